"""Event time stamps of the GPU engine.

Reference: GateVSource::PrepareNextTime / CalcNextTime (GateVSource.cpp:55-89) and
GateGenericSource::PrepareNextTime (GateGenericSource.cpp:143-158): a source with `number_of_primaries` emits all
its events of a run at the start of the run interval (CalcNextTime returns the current time unchanged); an
`activity` source draws exponential inter-event times at its CURRENT activity (UpdateActivity: A0 exp(-lambda
(t - t_start))) until the interval ends -- a (non-homogeneous) Poisson process on the interval.

On the GPU the histories of a run are independent and carry no order, and the number of events of the interval has
already been drawn from the Poisson law of the same process (sources.primaries_for_run).  Given that number, the
event times of a Poisson process are independent draws from the normalised rate over the interval, so every
history gets its own: t(id) = F^-1(u(id)) with u(id) a 53-bit uniform taken from a Philox block of that history
that the transport never uses (block number 0xFFFFFFF0).  The time of a history is therefore a pure function of
(seed, history id, interval, source) like everything else about it: it is evaluated on the host for the rows a
PhaseSpaceActor stored (`GlobalTime` = event time + time of flight), not carried through the kernel."""

from __future__ import annotations

import numpy as np

TIME_BLOCK = 0xFFFFFFF0     # Philox block of a history reserved for its event time (the runaway guard of the
                            # transport stops a history at 4e6 blocks)


def philox4x32_10(seed, hist, block):
    """Philox4x32-10 of (key = seed, counter = (hist lo, hist hi, block, 0)) for arrays of history ids: the
    generator of csrc/device_math.cuh / oracle philox_block, vectorised.  Returns uint32 [n][4]."""
    hist = np.asarray(hist, dtype=np.uint64)
    c0 = (hist & np.uint64(0xFFFFFFFF)).astype(np.uint64)
    c1 = (hist >> np.uint64(32)).astype(np.uint64)
    c2 = np.full(hist.shape, int(block) & 0xFFFFFFFF, dtype=np.uint64)
    c3 = np.zeros(hist.shape, dtype=np.uint64)
    k0 = np.uint64(int(seed) & 0xFFFFFFFF)
    k1 = np.uint64((int(seed) >> 32) & 0xFFFFFFFF)
    M0, M1, MASK = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0xFFFFFFFF)
    W0, W1 = np.uint64(0x9E3779B9), np.uint64(0xBB67AE85)
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        n0 = ((p1 >> np.uint64(32)) ^ c1 ^ k0) & MASK
        n1 = p1 & MASK
        n2 = ((p0 >> np.uint64(32)) ^ c3 ^ k1) & MASK
        n3 = p0 & MASK
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + W0) & MASK
        k1 = (k1 + W1) & MASK
    return np.stack([c0, c1, c2, c3], axis=-1).astype(np.uint32)


def uniform53(seed, hist, block=TIME_BLOCK):
    """uniform in (0,1) with 53 random bits per history, centred in its bin"""
    w = philox4x32_10(seed, hist, block).astype(np.uint64)
    bits = ((w[..., 0] >> np.uint64(5)) << np.uint64(26)) | (w[..., 1] >> np.uint64(6))
    return (bits.astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)


def event_times(seed, event_ids, interval, start_time=None, end_time=None, activity=False, half_life=-1.0):
    """Event time (G4 units: ns) of every history id of one source in one run interval.
    number_of_primaries sources: the start of the effective interval for all events (GateVSource.cpp:55-62).
    activity sources: independent draws from the normalised activity over the effective interval
    [max(t0, start), min(t1, end)]: uniform, or ~ exp(-lambda (t - t_a)) with a half-life."""
    ids = np.asarray(event_ids, dtype=np.uint64)
    t0, t1 = float(interval[0]), float(interval[1])
    if start_time is not None:
        t0 = max(t0, float(start_time))
    if end_time is not None:
        t1 = min(t1, float(end_time))
    if not activity or not t1 > t0:
        return np.full(ids.shape, t0, dtype=np.float64)
    u = uniform53(seed, ids)
    if half_life is not None and half_life > 0:
        lam = np.log(2.0) / float(half_life)
        # F(t) = (1 - exp(-lam (t - t0))) / (1 - exp(-lam (t1 - t0)))
        span = -np.expm1(-lam * (t1 - t0))
        return t0 - np.log1p(-u * span) / lam
    return t0 + u * (t1 - t0)
