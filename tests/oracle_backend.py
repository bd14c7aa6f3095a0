"""Test infrastructure: lets `Simulation.run(_backend=...)` drive the CPU oracle instead of the CUDA
engine, so the HOST logic of the drop-in layer (volumes, sources, actors, engines) is testable
without a GPU.  Never used by the product."""
import numpy as np

from oracle.oracle import TransportOracle


class OracleBackend(TransportOracle):
    def __init__(self, device, seed):
        super().__init__(seed)
        self.device = device

    def begin_run(self, run_id=0):
        super().begin_run()

    def run_source(self, source_id, first, n):
        super().run_source(source_id, first, n, 0)

    def sync(self):
        pass

    def dose_grid(self, actor, which, out=None, scale=1.0):
        g = np.array(super().dose_grid(actor, which), dtype=np.float64, copy=True)
        if scale != 1.0:
            g *= scale
        if out is not None:
            out[...] = g
            return out
        return g

    def uncertainty_grid(self, actor, which, n_events, out=None):
        """what gate_b200_get_uncertainty evaluates on the device, in numpy (reference algebra utility.py:551-560,
        dataitems.py:1956-1993)"""
        from opengate_b200.actors import uncertainty_from_squares
        v = np.array(TransportOracle.dose_grid(self, actor, which), dtype=np.float64)
        sq = np.array(TransportOracle.dose_grid(self, actor, which + 1), dtype=np.float64)
        u = uncertainty_from_squares(v, sq, n_events)
        if out is not None:
            out[...] = u
            return out
        return u
