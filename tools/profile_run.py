#!/usr/bin/env python
"""One warm-up launch + `--launches` timed launches of the transport kernel on a config; used
under ncu (profiles/) and for quick A/B timing on the GPU box."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

from bench import build_setup  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c3")
    ap.add_argument("--histories", type=int, default=5_000_000)
    ap.add_argument("--launches", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=1)
    a = ap.parse_args()
    from opengate_b200.core import Engine
    S = build_setup(a.config, a.histories)
    eng = Engine(0, S.seed)
    sids, _, _ = S.apply(eng)
    eng.begin_run()
    first = 0
    for _ in range(a.warmup):
        eng.timed_run(sids[0], first, a.histories)
        first += a.histories
    eng.begin_run()
    ms = []
    for _ in range(a.launches):
        ms.append(eng.timed_run(sids[0], first, a.histories))
        first += a.histories
    st = eng.stats()
    n = a.histories * a.launches
    print(json.dumps({"config": S.name, "histories_per_launch": a.histories, "ms": ms,
                      "primaries_per_s": n / (sum(ms) * 1e-3),
                      "tent_per_hist": st["tentative_collisions"] / n, "flush_per_hist": st["flushed_entries"] / n}))
    eng.close()


if __name__ == "__main__":
    main()
