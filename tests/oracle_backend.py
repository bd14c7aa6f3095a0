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

    def dose_grid(self, actor, which, out=None):
        g = np.array(super().dose_grid(actor, which), dtype=np.float64, copy=True)
        if out is not None:
            out[...] = g
            return out
        return g
