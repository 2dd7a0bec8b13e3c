"""
Parameter-sweep ensemble (BASELINE.json config 5: 256 members, 128 x 128, Ra sweep): independent
runs, sharded over the GPUs with NO data-path communication ("replicas only").
Every member is a NavierStokes object whose whole RK3 step is one CUDA graph; members are replayed
round-robin on a few streams so that the small kernels of different members overlap.
"""
import numpy as np
import torch

from .rbc2d import NavierStokes


class Ensemble:
    def __init__(self, ra_values, rank=0, world=1, streams=4, **kwargs):
        ra_values = np.asarray(ra_values, dtype=float)
        self.indices = [i for i in range(ra_values.size) if i % world == rank]     # round-robin sharding
        self.members = [NavierStokes(ra=float(ra_values[i]), graph=True, **kwargs) for i in self.indices]
        self.streams = [torch.cuda.Stream() for _ in range(max(1, min(streams, len(self.members))))]
        self.time = 0.0

    def for_each(self, fn):
        for m in self.members:
            fn(m)

    def update(self):
        cur = torch.cuda.current_stream()
        for s in self.streams:
            s.wait_stream(cur)
        for k, m in enumerate(self.members):
            with torch.cuda.stream(self.streams[k % len(self.streams)]):
                m.update()
        for s in self.streams:
            cur.wait_stream(s)
        if self.members:
            self.time += self.members[0].dt

    def nusselt(self):
        import contextlib
        import io
        out = []
        for m in self.members:
            with contextlib.redirect_stdout(io.StringIO()):
                out.append(m.eval_Nu())
        return out
