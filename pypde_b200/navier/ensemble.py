"""
Parameter-sweep ensemble (BASELINE.json configs[4]: 256 members, 128 x 128, Ra sweep): independent runs of one
grid, sharded over the GPUs with NO data-path communication ("replicas only").

batched=True (default on even grids): ONE launch list advances all members of this rank -- every axis pass
takes one job per member and array, the dense DCT / projection products run through the batched entry points,
and the whole ensemble step is one CUDA graph of 15 launches per stage (pass_stepper.PassStepper with a member
list).  batched=False: every member replays its own CUDA graph, round-robin on a few streams (round 1).
"""
import numpy as np
import torch

from .rbc2d import NavierStokes


class Ensemble:
    def __init__(self, ra_values, rank=0, world=1, streams=4, batched=None, **kwargs):
        from .pass_stepper import PassStepper
        ra_values = np.asarray(ra_values, dtype=float)
        self.indices = [i for i in range(ra_values.size) if i % world == rank]     # round-robin sharding
        self.members = [NavierStokes(ra=float(ra_values[i]), graph=False, **kwargs) for i in self.indices]
        ok = bool(self.members) and all(PassStepper.supported(m) and m._stepper_kind == "fast" for m in self.members)
        self.batched = ok if batched is None else (bool(batched) and ok)
        if not self.batched:
            for m in self.members:
                m._use_graph = True
        self.streams = [torch.cuda.Stream() for _ in range(max(1, min(streams, len(self.members))))]
        self.time = 0.0
        self._stepper = None
        self._graph = None

    def for_each(self, fn):
        for m in self.members:
            fn(m)

    @property
    def stepper(self):
        if self._stepper is None:
            from .pass_stepper import PassStepper
            self._stepper = PassStepper(self.members[0], self.members)
        return self._stepper

    def _run_eager(self):
        fs = self.stepper
        for rk in range(self.members[0].nstage):
            fs.stage(rk)

    def update(self):
        if not self.members:
            return
        if self.batched:
            fs = self.stepper
            if self._graph is None or fs._state_ptrs() != fs.bound:
                fs.bind()
                self._run_eager()              # warm-up outside the capture (plans, attributes); advances one step
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._run_eager()
                self._graph = g
            else:
                self._graph.replay()
            for m in self.members:
                m.ux, m.uz = None, None
        else:
            cur = torch.cuda.current_stream()
            for s in self.streams:
                s.wait_stream(cur)
            for k, m in enumerate(self.members):
                with torch.cuda.stream(self.streams[k % len(self.streams)]):
                    m.update()
            for s in self.streams:
                cur.wait_stream(s)
        self.time += self.members[0].dt

    def nusselt(self):
        import contextlib
        import io
        out = []
        for m in self.members:
            with contextlib.redirect_stdout(io.StringIO()):
                out.append(m.eval_Nu())
        return out
