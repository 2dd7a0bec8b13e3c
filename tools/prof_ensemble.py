"""Per-launch times of one eager batched-ensemble step (256 x 128^2 by default)."""
import json
import sys
import os
import numpy as np
import torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from pypde_b200 import _cabi
from pypde_b200.navier.ensemble import Ensemble

nm = int(sys.argv[1]) if len(sys.argv) > 1 else 256
kw = dict(case="rbc", shape=(128, 128), pr=1.0, dt=0.005, tsave=None, dealias=True, integrator="rk3", beta=1.0, aspect=1.0)
ens = Ensemble(np.logspace(4, 8, nm), **kw)
for m in ens.members:
    m.set_velocity(m=1, n=1, amplitude=0.2)
    m.set_temperature(amplitude=0.2)
for _ in range(3):
    ens.update()
torch.cuda.synchronize()
fs = ens.stepper
st = _cabi.stream()
rec = {}
for rep in range(3):
    for rk in range(3):
        lst = fs.stage_calls[rk]
        for (fn, args), name in zip(lst.calls, lst.labels):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _cabi.check(fn(*args, st))
            e1.record()
            rec.setdefault(name, []).append((e0, e1))
torch.cuda.synchronize()
out = {k: round(sum(a.elapsed_time(b) for a, b in v) / 3, 4) for k, v in rec.items()}
print(json.dumps({"members": nm, "ms_per_ensemble_step": out, "sum": sum(out.values())}, indent=1))
