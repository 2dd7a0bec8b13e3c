"""Transform sweep (BASELINE.json config 2): batched Chebyshev forward / backward transforms
Base(N, "CH").forward_fft / backward_fft on (N, batch) arrays, GB/s = 16 N batch / t against the
measured HBM copy peak, plus the CPU oracle (scipy pocketfft) on a bounded batch."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from pypde_b200 import Base  # noqa: E402
from oracle import pypde_port as P  # noqa: E402

peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0
out = []
for N in (64, 97, 128, 256, 512, 769, 1024, 2048, 3073, 4096):
    b = Base(N, "CH")
    batch = max(1000, min(1_000_000, (1 << 28) // N))       # up to ~2 GB per array, larger than L2
    x = torch.randn((N, batch), dtype=torch.float64, device="cuda")
    xt = x.T.contiguous()
    rec = {"N": N, "batch": batch, "algo": b.plan.algo}
    for name, fn, arr, axis in (("fwd_axis0", b.forward_fft, x, 0), ("bwd_axis0", b.backward_fft, x, 0),
                                ("fwd_axis1", b.forward_fft, xt, 1), ("bwd_axis1", b.backward_fft, xt, 1)):
        for _ in range(3):
            y = fn(arr, axis=axis)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for _ in range(reps):
            y = fn(arr, axis=axis)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        rec[name + "_gbs"] = 16.0 * N * batch / ms / 1e6
        rec[name + "_frac_hbm"] = rec[name + "_gbs"] / peak
    # round trip error and CPU (pocketfft, single thread) on a bounded batch
    back = b.backward_fft(b.forward_fft(x[:, :256].contiguous()))
    rec["roundtrip_rel"] = float(torch.linalg.norm(back - x[:, :256]) / torch.linalg.norm(x[:, :256]))
    cb = max(8, min(batch, 4_000_000 // N))
    xc = x[:, :cb].cpu().numpy()
    o = P.Basis(N, "CH")
    t0 = time.perf_counter()
    o.forward(xc)
    rec["cpu_fwd_gbs"] = 16.0 * N * cb / (time.perf_counter() - t0) / 1e9
    out.append(rec)
    print(json.dumps(rec))
    del x, xt
    torch.cuda.empty_cache()
os.makedirs("gpurun_out", exist_ok=True)
json.dump({"hbm_peak_gbs": peak, "algo": {"1": "dense DMMA", "2": "shared-memory FFT", "3": "Bluestein"}, "sweep": out},
          open("gpurun_out/dct_sweep.json", "w"), indent=1)
