"""GPU probe (run under gpurun): fp64 GEMM peak of the box (cuBLAS via torch.matmul, as
BASELINE.md asks) next to pypde_b200's DMMA kernel, and copy bandwidth."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from pypde_b200 import ops  # noqa: E402


def timeit(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) * 1e-3)
    return best


def main():
    out = {"gpu": torch.cuda.get_device_name(0)}
    dev = torch.device("cuda")
    for n in (2048, 4096, 8192):
        A = torch.randn((n, n), dtype=torch.float64, device=dev)
        B = torch.randn((n, n), dtype=torch.float64, device=dev)
        t = timeit(lambda: torch.matmul(A, B))
        out["cublas_fp64_tflops_%d" % n] = 2 * n ** 3 / t / 1e12
        t = timeit(lambda: ops.gemm(A, B))
        out["pde_gemm_nn_tflops_%d" % n] = 2 * n ** 3 / t / 1e12
        t = timeit(lambda: ops.gemm(A, B, transB=True))
        out["pde_gemm_nt_tflops_%d" % n] = 2 * n ** 3 / t / 1e12
        err = float(torch.linalg.norm(ops.gemm(A, B) - A @ B) / torch.linalg.norm(A @ B))
        out["pde_gemm_relerr_%d" % n] = err
    n = 2046
    A = torch.randn((n, n + 2), dtype=torch.float64, device=dev)
    Hm = torch.randn((n, n + 2), dtype=torch.float64, device=dev)
    t = timeit(lambda: ops.gemm(A, Hm, transB=True))
    out["pde_gemm_2046x2048x2046_tflops"] = 2 * n * n * (n + 2) / t / 1e12
    t = timeit(lambda: torch.matmul(A, Hm.T))
    out["cublas_2046x2048x2046_tflops"] = 2 * n * n * (n + 2) / t / 1e12
    x = torch.empty(1 << 28, dtype=torch.float64, device=dev)
    y = torch.empty_like(x)
    t = timeit(lambda: y.copy_(x))
    out["copy_gbs"] = 2 * x.numel() * 8 / t / 1e9
    print(json.dumps(out, indent=1))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/probe.json", "w"), indent=1)


if __name__ == "__main__":
    main()
