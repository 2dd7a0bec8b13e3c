#!/usr/bin/env python
"""
bench.py - headline benchmark of pypde_b200 (contract: see README / DESIGN.md §measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload rbc2048|rbc512|rbc64|ens128|diff1024|dct]
    python bench.py --impl reference ...      # the reference's CPU path (oracle port) on the host cores

Metric (BASELINE.json): RBC2D fp64 timesteps/sec at N x N.  A "step" is one full IMEX RK3 time
step (3 stages: transforms, nonlinear terms, Helmholtz/Poisson solves) of
navier.rbc2d.NavierStokes on synthetic initial fields.  One JSON line is printed by rank 0.

  value      steps/s, state resident in HBM, CUDA-event timed, max over ranks
  e2e        steps/s through the reference-facing API with HOST buffers: every step uploads the
             state (T, U, V, pres coefficients) from pinned host memory, steps, and reads it back
  roofline   dominant kernel of the step, timed live with CUDA events in an instrumented step
  cpu_baseline  the CPU oracle (bit-identical port of the reference) on a bounded sample
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (NavierStokes kwargs, description)
    "rbc2048": dict(case="rbc", shape=(2048, 2048), ra=1e10, pr=1.0, dt=1e-4, tsave=None, dealias=True,
                    integrator="rk3", beta=1.0, aspect=1.0),
    "rbc512": dict(case="rbc", shape=(512, 512), ra=1e8, pr=1.0, dt=1e-3, tsave=None, dealias=True,
                   integrator="rk3", beta=1.0, aspect=1.0),
    "rbc64": dict(case="rbc", shape=(64, 64), ra=1e5, pr=1.0, dt=0.01, tsave=None, dealias=True,
                  integrator="rk3", beta=1.0, aspect=1.0),
}
ENSEMBLE = {"members": 256, "cfg": dict(case="rbc", shape=(128, 128), pr=1.0, dt=0.005, tsave=None, dealias=True,
                                        integrator="rk3", beta=1.0, aspect=1.0)}
METRIC = "rbc2d_fp64_timesteps_per_sec"
UNIT = "steps/s"


def base_config(args, cfg):
    """The keys both arms (--impl b200 / reference) report under "config"."""
    return {"workload": args.workload, "shape": list(cfg["shape"]), "integrator": cfg["integrator"],
            "dealias": cfg["dealias"], "ra": cfg["ra"], "dt": cfg["dt"], "stages_per_step": 3 if cfg["integrator"] == "rk3" else 1}


def init_state(ns, shape, port=False):
    """Synthetic initial fields of SURVEY.md §8(d).1 (same seed as the golden fixtures)."""
    ns.set_velocity(m=1, n=1, amplitude=0.2)
    ns.set_temperature(amplitude=0.2)
    k0, k1 = min(16, shape[0] - 2), min(16, shape[1] - 2)
    pert = 1e-3 * np.random.default_rng(0).standard_normal((k0, k1))
    if port:
        ns.That_[:k0, :k1] += pert
    else:
        import torch
        ns.T.vhat[:k0, :k1] += torch.as_tensor(pert, device=ns.T.vhat.device)


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi sampled every 200 ms during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------- CPU arms
def cpu_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count()


def oracle_sample(cfg, full_steps):
    """Times `full_steps` complete time steps of the CPU oracle (bit-identical port of the reference's
    NumPy/SciPy/Fortran path) after its setup; the sample is measured, not extrapolated."""
    from oracle import pypde_port as P
    try:        # torchrun / the image may pin BLAS to one thread: the two dense products get all host cores
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=cpu_cores())
    except Exception:
        pass
    t0 = time.perf_counter()
    o = P.RBC2D(**cfg)
    init_state(o, cfg["shape"], port=True)
    setup = time.perf_counter() - t0
    t0 = time.perf_counter()
    for _ in range(full_steps):
        o.update()
    dt = (time.perf_counter() - t0) / full_steps
    return 1.0 / dt, setup, "%d full %s step(s) of %dx%d, %.2f s each (no warm-up step)" % (
        full_steps, cfg["integrator"].upper(), *cfg["shape"], dt)


def run_reference(args, cfg):
    """--impl reference: the reference's CPU path.  /root/reference does not exist on the GPU box
    and its Fortran cannot be built (no gfortran): the oracle port stands in (kind = "port").
    At least 3 timed steps (2048^2: ~55 s each on 16 cores) after one warm-up step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    budget = float(os.environ.get("PDE_BENCH_CPU_BUDGET_S", "300"))
    try:        # torchrun exports OMP_NUM_THREADS=1: give the two dense BLAS products all host cores back
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=cpu_cores())
    except Exception:
        pass
    from oracle import pypde_port as P
    o = P.RBC2D(**cfg)
    init_state(o, cfg["shape"], port=True)
    t0 = time.perf_counter()
    o.update()                       # warm-up + cost probe
    probe = time.perf_counter() - t0
    warm = 1
    steps = max(min(3, args.steps), min(args.steps, int(budget / max(probe, 1e-9)) - 1))
    while warm < args.warmup and probe * (warm + 1 + steps) < budget:
        o.update()
        warm += 1
    t0 = time.perf_counter()
    for _ in range(steps):
        o.update()
    dt = (time.perf_counter() - t0) / steps
    val = 1.0 / dt
    sample = "%d full step(s) of %dx%d after %d warm-up (requested %d/%d, clamped to a %.0f s budget, never below 3 timed steps)" % (
        steps, *cfg["shape"], warm, args.steps, args.warmup, budget)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": base_config(args, cfg),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cpu_cores(), "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------- GPU arm
NCU_STAGE_CSV = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r02_ncu_stage.csv")


def ncu_traffic(kernel_substr, path=NCU_STAGE_CSV):
    """(average dram__bytes_read + dram__bytes_write per launch in bytes, source) of the kernels whose name contains
    `kernel_substr` in the committed ncu --set full capture of ONE rbc2048 stage (tools/final_round2.sh writes it)."""
    import csv
    try:
        rows = [r for r in csv.reader(l for l in open(path) if not l.startswith('"#')) if r]
        hdr, units, body = rows[0], rows[1], rows[2:]
        ir, iw, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        tot = [float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]] for r in body if kernel_substr in r[ik]]
        if not tot:
            return None, None
        return sum(tot) / len(tot), "profiles/%s (%d launches of one stage)" % (os.path.basename(path), len(tot))
    except (OSError, ValueError, KeyError, IndexError):
        return None, None


class OpTimer:
    """CUDA-event timer around every C-ABI call of one (eager) time step: wraps the prebuilt launch
    lists of the stepper.  Events are recorded on the stream the kernels are launched on."""

    def __init__(self):
        self.records = []

    def run_step(self, ns):
        import torch
        from pypde_b200 import _cabi
        fs = ns._fast
        st = _cabi.stream()
        for rk in range(ns.nstage):
            lst = fs.stage_calls[rk]
            labels = getattr(lst, "labels", None) or [getattr(fn, "__name__", "call") for fn, _ in lst.calls]
            for (fn, args), name in zip(lst.calls, labels):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                _cabi.check(fn(*args, st))
                e1.record()
                self.records.append((name, getattr(fn, "__name__", "call"), args, e0, e1))
        torch.cuda.synchronize()

    def summary(self):
        out = {}
        for name, fname, a, e0, e1 in self.records:
            key = name
            work = None
            if fname == "pde_sweep":
                key = "pde_sweep[%s,axis%d]" % (["diff", "tdma_fwd", "tdma_bwd", "fdma_fwd", "fdma_bwd", "twodma"][a[0]], a[1])
            if fname == "pde_dct1_multi":
                key = "pde_dct1_multi[axis%d]" % a[10]
                # algorithmic bytes: every input element read once (n_in), every output element written once (n_out)
                work = 8.0 * a[2] * (a[5] + a[8]) * a[9]
            if fname == "pde_gemm_f64":
                work = 2.0 * a[7] * a[8] * a[9]
            d = out.setdefault(key, {"ms": 0.0, "launches": 0, "work": 0.0})
            d["ms"] += e0.elapsed_time(e1)
            d["launches"] += 1
            if work:
                d["work"] += work
        return out


def fp64_peak_tflops():
    """cuBLAS DGEMM 8192^3 (MEASURED_PEAKS.json has no fp64 entry, measure it)."""
    import torch
    n = 8192
    a = torch.randn((n, n), dtype=torch.float64, device="cuda")
    b = torch.randn((n, n), dtype=torch.float64, device="cuda")
    best = 1e9
    for i in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        if i:
            best = min(best, e0.elapsed_time(e1))
    del a, b
    torch.cuda.empty_cache()
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "MEASURED_PEAKS.json"
    return {"hbm_gbs": 6650.0}, "fallback of B200_PROFILING.md"


def run_ensemble(args):
    """--workload ens128: 256 independent 128 x 128 runs (Ra = logspace(4, 8, 256)) sharded over the GPUs,
    no data-path collective; a step advances EVERY member by one RK3 step."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        init_nccl(local)
    from pypde_b200 import _cabi
    from pypde_b200.navier.ensemble import Ensemble
    ra = np.logspace(4, 8, ENSEMBLE["members"])
    t0 = time.perf_counter()
    ens = Ensemble(ra, rank=rank, world=world, **ENSEMBLE["cfg"])
    ens.for_each(lambda m: init_state(m, ENSEMBLE["cfg"]["shape"]))
    setup_s = time.perf_counter() - t0

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        ens.update()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        ens.update()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clocks = sampler.stop()
    # ---- end to end: every member's state uploaded from pinned host memory before and read back after each step ----
    state = [t for m in ens.members for t in (m.T.vhat, m.U.vhat, m.V.vhat, m.pres.vhat)]
    host_a = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True).copy_(t) for t in state]
    host_b = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True) for t in state]
    h2d = d2h = sum(h.numel() * 8 for h in host_a)
    e2e_steps = max(1, min(args.steps, 5))
    barrier()
    e0.record()
    for _ in range(e2e_steps):
        for h, t in zip(host_a, state):
            t.copy_(h, non_blocking=True)
        ens.update()
        for h, t in zip(host_b, state):
            h.copy_(t, non_blocking=True)
        torch.cuda.synchronize()
        host_a, host_b = host_b, host_a
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1) / e2e_steps
    if world > 1:
        t = torch.tensor([e2e_ms, float(h2d), float(d2h)], dtype=torch.float64, device="cuda")
        tm = t.clone()
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        dist.all_reduce(t)
        e2e_ms, h2d, d2h = float(tm[0].item()), int(t[1].item()), int(t[2].item())
    finite = all(bool(torch.isfinite(m.T.vhat).all()) for m in ens.members)
    # whole step against the SURVEY 8(d) axis-pass model (all members): small grids, dense-matrix transforms
    peaks, peak_src = measured_peaks()
    N, D, M = 128, 193, 126
    stage_bytes = 8 * (31 * M * M + 18 * M * N + 5 * N * N + 12 * D * M + 6 * D * N + D * D)
    alg = stage_bytes * 3 * ENSEMBLE["members"]
    achieved = alg / (ms / args.steps * 1e-3) / 1e9
    roof = {"kernel": "whole ensemble step (45 launches: k_pass, batched k_gemm_f64 for the dense DCTs / projections, "
                      "k_conv_products_members)", "bound": "hbm", "achieved": achieved,
            "peak": peaks.get("hbm_gbs") * world, "unit": "GB/s", "frac": achieved / (peaks.get("hbm_gbs") * world),
            "traffic": None, "peak_source": peak_src,
            "note": "algorithmic bytes of the SURVEY.md 8(d) axis-pass model x 256 members; per-kernel view: tools/prof_ensemble.py"}
    # launches: one eager member step counted through the library, times members and steps
    m0 = ens.members[0]
    _cabi.launch_count_reset()
    if ens.batched:
        ens._run_eager()                # one eager ensemble step: every launch carries all members
        torch.cuda.synchronize()
        launches_step = _cabi.launch_count()
        per_member = launches_step / max(1, len(ens.members))
    else:
        for rk in range(m0.nstage):
            m0._fast.stage_calls[rk].run()
        torch.cuda.synchronize()
        per_member = _cabi.launch_count()
        launches_step = per_member * len(ens.members)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import pypde_port as P
        cfgm = dict(ENSEMBLE["cfg"], ra=float(ra[128]))
        o = P.RBC2D(**cfgm)
        init_state(o, cfgm["shape"], port=True)
        o.update()
        t0 = time.perf_counter()
        for _ in range(3):
            o.update()
        tm = (time.perf_counter() - t0) / 3
        cpu = {"value": 1.0 / (tm * ENSEMBLE["members"]), "unit": UNIT, "cores": cpu_cores(), "kind": "port",
               "sample": "3 RK3 steps of ONE 128x128 member (%.3f s each) x 256 members" % tm}
    line = {"metric": METRIC, "value": args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "ens128", "members": ENSEMBLE["members"], "members_per_gpu": len(ens.members),
                       "shape": [128, 128], "ra": "logspace(4, 8, 256)", "integrator": "rk3", "dealias": True,
                       "parallelism": "independent members sharded over %d GPU(s), no collective" % world,
                       "l2": "one member's working set fits L2 (small-grid regime by design)", "finite": finite,
                       "setup_s": setup_s, "cuda_graph": True},
            "member_steps_per_sec": args.steps * ENSEMBLE["members"] / (ms * 1e-3),
            "e2e": {"value": 1e3 / e2e_ms, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps},
            "roofline": roof,
            "clocks": clocks, "gpu_launches": int(launches_step * args.steps),
            "gpu_launches_per_ensemble_step": int(launches_step), "batched": bool(ens.batched),
            "gpu_launches_per_member_step": per_member, "cpu_baseline": cpu}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def quiet_nccl():
    """The image exports NCCL_DEBUG=VERSION, which prints a banner on stdout next to the JSON line."""
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"


def init_nccl(local):
    """init_process_group + the first collective with file descriptor 1 pointed at stderr: whatever the NCCL
    library prints while it creates the communicator (the version banner arrived on stdout even with
    NCCL_DEBUG=WARN set here) must not land next to the ONE JSON line rank 0 prints."""
    import torch
    import torch.distributed as dist
    quiet_nccl()
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier()
        torch.cuda.synchronize()
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)


def run_gpu(args, cfg):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    slab = world > 1
    if slab:
        init_nccl(local)
        # Capturing the NCCL transposes in the CUDA graph works (rbc512 x2: 2.74 vs 3.03 ms/step) but the
        # process then hung in teardown on this stack, so the multi-GPU step is launched eagerly.
        if os.environ.get("PDE_SLAB_GRAPH", "1") != "1":
            args.no_graph = True
    from pypde_b200 import _cabi
    from pypde_b200.navier import rbc2d

    def barrier():
        torch.cuda.synchronize()
        if slab:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(ms):
        if not slab:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    t0 = time.perf_counter()
    ns = rbc2d.NavierStokes(graph=not args.no_graph, slab=slab, **cfg)
    init_state(ns, cfg["shape"])
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t0

    for _ in range(args.warmup):
        ns.update()
        ns.update_time()
    barrier()
    # tensors holding the state that an end-to-end caller moves every step
    state_tensors = ns._fast.local_state() if slab else [f.vhat for f in (ns.T, ns.U, ns.V, ns.pres)]

    # ---- device-resident timing -------------------------------------------------------
    sampler = ClockSampler(local)
    sampler.start()
    _cabi.launch_count_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        ns.update()
        ns.update_time()
    e1.record()
    barrier()
    launches = _cabi.launch_count()
    ms = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop()
    ns.sync_fields()
    finite = bool(torch.isfinite(ns.T.vhat).all())

    # ---- end-to-end: host buffers in, host buffers out, every step -----------------------
    host_in = [torch.empty(t.shape, dtype=torch.float64).pin_memory() for t in state_tensors]
    host_out = [torch.empty(t.shape, dtype=torch.float64).pin_memory() for t in state_tensors]
    for h, t in zip(host_in, state_tensors):
        h.copy_(t)
    h2d = sum(h.numel() * 8 for h in host_in)
    d2h = sum(h.numel() * 8 for h in host_out)
    if slab:
        tt = torch.tensor([h2d, d2h], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt)
        h2d, d2h = int(tt[0].item()), int(tt[1].item())
    e2e_steps = max(1, min(args.steps, 5))
    barrier()
    e0.record()
    for _ in range(e2e_steps):
        for h, t in zip(host_in, state_tensors):
            t.copy_(h, non_blocking=True)
        ns.update()
        ns.update_time()
        for h, t in zip(host_out, state_tensors):
            h.copy_(t, non_blocking=True)
        torch.cuda.synchronize()
        host_in, host_out = host_out, host_in      # the next step's input is this step's host result
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1)) / e2e_steps

    # ---- instrumented step: per-kernel CUDA-event times, dominant kernel's roofline ----
    timer = OpTimer()
    _cabi.launch_count_reset()
    timer.run_step(ns)
    launches_per_step = _cabi.launch_count()
    if not args.no_graph:
        # graph replays do not pass the library's launch counter: the captured step holds exactly the
        # kernels of one eager step, counted here
        launches = launches_per_step * args.steps
    ops_ms = timer.summary()
    step_ms_instr = sum(v["ms"] for v in ops_ms.values())
    peaks, peak_src = measured_peaks()
    fp64_peak = fp64_peak_tflops()
    dct_ms = sum(v["ms"] for k, v in ops_ms.items() if k.startswith("pde_dct1"))
    dct_work = sum(v["work"] for k, v in ops_ms.items() if k.startswith("pde_dct1"))
    dct_launches = sum(v["launches"] for k, v in ops_ms.items() if k.startswith("pde_dct1"))
    gemm = ops_ms.get("pde_gemm_f64", {"ms": 0.0, "work": 0.0, "launches": 1})
    roof_dct = {"kernel": "k_dct_row_tma + k_dct_fft_t (shared-memory FFT DCT-I: persistent TMA-staged rows, 4-column strips; all batched transforms of the step)", "bound": "hbm",
                "achieved": dct_work / (dct_ms * 1e-3) / 1e9 if dct_ms else None, "peak": peaks.get("hbm_gbs"),
                "unit": "GB/s",
                "traffic": None,
                "traffic_note": "algorithmic bytes = 8 (n_in + n_out) batch per array: every input element read once, "
                                "every output element written once",
                "peak_source": peak_src, "launches_per_step": dct_launches,
                "avg_launch_ms": dct_ms / max(dct_launches, 1), "share_of_step": dct_ms / step_ms_instr,
                "algorithmic_bytes_per_launch": dct_work / max(dct_launches, 1)}
    roof_dct["frac"] = roof_dct["achieved"] / roof_dct["peak"] if roof_dct["achieved"] else None
    if not slab and tuple(cfg["shape"]) == (2048, 2048):
        # DRAM bytes per launch of the same kernels from the committed ncu --set full capture of one stage
        roof_dct["traffic"], roof_dct["traffic_source"] = ncu_traffic("k_dct")
        roof_gemm_traffic = ncu_traffic("k_gemm")
    else:
        roof_gemm_traffic = (None, None)
    roof_gemm = {"kernel": "k_gemm_f64 (DMMA; Poisson projections Hy, Qy)", "bound": "tensor",
                 "achieved": gemm["work"] / (gemm["ms"] * 1e-3) / 1e12 if gemm["ms"] else None, "peak": fp64_peak,
                 "unit": "TFLOP/s", "traffic": None,
                 "peak_source": "cuBLAS DGEMM 8192^3 measured in this run (MEASURED_PEAKS.json has no fp64 entry)",
                 "launches_per_step": gemm["launches"], "avg_launch_ms": gemm["ms"] / max(gemm["launches"], 1),
                 "share_of_step": gemm["ms"] / step_ms_instr}
    roof_gemm["frac"] = roof_gemm["achieved"] / fp64_peak if roof_gemm["achieved"] else None
    if roof_gemm_traffic[0]:
        roof_gemm["traffic"], roof_gemm["traffic_source"] = roof_gemm_traffic
    roof = roof_dct if dct_ms >= gemm["ms"] else roof_gemm
    roof_other = roof_gemm if roof is roof_dct else roof_dct

    # ---- multi-GPU: the slab-decomposed state against the single-GPU stepper, outside the timed region ----
    parity = None
    if slab:
        psteps = 2
        chk = rbc2d.NavierStokes(graph=False, slab=True, **cfg)
        init_state(chk, cfg["shape"])
        for _ in range(psteps):
            chk.update()
        chk.sync_fields()
        torch.cuda.synchronize()
        if hasattr(chk._fast, "check"):
            chk._fast.check()
        if rank == 0:
            one = rbc2d.NavierStokes(graph=False, slab=False, **cfg)
            init_state(one, cfg["shape"])
            for _ in range(psteps):
                one.update()
            torch.cuda.synchronize()
            errs = {}
            for k, a, b in (("T", chk.T.vhat, one.T.vhat), ("U", chk.U.vhat, one.U.vhat), ("V", chk.V.vhat, one.V.vhat),
                            ("pres", chk.pres.vhat, one.pres.vhat)):
                errs[k] = float(torch.linalg.norm(a - b) / torch.linalg.norm(b))
            parity = {"steps": psteps, "rel_l2": errs, "max": max(errs.values()),
                      "note": "slab-decomposed state (gathered) vs the single-GPU stepper on rank 0, same initial state; the "
                              "reference's own step moves by 3e-11 (U, V, pres) at 2048^2 when its input changes by half an ulp "
                              "(tests/test_gpu_large.py), so two roundings of the same step agree to that level"}
            del one
        chk.close()
        del chk

    # ---- CPU baseline: bounded sample of the same workload on rank 0 -----------------------
    cpu = None
    if not args.no_cpu_baseline and not slab and rank == 0:
        big = cfg["shape"][0] * cfg["shape"][1] > 600 * 600
        v, cpu_setup, sample = oracle_sample(cfg, 1 if big else 3)
        cpu = {"value": v, "unit": UNIT, "cores": cpu_cores(), "kind": "port", "sample": sample,
               "setup_s": cpu_setup}

    N = cfg["shape"][0]
    D = int(N * 1.5)
    M = N - 2
    stage_bytes = 8 * (31 * M * M + 18 * M * N + 5 * N * N + 12 * D * M + 6 * D * N + D * D)
    stage_flop = 2.0 * M * N * M + 2.0 * M ** 3
    nst = ns.nstage
    hbm = peaks.get("hbm_gbs", 6650.0) * 1e9 * world            # N GPUs: N x the peaks
    fl = fp64_peak * 1e12 * world
    t_step = ms / args.steps * 1e-3
    kern = {k: round(v["ms"], 4) for k, v in sorted(ops_ms.items(), key=lambda kv: -kv[1]["ms"])}
    config = base_config(args, cfg)
    config.update({
        "parallelism": ("slab x%d: peer-memory row passes (loads / stores over NVLink), device-side barriers, no NCCL on the "
                        "data path" % world) if slab else "single GPU",
        "stepper": type(ns._fast).__name__,
        "l2": "state + work arrays exceed the 126 MB L2 (no flush needed)" if N >= 1024
        else "working set fits L2 (small-grid regime, by design of the workload)",
        "finite": finite, "setup_s": setup_s, "cuda_graph": not args.no_graph,
        "dealias_grid": list(ns.deriv_field.dealias.shape_physical) if cfg["dealias"] else None})
    line = {
        "metric": METRIC, "value": args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config,
        "clocks": clocks,
        "e2e": {"value": 1e3 / e2e_ms, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps},
        "gpu_launches": launches,
        "gpu_launches_per_step": launches_per_step,
        "roofline": roof,
        "roofline_second": roof_other,
        "step_model": {"algorithmic_bytes_per_step": stage_bytes * nst, "dense_flop_per_step": stage_flop * nst,
                       "hbm_fraction": (stage_bytes * nst / hbm) / t_step,
                       "combined_fraction": (stage_bytes * nst / hbm + stage_flop * nst / fl) / t_step,
                       "fp64_peak_tflops_per_gpu": fp64_peak, "hbm_peak_gbs_per_gpu": peaks.get("hbm_gbs"), "n_gpus": world,
                       "note": "SURVEY.md 8(d) axis-pass model; peaks are per GPU x n_gpus"},
        "kernel_ms_per_step": kern,
        "cpu_baseline": cpu,
    }
    if slab:
        # exchange = the kernels that exist only because the grid is distributed: the two-instruction row passes that move the
        # inputs / outputs of the transforms and projections, and the device-side barriers; the other transposes are the
        # loads / stores of row passes that also compute (PY2, PY4, PY7)
        tms = sum(v for k, v in kern.items() if k.startswith("exchange") or k == "peer_barrier" or k.startswith("slab_transpose"))
        line["transpose_ms_per_step"] = tms
        line["transpose_note"] = "exchange-only passes + device barriers of rank 0 (eager, event-timed); the pulls / pushes fused " \
                                 "into the computing row passes are inside pass[PY*]"
        line["parity_vs_single_gpu"] = parity["max"] if parity else None
        line["parity_detail"] = parity
    if rank == 0:
        print(json.dumps(line))
    if slab:
        ns.close()
        dist.destroy_process_group()


def run_diffusion(args):
    """--workload diff1024 (BASELINE.json configs[2]): theta-scheme steps of the 2-D diffusion example
    with a Dirichlet wall at 1024 x 1024 through the generic Field / grad / SolverPlan API (no fused
    stepper: this is the plan-by-plan path a user script takes).  Single GPU."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "diffusion"))
    from diff_2d_bc import Diffusion2d
    from pypde_b200 import _cabi
    torch.cuda.set_device(0)
    cfg = dict(shape=(1024, 1024), dt=0.01, kappa=0.1, beta=0.5)
    D = Diffusion2d(tsave=None, **cfg)
    for _ in range(args.warmup):
        D.update()
    torch.cuda.synchronize()
    _cabi.launch_count_reset()
    sampler = ClockSampler(0)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        D.update()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = _cabi.launch_count()
    clocks = sampler.stop()
    # end to end: the coefficient array comes from / returns to pinned host memory every step
    hin = D.field.vhat.cpu().pin_memory()
    hout = torch.empty_like(hin).pin_memory()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        D.field.vhat.copy_(hin, non_blocking=True)
        D.update()
        hout.copy_(D.field.vhat, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        hin, hout = hout, hin
    e1.record()
    torch.cuda.synchronize()
    ms_e2e = e0.elapsed_time(e1)
    nb = hin.numel() * 8
    finite = bool(torch.isfinite(D.field.vhat).all())
    # the same step replayed from a CUDA graph (removes the Python / launch overhead of the generic path)
    static = D.field.vhat
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        D.update()
        static.copy_(D.field.vhat)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        D.update()
        static.copy_(D.field.vhat)
    D.field.vhat = static
    for _ in range(args.warmup):
        graph.replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    ms_graph = e0.elapsed_time(e1)
    cpu = None
    if not args.no_cpu_baseline:
        from oracle import pypde_port as P
        o = P.Diffusion2D(**cfg)
        o.update()
        t0 = time.perf_counter()
        for _ in range(20):
            o.update()
        tm = (time.perf_counter() - t0) / 20
        cpu = {"value": 1.0 / tm, "unit": UNIT, "cores": cpu_cores(), "kind": "port",
               "sample": "20 steps of the oracle Diffusion2D at 1024x1024 (%.3f s each)" % tm}
    # algorithmic traffic of one step: 2 mixed second derivatives (stencil + two recurrences each: 3 passes),
    # 2 axpy, solve_rhs (2 banded products), solve_old (2), 1 add, solve_lhs (2 sweeps): 16 passes, read + write
    N = 1024
    bytes_step = 16 * 2 * 8 * N * N
    peaks, peak_src = measured_peaks()
    achieved = bytes_step * args.steps / (ms * 1e-3) / 1e9
    line = {"metric": METRIC, "value": args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "diff1024", "shape": [N, N], "bases": ["CD", "CN"], "dt": 0.01, "kappa": 0.1,
                       "beta": 0.5, "path": "generic plan-by-plan API (grad, SolverPlan), eager launches",
                       "l2": "arrays are 8 MB each: the step's working set fits the 126 MB L2 (by the size of the "
                             "config)", "finite": finite},
            "clocks": clocks,
            "e2e": {"value": args.steps / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": nb,
                    "d2h_bytes_per_step": nb},
            "gpu_launches": launches, "graph_ms_per_step": ms_graph / args.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": achieved / peaks["hbm_gbs"], "traffic": None, "peak_source": peak_src,
                         "note": "whole step, 16 axis passes x (read + write) x 8 B x N^2 = %d B" % bytes_step},
            "cpu_baseline": cpu}
    print(json.dumps(line))


def run_dct(args):
    """--workload dct (BASELINE.json configs[1]): batched Chebyshev forward / backward transforms
    Base(N, "CH").forward_fft / backward_fft on (N, batch) arrays, N = 64 ... 4096 (and the FFT-friendly dealias
    lengths 769 / 3073 the time step uses), both axes; GB/s = 8 (n_in + n_out) batch / t = 16 N batch / t.
    value = median over the sweep.  A "step" is one pass over the whole sweep.  Single GPU."""
    import torch
    from pypde_b200 import Base, _cabi
    from oracle import pypde_port as P
    torch.cuda.set_device(0)
    peaks, peak_src = measured_peaks()
    sizes = (64, 128, 256, 512, 769, 1024, 2048, 3073, 4096)
    algo_name = {1: "dense DMMA matrix", 2: "shared-memory FFT", 3: "Bluestein"}
    cases = []
    for N in sizes:
        b = Base(N, "CH")
        batch = max(1000, min(1_000_000, (1 << 27) // N))       # ~1 GB per array: larger than L2
        x = torch.randn((N, batch), dtype=torch.float64, device="cuda")
        # rows with an even pitch (16-byte aligned rows, what the stepper allocates): odd N gets one pad column
        xt = torch.zeros((batch, N + (N & 1)), dtype=torch.float64, device="cuda")[:, :N]
        xt.copy_(x.T)
        for name, fn, arr, axis in (("fwd0", b.forward_fft, x, 0), ("bwd0", b.backward_fft, x, 0),
                                    ("fwd1", b.forward_fft, xt, 1), ("bwd1", b.backward_fft, xt, 1)):
            cases.append((N, batch, b.plan.algo, name, fn, arr, axis))
    sampler = ClockSampler(0)
    for _ in range(args.warmup):
        for N, batch, algo, name, fn, arr, axis in cases:
            fn(arr, axis=axis)
    torch.cuda.synchronize()
    sampler.start()
    _cabi.launch_count_reset()
    times = {}
    t_all0, t_all1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_all0.record()
    evs = []
    for _ in range(args.steps):
        for k, (N, batch, algo, name, fn, arr, axis) in enumerate(cases):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn(arr, axis=axis)
            e1.record()
            evs.append((k, e0, e1))
    t_all1.record()
    torch.cuda.synchronize()
    launches = _cabi.launch_count()
    clocks = sampler.stop()
    for k, e0, e1 in evs:
        times.setdefault(k, []).append(e0.elapsed_time(e1))
    sweep, gbs = [], []
    for k, (N, batch, algo, name, fn, arr, axis) in enumerate(cases):
        ms = float(np.median(times[k]))
        g = 16.0 * N * batch / ms / 1e6
        gbs.append(g)
        sweep.append({"N": N, "batch": batch, "algo": algo_name.get(algo, str(algo)), "case": name, "ms": round(ms, 4),
                      "GBps": round(g, 1), "frac_hbm": round(g / peaks["hbm_gbs"], 3)})
    ms_total = t_all0.elapsed_time(t_all1)
    # end to end on one size: host array in, host coefficients out
    N, batch = 1024, 65536
    b = Base(N, "CH")
    hin = torch.randn((N, batch), dtype=torch.float64).pin_memory()
    hout = torch.empty((N, batch), dtype=torch.float64).pin_memory()
    dev = torch.empty((N, batch), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        dev.copy_(hin, non_blocking=True)
        y = b.forward_fft(dev, axis=0)
        hout.copy_(y, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    e1.record()
    torch.cuda.synchronize()
    e2e_gbs = 3 * 16.0 * N * batch / e0.elapsed_time(e1) / 1e6
    cpu = None
    if not args.no_cpu_baseline:
        o = P.Basis(2048, "CH")
        xc = np.random.default_rng(0).standard_normal((2048, 2000))
        o.forward(xc[:, :50])
        t0 = time.perf_counter()
        o.forward(xc)
        tc = time.perf_counter() - t0
        cpu = {"value": 16.0 * 2048 * 2000 / tc / 1e9, "unit": "GB/s", "cores": 1, "kind": "port",
               "sample": "scipy pocketfft dctn(type=1) + scaling of the oracle Basis(2048).forward on 2000 columns (%.2f s)" % tc}
    best = max(sweep, key=lambda r: r["GBps"])
    val = float(np.median(gbs))
    line = {"metric": "dct1_fp64_GBps", "value": val, "unit": "GB/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "dct", "sizes": list(sizes), "cases": "forward / backward x axis 0 / axis 1", "value_is":
                       "median GB/s over the sweep (16 bytes per element and transform)",
                       "l2": "every array is ~1 GB (larger than the 126 MB L2)"},
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": e2e_gbs, "unit": "GB/s", "h2d_bytes_per_step": 8 * N * batch, "d2h_bytes_per_step": 8 * N * batch,
                    "note": "Base(1024, CH).forward_fft on a pinned host (1024 x 65536) array, copies inside the timed region"},
            "roofline": {"kernel": "batched DCT-I (best case of the sweep: N = %d %s, %s)" % (best["N"], best["case"], best["algo"]),
                         "bound": "hbm", "achieved": best["GBps"], "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": best["GBps"] / peaks["hbm_gbs"], "traffic": None, "peak_source": peak_src},
            "sweep": sweep, "cpu_baseline": cpu}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="rbc2048", choices=sorted(WORKLOADS) + ["ens128", "diff1024", "dct"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch eagerly instead of replaying a CUDA graph")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.workload == "ens128":
        if args.impl == "reference":
            raise SystemExit("--impl reference times the rbc workloads; the ensemble's CPU number is its cpu_baseline")
        return run_ensemble(args)
    if args.workload == "dct":
        if args.impl == "reference":
            raise SystemExit("--impl reference times the rbc workloads; the transform sweep's CPU number is its cpu_baseline")
        return run_dct(args)
    if args.workload == "diff1024":
        if args.impl == "reference":
            raise SystemExit("--impl reference times the rbc workloads; diff1024's CPU number is its cpu_baseline")
        return run_diffusion(args)
    cfg = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_gpu(args, cfg)


if __name__ == "__main__":
    main()
