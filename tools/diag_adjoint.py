"""prints the per-step relative errors of the device adjoint iteration against the CPU oracle"""
import os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "tests")))
from test_gpu_adjoint import run_pair
from test_oracle_cpu import ADJOINT_CASES
for name in sorted(ADJOINT_CASES):
    for step, e in enumerate(run_pair(name), 1):
        print(name, step, " ".join("%s=%.1e" % kv for kv in e.items()))
