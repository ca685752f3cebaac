"""Micro-benchmark of the SpMM kernels (v1 row-warp vs v2 TMA panel) on a slice of the bench matrix.
    python profiles/bench_spmm.py [cells] [peaks] [P]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from muon_b200 import _device  # noqa: E402
from muon_b200._synth import generate_device, make_tables  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 200_000
P = int(sys.argv[3]) if len(sys.argv) > 3 else 64
A = generate_device(n, d, 0.03, tables=make_tables(d, 0.03, 64, 1))
B = torch.randn((d, P), device="cuda")
ref = None
for algo in ("rowwarp", "panel"):
    for _ in range(10):   # warm up clocks and caches
        C = _device.spmm(A, B, algo=algo, dynamic=False)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(10):
        C = _device.spmm(A, B, algo=algo, dynamic=False)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / 10
    gb = (8 * A.nnz + 4 * P * (n + d)) / 1e9
    err = 0.0 if ref is None else float((C - ref).abs().max() / ref.abs().max())
    ref = C if ref is None else ref
    print(f"{algo:8s} n={n} d={d} P={P} nnz={A.nnz}: {ms:.3f} ms  {gb / ms * 1e3:.0f} GB/s algorithmic  "
          f"{ms * 1e9 / A.nnz:.1f} ps/nnz  maxrel diff vs v1 {err:.2e}")
