"""BASELINE.json configs[2]: mu.tl.mofa on 2 modalities (RNA n x 30k, 7 % + ATAC n x 200k, 3 %), k=30,
exactly 15 iterations, cells sharded over the ranks (weak scaling: --cells per GPU).

    python profiles/bench_mofa.py [--cells 1000000] [--iters 15] [--k 30]
    torchrun --nproc-per-node N profiles/bench_mofa.py ...

Prints one JSON line (rank 0): cells/s for the 15-iteration fit, ms per iteration, per-kernel CUDA-event
times, and the SpMM kernel's algorithmic GB/s (MOFA per-iteration bytes: SURVEY section 8d)."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from muon_b200 import _device, _lib  # noqa: E402
from muon_b200._mofa import run_mofa_device  # noqa: E402
from muon_b200._synth import generate_device, make_tables  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cells", type=int, default=1_000_000)
ap.add_argument("--iters", type=int, default=15)
ap.add_argument("--k", type=int, default=30)
ap.add_argument("--rna-genes", type=int, default=30_000)
ap.add_argument("--peaks", type=int, default=200_000)
args = ap.parse_args()

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n, N = args.cells, args.cells * world

views = []
for D, dens, seed in ((args.rna_genes, 0.07, 2), (args.peaks, 0.03, 1)):
    tb = make_tables(D, dens, 64, seed)
    A = generate_device(n, D, dens, tables=tb, row0=rank * n, n_total=N)
    # non-integer, log-normalised values (gaussian likelihood): the TF-IDF kernel, in place
    views.append(_device.tfidf_csr(A, inplace_values=True))
nnz = [v.nnz for v in views]
Z0 = torch.from_numpy(np.random.RandomState(1).normal(size=(N, args.k))[rank * n:(rank + 1) * n])
torch.cuda.synchronize()

run_mofa_device(views, args.k, 2, N, Z0, check_convergence=False)          # warm-up (builds transposes: cached)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
_lib.PROFILE = {}
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
res = run_mofa_device(views, args.k, args.iters, N, Z0, check_convergence=False)
e1.record()
torch.cuda.synchronize()
prof, _lib.PROFILE = _lib.PROFILE, None
ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
ms = float(ms[0])
kern = {k: float(np.sum([a.elapsed_time(b) for a, b in v])) for k, v in prof.items()}
P = _device.pad_width(args.k)
spmm_bytes = sum(2 * (8.0 * z + 4.0 * P * (n + v.shape[1])) for z, v in zip(nnz, views)) * args.iters
if rank == 0:
    print(json.dumps({
        "metric": "cells/sec for mu.tl.mofa (2 modalities, k=%d, %d iterations)" % (args.k, args.iters),
        "value": N / (ms / 1e3), "unit": "cells/s", "n_gpus": world, "ms_total": ms, "ms_per_iteration": ms / args.iters,
        "config": {"cells_per_gpu": n, "rna": [n, args.rna_genes, nnz[0]], "atac": [n, args.peaks, nnz[1]]},
        "kernel_ms": kern, "spmm_algorithmic_GBps": spmm_bytes / (kern.get("mub_spmm_csr_f32", float("nan")) * 1e-3) / 1e9,
        "elbo_first_last": [res["elbo"][0], res["elbo"][-1]], "variance_top5": [v[:5].tolist() for v in res["variance"]]}))
if world > 1:
    dist.destroy_process_group()
