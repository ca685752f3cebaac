#!/usr/bin/env python
"""bench.py -- TF-IDF + LSI(k=50) throughput on synthetic sparse ATAC (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A "step" is one pass of the hot path -- mu.atac.pp.tfidf + mu.atac.tl.lsi -- over the whole
synthetic matrix.  Default workload = BASELINE.json configs[1]: 1M cells x 200k peaks, 3 % nnz
per GPU (weak scaling: every rank owns 1M cells; peaks-space objects are replicated, one
allreduce of column sums, one of A^T Y per Lanczos step, one of the b x b Gram per QR).

Prints ONE JSON line (rank 0).  value = cells/s with the counts already resident in HBM;
e2e = the same calls on HOST scipy matrices (H2D of indices/values and D2H of the TF-IDF values
and the factors inside the timed region); roofline = the dominant kernel (CSR SpMM) measured
live with CUDA events; cpu_baseline = the scipy oracle on a bounded row-sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "cells/sec for TF-IDF+LSI(k=50) on 1Mx200k sparse ATAC"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cells", type=int, default=1_000_000, help="cells per GPU")
    ap.add_argument("--peaks", type=int, default=200_000)
    ap.add_argument("--density", type=float, default=0.03)
    ap.add_argument("--k", type=int, default=50)
    ap.add_argument("--topics", type=int, default=0,
                    help="planted topics of the synthetic matrix (0: max(64, k+14), so that the k wanted components "
                         "are separated from the noise bulk, SURVEY App. E)")
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--tol", type=float, default=1e-5)
    ap.add_argument("--sample-cells", type=int, default=2000, help="rows of the CPU-baseline sample")
    ap.add_argument("--e2e-cells", type=int, default=-1,
                    help="cells per GPU for the e2e leg (-1: same as --cells at 1 GPU; 250k per GPU under torchrun, "
                         "where N ranks share one host's PCIe switches, memory bandwidth and RAM)")
    ap.add_argument("--e2e-steps", type=int, default=1)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--breakdown", action="store_true", help="one extra, synchronising step with phase timers")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v.lower() == "active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def onchip_roofline(nnz, P, ms_per_pass, sm_mhz, n_sms=148):
    """Second roofline of the SpMM kernel: every non-zero moves 4*P bytes of the dense operand plus its 8-byte
    {index, value} entry through the SM's L1/LSU data path (128 B/clk/SM, shared with shared-memory traffic), so
    at P=64 a pass cannot take less than ~2 clocks per non-zero per SM however little DRAM traffic it causes.
    Reported next to the HBM figure because this, not DRAM, is what the kernel saturates."""
    if not sm_mhz or not ms_per_pass or ms_per_pass != ms_per_pass:
        return None
    bytes_per_pass = float(nnz) * (4.0 * P + 8.0)
    peak = n_sms * 128.0 * sm_mhz * 1e6 / 1e12                     # TB/s
    ach = bytes_per_pass / (ms_per_pass * 1e-3) / 1e12
    return {"bound": "l1-lsu data path (128 B/clk/SM)", "achieved": ach, "peak": peak, "unit": "TB/s",
            "frac": ach / peak, "bytes_per_nnz": 4.0 * P + 8.0, "sm_mhz": sm_mhz, "sms": n_sms}


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


# ------------------------------------------------------------------------------------------------
def cpu_oracle_step(X, k):
    """The reference's CPU path (scipy restatement of preproc.py:92-119 + svds + tools.py:56-65)."""
    from oracle.lsi_ref import lsi_ref
    from oracle.tfidf_ref import tfidf_ref
    t0 = time.perf_counter()
    Y = tfidf_ref(X)
    t1 = time.perf_counter()
    lsi_ref(Y, k)
    t2 = time.perf_counter()
    return t1 - t0, t2 - t1


def sample_matrix(args, n_rows):
    """First n_rows rows of rank 0's shard of the benchmark matrix, as a host scipy CSR."""
    import torch
    from muon_b200._synth import generate_device, generate_host, make_tables
    tb = make_tables(args.peaks, args.density, args.topics, args.seed)
    if torch.cuda.is_available():
        return generate_device(n_rows, args.peaks, args.density, tables=tb, row0=0).get()
    return generate_host(n_rows, args.peaks, args.density, tables=tb, row0=0)


def run_reference(args, rank):
    """--impl reference: the oracle (kind 'port': import muon is impossible here) timed on
    the host cores on a bounded row-sample of the same workload.  Rank 0 only."""
    if rank != 0:
        return
    from threadpoolctl import threadpool_limits  # noqa: F401  (BLAS uses all cores by default)
    S = args.sample_cells
    X = sample_matrix(args, S)
    for _ in range(args.warmup):
        cpu_oracle_step(X, args.k)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_oracle_step(X, args.k)
    dt = time.perf_counter() - t0
    v = S * args.steps / dt
    sample = f"first {S} cells x {args.peaks} peaks ({X.nnz} nnz) of the synthetic matrix"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "cells/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"TF-IDF + LSI k={args.k} on {args.cells}x{args.peaks} ATAC ({args.density:.0%} nnz) per GPU",
                   "sample": sample},
        "cpu_baseline": {"value": v, "unit": "cells/s", "cores": os.cpu_count(), "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# ------------------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.topics <= 0:
        args.topics = max(64, args.k + 14)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist

    import muon_b200 as mu
    from muon_b200 import _lib
    from muon_b200._synth import generate_device, make_tables

    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback in the product path)"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n_local, D, k = args.cells, args.peaks, args.k
    n_total = n_local * world
    tb = make_tables(D, args.density, args.topics, args.seed)
    A = generate_device(n_local, D, args.density, tables=tb, row0=rank * n_local, n_total=n_total)
    nnz = A.nnz
    torch.cuda.synchronize()

    import pandas as pd
    obs0 = pd.DataFrame(index=pd.RangeIndex(n_local).astype(str))
    var0 = pd.DataFrame(index=pd.RangeIndex(D).astype(str))

    def step():
        ad = mu.SimpleAnnData(A, obs=obs0, var=var0)  # counts stay untouched: tfidf writes a new matrix
        mu.atac.pp.tfidf(ad)
        info = mu.atac.tl.lsi(ad, n_comps=k, tol=args.tol, return_info=True)
        return ad, info

    for _ in range(args.warmup):
        ad, info = step()
        del ad
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    _lib.PROFILE = {}
    launches0 = _lib.LAUNCHES
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        ad, info = step()
        del ad
    e1.record()
    barrier()
    prof, _lib.PROFILE = _lib.PROFILE, None
    launches = _lib.LAUNCHES - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms[0])
    value = n_total * args.steps / (ms_total / 1e3)

    # ---- per-kernel device times (CUDA events on the launching stream, inside the timed region)
    kern = {name: [a.elapsed_time(b) for a, b in evs] for name, evs in prof.items()}
    spmm_ms = kern.get("mub_spmm_csr_f32", []) + kern.get("mub_spmm_csrp_f32", [])   # A*V and A^T*U (pairs layout)
    P = mu._device.pad_width(min(k + 8, 128))
    # one "pass" = one product with A or A^T over all nnz of the shard (A^T runs as several row-panel
    # launches).  Algorithmic bytes per pass: 8 B/nnz + the dense operands once.
    pass_bytes = 8.0 * nnz + 4.0 * P * (n_local + D)
    n_pass = info.passes * args.steps
    spmm_total = float(np.sum(spmm_ms)) if spmm_ms else float("nan")
    pk, pk_kind = peaks()
    hbm = float(pk.get("hbm_gbs", 6650.0))
    ach = n_pass * pass_bytes / (spmm_total * 1e-3) / 1e9 if spmm_ms else None
    roofline = {"bound": "hbm", "kernel": f"spmm_csr_rowwarp_kernel<{P},*> (A*V on CSR arrays, A^T*U on pair-layout panels)", "achieved": ach, "peak": hbm,
                "unit": "GB/s", "frac": (ach / hbm) if ach else None, "traffic": None, "peak_kind": pk_kind,
                "launches": len(spmm_ms), "passes": n_pass, "ms_per_pass": spmm_total / max(n_pass, 1),
                "bytes_per_pass": pass_bytes,
                "note": "algorithmic bytes = 8 B/nnz + 4*P*(n+D) per pass; every nnz also gathers 4*P B of the dense "
                        "operand through L2->L1 (ncu: ~80 % of L2 bandwidth), which is the practical limiter (DESIGN.md)"}
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        key = f"spmm_rowwarp{P}_n{n_local}_d{D}"
        if key in tr:
            roofline["traffic"] = tr[key]["dram_bytes_per_pass"]
            roofline["traffic_source"] = tr[key]["source"]
    except Exception:
        pass
    tf_red = kern.get("mub_tfidf_reduce_f32", [])
    tf_app = kern.get("mub_tfidf_apply_f32", [])
    if tf_red and tf_app:
        t = float(np.mean(tf_red) + np.mean(tf_app))
        roofline["tfidf"] = {"achieved": 20.0 * nnz / (t * 1e-3) / 1e9, "unit": "GB/s", "ms": t,
                             "frac": 20.0 * nnz / (t * 1e-3) / 1e9 / hbm}
    if rank == 0 and clocks and spmm_ms:
        roofline["onchip"] = onchip_roofline(nnz, P, spmm_total / max(n_pass, 1), clocks.get("sm_mhz"),
                                             torch.cuda.get_device_properties(local).multi_processor_count)
    phase_ms = {name: float(np.sum(v)) / args.steps for name, v in kern.items()}

    breakdown = None
    if args.breakdown:
        _lib.PHASES = {}
        t0 = time.perf_counter()
        with _lib.phase("total"):
            ad, _ = step()
        del ad
        breakdown = {k: round(1e3 * v, 2) for k, v in _lib.PHASES.items()}
        _lib.PHASES = None

    # ---- e2e: same public calls on HOST matrices -------------------------------------------------
    e2e = None
    if not args.no_e2e:
        import scipy.sparse as sp
        ne = (n_local if world == 1 else min(n_local, 250_000)) if args.e2e_cells < 0 else min(args.e2e_cells, n_local)
        Ae = A if ne == n_local else generate_device(ne, D, args.density, tables=tb, row0=rank * ne, n_total=ne * world)
        X = Ae.get()                                    # host scipy CSR (int64 indices when nnz >= 2^31)
        h2d = X.indptr.nbytes + X.indices.nbytes + X.data.nbytes
        d2h = X.data.nbytes + 4 * (ne * k + D * k + k)
        if Ae is not A:
            del Ae

        import pandas as pd
        obs_df = pd.DataFrame(index=pd.RangeIndex(ne).astype(str))
        var_df = pd.DataFrame(index=pd.RangeIndex(D).astype(str))

        def step_host():
            ad = mu.SimpleAnnData(X, obs=obs_df, var=var_df)   # tfidf rebinds ad.X; X itself is never modified
            mu.atac.pp.tfidf(ad)
            mu.atac.tl.lsi(ad, n_comps=k, tol=args.tol)
            return float(ad.uns["lsi"]["stdev"][0])

        step_host()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            step_host()
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": ne * world * args.e2e_steps / float(dt[0]), "unit": "cells/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "cells_per_gpu": ne, "steps": args.e2e_steps,
               "ms_per_step": 1e3 * float(dt[0]) / args.e2e_steps,
               "path": "scipy csr on host -> mu.atac.pp.tfidf -> mu.atac.tl.lsi -> numpy slots"}
        del X

    # ---- CPU baseline on a bounded sample (rank 0, N=1 only) --------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        S = args.sample_cells
        Xs = sample_matrix(args, S)
        t_tfidf, t_lsi = cpu_oracle_step(Xs, k)
        cpu = {"value": S / (t_tfidf + t_lsi), "unit": "cells/s", "cores": os.cpu_count(), "kind": "port",
               "sample": f"first {S} cells x {D} peaks ({Xs.nnz} nnz); tfidf {t_tfidf:.2f}s + svds {t_lsi:.2f}s; "
                         "scipy sparse kernels are single-threaded, BLAS tail uses all cores",
               "tfidf_nnz_per_s": Xs.nnz / t_tfidf}

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": "cells/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"TF-IDF + LSI k={k} on {n_local}x{D} ATAC ({args.density:.0%} nnz) per GPU",
                       "cells_total": n_total, "nnz_per_gpu": nnz, "parallelism": f"cells-sharded x{world}",
                       "l2": "inputs (8 B/nnz CSR stream) exceed L2 by >100x; no flush needed",
                       "lsi": {"block": info.block, "iterations": info.iterations, "passes": info.passes,
                               "tol": args.tol, "converged": info.converged, "stalled": info.stalled, "max_rel_residual": max(info.residuals),
                               "residual_history": [float("%.3g" % h) for h in info.history]}},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
            "phase_ms_per_step": phase_ms, "breakdown_ms": breakdown,
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
