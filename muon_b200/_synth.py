"""Synthetic ATAC count matrices (benchmark and test INPUT; SURVEY App. E, adapted).

A planted-topic Bernoulli model evaluated with a counter-based hash so that any row range
is reproducible on any number of GPUs.  ``generate_host`` (numpy) and ``generate_device``
(``csrc/synth.cu``) are bit-identical: only IEEE-exact float32 operations (+, *, min, float->
uint truncation) and 64-bit integer hashing are used, no transcendental functions.

    p_ij  = min(0.9, (0.5*beta_j + topic[t_i, j]) * (depth_i * kappa))
    keep  = hi32(h_ij) < uint32(p_ij * 2^32)      h_ij = mix64(mix64(seed) + i*n_cols + j)
    count = 1 + capped geometric(0.6) from lo32(h_ij)     ("mostly 1 and 2", docs atac.rst:72)

Cells belong to one of ``n_topics`` topics with unequal strengths so that the leading
singular values are separated; ``kappa`` is solved on the host so the expected density hits
the target.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def mix64(x):
    """splitmix64 finaliser on uint64 arrays (wrap-around arithmetic)."""
    with np.errstate(over="ignore"):
        x = (np.asarray(x, dtype=np.uint64) + np.uint64(0x9E3779B97F4A7C15))
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


def _u24(h):
    """24-bit uniform in [0,1) as float32 (exact)."""
    return (h >> np.uint64(40)).astype(np.float32) * np.float32(2.0 ** -24)


@dataclass
class SynthTables:
    n_cols: int
    n_topics: int
    seed: int
    kappa: np.float32
    beta: np.ndarray        # [n_cols] float32
    topic: np.ndarray       # [n_topics, n_cols] float32
    density: float

    def rows(self, row0: int, n_rows: int):
        """Per-cell topic id (int32) and scale depth*kappa (float32) for rows [row0, row0+n_rows)."""
        r = np.arange(row0, row0 + n_rows, dtype=np.uint64)
        base = mix64(np.uint64(self.seed)) ^ np.uint64(0x5851F42D4C957F2D)
        with np.errstate(over="ignore"):
            h = mix64(base + r * np.uint64(2))
            h2 = mix64(base + r * np.uint64(2) + np.uint64(1))
        # skewed topic sizes: topic id = floor(T * u^1.5)-like without pow: u*sqrt(u) is IEEE-exact
        u = _u24(h)
        t = np.minimum((np.float32(self.n_topics) * (u * np.sqrt(u))).astype(np.int32), self.n_topics - 1)
        u2 = _u24(h2)
        depth = np.float32(0.5) + np.float32(1.5) * (u2 * u2)           # mean 1.0, range [0.5, 2)
        scale = depth * self.kappa
        return t.astype(np.int32), scale.astype(np.float32)


def make_tables(n_cols: int, density: float, n_topics: int = 64, seed: int = 0) -> SynthTables:
    j = np.arange(n_cols, dtype=np.uint64)
    s0 = mix64(np.uint64(seed))
    with np.errstate(over="ignore"):
        ub = _u24(mix64(s0 ^ np.uint64(0xB5297A4D3F84D5B5) + j))
        beta = (ub * ub * ub).astype(np.float32)                        # skewed baseline in [0,1)
        topic = np.zeros((n_topics, n_cols), dtype=np.float32)
        for t in range(n_topics):
            h = mix64(s0 + np.uint64(0x1000003) * np.uint64(t + 1) + j * np.uint64(0x9E3779B1))
            member = (h >> np.uint64(32)).astype(np.uint32) < np.uint32(int(0.08 * 2**32))
            u = ((h & np.uint64(0xFFFFFF)).astype(np.float32) * np.float32(2.0 ** -24))
            strength = np.float32(1.0 - 0.6 * t / max(n_topics, 1))
            w = strength * (np.float32(0.2) + np.float32(0.8) * (u * u))
            topic[t] = np.where(member, w, np.float32(0)).astype(np.float32)
    # expected density (ignoring the 0.9 clip) = kappa * E[depth] * mean_t,j(0.5 beta + topic), E[depth]=1;
    # topic prior of a cell: P(t) for t = floor(T u^1.5)
    edges = (np.arange(n_topics + 1) / n_topics) ** (2.0 / 3.0)
    pt = np.diff(edges)
    mean_base = float(0.5 * beta.astype(np.float64).mean() + (pt[:, None] * topic.astype(np.float64)).sum(0).mean())
    kappa = np.float32(density / mean_base)
    return SynthTables(n_cols, n_topics, seed, kappa, beta, topic, density)


_GEOM_THR = np.array([2576980378, 3607772529, 4020089389, 4185016133], dtype=np.uint64)


def _host_chunk(tb, seedmix, j, half_beta, rt, rs, n_cols, row_lo, row_hi, c0, c1):
    """Rows [c0, c1) of the chunked generator (numpy releases the GIL inside these array operations)."""
    base = (half_beta[None, :] + tb.topic[rt[c0:c1]]).astype(np.float32)
    np.multiply(base, rs[c0:c1, None], out=base)
    np.minimum(base, np.float32(0.9), out=base)
    thr = (base * np.float32(4294967296.0)).astype(np.uint32)
    del base
    with np.errstate(over="ignore"):
        rowkey = seedmix + (np.arange(row_lo + c0, row_lo + c1, dtype=np.uint64) * np.uint64(n_cols))
        h = mix64(rowkey[:, None] + j[None, :])
    keep = (h >> np.uint64(32)).astype(np.uint32) < thr
    del thr
    r, c = np.nonzero(keep)
    lo = h[r, c] & np.uint64(0xFFFFFFFF)
    del h, keep
    val = np.float32(1.0) + (lo[:, None] >= _GEOM_THR[None, :]).sum(1).astype(np.float32)
    return np.bincount(r, minlength=c1 - c0), c.astype(np.int32), val.astype(np.float32)


def generate_host(n_rows: int, n_cols: int, density: float, n_topics: int = 64, seed: int = 0,
                  row0: int = 0, tables: SynthTables = None, chunk: int = 64, threads: int = None):
    """numpy twin of the device generator -> scipy.sparse.csr_matrix (float32, sorted int32 indices).
    Row chunks are generated on a thread pool (``threads``: default min(32, cores))."""
    import os

    import scipy.sparse as sp
    tb = tables or make_tables(n_cols, density, n_topics, seed)
    rt, rs = tb.rows(row0, n_rows)
    seedmix = mix64(np.uint64(tb.seed))
    j = np.arange(n_cols, dtype=np.uint64)
    half_beta = np.float32(0.5) * tb.beta
    indptr = np.zeros(n_rows + 1, dtype=np.int64)
    starts = list(range(0, n_rows, chunk))
    if threads is None:
        try:
            threads = len(os.sched_getaffinity(0))
        except Exception:
            threads = os.cpu_count() or 1
        threads = min(32, threads)
    work = lambda c0: _host_chunk(tb, seedmix, j, half_beta, rt, rs, n_cols, row0, None, c0, min(n_rows, c0 + chunk))  # noqa: E731
    if threads > 1 and len(starts) > 1:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(threads) as pool:
            parts = list(pool.map(work, starts))
    else:
        parts = [work(c0) for c0 in starts]
    for c0, (cnt, _, _) in zip(starts, parts):
        indptr[c0 + 1:c0 + 1 + cnt.shape[0]] = cnt
    indptr = np.cumsum(indptr)
    indices = np.concatenate([p[1] for p in parts]) if parts else np.zeros(0, np.int32)
    data = np.concatenate([p[2] for p in parts]) if parts else np.zeros(0, np.float32)
    if indptr[-1] < 2**31 - 1:
        indptr = indptr.astype(np.int32)
    X = sp.csr_matrix((data, indices, indptr), shape=(n_rows, n_cols))
    X.has_sorted_indices = True
    return X


def generate_device(n_rows: int, n_cols: int, density: float, n_topics: int = 64, seed: int = 0,
                    row0: int = 0, n_total: int = None, tables: SynthTables = None, device=None):
    """Generate rows [row0, row0+n_rows) directly in HBM -> DeviceCSR (float32)."""
    import torch

    from . import _device
    from ._lib import call, ptr, stream_ptr
    _device.require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    tb = tables or make_tables(n_cols, density, n_topics, seed)
    rt, rs = tb.rows(row0, n_rows)
    beta = torch.from_numpy(tb.beta).to(dev)
    topic = torch.from_numpy(tb.topic).to(dev)
    rt_d, rs_d = torch.from_numpy(rt).to(dev), torch.from_numpy(rs).to(dev)
    st = stream_ptr()
    row_nnz = torch.empty(n_rows, dtype=torch.int64, device=dev)
    call("mub_synth_count", row0, n_rows, n_cols, ptr(beta), ptr(topic), ptr(rt_d), ptr(rs_d), tb.seed,
         ptr(row_nnz), st)
    indptr = torch.zeros(n_rows + 1, dtype=torch.int64, device=dev)
    torch.cumsum(row_nnz, 0, out=indptr[1:])
    nnz = int(indptr[-1])
    indices = torch.empty(nnz, dtype=torch.int32, device=dev)
    data = torch.empty(nnz, dtype=torch.float32, device=dev)
    call("mub_synth_fill", row0, n_rows, n_cols, ptr(beta), ptr(topic), ptr(rt_d), ptr(rs_d), tb.seed,
         ptr(indptr), ptr(indices), ptr(data), st)
    return _device.DeviceCSR(indptr, indices, data, (n_rows, n_cols), row0=row0,
                             n_total=n_total if n_total is not None else n_rows)
