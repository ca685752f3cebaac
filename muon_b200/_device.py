"""Device-resident CSR container and thin wrappers over the C-ABI kernels.

torch supplies device memory, streams and (in ``_dist``) NCCL; every numerical kernel on
the hot path is a call into ``libmuon_b200.so``.  Nothing here has a CPU fallback.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import _dist
from ._lib import (TFIDF_LOG_IDF, TFIDF_LOG_TF, TFIDF_LOG_TFIDF, TFIDF_NO_SCALE, MuonB200Error, call, load, ptr,
                   stream_ptr)

PAD_WIDTHS = (32, 64, 128)


def require_cuda():
    if not torch.cuda.is_available():
        raise MuonB200Error("muon_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    load()


def pad_width(l: int) -> int:
    for p in PAD_WIDTHS:
        if l <= p:
            return p
    raise NotImplementedError(f"dense block width {l} > 128 is not supported yet")


_STAGE_BYTES = 256 << 20


def to_device(arr, device, dtype=None) -> torch.Tensor:
    """Host array -> device tensor of numpy dtype ``dtype`` (default: unchanged).

    Pinned inputs go in one async copy.  Pageable inputs are staged through two pinned buffers
    so the PCIe copy of one chunk overlaps the host memcpy of the next; a dtype change
    (e.g. scipy's int64 indices -> int32) happens on the device per chunk, never on the host."""
    if isinstance(arr, torch.Tensor):
        t = arr
    else:
        a = np.ascontiguousarray(arr)
        if not a.flags.writeable:
            a = a.copy()
        t = torch.from_numpy(a)
    tgt = t.dtype if dtype is None else getattr(torch, np.dtype(dtype).name)
    if t.device.type == "cuda":
        return t if t.dtype == tgt else t.to(tgt)
    nbytes = t.numel() * t.element_size()
    if t.is_pinned() or nbytes <= _STAGE_BYTES:
        out = t.to(device, non_blocking=True)
        return out if out.dtype == tgt else out.to(tgt)
    out = torch.empty(t.shape, dtype=tgt, device=device)
    flat_src, flat_dst = t.reshape(-1), out.reshape(-1)
    step = _STAGE_BYTES // t.element_size()
    stage = [torch.empty(step, dtype=t.dtype).pin_memory() for _ in range(2)]
    dstage = [torch.empty(step, dtype=t.dtype, device=device) for _ in range(2)] if tgt != t.dtype else None
    events = [None, None]
    for i, off in enumerate(range(0, flat_src.numel(), step)):
        s = stage[i & 1]
        if events[i & 1] is not None:
            events[i & 1].synchronize()
        n = min(step, flat_src.numel() - off)
        s[:n].copy_(flat_src[off:off + n])
        if dstage is None:
            flat_dst[off:off + n].copy_(s[:n], non_blocking=True)
        else:
            dstage[i & 1][:n].copy_(s[:n], non_blocking=True)
            flat_dst[off:off + n].copy_(dstage[i & 1][:n])
        ev = torch.cuda.Event()
        ev.record()
        events[i & 1] = ev
    torch.cuda.current_stream().synchronize()
    return out


def to_host(t: torch.Tensor, out: Optional[np.ndarray] = None) -> np.ndarray:
    """Device tensor -> numpy array, staged through pinned buffers for large transfers."""
    nbytes = t.numel() * t.element_size()
    if out is None:
        out = np.empty(tuple(t.shape), dtype=getattr(np, str(t.dtype).split(".")[1]))
    dst = torch.from_numpy(out).reshape(-1)
    src = t.reshape(-1)
    if nbytes <= _STAGE_BYTES or dst.is_pinned():
        dst.copy_(src)
        return out
    step = _STAGE_BYTES // t.element_size()
    stage = [torch.empty(step, dtype=t.dtype).pin_memory() for _ in range(2)]
    events = [None, None]
    pending = [None, None]
    chunks = list(range(0, src.numel(), step))
    for i, off in enumerate(chunks + [None]):
        if off is not None:
            n = min(step, src.numel() - off)
            stage[i & 1][:n].copy_(src[off:off + n], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            events[i & 1], pending[i & 1] = ev, (off, n)
        j = (i - 1) & 1
        if i >= 1 and pending[j] is not None:
            events[j].synchronize()
            o, n = pending[j]
            dst[o:o + n].copy_(stage[j][:n])
            pending[j] = None
    return out


class DeviceCSR:
    """CSR matrix resident in HBM: ``indptr`` int64 [n+1], ``indices`` int32 [nnz], ``data`` [nnz].

    Quacks enough like a scipy matrix (``shape``, ``dtype``, ``nnz``, ``get()``) that it can sit
    in ``adata.X`` between ``tfidf`` and ``lsi`` so that the matrix never leaves the GPU.
    ``row0``/``n_total`` describe a cell shard of a larger matrix (one shard per rank).
    """

    def __init__(self, indptr, indices, data, shape, row0: int = 0, n_total: Optional[int] = None,
                 sorted_indices: bool = True):
        self.indptr, self.indices, self.data = indptr, indices, data
        self.sorted_indices = sorted_indices
        self.shape = (int(shape[0]), int(shape[1]))
        self.row0 = int(row0)
        self._n_total = int(n_total) if n_total is not None else None
        self._t = None  # cached transpose (DeviceCSR of A^T), invalidated when data is rebound
        assert indptr.dtype == torch.int64 and indices.dtype == torch.int32
        assert indptr.numel() == self.shape[0] + 1

    @property
    def n_total(self) -> int:
        """Number of cells of the whole (possibly sharded) matrix: explicit, else the sum of the
        shard heights over the process group, else the local height."""
        if self._n_total is None:
            if _dist.is_distributed():
                t = torch.tensor([self.shape[0]], dtype=torch.int64, device=self.data.device)
                self._n_total = int(_dist.all_reduce_sum_(t)[0])
            else:
                self._n_total = self.shape[0]
        return self._n_total

    # -- scipy-ish surface -------------------------------------------------------------
    @property
    def dtype(self):
        return np.dtype(np.float32 if self.data.dtype == torch.float32 else np.float64)

    @property
    def nnz(self) -> int:
        return int(self.data.numel())

    @property
    def device(self):
        return self.data.device

    def with_data(self, data) -> "DeviceCSR":
        """Same sparsity pattern (shared index tensors), new values."""
        return DeviceCSR(self.indptr, self.indices, data, self.shape, self.row0, self._n_total,
                         self.sorted_indices)

    def copy(self) -> "DeviceCSR":
        return self.with_data(self.data.clone())

    @classmethod
    def from_scipy(cls, X, device=None, dtype=None) -> "DeviceCSR":
        """Upload a canonical scipy CSR (or any object exposing indptr/indices/data/shape)."""
        require_cuda()
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        data = X.data
        if dtype is None:
            dtype = np.float32 if data.dtype == np.float32 else (np.float64 if data.dtype == np.float64 else np.float32)
        return cls(to_device(X.indptr, device, np.int64), to_device(X.indices, device, np.int32),
                   to_device(data, device, dtype), X.shape)

    def get(self, indptr_host=None, indices_host=None):
        """Download as scipy.sparse.csr_matrix.  Host index arrays may be passed to be reused
        (the sparsity pattern is never modified on the device)."""
        import scipy.sparse as sp
        data = to_host(self.data)
        indptr = to_host(self.indptr) if indptr_host is None else indptr_host
        indices = to_host(self.indices) if indices_host is None else indices_host
        if self.nnz < 2**31 - 1 and indptr.dtype != np.int32:
            indptr = indptr.astype(np.int32)
        m = sp.csr_matrix((data, indices, indptr), shape=self.shape, copy=False)
        m.has_sorted_indices = bool(self.sorted_indices)
        return m

    # -- transpose (cached) --------------------------------------------------------------
    def transpose(self) -> "DeviceCSR":
        if self._t is None:
            self._t = csr_transpose(self)
        return self._t


# ------------------------------------------------------------------------------------------
def tfidf_csr(A: DeviceCSR, log_tf=True, log_idf=True, log_tfidf=False, scale_factor=1e4,
              inplace_values=False, check_canonical=False) -> Optional[DeviceCSR]:
    """K1: two fused passes (reduce, apply).  Column sums are allreduced across cell shards
    (muon/_atac/preproc.py:92-119; multi-GPU plan SURVEY section 8e)."""
    require_cuda()
    n, d = A.shape
    sfx = "f32" if A.data.dtype == torch.float32 else "f64"
    flags = (TFIDF_LOG_TF if log_tf else 0) | (TFIDF_LOG_IDF if log_idf else 0) | (TFIDF_LOG_TFIDF if log_tfidf else 0)
    if scale_factor is None or scale_factor == 0 or scale_factor == 1:
        flags |= TFIDF_NO_SCALE
        scale_factor = 1.0
    dev, dt = A.data.device, A.data.dtype
    row_sum = torch.empty(n, dtype=dt, device=dev)
    col_sum = torch.zeros(d, dtype=dt, device=dev)
    st = stream_ptr()
    status = torch.zeros(1, dtype=torch.int32, device=dev) if check_canonical else None
    call(f"mub_tfidf_reduce_{sfx}", ptr(A.indptr), ptr(A.indices), ptr(A.data), n, d, ptr(row_sum), ptr(col_sum),
         ptr(status), st)
    if check_canonical and int(status[0]) != 0:
        return None  # caller canonicalises (duplicates / explicit zeros / unsorted rows) and retries
    _dist.all_reduce_sum_(col_sum)
    idf = torch.empty(d, dtype=dt, device=dev)
    call(f"mub_tfidf_idf_{sfx}", ptr(col_sum), d, float(A.n_total), flags, ptr(idf), st)
    out = A.data if inplace_values else torch.empty_like(A.data)
    call(f"mub_tfidf_apply_{sfx}", ptr(A.indptr), ptr(A.indices), ptr(A.data), ptr(out), n, d, ptr(row_sum),
         ptr(idf), float(scale_factor), flags, st)
    res = A.with_data(out)
    res._aux = {"row_sum": row_sum, "col_sum": col_sum, "idf": idf}
    return res


def csr_transpose(A: DeviceCSR) -> DeviceCSR:
    """Build the CSR of A^T on the device (count -> scan -> atomic-cursor fill)."""
    require_cuda()
    assert A.data.dtype == torch.float32, "transpose: float32 values only"
    n, d = A.shape
    dev = A.data.device
    st = stream_ptr()
    t_count = torch.zeros(d + 1, dtype=torch.int64, device=dev)
    call("mub_csr_transpose_count", ptr(A.indices), A.nnz, d, ptr(t_count), st)
    t_indptr = torch.cumsum(t_count, 0)
    cursor = torch.empty(max(d, 1), dtype=torch.int64, device=dev)
    t_indices = torch.empty(A.nnz, dtype=torch.int32, device=dev)
    t_data = torch.empty(A.nnz, dtype=torch.float32, device=dev)
    call("mub_csr_transpose_fill", ptr(A.indptr), ptr(A.indices), ptr(A.data), n, d, 0, ptr(t_indptr), ptr(cursor),
         ptr(t_indices), ptr(t_data), st)
    return DeviceCSR(t_indptr, t_indices, t_data, (d, n), n_total=d, sorted_indices=False)


def spmm(A: DeviceCSR, B: torch.Tensor, out: Optional[torch.Tensor] = None, accumulate=False,
         dynamic=True) -> torch.Tensor:
    """K2/K3: C[n x P] (+)= A @ B[d x P];  P = B.shape[1] must be 32, 64 or 128."""
    n, d = A.shape
    P = B.shape[1]
    assert B.shape[0] == d and B.dtype == torch.float32 and B.is_contiguous(), (B.shape, d, B.dtype)
    if out is None:
        assert not accumulate
        out = torch.empty((n, P), dtype=torch.float32, device=B.device)
    counter = torch.zeros(1, dtype=torch.int64, device=B.device) if dynamic else None
    call("mub_spmm_csr_f32", ptr(A.indptr), ptr(A.indices), ptr(A.data), n, d, ptr(B), P, ptr(out),
         1 if accumulate else 0, ptr(counter), stream_ptr())
    return out


def gram(Y: torch.Tensor, l: Optional[int] = None, weights: Optional[torch.Tensor] = None,
         reduce=True) -> torch.Tensor:
    """K4: G[l x l] = Y^T diag(w) Y in float64 (allreduced over cell shards if ``reduce``)."""
    lib = load()
    n, P = Y.shape
    l = P if l is None else l
    assert Y.dtype == torch.float32 and Y.is_contiguous()
    ws = torch.empty(max(int(lib.mub_gram_workspace_bytes(n, P)), 4), dtype=torch.uint8, device=Y.device)
    G = torch.empty((l, l), dtype=torch.float64, device=Y.device)
    call("mub_gram_f32", ptr(Y), ptr(weights), n, P, l, ptr(G), ptr(ws), stream_ptr())
    if reduce:
        _dist.all_reduce_sum_(G)
    return G


# ------------------------------------------------------------------------------------------
# Host matrices produced by tfidf() keep a handle to their device twin so that a following
# lsi() on the same AnnData does not pay the PCIe upload again.  The handle is validated
# against the host values (sampled fingerprint) before use, so editing X on the host is safe.
_RESIDENT_ATTR = "_mub_resident"


def remember_resident(host_matrix, dev: "DeviceCSR"):
    try:
        setattr(host_matrix, _RESIDENT_ATTR, dev)
    except Exception:
        pass


def recall_resident(host_matrix) -> Optional["DeviceCSR"]:
    dev = getattr(host_matrix, _RESIDENT_ATTR, None)
    if dev is None or not isinstance(dev, DeviceCSR):
        return None
    try:
        if tuple(dev.shape) != tuple(host_matrix.shape) or dev.nnz != host_matrix.nnz:
            return None
        if host_matrix.data.dtype != np.float32 or dev.data.dtype != torch.float32:
            return None
        nnz = dev.nnz
        if nnz:
            pos = np.unique(np.linspace(0, nnz - 1, num=min(nnz, 4096), dtype=np.int64))
            idx = torch.from_numpy(pos).to(dev.data.device)
            if not np.array_equal(dev.data[idx].cpu().numpy(), host_matrix.data[pos]):
                return None
            if not np.array_equal(dev.indices[idx].cpu().numpy(), np.asarray(host_matrix.indices[pos], dtype=np.int32)):
                return None
        return dev
    except Exception:
        return None
