"""Device-resident CSR container and thin wrappers over the C-ABI kernels.

torch supplies device memory, streams and (in ``_dist``) NCCL; every numerical kernel on
the hot path is a call into ``libmuon_b200.so``.  Nothing here has a CPU fallback.
"""
from __future__ import annotations

import os
from typing import Optional

import numpy as np
import torch

from . import _dist
from ._lib import (TFIDF_BINARIZE, TFIDF_LOG_IDF, TFIDF_LOG_TF, TFIDF_LOG_TFIDF, TFIDF_NO_SCALE, MuonB200Error, call, load, ptr,
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
_PINNED = []          # two cached pinned staging buffers (uint8), allocated on first large transfer
_POOL = None          # thread pool for host-side memcpy into / out of the staging buffers
_COPY_THREADS = 16


def _staging():
    global _POOL
    if not _PINNED:
        _PINNED.extend(torch.empty(_STAGE_BYTES, dtype=torch.uint8).pin_memory() for _ in range(2))
    if _POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        _POOL = ThreadPoolExecutor(_COPY_THREADS)
    return _PINNED


def _parallel_copy(dst: np.ndarray, src: np.ndarray):
    """dst[:] = src for 1-D arrays of equal dtype, split over threads (numpy releases the GIL);
    a single-threaded memcpy (and first-touch page faults) would cap PCIe staging at ~3 GB/s."""
    n = dst.shape[0]
    parts = max(1, min(_COPY_THREADS, (n * dst.itemsize) >> 22))
    if parts == 1:
        np.copyto(dst, src)
        return
    bounds = [n * i // parts for i in range(parts + 1)]
    list(_POOL.map(lambda i: np.copyto(dst[bounds[i]:bounds[i + 1]], src[bounds[i]:bounds[i + 1]]), range(parts)))


def to_device(arr, device, dtype=None) -> torch.Tensor:
    """Host array -> device tensor of numpy dtype ``dtype`` (default: unchanged).

    Pinned inputs go in one async copy.  Pageable inputs are staged through two cached pinned
    buffers (multi-threaded host memcpy overlapping the PCIe copy of the previous chunk); a dtype
    change (e.g. scipy's int64 indices -> int32) happens on the device per chunk, never on the host."""
    if isinstance(arr, torch.Tensor):
        if arr.device.type == "cuda":
            tgt = arr.dtype if dtype is None else getattr(torch, np.dtype(dtype).name)
            return arr if arr.dtype == tgt else arr.to(tgt)
        arr = arr.numpy()
    a = np.ascontiguousarray(arr)
    tgt = getattr(torch, np.dtype(a.dtype if dtype is None else dtype).name)
    nbytes = a.nbytes
    if nbytes <= (8 << 20):
        if not a.flags.writeable:
            a = a.copy()
        out = torch.from_numpy(a).to(device, non_blocking=False)
        return out if out.dtype == tgt else out.to(tgt)
    src_t = getattr(torch, a.dtype.name)
    flat = a.reshape(-1)
    out = torch.empty(a.shape, dtype=tgt, device=device)
    flat_dst = out.reshape(-1)
    stage = _staging()
    step = _STAGE_BYTES // a.itemsize
    dstage = [torch.empty(step, dtype=src_t, device=device) for _ in range(2)] if tgt != src_t else None
    events = [None, None]
    for i, off in enumerate(range(0, flat.shape[0], step)):
        n = min(step, flat.shape[0] - off)
        s = stage[i & 1][: n * a.itemsize].view(src_t)
        if events[i & 1] is not None:
            events[i & 1].synchronize()
        _parallel_copy(s.numpy(), flat[off:off + n])
        if dstage is None:
            flat_dst[off:off + n].copy_(s, non_blocking=True)
        else:
            dstage[i & 1][:n].copy_(s, non_blocking=True)
            flat_dst[off:off + n].copy_(dstage[i & 1][:n])
        ev = torch.cuda.Event()
        ev.record()
        events[i & 1] = ev
    torch.cuda.current_stream().synchronize()
    return out


def to_host(t: torch.Tensor, out: Optional[np.ndarray] = None) -> np.ndarray:
    """Device tensor -> numpy array, staged through the pinned buffers for large transfers."""
    np_dt = np.dtype(str(t.dtype).split(".")[1])
    if out is None:
        out = np.empty(tuple(t.shape), dtype=np_dt)
    nbytes = t.numel() * t.element_size()
    src = t.reshape(-1)
    dst = out.reshape(-1)
    if nbytes <= (8 << 20):
        torch.from_numpy(dst).copy_(src)
        return out
    stage = _staging()
    step = _STAGE_BYTES // t.element_size()
    events, pending = [None, None], [None, None]
    offs = list(range(0, src.numel(), step))
    for i, off in enumerate(offs + [None]):
        if off is not None:
            n = min(step, src.numel() - off)
            stage[i & 1][: n * t.element_size()].view(t.dtype).copy_(src[off:off + n], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            events[i & 1], pending[i & 1] = ev, (off, n)
        j = (i - 1) & 1
        if i >= 1 and pending[j] is not None:
            events[j].synchronize()
            o, n = pending[j]
            _parallel_copy(dst[o:o + n], stage[j][: n * t.element_size()].view(t.dtype).numpy())
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
        # scipy's (data, indices, indptr) constructor scans the index arrays (min/max to pick an index
        # dtype): ~30 s for 6e9 int64 indices.  The arrays are known-good, so attach them to an empty matrix.
        if indices.dtype != indptr.dtype:
            if self.nnz < 2**31 - 1 and max(self.shape) < 2**31 - 1:
                indptr = indptr.astype(np.int32) if indptr.dtype != np.int32 else indptr
                indices = indices.astype(np.int32) if indices.dtype != np.int32 else indices
            else:
                indices = indices.astype(np.int64) if indices.dtype != np.int64 else indices
                indptr = indptr.astype(np.int64) if indptr.dtype != np.int64 else indptr
        m = sp.csr_matrix(self.shape, dtype=data.dtype)
        m.data, m.indices, m.indptr = data, indices, indptr
        m.has_sorted_indices = bool(self.sorted_indices)
        return m

    # -- transpose (cached) --------------------------------------------------------------
    def transpose(self) -> "DeviceCSR":
        if self._t is None:
            self._t = csr_transpose(self)
        return self._t

    def transpose_panels(self, pad: int = 64, side_stream: bool = False) -> "TransposedPanels":
        key = ("panels", pad)
        if getattr(self, "_tp", None) is None or self._tp[0] != key:
            self._tp = (key, TransposedPanels(self, pad, side_stream))
        return self._tp[1]


class DevicePairs:
    """CSR whose entries are interleaved 8-byte {int32 index, float32 value} pairs (``pairs``: int32 [nnz, 2]).
    Used for the transposed row panels: one scattered store per non-zero when building, one 8-byte load per
    entry when multiplying."""

    def __init__(self, indptr, pairs, shape):
        self.indptr, self.pairs = indptr, pairs
        self.shape = (int(shape[0]), int(shape[1]))
        self.sorted_indices = False

    @property
    def nnz(self) -> int:
        return int(self.pairs.shape[0])

    @property
    def indices(self):
        return self.pairs[:, 0]

    @property
    def data(self):
        return self.pairs[:, 1].view(torch.float32)


# ------------------------------------------------------------------------------------------
def tfidf_csr(A: DeviceCSR, log_tf=True, log_idf=True, log_tfidf=False, scale_factor=1e4,
              inplace_values=False, check_canonical=False, binarize=False) -> Optional[DeviceCSR]:
    """K1: two fused passes (reduce, apply).  Column sums are allreduced across cell shards
    (muon/_atac/preproc.py:92-119; multi-GPU plan SURVEY section 8e)."""
    require_cuda()
    n, d = A.shape
    sfx = "f32" if A.data.dtype == torch.float32 else "f64"
    flags = (TFIDF_LOG_TF if log_tf else 0) | (TFIDF_LOG_IDF if log_idf else 0) | (TFIDF_LOG_TFIDF if log_tfidf else 0)
    if scale_factor is None or scale_factor == 0 or scale_factor == 1:
        flags |= TFIDF_NO_SCALE
        scale_factor = 1.0
    if binarize:
        flags |= TFIDF_BINARIZE
    dev, dt = A.data.device, A.data.dtype
    row_sum = torch.empty(n, dtype=dt, device=dev)
    col_sum = torch.zeros(d, dtype=dt, device=dev)
    st = stream_ptr()
    status = torch.zeros(1, dtype=torch.int32, device=dev) if check_canonical else None
    call(f"mub_tfidf_reduce_{sfx}", ptr(A.indptr), ptr(A.indices), ptr(A.data), n, d, ptr(row_sum), ptr(col_sum),
         ptr(status), flags, st)
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


def csr_transpose(A: DeviceCSR, row0: int = 0, row1: Optional[int] = None, k0: Optional[int] = None,
                  k1: Optional[int] = None, pairs: bool = False):
    """Build the CSR of (A[row0:row1])^T on the device (count -> scan -> atomic-cursor fill).
    Row indices stored in the result are local to the panel (0 .. row1-row0).  ``k0``/``k1`` are the
    non-zero offsets of the row range if the caller already knows them (avoids a host sync)."""
    require_cuda()
    assert A.data.dtype == torch.float32, "transpose: float32 values only"
    n, d = A.shape
    row1 = n if row1 is None else row1
    dev = A.data.device
    st = stream_ptr()
    if k0 is None or k1 is None:
        if row0 == 0 and row1 == n:
            k0, k1 = 0, A.nnz
        else:
            k0, k1 = int(A.indptr[row0]), int(A.indptr[row1])
    nnz = k1 - k0
    t_count = torch.zeros(d + 1, dtype=torch.int64, device=dev)
    call("mub_csr_transpose_count", ptr(A.indices) + 4 * k0, nnz, d, ptr(t_count), st)
    t_indptr = torch.cumsum(t_count, 0)
    cursor = torch.empty(max(d, 1), dtype=torch.int64, device=dev)
    if pairs:
        t_pairs = torch.empty((nnz, 2), dtype=torch.int32, device=dev)
        call("mub_csr_transpose_fill_pairs", ptr(A.indptr) + 8 * row0, ptr(A.indices), ptr(A.data), row1 - row0, d, 0,
             ptr(t_indptr), ptr(cursor), ptr(t_pairs), st)
        return DevicePairs(t_indptr, t_pairs, (d, row1 - row0))
    t_indices = torch.empty(nnz, dtype=torch.int32, device=dev)
    t_data = torch.empty(nnz, dtype=torch.float32, device=dev)
    # indptr values are absolute offsets into indices/data, so only the indptr pointer is shifted;
    # slots claimed from `cursor` are panel-local, so the outputs are not
    call("mub_csr_transpose_fill", ptr(A.indptr) + 8 * row0, ptr(A.indices), ptr(A.data), row1 - row0, d, 0,
         ptr(t_indptr), ptr(cursor), ptr(t_indices), ptr(t_data), st)
    return DeviceCSR(t_indptr, t_indices, t_data, (d, row1 - row0), n_total=d, sorted_indices=False)


class TransposedPanels:
    """A^T as a list of row-panel transposes.  A^T Y gathers rows of Y (n x P); when n*P*4 bytes
    exceed what stays resident in L2 the gathers fall through to HBM, so the cells are cut into
    panels whose slice of Y fits in L2 and the products are accumulated panel by panel."""

    # bytes of the gathered operand per panel (B200 L2: 126 MB, shared with the CSR stream)
    L2_BUDGET = int(os.environ.get("MUON_B200_L2_BUDGET_MB", "32")) << 20

    def __init__(self, A: DeviceCSR, pad: int = 64, side_stream: bool = False):
        n = A.shape[0]
        rows = max(1, self.L2_BUDGET // (4 * pad))
        n_panels = max(1, -(-n // rows))
        bounds = [round(i * n / n_panels) for i in range(n_panels + 1)]
        self.shape = (A.shape[1], n)
        # non-zero offsets of the panel boundaries: one small D2H up front, no host syncs afterwards
        ks = A.indptr[torch.tensor(bounds, device=A.indptr.device)].tolist()
        self.ready = None
        if side_stream:
            # build on a second stream so that the first A*V product (main stream) overlaps it;
            # consumers call wait() before touching the panels
            main = torch.cuda.current_stream()
            side = torch.cuda.Stream()
            side.wait_stream(main)
            with torch.cuda.stream(side):
                self.panels = [(bounds[i], bounds[i + 1], csr_transpose(A, bounds[i], bounds[i + 1], ks[i], ks[i + 1], pairs=True))
                               for i in range(n_panels)]
                self.ready = side.record_event()
            for _, _, T in self.panels:           # memory is consumed on the main stream later on
                for t in (T.indptr, T.pairs):
                    t.record_stream(main)
        else:
            self.panels = [(bounds[i], bounds[i + 1], csr_transpose(A, bounds[i], bounds[i + 1], ks[i], ks[i + 1], pairs=True))
                           for i in range(n_panels)]

    def wait(self):
        if self.ready is not None:
            torch.cuda.current_stream().wait_event(self.ready)
            self.ready = None

    def spmm(self, Y: torch.Tensor, dynamic=True) -> torch.Tensor:
        self.wait()
        out = None
        for r0, r1, T in self.panels:
            if out is None:
                out = spmm(T, Y[r0:r1], dynamic=dynamic)
            else:
                spmm(T, Y[r0:r1], out=out, accumulate=True, dynamic=dynamic)
        return out


def spmm(A: DeviceCSR, B: torch.Tensor, out: Optional[torch.Tensor] = None, accumulate=False,
         dynamic=True, algo: Optional[str] = None) -> torch.Tensor:
    """K2/K3: C[n x P] (+)= A @ B[d x P];  P = B.shape[1] must be 32, 64 or 128.

    algo: "rowwarp" (v1: warp per row, B gathered from L2), "panel" (v2: B column panels staged in
    shared memory by TMA; needs sorted column indices) or None = $MUON_B200_SPMM or auto."""
    n, d = A.shape
    P = B.shape[1]
    assert B.shape[0] == d and B.dtype == torch.float32 and B.is_contiguous(), (B.shape, d, B.dtype)
    if out is None:
        assert not accumulate
        out = torch.empty((n, P), dtype=torch.float32, device=B.device)
    if isinstance(A, DevicePairs):
        counter = torch.zeros(1, dtype=torch.int64, device=B.device) if dynamic else None
        call("mub_spmm_csrp_f32", ptr(A.indptr), ptr(A.pairs), n, d, ptr(B), P, ptr(out), 1 if accumulate else 0,
             ptr(counter), stream_ptr())
        return out
    if algo is None:
        algo = os.environ.get("MUON_B200_SPMM", "auto")
    if algo == "auto":
        algo = "rowwarp"   # the panel kernel is correct but not yet faster
    if algo == "panel":
        if not A.sorted_indices:
            raise MuonB200Error("spmm(algo='panel') needs sorted column indices")
        call("mub_spmm_csr_panel_f32", ptr(A.indptr), ptr(A.indices), ptr(A.data), n, d, ptr(B), P, ptr(out),
             1 if accumulate else 0, stream_ptr())
        return out
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
# lsi() on the same AnnData does not pay the PCIe upload again.  Before the handle is used every
# value is compared with the host copy through a 64-bit checksum of the raw bits (one threaded pass
# over the host array, ~0.2 s for 24 GB, against >2 s for the upload), and the index arrays through
# a sampled fingerprint -- so editing X on the host between the two calls is safe.
_RESIDENT_ATTR = "_mub_resident"


def _bits_checksum_host(a: np.ndarray) -> int:
    """Sum of the array's bytes read as int64 words, modulo 2^64 (tail bytes zero-padded)."""
    if a.size == 0:
        return 0
    raw = np.ascontiguousarray(a).reshape(-1).view(np.uint8)
    n8 = raw.shape[0] // 8
    words = raw[: n8 * 8].view(np.int64)
    pool = _checksum_pool()
    parts = max(1, min(pool._max_workers, (n8 * 8) >> 22))
    bounds = [n8 * i // parts for i in range(parts + 1)]
    with np.errstate(over="ignore"):
        sums = list(pool.map(lambda i: int(np.add.reduce(words[bounds[i]:bounds[i + 1]], dtype=np.int64)), range(parts)))
    tail = bytes(raw[n8 * 8:]) + b"\0" * (8 - (raw.shape[0] - n8 * 8)) if raw.shape[0] > n8 * 8 else b""
    total = sum(sums) + (int.from_bytes(tail, "little", signed=True) if tail else 0)
    return total & 0xFFFFFFFFFFFFFFFF


def _bits_checksum_device(t: torch.Tensor) -> int:
    """Same checksum of a contiguous device (or CPU) tensor; chunked so that no large temporary is made."""
    if t.numel() == 0:
        return 0
    raw = t.contiguous().reshape(-1).view(torch.uint8)
    n8 = raw.numel() // 8
    words = raw[: n8 * 8].view(torch.int64)
    total = 0
    step = 1 << 28
    for off in range(0, n8, step):
        total += int(words[off:off + step].sum())          # int64 accumulation wraps like the host sum
    if raw.numel() > n8 * 8:
        tail = bytes(raw[n8 * 8:].cpu().numpy()) + b"\0" * (8 - (raw.numel() - n8 * 8))
        total += int.from_bytes(tail, "little", signed=True)
    return total & 0xFFFFFFFFFFFFFFFF


_SUM_POOL = None


def _checksum_pool():
    global _SUM_POOL
    if _SUM_POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        _SUM_POOL = ThreadPoolExecutor(max(1, min(32, os.cpu_count() or 1)))   # read-only pass: more threads than the memcpy
    return _SUM_POOL


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
            if not np.array_equal(dev.indices[idx].cpu().numpy(), np.asarray(host_matrix.indices[pos], dtype=np.int32)):
                return None
            rows = np.unique(np.linspace(0, dev.shape[0], num=min(dev.shape[0] + 1, 4096), dtype=np.int64))
            if not np.array_equal(dev.indptr[torch.from_numpy(rows).to(dev.data.device)].cpu().numpy(),
                                  np.asarray(host_matrix.indptr[rows], dtype=np.int64)):
                return None
            if _bits_checksum_device(dev.data) != _bits_checksum_host(host_matrix.data):
                return None
        return dev
    except Exception:
        return None


def knn_l2(X: torch.Tensor, k: int, Y: Optional[torch.Tensor] = None, algo: Optional[str] = None):
    """Exact k nearest neighbours in Euclidean distance (rows of Y closest to each row of X; Y defaults to X).
    Returns (indices int32 [n x k], distances float32 [n x k]), ascending, ties by lower index.  Groundwork for
    the WNN row (reference muon/_core/preproc.py:520-528).

    algo: "simt" (fp32 CUDA-core kernel), "tc" (tcgen05 TF32 candidate pass + fp32 re-rank, same result; falls back
    to "simt" -- another GPU kernel, not the host -- if a query has too many points inside the TF32 error band)
    or None = $MUON_B200_KNN, default "tc"."""
    require_cuda()
    Y = X if Y is None else Y
    assert X.dtype == torch.float32 and Y.dtype == torch.float32 and X.is_contiguous() and Y.is_contiguous()
    assert X.shape[1] == Y.shape[1]
    nq, d = X.shape
    idx = torch.empty((nq, k), dtype=torch.int32, device=X.device)
    dist = torch.empty((nq, k), dtype=torch.float32, device=X.device)
    algo = os.environ.get("MUON_B200_KNN", "tc") if algo is None else algo
    if algo == "tc" and d <= 128 and k <= 512:
        nbytes = int(load().mub_knn_l2_tc_workspace_bytes(nq, Y.shape[0], d))
        ws = torch.empty(max(nbytes, 4), dtype=torch.uint8, device=X.device)
        status = torch.zeros(1, dtype=torch.int32, device=X.device)
        call("mub_knn_l2_tc_f32", ptr(X), nq, ptr(Y), Y.shape[0], d, d, k, ptr(idx), ptr(dist), ptr(ws), nbytes,
             ptr(status), stream_ptr())
        st = int(status[0])
        if st & 4:
            raise MuonB200Error("knn_l2(algo='tc'): tcgen05 commit never arrived (internal error)")
        if st == 0:
            return idx, dist
        # st & 2: error band too crowded for some query -> exact SIMT kernel below
    elif algo not in ("simt", "tc"):
        raise ValueError(f"unknown kNN algorithm {algo!r}")
    call("mub_knn_l2_f32", ptr(X), nq, ptr(Y), Y.shape[0], d, d, k, ptr(idx), ptr(dist), stream_ptr())
    return idx, dist
