"""``mu.atac.tl`` -- ATAC tools.  Only ``lsi`` is on the hot path."""
from __future__ import annotations

import numpy as np

from .._containers import is_anndata, is_mudata


def lsi(data, scale_embeddings=True, n_comps=50, *, tol: float = 1e-5, seed: int = 0, return_info: bool = False):
    """Run Latent Semantic Indexing -- drop-in for ``muon.atac.tl.lsi``
    (reference muon/_atac/tools.py:29-71).

    Writes ``adata.obsm["X_lsi"]`` (n x k; z-scored left singular vectors when
    ``scale_embeddings``), ``adata.uns["lsi"]["stdev"]`` (singular values / sqrt(n-1)) and
    ``adata.varm["LSI"]`` (d x k right singular vectors), components in descending order,
    signs arbitrary (as with ARPACK).  Returns ``None``.

    The truncated SVD is a block Golub-Kahan-Lanczos iteration (``_lsi.py``) whose passes over
    the matrix are the CUDA SpMM kernels; ``tol`` (keyword-only, not in the reference) is the
    relative residual at which the k triplets are accepted -- the default makes singular values
    agree with ``scipy.sparse.linalg.svds`` to ~1e-7 and gap-resolved vectors to ~1e-5.
    """
    import torch

    from .. import _device, _dist
    from .._lib import phase
    from .._lsi import CsrOperator, truncated_svd

    if is_anndata(data):
        adata = data
    elif is_mudata(data) and "atac" in data.mod:
        adata = data.mod["atac"]
    else:
        raise TypeError("Expected AnnData or MuData object with 'atac' modality")

    X = adata.X
    n_comps = min(n_comps, X.shape[1])  # tools.py:50
    resident = isinstance(X, _device.DeviceCSR)
    if resident:
        A = X
        if A.data.dtype != torch.float32:
            A = A.with_data(A.data.to(torch.float32))
    else:
        import scipy.sparse as sp
        Xs = X if sp.issparse(X) else sp.csr_matrix(np.asarray(X))
        Xs = Xs.tocsr()  # row order inside a row is irrelevant to the SpMM kernels
        # tfidf() may have left a device twin of this matrix.  It is used only if EVERY host element still equals
        # what was downloaded; that check (a threaded read pass over the host arrays, GIL released) runs on a helper
        # thread while the device already works on the twin, and a mismatch discards that work for a fresh upload.
        cand = _device.resident_candidate(Xs)
        check = None
        if cand is not None:
            from concurrent.futures import ThreadPoolExecutor
            A = cand[0]
            pool = ThreadPoolExecutor(1)
            check = pool.submit(_device.resident_valid, Xs, cand[1])
            pool.shutdown(wait=False)
        else:
            A = _device.DeviceCSR.from_scipy(Xs, dtype=np.float32)
    n_total = A.n_total
    if n_comps >= min(n_total, A.shape[1]):  # svds' own requirement, scipy _svds.py:40-44
        if not resident and check is not None:
            check.result()
        raise ValueError(f"`k` must be an integer satisfying `0 < k < min(A.shape)` (k={n_comps})")

    P = _device.pad_width(min(n_comps + 8, 128) if n_comps + 8 <= 128 else n_comps)

    def solve(A):
        op = CsrOperator(A, P)
        return truncated_svd(op, n_comps, P, tol=tol, seed=seed)

    U, s, V, info = solve(A)
    if not resident:
        if check is not None and not check.result():          # the host matrix was edited after tfidf(): start over
            A._tp = None
            _device.release_resident(Xs)
            del U, s, V
            A = _device.DeviceCSR.from_scipy(Xs, dtype=np.float32)
            U, s, V, info = solve(A)
        # the matrix on the device is a hidden copy (fresh upload or the twin left by tfidf): do not leave the
        # 8 B/nnz transposed panels cached on it; a DeviceCSR the caller owns keeps them for the next call (mofa)
        A._tp = None
    if not info.converged:
        from warnings import warn
        warn(f"lsi: stopped after {info.passes} passes at relative residual {max(info.residuals):.1e} > tol={tol:g} "
             "(clustered singular values around the k-th component?); the leading components are still accurate")

    # post-processing of tools.py:60-65 on the device (moments allreduced over cell shards)
    with phase("lsi.post_and_d2h"):
        emb = U
        if scale_embeddings:
            mom = torch.stack([U.sum(0, dtype=torch.float64), (U.to(torch.float64) ** 2).sum(0)])
            _dist.all_reduce_sum_(mom)
            mean = mom[0] / n_total
            std = (mom[1] / n_total - mean**2).clamp_min(0).sqrt()     # numpy std, ddof=0
            emb = ((U.to(torch.float64) - mean) / std).to(torch.float32)
        stdev = s / np.sqrt(n_total - 1)

        out_dtype = np.float32 if resident or X.dtype == np.float32 else np.float64
        adata.obsm["X_lsi"] = _device.to_host(emb.contiguous()).astype(out_dtype, copy=False)
        adata.uns["lsi"] = {"stdev": stdev.cpu().numpy().astype(out_dtype, copy=False)}
        adata.varm["LSI"] = _device.to_host(V.contiguous()).astype(out_dtype, copy=False)
    if return_info:
        return info
    return None
