"""TEST INFRASTRUCTURE ONLY (CPU oracle) -- ATAC TF-IDF restated from the reference.

Restates the arithmetic of ``muon._atac.preproc.tfidf`` (reference
muon/_atac/preproc.py:92-119) with scipy/numpy only, so it runs where ``import muon``
cannot.  Pinned against the reference's own golden values
(tests/test_atac_preproc.py:19-20,52,63-64) in tests/test_oracle_tfidf.py, and against
the unmodified reference function executed in the build container
(tests/golden/make_golden.py -> tests/golden/tfidf_*.npz).

Nothing in the product path may import this module; it is the checker for the CUDA path
and the CPU arm of bench.py.
"""
from __future__ import annotations

import numpy as np
from scipy import sparse as sp


def tfidf_ref(counts, n_obs=None, log_tf=True, log_idf=True, log_tfidf=False, scale_factor=1e4):
    """Scipy restatement, same operation order and dtype flow as the reference.

    * sparse input: TF = diag(1/rowsum) @ counts  (preproc.py:92-96); dense input:
      counts / rowsum (preproc.py:97-99)
    * optional scale (skipped for None/0/1) and log1p          (preproc.py:101-104)
    * IDF = n_obs / colsum, optional log1p                     (preproc.py:106-108)
    * TF @ diag(IDF) (sparse) or csr(TF) @ csr(diag(IDF))      (preproc.py:110-114)
    * optional log1p of the product                            (preproc.py:116-117)
    Returns a ``scipy.sparse.csr_matrix`` exactly like the reference does (both branches).
    """
    if log_tfidf and (log_tf or log_idf):
        raise AttributeError("log_tfidf needs log_tf=False and log_idf=False")  # preproc.py:69-73
    n_obs = counts.shape[0] if n_obs is None else n_obs
    with np.errstate(divide="ignore", invalid="ignore"):
        if sp.issparse(counts):
            row_sum = np.asarray(counts.sum(axis=1)).reshape(-1)
            inv = sp.dia_matrix((1.0 / row_sum, 0), shape=(row_sum.size, row_sum.size))
            tf = inv @ counts
        else:
            row_sum = np.asarray(counts.sum(axis=1)).reshape(-1, 1)
            tf = counts / row_sum
        if scale_factor is not None and scale_factor != 0 and scale_factor != 1:
            tf = tf * scale_factor
        if log_tf:
            tf = np.log1p(tf)
        idf = np.asarray(n_obs / counts.sum(axis=0)).reshape(-1)
        if log_idf:
            idf = np.log1p(idf)
        if sp.issparse(tf):
            out = tf @ sp.dia_matrix((idf, 0), shape=(idf.size, idf.size))
        else:
            out = sp.csr_matrix(tf) @ sp.csr_matrix(np.diag(idf))
        if log_tfidf:
            out = np.log1p(out)
    return sp.csr_matrix(out)


def tfidf_closed_form(indptr, indices, data, n_rows, n_cols, n_obs=None, log_tf=True, log_idf=True,
                      log_tfidf=False, scale_factor=1e4):
    """Closed form on raw CSR arrays in canonical (input) index order, SURVEY App. A.1:

        out_ij = log1p(((1/r_i) * c_ij) * sf) * log1p(N / s_j)

    evaluated in that association order in the dtype of ``data`` (float32 stays float32).
    This is the form the CUDA kernels implement; tests check it is bit-equal to
    ``tfidf_ref`` after ``sort_indices()`` for canonical input.
    Returns (values, row_sum, col_sum).
    """
    data = np.asarray(data)
    dt = data.dtype if data.dtype.kind == "f" else np.dtype(np.float64)
    indptr = np.asarray(indptr, dtype=np.int64)
    indices = np.asarray(indices)
    n_obs = n_rows if n_obs is None else n_obs
    rows = np.repeat(np.arange(n_rows, dtype=np.int64), np.diff(indptr))
    c = data.astype(dt, copy=False)
    # float64 accumulation of exactly-representable partial sums is what numpy's pairwise
    # float32 reduction gives for integer-valued counts < 2**24
    row_sum = np.bincount(rows, weights=c, minlength=n_rows).astype(dt)
    col_sum = np.bincount(indices, weights=c, minlength=n_cols).astype(dt)
    with np.errstate(divide="ignore", invalid="ignore"):
        inv_r = (dt.type(1.0) / row_sum).astype(dt)
        tf = (inv_r[rows] * c).astype(dt)
        if scale_factor is not None and scale_factor != 0 and scale_factor != 1:
            tf = (tf * dt.type(scale_factor)).astype(dt)
        if log_tf:
            tf = np.log1p(tf)
        idf = (dt.type(n_obs) / col_sum).astype(dt)
        if log_idf:
            idf = np.log1p(idf)
        out = (tf * idf[indices]).astype(dt)
        if log_tfidf:
            out = np.log1p(out)
    return out, row_sum, col_sum
