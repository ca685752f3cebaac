"""TEST INFRASTRUCTURE ONLY -- exact stand-ins for the third-party pieces muon's WNN driver imports.

``muon._core.preproc.neighbors`` (reference muon/_core/preproc.py:264-640) leans on packages that are neither
vendored in the reference nor installable here: ``umap`` (NN-descent search, fuzzy simplicial set),
``pynndescent`` (numba distance helpers) and ``scanpy`` (representation choice, connectivities wrapper).  To run the
reference's *own control flow* in the build container (round-2 oracle for the WNN row) this
module restates just those call targets:

* ``nearest_neighbors``  -- umap.umap_.nearest_neighbors.  The real one is approximate (NN-descent, random
  projection forest); this one is an exact brute-force search with the same return convention
  (indices, distances, forest=None), honouring callable numba metrics called as ``metric(x, y, *metric_kwds.values())``
  (how pynndescent invokes custom distances) on the float32-converted data.  Parity against the real package can
  therefore only be statistical; exact-vs-exact it is deterministic.
* ``euclidean``, ``sparse_euclidean``, ``sparse_jaccard`` -- pynndescent.distances / pynndescent.sparse, written as
  plain Python so that muon can ``njit`` them itself (preproc.py:46-48); index arrays need not be sorted.
* ``choose_representation`` -- scanpy.tools._utils._choose_representation for explicit ``use_rep`` (and X fallback).
* ``umap_connectivities`` -- scanpy.neighbors._connectivity.umap = umap.umap_.fuzzy_simplicial_set on precomputed
  kNN with set_op_mix_ratio=1, local_connectivity=1 (McInnes et al. 2018, Algorithm 2/3), returned as CSR.

Nothing here is imported by the product path.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


# ---- pynndescent.distances.euclidean / pynndescent.sparse.* (plain Python; muon njit-compiles them) -----------
def euclidean(x, y):
    result = 0.0
    for i in range(x.shape[0]):
        result += (x[i] - y[i]) ** 2
    return np.sqrt(result)


def sparse_jaccard(ind1, data1, ind2, data2):
    """1 - |A n B| / |A u B| on the index sets (values ignored, as in pynndescent, whose arr_union /
    arr_intersect do not assume sorted input either)."""
    ind1 = np.sort(ind1)
    ind2 = np.sort(ind2)
    i1 = 0
    i2 = 0
    n_equal = 0
    n1 = ind1.shape[0]
    n2 = ind2.shape[0]
    while i1 < n1 and i2 < n2:
        if ind1[i1] == ind2[i2]:
            n_equal += 1
            i1 += 1
            i2 += 1
        elif ind1[i1] < ind2[i2]:
            i1 += 1
        else:
            i2 += 1
    n_union = n1 + n2 - n_equal
    if n_union == 0:
        return 0.0
    return float(n_union - n_equal) / n_union


def sparse_euclidean(ind1, data1, ind2, data2):
    o1 = np.argsort(ind1)
    o2 = np.argsort(ind2)
    ind1, data1, ind2, data2 = ind1[o1], data1[o1], ind2[o2], data2[o2]
    i1 = 0
    i2 = 0
    acc = 0.0
    n1 = ind1.shape[0]
    n2 = ind2.shape[0]
    while i1 < n1 and i2 < n2:
        if ind1[i1] == ind2[i2]:
            d = data1[i1] - data2[i2]
            acc += d * d
            i1 += 1
            i2 += 1
        elif ind1[i1] < ind2[i2]:
            acc += data1[i1] * data1[i1]
            i1 += 1
        else:
            acc += data2[i2] * data2[i2]
            i2 += 1
    while i1 < n1:
        acc += data1[i1] * data1[i1]
        i1 += 1
    while i2 < n2:
        acc += data2[i2] * data2[i2]
        i2 += 1
    return np.sqrt(acc)


# ---- umap.umap_.nearest_neighbors: exact ---------------------------------------------------------------------
def nearest_neighbors(X, n_neighbors, metric, metric_kwds=None, angular=False, random_state=None,
                      low_memory=True, **_ignored):
    """Exact k nearest neighbours, same return convention as umap (indices, distances, search_forest).
    Ties are broken by index (stable sort), so a point with distance 0 to itself comes first."""
    metric_kwds = metric_kwds or {}
    n = X.shape[0]
    k = min(n_neighbors, n)
    if callable(metric):
        Xf = np.ascontiguousarray(X, dtype=np.float32)          # pynndescent converts the data to float32
        args = tuple(metric_kwds.values())
        D = np.empty((n, n), dtype=np.float64)
        for i in range(n):
            for j in range(n):
                D[i, j] = metric(Xf[i], Xf[j], *args)
    else:
        from scipy.spatial.distance import cdist
        Xd = X.toarray() if sp.issparse(X) else np.asarray(X)
        D = cdist(Xd, Xd, metric=metric, **metric_kwds)
    idx = np.argsort(D, axis=1, kind="stable")[:, :k]
    dist = np.take_along_axis(D, idx, axis=1)
    return idx.astype(np.int64), dist.astype(np.float32 if callable(metric) else D.dtype), None


# ---- scanpy.tools._utils._choose_representation -----------------------------------------------------------------
def choose_representation(adata, use_rep=None, n_pcs=None, silent=False):
    if use_rep is None or use_rep == "X":
        if use_rep is None and "X_pca" in adata.obsm and adata.n_vars > 50:
            X = adata.obsm["X_pca"]
            return X[:, :n_pcs] if n_pcs else X
        return adata.X
    if use_rep in adata.obsm:
        X = adata.obsm[use_rep]
        return X[:, :n_pcs] if n_pcs else X
    raise ValueError(f"Did not find {use_rep} in `.obsm.keys()`.")


# ---- scanpy.neighbors._connectivity.umap == umap.umap_.fuzzy_simplicial_set on a precomputed kNN graph ----------
SMOOTH_K_TOLERANCE = 1e-5
MIN_K_DIST_SCALE = 1e-3


def smooth_knn_dist(distances, k, n_iter=64, local_connectivity=1.0, bandwidth=1.0):
    target = np.log2(k) * bandwidth
    n = distances.shape[0]
    rho = np.zeros(n, dtype=np.float32)
    result = np.zeros(n, dtype=np.float32)
    mean_distances = np.mean(distances)
    for i in range(n):
        lo, hi, mid = 0.0, np.inf, 1.0
        ith = distances[i]
        nz = ith[ith > 0.0]
        if nz.shape[0] >= local_connectivity:
            index = int(np.floor(local_connectivity))
            interpolation = local_connectivity - index
            if index > 0:
                rho[i] = nz[index - 1]
                if interpolation > SMOOTH_K_TOLERANCE:
                    rho[i] += interpolation * (nz[index] - nz[index - 1])
            else:
                rho[i] = interpolation * nz[0]
        elif nz.shape[0] > 0:
            rho[i] = np.max(nz)
        for _ in range(n_iter):
            psum = 0.0
            for j in range(1, distances.shape[1]):
                d = distances[i, j] - rho[i]
                psum += np.exp(-(d / mid)) if d > 0 else 1.0
            if np.fabs(psum - target) < SMOOTH_K_TOLERANCE:
                break
            if psum > target:
                hi = mid
                mid = (lo + hi) / 2.0
            else:
                lo = mid
                if hi == np.inf:
                    mid *= 2
                else:
                    mid = (lo + hi) / 2.0
        result[i] = mid
        if rho[i] > 0.0:
            mean_ith = np.mean(ith)
            if result[i] < MIN_K_DIST_SCALE * mean_ith:
                result[i] = MIN_K_DIST_SCALE * mean_ith
        else:
            if result[i] < MIN_K_DIST_SCALE * mean_distances:
                result[i] = MIN_K_DIST_SCALE * mean_distances
    return result, rho


def umap_connectivities(knn_indices, knn_dists, *, n_obs, n_neighbors, set_op_mix_ratio=1.0,
                        local_connectivity=1.0):
    knn_dists = np.asarray(knn_dists, dtype=np.float32)
    sigmas, rhos = smooth_knn_dist(knn_dists, float(n_neighbors), local_connectivity=float(local_connectivity))
    n, k = knn_indices.shape
    rows = np.repeat(np.arange(n), k)
    cols = knn_indices.reshape(-1).astype(np.int64)
    missing = cols == -1                 # umap leaves rows/cols/vals of missing slots at 0 (entry (0, 0) += 0)
    rows[missing] = 0
    cols[missing] = 0
    vals = np.zeros(n * k, dtype=np.float32)
    for i in range(n):
        for j in range(k):
            if knn_indices[i, j] == -1:
                continue
            if knn_indices[i, j] == i:
                v = 0.0
            elif knn_dists[i, j] - rhos[i] <= 0.0 or sigmas[i] == 0.0:
                v = 1.0
            else:
                v = np.exp(-((knn_dists[i, j] - rhos[i]) / sigmas[i]))
            vals[i * k + j] = v
    result = sp.coo_matrix((vals, (rows, cols)), shape=(n_obs, n_obs))
    result.eliminate_zeros()
    transpose = result.transpose()
    prod = result.multiply(transpose)
    result = set_op_mix_ratio * (result + transpose - prod) + (1.0 - set_op_mix_ratio) * prod
    result.eliminate_zeros()
    return result.tocsr()
