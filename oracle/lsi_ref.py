"""TEST INFRASTRUCTURE ONLY (CPU oracle) -- LSI restated from the reference.

``muon._atac.tools.lsi`` (reference muon/_atac/tools.py:42-69) is one call to
``scipy.sparse.linalg.svds`` plus ten lines of post-processing.  The numerics live in
scipy (third-party; installed here: scipy 1.18.1, ARPACK ``eigsh`` on the implicit
X^T X, ``tol=0``, ``ncv=max(2k+1,20)``, random ``v0`` -- _svds.py:428-533).  scipy *is*
importable, so ``svds`` itself is the live oracle; this file restates only muon's
post-processing and offers sign/gap-aware comparison helpers.

Parity status: the reference's tests never touch ``lsi``, so parity is
pinned by us: (1) ``lsi_ref`` == the unmodified reference ``lsi`` run in the build
container (tests/golden/make_golden.py, committed fixtures), (2) the CUDA path vs
``lsi_ref`` within 1e-4 relative, sign-aligned and gap-aware (see ``compare_lsi``).
"""
from __future__ import annotations

import numpy as np
from scipy.sparse.linalg import svds


def lsi_ref(X, n_comps=50, scale_embeddings=True, dtype=None, seed=0):
    """Returns dict(X_lsi, stdev, LSI, svalues, U) -- same quantities the reference stores.

    tools.py:50      n_comps = min(n_comps, n_vars)
    tools.py:53      svds(X, k)
    tools.py:56-58   reverse to descending order
    tools.py:60-63   z-score U columns (numpy std, ddof=0)
    tools.py:65      stdev = s / sqrt(n_obs - 1)
    tools.py:67-69   obsm["X_lsi"], uns["lsi"]["stdev"], varm["LSI"] = V (d x k)
    ``dtype`` lets tests ask for the float64 "truth"; ``seed`` fixes ARPACK's start vector
    (the reference leaves it unseeded; results agree to ~1e-15 up to sign).
    """
    if dtype is not None:
        X = X.astype(dtype)
    k = min(int(n_comps), X.shape[1])
    u, s, vt = svds(X, k=k, rng=np.random.default_rng(seed))
    u, s, vt = u[:, ::-1], s[::-1], vt[::-1, :]
    emb = u
    if scale_embeddings:
        emb = (u - u.mean(axis=0)) / u.std(axis=0)
    stdev = s / np.sqrt(X.shape[0] - 1)
    return {"X_lsi": emb, "stdev": stdev, "LSI": vt.T, "svalues": s, "U": u}


def sign_align(A, B):
    """Flip columns of ``A`` so that each has a positive inner product with ``B``'s column."""
    sgn = np.sign(np.sum(A * B, axis=0))
    sgn[sgn == 0] = 1.0
    return A * sgn


def relative_gaps(s):
    """min(|s_i - s_{i+-1}|) / s_i -- how well-defined each singular vector is."""
    s = np.asarray(s, dtype=np.float64)
    g = np.full(s.shape, np.inf)
    d = np.abs(np.diff(s))
    g[:-1] = np.minimum(g[:-1], d)
    g[1:] = np.minimum(g[1:], d)
    return g / s


def subspace_sin(A, B):
    """Largest principal-angle sine between the column spaces of A and B (orthonormalised)."""
    qa, _ = np.linalg.qr(np.asarray(A, dtype=np.float64))
    qb, _ = np.linalg.qr(np.asarray(B, dtype=np.float64))
    c = np.linalg.svd(qa.T @ qb, compute_uv=False)
    return float(np.sqrt(max(0.0, 1.0 - min(c.min(), 1.0) ** 2)))


def compare_lsi(got, ref, rtol=1e-4, gap=5e-3, s_next=None):
    """Gap-aware comparison used by every LSI parity test (all arithmetic in float64).

    * singular values: relative error <= rtol for every component;
    * singular vectors U, V (unit columns): for components whose relative gap to both
      neighbours exceeds ``gap``, the sign-aligned 2-norm error ||u - u_ref|| <= rtol and the
      max-norm error <= rtol * max|u_ref| * 10 -- i.e. "1e-4 relative on factor/loading
      matrices".  A perturbation eps of the matrix rotates a vector by ~eps/gap, so vectors
      inside a tighter cluster are only compared as a subspace;
    * all components: largest principal-angle sine between the spans of the first m columns
      <= 10*rtol, m = the widest prefix that ends at a resolved gap.
    ``s_next`` is sigma_{k+1} if known (the last kept component's lower gap).
    Returns a dict of measured errors; raises AssertionError with a readable message.
    """
    s_ref = np.asarray(ref["svalues"], dtype=np.float64)
    s_got = np.asarray(got["svalues"], dtype=np.float64)
    k = s_ref.size
    err_s = np.abs(s_got - s_ref) / s_ref
    assert err_s.max() <= rtol, f"singular values differ: max rel err {err_s.max():.3e} > {rtol}"
    ext = np.concatenate([s_ref, [s_next if s_next is not None else 0.0]])
    g = relative_gaps(ext)[:k]
    resolved = g > gap
    out = {"sigma_rel": float(err_s.max()), "n_resolved": int(resolved.sum())}
    bound = np.abs(np.diff(ext)) / ext[:-1] > gap
    m = int(np.max(np.nonzero(bound)[0]) + 1) if bound.any() else 0
    out["prefix"] = m
    for name in ("U", "LSI"):
        R = np.asarray(ref[name], dtype=np.float64)
        G = sign_align(np.asarray(got[name], dtype=np.float64), R)
        e2 = np.linalg.norm(G - R, axis=0) / np.linalg.norm(R, axis=0)
        emax = np.abs(G - R).max(axis=0) / np.abs(R).max(axis=0)
        w2 = float(e2[resolved].max()) if resolved.any() else 0.0
        wm = float(emax[resolved].max()) if resolved.any() else 0.0
        assert w2 <= rtol, f"{name}: relative 2-norm error {w2:.3e} > {rtol} on a gap-resolved component"
        assert wm <= 10 * rtol, f"{name}: relative max-norm error {wm:.3e} > {10 * rtol}"
        out[f"{name}_err2"], out[f"{name}_errmax"] = w2, wm
        if m > 0:
            sn = subspace_sin(G[:, :m], R[:, :m])
            assert sn <= 10 * rtol, f"{name}: subspace sin {sn:.3e} over first {m} comps"
            out[f"{name}_subspace_sin"] = sn
    return out
