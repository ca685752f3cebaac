"""GPU parity: mu.tl.mofa (sparse CUDA path) vs the float64 CPU restatement, same initial state."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

import muon_b200 as mu
from muon_b200._containers import SimpleAnnData, SimpleMuData
from muon_b200._mofa import run_mofa_device
from oracle.mofa_ref import mofa_ref, mofa_ref_general

pytestmark = pytest.mark.gpu


def _planted(N, dims, K, seed, noise=0.5):
    """Sparse non-negative views (about 19 % non-zero, non-integer values) with K planted factors."""
    rng = np.random.default_rng(seed)
    Z = rng.normal(size=(N, K))
    views = []
    for D in dims:
        W = rng.normal(size=(D, K)) * (rng.random((D, K)) < 0.5)
        Y = Z @ W.T + noise * rng.normal(size=(N, D)) - 1.0
        views.append(sp.csr_matrix(np.maximum(Y, 0).astype(np.float32)))
    return views


def _align(A, B):
    """Match columns of A to B by sign (same order expected)."""
    s = np.sign((A * B).sum(0))
    s[s == 0] = 1
    return A * s


@pytest.mark.parametrize("scale_views", [False, True])
def test_device_cavi_matches_oracle(cuda, scale_views):
    N, dims, K, T = 1500, [400, 250], 6, 15
    views = _planted(N, dims, 4, seed=7)
    Z0 = np.random.RandomState(1).normal(size=(N, K))
    ref = mofa_ref(views, n_factors=K, n_iterations=T, seed=1, Z0=Z0, scale_views=scale_views,
                   check_convergence=False, sort_factors=False)
    dv = [mu.DeviceCSR.from_scipy(v) for v in views]
    got = run_mofa_device(dv, K, T, N, torch.from_numpy(Z0), scale_views=scale_views, check_convergence=False,
                          sort_factors=False)
    # ELBO trajectory (float64 statistics on both sides) and the fp32 factor / loading matrices
    np.testing.assert_allclose(got["elbo"], ref["elbo"], rtol=1e-5)
    Zg, Zr = got["Z"].cpu().numpy().astype(np.float64), ref["Z"]
    scale = np.abs(Zr).max(0)
    active = ref["variance"][0] + ref["variance"][1] > 1.0       # factors that explain > 1 %
    assert active.sum() >= 3
    assert (np.abs(Zg - Zr).max(0) / scale)[active].max() < 1e-4
    for m in range(2):
        Wg, Wr = got["W"][m].cpu().numpy().astype(np.float64), ref["W"][m]
        assert (np.abs(Wg - Wr).max(0) / np.abs(Wr).max(0).clip(1e-12))[active].max() < 1e-4
        np.testing.assert_allclose(got["variance"][m].cpu().numpy()[active], ref["variance"][m][active], rtol=1e-3)
        np.testing.assert_allclose(got["intercepts"][m].cpu().numpy(), ref["intercepts"][m], rtol=1e-6, atol=1e-9)


def test_reference_structural_kat_through_api(cuda):
    # reference tests/test_muon_tools.py:12-54 (dense inputs, MuData and AnnData entry)
    np.random.seed(1000)
    z = np.random.normal(size=(100, 5))
    w1, w2 = np.random.normal(size=(90, 5)), np.random.normal(size=(50, 5))
    y1 = z @ w1.T + np.random.normal(size=(100, 90))
    y2 = z @ w2.T + np.random.normal(size=(100, 50))
    mdata = SimpleMuData({"y1": SimpleAnnData(y1), "y2": SimpleAnnData(y2)})
    with pytest.warns(UserWarning):                       # default use_var="highly_variable" is absent
        assert mu.tl.mofa(mdata, n_factors=10, quiet=True, verbose=False) is None
    assert mdata.obsm["X_mofa"].shape == (100, 10) and mdata.varm["LFs"].shape == (140, 10)
    y = np.concatenate([y1, y2], axis=1)
    r2 = [1 - np.sum((y - mdata.obsm["X_mofa"][:, [i]] @ mdata.varm["LFs"][:, [i]].T) ** 2) / np.sum(y ** 2)
          for i in range(10)]
    assert all(v > 0.1 for v in r2[:5]) and not any(v > 0.1 for v in r2[5:])
    assert set(mdata.uns["mofa"]["variance"]) == {"y1", "y2"}
    assert mdata.uns["mofa"]["params"]["model"]["n_factors"] == 10
    ad = SimpleAnnData(y1)
    with pytest.warns(UserWarning):
        mu.tl.mofa(ad, n_factors=10, quiet=True)
    assert "X_mofa" in ad.obsm and "LFs" in ad.varm


def test_api_semantics(cuda):
    views = _planted(300, [120, 80], 3, seed=3)
    a, b = SimpleAnnData(views[0]), SimpleAnnData(views[1])
    a.var["highly_variable"] = np.arange(120) % 3 != 0
    b.var["highly_variable"] = True
    md = SimpleMuData({"rna": a, "atac": b})
    md.var["highly_variable"] = np.concatenate([a.var["highly_variable"], b.var["highly_variable"]])
    out = mu.tl.mofa(md, n_factors=4, n_iterations=20, copy=True)
    assert out is not md and "X_mofa" not in md.obsm
    lfs = out.varm["LFs"]
    assert lfs.shape == (200, 4)
    assert np.all(lfs[:120][np.arange(120) % 3 == 0] == 0) and np.any(lfs[:120][np.arange(120) % 3 != 0] != 0)
    # the column only on the modalities (mudata would lift it into mdata.var): still honoured, no warning
    md_b = SimpleMuData({"rna": a, "atac": b})
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        out_b = mu.tl.mofa(md_b, n_factors=4, n_iterations=20, copy=True)
    # same fit (the transposed panels are filled in atomic-claim order: sums differ run to run in the last fp32 bits)
    np.testing.assert_allclose(out_b.varm["LFs"], lfs, rtol=2e-3, atol=1e-5)
    with pytest.raises(TypeError):
        mu.tl.mofa(np.ones((3, 3)))
    c = SimpleAnnData(views[1][:250])
    with pytest.raises(IndexError):
        mu.tl.mofa(SimpleMuData({"rna": a, "atac": c}), use_var=None)
    with pytest.raises(ValueError):
        mu.tl.mofa(md, use_var=None, likelihoods="negative_binomial")


def test_groups_and_union_match_general_oracle(cuda):
    """groups_label + use_obs="union" (cells missing from whole views) + scale_groups against the masked,
    textbook-form float64 oracle -- an independent formulation of the same model."""
    import pandas as pd
    N, dims, K = 300, [140, 90], 5
    views = _planted(N, dims, 3, seed=11)
    names = [f"c{i}" for i in range(N)]
    a = SimpleAnnData(views[0][:270], obs=names[:270])
    b = SimpleAnnData(views[1][30:], obs=names[30:])
    md = SimpleMuData({"rna": a, "atac": b})
    assert list(md.obs_names) == names
    groups = np.array(["g1"] * 120 + ["g2"] * 180)
    md.obs["grp"] = groups
    with pytest.raises(IndexError):
        mu.tl.mofa(md, use_var=None, n_factors=K)
    for kw in ({}, {"scale_groups": True}):
        mu.tl.mofa(md, use_var=None, n_factors=K, n_iterations=20, convergence_mode="slow", use_obs="union",
                   groups_label="grp", seed=3, **kw)
        Y1 = np.full((N, dims[0]), np.nan)
        Y1[:270] = views[0][:270].toarray()
        Y2 = np.full((N, dims[1]), np.nan)
        Y2[30:] = views[1][30:].toarray()
        ref = mofa_ref_general([Y1, Y2], groups=groups, n_factors=K, n_iterations=20, convergence_mode="slow",
                               seed=3, **kw)
        np.testing.assert_allclose(md.uns["mofa"]["_b200"]["elbo"], ref["elbo"], rtol=1e-5)
        Zg, Zr = md.obsm["X_mofa"], ref["Z"]
        act = np.sum([v.sum(0) for v in ref["variance"]], axis=0) > 2.0
        assert act.sum() >= 2
        assert (np.abs(Zg - Zr).max(0) / np.abs(Zr).max(0))[act].max() < 2e-4
        Wr = np.concatenate(ref["W"], axis=0)
        assert (np.abs(md.varm["LFs"] - Wr).max(0) / np.abs(Wr).max(0))[act].max() < 2e-4
        assert set(md.uns["mofa"]["variance"]["rna"]) == {"g1", "g2"}
        np.testing.assert_allclose(md.uns["mofa"]["variance"]["atac"]["g2"][act], ref["variance"][1][1][act], rtol=2e-3)


def test_reference_union_and_groups_api(cuda):
    # reference tests/test_muon_tools.py:56-87 (categorical groups on AnnData; obs union with dense/sparse mixes)
    import pandas as pd
    np.random.seed(1000)
    z = np.random.normal(size=(100, 5))
    y1 = z @ np.random.normal(size=(90, 5)).T + np.random.normal(size=(100, 90))
    y2 = z @ np.random.normal(size=(50, 5)).T + np.random.normal(size=(100, 50))
    ad = SimpleAnnData(y1.copy())
    ad.obs["ab"] = pd.Categorical(np.random.choice(["a", "b"], 100))
    with pytest.warns(UserWarning):
        mu.tl.mofa(ad, groups_label="ab", n_factors=10, quiet=True)
    assert ad.obsm["X_mofa"].shape == (100, 10) and ad.varm["LFs"].shape == (90, 10)
    assert np.all(np.isfinite(ad.obsm["X_mofa"]))
    for sparsity in (0, 1, 2):
        x1 = sp.csr_matrix(y1) if sparsity in (0, 2) else y1
        x2 = sp.csr_matrix(y2) if sparsity in (1, 2) else y2
        A1, A2 = SimpleAnnData(x1), SimpleAnnData(x2)
        md = SimpleMuData({"y1": A1[:-10], "y2": A2[10:]})
        with pytest.warns(UserWarning):
            mu.tl.mofa(md, n_factors=10, quiet=True, use_obs="union")
        assert md.obsm["X_mofa"].shape == (100, 10) and md.varm["LFs"].shape == (140, 10)
        md2 = SimpleMuData({"y1": A1[:-10], "y2": A2[10:]})
        with pytest.warns(UserWarning):
            mu.tl.mofa(md2, n_factors=10, quiet=True, use_obs="intersection")
        assert np.isnan(md2.obsm["X_mofa"][:10]).all() and np.isfinite(md2.obsm["X_mofa"][10:90]).all()


def test_device_resident_views_and_layers(cuda):
    views = _planted(400, [150, 100], 3, seed=5)
    a = SimpleAnnData(mu.DeviceCSR.from_scipy(views[0]))
    b = SimpleAnnData(None, layers={"norm": views[1]}, shape=views[1].shape)
    b.X = views[1] * 0                                     # X is ignored when use_layer is given
    a.layers["norm"] = a.X
    md = SimpleMuData({"rna": a, "atac": b})
    mu.tl.mofa(md, use_var=None, use_layer="norm", n_factors=4, n_iterations=15, likelihoods="gaussian")
    host = SimpleMuData({"rna": SimpleAnnData(views[0]), "atac": SimpleAnnData(views[1])})
    mu.tl.mofa(host, use_var=None, n_factors=4, n_iterations=15, likelihoods="gaussian")
    np.testing.assert_allclose(md.obsm["X_mofa"], host.obsm["X_mofa"], rtol=1e-4, atol=1e-5)
    assert md.uns["mofa"]["params"]["data"]["use_layer"] == "norm"


def _count_views(N, seed):
    rng = np.random.default_rng(seed)
    Z = rng.normal(size=(N, 3))
    W1 = rng.normal(size=(70, 3)) * (rng.random((70, 3)) < 0.5)
    W2 = rng.normal(size=(45, 3)) * (rng.random((45, 3)) < 0.5)
    W3 = rng.normal(size=(30, 3))
    Yp = rng.poisson(np.log1p(np.exp(Z @ W1.T + 0.5))).astype(np.float64)
    Yb = (rng.random((N, 45)) < 1 / (1 + np.exp(-(Z @ W2.T)))).astype(np.float64)
    Yg = Z @ W3.T + rng.normal(size=(N, 30))
    return Yp, Yb, Yg


def test_poisson_bernoulli_views_match_oracle(cuda):
    """Non-gaussian likelihoods (Seeger pseudo-data, dense views): guessed from the data exactly as mofapy2's
    guess_likelihoods does when the reference passes likelihoods=None (tools.py:272-280); the fit equals the
    float64 restatement from the same initial state."""
    N, K, T = 400, 5, 25
    Yp, Yb, Yg = _count_views(N, 0)
    md = SimpleMuData({"counts": SimpleAnnData(sp.csr_matrix(Yp)), "binary": SimpleAnnData(Yb), "cont": SimpleAnnData(Yg)})
    mu.tl.mofa(md, use_var=None, n_factors=K, n_iterations=T, convergence_mode="slow", seed=5)
    assert list(md.uns["mofa"]["params"]["data"]["likelihoods"]) == ["poisson", "bernoulli", "gaussian"]
    ref = mofa_ref_general([Yp, Yb, Yg], n_factors=K, n_iterations=T, convergence_mode="slow", seed=5,
                           likelihoods=["poisson", "bernoulli", "gaussian"])
    got_elbo = md.uns["mofa"]["_b200"]["elbo"]
    assert len(got_elbo) == len(ref["elbo"])
    np.testing.assert_allclose(got_elbo, ref["elbo"], rtol=2e-5)
    act = np.sum([v.sum(0) for v in ref["variance"]], axis=0) > 2.0
    assert act.sum() >= 2
    Zg, Zr = md.obsm["X_mofa"], ref["Z"]
    assert (np.abs(Zg - Zr).max(0) / np.abs(Zr).max(0))[act].max() < 5e-4
    Wr = np.concatenate(ref["W"], axis=0)
    assert (np.abs(md.varm["LFs"] - Wr).max(0) / np.abs(Wr).max(0))[act].max() < 5e-4
    # explicit likelihoods override the guess; the planted structure is found either way
    md2 = SimpleMuData({"counts": SimpleAnnData(sp.csr_matrix(Yp)), "cont": SimpleAnnData(Yg)})
    mu.tl.mofa(md2, use_var=None, n_factors=K, n_iterations=T, likelihoods=["gaussian", "gaussian"], seed=5)
    assert list(md2.uns["mofa"]["params"]["data"]["likelihoods"]) == ["gaussian", "gaussian"]
    # groups + poisson through the general path
    import pandas as pd
    md3 = SimpleMuData({"counts": SimpleAnnData(sp.csr_matrix(Yp)), "cont": SimpleAnnData(Yg)})
    grp = np.array(["a"] * 150 + ["b"] * 250)
    md3.obs["grp"] = grp
    mu.tl.mofa(md3, use_var=None, n_factors=K, n_iterations=15, convergence_mode="slow", groups_label="grp", seed=2)
    ref3 = mofa_ref_general([Yp, Yg], groups=grp, n_factors=K, n_iterations=15, convergence_mode="slow", seed=2,
                            likelihoods=["poisson", "gaussian"])
    np.testing.assert_allclose(md3.uns["mofa"]["_b200"]["elbo"], ref3["elbo"], rtol=2e-5)


def test_center_groups_false_uses_the_global_mean(cuda):
    """center_groups=False still centres every feature (mean over all groups): with one group the flag changes
    nothing, with two groups the fit equals the oracle's."""
    views = _planted(300, [90, 60], 3, seed=21)
    a = SimpleMuData({"rna": SimpleAnnData(views[0]), "atac": SimpleAnnData(views[1])})
    b = SimpleMuData({"rna": SimpleAnnData(views[0]), "atac": SimpleAnnData(views[1])})
    mu.tl.mofa(a, use_var=None, n_factors=4, n_iterations=12, convergence_mode="slow", center_groups=True)
    mu.tl.mofa(b, use_var=None, n_factors=4, n_iterations=12, convergence_mode="slow", center_groups=False)
    # (the transposed panels are filled in atomic-claim order: fp32 sums differ run to run in the last bits)
    np.testing.assert_allclose(a.uns["mofa"]["_b200"]["elbo"], b.uns["mofa"]["_b200"]["elbo"], rtol=1e-5)
    grp = np.array(["a"] * 100 + ["b"] * 200)
    c = SimpleMuData({"rna": SimpleAnnData(views[0]), "atac": SimpleAnnData(views[1])})
    c.obs["grp"] = grp
    mu.tl.mofa(c, use_var=None, n_factors=4, n_iterations=12, convergence_mode="slow", center_groups=False,
               groups_label="grp", seed=3)
    ref = mofa_ref_general([v.toarray() for v in views], groups=grp, n_factors=4, n_iterations=12,
                           convergence_mode="slow", center_groups=False, seed=3)
    np.testing.assert_allclose(c.uns["mofa"]["_b200"]["elbo"], ref["elbo"], rtol=1e-5)
