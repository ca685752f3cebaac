"""GPU parity: mu.atac.tl.lsi (CUDA SpMM/Gram under the block Golub-Kahan driver) vs svds."""
import numpy as np
import pytest
import scipy.sparse as sp

import muon_b200 as mu
from conftest import golden_csr, load_golden
from muon_b200._containers import SimpleAnnData
from muon_b200._synth import generate_host
from oracle.lsi_ref import compare_lsi, lsi_ref, sign_align
from oracle.tfidf_ref import tfidf_ref

pytestmark = pytest.mark.gpu


def _got(adata, n):
    stdev = np.asarray(adata.uns["lsi"]["stdev"], dtype=np.float64)
    return {"svalues": stdev * np.sqrt(n - 1), "LSI": adata.varm["LSI"]}


def test_lsi_vs_unmodified_reference_golden(cuda):
    z = load_golden("lsi_synth.npz")
    X = golden_csr(z, "x")
    a = SimpleAnnData(X.copy())
    assert mu.atac.tl.lsi(a, n_comps=8) is None
    assert a.obsm["X_lsi"].shape == (600, 8) and a.varm["LSI"].shape == (500, 8)
    assert a.uns["lsi"]["stdev"].shape == (8,)
    np.testing.assert_allclose(a.uns["lsi"]["stdev"], z["stdev"], rtol=1e-5)
    b = SimpleAnnData(X.copy())
    mu.atac.tl.lsi(b, n_comps=8, scale_embeddings=False)
    got = {"svalues": b.uns["lsi"]["stdev"], "U": b.obsm["X_lsi"], "LSI": b.varm["LSI"]}
    ref = {"svalues": z["stdev"], "U": z["U"], "LSI": z["LSI"]}
    compare_lsi(got, ref, rtol=1e-4)
    # z-scored embeddings (tools.py:60-63): zero mean, unit std (ddof=0), equal to reference up to sign
    emb = a.obsm["X_lsi"]
    np.testing.assert_allclose(emb.mean(0), 0, atol=1e-6)
    np.testing.assert_allclose(emb.std(0), 1, rtol=1e-5)
    e = sign_align(emb.astype(np.float64), z["X_lsi"])
    assert np.abs(e - z["X_lsi"]).max() / np.abs(z["X_lsi"]).max() < 1e-3


@pytest.mark.parametrize("shape,k", [((3000, 2000), 15), ((1500, 4000), 30), ((2500, 1500), 60)])
def test_lsi_vs_svds_float64_truth(cuda, shape, k):
    n, d = shape
    X = tfidf_ref(generate_host(n, d, 0.05, n_topics=24, seed=n)).astype(np.float32)
    X.sort_indices()
    ref = lsi_ref(X, k + 1, scale_embeddings=False, dtype=np.float64)
    s_next = ref["svalues"][k]
    ref = {"svalues": ref["svalues"][:k], "U": ref["U"][:, :k], "LSI": ref["LSI"][:, :k]}
    a = SimpleAnnData(X.copy())
    info = mu.atac.tl.lsi(a, n_comps=k, scale_embeddings=False, return_info=True)
    assert info.converged
    got = {"svalues": a.uns["lsi"]["stdev"].astype(np.float64) * np.sqrt(n - 1), "U": a.obsm["X_lsi"],
           "LSI": a.varm["LSI"]}
    out = compare_lsi(got, ref, rtol=1e-4, s_next=s_next)
    assert out["sigma_rel"] < 1e-5
    assert a.obsm["X_lsi"].dtype == np.float32


def test_lsi_device_resident_pipeline_and_errors(cuda):
    C = generate_host(2000, 1500, 0.05, n_topics=10, seed=8)
    ad = SimpleAnnData(mu.DeviceCSR.from_scipy(C))
    mu.atac.pp.tfidf(ad)
    mu.atac.tl.lsi(ad, n_comps=10)
    host = SimpleAnnData(C.copy())
    mu.atac.pp.tfidf(host)
    mu.atac.tl.lsi(host, n_comps=10)
    np.testing.assert_allclose(ad.uns["lsi"]["stdev"], host.uns["lsi"]["stdev"], rtol=1e-5)
    with pytest.raises(ValueError):
        mu.atac.tl.lsi(SimpleAnnData(sp.random(30, 6, 0.5, format="csr", dtype=np.float32)), n_comps=50)
    # n_comps is clipped to n_vars (tools.py:50) and then rejected by the svds rule k < min(shape)
    small = SimpleAnnData(sp.random(200, 12, 0.5, format="csr", random_state=0, dtype=np.float32))
    mu.atac.tl.lsi(small, n_comps=5)
    assert small.obsm["X_lsi"].shape == (200, 5)


def test_lsi_dense_input_and_mudata(cuda):
    from muon_b200._containers import SimpleMuData
    X = tfidf_ref(generate_host(500, 300, 0.1, n_topics=5, seed=2)).astype(np.float64)
    ref = lsi_ref(X, 6, dtype=np.float64)
    ad = SimpleAnnData(X.toarray())                       # dense ndarray X, float64 -> float64 slots
    md = SimpleMuData({"atac": ad, "rna": SimpleAnnData(np.ones((500, 3)))})
    mu.atac.tl.lsi(md, n_comps=6)
    assert ad.obsm["X_lsi"].dtype == np.float64 and ad.varm["LSI"].shape == (300, 6)
    np.testing.assert_allclose(ad.uns["lsi"]["stdev"], ref["stdev"], rtol=1e-5)


def test_lsi_does_not_reuse_the_device_copy_after_a_host_edit(cuda):
    """tfidf() on a host matrix leaves a device twin behind for lsi(); one edited value on the host (anywhere,
    not only at the sampled positions) must invalidate it."""
    from muon_b200 import _device
    C = generate_host(3000, 900, 0.05, n_topics=8, seed=4)
    ad = SimpleAnnData(C.copy())
    mu.atac.pp.tfidf(ad)
    assert _device.recall_resident(ad.X) is not None
    nnz = ad.X.nnz
    sampled = set(np.linspace(0, nnz - 1, num=min(nnz, 4096), dtype=np.int64).tolist())
    j = next(i for i in range(12345, nnz) if i not in sampled)  # a position the sampled fingerprint does not look at
    ad.X.data[j] *= 1.5
    assert _device.recall_resident(ad.X) is None
    ad.X.data[:400] *= 30.0                                      # make the edit visible in the spectrum
    mu.atac.tl.lsi(ad, n_comps=6)
    fresh = SimpleAnnData(ad.X.copy())
    mu.atac.tl.lsi(fresh, n_comps=6)
    np.testing.assert_allclose(ad.uns["lsi"]["stdev"], fresh.uns["lsi"]["stdev"], rtol=1e-6)
