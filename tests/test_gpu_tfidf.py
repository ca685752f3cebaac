"""GPU parity: CUDA TF-IDF (through the C ABI) vs the CPU oracle and the reference's goldens."""
import numpy as np
import pytest
import scipy.sparse as sp

import muon_b200 as mu
from conftest import golden_csr, load_golden
from muon_b200._containers import SimpleAnnData, SimpleMuData
from muon_b200._synth import generate_host
from oracle.tfidf_ref import tfidf_closed_form, tfidf_ref

pytestmark = pytest.mark.gpu

# float32 log1p on the device vs numpy differs by <= 2 ulp; sums of integer counts are exact
RTOL32, RTOL64 = 1e-6, 1e-12


def _canon(m):
    m = sp.csr_matrix(m)
    m.sort_indices()
    return m


def _assert_parity(got, ref, rtol):
    got, ref = _canon(got), _canon(ref)
    np.testing.assert_array_equal(got.indptr, ref.indptr)      # bit-exact pattern
    np.testing.assert_array_equal(got.indices, ref.indices)
    assert got.dtype == ref.dtype
    np.testing.assert_allclose(got.data, ref.data, rtol=rtol, atol=0)


def test_reference_kat_dense(cuda):
    # reference tests/test_atac_preproc.py:16-20
    np.random.seed(2020)
    x = np.abs(np.random.normal(size=(4, 5)))
    adata = SimpleAnnData(x.copy())
    assert mu.atac.pp.tfidf(adata, log_tf=True, log_idf=True) is None
    assert sp.isspmatrix_csr(adata.X)
    assert "%.3f" % adata.X[0, 0] == "4.659"
    assert "%.3f" % adata.X[3, 0] == "4.770"
    z = load_golden("tfidf_dense.npz")
    np.testing.assert_allclose(adata.X.toarray(), z["out"], rtol=RTOL64)


def test_reference_kat_view_copy_inplace_layers(cuda):
    np.random.seed(2020)
    x = np.abs(np.random.normal(size=(4, 5)))
    base = SimpleAnnData(x.copy())
    view = base[:, :]                                   # test_tfidf_view
    mu.atac.pp.tfidf(view)
    assert "%.3f" % view.X[0, 0] == "4.659" and not view.is_view
    adata = base.copy()                                 # test_tfidf_copy
    orig = adata.X[0, 0]
    cp = mu.atac.pp.tfidf(adata, copy=True)
    assert adata.X[0, 0] == orig and "%.3f" % cp.X[0, 0] == "4.659"
    res = mu.atac.pp.tfidf(adata, inplace=False)        # test_tfidf_inplace
    assert adata.X[0, 0] == orig and "%.3f" % res[0, 0] == "4.659"
    mu.atac.pp.tfidf(adata, to_layer="new")             # test_tfidf_to_layer
    assert adata.X[0, 0] == orig and "%.3f" % adata.layers["new"][0, 0] == "4.659"
    with pytest.warns(UserWarning):
        mu.atac.pp.tfidf(adata, to_layer="new")
    a2 = base.copy()                                    # test_tfidf_from_layer
    a2.layers["counts"] = a2.X.copy() + 1
    a2.X = None
    mu.atac.pp.tfidf(a2, from_layer="counts")
    assert "%.3f" % a2.X[0, 0] == "2.856"
    md = SimpleMuData({"atac": base.copy(), "rna": SimpleAnnData(np.ones((4, 2)))})
    mu.atac.pp.tfidf(md)
    assert "%.3f" % md.mod["atac"].X[0, 0] == "4.659"


@pytest.mark.parametrize("name,kw", [
    ("default", {}), ("nolog_tf", {"log_tf": False}), ("nolog_idf", {"log_idf": False}),
    ("log_tfidf", {"log_tf": False, "log_idf": False, "log_tfidf": True}),
    ("noscale", {"scale_factor": 1}), ("sf100", {"scale_factor": 100.0})])
def test_reference_kat_sparse_all_flags(cuda, name, kw):
    z = load_golden("tfidf_sparse.npz")
    adata = SimpleAnnData(golden_csr(z, "x"))
    mu.atac.pp.tfidf(adata, **kw)
    if name == "default":
        assert "%.3f" % adata.X[10, 9] == "18.749" and "%.3f" % adata.X[50, 5] == "0.000"
    _assert_parity(adata.X, golden_csr(z, f"out_{name}"), RTOL64)


def test_float32_synth_vs_unmodified_reference(cuda):
    z = load_golden("tfidf_synth.npz")
    adata = SimpleAnnData(golden_csr(z, "x"))
    mu.atac.pp.tfidf(adata)
    _assert_parity(adata.X, golden_csr(z, "out"), RTOL32)


@pytest.mark.parametrize("dtype,rtol", [(np.float32, RTOL32), (np.float64, RTOL64)])
def test_config0_shape_vs_oracle(cuda, dtype, rtol):
    # BASELINE.json configs[0] at its real size: 10k cells x 30k peaks, 5 % (15 M nnz) -- correctness only
    X = generate_host(10000, 30000, 0.05, n_topics=16, seed=0).astype(dtype)
    assert X.shape == (10000, 30000) and 13e6 < X.nnz < 17e6
    ref = tfidf_ref(X)
    adata = SimpleAnnData(X.copy())
    mu.atac.pp.tfidf(adata)
    _assert_parity(adata.X, ref, rtol)
    vals, rs, cs = tfidf_closed_form(X.indptr, X.indices, X.data, *X.shape)
    np.testing.assert_allclose(_canon(adata.X).data, vals, rtol=rtol)


def test_integer_counts_and_edge_cases(cuda):
    rng = np.random.default_rng(0)
    X = sp.random(200, 50, 0.1, format="csr", random_state=1)
    X.data = rng.integers(1, 5, X.nnz).astype(np.int64)
    X = sp.csr_matrix(X)
    lil = X.tolil()
    lil[3, :] = 0          # empty row
    lil[:, 7] = 0          # empty column
    X = lil.tocsr().astype(np.int64)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = tfidf_ref(X)
    adata = SimpleAnnData(X.copy())
    mu.atac.pp.tfidf(adata)
    assert adata.X.dtype == np.float64          # integer counts -> float64
    _assert_parity(adata.X, ref, RTOL64)
    # explicit zeros and duplicates are canonicalised like the reference's matmul does (App. A.3)
    Xd = sp.csr_matrix((np.array([1.0, 2.0, 0.0, 3.0]), np.array([0, 0, 1, 2]), np.array([0, 3, 4])), shape=(2, 3))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = tfidf_ref(Xd)
    got = mu.atac.pp.tfidf(SimpleAnnData(Xd.copy()), inplace=False)
    _assert_parity(got, ref, RTOL64)
    # empty matrix
    E = sp.csr_matrix((5, 4), dtype=np.float32)
    got = mu.atac.pp.tfidf(SimpleAnnData(E), inplace=False)
    assert got.nnz == 0 and got.shape == (5, 4)


def test_device_resident_and_properties(cuda):
    """Size-independent properties at a larger shape: pattern untouched, positivity, scale
    invariance of TF under duplication of cells, resident path == host path."""
    X = generate_host(4000, 5000, 0.03, n_topics=8, seed=2)
    dev = mu.DeviceCSR.from_scipy(X)
    ad = SimpleAnnData(dev)
    mu.atac.pp.tfidf(ad)
    assert isinstance(ad.X, mu.DeviceCSR) and ad.X.data.data_ptr() != dev.data.data_ptr()
    out = ad.X.get()
    host = mu.atac.pp.tfidf(SimpleAnnData(X.copy()), inplace=False)
    np.testing.assert_array_equal(out.indices, X.indices)
    np.testing.assert_array_equal(out.data, host.data)
    assert np.all(out.data > 0) and np.all(np.isfinite(out.data))
    # stacking the matrix on itself doubles N and every column sum: idf unchanged -> same values
    X2 = sp.vstack([X, X]).tocsr()
    out2 = mu.atac.pp.tfidf(SimpleAnnData(X2), inplace=False)
    np.testing.assert_allclose(out2[:4000].data, out.data, rtol=1e-6)


def test_canonical_input_never_takes_the_host_fallback(cuda, monkeypatch):
    """The device-side canonical-form check must not flag canonical input (rows of every length
    class around the 32/128-wide segment boundaries); non-canonical input must be flagged."""
    from muon_b200.atac import pp
    calls = []
    orig = pp._canonical_csr
    monkeypatch.setattr(pp, "_canonical_csr", lambda X: (calls.append(1), orig(X))[1])
    rng = np.random.default_rng(0)
    rows = []
    d = 1000
    for L in list(range(0, 140)) + [255, 256, 257, 511, 640, 999]:
        cols = np.sort(rng.choice(d, L, replace=False))
        rows.append(sp.csr_matrix((rng.integers(1, 4, L).astype(np.float32), (np.zeros(L, int), cols)), shape=(1, d)))
    X = sp.vstack(rows).tocsr()
    X.sort_indices()
    got = mu.atac.pp.tfidf(SimpleAnnData(X.copy()), inplace=False)
    assert not calls
    _assert_parity(got, tfidf_ref(X), RTOL32)
    Xu = X.copy()                                        # same matrix, two entries of row 130 stored out of order
    a = X.indptr[130]
    Xu.indices[a:a + 2] = X.indices[a:a + 2][::-1].copy()
    Xu.data[a:a + 2] = X.data[a:a + 2][::-1].copy()
    Xu.has_sorted_indices = False
    got = mu.atac.pp.tfidf(SimpleAnnData(Xu), inplace=False)
    assert calls
    _assert_parity(got, tfidf_ref(X), RTOL32)


def test_binarize_and_fused_binarized_tfidf(cuda):
    from muon_b200 import _device
    X = generate_host(600, 800, 0.05, n_topics=5, seed=6)
    ad = SimpleAnnData(X.copy())
    mu.atac.pp.binarize(ad)                               # reference preproc.py:149: in-place on the host buffer
    assert set(np.unique(ad.X.data)) == {1.0}
    ref = tfidf_ref(ad.X)
    dev = mu.DeviceCSR.from_scipy(X)
    fused = _device.tfidf_csr(dev, binarize=True).get()    # one fused pass on raw counts
    _assert_parity(fused, ref, RTOL32)
    d2 = SimpleAnnData(mu.DeviceCSR.from_scipy(X))
    mu.atac.pp.binarize(d2)
    assert float(d2.X.data.min()) == 1.0 and float(d2.X.data.max()) == 1.0
    with pytest.raises(TypeError):
        mu.atac.pp.binarize(X)


@pytest.mark.parametrize("dt", [np.int32, np.uint16, np.int64])
def test_integer_dtypes_follow_reference_dtype_flow(cuda, dt):
    # integer counts -> float64 result, values equal to the reference to 1e-12
    X = generate_host(300, 250, 0.08, n_topics=4, seed=3)
    Xi = sp.csr_matrix((X.data.astype(dt), X.indices, X.indptr), shape=X.shape)
    ref = tfidf_ref(Xi)
    got = mu.atac.pp.tfidf(SimpleAnnData(Xi.copy()), inplace=False)
    assert got.dtype == np.float64 == ref.dtype
    _assert_parity(got, ref, RTOL64)
