"""CPU: the TF-IDF oracle against the reference's own known answers and golden fixtures."""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import golden_csr, load_golden
from oracle.tfidf_ref import tfidf_closed_form, tfidf_ref


def test_reference_kat_dense():
    # reference tests/test_atac_preproc.py:11-20,47-52
    np.random.seed(2020)
    x = np.abs(np.random.normal(size=(4, 5)))
    out = tfidf_ref(x)
    assert "%.3f" % out[0, 0] == "4.659"
    assert "%.3f" % out[3, 0] == "4.770"
    out1 = tfidf_ref(x + 1)
    assert "%.3f" % out1[0, 0] == "2.856"
    assert sp.isspmatrix_csr(out)


def test_reference_kat_sparse():
    # reference tests/test_atac_preproc.py:55-64
    np.random.seed(2020)
    x = sp.rand(100, 10, density=0.2, format="csr")
    out = tfidf_ref(x)
    assert "%.3f" % out[10, 9] == "18.749"
    assert "%.3f" % out[50, 5] == "0.000"


def test_matches_unmodified_reference_dense():
    z = load_golden("tfidf_dense.npz")
    np.testing.assert_array_equal(tfidf_ref(z["x"]).toarray(), z["out"])
    np.testing.assert_array_equal(tfidf_ref(z["x"] + 1).toarray(), z["out_plus1"])


@pytest.mark.parametrize("name,kw", [
    ("default", {}), ("nolog_tf", {"log_tf": False}), ("nolog_idf", {"log_idf": False}),
    ("log_tfidf", {"log_tf": False, "log_idf": False, "log_tfidf": True}),
    ("noscale", {"scale_factor": 1}), ("sf100", {"scale_factor": 100.0})])
def test_matches_unmodified_reference_sparse(name, kw):
    z = load_golden("tfidf_sparse.npz")
    x = golden_csr(z, "x")
    ref = golden_csr(z, f"out_{name}")
    got = tfidf_ref(x, **kw)
    got.sort_indices()
    np.testing.assert_array_equal(got.indptr, ref.indptr)
    np.testing.assert_array_equal(got.indices, ref.indices)
    np.testing.assert_array_equal(got.data, ref.data)
    # closed form on raw arrays (what the CUDA kernels implement) == reference, canonical order
    vals, _, _ = tfidf_closed_form(x.indptr, x.indices, x.data, *x.shape, **kw)
    np.testing.assert_allclose(vals, ref.data, rtol=1e-14, atol=0)


def test_closed_form_float32_synth():
    z = load_golden("tfidf_synth.npz")
    x, ref = golden_csr(z, "x"), golden_csr(z, "out")
    assert x.dtype == np.float32 and ref.dtype == np.float32
    vals, rs, cs = tfidf_closed_form(x.indptr, x.indices, x.data, *x.shape)
    assert vals.dtype == np.float32
    np.testing.assert_array_equal(vals, ref.data)  # integer-valued counts: sums exact, bit-equal
    np.testing.assert_array_equal(rs, np.asarray(x.sum(1)).ravel())
    np.testing.assert_array_equal(cs, np.asarray(x.sum(0)).ravel())


def test_reversed_index_artifact_documented():
    # the reference output has per-row reversed index order; canonical form equal
    z = load_golden("tfidf_synth.npz")
    x = golden_csr(z, "x")
    out = tfidf_ref(x)
    assert not out.has_sorted_indices or out.nnz == 0
    out.sort_indices()
    np.testing.assert_array_equal(out.indices, x.indices)


def test_flag_conflict():
    with pytest.raises(AttributeError):
        tfidf_ref(np.ones((2, 2)), log_tfidf=True)
