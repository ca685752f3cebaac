"""CPU: the MOFA oracle against the reference's solver-agnostic structural test."""
import numpy as np

from oracle.mofa_ref import mofa_ref


def _reference_dataset():
    # reference tests/test_muon_tools.py:13-23
    np.random.seed(1000)
    z = np.random.normal(size=(100, 5))
    w1 = np.random.normal(size=(90, 5))
    w2 = np.random.normal(size=(50, 5))
    e1 = np.random.normal(size=(100, 90))
    e2 = np.random.normal(size=(100, 50))
    return np.dot(z, w1.T) + e1, np.dot(z, w2.T) + e2


def test_reference_structural_kat():
    # reference tests/test_muon_tools.py:25-44: 10 fitted factors, only the first 5 explain > 10 %
    y1, y2 = _reference_dataset()
    r = mofa_ref([y1, y2], n_factors=10, n_iterations=1000, seed=1)
    y = np.concatenate([y1, y2], axis=1)
    Z, W = r["Z"], np.concatenate(r["W"], axis=0)
    r2 = [1 - np.sum((y - Z[:, [i]] @ W[:, [i]].T) ** 2) / np.sum(y ** 2) for i in range(10)]
    assert all(v > 0.1 for v in r2[:5])
    assert not any(v > 0.1 for v in r2[5:])
    assert r["state"].converged
    tot = np.sum(r["variance"], axis=0)
    assert np.all(np.diff(tot) <= 1e-9)          # sorted by variance explained


def test_elbo_is_monotone():
    y1, y2 = _reference_dataset()
    r = mofa_ref([y1, y2], n_factors=6, n_iterations=60, seed=3, check_convergence=False)
    e = np.asarray(r["elbo"])
    assert np.all(np.diff(e) > -1e-8 * np.abs(e[0]))
