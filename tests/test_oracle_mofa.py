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


def _elbo_bruteforce(st, Ys, ard_weights=True, ard_factors=True, spikeslab=True):
    """ELBO = E_q[ln p(Y, Z, W, S, alpha, theta, tau)] - E_q[ln q] written out term by term from the posterior
    MOMENTS, with a dense N x D pass for the likelihood (no sufficient statistics, no 'tau trick') and the
    textbook closed forms of the Gamma / Beta / Gaussian KL divergences -- an independent evaluation of
    oracle.mofa_ref.elbo (MOFA: Argelaguet et al. 2018, Appendix 'Evidence lower bound'; MOFA+ 2020, Methods)."""
    from scipy.special import digamma, gammaln
    from oracle.mofa_ref import A0, B0, TH_A0, TH_B0
    N, K = st.Z.shape
    Ez, Ez2 = st.Z, st.Z ** 2 + st.Zvar[None, :]

    def kl_gamma(a, b, a0, b0):     # KL(Gamma(a,b) || Gamma(a0,b0)), rate parametrisation
        return np.sum((a - a0) * digamma(a) - gammaln(a) + gammaln(a0) + a0 * (np.log(b) - np.log(b0)) + a * (b0 - b) / b)

    def kl_beta(a, b, a0, b0):
        lnB = lambda x, y: gammaln(x) + gammaln(y) - gammaln(x + y)   # noqa: E731
        return np.sum(lnB(a0, b0) - lnB(a, b) + (a - a0) * digamma(a) + (b - b0) * digamma(b)
                      + (a0 - a + b0 - b) * digamma(a + b))

    total = 0.0
    for m, Y in enumerate(Ys):
        a, b = st.tau[m]
        Etau, Elntau = a / b, digamma(a) - np.log(b)
        Ew, Ew2 = st.W[m], st.WW[m]
        # E[(y_nd - sum_k z_nk w_dk)^2] = (y - E[zw])^2 + sum_k (E[z^2]E[w^2] - E[z]^2E[w]^2)
        mean = Ez @ Ew.T
        var = Ez2 @ Ew2.T - (Ez ** 2) @ (Ew ** 2).T
        total += np.sum(0.5 * Elntau[None, :] - 0.5 * np.log(2 * np.pi) - 0.5 * Etau[None, :] * ((Y - mean) ** 2 + var))
        total -= kl_gamma(a, b, A0, B0)
        if ard_weights:
            aa, ab = st.alphaW[m]
            Ea, Elna = aa / ab, digamma(aa) - np.log(ab)
            total -= kl_gamma(aa, ab, A0, B0)
        else:
            Ea, Elna = np.ones(K), np.zeros(K)
        S = st.S[m]
        # slab branch q(what | s=1) = N(m1, v1); spike branch q(what | s=0) = N(0, 1/E[alpha])
        with np.errstate(divide="ignore", invalid="ignore"):
            m1 = np.where(S > 0, Ew / S, 0.0)
            v1 = np.where(S > 0, Ew2 / S - m1 ** 2, 1.0)
        E_what2 = S * (m1 ** 2 + v1) + (1 - S) / Ea[None, :]
        total += np.sum(0.5 * Elna[None, :] - 0.5 * np.log(2 * np.pi) - 0.5 * Ea[None, :] * E_what2)       # E ln p(what|alpha)
        total += np.sum(S * 0.5 * np.log(2 * np.pi * np.e * np.maximum(v1, 1e-300))
                        + (1 - S) * 0.5 * np.log(2 * np.pi * np.e / Ea[None, :]))                          # H[q(what|s)]
        if spikeslab:
            ta, tb = st.theta[m]
            Elnth, Eln1mth = digamma(ta) - digamma(ta + tb), digamma(tb) - digamma(ta + tb)
            with np.errstate(divide="ignore", invalid="ignore"):
                Hs = -(np.where(S > 0, S * np.log(S), 0.0) + np.where(S < 1, (1 - S) * np.log1p(-S), 0.0))
            total += np.sum(S * Elnth[None, :] + (1 - S) * Eln1mth[None, :] + Hs)
            total -= kl_beta(ta, tb, TH_A0, TH_B0)
    if ard_factors:
        za, zb = st.alphaZ
        Ea, Elna = za / zb, digamma(za) - np.log(zb)
        total -= kl_gamma(za, zb, A0, B0)
    else:
        Ea, Elna = np.ones(K), np.zeros(K)
    total += np.sum(0.5 * Elna[None, :] - 0.5 * np.log(2 * np.pi) - 0.5 * Ea[None, :] * Ez2)                # E ln p(Z|alphaZ)
    total += N * np.sum(0.5 * np.log(2 * np.pi * np.e * st.Zvar))                                          # H[q(Z)]
    return float(total)


def test_elbo_equals_bruteforce_dense_evaluation():
    """Every term of the oracle's ELBO (likelihood through the 'tau trick', each KL) against the term-by-term dense
    evaluation above, on tiny problems, for the model variants the product supports."""
    from oracle.mofa_ref import preprocess
    rng = np.random.default_rng(5)
    z = rng.normal(size=(40, 3))
    views = [z @ rng.normal(size=(12, 3)).T + rng.normal(size=(40, 12)), z @ rng.normal(size=(7, 3)).T + rng.normal(size=(40, 7))]
    for kw in ({}, {"ard_weights": False}, {"ard_factors": False}, {"spikeslab_weights": False}, {"scale_views": True}):
        for T in (1, 4, 25):
            r = mofa_ref(views, n_factors=4, n_iterations=T, seed=2, check_convergence=False, sort_factors=False, **kw)
            Ys, _, _ = preprocess(views, True, kw.get("scale_views", False))
            want = _elbo_bruteforce(r["state"], Ys, kw.get("ard_weights", True), kw.get("ard_factors", True),
                                    kw.get("spikeslab_weights", True))
            np.testing.assert_allclose(r["elbo"][-1], want, rtol=1e-10, err_msg=f"{kw} after {T} iterations")


def test_non_gaussian_views_oracle_recovers_planted_factors():
    """Poisson / bernoulli views through Seeger pseudo-data (mofa_ref_general): the ELBO rises monotonically, the
    planted factors are found (and the superfluous ones switched off by the ARD prior), and for a gaussian-only model
    the general and the sufficient-statistics restatements still agree."""
    from oracle.mofa_ref import mofa_ref_general
    rng = np.random.default_rng(0)
    N, K = 200, 3
    Z = rng.normal(size=(N, K))
    W1 = rng.normal(size=(60, K)) * (rng.random((60, K)) < 0.5)
    W2 = rng.normal(size=(40, K)) * (rng.random((40, K)) < 0.5)
    Yp = rng.poisson(np.log1p(np.exp(Z @ W1.T + 0.5))).astype(float)
    Yb = (rng.random((N, 40)) < 1 / (1 + np.exp(-(Z @ W2.T)))).astype(float)
    Yg = Z @ rng.normal(size=(30, K)).T + rng.normal(size=(N, 30))
    r = mofa_ref_general([Yp, Yb, Yg], n_factors=5, n_iterations=40, likelihoods=["poisson", "bernoulli", "gaussian"],
                         check_convergence=False)
    e = np.asarray(r["elbo"])
    assert np.all(np.diff(e) > -1e-6 * np.abs(e[0]))
    tot = np.sum([v.sum(0) for v in r["variance"]], axis=0)
    assert (tot > 1.0).sum() == K                                   # 3 active factors, 2 pruned
    C = np.abs(np.corrcoef(np.c_[Z, r["Z"]].T)[:K, K:])
    assert C.max(1).min() > 0.6                                     # every planted factor has a fitted counterpart
    g = mofa_ref_general([Yg], n_factors=4, n_iterations=12, seed=2, check_convergence=False, sort_factors=False)
    s = mofa_ref([Yg], n_factors=4, n_iterations=12, seed=2, check_convergence=False, sort_factors=False)
    np.testing.assert_allclose(g["elbo"], s["elbo"], rtol=1e-12)
