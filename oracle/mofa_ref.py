"""TEST INFRASTRUCTURE ONLY (CPU oracle) -- MOFA+ variational inference restated in numpy float64.

``muon.tl.mofa`` (reference muon/_core/tools.py:290-708) densifies every modality
(tools.py:117-141), hands the arrays to the third-party **mofapy2** package
(``entry_point.build()/run()``, tools.py:583-585) and reads E[Z], E[W] and the variance
explained back (tools.py:604-701).  mofapy2 is NOT vendored in the reference, NOT pinned by it
(pyproject test-extra unpinned; CI installs git HEAD) and NOT installable here (no network).

PARITY UNPINNED against mofapy2's numerics: this file restates the published MOFA / MOFA+
coordinate-ascent updates (Argelaguet et al., Mol Syst Biol 2018, Appendix; Genome Biol 2020)
in the order mofapy2 schedules them (W -> Z -> AlphaW -> AlphaZ -> ThetaW -> Tau, ELBO every
iteration; SURVEY App. C, marked [recalled] there).  mofapy2's RNG stream cannot be reproduced,
so the reference's two golden factor values (tests/test_muon_tools.py:139-147) are unreachable;
what IS pinned is the reference's solver-agnostic structural test (tests/test_muon_tools.py:12-44:
planted 5 factors, 10 fitted, per-factor R^2 > 0.1 for the first five and <= 0.1 for the rest,
factors sorted by variance explained) -- see tests/test_oracle_mofa.py.  The CUDA path is then
checked against THIS restatement from an identical initial state.

What this restatement assumes about mofapy2, item by item (everything marked [recalled] is from memory of the
mofapy2 sources -- build_model/init_model.py, core/nodes/*.py, run/entry_point.py -- and of the papers'
supplementary methods; none of it can be checked in this container, which is why parity is called UNPINNED):

  item                     | value used here                                   | source
  -------------------------+---------------------------------------------------+------------------------------------------
  update order / iteration | Y(pseudo-data) -> W -> Z -> AlphaW -> AlphaZ ->    | MOFA 2018 Suppl. Methods "Inference: update
                           | ThetaW -> Tau, then ELBO                           | schedule"; mofapy2 entry_point default
                           |                                                   | schedule [recalled; mofapy2 may update Z
                           |                                                   | before W -- both are valid CAVI orders and
                           |                                                   | reach the same fixed points]
  init Z                   | N(0,1) draws from np.random.RandomState(seed)      | init_model.initZ(qmean="random") [recalled];
                           |                                                   | RNG stream not reproducible -> golden Z of
                           |                                                   | tests/test_muon_tools.py:139-147 unreachable
  init W                   | E[w] = 0, q(s=1) = 1                              | initW defaults [recalled]
  init alpha, tau          | E = 1 (qa = qb = 1)                               | initAlpha*/initTau(qa=1, qb=1) [recalled]
  prior alpha (W and Z)    | Gamma(a0 = 1e-3, b0 = 1e-3)                        | initAlphaW/Z(pa=1e-3, pb=1e-3) [recalled;
                           |                                                   | MOFA v1 used 1e-14]
  prior tau                | Gamma(1e-3, 1e-3)                                 | initTau(pa=1e-3, pb=1e-3) [recalled]
  prior theta              | Beta(1, 1); initial E ln theta ~ 0 (q(s=1)=1)      | initThetaW(pa=1, pb=1) [recalled]
  spike-and-slab W         | q(what, s): slab N(m, 1/a), spike N(0, 1/E[alpha]) | MOFA 2018 Suppl. "Spike-and-slab prior on the
                           | logit q(s=1) = E ln(theta/(1-theta)) + ln E[alpha]/2| weights" (reparametrised w = s * what)
                           |                 - ln(a)/2 + b^2/(2a)               |
  Z prior                  | N(0, 1/alphaZ_gk) per group (ard_factors) else N(0,1)| MOFA+ 2020 Methods "ARD prior on the factors"
  centring / scaling       | gaussian views: per-group means (center_groups) or | mofapy2 process_data [recalled]
                           | the global mean; scale_views: global std; scale_   |
                           | groups: per-group std; non-gaussian: neither       |
  non-gaussian views       | Seeger pseudo-data, fixed kappa (see               | Seeger & Bouchard 2012; MOFA 2018 Suppl.
                           | mofa_ref_general)                                  | "Non-gaussian likelihoods"; mofapy2
                           |                                                   | PseudoY_Seeger / Tau_Seeger nodes [recalled]
  convergence              | |dELBO| / |ELBO_0| * 100 < {fast 5e-4, medium 5e-5,| entry_point.set_train_options [recalled]
                           | slow 5e-6} (%), checked every iteration from the 2nd|
  factor order on output   | by total variance explained, descending            | implied by tests/test_muon_tools.py:42-44
  ELBO                     | every term verified against a dense term-by-term   | tests/test_oracle_mofa.py::
                           | evaluation from the posterior moments              | test_elbo_equals_bruteforce_dense_evaluation

Model (gaussian views m, one group, no missing values):
    y_nd = sum_k z_nk w_dk + eps,  eps ~ N(0, 1/tau_d)
    z_nk ~ N(0, 1/alphaZ_k)                      (ard_factors; fixed N(0,1) otherwise)
    w_dk = s_dk * what_dk,  what ~ N(0, 1/alphaW_k),  s ~ Bernoulli(theta_k)   (spike-and-slab)
    alpha, tau ~ Gamma(1e-3, 1e-3);  theta ~ Beta(1, 1)
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
from scipy.special import digamma, expit, gammaln

A0 = B0 = 1e-3          # Gamma hyper-prior of alpha and tau
TH_A0 = TH_B0 = 1.0     # Beta hyper-prior of theta
TOLERANCE = {"fast": 5e-4, "medium": 5e-5, "slow": 5e-6}   # % change of the ELBO


@dataclass
class MofaState:
    Z: np.ndarray                       # N x K   E[z]
    Zvar: np.ndarray                    # K       Var[z_nk] (same for all n: no missing values)
    W: list                             # per view D_m x K   E[s w]
    WW: list                            # per view D_m x K   E[(s w)^2]
    S: list                             # per view D_m x K   q(s=1)
    What2: list                         # per view D_m x K   E[what^2] (both branches), for AlphaW
    alphaW: list                        # per view (a[K], b[K])
    alphaZ: tuple                       # (a[K], b[K])
    theta: list                         # per view (a[K], b[K])
    tau: list                           # per view (a[D], b[D])
    elbo: list = field(default_factory=list)
    iterations: int = 0
    converged: bool = False


def preprocess(views, center=True, scale_views=False):
    """process_data of mofapy2 as reached from tools.py:283-287: per-feature centring of each view
    (one group), optional division by the view's global standard deviation.  Returns dense
    float64 arrays, the feature means (``intercepts``, tools.py:283-286) and the scale factors."""
    out, means, scales = [], [], []
    for Y in views:
        Y = np.asarray(Y.todense() if hasattr(Y, "todense") else Y, dtype=np.float64)  # tools.py:117-141
        # mofapy2's process_data centres gaussian views in either setting of center_groups: per group if set,
        # else with the mean over all samples ([recalled]; with one group both are the same mean)
        mu = Y.mean(axis=0)
        Yc = Y - mu
        sc = float(Yc.std()) if scale_views else 1.0
        out.append(Yc / sc)
        means.append(mu)
        scales.append(sc)
    return out, means, scales


def init_state(N, dims, K, seed=1, Z0=None):
    """Initial variational state.  Z ~ N(0,1) drawn from ``np.random.RandomState(seed)``
    (mofapy2 seeds numpy and draws Z at random [recalled]); everything else at prior means."""
    rs = np.random.RandomState(seed)
    Z = rs.normal(size=(N, K)) if Z0 is None else np.array(Z0, dtype=np.float64)
    st = MofaState(
        Z=Z, Zvar=np.ones(K),
        W=[np.zeros((D, K)) for D in dims], WW=[np.zeros((D, K)) for D in dims],
        S=[np.ones((D, K)) for D in dims], What2=[np.ones((D, K)) for D in dims],
        alphaW=[(np.ones(K), np.ones(K)) for _ in dims], alphaZ=(np.ones(K), np.ones(K)),
        theta=[(np.ones(K), np.full(K, 1e-8)) for _ in dims],      # E[theta] ~ 1 before sparsity kicks in
        tau=[(np.ones(D), np.ones(D)) for D in dims])
    return st


def _E_gamma(ab):
    a, b = ab
    return a / b, digamma(a) - np.log(b)


def _E_beta(ab):
    a, b = ab
    return digamma(a) - digamma(a + b), digamma(b) - digamma(a + b)


def update_W(P, ZZ, tau, alpha, lnth, ln1mth, W, spikeslab=True):
    """Spike-and-slab weights of one view, Gauss-Seidel over factors.
    P = Y^T E[Z] (D x K);  ZZ = E[Z^T Z] with E[z^2] on the diagonal (K x K)."""
    D, K = P.shape
    W = W.copy()
    WW = np.empty_like(W)
    S = np.empty_like(W)
    What2 = np.empty_like(W)
    for k in range(K):
        a = tau * ZZ[k, k] + alpha[k]
        cross = W @ ZZ[k, :] - W[:, k] * ZZ[k, k]
        b = tau * (P[:, k] - cross)
        m, v = b / a, 1.0 / a
        if spikeslab:
            logit = lnth[k] - ln1mth[k] + 0.5 * np.log(alpha[k]) - 0.5 * np.log(a) + 0.5 * b * b / a
            s = expit(logit)
        else:
            s = np.ones(D)
        S[:, k] = s
        W[:, k] = s * m
        WW[:, k] = s * (m * m + v)
        What2[:, k] = s * (m * m + v) + (1.0 - s) / alpha[k]
    return W, WW, S, What2


def update_Z(Q, GW, cW, alphaZ, Z):
    """Factors, Gauss-Seidel over k.  Q = sum_m Y (tau*W) (N x K);  GW = sum_m W^T diag(tau) W
    (off-diagonal use);  cW[k] = sum_m sum_d tau_d E[w_dk^2]."""
    N, K = Q.shape
    Z = Z.copy()
    var = 1.0 / (alphaZ + cW)
    for k in range(K):
        cross = Z @ GW[k, :] - Z[:, k] * GW[k, k]
        Z[:, k] = var[k] * (Q[:, k] - cross)
    return Z, var


def tau_b(ssq, P, ZZ, W, WW):
    """1/2 E||y_d - Z w_d||^2 per feature from sufficient statistics (no N x D pass)."""
    off = ZZ - np.diag(np.diag(ZZ))
    quad = np.einsum("dk,kj,dj->d", W, off, W) + WW @ np.diag(ZZ)
    return 0.5 * (ssq - 2.0 * np.sum(W * P, axis=1) + quad)


def _kl_gamma(ab, a0, b0):
    a, b = ab
    E, Eln = a / b, digamma(a) - np.log(b)
    lp = a0 * np.log(b0) - gammaln(a0) + (a0 - 1) * Eln - b0 * E
    lq = a * np.log(b) - gammaln(a) + (a - 1) * Eln - b * E
    return float(np.sum(lp - lq))


def _kl_beta(ab, a0, b0):
    a, b = ab
    Eln, Eln1 = _E_beta(ab)
    lp = gammaln(a0 + b0) - gammaln(a0) - gammaln(b0) + (a0 - 1) * Eln + (b0 - 1) * Eln1
    lq = gammaln(a + b) - gammaln(a) - gammaln(b) + (a - 1) * Eln + (b - 1) * Eln1
    return float(np.sum(lp - lq))


def elbo(st: MofaState, N, ard_weights=True, ard_factors=True, spikeslab=True):
    """Evidence lower bound with the 'tau trick' (valid right after the Tau update)."""
    total = 0.0
    K = st.Z.shape[1]
    for m in range(len(st.W)):
        Etau, Elntau = _E_gamma(st.tau[m])
        total += float(np.sum(0.5 * N * (Elntau - np.log(2 * np.pi)) - Etau * (st.tau[m][1] - B0)))
        total += _kl_gamma(st.tau[m], A0, B0)
        Ea, Elna = _E_gamma(st.alphaW[m]) if ard_weights else (np.ones(K), np.zeros(K))
        S = np.clip(st.S[m], 1e-300, 1.0)
        # E ln p(what|alpha) + entropy of q(what|s) (both branches)
        D = S.shape[0]
        # q(what | s=0) = N(0, 1/E[alpha]) is taken at the CURRENT E[alpha] in both terms that contain it (the
        # prior expectation below and the entropy), i.e. the bound is maximised over that free variance; What2
        # (frozen at the W update) only feeds the AlphaW update.  Checked term by term against a dense evaluation
        # in tests/test_oracle_mofa.py::test_elbo_equals_bruteforce_dense_evaluation.
        mW2 = st.WW[m] + (1.0 - st.S[m]) / Ea[None, :]
        lp = -0.5 * np.log(2 * np.pi) + 0.5 * Elna[None, :] - 0.5 * Ea[None, :] * mW2
        var1 = np.where(st.S[m] > 0, st.WW[m] / np.maximum(st.S[m], 1e-300) - (st.W[m] / np.maximum(st.S[m], 1e-300)) ** 2, 1.0)
        var1 = np.maximum(var1, 1e-300)
        ent = S * 0.5 * np.log(2 * np.pi * np.e * var1) + (1 - S) * 0.5 * np.log(2 * np.pi * np.e / Ea[None, :])
        total += float(np.sum(lp + ent))
        if spikeslab:
            lnth, ln1mth = _E_beta(st.theta[m])
            S1 = np.clip(1 - st.S[m], 1e-300, 1.0)
            total += float(np.sum(st.S[m] * lnth[None, :] + (1 - st.S[m]) * ln1mth[None, :]
                                  - st.S[m] * np.log(S) - (1 - st.S[m]) * np.log(S1)))
            total += _kl_beta(st.theta[m], TH_A0, TH_B0)
        if ard_weights:
            total += _kl_gamma(st.alphaW[m], A0, B0)
    Ea, Elna = _E_gamma(st.alphaZ) if ard_factors else (np.ones(K), np.zeros(K))
    Ez2 = (st.Z ** 2).sum(0) + N * st.Zvar
    total += float(np.sum(0.5 * N * Elna - 0.5 * Ea * Ez2 + 0.5 * N + 0.5 * N * np.log(st.Zvar)))
    if ard_factors:
        total += _kl_gamma(st.alphaZ, A0, B0)
    return total


def variance_explained(Yc, Z, W):
    """R^2 per factor for one (centred) view: 1 - SS(Y - z_k w_k^T) / SS(Y), in percent."""
    ss = float((Yc ** 2).sum())
    K = Z.shape[1]
    r2 = np.empty(K)
    for k in range(K):
        res = Yc - np.outer(Z[:, k], W[:, k])
        r2[k] = 1.0 - float((res ** 2).sum()) / ss
    return 100.0 * r2


def mofa_ref(views, n_factors=10, n_iterations=1000, center=True, scale_views=False, ard_weights=True,
             ard_factors=True, spikeslab_weights=True, convergence_mode="fast", seed=1, Z0=None,
             sort_factors=True, check_convergence=True):
    """Run CAVI.  ``views``: list of N x D_m arrays (dense or scipy sparse).  Returns a dict with
    Z (N x K), W (list of D_m x K), variance (list of K-vectors, %), elbo (list), state."""
    Ys, means, scales = preprocess(views, center, scale_views)
    N = Ys[0].shape[0]
    dims = [Y.shape[1] for Y in Ys]
    K = n_factors
    st = init_state(N, dims, K, seed, Z0)
    ssq = [(Y ** 2).sum(0) for Y in Ys]
    tol = TOLERANCE[convergence_mode]

    def stats_Z():
        ZZ = st.Z.T @ st.Z
        ZZ[np.diag_indices(K)] += N * st.Zvar
        return ZZ, [Y.T @ st.Z for Y in Ys]

    ZZ, P = stats_Z()
    for it in range(n_iterations):
        # ---- W (per view) -------------------------------------------------------------------
        for m in range(len(Ys)):
            Etau, _ = _E_gamma(st.tau[m])
            Ea = _E_gamma(st.alphaW[m])[0] if ard_weights else np.ones(K)
            lnth, ln1mth = _E_beta(st.theta[m])
            st.W[m], st.WW[m], st.S[m], st.What2[m] = update_W(P[m], ZZ, Etau, Ea, lnth, ln1mth, st.W[m],
                                                               spikeslab_weights)
        # ---- Z ------------------------------------------------------------------------------
        Q = np.zeros((N, K))
        GW = np.zeros((K, K))
        cW = np.zeros(K)
        for m in range(len(Ys)):
            Etau, _ = _E_gamma(st.tau[m])
            tw = st.W[m] * Etau[:, None]
            Q += Ys[m] @ tw
            GW += st.W[m].T @ tw
            cW += Etau @ st.WW[m]
        EaZ = _E_gamma(st.alphaZ)[0] if ard_factors else np.ones(K)
        st.Z, st.Zvar = update_Z(Q, GW, cW, EaZ, st.Z)
        ZZ, P = stats_Z()
        # ---- AlphaW, AlphaZ, ThetaW -------------------------------------------------------------
        for m in range(len(Ys)):
            if ard_weights:
                st.alphaW[m] = (np.full(K, A0 + 0.5 * dims[m]), B0 + 0.5 * st.What2[m].sum(0))
            if spikeslab_weights:
                s1 = st.S[m].sum(0)
                st.theta[m] = (TH_A0 + s1, TH_B0 + dims[m] - s1)
        if ard_factors:
            st.alphaZ = (np.full(K, A0 + 0.5 * N), B0 + 0.5 * np.diag(ZZ))
        # ---- Tau ------------------------------------------------------------------------------
        for m in range(len(Ys)):
            st.tau[m] = (np.full(dims[m], A0 + 0.5 * N), B0 + tau_b(ssq[m], P[m], ZZ, st.W[m], st.WW[m]))
        # ---- ELBO / convergence -----------------------------------------------------------------
        st.elbo.append(elbo(st, N, ard_weights, ard_factors, spikeslab_weights))
        st.iterations = it + 1
        if check_convergence and it >= 1:
            delta = 100.0 * abs((st.elbo[-1] - st.elbo[-2]) / st.elbo[0])
            if delta < tol:
                st.converged = True
                break
    var = [variance_explained(Ys[m], st.Z, st.W[m]) for m in range(len(Ys))]
    order = np.arange(K)
    if sort_factors:
        order = np.argsort(-np.sum(var, axis=0), kind="stable")
    return {"Z": st.Z[:, order], "W": [w[:, order] for w in st.W], "variance": [v[order] for v in var],
            "elbo": st.elbo, "order": order, "state": st, "intercepts": means, "scales": scales}


# ======================================================================================================
# General form: groups of cells and missing values, written the textbook way with an explicit N x D mask
# per view (exactly how the reference feeds mofapy2 after tools.py:144-169 expanded the "union" of
# observations with NaN rows and tools.py:243-255 ordered the cells by group).  Independent of the
# sufficient-statistics formulation above and of the CUDA path, which makes it a cross-check for both.
# ======================================================================================================
def mofa_ref_general(views, groups=None, n_factors=10, n_iterations=1000, center_groups=True, scale_views=False,
                     scale_groups=False, ard_weights=True, ard_factors=True, spikeslab_weights=True,
                     convergence_mode="fast", seed=1, Z0=None, sort_factors=True, check_convergence=True,
                     likelihoods=None):
    """``views``: list of N x D_m float arrays with NaN = missing; ``groups``: length-N array of group labels
    (None = one group).  Returns Z, W, variance[view] -> array (G x K, %), elbo, ...

    ``likelihoods``: per view "gaussian" (default), "poisson" or "bernoulli" -- what mofapy2's
    ``guess_likelihoods`` picks for integer / binary data when muon passes ``likelihoods=None``
    (muon/_core/tools.py:272-280).  Non-gaussian views follow Seeger & Bouchard (AISTATS 2012) as used by MOFA
    (Argelaguet et al. 2018, Appendix 'Non-gaussian likelihoods') [recalled from mofapy2's Poisson_PseudoY /
    Bernoulli_PseudoY / Tau_Seeger nodes, not checkable here]: each iteration starts by replacing the view with
    gaussian PSEUDO-DATA around zeta = E[Z]E[W]^T at a FIXED precision kappa_d,
        poisson   : rate(z) = ln(1+e^z),  yhat = zeta - sigmoid(zeta) (1 - y / rate(zeta)) / kappa,  kappa_d = 1/4 + 0.17 max_n y_nd
        bernoulli : yhat = zeta - (sigmoid(zeta) - y) / kappa,                                        kappa   = 1/4
    after which W and Z are updated exactly as for a gaussian view with tau = kappa; there is no Tau update,
    no centring and no scaling for such a view, and its ELBO term is the likelihood at zeta (sum y ln rate - rate,
    resp. sum y zeta - ln(1+e^zeta))."""
    Ys = [np.array(Y.todense() if hasattr(Y, "todense") else Y, dtype=np.float64) for Y in views]
    N = Ys[0].shape[0]
    M = len(Ys)
    K = n_factors
    glab = np.zeros(N, dtype=np.int64) if groups is None else np.unique(np.asarray(groups), return_inverse=True)[1]
    G = int(glab.max()) + 1
    gsel = [glab == g for g in range(G)]
    liks = ["gaussian"] * M if likelihoods is None else ([likelihoods] * M if isinstance(likelihoods, str) else list(likelihoods))
    assert all(l in ("gaussian", "poisson", "bernoulli") for l in liks), liks
    obs, kappa = [None] * M, [None] * M
    masks, means, scales = [], [], []
    for m in range(M):
        mask = ~np.isnan(Ys[m])
        Y = np.where(mask, Ys[m], 0.0)
        if liks[m] != "gaussian":                            # no centring / scaling; fixed precision kappa
            obs[m] = Y.copy()
            kappa[m] = 0.25 + 0.17 * Y.max(0) if liks[m] == "poisson" else np.full(Y.shape[1], 0.25)
            Ys[m] = Y
            masks.append(mask)
            means.append(np.zeros((G, Y.shape[1])))
            scales.append(np.ones(G))
            continue
        mu = np.zeros((G, Y.shape[1]))
        mu_all = Y.sum(0) / np.maximum(mask.sum(0), 1)        # center_groups=False: mean over all groups ([recalled])
        for g in range(G):
            cnt = mask[gsel[g]].sum(0)
            mu[g] = Y[gsel[g]].sum(0) / np.maximum(cnt, 1)
        for g in range(G):
            Y[gsel[g]] -= (mu[g] if center_groups else mu_all) * mask[gsel[g]]
        sc = np.ones(G)
        if scale_views:
            sc[:] = np.sqrt((Y ** 2).sum() / mask.sum() - (Y.sum() / mask.sum()) ** 2)
        if scale_groups:
            for g in range(G):
                yy, mm = Y[gsel[g]], mask[gsel[g]]
                sc[g] = np.sqrt((yy ** 2).sum() / mm.sum() - (yy.sum() / mm.sum()) ** 2)
        for g in range(G):
            Y[gsel[g]] /= sc[g]
        Ys[m] = Y
        masks.append(mask)
        means.append(mu)
        scales.append(sc)
    dims = [Y.shape[1] for Y in Ys]
    rs = np.random.RandomState(seed)
    Z = rs.normal(size=(N, K)) if Z0 is None else np.array(Z0, dtype=np.float64)
    Zvar = np.ones((N, K))
    W = [np.zeros((D, K)) for D in dims]
    WW = [np.zeros((D, K)) for D in dims]
    S = [np.ones((D, K)) for D in dims]
    What2 = [np.ones((D, K)) for D in dims]
    alphaW = [(np.ones(K), np.ones(K)) for _ in dims]
    alphaZ = (np.ones((G, K)), np.ones((G, K)))
    theta = [(np.ones(K), np.full(K, 1e-8)) for _ in dims]
    tau = [(np.ones((G, D)), np.ones((G, D))) if liks[m] == "gaussian" else
           (np.tile(kappa[m], (G, 1)), np.ones((G, D))) for m, D in enumerate(dims)]
    tol = TOLERANCE[convergence_mode]
    elbos, converged, it = [], False, -1

    def _softplus(x):
        return np.logaddexp(0.0, x)

    for it in range(n_iterations):
        for m in range(M):                                   # ---- Y: pseudo-data of the non-gaussian views
            if liks[m] == "gaussian":
                continue
            zeta = Z @ W[m].T
            if liks[m] == "poisson":
                rate = np.maximum(_softplus(zeta), 1e-300)
                Ys[m] = np.where(masks[m], zeta - expit(zeta) * (1.0 - obs[m] / rate) / kappa[m][None, :], 0.0)
            else:
                Ys[m] = np.where(masks[m], zeta - (expit(zeta) - obs[m]) / kappa[m][None, :], 0.0)
        ZZd = Z ** 2 + Zvar
        for m in range(M):
            Etau = (tau[m][0] / tau[m][1])[glab] * masks[m]                     # N x D, 0 where missing
            Ea = alphaW[m][0] / alphaW[m][1] if ard_weights else np.ones(K)
            lnth, ln1mth = _E_beta(theta[m])
            for k in range(K):
                a = Etau.T @ ZZd[:, k] + Ea[k]
                res = Ys[m] - Z @ W[m].T + np.outer(Z[:, k], W[m][:, k])
                b = ((Etau * res) * Z[:, [k]]).sum(0)
                mm_, v = b / a, 1.0 / a
                s = expit(lnth[k] - ln1mth[k] + 0.5 * np.log(Ea[k]) - 0.5 * np.log(a) + 0.5 * b * b / a) \
                    if spikeslab_weights else np.ones_like(a)
                S[m][:, k], W[m][:, k] = s, s * mm_
                WW[m][:, k] = s * (mm_ * mm_ + v)
                What2[m][:, k] = WW[m][:, k] + (1.0 - s) / Ea[k]
        EaZ = (alphaZ[0] / alphaZ[1])[glab] if ard_factors else np.ones((N, K))
        prec = EaZ.copy()
        for m in range(M):
            prec += ((tau[m][0] / tau[m][1])[glab] * masks[m]) @ WW[m]
        Zvar = 1.0 / prec
        for k in range(K):
            b = np.zeros(N)
            for m in range(M):
                Etau = (tau[m][0] / tau[m][1])[glab] * masks[m]
                res = Ys[m] - Z @ W[m].T + np.outer(Z[:, k], W[m][:, k])
                b += (Etau * res) @ W[m][:, k]
            Z[:, k] = Zvar[:, k] * b
        ZZd = Z ** 2 + Zvar
        for m in range(M):
            if ard_weights:
                alphaW[m] = (np.full(K, A0 + 0.5 * dims[m]), B0 + 0.5 * What2[m].sum(0))
            if spikeslab_weights:
                s1 = S[m].sum(0)
                theta[m] = (TH_A0 + s1, TH_B0 + dims[m] - s1)
        if ard_factors:
            alphaZ = (np.stack([np.full(K, A0 + 0.5 * gsel[g].sum()) for g in range(G)]),
                      np.stack([B0 + 0.5 * ZZd[gsel[g]].sum(0) for g in range(G)]))
        for m in range(M):
            if liks[m] != "gaussian":
                continue
            E2 = (Ys[m] - Z @ W[m].T) ** 2 + ZZd @ WW[m].T - (Z ** 2) @ (W[m] ** 2).T
            ta = np.stack([A0 + 0.5 * masks[m][gsel[g]].sum(0) for g in range(G)])
            tb = np.stack([B0 + 0.5 * (E2[gsel[g]] * masks[m][gsel[g]]).sum(0) for g in range(G)])
            tau[m] = (ta, tb)
        # ---- ELBO (same terms as ``elbo`` above, per group / per cell) ---------------------------------
        tot = 0.0
        for m in range(M):
            if liks[m] == "gaussian":
                Etau, Elntau = _E_gamma(tau[m])
                nmg = np.stack([masks[m][gsel[g]].sum(0) for g in range(G)])
                tot += float(np.sum(0.5 * nmg * (Elntau - np.log(2 * np.pi)) - Etau * (tau[m][1] - B0)))
                tot += _kl_gamma(tau[m], A0, B0)
            else:
                zeta = Z @ W[m].T
                if liks[m] == "poisson":
                    rate = np.maximum(_softplus(zeta), 1e-300)
                    tot += float(np.sum(np.where(masks[m], obs[m] * np.log(rate) - rate, 0.0)))
                else:
                    tot += float(np.sum(np.where(masks[m], obs[m] * zeta - _softplus(zeta), 0.0)))
            Ea, Elna = _E_gamma(alphaW[m]) if ard_weights else (np.ones(K), np.zeros(K))
            Sc = np.clip(S[m], 1e-300, 1.0)
            lp = -0.5 * np.log(2 * np.pi) + 0.5 * Elna[None, :] - 0.5 * Ea[None, :] * (WW[m] + (1.0 - S[m]) / Ea[None, :])
            var1 = np.maximum(np.where(S[m] > 0, WW[m] / np.maximum(S[m], 1e-300) - (W[m] / np.maximum(S[m], 1e-300)) ** 2, 1.0), 1e-300)
            ent = Sc * 0.5 * np.log(2 * np.pi * np.e * var1) + (1 - Sc) * 0.5 * np.log(2 * np.pi * np.e / Ea[None, :])
            tot += float(np.sum(lp + ent))
            if spikeslab_weights:
                lnth, ln1mth = _E_beta(theta[m])
                S1 = np.clip(1 - S[m], 1e-300, 1.0)
                tot += float(np.sum(S[m] * lnth[None, :] + (1 - S[m]) * ln1mth[None, :] - S[m] * np.log(Sc) - (1 - S[m]) * np.log(S1)))
                tot += _kl_beta(theta[m], TH_A0, TH_B0)
            if ard_weights:
                tot += _kl_gamma(alphaW[m], A0, B0)
        if ard_factors:
            EaZg, ElnaZg = _E_gamma(alphaZ)
        else:
            EaZg, ElnaZg = np.ones((G, K)), np.zeros((G, K))
        tot += float(np.sum(0.5 * ElnaZg[glab] - 0.5 * EaZg[glab] * ZZd + 0.5 + 0.5 * np.log(Zvar)))
        if ard_factors:
            tot += _kl_gamma(alphaZ, A0, B0)
        elbos.append(tot)
        if check_convergence and it >= 1 and 100.0 * abs((elbos[-1] - elbos[-2]) / elbos[0]) < tol:
            converged = True
            break
    var = []
    for m in range(M):
        r2 = np.zeros((G, K))
        for g in range(G):
            mk = masks[m][gsel[g]]
            yy = Ys[m][gsel[g]]
            ss = float((yy ** 2 * mk).sum())
            for k in range(K):
                r2[g, k] = 100.0 * (1.0 - float((((yy - np.outer(Z[gsel[g], k], W[m][:, k])) ** 2) * mk).sum()) / ss)
        var.append(r2)
    order = np.arange(K)
    if sort_factors:
        order = np.argsort(-np.sum([v.sum(0) for v in var], axis=0), kind="stable")
    return {"Z": Z[:, order], "W": [w[:, order] for w in W], "variance": [v[:, order] for v in var], "elbo": elbos,
            "order": order, "iterations": it + 1, "converged": converged, "intercepts": means, "scales": scales}
