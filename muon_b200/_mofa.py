"""``mu.tl.mofa`` on the GPU: MOFA+ coordinate-ascent VI over sparse, implicitly-centred modalities.

Reference: ``muon._core.tools.mofa`` (muon/_core/tools.py:290-708) marshals the MuData into dense
arrays (tools.py:117-141), trains mofapy2 (tools.py:583-585) and copies E[Z], E[W] and the variance
explained into ``obsm["X_mofa"]``, ``varm["LFs"]``, ``uns["mofa"]`` (tools.py:604-701).

Here the modalities stay sparse in HBM.  Per iteration and view there are exactly two passes over
the data -- ``P = Y^T E[Z]`` and ``Q = Y (tau * E[W])``, both the CSR SpMM kernel -- plus fused
per-row Gauss-Seidel kernels (csrc/mofa.cu) and K x K statistics from the Gram kernel.  Cells are
sharded across ranks; ``P``, the K x K Gram and K-vectors are sum-allreduced.
The update equations and their order are those of oracle/mofa_ref.py (see its header for the
parity status against mofapy2).

Supported: gaussian likelihood, one group, no missing values, ``use_var`` feature subsets,
``scale_views``, ``center_groups``, ARD on weights / factors, spike-and-slab weights, convergence
modes, ``copy``.  Not yet: ``groups_label``, ``use_obs`` (ragged observations), non-gaussian
likelihoods, SVI, MEFISTO smoothing, HDF5 ``outfile`` (h5py is not available) -- these raise
``NotImplementedError`` instead of silently doing something else.
"""
from __future__ import annotations

from functools import reduce
from typing import Optional
from warnings import warn

import numpy as np
import torch

from . import _device, _dist
from ._containers import SimpleMuData, is_anndata, is_mudata
from ._lib import call, ptr, stream_ptr

A0 = B0 = 1e-3
TH_A0 = TH_B0 = 1.0
TOLERANCE = {"fast": 5e-4, "medium": 5e-5, "slow": 5e-6}
f64 = torch.float64


class _View:
    """One modality on the device: CSR, its transpose, feature moments and variational state."""

    def __init__(self, A: _device.DeviceCSR, K: int, ld: int, n_total: int, center: bool, scale: bool):
        dev = A.data.device
        self.A, self.At = A, A.transpose_panels(ld)
        self.D = A.shape[1]
        D = self.D
        s1 = torch.empty(D, dtype=f64, device=dev)
        s2 = torch.empty(D, dtype=f64, device=dev)
        for i, (_, _, T) in enumerate(self.At.panels):
            t1, t2 = (s1, s2) if i == 0 else (torch.empty_like(s1), torch.empty_like(s2))
            call("mub_csr_row_stats_f32", ptr(T.indptr), ptr(T.data), D, ptr(t1), ptr(t2), stream_ptr())
            if i:
                s1 += t1
                s2 += t2
        _dist.all_reduce_sum_(s1)
        _dist.all_reduce_sum_(s2)
        self.mean = s1 / n_total                                   # intercepts, tools.py:283-286
        mu = self.mean if center else torch.zeros_like(self.mean)
        ssq = (s2 - 2.0 * mu * s1 + n_total * mu * mu).clamp_min(0)  # sum_n (y - mu)^2
        self.mu32 = mu.to(torch.float32) if center else None
        self.mu64 = mu
        self.inv_scale = 1.0
        if scale:                                                  # scale_views: global std of the centred view
            std = float(torch.sqrt(ssq.sum() / (float(n_total) * D)))
            self.inv_scale = 1.0 / std if std > 0 else 1.0
        self.ssq = ssq * self.inv_scale ** 2
        z = lambda: torch.zeros((D, ld), dtype=torch.float32, device=dev)  # noqa: E731
        self.W, self.WW, self.S, self.What2 = z(), z(), z(), z()
        self.S[:, :K] = 1.0
        self.What2[:, :K] = 1.0
        one = lambda n: torch.ones(n, dtype=f64, device=dev)       # noqa: E731
        self.alpha = (one(K), one(K))
        self.theta = (one(K), torch.full((K,), 1e-8, dtype=f64, device=dev))
        self.tau = (one(D), one(D))
        self.P = None                                              # Y^T E[Z], un-centred, allreduced


def _E_gamma(ab):
    a, b = ab
    return a / b, torch.special.digamma(a) - torch.log(b)


def _E_beta(ab):
    a, b = ab
    dg = torch.special.digamma
    return dg(a) - dg(a + b), dg(b) - dg(a + b)


def _kl_gamma(ab, a0, b0):
    a, b = ab
    E, Eln = _E_gamma(ab)
    lg = torch.special.gammaln
    a0t, b0t = torch.tensor(a0, dtype=f64, device=a.device), torch.tensor(b0, dtype=f64, device=a.device)
    lp = a0t * torch.log(b0t) - lg(a0t) + (a0 - 1) * Eln - b0 * E
    lq = a * torch.log(b) - lg(a) + (a - 1) * Eln - b * E
    return float((lp - lq).sum())


def _kl_beta(ab, a0, b0):
    a, b = ab
    Eln, Eln1 = _E_beta(ab)
    lg = torch.special.gammaln
    c = float(lg(torch.tensor(a0 + b0, dtype=f64)) - lg(torch.tensor(a0, dtype=f64)) - lg(torch.tensor(b0, dtype=f64)))
    lp = c + (a0 - 1) * Eln + (b0 - 1) * Eln1
    lq = lg(a + b) - lg(a) - lg(b) + (a - 1) * Eln + (b - 1) * Eln1
    return float((lp - lq).sum())


class MofaDevice:
    """CAVI state machine on the device (single group, gaussian, no missing values)."""

    def __init__(self, views, K, n_total, Z0, center=True, scale_views=False, ard_weights=True, ard_factors=True,
                 spikeslab_weights=True):
        self.K, self.N = int(K), int(n_total)
        self.ld = _device.pad_width(K)
        if K > 64:
            raise NotImplementedError("n_factors > 64 is not supported yet")
        self.opts = (ard_weights, ard_factors, spikeslab_weights)
        dev = views[0].data.device
        self.dev = dev
        self.views = [_View(A, K, self.ld, self.N, center, scale_views) for A in views]
        self.n_local = views[0].shape[0]
        self.Z = torch.zeros((self.n_local, self.ld), dtype=torch.float32, device=dev)
        self.Z[:, :K] = Z0.to(dev, torch.float32)
        self.zvar = torch.ones(K, dtype=f64, device=dev)
        self.alphaZ = (torch.ones(K, dtype=f64, device=dev), torch.ones(K, dtype=f64, device=dev))
        self.elbo = []
        self._stats_Z()

    # -- sufficient statistics of Z: two sparse passes (one per view) + Gram ---------------------------------
    def _stats_Z(self):
        K = self.K
        ZZ = _device.gram(self.Z, K, reduce=True)
        self.ZZ_mean = ZZ.clone()
        ZZ[range(K), range(K)] += self.N * self.zvar
        self.ZZ = ZZ.contiguous()
        self.zsum = _dist.all_reduce_sum_(self.Z[:, :K].sum(0, dtype=f64)).contiguous()
        for v in self.views:
            v.P = _dist.all_reduce_sum_(v.At.spmm(self.Z, dynamic=True))

    def step(self):
        K, ld, st = self.K, self.ld, stream_ptr()
        ard_w, ard_f, ss = self.opts
        onesK = torch.ones(K, dtype=f64, device=self.dev)
        # ---- W ----------------------------------------------------------------------------------------
        for v in self.views:
            Etau = _E_gamma(v.tau)[0].to(torch.float32).contiguous()
            Ea = (_E_gamma(v.alpha)[0] if ard_w else onesK).contiguous()
            lnth, ln1mth = _E_beta(v.theta)
            call("mub_mofa_update_w_f32", ptr(v.P), ptr(v.mu32), ptr(self.zsum), v.inv_scale, ptr(self.ZZ), ptr(Etau),
                 ptr(Ea), ptr(lnth.contiguous()), ptr(ln1mth.contiguous()), ptr(v.W), ptr(v.WW), ptr(v.S),
                 ptr(v.What2), v.D, ld, K, 1 if ss else 0, st)
            v.Etau32 = Etau
        # ---- Z ----------------------------------------------------------------------------------------
        Q = None
        GW = torch.zeros((K, K), dtype=f64, device=self.dev)
        cW = torch.zeros(K, dtype=f64, device=self.dev)
        qshift = torch.zeros(K, dtype=f64, device=self.dev)
        for v in self.views:
            tw = (v.W * (v.Etau32 * v.inv_scale)[:, None]).contiguous()        # D x ld operand of the SpMM
            if Q is None:
                Q = _device.spmm(v.A, tw, dynamic=False)
            else:
                _device.spmm(v.A, tw, out=Q, accumulate=True, dynamic=False)
            qshift += (v.mu64[:, None] * tw[:, :K].to(f64)).sum(0) if v.mu32 is not None else 0.0
            GW += _device.gram(v.W, K, weights=v.Etau32, reduce=False)
            cW += v.Etau32.to(f64) @ v.WW[:, :K].to(f64)
        EaZ = _E_gamma(self.alphaZ)[0] if ard_f else onesK
        self.zvar = (1.0 / (EaZ + cW)).contiguous()
        call("mub_mofa_update_z_f32", ptr(Q), ptr(qshift.contiguous()), ptr(GW.contiguous()), ptr(self.zvar),
             ptr(self.Z), self.n_local, ld, K, st)
        del Q
        self._stats_Z()
        # ---- AlphaW, ThetaW, AlphaZ ---------------------------------------------------------------------
        for v in self.views:
            if ard_w:
                v.alpha = (torch.full((K,), A0 + 0.5 * v.D, dtype=f64, device=self.dev),
                           B0 + 0.5 * v.What2[:, :K].sum(0, dtype=f64))
            if ss:
                s1 = v.S[:, :K].sum(0, dtype=f64)
                v.theta = (TH_A0 + s1, TH_B0 + v.D - s1)
        if ard_f:
            self.alphaZ = (torch.full((K,), A0 + 0.5 * self.N, dtype=f64, device=self.dev),
                           B0 + 0.5 * torch.diagonal(self.ZZ).clone())
        # ---- Tau ----------------------------------------------------------------------------------------
        for v in self.views:
            b = torch.empty(v.D, dtype=f64, device=self.dev)
            call("mub_mofa_tau_f32", ptr(v.P), ptr(v.mu32), ptr(self.zsum), v.inv_scale, ptr(self.ZZ), ptr(v.ssq),
                 ptr(v.W), ptr(v.WW), B0, ptr(b), v.D, ld, K, st)
            v.tau = (torch.full((v.D,), A0 + 0.5 * self.N, dtype=f64, device=self.dev), b)
        self.elbo.append(self._elbo())

    def _elbo(self):
        """Same expression as oracle/mofa_ref.py::elbo (tau trick; valid right after the Tau update)."""
        K, N = self.K, self.N
        ard_w, ard_f, ss = self.opts
        total = 0.0
        two_pi = 2.0 * np.pi
        for v in self.views:
            Etau, Elntau = _E_gamma(v.tau)
            total += float((0.5 * N * (Elntau - np.log(two_pi)) - Etau * (v.tau[1] - B0)).sum())
            total += _kl_gamma(v.tau, A0, B0)
            if ard_w:
                Ea, Elna = _E_gamma(v.alpha)
            else:
                Ea, Elna = torch.ones(K, dtype=f64, device=self.dev), torch.zeros(K, dtype=f64, device=self.dev)
            S0 = v.S[:, :K].to(f64)
            W, WW, W2 = v.W[:, :K].to(f64), v.WW[:, :K].to(f64), v.What2[:, :K].to(f64)
            S = S0.clamp(1e-300, 1.0)
            lp = -0.5 * np.log(two_pi) + 0.5 * Elna[None, :] - 0.5 * Ea[None, :] * W2
            Sd = S0.clamp_min(1e-300)
            var1 = torch.where(S0 > 0, WW / Sd - (W / Sd) ** 2, torch.ones_like(S0)).clamp_min(1e-300)
            ent = S * 0.5 * torch.log(two_pi * np.e * var1) + (1 - S) * 0.5 * torch.log(two_pi * np.e / Ea[None, :])
            total += float((lp + ent).sum())
            if ss:
                lnth, ln1mth = _E_beta(v.theta)
                S1 = (1 - S0).clamp(1e-300, 1.0)
                total += float((S0 * lnth[None, :] + (1 - S0) * ln1mth[None, :] - S0 * torch.log(S)
                                - (1 - S0) * torch.log(S1)).sum())
                total += _kl_beta(v.theta, TH_A0, TH_B0)
            if ard_w:
                total += _kl_gamma(v.alpha, A0, B0)
        if ard_f:
            Ea, Elna = _E_gamma(self.alphaZ)
        else:
            Ea, Elna = torch.ones(K, dtype=f64, device=self.dev), torch.zeros(K, dtype=f64, device=self.dev)
        Ez2 = torch.diagonal(self.ZZ)
        total += float((0.5 * N * Elna - 0.5 * Ea * Ez2 + 0.5 * N + 0.5 * N * torch.log(self.zvar)).sum())
        if ard_f:
            total += _kl_gamma(self.alphaZ, A0, B0)
        return total

    def variance_explained(self):
        """R^2 (%) per view and factor from sufficient statistics: 1 - SS(Y - z_k w_k^T)/SS(Y)."""
        K = self.K
        out = []
        zz = torch.diagonal(self.ZZ_mean)
        for v in self.views:
            W = v.W[:, :K].to(f64)
            P = (v.P[:, :K].to(f64) - (v.mu64[:, None] * self.zsum[None, :] if v.mu32 is not None else 0.0)) * v.inv_scale
            ss = v.ssq.sum()
            res = ss - 2.0 * (W * P).sum(0) + (W * W).sum(0) * zz
            out.append(100.0 * (1.0 - res / ss))
        return out


def run_mofa_device(views, n_factors, n_iterations, n_total, Z0, center=True, scale_views=False, ard_weights=True,
                    ard_factors=True, spikeslab_weights=True, convergence_mode="fast", check_convergence=True,
                    sort_factors=True, verbose=False):
    """Train on device-resident views; returns dict(Z, W (list), variance (list), elbo, iterations, converged)."""
    model = MofaDevice(views, n_factors, n_total, Z0, center, scale_views, ard_weights, ard_factors, spikeslab_weights)
    tol = TOLERANCE[convergence_mode]
    converged, it = False, 0
    for it in range(n_iterations):
        model.step()
        if verbose and _dist.rank() == 0:
            print(f"[mofa] iteration {it + 1}: ELBO = {model.elbo[-1]:.6e}")
        if check_convergence and it >= 1:
            delta = 100.0 * abs((model.elbo[-1] - model.elbo[-2]) / model.elbo[0])
            if delta < tol:
                converged = True
                break
    var = model.variance_explained()
    K = n_factors
    order = torch.arange(K, device=model.dev)
    if sort_factors:
        order = torch.argsort(-torch.stack(var).sum(0), stable=True)
    return {"Z": model.Z[:, :K][:, order], "W": [v.W[:, :K][:, order] for v in model.views],
            "variance": [x[order] for x in var], "elbo": model.elbo, "iterations": it + 1 if n_iterations else 0,
            "converged": converged, "order": order, "intercepts": [v.mean for v in model.views], "model": model}


# ------------------------------------------------------------------------------------------------------
def _to_device_view(X):
    import scipy.sparse as sp
    if isinstance(X, _device.DeviceCSR):
        return X if X.data.dtype == torch.float32 else X.with_data(X.data.to(torch.float32))
    Xs = X.tocsr() if sp.issparse(X) else sp.csr_matrix(np.asarray(X))
    return _device.DeviceCSR.from_scipy(Xs, dtype=np.float32)


def mofa(
    data,
    groups_label=None,
    use_raw: bool = False,
    use_layer: Optional[str] = None,
    use_var: Optional[str] = "highly_variable",
    use_obs: Optional[str] = None,
    likelihoods=None,
    n_factors: int = 10,
    scale_views: bool = False,
    scale_groups: bool = False,
    center_groups: bool = True,
    ard_weights: bool = True,
    ard_factors: bool = True,
    spikeslab_weights: bool = True,
    spikeslab_factors: bool = False,
    n_iterations: int = 1000,
    convergence_mode: str = "fast",
    use_float32: bool = False,
    gpu_mode: bool = False,
    gpu_device=None,
    svi_mode: bool = False,
    svi_batch_size: float = 0.5,
    svi_learning_rate: float = 1.0,
    svi_forgetting_rate: float = 0.5,
    svi_start_stochastic: int = 1,
    smooth_covariate: Optional[str] = None,
    smooth_warping: bool = False,
    smooth_kwargs=None,
    save_parameters: bool = False,
    save_data: bool = True,
    save_metadata: bool = True,
    seed: int = 1,
    outfile: Optional[str] = None,
    expectations=None,
    save_interrupted: bool = True,
    verbose: bool = False,
    quiet: bool = True,
    copy: bool = False,
):
    """Run Multi-Omics Factor Analysis -- drop-in for ``muon.tl.mofa`` (muon/_core/tools.py:290-708).

    Writes ``obsm["X_mofa"]`` (cells x factors), ``varm["LFs"]`` (features x factors, zero rows for
    features excluded by ``use_var``), ``uns["mofa"]["params"]`` and ``uns["mofa"]["variance"]``
    ({view: R^2 per factor in %}), factors ordered by total variance explained.  Returns a copy when
    ``copy=True``, else ``None``.  ``gpu_mode``/``gpu_device`` are accepted and ignored (always GPU).
    """
    if is_anndata(data):
        mdata = SimpleMuData({"data": data})            # tools.py:425-431
    elif is_mudata(data):
        mdata = data
    else:
        raise TypeError("Expected an MuData object")

    if use_var and (not hasattr(data.var, "columns") or use_var not in data.var.columns):
        warn(f"There is no column {use_var} in the provided object")     # tools.py:438-440
        use_var = None
    if is_mudata(data):
        common_obs = reduce(np.intersect1d, [np.asarray(v.obs_names) for v in mdata.mod.values()])
        if len(common_obs) != mdata.n_obs:
            if not use_obs:
                raise IndexError(
                    "Not all the observations are the same across modalities. Please run `mdata.intersect_obs()` "
                    "to subset the data or devise a strategy with `use_obs` ('union' or 'intersection')")
            elif use_obs not in ["union", "intersection"]:
                raise ValueError(f"Expected `use_obs` argument to be 'union' or 'intersection', not '{use_obs}'")
            raise NotImplementedError("use_obs='union'/'intersection' (ragged observations) is not supported yet")
        use_obs = None

    for flag, name in ((groups_label, "groups_label"), (svi_mode, "svi_mode"), (smooth_covariate, "smooth_covariate"),
                       (spikeslab_factors, "spikeslab_factors"), (scale_groups, "scale_groups"), (use_raw, "use_raw")):
        if flag:
            raise NotImplementedError(f"mofa(..., {name}=...) is not supported by the B200 path yet")
    lik = likelihoods
    if lik is not None:
        lik = [lik] * len(mdata.mod) if isinstance(lik, str) else list(lik)
        if any(l != "gaussian" for l in lik):
            raise NotImplementedError("only the gaussian likelihood is supported by the B200 path yet")
    if outfile is not None and not quiet:
        warn("outfile is ignored: the model is not written to HDF5 (h5py unavailable)")

    _device.require_cuda()
    # ---- marshal modalities (tools.py:104-176), sparse, no densification ---------------------------------
    mods = list(mdata.mod.keys())
    views, masks = [], []
    for m in mods:
        adata = mdata.mod[m]
        X = adata.layers[use_layer] if use_layer else adata.X
        mask = None
        if use_var and hasattr(adata.var, "columns") and use_var in adata.var.columns:
            mask = np.asarray(adata.var[use_var].astype(bool))
            if isinstance(X, _device.DeviceCSR):
                raise NotImplementedError("use_var with a device-resident matrix is not supported yet")
            import scipy.sparse as sp
            X = (X.tocsr() if sp.issparse(X) else np.asarray(X))[:, mask]
        elif use_var:
            mask = np.ones(adata.n_vars, dtype=bool)
        views.append(_to_device_view(X))
        masks.append(mask)
    n_local = views[0].shape[0]
    n_total = views[0].n_total
    row0 = views[0].row0

    rs = np.random.RandomState(seed)
    Z0 = torch.from_numpy(rs.normal(size=(n_total, n_factors))[row0:row0 + n_local])
    res = run_mofa_device(views, n_factors, n_iterations, n_total, Z0, center=center_groups, scale_views=scale_views,
                          ard_weights=ard_weights, ard_factors=ard_factors, spikeslab_weights=spikeslab_weights,
                          convergence_mode=convergence_mode, verbose=verbose and not quiet)

    if copy:
        data = data.copy()
    out_dt = np.float32 if use_float32 else np.float64
    data.obsm["X_mofa"] = res["Z"].cpu().numpy().astype(out_dt)                       # tools.py:628
    w = np.concatenate([W.cpu().numpy() for W in res["W"]], axis=0).astype(out_dt)
    if use_var:                                                                      # tools.py:636-641
        full = np.zeros((data.n_vars, w.shape[1]), dtype=out_dt)
        full[np.concatenate(masks)] = w
        data.varm["LFs"] = full
    else:
        data.varm["LFs"] = w
    data.uns["mofa"] = {
        "params": {
            "data": {"groups_label": groups_label, "use_raw": use_raw, "use_layer": use_layer,
                     "likelihoods": np.array(["gaussian"] * len(mods)), "features_subset": use_var,
                     "use_obs": use_obs, "scale_views": scale_views, "scale_groups": scale_groups,
                     "center_groups": center_groups, "use_float32": use_float32},
            "model": {"ard_factors": ard_factors, "ard_weights": ard_weights, "spikeslab_weights": spikeslab_weights,
                      "spikeslab_factors": spikeslab_factors, "n_factors": n_factors},
            "training": {"n_iterations": n_iterations, "convergence_mode": convergence_mode, "gpu_mode": gpu_mode,
                         "seed": seed},
        },
        "variance": {m: res["variance"][i].cpu().numpy() for i, m in enumerate(mods)},
    }
    data.uns["mofa"]["_b200"] = {"iterations": res["iterations"], "converged": res["converged"], "elbo": res["elbo"]}
    if copy:
        return data
    if not quiet:
        print("Saved MOFA embeddings in .obsm['X_mofa'] slot and their loadings in .varm['LFs'].")
    return None
