"""CPU emulation of the tensor-core kNN candidate filter (muon_b200/csrc/knn_tc.cu): TF32 truncation of the operands
(plain and 3xTF32 split), a(q,c) = |c|^2 - 2 S, the per-query buffer with the bucketed radix-select threshold
tau = (k-th smallest a seen so far, rounded up to a 2^12-ulp bucket) + slack, compaction when the buffer fills up.
Claim under test: with slack >= 2 x (error bound of a) no true k-nearest neighbour is ever
dropped, whatever the visiting order -- so the fp32 re-rank of the survivors equals the brute-force answer."""
import numpy as np
import pytest


def _trunc_tf32(x):
    return (np.ascontiguousarray(x, dtype=np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def _okey(a):
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
    return np.where(u & np.uint32(0x80000000), ~u, u | np.uint32(0x80000000)).astype(np.uint32)


def _from_okey(k):
    k = np.uint32(k)
    u = np.uint32(k & np.uint32(0x7FFFFFFF)) if (k & np.uint32(0x80000000)) else np.uint32(~k)
    return u.view(np.float32)


def _inner_products(Q, C, split):
    """What the tensor core accumulates: fp32 storage, inputs truncated to TF32, wide accumulation."""
    if not split:
        return _trunc_tf32(Q).astype(np.float64) @ _trunc_tf32(C).astype(np.float64).T
    qb, cb = _trunc_tf32(Q), _trunc_tf32(C)
    qs, cs = _trunc_tf32(Q - qb), _trunc_tf32(C - cb)       # the small parts are truncated by the hardware as well
    qb, cb, qs, cs = (m.astype(np.float64) for m in (qb, cb, qs, cs))
    return qb @ cb.T + qb @ cs.T + qs @ cb.T


def _survivors(a_row, k, slack, cap, tile, order):
    """Stream the candidates tile by tile like knn_tc_candidates_kernel does for one query."""
    keys, idx, tau = [], [], np.float32(np.inf)

    def compact():
        nonlocal keys, idx, tau
        if len(keys) < k:
            return
        pref = np.sort(_okey(np.asarray(keys, dtype=np.float32)) >> np.uint32(12))
        base = _from_okey((np.uint32(pref[k - 1]) << np.uint32(12)) | np.uint32(0xFFF))
        if not base <= np.finfo(np.float32).max:
            base = np.finfo(np.float32).max
        tau = np.float32(base + np.float32(slack))
        keep = [i for i, v in enumerate(keys) if v <= tau]
        keys, idx = [keys[i] for i in keep], [idx[i] for i in keep]

    overflow = False
    for t0 in range(0, len(order), tile):
        if len(keys) > cap - tile:
            compact()
            if len(keys) > cap - tile:
                overflow = True
                keys, idx = keys[: cap - tile], idx[: cap - tile]
        for j in order[t0:t0 + tile]:
            v = np.float32(a_row[j])
            if v <= tau:
                keys.append(v)
                idx.append(int(j))
    compact()
    return set(idx), overflow


@pytest.mark.parametrize("split,slack_rel", [(True, 1.2e-4), (False, 4 * 2.0 ** -8)])
@pytest.mark.parametrize("normalise", [True, False])
def test_no_true_neighbour_is_dropped(split, slack_rel, normalise):
    rng = np.random.default_rng(7 + split + 2 * normalise)
    n, d, k, cap, tile = 1500, 24, 12, 160, 32
    C = (rng.normal(size=(n, d)) + 2.0 * rng.normal(size=(6, d))[rng.integers(0, 6, n)]).astype(np.float32)
    if normalise:
        C /= np.linalg.norm(C, axis=1, keepdims=True)
    Q = C[:40]
    cn = (C.astype(np.float64) ** 2).sum(1).astype(np.float32)
    A = cn[None, :].astype(np.float64) - 2.0 * _inner_products(Q, C, split)
    exact = ((Q[:, None, :].astype(np.float64) - C[None, :, :].astype(np.float64)) ** 2).sum(-1)
    cmax = float(np.sqrt(cn.max()))
    n_over = 0
    for qi in range(Q.shape[0]):
        qn = float((Q[qi].astype(np.float64) ** 2).sum())
        # the bound the kernel relies on: |a - (exact - |q|^2)| <= slack / 2
        err = np.abs(A[qi] - (exact[qi] - qn)).max()
        slack = slack_rel * np.sqrt(qn) * cmax + 1e-5 * (qn + cmax * cmax)
        assert err <= slack / 2
        truth = set(np.argsort(exact[qi], kind="stable")[:k].tolist())
        for order in (np.arange(n), rng.permutation(n), np.argsort(exact[qi])[::-1]):   # incl. the worst case: far to near
            got, over = _survivors(A[qi], k, slack, cap, tile, order)
            n_over += over
            if not over:
                assert truth <= got
    assert n_over == 0 or not split       # the 3xTF32 band is narrow enough for this buffer; plain TF32 may overflow
