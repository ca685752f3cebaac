"""GPU: mu.pp.neighbors (WNN) against the golden produced by the reference's own driver
(muon/_core/preproc.py:264-640) executed with exact-search stand-ins (tests/golden/make_golden.py::dump_wnn)."""
import numpy as np
import pytest
import scipy.sparse as sp

import muon_b200 as mu
from conftest import golden_csr, load_golden
from muon_b200._containers import SimpleAnnData, SimpleMuData

pytestmark = pytest.mark.gpu


def _inputs(z):
    mods = {}
    for name in ("rna", "atac"):
        R = z[f"rep_{name}"]
        ad = SimpleAnnData(np.zeros((R.shape[0], 4)))
        ad.obsm["X_rep"] = R
        ad.obsp["distances"] = golden_csr(z, f"knn_{name}")
        ad.uns["neighbors"] = {"params": {"n_neighbors": 15, "use_rep": "X_rep", "metric": "euclidean"},
                               "distances_key": "distances", "connectivities_key": "connectivities"}
        mods[name] = ad
    return SimpleMuData(mods)


def test_wnn_matches_reference_driver_golden(cuda):
    z = load_golden("wnn_small.npz")
    md = _inputs(z)
    assert mu.pp.neighbors(md, n_multineighbors=40) is None
    n = z["rep_rna"].shape[0]
    np.testing.assert_allclose(md.obs["rna:mod_weight"].to_numpy(), z["w_rna"], rtol=2e-5, atol=1e-7)
    np.testing.assert_allclose(md.obs["atac:mod_weight"].to_numpy(), z["w_atac"], rtol=2e-5, atol=1e-7)
    got = sp.csr_matrix(md.obsp["distances"])
    k1 = int(z["n_neighbors"]) + 1
    assert np.all(np.diff(got.indptr) == k1)
    assert np.all(np.diff(got.data.reshape(n, k1), axis=1) >= 0)          # ascending per row (numba argsort order)
    got.sort_indices()
    ref = golden_csr(z, "wnn_dist")
    same_rows = np.array([np.array_equal(got.indices[got.indptr[i]:got.indptr[i + 1]],
                                         ref.indices[ref.indptr[i]:ref.indptr[i + 1]]) for i in range(n)])
    assert same_rows.mean() > 0.98                                        # fp32 embeddings: a tie may flip at the k-th place
    rows = np.where(same_rows)[0]
    for i in rows[:: max(1, len(rows) // 50)]:
        np.testing.assert_allclose(got.data[got.indptr[i]:got.indptr[i + 1]], ref.data[ref.indptr[i]:ref.indptr[i + 1]],
                                   rtol=1e-5, atol=1e-7)
    C = sp.csr_matrix(md.obsp["connectivities"])
    Cr = golden_csr(z, "wnn_conn")
    assert (C != C.T).nnz == 0
    if same_rows.all():
        assert abs(C - Cr).max() < 1e-4
    p = md.uns["neighbors"]["params"]
    assert p["n_neighbors"] == int(z["n_neighbors"]) and p["n_multineighbors"] == 40 and p["method"] == "umap"


def test_wnn_argument_errors(cuda):
    z = load_golden("wnn_small.npz")
    md = _inputs(z)
    del md.mod["atac"].uns["neighbors"]
    with pytest.raises(ValueError):
        mu.pp.neighbors(md)
    with pytest.raises(TypeError):
        mu.pp.neighbors(np.ones((3, 3)))
    md = _inputs(z)
    with pytest.raises(NotImplementedError):
        mu.pp.neighbors(md, metric="cosine")
    out = mu.pp.neighbors(md, n_multineighbors=30, key_added="wnn", add_weights_to_modalities=True, copy=True)
    assert "wnn_distances" in out.obsp and "wnn" in out.uns and "mod_weight" in out.mod["rna"].obs.columns
    assert "wnn_distances" not in md.obsp
    with pytest.raises(NotImplementedError):                 # candidate-table limit, checked before any device work
        mu.pp.neighbors(md, n_multineighbors=800)


def test_wnn_sparse_representation_and_low_memory_blocks(cuda, monkeypatch):
    """A sparse representation gives the same graph as its dense form (Euclidean distances between sparse rows,
    reference preproc.py:425-447), and low_memory=True (queries in blocks) the same as low_memory=False."""
    from muon_b200 import _device
    z = load_golden("wnn_small.npz")
    dense = _inputs(z)
    mu.pp.neighbors(dense, n_multineighbors=40, low_memory=False)
    sparse = _inputs(z)
    sparse.mod["rna"].obsm["X_rep"] = sp.csr_matrix(sparse.mod["rna"].obsm["X_rep"])
    orig = _device.knn_l2
    monkeypatch.setattr(_device, "knn_l2", lambda X, k, Y=None, algo=None, query_chunk=None:
                        orig(X, k, Y, algo, query_chunk=64 if query_chunk else None))     # 150 cells -> 3 blocks
    mu.pp.neighbors(sparse, n_multineighbors=40, low_memory=True)
    for key in ("distances", "connectivities"):
        assert abs(sp.csr_matrix(dense.obsp[key]) - sp.csr_matrix(sparse.obsp[key])).max() < 1e-12
    np.testing.assert_array_equal(dense.obs["rna:mod_weight"].to_numpy(), sparse.obs["rna:mod_weight"].to_numpy())


def test_wnn_second_case_vs_numpy_restatement(cuda):
    """Different sizes / k / candidate counts than the golden: CUDA path vs oracle/wnn_ref.py."""
    from oracle import _third_party as tp
    from oracle.wnn_ref import wnn_ref
    rng = np.random.default_rng(3)
    N, k = 260, 10
    c = rng.integers(0, 4, N)
    reps = [rng.normal(size=(N, 12)) + 2.5 * np.eye(4)[c] @ rng.normal(size=(4, 12)),
            rng.normal(size=(N, 5)) + 2.0 * np.eye(4)[c] @ rng.normal(size=(4, 5))]
    reps = [(r / np.linalg.norm(r, axis=1, keepdims=True)).astype(np.float32).astype(np.float64) for r in reps]
    mods, graphs = {}, []
    for name, R in zip(("a", "b"), reps):
        idx, dist, _ = tp.nearest_neighbors(R, k, "euclidean")
        g = sp.csr_matrix((dist[:, 1:].reshape(-1), idx[:, 1:].reshape(-1), np.arange(0, N * (k - 1) + 1, k - 1)),
                          shape=(N, N))
        graphs.append(g)
        ad = SimpleAnnData(np.zeros((N, 2)))
        ad.obsm["X_emb"] = R
        ad.obsp["d"] = g
        ad.uns["nn"] = {"params": {"n_neighbors": k, "use_rep": "X_emb"}, "distances_key": "d"}
        mods[name] = ad
    md = SimpleMuData(mods)
    mu.pp.neighbors(md, n_multineighbors=25, n_bandwidth_neighbors=12, neighbor_keys={"a": "nn", "b": "nn"})
    ref = wnn_ref(reps, graphs, n_neighbors=k, n_bandwidth_neighbors=12, n_multineighbors=25)
    np.testing.assert_allclose(md.obs["a:mod_weight"].to_numpy(), ref["weights"][:, 0], rtol=2e-5, atol=1e-7)
    got = sp.csr_matrix(md.obsp["distances"])
    got.sort_indices()
    rd = ref["distances"].copy()
    rd.sort_indices()
    same = np.array([np.array_equal(got.indices[got.indptr[i]:got.indptr[i + 1]], rd.indices[rd.indptr[i]:rd.indptr[i + 1]])
                     for i in range(N)])
    assert same.mean() > 0.98
    i = int(np.where(same)[0][0])
    np.testing.assert_allclose(got.data[got.indptr[i]:got.indptr[i + 1]], rd.data[rd.indptr[i]:rd.indptr[i + 1]], rtol=1e-5)


def test_wnn_hub_cells_take_the_global_table_fallback(cuda):
    """A cell that is everybody's neighbour makes every candidate set N-1 > 1536 cells: the shared-memory tables
    overflow and the bandwidth kernel's second pass (hash tables in global memory) must give the same sigma /
    weights as the exhaustive numpy restatement."""
    from oracle import _third_party as tp
    from oracle.wnn_ref import wnn_ref
    rng = np.random.default_rng(11)
    N, k = 1700, 8
    c = rng.integers(0, 5, N)
    reps = [rng.normal(size=(N, 10)) + 2.5 * np.eye(5)[c] @ rng.normal(size=(5, 10)),
            rng.normal(size=(N, 6)) + 2.0 * np.eye(5)[c] @ rng.normal(size=(5, 6))]
    reps = [(r / np.linalg.norm(r, axis=1, keepdims=True)).astype(np.float32).astype(np.float64) for r in reps]
    mods, graphs = {}, []
    for m, (name, R) in enumerate(zip(("a", "b"), reps)):
        idx, dist, _ = tp.nearest_neighbors(R, k, "euclidean")
        idx, dist = idx[:, 1:].copy(), dist[:, 1:].copy()
        if m == 0:                                             # plant the hub: cell 0 replaces the farthest neighbour
            for i in range(1, N):
                if 0 not in idx[i]:
                    idx[i, -1] = 0
                    dist[i, -1] = np.linalg.norm(R[i] - R[0])
        g = sp.csr_matrix((dist.reshape(-1), idx.reshape(-1), np.arange(0, N * (k - 1) + 1, k - 1)), shape=(N, N))
        graphs.append(g)
        ad = SimpleAnnData(np.zeros((N, 2)))
        ad.obsm["X_emb"] = R
        ad.obsp["d"] = g
        ad.uns["nn"] = {"params": {"n_neighbors": k, "use_rep": "X_emb"}, "distances_key": "d"}
        mods[name] = ad
    md = SimpleMuData(mods)
    mu.pp.neighbors(md, n_multineighbors=30, neighbor_keys={"a": "nn", "b": "nn"})
    ref = wnn_ref(reps, graphs, n_neighbors=k, n_multineighbors=30)
    np.testing.assert_allclose(md.obs["a:mod_weight"].to_numpy(), ref["weights"][:, 0], rtol=5e-5, atol=1e-6)
    np.testing.assert_allclose(md.obs["b:mod_weight"].to_numpy(), ref["weights"][:, 1], rtol=5e-5, atol=1e-6)
