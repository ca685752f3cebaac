"""Generates the committed golden fixtures by running the UNMODIFIED reference functions
(/root/reference/muon/_atac/preproc.py::tfidf and muon/_atac/tools.py::lsi) in the build
container through oracle/_refload.py (stubbed anndata/mudata/scanpy, nothing copied).

    python tests/golden/make_golden.py

Outputs (small, committed):
  tfidf_dense.npz   the reference test's dense 4x5 input (tests/test_atac_preproc.py:11-14) + outputs
  tfidf_sparse.npz  the reference test's sparse 100x10 input (:56-59) + output for several flag sets
  tfidf_synth.npz   300x400 synthetic counts (float32) + reference output
  lsi_synth.npz     reference lsi() on the TF-IDF of a 600x500 synthetic matrix, k=8
  signatures.json   parameter names/defaults of tfidf, binarize, lsi, mofa (ast, no import)
  wnn_small.npz     (round-2 groundwork) reference neighbors() with exact-search stand-ins, 150 cells x 2 modalities
"""
import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from muon_b200._containers import SimpleAnnData  # noqa: E402
from muon_b200._synth import generate_host  # noqa: E402
from oracle._refload import load_reference_lsi, load_reference_tfidf  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def csr_parts(X, prefix):
    X = sp.csr_matrix(X)
    X.sort_indices()
    return {f"{prefix}_indptr": X.indptr.astype(np.int64), f"{prefix}_indices": X.indices.astype(np.int32),
            f"{prefix}_data": X.data, f"{prefix}_shape": np.asarray(X.shape, dtype=np.int64)}


def main():
    import warnings
    warnings.simplefilter("ignore")
    tfidf = load_reference_tfidf()
    lsi = load_reference_lsi()

    # dense KAT of the reference tests
    np.random.seed(2020)
    x = np.abs(np.random.normal(size=(4, 5)))
    ad = SimpleAnnData(x.copy())
    tfidf(ad, log_tf=True, log_idf=True)
    ad2 = SimpleAnnData(None, layers={"counts": x.copy() + 1}, shape=x.shape)
    tfidf(ad2, from_layer="counts")
    np.savez(os.path.join(OUT, "tfidf_dense.npz"), x=x, out=ad.X.toarray(), out_plus1=ad2.X.toarray())

    # sparse KAT of the reference tests, all flag variants
    np.random.seed(2020)
    xs = sp.rand(100, 10, density=0.2, format="csr")
    parts = csr_parts(xs, "x")
    variants = {"default": {}, "nolog_tf": {"log_tf": False}, "nolog_idf": {"log_idf": False},
                "log_tfidf": {"log_tf": False, "log_idf": False, "log_tfidf": True},
                "noscale": {"scale_factor": 1}, "sf100": {"scale_factor": 100.0}}
    for name, kw in variants.items():
        a = SimpleAnnData(xs.copy())
        tfidf(a, **kw)
        parts.update(csr_parts(a.X, f"out_{name}"))
    np.savez(os.path.join(OUT, "tfidf_sparse.npz"), **parts)

    # synthetic float32 counts
    c = generate_host(300, 400, 0.06, n_topics=6, seed=11)
    a = SimpleAnnData(c.copy())
    tfidf(a)
    np.savez(os.path.join(OUT, "tfidf_synth.npz"), **csr_parts(c, "x"), **csr_parts(a.X, "out"))

    # lsi on TF-IDF of synthetic counts (float64 so that ARPACK's answer is the "truth")
    c = generate_host(600, 500, 0.08, n_topics=6, seed=5)
    a = SimpleAnnData(c.astype(np.float64))
    tfidf(a)
    X = a.X.copy()
    lsi(a, n_comps=8)
    b = SimpleAnnData(X.copy())
    lsi(b, n_comps=8, scale_embeddings=False)
    np.savez(os.path.join(OUT, "lsi_synth.npz"), **csr_parts(X, "x"), X_lsi=a.obsm["X_lsi"],
             stdev=a.uns["lsi"]["stdev"], LSI=a.varm["LSI"], U=b.obsm["X_lsi"])
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)))




def wnn_inputs(N=150, seed=0, k=15):
    """Two L2-normalised embeddings (8 and 6 dims) of 3 planted clusters + their exact kNN graphs, laid out the
    way ``sc.pp.neighbors`` leaves them (``obsp["distances"]`` with k-1 entries per row, ``uns["neighbors"]``)."""
    from oracle import _third_party as tp
    from muon_b200._containers import SimpleMuData
    rng = np.random.default_rng(seed)
    c = rng.integers(0, 3, N)
    reps = {"rna": rng.normal(size=(N, 8)) + 3 * np.eye(3)[c] @ rng.normal(size=(3, 8)),
            "atac": rng.normal(size=(N, 6)) + 2 * np.eye(3)[c] @ rng.normal(size=(3, 6))}
    mods = {}
    for name, R in reps.items():
        R = R / np.linalg.norm(R, axis=1, keepdims=True)
        ad = SimpleAnnData(np.zeros((N, 4)))
        ad.obsm["X_rep"] = R
        idx, dist, _ = tp.nearest_neighbors(R, k, "euclidean")
        ad.obsp["distances"] = sp.csr_matrix((dist[:, 1:].reshape(-1), idx[:, 1:].reshape(-1),
                                              np.arange(0, N * (k - 1) + 1, k - 1)), shape=(N, N))
        ad.uns["neighbors"] = {"params": {"n_neighbors": k, "use_rep": "X_rep", "metric": "euclidean"},
                               "distances_key": "distances", "connectivities_key": "connectivities"}
        mods[name] = ad
    return SimpleMuData(mods)


def dump_wnn():
    """Golden for the WNN row (round 2): the reference's own ``neighbors`` control flow
    (muon/_core/preproc.py:264-640) executed with the exact stand-ins of oracle/_third_party.py."""
    from oracle._refload import load_reference_neighbors
    neighbors = load_reference_neighbors()
    md = wnn_inputs()
    neighbors(md, n_multineighbors=40)
    out = {"rep_rna": md.mod["rna"].obsm["X_rep"], "rep_atac": md.mod["atac"].obsm["X_rep"],
           "w_rna": md.obs["rna:mod_weight"].to_numpy(), "w_atac": md.obs["atac:mod_weight"].to_numpy(),
           "n_neighbors": np.int64(md.uns["neighbors"]["params"]["n_neighbors"])}
    out.update(csr_parts(md.mod["rna"].obsp["distances"], "knn_rna"))
    out.update(csr_parts(md.mod["atac"].obsp["distances"], "knn_atac"))
    out.update(csr_parts(md.obsp["distances"], "wnn_dist"))
    out.update(csr_parts(md.obsp["connectivities"], "wnn_conn"))
    np.savez_compressed(os.path.join(OUT, "wnn_small.npz"), **out)
    print("wnn_small.npz", os.path.getsize(os.path.join(OUT, "wnn_small.npz")))


def dump_signatures():
    """Parameter names and defaults of the three reference entry points, read with ``ast`` (no import):
    muon/_atac/preproc.py::tfidf, ::binarize, muon/_atac/tools.py::lsi, muon/_core/tools.py::mofa."""
    import ast
    import json
    from oracle._refload import REF_ROOT
    out = {}
    for rel, names in (("muon/_atac/preproc.py", ("tfidf", "binarize")), ("muon/_atac/tools.py", ("lsi",)),
                       ("muon/_core/tools.py", ("mofa",)), ("muon/_core/preproc.py", ("neighbors",))):
        tree = ast.parse(open(os.path.join(REF_ROOT, rel)).read())
        for node in tree.body:
            if isinstance(node, ast.FunctionDef) and node.name in names:
                args = node.args
                params = [a.arg for a in args.args]
                defaults = [ast.literal_eval(d) for d in args.defaults]
                defaults = [None] * (len(params) - len(defaults)) + defaults
                required = len(params) - len(args.defaults)
                out[node.name] = {"params": params, "defaults": [repr(d) for d in defaults], "required": required}
    json.dump(out, open(os.path.join(OUT, "signatures.json"), "w"), indent=1)
    print("signatures.json", {k: len(v["params"]) for k, v in out.items()})


if __name__ == "__main__":
    if "--wnn-only" in sys.argv:
        dump_wnn()
    elif "--signatures-only" in sys.argv:
        dump_signatures()
    else:
        main()
        dump_wnn()
        dump_signatures()
