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




def dump_signatures():
    """Parameter names and defaults of the three reference entry points, read with ``ast`` (no import):
    muon/_atac/preproc.py::tfidf, ::binarize, muon/_atac/tools.py::lsi, muon/_core/tools.py::mofa."""
    import ast
    import json
    from oracle._refload import REF_ROOT
    out = {}
    for rel, names in (("muon/_atac/preproc.py", ("tfidf", "binarize")), ("muon/_atac/tools.py", ("lsi",)),
                       ("muon/_core/tools.py", ("mofa",))):
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
    if "--signatures-only" not in sys.argv:
        main()
    dump_signatures()
