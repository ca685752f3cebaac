"""BASELINE.json configs[4] (scaled): mu.pp.neighbors (WNN) on two L2-normalised embeddings (50 + 30 dims).

    python profiles/bench_wnn.py [cells=100000]

The per-modality kNN graphs that sc.pp.neighbors would leave behind (k=15) are produced with the same exact
kNN kernel and are NOT timed; the timed region is the mu.pp.neighbors call (host containers in, scipy graphs out).
"""
import json
import os
import sys
import time

import numpy as np
import scipy.sparse as sp
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import muon_b200 as mu  # noqa: E402
from muon_b200 import _device, _lib  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
k = 15
rng = np.random.default_rng(0)
c = rng.integers(0, 30, N)
mods = {}
for name, d in (("rna", 50), ("atac", 30)):
    R = rng.normal(size=(N, d)).astype(np.float32) + 3.0 * (np.eye(30, dtype=np.float32)[c] @ rng.normal(size=(30, d)).astype(np.float32))
    R /= np.linalg.norm(R, axis=1, keepdims=True)
    idx, dist = _device.knn_l2(torch.from_numpy(R).cuda(), k)
    g = sp.csr_matrix((dist[:, 1:].cpu().numpy().ravel().astype(np.float64), idx[:, 1:].cpu().numpy().ravel().astype(np.int64),
                       np.arange(0, N * (k - 1) + 1, k - 1)), shape=(N, N))
    ad = mu.SimpleAnnData(np.zeros((N, 1), dtype=np.float32))
    ad.obsm["X_emb"] = R
    ad.obsp["distances"] = g
    ad.uns["neighbors"] = {"params": {"n_neighbors": k, "use_rep": "X_emb"}, "distances_key": "distances"}
    mods[name] = ad
md = mu.SimpleMuData(mods)
mu.pp.neighbors(md)                       # warm-up
torch.cuda.synchronize()
_lib.PROFILE = {}
t0 = time.perf_counter()
mu.pp.neighbors(md)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
prof, _lib.PROFILE = _lib.PROFILE, None
kern = {name: float(sum(a.elapsed_time(b) for a, b in evs)) for name, evs in prof.items()}
print(json.dumps({"metric": "cells/sec for mu.pp.neighbors (WNN), 2 modalities (50+30 dims), n_multineighbors=200",
                  "value": N / dt, "unit": "cells/s", "cells": N, "seconds": dt, "kernel_ms": kern,
                  "weights_mean": [float(md.obs["rna:mod_weight"].mean()), float(md.obs["atac:mod_weight"].mean())]}))
