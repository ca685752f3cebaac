"""Where does the e2e (host scipy -> tfidf -> lsi -> numpy) time go?  python profiles/host_path_probe.py [cells]"""
import os
import sys
import time

import numpy as np
import scipy.sparse as sp
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import muon_b200 as mu  # noqa: E402
from muon_b200 import _device  # noqa: E402
from muon_b200._synth import generate_device, make_tables  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 250_000
d = 200_000
A = generate_device(n, d, 0.03, tables=make_tables(d, 0.03, 64, 1))
X = A.get()
del A
torch.cuda.synchronize()


def T(label, fn):
    torch.cuda.synchronize()
    t = time.perf_counter()
    r = fn()
    torch.cuda.synchronize()
    print(f"{label:40s} {time.perf_counter() - t:8.3f} s")
    return r


for rep in range(2):
    print("--- rep", rep, "nnz", X.nnz, "index dtype", X.indices.dtype)
    M = T("csr_matrix(copy=False)", lambda: sp.csr_matrix((X.data, X.indices, X.indptr), shape=X.shape, copy=False))
    di = T("to_device indices", lambda: _device.to_device(M.indices, torch.device("cuda"), np.int32))
    dd = T("to_device data", lambda: _device.to_device(M.data, torch.device("cuda"), np.float32))
    h = T("to_host data", lambda: _device.to_host(dd))
    del di, dd, h
    ad = mu.SimpleAnnData(M)
    T("tfidf(host)", lambda: mu.atac.pp.tfidf(ad))
    T("lsi(host, resident)", lambda: mu.atac.tl.lsi(ad, n_comps=50))
