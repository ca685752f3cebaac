"""CPU, world_size=2 over gloo: the cell-sharded host logic -- row shards per
rank, allreduce of A_r^T Y_r, of the b x b Gram and of the z-score moments -- gives the same
factorisation as a single process."""
import os
import socket

import numpy as np
import scipy.sparse as sp
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from muon_b200 import _dist
from muon_b200._lsi import truncated_svd
from muon_b200._synth import generate_host
from oracle.lsi_ref import lsi_ref
from oracle.tfidf_ref import tfidf_ref


class ShardOperator:
    def __init__(self, A, n_total):
        self.A = sp.csr_matrix(A).astype(np.float32)
        self.At = self.A.T.tocsr()
        self.n_local, self.d = self.A.shape
        self.n_total = n_total
        self.device = torch.device("cpu")

    def av(self, V):
        return torch.from_numpy(self.A @ V.numpy())

    def aty(self, Y):
        return _dist.all_reduce_sum_(torch.from_numpy(self.At @ Y.numpy()))

    def gram(self, Y, l):
        y = Y[:, :l].double()
        return _dist.all_reduce_sum_(y.T @ y)


def _worker(rank, world, port, X, k, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = X.shape[0]
        lo, hi = rank * n // world, (rank + 1) * n // world
        assert _dist.is_distributed() and _dist.world_size() == world
        U, s, V, info = truncated_svd(ShardOperator(X[lo:hi], n), k, 32, tol=1e-6)
        mom = torch.stack([U.sum(0, dtype=torch.float64), (U.double() ** 2).sum(0)])
        _dist.all_reduce_sum_(mom)
        out[rank] = (s.numpy(), U.numpy(), V.numpy(), mom.numpy(), info.converged)
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_svd_matches_single():
    X = tfidf_ref(generate_host(900, 700, 0.06, n_topics=8, seed=13)).astype(np.float32)
    k = 6
    ref = lsi_ref(X, k, scale_embeddings=False, dtype=np.float64)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, X, k, out), nprocs=2, join=True)
    s0, U0, V0, mom0, c0 = out[0]
    s1, U1, V1, mom1, c1 = out[1]
    assert c0 and c1
    np.testing.assert_array_equal(s0, s1)                 # replicated quantities identical on both ranks
    np.testing.assert_array_equal(V0, V1)
    np.testing.assert_allclose(s0, ref["svalues"], rtol=1e-5)
    U = np.vstack([U0, U1])
    cos = np.abs((U * ref["U"]).sum(0))
    assert np.all(1 - cos < 1e-6)
    np.testing.assert_allclose(mom0[1], np.ones(k), rtol=1e-5)   # sum of squares of unit columns
