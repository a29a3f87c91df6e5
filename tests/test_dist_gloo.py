"""CPU, world_size 2, gloo: the host logic of the sharded path — contiguous row
shards, one all-reduce of (hist, sse), identical global loss / perplexity on every rank and
equal to the 1-rank result (histogram exactly, loss to fp64 summation order)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as tdist
import torch.multiprocessing as mp

from oracle import vq_oracle as vo
from _cases import vq_inputs


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, out):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "d-vqvae_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from dvq import dist as ddist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    tdist.init_process_group("gloo", rank=rank, world_size=world)
    z, E, al, beta = vq_inputs(name)
    zf = z.reshape(-1, E.shape[1])
    lo, hi = ddist.shard_bounds(zf.shape[0], rank, world)
    idx, hist, sse = vo.forward_stats_chunked(zf[lo:hi], E)           # this rank's shard (oracle = checker)
    stats = torch.zeros(E.shape[0] + 1, dtype=torch.int64)
    stats[:E.shape[0]] = torch.from_numpy(hist)
    stats[E.shape[0]:].view(torch.float64)[0] = sse
    n_flag = ddist.allreduce_stats(stats, E.shape[0], hi - lo)
    assert n_flag == 0                                               # "take the histogram total"
    n_total = int(stats[:E.shape[0]].sum())
    g_hist = stats[:E.shape[0]].numpy().copy()
    g_sse = float(stats[E.shape[0]:].view(torch.float64)[0])
    out[rank] = (lo, hi, n_total, g_hist, g_sse, idx)
    tdist.destroy_process_group()


@pytest.mark.parametrize("name", ["vq_ragged", "vq_k512_d64"])
def test_two_rank_stats_equal_one_rank(name):
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, name, out), nprocs=world, join=True)
    z, E, al, beta = vq_inputs(name)
    zf = z.reshape(-1, E.shape[1])
    idx1, hist1, sse1 = vo.forward_stats_chunked(zf, E)
    (lo0, hi0, n0, h0, s0, i0), (lo1, hi1, n1, h1, s1, i1) = out[0], out[1]
    assert (lo0, hi1) == (0, zf.shape[0]) and hi0 == lo1 and abs((hi0 - lo0) - (hi1 - lo1)) <= 1
    assert n0 == n1 == zf.shape[0]
    assert np.array_equal(h0, h1) and np.array_equal(h0, hist1)       # integer histogram: exact
    assert s0 == s1 and abs(s0 - sse1) <= 1e-12 * abs(sse1)
    assert np.array_equal(np.concatenate([i0, i1]), idx1)             # row results independent of sharding
    loss = vo.loss_from_sse(s0, n0, E.shape[1], al, beta)
    assert abs(float(loss) - float(vo.loss_from_sse(sse1, zf.shape[0], E.shape[1], al, beta))) <= 1e-7 * abs(float(loss))


def test_shard_bounds_cover_rows_exactly():
    from dvq import dist as ddist
    for n in (0, 1, 7, 8, 4194304, 16777216 + 3):
        for w in (1, 2, 4, 8):
            edges = [ddist.shard_bounds(n, r, w) for r in range(w)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1
