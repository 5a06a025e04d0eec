"""CPU, world_size 2 over gloo: the host side of the row-sharded top-k (partition, all-gather, replicated
argmin).  The per-shard lists come from the oracle here; on the GPU they come from the kernels."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_topk(E, k, row0, nrows, sem):
    from oracle import lorentz as OL
    d = OL.batch_distance(E[row0:row0 + nrows], E, 1.0, sem)
    n = E.shape[0]
    d[torch.arange(nrows), torch.arange(row0, row0 + nrows)] = float("inf")
    # ascending by (d, j): stable sort on d over j-ordered columns
    order = torch.sort(d, dim=1, stable=True)
    idx = order.indices[:, :k].to(torch.int32)
    val = order.values[:, :k]
    idx[~torch.isfinite(val)] = -1
    return idx, val


def _worker(rank, world, port, n, k, sem, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hyptokenizer_b200 import knn
    from hyptokenizer_b200.synth import synthetic_embeddings
    E = synthetic_embeddings(n, 12, scale=0.3, seed=5)
    row0, nrows, per = knn.shard_rows(n, world, rank)
    li, ld = _oracle_topk(E, k, row0, nrows, sem)
    gi, gd = knn.gather_topk(li, ld, n)
    best = knn.best_pair_from_topk(gi, gd)
    if rank == 0:
        torch.save({"idx": gi, "d": gd, "best": best}, out)
    else:
        torch.save({"best": best}, out + ".1")
    dist.destroy_process_group()


@pytest.mark.parametrize("n,sem", [(101, "lorentz"), (64, "reference"), (7, "lorentz")])
def test_sharded_topk_host_logic(tmp_path, n, sem):
    from hyptokenizer_b200 import knn
    from hyptokenizer_b200.synth import synthetic_embeddings
    k = 5
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(2, _free_port(), n, k, sem, out), nprocs=2, join=True)
    got = torch.load(out)
    other = torch.load(out + ".1")
    E = synthetic_embeddings(n, 12, scale=0.3, seed=5)
    wi, wd = _oracle_topk(E, k, 0, n, sem)
    assert torch.equal(got["idx"], wi) and torch.equal(got["d"], wd)
    assert got["best"] == other["best"]                      # replicated reduction agrees across ranks
    # and equals the brute-force argmin over (d, i, j)
    from oracle import merge as OM
    tok = OM.OracleTokenizer([str(t) for t in range(n)], E, 1.0, float("inf"), n + 1, sem)
    ii, jj, dd = tok.find_candidates()
    want = OM.OracleTokenizer.pick(ii, jj, dd)
    assert (got["best"][0], got["best"][1]) == (want[0], want[1]) and got["best"][2] == want[2]


def test_shard_rows_cover():
    from hyptokenizer_b200.knn import shard_rows
    for n in (0, 1, 7, 64, 100000, 100001):
        for w in (1, 2, 3, 4, 8):
            covered = []
            for r in range(w):
                r0, nr, per = shard_rows(n, w, r)
                covered += list(range(r0, r0 + nr)) if n < 1000 else [(r0, nr)]
                assert 0 <= nr <= per
            if n < 1000:
                assert covered == list(range(n))
            else:
                assert sum(nr for _, nr in covered) == n
