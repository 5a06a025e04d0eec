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


# ---- pair counting over line-aligned byte shards -------------------------------------------------
def _corpus_bytes():
    import random
    rng = random.Random(3)
    words = ["".join(rng.choice("abcdefghij") for _ in range(rng.randint(1, 7))) for _ in range(60)]
    words += ["héllo", "naïve", "日本語", "x y"]
    lines = []
    for k in range(400):
        line = " ".join(rng.choice(words) for _ in range(rng.randint(0, 12)))
        lines.append(("  " if k % 7 == 0 else "") + line + ("\t " if k % 5 == 0 else ""))
    seps = ["\n", "\r\n", "\r"]
    text = "".join(ln + seps[k % 3] for k, ln in enumerate(lines)) + "tail without newline"
    return text.encode("utf-8")


def _oracle_counter(shard):
    from oracle.pair_count import count_pairs_c
    counts = count_pairs_c(np.ascontiguousarray(shard)) if shard.size else {}
    asc = torch.zeros(128 * 128, dtype=torch.int64)
    extra = {}
    for (a, b), v in counts.items():
        if ord(a) < 128 and ord(b) < 128:
            asc[ord(a) * 128 + ord(b)] = v
        else:
            extra[(ord(a), ord(b))] = v
    return asc, extra


def _pc_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hyptokenizer_b200.pair_count import count_pairs_sharded
    got = count_pairs_sharded(_corpus_bytes(), counter=_oracle_counter)
    torch.save(got, f"{out}.{rank}")
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_pair_count_host_logic(tmp_path, world):
    from oracle.pair_count import count_pairs_c
    out = str(tmp_path / "pc")
    mp.spawn(_pc_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    want = count_pairs_c(_corpus_bytes())
    assert len(want) > 50 and any(ord(a) > 127 for a, _ in want)
    for rank in range(world):
        assert torch.load(f"{out}.{rank}") == want


def test_shard_byte_ranges_are_line_aligned_and_cover():
    from hyptokenizer_b200.pair_count import shard_byte_range
    buf = np.frombuffer(_corpus_bytes(), dtype=np.uint8)
    for world in (1, 2, 3, 8, 64, 5000):
        cuts = [shard_byte_range(buf, r, world) for r in range(world)]
        assert cuts[0][0] == 0 and cuts[-1][1] == buf.size
        for (s0, e0), (s1, e1) in zip(cuts, cuts[1:]):
            assert e0 == s1 and s0 <= e0
        for s, e in cuts:
            assert s == 0 or s == buf.size or buf[s - 1] == 0x0A
    empty = np.zeros(0, dtype=np.uint8)
    assert shard_byte_range(empty, 0, 2) == (0, 0) and shard_byte_range(empty, 1, 2) == (0, 0)
    one_line = np.frombuffer(b"abcdef", dtype=np.uint8)
    assert [shard_byte_range(one_line, r, 2) for r in range(2)] == [(0, 6), (6, 6)]


def test_klein_l2_topk_and_recall_helpers():
    """The recall helpers of knn.py (pure torch, evaluation only): Klein-L2 kNN against a float64 numpy brute force."""
    import numpy as np
    import torch
    from hyptokenizer_b200.knn import klein_l2_topk, recall_at_k
    from oracle.lorentz import initialize_embeddings
    torch.manual_seed(3)
    E = initialize_embeddings(300, 12, 1.0, 0.3)
    rows = torch.tensor([0, 7, 150, 299])
    got = klein_l2_topk(E, 5, rows).numpy()
    K = (E[:, 1:] / (E[:, 0:1] + 1e-8)).double().numpy()
    for g, r in zip(got, rows.tolist()):
        d2 = ((K - K[r]) ** 2).sum(1)
        d2[r] = np.inf
        want = np.argsort(d2, kind="stable")[:5]
        assert set(g.tolist()) == set(want.tolist())
    a = torch.tensor([[1, 2, 3, 4], [5, 6, 7, 8]])
    b = torch.tensor([[4, 3, 9, 1], [0, 0, 0, 0]])
    assert recall_at_k(a, b) == (3 / 4 + 0) / 2
    assert recall_at_k(a, a) == 1.0
