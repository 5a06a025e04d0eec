"""GPU parity: K6 pair counting against the reference's golden dict and the oracle."""
import os
import tempfile

import pytest

pytestmark = pytest.mark.gpu


def test_golden_pair_counts(golden):
    from hyptokenizer_b200.pair_count import count_pairs
    gd = golden("pair_counts.json")
    data = ("\n".join(gd["lines"]) + "\n").encode("utf-8")
    got = count_pairs(data)
    want = {(a, b): n for a, b, n in gd["counts"]}
    assert got == want


def _lines(text):
    """What iterating a text-mode file (universal newlines) yields -- the reference's input."""
    fd, path = tempfile.mkstemp()
    with os.fdopen(fd, "wb") as f:
        f.write(text.encode("utf-8"))
    with open(path, "r", encoding="utf-8") as f:
        out = list(f)
    os.remove(path)
    return out


EDGE = {
    "empty": "",
    "no_newline": "  hello world  ",
    "crlf": "ab cd\r\nef\rgh\n\r\n  ij  \r",
    "ws_only": "   \n\t\t\n \x0b\x0c \n",
    "long_ws_runs": "a" + " " * 5000 + "b\n" + " " * 40000 + "c d" + "\t" * 33000 + "\n" + " " * 20000,
    "unicode_ws": "　ab cd \n x y z\n\x1cq\x1dr\x1f\n",
    "one_char_lines": "a\nb\n\nc\n",
    "multibyte": "éè \U0001F600\U0001F601x\n中文中文\n",
}


@pytest.mark.parametrize("case", sorted(EDGE))
def test_edge_cases(case):
    from hyptokenizer_b200.pair_count import count_pairs
    from oracle.merge import count_pairs_py
    text = EDGE[case]
    assert count_pairs(text.encode("utf-8")) == count_pairs_py(_lines(text))


def test_chunk_boundaries():
    """Lines and whitespace runs straddling the 16 KiB CTA chunk boundary."""
    from hyptokenizer_b200.pair_count import count_pairs
    from oracle.merge import count_pairs_py
    parts = []
    for k in range(40):
        parts.append("x" * (16384 - 3 + (k % 7)) + " " * (k % 5) + "\n" + " " * (k % 3) + "yz")
    text = "".join(parts)
    assert count_pairs(text.encode("utf-8")) == count_pairs_py(_lines(text))


def test_large_ascii_vs_oracle():
    """64 MB of synthetic config-4 text (lines of 20 random words).  Checksum property at full chunk
    counts: total pairs == sum over lines of max(len(strip(line)) - 1, 0); exact dict on a sample."""
    from hyptokenizer_b200.pair_count import count_pairs
    from hyptokenizer_b200.synth import synthetic_corpus
    from oracle.merge import count_pairs_py
    data = synthetic_corpus(64 << 20, seed=0)
    got = count_pairs(data)
    lines = data.tobytes().decode("ascii").split("\n")
    total = sum(max(len(ln.strip()) - 1, 0) for ln in lines)
    assert sum(got.values()) == total
    sample = lines[:20000]
    assert count_pairs(("\n".join(sample)).encode("ascii")) == count_pairs_py(sample)


def test_sharded_single_process_equals_plain():
    """Without a process group the sharded entry point is the plain one (one shard = the whole corpus)."""
    from hyptokenizer_b200.pair_count import count_pairs, count_pairs_sharded
    from hyptokenizer_b200.synth import synthetic_corpus
    data = synthetic_corpus(4 << 20, seed=2)
    assert count_pairs_sharded(data) == count_pairs(data)


def _pc_nccl_worker(rank, world, port, out):
    import socket  # noqa: F401
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from hyptokenizer_b200.pair_count import count_pairs_sharded
    from hyptokenizer_b200.synth import synthetic_corpus
    data = synthetic_corpus(32 << 20, seed=4)
    got = count_pairs_sharded(data, device=torch.device("cuda", rank))
    torch.save(got, f"{out}.{rank}")
    dist.destroy_process_group()


@pytest.mark.skipif(__import__("torch").cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_sharded_pair_count_nccl(tmp_path):
    import socket
    import torch
    import torch.multiprocessing as mp
    from hyptokenizer_b200.pair_count import count_pairs
    from hyptokenizer_b200.synth import synthetic_corpus
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    world = min(torch.cuda.device_count(), 8)
    out = str(tmp_path / "pc")
    mp.spawn(_pc_nccl_worker, args=(world, port, out), nprocs=world, join=True)
    want = count_pairs(synthetic_corpus(32 << 20, seed=4))
    for rank in range(world):
        assert torch.load(f"{out}.{rank}") == want
