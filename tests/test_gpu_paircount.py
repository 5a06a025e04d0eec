"""GPU parity: K6 pair counting against the reference's golden dict and the oracle."""
import os
import tempfile

import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["auto", "v3", "v2"])
def kernel_variant(request, monkeypatch):
    """Every test of this file runs with the default dispatch (the device picks v3 -- count first, correct later -- or
    v2 -- decide first -- by the stream's alphabet) and with each of the two forced; HYP_PAIR_COUNT is read at every
    call.  v1 has its own test below."""
    if request.param == "auto":
        monkeypatch.delenv("HYP_PAIR_COUNT", raising=False)
    else:
        monkeypatch.setenv("HYP_PAIR_COUNT", request.param)
    return request.param


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


@pytest.mark.parametrize("chunk", [16384, 8192, 6144, 4096])
def test_chunk_boundaries(chunk):
    """Lines and whitespace runs straddling the CTA chunk boundaries (16 KiB in the v1 kernel, 8 KiB -- 6 KiB with the
    32-symbol table -- in v2, 4 KiB in v3)."""
    from hyptokenizer_b200.pair_count import count_pairs
    from oracle.merge import count_pairs_py
    parts = []
    for k in range(40):
        parts.append("x" * (chunk - 3 + (k % 7)) + " " * (k % 5) + "\n" + " " * (k % 3) + "yz")
    text = "".join(parts)
    assert count_pairs(text.encode("utf-8")) == count_pairs_py(_lines(text))


def test_wide_alphabet_and_counter_wraps():
    """More distinct symbols than the private table of the v2 kernel holds (31) and than its CTA histogram holds (64),
    with enough repetitions for the one-byte private counters to wrap many times; exact against the C oracle."""
    import numpy as np
    from hyptokenizer_b200.pair_count import count_pairs
    from oracle.pair_count import count_pairs_c
    rng = np.random.default_rng(5)
    alphabet = np.frombuffer(bytes(range(33, 127)) + b" \t", np.uint8)               # 96 symbols
    w = 1.0 / (np.arange(len(alphabet)) + 1.0)
    data = rng.choice(alphabet, size=6 << 20, p=w / w.sum())
    data[rng.integers(0, data.size, data.size // 60)] = 10
    got = count_pairs(data)
    want = count_pairs_c(data)
    assert got == want


def test_strip_corrections_dense():
    """Leading / trailing white space on EVERY line, tabs and double spaces inside, CR LF breaks, symbols outside the
    private alphabet: the cases where v3's unconditional count has to be corrected (-1 / +1 in the global table)."""
    import numpy as np
    from hyptokenizer_b200.pair_count import count_pairs
    from oracle.pair_count import count_pairs_c
    rng = np.random.default_rng(11)
    words = ["".join(rng.choice(list("abcdefghij"), size=rng.integers(1, 7))) for _ in range(500)]
    seps = [" ", "  ", "\t", " \t ", " "]
    ends = ["\n", " \n", "  \r\n", "\t\n ", "\n\n", "\r", " \n  "]
    parts = []
    for k in range(120000):
        parts.append(words[rng.integers(0, len(words))])
        parts.append(seps[rng.integers(0, len(seps))] if k % 9 else ends[rng.integers(0, len(ends))])
        if k % 1013 == 0:
            parts.append("Q~#")                      # rare symbols: outside the 27 private ranks
    data = np.frombuffer("".join(parts).encode("ascii"), np.uint8)
    assert count_pairs(data) == count_pairs_c(data)


def test_counter_wraps_with_repeated_symbols():
    """One-byte private counters that wrap many times per lane, on runs of one symbol -- so that two consecutive pairs
    keep hitting the SAME bin (the case where the second store carries both increments and a wrap must be carried
    exactly once) -- and on a Zipf-skewed wide alphabet at a size where the hottest bins wrap in every lane."""
    import numpy as np
    from hyptokenizer_b200.pair_count import count_pairs
    from oracle.pair_count import count_pairs_c
    rng = np.random.default_rng(3)
    n = 48 << 20
    data = np.full(n, ord("x"), np.uint8)
    pos = rng.integers(0, n, n // 40)
    data[pos] = rng.choice(np.frombuffer(b"yz \n", np.uint8), size=pos.size, p=[0.4, 0.3, 0.2, 0.1])
    assert count_pairs(data) == count_pairs_c(data)
    alphabet = np.frombuffer((" etaoinshrdlcumwfgypbvkjxqz" "ETAOINSHRDLCUMWFGYPBVKJXQZ" "0123456789"
                              ".,;:!?'\"()-\t").encode(), np.uint8)
    w = 1.0 / (np.arange(len(alphabet)) + 1.5) ** 2
    wide = rng.choice(alphabet, size=96 << 20, p=w / w.sum())
    wide[rng.integers(0, wide.size, wide.size // 80)] = 10
    assert count_pairs(wide) == count_pairs_c(wide)


def test_v1_kernel_matches_too():
    """The atomics-only kernel stays selectable (HYP_PAIR_COUNT=v1) and exact."""
    import subprocess
    import sys
    code = (
        "import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "from test_gpu_paircount import EDGE, _lines\n"
        "from hyptokenizer_b200.pair_count import count_pairs\n"
        "from hyptokenizer_b200.synth import synthetic_corpus\n"
        "from oracle.merge import count_pairs_py\n"
        "for k, t in EDGE.items():\n"
        "    assert count_pairs(t.encode('utf-8')) == count_pairs_py(_lines(t)), k\n"
        "d = synthetic_corpus(8 << 20, seed=3)\n"
        "ls = d.tobytes().decode('ascii').split('\\n')\n"
        "assert sum(count_pairs(d).values()) == sum(max(len(l.strip()) - 1, 0) for l in ls)\n"
        "print('v1 ok')\n") % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, HYP_PAIR_COUNT="v1")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "v1 ok" in r.stdout, r.stdout + r.stderr


def test_large_ascii_vs_oracle():
    """64 MB of synthetic config-4 text (lines of 20 random words): the exact dict against the C oracle, the checksum
    property total pairs == sum over lines of max(len(strip(line)) - 1, 0), and a sample against the Python oracle."""
    from hyptokenizer_b200.pair_count import count_pairs
    from hyptokenizer_b200.synth import synthetic_corpus
    from oracle.merge import count_pairs_py
    from oracle.pair_count import count_pairs_c
    data = synthetic_corpus(64 << 20, seed=0)
    got = count_pairs(data)
    assert got == count_pairs_c(data)               # the full dict against the C restatement of the reference's loop
    lines = data.tobytes().decode("ascii").split("\n")
    total = sum(max(len(ln.strip()) - 1, 0) for ln in lines)
    assert sum(got.values()) == total
    sample = lines[:20000]
    assert count_pairs(("\n".join(sample)).encode("ascii")) == count_pairs_py(sample)


def test_english_letter_frequencies_vs_oracle():
    """48 MB with English letter frequencies: rare letters (0.07 % of the text) that a small sample may not contain, and
    hot bigrams whose one-byte counters pass 255 many times per lane -- the full dict against the C oracle.  A stream
    that starts with 64 KB of a different alphabet must not change the counts either (the rank table is a sample)."""
    import numpy as np
    from hyptokenizer_b200.pair_count import count_pairs
    from hyptokenizer_b200.synth import english_corpus, synthetic_corpus
    from oracle.pair_count import count_pairs_c
    data = english_corpus(48 << 20, seed=3)
    assert count_pairs(data) == count_pairs_c(data)
    mixed = np.concatenate([np.frombuffer(b"0123456789 ABCDEF\n" * 3641, np.uint8)[:65536], synthetic_corpus(8 << 20, seed=1),
                            english_corpus(8 << 20, seed=4)])
    assert count_pairs(mixed) == count_pairs_c(mixed)


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
