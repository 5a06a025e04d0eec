"""CPU: the C restatement of pair counting equals the Python restatement and the reference's golden dict."""
import random

from oracle.merge import count_pairs_py
from oracle.pair_count import count_pairs_c


def test_c_equals_golden(golden):
    gd = golden("pair_counts.json")
    data = ("\n".join(gd["lines"]) + "\n").encode("utf-8")
    assert count_pairs_c(data) == {(a, b): n for a, b, n in gd["counts"]}


def test_c_equals_python_on_random_text():
    rng = random.Random(3)
    alphabet = "abcdefg  \té中\U0001F600　 "
    lines = ["".join(rng.choice(alphabet) for _ in range(rng.randint(0, 60))) for _ in range(500)]
    text = "\n".join(lines) + "\n"
    assert count_pairs_c(text.encode("utf-8")) == count_pairs_py(ln + "\n" for ln in lines)


def test_synthetic_streams_are_what_the_bench_says():
    """The two bench streams: 26 lower-case letters, single spaces, a line break every 20 words; uniform letters in one,
    English letter frequencies in the other (deterministic by seed).  On both the C restatement equals the Python one."""
    import numpy as np
    from hyptokenizer_b200.synth import english_corpus, synthetic_corpus
    for make in (synthetic_corpus, english_corpus):
        data = make(1 << 18, seed=5)
        assert data.dtype == np.uint8 and data.size == 1 << 18 and np.array_equal(data, make(1 << 18, seed=5))
        assert set(np.unique(data).tolist()) <= set(b"abcdefghijklmnopqrstuvwxyz \n")
        text = data.tobytes().decode("ascii")
        assert "  " not in text and " \n" not in text and "\n " not in text
        lines = text.split("\n")
        assert all(len(ln.split(" ")) == 20 for ln in lines[:-1])
        assert count_pairs_c(data) == count_pairs_py(ln + "\n" for ln in lines)
    eng = english_corpus(1 << 20, seed=1)
    share = {c: float((eng == ord(c)).mean()) for c in "ezq"}
    assert 0.09 < share["e"] < 0.12 and 0.0003 < share["z"] < 0.001 and 0.0004 < share["q"] < 0.0015
    uni = synthetic_corpus(1 << 20, seed=1)
    assert abs(float((uni == ord("z")).mean()) - float((uni == ord("e")).mean())) < 0.002
