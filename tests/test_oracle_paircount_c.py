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
