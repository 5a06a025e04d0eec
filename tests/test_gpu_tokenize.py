"""GPU parity: batched tokenize / encode (apply_merges kernel) against the oracle's restatement of
HyperbolicTokenizer.tokenize and the reference's own golden strings."""
import random

import pytest
import torch

pytestmark = pytest.mark.gpu


def make_tok(vocab, history):
    from hyptokenizer_b200.synth import synthetic_embeddings
    from hyptokenizer_b200.tokenizer.hyperbolic_merge import HyperbolicTokenizer
    tok = HyperbolicTokenizer(list(vocab), torch.nn.Parameter(synthetic_embeddings(len(vocab), 4)), max_vocab_size=len(vocab) + 8)
    tok.merge_history = list(history)
    return tok


def oracle_tok(vocab, history):
    from oracle.merge import OracleTokenizer
    from hyptokenizer_b200.synth import synthetic_embeddings
    o = OracleTokenizer(list(vocab), synthetic_embeddings(len(vocab), 4), max_vocab_size=len(vocab) + 8)
    o.merge_history = list(history)
    return o


def test_reference_golden_strings():
    """reference tests/test_hyperbolic_tokenizer.py::test_tokenize_encode_decode."""
    vocab = ["<pad>", "<bos>", "<eos>", "<unk>", "a", "b", "c", "d", "e", "ab", "cd"]
    tok = make_tok(vocab, [("a", "b", "ab"), ("c", "d", "cd")])
    assert tok.tokenize_batch(["abcde", "", "e", "zzab"]) == [["ab", "cd", "e"], [], ["e"], ["z", "z", "ab"]]
    ids = tok.encode_batch(["abcde", "zab"])
    assert ids[0] == [tok.token2idx["ab"], tok.token2idx["cd"], tok.token2idx["e"]]
    assert ids[1] == [tok.token2idx["<unk>"], tok.token2idx["ab"]]
    assert tok.decode(ids[0]) == "abcde"


def test_random_rules_vs_oracle():
    rng = random.Random(5)
    alphabet = list("abcdefgh é中") + ["\U0001F600"]
    for trial in range(6):
        vocab = ["<pad>", "<bos>", "<eos>", "<unk>"] + alphabet[: rng.randint(4, len(alphabet))]
        pool = list(vocab[4:])
        history = []
        for _ in range(rng.randint(0, 25)):
            a, b = rng.choice(pool), rng.choice(pool)
            history.append((a, b, a + b))
            if len(a + b) < 9:
                pool.append(a + b)
        if history and trial % 2:
            history.append(history[0][:2] + ("X",))          # duplicate key, later one wins (:427-428)
        texts = ["".join(rng.choice(alphabet + ["z"]) for _ in range(rng.randint(0, 120))) for _ in range(300)]
        tok, ora = make_tok(vocab, history), oracle_tok(vocab, history)
        assert tok.tokenize_batch(texts) == [ora.tokenize(t) for t in texts]
        assert tok.encode_batch(texts) == [ora.encode(t) for t in texts]
        # the batch path and the reference-style single-string path agree
        assert [tok.tokenize(t) for t in texts[:50]] == tok.tokenize_batch(texts[:50])


def test_stale_rules_quirk():
    """`_merge_rules` is built once (hyperbolic_merge.py:425-428): merges made after the first tokenize call
    are ignored by later calls, single-string and batched alike."""
    vocab = ["<pad>", "<bos>", "<eos>", "<unk>", "a", "b"]
    tok = make_tok(vocab, [])
    assert tok.tokenize_batch(["abab"]) == [["a", "b", "a", "b"]]
    tok.merge_history.append(("a", "b", "ab"))
    assert tok.tokenize("abab") == ["a", "b", "a", "b"]
    assert tok.tokenize_batch(["abab"]) == [["a", "b", "a", "b"]]


def test_after_training_roundtrip():
    """Train a few merges on the device, then tokenize a corpus with the learnt rules: decode(encode(t)) == t for
    texts over the vocabulary's alphabet, and the batch equals the oracle's tokenize with the same history."""
    from hyptokenizer_b200.synth import synthetic_embeddings
    from hyptokenizer_b200.tokenizer.hyperbolic_merge import HyperbolicTokenizer
    from oracle.merge import OracleTokenizer
    chars = list("abcdefghijklmnopqrstuvwxyz ")
    vocab = ["<pad>", "<bos>", "<eos>", "<unk>"] + chars
    tok = HyperbolicTokenizer(vocab, torch.nn.Parameter(synthetic_embeddings(len(vocab), 16, scale=0.3, seed=4)),
                              merge_threshold=3.0, max_vocab_size=64, semantics="lorentz")
    tok.optimize_merges(steps=12)
    rng = random.Random(0)
    texts = ["".join(rng.choice(chars) for _ in range(rng.randint(1, 200))) for _ in range(2000)]
    ora = OracleTokenizer(vocab, synthetic_embeddings(len(vocab), 16), max_vocab_size=64)
    ora.merge_history = list(tok.merge_history)
    got = tok.tokenize_batch(texts)
    assert got == [ora.tokenize(t) for t in texts]
    for t, ids in zip(texts[:200], tok.encode_batch(texts[:200])):
        assert tok.decode(ids) == t


def test_token_pair_counts_vs_oracle():
    """General pair counting (hyp_pair_count_sorted): adjacent pairs of tokenize(line.strip()) WITH merge rules --
    multi-character tokens, a vocabulary-sized alphabet, characters without an id -- against the oracle's tokenize loop
    and a Python dict, the way _compute_pair_frequencies counts (frequency_aware_hyperbolic_merge.py:92-112)."""
    from hyptokenizer_b200.tokenizer.batch_tokenize import count_token_pairs
    rng = random.Random(11)
    alphabet = list("abcdefgh é中")
    vocab = ["<pad>", "<bos>", "<eos>", "<unk>"] + alphabet
    pool, history = list(alphabet[:8]), []
    for _ in range(300):
        x, y = rng.choice(pool), rng.choice(pool)
        if len(x + y) <= 5:
            history.append((x, y, x + y))
            pool.append(x + y)
    tok = make_tok(vocab + [h[2] for h in history], history)
    ora = oracle_tok(vocab + [h[2] for h in history], history)
    lines = []
    for k in range(4000):
        ln = "".join(rng.choice(alphabet + ["Z", "\U0001F600"]) for _ in range(rng.randint(0, 60)))
        lines.append(("  " if k % 7 == 0 else "") + ln + (" \t" if k % 5 == 0 else "") + "\n")
    lines += ["", "\n", "a\n", "   \n"]
    want = {}
    for ln in lines:
        toks = ora.tokenize(ln.strip())
        for a, b in zip(toks, toks[1:]):
            want[(a, b)] = want.get((a, b), 0) + 1
    got = count_token_pairs(tok, lines)
    assert got == want
    assert max(len(a) for a, _ in got) > 1              # multi-character tokens took part
    # empty rules: the character-bigram count of the dense kernel
    from hyptokenizer_b200.pair_count import count_pairs
    plain = make_tok(vocab, [])
    text = "".join(lines)
    assert count_token_pairs(plain, lines) == count_pairs(text.encode("utf-8"))
