"""CompressionAwareTokenizer (reference tokenizer/compression_aware_tokenizer.py) as a host policy over the device
candidate lists, against golden traces of the unmodified reference (tests/golden/trace_compress.json,
oracle/gen_golden.py gen_trace_compress).  Bar: identical merges, candidate counts and vocabulary; best scores and
appended rows within 1e-5 relative."""
import os

import numpy as np
import pytest
import torch

from helpers import from_bits

pytestmark = pytest.mark.gpu
REL = 1e-5


def _cls():
    from hyptokenizer_b200.tokenizer.compression_aware_tokenizer import CompressionAwareTokenizer
    return CompressionAwareTokenizer


@pytest.mark.parametrize("run", [0, 1, 2])
def test_compression_trace(golden, run):
    gd = golden("trace_compress.json")
    r = gd["runs"][run]
    vocab, d = gd["vocab0"], gd["d"]
    emb = from_bits(r["init"], len(vocab), d + 1)
    tok = _cls()(vocab, torch.nn.Parameter(emb), corpus_sample=list(gd["sample"]), sample_size=r["sample_size"],
                 merge_threshold=r["threshold0"], max_vocab_size=256, semantics=r["semantics"])
    merges, heads = [], []
    merge, find = tok._merge_tokens, tok._find_merge_candidates

    def spy_merge(i, j):
        merges.append([int(i), int(j)])
        return merge(i, j)

    def spy_find():
        c = find()
        heads.append([len(c), (float(c[0][2]) if c else None)])
        return c

    tok._merge_tokens, tok._find_merge_candidates = spy_merge, spy_find
    tok.optimize_merges(steps=10, log_every=10 ** 9)
    assert merges == r["merges_ij"]
    assert [h[0] for h in heads] == [h[0] for h in r["heads"]]
    for (_, got), (_, want) in zip(heads, r["heads"]):
        assert (got is None) == (want is None)
        if want is not None:
            assert abs(got - want) <= REL * abs(want)
    fin = r["final"]
    assert tok.current_vocab_size == fin["n"] and tok.vocab == fin["vocab"]
    want = from_bits(fin["embeddings"], fin["n"], d + 1).numpy()
    got = tok.embeddings[: fin["n"]].detach().cpu().numpy()
    assert np.array_equal(np.isnan(got), np.isnan(want))
    ok = ~np.isnan(want).any(axis=1)
    scale = np.abs(want[ok]).max(axis=1, keepdims=True)
    assert np.all(np.abs(got[ok].astype(np.float64) - want[ok]) <= REL * scale)


def test_greedy_tokenizer_and_persistence(golden, tmp_path):
    gd = golden("trace_compress.json")
    vocab = gd["vocab0"] + ["th", "the", "ing", "in", "at", "water"]
    emb = torch.zeros((len(vocab), gd["d"] + 1))
    emb[:, 0] = 1.0
    tok = _cls()(vocab, torch.nn.Parameter(emb), corpus_sample=list(gd["sample"]), max_vocab_size=64)

    def reference_greedy(text, vocab):             # compression_aware_tokenizer.py:91-122, as shipped
        out, i = [], 0
        ordered = sorted(vocab, key=len, reverse=True)
        while i < len(text):
            for t in ordered:
                if text[i:].startswith(t):
                    out.append(t)
                    i += len(t)
                    break
            else:
                out.append(text[i])
                i += 1
        return out

    for text in gd["sample"] + ["", "thethe <pad>water?", "ÿing at"]:
        assert tok._tokenize_with_vocab(text, vocab) == reference_greedy(text, vocab)
    assert tok._compression_aware_scoring([(4, 5, 0.5)])[0] > 0.0
    no_corpus = _cls()(vocab, torch.nn.Parameter(emb.clone()), max_vocab_size=64)
    assert no_corpus._compression_aware_scoring([(4, 5, 1.0), (4, 6, 3.0)]) == [0.5, 0.25]
    path = str(tmp_path / "comp")
    tok.compression_weight, tok.sample_size = 0.6, 17
    tok.save(path)
    assert os.path.exists(os.path.join(path, "compression_config.json"))
    back = _cls().load(path)
    assert (back.compression_weight, back.distance_weight, back.sample_size) == (0.6, 0.3, 17) and back.vocab == tok.vocab
