"""SURVEY.md 8f-4 (first half): HierarchicalHyperbolicTokenizer as a host policy over the device candidate lists,
against golden traces of the unmodified reference (tests/golden/trace_hier.json, oracle/gen_golden.py
gen_trace_hier: the three phase budgets shortened to 25/30/30 steps).

Bar: identical merge sequence, identical per-step candidate counts and thresholds, identical vocabulary; appended
rows within 1e-5 relative; the string criteria and both re-weighting filters bit-identical (host float arithmetic)."""
import os

import numpy as np
import pytest
import torch

from helpers import from_bits

pytestmark = pytest.mark.gpu
REL = 1e-5


def _cls():
    from hyptokenizer_b200.tokenizer.hierarchical_hyperbolic_merge import HierarchicalHyperbolicTokenizer
    return HierarchicalHyperbolicTokenizer


def _corpus(tmp_path, lines):
    path = str(tmp_path / "corpus.txt")
    with open(path, "w", encoding="utf-8") as f:
        f.write("\n".join(lines) + "\n")
    return path


def _record(tok):
    merges, steps = [], []
    merge, find = tok._merge_tokens, tok._find_merge_candidates

    def spy_merge(i, j):
        merges.append([int(i), int(j)])
        return merge(i, j)

    def spy_find():
        c = find()
        steps.append([len(c), tok.merge_threshold])
        return c

    tok._merge_tokens, tok._find_merge_candidates = spy_merge, spy_find
    return merges, steps


@pytest.mark.parametrize("run", [0, 1, 2])
def test_hierarchical_trace(golden, tmp_path, monkeypatch, run):
    gd = golden("trace_hier.json")
    r = gd["runs"][run]
    cls = _cls()
    monkeypatch.setattr(cls, "phase_steps", tuple(gd["phase_steps"]))
    vocab, d = gd["vocab0"], gd["d"]
    emb = from_bits(r["init"], len(vocab), d + 1)
    tok = cls(vocab, torch.nn.Parameter(emb), corpus_path=_corpus(tmp_path, gd["lines"]), max_vocab_size=400,
              semantics=r["semantics"])
    assert sorted(tok.common_morphemes) == r["common_morphemes"] and sorted(tok.common_words) == r["common_words"]
    merges, steps = _record(tok)
    tok.optimize_merges(hierarchical=True, target_vocab_size=r["target"])
    assert merges == r["merges_ij"]
    assert [s[0] for s in steps] == [s[0] for s in r["steps"]]
    assert [s[1] for s in steps] == [s[1] for s in r["steps"]]          # Python-float threshold arithmetic: exact
    fin = r["final"]
    assert tok.current_vocab_size == fin["n"] and tok.vocab == fin["vocab"]
    assert [list(m) for m in tok.merge_history] == fin["merges"]
    want = from_bits(fin["embeddings"], fin["n"], d + 1).numpy()
    got = tok.embeddings[: fin["n"]].detach().cpu().numpy()
    assert np.array_equal(np.isnan(got), np.isnan(want))
    ok = ~np.isnan(want).any(axis=1)
    scale = np.abs(want[ok]).max(axis=1, keepdims=True)
    assert np.all(np.abs(got[ok].astype(np.float64) - want[ok]) <= REL * scale)


def test_string_criteria_and_filters(golden, tmp_path):
    gd = golden("trace_hier.json")
    cr = gd["criteria"]
    vocab = cr["vocab"]
    emb = torch.zeros((len(vocab), gd["d"] + 1))
    emb[:, 0] = 1.0
    tok = _cls()(vocab, torch.nn.Parameter(emb), corpus_path=_corpus(tmp_path, gd["lines"]), max_vocab_size=400)
    assert [tok._is_potential_morpheme(t) for t in cr["strings"]] == cr["morpheme"]
    assert [tok._is_valid_word(t) for t in cr["strings"]] == cr["word"]
    cands = [(i, j, dist) for i, j, dist in cr["candidates"]]
    assert [list(c) for c in tok._filter_morphologically_valid(cands)] == cr["morph_filtered"]
    assert [list(c) for c in tok._filter_word_valid(cands)] == cr["word_filtered"]
    assert any(a[2] != b[2] for a, b in zip(cr["candidates"], cr["morph_filtered"]))      # the fixture does re-weight


def test_non_hierarchical_and_save_load(golden, tmp_path):
    """hierarchical=False is the base loop (:445-448); save/load round-trips the extra file (:450-513)."""
    from hyptokenizer_b200.tokenizer.hyperbolic_merge import HyperbolicTokenizer
    gd = golden("trace_hier.json")
    r = gd["runs"][1]
    vocab, d = gd["vocab0"], gd["d"]
    emb = from_bits(r["init"], len(vocab), d + 1)
    corpus = _corpus(tmp_path, gd["lines"])
    a = _cls()(vocab, torch.nn.Parameter(emb.clone()), corpus_path=corpus, merge_threshold=2.0, max_vocab_size=200,
               semantics="lorentz")
    b = HyperbolicTokenizer(vocab, torch.nn.Parameter(emb.clone()), merge_threshold=2.0, max_vocab_size=200,
                            semantics="lorentz")
    a.optimize_merges(steps=20, hierarchical=False)
    b.optimize_merges(steps=20)
    assert a.vocab == b.vocab and torch.equal(a.embeddings[:50], b.embeddings[:50])
    path = str(tmp_path / "hier")
    a.save(path)
    assert os.path.exists(os.path.join(path, "hierarchical_data.json"))
    back = _cls().load(path)
    assert isinstance(back, _cls()) and back.vocab == a.vocab
    assert back.common_words == a.common_words and back.common_morphemes == a.common_morphemes
    assert back.language == "english"
