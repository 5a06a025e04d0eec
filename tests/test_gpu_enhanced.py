"""EnhancedFastHyperbolicTokenizer (reference tokenizer/enhanced_fast_hyperbolic_merge.py, BASELINE config 5) through
the device kernels, against golden traces of the unmodified reference under the import shim of SURVEY.md 8c
(tests/golden/trace_enhanced.json, oracle/gen_golden.py gen_trace_enhanced).  The host policy alone is covered on the
CPU by tests/test_enhanced_host_logic.py; this file is the same comparison with K2 / K1 / K3 / K6 / K7 underneath."""
import pytest
import torch

import enhanced_common as EC
from helpers import from_bits

pytestmark = pytest.mark.gpu


def _cls():
    from hyptokenizer_b200.tokenizer.enhanced_fast_hyperbolic_merge import EnhancedFastHyperbolicTokenizer
    return EnhancedFastHyperbolicTokenizer


@pytest.fixture()
def corpus(golden, tmp_path):
    gd = golden("trace_enhanced.json")
    path = tmp_path / "corpus.txt"
    path.write_text("\n".join(gd["lines"]) + "\n", encoding="utf-8")
    return gd, str(path)


@pytest.mark.parametrize("run", [0, 1, 2, 3, 4])
def test_enhanced_trace(corpus, run):
    gd, path = corpus
    r = gd["runs"][run]
    tok, merges, heads, curv = EC.run_golden(_cls(), gd, r, path)
    assert tok.embeddings.is_cuda
    if EC.one_bit_tie_step(heads, r["heads"]) is None:
        EC.check_run(tok, merges, heads, curv, gd, r, strict=True)
        return
    # The plain run meets a one-bit tie (two copies of a merged row whose mutual product is 1 or 1 + 2^-23 depending
    # on the last bit of x0): identical up to there ...
    EC.check_run(tok, merges, heads, curv, gd, r)
    # ... and PROVED to be nothing else: with the reference's bits injected into every appended row (each within a few
    # ulp of what the device computed, asserted), the whole run is the reference's -- every merge, candidate count,
    # threshold, phase, score, the statistics and the final table.
    tok2, merges2, heads2, curv2 = EC.run_golden(_cls(), gd, r, path, inject_reference_rows=True)
    EC.check_run(tok2, merges2, heads2, curv2, gd, r, strict=True)
    assert tok2._row_injector.rows == len(r["merges_ij"])
    print(f"run {run}: strict parity with {tok2._row_injector.rows} injected rows, max device-vs-reference row "
          f"difference {tok2._row_injector.max_seen} ulp")


def test_shipped_curvature_step_raises_like_the_reference(corpus):
    gd, path = corpus
    s = gd["shipped_curvature_step"]
    vocab = gd["vocab0"]
    emb = from_bits(s["init"], len(vocab), gd["d"] + 1)
    tok = _cls()(vocab, torch.nn.Parameter(emb), max_vocab_size=160, use_approximate_search=False,
                 use_frequency_aware=False, use_compression_aware=False, optimize_curvature_freq=s["freq"],
                 semantics="reference")
    with pytest.raises(RuntimeError) as e:
        tok.optimize_merges(steps=20, log_every=1000)
    assert str(e.value) == s["error"]
    assert [[i, j] for i, j, _ in tok.last_trace] == s["merges_ij"]


def test_fresh_search_picks_the_best_scored_pair_of_the_whole_table(corpus):
    """cache_semantics="fresh": every step's head is the arg-max of the combined score over ALL pairs under the
    threshold at that step (checked against a per-step re-score of the full list)."""
    gd, path = corpus
    r = gd["runs"][3]                                   # compression-aware only: deterministic scores, no RNG
    vocab, d = gd["vocab0"], gd["d"]
    emb = from_bits(r["init"], len(vocab), d + 1)
    tok = _cls()(vocab, torch.nn.Parameter(emb), merge_threshold=r["threshold0"], max_vocab_size=160,
                 corpus_sample=list(gd["sample"]), semantics="lorentz", cache_semantics="fresh", **r["flags"])
    for _ in range(6):
        ii, jj, dd = tok._candidate_arrays()
        scored = tok._score_arrays(ii, jj, dd)
        want = min(range(len(scored)), key=lambda k: scored[k].combined_score)
        got = tok._find_merge_candidates_fast()[0]
        assert (got.token_i, got.token_j) == (scored[want].token_i, scored[want].token_j)
        tok._merge_tokens(got.token_i, got.token_j)
        for key in [k for k in tok.tokenize_cache if k.startswith("merge_")]:
            tok.tokenize_cache.pop(key)


def test_save_load_round_trip(corpus, tmp_path):
    gd, path = corpus
    tok, *_ = EC.run_golden(_cls(), gd, gd["runs"][4], path)
    EC.check_save_load(_cls(), tok, gd, str(tmp_path / "saved"))


@pytest.mark.parametrize("run", [0, 1, 2, 3])
def test_adaptive_curvature_trace(golden, run, tmp_path):
    """AdaptiveCurvatureTokenizer (reference tokenizer/adaptive_curvature_tokenizer.py) against
    tests/golden/trace_adaptive.json: shipped arithmetic below and at its first curvature step (which raises, as the
    reference does), and the corrected step."""
    from hyptokenizer_b200.tokenizer.adaptive_curvature_tokenizer import AdaptiveCurvatureTokenizer
    gd = golden("trace_adaptive.json")
    tok = EC.run_adaptive(AdaptiveCurvatureTokenizer, gd, gd["runs"][run])
    assert tok.embeddings.is_cuda
    if run == 3:
        EC.check_adaptive_save_load(AdaptiveCurvatureTokenizer, tok, str(tmp_path / "saved"))
