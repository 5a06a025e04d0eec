"""GPU parity: FrequencyAwareHyperbolicTokenizer (K6 + K2 + K7) against the reference's golden run."""
import numpy as np
import pytest
import torch

from helpers import from_bits

pytestmark = pytest.mark.gpu


def test_trace_freq(golden, tmp_path):
    from hyptokenizer_b200.tokenizer.frequency_aware_hyperbolic_merge import FrequencyAwareHyperbolicTokenizer
    gd = golden("trace_freq.json")
    corpus = tmp_path / "corpus.txt"
    corpus.write_text("\n".join(gd["lines"]) + "\n", encoding="utf-8")
    for run in gd["runs"]:
        emb = from_bits(run["init"], len(gd["vocab0"]), gd["d"] + 1)
        tok = FrequencyAwareHyperbolicTokenizer(gd["vocab0"], torch.nn.Parameter(emb), corpus_path=str(corpus),
                                                merge_threshold=run["threshold0"], max_vocab_size=256,
                                                semantics=run["semantics"])
        torch.manual_seed(123)
        tok.optimize_merges(steps=8, log_every=10 ** 9)
        assert [[a, b] for a, b, _ in tok.last_trace] == run["merges_ij"]
        for (_, _, s), (_, want) in zip(tok.last_trace, run["best_neg_score"]):
            if want is None:
                assert s != s                   # shipped semantics: every score is NaN
            else:
                assert abs(s - want) <= 1e-5 * abs(want)
        assert tok.vocab == run["final"]["vocab"]
        want = np.asarray(run["final"]["embeddings"], dtype=np.uint32).view(np.float32).reshape(run["final"]["n"], -1)
        got = tok.embeddings[: tok.current_vocab_size].detach().cpu().numpy()
        assert np.array_equal(np.isnan(got), np.isnan(want))
        ok = ~np.isnan(want).any(axis=1)
        assert np.all(np.abs(got[ok] - want[ok]) <= 1e-5 * np.abs(want[ok]).max(axis=1, keepdims=True))


def test_pair_frequencies_from_file(golden, tmp_path):
    from hyptokenizer_b200.tokenizer.frequency_aware_hyperbolic_merge import FrequencyAwareHyperbolicTokenizer
    from hyptokenizer_b200.synth import synthetic_embeddings
    gd = golden("pair_counts.json")
    corpus = tmp_path / "c.txt"
    corpus.write_text("\n".join(gd["lines"]) + "\n", encoding="utf-8")
    vocab = ["<pad>", "<bos>", "<eos>", "<unk>"] + list("abcdefghijklmnopqrstuvwxyz")
    tok = FrequencyAwareHyperbolicTokenizer(vocab, torch.nn.Parameter(synthetic_embeddings(30, 8)),
                                            corpus_path=str(corpus), max_vocab_size=64)
    assert tok.pair_frequencies == {(a, b): n for a, b, n in gd["counts"]}
    # save / load round trip of the extra files (reference :315-342)
    out = tmp_path / "saved"
    tok.save(str(out))
    back = FrequencyAwareHyperbolicTokenizer.load(str(out))
    ok = {k: v for k, v in tok.pair_frequencies.items() if "|" not in k[0] + k[1]}
    assert {k: back.pair_frequencies[k] for k in ok} == ok
    assert (back.alpha, back.beta, back.gamma) == (0.4, 0.4, 0.2)


def test_device_scoring_matches_host_scoring(tmp_path):
    """hyp_score_candidates (coherence mean in numpy's pairwise order, sigmoid and weighted score in float64 on the
    device) against the host path the class uses (numpy), same RNG draws: equal to the last bits of exp()."""
    import numpy as np
    import torch
    from hyptokenizer_b200.synth import synthetic_embeddings
    from hyptokenizer_b200.tokenizer.frequency_aware_hyperbolic_merge import FrequencyAwareHyperbolicTokenizer
    chars = list("abcdefghijklmnopqrstuvwxyz ")
    vocab = ["<pad>", "<bos>", "<eos>", "<unk>"] + chars + [a + b for a in "abcdef" for b in "xyz"]
    corpus = tmp_path / "c.txt"
    corpus.write_text("the quick brown fox jumps over the lazy dog\n" * 50 + "ax by cz ax ax\n" * 20)
    for sem, scale, thr in (("lorentz", 0.3, 2.5), ("lorentz", 0.05, 0.5), ("reference", 0.05, 0.5)):
        emb = synthetic_embeddings(len(vocab), 24, scale=scale, seed=4)
        tok = FrequencyAwareHyperbolicTokenizer(vocab, torch.nn.Parameter(emb), corpus_path=str(corpus),
                                                merge_threshold=thr, max_vocab_size=len(vocab) + 8, semantics=sem)
        cands = [(i, j, float(np.float32(0.01 * (i + j)))) for i in range(4, 40, 3) for j in range(i + 1, 49, 7)]
        cands += [(5, 6, 0.0), (7, 48, 1.5)]
        torch.manual_seed(77)
        host = tok._score_batch(cands)
        torch.manual_seed(77)
        dev = tok._score_batch_device(cands)
        assert len(host) == len(dev) == len(cands)
        for h, g in zip(host, dev):
            if h != h:
                assert g != g                                  # the shipped arithmetic: NaN scores stay NaN
            else:
                assert abs(h - g) <= 1e-13 * max(abs(h), 1.0), (sem, h, g)
