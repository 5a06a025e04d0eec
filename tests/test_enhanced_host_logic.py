"""Host policy of EnhancedFastHyperbolicTokenizer (scoring weights, phases, loop control, RNG consumption, sort with
NaN scores, compression scoring, corrected curvature step) checked on the CPU against the golden traces of the
reference: every DEVICE touch point of the product class is replaced here by the oracle's restatement (test
infrastructure), so what runs is exactly the product's host code.  The same runs go through the real kernels in
tests/test_gpu_enhanced.py."""
import numpy as np
import pytest
import torch

import enhanced_common as EC
from oracle import lorentz as OL
from hyptokenizer_b200.embedding import lorentz_model as LM
from oracle import merge as OM

from hyptokenizer_b200.tokenizer import adaptive_curvature_tokenizer as PA
from hyptokenizer_b200.tokenizer import enhanced_fast_hyperbolic_merge as PE
from hyptokenizer_b200.tokenizer import hyperbolic_merge as PH


def _host_init(self, vocab, embeddings, curvature=1.0, merge_threshold=0.1, lr=1e-3, device=None,
               max_vocab_size=100000, use_approximate_search=True, semantics=None):
    self.device = torch.device("cpu")
    self.semantics = semantics or "reference"
    self.vocab = vocab.copy()
    self.current_vocab_size = len(vocab)
    self.max_vocab_size = max_vocab_size
    self.curvature = curvature
    self.merge_threshold = merge_threshold
    self.lr = lr
    self.use_approximate_search = False
    full = torch.zeros((max_vocab_size, embeddings.size(1)), dtype=torch.float32)
    full[: self.current_vocab_size] = embeddings.detach()
    self.embeddings = torch.nn.Parameter(full)
    self.token2idx = {tok: k for k, tok in enumerate(self.vocab)}
    self.merge_history = []
    self.index = None
    self._ws = None


def _host_merge(self, i, j):
    E = self.embeddings.data
    ti, tj = self.vocab[i], self.vocab[j]
    if self.current_vocab_size >= self.max_vocab_size:
        raise ValueError("Maximum vocabulary size reached")
    E[self.current_vocab_size] = OL.midpoint(E[i], E[j], len(ti), len(tj), LM._curv(self.curvature), self.semantics)
    self._append_token(ti, tj)


def _host_find(self):
    """HyperbolicTokenizer._find_merge_candidates: row-major (i, j, d) under the threshold."""
    n = self.current_vocab_size
    E = self.embeddings.data[:n]
    dist = OL.batch_distance(E, E, LM._curv(self.curvature), self.semantics)
    thr = torch.tensor(PH._threshold_f32(self.merge_threshold, n), dtype=torch.float32)
    keep = (dist < thr) & torch.triu(torch.ones(n, n, dtype=torch.bool), diagonal=1)
    ii, jj = keep.nonzero(as_tuple=True)
    return list(zip(ii.tolist(), jj.tolist(), dist[ii, jj].tolist()))


class _HostCurvature:
    def _table(self):
        return self.embeddings.data

    def _project_table(self):
        self.embeddings.data = OL.project_to_hyperboloid(self._table(), LM._curv(self.curvature))

    def _pair_acosh(self, ii, jj):
        E = self._table()
        return OL.distance(E[torch.tensor(ii)], E[torch.tensor(jj)], 1.0, self.semantics)


class HostAdaptive(_HostCurvature, PA.AdaptiveCurvatureTokenizer):
    pass


class HostEnhanced(_HostCurvature, PE.EnhancedFastHyperbolicTokenizer):

    def _candidate_arrays(self):
        n = self.current_vocab_size
        E = self._table()[:n]
        dist = OL.batch_distance(E, E, LM._curv(self.curvature), self.semantics)
        thr = torch.tensor(PH._threshold_f32(self.merge_threshold, n), dtype=torch.float32)
        keep = (dist < thr) & torch.triu(torch.ones(n, n, dtype=torch.bool), diagonal=1)
        ii, jj = keep.nonzero(as_tuple=True)
        dd = dist[ii, jj].numpy()
        order = np.argsort(dd, kind="stable")
        return ii.numpy()[order], jj.numpy()[order], dd[order]

    def _coherence_batch(self, cands):
        E, n = self._table(), self.current_vocab_size
        out = []
        with np.errstate(over="ignore", invalid="ignore"):
            for i, j, _ in cands:
                merged = OL.midpoint(E[i], E[j], len(self.vocab[i]), len(self.vocab[j]), LM._curv(self.curvature),
                                     self.semantics, project=False)
                idx = torch.randperm(n)[:min(50, n)]
                keep = [int(t) for t in idx if int(t) != i and int(t) != j]
                if not keep:
                    out.append(0.0)
                    continue
                d = OL.distance(merged, E[torch.tensor(keep)], LM._curv(self.curvature), self.semantics).tolist()
                out.append(1.0 / (1.0 + np.exp(np.mean(d) - self.merge_threshold)))
        return out

    def _compute_distance_statistics(self, sample_size=1000):
        import random
        n = self.current_vocab_size
        k = min(sample_size, n * (n - 1) // 2)
        pairs = [random.sample(range(n), 2) for _ in range(k)]
        if not pairs:
            return {"min": 0.0, "max": 0.0, "mean": 0.0, "std": 0.0}
        E = self._table()
        a, b = torch.tensor([p[0] for p in pairs]), torch.tensor([p[1] for p in pairs])
        d = OL.distance(E[a], E[b], LM._curv(self.curvature), self.semantics).tolist()
        return {"min": min(d), "max": max(d), "mean": np.mean(d), "std": np.std(d)}

    def _project_embeddings(self):
        if self.use_adaptive_curvature:
            self._project_table()


@pytest.fixture()
def host(monkeypatch, golden, tmp_path):
    monkeypatch.setattr(PH.HyperbolicTokenizer, "__init__", _host_init)
    monkeypatch.setattr(PH.HyperbolicTokenizer, "_merge_tokens", _host_merge)
    monkeypatch.setattr(PH.HyperbolicTokenizer, "_find_merge_candidates", _host_find)
    monkeypatch.setattr(PE, "count_pairs",
                        lambda data, device=None: OM.count_pairs_py(data.decode("utf-8").splitlines(True)))
    gd = golden("trace_enhanced.json")
    path = tmp_path / "corpus.txt"
    path.write_text("\n".join(gd["lines"]) + "\n", encoding="utf-8")
    return gd, str(path)


@pytest.mark.parametrize("run", [0, 1, 2, 3, 4])
def test_enhanced_host_policy_against_reference_trace(host, run):
    gd, path = host
    r = gd["runs"][run]
    tok, merges, heads, curv = EC.run_golden(HostEnhanced, gd, r, path)
    EC.check_run(tok, merges, heads, curv, gd, r, strict=True)


def test_save_load_round_trip(host, tmp_path):
    gd, path = host
    tok, *_ = EC.run_golden(HostEnhanced, gd, gd["runs"][4], path)
    EC.check_save_load(HostEnhanced, tok, gd, str(tmp_path / "saved"))


def test_shipped_curvature_step_raises_like_the_reference(host):
    gd, path = host
    s = gd["shipped_curvature_step"]
    vocab = gd["vocab0"]
    from helpers import from_bits
    emb = from_bits(s["init"], len(vocab), gd["d"] + 1)
    tok = HostEnhanced(vocab, torch.nn.Parameter(emb), max_vocab_size=160, use_approximate_search=False,
                       use_frequency_aware=False, use_compression_aware=False, optimize_curvature_freq=s["freq"],
                       semantics="reference")
    with pytest.raises(RuntimeError) as e:
        tok.optimize_merges(steps=20, log_every=1000)
    assert str(e.value) == s["error"]
    assert [[i, j] for i, j, _ in tok.last_trace] == s["merges_ij"]


def test_length_index_matches_the_reference_scan():
    """_LengthIndex.count == len(reference greedy longest match) (enhanced_fast_hyperbolic_merge.py:813-847)."""
    import random
    rng = random.Random(3)
    alphabet = "abcd "
    for _ in range(200):
        vocab = list(alphabet[: rng.randint(2, 5)]) + ["".join(rng.choice(alphabet) for _ in range(rng.randint(2, 4)))
                                                        for _ in range(rng.randint(0, 6))]
        extra = "".join(rng.choice(alphabet) for _ in range(rng.randint(1, 5)))
        text = "".join(rng.choice(alphabet + "z") for _ in range(rng.randint(0, 30)))
        ref_vocab = sorted(vocab + [extra], key=len, reverse=True)
        tokens, i = [], 0
        while i < len(text):
            for t in ref_vocab:
                if text[i:].startswith(t):
                    tokens.append(t)
                    i += len(t)
                    break
            else:
                tokens.append(text[i])
                i += 1
        assert PE._LengthIndex(vocab).count(text, extra) == len(tokens)
        assert PE._LengthIndex(vocab).tokenize(text, extra) == tokens
        assert PE._LengthIndex(vocab + [extra]).tokenize(text) == tokens


@pytest.mark.parametrize("run", [0, 1, 2, 3])
def test_adaptive_curvature_host_policy_against_reference_trace(host, golden, run, tmp_path):
    gd = golden("trace_adaptive.json")
    tok = EC.run_adaptive(HostAdaptive, gd, gd["runs"][run])
    if run == 3:
        EC.check_adaptive_save_load(HostAdaptive, tok, str(tmp_path / "saved"))


def test_saved_files_are_byte_identical_to_the_reference(host, golden, tmp_path):
    """HyperbolicTokenizer.save against tests/golden/persistence.json (sha256 of every file the reference wrote after
    five merges, incl. the view-of-the-full-table embeddings.pt).  Device touch points replaced by the
    oracle, whose rows are the reference's bit for bit; on the GPU the appended rows differ in transcendental last bits,
    so only this host-side run can be byte-identical."""
    import hashlib
    import os
    from helpers import fbits, from_bits
    gd = golden("persistence.json")
    if gd["torch"] != torch.__version__:
        pytest.skip("torch.save bytes are pinned to the torch build that wrote the fixture")
    vocab, d = gd["vocab0"], gd["d"]
    emb = from_bits(gd["init"], len(vocab), d + 1)

    class HostTok(PH.HyperbolicTokenizer):
        def _table(self):
            return self.embeddings.data

    tok = HostTok(vocab, torch.nn.Parameter(emb), merge_threshold=gd["threshold"], max_vocab_size=gd["max_vocab_size"],
                  semantics="lorentz")
    for want in gd["merges"]:
        cands = tok._find_merge_candidates()
        cands.sort(key=lambda c: c[2])
        i, j, dist = cands[0]
        assert [i, j, fbits(dist), len(cands)] == want
        tok._merge_tokens(i, j)
    out = str(tmp_path / "saved")
    tok.save(out)
    assert sorted(os.listdir(out)) == sorted(gd["files"])
    for fn, want in gd["files"].items():
        data = open(os.path.join(out, fn), "rb").read()
        assert len(data) == want["size"], fn
        assert hashlib.sha256(data).hexdigest() == want["sha256"], fn
    back = HostTok.load(out)
    assert back.vocab == tok.vocab and back.current_vocab_size == tok.current_vocab_size
    assert torch.equal(back.embeddings[: tok.current_vocab_size], tok.embeddings[: tok.current_vocab_size])
