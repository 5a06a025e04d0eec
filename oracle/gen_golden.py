#!/usr/bin/env python
"""Generate tests/golden/*.json by running the UNMODIFIED reference (CPU, torch eager).

TEST INFRASTRUCTURE ONLY.  Run in the build container, where /root/reference is
mounted read-only:

    PYTHONDONTWRITEBYTECODE=1 python oracle/gen_golden.py [--only NAME] [--full-c1]

The reference cannot travel to the GPU box, so the vectors it produces are
committed.  fp32 values are stored as uint32 bit patterns so that comparisons are
exact.  `semantics="reference"` is the shipped code; `semantics="lorentz"` is the
shipped code with the three functions of SURVEY.md Appendix B monkey-patched in
(distance / batch_distance / log_map) -- nothing else is touched.
"""
from __future__ import annotations

import argparse
import contextlib
import json
import os
import random
import sys
import warnings

import numpy as np

# /root/reference in the build container; bench.py --impl reference points this at the staged copy (baseline/_ref,
# written by tools/stage_reference.py) on the GPU box
REF = os.environ.get("HYP_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

warnings.filterwarnings("ignore")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
import torch  # noqa: E402

import embedding.lorentz_model as RL  # noqa: E402
import tokenizer.hyperbolic_merge as RH  # noqa: E402
import tokenizer.fast_hyperbolic_merge as RF  # noqa: E402
import tokenizer.frequency_aware_hyperbolic_merge as RQ  # noqa: E402
import tokenizer.hierarchical_hyperbolic_merge as RHI  # noqa: E402
import tokenizer.compression_aware_tokenizer as RCA  # noqa: E402
import tqdm as _tqdm  # noqa: E402

# EnhancedFastHyperbolicTokenizer does not import as shipped: it asks embedding.lorentz_model for two functions that
# do not exist.  Inject placeholders (never called on the path), then import it unmodified.
for _name in ("poincare_to_lorentz", "lorentz_to_poincare"):
    if not hasattr(RL, _name):
        setattr(RL, _name, lambda *a, **k: (_ for _ in ()).throw(NotImplementedError("shim")))
import tokenizer.enhanced_fast_hyperbolic_merge as REN  # noqa: E402
import tokenizer.adaptive_curvature_tokenizer as RAC  # noqa: E402

# silence progress bars
for _m in (RH, RF, RQ, RHI, RCA, REN, RAC):
    if hasattr(_m, "tqdm"):
        _m.tqdm = lambda it, **kw: _Quiet(it)


class _Quiet:
    def __init__(self, it):
        self.it = it

    def __iter__(self):
        return iter(self.it)

    def set_postfix(self, *a, **k):
        pass


_tqdm.tqdm = lambda it=None, **kw: _Quiet(it)


def bits(t) -> list:
    a = t.detach().cpu().contiguous().numpy() if isinstance(t, torch.Tensor) else np.asarray(t, dtype=np.float32)
    return a.astype(np.float32).view(np.uint32).reshape(-1).tolist()


def fbits(x: float) -> int:
    return int(np.array([x], dtype=np.float32).view(np.uint32)[0])


# ---- Appendix B patch ---------------------------------------------------------------------
def _lz_distance(x, y, c=1.0):
    u = torch.clamp(RL.minkowski_dot(x, y), min=1.0 + 1e-8)
    return torch.acosh(u) / torch.sqrt(torch.tensor(c, device=x.device, dtype=x.dtype))


def _lz_batch_distance(x, y, c=1.0):
    xr, yr = x.unsqueeze(1), y.unsqueeze(0)
    t = xr[..., 0] * yr[..., 0]
    s = torch.sum(xr[..., 1:] * yr[..., 1:], dim=-1)
    u = torch.clamp(t - s, min=1.0 + 1e-8)
    return torch.acosh(u) / torch.sqrt(torch.tensor(c, device=x.device, dtype=x.dtype))


def _lz_log_map(x, y, c=1.0):
    m = RL.minkowski_dot(x, y)
    u = torch.clamp(m, min=1.0 + 1e-8)
    coef = torch.clamp(torch.acosh(u) / torch.sqrt(u * u - 1), max=1e4)
    coef = torch.where(torch.isnan(coef) | (coef > 1e4), torch.ones_like(coef), coef)
    return coef.unsqueeze(-1) * (y - m.unsqueeze(-1) * x)


_PATCH = {"distance": _lz_distance, "batch_distance": _lz_batch_distance, "log_map": _lz_log_map,
          "distance_compiled": _lz_distance, "batch_distance_compiled": _lz_batch_distance}


def _lz_distance_grad(x, y, c=1.0):
    """_lz_distance that keeps the graph of a tensor `c` (the shipped `torch.tensor(c)` detaches it, which is why
    the reference's curvature step raises): the one-line fix the corrected adaptive-curvature step is pinned to."""
    u = torch.clamp(RL.minkowski_dot(x, y), min=1.0 + 1e-8)
    cc = c if isinstance(c, torch.Tensor) else torch.tensor(c, device=x.device, dtype=x.dtype)
    return torch.acosh(u) / torch.sqrt(cc)


@contextlib.contextmanager
def semantics(name: str):
    saved = []
    if name == "lorentz+grad":
        with semantics("lorentz"):
            old = REN.distance, RAC.distance
            REN.distance = RAC.distance = _lz_distance_grad
            try:
                yield
            finally:
                REN.distance, RAC.distance = old
        return
    if name == "lorentz":
        for mod in (RL, RH, RF, RQ, RHI, RCA, REN, RAC):
            for k, fn in _PATCH.items():
                if hasattr(mod, k):
                    saved.append((mod, k, getattr(mod, k)))
                    setattr(mod, k, fn)
    try:
        yield
    finally:
        for mod, k, fn in saved:
            setattr(mod, k, fn)


# ---- synthetic inputs --------------------------------------------------------
def synth_words(n=2000, seed=0):
    rng = random.Random(seed)
    return ["".join(rng.choice("abcdefghijklmnopqrstuvwxyz") for _ in range(rng.randint(2, 10))) for _ in range(n)]


def c1_vocab():
    chars = sorted(set("".join(synth_words())))
    return ["<pad>", "<bos>", "<eos>", "<unk>"] + chars


def set_seeds(seed=42):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)


def ref_init(n, d, scale=0.01, c=1.0):
    """scripts/train_hyperbolic_tokenizer.py:64-109 with the 0.01 made a parameter
    (the script itself is not importable here: it needs typer)."""
    tangent = torch.zeros((n, d + 1), dtype=torch.float32)
    tangent[:, 1:] = torch.randn((n, d), dtype=torch.float32) * scale
    origin = torch.zeros(d + 1, dtype=torch.float32)
    origin[0] = 1.0
    emb = torch.zeros((n, d + 1), dtype=torch.float32)
    for i, t in enumerate(tangent):
        emb[i] = RL.exp_map(origin, t, c)
    return RL.project_to_hyperboloid(emb, c)


def dump(name, obj):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name)
    with open(path, "w") as f:
        json.dump(obj, f, separators=(",", ":"))
    print("wrote", path, os.path.getsize(path), "bytes")


# ---- 1. pointwise / all-pairs known answers ---------------------------------------------------
def gen_lorentz_ops():
    cases = []
    for d in (3, 5, 50, 100):
        for scale in (0.01, 0.5):
            for c in (1.0, 0.7):
                set_seeds(1000 + d)
                n = 7
                X = ref_init(n, d, scale, c)
                case = {"d": d, "scale": scale, "c": c, "n": n, "X": bits(X)}
                P = torch.randn(n, d + 1)
                case["P"] = bits(P)
                case["project"] = bits(RL.project_to_hyperboloid(P, c))
                case["project_row"] = bits(RL.project_to_hyperboloid(P[2], c))
                case["mdot"] = bits(RL.minkowski_dot(X.unsqueeze(1), X.unsqueeze(0)))
                case["mnorm"] = bits(RL.minkowski_norm(X))
                ia = torch.tensor([0, 1, 2, 3, 4, 5, 0, 6])
                ib = torch.tensor([1, 2, 3, 4, 5, 6, 6, 6])
                case["ia"], case["ib"] = ia.tolist(), ib.tolist()
                for sem in ("reference", "lorentz"):
                    with semantics(sem):
                        r = {}
                        r["distance"] = bits(RL.distance(X[ia], X[ib], c))
                        r["batch_distance"] = bits(RL.batch_distance(X, X, c))
                        lg = RL.log_map(X[ia], X[ib], c)
                        r["log_map"] = bits(lg)
                        # midpoint exactly as tokenizer/hyperbolic_merge.py:323-340 (len 1 + len 3)
                        rows = []
                        for a, b in zip(ia.tolist(), ib.tolist()):
                            v = RL.log_map(X[a].unsqueeze(0), X[b].unsqueeze(0), c) * (3 / (1 + 3))
                            m = RL.exp_map(X[a].unsqueeze(0), v, c)[0]
                            rows.append(RL.project_to_hyperboloid(m, c))
                        r["midpoint_1_3"] = bits(torch.stack(rows))
                        case[sem] = r
                with semantics("lorentz"):
                    V = RL.log_map(X[ia], X[ib], c) * 0.4
                case["V"] = bits(V)
                case["exp_map"] = bits(RL.exp_map(X[ia], V, c))
                cases.append(case)
    dump("lorentz_ops.json", {"cases": cases})


# ---- 2. traces ----------------------------------------------------------------------------------
def trace_of(tok, merges_before=0):
    """(i, j) of each merge recovered from merge_history is ambiguous (duplicate strings),
    so traces are recorded by wrapping _merge_tokens."""
    raise NotImplementedError


def record_merges(tok):
    rec = []
    orig = tok._merge_tokens

    def wrapped(i, j):
        rec.append([int(i), int(j)])
        return orig(i, j)

    tok._merge_tokens = wrapped
    return rec


def run_base_loop(tok, steps, script_loop=False, target=None):
    """Either HyperbolicTokenizer.optimize_merges (hyperbolic_merge.py:357-412) or the loop of
    scripts/train_hyperbolic_tokenizer.py:236-283 (log callback omitted), driven from outside so
    the chosen distance and the candidate count are captured."""
    out = []
    for step in range(steps):
        if script_loop and target is not None and len(tok.vocab) >= target:
            break
        cands = tok._find_merge_candidates()
        if not cands:
            break
        ncand = len(cands)
        cands.sort(key=lambda x: x[2])
        i, j, d = cands[0]
        tok._merge_tokens(i, j)
        out.append([int(i), int(j), fbits(d), ncand])
        if script_loop and step > 0 and step % 1000 == 0:
            tok.merge_threshold *= 1.05
    return out


def tok_state(tok):
    n = tok.current_vocab_size
    return {"n": n, "vocab": list(tok.vocab), "merges": [list(m) for m in tok.merge_history],
            "embeddings": bits(tok.embeddings[:n]), "merge_threshold": tok.merge_threshold}


def gen_trace_test9():
    """tests/test_hyperbolic_tokenizer.py:24-63 fixture (9 tokens, d=5, seed 42, thr 0.5)."""
    out = {}
    for sem in ("reference", "lorentz"):
        torch.manual_seed(42)
        vocab = ["<pad>", "<bos>", "<eos>", "<unk>", "a", "b", "c", "d", "e"]
        d = 5
        tangent = torch.randn((len(vocab), d), dtype=torch.float32) * 0.01
        origin = torch.zeros(d + 1)
        origin[0] = 1.0
        emb = torch.zeros((len(vocab), d + 1))
        for i, t in enumerate(tangent):
            emb[i] = RL.exp_map(origin.unsqueeze(0), torch.cat([torch.zeros(1), t]).unsqueeze(0))[0]
        emb = RL.project_to_hyperboloid(emb)
        with semantics(sem):
            tok = RH.HyperbolicTokenizer(vocab, torch.nn.Parameter(emb.clone()), curvature=1.0,
                                         merge_threshold=0.5, lr=1e-3, device=torch.device("cpu"), max_vocab_size=64)
            tok.merge_threshold = 10.0
            cands = tok._find_merge_candidates()
            tok.merge_threshold = 0.5
            trace = run_base_loop(tok, 12)
        out[sem] = {"init": bits(emb), "d": d, "vocab0": vocab,
                    "candidates_thr10": [[i, j, fbits(x)] for i, j, x in cands],
                    "trace": trace, "final": tok_state(tok)}
    dump("trace_test9.json", out)


def gen_trace_c1(target=330, scales=(0.01, 0.3), name="trace_c1.json", sems=("reference", "lorentz")):
    """Config 1 generator: 30-token char vocab, d=50, thr 0.1, script loop."""
    vocab = c1_vocab()
    out = {"vocab0": vocab, "d": 50, "runs": []}
    for scale in scales:
        for sem in sems:
            set_seeds(42)
            emb = ref_init(len(vocab), 50, scale)
            thr = 0.1 if scale == 0.01 else 3.0
            with semantics(sem):
                tok = RH.HyperbolicTokenizer(vocab, torch.nn.Parameter(emb.clone()), curvature=1.0, merge_threshold=thr,
                                             lr=1e-3, device=torch.device("cpu"), max_vocab_size=1000)
                trace = run_base_loop(tok, 100000, script_loop=True, target=target)
            run = {"semantics": sem, "scale": scale, "threshold": thr, "target": target, "init": bits(emb),
                   "trace": trace, "final": tok_state(tok)}
            if target > 400:      # keep the big fixture small: the trace is what matters
                run["final"].pop("embeddings")
                run["final_embeddings_tail"] = bits(tok.embeddings[tok.current_vocab_size - 4:tok.current_vocab_size])
            out["runs"].append(run)
            print("c1", sem, scale, "merges", len(trace), "first", trace[:7])
    dump(name, out)


def gen_trace_fast():
    """FastHyperbolicTokenizer, 300 tokens, d=16, 250 steps, no FAISS."""
    out = {"runs": []}
    for sem, scale, thr in (("reference", 0.01, 0.1), ("lorentz", 0.01, 0.1), ("lorentz", 0.3, 5.0)):
        set_seeds(42)
        vocab = [f"w{k}" for k in range(300)]
        emb = ref_init(300, 16, scale)
        with semantics(sem):
            tok = RF.FastHyperbolicTokenizer(vocab, torch.nn.Parameter(emb.clone()), curvature=1.0, merge_threshold=thr,
                                             device=torch.device("cpu"), max_vocab_size=1024,
                                             use_approximate_search=False)
            rec = record_merges(tok)
            tok.optimize_merges(steps=250, log_every=1000)
        out["runs"].append({"semantics": sem, "scale": scale, "threshold0": thr, "init": bits(emb), "d": 16,
                            "merges_ij": rec, "final": tok_state(tok),
                            "cache_stats": {k: (float(v) if isinstance(v, float) else v)
                                            for k, v in tok.cache.get_stats().items()}})
        print("fast", sem, scale, len(rec), rec[:6])
    dump("trace_fast300.json", out)


def synth_corpus_lines(n_lines=400, seed=7):
    rng = random.Random(seed)
    alpha = "abcdefghijklmnopqrstuvwxyz"
    lines = []
    for k in range(n_lines):
        words = ["".join(rng.choice(alpha) for _ in range(rng.randint(1, 9))) for _ in range(rng.randint(0, 20))]
        line = " ".join(words)
        if k % 7 == 0:
            line = "  \t" + line + "   "
        if k % 11 == 0:
            line = line + " café über naïve 中文 \U0001F600"
        if k % 13 == 0:
            line = ""
        lines.append(line)
    return lines


def gen_pair_counts(tmpdir="/tmp"):
    lines = synth_corpus_lines()
    path = os.path.join(tmpdir, "hyp_golden_corpus.txt")
    with open(path, "w", encoding="utf-8") as f:
        f.write("\n".join(lines) + "\n")
    vocab = c1_vocab()
    set_seeds(42)
    emb = ref_init(len(vocab), 8)
    tok = RQ.FrequencyAwareHyperbolicTokenizer(vocab, torch.nn.Parameter(emb), corpus_path=path,
                                               device=torch.device("cpu"), max_vocab_size=256)
    counts = [[a, b, n] for (a, b), n in tok.pair_frequencies.items()]
    dump("pair_counts.json", {"lines": lines, "counts": counts})


def gen_trace_freq(tmpdir="/tmp"):
    """FrequencyAwareHyperbolicTokenizer.optimize_merges (:236-313), small."""
    lines = synth_corpus_lines(120, seed=3)
    path = os.path.join(tmpdir, "hyp_golden_corpus2.txt")
    with open(path, "w", encoding="utf-8") as f:
        f.write("\n".join(lines) + "\n")
    vocab = c1_vocab()
    out = {"lines": lines, "vocab0": vocab, "d": 8, "runs": []}
    for sem, scale, thr in (("reference", 0.01, 1.0), ("lorentz", 0.01, 0.03), ("lorentz", 0.3, 1.0)):
        set_seeds(42)
        emb = ref_init(len(vocab), 8, scale)
        with semantics(sem):
            tok = RQ.FrequencyAwareHyperbolicTokenizer(vocab, torch.nn.Parameter(emb.clone()), corpus_path=path,
                                                       merge_threshold=thr, device=torch.device("cpu"),
                                                       max_vocab_size=256)
            rec = record_merges(tok)
            torch.manual_seed(123)
            scores = []
            orig = tok._find_merge_candidates

            def spy():
                c = orig()
                scores.append([len(c), (float(c[0][2]) if c else None)])
                return c

            tok._find_merge_candidates = spy
            tok.optimize_merges(steps=8, log_every=10 ** 9)
        out["runs"].append({"semantics": sem, "scale": scale, "threshold0": thr, "init": bits(emb),
                            "merges_ij": rec, "best_neg_score": [[n, (None if s is None or s != s else s)] for n, s in scores],
                            "final": tok_state(tok)})
        print("freq", sem, scale, rec)
    dump("trace_freq.json", out)


def hier_corpus_lines(n_lines=300, seed=5):
    """Word-like text over a-z with a skewed word distribution, so that the hierarchical tokenizer's corpus
    statistics (common words / common n-grams) are not degenerate."""
    rng = random.Random(seed)
    stems = ["the", "and", "ing", "tion", "er", "re", "un", "al", "ed", "in", "on", "at", "or", "st", "an",
             "light", "house", "water", "stone", "green", "able", "ment", "ness", "play", "work", "read"]
    words = stems + [a + b for a in stems[:12] for b in stems[12:]]
    weights = [1.0 / (k + 1) for k in range(len(words))]
    return [" ".join(rng.choices(words, weights=weights, k=rng.randint(3, 14))) for _ in range(n_lines)]


def gen_trace_hier(tmpdir="/tmp"):
    """HierarchicalHyperbolicTokenizer._hierarchical_merge_strategy (hierarchical_hyperbolic_merge.py:279-428) with
    the three phase budgets shortened from 2000/5000/10000 to 25/30/30 steps (the ranges go through tqdm, which is
    stubbed to truncate), so that all three phases, the threshold widening and both re-weighting filters run in
    seconds.  NLTK is absent in this image: the corpus-statistics branches are the ones pinned."""
    assert not RHI.NLTK_AVAILABLE
    lines = hier_corpus_lines()
    path = os.path.join(tmpdir, "hyp_golden_corpus3.txt")
    with open(path, "w", encoding="utf-8") as f:
        f.write("\n".join(lines) + "\n")
    vocab = c1_vocab()
    budgets = {2000: 25, 5000: 30, 10000: 30}
    saved = _tqdm.tqdm
    _tqdm.tqdm = lambda it=None, **kw: _Quiet(list(it)[:budgets.get(len(it), len(it))])
    out = {"lines": lines, "vocab0": vocab, "d": 16, "phase_steps": [25, 30, 30], "runs": []}
    try:
        for sem, scale, target in (("reference", 0.3, None), ("lorentz", 0.3, None), ("lorentz", 0.02, 70)):
            set_seeds(42)
            emb = ref_init(len(vocab), 16, scale)
            with semantics(sem):
                tok = RHI.HierarchicalHyperbolicTokenizer(vocab, torch.nn.Parameter(emb.clone()), corpus_path=path,
                                                          device=torch.device("cpu"), max_vocab_size=400)
                rec = record_merges(tok)
                thr_log = []
                orig = tok._find_merge_candidates

                def spy():
                    c = orig()
                    thr_log.append([len(c), tok.merge_threshold])
                    return c

                tok._find_merge_candidates = spy
                tok.optimize_merges(hierarchical=True, target_vocab_size=target)
            out["runs"].append({"semantics": sem, "scale": scale, "target": target, "init": bits(emb),
                                "merges_ij": rec, "steps": thr_log, "final": tok_state(tok),
                                "common_morphemes": sorted(tok.common_morphemes),
                                "common_words": sorted(tok.common_words)})
            print("hier", sem, scale, target, "merges", len(rec), "steps", len(thr_log), rec[:6], "thr", tok.merge_threshold)
        # the string criteria on their own (they decide nothing in the collapsed traces above, where the repeated pair
        # has distance 0): verdicts for a list of strings and both re-weighting filters on a fixed candidate list
        probe_vocab = vocab + ["th", "he", "in", "ing", "tio", "wat", "er", "stone", "ligh", "t", "xq", "zzz", "andan", "re"]
        emb = ref_init(len(probe_vocab), 16, 0.3)
        tok = RHI.HierarchicalHyperbolicTokenizer(probe_vocab, torch.nn.Parameter(emb), corpus_path=path,
                                                  device=torch.device("cpu"), max_vocab_size=400)
        strings = sorted(set(probe_vocab[4:] + [a + b for a in probe_vocab[30:] for b in probe_vocab[30:]]))
        rng = random.Random(9)
        n = len(probe_vocab)
        cands = [(i, j, rng.random()) for i in range(4, n) for j in range(i + 1, n) if rng.random() < 0.25]
        out["criteria"] = {"vocab": probe_vocab, "strings": strings,
                           "morpheme": [bool(tok._is_potential_morpheme(t)) for t in strings],
                           "word": [bool(tok._is_valid_word(t)) for t in strings],
                           "candidates": [[i, j, d] for i, j, d in cands],
                           "morph_filtered": [[i, j, d] for i, j, d in tok._filter_morphologically_valid(cands)],
                           "word_filtered": [[i, j, d] for i, j, d in tok._filter_word_valid(cands)]}
        print("criteria:", sum(out["criteria"]["morpheme"]), "morphemes,", sum(out["criteria"]["word"]), "words of",
              len(strings), "strings;", len(cands), "candidates")
    finally:
        _tqdm.tqdm = saved
    dump("trace_hier.json", out)


def gen_trace_compress():
    """CompressionAwareTokenizer.optimize_merges (compression_aware_tokenizer.py:217-276), small: the first
    `sample_size` candidates are re-scored by the greedy longest-match token count of a corpus sample."""
    rng = random.Random(11)
    stems = ["the", "and", "ing", "tion", "er", "re", "un", "al", "ed", "in", "light", "house", "water", "stone"]
    sample = [" ".join(rng.choice(stems) + rng.choice(stems) for _ in range(rng.randint(2, 6))) for _ in range(14)]
    sample.append(sample[0][:20] + " tail that shares the cache key of the first text")     # the text[:20] quirk
    vocab = c1_vocab()
    out = {"sample": sample, "vocab0": vocab, "d": 8, "runs": []}
    for sem, scale, thr, ss in (("reference", 0.3, 0.1, 40), ("lorentz", 0.3, 1.2, 25), ("lorentz", 0.02, 0.05, 100)):
        set_seeds(42)
        emb = ref_init(len(vocab), 8, scale)
        with semantics(sem):
            tok = RCA.CompressionAwareTokenizer(vocab, torch.nn.Parameter(emb.clone()), corpus_sample=list(sample),
                                                sample_size=ss, merge_threshold=thr, device=torch.device("cpu"),
                                                max_vocab_size=256)
            rec = record_merges(tok)
            heads = []
            orig = tok._find_merge_candidates

            def spy():
                c = orig()
                heads.append([len(c), (float(c[0][2]) if c else None)])
                return c

            tok._find_merge_candidates = spy
            tok.optimize_merges(steps=10, log_every=10 ** 9)
        out["runs"].append({"semantics": sem, "scale": scale, "threshold0": thr, "sample_size": ss, "init": bits(emb),
                            "merges_ij": rec, "heads": [[n, (None if s is None or s != s else s)] for n, s in heads],
                            "final": tok_state(tok)})
        print("compress", sem, scale, rec, heads[:3])
    dump("trace_compress.json", out)


def gen_trace_enhanced(tmpdir="/tmp"):
    """EnhancedFastHyperbolicTokenizer.optimize_merges (enhanced_fast_hyperbolic_merge.py:1015-1209) under the import
    shim above, constructed directly.  Runs: all four features in the shipped arithmetic (below the first curvature
    step, which raises as shipped -- its error text is recorded from a second short run); feature subsets in the
    corrected geometry; and the corrected curvature step (`lorentz+grad`: `distance` keeps the graph of `c`)."""
    assert not REN.NLTK_AVAILABLE
    lines = hier_corpus_lines()
    path = os.path.join(tmpdir, "hyp_golden_corpus4.txt")
    with open(path, "w", encoding="utf-8") as f:
        f.write("\n".join(lines) + "\n")
    vocab = c1_vocab()
    sample = lines[:5] + [lines[0][:20] + " shares the cache key of the first text"] + lines[5:17]
    out = {"lines": lines, "sample": sample, "vocab0": vocab, "d": 16, "runs": []}
    flags_all = dict(use_frequency_aware=True, use_hierarchical=True, use_adaptive_curvature=True,
                     use_compression_aware=True)
    runs = [
        dict(sem="reference", scale=0.3, thr=0.5, steps=45, log_every=20, phases={2: 15, 3: 30}, flags=flags_all, kw={}),
        dict(sem="lorentz", scale=0.02, thr=0.5, steps=60, log_every=25, phases={2: 20, 3: 40},
             flags=dict(flags_all, use_adaptive_curvature=False), kw={}),
        dict(sem="lorentz", scale=0.3, thr=1.3, steps=30, log_every=10, phases=None,
             flags=dict(use_frequency_aware=True, use_hierarchical=False, use_adaptive_curvature=False,
                        use_compression_aware=True), kw={}),
        dict(sem="lorentz", scale=0.3, thr=1.3, steps=24, log_every=1000, phases=None,
             flags=dict(use_frequency_aware=False, use_hierarchical=False, use_adaptive_curvature=False,
                        use_compression_aware=True), kw=dict(adaptive_threshold=False)),
        dict(sem="lorentz+grad", scale=0.05, thr=0.5, steps=50, log_every=1000, phases={2: 20, 3: 35},
             flags=dict(use_frequency_aware=False, use_hierarchical=True, use_adaptive_curvature=True,
                        use_compression_aware=False), kw={}, ctor=dict(optimize_curvature_freq=8, curvature_lr=0.05)),
    ]
    for r in runs:
        set_seeds(42)
        emb = ref_init(len(vocab), 16, r["scale"])
        with semantics(r["sem"]):
            tok = REN.EnhancedFastHyperbolicTokenizer(
                vocab, torch.nn.Parameter(emb.clone()), merge_threshold=r["thr"], device=torch.device("cpu"),
                max_vocab_size=160, use_approximate_search=False, corpus_path=path, corpus_sample=list(sample),
                **r["flags"], **r.get("ctor", {}))
            rec = record_merges(tok)
            heads, curv = [], []
            orig = tok._find_merge_candidates_fast

            def spy():
                c = orig()
                b = c[0] if c else None
                nn = lambda v: None if v != v else float(v)
                heads.append([len(c), tok.merge_threshold, tok.current_phase] +
                             ([nn(b.combined_score), b.distance, float(b.frequency_score), nn(b.semantic_score),
                               b.compression_score, b.morphology_score] if b is not None else []))
                return c

            tok._find_merge_candidates_fast = spy
            if r["flags"]["use_adaptive_curvature"]:
                oc = tok._optimize_curvature

                def spy_c(e):
                    oc(e)
                    curv.append(float(tok.curvature.item()))

                tok._optimize_curvature = spy_c
            torch.manual_seed(123)
            random.seed(5)
            kw = dict(r["kw"])
            if r["phases"] is not None:
                kw["phase_transition_steps"] = r["phases"]
            tok.optimize_merges(steps=r["steps"], log_every=r["log_every"], **kw)
        stats = {str(k): v for k, v in getattr(tok, "training_stats", {}).items()}
        out["runs"].append({"semantics": r["sem"], "scale": r["scale"], "threshold0": r["thr"], "steps": r["steps"],
                            "log_every": r["log_every"], "phases": r["phases"], "flags": r["flags"],
                            "ctor": r.get("ctor", {}), "kw": r["kw"], "init": bits(emb), "merges_ij": rec,
                            "heads": heads, "curvatures": curv, "stats": stats,
                            "final_threshold": tok.merge_threshold, "final": tok_state(tok)})
        print("enhanced", r["sem"], r["flags"], "merges", len(rec), rec[:8], "heads", len(heads), "curv", curv)
    # the shipped curvature step: what it raises, and after how many merges
    set_seeds(42)
    emb = ref_init(len(vocab), 16, 0.3)
    tok = REN.EnhancedFastHyperbolicTokenizer(vocab, torch.nn.Parameter(emb.clone()), device=torch.device("cpu"),
                                              max_vocab_size=160, use_approximate_search=False,
                                              use_frequency_aware=False, use_compression_aware=False,
                                              optimize_curvature_freq=7)
    rec = record_merges(tok)
    try:
        tok.optimize_merges(steps=20, log_every=1000)
        err = None
    except RuntimeError as e:
        err = str(e)
    out["shipped_curvature_step"] = {"init": bits(emb), "freq": 7, "merges_ij": rec, "error": err}
    print("shipped curvature step:", len(rec), "merges, then", err)
    dump("trace_enhanced.json", out)


def gen_trace_adaptive():
    """AdaptiveCurvatureTokenizer.optimize_merges (adaptive_curvature_tokenizer.py:267-330) under the same import shim:
    the shipped arithmetic below its first curvature step, the error that step raises, and the corrected step
    (`lorentz+grad`).  The loop takes candidates[0] of the UNSORTED row-major list."""
    vocab = c1_vocab()
    out = {"vocab0": vocab, "d": 16, "runs": []}
    for sem, scale, thr, steps, freq, lr in (("reference", 0.3, 0.1, 12, 100, 0.01), ("reference", 0.3, 0.1, 12, 5, 0.01),
                                             ("lorentz+grad", 0.05, 0.3, 30, 6, 0.05), ("lorentz+grad", 0.3, 1.4, 25, 4, 0.2)):
        set_seeds(42)
        emb = ref_init(len(vocab), 16, scale)
        with semantics(sem):
            tok = RAC.AdaptiveCurvatureTokenizer(vocab, torch.nn.Parameter(emb.clone()), merge_threshold=thr,
                                                 device=torch.device("cpu"), max_vocab_size=120, curvature_lr=lr,
                                                 optimize_freq=freq)
            rec = record_merges(tok)
            curv, ncand = [], []
            oc, of = tok._optimize_curvature, tok._find_merge_candidates

            def spy_c(e):
                oc(e)
                curv.append(float(tok.curvature.item()))

            def spy_f():
                c = of()
                ncand.append([len(c)] + ([c[0][0], c[0][1], fbits(c[0][2])] if c else []))
                return c

            tok._optimize_curvature, tok._find_merge_candidates = spy_c, spy_f
            torch.manual_seed(321)
            try:
                tok.optimize_merges(steps=steps, log_every=10 ** 9)
                err = None
            except RuntimeError as e:
                err = str(e)
        out["runs"].append({"semantics": sem, "scale": scale, "threshold0": thr, "steps": steps, "optimize_freq": freq,
                            "curvature_lr": lr, "init": bits(emb), "merges_ij": rec, "candidates": ncand,
                            "curvatures": curv, "error": err, "final": tok_state(tok)})
        print("adaptive", sem, "merges", len(rec), rec[:8], "curv", curv, "err", err)
    dump("trace_adaptive.json", out)


def gen_persistence(tmpdir="/tmp"):
    """HyperbolicTokenizer.save (hyperbolic_merge.py:473-504) after five merges: sha256 and size of every file it writes.
    `embeddings.pt` is `torch.save` of a VIEW of the full `[max_vocab_size, D]` table, so the file carries the whole
    storage; torch.save is deterministic, so a byte-identical file is a meaningful target."""
    import hashlib
    import shutil
    vocab = c1_vocab()
    set_seeds(42)
    emb = ref_init(len(vocab), 8, 0.3)
    out = {"vocab0": vocab, "d": 8, "max_vocab_size": 64, "threshold": 1.2, "init": bits(emb), "torch": torch.__version__}
    with semantics("lorentz"):
        tok = RH.HyperbolicTokenizer(vocab, torch.nn.Parameter(emb.clone()), merge_threshold=1.2, device=torch.device("cpu"),
                                     max_vocab_size=64)
        out["merges"] = run_base_loop(tok, 5)
        path = os.path.join(tmpdir, "hyp_golden_saved")
        shutil.rmtree(path, ignore_errors=True)
        tok.save(path)
    out["files"] = {}
    for fn in sorted(os.listdir(path)):
        data = open(os.path.join(path, fn), "rb").read()
        out["files"][fn] = {"size": len(data), "sha256": hashlib.sha256(data).hexdigest()}
    out["final"] = tok_state(tok)
    print("persistence", out["files"])
    dump("persistence.json", out)


GENS = {"persistence": gen_persistence, "trace_adaptive": gen_trace_adaptive, "trace_enhanced": gen_trace_enhanced, "trace_compress": gen_trace_compress, "trace_hier": gen_trace_hier, "lorentz_ops": gen_lorentz_ops, "trace_test9": gen_trace_test9, "trace_c1": gen_trace_c1,
        "trace_fast": gen_trace_fast, "pair_counts": gen_pair_counts, "trace_freq": gen_trace_freq}

def gen_trace_c2_feasible(steps=300, v0=2000, d=100, scale=0.01, thr=0.1):
    """BASELINE config 2 at the largest size the reference's own brute-force loop runs (V0 = 2 000, d = 100; V0 = 10 000
    needs a 40 GB broadcast temporary): HyperbolicTokenizer.optimize_merges, the sequence the north star gates on, for
    `steps` merges on the benchmark's synthetic vocabulary (hyptokenizer_b200/synth.py, seed 42).  The init is not
    stored (800 KB of random mantissas): its sha256 is, and the test regenerates it."""
    import hashlib
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from hyptokenizer_b200.synth import synthetic_embeddings, synthetic_vocab
    emb = synthetic_embeddings(v0, d, scale=scale, seed=42)
    digest = hashlib.sha256(emb.numpy().tobytes()).hexdigest()
    out = {"v0": v0, "d": d, "scale": scale, "threshold": thr, "seed": 42, "init_sha256": digest, "runs": []}
    for sem in ("lorentz", "reference"):
        n_steps = steps if sem == "lorentz" else 3          # as shipped every step is (0, 1): 2e6 candidates in a Python loop
        with semantics(sem):
            tok = RH.HyperbolicTokenizer(synthetic_vocab(v0), torch.nn.Parameter(emb.clone()), curvature=1.0,
                                         merge_threshold=thr, lr=1e-3, device=torch.device("cpu"),
                                         max_vocab_size=v0 + n_steps + 8, use_approximate_search=False)
            trace = run_base_loop(tok, n_steps)
        n = tok.current_vocab_size
        out["runs"].append({"semantics": sem, "trace": trace, "vocab_tail": list(tok.vocab[v0:]),
                            "rows_tail": bits(tok.embeddings[n - 4:n])})
        print("c2-feasible", sem, "merges", len(trace), "first", trace[:4], "last candidates", trace[-1][3])
    dump("trace_c2_feasible.json", out)


if __name__ == "__main__":
    GENS["trace_c2_feasible"] = gen_trace_c2_feasible
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    ap.add_argument("--full-c1", action="store_true",
                    help="also write trace_c1_full.json: config 1 at full size (30 -> 1000), slow in lorentz semantics")
    a = ap.parse_args()
    torch.set_num_threads(8)
    if a.full_c1:
        gen_trace_c1(target=1000, scales=(0.01,), name="trace_c1_full.json")
    else:
        for k, fn in GENS.items():
            if a.only in (None, k):
                fn()
