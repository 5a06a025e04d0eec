"""bench.py --workload c5: BASELINE configs[4] at size -- EnhancedFastHyperbolicTokenizer (frequency-aware +
hierarchical + adaptive-curvature + compression-aware), d=100, on ONE GPU for a fixed number of steps, with the time of a
step split into device work (every C-ABI call, synchronised on both sides) and host policy (string scoring, sorting,
Python control flow).

What this measures and what it does not: the reference class (tokenizer/enhanced_fast_hyperbolic_merge.py:1015-1209)
cannot be imported as shipped and its curvature step raises, so there is no reference number for this
configuration; its algorithm is reproduced as a host policy over the device kernels (parity: tests/test_gpu_enhanced.py at
test sizes).  At size the step is HOST-bound by construction: every candidate the cache hands out is scored with Python
string heuristics (morphology, greedy longest-match compression over a corpus sample), and `semantics="lorentz"` runs the
corrected curvature step every `optimize_curvature_freq` merges with a full re-projection of the table.  The line reports
steps/s, the device/host split, and which kernels the device time went to.
"""
from __future__ import annotations

import os
import tempfile
import time

import numpy as np
import torch


class _DeviceTimer:
    """Wraps every hyp_* entry point of the loaded library: synchronise, call, synchronise, accumulate per name."""

    def __init__(self, L, names):
        self.L, self.names = L, names
        self.t = {}
        self.calls = {}
        self.saved = {}

    def __enter__(self):
        for name in self.names:
            fn = getattr(self.L, name)
            self.saved[name] = fn

            def timed(*a, _fn=fn, _name=name):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                rc = _fn(*a)
                torch.cuda.synchronize()
                self.t[_name] = self.t.get(_name, 0.0) + time.perf_counter() - t0
                self.calls[_name] = self.calls.get(_name, 0) + 1
                return rc

            setattr(self.L, name, timed)
        return self

    def __exit__(self, *exc):
        for name, fn in self.saved.items():
            setattr(self.L, name, fn)


def run(a, env) -> dict:
    from . import _lib
    from .synth import synthetic_corpus, synthetic_embeddings
    from .tokenizer.enhanced_fast_hyperbolic_merge import EnhancedFastHyperbolicTokenizer
    import random
    import signal
    rank, dev = env.rank, env.dev
    v0, d, steps = a.c5_v0, a.dim, a.c5_steps
    rng = random.Random(0)
    words = set()
    while len(words) < v0 - 31:
        words.add("".join(rng.choice("abcdefghijklmnopqrstuvwxyz") for _ in range(rng.randint(2, 10))))
    vocab = ["<pad>", "<bos>", "<eos>", "<unk>"] + list("abcdefghijklmnopqrstuvwxyz ") + sorted(words)
    vocab = vocab[:v0]
    emb = synthetic_embeddings(len(vocab), d, scale=0.01, seed=42 + rank)
    text = synthetic_corpus(4 << 20, seed=rank).tobytes().decode("ascii")
    lines = text.split("\n")
    fd, path = tempfile.mkstemp(suffix=".txt")
    with os.fdopen(fd, "w") as f:
        f.write(text)
    try:
        t0 = time.perf_counter()
        tok = EnhancedFastHyperbolicTokenizer(vocab, torch.nn.Parameter(emb), merge_threshold=0.1,
                                              max_vocab_size=len(vocab) + steps + 8, device=dev,
                                              use_approximate_search=False, corpus_path=path,
                                              corpus_sample=lines[:100], semantics=a.semantics,
                                              optimize_curvature_freq=100)
        torch.cuda.synchronize()
        t_init = time.perf_counter() - t0
        L = _lib.lib()
        names = [n for n in _lib.exported_symbols() if n.startswith("hyp_") and n not in
                 ("hyp_abi_version", "hyp_last_error", "hyp_check_device") and "workspace_bytes" not in n]
        torch.manual_seed(123)
        random.seed(5)
        class _Budget(Exception):
            pass

        def _alarm(signum, frame):
            raise _Budget()

        old_handler = signal.signal(signal.SIGALRM, _alarm)
        with _DeviceTimer(L, names) as timer:
            t0 = time.perf_counter()
            err = None
            signal.alarm(int(a.c5_budget_s))          # one refill scores every candidate in Python: bound the run
            try:
                tok.optimize_merges(steps=steps, log_every=10 ** 9, adaptive_threshold=True)
            except RuntimeError as e:                 # semantics="reference": the shipped curvature step raises, as in the reference
                err = str(e)
            except _Budget:
                err = f"wall budget of {a.c5_budget_s} s reached"
            finally:
                signal.alarm(0)
                signal.signal(signal.SIGALRM, old_handler)
            torch.cuda.synchronize()
            wall = time.perf_counter() - t0
    finally:
        os.remove(path)
    merges = len(tok.last_trace)
    dev_s = float(sum(timer.t.values()))
    top = sorted(timer.t.items(), key=lambda kv: -kv[1])[:6]
    return {"metric": "enhanced tokenizer merges/s (config 5, 1 GPU)", "value": merges / wall if wall > 0 else 0.0,
            "unit": "merges/s", "n_gpus": 1, "steps": steps, "warmup": 0, "ms_per_step": 1e3 * wall / max(merges, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"c5: EnhancedFastHyperbolicTokenizer, all four features, V0={len(vocab)}, d={d}, "
                                   f"{steps} steps, 4 MiB corpus for the pair frequencies, 100 sample lines for the "
                                   "compression score, cache_semantics=snapshot (the shipped pop-100 control flow)",
                       "semantics": a.semantics, "merges_done": merges, "stopped_by": err,
                       "constructor_s": t_init,
                       "device_s": dev_s, "host_s": wall - dev_s, "device_share": dev_s / wall if wall > 0 else None,
                       "device_calls": int(sum(timer.calls.values())),
                       "device_top": [{"entry": k, "s": v, "calls": timer.calls[k]} for k, v in top],
                       "candidates_last_refill": int(getattr(tok, "_last_candidate_total", -1)),
                       "note": "host-bound by construction: Python string scoring of every candidate of a cache refill (the "
                               "compression score re-tokenises the corpus sample per candidate, ~ms each), which is why the "
                               "default size is V0=2000: at V0=10 000 one refill holds ~10^5 candidates and takes minutes of "
                               "Python; device time includes the synchronisation around each call"},
            "gpu_launches": int(sum(timer.calls.values()))}
