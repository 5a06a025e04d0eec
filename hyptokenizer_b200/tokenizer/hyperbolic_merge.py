"""Drop-in for the reference's ``tokenizer/hyperbolic_merge.py`` on B200.

`HyperbolicTokenizer` keeps the reference's constructor, attributes (plain Python / torch
objects that callers may mutate), methods and on-disk format; the work underneath is done by
sm_100a kernels through the C ABI.  `optimize_merges` does not rebuild a Python candidate list
every step: the whole loop runs on the device (all-pairs argmin once, then one new row scored
against the table per merge), and the strings are rebuilt on the host from the merge log.

CUDA only.  No FAISS, no torch.compile, no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import json
import logging
import os
from typing import List, Optional, Tuple

import numpy as np
import torch

from .. import _lib
from .._lib import SEM, HypBest, HypMergeState, check, ptr, stream_ptr
from ..embedding import lorentz_model as LM
from ..embedding.lorentz_model import batch_distance, distance

logger = logging.getLogger(__name__)

FAISS_AVAILABLE = False     # the HNSW / IndexFlatL2 paths are replaced by exact device search
TORCH_COMPILE_AVAILABLE = False
USING_COMPILED = False

_RECORD_DTYPE = np.dtype([("i", "<i4"), ("j", "<i4"), ("d", "<f4"), ("n_new", "<i4")])
_SEGMENT = 8192   # merges per queued hyp_merge_steps call (see _device_loop)


def _threshold_f32(thr: float, n: int) -> float:
    """The fp32 value `t` such that the reference's test equals `d < t` for every fp32 d.
    n > 100: `all_dists < thr` is a tensor compare, thr rounds to nearest fp32
    (hyperbolic_merge.py:262).  n <= 100: `dist < thr` compares Python floats (:288), which for
    fp32 `dist` equals comparing with thr rounded UP to fp32."""
    t = np.float32(thr)
    if n > 100 or float(t) >= thr or not np.isfinite(t):
        return float(t)
    return float(np.nextafter(t, np.float32(np.inf)))


class HyperbolicTokenizer:
    """reference tokenizer/hyperbolic_merge.py:96-625."""

    def __init__(self, vocab: List[str], embeddings: torch.nn.Parameter, curvature: float = 1.0,
                 merge_threshold: float = 0.1, lr: float = 1e-3, device: Optional[torch.device] = None,
                 max_vocab_size: int = 100000, use_approximate_search: bool = True,
                 semantics: Optional[str] = None):
        if device is None:
            if not torch.cuda.is_available():
                raise RuntimeError("hyptokenizer_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
            device = torch.device("cuda", torch.cuda.current_device())
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError(f"hyptokenizer_b200 runs on CUDA only, got device={device}")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = device
        _lib.check_device(device)
        self.semantics = LM.get_semantics() if semantics is None else semantics
        if self.semantics not in SEM:
            raise ValueError(f"semantics must be one of {tuple(SEM)}")

        self.vocab = vocab.copy()
        self.current_vocab_size = len(vocab)
        self.max_vocab_size = max_vocab_size
        self.curvature = curvature
        self.merge_threshold = merge_threshold
        self.lr = lr
        self.use_approximate_search = False   # reference: `use_approximate_search and FAISS_AVAILABLE`

        if embeddings.dtype != torch.float32:
            raise TypeError("embeddings must be float32")
        full = torch.zeros((max_vocab_size, embeddings.size(1)), dtype=torch.float32, device=self.device)
        full[: self.current_vocab_size] = embeddings.detach().to(self.device)
        self.embeddings = torch.nn.Parameter(full)

        self.token2idx = {tok: k for k, tok in enumerate(self.vocab)}
        self.merge_history: List[Tuple[str, str, str]] = []
        self.index = None
        self._ws = None

    # ---------------------------------------------------------------- device plumbing
    def _table(self) -> torch.Tensor:
        E = self.embeddings.data
        if not E.is_cuda or E.dtype != torch.float32 or not E.is_contiguous():
            raise RuntimeError("`embeddings` must stay a contiguous float32 CUDA tensor")
        return E

    def _workspace(self):
        if self._ws is None or self._ws["dev"] != self._table().device:
            dev = self._table().device
            L = _lib.lib()
            self._ws = {
                "dev": dev,
                "allpairs": torch.empty(L.hyp_allpairs_workspace_bytes(0), dtype=torch.uint8, device=dev),
                "loop": torch.empty(L.hyp_merge_workspace_bytes(), dtype=torch.uint8, device=dev),
                "best": torch.empty(32, dtype=torch.uint8, device=dev),
                "state": torch.empty(40, dtype=torch.uint8, device=dev),
                "count": torch.empty(1, dtype=torch.int64, device=dev),
            }
        return self._ws

    def _read_best(self) -> HypBest:
        raw = self._workspace()["best"].cpu().numpy().tobytes()
        return HypBest.from_buffer_copy(raw)

    def _global_best(self, thr_f32: float) -> HypBest:
        """argmin over all pairs i<j<n of (d, i, j), and how many are below the threshold."""
        E, ws = self._table(), self._workspace()
        n, D = self.current_vocab_size, E.shape[1]
        with torch.cuda.device(E.device):
            check(_lib.lib().hyp_allpairs_min(ptr(E), E.stride(0), n, D, LM._curv(self.curvature), SEM[self.semantics],
                                              thr_f32, ptr(ws["best"]), ptr(ws["allpairs"]), ws["allpairs"].numel(),
                                              stream_ptr()))
        return self._read_best()

    def _initial_best(self, thr_f32: float) -> HypBest:
        """The pair the loop starts from.  Large tables in the corrected geometry go through the tcgen05 Gram top-k
        (k = 1: each row's nearest neighbour, exact after the fp32 re-score; the argmin pair (i, j), i < j, is always
        row i's nearest neighbour) instead of the O(n^2 d) CUDA-core scan: 97 ms -> 3 ms at n = 60 000.  The count of
        pairs under the threshold, which only `_find_merge_candidates` needs, is not computed on this path."""
        E = self._table()
        n = self.current_vocab_size
        if n >= 30000 and self.semantics == "lorentz" and E.shape[1] - 1 <= 124:   # below, the exact scan is ~1-10 ms
            from ..knn import best_pair_from_topk, lorentz_topk
            idx, d = lorentz_topk(E, 1, LM._curv(self.curvature), "lorentz", n, engine="tc")
            i, j, dv = best_pair_from_topk(idx, d)
            if dv == float("inf"):
                return HypBest(d=float("inf"), i=-1, j=-1, count_lo=0, count_hi=0)
            return HypBest(d=dv, i=i, j=j, count_lo=0, count_hi=0)
        return self._global_best(thr_f32)

    # ---------------------------------------------------------------- reference API
    def _compute_pairwise_distances(self) -> torch.Tensor:
        """reference hyperbolic_merge.py:166-190 -> (n, n) distance matrix."""
        a = self._table()[: self.current_vocab_size]
        return batch_distance(a, a, self.curvature, self.semantics)

    def _find_merge_candidates(self) -> List[Tuple[int, int, float]]:
        """reference hyperbolic_merge.py:192-291: every (i, j, d) with i<j, d<threshold, row-major."""
        E, ws = self._table(), self._workspace()
        n, D = self.current_vocab_size, E.shape[1]
        thr = _threshold_f32(self.merge_threshold, n)
        count = self._global_best(thr)
        total = count.count_lo | (count.count_hi << 32)
        if total == 0:
            return []
        oi = torch.empty(total, dtype=torch.int32, device=E.device)
        oj = torch.empty(total, dtype=torch.int32, device=E.device)
        od = torch.empty(total, dtype=torch.float32, device=E.device)
        with torch.cuda.device(E.device):
            check(_lib.lib().hyp_allpairs_emit(ptr(E), E.stride(0), n, D, LM._curv(self.curvature), SEM[self.semantics],
                                               thr, ptr(oi), ptr(oj), ptr(od), total, ptr(ws["count"]),
                                               stream_ptr()))
        order = torch.argsort(oi.to(torch.int64) * n + oj.to(torch.int64))
        return list(zip(oi[order].tolist(), oj[order].tolist(), od[order].tolist()))

    def _is_valid_merge(self, token_i: str, token_j: str) -> bool:
        """reference hyperbolic_merge.py:293-307."""
        return True

    def _merge_tokens(self, i: int, j: int) -> None:
        """reference hyperbolic_merge.py:309-355."""
        E = self._table()
        token_i, token_j = self.vocab[i], self.vocab[j]
        if self.current_vocab_size >= min(self.max_vocab_size, E.shape[0]):
            raise ValueError(f"Maximum vocabulary size {self.max_vocab_size} reached. Cannot merge more tokens.")
        n, D = self.current_vocab_size, E.shape[1]
        idx = torch.tensor([i, j, len(token_i), len(token_j)], dtype=torch.int32, device=E.device)
        with torch.cuda.device(E.device):
            check(_lib.lib().hyp_midpoint(ptr(E), E.stride(0), idx[0:].data_ptr(), idx[1:].data_ptr(),
                                          idx[2:].data_ptr(), idx[3:].data_ptr(), E[n].data_ptr(), D, 1, D,
                                          LM._curv(self.curvature), SEM[self.semantics], 1, stream_ptr()))
        self._append_token(token_i, token_j)

    merge = _merge_tokens   # north-star name for the same operation

    def _append_token(self, token_i: str, token_j: str) -> None:
        merged = token_i + token_j
        self.vocab.append(merged)
        self.token2idx[merged] = self.current_vocab_size
        self.current_vocab_size += 1
        self.merge_history.append((token_i, token_j, merged))

    # ---------------------------------------------------------------- device-resident loop
    def _device_loop(self, max_steps: int, step0: int = 0, threshold_every: int = 0,
                     threshold_mul: float = 1.0, best: Optional[HypBest] = None) -> Tuple[np.ndarray, int]:
        """Run up to `max_steps` merges on the device from the current table.  Returns the merge
        log (structured array) and the stop code (0 ran out of steps, 1 nothing below the
        threshold, 2 table full).  Host strings are updated from the log."""
        E, ws = self._table(), self._workspace()
        n0, D = self.current_vocab_size, E.shape[1]
        if max_steps <= 0:
            return np.empty(0, _RECORD_DTYPE), 0
        if best is None:
            best = self._initial_best(_threshold_f32(self.merge_threshold, n0))
        cap = min(self.max_vocab_size, E.shape[0])   # callers may have swapped `embeddings` for a smaller tensor
        lens = self._token_lengths(cap, n0, E.device)
        st = HypMergeState(threshold=float(self.merge_threshold), n=n0, capacity=cap,
                           best_d=best.d, best_i=best.i, best_j=best.j, steps_done=0, stop=0, pad=0)
        state = ws["state"]
        state.copy_(torch.frombuffer(bytearray(bytes(st)), dtype=torch.uint8))
        log = torch.empty((max_steps, 4), dtype=torch.int32, device=E.device)
        # A long run is queued as segments of _SEGMENT merges, with no host round trip in between: a call on a
        # stopped state is a no-op.  Each segment's log and state snapshot are copied to pinned memory behind it,
        # so the host rebuilds the strings of segment s (what _merge_tokens does per merge, :347-355) while the
        # device runs segment s+1.
        segs = [(off, min(_SEGMENT, max_steps - off)) for off in range(0, max_steps, _SEGMENT)]
        pin = self._pinned(max_steps, len(segs))
        h_log, h_state = pin["log"], pin["state"]
        events = []
        hint = min(cap, n0 + max_steps)
        with torch.cuda.device(E.device):
            for s, (off, cnt) in enumerate(segs):
                check(_lib.lib().hyp_merge_steps(ptr(E), E.stride(0), ptr(lens), D, LM._curv(self.curvature),
                                                 SEM[self.semantics], ptr(state), log[off:].data_ptr(), cnt,
                                                 step0 + off, threshold_every, float(threshold_mul), hint,
                                                 ptr(ws["loop"]), ws["loop"].numel(), stream_ptr()))
                h_state[s].copy_(state, non_blocking=True)
                h_log[off:off + cnt].copy_(log[off:off + cnt], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
                events.append(ev)
        vocab, t2i, hist, n = self.vocab, self.token2idx, self.merge_history, self.current_vocab_size
        recs, out = [], None
        for s, (off, cnt) in enumerate(segs):
            events[s].synchronize()
            out = HypMergeState.from_buffer_copy(h_state[s].numpy().tobytes())
            rec = h_log[off:off + out.steps_done].numpy().view(_RECORD_DTYPE).reshape(-1).copy()
            recs.append(rec)
            for i, j in zip(rec["i"].tolist(), rec["j"].tolist()):
                a, b = vocab[i], vocab[j]
                m = a + b
                vocab.append(m)
                t2i[m] = n
                n += 1
                hist.append((a, b, m))
            if out.stop != 0:
                break
        torch.cuda.current_stream(E.device).synchronize()   # the queued no-op segments behind a stop
        rec = np.concatenate(recs) if len(recs) > 1 else recs[0]
        self._last_state = out
        self._lens_valid = out.n          # the kernel kept len[n] = len[i] + len[j] for every row it appended
        self.merge_threshold = out.threshold if threshold_every > 0 else self.merge_threshold
        self.current_vocab_size = n
        assert self.current_vocab_size == out.n
        return rec, out.stop

    def _token_lengths(self, cap: int, n: int, device) -> torch.Tensor:
        """Device array of len(token) per row (the midpoint weights, reference :323-324).  Built from the host
        strings once and then maintained by the loop kernel; rebuilt only when the vocabulary changed behind its back
        (a host-side _merge_tokens, a caller editing `vocab`)."""
        cached = getattr(self, "_lens_dev", None)
        if (cached is not None and cached.shape[0] == cap and cached.device == device and
                getattr(self, "_lens_valid", -1) == n and len(self.vocab) == n):
            return cached
        lens = torch.zeros(cap, dtype=torch.int32)
        lens[:n] = torch.tensor([len(t) for t in self.vocab[:n]], dtype=torch.int32)
        self._lens_dev = lens.to(device, non_blocking=True)
        self._lens_valid = n
        return self._lens_dev

    def _pinned(self, max_steps: int, nseg: int):
        """Pinned landing buffers for the merge log and the per-segment state snapshots (grown on demand)."""
        pin = getattr(self, "_pin", None)
        if pin is None or pin["log"].shape[0] < max_steps or pin["state"].shape[0] < nseg:
            pin = {"log": torch.empty((max(max_steps, 1024), 4), dtype=torch.int32).pin_memory(),
                   "state": torch.empty((max(nseg, 8), 40), dtype=torch.uint8).pin_memory()}
            self._pin = pin
        return pin

    def _overrides_candidate_search(self) -> bool:
        return type(self)._find_merge_candidates is not HyperbolicTokenizer._find_merge_candidates

    def optimize_merges(self, steps: int = 10000, log_every: int = 1000, parallel_eval: bool = True,
                        sample_ratio: float = 1.0) -> None:
        """reference hyperbolic_merge.py:357-412.  `parallel_eval` / `sample_ratio` never change the
        chosen merge there (the parallel evaluation returns candidates[0], the sample keeps the
        head of the sorted list), so they are accepted and ignored."""
        if self._overrides_candidate_search():
            return self._host_loop(steps, log_every)
        rec, stop = self._device_loop(steps)
        self.last_trace = rec
        if stop == 1:
            logger.info("No more merge candidates found. Stopping.")
        elif stop == 2:
            raise ValueError(f"Maximum vocabulary size {self.max_vocab_size} reached. Cannot merge more tokens.")

    def train(self, merge_steps: int = 100000, target_vocab_size: Optional[int] = None,
              threshold_every: int = 1000, threshold_mul: float = 1.05) -> np.ndarray:
        """The loop of scripts/train_hyperbolic_tokenizer.py:236-283 (the only place the reference
        honours `target_vocab_size`), device resident: stop at the target size or when nothing is
        below the threshold; threshold *= 1.05 after every 1000th step."""
        steps = merge_steps
        if target_vocab_size is not None:
            steps = min(steps, max(0, target_vocab_size - len(self.vocab)))
        rec, stop = self._device_loop(steps, 0, threshold_every, threshold_mul)
        self.last_trace = rec
        if stop == 2:
            raise ValueError(f"Maximum vocabulary size {self.max_vocab_size} reached. Cannot merge more tokens.")
        return rec

    def _host_loop(self, steps: int, log_every: int) -> None:
        """Generic find -> stable sort -> merge loop for subclasses that re-rank candidates."""
        for step in range(steps):
            candidates = self._find_merge_candidates()
            if not candidates:
                logger.info("No more merge candidates found. Stopping.")
                break
            candidates.sort(key=lambda x: x[2])
            i, j, dist = candidates[0]
            self._merge_tokens(i, j)

    # ---------------------------------------------------------------- tokenize / encode / decode
    def tokenize(self, text: str) -> List[str]:
        """reference hyperbolic_merge.py:414-446, including the build-once `_merge_rules` quirk."""
        if not hasattr(self, "_merge_rules"):
            self._merge_rules = {}
            for old1, old2, new in self.merge_history:
                self._merge_rules[(old1, old2)] = new
        tokens = list(text)
        rules = self._merge_rules
        changed = True
        while changed:
            changed = False
            k = 0
            while k < len(tokens) - 1:
                new = rules.get((tokens[k], tokens[k + 1]))
                if new is not None:
                    tokens[k] = new
                    tokens.pop(k + 1)
                    changed = True
                else:
                    k += 1
        return tokens

    def tokenize_batch(self, texts) -> List[List[str]]:
        """`[self.tokenize(t) for t in texts]` in one device launch (csrc/apply_merges.cu)."""
        from .batch_tokenize import tokenize_batch
        return tokenize_batch(self, texts)

    def encode_batch(self, texts) -> List[List[int]]:
        """`[self.encode(t) for t in texts]` in one device launch."""
        from .batch_tokenize import encode_batch
        return encode_batch(self, texts)

    def encode(self, text: str) -> List[int]:
        """reference hyperbolic_merge.py:448-459."""
        unk = self.token2idx.get("<unk>", 3)
        return [self.token2idx.get(tok, unk) for tok in self.tokenize(text)]

    def decode(self, indices: List[int]) -> str:
        """reference hyperbolic_merge.py:461-471."""
        return "".join(self.vocab[idx] for idx in indices)

    # ---------------------------------------------------------------- persistence
    def save(self, path: str) -> None:
        """reference hyperbolic_merge.py:473-504: vocab.json, embeddings.pt, merges.json, config.json."""
        os.makedirs(path, exist_ok=True)
        with open(f"{path}/vocab.json", "w") as f:
            json.dump(self.vocab, f)
        # The reference saves `embeddings[:n].detach().cpu()`, on its CPU device a VIEW of the whole table: torch.save
        # writes the full storage and the file loads as [n, D].  Same here (table to the host first, then the view), so
        # that the file is the reference's byte for byte when the rows are (tests/golden/persistence.json).
        active = self.embeddings.detach().cpu()[: self.current_vocab_size]
        torch.save(active, f"{path}/embeddings.pt")
        with open(f"{path}/merges.json", "w") as f:
            json.dump(self.merge_history, f)
        config = {
            "curvature": self.curvature,
            "merge_threshold": self.merge_threshold,
            "embedding_dim": self.embeddings.size(1) - 1,
            "max_vocab_size": self.max_vocab_size,
            "use_approximate_search": self.use_approximate_search,
        }
        with open(f"{path}/config.json", "w") as f:
            json.dump(config, f)

    @classmethod
    def load(cls, path: str, device: Optional[torch.device] = None) -> "HyperbolicTokenizer":
        """reference hyperbolic_merge.py:506-551."""
        with open(f"{path}/vocab.json", "r") as f:
            vocab = json.load(f)
        embeddings = torch.load(f"{path}/embeddings.pt")
        with open(f"{path}/config.json", "r") as f:
            config = json.load(f)
        tokenizer = cls(vocab=vocab, embeddings=torch.nn.Parameter(embeddings), curvature=config["curvature"],
                        merge_threshold=config["merge_threshold"], device=device,
                        max_vocab_size=config.get("max_vocab_size", 100000),
                        use_approximate_search=config.get("use_approximate_search", True))
        with open(f"{path}/merges.json", "r") as f:
            tokenizer.merge_history = json.load(f)
        tokenizer.current_vocab_size = len(tokenizer.vocab)
        return tokenizer

    def _evaluate_candidates_parallel(self, candidates):
        """reference hyperbolic_merge.py:553-591 computes simulated midpoints and then returns
        candidates[0]; only the return value is observable."""
        return candidates[0]
