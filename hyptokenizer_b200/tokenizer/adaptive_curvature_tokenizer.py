"""Drop-in for the reference's ``tokenizer/adaptive_curvature_tokenizer.py`` on B200, and the curvature step it shares
with ``enhanced_fast_hyperbolic_merge.py`` (the two reference files carry the same two losses, :102-243 and :637-782).

As shipped, the curvature step cannot run: `distance` detaches `c` (`torch.tensor(c)`, reference
lorentz_model.py:137), so `loss.backward()` raises at the first step that is a multiple of the optimisation frequency
(probed) -- and the module does not import at all without two functions `embedding.lorentz_model` does
not have.  `semantics="reference"` keeps the shipped behaviour, error included.  In `semantics="lorentz"` the step is the
corrected one: every distance of the two losses is `acosh(u) / sqrt(c)`, with `acosh(u)` of ALL
sampled pairs from one device re-score (K3 at c = 1) and `c` a host scalar that autograd differentiates in closed form
(d/dc = -d / 2c per distance); the draws (`torch.randperm`, `torch.randint`) are consumed in the reference's order, the
update is the reference's Adam step and clamp to [0.1, 10], then the whole table is re-projected (K1).
"""
from __future__ import annotations

import json
import logging
import os
from typing import List, Optional, Tuple

import torch

from .. import _lib
from .._lib import SEM, check, ptr, stream_ptr
from ..embedding import lorentz_model as LM
from .hyperbolic_merge import HyperbolicTokenizer

logger = logging.getLogger(__name__)


class CurvatureStepMixin:
    """Needs: `curvature` (a scalar Parameter), `curvature_optimizer`, `merge_pairs`, `hierarchy_weight`,
    `distortion_weight`, and the tokenizer's `_table()` / `semantics`."""

    def _pair_acosh(self, ii: List[int], jj: List[int]) -> torch.Tensor:
        """acosh(clamped product) of the given row pairs: one K3 launch with c = 1, back as a host fp32 tensor."""
        E = self._table()
        k = len(ii)
        idx = torch.tensor([ii, jj], dtype=torch.int32).to(E.device)
        out = torch.empty(k, dtype=torch.float32, device=E.device)
        with torch.cuda.device(E.device):
            check(_lib.lib().hyp_rescore_pairs(ptr(E), E.stride(0), idx[0].data_ptr(), idx[1].data_ptr(), ptr(out),
                                               None, k, E.shape[1], 1.0, SEM[self.semantics], stream_ptr()))
        return out.cpu()

    def _draw_curvature_samples(self, n: int):
        """The draws of reference :651-688 and :719-733, in their order: per tracked merge pair a `torch.randperm(n)`
        head of 10 with i, j removed, then up to 500 `torch.randint` pairs (i != j kept)."""
        hier = []
        for i, j in self.merge_pairs[-min(len(self.merge_pairs), 100):]:
            if i >= n or j >= n:
                continue
            sample = torch.randperm(n)[:min(10, n - 2)]
            sample = sample[(sample != i) & (sample != j)].tolist()
            if sample:
                hier.append((i, j, sample))
        dist_pairs = []
        for _ in range(min(500, n * (n - 1) // 2)):
            a, b = torch.randint(0, n, (2,)).tolist()
            if a != b:
                dist_pairs.append((a, b))
        return hier, dist_pairs

    def _curvature_loss(self, c: torch.Tensor, hier, hier_acosh: torch.Tensor, dist_acosh: torch.Tensor):
        """hierarchy_weight * (:637-702) + distortion_weight * (:704-751) as a function of the scalar `c`; every
        distance is acosh(u) / sqrt(c).  `hier_acosh` holds, per tracked pair, [pair, i-vs-samples..., j-vs-samples...]."""
        root = torch.sqrt(c)
        h_loss = torch.zeros((), dtype=torch.float32)
        off = 0
        for _, _, sample in hier:
            m = len(sample)
            pair = hier_acosh[off] / root
            oi = hier_acosh[off + 1: off + 1 + m] / root
            oj = hier_acosh[off + 1 + m: off + 1 + 2 * m] / root
            off += 1 + 2 * m
            h_loss = h_loss + torch.relu(pair - oi + 0.1).mean() + torch.relu(pair - oj + 0.1).mean()
        num_pairs = min(len(self.merge_pairs), 100)
        if num_pairs > 0:
            h_loss = h_loss / (2 * num_pairs)
        if len(dist_acosh):
            d = dist_acosh / root
            d_loss = torch.exp(-10 * d.mean()) + 0.1 * d.var()
        else:
            d_loss = torch.zeros((), dtype=torch.float32)
        return self.hierarchy_weight * h_loss + self.distortion_weight * d_loss, h_loss, d_loss

    def _curvature_step(self) -> None:
        """reference adaptive_curvature_tokenizer.py:216-242 / enhanced_fast_hyperbolic_merge.py:753-782."""
        if self.semantics == "reference":
            # the shipped step: the two losses are built from distances whose `c` is detached, so backward() has nothing
            # to differentiate (probed on the reference: this is the error it raises at this point)
            raise RuntimeError("element 0 of tensors does not require grad and does not have a grad_fn")
        # the reference hands the step `self.embeddings.detach()`, the WHOLE table: its samples range over all
        # max_vocab_size rows (unused rows are the origin after the constructor's projection)
        n = self._table().shape[0]
        hier, dist_pairs = self._draw_curvature_samples(n)
        ii: List[int] = []
        jj: List[int] = []
        for i, j, sample in hier:
            ii += [i] + [i] * len(sample) + [j] * len(sample)
            jj += [j] + sample + sample
        nh = len(ii)
        ii += [a for a, _ in dist_pairs]
        jj += [b for _, b in dist_pairs]
        acosh = self._pair_acosh(ii, jj) if ii else torch.empty(0)
        c_host = self.curvature.detach().cpu().clone().requires_grad_(True)
        loss, h_loss, d_loss = self._curvature_loss(c_host, hier, acosh[:nh], acosh[nh:])
        self.curvature_optimizer.zero_grad()
        if loss.requires_grad:
            loss.backward()
            self.curvature.grad = c_host.grad.to(self.curvature.device)
            self.curvature_optimizer.step()
        with torch.no_grad():
            self.curvature.clamp_(min=0.1, max=10.0)
        self.last_curvature_loss = tuple(float(t.detach()) for t in (loss, h_loss, d_loss))
        logger.info("Optimized curvature: %.4f, Loss: %.4f (H: %.4f, D: %.4f)", self.curvature.item(),
                    *self.last_curvature_loss)


    def _project_table(self) -> None:
        """project_to_hyperboloid over the whole `[max_vocab_size, D]` table with the current curvature (unused rows
        become the origin), reference adaptive_curvature_tokenizer.py:244-249."""
        with torch.no_grad():
            self.embeddings.data = LM.project_to_hyperboloid(self._table(), LM._curv(self.curvature))


class AdaptiveCurvatureTokenizer(CurvatureStepMixin, HyperbolicTokenizer):
    """reference tokenizer/adaptive_curvature_tokenizer.py:31-437."""

    def __init__(self, vocab: List[str], embeddings: torch.nn.Parameter, curvature: float = 1.0,
                 merge_threshold: float = 0.1, lr: float = 1e-3, curvature_lr: float = 0.01,
                 device: Optional[torch.device] = None, max_vocab_size: int = 100000,
                 use_approximate_search: bool = True, hierarchy_weight: float = 1.0,
                 distortion_weight: float = 0.1, optimize_freq: int = 100, semantics: Optional[str] = None):
        super().__init__(vocab=vocab, embeddings=embeddings, curvature=1.0, merge_threshold=merge_threshold, lr=lr,
                         device=device, max_vocab_size=max_vocab_size,
                         use_approximate_search=use_approximate_search, semantics=semantics)
        self.curvature = torch.nn.Parameter(torch.tensor(float(curvature), device=self.device))
        self.curvature_optimizer = torch.optim.Adam([self.curvature], lr=curvature_lr)
        self.hierarchy_weight = hierarchy_weight
        self.distortion_weight = distortion_weight
        self.optimize_freq = optimize_freq
        self._project_table()
        self.merge_pairs: List[Tuple[int, int]] = []

    def _optimize_curvature(self, embeddings: Optional[torch.Tensor] = None) -> None:
        """reference :216-242.  `embeddings` is accepted for signature compatibility; the table is read in place."""
        self._curvature_step()

    def _project_embeddings(self) -> None:
        self._project_table()

    def _merge_tokens(self, i: int, j: int) -> None:
        """reference :251-265."""
        self.merge_pairs.append((i, j))
        super()._merge_tokens(i, j)

    def optimize_merges(self, steps: int = 10000, log_every: int = 1000, parallel_eval: bool = True,
                        sample_ratio: float = 1.0) -> None:
        """reference :267-330.  Note what the shipped loop does: the candidate list is NOT sorted here, so the merged
        pair is the first one under the threshold in row-major order (`_evaluate_candidates_parallel` returns
        `candidates[0]` too), found BEFORE the step's curvature update and re-projection."""
        self.last_trace: List[Tuple[int, int, float]] = []
        for step in range(steps):
            candidates = self._find_merge_candidates()
            if not candidates:
                logger.info(f"No more merge candidates found after {step} steps")
                break
            if step > 0 and step % self.optimize_freq == 0:
                self._optimize_curvature()
                self._project_embeddings()
            i, j, dist = candidates[0]
            self.last_trace.append((i, j, dist))
            self._merge_tokens(i, j)

    # ---- persistence (reference :332-437) ---------------------------------------------------------------------------
    def save(self, path: str) -> None:
        os.makedirs(path, exist_ok=True)
        with open(f"{path}/vocab.json", "w") as f:
            json.dump(self.vocab, f)
        torch.save(self.embeddings, f"{path}/embeddings.pt")          # the full table, as the reference saves it
        torch.save(self.curvature, f"{path}/curvature.pt")
        with open(f"{path}/merges.json", "w") as f:
            json.dump(self.merge_history, f)
        torch.save(self.merge_pairs, f"{path}/merge_pairs.pt")
        config = {"curvature": self.curvature.item(), "merge_threshold": self.merge_threshold,
                  "max_vocab_size": self.max_vocab_size, "use_approximate_search": self.use_approximate_search,
                  "hierarchy_weight": self.hierarchy_weight, "distortion_weight": self.distortion_weight,
                  "optimize_freq": self.optimize_freq}
        with open(f"{path}/config.json", "w") as f:
            json.dump(config, f)

    @classmethod
    def load(cls, path: str, device: Optional[torch.device] = None) -> "AdaptiveCurvatureTokenizer":
        """The active rows are the first `len(vocab)` of the saved table (the reference passes the full table to the
        constructor and fails on the shape)."""
        with open(f"{path}/vocab.json", "r") as f:
            vocab = json.load(f)
        embeddings = torch.load(f"{path}/embeddings.pt", map_location="cpu")
        curvature = torch.load(f"{path}/curvature.pt", map_location="cpu").item()
        with open(f"{path}/config.json", "r") as f:
            config = json.load(f)
        tokenizer = cls(vocab=vocab, embeddings=torch.nn.Parameter(embeddings.detach()[: len(vocab)].clone()),
                        curvature=curvature, merge_threshold=config["merge_threshold"], device=device,
                        max_vocab_size=config.get("max_vocab_size", 100000),
                        use_approximate_search=config.get("use_approximate_search", True),
                        hierarchy_weight=config.get("hierarchy_weight", 1.0),
                        distortion_weight=config.get("distortion_weight", 0.1),
                        optimize_freq=config.get("optimize_freq", 100))
        with open(f"{path}/merges.json", "r") as f:
            tokenizer.merge_history = [tuple(m) for m in json.load(f)]
        try:
            tokenizer.merge_pairs = [tuple(p) for p in torch.load(f"{path}/merge_pairs.pt")]
        except FileNotFoundError:
            logger.warning("Merge pairs file not found")
        return tokenizer
