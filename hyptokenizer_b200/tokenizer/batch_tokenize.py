"""Batched tokenize / encode on the device.

`HyperbolicTokenizer.tokenize` (reference tokenizer/hyperbolic_merge.py:414-446) rewrites one Python
string at a time; `scripts/benchmark_efficiency.py:58-94` measures exactly that loop.  The methods
here apply the same rules to many texts in one kernel launch (csrc/apply_merges.cu), with the same
quirks: rules come from `_merge_rules` if it already exists (built once, possibly stale or empty),
otherwise they are built from `merge_history` the way the first `tokenize` call would; a later
duplicate of an `(old1, old2)` key overwrites an earlier one; unknown tokens encode to `<unk>`.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

from .. import _lib
from .._lib import check, ptr, stream_ptr


class RuleTable:
    """Host-built symbol ids + device rule hash table for one snapshot of the merge rules."""

    def __init__(self, rules: Dict[Tuple[str, str], str], vocab: Sequence[str], token2idx: Dict[str, int],
                 device: torch.device):
        self.strings: List[str] = []
        ids: Dict[str, int] = {}

        def sid(s: str) -> int:
            k = ids.get(s)
            if k is None:
                k = ids[s] = len(self.strings)
                self.strings.append(s)
            return k

        for (a, b), new in rules.items():
            sid(a), sid(b), sid(new)
        for tok in vocab:                      # every single-character vocabulary entry gets an id too,
            if len(tok) == 1:                  # so that a character without an id is always unknown
                sid(tok)
        self.ids = ids
        ascii_sym = np.full(128, -1, dtype=np.int32)
        other = []
        for s, k in ids.items():
            if len(s) == 1:
                cp = ord(s)
                if cp < 128:
                    ascii_sym[cp] = k
                else:
                    other.append((cp, k))
        other.sort()
        cap = 16
        while cap < 4 * max(1, len(rules)):
            cap *= 2
        keys = np.full(cap, np.uint64(0xFFFFFFFFFFFFFFFF), dtype=np.uint64)
        vals = np.full(cap, -1, dtype=np.int32)
        mask = cap - 1
        for (a, b), new in rules.items():
            key = (ids[a] << 32) | ids[b]
            h = ((key * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF) >> 32 & mask
            while keys[h] != np.uint64(0xFFFFFFFFFFFFFFFF) and int(keys[h]) != key:
                h = (h + 1) & mask
            keys[h] = np.uint64(key)
            vals[h] = ids[new]
        unk = token2idx.get("<unk>", 3)
        self.sym_to_vocab = np.array([token2idx.get(s, unk) for s in self.strings] or [unk], dtype=np.int64)
        self.unk = unk
        self.capacity = cap
        self.d_ascii = torch.from_numpy(ascii_sym).to(device)
        self.d_cp = torch.tensor([c for c, _ in other] or [0], dtype=torch.int64).to(torch.int32).to(device)
        self.d_cp_sym = torch.tensor([k for _, k in other] or [0], dtype=torch.int32).to(device)
        self.n_cp = len(other)
        self.d_keys = torch.from_numpy(keys.view(np.int64)).to(device)
        self.d_vals = torch.from_numpy(vals).to(device)
        self.device = device


def _rules_of(tok) -> Dict[Tuple[str, str], str]:
    """What tokenize() would use (hyperbolic_merge.py:425-428), including the build-once behaviour."""
    if not hasattr(tok, "_merge_rules"):
        tok._merge_rules = {}
        for old1, old2, new in tok.merge_history:
            tok._merge_rules[(old1, old2)] = new
    return tok._merge_rules


def apply_rules(tok, texts: Sequence[str]):
    """Run the kernel.  Returns (table, tokens int32 numpy (flat), starts, counts)."""
    table, d_tok, d_off, d_cnt, offsets, n = apply_rules_device(tok, texts)
    return table, d_tok.cpu().numpy(), offsets[:-1], d_cnt.cpu().numpy()[:n]


def count_token_pairs(tok, lines: Sequence[str]) -> Dict[Tuple[str, str], int]:
    """`_compute_pair_frequencies` (reference frequency_aware_hyperbolic_merge.py:92-112) for a tokenizer WITH merge
    rules: adjacent pairs of `tok.tokenize(line.strip())` over all lines, counted on the device -- batched rule
    application (hyp_apply_merges), then 64-bit pair keys, radix sort and run-length encode (hyp_pair_count_sorted).
    With empty rules this is the character-bigram count of pair_count.count_pairs."""
    texts = [ln.strip() for ln in lines]
    table, d_tok, d_off, d_cnt, offsets, n = apply_rules_device(tok, texts)
    n_slots = int(offsets[-1])
    if n == 0 or n_slots == 0:
        return {}
    dev = d_tok.device
    L = _lib.lib()
    nbytes = L.hyp_pair_count_sorted_workspace_bytes(n_slots)
    ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=dev)
    wsp = ws.data_ptr() + ((-ws.data_ptr()) % 256)
    cap = 1 << 16
    while True:
        keys = torch.empty(cap, dtype=torch.int64, device=dev)
        cnts = torch.empty(cap, dtype=torch.int64, device=dev)
        nu = torch.zeros(1, dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            check(L.hyp_pair_count_sorted(ptr(d_tok), ptr(d_off), ptr(d_cnt), n, n_slots, ptr(keys), ptr(cnts), cap,
                                          ptr(nu), wsp, nbytes, stream_ptr()))
        m = int(nu.item())
        if m <= cap:
            break
        cap = 1 << (m - 1).bit_length()
    k = keys[:m].cpu().numpy()
    c = cnts[:m].cpu().numpy()
    a = (k >> 32).astype(np.int32)
    b = (k & 0xFFFFFFFF).astype(np.uint32).view(np.int32) if m else np.empty(0, np.int32)
    strings = table.strings
    name = lambda v: strings[v] if v >= 0 else chr(-v - 1)
    out: Dict[Tuple[str, str], int] = {}
    for x, y, cnt in zip(a.tolist(), b.tolist(), c.tolist()):
        key = (name(x), name(y))
        out[key] = out.get(key, 0) + int(cnt)          # two ids never share a string; kept additive for safety
    return out


def apply_rules_device(tok, texts: Sequence[str]):
    """Run the kernel, results stay on the device: (table, tokens int32 [n_slots], offsets int64 [n+1], counts int32
    [n], host offsets, n)."""
    dev = tok.device
    rules = _rules_of(tok)
    cache = getattr(tok, "_rule_table", None)
    if cache is None or cache[0] is not rules or cache[1] != len(rules) or cache[2] != len(tok.vocab):
        table = RuleTable(rules, tok.vocab, tok.token2idx, dev)
        tok._rule_table = (rules, len(rules), len(tok.vocab), table)
    table = tok._rule_table[3]
    blobs = [t.encode("utf-8") for t in texts]
    lens = np.fromiter((len(b) for b in blobs), dtype=np.int64, count=len(blobs))
    offsets = np.zeros(len(blobs) + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    total = int(offsets[-1])
    host = np.frombuffer(b"".join(blobs), dtype=np.uint8).copy() if total else np.zeros(1, np.uint8)
    d_text = torch.from_numpy(host).to(dev)
    d_off = torch.from_numpy(offsets).to(dev)
    d_tok = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
    d_cnt = torch.zeros(max(len(blobs), 1), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.lib().hyp_apply_merges(ptr(d_text), ptr(d_off), len(blobs), ptr(table.d_ascii), ptr(table.d_cp),
                                          ptr(table.d_cp_sym), table.n_cp, ptr(table.d_keys), ptr(table.d_vals),
                                          table.capacity, ptr(d_tok), ptr(d_cnt), stream_ptr()))
    return table, d_tok, d_off, d_cnt, offsets, len(blobs)


def tokenize_batch(tok, texts: Sequence[str]) -> List[List[str]]:
    """[tokenize(t) for t in texts], one kernel launch."""
    table, flat, starts, counts = apply_rules(tok, texts)
    strings = table.strings
    out = []
    for s, c in zip(starts.tolist(), counts.tolist()):
        row = flat[s:s + c].tolist()
        out.append([strings[k] if k >= 0 else chr(-k - 1) for k in row])
    return out


def encode_batch(tok, texts: Sequence[str]) -> List[List[int]]:
    """[encode(t) for t in texts] (hyperbolic_merge.py:448-459), one kernel launch."""
    table, flat, starts, counts = apply_rules(tok, texts)
    ids = np.where(flat >= 0, table.sym_to_vocab[np.clip(flat, 0, len(table.sym_to_vocab) - 1)], table.unk)
    return [ids[s:s + c].tolist() for s, c in zip(starts.tolist(), counts.tolist())]
