"""Shared test helpers: bit-pattern <-> tensor conversion, synthetic inputs."""
import numpy as np
import torch


def from_bits(bits, *shape) -> torch.Tensor:
    a = np.asarray(bits, dtype=np.uint32).view(np.float32)
    t = torch.from_numpy(a.copy())
    return t.reshape(*shape) if shape else t


def to_bits(t) -> np.ndarray:
    if isinstance(t, torch.Tensor):
        t = t.detach().cpu().contiguous().numpy()
    return np.asarray(t, dtype=np.float32).view(np.uint32).reshape(-1)


def fbits(x: float) -> int:
    return int(np.array([x], dtype=np.float32).view(np.uint32)[0])


def same_bits(a, b) -> bool:
    """Bit equality, except that any NaN equals any NaN (payloads are not part of the contract)."""
    a, b = to_bits(a), to_bits(b)
    if a.shape != b.shape:
        return False
    fa, fb = a.view(np.float32), b.view(np.float32)
    return bool(np.all((a == b) | (np.isnan(fa) & np.isnan(fb))))


def ulp_diff(a, b) -> np.ndarray:
    """|a-b| in units in the last place of fp32 (monotone integer mapping); NaN==NaN -> 0."""
    a, b = to_bits(a).astype(np.int64), to_bits(b).astype(np.int64)
    fa = a.astype(np.uint32).view(np.float32)
    fb = b.astype(np.uint32).view(np.float32)
    ma = np.where(a & 0x80000000, 0x80000000 - a, a)
    mb = np.where(b & 0x80000000, 0x80000000 - b, b)
    d = np.abs(ma - mb)
    d[np.isnan(fa) & np.isnan(fb)] = 0
    return d


class RowInjector:
    """Tests-only strict-parity mode (hard part 1).

    The Minkowski products the kernels select on are bit-identical to the reference's; what is not reproducible
    between torch's CPU vector math and CUDA's libm is the last bit or two of acosh / cosh / sinh / sqrt, i.e. of
    every APPENDED row.  In tie-heavy regimes (duplicated midpoint rows whose mutual product is 1 or 1 + 2^-23
    depending on one bit of x0) that can flip a later choice.  To PROVE that a divergence has this cause and no
    other, the injector wraps `_merge_tokens`: after every merge it checks that the row the device appended is
    within `max_ulp` units in the last place OF THE ROW'S LARGEST ELEMENT of the reference's row (NaN pattern
    identical; the small spatial elements come out of a cancellation y - <x,y> x, so their own ulp is not the measure)
    and then REPLACES it by the reference's bits.  If the merge sequence, candidate counts and distances are then identical to the golden
    trace from the first step to the last, the only thing that separated the two runs was those last bits."""

    def __init__(self, tok, ref_rows: np.ndarray, n0: int, max_ulp: float = 16.0):
        self.max_seen = 0
        self.rows = 0
        self.max_ulp = max_ulp
        ref = np.asarray(ref_rows, dtype=np.float32)
        orig = tok._merge_tokens

        def wrapped(i, j):
            out = orig(i, j)
            r = tok.current_vocab_size - 1
            if n0 <= r < len(ref):
                got = tok.embeddings.data[r].detach().cpu().numpy()
                want = ref[r]
                assert np.array_equal(np.isnan(got), np.isnan(want)), f"row {r}: NaN pattern differs from the reference"
                ok = ~np.isnan(want)
                if ok.any():
                    scale = float(np.abs(want[ok]).max())
                    u = float(np.abs(got[ok].astype(np.float64) - want[ok]).max() / (scale * 2.0 ** -23))
                    self.max_seen = max(self.max_seen, u)
                    assert u <= self.max_ulp, f"row {r} is {u:.1f} ulp (of its scale) from the reference's row (> {self.max_ulp})"
                tok.embeddings.data[r] = torch.from_numpy(want.copy()).to(tok.embeddings.device)
                self.rows += 1
            return out

        tok._merge_tokens = wrapped
