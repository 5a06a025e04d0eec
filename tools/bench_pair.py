"""Micro-benchmark for the pair-count kernel variants (HYP_PAIR_COUNT=v1|v2): event-timed kernel GB/s on the config-4
stream and on a wider-alphabet stream, plus a digest of the counts so that two variants can be compared."""
import hashlib
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from hyptokenizer_b200 import _lib  # noqa: E402
from hyptokenizer_b200._lib import check, ptr  # noqa: E402
from hyptokenizer_b200.synth import english_corpus, synthetic_corpus  # noqa: E402


def wide_corpus(nbytes, seed=0):
    """Mixed-case text with digits and punctuation: ~75 distinct bytes, Zipf-like, lines of ~80 characters."""
    rng = np.random.default_rng(seed)
    alphabet = np.frombuffer((" etaoinshrdlcumwfgypbvkjxqz" "ETAOINSHRDLCUMWFGYPBVKJXQZ" "0123456789"
                              ".,;:!?'\"()-\t").encode(), np.uint8)
    w = 1.0 / (np.arange(len(alphabet)) + 1.5)
    out = rng.choice(alphabet, size=nbytes, p=w / w.sum())
    out[rng.integers(0, nbytes, nbytes // 80)] = 10
    return out


def run(name, host, iters=5):
    dev = torch.device("cuda", 0)
    text = torch.from_numpy(host).to(dev)
    cap = 1 << 20
    asc = torch.empty(128 * 128, dtype=torch.int64, device=dev)
    keys = torch.empty(cap, dtype=torch.int64, device=dev)
    vals = torch.empty(cap, dtype=torch.int64, device=dev)
    ovf = torch.empty(1, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream()
    L = _lib.lib()
    ts = []
    for it in range(iters + 2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        check(L.hyp_pair_count(ptr(text), text.numel(), ptr(asc), ptr(keys), ptr(vals), cap, ptr(ovf), st.cuda_stream))
        e1.record(st)
        torch.cuda.synchronize()
        if it >= 2:
            ts.append(e0.elapsed_time(e1))
    a = asc.cpu().numpy()
    print(f"{os.environ.get('HYP_PAIR_COUNT', 'default')} {name}: {text.numel() / (np.mean(ts) * 1e-3) / 1e9:8.1f} GB/s "
          f"({np.mean(ts):.3f} ms)  total={int(a.sum())} nz={int((a != 0).sum())} "
          f"digest={hashlib.sha1(a.tobytes()).hexdigest()[:12]}", flush=True)


if __name__ == "__main__":
    mb = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    run("c4-stream", synthetic_corpus(mb << 20, seed=0))
    run("english-letters", english_corpus(mb << 20))
    run("wide-alphabet", wide_corpus(mb << 20))
