"""Tuning aid for the tcgen05 Gram top-k: per-stage times (HYP_TC_TIMING) for bound-pass sampling steps (args: numbers)
or ablation masks (args: dN = HYP_TC_DEBUG=N; results are wrong under ablation, only the times mean something)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hyptokenizer_b200 import _lib
from hyptokenizer_b200._lib import SEM, check, ptr, stream_ptr
from hyptokenizer_b200.synth import synthetic_embeddings
n, k, D = 100000, 32, 101
E = synthetic_embeddings(n, 100, scale=0.01, seed=42, device="cuda")
L = _lib.lib()
ws = torch.empty(L.hyp_gram_topk_workspace_bytes(n, n, D), dtype=torch.uint8, device="cuda")
idx = torch.empty((n, k), dtype=torch.int32, device="cuda")
dd = torch.empty((n, k), dtype=torch.float32, device="cuda")
flags = torch.zeros(n, dtype=torch.int32, device="cuda")
os.environ["HYP_TC_TIMING"] = "1"
for arg in sys.argv[1:] or ["3"]:
    os.environ.pop("HYP_TC_DEBUG", None)
    if arg.startswith("d"):
        os.environ["HYP_TC_DEBUG"] = arg[1:]
    else:
        os.environ["HYP_TC_SUB"] = arg
    for rep in range(2):
        check(L.hyp_gram_topk(ptr(E), D, n, 0, n, D, 1.0, SEM["lorentz"], k, ptr(idx), ptr(dd), ptr(flags), ptr(ws), ws.numel(),
                              stream_ptr()))
        torch.cuda.synchronize()
    print(arg, "flagged", int(flags.sum()), flush=True)
