"""Tuning aid for the tcgen05 Gram top-k: flagged rows and per-stage times for several bound-pass sampling steps."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hyptokenizer_b200.knn import lorentz_topk
from hyptokenizer_b200.synth import synthetic_embeddings
n, k = 100000, 32
E = synthetic_embeddings(n, 100, scale=0.01, seed=42, device="cuda")
os.environ["HYP_TC_TIMING"] = "1"
for sub in sys.argv[1:] or ["1", "2", "3", "4"]:
    os.environ["HYP_TC_SUB"] = sub
    for rep in range(2):
        ti, td = lorentz_topk(E, k, 1.0, "lorentz", engine="tc")
    print("sub", sub, "flagged", lorentz_topk.last_flagged, flush=True)
