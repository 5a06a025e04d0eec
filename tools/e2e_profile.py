"""Host-side profile of the c2 end-to-end call (bench.py's e2e leg): where the time outside the kernels goes."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hyptokenizer_b200.synth import synthetic_embeddings, synthetic_vocab
from hyptokenizer_b200.tokenizer.fast_hyperbolic_merge import FastHyperbolicTokenizer

v0, d, target = 10000, 100, 50000
dev = torch.device("cuda", 0)
host_emb = synthetic_embeddings(v0, d, scale=0.01, seed=42).pin_memory()
vocab = synthetic_vocab(v0)

def run():
    tok = FastHyperbolicTokenizer(vocab, torch.nn.Parameter(host_emb), merge_threshold=0.1, max_vocab_size=target,
                                  device=dev, semantics="lorentz")
    tok.optimize_merges(steps=target - v0, log_every=10 ** 9, adaptive_threshold=True)
    out = tok.embeddings[: tok.current_vocab_size].detach().cpu()
    torch.cuda.synchronize()
    return tok, out

for _ in range(2):
    t0 = time.perf_counter(); run(); print("e2e s:", time.perf_counter() - t0)
pr = cProfile.Profile(); pr.enable(); run(); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
