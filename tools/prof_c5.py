"""cProfile of the config-5 host policy at V0 = 10 000 (40 steps, one cache refill): where the host time of
`bench.py --workload c5` goes.  Finding (round 2): one refill scores ~7e4 candidates at ~78 us each, of which 37 us are the
reference's own `torch.randperm(n)` per candidate (kept for RNG parity with the reference's coherence sample)."""
import cProfile, pstats, sys, io, argparse
sys.path.insert(0, '.')
import bench
from hyptokenizer_b200 import bench_c5
import torch
a = argparse.Namespace(c5_v0=10000, dim=100, c5_steps=40, c5_budget_s=200, semantics="lorentz")
class Env: pass
env = Env(); env.rank = 0; env.dev = torch.device("cuda", 0)
pr = cProfile.Profile()
pr.enable()
line = bench_c5.run(a, env)
pr.disable()
print(line["value"], line["config"]["merges_done"], line["config"]["host_s"])
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45); print(s.getvalue()[:9000])
