#!/usr/bin/env python
"""bench.py -- the HypTokenizer merge-loop hot path on B200 (BASELINE.json: "merges/sec at V=50k,d=100; all-pairs
Lorentz dist TFLOP/s + % roofline").

The default run measures, on N GPUs, every driver-visible figure of the path in ONE JSON line:

  c2 (headline; BASELINE configs[1])   FastHyperbolicTokenizer, d=100, V0=10 000 -> target_vocab_size=50 000: one
      "step" = one whole job (initial all-pairs argmin + 40 000 device-resident merges).
        value     merges/s, table resident in HBM, C-ABI calls only, CUDA events on the launch stream
        e2e       the same job through the public class: pinned host embeddings -> H2D, optimize_merges, D2H of
                  the merge log and of embeddings[:n], host string rebuild
        roofline  merge_loop_resident_kernel: algorithmic bytes sum_n 4 n (d+1) / its event-timed duration against
                  measured HBM copy bandwidth (+ smem_frac: the same bytes against the shared-memory bandwidth,
                  the bound that actually applies -- the table lives in shared memory)
      N > 1: the loop is sequential and fits one GPU ("replicas only"): N independent jobs.
  c3 (BASELINE configs[2])   all-pairs Lorentz distance + top-32 over V=100 000, d=100: tcgen05 TF32 Gram filter +
      fp32 re-score, rows SHARDED over the N ranks, the per-shard lists all-gathered INSIDE the timed region (the
      finishing kernels store into every rank's buffer over NVLink, hyp_ctx; NCCL all-gather timed beside it).
      config.c3_* / secondary.c3: ms, algorithmic TFLOP/s (2 V^2 (d+1) once), fraction of the TF32 peak measured
      here with torch.matmul 8192^3, recall against exact brute force and against Klein-L2 kNN.
  c4 (BASELINE configs[3], counting part)   pair counting over a 1 GiB stream per rank: GB/s and HBM fraction.

`--workload c2|c3|c4|tok|c2-sharded|c5` runs one of them alone with its own line.
`--impl reference` runs the UNMODIFIED reference (staged into baseline/_ref by tools/stage_reference.py) on the
host cores: FastHyperbolicTokenizer.optimize_merges at the largest size it can run (V0=2 000, d=100), one cache
cycle (101 merges) per step, measured, not scaled; the n=V0=10 000 figure is a cost model and labelled as such.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

V0, D_EMB, TARGET = 10000, 100, 50000
SCALE = 0.01            # the reference's init scale (scripts/train_hyperbolic_tokenizer.py:92)
THRESHOLD = 0.1
C3_V, C3_K = 100000, 32
REF_V0, REF_CYCLE = 2000, 101     # reference arm: the largest size the reference runs; one pop-100 cache cycle


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--semantics", default="lorentz", choices=["reference", "lorentz"])
    ap.add_argument("--v0", type=int, default=V0)
    ap.add_argument("--target", type=int, default=TARGET)
    ap.add_argument("--dim", type=int, default=D_EMB)
    ap.add_argument("--workload", default="all", choices=["all", "c2", "c3", "c4", "c5", "tok", "c2-sharded"],
                    help="all (default): c2 headline + c3 + c4 in one line; c2: merge loop (merges/s); c3: all-pairs "
                         "Lorentz distance + top-k=32 over V=100k, row-sharded, all-gather inside the timed region "
                         "(TFLOP/s); c4: pair counting; c5: enhanced tokenizer loop at size; tok: batched tokenize")
    ap.add_argument("--engine", default="tc", choices=["tc", "exact"])
    ap.add_argument("--exchange", default="auto", choices=["auto", "p2p", "nccl"])
    ap.add_argument("--ref-v0", type=int, default=REF_V0, help="vocabulary size the reference arm runs")
    ap.add_argument("--c3-min-steps", type=int, default=40)
    ap.add_argument("--c4-bytes", type=int, default=1 << 30)
    ap.add_argument("--c5-steps", type=int, default=150)
    ap.add_argument("--c5-v0", type=int, default=2000)
    ap.add_argument("--c5-budget-s", type=int, default=240)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-tf32-peak", action="store_true")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler:
    """nvidia-smi sampling DURING the timed region (B200_PROFILING.md clocks line).  nvidia-smi needs a few hundred ms
    before its first sample, longer than some timed regions: the process is started ahead of the warm-up (which runs
    the same kernels), `mark()` is called when the timed region begins, and only samples from then on are reported
    (all of them, flagged, if the region was too short to catch one)."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        self.t_mark = time.time()
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(gpu_index)], stdout=self.tmp,
                                         stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def mark(self):
        self.t_mark = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        import datetime
        t_end = time.time()
        time.sleep(0.1)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.tmp.flush()
        self.tmp.seek(0)
        rows = []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.tmp.read().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                ts = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(parts[1]), float(parts[2]),
                             {n for n, v in zip(names, parts[5:9]) if v.lower().startswith("active")}))
            except ValueError:
                continue
        os.unlink(self.tmp.name)
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        inside = [r for r in rows if self.t_mark - 0.05 <= r[0] <= t_end + 0.05]
        use = inside or rows[-3:]
        reasons = set().union(*[r[3] for r in use])
        out = {"sm_mhz": statistics.median(r[1] for r in use), "sm_max_mhz": max(r[2] for r in use),
               "reasons": sorted(reasons), "samples": len(use)}
        if not inside:
            out["note"] = "timed region shorter than the sampling interval: the last samples of the warm-up (same kernels)"
        return out


class Env:
    """One process per GPU (torchrun) or a single process: rank / world / device + NCCL for the plumbing."""

    def __init__(self):
        import torch.distributed as dist
        self.dist = dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        from hyptokenizer_b200 import _lib
        self.L = _lib.lib()
        _lib.check_device(self.dev)
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)   # > 126 MB L2

    def barrier(self):
        torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            torch.cuda.synchronize()

    def reduce(self, x: float, op: str) -> float:
        t = torch.tensor([x], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX if op == "max" else self.dist.ReduceOp.SUM)
        return float(t.item())

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def algorithmic_bytes(v0: int, merges: int, d: int) -> float:
    """each merge reads the n current rows once: 4*n*(d+1) bytes (+ one row written)."""
    n_sum = merges * v0 + merges * (merges - 1) // 2
    return 4.0 * (d + 1) * (n_sum + merges)


def measure_tf32_peak(dev) -> dict:
    """Dense TF32 throughput measured the way MEASURED_PEAKS.json measures bf16: torch.matmul
    8192^3 with TF32 allowed, best of 10 (burst) and back to back for 2 s (sustained), CUDA events."""
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        n = 8192
        a = torch.randn((n, n), device=dev)
        b = torch.randn((n, n), device=dev)
        c = torch.empty((n, n), device=dev)
        flops = 2.0 * n ** 3
        for _ in range(3):
            torch.matmul(a, b, out=c)
        torch.cuda.synchronize()
        best = float("inf")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(10):
            e0.record()
            torch.matmul(a, b, out=c)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        reps = max(10, int(2000.0 / best))
        e0.record()
        for _ in range(reps):
            torch.matmul(a, b, out=c)
        e1.record()
        torch.cuda.synchronize()
        sus = e0.elapsed_time(e1) / reps
        del a, b, c
        return {"tf32_tflops": flops / (best * 1e-3) / 1e12, "tf32_tflops_sustained": flops / (sus * 1e-3) / 1e12,
                "how": "torch.matmul fp32 8192^3 with allow_tf32 (2*N^3): best of 10 (burst), back to back for ~2 s (sustained)"}
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


# ------------------------------------------------------------------------------------------------
# c2: the merge loop (headline)
# ------------------------------------------------------------------------------------------------
def bench_c2(a, env: Env) -> dict:
    from hyptokenizer_b200._lib import SEM, HypMergeState, check, ptr
    from hyptokenizer_b200.synth import synthetic_embeddings, synthetic_vocab
    from hyptokenizer_b200.tokenizer.fast_hyperbolic_merge import FastHyperbolicTokenizer
    L, dev, rank, world = env.L, env.dev, env.rank, env.world
    v0, d, target = a.v0, a.dim, a.target
    merges = target - v0
    D = d + 1
    sem = SEM[a.semantics]
    host_emb = synthetic_embeddings(v0, d, scale=SCALE, seed=42 + rank).pin_memory()
    init = host_emb.to(dev)
    E = torch.zeros((target, D), dtype=torch.float32, device=dev)
    lens0 = torch.tensor([len(t) for t in synthetic_vocab(v0)], dtype=torch.int32)
    lens_init = torch.zeros(target, dtype=torch.int32)
    lens_init[:v0] = lens0
    lens_init = lens_init.to(dev)
    lens = torch.empty_like(lens_init)
    ws_ap = torch.empty(L.hyp_allpairs_workspace_bytes(v0), dtype=torch.uint8, device=dev)
    ws_lp = torch.empty(L.hyp_merge_workspace_bytes(), dtype=torch.uint8, device=dev)
    best = torch.empty(32, dtype=torch.uint8, device=dev)
    state = torch.empty(40, dtype=torch.uint8, device=dev)
    log = torch.empty((merges, 4), dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream()
    sp = stream.cuda_stream
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    thr32 = float(np.float32(THRESHOLD))

    def one_step(timed):
        # reset the table to the initial vocabulary (not part of the path; outside the events)
        E[:v0].copy_(init)
        E[v0:].zero_()
        lens.copy_(lens_init)
        env.flush.fill_(1)                               # L2 flush between timed iterations
        ev[0].record(stream)
        check(L.hyp_allpairs_min(ptr(E), D, v0, D, 1.0, sem, thr32, ptr(best), ptr(ws_ap), ws_ap.numel(), sp))
        check(L.hyp_merge_state_init(ptr(state), ptr(best), v0, target, THRESHOLD, sp))
        ev[1].record(stream)
        check(L.hyp_merge_steps(ptr(E), D, ptr(lens), D, 1.0, sem, ptr(state), ptr(log), merges, 0, 1000, 1.1,
                                target, ptr(ws_lp), ws_lp.numel(), sp))
        ev[2].record(stream)
        if not timed:
            return None
        torch.cuda.synchronize()
        return ev[0].elapsed_time(ev[2]), ev[1].elapsed_time(ev[2])

    sampler = ClockSampler(env.local) if rank == 0 else None
    for _ in range(a.warmup):
        one_step(False)
    env.barrier()
    if sampler:
        sampler.mark()
    t_total, t_loop = [], []
    for _ in range(a.steps):
        tt, tl = one_step(True)
        t_total.append(tt)
        t_loop.append(tl)
    env.barrier()
    clocks = sampler.stop() if sampler else None
    prof = ws_lp[32:96].cpu().numpy().view(np.int64)
    mprof = ws_lp[96:160].cpu().numpy().view(np.int64)
    if rank == 0 and mprof.any():      # only with -DHYP_LOOP_PROF builds
        print(f"[scan marks cta0] cycles/merge: {[round(int(v) / max(int(prof[3]), 1)) for v in mprof]}", file=sys.stderr)
    if rank == 0 and prof[3] > 0:
        for name, base in (("cta0", 0), ("ctaN", 4)):
            n_ = max(int(prof[base + 3]), 1)
            print(f"[phases {name}] cycles/merge: midpoint={prof[base] / n_:.0f} scan={prof[base + 1] / n_:.0f} "
                  f"barrier={prof[base + 2] / n_:.0f}", file=sys.stderr)
    st = HypMergeState.from_buffer_copy(state.cpu().numpy().tobytes())
    done = st.steps_done
    ms_total = float(sum(t_total))
    tmax = env.reduce(ms_total, "max")
    done_all = env.reduce(done * a.steps, "sum")
    value = done_all / (tmax * 1e-3)

    # ---- end to end through the public class, host buffers, copies inside the timed region ----------
    e2e = None
    if not a.no_e2e:
        vocab = synthetic_vocab(v0)
        out = torch.empty((target, D), dtype=torch.float32).pin_memory()    # landing buffer of the result table
        e2e_t = []
        n_e2e = n_rows = 0
        for it in range(2):
            env.barrier()
            t0 = time.perf_counter()
            tok = FastHyperbolicTokenizer(vocab, torch.nn.Parameter(host_emb), merge_threshold=THRESHOLD,
                                          max_vocab_size=target, device=dev, semantics=a.semantics)
            tok.optimize_merges(steps=merges, log_every=10 ** 9, adaptive_threshold=True)
            n_rows = tok.current_vocab_size
            out[:n_rows].copy_(tok.embeddings.detach()[:n_rows], non_blocking=True)
            torch.cuda.synchronize()
            e2e_t.append(time.perf_counter() - t0)
            n_e2e = len(tok.last_trace)
        te = env.reduce(min(e2e_t), "max")
        ne = env.reduce(float(n_e2e), "sum")
        e2e = {"value": ne / te, "unit": "merges/s",
               "h2d_bytes_per_step": int(host_emb.numel() * 4 + target * 4 + 40),
               "d2h_bytes_per_step": int(n_e2e * 16 + n_rows * D * 4 + 40 * ((n_e2e + 8191) // 8192))}

    pk, pk_kind = peaks()
    loop_ms = float(np.mean(t_loop))
    abytes = algorithmic_bytes(v0, done, d)
    achieved = abytes / (loop_ms * 1e-3) / 1e9
    smem_peak = 148 * 128 * float(pk.get("sm_max_mhz", 1965.0)) * 1e6 / 1e9     # GB/s: 128 B/clk/SM
    line = {
        "metric": "merges/sec at V=50k,d=100", "value": value, "unit": "merges/s", "n_gpus": world,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_total / a.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"c2: FastHyperbolicTokenizer d={d}, V0={v0} -> target_vocab_size={target} "
                               f"({merges} merges/step, exact device search in place of HNSW M=32/ef=100)",
                   "semantics": a.semantics, "init_scale": SCALE, "merge_threshold": THRESHOLD,
                   "merges_done_per_step": done, "stop_code": st.stop,
                   "l2": "flushed between timed steps (256 MiB write); within a step the table is on-chip by design",
                   "parallelism": "replicas only" if world > 1 else "1 GPU"},
        "gpu_launches": 3 * a.steps,
        "launches_note": "per step: allpairs_tile_kernel<min>, merge_state_init_kernel, "
                         "merge_loop_resident_kernel<100, false> (cooperative, persistent)",
        "roofline": {"bound": "hbm", "kernel": "merge_loop_resident_kernel<100, false>", "achieved": achieved,
                     "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": achieved / pk["hbm_gbs"],
                     "smem_frac": achieved / smem_peak, "smem_peak_gbs": smem_peak,
                     "traffic": 4107776, "traffic_source": "profiles/r01_prof_merge_raw.csv: dram__bytes_read.sum + "
                                                           "dram__bytes_write.sum of one launch (ncu --set full)",
                     "peak_kind": pk_kind,
                     "algorithmic_bytes_per_launch": abytes, "kernel_ms": loop_ms,
                     "note": "the table (<= 20.2 MB) lives in the shared memory of the persistent grid, so DRAM traffic is far "
                             "below the algorithmic bytes: frac is the algorithmic rate against HBM copy bandwidth (an "
                             "equivalence), smem_frac the same rate against 148 SMs x 128 B/clk, the bound that applies"},
        "clocks": clocks,
    }
    if e2e:
        line["e2e"] = e2e
    return line


# ------------------------------------------------------------------------------------------------
# c2 in the SHIPPED control flow: FastHyperbolicTokenizer(cache_semantics="snapshot"), the pop-100 cache of the reference
# ------------------------------------------------------------------------------------------------
def bench_snapshot(a, env: Env, steps: int = 505) -> dict:
    """The reference's FastHyperbolicTokenizer merges from a stale cache: one all-pairs candidate search per ~101
    merges, 10 000 candidates kept (fast_hyperbolic_merge.py:63-133, :253-376).  `cache_semantics="snapshot"` replays
    exactly that over the device kernels; this is the like-for-like of what `--impl reference` times.  Five cache cycles
    at the benchmark's V0, both arithmetics: in the shipped one every distance is 0.0, all n (n - 1) / 2 pairs are
    candidates and a refill is the device radix select over 5e7 ties (hyp_allpairs_hist / _row_ties / _emit_cut)."""
    import random
    from hyptokenizer_b200.synth import synthetic_embeddings, synthetic_vocab
    from hyptokenizer_b200.tokenizer.fast_hyperbolic_merge import FastHyperbolicTokenizer
    out = {}
    for sem in ("lorentz", "reference"):
        emb = synthetic_embeddings(a.v0, a.dim, scale=SCALE, seed=42)
        random.seed(42)
        tok = FastHyperbolicTokenizer(synthetic_vocab(a.v0), torch.nn.Parameter(emb), merge_threshold=THRESHOLD,
                                      max_vocab_size=a.v0 + steps + 8, device=env.dev, semantics=sem,
                                      cache_semantics="snapshot")
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        tok.optimize_merges(steps=steps, log_every=10 ** 9)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t1 = time.perf_counter()
        ii, jj, dd = tok._candidate_arrays()
        torch.cuda.synchronize()
        refill = time.perf_counter() - t1
        out[sem] = {"merges": len(tok.last_trace), "merges_per_s": len(tok.last_trace) / dt,
                    "candidates_of_a_refill": int(tok._last_candidate_total), "kept": int(len(dd)),
                    "refill_ms": 1e3 * refill}
    return out


# ------------------------------------------------------------------------------------------------
# c3: all-pairs Lorentz distance + top-k over V=100k, row-sharded, lists all-gathered inside the timed region
# ------------------------------------------------------------------------------------------------
def bench_c3(a, env: Env, tf32: dict | None) -> dict:
    import ctypes as C
    from hyptokenizer_b200 import knn
    from hyptokenizer_b200._lib import SEM, check, ptr
    from hyptokenizer_b200.synth import synthetic_embeddings
    L, dev, rank, world, dist = env.L, env.dev, env.rank, env.world, env.dist
    V, d, k = C3_V, a.dim, C3_K
    D = d + 1
    sem = SEM[a.semantics]
    E = synthetic_embeddings(V, d, scale=SCALE, seed=42).to(dev)       # every rank holds all columns
    row0, nrows, per = knn.shard_rows(V, world, rank)
    stream = torch.cuda.current_stream()
    sp = stream.cuda_stream
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx = None
    if a.engine == "tc" and a.exchange in ("auto", "p2p"):
        ctx = knn.topk_context(V, k, dev)                              # None: peers cannot map each other's memory
        if ctx is None and a.exchange == "p2p":
            raise RuntimeError("--exchange p2p: hyp_ctx could not be connected")
    # local-only / NCCL variants: the shard's lists as interleaved records in a single-rank context's buffer
    idx = torch.empty((per, k), dtype=torch.int32, device=dev)
    dd = torch.empty((per, k), dtype=torch.float32, device=dev)
    flags = torch.zeros(max(per, 1), dtype=torch.int32, device=dev)
    nbytes = L.hyp_gram_topk_workspace_bytes(V, max(nrows, 1), D)
    ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=dev)
    wsp = ws.data_ptr() + ((-ws.data_ptr()) % 256)
    rec_local = torch.empty((per, k), dtype=torch.int64, device=dev)
    rec_all = torch.empty((world * per, k), dtype=torch.int64, device=dev)
    result = {}

    def local_lists():
        if a.engine == "tc":
            check(L.hyp_gram_topk(ptr(E), D, V, row0, nrows, D, 1.0, sem, k, ptr(idx), ptr(dd), ptr(flags), wsp, nbytes, sp))
        else:
            check(L.hyp_allpairs_topk(ptr(E), D, V, row0, nrows, D, 1.0, sem, k, ptr(idx), ptr(dd), sp))

    def step_fused():
        if ctx is not None:
            result["rec"] = ctx.gram_topk(E, 1.0, a.semantics)
        else:
            step_nccl()

    def step_local():
        local_lists()

    def step_nccl():
        local_lists()
        rec_local.copy_(knn.join_records(idx, dd))
        if world > 1:
            dist.all_gather_into_tensor(rec_all, rec_local)
            result["rec"] = rec_all
        else:
            result["rec"] = rec_local

    n_steps = max(a.steps, a.c3_min_steps)   # a step is a few ms: enough of them for the clock sampler to see the region

    def timed(fn, with_clocks=False):
        sampler = ClockSampler(env.local) if (rank == 0 and with_clocks) else None
        for _ in range(a.warmup):
            fn()
        env.barrier()
        if sampler:
            sampler.mark()
        ts = []
        for _ in range(n_steps):
            env.flush.fill_(1)
            e0.record(stream)
            fn()
            e1.record(stream)
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        env.barrier()
        clocks = sampler.stop() if sampler else None
        return env.reduce(sum(ts), "max") / n_steps, clocks

    ms_fused, clocks = timed(step_fused, with_clocks=True)
    rec = result["rec"]
    gi, gd = knn.split_records(rec[:V])
    gi, gd = gi.contiguous(), gd.contiguous()
    ms_local, _ = timed(step_local)
    ms_nccl = None
    if world > 1 and ctx is not None:
        ms_nccl, _ = timed(step_nccl)
    flagged = int(flags[:nrows].sum().item())
    status = ctx.status() if ctx is not None else 0

    # correctness and recall on rank 0 (north star): the gathered lists against exact brute force on two 512-row blocks
    # (one of the first shard, one of the last), and the reference's own candidate generator -- exact L2 kNN over Klein
    # coordinates, which its FAISS-HNSW index approximates -- against the exact Lorentz lists
    recall = None
    if rank == 0:
        rs = 512
        blocks = [0, max(0, V - rs)] if V > rs else [0]
        hits_exact, hits_klein, total, bit_equal = 0.0, 0.0, 0, True
        for b0 in blocks:
            ex_i = torch.empty((rs, k), dtype=torch.int32, device=dev)
            ex_d = torch.empty((rs, k), dtype=torch.float32, device=dev)
            check(L.hyp_allpairs_topk(ptr(E), D, V, b0, rs, D, 1.0, sem, k, ptr(ex_i), ptr(ex_d), sp))
            torch.cuda.synchronize()
            got_i, got_d = gi[b0:b0 + rs], gd[b0:b0 + rs]
            bit_equal &= bool(torch.equal(got_i, ex_i)) and bool(torch.equal(got_d.view(torch.int32), ex_d.view(torch.int32)))
            hits_exact += knn.recall_at_k(got_i, ex_i) * rs
            rows = torch.arange(b0, b0 + rs, device=dev)
            hits_klein += knn.recall_at_k(knn.klein_l2_topk(E, k, rows), ex_i) * rs
            total += rs
        recall = {"rows": total, "k": k, "timed_engine_vs_exact_bruteforce": hits_exact / total,
                  "bit_identical_to_exact_kernel": bit_equal,
                  "reference_klein_l2_knn_vs_exact_lorentz": hits_klein / total,
                  "note": "FAISS is not installable in this image: the second figure is for EXACT L2 kNN over Klein "
                          "coordinates xs/(x0+1e-8), the list a perfect HNSW search of the reference's index returns "
                          "(fast_hyperbolic_merge.py:195-240, :286-304); Klein-L2 order is not Lorentz-distance order"}
    flops = 2.0 * V * V * D                                      # one full distance matrix
    pk, kind = peaks()
    if tf32:
        tf32_peak, peak_kind = tf32["tf32_tflops"], "measured here: " + tf32["how"]
    else:
        tf32_peak, peak_kind = pk.get("bf16_tflops", 1590.0) / 2.0, kind + " bf16/2 (TF32 peak not measured in this run)"
    # hardware FLOPs of the tc engine: the collect pass visits every column tile, the bound pass every `step`-th
    # (same rule as hyp_gram_topk: 2, fewer while that leaves under 4k sampled tiles); K padded d+1 -> 104
    col_tiles = (V + 127) // 128
    tc_step = max(1, int(os.environ.get("HYP_TC_SUB", "2")))
    while tc_step > 1 and (col_tiles + tc_step - 1) // tc_step < 4 * k:
        tc_step -= 1
    hw = (1.0 + 1.0 / tc_step) if a.engine == "tc" else 1.0
    shard_flops = 2.0 * nrows * V * D
    kern_s = ms_local * 1e-3
    ach = hw * shard_flops / kern_s / 1e12
    launches = (6 if a.engine == "tc" else 1) + (1 if (ctx is not None and world > 1) else 0)
    line = {"metric": "all-pairs Lorentz dist TFLOP/s (V=100k,d=100,top-k=32)", "value": flops / (ms_fused * 1e-3) / 1e12,
            "unit": "TFLOP/s", "n_gpus": world, "steps": n_steps, "warmup": a.warmup, "ms_per_step": ms_fused,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "tf32+f32 rescore",
            "data": "synthetic",
            "config": {"workload": f"c3: all-pairs Lorentz distance + top-{k}, V={V}, d={d}, rows sharded over "
                                   f"{world} GPU(s), (V/G, k) lists all-gathered inside the timed region",
                       "engine": a.engine, "semantics": a.semantics,
                       "exchange": ("peer memory (hyp_ctx): finishing kernels store into every rank's buffer + barrier kernel"
                                    if ctx is not None and world > 1 else "NCCL all_gather_into_tensor of interleaved records"
                                    if world > 1 else "none (1 GPU)"),
                       "rows_redone_exactly": flagged, "barrier_status": status, "l2": "flushed between timed steps",
                       "ms_local_only": ms_local, "ms_with_nccl_allgather": ms_nccl,
                       # fused call minus the local kernels alone (which write two separate arrays instead of records:
                       # at 2 GPUs the difference is within the noise and is clamped at 0)
                       "allgather_ms": max(0.0, ms_fused - ms_local) if world > 1 else 0.0},
            "gpu_launches": launches * n_steps,
            "launches_note": "per step: tc_pack, gram_tc<1>, kth_select, gram_tc<2>, tc_finish, allpairs_topk (redo, "
                             "returns at once when no row is flagged)" + (", ctx_barrier" if ctx is not None and world > 1 else ""),
            "roofline": {"bound": "tensor", "kernel": "gram_tc_kernel<1> + <2> (+pack/select/finish)",
                         "achieved": ach, "peak": tf32_peak, "unit": "TFLOP/s", "frac": ach / tf32_peak,
                         "algorithmic_tflops_local": shard_flops / kern_s / 1e12,
                         "algorithmic_frac": shard_flops / kern_s / 1e12 / tf32_peak,
                         "traffic": None, "peak_kind": peak_kind,
                         "note": "hardware FLOPs: the tc engine runs the Gram GEMM over every column tile once (collect pass) "
                                 "and over every 2nd tile once more (bound pass); `value` counts the algorithmic "
                                 "2*V^2*(d+1) once; achieved is per GPU on its shard, local kernels only"},
            "clocks": clocks}
    if recall:
        line["recall"] = recall
    if tf32:
        line["tf32_peak"] = tf32
    return line


# ------------------------------------------------------------------------------------------------
# c4 (pair counting part): 1 GiB synthetic token stream per rank
# ------------------------------------------------------------------------------------------------
def bench_c4(a, env: Env) -> dict:
    from hyptokenizer_b200._lib import check, ptr
    from hyptokenizer_b200.pair_count import pairs_to_dict
    from hyptokenizer_b200.synth import synthetic_corpus
    L, dev, rank, world, dist = env.L, env.dev, env.rank, env.world, env.dist
    nbytes = a.c4_bytes                                  # per rank: every GPU counts its own stream (weak scaling)
    host = torch.from_numpy(synthetic_corpus(nbytes, seed=rank)).pin_memory()
    text = host.to(dev)
    cap = 1 << 20
    asc = torch.empty(128 * 128, dtype=torch.int64, device=dev)
    keys = torch.empty(cap, dtype=torch.int64, device=dev)
    vals = torch.empty(cap, dtype=torch.int64, device=dev)
    ovf = torch.empty(1, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    ts, tk = [], []
    sampler = ClockSampler(env.local) if rank == 0 else None
    for it in range(a.warmup + a.steps):
        if it == a.warmup:
            env.barrier()
            if sampler:
                sampler.mark()
        e0.record(stream)
        check(L.hyp_pair_count(ptr(text), nbytes, ptr(asc), ptr(keys), ptr(vals), cap, ptr(ovf), stream.cuda_stream))
        e1.record(stream)
        if world > 1:
            dist.all_reduce(asc, op=dist.ReduceOp.SUM)   # the shards' only exchange: the dense 128x128 histogram
        e2.record(stream)
        torch.cuda.synchronize()
        if it >= a.warmup:
            ts.append(e0.elapsed_time(e2))
            tk.append(e0.elapsed_time(e1))
    env.barrier()
    clocks = sampler.stop() if sampler else None
    ms = env.reduce(float(sum(ts)), "max") / a.steps
    if rank != 0:
        return {}
    # end to end: pinned host bytes -> H2D -> kernel -> D2H of the tables -> dict
    t0 = time.perf_counter()
    t2 = host.to(dev, non_blocking=True)
    check(L.hyp_pair_count(ptr(t2), nbytes, ptr(asc), ptr(keys), ptr(vals), cap, ptr(ovf), stream.cuda_stream))
    counts = pairs_to_dict(asc, keys, vals)
    e2e_s = time.perf_counter() - t0
    pk, kind = peaks()
    gbs = world * nbytes / (ms * 1e-3) / 1e9
    kgbs = nbytes / (float(np.mean(tk)) * 1e-3) / 1e9
    variant = os.environ.get("HYP_PAIR_COUNT", "default")
    line = {"metric": "pair counting GB/s (1 GB token stream)", "value": gbs, "unit": "GB/s", "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": f"c4 (pair-count part): {nbytes / 2 ** 30:g} GiB ASCII stream per GPU, lines of 20 random "
                                   "words; input larger than L2" + ("; histograms summed with one all_reduce per step"
                                                                    if world > 1 else ""),
                       "kernel_variant": variant,
                       "distinct_pairs": len(counts), "total_pairs": int(sum(counts.values()))},
            "gpu_launches": a.steps,
            "roofline": {"bound": "hbm", "kernel": "pair_count kernel (" + variant + ")", "achieved": kgbs,
                         "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": kgbs / pk["hbm_gbs"],
                         "traffic": None, "peak_kind": kind},
            "clocks": clocks,
            "e2e": {"value": nbytes / e2e_s / 1e9, "unit": "GB/s", "h2d_bytes_per_step": nbytes,
                    # what pairs_to_dict reads back: the dense table and the USED entries of the open-addressing table
                    "d2h_bytes_per_step": int(asc.numel() * 8 + 16 * int((keys != -1).sum().item()))}}
    if not a.no_cpu_baseline:
        from oracle.pair_count import count_pairs_c
        sample = host[: min(nbytes, 256 << 20)].numpy()
        t0 = time.perf_counter()
        ref = count_pairs_c(sample)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": sample.size / dt / 1e9, "unit": "GB/s", "cores": 1, "kind": "port",
                                "sample": "first 256 MiB of the same stream, C restatement (oracle/pair_count.c), one thread; "
                                          "the reference's Python dict loop measured 2.5 MB/s (BASELINE.md)"}
        # parity at the benchmark's own size (not timed): the dict of the WHOLE stream against the C restatement
        full = ref if sample.size == nbytes else count_pairs_c(host.numpy())
        line["config"]["dict_equals_oracle"] = bool(full == counts)
    # The same call on English letter frequencies (same size): rare letters and counters that wrap, which the uniform
    # stream of configs[3] never produces.  Reported beside the headline, dict checked against the oracle.
    from hyptokenizer_b200.synth import english_corpus
    eb = nbytes
    ehost = english_corpus(eb, seed=7)
    etext = torch.from_numpy(ehost).to(dev)
    tz = []
    for it in range(a.warmup + a.steps):
        e0.record(stream)
        check(L.hyp_pair_count(ptr(etext), eb, ptr(asc), ptr(keys), ptr(vals), cap, ptr(ovf), stream.cuda_stream))
        e1.record(stream)
        torch.cuda.synchronize()
        if it >= a.warmup:
            tz.append(e0.elapsed_time(e1))
    line["config"]["english_letters_gbs"] = eb / (float(np.mean(tz)) * 1e-3) / 1e9
    line["config"]["english_letters_bytes"] = eb
    if not a.no_cpu_baseline:
        line["config"]["english_letters_dict_equals_oracle"] = bool(count_pairs_c(ehost) == pairs_to_dict(asc, keys, vals))
    return line


# ------------------------------------------------------------------------------------------------
# CPU side: cost model of one reference step at n = V0 (port), and the reference itself where it can run
# ------------------------------------------------------------------------------------------------
def _cpu_sample(a, rows: int):
    """One bounded sample of one reference merge step at n = V0: the reference recomputes all n x n
    distances and extracts candidates every step (hyperbolic_merge.py:247-269); rows are independent,
    so `rows` query rows x all n columns are timed and scaled by n/rows."""
    from oracle import lorentz as OL
    from hyptokenizer_b200.synth import synthetic_embeddings
    v0, d = a.v0, a.dim
    emb = synthetic_embeddings(v0, d, scale=SCALE, seed=42)
    rows = min(rows, v0)
    t0 = time.perf_counter()
    dist = OL.batch_distance(emb[:rows], emb, 1.0, a.semantics)
    keep = (dist < THRESHOLD) & (torch.arange(v0)[None, :] > torch.arange(rows)[:, None])
    ii, jj = keep.nonzero(as_tuple=True)
    dd = dist[ii, jj]
    if len(dd):
        int(np.argmin(dd.numpy()))
    dt = time.perf_counter() - t0
    return dt * (v0 / rows), int(len(dd))


def cpu_baseline(a, rows: int = 1024):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    _cpu_sample(a, 64)
    t, ncand = _cpu_sample(a, rows)
    return {"value": 1.0 / t, "unit": "merges/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"COST MODEL, not a run: one brute-force merge step of the reference at n={a.v0} (all-pairs recompute + "
                      f"candidate extraction, vectorised), {rows} of {a.v0} query rows timed and scaled by n/rows; the "
                      f"reference itself cannot run this size (n^2*d*4 B = {a.v0 ** 2 * a.dim * 4 / 1e9:.0f} GB temporary); "
                      "`bench.py --impl reference` times the real reference at the largest size it runs"}


def _reference_modules():
    """The unmodified reference from the staged copy (or /root/reference in the build container), imported through
    oracle/gen_golden.py, which also holds the three-function geometry correction (`lorentz` semantics)."""
    for root in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if os.path.isdir(os.path.join(root, "tokenizer")) and os.path.isdir(os.path.join(root, "embedding")):
            os.environ["HYP_REFERENCE_ROOT"] = root
            import logging
            import warnings
            warnings.filterwarnings("ignore")
            logging.disable(logging.CRITICAL)
            from oracle import gen_golden as G
            return G, root
    return None, None


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    G, root = _reference_modules()
    base = {"impl": "reference", "metric": "merges/sec at V=50k,d=100", "unit": "merges/s", "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic"}
    if G is None:
        # no staged reference: the oracle port's cost model (never on the GPU box of a normal round: build() stages it)
        rows = 1024
        for _ in range(a.warmup):
            _cpu_sample(a, rows)
        ts = [_cpu_sample(a, rows)[0] for _ in range(a.steps)]
        value = len(ts) / sum(ts)
        base.update({"value": value, "ms_per_step": 1e3 * sum(ts) / len(ts),
                     "config": {"workload": f"c2 cost model: d={a.dim}, n={a.v0}", "semantics": a.semantics,
                                "note": "baseline/_ref missing; oracle port, one step at n=V0 scaled from 1024 query rows"},
                     "cpu_baseline": {"value": value, "unit": "merges/s", "cores": torch.get_num_threads(), "kind": "port",
                                      "sample": "cost model (see config.note)"},
                     "e2e": {"value": value, "unit": "merges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        print(json.dumps(base))
        return
    from hyptokenizer_b200.synth import synthetic_embeddings, synthetic_vocab
    v0, d = a.ref_v0, a.dim
    emb = synthetic_embeddings(v0, d, scale=SCALE, seed=42)
    vocab = synthetic_vocab(v0)

    def one_cycle(sem: str) -> float:
        """FastHyperbolicTokenizer.optimize_merges(steps=101) on a fresh tokenizer: one candidate search (all-pairs
        batch_distance + Python extraction + sort + cache) and the 100 merges its pop-100 cache serves."""
        with G.semantics(sem):
            G.set_seeds(42)
            tok = G.RF.FastHyperbolicTokenizer(list(vocab), torch.nn.Parameter(emb.clone()), curvature=1.0,
                                               merge_threshold=THRESHOLD, device=torch.device("cpu"),
                                               max_vocab_size=v0 + REF_CYCLE + 8, use_approximate_search=False)
            t0 = time.perf_counter()
            tok.optimize_merges(steps=REF_CYCLE, log_every=10 ** 9)
            dt = time.perf_counter() - t0
        assert len(tok.merge_history) == REF_CYCLE, "the reference stopped early"
        return dt

    for _ in range(a.warmup):
        one_cycle(a.semantics)
    ts = [one_cycle(a.semantics) for _ in range(a.steps)]
    value = REF_CYCLE * len(ts) / sum(ts)
    shipped = REF_CYCLE / one_cycle("reference") if a.semantics != "reference" else value
    sample = (f"each step = FastHyperbolicTokenizer.optimize_merges(steps={REF_CYCLE}) of the UNMODIFIED reference ({root}) "
              f"at V0={v0}, d={d}, torch CPU eager fp32, {torch.get_num_threads()} threads: measured wall time, nothing scaled")
    base.update({"value": value, "ms_per_step": 1e3 * sum(ts) / len(ts),
                 "config": {"workload": f"c2 at the largest size the reference runs: FastHyperbolicTokenizer d={d}, V0={v0}, "
                                        f"{REF_CYCLE} merges per step (one pop-100 cache cycle); V0=10 000 needs a "
                                        f"{10000 ** 2 * d * 4 / 1e9:.0f} GB broadcast temporary (lorentz_model.py:155-163)",
                            "semantics": a.semantics,
                            "semantics_note": "lorentz = the reference with the three-function geometry correction of SURVEY.md "
                                              "Appendix B monkey-patched (distance / batch_distance / log_map); loop, candidate "
                                              "extraction, cache and threshold logic are the shipped code",
                            "as_shipped_merges_per_s": shipped,
                            "note": "the shipped class merges from a stale pop-100 cache (one all-pairs search per 101 merges); "
                                    "our arm's default is the always-fresh exact search (the brute-force HyperbolicTokenizer "
                                    "sequence), which costs the reference one all-pairs search per merge"},
                 "cpu_baseline": {"value": value, "unit": "merges/s", "cores": torch.get_num_threads(), "kind": "reference",
                                  "sample": sample},
                 "e2e": {"value": value, "unit": "merges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(json.dumps(base))


# ------------------------------------------------------------------------------------------------
# batched tokenize: what scripts/benchmark_efficiency.py measures, tokens/s
# ------------------------------------------------------------------------------------------------
def run_tok(a):
    import random
    from hyptokenizer_b200 import _lib
    from hyptokenizer_b200._lib import check, ptr
    from hyptokenizer_b200.synth import synthetic_corpus, synthetic_embeddings
    from hyptokenizer_b200.tokenizer.batch_tokenize import RuleTable
    from hyptokenizer_b200.tokenizer.hyperbolic_merge import HyperbolicTokenizer
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    L = _lib.lib()
    chars = list("abcdefghijklmnopqrstuvwxyz ")
    vocab = ["<pad>", "<bos>", "<eos>", "<unk>"] + chars
    rng = random.Random(0)
    pool, history = list(chars[:-1]), []
    for _ in range(2000):                           # 2000 BPE-like rules over the synthetic alphabet
        x, y = rng.choice(pool), rng.choice(pool)
        if len(x + y) <= 6:
            history.append((x, y, x + y))
            pool.append(x + y)
    tok = HyperbolicTokenizer(vocab + [h[2] for h in history], torch.nn.Parameter(
        synthetic_embeddings(len(vocab) + len(history), 8)), max_vocab_size=len(vocab) + len(history) + 1, device=dev)
    tok.merge_history = history
    rules = {(x, y): z for x, y, z in history}
    table = RuleTable(rules, tok.vocab, tok.token2idx, dev)
    nbytes = 256 << 20
    host = synthetic_corpus(nbytes, seed=0)
    nl = np.flatnonzero(host == 10)
    offsets = np.concatenate([[0], nl + 1, [nbytes]]).astype(np.int64)       # newline stays at the end of each text
    d_text = torch.from_numpy(host).to(dev)
    d_off = torch.from_numpy(offsets).to(dev)
    n_texts = len(offsets) - 1
    d_tok = torch.empty(nbytes, dtype=torch.int32, device=dev)
    d_cnt = torch.zeros(n_texts, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for it in range(a.warmup + a.steps):
        e0.record(stream)
        check(L.hyp_apply_merges(ptr(d_text), ptr(d_off), n_texts, ptr(table.d_ascii), ptr(table.d_cp), ptr(table.d_cp_sym),
                                 table.n_cp, ptr(table.d_keys), ptr(table.d_vals), table.capacity, ptr(d_tok), ptr(d_cnt),
                                 stream.cuda_stream))
        e1.record(stream)
        torch.cuda.synchronize()
        if it >= a.warmup:
            ts.append(e0.elapsed_time(e1))
    n_tokens = int(d_cnt.sum().item())
    ms = float(np.mean(ts))
    line = {"metric": "tokenize tokens/s (batched apply-merges)", "value": n_tokens / (ms * 1e-3), "unit": "tokens/s",
            "n_gpus": 1, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8/int32", "data": "synthetic",
            "config": {"workload": f"tok: {nbytes >> 20} MiB ASCII, {n_texts} lines, {len(rules)} merge rules",
                       "input_GBps": nbytes / (ms * 1e-3) / 1e9, "tokens_out": n_tokens}, "gpu_launches": a.steps}
    if not a.no_cpu_baseline:
        from oracle.merge import OracleTokenizer
        ora = OracleTokenizer(tok.vocab, synthetic_embeddings(len(tok.vocab), 8), max_vocab_size=len(tok.vocab) + 1)
        ora.merge_history = history
        sample = host[: int(offsets[20000])].tobytes().decode("ascii").split("\n")[:20000]
        t0 = time.perf_counter()
        cnt = sum(len(ora.tokenize(ln + "\n")) for ln in sample)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": cnt / dt, "unit": "tokens/s", "cores": 1, "kind": "port",
                                "sample": "first 20 000 lines through the oracle's tokenize() (the reference's Python loop)"}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# crossover experiment: the per-merge step with ROW-SHARDED tables and an NCCL exchange per merge
# ------------------------------------------------------------------------------------------------
def run_c2_sharded(a):
    """Every rank keeps rows r = rank (mod G) of the table (as a dense local table), scores the new row against
    its shard with hyp_row_min, and the per-shard minima are all-gathered (one 16-byte record per rank) so that
    every rank derives the same winner -- the multi-GPU form of the merge loop the north star describes.  Timed
    to show where it stands against the single-GPU device-resident loop (it never wins while the table fits one
    GPU: the step is one NVLink/NCCL latency, not bandwidth)."""
    from hyptokenizer_b200._lib import SEM, check, ptr
    from hyptokenizer_b200.synth import synthetic_embeddings
    env = Env()
    L, dev, rank, world, dist = env.L, env.dev, env.rank, env.world, env.dist
    V, d = a.target, a.dim
    D = d + 1
    sem = SEM[a.semantics]
    full = synthetic_embeddings(V, d, scale=SCALE, seed=42).to(dev)
    shard = full[rank::world].contiguous()                 # this rank's rows, dense
    n_local = shard.shape[0]
    table = torch.zeros((n_local + a.steps * 200 + 8, D), device=dev)
    table[:n_local] = shard
    ws = torch.empty(L.hyp_merge_workspace_bytes(), dtype=torch.uint8, device=dev)
    best = torch.empty(32, dtype=torch.uint8, device=dev)
    gathered = torch.empty((world, 32), dtype=torch.uint8, device=dev)
    idx = torch.tensor([0, 1, 2, 2], dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream()
    sp = stream.cuda_stream
    merges = a.steps * 200
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def one_merge(n_loc):
        # new row = midpoint of two (replicated) rows, written behind this rank's shard as the query row
        check(L.hyp_midpoint(ptr(full), D, idx[0:].data_ptr(), idx[1:].data_ptr(), idx[2:].data_ptr(), idx[3:].data_ptr(),
                             table[n_loc].data_ptr(), D, 1, D, 1.0, sem, 1, sp))
        check(L.hyp_row_min(ptr(table), D, n_loc, n_loc, D, 1.0, sem, 0.1, ptr(best), ptr(ws), ws.numel(), sp))
        if world > 1:
            dist.all_gather_into_tensor(gathered.view(-1), best)
        # (the replicated argmin over `world` records would run on the device here; it is a few instructions)

    for _ in range(50):
        one_merge(n_local)
    env.barrier()
    e0.record(stream)
    for m in range(merges):
        one_merge(n_local + (m // world))
    e1.record(stream)
    torch.cuda.synchronize()
    ms = env.reduce(e0.elapsed_time(e1), "max")
    if rank == 0:
        us = ms * 1e3 / merges
        print(json.dumps({"metric": "merges/sec, row-sharded per-merge step with NCCL all-gather (crossover experiment)",
                          "value": 1e6 / us, "unit": "merges/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                          "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "strong",
                          "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": {"workload": f"c2-sharded: V={V}, d={d}, rows r = rank mod {world}, host-driven step: "
                                                 "midpoint + shard scan + all-gather of one 32-byte record per rank",
                                     "us_per_merge": us}, "gpu_launches": merges * 2}))
    env.close()


# ------------------------------------------------------------------------------------------------
# c5 (BASELINE configs[4]) at size: the enhanced tokenizer's step, host / device split
# ------------------------------------------------------------------------------------------------
def run_c5(a):
    from hyptokenizer_b200.bench_c5 import run as run5
    env = Env()
    line = run5(a, env)
    if env.rank == 0:
        print(json.dumps(line))
    env.close()


def run_ours(a):
    env = Env()
    if a.workload == "c3":
        tf32 = None if a.no_tf32_peak else measure_tf32_peak(env.dev)
        line = bench_c3(a, env, tf32)
    elif a.workload == "c4":
        line = bench_c4(a, env)
    else:
        line = bench_c2(a, env)
        if env.rank == 0 and not a.no_cpu_baseline and env.world == 1:
            line["cpu_baseline"] = cpu_baseline(a)
        if env.rank == 0 and a.workload in ("all", "c2") and not a.no_e2e:
            line["config"]["snapshot_cache_semantics"] = bench_snapshot(a, env)
        if a.workload == "all":
            # the other driver-visible figures of the path ride in the same line: numeric fields under config, full
            # lines under `secondary`
            tf32 = None if a.no_tf32_peak else measure_tf32_peak(env.dev)
            c3 = bench_c3(a, env, tf32)
            c4 = bench_c4(a, env)
            if env.rank == 0:
                cfg = line["config"]
                cfg.update({"c3_ms": c3["ms_per_step"], "c3_tflops": c3["value"],
                            "c3_frac": c3["roofline"]["frac"], "c3_algorithmic_frac": c3["roofline"]["algorithmic_frac"],
                            "c3_allgather_ms": c3["config"]["allgather_ms"], "c3_local_ms": c3["config"]["ms_local_only"],
                            "c3_nccl_ms": c3["config"]["ms_with_nccl_allgather"],
                            "c3_recall_exact": c3["recall"]["timed_engine_vs_exact_bruteforce"],
                            "c3_bit_identical": c3["recall"]["bit_identical_to_exact_kernel"],
                            "c3_recall_klein": c3["recall"]["reference_klein_l2_knn_vs_exact_lorentz"],
                            "c3_tf32_peak_tflops": c3["roofline"]["peak"],
                            "c4_gbs": c4["value"], "c4_kernel_gbs": c4["roofline"]["achieved"],
                            "c4_frac": c4["roofline"]["frac"],
                            "c4_dict_equals_oracle": c4["config"].get("dict_equals_oracle"),
                            "c4_english_letters_gbs": c4["config"].get("english_letters_gbs")})
                line["secondary"] = {"c3": c3, "c4": c4}
                line["gpu_launches"] += c3["gpu_launches"] + c4["gpu_launches"]
    if env.rank == 0:
        print(json.dumps(line))
    env.close()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "c2-sharded":
        run_c2_sharded(args)
    elif args.workload == "tok":
        run_tok(args)
    elif args.workload == "c5":
        run_c5(args)
    else:
        run_ours(args)
