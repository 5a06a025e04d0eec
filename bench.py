#!/usr/bin/env python
"""bench.py -- merges/sec of the HypTokenizer merge loop on B200 (BASELINE.json configs[1]).

Workload ("c2"): FastHyperbolicTokenizer, embedding_dim=100, V0=10 000 synthetic tokens grown to
target_vocab_size=50 000 (40 000 merges), SURVEY.md 8d.  One "step" = one whole job: initial
all-pairs argmin over V0 rows + 40 000 device-resident merges.

  value   merges/s with the table already in HBM, C-ABI calls only, CUDA events on the launch stream
  e2e     the same job through the public class (host embeddings in pinned memory -> H2D,
          optimize_merges, D2H of the merge log and of embeddings[:n], host string rebuild)
  roofline  the merge-loop kernel against measured HBM bandwidth: algorithmic bytes
          sum_n 4*n*(d+1) over the merges of one launch / its CUDA-event duration
  cpu_baseline  the oracle port (reference algorithm, torch CPU eager) on a bounded sample

`--impl reference` times the reference's own algorithm (oracle port: the reference is pure Python
and cannot travel to the GPU box) on the host cores.

N > 1 (torchrun): the merge loop is inherently sequential and its table fits one GPU's L2, so the
path does not shard ("replicas only"): every rank runs an independent job on its own
seed; value = total merges of all ranks / max-over-ranks time.
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
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c4", "tok", "c2-sharded"],
                    help="c2: merge loop (headline, merges/s); c3: all-pairs Lorentz distance + top-k=32 over V=100k, "
                         "row-sharded over the ranks with an all-gather (TFLOP/s)")
    ap.add_argument("--engine", default="tc", choices=["tc", "exact"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler:
    """nvidia-smi sampling DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(gpu_index)], stdout=self.tmp,
                                         stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.tmp.flush()
        self.tmp.seek(0)
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.tmp.read().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.tmp.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "reasons": sorted(reasons),
                "samples": len(sm)}


def algorithmic_bytes(v0: int, merges: int, d: int) -> float:
    """each merge reads the n current rows once: 4*n*(d+1) bytes (+ one row written)."""
    n_sum = merges * v0 + merges * (merges - 1) // 2
    return 4.0 * (d + 1) * (n_sum + merges)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(a):
    import torch.distributed as dist
    from hyptokenizer_b200 import _lib
    from hyptokenizer_b200._lib import SEM, check, ptr
    from hyptokenizer_b200.synth import synthetic_embeddings, synthetic_vocab
    from hyptokenizer_b200.tokenizer.fast_hyperbolic_merge import FastHyperbolicTokenizer

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()
    _lib.check_device(dev)

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
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    stream = torch.cuda.current_stream()
    sp = stream.cuda_stream
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    thr32 = float(np.float32(THRESHOLD))

    def one_step(timed):
        # reset the table to the initial vocabulary (not part of the path; outside the events)
        E[:v0].copy_(init)
        E[v0:].zero_()
        lens.copy_(lens_init)
        flush.fill_(1)                                   # L2 flush between timed iterations
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

    for _ in range(a.warmup):
        one_step(False)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    sampler = ClockSampler(local) if rank == 0 else None
    t_total, t_loop = [], []
    for _ in range(a.steps):
        tt, tl = one_step(True)
        t_total.append(tt)
        t_loop.append(tl)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
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
    from hyptokenizer_b200._lib import HypMergeState
    st = HypMergeState.from_buffer_copy(state.cpu().numpy().tobytes())
    done = st.steps_done
    ms_total = float(sum(t_total))
    tmax = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    done_all = torch.tensor([done * a.steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(done_all, op=dist.ReduceOp.SUM)
    value = float(done_all.item()) / (float(tmax.item()) * 1e-3)

    # ---- end to end through the public class, host buffers, copies inside the timed region ----------
    e2e = None
    if not a.no_e2e:
        vocab = synthetic_vocab(v0)
        out = torch.empty((target, D), dtype=torch.float32).pin_memory()    # landing buffer of the result table
        e2e_t = []
        for it in range(2):
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            tok = FastHyperbolicTokenizer(vocab, torch.nn.Parameter(host_emb), merge_threshold=THRESHOLD,
                                          max_vocab_size=target, device=dev, semantics=a.semantics)
            tok.optimize_merges(steps=merges, log_every=10 ** 9, adaptive_threshold=True)
            n_rows = tok.current_vocab_size
            out[:n_rows].copy_(tok.embeddings.detach()[:n_rows], non_blocking=True)
            torch.cuda.synchronize()
            e2e_t.append(time.perf_counter() - t0)
            n_e2e = len(tok.last_trace)
        te = torch.tensor([min(e2e_t)], dtype=torch.float64, device=dev)
        ne = torch.tensor([float(n_e2e)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
            dist.all_reduce(ne, op=dist.ReduceOp.SUM)
        e2e = {"value": float(ne.item()) / float(te.item()), "unit": "merges/s",
               "h2d_bytes_per_step": int(host_emb.numel() * 4 + target * 4 + 40),
               "d2h_bytes_per_step": int(n_e2e * 16 + n_rows * D * 4 + 40 * ((n_e2e + 8191) // 8192))}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk, pk_kind = peaks()
    loop_ms = float(np.mean(t_loop))
    abytes = algorithmic_bytes(v0, done, d)
    achieved = abytes / (loop_ms * 1e-3) / 1e9
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
        "roofline": {"bound": "hbm", "kernel": "merge_loop_resident_kernel<100, false>", "achieved": achieved, "peak": pk["hbm_gbs"],
                     "unit": "GB/s", "frac": achieved / pk["hbm_gbs"],
                     "traffic": 4107776, "traffic_source": "profiles/r01_prof_merge_raw.csv: dram__bytes_read.sum + "
                                                           "dram__bytes_write.sum of one launch (ncu --set full)",
                     "peak_kind": pk_kind,
                     "algorithmic_bytes_per_launch": abytes, "kernel_ms": loop_ms,
                     "note": "the table (<= 20.2 MB) lives in the shared memory of the persistent grid, so DRAM traffic is far "
                             "below the algorithmic bytes; frac is the algorithmic rate against HBM copy bandwidth"},
        "clocks": clocks,
    }
    if e2e:
        line["e2e"] = e2e
    if not a.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline(a)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle port of the reference algorithm on the host cores
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
            "sample": f"one brute-force merge step of the reference at n={a.v0} (all-pairs recompute + candidate "
                      f"extraction, vectorised), {rows} of {a.v0} query rows timed and scaled by n/rows; the "
                      f"reference itself cannot run this size (n^2*d*4 B = {a.v0 ** 2 * a.dim * 4 / 1e9:.0f} GB temporary)"}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    rows = 1024
    for _ in range(a.warmup):
        _cpu_sample(a, rows)
    ts = []
    for _ in range(a.steps):
        t, _ = _cpu_sample(a, rows)
        ts.append(t)
    value = len(ts) / sum(ts)
    sample = (f"each step = one brute-force merge step of the reference algorithm at n={a.v0}, d={a.dim} "
              f"({rows} of {a.v0} query rows timed, scaled by n/rows); oracle port, torch CPU eager fp32")
    line = {"impl": "reference", "metric": "merges/sec at V=50k,d=100", "value": value, "unit": "merges/s",
            "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * sum(ts) / len(ts),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"c2: d={a.dim}, V0={a.v0} -> {a.target}", "semantics": a.semantics,
                       "note": "per-step cost grows as n^2; n=V0 is the CHEAPEST step of the job, so this flatters the reference"},
            "cpu_baseline": {"value": value, "unit": "merges/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": sample},
            "e2e": {"value": value, "unit": "merges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# config 3: all-pairs Lorentz distance + top-k over V=100k, row-sharded, all-gather of the lists
# ------------------------------------------------------------------------------------------------
def run_c3(a):
    import torch.distributed as dist
    from hyptokenizer_b200 import _lib, knn
    from hyptokenizer_b200._lib import SEM, check, ptr
    from hyptokenizer_b200.synth import synthetic_embeddings

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()
    _lib.check_device(dev)
    V, d, k = 100000, a.dim, 32
    D = d + 1
    sem = SEM[a.semantics]
    E = synthetic_embeddings(V, d, scale=SCALE, seed=42).to(dev)       # every rank holds all columns
    row0, nrows, per = knn.shard_rows(V, world, rank)
    idx = torch.empty((per, k), dtype=torch.int32, device=dev)
    dd = torch.empty((per, k), dtype=torch.float32, device=dev)
    all_i = torch.empty((world * per, k), dtype=torch.int32, device=dev)
    all_d = torch.empty((world * per, k), dtype=torch.float32, device=dev)
    flags = torch.zeros(per, dtype=torch.int32, device=dev)
    nbytes = L.hyp_gram_topk_workspace_bytes(V, nrows, D)
    ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=dev)
    wsp = ws.data_ptr() + ((-ws.data_ptr()) % 256)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()
    sp = stream.cuda_stream
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))

    def step():
        flush.fill_(1)
        e0.record(stream)
        if a.engine == "tc":
            check(L.hyp_gram_topk(ptr(E), D, V, row0, nrows, D, 1.0, sem, k, ptr(idx), ptr(dd), ptr(flags), wsp, nbytes, sp))
        else:
            check(L.hyp_allpairs_topk(ptr(E), D, V, row0, nrows, D, 1.0, sem, k, ptr(idx), ptr(dd), sp))
        e1.record(stream)
        if world > 1:
            dist.all_gather_into_tensor(all_i, idx)
            dist.all_gather_into_tensor(all_d, dd)
        e2.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e2), e0.elapsed_time(e1)

    for _ in range(a.warmup):
        step()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    tot, ker = [], []
    for _ in range(a.steps):
        t, t1 = step()
        tot.append(t)
        ker.append(t1)
    if world > 1:
        dist.barrier()
    clocks = sampler.stop() if sampler else None
    flagged = int(flags[:nrows].sum().item())
    tmax = torch.tensor([sum(tot)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    if rank == 0:
        flops = 2.0 * V * V * D                                      # one full distance matrix
        secs = float(tmax.item()) * 1e-3 / a.steps
        pk, kind = peaks()
        kern_s = float(np.mean(ker)) * 1e-3
        shard_flops = 2.0 * nrows * V * D
        tf32_peak = pk.get("bf16_tflops", 1590.0) / 2.0               # TF32 dense = half the bf16 rate
        # hardware FLOPs of the tc engine: the collect pass visits every column tile, the bound pass every
        # `step`-th (same rule as hyp_gram_topk: 2, fewer while that leaves under 4k sampled tiles)
        col_tiles = (V + 127) // 128
        tc_step = max(1, int(os.environ.get("HYP_TC_SUB", "2")))
        while tc_step > 1 and (col_tiles + tc_step - 1) // tc_step < 4 * k:
            tc_step -= 1
        hw = 1.0 + 1.0 / tc_step
        line = {"metric": "all-pairs Lorentz dist TFLOP/s (V=100k,d=100,top-k=32)", "value": flops / secs / 1e12,
                "unit": "TFLOP/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": secs * 1e3,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "tf32+f32 rescore",
                "data": "synthetic",
                "config": {"workload": f"c3: all-pairs Lorentz distance + top-{k}, V={V}, d={d}, rows sharded over "
                                       f"{world} GPU(s), all-gather of the (V/G, k) lists", "engine": a.engine,
                           "semantics": a.semantics, "rows_flagged_for_exact_redo": flagged,
                           "l2": "flushed between timed steps"},
                "gpu_launches": (5 if a.engine == "tc" else 1) * a.steps,
                "roofline": {"bound": "tensor", "kernel": "gram_tc_kernel x2 (+pack/select/finish)",
                             "achieved": hw * shard_flops / kern_s / 1e12 if a.engine == "tc" else shard_flops / kern_s / 1e12,
                             "peak": tf32_peak, "unit": "TFLOP/s",
                             "frac": (hw * shard_flops / kern_s / 1e12 if a.engine == "tc" else shard_flops / kern_s / 1e12) / tf32_peak,
                             "traffic": None, "peak_kind": kind + " bf16/2",
                             "note": "hardware FLOPs: the tc engine runs the Gram GEMM over every column tile once (collect pass) "
                                     "and over every 2nd tile once more (bound pass); "
                                     "`value` counts the algorithmic 2*V^2*(d+1) once"},
                "clocks": clocks}
        # recall (north star): the timed engine against exact brute force, and the reference's own candidate generator --
        # exact L2 kNN over Klein coordinates, which its FAISS-HNSW index approximates -- against the exact Lorentz lists
        rs = min(1024, nrows)
        g = torch.Generator().manual_seed(0)
        rows = (row0 + torch.randperm(nrows, generator=g)[:rs]).sort().values.to(dev)
        got = idx[(rows - row0)].to(torch.int64)
        ex_i = torch.empty((rs, k), dtype=torch.int32, device=dev)
        ex_d = torch.empty((rs, k), dtype=torch.float32, device=dev)
        hits = 0
        for j, r in enumerate(rows.tolist()):                         # the exact CUDA-core engine, row by row
            check(L.hyp_allpairs_topk(ptr(E), D, V, r, 1, D, 1.0, sem, k, ex_i[j].data_ptr(), ex_d[j].data_ptr(), sp))
        torch.cuda.synchronize()
        line["recall"] = {"rows": rs, "k": k,
                          "timed_engine_vs_exact_bruteforce": knn.recall_at_k(got, ex_i),
                          "reference_klein_l2_knn_vs_exact_lorentz": knn.recall_at_k(knn.klein_l2_topk(E, k, rows), ex_i),
                          "note": "FAISS is not installable in this image: the second figure is for EXACT L2 kNN over Klein "
                                  "coordinates xs/(x0+1e-8), the list a perfect HNSW search of the reference's index returns "
                                  "(fast_hyperbolic_merge.py:195-240, :286-304); Klein-L2 order is not Lorentz-distance order"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# config 4 (pair counting part): 1 GB synthetic token stream
# ------------------------------------------------------------------------------------------------
def run_c4(a):
    import torch.distributed as dist
    from hyptokenizer_b200 import _lib
    from hyptokenizer_b200._lib import check, ptr
    from hyptokenizer_b200.pair_count import pairs_to_dict
    from hyptokenizer_b200.synth import synthetic_corpus
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()
    _lib.check_device(dev)
    nbytes = 1 << 30                                   # per rank: every GPU counts its own GiB (weak scaling)
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
    sampler = None
    for it in range(a.warmup + a.steps):
        if it == a.warmup:
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
                torch.cuda.synchronize()
            sampler = ClockSampler(local) if rank == 0 else None
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
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None
    tot = torch.tensor([float(sum(ts))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    if rank != 0:
        dist.destroy_process_group()
        return
    # end to end: pinned host bytes -> H2D -> kernel -> D2H of the tables -> dict
    t0 = time.perf_counter()
    t2 = host.to(dev, non_blocking=True)
    check(L.hyp_pair_count(ptr(t2), nbytes, ptr(asc), ptr(keys), ptr(vals), cap, ptr(ovf), stream.cuda_stream))
    counts = pairs_to_dict(asc, keys, vals)
    e2e_s = time.perf_counter() - t0
    pk, kind = peaks()
    ms = float(tot.item()) / a.steps
    gbs = world * nbytes / (ms * 1e-3) / 1e9
    kgbs = nbytes / (float(np.mean(tk)) * 1e-3) / 1e9
    line = {"metric": "pair counting GB/s (1 GB token stream)", "value": gbs, "unit": "GB/s", "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "c4 (pair-count part): 1 GiB ASCII stream per GPU, lines of 20 random words; input "
                                   "larger than L2" + ("; histograms summed with one all_reduce per step" if world > 1 else ""),
                       "distinct_pairs": len(counts), "total_pairs": int(sum(counts.values()))},
            "gpu_launches": a.steps, "roofline": {"bound": "hbm", "kernel": "pair_count_v2_kernel" if os.environ.get("HYP_PAIR_COUNT", "v2") != "v1" else "pair_count_kernel", "achieved": kgbs,
                                                  "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": kgbs / pk["hbm_gbs"],
                                                  "traffic": 1078586856 if os.environ.get("HYP_PAIR_COUNT", "v2") != "v1" else None,
                                                  "traffic_source": "profiles/r01_prof_pair_v2_raw.csv: dram__bytes_read.sum + "
                                                                    "dram__bytes_write.sum of one launch over 1 GiB (ncu --set full)",
                                                  "peak_kind": kind},
            "clocks": clocks,
            "e2e": {"value": nbytes / e2e_s / 1e9, "unit": "GB/s", "h2d_bytes_per_step": nbytes,
                    # what pairs_to_dict reads back: the dense table and the USED entries of the open-addressing table
                    "d2h_bytes_per_step": int(asc.numel() * 8 + 16 * int((keys != -1).sum().item()))}}
    if not a.no_cpu_baseline:
        from oracle.pair_count import count_pairs_c
        sample = host[: 256 << 20].numpy()
        t0 = time.perf_counter()
        ref = count_pairs_c(sample)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": sample.size / dt / 1e9, "unit": "GB/s", "cores": 1, "kind": "port",
                                "sample": "first 256 MiB of the same stream, C restatement (oracle/pair_count.c), one thread; "
                                          "the reference's Python dict loop measured 2.5 MB/s (BASELINE.md)"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


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
    import torch.distributed as dist
    from hyptokenizer_b200 import _lib
    from hyptokenizer_b200._lib import SEM, check, ptr
    from hyptokenizer_b200.synth import synthetic_embeddings
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()
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
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0.record(stream)
    for m in range(merges):
        one_merge(n_local + (m // world))
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        us = float(t.item()) * 1e3 / merges
        print(json.dumps({"metric": "merges/sec, row-sharded per-merge step with NCCL all-gather (crossover experiment)",
                          "value": 1e6 / us, "unit": "merges/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                          "ms_per_step": float(t.item()) / a.steps, "higher_is_better": True, "scaling": "strong",
                          "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": {"workload": f"c2-sharded: V={V}, d={d}, rows r = rank mod {world}, host-driven step: "
                                                 "midpoint + shard scan + all-gather of one 32-byte record per rank",
                                     "us_per_merge": us}, "gpu_launches": merges * 2}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "c2-sharded":
        run_c2_sharded(args)
    elif args.workload == "tok":
        run_tok(args)
    elif args.workload == "c4":
        run_c4(args)
    elif args.workload == "c3":
        run_c3(args)
    else:
        run_ours(args)
