#!/usr/bin/env python
"""Headline benchmark: exact top-100 queries/s over a 1,007,000 x 2048 database (BASELINE.json
configs[1], the rParis6k+R1M shape), 70-query batches.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

ours       N=1: the whole database on one B200.  N>1: the SAME database row-sharded over the N
           GPUs (strong scaling: total work fixed), one NCCL all-gather of the per-shard top-100
           lists + a merge kernel per step.  A step = one 70-query batch through the hot path.
           `value`  = queries/s with queries already in HBM (CUDA events, max over ranks)
           `e2e`    = queries/s through the host-buffer C-ABI call (pinned host queries in, ids and
                      scores back to the host, every step)
           `roofline` = bf16 database bytes / duration of the dominant kernel vs measured HBM peak
reference  the reference's own CPU path for this metric -- np.dot(vecs.T, qvecs) +
           np.argsort(-scores, axis=0)[:100] (src/main_retrieve.py:175-176), restated in
           oracle/oracle.py -- on the host cores, on a bounded row sample of the same workload.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = "image-search-engine-for-historical-research_b200"

N_ROWS, DIM, N_QUERIES, TOPK = 1_007_000, 2048, 70, 100
METRIC, UNIT = "exact top-100 queries/s, 1M x 2048-d DB", "queries/s"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smmax, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smmax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [x for x in sm if x >= 0.5 * max(sm)] if sm else []
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(smmax) if smmax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def synth_rows_device(torch, n, d, device, seed):
    """Unit-norm Gaussian rows generated on the GPU in chunks (the host generator of synth.py would
    spend a minute on 2 G floats); same distribution family 'G' as the parity tests."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = torch.empty((n, d), dtype=torch.float32, device=device)
    for lo in range(0, n, 65536):
        hi = min(n, lo + 65536)
        blk = torch.randn((hi - lo, d), generator=g, dtype=torch.float32, device=device)
        out[lo:hi] = blk / blk.norm(dim=1, keepdim=True)
    return out


def reference_layout(db_rows: np.ndarray, queries: np.ndarray):
    """Row-major sample [s, D] / queries [Q, D] -> the reference's (D, s) and (D, Q) arrays."""
    return np.ascontiguousarray(db_rows.T), np.ascontiguousarray(queries.T)


def cpu_reference_step(vecs: np.ndarray, qvecs: np.ndarray) -> float:
    """One timed pass of np.dot + argsort[:100] exactly as the oracle restates
    main_retrieve.py:175-176; returns seconds."""
    oracle = importlib.import_module("oracle.oracle")
    t0 = time.time()
    _, ranks = oracle.rank_ip(vecs, qvecs)
    top = ranks[:TOPK]
    dt = time.time() - t0
    assert top.shape == (TOPK, qvecs.shape[1])
    return dt


def cpu_extras(vecs: np.ndarray, qvecs: np.ndarray, scale: float) -> dict:
    """The two other CPU timings SURVEY.md 8(d) asks for next to np.dot + argsort, on the same sample:
    the restated matching_L2 (nnsearch.py:687-706, one query -- it is seconds per query) and a best-effort exact
    top-K the reference does not have (torch.mm + topk on all host threads).  ``scale`` = full rows / sample rows."""
    import torch
    oracle = importlib.import_module("oracle.oracle")
    out = {}
    t0 = time.time()
    oracle.matching_L2(TOPK, vecs.T, qvecs.T[:1])
    out["matching_L2_s_per_query_full_db"] = (time.time() - t0) * scale
    v, q = torch.from_numpy(np.ascontiguousarray(vecs.T)), torch.from_numpy(np.ascontiguousarray(qvecs))
    best = 1e9
    for _ in range(2):
        t0 = time.time()
        torch.topk(torch.mm(v, q), TOPK, dim=0)
        best = min(best, time.time() - t0)
    out["torch_mm_topk_qps_full_db"] = qvecs.shape[1] / (best * scale)
    return out


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ----------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample_rows = int(os.environ.get("XS_BENCH_REF_ROWS", "125875"))     # 1/8 of the database per step (~1 GB fp32)
    rng = np.random.default_rng(0)
    db = rng.standard_normal((sample_rows, DIM), dtype=np.float32)
    db /= np.linalg.norm(db, axis=1, keepdims=True)
    q = np.random.default_rng(1).standard_normal((N_QUERIES, DIM), dtype=np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    vecs, qvecs = reference_layout(db, q)
    del db
    # keep the whole run within ~2 minutes whatever K is: shrink the row sample if one step is too slow
    t1 = cpu_reference_step(vecs, qvecs)
    budget = 90.0
    if args.steps * t1 > budget:
        sample_rows = max(8192, int(sample_rows * budget / (args.steps * t1)))
        vecs = np.ascontiguousarray(vecs[:, :sample_rows])
    for _ in range(max(1, min(args.warmup, 2))):
        cpu_reference_step(vecs, qvecs)
    dt = sum(cpu_reference_step(vecs, qvecs) for _ in range(args.steps)) / args.steps
    qps = N_QUERIES / (dt * N_ROWS / sample_rows)
    cores = host_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": qps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * N_ROWS / sample_rows * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg2: 1,007,000 x 2048 fp32 DB, 70-query batch, exact top-100",
                   "path": "np.dot(vecs.T, qvecs) + np.argsort(-scores, axis=0)[:100] (src/main_retrieve.py:175-176 as restated in oracle/oracle.py)"},
        "cpu_baseline": {"value": qps, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"each step = all 70 queries x {sample_rows} rows of the 1,007,000 (time scaled x{N_ROWS / sample_rows:.1f} to the full DB); numpy/OpenBLAS threads = host default"},
        "e2e": {"value": qps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ----------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    pkg = importlib.import_module(PKG)
    sharded = importlib.import_module(PKG + ".sharded")

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the matching path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---- database shard + queries (synthetic, unit-norm Gaussian) -----------------------------------
    replicas = args.sharding == "replicas" and world > 1     # every GPU holds the whole database and answers its own batches
    bounds = sharded.shard_bounds(N_ROWS, world)
    lo, hi = (0, N_ROWS) if replicas else (bounds[rank], bounds[rank + 1])
    full = synth_rows_device(torch, N_ROWS, DIM, dev, seed=0) if (world == 1 or replicas) else None
    if world == 1 or replicas:
        rows = full
    else:
        # every rank draws the same global matrix in the same chunks and keeps its slice
        g = torch.Generator(device=dev); g.manual_seed(0)
        rows = torch.empty((hi - lo, DIM), dtype=torch.float32, device=dev)
        for c0 in range(0, N_ROWS, 65536):
            c1 = min(N_ROWS, c0 + 65536)
            blk = torch.randn((c1 - c0, DIM), generator=g, dtype=torch.float32, device=dev)
            a, b = max(c0, lo), min(c1, hi)
            if a < b:
                rows[a - lo:b - lo] = blk[a - c0:b - c0] / blk[a - c0:b - c0].norm(dim=1, keepdim=True)
    queries = synth_rows_device(torch, N_QUERIES, DIM, dev, seed=1)
    torch.cuda.synchronize()
    index = pkg.ExactIndex.from_device(rows.data_ptr(), hi - lo, DIM, local, renormalise=False, id_offset=lo)
    shard = sharded.CudaShard(index, local, lanes=args.lanes)
    exchange = None
    if world > 1 and not replicas and args.exchange in ("auto", "peer"):
        try:                                                   # collective set-up: succeeds or fails on every rank together
            exchange = sharded.PeerExchange(local, sharded.packed_bytes(N_QUERIES, TOPK))
        except RuntimeError as e:
            print(f"rank {rank}: {e}; falling back to the NCCL all-gather", file=sys.stderr)
    searcher = sharded.ShardedSearcher(shard.local_search, shard.merge, exchange=exchange, exchange_pipelined=args.exchange == "peer",
                                       lane_stream=shard.lane_stream if args.lanes > 1 else None)
    if replicas:
        searcher.world = 1                                     # no exchange step at all

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput (`value`) --------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        ids, sims = searcher.search(queries, TOPK)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    pending = None                                   # two batches in flight: batch i's all-gather overlaps batch i+1's scan
    for _ in range(args.steps):
        nxt = searcher.search_async(queries, TOPK)
        if pending is not None:
            ids, sims = pending.result()
        pending = nxt
    ids, sims = pending.result()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    uncert = shard.uncertified(N_QUERIES, TOPK)
    launches_per_step = index.stats()["gpu_launches"] + ((2 if (exchange is not None and args.exchange == "peer") else 1) if (world > 1 and not replicas) else 0)

    # ---- end to end through the host-buffer call (`e2e`) ----------------------------------------------
    q_host = queries.cpu().pin_memory()
    q_np = q_host.numpy()
    if world > 1:       # pinned landing buffers for the sharded path (the host-buffer C-ABI call has its own)
        qd = torch.empty_like(queries)
        ids_pin = torch.empty((N_QUERIES, TOPK), dtype=torch.int64).pin_memory()
        sims_pin = torch.empty((N_QUERIES, TOPK), dtype=torch.float32).pin_memory()

    def e2e_step():
        if world == 1 or replicas:
            return index.search(q_np, TOPK)                       # pinned host queries in, ids + scores back on the host
        qd.copy_(q_host, non_blocking=True)
        i_d, s_d = searcher.search(qd, TOPK)
        ids_pin.copy_(i_d, non_blocking=True)
        sims_pin.copy_(s_d, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return ids_pin, sims_pin

    for _ in range(3):
        ids_h, sims_h = e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ids_h, sims_h = e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None

    # ---- dominant kernel, timed on its own stream by the library's CUDA events --------------------------
    coarse = []
    index.set_param("timing", 1)
    for _ in range(min(args.steps, 20)):
        shard.local_search(queries, TOPK)
        coarse.append(index.stats()["ms_coarse"])
    coarse_ms = sum(coarse) / len(coarse)
    stats = index.stats()
    index.set_param("timing", 0)

    # max over ranks
    t = torch.tensor([ms, e2e_s * 1e3, coarse_ms, float(uncert)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms, coarse_ms, uncert = [float(x) for x in t.tolist()]

    # ---- parity spot check of the timed configuration against the exact fp32 path ---------------------
    index.set_param("force_path", 3)
    ids_x, sims_x = index.search(q_np[:4], TOPK)
    index.set_param("force_path", 0)
    ids_l, sims_l = index.search(q_np[:4], TOPK)
    parity_ok = bool((ids_x == ids_l).all() and np.allclose(sims_x, sims_l, rtol=1e-6))

    if exchange is not None:
        exchange.close()                                       # collective: drains, barriers, unmaps
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- CPU baseline on the host cores (rank 0, N=1 only) -------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        sample_rows = 251_750
        vecs, qvecs = reference_layout(full[:sample_rows].cpu().numpy(), q_np)
        dt = min(cpu_reference_step(vecs, qvecs) for _ in range(2))
        qps_cpu = N_QUERIES / (dt * N_ROWS / sample_rows)
        cpu = {"value": qps_cpu, "unit": UNIT, "cores": host_threads(), "kind": "port",
               "sample": f"np.dot + argsort[:100] (oracle.rank_ip), all 70 queries x {sample_rows} rows (1/4 of the DB, {dt:.2f} s), scaled x4 to the full DB"}
        try:
            cpu["other_cpu_paths"] = cpu_extras(vecs, qvecs, N_ROWS / sample_rows)
        except Exception as e:                                   # informational only
            cpu["other_cpu_paths"] = {"error": str(e)[:200]}

    peak, peak_src = measured_peaks()
    shard_rows = hi - lo
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic_r1.json")
    if os.path.exists(tpath):
        t = json.load(open(tpath)).get("gemm_topk_kernel", {})
        if t.get("rows") == shard_rows:
            traffic = t["bytes"] / 1e9                 # GB per launch, from the committed ncu --set full capture
    algo_bytes = shard_rows * DIM * 2                    # one pass over the bf16 shard
    achieved = algo_bytes / (coarse_ms * 1e-3) / 1e9
    batches = world if replicas else 1                      # 70-query batches answered per step by the whole job
    qps = batches * N_QUERIES * args.steps / (ms * 1e-3)
    e2e_qps = batches * N_QUERIES * args.steps / (e2e_ms * 1e-3)
    line = {
        "metric": METRIC, "value": qps, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak" if replicas else "strong", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "cfg2: 1,007,000 x 2048 DB (unit-norm Gaussian, seed 0), 70-query batch, exact top-100",
                   "rows_per_gpu": shard_rows, "lanes": args.lanes, "sharding": "none" if world == 1 else (f"{world} replicas of the whole database, one 70-query batch per GPU and step, no collective" if replicas else (f"row-sharded x{world}; e2e call: per-shard top-100 pushed into every rank's mailbox over NVLink peer memory, merge kernel waits on arrival flags; value loop: " + ("the same push on a side stream" if args.exchange == "peer" else "NCCL all-gather on its own stream + merge kernel") if exchange is not None else f"row-sharded x{world}, NCCL all-gather of per-shard top-100 + merge kernel")),
                   "l2": "inputs larger than L2 (4.1 GB bf16 database per pass vs 126 MB L2)",
                   "arithmetic": "bf16 operands / fp32 accumulate (tcgen05) for the coarse pass, then fp32 operands / fp64 accumulate exact rescoring of ~120 candidates per query",
                   "path": {1: "scan", 2: "tcgen05 GEMM + fused top-K", 3: "exact"}.get(stats["path"], "?"),
                   "uncertified_queries_last_step": int(uncert), "parity_spot_check": parity_ok},
        "e2e": {"value": e2e_qps, "unit": UNIT, "h2d_bytes_per_step": N_QUERIES * DIM * 4,
                "d2h_bytes_per_step": N_QUERIES * TOPK * 12 + N_QUERIES * 4 + 4, "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": int(launches_per_step * args.steps),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_unit": "GB per launch (ncu dram read+write, profiles/)", "kernel": "gemm_topk_kernel", "kernel_ms": coarse_ms, "algorithmic_bytes": algo_bytes,
                     "peak_source": peak_src},
        "clocks": clocks,
    }
    if cpu:
        line["cpu_baseline"] = cpu
    emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line: dict):
    """The ONE JSON line goes to the real stdout; everything else any library prints (NCCL's version
    banner lands on fd 1) has been routed to stderr by main()."""
    data = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--lanes", type=int, default=2, choices=[1, 2],
                    help="search lanes per GPU in the pipelined (value) loop: 2 = consecutive batches alternate between the index "
                         "and a workspace clone on two streams, so one batch's selection/rescoring overlaps the next batch's scan")
    ap.add_argument("--exchange", default="auto", choices=["auto", "peer", "nccl"],
                    help="N>1, row sharding: how the per-shard lists meet -- peer-memory push + flag-waiting merge kernel, or NCCL "
                         "all-gather + merge; auto = push for the blocking (e2e) call, all-gather for the pipelined (value) loop")
    ap.add_argument("--sharding", default="rows", choices=["rows", "replicas"],
                    help="N > 1: 'rows' (default) = the database row-sharded over the GPUs + NCCL candidate merge (strong scaling, the "
                         "north-star layout); 'replicas' = every GPU holds the whole database and answers its own batches (weak scaling)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
