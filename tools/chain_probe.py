"""Where the 70-query step goes: per-step device time and coarse-kernel time with the in-kernel threshold bootstrap on
and off, at the full database and at the 8-GPU shard size.

    python tools/chain_probe.py [steps]
"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("image-search-engine-for-historical-research_b200")
bench = importlib.import_module("bench")

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
dev = torch.device("cuda", 0)
queries = bench.synth_rows_device(torch, 70, 2048, dev, 1)
ids = torch.empty((70, 100), dtype=torch.int64, device=dev)
sims = torch.empty((70, 100), dtype=torch.float32, device=dev)
status = torch.zeros((70,), dtype=torch.int32, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for n in (1_007_000, 125_875):
    rows = bench.synth_rows_device(torch, n, 2048, dev, 0)
    ix = pkg.ExactIndex.from_device(rows.data_ptr(), n, 2048, 0)
    for nq in (70, 1):
        for inline, stages in (((1, 4), (1, 3), (0, 4)) if nq > 1 else ((1, 4),)):
            ix.set_param("inline_boot", inline)
            ix.set_param("gemm_stages", stages)
            q = queries[:nq].contiguous()
            for _ in range(20):
                ix.search_device(q.data_ptr(), nq, 100, ids.data_ptr(), sims.data_ptr(), status_ptr=status.data_ptr())
            torch.cuda.synchronize()
            e0.record()
            for _ in range(steps):
                ix.search_device(q.data_ptr(), nq, 100, ids.data_ptr(), sims.data_ptr(), status_ptr=status.data_ptr())
            e1.record()
            torch.cuda.synchronize()
            step_ms = e0.elapsed_time(e1) / steps
            ix.set_param("timing", 1)
            ks, ts = [], []
            for _ in range(20):
                ix.search_device(q.data_ptr(), nq, 100, ids.data_ptr(), sims.data_ptr(), status_ptr=status.data_ptr())
                st = ix.stats()
                ks.append(st["ms_coarse"]); ts.append(st["ms_total"])
            ix.set_param("timing", 0)
            print(f"rows {n:8d} nq {nq:3d} inline_boot {inline} ring {stages}: step {step_ms*1e3:7.1f} us back to back; one call {sum(ts)/20*1e3:7.1f} us, coarse kernel {sum(ks)/20*1e3:7.1f} us, "
                  f"launches {st['gpu_launches']}, candidates/query {st['n_candidates']/nq:.0f}, uncertified {int(status[:nq].sum())}", flush=True)
    ix.close()
    del rows
    torch.cuda.empty_cache()

# timeline of the in-kernel bootstrap at the shard size (per-CTA globaltimer stamps)
import ctypes as C
import numpy as np
nat = importlib.import_module("image-search-engine-for-historical-research_b200._native")
for n in (125_875, 1_007_000):
    rows = bench.synth_rows_device(torch, n, 2048, dev, 0)
    ix = pkg.ExactIndex.from_device(rows.data_ptr(), n, 2048, 0)
    ix.set_param("boot_trace", 1)
    ix.set_param("gemm_stages", 4); ix.set_param("inline_boot", 1)
    for _ in range(5):
        ix.search_device(queries.data_ptr(), 70, 100, ids.data_ptr(), sims.data_ptr(), status_ptr=status.data_ptr())
    torch.cuda.synchronize()
    for which, width, names in ((0, 8, ["start", "tile1", "arrived", "all_arrived(owner)", "thr_out", "acc0_released", "done", "own_thr_selected(owner)"]),
                                (1, 10, ["start", "sizes", "gathered", "cut", "collected", "rescored", "ticket", "sorted(last)", "emitted(last)"])):
        buf = np.zeros((2048, width), dtype=np.uint64)
        g = C.c_int(0)
        nat.check(nat.load().xs_debug_trace(ix._h, which, buf.ctypes.data, 2048, C.byref(g)), "trace")
        t = buf[: g.value].astype(np.int64)
        t = t[t[:, 0] > 0]
        if not len(t):
            continue
        t0 = t[:, 0].min()
        rel = (t - t0) / 1e3
        print(f"{'bootstrap GEMM' if which == 0 else 'fused finalise'} timeline, {n} rows, {len(t)} CTAs (us after the first CTA's start): ", flush=True)
        for j, nm in enumerate(names):
            col = rel[:, j][t[:, j] > 0]
            if col.size:
                print(f"  {nm:24s} min {col.min():7.1f}  median {np.median(col):7.1f}  max {col.max():7.1f}   ({col.size} CTAs)")
    ix.close()
    del rows
    torch.cuda.empty_cache()
