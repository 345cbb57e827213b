"""Single GPU: per-step time of the native pipeline with one lane, two lanes (alternating) and two lanes in overlap mode
(3-stage scan ring + slim finalise CTAs), and the device-clock timeline of two consecutive steps.

    python tools/pipe_probe.py [rows] [steps]
"""
import ctypes as C
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("image-search-engine-for-historical-research_b200")
sharded = importlib.import_module("image-search-engine-for-historical-research_b200.sharded")
nat = importlib.import_module("image-search-engine-for-historical-research_b200._native")
bench = importlib.import_module("bench")

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_007_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
dev = torch.device("cuda", 0)
rows = bench.synth_rows_device(torch, n, 2048, dev, 0)
queries = bench.synth_rows_device(torch, 70, 2048, dev, 1)
index = pkg.ExactIndex.from_device(rows.data_ptr(), n, 2048, 0)
lib = nat.load()
for lanes, overlap in ((1, 0), (2, 0), (2, 1)):
    os.environ["XS_PIPE_OVERLAP"] = str(overlap)
    _, pipe = sharded.make_searcher(index, 0, lanes=lanes, pipeline=(70, 100))
    pend = [None]

    def piped():
        nxt = pipe.search_async(queries, 100)
        if pend[0] is not None:
            pend[0].result()
        pend[0] = nxt
    for _ in range(20):
        piped()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        piped()
    pend[0].result(); pend[0] = None
    e1.record()
    torch.cuda.synchronize()
    import time
    t_sub = t_col = 0.0
    for _ in range(100):
        ta = time.perf_counter(); nxt = pipe.search_async(queries, 100); tb = time.perf_counter()
        if pend[0] is not None:
            pend[0].result()
        tc = time.perf_counter()
        pend[0] = nxt
        t_sub += tb - ta; t_col += tc - tb
    pend[0].result(); pend[0] = None
    print(f"{n} rows, lanes {lanes}, overlap mode {overlap}: {e0.elapsed_time(e1) / steps * 1e3:.1f} us per 70-query step "
          f"(host: submit {t_sub / 100 * 1e6:.1f} us, collect incl. waiting {t_col / 100 * 1e6:.1f} us)", flush=True)
    hs = [C.c_void_p(lib.xs_pipeline_lane(pipe._h, l)) for l in range(lanes)]
    for h in hs:
        nat.check(lib.xs_set_param(h, b"boot_trace", 1.0), "set")
    for _ in range(8):
        piped()
    pend[0].result(); pend[0] = None
    torch.cuda.synchronize()
    ev = []
    for l, h in enumerate(hs):
        for which, width, name, c_end in ((0, 8, "scan+select GEMM", 6), (1, 10, "finalise", 8)):
            buf = np.zeros((2048, width), dtype=np.uint64)
            g = C.c_int(0)
            nat.check(lib.xs_debug_trace(h, which, buf.ctypes.data, 2048, C.byref(g)), "trace")
            t = buf[: g.value].astype(np.int64)
            t = t[t[:, 0] > 0]
            if len(t):
                ends = t[:, c_end][t[:, c_end] > 0]
                ev.append((int(t[:, 0].min()), int(np.median(t[:, 0])), int(t[:, 0].max()), int(ends.max() if len(ends) else t.max()), f"lane {l} {name}"))
    t0 = min(e[0] for e in ev)
    for a_, m_, x_, b_, nm in sorted(ev):
        print(f"    {nm:26s} first CTA {(a_ - t0) / 1e3:7.1f}  median CTA start {(m_ - t0) / 1e3:7.1f}  last CTA start {(x_ - t0) / 1e3:7.1f}  end {(b_ - t0) / 1e3:7.1f} us")
    for h in hs:
        nat.check(lib.xs_set_param(h, b"boot_trace", 0.0), "set")
    pipe.close()
