"""A/B timing of library builds: pipelined device-side steps for a few shapes.  python tools/ab_probe.py LABEL"""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("image-search-engine-for-historical-research_b200")
bench = importlib.import_module("bench")
label = sys.argv[1] if len(sys.argv) > 1 else "-"
dev = torch.device("cuda", 0)
K = 100
out = []
for n in (1_007_000, 125_875):
    rows = bench.synth_rows_device(torch, n, 2048, dev, 0)
    queries = bench.synth_rows_device(torch, 1024, 2048, dev, 1)
    ix = pkg.ExactIndex.from_device(rows.data_ptr(), n, 2048, 0)
    ids = torch.empty((1024, K), dtype=torch.int64, device=dev); sims = torch.empty((1024, K), dtype=torch.float32, device=dev)
    status = torch.zeros((1024,), dtype=torch.int32, device=dev)
    for nq, reps in ((70, 300), (1, 300), (1024, 30)):
        def step():
            ix.search_device(queries.data_ptr(), nq, K, ids.data_ptr(), sims.data_ptr(), status_ptr=status.data_ptr())
        for _ in range(5): step()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps): step()
            e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / reps)
        out.append(f"n={n} nq={nq}: {best*1e3:.1f} us")
        print(f"[{label}] {out[-1]}", file=sys.stderr, flush=True)
        assert int(status[:nq].sum()) == 0
    ix.close(); del rows
print(f"[{label}] " + " | ".join(out), flush=True)
