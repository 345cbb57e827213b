"""Would cross-step overlap pay?  Two indexes over the same rows (own workspaces), alternating on two streams, vs one
index on one stream.  python tools/overlap_probe.py [rows]"""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("image-search-engine-for-historical-research_b200")
bench = importlib.import_module("bench")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_007_000
dev = torch.device("cuda", 0)
K, NQ = 100, 70
rows = bench.synth_rows_device(torch, n, 2048, dev, 0)
queries = bench.synth_rows_device(torch, NQ, 2048, dev, 1)
ixs = [pkg.ExactIndex.from_device(rows.data_ptr(), n, 2048, 0) for _ in range(2)]
outs = [(torch.empty((NQ, K), dtype=torch.int64, device=dev), torch.empty((NQ, K), dtype=torch.float32, device=dev),
         torch.zeros((NQ,), dtype=torch.int32, device=dev)) for _ in range(2)]
streams = [torch.cuda.Stream(), torch.cuda.Stream()]

def step(i, two):
    j = i & 1 if two else 0
    st = streams[j] if two else streams[0]
    ids, sims, status = outs[j]
    ixs[j].search_device(queries.data_ptr(), NQ, K, ids.data_ptr(), sims.data_ptr(), status_ptr=status.data_ptr(), stream=st.cuda_stream)

for two in (False, True, False, True):
    for i in range(10): step(i, two)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(torch.cuda.default_stream())
        for s in streams: s.wait_stream(torch.cuda.default_stream())
        for i in range(300): step(i, two)
        for s in streams: torch.cuda.default_stream().wait_stream(s)
        e1.record(torch.cuda.default_stream()); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 300)
    print(f"n={n} {'two indexes / two streams' if two else 'one index / one stream  '}: {best*1e3:.1f} us/step", flush=True)
ok = torch.equal(outs[0][0], outs[1][0])
print("results equal across the two indexes:", ok)
