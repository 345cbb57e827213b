import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("image-search-engine-for-historical-research_b200")
bench = importlib.import_module("bench")
dev = torch.device("cuda", 0)
rows = bench.synth_rows_device(torch, 1_007_000, 2048, dev, 0)
q = bench.synth_rows_device(torch, 70, 2048, dev, 1).cpu().numpy()
ix = pkg.ExactIndex.from_device(rows.data_ptr(), 1_007_000, 2048, 0)
for _ in range(4):
    ix.search(q, 100)
