"""Sustained 70-query steps (power-capped regime): ms/step, SM clock and power over a long loop.  python tools/sustain_probe.py LABEL [steps]"""
import importlib, os, sys, time, subprocess, threading, statistics
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("image-search-engine-for-historical-research_b200")
bench = importlib.import_module("bench")
label = sys.argv[1] if len(sys.argv) > 1 else "-"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
dev = torch.device("cuda", 0)
rows = bench.synth_rows_device(torch, 1_007_000, 2048, dev, 0)
queries = bench.synth_rows_device(torch, 70, 2048, dev, 1)
ix = pkg.ExactIndex.from_device(rows.data_ptr(), 1_007_000, 2048, 0)
ids = torch.empty((70, 100), dtype=torch.int64, device=dev); sims = torch.empty((70, 100), dtype=torch.float32, device=dev)
status = torch.zeros((70,), dtype=torch.int32, device=dev)
def step():
    ix.search_device(queries.data_ptr(), 70, 100, ids.data_ptr(), sims.data_ptr(), status_ptr=status.data_ptr())
for _ in range(200): step()
torch.cuda.synchronize()
lines = []
p = subprocess.Popen(["nvidia-smi", "--id=0", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE, text=True)
threading.Thread(target=lambda: [lines.append(l) for l in p.stdout], daemon=True).start()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps): step()
e1.record(); torch.cuda.synchronize()
time.sleep(0.1); p.terminate()
vals = [tuple(float(x) for x in l.split(",")) for l in lines if "," in l]
busy = vals[len(vals)//3:] if len(vals) > 6 else vals
print(f"[{label}] {e0.elapsed_time(e1)/steps*1e3:.1f} us/step over {steps} steps; SM {statistics.median(v[0] for v in busy):.0f} MHz, {statistics.median(v[1] for v in busy):.0f} W ({len(vals)} samples)", flush=True)
