#!/bin/bash
# One GPU-box visit: tests, bench, chain probe, launch list.  Everything lands in gpurun_out/<tag>_*.
tag=${1:-r2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/${tag}_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -15 gpurun_out/${tag}_pytest.log
timeout 300 python tools/chain_probe.py > gpurun_out/${tag}_chain.log 2>&1
cat gpurun_out/${tag}_chain.log
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/${tag}_bench.json").read())
print({k:d[k] for k in ("value","ms_per_step")}, d["e2e"], d["roofline"]["kernel_ms"], d["clocks"])
x=d.get("extra_configs",{})
for k,v in x.items(): print(k, {a:b for a,b in v.items() if a not in ("workload",)})
PY
tail -3 gpurun_out/${tag}_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches.csv python tools/profile_step.py 1007000 2 > gpurun_out/${tag}_ncu.log 2>&1
python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/${tag}_launches.csv", errors="ignore")))
hdr=[i for i,r in enumerate(rows) if "Kernel Name" in r]
if hdr:
    h=rows[hdr[0]]; kn=h.index("Kernel Name"); mv=h.index("Metric Value"); gs=h.index("Grid Size") if "Grid Size" in h else None
    for r in rows[hdr[0]+1:]:
        if len(r)>mv and not r[kn].startswith(("void at::","at::")) and "vectorized" not in r[kn] and "distribution" not in r[kn]:
            print(r[kn][:70], r[gs] if gs is not None else "", r[mv])
PY
