#!/bin/bash
# Compact-index visit: its own tests first, the whole GPU suite (the default changed), scan layout A/B, a short bench line.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_compact.py -q -x > gpurun_out/cmp_pytest_compact.log 2>&1; tail -25 gpurun_out/cmp_pytest_compact.log
timeout 400 python tools/scan_probe.py > gpurun_out/cmp_scan_probe.log 2>&1; cat gpurun_out/cmp_scan_probe.log | tail -20
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/cmp_pytest.log 2>&1; tail -8 gpurun_out/cmp_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/cmp_bench_20.json 2> gpurun_out/cmp_bench_20.err
python - <<PY
import json
d=json.loads(open("gpurun_out/cmp_bench_20.json").read())
print(round(d["value"]), round(d["ms_per_step"]*1e3,1), "e2e", round(d["e2e"]["value"]), "gemm", round(d["roofline"]["kernel_ms"]*1e3,1), "frac", round(d["roofline"]["frac"],3), d["clocks"])
for k,v in d["extra_configs"].items(): print(k, {a:b for a,b in v.items() if a != "workload"})
PY
tail -3 gpurun_out/cmp_bench_20.err
