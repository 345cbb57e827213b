#!/bin/bash
# Very last visit: GPU suite + smoke on the final library, the driver-style bench line, ncu capture of the CTA-pair GEMM (1024 queries).
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/z_pytest.log 2>&1; tail -2 gpurun_out/z_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/z_bench_20.json 2> gpurun_out/z_bench_20.err
python - <<PY
import json
d=json.loads(open("gpurun_out/z_bench_20.json").read())
print(round(d["value"]), round(d["ms_per_step"]*1e3,1), "e2e", round(d["e2e"]["value"]), "gemm", round(d["roofline"]["kernel_ms"]*1e3,1), "frac", round(d["roofline"]["frac"],3), d["clocks"], {k:(round(v["roofline"].get("frac", v["roofline"].get("frac_whole_call_of_sustained",0)),3), v.get("e2e_ms_per_query")) for k,v in d["extra_configs"].items()})
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_topk -s 3 -c 1 -o gpurun_out/z_pair python tools/profile_step.py 1007000 2 > gpurun_out/z_ncu_pair.log 2>&1; tail -1 gpurun_out/z_ncu_pair.log
