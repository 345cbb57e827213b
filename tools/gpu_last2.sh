#!/bin/bash
# Closing visit after the compact-index change: GPU suite, smoke, both bench lines, ncu capture of the tiled scan, launch lists.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/fin_pytest.log 2>&1; tail -3 gpurun_out/fin_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/fin_bench_20.json 2> gpurun_out/fin_bench_20.err
timeout 600 python bench.py > gpurun_out/fin_bench_300.json 2> gpurun_out/fin_bench_300.err
python - <<PY
import json
for f in ("gpurun_out/fin_bench_20.json","gpurun_out/fin_bench_300.json"):
    d=json.loads(open(f).read())
    print(f, round(d["value"]), round(d["ms_per_step"]*1e3,1), "e2e", round(d["e2e"]["value"]), "gemm", round(d["roofline"]["kernel_ms"]*1e3,1), "frac", round(d["roofline"]["frac"],3), d["clocks"], {k:(round(v["roofline"].get("frac", v["roofline"].get("frac_whole_call_of_sustained",0)),3)) for k,v in d["extra_configs"].items()})
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:scan_scores -s 1 -c 1 -o gpurun_out/fin_scan python tools/profile_step.py 1007000 3 > gpurun_out/fin_ncu_scan.log 2>&1; tail -1 gpurun_out/fin_ncu_scan.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/fin_launches.csv python tools/profile_step.py 1007000 2 > gpurun_out/fin_ncu_list.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/fin_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-configs > gpurun_out/fin_ncu_bench.log 2>&1
ls -la gpurun_out/fin_*
