#!/bin/bash
# Last single-GPU visit of the round: whole GPU suite, smoke, build/store rates, the bench line, ncu captures of the final kernels.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/last_pytest.log 2>&1; tail -3 gpurun_out/last_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python tools/build_probe.py 250000 > gpurun_out/last_build.log 2>&1; tail -7 gpurun_out/last_build.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/last_bench_20.json 2> gpurun_out/last_bench_20.err
timeout 600 python bench.py > gpurun_out/last_bench_300.json 2> gpurun_out/last_bench_300.err
python - <<PY
import json
for f in ("gpurun_out/last_bench_20.json","gpurun_out/last_bench_300.json"):
    d=json.loads(open(f).read())
    print(f, round(d["value"]), round(d["ms_per_step"]*1e3,1), "e2e", round(d["e2e"]["value"]), "gemm", round(d["roofline"]["kernel_ms"]*1e3,1), "frac", round(d["roofline"]["frac"],3), d["clocks"], {k:(round(v["roofline"].get("frac", v["roofline"].get("frac_whole_call_of_sustained",0)),3)) for k,v in d["extra_configs"].items()})
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_topk -s 1 -c 1 -o gpurun_out/last_gemm python tools/profile_step.py 1007000 3 > gpurun_out/last_ncu_gemm.log 2>&1; tail -1 gpurun_out/last_ncu_gemm.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:finalise_cluster -s 1 -c 1 -o gpurun_out/last_fin python tools/profile_step.py 1007000 3 > gpurun_out/last_ncu_fin.log 2>&1; tail -1 gpurun_out/last_ncu_fin.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/last_launches.csv python tools/profile_step.py 1007000 2 > gpurun_out/last_ncu_list.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/last_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/last_ncu_bench.log 2>&1
ls gpurun_out/last_*
