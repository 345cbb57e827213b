#!/bin/bash
# One-GPU evidence run for profiles/: probes + ncu captures of the dominant kernels.
tag=${1:-ev}
mkdir -p gpurun_out
timeout 300 python tools/build_probe.py 250000 > gpurun_out/${tag}_build.log 2>&1; tail -8 gpurun_out/${tag}_build.log
timeout 300 python tools/diffusion_probe.py > gpurun_out/${tag}_diffusion.log 2>&1; tail -9 gpurun_out/${tag}_diffusion.log
timeout 300 python tools/rank_probe.py > gpurun_out/${tag}_rank.log 2>&1; tail -3 gpurun_out/${tag}_rank.log
timeout 300 python tools/online_probe.py > gpurun_out/${tag}_online.log 2>&1; tail -2 gpurun_out/${tag}_online.log
timeout 300 python tools/diag_p.py > gpurun_out/${tag}_families.log 2>&1; grep -c "reruns  0" gpurun_out/${tag}_families.log
timeout 300 python tools/eps_sweep.py > gpurun_out/${tag}_eps_sweep.log 2>&1; tail -12 gpurun_out/${tag}_eps_sweep.log
timeout 300 python tools/profile_step.py 1007000 3 > gpurun_out/${tag}_plain.log 2>&1; tail -3 gpurun_out/${tag}_plain.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_topk -s 1 -c 1 -o gpurun_out/${tag}_gemm python tools/profile_step.py 1007000 3 > gpurun_out/${tag}_ncu_gemm.log 2>&1; tail -1 gpurun_out/${tag}_ncu_gemm.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:finalise_cluster -s 1 -c 1 -o gpurun_out/${tag}_fin python tools/profile_step.py 1007000 3 > gpurun_out/${tag}_ncu_fin.log 2>&1; tail -1 gpurun_out/${tag}_ncu_fin.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_scores -s 1 -c 1 -o gpurun_out/${tag}_scan python tools/profile_step.py 1007000 3 > gpurun_out/${tag}_ncu_scan.log 2>&1; tail -1 gpurun_out/${tag}_ncu_scan.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_topk -s 9 -c 1 -o gpurun_out/${tag}_gemm_pair python tools/profile_step.py 1007000 3 > gpurun_out/${tag}_ncu_pair.log 2>&1; tail -1 gpurun_out/${tag}_ncu_pair.log
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_bench_short.json 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_ncu_bench.log 2>&1
ls -la gpurun_out/${tag}_*
