#!/bin/bash
tag=${1:-d}
mkdir -p gpurun_out
timeout 300 python tools/chain_probe.py 100 > gpurun_out/${tag}_chain.log 2>&1; cat gpurun_out/${tag}_chain.log
timeout 300 python tools/diag_p.py > gpurun_out/${tag}_diagp.log 2>&1; cat gpurun_out/${tag}_diagp.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:finalise_fused -s 2 -c 1 -o gpurun_out/${tag}_fused python tools/profile_step.py 1007000 3 > gpurun_out/${tag}_ncu_fused.log 2>&1
tail -3 gpurun_out/${tag}_ncu_fused.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:prep_queries -s 2 -c 1 -o gpurun_out/${tag}_prep python tools/profile_step.py 1007000 3 > gpurun_out/${tag}_ncu_prep.log 2>&1
tail -3 gpurun_out/${tag}_ncu_prep.log
ls -la gpurun_out/${tag}_*
