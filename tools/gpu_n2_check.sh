#!/bin/bash
# Two-GPU sanity visit after the compact-index change: NCCL exchange tests and the bench line the driver runs at N=2.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_exchange.py -q -m gpu > gpurun_out/n2c_pytest.log 2>&1; tail -2 gpurun_out/n2c_pytest.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29617 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/n2c_bench.json 2> gpurun_out/n2c_bench.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/n2c_bench.json").read())
    print({k:d[k] for k in ("value","ms_per_step","n_gpus")}, d["e2e"]["value"], d["roofline"]["kernel_ms"], d["config"]["merged_parity_vs_exact_shards"], d["config"]["exact_reruns_total"], d["clocks"])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/n2c_bench.err").read()[-3000:])
PY
