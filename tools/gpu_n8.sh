#!/bin/bash
# N-GPU bench matrix: lanes x merge form
tag=${1:-n8}; n=${2:-8}
mkdir -p gpurun_out
for cfg in "2 2" "2 1" "1 1" "1 2"; do
  set -- $cfg; lanes=$1; mg=$2
  XS_PIPE_MERGE=$mg timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $n --steps 300 --warmup 5 --lanes $lanes > gpurun_out/${tag}_l${lanes}_m${mg}.json 2> gpurun_out/${tag}_l${lanes}_m${mg}.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_l${lanes}_m${mg}.json").read())
    print("lanes ${lanes} merge ${mg} (1 fused, 2 separate):", round(d["value"]), round(d["ms_per_step"]*1e3,1), "us | e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"]*1e3,1), "us | gemm", round(d["roofline"]["kernel_ms"]*1e3,1), d["config"]["merged_parity_vs_exact_shards"], d["config"]["exact_reruns_total"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/${tag}_l${lanes}_m${mg}.err").read()[-2000:])
PY
done
