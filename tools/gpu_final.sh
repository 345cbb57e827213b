#!/bin/bash
# Final record at N GPUs: bench.py as the driver runs it (20 steps) and over 300 steps; at 8 GPUs also cfg4 and the exchange timeline.
n=${1:-8}; shift
mkdir -p gpurun_out
tr() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) "$@"; }
for steps in 300 20; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $n --steps $steps --warmup 5 > gpurun_out/final_n${n}_s${steps}.json 2> gpurun_out/final_n${n}_s${steps}.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/final_n${n}_s${steps}.json").read())
    print("N=${n} steps ${steps}:", round(d["value"]), round(d["ms_per_step"]*1e3,1), "us | e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"]*1e3,1), "us | gemm", round(d["roofline"]["kernel_ms"]*1e3,1), "| lanes", d["config"]["lanes"], d["config"]["merged_parity_vs_exact_shards"], d["config"]["exact_reruns_total"], d["clocks"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/final_n${n}_s${steps}.err").read()[-2500:])
PY
done
for extra in "$@"; do
  case $extra in
    cfg4) timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) tools/cfg4_probe.py 10000000 10000 3 peer > gpurun_out/final_cfg4_n${n}.log 2>&1; grep '"probe"' gpurun_out/final_cfg4_n${n}.log || tail -20 gpurun_out/final_cfg4_n${n}.log;;
    xprobe) XS_LANES=2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) tools/exchange_probe.py 200 > gpurun_out/final_xprobe_n${n}.log 2>&1; tail -16 gpurun_out/final_xprobe_n${n}.log;;
  esac
done
