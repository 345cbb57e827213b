#!/bin/bash
# Multi-GPU visit: exchange tests, bench at N GPUs (both lane settings), cfg4 / cfg5 probes.  Logs in gpurun_out/<tag>_*.
tag=${1:-m}; n=${2:-2}; shift 2
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) "$@"; }
timeout 600 python -m pytest tests/test_gpu_exchange.py -q -m gpu > gpurun_out/${tag}_pytest.log 2>&1; tail -3 gpurun_out/${tag}_pytest.log
for lanes in 2 1; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $n --steps 300 --warmup 5 --lanes $lanes > gpurun_out/${tag}_bench_n${n}_l${lanes}.json 2> gpurun_out/${tag}_bench_n${n}_l${lanes}.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_bench_n${n}_l${lanes}.json").read())
    print("lanes ${lanes}:", {k:d[k] for k in ("value","ms_per_step","n_gpus")}, d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["kernel_ms"], d["config"]["merged_parity_vs_exact_shards"], d["config"]["exact_reruns_total"], d["clocks"])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/${tag}_bench_n${n}_l${lanes}.err").read()[-3000:])
PY
done
for extra in "$@"; do
  case $extra in
    cfg4) timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) tools/cfg4_probe.py 10000000 10000 3 peer > gpurun_out/${tag}_cfg4_n${n}.log 2>&1; grep '"probe"' gpurun_out/${tag}_cfg4_n${n}.log || tail -20 gpurun_out/${tag}_cfg4_n${n}.log;;
    cfg5) timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) tools/cfg5_probe.py 1000000 100 both > gpurun_out/${tag}_cfg5_n${n}.log 2>&1; grep '"probe"' gpurun_out/${tag}_cfg5_n${n}.log || tail -20 gpurun_out/${tag}_cfg5_n${n}.log;;
    ref) timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --impl reference --gpus $n --steps 5 --warmup 1 > gpurun_out/${tag}_ref_n${n}.json 2> gpurun_out/${tag}_ref_n${n}.err; cut -c1-600 gpurun_out/${tag}_ref_n${n}.json;;
  esac
done
