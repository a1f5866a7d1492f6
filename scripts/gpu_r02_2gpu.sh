#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_graph_and_state.py tests/test_gpu_configs.py -q 2>&1 | tail -3
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-full-generate > gpurun_out/bench_r02_${N}gpu.json 2> gpurun_out/bench_r02_${N}gpu.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/bench_r02_${N}gpu.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r02_${N}gpu.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'train', {k:d['train'][k] for k in ('ms_per_step','value')}, 'strong', d['train'].get('strong'))
PY
