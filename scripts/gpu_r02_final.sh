#!/bin/bash
# round 2 evidence run: tests + margins, smoke, sanitizers, per-layer table, ncu launch lists + conv counters, full bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_margins.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/smoke.log
for tool in memcheck racecheck synccheck; do
  VF_SANITIZE_GRAPHS=0 timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_smoke.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/sanitize_$tool.log | tail -1
done
timeout 300 python scripts/layer_table.py > gpurun_out/r02_layer_table.txt 2>&1; tail -1 gpurun_out/r02_layer_table.txt
bash scripts/gpu_r02_prof.sh
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02_final.out 2> gpurun_out/bench_r02_final.err; echo "bench rc=$?"; tail -c 800 gpurun_out/bench_r02_final.err
tail -1 gpurun_out/bench_r02_final.out > gpurun_out/r02_bench.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 2>/dev/null | tail -1 > gpurun_out/r02_bench_reference_arm.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench.json'))
print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'gen_full', d['generate_full']['value'], 'lib', {k:v.get('ms_per_step') for k,v in d['library_baseline'].items() if isinstance(v,dict)}, 'train', d['train']['ms_per_step'], d['train']['value'], 'strong', d['train']['strong']['value'])
PY
