#!/bin/bash
# round 2, run H: full GPU suite, compute-sanitizer passes, full bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_margins.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
for tool in memcheck racecheck synccheck; do
  VF_SANITIZE_GRAPHS=0 timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_smoke.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|ok " gpurun_out/sanitize_$tool.log | tail -4
done
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02_h.json 2> gpurun_out/bench_r02_h.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/bench_r02_h.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02_h.json'))
print({k:d[k] for k in ('value','ms_per_step')}, 'frac', d['roofline']['frac'], 'gen_full', d['generate_full']['value'], 'train', d['train']['ms_per_step'], d['train']['value'], 'strong', d['train'].get('strong'))
PY
