#!/bin/bash
# round 2, run C: full GPU suite; A/B of the exact SiLU variant on the headline margins and the gn_apply class
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_margins.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
cp gpurun_out/parity_margins.txt gpurun_out/parity_margins_default.txt
export VF_B200_LIB=$PWD/view_fusion_b200/libvf_ab_silu.so
VF_MARGINS_FILE=$PWD/gpurun_out/parity_margins_silu_exact.txt timeout 600 python -m pytest tests/test_gpu_headline_n6.py tests/test_gpu_unet.py -q 2>&1 | tail -3
grep "bf16" gpurun_out/parity_margins_silu_exact.txt | grep "eps\|PSNR"
timeout 600 python bench.py --steps 20 --warmup 5 --no-library-baseline --no-extra-configs --no-full-generate --no-cpu-baseline --no-train > gpurun_out/bench_r02_c_silu.json 2> gpurun_out/bench_r02_c.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02_c_silu.json'))
print("exact silu:", {k:d[k] for k in ('value','ms_per_step')}, d['kernel_classes']['gn_apply'])
PY
unset VF_B200_LIB
timeout 600 python bench.py --steps 20 --warmup 5 --no-library-baseline --no-extra-configs --no-full-generate --no-cpu-baseline --no-train > gpurun_out/bench_r02_c.json 2>> gpurun_out/bench_r02_c.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02_c.json'))
print("default:", {k:d[k] for k in ('value','ms_per_step')}, d['roofline']['frac'], d['kernel_classes'])
PY
