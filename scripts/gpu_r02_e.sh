#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_bwd_ops.py tests/test_gpu_ops.py -q -k "groupnorm or gn" 2>&1 | tail -3
(
GN_TAG=old VF_GN_STREAM=0 python scripts/gn_bench.py
GN_TAG=stream python scripts/gn_bench.py
GN_TAG=stream_16KB VF_GS_TILE_KB=16 python scripts/gn_bench.py
GN_TAG=stream_48KB_1cta VF_GS_TILE_KB=48 VF_GS_CTAS=1 python scripts/gn_bench.py
GN_TAG=stream_2stage VF_GS_STAGES=2 python scripts/gn_bench.py
GN_TAG=stream_8KB VF_GS_TILE_KB=8 python scripts/gn_bench.py
) 2>&1 | tee gpurun_out/gn_bench.txt
