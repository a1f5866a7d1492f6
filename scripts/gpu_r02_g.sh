#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
(
GN_TAG=old VF_GN_STREAM=0 python scripts/gn_bench.py
GN_TAG=stream_2x VF_GS_STAGES=2 python scripts/gn_bench.py
GN_TAG=stream_3cta VF_B200_LIB=$PWD/view_fusion_b200/libvf_ab_gs3.so VF_GS_STAGES=2 VF_GS_TILE_KB=24 python scripts/gn_bench.py
GN_TAG=stream_3cta16 VF_B200_LIB=$PWD/view_fusion_b200/libvf_ab_gs3.so VF_GS_STAGES=2 VF_GS_TILE_KB=16 python scripts/gn_bench.py
GN_TAG=stream_4cta VF_B200_LIB=$PWD/view_fusion_b200/libvf_ab_gs4.so VF_GS_STAGES=2 VF_GS_TILE_KB=16 python scripts/gn_bench.py
GN_TAG=stream_4cta3s VF_B200_LIB=$PWD/view_fusion_b200/libvf_ab_gs4.so VF_GS_STAGES=3 VF_GS_TILE_KB=12 python scripts/gn_bench.py
) 2>&1 | grep -v "^$" | tee gpurun_out/gn_bench2.txt
