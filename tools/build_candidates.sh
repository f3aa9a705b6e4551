#!/bin/bash
# Builds (here, on the CPU: nvcc cross-compiles sm_100a) the A/B partners of the default module that DESIGN.md section 4
# lists as candidates, plus the counting build, into gpurun_variants/ (git-ignored, travels to the GPU box):
#   bash tools/build_candidates.sh && gpurun -- 'bash tools/visits/ab_candidates.sh'
set -eu
cd "$(dirname "$0")/.."
python tools/build_variant.py stats     ONLY=all NVCCDEF=SDQLB200_STATS
python tools/build_variant.py idx32     ONLY=all SDQLB200_IDX32=1
python tools/build_variant.py runagg    ONLY=all SDQLB200_RUNAGG=1
python tools/build_variant.py tier0smem ONLY=all SDQLB200_TIER0_SMEM=1
python tools/build_variant.py next3     ONLY=all SDQLB200_TIER0_SMEM=1 SDQLB200_IDX32=1 SDQLB200_RUNAGG=1
python tools/build_variant.py mat       ONLY=all SDQLB200_MATERIALISE=1
python tools/build_variant.py pack32    ONLY=all SDQLB200_PACK32=1
