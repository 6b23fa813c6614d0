#!/bin/bash
# `ncu --set full` over EVERY kernel launch of a shallow, full-width run of the path (tools/profile_step.py), one report per batch
# size, plus the image pre-processing kernels.  Run under gpurun on ONE GPU; summarise here (no GPU) with
#     python tools/ncu_kernel_table.py gpurun_out/ncu_all_*.ncu-rep > profiles/r2_ncu_kernels.json
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NCU=${NCU:-ncu}
COMMON="--set full --clock-control none --import-source on --kernel-name-base demangled -k regex:vcla:: -f"
timeout 900 $NCU $COMMON -c 260 -o gpurun_out/ncu_all_B8 python tools/profile_step.py 8 > gpurun_out/ncu_all_B8.log 2>&1; echo "B8 rc=$?"
timeout 900 $NCU $COMMON -c 260 -o gpurun_out/ncu_all_B32 python tools/profile_step.py 32 > gpurun_out/ncu_all_B32.log 2>&1; echo "B32 rc=$?"
VCLA_PROFILE_T=1024 timeout 900 $NCU $COMMON -c 260 -o gpurun_out/ncu_all_B16_long python tools/profile_step.py 16 > gpurun_out/ncu_all_B16_long.log 2>&1; echo "B16 long rc=$?"
timeout 600 $NCU $COMMON -k regex:dec_sample -c 4 -o gpurun_out/ncu_all_sampler python tools/profile_step.py 8 --sample > gpurun_out/ncu_all_sampler.log 2>&1; echo "sampler rc=$?"
timeout 600 $NCU $COMMON -k regex:pp_ -c 6 -o gpurun_out/ncu_all_preprocess python tools/preprocess_bench.py --sizes 1080x1920 --reps 3 --no-pil > gpurun_out/ncu_all_preprocess.log 2>&1; echo "preprocess rc=$?"
ls -la gpurun_out/*.ncu-rep
