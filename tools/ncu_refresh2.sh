#!/bin/bash
# final re-capture of the prefill attention kernels (tcgen05 at head dim 128, mma.sync at 64) and two pinned --set full reports
set -u
cd "$(dirname "$0")/.."
NCU=${NCU:-ncu}
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__shared_mem_per_block_dynamic,launch__grid_size,launch__block_size,lts__t_sector_hit_rate.pct"
C="--metrics $M --clock-control none --kernel-name-base demangled --csv"
VCLA_ATTN_TC=2 timeout 300 $NCU $C -k "regex:attn_prefill" -c 12 --log-file gpurun_out/ncu_metrics_attn_B8.csv python tools/profile_step.py 8 > /dev/null 2>&1; echo "attn B8 rc=$?"
VCLA_ATTN_TC=2 VCLA_PROFILE_T=1024 timeout 300 $NCU $C -k "regex:attn_prefill" -c 12 --log-file gpurun_out/ncu_metrics_attn_B16_long.csv python tools/profile_step.py 16 > /dev/null 2>&1; echo "attn long rc=$?"
F="--set full --clock-control none --import-source on --kernel-name-base demangled -f"
timeout 300 $NCU $F -k "regex:gemm_tc_kernel.*256.*6.*2>" -s 8 -c 1 -o gpurun_out/ncu_full_gemm_2cta python tools/profile_step.py 8 > gpurun_out/ncu_full_2cta.log 2>&1; echo "full 2cta rc=$?"
VCLA_PROFILE_T=1024 timeout 300 $NCU $F -k "regex:attn_prefill_tc_kernel.*128" -s 0 -c 1 -o gpurun_out/ncu_full_attn_tc python tools/profile_step.py 16 > gpurun_out/ncu_full_attn_tc.log 2>&1; echo "full attn rc=$?"
ls -la gpurun_out/*.ncu-rep
