#!/bin/bash
# Re-capture, with the final code, the kernels that changed after tools/ncu_metrics.sh ran (tcgen05 attention, sampler) and the three
# `--set full` reports, each pinned to the intended launch.  Run under gpurun on ONE GPU.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NCU=${NCU:-ncu}
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__shared_mem_per_block_dynamic,launch__grid_size,launch__block_size,lts__t_sector_hit_rate.pct"
C="--metrics $M --clock-control none --kernel-name-base demangled --csv"
timeout 300 $NCU $C -k "regex:attn_prefill" -c 12 --log-file gpurun_out/ncu_metrics_attn_B8.csv python tools/profile_step.py 8 > /dev/null 2>&1; echo "attn B8 rc=$?"
VCLA_ATTN_TC=0 timeout 300 $NCU $C -k "regex:attn_prefill" -c 12 --log-file gpurun_out/ncu_metrics_attn_B8_mma.csv python tools/profile_step.py 8 > /dev/null 2>&1; echo "attn B8 mma rc=$?"
VCLA_PROFILE_T=1024 timeout 300 $NCU $C -k "regex:attn_prefill" -c 12 --log-file gpurun_out/ncu_metrics_attn_B16_long.csv python tools/profile_step.py 16 > /dev/null 2>&1; echo "attn long rc=$?"
VCLA_ATTN_TC=0 VCLA_PROFILE_T=1024 timeout 300 $NCU $C -k "regex:attn_prefill" -c 12 --log-file gpurun_out/ncu_metrics_attn_B16_long_mma.csv python tools/profile_step.py 16 > /dev/null 2>&1; echo "attn long mma rc=$?"
timeout 300 $NCU $C -k regex:dec_sample -c 4 --log-file gpurun_out/ncu_metrics_sampler.csv python tools/profile_step.py 8 --sample > /dev/null 2>&1; echo "sampler rc=$?"
F="--set full --clock-control none --import-source on --kernel-name-base demangled -f"
# decode GEMMs of profile_step.py (2 LLaMA layers): prefill's lm_head is csk launch 0; per decode step the order is qkv, o, gate_up, down (x2 layers), lm_head
timeout 300 $NCU $F -k regex:gemm_csk -s 3 -c 1 -o gpurun_out/ncu_full_csk_gate_up python tools/profile_step.py 8 > gpurun_out/ncu_full_csk.log 2>&1; echo "full csk rc=$?"
timeout 300 $NCU $F -k "regex:gemm_tc_kernel<256, 6, false, 2>" -s 8 -c 1 -o gpurun_out/ncu_full_gemm_2cta python tools/profile_step.py 8 > gpurun_out/ncu_full_2cta.log 2>&1; echo "full 2cta rc=$?"
timeout 300 $NCU $F -k "regex:attn_prefill_tc_kernel<128>" -s 0 -c 1 -o gpurun_out/ncu_full_attn_tc python tools/profile_step.py 8 > gpurun_out/ncu_full_attn_tc.log 2>&1; echo "full attn rc=$?"
grep -h "gemm_csk\|gemm_tc_kernel\|attn_prefill" gpurun_out/ncu_full_*.log | head -6
ls -la gpurun_out/*.ncu-rep gpurun_out/ncu_metrics_attn* gpurun_out/ncu_metrics_sampler.csv
