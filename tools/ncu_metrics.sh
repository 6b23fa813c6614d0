#!/bin/bash
# Per-kernel ncu evidence with a SHORT metric list (a few passes per launch instead of the ~40 of --set full, CSV instead of reports):
# duration, DRAM bytes read / written, DRAM throughput %, tensor-pipe %, occupancy, registers -- for every kernel launch of a shallow
# full-width run of the path (tools/profile_step.py).  Plus ONE `--set full` report of the dominant decode GEMM.
# Run under gpurun on ONE GPU; summarise here with  python tools/ncu_csv_table.py gpurun_out/ncu_metrics_*.csv > profiles/r2_ncu_kernels.json
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NCU=${NCU:-ncu}
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__shared_mem_per_block_dynamic,launch__grid_size,launch__block_size,lts__t_sector_hit_rate.pct"
COMMON="--metrics $M --clock-control none --kernel-name-base demangled -k regex:vcla:: --csv"
run() { tag=$1; shift; timeout 420 $NCU $COMMON -c 400 --log-file gpurun_out/ncu_metrics_$tag.csv "$@" > gpurun_out/ncu_metrics_$tag.log 2>&1; echo "$tag rc=$? $(wc -l < gpurun_out/ncu_metrics_$tag.csv) lines"; }
run B8 python tools/profile_step.py 8
if [ "${VCLA_NCU_R1:-0}" = "1" ]; then VCLA_PREFILL_FUSED=0 VCLA_DECODE_SCHEDULE=unfused VCLA_GEMM_2CTA=0 VCLA_ATTN_TC=0 run B8_r1schedule python tools/profile_step.py 8; fi
run B32 python tools/profile_step.py 32
VCLA_PROFILE_T=1024 run B16_long python tools/profile_step.py 16
timeout 300 $NCU --metrics $M --clock-control none --kernel-name-base demangled -k regex:dec_sample --csv -c 4 --log-file gpurun_out/ncu_metrics_sampler.csv python tools/profile_step.py 8 --sample > gpurun_out/ncu_metrics_sampler.log 2>&1; echo "sampler rc=$?"
timeout 300 $NCU --metrics $M --clock-control none --kernel-name-base demangled -k regex:pp_ --csv -c 12 --log-file gpurun_out/ncu_metrics_preprocess.csv python tools/preprocess_bench.py --sizes 1080x1920 --reps 3 --no-pil > gpurun_out/ncu_metrics_preprocess.log 2>&1; echo "preprocess rc=$?"
# one full-set report of the dominant kernel (gate/up decode GEMM = 3rd csk launch of a layer), source view included
timeout 300 $NCU --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:gemm_csk -s 12 -c 1 -f -o gpurun_out/ncu_full_csk_gate_up python tools/profile_step.py 8 > gpurun_out/ncu_full_csk.log 2>&1; echo "full rc=$?"
timeout 300 $NCU --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:gemm_tc_kernel -s 8 -c 1 -f -o gpurun_out/ncu_full_gemm_2cta python tools/profile_step.py 8 > gpurun_out/ncu_full_2cta.log 2>&1; echo "full2 rc=$?"
timeout 300 $NCU --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:attn_prefill_tc -s 2 -c 1 -f -o gpurun_out/ncu_full_attn_tc python tools/profile_step.py 8 > gpurun_out/ncu_full_attn_tc.log 2>&1; echo "full3 rc=$?"
ls -la gpurun_out | tail -15; du -sh gpurun_out
