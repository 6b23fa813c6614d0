#!/bin/bash
# One `ncu --set full` capture per kernel of the path (run under gpurun, ONE GPU; ~40 replays per profiled launch, so each
# run profiles 2 launches of one kernel inside a short prefill + 5-token decode at batch 8).  Reports land in
# gpurun_out/ncu_<tag>.ncu-rep; summarise them here (no GPU needed) with
#     python tools/ncu_summarize.py gpurun_out/ncu_*.ncu-rep > profiles/rN_ncu_kernels.json
# Usage: bash tools/ncu_path.sh [tag ...]        (no tags = all)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NCU=${NCU:-ncu}
COMMON="--set full --clock-control none --import-source on --kernel-name-base demangled"
STEP="python tools/profile_gemm.py step 8"

# tag | demangled-name regex | launches to skip (lands on a warm, mid-run instance) | command
KERNELS=(
  "gemm_prefill_bn256|gemm_tc_kernel<.*256, .*4, .*0>|20|$STEP"
  "gemm_prefill_bn128|gemm_tc_kernel<.*128, .*6, .*0>|10|$STEP"
  "gemm_decode_swapab16|gemm_tc_kernel<.*16, .*5, .*1>|200|$STEP"
  "attn_prefill|attn_prefill_kernel|10|$STEP"
  "attn_decode|attn_decode_kernel|40|$STEP"
  "dec_resid_norm|dec_resid_norm_kernel|80|$STEP"
  "dec_silu_mul|dec_silu_mul_kernel|40|$STEP"
  "dec_logits|dec_logits_stage1_kernel|2|$STEP"
  "layernorm|layernorm_kernel|10|$STEP"
  "rmsnorm|rmsnorm_kernel|4|$STEP"
  "rope_and_cache|rope_and_cache_kernel|4|$STEP"
  "im2col|im2col_kernel|0|$STEP"
  "attn_decode_persistent|attn_decode_persistent_kernel|40|python tools/profile_gemm.py step 32"
  "pp_hpass|pp_hpass_kernel|3|python tools/preprocess_bench.py --sizes 1080x1920 --reps 5 --no-pil"
  "pp_vpass|pp_vpass_kernel|3|python tools/preprocess_bench.py --sizes 1080x1920 --reps 5 --no-pil"
)

want=("$@")
for row in "${KERNELS[@]}"; do
  IFS='|' read -r tag regex skip cmd <<<"$row"
  if [ ${#want[@]} -gt 0 ] && [[ ! " ${want[*]} " =~ " $tag " ]]; then continue; fi
  echo "=== $tag ($regex, skip $skip): $cmd"
  timeout 600 $NCU $COMMON -k "regex:$regex" -s "$skip" -c 2 -f -o "gpurun_out/ncu_$tag" $cmd > "gpurun_out/ncu_$tag.log" 2>&1
  echo "    rc=$? $(ls -la gpurun_out/ncu_$tag.ncu-rep 2>/dev/null | awk '{print $5" bytes"}')"
done
