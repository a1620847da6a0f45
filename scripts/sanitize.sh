#!/usr/bin/env bash
# compute-sanitizer pass over one tiny invocation of every kernel family (run under gpurun; slow: keep shapes tiny).
# usage: scripts/sanitize.sh [memcheck|racecheck|synccheck]
tool="${1:-memcheck}"
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool "$tool" --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_gpu_model.py -x -q -m gpu -k "test_bf16_loss_and_grad_vs_oracle or test_fp32_loss_and_grad_match_oracle" \
  -p no:cacheprovider > "gpurun_out/sanitize_${tool}_model.log" 2>&1
echo "model: exit $? : $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/sanitize_${tool}_model.log | tail -n 2 | tr '\n' ' ')"
timeout 900 compute-sanitizer --tool "$tool" --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_gpu_decode.py -x -q -m gpu -k "test_decode_logits_match_oracle_forward" \
  -p no:cacheprovider > "gpurun_out/sanitize_${tool}_decode.log" 2>&1
echo "decode: exit $? : $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/sanitize_${tool}_decode.log | tail -n 2 | tr '\n' ' ')"
# kernels the tiny model configs do not reach: CTA-pair GEMM (all epilogues), paired tcgen05 attention, streaming LN backward
timeout 1200 compute-sanitizer --tool "$tool" --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_attn_mma.py tests/test_gpu_elementwise.py -x -q -m gpu \
  -k "tc2_pair or tcgen05 or stream_path" -p no:cacheprovider > "gpurun_out/sanitize_${tool}_kernels.log" 2>&1
echo "kernels: exit $? : $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/sanitize_${tool}_kernels.log | tail -n 2 | tr '\n' ' ')"
