#!/bin/bash
# Runs the GPU parity suite in separate processes (a CUDA fault in one group must not poison the others).
# usage: tools/gpu_suite.sh [outdir]
OUT=${1:-gpurun_out}
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > "$OUT/nvidia_smi.csv" 2>&1
run() {  # name, timeout, pytest args...
  local name=$1 to=$2; shift 2
  timeout "$to" python -m pytest "$@" -q -rA --tb=short -s -m gpu -p no:cacheprovider > "$OUT/$name.log" 2>&1
  echo "$name exit=$?" | tee -a "$OUT/summary.txt"
  grep -E "passed|failed|error" "$OUT/$name.log" | tail -1 | tee -a "$OUT/summary.txt"
}
: > "$OUT/summary.txt"
run ops_gemm 600 tests/test_gpu_ops.py -k "gemm"
run ops_ln 300 tests/test_gpu_ops.py -k layer_norm
run ops_ola 300 tests/test_gpu_ops.py -k overlap_add
run ops_resample 300 tests/test_gpu_ops.py -k resample
run ops_attn_warp 600 tests/test_gpu_ops.py -k "attention and warp"
run ops_attn_tc 600 tests/test_gpu_ops.py -k "attention and tc"
run decode_simt 900 tests/test_gpu_decode.py -k "offline_decode and simt"
run decode_tcgemm 900 tests/test_gpu_decode.py -k "offline_decode and tc_gemm"
run decode_product 900 tests/test_gpu_decode.py -k "offline_decode and product"
run decode_misc 900 tests/test_gpu_decode.py -k "rvq or index or varlen or pcm16"
run decode_stream 900 tests/test_gpu_decode.py -k "stream or graph"
run decode_c0 900 tests/test_gpu_decode.py -k "c0"
run pool 900 tests/test_gpu_pool.py
run streams 900 tests/test_gpu_streams.py
run rvq_encode 900 tests/test_gpu_rvq_encode.py
run encoder 900 tests/test_gpu_encoder.py
run frame_decoder 900 tests/test_gpu_frame_decoder.py
timeout 600 python __graft_entry__.py --smoke > "$OUT/smoke.log" 2>&1; echo "smoke exit=$?" | tee -a "$OUT/summary.txt"
grep -h "\[parity\]\|^gemm\|^attention\|smoke:" "$OUT"/*.log > "$OUT/parity_lines.txt" 2>/dev/null
cat "$OUT/summary.txt"
