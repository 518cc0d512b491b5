#!/bin/bash
# ncu evidence of one round: launch lists (gpu__time_duration, cold-cache and serialised: compare SHARES) and --set full
# captures of the top kernels of the throughput step and of the streaming token step.  Run under gpurun on ONE GPU:
#   gpurun --timeout 1200 -- 'bash tools/profile_pass.sh gpurun_out/prof'
# Each ncu command runs only after the same command line has exited 0 without ncu.
OUT=${1:-gpurun_out/prof}
mkdir -p "$OUT"
BENCH="python bench.py --quick --steps 1 --warmup 1"
STREAM="python tools/stream_bench.py --eager"
export STEPS=17
$BENCH > "$OUT/plain_bench.log" 2>&1 || { echo "plain bench failed"; exit 1; }
$STREAM > "$OUT/plain_stream.log" 2>&1 || { echo "plain stream failed"; exit 1; }
# 1. every launch of the second (timed) decode step: 111 launches after the 111 of the warm-up step
ncu --metrics gpu__time_duration.sum --clock-control none -s 111 -c 111 --csv --log-file "$OUT/launches.csv" $BENCH > "$OUT/ncu_list.log" 2>&1
echo "list=$?"
# 2. one transformer layer of the step (row_stats, QKV, attention, out-proj, row_stats, fc1, fc2 ...) + the tail kernels
ncu --set full --clock-control none -k regex:"gemm_tc|attention_t|row_stats" -s 9 -c 9 -o "$OUT/prof_layer" -f $BENCH > "$OUT/ncu_layer.log" 2>&1
echo "layer=$?"
ncu --set full --clock-control none -k regex:"overlap_add|layer_norm|rvq_gather" -c 4 -o "$OUT/prof_misc" -f $BENCH > "$OUT/ncu_misc.log" 2>&1
echo "misc=$?"
# 3. the streaming token step, launch by launch (eager: the graph's kernels one by one), two tokens
ncu --metrics gpu__time_duration.sum --clock-control none -s 180 -c 180 --csv --log-file "$OUT/stream_launches.csv" $STREAM > "$OUT/ncu_stream_list.log" 2>&1
echo "streamlist=$?"
ncu --set full --clock-control none -k regex:"gemm_skinny_kernel|attention_warp" -s 100 -c 6 -o "$OUT/prof_stream" -f $STREAM > "$OUT/ncu_stream.log" 2>&1
echo "stream=$?"
ls -la "$OUT"
# 4. DRAM bytes of every launch of the step (bench.py reports the GEMM's per-launch traffic from this: tools/dram_traffic.py)
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -s 111 -c 111 --csv --log-file "$OUT/launches_dram.csv" $BENCH > "$OUT/ncu_dram.log" 2>&1
echo "dram=$?"
