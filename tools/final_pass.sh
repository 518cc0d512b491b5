#!/bin/bash
# Driver-style pass of the whole repository on one GPU: `pytest -m gpu`, smoke(), both bench arms, the frame-tail and
# serving-loop tools.  usage (under gpurun): bash tools/final_pass.sh <prefix>   -> gpurun_out/<prefix>_*
P=${1:-final}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/ -x -q -m gpu > $O/${P}_pytest_gpu.log 2>&1; tail -2 $O/${P}_pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > $O/${P}_smoke.log 2>&1; tail -1 $O/${P}_smoke.log
timeout 400 python bench.py > $O/${P}_bench_1gpu.json 2> $O/${P}_bench.err; tail -c 200 $O/${P}_bench_1gpu.json; echo
timeout 300 python bench.py --impl reference > $O/${P}_bench_reference.json 2>> $O/${P}_bench.err; tail -c 150 $O/${P}_bench_reference.json; echo
timeout 300 python tools/frame_decoder_bench.py --batch 1 2 4 8 16 32 128 256 > $O/${P}_fd_bench.jsonl 2>&1
timeout 200 python tools/frame_decoder_bench.py --preset FD_500M --batch 1 8 >> $O/${P}_fd_bench.jsonl 2>&1
cut -c1-140 $O/${P}_fd_bench.jsonl
timeout 300 python tools/serve_loop_bench.py > $O/${P}_serve_loop.jsonl 2>> $O/${P}_bench.err; cut -c1-200 $O/${P}_serve_loop.jsonl
