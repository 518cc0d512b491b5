#!/bin/bash
# A/B of the tcgen05 attention variants on the benchmark shape (B=64, H=16, T=3000, hd=64); one line per variant.
run() { echo -n "$* : "; env "$@" timeout 120 python tools/op_bench.py attn 10 2>&1 | tail -1; }
run FRT2_ATTN_VER=3
run FRT2_ATTN_VER=4
run FRT2_ATTN_VER=4 FRT2_A4_EMU=0
run FRT2_ATTN_VER=4 FRT2_A4_PROBE=0
run FRT2_ATTN_VER=4 FRT2_A4_ONE_ITEM=1
