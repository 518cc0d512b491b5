#!/bin/bash
# ncu evidence of the encode side (round 2): launch list of one whole-path encode (log-mel, both encoders, adaptor,
# downsample, RVQ chain) and --set full captures of its own kernels.  Run under gpurun on ONE GPU:
#   gpurun --timeout 1500 -- 'bash tools/profile_encode.sh gpurun_out/prof_enc'
OUT=${1:-gpurun_out/prof_enc}
mkdir -p "$OUT"
CMD="python tools/encode_bench.py --audio --batch 32 --reps 1"
$CMD > "$OUT/plain.log" 2>&1 || { echo "plain run failed"; exit 1; }
# the bench makes 2 warm-up passes + 1 timed pass of (354 feature launches + 100 RVQ launches): list the last pass
ncu --metrics gpu__time_duration.sum --clock-control none -s 908 -c 454 --csv --log-file "$OUT/enc_launches.csv" $CMD > "$OUT/ncu_list.log" 2>&1
echo "list=$?"
ncu --set full --clock-control none -k regex:"mel_power|mel_norm|silu_mul|cvt_rows|add_pos|layer_norm_f32|split_rows|argmax_update|gather_z" -c 12 -o "$OUT/prof_enc_kernels" -f $CMD > "$OUT/ncu_k.log" 2>&1
echo "kernels=$?"
ls -la "$OUT"
