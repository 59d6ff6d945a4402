#!/bin/bash
# Launch list of ONE timed step of the bench (run on the GPU box, after `python bench.py` exited 0 without ncu).
# usage: tools/ncu_launch_list.sh <launches per step> <out.csv>
# --profile: 1 warm-up + 1 timed step, e2e arm and CPU baseline skipped; -s skips the warm-up step's launches.
N=${1:-884}
OUT=${2:-gpurun_out/launches.csv}
K='regex:attention_tcgen05|gemm_bf16_tcgen05|gemv_bf16|camera_pose|cast_f32_bf16|dino_embed|gather_rows|im2col|layernorm|mean_pool|mrope_table|points_epilogue|qknorm_mrope|rmsnorm_routed|rope2d|split3'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s "$N" -c "$N" --csv --log-file "$OUT" \
    python bench.py --profile --steps 1 --warmup 1 > gpurun_out/ncu_bench.log 2>&1
