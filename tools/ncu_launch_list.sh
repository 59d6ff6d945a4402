#!/bin/bash
# Launch list of the bench (run on the GPU box, after `python bench.py --profile ...` exited 0 without ncu).
# usage: tools/ncu_launch_list.sh <out.csv>
# --profile: 1 warm-up + 1 timed step, e2e arm and CPU baseline skipped.  Every launch of OUR kernels is listed
# (both steps); tools/summarize_profiles.py launches --last N keeps the N launches of the timed step.
OUT=${1:-gpurun_out/launches.csv}
K='regex:attention_tcgen05|gemm_bf16_tcgen05|gemv_|camera_pose|cast_f32_bf16|dino_embed|gather_rows|im2col|layernorm|mean_pool|mrope_table|points_epilogue|qknorm_mrope|rmsnorm_routed|rope2d|split3'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 4000 --csv --log-file "$OUT" \
    python bench.py --profile --steps 1 --warmup 1 > gpurun_out/ncu_bench.log 2>&1
