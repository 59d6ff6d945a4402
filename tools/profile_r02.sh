#!/bin/bash
# Round-2 evidence run (on the GPU box, one GPU): plain bench first, then the ncu launch list of the same command,
# then ONE `ncu --set full` capture per kernel of interest at its config-2 shape (each after its own plain run exited 0).
# Outputs under gpurun_out/r02_prof/ ; summarised here by tools/summarize_profiles.py into profiles/r02_*.
set -u
O=gpurun_out/r02_prof; mkdir -p $O
python bench.py --profile --steps 1 --warmup 1 > $O/bench_profile_plain.json 2> $O/bench_profile_plain.err || exit 1
K='regex:attention_|gemm_bf16_tcgen05|gemv_|camera_pose|cast_f32_bf16|dino_embed|gather_rows|im2col|layernorm|mean_pool|mrope_table|points_epilogue|qknorm_mrope|rmsnorm_routed|rope2d|split3'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 4000 --csv --log-file $O/launches.csv \
    python bench.py --profile --steps 1 --warmup 1 > $O/ncu_bench.log 2>&1
for k in attention attention64 oproj dino_dense dino_fc1 gateup; do
  python tools/run_kernel.py $k 5 > $O/plain_$k.txt 2>&1 || continue
  case $k in attention*) pat='regex:attention_tcgen05';; *) pat='regex:gemm_bf16_tcgen05';; esac
  ncu --set full --clock-control none --import-source on -k "$pat" -s 3 -c 1 -f -o $O/$k python tools/run_kernel.py $k 2 > $O/ncu_$k.log 2>&1
done
cat $O/plain_*.txt
ls -la $O
