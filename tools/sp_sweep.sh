#!/bin/bash
# view-sharded K/V exchange A/B on N GPUs (builder tool): overlap (P2P + local-first attention + LSE merge) at several
# SM margins vs the v1 blocking all-gather, 64- and 256-view scenes.  usage: tools/sp_sweep.sh N out_prefix
N=${1:-2}; OUT=${2:-gpurun_out/r02_sp}; MARGINS=${3:-"0 8 24"}; VIEWS=${4:-"64 256"}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for V in $VIEWS; do
  for M in $MARGINS; do
    $T bench.py --gpus $N --workload views --views $V --steps 2 --warmup 1 --sp-mode overlap --sp-sm-margin $M 2>/dev/null | tail -1 > ${OUT}_n${N}_v${V}_overlap_m${M}.json
  done
  $T bench.py --gpus $N --workload views --views $V --steps 2 --warmup 1 --sp-mode allgather 2>/dev/null | tail -1 > ${OUT}_n${N}_v${V}_allgather.json
  $T bench.py --gpus $N --workload views --views $V --steps 2 --warmup 1 --sp-mode peer 2>/dev/null | tail -1 > ${OUT}_n${N}_v${V}_peer.json
done
for f in ${OUT}_n${N}_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1].split('/')[-1], "ms/scene %.1f  views/s %.2f  exposed exchange ms/layer %.4f  frac_of_peak %.3f" % (d["ms_per_step"], d["value"], d["exposed_exchange_ms_per_layer"], d["whole_step"]["frac_of_peak"]))
except Exception as e: print(sys.argv[1], "FAILED", e)
PY
done
