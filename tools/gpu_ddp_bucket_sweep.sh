#!/bin/bash
# 8 x B200 (gpurun --gpus 8): gradient all-reduce bucketing sweep of bvc_b200.DistributedDataParallel.
#   BVC_DDP_BUCKET_MB=0      one NCCL all-reduce per backward stage (19 per ViT-B step)
#   BVC_DDP_BUCKET_MB=<n>    consecutive stages coalesced into one NCCL group launch per >= n MB
# then the 2-rank hardware parity script under the coalesced mode.  Results under gpurun_out/.
N=${1:-8}
shift
mkdir -p gpurun_out
for mb in "$@"; do
  echo "=== BVC_DDP_BUCKET_MB=$mb" >&2
  BVC_DDP_BUCKET_MB=$mb timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port 29671 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu --no-hf-gpu \
    > gpurun_out/bench_${N}gpu_bucket${mb}.json 2> gpurun_out/bench_${N}gpu_bucket${mb}.err
  python - <<P
import json
try:
    d = json.load(open("gpurun_out/bench_${N}gpu_bucket${mb}.json"))
    print("bucket_mb=$mb value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1),
          "e2e_u8", round(d.get("e2e_uint8_input", {}).get("value", 0), 1), "clocks", d.get("clocks", {}).get("sm_mhz"))
except Exception as e:
    print("bucket_mb=$mb FAILED", e)
P
done
BVC_DDP_BUCKET_MB=96 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
  --master-port 29655 tools/ddp_parity.py --config base --batch 8 > gpurun_out/ddp_parity_bucket96.log 2>&1
grep -E "^(PASS|FAIL|DDP-PARITY)" gpurun_out/ddp_parity_bucket96.log | tail -25
