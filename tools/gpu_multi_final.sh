#!/bin/bash
# 8 x B200 (gpurun --gpus 8): the round's multi-GPU evidence in one call.  Results under gpurun_out/.
mkdir -p gpurun_out
run() {  # name, nproc, extra bench args...
  local name=$1 n=$2; shift 2
  if [ "$n" = 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu --no-hf-gpu "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29671 \
      bench.py --gpus $n --steps 10 --warmup 3 --no-cpu --no-hf-gpu "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  fi
  python - <<P
import json
try:
    d = json.load(open("gpurun_out/$name.json"))
    print("$name value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1),
          "e2e_u8", round(d.get("e2e_uint8_input", {}).get("value", 0), 1), "clocks", d.get("clocks", {}).get("sm_mhz"),
          "numa", d["config"].get("numa_cpus_bound_rank0"))
except Exception as e:
    print("$name FAILED", e)
P
}
run bench_1gpu_samebox 1
run bench_8gpu 8
run bench_8gpu_vitl_b32 8 --config large --batch 32
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29655 \
  tools/ddp_parity.py --config base --batch 8 > gpurun_out/ddp_parity_8gpu.log 2>&1
grep -E "^(PASS|FAIL|DDP-PARITY)" gpurun_out/ddp_parity_8gpu.log | cut -c1-200 | tail -20
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
