set -x
python tools/gpu_attn_probe.py 16 1568 6 3 > gpurun_out/plain_probe_1568.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_kernel -s 2 -c 1 -o gpurun_out/prof_fwd python tools/gpu_attn_probe.py 16 1568 6 3 > gpurun_out/ncu_fwd.log 2>&1
python tools/gpu_attn_probe.py 64 160 12 3 > gpurun_out/plain_probe_160.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_small -s 4 -c 2 -o gpurun_out/prof_small python tools/gpu_attn_probe.py 64 160 12 3 > gpurun_out/ncu_small.log 2>&1
cat gpurun_out/plain_probe_160.log; tail -2 gpurun_out/ncu_small.log
