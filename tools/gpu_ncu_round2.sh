set -x
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed
python bench.py --steps 2 --warmup 3 --no-cpu --no-hf-gpu > gpurun_out/plain_bench.log 2>&1 &&
timeout 1500 ncu --metrics $M --clock-control none -s 3600 -c 900 --csv --log-file gpurun_out/step_metrics.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-hf-gpu > gpurun_out/ncu_bench.log 2>&1
python tools/gpu_attn_probe.py 16 1568 6 3 > gpurun_out/plain_probe.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_bwd1 -s 2 -c 2 -o gpurun_out/prof_bwd1 python tools/gpu_attn_probe.py 16 1568 6 3 > gpurun_out/ncu_probe.log 2>&1
ls -la gpurun_out | tail; tail -3 gpurun_out/ncu_probe.log; cat gpurun_out/plain_probe.log
