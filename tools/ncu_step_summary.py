#!/usr/bin/env python
"""Summarise an ncu launch list of bench.py into per-kernel-family counters for ONE training step.

    python tools/ncu_step_summary.py gpurun_out/step_metrics.csv profiles/r01_ncu_step_metrics.json

Input: `ncu --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,
gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.
avg.pct_of_peak_sustained_elapsed ... python bench.py --steps 2 --warmup 3 --no-cpu`.  One step = the launches
between two consecutive patchify_target kernels.  ncu's times are cold-cache and serialised: the SHARES are the evidence,
the absolute step time comes from bench.py's CUDA events."""
import csv
import json
import re
import sys
from collections import OrderedDict


def family(name):
    for pat, fam in (("gemm_kernel", "gemm"), ("attn(_small)?_fwd", "attn_fwd"), ("attn(_small)?_bwd|attn_dq_convert", "attn_bwd"),
                     ("attn_delta", "attn_delta"),
                     ("patchify_target", "patchify_target"), ("mask_to_index|mask_count|mask_", "mask"),
                     ("layernorm_fwd", "layernorm_fwd"), ("layernorm_bwd", "layernorm_bwd"), ("colsum", "colsum"),
                     ("sgd", "sgd_step"), ("decoder_mask_rows", "decoder_mask_rows"), ("rows_to_bf16", "rows_to_bf16"),
                     ("loss_finalize", "loss_finalize"), ("cast", "cast")):
        if re.search(pat, name):
            return fam
    return "other (torch / NCCL)"


def main(src, dst):
    rows = []
    with open(src, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    by_id = OrderedDict()
    for r in rd:
        if "ID" not in r or not r.get("Metric Name"):
            continue
        e = by_id.setdefault(r["ID"], {"name": r["Kernel Name"]})
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        unit = r.get("Metric Unit", "")
        m = r["Metric Name"]
        if m == "gpu__time_duration.sum":
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)      # -> us
        if m.startswith("dram__bytes"):
            v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)  # -> bytes
        e[m] = v
    rows = list(by_id.values())
    marks = [i for i, r in enumerate(rows) if "patchify_target" in r["name"]]
    if len(marks) < 2:
        raise SystemExit(f"need two patchify_target launches to delimit a step, found {len(marks)} in {len(rows)} launches")
    step = rows[marks[0]:marks[1]]
    fams = OrderedDict()
    T, RD, WR = "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum"
    DP = "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"
    TP = "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed"
    for r in step:
        f_ = fams.setdefault(family(r["name"]), {"launches": 0, "time_us": 0.0, "dram_read": 0.0, "dram_write": 0.0,
                                                 "dram_pct_x_time": 0.0, "tensor_pct_x_time": 0.0})
        t = r.get(T, 0.0)
        f_["launches"] += 1
        f_["time_us"] += t
        f_["dram_read"] += r.get(RD, 0.0)
        f_["dram_write"] += r.get(WR, 0.0)
        f_["dram_pct_x_time"] += r.get(DP, 0.0) * t
        f_["tensor_pct_x_time"] += r.get(TP, 0.0) * t
    total = sum(f_["time_us"] for f_ in fams.values())
    out = {"source": src, "launches_in_step": len(step), "ncu_step_time_us": total, "families": OrderedDict()}
    for k, f_ in sorted(fams.items(), key=lambda kv: -kv[1]["time_us"]):
        t = max(f_["time_us"], 1e-9)
        out["families"][k] = {
            "launches": f_["launches"], "time_us": round(f_["time_us"], 1), "share_of_step": round(f_["time_us"] / total, 4),
            "dram_bytes_per_launch": round((f_["dram_read"] + f_["dram_write"]) / f_["launches"]),
            "dram_read_bytes": round(f_["dram_read"]), "dram_write_bytes": round(f_["dram_write"]),
            "achieved_dram_GBps": round((f_["dram_read"] + f_["dram_write"]) / t / 1e3, 1),
            "dram_throughput_pct_time_weighted": round(f_["dram_pct_x_time"] / t, 2),
            "tensor_pipe_utchmma_pct_time_weighted": round(f_["tensor_pct_x_time"] / t, 2),
        }
    with open(dst, "w") as f:
        json.dump(out, f, indent=1)
    for k, v in out["families"].items():
        print(f"{k:24s} n={v['launches']:4d} {v['time_us']:9.1f} us {100*v['share_of_step']:5.1f}%  dram {v['achieved_dram_GBps']:7.1f} GB/s "
              f"({v['dram_throughput_pct_time_weighted']:5.1f}%)  tensor {v['tensor_pipe_utchmma_pct_time_weighted']:5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
