#!/usr/bin/env python
"""Roofline model of one training step, per kernel shape: measured time (bench.py --detail) against the time the same
algorithmic work would take at the measured peaks (MEASURED_PEAKS.json: sustained bf16 tensor rate, copy bandwidth).

    python tools/step_model.py profiles/r01_kernel_shapes_v6.json > profiles/r01_step_model_v6.md

`rate` in the detail file is TFLOP/s for GEMM / attention shapes and TB/s for the memory-bound ones (algorithmic FLOPs
or bytes of the launch, the same accounting as bench.py's roofline), so floor = measured * rate / peak."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main(path):
    peaks = {"tc": 1387.9, "hbm": 6.5552}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            mp = json.load(f)
        peaks["tc"] = float(mp.get("bf16_tflops_sustained", peaks["tc"]))
        peaks["hbm"] = float(mp.get("hbm_gbs", peaks["hbm"] * 1e3)) / 1e3
    except (OSError, ValueError):
        pass
    rows = json.load(open(path))
    rows.sort(key=lambda r: -r["ms_per_step"])
    total = sum(r["ms_per_step"] for r in rows)
    floor_total = 0.0
    print(f"# Step model from `{os.path.relpath(path, ROOT)}` (peaks: {peaks['tc']:.0f} TFLOP/s sustained bf16, "
          f"{peaks['hbm']:.2f} TB/s copy)\n")
    print("| kernel shape | launches | measured ms/step | achieved | bound | floor ms/step | gap ms | of peak |")
    print("|---|---|---|---|---|---|---|---|")
    for r in rows:
        tensor = r["what"].startswith(("gemm", "attn_fwd", "attn_bwd"))
        peak = peaks["tc"] if tensor else peaks["hbm"]
        frac = r["rate"] / peak if peak else 0.0
        floor = r["ms_per_step"] * frac
        floor_total += floor
        unit = "TFLOP/s" if tensor else "TB/s"
        print(f"| {r['what']} | {r['launches_per_step']:.0f} | {r['ms_per_step']:.3f} | {r['rate']:.1f} {unit} | "
              f"{'tensor' if tensor else 'hbm'} | {floor:.3f} | {r['ms_per_step'] - floor:.3f} | {100 * frac:.0f} % |")
    print(f"\nSum of kernel time {total:.2f} ms per step; the same work at the measured peaks {floor_total:.2f} ms "
          f"({100 * floor_total / total:.0f} %).  The gap column ranks where the next milliseconds are.")


if __name__ == "__main__":
    main(sys.argv[1])
