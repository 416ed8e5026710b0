#!/usr/bin/env python
"""Verbose bring-up of the full model on the GPU: loss / logits / per-tensor gradient errors vs the CPU oracle."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import videomae_oracle as O  # noqa: E402
from tests.helpers import grad_report, rel_l2, run_bvc  # noqa: E402

torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
for name, batch, ratio in (("tiny", 3, 0.5), ("small", 2, 0.9)) if len(sys.argv) < 2 else [(sys.argv[1], int(sys.argv[2]), 0.9)]:
    for perturb in (False, True):
        cfg = O.make_config(name)
        params = O.init_params(cfg, seed=0, perturb=perturb)
        x = O.synthetic_clip(batch, cfg, seed=0, image_like=perturb)
        np.random.seed(0)
        mask = O.batch_tube_masks(batch, cfg.grid, ratio)
        ref_loss, ref_logits, ref_grads = O.grads_of(params, x, mask, cfg)
        loss, logits, grads, _ = run_bvc(cfg, params, x, mask)
        rows, g_all = grad_report(grads, ref_grads)
        print(f"== {name} B{batch} perturb={perturb}: loss {float(loss):.7f} ref {float(ref_loss):.7f} "
              f"rel {abs(float(loss)-float(ref_loss))/float(ref_loss):.2e}; logits rel-L2 {rel_l2(logits, ref_logits):.2e}; "
              f"grads global rel-L2 {g_all:.2e}", flush=True)
        for k, (e, ne, n) in sorted(rows.items(), key=lambda kv: -kv[1][0])[:12]:
            print(f"   {k:75s} rel-L2 {e:.2e} norm-rel {ne:.2e} |ref| {n:.3e}")
