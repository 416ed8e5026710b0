#!/usr/bin/env python
"""Secondary bar: the unmodified HF VideoMAEForPreTraining on the same B200 under
torch.autocast(bf16) with torch's library kernels (cuDNN conv3d, cuBLASLt, SDPA) -- same step as bench.py."""
import json
import os
import sys
import time

import numpy as np
import torch
import transformers

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps, warm = 8, 3
dev = torch.device("cuda:0")
c = bench.CONFIGS["base"]
torch.manual_seed(0)
model = transformers.VideoMAEForPreTraining(transformers.VideoMAEConfig(
    image_size=224, patch_size=16, num_channels=3, num_frames=16, tubelet_size=2, initializer_range=0.02,
    use_mean_pooling=True, norm_pix_loss=True, **c)).to(dev).train()
opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9, nesterov=True)
scaler = torch.amp.GradScaler("cuda")
x = [torch.randn(B, 16, 3, 224, 224, device=dev) for _ in range(2)]
masks = [bench.make_masks(B, i).to(dev) for i in range(4)]


def step(i):
    with torch.autocast("cuda", dtype=torch.bfloat16):
        opt.zero_grad()
        loss = model(x[i % 2], bool_masked_pos=masks[i % 4]).loss
    scaler.scale(loss).backward()
    scaler.step(opt)
    scaler.update()
    return loss


for i in range(warm):
    step(i)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for i in range(steps):
    loss = step(i)
e.record()
torch.cuda.synchronize()
ms = s.elapsed_time(e) / steps
print(json.dumps({"impl": "hf_gpu_bf16_autocast", "batch": B, "ms_per_step": ms, "clips_per_s": B / ms * 1e3,
                  "loss": float(loss.detach()), "torch": torch.__version__, "transformers": transformers.__version__,
                  "max_mem_gb": torch.cuda.max_memory_allocated() / 2**30}))
