"""Profiling driver (run under ncu): the bench workload (base Unet dim=128 64x64, bs=16, cond_scale 3 -> 32 U-Net rows),
a few EAGER denoising steps (no CUDA graph, so every kernel is a separate launch ncu can see)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ['B200_IMAGEN_NO_GRAPH'] = '1'
from bench import build_model  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
bs = int(sys.argv[2]) if len(sys.argv) > 2 else 16
im = build_model(128, steps, torch.device('cuda'))
torch.manual_seed(0)
out = im.sample(text_embeds=torch.randn(bs, 256, 768, device='cuda'), cond_scale=3., use_tqdm=False)
torch.cuda.synchronize()
print('ok', out.shape, im.last_launch_count)
import json
plan = next(iter(im.unets[0]._plans.values()))
os.makedirs('gpurun_out', exist_ok=True)
json.dump(plan.describe_gemms(), open('gpurun_out/plan_gemms.json', 'w'))
