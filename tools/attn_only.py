"""Run only the 64x64-level attention kernel (for ncu)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import time_attention_kernel
print(time_attention_kernel(32, torch.device('cuda'), iters=2))
