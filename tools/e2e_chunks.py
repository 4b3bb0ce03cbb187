#!/usr/bin/env python
"""Sensitivity of the host-tensor pipeline (HostPipeline.filtered_gelu) to chunk count and slots."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import aliasfree_b200 as afr
B, C, H, W = 256, 128, 64, 64
hx = torch.randn(B, C, H, W).pin_memory(); hy = torch.empty(B, C, H, W).pin_memory()
k = afr.circularLowpassKernel(np.pi / 2, 3, 2)
for chunks in (8, 16, 32, 64):
    for slots in (2, 3, 4):
        pipe = afr.HostPipeline((B // chunks, C, H, W), slots=slots)
        for _ in range(2): pipe.filtered_gelu(hx, hy, k, k)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5): pipe.filtered_gelu(hx, hy, k, k)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        print(f"chunks {chunks:3d} slots {slots}: {ms:7.3f} ms  {2 * hx.numel() * 4 / ms / 1e6:6.1f} GB/s")
