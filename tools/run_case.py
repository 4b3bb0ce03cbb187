#!/usr/bin/env python
"""Run one op on one shape a few times (for ncu / quick timing).
usage: python tools/run_case.py OP B C H W [dtype=f32] [path=auto] [reps=5]
OP in: fgelu_fwd fgelu_bwd fgelu_fwd_res up down up_bwd down_bwd gelu_down gelu_down_bwd"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import aliasfree_b200 as afr

op = sys.argv[1]; B, C, H, W = map(int, sys.argv[2:6])
dt = torch.bfloat16 if (len(sys.argv) > 6 and sys.argv[6] == "bf16") else torch.float32
path = sys.argv[7] if len(sys.argv) > 7 else "auto"
reps = int(sys.argv[8]) if len(sys.argv) > 8 else 5
afr.set_path(path)
NTAPS = int(os.environ.get("AFR_CASE_N", "3"))
k = afr.Taps(afr.circularLowpassKernel(np.pi / 2, NTAPS, 2))
x = torch.randn(B, C, H, W, device="cuda").to(dt)
dy = torch.randn_like(x); r = torch.randn_like(x)
if os.environ.get("AFR_CASE_CL") == "1":          # channels-last tensors (fgelu3_nhwc_kernel)
    x, dy, r = (t.contiguous(memory_format=torch.channels_last) for t in (x, dy, r))
big = torch.randn(B, C, 2 * H, 2 * W, device="cuda").to(dt) if op == "up_bwd" else None
small = torch.randn(B, C, H // 2, W // 2, device="cuda").to(dt) if op == "down_bwd" else None
fn = {"fgelu_fwd": lambda: afr.ops._fgelu_fwd(x, None, k, k), "fgelu_bwd": lambda: afr.ops._fgelu_bwd(x, None, dy, k, k),
      "fgelu_fwd_res": lambda: afr.ops._fgelu_fwd(x, r, k, k), "up": lambda: afr.ops._up_fwd(x, k, dt),
      "down": lambda: afr.ops._down_fwd(x, k), "up_bwd": lambda: afr.ops._up_bwd(big, k, H, W),
      "down_bwd": lambda: afr.ops._down_bwd(small, k, H, W),
      "gelu_down": lambda: afr.ops._GeluDown2x.apply(x, k),
      "gelu_down_bwd": lambda: afr.ops._GeluDown2x.apply(x.requires_grad_(True), k).backward(torch.ones(B, C, H // 2, W // 2, device="cuda").to(dt))}[op]
flush = torch.empty(64 << 20, dtype=torch.float32, device="cuda")
ts = []
for _ in range(reps):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
print(op, (B, C, H, W), str(dt), path, afr.last_kernel(), "ms:", " ".join(f"{t:.4f}" for t in ts))
