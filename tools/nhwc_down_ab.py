#!/usr/bin/env python
"""A/B of the channels-last down2x kernel's outputs-per-thread block (AFR_NHWC_DOWN = 11 | 12 | 21 | 22 | auto):
fraction of the measured HBM peak per shape and dtype, clean-L2 flush between launches; the up kernel beside it."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import aliasfree_b200 as afr

try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    PEAK = 6555.2
k = afr.Taps(afr.circularLowpassKernel(np.pi / 2, 3, 2))
flush = torch.empty(64 << 20, dtype=torch.float32, device="cuda")
flush_r = torch.zeros(64 << 20, dtype=torch.float32, device="cuda")
sink = torch.zeros((), device="cuda")


def tm(fn, reps=7):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_(); sink.add_(flush_r.sum())
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


SHAPES = [(4096, 256, 4, 4), (4096, 128, 8, 8), (4096, 64, 16, 16), (2048, 32, 32, 32), (256, 128, 64, 64), (64, 64, 128, 128),
          (512, 64, 16, 16), (512, 32, 32, 32), (16384, 64, 16, 16), (1024, 128, 64, 64)]
VARIANTS = ["11", "12", "21", "22", "auto"]
print("shape dtype | down2x: " + " ".join(f"{v:>6s}" for v in VARIANTS) + " | up2x | up2x-adjoint(auto)")
for dt, es, dn in ((torch.float32, 4, "f32"), (torch.bfloat16, 2, "bf16")):
    for shp in SHAPES:
        x = torch.randn(shp, device="cuda").to(dt).contiguous(memory_format=torch.channels_last); n = x.numel()
        fr = []
        for v in VARIANTS:
            if v == "auto":
                os.environ.pop("AFR_NHWC_DOWN", None)
            else:
                os.environ["AFR_NHWC_DOWN"] = v
            t = tm(lambda: afr.ops._down_fwd(x, k))
            assert afr.last_kernel() == "down3_nhwc_kernel"
            fr.append(1.25 * n * es / t / 1e6 / PEAK)
        os.environ.pop("AFR_NHWC_DOWN", None)
        u = tm(lambda: afr.ops._up_fwd(x, k, dt))
        B, C, H, W = shp
        ub = tm(lambda: afr.ops._up_bwd(x, k, H // 2, W // 2))       # x as the gradient of a [B,C,H/2,W/2] input
        print(f"{str(list(shp)):22s} {dn:5s} | " + " ".join(f"{f:6.3f}" for f in fr) +
              f" | {5 * n * es / u / 1e6 / PEAK:6.3f} | {1.25 * n * es / ub / 1e6 / PEAK:6.3f}")
        del x
