#!/usr/bin/env python
"""Compact timing table for A/B builds: fraction of the measured HBM peak per (op, shape, dtype).
usage: [AFR_LIB_PATH=tools/variants/libafr_X.so] [AFR_...=..] python tools/ab_sweep.py [fused|resample|all|fused_cl|resample_cl] [tag]
(fused_cl / resample_cl: the same ops on channels-last tensors)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import aliasfree_b200 as afr

what = sys.argv[1] if len(sys.argv) > 1 else "all"
tag = sys.argv[2] if len(sys.argv) > 2 else os.environ.get("AFR_LIB_PATH", "default")
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    PEAK = 6555.2
k = afr.Taps(afr.circularLowpassKernel(np.pi / 2, 3, 2))
flush = torch.empty(64 << 20, dtype=torch.float32, device="cuda")
flush_r = torch.zeros(64 << 20, dtype=torch.float32, device="cuda")
CLEAN = os.environ.get("AB_CLEAN_FLUSH", "1") != "0"     # read pass after the write pass: L2 is left full of CLEAN lines
sink = torch.zeros((), device="cuda")


def tm(fn, reps=7):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        if CLEAN:
            sink.add_(flush_r.sum())
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


FUSED = [(256, 128, 64, 64), (1024, 64, 32, 32), (64, 64, 128, 128), (16, 64, 256, 256), (4096, 64, 16, 16), (4096, 128, 8, 8),
         (4096, 256, 4, 4)]
RES = [(4096, 256, 4, 4), (4096, 128, 8, 8), (4096, 64, 16, 16), (2048, 32, 32, 32), (256, 128, 64, 64), (64, 64, 128, 128),
       (16, 64, 256, 256)]
rows = []
for dt, es, dn in ((torch.float32, 4, "f32"), (torch.bfloat16, 2, "bf16")):
    if what in ("fused", "all", "fused_cl"):
        for shp in FUSED:
            x = torch.randn(shp, device="cuda").to(dt); dy = torch.randn_like(x); n = x.numel()
            if what == "fused_cl":
                x = x.contiguous(memory_format=torch.channels_last); dy = dy.contiguous(memory_format=torch.channels_last)
            f = tm(lambda: afr.ops._fgelu_fwd(x, None, k, k)); kf = afr.last_kernel()
            b = tm(lambda: afr.ops._fgelu_bwd(x, None, dy, k, k))
            rows.append(("fgelu", shp, dn, 2 * n * es / f / 1e6 / PEAK, 3 * n * es / b / 1e6 / PEAK, kf))
            del x, dy
    if what in ("resample", "all", "resample_cl"):
        for shp in RES:
            x = torch.randn(shp, device="cuda").to(dt); n = x.numel()
            if what == "resample_cl":
                x = x.contiguous(memory_format=torch.channels_last)
            u = tm(lambda: afr.ops._up_fwd(x, k, dt)); ku = afr.last_kernel()
            d = tm(lambda: afr.ops._down_fwd(x, k)); kd = afr.last_kernel()
            rows.append(("up/down", shp, dn, 5 * n * es / u / 1e6 / PEAK, 1.25 * n * es / d / 1e6 / PEAK, ku + " / " + kd))
            del x
print(f"== {tag}  env: " + " ".join(f"{a}={b}" for a, b in os.environ.items() if a.startswith("AFR_")))
for op, shp, dn, a, b, kn in rows:
    print(f"{op:8s} {str(list(shp)):22s} {dn:5s} {a:6.3f} {b:6.3f}   {kn}")
