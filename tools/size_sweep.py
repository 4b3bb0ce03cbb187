#!/usr/bin/env python
"""t(bytes) = a + bytes / bw for the standalone resamplers: separates the fixed cost of a launch measured between two
events (launch latency, ramp, tail, cold TLB after the L2 flush) from the streaming rate.  Prints per layout / dtype /
op the least-squares a (us) and bw (fraction of the measured HBM peak) over a batch sweep of [B, 64, 16, 16]."""
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


def tm(fn, reps=9):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_(); sink.add_(flush_r.sum())
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)) * 1e3          # us


C, H, W = int(os.environ.get("SWEEP_C", 64)), int(os.environ.get("SWEEP_H", 16)), int(os.environ.get("SWEEP_W", 16))
BS = [256, 512, 1024, 2048, 4096, 8192, 16384, 32768]
tiny = torch.randn(1, 4, 2, 2, device="cuda")
print(f"floor: down2x on [1,4,2,2] between two events: {tm(lambda: afr.ops._down_fwd(tiny, k)):.2f} us")
for layout in ("nchw", "nhwc"):
    for dt, es, dn in ((torch.float32, 4, "f32"), (torch.bfloat16, 2, "bf16")):
        for op, per in (("down", 1.25), ("up", 5.0)):
            xs, ys, names = [], [], set()
            for B in BS:
                if op == "up" and B > 16384:
                    continue
                x = torch.randn(B, C, H, W, device="cuda").to(dt)
                if layout == "nhwc":
                    x = x.contiguous(memory_format=torch.channels_last)
                fn = (lambda: afr.ops._down_fwd(x, k)) if op == "down" else (lambda: afr.ops._up_fwd(x, k, dt))
                t = tm(fn); names.add(afr.last_kernel())
                xs.append(per * x.numel() * es); ys.append(t)
                del x
            A = np.stack([np.ones(len(xs)), np.array(xs)], 1)
            (a, b), *_ = np.linalg.lstsq(A, np.array(ys), rcond=None)
            pts = " ".join(f"{B}:{t:.1f}" for B, t in zip(BS, ys))
            print(f"{layout} {dn:5s} {op:5s} a = {a:6.2f} us  bw = {1 / b / 1e3 / PEAK:5.3f} of peak   [{pts}]  {sorted(names)}")
