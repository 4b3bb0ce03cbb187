#!/usr/bin/env python
"""Eager (not graph-replayed) Config-D reverse step at a launch-bound batch: what does the TMA descriptor cache buy?
usage: [AFR_NO_DESC_CACHE=1] python tools/desc_cache_ab.py [batch] [steps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import aliasfree_b200 as afr
FS = dict(kernel_size=3, kaiser_beta=2, omega_c_down=np.pi / 2, omega_c_up=np.pi / 2)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
torch.manual_seed(0)
net = afr.UNet(c_in=3, c_out=3, image_size=32, f_settings=FS, variant=3).cuda().eval()
diff = afr.Diffusion(noise_steps=1000, img_size=32, device="cuda")
x = torch.randn(n, 3, 32, 32, device="cuda")
k = afr.Taps(afr.circularLowpassKernel(np.pi / 2, 3, 2))
a = torch.randn(n, 32, 32, 32, device="cuda")
with torch.no_grad():
    for _ in range(20):
        diff._reverse_step(net, x, 500, torch.randn_like(x))
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(steps):
        diff._reverse_step(net, x, 500, torch.randn_like(x))
    torch.cuda.synchronize(); step_ms = 1e3 * (time.perf_counter() - t) / steps
    # host cost of one fused launch (TMA path) with nothing else in the loop
    for _ in range(50):
        afr.ops._fgelu_fwd(a, None, k, k)
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(2000):
        afr.ops._fgelu_fwd(a, None, k, k)
    host_us = 1e6 * (time.perf_counter() - t) / 2000
    torch.cuda.synchronize()
print(f"desc cache {'OFF' if os.environ.get('AFR_NO_DESC_CACHE') == '1' else 'ON '}: eager reverse step n={n}: {step_ms:.3f} ms ;"
      f" one fused launch incl. output allocation, host side: {host_us:.2f} us")
