#!/usr/bin/env python
"""What do PyTorch-level switches do to one Config-D reverse step?  usage: python tools/step_experiment.py [batch]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import aliasfree_b200 as afr
FS = dict(kernel_size=3, kaiser_beta=2, omega_c_down=np.pi / 2, omega_c_up=np.pi / 2)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
net = afr.UNet(c_in=3, c_out=3, image_size=32, f_settings=FS, variant=3).cuda().eval()
diff = afr.Diffusion(noise_steps=1000, img_size=32, device="cuda")
x = torch.randn(n, 3, 32, 32, device="cuda")
def tm(tag, reps=4):
    with torch.no_grad():
        for _ in range(2): diff._reverse_step(net, x, 500, torch.randn_like(x))
        torch.cuda.synchronize(); t = time.perf_counter()
        for _ in range(reps): diff._reverse_step(net, x, 500, torch.randn_like(x))
        torch.cuda.synchronize()
    print(f"{tag:40s} {1e3 * (time.perf_counter() - t) / reps:8.2f} ms/step", flush=True)
tm("default")
torch.backends.cudnn.benchmark = True
tm("cudnn.benchmark=True")
torch.backends.cudnn.benchmark = False
torch.backends.cuda.matmul.allow_tf32 = True
tm("matmul tf32")
torch.backends.cuda.matmul.allow_tf32 = False
from torch.nn.attention import sdpa_kernel, SDPBackend
for be in (SDPBackend.EFFICIENT_ATTENTION, SDPBackend.MATH, SDPBackend.CUDNN_ATTENTION):
    try:
        with sdpa_kernel(be):
            tm("sdpa " + str(be), reps=2)
    except Exception as e:
        print("sdpa", be, "failed:", repr(e)[:100])
