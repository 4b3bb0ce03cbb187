"""Experiment: Config-D reverse step under torch.autocast(bf16) vs fp32 (not a bench line)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import aliasfree_b200 as afr
FS = dict(kernel_size=3, kaiser_beta=2, omega_c_down=np.pi / 2, omega_c_up=np.pi / 2)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
net = afr.UNet(c_in=3, c_out=3, image_size=32, f_settings=FS, variant=3).cuda().eval()
x = torch.randn(B, 3, 32, 32, device="cuda"); t = torch.full((B,), 500, device="cuda")
def tm(fn, n=5):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - a) / n * 1e3
with torch.no_grad():
    y32 = net(x, t)
    ms32 = tm(lambda: net(x, t))
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y16 = net(x, t)
        ms16 = tm(lambda: net(x, t))
    torch.backends.cuda.matmul.allow_tf32 = True
    ms32tf = tm(lambda: net(x, t))
print(f"B={B} fp32 {ms32:.2f} ms  fp32+tf32 matmul {ms32tf:.2f} ms  bf16 autocast {ms16:.2f} ms  rel diff {float((y16.float()-y32).abs().max()/y32.abs().max()):.3e}")
