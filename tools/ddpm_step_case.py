"""A few eager Config-D reverse steps (for ncu launch lists).  usage: python tools/ddpm_step_case.py [batch] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import aliasfree_b200 as afr
FS = dict(kernel_size=3, kaiser_beta=2, omega_c_down=np.pi / 2, omega_c_up=np.pi / 2)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
torch.manual_seed(0)
net = afr.UNet(c_in=3, c_out=3, image_size=32, f_settings=FS, variant=3).cuda().eval()
if os.environ.get("AFR_CASE_CL") == "1":          # channels-last model
    net = net.to(memory_format=torch.channels_last)
diff = afr.Diffusion(noise_steps=1000, img_size=32, device="cuda")
x = torch.randn(B, 3, 32, 32, device="cuda")
with torch.no_grad():
    for i in range(steps):
        diff._reverse_step(net, x, 500 - i, torch.randn_like(x))
torch.cuda.synchronize()
print("ok", float(x.abs().mean()))
