"""A few eager Config-D training steps (for ncu launch lists).  usage: python tools/train_step_case.py [batch] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import aliasfree_b200 as afr
from aliasfree_b200 import parallel
FS = dict(kernel_size=3, kaiser_beta=2, omega_c_down=np.pi / 2, omega_c_up=np.pi / 2)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
torch.manual_seed(0)
net = afr.UNet(c_in=3, c_out=3, image_size=32, f_settings=FS, variant=3).cuda().train()
diff = afr.Diffusion(noise_steps=1000, img_size=32, device="cuda")
opt = torch.optim.AdamW(net.parameters(), lr=3e-4)
ddp = parallel.FlatGradAllReduce(net)
imgs = torch.rand(B, 3, 32, 32, device="cuda") * 2 - 1
for _ in range(steps):
    loss = parallel.train_step(net, diff, opt, imgs, ddp=ddp)
torch.cuda.synchronize()
print("loss", float(loss))
