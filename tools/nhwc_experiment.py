#!/usr/bin/env python
"""Config-D UNet in channels-last vs NCHW: reverse step (eager / graph replay) and training step.
usage: python tools/nhwc_experiment.py [batch] [train_batch]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import aliasfree_b200 as afr
from aliasfree_b200 import parallel
FS = dict(kernel_size=3, kaiser_beta=2, omega_c_down=np.pi / 2, omega_c_up=np.pi / 2)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
nt = int(sys.argv[2]) if len(sys.argv) > 2 else 256
torch.manual_seed(0)
base = afr.UNet(c_in=3, c_out=3, image_size=32, f_settings=FS, variant=3).cuda()
x0 = torch.randn(n, 3, 32, 32, device="cuda")
t0 = torch.full((n,), 500, dtype=torch.long, device="cuda")
outs = {}
for fmt in ("nchw", "channels_last"):
    import copy
    net = copy.deepcopy(base).eval()
    if fmt == "channels_last":
        net = net.to(memory_format=torch.channels_last)
    diff = afr.Diffusion(noise_steps=1000, img_size=32, device="cuda")
    x = x0.clone()
    with torch.no_grad():
        outs[fmt] = net(x0, t0).float()
        l0 = afr.launch_count()
        for _ in range(2): diff._reverse_step(net, x, 500, torch.randn_like(x))
        torch.cuda.synchronize(); t = time.perf_counter()
        for _ in range(4): diff._reverse_step(net, x, 500, torch.randn_like(x))
        torch.cuda.synchronize(); eager = 1e3 * (time.perf_counter() - t) / 4
        launches = (afr.launch_count() - l0) // 6
        g, step = diff.capture_reverse_step(net, x)
        for _ in range(2): g.replay()
        torch.cuda.synchronize(); t = time.perf_counter()
        for _ in range(6): g.replay()
        torch.cuda.synchronize(); graph = 1e3 * (time.perf_counter() - t) / 6
    print(f"{fmt:14s} reverse step n={n}: eager {eager:8.2f} ms  graph {graph:8.2f} ms  afr launches/step {launches}  last kernel {afr.last_kernel()}", flush=True)
    del g
    # training step
    tn = copy.deepcopy(base).train()
    if fmt == "channels_last":
        tn = tn.to(memory_format=torch.channels_last)
    opt = torch.optim.AdamW(tn.parameters(), lr=3e-4, capturable=True, fused=True)
    ddp = parallel.FlatGradAllReduce(tn)
    imgs = torch.rand(nt, 3, 32, 32, device="cuda") * 2 - 1
    gs = parallel.GraphedTrainStep(tn, diff, opt, tuple(imgs.shape), ddp=ddp)
    for _ in range(3): gs(imgs)
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(10): gs(imgs)
    torch.cuda.synchronize()
    print(f"{fmt:14s} train step batch {nt}: graphed {1e3 * (time.perf_counter() - t) / 10:8.2f} ms  loss {float(gs.loss):.4f}", flush=True)
    del gs, tn, opt, ddp
    torch.cuda.empty_cache()
d = (outs["nchw"] - outs["channels_last"]).abs().max() / outs["nchw"].abs().max()
print("UNet output rel-max difference channels_last vs nchw (TF32 convs):", float(d))
