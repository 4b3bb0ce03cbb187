#!/usr/bin/env python
"""Generate golden fixtures by RUNNING THE REFERENCE ITSELF (CPU, fp32).

Run in the build container only:  python tests/golden/make_golden.py
It imports the unmodified reference from /root/reference (read-only); the two
plotting packages the reference imports at module top (matplotlib, imageio) are not
installed, so empty stand-ins are placed in sys.modules first.  Nothing of the
reference is copied: only inputs and the arrays it returns are stored.

Outputs (all small .npz files next to this script):
  taps.npz      circularLowpassKernel for a sweep of (omega_c, N, beta)
  resample.npz  custom_upsample / custom_downsample / up->gelu->down, fwd and autograd dx
  rotate.npz    Diffusion.rotate_2d_matrix (scipy.ndimage.rotate grid-wrap)
  blocks.npz    DoubleConv_F, Down_F/FF/FFF, Up_F/FF/FFF forward + input grads
  unet.npz      UNet variants 1/2/3 forward (+ variant 3 input grad)
  sampler.npz   Diffusion.sample short schedules (with and without theta), per-step states
  train.npz     the reference's own train() (modules/ddpm_utils.py:483-519), one epoch of three batches
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from _fill import fill_params_, seeded_array  # noqa: E402

REF = "/root/reference"


def import_reference():
    for name in ("matplotlib", "matplotlib.pyplot", "imageio"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.path.insert(0, REF)
    import modules.filtrs as filtrs
    import modules.ddpm_utils as ddpm_utils
    import modules.ddpm_models as ddpm_models
    return filtrs, ddpm_utils, ddpm_models


TAP_SWEEP = [  # (omega_c, N, beta)
    (np.pi / 2, 3, 2.0), (np.pi / 2, 3, None), (np.pi / 2, 3, 0.0), (np.pi / 2, 3, 1.0),
    (np.pi, 6, None), (np.pi / 2, 6, 2.0), (np.pi / 4, 11, None), (np.pi / 2, 4, 2.0),
    (np.pi / 2, 5, 1.0), (np.pi / 3, 7, 3.0), (np.pi / 2, 1, 2.0), (np.pi / 2, 2, None),
    (np.pi / 2, 8, 14.0), (0.8 * np.pi, 3, 2.0),
]

RESAMPLE_CASES = [  # (name, shape, (omega_up, N_up, beta_up), (omega_dn, N_dn, beta_dn))
    ("n3_8x8", (2, 3, 8, 8), (np.pi / 2, 3, 2.0), (np.pi / 2, 3, 2.0)),
    ("n3_odd", (1, 2, 7, 9), (np.pi / 2, 3, 2.0), (np.pi / 2, 3, 2.0)),
    ("n3_4x4", (2, 5, 4, 4), (np.pi / 2, 3, 2.0), (np.pi / 2, 3, 2.0)),
    ("n3_1x1", (1, 1, 1, 1), (np.pi / 2, 3, 2.0), (np.pi / 2, 3, 2.0)),
    ("n3_2x3", (1, 1, 2, 3), (np.pi / 2, 3, 2.0), (np.pi / 2, 3, 2.0)),
    ("n3_32x32", (1, 2, 32, 32), (np.pi / 2, 3, 2.0), (0.8 * np.pi, 3, 2.0)),
    ("n3_40x72", (1, 1, 40, 72), (np.pi / 2, 3, 1.0), (np.pi / 2, 3, None)),
    ("n6_5x6", (1, 1, 5, 6), (np.pi, 6, None), (np.pi, 6, None)),
    ("n4_6x5", (1, 2, 6, 5), (np.pi / 2, 4, 2.0), (np.pi / 2, 4, 2.0)),
    ("n5_8x8", (1, 2, 8, 8), (np.pi / 2, 5, 1.0), (np.pi / 2, 5, 1.0)),
    ("n8_9x9", (1, 1, 9, 9), (np.pi / 2, 8, 14.0), (np.pi / 2, 8, 14.0)),
    ("n6_16x16", (1, 2, 16, 16), (np.pi / 2, 6, 2.0), (np.pi / 2, 6, 2.0)),
    ("mixed_n3_n6", (1, 2, 12, 10), (np.pi / 2, 3, 2.0), (np.pi / 2, 6, 2.0)),
    ("n2_4x6", (1, 1, 4, 6), (np.pi / 2, 2, None), (np.pi / 2, 2, None)),
    ("n1_4x4", (1, 1, 4, 4), (np.pi / 2, 1, 2.0), (np.pi / 2, 1, 2.0)),
]

F_SETTINGS = dict(kernel_size=3, kaiser_beta=2, omega_c_down=np.pi / 2, omega_c_up=np.pi / 2)


def gen_taps(filtrs):
    out = {}
    for idx, (w, N, beta) in enumerate(TAP_SWEEP):
        out[f"k{idx}"] = filtrs.circularLowpassKernel(w, N, beta).numpy()
        out[f"p{idx}"] = np.array([w, N, -1.0 if beta is None else beta, beta is not None],
                                  np.float64)
    np.savez_compressed(os.path.join(HERE, "taps.npz"), **out)


def gen_resample(filtrs):
    import torch.nn.functional as F
    out, names = {}, []
    for name, shape, up, dn in RESAMPLE_CASES:
        names.append(name)
        ku, kd = filtrs.circularLowpassKernel(*up), filtrs.circularLowpassKernel(*dn)
        x = torch.from_numpy(seeded_array("x_" + name, shape)).requires_grad_(True)
        B, C, H, W = shape
        u = filtrs.custom_upsample(x, ku)
        du = torch.from_numpy(seeded_array("du_" + name, tuple(u.shape)))
        (dx_up,) = torch.autograd.grad(u, x, du)
        d = filtrs.custom_downsample(x, kd)
        dd = torch.from_numpy(seeded_array("dd_" + name, tuple(d.shape)))
        (dx_dn,) = torch.autograd.grad(d, x, dd)
        y = filtrs.custom_downsample(F.gelu(filtrs.custom_upsample(x, ku)), kd)
        dy = torch.from_numpy(seeded_array("dy_" + name, tuple(y.shape)))
        (dx_f,) = torch.autograd.grad(y, x, dy)
        for key, val in dict(ku=ku, kd=kd, x=x, up=u, du=du, dx_up=dx_up, down=d, dd=dd,
                             dx_down=dx_dn, fused=y, dy=dy, dx_fused=dx_f).items():
            out[f"{name}.{key}"] = val.detach().contiguous().numpy()
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "resample.npz"), **out)


def gen_rotate(models):
    out = {}
    x = torch.from_numpy(seeded_array("rot_x", (2, 3, 32, 32)))
    angles = [0.0225, 0.09, -0.05, 22.5, 45.0, -90.0, 180.0]
    out["x"] = x.numpy(); out["angles"] = np.array(angles)
    for i, a in enumerate(angles):
        out[f"y{i}"] = models.Diffusion.rotate_2d_matrix(x, a).numpy()
    xr = torch.from_numpy(seeded_array("rot_xr", (1, 2, 8, 12)))
    out["xr"] = xr.numpy()
    out["yr"] = models.Diffusion.rotate_2d_matrix(xr, 7.5).numpy()
    np.savez_compressed(os.path.join(HERE, "rotate.npz"), **out)


def gen_blocks(utils):
    out = {}
    fs = F_SETTINGS
    temb = torch.from_numpy(seeded_array("blk_t", (2, 256)))
    out["t"] = temb.numpy()

    def run(tag, mod, *inputs):
        fill_params_(mod, salt=tag)
        ins = [i.clone().requires_grad_(True) for i in inputs]
        y = mod(*ins, temb) if tag.startswith(("Down", "Up")) else mod(*ins)
        dy = torch.from_numpy(seeded_array("dy_" + tag, tuple(y.shape)))
        grads = torch.autograd.grad(y, ins, dy)
        out[f"{tag}.y"] = y.detach().numpy(); out[f"{tag}.dy"] = dy.numpy()
        for n, (i, g) in enumerate(zip(inputs, grads)):
            out[f"{tag}.in{n}"] = i.numpy(); out[f"{tag}.din{n}"] = g.numpy()

    x8 = torch.from_numpy(seeded_array("blk_x8", (2, 8, 8, 8)))
    x4 = torch.from_numpy(seeded_array("blk_x4", (2, 8, 4, 4)))
    sk = torch.from_numpy(seeded_array("blk_sk", (2, 8, 8, 8)))
    run("DoubleConv_F.plain", utils.DoubleConv_F(8, 12, f_settings=fs), x8)
    run("DoubleConv_F.mid", utils.DoubleConv_F(8, 6, 4, f_settings=fs), x8)
    run("DoubleConv_F.res", utils.DoubleConv_F(8, 8, residual=True, f_settings=fs), x8)
    run("Down_F", utils.Down_F(8, 16, f_settings=fs), x8)
    run("Down_FF", utils.Down_FF(8, 16, f_settings=fs), x8)
    run("Down_FFF", utils.Down_FFF(8, 16, f_settings=fs), x8)
    run("Up_F", utils.Up_F(16, 8, f_settings=fs), x4, sk)
    run("Up_FF", utils.Up_FF(16, 8, f_settings=fs), x4, sk)
    run("Up_FFF", utils.Up_FFF(16, 8, f_settings=fs), x4, sk)
    run("DoubleConv_F4.res", utils.DoubleConv_F4(8, 8, residual=True, f_settings=fs), x8)
    run("Down_F4", utils.Down_F4(8, 16, f_settings=fs), x8)
    run("Up_F4", utils.Up_F4(16, 8, f_settings=fs), x4, sk)
    np.savez_compressed(os.path.join(HERE, "blocks.npz"), **out)


def gen_unet(models):
    out = {}
    for variant, size, c in [(1, 16, 3), (2, 16, 3), (3, 16, 3), (3, 32, 1), (4, 16, 3)]:
        tag = f"v{variant}_s{size}_c{c}"
        net = models.UNet(c_in=c, c_out=c, image_size=size, device="cpu",
                          f_settings=F_SETTINGS, variant=variant)
        fill_params_(net)
        x = torch.from_numpy(seeded_array("unet_x_" + tag, (2, c, size, size)))
        t = torch.tensor([500, 37], dtype=torch.long)
        x.requires_grad_(True)
        y = net(x, t)
        out[f"{tag}.x"] = x.detach().numpy(); out[f"{tag}.t"] = t.numpy()
        out[f"{tag}.y"] = y.detach().numpy()
        out[f"{tag}.nparams"] = np.array(sum(p.numel() for p in net.parameters()))
        if variant == 3 and size == 16:
            dy = torch.from_numpy(seeded_array("unet_dy_" + tag, tuple(y.shape)))
            y.backward(dy)
            out[f"{tag}.dy"] = dy.numpy(); out[f"{tag}.dx"] = x.grad.numpy()
            # a few parameter grads pin the weight-gradient path through our backward
            for key in ("inc.conv1.weight", "down1.conv.0.conv2.weight", "bot2.norm1.bias",
                        "up3.conv.1.conv1.weight", "outc.bias"):
                out[f"{tag}.grad.{key}"] = dict(net.named_parameters())[key].grad.numpy()
    np.savez_compressed(os.path.join(HERE, "unet.npz"), **out)


def gen_sampler(models):
    out = {}
    for tag, variant, theta in [("v3_plain", 3, None), ("v3_theta", 3, 30.0), ("v1_plain", 1, None)]:
        net = models.UNet(c_in=3, c_out=3, image_size=16, device="cpu",
                          f_settings=F_SETTINGS, variant=variant)
        fill_params_(net)
        diff = models.Diffusion(noise_steps=12, img_size=16, device="cpu")
        states = []
        hook = net.register_forward_pre_hook(lambda m, args: states.append(args[0].clone()))
        torch.manual_seed(1234)
        x_u8, result_u8 = diff.sample(net, n=2, image_channels=3, theta=theta)
        hook.remove()
        out[f"{tag}.states"] = torch.stack(states).numpy()      # x fed to the model at each step
        out[f"{tag}.x_u8"] = x_u8.numpy(); out[f"{tag}.result_u8"] = result_u8.numpy()
        out[f"{tag}.theta"] = np.array(np.nan if theta is None else theta)
    np.savez_compressed(os.path.join(HERE, "sampler.npz"), **out)


class _TensorSample:
    """Mixin for the Diffusion handed to the reference's train(): ``sample`` returns the image tensor
    only.  Upstream ``sample`` returns a tuple which train() passes to torchvision's make_grid -- a
    TypeError at the end of epoch 0; the train() code itself runs unmodified."""

    def sample(self, *a, **kw):
        return super().sample(*a, **kw)[0]


TRAIN_CFG = dict(steps=3, batch=4, size=16, noise_steps=40, lr=3e-4, seed=77, image_gen_n=2)


def train_batches():
    c = TRAIN_CFG
    return [(torch.from_numpy(seeded_array(f"train_images_{i}", (c["batch"], 3, c["size"], c["size"]))).clamp(-1, 1),
             torch.zeros(c["batch"], dtype=torch.long)) for i in range(c["steps"])]


def gen_train(utils, models):
    """Run the reference's train() on CPU: timesteps, q-sample noise and the end-of-epoch sampling all
    draw from torch's default CPU generator (seeded), which the GPU test replays."""
    import tempfile
    c = TRAIN_CFG
    net = models.UNet(c_in=3, c_out=3, image_size=c["size"], device="cpu", f_settings=F_SETTINGS, variant=3)
    fill_params_(net)
    Diff = type("Diffusion", (_TensorSample, models.Diffusion), {})
    diff = Diff(noise_steps=c["noise_steps"], img_size=c["size"], device="cpu")
    args = utils.argument(run_name="golden_train", epochs=1, batch_size=c["batch"], image_size=c["size"],
                          image_channels=3, device="cpu", lr=c["lr"], noise_steps=c["noise_steps"],
                          image_gen_n=c["image_gen_n"])
    preds = []
    hook = net.register_forward_hook(lambda m, a, o: preds.append(o.detach().clone()) if m.training else None)
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            torch.manual_seed(c["seed"])
            losses = utils.train(args, os.path.join(tmp, "ckpt.pt"), train_batches(), net, diff)
            saved = sorted(os.listdir(os.path.join(tmp, "results", "golden_train")))
        finally:
            os.chdir(cwd)
    hook.remove()
    out = {"preds": torch.stack(preds).numpy(), "saved": np.array(len(saved))}
    if losses is not None:
        out["loss_all"] = np.asarray(losses, np.float64)
    sd = net.state_dict()
    for key in ("inc.conv1.weight", "down2.conv.1.norm2.weight", "bot2.norm2.bias", "up1.emb_layer.1.bias",
                "sa3.mha.in_proj_weight", "outc.weight"):
        out["param." + key] = sd[key].numpy()
    np.savez_compressed(os.path.join(HERE, "train.npz"), **out)


def main():
    torch.set_num_threads(8)
    torch.use_deterministic_algorithms(False)
    filtrs, utils, models = import_reference()
    gen_taps(filtrs)
    gen_resample(filtrs)
    gen_rotate(models)
    gen_blocks(utils)
    gen_unet(models)
    gen_sampler(models)
    gen_train(utils, models)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
