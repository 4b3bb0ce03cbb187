"""The reference's OWN UNet, sampler and train loop running unchanged on the CUDA kernels.

``afr.patch()`` rebinds the hot-path names inside the unmodified reference modules (imported from
``/root/reference`` in the build container, from the byte-for-byte install ``baseline/_ref`` on the
GPU box -- see tools/install_ref.py); everything else that executes below is reference code:
``modules.ddpm_models.UNet`` (:41-298), ``Diffusion.sample`` (:352-386) and ``modules.ddpm_utils.train``
(:483-519).  Results are compared with fixtures the same reference produced on CPU
(tests/golden/make_golden.py); random draws are replayed from torch's CPU generator so that both
sides see identical seeds.
"""
import os
import sys
from unittest import mock

import numpy as np
import pytest
import torch

from _fill import fill_params_
from conftest import ROOT, golden, relmax

sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

pytestmark = pytest.mark.gpu

FS = dict(kernel_size=3, kaiser_beta=2, omega_c_down=np.pi / 2, omega_c_up=np.pi / 2)
UNET_TOL = 3e-4      # cuDNN / cuBLAS summation order vs the CPU reference through ~60 layers (see test_gpu_models)


@pytest.fixture(scope="module")
def afr():
    import aliasfree_b200 as m
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return m


@pytest.fixture()
def ref(afr):
    """(filtrs, ddpm_utils, ddpm_models) of the unmodified reference with our path patched in."""
    from baseline import ref_loader
    if not ref_loader.available():
        pytest.skip("no reference checkout and no baseline/_ref install (run tools/install_ref.py)")
    mods = ref_loader.load()
    names = afr.patch()
    assert "modules.ddpm_models.DoubleConv_F" in names and "modules.ddpm_utils.custom_upsample" in names
    try:
        yield mods
    finally:
        afr.unpatch()


def dev(a, grad=False):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t.requires_grad_(True) if grad else t


def host(t):
    return t.detach().float().cpu().numpy()


def _cpu_randn_like(x, **kw):
    """torch.randn_like drawing from the default CPU generator, like the reference run that made the
    fixtures (its tensors lived on the CPU)."""
    return torch.randn(x.shape, dtype=x.dtype).to(x.device)


@pytest.mark.parametrize("variant,size,c", [(1, 16, 3), (2, 16, 3), (3, 16, 3), (3, 32, 1), (4, 16, 3)])
def test_reference_unet_runs_on_our_kernels(afr, ref, variant, size, c):
    """modules.ddpm_models.UNet forward (+ backward for the variant-3 fixture) under patch()."""
    _, _, rm = ref
    tag = f"v{variant}_s{size}_c{c}"
    g = golden("unet.npz")
    net = rm.UNet(c_in=c, c_out=c, image_size=size, device="cuda", f_settings=FS, variant=variant)
    assert type(net).__module__ == "modules.ddpm_models"             # the reference's class, not ours
    if variant in (2, 3):
        assert isinstance(net.inc, afr.DoubleConv_F)
    if variant in (1, 3):
        assert isinstance(net.down1, (afr.Down_FF, afr.Down_FFF))
    net = fill_params_(net).cuda()
    l0 = afr.launch_count()
    x = dev(g[f"{tag}.x"], grad=True)
    y = net(x, dev(g[f"{tag}.t"]))
    assert afr.launch_count() > l0                                    # our kernels did run
    assert relmax(host(y), g[f"{tag}.y"]) <= UNET_TOL
    if f"{tag}.dx" in g.files:
        y.backward(dev(g[f"{tag}.dy"]))
        assert relmax(host(x.grad), g[f"{tag}.dx"]) <= UNET_TOL
        params = dict(net.named_parameters())
        for key in [k for k in g.files if k.startswith(f"{tag}.grad.")]:
            assert relmax(host(params[key[len(tag) + 6:]].grad), g[key]) <= UNET_TOL, key


@pytest.mark.parametrize("tag,variant", [("v3_plain", 3), ("v3_theta", 3), ("v1_plain", 1)])
def test_reference_sampler_runs_on_our_kernels(afr, ref, tag, variant):
    """modules.ddpm_models.Diffusion.sample (Algorithm 1, with the Config-E rotation for v3_theta:
    its rotate_2d_matrix is rebound to the device kernel) against the reference's CPU run."""
    _, _, rm = ref
    g = golden("sampler.npz")
    theta = None if np.isnan(g[f"{tag}.theta"]) else float(g[f"{tag}.theta"])
    net = fill_params_(rm.UNet(c_in=3, c_out=3, image_size=16, device="cuda", f_settings=FS, variant=variant)).cuda()
    diff = rm.Diffusion(noise_steps=12, img_size=16, device="cuda")
    states = []
    hook = net.register_forward_pre_hook(lambda m, a: states.append(a[0].detach().clone()))
    torch.manual_seed(1234)
    with mock.patch.object(torch, "randn_like", _cpu_randn_like):
        x_u8, result_u8 = diff.sample(net, n=2, image_channels=3, theta=theta)
    hook.remove()
    want = g[f"{tag}.states"]
    assert len(states) == want.shape[0] == 11
    for s, w in zip(states, want):
        assert relmax(host(s), w) <= 1e-3          # 11 chained UNet calls; per-call error is ~1e-5
    assert x_u8.is_cuda and x_u8.dtype == torch.uint8
    assert tuple(result_u8.shape) == g[f"{tag}.result_u8"].shape
    assert np.abs(x_u8.cpu().numpy().astype(int) - g[f"{tag}.x_u8"].astype(int)).max() <= 1


def test_reference_train_loop_runs_on_our_kernels(afr, ref, tmp_path, monkeypatch):
    """modules.ddpm_utils.train -- setup, AdamW, the inner step (:498-509), the end-of-epoch sampling,
    image and checkpoint writes -- for one epoch of three batches, against the same call on CPU."""
    import make_golden as mg
    _, ru, rm = ref
    c = mg.TRAIN_CFG
    g = golden("train.npz")
    net = fill_params_(rm.UNet(c_in=3, c_out=3, image_size=c["size"], device="cuda", f_settings=FS, variant=3)).cuda()
    Diff = type("Diffusion", (mg._TensorSample, rm.Diffusion), {})
    diff = Diff(noise_steps=c["noise_steps"], img_size=c["size"], device="cuda")
    args = ru.argument(run_name="golden_train", epochs=1, batch_size=c["batch"], image_size=c["size"],
                       image_channels=3, device="cuda", lr=c["lr"], noise_steps=c["noise_steps"],
                       image_gen_n=c["image_gen_n"])
    preds = []
    hook = net.register_forward_hook(lambda m, a, o: preds.append(o.detach().clone()) if m.training else None)
    monkeypatch.chdir(tmp_path)
    l0 = afr.launch_count()
    torch.manual_seed(c["seed"])
    with mock.patch.object(torch, "randn_like", _cpu_randn_like):
        losses = ru.train(args, str(tmp_path / "ckpt.pt"), mg.train_batches(), net, diff)
    hook.remove()
    assert afr.launch_count() - l0 > 3 * 2 * 22                      # fwd + bwd kernels of three steps at least
    assert len(preds) == c["steps"]
    # step 0 sees identical parameters; later steps sit behind AdamW updates (lr * m / (sqrt(v) + eps) is a
    # sign function on its first steps, so last-bit gradient differences move a few weights by up to 2 lr)
    assert relmax(host(preds[0]), g["preds"][0]) <= UNET_TOL
    for k in range(1, c["steps"]):
        assert relmax(host(preds[k]), g["preds"][k]) <= 5e-3, k
    assert abs(losses[0] - float(g["loss_all"][0])) <= 1e-3 * float(g["loss_all"][0])
    sd = net.state_dict()
    for key in [k for k in g.files if k.startswith("param.")]:
        assert float(np.abs(host(sd[key[6:]]) - g[key]).max()) <= 2.5 * c["lr"] * c["steps"], key
    assert len(os.listdir(tmp_path / "results" / "golden_train")) == int(g["saved"])
    assert os.path.exists(tmp_path / "ckpt.pt")
    assert set(torch.load(tmp_path / "ckpt.pt").keys()) == set(sd.keys())


def test_reference_functions_rebound_at_op_level(afr, ref, oracle):
    """modules.ddpm_utils.custom_upsample / custom_downsample are ours after patch(); the reference's own
    (saved before patching) agree with them on a CUDA tensor, and both with the oracle."""
    rf, ru, _ = ref
    assert ru.custom_upsample is afr.custom_upsample and ru.custom_downsample is afr.custom_downsample
    afr.unpatch()                                   # the fixture's finaliser tolerates a second unpatch
    k = rf.circularLowpassKernel(np.pi / 2, 3, 2)
    x = torch.randn(3, 5, 12, 20, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        want = rf.custom_downsample(torch.nn.functional.gelu(rf.custom_upsample(x.cuda(), k)), k)
        got = afr.filtered_gelu(x.cuda(), k, k)
    o = oracle.filtered_gelu(x.numpy(), k.numpy(), k.numpy())
    assert relmax(host(got), o) <= 1e-5
    assert relmax(host(want), o) <= 1e-5
