"""CPU-only tests: host logic, filter design, module/state_dict compatibility, the C-ABI
library's exported symbols, multi-rank plumbing on gloo.  No compute calls into the CUDA
library here (there is no GPU in the build container)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, golden

FS = dict(kernel_size=3, kaiser_beta=2, omega_c_down=np.pi / 2, omega_c_up=np.pi / 2)


def _reference():
    """The unmodified reference: /root/reference here, baseline/_ref elsewhere (tools/install_ref.py)."""
    from baseline import ref_loader
    if not ref_loader.available():
        pytest.skip("no reference checkout on this machine")
    return ref_loader.load()


@pytest.fixture(scope="module")
def afr():
    import aliasfree_b200 as m
    return m


def test_filter_design_matches_reference(afr):
    g = golden("taps.npz")
    n = len([k for k in g.files if k.startswith("k")])
    for i in range(n):
        w, N, beta, has = g[f"p{i}"]
        k = afr.circularLowpassKernel(w, int(N), beta if has else None)
        assert k.dtype == torch.float32 and not k.requires_grad and k.device.type == "cpu"
        np.testing.assert_allclose(k.numpy(), g[f"k{i}"], rtol=1e-6, atol=1e-9)


def test_library_loads_and_exports_every_declared_symbol(afr):
    header = open(os.path.join(ROOT, "include", "afr.h")).read()
    declared = set(re.findall(r"\b(afr_[a-z0-9_]+)\s*\(", header))
    declared -= {"afr_status", "afr_dtype", "afr_path"}
    assert {"afr_up2x_fwd", "afr_filtered_gelu_bwd", "afr_rotate_periodic_cubic"} <= declared
    path = afr.build()
    L = ctypes.CDLL(path)
    for name in sorted(declared):
        assert hasattr(L, name), name
    assert set(afr._native.EXPORTS) == declared
    assert afr._native.lib().afr_version() == 100


def test_c_abi_argument_errors_without_gpu(afr):
    """Argument validation happens before any CUDA call, so it is testable on CPU."""
    L = afr._native.lib()
    taps = (ctypes.c_float * 9)(*([1 / 9.0] * 9))
    assert L.afr_up2x_fwd(None, None, 1, 1, 0, 4, taps, 3, 0, 0, None) == 1      # bad shape
    assert L.afr_up2x_fwd(None, None, 1, 1, 4, 4, taps, 17, 0, 0, None) == 2     # bad taps
    assert L.afr_up2x_fwd(None, None, 1, 1, 4, 4, None, 3, 0, 0, None) == 2
    assert L.afr_up2x_fwd(None, None, 1, 1, 4, 4, taps, 3, 5, 0, None) == 3      # bad dtype
    assert L.afr_up2x_fwd(None, None, 1, 1, 4, 4, taps, 3, 0, 0, None) == 4      # null pointer
    assert L.afr_up2x_fwd(None, None, 0, 1, 4, 4, taps, 3, 0, 0, None) == 0      # empty batch: no-op
    assert L.afr_down2x_fwd(ctypes.c_void_p(2), ctypes.c_void_p(4), 1, 1, 4, 4, taps, 3, 0, None) == 5
    assert b"misaligned" in L.afr_last_error()
    assert L.afr_status_string(6) == b"CUDA error"


def test_c_abi_argument_errors_of_the_round2_entry_points(afr):
    """Strided upsample, variant-4 GELU + down, GroupNorm statistics / apply, affine adjoint: validation before CUDA."""
    L = afr._native.lib()
    taps = (ctypes.c_float * 9)(*([1 / 9.0] * 9))
    taps5 = (ctypes.c_float * 25)(*([1 / 25.0] * 25))
    p = ctypes.c_void_p(1 << 20)                       # never dereferenced: every call below fails validation first
    assert L.afr_up2x_fwd_strided(p, p, 2, 3, 4, 4, 10, taps, 3, 0, 0, None) == 1           # batch stride < C*2H*2W
    assert b"out_batch_stride" in L.afr_last_error()
    assert L.afr_up2x_fwd_strided(p, p, 2, 3, 4, 4, 4 * 3 * 16, taps5, 5, 0, 0, None) == 7  # N != 3: caller falls back
    assert L.afr_up2x_fwd_strided(None, None, 0, 3, 4, 4, 4 * 3 * 16, taps, 3, 0, 0, None) == 0
    assert L.afr_up2x_bwd_strided(p, p, 1, 1, 4, 4, 1, taps, 3, 0, None) == 1
    assert L.afr_gelu_down2x_fwd(p, None, None, p, 1, 1, 8, 8, taps5, 5, 0, None) == 7      # N != 3
    assert L.afr_gelu_down2x_fwd(p, None, None, p, 1, 1, 7, 8, taps, 3, 0, None) == 7       # odd H
    assert L.afr_gelu_down2x_fwd(p, p, None, p, 1, 1, 8, 8, taps, 3, 0, None) == 4          # scale without shift
    assert L.afr_gelu_down2x_bwd(p, None, p, 1, 1, 8, 8, taps, 3, 0, None) == 4
    assert L.afr_groupnorm1_stats(p, p, p, 1e-5, None, None, p, None, None, 1, 4, 4, 4, 0, None) == 4
    assert L.afr_groupnorm1_stats(p, p, p, 1e-5, None, p, p, None, None, 1, 3, 1, 1, 0, None) == 7   # C*H*W % 4
    assert L.afr_affine_apply(p, p, p, p, 1, 1, 3, 3, 0, None) == 7                          # H*W % 4
    assert L.afr_affine_apply(p, p, p, p, 0, 1, 4, 4, 0, None) == 0
    assert L.afr_filtered_gelu_affine_bwd(p, None, None, None, p, p, 1, 1, 8, 8, taps, 3, taps, 3, 0, None) == 4
    assert L.afr_set_path(-1) == 0 and afr._native.PATHS_AUTO()                               # pure query


def test_reference_install_is_byte_identical():
    """baseline/_ref (what travels to the GPU box) is the unmodified reference: every file's SHA-256 equals the
    manifest, and equals /root/reference when that checkout is present."""
    import hashlib
    import json
    ref = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref, "modules")):
        pytest.skip("baseline/_ref not installed (run tools/install_ref.py where /root/reference exists)")
    man = json.load(open(os.path.join(ref, "MANIFEST.json")))["sha256"]
    assert {"modules/filtrs.py", "modules/ddpm_utils.py", "modules/ddpm_models.py"} <= set(man)
    for rel, sha in man.items():
        assert hashlib.sha256(open(os.path.join(ref, rel), "rb").read()).hexdigest() == sha, rel
        live = os.path.join("/root/reference", rel)
        if os.path.exists(live):
            assert open(live, "rb").read() == open(os.path.join(ref, rel), "rb").read(), rel
    tracked = subprocess.run(["git", "-C", ROOT, "ls-files", "baseline/_ref"], capture_output=True, text=True).stdout
    assert tracked.strip() == ""                        # never committed


def test_groupnorm_fold_is_cuda_only_and_modules_fall_back_cleanly(afr):
    """The fold is decided per call (ops.norm_fusable): CPU tensors are never fusable, so the blocks reach the plain
    GroupNorm + filtered_gelu path, which then refuses CPU tensors loudly (no CPU fallback anywhere)."""
    blk = afr.DoubleConv_F(4, 4, f_settings=FS)
    h = torch.randn(2, 4, 8, 8)
    assert not afr.ops.norm_fusable(h, blk.norm1)
    assert not afr.ops.norm_fusable(h.to(torch.float64), blk.norm1)
    with pytest.raises(RuntimeError, match="CUDA"):
        blk(h)
    with pytest.raises(RuntimeError, match="CUDA"):
        afr.gelu_down2x(h, afr.circularLowpassKernel(np.pi / 2, 3, 2))
    with pytest.raises(RuntimeError, match="CUDA"):
        afr.up2x_cat(torch.randn(2, 4, 16, 16), h, afr.circularLowpassKernel(np.pi / 2, 3, 2))


def test_no_cpu_fallback(afr):
    k = afr.circularLowpassKernel(np.pi / 2, 3, 2)
    x = torch.randn(1, 2, 8, 8)
    for fn in (lambda: afr.up2x(x, k), lambda: afr.down2x(x, k), lambda: afr.filtered_gelu(x, k, k),
               lambda: afr.rotate(x, 3.0), lambda: afr.custom_upsample(x, k)):
        with pytest.raises(RuntimeError, match="CUDA"):
            fn()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "aliasfree-diffusion-models-pytorch_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.lower(), f
                assert "import baseline" not in src and "from baseline" not in src, f


def test_state_dict_keys_match_reference_layout(afr):
    net = afr.UNet(c_in=3, c_out=3, image_size=16, f_settings=FS, variant=3)
    keys = set(net.state_dict())
    for k in ("inc.conv1.weight", "inc.norm1.weight", "inc.norm1.bias", "down1.conv.0.conv1.weight",
              "down1.emb_layer.1.weight", "down1.emb_layer.1.bias", "up1.conv.1.norm2.weight",
              "sa1.mha.in_proj_weight", "sa3.ff_self.3.bias", "outc.weight", "outc.bias"):
        assert k in keys, k
    assert not any("filter" in k or "taps" in k for k in keys)
    k2 = set(afr.UNet(c_in=3, c_out=3, image_size=16, f_settings=FS, variant=2).state_dict())
    assert "down1.maxpool_conv.1.conv1.weight" in k2 and "up1.conv.0.norm1.weight" in k2
    k1 = set(afr.UNet(c_in=3, c_out=3, image_size=16, f_settings=FS, variant=1).state_dict())
    assert "down1.conv.0.double_conv.0.weight" in k1 and "inc.double_conv.1.weight" in k1
    with pytest.raises(ValueError):
        afr.UNet(variant=3)                     # f_settings missing, like the reference
    with pytest.raises(ValueError):
        afr.UNet(variant=7)
    n4 = afr.UNet(c_in=3, c_out=3, image_size=32, f_settings=FS, variant=4)
    assert sum(p.numel() for p in n4.parameters()) == 5898051      # variant 4 count


def test_state_dict_is_strictly_loadable_from_reference():
    rm = _reference()[2]
    import aliasfree_b200 as afr
    for v in (0, 1, 2, 3, 4):
        ref = rm.UNet(c_in=3, c_out=3, image_size=16, device="cpu", f_settings=FS, variant=v)
        ours = afr.UNet(c_in=3, c_out=3, image_size=16, f_settings=FS, variant=v)
        ours.load_state_dict(ref.state_dict(), strict=True)
        assert [tuple(p.shape) for p in ours.parameters()] == [tuple(p.shape) for p in ref.parameters()]


def test_patch_rebinds_reference_names_and_restores():
    """patch() makes the reference's own UNet build from our block classes (checked on CPU by
    construction + strict state_dict exchange; the forward needs a GPU) and unpatch() undoes it."""
    ru, rm = _reference()[1:]
    import aliasfree_b200 as afr
    orig = (rm.DoubleConv_F, ru.custom_upsample, rm.Diffusion.__dict__["rotate_2d_matrix"])
    ref_sd = rm.UNet(c_in=3, c_out=3, image_size=16, device="cpu", f_settings=FS, variant=3).state_dict()
    names = afr.patch()
    try:
        assert {"modules.ddpm_models.DoubleConv_F", "modules.ddpm_utils.custom_upsample",
                "modules.ddpm_utils.custom_downsample", "modules.ddpm_models.Up_FFF",
                "modules.ddpm_models.Diffusion.rotate_2d_matrix"} <= set(names)
        net = rm.UNet(c_in=3, c_out=3, image_size=16, device="cpu", f_settings=FS, variant=3)
        assert isinstance(net.inc, afr.DoubleConv_F) and isinstance(net.down1, afr.Down_FFF)
        assert isinstance(net.up3.conv[1], afr.DoubleConv_F)
        net.load_state_dict(ref_sd, strict=True)                 # reference checkpoint -> patched model
        assert ru.custom_upsample is afr.custom_upsample
        with pytest.raises(RuntimeError, match="CUDA"):          # and it really is our (GPU-only) path
            net(torch.randn(1, 3, 16, 16), torch.tensor([5]))
    finally:
        afr.unpatch()
    assert (rm.DoubleConv_F, ru.custom_upsample, rm.Diffusion.__dict__["rotate_2d_matrix"]) == orig


def test_variant0_unet_runs_on_cpu_and_matches_reference():
    """Variant 0 has no filters, so the wiring itself can be checked on CPU against the reference."""
    rm = _reference()[2]
    import aliasfree_b200 as afr
    from _fill import fill_params_
    ref = fill_params_(rm.UNet(c_in=3, c_out=3, image_size=16, device="cpu", variant=0))
    ours = fill_params_(afr.UNet(c_in=3, c_out=3, image_size=16, variant=0))
    x = torch.randn(2, 3, 16, 16); t = torch.tensor([3, 700])
    with torch.no_grad():
        assert torch.allclose(ours(x, t), ref(x, t), atol=1e-5, rtol=1e-5)


def test_diffusion_schedule_and_tables(afr):
    d = afr.Diffusion(noise_steps=50, img_size=8, device="cpu")
    assert d.beta.shape == (50,) and abs(d.beta[0].item() - 1e-4) < 1e-9 and abs(d.beta[-1].item() - 0.02) < 1e-9
    assert torch.allclose(d.alpha_hat, torch.cumprod(1 - d.beta, 0))
    i = 17
    assert abs(d._ca[i] - (1 / torch.sqrt(d.alpha[i])).item()) < 1e-7
    assert abs(d._cb[i] - ((1 - d.alpha[i]) / torch.sqrt(1 - d.alpha_hat[i])).item()) < 1e-7
    t = d.sample_timesteps(1000)
    assert t.min() >= 1 and t.max() < 50
    x = torch.randn(4, 1, 8, 8)
    xt, eps = d.noise_images(x, torch.tensor([1, 5, 20, 49]))
    assert xt.shape == x.shape and eps.shape == x.shape


def test_shard_bounds(afr):
    from aliasfree_b200.parallel import shard_bounds
    for n, ws in [(4096, 8), (10, 4), (3, 8), (0, 2)]:
        spans = [shard_bounds(n, r, ws) for r in range(ws)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1


_GLOO_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["AFR_ROOT"])
import aliasfree_b200 as afr
from aliasfree_b200 import parallel
dist.init_process_group("gloo")
rank, ws = dist.get_rank(), dist.get_world_size()
torch.manual_seed(100 + rank)                       # different init per rank on purpose
net = afr.UNet(c_in=1, c_out=1, image_size=8, variant=0)     # variant 0: runs on CPU
ddp = parallel.FlatGradAllReduce(net)               # broadcasts rank 0's parameters
flat0 = torch.cat([p.detach().flatten() for p in net.parameters()])
gathered = [torch.empty_like(flat0) for _ in range(ws)]
dist.all_gather(gathered, flat0)
assert all(torch.equal(gathered[0], g) for g in gathered), "broadcast failed"
diff = afr.Diffusion(noise_steps=20, img_size=8, device="cpu")
opt = torch.optim.AdamW(net.parameters(), lr=1e-3)
torch.manual_seed(7)
full = torch.rand(4 * ws, 1, 8, 8) * 2 - 1          # identical global batch on every rank
lo, hi = parallel.shard_bounds(full.shape[0], rank, ws)
# local gradient on the shard, then the single all-reduce
t = torch.arange(1, 1 + full.shape[0]) % 19 + 1
torch.manual_seed(11); eps = torch.randn_like(full)
def loss_on(sl):
    sa = torch.sqrt(diff.alpha_hat[t[sl]])[:, None, None, None]; sb = torch.sqrt(1 - diff.alpha_hat[t[sl]])[:, None, None, None]
    return torch.nn.functional.mse_loss(net(sa * full[sl] + sb * eps[sl], t[sl]), eps[sl])
ddp.zero_grad(); loss_on(slice(lo, hi)).backward(); ddp.sync()
sharded = ddp.flat.clone()
ddp.zero_grad(); loss_on(slice(0, full.shape[0])).backward()
assert torch.allclose(sharded, ddp.flat, atol=1e-6, rtol=1e-4), (sharded - ddp.flat).abs().max()
# sharded sampling: shards tile the global batch, no communication needed
x, res = None, None
lo2, hi2 = parallel.shard_bounds(6, rank, ws)
assert hi2 - lo2 == 6 // ws
dist.barrier()
if rank == 0: print("GLOO_OK")
'''


def test_two_rank_gloo_data_parallel(tmp_path):
    """world_size=2 on CPU: parameter broadcast and flat-gradient all-reduce reproduce the
    global-batch gradient (equal shards, MSE mean)."""
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    env = dict(os.environ, AFR_ROOT=ROOT, OMP_NUM_THREADS="2")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "GLOO_OK" in out.stdout


def test_shift_matches_scipy_grid_wrap(afr):
    """Diffusion.shift_2d_matrix (torch.roll) == scipy.ndimage.shift(..., mode='grid-wrap') for the
    integer shifts the reference uses (modules/ddpm_models.py:431-436)."""
    from scipy import ndimage
    x = torch.randn(2, 3, 8, 10)
    for h, v in [(1, 0), (-1, 0), (3, -2), (0, 5)]:
        want = ndimage.shift(input=x.numpy(), shift=(0, 0, v, h), mode="grid-wrap")
        got = afr.Diffusion.shift_2d_matrix(x, h, v, "cpu").numpy()
        np.testing.assert_allclose(got, want, atol=2e-6)


def test_blocks_deepcopy_and_pickle_after_taps_were_cached(afr):
    """The reference's EMA pattern deep-copies the model and users torch.save whole models; a cached
    ``Taps`` must not get in the way (it holds a tensor, the C pointer is taken at launch time)."""
    import copy
    import io
    import pickle
    blk = afr.DoubleConv_F(4, 4, residual=True, f_settings=FS)
    up, dn = blk._filters()                                  # what the first forward caches
    assert up.n == 3 and up.ptr.value == up.t.data_ptr()
    clone = copy.deepcopy(blk)
    up2, _ = clone._filters()
    assert up2.t.data_ptr() != up.t.data_ptr() and torch.equal(up2.t, up.t)   # its own storage, same taps
    back = pickle.loads(pickle.dumps(blk))
    assert torch.equal(back._filters()[0].t, up.t)
    buf = io.BytesIO()
    net = afr.UNet(c_in=3, c_out=3, image_size=16, f_settings=FS, variant=3)
    for m in net.modules():
        if isinstance(m, afr.DoubleConv_F):
            m._filters()
    torch.save(net, buf)
    buf.seek(0)
    again = torch.load(buf, weights_only=False)
    assert sorted(again.state_dict()) == sorted(net.state_dict())
    t = afr.Taps(afr.circularLowpassKernel(np.pi / 2, 3, 2))
    assert pickle.loads(pickle.dumps(t)).n == 3 and copy.deepcopy(t).t.data_ptr() != t.t.data_ptr()
