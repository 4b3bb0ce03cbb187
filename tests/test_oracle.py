"""Pins the CPU oracle (oracle/afr_oracle.c) against outputs of the reference itself
(tests/golden/*.npz, produced by tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from conftest import golden, relmax

TOL = 2e-6   # oracle accumulates in double; the reference conv is fp32


def test_taps_match_reference(oracle):
    g = golden("taps.npz")
    n = len([k for k in g.files if k.startswith("k")])
    assert n >= 10
    for i in range(n):
        w, N, beta, has = g[f"p{i}"]
        got = oracle.lowpass_taps(w, int(N), beta if has else None)
        assert got.dtype == np.float32 and got.shape == (int(N), int(N))
        np.testing.assert_allclose(got, g[f"k{i}"], rtol=2e-6, atol=1e-8)
        assert abs(float(got.astype(np.float64).sum()) - 1.0) < 1e-6


def test_canonical_taps_values(oracle):
    # N=3, beta=2, omega=pi/2
    k = oracle.lowpass_taps(np.pi / 2, 3, 2.0)
    np.testing.assert_allclose(k[1, 1], 0.37743, atol=1e-5)
    np.testing.assert_allclose(k[0, 1], 0.11949, atol=1e-5)
    np.testing.assert_allclose(k[0, 0], 0.03615, atol=1e-5)
    np.testing.assert_array_equal(k, k.T)


def _cases():
    return [str(s) for s in golden("resample.npz")["names"]]


@pytest.mark.parametrize("name", _cases())
def test_resample_forward(oracle, name):
    g = golden("resample.npz")
    x, ku, kd = g[f"{name}.x"], g[f"{name}.ku"], g[f"{name}.kd"]
    assert relmax(oracle.up2x(x, ku), g[f"{name}.up"]) <= TOL
    assert relmax(oracle.down2x(x, kd), g[f"{name}.down"]) <= TOL
    assert relmax(oracle.filtered_gelu(x, ku, kd), g[f"{name}.fused"]) <= TOL


@pytest.mark.parametrize("name", _cases())
def test_resample_backward(oracle, name):
    g = golden("resample.npz")
    x, ku, kd = g[f"{name}.x"], g[f"{name}.ku"], g[f"{name}.kd"]
    H, W = x.shape[-2:]
    assert relmax(oracle.up2x_bwd(g[f"{name}.du"], ku), g[f"{name}.dx_up"]) <= TOL
    assert relmax(oracle.down2x_bwd(g[f"{name}.dd"], kd, H, W), g[f"{name}.dx_down"]) <= TOL
    assert relmax(oracle.filtered_gelu_bwd(x, g[f"{name}.dy"], ku, kd), g[f"{name}.dx_fused"]) <= TOL


def test_fused_equals_composition(oracle):
    # the fused op is literally down(gelu(up(x))) with the GELU output zero padded
    from math import erf, sqrt
    rng = np.random.default_rng(0)
    x = rng.standard_normal((3, 6, 10)).astype(np.float32)
    ku = oracle.lowpass_taps(np.pi / 2, 3, 2.0); kd = oracle.lowpass_taps(np.pi / 3, 6, 1.0)
    u = oracle.up2x(x, ku).astype(np.float64)
    gel = 0.5 * u * (1.0 + np.vectorize(erf)(u / sqrt(2.0)))
    assert relmax(oracle.filtered_gelu(x, ku, kd), oracle.down2x(gel.astype(np.float32), kd)) <= 1e-6


def test_adjoint_identity(oracle):
    # <up(x), d> == <x, up^T(d)> and the same for down: checks the bwd restatements
    rng = np.random.default_rng(1)
    for N in (3, 4, 6):
        k = oracle.lowpass_taps(np.pi / 2, N, 2.0)
        x = rng.standard_normal((2, 7, 6)).astype(np.float32)
        d = rng.standard_normal((2, 14, 12)).astype(np.float32)
        lhs = float((oracle.up2x(x, k).astype(np.float64) * d).sum())
        rhs = float((x.astype(np.float64) * oracle.up2x_bwd(d, k)).sum())
        assert abs(lhs - rhs) <= 1e-5 * max(1.0, abs(lhs))
        d2 = rng.standard_normal((2, 4, 3)).astype(np.float32)
        lhs = float((oracle.down2x(x, k).astype(np.float64) * d2).sum())
        rhs = float((x.astype(np.float64) * oracle.down2x_bwd(d2, k, 7, 6)).sum())
        assert abs(lhs - rhs) <= 1e-5 * max(1.0, abs(lhs))


def test_rotate_matches_scipy_via_reference(oracle):
    g = golden("rotate.npz")
    x = g["x"]
    for i, a in enumerate(g["angles"]):
        got = oracle.rotate(x, float(a))
        assert np.abs(got - g[f"y{i}"]).max() <= 2e-6, a
    assert np.abs(oracle.rotate(g["xr"], 7.5) - g["yr"]).max() <= 2e-6


def test_empty_and_shapes(oracle):
    k = oracle.lowpass_taps(np.pi / 2, 3, 2.0)
    assert oracle.up2x(np.zeros((0, 4, 4), np.float32), k).shape == (0, 8, 8)
    assert oracle.down2x(np.zeros((2, 7, 9), np.float32), k).shape == (2, 4, 5)
    assert oracle.filtered_gelu(np.zeros((1, 1, 1), np.float32), k, k).shape == (1, 1, 1)


def test_installed_reference_matches_goldens():
    """The unmodified reference (baseline/_ref or /root/reference, loaded by baseline/ref_loader.py -- what bench.py
    times as the eager-GPU arm) reproduces, on CPU, the fixtures it generated: the loader hands out the real
    modules/filtrs.py and the fused sequence of modules/ddpm_utils.py:123-125."""
    import torch
    import torch.nn.functional as F
    from baseline import ref_loader
    assert ref_loader.available(), "run tools/install_ref.py (or __graft_entry__.build()) where /root/reference exists"
    filtrs, _, _ = ref_loader.load()
    g = golden("resample.npz")
    for name in _cases():
        x = torch.from_numpy(g[f"{name}.x"]); ku = torch.from_numpy(g[f"{name}.ku"]); kd = torch.from_numpy(g[f"{name}.kd"])
        up = filtrs.custom_upsample(x, ku)
        assert relmax(up.numpy(), g[f"{name}.up"]) <= 1e-6
        assert relmax(filtrs.custom_downsample(x, kd).contiguous().numpy(), g[f"{name}.down"]) <= 1e-6
        assert relmax(filtrs.custom_downsample(F.gelu(up), kd).contiguous().numpy(), g[f"{name}.fused"]) <= 1e-6
