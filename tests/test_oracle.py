"""CPU tests: the oracle restatement is pinned to the reference's own outputs (tests/golden, produced by
oracle/make_golden.py from the unmodified reference) and to the reference's known-answer tests."""
import json
import os

import pytest
import torch

from fcvsr_b200.arch import FCVSR_SNet, FCVSRNet, GShiftNet, GShiftNet_S, seeded_state_dict
from oracle import fcvsr_oracle as O
from oracle import ref_loader
from tests.util import GOLD, load_golden, make_clip, make_clip_rgb


@pytest.mark.parametrize("name", ["fcvsr_s_64", "fcvsr_s_36x40"])
def test_oracle_matches_reference_golden(name):
    g = load_golden(name)
    c = g["case"]
    sd = seeded_state_dict(c["variant"], c["seed"])
    x = make_clip(c["clip_seed"], c["b"], c["h"], c["w"])
    with torch.no_grad():
        y, taps = O.forward(sd, x, return_taps=True)
    assert (y - g["out"]).abs().max().item() <= 2e-5
    for k in ("mgaa1", "mgaa2", "mffr", "sc_l1", "fuse"):
        assert (taps[k][..., ::4, ::4] - g[k]).abs().max().item() <= 5e-5, k
    assert (taps["sc_l3"] - g["sc_l3"]).abs().max().item() <= 5e-5


def test_oracle_matches_reference_golden_full():
    g = load_golden("fcvsr_full_64")
    c = g["case"]
    sd = seeded_state_dict(c["variant"], c["seed"])
    x = make_clip(c["clip_seed"], c["b"], c["h"], c["w"])
    with torch.no_grad():
        y = O.forward(sd, x)
    assert (y - g["out"]).abs().max().item() <= 2e-5


@pytest.mark.parametrize("name", ["fcvsrnet_s_32x40", "fcvsrnet_32"])
def test_oracle_matches_reference_golden_rgb(name):
    """The mmedit backbones FCVSRNet / FCVSR_SNet (RGB, 21 -> 3 channels): goldens made by the unmodified GShiftNet.forward with
    the two re-shaped layers (oracle/make_golden.py:build_reference; the mmedit file itself needs mmcv)."""
    g = load_golden(name)
    c = g["case"]
    sd = seeded_state_dict(c["variant"], c["seed"])
    x = make_clip_rgb(c["clip_seed"], c["b"], c["h"], c["w"])
    with torch.no_grad():
        y, taps = O.forward(sd, x, return_taps=True)
    assert y.shape == g["out"].shape and y.shape[1] == 3
    assert (y - g["out"]).abs().max().item() <= 2e-5
    for k in ("mgaa1", "mgaa2", "mffr", "sc_l1", "fuse"):
        assert (taps[k][..., ::4, ::4] - g[k]).abs().max().item() <= 5e-5, k


def test_state_dict_matches_reference_keys_and_shapes():
    """Drop-in contract: same keys, order and shapes as the reference state_dict, including
    the aliased RCB / body.3 entries."""
    with open(os.path.join(GOLD, "state_dict_shapes.json")) as f:
        ref = json.load(f)
    for variant, cls in (("S", GShiftNet_S), ("full", GShiftNet), ("rgb", FCVSRNet), ("rgb_S", FCVSR_SNet)):
        sd = cls().state_dict()
        mine = [[k, list(v.shape)] for k, v in sd.items()]
        assert mine == ref[variant]
        k0 = "recorb1.body.0.body.0"
        assert sd[k0 + ".RCB.body.0.weight"].data_ptr() == sd[k0 + ".body.3.body.0.weight"].data_ptr()
    assert sum(p.numel() for p in GShiftNet().parameters()) == 8811336
    assert sum(p.numel() for p in GShiftNet_S().parameters()) == 3704709


def test_dcn_oracle_known_answer():
    """ops/dcn/simple_check.py:8-22 (the reference's only KAT on this path)."""
    with open(os.path.join(GOLD, "dcn_simple_check.json")) as f:
        kat = json.load(f)
    x = torch.tensor(kat["input"])
    off = torch.tensor(kat["offset18"], dtype=torch.float32).view(1, 18, 1, 1).repeat(1, 2, 3, 3)
    w = torch.ones(1, 2, 3, 3)
    y = O.modulated_deform_conv(x, off, None, w, padding=1, deformable_groups=2)
    assert torch.equal(y.flatten(), torch.tensor(kat["expected"], dtype=torch.float32))


def test_dcn_oracle_matches_torchvision():
    """torchvision.ops.deform_conv2d shares the reference's offset/mask layout."""
    tv = pytest.importorskip("torchvision.ops")
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 8, 9, 11, generator=g)
    w = torch.randn(6, 4, 3, 3, generator=g)
    b = torch.randn(6, generator=g)
    off = 2.5 * torch.randn(2, 4 * 18, 9, 11, generator=g)
    msk = torch.rand(2, 4 * 9, 9, 11, generator=g)
    y = O.modulated_deform_conv(x, off, msk, w, b, padding=1, groups=2, deformable_groups=4)
    ref = tv.deform_conv2d(x, off, w, b, padding=1, mask=msk)
    assert (y - ref).abs().max().item() < 1e-4


def test_flow_warp_integer_shift():
    """mmedit_train/tests/test_models/test_common/test_flow_warp.py:31-52: flow = -1 is an exact
    one-pixel shift with zero fill."""
    x = torch.arange(16.0).view(1, 1, 4, 4)
    flow = -torch.ones(1, 2, 4, 4)
    y = O.warp_bilinear(x, flow)
    ref = torch.zeros_like(x)
    ref[..., 1:, 1:] = x[..., :-1, :-1]
    assert torch.allclose(y, ref, atol=1e-6)


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree only exists in the build container")
def test_oracle_blocks_against_live_reference():
    ref = ref_loader.load()
    g = torch.Generator().manual_seed(11)
    feat = torch.randn(1, 64, 12, 16, generator=g)
    taps = torch.randn(1, 192, 12, 16, generator=g)
    assert torch.allclose(O.sac(feat, taps), ref.SAC(feat, taps, torch.zeros_like(taps), 3), atol=1e-6)
    a = torch.randn(1, 128, 10, 9, generator=g)
    b = torch.randn(1, 128, 10, 9, generator=g)
    coords = ref.coords_grid(1, 10, 9, "cpu")
    assert torch.allclose(O.corr_lookup(a, b), ref.CorrBlock(a, b)(coords), atol=1e-5)


def test_dcn_oracle_gradients_match_torchvision():
    """Backward pin: autograd through the oracle restatement gives the gradients of the reference's backward kernels
    (dmcn_get_gradient_weight / dmcn_get_coordinate_weight, deform_conv_cuda_kernel.cu:499-567); checked here against the
    autograd of torchvision.ops.deform_conv2d, which shares the reference's layout.  float64: exact up to
    summation order."""
    import torchvision.ops as tv
    g = torch.Generator().manual_seed(11)
    x = torch.randn(2, 8, 9, 11, generator=g, dtype=torch.float64)
    w = torch.randn(6, 4, 3, 3, generator=g, dtype=torch.float64)
    b = torch.randn(6, generator=g, dtype=torch.float64)
    off = 2.5 * torch.randn(2, 4 * 2 * 9, 9, 11, generator=g, dtype=torch.float64)
    msk = torch.rand(2, 4 * 9, 9, 11, generator=g, dtype=torch.float64)
    gy = torch.randn(2, 6, 9, 11, generator=g, dtype=torch.float64)
    grads = []
    for fn in (lambda *a: O.modulated_deform_conv(a[0], a[1], a[2], a[3], a[4], padding=1, groups=2, deformable_groups=4),
               lambda *a: tv.deform_conv2d(a[0], a[1], a[3], a[4], padding=1, mask=a[2])):
        leaves = [t.clone().requires_grad_(True) for t in (x, off, msk, w, b)]
        (fn(*leaves) * gy).sum().backward()
        grads.append([t.grad for t in leaves])
    for name, a, r in zip(("input", "offset", "mask", "weight", "bias"), *grads):
        assert (a - r).abs().max().item() <= 1e-10 * max(1.0, r.abs().max().item()), name


@pytest.mark.parametrize("name", ["fcvsr_s_32_grads", "fcvsr_full_32_grads"])
def test_oracle_autograd_matches_reference_gradients(name):
    """Backward pin of the training step (BASELINE config 4): autograd through the oracle restatement reproduces the gradients
    of the UNMODIFIED reference (tests/golden/*_grads.pt, made by oracle/make_golden_grads.py) for every parameter and for the
    input clip; the parameters the reference leaves without gradient (DivEnh.Conv) get none here either,
    and the dead half of MGAA.F.1 (rows i*384+192 .. +383) has exactly zero gradient."""
    from oracle.make_golden_grads import charbonnier_sum, strided, target
    g = load_golden(name)
    c = g["case"]
    sd = seeded_state_dict(c["variant"], c["seed"])
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    x = make_clip(c["clip_seed"], c["b"], c["h"], c["w"]).requires_grad_()
    loss = charbonnier_sum(O.forward(leaves, x), target(c["target_seed"], c["b"], c["h"], c["w"]))
    loss.backward()
    assert abs(float(loss.detach()) - g["loss"]) <= 1e-5 * g["loss"]

    def grad_of(k):                      # the reference names an aliased module once (RCB); the oracle reads it as body.3
        t = leaves[k].grad
        if t is None and ".RCB." in k:
            t = leaves[k.replace(".RCB.", ".body.3.")].grad
        return t

    for k, ref in g["grads"].items():
        got = grad_of(k)
        assert got is not None, k
        tol = 2e-3 * max(ref["amax"], 1e-6)
        assert float((strided(got) - ref["samples"]).abs().max()) <= tol, k
        assert abs(float(got.norm()) - ref["norm"]) <= 2e-3 * max(ref["norm"], 1e-6), k
    for k in g["no_grad"]:
        t = grad_of(k)
        assert t is None or float(t.abs().max()) == 0.0, k
    assert float((strided(x.grad, 64) - g["dx"]).abs().max()) <= 2e-3 * float(g["dx"].abs().max())
    f1 = leaves["MGAA.F.1.weight"].grad
    a = f1.shape[0] // 384
    dead = torch.cat([f1[i * 384 + 192:(i + 1) * 384] for i in range(a)])
    assert float(dead.abs().max()) == 0.0


def test_metrics_oracle_matches_live_reference():
    """oracle/metrics_oracle.py against the reference's calculate_psnr / calculate_ssim (metric/psnr_ssim.py:278-399) called as
    the evaluation driver calls them (:470-471).  Needs the reference tree and cv2: skipped on the GPU box."""
    import importlib.util
    import numpy as np
    from oracle import metrics_oracle as M
    path = "/root/reference/CVSR_train/metric/psnr_ssim.py"
    if not os.path.isfile(path):
        pytest.skip("reference tree not present")
    pytest.importorskip("cv2")
    spec = importlib.util.spec_from_file_location("ref_psnr_ssim", path)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    rng = np.random.default_rng(3)
    for (h, w, noise) in ((60, 70, 20), (48, 33, 3), (40, 40, 0)):
        a = rng.integers(0, 256, (h, w, 1)).astype(np.uint8)
        b = np.clip(a.astype(int) + (rng.integers(-noise, noise + 1, a.shape) if noise else 0), 0, 255).astype(np.uint8)
        af, bf = a.astype(np.float64), b.astype(np.float64)
        p_ref, s_ref = ref.calculate_psnr(af, bf, 4, test_y_channel=True), ref.calculate_ssim(af, bf, 4, test_y_channel=True)
        p, s = M.calculate_psnr(af[..., 0], bf[..., 0], 4), M.calculate_ssim(af[..., 0], bf[..., 0], 4)
        assert (p == p_ref) or abs(p - p_ref) <= 1e-5 * abs(p_ref)
        assert abs(s - s_ref) <= 1e-12


@pytest.mark.parametrize("name", ["fcvsr_rgb_s_32x40", "fcvsr_rgb_full_32"])
def test_rgb_oracle_matches_reference_golden(name):
    """oracle/fcvsr_rgb_oracle.py (FCVSR / FCVSR_S of CVSR_freq_RGB.py) against goldens made by the unmodified reference
    (oracle/make_golden_rgb.py): output and stage taps."""
    pytest.importorskip("cv2")             # the 'ideal' band masks are rasterised with cv2.circle, as in the reference
    from fcvsr_b200.arch_rgb import seeded_state_dict_rgb
    from oracle import fcvsr_rgb_oracle as R
    g = load_golden(name)
    c = g["case"]
    sd = seeded_state_dict_rgb(c["variant"], c["seed"])
    x = make_clip_rgb(c["clip_seed"], c["b"], c["h"], c["w"])
    with torch.no_grad():
        y, taps = R.forward(sd, x, return_taps=True)
    assert (y - g["out"]).abs().max().item() <= 2e-5
    for k in ("mgaa1", "mgaa2", "mffr", "sc_l1", "fuse"):
        assert (taps[k][..., ::4, ::4] - g[k]).abs().max().item() <= 5e-5, k


def test_rgb_state_dict_matches_reference_keys_and_shapes():
    from fcvsr_b200.arch_rgb import FCVSR, FCVSR_S
    with open(os.path.join(GOLD, "state_dict_shapes_rgb.json")) as f:
        ref = json.load(f)
    for variant, cls in (("S", FCVSR_S), ("full", FCVSR)):
        assert [[k, list(v.shape)] for k, v in cls().state_dict().items()] == ref[variant]
    assert sum(p.numel() for p in FCVSR().parameters()) == 9042890
    assert sum(p.numel() for p in FCVSR_S().parameters()) == 4024999
