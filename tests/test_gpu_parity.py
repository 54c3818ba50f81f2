"""GPU parity tests (run on the B200 box): every call goes through the C ABI of libfcvsr_b200.so and is
compared with the CPU oracle (oracle/fcvsr_oracle.py) and with the committed golden vectors produced
by the unmodified reference (tests/golden).

Tolerances (BASELINE.json north_star / SURVEY 8d):
  * fp32 CUDA-core path (``use_tc=False``): max-abs <= 2e-5 on the unclamped output.
  * TF32 tensor-core path (default; fp32 storage, TF32 operands, fp32 accumulate) -- the "fp32" mode of
    the contract: max-abs <= 1e-3 and |PSNR(ours,T) - PSNR(ref,T)| <= 0.01 dB.
"""
import json
import os

import pytest
import torch
import torch.nn.functional as F

from fcvsr_b200 import _capi as C, arch, bands
from fcvsr_b200.engine import Engine, _ConvPack
from oracle import fcvsr_oracle as O
from tests.util import GOLD, load_golden, make_clip, nchw, nhwc, psnr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev(lib):
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


def _st():
    return torch.cuda.current_stream().cuda_stream


def _model(variant, sd, dev, use_tc=True):
    m = (arch.GShiftNet_S if variant == "S" else arch.GShiftNet)().to(dev).eval()
    m.load_state_dict(sd)
    m._engine = Engine(m, use_tc=use_tc)
    m.compute_dtype = m._engine.mode
    return m


# ------------------------------------------------------------------------------------------------
# kernels
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("H,W,Cc", [(64, 64, 64), (36, 40, 24), (180, 320, 16), (272, 480, 12), (4, 4, 2), (38, 76, 8), (46, 124, 4)])
def test_fft_kernels_match_torch_fft(dev, H, W, Cc):
    g = torch.Generator().manual_seed(H * W)
    x = torch.randn(2, Cc, H, W, generator=g)
    Wf = W // 2 + 1
    xd = nhwc(x).to(dev)
    spec = torch.empty(2, H, Wf, Cc, 2, device=dev)
    tw_w, tw_h = bands.twiddles(W, dev), bands.twiddles(H, dev)
    C.call("fcvsr_fft_r2c_w", xd.data_ptr(), Cc, spec.data_ptr(), tw_w.data_ptr(), 2, H, W, Cc, _st())
    C.call("fcvsr_fft_c2c_h", spec.data_ptr(), spec.data_ptr(), tw_h.data_ptr(), 0, 2, H, Wf, Cc, 0, 1.0, 0, 1, 0, _st())
    ref = torch.fft.rfft2(x)
    got = torch.view_as_complex(spec.cpu()).permute(0, 3, 1, 2)
    assert float((got - ref).abs().max()) <= 2e-6 * float(ref.abs().max())
    # irfft2 on a learned (non-Hermitian) spectrum: Im of DC / Nyquist columns must be ignored
    z = torch.randn(2, Cc, H, Wf, dtype=torch.complex64, generator=g)
    zd = torch.view_as_real(z.permute(0, 2, 3, 1).contiguous()).contiguous().to(dev)
    y = torch.empty(2, H, W, Cc, device=dev)
    C.call("fcvsr_fft_c2c_h", zd.data_ptr(), zd.data_ptr(), tw_h.data_ptr(), 0, 2, H, Wf, Cc, 1, 1.0, 0, 1, 0, _st())
    C.call("fcvsr_fft_c2r_w", zd.data_ptr(), y.data_ptr(), Cc, tw_w.data_ptr(), 2, H, W, Cc, 1.0 / (H * W), _st())
    ref2 = torch.fft.irfft2(z, s=(H, W))
    assert float((nchw(y.cpu()) - ref2).abs().max()) <= 2e-6 * float(ref2.abs().max())


def test_fft_accepts_any_even_length_and_rejects_odd_width(dev):
    """torch.fft takes every length: prime factors above 17 run on the generic radix pass (38 = 2 * 19); a real transform of odd
    width (no packed two-channel form) is the one shape the operator refuses."""
    x = torch.randn(1, 4, 38, 2, device=dev)
    out = torch.zeros(1, 4, 20, 2, 2, device=dev)
    tw = bands.twiddles(38, dev)
    rc = C.try_call("fcvsr_fft_r2c_w", x.data_ptr(), 2, out.data_ptr(), tw.data_ptr(), 1, 4, 38, 2, _st())
    assert rc == 0
    torch.cuda.synchronize()
    ref = torch.fft.rfft(x.cpu(), dim=2)                      # [1, 4, 20, 2] complex (channels last)
    got = torch.view_as_complex(out.cpu().contiguous())
    assert float((got - ref).abs().max()) <= 2e-5 * float(ref.abs().max())
    rc = C.try_call("fcvsr_fft_r2c_w", x.data_ptr(), 2, out.data_ptr(), tw.data_ptr(), 1, 4, 37, 2, _st())
    assert rc == C.ERR_ARG


def _run_conv(dev, x, w, b, tc, act=0, slope=0.0, res=None, ps=False, stride=1):
    B, Cin, H, W = x.shape
    pk = _ConvPack(w.to(dev), b.to(dev) if b is not None else None, stride=stride, ps=ps)
    xd = nhwc(x).to(dev)
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    cout = w.shape[0]
    y = torch.empty((B, 2 * Ho, 2 * Wo, cout // 4) if ps else (B, Ho, Wo, cout), device=dev)
    rd = nhwc(res).to(dev) if res is not None else None
    bias = pk.bias.data_ptr() if b is not None else 0
    rptr = rd.data_ptr() if rd is not None else 0
    if tc:
        C.call("fcvsr_conv2d_tc", xd.data_ptr(), Cin, pk.w_tc.data_ptr(), bias, rptr, cout, 0, 0, y.data_ptr(),
               y.shape[-1], B, H, W, Cin, cout, w.shape[-1], act, slope, 0, int(ps), 0, 0, 0, 0, 0, _st())
    else:
        C.call("fcvsr_conv2d_direct", xd.data_ptr(), Cin, 0, pk.w_direct.data_ptr(), bias, rptr, cout, 0, 0,
               y.data_ptr(), y.shape[-1], B, H, W, Cin, cout, w.shape[-1], stride, act, slope, 0, int(ps), 0, 0, 0, 0, 0, _st())
    torch.cuda.synchronize()
    return nchw(y.cpu())


CONV_CASES = [(1, 64, 64, 16, 16, 3), (2, 64, 128, 20, 36, 3), (1, 128, 64, 45, 80, 3), (1, 64, 256, 12, 20, 3),
              (1, 96, 64, 9, 17, 3), (2, 256, 128, 10, 33, 1), (1, 64, 4, 11, 19, 1), (1, 64, 1, 24, 40, 3),
              (1, 64, 576, 8, 16, 1), (1, 64, 1152, 5, 7, 1)]


@pytest.mark.parametrize("case", CONV_CASES + [(1, 7, 448, 12, 12, 3), (1, 209, 64, 6, 9, 1)])
def test_conv_direct_matches_fp32_reference(dev, case):
    B, ci, co, H, W, k = case
    g = torch.Generator().manual_seed(sum(case))
    x = torch.randn(B, ci, H, W, generator=g)
    w = torch.randn(co, ci, k, k, generator=g) / (ci * k * k) ** 0.5
    b = torch.randn(co, generator=g)
    res = torch.randn(B, co, H, W, generator=g)
    y = _run_conv(dev, x, w, b, False, act=2, slope=0.1, res=res)
    ref = F.leaky_relu(F.conv2d(x, w, b, padding=k // 2), 0.1) + res
    assert float((y - ref).abs().max()) <= 2e-5


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_tcgen05_matches_fp32_reference(dev, case):
    """TF32 operands: tolerance 2e-3 relative to the output scale (K up to 1152 products of ~N(0,1))."""
    B, ci, co, H, W, k = case
    g = torch.Generator().manual_seed(sum(case))
    x = torch.randn(B, ci, H, W, generator=g)
    w = torch.randn(co, ci, k, k, generator=g) / (ci * k * k) ** 0.5
    b = torch.randn(co, generator=g)
    res = torch.randn(B, co, H, W, generator=g)
    y = _run_conv(dev, x, w, b, True, act=2, slope=0.1, res=res)
    ref = F.leaky_relu(F.conv2d(x, w, b, padding=k // 2), 0.1) + res
    assert float((y - ref).abs().max()) <= 2e-3 * max(1.0, float(ref.abs().max()))
    if co % 64 == 0:
        y = _run_conv(dev, x, w, b, True, ps=True)
        ref = F.pixel_shuffle(F.conv2d(x, w, b, padding=k // 2), 2)
        assert float((y - ref).abs().max()) <= 2e-3 * max(1.0, float(ref.abs().max()))


def test_conv_direct_stride2(dev):
    g = torch.Generator().manual_seed(2)
    x = torch.randn(1, 64, 16, 20, generator=g)
    w = torch.randn(64, 64, 3, 3, generator=g) / 24
    b = torch.randn(64, generator=g)
    y = _run_conv(dev, x, w, b, False, stride=2)
    assert float((y - F.conv2d(x, w, b, stride=2, padding=1)).abs().max()) <= 2e-5


def test_conv_tc_reports_unsupported_shapes(dev):
    x = torch.zeros(1, 8, 8, 48, device=dev)
    w = torch.zeros(64, 48, device=dev)
    y = torch.zeros(1, 8, 8, 64, device=dev)
    rc = C.try_call("fcvsr_conv2d_tc", x.data_ptr(), 48, w.data_ptr(), 0, 0, 0, 0, 0, y.data_ptr(), 64, 1, 8, 8, 48, 64, 1,
                    0, 0.0, 0, 0, 0, 0, 0, 0, 0, _st())
    assert rc == C.ERR_UNSUPPORTED          # Cin % 32 != 0 -> caller must use fcvsr_conv2d_direct


# ------------------------------------------------------------------------------------------------
# model-level parity
# ------------------------------------------------------------------------------------------------
def _stage_taps(model, B, H, W, dev):
    ws = model._engine._workspace(B, H, W, dev)

    def tap(t, c0, c1, h, w):
        return nchw(t.view(B, h, w, -1)[..., c0:c1].cpu())

    return {"mgaa1": tap(ws["feat"], 128, 192, H, W), "mgaa2": tap(ws["m2"], 0, 64, H, W),
            "mffr": tap(ws["xs0"], 0, 64, H, W), "sc_l1": tap(ws["fuse"], 0, 64, H, W),
            "sc_l3": tap(ws["o3"], 0, 64, H // 4, W // 4), "fuse": tap(ws["f2"], 0, 64, H, W)}


@pytest.mark.parametrize("name", ["fcvsr_s_64", "fcvsr_s_36x40", "fcvsr_full_64"])
def test_fp32_path_matches_reference_golden(dev, name):
    """CUDA-core fp32 path against the golden outputs of the unmodified reference, stage by stage."""
    g = load_golden(name)
    c = g["case"]
    sd = arch.seeded_state_dict(c["variant"], c["seed"])
    x = make_clip(c["clip_seed"], c["b"], c["h"], c["w"])
    m = _model(c["variant"], sd, dev, use_tc=False)
    with torch.no_grad():
        y = m(x.to(dev)).cpu()
    taps = _stage_taps(m, c["b"], c["h"], c["w"], dev)
    for k in ("mgaa1", "mgaa2", "mffr", "sc_l1", "fuse"):
        assert float((taps[k][..., ::4, ::4] - g[k]).abs().max()) <= 1e-4, k
    assert float((taps["sc_l3"] - g["sc_l3"]).abs().max()) <= 1e-4
    assert float((y - g["out"]).abs().max()) <= 2e-5
    assert m._engine.tc_launches == 0 and m._engine.launches > 100


@pytest.mark.parametrize("name", ["fcvsr_s_64", "fcvsr_s_36x40", "fcvsr_full_64"])
def test_tf32_path_matches_reference_golden(dev, name):
    """Default (tcgen05, TF32 operands) path: the contract's fp32 tolerance, max-abs <= 1e-3 and
    PSNR delta <= 0.01 dB against a fixed random target."""
    g = load_golden(name)
    c = g["case"]
    sd = arch.seeded_state_dict(c["variant"], c["seed"])
    x = make_clip(c["clip_seed"], c["b"], c["h"], c["w"])
    m = _model(c["variant"], sd, dev, use_tc=True)
    with torch.no_grad():
        y = m(x.to(dev)).cpu()
    ref = g["out"]
    assert m._engine.tc_launches > 100            # the tensor-core kernel is the one that ran
    assert float((y - ref).abs().max()) <= 1e-3
    tgt = torch.rand(ref.shape, generator=torch.Generator().manual_seed(99))
    assert abs(psnr(y, tgt) - psnr(ref, tgt)) <= 0.01


@pytest.mark.parametrize("name", ["fcvsrnet_s_32x40", "fcvsrnet_32"])
def test_mmedit_rgb_variants_match_reference_golden(dev, name):
    """FCVSRNet / FCVSR_SNet (mmedit backbones, [B,7,3,H,W] -> [B,3,4H,4W]) in the three compute modes against the reference
    golden, and the differentiable forward in fp32 mode."""
    from tests.util import make_clip_rgb
    g = load_golden(name)
    c = g["case"]
    sd = arch.seeded_state_dict(c["variant"], c["seed"])
    x = make_clip_rgb(c["clip_seed"], c["b"], c["h"], c["w"]).to(dev)
    m = (arch.FCVSRNet if c["variant"] == "rgb" else arch.FCVSR_SNet)().to(dev).eval()
    m.load_state_dict(sd)
    for mode, tol in (("fp32", 2e-5), ("tf32", 1e-3), ("bf16", 5e-3)):
        m.compute_dtype = mode
        with torch.no_grad():
            y = m(x).cpu()
        assert y.shape == g["out"].shape
        err = float((y - g["out"]).abs().max())
        assert err <= tol, (mode, err)
        if mode == "bf16":
            assert psnr(y, g["out"]) >= 60.0
    m.compute_dtype = "fp32"
    yt = m(x)                                             # autograd recording: fcvsr_b200.train_forward
    assert yt.requires_grad and float((yt.detach().cpu() - g["out"]).abs().max()) <= 2e-5


def test_forward_against_oracle_unseen_shape(dev):
    """A shape with no golden file (ragged tiles: 44 x 52, batch 2) against the live oracle."""
    sd = arch.seeded_state_dict("S", 5)
    x = make_clip(321, 2, 44, 52)
    with torch.no_grad():
        ref = O.forward(sd, x)
    for use_tc, tol in ((False, 2e-5), (True, 1e-3)):
        m = _model("S", sd, dev, use_tc=use_tc)
        with torch.no_grad():
            y = m(x.to(dev)).cpu()
        assert float((y - ref).abs().max()) <= tol


def test_drop_in_api_errors(dev):
    m = arch.GShiftNet_S().to(dev).eval()
    with torch.no_grad():
        with pytest.raises(ValueError):
            m(torch.zeros(1, 7, 64, 64, device=dev))                 # not 5-D
        with pytest.raises(ValueError):
            m(torch.zeros(1, 7, 1, 30, 32, device=dev))              # H % 4 != 0
        with pytest.raises(RuntimeError):
            m(torch.zeros(1, 7, 1, 32, 32))                          # CPU tensor: no fallback
    y = m(torch.zeros(1, 7, 1, 32, 32, device=dev))                  # autograd recording: the differentiable forward runs
    assert y.requires_grad and y.shape == (1, 1, 128, 128)
    with pytest.raises(ValueError):
        m(torch.zeros(1, 7, 1, 30, 32, device=dev))                  # same shape contract under autograd


def test_weights_repack_after_load(dev):
    """load_state_dict after a forward must invalidate the packed weights."""
    sd_a, sd_b = arch.seeded_state_dict("S", 0), arch.seeded_state_dict("S", 7)
    x = make_clip(5, 1, 32, 32).to(dev)
    m = _model("S", sd_a, dev)
    with torch.no_grad():
        ya = m(x).clone()
        m.load_state_dict(sd_b)
        yb = m(x).clone()
        m.load_state_dict(sd_a)
        ya2 = m(x)
    assert float((ya - yb).abs().max()) > 1e-3
    assert torch.equal(ya, ya2)               # deterministic: same weights, same bits


# ------------------------------------------------------------------------------------------------
# DCN operator
# ------------------------------------------------------------------------------------------------
def test_dcn_known_answer(dev):
    """ops/dcn/simple_check.py:8-22 through the C ABI."""
    from fcvsr_b200.ops.dcn import DeformConv
    with open(os.path.join(GOLD, "dcn_simple_check.json")) as f:
        kat = json.load(f)
    conv = DeformConv(2, 1, kernel_size=3, padding=1, deformable_groups=2).to(dev)
    torch.nn.init.constant_(conv.weight, 1)
    off = torch.tensor(kat["offset18"], dtype=torch.float32, device=dev).view(1, 18, 1, 1).repeat(1, 2, 3, 3)
    x = torch.tensor(kat["input"], device=dev)
    with torch.no_grad():
        y = conv(x, off)
    assert torch.equal(y.cpu().flatten(), torch.tensor(kat["expected"], dtype=torch.float32))


@pytest.mark.parametrize("cfg", [dict(B=2, cin=16, cout=16, H=9, W=11, k=3, g=1, dg=16, pad=1, stride=1, dil=1),
                                 dict(B=1, cin=64, cout=64, H=20, W=24, k=3, g=1, dg=16, pad=1, stride=1, dil=1),
                                 dict(B=2, cin=8, cout=6, H=9, W=11, k=3, g=2, dg=4, pad=1, stride=2, dil=1),
                                 dict(B=1, cin=4, cout=4, H=12, W=10, k=5, g=1, dg=4, pad=4, stride=1, dil=2)])
def test_modulated_dcn_matches_oracle(dev, cfg, monkeypatch):
    """Exact-fp32 kernel (dcn.cu) against the loop oracle."""
    import fcvsr_b200.ops.dcn as dcn_mod
    from fcvsr_b200.ops.dcn import modulated_deform_conv
    monkeypatch.setattr(dcn_mod, "PRECISION", "fp32")
    g = torch.Generator().manual_seed(cfg["H"] * cfg["W"])
    k, dg = cfg["k"], cfg["dg"]
    x = torch.randn(cfg["B"], cfg["cin"], cfg["H"], cfg["W"], generator=g)
    w = torch.randn(cfg["cout"], cfg["cin"] // cfg["g"], k, k, generator=g) / 6
    b = torch.randn(cfg["cout"], generator=g)
    ho = (cfg["H"] + 2 * cfg["pad"] - (cfg["dil"] * (k - 1) + 1)) // cfg["stride"] + 1
    wo = (cfg["W"] + 2 * cfg["pad"] - (cfg["dil"] * (k - 1) + 1)) // cfg["stride"] + 1
    off = 3.0 * torch.randn(cfg["B"], dg * 2 * k * k, ho, wo, generator=g)
    msk = torch.rand(cfg["B"], dg * k * k, ho, wo, generator=g)
    ref = O.modulated_deform_conv(x, off, msk, w, b, cfg["stride"], cfg["pad"], cfg["dil"], cfg["g"], dg)
    with torch.no_grad():
        y = modulated_deform_conv(x.to(dev), off.to(dev), msk.to(dev), w.to(dev), b.to(dev), cfg["stride"], cfg["pad"],
                                  cfg["dil"], cfg["g"], dg)
    assert float((y.cpu() - ref).abs().max()) <= 1e-4


@pytest.mark.parametrize("cfg", [dict(B=2, cin=16, cout=16, H=9, W=11, k=3, g=1, dg=16, pad=1, stride=1, dil=1, mask=True),
                                 dict(B=1, cin=64, cout=64, H=20, W=24, k=3, g=1, dg=16, pad=1, stride=1, dil=1, mask=True),
                                 dict(B=2, cin=8, cout=6, H=9, W=11, k=3, g=2, dg=4, pad=1, stride=2, dil=1, mask=True),
                                 dict(B=1, cin=4, cout=4, H=12, W=10, k=5, g=1, dg=4, pad=4, stride=1, dil=2, mask=True),
                                 dict(B=3, cin=6, cout=10, H=14, W=13, k=3, g=1, dg=2, pad=1, stride=1, dil=1, mask=True),
                                 dict(B=2, cin=12, cout=72, H=11, W=17, k=3, g=1, dg=1, pad=0, stride=1, dil=1, mask=False),
                                 dict(B=1, cin=2, cout=1, H=4, W=4, k=3, g=1, dg=2, pad=1, stride=1, dil=1, mask=False),
                                 dict(B=2, cin=16, cout=12, H=15, W=18, k=3, g=2, dg=2, pad=2, stride=2, dil=2, mask=True)])
def test_dcn_backward_matches_oracle(dev, cfg, monkeypatch):
    """fcvsr_modulated_deform_conv_backward (dcn_bwd.cu) through the autograd Functions of ops.dcn against autograd through
    the CPU oracle (pinned to torchvision's deform_conv2d backward in tests/test_oracle.py): all five gradients, groups,
    stride / dilation, ragged tiles, deformable-group widths 1, 2, 3, 4, 8 and 12, DCNv1 (no mask); shapes with
    (Cin/groups) % 4 == 0 and (Cin/dg) % 4 == 0 take the NHWC vector-reduction path, and are run on the scalar NCHW path as
    well.  fp32 with atomics: max-abs <= 2e-4 of each gradient's scale."""
    import fcvsr_b200.ops.dcn as dcn_mod
    monkeypatch.setattr(dcn_mod, "PRECISION", "fp32")
    g = torch.Generator().manual_seed(cfg["H"] * cfg["W"] + 7)
    k, dg = cfg["k"], cfg["dg"]
    x = torch.randn(cfg["B"], cfg["cin"], cfg["H"], cfg["W"], generator=g)
    w = torch.randn(cfg["cout"], cfg["cin"] // cfg["g"], k, k, generator=g) / 6
    b = torch.randn(cfg["cout"], generator=g)
    ho = (cfg["H"] + 2 * cfg["pad"] - (cfg["dil"] * (k - 1) + 1)) // cfg["stride"] + 1
    wo = (cfg["W"] + 2 * cfg["pad"] - (cfg["dil"] * (k - 1) + 1)) // cfg["stride"] + 1
    off = 3.0 * torch.randn(cfg["B"], dg * 2 * k * k, ho, wo, generator=g)
    msk = torch.rand(cfg["B"], dg * k * k, ho, wo, generator=g)
    gy = torch.randn(cfg["B"], cfg["cout"], ho, wo, generator=g)
    if cfg["mask"]:
        names = ("input", "offset", "mask", "weight", "bias")
        cpu = [t.clone().requires_grad_(True) for t in (x, off, msk, w, b)]
        (O.modulated_deform_conv(cpu[0], cpu[1], cpu[2], cpu[3], cpu[4], cfg["stride"], cfg["pad"], cfg["dil"], cfg["g"],
                                 dg) * gy).sum().backward()
        gpu = [t.to(dev).requires_grad_(True) for t in (x, off, msk, w, b)]
        y = dcn_mod.modulated_deform_conv(gpu[0], gpu[1], gpu[2], gpu[3], gpu[4], cfg["stride"], cfg["pad"], cfg["dil"],
                                          cfg["g"], dg)
    else:
        names = ("input", "offset", "weight")
        cpu = [t.clone().requires_grad_(True) for t in (x, off, w)]
        (O.modulated_deform_conv(cpu[0], cpu[1], None, cpu[2], None, cfg["stride"], cfg["pad"], cfg["dil"], cfg["g"], dg)
         * gy).sum().backward()
        gpu = [t.to(dev).requires_grad_(True) for t in (x, off, w)]
        y = dcn_mod.deform_conv(gpu[0], gpu[1], gpu[2], cfg["stride"], cfg["pad"], cfg["dil"], cfg["g"], dg, 64)
    for fast in (True, False):
        monkeypatch.setattr(dcn_mod, "BACKWARD_NHWC", fast)
        for t in gpu:
            t.grad = None
        (y * gy.to(dev)).sum().backward(retain_graph=True)
        torch.cuda.synchronize()
        for name, a, r in zip(names, gpu, cpu):
            err = float((a.grad.cpu() - r.grad).abs().max())
            assert err <= 2e-4 * max(1.0, float(r.grad.abs().max())), (name, fast, err)


def test_dcn_backward_partial_needs_and_module(dev, monkeypatch):
    """Only the requested gradients are produced (frozen offset / mask), and the ModulatedDeformConv module trains."""
    import fcvsr_b200.ops.dcn as dcn_mod
    monkeypatch.setattr(dcn_mod, "PRECISION", "fp32")
    g = torch.Generator().manual_seed(5)
    m = dcn_mod.ModulatedDeformConv(8, 8, 3, padding=1, deformable_groups=2).to(dev)
    x = torch.randn(1, 8, 10, 12, generator=g).to(dev)
    off = torch.randn(1, 36, 10, 12, generator=g).to(dev)
    msk = torch.rand(1, 18, 10, 12, generator=g).to(dev)
    y = m(x, off, msk)
    y.square().sum().backward()
    assert m.weight.grad is not None and m.bias.grad is not None and x.grad is None and off.grad is None
    wr = m.weight.detach().cpu().clone().requires_grad_(True)
    br = m.bias.detach().cpu().clone().requires_grad_(True)
    O.modulated_deform_conv(x.cpu(), off.cpu(), msk.cpu(), wr, br, 1, 1, 1, 1, 2).square().sum().backward()
    assert float((m.weight.grad.cpu() - wr.grad).abs().max()) <= 2e-4 * float(wr.grad.abs().max())
    assert float((m.bias.grad.cpu() - br.grad).abs().max()) <= 2e-4 * float(br.grad.abs().max())


def test_dcn_backward_adjoint_identity_full_size(dev, monkeypatch):
    """Size-independent property at the benchmark shape (64 -> 64, 3x3, dg 16, 180x320): the operator is linear in its
    input and in its weight, so <gy, DCN(dx; w)> == <grad_input, dx> and <gy, DCN(x; dw)> == <grad_weight, dw> (no bias)."""
    import fcvsr_b200.ops.dcn as dcn_mod
    monkeypatch.setattr(dcn_mod, "PRECISION", "fp32")
    g = torch.Generator(device=dev).manual_seed(3)
    B, Cc, H, W, dg = 1, 64, 180, 320, 16
    x = torch.randn(B, Cc, H, W, device=dev, generator=g, requires_grad=True)
    w = (torch.randn(Cc, Cc, 3, 3, device=dev, generator=g) / 24).requires_grad_(True)
    off = 2.0 * torch.randn(B, dg * 18, H, W, device=dev, generator=g)
    msk = torch.rand(B, dg * 9, H, W, device=dev, generator=g)
    gy = torch.randn(B, Cc, H, W, device=dev, generator=g)
    dx = torch.randn(B, Cc, H, W, device=dev, generator=g)
    dw = torch.randn(Cc, Cc, 3, 3, device=dev, generator=g) / 24
    y = dcn_mod.modulated_deform_conv(x, off, msk, w, None, 1, 1, 1, 1, dg)
    (y * gy).sum().backward()
    with torch.no_grad():
        lhs_x = float((gy.double() * dcn_mod.modulated_deform_conv(dx, off, msk, w, None, 1, 1, 1, 1, dg).double()).sum())
        rhs_x = float((x.grad.double() * dx.double()).sum())
        lhs_w = float((gy.double() * dcn_mod.modulated_deform_conv(x, off, msk, dw, None, 1, 1, 1, 1, dg).double()).sum())
        rhs_w = float((w.grad.double() * dw.double()).sum())
    assert abs(lhs_x - rhs_x) <= 1e-4 * max(1.0, abs(lhs_x)), (lhs_x, rhs_x)
    assert abs(lhs_w - rhs_w) <= 1e-4 * max(1.0, abs(lhs_w)), (lhs_w, rhs_w)


@pytest.mark.parametrize("cfg", [dict(B=1, cin=64, cout=64, H=20, W=24, k=3, dg=16, pad=1, stride=1, dil=1, mask=True),
                                 dict(B=2, cin=32, cout=48, H=17, W=23, k=3, dg=1, pad=2, stride=2, dil=2, mask=True),
                                 dict(B=2, cin=64, cout=128, H=13, W=9, k=1, dg=8, pad=0, stride=1, dil=1, mask=True),
                                 dict(B=1, cin=96, cout=16, H=30, W=31, k=3, dg=3, pad=1, stride=1, dil=1, mask=False),
                                 dict(B=1, cin=32, cout=256, H=12, W=20, k=5, dg=2, pad=2, stride=1, dil=1, mask=True)])
def test_modulated_dcn_tensor_core_path(dev, cfg):
    """dcn_tc.cu (tcgen05, TF32-rounded operands, fp32 accumulate) against the loop oracle: the tolerance is the
    model's "tf32" contract, 1e-3 of the output scale; ragged last tiles, stride / dilation, v1 (no mask), all
    deformable-group widths (several groups per 16-channel run, one group for all channels)."""
    import fcvsr_b200.ops.dcn as dcn_mod
    from fcvsr_b200.ops.dcn import deform_conv, modulated_deform_conv
    assert dcn_mod.PRECISION == "tf32"
    g = torch.Generator().manual_seed(cfg["H"] * cfg["W"] + cfg["cin"])
    k, dg = cfg["k"], cfg["dg"]
    x = torch.randn(cfg["B"], cfg["cin"], cfg["H"], cfg["W"], generator=g)
    w = torch.randn(cfg["cout"], cfg["cin"], k, k, generator=g) / (cfg["cin"] * k * k) ** 0.5
    b = torch.randn(cfg["cout"], generator=g)
    ho = (cfg["H"] + 2 * cfg["pad"] - (cfg["dil"] * (k - 1) + 1)) // cfg["stride"] + 1
    wo = (cfg["W"] + 2 * cfg["pad"] - (cfg["dil"] * (k - 1) + 1)) // cfg["stride"] + 1
    off = 3.0 * torch.randn(cfg["B"], dg * 2 * k * k, ho, wo, generator=g)
    if cfg["mask"]:
        msk = torch.rand(cfg["B"], dg * k * k, ho, wo, generator=g)
        ref = O.modulated_deform_conv(x, off, msk, w, b, cfg["stride"], cfg["pad"], cfg["dil"], 1, dg)
        with torch.no_grad():
            y = modulated_deform_conv(x.to(dev), off.to(dev), msk.to(dev), w.to(dev), b.to(dev), cfg["stride"], cfg["pad"],
                                      cfg["dil"], 1, dg)
    else:
        ref = O.modulated_deform_conv(x, off, torch.ones(cfg["B"], dg * k * k, ho, wo), w, None, cfg["stride"], cfg["pad"],
                                      cfg["dil"], 1, dg)
        with torch.no_grad():
            y = deform_conv(x.to(dev), off.to(dev), w.to(dev), cfg["stride"], cfg["pad"], cfg["dil"], 1, dg, 64)
    err = float((y.cpu() - ref).abs().max())
    assert err <= 1e-3 * max(1.0, float(ref.abs().max())), err
    # and the two kernels agree with each other to the same bound
    dcn_mod.PRECISION = "fp32"
    try:
        with torch.no_grad():
            if cfg["mask"]:
                y32 = modulated_deform_conv(x.to(dev), off.to(dev), msk.to(dev), w.to(dev), b.to(dev), cfg["stride"],
                                            cfg["pad"], cfg["dil"], 1, dg)
            else:
                y32 = deform_conv(x.to(dev), off.to(dev), w.to(dev), cfg["stride"], cfg["pad"], cfg["dil"], 1, dg, 64)
    finally:
        dcn_mod.PRECISION = "tf32"
    assert float((y32.cpu() - ref).abs().max()) <= 1e-4
    assert not torch.equal(y32, y)            # the tensor-core kernel really ran


def test_modulated_dcn_pack_module(dev):
    """ModulatedDeformConvPack (deform_conv.py:311-337): conv_offset_mask -> chunk/cat/sigmoid -> DCN."""
    from fcvsr_b200.ops.dcn import ModulatedDeformConvPack
    torch.manual_seed(3)                       # 16 channels: outside the tensor-core kernel's class -> exact kernel
    m = ModulatedDeformConvPack(16, 16, 3, stride=1, padding=1, deformable_groups=4).to(dev)
    with torch.no_grad():
        m.conv_offset_mask.weight.normal_(0, 0.05)
        m.conv_offset_mask.bias.normal_(0, 0.5)
        m.bias.normal_(0, 0.1)
        x = torch.randn(2, 16, 10, 12, device=dev)
        y = m(x).cpu()
        o = F.conv2d(x.cpu(), m.conv_offset_mask.weight.cpu(), m.conv_offset_mask.bias.cpu(), padding=1)
        o1, o2, mk = torch.chunk(o, 3, dim=1)
        ref = O.modulated_deform_conv(x.cpu(), torch.cat((o1, o2), 1), torch.sigmoid(mk), m.weight.cpu(), m.bias.cpu(),
                                      1, 1, 1, 1, 4)
    assert float((y - ref).abs().max()) <= 1e-4


@pytest.mark.parametrize("shape", [(8, 1, 256, 256), (3, 1, 37, 53), (2, 7, 1, 16, 20)])
def test_charbonnier_loss_matches_oracle(dev, shape):
    """CharbonnierLoss forward and backward (opt/loss.py:20-31), sum reduction and the mean_res variant; relative 1e-6 (fp32 sum)."""
    from fcvsr_b200.ops.loss import CharbonnierLoss
    g = torch.Generator().manual_seed(sum(shape))
    x, y = torch.rand(*shape, generator=g), torch.rand(*shape, generator=g)
    with torch.no_grad():
        got = CharbonnierLoss(x.to(dev), y.to(dev)).item()
        got_m = CharbonnierLoss(x.to(dev), y.to(dev), mean_res=True).item()
    ref = O.charbonnier_sum(x.double(), y.double()).item()
    d = (x - y).double().view(shape[0], -1).mean(1)
    ref_m = torch.sqrt(d * d + 1e-4).sum().item()
    assert abs(got - ref) <= 1e-6 * ref and abs(got_m - ref_m) <= 1e-6 * ref_m
    # backward (fcvsr_charbonnier_loss_backward) against autograd through the oracle expression, both variants
    for mean_res in (False, True):
        xg, yg = x.to(dev).requires_grad_(), y.to(dev).requires_grad_()
        (3.0 * CharbonnierLoss(xg, yg, mean_res)).backward()
        xr, yr = x.clone().requires_grad_(), y.clone().requires_grad_()
        if mean_res:
            dr = (xr - yr).view(shape[0], -1).mean(1, keepdim=True)
            (3.0 * torch.sqrt(dr * dr + 1e-4).sum()).backward()
        else:
            (3.0 * O.charbonnier_sum(xr, yr)).backward()
        assert float((xg.grad.cpu() - xr.grad).abs().max()) <= 1e-5 * float(xr.grad.abs().max())
        assert float((yg.grad.cpu() - yr.grad).abs().max()) <= 1e-5 * float(yr.grad.abs().max())


@pytest.mark.parametrize("variant,b,h,w", [("full", 1, 180, 320), ("S", 2, 96, 128), ("S", 1, 272, 480)])
def test_baseline_size_parity_against_oracle(dev, variant, b, h, w):
    """BASELINE config 2 at its full size (FCVSR, 7x180x320 -> 720x1280) and two more shapes the goldens do not cover
    (two-phase FFT radix pairs 15*12 / 20*16, 12*8 / 16*8, 17*16 / 24*20), all three compute modes against the CPU oracle
    (itself pinned to the reference by tests/test_oracle.py): fp32 <= 2e-5, tf32 <= 1e-3, bf16 <= 5e-3."""
    sd = arch.seeded_state_dict(variant, 0)
    x = make_clip(4321 + h, b, h, w)
    with torch.no_grad():
        ref = O.forward(sd, x)
    for mode, tol in (("fp32", 2e-5), ("tf32", 1e-3), ("bf16", 5e-3)):
        m = (arch.GShiftNet_S if variant == "S" else arch.GShiftNet)().to(dev).eval()
        m.load_state_dict(sd)
        m.compute_dtype = mode
        with torch.no_grad():
            y = m(x.to(dev)).cpu()
        assert float((y - ref).abs().max()) <= tol, mode
        del m


@pytest.mark.parametrize("B,H,W", [(1, 8, 64), (2, 19, 70), (1, 40, 200)])
def test_conv_last_to1_kernel(dev, B, H, W):
    """fcvsr_conv3x3_c64_to1 (conv_last0 + skip, bf16 input): exact on the bf16-rounded input and weights up to fp32 summation order."""
    import ctypes
    g = torch.Generator().manual_seed(H * W)
    x = torch.randn(B, 64, H, W, generator=g)
    w = torch.randn(1, 64, 3, 3, generator=g) / 24
    res = torch.randn(B, 1, H, W, generator=g)
    xd = nhwc(x).to(dev).to(torch.bfloat16)
    rd = res.to(dev).contiguous()
    y = torch.empty(B, 1, H, W, device=dev)
    wh = (ctypes.c_float * 576)(*w[0].permute(1, 2, 0).reshape(-1).tolist())
    C.call("fcvsr_conv3x3_c64_to1", xd.data_ptr(), 64, ctypes.addressof(wh), 0.25, rd.data_ptr(), y.data_ptr(), B, H, W, _st())
    torch.cuda.synchronize()
    # the kernel rounds the weights to bf16 (mma.sync operands, like every other convolution weight of the bf16 mode)
    ref = F.conv2d(x.to(torch.bfloat16).float(), w.to(torch.bfloat16).float(), torch.tensor([0.25]), padding=1) + res
    assert float((y.cpu() - ref).abs().max()) <= 2e-5 * max(1.0, float(ref.abs().max()))


def test_gshiftnet_etc_matches_per_window_forward(dev):
    """GShiftNet_ETC (CVSR_freq.py:2760-2843): 7 windows of a 13-frame clip == 7 GShiftNet forwards, x_up == bilinear x4."""
    sd = arch.seeded_state_dict("full", 2, ACNum=2, Freq_Inv=2, SCGroupN=1)
    m = arch.GShiftNet_ETC(ACNum=2, Freq_Inv=2, SCGroupN=1).to(dev).eval()
    m.load_state_dict(sd)
    ref_m = arch.GShiftNet(ACNum=2, Freq_Inv=2, SCGroupN=1).to(dev).eval()
    ref_m.load_state_dict(sd)
    g = torch.Generator().manual_seed(5)
    x = torch.rand(1, 13, 1, 16, 20, generator=g).to(dev)
    with torch.no_grad():
        out, up = m(x)
        assert out.shape == (1, 7, 1, 64, 80) and up.shape == (1, 7, 1, 64, 80)
        for i in (0, 3, 6):
            yi = ref_m(x[:, i:i + 7])
            assert float((out[:, i] - yi).abs().max()) <= 1e-5
            base = F.interpolate(x[:, i + 3].cpu(), scale_factor=4, mode="bilinear")
            assert float((up[:, i].cpu() - base).abs().max()) <= 1e-6


def test_sequence_inference_matches_oracle(dev):
    """Sliding-window driver (replicate edges, 30 -> 32 row padding and crop) against per-window oracle calls."""
    from fcvsr_b200 import sequence as S
    sd = arch.seeded_state_dict("S", 2)
    frames = make_clip(9, 1, 30, 36)[0][:5]                # 5 frames [5,1,30,36]; 30 rows -> padded to 32
    m = _model("S", sd, dev, use_tc=False)
    out, (lo, hi) = S.super_resolve_sequence(m, frames, batch=2)
    assert (lo, hi) == (0, 5) and out.shape == (5, 1, 120, 144)
    padded, _, _ = S.pad_to_multiple(frames)
    for t in (0, 2, 4):
        clip = padded[S.window_indices(t, 5)].unsqueeze(0)
        with torch.no_grad():
            ref = O.forward(sd, clip)[..., :120, :144]
        assert float((out[t].cpu() - ref[0]).abs().max()) <= 2e-5


def test_radix17_height_in_the_model(dev):
    """272-row inputs (270 padded) need the radix-17 FFT pass inside MGAA / MFFR."""
    sd = arch.seeded_state_dict("S", 4)
    x = make_clip(11, 1, 272, 16)
    with torch.no_grad():
        ref = O.forward(sd, x)
    m = _model("S", sd, dev, use_tc=False)
    with torch.no_grad():
        y = m(x.to(dev)).cpu()
    assert float((y - ref).abs().max()) <= 2e-5


def test_cuda_graph_replay_is_bit_identical(dev):
    sd = arch.seeded_state_dict("S", 0)
    x = make_clip(1234, 1, 32, 32).to(dev)
    m = _model("S", sd, dev)
    with torch.no_grad():
        y0 = m(x).clone()
        m._engine.use_graph = True
        y1 = m(x).clone()
        y2 = m(x * 0.5).clone()
        y3 = m(x).clone()
    assert torch.equal(y0, y1) and torch.equal(y1, y3) and not torch.equal(y1, y2)


# ------------------------------------------------------------------------------------------------
# bf16 operand mode (BASELINE config 2 "fp32 and bf16"): bf16 operand tensors, fp32 accumulate / residual streams.
# Tolerance stated separately from fp32: max-abs <= 5e-3 and PSNR(ours, reference) >= 60 dB.
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", [(1, 64, 64, 16, 16, 3), (2, 64, 128, 20, 36, 3), (1, 128, 64, 45, 80, 3),
                                  (1, 64, 256, 12, 20, 3), (2, 256, 128, 10, 33, 1), (1, 64, 4, 11, 19, 1),
                                  (1, 64, 1, 24, 40, 3), (1, 64, 576, 8, 16, 1)])
def test_conv_tcgen05_bf16_operands(dev, case):
    B, ci, co, H, W, k = case
    g = torch.Generator().manual_seed(sum(case))
    x = torch.randn(B, ci, H, W, generator=g)
    w = torch.randn(co, ci, k, k, generator=g) / (ci * k * k) ** 0.5
    b = torch.randn(co, generator=g)
    res = torch.randn(B, co, H, W, generator=g)
    pk = _ConvPack(w.to(dev), b.to(dev), op16=True)
    xd = nhwc(x).to(dev).to(torch.bfloat16)
    rd = nhwc(res).to(dev)
    y = torch.empty(B, H, W, co, device=dev)
    C.call("fcvsr_conv2d_tc", xd.data_ptr(), ci, pk.w_tc.data_ptr(), pk.bias.data_ptr(), rd.data_ptr(), co, 0, 0,
           y.data_ptr(), co, B, H, W, ci, co, k, 2, 0.1, 0, 0, 0, 0, 0, 0, 1, _st())
    torch.cuda.synchronize()
    xr, wr = x.to(torch.bfloat16).float(), w.to(torch.bfloat16).float()      # the kernel is exact on bf16-rounded operands
    ref = F.leaky_relu(F.conv2d(xr, wr, b, padding=k // 2), 0.1) + res
    assert float((nchw(y.cpu()) - ref).abs().max()) <= 2e-4 * max(1.0, float(ref.abs().max()))
    if co % 64 == 0:                                                        # bf16 output through pixel shuffle
        pk2 = _ConvPack(w.to(dev), b.to(dev), ps=True, op16=True)
        y2 = torch.empty(B, 2 * H, 2 * W, co // 4, device=dev, dtype=torch.bfloat16)
        C.call("fcvsr_conv2d_tc", xd.data_ptr(), ci, pk2.w_tc.data_ptr(), pk2.bias.data_ptr(), 0, 0, 0, 0,
               y2.data_ptr(), co // 4, B, H, W, ci, co, k, 0, 0.0, 0, 1, 0, 0, 1, 0, 1, _st())
        torch.cuda.synchronize()
        ref2 = F.pixel_shuffle(F.conv2d(xr, wr, b, padding=k // 2), 2)
        assert float((nchw(y2.float().cpu()) - ref2).abs().max()) <= 1e-2 * max(1.0, float(ref2.abs().max()))


@pytest.mark.parametrize("case", [(2, 128, 128, 13, 33, 1, 216, 1), (1, 128, 128, 20, 16, 1, 128, 2), (2, 64, 64, 11, 21, 1, 64, 0),
                                  (1, 64, 64, 17, 35, 3, 96, 2), (1, 256, 128, 9, 18, 1, 128, 1), (1, 64, 128, 12, 20, 3, 128, 2)])
def test_conv_tcgen05_bf16_output_tensor(dev, case):
    """bf16 operands AND a bf16 output tensor (round_out = 1): 64- and 128-channel outputs of resident-weight convolutions leave
    through the staged epilogue (SWIZZLE_128B tile in shared memory, one or two TMA stores per tile, clipped at the image border);
    the 64 -> 128 3x3 case streams its filter and keeps the direct stores.  Output rows wider than Cout (ldy), every activation,
    the res - res2 skip of MGAA.convfuse (CVSR_freq.py:1472-1473)."""
    B, ci, co, H, W, k, ldy, act = case
    g = torch.Generator().manual_seed(sum(case))
    x = torch.randn(B, ci, H, W, generator=g)
    w = torch.randn(co, ci, k, k, generator=g) / (ci * k * k) ** 0.5
    b = torch.randn(co, generator=g)
    res, res2 = torch.randn(B, co, H, W, generator=g), torch.randn(B, co, H, W, generator=g)
    pk = _ConvPack(w.to(dev), b.to(dev), op16=True)
    xd = nhwc(x).to(dev).to(torch.bfloat16)
    rd, r2d = nhwc(res).to(dev), nhwc(res2).to(dev)
    y = torch.full((B, H, W, ldy), 7.0, device=dev, dtype=torch.bfloat16)
    C.call("fcvsr_conv2d_tc", xd.data_ptr(), ci, pk.w_tc.data_ptr(), pk.bias.data_ptr(), rd.data_ptr(), co, r2d.data_ptr(), co,
           y.data_ptr(), ldy, B, H, W, ci, co, k, act, 0.2, 0, 0, 0, 0, 1, 0, 1, _st())
    torch.cuda.synchronize()
    xr, wr = x.to(torch.bfloat16).float(), w.to(torch.bfloat16).float()
    pre = F.conv2d(xr, wr, b, padding=k // 2)
    ref = {0: pre, 1: F.relu(pre), 2: F.leaky_relu(pre, 0.2)}[act] + res - res2
    got = y.float().cpu()
    assert float((nchw(got[..., :co]) - ref).abs().max()) <= (2.0 ** -8 + 2e-4) * max(1.0, float(ref.abs().max()))
    assert bool((got[..., co:] == 7.0).all())            # channels beyond Cout of a wider row are not touched


@pytest.mark.parametrize("name", ["fcvsr_s_64", "fcvsr_s_36x40", "fcvsr_full_64"])
def test_bf16_path_matches_reference_golden(dev, name):
    g = load_golden(name)
    c = g["case"]
    sd = arch.seeded_state_dict(c["variant"], c["seed"])
    x = make_clip(c["clip_seed"], c["b"], c["h"], c["w"])
    m = (arch.GShiftNet_S if c["variant"] == "S" else arch.GShiftNet)().to(dev).eval()
    m.load_state_dict(sd)
    m.compute_dtype = "bf16"
    with torch.no_grad():
        y = m(x.to(dev)).cpu()
    ref = g["out"]
    assert m._engine.mode == "bf16" and m._engine.tc_launches > 100
    err = float((y - ref).abs().max())
    mse = float(((y - ref) ** 2).mean())
    psnr_vs_ref = 10.0 * torch.log10(torch.tensor(1.0 / max(mse, 1e-20))).item()
    assert err <= 5e-3, err
    assert psnr_vs_ref >= 60.0, psnr_vs_ref


def _bf16r(t):
    return t.to(torch.bfloat16).float()


@pytest.mark.parametrize("B,H,W", [(1, 8, 12), (2, 18, 22)])
def test_subsample2_kernel(dev, B, H, W):
    """fcvsr_subsample2: y[b,i,j,:] = x[b,2i,2j,:] (fp32 stream) and its bf16 / TF32-rounded operand copy -- with the stride-1
    tcgen05 convolution in front this is rconcat1/2 (CVSR_freq.py:2671-2672): checked against F.conv2d(stride=2)."""
    g = torch.Generator().manual_seed(B * H * W)
    x = torch.randn(B, 64, H, W, generator=g)
    xd = nhwc(x).to(dev)
    y = torch.zeros(B, H // 2, W // 2, 64, device=dev)
    y2 = torch.zeros(B, H // 2, W // 2, 64, device=dev, dtype=torch.bfloat16)
    C.call("fcvsr_subsample2", xd.data_ptr(), 64, y.data_ptr(), 64, y2.data_ptr(), 64, B, H, W, 64, 1, _st())
    torch.cuda.synchronize()
    ref = nhwc(x)[:, ::2, ::2, :]
    assert torch.equal(y.cpu(), ref)
    assert torch.equal(y2.float().cpu(), _bf16r(ref))
    w = torch.randn(64, 64, 3, 3, generator=g) / 24
    full = F.conv2d(x, w, padding=1)
    assert torch.allclose(full[:, :, ::2, ::2], F.conv2d(x, w, padding=1, stride=2), atol=1e-5)


@pytest.mark.parametrize("b16", [0, 1])
def test_scnet_helpers_bf16_inputs(dev, b16):
    """ContextBlock / RCB tail / cross-level mix (CVSR_freq.py:657-777) through the C ABI with fp32 and with bf16 side tensors
    (the bf16 mode's res, r0, rr, td, tu): exact formulas on the (rounded) inputs, fp32 arithmetic."""
    g = torch.Generator().manual_seed(17 + b16)
    B, H, W = 2, 12, 20
    P = H * W
    rnd = _bf16r if b16 else (lambda t: t)
    res, r0 = rnd(torch.randn(B, P, 64, generator=g)), rnd(torch.randn(B, P, 64, generator=g))
    wm = torch.randn(64, generator=g) / 8
    w1, w2 = torch.randn(64, 64, generator=g) / 8, torch.randn(64, 64, generator=g) / 8
    dt = torch.bfloat16 if b16 else torch.float32
    res_d, r0_d = res.to(dev).to(dt), r0.to(dev).to(dt)
    nblk = (P + 127) // 128
    part = torch.zeros(B * nblk * 66, device=dev)
    add = torch.zeros(B, 64, device=dev)
    wm_d, w1_d, w2_d = wm.to(dev), w1.to(dev), w2.to(dev)
    cnt = torch.zeros(B, device=dev, dtype=torch.int32)
    for _ in range(2):          # twice: the block counters reset themselves
        C.call("fcvsr_context_block", res_d.data_ptr(), 64, wm_d.data_ptr(), w1_d.data_ptr(), w2_d.data_ptr(),
               part.data_ptr(), add.data_ptr(), cnt.data_ptr(), B, P, b16, _st())
    assert int(cnt.abs().sum()) == 0
    att = torch.softmax(res @ wm, dim=1)                                   # [B,P]
    ctx = torch.einsum("bp,bpc->bc", att, res)
    add_ref = F.leaky_relu(ctx @ w1.t(), 0.2) @ w2.t()
    torch.cuda.synchronize()
    assert float((add.cpu() - add_ref).abs().max()) <= 1e-4 * max(1.0, float(add_ref.abs().max()))
    # RCB tail: rr = lrelu_0.2(res + add) + r0, with the 2x2 mean; fp32 output and / or the bf16 operand copy
    rr = torch.zeros(B, P, 64, device=dev)
    rrh = torch.zeros(B, P, 64, device=dev, dtype=torch.bfloat16)
    rrp = torch.zeros(B, P // 4, 64, device=dev, dtype=torch.bfloat16)
    add_d = add_ref.to(dev)
    C.call("fcvsr_rcb_finish", res_d.data_ptr(), add_d.data_ptr(), r0_d.data_ptr(), 0 if b16 else rr.data_ptr(), B, P,
           rrh.data_ptr(), 1, rrp.data_ptr(), H, W, 0, 3 * b16, _st())
    rr_ref = F.leaky_relu(res + add_ref[:, None, :], 0.2) + r0
    pool_ref = rr_ref.view(B, H // 2, 2, W // 2, 2, 64).mean(dim=(2, 4)).reshape(B, P // 4, 64)
    torch.cuda.synchronize()
    if not b16:
        assert float((rr.cpu() - rr_ref).abs().max()) <= 1e-5
    assert float((rrh.float().cpu() - rr_ref).abs().max()) <= 2 ** -8 * float(rr_ref.abs().max())
    assert float((rrp.float().cpu() - pool_ref).abs().max()) <= 2 ** -8 * float(pool_ref.abs().max())
    # cross-level mix at the middle level: x + 1.0 * rr + td (already pooled) + bilinear_x2(tu)
    xin = torch.randn(B, P, 64, generator=g)
    rrv = rnd(rr_ref)
    td = rnd(torch.randn(B, P, 64, generator=g))
    tu = rnd(torch.randn(B, (H // 2) * (W // 2), 64, generator=g))
    out = torch.zeros(B, P, 64, device=dev)
    out_r = torch.zeros(B, P, 64, device=dev, dtype=torch.bfloat16)
    xin_d, rr_d, td_d, tu_d = xin.to(dev), rrv.to(dev).to(dt), td.to(dev).to(dt), tu.to(dev).to(dt)
    C.call("fcvsr_level_mix", xin_d.data_ptr(), 64, out.data_ptr(), 64, rr_d.data_ptr(), 1.0, td_d.data_ptr(), tu_d.data_ptr(),
           B, H, W, out_r.data_ptr(), 64, 0, 1, 1 + 6 * b16, _st())
    up = F.interpolate(tu.view(B, H // 2, W // 2, 64).permute(0, 3, 1, 2), scale_factor=2, mode="bilinear")
    ref = xin + rrv + td + up.permute(0, 2, 3, 1).reshape(B, P, 64)
    torch.cuda.synchronize()
    assert float((out.cpu() - ref).abs().max()) <= 1e-5 * max(1.0, float(ref.abs().max()))
    assert float((out_r.float().cpu() - ref).abs().max()) <= 2 ** -8 * float(ref.abs().max())
    if b16:     # x carried as the bf16 operand copy, in place, no fp32 output (td_pooled & 8; blocks 2 and 3 of a group)
        xr = rnd(xin)
        x16 = xr.to(dev).to(dt)
        C.call("fcvsr_level_mix", x16.data_ptr(), 64, 0, 64, rr_d.data_ptr(), 1.0, td_d.data_ptr(), tu_d.data_ptr(),
               B, H, W, x16.data_ptr(), 64, 0, 1, 15, _st())
        ref16 = xr + rrv + td + up.permute(0, 2, 3, 1).reshape(B, P, 64)
        torch.cuda.synchronize()
        assert float((x16.float().cpu() - ref16).abs().max()) <= 2 ** -8 * float(ref16.abs().max())
        # fp32 xin, only the operand copy written: the eight-channel kernel, bit-identical to the four-channel one above
        out_r8 = torch.zeros_like(out_r)
        C.call("fcvsr_level_mix", xin_d.data_ptr(), 64, 0, 64, rr_d.data_ptr(), 1.0, td_d.data_ptr(), tu_d.data_ptr(),
               B, H, W, out_r8.data_ptr(), 64, 0, 1, 7, _st())
        torch.cuda.synchronize()
        assert torch.equal(out_r8, out_r)
        # a bf16 xin without bf16 r / td / tu is refused, and so is a call without any output
        assert C.try_call("fcvsr_level_mix", x16.data_ptr(), 64, 0, 64, rr_d.data_ptr(), 1.0, td_d.data_ptr(), tu_d.data_ptr(),
                          B, H, W, x16.data_ptr(), 64, 0, 1, 9, _st()) == C.ERR_UNSUPPORTED
        assert C.try_call("fcvsr_level_mix", x16.data_ptr(), 64, 0, 64, rr_d.data_ptr(), 1.0, td_d.data_ptr(), tu_d.data_ptr(),
                          B, H, W, 0, 64, 0, 1, 15, _st()) == C.ERR_ARG


@pytest.mark.parametrize("kind,cin,cout,stride", [("v2", 6, 4, 1), ("v2", 8, 8, 1), ("v1", 8, 6, 2)])
def test_dcn_pack_modules_train(dev, monkeypatch, kind, cin, cout, stride):
    """DeformConvPack / ModulatedDeformConvPack (deform_conv.py:239-261,:311-337) under autograd: gradients of the input, the DCN
    weight / bias and the conv_offset[_mask] layer (its backward = the DCN backward entry with zero offsets) against autograd
    through F.conv2d + the oracle DCN on the CPU."""
    import fcvsr_b200.ops.dcn as dcn_mod
    monkeypatch.setattr(dcn_mod, "PRECISION", "fp32")
    torch.manual_seed(11 + cin)
    dg = 2
    if kind == "v2":
        m = dcn_mod.ModulatedDeformConvPack(cin, cout, 3, stride=stride, padding=1, deformable_groups=dg).to(dev)
        off_layer = m.conv_offset_mask
    else:
        m = dcn_mod.DeformConvPack(cin, cout, 3, stride=stride, padding=1, deformable_groups=dg).to(dev)
        off_layer = m.conv_offset
    with torch.no_grad():
        off_layer.weight.normal_(0, 0.08)
        off_layer.bias.normal_(0, 0.4)
        if kind == "v2":
            m.bias.normal_(0, 0.1)
    x = torch.randn(2, cin, 10, 12, device=dev, requires_grad=True)
    y = m(x)
    gy = torch.randn_like(y)
    (y * gy).sum().backward()
    torch.cuda.synchronize()
    # CPU reference
    xr = x.detach().cpu().requires_grad_()
    w = m.weight.detach().cpu().requires_grad_()
    cw, cb = off_layer.weight.detach().cpu().requires_grad_(), off_layer.bias.detach().cpu().requires_grad_()
    o = F.conv2d(xr, cw, cb, stride=stride, padding=1)
    if kind == "v2":
        b = m.bias.detach().cpu().requires_grad_()
        o1, o2, mk = torch.chunk(o, 3, dim=1)
        ref = O.modulated_deform_conv(xr, torch.cat((o1, o2), 1), torch.sigmoid(mk), w, b, stride, 1, 1, 1, dg)
    else:
        ref = O.modulated_deform_conv(xr, o, None, w, None, stride, 1, 1, 1, dg)
    assert float((y.detach().cpu() - ref.detach()).abs().max()) <= 1e-4
    (ref * gy.cpu()).sum().backward()
    pairs = [("input", x.grad, xr.grad), ("weight", m.weight.grad, w.grad), ("conv_offset.weight", off_layer.weight.grad, cw.grad),
             ("conv_offset.bias", off_layer.bias.grad, cb.grad)]
    if kind == "v2":
        pairs.append(("bias", m.bias.grad, b.grad))
    for name, a, r in pairs:
        assert a is not None, name
        err = float((a.cpu() - r).abs().max())
        assert err <= 3e-4 * max(1.0, float(r.abs().max())), (name, err)


def test_adam_step_matches_torch_adam(dev):
    """fcvsr_adam_step (multi-tensor, csrc/optim.cu) behind ops.optim.Adam against torch.optim.Adam with the reference's
    settings (train_LD_freqCVSR_22.py:204: lr 5e-6 scaled up here so that the updates are visible, weight_decay 1e-5), over 70
    tensors (two launches) of ragged sizes incl. unaligned views and a parameter without gradient, 3 steps + an lr change."""
    from fcvsr_b200.ops.optim import Adam
    g = torch.Generator().manual_seed(9)
    shapes = [(64, 64, 3, 3), (1152, 64, 1, 1), (64,), (1,), (4, 4, 11, 11), (7, 3), (4097,)] * 10
    ours = [torch.nn.Parameter(torch.randn(*s, generator=g).to(dev)) for s in shapes]
    refs = [torch.nn.Parameter(p.detach().clone()) for p in ours]
    opt = Adam(ours, lr=2e-3, weight_decay=1e-5)
    ref = torch.optim.Adam(refs, lr=2e-3, weight_decay=1e-5, foreach=False, fused=False)
    sched = torch.optim.lr_scheduler.MultiStepLR(opt, milestones=[2], gamma=0.25)
    sched_ref = torch.optim.lr_scheduler.MultiStepLR(ref, milestones=[2], gamma=0.25)
    for step in range(3):
        for i, (p, r) in enumerate(zip(ours, refs)):
            if i == 5:                       # never gets a gradient: skipped by both
                continue
            gr = torch.randn(p.shape, generator=g).to(dev)
            p.grad, r.grad = gr.clone(), gr.clone()
        opt.step()
        ref.step()
        sched.step()
        sched_ref.step()
    torch.cuda.synchronize()
    for i, (p, r) in enumerate(zip(ours, refs)):
        assert float((p.detach() - r.detach()).abs().max()) <= 2e-6 * max(1.0, float(r.detach().abs().max())), i
        if i != 5:
            assert torch.allclose(opt.state[p]["exp_avg_sq"], ref.state[r]["exp_avg_sq"], rtol=2e-6, atol=1e-12)
            assert torch.allclose(opt.state[p]["exp_avg"], ref.state[r]["exp_avg"], rtol=2e-6, atol=1e-6)
    assert opt.param_groups[0]["lr"] == ref.param_groups[0]["lr"] == 5e-4


def test_cuda_train_step_on_dcn_pack_module(dev, monkeypatch):
    """fcvsr_b200.train.train_step with every piece on this library's kernels -- ModulatedDeformConvPack forward / backward (DCN
    + conv_offset_mask), CharbonnierLoss forward / backward, multi-tensor Adam -- against a CPU twin built from F.conv2d, the
    oracle DCN and torch.optim.Adam: same loss trajectory over 4 steps, and the loss goes down."""
    import fcvsr_b200.ops.dcn as dcn_mod
    from fcvsr_b200.ops.loss import CharbonnierLoss
    from fcvsr_b200.ops.optim import Adam
    from fcvsr_b200.train import train_step
    monkeypatch.setattr(dcn_mod, "PRECISION", "fp32")
    torch.manual_seed(21)
    m = dcn_mod.ModulatedDeformConvPack(8, 8, 3, stride=1, padding=1, deformable_groups=2).to(dev)
    with torch.no_grad():
        m.conv_offset_mask.weight.normal_(0, 0.05)
        m.conv_offset_mask.bias.normal_(0, 0.3)
    twin = {k: v.detach().cpu().clone().requires_grad_() for k, v in m.named_parameters()}
    opt = Adam(m.parameters(), lr=2e-3, weight_decay=1e-5)
    opt_twin = torch.optim.Adam(list(twin.values()), lr=2e-3, weight_decay=1e-5)
    g = torch.Generator().manual_seed(4)
    x = torch.randn(2, 8, 12, 14, generator=g)
    hr = torch.randn(2, 8, 12, 14, generator=g)
    xd, hd = x.to(dev), hr.to(dev)
    losses, losses_twin = [], []
    for _ in range(4):
        losses.append(float(train_step(m, opt, xd, hd, CharbonnierLoss)))
        opt_twin.zero_grad(set_to_none=True)
        o = F.conv2d(x, twin["conv_offset_mask.weight"], twin["conv_offset_mask.bias"], padding=1)
        o1, o2, mk = torch.chunk(o, 3, dim=1)
        y = O.modulated_deform_conv(x, torch.cat((o1, o2), 1), torch.sigmoid(mk), twin["weight"], twin["bias"], 1, 1, 1, 1, 2)
        lt = O.charbonnier_sum(y, hr)
        lt.backward()
        opt_twin.step()
        losses_twin.append(float(lt.detach()))
    assert losses[-1] < losses[0]
    for a, b in zip(losses, losses_twin):
        assert abs(a - b) <= 2e-3 * abs(b), (losses, losses_twin)
