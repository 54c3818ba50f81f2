"""GPU parity tests of single hot-path kernels against the oracle's block functions (run on the B200 box).

Round 1 covered `corr_gather`, `offset_blocks`, `iac_step` and the DivEnh chain only through whole-model stage taps at
64x64 / 36x40.  Here each kernel is called directly through the C ABI on seeded inputs at ragged sizes (tile remainders in
both directions), with every storage flavour the engine uses (fp32 / TF32-rounded / bf16 tensors, fp16 filter taps), and
compared with the matching function of oracle/fcvsr_oracle.py (itself pinned to the live reference in tests/test_oracle.py).
Also: FCVSR-S at BASELINE config 5b's 540x960 (FFT radix pairs 27*20 and 32*30), and the DCN backward on the tensors the
tcgen05 forward saved.
"""
import math

import pytest
import torch
import torch.nn.functional as F

from fcvsr_b200 import _capi as C, arch
from fcvsr_b200.engine import Engine, _round_tf32 as _tf32r
from oracle import fcvsr_oracle as O
from tests.util import make_clip, nchw, nhwc, psnr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev(lib):
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


def _st():
    return torch.cuda.current_stream().cuda_stream


def _bf16r(t):
    return t.to(torch.bfloat16).float()


# ------------------------------------------------------------------------------------------------------------------------
# CorrBlock lookup (CVSR_freq.py:1279-1337)
# ------------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,H,W,mode", [(1, 12, 20, 0), (2, 70, 18, 0), (2, 9, 14, 1), (1, 70, 12, 2), (2, 70, 18, 4), (2, 9, 14, 5),
                                        (1, 70, 12, 6), (2, 24, 36, 6)])
def test_corr_gather_kernel(dev, B, H, W, mode):
    """fcvsr_corr_gather on an interleaved spectrum against oracle.corr_lookup on the reference's cat([imag, real]) packing;
    H = 70 > 68 exercises the rows where the 64x2 'image' reinterpretation runs out; op_mode 1 / 2 store
    TF32-rounded / bf16 values (the engine's tensor-core operand flavours)."""
    Wf = W // 2 + 1
    g = torch.Generator().manual_seed(B * 100 + H + mode)
    z1 = torch.complex(torch.randn(B, 64, H, Wf, generator=g), torch.randn(B, 64, H, Wf, generator=g))
    z2 = torch.complex(torch.randn(B, 64, H, Wf, generator=g), torch.randn(B, 64, H, Wf, generator=g))
    ref = O.corr_lookup(torch.cat([z1.imag, z1.real], 1), torch.cat([z2.imag, z2.real], 1))        # [B,81,H,Wf]
    S = torch.zeros(B, H * Wf, 384)
    for k, z in ((0, z1), (1, z2)):
        zz = z.permute(0, 2, 3, 1).reshape(B, H * Wf, 64)
        S[:, :, 128 * k:128 * k + 128:2] = zz.real
        S[:, :, 128 * k + 1:128 * k + 128:2] = zz.imag
    Sd = S.to(dev)
    ldo = 96
    out = torch.zeros(B, H * Wf, ldo, device=dev, dtype=torch.bfloat16 if (mode & 3) == 2 else torch.float32)
    C.call("fcvsr_corr_gather", Sd.data_ptr(), 384, 0, 128, out.data_ptr() + 8 * out.element_size(), ldo, B, H, Wf, 128, mode, _st())
    torch.cuda.synchronize()
    got = out.float().cpu()[:, :, 8:89].reshape(B, H, Wf, 81).permute(0, 3, 1, 2)
    tol = {0: 1e-6, 1: 2.0 ** -11, 2: 2.0 ** -8}[mode & 3] * max(1.0, float(ref.abs().max()))
    assert float((got - ref).abs().max()) <= tol
    # mode + 4 (vector stores) zero-fills the padding channels behind the 81 real ones; nothing else is touched
    assert float(out.float().cpu()[:, :, :8].abs().max()) == 0.0 and float(out.float().cpu()[:, :, 89:].abs().max()) == 0.0


# ------------------------------------------------------------------------------------------------------------------------
# ConvBlk x ACNum on the 4-channel frequency maps (CVSR_freq.py:344-357, :1494-1498)
# ------------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("variant,B,H,W", [("full", 1, 20, 24), ("S", 2, 13, 30), ("full", 2, 9, 10)])
def test_offset_blocks_kernel(dev, variant, B, H, W):
    """fcvsr_offset_blocks (all iterations, both directions, kernel sizes 1..11) against oracle.conv_blk * sim, incl. maps
    smaller than the largest kernel (9x6 bins under an 11x11 filter) and quad / block remainders."""
    Wf = W // 2 + 1
    P = H * Wf
    sd = arch.seeded_state_dict(variant, 0)
    m = (arch.GShiftNet if variant == "full" else arch.GShiftNet_S)().to(dev).eval()
    m.load_state_dict(sd)
    eng = Engine(m, mode="fp32")
    eng._ensure_packs(dev)
    Pk, A = eng.packs, m.ACNum
    g = torch.Generator().manual_seed(17 + H)
    off = torch.randn(2, B, 4, H, Wf, generator=g)
    sim = torch.randn(B, 4, H, Wf, generator=g)
    offd = off.permute(0, 1, 3, 4, 2).contiguous().to(dev)                 # [2][B][P][4]
    simd = nhwc(sim).to(dev)
    t1 = torch.empty(A, 2 * B, P, 4, device=dev)
    t2 = torch.empty_like(t1)
    part = torch.empty(A * 2 * B * max((P + 127) // 128, ((Wf + 63) // 64) * ((H + 7) // 8)) * 4, device=dev)
    z = torch.empty(B, P, 8 * A, device=dev)
    C.call("fcvsr_offset_blocks", offd.data_ptr(), Pk["ob_w1"].data_ptr(), Pk["ob_w2"].data_ptr(), Pk["ob_prelu"].data_ptr(),
           Pk["ob_ca"].data_ptr(), simd.data_ptr(), 4, t1.data_ptr(), t2.data_ptr(), part.data_ptr(), z.data_ptr(), B, H, Wf, A,
           _st())
    torch.cuda.synchronize()
    zc = z.cpu().view(B, H, Wf, 2 * A, 2, 2)                               # [..., it*2+dir, m, (re, im)]
    for i in range(A):
        for d in range(2):
            o = O.conv_blk(sd, f"MGAA.MConvB.{i}", off[d]) * sim           # [B,4,H,Wf]; complex(o[0:2], o[2:4]) (:1497-1498)
            want = torch.stack([o[:, 0:2], o[:, 2:4]], -1).permute(0, 2, 3, 1, 4)       # [B,H,Wf,m,(re,im)]
            err = float((zc[:, :, :, i * 2 + d] - want).abs().max())
            assert err <= 2e-5 * max(1.0, float(want.abs().max())), (i, d, err)


# ------------------------------------------------------------------------------------------------------------------------
# IAC iteration (flow_warp + SAC + residual + LeakyReLU, CVSR_freq.py:1188-1276)
# ------------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,H,W,taps_half,prev16,ro", [(1, 19, 37, 0, 0, 0), (2, 8, 16, 1, 0, 0), (1, 25, 18, 1, 0, 1),
                                                      (1, 19, 37, 1, 1, 2), (2, 12, 20, 1, 1, 0), (1, 33, 35, 1, 0, 2)])
def test_iac_step_kernel(dev, B, H, W, taps_half, prev16, ro):
    """fcvsr_iac_step for both directions against oracle.warp_bilinear + oracle.sac (+ x_in, LeakyReLU 0.1) with offsets that
    leave the image (zero fill), ragged 8x16 tiles, fp16 taps (the tensor-core modes), bf16 ping-pong inputs (flag 4) and
    TF32-rounded / bf16 outputs -- every flavour engine._mgaa uses."""
    g = torch.Generator().manual_seed(H * W + taps_half + 2 * prev16 + 4 * ro)
    prev = [torch.randn(B, 64, H, W, generator=g) for _ in range(2)]
    xin = [torch.randn(B, 64, H, W, generator=g) for _ in range(2)]
    off = [3.0 * torch.randn(B, 2, H, W, generator=g) for _ in range(2)]
    off[0][:, :, 0, 0] = torch.tensor([-40.0, 55.0])                      # far outside
    taps = 0.6 * torch.randn(B, 64, 3, H, W, generator=g)                 # reference channel c*3 + t
    if taps_half:
        taps = taps.half().float()
    if prev16:
        prev = [_bf16r(p) for p in prev]
    want = [F.leaky_relu(O.sac(O.warp_bilinear(prev[d], off[d]), taps.reshape(B, 192, H, W)) + xin[d], 0.1) for d in range(2)]
    pdt = torch.bfloat16 if prev16 else torch.float32
    prev_d = [nhwc(p).to(dev).to(pdt) for p in prev]
    xin_d = [nhwc(t).to(dev) for t in xin]
    offs_d = torch.zeros(B, H, W, 8, device=dev)
    offs_d[..., 2:4] = nhwc(off[0]).to(dev)
    offs_d[..., 6:8] = nhwc(off[1]).to(dev)
    taps_d = taps.permute(0, 3, 4, 2, 1).reshape(B, H, W, 192).contiguous().to(dev)      # ours: [t][c]
    if taps_half:
        taps_d = taps_d.half()
    odt = torch.bfloat16 if ro == 2 else torch.float32
    nxt = [torch.empty(B, H, W, 64, device=dev, dtype=odt) for _ in range(2)]
    C.call("fcvsr_iac_step", prev_d[0].data_ptr(), 64, prev_d[1].data_ptr(), 64, xin_d[0].data_ptr(), 64, xin_d[1].data_ptr(), 64,
           nxt[0].data_ptr(), 64, nxt[1].data_ptr(), 64, offs_d.data_ptr(), 8, 2, 6, taps_d.data_ptr(), 192, taps_half, B, H, W,
           ro | (4 if prev16 else 0), _st())
    torch.cuda.synchronize()
    for d in range(2):
        got = nchw(nxt[d].float().cpu())
        scale = max(1.0, float(want[d].abs().max()))
        tol = {0: 2e-5, 1: 2.0 ** -11 + 2e-5, 2: 2.0 ** -8 + 2e-5}[ro] * scale
        assert float((got - want[d]).abs().max()) <= tol, d


@pytest.mark.parametrize("B,H,W,prev16", [(1, 19, 37, 0), (2, 8, 14, 1), (1, 25, 18, 1), (2, 33, 45, 0), (1, 7, 13, 1)])
def test_iac_step_tc_kernel(dev, B, H, W, prev16):
    """fcvsr_iac_step_tc (taps = MGAA.F.1 slice of kp2 on tcgen05 inside the IAC kernel, CVSR_freq.py:1522-1527) for both
    directions against oracle.warp_bilinear + oracle.sac with taps from an fp32 matmul of the same bf16 operands: ragged
    8x14 tiles, images smaller than a tile, offsets leaving the image, fp32 (iteration 0) and bf16 (ping-pong) inputs."""
    g = torch.Generator().manual_seed(H * W + prev16)
    prev = [torch.randn(B, 64, H, W, generator=g) for _ in range(2)]
    xin = [torch.randn(B, 64, H, W, generator=g) for _ in range(2)]
    off = [3.0 * torch.randn(B, 2, H, W, generator=g) for _ in range(2)]
    off[1][:, :, H - 1, W - 1] = torch.tensor([70.0, -31.0])
    kp = _bf16r(torch.randn(B, 64, H, W, generator=g))
    wt = _bf16r(0.1 * torch.randn(64, 3, 64, generator=g))               # [c][t][k]: reference row c*3 + t
    bias = 0.2 * torch.randn(64, 3, generator=g)
    taps = torch.einsum("ctk,bkhw->bcthw", wt, kp) + bias[None, :, :, None, None]        # [B,64,3,H,W]
    if prev16:
        prev = [_bf16r(p) for p in prev]
    want = [F.leaky_relu(O.sac(O.warp_bilinear(prev[d], off[d]), taps.reshape(B, 192, H, W)) + xin[d], 0.1) for d in range(2)]
    pdt = torch.bfloat16 if prev16 else torch.float32
    prev_d = [nhwc(p).to(dev).to(pdt) for p in prev]
    xin_d = [nhwc(t).to(dev) for t in xin]
    offs_d = torch.zeros(B, H, W, 8, device=dev)
    offs_d[..., 2:4] = nhwc(off[0]).to(dev)
    offs_d[..., 6:8] = nhwc(off[1]).to(dev)
    kp_d = nhwc(kp).to(dev).to(torch.bfloat16)
    # kernel row order c4*12 + t*4 + cc, channel c = 4 c4 + cc
    w_d = wt.view(16, 4, 3, 64).permute(0, 2, 1, 3).reshape(192, 64).contiguous().to(dev).to(torch.bfloat16)
    b_d = bias.view(16, 4, 3).permute(0, 2, 1).reshape(192).contiguous().to(dev)
    nxt = [torch.empty(B, H, W, 64, device=dev, dtype=torch.bfloat16) for _ in range(2)]
    C.call("fcvsr_iac_step_tc", prev_d[0].data_ptr(), 64, prev_d[1].data_ptr(), 64, prev16, xin_d[0].data_ptr(), 64,
           xin_d[1].data_ptr(), 64, nxt[0].data_ptr(), 64, nxt[1].data_ptr(), 64, offs_d.data_ptr(), 8, 2, 6, kp_d.data_ptr(), 64,
           w_d.data_ptr(), b_d.data_ptr(), B, H, W, _st())
    torch.cuda.synchronize()
    for d in range(2):
        got = nchw(nxt[d].float().cpu())
        scale = max(1.0, float(want[d].abs().max()))
        # bf16 output rounding; the taps stay fp32 (TMEM accumulator) on the device
        assert float((got - want[d]).abs().max()) <= (2.0 ** -8 + 2e-4) * scale, d


@pytest.mark.parametrize("B,H,W,rnd", [(1, 19, 37, 0), (2, 8, 14, 4), (2, 33, 45, 0), (1, 7, 13, 4)])
def test_iac_step_tc_kernel_tf32(dev, B, H, W, rnd):
    """fcvsr_iac_step_tc in the fp32-contract mode (prev16 = 2 [+ 4]): TF32-rounded fp32 kp2 / F.1 slice on kind::tf32 MMAs, fp32
    prev / next, against the oracle with taps from an fp32 matmul of the same rounded operands."""
    g = torch.Generator().manual_seed(H * W + 5)
    prev = [torch.randn(B, 64, H, W, generator=g) for _ in range(2)]
    xin = [torch.randn(B, 64, H, W, generator=g) for _ in range(2)]
    off = [3.0 * torch.randn(B, 2, H, W, generator=g) for _ in range(2)]
    off[0][:, :, 0, 0] = torch.tensor([-55.0, 12.0])
    kp = _tf32r(torch.randn(B, 64, H, W, generator=g))
    wt = _tf32r(0.1 * torch.randn(64, 3, 64, generator=g))               # [c][t][k]: reference row c*3 + t
    bias = 0.2 * torch.randn(64, 3, generator=g)
    taps = torch.einsum("ctk,bkhw->bcthw", wt, kp) + bias[None, :, :, None, None]
    want = [F.leaky_relu(O.sac(O.warp_bilinear(prev[d], off[d]), taps.reshape(B, 192, H, W)) + xin[d], 0.1) for d in range(2)]
    prev_d = [nhwc(p).to(dev) for p in prev]
    xin_d = [nhwc(t).to(dev) for t in xin]
    offs_d = torch.zeros(B, H, W, 8, device=dev)
    offs_d[..., 2:4] = nhwc(off[0]).to(dev)
    offs_d[..., 6:8] = nhwc(off[1]).to(dev)
    kp_d = nhwc(kp).to(dev)
    w_d = wt.view(16, 4, 3, 64).permute(0, 2, 1, 3).reshape(192, 64).contiguous().to(dev)
    b_d = bias.view(16, 4, 3).permute(0, 2, 1).reshape(192).contiguous().to(dev)
    nxt = torch.zeros(B, H, W, 128, device=dev)                         # both directions side by side (ld 128), as cat128
    C.call("fcvsr_iac_step_tc", prev_d[0].data_ptr(), 64, prev_d[1].data_ptr(), 64, 2 + rnd, xin_d[0].data_ptr(), 64,
           xin_d[1].data_ptr(), 64, nxt.data_ptr(), 128, nxt.data_ptr() + 64 * 4, 128, offs_d.data_ptr(), 8, 2, 6, kp_d.data_ptr(), 64,
           w_d.data_ptr(), b_d.data_ptr(), B, H, W, _st())
    torch.cuda.synchronize()
    for d in range(2):
        got = nchw(nxt[..., 64 * d:64 * d + 64].cpu())
        scale = max(1.0, float(want[d].abs().max()))
        assert float((got - want[d]).abs().max()) <= ((2.0 ** -11 if rnd else 0.0) + 1e-4) * scale, d
    # a bf16 prev together with TF32 operands is refused
    assert C.try_call("fcvsr_iac_step_tc", prev_d[0].data_ptr(), 64, prev_d[1].data_ptr(), 64, 3, xin_d[0].data_ptr(), 64,
                      xin_d[1].data_ptr(), 64, nxt.data_ptr(), 128, nxt.data_ptr() + 256, 128, offs_d.data_ptr(), 8, 2, 6,
                      kp_d.data_ptr(), 64, w_d.data_ptr(), b_d.data_ptr(), B, H, W, _st()) == C.ERR_ARG


# ------------------------------------------------------------------------------------------------------------------------
# DivEnh chain + CALayer gates (CVSR_freq.py:2104-2133, :1812-1828, :2201-2254)
# ------------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("variant,B,H,W", [("S", 2, 12, 20), ("full", 1, 19, 27), ("full", 3, 16, 16)])
def test_divenh_chain_kernels(dev, variant, B, H, W):
    """chansum64 / divenh_step / reduce_finalize / mffr_final driven exactly as engine._mffr drives them, on band tensors taken
    from the oracle's split_freq (so the FFT kernels are not in the loop): the running sum of the enhanced bands and the
    block output against oracle.div_enh / oracle.mffr; P not a multiple of the 256-pixel block, Freq_Inv 4 and 8."""
    sd = arch.seeded_state_dict(variant, 0)
    m = (arch.GShiftNet if variant == "full" else arch.GShiftNet_S)().to(dev).eval()
    m.load_state_dict(sd)
    eng = Engine(m, mode="fp32")
    eng._ensure_packs(dev)
    Pk, Q = eng.packs, m.Freq_Inv
    g = torch.Generator().manual_seed(3 + H)
    x = torch.randn(B, 64, H, W, generator=g)
    bands = O.split_freq(x, Q)[::-1]
    outs = []
    for i in range(Q):
        outs.append(O.div_enh(sd, f"MFFRblock.DivEnh_block.{i}", bands[i], bands[:i], outs[:i]))
    so_ref = torch.stack(outs, 0).sum(0)
    y_ref = O.mffr(sd, x, Q)
    npix = H * W
    nblk = (npix + 255) // 256
    bd = [nhwc(b).to(dev) for b in bands]
    xd = nhwc(x).to(dev)
    sb = torch.empty(B, npix, 64, device=dev)
    so = torch.empty(B, npix, 64, device=dev)
    part = torch.empty(B * nblk * 128, device=dev)
    mean0 = torch.empty(B, 128, device=dev)
    gates = torch.empty(Q + 1, B, 128, device=dev)
    y = torch.empty(B, npix, 64, device=dev)
    st = _st()
    de = lambda i, s: Pk[f"de{i}.{s}"].data_ptr()  # noqa: E731
    gate = lambda i: gates.data_ptr() + i * B * 128 * 4  # noqa: E731
    inv = 1.0 / npix
    C.call("fcvsr_chansum64", bd[0].data_ptr(), 64, part.data_ptr(), B, npix, st)
    C.call("fcvsr_reduce_finalize", part.data_ptr(), nblk, 1, inv, 0, 0, 0, mean0.data_ptr(), B, st)
    C.call("fcvsr_divenh_step", 0, 0, 0, 0, 0, 0, 1, 1, bd[0].data_ptr(), de(0, "a"), de(0, "b"), mean0.data_ptr(), sb.data_ptr(),
           so.data_ptr(), part.data_ptr(), B, npix, st)
    C.call("fcvsr_reduce_finalize", part.data_ptr(), nblk, 1, inv, 1, de(0, "w1"), de(0, "w2"), gate(0), B, st)
    for i in range(1, Q):
        C.call("fcvsr_divenh_step", bd[i - 1].data_ptr(), de(i - 1, "a"), de(i - 1, "b"), mean0.data_ptr(), gate(i - 1),
               int(i - 1 == 0), 1, 0, bd[i].data_ptr(), de(i, "a"), de(i, "b"), 0, sb.data_ptr(), so.data_ptr(), part.data_ptr(),
               B, npix, st)
        C.call("fcvsr_reduce_finalize", part.data_ptr(), nblk, 2, inv, 1, de(i, "w1"), de(i, "w2"), gate(i), B, st)
    C.call("fcvsr_divenh_step", bd[Q - 1].data_ptr(), de(Q - 1, "a"), de(Q - 1, "b"), mean0.data_ptr(), gate(Q - 1), int(Q == 1), 2,
           0, 0, 0, 0, 0, sb.data_ptr(), so.data_ptr(), part.data_ptr(), B, npix, st)
    C.call("fcvsr_reduce_finalize", part.data_ptr(), nblk, 1, inv, 1, Pk["mffr.w1"].data_ptr(), Pk["mffr.w2"].data_ptr(), gate(Q), B, st)
    C.call("fcvsr_mffr_final", so.data_ptr(), gate(Q), xd.data_ptr(), 64, y.data_ptr(), 64, B, npix, st)
    torch.cuda.synchronize()
    so_got = nchw(so.cpu().view(B, H, W, 64))
    y_got = nchw(y.cpu().view(B, H, W, 64))
    assert float((so_got - so_ref).abs().max()) <= 2e-5 * max(1.0, float(so_ref.abs().max()))
    assert float((y_got - y_ref).abs().max()) <= 2e-5 * max(1.0, float(y_ref.abs().max()))


# ------------------------------------------------------------------------------------------------------------------------
# BASELINE config 5b: FCVSR-S at 540x960 -> 2160x3840
# ------------------------------------------------------------------------------------------------------------------------
def test_fcvsr_s_540x960_parity_against_oracle(dev):
    """The 4K-output case benchmarked in round 1 without a parity test: H = 540 = 27*20 and W = 960 = 32*30 take FFT radix pairs
    no other model-level test reaches.  All three compute modes against the CPU oracle with the tolerances of SURVEY 8(d)."""
    sd = arch.seeded_state_dict("S", 0)
    x = make_clip(77, 1, 540, 960)
    with torch.no_grad():
        ref = O.forward(sd, x)
    tgt = torch.rand(ref.shape, generator=torch.Generator().manual_seed(5))
    m = arch.GShiftNet_S().to(dev).eval()
    m.load_state_dict(sd)
    xd = x.to(dev)
    for mode, tol in (("fp32", 2e-5), ("tf32", 1e-3), ("bf16", 5e-3)):
        m.compute_dtype = mode
        with torch.no_grad():
            y = m(xd).cpu()
        m._engine = None                                   # drop the 2.5 GB workspace before the next mode
        torch.cuda.empty_cache()
        err = float((y - ref).abs().max())
        assert err <= tol, (mode, err)
        if mode == "tf32":
            assert abs(psnr(y, tgt) - psnr(ref, tgt)) <= 0.01
        if mode == "bf16":
            assert psnr(y, ref) >= 60.0


# ------------------------------------------------------------------------------------------------------------------------
# DCN: backward on the tensors saved by the tensor-core forward (the default forward)
# ------------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cfg", [dict(B=1, cin=64, cout=64, H=20, W=24, dg=16), dict(B=2, cin=32, cout=48, H=13, W=17, dg=4)])
def test_dcn_backward_after_tensor_core_forward(dev, cfg):
    """Default configuration of ops.dcn (PRECISION = "tf32"): the forward runs on tcgen05 with TF32 operands, the backward kernels
    (fp32) consume the tensors it saved.  Forward within the TF32 bound, all five gradients within 2e-4 of their scale of the
    oracle's autograd (the backward does not depend on the forward's rounding: it re-samples from the saved fp32 input)."""
    import fcvsr_b200.ops.dcn as dcn_mod
    assert dcn_mod.PRECISION == "tf32"
    g = torch.Generator().manual_seed(cfg["H"] + cfg["dg"])
    B, ci, co, H, W, dg = cfg["B"], cfg["cin"], cfg["cout"], cfg["H"], cfg["W"], cfg["dg"]
    x = torch.randn(B, ci, H, W, generator=g)
    w = torch.randn(co, ci, 3, 3, generator=g) / math.sqrt(ci * 9)
    b = torch.randn(co, generator=g)
    off = 2.0 * torch.randn(B, dg * 18, H, W, generator=g)
    msk = torch.rand(B, dg * 9, H, W, generator=g)
    gy = torch.randn(B, co, H, W, generator=g)
    cpu = [t.clone().requires_grad_(True) for t in (x, off, msk, w, b)]
    ref = O.modulated_deform_conv(cpu[0], cpu[1], cpu[2], cpu[3], cpu[4], 1, 1, 1, 1, dg)
    (ref * gy).sum().backward()
    gpu = [t.to(dev).requires_grad_(True) for t in (x, off, msk, w, b)]
    y = dcn_mod.modulated_deform_conv(gpu[0], gpu[1], gpu[2], gpu[3], gpu[4], 1, 1, 1, 1, dg)
    (y * gy.to(dev)).sum().backward()
    torch.cuda.synchronize()
    assert float((y.detach().cpu() - ref.detach()).abs().max()) <= 2e-3 * max(1.0, float(ref.abs().max()))
    for name, a, r in zip(("input", "offset", "mask", "weight", "bias"), gpu, cpu):
        err = float((a.grad.cpu() - r.grad).abs().max())
        assert err <= 2e-4 * max(1.0, float(r.grad.abs().max())), (name, err)


def test_quantize_u8_matches_numpy_truncation(dev):
    """fcvsr_quantize_u8 = crop + clamp + *255 + astype(uint8) of the evaluation driver (test_LD_freqCVSR.py:85-93)."""
    from fcvsr_b200.sequence import quantize_u8
    g = torch.Generator().manual_seed(2)
    y = torch.rand(3, 1, 40, 52, generator=g) * 1.4 - 0.2
    y[0, 0, 0, :4] = torch.tensor([0.0, 1.0, 254.9999 / 255.0, 0.5])
    want = torch.from_numpy((torch.clamp(y[..., :36, :50], 0, 1).numpy() * 255.0).astype("uint8"))
    got = quantize_u8(y.to(dev), 36, 50).cpu()
    assert torch.equal(got, want)


@pytest.mark.parametrize("B,H,W,crop", [(2, 60, 70, 4), (1, 144, 160, 4), (3, 19, 23, 4), (1, 32, 40, 0)])
def test_psnr_ssim_kernel_matches_oracle(dev, B, H, W, crop):
    """fcvsr_psnr_ssim_u8 (Y-PSNR / SSIM with a cropped border, metric/psnr_ssim.py:278-399) against the numpy restatement that
    tests/test_oracle.py pins to the live reference; ragged 16-pixel tiles, one identical pair (PSNR = inf, SSIM = 1)."""
    import numpy as np
    from fcvsr_b200.metrics import psnr_ssim_u8
    from oracle import metrics_oracle as M
    g = torch.Generator().manual_seed(H + W)
    a = torch.randint(0, 256, (B, 1, H, W), generator=g, dtype=torch.uint8)
    b = (a.int() + torch.randint(-25, 26, a.shape, generator=g)).clamp(0, 255).to(torch.uint8)
    b[0] = a[0] if B > 1 else b[0]
    p, s = psnr_ssim_u8(a.to(dev), b.to(dev), crop)
    assert p.shape == (B, 1) and s.shape == (B, 1)
    for i in range(B):
        af, bf = a[i, 0].numpy().astype(np.float64), b[i, 0].numpy().astype(np.float64)
        pr, sr = M.calculate_psnr(af, bf, crop), M.calculate_ssim(af, bf, crop)
        if pr == float("inf"):
            assert float(p[i, 0]) == float("inf") and abs(float(s[i, 0]) - 1.0) <= 1e-6
        else:
            assert abs(float(p[i, 0]) - pr) <= 1e-4 and abs(float(s[i, 0]) - sr) <= 1e-6


@pytest.mark.parametrize("B,cin,cout,H,W,dg", [(1, 64, 64, 20, 24, 16), (2, 64, 48, 13, 17, 16)])
def test_modulated_dcn_pack_tensor_core_path(dev, B, cin, cout, H, W, dg):
    """ModulatedDeformConvPack inference on the tensor cores (the reference's canonical use of the operator, SURVEY 8 a10:
    MVDeformableAlignment(64, 64, 3, padding=1, deformable_groups=16)): one pixel-major TF32 copy of x, conv_offset_mask on
    tcgen05 with NHWC output, fused gather + GEMM DCN reading offsets / mask from that output (sigmoid on the fly).  Against the
    oracle (F.conv2d + sigmoid + loop-level DCN) within the TF32 contract; the offsets depend on TF32-rounded convolution
    outputs, so the bound is 5e-3 of the output scale."""
    import fcvsr_b200.ops.dcn as dcn_mod
    torch.manual_seed(B + dg)
    m = dcn_mod.ModulatedDeformConvPack(cin, cout, 3, stride=1, padding=1, deformable_groups=dg).to(dev)
    with torch.no_grad():
        m.conv_offset_mask.weight.normal_(0, 0.02)
        m.conv_offset_mask.bias.normal_(0, 0.5)
        m.bias.normal_(0, 0.1)
        x = torch.randn(B, cin, H, W, device=dev)
        calls = []
        orig = dcn_mod.C.try_call
        dcn_mod.C.try_call = lambda name, *a: (calls.append(name), orig(name, *a))[1]
        try:
            y = m(x).cpu()
        finally:
            dcn_mod.C.try_call = orig
        assert calls == ["fcvsr_conv2d_tc", "fcvsr_modulated_deform_conv_forward_tc"], calls
        o = F.conv2d(x.cpu(), m.conv_offset_mask.weight.cpu(), m.conv_offset_mask.bias.cpu(), padding=1)
        o1, o2, mk = torch.chunk(o, 3, dim=1)
        ref = O.modulated_deform_conv(x.cpu(), torch.cat((o1, o2), 1), torch.sigmoid(mk), m.weight.cpu(), m.bias.cpu(), 1, 1, 1, 1, dg)
    err = float((y - ref).abs().max())
    assert err <= 5e-3 * max(1.0, float(ref.abs().max())), err


def test_deform_conv_cuda_module_mirrors_reference_entry_points(dev):
    """fcvsr_b200.ops.deform_conv_cuda: the five pybind entry names of the reference extension with its argument order
    (deform_conv_cuda.cpp:681-695), driven exactly as deform_conv.py:52-57,:76-92,:144-166 drives them, against the oracle's
    autograd; plus fp16 tensors through the public op (computed in fp32, returned as fp16)."""
    import fcvsr_b200.ops.dcn as dcn_mod
    from fcvsr_b200.ops import deform_conv_cuda as ext
    assert all(hasattr(ext, n) for n in ("deform_conv_forward_cuda", "deform_conv_backward_input_cuda", "deform_conv_backward_parameters_cuda",
                                          "modulated_deform_conv_cuda_forward", "modulated_deform_conv_cuda_backward"))
    g = torch.Generator().manual_seed(12)
    B, ci, co, H, W, dg = 2, 8, 6, 9, 11, 4
    x = torch.randn(B, ci, H, W, generator=g)
    w = torch.randn(co, ci, 3, 3, generator=g) / 6
    b = torch.randn(co, generator=g)
    off = 2.0 * torch.randn(B, dg * 18, H, W, generator=g)
    msk = torch.rand(B, dg * 9, H, W, generator=g)
    gy = torch.randn(B, co, H, W, generator=g)
    old = dcn_mod.PRECISION
    dcn_mod.PRECISION = "fp32"
    try:
        # --- modulated (v2), as ModulatedDeformConvFunction.forward / backward call it ---
        cpu = [t.clone().requires_grad_(True) for t in (x, off, msk, w, b)]
        ref = O.modulated_deform_conv(cpu[0], cpu[1], cpu[2], cpu[3], cpu[4], 1, 1, 1, 1, dg)
        (ref * gy).sum().backward()
        xd, od, md, wd, bd, gyd = (t.to(dev) for t in (x, off, msk, w, b, gy))
        out = xd.new_empty(B, co, H, W)
        bufs = [xd.new_empty(0), xd.new_empty(0)]
        ext.modulated_deform_conv_cuda_forward(xd, wd, bd, bufs[0], od, md, out, bufs[1], 3, 3, 1, 1, 1, 1, 1, 1, 1, dg, True)
        assert float((out.cpu() - ref.detach()).abs().max()) <= 1e-4
        gi, go, gm, gw, gb = (torch.zeros_like(t) for t in (xd, od, md, wd, bd))
        ext.modulated_deform_conv_cuda_backward(xd, wd, bd, bufs[0], od, md, bufs[1], gi, gw, gb, go, gm, gyd, 3, 3, 1, 1, 1, 1, 1, 1, 1,
                                                dg, True)
        for name, a, r in zip(("input", "offset", "mask", "weight", "bias"), (gi, go, gm, gw, gb), cpu):
            assert float((a.cpu() - r.grad).abs().max()) <= 2e-4 * max(1.0, float(r.grad.abs().max())), name
        # --- v1, as DeformConvFunction calls it (W / H argument order) ---
        cpu = [t.clone().requires_grad_(True) for t in (x, off, w)]
        ref = O.modulated_deform_conv(cpu[0], cpu[1], None, cpu[2], None, 1, 1, 1, 1, dg)
        (ref * gy).sum().backward()
        out = xd.new_empty(B, co, H, W)
        assert ext.deform_conv_forward_cuda(xd, wd, od, out, bufs[0], bufs[1], 3, 3, 1, 1, 1, 1, 1, 1, 1, dg, B) == 1
        assert float((out.cpu() - ref.detach()).abs().max()) <= 1e-4
        gi, go, gw = torch.zeros_like(xd), torch.zeros_like(od), torch.zeros_like(wd)
        ext.deform_conv_backward_input_cuda(xd, od, gyd, gi, go, wd, bufs[0], 3, 3, 1, 1, 1, 1, 1, 1, 1, dg, B)
        ext.deform_conv_backward_parameters_cuda(xd, od, gyd, gw, bufs[0], bufs[1], 3, 3, 1, 1, 1, 1, 1, 1, 1, dg, 1, B)
        for name, a, r in zip(("input", "offset", "weight"), (gi, go, gw), cpu):
            assert float((a.cpu() - r.grad).abs().max()) <= 2e-4 * max(1.0, float(r.grad.abs().max())), name
        # --- fp16 tensors through the public operator ---
        yh = dcn_mod.modulated_deform_conv(xd.half(), od.half(), md.half(), wd.half(), bd.half(), 1, 1, 1, 1, dg)
        assert yh.dtype == torch.float16
        ref16 = O.modulated_deform_conv(x.half().float(), off.half().float(), msk.half().float(), w.half().float(), b.half().float(),
                                        1, 1, 1, 1, dg)
        assert float((yh.float().cpu() - ref16).abs().max()) <= 2.0 ** -9 * max(1.0, float(ref16.abs().max()))
    finally:
        dcn_mod.PRECISION = old


# ------------------------------------------------------------------------------------------------------------------------
# fcvsr_conv2d_tc_multi_w: the 1x1 `down` and `up` convolutions of a BlockRCB (CVSR_freq.py:753-763) as one launch
# ------------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("op16", [0, 1])
def test_conv_multi_stacked_filters(dev, op16):
    """Four problems (two pyramid levels x two stacked 64 -> 64 1x1 filters with bias) in one launch against F.conv2d on the
    operand-rounded inputs / weights; ragged tiles (sizes that are not multiples of the 8 x 16 tile)."""
    import ctypes
    g = torch.Generator().manual_seed(3 + op16)
    B = 2
    dims = [(18, 20), (9, 10), (18, 20), (9, 10)]
    wrow = [0, 0, 64, 64]
    w = torch.randn(128, 64, generator=g) / 8
    b = torch.randn(128, generator=g)
    xs = [torch.randn(B, 64, h, ww, generator=g) for h, ww in dims]
    if op16:
        rnd = lambda t: t.to(torch.bfloat16).float()  # noqa: E731
    else:
        rnd = lambda t: ((t.contiguous().view(torch.int32) + 0x1000) & -8192).view(torch.float32)  # noqa: E731
    wr = rnd(w)
    dt = torch.bfloat16 if op16 else torch.float32
    xd = [nhwc(rnd(x)).to(dev).to(dt).contiguous() for x in xs]
    wd = wr.to(dev).to(dt).contiguous()
    bd = b.to(dev)
    yd = [torch.zeros(B, h, ww, 64, device=dev, dtype=dt) for h, ww in dims]
    n = 4
    vp = lambda ts: (ctypes.c_void_p * n)(*[t.data_ptr() for t in ts])  # noqa: E731
    ia = lambda v: (ctypes.c_int * n)(*v)  # noqa: E731
    C.call("fcvsr_conv2d_tc_multi_w", n, vp(xd), 64, wd.data_ptr(), bd.data_ptr(), ia(wrow), 128, vp(yd), 64, ia([d[0] for d in dims]),
           ia([d[1] for d in dims]), B, 64, 64, 1, C.ACT_NONE, 0.0, op16, op16, _st())
    torch.cuda.synchronize()
    for i in range(n):
        ref = F.conv2d(rnd(xs[i]), wr[wrow[i]:wrow[i] + 64].view(64, 64, 1, 1), b[wrow[i]:wrow[i] + 64])
        got = nchw(yd[i].float().cpu())
        tol = (2 ** -8 if op16 else 2e-5) * max(1.0, float(ref.abs().max()))
        assert float((got - ref).abs().max()) <= tol, (i, float((got - ref).abs().max()))
