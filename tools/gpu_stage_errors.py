"""Bring-up helper (GPU box): per-kernel and per-stage errors of the CUDA path against the CPU oracle.

    python tools/gpu_stage_errors.py [--no-tc] [--variant S|full] [--hw 64 64]
"""
import argparse
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fcvsr_b200 import _capi as C, arch, bands  # noqa: E402
from fcvsr_b200.engine import Engine, _ConvPack  # noqa: E402
from oracle import fcvsr_oracle as O  # noqa: E402
from oracle.make_golden import make_clip  # noqa: E402

dev = torch.device("cuda:0")


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def nchw(t):
    return t.permute(0, 3, 1, 2).contiguous()


def st():
    return torch.cuda.current_stream().cuda_stream


def err(a, b):
    return float((a - b).abs().max()), float(b.abs().max())


def check_fft():
    for (H, W, Cc) in ((64, 64, 64), (36, 40, 24), (180, 320, 64), (272, 480, 12)):
        x = torch.randn(2, Cc, H, W)
        xd = nhwc(x).to(dev)
        Wf = W // 2 + 1
        spec = torch.empty(2, H, Wf, Cc, 2, device=dev)
        tw_w, tw_h = bands.twiddles(W, dev), bands.twiddles(H, dev)
        C.call("fcvsr_fft_r2c_w", xd.data_ptr(), Cc, spec.data_ptr(), tw_w.data_ptr(), 2, H, W, Cc, st())
        C.call("fcvsr_fft_c2c_h", spec.data_ptr(), spec.data_ptr(), tw_h.data_ptr(), 0, 2, H, Wf, Cc, 0, 1.0, 0, 1, 0, st())
        ref = torch.fft.rfft2(x)
        got = torch.view_as_complex(spec.cpu()).permute(0, 3, 1, 2)
        e1 = float((got - ref).abs().max()) / float(ref.abs().max())
        # inverse with torch c2r semantics on a NON-hermitian spectrum
        z = torch.randn(2, Cc, H, Wf, dtype=torch.complex64)
        zd = torch.view_as_real(z.permute(0, 2, 3, 1).contiguous()).contiguous().to(dev)
        y = torch.empty(2, H, W, Cc, device=dev)
        C.call("fcvsr_fft_c2c_h", zd.data_ptr(), zd.data_ptr(), tw_h.data_ptr(), 0, 2, H, Wf, Cc, 1, 1.0, 0, 1, 0, st())
        C.call("fcvsr_fft_c2r_w", zd.data_ptr(), y.data_ptr(), Cc, tw_w.data_ptr(), 2, H, W, Cc, 1.0 / (H * W), st())
        ref2 = torch.fft.irfft2(z, s=(H, W))
        e2 = float((nchw(y.cpu()) - ref2).abs().max()) / float(ref2.abs().max())
        print(f"fft {H}x{W} C={Cc}: rfft2 rel err {e1:.2e}  irfft2 rel err {e2:.2e}")


def run_conv(x, w, b, tc, act=0, slope=0.0, res=None, ps=False, stride=1):
    B, Cin, H, W = x.shape
    pk = _ConvPack(w.to(dev), b.to(dev) if b is not None else None, stride=stride, ps=ps)
    xd = nhwc(x).to(dev)
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    cout = w.shape[0]
    if ps:
        y = torch.empty(B, 2 * Ho, 2 * Wo, cout // 4, device=dev)
    else:
        y = torch.empty(B, Ho, Wo, cout, device=dev)
    rd = nhwc(res).to(dev) if res is not None else None
    if tc:
        C.call("fcvsr_conv2d_tc", xd.data_ptr(), Cin, pk.w_tc.data_ptr(), pk.bias.data_ptr() if b is not None else 0,
               rd.data_ptr() if rd is not None else 0, cout, 0, 0, y.data_ptr(), y.shape[-1], B, H, W, Cin, cout,
               w.shape[-1], act, slope, 0, int(ps), 0, 0, 0, 0, 0, st())
    else:
        C.call("fcvsr_conv2d_direct", xd.data_ptr(), Cin, 0, pk.w_direct.data_ptr(),
               pk.bias.data_ptr() if b is not None else 0, rd.data_ptr() if rd is not None else 0, cout, 0, 0,
               y.data_ptr(), y.shape[-1], B, H, W, Cin, cout, w.shape[-1], stride, act, slope, 0, int(ps), 0, 0, 0, 0, 0, st())
    torch.cuda.synchronize()
    return nchw(y.cpu())


def check_conv(tc):
    g = torch.Generator().manual_seed(0)
    cases = [(1, 64, 64, 16, 16, 3), (2, 64, 128, 20, 36, 3), (1, 128, 64, 45, 80, 3), (1, 64, 256, 12, 20, 3),
             (1, 96, 64, 9, 17, 3), (2, 256, 128, 10, 33, 1), (1, 64, 4, 11, 19, 1), (1, 64, 1, 24, 40, 3),
             (1, 64, 576, 8, 16, 1)]
    if not tc:
        cases += [(1, 7, 448, 12, 12, 3), (1, 209, 64, 6, 9, 1)]
    for (B, ci, co, H, W, k) in cases:
        x = torch.randn(B, ci, H, W, generator=g)
        w = torch.randn(co, ci, k, k, generator=g) / (ci * k * k) ** 0.5
        b = torch.randn(co, generator=g)
        res = torch.randn(B, co, H, W, generator=g)
        y = run_conv(x, w, b, tc, act=2, slope=0.1, res=res)
        ref = F.leaky_relu(F.conv2d(x, w, b, padding=k // 2), 0.1) + res
        e = err(y, ref)
        print(f"conv{'_tc' if tc else '_direct'} B{B} {ci}->{co} {H}x{W} k{k}: max err {e[0]:.2e} (ref max {e[1]:.2f})")
        if co % 64 == 0 and co >= 64:
            y = run_conv(x, w, b, tc, ps=True)
            ref = F.pixel_shuffle(F.conv2d(x, w, b, padding=k // 2), 2)
            e = err(y, ref)
            print(f"     + pixel_shuffle: max err {e[0]:.2e}")
    if not tc:
        x = torch.randn(1, 64, 16, 20, generator=g)
        w = torch.randn(64, 64, 3, 3, generator=g) / 24
        y = run_conv(x, w, None, False, stride=2)
        print("conv_direct stride 2: max err %.2e" % err(y, F.conv2d(x, w, None, stride=2, padding=1))[0])


def check_model(variant, H, W, use_tc, B=1, mode=None):
    sd = arch.seeded_state_dict(variant, 0)
    x = make_clip(1234, B, H, W)
    t0 = time.time()
    with torch.no_grad():
        ref, taps = O.forward(sd, x, return_taps=True)
    t_cpu = time.time() - t0
    model = (arch.GShiftNet_S if variant == "S" else arch.GShiftNet)().to(dev).eval()
    model.load_state_dict(sd)
    model._engine = Engine(model, use_tc=use_tc, mode=mode)
    model.compute_dtype = model._engine.mode
    with torch.no_grad():
        y = model(x.to(dev))
        torch.cuda.synchronize()
        t0 = time.time()
        y = model(x.to(dev))
        torch.cuda.synchronize()
    t_gpu = time.time() - t0
    eng = model._engine
    ws = eng._workspace(B, H, W, dev)
    P = H * W

    def tap(t, c0, c1, h, w):
        return nchw(t.view(B, h, w, -1)[..., c0:c1].cpu())

    stages = [("mgaa1", tap(ws["feat"], 128, 192, H, W)), ("mgaa2", tap(ws["m2"], 0, 64, H, W)),
              ("mffr", tap(ws["xs0"], 0, 64, H, W)), ("sc_l1", tap(ws["fuse"], 0, 64, H, W)),
              ("sc_l3", tap(ws["o3"], 0, 64, H // 4, W // 4)), ("fuse", tap(ws["f2"], 0, 64, H, W))]
    print(f"model {variant} {H}x{W} mode={model._engine.mode}: cpu oracle {t_cpu:.2f}s, gpu eager {t_gpu * 1e3:.1f} ms, "
          f"launches {eng.launches} (tc {eng.tc_launches})")
    for name, got in stages:
        e = err(got, taps[name])
        print(f"   {name:6s} max err {e[0]:.3e} (ref max {e[1]:.2f})")
    e = err(y.cpu(), ref)
    print(f"   out    max err {e[0]:.3e} (ref max {e[1]:.2f})")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--no-tc", action="store_true")
    ap.add_argument("--variant", default="S")
    ap.add_argument("--hw", type=int, nargs=2, default=[64, 64])
    ap.add_argument("--skip-kernels", action="store_true")
    a = ap.parse_args()
    print(C.version(), torch.cuda.get_device_name(0))
    if not a.skip_kernels:
        check_fft()
        check_conv(False)
        if not a.no_tc:
            check_conv(True)
    check_model(a.variant, a.hw[0], a.hw[1], False)
    if not a.no_tc:
        check_model(a.variant, a.hw[0], a.hw[1], True)
        check_model(a.variant, a.hw[0], a.hw[1], True, mode="bf16")
