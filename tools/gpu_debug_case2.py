"""Bring-up helper: order-dependence hunt (run 64x64 first, then 36x40 B=2) with MFFR internals."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fcvsr_b200 import arch  # noqa: E402
from fcvsr_b200.engine import Engine  # noqa: E402
from oracle import fcvsr_oracle as O  # noqa: E402
from oracle.make_golden import make_clip  # noqa: E402

dev = torch.device("cuda:0")
junk = [torch.full((1 << 24,), float("nan"), device=dev) for _ in range(16)]   # poison 1 GiB of cached blocks
del junk


def run(seed, cseed, B, H, W):
    sd = arch.seeded_state_dict("S", seed)
    x = make_clip(cseed, B, H, W)
    with torch.no_grad():
        ref, taps = O.forward(sd, x, return_taps=True)
    m = arch.GShiftNet_S().to(dev).eval()
    m.load_state_dict(sd)
    m._engine = Engine(m, use_tc=False)
    m.compute_dtype = 'fp32'
    with torch.no_grad():
        y = m(x.to(dev)).cpu()
    ws = m._engine._workspace(B, H, W, dev)
    Q = 4

    def tap(t, c0, c1, h, w):
        return t.view(B, h, w, -1)[..., c0:c1].permute(0, 3, 1, 2).cpu()

    print(f"seed {seed} B{B} {H}x{W}")
    for k, v in {"mgaa1": tap(ws["feat"], 128, 192, H, W), "mgaa2": tap(ws["m2"], 0, 64, H, W),
                 "mffr": tap(ws["xs0"], 0, 64, H, W)}.items():
        print("   ", k, ["%.2e" % float((v[b] - taps[k][b]).abs().max()) for b in range(B)])
    bands_ref = O.split_freq(taps["mgaa2"], Q)
    got = ws["bands"].view(Q, B, H, W, 64).permute(0, 1, 4, 2, 3).cpu()
    print("    bands", ["%.1e" % float((got[q] - bands_ref[q]).abs().max()) for q in range(Q)])
    print("    gates finite:", bool(torch.isfinite(ws["gates"][:, :, :64]).all()), "mean0 finite:",
          bool(torch.isfinite(ws["mean0"][:, :64]).all()))
    print("    out", ["%.2e" % float((y[b] - ref[b]).abs().max()) for b in range(B)])


run(0, 1234, 1, 64, 64)
run(3, 77, 2, 36, 40)

# ---- which stage corrupts the band masks? -------------------------------------------------------
from fcvsr_b200 import bands as _b  # noqa: E402
sd = arch.seeded_state_dict("S", 3)
B, H, W = 2, 36, 40
x = make_clip(77, B, H, W).to(dev)
m = arch.GShiftNet_S().to(dev).eval()
m.load_state_dict(sd)
eng = Engine(m, use_tc=False)
eng._ensure_packs(dev)
ws = eng._workspace(B, H, W, dev)
ref_masks = ws["masks"].clone()
ref_tw = (ws["tw_w"].clone(), ws["tw_h"].clone())
eng.st = torch.cuda.current_stream().cuda_stream
p = {k: v.data_ptr() for k, v in ws.items()}
f = p["feat"]
P = eng.packs


def chk(tag):
    torch.cuda.synchronize()
    bad = [k for k in ("masks", "tw_w", "tw_h") if not torch.equal(ws[k], {"masks": ref_masks, "tw_w": ref_tw[0], "tw_h": ref_tw[1]}[k])]
    print(f"after {tag}: corrupted {bad}", flush=True)


eng._conv(P["feat"], x.data_ptr(), 0, f, 448, B, H, W, nchw=True); chk("feat")
eng._mgaa(ws, p, f, 448, f + 128 * 4, 448, B, H, W); chk("mgaa1")
eng._mgaa(ws, p, f + 256 * 4, 448, f + 256 * 4, 448, B, H, W); chk("mgaa3")
eng._mgaa(ws, p, f + 128 * 4, 448, p["m2"], 64, B, H, W); chk("mgaa2")
eng._mffr(ws, p, B, H, W); chk("mffr")
# address map of the small buffers
for k, v in sorted(ws.items(), key=lambda kv: kv[1].data_ptr()):
    print(f"{v.data_ptr():#x} {v.numel() * 4:>10d} {k}")
