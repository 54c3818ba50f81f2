"""Bring-up helper: stage errors for a golden case, per batch element."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fcvsr_b200 import arch  # noqa: E402
from fcvsr_b200.engine import Engine  # noqa: E402
from oracle import fcvsr_oracle as O  # noqa: E402
from oracle.make_golden import make_clip  # noqa: E402

dev = torch.device("cuda:0")
for (seed, cseed, B, H, W) in ((3, 77, 2, 36, 40), (3, 77, 1, 36, 40), (0, 77, 2, 36, 40), (3, 77, 2, 64, 64)):
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

    def tap(t, c0, c1, h, w):
        return t.view(B, h, w, -1)[..., c0:c1].permute(0, 3, 1, 2).cpu()

    st = {"mgaa1": tap(ws["feat"], 128, 192, H, W), "mgaa2": tap(ws["m2"], 0, 64, H, W), "mffr": tap(ws["xs0"], 0, 64, H, W)}
    print(f"seed {seed} B{B} {H}x{W}")
    for k, v in st.items():
        print("   ", k, ["%.2e" % float((v[b] - taps[k][b]).abs().max()) for b in range(B)])
    print("    out", ["%.2e" % float((y[b] - ref[b]).abs().max()) for b in range(B)])
