"""Bring-up helper: MFFR internals (bands, final) vs the oracle for several shapes / batch sizes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fcvsr_b200 import arch  # noqa: E402
from fcvsr_b200.engine import Engine  # noqa: E402
from oracle import fcvsr_oracle as O  # noqa: E402

dev = torch.device("cuda:0")
for (seed, B, H, W) in ((0, 1, 64, 64), (0, 2, 64, 64), (3, 1, 36, 40), (3, 2, 36, 40), (0, 1, 36, 40)):
    sd = arch.seeded_state_dict("S", seed)
    m = arch.GShiftNet_S().to(dev).eval()
    m.load_state_dict(sd)
    eng = Engine(m, use_tc=False)
    eng._ensure_packs(dev)
    ws = eng._workspace(B, H, W, dev)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, 64, H, W, generator=g)
    ws["m2"].copy_(x.permute(0, 2, 3, 1).reshape(B, H * W, 64))
    eng.st = torch.cuda.current_stream().cuda_stream
    p = {k: v.data_ptr() for k, v in ws.items()}
    eng._mffr(ws, p, B, H, W)
    torch.cuda.synchronize()
    Q = m.Freq_Inv
    bands_ref = O.split_freq(x, Q)
    got = ws["bands"].view(Q, B, H, W, 64).permute(0, 1, 4, 2, 3).cpu()
    be = [float((got[q] - bands_ref[q]).abs().max()) for q in range(Q)]
    ref = O.mffr(sd, x, Q)
    out = ws["xs0"].view(B, H, W, 64).permute(0, 3, 1, 2).cpu()
    per_b = [float((out[b] - ref[b]).abs().max()) for b in range(B)]
    print(f"seed {seed} B{B} {H}x{W}: band errs {['%.1e' % e for e in be]}  out err per batch {['%.1e' % e for e in per_b]}")
