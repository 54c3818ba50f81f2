"""Whole-forward parity against the CPU oracle at sizes the committed goldens do not cover (exercises the two-phase FFT
radix pairs, ragged conv tiles and the pyramid at realistic shapes).  Prints max-abs error per (variant, shape, mode)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fcvsr_b200 import arch  # noqa: E402
from oracle import fcvsr_oracle as O  # noqa: E402
from oracle.make_golden import make_clip  # noqa: E402

dev = torch.device("cuda:0")
torch.set_num_threads(os.cpu_count() or 1)
cases = [("S", 1, 64, 96), ("S", 2, 96, 128), ("S", 1, 120, 160), ("S", 1, 180, 320), ("full", 1, 180, 320), ("S", 1, 272, 480),
         ("full", 2, 68, 100), ("S", 3, 52, 76), ("full", 1, 100, 132)]      # ragged conv / IAC / ConvBlk tiles, odd FFT radix pairs
if len(sys.argv) > 1:
    cases = cases[: int(sys.argv[1])]
worst = {"fp32": 0.0, "tf32": 0.0, "bf16": 0.0}
for variant, b, h, w in cases:
    sd = arch.seeded_state_dict(variant, 0)
    x = make_clip(4321 + h, b, h, w)
    t0 = time.time()
    with torch.no_grad():
        ref = O.forward(sd, x)
    line = f"{variant:4s} B{b} {h}x{w} (oracle {time.time() - t0:.1f} s):"
    for mode in ("fp32", "tf32", "bf16"):
        m = (arch.GShiftNet_S if variant == "S" else arch.GShiftNet)().to(dev).eval()
        m.load_state_dict(sd)
        m.compute_dtype = mode
        with torch.no_grad():
            y = m(x.to(dev)).cpu()
        err = float((y - ref).abs().max())
        worst[mode] = max(worst[mode], err)
        line += f"  {mode} {err:.2e}"
    print(line, flush=True)
print("worst:", worst)
assert worst["fp32"] <= 2e-5 and worst["tf32"] <= 1e-3 and worst["bf16"] <= 5e-3
