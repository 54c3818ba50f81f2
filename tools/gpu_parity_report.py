"""Print max-abs error / PSNR of the three compute modes against the committed reference goldens."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fcvsr_b200 import arch  # noqa: E402
from tests.util import load_golden, make_clip  # noqa: E402

dev = torch.device("cuda:0")
for name in ("fcvsr_s_64", "fcvsr_s_36x40", "fcvsr_full_64"):
    g = load_golden(name)
    c = g["case"]
    sd = arch.seeded_state_dict(c["variant"], c["seed"])
    x = make_clip(c["clip_seed"], c["b"], c["h"], c["w"]).to(dev)
    ref = g["out"]
    for mode in ("fp32", "tf32", "bf16"):
        m = (arch.GShiftNet_S if c["variant"] == "S" else arch.GShiftNet)().to(dev).eval()
        m.load_state_dict(sd)
        m.compute_dtype = mode
        with torch.no_grad():
            y = m(x).cpu()
        err = float((y - ref).abs().max())
        mse = float(((y - ref) ** 2).mean())
        print(f"{name:14s} {mode}: max-abs {err:.3e}  PSNR(ours, reference) {10 * torch.log10(torch.tensor(1.0 / max(mse, 1e-20))).item():.1f} dB"
              f"  (|ref| max {float(ref.abs().max()):.2f})")
