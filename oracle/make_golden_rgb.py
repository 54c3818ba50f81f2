"""Generate tests/golden/fcvsr_rgb_*.pt by running the UNMODIFIED reference RGB models (build container only).

    python oracle/make_golden_rgb.py

`FCVSR` / `FCVSR_S` of CVSR_train/arch/CVSR_freq_RGB.py:2135-2202 / :2059-2128 are imported under the matplotlib stub of
oracle/ref_loader.py, the seeded weights (fcvsr_b200.arch_rgb.seeded_state_dict_rgb) are loaded strictly, and the forward runs
on a seeded RGB clip.  Stored: output, strided stage taps (forward hooks), the reference's state-dict key / shape list.
"""
from __future__ import annotations

import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from fcvsr_b200.arch_rgb import seeded_state_dict_rgb  # noqa: E402
from oracle import ref_loader  # noqa: E402
from oracle.make_golden import GOLD, make_clip_rgb, sample  # noqa: E402

CASES = [
    dict(name="fcvsr_rgb_s_32x40", variant="S", seed=2, clip_seed=31, b=2, h=32, w=40),
    dict(name="fcvsr_rgb_full_32", variant="full", seed=4, clip_seed=32, b=1, h=32, w=32),
]


def main() -> None:
    ref_loader.load()
    rgb = importlib.import_module("arch.CVSR_freq_RGB")
    shapes = {}
    for case in CASES:
        sd = seeded_state_dict_rgb(case["variant"], case["seed"])
        model = (rgb.FCVSR_S if case["variant"] == "S" else rgb.FCVSR)().eval()
        res = model.load_state_dict(sd, strict=True)
        assert not res.missing_keys and not res.unexpected_keys
        shapes[case["variant"]] = [[k, list(v.shape)] for k, v in model.state_dict().items()]
        taps, calls = {}, {"n": 0}

        def hook_mgaa(_m, _i, o):
            calls["n"] += 1
            taps[f"mgaa{calls['n']}"] = sample(o)

        hs = [model.MGAA.register_forward_hook(hook_mgaa),
              model.MFFRblock.register_forward_hook(lambda _m, _i, o: taps.__setitem__("mffr", sample(o))),
              model.recorb1.register_forward_hook(lambda _m, _i, o: taps.update(sc_l1=sample(o[0]), sc_l3=o[2].clone())),
              model.recorb0.register_forward_hook(lambda _m, _i, o: taps.__setitem__("fuse", sample(o)))]
        x = make_clip_rgb(case["clip_seed"], case["b"], case["h"], case["w"])
        with torch.no_grad():
            y = model(x)
        for hnd in hs:
            hnd.remove()
        out = {"case": case, "out": y.clone(), "mgaa1": taps["mgaa1"], "mgaa2": taps["mgaa3"], "mffr": taps["mffr"],
               "sc_l1": taps["sc_l1"], "sc_l3": taps["sc_l3"], "fuse": taps["fuse"]}
        torch.save(out, os.path.join(GOLD, case["name"] + ".pt"))
        print(case["name"], tuple(y.shape), float(y.abs().max()), float(y.mean()))
    with open(os.path.join(GOLD, "state_dict_shapes_rgb.json"), "w") as f:
        json.dump(shapes, f)


if __name__ == "__main__":
    main()
