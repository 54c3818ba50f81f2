"""Generate tests/golden/*.pt by running the UNMODIFIED reference (build container only).

    python oracle/make_golden.py

For each case the seeded weights (fcvsr_b200.arch.seeded_state_dict) are loaded *strictly* into
the reference ``GShiftNet`` / ``GShiftNet_S`` (CVSR_train/arch/CVSR_freq.py:2653,2577) and the
reference forward is run on a seeded 8-bit-quantised clip.  Stored: the output, and
strided samples of intermediate tensors captured with forward hooks (MGAA calls, MFFRblock,
recorb1 levels), so that a parity failure can be localised.  Also stores the reference's
state-dict key/shape list and the DCN known-answer vector of ops/dcn/simple_check.py:8-22.
"""
from __future__ import annotations

import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from fcvsr_b200.arch import seeded_state_dict  # noqa: E402
from oracle import ref_loader  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def make_clip(seed: int, b: int, h: int, w: int) -> torch.Tensor:
    """8-bit decoded frames / 255 (test_LD_freqCVSR_S_FPS.py:28), smooth + shifted per frame so the
    7 frames look like a video (low-pass 3x3 box, <=2 px shifts)."""
    g = torch.Generator().manual_seed(seed)
    base = torch.rand(b, 1, h + 8, w + 8, generator=g)
    base = torch.nn.functional.avg_pool2d(base, 3, 1, 1)
    base = (base - base.min()) / (base.max() - base.min())
    frames = []
    for t in range(7):
        dy, dx = (t * 2) % 5, (t * 3) % 5
        frames.append(base[:, :, dy:dy + h, dx:dx + w])
    x = torch.stack(frames, 1)
    x = x + 0.05 * torch.rand(x.shape, generator=g)
    return torch.round(255.0 * x.clamp(0, 1)) / 255.0


def make_clip_rgb(seed: int, b: int, h: int, w: int) -> torch.Tensor:
    """[B,7,3,H,W]: three independently seeded planes of make_clip (the mmedit variants take RGB frames)."""
    return torch.cat([make_clip(seed + 1000 * k, b, h, w) for k in range(3)], 2)


def build_reference(ref, variant: str):
    """The reference module for a variant.  "rgb" / "rgb_S" are the mmedit backbones FCVSRNet / FCVSR_SNet
    (mmedit_train/mmedit/models/backbones/sr_backbones/fcvsr.py:38-142, fcvsr_s.py:40-): that file needs mmcv (absent here), and it
    is GShiftNet's code with feat_extract 21 -> 448 and conv_last0 64 -> 3, so the UNMODIFIED GShiftNet.forward
    (CVSR_freq.py:2688-2756, channel-count agnostic) is run with those two layers re-shaped."""
    if variant == "S":
        return ref.GShiftNet_S()
    if variant == "full":
        return ref.GShiftNet()
    m = ref.GShiftNet() if variant == "rgb" else ref.GShiftNet(ACNum=3, Freq_Inv=4, SCGroupN=4)
    m.feat_extract = torch.nn.Sequential(torch.nn.Conv2d(21, 7 * 64, 3, 1, 1))
    m.conv_last0 = torch.nn.Conv2d(64, 3, 3, 1, 1)
    return m


CASES = [
    dict(name="fcvsr_s_64", variant="S", seed=0, clip_seed=1234, b=1, h=64, w=64),
    dict(name="fcvsr_full_64", variant="full", seed=0, clip_seed=1234, b=1, h=64, w=64),
    dict(name="fcvsr_s_36x40", variant="S", seed=3, clip_seed=77, b=2, h=36, w=40),
    dict(name="fcvsrnet_s_32x40", variant="rgb_S", seed=5, clip_seed=21, b=2, h=32, w=40),
    dict(name="fcvsrnet_32", variant="rgb", seed=6, clip_seed=22, b=1, h=32, w=32),
]


def sample(t: torch.Tensor) -> torch.Tensor:
    return t[..., ::4, ::4].contiguous().clone()


def main() -> None:
    os.makedirs(GOLD, exist_ok=True)
    ref = ref_loader.load()
    shapes = {}
    for case in CASES:
        sd = seeded_state_dict(case["variant"], case["seed"])
        model = build_reference(ref, case["variant"]).eval()
        missing = model.load_state_dict(sd, strict=True)
        assert not missing.missing_keys and not missing.unexpected_keys
        shapes[case["variant"]] = [[k, list(v.shape)] for k, v in model.state_dict().items()]
        taps = {}
        calls = {"n": 0}

        def hook_mgaa(_m, _i, o):
            calls["n"] += 1
            taps[f"mgaa{calls['n']}"] = sample(o[0])

        hs = [model.MGAA.register_forward_hook(hook_mgaa),
              model.MFFRblock.register_forward_hook(lambda _m, _i, o: taps.__setitem__("mffr", sample(o))),
              model.recorb1.register_forward_hook(
                  lambda _m, _i, o: taps.update(sc_l1=sample(o[0]), sc_l3=o[2].clone())),
              model.recorb0.register_forward_hook(lambda _m, _i, o: taps.__setitem__("fuse", sample(o)))]
        mk = make_clip_rgb if case["variant"].startswith("rgb") else make_clip
        x = mk(case["clip_seed"], case["b"], case["h"], case["w"])
        with torch.no_grad():
            y = model(x)
        for hnd in hs:
            hnd.remove()
        # reference call order is MGAA(f1), MGAA(f3), MGAA(cat): keep 1st and 3rd under oracle names
        out = {"case": case, "out": y.clone(), "mgaa1": taps["mgaa1"], "mgaa2": taps["mgaa3"],
               "mffr": taps["mffr"], "sc_l1": taps["sc_l1"], "sc_l3": taps["sc_l3"], "fuse": taps["fuse"]}
        torch.save(out, os.path.join(GOLD, case["name"] + ".pt"))
        print(case["name"], tuple(y.shape), float(y.abs().max()), float(y.mean()))
    with open(os.path.join(GOLD, "state_dict_shapes.json"), "w") as f:
        json.dump(shapes, f)
    # DCN known-answer test of the reference (ops/dcn/simple_check.py:8-22)
    kat = {"input": torch.arange(18, dtype=torch.float32).view(1, 2, 3, 3).tolist(),
           "offset18": [1, 1, 1, 0, 1, -1, 0, 1, 0, 0, 0, -1, -1, 1, -1, 0, -1, -1],
           "expected": [81, 99, 117, 135, 153, 171, 189, 207, 225]}
    with open(os.path.join(GOLD, "dcn_simple_check.json"), "w") as f:
        json.dump(kat, f)


if __name__ == "__main__":
    main()
