"""Drop-in classes for the RGB FCVSR family of the reference: `FCVSR` / `FCVSR_S` of
CVSR_train/arch/CVSR_freq_RGB.py:2135-2202 / :2059-2128.

Same constructor keywords, parameter names and shapes as the reference modules (the state dict loads strictly both ways),
same `forward(x[B,7,3,H,W]) -> [B,3,4H,4W]`.  This family differs from GShiftNet (CVSR_freq.py) in every stage: `MGAA` has
no CorrBlock / convcorr, its offset blocks are dense 128 -> 64 -> 4 convolutions with kernel size 2i+1 (:298-310), IAC adds a
predicted per-pixel bias instead of the input (:1009-1023), the DivEnh blocks use their 3x3 `Conv` with a sigmoid gate
(:1575-1612), the band masks are "ideal" discs (:1493-1506), and SCNet's blocks are plain 64 -> 256 -> 64 convolution pairs
(:609-657).  The modules below only own parameters; the forward is `fcvsr_b200.rgb_forward` (kernel-library convolutions /
FFTs / flow_warp / SAC behind autograd Functions, PyTorch elementwise glue), used for inference and training alike.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .arch import _ChannelAttn, _conv, _Holder, _scaled_kaiming


class _DenseOffsetBlk(_Holder):       # ConvBlk, CVSR_freq_RGB.py:298-310
    def __init__(self, dim, index):
        super().__init__()
        k = 2 * index + 1
        self.conv1 = _conv(2 * dim, dim, k, bias=False)
        self.conv2 = _conv(dim, 4, k, bias=False)
        self.relu = nn.PReLU()
        self.CA = _ChannelAttn(4, 1)


class _MGAA_RGB(_Holder):             # MGAA, CVSR_freq_RGB.py:1059-1100
    def __init__(self, dim, ACNum):
        super().__init__()
        self.convfuse = nn.Sequential(_conv(4 * dim, 2 * dim, 1, bias=False), nn.ReLU(), _conv(2 * dim, 2 * dim, 1, bias=False))
        self.MConvB = nn.ModuleList([_DenseOffsetBlk(dim, i) for i in range(ACNum)])
        self.convcrt = nn.Sequential(_conv(2 * dim, dim, 1, bias=False), nn.ReLU(), _conv(dim, 4, 1, bias=False))
        self.conv_KP = _conv(dim, dim, 3)
        self.F = nn.Sequential(_conv(dim, dim, 3), _conv(dim, ACNum * dim * 3 * 2 + ACNum * dim, 1))
        self.sigmoid = nn.Sigmoid()
        self.conv3 = _conv(2 * dim, dim, 3, bias=False)


class _Block(_Holder):                # Block, CVSR_freq_RGB.py:609-657 (width multiplier 4)
    def __init__(self, ch):
        super().__init__()
        self.body = nn.Sequential(_conv(ch, 4 * ch, 3), nn.LeakyReLU(0.1), _conv(4 * ch, ch, 3))
        _scaled_kaiming(self.body, 0.1)
        self.down = nn.Sequential(_conv(ch, ch, 1))
        self.up = nn.Sequential(_conv(ch, ch, 1))
        _scaled_kaiming(self.up, 0.1)
        _scaled_kaiming(self.down, 0.1)


class _SCGroupRGB(_Holder):           # SCGroup :659-683
    def __init__(self, ch):
        super().__init__()
        self.conv = _conv(ch, ch, 3)
        self.body = nn.Sequential(*[_Block(ch) for _ in range(3)])


class _SCNetRGB(_Holder):             # SCNet :685-700
    def __init__(self, ch, groups):
        super().__init__()
        self.body = nn.Sequential(*[_SCGroupRGB(ch) for _ in range(groups)])


class _DivEnhRGB(_Holder):            # DivEnh :1575-1612 (a, b, Conv, ca in the reference's order)
    def __init__(self, ch):
        super().__init__()
        self.Conv = _conv(ch, ch, 3)
        self.sig = nn.Sigmoid()
        self.a = nn.Parameter(torch.zeros(ch, 1, 1))
        self.b = nn.Parameter(torch.ones(ch, 1, 1))
        self.ca = _ChannelAttn(ch, 16)


class _MFFR_RGB(_Holder):             # MultiFreq_Refinment :1615-1655
    def __init__(self, ch, Freq_Inv):
        super().__init__()
        self.DivEnh_block = nn.ModuleList([_DivEnhRGB(ch) for _ in range(Freq_Inv)])
        self.ca = _ChannelAttn(ch, 16)


class _RGBBase(nn.Module):
    _SMALL = False

    def __init__(self, n_features, wiF, AC_Ks, ACNum, Freq_Inv, SCGroupN):
        super().__init__()
        if AC_Ks != 3:
            raise ValueError("AC_Ks must be 3 (the reference SAC/IAC path is only defined for 3 taps)")
        if n_features != 64:
            raise ValueError("the sm_100a kernels are built for n_features=64 (reference default)")
        n = n_features
        self.n_feats, self.wiF, self.AC_Ks = n, wiF, AC_Ks
        self.ACNum, self.Freq_Inv, self.SCGroupN = ACNum, Freq_Inv, SCGroupN
        self.in_ch = 3
        ku = 1                                # both classes use 1x1 up-convolutions (:2083-2088, :2156-2161)
        self.feat_extract = nn.Sequential(_conv(21, 7 * n, 3))
        self.lrelu = nn.PReLU()
        self.MGAA = _MGAA_RGB(n, ACNum)
        self.rconcat1 = _conv(n, n, 3, stride=2)
        self.rconcat2 = _conv(n, n, 3, stride=2)
        self.recorb1 = _SCNetRGB(n, SCGroupN)
        self.recorb0 = _conv(n, n, 3 if self._SMALL else 1)     # FCVSR_S 3x3 (:2081), FCVSR 1x1 (:2153)
        self.upconv1_L2 = _conv(n, n, ku)
        self.upconv1_L2_2 = _conv(n + n // 4, n, ku)
        self.upconv1_L3 = _conv(n, n, ku)
        self.upconv1 = _conv(n, 4 * n, ku)
        self.upconv2 = _conv(n, 4 * n, ku)
        self.pixel_shuffle = nn.PixelShuffle(2)
        self.conv_last0 = _conv(n, 3, 3)
        # FCVSR_S builds MultiFreq_Refinment(dim, Freq_Inv=8, mode="ideal") regardless of its Freq_Inv argument (:2092); FCVSR
        # passes its own (:2164, default 8)
        self.mffr_bands = 8 if self._SMALL else Freq_Inv
        self.MFFRblock = _MFFR_RGB(n, self.mffr_bands)
        self.upconv_fuse = _conv(n + n // 4 + n // 16, n, 3)
        self.compute_dtype = "tf32"       # "tf32" (tcgen05 where the shape fits) or "fp32" (CUDA-core convolutions)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if x.dim() != 5:
            raise ValueError(f"expected [B,T,C,H,W], got {tuple(x.shape)}")
        if not x.is_cuda:
            raise RuntimeError("fcvsr_b200 runs only on CUDA (sm_100a); there is no CPU fallback")
        if x.shape[1] != 7 or x.shape[2] != 3 or x.shape[3] % 4 or x.shape[4] % 4:
            raise ValueError("expected [B, 7, 3, H, W] with H and W multiples of 4")
        if x.dtype != torch.float32:
            raise TypeError("fcvsr_b200 expects float32 input")
        from . import _capi
        from .rgb_forward import forward_rgb
        _capi.lib()
        return forward_rgb(self, x, "fp32" if self.compute_dtype == "fp32" else "tf32")


class FCVSR(_RGBBase):
    """CVSR_freq_RGB.py:2135-2202."""

    def __init__(self, n_features=64, wiF=1.5, AC_Ks=3, ACNum=6, Freq_Inv=8, SCGroupN=6):
        super().__init__(n_features, wiF, AC_Ks, ACNum, Freq_Inv, SCGroupN)


class FCVSR_S(_RGBBase):
    """CVSR_freq_RGB.py:2059-2128."""
    _SMALL = True

    def __init__(self, n_features=64, wiF=1.5, AC_Ks=3, ACNum=3, Freq_Inv=4, SCGroupN=3):
        super().__init__(n_features, wiF, AC_Ks, ACNum, Freq_Inv, SCGroupN)


def seeded_state_dict_rgb(variant: str = "S", seed: int = 0):
    """Deterministic random-init weights for FCVSR_S ("S") / FCVSR ("full"), perturbing all-zero / all-one parameters like
    fcvsr_b200.arch.seeded_state_dict so that no branch is dead (DivEnh.a is zero-initialised in the reference)."""
    cls = {"S": FCVSR_S, "full": FCVSR}[variant]
    rng_state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    m = cls()
    torch.random.set_rng_state(rng_state)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for _, p in m.named_parameters():
            if bool((p == 0).all()) or bool((p == 1).all()):
                p.add_(0.1 * torch.randn(p.shape, generator=g))
    return {k: v.detach().clone() for k, v in m.state_dict().items()}
