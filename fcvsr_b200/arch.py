"""Drop-in model classes for the FCVSR per-clip x4 super-resolution forward on B200.

Mirrors the reference constructor / ``forward(lr_clip)`` surface of
``CVSR_train/arch/CVSR_freq.py``: ``GShiftNet`` (:2653-2756, "FCVSR") and ``GShiftNet_S``
(:2577-2646, "FCVSR-S"): same keyword arguments, same parameter names and shapes (the
``state_dict`` -- including the aliased ``...RCB.*`` == ``...body.3.*`` keys of ``BlockRCB``
:736,:751 and the never-used ``DivEnh.Conv`` -- loads strictly both ways), same
``forward(x[B,7,1,H,W]) -> [B,1,4H,4W]`` (unclamped).

The modules below are *parameter containers only*: none of the sub-modules has a forward.  All
arithmetic is done by the hand-written sm_100a kernels in ``fcvsr_b200/csrc`` through the C-ABI
library (``fcvsr_b200._lib``), orchestrated by ``fcvsr_b200.engine.Engine``.  There is no
PyTorch/ATen or CPU fallback: a missing library or a non-CUDA input raises.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.init as init


def _conv(cin, cout, k, stride=1, bias=True):
    return nn.Conv2d(cin, cout, k, stride, k // 2, bias=bias)


def _scaled_kaiming(mod: nn.Module, scale: float) -> None:
    """initialize_weights(net, scale) of the reference (:635-652): kaiming-normal fan_in * scale,
    zero bias, for every conv below `mod`."""
    for m in mod.modules():
        if isinstance(m, nn.Conv2d):
            init.kaiming_normal_(m.weight, a=0, mode="fan_in")
            m.weight.data *= scale
            if m.bias is not None:
                m.bias.data.zero_()


class _Holder(nn.Module):
    """A module that only owns parameters / children (no forward)."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter container: the forward runs in fcvsr_b200.engine")


class _ChannelAttn(_Holder):          # CALayer :1812-1828
    def __init__(self, ch, reduction):
        super().__init__()
        self.conv_du = nn.Sequential(_conv(ch, ch // reduction, 1, bias=False), nn.ReLU(),
                                     _conv(ch // reduction, ch, 1, bias=False), nn.Sigmoid())


class _OffsetConvBlk(_Holder):        # ConvBlk :344-357
    def __init__(self, dim, index):
        super().__init__()
        k = 2 * index + 1
        self.conv1 = _conv(dim, dim, k, bias=False)
        self.conv2 = _conv(dim, dim, k, bias=False)
        self.relu = nn.PReLU()
        self.CA = _ChannelAttn(dim, 1)


class _MGAA(_Holder):                 # MGAAbk :1365-1430
    def __init__(self, dim, ACNum):
        super().__init__()

        def mlp(*chs):
            layers = []
            for i in range(len(chs) - 1):
                layers.append(_conv(chs[i], chs[i + 1], 1, bias=False))
                if i + 2 < len(chs):
                    layers.append(nn.ReLU())
            return nn.Sequential(*layers)

        self.convfuse = mlp(4 * dim, 2 * dim, 2 * dim, 2 * dim)
        self.convcorr = mlp(2 * dim + 83, dim, dim, 4)
        self.MConvB = nn.ModuleList([_OffsetConvBlk(4, i) for i in range(ACNum)])
        self.convcrt = mlp(2 * dim, dim, 4)
        self.conv_KP = _conv(dim, dim, 3)
        self.F = nn.Sequential(_conv(dim, dim, 3), _conv(dim, ACNum * dim * 3 * 2, 1))
        self.conv3 = _conv(2 * dim, dim, 3, bias=False)


class _Context(_Holder):              # ContextBlock :657-669
    def __init__(self, ch):
        super().__init__()
        self.conv_mask = _conv(ch, 1, 1, bias=False)
        self.channel_add_conv = nn.Sequential(_conv(ch, ch, 1, bias=False), nn.LeakyReLU(0.2),
                                              _conv(ch, ch, 1, bias=False))


class _RCB(_Holder):                  # RCB :705-718
    def __init__(self, ch):
        super().__init__()
        self.body = nn.Sequential(_conv(ch, ch, 3, bias=False), nn.LeakyReLU(0.2), _conv(ch, ch, 3, bias=False))
        self.gcnet = _Context(ch)


class _BlockRCB(_Holder):             # BlockRCB :729-764 (RCB registered twice -> aliased keys)
    def __init__(self, ch):
        super().__init__()
        self.RCB = _RCB(ch)
        self.body = nn.Sequential(_conv(ch, 2 * ch, 3), nn.LeakyReLU(0.1), _conv(2 * ch, ch, 3), self.RCB)
        _scaled_kaiming(self.body, 0.1)
        self.down = nn.Sequential(_conv(ch, ch, 1))
        self.up = nn.Sequential(_conv(ch, ch, 1))
        _scaled_kaiming(self.down, 0.1)
        _scaled_kaiming(self.up, 0.1)


class _SCGroup(_Holder):              # SCGroupbk :781-795
    def __init__(self, ch):
        super().__init__()
        self.conv = _conv(ch, ch, 3)
        self.body = nn.Sequential(*[_BlockRCB(ch) for _ in range(3)])


class _SCNet(_Holder):                # SCNetbk :807-814
    def __init__(self, ch, groups):
        super().__init__()
        self.body = nn.Sequential(*[_SCGroup(ch) for _ in range(groups)])


class _DivEnh(_Holder):               # DivEnh :2104-2112 (Conv is a dead parameter, kept for the state dict)
    def __init__(self, ch):
        super().__init__()
        self.Conv = _conv(ch, ch, 3)
        self.a = nn.Parameter(torch.zeros(ch, 1, 1))
        self.b = nn.Parameter(torch.ones(ch, 1, 1))
        self.ca = _ChannelAttn(ch, 16)


class _MFFR(_Holder):                 # MultiFreq_Refinment :2183-2199
    def __init__(self, ch, Freq_Inv):
        super().__init__()
        self.DivEnh_block = nn.ModuleList([_DivEnh(ch) for _ in range(Freq_Inv)])
        self.ca = _ChannelAttn(ch, 16)


class _FCVSRBase(nn.Module):
    _SMALL = False      # 1x1 up-convolutions (GShiftNet_S only)
    _CH = 1             # image channels: 1 (Y, CVSR_train) or 3 (RGB, the mmedit backbones)

    def __init__(self, n_features, wiF, AC_Ks, ACNum, Freq_Inv, SCGroupN):
        super().__init__()
        if AC_Ks != 3:
            raise ValueError("AC_Ks must be 3 (the reference SAC/IAC path is only defined for 3 taps)")
        if n_features != 64:
            raise ValueError("the sm_100a kernels are built for n_features=64 (reference default)")
        n = n_features
        self.n_feats, self.wiF, self.AC_Ks = n, wiF, AC_Ks
        self.ACNum, self.Freq_Inv, self.SCGroupN = ACNum, Freq_Inv, SCGroupN
        ku = 1 if self._SMALL else 3          # GShiftNet_S uses 1x1 up-convs (:2600-2605)
        self.in_ch = self._CH
        self.feat_extract = nn.Sequential(_conv(7 * self._CH, 7 * n, 3))
        self.lrelu = nn.PReLU()
        self.MGAA = _MGAA(n, ACNum)
        self.rconcat1 = _conv(n, n, 3, stride=2)
        self.rconcat2 = _conv(n, n, 3, stride=2)
        self.recorb1 = _SCNet(n, SCGroupN)
        self.recorb0 = _conv(n, n, 3)
        self.upconv1_L2 = _conv(n, n, ku)
        self.upconv1_L2_2 = _conv(n + n // 4, n, ku)
        self.upconv1_L3 = _conv(n, n, ku)
        self.upconv1 = _conv(n, 4 * n, ku)
        self.upconv2 = _conv(n, 4 * n, ku)
        self.pixel_shuffle = nn.PixelShuffle(2)
        self.conv_last0 = _conv(n, self._CH, 3)
        self.MFFRblock = _MFFR(n, Freq_Inv)
        self.upconv_fuse = _conv(n + n // 4 + n // 16, n, 3)
        self._engine = None
        # "tf32": fp32 storage, TF32 tensor-core operands (the contract's fp32 mode, max-abs <= 1e-3);
        # "bf16": bf16 operand tensors, fp32 accumulate and residual streams (max-abs <= 5e-3, PSNR >= 60 dB);
        # "fp32": CUDA-core FFMA convolutions, bit-level cross-check of the other two
        self.compute_dtype = "tf32"

    # ------------------------------------------------------------------------------------------
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x [B,7,1,H,W] float32 on a CUDA device, H % 4 == W % 4 == 0 -> [B,1,4H,4W]
        (GShiftNet.forward :2688-2756).  Under torch.no_grad() the call runs the inference engine (one CUDA-graph-able launch
        sequence); when autograd is recording and the input or a parameter requires grad it runs the differentiable forward
        (fcvsr_b200.train_forward: the same kernels behind autograd Functions with backward kernels), so
        `loss(model(x), hr).backward()` works as in the reference's training loop (train_LD_freqCVSR_22.py:243-251).  The
        training forward computes in the contract's fp32 mode ("tf32": TF32 tensor-core operands, fp32 everywhere else) unless
        compute_dtype is "fp32"."""
        if x.dim() != 5:
            raise ValueError(f"expected [B,T,C,H,W], got {tuple(x.shape)}")   # reference: unpack error :2690
        if not x.is_cuda:
            raise RuntimeError("fcvsr_b200 runs only on CUDA (sm_100a); there is no CPU fallback")
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            if x.shape[1] != 7 or x.shape[2] != self.in_ch or x.shape[3] % 4 or x.shape[4] % 4:
                raise ValueError(f"expected [B, 7, {self.in_ch}, H, W] with H and W multiples of 4")
            if x.dtype != torch.float32:
                raise TypeError("fcvsr_b200 expects float32 input")
            from . import _capi
            from .train_forward import forward_train
            _capi.lib()                      # fail loudly if the kernel library is missing
            return forward_train(self, x, "fp32" if self.compute_dtype == "fp32" else "tf32")
        from .engine import Engine
        if self._engine is None or self._engine.mode != self.compute_dtype:
            self._engine = Engine(self, mode=self.compute_dtype)
        return self._engine.forward(x)


class GShiftNet(_FCVSRBase):
    """FCVSR (CVSR_freq.py:2653-2756)."""

    def __init__(self, n_features=64, wiF=1.5, AC_Ks=3, ACNum=6, Freq_Inv=8, SCGroupN=10):
        super().__init__(n_features, wiF, AC_Ks, ACNum, Freq_Inv, SCGroupN)


class GShiftNet_S(_FCVSRBase):
    """FCVSR-S (CVSR_freq.py:2577-2646)."""
    _SMALL = True

    def __init__(self, n_features=64, wiF=1.5, AC_Ks=3, ACNum=3, Freq_Inv=4, SCGroupN=4):
        super().__init__(n_features, wiF, AC_Ks, ACNum, Freq_Inv, SCGroupN)


class FCVSRNet(_FCVSRBase):
    """mmedit backbone `FCVSRNet` (mmedit_train/mmedit/models/backbones/sr_backbones/fcvsr.py:38-142): GShiftNet with RGB
    input / output -- feat_extract 21 -> 448, conv_last0 64 -> 3, forward(x[B,7,3,H,W]) -> [B,3,4H,4W].  The mmcv registry
    decorator of the reference needs mmcv; `register_mmedit_backbones()` below does the registration when mmedit is importable."""
    _CH = 3

    def __init__(self, n_features=64, wiF=1.5, AC_Ks=3, ACNum=6, Freq_Inv=8, SCGroupN=10):
        super().__init__(n_features, wiF, AC_Ks, ACNum, Freq_Inv, SCGroupN)

    def init_weights(self, pretrained=None, strict=True):
        """fcvsr.py:138-153: load a checkpoint when `pretrained` is a path, keep the constructor's init when it is None."""
        if isinstance(pretrained, str):
            ckpt = torch.load(pretrained, map_location="cpu", weights_only=True)
            sd = ckpt.get("state_dict", ckpt)
            sd = {(k[len("generator."):] if k.startswith("generator.") else k): v for k, v in sd.items()}
            self.load_state_dict(sd, strict=strict)
        elif pretrained is not None:
            raise TypeError(f'"pretrained" must be a str or None. But received {type(pretrained)}.')


class FCVSR_SNet(FCVSRNet):
    """mmedit backbone `FCVSR_SNet` (sr_backbones/fcvsr_s.py:40-): the FCVSR-S hyper-parameters with RGB I/O; unlike GShiftNet_S
    its up-convolutions are 3x3 (:65-70)."""

    def __init__(self, n_features=64, wiF=1.5, AC_Ks=3, ACNum=3, Freq_Inv=4, SCGroupN=4):
        super().__init__(n_features, wiF, AC_Ks, ACNum, Freq_Inv, SCGroupN)


def register_mmedit_backbones() -> bool:
    """Register FCVSRNet / FCVSR_SNet in mmedit's BACKBONES registry (what `@BACKBONES.register_module()` does at
    sr_backbones/fcvsr.py:37) so that `configs/restorers/fcvsr/*.py` build this implementation.  Returns False when mmedit /
    mmcv are not installed (they are not in the build image)."""
    try:
        from mmedit.models.registry import BACKBONES
    except Exception:
        return False
    for cls in (FCVSRNet, FCVSR_SNet):
        BACKBONES.register_module(module=cls, force=True)
    return True


class GShiftNet_ETC(GShiftNet):
    """GShiftNet_ETC (CVSR_freq.py:2760-2843): the FCVSR parameters applied to the 7 sliding windows of a 13-frame clip
    in one call.  forward(x[B,13,1,H,W]) -> (out_seq[B,7,1,4H,4W], x_up[B,7,1,4H,4W]) where x_up is the bilinear x4
    up-sampling of each window's centre frame (:2834-2842).  The windows are independent, so they run as one batch of
    7B clips through the same engine."""

    def forward(self, x: torch.Tensor):
        if x.dim() != 5:
            raise ValueError(f"expected [B,T,C,H,W], got {tuple(x.shape)}")
        b, t, c, h, w = x.shape
        if t < 13 or c != 1:
            raise ValueError("GShiftNet_ETC expects [B, 13, 1, H, W] (seven 7-frame windows)")
        if not x.is_cuda:
            raise RuntimeError("fcvsr_b200 runs only on CUDA (sm_100a); there is no CPU fallback")
        from . import _capi as C
        x = x.contiguous()
        clips = torch.stack([x[:, i:i + 7] for i in range(7)], dim=1).reshape(b * 7, 7, 1, h, w)
        out = super().forward(clips).view(b, 7, 1, 4 * h, 4 * w)
        centres = x[:, 3:10, 0].contiguous()                                  # centre frame of window i is frame i + 3
        up = torch.empty(b * 7, 1, 4 * h, 4 * w, device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            C.call("fcvsr_bilinear_up4", centres.data_ptr(), h * w, up.data_ptr(), b * 7, h, w,
                   torch.cuda.current_stream().cuda_stream)
        return out, up.view(b, 7, 1, 4 * h, 4 * w)


def seeded_state_dict(variant: str = "S", seed: int = 0, **kw):
    """Deterministic random-init weights shared by the tests, the golden generator and bench.py
    (SURVEY 8d): construct on CPU under torch.manual_seed(seed); every parameter whose init is
    all-zero or all-one (DivEnh.a/.b, biases zeroed by the scaled kaiming init) gets N(0, 0.1^2)
    noise from Generator(seed+1) so that no branch of the forward is dead."""
    cls = {"S": GShiftNet_S, "full": GShiftNet, "rgb": FCVSRNet, "rgb_S": FCVSR_SNet}[variant]
    rng_state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    m = cls(**kw)
    torch.random.set_rng_state(rng_state)
    g = torch.Generator().manual_seed(seed + 1)
    seen = set()
    with torch.no_grad():
        for _, p in m.named_parameters():
            if p.data_ptr() in seen:
                continue
            seen.add(p.data_ptr())
            if bool((p == 0).all()) or bool((p == 1).all()):
                p.add_(0.1 * torch.randn(p.shape, generator=g))
    return {k: v.detach().clone() for k, v in m.state_dict().items()}
