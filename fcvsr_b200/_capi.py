"""ctypes binding of the C ABI declared in include/fcvsr_b200.h (libfcvsr_b200.so).

There is deliberately no fallback: if the library is missing or a call returns a non-zero status a
RuntimeError is raised.  Pointers are passed as integers (``tensor.data_ptr()``), the stream as
``torch.cuda.current_stream().cuda_stream``.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_lib", "libfcvsr_b200.so")

ACT_NONE, ACT_RELU, ACT_LEAKY, ACT_PRELU = 0, 1, 2, 3
OK, ERR_ARG, ERR_CUDA, ERR_UNSUPPORTED = 0, -1, -2, -3

_T = {"p": ctypes.c_void_p, "i": ctypes.c_int, "f": ctypes.c_float, "d": ctypes.c_double, "l": ctypes.c_longlong, "s": ctypes.c_void_p}

# name -> argument codes, in the order of include/fcvsr_b200.h
SIGNATURES = {
    "fcvsr_conv2d_direct": "pii pp pi pi pi iiiiiii if p ii pii i s",
    "fcvsr_conv2d_tc": "pi pp pi pi pi iiiiii if p i pii i i s",
    "fcvsr_conv2d_tc_multi": "i pi pp pi pi pp iiii i f p pi i i s",
    "fcvsr_conv2d_tc_multi_w": "i pi pp pi pi pp iiii i f i i s",
    "fcvsr_fft_r2c_w": "pi p p iiii s",
    "fcvsr_fft_c2c_h": "p p p p iiii i f ii p s",
    "fcvsr_fft_c2r_w": "p pi p iiii f s",
    "fcvsr_corr_gather": "piii pi iiii i s",
    "fcvsr_corr_gather2": "piii ppi iiii i s",
    "fcvsr_offset_blocks": "ppppp pi pppp iiii s",
    "fcvsr_iac_step": "pi pi pi pi pi pi pi ii pi i iii i s",
    "fcvsr_iac_step_tc": "pi pi i pi pi pi pi pi ii pi p p iii s",
    "fcvsr_round_copy": "pi pi ii l i s",
    "fcvsr_chansum64": "pi p ii s",
    "fcvsr_reduce_finalize": "p ii f i pp p i s",
    "fcvsr_divenh_step": "ppppp i ii pppp ppp ii s",
    "fcvsr_mffr_final": "pp pi pi ii s",
    "fcvsr_context_block": "pi ppp ppp ii i s",
    "fcvsr_context_block_multi": "i pi ppp ppp i p i s",
    "fcvsr_context_pool_multi": "i pi p pp p i p i s",
    "fcvsr_context_pool_backward_multi": "i pi p pp pp i p s",
    "fcvsr_context_pool_backward_blocks": "i p",          # returns a block count, not a status
    "fcvsr_rcb_finish_multi": "i ppp ppp pp i iii s",
    "fcvsr_rcb_finish_backward_multi": "i ppp pp p i s",
    "fcvsr_level_mix_multi": "i pi pi p p pp i pp pi iii s",
    "fcvsr_rcb_finish": "ppp p ii p i p ii i i s",
    "fcvsr_level_mix": "pi pi p f pp iii pi i i i s",
    "fcvsr_pixel_shuffle": "pi pi iiii i s",
    "fcvsr_bilinear_up4": "p l p iii s",
    "fcvsr_fill_channels": "p iii f l s",
    "fcvsr_quantize_u8": "pp iiiii s",
    "fcvsr_rgb_tail": "pi p l p iiii s",
    "fcvsr_pack_clip": "p p iiii i s",
    "fcvsr_subsample2": "pi pi pi iiii i s",
    "fcvsr_conv3x3_c64_to1": "pi p f p p iii s",
    "fcvsr_conv2d_dgrad_direct": "pi p pi iiiiiii s",
    "fcvsr_conv2d_wgrad": "pi pi p iiiiiii s",
    "fcvsr_conv2d_wgrad_tc": "pi pi p iiiiii s",
    "fcvsr_conv2d_wgrad_tc_multi": "i pi pi p i pp iii s",
    "fcvsr_round_copy_dual": "ppp l s",
    "fcvsr_round_copy_dual_multi": "i ppp p s",
    "fcvsr_conv4x4": "ppp iiii s",
    "fcvsr_conv4x4_wgrad": "ppp iiii s",
    "fcvsr_pack_conv_weight": "pp iiiii s",
    "fcvsr_colsum": "pi i l pp i s",
    "fcvsr_flow_warp": "pi pi pi iiii s",
    "fcvsr_flow_warp_backward": "pi pi pi pi p iiii s",
    "fcvsr_sac": "pi pi pi iiii s",
    "fcvsr_sac_backward": "pi pi pi p pi pi iiii s",
    "fcvsr_corr_gather_backward": "piii pi pi iiii s",
    "fcvsr_psnr_ssim_u8": "pp iiii pp s",
    "fcvsr_charbonnier_loss": "pp li i f pp s",
    "fcvsr_pixel_loss": "pp l i f d pp s",
    "fcvsr_pixel_loss_backward": "pp l i f d p pp s",
    "fcvsr_adam_step": "pppp p i dddd d i s",
    "fcvsr_charbonnier_loss_backward": "pp li i f pp pp s",
    "fcvsr_modulated_deform_conv_forward": "ppppp p iiii i ii ii ii ii ii ll i s",
    "fcvsr_modulated_deform_conv_backward": "ppppp ppppp iiii i ii ii ii ii ii p s",
    "fcvsr_modulated_deform_conv_forward_tc": "ppppp p iiii i ii ii ii ii ii ll i p i s",
    "fcvsr_nchw_to_nhwc": "pp iiii i s",
}

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"fcvsr_b200: {LIB_PATH} is missing -- build it with `python -m fcvsr_b200.build` "
                "(there is no CPU / PyTorch fallback)")
        import torch  # noqa: F401  (loads libcudart before our library resolves it)
        _lib = ctypes.CDLL(LIB_PATH)
        for name, sig in SIGNATURES.items():
            fn = getattr(_lib, name)
            fn.restype = ctypes.c_int
            fn.argtypes = [_T[c] for c in sig.replace(" ", "")]
        _lib.fcvsr_version.restype = ctypes.c_char_p
        _lib.fcvsr_version.argtypes = []
    return _lib


_ERR = {ERR_ARG: "invalid argument", ERR_CUDA: "CUDA error", ERR_UNSUPPORTED: "unsupported shape"}


def call(name: str, *args) -> None:
    rc = getattr(lib(), name)(*args)
    if rc != 0:
        raise RuntimeError(f"{name} failed: {_ERR.get(rc, rc)}")


def try_call(name: str, *args) -> int:
    """Returns the status; used where ERR_UNSUPPORTED selects another kernel."""
    return getattr(lib(), name)(*args)


def version() -> str:
    return lib().fcvsr_version().decode()
