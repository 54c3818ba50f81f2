"""Host-side orchestration of the FCVSR forward on the sm_100a kernel library.

``Engine`` owns (1) the packed weights (reference OIHW fp32 -> the layouts the kernels read, with
the channel permutations that let concatenations / pixel shuffles / the [imag|real] spectrum
packing of the reference disappear), (2) a per-shape workspace of NHWC device buffers, and (3) the
launch sequence that restates ``GShiftNet.forward`` (CVSR_train/arch/CVSR_freq.py:2688-2756) kernel
by kernel.  Every arithmetic step is a call into ``libfcvsr_b200.so`` through the C ABI
(include/fcvsr_b200.h); no ATen op touches the data path.  The whole launch sequence can be captured
into a CUDA graph (``use_graph``), which removes the Python launch overhead of the 378 launches (FCVSR).

Line numbers in comments refer to CVSR_train/arch/CVSR_freq.py.
"""
from __future__ import annotations

import ctypes
import os

from typing import Dict, Optional

import torch

from . import _capi as C
from . import bands

F32 = torch.float32


def _spec_perm(n: int = 64) -> torch.Tensor:
    """reference xk_f channel j (cat[imag, real], :1456-1465) -> float index in our complex-interleaved
    spectrum (2c real, 2c+1 imag)."""
    j = torch.arange(2 * n)
    return torch.where(j < n, 2 * j + 1, 2 * (j - n))


def _round_tf32(t: torch.Tensor) -> torch.Tensor:
    """Round fp32 to the nearest TF32 value (ties away from zero), keeping fp32 storage."""
    bits = t.contiguous().view(torch.int32)
    return ((bits + 0x1000) & -8192).view(torch.float32)


class _ConvPack:
    """One convolution's weights in both kernel layouts."""

    def __init__(self, w: torch.Tensor, b: Optional[torch.Tensor], stride: int = 1, ps: bool = False,
                 cin_pad: Optional[int] = None, in_perm: Optional[torch.Tensor] = None,
                 out_perm: Optional[torch.Tensor] = None, op16: bool = False):
        """op16: pack the tensor-core weights as bf16 with Cin padded to a multiple of 64 (one 128-byte operand row);
        otherwise TF32-rounded fp32 with Cin padded to a multiple of 32 when `cin_pad` asks for it."""
        w = w.detach().to(F32)
        self.op16 = op16
        if in_perm is not None:       # new input position in_perm[j] <- reference input channel j
            w2 = torch.zeros_like(w)
            w2[:, in_perm] = w
            w = w2
        if out_perm is not None:      # new output position out_perm[j] <- reference output channel j
            w2 = torch.zeros_like(w)
            w2[out_perm] = w
            w = w2
            if b is not None:
                b2 = torch.zeros_like(b)
                b2[out_perm] = b.detach()
                b = b2
        cout, cin, k, _ = w.shape
        self.cin_logical = cin
        if ps:                        # GEMM column ij*C4 + c <- reference channel c*4 + ij (pixel_shuffle)
            c4 = cout // 4
            idx = (torch.arange(c4).view(1, c4) * 4 + torch.arange(4).view(4, 1)).reshape(-1).to(w.device)
            w = w[idx]
            if b is not None:
                b = b.detach()[idx]
        if op16 and cin_pad is None and cin % 64 and cin >= 32:
            cin_pad = -(-cin // 64) * 64
        if cin_pad is not None and cin_pad > cin:
            w = torch.cat([w, w.new_zeros(cout, cin_pad - cin, k, k)], 1)
            cin = cin_pad
        self.cin, self.cout, self.k, self.stride, self.ps = cin, cout, k, stride, ps
        self.bias = b.detach().to(F32).contiguous() if b is not None else None
        self.w_direct = w.permute(2, 3, 1, 0).contiguous()                 # [k*k][Cin][Cout]
        wt = w.permute(0, 2, 3, 1).reshape(cout, k * k * cin)              # [Cout][k*k*Cin]
        if cout < 16:
            wt = torch.cat([wt, wt.new_zeros(16 - cout, wt.shape[1])], 0)
        # tcgen05 kind::tf32 truncates its operands to 10 mantissa bits; rounding the weights to
        # nearest here (cvt.rna semantics) removes their share of the truncation bias for free.
        self.w_tc = wt.contiguous().to(torch.bfloat16) if op16 else _round_tf32(wt.contiguous())
        self.tc_ok = stride == 1 and k in (1, 3) and cin % (64 if op16 else 32) == 0 and (cout < 16 or cout % 16 == 0)


class Engine:
    def __init__(self, model, use_tc: bool = True, mode: Optional[str] = None):
        """mode: "fp32" (CUDA-core FFMA convolutions, exact fp32), "tf32" (tcgen05, fp32 storage, TF32 operands;
        the contract's fp32 mode) or "bf16" (tcgen05, bf16 operand tensors, fp32 accumulate / residual streams)."""
        self.model = model
        self.mode = mode or ("tf32" if use_tc else "fp32")
        if self.mode not in ("fp32", "tf32", "bf16"):
            raise ValueError(f"unknown compute mode {self.mode!r}")
        self.use_tc = self.mode != "fp32"
        self.op16 = self.mode == "bf16"
        self.OPD = torch.bfloat16 if self.op16 else F32       # dtype of tensors that only feed tensor-core convs
        self.E = 2 if self.op16 else 4
        self.cc_ld = 256 if self.op16 else 224                # convcorr[0] input: 128 + 81 channels padded to the K chunk
        self.cat_ld = 128 if self.op16 else 96                # 80 / 84-channel concat buffers of the tail
        self.packs: Optional[Dict[str, object]] = None
        self._pack_key = None
        self._ws: Dict[tuple, Dict[str, torch.Tensor]] = {}
        self.launches = 0            # kernel launches issued by the last forward (for bench.py)
        self.tc_launches = 0
        self.use_graph = False       # replay the launch sequence from a CUDA graph (one per input shape)
        self._graphs: Dict[tuple, tuple] = {}
        self.max_ctas = 0            # grid cap of the tensor-core convs on the current stream (0 = all SMs)
        self.multi_stream = True     # run the three pyramid levels of SCNetbk on three streams
        # layout choices of the bf16 mode, each measured as an A/B in round 1 (profiles/r1_notes.md 6b); plain attributes so
        # that the tools under tools/ can flip them -- nothing in the product path reads the environment
        self.tc_pyramid = True       # rconcat1/2 as stride-1 tcgen05 convs + sampling
        self.iac16 = True            # IAC ping-pong tensors in bf16
        self.iac_tc = True           # IAC taps computed on chip (fcvsr_iac_step_tc), no Pred_K tensor
        self.iac_tc_tf32 = True      # ... also in the fp32-contract mode (TF32 operands, fp32 ping-pong)
        self.res16 = True            # RCB body output as a bf16 tensor
        self.r016 = True             # RCB input / skip r0 only as a bf16 tensor
        self.rr16 = True             # RCB output rr only as a bf16 tensor
        self.t16 = True              # cross-level terms td / tu as bf16 tensors
        self.x16 = True              # x carried through a group's three BlockRCBs only as the bf16 operand copy
        self.use_last_kernel = True  # dedicated Cout = 1 kernel
        self.merge_down_up = True    # the 1x1 down / up convolutions of a BlockRCB as one launch (stacked filters)
        self.mgaa_ctas = 74          # SM cap of each of the two concurrently running MGAA calls
        self.stop_after = None       # tools/gpu_phase_times.py: return after "mgaa_pair" | "mgaa" | "mffr" | "scnet"
        self.profile_flavor = False  # tools: per-launch profile entries also name the output flavour
        self.clone_output = True     # graph mode: return a copy of the static output buffer (False: the buffer itself,
                                     # valid until the next call)
        self._streams = {}
        self.profile = None          # optional list: (kind, flops, bytes, start_event, end_event) per launch
        self.profile_in_graph = False  # the events are graph nodes (external events) and the pyramid streams stay on
        C.lib()                      # fail loudly now if the library is missing

    # -------------------------------------------------------------------------------------------
    # weights
    # -------------------------------------------------------------------------------------------
    def _key(self):
        return tuple((p.data_ptr(), p._version) for p in self.model.parameters())

    def _ensure_packs(self, device):
        key = (str(device),) + self._key()
        if self.packs is not None and key == self._pack_key:
            return
        m = self.model
        sd = {k: v.detach().to(device=device, dtype=F32) for k, v in m.state_dict().items()}
        P: Dict[str, object] = {}
        A = m.ACNum
        n = m.n_feats
        pi = _spec_perm(n).to(device)

        op16 = self.op16
        kch = 64 if op16 else 32

        def cp(name, **kw):
            return _ConvPack(sd[name + ".weight"], sd.get(name + ".bias"), op16=op16, **kw)

        P["feat"] = cp("feat_extract.0")
        # tensor-core path: clip packed to one NHWC-32 chunk; kept on TF32 operands in bf16 mode too (the 8-bit
        # pixel values / 255 would lose three bits as bf16 and the conv is 1 % of the FLOPs)
        P["feat_tc"] = _ConvPack(sd["feat_extract.0.weight"], sd["feat_extract.0.bias"], cin_pad=32)
        # --- MGAA per-bin MLPs (:1371-1396) on the interleaved spectrum ---
        ar = torch.arange(2 * n, device=device)
        perm_f = torch.cat([pi, 2 * n + pi])                     # cat[x1_f, x2_f] -> [grp0 | grp1]
        perm_b = torch.cat([2 * n + pi, pi])                     # cat[x3_f, x2_f] -> [grp1 | grp2] slice
        P["fuse0_f"] = cp("MGAA.convfuse.0", in_perm=perm_f)
        P["fuse0_b"] = cp("MGAA.convfuse.0", in_perm=perm_b)
        P["fuse2"] = cp("MGAA.convfuse.2")
        P["fuse4"] = cp("MGAA.convfuse.4", out_perm=pi)
        P["crt0"] = cp("MGAA.convcrt.0", in_perm=pi)
        P["crt2"] = cp("MGAA.convcrt.2")
        wcc = sd["MGAA.convcorr.0.weight"][:, : 2 * n + 81]       # the 2 flow channels are zeros (:1484-1485)
        perm_cc = torch.cat([pi, 2 * n + torch.arange(81, device=device)])
        P["corr0"] = _ConvPack(wcc, None, in_perm=perm_cc, cin_pad=self.cc_ld, op16=op16)
        P["corr2"] = cp("MGAA.convcorr.2")
        P["corr4"] = cp("MGAA.convcorr.4")
        # ConvBlk stack (:344-357): weights [tap][ci][co] per iteration, back to back
        w1 = [sd[f"MGAA.MConvB.{i}.conv1.weight"].permute(2, 3, 1, 0).reshape(-1) for i in range(A)]
        w2 = [sd[f"MGAA.MConvB.{i}.conv2.weight"].permute(2, 3, 1, 0).reshape(-1) for i in range(A)]
        P["ob_w1"] = torch.cat(w1).contiguous()
        P["ob_w2"] = torch.cat(w2).contiguous()
        P["ob_prelu"] = torch.cat([sd[f"MGAA.MConvB.{i}.relu.weight"].reshape(1) for i in range(A)]).contiguous()
        P["ob_ca"] = torch.cat([torch.cat([sd[f"MGAA.MConvB.{i}.CA.conv_du.0.weight"].reshape(-1),
                                           sd[f"MGAA.MConvB.{i}.CA.conv_du.2.weight"].reshape(-1)])
                                for i in range(A)]).contiguous()
        P["kp"] = cp("MGAA.conv_KP")
        P["F0"] = cp("MGAA.F.0")
        # F.1: keep only the live rows i*384 + c*3 + t (:1231-1235, SAC uses kernel1 twice :1272-1273),
        # re-ordered to [i][t][c] so the IAC kernel reads 64 contiguous channels per tap
        ii, tt, cc = torch.meshgrid(torch.arange(A), torch.arange(3), torch.arange(n), indexing="ij")
        rows = (ii * 6 * n + cc * 3 + tt).reshape(-1).to(device)
        P["F1"] = _ConvPack(sd["MGAA.F.1.weight"][rows], sd["MGAA.F.1.bias"][rows], op16=op16)
        if self.use_tc:
            # the same live rows per iteration as the B operand of fcvsr_iac_step_tc: row c4*12 + t*4 + cc, channel c = 4 c4 + cc
            # (bf16, or TF32-rounded fp32 in the fp32-contract mode)
            ii, gg, tt, cc = torch.meshgrid(torch.arange(A), torch.arange(n // 4), torch.arange(3), torch.arange(4), indexing="ij")
            rows_tc = (ii * 6 * n + (gg * 4 + cc) * 3 + tt).reshape(-1).to(device)
            w_tc = sd["MGAA.F.1.weight"][rows_tc].reshape(A, 3 * n, n).float().contiguous()
            P["F1tc_w"] = w_tc.to(torch.bfloat16).contiguous() if op16 else _round_tf32(w_tc)
            P["F1tc_b"] = sd["MGAA.F.1.bias"][rows_tc].float().reshape(A, 3 * n).contiguous()
        P["conv3"] = cp("MGAA.conv3")
        # --- MFFR (:2104-2133) ---
        for i in range(m.Freq_Inv):
            P[f"de{i}.a"] = sd[f"MFFRblock.DivEnh_block.{i}.a"].reshape(-1).contiguous()
            P[f"de{i}.b"] = sd[f"MFFRblock.DivEnh_block.{i}.b"].reshape(-1).contiguous()
            P[f"de{i}.w1"] = sd[f"MFFRblock.DivEnh_block.{i}.ca.conv_du.0.weight"].reshape(4, n).contiguous()
            P[f"de{i}.w2"] = sd[f"MFFRblock.DivEnh_block.{i}.ca.conv_du.2.weight"].reshape(n, 4).contiguous()
        P["mffr.w1"] = sd["MFFRblock.ca.conv_du.0.weight"].reshape(4, n).contiguous()
        P["mffr.w2"] = sd["MFFRblock.ca.conv_du.2.weight"].reshape(n, 4).contiguous()
        P["rc1"] = cp("rconcat1", stride=2)
        P["rc2"] = cp("rconcat2", stride=2)
        # tensor-core path: the same weights as stride-1 convolutions, sampled at the even pixels afterwards (tail.cu)
        P["rc1_s1"] = cp("rconcat1")
        P["rc2_s1"] = cp("rconcat2")
        # --- SCNetbk (:705-822) ---
        for g in range(m.SCGroupN):
            P[f"g{g}.conv"] = cp(f"recorb1.body.{g}.conv")
            for k in range(3):
                pre = f"recorb1.body.{g}.body.{k}"
                q = f"g{g}.b{k}."
                P[q + "c0"] = cp(pre + ".body.0")
                P[q + "c2"] = cp(pre + ".body.2")
                P[q + "r0"] = cp(pre + ".RCB.body.0")
                P[q + "r2"] = cp(pre + ".RCB.body.2")
                P[q + "mask"] = sd[pre + ".RCB.gcnet.conv_mask.weight"].reshape(-1).contiguous()
                P[q + "a0"] = sd[pre + ".RCB.gcnet.channel_add_conv.0.weight"].reshape(n, n).contiguous()
                P[q + "a2"] = sd[pre + ".RCB.gcnet.channel_add_conv.2.weight"].reshape(n, n).contiguous()
                P[q + "down"] = cp(pre + ".down.0")
                P[q + "up"] = cp(pre + ".up.0")
                if self.use_tc:       # both 1x1 filters stacked: one launch runs `down` on two levels and `up` on two levels
                    dn, up = P[q + "down"], P[q + "up"]
                    zb = torch.zeros(n, device=device, dtype=F32)
                    P[q + "du_w"] = torch.cat([dn.w_tc, up.w_tc], 0).contiguous()
                    P[q + "du_b"] = torch.cat([dn.bias if dn.bias is not None else zb, up.bias if up.bias is not None else zb]).contiguous()
        # --- tail (:2739-2749) ---
        c4 = n // 4
        ps_pos = torch.empty(n, dtype=torch.long, device=device)   # position of reference channel c*4+ij
        ps_pos[(torch.arange(c4).view(1, c4) * 4 + torch.arange(4).view(4, 1)).reshape(-1).to(device)] = torch.arange(n, device=device)
        P["up_l3"] = cp("upconv1_L3", ps=True)
        # u2 is kept in pixel-shuffle order (position ij*16+c) so that `u2 + conv(cat)` lines up with the
        # shuffled GEMM columns of upconv1_L2_2 (:2743)
        P["up_l2"] = cp("upconv1_L2", out_perm=ps_pos)
        perm_l22 = torch.cat([ps_pos, n + torch.arange(c4, device=device)])
        P["up_l2_2"] = cp("upconv1_L2_2", in_perm=perm_l22, cin_pad=self.cat_ld, ps=True)
        P["fuse"] = cp("upconv_fuse", cin_pad=self.cat_ld)
        P["rec0"] = cp("recorb0")
        P["up1"] = cp("upconv1", ps=True)
        P["up2"] = cp("upconv2", ps=True)
        P["last"] = cp("conv_last0")
        if self.op16:                 # Cout = 1: CUDA-core kernel, weights as kernel parameters (host copy [ky][kx][c])
            wl = sd["conv_last0.weight"].cpu()                                          # [1, 64, 3, 3]
            self._last_w = (ctypes.c_float * 576)(*wl[0].permute(1, 2, 0).reshape(-1).tolist())
            self._last_b = float(sd["conv_last0.bias"].cpu()[0])
        P["prelu"] = sd["lrelu.weight"].reshape(1).contiguous()
        self.packs = P
        self._pack_key = key

    # -------------------------------------------------------------------------------------------
    # workspace
    # -------------------------------------------------------------------------------------------
    def _iac_on_chip(self) -> bool:
        """fcvsr_iac_step_tc computes the taps from kp2 in its own prologue: bf16 mode with bf16 ping-pong tensors, or the
        fp32-contract mode with TF32 operands and fp32 ping-pong tensors."""
        if not (self.use_tc and self.iac_tc and self.model.n_feats == 64):
            return False
        return bool(self.iac16) if self.op16 else bool(self.iac_tc_tf32)

    def _workspace(self, B, H, W, device):
        key = (B, H, W, str(device), self._iac_on_chip())
        if key in self._ws:
            return self._ws[key]
        m = self.model
        A, Q = m.ACNum, m.Freq_Inv
        Wf = W // 2 + 1
        P, Pf = H * W, H * Wf
        ws: Dict[str, torch.Tensor] = {}
        OPD = self.OPD

        def buf(name, *shape, zero=False, op=False):
            """op=True: tensor-core operand tensor (TF32-rounded fp32, or bf16 in bf16 mode)."""
            assert name not in ws, f"workspace name collision: {name}"
            ws[name] = (torch.zeros if zero else torch.empty)(*shape, device=device, dtype=OPD if op else F32)

        buf("feat", B, P, 448)
        buf("clip", B, P, 32)
        for sfx in ("", "_b"):         # two scratch sets: MGAA(f1) and MGAA(f3) run concurrently
            buf("spec" + sfx, B, Pf, 384)
            if self.op16:
                buf("spec_op" + sfx, B, Pf, 384, op=True)
            buf("h1" + sfx, 2 * B, Pf, 128, op=True)
            buf("h2" + sfx, 2 * B, Pf, 128, op=True)
            buf("cc" + sfx, 2 * B, Pf, self.cc_ld, zero=True, op=True)
            buf("c1" + sfx, 2 * B, Pf, 64, op=True)
            buf("c2" + sfx, 2 * B, Pf, 64, op=True)
            buf("off" + sfx, 2 * B, Pf, 4)
            buf("simh" + sfx, B, Pf, 64, op=True)
            buf("sim" + sfx, B, Pf, 4)
            buf("ob_t1" + sfx, A, 2 * B, Pf, 4)
            buf("ob_t2" + sfx, A, 2 * B, Pf, 4)
            buf("ob_partial" + sfx, A * 2 * B * max((Pf + 127) // 128, ((Wf + 63) // 64) * ((H + 7) // 8)) * 4)
            buf("z" + sfx, B, Pf, 8 * A)
            buf("offs" + sfx, B, P, 4 * A)
            buf("x2r" + sfx, B, P, 64, op=True)
            buf("kp1" + sfx, B, P, 64, op=True)
            buf("kp2" + sfx, B, P, 64, op=True)
            # per-pixel filter taps (the largest tensor): fp16 in the tensor-core modes (10-bit mantissa = TF32)
            if not self._iac_on_chip():
                ws["pk" + sfx] = torch.empty(B, P, A * 192, device=device, dtype=torch.float16 if self.use_tc else F32)
            buf("ping" + sfx, 2, 2, B, P, 64)
            buf("cat128" + sfx, B, P, 128, op=self.use_tc)
        buf("m2", B, P, 64)
        buf("specx", B, Pf, 128)
        buf("tmpc", Q, B, Pf, 128)
        buf("bands", Q, B, P, 64)
        buf("sb", B, P, 64)
        buf("so", B, P, 64)
        buf("mf_partial", B * ((P + 255) // 256) * 128)
        buf("mean0", B, 128)
        buf("gates", Q + 1, B, 128)
        dims = [(H, W), (H // 2, W // 2), (H // 4, W // 4)]
        for l, (h, w) in enumerate(dims):
            p = h * w
            for nm in ("xs", "cur", "t", "r0", "res", "rr"):          # fp32 streams / residual inputs
                buf(f"{nm}{l}", B, p, 64)
            for nm in ("xsr", "curr", "tr", "r0h", "rrh", "c1"):      # tensor-core operand copies / conv-only tensors
                buf(f"{nm}{l}", B, p, 64, op=True)
            buf(f"a128_{l}", B, p, 128, op=True)
            buf(f"td{l}", B, p // 4, 64)                           # down conv output, already at the next level's size
            buf(f"rrp{l}", B, p // 4, 64, op=self.use_tc)          # 2x2 mean of rr: input of the down conv
            buf(f"tu{l}", B, p, 64)
        buf("ctxall", sum(B * ((h * w + 127) // 128) * 66 for h, w in dims))       # ContextBlock partials of the three levels
        buf("addall", 3, B, 64)
        ws["ctxcnt"] = torch.zeros(3 * B, device=device, dtype=torch.int32)      # per (level, image) block counters, self-resetting
        p2, p3 = dims[1][0] * dims[1][1], dims[2][0] * dims[2][1]
        buf("o2", B, p2, 64, op=self.use_tc)
        buf("o3", B, p3, 64, op=self.use_tc)
        buf("u2", B, p2, 64)                                       # out_L2 in full precision (residual of upconv1_L2_2)
        buf("cat2", B, p2, self.cat_ld, zero=True, op=self.use_tc)
        buf("fuse", B, P, self.cat_ld, zero=True, op=self.use_tc)
        buf("f1", B, P, 64, op=self.use_tc)
        buf("f2", B, P, 64, op=self.use_tc)
        buf("up1", B, 4 * P, 64, op=self.use_tc)
        buf("up2", B, 16 * P, 64, op=self.use_tc)
        buf("base", B, 16 * P)
        if m.in_ch > 1:
            buf("lastc", B, 16 * P, 4)
        ws["masks"] = bands.symmetric_half_masks(Q, H, W, device)
        ws["tw_w"] = bands.twiddles(W, device)
        ws["tw_h"] = bands.twiddles(H, device)
        self._ws[key] = ws
        return ws

    # -------------------------------------------------------------------------------------------
    # launch helpers
    # -------------------------------------------------------------------------------------------
    def _conv(self, pk: _ConvPack, x, ldx, y, ldy, B, H, W, act=C.ACT_NONE, slope=0.0, slope_ptr=0, res=0, ldres=0,
              res2=0, ldres2=0, nchw=False, cin=None, y2=0, ldy2=0, rnd=False, op16=None):
        """x, y, res*: integer device addresses.  `cin`: logical Cin for the direct kernel when the pack
        was padded for the tensor-core path."""
        st = self.st
        self.launches += 1
        prof = self.profile
        if prof is not None:
            ho, wo = (H - 1) // pk.stride + 1, (W - 1) // pk.stride + 1
            e0, e1 = self._event_pair()
            e0.record()
            try:
                self.profile = None
                self._conv(pk, x, ldx, y, ldy, B, H, W, act, slope, slope_ptr, res, ldres, res2, ldres2, nchw, cin, y2,
                           ldy2, rnd, op16)
            finally:
                self.profile = prof
            self.launches -= 1
            e1.record()
            kind = "tc" if (self.use_tc and pk.tc_ok and not nchw) else "direct"
            if kind == "tc":
                kind = f"tc {pk.cin_logical}->{pk.cout} k{pk.k} {H}x{W}"
                if self.profile_flavor:
                    kind += f" out={'op' if rnd is True or rnd == 1 else ('f16' if rnd == 2 else 'f32')}{'+y2' if y2 else ''}{'+res' if res else ''}"
            prof.append((kind, 2.0 * B * ho * wo * pk.cin_logical * pk.cout * pk.k * pk.k,
                         4.0 * B * (H * W * pk.cin_logical + ho * wo * pk.cout), e0, e1))
            return
        o16 = int(self.op16 if op16 is None else op16)
        if not self.use_tc:          # exact-fp32 mode: nothing is rounded to TF32
            y2, ldy2, rnd = 0, 0, False
        if self.op16 and (rnd or y2) and not (pk.tc_ok and not nchw) and rnd:
            raise RuntimeError("bf16 operand output requested from a convolution the tensor-core kernel cannot run")
        if self.use_tc and pk.tc_ok and not nchw:
            rc = C.try_call("fcvsr_conv2d_tc", x, ldx, pk.w_tc.data_ptr(), pk.bias.data_ptr() if pk.bias is not None else 0,
                            res, ldres, res2, ldres2, y, ldy, B, H, W, pk.cin, pk.cout, pk.k, act, slope, slope_ptr,
                            int(pk.ps), y2, ldy2, int(rnd), self.max_ctas, o16, st)
            if rc == 0:
                self.tc_launches += 1
                return
            if rc != C.ERR_UNSUPPORTED or o16:           # bf16 operand tensors cannot fall back to the fp32 kernel
                raise RuntimeError(f"fcvsr_conv2d_tc failed with status {rc}")
        C.call("fcvsr_conv2d_direct", x, ldx, int(nchw), pk.w_direct.data_ptr(),
               pk.bias.data_ptr() if pk.bias is not None else 0, res, ldres, res2, ldres2, y, ldy, B, H, W, pk.cin,
               pk.cout, pk.k, pk.stride, act, slope, slope_ptr, int(pk.ps), 0, y2, ldy2, int(rnd and not self.op16),
               int(self.op16), st)

    def _conv_multi(self, pk: _ConvPack, xs, ldx, ys, ldy, B, dims, act=C.ACT_NONE, slope=0.0, res=None, ldres=0, y2=None,
                    ldy2=0, rnd=False):
        """The same convolution on several tensors of different spatial size (pyramid levels) in one launch.  xs / ys / res /
        y2: lists of device addresses, dims: list of (H, W)."""
        n = len(xs)
        if not (self.use_tc and pk.tc_ok):          # exact-fp32 mode (CUDA-core kernel): one launch per level
            for i in range(n):
                self._conv(pk, xs[i], ldx, ys[i], ldy, B, dims[i][0], dims[i][1], act=act, slope=slope,
                           res=res[i] if res else 0, ldres=ldres, y2=y2[i] if y2 else 0, ldy2=ldy2, rnd=rnd)
            return
        self.launches += 1
        self.tc_launches += 1
        vp = lambda ptrs: (ctypes.c_void_p * n)(*ptrs)  # noqa: E731
        args = (n, vp(xs), ldx, pk.w_tc.data_ptr(), pk.bias.data_ptr() if pk.bias is not None else 0, vp(res) if res else None,
                ldres, vp(ys), ldy, (ctypes.c_int * n)(*[d[0] for d in dims]), (ctypes.c_int * n)(*[d[1] for d in dims]), B,
                pk.cin, pk.cout, pk.k, act, slope, 0, vp(y2) if y2 else None, ldy2, int(rnd), int(self.op16), self.st)
        prof = self.profile
        if prof is not None:
            e0, e1 = self._event_pair()
            e0.record()
            C.call("fcvsr_conv2d_tc_multi", *args)
            e1.record()
            npix = sum(h * w for h, w in dims)
            prof.append((f"tc {pk.cin_logical}->{pk.cout} k{pk.k} pyramid x{n}", 2.0 * B * npix * pk.cin_logical * pk.cout * pk.k * pk.k,
                         4.0 * B * npix * (pk.cin_logical + pk.cout), e0, e1))
            return
        C.call("fcvsr_conv2d_tc_multi", *args)

    def _conv_down_up(self, w, b, xs, ys, B, dims, rnd):
        """The 1x1 `down` (filter rows 0..63) and `up` (rows 64..127) convolutions of a BlockRCB on two pyramid levels each, one launch."""
        n = len(xs)
        self.launches += 1
        self.tc_launches += 1
        vp = lambda ptrs: (ctypes.c_void_p * n)(*ptrs)  # noqa: E731
        args = (n, vp(xs), 64, w.data_ptr(), b.data_ptr(), (ctypes.c_int * n)(0, 0, 64, 64), 128, vp(ys), 64,
                (ctypes.c_int * n)(*[d[0] for d in dims]), (ctypes.c_int * n)(*[d[1] for d in dims]), B, 64, 64, 1, C.ACT_NONE, 0.0,
                int(rnd), int(self.op16), self.st)
        prof = self.profile
        if prof is not None:
            e0, e1 = self._event_pair()
            e0.record()
            C.call("fcvsr_conv2d_tc_multi_w", *args)
            e1.record()
            npix = sum(h * w_ for h, w_ in dims)
            prof.append(("tc 64->64 k1 down+up x4", 2.0 * B * npix * 64 * 64, 4.0 * B * npix * 128, e0, e1))
            return
        C.call("fcvsr_conv2d_tc_multi_w", *args)

    def _event_pair(self):
        # inside a graph capture only "external" events become event-record nodes whose times can be read after a replay
        kw = dict(enable_timing=True, external=True) if self.profile_in_graph else dict(enable_timing=True)
        return torch.cuda.Event(**kw), torch.cuda.Event(**kw)

    def profile_graph_replay(self, x: torch.Tensor, reps: int = 5):
        """Per-launch device times of the GRAPH-REPLAYED step: the launch sequence is captured once more with an event-record
        node before and after every launch (on the stream it is launched on; the pyramid / MGAA streams stay concurrent, so
        the times include the interference between concurrently running kernels), replayed `reps` times, and the elapsed
        times of each pair are averaged.  Returns [(kind, flops, bytes, mean_ms)] in launch order plus the mean replay time.
        Differences to the production graph: the event nodes cut the programmatic-dependent-launch edges between
        back-to-back convolutions (~2 us per launch)."""
        B, T, Cc, H, W = x.shape
        dev = x.device
        with torch.cuda.device(dev), torch.no_grad():
            self._dev = dev
            self._ensure_packs(dev)
            ws = self._workspace(B, H, W, dev)
            sx = x.clone()
            so = torch.empty(B, Cc, 4 * H, 4 * W, device=dev, dtype=F32)
            self.st = torch.cuda.current_stream().cuda_stream
            self._run(sx, so, ws, B, H, W)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            self.profile, self.profile_in_graph = [], True
            try:
                with torch.cuda.graph(g):
                    self.st = torch.cuda.current_stream().cuda_stream
                    self._run(sx, so, ws, B, H, W)
            finally:
                prof, self.profile, self.profile_in_graph = self.profile, None, False
            g.replay()
            torch.cuda.synchronize()
            acc = [0.0] * len(prof)
            total = 0.0
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for _ in range(reps):
                t0.record()
                g.replay()
                t1.record()
                torch.cuda.synchronize()
                total += t0.elapsed_time(t1)
                for i, (_, _, _, e0, e1) in enumerate(prof):
                    acc[i] += e0.elapsed_time(e1)
        return [(k, fl, by, a / reps) for (k, fl, by, _, _), a in zip(prof, acc)], total / reps

    def _k(self, name, *args):
        self.launches += 1
        prof = self.profile
        if prof is not None:        # per-launch CUDA events (bench.py roofline pass / tools/gpu_breakdown.py)
            e0, e1 = self._event_pair()
            e0.record()
            C.call(name, *args, self.st)
            e1.record()
            prof.append((name, 0.0, 0.0, e0, e1))
            return
        C.call(name, *args, self.st)

    # -------------------------------------------------------------------------------------------
    # forward
    # -------------------------------------------------------------------------------------------
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        B, T, Cc, H, W = x.shape
        if T != 7 or Cc != self.model.in_ch:
            raise ValueError(f"expected [B, 7, {self.model.in_ch}, H, W] (clips of 7 frames)")
        if H % 4 or W % 4:
            raise ValueError("H and W must be multiples of 4 (two stride-2 levels, reference :2671-2672)")
        if x.dtype != F32:
            raise TypeError("fcvsr_b200 expects float32 input")
        x = x.contiguous()
        dev = x.device
        self._dev = dev
        with torch.cuda.device(dev):
            self._ensure_packs(dev)
            ws = self._workspace(B, H, W, dev)
            if self.use_graph:
                return self._forward_graph(x, ws, B, H, W)
            out = torch.empty(B, Cc, 4 * H, 4 * W, device=dev, dtype=F32)
            self.st = torch.cuda.current_stream().cuda_stream
            self.launches = 0
            self.tc_launches = 0
            self._run(x, out, ws, B, H, W)
        return out

    def _forward_graph(self, x, ws, B, H, W):
        """CUDA-graph replay of the same launch sequence (static input/output buffers per shape)."""
        key = (B, H, W, str(x.device), self._pack_key)
        if key not in self._graphs:
            sx = x.clone()
            so = torch.empty(B, x.shape[2], 4 * H, 4 * W, device=x.device, dtype=F32)
            self.st = torch.cuda.current_stream().cuda_stream
            self.launches = self.tc_launches = 0
            self._run(sx, so, ws, B, H, W)            # eager warm-up: function attributes, error flag, ...
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.st = torch.cuda.current_stream().cuda_stream
                self.launches = self.tc_launches = 0
                self._run(sx, so, ws, B, H, W)
            self._graphs = {k: v for k, v in self._graphs.items() if k[-1] == self._pack_key}
            self._graphs[key] = (g, sx, so, self.launches, self.tc_launches)
        g, sx, so, self.launches, self.tc_launches = self._graphs[key]
        if x.data_ptr() != sx.data_ptr():          # a caller that fills `static_input()` in place skips this copy
            sx.copy_(x)
        g.replay()
        return so.clone() if self.clone_output else so

    def static_input(self, B, H, W, device) -> Optional[torch.Tensor]:
        """Graph mode: the captured input buffer of shape [B,7,1,H,W] (None before the first call with that shape).  Writing
        the clip into it (e.g. the H2D copy of a streaming driver) and passing it to forward() avoids one device copy."""
        for k, v in self._graphs.items():
            if k[:4] == (B, H, W, str(device)):
                return v[1]
        return None

    def _run(self, x, out, ws, B, H, W):
        m, P = self.model, self.packs
        p = {k: v.data_ptr() for k, v in ws.items()}
        O16 = int(self.op16)
        f = p["feat"]
        # feat_extract (:2663): NCHW clip -> NHWC 448 channels
        if self.use_tc:
            self._k("fcvsr_pack_clip", x.data_ptr(), p["clip"], B, 7 * self.model.in_ch, H, W, 0)
            self._conv(P["feat_tc"], p["clip"], 32, f, 448, B, H, W, op16=0)
        else:
            self._conv(P["feat"], x.data_ptr(), 0, f, 448, B, H, W, nchw=True)
        # MGAA(f1) -> feat[128:192], MGAA(f3) -> feat[256:320]: cat[o1, f2, o3] (:2720) is then the
        # contiguous channel slice feat[128:320] and no concatenation is materialised.
        main = torch.cuda.current_stream()
        if self.multi_stream and (self.profile is None or self.profile_in_graph):
            dev = main.device
            if dev not in self._streams:
                self._streams[dev] = (torch.cuda.Stream(dev), torch.cuda.Stream(dev))
            side = self._streams[dev][0]
            side.wait_stream(main)
            self.max_ctas = self.mgaa_ctas                                    # two tensor-core conv grids share the 148 SMs
            self._mgaa(ws, p, f + 0 * 4, 448, f + 128 * 4, 448, B, H, W)
            with torch.cuda.stream(side):
                self.st = side.cuda_stream
                self._mgaa(ws, p, f + 256 * 4, 448, f + 256 * 4, 448, B, H, W, sfx="_b")
            self.st = main.cuda_stream
            self.max_ctas = 0
            main.wait_stream(side)
        else:
            self._mgaa(ws, p, f + 0 * 4, 448, f + 128 * 4, 448, B, H, W)
            self._mgaa(ws, p, f + 256 * 4, 448, f + 256 * 4, 448, B, H, W)
        stop = self.stop_after
        if stop == "mgaa_pair":
            return
        self._mgaa(ws, p, f + 128 * 4, 448, p["m2"], 64, B, H, W)
        if stop == "mgaa":
            return
        self._mffr(ws, p, B, H, W)                                   # m2 -> xs0 (+ operand copy xsr0)
        if stop == "mffr":
            return
        if self.use_tc and self.tc_pyramid:
            # stride 2 = stride 1 on the tensor cores + even-pixel sampling; t0 / t1 are free until SCNetbk starts
            O16 = int(self.op16)
            self._conv(P["rc1_s1"], p["xsr0"], 64, p["t0"], 64, B, H, W)                                      # :2735
            self._k("fcvsr_subsample2", p["t0"], 64, p["xs1"], 64, p["xsr1"], 64, B, H, W, 64, O16)
            self._conv(P["rc2_s1"], p["xsr1"], 64, p["t1"], 64, B, H // 2, W // 2)                            # :2736
            self._k("fcvsr_subsample2", p["t1"], 64, p["xs2"], 64, p["xsr2"], 64, B, H // 2, W // 2, 64, O16)
        else:
            self._conv(P["rc1"], p["xs0"], 64, p["xs1"], 64, B, H, W, y2=p["xsr1"], ldy2=64)                  # :2735
            self._conv(P["rc2"], p["xs1"], 64, p["xs2"], 64, B, H // 2, W // 2, y2=p["xsr2"], ldy2=64)        # :2736
        self._scnet(ws, p, B, H, W)
        if stop == "scnet":
            return
        self._tail(x, out, ws, p, B, H, W)

    # MGAAbk.forward (:1442-1547) on the 192-channel slice at `src`; result (64 ch) to `dst`.
    # `sfx` selects the scratch set ("" or "_b"): MGAA(f1) and MGAA(f3) are independent and run concurrently.
    def _mgaa(self, ws, p0, src, lds, dst, ldd, B, H, W, sfx=""):
        P, A = self.packs, self.model.ACNum
        Wf = W // 2 + 1
        Pf = H * Wf
        shared = ("tw_w", "tw_h")
        p = {k[: -len(sfx)] if sfx and k.endswith(sfx) else k: v for k, v in p0.items()
             if (not sfx) or k.endswith(sfx) or k in shared}
        spec = p["spec"]
        RELU = C.ACT_RELU
        R, O16, E = int(self.use_tc), int(self.op16), self.E
        CC = self.cc_ld
        # rfft2 of x1|x2|x3 (:1452-1454).  TF32 mode rounds the spectrum in place (it is operand and residual at
        # once); bf16 mode keeps it in fp32 and makes a bf16 operand copy.
        self._k("fcvsr_fft_r2c_w", src, lds, spec, p["tw_w"], B, H, W, 192)
        sop = p["spec_op"] if O16 else spec                      # bf16 operand copy written by the same pass
        self._k("fcvsr_fft_c2c_h", spec, spec, p["tw_h"], 0, B, H, Wf, 192, 0, 1.0, int(R and not O16), 1, sop if O16 else 0)
        h1, h2, cc = p["h1"], p["h2"], p["cc"]
        half = B * Pf
        # convfuse (:1472-1473); the diff skip rides in the last layer's epilogue (res - res2)
        self._conv(P["fuse0_f"], sop, 384, h1, 128, B, H, Wf, act=RELU, rnd=True)
        self._conv(P["fuse0_b"], sop + 128 * E, 384, h1 + half * 128 * E, 128, B, H, Wf, act=RELU, rnd=True)
        self._conv(P["fuse2"], h1, 128, h2, 128, 2 * B, H, Wf, act=RELU, rnd=True)
        self._conv(P["fuse4"], h2, 128, cc, CC, B, H, Wf, res=spec, ldres=384, res2=spec + 128 * 4, ldres2=384, rnd=True)
        self._conv(P["fuse4"], h2 + half * 128 * E, 128, cc + half * CC * E, CC, B, H, Wf,
                   res=spec + 256 * 4, ldres=384, res2=spec + 128 * 4, ldres2=384, rnd=True)
        # convcrt (:1474)
        self._conv(P["crt0"], sop + 128 * E, 384, p["simh"], 64, B, H, Wf, act=RELU, rnd=True)
        self._conv(P["crt2"], p["simh"], 64, p["sim"], 4, B, H, Wf)
        # CorrBlock lookup (:1475-1483); corr_f feeds both branches (:1487-1488)
        om = 2 if O16 else R
        self._k("fcvsr_corr_gather2", spec, 384, 0, 128, cc + 128 * E, cc + (half * CC + 128) * E, CC, B, H, Wf, 128, om | 4)
        # convcorr (:1487-1488), both branches as a batch of 2B
        self._conv(P["corr0"], cc, CC, p["c1"], 64, 2 * B, H, Wf, act=RELU, rnd=True)
        self._conv(P["corr2"], p["c1"], 64, p["c2"], 64, 2 * B, H, Wf, act=RELU, rnd=True)
        self._conv(P["corr4"], p["c2"], 64, p["off"], 4, 2 * B, H, Wf)
        # ConvBlk_i * x2_f_sim for all i (:1494-1498), then irfft2 (:1499-1505)
        z = p["z"]
        self.launches += 2
        self._k("fcvsr_offset_blocks", p["off"], P["ob_w1"].data_ptr(), P["ob_w2"].data_ptr(), P["ob_prelu"].data_ptr(),
                P["ob_ca"].data_ptr(), p["sim"], 4, p["ob_t1"], p["ob_t2"], p["ob_partial"], z, B, H, Wf, A)
        self._k("fcvsr_fft_c2c_h", z, z, p["tw_h"], 0, B, H, Wf, 4 * A, 1, 1.0, 0, 1, 0)
        self._k("fcvsr_fft_c2r_w", z, p["offs"], 4 * A, p["tw_w"], B, H, W, 4 * A, 1.0 / (H * W))
        # kernel predictor (:1522-1523)
        x2, ldx2 = src + 64 * 4, lds
        if R:                       # x2 is both the conv_KP operand and the full-precision skip of conv3
            self._k("fcvsr_round_copy", x2, lds, p["x2r"], 64, 64, 64, B * H * W, O16)
            x2, ldx2 = p["x2r"], 64
        self._conv(P["kp"], x2, ldx2, p["kp1"], 64, B, H, W, rnd=True)
        self._conv(P["F0"], p["kp1"], 64, p["kp2"], 64, B, H, W, rnd=True)
        on_chip = self._iac_on_chip()
        if not on_chip:
            self._conv(P["F1"], p["kp2"], 64, p["pk"], A * 192, B, H, W, rnd=2 if R else False)
        TE = 2 if R else 4                                # bytes per tap element
        # IAC (:1526-1527); the last iteration writes the operand-typed conv3 input
        ping = p["ping"]
        sz = B * H * W * 64 * 4
        prev_f, ldpf, prev_b, ldpb = src, lds, src + 128 * 4, lds
        iac16 = bool(O16) and self.iac16       # bf16 mode: the intermediate iterations ping-pong through bf16 tensors
        for i in range(A):
            if i == A - 1:
                nf, nb, ldn = p["cat128"], p["cat128"] + 64 * (E if R else 4), 128
            else:
                nf, nb, ldn = ping + (i % 2) * 2 * sz, ping + ((i % 2) * 2 + 1) * sz, 64
            ro = (2 if O16 else R) if i == A - 1 else (2 if iac16 else 0)     # + 4: prev is a bf16 ping buffer
            if on_chip:
                # taps = F.1 slice i of kp2 on tcgen05 inside the IAC kernel (Pred_K never materialised)
                self.tc_launches += 1
                pm = int(i > 0) if O16 else (2 + (4 if i == A - 1 else 0))      # TF32 operands; the last iteration is conv3's operand
                self._k("fcvsr_iac_step_tc", prev_f, ldpf, prev_b, ldpb, pm, src, lds, src + 128 * 4, lds, nf, ldn, nb, ldn,
                        p["offs"], 4 * A, (i * 2) * 2, (i * 2 + 1) * 2, p["kp2"], 64, P["F1tc_w"].data_ptr() + i * 192 * 64 * E,
                        P["F1tc_b"].data_ptr() + i * 192 * 4, B, H, W)
                prev_f, ldpf, prev_b, ldpb = nf, ldn, nb, ldn
                continue
            self._k("fcvsr_iac_step", prev_f, ldpf, prev_b, ldpb, src, lds, src + 128 * 4, lds, nf, ldn, nb, ldn,
                    p["offs"], 4 * A, (i * 2) * 2, (i * 2 + 1) * 2, p["pk"] + i * 192 * TE, A * 192, R, B, H, W,
                    ro | (4 if (iac16 and i > 0) else 0))
            prev_f, ldpf, prev_b, ldpb = nf, ldn, nb, ldn
        # conv3(cat) + x2 (:1529)
        self._conv(P["conv3"], p["cat128"], 128, dst, ldd, B, H, W, res=src + 64 * 4, ldres=lds)

    # MultiFreq_Refinment.forward (:2201-2254): m2 -> xs0
    def _mffr(self, ws, p, B, H, W):
        P, Q = self.packs, self.model.Freq_Inv
        Wf = W // 2 + 1
        npix = H * W
        nblk = (npix + 255) // 256
        x = p["m2"]
        self._k("fcvsr_fft_r2c_w", x, 64, p["specx"], p["tw_w"], B, H, W, 64)
        self._k("fcvsr_fft_c2c_h", p["specx"], p["specx"], p["tw_h"], 0, B, H, Wf, 64, 0, 1.0, 0, 1, 0)
        bsz = B * npix * 64 * 4
        # Split_freq (:2075-2101): band_q = irfft2(spectrum * Msym_q), all Q bands in one launch per pass
        self._k("fcvsr_fft_c2c_h", p["specx"], p["tmpc"], p["tw_h"], p["masks"], B, H, Wf, 64, 1, 1.0, 0, Q, 0)
        self._k("fcvsr_fft_c2r_w", p["tmpc"], p["bands"], 64, p["tw_w"], Q * B, H, W, 64, 1.0 / npix)
        band = lambda i: p["bands"] + (Q - 1 - i) * bsz            # freq[::-1] (:2204-2205)
        gate = lambda i: p["gates"] + i * B * 128 * 4
        inv = 1.0 / npix
        self._k("fcvsr_chansum64", band(0), 64, p["mf_partial"], B, npix)
        self._k("fcvsr_reduce_finalize", p["mf_partial"], nblk, 1, inv, 0, 0, 0, p["mean0"], B)
        de = lambda i, s: P[f"de{i}.{s}"].data_ptr()
        self._k("fcvsr_divenh_step", 0, 0, 0, 0, 0, 0, 1, 1, band(0), de(0, "a"), de(0, "b"), p["mean0"], p["sb"], p["so"],
                p["mf_partial"], B, npix)
        self._k("fcvsr_reduce_finalize", p["mf_partial"], nblk, 1, inv, 1, de(0, "w1"), de(0, "w2"), gate(0), B)
        for i in range(1, Q):
            self._k("fcvsr_divenh_step", band(i - 1), de(i - 1, "a"), de(i - 1, "b"), p["mean0"], gate(i - 1),
                    int(i - 1 == 0), 1, 0, band(i), de(i, "a"), de(i, "b"), 0, p["sb"], p["so"], p["mf_partial"], B, npix)
            self._k("fcvsr_reduce_finalize", p["mf_partial"], nblk, 2, inv, 1, de(i, "w1"), de(i, "w2"), gate(i), B)
        self._k("fcvsr_divenh_step", band(Q - 1), de(Q - 1, "a"), de(Q - 1, "b"), p["mean0"], gate(Q - 1), int(Q == 1), 2,
                0, 0, 0, 0, 0, p["sb"], p["so"], p["mf_partial"], B, npix)
        self._k("fcvsr_reduce_finalize", p["mf_partial"], nblk, 1, inv, 1, P["mffr.w1"].data_ptr(),
                P["mffr.w2"].data_ptr(), gate(Q), B)
        self._k("fcvsr_mffr_final", p["so"], gate(Q), x, 64, p["xs0"], 64, B, npix)
        if self.use_tc:
            self._k("fcvsr_round_copy", p["xs0"], 64, p["xsr0"], 64, 64, 64, B * npix, int(self.op16))

    # SCNetbk (:807-822).  BlockRCB applies ONE body to the three pyramid levels (:766-770), so every convolution and every
    # helper kernel of a block runs the three levels in one launch (fcvsr_conv2d_tc_multi / *_multi: one tile or thread list
    # spanning the levels) on a single stream: the small levels (1/4 and 1/16 of the pixels) fill SMs next to level 0 instead
    # of running as under-filled launches on side streams, and back-to-back convolutions keep their programmatic-dependent-
    # launch overlap.  Full-precision streams (xs, cur, t, r0, rr) stay fp32; every tensor-core convolution reads the
    # operand-typed copy written next to it (xsr, curr, tr, r0h, rrh: TF32-rounded fp32 or bf16).
    def _scnet(self, ws, p, B, H, W):
        P, G = self.packs, self.model.SCGroupN
        dims = [(H, W), (H // 2, W // 2), (H // 4, W // 4)]
        LK = C.ACT_LEAKY
        R, O16 = int(self.use_tc), int(self.op16)
        res16 = int(bool(O16) and self.res16)
        r016 = bool(O16) and self.r016
        rr16 = int(bool(O16) and self.rr16)
        t16 = 4 if (O16 and self.t16) else 0      # down / up conv outputs td, tu as bf16 tensors (flag bit of level_mix)
        x16 = bool(O16 and R and rr16 and t16 and self.x16)
        L3 = (0, 1, 2)
        vp = lambda *ptrs: (ctypes.c_void_p * len(ptrs))(*ptrs)  # noqa: E731
        ia = lambda *v: (ctypes.c_int * len(v))(*v)  # noqa: E731
        H3, W3 = ia(*[d[0] for d in dims]), ia(*[d[1] for d in dims])
        P3 = ia(*[d[0] * d[1] for d in dims])
        coef3 = (ctypes.c_float * 3)(2.0, 1.0, 2.0)              # level 0 has d = r, level 2 has u = r (:771-776)
        adds = [p["addall"] + l * B * 64 * 4 for l in L3]
        for g in range(G):
            inp = [p[f"xs{l}"] if g == 0 else p[f"cur{l}"] for l in L3]
            inp_r = [p[f"xsr{l}"] if g == 0 else p[f"curr{l}"] for l in L3] if R else inp
            for k in range(3):
                q = f"g{g}.b{k}."
                src = inp if k == 0 else [p[f"t{l}"] for l in L3]
                src_r = (inp_r if k == 0 else [p[f"tr{l}"] for l in L3]) if R else src
                r0_op = [p[f"r0h{l}"] if R else p[f"r0{l}"] for l in L3]
                rr_op = [p[f"rrh{l}"] if R else p[f"rr{l}"] for l in L3]
                a128 = [p[f"a128_{l}"] for l in L3]
                # BlockRCB body (:729-751) + RCB (:705-725)
                self._conv_multi(P[q + "c0"], src_r, 64, a128, 128, B, dims, act=LK, slope=0.1, rnd=True)
                if r016:                        # RCB input r0 only as the bf16 operand tensor (also the RCB skip)
                    self._conv_multi(P[q + "c2"], a128, 128, [p[f"r0h{l}"] for l in L3], 64, B, dims, rnd=True)
                else:
                    self._conv_multi(P[q + "c2"], a128, 128, [p[f"r0{l}"] for l in L3], 64, B, dims,
                                     y2=[p[f"r0h{l}"] for l in L3], ldy2=64)
                self._conv_multi(P[q + "r0"], r0_op, 64, [p[f"c1{l}"] for l in L3], 64, B, dims, act=LK, slope=0.2, rnd=True)
                # res (RCB body output, consumed by the ContextBlock and the RCB tail only) is a bf16 tensor in bf16 mode
                res = [p[f"res{l}"] for l in L3]
                self._conv_multi(P[q + "r2"], [p[f"c1{l}"] for l in L3], 64, res, 64, B, dims, rnd=bool(res16))
                self._k("fcvsr_context_block_multi", 3, vp(*res), 64, P[q + "mask"].data_ptr(), P[q + "a0"].data_ptr(),
                        P[q + "a2"].data_ptr(), p["ctxall"], p["addall"], p["ctxcnt"], B, P3, res16)
                self._k("fcvsr_rcb_finish_multi", 3, vp(*res), vp(*adds), vp(*[p[f"r0h{l}"] if r016 else p[f"r0{l}"] for l in L3]),
                        vp(*[0 if rr16 else p[f"rr{l}"] for l in L3]),
                        vp(*[p[f"rrh{l}"] if (R and (l > 0 or rr16)) else 0 for l in L3]),
                        vp(p["rrp0"], p["rrp1"], 0), H3, W3, B, O16, int(not R), res16 | (2 if r016 else 0))
                # down: 1x1 conv on the 2x2 mean == mean of the conv (:753-757); up: 1x1 conv, interpolated in level_mix (:759-763)
                if R and self.merge_down_up:
                    self._conv_down_up(P[q + "du_w"], P[q + "du_b"], [p["rrp0"], p["rrp1"], rr_op[1], rr_op[2]],
                                       [p["td0"], p["td1"], p["tu1"], p["tu2"]], B, [dims[1], dims[2], dims[1], dims[2]], bool(t16))
                else:
                    self._conv_multi(P[q + "down"], [p["rrp0"], p["rrp1"]], 64, [p["td0"], p["td1"]], 64, B, dims[1:], rnd=bool(t16))
                    self._conv_multi(P[q + "up"], rr_op[1:], 64, [p["tu1"], p["tu2"]], 64, B, dims[1:], rnd=bool(t16))
                # x + r + d + u (:771-776)
                # bf16 mode: the x carried through the group's three blocks lives only as the bf16 operand copy (what the block's
                # first convolution read); the fp32 copy of every block (116 MB read + 116 MB written at 6 windows) is gone
                xin = [p[f"tr{l}"] for l in L3] if (x16 and k > 0) else src
                self._k("fcvsr_level_mix_multi", 3, vp(*xin), 64, vp(*[0 if x16 else p[f"t{l}"] for l in L3]), 64,
                        vp(*[p[f"rrh{l}" if rr16 else f"rr{l}"] for l in L3]), coef3, vp(0, p["td0"], p["td1"]),
                        vp(p["tu1"], p["tu2"], 0), B, H3, W3, vp(*[p[f"tr{l}"] if R else 0 for l in L3]), 64, 0, O16,
                        1 + 2 * rr16 + t16 + (8 if (x16 and k > 0) else 0))
            # SCGroupbk tail: x + conv(res) (:797-803)
            self._conv_multi(P[f"g{g}.conv"], [p[f"tr{l}"] if R else p[f"t{l}"] for l in L3], 64, [p[f"cur{l}"] for l in L3], 64,
                             B, dims, res=inp, ldres=64, y2=[p[f"curr{l}"] for l in L3], ldy2=64)
        # SCNetbk skip (:816-822): outputs feed only convolutions -> operand-typed; level 0 lands in the concat buffer
        CL = self.cat_ld
        self._k("fcvsr_level_mix", p["xs0"], 64, p["fuse"], CL, p["cur0"], 1.0, 0, 0, B, *dims[0], 0, 0, R, O16, 0)
        self._k("fcvsr_level_mix_multi", 2, vp(p["xs1"], p["xs2"]), 64, vp(p["o2"], p["o3"]), 64, vp(p["cur1"], p["cur2"]),
                (ctypes.c_float * 2)(1.0, 1.0), vp(0, 0), vp(0, 0), B, ia(*[d[0] for d in dims[1:]]), ia(*[d[1] for d in dims[1:]]),
                vp(0, 0), 0, R, O16, 0)

    # pyramid fuse + up-sampler (:2739-2751)
    def _tail(self, x, out, ws, p, B, H, W):
        P = self.packs
        h2, w2, h3, w3 = H // 2, W // 2, H // 4, W // 4
        PR = C.ACT_PRELU
        sl = P["prelu"].data_ptr()
        R, O16 = int(self.use_tc), int(self.op16)
        E = self.E if R else 4
        CL = self.cat_ld
        cat2, fuse = p["cat2"], p["fuse"]
        # out_L3 -> PS -> cat2[64:80] (L2 res) -> PS -> fuse[80:84]
        self._conv(P["up_l3"], p["o3"], 64, cat2 + 64 * E, CL, B, h3, w3, act=PR, slope_ptr=sl, rnd=True)
        self._k("fcvsr_pixel_shuffle", cat2 + 64 * E, CL, fuse + 80 * E, CL, B, h2, w2, 4, O16)
        # out_L2 (kept in shuffle order): full precision in u2 (residual below), operand copy in cat2[0:64]
        if R:
            self._conv(P["up_l2"], p["o2"], 64, p["u2"], 64, B, h2, w2, act=PR, slope_ptr=sl, y2=cat2, ldy2=CL)
            u2, ldu2 = p["u2"], 64
        else:
            self._conv(P["up_l2"], p["o2"], 64, cat2, CL, B, h2, w2, act=PR, slope_ptr=sl)
            u2, ldu2 = cat2, CL
        # PS(out_L2 + upconv1_L2_2(cat)) -> fuse[64:80]
        self._conv(P["up_l2_2"], cat2, CL, fuse + 64 * E, CL, B, h2, w2, res=u2, ldres=ldu2, rnd=True)
        self._conv(P["fuse"], fuse, CL, p["f1"], 64, B, H, W, rnd=True)
        self._conv(P["rec0"], p["f1"], 64, p["f2"], 64, B, H, W, rnd=True)
        self._conv(P["up1"], p["f2"], 64, p["up1"], 64, B, H, W, act=PR, slope_ptr=sl, rnd=True)
        self._conv(P["up2"], p["up1"], 64, p["up2"], 64, B, 2 * H, 2 * W, act=PR, slope_ptr=sl, rnd=True)
        Cc = self.model.in_ch
        if Cc > 1:
            # mmedit variants (3-channel output, sr_backbones/fcvsr.py:133-136): conv_last0 into an NHWC scratch (pixel stride 4),
            # then one pass transposes to the module's NCHW output and adds the bilinear x4 of the centre frame
            self._conv(P["last"], p["up2"], 64, p["lastc"], 4, B, 4 * H, 4 * W)
            self._k("fcvsr_rgb_tail", p["lastc"], 4, x.data_ptr() + 3 * Cc * H * W * 4, 7 * Cc * H * W, out.data_ptr(), B, Cc, H, W)
            return
        # bilinear x4 of the centre LR frame (:2750) rides in conv_last0's epilogue as the residual
        self._k("fcvsr_bilinear_up4", x.data_ptr() + 3 * H * W * 4, 7 * H * W, p["base"], B, H, W)
        if self.op16 and self.use_last_kernel:
            self._k("fcvsr_conv3x3_c64_to1", p["up2"], 64, ctypes.addressof(self._last_w), self._last_b, p["base"],
                    out.data_ptr(), B, 4 * H, 4 * W)
        else:
            self._conv(P["last"], p["up2"], 64, out.data_ptr(), 1, B, 4 * H, 4 * W, res=p["base"], ldres=1)
