/* fcvsr_b200 -- C ABI of the B200-native FCVSR forward kernels (libfcvsr_b200.so).
 *
 * Plain pointers and sizes only: every pointer is a DEVICE pointer to fp32 data unless stated,
 * the caller owns all buffers (nothing here allocates), every call enqueues work on `stream` and
 * returns immediately with FCVSR_OK (0) or a negative error code.  This is the boundary the Python
 * host layer (fcvsr_b200/_capi.py, ctypes) binds, and the one a maintainer of the reference would
 * bind in place of the ATen / deform_conv_cuda calls cited per entry.
 *
 * Tensor layout: NHWC ("pixel-major") fp32.  Element (b,y,x,c) of a tensor with pixel stride `ld`
 * lives at base[((b*H + y)*W + x)*ld + c]; a channel slice is base+offset with the same ld.
 * Spectra are complex-interleaved NHWC: [B,H,Wf,C] float2, Wf = W/2+1.
 *
 * Process model: one process per GPU (the bench and the sequence driver launch one rank per device).  The library keeps a
 * few per-process statics (SM count, function attributes, a 4-byte device error word for the tensor-core kernels), so a
 * process must not drive more than one device through it.  All entries are asynchronous on `stream`, never allocate
 * (except that error word, once) and never synchronise.
 *
 * Reference files are relative to /root/reference (QZ1-boy/FCVSR).
 */
#ifndef FCVSR_B200_H
#define FCVSR_B200_H

#include <cuda_runtime_api.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FCVSR_OK 0
#define FCVSR_ERR_ARG (-1)          /* invalid argument / unsupported shape */
#define FCVSR_ERR_CUDA (-2)         /* a CUDA runtime call or launch failed */
#define FCVSR_ERR_UNSUPPORTED (-3)  /* shape outside the kernel's envelope (caller must use another entry) */

#define FCVSR_ACT_NONE 0
#define FCVSR_ACT_RELU 1
#define FCVSR_ACT_LEAKY 2 /* slope by value */
#define FCVSR_ACT_PRELU 3 /* slope read from device scalar slope_ptr (nn.PReLU weight) */

/* ---- convolutions --------------------------------------------------------------------------- */

/* Generic NHWC convolution on the CUDA cores: y = act(conv(x,w)+bias) + res - res2, optionally stored
 * through pixel_shuffle(2) (columns then ordered (i,j,c), y is [B,2Ho,2Wo,ldy]).  Padding k/2.
 * w packed [k*k][Cin][Cout].  x may be NCHW (x_nchw=1, ldx ignored), y may be NCHW (y_nchw=1, ldy ignored).
 * Replaces F.conv2d at CVSR_train/arch/CVSR_freq.py:2663 (feat_extract), :2671-2672 (stride 2),
 * :2684 (conv_last0), :1380-1395 (per-bin MLP heads) and is the fp32 cross-check of the tensor-core
 * kernel.  round_out=1 stores y rounded to TF32 (for tensors that only feed tensor-core convs); y2 != NULL
 * additionally stores a TF32-rounded copy (for tensors that are also full-precision residuals). */
int fcvsr_conv2d_direct(const float* x, int ldx, int x_nchw, const float* w, const float* bias,
                        const float* res, int ldres, const float* res2, int ldres2, float* y, int ldy,
                        int B, int H, int W, int Cin, int Cout, int ksize, int stride, int act, float slope,
                        const float* slope_ptr, int pixel_shuffle, int y_nchw, void* y2, int ldy2,
                        int round_out, int op16, cudaStream_t stream);

/* tcgen05/TMEM implicit-GEMM convolution fed by TMA (stride 1, k in {1,3}, Cin % 32 == 0, Cout % 16 == 0,
 * Cout <= 256), TF32 operands / fp32 accumulate, same fused epilogue as above.
 * w packed [Cout][k*k][Cin] (K-major).  x, y, w, res must be 16-byte aligned, ld multiples of 4.
 * Replaces every 3x3 / 1x1 nn.Conv2d of MGAAbk (CVSR_freq.py:1371-1430), SCNetbk (:705-822) and the
 * up-sampling tail (:2739-2749).  Returns FCVSR_ERR_UNSUPPORTED for shapes outside the envelope.
 * max_ctas > 0 caps the persistent grid (pyramid levels run concurrently on separate streams).
 * op16 = 1 selects bf16 operands (kind::f16, K = 16): x and w are bf16 (Cin % 64 == 0), a `round_out` y and y2 are
 * bf16 tensors (ld in elements); op16 = 0: TF32 operands, `round_out` / y2 store TF32-rounded fp32.
 * round_out = 2 stores y as fp16 (ld in elements) in either mode (the per-pixel filter taps). */
int fcvsr_conv2d_tc(const float* x, int ldx, const float* w, const float* bias, const float* res, int ldres,
                    const float* res2, int ldres2, float* y, int ldy, int B, int H, int W, int Cin, int Cout,
                    int ksize, int act, float slope, const float* slope_ptr, int pixel_shuffle,
                    float* y2, int ldy2, int round_out, int max_ctas, int op16, cudaStream_t stream);

/* The same convolution (weights, bias, activation) on up to three NHWC tensors of different spatial size in ONE launch: the
 * pyramid levels of SCNetbk (CVSR_freq.py:766-770, BlockRCB applies one body to every level).  The persistent CTAs walk one tile
 * list that spans the levels.  x / res / y / y2: HOST arrays of nprob device pointers (res, y2 may be NULL or hold NULL
 * entries); H, W: HOST arrays of nprob ints; everything else as fcvsr_conv2d_tc (no pixel shuffle, no res2). */
int fcvsr_conv2d_tc_multi(int nprob, const void* const* x, int ldx, const float* w, const float* bias, const float* const* res,
                          int ldres, float* const* y, int ldy, const int* H, const int* W, int B, int Cin, int Cout, int ksize,
                          int act, float slope, const float* slope_ptr, float* const* y2, int ldy2, int round_out, int op16,
                          cudaStream_t stream);

/* ---- FFT (torch.fft.rfft2 / irfft2 / fftn / ifftn of CVSR_freq.py:1452-1454, :1499-1504, :2082-2088) */

/* tw: float2[N] = exp(-2 pi i k / N) for the transform length N (W for *_w, H for *_h). */
int fcvsr_fft_r2c_w(const float* x, int ldx, float* out_c, const float* tw, int B, int H, int W, int C,
                    cudaStream_t stream);
/* in/out complex [B,H,Wf,C] (C complex channels); optional real mask [H*Wf] multiplied at load;
 * in == out allowed.  inverse: 0 forward, 1 inverse (unnormalised); result * scale.  nrep > 1 transforms the
 * same input nrep times with mask + r*H*Wf into out + r*B*H*Wf*C (all band masks of Split_freq in one launch).
 * out_bf16 (optional, nrep == 1): bf16 copy of the result, same layout (the operand tensor of the per-bin 1x1 convs). */
int fcvsr_fft_c2c_h(const float* in_c, float* out_c, const float* tw, const float* mask, int B, int H, int Wf,
                    int C, int inverse, float scale, int round_out, int nrep, void* out_bf16, cudaStream_t stream);
/* complex [B,H,Wf,C] -> real [B,H,W,ldy] (C real channels), torch c2r semantics, result * scale. */
int fcvsr_fft_c2r_w(const float* in_c, float* y, int ldy, const float* tw, int B, int H, int W, int C,
                    float scale, cudaStream_t stream);

/* ---- MGAAbk (CVSR_freq.py:1365-1547) ----------------------------------------------------------- */

/* CorrBlock lookup (:1279-1337): S [B,H*Wf,ldS] floats with the two packed spectra at float offsets
 * a_off / b_off (C2 floats each, complex-interleaved); out [B,H*Wf,ldo], 81 channels. */
int fcvsr_corr_gather(const float* S, int ldS, int a_off, int b_off, void* out, int ldo, int B, int H, int Wf,
                      int C2, int op_mode /* 0 fp32, 1 TF32-rounded, 2 bf16; + 4: 16-byte aligned rows whose channels 81 .. 83 (fp32) /
                      81 .. 87 (bf16) may be zero-filled: vector stores */, cudaStream_t stream);
/* same, writing the result to two tensors of the same layout (out2 may be NULL): corr_f feeds both offset branches (:1487-1488) */
int fcvsr_corr_gather2(const float* S, int ldS, int a_off, int b_off, void* out, void* out2, int ldo, int B, int H, int Wf,
                       int C2, int op_mode, cudaStream_t stream);

/* ConvBlk(4, index=i) for i < A and both directions (:344-357, :1494-1498):
 * off [2][B][H*Wf][4] (dir-major), w1/w2 packed per iteration [k*k][ci][co] back to back, prelu [A],
 * ca_w [A][2][4][4], sim [B,H*Wf,ldsim] (4 ch).  Scratch t1,t2 [A][2][B][H*Wf][4], partial
 * [A][2B][max(ceil(H*Wf/128), ceil(Wf/64)*ceil(H/8))][4].  z: complex [B,H*Wf,4A], channel (i*2+dir)*2+m = (v[m], v[2+m]). */
int fcvsr_offset_blocks(const float* off, const float* w1, const float* w2, const float* prelu,
                        const float* ca_w, const float* sim, int ldsim, float* t1, float* t2, float* partial,
                        float* z, int B, int H, int Wf, int A, cudaStream_t stream);

/* One IAC iteration (:1230-1250 = flow_warp :1188-1227 + SAC :1253-1276 + residual + LeakyReLU 0.1) for
 * the forward (f) and backward (b) neighbour, 64 channels.  offs [B,H,W,ldoffs]: (dx,dy) at channel
 * ch_f / ch_b.  taps [B,H,W,ldtaps]: 192 channels [t][c] of this iteration (fp16 when taps_half = 1).
 * round_out: 0 fp32 outputs, 1 TF32-rounded fp32, 2 bf16 (next_* are then bf16 tensors); + 4: prev_* are bf16 tensors
 * (ld in elements) -- the ping-pong between iterations of the bf16 mode. */
int fcvsr_iac_step(const float* prev_f, int ldprev_f, const float* prev_b, int ldprev_b, const float* xin_f,
                   int ldxin_f, const float* xin_b, int ldxin_b, float* next_f, int ldnext_f, float* next_b,
                   int ldnext_b, const float* offs, int ldoffs, int ch_f, int ch_b, const float* taps, int ldtaps,
                   int taps_half, int B, int H, int W, int round_out, cudaStream_t stream);

/* The same IAC iteration with the taps computed on chip (bf16 mode): the iteration's 64 -> 192 slice of the kernel
 * predictor's last 1x1 convolution (MGAA.F.1, :1522-1523; only rows i*384 + c*3 + t are live, :1231-1235) is one
 * 128 x 192 x 64 tcgen05 GEMM per 8 x 14 tile, its TMEM accumulator is the tap set, and `Pred_K` never reaches memory.
 * kp [B,H,W,ldkp] bf16: the F.0 output (64 ch).  w: bf16 [192][64], row c4*12 + t*4 + cc = F.1 row i*384 + (4 c4 + cc)*3 + t;
 * bias [192] fp32 in the same order.  prev_* fp32 (prev16 = 0) or bf16 (prev16 = 1), ld in elements; next_* bf16.
 * prev16 = 2 (+ 4): the fp32-contract mode -- kp and w are TF32-rounded fp32 tensors (kind::tf32 MMAs), prev_* and next_* fp32;
 * with + 4 the outputs are TF32-rounded (the last iteration feeds conv3). */
int fcvsr_iac_step_tc(const void* prev_f, int ldprev_f, const void* prev_b, int ldprev_b, int prev16,
                      const float* xin_f, int ldxin_f, const float* xin_b, int ldxin_b, void* next_f, int ldnext_f,
                      void* next_b, int ldnext_b, const float* offs, int ldoffs, int ch_f, int ch_b, const void* kp,
                      int ldkp, const void* w, const float* bias, int B, int H, int W, cudaStream_t stream);

/* y[pix,0:Cy] = operand-typed copy (TF32-rounded fp32 or bf16) of x[pix,0:C], channels C..Cy-1 zero: the
 * tensor-core operand copy of a tensor that is also a full-precision residual */
int fcvsr_round_copy(const float* x, int ldx, void* y, int ldy, int C, int Cy, long long npix, int op16,
                     cudaStream_t stream);

/* ---- MultiFreq_Refinment (CVSR_freq.py:2104-2133, :2183-2254) ---------------------------------- */

int fcvsr_chansum64(const float* x, int ldx, float* partial, int B, int P, cudaStream_t stream);
/* partial [B][nblk][128] -> out [B][128]: mode 0 mean, mode 1 CALayer gate sigmoid(W2 relu(W1 mean)) */
int fcvsr_reduce_finalize(const float* partial, int nblk, int nvec, float inv_count, int mode, const float* w1,
                          const float* w2, float* out, int B, cudaStream_t stream);
/* apply DivEnh step i-1 (x_prev != NULL) and reduce step i (mode 1) or sum So (mode 2); see mffr.cu */
int fcvsr_divenh_step(const float* x_prev, const float* a_prev, const float* b_prev, const float* mean_prev,
                      const float* gate_prev, int prev_is_first, int mode, int cur_is_first, const float* x_cur,
                      const float* a_cur, const float* b_cur, const float* mean_cur, float* sb, float* so,
                      float* partial, int B, int P, cudaStream_t stream);
int fcvsr_mffr_final(const float* so, const float* gate, const float* x, int ldx, float* y, int ldy, int B, int P,
                     cudaStream_t stream);

/* fcvsr_conv2d_tc_multi with a different filter per problem (csrc/conv_tc.cu): w / bias hold w_rows >= Cout rows (several
 * [Cout][k*k*Cin] filters stacked), problem i uses rows wrow[i] .. wrow[i] + Cout (HOST array, multiples of 16); up to four
 * problems.  The 1x1 `down` and `up` convolutions of a BlockRCB (CVSR_freq.py:753-763) run as one launch this way. */
int fcvsr_conv2d_tc_multi_w(int nprob, const void* const* x, int ldx, const float* w, const float* bias, const int* wrow,
                            int w_rows, float* const* y, int ldy, const int* H, const int* W, int B, int Cin, int Cout,
                            int ksize, int act, float slope, int round_out, int op16, cudaStream_t stream);

/* ---- SCNetbk helpers (CVSR_freq.py:657-777) ----------------------------------------------------- */

/* ContextBlock (:657-701): add[b][64] = W2 lrelu_0.2(W1 softmax-pool(x)); partial [B][ceil(P/128)][66] (scratch).
 * x_bf16 = 1: x is a bf16 tensor (ld in elements, a multiple of 8; x 16-byte aligned) -- the bf16 mode stores the RCB's second
 * conv output that way.  counters: B ints of scratch that must be ZERO before the first call; the kernel leaves them zero
 * (the last block of an image to finish merges the partials and resets its counter, so the launch replays from a CUDA graph).
 * Two launches that share `counters` / `partial` must not run concurrently. */
int fcvsr_context_block(const void* x, int ldx, const float* wmask, const float* w1, const float* w2,
                        float* partial, float* add, int* counters, int B, int P, int x_bf16, cudaStream_t stream);
/* RCB tail (:720-724): r = lrelu_0.2(res + add[b]) + r0 (64 ch, ld 64).  r_pool (optional, needs even H, W with
 * H*W == P): 2x2 mean of r, [B,H/2,W/2,64], operand-typed or plain fp32 (pool_plain) -- the input of the 1x1 `down`
 * convolution, which commutes with the reference's Interpolate(0.5) (:753-757).  res_bf16: bit 0 = res is a bf16 tensor, bit 1 = r0 is. */
int fcvsr_rcb_finish(const void* res, const float* add, const void* r0, float* r, int B, int P,
                     void* r_operand_copy, int op16, void* r_pool, int H, int W, int pool_plain, int res_bf16,
                     cudaStream_t stream);
/* BlockRCB cross-level sum (:766-777): xout = xin + coef*r + d + bilinear_x2(tu[B,H/2,W/2,64]) with
 * d = mean2x2(td[B,2H,2W,64]), or d = td[B,H,W,64] when td_pooled & 1 (down conv applied to the pooled tensor);
 * td_pooled & 2: r is a bf16 tensor; td_pooled & 4: td and tu are; td_pooled & 8 (only with 2 and 4): xin is a bf16 tensor too
 * (pass its address through the float pointer; ldx in elements).  xout may be NULL when xout_r is given: only the operand copy
 * is written (xin == xout_r in place is fine, a thread reads and writes its own four channels).
 * fcvsr_rcb_finish: r may be NULL when r_operand_copy is given. */
int fcvsr_level_mix(const float* xin, int ldx, float* xout, int ldo, const void* r, float coef, const void* td,
                    const void* tu, int B, int H, int W, void* xout_r, int ldr, int round_main, int op16,
                    int td_pooled, cudaStream_t stream);

/* The three SCNet helpers over up to three pyramid levels in one launch (same weights / flags for every level; pointer
 * arguments are HOST arrays of nlev device pointers, H / W / P / coef HOST arrays).  partial: sum over levels of
 * B*ceil(P_l/128)*66 floats; add: [nlev][B][64]; counters: nlev*B zeroed ints (see fcvsr_context_block).  NULL entries in
 * r / r_op / r_pool / td / tu / xout_r skip that output or term. */
int fcvsr_context_block_multi(int nlev, const void* const* x, int ldx, const float* wmask, const float* w1, const float* w2,
                              float* partial, float* add, int* counters, int B, const int* P, int x_bf16, cudaStream_t stream);
/* Training step: the ContextBlock's soft-max pooling alone and its backward (csrc/scnet.cu); the 64 -> 64 -> 64 MLP stays with
 * autograd.  pool: [nlev][B][66] = context[64], soft-max running max (log2 domain) and sum; partial / counters as above.
 * Backward: gctx = d loss / d context [nlev][B][64]; dx[l] ([B,P_l,ldx] fp32) is WRITTEN with d loss / d x; dwpart receives
 * B * fcvsr_context_pool_backward_blocks(nlev, P) rows of 64 partial sums of d loss / d wmask (the caller adds them up). */
int fcvsr_context_pool_multi(int nlev, const void* const* x, int ldx, const float* wmask, float* partial, float* pool,
                             int* counters, int B, const int* P, int x_bf16, cudaStream_t stream);
int fcvsr_context_pool_backward_blocks(int nlev, const int* P);
int fcvsr_context_pool_backward_multi(int nlev, const void* const* x, int ldx, const float* wmask, const float* pool,
                                      const float* gctx, float* const* dx, float* dwpart, int B, const int* P, cudaStream_t stream);
/* Backward of the RCB tail r = lrelu_0.2(res + add[b]) + r0 (training step): gres[l] = g[l] * (res + add >= 0 ? 1 : 0.2) is WRITTEN,
 * gadd[l] ([B,64]) ACCUMULATES the sum of gres over the pixels (zero it first); the gradient of r0 is g itself.  HOST arrays of
 * nlev device pointers, 64-channel fp32 tensors [B,P_l,64]. */
int fcvsr_rcb_finish_backward_multi(int nlev, const float* const* res, const float* const* add, const float* const* g,
                                    float* const* gres, float* const* gadd, const int* P, int B, cudaStream_t stream);
int fcvsr_rcb_finish_multi(int nlev, const void* const* res, const float* const* add, const void* const* r0, float* const* r,
                           void* const* r_op, void* const* r_pool, const int* H, const int* W, int B, int op16, int pool_plain,
                           int res_bf16, cudaStream_t stream);
int fcvsr_level_mix_multi(int nlev, const float* const* xin, int ldx, float* const* xout, int ldo, const void* const* r,
                          const float* coef, const void* const* td, const void* const* tu, int B, const int* H, const int* W,
                          void* const* xout_r, int ldr, int round_main, int op16, int td_pooled, cudaStream_t stream);

/* ---- tail (CVSR_freq.py:2739-2751) -------------------------------------------------------------- */
int fcvsr_pixel_shuffle(const void* in, int ldi, void* out, int ldo, int B, int H, int W, int Co, int half,
                        cudaStream_t stream);
int fcvsr_bilinear_up4(const float* in, long long bstride, float* out, int B, int H, int W, cudaStream_t stream);
/* conv_last0 + skip (CVSR_freq.py:2749-2751) in bf16 mode: y[b,y,x] = bias + res[b,y,x] + sum_{ky,kx,c} w[ky][kx][c] *
 * x[b, y+ky-1, x+kx-1, c] for a bf16 NHWC input with 64 channels (zero padding).  w_host is a HOST pointer to 9*64 floats:
 * the weights travel in the kernel parameter space.  CUDA cores: the tensor-core path needs an N = 16 MMA for one column. */
int fcvsr_conv3x3_c64_to1(const void* x_bf16, int ldx, const float* w_host, float bias, const float* res, float* y, int B,
                          int H, int W, cudaStream_t stream);
/* y[b,i,j,0:C] = x[b,2i,2j,0:C] (fp32 NHWC, H and W even, C % 4 == 0) plus an optional operand-typed copy y2: turns the
 * stride-1 tensor-core convolution into the stride-2 rconcat1/2 of the pyramid (CVSR_freq.py:2671-2672, :2735-2736). */
int fcvsr_subsample2(const float* x, int ldx, float* y, int ldy, void* y2, int ldy2, int B, int H, int W, int C, int op16,
                     cudaStream_t stream);
/* NCHW clip [B,T,H,W] -> NHWC [B,H,W,32] (channels >= T zero, TF32-rounded): tensor-core operand of feat_extract */
int fcvsr_pack_clip(const float* x, void* y, int B, int T, int H, int W, int op16, cudaStream_t stream);
int fcvsr_fill_channels(float* x, int ld, int c0, int nc, float v, long long npix, cudaStream_t stream);
/* Output stage of the 3-channel mmedit variants (FCVSRNet / FCVSR_SNet, sr_backbones/fcvsr.py:133-136): out [B,C,4H,4W] NCHW =
 * t [B,4H,4W,ldt] (NHWC conv_last0 result) + bilinear x4 (align_corners=False) of the centre LR frame center [B,C,H,W] with batch
 * stride bstride (elements). */
int fcvsr_rgb_tail(const float* t, int ldt, const float* center, long long bstride, float* out, int B, int C, int H, int W,
                   cudaStream_t stream);
/* 8-bit output of the evaluation driver (CVSR_train/test_LD_freqCVSR.py:85-93: crop the padded rows, clamp to [0,1], * 255,
 * numpy astype(uint8) = truncation): out [B,Ho,Wo] uint8 = trunc(clamp(v[b, y < Ho, x < Wo], 0, 1) * 255) of v [B,H,W] fp32. */
int fcvsr_quantize_u8(const float* v, unsigned char* out, int B, int H, int W, int Ho, int Wo, cudaStream_t stream);

/* ---- training loss (CVSR_train/opt/loss.py:20-31)       ---------------------------------------- */

/* CharbonnierLoss(x, y, mean_res): out[0] = sum sqrt((x - y)^2 + eps) over `numel` fp32 elements, or with mean_res the
 * per-sample mean of x - y first (:27-29; `batch` samples of numel / batch elements).  Deterministic two-stage reduction
 * with double-precision partials; scratch: max(592, batch) doubles; out: one float on the device. */
int fcvsr_charbonnier_loss(const float* x, const float* y, long long numel, int batch, int mean_res, float eps,
                           double* scratch, float* out, cudaStream_t stream);

/* Backward of CharbonnierLoss: grad_x[i] = grad_out[0] * d / sqrt(d^2 + eps), d = x[i] - y[i]; grad_y = -grad_x (either may be
 * NULL).  grad_out: the upstream gradient of the scalar loss on the DEVICE.  With mean_res, `scratch` must be the buffer the
 * forward call filled (it holds the per-sample means).  What autograd derives for opt/loss.py:20-31 in the reference. */
int fcvsr_charbonnier_loss_backward(const float* x, const float* y, long long numel, int batch, int mean_res, float eps,
                                    const float* grad_out, const double* scratch, float* grad_x, float* grad_y,
                                    cudaStream_t stream);

/* mmedit pixel losses (mmedit_train/mmedit/models/losses/pixelwise_loss.py:13-51,:54-190; the FCVSR REDS configuration uses
 * MSELoss(mean), configs/restorers/fcvsr/fcvsr_redsLD_QP22.py:7): out[0] = scale * sum_i f(x_i - y_i) with f = sqrt(d^2 + eps)
 * (kind 0, CharbonnierLoss, eps 1e-12), d^2 (kind 1, MSELoss) or |d| (kind 2, L1Loss); reduction 'mean' and loss_weight are the
 * caller's `scale` (loss_weight / numel).  Deterministic; scratch: 592 doubles.  The backward reads the upstream gradient of the
 * scalar from the device: grad_x = grad_out[0] * scale * f'(d), grad_y = -grad_x (either may be NULL). */
int fcvsr_pixel_loss(const float* x, const float* y, long long numel, int kind, float eps, double scale, double* scratch, float* out,
                     cudaStream_t stream);
int fcvsr_pixel_loss_backward(const float* x, const float* y, long long numel, int kind, float eps, double scale,
                              const float* grad_out, float* grad_x, float* grad_y, cudaStream_t stream);

/* Adam step of the reference's training loop (train_LD_freqCVSR_22.py:204,251: torch.optim.Adam with L2 weight decay, no
 * amsgrad), multi-tensor: params / grads / exp_avg / exp_avg_sq are HOST arrays of `count` device pointers (fp32), numels
 * their element counts, step the 1-based step number; hyper-parameters as doubles (1 - beta and the bias corrections are formed
 * in double on the host, as torch does).  One launch per 64 tensors. */
int fcvsr_adam_step(float* const* params, const float* const* grads, float* const* exp_avg, float* const* exp_avg_sq,
                    const long long* numels, int count, double lr, double beta1, double beta2, double eps,
                    double weight_decay, int step, cudaStream_t stream);

/* ---- adjoints of the forward kernels: the training step (CVSR_train/train_LD_freqCVSR_22.py:243-251, loss.backward()) -------
 * The reference derives these with ATen's autograd (cuDNN dgrad / wgrad, grid_sampler_2d_backward, elementwise kernels); the
 * entries below are what fcvsr_b200/autograd.py binds in their place. */

/* Data gradient of y = conv(x, w) (k x k, stride s, padding k/2), any shape, CUDA cores: dy [B,Ho,Wo,lddy] (Cout channels),
 * wt packed [k*k][Cout][Cin] (w.permute(2,3,0,1)), dx [B,H,W,lddx] (Cin channels) is WRITTEN.  Stride-1 layers whose shapes fit
 * use fcvsr_conv2d_tc on flipped / transposed weights instead. */
int fcvsr_conv2d_dgrad_direct(const float* dy, int lddy, const float* wt, float* dx, int lddx, int B, int H, int W, int Cin,
                              int Cout, int ksize, int stride, cudaStream_t stream);
/* Weight gradient: dw [k*k][Cin][Cout] += sum over output pixels of x[pix*s + tap - pad][ci] * dy[pix][co].  ACCUMULATES with
 * fp32 atomics over pixel slices (zero-fill first; run-to-run differences at rounding level, as cuDNN's atomic wgrad). */
int fcvsr_conv2d_wgrad(const float* x, int ldx, const float* dy, int lddy, float* dw, int B, int H, int W, int Cin, int Cout,
                       int ksize, int stride, cudaStream_t stream);
/* Both operand-typed copies of a dense fp32 tensor in one pass: y_tf32 = values rounded to nearest TF32 (fp32 storage), y_bf16 =
 * bf16; either may be NULL; numel % 4 == 0.  (Forward / data-gradient operands and the tcgen05 weight gradient's operands.) */
int fcvsr_round_copy_dual(const float* x, float* y_tf32, void* y_bf16, long long numel, cudaStream_t stream);
/* fcvsr_round_copy_dual for up to three tensors in one launch (HOST arrays of n device pointers / element counts; entries of
 * y_tf32 / y_bf16, or the arrays themselves, may be NULL). */
int fcvsr_round_copy_dual_multi(int n, const float* const* x, float* const* y_tf32, void* const* y_bf16, const long long* numel,
                                cudaStream_t stream);
/* fcvsr_conv2d_wgrad_tc over up to three tensor pairs of different spatial size (same batch and channels: the pyramid levels of a
 * BlockRCB convolution) in one launch; dw accumulates over all of them.  x, dy, H, W: HOST arrays. */
int fcvsr_conv2d_wgrad_tc_multi(int nprob, const void* const* x, int ldx, const void* const* dy, int lddy, float* dw, int B,
                                const int* H, const int* W, int Cin, int Cout, int ksize, cudaStream_t stream);
/* 4 -> 4 channel convolutions (the ConvBlk convolutions of CVSR_freq.py:344-357 under autograd; csrc/mgaa.cu): x, y, dy
 * [B,H,W,4] fp32 (16-byte aligned), w [k*k][ci][co], k odd <= 11, stride 1, zero padding k / 2, no bias.  The data gradient is
 * fcvsr_conv4x4 on dy with w'[tap][co][ci] = w[k*k - 1 - tap][ci][co]; fcvsr_conv4x4_wgrad ACCUMULATES dw [k*k][4][4] with
 * fp32 atomics (zero it first).  They replace the generic 64 x 64-tiled kernels above for this shape (1/16 .. 1/256 useful work). */
int fcvsr_conv4x4(const float* x, const float* w, float* y, int B, int H, int W, int ksize, cudaStream_t stream);
int fcvsr_conv4x4_wgrad(const float* x, const float* dy, float* dw, int B, int H, int W, int ksize, cudaStream_t stream);
/* The same weight gradient on the tcgen05 tensor cores (csrc/wgrad_tc.cu): x and dy are BF16 NHWC tensors (ld in elements,
 * % 8 == 0, 16-byte aligned), fp32 accumulation in TMEM over a split of the pixel tiles, fp32 vector reductions into dw.
 * k in {1, 3}, stride 1, Cin % 64 == 0, Cout % 64 == 0; other shapes return FCVSR_ERR_UNSUPPORTED (use fcvsr_conv2d_wgrad). */
int fcvsr_conv2d_wgrad_tc(const void* x_bf16, int ldx, const void* dy_bf16, int lddy, float* dw, int B, int H, int W, int Cin,
                          int Cout, int ksize, cudaStream_t stream);
/* Weight packing for fcvsr_conv2d_tc in one launch: w [Cout][Cin][k][k] (the reference's nn.Conv2d layout) -> K-major,
 * TF32-rounded out [rows][k*k*C]: transposed = 0: out[co][tap][ci] (forward, C = Cin); transposed = 1: out[ci][tap][co] with
 * flipped taps (the data gradient of a stride-1 convolution is the convolution with these weights, C = Cout).  rows =
 * max(real rows, rows_pad), extra rows zero (thin heads are padded to 16). */
int fcvsr_pack_conv_weight(const float* w, float* out, int Cout, int Cin, int ksize, int transposed, int rows_pad,
                           cudaStream_t stream);
/* out[c] (+)= sum over npix rows of x[row*ldx + c] (bias gradient), deterministic; scratch: ceil(npix / 64) * C floats. */
int fcvsr_colsum(const float* x, int ldx, int C, long long npix, float* scratch, float* out, int accumulate, cudaStream_t stream);
/* flow_warp (CVSR_freq.py:1188-1227) on NHWC maps: y[b,py,px,:] = bilinear(x[b], px + off[b,py,px,0], py + off[b,py,px,1]), zero
 * outside, align_corners=True; C % 4 == 0.  (The inference path fuses this into fcvsr_iac_step.) */
int fcvsr_flow_warp(const float* x, int ldx, const float* off, int ldoff, float* y, int ldy, int B, int H, int W, int C,
                    cudaStream_t stream);
/* its adjoint: dx [B,H,W,lddx] is ACCUMULATED with atomics (zero-fill first; may be NULL), doff [B,H,W,2] is written (may be NULL) */
int fcvsr_flow_warp_backward(const float* x, int ldx, const float* off, int ldoff, const float* dy, int lddy, float* dx, int lddx,
                             float* doff, int B, int H, int W, int C, cudaStream_t stream);
/* SAC (CVSR_freq.py:1253-1276): vertical then horizontal per-pixel, per-channel 3-tap filter, replicate padding, the SAME taps
 * in both passes (the reference's kernel1); taps [B,H,W,ldk] with channel t*C + c. */
int fcvsr_sac(const float* wp, int ldw, const float* taps, int ldk, float* out, int ldo, int B, int H, int W, int C,
              cudaStream_t stream);
/* its adjoint for upstream gradient g: dtaps [B,H,W,lddk] and dwp [B,H,W,lddw] are written (either may be NULL);
 * scratch: B*H*W*C floats. */
int fcvsr_sac_backward(const float* wp, int ldw, const float* taps, int ldk, const float* g, int ldg, float* scratch, float* dtaps,
                       int lddk, float* dwp, int lddw, int B, int H, int W, int C, cudaStream_t stream);
/* adjoint of fcvsr_corr_gather (fp32): dS [B,H*Wf,lddS] is written at the float ranges [a_off, a_off+C2) and [b_off, b_off+C2) */
int fcvsr_corr_gather_backward(const float* S, int ldS, int a_off, int b_off, const float* dout, int ldo, float* dS, int lddS,
                               int B, int H, int Wf, int C2, cudaStream_t stream);

/* ---- evaluation metrics next to the output (CVSR_train/metric/psnr_ssim.py:278-399, as called at :447-478) ------------------
 * a, b: [B,H,W] uint8 single-channel frames; out [B][2] floats = (PSNR dB, SSIM) over the frame without its `crop` border
 * pixels (SSIM: 11 x 11 Gaussian window, sigma 1.5, float64, valid region); scratch: B*ceil((H-2crop)/16)*ceil((W-2crop)/16)*2
 * doubles.  Deterministic. */
int fcvsr_psnr_ssim_u8(const unsigned char* a, const unsigned char* b, int B, int H, int W, int crop, double* scratch, float* out,
                       cudaStream_t stream);

/* ---- deformable convolution operator (CVSR_train/ops/dcn) --------------------------------------- */

/* Fused bilinear-gather + GEMM modulated deformable convolution forward, NCHW fp32 exactly as the
 * reference extension: replaces modulated_deform_conv_cuda_forward (ops/dcn/src/deform_conv_cuda.cpp:486-564;
 * call site ops/dcn/deform_conv.py:144-148) and, with mask == NULL, deform_conv_forward_cuda (:151-258;
 * call site deform_conv.py:52-57).  No column buffer: `columns`/`ones` scratch of the reference ABI are
 * not needed.  offset [B, dg*2*kh*kw, Ho, Wo] (dh, dw interleaved per tap), mask [B, dg*kh*kw, Ho, Wo].
 * offset_batch_stride / mask_batch_stride (elements, 0 = dense) let offset and mask be channel slices of
 * one conv_offset_mask output (ModulatedDeformConvPack, deform_conv.py:330-337); mask_sigmoid=1 applies
 * the sigmoid of :334 on the fly. */
int fcvsr_modulated_deform_conv_forward(const float* input, const float* weight, const float* bias,
                                        const float* offset, const float* mask, float* output, int B, int Cin,
                                        int H, int W, int Cout, int kh, int kw, int stride_h, int stride_w,
                                        int pad_h, int pad_w, int dil_h, int dil_w, int groups,
                                        int deformable_groups, long long offset_batch_stride,
                                        long long mask_batch_stride, int mask_sigmoid, cudaStream_t stream);

/* Same operator on the tensor cores (csrc/dcn_tc.cu): bilinear gather straight into the tcgen05 operand layout,
 * TF32-rounded operands, fp32 accumulate in TMEM, NCHW fp32 in and out.  groups == 1, Cin % 32 == 0, Cout % 16 == 0,
 * (Cin / deformable_groups) % 4 == 0, Cout <= 256; other shapes return FCVSR_ERR_UNSUPPORTED (use the exact-fp32 entry
 * above).  scratch_nhwc: caller-provided B*Cin*H*W floats (the kernel gathers from an NHWC copy of the input it makes
 * there; input == scratch_nhwc means the caller already made that copy with fcvsr_nchw_to_nhwc).  offset_pixel_stride > 0:
 * offset and mask are channel slices of an NHWC tensor with that pixel stride (the tcgen05 conv_offset_mask output of
 * ModulatedDeformConvPack, deform_conv.py:330-337) instead of NCHW planes.  Same reference call sites. */
int fcvsr_modulated_deform_conv_forward_tc(const float* input, const float* weight, const float* bias,
                                        const float* offset, const float* mask, float* output, int B, int Cin,
                                        int H, int W, int Cout, int kh, int kw, int stride_h, int stride_w,
                                        int pad_h, int pad_w, int dil_h, int dil_w, int groups,
                                        int deformable_groups, long long offset_batch_stride,
                                        long long mask_batch_stride, int mask_sigmoid, float* scratch_nhwc,
                                        int offset_pixel_stride, cudaStream_t stream);
/* NCHW fp32 [B,C,H,W] -> NHWC fp32 [B,H,W,C], optionally TF32-rounded (the operand layout of the tcgen05 kernels): the DCN
 * modules keep the reference's NCHW tensor contract (ops/dcn/deform_conv.py), the kernels gather from pixel-major copies. */
int fcvsr_nchw_to_nhwc(const float* x, float* y, int B, int C, int H, int W, int round_tf32, cudaStream_t stream);

/* Backward of the operator above (csrc/dcn_bwd.cu), NCHW fp32: replaces modulated_deform_conv_cuda_backward
 * (ops/dcn/src/deform_conv_cuda.cpp:566-700; call site ops/dcn/deform_conv.py:161-166) and, with mask == NULL,
 * deform_conv_backward_input_cuda + deform_conv_backward_parameters_cuda (:260-484; call sites deform_conv.py:76-82,:86-92,
 * scale = 1).  No column buffer.  Every grad_* pointer may be NULL (that gradient is skipped); the ones given ACCUMULATE,
 * so the caller zero-fills them first (as deform_conv.py:155-159 does with zeros_like).  grad_input [B,Cin,H,W],
 * grad_weight [Cout,Cin/groups,kh,kw], grad_bias [Cout], grad_offset / grad_mask dense, shaped like offset / mask.
 * grad_input, grad_offset, grad_mask and grad_weight are combined with fp32 atomics (as the reference's col2im is):
 * run-to-run differences are at rounding level.  offset == NULL (with grad_offset == NULL) means zero offsets: the backward of
 * a plain convolution, used for the conv_offset / conv_offset_mask layers of the *Pack modules (deform_conv.py:243-250,
 * :315-323).  scratch: NULL, or 2*B*Cin*H*W floats (16-byte aligned) that enable the
 * NHWC fast path for (Cin/groups) % 4 == 0 and (Cin/deformable_groups) % 4 == 0: 16-byte corner loads and vector
 * reductions (red.global.add.v4.f32) on an NHWC copy of input / grad_input. */
int fcvsr_modulated_deform_conv_backward(const float* input, const float* weight, const float* offset, const float* mask,
                                         const float* grad_output, float* grad_input, float* grad_weight, float* grad_bias,
                                         float* grad_offset, float* grad_mask, int B, int Cin, int H, int W, int Cout,
                                         int kh, int kw, int stride_h, int stride_w, int pad_h, int pad_w, int dil_h,
                                         int dil_w, int groups, int deformable_groups, float* scratch, cudaStream_t stream);

/* library / build info: returns a static string "fcvsr_b200 <version> sm_100a" */
const char* fcvsr_version(void);

#ifdef __cplusplus
}
#endif
#endif /* FCVSR_B200_H */
