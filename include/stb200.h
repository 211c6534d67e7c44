/*
 * stb200.h -- C ABI of libstb200.so: hand-written sm_100a CUDA kernels for the cost-volume hot
 * path of xxxupeng/stereo_toolbox.
 *
 * The reference has no FFI / plugin boundary for this path: it is reached through ordinary
 * Python calls into torch.  Each entry point below therefore names the reference *function*
 * (file:line under /root/reference/stereo_toolbox/models/) whose arithmetic it replaces; the
 * Python binding that the toolbox-side maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers + sizes, no torch types; all pointers are DEVICE pointers unless noted
 *   - `stream` is a cudaStream_t passed as void*; every call is stream-ordered and re-entrant,
 *     allocates nothing and keeps no global state (tensor maps are built per call on the host)
 *   - return value: 0 = ok; <0 = STB_E_* (the Python wrapper raises RuntimeError with
 *     stb_error_string()); a CUDA launch error is returned as -(1000 + cudaError_t)
 *   - fp32 tensors are NCDHW / NCHW contiguous exactly like the reference's; the bf16 fast path
 *     uses channels-last NDHWC (documented per function)
 */
#ifndef STB200_H
#define STB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STB_OK 0
#define STB_E_BADARG (-1)
#define STB_E_UNSUPPORTED (-2)
#define STB_E_SMEM (-3)
#define STB_E_DRIVER (-4)

/* activation codes (epilogue of the conv family) */
#define STB_ACT_NONE 0
#define STB_ACT_RELU 1
#define STB_ACT_LEAKY 2   /* LeakyReLU(0.01), IGEVStereo/submodule.py:36 */
#define STB_ACT_MISH 3    /* x*tanh(softplus(x)), CFNet/submodule.py:99-106 */
#define STB_ACT_SIGMOID 4 /* ConvGRU gates z, r (RAFTStereo/update.py:40-41); stb_conv3d_umma: split storage only */
#define STB_ACT_TANH 5    /* ConvGRU candidate q (RAFTStereo/update.py:42); stb_conv3d_umma: split storage only */

const char* stb_error_string(int code);
int stb_version(void);

/* ---- cost-volume builders (fp32, NCDHW out) ------------------------------------------------
 * build_gwc_volume: GwcNet/submodule.py:53-63 (+ACVNet:228-238, CFNet:171-181, PCWNet:109-119,
 * IGEVStereo:161-171) incl. groupwise_correlation :44-50.
 *   vol[b, c_off+g, d, h, w] = mean_c L[b,g*k+c,h,w] * R[b,g*k+c,h,w-d]   (w>=d), else 0
 * `vol` has c_total channels; only channels [c_off, c_off+G) are written (lets GwcNet_GC build
 * gwc+concat into one buffer instead of torch.cat, GwcNet/gwcnet.py:180). */
int stb_gwc_volume_f32(const float* left, const float* right, float* vol,
                       int B, int C, int H, int W, int D, int G, int c_total, int c_off, void* stream);

/* build_concat_volume: variant A (mask_left=1) GwcNet/submodule.py:30-41, CFNet:141-152,
 * PCWNet:86-97, inline PSMNet/stackhourglass.py:111-120; variant B (mask_left=0)
 * ACVNet/submodule.py:180-191, IGEVStereo/submodule.py:208-219.
 * Writes channels [c_off, c_off+2C) of a c_total-channel volume.
 * att_prob (nullable): [B,1,D,H,W] probabilities = softmax_d(att_weights) from stb_softmax_d_f32;
 * when given every written value is multiplied by it -- ACVNet/acv.py:196 fused into the build. */
int stb_concat_volume_f32(const float* left, const float* right, const float* att_prob, float* vol,
                          int B, int C, int H, int W, int D, int mask_left, int c_total, int c_off,
                          void* stream);

/* CFNet sampled cost volume (cascade stages): cost_volume_generator 'gwc' + 'concat' + the sample channel, concatenated
 * (CFNet/cfnet.py:472-496, 545-550; SpatialTransformer CFNet/submodule.py:302-349; groupwise_correlation_4D :162-168).
 *   gw_*  [B,Cg,H,W], cat_* [B,Cc,H,W] (nullable when Cc == 0), samples [B,S,H,W] fp32 holding integer disparities
 *   vol   [B, G + 2*Cc + 1, S, H, W]: G group-wise correlations at x = w - sample, Cc left channels (broadcast over S),
 *         Cc right channels gathered at x, the samples themselves; x clamped to [0, W-1], gathered values zeroed where
 *         w - sample falls outside the row.  Opt-in on the host side (STB_CFNET_SAMPLED), not yet run on hardware. */
int stb_sampled_volume_f32(const float* gw_left, const float* gw_right, const float* cat_left, const float* cat_right,
                           const float* samples, float* vol, int B, int Cg, int G, int Cc, int S, int H, int W,
                           void* stream);

/* F.softmax(x, dim=D axis) of a [B,D,plane] tensor (ACVNet/acv.py:196, plane = H*W). */
int stb_softmax_d_f32(const float* x, float* y, int B, int D, long long plane, void* stream);

/* ---- head: F.upsample(trilinear) + softmax over D + disparity_regression, fused -------------
 * GwcNet/gwcnet.py:220-223 + GwcNet/submodule.py:23-27; PSMNet/stackhourglass.py:150-156;
 * align_corners=1: CFNet/cfnet.py:605-613, PCWNet/pcwnet.py:486.  With out sizes equal to the
 * input sizes it is the plain softmax+regression of IGEVStereo/igev_stereo.py:212-213.
 * cost [B,D,H,W] fp32  ->  disp [B,outH,outW] fp32 ; the [B,outD,outH,outW] tensor is never
 * materialised. */
int stb_upsample_softargmin_f32(const float* cost, float* disp, int B, int D, int H, int W,
                                int outD, int outH, int outW, int align_corners, void* stream);

/* disparity_regression on an explicit probability volume: GwcNet/submodule.py:23-27,
 * PSMNet/submodule.py:46-54.  prob [B,D,plane] -> disp [B,plane], disp = sum_d d*prob[d]. */
int stb_disparity_regression_f32(const float* prob, float* disp, int B, int D, long long plane, void* stream);

/* ---- 3-D convolution family, exact fp32 path (CUDA cores) ------------------------------------
 * convbn_3d (+ReLU/Mish/LeakyReLU, + residual before the activation): PSMNet/submodule.py:16-19,
 * GwcNet/gwcnet.py:72-105; ConvTranspose3d(k3,s2,p1,op1) PSMNet/stackhourglass.py:25-29;
 * k4 s2 p1 IGEVStereo/igev_stereo.py:43-50.  One call computes
 *   out[b,co, jd*os+od0, jh*os+oh0, jw*os+ow0] =
 *       act( sum_t sum_ci wt[t][ci][co] * x[b,ci, jd*is+dd[t], jh*is+dh[t], jw*is+dw[t]]
 *            + shift[co] + residual[same index] )            for (jd,jh,jw) in [0,nd)x[0,nh)x[0,nw)
 * with zero padding outside x.  A strided conv is is=2; a transposed conv is one call per output
 * parity class (os=2).  wt is the tap-major repack [ntaps][Cin][Cout] (BN scale pre-multiplied),
 * dd/dh/dw are HOST arrays of ntaps offsets (ntaps <= 64).  x [B,Cin,Di,Hi,Wi], out/residual
 * [B,Cout,Do,Ho,Wo], all fp32 NCDHW. shift/residual may be NULL. */
int stb_conv3d_taps_f32(const float* x, const float* wt, const float* shift, const float* residual,
                        float* out, int B, int Cin, int Di, int Hi, int Wi, int Cout, int Do, int Ho,
                        int Wo, int ntaps, const int* dd, const int* dh, const int* dw, int in_stride,
                        int out_stride, int od0, int oh0, int ow0, int nd, int nh, int nw, int act,
                        void* stream);

/* ---- tensor-core path (tcgen05 / TMEM / TMA), channels-last 16-bit ---------------------------------
 * Layout: activations NDHWC x[b][d][h][w][c], 16-bit: f16 = 0 -> bf16, f16 = 1 -> fp16 (same tensor-core
 * rate, 3 more mantissa bits).  Same reference layers as stb_conv3d_taps_f32.
 *
 * stb_volume_cl16: build_gwc_volume (+ build_concat_volume variant A/B + torch.cat,
 * GwcNet/gwcnet.py:175-180; PSMNet/stackhourglass.py:111-120 with G=0) written once as
 * vol[b][d][h][w][Ct_pad] (channels [0,G) correlation groups, [G,G+Cc) left, [G+Cc,G+2Cc) right,
 * the rest zero).  Inputs fp32 NCHW. Requires G % 8 == 0, Ct_pad % 8 == 0. */
int stb_volume_cl16(const float* gwc_l, const float* gwc_r, const float* cat_l, const float* cat_r,
                    void* vol, int f16, int B, int Cg, int G, int Cc, int H, int W, int D, int Ct_pad,
                    int mask_left, void* stream);

/* Same volume from CHANNELS-LAST 16-bit features, as the tensor-core extractor produces them: feats[i] is
 * [2B][H][W][feat_ch[i]] (left = image b, right = image B+b; the channels of the nfeat <= 4 tensors concatenate to the
 * 8*G correlation channels -- GwcNet's layer2/3/4 outputs, no torch.cat), cat [2B][H][W][cat_c] holds the Cc concat
 * channels (nullable when Cc = 0).  feats / feat_ch are HOST arrays.  Saves the NCHW fp32 copies of the features. */
int stb_volume_cl16_from_cl16(const void* const* feats, const int* feat_ch, int nfeat, const void* cat, int cat_c,
                              void* vol, int f16, int B, int G, int Cc, int H, int W, int D, int Ct_pad,
                              int mask_left, void* stream);

/* Layout boundary helpers: [B,C,S] fp32 <-> [B,S,Cpad] 16-bit (S = D*H*W). */
int stb_ncdhw_to_cl16(const float* src, void* dst, int f16, int B, int C, long long S, int Cpad, void* stream);
int stb_cl16_to_ncdhw(const void* src, float* dst, int f16, int B, int C, long long S, int Cpad, void* stream);

/* Implicit-GEMM conv on tcgen05: every tap is a row-shifted view of a TMA-staged zero-padded tile
 * (see csrc/conv3d_umma.cu).  x [B,Di,Hi,Wi,Cin]; KC in {16,32,64} = channels per K-chunk (Cin % KC == 0;
 * Cin/KC > 1 runs K-split passes chained through the fp32 workspace ws[B,Do,Ho,Wo,Cout_total]);
 * wt [nwtiles][Cin/KC][Cpad][KC], Cpad = Cout_total rounded up to 16 (BN scale folded);
 * out/residual [B,Do,Ho,Wo,Cout_total] 16-bit (out fp32 if out_fp32).
 * `nclass` output classes: class c owns taps [tap_begin[c],tap_end[c]) and writes output
 * (s*os+od0[c], jh*os+oh0[c], jw*os+ow0[c]) for s in [0,nsteps), jh in [0,nclass_h), jw in [0,nclass_w).
 * in_stride 1: tap t reads input plane s+dz[t] at (jh+in_h_off+dh[t], jw+in_w_off+dw[t]).
 * in_stride 2: tap t reads plane 2s+dz[t], parity sub-tile sub[t] = 2*ph+pw, at half-resolution index
 * (jh+in_h_off+dh[t], jw+in_w_off+dw[t]) i.e. input (2*(..)+ph, 2*(..)+pw).   dh,dw in [0,3]; weight tile
 * widx[t].  nblk[t]/cls0[t] (nullable): the tap's MMA spans nblk consecutive weight tiles (N = nblk*Cn) and
 * accumulates into column blocks [cls0, cls0+nblk) -- kw-merge uses 3 blocks, the merged transposed conv
 * (flags bit3) 8 parity-class blocks drained in one round.  All index arrays are HOST arrays.
 * flags: bit0 UMMA base-offset convention, bit1 TMA element-stride box convention (both settled by
 * csrc/probe/umma_probe.cu), bit2 kw-merge, bit3 merged transposed conv, bit4 2-D convolution (dz = 0,
 * the image index is the depth axis and is never strided), bits 8-10 dilation of the kw-merged taps (0 = 1),
 * bit5 K-chunks along a pseudo-depth axis (plane P = depth*G + chunk, taps carry dz*G + chunk: the chunks
 * accumulate in TMEM; weight tiles ordered [pass][chunk][tile], nwtiles = tiles of one pass) with bits 11-13 = G
 * chunks per pass (0 = all Cin/KC chunks in one pass; else (Cin/KC)/G passes chained through ws), bit6 operand-split
 * fp16 storage ("fp16x2"; bits 16-22: exponent of the weight pre-scale), bit7 stride-2 kw = 0 / 2 pair merge,
 * bit14 (with bit3) the 8 parity-class column blocks hold the classes 0,1,3,2,6,7,5,4 (Gray order) instead of 0..7;
 * dchunk: depth steps per CTA
 * (0 = auto).  Cout_valid = real (unpadded) channels. */
int stb_conv3d_umma(const void* x, const void* wt, const float* shift, const void* residual, void* out,
                    float* ws, int f16, int B, int Cin, int KC, int Di, int Hi, int Wi, int Cout_total,
                    int Cout_valid, int Do, int Ho, int Wo, int ntaps, const int* dz, const int* dh,
                    const int* dw, const int* sub, const int* widx, const int* nblk, const int* cls0,
                    int nwtiles, int nclass,
                    const int* tap_begin, const int* tap_end, const int* od0, const int* oh0,
                    const int* ow0, int in_stride, int out_stride, int nsteps, int nclass_h, int nclass_w,
                    int in_h_off, int in_w_off, int act, int out_fp32, int flags, int dchunk,
                    void* stream);

/* Debug aid: arm (dev_buf = [rounds][4] int64 device buffer) or disarm (NULL) the in-kernel timeline of CTA 0 of
 * subsequent stb_conv3d_umma launches: per accumulator round clock64() at issue start / commit / epilogue wake /
 * buffer release.  Not used by the product path. */
int stb_conv3d_umma_set_trace(long long* dev_buf, int rounds);

/* CUDA-core companion on the same channels-last 16-bit tensors (fp32 weights wt[ntaps][Cin][Cout], fp32
 * accumulate); arguments as stb_conv3d_taps_f32. */
int stb_conv3d_taps_cl16(const void* x, const float* wt, const float* shift, const void* residual,
                         void* out, int out_fp32, int f16, int B, int Cin, int Di, int Hi, int Wi, int Cout,
                         int Do, int Ho, int Wo, int ntaps, const int* dd, const int* dh, const int* dw,
                         int in_stride, int out_stride, int od0, int oh0, int ow0, int nd, int nh,
                         int nw, int act, void* stream);

/* ---- 1-D all-pairs correlation + pyramid + lookup --------------------------------------------
 * CorrBlock1D.corr (RAFTStereo/corr.py:148-156, scale=1/sqrt(C)) and
 * Combined_Geo_Encoding_Volume.corr (IGEVStereo/geometry.py:62-70, scale=1).
 * f1 [B,C,H,W1], f2 [B,C,H,W2] fp32 -> corr [B,H,W1,W2] fp32 (level 0), corr *= scale. */
int stb_corr1d_f32(const float* f1, const float* f2, float* corr, int B, int C, int H, int W1, int W2,
                   float scale, void* stream);

/* F.avg_pool2d(x,[1,2],stride=[1,2]) along the last axis (RAFTStereo/corr.py:123-125,
 * IGEVStereo/geometry.py:24-30): src [rows, Wsrc] -> dst [rows, Wsrc/2]. */
int stb_avgpool_last_f32(const float* src, float* dst, long long rows, int Wsrc, void* stream);

/* CorrBlock1D.__call__ (RAFTStereo/corr.py:127-146) with bilinear_sampler
 * (RAFTStereo/utils/utils.py:59-74): all levels x (2r+1) taps in one launch.
 * pyr[l] (HOST array of device pointers) [B,H,W1,W2>>l]; coords_x [B,H,W1] (channel 0 of the
 * coords tensor); out [B, levels*(2r+1), H, W1] fp32. Out-of-range taps contribute 0. */
int stb_corr1d_lookup_f32(const float* const* pyr, const float* coords_x, long long coords_bstride,
                          float* out, int B, int H, int W1, int W2, int levels, int radius, void* stream);

/* Combined_Geo_Encoding_Volume.__call__ (IGEVStereo/geometry.py:35-59).
 * geo[l] [B,H,W,C,D>>l], corr[l] [B,H,W,W2>>l] (HOST arrays of device pointers);
 * disp, coords_x [B,H,W]; out [B, levels*(2r+1)*(C+1), H, W]. */
int stb_geo_lookup_f32(const float* const* geo, const float* const* corr, const float* disp,
                       const float* coords_x, float* out, int B, int H, int W, int C, int D, int W2,
                       int levels, int radius, void* stream);

/* geo volume [B,C,D,H,W] -> [B,H,W,C,D] (IGEVStereo/geometry.py:19). */
int stb_geo_permute_f32(const float* src, float* dst, int B, int C, int D, int H, int W, void* stream);

/* ---- ACVNet pieces -----------------------------------------------------------------------------
 * Depthwise (1,3,3) dilated "patch" convolution, ACVNet/acv.py:109-112 (nn.Conv3d(C, C, (1,3,3), groups=C,
 * dilation=d, padding=(0,d,d), bias=False)) as used at :169-173.  x, out [B,c_total,D,H,W] fp32 (out != x);
 * channels [c_off, c_off+C) are processed with weight [C,1,1,3,3]; the other channels of `out` are untouched, so
 * patch_l1/l2/l3 write their slices of one buffer (replaces the torch.cat of :173). */
int stb_patch_dw_f32(const float* x, const float* weight, float* out, int B, int c_total, int c_off, int C,
                     int D, int H, int W, int dilation, void* stream);

/* Block self-attention core of attention_block.forward, ACVNet/submodule.py:381-428, between the qkv Linear and
 * final1x1: softmax(q k^T * head_dim^-0.5 + mask) v inside non-overlapping (b0,b1,b2) blocks, `heads` heads.
 * qkv holds the Linear's output over the UN-padded volume, logical shape [B,3C,D,H0,W0] with the element strides
 * qkv_strides[5] = (b,c,d,h,w) (so NCDHW fp32 and NDHWC 16-bit tensors both bind); channel = which*C + head*hd + e.
 * H0/W0 are padded up to multiples of b1/b2 implicitly: padded tokens take q,k,v = qkv_bias [3C] (the reference
 * pads with zeros BEFORE the Linear), and the pad mask follows :403-409 including its "-0:" slice behaviour.
 * out: logical [B,C,D,H0,W0] with out_strides[5]; dtype 0 = fp32, 1 = fp16, 2 = bf16 (both tensors).
 * D must be a multiple of b0 (the reference's view() requires it). */
int stb_block_attention(const void* qkv, const float* qkv_bias, void* out, int dtype, int B, int C, int heads,
                        int D, int H0, int W0, int b0, int b1, int b2, const long long* qkv_strides,
                        const long long* out_strides, void* stream);

/* ---- IGEV / CFNet elementwise pieces ---------------------------------------------------------------
 * FeatureAtt gate, IGEVStereo/submodule.py:228-241 (used at igev_stereo.py:71-88,208): out = x * sigmoid(gate)
 * with the 2-D gate logits [B,C,H,W] (fp32) broadcast over D.  x/out [B,C,D,H,W] fp32, or channels-last 16-bit
 * [B,D,H,W,Cpad] (channels >= C are padding and pass through). out may alias x. */
int stb_feature_gate_f32(const float* x, const float* gate, float* out, int B, int C, int D, int H, int W, void* stream);
int stb_feature_gate_cl16(const void* x, const float* gate, void* out, int f16, int B, int C, int Cpad, int D, int H,
                          int W, void* stream);

/* disparity_variance, CFNet/submodule.py:127-133: prob [B,D,plane], disp [B,plane] -> var [B,plane]
 * var = sum_d prob[d] * (d - disp)^2. */
int stb_disparity_variance_f32(const float* prob, const float* disp, float* var, int B, int D, long long plane,
                               void* stream);

/* ---- learned convex 9-tap upsampling of the iterative models -----------------------------
 * stb_convex_upsample_f32 replaces RAFTStereo.upsample_flow (RAFTStereo/raft_stereo.py:81-93): softmax over the 9 mask
 * logits of every fine pixel + convex combination of the 3x3 coarse neighbourhood of factor*flow (zero outside).
 * flow [N,D,H,W], mask [N, 9*factor^2, H, W] (channel = tap*factor^2 + fy*factor + fx) -> out [N,D,factor*H,factor*W].
 * factor 2, 4 or 8.
 * stb_context_upsample_f32 replaces context_upsample (IGEVStereo/submodule.py:243-255): out[b,Y,X] = sum_t w[b,t,Y,X] *
 * scale * disp_low[b, Y/factor + t/3 - 1, X/factor + t%3 - 1]; apply_softmax != 0 soft-maxes the 9 weights first (the
 * F.softmax(..., 1) the caller applies at igev_stereo.py:164). disp_low [B,1,h,w], weights [B,9,factor*h,factor*w]. */
int stb_convex_upsample_f32(const float* flow, const float* mask, float* out, int N, int D, int H, int W, int factor,
                            void* stream);
int stb_context_upsample_f32(const float* disp_low, const float* weights, float* out, int B, int h, int w, int factor,
                             float scale, int apply_softmax, void* stream);

/* ---- backward kernels (training path, exact fp32, reference layouts) --------------------------------------
 * The DATA gradient of the conv family needs no entry point of its own: the adjoint of Conv3d(k,s,p) is
 * ConvTranspose3d(k,s,p,op) with the same weight tensor and vice versa, i.e. another stb_conv3d_taps_f32 call.
 *
 * Weight gradient of Conv3d / ConvTranspose3d (PSMNet/submodule.py:16-19, PSMNet/stackhourglass.py:25-29):
 *   dW[kd][kh][kw][cp][cq] += sum_{b,q} P[b][cp][stride*q + k - pad] * Q[b][cq][q]        (dW must be zeroed by the caller)
 * Conv3d: P = layer input, Q = grad of the output (weight.grad[co,ci,k] = dW[k][ci][co]);
 * ConvTranspose3d: P = grad of the output, Q = layer input (weight.grad[ci,co,k] = dW[k][co][ci]).
 * P [B,Cp,Dp,Hp,Wp], Q [B,Cq,Dq,Hq,Wq] fp32; (K,stride) in {(1,1),(3,1),(3,2),(4,2)}. */
int stb_conv3d_wgrad_f32(const float* P, const float* Q, float* dW, int B, int Cp, int Dp, int Hp, int Wp, int Cq,
                         int Dq, int Hq, int Wq, int K, int pad, int stride, void* stream);

/* The same weight gradient on the tensor cores, straight from the channels-last 16-bit tensors of the 16-bit training path
 * (train16.py; BASELINE config 3, reference: trainer/trainer_torchrun.py:105-123 runs the step under autocast):
 *   P [B,Dp,Hp,Wp,Cp], Q [B,Dq,Hq,Wq,Cq] bf16 (f16 = 0) or fp16 (f16 = 1), Cp and Cq multiples of 32 (zero-padded channels
 *   give zero rows / columns); dW [K^3][Cp][Cq] fp32, zeroed by the caller (fp32 accumulation, atomics across CTAs).
 * K in 1..3, stride in {1, 2}; max_ctas <= 0: one persistent CTA per SM (and per tap / channel group). */
int stb_conv3d_wgrad_cl16(const void* P, const void* Q, float* dW, int f16, int B, int Dp, int Hp, int Wp, int Cp, int Dq,
                          int Hq, int Wq, int Cq, int K, int pad, int stride, int max_ctas, void* stream);

/* Between the convolutions of the iterative models' update block (RAFTStereo/update.py:29-44 ConvGRU, :91-95 pool2x /
 * interp, :115-138 BasicMultiUpdateBlock.forward), on channels-last operand-split fp16 rows ([.., 2*C] halves, C % 16 == 0):
 *   stb_gru_rh_split     out[p][c] = r * h, r = logical channels [C, 2C) of zr ([npix][2C] logical: sigmoid(convz | convr))
 *   stb_gru_blend_split  out = (1 - z) * h + z * q, z = logical channels [0, C) of zr
 *   stb_pool2x_split     F.avg_pool2d(x, 3, stride=2, padding=1): x [N,H,W,C] -> out [N,(H-1)/2+1,(W-1)/2+1,C]
 *   stb_interp_split     F.interpolate(x, (Ho,Wo), mode="bilinear", align_corners=True): x [N,H,W,C] -> out [N,Ho,Wo,C] */
int stb_gru_rh_split(const void* zr, const void* h, void* out, long long npix, int C, void* stream);
int stb_gru_blend_split(const void* zr, const void* h, const void* q, void* out, long long npix, int C, void* stream);
int stb_pool2x_split(const void* x, void* out, int N, int H, int W, int C, void* stream);
int stb_interp_split(const void* x, void* out, int N, int H, int W, int Ho, int Wo, int C, void* stream);

/* PCWNet's full-resolution refinement inputs (PCWNet/submodule.py:122-152 `warp`, :104-120 `build_corrleation_volume` with
 * num_groups = 1; PCWNet/pcwnet.py:491-506):
 *   stb_warp_disp_f32: x [B,C,H,W], disp [B,1,H,W] -> out [B,C,H,W] = grid_sample(x, grid(x - disp)) * (grid_sample(1) >= 0.999)
 *     with the reference's grid arithmetic (normalised with W-1 / H-1, sampled with align_corners=False, zero padding);
 *   stb_corr_volume_1d_f32: left, right [B,C,H,W] -> vol [B,2*maxdisp+1,H,W] (every element written); maxdisp in {4, 8, 24}. */
int stb_warp_disp_f32(const float* x, const float* disp, float* out, int B, int C, int H, int W, void* stream);
int stb_corr_volume_1d_f32(const float* left, const float* right, float* vol, int B, int C, int H, int W, int maxdisp,
                           void* stream);

/* Adjoint of build_concat_volume (variant A mask_left=1 / B mask_left=0): dvol [B,c_total,D,H,W] channels
 * [c_off, c_off+2C) -> dleft, dright [B,C,H,W]. */
int stb_concat_volume_bwd_f32(const float* dvol, float* dleft, float* dright, int B, int C, int H, int W, int D,
                              int mask_left, int c_total, int c_off, void* stream);

/* Adjoint of build_gwc_volume: dvol channels [c_off, c_off+G) of [B,c_total,D,H,W]; left/right [B,C,H,W] are the
 * forward inputs -> dleft, dright [B,C,H,W]. */
int stb_gwc_volume_bwd_f32(const float* dvol, const float* left, const float* right, float* dleft, float* dright,
                           int B, int C, int H, int W, int D, int G, int c_total, int c_off, void* stream);

/* Adjoint of stb_upsample_softargmin_f32: cost [B,D,H,W] (the forward input), gdisp [B,outH,outW] -> dcost [B,D,H,W]
 * is ACCUMULATED into (zero it first). */
int stb_upsample_softargmin_bwd_f32(const float* cost, const float* gdisp, float* dcost, int B, int D, int H, int W,
                                    int outD, int outH, int outW, int align_corners, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* STB200_H */
