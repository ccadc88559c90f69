/*
 * shgan_b200 -- C ABI of the B200-native SH-GAN generator-forward hot path.
 *
 * Drop-in boundary.  Every entry point takes raw DEVICE pointers to
 * caller-allocated, contiguous buffers, plain integer sizes, and an explicit cudaStream_t
 * (passed as void*), returns 0 on success or a non-zero code, and never synchronises the
 * host.  On failure `shgan_last_error()` returns a human-readable message (thread-local).
 * No torch types appear in any signature.  The reference interfaces each function replaces
 * are cited as reference file:line (relative to SHI-Labs/SH-GAN @ a9ba83c5).
 *
 * Activation layout used between fused layers ("split planes"): a tensor [N,H,W,C] is held
 * as two fp16 NHWC planes hi/lo with value = hi + lo (22 significant bits).  The planes are
 * directly the A operands of the 3-pass fp16 tcgen05 MMA (hi*hi + hi*lo + lo*hi) that gives
 * fp32-class accuracy on the tensor cores.
 */
#ifndef SHGAN_B200_H_
#define SHGAN_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SHGAN_B200_ABI_VERSION 3
#define SHGAN_MAX_TAPS 16
#define SHGAN_MAX_SRC 4

/* ---- library ------------------------------------------------------------------------- */
int shgan_abi_version(void);
const char* shgan_last_error(void);
/* number of kernels launched by this library in the calling process so far (bench.py:gpu_launches) */
uint64_t shgan_launch_count(void);

/* ---- upfirdn2d ------------------------------------------------------------------------
 * replaces: pybind `upfirdn2d(x,f,upx,upy,downx,downy,padx0,padx1,pady0,pady1,flip,gain)`
 *           lib/model_zoo/stylegan_utils/upfirdn2d.cpp:16-94 (+ kernels upfirdn2d.cu:29-200)
 * x [N,C,H,W] fp32 NCHW contiguous; f [fH,fW] fp32; y [N,C,outH,outW] caller-allocated with
 * outW = (W*upx + padx0 + padx1 - fW + downx) / downx  (upfirdn2d.cpp:32-33), same for H. */
int shgan_upfirdn2d_fwd(const float* x, const float* f, float* y,
                        int N, int C, int H, int W, int fH, int fW,
                        int upx, int upy, int downx, int downy,
                        int padx0, int padx1, int pady0, int pady1,
                        int flip, float gain, void* stream);

/* ---- layout conversion ------------------------------------------------------------------
 * NCHW fp32 -> split planes NHWC.  Optional fused terms, all in fp32 before the split:
 *   v = x[n,c,y,x] (+ add_hi+add_lo planes) (* scale[n,c])
 * c_off/c_tot allow writing a channel slice [c_off, c_off+C) of planes that have c_tot channels. */
int shgan_nchw_to_planes(const float* x, const void* add_hi, const void* add_lo, const float* scale,
                         void* out_hi, void* out_lo, int N, int C, int H, int W,
                         int c_off, int c_tot, void* stream);
/* split planes (channel slice [c_off,c_off+C) of c_tot) -> NCHW fp32 */
int shgan_planes_to_nchw(const void* in_hi, const void* in_lo, float* y, int N, int C, int H, int W,
                         int c_off, int c_tot, void* stream);
/* planes[.., c_off:c_off+C] += x (NCHW fp32), re-split in place.  replaces the split/add/cat of
 * lib/model_zoo/shgan.py:378-382 without copying the untouched channels. */
int shgan_planes_add_nchw(void* hi, void* lo, const float* x, int N, int C, int H, int W,
                          int c_off, int c_tot, void* stream);
/* shgan_planes_add_nchw for up to SHGAN_MAX_ADD tensors of different spatial size in one launch: band k adds
 * x[k] [N,C,hw[k]] (NCHW fp32) into channels [c_off[k], c_off[k]+C) of the planes hi[k]/lo[k] [N,hw[k],c_tot[k]].
 * The SHU add-back over all of its bands (lib/model_zoo/shgan.py:378-382).  work_start is filled in by the library. */
#define SHGAN_MAX_ADD 8
typedef struct {
    int num;
    void* hi[SHGAN_MAX_ADD];
    void* lo[SHGAN_MAX_ADD];
    const void* x[SHGAN_MAX_ADD];
    int hw[SHGAN_MAX_ADD], c_off[SHGAN_MAX_ADD], c_tot[SHGAN_MAX_ADD];
    long long work_start[SHGAN_MAX_ADD + 1];
} shgan_add_batch;
int shgan_planes_add_nchw_multi(shgan_add_batch* b, int N, int C, void* stream);
/* NHWC fp32 -> NCHW fp32 (op-level API glue) */
int shgan_nhwc_to_nchw_f32(const float* x, float* y, int N, int C, int H, int W, void* stream);

/* ---- pointwise epilogue parameters shared by the conv and FIR kernels -------------------
 *   v = acc * (dcoef ? dcoef[n,o] : 1) * wgain
 *   v += noise[n*noise_sn + y*W + x] * (*noise_strength)          (if noise)
 *   v += bias[o]                                                  (if bias)
 *   v = clamp(lrelu(v, act_alpha) * act_gain, +-act_clamp)        (if act; act_clamp<=0: none)
 *       else v *= act_gain
 *   v += skip_hi+skip_lo [n,y,x,o]                                (if skip)
 *   rgb[n,y,x,blk,j] = sum_o v * rgb_w[j,o] * rgb_style[n,o]      (if rgb_w; partial per N-block)
 *   v_out = v * next_scale[n,o]                                   (if next_scale)
 * replaces the separate ATen passes of lib/model_zoo/stylegan.py:191-192,226-238,298-303 and
 * common/utils.py:135-143, the skip add of comodgan.py:319-327 and torgb of stylegan.py:325-337. */
typedef struct {
    const float* dcoef;          /* [N,Co] or NULL */
    float wgain;
    const float* noise;          /* fp32, indexed n*noise_sn + y*OW + x, or NULL */
    int64_t noise_sn;            /* 0 = shared const noise */
    const float* noise_strength; /* device scalar (required when noise != NULL) */
    const float* bias;           /* [Co] or NULL */
    int act;
    float act_alpha, act_gain, act_clamp;
    const void* skip_hi;         /* planes [N,OH,OW,Co] or NULL */
    const void* skip_lo;
    const float* next_scale;     /* [N,Co] or NULL */
    const float* rgb_w;          /* [3,Co] or NULL */
    const float* rgb_style;      /* [N,Co] */
    float* rgb_out;              /* [N,OH,OW,n_blocks,4] */
    void* out_hi;                /* planes [N,OH,OW,Co] or NULL */
    void* out_lo;
    float* out_f32;              /* NHWC fp32 [N,OH,OW,Co] of v (before next_scale) or NULL */
} shgan_epilogue;

/* ---- implicit-GEMM convolution on tcgen05 tensor cores ----------------------------------
 * replaces: F.conv2d / F.conv_transpose2d (cuDNN) reached through
 *           lib/model_zoo/stylegan_utils/conv2d_resample.py:26-51,57-154 and the per-sample
 *           weight materialisation of lib/model_zoo/stylegan.py:149-190 (modulated_conv2d).
 * acc[n,oy,ox,o] = sum_{t<ntaps} sum_{c<C} src[tap_src[t]][n, oy+tap_dy[t], ox+tap_dx[t], c]
 *                                          * w[tap_w[t], o, c]        (zero outside a source)
 * computed as hi*hi + hi*lo + lo*hi in fp16 with fp32 accumulation in TMEM.
 * mode 0 (ACT): apply `epi`, write planes/fp32/rgb.   mode 1 (RAW): write acc as fp32 NHWC into
 * z[n, oy*zsy+zoy, ox*zsx+zox, o] of a [N,ZH,ZW,Co] tensor (used by the 4 parity passes of the
 * stride-2 transposed convolution).  C and Co must be multiples of 64.  In ACT mode with rgb_w set, the
 * torgb partial sums of each block of 32 output channels are written to
 * rgb_out[n,y,x,blk,0..2] (blk < shgan_conv_num_nblocks); shgan_torgb_combine adds them up.
 * ACT mode with z != NULL: z is an optional split-K scratch of ZH >= 2 buffers of ZW >= N*OH*OW*Co floats (16-byte aligned).
 * Layers with too few output tiles to fill the GPU (4x4 / 8x8) are then computed by several CTAs per tile, each over a slice of
 * the (tap, channel) loop, their fp32 partials summed in a fixed order by a second launch that applies `epi`; any other layer
 * ignores the scratch.  Results are deterministic and agree with the un-split launch to fp32 rounding. */
typedef struct {
    int num_src;
    const void* src_hi[SHGAN_MAX_SRC];
    const void* src_lo[SHGAN_MAX_SRC];
    int src_h[SHGAN_MAX_SRC], src_w[SHGAN_MAX_SRC];
    int N, C, Co;
    const void* w_hi;            /* packed [w_taps, Co, C] fp16 */
    const void* w_lo;
    int w_taps;
    int ntaps;
    int tap_src[SHGAN_MAX_TAPS], tap_dy[SHGAN_MAX_TAPS], tap_dx[SHGAN_MAX_TAPS], tap_w[SHGAN_MAX_TAPS];
    int OH, OW;
    int mode;
    float* z; int ZH, ZW, zsy, zsx, zoy, zox;
    shgan_epilogue epi;
    int block_n;                 /* GEMM tile width: 0 = auto (widest of 64/128/256 that still fills the SMs) */
    int passes;                  /* 3 = fp32-class (default when 0), 1 = hi*hi only (fast, ~fp16 accuracy) */
    int impl;                    /* 0 = tcgen05 tensor-core path (the product): the halo-tile kernel for the large layers,
                                        the per-tap kernel for the small ones, chosen per layer;
                                    1 = rejected by this library: the fp32 FMA cross-check kernel with identical operands /
                                        epilogue is built into the TEST-ONLY libshgan_b200_check.so (shgan_check_conv_igemm);
                                    2 / 3 = force the per-tap / the halo-tile tensor-core kernel (tests, profiling);
                                    4 = the two-SM (cta_group::2) kernel wherever Co % 256 == 0, per-tap elsewhere */
    float acc_comp;              /* compensation of the tensor core's TRUNCATING fp32 accumulate (measured on B200: every
                                    chained tcgen05.mma loses ~half an ulp of the running sum towards zero, a relative
                                    shrink of 1.6e-8 .. 1.8e-8 per chained MMA, tools/acc_bias_probe.py): each chunk of L chained MMAs is scaled by
                                    (1 + acc_comp * L) when it is drained into the fp32 register sum.
                                    0 = library default (SHGAN_ACC_COMP_DEFAULT), < 0 = off, > 0 = this value */
} shgan_conv_desc;
#define SHGAN_ACC_COMP_DEFAULT 1.6e-8f
int shgan_conv_igemm(const shgan_conv_desc* d, void* stream);
/* size of the rgb partial axis: Co / 32 (independent of block_n, kept in the signature for ABI stability) */
int shgan_conv_num_nblocks(int Co, int block_n);

/* ---- up-sampling (stride-2 transposed) 3x3 convolution + 4x4 blur + pointwise epilogue, one launch --------------
 * replaces: the `up == 2` path of conv2d_resample (lib/model_zoo/stylegan_utils/conv2d_resample.py:123-142:
 *           conv_transpose2d stride 2 with the un-flipped weights -> upfirdn2d(f, pad 1, gain up^2)) as reached from
 *           modulated_conv2d (lib/model_zoo/stylegan.py:187-190), and every elementwise pass after it
 *           (stylegan.py:191-193,298-303; the `+ feats[res]` of comodgan.py:319-320).
 * z[n, 2i+ky, 2j+kx, o] += sum_c src[n,i,j,c] * w[ky*3+kx, o, c]          (z is never materialised)
 * out[n,y,x,o] = gain * sum_{a,b<4} fy[a]*fx[b] * z[n, y+a-1, x+b-1, o]   (z = 0 outside [0,2H] x [0,2W]),  y < 2H, x < 2W
 * then `epi` as documented on shgan_epilogue (no torgb terms).  The blur must be separable (fy (x) fx, as applied, i.e.
 * already flipped by the caller); src are split planes [N,H,W,C]; weights are packed per block of 64 output channels in
 * the order the kernel's stacked-N MMAs consume them:
 *   w_hi/w_lo fp16 [Co/64][9][64][C], slot s holding tap (ky*3+kx) = {3,0,1,4,5,2,6,7,8}[s].   C, Co multiples of 64. */
typedef struct {
    const void* src_hi;
    const void* src_lo;
    int N, H, W, C, Co;
    const void* w_hi;
    const void* w_lo;
    float fx[4], fy[4];
    float gain;
    shgan_epilogue epi;
    int passes;                  /* 3 (default when 0) or 1, as in shgan_conv_desc; | SHGAN_UP2_NARROW forces the 8-warp
                                    epilogue instance where the library would pick the 16-warp one (C <= 256): tests, profiling */
    float acc_comp;              /* as in shgan_conv_desc */
} shgan_up2_desc;
#define SHGAN_UP2_NARROW 0x100
#define SHGAN_UP2_CLUSTER 0x200     /* | into passes: run as two-CTA clusters that share each weight load by TMA multicast (measured: no */
#define SHGAN_UP2_NO_CLUSTER 0x400  /* faster than single CTAs, so off by default; kept selectable and tested) / forbid them */
int shgan_conv_up2(const shgan_up2_desc* d, void* stream);

/* ---- FIR (blur) on NHWC data with the fused pointwise epilogue ---------------------------
 * replaces: upfirdn2d blur passes around resampled convs, conv2d_resample.py:117-120 (down path,
 * pad 2) and :139 (up path, pad 1, gain 4), fused with everything that follows the blur.
 * in: either fp32 NHWC `in_f32` or split planes in_hi/in_lo, [N,IH,IW,C].
 * out[n,y,x,c] = gain * sum_{i<fH,j<fW} f[i,j] * in[n, y+i-pad_y0, x+j-pad_x0, c]  (f already flipped
 * as needed by the caller), OH = IH+pad_y0+pad_y1-fH+1, then `epi` (Co == C).
 * parity_split == 1: out planes are written de-interleaved as 4 tensors [N,PH,PW,C], plane
 * q=(y&1)*2+(x&1) at out_hi + q*N*PH*PW*C, element (y>>1, x>>1); PH=(OH+1)/2, PW=(OW+1)/2.
 * parity_split == 2: only the even/even samples are kept, out planes [N,PH,PW,C] (= upfirdn2d with down=2:
 * the discriminator's skip path, conv2d_resample.py:105-108).
 * parity_split | SHGAN_FIR_RANK1: the caller guarantees that f is an outer product fy (x) fx (the reference's
 * [1,3,3,1] blur is).  Without the flag the library decides on the device, and a second, general-filter kernel is always
 * launched behind the separable one (it returns at once for rank-1 filters); with it that launch is skipped. */
#define SHGAN_FIR_RANK1 0x100
/* parity_split | SHGAN_FIR_TWO_PHASE: with SHGAN_FIR_RANK1, planes input and an identity epilogue (the blur in front of the
 * stride-2 convolutions) the library runs a row-walking kernel; this flag selects the TMA-staged two-phase kernel instead
 * (tests, profiling). */
#define SHGAN_FIR_TWO_PHASE 0x200
int shgan_fir_nhwc(const float* in_f32, const void* in_hi, const void* in_lo,
                   const float* f, int fH, int fW, float gain,
                   int N, int C, int IH, int IW, int pad_x0, int pad_x1, int pad_y0, int pad_y1,
                   const shgan_epilogue* epi, int parity_split, void* stream);

/* ---- small-channel pointwise convs ----------------------------------------------------------
 * fromrgb: 1x1 conv Ci(<=8) -> Co with bias + lrelu_agc, NCHW fp32 in, split planes out.
 * replaces conv2d_layer.forward for the encoder's fromrgb, lib/model_zoo/stylegan.py:226-238. */
int shgan_fromrgb(const float* x, const float* w /*[Co,Ci]*/, const float* bias, float wgain,
                  float act_alpha, float act_gain, float act_clamp,
                  void* out_hi, void* out_lo, int N, int Ci, int Co, int H, int W, void* stream);
/* fromrgb with the eval loop's input preparation fused in (lib/experiments/shgan_default.py:269-274): the layer input is
 * x = cat([mask - 0.5, real * mask]) formed in registers from real [N,Ci-1,H,W] and mask [N,1,H,W]; x is also written to
 * x_out [N,Ci,H,W] (the composite fused into the last torgb reads it back).  w [Co,Ci] as in shgan_fromrgb. */
int shgan_fromrgb_masked(const float* real, const float* mask, float* x_out, const float* w, const float* bias, float wgain,
                         float act_alpha, float act_gain, float act_clamp,
                         void* out_hi, void* out_lo, int N, int Ci, int Co, int H, int W, void* stream);
/* img_out[n,j,y,x] = (img_prev ? upfirdn2d(img_prev, f, up=2, pad=[2,1,2,1], gain=4) : 0)
 *                    + sum_blk rgb_partial[n,y,x,blk,j] + bias[j]
 * replaces upsample2d + torgb add, lib/model_zoo/comodgan.py:331-338.  When `comp_x` != NULL also
 * writes the eval-loop composite uint8 image (lib/experiments/shgan_default.py:257-262). */
int shgan_torgb_combine(const float* img_prev, const float* rgb_partial, int n_blocks, const float* bias,
                        const float* f /*[4,4]*/, float* img_out, int N, int H, int W,
                        const float* comp_x /*[N,4,H,W] or NULL*/, uint8_t* comp_out, void* stream);

/* eval-loop input preparation: x[n,0] = mask - 0.5 ; x[n,1+j] = real[n,j] * mask   (all NCHW fp32; real [N,3,H,W], mask
 * [N,1,H,W] in {0,1}, x [N,4,H,W]).  replaces the sub / mul / torch.cat of lib/experiments/shgan_default.py:269-274. */
int shgan_prepare_input(const float* real, const float* mask, float* x, int N, int H, int W, void* stream);
/* out[n,0] = x[n,0] ; out[n,1+j] = x[n,1+j]*m + img[n,j]*(1-m), m = x[n,0] + 0.5: the float composite of
 * lib/experiments/shgan_default.py:257-260 concatenated with the mask channel = the discriminator's input of the
 * generator+discriminator step (BASELINE.json config C4).  x, out [N,4,H,W]; img [N,3,H,W]; NCHW fp32. */
int shgan_composite_cat(const float* x, const float* img, float* out, int N, int H, int W, void* stream);

/* minibatch_std_layer (num_channels = 1) fused with the channel concat of lib/model_zoo/stylegan.py:686-705:
 * out planes [N,H,W,C_out] = [ in planes [N,H,W,C] | std statistic of the sample's group | zeros ]. */
int shgan_mbstd_append(const void* in_hi, const void* in_lo, void* out_hi, void* out_lo, int N, int H, int W, int C,
                       int C_out, int group_size, void* stream);

/* ---- dense / styles ------------------------------------------------------------------------
 * y[b,o] = act( (sum_i x[b,i] w[o,i]) * wgain + bias[o]*bgain )   replaces dense.forward
 * (torch.addmm), lib/model_zoo/stylegan.py:87-98.  x rows are read with stride x_stride floats;
 * x may be the concatenation [x0 (I0 floats) ; x1 (I-I0 floats)] (cat of comodgan.py:252,323).
 * Layers with few outputs split I over a thread-block cluster; the partial sums are added in a fixed order
 * (results are bit-identical from run to run). */
int shgan_dense_fwd(const float* x0, int64_t x0_stride, int I0, const float* x1, int64_t x1_stride,
                    const float* w, const float* bias, float* y, int64_t y_stride,
                    int B, int I, int O, float wgain, float bgain,
                    int act, float act_alpha, float act_gain, float act_clamp, void* stream);
/* z -> z * rsqrt(mean(z^2, dim=1) + 1e-8)   (normalize_2nd_moment, stylegan.py:343-344) */
int shgan_normalize_2nd_moment(const float* z, float* y, int B, int D, void* stream);
/* style preparation for one modulated conv (stylegan.py:145-155):
 *   demod:  s_hat = s * rsqrt(mean_{all n,i} s^2) ; dcoef[n,o] = rsqrt(sum_i s_hat[n,i]^2 wsq[o,i] + 1e-8)
 *   !demod: s_hat = s * pre_scale ; dcoef untouched
 * wsq[o,i] = sum_k w_hat[o,i,k]^2 is precomputed when the weights are packed. */
int shgan_style_prep(const float* styles, const float* wsq, float* s_hat, float* dcoef,
                     int N, int Ci, int Co, int demod, float pre_scale, void* stream);
/* shgan_style_prep for every style layer of a forward in one launch.  Layer l reads its raw styles from
 * raw[n*raw_stride + offset[l] + i], i < ci[l] (the columns of ONE dense call over the concatenated affine weights of
 * all layers: stylegan.py:266,280,323,331 run on the same input [w ; x_global] whenever ws is a broadcast w), and
 * writes s_hat[l] [N,ci[l]] and, when demod[l], dcoef[l] [N,co[l]].  block_start is filled in by the library. */
#define SHGAN_MAX_STYLE_LAYERS 40
typedef struct {
    int num_layers;
    int64_t offset[SHGAN_MAX_STYLE_LAYERS];
    int ci[SHGAN_MAX_STYLE_LAYERS], co[SHGAN_MAX_STYLE_LAYERS], demod[SHGAN_MAX_STYLE_LAYERS];
    float pre_scale[SHGAN_MAX_STYLE_LAYERS];
    const void* wsq[SHGAN_MAX_STYLE_LAYERS];
    void* s_hat[SHGAN_MAX_STYLE_LAYERS];
    void* dcoef[SHGAN_MAX_STYLE_LAYERS];
    int block_start[SHGAN_MAX_STYLE_LAYERS + 1];
} shgan_style_batch;
int shgan_style_prep_batched(const float* raw, int64_t raw_stride, int N, const shgan_style_batch* tb, void* stream);

/* ---- Spectral Hint Unit ----------------------------------------------------------------------
 * replaces SHU.forward, lib/model_zoo/shgan.py:312-336 (cuFFT rfftn/irfftn + ~340 ATen calls).
 * x [N,C,R,R] fp32 NCHW (R = input_res, power of two, 4..512; C*2 <= 64)
 * conv0_w [2C,2C], conv0_b [2C], df1_w [2C, 2C*6] (reference parameter layouts),
 * cw [6,R,R/2+1] and the Gaussian band masks gauss[r] ([r, r/2+1], r = lowest_res..R, concatenated
 * lowest band first) are the constants of shgan.py:70-121,280-310.
 * packed_w: the fp16 hi/lo operands of the tensor-core channel mix (C == 32) as the kernel's pre-swizzled shared-memory
 *   image, written ONCE per parameter set by shgan_shu_pack into a 16-byte aligned buffer of shgan_shu_packed_bytes(C)
 *   bytes; NULL = pack into the workspace on every call.
 * spec_ws: 16-byte aligned workspace of shgan_shu_workspace_bytes(N,C,R) bytes (the two spectra, scratch for R > 128, and
 *   for R == 64 the blend weights in the spectrum's bin order).
 * R == 64, C == 32, lowest_res >= 4 with x and every outs[k] 16-byte aligned runs the register-resident radix-8 transforms
 *   (whole planes move by bulk copies); any other case runs the generic shared-memory transforms, same results.
 * outs[k] -> [N,C,r_k,r_k] fp32 for r_k = lowest_res * 2^k. */
int64_t shgan_shu_packed_bytes(int C);
int shgan_shu_pack(const float* conv0_w, const float* df1_w, void* packed, int C, void* stream);
int64_t shgan_shu_workspace_bytes(int N, int C, int R);
int shgan_shu_fwd(const float* x, const float* conv0_w, const float* conv0_b, const float* df1_w,
                  const float* cw, const float* gauss, const void* packed_w, void* spec_ws,
                  float* const* outs, int num_bands, int N, int C, int R, int lowest_res, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SHGAN_B200_H_ */
