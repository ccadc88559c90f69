// Geometry shared by the implementations of shgan_conv_igemm (conv_tc.cu / conv_halo.cu / conv_pair.cu: tcgen05 tensor-core
// kernels, the product path; check/conv_simt.cu: plain fp32 FMA kernel, built into the TEST-ONLY libshgan_b200_check.so).
#pragma once
#include "common.cuh"

namespace shgan {

struct ConvGeom {
    int num_src;
    const __half* src_hi[SHGAN_MAX_SRC];
    const __half* src_lo[SHGAN_MAX_SRC];
    int src_h[SHGAN_MAX_SRC], src_w[SHGAN_MAX_SRC];
    int N, C, Co;
    const __half* w_hi;
    const __half* w_lo;
    int ntaps;
    int tap_src[SHGAN_MAX_TAPS], tap_dy[SHGAN_MAX_TAPS], tap_dx[SHGAN_MAX_TAPS], tap_w[SHGAN_MAX_TAPS];
    int OH, OW;
    int mode;  // 0: epilogue, 1: raw fp32 scatter into z
    float* z;
    int ZH, ZW, zsy, zsx, zoy, zox;
    // library-internal (not in the C ABI; conv_tc.cu only): per-tap, per-pixel weights of the accumulation,
    // acc[n,y,x,o] = sum_t chunk_scale[t, y*OW + x] * (tap t's contribution).  Used by the SHU's heterogeneous filter
    // (shu.cu), whose blend over 6 anchor filters is exactly that.  NULL: plain sum.
    const float* chunk_scale;
    // relative compensation per chained MMA of the truncating tensor-core accumulate (shgan_conv_desc::acc_comp, resolved)
    float acc_comp;
};

static inline ConvGeom make_geom(const shgan_conv_desc& d) {
    ConvGeom g;
    g.num_src = d.num_src;
    for (int i = 0; i < SHGAN_MAX_SRC; ++i) {
        g.src_hi[i] = (const __half*)d.src_hi[i]; g.src_lo[i] = (const __half*)d.src_lo[i];
        g.src_h[i] = d.src_h[i]; g.src_w[i] = d.src_w[i];
    }
    g.N = d.N; g.C = d.C; g.Co = d.Co; g.w_hi = (const __half*)d.w_hi; g.w_lo = (const __half*)d.w_lo;
    g.ntaps = d.ntaps;
    for (int i = 0; i < SHGAN_MAX_TAPS; ++i) {
        g.tap_src[i] = d.tap_src[i]; g.tap_dy[i] = d.tap_dy[i]; g.tap_dx[i] = d.tap_dx[i]; g.tap_w[i] = d.tap_w[i];
    }
    g.OH = d.OH; g.OW = d.OW; g.mode = d.mode; g.z = d.z; g.ZH = d.ZH; g.ZW = d.ZW;
    g.zsy = d.zsy; g.zsx = d.zsx; g.zoy = d.zoy; g.zox = d.zox;
    g.chunk_scale = nullptr;
    g.acc_comp = d.acc_comp == 0.f ? SHGAN_ACC_COMP_DEFAULT : (d.acc_comp < 0.f ? 0.f : d.acc_comp);
    return g;
}

int launch_conv_simt(const ConvGeom& g, const EpiParams& epi, int block_n, cudaStream_t stream);
int launch_conv_tc(const ConvGeom& g, const EpiParams& epi, int block_n, int passes, cudaStream_t stream);
// split-K for the 4x4 / 8x8 layers (conv_tc.cu): scratch = max_splits partial buffers of N*OH*OW*Co floats; returns -1 when the
// layer is not worth splitting (the caller then takes the normal path)
int launch_conv_tc_splitk(const ConvGeom& g, const EpiParams& epi, int passes, float* scratch, int max_splits, cudaStream_t stream);
int launch_conv_halo(const ConvGeom& g, const EpiParams& epi, int block_n, int passes, cudaStream_t stream);
int launch_conv_pair(const ConvGeom& g, const EpiParams& epi, int passes, cudaStream_t stream);
bool conv_pair_supported(const ConvGeom& g);
bool conv_prefers_pair(const ConvGeom& g);
double conv_halo_efficiency(const ConvGeom& g);
bool conv_prefers_halo(const ConvGeom& g);

// torgb partial sums are produced per block of 32 output channels, independent of the GEMM tile width
constexpr int CONV_RGB_BLOCK = 32;

static inline int conv_block_n(int Co, int block_n) {
    if (block_n == 0) block_n = Co >= 256 ? 256 : (Co >= 128 ? 128 : 64);
    return block_n;
}

// vectors staged per tile by the tensor-core kernels: dcoef*wgain, bias, next_scale, rgb_w[0..2]*rgb_style
constexpr int CONV_STG_VECS = 6;

#ifdef __CUDACC__
// The per-(sample, channel) vectors of the fused epilogue (semantics: `shgan_epilogue` in include/shgan_b200.h) are
// staged in shared memory once per tile: the tile's final epilogue then runs on LDS broadcasts instead of ~20 dependent
// L2 round trips per 16 channels.  That matters because the epilogue warps also drain the TMEM accumulation chunks: while
// they are in a tile's final epilogue the MMA issuer can only run two chunks ahead.
//   stg[0][i] = (dcoef ? dcoef[n,o] : 1) * wgain     stg[1][i] = bias ? bias[o] : 0     stg[2][i] = next_scale ? .. : 1
//   stg[3+j][i] = rgb_w[j,o] * rgb_style[n,o]        (o = o_base + i)
template <int BN, int NTHREADS>
__device__ __forceinline__ void stage_epilogue_vectors(const EpiParams& p, float* stg, int n, int Co, int o_base, int et) {
    for (int i = et; i < BN; i += NTHREADS) {
        const int o = o_base + i;
        const long long no = (long long)n * Co + o;
        stg[i] = (p.dcoef ? __ldg(p.dcoef + no) : 1.f) * p.wgain;
        stg[BN + i] = p.bias ? __ldg(p.bias + o) : 0.f;
        stg[2 * BN + i] = p.next_scale ? __ldg(p.next_scale + no) : 1.f;
        if (p.rgb_w) {
            const float st = __ldg(p.rgb_style + no);
#pragma unroll
            for (int j = 0; j < 3; ++j) stg[(3 + j) * BN + i] = __ldg(p.rgb_w + (long long)j * Co + o) * st;
        }
    }
}

// fused epilogue on CH consecutive channels (tile-local index oi, global index o0) of output pixel (n,y,x); nz = the
// pixel's noise value already multiplied by the noise strength (0 when the layer has no noise)
template <int BN, int CH>
__device__ __forceinline__ void epilogue_apply_staged(const EpiParams& p, const float* stg, float* v, float nz, int Co, int o0, int oi,
                                                      float* rgb, long long pix) {
#pragma unroll
    for (int i = 0; i < CH; i += 4) {
        const float4 d = *reinterpret_cast<const float4*>(stg + oi + i);
        const float4 b = *reinterpret_cast<const float4*>(stg + BN + oi + i);
        v[i] = v[i] * d.x + nz + b.x; v[i + 1] = v[i + 1] * d.y + nz + b.y;
        v[i + 2] = v[i + 2] * d.z + nz + b.z; v[i + 3] = v[i + 3] * d.w + nz + b.w;
    }
    if (p.act) {
#pragma unroll
        for (int i = 0; i < CH; ++i) v[i] = lrelu_agc(v[i], p.act_alpha, p.act_gain, p.act_clamp);
    } else if (p.act_gain != 1.f) {
#pragma unroll
        for (int i = 0; i < CH; ++i) v[i] *= p.act_gain;
    }
    if (p.skip_hi) {
#pragma unroll
        for (int i = 0; i < CH; i += 8) {
            float s[8];
            load_planes8(p.skip_hi, p.skip_lo, pix * Co + o0 + i, s);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[i + j] += s[j];
        }
    }
    if (p.rgb_w) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
#pragma unroll
            for (int i = 0; i < CH; i += 4) {
                const float4 w = *reinterpret_cast<const float4*>(stg + (3 + j) * BN + oi + i);
                rgb[j] = fmaf(v[i], w.x, rgb[j]); rgb[j] = fmaf(v[i + 1], w.y, rgb[j]);
                rgb[j] = fmaf(v[i + 2], w.z, rgb[j]); rgb[j] = fmaf(v[i + 3], w.w, rgb[j]);
            }
        }
    }
    if (p.out_f32) {
#pragma unroll
        for (int i = 0; i < CH; i += 4)
            *reinterpret_cast<float4*>(p.out_f32 + pix * Co + o0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    }
    if (p.out_hi) {
#pragma unroll
        for (int i = 0; i < CH; i += 4) {
            const float4 sc = *reinterpret_cast<const float4*>(stg + 2 * BN + oi + i);
            v[i] *= sc.x; v[i + 1] *= sc.y; v[i + 2] *= sc.z; v[i + 3] *= sc.w;
        }
        if (CH % 16 == 0) {
#pragma unroll
            for (int i = 0; i < CH; i += 16) store_planes16(p.out_hi, p.out_lo, pix * Co + o0 + i, v + i);
        } else {
#pragma unroll
            for (int i = 0; i < CH; i += 8) store_planes8(p.out_hi, p.out_lo, pix * Co + o0 + i, v + i);
        }
    }
}


// raw-mode store of CH consecutive channels of output pixel (n,y,x) into the strided z tensor
template <int CH>
__device__ __forceinline__ void raw_store(const ConvGeom& g, const float* v, int n, int y, int x, int o0, long long z_off = 0) {
    const long long zi = z_off + (((long long)n * g.ZH + (y * g.zsy + g.zoy)) * g.ZW + (x * g.zsx + g.zox)) * g.Co + o0;
#pragma unroll
    for (int i = 0; i < CH; i += 4)
        *reinterpret_cast<float4*>(g.z + zi + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
}
#endif

}  // namespace shgan
