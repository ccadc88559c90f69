// Fully-connected layers and style preparation (sm_100a).
//
// Batch sizes on this path are tiny (B <= 32 per GPU), so every dense layer is a weight-streaming
// GEMV-like problem bounded by reading W once from HBM: one warp owns one output feature, streams
// its weight row with 128-bit loads and keeps up to 8 batch accumulators in registers; the input
// rows (a few hundred KB at most) are served by L1/L2; DU independent weight loads per lane keep enough bytes in flight
// to stream the row at HBM speed instead of paying one DRAM latency per 512 B.
//   shgan_dense_fwd            <- dense.forward (torch.addmm), lib/model_zoo/stylegan.py:87-98
//   shgan_normalize_2nd_moment <- normalize_2nd_moment, stylegan.py:343-344
//   shgan_style_prep           <- the style / demodulation arithmetic of modulated_conv2d, stylegan.py:145-155
#include "common.cuh"

namespace shgan {

constexpr int DCH = 512;  // input features per staged chunk: 4 x 128-bit weight loads in flight per lane

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block = 8 warps = 8 output features; the NB input rows of the current chunk are staged once per block in shared
// memory (coalesced), each warp streams its own weight-row chunk from HBM with 4 independent 128-bit loads per lane.
template <int NB>
__global__ void __launch_bounds__(256)
dense_kernel(const float* __restrict__ x0, long long x0_stride, int I0, const float* __restrict__ x1, long long x1_stride,
             const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ y, long long y_stride, int B,
             int I, int O, float wgain, float bgain, int act, float act_alpha, float act_gain, float act_clamp) {
    __shared__ __align__(16) float xs[NB][DCH];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int o = blockIdx.x * 8 + warp;
    const bool live = o < O;
    const float* wr = w + (long long)(live ? o : 0) * I;
    for (int b0 = 0; b0 < B; b0 += NB) {
        float acc[NB];
#pragma unroll
        for (int j = 0; j < NB; ++j) acc[j] = 0.f;
        for (int ic = 0; ic < I; ic += DCH) {
            float4 wv[DCH / 128];
#pragma unroll
            for (int u = 0; u < DCH / 128; ++u) {
                const int i = ic + u * 128 + lane * 4;
                wv[u] = (live && i < I) ? __ldg(reinterpret_cast<const float4*>(wr + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            __syncthreads();   // previous chunk fully consumed
            for (int q = threadIdx.x; q < NB * (DCH / 4); q += 256) {
                const int j = q / (DCH / 4), i = ic + (q % (DCH / 4)) * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (b0 + j < B && i < I) {
                    // I0 is a multiple of 4, so a quad never straddles the two concatenated inputs
                    const float* xp = i < I0 ? x0 + (long long)(b0 + j) * x0_stride + i
                                             : x1 + (long long)(b0 + j) * x1_stride + (i - I0);
                    v = __ldg(reinterpret_cast<const float4*>(xp));
                }
                *reinterpret_cast<float4*>(&xs[j][(q % (DCH / 4)) * 4]) = v;
            }
            __syncthreads();
#pragma unroll
            for (int u = 0; u < DCH / 128; ++u) {
#pragma unroll
                for (int j = 0; j < NB; ++j) {
                    const float4 xv = *reinterpret_cast<const float4*>(&xs[j][u * 128 + lane * 4]);
                    acc[j] = fmaf(xv.x, wv[u].x, acc[j]);
                    acc[j] = fmaf(xv.y, wv[u].y, acc[j]);
                    acc[j] = fmaf(xv.z, wv[u].z, acc[j]);
                    acc[j] = fmaf(xv.w, wv[u].w, acc[j]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < NB; ++j) acc[j] = warp_sum(acc[j]);
        if (lane == 0 && live) {
            const float bv = bias ? __ldg(bias + o) * bgain : 0.f;
#pragma unroll
            for (int j = 0; j < NB; ++j) {
                if (b0 + j < B) {
                    float v = acc[j] * wgain + bv;
                    if (act) v = lrelu_agc(v, act_alpha, act_gain, act_clamp);
                    y[(long long)(b0 + j) * y_stride + o] = v;
                }
            }
        }
    }
}

__global__ void __launch_bounds__(256)
normalize_2nd_moment_kernel(const float* __restrict__ z, float* __restrict__ y, int D) {
    __shared__ float red[8];
    const int b = blockIdx.x;
    float s = 0.f;
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
        const float v = z[(long long)b * D + i];
        s = fmaf(v, v, s);
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    float tot = 0.f;
    for (int i = 0; i < 8; ++i) tot += red[i];
    const float sc = rsqrtf(tot / (float)D + 1e-8f);
    for (int i = threadIdx.x; i < D; i += blockDim.x) y[(long long)b * D + i] = z[(long long)b * D + i] * sc;
}

// one block: s_hat = s * (demod ? rsqrt(mean over the WHOLE [N,Ci] tensor of s^2) : pre_scale)
__global__ void __launch_bounds__(1024)
style_scale_kernel(const float* __restrict__ styles, float* __restrict__ s_hat, int total, int demod, float pre_scale) {
    __shared__ float red[32];
    float sc = pre_scale;
    if (demod) {
        float s = 0.f;
        for (int i = threadIdx.x; i < total; i += blockDim.x) {
            const float v = styles[i];
            s = fmaf(v, v, s);
        }
        s = warp_sum(s);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
        __syncthreads();
        float tot = 0.f;
        for (int i = 0; i < 32; ++i) tot += red[i];
        sc = rsqrtf(tot / (float)total);
    }
    for (int i = threadIdx.x; i < total; i += blockDim.x) s_hat[i] = styles[i] * sc;
}

// dcoef[n,o] = rsqrt(sum_i s_hat[n,i]^2 * wsq[o,i] + 1e-8); one warp per (n,o)
__global__ void __launch_bounds__(256)
dcoef_kernel(const float* __restrict__ s_hat, const float* __restrict__ wsq, float* __restrict__ dcoef, int N, int Ci, int Co) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long idx = (long long)blockIdx.x * 8 + warp;
    if (idx >= (long long)N * Co) return;
    const int n = (int)(idx / Co), o = (int)(idx % Co);
    float s = 0.f;
    for (int i = lane; i < Ci; i += 32) {
        const float v = __ldg(s_hat + (long long)n * Ci + i);
        s = fmaf(v * v, __ldg(wsq + (long long)o * Ci + i), s);
    }
    s = warp_sum(s);
    if (lane == 0) dcoef[idx] = rsqrtf(s + 1e-8f);
}

// All style layers of a forward in ONE launch (shgan_style_prep_batched): the per-layer pair of launches above costs
// ~40 launches of a few microseconds each per generator call.  Block (layer l, chunk c) recomputes the layer's
// batch-global scale from the (L2-resident) raw styles -- every block of a layer reduces in the same order, so they all
// obtain the same bits -- writes its slice of s_hat and 64 demodulation coefficients.
constexpr int SB_THREADS = 256;
constexpr int SB_OUT_PER_BLOCK = 64;    // (n, o) pairs per block: 8 warps x 8

__global__ void __launch_bounds__(SB_THREADS)
style_prep_batched_kernel(const float* __restrict__ raw, long long raw_stride, int N, const shgan_style_batch tb) {
    __shared__ float red[SB_THREADS / 32];
    __shared__ float s_sc;
    int l = 0;
    while (l + 1 < tb.num_layers && (int)blockIdx.x >= tb.block_start[l + 1]) ++l;
    const int chunk = blockIdx.x - tb.block_start[l];
    const int nchunks = tb.block_start[l + 1] - tb.block_start[l];
    const int Ci = tb.ci[l], Co = tb.co[l], total = N * Ci;
    const float* rl = raw + tb.offset[l];
    float sc = tb.pre_scale[l];
    if (tb.demod[l]) {
        float s = 0.f;
        for (int i = threadIdx.x; i < total; i += SB_THREADS) {
            const float v = __ldg(rl + (long long)(i / Ci) * raw_stride + (i % Ci));
            s = fmaf(v, v, s);
        }
        s = warp_sum(s);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            float tot = 0.f;
            for (int i = 0; i < SB_THREADS / 32; ++i) tot += red[i];
            s_sc = rsqrtf(tot / (float)total);
        }
        __syncthreads();
        sc = s_sc;
    }
    float* sh = (float*)tb.s_hat[l];
    for (int i = chunk * SB_THREADS + threadIdx.x; i < total; i += nchunks * SB_THREADS)
        sh[i] = __ldg(rl + (long long)(i / Ci) * raw_stride + (i % Ci)) * sc;
    if (!tb.demod[l]) return;
    const float* wsq = (const float*)tb.wsq[l];
    float* dc = (float*)tb.dcoef[l];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int k = 0; k < SB_OUT_PER_BLOCK / 8; ++k) {
        const long long idx = (long long)chunk * SB_OUT_PER_BLOCK + k * 8 + warp;
        if (idx >= (long long)N * Co) break;
        const int n = (int)(idx / Co), o = (int)(idx % Co);
        float s = 0.f;
        for (int i = lane; i < Ci; i += 32) {
            const float v = __ldg(rl + (long long)n * raw_stride + i) * sc;
            s = fmaf(v * v, __ldg(wsq + (long long)o * Ci + i), s);
        }
        s = warp_sum(s);
        if (lane == 0) dc[idx] = rsqrtf(s + 1e-8f);
    }
}

}  // namespace shgan

using namespace shgan;

extern "C" int shgan_style_prep_batched(const float* raw, int64_t raw_stride, int N, const shgan_style_batch* tb_, void* stream) {
    SHGAN_CHECK(raw && tb_, "null pointer");
    SHGAN_CHECK(tb_->num_layers >= 1 && tb_->num_layers <= SHGAN_MAX_STYLE_LAYERS, "num_layers out of range");
    if (N == 0) return 0;
    shgan_style_batch tb = *tb_;
    int blocks = 0;
    for (int l = 0; l < tb.num_layers; ++l) {
        SHGAN_CHECK(tb.ci[l] >= 1 && tb.s_hat[l], "bad layer description");
        SHGAN_CHECK(!tb.demod[l] || (tb.wsq[l] && tb.dcoef[l] && tb.co[l] >= 1), "demodulation needs wsq and dcoef");
        tb.block_start[l] = blocks;
        const int by_out = tb.demod[l] ? ceil_div(N * tb.co[l], SB_OUT_PER_BLOCK) : 1;
        const int by_in = ceil_div(N * tb.ci[l], SB_THREADS * 4);
        blocks += by_out > by_in ? by_out : by_in;
    }
    tb.block_start[tb.num_layers] = blocks;
    style_prep_batched_kernel<<<blocks, SB_THREADS, 0, (cudaStream_t)stream>>>(raw, raw_stride, N, tb);
    SHGAN_LAUNCH_CHECK();
    return 0;
}

extern "C" int shgan_dense_fwd(const float* x0, int64_t x0_stride, int I0, const float* x1, int64_t x1_stride,
                               const float* w, const float* bias, float* y, int64_t y_stride, int B, int I, int O,
                               float wgain, float bgain, int act, float act_alpha, float act_gain, float act_clamp,
                               void* stream) {
    SHGAN_CHECK(x0 && w && y, "null pointer");
    SHGAN_CHECK(B >= 0 && I >= 4 && O >= 1, "bad sizes");
    SHGAN_CHECK(I % 4 == 0 && I0 % 4 == 0 && x0_stride % 4 == 0 && x1_stride % 4 == 0, "feature counts/strides must be multiples of 4");
    SHGAN_CHECK(I0 >= 0 && I0 <= I && (I0 == I || x1), "second input missing");
    if (B == 0) return 0;
    if (B <= 8)
        dense_kernel<8><<<ceil_div(O, 8), 256, 0, (cudaStream_t)stream>>>(x0, x0_stride, I0, x1, x1_stride, w, bias, y, y_stride,
                                                                          B, I, O, wgain, bgain, act, act_alpha, act_gain, act_clamp);
    else
        dense_kernel<16><<<ceil_div(O, 8), 256, 0, (cudaStream_t)stream>>>(x0, x0_stride, I0, x1, x1_stride, w, bias, y, y_stride,
                                                                           B, I, O, wgain, bgain, act, act_alpha, act_gain, act_clamp);
    SHGAN_LAUNCH_CHECK();
    return 0;
}

extern "C" int shgan_normalize_2nd_moment(const float* z, float* y, int B, int D, void* stream) {
    SHGAN_CHECK(z && y && B >= 0 && D >= 1, "bad arguments");
    if (B == 0) return 0;
    normalize_2nd_moment_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(z, y, D);
    SHGAN_LAUNCH_CHECK();
    return 0;
}

extern "C" int shgan_style_prep(const float* styles, const float* wsq, float* s_hat, float* dcoef, int N, int Ci, int Co,
                                int demod, float pre_scale, void* stream) {
    SHGAN_CHECK(styles && s_hat, "null pointer");
    SHGAN_CHECK(N >= 0 && Ci >= 1 && Co >= 1, "bad sizes");
    SHGAN_CHECK(!demod || (wsq && dcoef), "demodulation needs wsq and dcoef");
    if (N == 0) return 0;
    style_scale_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(styles, s_hat, N * Ci, demod, pre_scale);
    SHGAN_LAUNCH_CHECK();
    if (demod) {
        dcoef_kernel<<<(unsigned)ceil_div64((long long)N * Co, 8), 256, 0, (cudaStream_t)stream>>>(s_hat, wsq, dcoef, N, Ci, Co);
        SHGAN_LAUNCH_CHECK();
    }
    return 0;
}
