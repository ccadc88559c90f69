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

// barrier over all threads of the thread-block cluster (release / acquire: shared-memory writes before it are visible to
// distributed-shared-memory reads after it)
__device__ __forceinline__ void dense_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// reads a float at the same shared-memory offset as `p` in the block of rank `rank` of this cluster
__device__ __forceinline__ float dense_ld_dsmem_f32(const float* p, int rank) {
    uint32_t remote;
    float v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"((uint32_t)__cvta_generic_to_shared(p)), "r"(rank));
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(remote) : "memory");
    return v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block = 8 warps x 2 output features each over ONE slice of the input features; the KS blocks of a thread-block cluster
// share the same 16 outputs and split I between them.  The NB input rows of the current 512-feature chunk are staged
// once per block in shared memory (coalesced); each lane keeps 2 rows x 4 x 128-bit weight loads in flight and requests
// the NEXT chunk's weights before it consumes the current one, so a block always has 32 KB of HBM requests outstanding
// (measured before: one row per warp, loads issued only after the previous chunk's FMAs, <= 148 blocks for the
// 8192 -> 1024 layer: 0.4 TB/s).  Two rows per warp halve the shared-memory reads per weight byte (the x fragment is
// reused from registers).  The cluster's partial sums are reduced in a fixed order by its first block through
// distributed shared memory: deterministic, no workspace, one launch.
constexpr int DROWS = 16;     // output features per block

template <int NB>
__global__ void __launch_bounds__(256, 2)
dense_kernel(const float* __restrict__ x0, long long x0_stride, int I0, const float* __restrict__ x1, long long x1_stride,
             const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ y, long long y_stride, int B,
             int I, int O, float wgain, float bgain, int act, float act_alpha, float act_gain, float act_clamp, int kslice, int KS) {
    __shared__ __align__(16) float xs[NB][DCH];
    __shared__ float part[DROWS][NB];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int o = blockIdx.y * DROWS + warp * 2;
    const bool live0 = o < O, live1 = o + 1 < O;
    const int kbeg = blockIdx.x * kslice, kend = min(I, kbeg + kslice);      // gridDim.x == KS == the cluster size
    const float* wr0 = w + (long long)(live0 ? o : 0) * I;
    const float* wr1 = w + (long long)(live1 ? o + 1 : 0) * I;
    auto load_w = [&](int ic, float4 (&a)[DCH / 128], float4 (&b)[DCH / 128]) {
#pragma unroll
        for (int u = 0; u < DCH / 128; ++u) {
            const int i = ic + u * 128 + lane * 4;
            const bool in = i < kend;
            a[u] = (live0 && in) ? __ldg(reinterpret_cast<const float4*>(wr0 + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
            b[u] = (live1 && in) ? __ldg(reinterpret_cast<const float4*>(wr1 + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    for (int b0 = 0; b0 < B; b0 += NB) {
        float acc0[NB], acc1[NB];
#pragma unroll
        for (int j = 0; j < NB; ++j) { acc0[j] = 0.f; acc1[j] = 0.f; }
        float4 wa[DCH / 128], wb[DCH / 128], na[DCH / 128], nb[DCH / 128];
        if (kbeg < kend) load_w(kbeg, wa, wb);
        for (int ic = kbeg; ic < kend; ic += DCH) {
            if (ic + DCH < kend) load_w(ic + DCH, na, nb);
            __syncthreads();   // previous chunk fully consumed
            {
                // all of a thread's input quads are requested before the first one is stored (one L2 round trip per chunk,
                // not NB / 2 dependent ones)
                constexpr int XQ = NB * (DCH / 4) / 256;
                float4 xv[XQ];
#pragma unroll
                for (int t = 0; t < XQ; ++t) {
                    const int q = threadIdx.x + t * 256;
                    const int j = q / (DCH / 4), i = ic + (q % (DCH / 4)) * 4;
                    xv[t] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (b0 + j < B && i < kend) {
                        // I0 is a multiple of 4, so a quad never straddles the two concatenated inputs
                        const float* xp = i < I0 ? x0 + (long long)(b0 + j) * x0_stride + i
                                                 : x1 + (long long)(b0 + j) * x1_stride + (i - I0);
                        xv[t] = __ldg(reinterpret_cast<const float4*>(xp));
                    }
                }
#pragma unroll
                for (int t = 0; t < XQ; ++t) {
                    const int q = threadIdx.x + t * 256;
                    *reinterpret_cast<float4*>(&xs[q / (DCH / 4)][(q % (DCH / 4)) * 4]) = xv[t];
                }
            }
            __syncthreads();
#pragma unroll
            for (int u = 0; u < DCH / 128; ++u) {
#pragma unroll
                for (int j = 0; j < NB; ++j) {
                    const float4 xv = *reinterpret_cast<const float4*>(&xs[j][u * 128 + lane * 4]);
                    acc0[j] = fmaf(xv.x, wa[u].x, acc0[j]); acc1[j] = fmaf(xv.x, wb[u].x, acc1[j]);
                    acc0[j] = fmaf(xv.y, wa[u].y, acc0[j]); acc1[j] = fmaf(xv.y, wb[u].y, acc1[j]);
                    acc0[j] = fmaf(xv.z, wa[u].z, acc0[j]); acc1[j] = fmaf(xv.z, wb[u].z, acc1[j]);
                    acc0[j] = fmaf(xv.w, wa[u].w, acc0[j]); acc1[j] = fmaf(xv.w, wb[u].w, acc1[j]);
                }
            }
#pragma unroll
            for (int u = 0; u < DCH / 128; ++u) { wa[u] = na[u]; wb[u] = nb[u]; }
        }
#pragma unroll
        for (int j = 0; j < NB; ++j) { acc0[j] = warp_sum(acc0[j]); acc1[j] = warp_sum(acc1[j]); }
        __syncthreads();        // (the previous batch group's partials have been read)
        if (lane == 0) {
#pragma unroll
            for (int j = 0; j < NB; ++j) { part[warp * 2][j] = acc0[j]; part[warp * 2 + 1][j] = acc1[j]; }
        }
        if (KS > 1) dense_cluster_sync(); else __syncthreads();
        if (blockIdx.x == 0 && threadIdx.x < DROWS * NB) {
            const int rr = threadIdx.x / NB, j = threadIdx.x % NB, oo = blockIdx.y * DROWS + rr;
            float v = part[rr][j];
            for (int r = 1; r < KS; ++r) v += dense_ld_dsmem_f32(&part[rr][j], r);      // fixed order: deterministic
            if (oo < O && b0 + j < B) {
                v = v * wgain + (bias ? __ldg(bias + oo) * bgain : 0.f);
                if (act) v = lrelu_agc(v, act_alpha, act_gain, act_clamp);
                y[(long long)(b0 + j) * y_stride + oo] = v;
            }
        }
        if (KS > 1) dense_cluster_sync();     // the peers' partials stay alive until the first block has read them
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
    // split I over a cluster of KS blocks until the grid fills the GPU twice over (slices stay multiples of 128 features:
    // one 128-bit weight load per lane)
    const int row_blocks = ceil_div(O, DROWS);
    int KS = 1;
    while (KS < 8 && row_blocks * KS < 2 * 148 && I % (KS * 2 * 128) == 0 && I / (KS * 2) >= 128) KS *= 2;
    const int kslice = I / KS;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)KS, (unsigned)row_blocks, 1);
    cfg.blockDim = dim3(256, 1, 1);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)KS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = KS > 1 ? 1 : 0;
    const long long x0s = x0_stride, x1s = x1_stride, ys = y_stride;
    if (B <= 8)
        SHGAN_CUDA(cudaLaunchKernelEx(&cfg, dense_kernel<8>, x0, x0s, I0, x1, x1s, w, bias, y, ys, B, I, O, wgain, bgain, act, act_alpha,
                                      act_gain, act_clamp, kslice, KS));
    else
        SHGAN_CUDA(cudaLaunchKernelEx(&cfg, dense_kernel<16>, x0, x0s, I0, x1, x1s, w, bias, y, ys, B, I, O, wgain, bgain, act, act_alpha,
                                      act_gain, act_clamp, kslice, KS));
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
