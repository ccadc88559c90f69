// Spectral Hint Unit forward (sm_100a), cuFFT-free.  Replaces SHU.forward, lib/model_zoo/shgan.py:312-336
// (torch.fft.rfftn -> row shift -> cat(re,im) -> conv0 1x1 + bias + ReLU -> heterogeneous_filter (:143-160)
// -> complex -> per band: crop, Gaussian mask, un-shift, torch.fft.irfftn).
//
// Three kernels, all spectra stay in shared memory / registers inside a kernel:
//   1. shu_rfft2_kernel   one CTA per (n,c) plane: radix-2 shared-memory FFT, two real rows packed into one
//                         complex transform, columns transformed in place, 1/(R*R) scaling ('forward' norm) and
//                         the DC-to-centre row shift folded into the store.  -> spec1 [N, 2C, R, R/2+1] (re | im)
//   2. channel mixing.  C == 32 (the released model): on the tensor cores -- the spectrum is re-laid as NHWC hi/lo planes
//                         (x R so that the 'forward'-normalised values sit in fp16's normal range) and the two 1x1
//                         convolutions run through the tcgen05 igemm (conv_tc.cu) with its fp32-class 3-pass scheme:
//                         conv0 + bias + ReLU as a 1-tap conv, the heterogeneous filter as a 6-"tap" conv whose taps all
//                         read the same pixel, use the 6 anchor filters df1[:, o*6+k] as weights and are blended per
//                         frequency bin by cw[k,bin] in the register-level accumulation (ConvGeom::chunk_scale).
//                         Any other C: shu_mix_kernel, per-frequency-bin fp32 FMA mixing for a tile of 32 bins: conv0 (2C x 2C) + bias + ReLU,
//                         then the heterogeneous filter out[o] = sum_k cw[k,bin] * sum_i t[i] * df1[i, o*6+k]
//                         with all weights resident in shared memory (broadcast 128-bit reads, 4 FMA per LDS).
//                         -> spec2 [N, 2C, R, R/2+1]
//   3. shu_irfft2_kernel  one CTA per (n,c,band): crop + Gaussian band mask + un-shift folded into the load,
//                         inverse column FFTs, Hermitian extension with the DC/Nyquist imaginary parts dropped
//                         (C2R semantics of pocketfft/cuFFT on non-Hermitian input), two rows per complex FFT.
#include "conv_common.cuh"

namespace shgan {

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ int brev(int i, int log2n) { return (int)(__brev((unsigned)i) >> (32 - log2n)); }

// `count` independent in-place radix-2 DIT FFTs of length L = 1 << log2L on bit-reversed input.
// element e of transform b lives at data[b * batch_stride + e * elem_stride].  tw[m] = exp(+2*pi*i*m/Lmax),
// sign = -1 forward / +1 inverse.  All threads of the block must call this.
__device__ void fft_smem(float2* data, int log2L, int count, int elem_stride, int batch_stride, float sign,
                         const float2* tw, int log2Lmax) {
    const int L = 1 << log2L;
    const int nb = count * (L >> 1);
    for (int s = 0; s < log2L; ++s) {
        const int half = 1 << s;
        const int tw_shift = log2Lmax - (s + 1);
        for (int t = threadIdx.x; t < nb; t += blockDim.x) {
            const int b = t >> (log2L - 1);
            const int u = t & ((L >> 1) - 1);
            const int j = u & (half - 1);
            const int i0 = ((u >> s) << (s + 1)) + j;
            float2* p0 = data + b * batch_stride + i0 * elem_stride;
            float2* p1 = p0 + half * elem_stride;
            float2 w = tw[j << tw_shift];
            w.y *= sign;
            const float2 a = *p0, v = cmul(*p1, w);
            *p0 = make_float2(a.x + v.x, a.y + v.y);
            *p1 = make_float2(a.x - v.x, a.y - v.y);
        }
        __syncthreads();
    }
}

__device__ __forceinline__ void fill_twiddles(float2* tw, int Lmax) {
    for (int m = threadIdx.x; m < Lmax / 2; m += blockDim.x) {
        float s, c;
        sincospif(2.0f * (float)m / (float)Lmax, &s, &c);
        tw[m] = make_float2(c, s);
    }
}

// ---- 1. forward rfft2 ------------------------------------------------------------------------
// grid (N*C), dynamic smem: rowbuf [R/2][R] + colbuf [R][Rh] + tw [R/2] float2
__global__ void __launch_bounds__(256)
shu_rfft2_kernel(const float* __restrict__ x, float* __restrict__ spec1, int C, int R, int log2R) {
    extern __shared__ float2 sm[];
    const int Rh = R / 2 + 1;
    float2* rowbuf = sm;
    float2* colbuf = rowbuf + (R / 2) * R;
    float2* tw = colbuf + R * Rh;
    const int n = blockIdx.x / C, c = blockIdx.x % C;
    const float* xp = x + (long long)blockIdx.x * R * R;
    fill_twiddles(tw, R);
    for (int i = threadIdx.x; i < (R / 2) * R; i += blockDim.x) {
        const int p = i / R, xx = i - p * R;
        rowbuf[p * R + brev(xx, log2R)] = make_float2(__ldg(xp + (2 * p) * R + xx), __ldg(xp + (2 * p + 1) * R + xx));
    }
    __syncthreads();
    fft_smem(rowbuf, log2R, R / 2, 1, R, -1.f, tw, log2R);
    // untangle the two real rows of each packed transform; store rows bit-reversed for the column pass
    for (int i = threadIdx.x; i < (R / 2) * Rh; i += blockDim.x) {
        const int p = i / Rh, k = i - p * Rh;
        const float2 z = rowbuf[p * R + k], zz = rowbuf[p * R + ((R - k) & (R - 1))];
        colbuf[brev(2 * p, log2R) * Rh + k] = make_float2(0.5f * (z.x + zz.x), 0.5f * (z.y - zz.y));
        colbuf[brev(2 * p + 1, log2R) * Rh + k] = make_float2(0.5f * (z.y + zz.y), -0.5f * (z.x - zz.x));
    }
    __syncthreads();
    fft_smem(colbuf, log2R, Rh, Rh, 1, -1.f, tw, log2R);
    // norm='forward' scaling and the row shift of shgan.py:315-317: out row j holds X[(j + R/2 + 1) mod R]
    const float sc = 1.f / ((float)R * (float)R);
    float* re = spec1 + ((long long)n * 2 * C + c) * R * Rh;
    float* im = spec1 + ((long long)n * 2 * C + C + c) * R * Rh;
    for (int i = threadIdx.x; i < R * Rh; i += blockDim.x) {
        const int j = i / Rh, k = i - j * Rh;
        const float2 v = colbuf[((j + R / 2 + 1) & (R - 1)) * Rh + k];
        re[i] = v.x * sc;
        im[i] = v.y * sc;
    }
}

// ---- 2. per-bin channel mixing -----------------------------------------------------------------
constexpr int MIX_TB = 32;  // bins per CTA

// grid (ceil(bins/32), N), 256 threads, dynamic smem: W0T [K2][K2] | b0 [K2] | df1 [K2][K2*6] | t [K2][32] | t0 [K2][32]
__global__ void __launch_bounds__(256)
shu_mix_kernel(const float* __restrict__ spec1, const float* __restrict__ conv0_w, const float* __restrict__ conv0_b,
               const float* __restrict__ df1_w, const float* __restrict__ cw, float* __restrict__ spec2, int K2, int bins) {
    extern __shared__ float smf[];
    float* W0T = smf;                     // [i][o]
    float* b0 = W0T + K2 * K2;
    float* df1 = b0 + K2;                 // [i][o*6+k]
    float* ts = df1 + K2 * K2 * 6;        // [i][b]
    float* t0s = ts + K2 * MIX_TB;        // [i][b]
    const int n = blockIdx.y, bin0 = blockIdx.x * MIX_TB;
    for (int i = threadIdx.x; i < K2 * K2; i += blockDim.x) {
        const int o = i / K2, ii = i - o * K2;
        W0T[ii * K2 + o] = __ldg(conv0_w + i);     // conv0.weight [o, i, 1, 1]
    }
    for (int i = threadIdx.x; i < K2; i += blockDim.x) b0[i] = __ldg(conv0_b + i);
    for (int i = threadIdx.x; i < K2 * K2 * 6 / 4; i += blockDim.x)
        reinterpret_cast<float4*>(df1)[i] = __ldg(reinterpret_cast<const float4*>(df1_w) + i);
    for (int i = threadIdx.x; i < K2 * MIX_TB; i += blockDim.x) {
        const int ch = i / MIX_TB, b = i - ch * MIX_TB;
        ts[i] = bin0 + b < bins ? __ldg(spec1 + ((long long)n * K2 + ch) * bins + bin0 + b) : 0.f;
    }
    __syncthreads();
    const int b = threadIdx.x & 31, og = threadIdx.x >> 5;  // bin, group of 8 output channels
    const bool active = og * 8 < K2;
    if (active) {
        float a[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = b0[og * 8 + j];
        for (int i = 0; i < K2; ++i) {
            const float t = ts[i * MIX_TB + b];
            const float4 w0 = *reinterpret_cast<const float4*>(W0T + i * K2 + og * 8);
            const float4 w1 = *reinterpret_cast<const float4*>(W0T + i * K2 + og * 8 + 4);
            a[0] = fmaf(t, w0.x, a[0]); a[1] = fmaf(t, w0.y, a[1]); a[2] = fmaf(t, w0.z, a[2]); a[3] = fmaf(t, w0.w, a[3]);
            a[4] = fmaf(t, w1.x, a[4]); a[5] = fmaf(t, w1.y, a[5]); a[6] = fmaf(t, w1.z, a[6]); a[7] = fmaf(t, w1.w, a[7]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) t0s[(og * 8 + j) * MIX_TB + b] = fmaxf(a[j], 0.f);  // ReLU, shgan.py:321
    }
    __syncthreads();
    if (active && bin0 + b < bins) {
        float acc[48];
#pragma unroll
        for (int j = 0; j < 48; ++j) acc[j] = 0.f;
        for (int i = 0; i < K2; ++i) {
            const float t = t0s[i * MIX_TB + b];
            const float4* wp = reinterpret_cast<const float4*>(df1 + i * K2 * 6 + og * 48);
#pragma unroll
            for (int q = 0; q < 12; ++q) {
                const float4 w = wp[q];
                acc[4 * q] = fmaf(t, w.x, acc[4 * q]);
                acc[4 * q + 1] = fmaf(t, w.y, acc[4 * q + 1]);
                acc[4 * q + 2] = fmaf(t, w.z, acc[4 * q + 2]);
                acc[4 * q + 3] = fmaf(t, w.w, acc[4 * q + 3]);
            }
        }
        float cwv[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) cwv[k] = __ldg(cw + (long long)k * bins + bin0 + b);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float o = 0.f;
#pragma unroll
            for (int k = 0; k < 6; ++k) o = fmaf(acc[j * 6 + k], cwv[k], o);
            spec2[((long long)n * K2 + og * 8 + j) * bins + bin0 + b] = o;
        }
    }
}

// ---- 3. per-band inverse rfft2 -----------------------------------------------------------------
struct ShuBands {
    float* out[8];
    int gauss_off[8];   // float offset of band k's mask inside `gauss`
    int num_bands, lowest_log2;
};

// grid (N*C, num_bands); dynamic smem sized for the largest band: colbuf [r][rh] + rowbuf [r/2][r] + tw [R/2]
__global__ void __launch_bounds__(256)
shu_irfft2_kernel(const float* __restrict__ spec2, const float* __restrict__ gauss, ShuBands bands, int C, int R, int log2R) {
    extern __shared__ float2 sm[];
    const int band = blockIdx.y;
    const int log2r = bands.lowest_log2 + band;
    const int r = 1 << log2r, rh = r / 2 + 1, Rh = R / 2 + 1;
    float2* colbuf = sm;
    float2* rowbuf = colbuf + r * rh;
    float2* tw = rowbuf + (r / 2) * r;
    const int n = blockIdx.x / C, c = blockIdx.x % C;
    const float* re = spec2 + ((long long)n * 2 * C + c) * R * Rh;
    const float* im = spec2 + ((long long)n * 2 * C + C + c) * R * Rh;
    const float* gm = gauss + bands.gauss_off[band];
    fill_twiddles(tw, r);
    // crop rows [R/2 - r/2, R/2 + r/2), cols [0, rh) (shgan.py:328), mask (:329), un-shift rows (:331-333):
    // un-shifted row j holds cropped row (j + r/2 - 1) mod r
    for (int i = threadIdx.x; i < r * rh; i += blockDim.x) {
        const int j = i / rh, k = i - j * rh;
        const int cj = (j + r / 2 - 1) & (r - 1);
        const int src = (R / 2 - r / 2 + cj) * Rh + k;
        const float g = __ldg(gm + cj * rh + k);
        colbuf[brev(j, log2r) * rh + k] = make_float2(__ldg(re + src) * g, __ldg(im + src) * g);
    }
    __syncthreads();
    fft_smem(colbuf, log2r, rh, rh, 1, +1.f, tw, log2r);
    // Hermitian extension along the last axis (imaginary parts of the DC and Nyquist bins dropped),
    // rows 2p and 2p+1 packed as real and imaginary part of one complex inverse transform
    for (int i = threadIdx.x; i < (r / 2) * r; i += blockDim.x) {
        const int p = i / r, k = i - p * r;
        const int kk = k <= r / 2 ? k : r - k;
        float2 ya = colbuf[(2 * p) * rh + kk], yb = colbuf[(2 * p + 1) * rh + kk];
        if (k > r / 2) { ya.y = -ya.y; yb.y = -yb.y; }
        if (k == 0 || k == r / 2) { ya.y = 0.f; yb.y = 0.f; }
        rowbuf[p * r + brev(k, log2r)] = make_float2(ya.x - yb.y, ya.y + yb.x);
    }
    __syncthreads();
    fft_smem(rowbuf, log2r, r / 2, 1, r, +1.f, tw, log2r);
    float* op = bands.out[band] + (long long)blockIdx.x * r * r;
    for (int i = threadIdx.x; i < r * r; i += blockDim.x) {
        const int j = i / r, xx = i - j * r;
        const float2 v = rowbuf[(j >> 1) * r + xx];
        op[i] = (j & 1) ? v.y : v.x;
    }
}

// ---- 2b. operand packing for the tensor-core mix (C == 32) -------------------------------------
// w0 hi/lo [64][64] = conv0.weight [o][i]; w1 hi/lo [6][64][64]: w1[k][o][i] = df1[i][o*6+k]; sc [N*64] = `scale`
__global__ void __launch_bounds__(256)
shu_pack_kernel(const float* __restrict__ conv0_w, const float* __restrict__ df1_w, __half* w0_hi, __half* w0_lo, __half* w1_hi,
                __half* w1_lo, float* sc, int n_sc, float scale) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    for (int i = gid; i < 64 * 64; i += stride) split_f32(__ldg(conv0_w + i), w0_hi[i], w0_lo[i]);
    for (int i = gid; i < 6 * 64 * 64; i += stride) {
        const int k = i / 4096, o = (i >> 6) & 63, ii = i & 63;
        split_f32(__ldg(df1_w + ii * 384 + o * 6 + k), w1_hi[i], w1_lo[i]);
    }
    for (int i = gid; i < n_sc; i += stride) sc[i] = scale;
}

static int shu_mix_tensor(const float* spec1, float* spec2, const float* conv0_w, const float* conv0_b, const float* df1_w,
                          const float* cw, uint8_t* ws, int N, int R, cudaStream_t stream) {
    const int Rh = R / 2 + 1;
    const long long px = (long long)N * R * Rh;
    // workspace: P1 hi | P1 lo | P2 hi | P2 lo (fp16 [N,R,Rh,64]) | Y fp32 [N,R,Rh,64] | w0 hi/lo | w1 hi/lo | sc [N*64]
    __half* p1_hi = (__half*)ws;
    __half* p1_lo = p1_hi + px * 64;
    __half* p2_hi = p1_lo + px * 64;
    __half* p2_lo = p2_hi + px * 64;
    float* y = (float*)(p2_lo + px * 64);
    __half* w0_hi = (__half*)(y + px * 64);
    __half* w0_lo = w0_hi + 4096;
    __half* w1_hi = w0_lo + 4096;
    __half* w1_lo = w1_hi + 6 * 4096;
    float* sc = (float*)(w1_lo + 6 * 4096);
    const float scale = (float)R;      // |X_forward-normalised| <= max|x| <= 256 (lrelu_agc clamp): x R stays below fp16 max
    shu_pack_kernel<<<32, 256, 0, stream>>>(conv0_w, df1_w, w0_hi, w0_lo, w1_hi, w1_lo, sc, N * 64, scale);
    SHGAN_LAUNCH_CHECK();
    if (int e = shgan_nchw_to_planes(spec1, nullptr, nullptr, sc, p1_hi, p1_lo, N, 64, R, Rh, 0, 64, stream)) return e;

    ConvGeom g{};
    g.num_src = 1; g.src_hi[0] = p1_hi; g.src_lo[0] = p1_lo; g.src_h[0] = R; g.src_w[0] = Rh;
    g.N = N; g.C = 64; g.Co = 64; g.w_hi = w0_hi; g.w_lo = w0_lo;
    g.ntaps = 1; g.tap_src[0] = 0; g.tap_dy[0] = 0; g.tap_dx[0] = 0; g.tap_w[0] = 0;
    g.OH = R; g.OW = Rh; g.mode = 0; g.chunk_scale = nullptr; g.acc_comp = SHGAN_ACC_COMP_DEFAULT;
    EpiParams e0{};
    e0.wgain = 1.f / scale; e0.bias = conv0_b; e0.act = 1; e0.act_alpha = 0.f; e0.act_gain = 1.f; e0.act_clamp = -1.f;   // ReLU
    e0.next_scale = sc;                                                                                                  // x R again
    e0.out_hi = p2_hi; e0.out_lo = p2_lo;
    if (int e = launch_conv_tc(g, e0, 0, 3, stream)) return e;

    g.src_hi[0] = p2_hi; g.src_lo[0] = p2_lo; g.w_hi = w1_hi; g.w_lo = w1_lo;
    g.ntaps = 6;
    for (int k = 0; k < 6; ++k) { g.tap_src[k] = 0; g.tap_dy[k] = 0; g.tap_dx[k] = 0; g.tap_w[k] = k; }
    g.chunk_scale = cw;                 // [6][R*Rh]
    EpiParams e1{};
    e1.wgain = 1.f / scale; e1.act = 0; e1.act_gain = 1.f; e1.act_clamp = -1.f; e1.out_f32 = y;
    if (int e = launch_conv_tc(g, e1, 0, 3, stream)) return e;
    return shgan_nhwc_to_nchw_f32(y, spec2, N, 64, R, Rh, stream);
}

static inline int log2_exact(int v) {
    int l = 0;
    while ((1 << l) < v) ++l;
    return (1 << l) == v ? l : -1;
}

}  // namespace shgan

using namespace shgan;

extern "C" int64_t shgan_shu_workspace_bytes(int N, int C, int R) {
    const int64_t bins = (int64_t)R * (R / 2 + 1);
    int64_t b = 2LL * N * 2 * C * bins * (int64_t)sizeof(float);                  // spec1, spec2
    if (C == 32) b += N * bins * 64 * (4 * 2 + 4) + 2 * 7 * 4096 * 2 + (int64_t)N * 64 * 4 + 256;   // tensor-core mix operands
    return b;
}

extern "C" int shgan_shu_fwd(const float* x, const float* conv0_w, const float* conv0_b, const float* df1_w, const float* cw,
                             const float* gauss, void* spec_ws, float* const* outs, int num_bands, int N, int C, int R,
                             int lowest_res, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SHGAN_CHECK(x && conv0_w && conv0_b && df1_w && cw && gauss && spec_ws && outs, "null pointer");
    const int log2R = log2_exact(R), log2low = log2_exact(lowest_res);
    SHGAN_CHECK(log2R >= 2 && R <= 128, "input_res must be a power of two in 4..128");
    SHGAN_CHECK(log2low >= 1 && lowest_res <= R, "lowest_res must be a power of two in 2..input_res");
    SHGAN_CHECK(num_bands == log2R - log2low + 1 && num_bands <= 8, "num_bands must be log2(input_res/lowest_res)+1");
    SHGAN_CHECK(C >= 4 && C <= 32 && C % 4 == 0, "C must be a multiple of 4 in 4..32");
    SHGAN_CHECK(N >= 0 && (long long)N * C <= INT32_MAX / (R * R), "bad batch size");
    if (N == 0) return 0;
    const int Rh = R / 2 + 1, K2 = 2 * C, bins = R * Rh;
    float* spec1 = (float*)spec_ws;
    float* spec2 = spec1 + (long long)N * K2 * bins;

    static DeviceInit once;
    if (int e = device_init(once, nullptr, []() -> int {
            SHGAN_CUDA(cudaFuncSetAttribute(shu_rfft2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
            SHGAN_CUDA(cudaFuncSetAttribute(shu_irfft2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
            SHGAN_CUDA(cudaFuncSetAttribute(shu_mix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
            return 0;
        })) return e;
    const size_t fft_smem_bytes = ((size_t)(R / 2) * R + (size_t)R * Rh + R / 2) * sizeof(float2);
    shu_rfft2_kernel<<<N * C, 256, fft_smem_bytes, stream>>>(x, spec1, C, R, log2R);
    SHGAN_LAUNCH_CHECK();

    if (C == 32) {
        uint8_t* mix_ws = (uint8_t*)(spec2 + (long long)N * K2 * bins);
        mix_ws = (uint8_t*)(((uintptr_t)mix_ws + 255) & ~(uintptr_t)255);
        if (int e = shu_mix_tensor(spec1, spec2, conv0_w, conv0_b, df1_w, cw, mix_ws, N, R, stream)) return e;
    } else {
        const size_t mix_smem = ((size_t)K2 * K2 + K2 + (size_t)K2 * K2 * 6 + 2 * (size_t)K2 * MIX_TB) * sizeof(float);
        dim3 mgrid(ceil_div(bins, MIX_TB), N);
        shu_mix_kernel<<<mgrid, 256, mix_smem, stream>>>(spec1, conv0_w, conv0_b, df1_w, cw, spec2, K2, bins);
        SHGAN_LAUNCH_CHECK();
    }

    ShuBands bands;
    bands.num_bands = num_bands;
    bands.lowest_log2 = log2low;
    int off = 0;
    for (int k = 0; k < num_bands; ++k) {
        const int r = lowest_res << k;
        SHGAN_CHECK(outs[k], "null output pointer");
        bands.out[k] = outs[k];
        bands.gauss_off[k] = off;
        off += r * (r / 2 + 1);
    }
    dim3 igrid(N * C, num_bands);
    shu_irfft2_kernel<<<igrid, 256, fft_smem_bytes, stream>>>(spec2, gauss, bands, C, R, log2R);
    SHGAN_LAUNCH_CHECK();
    return 0;
}
