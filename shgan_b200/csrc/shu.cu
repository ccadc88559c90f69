// Spectral Hint Unit forward (sm_100a), cuFFT-free.  Replaces SHU.forward, lib/model_zoo/shgan.py:312-336
// (torch.fft.rfftn -> row shift -> cat(re,im) -> conv0 1x1 + bias + ReLU -> heterogeneous_filter (:143-160)
// -> complex -> per band: crop, Gaussian mask, un-shift, torch.fft.irfftn).
//
// Three launches for input_res <= 128 (the released model: 64), spectra in shared memory / registers inside each:
//   1. forward rfft2.  input_res 64 (the released model): shu_fft64.cu -- register-resident radix-8 transforms, bulk-copied
//                         planes, spectrum written kx-major.  Other sizes: shu_rfft2_kernel, one CTA per (n,c) plane, radix-2
//                         shared-memory FFT, two real rows packed into one complex transform, 1/(R*R) scaling ('forward' norm)
//                         and the DC-to-centre row shift folded into the store.  -> spec1 [N, 2C, bins] (re | im)
//   2. channel mixing, ONE kernel.  C == 32 (the released model): shu_mix_tc.cu -- both 1x1 convolutions on the tcgen05
//                         tensor cores (conv0 + bias + ReLU, whose result never leaves shared memory, then the anchor filters
//                         of the heterogeneous filter blended per bin by cw in registers); the packed fp16 hi/lo weights are
//                         prepared ONCE by shgan_shu_pack (engine.refresh), not per forward.
//                         Any other C: shu_mix_kernel, per-bin fp32 FMA mixing with the weights in shared memory.
//                         -> spec2 [N, 2C, bins]
//   3. per-band inverse.  input_res 64: shu_fft64.cu, all bands of a plane in one CTA.  Other sizes: shu_irfft2_kernel, one CTA
//                         per (n,c,band): crop + Gaussian band mask + un-shift folded into the load,
//                         inverse column FFTs, Hermitian extension with the DC/Nyquist imaginary parts dropped
//                         (C2R semantics of pocketfft/cuFFT on non-Hermitian input), two rows per complex FFT.
// input_res 256 / 512 (BASELINE.json config C5 sweeps the unit up to 512): a plane no longer fits one SM's shared memory;
// the transforms then run as a row pass and a column pass through a float2 scratch in global memory
// (shu_fft_rows_kernel / shu_fft_cols_kernel and their inverses), same arithmetic, five launches more.
#include "shu_internal.cuh"

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
// grid ceil(N*C / P): P planes per CTA (small planes: more work per block barrier of the radix-2 stages);
// dynamic smem: rowbuf [P][R/2][R] + colbuf [R][P][Rh] (planes interleaved per row: the P * Rh column transforms of a CTA are
// then one uniformly strided batch) + tw [R/2] float2
template <bool MULTI>
__global__ void __launch_bounds__(1024)
shu_rfft2_kernel(const float* __restrict__ x, float* __restrict__ spec1, int NC, int C, int R, int log2R, int P_) {
    extern __shared__ float2 sm[];
    const int P = MULTI ? P_ : 1;          // MULTI == false: one plane per CTA, the plane index below folds to 0 at compile time
    const int Rh = R / 2 + 1, RR = (R / 2) * R, RC = R * Rh;
    float2* rowbuf = sm;
    float2* colbuf = rowbuf + P * RR;
    float2* tw = colbuf + P * RC;
    const int plane0 = blockIdx.x * P;
    const int np = NC - plane0 < P ? NC - plane0 : P;          // planes of this CTA
    const float* xp = x + (long long)plane0 * R * R;
    fill_twiddles(tw, R);
    for (int i = threadIdx.x; i < np * RR; i += blockDim.x) {
        const int pl = MULTI ? i / RR : 0, r = i - pl * RR, p = r / R, xx = r - p * R;
        const float* xq = xp + (long long)pl * R * R;
        rowbuf[pl * RR + p * R + brev(xx, log2R)] = make_float2(__ldg(xq + (2 * p) * R + xx), __ldg(xq + (2 * p + 1) * R + xx));
    }
    __syncthreads();
    fft_smem(rowbuf, log2R, np * (R / 2), 1, R, -1.f, tw, log2R);
    // untangle the two real rows of each packed transform; store rows bit-reversed for the column pass
    for (int i = threadIdx.x; i < np * (R / 2) * Rh; i += blockDim.x) {
        const int pl = MULTI ? i / ((R / 2) * Rh) : 0, r = i - pl * (R / 2) * Rh, p = r / Rh, k = r - p * Rh;
        const float2 z = rowbuf[pl * RR + p * R + k], zz = rowbuf[pl * RR + p * R + ((R - k) & (R - 1))];
        colbuf[(brev(2 * p, log2R) * P + pl) * Rh + k] = make_float2(0.5f * (z.x + zz.x), 0.5f * (z.y - zz.y));
        colbuf[(brev(2 * p + 1, log2R) * P + pl) * Rh + k] = make_float2(0.5f * (z.y + zz.y), -0.5f * (z.x - zz.x));
    }
    __syncthreads();
    fft_smem(colbuf, log2R, P * Rh, P * Rh, 1, -1.f, tw, log2R);
    // norm='forward' scaling and the row shift of shgan.py:315-317: out row j holds X[(j + R/2 + 1) mod R]
    const float sc = 1.f / ((float)R * (float)R);
    for (int i = threadIdx.x; i < np * RC; i += blockDim.x) {
        const int pl = MULTI ? i / RC : 0, r = i - pl * RC, j = r / Rh, k = r - j * Rh;
        const int plane = plane0 + pl, n = plane / C, c = plane - n * C;
        const float2 v = colbuf[(((j + R / 2 + 1) & (R - 1)) * P + pl) * Rh + k];
        spec1[((long long)n * 2 * C + c) * RC + r] = v.x * sc;
        spec1[((long long)n * 2 * C + C + c) * RC + r] = v.y * sc;
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

// grid (ceil(N*C / P), num_bands), P planes per CTA; dynamic smem sized for the largest band: colbuf [r][P][rh] + rowbuf [P][r/2][r] + tw [R/2]
template <bool MULTI>
__global__ void __launch_bounds__(1024)
shu_irfft2_kernel(const float* __restrict__ spec2, const float* __restrict__ gauss, ShuBands bands, int NC, int C, int R, int log2R, int P_,
                  int skip_upto) {
    extern __shared__ float2 sm[];
    const int P = MULTI ? P_ : 1;
    const int band = blockIdx.y;
    const int log2r = bands.lowest_log2 + band;
    const int r = 1 << log2r, rh = r / 2 + 1, Rh = R / 2 + 1, rc = r * rh, rr = (r / 2) * r;
    if (r > 128) return;      // large bands run as a column pass + a row pass through global memory
    if (r <= skip_upto) return;   // bands of at most 32 x 32 run in shu_small.cu (one thread per transform)
    float2* colbuf = sm;
    float2* rowbuf = colbuf + P * rc;
    float2* tw = rowbuf + P * rr;
    const int plane0 = blockIdx.x * P;
    const int np = NC - plane0 < P ? NC - plane0 : P;
    const float* gm = gauss + bands.gauss_off[band];
    fill_twiddles(tw, r);
    // crop rows [R/2 - r/2, R/2 + r/2), cols [0, rh) (shgan.py:328), mask (:329), un-shift rows (:331-333):
    // un-shifted row j holds cropped row (j + r/2 - 1) mod r
    for (int i = threadIdx.x; i < np * rc; i += blockDim.x) {
        const int pl = MULTI ? i / rc : 0, q = i - pl * rc, j = q / rh, k = q - j * rh;
        const int plane = plane0 + pl, n = plane / C, c = plane - n * C;
        const float* re = spec2 + ((long long)n * 2 * C + c) * R * Rh;
        const float* im = spec2 + ((long long)n * 2 * C + C + c) * R * Rh;
        const int cj = (j + r / 2 - 1) & (r - 1);
        const int src = (R / 2 - r / 2 + cj) * Rh + k;
        const float g = __ldg(gm + cj * rh + k);
        colbuf[(brev(j, log2r) * P + pl) * rh + k] = make_float2(__ldg(re + src) * g, __ldg(im + src) * g);
    }
    for (int i = threadIdx.x + np * rc; i < P * rc; i += blockDim.x) {      // planes past the end of the batch: zeros
        const int pl = i / rc, q = i - pl * rc;
        colbuf[((q / rh) * P + pl) * rh + (q % rh)] = make_float2(0.f, 0.f);
    }
    __syncthreads();
    fft_smem(colbuf, log2r, P * rh, P * rh, 1, +1.f, tw, log2r);
    // Hermitian extension along the last axis (imaginary parts of the DC and Nyquist bins dropped),
    // rows 2p and 2p+1 packed as real and imaginary part of one complex inverse transform
    for (int i = threadIdx.x; i < np * rr; i += blockDim.x) {
        const int pl = MULTI ? i / rr : 0, q = i - pl * rr, p = q / r, k = q - p * r;
        const int kk = k <= r / 2 ? k : r - k;
        float2 ya = colbuf[((2 * p) * P + pl) * rh + kk], yb = colbuf[((2 * p + 1) * P + pl) * rh + kk];
        if (k > r / 2) { ya.y = -ya.y; yb.y = -yb.y; }
        if (k == 0 || k == r / 2) { ya.y = 0.f; yb.y = 0.f; }
        rowbuf[pl * rr + p * r + brev(k, log2r)] = make_float2(ya.x - yb.y, ya.y + yb.x);
    }
    __syncthreads();
    fft_smem(rowbuf, log2r, np * (r / 2), 1, r, +1.f, tw, log2r);
    float* op = bands.out[band] + (long long)plane0 * r * r;
    for (int i = threadIdx.x; i < np * r * r; i += blockDim.x) {
        const int pl = MULTI ? i / (r * r) : 0, q = i - pl * r * r, j = q / r, xx = q - j * r;
        const float2 v = rowbuf[pl * rr + (j >> 1) * r + xx];
        op[i] = (j & 1) ? v.y : v.x;
    }
}

// ---- 1b / 3b. transforms through global memory for input_res (or band size) > 128 ------------------------------
constexpr int BIG_RP = 8;     // row pairs per CTA in the row passes
constexpr int BIG_CB = 8;     // columns per CTA in the column passes

// forward row pass: grid (N*C, R/2/BIG_RP); smem [BIG_RP][R] + tw[R/2] float2.  tmp[plane][row][k], k < Rh (natural order)
__global__ void __launch_bounds__(256)
shu_fft_rows_kernel(const float* __restrict__ x, float2* __restrict__ tmp, int R, int log2R) {
    extern __shared__ float2 sm[];
    const int Rh = R / 2 + 1;
    float2* rowbuf = sm;
    float2* tw = rowbuf + BIG_RP * R;
    const float* xp = x + (long long)blockIdx.x * R * R;
    const int p0 = blockIdx.y * BIG_RP;
    fill_twiddles(tw, R);
    for (int i = threadIdx.x; i < BIG_RP * R; i += blockDim.x) {
        const int pp = i / R, xx = i - pp * R;
        rowbuf[pp * R + brev(xx, log2R)] = make_float2(__ldg(xp + (long long)(2 * (p0 + pp)) * R + xx), __ldg(xp + (long long)(2 * (p0 + pp) + 1) * R + xx));
    }
    __syncthreads();
    fft_smem(rowbuf, log2R, BIG_RP, 1, R, -1.f, tw, log2R);
    float2* tp = tmp + (long long)blockIdx.x * R * Rh;
    for (int i = threadIdx.x; i < BIG_RP * Rh; i += blockDim.x) {
        const int pp = i / Rh, k = i - pp * Rh;
        const float2 z = rowbuf[pp * R + k], zz = rowbuf[pp * R + ((R - k) & (R - 1))];
        tp[(long long)(2 * (p0 + pp)) * Rh + k] = make_float2(0.5f * (z.x + zz.x), 0.5f * (z.y - zz.y));
        tp[(long long)(2 * (p0 + pp) + 1) * Rh + k] = make_float2(0.5f * (z.y + zz.y), -0.5f * (z.x - zz.x));
    }
}

// forward column pass: grid (N*C, ceil(Rh/BIG_CB)); smem [R][BIG_CB] + tw.  Writes spec1 (scaled, rows shifted) like shu_rfft2_kernel
__global__ void __launch_bounds__(256)
shu_fft_cols_kernel(const float2* __restrict__ tmp, float* __restrict__ spec1, int C, int R, int log2R) {
    extern __shared__ float2 sm[];
    const int Rh = R / 2 + 1;
    float2* colbuf = sm;
    float2* tw = colbuf + R * BIG_CB;
    const int n = blockIdx.x / C, c = blockIdx.x % C, k0 = blockIdx.y * BIG_CB;
    const float2* tp = tmp + (long long)blockIdx.x * R * Rh;
    fill_twiddles(tw, R);
    for (int i = threadIdx.x; i < R * BIG_CB; i += blockDim.x) {
        const int j = i / BIG_CB, kk = i - j * BIG_CB;
        colbuf[brev(j, log2R) * BIG_CB + kk] = k0 + kk < Rh ? tp[(long long)j * Rh + k0 + kk] : make_float2(0.f, 0.f);
    }
    __syncthreads();
    fft_smem(colbuf, log2R, BIG_CB, BIG_CB, 1, -1.f, tw, log2R);
    const float sc = 1.f / ((float)R * (float)R);
    float* re = spec1 + ((long long)n * 2 * C + c) * R * Rh;
    float* im = spec1 + ((long long)n * 2 * C + C + c) * R * Rh;
    for (int i = threadIdx.x; i < R * BIG_CB; i += blockDim.x) {
        const int j = i / BIG_CB, kk = i - j * BIG_CB;
        if (k0 + kk < Rh) {
            const float2 v = colbuf[((j + R / 2 + 1) & (R - 1)) * BIG_CB + kk];
            re[(long long)j * Rh + k0 + kk] = v.x * sc;
            im[(long long)j * Rh + k0 + kk] = v.y * sc;
        }
    }
}

// inverse column pass of one band of size r: grid (N*C, ceil(rh/BIG_CB)); crop + mask + un-shift folded into the load
__global__ void __launch_bounds__(256)
shu_ifft_cols_kernel(const float* __restrict__ spec2, const float* __restrict__ gm, float2* __restrict__ tmp, int C, int R, int r, int log2r) {
    extern __shared__ float2 sm[];
    const int rh = r / 2 + 1, Rh = R / 2 + 1;
    float2* colbuf = sm;
    float2* tw = colbuf + r * BIG_CB;
    const int n = blockIdx.x / C, c = blockIdx.x % C, k0 = blockIdx.y * BIG_CB;
    const float* re = spec2 + ((long long)n * 2 * C + c) * R * Rh;
    const float* im = spec2 + ((long long)n * 2 * C + C + c) * R * Rh;
    fill_twiddles(tw, r);
    for (int i = threadIdx.x; i < r * BIG_CB; i += blockDim.x) {
        const int j = i / BIG_CB, kk = i - j * BIG_CB, k = k0 + kk;
        float2 v = make_float2(0.f, 0.f);
        if (k < rh) {
            const int cj = (j + r / 2 - 1) & (r - 1);
            const long long src = (long long)(R / 2 - r / 2 + cj) * Rh + k;
            const float gq = __ldg(gm + (long long)cj * rh + k);
            v = make_float2(__ldg(re + src) * gq, __ldg(im + src) * gq);
        }
        colbuf[brev(j, log2r) * BIG_CB + kk] = v;
    }
    __syncthreads();
    fft_smem(colbuf, log2r, BIG_CB, BIG_CB, 1, +1.f, tw, log2r);
    float2* tp = tmp + (long long)blockIdx.x * r * rh;
    for (int i = threadIdx.x; i < r * BIG_CB; i += blockDim.x) {
        const int j = i / BIG_CB, kk = i - j * BIG_CB;
        if (k0 + kk < rh) tp[(long long)j * rh + k0 + kk] = colbuf[j * BIG_CB + kk];
    }
}

// inverse row pass: grid (N*C, r/2/BIG_RP); Hermitian extension, two rows per complex transform
__global__ void __launch_bounds__(256)
shu_ifft_rows_kernel(const float2* __restrict__ tmp, float* __restrict__ out, int r, int log2r) {
    extern __shared__ float2 sm[];
    const int rh = r / 2 + 1;
    float2* rowbuf = sm;
    float2* tw = rowbuf + BIG_RP * r;
    const float2* tp = tmp + (long long)blockIdx.x * r * rh;
    const int p0 = blockIdx.y * BIG_RP;
    fill_twiddles(tw, r);
    for (int i = threadIdx.x; i < BIG_RP * r; i += blockDim.x) {
        const int pp = i / r, k = i - pp * r;
        const int kk = k <= r / 2 ? k : r - k;
        float2 ya = tp[(long long)(2 * (p0 + pp)) * rh + kk], yb = tp[(long long)(2 * (p0 + pp) + 1) * rh + kk];
        if (k > r / 2) { ya.y = -ya.y; yb.y = -yb.y; }
        if (k == 0 || k == r / 2) { ya.y = 0.f; yb.y = 0.f; }
        rowbuf[pp * r + brev(k, log2r)] = make_float2(ya.x - yb.y, ya.y + yb.x);
    }
    __syncthreads();
    fft_smem(rowbuf, log2r, BIG_RP, 1, r, +1.f, tw, log2r);
    float* op = out + (long long)blockIdx.x * r * r;
    for (int i = threadIdx.x; i < BIG_RP * 2 * r; i += blockDim.x) {
        const int jj = i / r, xx = i - jj * r;          // jj: local output row 0 .. 2*BIG_RP-1
        const float2 v = rowbuf[(jj >> 1) * r + xx];
        op[(long long)(2 * p0 + jj) * r + xx] = (jj & 1) ? v.y : v.x;
    }
}

static inline int log2_exact(int v) {
    int l = 0;
    while ((1 << l) < v) ++l;
    return (1 << l) == v ? l : -1;
}

}  // namespace shgan

using namespace shgan;

extern "C" int64_t shgan_shu_packed_bytes(int C) { return C == 32 ? (int64_t)SHU_PACKED_BYTES : 16; }

namespace shgan { int launch_shu_pack_tc(const float* conv0_w, const float* df1_w, void* packed, cudaStream_t stream); }

extern "C" int shgan_shu_pack(const float* conv0_w, const float* df1_w, void* packed, int C, void* stream) {
    SHGAN_CHECK(conv0_w && df1_w && packed, "null pointer");
    if (C != 32) return 0;                 // only the tensor-core mix has packed operands
    SHGAN_CHECK(((uintptr_t)packed & 15) == 0, "packed buffer must be 16-byte aligned");
    return launch_shu_pack_tc(conv0_w, df1_w, packed, (cudaStream_t)stream);
}

extern "C" int64_t shgan_shu_workspace_bytes(int N, int C, int R) {
    const int64_t bins = (int64_t)R * (R / 2 + 1);
    int64_t b = 2LL * N * 2 * C * bins * (int64_t)sizeof(float) + 256;            // spec1, spec2
    if (R > 128) b += (int64_t)N * C * bins * (int64_t)sizeof(float2);            // row/column pass scratch
    if (C == 32) b += shgan_shu_packed_bytes(C) + 256;                            // on-the-fly weight packing when none is passed
    if (C == 32 && R == 64) b += 6 * bins * (int64_t)sizeof(float) + 256 + 8192;  // blend weights in the kx-major bin order (shu_fft64.cu) + trace
    return b;
}

extern "C" int shgan_shu_fwd(const float* x, const float* conv0_w, const float* conv0_b, const float* df1_w, const float* cw,
                             const float* gauss, const void* packed_w, void* spec_ws, float* const* outs, int num_bands, int N,
                             int C, int R, int lowest_res, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SHGAN_CHECK(x && conv0_w && conv0_b && df1_w && cw && gauss && spec_ws && outs, "null pointer");
    const int log2R = log2_exact(R), log2low = log2_exact(lowest_res);
    SHGAN_CHECK(log2R >= 2 && R <= 512, "input_res must be a power of two in 4..512");
    SHGAN_CHECK(log2low >= 1 && lowest_res <= R, "lowest_res must be a power of two in 2..input_res");
    SHGAN_CHECK(num_bands == log2R - log2low + 1 && num_bands <= 8, "num_bands must be log2(input_res/lowest_res)+1");
    SHGAN_CHECK(C >= 4 && C <= 32 && C % 4 == 0, "C must be a multiple of 4 in 4..32");
    SHGAN_CHECK(N >= 0 && (long long)N * C <= INT32_MAX / (R * R), "bad batch size");
    SHGAN_CHECK(((uintptr_t)spec_ws & 15) == 0, "workspace must be 16-byte aligned");
    if (N == 0) return 0;
    const int Rh = R / 2 + 1, K2 = 2 * C, bins = R * Rh;
    float* spec1 = (float*)spec_ws;
    float* spec2 = spec1 + (long long)N * K2 * bins;
    uint8_t* extra = (uint8_t*)(((uintptr_t)(spec2 + (long long)N * K2 * bins) + 255) & ~(uintptr_t)255);
    float2* tmp = (float2*)extra;                                  // only when R > 128
    if (R > 128) extra += (long long)N * C * bins * sizeof(float2);

    ShuBands bands;
    bands.num_bands = num_bands;
    bands.lowest_log2 = log2low;
    int off = 0;
    bool aligned = ((uintptr_t)x & 15) == 0;
    for (int k = 0; k < num_bands; ++k) {
        const int r = lowest_res << k;
        SHGAN_CHECK(outs[k], "null output pointer");
        bands.out[k] = outs[k];
        bands.gauss_off[k] = off;
        off += r * (r / 2 + 1);
        aligned = aligned && ((uintptr_t)outs[k] & 15) == 0;
    }
    // the released model's size: register-resident radix-8 transforms (shu_fft64.cu), spectra kx-major between the kernels;
    // needs 16-byte aligned planes for its bulk copies
    const bool fast64 = R == 64 && C == 32 && lowest_res >= 4 && aligned;

    static DeviceInit once;
    int num_sms = 148;
    if (int e = device_init(once, &num_sms, []() -> int {
            SHGAN_CUDA(cudaFuncSetAttribute(shu_rfft2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
            SHGAN_CUDA(cudaFuncSetAttribute(shu_irfft2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
            SHGAN_CUDA(cudaFuncSetAttribute(shu_rfft2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
            SHGAN_CUDA(cudaFuncSetAttribute(shu_irfft2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
            SHGAN_CUDA(cudaFuncSetAttribute(shu_mix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
            return 0;
        })) return e;
    const int Rs = R < 128 ? R : 128;       // size the single-CTA transforms see
    const size_t fft_smem_bytes = ((size_t)(Rs / 2) * Rs + (size_t)Rs * (Rs / 2 + 1) + Rs / 2) * sizeof(float2);
    // planes per CTA of the generic single-CTA transforms: small planes share a CTA (more work per block barrier of the radix-2
    // stages; measured at batch 4096: input_res 4 / 8 / 16 0.78 / 1.31 / 2.12 -> 0.17 / 0.30 / 0.94 ms, no gain at 32 and a loss at
    // 128 from the extra index arithmetic, hence R <= 16), as long as every SM still gets CTAs
    int fft_planes = 1;
    while (R <= 16 && fft_planes < 16 && (size_t)(2 * fft_planes) * fft_smem_bytes <= 40 * 1024 && (long long)N * C >= 2LL * num_sms * 2 * fft_planes) fft_planes *= 2;
    // a 128^2 plane takes 133 KB of shared memory: one CTA per SM, so it gets 1024 threads (8 warps per SM could not hide the
    // shared-memory latency of the radix-2 stages between their block barriers); 64^2: two CTAs of 512
    const int fft_threads = R >= 128 ? 1024 : (R >= 64 ? 512 : 256);
    float* cw_kx = (float*)(((uintptr_t)extra + SHU_PACKED_BYTES + 255) & ~(uintptr_t)255);   // fast64 only (workspace sized for it)
    if (fast64) {
        if (int e = launch_shu_rfft2_r64(x, spec1, cw, cw_kx, N, C, stream)) return e;
    } else if (R <= 32 && aligned) {
        if (int e = launch_shu_rfft2_small(x, spec1, N, C, R, stream)) return e;      // one thread per 4 ... 32-point transform
    } else if (R <= 128) {
        if (fft_planes > 1) shu_rfft2_kernel<true><<<ceil_div(N * C, fft_planes), fft_threads, fft_planes * fft_smem_bytes, stream>>>(x, spec1, N * C, C, R, log2R, fft_planes);
        else shu_rfft2_kernel<false><<<N * C, fft_threads, fft_smem_bytes, stream>>>(x, spec1, N * C, C, R, log2R, 1);
        SHGAN_LAUNCH_CHECK();
    } else {
        const size_t sm_rows = ((size_t)BIG_RP * R + R / 2) * sizeof(float2), sm_cols = ((size_t)R * BIG_CB + R / 2) * sizeof(float2);
        shu_fft_rows_kernel<<<dim3(N * C, R / 2 / BIG_RP), 256, sm_rows, stream>>>(x, tmp, R, log2R);
        SHGAN_LAUNCH_CHECK();
        shu_fft_cols_kernel<<<dim3(N * C, ceil_div(Rh, BIG_CB)), 256, sm_cols, stream>>>(tmp, spec1, C, R, log2R);
        SHGAN_LAUNCH_CHECK();
    }

    if (C == 32) {
        const void* packed = packed_w;
        if (!packed) {                       // op-level callers without a prepared weight set: pack into the workspace
            if (int e = shgan_shu_pack(conv0_w, df1_w, extra, C, stream)) return e;
            packed = extra;
        }
        // |X_forward-normalised| <= max|x| <= 256 (lrelu_agc clamp): x R stays far below fp16 max for R <= 128; larger R use 128
        if (int e = launch_shu_mix_tc(spec1, packed, conv0_b, fast64 ? cw_kx : cw, spec2, N, R, (float)(R < 128 ? R : 128), stream,
                                      fast64 ? (void*)(cw_kx + 6 * bins) : nullptr /* cycle trace (development aid), see shu_workspace_bytes */)) return e;
    } else {
        const size_t mix_smem = ((size_t)K2 * K2 + K2 + (size_t)K2 * K2 * 6 + 2 * (size_t)K2 * MIX_TB) * sizeof(float);
        dim3 mgrid(ceil_div(bins, MIX_TB), N);
        shu_mix_kernel<<<mgrid, 256, mix_smem, stream>>>(spec1, conv0_w, conv0_b, df1_w, cw, spec2, K2, bins);
        SHGAN_LAUNCH_CHECK();
    }

    if (fast64) return launch_shu_irfft2_r64(spec2, gauss, bands, N, C, stream);
    // bands of at most 32 x 32: one thread per transform (shu_small.cu), one launch per band
    int skip_upto = 0;
    if (aligned) {
        for (int k = 0; k < num_bands; ++k) {
            const int r = lowest_res << k;
            if (r > 32) break;
            if (int e = launch_shu_irfft2_small(spec2, gauss + bands.gauss_off[k], bands.out[k], N, C, R, r, stream)) return e;
            skip_upto = r;
        }
    }
    if (lowest_res <= 128 && (lowest_res << (num_bands - 1)) > skip_upto) {
        int small_bands = 0;
        while (small_bands < num_bands && (lowest_res << small_bands) <= 128) ++small_bands;
        dim3 igrid(ceil_div(N * C, fft_planes), small_bands);
        if (fft_planes > 1) shu_irfft2_kernel<true><<<igrid, fft_threads, fft_planes * fft_smem_bytes, stream>>>(spec2, gauss, bands, N * C, C, R, log2R, fft_planes, skip_upto);
        else shu_irfft2_kernel<false><<<igrid, fft_threads, fft_smem_bytes, stream>>>(spec2, gauss, bands, N * C, C, R, log2R, 1, skip_upto);
        SHGAN_LAUNCH_CHECK();
    }
    for (int k = 0; k < num_bands; ++k) {
        const int r = lowest_res << k;
        if (r <= 128) continue;
        const int rh = r / 2 + 1, log2r = log2low + k;
        const size_t sm_rows = ((size_t)BIG_RP * r + r / 2) * sizeof(float2), sm_cols = ((size_t)r * BIG_CB + r / 2) * sizeof(float2);
        shu_ifft_cols_kernel<<<dim3(N * C, ceil_div(rh, BIG_CB)), 256, sm_cols, stream>>>(spec2, gauss + bands.gauss_off[k], tmp, C, R, r, log2r);
        SHGAN_LAUNCH_CHECK();
        shu_ifft_rows_kernel<<<dim3(N * C, r / 2 / BIG_RP), 256, sm_rows, stream>>>(tmp, bands.out[k], r, log2r);
        SHGAN_LAUNCH_CHECK();
    }
    return 0;
}
