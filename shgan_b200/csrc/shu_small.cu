// Spectral Hint Unit, transforms of 4 ... 32 points (sm_100a): one THREAD per transform, data in registers.
//
// The generic shared-memory radix-2 kernels of shu.cu spend ~20 instructions and a block barrier per butterfly stage; for
// planes of up to 32 x 32 a whole row (or column) transform fits the registers of one thread (32 complex values), so a
// transform is a fully unrolled butterfly network with compile-time twiddles and the only exchange is the row/column corner
// turn through shared memory (one block barrier per plane set).  Used for input_res <= 32 (forward, BASELINE.json config C5
// sweeps the unit from 4 to 512) and for every band of at most 32 x 32 of the inverse when the input_res-64 kernels of
// shu_fft64.cu do not apply.  Same arithmetic and data layout as shu.cu (shgan.py:312-336):
//   forward: rfft2(norm='forward') of [N,C,R,R], two real rows packed per complex transform, spectrum rows shifted so that
//            output row j holds frequency (j + R/2 + 1) mod R, written as [N, 2C, R, Rh] (real planes, then imaginary planes);
//   inverse: crop rows [R/2 - r/2, R/2 + r/2) x cols [0, rh), Gaussian mask, un-shift, irfft2 (imaginary parts of the DC and
//            Nyquist bins dropped), rows 2p / 2p+1 recovered as real / imaginary part of one complex inverse transform.
#include "shu_internal.cuh"

namespace shgan {

namespace {

// exp(2*pi*i*m/32), m = 0..15
__device__ constexpr float TW_C[16] = {1.000000000e+00f, 9.807852804e-01f, 9.238795325e-01f, 8.314696123e-01f, 7.071067812e-01f, 5.555702330e-01f, 3.826834324e-01f, 1.950903220e-01f, 6.123233996e-17f, -1.950903220e-01f, -3.826834324e-01f, -5.555702330e-01f, -7.071067812e-01f, -8.314696123e-01f, -9.238795325e-01f, -9.807852804e-01f};
__device__ constexpr float TW_S[16] = {0.000000000e+00f, 1.950903220e-01f, 3.826834324e-01f, 5.555702330e-01f, 7.071067812e-01f, 8.314696123e-01f, 9.238795325e-01f, 9.807852804e-01f, 1.000000000e+00f, 9.807852804e-01f, 9.238795325e-01f, 8.314696123e-01f, 7.071067812e-01f, 5.555702330e-01f, 3.826834324e-01f, 1.950903220e-01f};

__host__ __device__ constexpr int brev_c(int i, int log2n) {
    int r = 0;
    for (int b = 0; b < log2n; ++b) r |= ((i >> b) & 1) << (log2n - 1 - b);
    return r;
}
__host__ __device__ constexpr int ilog2_c(int n) { return n <= 1 ? 0 : 1 + ilog2_c(n >> 1); }

// in-place complex FFT of length L (4 ... 32) on registers, natural order in and out; SIGN = -1 forward, +1 inverse (unscaled).
// Every index below is a compile-time constant after unrolling: the permutation is register renaming, the twiddles immediates.
template <int L, int SIGN>
__device__ __forceinline__ void fft_reg(float2 (&v)[L]) {
    constexpr int LOG2 = ilog2_c(L);
    float2 t[L];
#pragma unroll
    for (int i = 0; i < L; ++i) t[i] = v[brev_c(i, LOG2)];
#pragma unroll
    for (int s = 0; s < LOG2; ++s) {
        const int half = 1 << s;
#pragma unroll
        for (int u = 0; u < L / 2; ++u) {
            const int j = u & (half - 1);
            const int i0 = ((u >> s) << (s + 1)) + j, i1 = i0 + half;
            const int m = j * (32 >> (s + 1));             // twiddle exp(SIGN * 2*pi*i * j / (2*half)) = exp(SIGN * 2*pi*i * m / 32)
            float2 w;
            if (m == 0) {
                w = t[i1];
            } else if (m == 8) {                           // multiply by SIGN * i
                w = SIGN > 0 ? make_float2(-t[i1].y, t[i1].x) : make_float2(t[i1].y, -t[i1].x);
            } else {
                const float wc = TW_C[m], ws = SIGN > 0 ? TW_S[m] : -TW_S[m];
                w = make_float2(fmaf(t[i1].x, wc, -t[i1].y * ws), fmaf(t[i1].x, ws, t[i1].y * wc));
            }
            const float2 a = t[i0];
            t[i0] = make_float2(a.x + w.x, a.y + w.y);
            t[i1] = make_float2(a.x - w.x, a.y - w.y);
        }
    }
#pragma unroll
    for (int i = 0; i < L; ++i) v[i] = t[i];
}

}  // namespace

// ---- forward: grid ceil(N*C / PP), 256 threads, dynamic smem PP * R * Rh float2 ---------------------------------------
template <int R>
__global__ void __launch_bounds__(256)
shu_rfft2_small_kernel(const float* __restrict__ x, float* __restrict__ spec1, int NC, int C, int PP) {
    constexpr int Rh = R / 2 + 1, RC = R * Rh;
    extern __shared__ float2 sm_small[];
    const int plane0 = blockIdx.x * PP;
    const int np = NC - plane0 < PP ? NC - plane0 : PP;
    // rows: thread = (plane, row pair)
    {
        const int pl = threadIdx.x / (R / 2), p = threadIdx.x - pl * (R / 2);
        if (pl < np) {
            const float* xr = x + ((long long)(plane0 + pl) * R + 2 * p) * R;
            float2 v[R];
#pragma unroll
            for (int q = 0; q < R / 4; ++q) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(xr) + q), b = __ldg(reinterpret_cast<const float4*>(xr + R) + q);
                v[4 * q] = make_float2(a.x, b.x); v[4 * q + 1] = make_float2(a.y, b.y);
                v[4 * q + 2] = make_float2(a.z, b.z); v[4 * q + 3] = make_float2(a.w, b.w);
            }
            fft_reg<R, -1>(v);
            // untangle the two real rows of the packed transform
            float2* c0 = sm_small + (pl * R + 2 * p) * Rh;
#pragma unroll
            for (int k = 0; k < Rh; ++k) {
                const float2 z = v[k], zz = v[(R - k) & (R - 1)];
                c0[k] = make_float2(0.5f * (z.x + zz.x), 0.5f * (z.y - zz.y));
                c0[Rh + k] = make_float2(0.5f * (z.y + zz.y), -0.5f * (z.x - zz.x));
            }
        }
    }
    __syncthreads();
    // columns: thread = (plane, kx); norm='forward' scaling and the row shift of shgan.py:315-317 on the way out
    {
        const int pl = threadIdx.x / Rh, k = threadIdx.x - pl * Rh;
        if (pl < np) {
            float2 v[R];
            const float2* cb = sm_small + pl * RC + k;
#pragma unroll
            for (int j = 0; j < R; ++j) v[j] = cb[j * Rh];
            fft_reg<R, -1>(v);
            const int plane = plane0 + pl, n = plane / C, c = plane - n * C;
            float* re = spec1 + ((long long)n * 2 * C + c) * RC + k;
            float* im = re + (long long)C * RC;
            const float sc = 1.f / (float)(R * R);
#pragma unroll
            for (int j = 0; j < R; ++j) {
                const float2 o = v[(j + R / 2 + 1) & (R - 1)];
                re[j * Rh] = o.x * sc;
                im[j * Rh] = o.y * sc;
            }
        }
    }
}

// ---- inverse of one band of r x r (r <= 32): grid ceil(N*C / PP), 256 threads, dynamic smem PP * r * rh float2 ----------
template <int r>
__global__ void __launch_bounds__(256)
shu_irfft2_small_kernel(const float* __restrict__ spec2, const float* __restrict__ gm, float* __restrict__ out, int NC, int C, int R, int PP) {
    constexpr int rh = r / 2 + 1, rc = r * rh;
    extern __shared__ float2 sm_small[];
    const int Rh = R / 2 + 1;
    const int plane0 = blockIdx.x * PP;
    const int np = NC - plane0 < PP ? NC - plane0 : PP;
    // columns: thread = (plane, kx): crop rows [R/2 - r/2, R/2 + r/2) x cols [0, rh) (shgan.py:328), mask (:329), un-shift rows
    // (:331-333: un-shifted row j holds cropped row (j + r/2 - 1) mod r), inverse transform along the rows' index
    {
        const int pl = threadIdx.x / rh, k = threadIdx.x - pl * rh;
        if (pl < np) {
            const int plane = plane0 + pl, n = plane / C, c = plane - n * C;
            const float* re = spec2 + ((long long)n * 2 * C + c) * R * Rh + k;
            const float* im = re + (long long)C * R * Rh;
            float2 v[r];
#pragma unroll
            for (int j = 0; j < r; ++j) {
                const int cj = (j + r / 2 - 1) & (r - 1);
                const int src = (R / 2 - r / 2 + cj) * Rh;
                const float g = __ldg(gm + cj * rh + k);
                v[j] = make_float2(__ldg(re + src) * g, __ldg(im + src) * g);
            }
            fft_reg<r, +1>(v);
            float2* cb = sm_small + pl * rc + k;
#pragma unroll
            for (int j = 0; j < r; ++j) cb[j * rh] = v[j];
        }
    }
    __syncthreads();
    // rows: thread = (plane, row pair): Hermitian extension along the last axis (imaginary parts of the DC and Nyquist bins
    // dropped), rows 2p and 2p+1 packed as real and imaginary part of one complex inverse transform
    {
        const int pl = threadIdx.x / (r / 2), p = threadIdx.x - pl * (r / 2);
        if (pl < np) {
            const float2* ra = sm_small + (pl * r + 2 * p) * rh;
            float2 v[r];
#pragma unroll
            for (int k = 0; k < r; ++k) {
                const int kk = k <= r / 2 ? k : r - k;
                float2 ya = ra[kk], yb = ra[rh + kk];
                if (k > r / 2) { ya.y = -ya.y; yb.y = -yb.y; }
                if (k == 0 || k == r / 2) { ya.y = 0.f; yb.y = 0.f; }
                v[k] = make_float2(ya.x - yb.y, ya.y + yb.x);
            }
            fft_reg<r, +1>(v);
            float* o0 = out + ((long long)(plane0 + pl) * r + 2 * p) * r;
#pragma unroll
            for (int q = 0; q < r / 4; ++q) {
                reinterpret_cast<float4*>(o0)[q] = make_float4(v[4 * q].x, v[4 * q + 1].x, v[4 * q + 2].x, v[4 * q + 3].x);
                reinterpret_cast<float4*>(o0 + r)[q] = make_float4(v[4 * q].y, v[4 * q + 1].y, v[4 * q + 2].y, v[4 * q + 3].y);
            }
        }
    }
}

// planes per CTA: both thread mappings (R/2 row pairs, Rh columns per plane) fit 256 threads
static inline int small_pp(int R) { return 256 / (R / 2 + 1); }

template <int R>
static int launch_fwd_small(const float* x, float* spec1, int N, int C, cudaStream_t stream) {
    static DeviceInit once;
    int num_sms = 0;
    if (int e = device_init(once, &num_sms, []() -> int {
            SHGAN_CUDA(cudaFuncSetAttribute(shu_rfft2_small_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
            return 0;
        })) return e;
    const int PP = small_pp(R), NC = N * C;
    const size_t smem = (size_t)PP * R * (R / 2 + 1) * sizeof(float2);
    shu_rfft2_small_kernel<R><<<ceil_div(NC, PP), 256, smem, stream>>>(x, spec1, NC, C, PP);
    SHGAN_LAUNCH_CHECK();
    return 0;
}

template <int r>
static int launch_inv_small(const float* spec2, const float* gm, float* out, int N, int C, int R, cudaStream_t stream) {
    static DeviceInit once;
    int num_sms = 0;
    if (int e = device_init(once, &num_sms, []() -> int {
            SHGAN_CUDA(cudaFuncSetAttribute(shu_irfft2_small_kernel<r>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
            return 0;
        })) return e;
    const int PP = small_pp(r), NC = N * C;
    const size_t smem = (size_t)PP * r * (r / 2 + 1) * sizeof(float2);
    shu_irfft2_small_kernel<r><<<ceil_div(NC, PP), 256, smem, stream>>>(spec2, gm, out, NC, C, R, PP);
    SHGAN_LAUNCH_CHECK();
    return 0;
}

int launch_shu_rfft2_small(const float* x, float* spec1, int N, int C, int R, cudaStream_t stream) {
    switch (R) {
        case 4: return launch_fwd_small<4>(x, spec1, N, C, stream);
        case 8: return launch_fwd_small<8>(x, spec1, N, C, stream);
        case 16: return launch_fwd_small<16>(x, spec1, N, C, stream);
        case 32: return launch_fwd_small<32>(x, spec1, N, C, stream);
    }
    set_error("launch_shu_rfft2_small: unsupported size");
    return 1;
}

int launch_shu_irfft2_small(const float* spec2, const float* gm, float* out, int N, int C, int R, int r, cudaStream_t stream) {
    switch (r) {
        case 4: return launch_inv_small<4>(spec2, gm, out, N, C, R, stream);
        case 8: return launch_inv_small<8>(spec2, gm, out, N, C, R, stream);
        case 16: return launch_inv_small<16>(spec2, gm, out, N, C, R, stream);
        case 32: return launch_inv_small<32>(spec2, gm, out, N, C, R, stream);
    }
    set_error("launch_shu_irfft2_small: unsupported size");
    return 1;
}

}  // namespace shgan
