// Register-resident radix-8 transforms of the Spectral Hint Unit at input_res 64 (the released model's size), sm_100a.
// Replace torch.fft.rfftn / irfftn of SHU.forward (lib/model_zoo/shgan.py:312-317 and :326-334) together with the row shift,
// the per-band crop, the Gaussian band masks and the un-shift.  cuFFT-free.
//
// A length-L transform (L = 8 M, M = 1, 2, 4, 8) is held by M consecutive lanes, 8 values per lane: radix-8 butterflies in
// registers, one twiddle multiply, ONE exchange through a conflict-free shared-memory scratch private to the M lanes
// (__syncwarp, no block barrier), radix-M butterflies in registers.  Lane t holds elements t + M a on input and on output.
// Register ROTATIONS of the input (element t + M ((a + rin) & 7) in register a) and of the output are folded into the
// per-thread twiddle table (a rotation of a DFT's input is a modulation of its output and vice versa), which is how every
// shared-memory access pattern below is made bank-conflict free without padding the bulk-copied planes.
//
// Forward, one (sample, channel) plane per iteration, 256 threads, persistent, planes double-buffered by 1-D bulk copies:
//   row pass     32 complex FFT-64: rows p and p + 32 packed as real and imaginary part
//   column pass  32 complex FFT-64: columns 1..31, and columns 0 and 32 (both real sequences) packed into one; the
//                untangling of the packed rows is folded into the loads (both rows of a pair land in the same thread)
//   the 1/R^2 'forward' normalisation rides on the twiddles; the DC-to-centre row shift on the store index; the plane leaves
//   through a staging buffer and two bulk stores in the kx-MAJOR layout spec[n, ch, kx, s] that the channel mix tiles over.
// Inverse, one plane per iteration, 384 threads: every band r = lowest_res..64 at once (thread ranges per band), crop +
//   Gaussian mask (per-thread registers, constant over planes) + un-shift folded into the loads, r/2+1 column transforms,
//   Hermitian extension with the DC/Nyquist imaginary parts dropped (C2R semantics on non-Hermitian input) and the pairing of
//   output rows y, y + r/2 done by the thread that holds both, r/2 packed row transforms, bulk stores of the r x r planes.
#include "shu_internal.cuh"
#include "tma_util.cuh"

namespace shgan {

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmulf(float2 a, float2 b) { return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x)); }
// a * (i * S)
template <int S>
__device__ __forceinline__ float2 mul_i(float2 a) { return S > 0 ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x); }

// natural-order DFTs, kernel exp(S * 2 pi i n k / L)
template <int S>
__device__ __forceinline__ void dft2(float2& y0, float2& y1) {
    const float2 a = y0;
    y0 = cadd(a, y1);
    y1 = csub(a, y1);
}
template <int S>
__device__ __forceinline__ void dft4(float2& y0, float2& y1, float2& y2, float2& y3) {
    const float2 s0 = cadd(y0, y2), s1 = csub(y0, y2), s2 = cadd(y1, y3), s3 = mul_i<S>(csub(y1, y3));
    y0 = cadd(s0, s2);
    y2 = csub(s0, s2);
    y1 = cadd(s1, s3);
    y3 = csub(s1, s3);
}
template <int S>
__device__ __forceinline__ void dft8(float2 (&v)[8]) {
    float2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6], o0 = v[1], o1 = v[3], o2 = v[5], o3 = v[7];
    dft4<S>(e0, e1, e2, e3);
    dft4<S>(o0, o1, o2, o3);
    constexpr float r = 0.70710678118654752f, s = (float)S;
    o1 = make_float2((o1.x - s * o1.y) * r, (o1.y + s * o1.x) * r);        // * (1 + iS) / sqrt 2
    o2 = mul_i<S>(o2);
    o3 = make_float2((-o3.x - s * o3.y) * r, (s * o3.x - o3.y) * r);       // * (-1 + iS) / sqrt 2
    v[0] = cadd(e0, o0); v[4] = csub(e0, o0);
    v[1] = cadd(e1, o1); v[5] = csub(e1, o1);
    v[2] = cadd(e2, o2); v[6] = csub(e2, o2);
    v[3] = cadd(e3, o3); v[7] = csub(e3, o3);
}

// tw[c] = scale * exp(S 2 pi i (t c / L + rin c / 8 + t rout / M)), L = 8 M
template <int M>
__device__ __forceinline__ void make_tw(float2 (&tw)[8], int S, int t, int rin, int rout, float scale) {
    constexpr int L = 8 * M;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const int ex = (t * c + rin * c * M + t * rout * 8) & (L - 1);
        float sn, cs;
        sincospif(2.f * (float)ex / (float)L, &sn, &cs);
        tw[c] = make_float2(cs * scale, (float)S * sn * scale);
    }
}

// scratch of one transform group: [c = 0..7][b = 0..M-1] at a pitch of M + 1 float2
// and the groups of a warp at a pitch that spreads them over the banks (72 / 52 / 34 float2 for M = 8 / 4 / 2: the 16 lanes
// of a half-warp -- 2 / 4 / 8 groups -- then touch 16 distinct bank pairs in both the column-of-scratch writes and the row reads)
template <int M> struct FftScratch { static constexpr int PITCH = M + 1, SIZE = M == 8 ? 72 : (M == 4 ? 52 : (M == 2 ? 34 : 16)); };

// In: v[a] = x[t + M ((a + rin) & 7)].  Out: v[q] = X[t + M e + 8 ((d + rout) % M)], q = e + (8 / M) d.
template <int M, int S>
__device__ __forceinline__ void fft_8xM(float2 (&v)[8], const float2 (&tw)[8], float2* ex, int t, unsigned gmask) {
    dft8<S>(v);
    if (M == 1) return;
    constexpr int P = FftScratch<M>::PITCH, E = 8 / M;
#pragma unroll
    for (int c = 0; c < 8; ++c) ex[c * P + t] = cmulf(v[c], tw[c]);
    __syncwarp(gmask);
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const float2* src = ex + (t + M * e) * P;
        float2 u[M];
#pragma unroll
        for (int b = 0; b < M; ++b) u[b] = src[b];
        if (M == 8) {
            float2 w[8];
#pragma unroll
            for (int b = 0; b < 8; ++b) w[b] = u[b % M];
            dft8<S>(w);
#pragma unroll
            for (int d = 0; d < 8; ++d) v[d] = w[d];
        } else if (M == 4) {
            dft4<S>(u[0], u[1 % M], u[2 % M], u[3 % M]);
#pragma unroll
            for (int d = 0; d < M; ++d) v[e + E * d] = u[d];
        } else {
            dft2<S>(u[0], u[1 % M]);
#pragma unroll
            for (int d = 0; d < M; ++d) v[e + E * d] = u[d];
        }
    }
    __syncwarp(gmask);
}

// =============================================== forward =====================================================================
constexpr int F64_THREADS = 256;
constexpr int F64_ZP = 66;                        // float2 pitch of the row-pass result Z[pair][k]
constexpr int F64_BINS = 33 * 64;
constexpr int F64_OFF_Z = 2 * 4096 * 4;
constexpr int F64_OFF_EX = F64_OFF_Z + 32 * F64_ZP * 8;
constexpr int F64_OFF_ST = F64_OFF_EX + 32 * FftScratch<8>::SIZE * 8;
constexpr int F64_OFF_BAR = F64_OFF_ST + 2 * 2 * F64_BINS * 4;
constexpr int F64_SMEM = F64_OFF_BAR + 16;

__global__ void __launch_bounds__(F64_THREADS, 2)
shu_rfft2_r64_kernel(const float* __restrict__ x, float* __restrict__ spec1, const float* __restrict__ cw, float* __restrict__ cw_kxmajor,
                     int planes, int C) {
    extern __shared__ __align__(128) uint8_t sm[];
    float* xin = reinterpret_cast<float*>(sm);                          // [2][64][64]
    float2* zs = reinterpret_cast<float2*>(sm + F64_OFF_Z);             // [32][F64_ZP]
    float2* exs = reinterpret_cast<float2*>(sm + F64_OFF_EX);           // [32 groups]
    float* stage = reinterpret_cast<float*>(sm + F64_OFF_ST);           // [2][re | im][33][64]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + F64_OFF_BAR);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, fl = lane >> 3, t = lane & 7;
    // row pass: transform p packs rows p and p + 32 (half-warp partners 4 apart: conflict-free Z stores at pitch 66)
    const int p = (warp & 3) + 4 * (fl & 1) + 8 * (fl >> 1) + 16 * (warp >> 2);
    // column pass: spectrum column k (k = 0: columns 0 and 32 packed)
    const int k = 4 * warp + fl;
    const int kp = k == 0 ? 32 : 64 - k;
    float2 twr[8], twc[8];
    make_tw<8>(twr, -1, t, fl, 0, 1.f);
    make_tw<8>(twc, -1, t, 0, fl, (k == 0 ? 1.f : 0.5f) / 4096.f);
    float2* ex = exs + (warp * 4 + fl) * FftScratch<8>::SIZE;
    const unsigned gmask = 0xFFu << (fl * 8);

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    // the blend weights of the channel mix in the bin order of the spectrum written here (kx-major), for the launch that follows
    if (blockIdx.x == 0 && cw_kxmajor) {
        for (int i = tid; i < 6 * F64_BINS; i += F64_THREADS) {
            const int k6 = i / F64_BINS, e = i - k6 * F64_BINS;
            cw_kxmajor[i] = __ldg(cw + k6 * F64_BINS + (e & 63) * 33 + (e >> 6));
        }
    }
    if (tid == 0) {
        for (int j = 0; j < 2; ++j) {
            const long long pl = (long long)blockIdx.x + (long long)j * gridDim.x;
            if (pl < planes) {
                mbar_expect_tx(&bars[j], 16384u);
                bulk_g2s(xin + j * 4096, x + pl * 4096, 16384u, &bars[j]);
            }
        }
    }
    int it = 0;
    for (int pl = blockIdx.x; pl < planes; pl += gridDim.x, ++it) {
        const int buf = it & 1;
        mbar_wait(&bars[buf], (uint32_t)((it >> 1) & 1));
        const float* xb = xin + buf * 4096;
        float2 v[8];
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            const int col = 8 * ((a + fl) & 7) + t;
            v[a] = make_float2(xb[p * 64 + col], xb[(p + 32) * 64 + col]);
        }
        fft_8xM<8, -1>(v, twr, ex, t, gmask);
#pragma unroll
        for (int d = 0; d < 8; ++d) zs[p * F64_ZP + t + 8 * d] = v[d];
        if (tid == 0) bulk_wait_read<1>();                 // the staging buffer of two planes ago has left
        __syncthreads();
        if (tid == 0) {
            const long long nxt = (long long)pl + 2LL * gridDim.x;
            if (nxt < planes) {
                mbar_expect_tx(&bars[buf], 16384u);
                bulk_g2s(xin + buf * 4096, x + nxt * 4096, 16384u, &bars[buf]);
            }
        }
        // column pass; rows t + 8a (a < 4) are the real parts' transforms A of pairs t + 8a, rows t + 8a + 32 the B's
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const float2 za = zs[(t + 8 * a) * F64_ZP + k], zb = zs[(t + 8 * a) * F64_ZP + kp];
            if (k != 0) {
                v[a] = make_float2(za.x + zb.x, za.y - zb.y);          // 2 A[k]   (the 1/2 rides on the twiddles)
                v[a + 4] = make_float2(za.y + zb.y, zb.x - za.x);      // 2 B[k]
            } else {
                v[a] = make_float2(za.x, zb.x);                        // A[0] + i A[32]
                v[a + 4] = make_float2(za.y, zb.y);                    // B[0] + i B[32]
            }
        }
        fft_8xM<8, -1>(v, twc, ex, t, gmask);
        float* sre = stage + buf * 2 * F64_BINS;
        float* sim = sre + F64_BINS;
        if (k != 0) {
#pragma unroll
            for (int d = 0; d < 8; ++d) {
                const int ky = t + 8 * ((d + fl) & 7);
                const int s = (ky - 33) & 63;                          // row shift of shgan.py:315-317
                sre[k * 64 + s] = v[d].x;
                sim[k * 64 + s] = v[d].y;
            }
        } else {
            // C[ky] = col0[ky] + i col32[ky], both Hermitian: untangle with C[-ky], held by lane (8 - t) & 7
#pragma unroll
            for (int d = 0; d < 8; ++d) {
                float px = __shfl_sync(0xFFu, v[7 - d].x, (8 - t) & 7);
                float py = __shfl_sync(0xFFu, v[7 - d].y, (8 - t) & 7);
                if (t == 0) { px = v[(8 - d) & 7].x; py = v[(8 - d) & 7].y; }
                const int s = (t + 8 * d - 33) & 63;
                sre[s] = 0.5f * (v[d].x + px);
                sim[s] = 0.5f * (v[d].y - py);
                sre[32 * 64 + s] = 0.5f * (v[d].y + py);
                sim[32 * 64 + s] = 0.5f * (px - v[d].x);
            }
        }
        fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            const int n = pl / C, c = pl - n * C;
            bulk_s2g(spec1 + ((long long)n * 2 * C + c) * F64_BINS, sre, F64_BINS * 4);
            bulk_s2g(spec1 + ((long long)n * 2 * C + C + c) * F64_BINS, sim, F64_BINS * 4);
            bulk_commit();
        }
    }
    if (tid == 0) bulk_wait_read<0>();
}

// =============================================== inverse =====================================================================
constexpr int I64_THREADS = 448;
// thread ranges of the bands: (r/2 + 1) * M threads in the column stage, (r/2) * M in the row stage, M = max(r / 8, 1)
// band starts are warp-aligned so that no warp runs two long roles one after the other (its lanes would serialise them and every
// other warp would wait for it at the block barrier): warp 8 holds the 33rd column transform of r = 64 only, r = 32 starts at
// warp 9, r = 16 has warp 12, r = 8 and r = 4 (short roles) share warp 13
constexpr int I64_B64 = 0, I64_B32 = 288, I64_B16 = 384, I64_B8 = 416, I64_B4 = 421, I64_END = 424;
template <int r> struct InvBand {
    static constexpr int M = r >= 8 ? r / 8 : 1;
    static constexpr int RH = r / 2 + 1;
    static constexpr int ZP = r == 64 ? 66 : (r == 32 ? 36 : (r == 16 ? 20 : r + 2));   // float2 pitch of Zrow[pair][k] (bank spreading)
    static constexpr int NCOL = RH * M, NROW = (r / 2) * M;
    static constexpr int ZROW_F2 = (r / 2) * ZP;
    static constexpr int EX_F2 = RH * FftScratch<M>::SIZE;
};
constexpr int I64_OFF_IN = 0;                                                   // [2][re | im][33][64] fp32
constexpr int I64_OFF_Z64 = I64_OFF_IN + 2 * 2 * F64_BINS * 4;
constexpr int I64_OFF_Z32 = I64_OFF_Z64 + InvBand<64>::ZROW_F2 * 8;
constexpr int I64_OFF_Z16 = I64_OFF_Z32 + InvBand<32>::ZROW_F2 * 8;
constexpr int I64_OFF_Z8 = I64_OFF_Z16 + InvBand<16>::ZROW_F2 * 8;
constexpr int I64_OFF_Z4 = I64_OFF_Z8 + InvBand<8>::ZROW_F2 * 8;
constexpr int I64_OFF_EX64 = I64_OFF_Z4 + InvBand<4>::ZROW_F2 * 8;
constexpr int I64_OFF_EX32 = I64_OFF_EX64 + InvBand<64>::EX_F2 * 8;
constexpr int I64_OFF_EX16 = I64_OFF_EX32 + InvBand<32>::EX_F2 * 8;
constexpr int I64_OFF_OUT = (I64_OFF_EX16 + InvBand<16>::EX_F2 * 8 + 127) & ~127;   // out planes 64 | 32 | 16 | 8 | 4
constexpr int I64_OUT_FLOATS = 64 * 64 + 32 * 32 + 16 * 16 + 8 * 8 + 4 * 4;
constexpr int I64_OFF_BAR = I64_OFF_OUT + I64_OUT_FLOATS * 4;
constexpr int I64_SMEM = I64_OFF_BAR + 32;                                       // full[2] + empty[2]

__device__ __forceinline__ unsigned group_mask(int lane, int M) { return (M >= 32 ? 0xFFFFFFFFu : ((1u << M) - 1u)) << (lane & ~(M - 1)); }

// per-thread constants of a band's column role: Gaussian factors of its 8 elements, twiddles
template <int r>
__device__ __forceinline__ void inv_col_init(int lt, const float* __restrict__ gm, float (&g)[8], float2 (&tw)[8]) {
    constexpr int M = InvBand<r>::M, RH = InvBand<r>::RH;
    const int kx = lt / M, t = lt % M;
    const int rin = M > 1 ? (kx & 7) : 0;
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        const int j = t + M * ((a + rin) & 7);
        const int cj = (j + r / 2 - 1) & (r - 1);
        g[a] = (r >= 8 || a < 4) ? __ldg(gm + cj * RH + kx) : 0.f;
    }
    make_tw<M>(tw, +1, t, rin, 0, 1.f);
}

// column stage of band r >= 8: crop rows [32 - r/2, 32 + r/2), cols [0, r/2] (shgan.py:328), Gaussian mask (:329), un-shift
// (:331-333: un-shifted row j holds cropped row (j + r/2 - 1) mod r), inverse transforms over j; rows y and y + r/2 (held by
// the same thread) leave paired and Hermitian-extended as the packed row-stage input Z[y][k] = Ya[k] + i Yb[k]
template <int r>
__device__ __forceinline__ void inv_cols(int lt, int lane, const float* __restrict__ in_re, const float* __restrict__ in_im,
                                         const float (&g)[8], const float2 (&tw)[8], float2* zrow, float2* exb) {
    constexpr int M = InvBand<r>::M, ZP = InvBand<r>::ZP;
    const int kx = lt / M, t = lt % M;
    const int rin = M > 1 ? (kx & 7) : 0;
    float2 v[8];
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        const int j = t + M * ((a + rin) & 7);
        const int s = 32 - r / 2 + ((j + r / 2 - 1) & (r - 1));
        v[a] = make_float2(in_re[kx * 64 + s] * g[a], in_im[kx * 64 + s] * g[a]);
    }
    fft_8xM<M, +1>(v, tw, exb + kx * FftScratch<M>::SIZE, t, group_mask(lane, M));
    const bool edge = kx == 0 || kx == r / 2;               // C2R: imaginary parts of the DC / Nyquist columns are dropped
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int y = t + M * q;
        const float2 ya = v[q], yb = v[q + 4];
        if (edge) {
            zrow[y * ZP + kx] = make_float2(ya.x, yb.x);
        } else {
            zrow[y * ZP + kx] = make_float2(ya.x - yb.y, ya.y + yb.x);
            zrow[y * ZP + r - kx] = make_float2(ya.x + yb.y, yb.x - ya.y);
        }
    }
}

template <int r>
__device__ __forceinline__ void inv_row_role(int lt, int& y1, int& t, int& rout) {
    constexpr int M = InvBand<r>::M;
    const int fi = lt / M;
    t = lt % M;
    if (M == 8) {
        const int w = fi >> 2, f = fi & 3;
        y1 = (w & 3) + 4 * (f & 1) + 8 * (f >> 1) + 16 * (w >> 2);
        rout = f;
    } else {
        y1 = fi;
        rout = M > 1 ? (fi & (M - 1)) : 0;
    }
}

// row stage of band r >= 8: r/2 packed inverse transforms, real part -> row y1, imaginary part -> row y1 + r/2
template <int r>
__device__ __forceinline__ void inv_rows(int lt, int lane, const float2* zrow, float* out, const float2 (&tw)[8], float2* exb) {
    constexpr int M = InvBand<r>::M, ZP = InvBand<r>::ZP, E = 8 / M;
    int y1, t, rout;
    inv_row_role<r>(lt, y1, t, rout);
    float2 v[8];
#pragma unroll
    for (int a = 0; a < 8; ++a) v[a] = zrow[y1 * ZP + t + M * a];
    fft_8xM<M, +1>(v, tw, exb + (lt / M) * FftScratch<M>::SIZE, t, group_mask(lane, M));
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int e = q % E, d = q / E;
        const int xx = M > 1 ? t + M * e + 8 * ((d + rout) & (M - 1)) : q;
        out[y1 * r + xx] = v[q].x;
        out[(y1 + r / 2) * r + xx] = v[q].y;
    }
}

__global__ void __launch_bounds__(I64_THREADS, 2)
shu_irfft2_r64_kernel(const float* __restrict__ spec2, const float* __restrict__ gauss, const ShuBands bands, int planes, int C) {
    extern __shared__ __align__(128) uint8_t sm[];
    float* in = reinterpret_cast<float*>(sm + I64_OFF_IN);
    float2* z64 = reinterpret_cast<float2*>(sm + I64_OFF_Z64);
    float2* z32 = reinterpret_cast<float2*>(sm + I64_OFF_Z32);
    float2* z16 = reinterpret_cast<float2*>(sm + I64_OFF_Z16);
    float2* z8 = reinterpret_cast<float2*>(sm + I64_OFF_Z8);
    float2* z4 = reinterpret_cast<float2*>(sm + I64_OFF_Z4);
    float2* ex64 = reinterpret_cast<float2*>(sm + I64_OFF_EX64);
    float2* ex32 = reinterpret_cast<float2*>(sm + I64_OFF_EX32);
    float2* ex16 = reinterpret_cast<float2*>(sm + I64_OFF_EX16);
    float* o64 = reinterpret_cast<float*>(sm + I64_OFF_OUT);
    float* o32 = o64 + 64 * 64;
    float* o16 = o32 + 32 * 32;
    float* o8 = o16 + 16 * 16;
    float* o4 = o8 + 8 * 8;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + I64_OFF_BAR);
    const int tid = threadIdx.x, lane = tid & 31;

    // role: band (log2 r) and thread index inside the band
    int lr = 0, lt = 0;
    if (tid < I64_B32) { lr = 6; lt = tid - I64_B64; }
    else if (tid < I64_B16) { lr = 5; lt = tid - I64_B32; }
    else if (tid < I64_B8) { lr = 4; lt = tid - I64_B16; }
    else if (tid < I64_B4) { lr = 3; lt = tid - I64_B8; }
    else if (tid < I64_END) { lr = 2; lt = tid - I64_B4; }
    // (threads past a band's last transform fall out through col_on / row_on below)
    if (lr < bands.lowest_log2) lr = 0;                    // this band is not produced
    const int bi = lr - bands.lowest_log2;
    const float* gm = gauss + (lr ? bands.gauss_off[bi] : 0);
    float g[8];
    float2 twc[8], twr[8];
    bool col_on = false, row_on = false;
    switch (lr) {
    case 6: col_on = lt < InvBand<64>::NCOL; row_on = lt < InvBand<64>::NROW; break;
    case 5: col_on = lt < InvBand<32>::NCOL; row_on = lt < InvBand<32>::NROW; break;
    case 4: col_on = lt < InvBand<16>::NCOL; row_on = lt < InvBand<16>::NROW; break;
    case 3: col_on = lt < InvBand<8>::NCOL; row_on = lt < InvBand<8>::NROW; break;
    case 2: col_on = lt < InvBand<4>::NCOL; row_on = lt < InvBand<4>::NROW; break;
    default: break;
    }
#pragma unroll
    for (int a = 0; a < 8; ++a) { g[a] = 0.f; twc[a] = twr[a] = make_float2(1.f, 0.f); }
    if (col_on) {
        switch (lr) {
        case 6: inv_col_init<64>(lt, gm, g, twc); break;
        case 5: inv_col_init<32>(lt, gm, g, twc); break;
        case 4: inv_col_init<16>(lt, gm, g, twc); break;
        case 3: inv_col_init<8>(lt, gm, g, twc); break;
        case 2:
#pragma unroll
            for (int a = 0; a < 4; ++a) g[a] = __ldg(gm + ((a + 1) & 3) * 3 + lt);     // cropped row (j + r/2 - 1) mod 4 of column lt
            break;
        }
    }
    if (row_on) {
        int y1, t, rout;
        switch (lr) {
        case 6: inv_row_role<64>(lt, y1, t, rout); make_tw<8>(twr, +1, t, 0, rout, 1.f); break;
        case 5: inv_row_role<32>(lt, y1, t, rout); make_tw<4>(twr, +1, t, 0, rout, 1.f); break;
        case 4: inv_row_role<16>(lt, y1, t, rout); make_tw<2>(twr, +1, t, 0, rout, 1.f); break;
        default: break;
        }
    }

    // Two decoupled groups share the double-buffered input plane: A = warps 0-8 (band 64), B = warps 9-13 (the smaller bands).
    // Each group has its own named barrier, its own output staging and its own leader issuing its bulk stores, so that the
    // long pole of one group (B's warps run short roles with few active lanes) does not stall the other at a block barrier
    // twice per plane; they only meet through the input buffers' empty barriers (two arrivals: one per group), two planes
    // apart.  (Measured: 200 -> 183 us at batch 512.  Giving every band its own plane loop -- so that its constants become
    // registers of that loop instead of a 160-byte stack frame indexed through the switch below -- was built twice and was
    // SLOWER both times, 257 and 284 us: five loop bodies no longer share an instruction-cache footprint.)
    constexpr int GA_THREADS = I64_B32, GB_THREADS = I64_THREADS - I64_B32;
    static_assert(GA_THREADS % 32 == 0 && GB_THREADS % 32 == 0, "groups are whole warps");
    uint64_t* in_empty = bars + 2;
    const bool group_a = tid < GA_THREADS;
    const bool leader = tid == 0 || tid == GA_THREADS;            // first thread of each group
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_init(&in_empty[0], 2);
        mbar_init(&in_empty[1], 2);
        mbar_fence_init();
    }
    __syncthreads();
    auto group_barrier = [&]() {
        if (group_a) asm volatile("bar.sync 1, %0;" ::"n"(GA_THREADS) : "memory");
        else asm volatile("bar.sync 2, %0;" ::"n"(GB_THREADS) : "memory");
    };
    auto issue_load = [&](long long pl, int buf) {
        const long long n = pl / C, c = pl - n * C;
        mbar_expect_tx(&bars[buf], 2u * F64_BINS * 4u);
        bulk_g2s(in + buf * 2 * F64_BINS, spec2 + (n * 2 * C + c) * F64_BINS, F64_BINS * 4u, &bars[buf]);
        bulk_g2s(in + buf * 2 * F64_BINS + F64_BINS, spec2 + (n * 2 * C + C + c) * F64_BINS, F64_BINS * 4u, &bars[buf]);
    };
    if (tid == 0) {
        for (int j = 0; j < 2; ++j) {
            const long long pl = (long long)blockIdx.x + (long long)j * gridDim.x;
            if (pl < planes) issue_load(pl, j);
        }
    }
    int it = 0;
    for (int pl = blockIdx.x; pl < planes; pl += gridDim.x, ++it) {
        const int buf = it & 1;
        mbar_wait(&bars[buf], (uint32_t)((it >> 1) & 1));
        const float* in_re = in + buf * 2 * F64_BINS;
        const float* in_im = in_re + F64_BINS;
        if (col_on) {
            switch (lr) {
            case 6: inv_cols<64>(lt, lane, in_re, in_im, g, twc, z64, ex64); break;
            case 5: inv_cols<32>(lt, lane, in_re, in_im, g, twc, z32, ex32); break;
            case 4: inv_cols<16>(lt, lane, in_re, in_im, g, twc, z16, ex16); break;
            case 3: inv_cols<8>(lt, lane, in_re, in_im, g, twc, z8, ex16); break;
            case 2: {
                // r = 4: one thread per column kx = lt, DFT-4 over the un-shifted rows j (cropped rows (j + 1) & 3 at s = 30 + ..)
                float2 y[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int s = 30 + ((j + 1) & 3);
                    y[j] = make_float2(in_re[lt * 64 + s] * g[j], in_im[lt * 64 + s] * g[j]);
                }
                dft4<+1>(y[0], y[1], y[2], y[3]);
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    if (lt != 1) z4[q * 6 + lt] = make_float2(y[q].x, y[q + 2].x);
                    else {
                        z4[q * 6 + 1] = make_float2(y[q].x - y[q + 2].y, y[q].y + y[q + 2].x);
                        z4[q * 6 + 3] = make_float2(y[q].x + y[q + 2].y, y[q + 2].x - y[q].y);
                    }
                }
                break;
            }
            }
        }
        if (leader) bulk_wait_read<0>();                   // this group's previous output staging has left
        group_barrier();                                   // the group is done reading the input plane
        if (leader) mbar_arrive(&in_empty[buf]);
        if (tid == 0) {
            const long long nxt = (long long)pl + 2LL * gridDim.x;
            if (nxt < planes) {
                mbar_wait(&in_empty[buf], (uint32_t)((it >> 1) & 1));   // ... and so is the other group
                issue_load(nxt, buf);
            }
        }
        if (row_on) {
            switch (lr) {
            case 6: inv_rows<64>(lt, lane, z64, o64, twr, ex64); break;
            case 5: inv_rows<32>(lt, lane, z32, o32, twr, ex32); break;
            case 4: inv_rows<16>(lt, lane, z16, o16, twr, ex16); break;
            case 3: inv_rows<8>(lt, lane, z8, o8, twr, ex16); break;
            case 2: {
                float2 y[4];
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) y[kk] = z4[lt * 6 + kk];
                dft4<+1>(y[0], y[1], y[2], y[3]);
#pragma unroll
                for (int xx = 0; xx < 4; ++xx) {
                    o4[lt * 4 + xx] = y[xx].x;
                    o4[(lt + 2) * 4 + xx] = y[xx].y;
                }
                break;
            }
            }
        }
        fence_async_smem();
        group_barrier();
        if (tid == 0) {
            bulk_s2g(bands.out[6 - bands.lowest_log2] + (long long)pl * 64 * 64, o64, 64u * 64u * 4u);
            bulk_commit();
        } else if (tid == GA_THREADS) {
            const float* src[4] = {o4, o8, o16, o32};
            for (int l2 = bands.lowest_log2; l2 < 6; ++l2) {
                const int r = 1 << l2;
                bulk_s2g(bands.out[l2 - bands.lowest_log2] + (long long)pl * r * r, src[l2 - 2], (uint32_t)(r * r * 4));
            }
            bulk_commit();
        }
    }
    if (leader) bulk_wait_read<0>();
}

// =============================================== host ========================================================================
int launch_shu_rfft2_r64(const float* x, float* spec1, const float* cw, float* cw_kxmajor, int N, int C, cudaStream_t stream) {
    static DeviceInit once;
    int num_sms = 148;
    if (int e = device_init(once, &num_sms, []() -> int {
            SHGAN_CUDA(cudaFuncSetAttribute(shu_rfft2_r64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, F64_SMEM));
            return 0;
        })) return e;
    const int planes = N * C;
    const int grid = planes < 2 * num_sms ? planes : 2 * num_sms;
    shu_rfft2_r64_kernel<<<grid, F64_THREADS, F64_SMEM, stream>>>(x, spec1, cw, cw_kxmajor, planes, C);
    SHGAN_LAUNCH_CHECK();
    return 0;
}

int launch_shu_irfft2_r64(const float* spec2, const float* gauss, const ShuBands& bands, int N, int C, cudaStream_t stream) {
    static DeviceInit once;
    int num_sms = 148;
    if (int e = device_init(once, &num_sms, []() -> int {
            SHGAN_CUDA(cudaFuncSetAttribute(shu_irfft2_r64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, I64_SMEM));
            return 0;
        })) return e;
    const int planes = N * C;
    const int grid = planes < 2 * num_sms ? planes : 2 * num_sms;      // (one 448-thread CTA per SM at 128 registers measured 238 us against 187)
    shu_irfft2_r64_kernel<<<grid, I64_THREADS, I64_SMEM, stream>>>(spec2, gauss, bands, planes, C);
    SHGAN_LAUNCH_CHECK();
    return 0;
}

}  // namespace shgan
