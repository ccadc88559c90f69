// 4x4 FIR (blur) on NHWC activations with the fused pointwise epilogue (sm_100a).
//
// Replaces the upfirdn2d passes that surround the resampled convolutions of the reference
// (lib/model_zoo/stylegan_utils/conv2d_resample.py:117-120 blur before the stride-2 conv, :139 blur
// after the stride-2 transposed conv) together with every elementwise pass that follows them
// (stylegan.py:191-192, 298-303; comodgan.py:319-327).
//
// HBM-bound; the first version of this kernel was instruction-bound instead (ncu: 14% DRAM, ~94 instructions per
// output element: every output re-loaded and re-converted a 4x7 input window).  This version is a sliding window
// down the image: one thread owns 2 horizontally adjacent output pixels x 8 channels (channels fastest across
// lanes, so every global access is a 16 B / 32 B vector and a warp covers whole 128 B lines) and walks FIR_TY output
// rows.  Each input row is loaded and converted once per thread (5 pixel vectors) and scattered into the four
// partially accumulated output rows it contributes to, which live in registers as a sliding window; a finished
// row goes through the fused epilogue and is stored.  The row loop is deliberately not unrolled (one copy of the
// body and of the epilogue): ncu showed the unrolled variant stalled on instruction fetch for most issue slots.
#include "common.cuh"

namespace shgan {

constexpr int FIR_T = 4;     // filter taps per axis
constexpr int FIR_SX = 2;    // output pixels per thread along x
constexpr int FIR_TY = 16;   // output rows per thread
constexpr int FIR_THREADS = 128;

template <bool IN_F32>
__global__ void __launch_bounds__(FIR_THREADS, 4)
fir4x4_nhwc_kernel(const float* __restrict__ in_f32, const __half* __restrict__ in_hi, const __half* __restrict__ in_lo,
                   const float* __restrict__ f, float gain, int N, int C, int IH, int IW, int OH, int OW,
                   int pad_x0, int pad_y0, EpiParams epi, int parity_split, long long total) {
    __shared__ float s_f[FIR_T * FIR_T];
    if (threadIdx.x < FIR_T * FIR_T) s_f[threadIdx.x] = f[threadIdx.x] * gain;
    __syncthreads();
    float fk[FIR_T][FIR_T];
#pragma unroll
    for (int i = 0; i < FIR_T; ++i)
#pragma unroll
        for (int j = 0; j < FIR_T; ++j) fk[i][j] = s_f[i * FIR_T + j];
    const int cgs = C / 8;
    const int xs = (OW + FIR_SX - 1) / FIR_SX;
    const int ys = (OH + FIR_TY - 1) / FIR_TY;
    const int PH = (OH + 1) / 2, PW = (OW + 1) / 2;
    for (long long gid = blockIdx.x * (long long)blockDim.x + threadIdx.x; gid < total;
         gid += (long long)gridDim.x * blockDim.x) {
        const int cg = (int)(gid % cgs);
        long long t = gid / cgs;
        const int sx = (int)(t % xs);
        t /= xs;
        const int sy = (int)(t % ys);
        const int n = (int)(t / ys);
        const int x0 = sx * FIR_SX, y0 = sy * FIR_TY, c0 = cg * 8;
        const int y1 = y0 + FIR_TY < OH ? y0 + FIR_TY : OH;      // output rows [y0, y1)
        const int nin = y1 - y0 + FIR_T - 1;                      // input rows to stream
        const int iy0 = y0 - pad_y0, ix0 = x0 - pad_x0;

        // acc[s] = partially accumulated output row (jrow - 3 + s) while input row jrow is being consumed
        float acc[FIR_T][FIR_SX][8];
#pragma unroll
        for (int a = 0; a < FIR_T; ++a)
#pragma unroll
            for (int b = 0; b < FIR_SX; ++b)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[a][b][j] = 0.f;

        // NOT unrolled: one copy of the row body and of the epilogue keeps the kernel inside the instruction cache
        // (the 4x-unrolled first attempt spent most of its issue slots stalled on instruction fetch)
#pragma unroll 1
        for (int jrow = 0; jrow < nin; ++jrow) {
            const int iy = iy0 + jrow;
            if (iy >= 0 && iy < IH) {
                float row[FIR_SX + FIR_T - 1][8];
#pragma unroll
                for (int c = 0; c < FIR_SX + FIR_T - 1; ++c) {
                    const int ix = ix0 + c;
                    if (ix >= 0 && ix < IW) {
                        const long long idx = (((long long)n * IH + iy) * IW + ix) * C + c0;
                        if (IN_F32) {
                            const float4 a = __ldg(reinterpret_cast<const float4*>(in_f32 + idx));
                            const float4 b = __ldg(reinterpret_cast<const float4*>(in_f32 + idx + 4));
                            row[c][0] = a.x; row[c][1] = a.y; row[c][2] = a.z; row[c][3] = a.w;
                            row[c][4] = b.x; row[c][5] = b.y; row[c][6] = b.z; row[c][7] = b.w;
                        } else {
                            load_planes8(in_hi, in_lo, idx, row[c]);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) row[c][j] = 0.f;
                    }
                }
                // input row jrow carries filter row r = 3 - s for the output row held in slot s
#pragma unroll
                for (int sl = 0; sl < FIR_T; ++sl)
#pragma unroll
                    for (int ox = 0; ox < FIR_SX; ++ox)
#pragma unroll
                        for (int fx = 0; fx < FIR_T; ++fx)
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                acc[sl][ox][j] = fmaf(row[ox + fx][j], fk[FIR_T - 1 - sl][fx], acc[sl][ox][j]);
            }
            // slot 0 (output row jrow - 3) has now received all four filter rows
            const int yo = y0 + jrow - (FIR_T - 1);
            if (jrow >= FIR_T - 1 && yo < y1) {
#pragma unroll
                for (int ox = 0; ox < FIR_SX; ++ox) {
                    const int x = x0 + ox;
                    if (x < OW) {
                        long long out_pix = ((long long)n * OH + yo) * OW + x;
                        if (parity_split) {
                            const int q = (yo & 1) * 2 + (x & 1);
                            out_pix = (long long)q * N * PH * PW + ((long long)n * PH + (yo >> 1)) * PW + (x >> 1);
                        }
                        float rgb[3] = {0.f, 0.f, 0.f};
                        epilogue_apply<8>(epi, acc[0][ox], n, yo, x, OH, OW, C, c0, rgb, out_pix);
                    }
                }
            }
            // slide the window down by one output row
#pragma unroll
            for (int sl = 0; sl < FIR_T - 1; ++sl)
#pragma unroll
                for (int ox = 0; ox < FIR_SX; ++ox)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[sl][ox][j] = acc[sl + 1][ox][j];
#pragma unroll
            for (int ox = 0; ox < FIR_SX; ++ox)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[FIR_T - 1][ox][j] = 0.f;
        }
    }
}

}  // namespace shgan

using namespace shgan;

extern "C" int shgan_fir_nhwc(const float* in_f32, const void* in_hi, const void* in_lo, const float* f, int fH, int fW,
                              float gain, int N, int C, int IH, int IW, int pad_x0, int pad_x1, int pad_y0, int pad_y1,
                              const shgan_epilogue* epi_, int parity_split, void* stream) {
    SHGAN_CHECK(f && epi_, "null pointer");
    SHGAN_CHECK((in_f32 != nullptr) != (in_hi != nullptr), "exactly one of in_f32 / in_hi must be given");
    SHGAN_CHECK(in_f32 || in_lo, "in_lo missing");
    SHGAN_CHECK(fH == FIR_T && fW == FIR_T, "only 4x4 filters are supported on NHWC data");
    SHGAN_CHECK(N >= 0 && C >= 8 && C % 8 == 0 && IH >= 1 && IW >= 1, "bad tensor size (C must be a multiple of 8)");
    const int OH = IH + pad_y0 + pad_y1 - fH + 1, OW = IW + pad_x0 + pad_x1 - fW + 1;
    SHGAN_CHECK(OH >= 1 && OW >= 1, "output must be at least 1x1");
    if (const char* m = check_epi(*epi_, C)) SHGAN_CHECK(false, m);
    SHGAN_CHECK(!epi_->rgb_w, "fused torgb is not available in the FIR epilogue");
    SHGAN_CHECK(!parity_split || (!epi_->out_f32 && !epi_->skip_hi), "parity_split supports plane output only");
    SHGAN_CHECK((long long)N * C * ((long long)OH + 1) * (OW + 1) <= INT32_MAX, "tensor is too large");
    if (N == 0) return 0;
    const long long total = (long long)N * ceil_div(OH, FIR_TY) * ceil_div(OW, FIR_SX) * (C / 8);
    long long blocks = ceil_div64(total, FIR_THREADS);
    if (blocks > 148LL * 96) blocks = 148LL * 96;
    EpiParams epi = make_epi(*epi_);
    if (in_f32)
        fir4x4_nhwc_kernel<true><<<(unsigned)blocks, FIR_THREADS, 0, (cudaStream_t)stream>>>(
            in_f32, nullptr, nullptr, f, gain, N, C, IH, IW, OH, OW, pad_x0, pad_y0, epi, parity_split, total);
    else
        fir4x4_nhwc_kernel<false><<<(unsigned)blocks, FIR_THREADS, 0, (cudaStream_t)stream>>>(
            nullptr, (const __half*)in_hi, (const __half*)in_lo, f, gain, N, C, IH, IW, OH, OW, pad_x0, pad_y0, epi,
            parity_split, total);
    SHGAN_LAUNCH_CHECK();
    return 0;
}
