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
#include "tma_util.cuh"

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
                    if (x < OW && !(parity_split == 2 && ((yo | x) & 1))) {
                        long long out_pix = ((long long)n * OH + yo) * OW + x;
                        if (parity_split == 1) {
                            const int q = (yo & 1) * 2 + (x & 1);
                            out_pix = (long long)q * N * PH * PW + ((long long)n * PH + (yo >> 1)) * PW + (x >> 1);
                        } else if (parity_split == 2) {
                            out_pix = ((long long)n * PH + (yo >> 1)) * PW + (x >> 1);
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


// ---------------------------------------------------------------------------------------------------------------------
// TMA-staged variant (default).  The sliding-window kernel above is bound by exposed global-load latency (ncu: the top
// stall is long_scoreboard at ~18 % occupancy, 2-3 TB/s).  Here a persistent CTA walks output tiles of
// FT_W x FT_H pixels x FT_C channels: one elected thread prefetches the NEXT tile's (FT_W+3) x (FT_H+3) input window with
// TMA (zero-filled outside the image = the blur's zero padding) into the other half of a 2-stage shared-memory ring
// while all threads filter the current one out of shared memory, so HBM latency is covered by the copy engine rather
// than by occupancy.  Each thread owns 1 pixel column x 8 channels and slides down the tile rows.
constexpr int FT_W = 32, FT_H = 8, FT_C = 32;
constexpr int FT_IW = FT_W + FIR_T - 1, FT_IH = FT_H + FIR_T - 1;
constexpr int FT_THREADS = FT_W * (FT_C / 8);                       // 128
constexpr int FT_PLANE_BYTES = FT_IH * FT_IW * FT_C * 2;            // one fp16 plane of one stage
constexpr int FT_PLANE_STRIDE = ((FT_PLANE_BYTES + 127) / 128) * 128;   // TMA destinations are 128 B aligned
constexpr int FT_STAGE_BYTES = 2 * FT_PLANE_STRIDE;                     // hi+lo planes, or one fp32 tile of the same bytes
constexpr int FT_SMEM_BYTES = 2 * FT_STAGE_BYTES + 128 + 64;

struct FirMaps {
    CUtensorMap a;   // planes: hi   | fp32 input: the tensor
    CUtensorMap b;   // planes: lo   | unused
};

struct FirTiles {
    int tiles_x, tiles_y, tiles_c, total;
};

template <bool IN_F32>
__global__ void __launch_bounds__(FT_THREADS, 2)
fir4x4_tma_kernel(const __grid_constant__ FirMaps maps, const float* __restrict__ f, float gain, int N, int C, int OH, int OW,
                  int pad_x0, int pad_y0, EpiParams epi, int parity_split, FirTiles ft) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + 2 * FT_STAGE_BYTES);
    __shared__ float s_f[FIR_T * FIR_T];
    if (threadIdx.x < FIR_T * FIR_T) s_f[threadIdx.x] = f[threadIdx.x] * gain;
    if (threadIdx.x == 0) {
        prefetch_tmap(&maps.a);
        if (!IN_F32) prefetch_tmap(&maps.b);
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    float fk[FIR_T][FIR_T];
#pragma unroll
    for (int i = 0; i < FIR_T; ++i)
#pragma unroll
        for (int j = 0; j < FIR_T; ++j) fk[i][j] = s_f[i * FIR_T + j];

    const int cg = threadIdx.x & (FT_C / 8 - 1);     // 8-channel group inside the tile's channel block
    const int px = threadIdx.x / (FT_C / 8);         // pixel column inside the tile
    const int PH = (OH + 1) / 2, PW = (OW + 1) / 2;

    auto issue = [&](int tile, int stage) {
        int t = tile;
        const int cb = t % ft.tiles_c; t /= ft.tiles_c;
        const int tx = t % ft.tiles_x; t /= ft.tiles_x;
        const int ty = t % ft.tiles_y;
        const int n = t / ft.tiles_y;
        uint8_t* dst = smem + stage * FT_STAGE_BYTES;
        const int cx = tx * FT_W - pad_x0, cy = ty * FT_H - pad_y0;
        if (IN_F32) {
            mbar_expect_tx(&full[stage], 2 * FT_PLANE_BYTES);
            tma_load_4d(dst, &maps.a, &full[stage], cb * FT_C, cx, cy, n);
        } else {
            mbar_expect_tx(&full[stage], 2 * FT_PLANE_BYTES);
            tma_load_4d(dst, &maps.a, &full[stage], cb * FT_C, cx, cy, n);
            tma_load_4d(dst + FT_PLANE_STRIDE, &maps.b, &full[stage], cb * FT_C, cx, cy, n);
        }
    };

    if (threadIdx.x == 0 && (int)blockIdx.x < ft.total) issue(blockIdx.x, 0);
    int it = 0;
    for (int tile = blockIdx.x; tile < ft.total; tile += gridDim.x, ++it) {
        const int stage = it & 1;
        if (threadIdx.x == 0 && tile + (int)gridDim.x < ft.total) issue(tile + gridDim.x, stage ^ 1);
        int t = tile;
        const int cb = t % ft.tiles_c; t /= ft.tiles_c;
        const int tx = t % ft.tiles_x; t /= ft.tiles_x;
        const int ty = t % ft.tiles_y;
        const int n = t / ft.tiles_y;
        const int x = tx * FT_W + px, y0 = ty * FT_H, c0 = cb * FT_C + cg * 8;
        mbar_wait(&full[stage], (it >> 1) & 1);
        const uint8_t* sbase = smem + stage * FT_STAGE_BYTES;

        float acc[FIR_T][8];
#pragma unroll
        for (int a = 0; a < FIR_T; ++a)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[a][j] = 0.f;
#pragma unroll 1
        for (int jrow = 0; jrow < FT_IH; ++jrow) {
            float row[FIR_T][8];
#pragma unroll
            for (int c = 0; c < FIR_T; ++c) {
                const int pix = jrow * FT_IW + px + c;
                if (IN_F32) {
                    const float4* sp = reinterpret_cast<const float4*>(sbase + ((size_t)pix * FT_C + cg * 8) * 4);
                    const float4 a = sp[0], b = sp[1];
                    row[c][0] = a.x; row[c][1] = a.y; row[c][2] = a.z; row[c][3] = a.w;
                    row[c][4] = b.x; row[c][5] = b.y; row[c][6] = b.z; row[c][7] = b.w;
                } else {
                    const uint4 h = *reinterpret_cast<const uint4*>(sbase + ((size_t)pix * FT_C + cg * 8) * 2);
                    const uint4 l = *reinterpret_cast<const uint4*>(sbase + FT_PLANE_STRIDE + ((size_t)pix * FT_C + cg * 8) * 2);
                    const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float2 fa = unpack_h2(hw[i]), fb = unpack_h2(lw[i]);
                        row[c][2 * i] = fa.x + fb.x;
                        row[c][2 * i + 1] = fa.y + fb.y;
                    }
                }
            }
#pragma unroll
            for (int sl = 0; sl < FIR_T; ++sl)
#pragma unroll
                for (int fx = 0; fx < FIR_T; ++fx)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[sl][j] = fmaf(row[fx][j], fk[FIR_T - 1 - sl][fx], acc[sl][j]);
            const int yo = y0 + jrow - (FIR_T - 1);
            if (jrow >= FIR_T - 1 && yo < OH && x < OW && !(parity_split == 2 && ((yo | x) & 1))) {
                long long out_pix = ((long long)n * OH + yo) * OW + x;
                if (parity_split == 1) {
                    const int q = (yo & 1) * 2 + (x & 1);
                    out_pix = (long long)q * N * PH * PW + ((long long)n * PH + (yo >> 1)) * PW + (x >> 1);
                } else if (parity_split == 2) {
                    out_pix = ((long long)n * PH + (yo >> 1)) * PW + (x >> 1);
                }
                float rgb[3] = {0.f, 0.f, 0.f};
                epilogue_apply<8>(epi, acc[0], n, yo, x, OH, OW, C, c0, rgb, out_pix);
            }
#pragma unroll
            for (int sl = 0; sl < FIR_T - 1; ++sl)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[sl][j] = acc[sl + 1][j];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[FIR_T - 1][j] = 0.f;
        }
        __syncthreads();   // every thread is done with this stage before the next-but-one TMA overwrites it
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
    SHGAN_CHECK(parity_split >= 0 && parity_split <= 2, "parity_split must be 0, 1 or 2");
    SHGAN_CHECK(!parity_split || (!epi_->out_f32 && !epi_->skip_hi && !epi_->noise), "parity_split supports plane output only");
    SHGAN_CHECK((long long)N * C * ((long long)OH + 1) * (OW + 1) <= INT32_MAX, "tensor is too large");
    if (N == 0) return 0;
    EpiParams epi = make_epi(*epi_);
    // measured on B200 (tools/microbench.py, batch 16): planes input 2.3-2.5 TB/s with the TMA-staged kernel vs 2.0 TB/s with the
    // register kernel; fp32 input + full epilogue 2.0 TB/s vs 3.1 TB/s (its exposed skip/parameter loads want occupancy)
    if (C % FT_C == 0 && !in_f32) {
        // TMA-staged kernel
        static bool attr_set = false;
        static int num_sms = 148;
        if (!attr_set) {
            SHGAN_CUDA(cudaFuncSetAttribute(fir4x4_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_SMEM_BYTES));
            SHGAN_CUDA(cudaFuncSetAttribute(fir4x4_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_SMEM_BYTES));
            int dev = 0;
            SHGAN_CUDA(cudaGetDevice(&dev));
            SHGAN_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
            attr_set = true;
        }
        FirTiles ft;
        ft.tiles_x = ceil_div(OW, FT_W); ft.tiles_y = ceil_div(OH, FT_H); ft.tiles_c = C / FT_C;
        const long long tot = (long long)ft.tiles_x * ft.tiles_y * ft.tiles_c * N;
        SHGAN_CHECK(tot <= INT32_MAX, "too many tiles");
        ft.total = (int)tot;
        FirMaps maps;
        const uint64_t dims[4] = {(uint64_t)C, (uint64_t)IW, (uint64_t)IH, (uint64_t)N};
        const uint32_t box[4] = {(uint32_t)FT_C, (uint32_t)FT_IW, (uint32_t)FT_IH, 1u};
        if (in_f32) {
            if (int e = encode_tmap(&maps.a, in_f32, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, 4, dims, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return e;
            maps.b = maps.a;
        } else {
            if (int e = encode_tmap(&maps.a, in_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, 4, dims, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return e;
            if (int e = encode_tmap(&maps.b, in_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, 4, dims, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return e;
        }
        const int grid = ft.total < 2 * num_sms ? ft.total : 2 * num_sms;
        if (in_f32)
            fir4x4_tma_kernel<true><<<grid, FT_THREADS, FT_SMEM_BYTES, (cudaStream_t)stream>>>(maps, f, gain, N, C, OH, OW, pad_x0, pad_y0,
                                                                                             epi, parity_split, ft);
        else
            fir4x4_tma_kernel<false><<<grid, FT_THREADS, FT_SMEM_BYTES, (cudaStream_t)stream>>>(maps, f, gain, N, C, OH, OW, pad_x0, pad_y0,
                                                                                              epi, parity_split, ft);
        SHGAN_LAUNCH_CHECK();
        return 0;
    }
    // channel counts that are not a multiple of 32: register sliding-window kernel
    const long long total = (long long)N * ceil_div(OH, FIR_TY) * ceil_div(OW, FIR_SX) * (C / 8);
    long long blocks = ceil_div64(total, FIR_THREADS);
    if (blocks > 148LL * 96) blocks = 148LL * 96;
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
