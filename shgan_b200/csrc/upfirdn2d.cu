// upfirdn2d forward for NCHW fp32 tensors (sm_100a).
//
// Semantics follow the reference op exactly (lib/model_zoo/stylegan_utils/upfirdn2d.py:98-138 and
// upfirdn2d.cu:97-200): zero-insert upsample, pad/crop, correlate with the (flipped unless `flip`)
// filter, keep every down-th sample, scale by gain; fp32 accumulation.
//
// Three kernels:
//  * upfirdn2d_sep4_kernel -- up == down == 1 with a 4x4 filter that is an outer product (the reference's
//    setup_filter([1,3,3,1]) is; decided on the device, no host sync): the blur passes that carry all the
//    traffic.  A warp owns 128 output columns of one (n,c) plane and walks down a segment of 64 rows: each
//    input row is read ONCE, requested seven rows ahead by fully coalesced 4-byte cp.async copies straight into a
//    per-warp shared-memory ring that doubles as the staged row, so that a lane gets the 7 inputs of its 4
//    outputs as two 128-bit reads, filtered horizontally (16 FMA) and vertically over a 4-row ring in
//    registers (16 FMA), and stored as one 128-bit vector: 8 FMA per output instead of 16, no block barrier.
//  * upfirdn2d_tile_kernel<FH,FW,DOWN>  -- up == 1 with any other 4x4 filter, and the blur+decimate
//    passes: one CTA stages a (TILE_H*DOWN+FH-1) x (TILE_W*DOWN+FW-1) input window of
//    one (n,c) plane in shared memory with coalesced loads, each thread produces a 2x4 register strip
//    of outputs (filter taps held in registers), 128-bit stores when the row is 16 B aligned.
//  * upfirdn2d_gather_kernel -- any up/down/filter size: one thread per output, taps that hit an
//    inserted zero are skipped by stepping the tap loop with stride `up` (polyphase); 32-bit index
//    arithmetic (the host checks that every tensor has fewer than 2^31 elements).
#include "common.cuh"

namespace shgan {

constexpr int UF_TILE_W = 64;
constexpr int UF_TILE_H = 32;
constexpr int UF_THREADS = 256;

// the filter as applied (correlation order: flipped unless `flip`, upfirdn2d.py:122-124) times gain, and its rank-1
// factorisation a = u (x) v around the largest tap; false when |a - u (x) v| > 1e-6 max|a|
__device__ __forceinline__ bool uf_factorise4(const float* __restrict__ f, int flip, float gain, float* u, float* v) {
    float a[16];
    int bi = 0;
    float best = -1.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        a[i] = __ldg(f + (flip ? i : 15 - i)) * gain;
        if (fabsf(a[i]) > best) { best = fabsf(a[i]); bi = i; }
    }
    const int pr = bi >> 2, pc = bi & 3;
    const float piv = a[bi];
    float resid = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = a[pr * 4 + j];
#pragma unroll
    for (int i = 0; i < 4; ++i) u[i] = piv != 0.f ? a[i * 4 + pc] / piv : 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) resid = fmaxf(resid, fabsf(a[i * 4 + j] - u[i] * v[j]));
    return resid <= 1e-6f * best;
}

constexpr int US_SEG = 64;        // output rows per work item
constexpr int US_COLS = 128;      // output columns per work item (4 per lane)
constexpr int US_ROWBUF = 136;    // floats per staged input row: 128 + 3 halo, padded to a multiple of 4
constexpr int US_STAGES = 8;      // input rows in flight per warp (cp.async ring; the ring IS the staged row: fp32 in, fp32 out)
constexpr int US_SMEM_BYTES = 8 * US_STAGES * US_ROWBUF * 4;      // 8 warps

__device__ __forceinline__ void us_cp_async4(uint32_t dst, const void* src, bool valid) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(valid ? 4 : 0) : "memory");
}

__global__ void __launch_bounds__(256)
upfirdn2d_sep4_kernel(const float* __restrict__ x, const float* __restrict__ f, float* __restrict__ y, int H, int W, int OH, int OW,
                      int padx0, int pady0, int flip, float gain, int xblocks, int segs, int items) {
    extern __shared__ __align__(16) float us_ring[];       // [8 warps][US_STAGES][US_ROWBUF]
    float u[4], v[4];
    const bool sep = uf_factorise4(f, flip, gain, u, v);   // warp-uniform (same filter for every thread)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int item = blockIdx.x * 8 + warp;
    if (item >= items) return;
    const int xb = item % xblocks;
    int t = item / xblocks;
    const int seg = t % segs;
    const int plane = t / segs;
    const float* xp = x + (size_t)plane * H * W;
    float* yp = y + (size_t)plane * OH * OW;
    const int oy0 = seg * US_SEG, oy1 = min(OH, oy0 + US_SEG);
    const int ox = xb * US_COLS + lane * 4;                 // my first output column
    const int ixb = xb * US_COLS - padx0;                   // input column of ring[.][0]
    float* ring = us_ring + warp * (US_STAGES * US_ROWBUF);
    const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(ring);
    // one input row -> ring stage, straight from global memory with fully coalesced 4-byte copies (rows of odd widths are not
    // 16-byte aligned), zero-filled outside the image: columns ixb + k*32 + lane (k < 4) and the 3 halo columns (lane < 3)
    auto request_row = [&](int iy, int stage) {
        const bool rv = iy >= 0 && iy < H;
        const float* rp = xp + (size_t)(rv ? iy : 0) * W;
        const uint32_t d = ring_s + (uint32_t)(stage * US_ROWBUF) * 4u;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int ix = ixb + k * 32 + lane;
            const bool ok = rv && ix >= 0 && ix < W;
            us_cp_async4(d + (uint32_t)(k * 32 + lane) * 4u, rp + (ok ? ix : 0), ok);
        }
        if (lane < 3) {
            const int ix = ixb + 128 + lane;
            const bool ok = rv && ix >= 0 && ix < W;
            us_cp_async4(d + (uint32_t)(128 + lane) * 4u, rp + (ok ? ix : 0), ok);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // row r of the walk (r = 0 is input row oy0 - pady0) lives in stage r % US_STAGES; rows up to r + US_STAGES - 2 have been
    // requested when row r is consumed.  The warp barrier makes every lane's copies of row r visible to the others and proves
    // that all lanes are done with row r - 1, whose stage the next request overwrites.
    auto acquire = [&](int r, int iy_base) -> const float* {
        asm volatile("cp.async.wait_group %0;" ::"n"(US_STAGES - 2) : "memory");
        __syncwarp();
        request_row(iy_base + r + US_STAGES - 1, (r + US_STAGES - 1) % US_STAGES);
        return ring + (r % US_STAGES) * US_ROWBUF + lane * 4;        // my 7 inputs: [0..6]
    };
    const bool vec_ok = (OW & 3) == 0 && ((reinterpret_cast<uintptr_t>(y) & 15) == 0);
    auto store4 = [&](int oy, const float4& o) {
        if (ox >= OW) return;
        float* dst = yp + (size_t)oy * OW + ox;
        if (vec_ok && ox + 3 < OW) {
            *reinterpret_cast<float4*>(dst) = o;
        } else {
            dst[0] = o.x;
            if (ox + 1 < OW) dst[1] = o.y;
            if (ox + 2 < OW) dst[2] = o.z;
            if (ox + 3 < OW) dst[3] = o.w;
        }
    };
    const int iy0 = oy0 - pady0;
#pragma unroll
    for (int r = 0; r < US_STAGES - 1; ++r) request_row(iy0 + r, r);
    if (!sep) {
        // any other 4x4 filter: same walk, the ring of registers holds the last four RAW rows (7 inputs each), 16 taps per output
        float a[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = __ldg(f + (flip ? i : 15 - i)) * gain;
        auto fetch = [&](int r, float (&w)[7]) {
            const float* s = acquire(r, iy0);
            const float4 p = *reinterpret_cast<const float4*>(s), q = *reinterpret_cast<const float4*>(s + 4);
            w[0] = p.x; w[1] = p.y; w[2] = p.z; w[3] = p.w; w[4] = q.x; w[5] = q.y; w[6] = q.z;
        };
        float w0[7], w1[7], w2[7], w3[7];
        fetch(0, w0); fetch(1, w1); fetch(2, w2);
        int r = 3;
        for (int oy = oy0; oy < oy1; ++oy, ++r) {
            fetch(r, w3);
            float o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float acc = 0.f;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    acc = fmaf(a[k], w0[j + k], acc);
                    acc = fmaf(a[4 + k], w1[j + k], acc);
                    acc = fmaf(a[8 + k], w2[j + k], acc);
                    acc = fmaf(a[12 + k], w3[j + k], acc);
                }
                o[j] = acc;
            }
            store4(oy, make_float4(o[0], o[1], o[2], o[3]));
#pragma unroll
            for (int k = 0; k < 7; ++k) { w0[k] = w1[k]; w1[k] = w2[k]; w2[k] = w3[k]; }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        return;
    }
    // separable: horizontal 4-tap pass of each row (my 4 outputs read ring[lane*4 .. lane*4+6]), vertical pass over a ring of the
    // last four horizontal rows in registers: output row oy = sum_k u[k] * hrow(oy - pady0 + k)
    auto hpass = [&](int r) -> float4 {
        const float* s = acquire(r, iy0);
        const float4 a = *reinterpret_cast<const float4*>(s), c = *reinterpret_cast<const float4*>(s + 4);
        float4 h;
        h.x = fmaf(v[3], a.w, fmaf(v[2], a.z, fmaf(v[1], a.y, v[0] * a.x)));
        h.y = fmaf(v[3], c.x, fmaf(v[2], a.w, fmaf(v[1], a.z, v[0] * a.y)));
        h.z = fmaf(v[3], c.y, fmaf(v[2], c.x, fmaf(v[1], a.w, v[0] * a.z)));
        h.w = fmaf(v[3], c.z, fmaf(v[2], c.y, fmaf(v[1], c.x, v[0] * a.w)));
        return h;
    };
    float4 h0 = hpass(0), h1 = hpass(1), h2 = hpass(2), h3;
    int r = 3;
    for (int oy = oy0; oy < oy1; ++oy, ++r) {
        h3 = hpass(r);
        float4 o;
        o.x = fmaf(u[3], h3.x, fmaf(u[2], h2.x, fmaf(u[1], h1.x, u[0] * h0.x)));
        o.y = fmaf(u[3], h3.y, fmaf(u[2], h2.y, fmaf(u[1], h1.y, u[0] * h0.y)));
        o.z = fmaf(u[3], h3.z, fmaf(u[2], h2.z, fmaf(u[1], h1.z, u[0] * h0.z)));
        o.w = fmaf(u[3], h3.w, fmaf(u[2], h2.w, fmaf(u[1], h1.w, u[0] * h0.w)));
        store4(oy, o);
        h0 = h1; h1 = h2; h2 = h3;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

template <int FH, int FW, int DOWN>
__global__ void __launch_bounds__(UF_THREADS)
upfirdn2d_tile_kernel(const float* __restrict__ x, const float* __restrict__ f, float* __restrict__ y,
                      int H, int W, int OH, int OW, int padx0, int pady0, int flip, float gain,
                      int tiles_x, int tiles_y) {
    constexpr int IN_W = UF_TILE_W * DOWN + FW - 1;
    constexpr int IN_H = UF_TILE_H * DOWN + FH - 1;
    constexpr int PITCH = (IN_W + 3) & ~3;
    __shared__ float s_in[IN_H * PITCH];
    __shared__ float s_f[FH * FW];

    int tile = blockIdx.x;
    const int tx_i = tile % tiles_x; tile /= tiles_x;
    const int ty_i = tile % tiles_y; tile /= tiles_y;
    const long long plane = tile;  // n*C + c
    const float* xp = x + plane * (long long)H * W;
    float* yp = y + plane * (long long)OH * OW;

    const int ox0 = tx_i * UF_TILE_W, oy0 = ty_i * UF_TILE_H;
    const int ix0 = ox0 * DOWN - padx0, iy0 = oy0 * DOWN - pady0;

    if (threadIdx.x < FH * FW) {
        int fy = threadIdx.x / FW, fx = threadIdx.x % FW;
        // correlation with the flipped filter unless `flip` (upfirdn2d.py:122-124)
        s_f[threadIdx.x] = flip ? f[fy * FW + fx] : f[(FH - 1 - fy) * FW + (FW - 1 - fx)];
    }
    for (int i = threadIdx.x; i < IN_H * PITCH; i += UF_THREADS) {
        int r = i / PITCH, c = i - r * PITCH;
        int iy = iy0 + r, ix = ix0 + c;
        float v = 0.f;
        if (c < IN_W && iy >= 0 && iy < H && ix >= 0 && ix < W) v = __ldg(xp + (long long)iy * W + ix);
        s_in[i] = v;
    }
    __syncthreads();

    float fk[FH][FW];
#pragma unroll
    for (int i = 0; i < FH; ++i)
#pragma unroll
        for (int j = 0; j < FW; ++j) fk[i][j] = s_f[i * FW + j] * gain;

    // thread -> 2 rows x 4 cols strip
    const int sx = (threadIdx.x % (UF_TILE_W / 4)) * 4;
    const int sy = (threadIdx.x / (UF_TILE_W / 4)) * 2;
    constexpr int RW = 3 * DOWN + FW;   // input cols needed by 4 outputs
    constexpr int RH = 1 * DOWN + FH;   // input rows needed by 2 outputs
    float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
    for (int r = 0; r < RH; ++r) {
        float row[RW];
        const float* sp = s_in + (sy * DOWN + r) * PITCH + sx * DOWN;
#pragma unroll
        for (int c = 0; c < RW; ++c) row[c] = sp[c];
#pragma unroll
        for (int oy = 0; oy < 2; ++oy) {
            const int fy = r - oy * DOWN;
            if (fy >= 0 && fy < FH) {
#pragma unroll
                for (int ox = 0; ox < 4; ++ox)
#pragma unroll
                    for (int fx = 0; fx < FW; ++fx) acc[oy][ox] = fmaf(row[ox * DOWN + fx], fk[fy][fx], acc[oy][ox]);
            }
        }
    }
#pragma unroll
    for (int oy = 0; oy < 2; ++oy) {
        const int gy = oy0 + sy + oy, gx = ox0 + sx;
        if (gy >= OH || gx >= OW) continue;
        float* dst = yp + (long long)gy * OW + gx;
        if (gx + 3 < OW && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
            *reinterpret_cast<float4*>(dst) = make_float4(acc[oy][0], acc[oy][1], acc[oy][2], acc[oy][3]);
        } else {
#pragma unroll
            for (int ox = 0; ox < 4; ++ox)
                if (gx + ox < OW) dst[ox] = acc[oy][ox];
        }
    }
}

// up == 2, down == 1, 4x4 filter (upsample2d of the image, comodgan.py:331-338): polyphase -- an output row takes 2 of the 4
// filter rows and every output 2 of the 4 taps of a row.  Thread = 4 adjacent outputs of one row: 2 input rows x 4 input
// columns, 16 FMA, one 128-bit store.  E = parity of the first output's position in the zero-inserted domain (X0 & 1; the
// same for every thread because a thread's first output column is a multiple of 4).
template <int E>
__global__ void __launch_bounds__(256)
upfirdn2d_up2_kernel(const float* __restrict__ x, const float* __restrict__ f, float* __restrict__ y, unsigned total4, int H, int W,
                     int OH, int OW, int padx0, int pady0, int flip, float gain) {
    __shared__ float s_a[16];
    if (threadIdx.x < 16) s_a[threadIdx.x] = __ldg(f + (flip ? threadIdx.x : 15 - threadIdx.x)) * gain;
    __syncthreads();
    const unsigned ow4 = (unsigned)(OW + 3) >> 2;
    const bool vec_ok = (OW & 3) == 0 && ((reinterpret_cast<uintptr_t>(y) & 15) == 0);
    const unsigned stride = gridDim.x * blockDim.x;
    for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total4; idx += stride) {
        const unsigned t = idx / ow4;
        const int ox = (int)(idx - t * ow4) * 4;
        const unsigned plane = t / (unsigned)OH;
        const int oy = (int)(t - plane * (unsigned)OH);
        const float* xp = x + (size_t)plane * H * W;
        const int X0 = ox - padx0, Y0 = oy - pady0;
        const int fb = (X0 - E) >> 1;                  // input column of c[0]: zero-inserted position X0 - E (even)
        const int q = Y0 & 1;                          // first filter row that hits a real sample
        float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int fy = q + 2 * r;
            const int iy = (Y0 + fy) >> 1;             // exact: Y0 + fy is even
            if (iy < 0 || iy >= H) continue;
            float c[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) c[k] = (fb + k >= 0 && fb + k < W) ? __ldg(xp + iy * W + fb + k) : 0.f;
            const float a0 = s_a[fy * 4], a1 = s_a[fy * 4 + 1], a2 = s_a[fy * 4 + 2], a3 = s_a[fy * 4 + 3];
            const float a[4] = {a0, a1, a2, a3};
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int fx = 0; fx < 4; ++fx)
                    if (((j + fx + E) & 1) == 0) o[j] = fmaf(a[fx], c[(j + fx + E) >> 1], o[j]);
        }
        float* dst = y + (size_t)t * OW + ox;
        if (vec_ok) {
            *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (ox + j < OW) dst[j] = o[j];
        }
    }
}

__global__ void __launch_bounds__(256)
upfirdn2d_gather_kernel(const float* __restrict__ x, const float* __restrict__ f, float* __restrict__ y,
                        unsigned total, int H, int W, int OH, int OW, int fH, int fW,
                        int upx, int upy, int downx, int downy, int padx0, int pady0, int flip, float gain) {
    // every tensor has fewer than 2^31 elements (checked on the host): 32-bit index arithmetic throughout (64-bit divisions were
    // most of this kernel's instructions)
    const unsigned stride = gridDim.x * blockDim.x;
    for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        const unsigned t = idx / (unsigned)OW;
        const int ox = (int)(idx - t * (unsigned)OW);
        const unsigned plane = t / (unsigned)OH;
        const int oy = (int)(t - plane * (unsigned)OH);
        const float* xp = x + (size_t)plane * H * W;
        // position in the upsampled+padded domain of filter tap (fy,fx): (oy*downy+fy, ox*downx+fx);
        // it reads input sample ((Y-pady0)/upy, (X-padx0)/upx) when divisible, else an inserted zero.
        const int Y0 = oy * downy - pady0, X0 = ox * downx - padx0;
        const int fy0 = ((-Y0) % upy + upy) % upy;  // first tap row with (Y0+fy) % upy == 0
        const int fx0 = ((-X0) % upx + upx) % upx;
        float acc = 0.f;
        int iy = (Y0 + fy0) >= 0 ? (Y0 + fy0) / upy : -((-(Y0 + fy0)) / upy);     // exact: Y0 + fy0 is a multiple of upy
        for (int fy = fy0; fy < fH; fy += upy, ++iy) {
            if (iy < 0 || iy >= H) continue;
            int ix = (X0 + fx0) >= 0 ? (X0 + fx0) / upx : -((-(X0 + fx0)) / upx);
            for (int fx = fx0; fx < fW; fx += upx, ++ix) {
                if (ix < 0 || ix >= W) continue;
                const float fv = flip ? __ldg(f + fy * fW + fx) : __ldg(f + (fH - 1 - fy) * fW + (fW - 1 - fx));
                acc = fmaf(__ldg(xp + iy * W + ix), fv, acc);
            }
        }
        y[idx] = acc * gain;
    }
}

}  // namespace shgan

using namespace shgan;

extern "C" int shgan_upfirdn2d_fwd(const float* x, const float* f, float* y, int N, int C, int H, int W, int fH,
                                   int fW, int upx, int upy, int downx, int downy, int padx0, int padx1, int pady0,
                                   int pady1, int flip, float gain, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    // argument validation mirrors the TORCH_CHECKs of upfirdn2d.cpp:19-36
    SHGAN_CHECK(f && ((x && y) || N == 0 || C == 0), "null pointer");
    SHGAN_CHECK(N >= 0 && C >= 0 && H >= 1 && W >= 1, "bad input size");
    SHGAN_CHECK(fH >= 1 && fW >= 1, "f must be at least 1x1");
    SHGAN_CHECK(upx >= 1 && upy >= 1, "upsampling factor must be at least 1");
    SHGAN_CHECK(downx >= 1 && downy >= 1, "downsampling factor must be at least 1");
    const int OW = (W * upx + padx0 + padx1 - fW + downx) / downx;
    const int OH = (H * upy + pady0 + pady1 - fH + downy) / downy;
    SHGAN_CHECK(OW >= 1 && OH >= 1, "output must be at least 1x1");
    const long long total = (long long)N * C * OH * OW;
    SHGAN_CHECK((long long)N * C * H * W <= INT32_MAX && total <= INT32_MAX, "tensor is too large");
    if (total == 0) return 0;

    if (upx == 1 && upy == 1 && downx == 1 && downy == 1 && fH == 4 && fW == 4) {
        // the blur passes: row-walking kernel (separable filters take its 8-FMA path, decided on the device)
        const int xblocks = ceil_div(OW, US_COLS), segs = ceil_div(OH, US_SEG);
        const long long items = (long long)xblocks * segs * N * C;
        SHGAN_CHECK(items <= INT32_MAX, "grid too large");
        upfirdn2d_sep4_kernel<<<(unsigned)ceil_div64(items, 8), 256, US_SMEM_BYTES, stream>>>(x, f, y, H, W, OH, OW, padx0, pady0, flip, gain, xblocks,
                                                                                  segs, (int)items);
        SHGAN_LAUNCH_CHECK();
        return 0;
    }
    if (upx == 1 && upy == 1 && downx == 2 && downy == 2 && fH == 4 && fW == 4) {
        const int tiles_x = ceil_div(OW, UF_TILE_W), tiles_y = ceil_div(OH, UF_TILE_H);
        const long long grid = (long long)tiles_x * tiles_y * N * C;
        SHGAN_CHECK(grid <= INT32_MAX, "grid too large");
        upfirdn2d_tile_kernel<4, 4, 2><<<(unsigned)grid, UF_THREADS, 0, stream>>>(x, f, y, H, W, OH, OW, padx0, pady0, flip, gain, tiles_x, tiles_y);
        SHGAN_LAUNCH_CHECK();
        return 0;
    }
    if (upx == 2 && upy == 2 && downx == 1 && downy == 1 && fH == 4 && fW == 4) {
        const long long total4 = (long long)N * C * OH * ((OW + 3) / 4);
        long long blocks = ceil_div64(total4, 256);
        if (blocks > 148LL * 32) blocks = 148LL * 32;
        if ((padx0 & 1) == 0)
            upfirdn2d_up2_kernel<0><<<(unsigned)blocks, 256, 0, stream>>>(x, f, y, (unsigned)total4, H, W, OH, OW, padx0, pady0, flip, gain);
        else
            upfirdn2d_up2_kernel<1><<<(unsigned)blocks, 256, 0, stream>>>(x, f, y, (unsigned)total4, H, W, OH, OW, padx0, pady0, flip, gain);
        SHGAN_LAUNCH_CHECK();
        return 0;
    }
    const int threads = 256;
    long long blocks = ceil_div64(total, threads);
    if (blocks > 148LL * 32) blocks = 148LL * 32;
    upfirdn2d_gather_kernel<<<(unsigned)blocks, threads, 0, stream>>>(x, f, y, (unsigned)total, H, W, OH, OW, fH, fW, upx, upy, downx, downy, padx0, pady0, flip, gain);
    SHGAN_LAUNCH_CHECK();
    return 0;
}
