// upfirdn2d forward for NCHW fp32 tensors (sm_100a).
//
// Semantics follow the reference op exactly (lib/model_zoo/stylegan_utils/upfirdn2d.py:98-138 and
// upfirdn2d.cu:97-200): zero-insert upsample, pad/crop, correlate with the (flipped unless `flip`)
// filter, keep every down-th sample, scale by gain; fp32 accumulation.
//
// Two kernels:
//  * upfirdn2d_tile_kernel<FH,FW,DOWN>  -- up == 1 fast path (the blur / blur+decimate passes that
//    carry all the traffic): one CTA stages a (TILE_H*DOWN+FH-1) x (TILE_W*DOWN+FW-1) input window of
//    one (n,c) plane in shared memory with coalesced loads, each thread produces a 2x4 register strip
//    of outputs (filter taps held in registers), 128-bit stores when the row is 16 B aligned.
//  * upfirdn2d_gather_kernel -- any up/down/filter size: one thread per output, taps that hit an
//    inserted zero are skipped by stepping the tap loop with stride `up` (polyphase).
#include "common.cuh"

namespace shgan {

constexpr int UF_TILE_W = 64;
constexpr int UF_TILE_H = 32;
constexpr int UF_THREADS = 256;

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

__global__ void __launch_bounds__(256)
upfirdn2d_gather_kernel(const float* __restrict__ x, const float* __restrict__ f, float* __restrict__ y,
                        long long total, int H, int W, int OH, int OW, int fH, int fW,
                        int upx, int upy, int downx, int downy, int padx0, int pady0, int flip, float gain) {
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int ox = (int)(idx % OW);
        long long t = idx / OW;
        const int oy = (int)(t % OH);
        const long long plane = t / OH;
        const float* xp = x + plane * (long long)H * W;
        // position in the upsampled+padded domain of filter tap (fy,fx): (oy*downy+fy, ox*downx+fx);
        // it reads input sample ((Y-pady0)/upy, (X-padx0)/upx) when divisible, else an inserted zero.
        const int Y0 = oy * downy - pady0, X0 = ox * downx - padx0;
        int fy0 = ((-Y0) % upy + upy) % upy;  // first tap row with (Y0+fy) % upy == 0
        int fx0 = ((-X0) % upx + upx) % upx;
        float acc = 0.f;
        for (int fy = fy0; fy < fH; fy += upy) {
            const int Yn = Y0 + fy;
            const int iy = Yn >= 0 ? Yn / upy : -1;
            if (iy < 0 || iy >= H) continue;
            for (int fx = fx0; fx < fW; fx += upx) {
                const int Xn = X0 + fx;
                const int ix = Xn >= 0 ? Xn / upx : -1;
                if (ix < 0 || ix >= W) continue;
                const float fv = flip ? f[fy * fW + fx] : f[(fH - 1 - fy) * fW + (fW - 1 - fx)];
                acc = fmaf(__ldg(xp + (long long)iy * W + ix), fv, acc);
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

    const bool fast = (upx == 1 && upy == 1 && downx == downy && (downx == 1 || downx == 2) && fH == 4 && fW == 4);
    if (fast) {
        const int tiles_x = ceil_div(OW, UF_TILE_W), tiles_y = ceil_div(OH, UF_TILE_H);
        const long long grid = (long long)tiles_x * tiles_y * N * C;
        SHGAN_CHECK(grid <= INT32_MAX, "grid too large");
        if (downx == 1)
            upfirdn2d_tile_kernel<4, 4, 1><<<(unsigned)grid, UF_THREADS, 0, stream>>>(x, f, y, H, W, OH, OW, padx0, pady0, flip, gain, tiles_x, tiles_y);
        else
            upfirdn2d_tile_kernel<4, 4, 2><<<(unsigned)grid, UF_THREADS, 0, stream>>>(x, f, y, H, W, OH, OW, padx0, pady0, flip, gain, tiles_x, tiles_y);
        SHGAN_LAUNCH_CHECK();
        return 0;
    }
    const int threads = 256;
    long long blocks = ceil_div64(total, threads);
    if (blocks > 148LL * 32) blocks = 148LL * 32;
    upfirdn2d_gather_kernel<<<(unsigned)blocks, threads, 0, stream>>>(x, f, y, total, H, W, OH, OW, fH, fW, upx, upy, downx, downy, padx0, pady0, flip, gain);
    SHGAN_LAUNCH_CHECK();
    return 0;
}
