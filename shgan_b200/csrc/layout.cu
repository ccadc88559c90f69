// Layout conversion between the reference's NCHW fp32 tensors and the split-plane NHWC fp16 hi/lo
// activations used between the fused kernels (include/shgan_b200.h, "split planes").
//
// All four kernels are 32x32 shared-memory transposes: global reads are coalesced along the
// source's fastest axis (pixels for NCHW, channels for NHWC) and global writes along the
// destination's fastest axis.
#include "common.cuh"

namespace shgan {

constexpr int LT = 32;  // transpose tile edge

// grid: (ceil(HW/32), ceil(C/32), N), block (32, 8)
__global__ void __launch_bounds__(256)
nchw_to_planes_kernel(const float* __restrict__ x, const __half* add_hi, const __half* add_lo,
                      const float* __restrict__ scale, __half* out_hi, __half* out_lo,  // add_* may alias out_*
                      int C, int HW, int c_off, int c_tot) {
    __shared__ float tile[LT][LT + 1];
    const int n = blockIdx.z, p0 = blockIdx.x * LT, c0 = blockIdx.y * LT;
    for (int r = threadIdx.y; r < LT; r += 8) {
        const int c = c0 + r, p = p0 + threadIdx.x;
        tile[r][threadIdx.x] = (c < C && p < HW) ? __ldg(x + ((long long)n * C + c) * HW + p) : 0.f;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < LT; r += 8) {
        const int p = p0 + r, c = c0 + threadIdx.x;
        if (p < HW && c < C) {
            float v = tile[threadIdx.x][r];
            const long long o = ((long long)n * HW + p) * c_tot + c_off + c;
            if (add_hi) v += __half2float(add_hi[o]) + __half2float(add_lo[o]);
            if (scale) v *= __ldg(scale + (long long)n * C + c);
            __half h, l;
            split_f32(v, h, l);
            out_hi[o] = h;
            out_lo[o] = l;
        }
    }
}

__global__ void __launch_bounds__(256)
planes_to_nchw_kernel(const __half* __restrict__ in_hi, const __half* __restrict__ in_lo, float* __restrict__ y,
                      int C, int HW, int c_off, int c_tot) {
    __shared__ float tile[LT][LT + 1];
    const int n = blockIdx.z, p0 = blockIdx.x * LT, c0 = blockIdx.y * LT;
    for (int r = threadIdx.y; r < LT; r += 8) {
        const int p = p0 + r, c = c0 + threadIdx.x;
        float v = 0.f;
        if (p < HW && c < C) {
            const long long i = ((long long)n * HW + p) * c_tot + c_off + c;
            v = __half2float(in_hi[i]) + __half2float(in_lo[i]);
        }
        tile[r][threadIdx.x] = v;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < LT; r += 8) {
        const int c = c0 + r, p = p0 + threadIdx.x;
        if (c < C && p < HW) y[((long long)n * C + c) * HW + p] = tile[threadIdx.x][r];
    }
}

__global__ void __launch_bounds__(256)
nhwc_to_nchw_f32_kernel(const float* __restrict__ x, float* __restrict__ y, int C, int HW) {
    __shared__ float tile[LT][LT + 1];
    const int n = blockIdx.z, p0 = blockIdx.x * LT, c0 = blockIdx.y * LT;
    for (int r = threadIdx.y; r < LT; r += 8) {
        const int p = p0 + r, c = c0 + threadIdx.x;
        tile[r][threadIdx.x] = (p < HW && c < C) ? __ldg(x + ((long long)n * HW + p) * C + c) : 0.f;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < LT; r += 8) {
        const int c = c0 + r, p = p0 + threadIdx.x;
        if (c < C && p < HW) y[((long long)n * C + c) * HW + p] = tile[threadIdx.x][r];
    }
}

}  // namespace shgan

using namespace shgan;

static int check_dims(int N, int C, int H, int W, int c_off, int c_tot) {
    SHGAN_CHECK(N >= 0 && C >= 1 && H >= 1 && W >= 1, "bad tensor size");
    SHGAN_CHECK(c_off >= 0 && c_off + C <= c_tot, "channel slice out of range");
    SHGAN_CHECK((long long)N * c_tot * H * W <= INT32_MAX, "tensor is too large");
    SHGAN_CHECK(N <= 65535, "batch too large");
    return 0;
}

extern "C" int shgan_nchw_to_planes(const float* x, const void* add_hi, const void* add_lo, const float* scale,
                                    void* out_hi, void* out_lo, int N, int C, int H, int W, int c_off, int c_tot,
                                    void* stream) {
    SHGAN_CHECK(x && out_hi && out_lo, "null pointer");
    SHGAN_CHECK((add_hi == nullptr) == (add_lo == nullptr), "add_hi/add_lo must both be set");
    if (int e = check_dims(N, C, H, W, c_off, c_tot)) return e;
    if (N == 0) return 0;
    dim3 grid(ceil_div(H * W, LT), ceil_div(C, LT), N), block(LT, 8);
    nchw_to_planes_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(x, (const __half*)add_hi, (const __half*)add_lo, scale,
                                                                    (__half*)out_hi, (__half*)out_lo, C, H * W, c_off, c_tot);
    SHGAN_LAUNCH_CHECK();
    return 0;
}

extern "C" int shgan_planes_to_nchw(const void* in_hi, const void* in_lo, float* y, int N, int C, int H, int W,
                                    int c_off, int c_tot, void* stream) {
    SHGAN_CHECK(in_hi && in_lo && y, "null pointer");
    if (int e = check_dims(N, C, H, W, c_off, c_tot)) return e;
    if (N == 0) return 0;
    dim3 grid(ceil_div(H * W, LT), ceil_div(C, LT), N), block(LT, 8);
    planes_to_nchw_kernel<<<grid, block, 0, (cudaStream_t)stream>>>((const __half*)in_hi, (const __half*)in_lo, y, C, H * W,
                                                                    c_off, c_tot);
    SHGAN_LAUNCH_CHECK();
    return 0;
}

extern "C" int shgan_planes_add_nchw(void* hi, void* lo, const float* x, int N, int C, int H, int W, int c_off,
                                     int c_tot, void* stream) {
    // in-place: every element is read and written by the same thread
    return shgan_nchw_to_planes(x, hi, lo, nullptr, hi, lo, N, C, H, W, c_off, c_tot, stream);
}

// feats[r][:, c_off:c_off+C] += x_r for several resolutions in ONE launch (the SHU add-back of shgan.py:378-382 over its
// five bands): thread = (band, n, pixel, group of 8 channels); x reads are coalesced over the pixels, each thread
// re-splits its 8 channels in place.
namespace shgan {
__global__ void __launch_bounds__(256)
planes_add_multi_kernel(const shgan_add_batch b, int N, int C) {
    const int cgs = C / 8;
    long long gid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; gid < b.work_start[b.num]; gid += stride) {
        int k = 0;
        while (gid >= b.work_start[k + 1]) ++k;
        const long long loc = gid - b.work_start[k];
        const int hw = b.hw[k];
        const int p = (int)(loc % hw);
        const long long t = loc / hw;
        const int cg = (int)(t % cgs), n = (int)(t / cgs);
        const float* xp = (const float*)b.x[k] + ((long long)n * C + cg * 8) * hw + p;
        const long long idx = ((long long)n * hw + p) * b.c_tot[k] + b.c_off[k] + cg * 8;
        float v[8];
        load_planes8((const __half*)b.hi[k], (const __half*)b.lo[k], idx, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += __ldg(xp + (long long)j * hw);
        store_planes8((__half*)b.hi[k], (__half*)b.lo[k], idx, v);
    }
}
}  // namespace shgan

extern "C" int shgan_planes_add_nchw_multi(shgan_add_batch* b, int N, int C, void* stream) {
    SHGAN_CHECK(b && b->num >= 1 && b->num <= SHGAN_MAX_ADD, "bad batch");
    SHGAN_CHECK(N >= 0 && C >= 8 && C % 8 == 0, "C must be a multiple of 8");
    if (N == 0) return 0;
    b->work_start[0] = 0;
    for (int k = 0; k < b->num; ++k) {
        SHGAN_CHECK(b->hi[k] && b->lo[k] && b->x[k] && b->hw[k] >= 1, "null pointer / empty band");
        SHGAN_CHECK(b->c_off[k] >= 0 && b->c_off[k] % 8 == 0 && b->c_off[k] + C <= b->c_tot[k] && b->c_tot[k] % 8 == 0, "bad channel slice");
        b->work_start[k + 1] = b->work_start[k] + (long long)N * (C / 8) * b->hw[k];
    }
    const long long total = b->work_start[b->num];
    long long blocks = (total + 255) / 256;
    if (blocks > 148LL * 8) blocks = 148LL * 8;
    shgan::planes_add_multi_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(*b, N, C);
    SHGAN_LAUNCH_CHECK();
    return 0;
}

extern "C" int shgan_nhwc_to_nchw_f32(const float* x, float* y, int N, int C, int H, int W, void* stream) {
    SHGAN_CHECK(x && y, "null pointer");
    if (int e = check_dims(N, C, H, W, 0, C)) return e;
    if (N == 0) return 0;
    dim3 grid(ceil_div(H * W, LT), ceil_div(C, LT), N), block(LT, 8);
    nhwc_to_nchw_f32_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(x, y, C, H * W);
    SHGAN_LAUNCH_CHECK();
    return 0;
}
