// Geometry shared by the two implementations of shgan_conv_igemm (conv_tc.cu: tcgen05 tensor-core
// kernel, the product path; conv_simt.cu: plain fp32 FMA kernel kept as the on-device cross-check).
#pragma once
#include "common.cuh"

namespace shgan {

struct ConvGeom {
    int num_src;
    const __half* src_hi[SHGAN_MAX_SRC];
    const __half* src_lo[SHGAN_MAX_SRC];
    int src_h[SHGAN_MAX_SRC], src_w[SHGAN_MAX_SRC];
    int N, C, Co;
    const __half* w_hi;
    const __half* w_lo;
    int ntaps;
    int tap_src[SHGAN_MAX_TAPS], tap_dy[SHGAN_MAX_TAPS], tap_dx[SHGAN_MAX_TAPS], tap_w[SHGAN_MAX_TAPS];
    int OH, OW;
    int mode;  // 0: epilogue, 1: raw fp32 scatter into z
    float* z;
    int ZH, ZW, zsy, zsx, zoy, zox;
};

static inline ConvGeom make_geom(const shgan_conv_desc& d) {
    ConvGeom g;
    g.num_src = d.num_src;
    for (int i = 0; i < SHGAN_MAX_SRC; ++i) {
        g.src_hi[i] = (const __half*)d.src_hi[i]; g.src_lo[i] = (const __half*)d.src_lo[i];
        g.src_h[i] = d.src_h[i]; g.src_w[i] = d.src_w[i];
    }
    g.N = d.N; g.C = d.C; g.Co = d.Co; g.w_hi = (const __half*)d.w_hi; g.w_lo = (const __half*)d.w_lo;
    g.ntaps = d.ntaps;
    for (int i = 0; i < SHGAN_MAX_TAPS; ++i) {
        g.tap_src[i] = d.tap_src[i]; g.tap_dy[i] = d.tap_dy[i]; g.tap_dx[i] = d.tap_dx[i]; g.tap_w[i] = d.tap_w[i];
    }
    g.OH = d.OH; g.OW = d.OW; g.mode = d.mode; g.z = d.z; g.ZH = d.ZH; g.ZW = d.ZW;
    g.zsy = d.zsy; g.zsx = d.zsx; g.zoy = d.zoy; g.zox = d.zox;
    return g;
}

int launch_conv_simt(const ConvGeom& g, const EpiParams& epi, int block_n, cudaStream_t stream);
int launch_conv_tc(const ConvGeom& g, const EpiParams& epi, int block_n, int passes, cudaStream_t stream);

// torgb partial sums are produced per block of 32 output channels, independent of the GEMM tile width
constexpr int CONV_RGB_BLOCK = 32;

static inline int conv_block_n(int Co, int block_n) {
    if (block_n == 0) block_n = Co >= 256 ? 256 : (Co >= 128 ? 128 : 64);
    return block_n;
}

#ifdef __CUDACC__
// raw-mode store of CH consecutive channels of output pixel (n,y,x) into the strided z tensor
template <int CH>
__device__ __forceinline__ void raw_store(const ConvGeom& g, const float* v, int n, int y, int x, int o0) {
    const long long zi = (((long long)n * g.ZH + (y * g.zsy + g.zoy)) * g.ZW + (x * g.zsx + g.zox)) * g.Co + o0;
#pragma unroll
    for (int i = 0; i < CH; i += 4)
        *reinterpret_cast<float4*>(g.z + zi + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
}
#endif

}  // namespace shgan
