// fp32 FMA implementation of shgan_conv_igemm (desc.impl == 1).  NOT the product path: it exists so
// that the tcgen05 kernel (conv_tc.cu) can be cross-checked on the device at full layer sizes, where
// the CPU oracle would take minutes.  It consumes exactly the same operands (split planes, packed
// fp16 hi/lo weights) and runs the same epilogue, so the two kernels must agree to fp32 rounding.
//
// Classic shared-memory tiled SGEMM over the implicit im2col matrix: CTA tile = 64 output pixels
// (flattened n,y,x) x 64 output channels, K step = 16 input channels of one tap; each of the 256
// threads owns 2 pixels x 8 channels.
#include "../conv_common.cuh"

namespace shgan {

constexpr int SM_PIX = 64, SM_CO = 64, SM_K = 16;

__global__ void __launch_bounds__(256)
conv_simt_kernel(ConvGeom g, EpiParams epi, int block_n) {
    __shared__ float As[SM_K][SM_PIX + 4];
    __shared__ float Ws[SM_K][SM_CO + 4];
    const long long npix = (long long)g.N * g.OH * g.OW;
    const long long pix0 = (long long)blockIdx.x * SM_PIX;
    const int co0 = blockIdx.y * SM_CO;
    const int tid = threadIdx.x;

    // loader role: pixel lp (0..63), channel quad lq (0..3)
    const int lp = tid >> 2, lq = tid & 3;
    const long long lpix = pix0 + lp;
    int ln = 0, ly = 0, lx = 0;
    const bool lvalid = lpix < npix;
    if (lvalid) {
        lx = (int)(lpix % g.OW);
        long long t = lpix / g.OW;
        ly = (int)(t % g.OH);
        ln = (int)(t / g.OH);
    }
    // compute role
    const int tx = tid & 7, ty = tid >> 3;
    float acc[2][8];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    for (int t = 0; t < g.ntaps; ++t) {
        const int s = g.tap_src[t];
        const int iy = ly + g.tap_dy[t], ix = lx + g.tap_dx[t];
        const bool in = lvalid && iy >= 0 && iy < g.src_h[s] && ix >= 0 && ix < g.src_w[s];
        const long long abase = in ? (((long long)ln * g.src_h[s] + iy) * g.src_w[s] + ix) * g.C : 0;
        const long long wbase = ((long long)g.tap_w[t] * g.Co + co0 + lp) * g.C;
        for (int c0 = 0; c0 < g.C; c0 += SM_K) {
            float a[4] = {0.f, 0.f, 0.f, 0.f}, w[4];
            if (in) {
                const uint2 h = __ldg(reinterpret_cast<const uint2*>(g.src_hi[s] + abase + c0 + lq * 4));
                const uint2 l = __ldg(reinterpret_cast<const uint2*>(g.src_lo[s] + abase + c0 + lq * 4));
                const float2 h0 = unpack_h2(h.x), h1 = unpack_h2(h.y), l0 = unpack_h2(l.x), l1 = unpack_h2(l.y);
                a[0] = h0.x + l0.x; a[1] = h0.y + l0.y; a[2] = h1.x + l1.x; a[3] = h1.y + l1.y;
            }
            {
                const uint2 h = __ldg(reinterpret_cast<const uint2*>(g.w_hi + wbase + c0 + lq * 4));
                const uint2 l = __ldg(reinterpret_cast<const uint2*>(g.w_lo + wbase + c0 + lq * 4));
                const float2 h0 = unpack_h2(h.x), h1 = unpack_h2(h.y), l0 = unpack_h2(l.x), l1 = unpack_h2(l.y);
                w[0] = h0.x + l0.x; w[1] = h0.y + l0.y; w[2] = h1.x + l1.x; w[3] = h1.y + l1.y;
            }
            __syncthreads();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                As[lq * 4 + i][lp] = a[i];
                Ws[lq * 4 + i][lp] = w[i];
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < SM_K; ++k) {
                const float a0 = As[k][ty * 2], a1 = As[k][ty * 2 + 1];
                const float4 w0 = *reinterpret_cast<const float4*>(&Ws[k][tx * 8]);
                const float4 w1 = *reinterpret_cast<const float4*>(&Ws[k][tx * 8 + 4]);
                const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    acc[0][j] = fmaf(a0, wv[j], acc[0][j]);
                    acc[1][j] = fmaf(a1, wv[j], acc[1][j]);
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const long long pix = pix0 + ty * 2 + i;
        if (pix >= npix) continue;
        const int x = (int)(pix % g.OW);
        long long t = pix / g.OW;
        const int y = (int)(t % g.OH);
        const int n = (int)(t / g.OH);
        const int o0 = co0 + tx * 8;
        if (g.mode == 1) {
            raw_store<8>(g, acc[i], n, y, x, o0);
        } else {
            float rgb[3] = {0.f, 0.f, 0.f};
            epilogue_apply<8>(epi, acc[i], n, y, x, g.OH, g.OW, g.Co, o0, rgb, pix);
            if (epi.rgb_w) {
                float* dst = epi.rgb_out + (pix * (g.Co / CONV_RGB_BLOCK) + o0 / CONV_RGB_BLOCK) * 4;
#pragma unroll
                for (int j = 0; j < 3; ++j) atomicAdd(dst + j, rgb[j]);
            }
        }
    }
}

int launch_conv_simt(const ConvGeom& g, const EpiParams& epi, int block_n, cudaStream_t stream) {
    const long long npix = (long long)g.N * g.OH * g.OW;
    if (epi.rgb_w && g.mode == 0)
        SHGAN_CUDA(cudaMemsetAsync(epi.rgb_out, 0, (size_t)npix * (g.Co / CONV_RGB_BLOCK) * 4 * sizeof(float), stream));
    dim3 grid((unsigned)ceil_div64(npix, SM_PIX), g.Co / SM_CO);
    conv_simt_kernel<<<grid, 256, 0, stream>>>(g, epi, block_n);
    SHGAN_LAUNCH_CHECK();
    return 0;
}

}  // namespace shgan
