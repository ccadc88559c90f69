// Small-channel pointwise kernels at the two ends of the generator (sm_100a):
//   * fromrgb : 1x1 conv Ci(<=8) -> Co with bias + lrelu_agc, reading the NCHW fp32 network input and
//               writing the split-plane NHWC activation (lib/model_zoo/stylegan.py:226-238 for the
//               encoder's fromrgb layer, comodgan.py:44-52).
//   * torgb_combine : img = upsample2d(img_prev) + sum of the torgb partial sums produced by the conv
//               epilogue + bias (comodgan.py:331-338, stylegan.py:325-337), optionally fused with the
//               eval loop's composite + uint8 quantisation (lib/experiments/shgan_default.py:257-262).
// Both are pure HBM streaming kernels; each thread owns one output pixel (x fastest) so the NCHW
// reads/writes are coalesced, and fromrgb's NHWC plane stores are 16 B vectors.
#include "common.cuh"

namespace shgan {

constexpr int FRGB_MAX_CI = 8;
constexpr int FRGB_MAX_CO = 128;

// mask != NULL: fused input preparation of the eval loop (shgan_default.py:269-274): `x` then is the real image [N,Ci-1,HW],
// the layer input is [mask - 0.5, real * mask], and that 4-channel tensor is also written to x_out (the composite at the
// end of the generator reads it) by the thread that owns the pixel's first channel group.
__global__ void __launch_bounds__(256)
fromrgb_kernel(const float* __restrict__ x, const float* __restrict__ mask, float* __restrict__ x_out, const float* __restrict__ w,
               const float* __restrict__ bias, float wgain, float act_alpha, float act_gain, float act_clamp,
               __half* __restrict__ out_hi, __half* __restrict__ out_lo, int N, int Ci, int Co, int HW) {
    __shared__ __align__(16) float s_w[FRGB_MAX_CI * FRGB_MAX_CO];   // transposed [i][o]: a thread's 8 outputs are 2 x LDS.128,
    __shared__ __align__(16) float s_b[FRGB_MAX_CO];                 // conflict-free across the warp's 8 channel groups
    for (int i = threadIdx.x; i < Co * Ci; i += blockDim.x) s_w[(i % Ci) * Co + i / Ci] = w[i] * wgain;
    for (int i = threadIdx.x; i < Co; i += blockDim.x) s_b[i] = bias ? bias[i] : 0.f;
    __syncthreads();
    // 32-bit index arithmetic without divisions in the loop (the host checks N*Co*H*W <= INT32_MAX; 64-bit modulo / division per
    // element were ~half of this kernel's instructions): the channel group of a thread is loop invariant when the grid stride is
    // a multiple of the groups per pixel, and (n, p) advance incrementally
    const unsigned cgs = (unsigned)Co / 8u, total_pix = (unsigned)N * (unsigned)HW;
    const unsigned gstride = gridDim.x * blockDim.x;
    const bool inv = gstride % cgs == 0;
    const unsigned gid0 = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned pix_u = gid0 / cgs, cg_u = gid0 - pix_u * cgs;
    unsigned n_u = pix_u / (unsigned)HW, p_u = pix_u - n_u * (unsigned)HW;
    const unsigned pstride = gstride / cgs;
    for (unsigned gid = gid0; pix_u < total_pix; gid += gstride) {
        const int cg = (int)cg_u, n = (int)n_u, p = (int)p_u;
        const long long pix = (long long)pix_u;    // n*HW + p
        if (inv) {
            pix_u += pstride;
            p_u += pstride;
            while (p_u >= (unsigned)HW) { p_u -= (unsigned)HW; ++n_u; }
        } else {
            const unsigned g2 = gid + gstride;
            pix_u = g2 < gid ? total_pix : g2 / cgs;   // (overflow guard)
            cg_u = g2 - pix_u * cgs;
            n_u = pix_u / (unsigned)HW;
            p_u = pix_u - n_u * (unsigned)HW;
        }
        float xin[FRGB_MAX_CI];
        if (mask) {
            const float m = __ldg(mask + (long long)n * HW + p);
            xin[0] = m - 0.5f;
#pragma unroll
            for (int i = 1; i < FRGB_MAX_CI; ++i) xin[i] = i < Ci ? __ldg(x + ((long long)n * (Ci - 1) + (i - 1)) * HW + p) * m : 0.f;
            if (cg == 0) {
#pragma unroll
                for (int i = 0; i < FRGB_MAX_CI; ++i)
                    if (i < Ci) x_out[((long long)n * Ci + i) * HW + p] = xin[i];
            }
        } else {
#pragma unroll
            for (int i = 0; i < FRGB_MAX_CI; ++i) xin[i] = i < Ci ? __ldg(x + ((long long)n * Ci + i) * HW + p) : 0.f;
        }
        float v[8];
        {
            const float4 b0 = *reinterpret_cast<const float4*>(s_b + cg * 8), b1 = *reinterpret_cast<const float4*>(s_b + cg * 8 + 4);
            v[0] = b0.x; v[1] = b0.y; v[2] = b0.z; v[3] = b0.w; v[4] = b1.x; v[5] = b1.y; v[6] = b1.z; v[7] = b1.w;
        }
#pragma unroll
        for (int i = 0; i < FRGB_MAX_CI; ++i) {
            if (i < Ci) {
                const float4 w0 = *reinterpret_cast<const float4*>(s_w + i * Co + cg * 8);
                const float4 w1 = *reinterpret_cast<const float4*>(s_w + i * Co + cg * 8 + 4);
                v[0] = fmaf(xin[i], w0.x, v[0]); v[1] = fmaf(xin[i], w0.y, v[1]); v[2] = fmaf(xin[i], w0.z, v[2]); v[3] = fmaf(xin[i], w0.w, v[3]);
                v[4] = fmaf(xin[i], w1.x, v[4]); v[5] = fmaf(xin[i], w1.y, v[5]); v[6] = fmaf(xin[i], w1.z, v[6]); v[7] = fmaf(xin[i], w1.w, v[7]);
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = lrelu_agc(v[j], act_alpha, act_gain, act_clamp);
        store_planes8(out_hi, out_lo, pix * Co + cg * 8, v);
    }
}

// Fast instance for the shapes the models use (CI = 3 / 4 input channels known at compile time, grid stride a multiple of the
// channel groups per pixel so that a thread's 8 output channels never change): weights x gain and bias live in registers for
// the whole kernel and the activation gain is folded into them (lrelu(v) * g == lrelu(v * g) for g > 0).  ncu on the generic
// kernel above (profiles/r2_fromrgb_ncu.md): 300 issued instructions per 8 outputs -- every iteration re-read its 4x8 weights
// from shared memory (93 % L1 pipe, short-scoreboard stalls) and ran the MAX_CI = 8 loop with half of it predicated off.
#ifndef FRGB_MINB
#define FRGB_MINB 4
#endif
template <int CI>
__global__ void __launch_bounds__(256, FRGB_MINB)
fromrgb_fast_kernel(const float* __restrict__ x, const float* __restrict__ mask, float* __restrict__ x_out, const float* __restrict__ w,
                    const float* __restrict__ bias, float wgain, float act_alpha, float act_gain, float act_clamp,
                    __half* __restrict__ out_hi, __half* __restrict__ out_lo, int N, int Co, int HW) {
    const unsigned cgs = (unsigned)Co / 8u, total_pix = (unsigned)N * (unsigned)HW;
    const unsigned gstride = gridDim.x * blockDim.x;        // a multiple of cgs (host)
    const unsigned gid0 = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned pix = gid0 / cgs;
    const int cg = (int)(gid0 - pix * cgs);
    const unsigned pstride = gstride / cgs;
    float wr[CI][8], br[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        br[j] = bias ? __ldg(bias + cg * 8 + j) * act_gain : 0.f;
#pragma unroll
        for (int i = 0; i < CI; ++i) wr[i][j] = __ldg(w + (cg * 8 + j) * CI + i) * (wgain * act_gain);
    }
    const float clampv = act_clamp > 0.f ? act_clamp : __int_as_float(0x7f800000);
    unsigned n = pix / (unsigned)HW, p = pix - n * (unsigned)HW;
    // the layer input of pixel (n, p): [mask - 0.5, real * mask] when the input preparation is fused, the NCHW input otherwise
    auto load_in = [&](unsigned nn, unsigned pp, float (&xin)[CI]) {
        if (mask) {
            const float m = __ldg(mask + (size_t)nn * HW + pp);
            xin[0] = m - 0.5f;
#pragma unroll
            for (int i = 1; i < CI; ++i) xin[i] = __ldg(x + ((size_t)nn * (CI - 1) + (i - 1)) * HW + pp) * m;
        } else {
#pragma unroll
            for (int i = 0; i < CI; ++i) xin[i] = __ldg(x + ((size_t)nn * CI + i) * HW + pp);
        }
    };
    // software pipeline: the next pixel's inputs are requested before the current pixel is computed and stored.  One dependent
    // DRAM round trip per iteration held this kernel at 2.7 TB/s of stores (394 us); a plain fill kernel writes 7.4 TB/s on this
    // GPU (tools/hbm_write_probe.py).  Pairs of pixels per iteration measured slower (375 us: 99 registers, two CTAs per SM).
    float cur[CI], nxt[CI];
    if (pix < total_pix) load_in(n, p, cur);
    while (pix < total_pix) {
        const unsigned pix_n = pix + pstride;               // (pixel indices < 2^31 and pstride < 2^28: no wrap-around)
        unsigned n_n = n, p_n = p + pstride;
        while (p_n >= (unsigned)HW) { p_n -= (unsigned)HW; ++n_n; }
#pragma unroll
        for (int i = 0; i < CI; ++i) nxt[i] = 0.f;
        if (pix_n < total_pix) load_in(n_n, p_n, nxt);
        if (mask && cg == 0) {
#pragma unroll
            for (int i = 0; i < CI; ++i) x_out[((size_t)n * CI + i) * HW + p] = cur[i];
        }
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float a = br[j];
#pragma unroll
            for (int i = 0; i < CI; ++i) a = fmaf(cur[i], wr[i][j], a);
            a = fmaxf(a, a * act_alpha);                                   // lrelu (0 <= alpha <= 1), gain already applied
            v[j] = fminf(fmaxf(a, -clampv), clampv);
        }
        store_planes8(out_hi, out_lo, (long long)((size_t)pix * Co + cg * 8), v);
#pragma unroll
        for (int i = 0; i < CI; ++i) cur[i] = nxt[i];
        pix = pix_n; n = n_n; p = p_n;
    }
}

__global__ void __launch_bounds__(256)
torgb_combine_kernel(const float* __restrict__ img_prev, const float* __restrict__ rgb_partial, int n_blocks,
                     const float* __restrict__ bias, const float* __restrict__ f, float* __restrict__ img_out, int N, int H,
                     int W, const float* __restrict__ comp_x, uint8_t* __restrict__ comp_out) {
    __shared__ float s_f[16];
    if (threadIdx.x < 16) s_f[threadIdx.x] = f ? f[15 - threadIdx.x] * 4.f : 0.f;  // flipped filter * gain(up^2)
    __syncthreads();
    const long long total = (long long)N * H * W;
    const int h2 = H / 2, w2 = W / 2;
    for (long long gid = blockIdx.x * (long long)blockDim.x + threadIdx.x; gid < total;
         gid += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(gid % W);
        long long t = gid / W;
        const int y = (int)(t % H);
        const int n = (int)(t / H);
        float v[3] = {0.f, 0.f, 0.f};
        if (img_prev) {
            // upfirdn2d(up=2, pad=[2,1,2,1]): tap (fy,fx) reads zero-inserted position (y+fy-2, x+fx-2), which holds
            // img_prev[(y+fy-2)/2, (x+fx-2)/2] when both are even and in range (upfirdn2d.py:98-138, :279-314)
            const int fy0 = y & 1, fx0 = x & 1;
#pragma unroll
            for (int a = 0; a < 2; ++a) {
                const int fy = fy0 + 2 * a;
                const int Y = y + fy - 2;
                if (Y < 0 || (Y >> 1) >= h2) continue;
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    const int fx = fx0 + 2 * b;
                    const int X = x + fx - 2;
                    if (X < 0 || (X >> 1) >= w2) continue;
                    const float fv = s_f[fy * 4 + fx];
                    const long long ip = (long long)(Y >> 1) * w2 + (X >> 1);
#pragma unroll
                    for (int j = 0; j < 3; ++j)
                        v[j] = fmaf(__ldg(img_prev + ((long long)n * 3 + j) * h2 * w2 + ip), fv, v[j]);
                }
            }
        }
        float r[3] = {0.f, 0.f, 0.f};
        const float4* pp = reinterpret_cast<const float4*>(rgb_partial) + gid * n_blocks;
        for (int b = 0; b < n_blocks; ++b) {
            const float4 q = __ldg(pp + b);
            r[0] += q.x; r[1] += q.y; r[2] += q.z;
        }
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            // reference order: img = upsample(img) + (conv + bias)
            v[j] = __fadd_rn(v[j], __fadd_rn(r[j], bias ? __ldg(bias + j) : 0.f));
            img_out[(((long long)n * 3 + j) * H + y) * W + x] = v[j];
        }
        if (comp_x) {
            const long long hw = (long long)H * W, p = (long long)y * W + x;
            const float m = __fadd_rn(__ldg(comp_x + (long long)n * 4 * hw + p), 0.5f);
            const float om = __fsub_rn(1.f, m);
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const float xin = __ldg(comp_x + ((long long)n * 4 + 1 + j) * hw + p);
                float o = __fadd_rn(__fmul_rn(xin, m), __fmul_rn(v[j], om));
                o = __fadd_rn(__fmul_rn(o, 127.5f), 127.5f);
                o = fminf(fmaxf(o, 0.f), 255.f);
                comp_out[((long long)n * 3 + j) * hw + p] = (uint8_t)o;  // truncation, like .to(torch.uint8)
            }
        }
    }
}

// minibatch_std_layer (stylegan.py:686-705) with num_channels == 1, fused with the channel concat: one CTA per group
// of G samples {g*(N/G) + grp}: stat = mean over (c,y,x) of sqrt(var_over_group + 1e-8); the output planes carry the
// C input channels, the statistic in channel C, zeros up to C_out (channel padding for the tensor-core conv).
__global__ void __launch_bounds__(256)
mbstd_append_kernel(const __half* __restrict__ in_hi, const __half* __restrict__ in_lo, __half* __restrict__ out_hi,
                    __half* __restrict__ out_lo, int N, int G, int HW, int C, int C_out) {
    __shared__ float red[8];
    __shared__ float s_stat;
    const int groups = N / G, grp = blockIdx.x;
    const int E = HW * C;
    float sum = 0.f;
    for (int e = threadIdx.x; e < E; e += blockDim.x) {
        float v[8];
        float m = 0.f;
        for (int g = 0; g < G; ++g) {
            const long long i = (long long)(g * groups + grp) * E + e;
            v[g] = __half2float(in_hi[i]) + __half2float(in_lo[i]);
            m += v[g];
        }
        m /= (float)G;
        float var = 0.f;
        for (int g = 0; g < G; ++g) var += (v[g] - m) * (v[g] - m);
        sum += sqrtf(var / (float)G + 1e-8f);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += red[i];
        s_stat = t / (float)E;
    }
    __syncthreads();
    const float stat = s_stat;
    for (int g = 0; g < G; ++g) {
        const long long n = g * groups + grp;
        for (int e = threadIdx.x; e < HW * C_out; e += blockDim.x) {
            const int p = e / C_out, c = e - p * C_out;
            __half h, l;
            if (c < C) {
                h = in_hi[(n * HW + p) * C + c];
                l = in_lo[(n * HW + p) * C + c];
            } else {
                split_f32(c == C ? stat : 0.f, h, l);
            }
            out_hi[(n * HW + p) * C_out + c] = h;
            out_lo[(n * HW + p) * C_out + c] = l;
        }
    }
}

// eval-loop input preparation (lib/experiments/shgan_default.py:269-274): x = cat([mask - 0.5, real * mask])
__global__ void __launch_bounds__(256)
prepare_input_kernel(const float* __restrict__ real, const float* __restrict__ mask, float* __restrict__ x, long long total4, int hw4) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
        const long long n = i / hw4;
        const int p = (int)(i - n * hw4);
        const float4 m = __ldg(reinterpret_cast<const float4*>(mask) + i);
        float4* xo = reinterpret_cast<float4*>(x) + n * 4 * hw4 + p;
        xo[0] = make_float4(m.x - 0.5f, m.y - 0.5f, m.z - 0.5f, m.w - 0.5f);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const float4 r = __ldg(reinterpret_cast<const float4*>(real) + (n * 3 + j) * hw4 + p);
            xo[(long long)(1 + j) * hw4] = make_float4(r.x * m.x, r.y * m.y, r.z * m.z, r.w * m.w);
        }
    }
}

// float composite of the eval loop (shgan_default.py:257-260) concatenated with the mask channel: the discriminator's input
__global__ void __launch_bounds__(256)
composite_cat_kernel(const float* __restrict__ x, const float* __restrict__ img, float* __restrict__ out, long long total4, int hw4) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
        const long long n = i / hw4;
        const int p = (int)(i - n * hw4);
        const float4* xi = reinterpret_cast<const float4*>(x) + n * 4 * hw4 + p;
        float4* o = reinterpret_cast<float4*>(out) + n * 4 * hw4 + p;
        const float4 c0 = __ldg(xi);
        o[0] = c0;
        const float4 m = make_float4(__fadd_rn(c0.x, 0.5f), __fadd_rn(c0.y, 0.5f), __fadd_rn(c0.z, 0.5f), __fadd_rn(c0.w, 0.5f));
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const float4 a = __ldg(xi + (long long)(1 + j) * hw4);
            const float4 g = __ldg(reinterpret_cast<const float4*>(img) + (n * 3 + j) * hw4 + p);
            // x*m + img*(1-m) with the reference's operation order and no contraction into FMAs
            o[(long long)(1 + j) * hw4] = make_float4(__fadd_rn(__fmul_rn(a.x, m.x), __fmul_rn(g.x, __fsub_rn(1.f, m.x))),
                                                      __fadd_rn(__fmul_rn(a.y, m.y), __fmul_rn(g.y, __fsub_rn(1.f, m.y))),
                                                      __fadd_rn(__fmul_rn(a.z, m.z), __fmul_rn(g.z, __fsub_rn(1.f, m.z))),
                                                      __fadd_rn(__fmul_rn(a.w, m.w), __fmul_rn(g.w, __fsub_rn(1.f, m.w))));
        }
    }
}

}  // namespace shgan

using namespace shgan;

extern "C" int shgan_prepare_input(const float* real, const float* mask, float* x, int N, int H, int W, void* stream) {
    SHGAN_CHECK(real && mask && x, "null pointer");
    SHGAN_CHECK(N >= 0 && H >= 1 && W >= 1 && ((long long)H * W) % 4 == 0, "H*W must be a multiple of 4");
    if (N == 0) return 0;
    const int hw4 = (int)((long long)H * W / 4);
    const long long total4 = (long long)N * hw4;
    long long blocks = ceil_div64(total4, 256);
    if (blocks > 148LL * 16) blocks = 148LL * 16;
    prepare_input_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(real, mask, x, total4, hw4);
    SHGAN_LAUNCH_CHECK();
    return 0;
}

extern "C" int shgan_composite_cat(const float* x, const float* img, float* out, int N, int H, int W, void* stream) {
    SHGAN_CHECK(x && img && out, "null pointer");
    SHGAN_CHECK(N >= 0 && H >= 1 && W >= 1 && ((long long)H * W) % 4 == 0, "H*W must be a multiple of 4");
    if (N == 0) return 0;
    const int hw4 = (int)((long long)H * W / 4);
    const long long total4 = (long long)N * hw4;
    long long blocks = ceil_div64(total4, 256);
    if (blocks > 148LL * 16) blocks = 148LL * 16;
    composite_cat_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, img, out, total4, hw4);
    SHGAN_LAUNCH_CHECK();
    return 0;
}

extern "C" int shgan_mbstd_append(const void* in_hi, const void* in_lo, void* out_hi, void* out_lo, int N, int H, int W, int C,
                                  int C_out, int group_size, void* stream) {
    SHGAN_CHECK(in_hi && in_lo && out_hi && out_lo, "null pointer");
    SHGAN_CHECK(N >= 1 && H >= 1 && W >= 1 && C >= 1 && C_out > C, "bad sizes (C_out must exceed C)");
    const int G = group_size < N ? group_size : N;
    SHGAN_CHECK(G >= 1 && G <= 8 && N % G == 0, "batch must be a multiple of the group size (<= 8)");
    mbstd_append_kernel<<<N / G, 256, 0, (cudaStream_t)stream>>>((const __half*)in_hi, (const __half*)in_lo, (__half*)out_hi,
                                                               (__half*)out_lo, N, G, H * W, C, C_out);
    SHGAN_LAUNCH_CHECK();
    return 0;
}

static int launch_fromrgb(const float* x, const float* mask, float* x_out, const float* w, const float* bias, float wgain, float act_alpha,
                          float act_gain, float act_clamp, void* out_hi, void* out_lo, int N, int Ci, int Co, int HW, void* stream) {
    const long long total = (long long)N * HW * (Co / 8);
    long long blocks = ceil_div64(total, 256);
    if (blocks > 148LL * 32) blocks = 148LL * 32;
    const int cgs = Co / 8;
    // fast instance: 3 / 4 input channels, channel groups dividing the block size (so the grid stride is a multiple of them),
    // the leaky-ReLU form max(v, alpha v) (0 <= alpha <= 1) and a positive gain (folded into the weights)
    const bool fast = (Ci == 3 || Ci == 4) && 256 % cgs == 0 && act_alpha >= 0.f && act_alpha <= 1.f && act_gain > 0.f && (!mask || Ci >= 2);
    if (fast && Ci == 4)
        fromrgb_fast_kernel<4><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, mask, x_out, w, bias, wgain, act_alpha, act_gain, act_clamp,
                                                                                   (__half*)out_hi, (__half*)out_lo, N, Co, HW);
    else if (fast)
        fromrgb_fast_kernel<3><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, mask, x_out, w, bias, wgain, act_alpha, act_gain, act_clamp,
                                                                                   (__half*)out_hi, (__half*)out_lo, N, Co, HW);
    else
        fromrgb_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, mask, x_out, w, bias, wgain, act_alpha, act_gain, act_clamp,
                                                                           (__half*)out_hi, (__half*)out_lo, N, Ci, Co, HW);
    SHGAN_LAUNCH_CHECK();
    return 0;
}

extern "C" int shgan_fromrgb(const float* x, const float* w, const float* bias, float wgain, float act_alpha,
                             float act_gain, float act_clamp, void* out_hi, void* out_lo, int N, int Ci, int Co, int H,
                             int W, void* stream) {
    SHGAN_CHECK(x && w && out_hi && out_lo, "null pointer");
    SHGAN_CHECK(Ci >= 1 && Ci <= FRGB_MAX_CI, "Ci must be in 1..8");
    SHGAN_CHECK(Co >= 8 && Co % 8 == 0 && Co <= FRGB_MAX_CO, "Co must be a multiple of 8, at most 128");
    SHGAN_CHECK(N >= 0 && H >= 1 && W >= 1 && (long long)N * Co * H * W <= INT32_MAX, "bad tensor size");
    if (N == 0) return 0;
    return launch_fromrgb(x, nullptr, nullptr, w, bias, wgain, act_alpha, act_gain, act_clamp, out_hi, out_lo, N, Ci, Co, H * W, stream);
}

extern "C" int shgan_fromrgb_masked(const float* real, const float* mask, float* x_out, const float* w, const float* bias,
                                    float wgain, float act_alpha, float act_gain, float act_clamp, void* out_hi, void* out_lo,
                                    int N, int Ci, int Co, int H, int W, void* stream) {
    SHGAN_CHECK(real && mask && x_out && w && out_hi && out_lo, "null pointer");
    SHGAN_CHECK(Ci >= 2 && Ci <= FRGB_MAX_CI, "Ci (mask channel + image channels) must be in 2..8");
    SHGAN_CHECK(Co >= 8 && Co % 8 == 0 && Co <= FRGB_MAX_CO, "Co must be a multiple of 8, at most 128");
    SHGAN_CHECK(N >= 0 && H >= 1 && W >= 1 && (long long)N * Co * H * W <= INT32_MAX, "bad tensor size");
    if (N == 0) return 0;
    return launch_fromrgb(real, mask, x_out, w, bias, wgain, act_alpha, act_gain, act_clamp, out_hi, out_lo, N, Ci, Co, H * W, stream);
}

extern "C" int shgan_torgb_combine(const float* img_prev, const float* rgb_partial, int n_blocks, const float* bias,
                                   const float* f, float* img_out, int N, int H, int W, const float* comp_x,
                                   uint8_t* comp_out, void* stream) {
    SHGAN_CHECK(rgb_partial && img_out, "null pointer");
    SHGAN_CHECK(n_blocks >= 1, "n_blocks must be at least 1");
    SHGAN_CHECK(!img_prev || f, "upsampling img_prev needs the 4x4 filter");
    SHGAN_CHECK(!img_prev || (H % 2 == 0 && W % 2 == 0), "H and W must be even when img_prev is given");
    SHGAN_CHECK((comp_x == nullptr) == (comp_out == nullptr), "comp_x/comp_out must both be set");
    SHGAN_CHECK(N >= 0 && H >= 1 && W >= 1 && (long long)N * 4 * H * W * n_blocks <= INT32_MAX, "bad tensor size");
    if (N == 0) return 0;
    const long long total = (long long)N * H * W;
    long long blocks = ceil_div64(total, 256);
    if (blocks > 148LL * 32) blocks = 148LL * 32;
    torgb_combine_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(img_prev, rgb_partial, n_blocks, bias, f, img_out,
                                                                             N, H, W, comp_x, comp_out);
    SHGAN_LAUNCH_CHECK();
    return 0;
}
