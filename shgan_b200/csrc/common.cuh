// Shared host/device helpers for the shgan_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <mutex>
#include <string>

#include "../../include/shgan_b200.h"

namespace shgan {

// ---- error reporting (thread-local message returned by shgan_last_error) ----------------
void set_error(const std::string& msg);
extern std::atomic<uint64_t> g_launch_count;

#define SHGAN_CHECK(cond, msg)                                                          \
    do {                                                                                \
        if (!(cond)) {                                                                  \
            ::shgan::set_error(std::string(__func__) + ": " + (msg));                    \
            return 1;                                                                   \
        }                                                                               \
    } while (0)

#define SHGAN_CUDA(expr)                                                                \
    do {                                                                                \
        cudaError_t _e = (expr);                                                        \
        if (_e != cudaSuccess) {                                                        \
            ::shgan::set_error(std::string(__func__) + ": " #expr " -> " + cudaGetErrorString(_e)); \
            return 2;                                                                   \
        }                                                                               \
    } while (0)

// every kernel launch goes through this: counts the launch and surfaces launch errors
#define SHGAN_LAUNCH_CHECK()                                                            \
    do {                                                                                \
        ::shgan::g_launch_count.fetch_add(1, std::memory_order_relaxed);                \
        SHGAN_CUDA(cudaGetLastError());                                                 \
    } while (0)

// One-time per-DEVICE initialisation of a kernel family (cudaFuncSetAttribute is a per-device setting, and one process may
// drive several GPUs): runs `set_attrs` the first time the calling thread's current device is seen, under a mutex, and
// returns that device's SM count in *num_sms.
constexpr int SHGAN_MAX_DEVICES = 64;
struct DeviceInit {
    std::mutex mu;
    int sms[SHGAN_MAX_DEVICES] = {0};      // 0 = this device has not been initialised yet
};
template <class F>
static inline int device_init(DeviceInit& st, int* num_sms, F&& set_attrs) {
    int dev = 0;
    SHGAN_CUDA(cudaGetDevice(&dev));
    SHGAN_CHECK(dev >= 0 && dev < SHGAN_MAX_DEVICES, "device ordinal out of range");
    std::lock_guard<std::mutex> lock(st.mu);
    if (!st.sms[dev]) {
        if (int e = set_attrs()) return e;
        int n = 0;
        SHGAN_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
        st.sms[dev] = n;
    }
    if (num_sms) *num_sms = st.sms[dev];
    return 0;
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// device-side mirror of shgan_epilogue
struct EpiParams {
    const float* dcoef;
    float wgain;
    const float* noise;
    long long noise_sn;
    const float* noise_strength;
    const float* bias;
    int act;
    float act_alpha, act_gain, act_clamp;
    const __half* skip_hi;
    const __half* skip_lo;
    const float* next_scale;
    const float* rgb_w;
    const float* rgb_style;
    float* rgb_out;
    __half* out_hi;
    __half* out_lo;
    float* out_f32;
};

static inline EpiParams make_epi(const shgan_epilogue& e) {
    EpiParams p;
    p.dcoef = e.dcoef; p.wgain = e.wgain; p.noise = e.noise; p.noise_sn = e.noise_sn;
    p.noise_strength = e.noise_strength; p.bias = e.bias; p.act = e.act; p.act_alpha = e.act_alpha;
    p.act_gain = e.act_gain; p.act_clamp = e.act_clamp;
    p.skip_hi = (const __half*)e.skip_hi; p.skip_lo = (const __half*)e.skip_lo;
    p.next_scale = e.next_scale; p.rgb_w = e.rgb_w; p.rgb_style = e.rgb_style; p.rgb_out = e.rgb_out;
    p.out_hi = (__half*)e.out_hi; p.out_lo = (__half*)e.out_lo; p.out_f32 = e.out_f32;
    return p;
}

// host-side validation shared by every entry point that takes a shgan_epilogue
static inline const char* check_epi(const shgan_epilogue& e, int Co) {
    if (e.noise && !e.noise_strength) return "noise given without noise_strength";
    if ((e.skip_hi == nullptr) != (e.skip_lo == nullptr)) return "skip_hi/skip_lo must both be set";
    if ((e.out_hi == nullptr) != (e.out_lo == nullptr)) return "out_hi/out_lo must both be set";
    if (e.rgb_w && (!e.rgb_style || !e.rgb_out)) return "rgb_w given without rgb_style/rgb_out";
    if (!e.out_hi && !e.out_f32 && !e.rgb_out) return "epilogue has no output";
    if (Co % 8 != 0) return "Co must be a multiple of 8";
    return nullptr;
}

// ---- device helpers ----------------------------------------------------------------------
#ifdef __CUDACC__

// value = hi + lo with hi = fp16(v), lo = fp16(v - hi): 22 significant bits, |err| <= 2^-25 near 0
__device__ __forceinline__ void split_f32(float v, __half& hi, __half& lo) {
    hi = __float2half_rn(v);
    lo = __float2half_rn(v - __half2float(hi));
}

__device__ __forceinline__ uint32_t pack_h2(__half a, __half b) {
    return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}

__device__ __forceinline__ float2 unpack_h2(uint32_t v) {
    return make_float2(__half2float(__ushort_as_half((unsigned short)(v & 0xffff))),
                       __half2float(__ushort_as_half((unsigned short)(v >> 16))));
}

// 8 consecutive channels of a split-plane tensor -> fp32 (hi + lo)
__device__ __forceinline__ void load_planes8(const __half* __restrict__ hi, const __half* __restrict__ lo,
                                             long long idx, float* v) {
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(hi + idx));
    const uint4 b = __ldg(reinterpret_cast<const uint4*>(lo + idx));
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 fa = unpack_h2(aw[i]), fb = unpack_h2(bw[i]);
        v[2 * i] = fa.x + fb.x;
        v[2 * i + 1] = fa.y + fb.y;
    }
}

// two values -> one packed word of each plane.  Packed conversions (F2FP.PACK_AB, two results per instruction) instead of four
// scalar F2F per pair: same round-to-nearest results bit for bit, half the instructions on the conversion pipe (the scalar
// form was a third of fromrgb's issue slots and sits in every convolution / FIR epilogue).
__device__ __forceinline__ void split2_f32(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __half2 h2 = __floats2half2_rn(a, b);
    const float2 f2 = __half22float2(h2);
    const __half2 l2 = __floats2half2_rn(a - f2.x, b - f2.y);
    hi = *reinterpret_cast<const uint32_t*>(&h2);
    lo = *reinterpret_cast<const uint32_t*>(&l2);
}

__device__ __forceinline__ void store_planes8(__half* __restrict__ hi, __half* __restrict__ lo, long long idx,
                                              const float* v) {
    uint32_t hw[4], lw[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) split2_f32(v[2 * i], v[2 * i + 1], hw[i], lw[i]);
    *reinterpret_cast<uint4*>(hi + idx) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    *reinterpret_cast<uint4*>(lo + idx) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
}

// 16 consecutive channels: one 256-bit store per plane (STG.256, sm_100: a lane's 32 bytes are one full sector, so the
// store path handles half the sector writes of two 128-bit stores); idx must be a multiple of 16 elements
__device__ __forceinline__ void store_planes16(__half* __restrict__ hi, __half* __restrict__ lo, long long idx, const float* v) {
    uint32_t hw[8], lw[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) split2_f32(v[2 * i], v[2 * i + 1], hw[i], lw[i]);
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(hi + idx), "r"(hw[0]), "r"(hw[1]), "r"(hw[2]),
                 "r"(hw[3]), "r"(hw[4]), "r"(hw[5]), "r"(hw[6]), "r"(hw[7]) : "memory");
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(lo + idx), "r"(lw[0]), "r"(lw[1]), "r"(lw[2]),
                 "r"(lw[3]), "r"(lw[4]), "r"(lw[5]), "r"(lw[6]), "r"(lw[7]) : "memory");
}

__device__ __forceinline__ float lrelu_agc(float v, float alpha, float gain, float clampv) {
    v = (v >= 0.f ? v : v * alpha) * gain;
    if (clampv > 0.f) v = fminf(fmaxf(v, -clampv), clampv);
    return v;
}

// Pointwise epilogue on CH (multiple of 8) consecutive output channels [o0, o0+CH) of output pixel
// (n, y, x) of an [N,OH,OW,Co] tensor.  v[] holds the raw accumulators on entry.  Semantics are the
// ones documented on `shgan_epilogue` in include/shgan_b200.h.  rgb[3] accumulates the fused torgb
// partial sums (caller zero-initialises and stores them).  out_pix is the pixel index used for the
// plane/fp32 outputs (differs from the natural index when the caller de-interleaves by parity).
template <int CH>
__device__ __forceinline__ void epilogue_apply(const EpiParams& p, float* v, int n, int y, int x, int OH, int OW,
                                               int Co, int o0, float* rgb, long long out_pix) {
    const long long pix = ((long long)n * OH + y) * OW + x;
    const long long no = (long long)n * Co + o0;
    if (p.dcoef) {
#pragma unroll
        for (int i = 0; i < CH; i += 4) {
            const float4 d = __ldg(reinterpret_cast<const float4*>(p.dcoef + no + i));
            v[i] *= d.x * p.wgain; v[i + 1] *= d.y * p.wgain; v[i + 2] *= d.z * p.wgain; v[i + 3] *= d.w * p.wgain;
        }
    } else if (p.wgain != 1.f) {
#pragma unroll
        for (int i = 0; i < CH; ++i) v[i] *= p.wgain;
    }
    if (p.noise) {
        const float nz = __ldg(p.noise + (long long)n * p.noise_sn + (long long)y * OW + x) * __ldg(p.noise_strength);
#pragma unroll
        for (int i = 0; i < CH; ++i) v[i] += nz;
    }
    if (p.bias) {
#pragma unroll
        for (int i = 0; i < CH; i += 4) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + o0 + i));
            v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
        }
    }
    if (p.act) {
#pragma unroll
        for (int i = 0; i < CH; ++i) v[i] = lrelu_agc(v[i], p.act_alpha, p.act_gain, p.act_clamp);
    } else if (p.act_gain != 1.f) {
#pragma unroll
        for (int i = 0; i < CH; ++i) v[i] *= p.act_gain;
    }
    if (p.skip_hi) {
#pragma unroll
        for (int i = 0; i < CH; i += 8) {
            float s[8];
            load_planes8(p.skip_hi, p.skip_lo, pix * Co + o0 + i, s);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[i + j] += s[j];
        }
    }
    if (p.rgb_w) {
#pragma unroll
        for (int i = 0; i < CH; i += 4) {
            const float4 st = __ldg(reinterpret_cast<const float4*>(p.rgb_style + no + i));
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const float4 w = __ldg(reinterpret_cast<const float4*>(p.rgb_w + (long long)j * Co + o0 + i));
                rgb[j] = fmaf(v[i] * st.x, w.x, rgb[j]);
                rgb[j] = fmaf(v[i + 1] * st.y, w.y, rgb[j]);
                rgb[j] = fmaf(v[i + 2] * st.z, w.z, rgb[j]);
                rgb[j] = fmaf(v[i + 3] * st.w, w.w, rgb[j]);
            }
        }
    }
    if (p.out_f32) {
#pragma unroll
        for (int i = 0; i < CH; i += 4)
            *reinterpret_cast<float4*>(p.out_f32 + out_pix * Co + o0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    }
    if (p.out_hi) {
        if (p.next_scale) {
#pragma unroll
            for (int i = 0; i < CH; i += 4) {
                const float4 s = __ldg(reinterpret_cast<const float4*>(p.next_scale + no + i));
                v[i] *= s.x; v[i + 1] *= s.y; v[i + 2] *= s.z; v[i + 3] *= s.w;
            }
        }
        if (CH % 16 == 0) {
#pragma unroll
            for (int i = 0; i < CH; i += 16) store_planes16(p.out_hi, p.out_lo, out_pix * Co + o0 + i, v + i);
        } else {
#pragma unroll
            for (int i = 0; i < CH; i += 8) store_planes8(p.out_hi, p.out_lo, out_pix * Co + o0 + i, v + i);
        }
    }
}

#endif  // __CUDACC__

}  // namespace shgan
