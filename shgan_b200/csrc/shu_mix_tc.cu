// Channel mix of the Spectral Hint Unit on the 5th-generation tensor cores (sm_100a): per frequency bin
//     t   = ReLU(conv0 . x + b)                          (shgan.py:319-321, 1x1 convolution over the 64 spectrum channels)
//     out = sum_k cw[k, bin] * (W1_k . t)                 (heterogeneous_filter.forward, shgan.py:143-160, six anchor filters)
// as two chained GEMMs per tile of 128 bins, everything between them on chip:
//     D1[128 bins, 64]  = X[128, 64] . conv0^T            tcgen05.mma M = 128, N = 64,  K = 64  (x3 split-precision passes)
//     T = ReLU(D1 + b) -> fp16 hi/lo, written by the epilogue warps straight into the K-major SWIZZLE_128B shared-memory
//                         tile that is the A operand of the second GEMM (never leaves the SM)
//     D2_p[128, 128]    = T[128, 64] . [W1_p ; W1_{3+p}]^T   one N = 128 accumulator per active anchor PAIR p (width node p of
//                         the [2,3] filter: anchors p and 3+p); blended per bin by cw in the draining threads' registers.
// The piecewise-linear blend has at most two active width nodes per bin; tiles are scanned once (kernel prologue) for the
// anchor pairs whose cw is non-zero anywhere in the tile and only those are multiplied (2 of 3 for the reference's cw when
// the spectrum is stored kx-major, which shu_fft64.cu does) -- any cw is handled, all-active tiles just run three pairs.
//
// Warp roles (288 threads, one CTA per SM, persistent over (sample, tile); a warp reaches TMEM lanes 32 (warp % 4) ..):
//   warp 0      TMEM allocation, weight image bulk copy (114 KB, once), MMA issue (one elected lane)
//   warps 1-4   producer + first epilogue, thread = bin: coalesced loads of the 64 channels of its bin (next tile's loads in
//               flight across the current tile's epilogue), fp16 hi/lo split into the X operand tile; tcgen05.ld of D1,
//               bias + ReLU, split into the T operand tile
//   warps 5-8   second epilogue, thread = bin: tcgen05.ld of each finished pair accumulator, cw blend in
//               registers, coalesced fp32 stores of spec2
// TMEM: D1 double-buffered (2 x 64 columns) + a ring of three 128-column pair accumulators = 512 columns.  The issue order
// per tile is G2(pair a) -> G1(next tile) -> G2(pair b): the next tile's ReLU epilogue overlaps the second pair's MMAs and
// the blend of pair a overlaps both, so the tensor pipe does not wait for the CUDA cores in steady state.
#include "shu_internal.cuh"
#include "tc_ptx.cuh"

namespace shgan {

constexpr int MX_THREADS = 288;                 // 9 warps, at most 3 per SM sub-partition -> 168 registers per thread
constexpr int MX_TILE = 128;                      // bins per tile == UMMA M
constexpr int MX_A_BYTES = MX_TILE * 128;         // one fp16 operand plane of a tile: 128 rows x 64 halves
constexpr int MX_W0_BYTES = 64 * 128;             // conv0 plane
constexpr int MX_PAIR_BYTES = 128 * 128;          // one anchor-pair plane
constexpr int MX_MAX_TILES = 1032;                // tiles per sample at input_res 512 (131584 bins), rounded up
constexpr int MX_OFF_X = SHU_PACKED_BYTES;                    // X hi | lo
constexpr int MX_OFF_T = MX_OFF_X + 2 * MX_A_BYTES;           // T[2] (hi | lo)
constexpr int MX_OFF_BIAS = MX_OFF_T + 4 * MX_A_BYTES;        // 64 floats
constexpr int MX_OFF_BARS = MX_OFF_BIAS + 256;                // 17 mbarriers + tmem slot
constexpr int MX_OFF_MASKS = MX_OFF_BARS + 256;               // per-tile active-pair masks
constexpr int MX_SMEM_BYTES = 1024 + MX_OFF_MASKS + MX_MAX_TILES * 4;
static_assert(MX_SMEM_BYTES <= 227 * 1024, "mix kernel shared memory exceeds the sm_100a limit");
// 8 fp32 -> one 16-byte chunk of the hi plane and one of the lo plane
__device__ __forceinline__ void split8(const float* v, uint4& hi, uint4& lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __half2 h2 = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
        const float2 f2 = __half22float2(h2);
        const __half2 l2 = __floats2half2_rn(v[2 * i] - f2.x, v[2 * i + 1] - f2.y);
        h[i] = *reinterpret_cast<const uint32_t*>(&h2);
        l[i] = *reinterpret_cast<const uint32_t*>(&l2);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

__device__ __forceinline__ int cw_index(int e, int R, int log2R, int Rh, int transposed) {
    return transposed ? (e & (R - 1)) * Rh + (e >> log2R) : e;
}

__global__ void __launch_bounds__(MX_THREADS, 1)
shu_mix_tc_kernel(const float* __restrict__ spec1, const uint8_t* __restrict__ packed, const float* __restrict__ conv0_b,
                  const float* __restrict__ cw, float* __restrict__ spec2, int N, int bins, int R, int log2R, int transposed,
                  float scale) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* w0 = smem;                                   // conv0 hi | lo
    uint8_t* wp = smem + 2 * MX_W0_BYTES;                 // pairs: [p][hi | lo]
    uint8_t* xs = smem + MX_OFF_X;
    uint8_t* ts = smem + MX_OFF_T;
    float* b_s = reinterpret_cast<float*>(smem + MX_OFF_BIAS);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + MX_OFF_BARS);
    uint64_t* w_bar = bars;
    uint64_t* x_full = bars + 1;
    uint64_t* x_empty = bars + 2;
    uint64_t* d1_full = bars + 3;      // [2]
    uint64_t* d1_empty = bars + 5;     // [2]
    uint64_t* t_full = bars + 7;       // [2]
    uint64_t* t_empty = bars + 9;      // [2]
    uint64_t* d2_full = bars + 11;     // [3]
    uint64_t* d2_empty = bars + 14;    // [3]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);
    uint32_t* masks = reinterpret_cast<uint32_t*>(smem + MX_OFF_MASKS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Rh = R / 2 + 1;
    const int tiles_per_n = (bins + MX_TILE - 1) / MX_TILE;
    const int total = N * tiles_per_n;
    const float inv_scale = 1.f / scale;

    if (threadIdx.x == 0) {
        mbar_init(w_bar, 1);
        mbar_init(x_full, 128);
        mbar_init(x_empty, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&d1_full[i], 1);
            mbar_init(&d1_empty[i], 128);
            mbar_init(&t_full[i], 128);
            mbar_init(&t_empty[i], 1);
        }
        for (int i = 0; i < 3; ++i) {
            mbar_init(&d2_full[i], 1);
            mbar_init(&d2_empty[i], 128);
        }
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < tiles_per_n; i += MX_THREADS) masks[i] = 0;
    if (threadIdx.x < 64) b_s[threadIdx.x] = __ldg(conv0_b + threadIdx.x);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (warp == 0 && elect_one()) {                       // weight image: seven 16 KB bulk copies on one barrier
        mbar_expect_tx(w_bar, (uint32_t)SHU_PACKED_BYTES);
        for (int i = 0; i < SHU_PACKED_BYTES / 16384; ++i) bulk_g2s(smem + i * 16384, packed + i * 16384, 16384u, w_bar);
    }
    // anchor pairs with a non-zero blend weight anywhere in each tile (same for every sample)
    for (int e = threadIdx.x; e < bins; e += MX_THREADS) {
        const int idx = cw_index(e, R, log2R, Rh, transposed);
        uint32_t bits = 0;
#pragma unroll
        for (int p = 0; p < 3; ++p)
            if (__ldg(cw + (long long)p * bins + idx) != 0.f || __ldg(cw + (long long)(p + 3) * bins + idx) != 0.f) bits |= 1u << p;
        if (bits) atomicOr(&masks[e >> 7], bits);
    }
    __syncthreads();

    if (warp == 0) {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            // instruction descriptors (cute::UMMA::InstrDescriptor): D fp32, A/B fp16, both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
            constexpr uint32_t idesc1 = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            constexpr uint32_t idesc2 = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t x_hi = smem_u32(xs), x_lo = x_hi + MX_A_BYTES;
            const uint32_t w0_hi = smem_u32(w0), w0_lo = w0_hi + MX_W0_BYTES;
            mbar_wait(w_bar, 0);
            auto issue_g1 = [&](int it) {
                const int buf = it & 1;
                mbar_wait(x_full, (uint32_t)(it & 1));
                mbar_wait(&d1_empty[buf], (uint32_t)(((it >> 1) & 1) ^ 1));
                tc_fence_after();
                const uint32_t d = tmem_base + (uint32_t)(buf * 64);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t ko = k * 32;
                    const uint64_t dah = umma_desc_sw128(x_hi + ko), dbh = umma_desc_sw128(w0_hi + ko);
                    umma_f16(d, dah, dbh, idesc1, k != 0);
                    umma_f16(d, umma_desc_sw128(x_lo + ko), dbh, idesc1, 1);
                    umma_f16(d, dah, umma_desc_sw128(w0_lo + ko), idesc1, 1);
                }
                umma_commit(&d1_full[buf]);
                umma_commit(x_empty);
            };
            int ring = 0;
            uint32_t ring_phase = 0;
            int it = 0;
            if ((int)blockIdx.x < total) issue_g1(0);
            for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
                const int buf = it & 1;
                uint32_t mask = masks[tile % tiles_per_n];
                if (mask == 0) mask = 1;
                const bool has_next = tile + (int)gridDim.x < total;
                mbar_wait(&t_full[buf], (uint32_t)((it >> 1) & 1));
                tc_fence_after();
                const uint32_t t_hi = smem_u32(ts) + (uint32_t)(buf * 2 * MX_A_BYTES), t_lo = t_hi + MX_A_BYTES;
                bool first = true;
                for (int p = 0; p < 3; ++p) {
                    if (!(mask & (1u << p))) continue;
                    mbar_wait(&d2_empty[ring], ring_phase ^ 1);
                    tc_fence_after();
                    const uint32_t d = tmem_base + (uint32_t)(128 + ring * 128);
                    const uint32_t b_hi = smem_u32(wp) + (uint32_t)(p * 2 * MX_PAIR_BYTES), b_lo = b_hi + MX_PAIR_BYTES;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t ko = k * 32;
                        const uint64_t dah = umma_desc_sw128(t_hi + ko), dbh = umma_desc_sw128(b_hi + ko);
                        umma_f16(d, dah, dbh, idesc2, k != 0);
                        umma_f16(d, umma_desc_sw128(t_lo + ko), dbh, idesc2, 1);
                        umma_f16(d, dah, umma_desc_sw128(b_lo + ko), idesc2, 1);
                    }
                    umma_commit(&d2_full[ring]);
                    if (++ring == 3) { ring = 0; ring_phase ^= 1; }
                    if (first) {
                        first = false;
                        if (has_next) issue_g1(it + 1);
                    }
                }
                umma_commit(&t_empty[buf]);
            }
        }
    } else if (warp < 5) {
        // ===================== producer + first epilogue: thread = bin =====================
        const int q = warp & 3, m = q * 32 + lane;
        const uint32_t row_off = (uint32_t)m * 128u, sw = (uint32_t)(m & 7);
        float xr[64];
        auto load_x = [&](int tile) {
            const int n = tile / tiles_per_n, e = (tile - n * tiles_per_n) * MX_TILE + m;
            const float* src = spec1 + (long long)n * 64 * bins + e;
            if (e < bins) {
#pragma unroll
                for (int ch = 0; ch < 64; ++ch) xr[ch] = __ldg(src + (long long)ch * bins);
            } else {
#pragma unroll
                for (int ch = 0; ch < 64; ++ch) xr[ch] = 0.f;
            }
        };
        auto write_x = [&]() {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = xr[j * 8 + i] * scale;   // the 'forward'-normalised spectrum is tiny: keep it in fp16's normal range
                uint4 hi, lo;
                split8(v, hi, lo);
                const uint32_t off = row_off + (((uint32_t)j ^ sw) << 4);
                *reinterpret_cast<uint4*>(xs + off) = hi;
                *reinterpret_cast<uint4*>(xs + MX_A_BYTES + off) = lo;
            }
            fence_async_smem();
            mbar_arrive(x_full);
        };
        int it = 0;
        if ((int)blockIdx.x < total) {
            load_x(blockIdx.x);
            write_x();
        }
        for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
            const int buf = it & 1;
            const uint32_t ph = (uint32_t)((it >> 1) & 1);
            const bool has_next = tile + (int)gridDim.x < total;
            if (has_next) load_x(tile + gridDim.x);
            mbar_wait(&d1_full[buf], ph);
            mbar_wait(&t_empty[buf], ph ^ 1);
            tc_fence_after();
            uint8_t* t_hi = ts + buf * 2 * MX_A_BYTES;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 64 + c * 32), v);
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = fmaxf(fmaf(v[i], inv_scale, b_s[c * 32 + i]), 0.f) * scale;   // bias + ReLU, shgan.py:319-321
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint4 hi, lo;
                    split8(v + j * 8, hi, lo);
                    const uint32_t off = row_off + (((uint32_t)(c * 4 + j) ^ sw) << 4);
                    *reinterpret_cast<uint4*>(t_hi + off) = hi;
                    *reinterpret_cast<uint4*>(t_hi + MX_A_BYTES + off) = lo;
                }
            }
            fence_async_smem();
            tc_fence_before();
            mbar_arrive(&d1_empty[buf]);
            mbar_arrive(&t_full[buf]);
            if (has_next) {
                mbar_wait(x_empty, (uint32_t)(it & 1));
                write_x();
            }
        }
    } else {
        // ===================== second epilogue: thread = bin, all 64 outputs =====================
        const int q = warp & 3, m = q * 32 + lane;
        int ring = 0;
        uint32_t ring_phase = 0;
        for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
            const int n = tile / tiles_per_n, tj = tile - n * tiles_per_n, e = tj * MX_TILE + m;
            const bool valid = e < bins;
            uint32_t mask = masks[tj];
            if (mask == 0) mask = 1;
            const int idx = valid ? cw_index(e, R, log2R, Rh, transposed) : 0;
            float acc[64];
#pragma unroll
            for (int i = 0; i < 64; ++i) acc[i] = 0.f;
            for (int p = 0; p < 3; ++p) {
                if (!(mask & (1u << p))) continue;
                const float ca = valid ? __ldg(cw + (long long)p * bins + idx) : 0.f;
                const float cb = valid ? __ldg(cw + (long long)(p + 3) * bins + idx) : 0.f;
                mbar_wait(&d2_full[ring], ring_phase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(128 + ring * 128);
#pragma unroll
                for (int c = 0; c < 4; ++c) {                       // columns: anchor p outputs 0..63, anchor 3+p outputs 0..63
                    float v[32];
                    tmem_ld32(taddr + c * 32, v);
                    const float cf = c < 2 ? ca : cb;
#pragma unroll
                    for (int i = 0; i < 32; ++i) acc[(c & 1) * 32 + i] = fmaf(v[i], cf, acc[(c & 1) * 32 + i]);
                }
                tc_fence_before();
                mbar_arrive(&d2_empty[ring]);
                if (++ring == 3) { ring = 0; ring_phase ^= 1; }
            }
            if (valid) {
                float* dst = spec2 + (long long)n * 64 * bins + e;
#pragma unroll
                for (int i = 0; i < 64; ++i) dst[(long long)i * bins] = acc[i] * inv_scale;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

// one-time packing of the SHU weights into the kernel's shared-memory image (layout: shu_internal.cuh)
__global__ void __launch_bounds__(256)
shu_pack_tc_kernel(const float* __restrict__ conv0_w, const float* __restrict__ df1_w, uint8_t* __restrict__ packed) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    // rows: 64 of conv0, then 3 pairs x 128; one 16-byte chunk (8 input channels) per work item
    for (int i = gid; i < (64 + 3 * 128) * 8; i += stride) {
        const int row = i >> 3, j = i & 7;
        float v[8];
        uint8_t* base;
        int r;
        if (row < 64) {
            r = row;
            base = packed;
#pragma unroll
            for (int t = 0; t < 8; ++t) v[t] = __ldg(conv0_w + r * 64 + j * 8 + t);          // conv0.weight [o, i]
        } else {
            const int pr = row - 64, p = pr >> 7;
            r = pr & 127;
            const int k = r < 64 ? p : 3 + p, o2 = r & 63;
            base = packed + 2 * MX_W0_BYTES + p * 2 * MX_PAIR_BYTES;
#pragma unroll
            for (int t = 0; t < 8; ++t) v[t] = __ldg(df1_w + (j * 8 + t) * 384 + o2 * 6 + k);  // df1.weight [i, o2*6 + k]
        }
        uint4 hi, lo;
        split8(v, hi, lo);
        const uint32_t off = (uint32_t)r * 128u + (((uint32_t)j ^ (uint32_t)(r & 7)) << 4);
        *reinterpret_cast<uint4*>(base + off) = hi;
        *reinterpret_cast<uint4*>(base + (row < 64 ? MX_W0_BYTES : MX_PAIR_BYTES) + off) = lo;
    }
}

int launch_shu_pack_tc(const float* conv0_w, const float* df1_w, void* packed, cudaStream_t stream) {
    shu_pack_tc_kernel<<<16, 256, 0, stream>>>(conv0_w, df1_w, (uint8_t*)packed);
    SHGAN_LAUNCH_CHECK();
    return 0;
}

int launch_shu_mix_tc(const float* spec1, const void* packed, const float* conv0_b, const float* cw, float* spec2, int N, int R,
                      int transposed, float scale, cudaStream_t stream) {
    static DeviceInit once;
    int num_sms = 148;
    if (int e = device_init(once, &num_sms, []() -> int {
            SHGAN_CUDA(cudaFuncSetAttribute(shu_mix_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MX_SMEM_BYTES));
            return 0;
        })) return e;
    SHGAN_CHECK(((uintptr_t)packed & 15) == 0, "packed SHU weights must be 16-byte aligned");
    const int bins = R * (R / 2 + 1);
    int log2R = 0;
    while ((1 << log2R) < R) ++log2R;
    const long long tiles = (long long)N * ((bins + MX_TILE - 1) / MX_TILE);
    SHGAN_CHECK((bins + MX_TILE - 1) / MX_TILE <= MX_MAX_TILES && tiles <= INT32_MAX, "too many tiles");
    const int grid = (int)(tiles < num_sms ? tiles : num_sms);
    shu_mix_tc_kernel<<<grid, MX_THREADS, MX_SMEM_BYTES, stream>>>(spec1, (const uint8_t*)packed, conv0_b, cw, spec2, N, bins, R, log2R,
                                                                   transposed, scale);
    SHGAN_LAUNCH_CHECK();
    return 0;
}

}  // namespace shgan
