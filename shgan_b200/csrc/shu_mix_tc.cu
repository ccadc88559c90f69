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
// Warp roles (576 threads, one CTA per SM, persistent over (sample, tile); a warp reaches TMEM lanes 32 (warp % 4) ..):
//   warp 0      TMEM allocation, weight image bulk copy (114 KB, once), MMA issue of the second GEMM (one elected lane)
//   warp 17     MMA issue of the first GEMM (one elected lane)
//   warps 1-4   producer: warp = 16 channels (two 16-byte operand chunks), lane = 4 bins; 128-bit loads of the NEXT tile in
//               flight while it waits for the operand tile to be consumed; fp16 hi/lo split into the X operand tile
//   warps 5-8   first epilogue, thread = bin: tcgen05.ld of D1 the moment it completes, bias + ReLU, split into the T operand tile
//   warps 9-16  second epilogue, thread = (bin, 32 of the 64 outputs): tcgen05.ld of each finished pair accumulator (both anchors
//               requested before one wait), cw blend in registers, coalesced fp32 stores of spec2
// What the cycle accounting (MX_TRACE, tools/mix_trace.py) taught, in order of cost: fence.proxy.async waits for every memory
// operation in flight (see the producer); shared-memory accesses through a pointer that went through an integer became generic
// LD / ST; register spills in the draining roles (25 warps at 72 registers) cost more than the extra warps hid; spinning
// mbarrier waits executed half the kernel's instructions (parked waits now); the first GEMM must run a tile ahead.
// TMEM: D1 double-buffered (2 x 64 columns) + a ring of three 128-column pair accumulators = 512 columns.  The issue order
// is G1(tile + 1) -> G2(tile, pair a) -> G2(tile, pair b): the first GEMM runs a tile ahead, so its ReLU epilogue overlaps the
// second GEMM of the tile before and the blend of each pair overlaps the MMAs that follow.
#include "shu_internal.cuh"
#include "tc_ptx.cuh"
#include <stdlib.h>

namespace shgan {

constexpr int MX_THREADS = 576;                 // 18 warps -> 96 registers per thread
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

// cycle accounting (development aid, SHGAN_MIX_TRACE=1 + tools/mix_trace.py): CTA 0 records clock64() at the hand-offs of its first 32 tiles
#define MX_TRACE(role, it_, slot)                                                                             \
    do {                                                                                                      \
        if (trace && blockIdx.x == 0 && (it_) < 32 && (threadIdx.x & 31) == 0) trace[((role) * 32 + (it_)) * 8 + (slot)] = clock64(); \
    } while (0)

// mbarrier wait with a suspend-time hint: the waiting thread is parked by the hardware until the phase completes (or the hint
// expires) instead of spinning through try_wait + branch.  With ~20 of this kernel's 25 warps waiting at any time, spinning
// waits executed half of all instructions of the kernel and starved the working warps of issue slots (measured: 15 cycles per
// instruction in the epilogue math).
__device__ __forceinline__ bool mbar_try_wait_parked(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(1000000u)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_parked(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait_parked(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait_parked(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) {
            printf("shgan: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
            __trap();
        }
    }
}
#define mbar_wait mbar_wait_parked
#define mbar_wait_relaxed mbar_wait_parked

// cw: [6, bins] in the bin order of spec1 / spec2 (the caller passes the kx-major copy when the spectra are kx-major)
__global__ void __launch_bounds__(MX_THREADS, 1)
shu_mix_tc_kernel(const float* __restrict__ spec1, const uint8_t* __restrict__ packed, const float* __restrict__ conv0_b,
                  const float* __restrict__ cw, float* __restrict__ spec2, int N, int bins, float scale,
                  unsigned long long* __restrict__ trace) {
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment by pointer arithmetic on the __shared__ array (an integer round trip would make every access below a
    // generic LD / ST instead of LDS / STS: measured 6 cycles per instruction in the epilogues)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
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
    const int tiles_per_n = (bins + MX_TILE - 1) / MX_TILE;
    const int total = N * tiles_per_n;
    const int first_tile = blockIdx.x, tile_step = gridDim.x;
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
            mbar_init(&d2_empty[i], 256);
        }
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < tiles_per_n; i += MX_THREADS) masks[i] = 0;
    if (threadIdx.x < 64) b_s[threadIdx.x] = __ldg(conv0_b + threadIdx.x) * scale;   // max(acc / scale + b, 0) * scale == max(acc + b * scale, 0)
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
        uint32_t bits = 0;
#pragma unroll
        for (int p = 0; p < 3; ++p)
            if (__ldg(cw + (long long)p * bins + e) != 0.f || __ldg(cw + (long long)(p + 3) * bins + e) != 0.f) bits |= 1u << p;
        if (bits) atomicOr(&masks[e >> 7], bits);
    }
    __syncthreads();

    // instruction descriptors (cute::UMMA::InstrDescriptor): D fp32, A/B fp16, both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
    constexpr uint32_t idesc1 = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    constexpr uint32_t idesc2 = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (warp == 0) {
        // ===================== MMA issuer of the second GEMM =====================
        // Two issuing threads (this one and warp 17 for the first GEMM): a single thread's waits and loop overhead between
        // its batches of 12 MMAs left the tensor pipe idle for half of every tile period (cycle accounting); the MMAs of the
        // two GEMMs touch disjoint TMEM columns and operand tiles, so their order does not matter.
        if (elect_one()) {
            mbar_wait(w_bar, 0);
            int ring = 0;
            uint32_t ring_phase = 0;
            int it = 0;
            int tj = first_tile % tiles_per_n;
            const int dj = tile_step % tiles_per_n;
            for (int tile = first_tile; tile < total; tile += tile_step, ++it) {
                const int buf = it & 1;
                uint32_t mask = masks[tj];
                if (mask == 0) mask = 1;
                MX_TRACE(0, it, 2);
                mbar_wait(&t_full[buf], (uint32_t)((it >> 1) & 1));
                tc_fence_after();
                MX_TRACE(0, it, 3);
                const uint64_t a_hi = umma_desc_sw128(smem_u32(ts) + (uint32_t)(buf * 2 * MX_A_BYTES)), a_lo = a_hi + (MX_A_BYTES >> 4);
                bool first = true;
                for (int p = 0; p < 3; ++p) {
                    if (!(mask & (1u << p))) continue;
                    mbar_wait(&d2_empty[ring], ring_phase ^ 1);
                    tc_fence_after();
                    MX_TRACE(0, it, first ? 4 : 5);
                    first = false;
                    const uint32_t d = tmem_base + (uint32_t)(128 + ring * 128);
                    const uint64_t b_hi = umma_desc_sw128(smem_u32(wp) + (uint32_t)(p * 2 * MX_PAIR_BYTES)), b_lo = b_hi + (MX_PAIR_BYTES >> 4);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {                  // 32 bytes (16 fp16) along K = +2 in the descriptor's address field
                        umma_f16(d, a_hi + 2 * k, b_hi + 2 * k, idesc2, k != 0);
                        umma_f16(d, a_lo + 2 * k, b_hi + 2 * k, idesc2, 1);
                        umma_f16(d, a_hi + 2 * k, b_lo + 2 * k, idesc2, 1);
                    }
                    umma_commit(&d2_full[ring]);
                    if (++ring == 3) { ring = 0; ring_phase ^= 1; }
                }
                umma_commit(&t_empty[buf]);
                MX_TRACE(0, it, 6);
                tj += dj;
                if (tj >= tiles_per_n) tj -= tiles_per_n;
            }
        }
    } else if (warp == 17) {
        // ===================== MMA issuer of the first GEMM (runs ahead of the second: D1 and T are double-buffered) ===========
        if (elect_one()) {
            mbar_wait(w_bar, 0);
            const uint64_t a_hi = umma_desc_sw128(smem_u32(xs)), a_lo = a_hi + (MX_A_BYTES >> 4);
            const uint64_t b_hi = umma_desc_sw128(smem_u32(w0)), b_lo = b_hi + (MX_W0_BYTES >> 4);
            int it = 0;
            for (int tile = first_tile; tile < total; tile += tile_step, ++it) {
                const int buf = it & 1;
                MX_TRACE(0, it, 0);
                mbar_wait(x_full, (uint32_t)(it & 1));
                mbar_wait(&d1_empty[buf], (uint32_t)(((it >> 1) & 1) ^ 1));
                tc_fence_after();
                MX_TRACE(0, it, 1);
                const uint32_t d = tmem_base + (uint32_t)(buf * 64);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    umma_f16(d, a_hi + 2 * k, b_hi + 2 * k, idesc1, k != 0);
                    umma_f16(d, a_lo + 2 * k, b_hi + 2 * k, idesc1, 1);
                    umma_f16(d, a_hi + 2 * k, b_lo + 2 * k, idesc1, 1);
                }
                umma_commit(&d1_full[buf]);
                umma_commit(x_empty);
            }
        }
    } else if (warp < 5) {
        // ===================== producer (4 warps): warp = 16 channels (two 16-byte operand chunks), lane = 4 consecutive bins ====
        // fence.proxy.async compiles to MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC: it waits for EVERY memory operation the thread has in
        // flight, global loads and stores included (measured: a fence after the operand-tile stores with the next tile's loads
        // in flight cost a DRAM round trip per tile).  Hence the role split of this kernel: a warp that fences has nothing else
        // in flight -- here the loads of tile it+1 are issued AFTER the fence of tile it and land during the wait for x_empty;
        // 128-bit loads because the SM's load path holds a limited number of requests (64 scalar loads per thread took 5 k
        // cycles just to issue).
        const int pw = warp - 1;
        float4 xr[16];
        auto load_x = [&](int tile) {
            const int n = tile / tiles_per_n, e = (tile - n * tiles_per_n) * MX_TILE + 4 * lane;
            const float* src = spec1 + ((long long)n * 64 + pw * 16) * bins + e;
            if (e + 3 < bins && (bins & 3) == 0) {
#pragma unroll
                for (int ch = 0; ch < 16; ++ch) xr[ch] = __ldg(reinterpret_cast<const float4*>(src + (long long)ch * bins));
            } else {
#pragma unroll
                for (int ch = 0; ch < 16; ++ch) {
                    const float* sp = src + (long long)ch * bins;
                    xr[ch].x = e < bins ? __ldg(sp) : 0.f;
                    xr[ch].y = e + 1 < bins ? __ldg(sp + 1) : 0.f;
                    xr[ch].z = e + 2 < bins ? __ldg(sp + 2) : 0.f;
                    xr[ch].w = e + 3 < bins ? __ldg(sp + 3) : 0.f;
                }
            }
        };
        // L2 prefetch two tiles ahead (one lane per 128-byte line): the register loads of the next tile then cost an L2 round
        // trip instead of a DRAM one, which is what this role's period is made of (load latency + conversion, in series)
        auto prefetch_x = [&](int tile) {
            const int n = tile / tiles_per_n, e = (tile - n * tiles_per_n) * MX_TILE + 4 * lane;
            if ((lane & 7) == 0 && e < bins) {
                const float* src = spec1 + ((long long)n * 64 + pw * 16) * bins + e;
#pragma unroll
                for (int ch = 0; ch < 16; ++ch) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + (long long)ch * bins));
            }
        };
        int it = 0;
        if (first_tile < total) load_x(first_tile);
        if (first_tile + tile_step < total) prefetch_x(first_tile + tile_step);
        for (int tile = first_tile; tile < total; tile += tile_step, ++it) {
            if (tile + 2 * tile_step < total) prefetch_x(tile + 2 * tile_step);
            MX_TRACE(1, it, 0);
            if (it > 0) mbar_wait(x_empty, (uint32_t)((it - 1) & 1));       // G1 of the previous tile has read the operand tile
            MX_TRACE(1, it, 1);
#pragma unroll
            for (int i = 0; i < 4; ++i) {                                   // bin 4 lane + i = operand row
                const uint32_t row = (uint32_t)(4 * lane + i);
#pragma unroll
                for (int jj = 0; jj < 2; ++jj) {
                    float v[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const float4 t4 = xr[jj * 8 + c];
                        v[c] = (i == 0 ? t4.x : i == 1 ? t4.y : i == 2 ? t4.z : t4.w) * scale;   // the 'forward'-normalised spectrum is tiny: keep it in fp16's normal range
                    }
                    uint4 hi, lo;
                    split8(v, hi, lo);
                    const uint32_t off = row * 128u + (((uint32_t)(pw * 2 + jj) ^ (row & 7u)) << 4);
                    *reinterpret_cast<uint4*>(xs + off) = hi;
                    *reinterpret_cast<uint4*>(xs + MX_A_BYTES + off) = lo;
                }
            }
            MX_TRACE(1, it, 3);
            fence_async_smem();
            MX_TRACE(1, it, 4);
            mbar_arrive(x_full);
            MX_TRACE(1, it, 2);
            if (tile + tile_step < total) load_x(tile + tile_step);
        }
    } else if (warp < 9) {
        // ===================== first epilogue (4 warps): thread = bin.  D1 -> bias + ReLU (shgan.py:319-321) -> operand tile T ====
        const int q = warp & 3, m = q * 32 + lane;
        const uint32_t row_off = (uint32_t)m * 128u, sw = (uint32_t)(m & 7);
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        int it = 0;
        for (int tile = first_tile; tile < total; tile += tile_step, ++it) {
            const int buf = it & 1;
            const uint32_t ph = (uint32_t)((it >> 1) & 1);
            MX_TRACE(3, it, 0);
            mbar_wait(&d1_full[buf], ph);
            mbar_wait(&t_empty[buf], ph ^ 1);
            tc_fence_after();
            MX_TRACE(3, it, 1);
            uint8_t* t_hi = ts + buf * 2 * MX_A_BYTES;
            float va[32], vb[32];
            tmem_ld32_nowait(tmem_base + lane_base + (uint32_t)(buf * 64), va);
            tmem_ld32_nowait(tmem_base + lane_base + (uint32_t)(buf * 64 + 32), vb);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(&d1_empty[buf]);                             // D1 is drained: the first GEMM after next may overwrite it
            MX_TRACE(3, it, 3);
#pragma unroll
            for (int c = 0; c < 2; ++c) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float v[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = fmaxf((c ? vb[j * 8 + i] : va[j * 8 + i]) + b_s[c * 32 + j * 8 + i], 0.f);
                    uint4 hi, lo;
                    split8(v, hi, lo);
                    const uint32_t off = row_off + (((uint32_t)(c * 4 + j) ^ sw) << 4);
                    *reinterpret_cast<uint4*>(t_hi + off) = hi;
                    *reinterpret_cast<uint4*>(t_hi + MX_A_BYTES + off) = lo;
                }
            }
            MX_TRACE(3, it, 5);
            fence_async_smem();
            mbar_arrive(&t_full[buf]);
            MX_TRACE(3, it, 2);
        }
    } else if (warp < 17) {
        // ===================== second epilogue (8 warps): thread = (bin, 32 of the 64 outputs) =====================
        const int q = warp & 3, h = (warp - 9) >> 2, m = q * 32 + lane;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        int ring = 0;
        uint32_t ring_phase = 0;
        int it = 0;
        int n = first_tile / tiles_per_n, tj = first_tile - n * tiles_per_n;       // advanced incrementally: no division per tile
        const int dn = tile_step / tiles_per_n, dj = tile_step - dn * tiles_per_n;
        // blend weights of the NEXT tile are requested while this one is drained (an L2 round trip, and the tile's first
        // arithmetic depends on them)
        float cwn[6];
        auto load_cw = [&](int tjx) {
            const int ex = tjx * MX_TILE + m;
#pragma unroll
            for (int k = 0; k < 6; ++k) cwn[k] = ex < bins ? __ldg(cw + (long long)k * bins + ex) : 0.f;
        };
        if (first_tile < total) load_cw(tj);
        for (int tile = first_tile; tile < total; tile += tile_step, ++it) {
            const int e = tj * MX_TILE + m;
            const bool valid = e < bins;
            uint32_t mask = masks[tj];
            if (mask == 0) mask = 1;
            float cwv[6];
#pragma unroll
            for (int k = 0; k < 6; ++k) cwv[k] = cwn[k] * inv_scale;
            int n2 = n + dn, tj2 = tj + dj;
            if (tj2 >= tiles_per_n) { tj2 -= tiles_per_n; ++n2; }
            if (tile + tile_step < total) load_cw(tj2);
            float acc[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[i] = 0.f;
            bool first = true;
#pragma unroll
            for (int p = 0; p < 3; ++p) {
                if (!(mask & (1u << p))) continue;
                MX_TRACE(2, it, first ? 0 : 3);
                mbar_wait(&d2_full[ring], ring_phase);
                tc_fence_after();
                MX_TRACE(2, it, first ? 1 : 4);
                // columns of the pair accumulator: anchor p outputs 0..63 | anchor 3+p outputs 0..63
                const uint32_t taddr = tmem_base + lane_base + (uint32_t)(128 + ring * 128 + h * 32);
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    float va[16], vb[16];
                    tmem_ld16_nowait(taddr + hh * 16, va);
                    tmem_ld16_nowait(taddr + 64 + hh * 16, vb);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) acc[hh * 16 + i] = fmaf(va[i], cwv[p], fmaf(vb[i], cwv[p + 3], acc[hh * 16 + i]));
                }
                tc_fence_before();
                mbar_arrive(&d2_empty[ring]);
                MX_TRACE(2, it, first ? 2 : 5);
                first = false;
                if (++ring == 3) { ring = 0; ring_phase ^= 1; }
            }
            if (valid) {
                float* dst = spec2 + ((long long)n * 64 + h * 32) * bins + e;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    *dst = acc[i];
                    dst += bins;
                }
            }
            MX_TRACE(2, it, 6);
            n = n2;
            tj = tj2;
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

int launch_shu_mix_tc(const float* spec1, const void* packed, const float* conv0_b, const float* cw_binorder, float* spec2, int N, int R,
                      float scale, cudaStream_t stream, void* trace) {
    static DeviceInit once;
    int num_sms = 148;
    if (int e = device_init(once, &num_sms, []() -> int {
            SHGAN_CUDA(cudaFuncSetAttribute(shu_mix_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MX_SMEM_BYTES));
            return 0;
        })) return e;
    SHGAN_CHECK(((uintptr_t)packed & 15) == 0, "packed SHU weights must be 16-byte aligned");
    const int bins = R * (R / 2 + 1);
    const long long tiles = (long long)N * ((bins + MX_TILE - 1) / MX_TILE);
    SHGAN_CHECK((bins + MX_TILE - 1) / MX_TILE <= MX_MAX_TILES && tiles <= INT32_MAX, "too many tiles");
    const int grid = (int)(tiles < num_sms ? tiles : num_sms);
    shu_mix_tc_kernel<<<grid, MX_THREADS, MX_SMEM_BYTES, stream>>>(spec1, (const uint8_t*)packed, conv0_b, cw_binorder, spec2, N, bins, scale,
                                                                   getenv("SHGAN_MIX_TRACE") ? (unsigned long long*)trace : nullptr);
    SHGAN_LAUNCH_CHECK();
    return 0;
}

}  // namespace shgan
