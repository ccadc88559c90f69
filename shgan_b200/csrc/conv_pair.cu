// Two-SM (cta_group::2) tcgen05 implementation of shgan_conv_igemm for the wide layers (Co a multiple of 256):
// the product path for the 256- and 512-channel convolutions of the generator (replaces cuDNN conv / conv_transpose
// reached from conv2d_resample.py:26-51 and the per-sample weight materialisation of stylegan.py:149-190).
//
// Why: the single-CTA kernel (conv_tc.cu) moves 96 KB of operands per (tap, 64-channel slab) step into each SM for
// 1536 clocks of tensor work (128 pixels x 256 channels, hi/lo split x 3 passes) = 62 B/clk/SM; the L2->SM fabric
// delivers ~45-49 B/clk/SM when all 148 SMs pull (measured, tools/mma_rate_probe.cu), so those layers sit at 60-64 %
// tensor-pipe utilisation.  Here a cluster of two CTAs computes one 256-pixel x 256-channel tile with
// tcgen05.mma.cta_group::2: each SM multiplies ITS 128 pixels by all 256 output channels, but stages only HALF of the
// weight tile (128 of the 256 rows; the tensor cores read the other half from the peer SM's shared memory).  Per SM and
// step that is 32 KB of activations + 32 KB of weights = 42 B/clk, one MMA instruction per 128 clocks issued by a
// single thread for both SMs, and a 3-deep ring of 64 KB stages instead of 2 x 96 KB.
//
// Warp roles per CTA: warp 0 = TMA producer, warp 1 = TMEM allocator + (leader) MMA issuer, warps 4-11 = epilogue (TMEM lane
// quarter warp & 3, column half (warp - 4) >> 2).
// Protocol (barrier arrays live at the same shared-memory offsets in both CTAs; "leader" = cluster rank 0):
//   full[s]   leader only.  Both CTAs' TMA loads complete_tx on the LEADER's barrier (cp.async.bulk.tensor ...
//             .cta_group::2 with the barrier's shared::cluster address in the leader); the leader's producer arrives once
//             with expect_tx = the bytes of both CTAs.
//   empty[s]  one per CTA; tcgen05.commit.cta_group::2 ... multicast::cluster signals both when the MMAs that read stage s
//             have retired.
//   tfull[a]  one per CTA (multicast commit): chunk accumulator a is complete, each CTA's epilogue warps drain their own
//             128 TMEM lanes.
//   tempty[a] leader only, 16 arrivals: one per epilogue warp of either CTA (the peer's arrive remotely).
// Everything else (fp16 hi/lo operands and the hi*hi + lo*hi + hi*lo passes, chunked two-level accumulation, fused
// epilogue, RAW scatter mode, persistent static tile schedule) is as in conv_tc.cu.
#include "conv_common.cuh"
#include "tc_ptx.cuh"

#include <cuda_runtime.h>

namespace shgan {

// Development-only cycle accounting (compile with -DSHGAN_PAIR_PROFILE; tools/pair_profile.py reads it back)
#ifdef SHGAN_PAIR_PROFILE
__device__ long long g_pair_prof[148 * 16];
#define PPROF_DECL long long hp_t0 = 0, hp_acc0 = 0, hp_acc1 = 0, hp_acc2 = 0; const long long hp_start = clock64();
#define PPROF_BEGIN hp_t0 = clock64();
#define PPROF_END(k) hp_acc##k += clock64() - hp_t0;
#define PPROF_STORE(base) { long long* d = g_pair_prof + blockIdx.x * 16 + (base); d[0] = clock64() - hp_start; d[1] = hp_acc0; d[2] = hp_acc1; d[3] = hp_acc2; }
#else
#define PPROF_DECL
#define PPROF_BEGIN
#define PPROF_END(k)
#define PPROF_STORE(base)
#endif

struct PairTmaps {
    CUtensorMap a_hi[SHGAN_MAX_SRC];
    CUtensorMap a_lo[SHGAN_MAX_SRC];
    CUtensorMap w_hi;
    CUtensorMap w_lo;
};

struct PairTile {
    int tw_log2, th_log2;
    int TW, TH, TN;             // M tile of one CTA = TN images x TH rows x TW cols = 128 pixels
    int tiles_x, tiles_y, tiles_n, nblk;
    int m_tiles;                // tiles_x * tiles_y * tiles_n
    int total_pairs;            // ceil(m_tiles / 2) * nblk
};

// 4 non-epilogue warps + 8 epilogue warps (16 epilogue warps with 64 values per thread were measured 3 % slower)
constexpr int CP_THREADS = 384;
constexpr int CP_EPI_THREADS = 256;
constexpr int CP_COL_SPLIT = CP_EPI_THREADS / 128;      // column slices of the tile, one per epilogue warpgroup
// setmaxnreg budget: 384 threads are launched with 168 registers (ptxas cap for that block size): 128*56 + 256*224 <= 384*168
constexpr int CP_REGS_LAUNCH = 168, CP_REGS_DEC = 56, CP_REGS_INC = 224;
static_assert(128 * CP_REGS_DEC + CP_EPI_THREADS * CP_REGS_INC <= CP_THREADS * CP_REGS_LAUNCH, "setmaxnreg budget exceeds the CTA register pool");
constexpr int CP_M = 128;            // pixels per CTA (UMMA M = 256 over the pair)
constexpr int CP_KC = 64;
constexpr int CP_A_BYTES = CP_M * CP_KC * 2;            // 16 KB per plane

// BN = output channels per tile (UMMA N): 256 for the 256/512-channel layers, 128 for the 128-channel ones
template <int BN> struct PairCfg {
    static constexpr int B_BYTES = (BN / 2) * CP_KC * 2;            // this CTA's half of the weight tile, per plane
    static constexpr int STAGE_BYTES = 2 * CP_A_BYTES + 2 * B_BYTES;   // 64 KB (BN = 256) / 48 KB (BN = 128)
    static constexpr int STAGES = BN == 256 ? 3 : 4;
    static constexpr int NACC = 512 / BN;                            // TMEM accumulator buffers
    static constexpr int MAX_CHUNK = BN == 256 ? 6 : 4;              // (tap, slab) steps chained in one TMEM accumulator
    static constexpr int STG_BYTES = CONV_STG_VECS * BN * 4;         // per-tile epilogue vectors (stage_epilogue_vectors)
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256 + STG_BYTES;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_addr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA loads of the 2-CTA form: data lands in THIS CTA's shared memory, the transaction bytes are credited to the barrier at
// `bar_cluster_addr` (the leader's)
__device__ __forceinline__ void tma2_load_4d(void* dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma2_load_2d(void* dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
        : "memory");
}
// D[tmem, both CTAs] (+)= A * B^T over the CTA pair: M = 256 (128 rows per CTA), each CTA supplies N/2 rows of B
__device__ __forceinline__ void umma2_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n"
        "}\n"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
// arrives on the barrier at the same offset in BOTH CTAs once every tcgen05.mma issued so far has completed
__device__ __forceinline__ void umma2_commit_both(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3)
                 : "memory");
}

template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(CP_THREADS, 1)
conv_pair_kernel(const __grid_constant__ PairTmaps maps, const ConvGeom g, const EpiParams epi, const PairTile ti, const int passes,
                 const int chunk_iters) {
    using Cfg = PairCfg<BN>;
    constexpr int CP_BN = BN, CP_STAGES = Cfg::STAGES, CP_NACC = Cfg::NACC, CP_STAGE_BYTES = Cfg::STAGE_BYTES, CP_B_BYTES = Cfg::B_BYTES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS / STS, not generic LD / ST)
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + CP_STAGES * CP_STAGE_BYTES);
    uint64_t* full_bar = bars;                           // [STAGES] (leader's are the live ones)
    uint64_t* empty_bar = bars + CP_STAGES;              // [STAGES]
    uint64_t* tfull_bar = bars + 2 * CP_STAGES;          // [NACC]
    uint64_t* tempty_bar = bars + 2 * CP_STAGES + CP_NACC;   // [NACC] (leader's are the live ones)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * CP_STAGES + 2 * CP_NACC);
    float* stg = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);   // [CONV_STG_VECS][CP_BN]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
    const int kslabs = g.C / CP_KC;
    const int kiters = g.ntaps * kslabs;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < g.num_src; ++s) {
            prefetch_tmap(&maps.a_hi[s]);
            prefetch_tmap(&maps.a_lo[s]);
        }
        prefetch_tmap(&maps.w_hi);
        prefetch_tmap(&maps.w_lo);
        for (int s = 0; s < CP_STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < CP_NACC; ++a) {
            mbar_init(&tfull_bar[a], 1);
            mbar_init(&tempty_bar[a], 2 * (CP_EPI_THREADS / 32));
        }
        mbar_fence_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();        // the peer's barriers are initialised before anything signals them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // tile of this CTA for pair-tile index pt: m = 2 * (pt / nblk) + rank, nb = pt % nblk
    if (warp < 4) {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(CP_REGS_DEC));
      if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t tx_bytes = 2u * (passes == 3 ? 2u : 1u) * (uint32_t)(CP_A_BYTES + CP_B_BYTES);   // both CTAs
            for (int pt = cluster_id; pt < ti.total_pairs; pt += num_clusters) {
                int m = 2 * (pt / ti.nblk) + (int)rank;
                const int nb = pt % ti.nblk;
                const int x0 = (m % ti.tiles_x) * ti.TW;
                m /= ti.tiles_x;
                const int y0 = (m % ti.tiles_y) * ti.TH;
                const int n0 = (m / ti.tiles_y) * ti.TN;      // >= N for the odd tile out: the box is zero-filled
                for (int t = 0; t < g.ntaps; ++t) {
                    const int s = g.tap_src[t];
                    const int cx = x0 + g.tap_dx[t], cy = y0 + g.tap_dy[t];
                    const int wrow = g.tap_w[t] * g.Co + nb * CP_BN + (int)rank * (CP_BN / 2);
                    for (int ks = 0; ks < kslabs; ++ks) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        uint8_t* sa = smem + stage * CP_STAGE_BYTES;
                        const uint32_t fb = mapa_u32(smem_u32(&full_bar[stage]), 0);
                        if (rank == 0) mbar_expect_tx(&full_bar[stage], tx_bytes);
                        tma2_load_4d(sa, &maps.a_hi[s], fb, ks * CP_KC, cx, cy, n0);
                        tma2_load_2d(sa + 2 * CP_A_BYTES, &maps.w_hi, fb, ks * CP_KC, wrow);
                        if (passes == 3) {
                            tma2_load_4d(sa + CP_A_BYTES, &maps.a_lo[s], fb, ks * CP_KC, cx, cy, n0);
                            tma2_load_2d(sa + 2 * CP_A_BYTES + CP_B_BYTES, &maps.w_lo, fb, ks * CP_KC, wrow);
                        }
                        if (++stage == CP_STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
      } else if (warp == 1 && rank == 0) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (elect_one()) {
            // instruction descriptor: D fp32, A/B fp16 K-major, N>>3 in [17,23), M>>4 in [24,29) with M = 256 over the pair
            constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(CP_BN >> 3) << 17) | ((uint32_t)((2 * CP_M) >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            PPROF_DECL
            for (int pt = cluster_id; pt < ti.total_pairs; pt += num_clusters) {
                for (int it0 = 0; it0 < kiters; it0 += chunk_iters) {
                    const int n_it = kiters - it0 < chunk_iters ? kiters - it0 : chunk_iters;
                    PPROF_BEGIN
                    mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                    PPROF_END(1)
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(acc * CP_BN);
                    for (int it = 0; it < n_it; ++it) {
                        PPROF_BEGIN
                        mbar_wait(&full_bar[stage], phase);
                        PPROF_END(0)
                        tc_fence_after();
                        const uint32_t sa = smem_u32(smem + stage * CP_STAGE_BYTES);
                        const uint32_t a_hi = sa, a_lo = sa + CP_A_BYTES, b_hi = sa + 2 * CP_A_BYTES, b_lo = b_hi + CP_B_BYTES;
#pragma unroll
                        for (int k = 0; k < CP_KC / 16; ++k) {
                            const uint32_t ko = k * 32;
                            const uint64_t dah = umma_desc_sw128(a_hi + ko), dbh = umma_desc_sw128(b_hi + ko);
                            umma2_f16(d_tmem, dah, dbh, idesc, (it | k) != 0);
                            if (passes == 3) {
                                umma2_f16(d_tmem, umma_desc_sw128(a_lo + ko), dbh, idesc, 1);
                                umma2_f16(d_tmem, dah, umma_desc_sw128(b_lo + ko), idesc, 1);
                            }
                        }
                        umma2_commit_both(&empty_bar[stage]);
                        if (++stage == CP_STAGES) { stage = 0; phase ^= 1; }
                    }
                    umma2_commit_both(&tfull_bar[acc]);
                    if (++acc == CP_NACC) { acc = 0; acc_phase ^= 1; }
                }
            }
            PPROF_STORE(0)
        }
      }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(CP_REGS_INC));
        // ===================== epilogue (warps 4..11 of both CTAs) =====================
        constexpr int HN = CP_BN / CP_COL_SPLIT;
        const int q = warp & 3;
        const int half = (warp - 4) >> 2;
        const int row = q * 32 + lane;
        const int tx_i = row & (ti.TW - 1);
        const int ty_i = (row >> ti.tw_log2) & (ti.TH - 1);
        const int tn_i = row >> (ti.tw_log2 + ti.th_log2);
        const int nchunks = (kiters + chunk_iters - 1) / chunk_iters;
        int acc = 0;
        uint32_t acc_phase = 0;
        PPROF_DECL
        for (int pt = cluster_id; pt < ti.total_pairs; pt += num_clusters) {
            int m = 2 * (pt / ti.nblk) + (int)rank;
            const int nb = pt % ti.nblk;
            const int x = (m % ti.tiles_x) * ti.TW + tx_i;
            m /= ti.tiles_x;
            const int y = (m % ti.tiles_y) * ti.TH + ty_i;
            const int n = (m / ti.tiles_y) * ti.TN + tn_i;
            const bool valid = n < g.N && y < g.OH && x < g.OW;
            const long long pix = ((long long)n * g.OH + y) * g.OW + x;

            // one image per tile: stage the per-(sample, channel) epilogue vectors and fetch the pixel's noise now, so that the
            // tile's final epilogue (during which the MMA issuer can only run CP_NACC chunks ahead) has no dependent L2 trips
            const bool staged = ti.TN == 1 && g.mode == 0;
            float nz = 0.f;
            if (staged) {
                asm volatile("bar.sync 1, %0;" ::"n"(CP_EPI_THREADS) : "memory");
                stage_epilogue_vectors<CP_BN, CP_EPI_THREADS>(epi, stg, n, g.Co, nb * CP_BN, (int)threadIdx.x - (CP_THREADS - CP_EPI_THREADS));
                asm volatile("bar.sync 1, %0;" ::"n"(CP_EPI_THREADS) : "memory");
                if (epi.noise && valid) nz = __ldg(epi.noise + (long long)n * epi.noise_sn + (long long)y * g.OW + x) * __ldg(epi.noise_strength);
            }

            float accv[HN];
            for (int c = 0; c < nchunks; ++c) {
                PPROF_BEGIN
                mbar_wait(&tfull_bar[acc], acc_phase);
                PPROF_END(0)
                PPROF_BEGIN
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * CP_BN + half * HN);
                // compensation of the truncating accumulate for this chunk's chain of MMAs (ConvGeom::acc_comp)
                const int n_it = kiters - c * chunk_iters < chunk_iters ? kiters - c * chunk_iters : chunk_iters;
                const float comp = 1.f + g.acc_comp * (float)(n_it * (passes == 3 ? 12 : 4));
#pragma unroll
                for (int p = 0; p < HN / 16; ++p) {
                    float v[16];
                    tmem_ld16(taddr + p * 16, v);
#pragma unroll
                    for (int i = 0; i < 16; ++i) accv[p * 16 + i] = c == 0 ? v[i] * comp : fmaf(v[i], comp, accv[p * 16 + i]);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[acc]), 0));
                PPROF_END(1)
                if (++acc == CP_NACC) { acc = 0; acc_phase ^= 1; }
            }
            PPROF_BEGIN
            if (valid) {
                float rgb[3] = {0.f, 0.f, 0.f};
#pragma unroll
                for (int p = 0; p < HN / 16; ++p) {
                    const int oi = half * HN + p * 16;
                    const int o0 = nb * CP_BN + oi;
                    if (g.mode == 1) raw_store<16>(g, accv + p * 16, n, y, x, o0);
                    else if (staged) epilogue_apply_staged<CP_BN, 16>(epi, stg, accv + p * 16, nz, g.Co, o0, oi, rgb, pix);
                    else epilogue_apply<16>(epi, accv + p * 16, n, y, x, g.OH, g.OW, g.Co, o0, rgb, pix);
                    if ((p & 1) && g.mode == 0 && epi.rgb_w) {
                        float* dst = epi.rgb_out + (pix * (g.Co / CONV_RGB_BLOCK) + (o0 - 16) / CONV_RGB_BLOCK) * 4;
                        *reinterpret_cast<float4*>(dst) = make_float4(rgb[0], rgb[1], rgb[2], 0.f);
                        rgb[0] = rgb[1] = rgb[2] = 0.f;
                    }
                }
            }
            PPROF_END(2)
        }
#ifdef SHGAN_PAIR_PROFILE
        if (threadIdx.x == CP_THREADS - CP_EPI_THREADS) PPROF_STORE(4)
#endif
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();        // no CTA leaves (and frees its shared memory / TMEM) while the peer may still signal it
    tc_fence_after();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---- host side -------------------------------------------------------------------------------
static int encode_map_f16(CUtensorMap* map, const void* ptr, int rank, const uint64_t* dims, const uint32_t* box) {
    return encode_tmap(map, ptr, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, rank, dims, box, CU_TENSOR_MAP_SWIZZLE_128B);
}
static int pow2_ceil_i(int v) { int p = 1; while (p < v) p <<= 1; return p; }
static int ilog2_i(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

static int pair_bn(const ConvGeom& g) { return g.Co % 256 == 0 ? 256 : 128; }
bool conv_pair_supported(const ConvGeom& g) { return g.Co % 128 == 0 && g.C % CP_KC == 0; }

// enough 256-pixel x 256-channel tiles for (nearly) every cluster: measured faster than the single-CTA kernels from 32 x 32
// x batch 16 upwards (-15 .. -18 % on the 256- and 512-channel layers), slower below (4 x 4 .. 16 x 16: too few pairs)
bool conv_prefers_pair(const ConvGeom& g) {
    if (!conv_pair_supported(g)) return false;
    const int tw = pow2_ceil_i(g.OW) < 16 ? pow2_ceil_i(g.OW) : 16;
    const int th = pow2_ceil_i(g.OH) < CP_M / tw ? pow2_ceil_i(g.OH) : CP_M / tw;
    const int tn = CP_M / (tw * th);
    const long long m_tiles = (long long)ceil_div(g.OW, tw) * ceil_div(g.OH, th) * ceil_div(g.N, tn);
    return ((m_tiles + 1) / 2) * (g.Co / pair_bn(g)) >= 64;
}

template <int BN>
static int launch_pair_bn(const PairTmaps& maps, const ConvGeom& g, const EpiParams& epi, const PairTile& ti, int passes, cudaStream_t stream) {
    using Cfg = PairCfg<BN>;
    static DeviceInit once;
    int num_sms = 0;
    if (int e = device_init(once, &num_sms, []() -> int {
            SHGAN_CUDA(cudaFuncSetAttribute(conv_pair_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
            return 0;
        })) return e;
    const int max_clusters = num_sms / 2;
    const int clusters = ti.total_pairs < max_clusters ? ti.total_pairs : max_clusters;
    const int kiters = g.ntaps * (g.C / CP_KC);
    const int nchunks = ceil_div(kiters, Cfg::MAX_CHUNK);
    const int chunk_iters = ceil_div(kiters, nchunks);
    conv_pair_kernel<BN><<<2 * clusters, CP_THREADS, Cfg::SMEM_BYTES, stream>>>(maps, g, epi, ti, passes, chunk_iters);
    SHGAN_LAUNCH_CHECK();
    return 0;
}

int launch_conv_pair(const ConvGeom& g, const EpiParams& epi, int passes, cudaStream_t stream) {
    SHGAN_CHECK(conv_pair_supported(g), "the two-SM kernel needs Co % 128 == 0 and C % 64 == 0");
    const int CP_BN = pair_bn(g);
    SHGAN_CHECK(passes == 1 || passes == 3, "passes must be 1 or 3");
    PairTile ti;
    ti.TW = pow2_ceil_i(g.OW) < 16 ? pow2_ceil_i(g.OW) : 16;
    const int th_max = CP_M / ti.TW;
    ti.TH = pow2_ceil_i(g.OH) < th_max ? pow2_ceil_i(g.OH) : th_max;
    ti.TN = CP_M / (ti.TW * ti.TH);
    ti.tw_log2 = ilog2_i(ti.TW);
    ti.th_log2 = ilog2_i(ti.TH);
    ti.tiles_x = ceil_div(g.OW, ti.TW);
    ti.tiles_y = ceil_div(g.OH, ti.TH);
    ti.tiles_n = ceil_div(g.N, ti.TN);
    ti.nblk = g.Co / CP_BN;
    const long long m_tiles = (long long)ti.tiles_x * ti.tiles_y * ti.tiles_n;
    const long long pairs = ((m_tiles + 1) / 2) * ti.nblk;
    SHGAN_CHECK(pairs <= INT32_MAX / 2, "too many tiles");
    ti.m_tiles = (int)m_tiles;
    ti.total_pairs = (int)pairs;

    PairTmaps maps;
    const uint32_t abox[4] = {(uint32_t)CP_KC, (uint32_t)ti.TW, (uint32_t)ti.TH, (uint32_t)ti.TN};
    for (int s = 0; s < g.num_src; ++s) {
        const uint64_t dims[4] = {(uint64_t)g.C, (uint64_t)g.src_w[s], (uint64_t)g.src_h[s], (uint64_t)g.N};
        if (int e = encode_map_f16(&maps.a_hi[s], g.src_hi[s], 4, dims, abox)) return e;
        if (int e = encode_map_f16(&maps.a_lo[s], g.src_lo[s], 4, dims, abox)) return e;
    }
    int w_taps = 0;
    for (int t = 0; t < g.ntaps; ++t) w_taps = g.tap_w[t] + 1 > w_taps ? g.tap_w[t] + 1 : w_taps;
    const uint64_t wdims[2] = {(uint64_t)g.C, (uint64_t)w_taps * g.Co};
    const uint32_t wbox[2] = {(uint32_t)CP_KC, (uint32_t)(CP_BN / 2)};
    if (int e = encode_map_f16(&maps.w_hi, g.w_hi, 2, wdims, wbox)) return e;
    if (int e = encode_map_f16(&maps.w_lo, g.w_lo, 2, wdims, wbox)) return e;

    if (CP_BN == 256) return launch_pair_bn<256>(maps, g, epi, ti, passes, stream);
    return launch_pair_bn<128>(maps, g, epi, ti, passes, stream);
}

}  // namespace shgan

#ifdef SHGAN_PAIR_PROFILE
extern "C" int shgan_debug_pair_profile(long long* host_out /*[148*16]*/) {
    return (int)cudaMemcpyFromSymbol(host_out, shgan::g_pair_prof, sizeof(long long) * 148 * 16);
}
#endif
