// Halo-tile tcgen05 implementation of shgan_conv_igemm (the product path for the large layers of the generator:
// replaces cuDNN conv / conv_transpose reached from conv2d_resample.py:26-51 and the per-sample weight
// materialisation of stylegan.py:149-190).
//
// Why a second tensor-core kernel: conv_tc.cu fetches one 128-pixel A box PER TAP, i.e. every activation byte crosses
// the L2->SM fabric nine times and every weight tile serves only 128 output pixels.  Measured on B200 that fabric
// saturates at ~10 TB/s (~36 B/clk/SM) and caps conv_tc.cu at 30 % (64 channels) .. 71 % (512 channels) tensor-pipe
// utilisation.  This kernel moves each operand byte once per tile:
//   * one TMA box {64 ch, P = TW+ex, TH+ey} per (64-channel slab, source) stages the output tile's whole input halo in
//     shared memory (SWIZZLE_128B, one pixel = one 128 B row, rows in raster order of the haloed tile);
//   * the A operand of tap (dy,dx) is that same tile read through a UMMA descriptor whose start address is shifted by
//     (dy*P + dx) pixels.  The tensor core swizzles on absolute shared-memory address bits, so a start address that is
//     a multiple of 128 B but not of 1024 B is legal (probed on hardware: tools/desc_probe.cu).  Accumulator row m of
//     M-block j is therefore the "flattened" tile position p = 128 j + m = r*P + c; rows with c >= TW or r >= TH are
//     halo columns / tile overrun, computed and discarded (TW/P of the tensor work is useful);
//   * a tile is 2 M-blocks (256 positions) x BN output channels, so a weight tile [BN x 64] serves twice the rows;
//     weights stream through a ring of [BN x 64] fp16 slots (hi and lo planes are separate slots).
// Per (tap, slab) step the L2->SM traffic drops from 64..96 KB to <= 20 KB (~25 B/clk/SM), under the fabric limit.
//
// Everything else follows conv_tc.cu: fp16 hi/lo split operands with hi*hi + lo*hi + hi*lo passes, two-level
// accumulation (chunks of <= 4 (tap, slab) steps chained in TMEM, drained into fp32 registers with round-to-nearest
// adds by the epilogue warps while the other TMEM buffer accumulates), the fused epilogue of common.cuh, RAW scatter
// mode for the transposed-conv parity passes, persistent CTAs over a static tile schedule.
// Warp roles: warp 0 = weight-slot TMA producer, warp 1 = TMEM allocator + MMA issuer of M-block 0, warp 2 = halo-tile TMA
// producer, warp 3 = MMA issuer of M-block 1, warps 4-11 = epilogue.
#include "conv_common.cuh"
#include "tc_ptx.cuh"

namespace shgan {

// Development-only cycle accounting (compile with -DSHGAN_HALO_PROFILE): per-CTA clock64 totals of the time each role
// spends blocked on each barrier, written to a device buffer that tools/halo_profile.py reads back.
#ifdef SHGAN_HALO_PROFILE
__device__ long long g_halo_prof[148 * 16];
#define HPROF_DECL long long hp_t0 = 0, hp_acc0 = 0, hp_acc1 = 0, hp_acc2 = 0, hp_acc3 = 0; const long long hp_start = clock64();
#define HPROF_BEGIN hp_t0 = clock64();
#define HPROF_END(k) hp_acc##k += clock64() - hp_t0;
#define HPROF_STORE(base) { long long* d = g_halo_prof + blockIdx.x * 16 + (base); d[0] = clock64() - hp_start; d[1] = hp_acc0; d[2] = hp_acc1; d[3] = hp_acc2; }
#else
#define HPROF_DECL
#define HPROF_BEGIN
#define HPROF_END(k)
#define HPROF_STORE(base)
#endif

struct HaloTmaps {
    CUtensorMap a_hi[SHGAN_MAX_SRC];
    CUtensorMap a_lo[SHGAN_MAX_SRC];
    CUtensorMap w_hi;
    CUtensorMap w_lo;
};

struct HaloTile {
    int TW, TH, P;              // output tile TW x TH, haloed row pitch P = TW + ex
    int hx0, hy0;               // halo origin relative to the tile origin (min tap dx / dy)
    int tiles_x, tiles_y, nblk; // nblk = Co / BN
    int total;
    int a_px;                   // pixels (128 B rows) allocated per plane per A buffer, multiple of 8
    int a_box_bytes;            // bytes one TMA box delivers per plane
    int ngroups;                // taps are sorted by source; group g = taps [grp_start[g], grp_start[g+1]) of source grp_src[g]
    int grp_src[SHGAN_MAX_SRC], grp_start[SHGAN_MAX_SRC + 1];
    int tap_off16[SHGAN_MAX_TAPS]; // offset of the tap's A view inside the staged tile, in 16-byte units (descriptor units)
    int tap_wrow[SHGAN_MAX_TAPS];  // tap_w * Co
};

constexpr int HL_THREADS = 384;
constexpr int HL_EPI_THREADS = 256;
constexpr int HL_REGS_LAUNCH = 168, HL_REGS_DEC = 56, HL_REGS_INC = 224;
static_assert(128 * HL_REGS_DEC + 256 * HL_REGS_INC <= 384 * HL_REGS_LAUNCH, "setmaxnreg budget exceeds the CTA register pool");
constexpr int HL_M = 128;            // UMMA M
constexpr int HL_NBLK = 2;           // M-blocks per tile
constexpr int HL_KC = 64;            // channels per slab = one 128 B swizzle row
constexpr int HL_MAX_CHUNK = 4;      // (tap, slab) steps chained in one TMEM accumulator = 48 MMAs (see conv_tc.cu)
constexpr int HL_SMEM_MAX = 232448;  // 227 KB opt-in limit per CTA
constexpr int HL_STG_VECS = CONV_STG_VECS;

template <int BN> struct HaloCfg {
    static constexpr int W_SLOT_BYTES = BN * HL_KC * 2;
    static constexpr int W_SLOTS = BN == 128 ? 4 : 6;
    static constexpr int TMEM_COLS = 512;                   // the whole tensor memory (one CTA per SM)
    static constexpr int NACC = TMEM_COLS / (HL_NBLK * BN); // accumulator buffers of HL_NBLK x BN columns: 2 (BN=128) or 4 (BN=64);
                                                            // the MMA issuer may run NACC chunks ahead of the epilogue warps
    static constexpr int BAR_BYTES = 256;
    static constexpr int STG_BYTES = HL_STG_VECS * BN * 4;  // per-tile epilogue vectors (see stage_epilogue_vectors)
    static constexpr int FIXED_BYTES = W_SLOTS * W_SLOT_BYTES + 1024 /*align slack*/ + BAR_BYTES + STG_BYTES;
    static constexpr int A_PX_MAX = ((HL_SMEM_MAX - FIXED_BYTES) / (4 * 128)) & ~7;   // 2 buffers x 2 planes
};

__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}

template <int BN>
__global__ void __launch_bounds__(HL_THREADS, 1)
conv_halo_kernel(const __grid_constant__ HaloTmaps maps, const ConvGeom g, const EpiParams epi, const __grid_constant__ HaloTile ti,
                 const int passes, const int chunk_iters) {
    using Cfg = HaloCfg<BN>;
    constexpr int W_SLOTS = Cfg::W_SLOTS;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS / STS, not generic LD / ST)
    const int a_plane = ti.a_px * 128;                    // multiple of 1024
    uint8_t* a_base = smem;                               // [buf][plane hi/lo][a_px][128 B]
    uint8_t* w_base = smem + 4 * a_plane;                 // [W_SLOTS][BN][128 B]
    uint64_t* bars = reinterpret_cast<uint64_t*>(w_base + W_SLOTS * Cfg::W_SLOT_BYTES);
    uint64_t* a_full = bars;                  // [2]
    uint64_t* a_empty = bars + 2;             // [2]
    uint64_t* w_full = bars + 4;              // [W_SLOTS]
    uint64_t* w_empty = w_full + W_SLOTS;     // [W_SLOTS]
    constexpr int NACC = Cfg::NACC;
    uint64_t* t_full = w_empty + W_SLOTS;     // [NACC] MMA -> epilogue
    uint64_t* t_empty = t_full + NACC;        // [NACC] epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + NACC);
    float* stg = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + Cfg::BAR_BYTES);   // [HL_STG_VECS][BN]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kslabs = g.C / HL_KC;
    const int kiters = g.ntaps * kslabs;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < g.num_src; ++s) {
            prefetch_tmap(&maps.a_hi[s]);
            prefetch_tmap(&maps.a_lo[s]);
        }
        prefetch_tmap(&maps.w_hi);
        prefetch_tmap(&maps.w_lo);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&a_full[s], 1);
            mbar_init(&a_empty[s], HL_NBLK);     // one commit per MMA-issuing thread
        }
        for (int s = 0; s < NACC; ++s) {
            mbar_init(&t_full[s], HL_NBLK);
            mbar_init(&t_empty[s], HL_EPI_THREADS);
        }
        for (int s = 0; s < W_SLOTS; ++s) {
            mbar_init(&w_full[s], 1);
            mbar_init(&w_empty[s], HL_NBLK);
        }
        mbar_fence_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)Cfg::TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(HL_REGS_DEC));
        if (warp == 2) {
            // ===================== halo-tile TMA producer =====================
            if (elect_one()) {
                int buf = 0;
                uint32_t phase = 0;
                const uint32_t tx_bytes = (passes == 3 ? 2u : 1u) * (uint32_t)ti.a_box_bytes;
                for (int tile = blockIdx.x; tile < ti.total; tile += gridDim.x) {
                    int m = tile / ti.nblk;
                    const int x0 = (m % ti.tiles_x) * ti.TW + ti.hx0;
                    m /= ti.tiles_x;
                    const int y0 = (m % ti.tiles_y) * ti.TH + ti.hy0;
                    const int n0 = m / ti.tiles_y;
                    for (int ks = 0; ks < kslabs; ++ks) {
                        for (int gi = 0; gi < ti.ngroups; ++gi) {
                            const int s = ti.grp_src[gi];
                            mbar_wait(&a_empty[buf], phase ^ 1);
                            uint8_t* sa = a_base + buf * 2 * a_plane;
                            mbar_expect_tx(&a_full[buf], tx_bytes);
                            tma_load_4d(sa, &maps.a_hi[s], &a_full[buf], ks * HL_KC, x0, y0, n0);
                            if (passes == 3) tma_load_4d(sa + a_plane, &maps.a_lo[s], &a_full[buf], ks * HL_KC, x0, y0, n0);
                            buf ^= 1;
                            if (buf == 0) phase ^= 1;
                        }
                    }
                }
            }
        } else if (warp == 0) {
            // ===================== weight-slot TMA producer =====================
            if (elect_one()) {
                int slot = 0;
                uint32_t phase = 0;
                for (int tile = blockIdx.x; tile < ti.total; tile += gridDim.x) {
                    const int nb = tile % ti.nblk;
                    for (int ks = 0; ks < kslabs; ++ks) {
                        for (int t = 0; t < g.ntaps; ++t) {
                            const int wrow = ti.tap_wrow[t] + nb * BN;
                            for (int h = 0; h < (passes == 3 ? 2 : 1); ++h) {
                                mbar_wait(&w_empty[slot], phase ^ 1);
                                mbar_expect_tx(&w_full[slot], Cfg::W_SLOT_BYTES);
                                tma_load_2d(w_base + slot * Cfg::W_SLOT_BYTES, h == 0 ? &maps.w_hi : &maps.w_lo, &w_full[slot],
                                            ks * HL_KC, wrow);
                                if (++slot == W_SLOTS) { slot = 0; phase ^= 1; }
                            }
                        }
                    }
                }
            }
        } else if (warp == 1 || warp == 3) {
            // ===================== MMA issuers: warp 1 owns M-block 0, warp 3 owns M-block 1 =====================
            // Two issuing threads because one thread gets an N <= 128 MMA out only every ~66-80 clocks once the barrier
            // waits and commits are in its loop; two threads add up to 50 (N = 64) / 65 (N = 128, the tensor floor) clocks
            // per MMA (tools/mma_rate_probe.cu).  They wait on the same full barriers; every barrier the MMAs release
            // (w_empty, a_empty, t_full) counts one commit per issuer.
            const int blk = warp == 1 ? 0 : 1;
            // The issue loop is kept lean: the tensor pipe retires an N = 128 MMA every 64 clocks, and (measured) its issue
            // queue is shallow, so ~60 uniform-datapath instructions of descriptor arithmetic after each barrier wait show
            // up as idle tensor cycles.  Descriptors are therefore formed by ADDING 16-byte-unit offsets to low words
            // computed once (the start-address field is the low 14 bits; every operand lives below 256 KB, so the sums
            // cannot carry into the LBO field at bit 16): one add per descriptor.
            if (elect_one()) {
                constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(HL_M >> 4) << 24);
                constexpr uint32_t DESC_HI = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);   // SBO 1024 B, version 1, SWIZZLE_128B
                const uint32_t a_lo0 = ((smem_u32(a_base) & 0x3FFFF) >> 4) | (1u << 16);      // buffer 0, hi plane, pixel 0
                const uint32_t w_lo0 = ((smem_u32(w_base) & 0x3FFFF) >> 4) | (1u << 16);      // slot 0
                const uint32_t a_plane16 = (uint32_t)a_plane >> 4;
                constexpr uint32_t W_SLOT16 = Cfg::W_SLOT_BYTES >> 4, BLK16 = (HL_M * 128) >> 4, K16 = 32 >> 4;
                int buf = 0, slot = 0, acc = 0;
                uint32_t a_phase = 0, w_phase = 0, acc_phase = 0;
                HPROF_DECL
                for (int tile = blockIdx.x; tile < ti.total; tile += gridDim.x) {
                    int step = 0;   // (tap, slab) steps of this tile, chunked into TMEM accumulation chains
                    int in_chunk = 0;
                    for (int ks = 0; ks < kslabs; ++ks) {
                        for (int gi = 0; gi < ti.ngroups; ++gi) {
                            HPROF_BEGIN
                            mbar_wait(&a_full[buf], a_phase);
                            HPROF_END(0)
                            tc_fence_after();
                            const uint32_t ah = a_lo0 + (uint32_t)buf * 2u * a_plane16, al = ah + a_plane16;
                            for (int t = ti.grp_start[gi]; t < ti.grp_start[gi + 1]; ++t) {
                                if (in_chunk == 0) {
                                    HPROF_BEGIN
                                    mbar_wait(&t_empty[acc], acc_phase ^ 1);
                                    HPROF_END(2)
                                    tc_fence_after();
                                }
                                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * HL_NBLK * BN);
                                const uint32_t th = ah + (uint32_t)ti.tap_off16[t], tl = al + (uint32_t)ti.tap_off16[t];
                                // weights hi: A_hi*W_hi and A_lo*W_hi
                                HPROF_BEGIN
                                mbar_wait(&w_full[slot], w_phase);
                                HPROF_END(1)
                                tc_fence_after();
                                uint32_t wb = w_lo0 + (uint32_t)slot * W_SLOT16;
#pragma unroll
                                for (int k = 0; k < HL_KC / 16; ++k) {
                                    const uint64_t db = ((uint64_t)DESC_HI << 32) | (wb + k * K16);
                                    umma_f16(d_tmem + blk * BN, ((uint64_t)DESC_HI << 32) | (th + blk * BLK16 + k * K16), db, idesc,
                                             (in_chunk | k) != 0);
                                    if (passes == 3)
                                        umma_f16(d_tmem + blk * BN, ((uint64_t)DESC_HI << 32) | (tl + blk * BLK16 + k * K16), db, idesc, 1);
                                }
                                umma_commit(&w_empty[slot]);
                                if (++slot == W_SLOTS) { slot = 0; w_phase ^= 1; }
                                if (passes == 3) {
                                    // weights lo: A_hi*W_lo
                                    HPROF_BEGIN
                                    mbar_wait(&w_full[slot], w_phase);
                                    HPROF_END(1)
                                    tc_fence_after();
                                    wb = w_lo0 + (uint32_t)slot * W_SLOT16;
#pragma unroll
                                    for (int k = 0; k < HL_KC / 16; ++k)
                                        umma_f16(d_tmem + blk * BN, ((uint64_t)DESC_HI << 32) | (th + blk * BLK16 + k * K16),
                                                 ((uint64_t)DESC_HI << 32) | (wb + k * K16), idesc, 1);
                                    umma_commit(&w_empty[slot]);
                                    if (++slot == W_SLOTS) { slot = 0; w_phase ^= 1; }
                                }
                                ++step;
                                if (++in_chunk == chunk_iters || step == kiters) {
                                    umma_commit(&t_full[acc]);   // chunk accumulators complete -> epilogue warps
                                    in_chunk = 0;
                                    if (++acc == NACC) { acc = 0; acc_phase ^= 1; }
                                }
                            }
                            umma_commit(&a_empty[buf]);          // frees the halo tile once these MMAs have read it
                            buf ^= 1;
                            if (buf == 0) a_phase ^= 1;
                        }
                    }
                }
                if (blk == 0) { HPROF_STORE(0) }
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(HL_REGS_INC));
        // ===================== epilogue (warps 4..11) =====================
        constexpr int HN = BN / 2;              // columns owned by this thread (per M-block)
        const int q = warp & 3;                 // TMEM lane quarter this warp may access
        const int half = (warp - 4) >> 2;       // column half of the tile
        const int row = q * 32 + lane;          // accumulator row inside an M-block
        int pr[HL_NBLK], pc[HL_NBLK];           // flattened tile position of (block, row) -> (tile row, tile column)
#pragma unroll
        for (int blk = 0; blk < HL_NBLK; ++blk) {
            const int p = blk * HL_M + row;
            pr[blk] = p / ti.P;
            pc[blk] = p - pr[blk] * ti.P;
        }
        const int nchunks = (kiters + chunk_iters - 1) / chunk_iters;
        int acc = 0;
        uint32_t acc_phase = 0;
        HPROF_DECL
        for (int tile = blockIdx.x; tile < ti.total; tile += gridDim.x) {
            int m = tile / ti.nblk;
            const int nb = tile - m * ti.nblk;
            const int x0 = (m % ti.tiles_x) * ti.TW;
            m /= ti.tiles_x;
            const int y0 = (m % ti.tiles_y) * ti.TH;
            const int n = m / ti.tiles_y;

            // stage this tile's epilogue vectors (named barrier 1 = the 256 epilogue threads: the previous tile's final
            // epilogue has finished reading the staging area / the new vectors are visible) and fetch the noise values
            float nz[HL_NBLK];
            bool valid[HL_NBLK];
#pragma unroll
            for (int blk = 0; blk < HL_NBLK; ++blk) {
                const int y = y0 + pr[blk], x = x0 + pc[blk];
                valid[blk] = pc[blk] < ti.TW && pr[blk] < ti.TH && y < g.OH && x < g.OW;
                nz[blk] = 0.f;
            }
            if (g.mode == 0) {
                asm volatile("bar.sync 1, %0;" ::"n"(HL_EPI_THREADS) : "memory");
                stage_epilogue_vectors<BN, HL_EPI_THREADS>(epi, stg, n, g.Co, nb * BN, (int)threadIdx.x - (HL_THREADS - HL_EPI_THREADS));
                asm volatile("bar.sync 1, %0;" ::"n"(HL_EPI_THREADS) : "memory");
                if (epi.noise) {
                    const float ns = __ldg(epi.noise_strength);
#pragma unroll
                    for (int blk = 0; blk < HL_NBLK; ++blk)
                        if (valid[blk])
                            nz[blk] = __ldg(epi.noise + (long long)n * epi.noise_sn + (long long)(y0 + pr[blk]) * g.OW + x0 + pc[blk]) * ns;
                }
            }

            float accv[HL_NBLK][HN];
            for (int c = 0; c < nchunks; ++c) {
                HPROF_BEGIN
                mbar_wait(&t_full[acc], acc_phase);
                HPROF_END(0)
                HPROF_BEGIN
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * HL_NBLK * BN + half * HN);
                // compensation of the truncating accumulate for this chunk's chain of MMAs (ConvGeom::acc_comp)
                const int n_it = kiters - c * chunk_iters < chunk_iters ? kiters - c * chunk_iters : chunk_iters;
                const float comp = 1.f + g.acc_comp * (float)(n_it * (passes == 3 ? 12 : 4));
#pragma unroll
                for (int blk = 0; blk < HL_NBLK; ++blk) {
#pragma unroll
                    for (int p = 0; p < HN / 16; ++p) {
                        float v[16];
                        tmem_ld16(taddr + blk * BN + p * 16, v);
#pragma unroll
                        for (int i = 0; i < 16; ++i) accv[blk][p * 16 + i] = c == 0 ? v[i] * comp : fmaf(v[i], comp, accv[blk][p * 16 + i]);
                    }
                }
                tc_fence_before();
                mbar_arrive(&t_empty[acc]);
                HPROF_END(1)
                if (++acc == NACC) { acc = 0; acc_phase ^= 1; }
            }
            HPROF_BEGIN
#pragma unroll
            for (int blk = 0; blk < HL_NBLK; ++blk) {
                if (valid[blk]) {
                    const int y = y0 + pr[blk], x = x0 + pc[blk];
                    const long long pix = ((long long)n * g.OH + y) * g.OW + x;
                    float rgb[3] = {0.f, 0.f, 0.f};
#pragma unroll
                    for (int p = 0; p < HN / 16; ++p) {
                        const int oi = half * HN + p * 16;
                        const int o0 = nb * BN + oi;
                        if (g.mode == 1) raw_store<16>(g, accv[blk] + p * 16, n, y, x, o0);
                        else epilogue_apply_staged<BN, 16>(epi, stg, accv[blk] + p * 16, nz[blk], g.Co, o0, oi, rgb, pix);
                        if ((p & 1) && g.mode == 0 && epi.rgb_w) {   // one torgb partial per CONV_RGB_BLOCK = 32 channels
                            float* dst = epi.rgb_out + (pix * (g.Co / CONV_RGB_BLOCK) + (o0 - 16) / CONV_RGB_BLOCK) * 4;
                            *reinterpret_cast<float4*>(dst) = make_float4(rgb[0], rgb[1], rgb[2], 0.f);
                            rgb[0] = rgb[1] = rgb[2] = 0.f;
                        }
                    }
                }
            }
            HPROF_END(2)
        }
#ifdef SHGAN_HALO_PROFILE
        if (threadIdx.x == HL_THREADS - HL_EPI_THREADS) HPROF_STORE(4)
        if (threadIdx.x == HL_THREADS - 1) HPROF_STORE(8)
#endif
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS)
                     : "memory");
    }
}

// ---- host side -------------------------------------------------------------------------------
static int encode_map_f16(CUtensorMap* map, const void* ptr, int rank, const uint64_t* dims, const uint32_t* box) {
    return encode_tmap(map, ptr, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, rank, dims, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

// Tile picker: TW x TH output pixels with P*TH <= 256 flattened positions (P = TW + ex), maximising the fraction of
// tensor-core rows that land on real output pixels (halo columns, the P*TH < 256 remainder and partial edge tiles are
// waste), within the shared-memory budget of the two halo buffers.
static bool pick_tile(int OW, int OH, int ex, int ey, int a_px_max, HaloTile& ti, double& eff) {
    const int M = HL_M * HL_NBLK;
    double best = 0.0;
    long long best_bytes = 0;
    for (int tw = 1; tw <= OW && tw + ex <= 256; ++tw) {
        const int P = tw + ex;
        int th = M / P;
        if (th < 1) break;
        if (th > OH) th = OH;
        if (th + ey > 256) th = 256 - ey;
        const int need = (M + ey * P + ex + 7) & ~7;
        const int box = (P * (th + ey) + 7) & ~7;
        const int a_px = need > box ? need : box;
        if (a_px > a_px_max) continue;
        const double e = (double)OW * OH / ((double)ceil_div(OW, tw) * ceil_div(OH, th) * M);
        const long long bytes = (long long)ceil_div(OW, tw) * ceil_div(OH, th) * P * (th + ey);
        if (e > best + 1e-9 || (e > best - 1e-9 && bytes < best_bytes)) {
            best = e; best_bytes = bytes;
            ti.TW = tw; ti.TH = th; ti.P = P; ti.a_px = a_px;
        }
    }
    eff = best;
    return best > 0.0;
}

// Fraction of the issued tensor-core rows that are real output pixels if this layer ran on the halo kernel
// (0 when the layer cannot).  launch_conv_tc uses it to choose between the two tensor-core kernels.
double conv_halo_efficiency(const ConvGeom& g) {
    int hx0 = 1 << 30, hx1 = -(1 << 30), hy0 = 1 << 30, hy1 = -(1 << 30);
    for (int t = 0; t < g.ntaps; ++t) {
        hx0 = g.tap_dx[t] < hx0 ? g.tap_dx[t] : hx0; hx1 = g.tap_dx[t] > hx1 ? g.tap_dx[t] : hx1;
        hy0 = g.tap_dy[t] < hy0 ? g.tap_dy[t] : hy0; hy1 = g.tap_dy[t] > hy1 ? g.tap_dy[t] : hy1;
    }
    HaloTile ti;
    double eff = 0.0;
    const int bn = g.Co % 128 == 0 ? 128 : 64;
    const int a_px_max = bn == 128 ? HaloCfg<128>::A_PX_MAX : HaloCfg<64>::A_PX_MAX;
    if (!pick_tile(g.OW, g.OH, hx1 - hx0, hy1 - hy0, a_px_max, ti, eff)) return 0.0;
    return eff;
}

// Halo vs per-tap kernel for the layers the two-SM kernel does not take (impl == 0), from per-layer timings on B200
// (profiles/r1_conv_findings.md section 4).  Both single-CTA kernels are bound by the rate at which ONE thread can issue
// tcgen05.mma (~70-90 clocks per instruction once barrier waits and descriptor set-up are in the loop), which only N = 256
// tiles hide: the per-tap kernel (BN up to 256) wins wherever the halo waste is large, and the halo kernel wins for the wide
// transposed-convolution passes at 17^2 .. 65^2 (too few tile pairs for the two-SM kernel at 17^2), whose 1-4 tap K loops
// are too short to amortise the per-tap kernel's per-tile operand latency.
bool conv_prefers_halo(const ConvGeom& g) {
    if (g.num_src != 1) return false;
    // 64 -> 64 channel layers at 256^2 and above (the first encoder / last synthesis convolutions): with two issuing threads the
    // halo kernel's N = 64 MMAs go out every ~50 clocks and its operand traffic is a third of the per-tap kernel's
    // (2.04 vs 2.18 ms for the two 512^2 layers at batch 16)
    if (g.mode == 0 && g.C == 64 && g.Co == 64 && g.ntaps == 9 && g.OW >= 256 && g.OH >= 256) return conv_halo_efficiency(g) >= 0.8;
    if (g.mode != 1 || g.C < 512 || g.ntaps < 2) return false;
    if (g.OW < 17 || g.OW > 65 || g.OH < 17 || g.OH > 65) return false;
    return conv_halo_efficiency(g) >= 0.6;
}

template <int BN>
static int launch_halo_bn(const HaloTmaps& maps, const ConvGeom& g, const EpiParams& epi, const HaloTile& ti, int passes,
                          cudaStream_t stream) {
    using Cfg = HaloCfg<BN>;
    static DeviceInit once;
    int num_sms = 0;
    if (int e = device_init(once, &num_sms, []() -> int {
            SHGAN_CUDA(cudaFuncSetAttribute(conv_halo_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, HL_SMEM_MAX));
            return 0;
        })) return e;
    const int smem_bytes = 4 * ti.a_px * 128 + Cfg::FIXED_BYTES;
    SHGAN_CHECK(smem_bytes <= HL_SMEM_MAX, "halo tile does not fit in shared memory");
    const int grid = ti.total < num_sms ? ti.total : num_sms;
    const int kiters = g.ntaps * (g.C / HL_KC);
    const int nchunks = ceil_div(kiters, HL_MAX_CHUNK);
    const int chunk_iters = ceil_div(kiters, nchunks);
    conv_halo_kernel<BN><<<grid, HL_THREADS, smem_bytes, stream>>>(maps, g, epi, ti, passes, chunk_iters);
    SHGAN_LAUNCH_CHECK();
    return 0;
}

int launch_conv_halo(const ConvGeom& g_in, const EpiParams& epi, int block_n, int passes, cudaStream_t stream) {
    SHGAN_CHECK(g_in.C % HL_KC == 0, "C must be a multiple of 64 for the tensor-core path");
    SHGAN_CHECK(passes == 1 || passes == 3, "passes must be 1 or 3");
    if (block_n == 0 || block_n > 128) block_n = g_in.Co % 128 == 0 ? 128 : 64;
    SHGAN_CHECK(g_in.Co % block_n == 0, "Co must be a multiple of block_n");

    // sort the taps by source (stable) so that each source's halo tile is staged once per slab
    ConvGeom g = g_in;
    HaloTile ti;
    int nt = 0;
    ti.ngroups = 0;
    for (int s = 0; s < g_in.num_src; ++s) {
        const int start = nt;
        for (int t = 0; t < g_in.ntaps; ++t)
            if (g_in.tap_src[t] == s) {
                g.tap_src[nt] = s; g.tap_dy[nt] = g_in.tap_dy[t]; g.tap_dx[nt] = g_in.tap_dx[t]; g.tap_w[nt] = g_in.tap_w[t];
                ++nt;
            }
        if (nt > start) {
            ti.grp_src[ti.ngroups] = s;
            ti.grp_start[ti.ngroups] = start;
            ++ti.ngroups;
        }
    }
    ti.grp_start[ti.ngroups] = nt;
    int hx0 = 1 << 30, hx1 = -(1 << 30), hy0 = 1 << 30, hy1 = -(1 << 30);
    for (int t = 0; t < g.ntaps; ++t) {
        hx0 = g.tap_dx[t] < hx0 ? g.tap_dx[t] : hx0; hx1 = g.tap_dx[t] > hx1 ? g.tap_dx[t] : hx1;
        hy0 = g.tap_dy[t] < hy0 ? g.tap_dy[t] : hy0; hy1 = g.tap_dy[t] > hy1 ? g.tap_dy[t] : hy1;
    }
    const int ex = hx1 - hx0, ey = hy1 - hy0;
    double eff = 0.0;
    const int a_px_max = block_n == 128 ? HaloCfg<128>::A_PX_MAX : HaloCfg<64>::A_PX_MAX;
    SHGAN_CHECK(pick_tile(g.OW, g.OH, ex, ey, a_px_max, ti, eff), "no halo tile fits this layer");
    ti.hx0 = hx0; ti.hy0 = hy0;
    ti.tiles_x = ceil_div(g.OW, ti.TW);
    ti.tiles_y = ceil_div(g.OH, ti.TH);
    ti.nblk = g.Co / block_n;
    const long long total = (long long)ti.tiles_x * ti.tiles_y * g.N * ti.nblk;
    SHGAN_CHECK(total <= INT32_MAX, "too many tiles");
    ti.total = (int)total;
    ti.a_box_bytes = ti.P * (ti.TH + ey) * 128;
    for (int t = 0; t < g.ntaps; ++t) {
        ti.tap_off16[t] = ((g.tap_dy[t] - hy0) * ti.P + (g.tap_dx[t] - hx0)) * (128 / 16);
        ti.tap_wrow[t] = g.tap_w[t] * g.Co;
    }

    HaloTmaps maps;
    const uint32_t abox[4] = {(uint32_t)HL_KC, (uint32_t)ti.P, (uint32_t)(ti.TH + ey), 1u};
    for (int s = 0; s < g.num_src; ++s) {
        const uint64_t dims[4] = {(uint64_t)g.C, (uint64_t)g.src_w[s], (uint64_t)g.src_h[s], (uint64_t)g.N};
        if (int e = encode_map_f16(&maps.a_hi[s], g.src_hi[s], 4, dims, abox)) return e;
        if (int e = encode_map_f16(&maps.a_lo[s], g.src_lo[s], 4, dims, abox)) return e;
    }
    int w_taps = 0;
    for (int t = 0; t < g.ntaps; ++t) w_taps = g.tap_w[t] + 1 > w_taps ? g.tap_w[t] + 1 : w_taps;
    const uint64_t wdims[2] = {(uint64_t)g.C, (uint64_t)w_taps * g.Co};
    const uint32_t wbox[2] = {(uint32_t)HL_KC, (uint32_t)block_n};
    if (int e = encode_map_f16(&maps.w_hi, g.w_hi, 2, wdims, wbox)) return e;
    if (int e = encode_map_f16(&maps.w_lo, g.w_lo, 2, wdims, wbox)) return e;

    if (block_n == 64) return launch_halo_bn<64>(maps, g, epi, ti, passes, stream);
    return launch_halo_bn<128>(maps, g, epi, ti, passes, stream);
}

}  // namespace shgan

#ifdef SHGAN_HALO_PROFILE
extern "C" int shgan_debug_halo_profile(long long* host_out /*[148*16]*/) {
    return (int)cudaMemcpyFromSymbol(host_out, shgan::g_halo_prof, sizeof(long long) * 148 * 16);
}
#endif
