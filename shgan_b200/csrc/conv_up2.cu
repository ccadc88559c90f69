// shgan_conv_up2: the up-sampling modulated convolution of the synthesis network in ONE launch --
//   conv_transpose2d(stride 2, 3x3, un-flipped weights) -> 4x4 blur (pad 1, gain 4) -> demod, noise, bias, lrelu,
//   + feats[res] skip, next-layer style, hi/lo split
// (replaces conv2d_resample.py:123-142 + stylegan.py:187-193,298-303 + comodgan.py:319-320; before: four RAW parity
// launches of shgan_conv_igemm that wrote the fp32 (2h+1)^2 intermediate `z` to HBM plus one shgan_fir_nhwc launch that
// read it back).
//
// The transposed convolution runs at its algorithmic cost on the INPUT grid.  z[2i+py, 2j+px] (parity (py,px)) only
// receives the taps ky = py, kx = px (mod 2), read at x[i - (ky>>1), j - (kx>>1)]: 4 + 2 + 2 + 1 = 9 taps over four
// operand shifts (0,0) (0,-1) (-1,0) (-1,-1).  One CTA tile is 128 flattened positions of a haloed input tile (8 rows x
// 16 columns, pitch 16, staged ONCE per 64-channel slab by one TMA box; shifted A views are UMMA descriptor start
// offsets, cf. conv_halo.cu) and its accumulator holds all four parities side by side in TMEM:
//     columns   0.. 63  parity (1,0)      64..127  parity (0,0)      128..191  parity (0,1)      192..255  parity (1,1)
// so that every shift is ONE tcgen05.mma per k-step whose B operand stacks the taps of the parities it feeds along N:
//     shift ( 0, 0): N = 256 at column   0  taps [3,0,1,4]        shift ( 0,-1): N = 128 at column  0  taps [5,2]
//     shift (-1, 0): N = 128 at column  64  taps [6,7]            shift (-1,-1): N =  64 at column 64  tap  [8]
// (the N = 64 half-rate MMA shape carries 1/9 of the work instead of all of it).  Weights are pre-packed in exactly that
// row order (packing.pack_up2_weight).  Precision scheme as everywhere: fp16 hi/lo operands, hi*hi + lo*hi + hi*lo, one
// 64-channel slab (<= 48 chained MMAs per column) per TMEM chunk, chunks summed in fp32 registers with the truncation
// compensation (shgan_conv_desc::acc_comp).
//
// Epilogue: the 8 epilogue warps drain the chunks (thread = tile position x two parity blocks), then, in two rounds of
// 32 channels, scatter their parities into a shared-memory z tile [16 z-rows][2 column parities][16][32 ch] (float4 slots
// XOR-swizzled by the column so that both the position-major writes and the channel-major reads are conflict free),
// and 208 of the 256 threads run the separable 4-tap blur over it (thread = 4 channels x 2 output columns x 6 output
// rows, sliding window in registers) followed by the pointwise epilogue, storing 64 contiguous bytes per pixel and plane.
// A tile owns 12 x 26 outputs and recomputes a 3-row / 3-column z halo (the 15th input column is the wrap-around column
// of the flattened tile): 61 % of the tensor rows are useful, in exchange for never writing z.
//
// Warp roles: warp 0 = weight TMA producer, warp 1 = TMEM allocator + MMA issuer, warp 2 = halo-tile TMA producer,
// warps 4-11 = epilogue.
#include "conv_common.cuh"
#include "tc_ptx.cuh"

namespace shgan {

struct Up2Tmaps {
    CUtensorMap a_hi, a_lo, w_hi, w_lo;
};

struct Up2Params {
    int N, H, W, C, Co;          // input planes [N,H,W,C]; output [N,2H,2W,Co]
    int OH, OW;
    int tiles_x, tiles_y, nblk, total;
    int cluster;                 // 1: launched as clusters of two CTAs that run the same channel block in lock step and share every
                                 // weight load (each loads half of the units and multicasts them to both)
    int total_m;                 // spatial tiles (tiles_x * tiles_y * N)
    float fx[4], fy[4];          // separable blur taps as applied (correlation order); gain folded into fy
    float acc_comp;
    int passes;
    int chunk_slabs;             // 64-channel slabs chained in one TMEM accumulator (1, or 4 for C >= 256: fewer, longer chunks
                                 // let the MMAs run eight slabs ahead while the epilogue warps are busy with the blur)
    // per shift group (issue order), read by the single-thread producer / issuer loops from the constant bank:
    // instruction descriptor, TMEM column offset, A start offset and weight region offset in 16-byte descriptor units,
    // number of 64-row weight units and the first unit
    uint32_t sg_idesc[4], sg_col[4], sg_aoff16[4], sg_wb16[4];
    int sg_units[4], sg_ubase[4];
};

// Two instances.  WIDE = false: 8 epilogue warps with the whole tile's accumulator sums in registers (any number of TMEM chunks per
// tile).  WIDE = true (C <= 256: the tile is ONE chunk, so the accumulator can stay in TMEM until the z tile is free): 16 epilogue
// warps, thread = tile position x ONE parity block for the drain and 4 channels x 1 output column x 6 output rows for the blur --
// the 512^2 / 256^2 layers are bound by the epilogue's instruction issue (profiles/r2_up2_findings.md), and twice the warps at
// half the instructions each halve it.  Register pool of a CTA = threads x the launch allocation (168 / 96):
// 128 x 56 + 256 x 224 = 384 x 168;  128 x 56 + 512 x 104 <= 640 x 96.
template <bool WIDE> struct U2Cfg {
    static constexpr int EPI_THREADS = WIDE ? 512 : 256;
    static constexpr int THREADS = 128 + EPI_THREADS;
    static constexpr int REGS_DEC = 56, REGS_INC = WIDE ? 104 : 224;
};
constexpr int U2_KC = 64;
constexpr int U2_P = 16;                         // tile pitch (input columns per flattened row)
constexpr int U2_ROWS = 8;                       // input rows per tile: U2_ROWS * U2_P = 128 = UMMA M
constexpr int U2_OWN_Y = 12, U2_OWN_X = 26;      // outputs owned by a tile
constexpr int U2_STEP_I = 6, U2_STEP_J = 13;     // tile step on the input grid
constexpr int U2_A_BOX_ROWS = (U2_ROWS + 1) * U2_P;          // 144 positions delivered by the TMA box
constexpr int U2_A_PX = 152;                     // allocated positions per plane (reads reach position 127 + 17)
constexpr int U2_A_PLANE = U2_A_PX * 128;        // 19456 B, a multiple of 1024
constexpr int U2_W_UNIT = 64 * 128;              // one [64 co x 64 ci] fp16 block
constexpr int U2_W_BYTES = 9 * U2_W_UNIT;        // one plane of one slab: 72 KB
constexpr int U2_ZT_PITCH = 9;                   // float4 slots per z position: 8 channel quads + 1 pad (bank spreading)
constexpr int U2_ZT_ROW = 32 * U2_ZT_PITCH;      // float4 slots per z row (32 z columns)
constexpr int U2_ZT_BYTES = 16 * U2_ZT_ROW * 16; // 72 KB
constexpr int U2_BAR_BYTES = 256;
constexpr int U2_STG_BYTES = 3 * 64 * 4;
constexpr int U2_SMEM_BYTES = 1024 + 4 * U2_A_PLANE + U2_W_BYTES + U2_ZT_BYTES + U2_BAR_BYTES + U2_STG_BYTES;
static_assert(U2_SMEM_BYTES <= 232448, "conv_up2 shared memory exceeds the 227 KB opt-in limit");
constexpr int U2_ROWS_PER_NBLK = 9 * 64;         // packed weight rows per block of 64 output channels

// shift groups in issue order: N, TMEM column offset, A start offset (tile positions), weight region offset, 64-row units
__host__ __device__ constexpr int u2_n(int sg) { return sg == 0 ? 256 : (sg == 3 ? 64 : 128); }
__host__ __device__ constexpr int u2_col(int sg) { return sg < 2 ? 0 : 64; }
__host__ __device__ constexpr int u2_aoff(int sg) { return sg == 0 ? U2_P + 1 : (sg == 1 ? U2_P : (sg == 2 ? 1 : 0)); }
__host__ __device__ constexpr int u2_units(int sg) { return sg == 0 ? 4 : (sg == 3 ? 1 : 2); }
__host__ __device__ constexpr int u2_ubase(int sg) { return sg == 0 ? 0 : (sg == 1 ? 4 : (sg == 2 ? 6 : 8)); }

__device__ __forceinline__ uint32_t u2_cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void u2_cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// the box lands at the same shared-memory offset in BOTH CTAs of the pair and its bytes are credited to the barrier at the same
// offset in both
__device__ __forceinline__ void u2_tma_load_2d_mc(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"((uint16_t)3)
        : "memory");
}
// arrives on the barrier at the same offset in BOTH CTAs once every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void u2_commit_both(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3)
                 : "memory");
}
// Work distribution.  Without clusters a CTA walks tiles t = blockIdx.x, + gridDim.x, ... with t = m * nblk + nb.  With
// clusters the PAIR walks virtual tiles v = (m-pair) * nblk + nb and CTA `rank` takes spatial tile m = 2 * (m-pair) + rank, so
// both CTAs need the same weights at the same time; an odd last spatial tile leaves one CTA with an invalid tile that it still
// runs (zero operands, nothing stored) to keep the weight pipeline in lock step.
struct U2Walk {
    int first, step, total, rank, cluster;
};
__device__ __forceinline__ U2Walk u2_walk(const Up2Params& P) {
    U2Walk w;
    w.cluster = P.cluster;
    w.rank = P.cluster ? (int)u2_cluster_ctarank() : 0;
    w.first = P.cluster ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    w.step = P.cluster ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    w.total = P.total;
    return w;
}
__device__ __forceinline__ bool u2_decode(const Up2Params& P, const U2Walk& w, int v, int& nb, int& kx, int& ky, int& n) {
    int m = v / P.nblk;
    nb = v - m * P.nblk;
    if (w.cluster) m = 2 * m + w.rank;
    const bool valid = m < P.total_m;
    kx = m % P.tiles_x;
    m /= P.tiles_x;
    ky = m % P.tiles_y;
    n = m / P.tiles_y;
    return valid;
}

__device__ __forceinline__ float4 f4_fma(float a, const float4& x, const float4& acc) {
    return make_float4(fmaf(a, x.x, acc.x), fmaf(a, x.y, acc.y), fmaf(a, x.z, acc.z), fmaf(a, x.w, acc.w));
}
__device__ __forceinline__ float4 f4_mul(float a, const float4& x) { return make_float4(a * x.x, a * x.y, a * x.z, a * x.w); }

template <bool WIDE>
__global__ void __launch_bounds__(U2Cfg<WIDE>::THREADS, 1)
conv_up2_kernel(const __grid_constant__ Up2Tmaps maps, const __grid_constant__ Up2Params P, const EpiParams epi) {
    extern __shared__ uint8_t smem_raw[];
    // aligned by pointer arithmetic on the __shared__ array (not an integer round trip), so that the compiler keeps the
    // shared address space and emits LDS/STS with immediate offsets for the z tile
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* a_base = smem;                                  // [buf 2][plane hi/lo][U2_A_PX][128 B]
    uint8_t* w_base = smem + 4 * U2_A_PLANE;                 // 9 units of [64][128 B]: S0 (4) | S1 (2) | S2 (2) | S3 (1)
    float* zt = reinterpret_cast<float*>(w_base + U2_W_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(zt) + U2_ZT_BYTES);
    uint64_t* a_full = bars;            // [2]
    uint64_t* a_empty = bars + 2;       // [2]
    uint64_t* w_full = bars + 4;        // [4] one per shift group region
    uint64_t* w_empty = bars + 8;       // [4]
    uint64_t* t_full = bars + 12;       // [2]
    uint64_t* t_empty = bars + 14;      // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
    volatile int* epi_progress = reinterpret_cast<volatile int*>(bars + 17);   // index of the tile the epilogue is working on
    float* stg = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + U2_BAR_BYTES);   // [3][64]

    constexpr int U2_THREADS = U2Cfg<WIDE>::THREADS, U2_EPI_THREADS = U2Cfg<WIDE>::EPI_THREADS;
    constexpr int U2_REGS_DEC = U2Cfg<WIDE>::REGS_DEC, U2_REGS_INC = U2Cfg<WIDE>::REGS_INC;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kslabs = P.C / U2_KC;
    const int nplanes = P.passes == 3 ? 2 : 1;

    if (warp == 0 && lane == 0) {
        *epi_progress = -1;
        prefetch_tmap(&maps.a_hi); prefetch_tmap(&maps.a_lo); prefetch_tmap(&maps.w_hi); prefetch_tmap(&maps.w_lo);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&a_full[s], 1);
            mbar_init(&a_empty[s], 1);
            mbar_init(&t_full[s], 1);
            mbar_init(&t_empty[s], U2_EPI_THREADS);
        }
        for (int s = 0; s < 4; ++s) {
            mbar_init(&w_full[s], 1);
            mbar_init(&w_empty[s], P.cluster ? 2 : 1);      // a weight region is free when BOTH CTAs' MMAs have read it
        }
        mbar_fence_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (P.cluster) u2_cluster_sync_all();        // the peer's barriers are initialised before anything signals them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const U2Walk walk = u2_walk(P);

    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(U2_REGS_DEC));
        if (warp == 2) {
            // ===================== halo-tile TMA producer =====================
            if (elect_one()) {
                int buf = 0;
                uint32_t phase = 0;
                const uint32_t tx_bytes = (uint32_t)nplanes * (uint32_t)(U2_A_BOX_ROWS * 128);
                for (int tile = walk.first; tile < walk.total; tile += walk.step) {
                    int nb, kx, ky, n;
                    u2_decode(P, walk, tile, nb, kx, ky, n);      // an invalid tile has n >= N: the TMA unit zero-fills the box
                    const int x0 = U2_STEP_J * kx - 2, y0 = U2_STEP_I * ky - 2;     // box origin = tile origin - 1 (halo)
                    for (int ks = 0; ks < kslabs; ++ks) {
                        mbar_wait(&a_empty[buf], phase ^ 1);
                        uint8_t* sa = a_base + buf * 2 * U2_A_PLANE;
                        mbar_expect_tx(&a_full[buf], tx_bytes);
                        tma_load_4d(sa, &maps.a_hi, &a_full[buf], ks * U2_KC, x0, y0, n);
                        if (nplanes == 2) tma_load_4d(sa + U2_A_PLANE, &maps.a_lo, &a_full[buf], ks * U2_KC, x0, y0, n);
                        buf ^= 1;
                        if (buf == 0) phase ^= 1;
                    }
                }
            }
        } else if (warp == 0) {
            // ===================== weight TMA producer =====================
            if (elect_one()) {
                uint32_t phase = 0;          // the four regions are used in lock step: one phase bit serves all
                for (int tile = walk.first; tile < walk.total; tile += walk.step) {
                    const int nb = tile % P.nblk;
                    for (int ks = 0; ks < kslabs; ++ks) {
                        for (int h = 0; h < nplanes; ++h) {
                            const CUtensorMap* wm = h == 0 ? &maps.w_hi : &maps.w_lo;
#pragma unroll 1
                            for (int sg = 0; sg < 4; ++sg) {
                                mbar_wait(&w_empty[sg], phase ^ 1);
                                // the whole region arrives in this CTA's shared memory whoever loads it
                                mbar_expect_tx(&w_full[sg], (uint32_t)(P.sg_units[sg] * U2_W_UNIT));
                                for (int u = 0; u < P.sg_units[sg]; ++u) {
                                    const int unit = P.sg_ubase[sg] + u;
                                    if (!walk.cluster)
                                        tma_load_2d(w_base + unit * U2_W_UNIT, wm, &w_full[sg], ks * U2_KC, nb * U2_ROWS_PER_NBLK + unit * 64);
                                    else if ((unit & 1) == walk.rank)       // units 0,2,4,6,8 by rank 0; 1,3,5,7 by rank 1
                                        u2_tma_load_2d_mc(w_base + unit * U2_W_UNIT, wm, &w_full[sg], ks * U2_KC, nb * U2_ROWS_PER_NBLK + unit * 64);
                                }
                            }
                            phase ^= 1;
                        }
                    }
                }
            }
        } else if (warp == 3) {
            // ===================== L2 prefetcher of the epilogue's global operands =====================
            // The blur warps consume the skip planes (feats[res]) and the noise of a tile's 12 x 26 outputs straight from
            // global memory; this otherwise idle warp pulls the NEXT tile's lines into L2 so that those loads are L2 hits.
            if (epi.skip_hi || epi.noise) {
                int k = 1;
                for (int tile = walk.first + walk.step; tile < walk.total; tile += walk.step, ++k) {
                    while (*epi_progress < k - 1) __nanosleep(500);     // stay exactly one tile ahead of the epilogue
                    int nb, kx, ky, n;
                    if (!u2_decode(P, walk, tile, nb, kx, ky, n)) continue;
                    for (int i = lane; i < U2_OWN_Y * U2_OWN_X; i += 32) {
                        const int oy = i / U2_OWN_X, oxx = i - oy * U2_OWN_X;
                        const int y = U2_OWN_Y * ky + oy, x = U2_OWN_X * kx + oxx;
                        if (y < P.OH && x < P.OW) {
                            const long long el = (((long long)n * P.OH + y) * P.OW + x) * P.Co + nb * 64;
                            if (epi.skip_hi) {
                                asm volatile("prefetch.global.L2 [%0];" ::"l"(epi.skip_hi + el));
                                asm volatile("prefetch.global.L2 [%0];" ::"l"(epi.skip_lo + el));
                            }
                            if (epi.noise && (oxx & 7) == 0 && nb == 0)
                                asm volatile("prefetch.global.L2 [%0];" ::"l"(epi.noise + (long long)n * epi.noise_sn + (long long)y * P.OW + x));
                        }
                    }
                }
            }
        } else if (warp == 1) {
            // ===================== MMA issuer =====================
            if (elect_one()) {
                constexpr uint32_t DESC_HI = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);   // SBO 1024 B, version 1, SWIZZLE_128B
                const uint32_t a_lo0 = ((smem_u32(a_base) & 0x3FFFF) >> 4) | (1u << 16);
                const uint32_t w_lo0 = ((smem_u32(w_base) & 0x3FFFF) >> 4) | (1u << 16);
                constexpr uint32_t A_PLANE16 = U2_A_PLANE >> 4, K16 = 32 >> 4;
                int buf = 0, acc = 0;
                uint32_t a_phase = 0, w_phase = 0, acc_phase = 0;
                for (int tile = walk.first; tile < walk.total; tile += walk.step) {
                    for (int ks = 0; ks < kslabs; ++ks) {
                        const bool chunk_first = ks % P.chunk_slabs == 0, chunk_last = (ks + 1) % P.chunk_slabs == 0;
                        if (chunk_first) mbar_wait(&t_empty[acc], acc_phase ^ 1);
                        mbar_wait(&a_full[buf], a_phase);
                        tc_fence_after();
                        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 256);
                        const uint32_t ah = a_lo0 + (uint32_t)buf * 2u * A_PLANE16, al = ah + A_PLANE16;
#pragma unroll 1
                        for (int h = 0; h < nplanes; ++h) {
#pragma unroll 1
                            for (int sg = 0; sg < 4; ++sg) {
                                const uint32_t idesc = P.sg_idesc[sg];
                                const uint32_t aoff16 = P.sg_aoff16[sg];
                                const uint32_t wb = w_lo0 + P.sg_wb16[sg];
                                const uint32_t dcol = d_tmem + P.sg_col[sg];
                                mbar_wait(&w_full[sg], w_phase);
                                tc_fence_after();
#pragma unroll
                                for (int k = 0; k < U2_KC / 16; ++k) {
                                    const uint64_t db = ((uint64_t)DESC_HI << 32) | (wb + k * K16);
                                    umma_f16(dcol, ((uint64_t)DESC_HI << 32) | (ah + aoff16 + k * K16), db, idesc, (h | sg | k) != 0 || !chunk_first);
                                    if (h == 0 && nplanes == 2)
                                        umma_f16(dcol, ((uint64_t)DESC_HI << 32) | (al + aoff16 + k * K16), db, idesc, 1);
                                }
                                if (walk.cluster) u2_commit_both(&w_empty[sg]);
                                else umma_commit(&w_empty[sg]);
                            }
                            w_phase ^= 1;
                        }
                        umma_commit(&a_empty[buf]);
                        buf ^= 1;
                        if (buf == 0) a_phase ^= 1;
                        if (chunk_last) {
                            umma_commit(&t_full[acc]);
                            acc ^= 1;
                            if (acc == 0) acc_phase ^= 1;
                        }
                    }
                }
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(U2_REGS_INC));
        if constexpr (WIDE) {
        // ===================== epilogue, 16 warps (warps 4..19), one TMEM chunk per tile =====================
        const int e = (int)threadIdx.x - (U2_THREADS - U2_EPI_THREADS);
        const int q = warp & 3;                 // TMEM lane quarter
        const int pb = (warp - 4) >> 2;         // parity block = 64 accumulator columns: (1,0) (0,0) (0,1) (1,1)
        const int row = q * 32 + lane;          // tile position
        const int r = row >> 4, c = row & 15;
        const int py = (pb == 0 || pb == 3) ? 1 : 0, px = pb >> 1;
        // chained MMAs per column of my block: taps {2,4,2,1} x 4 k-steps x passes x slabs
        const float comp = 1.f + P.acc_comp * (float)((P.passes == 3 ? 12 : 4) * P.chunk_slabs) * (pb == 1 ? 4.f : (pb == 3 ? 1.f : 2.f));
        float4* const zt4 = reinterpret_cast<float4*>(zt);
        float4* const zw = zt4 + ((2 * r + py) * 32 + 2 * c + px) * U2_ZT_PITCH;
        // blur role: thread = 4 channels (quad q4) of one output column ox and one half (6 rows) of the tile's 12 output rows;
        // output (oy, ox) reads z rows oy+1 .. oy+4 and z columns ox+1 .. ox+4
        const int q4 = e & 7, ox = (e >> 3) & 31, hr = e >> 8;
        const bool blur_active = ox < U2_OWN_X;
        const float4* const zrd = zt4 + ((6 * hr + 1) * 32 + ox + 1) * U2_ZT_PITCH + q4;
        const float gain_e = epi.act_gain;
        const float alpha_e = epi.act ? epi.act_alpha : 1.f;
        const float clamp_e = (epi.act && epi.act_clamp > 0.f) ? epi.act_clamp : __int_as_float(0x7f800000);
        const float nstr = epi.noise ? __ldg(epi.noise_strength) * gain_e : 0.f;
        const bool has_skip = epi.skip_hi != nullptr, has_noise = epi.noise != nullptr, has_f32 = epi.out_f32 != nullptr,
                   has_planes = epi.out_hi != nullptr;
        const int row_elems = P.OW * P.Co;
        int acc = 0;
        uint32_t acc_phase = 0;
        int tile_iter = 0;
        for (int tile = walk.first; tile < walk.total; tile += walk.step, ++tile_iter) {
            int nb, kx, ky, n;
            const bool tile_valid = u2_decode(P, walk, tile, nb, kx, ky, n);
            if (!tile_valid) n = 0;            // (keeps the staged per-sample vectors in bounds; nothing of this tile is stored)
            if (e == 0) *epi_progress = tile_iter;
            const int y0 = U2_OWN_Y * ky + 6 * hr, x = U2_OWN_X * kx + ox;     // my first output row, my output column
            const int rows_valid = P.OH - y0 < 6 ? P.OH - y0 : 6;              // (may be <= 0)
            const bool col_valid = tile_valid && blur_active && x < P.OW && rows_valid > 0;

            mbar_wait(&t_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 256 + pb * 64);
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                // the z tile / staging area is free again (previous round's or tile's readers are done)
                asm volatile("bar.sync 1, %0;" ::"n"(U2_EPI_THREADS) : "memory");
                if (g == 0 && e < 64) {
                    const int o = nb * 64 + e;
                    const long long no = (long long)n * P.Co + o;
                    stg[e] = (epi.dcoef ? __ldg(epi.dcoef + no) : 1.f) * epi.wgain * gain_e;
                    stg[64 + e] = epi.bias ? __ldg(epi.bias + o) * gain_e : 0.f;
                    stg[128 + e] = epi.next_scale ? __ldg(epi.next_scale + no) : 1.f;
                }
                {
                    float v[32];
                    tmem_ld32(taddr + g * 32, v);
#pragma unroll
                    for (int qq = 0; qq < 8; ++qq)
                        zw[qq] = make_float4(v[4 * qq] * comp, v[4 * qq + 1] * comp, v[4 * qq + 2] * comp, v[4 * qq + 3] * comp);
                }
                if (g == 1) {               // the accumulator buffer is drained: the MMAs of the next-but-one tile may overwrite it
                    tc_fence_before();
                    mbar_arrive(&t_empty[acc]);
                }
                asm volatile("bar.sync 1, %0;" ::"n"(U2_EPI_THREADS) : "memory");
                if (!col_valid) continue;

                const int lc = g * 32 + q4 * 4;                      // my 4 channels inside this block of 64
                const float4 sd = *reinterpret_cast<const float4*>(stg + lc);
                const float4 sb = *reinterpret_cast<const float4*>(stg + 64 + lc);
                const float4 sn = *reinterpret_cast<const float4*>(stg + 128 + lc);
                const int eo = ((n * P.OH + y0) * P.OW + x) * P.Co + nb * 64 + lc;      // (n, y0, x, first channel)
                const float* p_nz = epi.noise + ((long long)n * epi.noise_sn + (long long)y0 * P.OW + x);
                auto hrow = [&](const float4* zp) -> float4 {
                    const float4 l0 = zp[0], l1 = zp[U2_ZT_PITCH], l2 = zp[2 * U2_ZT_PITCH], l3 = zp[3 * U2_ZT_PITCH];
                    return f4_fma(P.fx[3], l3, f4_fma(P.fx[2], l2, f4_fma(P.fx[1], l1, f4_mul(P.fx[0], l0))));
                };
                // skip planes / noise of a row are requested two rows before the row that consumes them
                uint2 sh[6], sl[6];
                float sz[6];
                auto fetch = [&](int j) {
                    sh[j] = make_uint2(0u, 0u); sl[j] = make_uint2(0u, 0u); sz[j] = 0.f;
                    if (j < rows_valid) {
                        if (has_skip) {
                            sh[j] = __ldg(reinterpret_cast<const uint2*>(epi.skip_hi + eo + j * row_elems));
                            sl[j] = __ldg(reinterpret_cast<const uint2*>(epi.skip_lo + eo + j * row_elems));
                        }
                        if (has_noise) sz[j] = __ldg(p_nz + j * P.OW);
                    }
                };
                fetch(0); fetch(1);
                float4 h[4];
                h[0] = hrow(zrd); h[1] = hrow(zrd + U2_ZT_ROW); h[2] = hrow(zrd + 2 * U2_ZT_ROW);
#pragma unroll
                for (int j = 0; j < 6; ++j) {
                    if (j + 2 < 6) fetch(j + 2);
                    h[(j + 3) & 3] = hrow(zrd + (j + 3) * U2_ZT_ROW);
                    if (j < rows_valid) {
                        const float4 o = f4_fma(P.fy[3], h[(j + 3) & 3], f4_fma(P.fy[2], h[(j + 2) & 3], f4_fma(P.fy[1], h[(j + 1) & 3], f4_mul(P.fy[0], h[j & 3]))));
                        const float nzs = sz[j] * nstr;
                        float v0 = fmaf(o.x, sd.x, nzs) + sb.x, v1 = fmaf(o.y, sd.y, nzs) + sb.y;
                        float v2 = fmaf(o.z, sd.z, nzs) + sb.z, v3 = fmaf(o.w, sd.w, nzs) + sb.w;
                        v0 = fminf(fmaxf(fmaxf(v0, v0 * alpha_e), -clamp_e), clamp_e);
                        v1 = fminf(fmaxf(fmaxf(v1, v1 * alpha_e), -clamp_e), clamp_e);
                        v2 = fminf(fmaxf(fmaxf(v2, v2 * alpha_e), -clamp_e), clamp_e);
                        v3 = fminf(fmaxf(fmaxf(v3, v3 * alpha_e), -clamp_e), clamp_e);
                        {
                            const float2 a0 = __half22float2(*reinterpret_cast<const __half2*>(&sh[j].x));
                            const float2 a1 = __half22float2(*reinterpret_cast<const __half2*>(&sh[j].y));
                            const float2 b0 = __half22float2(*reinterpret_cast<const __half2*>(&sl[j].x));
                            const float2 b1 = __half22float2(*reinterpret_cast<const __half2*>(&sl[j].y));
                            v0 += a0.x + b0.x; v1 += a0.y + b0.y; v2 += a1.x + b1.x; v3 += a1.y + b1.y;
                        }
                        const int eoj = eo + j * row_elems;
                        if (has_f32) *reinterpret_cast<float4*>(epi.out_f32 + eoj) = make_float4(v0, v1, v2, v3);
                        if (has_planes) {
                            v0 *= sn.x; v1 *= sn.y; v2 *= sn.z; v3 *= sn.w;
                            const __half2 h01 = __floats2half2_rn(v0, v1), h23 = __floats2half2_rn(v2, v3);
                            const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                            const __half2 l01 = __floats2half2_rn(v0 - f01.x, v1 - f01.y), l23 = __floats2half2_rn(v2 - f23.x, v3 - f23.y);
                            *reinterpret_cast<uint2*>(epi.out_hi + eoj) =
                                make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
                            *reinterpret_cast<uint2*>(epi.out_lo + eoj) =
                                make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
                        }
                    }
                }
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
        } else {
        // ===================== epilogue (warps 4..11) =====================
        const int e = (int)threadIdx.x - (U2_THREADS - U2_EPI_THREADS);
        const int q = warp & 3;                 // TMEM lane quarter
        const int half = (warp - 4) >> 2;       // column half: parity blocks {(1,0),(0,0)} or {(0,1),(1,1)}
        const int row = q * 32 + lane;          // tile position
        const int r = row >> 4, c = row & 15;
        // chained MMAs per chunk and column block: taps {2,4} (half 0) / {2,1} (half 1) x 4 k-steps x passes
        const float mm = (float)((P.passes == 3 ? 12 : 4) * P.chunk_slabs);
        const float comp0 = 1.f + P.acc_comp * mm * 2.f;
        const float comp1 = 1.f + P.acc_comp * mm * (half == 0 ? 4.f : 1.f);
        // z-tile slots of my two parity blocks: z row 2r+py, z column 2c+half; a position holds its 32 channels as 8
        // float4 quads at a pitch of 9 slots (conflict-free channel-major reads, 2-way conflicts on the position-major writes)
        const int py0 = half == 0 ? 1 : 0, py1 = 1 - py0;
        float4* const zt4 = reinterpret_cast<float4*>(zt);
        float4* const zw0 = zt4 + ((2 * r + py0) * 32 + 2 * c + half) * U2_ZT_PITCH;
        float4* const zw1 = zt4 + ((2 * r + py1) * 32 + 2 * c + half) * U2_ZT_PITCH;
        // blur role: thread = 4 channels (quad q4) of one output column ox, walking the 15 z rows of the tile top to bottom;
        // output (oy, ox) reads z rows oy+1 .. oy+4 and z columns ox+1 .. ox+4
        const int q4 = e & 7, ox = e >> 3;
        const bool blur_active = ox < U2_OWN_X;
        const float4* const zrd = zt4 + (32 + ox + 1) * U2_ZT_PITCH + q4;       // z row 1, column ox+1
        const float fx0 = P.fx[0], fx1 = P.fx[1], fx2 = P.fx[2], fx3 = P.fx[3];
        const float fy0 = P.fy[0], fy1 = P.fy[1], fy2 = P.fy[2], fy3 = P.fy[3];
        // pointwise parameters, normalised once: lrelu(u) * gain == lrelu(u * gain) for gain > 0, so the gain is folded into
        // the staged scale / bias / noise; no activation = alpha 1; no clamp = +inf
        const float gain_e = epi.act_gain;
        const float alpha_e = epi.act ? epi.act_alpha : 1.f;
        const float clamp_e = (epi.act && epi.act_clamp > 0.f) ? epi.act_clamp : __int_as_float(0x7f800000);
        const float nstr = epi.noise ? __ldg(epi.noise_strength) * gain_e : 0.f;
        const bool has_skip = epi.skip_hi != nullptr, has_noise = epi.noise != nullptr, has_f32 = epi.out_f32 != nullptr,
                   has_planes = epi.out_hi != nullptr;
        const int row_elems = P.OW * P.Co;      // elements between vertically adjacent pixels (tensor sizes are < 2^31)
        int acc = 0;
        uint32_t acc_phase = 0;
        int tile_iter = 0;
        for (int tile = walk.first; tile < walk.total; tile += walk.step, ++tile_iter) {
            int nb, kx, ky, n;
            const bool tile_valid = u2_decode(P, walk, tile, nb, kx, ky, n);
            if (!tile_valid) n = 0;            // (keeps the staged per-sample vectors in bounds; nothing of this tile is stored)
            if (e == 0) *epi_progress = tile_iter;

            float accv[128];
            for (int ks = 0; ks < kslabs; ks += P.chunk_slabs) {
                mbar_wait(&t_full[acc], acc_phase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 256 + half * 128);
#pragma unroll
                for (int p = 0; p < 8; ++p) {
                    float v[16];
                    tmem_ld16(taddr + p * 16, v);
                    const float comp = p < 4 ? comp0 : comp1;
#pragma unroll
                    for (int i = 0; i < 16; ++i) accv[p * 16 + i] = ks == 0 ? v[i] * comp : fmaf(v[i], comp, accv[p * 16 + i]);
                }
                tc_fence_before();
                mbar_arrive(&t_empty[acc]);
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }

            const int y0 = U2_OWN_Y * ky, x = U2_OWN_X * kx + ox;       // first output row of the tile, my output column
            const int rows_valid = P.OH - y0 < U2_OWN_Y ? P.OH - y0 : U2_OWN_Y;
            const bool col_valid = tile_valid && blur_active && x < P.OW;
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                // the z tile / staging area is free again (previous round's or tile's readers are done)
                asm volatile("bar.sync 1, %0;" ::"n"(U2_EPI_THREADS) : "memory");
                if (g == 0 && e < 64) {
                    const int o = nb * 64 + e;
                    const long long no = (long long)n * P.Co + o;
                    stg[e] = (epi.dcoef ? __ldg(epi.dcoef + no) : 1.f) * epi.wgain * gain_e;
                    stg[64 + e] = epi.bias ? __ldg(epi.bias + o) * gain_e : 0.f;
                    stg[128 + e] = epi.next_scale ? __ldg(epi.next_scale + no) : 1.f;
                }
#pragma unroll
                for (int qq = 0; qq < 8; ++qq) {
                    const int i0 = g * 32 + qq * 4;
                    zw0[qq] = make_float4(accv[i0], accv[i0 + 1], accv[i0 + 2], accv[i0 + 3]);
                    zw1[qq] = make_float4(accv[64 + i0], accv[64 + i0 + 1], accv[64 + i0 + 2], accv[64 + i0 + 3]);
                }
                asm volatile("bar.sync 1, %0;" ::"n"(U2_EPI_THREADS) : "memory");
                if (!col_valid) continue;

                const int lc = g * 32 + q4 * 4;                      // my 4 channels inside this block of 64
                const float4 sd = *reinterpret_cast<const float4*>(stg + lc);
                const float4 sb = *reinterpret_cast<const float4*>(stg + 64 + lc);
                const float4 sn = *reinterpret_cast<const float4*>(stg + 128 + lc);
                // element index of (n, y0, x, first channel); one output row further = + row_elems
                int eo = ((n * P.OH + y0) * P.OW + x) * P.Co + nb * 64 + lc;
                const float* p_nz = epi.noise + ((long long)n * epi.noise_sn + (long long)y0 * P.OW + x);

                // horizontal 4-tap pass of one z row
                auto hrow = [&](const float4* zp) -> float4 {
                    const float4 l0 = zp[0], l1 = zp[U2_ZT_PITCH], l2 = zp[2 * U2_ZT_PITCH], l3 = zp[3 * U2_ZT_PITCH];
                    return f4_fma(fx3, l3, f4_fma(fx2, l2, f4_fma(fx1, l1, f4_mul(fx0, l0))));
                };
                // the skip planes and the noise are fetched one block of 4 output rows ahead (raw values; nothing waits on them
                // before the block that consumes them)
                uint2 ch_[4], cl_[4], nh_[4], nl_[4];
                float cn_[4], nn_[4];
                auto prefetch = [&](int i0, int eoff, uint2* sh, uint2* sl, float* sz) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        sh[j] = make_uint2(0u, 0u); sl[j] = make_uint2(0u, 0u); sz[j] = 0.f;
                        if (i0 + j < rows_valid) {
                            if (has_skip) {
                                sh[j] = __ldg(reinterpret_cast<const uint2*>(epi.skip_hi + eoff + j * row_elems));
                                sl[j] = __ldg(reinterpret_cast<const uint2*>(epi.skip_lo + eoff + j * row_elems));
                            }
                            if (has_noise) sz[j] = __ldg(p_nz + (i0 + j) * P.OW);
                        }
                    }
                };
                prefetch(0, eo, ch_, cl_, cn_);
                const float4* zp = zrd;
                float4 h0 = hrow(zp), h1 = hrow(zp + U2_ZT_ROW), h2 = hrow(zp + 2 * U2_ZT_ROW), h3;
                zp += 3 * U2_ZT_ROW;
#pragma unroll 1
                for (int blk = 0; blk < 3; ++blk) {
                    if (blk < 2) prefetch(4 * blk + 4, eo + 4 * row_elems, nh_, nl_, nn_);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        // ring of the last four horizontal rows, rotated statically over the 4 unrolled steps
                        float4& ha = j == 0 ? h0 : (j == 1 ? h1 : (j == 2 ? h2 : h3));
                        float4& hb = j == 0 ? h1 : (j == 1 ? h2 : (j == 2 ? h3 : h0));
                        float4& hc = j == 0 ? h2 : (j == 1 ? h3 : (j == 2 ? h0 : h1));
                        float4& hd = j == 0 ? h3 : (j == 1 ? h0 : (j == 2 ? h1 : h2));
                        hd = hrow(zp + j * U2_ZT_ROW);
                        if (4 * blk + j < rows_valid) {
                            const float4 o = f4_fma(fy3, hd, f4_fma(fy2, hc, f4_fma(fy1, hb, f4_mul(fy0, ha))));
                            const float nzs = cn_[j] * nstr;
                            float v0 = fmaf(o.x, sd.x, nzs) + sb.x, v1 = fmaf(o.y, sd.y, nzs) + sb.y;
                            float v2 = fmaf(o.z, sd.z, nzs) + sb.z, v3 = fmaf(o.w, sd.w, nzs) + sb.w;
                            v0 = fminf(fmaxf(fmaxf(v0, v0 * alpha_e), -clamp_e), clamp_e);
                            v1 = fminf(fmaxf(fmaxf(v1, v1 * alpha_e), -clamp_e), clamp_e);
                            v2 = fminf(fmaxf(fmaxf(v2, v2 * alpha_e), -clamp_e), clamp_e);
                            v3 = fminf(fmaxf(fmaxf(v3, v3 * alpha_e), -clamp_e), clamp_e);
                            {
                                const float2 a0 = __half22float2(*reinterpret_cast<const __half2*>(&ch_[j].x));
                                const float2 a1 = __half22float2(*reinterpret_cast<const __half2*>(&ch_[j].y));
                                const float2 b0 = __half22float2(*reinterpret_cast<const __half2*>(&cl_[j].x));
                                const float2 b1 = __half22float2(*reinterpret_cast<const __half2*>(&cl_[j].y));
                                v0 += a0.x + b0.x; v1 += a0.y + b0.y; v2 += a1.x + b1.x; v3 += a1.y + b1.y;
                            }
                            const int eoj = eo + j * row_elems;
                            if (has_f32) *reinterpret_cast<float4*>(epi.out_f32 + eoj) = make_float4(v0, v1, v2, v3);
                            if (has_planes) {
                                v0 *= sn.x; v1 *= sn.y; v2 *= sn.z; v3 *= sn.w;
                                const __half2 h01 = __floats2half2_rn(v0, v1), h23 = __floats2half2_rn(v2, v3);
                                const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                                const __half2 l01 = __floats2half2_rn(v0 - f01.x, v1 - f01.y), l23 = __floats2half2_rn(v2 - f23.x, v3 - f23.y);
                                *reinterpret_cast<uint2*>(epi.out_hi + eoj) =
                                    make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
                                *reinterpret_cast<uint2*>(epi.out_lo + eoj) =
                                    make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
                            }
                        }
                    }
                    zp += 4 * U2_ZT_ROW;
                    eo += 4 * row_elems;
#pragma unroll
                    for (int j = 0; j < 4; ++j) { ch_[j] = nh_[j]; cl_[j] = nl_[j]; cn_[j] = nn_[j]; }
                }
            }
        }
        }   // !WIDE
    }

    tc_fence_before();
    __syncthreads();
    if (P.cluster) u2_cluster_sync_all();        // no CTA leaves while the peer may still write its shared memory / signal its barriers
    tc_fence_after();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

}  // namespace shgan

using namespace shgan;

extern "C" int shgan_conv_up2(const shgan_up2_desc* d, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SHGAN_CHECK(d, "null descriptor");
    SHGAN_CHECK(d->src_hi && d->src_lo && d->w_hi && d->w_lo, "null operand");
    SHGAN_CHECK(d->N >= 0 && d->H >= 1 && d->W >= 1, "bad input size");
    SHGAN_CHECK(d->C >= 64 && d->C % 64 == 0 && d->Co >= 64 && d->Co % 64 == 0, "C and Co must be multiples of 64");
    SHGAN_CHECK((long long)d->N * d->C * d->H * d->W <= INT32_MAX, "input tensor is too large");
    SHGAN_CHECK(4LL * d->N * d->Co * d->H * d->W <= INT32_MAX, "output tensor is too large");
    const int passes_arg = d->passes & ~(SHGAN_UP2_NARROW | SHGAN_UP2_CLUSTER | SHGAN_UP2_NO_CLUSTER);
    SHGAN_CHECK(passes_arg == 0 || passes_arg == 1 || passes_arg == 3, "passes must be 0, 1 or 3");
    if (const char* m = check_epi(d->epi, d->Co)) SHGAN_CHECK(false, m);
    SHGAN_CHECK(!d->epi.rgb_w, "the up-sampling convolution has no fused torgb");
    if (d->N == 0) return 0;

    Up2Params P;
    P.N = d->N; P.H = d->H; P.W = d->W; P.C = d->C; P.Co = d->Co;
    P.OH = 2 * d->H; P.OW = 2 * d->W;
    P.tiles_x = ceil_div(P.OW, U2_OWN_X);
    P.tiles_y = ceil_div(P.OH, U2_OWN_Y);
    P.nblk = d->Co / 64;
    const long long total_m = (long long)P.tiles_x * P.tiles_y * d->N;
    SHGAN_CHECK(total_m * P.nblk <= INT32_MAX, "too many tiles");
    P.total_m = (int)total_m;
    P.cluster = 0;
    P.total = (int)(total_m * P.nblk);
    for (int i = 0; i < 4; ++i) { P.fx[i] = d->fx[i]; P.fy[i] = d->fy[i] * d->gain; }
    P.acc_comp = d->acc_comp == 0.f ? SHGAN_ACC_COMP_DEFAULT : (d->acc_comp < 0.f ? 0.f : d->acc_comp);
    P.passes = passes_arg == 0 ? 3 : passes_arg;
    // C >= 256: four slabs (<= 192 chained MMAs per column, compensated by acc_comp) per TMEM chunk, so that the two
    // accumulator buffers let the MMAs run a whole 512-channel tile ahead of the epilogue warps
    P.chunk_slabs = (d->C / 64) % 4 == 0 ? 4 : ((d->C / 64) % 2 == 0 && d->C >= 256 ? 2 : 1);
    // C <= 256: the whole tile is one chunk (<= 192 chained MMAs per column) and the 16-warp epilogue instance runs
    const bool wide = d->C <= 256 && !(d->passes & SHGAN_UP2_NARROW);
    if (wide) P.chunk_slabs = d->C / 64;
    for (int sg = 0; sg < 4; ++sg) {
        P.sg_idesc[sg] = (1u << 4) | ((uint32_t)(u2_n(sg) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        P.sg_col[sg] = (uint32_t)u2_col(sg);
        P.sg_aoff16[sg] = (uint32_t)u2_aoff(sg) * (128 / 16);
        P.sg_wb16[sg] = (uint32_t)(u2_ubase(sg) * U2_W_UNIT) >> 4;
        P.sg_units[sg] = u2_units(sg);
        P.sg_ubase[sg] = u2_ubase(sg);
    }

    Up2Tmaps maps;
    const uint64_t adims[4] = {(uint64_t)d->C, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->N};
    const uint32_t abox[4] = {(uint32_t)U2_KC, (uint32_t)U2_P, (uint32_t)(U2_ROWS + 1), 1u};
    if (int e = encode_tmap(&maps.a_hi, d->src_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, 4, adims, abox, CU_TENSOR_MAP_SWIZZLE_128B)) return e;
    if (int e = encode_tmap(&maps.a_lo, d->src_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, 4, adims, abox, CU_TENSOR_MAP_SWIZZLE_128B)) return e;
    const uint64_t wdims[2] = {(uint64_t)d->C, (uint64_t)P.nblk * U2_ROWS_PER_NBLK};
    const uint32_t wbox[2] = {(uint32_t)U2_KC, 64u};
    if (int e = encode_tmap(&maps.w_hi, d->w_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, 2, wdims, wbox, CU_TENSOR_MAP_SWIZZLE_128B)) return e;
    if (int e = encode_tmap(&maps.w_lo, d->w_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, 2, wdims, wbox, CU_TENSOR_MAP_SWIZZLE_128B)) return e;

    static DeviceInit once;
    int num_sms = 0;
    if (int e = device_init(once, &num_sms, []() -> int {
            SHGAN_CUDA(cudaFuncSetAttribute(conv_up2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, U2_SMEM_BYTES));
            SHGAN_CUDA(cudaFuncSetAttribute(conv_up2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, U2_SMEM_BYTES));
            return 0;
        })) return e;
    // Clusters of two CTAs that share each weight load by TMA multicast: only on request.  Measured on B200 (batch 16, C = 256 / 512
    // layers, profiles/r2_up2_findings.md): 0.695 / 0.646 / 0.431 ms with, 0.699 / 0.654 / 0.427 ms without -- halving the weight
    // requests changes nothing, so these layers are bound by the LATENCY of the single-buffered weight regions, not by L2 -> SM bytes.
    const bool cluster = (d->passes & SHGAN_UP2_CLUSTER) && !(d->passes & SHGAN_UP2_NO_CLUSTER);
    int grid = P.total < num_sms ? P.total : num_sms;
    if (cluster) {
        P.cluster = 1;
        P.total = (int)((total_m + 1) / 2) * P.nblk;        // virtual tiles walked by a pair
        const int pairs = P.total < num_sms / 2 ? P.total : num_sms / 2;
        grid = 2 * pairs;
    }
    const EpiParams epi = make_epi(d->epi);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid, 1, 1);
    cfg.blockDim = dim3((unsigned)(wide ? U2Cfg<true>::THREADS : U2Cfg<false>::THREADS), 1, 1);
    cfg.dynamicSmemBytes = U2_SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = cluster ? 1 : 0;
    if (wide) SHGAN_CUDA(cudaLaunchKernelEx(&cfg, conv_up2_kernel<true>, maps, P, epi));
    else SHGAN_CUDA(cudaLaunchKernelEx(&cfg, conv_up2_kernel<false>, maps, P, epi));
    SHGAN_LAUNCH_CHECK();
    return 0;
}
