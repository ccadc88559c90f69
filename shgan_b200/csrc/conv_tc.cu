// Single-CTA, per-tap tcgen05 tensor-core implementation of shgan_conv_igemm (replaces cuDNN conv / conv_transpose
// reached from conv2d_resample.py:26-51 and the per-sample weight materialisation of stylegan.py:149-190).  Of the
// three tensor-core kernels behind that entry point (conv_api.cu picks per layer) this one serves the Co = 64
// transposed-conv passes, the 4x4 .. 16x16 layers (multi-image tiles) and the SHU's channel mix; the two-SM kernel
// (conv_pair.cu) takes the Co % 128 == 0 layers, the halo kernel (conv_halo.cu) the 64 -> 64 layers at >= 256^2.
//
// Im2col-free implicit GEMM, one CTA per SM, persistent over output tiles:
//   D[128 pixels, BN out-channels] += A_tap[128 pixels, 64 ch] * W_tap[BN, 64 ch]^T   for every tap and 64-ch slab
// * A_tap is fetched straight from the NHWC split planes by a 4-D TMA box {64 ch, TW, TH, TN} whose
//   (x,y) origin is shifted by the tap offset; out-of-image pixels are zero-filled by the TMA unit, which
//   is the convolution's zero padding.  The box lands in shared memory as 128 rows of 128 B in the
//   SWIZZLE_128B K-major layout that tcgen05.mma consumes directly -- no im2col buffer, no register staging.
// * W_tap is a 2-D TMA box {64 ch, BN} of the pre-packed [tap, Co, C] fp16 weights, same layout.
// * fp32-class accuracy on fp16 tensor cores: activations and weights are held as fp16 hi+lo pairs
//   and each slab issues hi*hi + lo*hi + hi*lo into the same fp32 TMEM accumulator (the dropped lo*lo
//   term is < 2^-22 relative).  passes == 1 issues hi*hi only.
// * Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one elected lane),
//   warps 4-11 = epilogue (tcgen05.ld 32 lanes x 32 columns per instruction -> registers -> fused
//   demod / noise / bias / lrelu / clamp / skip-add / torgb / next-layer modulation -> split planes);
//   epilogue warp w owns TMEM lane quarter (w & 3) and column half (w - 4) >> 2 of the tile;
//   setmaxnreg moves registers from the producer/MMA warpgroup to the two epilogue warpgroups.
//   An mbarrier ring of STAGES smem slots decouples TMA from MMA.
// * Two-level accumulation.  The tensor core truncates (does not round) when it adds into the fp32 TMEM
//   accumulator, which shows up as a bias that grows linearly with the number of chained MMAs (measured:
//   2.5e-6 relative after 108 MMAs, 1.5e-5 after 864).  The K loop is therefore cut into chunks of at most
//   4 (tap, slab) steps; each chunk accumulates in one of the 512 / BN TMEM accumulator buffers while the
//   epilogue warps drain a finished one and add it into fp32 registers with round-to-nearest FMAs.  This
//   bounds the truncation chain for every layer width and is also what overlaps epilogue and MMA.
// * The per-(sample, channel) epilogue vectors are staged in shared memory per tile for BN < 256 (conv_common.cuh),
//   plane outputs leave through 256-bit stores, and producer / issuer loops run under elect.sync (tc_ptx.cuh).
#include "conv_common.cuh"
#include "tc_ptx.cuh"

namespace shgan {

struct ConvTmaps {
    CUtensorMap a_hi[SHGAN_MAX_SRC];
    CUtensorMap a_lo[SHGAN_MAX_SRC];
    CUtensorMap w_hi;
    CUtensorMap w_lo;
};

struct TileInfo {
    int tw_log2, th_log2;       // tile = TN images x TH rows x TW cols = 128 pixels
    int TW, TH, TN;
    int tiles_x, tiles_y, tiles_n, nblk;
    int total;
    // split-K (RAW mode into per-split partial buffers, summed by conv_splitk_epilogue_kernel): ksplit CTAs share one output
    // tile, each running kit_per of the (tap, slab) iterations; ksplit == 1: off
    int ksplit, kit_per;
    long long split_stride;     // floats between partial buffers
};

constexpr int TC_THREADS = 384;      // warpgroup 0: warp 0 TMA, warp 1 MMA (2 idle); warpgroups 1-2 (warps 4..11): epilogue
constexpr int TC_EPI_THREADS = 256;
// setmaxnreg budget: the CTA is launched with 168 registers/thread (the cap ptxas applies for 384 threads); the
// re-partition must fit in that pool or setmaxnreg.inc blocks forever: 128*DEC + 256*INC <= 384*168.
constexpr int TC_REGS_LAUNCH = 168, TC_REGS_DEC = 56, TC_REGS_INC = 224;
static_assert(128 * TC_REGS_DEC + 256 * TC_REGS_INC <= 384 * TC_REGS_LAUNCH, "setmaxnreg budget exceeds the CTA register pool");
constexpr int TC_M = 128;       // output pixels per tile == UMMA M
constexpr int TC_MAX_CHUNK_ITERS = 4;   // (tap, slab) steps chained in one TMEM accumulator = 48 MMAs
constexpr int TC_KC = 64;       // channels per K slab (= 128 B of fp16 = one swizzle row)
constexpr int A_BYTES = TC_M * TC_KC * 2;
constexpr int TC_STG_VECS = CONV_STG_VECS;

template <int BN> struct TcCfg {
    static constexpr int STAGES = BN == 256 ? 2 : (BN == 128 ? 3 : 4);
    static constexpr int B_BYTES = BN * TC_KC * 2;
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    static constexpr int STG_BYTES = TC_STG_VECS * BN * 4;   // per-tile epilogue vectors (stage_epilogue_vectors)
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + STG_BYTES;
    static constexpr int TMEM_COLS = 512;      // the whole tensor memory (one CTA per SM)
    static constexpr int NACC = TMEM_COLS / BN;   // accumulator buffers: the MMA issuer may run NACC chunks ahead of the epilogue
};

// ---- kernel -------------------------------------------------------------------------------------
template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ ConvTmaps maps, const ConvGeom g, const EpiParams epi, const TileInfo ti,
               const int passes, const int chunk_iters) {
    using Cfg = TcCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS / STS, not generic LD / ST)
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
    uint64_t* full_bar = bars;                    // [STAGES] TMA -> MMA
    uint64_t* empty_bar = bars + STAGES;          // [STAGES] MMA -> TMA
    constexpr int NACC = Cfg::NACC;
    uint64_t* tfull_bar = bars + 2 * STAGES;             // [NACC] MMA -> epilogue
    uint64_t* tempty_bar = bars + 2 * STAGES + NACC;     // [NACC] epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 2 * NACC);
    float* stg = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);   // [TC_STG_VECS][BN]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kslabs = g.C / TC_KC;
    const int kiters = g.ntaps * kslabs;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < g.num_src; ++s) {
            prefetch_tmap(&maps.a_hi[s]);
            prefetch_tmap(&maps.a_lo[s]);
        }
        prefetch_tmap(&maps.w_hi);
        prefetch_tmap(&maps.w_lo);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < NACC; ++a) {
            mbar_init(&tfull_bar[a], 1);
            mbar_init(&tempty_bar[a], TC_EPI_THREADS);
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
      asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(TC_REGS_DEC));
      if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t tx_bytes = (passes == 3 ? 2u : 1u) * (uint32_t)(A_BYTES + Cfg::B_BYTES);
            for (int tile = blockIdx.x; tile < ti.total; tile += gridDim.x) {
                const int sp = tile % ti.ksplit;
                int m = (tile / ti.ksplit) / ti.nblk;
                const int nb = (tile / ti.ksplit) - m * ti.nblk;
                const int x0 = (m % ti.tiles_x) * ti.TW;
                m /= ti.tiles_x;
                const int y0 = (m % ti.tiles_y) * ti.TH;
                const int n0 = (m / ti.tiles_y) * ti.TN;
                const int kit0 = sp * ti.kit_per, kit1 = kit0 + ti.kit_per < kiters ? kit0 + ti.kit_per : kiters;
                for (int t = kit0 / kslabs; t * kslabs < kit1; ++t) {
                    const int s = g.tap_src[t];
                    const int cx = x0 + g.tap_dx[t], cy = y0 + g.tap_dy[t];
                    const int wrow = g.tap_w[t] * g.Co + nb * BN;
                    const int ks0 = t * kslabs < kit0 ? kit0 - t * kslabs : 0, ks1 = (t + 1) * kslabs > kit1 ? kit1 - t * kslabs : kslabs;
                    for (int ks = ks0; ks < ks1; ++ks) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
                        mbar_expect_tx(&full_bar[stage], tx_bytes);
                        tma_load_4d(sa, &maps.a_hi[s], &full_bar[stage], ks * TC_KC, cx, cy, n0);
                        tma_load_2d(sa + 2 * A_BYTES, &maps.w_hi, &full_bar[stage], ks * TC_KC, wrow);
                        if (passes == 3) {
                            tma_load_4d(sa + A_BYTES, &maps.a_lo[s], &full_bar[stage], ks * TC_KC, cx, cy, n0);
                            tma_load_2d(sa + 2 * A_BYTES + Cfg::B_BYTES, &maps.w_lo, &full_bar[stage], ks * TC_KC, wrow);
                        }
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            // instruction descriptor (cute::UMMA::InstrDescriptor): D fp32 [4,6)=1, A/B fp16 (0), both K-major,
            // N>>3 in [17,23), M>>4 in [24,29)
            constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < ti.total; tile += gridDim.x) {
                const int kit0s = (tile % ti.ksplit) * ti.kit_per;
                const int kloc = (kit0s + ti.kit_per < kiters ? kit0s + ti.kit_per : kiters) - kit0s;   // this CTA's share of the K loop
                for (int it0 = 0; it0 < kloc; it0 += chunk_iters) {
                    const int n_it = kloc - it0 < chunk_iters ? kloc - it0 : chunk_iters;
                    mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
                    for (int it = 0; it < n_it; ++it) {
                        mbar_wait(&full_bar[stage], phase);
                        tc_fence_after();
                        const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
                        const uint32_t a_hi = sa, a_lo = sa + A_BYTES, b_hi = sa + 2 * A_BYTES, b_lo = b_hi + Cfg::B_BYTES;
#pragma unroll
                        for (int k = 0; k < TC_KC / 16; ++k) {
                            const uint32_t ko = k * 32;  // 16 fp16 = 32 B inside the 128 B swizzle row
                            const uint64_t dah = umma_desc_sw128(a_hi + ko), dbh = umma_desc_sw128(b_hi + ko);
                            umma_f16(d_tmem, dah, dbh, idesc, (it | k) != 0);
                            if (passes == 3) {
                                umma_f16(d_tmem, umma_desc_sw128(a_lo + ko), dbh, idesc, 1);
                                umma_f16(d_tmem, dah, umma_desc_sw128(b_lo + ko), idesc, 1);
                            }
                        }
                        umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                    umma_commit(&tfull_bar[acc]);        // chunk accumulator complete -> epilogue warps
                    if (++acc == NACC) { acc = 0; acc_phase ^= 1; }
                }
            }
        }
      }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(TC_REGS_INC));
        // ===================== epilogue (warps 4..11) =====================
        constexpr int HN = BN / 2;              // columns owned by this thread
        const int q = warp & 3;                 // TMEM lane quarter this warp may access
        const int half = (warp - 4) >> 2;       // column half of the tile
        const int row = q * 32 + lane;          // accumulator row == pixel index inside the tile
        const int tx_i = row & (ti.TW - 1);
        const int ty_i = (row >> ti.tw_log2) & (ti.TH - 1);
        const int tn_i = row >> (ti.tw_log2 + ti.th_log2);
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < ti.total; tile += gridDim.x) {
            const int sp = tile % ti.ksplit;
            const int kit0s = sp * ti.kit_per;
            const int kloc = (kit0s + ti.kit_per < kiters ? kit0s + ti.kit_per : kiters) - kit0s;
            const int nchunks = (kloc + chunk_iters - 1) / chunk_iters;
            int m = (tile / ti.ksplit) / ti.nblk;
            const int nb = (tile / ti.ksplit) - m * ti.nblk;
            const int x = (m % ti.tiles_x) * ti.TW + tx_i;
            m /= ti.tiles_x;
            const int y = (m % ti.tiles_y) * ti.TH + ty_i;
            const int n = (m / ti.tiles_y) * ti.TN + tn_i;
            const bool valid = n < g.N && y < g.OH && x < g.OW;
            const long long pix = ((long long)n * g.OH + y) * g.OW + x;

            // one image per tile (every layer from 16x16 up): stage the per-(sample, channel) epilogue vectors in shared
            // memory and fetch the pixel's noise now, so that the tile's final epilogue needs no dependent L2 round trips
            // (measured: worth 5-13 % for BN = 64 / 128, whose tiles are short; BN = 256 tiles are long enough to hide the
            // global-load epilogue and lose 4 % to the two barriers, so they keep it)
            const bool staged = BN < 256 && ti.TN == 1 && g.mode == 0;
            float nz = 0.f;
            if (staged) {
                asm volatile("bar.sync 1, %0;" ::"n"(TC_EPI_THREADS) : "memory");
                stage_epilogue_vectors<BN, TC_EPI_THREADS>(epi, stg, n, g.Co, nb * BN, (int)threadIdx.x - (TC_THREADS - TC_EPI_THREADS));
                asm volatile("bar.sync 1, %0;" ::"n"(TC_EPI_THREADS) : "memory");
                if (epi.noise && valid) nz = __ldg(epi.noise + (long long)n * epi.noise_sn + (long long)y * g.OW + x) * __ldg(epi.noise_strength);
            }

            float accv[HN];
            for (int c = 0; c < nchunks; ++c) {
                // weight of this chunk in the register-level sum: 1 (fmaf(v, 1, acc) == acc + v exactly), or the per-pixel
                // blend weight of the chunk's tap (ConvGeom::chunk_scale)
                // times the compensation of the truncating accumulate for this chunk's chain of MMAs (ConvGeom::acc_comp)
                const int n_it = kloc - c * chunk_iters < chunk_iters ? kloc - c * chunk_iters : chunk_iters;
                const float csc = ((g.chunk_scale && valid) ? __ldg(g.chunk_scale + (long long)c * g.OH * g.OW + (long long)y * g.OW + x) : 1.f) *
                                  (1.f + g.acc_comp * (float)(n_it * (passes == 3 ? 12 : 4)));
                mbar_wait(&tfull_bar[acc], acc_phase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + half * HN);
#pragma unroll
                for (int p = 0; p < HN / 16; ++p) {
                    float v[16];
                    tmem_ld16(taddr + p * 16, v);
#pragma unroll
                    for (int i = 0; i < 16; ++i) accv[p * 16 + i] = c == 0 ? v[i] * csc : fmaf(v[i], csc, accv[p * 16 + i]);
                }
                tc_fence_before();
                mbar_arrive(&tempty_bar[acc]);
                if (++acc == NACC) { acc = 0; acc_phase ^= 1; }
            }
            if (valid) {
                float rgb[3] = {0.f, 0.f, 0.f};
#pragma unroll
                for (int p = 0; p < HN / 16; ++p) {
                    const int oi = half * HN + p * 16;
                    const int o0 = nb * BN + oi;
                    if (g.mode == 1) raw_store<16>(g, accv + p * 16, n, y, x, o0, (long long)sp * ti.split_stride);
                    else if (staged) epilogue_apply_staged<BN, 16>(epi, stg, accv + p * 16, nz, g.Co, o0, oi, rgb, pix);
                    else epilogue_apply<16>(epi, accv + p * 16, n, y, x, g.OH, g.OW, g.Co, o0, rgb, pix);
                    if ((p & 1) && g.mode == 0 && epi.rgb_w) {   // one torgb partial per CONV_RGB_BLOCK = 32 channels
                        float* dst = epi.rgb_out + (pix * (g.Co / CONV_RGB_BLOCK) + (o0 - 16) / CONV_RGB_BLOCK) * 4;
                        *reinterpret_cast<float4*>(dst) = make_float4(rgb[0], rgb[1], rgb[2], 0.f);
                        rgb[0] = rgb[1] = rgb[2] = 0.f;
                    }
                }
            }
        }
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
static int encode_map(CUtensorMap* map, const void* ptr, int rank, const uint64_t* dims, const uint32_t* box) {
    return encode_tmap(map, ptr, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, rank, dims, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

static int pow2_ceil(int v) { int p = 1; while (p < v) p <<= 1; return p; }
static int ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

template <int BN>
static int launch_bn(const ConvTmaps& maps, const ConvGeom& g, const EpiParams& epi, const TileInfo& ti, int passes,
                     cudaStream_t stream) {
    using Cfg = TcCfg<BN>;
    static DeviceInit once;
    int num_sms = 0;
    if (int e = device_init(once, &num_sms, []() -> int {
            SHGAN_CUDA(cudaFuncSetAttribute(conv_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
            return 0;
        })) return e;
    const int grid = ti.total < num_sms ? ti.total : num_sms;
    // two-level accumulation: at most TC_MAX_CHUNK_ITERS (tap, slab) steps are chained inside one TMEM accumulator
    const int kiters = ti.ksplit > 1 ? ti.kit_per : g.ntaps * (g.C / TC_KC);
    const int nchunks = ceil_div(kiters, TC_MAX_CHUNK_ITERS);
    int chunk_iters = ceil_div(kiters, nchunks);
    if (g.chunk_scale) {      // one chunk per tap (taps are the outer loop of the K order)
        chunk_iters = g.C / TC_KC;
        SHGAN_CHECK(chunk_iters <= TC_MAX_CHUNK_ITERS, "chunk_scale needs C <= 256");
    }
    conv_tc_kernel<BN><<<grid, TC_THREADS, Cfg::SMEM_BYTES, stream>>>(maps, g, epi, ti, passes, chunk_iters);
    SHGAN_LAUNCH_CHECK();
    return 0;
}

static int launch_conv_tc_impl(const ConvGeom& g, const EpiParams& epi, int block_n, int passes, cudaStream_t stream, int ksplit,
                               long long split_stride) {
    SHGAN_CHECK(g.C % TC_KC == 0, "C must be a multiple of 64 for the tensor-core path");
    SHGAN_CHECK(passes == 1 || passes == 3, "passes must be 1 or 3");
    TileInfo ti;
    ti.TW = pow2_ceil(g.OW) < 16 ? pow2_ceil(g.OW) : 16;
    const int th_max = TC_M / ti.TW;
    ti.TH = pow2_ceil(g.OH) < th_max ? pow2_ceil(g.OH) : th_max;
    ti.TN = TC_M / (ti.TW * ti.TH);
    ti.tw_log2 = ilog2(ti.TW);
    ti.th_log2 = ilog2(ti.TH);
    ti.tiles_x = ceil_div(g.OW, ti.TW);
    ti.tiles_y = ceil_div(g.OH, ti.TH);
    ti.tiles_n = ceil_div(g.N, ti.TN);
    if (block_n == 0) {
        // tile width by a two-term cost model: waves of tiles over the SMs x clocks per MMA.  One thread issues an MMA every
        // ~90 clocks at best (profiles/r1_conv_findings.md section 2), so N = 64 and N = 128 instructions cost the same and
        // only N = 256 is bound by the tensor pipe (128 clk): the widest tile wins on large layers (least operand re-fetch,
        // fewest instructions), and on the 4x4 .. 16x16 layers the width that needs the fewest waves does.
        int dev = 0, num_sms = 148;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        const long long mt = (long long)ti.tiles_x * ti.tiles_y * ti.tiles_n;
        long long best_cost = -1;
        for (int bn = conv_block_n(g.Co, 0); bn >= 64; bn >>= 1) {
            if (g.Co % bn) continue;
            const long long waves = (mt * (g.Co / bn) + num_sms - 1) / num_sms;
            const long long cost = waves * (bn / 2 > 90 ? bn / 2 : 90);
            if (best_cost < 0 || cost <= best_cost) { best_cost = cost; block_n = bn; }   // ties: more, narrower tiles
        }
    }
    SHGAN_CHECK(g.Co % block_n == 0, "Co must be a multiple of block_n");
    ti.nblk = g.Co / block_n;
    ti.ksplit = ksplit;
    ti.kit_per = ceil_div(g.ntaps * (g.C / TC_KC), ksplit);
    ti.split_stride = split_stride;
    const long long total = (long long)ti.tiles_x * ti.tiles_y * ti.tiles_n * ti.nblk * ksplit;
    SHGAN_CHECK(total <= INT32_MAX, "too many tiles");
    ti.total = (int)total;

    ConvTmaps maps;
    const uint32_t abox[4] = {(uint32_t)TC_KC, (uint32_t)ti.TW, (uint32_t)ti.TH, (uint32_t)ti.TN};
    for (int s = 0; s < g.num_src; ++s) {
        const uint64_t dims[4] = {(uint64_t)g.C, (uint64_t)g.src_w[s], (uint64_t)g.src_h[s], (uint64_t)g.N};
        if (int e = encode_map(&maps.a_hi[s], g.src_hi[s], 4, dims, abox)) return e;
        if (int e = encode_map(&maps.a_lo[s], g.src_lo[s], 4, dims, abox)) return e;
    }
    int w_taps = 0;
    for (int t = 0; t < g.ntaps; ++t) w_taps = g.tap_w[t] + 1 > w_taps ? g.tap_w[t] + 1 : w_taps;
    // the weight map covers exactly the taps this launch references (tap_w < w_taps was validated by shgan_conv_igemm)
    const uint64_t wdims[2] = {(uint64_t)g.C, (uint64_t)w_taps * g.Co};
    const uint32_t wbox[2] = {(uint32_t)TC_KC, (uint32_t)block_n};
    if (int e = encode_map(&maps.w_hi, g.w_hi, 2, wdims, wbox)) return e;
    if (int e = encode_map(&maps.w_lo, g.w_lo, 2, wdims, wbox)) return e;

    if (block_n == 64) return launch_bn<64>(maps, g, epi, ti, passes, stream);
    if (block_n == 128) return launch_bn<128>(maps, g, epi, ti, passes, stream);
    return launch_bn<256>(maps, g, epi, ti, passes, stream);
}

int launch_conv_tc(const ConvGeom& g, const EpiParams& epi, int block_n, int passes, cudaStream_t stream) {
    return launch_conv_tc_impl(g, epi, block_n, passes, stream, 1, 0);
}

// ---- split-K for the 4x4 / 8x8 layers ---------------------------------------------------------------------------------------
// Those layers have 2-8 pixel tiles: 16-64 CTAs each chaining all 72 (tap, slab) steps (864 MMAs at the single-thread issue rate,
// 78 us measured for 3.6 / 14.5 GFLOP).  Here `ksplit` CTAs share an output tile, each runs a slice of the K loop and writes its
// fp32 partial tile into its own buffer (RAW mode); one small kernel sums the partials and applies the fused epilogue.  The
// summation order is fixed (split 0, 1, 2, ...): results are deterministic.
// thread = (pixel, 8 channels): 2 independent 128-bit loads per split; the torgb partial of a 32-channel block is summed over
// the 4 adjacent lanes that hold it
__global__ void __launch_bounds__(256)
conv_splitk_epilogue_kernel(const float* __restrict__ z, int ksplit, long long split_stride, const EpiParams epi, int N, int OH, int OW,
                            int Co) {
    const int cgs = Co / 8;
    const long long total = (long long)N * OH * OW * cgs;
    for (long long gid0 = blockIdx.x * (long long)blockDim.x; gid0 < total; gid0 += (long long)gridDim.x * blockDim.x) {
        const long long gid = gid0 + threadIdx.x;
        const bool live = gid < total;                       // (total is a multiple of 4: the shuffles below stay inside a 32-channel block)
        const int cg = (int)((live ? gid : 0) % cgs);
        const long long pix = (live ? gid : 0) / cgs;
        const int x = (int)(pix % OW), y = (int)((pix / OW) % OH), n = (int)(pix / ((long long)OW * OH));
        const int o0 = cg * 8;
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.f;
        if (live) {
            const float4* src = reinterpret_cast<const float4*>(z + pix * Co + o0);
            const long long st4 = split_stride / 4;
#pragma unroll 6
            for (int sp = 0; sp < ksplit; ++sp) {
                const float4 a = __ldg(src + sp * st4), b = __ldg(src + sp * st4 + 1);
                v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w;
                v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
            }
        }
        float rgb[3] = {0.f, 0.f, 0.f};
        if (live) epilogue_apply<8>(epi, v, n, y, x, OH, OW, Co, o0, rgb, pix);
        if (epi.rgb_w) {
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                rgb[j] += __shfl_xor_sync(0xffffffffu, rgb[j], 1);
                rgb[j] += __shfl_xor_sync(0xffffffffu, rgb[j], 2);
            }
            if (live && (cg & 3) == 0)
                *reinterpret_cast<float4*>(epi.rgb_out + (pix * (Co / CONV_RGB_BLOCK) + cg / 4) * 4) = make_float4(rgb[0], rgb[1], rgb[2], 0.f);
        }
    }
}

int launch_conv_tc_splitk(const ConvGeom& g0, const EpiParams& epi, int passes, float* scratch, int max_splits, cudaStream_t stream) {
    if (g0.C % TC_KC != 0 || g0.Co % 128 != 0 || g0.mode != 0 || g0.chunk_scale) return -1;
    int dev = 0, num_sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    // tiles of the un-split launch at BN = 128 (the MMA issue rate of one thread is the same for N = 64 and N = 128).  Layers whose
    // BN = 128 tiles already fill the SMs (16x16 at batch 16) are left alone: splitting them at BN = 256 measured 53 + 9 us
    // against 55 us un-split (two-stage operand pipeline at that width).
    const int TW = pow2_ceil(g0.OW) < 16 ? pow2_ceil(g0.OW) : 16;
    const int th_max = TC_M / TW;
    const int TH = pow2_ceil(g0.OH) < th_max ? pow2_ceil(g0.OH) : th_max;
    const int TN = TC_M / (TW * TH);
    const long long tiles = (long long)ceil_div(g0.OW, TW) * ceil_div(g0.OH, TH) * ceil_div(g0.N, TN) * (g0.Co / 128);
    const int kiters = g0.ntaps * (g0.C / TC_KC);
    const int bn = 128;
    int ksplit = (int)(num_sms / tiles);
    if (ksplit > max_splits) ksplit = max_splits;
    if (ksplit > kiters / 2) ksplit = kiters / 2;          // at least two K steps per CTA
    if (ksplit < 2) return -1;                             // enough tiles already: not worth a second launch
    ksplit = ceil_div(kiters, ceil_div(kiters, ksplit));   // no empty split
    ConvGeom g = g0;
    g.mode = 1;
    g.z = scratch;
    g.ZH = g0.OH; g.ZW = g0.OW; g.zsy = 1; g.zsx = 1; g.zoy = 0; g.zox = 0;
    const long long stride = (long long)g0.N * g0.OH * g0.OW * g0.Co;
    if (int e = launch_conv_tc_impl(g, EpiParams{}, bn, passes, stream, ksplit, stride)) return e;
    const long long items = (long long)g0.N * g0.OH * g0.OW * (g0.Co / 8);
    long long blocks = ceil_div64(items, 256);
    if (blocks > 148LL * 8) blocks = 148LL * 8;
    conv_splitk_epilogue_kernel<<<(unsigned)blocks, 256, 0, stream>>>(scratch, ksplit, stride, epi, g0.N, g0.OH, g0.OW, g0.Co);
    SHGAN_LAUNCH_CHECK();
    return 0;
}

}  // namespace shgan
