// shgan_conv_igemm: argument validation and dispatch to the tensor-core kernel (conv_tc.cu).
#include "conv_common.cuh"

using namespace shgan;

extern "C" int shgan_conv_num_nblocks(int Co, int block_n) {
    (void)block_n;
    return Co % CONV_RGB_BLOCK == 0 ? Co / CONV_RGB_BLOCK : 0;
}

extern "C" int shgan_conv_igemm(const shgan_conv_desc* d, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SHGAN_CHECK(d, "null descriptor");
    SHGAN_CHECK(d->num_src >= 1 && d->num_src <= SHGAN_MAX_SRC, "num_src out of range");
    SHGAN_CHECK(d->ntaps >= 1 && d->ntaps <= SHGAN_MAX_TAPS, "ntaps out of range");
    SHGAN_CHECK(d->N >= 0 && d->OH >= 1 && d->OW >= 1, "bad output size");
    SHGAN_CHECK(d->C >= 64 && d->C % 64 == 0 && d->Co >= 64 && d->Co % 64 == 0, "C and Co must be multiples of 64");
    SHGAN_CHECK(d->w_hi && d->w_lo && d->w_taps >= 1, "weights missing");
    for (int s = 0; s < d->num_src; ++s) {
        SHGAN_CHECK(d->src_hi[s] && d->src_lo[s], "source planes missing");
        SHGAN_CHECK(d->src_h[s] >= 1 && d->src_w[s] >= 1, "bad source size");
        SHGAN_CHECK((long long)d->N * d->C * d->src_h[s] * d->src_w[s] <= INT32_MAX, "source tensor is too large");
    }
    for (int t = 0; t < d->ntaps; ++t) {
        SHGAN_CHECK(d->tap_src[t] >= 0 && d->tap_src[t] < d->num_src, "tap_src out of range");
        SHGAN_CHECK(d->tap_w[t] >= 0 && d->tap_w[t] < d->w_taps, "tap_w out of range");
    }
    SHGAN_CHECK((long long)d->N * d->Co * d->OH * d->OW <= INT32_MAX, "output tensor is too large");
    SHGAN_CHECK(d->mode == 0 || d->mode == 1, "mode must be 0 (ACT) or 1 (RAW)");
    if (d->mode == 1) {
        SHGAN_CHECK(d->z && d->zsy >= 1 && d->zsx >= 1 && d->zoy >= 0 && d->zox >= 0, "bad RAW output description");
        SHGAN_CHECK((d->OH - 1) * d->zsy + d->zoy < d->ZH && (d->OW - 1) * d->zsx + d->zox < d->ZW, "RAW output out of range");
        SHGAN_CHECK((long long)d->N * d->Co * d->ZH * d->ZW <= INT32_MAX, "RAW output tensor is too large");
    } else {
        if (const char* m = check_epi(d->epi, d->Co)) SHGAN_CHECK(false, m);
    }
    SHGAN_CHECK(d->block_n == 0 || d->block_n == 64 || d->block_n == 128 || d->block_n == 256, "block_n must be 0, 64, 128 or 256");
    SHGAN_CHECK(d->block_n == 0 || d->Co % d->block_n == 0, "Co must be a multiple of block_n");
    const int bn = d->block_n;
    if (d->N == 0) return 0;
    const ConvGeom g = make_geom(*d);
    const EpiParams epi = d->mode == 0 ? make_epi(d->epi) : EpiParams{};
    SHGAN_CHECK(d->impl != 1, "impl 1 (fp32 FMA cross-check) is not in the product library: it lives in the test-only libshgan_b200_check.so");
    SHGAN_CHECK(d->impl == 0 || (d->impl >= 2 && d->impl <= 4), "impl must be 0, 1, 2, 3 or 4");
    const int passes = d->passes == 0 ? 3 : d->passes;
    // ACT mode with a scratch (z = ZH partial buffers of ZW floats): split-K for layers with too few tiles to fill the GPU
    if (d->mode == 0 && d->z && d->impl == 0 && bn == 0 && !conv_prefers_pair(g)) {
        SHGAN_CHECK(d->ZH >= 2 && (long long)d->ZW >= (long long)d->N * d->OH * d->OW * d->Co, "split-K scratch: z must hold ZH >= 2 buffers of ZW >= N*OH*OW*Co floats");
        SHGAN_CHECK(((uintptr_t)d->z & 15) == 0 && (d->ZW & 3) == 0, "split-K scratch must be 16-byte aligned");
        const int rc = launch_conv_tc_splitk(g, epi, passes, d->z, d->ZH, stream);
        if (rc >= 0) return rc;
    }
    if (d->impl == 4 && bn == 0 && conv_pair_supported(g)) return launch_conv_pair(g, epi, passes, stream);
    if (d->impl == 4) return launch_conv_tc(g, epi, bn, passes, stream);
    // per-layer choice of the product path (impl == 0), from per-layer timings of the three kernels on B200
    // (profiles/r1_conv_findings.md section 4): the two-SM kernel wherever Co % 128 == 0 and there are enough tile pairs to
    // occupy the 74 clusters; the halo kernel for the remaining wide transposed-conv passes; the per-tap kernel elsewhere
    if (d->impl == 0 && bn == 0 && conv_prefers_pair(g)) return launch_conv_pair(g, epi, passes, stream);
    if (d->impl == 3 || (d->impl == 0 && bn == 0 && conv_prefers_halo(g))) return launch_conv_halo(g, epi, bn, passes, stream);
    return launch_conv_tc(g, epi, bn, passes, stream);
}
