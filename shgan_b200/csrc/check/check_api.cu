// TEST-ONLY library (libshgan_b200_check.so): the fp32 FMA cross-check implementation of shgan_conv_igemm
// (check/conv_simt.cu) behind its own C entry point.  It is built next to the product library but is NOT part of it:
// libshgan_b200.so contains no CUDA-core convolution, and `impl = 1` in a shgan_conv_desc is rejected there.  The GPU
// tests load this library to compare the tcgen05 kernels with plain fp32 FMA arithmetic on identical operands at full
// layer sizes (where the CPU oracle would take minutes).
#include "../conv_common.cuh"

using namespace shgan;

extern "C" int shgan_check_conv_igemm(const shgan_conv_desc* d, void* stream) {
    SHGAN_CHECK(d, "null descriptor");
    SHGAN_CHECK(d->num_src >= 1 && d->num_src <= SHGAN_MAX_SRC && d->ntaps >= 1 && d->ntaps <= SHGAN_MAX_TAPS, "bad descriptor");
    SHGAN_CHECK(d->C >= 64 && d->C % 64 == 0 && d->Co >= 64 && d->Co % 64 == 0, "C and Co must be multiples of 64");
    SHGAN_CHECK(d->mode == 0 || d->mode == 1, "mode must be 0 (ACT) or 1 (RAW)");
    if (d->mode == 0) {
        if (const char* m = check_epi(d->epi, d->Co)) SHGAN_CHECK(false, m);
    }
    if (d->N == 0) return 0;
    const ConvGeom g = make_geom(*d);
    const EpiParams epi = d->mode == 0 ? make_epi(d->epi) : EpiParams{};
    return launch_conv_simt(g, epi, d->block_n, (cudaStream_t)stream);
}
