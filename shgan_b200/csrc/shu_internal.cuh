// Internal interface between the pieces of the Spectral Hint Unit (shu.cu: entry points, generic-size transforms;
// shu_fft64.cu: register-resident radix-8 transforms for input_res 64; shu_mix_tc.cu: tcgen05 channel mix).
#pragma once
#include "common.cuh"

namespace shgan {

struct ShuBands {
    float* out[8];
    int gauss_off[8];   // float offset of band k's mask inside `gauss`
    int num_bands, lowest_log2;
};

// Packed operands of the tensor-core channel mix (C == 32), the exact shared-memory image the kernel bulk-copies in:
// K-major SWIZZLE_128B fp16 tiles (row = output channel, 128 B = 64 input channels, 16-byte chunk j stored at j ^ (row & 7))
//   [conv0 hi 8 KB][conv0 lo 8 KB]  then for the anchor pair p = 0..2 (anchors k = p and k = 3 + p, i.e. the two height
//   nodes of width node p of the [2,3] heterogeneous filter): [pair hi 16 KB][pair lo 16 KB], rows 0..63 = W1_p, 64..127 = W1_{3+p}
constexpr int SHU_PACKED_BYTES = 16384 + 3 * 32768;

// spectra between the three kernels: [N, 2C, bins] fp32.  Generic-size transforms: bin = s * Rh + kx (row-shifted spectrum,
// row-major); shu_fft64.cu: bin = kx * R + s (kx-major).  cw_binorder: the blend weights [6, bins] in that same bin order.
int launch_shu_mix_tc(const float* spec1, const void* packed, const float* conv0_b, const float* cw_binorder, float* spec2, int N, int R,
                      float scale, cudaStream_t stream, void* trace = nullptr);

// input_res == 64, C == 32, lowest_res >= 4: spectra in the transposed layout
// cw [6, 64, 33] -> cw_kxmajor [6, 33 * 64] is written by block 0 of the forward launch (consumed by the mix launch that follows)
int launch_shu_rfft2_r64(const float* x, float* spec1, const float* cw, float* cw_kxmajor, int N, int C, cudaStream_t stream);
int launch_shu_irfft2_r64(const float* spec2, const float* gauss, const ShuBands& bands, int N, int C, cudaStream_t stream);

// transforms of 4 ... 32 points with one thread per transform (shu_small.cu): the forward pass for input_res <= 32 and the
// inverse of one band of r <= 32 (spectra in the generic row-major layout of shu.cu; planes and rows 16-byte aligned)
int launch_shu_rfft2_small(const float* x, float* spec1, int N, int C, int R, cudaStream_t stream);
int launch_shu_irfft2_small(const float* spec2, const float* gm, float* out, int N, int C, int R, int r, cudaStream_t stream);

}  // namespace shgan
