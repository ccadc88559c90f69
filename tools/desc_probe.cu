// Hardware probe (development aid, not product code): does a SWIZZLE_128B K-major UMMA shared-memory descriptor accept a
// start address that is a multiple of 128 B but not of 1024 B (i.e. a matrix that starts in the middle of a swizzle atom)?
// The halo-tile convolution relies on it: tap (dy,dx) of a TMA-staged halo tile is the same smem tile read from a start
// address shifted by (dy*P + dx) pixels of 128 B.  mode 0: base_offset field = 0; mode 1: base_offset = (addr >> 7) & 7.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I shgan_b200/csrc -I include tools/desc_probe.cu \
//        shgan_b200/csrc/tma_util.cu -o gpurun_out/desc_probe
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <string>
#include <vector>
#include "tma_util.cuh"

namespace shgan { thread_local std::string g_err; void set_error(const std::string& s) { g_err = s; } }
using namespace shgan;

constexpr int ROWS = 320, KC = 64, BN = 64;

__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr, int mode) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    if (mode == 1) d |= (uint64_t)((saddr >> 7) & 7) << 49;
    d |= (uint64_t)2 << 61;
    return d;
}

__global__ void probe(const __grid_constant__ CUtensorMap ma, const __grid_constant__ CUtensorMap mb, float* out, int shift, int mode) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sa = smem;                       // ROWS x 128 B
    uint8_t* sb = smem + 512 * 128;           // 64 x 128 B (A is staged by two 256-row boxes; rows >= ROWS are zero-filled)
    uint64_t* bar = (uint64_t*)(sb + BN * 128);
    uint32_t* slot = (uint32_t*)(bar + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_fence_init(); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar[0], 512 * 128 + BN * 128);
        tma_load_2d(sa, &ma, &bar[0], 0, 0);
        tma_load_2d(sa + 256 * 128, &ma, &bar[0], 0, 256);
        tma_load_2d(sb, &mb, &bar[0], 0, 0);
        mbar_wait(&bar[0], 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        for (int k = 0; k < 4; ++k) {
            const uint64_t da = desc_sw128(smem_u32(sa) + shift * 128 + k * 32, mode);
            const uint64_t db = desc_sw128(smem_u32(sb) + k * 32, 0);
            const uint32_t accum = k != 0;
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                         ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(accum) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[1])) : "memory");
    }
    mbar_wait(&bar[1], 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c = 0; c < BN; c += 16) {
        uint32_t r[16];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                       "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 16; ++i) out[(warp * 32 + lane) * BN + c + i] = __uint_as_float(r[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
}

int main() {
    std::vector<__half> a(ROWS * KC), b(BN * KC);
    std::vector<float> af(ROWS * KC), bf(BN * KC);
    srand(1);
    for (size_t i = 0; i < a.size(); ++i) { a[i] = __float2half((rand() % 2001 - 1000) / 1000.f); af[i] = __half2float(a[i]); }
    for (size_t i = 0; i < b.size(); ++i) { b[i] = __float2half((rand() % 2001 - 1000) / 1000.f); bf[i] = __half2float(b[i]); }
    __half *da, *db; float* dout;
    cudaMalloc(&da, a.size() * 2); cudaMalloc(&db, b.size() * 2); cudaMalloc(&dout, 128 * BN * 4);
    cudaMemcpy(da, a.data(), a.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(db, b.data(), b.size() * 2, cudaMemcpyHostToDevice);
    CUtensorMap ma, mb;
    const uint64_t adims[2] = {KC, ROWS}, bdims[2] = {KC, BN};
    const uint32_t abox[2] = {KC, 256};
    if (encode_tmap(&ma, da, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, 2, adims, abox, CU_TENSOR_MAP_SWIZZLE_128B)) { printf("encode failed: %s\n", g_err.c_str()); return 1; }
    const uint32_t bbox[2] = {KC, BN};
    if (encode_tmap(&mb, db, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, 2, bdims, bbox, CU_TENSOR_MAP_SWIZZLE_128B)) { printf("encode failed: %s\n", g_err.c_str()); return 1; }
    const int smem_bytes = 1024 + 512 * 128 + BN * 128 + 64;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    const int shifts[] = {0, 1, 2, 3, 5, 7, 8, 9, 17, 18, 19, 37, 66, 130, 191};
    std::vector<float> out(128 * BN);
    for (int mode = 0; mode < 2; ++mode)
        for (int s : shifts) {
            cudaMemset(dout, 0, out.size() * 4);
            probe<<<1, 128, smem_bytes>>>(ma, mb, dout, s, mode);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("mode %d shift %d: CUDA error %s\n", mode, s, cudaGetErrorString(e)); return 1; }
            cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
            double maxerr = 0;
            for (int m = 0; m < 128; ++m)
                for (int n = 0; n < BN; ++n) {
                    double ref = 0;
                    for (int k = 0; k < KC; ++k) ref += (double)af[(s + m) * KC + k] * bf[n * KC + k];
                    maxerr = fmax(maxerr, fabs(ref - out[m * BN + n]));
                }
            printf("mode %d shift %3d: max err %.3e %s\n", mode, s, maxerr, maxerr < 1e-3 ? "OK" : "MISMATCH");
        }
    return 0;
}
