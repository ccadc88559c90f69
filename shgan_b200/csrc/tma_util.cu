// Host-side tensor-map encoding through the driver entry point (no link-time dependency on libcuda).
#include "tma_util.cuh"

namespace shgan {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// dims/box: fastest dimension first; strides are the dense strides of a contiguous tensor; out-of-bounds elements read as 0
int encode_tmap(CUtensorMap* map, const void* ptr, CUtensorMapDataType dtype, int elem_bytes, int rank, const uint64_t* dims,
                const uint32_t* box, CUtensorMapSwizzle swizzle) {
    EncodeTiledFn fn = get_encode_fn();
    SHGAN_CHECK(fn, "cuTensorMapEncodeTiled is not available from the CUDA driver");
    cuuint64_t gdim[5], gstr[4];
    cuuint32_t bdim[5], estr[5];
    uint64_t stride = (uint64_t)elem_bytes;
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bdim[i] = box[i];
        estr[i] = 1;
        stride *= dims[i];
        if (i < rank - 1) gstr[i] = stride;
    }
    CUresult r = fn(map, dtype, rank, const_cast<void*>(ptr), gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SHGAN_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
    return 0;
}

}  // namespace shgan
