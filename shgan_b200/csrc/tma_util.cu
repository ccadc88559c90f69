// Host-side tensor-map encoding through the driver entry point (no link-time dependency on libcuda).
#include "tma_util.cuh"

#include <cstring>
#include <mutex>
#include <unordered_map>

namespace shgan {

// A tensor map depends only on (pointer, element type, rank, dims, box, swizzle): the engine reuses its activation buffers
// and packed weights from step to step, so the encoded descriptors are cached instead of being re-encoded by the driver on
// every launch (about six maps per convolution launch; matters for eager / batch-1 latency, a replayed CUDA graph never
// re-encodes).  Bounded: the cache is dropped when it reaches TMAP_CACHE_MAX entries.
struct TmapKey {
    uint64_t ptr, dims[5];
    uint32_t box[5], dtype, elem_bytes, rank, swizzle;
    bool operator==(const TmapKey& o) const { return std::memcmp(this, &o, sizeof(TmapKey)) == 0; }
};
struct TmapKeyHash {
    size_t operator()(const TmapKey& k) const {
        const uint64_t* w = reinterpret_cast<const uint64_t*>(&k);
        uint64_t h = 1469598103934665603ull;
        for (size_t i = 0; i < sizeof(TmapKey) / 8; ++i) h = (h ^ w[i]) * 1099511628211ull;
        return (size_t)h;
    }
};
static_assert(sizeof(TmapKey) % 8 == 0, "TmapKey is hashed as 64-bit words");
constexpr size_t TMAP_CACHE_MAX = 4096;
static std::mutex g_tmap_mu;
static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmap_cache;

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
    TmapKey key;
    std::memset(&key, 0, sizeof(key));
    key.ptr = (uint64_t)(uintptr_t)ptr; key.dtype = (uint32_t)dtype; key.elem_bytes = (uint32_t)elem_bytes;
    key.rank = (uint32_t)rank; key.swizzle = (uint32_t)swizzle;
    for (int i = 0; i < rank; ++i) { key.dims[i] = dims[i]; key.box[i] = box[i]; }
    {
        std::lock_guard<std::mutex> lock(g_tmap_mu);
        auto it = g_tmap_cache.find(key);
        if (it != g_tmap_cache.end()) {
            *map = it->second;
            return 0;
        }
    }
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
    {
        std::lock_guard<std::mutex> lock(g_tmap_mu);
        if (g_tmap_cache.size() >= TMAP_CACHE_MAX) g_tmap_cache.clear();
        g_tmap_cache.emplace(key, *map);
    }
    return 0;
}

}  // namespace shgan
