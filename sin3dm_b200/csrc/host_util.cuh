// Host-side helpers shared by the translation units of libsin3dm_b200: error reporting, kernel launch, TMA descriptors.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <string>
#include <utility>

// ------------------------------------------------------------------------------------ errors
inline thread_local std::string g_err;     // one per thread for the whole library (s3d_last_error)
inline int fail(const std::string& m) {
    g_err = m;
    return 1;
}
struct S3dError {
    std::string msg;
};
#define S3D_CHECK(cond, msg)                                                      \
    do {                                                                          \
        if (!(cond)) throw S3dError{std::string(msg) + " (" #cond ")"};          \
    } while (0)
#define CUDA_TRY(expr)                                                                                    \
    do {                                                                                                  \
        cudaError_t _e = (expr);                                                                          \
        if (_e != cudaSuccess)                                                                            \
            throw S3dError{std::string(#expr) + ": " + cudaGetErrorName(_e) + ": " + cudaGetErrorString(_e)}; \
    } while (0)
#define LAUNCH_CHECK(name)                                                                 \
    do {                                                                                   \
        cudaError_t _e = cudaGetLastError();                                               \
        if (_e != cudaSuccess) throw S3dError{std::string("launch ") + name + ": " + cudaGetErrorString(_e)}; \
    } while (0)

// ------------------------------------------------------------------------------------ launches
// S3D_PDL=1 sends every kernel out with programmatic stream serialization (PDL): see pdl_wait()/pdl_trigger().
inline bool g_pdl = false;     // measured on B200 (r1): PDL on every launch is ~2.5% slower than plain graph edges at batch 1
template <typename... KArgs, typename... Args>
static void launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_pdl ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(std::forward<Args>(args))...);
    if (e != cudaSuccess) throw S3dError{std::string("cudaLaunchKernelEx: ") + cudaGetErrorString(e)};
}
// For kernels that do NOT execute griddepcontrol.wait (helpers off the sampling loop, the decoder / encoder and the training
// kernels): never launched with programmatic stream serialization, so S3D_PDL=1 cannot let them overtake their producer.
template <typename... KArgs, typename... Args>
static void launch_plain(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(std::forward<Args>(args))...);
    if (e != cudaSuccess) throw S3dError{std::string("cudaLaunchKernelEx: ") + cudaGetErrorString(e)};
}

// ------------------------------------------------------------------------------------ driver entry (TMA descriptors)
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
        S3D_CHECK(p != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available");
        fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}
// fp16 tensor, dims innermost-first, SWIZZLE_128B, zero OOB fill
inline void make_tmap(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint32_t* box) {
    cuuint64_t gdim[5], gstride[4];
    cuuint32_t bx[5], es[5];
    uint64_t stride = 2;
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bx[i] = box[i];
        es[i] = 1;
        stride *= dims[i];
        if (i < rank - 1) gstride[i] = stride;
    }
    CUresult r = get_encode()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, const_cast<void*>(base), gdim, gstride, bx, es,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw S3dError{"cuTensorMapEncodeTiled failed with CUresult " + std::to_string(static_cast<int>(r))};
}

// dense fp32 tensor, dims innermost-first, no swizzle, zero OOB fill (row bytes of the box must be a multiple of 16)
inline void make_tmap_f32(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint32_t* box) {
    cuuint64_t gdim[5], gstride[4];
    cuuint32_t bx[5], es[5];
    uint64_t stride = 4;
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bx[i] = box[i];
        es[i] = 1;
        stride *= dims[i];
        if (i < rank - 1) gstride[i] = stride;
    }
    CUresult r = get_encode()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, const_cast<void*>(base), gdim, gstride, bx, es,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw S3dError{"cuTensorMapEncodeTiled (fp32) failed with CUresult " + std::to_string(static_cast<int>(r))};
}

// same, with explicit byte strides for dims 1..rank-1
inline void make_tmap_strided(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                              const uint32_t* box) {
    cuuint64_t gdim[5], gstride[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bx[i] = box[i];
        es[i] = 1;
        if (i < rank - 1) gstride[i] = strides_bytes[i];
    }
    CUresult r = get_encode()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, const_cast<void*>(base), gdim, gstride, bx, es,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw S3dError{"cuTensorMapEncodeTiled failed with CUresult " + std::to_string(static_cast<int>(r))};
}

