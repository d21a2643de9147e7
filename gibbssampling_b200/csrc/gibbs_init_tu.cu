// gibbssampling_b200/csrc/gibbs_init_tu.cu -- the grid-wide random-start kernels, one kind per translation unit.
// Compiled four times by _build.py with -DGIBBS_INIT_TU_KIND=0 (init_kernel, fixed background), 1 (init_kernel with the
// data-derived background), 2 (init_smem_kernel) or 3 (init_tiled_kernel), each time for the 16 k-widths. Like the chain kernels
// (gibbs_chain_tu.cu) these are register-limited; inside the big API module under nvcc --split-compile their code
// generation (spills or none) changed from build to build with the kernels that shared their compiler partition.
#include "gibbs_kernels.cuh"

#if !defined(GIBBS_INIT_TU_KIND)
#error "compile with -DGIBBS_INIT_TU_KIND=0|1|2|3 (see _build.py)"
#endif

namespace gibbs {

#if GIBBS_INIT_TU_KIND == 3
template <int KPV>
static cudaError_t launch_tiled_one(const ChainArgs &a, int grid, int smem, cudaStream_t stream, int tile_rows) {
    auto kernel = init_tiled_kernel<KPV>;
    if (smem > 48 * 1024) {
        const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
    }
    kernel<<<grid, TILED_WARPS * 32, smem, stream>>>(a, tile_rows, philox_round_keys(a.seed));
    return cudaGetLastError();
}

// smem = init_tiled_total_bytes(tile_rows, row_words, KP)
cudaError_t launch_init_tiled(const ChainArgs &a, int grid, int smem, cudaStream_t stream, int tile_rows) {
    switch ((a.k + 1) / 2) {
#define X(KPV) case KPV: return launch_tiled_one<KPV>(a, grid, smem, stream, tile_rows);
        X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15) X(16)
#undef X
    default: return cudaErrorInvalidValue;
    }
}
#else
template <int KPV>
static cudaError_t launch_one(const ChainArgs &a, int grid, int smem, cudaStream_t stream) {
#if GIBBS_INIT_TU_KIND == 2
    auto kernel = init_smem_kernel<KPV>;
    constexpr int threads = ISM_WARPS * 32;
#else
    auto kernel = init_kernel<KPV, GIBBS_INIT_TU_KIND == 1>;
    constexpr int threads = INIT_WARPS * 32;
#endif
    if (smem > 48 * 1024) {
        const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
    }
    kernel<<<grid, threads, smem, stream>>>(a);
    return cudaGetLastError();
}

#if GIBBS_INIT_TU_KIND == 0
#define GIBBS_INIT_TU_NAME launch_init_wide
#elif GIBBS_INIT_TU_KIND == 1
#define GIBBS_INIT_TU_NAME launch_init_wide_drift
#else
#define GIBBS_INIT_TU_NAME launch_init_smem
#endif

cudaError_t GIBBS_INIT_TU_NAME(const ChainArgs &a, int grid, int smem, cudaStream_t stream) {
    switch ((a.k + 1) / 2) {
#define X(KPV) case KPV: return launch_one<KPV>(a, grid, smem, stream);
        X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15) X(16)
#undef X
    default: return cudaErrorInvalidValue;
    }
}
#endif

} // namespace gibbs
