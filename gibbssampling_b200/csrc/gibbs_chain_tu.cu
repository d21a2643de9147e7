// gibbssampling_b200/csrc/gibbs_chain_tu.cu -- one group of chain_kernel instantiations per translation unit.
// Compiled several times by _build.py with -DGIBBS_TU_NAME=launch_chain_xxx -DGIBBS_TU_T=<warps per chain>
// -DGIBBS_TU_MASKED=0|1 -DGIBBS_TU_DRIFT=0|1 [-DGIBBS_TU_INIT_ONLY=1], each time for the 16 k-widths.
//
// Why not one module: the 4-warp fixed-background kernel sits at its register limit (72), and with every
// instantiation in one module under nvcc --split-compile its code generation (spills or none) changed with the
// unrelated kernels that happened to share its compiler partition. Small modules compiled without --split-compile
// are reproducible, build in parallel, and give the best allocation measured (k = 12: no spill; k = 20: 24 B).
#include "gibbs_kernels.cuh"

#ifndef GIBBS_TU_INIT_ONLY
#define GIBBS_TU_INIT_ONLY 0 // 1: the instantiations that run the random starts on the chain's own team (nothing else)
#endif
#if !defined(GIBBS_TU_NAME) || !defined(GIBBS_TU_T) || !defined(GIBBS_TU_MASKED) || !defined(GIBBS_TU_DRIFT)
#error "compile with -DGIBBS_TU_NAME=... -DGIBBS_TU_T=... -DGIBBS_TU_MASKED=... -DGIBBS_TU_DRIFT=... (see _build.py)"
#endif

namespace gibbs {

template <int KPV>
static cudaError_t launch_one(const ChainArgs &a, int grid, int smem, cudaStream_t stream) {
    auto kernel = chain_kernel<KPV, GIBBS_TU_T, GIBBS_TU_MASKED != 0, GIBBS_TU_DRIFT != 0, GIBBS_TU_INIT_ONLY != 0>;
    if (smem > 48 * 1024) {
        const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
    }
    kernel<<<grid, 32 * GIBBS_TU_T, smem, stream>>>(a);
    return cudaGetLastError();
}

// smem = team_smem_bytes(row_words, GIBBS_TU_T)
cudaError_t GIBBS_TU_NAME(const ChainArgs &a, int grid, int smem, cudaStream_t stream) {
    switch ((a.k + 1) / 2) {
#define X(KPV) case KPV: return launch_one<KPV>(a, grid, smem, stream);
        X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15) X(16)
#undef X
    default: return cudaErrorInvalidValue;
    }
}

} // namespace gibbs
