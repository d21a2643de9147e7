// gibbssampling_b200/csrc/gibbs_motif_tu.cu -- motif_kernel<KP, T, MASKED> for the 16 k-widths, one team size per
// translation unit (-DGIBBS_MOTIF_TU_T=1|4, -DGIBBS_MOTIF_TU_MASKED=0|1, see _build.py): register-limited like the chain
// kernels, so it gets a module of its own (reproducible code generation, see gibbs_chain_tu.cu).
#include "gibbs_motif.cuh"

#if !defined(GIBBS_MOTIF_TU_T)
#error "compile with -DGIBBS_MOTIF_TU_T=1, 4, 8 or 16 (see _build.py)"
#endif
#if !defined(GIBBS_MOTIF_TU_MASKED)
#define GIBBS_MOTIF_TU_MASKED 0
#endif

namespace gibbs {

template <int KPV>
static cudaError_t launch_one(const MotifArgs &m, int grid, int smem, cudaStream_t stream) {
    auto kernel = motif_kernel<KPV, GIBBS_MOTIF_TU_T, (GIBBS_MOTIF_TU_MASKED != 0)>;
    if (smem > 48 * 1024) {
        const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
    }
    kernel<<<grid, 32 * GIBBS_MOTIF_TU_T, smem, stream>>>(m);
    return cudaGetLastError();
}

#if GIBBS_MOTIF_TU_MASKED
#if GIBBS_MOTIF_TU_T == 4
#define GIBBS_MOTIF_TU_NAME launch_motif_masked_t4
#else
#define GIBBS_MOTIF_TU_NAME launch_motif_masked_t1
#endif
#elif GIBBS_MOTIF_TU_T == 16
#define GIBBS_MOTIF_TU_NAME launch_motif_t16
#elif GIBBS_MOTIF_TU_T == 8
#define GIBBS_MOTIF_TU_NAME launch_motif_t8
#elif GIBBS_MOTIF_TU_T == 4
#define GIBBS_MOTIF_TU_NAME launch_motif_t4
#else
#define GIBBS_MOTIF_TU_NAME launch_motif_t1
#endif

cudaError_t GIBBS_MOTIF_TU_NAME(const MotifArgs &m, int grid, int smem, cudaStream_t stream) {
    switch ((m.c.k + 1) / 2) {
#define X(KPV) case KPV: return launch_one<KPV>(m, grid, smem, stream);
        X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15) X(16)
#undef X
    default: return cudaErrorInvalidValue;
    }
}

} // namespace gibbs
