// gibbssampling_b200/csrc/gibbs_motif2_tu.cu -- motif2_kernel<KP> (MotifSampler, motifAmount = 2) for the 16 k-widths.
#include "gibbs_motif2.cuh"

namespace gibbs {

template <int KPV>
static cudaError_t launch_one(const Motif2Args &q, int grid, int smem, cudaStream_t stream) {
    auto kernel = motif2_kernel<KPV>;
    if (smem > 48 * 1024) {
        const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
    }
    kernel<<<grid, 32, smem, stream>>>(q);
    return cudaGetLastError();
}

cudaError_t launch_motif2(const Motif2Args &q, int grid, int smem, cudaStream_t stream) {
    switch ((q.m.c.k + 1) / 2) {
#define X(KPV) case KPV: return launch_one<KPV>(q, grid, smem, stream);
        X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15) X(16)
#undef X
    default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_motif2_seed(const int32_t *sites, long long cells, int32_t *pos2, cudaStream_t stream) {
    motif2_seed_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, stream>>>(sites, cells, pos2);
    return cudaGetLastError();
}

} // namespace gibbs
