// gibbssampling_b200/csrc/gibbs_cluster_tu.cu -- chain_cluster_kernel<KP, C> for the 16 k-widths, one cluster size per
// translation unit (-DGIBBS_CLUSTER_C=4|8, see _build.py).
#include "gibbs_cluster.cuh"

#if !defined(GIBBS_CLUSTER_C)
#error "compile with -DGIBBS_CLUSTER_C=4 or 8 (see _build.py)"
#endif

namespace gibbs {

template <int KPV>
static cudaError_t launch_one(const ChainArgs &a, int n_clusters, cudaStream_t stream, int *capacity_out) {
    auto kernel = chain_cluster_kernel<KPV, GIBBS_CLUSTER_C>;
    const size_t smem = cluster_smem_bytes<GIBBS_CLUSTER_C>(a.s.n, a.s.row_words);
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(n_clusters * GIBBS_CLUSTER_C), 1, 1);
    cfg.blockDim = dim3(32 * CL_T, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = GIBBS_CLUSTER_C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (capacity_out) { // how many clusters the device can hold at once (the hand-over threshold of the stage before)
        int n = 0;
        e = cudaOccupancyMaxActiveClusters(&n, kernel, &cfg);
        if (e != cudaSuccess) return e;
        *capacity_out = n;
        return cudaSuccess;
    }
    return cudaLaunchKernelEx(&cfg, kernel, a);
}

#if GIBBS_CLUSTER_C == 4
#define GIBBS_CLUSTER_NAME launch_chain_cluster4
#define GIBBS_CLUSTER_SMEM_NAME launch_chain_cluster4_smem
#else
#define GIBBS_CLUSTER_NAME launch_chain_cluster8
#define GIBBS_CLUSTER_SMEM_NAME launch_chain_cluster8_smem
#endif

// capacity_out != null: only report how many clusters fit the device at once
cudaError_t GIBBS_CLUSTER_NAME(const ChainArgs &a, int n_clusters, cudaStream_t stream, int *capacity_out) {
    switch ((a.k + 1) / 2) {
#define X(KPV) case KPV: return launch_one<KPV>(a, n_clusters, stream, capacity_out);
        X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15) X(16)
#undef X
    default: return cudaErrorInvalidValue;
    }
}

size_t GIBBS_CLUSTER_SMEM_NAME(int n, int row_words) { return cluster_smem_bytes<GIBBS_CLUSTER_C>(n, row_words); }

} // namespace gibbs
