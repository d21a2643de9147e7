// gibbssampling_b200/csrc/gibbs_drift_launch.cu -- instantiations of chain_kernel<KP, T, false, DRIFT = true>.
// A translation unit of its own: the 4-warp fixed-background kernel sits at its register limit and its code generation
// changed (spills) whenever the drifting-background instantiations were compiled in the same module.
#include "gibbs_kernels.cuh"

namespace gibbs {

template <int KPV, int TV>
static cudaError_t launch_one(const ChainArgs &a, int grid, int smem, cudaStream_t stream) {
    if (smem > 48 * 1024) {
        const cudaError_t e = cudaFuncSetAttribute(chain_kernel<KPV, TV, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
    }
    chain_kernel<KPV, TV, false, true><<<grid, 32 * TV, smem, stream>>>(a);
    return cudaGetLastError();
}

template <int KPV>
static cudaError_t launch_kp(int team, const ChainArgs &a, int grid, int smem, cudaStream_t stream) {
    switch (team) {
    case 8: return launch_one<KPV, 8>(a, grid, smem, stream);
    case 4: return launch_one<KPV, 4>(a, grid, smem, stream);
    case 1: return launch_one<KPV, 1>(a, grid, smem, stream);
    default: return cudaErrorInvalidValue;
    }
}

// team in {1, 4, 8}; smem = team_smem_bytes(row_words, team)
cudaError_t launch_drift_team(int team, const ChainArgs &a, int grid, int smem, cudaStream_t stream) {
    switch ((a.k + 1) / 2) {
#define X(KPV) case KPV: return launch_kp<KPV>(team, a, grid, smem, stream);
        X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15) X(16)
#undef X
    default: return cudaErrorInvalidValue;
    }
}

} // namespace gibbs
