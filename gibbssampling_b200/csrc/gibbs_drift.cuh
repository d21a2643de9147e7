// gibbssampling_b200/csrc/gibbs_drift.cuh -- SiteSampler with the DATA-DERIVED background
// (doSiteSampling fs:697-701, getBestPWMSs fs:462-479, sweeps fs:554-611, shifts fs:483-550).
//
// The reference's background counts are mutated in place from window to window (quirk A.6-1):
//   F_w[b] = F0[b] + (w + 1) * count_n[b] - sum_{m <= w} occ_m[b]
// with F0 = base counts of the other sequences outside their sites (fs:565-568), count_n = base counts of
// the held-out sequence (re-added once per window, fs:471) and occ_m = base counts of window m (fs:472;
// the max(c-1, 0) clamp of fs:87 can never bind because every window's bases were just added). Each
// window then gets its own background pcv_w (fs:473) and PWM = PPM / pcv_w (fs:474), so a window score
// costs 4 + k float64 divisions. The scan itself lives in gibbs_drift_dev.cuh and runs inside
// chain_kernel<.., DRIFT = true> (gibbs_drift_launch.cu); this header keeps the two setup kernels.
#pragma once
#include "gibbs_device.cuh"
#include "gibbs_drift_dev.cuh"
#include "gibbs_kernels.cuh"

namespace gibbs {


// one thread per sequence: A,C,G,T counts, and the number of symbols outside A,C,G,T (maskcnt may be null)
static __global__ void basecount_kernel(DeviceSeqs s, int32_t *basecnt, int32_t *maskcnt) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= s.n) return;
    const uint32_t *row = s.packed + (size_t)n * s.row_words;
    const uint32_t *mrow = s.mask ? s.mask + (size_t)n * s.row_words : nullptr;
    const int L = s.len[n];
    int c[4] = {0, 0, 0, 0}, m = 0;
    for (int b = 0; b < L; ++b) {
        const int code = (row[b >> 4] >> ((b & 15) * 2)) & 3;
        if (mrow && ((mrow[b >> 4] >> ((b & 15) * 2)) & 1)) {
            ++m;
            continue;
        }
        c[0] += code == 0; c[1] += code == 1; c[2] += code == 2; c[3] += code == 3;
    }
    for (int b = 0; b < 4; ++b) basecnt[n * 4 + b] = c[b];
    if (maskcnt) maskcnt[n] = m;
}

// ss[n][t] = sum_{u < t} P_n[u] with P_n[u] = base counts (A,C,G,T; other symbols skipped) of the first u bases of
// sequence n, t = 0 .. len + 1: the prefix-of-prefix table of scan_drifting_tables. One thread per sequence (run once).
static __global__ void prefix_kernel(DeviceSeqs s, int stride, int4 *ss) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= s.n) return;
    const uint32_t *row = s.packed + (size_t)n * s.row_words;
    const uint32_t *mrow = s.mask ? s.mask + (size_t)n * s.row_words : nullptr;
    int4 *out = ss + (size_t)n * stride;
    const int L = s.len[n];
    int4 P = make_int4(0, 0, 0, 0), S = make_int4(0, 0, 0, 0);
    for (int t = 0; t <= L + 1 && t < stride; ++t) {
        out[t] = S;                     // S = sum_{u < t} P[u]
        S.x += P.x; S.y += P.y; S.z += P.z; S.w += P.w;
        if (t < L) {                    // P becomes the counts of the first t + 1 bases
            const int code = (row[t >> 4] >> ((t & 15) * 2)) & 3;
            const bool masked = mrow && ((mrow[t >> 4] >> ((t & 15) * 2)) & 1);
            if (!masked) {
                P.x += code == 0; P.y += code == 1; P.z += code == 2; P.w += code == 3;
            }
        }
    }
}

static __global__ void pvals_kernel(int n, double pc, double den, double *pvals) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n) pvals[c] = __ddiv_rn(__dadd_rn((double)c, pc), den);
}

} // namespace gibbs
