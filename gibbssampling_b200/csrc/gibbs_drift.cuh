// gibbssampling_b200/csrc/gibbs_drift.cuh -- SiteSampler with the DATA-DERIVED background
// (doSiteSampling fs:697-701, getBestPWMSs fs:462-479, sweeps fs:554-611, shifts fs:483-550).
//
// The reference's background counts are mutated in place from window to window (quirk A.6-1):
//   F_w[b] = F0[b] + (w + 1) * count_n[b] - sum_{m <= w} occ_m[b]
// with F0 = base counts of the other sequences outside their sites (fs:565-568), count_n = base counts of
// the held-out sequence (re-added once per window, fs:471) and occ_m = base counts of window m (fs:472;
// the max(c-1, 0) clamp of fs:87 can never bind because every window's bases were just added). Each
// window then gets its own background pcv_w (fs:473) and PWM = PPM / pcv_w (fs:474), so a window score
// costs 4 + k float64 divisions. This kernel evaluates every window exactly in float64 (no ranking
// pass): one warp per chain, sweeps in the reference's sequential order.
#pragma once
#include "gibbs_device.cuh"
#include "gibbs_drift_dev.cuh"
#include "gibbs_kernels.cuh"

namespace gibbs {

struct DriftArgs {
    ChainArgs c;
    const double *pvals;     // [n] normalizePPM value of a count: (c + pc) / ((N-1) + |A| pc), fs:260
    const int32_t *basecnt;  // [n][4] base counts of every sequence
    int32_t gcnt[4];         // base counts of the whole set
    double alpha_pc;         // float alphabet.Length * pseudoCount (fs:117)
    double pc;
    int32_t fast_ok;         // the float32 ranking pass may be used (pc > 0 and no float64 under/overflow possible)
};

// one thread per sequence: A,C,G,T counts
__global__ void basecount_kernel(DeviceSeqs s, int32_t *basecnt) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= s.n) return;
    const uint32_t *row = s.packed + (size_t)n * s.row_words;
    const int L = s.len[n];
    int c[4] = {0, 0, 0, 0};
    for (int b = 0; b < L; ++b) {
        const int code = (row[b >> 4] >> ((b & 15) * 2)) & 3;
        c[0] += code == 0; c[1] += code == 1; c[2] += code == 2; c[3] += code == 3;
    }
    for (int b = 0; b < 4; ++b) basecnt[n * 4 + b] = c[b];
}

__global__ void pvals_kernel(int n, double pc, double den, double *pvals) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n) pvals[c] = __ddiv_rn(__dadd_rn((double)c, pc), den);
}

template <int KP>
__global__ void __launch_bounds__(32) drift_kernel(const DriftArgs d) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const ChainArgs &a = d.c;
    const int lane = threadIdx.x;
    const int chain = blockIdx.x;
    const TeamSmem S = carve_smem(smem_raw, 1);
    const WarpTables WT = warp_tables(S, 0);
    const int N = a.s.n, k = a.k;
    int32_t *sites = a.sites + (size_t)chain * N;
    double *hv = a.hv + (size_t)chain * N;
    double *scores = a.scores + (size_t)chain * N;
    const uint64_t chain_uid = (uint64_t)a.chain_id_base + (uint64_t)chain;

    RowRing<4> ring;
    ring.init(S, a.s, 0, lane);
    if (lane < 16) S.lut[lane] = hist_lut_entry(lane);
    __syncwarp();
    if (lane == 0) ring.fill(4);

    unsigned long long st_updates = 0, st_windows = 0, st_slow = 0;
    int st_sweeps = 0, capped = 0;
    uint32_t v = 0;
    int phase = next_phase(PH_INIT, a.phase_mask);
    int sweeps_in_phase = 0;
    while (phase != PH_DONE) {
        const int mode = phase == PH_LEFT ? SHIFT_LEFT : phase == PH_RIGHT ? SHIFT_RIGHT : SHIFT_NONE;
        if (phase == PH_LEFT || phase == PH_RIGHT || (phase == PH_GREEDY && sweeps_in_phase == 0))
            site_counts<KP, 1>(a.s, sites, -1, k, mode, S.total, S.lut, S.fix, lane);
        bool changed = false;
        for (int n = 0; n < N; ++n, ++v) {
            const uint32_t *row = ring.wait(v);
            const int len_n = __ldg(a.s.len + n);
            const int W = len_n - k + 1;
            int site_n = 0;
            double hv_n = 0.0;
            uint64_t own = 0;
            const int32_t *counts;
            if (phase == PH_INIT) {
                random_loo_counts_impl<KP, GIBBS_P0_NB_CHAIN, false>(a, chain_uid, chain, n, WT.counts, S.lut, lane, WT.lgcol);
                counts = WT.counts;
            } else {
                site_n = __ldcg(sites + n);
                hv_n = __ldcg(hv + n);
                own = kmer_shared<KP>(row, shifted_site(site_n, len_n, k, mode));
                counts = S.total;
            }
            // leave-one-out counts -> PPM (fs:573-575) in WT.wcol; bases inside the others' sites per base (column sums)
            int insite[4] = {0, 0, 0, 0};
            for (int e = lane; e < 4 * k; e += 32) {
                int c = counts[e];
                if (phase != PH_INIT && (int)((own >> (2 * (e >> 2))) & 3u) == (e & 3)) c -= 1;
                WT.wcol[e] = (phase == PH_INIT && a.ppm_given) ? __ldg(a.ppm_given + e) : __ldg(d.pvals + c); // fs:661 / fs:573-575
                WT.lgcol[e] = c;
            }
            __syncwarp();
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                int s = 0;
                for (int j = lane; j < k; j += 32) s += WT.lgcol[j * 4 + b];
                insite[b] = __reduce_add_sync(FULL, s);
            }
            // fused createFCVWithout of the others (fs:565-568): their bases outside their sites
            int f0[4], cn[4];
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                cn[b] = __ldg(d.basecnt + n * 4 + b);
                f0[b] = (d.gcnt[b] - cn[b]) - insite[b];
            }
            double p;
            int w;
            bool ranked = false;
            if (d.fast_ok && !(phase == PH_INIT && a.ppm_given)) { // (a supplied PPM may hold zeros or denormals: exact scan)
                // WT.counts was consumed above: its space holds the float32 log table
                drift_pair_table<KP>(WT.wcol, reinterpret_cast<float *>(WT.counts), reinterpret_cast<float *>(WT.ptab), k, lane);
                ranked = scan_drifting_fast<KP>(row, W, k, WT.wcol, reinterpret_cast<const float *>(WT.ptab), f0, cn, d.pc,
                                                d.alpha_pc, lane, p, w);
            }
            if (!ranked) {
                scan_drifting(row, W, k, WT.wcol, f0, cn, d.pc, d.alpha_pc, lane, p, w);
                st_slow += 1;
            }
            st_updates += 1;
            st_windows += (unsigned long long)W;
            if (phase == PH_INIT) {
                if (lane == 0) {
                    sites[n] = w;
                    hv[n] = p;
                }
            } else if (score_improves(p, hv_n, hv_n != hv_n ? __ldcg(scores + n) : 0.0)) { // fs:579
                if (w != site_n) {
                    changed = true;
                    if (phase == PH_GREEDY) {
                        const uint64_t neu = kmer_shared<KP>(row, w);
                        if (lane < k) {
                            const int bo = (int)((own >> (2 * lane)) & 3u), bn = (int)((neu >> (2 * lane)) & 3u);
                            if (bo != bn) {
                                S.total[lane * 4 + bo] -= 1;
                                S.total[lane * 4 + bn] += 1;
                            }
                        }
                    }
                }
                if (lane == 0) {
                    sites[n] = w;
                    hv[n] = p;
                }
            }
            __syncwarp();
            if (lane == 0) ring.fill(v + 1 + 4);
        }
        st_sweeps += 1;
        if (phase == PH_INIT) {
            phase = next_phase(PH_GREEDY, a.phase_mask);
            sweeps_in_phase = 0;
            continue;
        }
        ++sweeps_in_phase;
        bool next = !changed;
        if (!next && sweeps_in_phase >= a.max_sweeps) {
            next = true;
            capped = 1;
        }
        if (next) {
            sweeps_in_phase = 0;
            phase = next_phase(phase + 1, a.phase_mask);
        }
    }
    if (lane == 0)
        for (int i = 0; i < 4; ++i) ring.wait(v + (uint32_t)i);
    __syncwarp();
    for (int n = lane; n < N; n += 32) {
        const double x = __ldcg(hv + n);
        if (x == x) scores[n] = log2_ref(x);
    }
    __syncwarp();
    if (lane == 0) {
        double sum = 0.0;
        for (int n = 0; n < N; ++n) sum = __dadd_rn(sum, __ldcg(scores + n));
        a.sums[chain] = sum;
        atomicAdd(a.stats + ST_SITE_UPDATES, st_updates);
        atomicAdd(a.stats + ST_WINDOW_SCORES, st_windows);
        atomicAdd(a.stats + ST_EXACT_RESCANS, st_slow);
        atomicAdd(a.stats + ST_SWEEPS, (unsigned long long)st_sweeps);
        atomicAdd(a.stats + ST_CAPPED, (unsigned long long)capped);
    }
}

} // namespace gibbs
