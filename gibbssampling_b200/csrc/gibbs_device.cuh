// gibbssampling_b200/csrc/gibbs_device.cuh -- device-side building blocks (sm_100a only).
//
// One warp owns one chain (= one restart of the reference, fs:691-695). Everything a site update
// needs lives in that warp's shared memory: the k x 4 count matrix, the float64 PWM column table,
// the fixed-point log2-odds pair table and two TMA-staged rows of 2-bit packed sequence.
//
// Exactness strategy (DESIGN.md "Kernels"): every window is first scored with an int32
// fixed-point sum of log2-odds looked up two bases at a time (one LDS + half an IADD3 per
// column pair). That ranks windows up to a rigorous rounding margin; the few windows inside the
// margin of the maximum are re-scored with the reference's own arithmetic -- a left-to-right
// float64 product of odds ratios (fs:290-293) -- so the argmax and its score are bit-identical
// to the reference's float64 scan (fs:301-314).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gibbs {

constexpr unsigned FULL = 0xffffffffu;
constexpr int LG_FRAC_BITS = 11;          // log2-odds fixed point: 2^-11 units ...
constexpr int KEY_IDX_BITS = 8;           // ... shifted left by 8 to leave room for a window index
constexpr int LG_ENTRY_SHIFT = LG_FRAC_BITS + KEY_IDX_BITS;
constexpr int MAX_COLS = 32;              // GIBBS_MAX_K
constexpr double LN2 = 0.6931471805599453; // log 2.0 as float64 (FSharpAux log2 = ln x / ln 2)

enum ShiftMode { SHIFT_NONE = 0, SHIFT_LEFT = 1, SHIFT_RIGHT = 2 };
enum StatSlot { ST_SITE_UPDATES = 0, ST_WINDOW_SCORES = 1, ST_SWEEPS = 2, ST_EXACT_RESCANS = 3, ST_CAPPED = 4, ST_NSLOTS = 8 };

// one entry per (count, base): the odds ratio W = ((c + pc) / den) / q[b] of fs:260 + fs:286 as
// float64, and round(log2 W * 2^11) << 8 for the ranking pass
struct __align__(16) WEnt {
    double w;
    int32_t lg;
    int32_t pad;
};

struct DeviceSeqs {
    const uint32_t *packed; // [n][row_words], base b of a row in bits [2(b%16), +2) of word b/16
    const int32_t *len;     // [n]
    int32_t n;
    int32_t row_words;      // multiple of 4 (16 B) and >= ceil(max_len/16) + 4 zero words
};

struct ChainArgs {
    DeviceSeqs s;
    const WEnt *wtab;       // [n][4]
    int32_t k;
    int32_t phase_shifts;
    int32_t max_sweeps;
    int32_t fast_ok;
    int32_t sampler;
    int32_t phase_mask;     // GIBBS_PHASE_* bits to run, in pipeline order
    int32_t rng_mode;       // 0 Philox, 1 injected
    uint64_t seed;
    int64_t chain_id_base;
    const double *uniforms;
    int64_t uniforms_per_chain;
    int32_t n_chains;
    int32_t *sites;         // [chains][n]
    double *hv;             // [chains][n] raw float64 score (odds product or background product)
    double *scores;         // [chains][n] log2 / PWMS
    double *sums;           // [chains]
    unsigned long long *stats;
    double cutoff;
    double bg[4];
};

// ------------------------------------------------------------------------------------------------
// per-warp shared memory
// ------------------------------------------------------------------------------------------------
struct WarpSmem {
    uint64_t *bar;   // [2] mbarriers of the two row buffers
    int32_t *total;  // [32*4] counts, entry j*4+b
    int32_t *lgcol;  // [32*4] fixed-point log2 odds per column
    int32_t *ptab;   // [16*16] pair table: ptab[p*16 + nib] = lgcol[2p][nib&3] + lgcol[2p+1][nib>>2]
    double *wcol;    // [32*4] float64 odds per column
    double *scratch; // [32]
    uint32_t *row[2];
};

constexpr int WARP_SMEM_FIXED = 16 + 512 + 512 + 1024 + 1024 + 256; // bytes before the row buffers

__host__ __device__ inline int warp_smem_bytes(int row_words) { return WARP_SMEM_FIXED + 2 * row_words * 4; }

__device__ __forceinline__ WarpSmem carve_smem(unsigned char *base, int row_words) {
    WarpSmem s;
    s.bar = reinterpret_cast<uint64_t *>(base);
    s.total = reinterpret_cast<int32_t *>(base + 16);
    s.lgcol = reinterpret_cast<int32_t *>(base + 16 + 512);
    s.ptab = reinterpret_cast<int32_t *>(base + 16 + 1024);
    s.wcol = reinterpret_cast<double *>(base + 16 + 2048);
    s.scratch = reinterpret_cast<double *>(base + 16 + 3072);
    s.row[0] = reinterpret_cast<uint32_t *>(base + WARP_SMEM_FIXED);
    s.row[1] = s.row[0] + row_words;
    return s;
}

// ------------------------------------------------------------------------------------------------
// PTX wrappers: mbarrier + 1-D bulk async copy (TMA, SASS UBLKCP)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// Double-buffered row staging: one elected lane issues the bulk copy of the NEXT held-out row while
// the warp scores the current one.
struct RowPipe {
    uint64_t *bar;    // [2]
    uint32_t *row0;   // slot s lives at row0 + s * row_words
    const uint32_t *gpacked;
    int row_words;
    uint32_t phase;   // bit s = parity to wait for on slot s
    uint32_t visit;

    __device__ __forceinline__ void init(const WarpSmem &s, const DeviceSeqs &d, int lane) {
        bar = s.bar;
        row0 = s.row[0];
        gpacked = d.packed;
        row_words = d.row_words;
        phase = 0;
        visit = 0;
        if (lane == 0) {
            mbar_init(bar, 1);
            mbar_init(bar + 1, 1);
            fence_barrier_init();
        }
        __syncwarp();
    }
    __device__ __forceinline__ void issue(int slot, int seq, int lane) {
        if (lane == 0) {
            const uint32_t bytes = (uint32_t)row_words * 4u;
            mbar_expect_tx(bar + slot, bytes);
            bulk_g2s(row0 + slot * row_words, gpacked + (size_t)seq * row_words, bytes, bar + slot);
        }
    }
    // row of the current visit is (or was prefetched) in slot visit&1; prefetch `next` into the other slot
    __device__ __forceinline__ const uint32_t *acquire(int next, int lane) {
        const int slot = visit & 1;
        mbar_wait(bar + slot, (phase >> slot) & 1u);
        phase ^= 1u << slot;
        __syncwarp(); // every lane is done reading the other slot (previous visit)
        issue(slot ^ 1, next, lane);
        ++visit;
        return row0 + slot * row_words;
    }
    __device__ __forceinline__ void drain() { // one prefetch is always outstanding
        const int slot = visit & 1;
        mbar_wait(bar + slot, (phase >> slot) & 1u);
        phase ^= 1u << slot;
    }
};

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al.), the counter-based uniform stream; same constants as the oracle
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

// ------------------------------------------------------------------------------------------------
// 2-bit k-mer extraction
// ------------------------------------------------------------------------------------------------
// 64 bits starting at base `pos` of a row (global memory, read-only path)
template <int KP>
__device__ __forceinline__ uint64_t kmer_global(const uint32_t *__restrict__ rowp, int pos) {
    const uint32_t *p = rowp + (pos >> 4);
    const int sh = (pos & 15) * 2;
    const uint32_t w0 = __ldg(p), w1 = __ldg(p + 1);
    uint32_t lo = __funnelshift_r(w0, w1, sh), hi = w1 >> sh;
    if (KP > 8) {
        const uint32_t w2 = __ldg(p + 2);
        hi = __funnelshift_r(w1, w2, sh);
    }
    return ((uint64_t)hi << 32) | lo;
}
template <int KP>
__device__ __forceinline__ uint64_t kmer_shared(const uint32_t *rowp, int pos) {
    const uint32_t *p = rowp + (pos >> 4);
    const int sh = (pos & 15) * 2;
    const uint32_t w0 = p[0], w1 = p[1];
    uint32_t lo = __funnelshift_r(w0, w1, sh), hi = w1 >> sh;
    if (KP > 8) {
        const uint32_t w2 = p[2];
        hi = __funnelshift_r(w1, w2, sh);
    }
    return ((uint64_t)hi << 32) | lo;
}

__device__ __forceinline__ int shifted_site(int pos, int len, int k, int mode) {
    if (mode == SHIFT_LEFT) return pos > 0 ? pos - 1 : pos;             // fs:358
    if (mode == SHIFT_RIGHT) return pos <= len - k - 1 ? pos + 1 : pos; // fs:327
    return pos;
}

// ------------------------------------------------------------------------------------------------
// per-lane packed histogram of k-mers: 4 byte-wide counters (A,C,G,T) per column register
// (replaces createPFMOf + fusePositionFrequencyMatrices, fs:211-226)
// ------------------------------------------------------------------------------------------------
template <int KP>
struct Hist {
    uint32_t c[2 * KP];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int j = 0; j < 2 * KP; ++j) c[j] = 0;
    }
    __device__ __forceinline__ void add(uint64_t kmer) {
#pragma unroll
        for (int j = 0; j < 2 * KP; ++j) {
            const uint32_t b8 = ((uint32_t)(kmer >> (2 * j)) & 3u) * 8u;
            c[j] += 1u << b8;
        }
    }
    // at most 255 adds per lane since the last clear; total[] += warp sums
    __device__ __forceinline__ void flush_add(int32_t *total, int k, int lane) {
#pragma unroll
        for (int j = 0; j < 2 * KP; ++j) {
            if (j < k) {
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const uint32_t s = __reduce_add_sync(FULL, (c[j] >> (8 * b)) & 255u);
                    if (lane == ((j * 4 + b) & 31)) total[j * 4 + b] += (int32_t)s;
                }
            }
            c[j] = 0;
        }
        __syncwarp();
    }
};

__device__ __forceinline__ void zero_total(int32_t *total, int lane) {
#pragma unroll
    for (int e = lane; e < MAX_COLS * 4; e += 32) total[e] = 0;
    __syncwarp();
}

// counts over the sites of all sequences except `exclude` (sites < 0 = no site), positions shifted
// by `mode`: the fused PFM of fs:392-396 (exclude = held-out) or the all-sites total
template <int KP>
__device__ __forceinline__ void site_counts(const DeviceSeqs &s, const int32_t *sites, int exclude, int k, int mode,
                                            int32_t *total, int lane) {
    zero_total(total, lane);
    Hist<KP> h;
    h.clear();
    const int iters = (s.n + 31) >> 5;
    for (int it0 = 0; it0 < iters; it0 += 255) { // byte counters hold 255 adds per lane
        const int it1 = min(iters, it0 + 255);
        for (int it = it0; it < it1; ++it) {
            const int i = it * 32 + lane;
            if (i < s.n && i != exclude) {
                const int site = __ldcg(sites + i);
                if (site >= 0) {
                    const int pos = shifted_site(site, __ldg(s.len + i), k, mode);
                    h.add(kmer_global<KP>(s.packed + (size_t)i * s.row_words, pos));
                }
            }
        }
        h.flush_add(total, k, lane);
    }
}

// ------------------------------------------------------------------------------------------------
// PWM tables for one held-out sequence
// ------------------------------------------------------------------------------------------------
// Leave-one-out count c = total - own, then W(c, b) and its fixed-point log2 are gathered from the
// precomputed table (normalizePPM fs:255-261 + createPositionWeightMatrix fs:282-287 evaluated once
// per distinct count instead of once per window).
template <int KP>
__device__ __forceinline__ void build_tables(const WarpSmem &S, bool has_own, uint64_t own, int k,
                                             const WEnt *__restrict__ wtab, int lane) {
#pragma unroll
    for (int e = lane; e < 8 * KP; e += 32) {
        const int j = e >> 2, b = e & 3;
        double w = 1.0;
        int32_t lg = 0;
        if (j < k) {
            int c = S.total[e];
            if (has_own && (int)((own >> (2 * j)) & 3u) == b) c -= 1;
            const int4 raw = __ldg(reinterpret_cast<const int4 *>(wtab + (size_t)c * 4 + b));
            w = __hiloint2double(raw.y, raw.x);
            lg = raw.z;
        }
        S.wcol[e] = w;
        S.lgcol[e] = lg;
    }
    __syncwarp();
#pragma unroll
    for (int idx = lane; idx < 16 * KP; idx += 32) {
        const int p = idx >> 4, nib = idx & 15;
        S.ptab[idx] = S.lgcol[(2 * p) * 4 + (nib & 3)] + S.lgcol[(2 * p + 1) * 4 + (nib >> 2)];
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------------------------
// exact float64 window score: Array.fold (fun (pos, v) b -> pos + 1, v * pwm.[b, pos]) (0, 1.) (fs:290-293)
// ------------------------------------------------------------------------------------------------
template <int KP>
__device__ __forceinline__ double exact_window(const uint32_t *row, int w, int k, const double *wcol) {
    const uint64_t kmer = kmer_shared<KP>(row, w);
    double f[2 * KP];
#pragma unroll
    for (int j = 0; j < 2 * KP; ++j) f[j] = wcol[j * 4 + (int)((kmer >> (2 * j)) & 3u)]; // dummy column = 1.0
    double p = 1.0;
#pragma unroll
    for (int j = 0; j < 2 * KP; ++j)
        if (j < k) p = __dmul_rn(p, f[j]);
    return p;
}

// (max value, lowest index) over the warp; every lane ends with the result
__device__ __forceinline__ void warp_argmax(double &hv, int &w) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const double ohv = __shfl_xor_sync(FULL, hv, o);
        const int ow = __shfl_xor_sync(FULL, w, o);
        if (ohv > hv || (ohv == hv && ow < w)) {
            hv = ohv;
            w = ow;
        }
    }
}

// the reference loop itself (fs:302-314): every window in float64, first strict maximum from (0., 0)
template <int KP>
__device__ __forceinline__ void scan_exact_all(const uint32_t *row, int W, int k, const double *wcol, int lane,
                                               double &hv_out, int &w_out) {
    double hv = 0.0;
    int hw = 0;
    for (int w = lane; w < W; w += 32) {
        const double p = exact_window<KP>(row, w, k, wcol);
        if (p > hv) {
            hv = p;
            hw = w;
        }
    }
    warp_argmax(hv, hw);
    hv_out = hv;
    w_out = hw;
}

// ------------------------------------------------------------------------------------------------
// ranking pass: int32 fixed-point log2-odds, two columns per lookup
// ------------------------------------------------------------------------------------------------
// A chunk is CH consecutive windows handled by one lane. Even windows 2e use the nibbles at bit
// 4(e+p) of the chunk's aligned bit string, odd windows those at bit 4(e+p)+2, so CH/2 + KP - 1
// nibble extractions per parity serve CH/2 windows x KP lookups. key = (sum << 8) | (255 - local).
template <int KP, int CH, bool MASK>
__device__ __forceinline__ void score_chunk(const uint32_t *a, const uint32_t *b, const int32_t *ptab, int idx0, int lim,
                                            int32_t &m1, int32_t &m2) {
    constexpr int NE = CH / 2 + KP - 1;
    int nibE[NE], nibO[NE];
#pragma unroll
    for (int t = 0; t < NE; ++t) {
        nibE[t] = (a[t >> 3] >> (4 * (t & 7))) & 15;
        nibO[t] = (b[t >> 3] >> (4 * (t & 7))) & 15;
    }
#pragma unroll
    for (int i = 0; i < CH; ++i) {
        const int e = i >> 1;
        int32_t key = idx0 - i;
#pragma unroll
        for (int p = 0; p < KP; ++p) key += ptab[p * 16 + ((i & 1) ? nibO[e + p] : nibE[e + p])];
        if (MASK) key = (i < lim) ? key : INT32_MIN;
        m2 = max(m2, min(m1, key));
        m1 = max(m1, key);
    }
}

template <int KP, int CH>
struct ScanGeom {
    static constexpr int NB = CH + 2 * KP - 1;  // bases a chunk touches
    static constexpr int NWA = (NB + 15) / 16;  // aligned words
    static constexpr int RPS = 256 / CH;        // rounds per segment (local index fits 8 bits)
    static constexpr int CPS = 32 * RPS;        // chunks per segment
};

// one segment: chunks [c_begin, c_end), at most CPS of them
template <int KP, int CH>
__device__ __forceinline__ void scan_segment(const uint32_t *row, int W, const int32_t *ptab, int c_begin, int c_end,
                                             int lane, int32_t &m1, int32_t &m2) {
    using G = ScanGeom<KP, CH>;
    m1 = INT32_MIN;
    m2 = INT32_MIN;
    int r = 0;
    for (int c = c_begin + lane; c < c_end; c += 32, ++r) {
        const int base0 = c * CH;
        const uint32_t *p = row + (base0 >> 4);
        uint32_t a[G::NWA + 1], b[G::NWA];
        if (CH == 16) {
#pragma unroll
            for (int i = 0; i < G::NWA; ++i) a[i] = p[i];
        } else {
            const int sh = (base0 & 15) * 2;
            uint32_t raw[G::NWA + 1];
#pragma unroll
            for (int i = 0; i <= G::NWA; ++i) raw[i] = p[i];
#pragma unroll
            for (int i = 0; i < G::NWA; ++i) a[i] = __funnelshift_r(raw[i], raw[i + 1], sh);
        }
        a[G::NWA] = 0;
#pragma unroll
        for (int i = 0; i < G::NWA; ++i) b[i] = __funnelshift_r(a[i], a[i + 1], 2);
        const int lim = W - base0;
        const int idx0 = 255 - r * CH;
        if (lim >= CH) score_chunk<KP, CH, false>(a, b, ptab, idx0, lim, m1, m2);
        else score_chunk<KP, CH, true>(a, b, ptab, idx0, lim, m1, m2);
    }
}

// per-lane best key M1 (from segment S1), second best M2, over all windows of the row
template <int KP, int CH>
__device__ __forceinline__ void scan_fast(const uint32_t *row, int W, const int32_t *ptab, int lane, int32_t &M1,
                                          int32_t &M2, int &S1) {
    using G = ScanGeom<KP, CH>;
    const int n_chunks = (W + CH - 1) / CH;
    M1 = INT32_MIN;
    M2 = INT32_MIN;
    S1 = 0;
    for (int seg = 0, c0 = 0; c0 < n_chunks; ++seg, c0 += G::CPS) {
        int32_t m1, m2;
        scan_segment<KP, CH>(row, W, ptab, c0, min(n_chunks, c0 + G::CPS), lane, m1, m2);
        if (m1 > M1) {
            M2 = max(max(M2, M1), m2);
            M1 = m1;
            S1 = seg;
        } else {
            M2 = max(M2, m1);
        }
    }
}

template <int KP, int CH>
__device__ __forceinline__ int decode_window(int32_t key, int seg, int lane) {
    using G = ScanGeom<KP, CH>;
    const int local = 255 - (key & 255);
    const int r = local / CH, i = local % CH;
    return (seg * G::CPS + lane + 32 * r) * CH + i;
}

// getBestPWMSsWithBPV (fs:301-314) for the staged row: (raw float64 maximum, first argmax).
// Returns true when the all-windows float64 path had to be taken.
template <int KP, int CH>
__device__ __forceinline__ bool pick_argmax_ch(const WarpSmem &S, const uint32_t *row, int W, int k, int lane,
                                               double &hv_out, int &w_out) {
    int32_t M1, M2;
    int S1;
    scan_fast<KP, CH>(row, W, S.ptab, lane, M1, M2, S1);
    const int32_t M = __reduce_max_sync(FULL, M1);
    // |key/256 - true log2 score * 2^11| <= k/2 units for every window, so any window whose exact
    // product can reach the maximum has key >= M - (k + 1) units (index bits: 255 more)
    const int32_t thr = M - (((k + 1) << KEY_IDX_BITS) + 255);
    if (__ballot_sync(FULL, M2 >= thr)) return true; // two candidates in one lane: rescan exactly
    const bool is_cand = M1 >= thr;
    const unsigned cand = __ballot_sync(FULL, is_cand);
    double p = 0.0;
    int w = INT32_MAX;
    if (is_cand) {
        w = decode_window<KP, CH>(M1, S1, lane);
        p = exact_window<KP>(row, w, k, S.wcol);
    }
    if (__popc(cand) == 1) {
        const int src = __ffs(cand) - 1;
        p = __shfl_sync(FULL, p, src);
        w = __shfl_sync(FULL, w, src);
    } else {
        warp_argmax(p, w);
    }
    hv_out = p;
    w_out = w;
    return false;
}

template <int KP>
__device__ __forceinline__ bool pick_argmax(const WarpSmem &S, const uint32_t *row, int W, int k, int fast_ok, int lane,
                                            double &hv_out, int &w_out) {
    bool slow = !fast_ok;
    if (!slow) {
        if (W > 256) slow = pick_argmax_ch<KP, 16>(S, row, W, k, lane, hv_out, w_out);
        else if (W > 128) slow = pick_argmax_ch<KP, 8>(S, row, W, k, lane, hv_out, w_out);
        else slow = pick_argmax_ch<KP, 4>(S, row, W, k, lane, hv_out, w_out);
    }
    if (slow) scan_exact_all<KP>(row, W, k, S.wcol, lane, hv_out, w_out);
    return slow;
}

__device__ __forceinline__ double log2_ref(double x) { return log(x) / LN2; } // fs:303 via FSharpAux

// `if fst tmp > fst acc.[n]` (fs:402) on log2 scores, decided from the raw products whenever the
// two logs cannot collide
// hv_old = NaN marks a caller-supplied start state: only its log2 score (score_old) is known.
__device__ __forceinline__ bool score_improves(double hv_new, double hv_old, double score_old) {
    if (hv_old != hv_old) return log2_ref(hv_new) > score_old;
    if (!(hv_new > hv_old)) return false;
    if (hv_new * (1.0 - 0x1p-30) > hv_old) return true;
    return log2_ref(hv_new) > log2_ref(hv_old);
}

} // namespace gibbs
