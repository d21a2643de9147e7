// gibbssampling_b200/csrc/gibbs_pool.cuh -- warp-pool scheduling of the SiteSampler chains.
//
// One persistent CTA per SM holds up to POOL_SLOTS chains in shared memory. Its warps are a pool of
// workers: a worker claims one site update (slot, index) of some chain's open ROUND, computes it with
// the same device routines as chain_kernel (leave-one-out tables, ranking pass, float64 verification)
// and posts the result; the worker that posts the last result of a round closes it: validates the
// speculation, commits, updates the counts, advances the sweep / phase state machine and opens the next
// round. No warp ever waits at a barrier for a slower one, and when chains finish early their share of
// the pool goes to the chains still running (rounds grow from POOL_WARPS/slots up to POOL_MAXW wide),
// which removes most of the straggler tail of a one-wave launch.
//
// The committed sequence of site updates of every chain is the same as in chain_kernel and in the
// reference's sequential sweeps; only who computes what, and when, differs.
#pragma once
#include "gibbs_kernels.cuh"

namespace gibbs {

constexpr int POOL_WARPS = 28;  // 896 threads x 72 registers = one SM's register file
constexpr int POOL_SLOTS = 8;   // chains resident per CTA
constexpr int POOL_RING = 16;   // staged rows per chain
constexpr int POOL_MAXW = 8;    // widest round

enum SlotState { SLOT_OPEN = 1, SLOT_CLOSED = 2, SLOT_FINISHED = 3 };

struct SlotCtl {
    int state, taken, done, width;
    int chain, phase, n0, cur_blk;
    uint32_t vbase, issued;
    int next_seq, sweeps_in_phase, changed, gwidth, capped, st_sweeps;
    unsigned long long st_updates, st_windows, st_slow, st_spec;
};

struct SlotRes {
    double p;
    uint64_t own, neu;
    int w, flag, Wn, pad;
};

struct SlotMem {
    SlotCtl *ctl;
    SlotRes *res;       // [POOL_MAXW]
    int32_t *total;     // [128]
    double *blk_hv;     // [2][32]
    int32_t *blk_site;  // [2][32]
    int32_t *blk_len;   // [2][32]
    uint64_t *bar;      // [POOL_RING]
    uint32_t *rows;     // [POOL_RING][row_words]
};

constexpr int SLOT_FIXED_BYTES = 128 + POOL_MAXW * 40 + 512 + 512 + 256 + 256 + POOL_RING * 8; // 2112
static_assert(sizeof(SlotCtl) <= 128, "SlotCtl must fit its 128-byte cell");
static_assert(sizeof(SlotRes) == 40, "SlotRes layout");

__host__ __device__ inline int pool_slot_bytes(int row_words) { return SLOT_FIXED_BYTES + POOL_RING * row_words * 4; }
__host__ __device__ inline int pool_smem_bytes(int row_words) {
    return 128 /* lut + counters */ + POOL_WARPS * WARP_TABLE_BYTES + POOL_SLOTS * pool_slot_bytes(row_words);
}

__device__ __forceinline__ SlotMem slot_mem(unsigned char *base, int row_words, int slot) {
    unsigned char *b = base + 128 + POOL_WARPS * WARP_TABLE_BYTES + slot * pool_slot_bytes(row_words);
    SlotMem m;
    m.ctl = reinterpret_cast<SlotCtl *>(b);
    m.res = reinterpret_cast<SlotRes *>(b + 128);
    m.total = reinterpret_cast<int32_t *>(b + 128 + POOL_MAXW * 40);
    m.blk_hv = reinterpret_cast<double *>(b + 128 + POOL_MAXW * 40 + 512);
    m.blk_site = reinterpret_cast<int32_t *>(b + 128 + POOL_MAXW * 40 + 1024);
    m.blk_len = reinterpret_cast<int32_t *>(b + 128 + POOL_MAXW * 40 + 1280);
    m.bar = reinterpret_cast<uint64_t *>(b + 128 + POOL_MAXW * 40 + 1536);
    m.rows = reinterpret_cast<uint32_t *>(b + SLOT_FIXED_BYTES);
    return m;
}

struct PoolArgs {
    ChainArgs c;
    int *next_chain; // global counter of the next chain to start
};

// ---- ring helpers (one ring per slot; visit v <-> slot v % POOL_RING, phase (v / POOL_RING) & 1) ----
__device__ __forceinline__ void pool_ring_fill(const SlotMem &M, const DeviceSeqs &s, uint32_t upto) { // one lane
    SlotCtl *c = M.ctl;
    const uint32_t bytes = (uint32_t)s.row_words * 4u;
    uint32_t issued = c->issued;
    int seq = c->next_seq;
    while (issued != upto) {
        const int r = (int)(issued & (POOL_RING - 1));
        mbar_expect_tx(M.bar + r, bytes);
        bulk_g2s(M.rows + r * s.row_words, s.packed + (size_t)seq * s.row_words, bytes, M.bar + r);
        seq = seq + 1 < s.n ? seq + 1 : 0;
        ++issued;
    }
    c->issued = issued;
    c->next_seq = seq;
}
__device__ __forceinline__ const uint32_t *pool_ring_wait(const SlotMem &M, int row_words, uint32_t v) {
    const int r = (int)(v & (POOL_RING - 1));
    mbar_wait(M.bar + r, (v / POOL_RING) & 1u);
    return M.rows + r * row_words;
}

__device__ __forceinline__ void pool_load_block(const SlotMem &M, const ChainArgs &a, int b, int phase, int lane) {
    const int N = a.s.n;
    const int i = b * 32 + lane;
    if (i < N) {
        const int o = (b & 1) * 32 + lane;
        M.blk_len[o] = __ldg(a.s.len + i);
        if (phase != PH_INIT) {
            const size_t off = (size_t)M.ctl->chain * N + i;
            M.blk_site[o] = __ldcg(a.sites + off);
            M.blk_hv[o] = __ldcg(a.hv + off);
        }
    }
}

// round width this slot may open: the pool divided by the chains still running
__device__ __forceinline__ int pool_target_width(int active_slots) {
    const int w = POOL_WARPS / max(1, active_slots);
    return min(POOL_MAXW, max(2, w));
}

// start a sweep of the slot's current phase (closer warp, all lanes)
template <int KP>
__device__ __forceinline__ void pool_start_sweep(const SlotMem &M, const ChainArgs &a, const uint32_t *lut, int lane) {
    SlotCtl *c = M.ctl;
    const int phase = c->phase;
    const int mode = phase == PH_LEFT ? SHIFT_LEFT : phase == PH_RIGHT ? SHIFT_RIGHT : SHIFT_NONE;
    if (phase == PH_LEFT || phase == PH_RIGHT || (phase == PH_GREEDY && c->sweeps_in_phase == 0))
        site_counts<KP, 1>(a.s, a.sites + (size_t)c->chain * a.s.n, -1, a.k, mode, M.total, lut, lane);
    pool_load_block(M, a, 0, phase, lane);
    pool_load_block(M, a, 1, phase, lane);
    if (lane == 0) {
        c->n0 = 0;
        c->cur_blk = 0;
        c->changed = 0;
        c->gwidth = 1;
    }
    __syncwarp();
}

// bind a chain to the slot (closer warp). The ring keeps its visit numbering across chains: the rows
// prefetched past the end of the previous chain's last sweep are sequences 0,1,2,... = what the new chain needs.
__device__ __forceinline__ void pool_bind_chain(const SlotMem &M, const ChainArgs &a, int chain, int lane) {
    SlotCtl *c = M.ctl;
    if (lane == 0) {
        c->chain = chain;
        c->phase = next_phase(PH_INIT, a.phase_mask);
        c->sweeps_in_phase = 0;
        c->capped = 0;
        c->st_sweeps = 0;
        c->st_updates = c->st_windows = c->st_slow = c->st_spec = 0;
    }
    __syncwarp();
}

// open the next round of a slot whose n0 / phase are set (closer warp)
__device__ __forceinline__ void pool_open_round(const SlotMem &M, const ChainArgs &a, int active_slots, int lane) {
    SlotCtl *c = M.ctl;
    const int N = a.s.n;
    if ((c->n0 >> 5) != c->cur_blk) { // entering a new block: fetch the one after it
        const int nb = c->n0 >> 5;
        __syncwarp();
        if (lane == 0) c->cur_blk = nb;
        pool_load_block(M, a, nb + 1, c->phase, lane);
    }
    __syncwarp();
    if (lane == 0) {
        int w = pool_target_width(active_slots);
        if (c->phase == PH_GREEDY) w = min(w, c->gwidth);
        w = min(w, N - c->n0);
        pool_ring_fill(M, a.s, c->vbase + (uint32_t)c->n0 + POOL_RING);
        c->width = w;
        c->done = 0;
        __threadfence_block();
        c->taken = 0;
        __threadfence_block();
        c->state = SLOT_OPEN;
    }
    __syncwarp();
}

// chain epilogue (closer warp): (log2 highValue, highIndex) fs:303, Array.sum fs:445, statistics
__device__ __forceinline__ void pool_finish_chain(const SlotMem &M, const ChainArgs &a, int lane) {
    SlotCtl *c = M.ctl;
    const int N = a.s.n;
    const size_t off = (size_t)c->chain * N;
    for (int n = lane; n < N; n += 32) {
        const double v = __ldcg(a.hv + off + n);
        if (v == v) a.scores[off + n] = log2_ref(v);
    }
    __syncwarp();
    if (lane == 0) {
        double sum = 0.0;
        for (int n = 0; n < N; ++n) sum = __dadd_rn(sum, __ldcg(a.scores + off + n));
        a.sums[c->chain] = sum;
        atomicAdd(a.stats + ST_SITE_UPDATES, c->st_updates);
        atomicAdd(a.stats + ST_WINDOW_SCORES, c->st_windows);
        atomicAdd(a.stats + ST_EXACT_RESCANS, c->st_slow);
        atomicAdd(a.stats + ST_SPECULATED, c->st_spec);
        atomicAdd(a.stats + ST_SWEEPS, (unsigned long long)c->st_sweeps);
        atomicAdd(a.stats + ST_CAPPED, (unsigned long long)c->capped);
    }
    __syncwarp();
}

template <int KP>
__global__ void __launch_bounds__(POOL_WARPS * 32, 1) pool_kernel(const PoolArgs pa) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const ChainArgs &a = pa.c;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = a.s.n, k = a.k, row_words = a.s.row_words;
    uint32_t *lut = reinterpret_cast<uint32_t *>(smem_raw);
    int *finished = reinterpret_cast<int *>(smem_raw + 64);
    WarpTables WT;
    {
        unsigned char *b = smem_raw + 128 + warp * WARP_TABLE_BYTES;
        WT.wcol = reinterpret_cast<double *>(b);
        WT.ptab = reinterpret_cast<int32_t *>(b + 1024);
        WT.lgcol = reinterpret_cast<int32_t *>(b + 2048);
        WT.counts = reinterpret_cast<int32_t *>(b + 2560);
    }

    // ---- set-up: slot j of CTA b starts with chain b + j * gridDim.x ----
    if (tid < 16) lut[tid] = hist_lut_entry(tid);
    if (tid == 0) *finished = 0;
    __syncthreads();
    if (warp < POOL_SLOTS) {
        const SlotMem M = slot_mem(smem_raw, row_words, warp);
        const int chain = blockIdx.x + warp * gridDim.x;
        if (lane == 0) {
            for (int i = 0; i < POOL_RING; ++i) mbar_init(M.bar + i, 1);
            fence_barrier_init();
            M.ctl->state = SLOT_CLOSED;
            M.ctl->vbase = 0;
            M.ctl->issued = 0;
            M.ctl->next_seq = 0;
            M.ctl->taken = 0;
            M.ctl->done = 0;
            M.ctl->width = 0;
        }
        __syncwarp();
        if (chain < a.n_chains) {
            pool_bind_chain(M, a, chain, lane);
            if (M.ctl->phase == PH_DONE) { // nothing to run
                pool_finish_chain(M, a, lane);
                if (lane == 0) {
                    M.ctl->state = SLOT_FINISHED;
                    atomicAdd(finished, 1);
                }
            } else {
                pool_start_sweep<KP>(M, a, lut, lane);
            }
        } else if (lane == 0) {
            M.ctl->state = SLOT_FINISHED;
            atomicAdd(finished, 1);
        }
    }
    __syncthreads();
    if (warp < POOL_SLOTS) {
        const SlotMem M = slot_mem(smem_raw, row_words, warp);
        if (M.ctl->state == SLOT_CLOSED) pool_open_round(M, a, POOL_SLOTS - *finished, lane);
    }
    __syncthreads();

    // ---- worker loop ----
    int rr = warp; // round-robin start
    for (;;) {
        // claim a work item
        int slot = -1, idx = 0;
        if (lane == 0) {
            for (int t = 0; t < POOL_SLOTS; ++t) {
                const int s = (rr + t) & (POOL_SLOTS - 1);
                SlotCtl *c = slot_mem(smem_raw, row_words, s).ctl;
                if (*(volatile int *)&c->state == SLOT_OPEN && *(volatile int *)&c->taken < *(volatile int *)&c->width) {
                    const int i = atomicAdd(&c->taken, 1);
                    if (i < *(volatile int *)&c->width) {
                        slot = s;
                        idx = i;
                        break;
                    }
                }
            }
            if (slot < 0 && *(volatile int *)finished >= POOL_SLOTS) slot = -2;
        }
        slot = __shfl_sync(FULL, slot, 0);
        idx = __shfl_sync(FULL, idx, 0);
        if (slot == -2) break;
        if (slot < 0) {
            __nanosleep(40);
            continue;
        }
        rr = slot + 1;
        __threadfence_block();
        const SlotMem M = slot_mem(smem_raw, row_words, slot);
        SlotCtl *c = M.ctl;
        const int phase = *(volatile int *)&c->phase;
        const int n = *(volatile int *)&c->n0 + idx;
        const int chain = *(volatile int *)&c->chain;
        const uint32_t vbase = *(volatile uint32_t *)&c->vbase;
        const int mode = phase == PH_LEFT ? SHIFT_LEFT : phase == PH_RIGHT ? SHIFT_RIGHT : SHIFT_NONE;

        // ---- one site update (same arithmetic as chain_kernel) ----
        const uint32_t *row = pool_ring_wait(M, row_words, vbase + (uint32_t)n);
        const int o = ((n >> 5) & 1) * 32 + (n & 31);
        const int len_n = M.blk_len[o];
        const int Wn = len_n - k + 1;
        int site_n = 0, w = 0;
        double hv_n = 0.0, p = 0.0;
        uint64_t own = 0, neu = 0;
        if (phase == PH_INIT) {
            random_loo_counts<KP>(a, (uint64_t)a.chain_id_base + (uint64_t)chain, chain, n, WT.counts, lut, lane);
            build_tables<KP>(WT, WT.counts, false, 0, k, a.wtab, lane);
        } else {
            site_n = M.blk_site[o];
            hv_n = M.blk_hv[o];
            own = kmer_shared<KP>(row, shifted_site(site_n, len_n, k, mode));
            build_tables<KP>(WT, M.total, true, own, k, a.wtab, lane);
        }
        const bool slow = pick_argmax<KP>(WT, row, Wn, k, a.fast_ok, lane, p, w);
        bool accept = true, moved = false;
        if (phase != PH_INIT) {
            accept = score_improves(p, hv_n, hv_n != hv_n ? __ldcg(a.scores + (size_t)chain * N + n) : 0.0); // fs:402
            moved = accept && (w != site_n);
            if (moved && phase == PH_GREEDY) neu = kmer_shared<KP>(row, w);
        }
        int closer = 0;
        if (lane == 0) {
            SlotRes r;
            r.p = p;
            r.own = own;
            r.neu = neu;
            r.w = w;
            r.flag = (accept ? 1 : 0) | (moved ? 2 : 0) | (slow ? 4 : 0);
            r.Wn = Wn;
            r.pad = 0;
            M.res[idx] = r;
            __threadfence_block();
            closer = (atomicAdd(&c->done, 1) + 1 == *(volatile int *)&c->width) ? 1 : 0;
        }
        closer = __shfl_sync(FULL, closer, 0);
        if (!closer) continue;

        // ---- close the round (this warp only; the slot is invisible to claimers: taken >= width) ----
        __threadfence_block();
        if (lane == 0) c->state = SLOT_CLOSED;
        const int width = c->width;
        int flag = 0;
        if (lane < width) flag = M.res[lane].flag;
        const unsigned movers = __ballot_sync(FULL, (flag & 2) != 0);
        int first_mover = POOL_MAXW + 1; // greedy only: later items of the round saw stale counts
        if (phase == PH_GREEDY && movers) first_mover = __ffs(movers) - 1;
        const int last_commit = min(first_mover, width - 1);
        const bool committed = lane <= last_commit;
        if (committed && (flag & 1)) { // lane i commits item i
            const size_t off = (size_t)chain * N + (c->n0 + lane);
            a.sites[off] = M.res[lane].w;
            a.hv[off] = M.res[lane].p;
        }
        const int n_commit = last_commit + 1;
        const unsigned cm = n_commit >= 32 ? FULL : ((1u << n_commit) - 1u);
        const int wsum = __reduce_add_sync(FULL, committed ? M.res[lane].Wn : 0);
        const int nslow = __popc(__ballot_sync(FULL, committed && (flag & 4)));
        const bool any_moved = (movers & cm) != 0;
        if (phase == PH_GREEDY && first_mover < width) { // in-place sweep: -old k-mer, +new k-mer (fs:388)
            const uint64_t o_k = M.res[first_mover].own, n_k = M.res[first_mover].neu;
            if (lane < k) {
                const int bo = (int)((o_k >> (2 * lane)) & 3u), bn = (int)((n_k >> (2 * lane)) & 3u);
                if (bo != bn) {
                    M.total[lane * 4 + bo] -= 1;
                    M.total[lane * 4 + bn] += 1;
                }
            }
        }
        __syncwarp();
        if (lane == 0) {
            c->st_updates += (unsigned long long)n_commit;
            c->st_windows += (unsigned long long)wsum;
            c->st_slow += (unsigned long long)nslow;
            c->st_spec += (unsigned long long)(width - n_commit);
            c->changed |= any_moved ? 1 : 0;
            c->n0 += n_commit;
            if (phase == PH_GREEDY) c->gwidth = (first_mover < width) ? max(1, c->gwidth >> 1) : min(POOL_MAXW, c->gwidth * 2);
        }
        __syncwarp();
        bool chain_done = false;
        if (c->n0 >= N) { // ---- sweep finished: advance the phase state machine ----
            if (lane == 0) {
                c->vbase += (uint32_t)N;
                c->st_sweeps += 1;
                if (phase == PH_INIT) {
                    c->phase = next_phase(PH_GREEDY, a.phase_mask);
                    c->sweeps_in_phase = 0;
                } else {
                    c->sweeps_in_phase += 1;
                    bool next = !c->changed; // positions(acc) = positions(bestMotif), fs:384
                    if (!next && c->sweeps_in_phase >= a.max_sweeps) {
                        next = true;
                        c->capped = 1;
                    }
                    if (next) {
                        c->sweeps_in_phase = 0;
                        c->phase = next_phase(phase + 1, a.phase_mask);
                    }
                }
            }
            __syncwarp();
            __threadfence(); // the committed sites / scores are read back by site_counts and block loads
            if (c->phase == PH_DONE) {
                pool_finish_chain(M, a, lane);
                int next_chain = 0;
                if (lane == 0) next_chain = atomicAdd(pa.next_chain, 1);
                next_chain = __shfl_sync(FULL, next_chain, 0);
                if (next_chain < a.n_chains) {
                    pool_bind_chain(M, a, next_chain, lane);
                    if (c->phase == PH_DONE) chain_done = true; // empty pipeline: treat as finished below
                } else {
                    chain_done = true;
                }
                if (chain_done) {
                    if (lane == 0) { // let the prefetched rows land before the slot dies
                        for (uint32_t v = c->vbase; v != c->issued; ++v) pool_ring_wait(M, row_words, v);
                        c->state = SLOT_FINISHED;
                        __threadfence_block();
                        atomicAdd(finished, 1);
                    }
                    __syncwarp();
                    continue;
                }
            }
            pool_start_sweep<KP>(M, a, lut, lane);
        }
        pool_open_round(M, a, POOL_SLOTS - *(volatile int *)finished, lane);
    }
}

} // namespace gibbs
