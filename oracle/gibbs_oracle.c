/*
 * oracle/gibbs_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement, in plain C, of the hot path of Etschbeijer/GibbsSampling
 * (/root/reference/GibbsSampling/GibbsSampling.fs, cited as fs:N). It follows the reference
 * function by function, including its quirks: 49-slot vectors and
 * 49 x k matrices indexed by ASCII-42, from-scratch leave-one-out rebuilds, a PWM rebuilt for
 * every window, float64 left-to-right products, log2 x = ln x / ln 2, the drifting in-place
 * background of getBestPWMSs, first-strict-maximum argmax, sequential roulette accumulation and
 * the promote-or-restart loop.
 *
 * PARITY UNPINNED: the reference has no tests or golden vectors and cannot run here (no .NET).
 * See gibbs_oracle.h. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product path never does.
 *
 * Build: see oracle/Makefile (-O2 -ffp-contract=off: no FMA contraction, IEEE float64).
 */
#include "gibbs_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define NS OR_NSLOT

/* ------------------------------------------------------------------------------------------ */
/* context                                                                                    */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    const uint8_t *seqs;
    const int64_t *off;
    int32_t n;
    int32_t k;
    double pc;
    const uint8_t *alpha;
    int32_t alen;
} ctx_t;

static inline const uint8_t *seq_of(const ctx_t *c, int32_t i) { return c->seqs + c->off[i]; }
static inline int32_t len_of(const ctx_t *c, int32_t i) { return (int32_t)(c->off[i + 1] - c->off[i]); }
static inline int slot(uint8_t s) { return (int)s - 42; } /* fs:17, fs:176 */

static int check_symbols(const uint8_t *s, int64_t n) {
    for (int64_t i = 0; i < n; ++i)
        if (s[i] < 42 || s[i] > 90) return OR_ERR_SYMBOL; /* 49-slot array, fs:20 */
    return OR_OK;
}

static int check_ctx(const ctx_t *c) {
    if (!c->seqs || !c->off || !c->alpha || c->n < 1 || c->k < 1 || c->alen < 1) return OR_ERR_ARG;
    for (int32_t i = 0; i < c->n; ++i)
        if (len_of(c, i) < c->k) return OR_ERR_SHORT_SEQ; /* Array.take, fs:152 / rnd.Next, fs:145 */
    int rc = check_symbols(c->seqs + c->off[0], c->off[c->n] - c->off[0]);
    if (rc) return rc;
    return check_symbols(c->alpha, c->alen);
}

static double log2_ref(double x) { return log(x) / log(2.0); } /* FSharpAux log2 = Math.Log(x, 2.0) */

/* ------------------------------------------------------------------------------------------ */
/* RNG: injected stream or Philox4x32-10                                                      */
/* ------------------------------------------------------------------------------------------ */
void or_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* Draw `draw` of stream (seed, chain): Philox block draw/4 -> word draw%4 -> u = word * 2^-32. */
double or_uniform_at(uint64_t seed, uint64_t chain, uint64_t draw) {
    uint64_t blk = draw >> 2;
    uint32_t ctr[4] = {(uint32_t)blk, (uint32_t)(blk >> 32), (uint32_t)chain, (uint32_t)(chain >> 32)};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t out[4];
    or_philox4x32_10(ctr, key, out);
    return (double)out[draw & 3] * (1.0 / 4294967296.0);
}

double or_next_uniform(or_rng *rng) {
    double u;
    if (rng->mode == 0) {
        if (rng->next >= rng->n_u) { rng->exhausted = 1; u = 0.0; }
        else u = rng->u[rng->next];
    } else {
        u = or_uniform_at(rng->seed, rng->chain, (uint64_t)rng->next);
    }
    rng->next++;
    return u;
}

/* fs:143-146: rnd.Next(0, L-k+1); .NET maps Sample() * range -> int (truncation). */
int32_t or_draw_to_position(double u, int32_t len, int32_t k) {
    return (int32_t)(u * (double)(len - k + 1));
}

/* ------------------------------------------------------------------------------------------ */
/* CompositeVector (fs:11-124)                                                                */
/* ------------------------------------------------------------------------------------------ */
static void fcv_zero(int32_t *v) { memset(v, 0, NS * sizeof(int32_t)); }

/* fs:79-81 (and fs:60-62 when v starts at zero): in-place +1 per symbol */
static void fcv_increase_in_place_of(const uint8_t *s, int32_t len, int32_t *v) {
    for (int32_t i = 0; i < len; ++i) v[slot(s[i])] += 1;
}

/* fs:73-76: counts of source[0..pos-1] ++ source[pos+k..] */
static void fcv_create_without(int32_t k, int32_t pos, const uint8_t *s, int32_t len, int32_t *v) {
    fcv_zero(v);
    /* Array.append allocates the concatenation (L-k symbols) before folding */
    int32_t m = len - k;
    uint8_t *tmp = (uint8_t *)malloc((size_t)(m > 0 ? m : 1));
    memcpy(tmp, s, (size_t)pos);
    memcpy(tmp + pos, s + pos + k, (size_t)(len - pos - k));
    fcv_increase_in_place_of(tmp, m, v);
    free(tmp);
}

/* fs:65-70: only alphabet slots are summed (quirk A.6-10) */
static void fcv_fuse_add(const ctx_t *c, const int32_t *src, int32_t *dst) {
    for (int32_t a = 0; a < c->alen; ++a) dst[slot(c->alpha[a])] += src[slot(c->alpha[a])];
}

/* fs:84-88: same array, max(c-1, 0) per symbol of the segment */
static void fcv_subtract_segment(const uint8_t *seg, int32_t k, int32_t *v) {
    for (int32_t j = 0; j < k; ++j) {
        int s = slot(seg[j]);
        v[s] = (v[s] - 1 > 0) ? v[s] - 1 : 0;
    }
}

/* fs:115-120: sum over all 49 slots; alphabet slots normalised, others keep the raw float count */
static void pcv_normalized_of_fcv(const ctx_t *c, const int32_t *fcv, double *pcv) {
    int32_t isum = 0;
    for (int s = 0; s < NS; ++s) { pcv[s] = (double)fcv[s]; isum += fcv[s]; }
    double sum = (double)isum + ((double)c->alen * c->pc);
    for (int32_t a = 0; a < c->alen; ++a) {
        int s = slot(c->alpha[a]);
        pcv[s] = (pcv[s] + c->pc) / sum;
    }
}

/* fs:123-124 */
static double pcv_segment_score(const double *pcv, const uint8_t *seg, int32_t k) {
    double v = 1.0;
    for (int32_t j = 0; j < k; ++j) v = v * pcv[slot(seg[j])];
    return v;
}

/* ------------------------------------------------------------------------------------------ */
/* PositionMatrix (fs:126-293)                                                                */
/* ------------------------------------------------------------------------------------------ */
/* fs:149-153: Array.skip start |> Array.take k  (skip copies the tail, take copies k) */
static void get_segment(const uint8_t *s, int32_t len, int32_t start, int32_t k, uint8_t *tail_buf,
                        uint8_t *seg) {
    memcpy(tail_buf, s + start, (size_t)(len - start)); /* Array.skip allocation */
    memcpy(seg, tail_buf, (size_t)k);
}

/* fs:211-215: one-hot 49 x k */
static void pfm_create_of(const uint8_t *seg, int32_t k, int32_t *pfm) {
    memset(pfm, 0, (size_t)NS * k * sizeof(int32_t));
    for (int32_t j = 0; j < k; ++j) pfm[slot(seg[j]) * k + j] += 1;
}

/* fs:218-226: dst += src over all 49 x k cells */
static void pfm_fuse_add(const int32_t *src, int32_t k, int32_t *dst) {
    for (int i = 0; i < NS * k; ++i) dst[i] += src[i];
}

/* fs:249-261: createPPMOf then normalizePPM sourceCount alphabet pc (alphabet rows only) */
static void ppm_of_pfm(const ctx_t *c, const int32_t *pfm, int32_t source_count, double *ppm) {
    int32_t k = c->k;
    for (int i = 0; i < NS * k; ++i) ppm[i] = (double)pfm[i];
    double sum = (double)source_count + ((double)c->alen * c->pc);
    for (int32_t a = 0; a < c->alen; ++a) {
        int s = slot(c->alpha[a]);
        for (int32_t j = 0; j < k; ++j) ppm[s * k + j] = (ppm[s * k + j] + c->pc) / sum;
    }
}

/* fs:282-287: PWM[s,j] = PPM[s,j] / pcv[s] for alphabet rows, 0 elsewhere */
static void pwm_create(const ctx_t *c, const double *pcv, const double *ppm, double *pwm) {
    int32_t k = c->k;
    memset(pwm, 0, (size_t)NS * k * sizeof(double));
    for (int32_t a = 0; a < c->alen; ++a) {
        int s = slot(c->alpha[a]);
        for (int32_t j = 0; j < k; ++j) pwm[s * k + j] = ppm[s * k + j] / pcv[s];
    }
}

/* fs:290-293 */
static double pwm_segment_score(const double *pwm, const uint8_t *seg, int32_t k) {
    double v = 1.0;
    for (int32_t j = 0; j < k; ++j) v = v * pwm[slot(seg[j]) * k + j];
    return v;
}

/* ------------------------------------------------------------------------------------------ */
/* leave-one-out builders shared by the sweeps                                                */
/* ------------------------------------------------------------------------------------------ */
/* getSegment -> createPFMOf -> fuse for a list of (sequence, position) pairs (fs:392-396 etc.) */
typedef struct {
    int32_t *pfm_one; /* 49*k scratch: one-hot of one site */
    int32_t *pfm;     /* 49*k fused                          */
    int32_t *fcv_one; /* 49                                   */
    uint8_t *tail;    /* max_len scratch for Array.skip       */
    uint8_t *seg;     /* k                                    */
    double *ppm;      /* 49*k                                 */
    double *pwm;      /* 49*k                                 */
    int32_t max_len;
} scratch_t;

static int scratch_init(scratch_t *s, const ctx_t *c) {
    int32_t ml = 0;
    for (int32_t i = 0; i < c->n; ++i) if (len_of(c, i) > ml) ml = len_of(c, i);
    s->max_len = ml;
    size_t cells = (size_t)NS * c->k;
    s->pfm_one = (int32_t *)malloc(cells * sizeof(int32_t));
    s->pfm = (int32_t *)malloc(cells * sizeof(int32_t));
    s->fcv_one = (int32_t *)malloc(NS * sizeof(int32_t));
    s->tail = (uint8_t *)malloc((size_t)ml + 1);
    s->seg = (uint8_t *)malloc((size_t)c->k);
    s->ppm = (double *)malloc(cells * sizeof(double));
    s->pwm = (double *)malloc(cells * sizeof(double));
    if (!s->pfm_one || !s->pfm || !s->fcv_one || !s->tail || !s->seg || !s->ppm || !s->pwm) return OR_ERR_NOMEM;
    return OR_OK;
}

static void scratch_free(scratch_t *s) {
    free(s->pfm_one); free(s->pfm); free(s->fcv_one); free(s->tail); free(s->seg); free(s->ppm); free(s->pwm);
}

static void pfm_begin(const ctx_t *c, scratch_t *s) { memset(s->pfm, 0, (size_t)NS * c->k * sizeof(int32_t)); }

static void pfm_add_site(const ctx_t *c, scratch_t *s, int32_t i, int32_t pos) {
    get_segment(seq_of(c, i), len_of(c, i), pos, c->k, s->tail, s->seg);
    pfm_create_of(s->seg, c->k, s->pfm_one);
    pfm_fuse_add(s->pfm_one, c->k, s->pfm);
}

static void fcv_add_without(const ctx_t *c, scratch_t *s, int32_t i, int32_t pos, int32_t *fcv) {
    fcv_create_without(c->k, pos, seq_of(c, i), len_of(c, i), s->fcv_one);
    fcv_fuse_add(c, s->fcv_one, fcv);
}

/* ------------------------------------------------------------------------------------------ */
/* scans                                                                                      */
/* ------------------------------------------------------------------------------------------ */
/* fs:301-314 */
static void best_pwms_with_bpv(const ctx_t *c, scratch_t *s, const uint8_t *src, int32_t len,
                               const double *pcv, const double *ppm, double *score_out,
                               int32_t *pos_out, or_stats *st) {
    double high = 0.0;
    int32_t hi = 0;
    int32_t k = c->k;
    for (int32_t n = 0; n + k <= len; ++n) {
        get_segment(src, len, n, k, s->tail, s->seg);   /* Array.skip n |> Array.take k */
        pwm_create(c, pcv, ppm, s->pwm);                /* rebuilt per window, fs:309  */
        double tmp = pwm_segment_score(s->pwm, s->seg, k);
        if (tmp > high) { high = tmp; hi = n; }
        if (st) st->window_scores++;
    }
    if (st) st->site_updates++;
    *score_out = log2_ref(high);
    *pos_out = hi;
}

/* fs:462-479: drifting background (quirk A.6-1). fcv is mutated in place. */
static void best_pwms_drifting(const ctx_t *c, scratch_t *s, const uint8_t *src, int32_t len,
                               int32_t *fcv, const double *ppm, double *score_out, int32_t *pos_out,
                               double *raw_out, or_stats *st) {
    double high = 0.0;
    int32_t hi = 0;
    int32_t k = c->k;
    double pcv[NS];
    for (int32_t n = 0; n + k <= len; ++n) {
        get_segment(src, len, n, k, s->tail, s->seg);
        fcv_increase_in_place_of(src, len, fcv);        /* fs:471, in place */
        fcv_subtract_segment(s->seg, k, fcv);           /* fs:472, same array (fs:85) */
        pcv_normalized_of_fcv(c, fcv, pcv);             /* fs:473 */
        pwm_create(c, pcv, ppm, s->pwm);                /* fs:474 */
        double tmp = pwm_segment_score(s->pwm, s->seg, k);
        if (raw_out) raw_out[n] = tmp;
        if (tmp > high) { high = tmp; hi = n; }
        if (st) st->window_scores++;
    }
    if (st) st->site_updates++;
    *score_out = log2_ref(high);
    *pos_out = hi;
}

/* ------------------------------------------------------------------------------------------ */
/* exported primitives                                                                        */
/* ------------------------------------------------------------------------------------------ */
#define MAKE_CTX(c)                                                                            \
    ctx_t c;                                                                                   \
    c.seqs = seqs; c.off = off; c.n = n_seqs; c.k = k; c.pc = pc; c.alpha = alphabet; c.alen = alen

int or_loo_pfm(const uint8_t *seqs, const int64_t *off, int32_t n_seqs, const int32_t *sites,
               int32_t heldout, int32_t k, int32_t *pfm_out) {
    static const uint8_t dummy_alpha[1] = {'A'};
    const uint8_t *alphabet = dummy_alpha; int32_t alen = 1; double pc = 0.0;
    MAKE_CTX(c);
    int rc = check_ctx(&c); if (rc) return rc;
    if (!sites || !pfm_out || heldout < 0 || heldout >= n_seqs) return OR_ERR_ARG;
    for (int32_t i = 0; i < n_seqs; ++i)
        if (i != heldout && (sites[i] < 0 || sites[i] + k > len_of(&c, i))) return OR_ERR_ARG;
    scratch_t s; rc = scratch_init(&s, &c); if (rc) { scratch_free(&s); return rc; }
    pfm_begin(&c, &s);
    for (int32_t i = 0; i < n_seqs; ++i) if (i != heldout) pfm_add_site(&c, &s, i, sites[i]);
    memcpy(pfm_out, s.pfm, (size_t)NS * k * sizeof(int32_t));
    scratch_free(&s);
    return OR_OK;
}

int or_ppm_of_pfm(const int32_t *pfm, int32_t k, int32_t source_count, const uint8_t *alphabet,
                  int32_t alen, double pc, double *ppm_out) {
    if (!pfm || !ppm_out || !alphabet || k < 1) return OR_ERR_ARG;
    if (check_symbols(alphabet, alen)) return OR_ERR_SYMBOL;
    ctx_t c; memset(&c, 0, sizeof c); c.k = k; c.pc = pc; c.alpha = alphabet; c.alen = alen;
    ppm_of_pfm(&c, pfm, source_count, ppm_out);
    return OR_OK;
}

int or_pcv_of_sources(const uint8_t *seqs, const int64_t *off, int32_t n_seqs,
                      const uint8_t *alphabet, int32_t alen, double pc, double *pcv_out) {
    int32_t k = 1;
    MAKE_CTX(c);
    if (!seqs || !off || !alphabet || !pcv_out || n_seqs < 1) return OR_ERR_ARG;
    int rc = check_symbols(seqs + off[0], off[n_seqs] - off[0]); if (rc) return rc;
    rc = check_symbols(alphabet, alen); if (rc) return rc;
    int32_t fused[NS], one[NS];
    fcv_zero(fused);
    for (int32_t i = 0; i < n_seqs; ++i) {
        fcv_zero(one);
        fcv_increase_in_place_of(seq_of(&c, i), len_of(&c, i), one); /* createFCVOf fs:60 */
        fcv_fuse_add(&c, one, fused);                                /* fs:65 */
    }
    pcv_normalized_of_fcv(&c, fused, pcv_out);                       /* fs:115 */
    return OR_OK;
}

static int single_seq_ctx(ctx_t *c, int64_t *off2, const uint8_t *src, int32_t len, int32_t k,
                          const uint8_t *alphabet, int32_t alen, double pc) {
    off2[0] = 0; off2[1] = len;
    c->seqs = src; c->off = off2; c->n = 1; c->k = k; c->pc = pc; c->alpha = alphabet; c->alen = alen;
    return check_ctx(c);
}

int or_window_scores_bpv(const uint8_t *src, int32_t len, int32_t k, const uint8_t *alphabet,
                         int32_t alen, const double *pcv, const double *ppm, double *scores_out) {
    ctx_t c; int64_t off2[2];
    int rc = single_seq_ctx(&c, off2, src, len, k, alphabet, alen, 0.0); if (rc) return rc;
    if (!pcv || !ppm || !scores_out) return OR_ERR_ARG;
    scratch_t s; rc = scratch_init(&s, &c); if (rc) { scratch_free(&s); return rc; }
    for (int32_t n = 0; n + k <= len; ++n) {
        get_segment(src, len, n, k, s.tail, s.seg);
        pwm_create(&c, pcv, ppm, s.pwm);
        scores_out[n] = pwm_segment_score(s.pwm, s.seg, k);
    }
    scratch_free(&s);
    return OR_OK;
}

int or_best_pwms_with_bpv(const uint8_t *src, int32_t len, int32_t k, const uint8_t *alphabet,
                          int32_t alen, const double *pcv, const double *ppm, double *score_out,
                          int32_t *pos_out) {
    ctx_t c; int64_t off2[2];
    int rc = single_seq_ctx(&c, off2, src, len, k, alphabet, alen, 0.0); if (rc) return rc;
    if (!pcv || !ppm || !score_out || !pos_out) return OR_ERR_ARG;
    scratch_t s; rc = scratch_init(&s, &c); if (rc) { scratch_free(&s); return rc; }
    best_pwms_with_bpv(&c, &s, src, len, pcv, ppm, score_out, pos_out, NULL);
    scratch_free(&s);
    return OR_OK;
}

int or_best_pwms(const uint8_t *src, int32_t len, int32_t k, const uint8_t *alphabet, int32_t alen,
                 double pc, int32_t *fcv, const double *ppm, double *score_out, int32_t *pos_out,
                 double *window_scores_out) {
    ctx_t c; int64_t off2[2];
    int rc = single_seq_ctx(&c, off2, src, len, k, alphabet, alen, pc); if (rc) return rc;
    if (!fcv || !ppm || !score_out || !pos_out) return OR_ERR_ARG;
    scratch_t s; rc = scratch_init(&s, &c); if (rc) { scratch_free(&s); return rc; }
    best_pwms_drifting(&c, &s, src, len, fcv, ppm, score_out, pos_out, window_scores_out, NULL);
    scratch_free(&s);
    return OR_OK;
}

int or_loo_fcv(const uint8_t *seqs, const int64_t *off, int32_t n_seqs, const int32_t *sites,
               int32_t heldout, int32_t k, const uint8_t *alphabet, int32_t alen, int32_t *fcv_out) {
    double pc = 0.0;
    MAKE_CTX(c);
    int rc = check_ctx(&c); if (rc) return rc;
    if (!sites || !fcv_out || heldout < 0 || heldout >= n_seqs) return OR_ERR_ARG;
    scratch_t s; rc = scratch_init(&s, &c); if (rc) { scratch_free(&s); return rc; }
    fcv_zero(fcv_out);
    for (int32_t i = 0; i < n_seqs; ++i) if (i != heldout) fcv_add_without(&c, &s, i, sites[i], fcv_out);
    scratch_free(&s);
    return OR_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* SiteSampler sweeps (fs:318-611)                                                            */
/* ------------------------------------------------------------------------------------------ */
enum { SHIFT_NONE = 0, SHIFT_LEFT = 1, SHIFT_RIGHT = 2 };

static int32_t shifted(const ctx_t *c, int32_t i, int32_t pos, int mode) {
    if (mode == SHIFT_LEFT) return pos > 0 ? pos - 1 : pos;                              /* fs:358 */
    if (mode == SHIFT_RIGHT) return pos <= len_of(c, i) - c->k - 1 ? pos + 1 : pos;      /* fs:327 */
    return pos;
}

static int positions_equal(const int32_t *a, const int32_t *b, int32_t n) {
    return memcmp(a, b, (size_t)n * sizeof(int32_t)) == 0;
}

/*
 * One family of sweeps covers fs:381 (greedy, BPV), fs:350 / fs:318 (shifts, BPV) and their
 * data-derived twins fs:554 / fs:519 / fs:483:
 *   greedy : other sites are read from acc (updated in place)
 *   shifts : other sites are read from the snapshot bestMotif, shifted by -1 / +1 (clamped)
 * acc[n] <- tmp iff fst tmp > fst acc[n]; repeat until positions(acc) = positions(snapshot).
 */
static int sweep_until_stable(const ctx_t *c, scratch_t *s, const double *pcv_or_null, int mode,
                              double *score, int32_t *pos, or_stats *st) {
    int32_t n = c->n;
    int32_t *snap_pos = (int32_t *)malloc((size_t)n * sizeof(int32_t));
    if (!snap_pos) return OR_ERR_NOMEM;
    memcpy(snap_pos, pos, (size_t)n * sizeof(int32_t)); /* bestMotif = Array.copy startPositions */
    int32_t fcv[NS];
    for (;;) {
        for (int32_t h = 0; h < n; ++h) {
            const int32_t *from = (mode == SHIFT_NONE) ? pos : snap_pos;
            pfm_begin(c, s);
            if (!pcv_or_null) fcv_zero(fcv);
            for (int32_t i = 0; i < n; ++i) {
                if (i == h) continue;
                int32_t p = shifted(c, i, from[i], mode);
                if (!pcv_or_null) fcv_add_without(c, s, i, p, fcv);
                pfm_add_site(c, s, i, p);
            }
            ppm_of_pfm(c, s->pfm, n - 1, s->ppm);
            double ts; int32_t tp;
            if (pcv_or_null) best_pwms_with_bpv(c, s, seq_of(c, h), len_of(c, h), pcv_or_null, s->ppm, &ts, &tp, st);
            else best_pwms_drifting(c, s, seq_of(c, h), len_of(c, h), fcv, s->ppm, &ts, &tp, NULL, st);
            if (ts > score[h]) { score[h] = ts; pos[h] = tp; }   /* fs:402 / fs:579 */
        }
        if (st) st->sweeps++;
        if (positions_equal(pos, snap_pos, n)) break;            /* fs:384 / fs:557 */
        memcpy(snap_pos, pos, (size_t)n * sizeof(int32_t));
    }
    free(snap_pos);
    return OR_OK;
}

/* fs:412-430 (pcv given), fs:589-611 (pcv NULL), fs:644-660 (fixed_ppm given, data background) */
static int random_starts(const ctx_t *c, scratch_t *s, const double *pcv_or_null,
                         const double *fixed_ppm_or_null, or_rng *rng, double *score, int32_t *pos,
                         or_stats *st) {
    int32_t n = c->n;
    int32_t *rp = (int32_t *)malloc((size_t)n * sizeof(int32_t));
    if (!rp) return OR_ERR_NOMEM;
    int32_t fcv[NS];
    for (int32_t h = 0; h < n; ++h) {
        for (int32_t i = 0; i < n; ++i) {            /* fresh draw for every other sequence (A.6-5) */
            if (i == h) continue;
            rp[i] = or_draw_to_position(or_next_uniform(rng), len_of(c, i), c->k);
        }
        if (!pcv_or_null) {
            fcv_zero(fcv);
            for (int32_t i = 0; i < n; ++i) if (i != h) fcv_add_without(c, s, i, rp[i], fcv);
        }
        const double *ppm = fixed_ppm_or_null;
        if (!ppm) {
            pfm_begin(c, s);
            for (int32_t i = 0; i < n; ++i) if (i != h) pfm_add_site(c, s, i, rp[i]);
            ppm_of_pfm(c, s->pfm, n - 1, s->ppm);
            ppm = s->ppm;
        }
        if (pcv_or_null) best_pwms_with_bpv(c, s, seq_of(c, h), len_of(c, h), pcv_or_null, ppm, &score[h], &pos[h], st);
        else best_pwms_drifting(c, s, seq_of(c, h), len_of(c, h), fcv, ppm, &score[h], &pos[h], NULL, st);
    }
    if (st) st->sweeps++;
    free(rp);
    return OR_OK;
}

/* pipeline variants: 0 = WithBPV fs:691, 1 = data-derived fs:697, 2 = WithPPM fs:703 */
static int site_pipeline(int variant, const ctx_t *c, scratch_t *s, const double *pcv,
                         const double *ppm, or_rng *rng, double *score, int32_t *pos, or_stats *st) {
    int rc;
    const double *bg = (variant == 0) ? pcv : NULL;
    rc = random_starts(c, s, bg, variant == 2 ? ppm : NULL, rng, score, pos, st); if (rc) return rc;
    rc = sweep_until_stable(c, s, bg, SHIFT_NONE, score, pos, st); if (rc) return rc;
    rc = sweep_until_stable(c, s, bg, SHIFT_LEFT, score, pos, st); if (rc) return rc;
    rc = sweep_until_stable(c, s, bg, SHIFT_RIGHT, score, pos, st); if (rc) return rc;
    if (st) st->restarts++;
    return OR_OK;
}

#define ENTER(need_pcv)                                                                        \
    MAKE_CTX(c);                                                                               \
    int rc = check_ctx(&c); if (rc) return rc;                                                 \
    if (!score || !pos) return OR_ERR_ARG;                                                     \
    if ((need_pcv) && !pcv) return OR_ERR_ARG;                                                 \
    scratch_t s; rc = scratch_init(&s, &c); if (rc) { scratch_free(&s); return rc; }
#define LEAVE scratch_free(&s); return rc

int or_random_starts_with_bpv(const uint8_t *seqs, const int64_t *off, int32_t n_seqs, int32_t k,
                              double pc, const uint8_t *alphabet, int32_t alen, const double *pcv,
                              or_rng *rng, double *score, int32_t *pos, or_stats *st) {
    ENTER(1); rc = random_starts(&c, &s, pcv, NULL, rng, score, pos, st); LEAVE;
}
int or_find_best_motif_with_start_position(const uint8_t *seqs, const int64_t *off, int32_t n_seqs,
                              int32_t k, double pc, const uint8_t *alphabet, int32_t alen,
                              const double *pcv, double *score, int32_t *pos, or_stats *st) {
    ENTER(1); rc = sweep_until_stable(&c, &s, pcv, SHIFT_NONE, score, pos, st); LEAVE;
}
int or_left_shifted_with_bpv(const uint8_t *seqs, const int64_t *off, int32_t n_seqs, int32_t k,
                              double pc, const uint8_t *alphabet, int32_t alen, const double *pcv,
                              double *score, int32_t *pos, or_stats *st) {
    ENTER(1); rc = sweep_until_stable(&c, &s, pcv, SHIFT_LEFT, score, pos, st); LEAVE;
}
int or_right_shifted_with_bpv(const uint8_t *seqs, const int64_t *off, int32_t n_seqs, int32_t k,
                              double pc, const uint8_t *alphabet, int32_t alen, const double *pcv,
                              double *score, int32_t *pos, or_stats *st) {
    ENTER(1); rc = sweep_until_stable(&c, &s, pcv, SHIFT_RIGHT, score, pos, st); LEAVE;
}
int or_do_site_sampling_with_bpv(const uint8_t *seqs, const int64_t *off, int32_t n_seqs, int32_t k,
                              double pc, const uint8_t *alphabet, int32_t alen, const double *pcv,
                              or_rng *rng, double *score, int32_t *pos, or_stats *st) {
    ENTER(1); rc = site_pipeline(0, &c, &s, pcv, NULL, rng, score, pos, st); LEAVE;
}
int or_random_starts(const uint8_t *seqs, const int64_t *off, int32_t n_seqs, int32_t k, double pc,
                     const uint8_t *alphabet, int32_t alen, or_rng *rng, double *score,
                     int32_t *pos, or_stats *st) {
    const double *pcv = NULL;
    ENTER(0); rc = random_starts(&c, &s, NULL, NULL, rng, score, pos, st); LEAVE;
}
int or_best_pwms_with_start_positions(const uint8_t *seqs, const int64_t *off, int32_t n_seqs,
                     int32_t k, double pc, const uint8_t *alphabet, int32_t alen, double *score,
                     int32_t *pos, or_stats *st) {
    const double *pcv = NULL;
    ENTER(0); rc = sweep_until_stable(&c, &s, NULL, SHIFT_NONE, score, pos, st); LEAVE;
}
int or_left_shifted(const uint8_t *seqs, const int64_t *off, int32_t n_seqs, int32_t k, double pc,
                     const uint8_t *alphabet, int32_t alen, double *score, int32_t *pos,
                     or_stats *st) {
    const double *pcv = NULL;
    ENTER(0); rc = sweep_until_stable(&c, &s, NULL, SHIFT_LEFT, score, pos, st); LEAVE;
}
int or_right_shifted(const uint8_t *seqs, const int64_t *off, int32_t n_seqs, int32_t k, double pc,
                     const uint8_t *alphabet, int32_t alen, double *score, int32_t *pos,
                     or_stats *st) {
    const double *pcv = NULL;
    ENTER(0); rc = sweep_until_stable(&c, &s, NULL, SHIFT_RIGHT, score, pos, st); LEAVE;
}
int or_do_site_sampling(const uint8_t *seqs, const int64_t *off, int32_t n_seqs, int32_t k,
                     double pc, const uint8_t *alphabet, int32_t alen, or_rng *rng, double *score,
                     int32_t *pos, or_stats *st) {
    const double *pcv = NULL;
    ENTER(0); rc = site_pipeline(1, &c, &s, NULL, NULL, rng, score, pos, st); LEAVE;
}
int or_motifs_with_best_pwms_of_ppm(const uint8_t *seqs, const int64_t *off, int32_t n_seqs,
                     int32_t k, double pc, const uint8_t *alphabet, int32_t alen, const double *ppm,
                     or_rng *rng, double *score, int32_t *pos, or_stats *st) {
    const double *pcv = NULL;
    if (!ppm) return OR_ERR_ARG;
    ENTER(0); rc = random_starts(&c, &s, NULL, ppm, rng, score, pos, st); LEAVE;
}
int or_do_site_sampling_with_ppm(const uint8_t *seqs, const int64_t *off, int32_t n_seqs, int32_t k,
                     double pc, const uint8_t *alphabet, int32_t alen, const double *ppm,
                     or_rng *rng, double *score, int32_t *pos, or_stats *st) {
    const double *pcv = NULL;
    if (!ppm) return OR_ERR_ARG;
    ENTER(0); rc = site_pipeline(2, &c, &s, NULL, ppm, rng, score, pos, st); LEAVE;
}

/* ------------------------------------------------------------------------------------------ */
/* restart loop (fs:434-459, fs:615-640, fs:664-689; quirk A.6-8)                             */
/* ------------------------------------------------------------------------------------------ */
static double seq_sum(const double *v, int32_t n) { /* Array.sum: left to right from 0.0 */
    double a = 0.0;
    for (int32_t i = 0; i < n; ++i) a = a + v[i];
    return a;
}

static int site_arrays_equal(const double *sa, const int32_t *pa, int32_t na, const double *sb,
                             const int32_t *pb, int32_t nb) {
    if (na != nb) return 0;
    for (int32_t i = 0; i < na; ++i)
        if (!(sa[i] == sb[i]) || pa[i] != pb[i]) return 0; /* F# (=) on float: NaN <> NaN */
    return 1;
}

int or_best_information_content(int32_t variant, int32_t reps, const uint8_t *seqs,
                     const int64_t *off, int32_t n_seqs, int32_t k, double pc,
                     const uint8_t *alphabet, int32_t alen, const double *pcv, const double *ppm,
                     or_rng *rng, double *score, int32_t *pos, int32_t *n_out, or_stats *st) {
    if (variant < 0 || variant > 2 || !n_out || !rng) return OR_ERR_ARG;
    if (variant == 2 && !ppm) return OR_ERR_ARG;
    ENTER(variant == 0);
    int32_t n = c.n;
    double *acc_s = (double *)malloc((size_t)n * sizeof(double));
    int32_t *acc_p = (int32_t *)malloc((size_t)n * sizeof(int32_t));
    double *best_s = (double *)malloc((size_t)(n > 1 ? n : 1) * sizeof(double));
    int32_t *best_p = (int32_t *)malloc((size_t)(n > 1 ? n : 1) * sizeof(int32_t));
    if (!acc_s || !acc_p || !best_s || !best_p) { free(acc_s); free(acc_p); free(best_s); free(best_p); rc = OR_ERR_NOMEM; LEAVE; }
    int32_t acc_n = 0, best_n = 1;           /* loop 0 [||] [|0., 0|] */
    best_s[0] = 0.0; best_p[0] = 0;
    uint64_t chain0 = rng->chain;
    int64_t restart_index = 0;
    for (int32_t it = 0;; ++it) {
        if (it > reps) break;                                                    /* fs:436 */
        if (site_arrays_equal(acc_s, acc_p, acc_n, best_s, best_p, best_n)) break; /* fs:439 */
        double ia = seq_sum(acc_s, acc_n), ib = seq_sum(best_s, best_n);
        if (ia > ib) {                                                           /* fs:450-451 */
            if (acc_n != 0) { memcpy(best_s, acc_s, (size_t)acc_n * sizeof(double)); memcpy(best_p, acc_p, (size_t)acc_n * sizeof(int32_t)); best_n = acc_n; }
            acc_n = 0;
        } else {
            if (rng->mode == 1) { rng->chain = chain0 + (uint64_t)restart_index; rng->next = 0; }
            restart_index++;
            rc = site_pipeline(variant, &c, &s, pcv, ppm, rng, acc_s, acc_p, st);
            if (rc) break;
            acc_n = n;
        }
    }
    rng->chain = chain0;
    if (!rc) { memcpy(score, best_s, (size_t)best_n * sizeof(double)); memcpy(pos, best_p, (size_t)best_n * sizeof(int32_t)); *n_out = best_n; }
    free(acc_s); free(acc_p); free(best_s); free(best_p);
    LEAVE;
}

/* ------------------------------------------------------------------------------------------ */
/* MotifSampler (fs:709-1038)                                                                 */
/* ------------------------------------------------------------------------------------------ */
typedef struct { double pwms; int32_t npos; int32_t pos[OR_MAX_M]; } midx_t;

typedef struct { midx_t *v; int64_t n, cap; int err; } mlist_t;

static void mlist_push(mlist_t *l, double pwms, const int32_t *pos, int32_t npos) {
    if (l->err) return;
    if (l->n == l->cap) {
        int64_t nc = l->cap ? l->cap * 2 : 256;
        midx_t *nv = (midx_t *)realloc(l->v, (size_t)nc * sizeof(midx_t));
        if (!nv) { l->err = OR_ERR_NOMEM; return; }
        l->v = nv; l->cap = nc;
    }
    midx_t *e = &l->v[l->n++];
    e->pwms = pwms; e->npos = npos;
    for (int32_t i = 0; i < npos; ++i) e->pos[i] = pos[i];
}

/* fs:129-140: every pair of list items must be more than `width` apart */
static int check_for_distance(int32_t width, const int32_t *items, int32_t n) {
    if (n <= 1) return 1;
    for (int32_t a = 0; a < n - 1; ++a)
        for (int32_t b = a + 1; b < n; ++b) {
            int32_t d = items[a] - items[b];
            if (d < 0) d = -d;
            if (!(d > width)) return 0;
        }
    return 1;
}

/* fs:727-742. `positions` is the cons-list (newest first). */
static void combos(double cutoff, int32_t width, const double *sc, const int32_t *ps, int32_t nset,
                   int32_t from, double prob, const int32_t *positions, int32_t npos, int32_t size,
                   mlist_t *out) {
    if (from < nset) {                                   /* | n, x::xs -> */
        if (size > 0) {
            int32_t np[OR_MAX_M + 1];
            np[0] = ps[from];
            for (int32_t i = 0; i < npos; ++i) np[i + 1] = positions[i];
            if (check_for_distance(width, np, npos + 1))
                if (log2_ref(sc[from] * prob) > cutoff)
                    combos(cutoff, width, sc, ps, nset, from + 1, sc[from] * prob, np, npos + 1, size - 1, out);
        }
        if (size >= 0) combos(cutoff, width, sc, ps, nset, from + 1, prob, positions, npos, size, out);
    } else if (size == 0) {                              /* | 0, [] -> */
        mlist_push(out, log2_ref(prob), positions, npos);
    }                                                    /* | _, [] -> () */
}

/* fs:759-784 */
static void normalized_segment_scores(const ctx_t *c, scratch_t *s, double cutoff, int32_t m,
                                      const uint8_t *src, int32_t len, const double *pcv,
                                      const double *pwm, mlist_t *out, or_stats *st) {
    int32_t k = c->k;
    int32_t w = len - k + 1;
    double *sc = (double *)malloc((size_t)w * sizeof(double));
    int32_t *ps = (int32_t *)malloc((size_t)w * sizeof(int32_t));
    if (!sc || !ps) { free(sc); free(ps); out->err = OR_ERR_NOMEM; return; }
    for (int32_t n = 0; n < w; ++n) {                    /* segments, fs:760-769 */
        get_segment(src, len, n, k, s->tail, s->seg);
        sc[n] = pwm_segment_score(pwm, s->seg, k);       /* fs:773 */
        ps[n] = n;
        if (st) st->window_scores++;
    }
    for (int32_t n = 0; n < w; ++n) {                    /* backGroundScores, fs:774-777 */
        get_segment(src, len, n, k, s->tail, s->seg);
        mlist_push(out, pcv_segment_score(pcv, s->seg, k), NULL, 0);
    }
    for (int32_t size = 1; size <= m; ++size)            /* fs:778-782 */
        combos(cutoff, k, sc, ps, w, 0, 1.0, NULL, 0, size, out);
    if (st) st->site_updates++;
    free(sc); free(ps);
}

int or_candidates(const uint8_t *src, int32_t len, int32_t k, int32_t m, double cutoff,
                  const uint8_t *alphabet, int32_t alen, const double *pcv, const double *ppm,
                  double *pwms_out, int32_t *npos_out, int32_t *pos_out, int64_t cap,
                  int64_t *n_out) {
    ctx_t c; int64_t off2[2];
    int rc = single_seq_ctx(&c, off2, src, len, k, alphabet, alen, 0.0); if (rc) return rc;
    if (!pcv || !ppm || !n_out || m < 1 || m > OR_MAX_M) return OR_ERR_ARG;
    scratch_t s; rc = scratch_init(&s, &c); if (rc) { scratch_free(&s); return rc; }
    pwm_create(&c, pcv, ppm, s.pwm);
    mlist_t l; memset(&l, 0, sizeof l);
    normalized_segment_scores(&c, &s, cutoff, m, src, len, pcv, s.pwm, &l, NULL);
    rc = l.err;
    if (!rc) {
        *n_out = l.n;
        for (int64_t i = 0; i < l.n && i < cap; ++i) {
            if (pwms_out) pwms_out[i] = l.v[i].pwms;
            if (npos_out) npos_out[i] = l.v[i].npos;
            if (pos_out) for (int32_t j = 0; j < OR_MAX_M; ++j) pos_out[i * OR_MAX_M + j] = j < l.v[i].npos ? l.v[i].pos[j] : -1;
        }
    }
    free(l.v);
    scratch_free(&s);
    return rc;
}

/* fs:746-754 */
static int roulette(const midx_t *items, int64_t n, double pick, int64_t *index_out) {
    double sum = 0.0;
    for (int64_t i = 0; i < n; ++i) sum = sum + items[i].pwms;     /* List.sum */
    double acc = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        double w = items[i].pwms / sum;
        if (acc <= pick && pick <= acc + w) { *index_out = i; return OR_OK; }
        acc = acc + w;
    }
    return OR_ERR_ROULETTE;                                        /* list index out of range */
}

int or_roulette(const double *pwms, int64_t n, double pick, int64_t *index_out) {
    if (!pwms || !index_out || n < 1) return OR_ERR_ARG;
    midx_t *it = (midx_t *)malloc((size_t)n * sizeof(midx_t));
    if (!it) return OR_ERR_NOMEM;
    for (int64_t i = 0; i < n; ++i) { it[i].pwms = pwms[i]; it[i].npos = 0; }
    int rc = roulette(it, n, pick, index_out);
    free(it);
    return rc;
}

/* List.sortByDescending PWMS |> List.head: stable => first maximum (quirk A.6-7) */
static int64_t first_max(const midx_t *items, int64_t n) {
    int64_t b = 0;
    for (int64_t i = 1; i < n; ++i) if (items[i].pwms > items[b].pwms) b = i;
    return b;
}

static int midx_positions_equal(const midx_t *a, const midx_t *b, int32_t n) {
    for (int32_t i = 0; i < n; ++i) {
        if (a[i].npos != b[i].npos) return 0;
        for (int32_t j = 0; j < a[i].npos; ++j) if (a[i].pos[j] != b[i].pos[j]) return 0;
    }
    return 1;
}

/* leave-one-out PPM / background from MotifIndex state `from` (fs:794-808, fs:891-915) */
static void motif_loo(const ctx_t *c, scratch_t *s, int variant, const midx_t *from, int32_t h,
                      const double *pcv_fixed, double *pcv_out) {
    pfm_begin(c, s);
    int32_t fcv[NS];
    fcv_zero(fcv);
    for (int32_t i = 0; i < c->n; ++i) {
        if (i == h) continue;
        for (int32_t j = 0; j < from[i].npos; ++j) {
            if (variant == 1) fcv_add_without(c, s, i, from[i].pos[j], fcv);    /* fs:897-903 */
            pfm_add_site(c, s, i, from[i].pos[j]);
        }
    }
    ppm_of_pfm(c, s->pfm, c->n - 1, s->ppm);                                    /* fs:808 / fs:915 */
    if (variant == 1) {
        fcv_increase_in_place_of(seq_of(c, h), len_of(c, h), fcv);              /* fs:904 */
        pcv_normalized_of_fcv(c, fcv, pcv_out);                                 /* fs:905 */
    } else {
        memcpy(pcv_out, pcv_fixed, NS * sizeof(double));
    }
    pwm_create(c, pcv_out, s->ppm, s->pwm);                                     /* fs:809 / fs:916 */
}

/* fs:788-822 (variant 0) / fs:885-929 (variant 1) */
static int motif_greedy(int variant, const ctx_t *c, scratch_t *s, int32_t m, double cutoff,
                        const double *pcv_fixed, midx_t *acc, or_stats *st) {
    int32_t n = c->n;
    midx_t *snap = (midx_t *)malloc((size_t)n * sizeof(midx_t));
    if (!snap) return OR_ERR_NOMEM;
    memcpy(snap, acc, (size_t)n * sizeof(midx_t));
    double pcv[NS];
    int rc = OR_OK;
    for (;;) {
        for (int32_t h = 0; h < n; ++h) {
            motif_loo(c, s, variant, acc, h, pcv_fixed, pcv);
            mlist_t l; memset(&l, 0, sizeof l);
            normalized_segment_scores(c, s, cutoff, m, seq_of(c, h), len_of(c, h), pcv, s->pwm, &l, st);
            if (l.err) { rc = l.err; free(l.v); goto done; }
            midx_t tmp = l.v[first_max(l.v, l.n)];
            free(l.v);
            if (tmp.pwms > acc[h].pwms) acc[h] = tmp;                           /* fs:816 / fs:923 */
        }
        if (st) st->sweeps++;
        if (midx_positions_equal(acc, snap, n)) break;
        memcpy(snap, acc, (size_t)n * sizeof(midx_t));
    }
done:
    free(snap);
    return rc;
}

/* fs:828-853 (variant 0) / fs:935-970 (variant 1): synchronous, every n reads the INPUT state */
static int motif_stochastic(int variant, const ctx_t *c, scratch_t *s, int32_t m, double cutoff,
                            const double *pcv_fixed, or_rng *rng, midx_t *state, or_stats *st) {
    int32_t n = c->n;
    midx_t *out = (midx_t *)malloc((size_t)n * sizeof(midx_t));
    if (!out) return OR_ERR_NOMEM;
    double pcv[NS];
    int rc = OR_OK;
    for (int32_t h = 0; h < n; ++h) {
        motif_loo(c, s, variant, state, h, pcv_fixed, pcv);
        mlist_t l; memset(&l, 0, sizeof l);
        normalized_segment_scores(c, s, cutoff, m, seq_of(c, h), len_of(c, h), pcv, s->pwm, &l, st);
        if (l.err) { rc = l.err; free(l.v); break; }
        int64_t idx;
        rc = roulette(l.v, l.n, or_next_uniform(rng), &idx);                    /* fs:851 / fs:968 */
        if (rc) { free(l.v); break; }
        out[h] = l.v[idx];
        free(l.v);
    }
    if (st) st->sweeps++;
    if (!rc) memcpy(state, out, (size_t)n * sizeof(midx_t));
    free(out);
    return rc;
}

static void midx_unpack(const midx_t *v, int32_t n, double *pwms, int32_t *npos, int32_t *pos) {
    for (int32_t i = 0; i < n; ++i) {
        pwms[i] = v[i].pwms; npos[i] = v[i].npos;
        for (int32_t j = 0; j < OR_MAX_M; ++j) pos[i * OR_MAX_M + j] = j < v[i].npos ? v[i].pos[j] : -1;
    }
}
static int midx_pack(midx_t *v, int32_t n, const double *pwms, const int32_t *npos, const int32_t *pos) {
    for (int32_t i = 0; i < n; ++i) {
        if (npos[i] < 0 || npos[i] > OR_MAX_M) return OR_ERR_ARG;
        v[i].pwms = pwms[i]; v[i].npos = npos[i];
        for (int32_t j = 0; j < npos[i]; ++j) v[i].pos[j] = pos[i * OR_MAX_M + j];
    }
    return OR_OK;
}

#define MENTER                                                                                 \
    MAKE_CTX(c);                                                                               \
    int rc = check_ctx(&c); if (rc) return rc;                                                 \
    if (!pwms || !npos || !pos || m < 1 || m > OR_MAX_M) return OR_ERR_ARG;                    \
    if (variant == 0 && !pcv) return OR_ERR_ARG;                                               \
    scratch_t s; rc = scratch_init(&s, &c); if (rc) { scratch_free(&s); return rc; }           \
    midx_t *state = (midx_t *)malloc((size_t)c.n * sizeof(midx_t));                            \
    if (!state) { scratch_free(&s); return OR_ERR_NOMEM; }
#define MLEAVE free(state); scratch_free(&s); return rc

int or_motif_greedy(int32_t variant, const uint8_t *seqs, const int64_t *off, int32_t n_seqs,
                    int32_t m, int32_t k, double pc, double cutoff, const uint8_t *alphabet,
                    int32_t alen, const double *pcv, double *pwms, int32_t *npos, int32_t *pos,
                    or_stats *st) {
    MENTER;
    rc = midx_pack(state, c.n, pwms, npos, pos);
    if (!rc) rc = motif_greedy(variant, &c, &s, m, cutoff, pcv, state, st);
    if (!rc) midx_unpack(state, c.n, pwms, npos, pos);
    MLEAVE;
}

int or_motif_stochastic(int32_t variant, const uint8_t *seqs, const int64_t *off, int32_t n_seqs,
                    int32_t m, int32_t k, double pc, double cutoff, const uint8_t *alphabet,
                    int32_t alen, const double *pcv, or_rng *rng, double *pwms, int32_t *npos,
                    int32_t *pos, or_stats *st) {
    MENTER;
    rc = midx_pack(state, c.n, pwms, npos, pos);
    if (!rc) rc = motif_stochastic(variant, &c, &s, m, cutoff, pcv, rng, state, st);
    if (!rc) midx_unpack(state, c.n, pwms, npos, pos);
    MLEAVE;
}

/* variant 0: fs:876-879, 1: fs:1034-1038, 2: fs:1028-1032 */
static int motif_pipeline(int variant, const ctx_t *c, scratch_t *s, int32_t m, double cutoff,
                          const double *pcv, const double *ppm, or_rng *rng, midx_t *state,
                          or_stats *st) {
    int32_t n = c->n;
    double *sc = (double *)malloc((size_t)n * sizeof(double));
    int32_t *ps = (int32_t *)malloc((size_t)n * sizeof(int32_t));
    if (!sc || !ps) { free(sc); free(ps); return OR_ERR_NOMEM; }
    int rc = random_starts(c, s, variant == 0 ? pcv : NULL, variant == 2 ? ppm : NULL, rng, sc, ps, st);
    if (!rc) {
        for (int32_t i = 0; i < n; ++i) { state[i].pwms = sc[i]; state[i].npos = 1; state[i].pos[0] = ps[i]; }
        int sub = variant == 0 ? 0 : 1;
        rc = motif_stochastic(sub, c, s, m, cutoff, pcv, rng, state, st);
        if (!rc) rc = motif_greedy(sub, c, s, m, cutoff, pcv, state, st);
    }
    if (!rc && st) st->restarts++;
    free(sc); free(ps);
    return rc;
}

int or_do_motif_sampling(int32_t variant, const uint8_t *seqs, const int64_t *off, int32_t n_seqs,
                    int32_t m, int32_t k, double pc, double cutoff, const uint8_t *alphabet,
                    int32_t alen, const double *pcv, const double *ppm, or_rng *rng, double *pwms,
                    int32_t *npos, int32_t *pos, or_stats *st) {
    if (variant < 0 || variant > 2 || !rng || (variant == 2 && !ppm)) return OR_ERR_ARG;
    MENTER;
    rc = motif_pipeline(variant, &c, &s, m, cutoff, pcv, ppm, rng, state, st);
    if (!rc) midx_unpack(state, c.n, pwms, npos, pos);
    MLEAVE;
}

static int midx_arrays_equal(const midx_t *a, int32_t na, const midx_t *b, int32_t nb) {
    if (na != nb) return 0;
    for (int32_t i = 0; i < na; ++i) {
        if (!(a[i].pwms == b[i].pwms) || a[i].npos != b[i].npos) return 0;
        for (int32_t j = 0; j < a[i].npos; ++j) if (a[i].pos[j] != b[i].pos[j]) return 0;
    }
    return 1;
}

int or_best_motif_information_content(int32_t variant, int32_t reps, const uint8_t *seqs,
                    const int64_t *off, int32_t n_seqs, int32_t m, int32_t k, double pc,
                    double cutoff, const uint8_t *alphabet, int32_t alen, const double *pcv,
                    const double *ppm, or_rng *rng, double *pwms, int32_t *npos, int32_t *pos,
                    int32_t *n_out, or_stats *st) {
    if (variant < 0 || variant > 2 || !rng || !n_out || (variant == 2 && !ppm)) return OR_ERR_ARG;
    MENTER;
    int32_t n = c.n;
    midx_t *best = (midx_t *)malloc((size_t)(n > 1 ? n : 1) * sizeof(midx_t));
    if (!best) { rc = OR_ERR_NOMEM; MLEAVE; }
    int32_t acc_n = 0, best_n = 1;
    best[0].pwms = 0.0; best[0].npos = 0;               /* [|createMotifIndex 0. []|] */
    uint64_t chain0 = rng->chain;
    int64_t restart_index = 0;
    for (int32_t it = 0;; ++it) {
        if (it > reps) break;
        if (midx_arrays_equal(state, acc_n, best, best_n)) break;
        double ia = 0.0, ib = 0.0;
        for (int32_t i = 0; i < acc_n; ++i) ia = ia + state[i].pwms;
        for (int32_t i = 0; i < best_n; ++i) ib = ib + best[i].pwms;
        if (ia > ib) {
            if (acc_n != 0) { memcpy(best, state, (size_t)acc_n * sizeof(midx_t)); best_n = acc_n; }
            acc_n = 0;
        } else {
            if (rng->mode == 1) { rng->chain = chain0 + (uint64_t)restart_index; rng->next = 0; }
            restart_index++;
            rc = motif_pipeline(variant, &c, &s, m, cutoff, pcv, ppm, rng, state, st);
            if (rc) break;
            acc_n = n;
        }
    }
    rng->chain = chain0;
    if (!rc) { midx_unpack(best, best_n, pwms, npos, pos); *n_out = best_n; }
    free(best);
    MLEAVE;
}

/* fs:156-170: strict '>' over sums, starting from the empty array (sum 0.0) */
int or_get_best_information_content(const double *scores, const int32_t *lens, int32_t n_items,
                                    int32_t *best_index_out) {
    if (!scores || !lens || !best_index_out) return OR_ERR_ARG;
    int32_t best = -1;
    double best_sum = 0.0;
    const double *p = scores;
    for (int32_t i = 0; i < n_items; ++i) {
        double sum = 0.0;
        for (int32_t j = 0; j < lens[i]; ++j) sum = p[j] + sum;  /* fold (fun b (pwms,_) -> pwms + b) */
        if (sum > best_sum) { best = i; best_sum = sum; }
        p += lens[i];
    }
    *best_index_out = best;
    return OR_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* INCREMENTAL MODE of the SiteSampler pipelines (test infrastructure for full-size parity)    */
/* ------------------------------------------------------------------------------------------ */
/*
 * The faithful functions above cost O(N (L + 49 k) + W (L + 49 k)) per site update because they
 * repeat the reference's redundant work (from-scratch leave-one-out rebuilds fs:392-396, a PWM
 * per window fs:309, Array.skip / Array.append copies). At the BASELINE sizes (1000 x 500 bp x
 * 1024 chains, 10 k x 1 kb, 100 k x 200 bp) that is hours of CPU. This section computes the SAME
 * values with the SAME float64 operations -- it calls the same static helpers ppm_of_pfm
 * (fs:249-261), pwm_create (fs:282-287), pwm_segment_score (fs:290-293), pcv_normalized_of_fcv
 * (fs:115-120), fcv_subtract_segment (fs:84-88) -- and only changes how the INTEGER inputs of
 * those helpers are obtained:
 *   - leave-one-out counts = counts over all current sites minus the held-out one-hot (integer
 *     adds: exact), updated -old site / +new site when a greedy sweep moves a site;
 *   - fixed background: the PWM is built once per site update instead of once per window (same
 *     inputs, same divisions, same values);
 *   - data-derived background: symbol counts per sequence are tabulated once, so
 *     increaseInPlaceFCVOf (fs:79) is 49 integer adds instead of a pass over L symbols, and
 *     createFCVWithout |> fuse (fs:565-568) is a running integer total;
 *   - random starts: the Philox block of four draws is computed once, not once per draw; held-out
 *     sequences are independent there (fresh draws for every other sequence, fs:595-598) and run
 *     on n_threads OpenMP threads, each addressing the uniform stream by draw index.
 * tests/test_oracle_fast.py requires bit-identical (scores, positions) from both modes on every
 * small case of the existing suites; the golden fixtures of the BASELINE sizes
 * (tests/golden/make_fullsize_golden.py) are generated with this mode.
 */
#include <pthread.h>

/* dynamic parallel-for over [0, n_items) on n_threads POSIX threads (chunked, atomic cursor) */
typedef void (*pf_body)(void *ctx, int32_t item, int32_t thread);
typedef struct { pf_body body; void *ctx; int32_t n_items, chunk, thread; volatile int32_t *cursor; } pf_arg;
static void *pf_worker(void *p) {
    pf_arg *a = (pf_arg *)p;
    for (;;) {
        int32_t i0 = __sync_fetch_and_add(a->cursor, a->chunk);
        if (i0 >= a->n_items) break;
        int32_t i1 = i0 + a->chunk < a->n_items ? i0 + a->chunk : a->n_items;
        for (int32_t i = i0; i < i1; ++i) a->body(a->ctx, i, a->thread);
    }
    return NULL;
}
static void parallel_for(int32_t n_items, int32_t n_threads, int32_t chunk, pf_body body, void *ctx) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    volatile int32_t cursor = 0;
    pf_arg args[256];
    pthread_t th[256];
    for (int32_t t = 0; t < n_threads; ++t) {
        args[t].body = body; args[t].ctx = ctx; args[t].n_items = n_items; args[t].chunk = chunk > 0 ? chunk : 1;
        args[t].thread = t; args[t].cursor = &cursor;
    }
    int32_t started = 0;
    for (int32_t t = 1; t < n_threads; ++t) {
        if (pthread_create(&th[t], NULL, pf_worker, &args[t]) != 0) break;
        started = t;
    }
    pf_worker(&args[0]);
    for (int32_t t = 1; t <= started; ++t) pthread_join(th[t], NULL);
}

typedef struct {
    const ctx_t *c;
    int32_t *tot;      /* 49 * k: counts over all (shifted) sites                              */
    int32_t *cnt;      /* n * 49: symbol counts per sequence (data-derived background only)    */
    int64_t gtot[NS];  /* data background: sum over sequences of (cnt_i - k-mer counts of site) */
    int32_t mult[NS];  /* how often a slot occurs in `alphabet` (fuse adds once per occurrence) */
    double *ppm, *pwm; /* 49 * k                                                               */
    int data_bg;
} fast_t;

static void fast_free(fast_t *f) { free(f->tot); free(f->cnt); free(f->ppm); free(f->pwm); }

static int fast_init(fast_t *f, const ctx_t *c, int data_bg) {
    memset(f, 0, sizeof *f);
    f->c = c;
    f->data_bg = data_bg;
    size_t cells = (size_t)NS * c->k;
    f->tot = (int32_t *)calloc(cells, sizeof(int32_t));
    f->ppm = (double *)malloc(cells * sizeof(double));
    f->pwm = (double *)calloc(cells, sizeof(double));
    if (data_bg) f->cnt = (int32_t *)calloc((size_t)c->n * NS, sizeof(int32_t));
    if (!f->tot || !f->ppm || !f->pwm || (data_bg && !f->cnt)) return OR_ERR_NOMEM;
    for (int32_t a = 0; a < c->alen; ++a) f->mult[slot(c->alpha[a])] += 1;
    if (data_bg)
        for (int32_t i = 0; i < c->n; ++i) fcv_increase_in_place_of(seq_of(c, i), len_of(c, i), f->cnt + (size_t)i * NS);
    return OR_OK;
}

static void onehot_add(const ctx_t *c, int32_t *pfm, int32_t i, int32_t p, int32_t d) {
    const uint8_t *s = seq_of(c, i) + p;
    for (int32_t j = 0; j < c->k; ++j) pfm[slot(s[j]) * c->k + j] += d;
}

/* fs:301-314 with the PWM hoisted out of the window loop */
static void fast_scan_bpv(const ctx_t *c, const double *pwm, const uint8_t *src, int32_t len, double *score_out,
                          int32_t *pos_out, or_stats *st) {
    double high = 0.0;
    int32_t hi = 0, k = c->k;
    for (int32_t n = 0; n + k <= len; ++n) {
        double tmp = pwm_segment_score(pwm, src + n, k);
        if (tmp > high) { high = tmp; hi = n; }
    }
    if (st) { st->window_scores += len - k + 1; st->site_updates++; }
    *score_out = log2_ref(high);
    *pos_out = hi;
}

/* pwm_create (fs:282-287) for a matrix whose non-alphabet rows are already 0 */
static void pwm_rows(const ctx_t *c, const double *pcv, const double *ppm, double *pwm) {
    int32_t k = c->k;
    for (int32_t a = 0; a < c->alen; ++a) {
        int s = slot(c->alpha[a]);
        for (int32_t j = 0; j < k; ++j) pwm[s * k + j] = ppm[s * k + j] / pcv[s];
    }
}

/* fs:462-479: fcv starts as the fused background of the others; cnt_h = symbol counts of src */
static void fast_scan_drift(const ctx_t *c, int32_t *fcv, const int32_t *cnt_h, const double *ppm, double *pwm,
                            const uint8_t *src, int32_t len, double *score_out, int32_t *pos_out, or_stats *st) {
    double high = 0.0, pcv[NS];
    int32_t hi = 0, k = c->k;
    for (int32_t n = 0; n + k <= len; ++n) {
        for (int s = 0; s < NS; ++s) fcv[s] += cnt_h[s];   /* increaseInPlaceFCVOf source, fs:471 */
        fcv_subtract_segment(src + n, k, fcv);             /* fs:472 */
        pcv_normalized_of_fcv(c, fcv, pcv);                /* fs:473 */
        pwm_rows(c, pcv, ppm, pwm);                        /* fs:474 */
        double tmp = pwm_segment_score(pwm, src + n, k);
        if (tmp > high) { high = tmp; hi = n; }
    }
    if (st) { st->window_scores += len - k + 1; st->site_updates++; }
    *score_out = log2_ref(high);
    *pos_out = hi;
}

/* one site update from leave-one-out counts `pfm` (49 x k) and, data background, the fused fcv of the others */
static void fast_update(const fast_t *f, const int32_t *pfm, double *ppm, double *pwm, int32_t *fcv, const double *pcv,
                        int32_t h, double *ts, int32_t *tp, or_stats *st) {
    const ctx_t *c = f->c;
    ppm_of_pfm(c, pfm, c->n - 1, ppm);
    if (!f->data_bg) {
        pwm_rows(c, pcv, ppm, pwm);
        fast_scan_bpv(c, pwm, seq_of(c, h), len_of(c, h), ts, tp, st);
    } else {
        fast_scan_drift(c, fcv, f->cnt + (size_t)h * NS, ppm, pwm, seq_of(c, h), len_of(c, h), ts, tp, st);
    }
}

/* sweeps of fs:381 / fs:350 / fs:318 (pcv given) and fs:554 / fs:519 / fs:483 (data background) */
static int fast_sweeps(fast_t *f, const double *pcv, int mode, int32_t max_sweeps, double *score, int32_t *pos,
                       or_stats *st, int32_t *log, int32_t log_cap, int32_t *log_n) {
    const ctx_t *c = f->c;
    int32_t n = c->n, k = c->k;
    int32_t *snap = (int32_t *)malloc((size_t)n * sizeof(int32_t));
    if (!snap) return OR_ERR_NOMEM;
    memcpy(snap, pos, (size_t)n * sizeof(int32_t));
    int32_t fcv[NS], kc[NS];
    for (int32_t sweep = 0;; ++sweep) {
        /* all-sites counts of what the sweep reads: acc itself (greedy) or the shifted snapshot */
        memset(f->tot, 0, (size_t)NS * k * sizeof(int32_t));
        for (int32_t i = 0; i < n; ++i) onehot_add(c, f->tot, i, shifted(c, i, snap[i], mode), +1);
        if (f->data_bg) {
            memset(f->gtot, 0, sizeof f->gtot);
            for (int32_t i = 0; i < n; ++i)
                for (int s = 0; s < NS; ++s) f->gtot[s] += f->cnt[(size_t)i * NS + s];
            for (int s = 0; s < NS; ++s)
                for (int32_t j = 0; j < k; ++j) f->gtot[s] -= f->tot[s * k + j];
        }
        int32_t movers = 0, accepted = 0;
        for (int32_t h = 0; h < n; ++h) {
            int32_t own = shifted(c, h, mode == SHIFT_NONE ? pos[h] : snap[h], mode);
            onehot_add(c, f->tot, h, own, -1);
            if (f->data_bg) {
                memset(kc, 0, sizeof kc);
                for (int32_t j = 0; j < k; ++j) kc[slot(seq_of(c, h)[own + j])] += 1;
                for (int s = 0; s < NS; ++s)
                    fcv[s] = f->mult[s] * (int32_t)(f->gtot[s] - (f->cnt[(size_t)h * NS + s] - kc[s]));
            }
            double ts; int32_t tp;
            fast_update(f, f->tot, f->ppm, f->pwm, fcv, pcv, h, &ts, &tp, st);
            int32_t now = own;
            if (ts > score[h]) {                                   /* fs:402 / fs:579 */
                accepted++;
                if (tp != pos[h]) movers++;
                score[h] = ts; pos[h] = tp;
                if (mode == SHIFT_NONE) now = tp;                  /* in place: later h read acc (fs:388) */
            }
            onehot_add(c, f->tot, h, now, +1);
            if (f->data_bg && now != own) {
                for (int32_t j = 0; j < k; ++j) {
                    f->gtot[slot(seq_of(c, h)[own + j])] += 1;
                    f->gtot[slot(seq_of(c, h)[now + j])] -= 1;
                }
            }
        }
        if (st) st->sweeps++;
        if (log && *log_n < log_cap) { log[*log_n * 3] = mode; log[*log_n * 3 + 1] = movers; log[*log_n * 3 + 2] = accepted; ++*log_n; }
        if (positions_equal(pos, snap, n)) break;                  /* fs:384 / fs:557 */
        if (max_sweeps > 0 && sweep + 1 >= max_sweeps) break;
        memcpy(snap, pos, (size_t)n * sizeof(int32_t));
    }
    free(snap);
    return OR_OK;
}

/* uniform `draw` of the stream without touching rng->next (threads address the stream by index) */
static double rng_at(const or_rng *rng, int64_t draw, int *exhausted) {
    if (rng->mode == 0) {
        if (draw >= rng->n_u) { *exhausted = 1; return 0.0; }
        return rng->u[draw];
    }
    return or_uniform_at(rng->seed, rng->chain, (uint64_t)draw);
}

/* fs:412-430 / fs:589-611 */
typedef struct {
    fast_t *f; const double *pcv; const or_rng *rng; double *score; int32_t *pos; int64_t base0;
    int64_t ctot[NS];
    int32_t n_threads;
    int32_t **pfm; double **ppm, **pwm;   /* per-thread scratch */
    int64_t *ws; int *exhausted;          /* per-thread results */
} rs_ctx;

static void rs_body(void *p, int32_t h, int32_t t) {
    rs_ctx *r = (rs_ctx *)p;
    const fast_t *f = r->f;
    const ctx_t *c = f->c;
    const or_rng *rng = r->rng;
    const int32_t n = c->n, k = c->k;
    const size_t cells = (size_t)NS * k;
    int32_t *pfm = r->pfm[t];
    memset(pfm, 0, cells * sizeof(int32_t));
    int64_t d = r->base0 + (int64_t)h * (n - 1);
    uint32_t blk_words[4] = {0, 0, 0, 0};
    int64_t blk_have = -1;
    for (int32_t i = 0; i < n; ++i) {
        if (i == h) continue;
        double u;
        if (rng->mode == 1) {           /* Philox block of draws 4b .. 4b+3, computed once */
            if ((d >> 2) != blk_have) {
                blk_have = d >> 2;
                uint32_t ctr[4] = {(uint32_t)blk_have, (uint32_t)((uint64_t)blk_have >> 32), (uint32_t)rng->chain, (uint32_t)(rng->chain >> 32)};
                uint32_t key[2] = {(uint32_t)rng->seed, (uint32_t)(rng->seed >> 32)};
                or_philox4x32_10(ctr, key, blk_words);
            }
            u = (double)blk_words[d & 3] * (1.0 / 4294967296.0);
        } else {
            int ex = 0;
            u = rng_at(rng, d, &ex);
            r->exhausted[t] |= ex;
        }
        ++d;
        onehot_add(c, pfm, i, or_draw_to_position(u, len_of(c, i), k), +1);
    }
    int32_t fcv[NS];
    if (f->data_bg)
        for (int s = 0; s < NS; ++s) {
            int64_t site_cnt = 0;
            for (int32_t j = 0; j < k; ++j) site_cnt += pfm[s * k + j];
            fcv[s] = f->mult[s] * (int32_t)(r->ctot[s] - f->cnt[(size_t)h * NS + s] - site_cnt);
        }
    or_stats local = {0, 0, 0, 0};
    fast_update(f, pfm, r->ppm[t], r->pwm[t], fcv, r->pcv, h, &r->score[h], &r->pos[h], &local);
    r->ws[t] += local.window_scores;
}

static int fast_random_starts(fast_t *f, const double *pcv, or_rng *rng, double *score, int32_t *pos, or_stats *st,
                              int32_t n_threads) {
    const ctx_t *c = f->c;
    const int32_t n = c->n, k = c->k;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    rs_ctx r;
    memset(&r, 0, sizeof r);
    r.f = f; r.pcv = pcv; r.rng = rng; r.score = score; r.pos = pos; r.base0 = rng->next; r.n_threads = n_threads;
    if (f->data_bg)
        for (int32_t i = 0; i < n; ++i)
            for (int s = 0; s < NS; ++s) r.ctot[s] += f->cnt[(size_t)i * NS + s];
    int32_t *pfm[256]; double *ppm[256], *pwm[256]; int64_t ws[256]; int exhausted[256];
    size_t cells = (size_t)NS * k;
    int err = 0;
    for (int32_t t = 0; t < n_threads; ++t) {
        pfm[t] = (int32_t *)malloc(cells * sizeof(int32_t));
        ppm[t] = (double *)malloc(cells * sizeof(double));
        pwm[t] = (double *)calloc(cells, sizeof(double));
        ws[t] = 0; exhausted[t] = 0;
        if (!pfm[t] || !ppm[t] || !pwm[t]) err = 1;
    }
    r.pfm = pfm; r.ppm = ppm; r.pwm = pwm; r.ws = ws; r.exhausted = exhausted;
    if (!err) parallel_for(n, n_threads, 16, rs_body, &r);
    int64_t wsum = 0; int ex = 0;
    for (int32_t t = 0; t < n_threads; ++t) { free(pfm[t]); free(ppm[t]); free(pwm[t]); wsum += ws[t]; ex |= exhausted[t]; }
    if (err) return OR_ERR_NOMEM;
    rng->next = r.base0 + (int64_t)n * (n - 1);
    if (ex) rng->exhausted = 1;
    if (st) { st->sweeps++; st->site_updates += n; st->window_scores += wsum; }
    return OR_OK;
}

/*
 * variant 0 = WithBPV family (fs:691), 1 = data-derived background (fs:697). phase_mask: 1 random starts,
 * 2 greedy sweeps, 4 left shifts, 8 right shifts (0 = all four, the pipeline of doSiteSampling[WithBPV]);
 * without bit 1 the sweeps start from the (score, pos) passed in. sweep_log (nullable) receives
 * (mode, sites moved, updates accepted) per sweep of the greedy / shift phases.
 */
int or_fast_site_pipeline(int32_t variant, int32_t phase_mask, const uint8_t *seqs, const int64_t *off, int32_t n_seqs,
                          int32_t k, double pc, const uint8_t *alphabet, int32_t alen, const double *pcv, or_rng *rng,
                          double *score, int32_t *pos, or_stats *st, int32_t n_threads, int32_t *sweep_log,
                          int32_t sweep_log_cap, int32_t *sweep_log_n) {
    if (variant < 0 || variant > 1) return OR_ERR_ARG;
    MAKE_CTX(c);
    int rc = check_ctx(&c); if (rc) return rc;
    if (!score || !pos || (variant == 0 && !pcv)) return OR_ERR_ARG;
    if (phase_mask == 0) phase_mask = 15;
    if ((phase_mask & 1) && !rng) return OR_ERR_ARG;
    fast_t f;
    rc = fast_init(&f, &c, variant == 1);
    int32_t ln = 0;
    if (!rc && (phase_mask & 1)) rc = fast_random_starts(&f, pcv, rng, score, pos, st, n_threads);
    if (!rc && (phase_mask & 2)) rc = fast_sweeps(&f, pcv, SHIFT_NONE, 0, score, pos, st, sweep_log, sweep_log_cap, &ln);
    if (!rc && (phase_mask & 4)) rc = fast_sweeps(&f, pcv, SHIFT_LEFT, 0, score, pos, st, sweep_log, sweep_log_cap, &ln);
    if (!rc && (phase_mask & 8)) rc = fast_sweeps(&f, pcv, SHIFT_RIGHT, 0, score, pos, st, sweep_log, sweep_log_cap, &ln);
    if (!rc && st && phase_mask == 15) st->restarts++;
    if (sweep_log_n) *sweep_log_n = ln;
    fast_free(&f);
    return rc;
}

/* n_chains restarts (Philox streams (seed, chain_base + c)) on n_threads threads; results [n_chains][n_seqs],
 * sums = Array.sum of each restart's scores, left to right (fs:445) */
typedef struct {
    int32_t variant; const uint8_t *seqs; const int64_t *off; int32_t n_seqs, k; double pc; const uint8_t *alphabet;
    int32_t alen; const double *pcv; uint64_t seed; int64_t chain_base; double *scores; int32_t *pos; double *sums;
    or_stats *per_thread; int *rc;
} ch_ctx;

static void ch_body(void *p, int32_t ch, int32_t t) {
    ch_ctx *a = (ch_ctx *)p;
    or_rng rng;
    memset(&rng, 0, sizeof rng);
    rng.mode = 1; rng.seed = a->seed; rng.chain = (uint64_t)(a->chain_base + ch);
    int rc = or_fast_site_pipeline(a->variant, 15, a->seqs, a->off, a->n_seqs, a->k, a->pc, a->alphabet, a->alen, a->pcv,
                                   &rng, a->scores + (size_t)ch * a->n_seqs, a->pos + (size_t)ch * a->n_seqs,
                                   &a->per_thread[t], 1, NULL, 0, NULL);
    if (rc) a->rc[t] = rc;
    if (a->sums) a->sums[ch] = seq_sum(a->scores + (size_t)ch * a->n_seqs, a->n_seqs);
}

int or_fast_site_chains(int32_t variant, const uint8_t *seqs, const int64_t *off, int32_t n_seqs, int32_t k, double pc,
                        const uint8_t *alphabet, int32_t alen, const double *pcv, uint64_t seed, int64_t chain_base,
                        int32_t n_chains, int32_t n_threads, double *scores, int32_t *pos, double *sums, or_stats *st) {
    if (n_chains < 1 || !scores || !pos) return OR_ERR_ARG;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    or_stats per_thread[256];
    int rcs[256];
    memset(per_thread, 0, sizeof per_thread);
    memset(rcs, 0, sizeof rcs);
    ch_ctx a = {variant, seqs, off, n_seqs, k, pc, alphabet, alen, pcv, seed, chain_base, scores, pos, sums, per_thread, rcs};
    parallel_for(n_chains, n_threads, 1, ch_body, &a);
    int rc_all = OR_OK;
    for (int32_t t = 0; t < n_threads; ++t) {
        if (rcs[t]) rc_all = rcs[t];
        if (st) { st->site_updates += per_thread[t].site_updates; st->window_scores += per_thread[t].window_scores;
                  st->sweeps += per_thread[t].sweeps; st->restarts += per_thread[t].restarts; }
    }
    return rc_all;
}
