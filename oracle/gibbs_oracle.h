/*
 * oracle/gibbs_oracle.h -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
 *
 * C-ABI of the CPU oracle: a plain-C restatement of the reference's algorithm
 * (/root/reference/GibbsSampling/GibbsSampling.fs, cited below as fs:N).
 *
 * PARITY UNPINNED: the reference ships no tests, golden vectors or recorded outputs
 * and cannot be executed in this environment (no .NET runtime). The oracle is pinned
 * only by SURVEY.md Appendix A (behavioural spec) and Appendix B (derived known-answer
 * values, see tests/golden/appendix_b.json) plus an independent pure-Python model in
 * tests/pymodel.py.
 *
 * Conventions
 *   symbols   : ASCII bytes, slot index = byte - 42 (fs:17, fs:176), 49 slots (fs:20, fs:179)
 *   sequences : one concatenated byte buffer + int64 offsets[n_seqs + 1]
 *   alphabet  : ASCII bytes (the script uses "ATGC-", fsx:368-369)
 *   (float*int)[] : parallel arrays double score[n], int32 pos[n]
 *   MotifIndex[]  : double pwms[n], int32 npos[n], int32 pos[n * OR_MAX_M] (list order of fs:736)
 *   uniforms  : or_rng -- either an injected array consumed in call order, or Philox4x32-10
 */
#ifndef GIBBS_ORACLE_H
#define GIBBS_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OR_NSLOT 49
#define OR_MAX_M 4

enum {
    OR_OK = 0,
    OR_ERR_ARG = 1,        /* ArgumentNullException / bad sizes                          */
    OR_ERR_SYMBOL = 2,     /* IndexOutOfRangeException: symbol outside '*'..'Z' (fs:17)  */
    OR_ERR_SHORT_SEQ = 3,  /* InvalidOperationException from Array.take (fs:152, fs:308) */
    OR_ERR_ROULETTE = 4,   /* ArgumentException: pick beyond accumulated mass (fs:753)   */
    OR_ERR_NOMEM = 6
};

typedef struct {
    int32_t mode;          /* 0 = injected array, 1 = Philox4x32-10(seed, chain)          */
    const double *u;       /* mode 0: uniforms in [0,1) consumed sequentially             */
    int64_t n_u;
    int64_t next;          /* draws consumed so far (also the Philox draw index)          */
    uint64_t seed;
    uint64_t chain;
    int32_t exhausted;     /* set when mode 0 ran out (draw returns 0.0)                  */
} or_rng;

typedef struct {
    int64_t site_updates;  /* scans of one held-out sequence (fs:301 / fs:462 / fs:759)   */
    int64_t window_scores; /* windows scored inside those scans                           */
    int64_t sweeps;        /* passes n = 0..N-1                                           */
    int64_t restarts;      /* restart pipelines executed                                  */
} or_stats;

/* ---- RNG ---- */
void or_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
double or_uniform_at(uint64_t seed, uint64_t chain, uint64_t draw);
double or_next_uniform(or_rng *rng);
int32_t or_draw_to_position(double u, int32_t len, int32_t k);          /* fs:143-146 */

/* ---- primitives (L1/L2 of the reference) ---- */
/* fs:392-396 / fs:218: leave-one-out PFM, int32 [49*k] row-major (slot, column) */
int or_loo_pfm(const uint8_t *seqs, const int64_t *off, int32_t n_seqs, const int32_t *sites,
               int32_t heldout, int32_t k, int32_t *pfm_out);
/* fs:249-261: PPM from PFM */
int or_ppm_of_pfm(const int32_t *pfm, int32_t k, int32_t source_count, const uint8_t *alphabet,
                  int32_t alen, double pc, double *ppm_out);
/* A.3 fixed mode: createFCVOf (fs:60) per sequence -> fuse (fs:65) -> normalise (fs:115) */
int or_pcv_of_sources(const uint8_t *seqs, const int64_t *off, int32_t n_seqs,
                      const uint8_t *alphabet, int32_t alen, double pc, double *pcv_out);
/* fs:282-293 applied to every window of one sequence: raw float64 products, W = len-k+1 */
int or_window_scores_bpv(const uint8_t *src, int32_t len, int32_t k, const uint8_t *alphabet,
                         int32_t alen, const double *pcv, const double *ppm, double *scores_out);
/* fs:301-314 */
int or_best_pwms_with_bpv(const uint8_t *src, int32_t len, int32_t k, const uint8_t *alphabet,
                          int32_t alen, const double *pcv, const double *ppm, double *score_out,
                          int32_t *pos_out);
/* fs:462-479 (drifting background; fcv is mutated in place like the reference) */
int or_best_pwms(const uint8_t *src, int32_t len, int32_t k, const uint8_t *alphabet, int32_t alen,
                 double pc, int32_t *fcv, const double *ppm, double *score_out, int32_t *pos_out,
                 double *window_scores_out /* nullable, raw products */);
/* fs:565-568: fused createFCVWithout of every other sequence */
int or_loo_fcv(const uint8_t *seqs, const int64_t *off, int32_t n_seqs, const int32_t *sites,
               int32_t heldout, int32_t k, const uint8_t *alphabet, int32_t alen, int32_t *fcv_out);
/* fs:759-784: candidate list for one sequence. Outputs up to cap entries. */
int or_candidates(const uint8_t *src, int32_t len, int32_t k, int32_t m, double cutoff,
                  const uint8_t *alphabet, int32_t alen, const double *pcv, const double *ppm,
                  double *pwms_out, int32_t *npos_out, int32_t *pos_out, int64_t cap,
                  int64_t *n_out);
/* fs:746-754 */
int or_roulette(const double *pwms, int64_t n, double pick, int64_t *index_out);

/* ---- SiteSampler (fs:298-707) ---- */
int or_random_starts_with_bpv(const uint8_t *seqs, const int64_t *off, int32_t n_seqs, int32_t k,
                              double pc, const uint8_t *alphabet, int32_t alen, const double *pcv,
                              or_rng *rng, double *score, int32_t *pos, or_stats *st);      /* fs:412 */
int or_find_best_motif_with_start_position(const uint8_t *seqs, const int64_t *off, int32_t n_seqs,
                              int32_t k, double pc, const uint8_t *alphabet, int32_t alen,
                              const double *pcv, double *score, int32_t *pos, or_stats *st); /* fs:381 */
int or_left_shifted_with_bpv(const uint8_t *seqs, const int64_t *off, int32_t n_seqs, int32_t k,
                              double pc, const uint8_t *alphabet, int32_t alen, const double *pcv,
                              double *score, int32_t *pos, or_stats *st);                   /* fs:350 */
int or_right_shifted_with_bpv(const uint8_t *seqs, const int64_t *off, int32_t n_seqs, int32_t k,
                              double pc, const uint8_t *alphabet, int32_t alen, const double *pcv,
                              double *score, int32_t *pos, or_stats *st);                   /* fs:318 */
int or_do_site_sampling_with_bpv(const uint8_t *seqs, const int64_t *off, int32_t n_seqs, int32_t k,
                              double pc, const uint8_t *alphabet, int32_t alen, const double *pcv,
                              or_rng *rng, double *score, int32_t *pos, or_stats *st);      /* fs:691 */

int or_random_starts(const uint8_t *seqs, const int64_t *off, int32_t n_seqs, int32_t k, double pc,
                     const uint8_t *alphabet, int32_t alen, or_rng *rng, double *score,
                     int32_t *pos, or_stats *st);                                           /* fs:589 */
int or_best_pwms_with_start_positions(const uint8_t *seqs, const int64_t *off, int32_t n_seqs,
                     int32_t k, double pc, const uint8_t *alphabet, int32_t alen, double *score,
                     int32_t *pos, or_stats *st);                                           /* fs:554 */
int or_left_shifted(const uint8_t *seqs, const int64_t *off, int32_t n_seqs, int32_t k, double pc,
                     const uint8_t *alphabet, int32_t alen, double *score, int32_t *pos,
                     or_stats *st);                                                         /* fs:519 */
int or_right_shifted(const uint8_t *seqs, const int64_t *off, int32_t n_seqs, int32_t k, double pc,
                     const uint8_t *alphabet, int32_t alen, double *score, int32_t *pos,
                     or_stats *st);                                                         /* fs:483 */
int or_do_site_sampling(const uint8_t *seqs, const int64_t *off, int32_t n_seqs, int32_t k,
                     double pc, const uint8_t *alphabet, int32_t alen, or_rng *rng, double *score,
                     int32_t *pos, or_stats *st);                                           /* fs:697 */
int or_motifs_with_best_pwms_of_ppm(const uint8_t *seqs, const int64_t *off, int32_t n_seqs,
                     int32_t k, double pc, const uint8_t *alphabet, int32_t alen, const double *ppm,
                     or_rng *rng, double *score, int32_t *pos, or_stats *st);               /* fs:644 */
int or_do_site_sampling_with_ppm(const uint8_t *seqs, const int64_t *off, int32_t n_seqs, int32_t k,
                     double pc, const uint8_t *alphabet, int32_t alen, const double *ppm,
                     or_rng *rng, double *score, int32_t *pos, or_stats *st);               /* fs:703 */

/* restart loops: variant 0 = WithBPV (fs:434), 1 = data-derived (fs:615), 2 = OfPPM (fs:664).
 * Each restart r consumes the uniform stream (seed, chain_base + r) in Philox mode, or continues
 * the injected array in mode 0. n_out = length of the returned array (1 when the initial
 * [|(0.,0)|] survives, quirk A.6-8). */
int or_best_information_content(int32_t variant, int32_t reps, const uint8_t *seqs,
                     const int64_t *off, int32_t n_seqs, int32_t k, double pc,
                     const uint8_t *alphabet, int32_t alen, const double *pcv_or_null,
                     const double *ppm_or_null, or_rng *rng, double *score, int32_t *pos,
                     int32_t *n_out, or_stats *st);

/* ---- MotifSampler (fs:709-1038) ---- */
/* variant 0 = fixed pcv (fs:788 / fs:828), 1 = data-derived once-per-n background (fs:885 / fs:935) */
int or_motif_greedy(int32_t variant, const uint8_t *seqs, const int64_t *off, int32_t n_seqs,
                    int32_t m, int32_t k, double pc, double cutoff, const uint8_t *alphabet,
                    int32_t alen, const double *pcv_or_null, double *pwms, int32_t *npos,
                    int32_t *pos, or_stats *st);
int or_motif_stochastic(int32_t variant, const uint8_t *seqs, const int64_t *off, int32_t n_seqs,
                    int32_t m, int32_t k, double pc, double cutoff, const uint8_t *alphabet,
                    int32_t alen, const double *pcv_or_null, or_rng *rng, double *pwms,
                    int32_t *npos, int32_t *pos, or_stats *st);
/* pipelines: variant 0 = with pcv (fs:876-879), 1 = doMotifSampling (fs:1034),
 * 2 = doMotifSamplingWithPPM (fs:1028) */
int or_do_motif_sampling(int32_t variant, const uint8_t *seqs, const int64_t *off, int32_t n_seqs,
                    int32_t m, int32_t k, double pc, double cutoff, const uint8_t *alphabet,
                    int32_t alen, const double *pcv_or_null, const double *ppm_or_null,
                    or_rng *rng, double *pwms, int32_t *npos, int32_t *pos, or_stats *st);
/* restart loops fs:856 (variant 0), fs:973 (variant 1), fs:1001 (variant 2) */
int or_best_motif_information_content(int32_t variant, int32_t reps, const uint8_t *seqs,
                    const int64_t *off, int32_t n_seqs, int32_t m, int32_t k, double pc,
                    double cutoff, const uint8_t *alphabet, int32_t alen,
                    const double *pcv_or_null, const double *ppm_or_null, or_rng *rng,
                    double *pwms, int32_t *npos, int32_t *pos, int32_t *n_out, or_stats *st);

/* ---- incremental mode (see the section of that name in gibbs_oracle.c): the same float64 operations on the same
 * integer counts, obtained without the reference's from-scratch rebuilds; bit-identical to the functions above
 * (tests/test_oracle_fast.py) and fast enough for the BASELINE sizes. variant 0 = WithBPV (fs:691), 1 = data-derived
 * background (fs:697). phase_mask: 1 random starts | 2 greedy | 4 left shifts | 8 right shifts (0 = all). ---- */
int or_fast_site_pipeline(int32_t variant, int32_t phase_mask, const uint8_t *seqs, const int64_t *off,
                          int32_t n_seqs, int32_t k, double pc, const uint8_t *alphabet, int32_t alen,
                          const double *pcv_or_null, or_rng *rng, double *score, int32_t *pos, or_stats *st,
                          int32_t n_threads, int32_t *sweep_log /* nullable [cap][3]: mode, moved, accepted */,
                          int32_t sweep_log_cap, int32_t *sweep_log_n);
int or_fast_site_chains(int32_t variant, const uint8_t *seqs, const int64_t *off, int32_t n_seqs, int32_t k,
                        double pc, const uint8_t *alphabet, int32_t alen, const double *pcv_or_null,
                        uint64_t seed, int64_t chain_base, int32_t n_chains, int32_t n_threads,
                        double *scores, int32_t *pos, double *sums, or_stats *st);

/* fs:156-170 */
int or_get_best_information_content(const double *scores, const int32_t *lens, int32_t n_items,
                                    int32_t *best_index_out /* -1 = the empty start value */);

#ifdef __cplusplus
}
#endif
#endif
