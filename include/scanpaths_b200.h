/* scanpaths_b200 -- C ABI of the B200-native scanpath sampling + scoring path.
 *
 * This is the drop-in boundary.  The reference (chenxy99/Scanpaths) is pure
 * Python and has no FFI of its own; each entry point below replaces one of the
 * Python callables that the reference's train.py / test.py import (file:line
 * cited per function, paths relative to the reference root, identical in
 * OSIE/, AiR/, COCO_Search18/ unless noted).  INTEGRATION.md shows the ctypes
 * stubs a maintainer adds on the reference side.
 *
 * Conventions
 *   - every pointer named d_* is a DEVICE pointer owned by the caller (e.g. a
 *     torch tensor's data_ptr()); h_* is a HOST pointer.  Nothing is allocated,
 *     freed or synchronised inside; work is enqueued on `stream`.
 *   - every function returns 0 on success or a negative spb_status; the message
 *     of the last failure on the calling thread is spb_last_error().
 *   - plain C types only.  There is no CPU fallback: without a CUDA device the
 *     launch functions return SPB_ERR_CUDA.
 */
#ifndef SCANPATHS_B200_H
#define SCANPATHS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void *spb_stream; /* cudaStream_t */

enum spb_status {
    SPB_OK = 0,
    SPB_ERR_ARG = -1,       /* bad argument (null pointer, size out of range) */
    SPB_ERR_CUDA = -2,      /* CUDA runtime / driver error (message has the cudaError string) */
    SPB_ERR_WORKSPACE = -3, /* caller-provided workspace too small */
    SPB_ERR_UNSUPPORTED = -4
};

int spb_version(void);
const char *spb_last_error(void);

/* Number of kernels this library has launched since it was loaded (bench.py's gpu_launches). */
int64_t spb_kernel_launches(void);

/* Live profiling for bench.py's roofline: with max_pairs > 0 the decode path brackets its
 * tagged launches (1 = x-conv, 2 = 3x3 gate conv, 3 = 5x5 conv, 4 = LSTM cell, 5 = head,
 * 6 = feedback + attention, 7 = rank-1 projection, 8 = operand prep, 9 = Winograd input transform,
 * 10 = score_pairs kernel, 11 = sampling kernel) with CUDA events on the
 * launching stream; spb_profile_collect waits for them and returns (milliseconds, tag) per
 * bracket, then resets.  max_pairs = 0 turns it off. */
int spb_profile_enable(int32_t max_pairs);
int spb_profile_collect(float *h_ms, int32_t *h_tags, int32_t cap, int32_t *n_out);

/* ------------------------------------------------------------------------
 * ScanMatch configuration and tables  (utils/evaltools/scanmatch.py:43-114)
 * ---------------------------------------------------------------------- */
typedef struct spb_scanmatch_cfg {
    int32_t Xres, Yres, Xbin, Ybin;
    double Threshold, GapValue, TempBin, OffsetX, OffsetY;
} spb_scanmatch_cfg;

/* Host helper mirroring ScanMatch.CreateSubMatrix / GridMask.
 *   h_sub_delta [Ybin*Xbin]: substitution score as a function of (|d row|, |d col|)
 *                            (SubMatrix[a,b] depends only on those; bit-equal to it)
 *   h_sub_full  [nb*nb] or NULL: the full SubMatrix, nb = Xbin*Ybin
 *   h_xlut [Xres], h_ylut [Yres]: pixel -> bin column / row (the float-arange LUT)
 *   h_max_sub: numpy.max(SubMatrix), the score normaliser (scanmatch.py:191)      */
int spb_scanmatch_tables(const spb_scanmatch_cfg *cfg, double *h_sub_delta, double *h_sub_full,
                         uint8_t *h_xlut, uint8_t *h_ylut, double *h_max_sub);

/* Everything the scoring kernels need besides the scanpaths. */
typedef struct spb_score_cfg {
    spb_scanmatch_cfg sm;          /* sm.TempBin is the with-duration bin (50 ms in evaluation.py:159) */
    int32_t sed_height, sed_width, sed_n; /* stimulus (240, 320), n = 5 (visual_attention_metrics.py:301-309) */
    int32_t reserved;
    double stde_max_dim;           /* max(np.shape(stimulus)) = 320 (:407) */
    double dur_scale;              /* duration multiplier applied first: 1000 (s -> ms, evaluation.py:182) */
    double max_sub;                /* from spb_scanmatch_tables */
    const double *d_sub_delta;     /* device copies of the tables */
    const uint8_t *d_xlut;
    const uint8_t *d_ylut;
    const uint8_t *d_mask;         /* optional [Yres*Xres] pixel -> symbol table installed by ScanMatch.maskFromArray */
                                   /* (scanmatch.py:199-200); NULL = the regular grid (ylut[y]*Xbin + xlut[x])         */
} spb_score_cfg;

/* ------------------------------------------------------------------------
 * K1 scanpath_prep: fixations -> symbols, once per scanpath (not per pair).
 * Replaces ScanMatch.fixationToSequence x2 (scanmatch.py:116-133),
 * _scanpath_to_string (visual_attention_metrics.py:288-298) and the STDE
 * rescaling (:409-415).
 *   d_xyd [n_paths, lmax, 3] f64 (x, y, duration), d_len [n_paths]
 * out (the "symbol pack"):
 *   d_sym [n_paths, lmax] u8      ScanMatch bin symbol of each fixation (w/o-duration string)
 *   d_run [n_paths, lmax] i32     temporal-bin repeat count (with-duration string = sym repeated run times)
 *   d_nwd [n_paths] i32           length of the with-duration string (sum of runs)
 *   d_sed [n_paths, lmax] i32     SED grid symbol (chr(97+sq) in the reference)
 *   d_xyn [n_paths, lmax, 2] f64  x / max_dim, y / max_dim                                  */
int spb_prep_paths(const double *d_xyd, const int32_t *d_len, int64_t n_paths, int32_t lmax,
                   const spb_score_cfg *cfg, uint8_t *d_sym, int32_t *d_run, int32_t *d_nwd,
                   int32_t *d_sed, double *d_xyn, spb_stream stream);

typedef struct spb_path_pack {
    const uint8_t *d_sym;
    const int32_t *d_run;
    const int32_t *d_nwd;
    const int32_t *d_sed;
    const double *d_xyn;
    const int32_t *d_len;
    int64_t n_paths;
    int32_t lmax;
    int32_t reserved;
} spb_path_pack;

/* ------------------------------------------------------------------------
 * K2-K4 score_pairs: one warp per (human, simulated) pair, anti-diagonal
 * wavefront DPs.  Replaces, per pair, ScanMatch.match with and without
 * duration (scanmatch.py:135-150, 190-193; score only -- every caller drops
 * align and F), string_edit_distance (visual_attention_metrics.py:301-317) and
 * scaled_time_delay_embedding_similarity (:393-441).
 *   pair p scores human path d_pair_h[p] (first argument of the reference
 *   calls) against simulated path d_pair_s[p]; both packs may be the same
 *   (human_evaluation).
 *   d_scores [n_pairs, 4] f64 = (ScanMatch with duration, ScanMatch w/o duration, SED, STDE);
 *   NaN where the reference yields NaN / None (both strings empty; a path empty).
 *   d_workspace: only used when a simulated with-duration string exceeds 256
 *   symbols; spb_score_workspace_bytes(max human nwd) bytes, may be NULL otherwise.
 *   d_err: device int32 set non-zero if a pair needed more workspace than given.   */
int64_t spb_score_workspace_bytes(int64_t max_human_nwd);

/* The F matrix of ScanMatch.match for ONE pair of symbol strings (scanmatch.py:138-150), what the single-pair
 * API needs for the alignment it returns (scanmatch.py:152-195; the O(n + m) traceback itself runs on the host):
 *   d_F [(n+1), (m+1)] f64 row-major, F[i][0] = GapValue * (i + 1), F[0][j] = GapValue * (j + 1),
 *   F[i][j] = max(F[i-1][j-1] + Sub[a[i-1]][b[j-1]], F[i][j-1] + GapValue, F[i-1][j] + GapValue); bit-identical
 *   to the reference's numpy loop.  Symbols must lie in [0, Xbin * Ybin). */
int spb_scanmatch_matrix(const int32_t *d_a, int32_t n, const int32_t *d_b, int32_t m, const spb_score_cfg *cfg,
                         double *d_F, spb_stream stream);

/* The time-delay-embedding distances of ONE pair for every window length k = 1 .. min(Lh, Ls)
 * (time_delay_embedding_distance, visual_attention_metrics.py:332-390; its callers euclidean_distance :205-218
 * and scaled_time_delay_embedding_distance :444-492 are host arithmetic on this table):
 *   d_h_xy [Lh, 2], d_s_xy [Ls, 2] f64 (x, y), no rescaling here;
 *   d_out [min(Lh, Ls), 3] f64: (distance_mode 'Mean', distance_mode 'Hausdorff', sum of the first k point distances);
 *   d_work: spb_tde_work_bytes(Lh, Ls) bytes. */
int64_t spb_tde_work_bytes(int32_t Lh, int32_t Ls);
int spb_tde_distances(const double *d_h_xy, int32_t Lh, const double *d_s_xy, int32_t Ls, double *d_work,
                      int64_t work_bytes, double *d_out, spb_stream stream);
int spb_score_pairs(const spb_path_pack *human, const spb_path_pack *sim, const int32_t *d_pair_h,
                    const int32_t *d_pair_s, int64_t n_pairs, const spb_score_cfg *cfg, double *d_scores,
                    void *d_workspace, int64_t workspace_bytes, int32_t *d_err, spb_stream stream);

/* ------------------------------------------------------------------------
 * a7 reductions of the score table.  Replaces the aggregation halves of
 * pairs_eval (OSIE/utils/evaluation.py:325-338) and pairs_eval_scanmatch
 * (COCO_Search18/utils/evaluation.py:342-349) for a table laid out
 * [n_groups, group_size, 4] (one group = one simulated path against the S
 * subjects of its image).
 *   d_valid [n_groups, group_size] u8 or NULL: rows dropped by the MultiMatch
 *   NaN rule (either path shorter than min_len_valid); rows whose
 *   scores are NaN are dropped too.
 *   d_out [n_groups, 11] f32: slots 5..10 = SM w/o duration, SM with duration,
 *   SED mean, STDE mean, SED best (min), STDE best (max); means divide by
 *   group_size; slots 0..4 (MultiMatch, out of scope) hold the placeholder 0; the whole row is NaN
 *   if no row survives (what the reference returns, and what train.py:237 rejects a trial on).
 *   d_reward [n_groups] f64 or NULL: harmonic mean of slots 5, 6 (train.py:252).  */
int spb_reduce_pairs_eval(const double *d_scores, const uint8_t *d_valid, int64_t n_groups, int32_t group_size,
                          float *d_out, double *d_reward, spb_stream stream);

/* The same reduction with everything the drivers need in ONE pass over the table
 * (pairs_eval :284-340, evaluation's aggregation :211-237, the SCST trial-rejection rule train.py:237-238):
 *   - images with fewer subjects than group_size (lists padded by the packer): d_group_count[g % n_images]
 *     real subjects; padded pairs are skipped and the table divides by the real count (len(gt), :329);
 *   - the MultiMatch NaN rule computed here from the path lengths (min_len_valid = 3; 0 = off) instead of a
 *     caller-built mask: needs the pair maps of spb_score_pairs and both packs' d_len;
 *   - d_acc: running sums over ALL real pairs (evaluation() eliminates nothing), added to on every call:
 *     [0..3] sum of (SM-wd, SM-wod, SED, STDE), [4..7] sums of squares, [8..9] sum of the per-group SED min /
 *     STDE max, [10..11] their squares, [12] pairs, [13] groups, [14] groups with a surviving row.  The buffer
 *     holds spb_reduce_acc_bytes() bytes (result slots + per-block partials + a counter), zero-initialised
 *     once by the caller; the sum order is fixed (no floating-point atomics), so results are reproducible;
 *   - d_group_valid [n_groups] u8: 1 where at least one row survives (a trial is accepted iff all are 1).
 * Slots 0..4 of d_out are 0 where a row survives (MultiMatch placeholder) and the whole row is NaN otherwise. */
typedef struct spb_reduce_args {
    const double *d_scores;        /* [n_groups * group_size, 4] */
    const uint8_t *d_valid;        /* optional caller mask [n_groups * group_size] */
    const int32_t *d_pair_h;       /* optional, with d_pair_s / d_len_h / d_len_s and min_len_valid > 0 */
    const int32_t *d_pair_s;
    const int32_t *d_len_h;
    const int32_t *d_len_s;
    const int32_t *d_group_count;  /* optional [n_images] */
    float *d_out;                  /* optional [n_groups, 11] */
    double *d_reward;              /* optional [n_groups] */
    uint8_t *d_group_valid;        /* optional [n_groups] */
    double *d_acc;                 /* optional, spb_reduce_acc_bytes() bytes */
    int64_t acc_bytes;
    int64_t n_groups;
    int32_t group_size, n_images, min_len_valid;
    int32_t acc_blocks;            /* set by the library */
    int32_t mean_over_kept;        /* 0: table means divide by the subject count (pairs_eval, :329); 1: by the number of */
    int32_t reserved;              /*    surviving rows (AiR pairs_eval_scanmatch_performance_related, evaluation.py:405-416) */
} spb_reduce_args;

int64_t spb_reduce_acc_bytes(void);
int spb_reduce_pairs(const spb_reduce_args *args, spb_stream stream);

/* ------------------------------------------------------------------------
 * K5 sample_paths.  Replaces Sampling.random_sample + generate_scanpath
 * (models/sampling.py:16-77) for K independent samples per image.
 *   d_probs [N,T,A] f32, d_mu, d_sigma2 [N,T] f32 (decoder outputs)
 *   d_q [K,N,T,A] f32 or NULL: injected Exp(1) draws of the categorical race
 *   d_z [K,N,T]   f32 or NULL: injected N(0,1) draws of the duration
 *   (NULL -> Philox4x32-10 streams from `seed`)
 * out, all [K,N,T] unless noted:
 *   d_actions i32, d_sel_prob f32 (gathered from the UNMASKED probs), d_dur f32,
 *   d_action_mask, d_duration_mask f32, d_length [K,N] f32 (scanpath_length quirk kept),
 *   d_xyd [K*N, T, 3] f64 (x px, y px, duration s) + d_len [K*N] i32: the predicted
 *   scanpaths truncated at the first stop action, ready for spb_prep_paths.            */
typedef struct spb_sample_geom {
    int32_t map_width, map_height; /* 40 x 30 action map */
    int32_t width, height;         /* 320 x 240 image */
    int32_t min_length;            /* first min_length steps cannot stop (sampling.py:20) */
    int32_t reserved;
} spb_sample_geom;

int spb_sample_paths(const float *d_probs, const float *d_mu, const float *d_sigma2, const float *d_q,
                     const float *d_z, uint64_t seed, int32_t N, int32_t T, int32_t A, int32_t K,
                     const spb_sample_geom *geom, int32_t *d_actions, float *d_sel_prob, float *d_dur,
                     float *d_action_mask, float *d_duration_mask, float *d_length, double *d_xyd,
                     int32_t *d_len, spb_stream stream);

/* generate_scanpath alone (models/sampling.py:48-77) for given actions / durations
 * [n_samples, T]: masks, scanpath_length and the packed scanpaths as above. */
int spb_generate_scanpaths(const int32_t *d_actions, const float *d_dur, int64_t n_samples, int32_t T,
                           const spb_sample_geom *geom, float *d_action_mask, float *d_duration_mask,
                           float *d_length, double *d_xyd, int32_t *d_len, spb_stream stream);

/* ------------------------------------------------------------------------
 * a14 log-likelihoods (models/loss.py:10-45) and f4, the self-critical (SCST) loss tail
 * (OSIE/train.py:242-258; COCO_Search18/train.py:255-287; AiR/train.py:281-342), forward and
 * analytic backward.  eps = 1e-7; log-normal density with sigma2 as the VARIANCE;
 * LogAction / LogDuration rows are divided by the whole batch's mask.sum() (quirks kept).
 * ---------------------------------------------------------------------- */

/* LogAction (loss.py:34-37) and LogDuration (:39-45) for K stacked calls ("trials") over the same N images:
 *   d_log_actions[k,n]   = sum_t log(p[k,n,t] + eps) action_mask[k,n,t] / sum(action_mask[k])
 *   d_log_durations[k,n] = sum_t logpdf(x[k,n,t]; mu[n,t], sigma2[n,t]) duration_mask[k,n,t] / sum(duration_mask[k])
 * p is d_p [K,N,T] or, when d_p is NULL, gathered from d_probs [N,T,A] at d_actions [K,N,T] (sampling.py:24).
 * Either output may be NULL.  d_mask_sums [K,2] (or NULL) receives the two mask sums per trial. */
int spb_loglik_rows(const float *d_p, const float *d_probs, const int32_t *d_actions, const float *d_x,
                    const float *d_mu, const float *d_sigma2, const float *d_action_mask,
                    const float *d_duration_mask, int32_t K, int32_t N, int32_t T, int32_t A,
                    float *d_log_actions, float *d_log_durations, float *d_mask_sums, spb_stream stream);
/* Backward for upstream row gradients [K,N]: d_grad_p [K,N,T]; d_grad_mu / d_grad_sigma2 [N,T] (summed over
 * the trials in order); d_grad_x [K,N,T].  Outputs may be NULL. */
int spb_loglik_rows_backward(const float *d_p, const float *d_x, const float *d_mu, const float *d_sigma2,
                             const float *d_action_mask, const float *d_duration_mask, const float *d_mask_sums,
                             const float *d_grad_log_actions, const float *d_grad_log_durations, int32_t K, int32_t N,
                             int32_t T, float *d_grad_p, float *d_grad_mu, float *d_grad_sigma2, float *d_grad_x,
                             spb_stream stream);

/* CrossEntropyLoss (loss.py:10-14): logits [rows, A], dense target gt [rows, A], mask [rows].
 * Forward (d_grad_logits NULL): d_row_loss [rows] scratch, d_loss [1], d_mask_sum [1] written.
 * Backward (d_grad_logits given): d_grad_out [1] upstream gradient, d_mask_sum from the forward call. */
int spb_cross_entropy(const float *d_logits, const float *d_gt, const float *d_mask, int64_t rows, int32_t A,
                      float *d_row_loss, float *d_loss, float *d_mask_sum, const float *d_grad_out,
                      float *d_grad_logits, spb_stream stream);

/* MLPLogNormalDistribution (loss.py:27-32): -sum_{mask == 1} logpdf(gt; mu, sigma2) / mask.sum(), n elements.
 * Forward (d_grad_mu NULL): d_item [n] scratch, d_loss [1], d_mask_sum [1]; backward: d_grad_mu, d_grad_sigma2 [n]. */
int spb_lognormal_nll(const float *d_mu, const float *d_sigma2, const float *d_gt, const float *d_mask, int64_t n,
                      float *d_item, float *d_loss, float *d_mask_sum, const float *d_grad_out, float *d_grad_mu,
                      float *d_grad_sigma2, spb_stream stream);

/* The SCST tail in two launches.  Inputs: the decoder outputs of N images, K sampled trials
 * (spb_sample_paths) and their rewards (spb_reduce_pairs: harmonic mean of ScanMatch w/o and with duration).
 *   accepted trial : all of its N images kept a scored pair (d_group_valid, train.py:237-238) and no reward is NaN;
 *                    the first k_use accepted trials are used (d_trial_used[k] = 1; d_trial_used[K] = how many --
 *                    fewer than k_use means the caller has to sample more, the reference loops until it has them)
 *   d_adv[k,n]     = reward[k,n] - mean over used trials of reward[.,n]  (+ d_extra_adv[k,n] if given: the hook
 *                    for AiR's consistency-divergence term, AiR/train.py:322-333)
 *   d_loss[0..2]   = loss, loss_actions, loss_duration = sum (-LogAction - LogDuration) * adv    (train.py:256-258)
 * spb_scst_loss_backward writes d loss / d all_actions_prob [N,T,A] (zero-filled, then the gather's scatter),
 * d loss / d log_normal_mu and / d log_normal_sigma2 [N,T] for the upstream gradient d_grad_loss [3] on
 * (loss, loss_actions, loss_duration) (NULL = back-propagate `loss` with gradient 1). */
typedef struct spb_scst_args {
    const float *d_probs;          /* [N,T,A] all_actions_prob */
    const float *d_mu, *d_sigma2;  /* [N,T] */
    const int32_t *d_actions;      /* [K,N,T] */
    const float *d_dur;            /* [K,N,T] sampled durations (constants: train.py:233 clones .data) */
    const float *d_action_mask, *d_duration_mask; /* [K,N,T] */
    const double *d_reward;        /* [K,N] */
    const uint8_t *d_group_valid;  /* [K,N] or NULL */
    const float *d_extra_adv;      /* [K,N] or NULL */
    float *d_loss;                 /* [3] */
    float *d_adv;                  /* [K,N] */
    float *d_log_actions, *d_log_durations; /* [K,N] */
    float *d_mask_sums;            /* [K,2] */
    int32_t *d_trial_used;         /* [K+1] */
    int32_t N, T, A, K, k_use, reserved;
} spb_scst_args;

int spb_scst_loss(const spb_scst_args *args, spb_stream stream);
int spb_scst_loss_backward(const spb_scst_args *args, const float *d_grad_loss, float *d_grad_probs,
                           float *d_grad_mu, float *d_grad_sigma2, spb_stream stream);

/* ------------------------------------------------------------------------
 * K6/K7 decode: the 16-step ConvLSTM rollout + prediction head for a wave of
 * images, everything after the once-per-image encoder.  Replaces
 * baseline.inference from `state = init_hidden` on
 * (OSIE/models/baseline_attention.py:333-396; AiR/models/baseline_attention.py:392-491;
 * COCO_Search18/models/baseline_attention_multihead.py:339-405) incl. ConvLSTM.forward
 * (:33-48), predict_head.forward (:141-166), the feedback features (:226-236) and the
 * spatial / semantic memory attention (:69-80, :103-116).
 *
 * Weights are passed in PREPARED device layouts (scanpaths_b200/models/
 * baseline_attention.py::prepare_weights builds them from a reference state_dict):
 *   conv weights  fp16 pairs (hi, lo): w * scale = hi + lo / 2^11, rows = output
 *                 channel, K index = (ky*ks + kx)*512 + ci; gate matrices have 2048
 *                 rows ordered [channel block of 64][half of 32][gate i,f,o,g][32].
 * ---------------------------------------------------------------------- */
typedef struct spb_decoder_weights {
    const void *wx_hi, *wx_lo;       /* fp16 [2048, 4608]  lstm.*_x                     */
    const void *wh_hi, *wh_lo;       /* fp16 [2048, 4608]  lstm.*_h                     */
    const void *wp_hi, *wp_lo;       /* fp16 [n_weight_sets*512, 12800]  5x5 layer(s)   */
    const void *ww_hi, *ww_lo;       /* fp16 [24*2048, 512] Winograd F(2x4,3x3)-transformed lstm.*_h: G2 g G4^T,     */
                                     /*   position-major, position = 4*(column position j) + (row position i)   */
    const void *wwx_hi, *wwx_lo;     /* the same for lstm.*_x (the loop-invariant x-gate convolution)              */
    const void *wwx2_hi, *wwx2_lo;   /* fp16 [16*2048, 512] Winograd F(2x2,3x3)-transformed lstm.*_x: G2 g G2^T,   */
                                     /*   position = 4*(column position j) + (row position i): the product path's  */
                                     /*   x-gate convolution                                                       */
    const float *bias_gate;          /* [2048]  b_x + b_h + sum over streams b_m         */
    const float *bias_p;             /* [n_weight_sets*512]                              */
    const float *wm;                 /* [n_streams*3*512*9, 512] rank-1 gate weights:    */
                                     /*   row ((s*3+g)*512+co)*9+tap, col ci             */
    const void *wm_hi, *wm_lo;       /* the same as an fp16 pair (tensor-core route)     */
    const float *w2, *w3;            /* [512] object_head.sal_layer_2 / sal_layer_3      */
    const float *wd1;                /* [49, 512] object_head.drt_layer_1, tap-major     */
    const float *wd2;                /* [2, 48]   object_head.drt_layer_2                */
    const float *w_spatial_embed;    /* [1200, 1200] */
    const float *b_spatial_embed;    /* [1200] */
    const float *w_semantic_embed;   /* [512, 512] */
    const float *b_semantic_embed;   /* [512] */
    const void *wse_hi, *wse_lo;     /* w_semantic_embed as an fp16 pair (tensor-core route) */
    /* composed head (tensor-core path): the 5x5 layer feeds sal_layer_2, sal_layer_3 and drt_layer_1
     * with no nonlinearity in between, so they collapse into effective kernels on h:          */
    const void *w23_hi, *w23_lo;     /* fp16 [n_weight_sets*256, 512] the head as a per-pixel GEMM: rows [0,50) =  */
                                     /*   tap*2 + map of the 5x5 -> (stop map, action map), rows [128,249) = the  */
                                     /*   121 taps of wd_eff variant 0, the rest zero                             */
    const float *b23_eff;            /* [n_weight_sets, 2]  incl. sal_layer_2/3 bias             */
    const float *wd_eff;             /* [n_weight_sets, 4, 121, 512] 11x11 stride-5 duration conv; variant = */
                                     /*   2*(window in top row) + (window in left column): taps of drt_layer_1 */
                                     /*   that fall on the zero padding of the 5x5 output are excluded         */
    const float *bd_eff;             /* [n_weight_sets, 4]  incl. drt_layer_1 bias               */
    const float *w_eff_spatial;      /* [1200] spatial_att: spatial_attention correlated with spatial_lists */
    const float *u_semantic;         /* [512]  semantic_att: semantic_lists^T semantic_attention            */
    float b2, b3, bd1, bd2_mu, bd2_sigma;
    float inv_scale_x, inv_scale_h, inv_scale_p, inv_scale_w, inv_scale_wx, inv_scale_23, inv_scale_m, inv_scale_se,
          inv_scale_wx2;             /* 1 / (power-of-two scale of the fp16 weight pairs) */
    int32_t n_streams;               /* 1 (OSIE, COCO) or 2 (AiR pos / neg)             */
    int32_t n_heads;                 /* 1 or 2 (AiR good / poor)                        */
    int32_t n_weight_sets;           /* 1, 2 (AiR: True, False) or 18 (COCO tasks)      */
    int32_t reserved;
} spb_decoder_weights;

typedef struct spb_decoder_io {
    int32_t n_images, steps;
    int32_t use_tensor_cores;        /* 1 (product path): tcgen05, Winograd F(2x4,3x3) GEMMs for the h-gates,        */
                                     /*   Winograd F(2x2,3x3) for the loop-invariant x-gates, composed head;          */
                                     /* 2: direct 3x3 implicit GEMM for both; 3: Winograd F(2x4) for both (4: with    */
                                     /*   8-k-step accumulators for the x-gates); 5: F(2x4) h-gates + direct x-gates; */
                                     /* 0: SIMT fp32 check kernels with the explicit 5x5 layer                        */
    int32_t reserved;
    const float *d_vf;               /* [N, 512, 30, 40] visual_feature (NCHW, as the encoder emits it) */
    const float *d_att;              /* [N, 1200] initial attention map, or NULL = zeros (OSIE)         */
    const int32_t *d_w_row_base;     /* [N] first row of the 5x5 weight set of each image (task*512), or NULL */
    void *d_workspace;               /* spb_decoder_workspace_bytes() bytes, 1024-byte aligned */
    int64_t workspace_bytes;
    float *d_probs;                  /* [n_heads, N, steps, 1201] */
    float *d_mu, *d_sigma2;          /* [n_heads, N, steps] */
    float *d_action_map;             /* [n_heads, N, steps, 1200] */
} spb_decoder_io;

int64_t spb_decoder_workspace_bytes(int32_t n_images, int32_t n_streams, int32_t n_heads, int32_t steps);
int spb_decode(const spb_decoder_weights *w, const spb_decoder_io *io, spb_stream stream);

/* f3, last layer of the once-per-image encoder: d_vf = relu(sal_conv(d_x)) -- Conv2d(2048, 512, 3, padding 1) on
 * the dilated ResNet-50's [N, 2048, 30, 40] map (OSIE/models/baseline_attention.py:194, :328), as a tcgen05
 * implicit GEMM with fp32-equivalent fp16 operand pairs (w: [512, 9*2048], K index (ky*3+kx)*2048 + ci, prepared
 * like the decoder's conv weights).  d_x, d_vf are channel-major (NCHW) like the reference's tensors; the
 * workspace (spb_sal_conv_workspace_bytes) holds the NHWC operand pairs of d_x. */
int64_t spb_sal_conv_workspace_bytes(int32_t n_images);
int spb_sal_conv(const float *d_x, const void *d_w_hi, const void *d_w_lo, const float *d_bias, float inv_scale,
                 int32_t n_images, void *d_workspace, int64_t workspace_bytes, float *d_vf, spb_stream stream);

/* One implicit-GEMM convolution on its own (unit tests / profiling of the kernel):
 * d_out[(n*1200+p)*ldo + col] = inv_scale * conv_ks(a, w)[p, col] + bias[col]  (ks = 1, 3, 5; operand pairs
 * x = hi + lo/2^11, a NHWC [N,30,40,512], w [rows, ks*ks*512]). */
int spb_conv_gemm(const void *d_a_hi, const void *d_a_lo, const void *d_w_hi, const void *d_w_lo,
                  const int32_t *d_w_row_base, int64_t w_rows, const float *d_bias, float *d_out, int64_t ldo,
                  int32_t n_images, int32_t cols, int32_t ks, float inv_scale, int32_t use_tensor_cores,
                  spb_stream stream);

/* The Winograd F(2x4,3x3) gate GEMMs of the product path on their own (unit tests / profiling), position
 * p = 4j + i with i = 0..3 the row position (F(2,3)) and j = 0..5 the column position (F(4,3)):
 *   m[i][j] = u[p] . w[p]^T  (K = 512),  out[2j] = m[0][j]+m[1][j]+m[2][j],  out[2j+1] = m[1][j]-m[2][j]-m[3][j]
 * d_u  [24][rows_pad][512], d_w [24*cols][512]: fp16 pairs x = hi + lo/2^11; rows_pad, cols % 128 == 0;
 * d_out[((k*cols/128 + col/128)*rows_pad + row)*128 + col%128], k = 0..11, scaled by inv_scale.
 * fine_drain != 0: accumulators are drained every 8 k-steps instead of 32 (less truncation noise, 4x the TMEM reads). */
int spb_wino_gemm(const void *d_u_hi, const void *d_u_lo, const void *d_w_hi, const void *d_w_lo, float *d_out,
                  int64_t rows_pad, int32_t cols, float inv_scale, int32_t fine_drain, spb_stream stream);

/* Compensation of the tensor core's truncating fp32 accumulation (csrc/decoder.cuh): every drained "main"
 * accumulator x becomes x + x * fix.  Default 5.5e-7 (measured on B200, 32 accumulation steps, mixed-sign
 * operands); CudaDecoder re-measures it at construction with a probe through spb_wino_gemm (fix = 0) against
 * an exact float64 product and installs the result here.  Process-wide. */
float spb_get_acc_trunc_fix(void);
int spb_set_acc_trunc_fix(float fix);
float spb_get_acc_trunc_fix_fine(void);          /* the same for the 8-k-step accumulators (spb_wino_gemm fine_drain) */
int spb_set_acc_trunc_fix_fine(float fix);

/* fp32 -> (hi, lo) fp16 pair: x*scale = hi + lo / 2^11.  NCHW [N,C,HW] -> NHWC when `transpose`. */
int spb_split_fp16(const float *d_x, void *d_hi, void *d_lo, int64_t n_outer, int32_t C, int32_t HW,
                   int32_t transpose, float scale, spb_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* SCANPATHS_B200_H */
