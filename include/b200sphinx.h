/* b200sphinx.h -- C ABI of the B200-native acoustic-scoring / HMM-evaluation
 * engine (libb200sphinx.so).
 *
 * Plain pointers and sizes only; no torch, no C++ types.  Every entry point
 * names the reference interface it stands in for (paths relative to the
 * cjac/cmusphinx tree; PS = pocketsphinx/src/libpocketsphinx,
 * SB = sphinxbase/src/libsphinxbase, S3 = sphinx3/src/libs3decoder).
 *
 * Conventions
 *   - return 0 on success, <0 on error (b200_last_error() gives the text);
 *     constructors return NULL on error (the reference's *_init convention,
 *     PS/acmod.c:110-127).
 *   - "host" pointers are ordinary (pinned or pageable) CPU memory; "dev"
 *     pointers are CUDA device memory on the context's device.
 *   - There is NO CPU fallback: every scoring entry point launches sm_100a
 *     kernels and fails with B200_ERR_CUDA if no device is present.
 *   - Senone scores are the reference's negated, shifted (>>10) int16 log
 *     scores, 0 = best (PS/hmm.h:63, PS/acmod.c:1075-1131).
 */
#ifndef B200SPHINX_H
#define B200SPHINX_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_OK            0
#define B200_ERR_ARG      -1
#define B200_ERR_CUDA     -2
#define B200_ERR_IO       -3
#define B200_ERR_UNSUP    -4

#define B200_SENSCR_SHIFT 10                      /* PS/hmm.h:63 */
#define B200_WORST_SCORE  ((int32_t)0xE0000000)   /* PS/hmm.h:74 */
#define B200_TMAT_WORST   (-255)                  /* PS/hmm.h:80 */
#define B200_BAD_SSID     0xffff                  /* PS/bin_mdef.h:96 */
#define B200_MAX_TOPN     8
#define B200_MAX_STREAMS  4

/* ------------------------------------------------------------------ misc */
const char *b200_last_error(void);
/* Library/ABI version and the number of CUDA devices visible (0 if none). */
int  b200_abi_version(void);
int  b200_device_count(void);
/* Number of kernels this library has launched since load (bench.py's
 * gpu_launches claim is read from here). */
long long b200_launch_count(void);

/* ------------------------------------------------------------- logmath */
/* SB/util/logmath.c:61-161  logmath_init(base, shift, use_table=1): writes the
 * log-add table (widened to uint32) and returns its size, or <0. */
int  b200_logadd_table(double base, int shift, uint32_t *out, int max_out);
/* SB/util/logmath.c:446-452 logmath_log(), :391-436 logmath_add() */
int32_t b200_logmath_log(double base, int shift, double p);
int32_t b200_logmath_add(double base, int shift, int32_t x, int32_t y);

/* -------------------------------------------------- load-time precompute */
/* PS/ms_gauden.c:314-359 gauden_dist_precompute: floors raw variances at
 * varfloor, writes det[n_vec] and overwrites var[n_vec*len] with the
 * integer-valued scaled 1/(2 var).  Host arrays. */
int  b200_gauden_precompute(float *var, float *det, long n_vec, int len,
                            float varfloor, double logbase);
/* PS/ms_senone.c:236-258: normalise/floor/renormalise float mixture weights
 * [n_sen][n_feat][n_cw] (modified in place) and quantise to uint8 with the
 * +511 rounding bias, same logical order. */
int  b200_mixw_quantize_ms(float *mixw, uint8_t *out, int n_sen, int n_feat,
                           int n_cw, float mixwfloor, double logbase);
/* PS/ptm_mgau.c:720-742, PS/s2_semi_mgau.c:1155-1177: same normalisation but
 * quantised as -logmath_log(lmath_8b) clamped to [0,159]; output TRANSPOSED
 * to [n_feat][n_cw][n_sen]. */
int  b200_mixw_quantize_tied(float *mixw, uint8_t *out, int n_sen, int n_feat,
                             int n_cw, float mixwfloor, double logbase);
/* PS/tmat.c:275-296: float [n_tmat][n_src][n_src+1] -> uint8 negated
 * log>>10 (255 = "zero"). */
int  b200_tmat_quantize(float *tp, uint8_t *out, int n_tmat, int n_src,
                        double tmatfloor, double logbase);

/* ------------------------------------------------------ S3 binary files */
/* SB/util/bio.c:137-262 + PS/ms_gauden.c:160-298 gauden_param_read.  First
 * call with data=NULL to get dims = {n_mgau, n_feat, n_density, total_floats}
 * and veclen[n_feat]; second call fills data[total_floats]. */
int  b200_s3_read_gauden(const char *path, int32_t dims[4], int32_t *veclen,
                         float *data);
/* PS/ms_senone.c:149-282 header part: dims = {n_sen, n_feat, n_cw, total}. */
int  b200_s3_read_mixw(const char *path, int32_t dims[4], float *data);
/* PS/tmat.c:191-274: dims = {n_tmat, n_src, n_dst, total}. */
int  b200_s3_read_tmat(const char *path, int32_t dims[4], float *data);
/* PS/s2_semi_mgau.c:888-1089 read_sendump.  On input dims[0..2] hold the
 * model's {n_feat, n_density, n_sen} (defaults when the file does not name
 * them); on output dims = {n_feat, n_density, n_sen, n_clust (0 if none),
 * row_bytes}.  mixw (NULL = query only) receives [n_feat][n_density][row_bytes]
 * raw rows (4-bit packed, even senone in the low nibble, when cluster_bits is
 * 4), cb the 16 cluster values when n_clust != 0. */
int  b200_s3_read_sendump(const char *path, int32_t dims[5], uint8_t *mixw,
                          uint8_t cb[16]);

/* Model-definition maps (PS/bin_mdef.c:330-507 binary "BMDF", PS/mdef.c:505-602
 * text 0.3): dims = {n_sen, n_ci_sen, n_ciphone, n_emit_state};
 * sen2cimap[n_sen] = CI phone owning each senone (the ptm senone -> codebook
 * map, PS/ptm_mgau.c:836-848), cd2cisen[n_sen] = CI senone at the same state
 * position (sphinx3's mdef_t.cd2cisen).  Either pointer may be NULL. */
int  b200_mdef_read_maps(const char *path, int32_t dims[4], int16_t *sen2cimap,
                         int16_t *cd2cisen);

/* PS/acmod.c:349-361,885-982: senone-score dump files (`-senlogdir` output,
 * `-senin yes` / ps_decode_senscr input).  Write: scores [n_frames][n_sen]; with
 * active == NULL every frame is written dense, else frame t carries the
 * n_active[t] uint8 deltas active[t] and only those scores.  Read: first call
 * with scores == NULL fills dims = {n_sen, n_frames}; second call (dims[1] =
 * capacity in frames) fills scores [n_frames][n_sen] (unlisted senones =
 * SENSCR_DUMMY 0x7fff) and n_active [n_frames]. */
int  b200_sen_write(const char *path, const char *mdef_file, int n_sen, double logbase,
                    const int16_t *scores, int n_frames,
                    const uint8_t *const *active, const int32_t *n_active);
int  b200_sen_read(const char *path, int32_t dims[2], double *logbase, int16_t *scores,
                   int32_t *n_active);

/* ============================================================ GMM scoring
 * One handle type for the three pocketsphinx back-ends.  The handle owns a
 * device-resident copy of the (precomputed) parameters.
 */
typedef struct b200_mgau b200_mgau_t;

typedef struct {
    int32_t n_mgau, n_feat, n_density, n_sen;
    int32_t featlen[B200_MAX_STREAMS];
    int32_t topn;       /* -topn  (PS cmdln_macro.h:351) */
    int32_t aw;         /* -aw    (:331), ms only */
    int32_t ds_ratio;   /* -ds    (:347): s2_semi implements it, ms ignores it (as the reference), ptm refuses > 1 */
    double  logbase;    /* -logbase (:371) */
    int32_t device;     /* CUDA device ordinal */
    /* -topn_beam (PS cmdln_macro.h:355), s2_semi only: per stream, the list is cut at the
     * first normalised score above the beam (mgau_norm, PS/s2_semi_mgau.c:189-207); 0 = off.
     * Values as split_topn leaves them (uint8 range; PS/s2_semi_mgau.c:1204-1231). */
    int32_t topn_beam[B200_MAX_STREAMS];
} b200_mgau_cfg_t;

/* ms back-end: PS/ms_mgau.c:79-141 ms_mgau_init.  mean/var/det are the
 * PRECOMPUTED arrays [mgau][feat][density][featlen f] / [mgau][feat][density]
 * (b200_gauden_precompute), mixw the ms-quantised uint8 [sen][feat][cw],
 * sen2mgau[n_sen] the senone->codebook map (PS/ms_senone.c:297-340). */
b200_mgau_t *b200_ms_create(const b200_mgau_cfg_t *cfg, const float *mean,
                            const float *var, const float *det,
                            const uint8_t *mixw, const uint32_t *sen2mgau);
/* Convenience mirroring ms_mgau_init(config): reads S3 files, precomputes on
 * the host exactly as the reference does, uploads.  senmgau is ".cont.",
 * ".semi." or ".ptm." (then sen2cb must be given), as PS/ms_senone.c:297-340. */
b200_mgau_t *b200_ms_load(const char *meanfile, const char *varfile,
                          const char *mixwfile, const char *senmgau,
                          const uint8_t *sen2cb, double varfloor,
                          double mixwfloor, int topn, int aw, double logbase,
                          int device);

/* ptm / s2_semi back-ends: PS/ptm_mgau.c:775-872 ptm_mgau_init,
 * PS/s2_semi_mgau.c:1240-1330 s2_semi_mgau_init.  mixw is
 * [n_feat][n_density][row_bytes]: 8-bit rows of n_sen bytes, or (n_clust!=0)
 * 4-bit packed rows of (n_sen+1)/2 bytes with the 16-entry cluster codebook
 * mixw_cb.  sen2cb[n_sen] (ptm only; NULL for s2_semi = one codebook). */
b200_mgau_t *b200_ptm_create(const b200_mgau_cfg_t *cfg, const float *mean,
                             const float *var, const float *det,
                             const uint8_t *mixw, int n_clust,
                             const uint8_t *mixw_cb, const uint8_t *sen2cb);
b200_mgau_t *b200_semi_create(const b200_mgau_cfg_t *cfg, const float *mean,
                              const float *var, const float *det,
                              const uint8_t *mixw, int n_clust,
                              const uint8_t *mixw_cb);

/* ps_mgaufuncs_t.free (PS/acmod.h:111) */
void b200_mgau_free(b200_mgau_t *m);
/* ps_mgaufuncs_t.name: "b200_ms" | "b200_ptm" | "b200_semi" */
const char *b200_mgau_name(const b200_mgau_t *m);
int  b200_mgau_n_sen(const b200_mgau_t *m);
int  b200_mgau_featdim(const b200_mgau_t *m);   /* sum of stream lengths */

/* ps_mgaufuncs_t.transform (PS/acmod.h:108; gauden_mllr_transform
 * PS/ms_gauden.c:551-605): replace the device parameters by freshly
 * transformed + precomputed host arrays. */
int  b200_mgau_update_params(b200_mgau_t *m, const float *mean,
                             const float *var, const float *det);

/* Which kernel family scores dense batches for the ms back-end:
 *   0 = exact CUDA-core path (sequential f32, bit-exact vs the reference),
 *   1 = tcgen05 tensor-core Mahalanobis GEMM (TF32x3 split, fused top-N /
 *       log-add epilogue; .cont. single-stream models only).
 * Default: 1 when the model shape allows it, else 0. */
/* For the ptm / s2_semi back-ends path 1 runs the codebook stage (eval_topn +
 * eval_cb, PS/ptm_mgau.c:98-231, PS/s2_semi_mgau.c:80-187) as the same
 * tensor-core GEMM followed by an exact float32 re-scoring of the 8 best
 * candidates per (frame, codebook); pairs for which exactness cannot be proven
 * (integer-score ties near rank N) are re-done with the reference's literal
 * scan, so both paths return identical top-N lists and scores.  Needs
 * n_density % 256 == 0 and topn <= 4; default 1 when available. */
int  b200_mgau_set_path(b200_mgau_t *m, int path);
int  b200_mgau_get_path(const b200_mgau_t *m);
/* ms tensor-core path: operand format of the last scoring call, decided per
 * n-tile (256 Gaussians) and per batch -- 1 = every tile on fp16 hi/lo pairs
 * (kind::f16, the default), 0 = every tile on TF32 hi/lo pairs (no fp16 operand
 * for this model, B200_TC_F16=0, or the batch's features exceed every tile's
 * scaled fp16 range), 2 = mixed (tiles with very sharp Gaussians, or whose
 * limits this batch exceeds, ran on TF32), -1 = no tensor-core plan.
 * Synchronises. */
int  b200_mgau_tc_last_format(b200_mgau_t *m);
/* Tied back-ends, tensor-core path: {(frame, codebook, stream) lists produced,
 * lists that went through the exact-scan fallback, largest |GEMM - exact|
 * distance observed on a best candidate in raw log units (0 if <= 4)} of the
 * last scoring call. */
int  b200_mgau_tied_stats(b200_mgau_t *m, long long out[3]);
/* ms tensor-core path (fully continuous models), last scoring call:
 * {frame x senone pairs scored, pairs that took the full top-N network (the
 * rest were settled by the single-density certificate), pairs whose fden had to
 * be re-derived from the reference's float32 distance (queue A), pairs whose
 * top-N set or order was re-derived by the literal scan (queue B), 1 if a queue
 * overflowed and every pair was redone by the literal scan, largest
 * |GEMM - reference| distance seen on a re-scored density in raw log units (0
 * if <= 2; the certificates assume eps0 + (|d| >> eps_shift), see
 * B200_TC_EPS0 / B200_TC_EPS_SHIFT), hard pairs finished by the finding lane
 * because the tile's queue was full}.  Every score of this path is bit-identical
 * to ms_cont_mgau_frame_eval's as long as the last entry stays below that
 * bound.  Synchronises. */
int  b200_mgau_cont_stats(b200_mgau_t *m, long long out[7]);

/* Batched dense scoring == ps_mgau_frame_eval(..., compallsen=1) for frames
 * 0..T-1 (PS/ms_mgau.c:162-205, PS/ptm_mgau.c:405-450,
 * PS/s2_semi_mgau.c:840-886): feat [T][featdim] float32 (streams
 * concatenated as in acmod's feat_buf, SB/feat/feat.c:519-549),
 * out [T][n_sen] int16.  Host version copies in/out (pinned staging inside). */
int  b200_mgau_score_host(b200_mgau_t *m, const float *feat, int T, int16_t *out);
/* Same with device-resident input/output on `stream` (a cudaStream_t cast to
 * void*, NULL = default stream); asynchronous. */
int  b200_mgau_score_dev(b200_mgau_t *m, const float *d_feat, int T,
                         int16_t *d_out, void *stream);

/* Drop-in for one ps_mgaufuncs_t.frame_eval call (PS/acmod.h:99-105): feat is
 * the reference's mfcc_t** (one pointer per stream), senone_active the uint8
 * delta list of PS/acmod.c:1219-1271.  With compallsen==0 only active entries
 * of senscr are meaningful (ms) / others are zero-minus-best (ptm) / zero
 * (s2_semi), as in the reference. */
int  b200_mgau_frame_eval(b200_mgau_t *m, int16_t *senscr,
                          const uint8_t *senone_active, int32_t n_senone_active,
                          const float *const *feat, int32_t frame,
                          int32_t compallsen);

/* Utterance-batched variant used by the plug-in: score all T frames of an
 * utterance once, then serve frame_eval calls from the cache with the caller's
 * active list.
 *   ms      : the dense un-normalised rows are copied to pinned host memory at
 *             utt_begin; utt_frame applies the active-set-dependent
 *             normalisation (PS/ms_mgau.c:226-248) on the host -- no GPU work
 *             per frame;
 *   s2_semi : the final scores (normalised per stream, independent of the active
 *             set) are copied to the host at utt_begin; utt_frame copies / masks
 *             a row -- no GPU work per frame;
 *   ptm     : the top-N lists stay on the device; utt_frame runs the mixing stage
 *             there, because its normalisation depends on which codebooks the
 *             active senones touch (PS/ptm_mgau.c:267-288,390-397). */
int  b200_mgau_utt_begin(b200_mgau_t *m, const float *feat, int T);
/* Same, for a block of frames that starts at utterance frame `frame0` (live
 * mode hands the decoder's buffered frames over piecewise).  Only s2_semi with
 * -ds > 1 looks at the number: frames that are not a multiple of ds_ratio
 * re-score the previous frame's codewords (PS/s2_semi_mgau.c:176-186), so a
 * block starting on such a frame continues from the last frame of the block
 * before it. */
int  b200_mgau_utt_begin_at(b200_mgau_t *m, const float *feat, int T, int frame0);
int  b200_mgau_utt_frame(b200_mgau_t *m, int16_t *senscr,
                         const uint8_t *senone_active, int32_t n_senone_active,
                         int32_t frame, int32_t compallsen);

/* Timing of the last score_* call's dominant kernel(s) in milliseconds
 * (CUDA events on the launching stream). which: 0 = total device time,
 * 1 = operand-prep kernel, 2 = main scoring kernel(s), 3 = normalise kernel;
 * b200_mgau_timing_avg also takes 4 = the exact fix-up kernels of the tensor-core
 * ms path (the tail of section 2; 2 minus 4 = the tcgen05 score kernel alone). */
float b200_mgau_last_ms(const b200_mgau_t *m, int which);
/* Average of the same quantity over the last n_calls (<= 64) score_dev calls,
 * including asynchronous ones on a caller's stream; the caller must have
 * synchronised that stream.  <0 on error. */
float b200_mgau_timing_avg(b200_mgau_t *m, int n_calls, int which);

/* ======================================================= HMM evaluation
 * Batched hmm_vit_eval (PS/hmm.c:224-807) over a structure-of-arrays
 * population, plus the active-senone gather (PS/acmod.c:1178-1271,
 * PS/ngram_search_fwdtree.c:518-555) and the beam/compaction step
 * (PS/ngram_search_fwdtree.c:714-869 beam test only).
 */
typedef struct b200_hmmctx b200_hmmctx_t;

/* hmm_context_init (PS/hmm.c:55-77): tp [n_tmat][n_emit][n_emit+1] uint8,
 * sseq [n_sseq][n_emit] uint16.  n_emit must be 3 or 5. */
b200_hmmctx_t *b200_hmm_ctx_create(int n_emit, const uint8_t *tp, int n_tmat,
                                   const uint16_t *sseq, int n_sseq, int n_sen,
                                   int device);
void b200_hmm_ctx_free(b200_hmmctx_t *c);

/* The SoA population (all arrays length n_hmm unless noted; state-major:
 * score[st*n_hmm + i]).  Same fields as hmm_t (PS/hmm.h:156-173). */
typedef struct {
    int32_t  n_hmm;
    int32_t *score;        /* [n_emit][n_hmm] */
    int32_t *history;      /* [n_emit][n_hmm] */
    int32_t *out_score;    /* [n_hmm] */
    int32_t *out_history;  /* [n_hmm] */
    uint16_t *senid;       /* [n_emit][n_hmm]: senone ids, or ssids if mpx */
    int16_t *tmatid;       /* [n_hmm] */
    uint8_t *mpx;          /* [n_hmm] */
    int32_t *bestscore;    /* [n_hmm] out */
} b200_hmm_soa_t;

/* AoS <-> SoA: the reference's hmm_t (PS/hmm.h:156-173; 80 bytes on LP64,
 * embedded as the first member of chan_t / root_chan_t, PS/ngram_search.h:64-104)
 * as this library reads it.  The binding checks sizeof / offsetof against the
 * real header at compile time (plugin/b200_hmm.c). */
typedef struct {
    void    *ctx;
    int32_t  score[5];
    int32_t  history[5];
    int32_t  out_score;
    int32_t  out_history;
    uint16_t ssid;
    uint16_t senid[5];
    int32_t  bestscore;
    int16_t  tmatid;
    int16_t  frame;
    uint8_t  mpx;
    uint8_t  n_emit_state;
} b200_ps_hmm_t;
/* Gather n HMMs (an array of hmm_t pointers, e.g. the channels evaluate_channels
 * walks, PS/ngram_search_fwdtree.c:598-691) into a caller-allocated SoA with
 * room for n HMMs (arrays of n_emit * n resp. n entries); sets soa->n_hmm. */
int  b200_hmm_pack(const void *const *hmms, int n, int n_emit, b200_hmm_soa_t *soa);
/* Scatter the fields hmm_vit_eval writes (score, history, out_score, out_history,
 * bestscore, and the per-state ssids of mpx HMMs) of SoA entry i back into the
 * hmm_t. */
void b200_hmm_unpack_one(const b200_hmm_soa_t *soa, int n_emit, int i, void *hmm);
int  b200_hmm_unpack(const b200_hmm_soa_t *soa, int n_emit, void *const *hmms);

/* Host round trip: upload SoA, run n_frames steps with senscr[f] =
 * senscr + f*n_sen (hmm_context_set_senscore per frame), download.
 * best_out[n_frames] gets max bestscore per frame. */
int  b200_hmm_eval_host(b200_hmmctx_t *c, b200_hmm_soa_t *h,
                        const int16_t *senscr, int n_frames, int32_t *best_out);

/* Device-resident population for benchmarking / resident search state. */
int  b200_hmm_pop_upload(b200_hmmctx_t *c, const b200_hmm_soa_t *h);
int  b200_hmm_pop_download(b200_hmmctx_t *c, b200_hmm_soa_t *h);
/* Device addresses of the resident population (state stride = n_hmm) and the stream its kernels
 * run on: for device-side stages chained behind the step kernels (b200_fwdtree_prune_dev). */
int  b200_hmm_pop_device(b200_hmmctx_t *c, b200_hmm_soa_t *dev, void **stream);
/* Batched search state: split the resident population into n_utt utterances,
 * utterance u owning HMMs [utt_off[u], utt_off[u+1]) (utt_off[0] = 0,
 * utt_off[n_utt] = n_hmm).  Afterwards b200_hmm_step_* take n_utt rows of
 * senone scores ([n_utt][n_sen]) and produce per-utterance best scores, beam
 * thresholds, survivor counts and active-senone masks; the survivor list is
 * ordered by (utterance, HMM index).  pop_upload resets to one utterance. */
int  b200_hmm_pop_set_utts(b200_hmmctx_t *c, int n_utt, const int32_t *utt_off);
/* One frame on the resident population: hmm_vit_eval for every HMM, frame
 * best (max), then beam test bestscore > best + beam
 * (PS/ngram_search_fwdtree.c:741,800) -> keep[n_hmm] uint8 flags and an
 * ORDER-PRESERVING compacted index list; then active-senone bitmask of the
 * survivors (acmod_activate_hmm) -> mask[(n_sen+31)/32].
 * d_senscr: device int16[n_sen].  Results stay on the device; fetch with
 * b200_hmm_step_results. */
/* n_frames consecutive b200_hmm_step_dev calls; frame f reads the senone scores at
 * d_senscr + (f % n_cycle) * frame_stride (int16 elements).  Same results as the
 * single steps, issued as ONE persistent cooperative launch for the whole run
 * (phases separated by barriers among the CTAs of an utterance; a population of at
 * most one HMM per resident thread -- e.g. one utterance x 50 000 HMMs -- keeps its
 * state in registers from the first frame to the last).  The results of the LAST
 * frame are read with b200_hmm_step_results. */
int  b200_hmm_run_dev(b200_hmmctx_t *c, const int16_t *d_senscr, long frame_stride,
                      int n_cycle, int n_frames, int32_t beam, void *stream);
/* (n_emit_state may be 1..5 = HMM_MAX_NSTATE: 3 and 5 run the reference's unrolled
 * specialisations, 1, 2 and 4 hmm_vit_eval_anytopo, PS/hmm.c:711-786.) */
int  b200_hmm_step_dev(b200_hmmctx_t *c, const int16_t *d_senscr, int32_t beam,
                       void *stream);
int  b200_hmm_step_results(b200_hmmctx_t *c, int32_t *best /* [n_utt] */, int32_t *n_keep /* [n_utt] */,
                           int32_t *keep_idx /* [n_hmm] or NULL */,
                           uint32_t *sen_mask /* [n_utt][(n_sen+31)/32] or NULL */);
/* Host-input convenience for tests: senscr on the host. */
int  b200_hmm_step_host(b200_hmmctx_t *c, const int16_t *senscr, int32_t beam);
float b200_hmm_last_ms(const b200_hmmctx_t *c);

/* Maintenance of the resident population (PS/hmm.c:169-218), batched:
 *  - hmm_normalize: every state / exit score BETTER_THAN WORST_SCORE loses its
 *    utterance's value d_best_per_utt[u] (NULL: the best score of the last step),
 *    as renormalize_scores does (PS/ngram_search_fwdtree.c:560-594);
 *  - hmm_clear_scores for the HMMs the last beam step dropped (the `else` arm of
 *    prune_nonroot_chan, PS/ngram_search_fwdtree.c:864-866);
 *  - hmm_enter for a list of (HMM index, score, history id) with the test its
 *    callers make first (`score BETTER_THAN hmm_in_score`, :757, :846): same
 *    result as walking the list in order -- the best score wins, the first of
 *    equal scores keeps its history, nothing changes if the resident in-score
 *    is not beaten. */
int  b200_hmm_normalize_dev(b200_hmmctx_t *c, const int32_t *d_best_per_utt, void *stream);
int  b200_hmm_clear_pruned_dev(b200_hmmctx_t *c, void *stream);
int  b200_hmm_enter_dev(b200_hmmctx_t *c, const int32_t *d_idx, const int32_t *d_score,
                        const int32_t *d_hist, int n, void *stream);
int  b200_hmm_enter_host(b200_hmmctx_t *c, const int32_t *idx, const int32_t *score,
                         const int32_t *hist, int n);

/* ----------------------------------------------------------------------------
 * Prune / phone-transition stage of the forward tree search:
 * prune_root_chan followed by prune_nonroot_chan, PS/ngram_search_fwdtree.c:714-790 and
 * :792-869 (they are `static`; their only caller is prune_channels, :1125-1160), with
 * hmm_enter (PS/hmm.c:197-203) and hmm_clear_scores (PS/hmm.c:169-181), for a batch of
 * utterances that share one lexical tree.  Results -- channel states, the ORDER of the next
 * frame's active list and the order of the last-phone candidates -- are identical to the
 * reference's sequential walk (csrc/fwdtree_prune.cu explains how).
 *
 * The tree (root_chan_t / chan_t, PS/ngram_search.h:64-104) as arrays: channels
 * 0..n_root-1 are ngs->root_chan[]; the children of channel c -- its `next` channel and
 * that channel's `alt` chain, in chain order -- are child[child_off[c] .. child_off[c+1]);
 * ciphone[c] is chan_t.ciphone (the phone-loop look-ahead is indexed with it, :748, :827);
 * the words whose last phone follows c -- penult_phn_wid and its homophone_set chain
 * (:765-766, :847-848) -- are pw_wid[pw_off[c] .. pw_off[c+1]) with
 * pw_lastphone = dict_last_phone(w).  Every non-root channel must have exactly one
 * parent (else NULL + b200_last_error).  n_emit = hmm_n_emit_state. */
typedef struct b200_chantree b200_chantree_t;
b200_chantree_t *b200_chantree_create(int n_root, int n_chan, const int32_t *child_off, const int32_t *child,
                                      const int32_t *ciphone, const int32_t *pw_off, const int32_t *pw_wid,
                                      const int32_t *pw_lastphone, int n_ci, int n_emit, int device);
void b200_chantree_free(b200_chantree_t *t);
int  b200_chantree_cand_cap(const b200_chantree_t *t);   /* the most candidates one frame can produce (= n_pw) */

/* Per utterance and frame: par[8] = {frame_idx, ngs->best_score, ngs->dynamic_beam, ngs->pbeam,
 * ngs->lpbeam, ngs->pip, ngs->nwpen, pls != NULL} (:725-730), pls_pen[n_ci] =
 * phone_loop_search_score(pls, ci) (PS/phone_loop_search.h:104; may be NULL when no utterance
 * has a look-ahead), acl = ngs->active_chan_list[frame_idx & 1] as channel ids, n_act its length.
 * State: the hmm_t fields of every channel, state-major like b200_hmm_soa_t
 * (score / history [n_emit][n_utt * n_chan], the others [n_utt * n_chan]) plus hmm_frame.
 * Out: nacl / n_nacl = active_chan_list[(frame_idx + 1) & 1] in the reference's order,
 * cand = ngs->lastphn_cand {wid, score, bp} in the reference's order (last_phone_transition,
 * :876, consumes it on the host), the states updated in place.
 * list_cap >= the longest list in or out (n_chan - n_root always suffices); cand_cap >=
 * b200_chantree_cand_cap().  The device form takes device pointers and a stream
 * (state_stride = the distance between two states' rows, >= n_utt * n_chan: the arrays may be
 * the resident population of a b200_hmmctx_t). */
typedef struct {
    int32_t *score, *history, *out_score, *out_history, *bestscore, *frame;
    long state_stride;
    const int32_t *par;        /* [n_utt][8] */
    const int32_t *pls_pen;    /* [n_utt][n_ci] or NULL */
    const int32_t *acl;        /* [n_utt][list_cap] */
    const int32_t *n_act;      /* [n_utt] */
    int32_t list_cap;
    int32_t *nacl, *n_nacl;    /* [n_utt][list_cap], [n_utt] */
    int32_t *cand, *n_cand;    /* [n_utt][cand_cap][3], [n_utt] */
    int32_t cand_cap;
} b200_prune_dev_t;
/* The stage in front of it, on the same lists: eval_root_chan + eval_nonroot_chan
 * (PS/ngram_search_fwdtree.c:598-634) -- hmm_vit_eval for the root channels whose hmm_frame is
 * the current frame (d_par[u][0]) and for every entry of the active list, on the RESIDENT
 * population of `c` (n_utt copies of the tree's n_chan channels: b200_hmm_pop_upload), senone
 * scores d_senscr [n_utt][n_sen]; d_best[u] = the best hmm_vit_eval result of utterance u
 * (WORST_SCORE when nothing is active).  All pointers are device pointers.  Evaluate and prune
 * alternate on the device: the list b200_fwdtree_prune_dev writes is the next frame's d_acl. */
int  b200_hmm_eval_list_dev(b200_hmmctx_t *c, int n_root, int n_chan, const int32_t *d_frame, const int32_t *d_par,
                            const int32_t *d_acl, const int32_t *d_n_act, int list_cap, const int16_t *d_senscr,
                            int32_t *d_best, void *stream);
int  b200_fwdtree_prune_dev(b200_chantree_t *t, int n_utt, const b200_prune_dev_t *d, void *stream);
/* The two other places where ngram_fwdtree_search touches the tree's channels, on the same arrays
 * (only score / out_score / bestscore / frame / par / acl / n_act / list_cap of `d` are read):
 * renormalize_scores' tree part (PS/ngram_search_fwdtree.c:557-576: hmm_normalize of the roots stamped
 * with the current frame and of the active list, by d_norm[u]) and deactivate_channels' root loop
 * (:1418-1431: hmm_clear_scores of the roots still stamped with the current frame). */
int  b200_fwdtree_renorm_dev(b200_chantree_t *t, int n_utt, const b200_prune_dev_t *d, const int32_t *d_norm, void *stream);
int  b200_fwdtree_deactivate_dev(b200_chantree_t *t, int n_utt, const b200_prune_dev_t *d, void *stream);
int  b200_fwdtree_prune_host(b200_chantree_t *t, int n_utt, const int32_t *par, const int32_t *pls_pen,
                             const int32_t *acl, const int32_t *n_act, int list_cap, int32_t *score,
                             int32_t *history, int32_t *out_score, int32_t *out_history, int32_t *bestscore,
                             int32_t *frame, int32_t *nacl, int32_t *n_nacl, int32_t *cand, int32_t *n_cand,
                             int cand_cap);

/* ----------------------------------------------------------------------------
 * The phone-loop look-ahead search (-pl_window > 0; another caller of hmm_vit_eval,
 * PS/phone_loop_search.c:201): phone_loop_search_step (:253-291) minus its acmod calls --
 * renormalize_hmms (:171-184), evaluate_hmms (:186-210), prune_hmms (:212-233),
 * phone_transition (:235-268) -- for n_utt utterances in lock step, the phone HMMs resident on
 * the device; its product is what the forward tree search reads through
 * phone_loop_search_score (PS/phone_loop_search.h:104): pls_pen[u][ci] =
 * hmm_bestscore(phone ci) - best_score, the pls_pen input of b200_fwdtree_prune_*.
 * create = phone_loop_search_reinit (:66-106): n_phones = bin_mdef_n_ciphone, non-mpx HMMs,
 * senid [n_emit][n_phones] = sseq[pid2ssid(ci)], tmatid = pid2tmatid(ci), beam / pbeam / pip =
 * logmath_log of -pl_beam / -pl_pbeam / -pip; start = phone_loop_search_start (:153-169).
 * State arrays are state-major: score / history [n_emit][n_utt * n_phones], the rest
 * [n_utt * n_phones]; best [n_utt] = pls->best_score; renorm [n_utt] = the norm the last step
 * applied (0: none; the reference keeps them in pls->renorm). */
typedef struct b200_phoneloop b200_phoneloop_t;
b200_phoneloop_t *b200_phone_loop_create(int n_phones, int n_emit, const uint8_t *tp, int n_tmat, const uint16_t *senid,
                                         const int16_t *tmatid, int n_sen, int32_t beam, int32_t pbeam, int32_t pip,
                                         int n_utt, int device);
void b200_phone_loop_free(b200_phoneloop_t *h);
int  b200_phone_loop_start(b200_phoneloop_t *h);
int  b200_phone_loop_set_state(b200_phoneloop_t *h, const int32_t *score, const int32_t *history, const int32_t *out_score,
                               const int32_t *out_history, const int32_t *bestscore, const int32_t *frame, const int32_t *best);
int  b200_phone_loop_get_state(b200_phoneloop_t *h, int32_t *score, int32_t *history, int32_t *out_score, int32_t *out_history,
                               int32_t *bestscore, int32_t *frame, int32_t *best, int32_t *renorm);
/* One frame.  senscr [n_utt][n_sen] int16 (device pointer for _dev); pls_pen [n_utt][n_phones] or NULL. */
int  b200_phone_loop_step_dev(b200_phoneloop_t *h, const int16_t *d_senscr, int frame_idx, int32_t *d_pls_pen, void *stream);
int  b200_phone_loop_step_host(b200_phoneloop_t *h, const int16_t *senscr, int frame_idx, int32_t *pls_pen, int32_t *best);

/* acmod_flags2list (PS/acmod.c:1219-1271): bitmask -> uint8 delta list with
 * the reference's lossy >255 bridging.  Host utility; returns n written. */
int  b200_flags2list(const uint32_t *mask, int n_sen, uint8_t *deltas, int max_out);

/* ================================================= sphinx3 GMM scoring
 * sphinx3's flavour of the path (S3/libam/approx_cont_mgau.c, cont_mgau.c):
 * one mixture per senone, float64 Mahalanobis accumulation, int32 log scores
 * (base -logbase, default 1.0003, unshifted; HIGHER = better), CI-senone beam
 * with best-Gaussian / CI back-off, optional frame down-sampling.  Stands in
 * for srch_funcs_t.gmm_compute_lv1 / gmm_compute_lv2 (sphinx3/include/srch.h:
 * 590-612 -> S3/libsearch/gmm_wrap.c:80-214).  Diagonal covariances, no
 * Gaussian-selection / sub-VQ shortlists.
 */
typedef struct b200_s3mgau b200_s3mgau_t;

/* mgau_init(meanfile, varfile, varfloor, mixwfile, mixwfloor, precomp=1,
 * ".cont.", MIX_INT_FLOAT_COMP, logmath) -- S3/libam/cont_mgau.c:900-958 --
 * on arrays: mean/var [n_sen][n_comp][veclen] RAW values as stored in the
 * files, mixw [n_sen][n_comp] raw counts.  Does mixw normalisation,
 * mgau_uninit_compact, mgau_var_floor and mgau_precomp on the host exactly as
 * the reference.  cd2cisen[n_sen] / n_ci_sen come from the mdef
 * (sphinx3/include/mdef.h:188-201; CI senones first). */
b200_s3mgau_t *b200_s3_create(int n_sen, int n_comp, int veclen, const float *mean,
                              const float *var, const float *mixw, double varfloor,
                              double mixwfloor, double logbase,
                              const int32_t *cd2cisen, int n_ci_sen, int device);
/* Same from the S3 binary files (mgau_file_read :160-400, mgau_mixw_read :480-680). */
b200_s3mgau_t *b200_s3_load(const char *meanfile, const char *varfile,
                            const char *mixwfile, double varfloor, double mixwfloor,
                            double logbase, const int32_t *cd2cisen, int n_ci_sen,
                            int device);
void b200_s3_free(b200_s3mgau_t *m);           /* mgau_free */
/* dims = {n_sen, max_comp, veclen, n_ci_sen, ci_pbeam (log)} */
int  b200_s3_dims(const b200_s3mgau_t *m, int32_t dims[5]);
/* fast_gmm_init (S3/libam/fast_algo_struct.c:420-467): -ci_pbeam (a
 * probability), -maxcdsenpf, -ds, -tighten_factor. */
int  b200_s3_set_fast(b200_s3mgau_t *m, double ci_pbeam, int max_cd, int ds_ratio,
                      float tighten_factor);
/* The same with the beams as the logs3 integers fast_gmm_t holds (S3 include/fast_algo_struct.h:204-262:
 * fg->gmms->ci_pbeam, ->max_cd, ->tighten_factor, fg->downs->ds_ratio, fg->gaus->subvqbeam) -- what a binding
 * inside the decoder (plugin/b200_s3_mgau.c) has at hand. */
int  b200_s3_set_fast_log(b200_s3mgau_t *m, int32_t ci_pbeam_log, int max_cd, int ds_ratio, float tighten_factor,
                          int32_t subvqbeam_log);
/* -subvq FILE -svmax N -vqeval N -subvqbeam P: sub-vector quantised Gaussian selection for the approx path --
 * subvq_init, S3/libam/subvq.c:206-373 (file format, variance floor, vector_maha_precomp, map compaction and
 * linearisation), subvq_gautbl_eval_logs3 :488-506, subvq_mgau_shortlist :383-468 and approx_mgau_eval's use of
 * the shortlist incl. its full re-evaluation below S3_LOGPROB_ZERO + 100000 (approx_cont_mgau.c:187-284).  On the
 * device the shortlist is a mask over the dense component scores.  -svq4svq (quantised scores AS the Gaussian
 * scores) is not implemented.
 * file == NULL removes the layer.  The model must have been created with the same component layout the file
 * was made for (the reference's own check: #valid components per mixture). */
int  b200_s3_set_subvq(b200_s3mgau_t *m, const char *file, double varfloor, int max_sv, int vqeval, double subvqbeam);
/* -gs FILE: the Gaussian selector -- gs_read (S3/libam/gs.c:156-218), gc_compute_closest_cw (:221-259: the frame's
 * nearest codeword) and gs_mgau_shortlist (:263-300: the bit map of (mixture, codeword), every component when it
 * is empty), used by approx_mgau_eval ahead of the sub-VQ shortlist (approx_cont_mgau.c:207-212).  Refused: more
 * than 32 densities (the reference keeps one 32-bit word) and an odd codeword count (its pairwise loop reads past
 * the table).  The reference asserts best_cid > 0 (:209), i.e. aborts a debug build when codeword 0 is nearest;
 * this implementation -- like a release build -- goes on.  file == NULL removes the layer. */
int  b200_s3_set_gs(b200_s3mgau_t *m, const char *file);
/* per-utterance reset of bstidx / updatetime (S3/libsearch/srch_time_switch_tree.c:484-490) */
int  b200_s3_utt_reset(b200_s3mgau_t *m);
/* Host copies of the precomputed parameters in the reference's order, padded
 * to max_comp: n_comp[n_sen], mean/var [n_sen][max_comp][veclen] (var =
 * 1/(2 var)), lrd/mixw [n_sen][max_comp], scal = {distfloor, 1/ln(base)}.
 * Any pointer may be NULL. */
int  b200_s3_params(const b200_s3mgau_t *m, int32_t *n_comp, float *mean, float *var,
                    float *lrd, int32_t *mixw, double scal[2]);
/* mgau_t.bstidx / updatetime of every senone (cont_mgau.h:174-176) */
int  b200_s3_state(b200_s3mgau_t *m, int32_t *bstidx, int32_t *updatetime);

/* Dense scoring: out[t][s] = mgau_eval(g, s, NULL, feat[t], t, 1)
 * (cont_mgau.c:1171-1205) for every senone, un-normalised.  [T][veclen] float32
 * in, [T][n_sen] int32 out. */
int  b200_s3_dense_host(b200_s3mgau_t *m, const float *feat, int T, int32_t *out);
int  b200_s3_dense_dev(b200_s3mgau_t *m, const float *d_feat, int T, int32_t *d_out,
                       void *stream);

/* One utterance (or a run of T consecutive frames starting at frame number
 * frame0): per frame approx_cont_mgau_ci_eval (approx_cont_mgau.c:368-431) then
 * approx_cont_mgau_frame_eval (:433-616).  sen_active [T][n_sen] uint8 is
 * ascr_t.sen_active per frame, in/out (CI entries are set to 1); NULL = every
 * senone active.  senscr_io [n_sen] is ascr_t.senscr before the first frame /
 * after the last one (entries of inactive senones keep their old value, as in
 * the reference); NULL = continue from the handle's own copy.  out
 * [T][n_sen] = ascr_t.senscr after each frame, best [T] = the return values
 * (srch_t.senscale). */
int  b200_s3_score_utt_host(b200_s3mgau_t *m, const float *feat, int T, int frame0,
                            uint8_t *sen_active, int32_t *senscr_io, int32_t *out,
                            int32_t *best);
int  b200_s3_score_utt_dev(b200_s3mgau_t *m, const float *d_feat, int T, int frame0,
                           uint8_t *d_sen_active, int32_t *d_out, int32_t *d_best,
                           void *stream);
/* Drop-in for one frame: s3_cd_gmm_compute_sen (gmm_wrap.c:103-171) with the
 * CI pass of approx_ci_gmm_compute (:174-214) folded in.  senscr is updated in
 * place; *best receives the return value of approx_cont_mgau_frame_eval. */
int  b200_s3_frame_eval(b200_s3mgau_t *m, const float *feat, int32_t frame,
                        uint8_t *sen_active, int32_t *senscr, int32_t *best);
/* Device time of the last *_dev / score call in ms (CUDA events). */
float b200_s3_last_ms(b200_s3mgau_t *m);
/* Measurement utility for the roofline of the float64 kernel: sustained
 * thread-level FP64 instructions per second (independent DMUL/DADD chains, the
 * non-fused mix the reference's arithmetic forces), measured on `device`. */
double b200_fp64_issue_rate(int device);

/* ===================================================== feature stage
 * Full-utterance cepstra -> `1s_c_d_dd` dynamic features with `-cmn current`
 * (or none), as acmod_process_cep(full_utt) computes them for batch decoding:
 * feat_s2mfc2feat_block_utt (SB/feat/feat.c:1241-1265: pad with 3 copies of the
 * first/last frame, then normalise), cmn (SB/feat/cmn.c:150-186),
 * feat_1s_c_d_dd_cep2feat (feat.c:726-769).  Bit-exact.  The 3-stream
 * `-svspec 0-12/13-25/26-38` models use the same layout.
 * cep [T_total][cepsize] holds n_utt utterances back to back, utterance u =
 * frames [utt_off[u], utt_off[u+1]); feat [T_total][3*cepsize].
 * d_mean_scratch: n_utt*cepsize floats (cmn == 1). */
int  b200_feat_1s_c_d_dd_dev(const float *d_cep, const int32_t *d_utt_off, int n_utt,
                             int T_total, int cepsize, int cmn, float *d_mean_scratch,
                             float *d_feat, void *stream);
int  b200_feat_1s_c_d_dd_host(const float *cep, const int32_t *utt_off, int n_utt,
                              int cepsize, int cmn, float *feat, int device);

/* General form of the same stage: every `-feat` type the reference computes
 * from cepstra, with its whole-utterance normalisations and the two linear
 * post-steps.  Replaces feat_s2mfc2feat_block_utt -> feat_compute_utt
 * (SB/feat/feat.c:1110-1135, 1241-1265) for a batch of utterances:
 *   cmn      0 = none, 1 = current           (feat_cmn feat.c:1060-1081; cmn.c:150-213)
 *   varnorm  unit variance per dimension      (cmn.c:186-212; needs cmn == 1)
 *   agc      0 = none, 1 = max on c0, applied after cmn (feat_agc feat.c:1084-1108; agc.c:108-126)
 *   type     the compute_feat function        (feat.c:559-849)
 *   lda      [lda_rows][lda_cols] float32, lda_cols == the type's stream length, rows used =
 *            lda_dim (0: all) -- feat_lda_transform (SB/feat/lda.c:141-160), single stream only
 *   subvec   n_subvec indices into the (LDA'd) vector -- feat_subvec_project (feat.c:334-355);
 *            the -svspec stream boundaries do not change the memory layout
 * `prior` CMN and `emax`/`noise` AGC are live-mode running estimates that the
 * reference does not use in whole-utterance mode (feat.c:1064-1066, 1088-1090)
 * and are refused.  Bit-exact: each float operation is the reference's, in its order. */
enum {
    B200_FEAT_1S_C_D_DD = 0,    /* c | d | dd, window 3                     feat.c:726-769 */
    B200_FEAT_S3_1X39 = 1,      /* c1-12 | d1-12 | c0 d0 dd0 | dd1-12, w 3  feat.c:622-672 */
    B200_FEAT_S2_4X = 2,        /* 4 streams 12|24|3|12, window 4           feat.c:559-618 */
    B200_FEAT_1S_C_D_LD_DD = 3, /* c | d | long d | dd, window 4            feat.c:772-825 */
    B200_FEAT_1S_C = 4,         /* cepstra only, window 0                   feat.c:676-683 */
    B200_FEAT_1S_C_D = 5,       /* c | d, window 2                          feat.c:700-723 */
    B200_FEAT_COPY = 6          /* the numeric types "n[,n..][:w]": frames t-w..t+w concatenated per stream,
                                 * feat_copy feat.c:828-849, 952-1000 (copy_window, copy_streams, copy_len).
                                 * (`1s_3c` / `1s_4c`, feat_s3_cepwin feat.c:687-696, copy contiguous memory
                                 * across the non-contiguous padded utterance: their edge frames are
                                 * undefined in the reference and the types are not offered.) */
};
typedef struct b200_feat_cfg {
    int32_t type, cepsize, cmn, varnorm, agc;
    int32_t lda_rows, lda_cols, lda_dim;   /* lda_rows == 0: no LDA */
    const float *lda;                      /* host pointer (both entry points) */
    int32_t n_subvec;                      /* 0: no projection */
    const int32_t *subvec;                 /* host pointer */
    /* B200_FEAT_COPY only: window w (0..7); copy_streams == 0 means one stream of cepsize
     * values, else the cepstral vector is cut into copy_streams pieces of copy_len[] values */
    int32_t copy_window, copy_streams;
    int32_t copy_len[B200_MAX_STREAMS];
} b200_feat_cfg_t;
/* dims = {window, stream-concatenated length before LDA, output length per frame};
 * B200_ERR_ARG / B200_ERR_UNSUP for a configuration the reference would reject. */
int  b200_feat_dims(const b200_feat_cfg_t *cfg, int32_t dims[3]);
/* d_cep [T_total][cepsize] -> d_feat [T_total][dims[2]], both on the device.
 * d_scratch: b200_feat_scratch_bytes(cfg, n_utt, T_total) bytes. */
size_t b200_feat_scratch_bytes(const b200_feat_cfg_t *cfg, int n_utt, int T_total);
int  b200_feat_compute_dev(const b200_feat_cfg_t *cfg, const float *d_cep, const int32_t *d_utt_off,
                           int n_utt, int T_total, void *d_scratch, float *d_feat, void *stream);
int  b200_feat_compute_host(const b200_feat_cfg_t *cfg, const float *cep, const int32_t *utt_off,
                            int n_utt, float *feat, int device);

/* -------------------------------------------------- device memory helpers
 * (so a non-torch host can keep buffers resident) */
void *b200_dev_alloc(size_t bytes, int device);
void  b200_dev_free(void *p);
int   b200_dev_upload(void *dst, const void *src, size_t bytes);
int   b200_dev_download(void *dst, const void *src, size_t bytes);
void *b200_host_alloc_pinned(size_t bytes);
void  b200_host_free_pinned(void *p);
int   b200_dev_sync(int device);

/* ------------------------------------------------------------------------
 * sphinx3's flavour of hmm_vit_eval (sphinx3/src/libs3decoder/libam/hmm.c:285-873,
 * dispatcher :852-873): int32 transition log-probabilities and int32 senone scores
 * that are ADDED, WORST_SCORE = S3_LOGPROB_ZERO = 0xc8000000 (s3types.h:192),
 * int32 senone-sequence ids (-1 = none), senone ids through sseq[ssid][state] for
 * mpx and non-mpx HMMs alike (hmm.h:218-226).  State-major SoA mirror of
 * sphinx3's hmm_t (hmm.h:184-197):
 *   score/history/ssid [n_emit][n_hmm] (non-mpx HMMs use ssid row 0 only),
 *   out_score/out_history/bestscore [n_hmm], tmatid [n_hmm], mpx [n_hmm].
 * tp: [n_tmat][n_emit][n_emit + 1] (what ctx->tp[tmatid][0] points at, row stride
 * n_emit + 1: hmm_tprob_3st(i, j) = tp[i * 4 + j]); sseq: s3senid_t
 * [n_sseq][n_emit]; senscr: [n_frames][n_sen].  Evaluates every HMM once per
 * frame (the population is updated in place), best_out[f] = best score of frame f. */
typedef struct {
    int32_t n_hmm;
    int32_t *score, *history, *ssid;
    int32_t *out_score, *out_history, *bestscore;
    const int32_t *tmatid;
    const uint8_t *mpx;
} b200_s3hmm_soa_t;
int  b200_s3hmm_eval_host(int n_emit, const int32_t *tp, int n_tmat, const int16_t *sseq, int n_sseq, int n_sen,
                          b200_s3hmm_soa_t *h, const int32_t *senscr, int n_frames, int32_t *best_out, int device);

#ifdef __cplusplus
}
#endif
#endif /* B200SPHINX_H */
