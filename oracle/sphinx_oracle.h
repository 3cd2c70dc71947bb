/* oracle/sphinx_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement ("port") of the pocketsphinx / sphinx3 hot path used as the
 * parity checker for the CUDA engine.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library;
 * the product (cmusphinx_b200/) never does.
 *
 * Pinned against the reference itself (oracle/_ref, built from
 * /root/reference by oracle/Makefile) in tests/test_oracle_vs_ref.py and
 * against committed golden vectors in tests/golden/.
 */
#ifndef SPHINX_ORACLE_H
#define SPHINX_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_SENSCR_SHIFT 10
#define ORC_WORST_SCORE ((int32_t)0xE0000000)
#define ORC_TMAT_WORST (-255)
#define ORC_BAD_SSID 0xffff
#define ORC_WORST_DIST ((int32_t)0x80000000)

/* ---- logmath (sphinxbase/src/libsphinxbase/util/logmath.c) ---- */
typedef struct {
    double base, inv_log_of_base;
    int shift;
    int32_t zero;
    uint32_t table_size;
    uint32_t *table; /* widened to 32 bit whatever the reference width */
} orc_logmath_t;

orc_logmath_t *orc_logmath_init(double base, int shift, int use_table);
void orc_logmath_free(orc_logmath_t *lm);
int orc_logmath_table(const orc_logmath_t *lm, int32_t *out, int max_out);
int32_t orc_logmath_log(const orc_logmath_t *lm, double p);
int32_t orc_logmath_ln_to_log(const orc_logmath_t *lm, double ln_p);
int32_t orc_logmath_add(const orc_logmath_t *lm, int32_t x, int32_t y);

/* ---- load-time precompute ---- */
/* ms_gauden.c:314-359. var: in = raw variances, out = scaled 1/(2 var);
 * det: out.  n_vec = number of (mgau,feat,density) vectors of length len. */
int orc_gauden_precompute(float *var, float *det, long n_vec, int len,
                          float varfloor, double logbase);
/* ms_senone.c:236-258. in: [n_sen][n_feat][n_cw] float32 (rows normalised in
 * place), out: same logical order uint8. */
int orc_mixw_quantize(float *mixw, uint8_t *out, int n_sen, int n_feat,
                      int n_cw, float mixwfloor, double logbase);
/* tmat.c:275-296.  in: [n_tmat][n_src][n_src+1] float32, out: uint8. */
int orc_tmat_quantize(float *tp, uint8_t *out, int n_tmat, int n_src,
                      double tpfloor, double logbase);

/* ---- multi-stream GMM scoring (ms_mgau.c / ms_gauden.c / ms_senone.c) ---- */
typedef struct {
    int n_mgau, n_feat, n_density, n_sen, topn, aw;
    int featlen[8];
    int featoff[8];          /* offset of stream f inside a frame vector */
    int veclen;              /* sum of featlen */
    const float *mean;       /* [mgau][feat][density][featlen[f]] */
    const float *var;        /* precomputed, same layout */
    const float *det;        /* [mgau][feat][density] */
    const uint8_t *mixw;     /* [sen][feat][cw] */
    const uint32_t *sen2mgau;/* [sen] */
    orc_logmath_t *lmath10;  /* base, shift 10, table */
} orc_ms_model_t;

orc_ms_model_t *orc_ms_model_new(int n_mgau, int n_feat, const int *featlen,
                                 int n_density, int n_sen, int topn, int aw,
                                 const float *mean, const float *var,
                                 const float *det, const uint8_t *mixw,
                                 const uint32_t *sen2mgau, double logbase);
void orc_ms_model_free(orc_ms_model_t *m);

/* Top-N of one codebook/stream, ms_gauden.c:417-523.  ids/dists: [topn]. */
void orc_ms_compute_dist(const orc_ms_model_t *m, int mgau, int feat,
                         const float *obs, int32_t *ids, float *dists);
/* One frame, ms_mgau.c:162-252.  feat: [veclen]; senscr: [n_sen]. */
int orc_ms_frame_eval(const orc_ms_model_t *m, const float *feat,
                      const uint8_t *senone_active, int n_senone_active,
                      int compallsen, int16_t *senscr);
/* T frames with compallsen=1.  feat [T][veclen], out [T][n_sen]. */
int orc_ms_eval_all(const orc_ms_model_t *m, const float *feat, int T,
                    int16_t *out);

/* ---- tied-mixture scoring: ptm_mgau.c (kind 1) / s2_semi_mgau.c (kind 2) ----
 * Sequential, stateful restatement: the top-N lists of the previous frame
 * seed the current one exactly as the reference's rotating history does
 * (ptm_mgau.c:405-450, s2_semi_mgau.c:840-886; -pl_window 0 => 2 slots). */
typedef struct orc_tied_model orc_tied_model_t;
orc_tied_model_t *orc_tied_new(int kind, int n_mgau, int n_feat, const int *featlen,
                               int n_density, int n_sen, int topn,
                               const float *mean, const float *var, const float *det,
                               const uint8_t *mixw /* [feat][density][row_bytes] */,
                               int row_bytes, int n_clust, const uint8_t *mixw_cb,
                               const uint8_t *sen2cb, double logbase);
void orc_tied_free(orc_tied_model_t *m);
void orc_tied_reset(orc_tied_model_t *m);   /* fresh history, as after *_init */
/* One frame_eval call; frames must be fed in increasing order. */
void orc_tied_set_ds(orc_tied_model_t *m, int ds_ratio);                 /* -ds, s2_semi only */
void orc_tied_set_topn_beam(orc_tied_model_t *m, const int *beam);   /* -topn_beam, s2_semi only */
int orc_tied_frame_eval(orc_tied_model_t *m, const float *feat,
                        const uint8_t *senone_active, int n_senone_active,
                        int compallsen, int frame, int16_t *senscr);
int orc_tied_eval_all(orc_tied_model_t *m, const float *feat, int T, int16_t *out);
/* Copies the current frame's lists: cw/score [n_mgau][n_feat][topn]. */
void orc_tied_lists(const orc_tied_model_t *m, int32_t *cw, int32_t *score);

/* tied mixw quantiser (ptm_mgau.c:720-742): in [sen][feat][cw] float (modified),
 * out [feat][cw][sen] uint8. */
int orc_mixw_quantize_tied(float *mixw, uint8_t *out, int n_sen, int n_feat,
                           int n_cw, float mixwfloor, double logbase);

/* acmod_flags2list (acmod.c:1219-1271) over a 32-bit-word bitmask. */
int orc_flags2list(const uint32_t *mask, int n_sen, uint8_t *deltas);

/* ---- HMM Viterbi step (hmm.c:224-807), SoA batch ---- */
/* Arrays are HMM-major ([hmm][state]) like the reference shim's. */
int32_t orc_hmm_eval_batch(int n_emit, int n_hmm, const uint8_t *tp, int n_tmat,
                           const uint16_t *sseq, int n_sseq,
                           const int16_t *senscr, int32_t *score,
                           int32_t *history, int32_t *out_score,
                           int32_t *out_history, uint16_t *senid,
                           const uint16_t *ssid, const int16_t *tmatid,
                           const uint8_t *mpx, int32_t *bestscore,
                           int n_frames_repeat);

/* ---- sphinx3 flavour: approx_cont_mgau_frame_eval (S3/libam/approx_cont_mgau.c,
 * cont_mgau.c, fast_algo_struct.c); int32 scores, float64 accumulation ---- */
typedef struct orc_s3_model orc_s3_model_t;
/* mgau_init on arrays: mean/var [n_sen][n_comp][veclen] RAW (variances not yet
 * inverted), mixw [n_sen][n_comp] raw counts. */
orc_s3_model_t *orc_s3_new(int n_sen, int n_comp, int veclen, const float *mean,
                           const float *var, const float *mixw, double varfloor,
                           double mixwfloor, double logbase,
                           const int32_t *cd2cisen, int n_ci_sen);
void orc_s3_free(orc_s3_model_t *m);
void orc_s3_set_fast(orc_s3_model_t *m, double ci_pbeam, int max_cd, int ds_ratio,
                     float tighten_factor);
int32_t orc_s3_ci_pbeam(const orc_s3_model_t *m);
void orc_s3_utt_reset(orc_s3_model_t *m);
void orc_s3_params(const orc_s3_model_t *m, int32_t *n_comp, float *mean, float *var,
                   float *lrd, int32_t *mixw, double *scal /* distfloor, f */);
void orc_s3_state(const orc_s3_model_t *m, int32_t *bstidx, int32_t *updatetime);
int32_t orc_s3_mgau_eval(orc_s3_model_t *m, int s, const int32_t *active,
                         const float *x, int fr, int update_best_id);
void orc_s3_ci_eval(orc_s3_model_t *m, const float *x, int32_t *ci_senscr,
                    int32_t *best, int fr);
int32_t orc_s3_frame_eval(orc_s3_model_t *m, const float *x, int frame,
                          const int32_t *cache_ci_senscr, uint8_t *sen_active,
                          int32_t *senscr);
void orc_s3_eval_utt(orc_s3_model_t *m, const float *feat, int T, int frame0,
                     uint8_t *sen_active, int32_t *senscr_io, int32_t *out,
                     int32_t *best);
void orc_s3_counts(const orc_s3_model_t *m, int64_t *c);

/* ---- feature stage: full-utterance cepstra -> 1s_c_d_dd features, -cmn current
 * (feat.c:726-769, 1241-1265; cmn.c:150-186).  cep [T][cepsize], feat [T][3*cepsize]. */
void orc_feat_1s_c_d_dd(const float *cep, int T, int cepsize, int cmn, float *feat);

/* General form (feat.c:1110-1135 feat_compute_utt on the padded utterance of
 * feat.c:1241-1265): type 0 1s_c_d_dd, 1 s3_1x39, 2 s2_4x, 3 1s_c_d_ld_dd, 4 1s_c,
 * 5 1s_c_d (feat.c:559-849); cmn 0/1 (+ varnorm, cmn.c:150-213); agc 0 none, 1 max
 * (agc.c:108-126); lda [lda_dim][k] or NULL (lda.c:141-160); subvec indices or NULL
 * (feat.c:334-355).  out [T][out_len]; returns out_len (<0: bad configuration). */
void orc_feat_set_copy(int window, int n_streams, const int *len);   /* parameters of type 6 (feat_copy types) */
int orc_feat_compute(int type, int cepsize, int cmn, int varnorm, int agc,
                     const float *lda, int lda_dim, const int *subvec, int n_subvec,
                     const float *cep, int T, float *out);

/* ---- prune / phone-transition stage of the forward tree search: prune_root_chan then
 * prune_nonroot_chan (ngram_search_fwdtree.c:714-790, 792-869), sequential, on the lexical tree as
 * arrays.  Channels 0..n_root-1 are the roots; children of channel c (its `next` channel and that
 * channel's `alt` chain, in chain order) are child[child_off[c] .. child_off[c+1]); the words whose
 * last phone follows c (penult_phn_wid and its homophone_set chain, :765, :847) are
 * pw_wid[pw_off[c]..], with dict_last_phone in pw_lastphone.  State arrays are state-major
 * ([ne][n_chan]); frame[] is hmm_frame.  par = {frame_idx, best_score, dynamic_beam, pbeam, lpbeam,
 * pip, nwpen, has_pls}; pls_pen[ciphone] = phone_loop_search_score (phone_loop_search.h:104).
 * acl = the current frame's active non-root channels in list order.  Out: nacl (next frame's list,
 * in the order the reference appends), cand = lastphn_cand {wid, score, bp} in append order. */
void orc_fwdtree_prune(int n_root, int n_chan, int ne, const int32_t *child_off, const int32_t *child,
                       const int32_t *ciphone, const int32_t *pw_off, const int32_t *pw_wid,
                       const int32_t *pw_lastphone, const int32_t *par, const int32_t *pls_pen,
                       const int32_t *acl, int n_act, int32_t *score, int32_t *history, int32_t *out_score,
                       int32_t *out_history, int32_t *bestscore, int32_t *frame, int32_t *nacl, int32_t *n_nacl,
                       int32_t *cand, int32_t *n_cand);

/* ---- the phone-loop look-ahead search, one frame: phone_loop_search_step (phone_loop_search.c:253-291)
 * minus its acmod calls = renormalize_hmms (:171-184) if best + 2 beam underflows, evaluate_hmms
 * (:186-210), prune_hmms (:212-233), phone_transition (:235-268).  HMM-major arrays ([n_phones][ne]),
 * non-mpx HMMs, senid = senone ids into senscr.  par = {frame_idx, best_score, beam, pbeam, pip}:
 * best_score is the previous frame's on entry and this frame's on return (also the return value);
 * *renorm = 1 when the frame renormalised (by the previous best). */
int32_t orc_phone_loop_step(int n_phones, int ne, const uint8_t *tp, const int16_t *senscr, int32_t *par,
                            int32_t *score, int32_t *history, int32_t *out_score, int32_t *out_history,
                            int32_t *bestscore, int32_t *frame, uint16_t *senid, const int16_t *tmatid,
                            int32_t *renorm);

#ifdef __cplusplus
}
#endif
/* sphinx3's hmm_vit_eval flavour (libs3decoder/libam/hmm.c:285-873), HMM-major arrays */
int32_t orc_s3hmm_eval_batch(int ne, int n_hmm, const int32_t *tp, int n_tmat, const int16_t *sseq, int n_sseq,
                             const int32_t *sen, int32_t *score, int32_t *history, int32_t *out_score,
                             int32_t *out_history, int32_t *ssid, const int32_t *tmatid, const uint8_t *mpx,
                             int32_t *bestscore, int repeat);

#endif
