// hmm_dev.cuh -- device-side views for the HMM kernels.
#pragma once
#include "dev_common.cuh"

namespace b200 {

struct HmmDev {               // hmm_context_t (PS/hmm.h:136-151)
    int n_emit, n_tmat, n_sseq, n_sen;
    const uint8_t *tp;        // [n_tmat][n_emit][n_emit+1]
    const uint16_t *sseq;     // [n_sseq][n_emit]
};

struct HmmPop {               // SoA mirror of hmm_t (PS/hmm.h:156-173), state-major
    int n_hmm;
    int n_utt, max_per_utt;   // utterances sharing the population; largest range
    const int32_t *utt_off;   // [n_utt + 1] device
    int32_t *score, *history, *out_score, *out_history, *bestscore;
    uint16_t *senid;
    int16_t *tmatid;
    uint8_t *mpx;
};

// per utterance and frame: best score, survivors, the beam threshold they were tested against
struct HmmFrame { int32_t best; int32_t n_keep; int32_t thresh; int32_t pad; };

// one launch = a run of frames (see hmm_run_kernel)
constexpr int kHmmBarRows = 4096;
struct HmmRun {
    const int16_t *sen_base; long frame_stride; int n_cycle, frame0, n_frames;   // frame f scores: sen_base + ((frame0 + f) % n_cycle) * frame_stride + utt * n_sen
    int32_t beam; int do_beam;
    HmmFrame *fr3; int slot0;            // [3][n_utt] frame records, frame f uses slot (slot0 + f) % 3
    int32_t *tile_count;                 // [n_utt][tpu] survivors per 256-HMM tile
    int tpu;                             // tiles per utterance (largest)
    int32_t *keep_idx;
    int32_t *keep_tmp;                   // [n_hmm] per-utterance survivor lists of the cluster form (packed into keep_idx after the run)
    uint32_t *mask2; int mask0;          // [2][n_utt][n_words], frame f uses (mask0 + f) & 1
    uint32_t *mask_part; size_t mask_part_words;   // [2][n_utt][gx][n_words] per-CTA partial masks
    int32_t *total;
    unsigned *bar;                       // barrier counters: [0] for the grid, [row * 32] per grid row (kHmmBarRows rows)
    int pre_off;                         // set by the launcher: byte offset of phase A's prefetch slots in dynamic shared memory
    int pair;                            // set by the launcher: phase A may take two adjacent HMMs per thread (see hmm_run_kernel)
    int row_sync;                        // set by the launcher: barriers per grid row (see hmm_run_kernel)
    long long *probe;                    // development: phase time stamps of CTA 0 on the last frame (or null)
};
int hmm_launch_run(const HmmDev &c, const HmmPop &p, const HmmRun &run, cudaStream_t st);

// eval_root_chan + eval_nonroot_chan over per-utterance channel lists (see hmm_eval_list_kernel)
struct HmmList {
    int n_root, n_chan, list_cap;
    const int32_t *frame;      // [n_utt * n_chan] hmm_frame
    const int32_t *par;        // [n_utt][8], par[0] = frame_idx (the b200_fwdtree_prune_* parameter rows)
    const int32_t *acl, *n_act;
    const int16_t *senscr;     // [n_utt][n_sen]
    int32_t *best;             // [n_utt] out
};
int hmm_launch_eval_list(const HmmDev &c, const HmmPop &p, const HmmList &l, int n_utt, cudaStream_t st);

int hmm_launch_normalize(const HmmPop &p, int n_emit, const int32_t *d_best_per_utt, const HmmFrame *fr, cudaStream_t st);
int hmm_launch_clear_pruned(const HmmPop &p, int n_emit, const HmmFrame *fr, cudaStream_t st);
int hmm_launch_enter(const HmmPop &p, const int32_t *d_idx, const int32_t *d_score, const int32_t *d_hist, int n,
                     int32_t *d_winner, int32_t *d_old0, uint8_t *d_entered, cudaStream_t st);

}  // namespace b200
