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

struct HmmFrame { int32_t best; int32_t n_keep; };

int hmm_launch_step(const HmmDev &c, const HmmPop &p, const int16_t *d_senscr, int32_t beam,
                    HmmFrame *fr, uint8_t *keep, int32_t *block_count, int32_t *keep_idx,
                    uint32_t *mask, int32_t *total, int do_beam, cudaStream_t st);

int hmm_launch_normalize(const HmmPop &p, int n_emit, const int32_t *d_best_per_utt, const HmmFrame *fr, cudaStream_t st);
int hmm_launch_clear_pruned(const HmmPop &p, int n_emit, const uint8_t *keep, cudaStream_t st);
int hmm_launch_enter(const HmmPop &p, const int32_t *d_idx, const int32_t *d_score, const int32_t *d_hist, int n,
                     int32_t *d_winner, int32_t *d_old0, uint8_t *d_entered, cudaStream_t st);

}  // namespace b200
