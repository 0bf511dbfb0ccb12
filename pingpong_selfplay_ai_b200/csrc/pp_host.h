// pp_host.h — launchers behind the C ABI (one per kernel family); all return a cudaError_t as int.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/pong_b200.h"

namespace pp {

int env_step_launch(int mode, int64_t n, const PPParams &p, const PPEnvState &st, const uint8_t *aa, const uint8_t *ab,
                    float *oa, float *ob, float *ra, float *rb, uint8_t *dn, cudaStream_t stream);
int env_observe_launch(int mode, int64_t n, const PPEnvState &st, float *oa, float *ob, cudaStream_t stream);
int env_serve_launch(int mode, int64_t n, const PPEnvState &st, const uint8_t *mask, const void *vx, const void *vy,
                     const void *spin, cudaStream_t stream);
int env_reset_launch(int mode, int64_t n, const PPParams &p, const PPEnvState &st, const uint8_t *mask,
                     const PPServeSource &src, int64_t env_id_base, int advance, cudaStream_t stream);
int collide_launch(int mode, int64_t n, const PPParams &p, const void *vn, const void *vt, const void *u, const void *om,
                   void *vn_out, void *vt_out, void *om_out, cudaStream_t stream);
int env_rollout_launch(int mode, int64_t n, int64_t k, const PPParams &p, const PPEnvState &st, const uint8_t *actions,
                       const PPServeSource &src, int32_t quota, int64_t env_id_base, const PPRolloutOut &out,
                       cudaStream_t stream);

int qnet_act_launch(int64_t n, const float *obs, const PPPolicy &pol, uint64_t seed, int64_t step_index,
                    int64_t env_id_base, int32_t stream_id, uint8_t *actions, float *q_out, cudaStream_t stream);
int qnetrnn_act_launch(int64_t n, const float *obs, const PPPolicy &pol, const uint8_t *reset_mask, uint64_t seed,
                       int64_t step_index, int64_t env_id_base, int32_t stream_id, uint8_t *actions, float *q_out,
                       cudaStream_t stream);
int selfplay_launch(int mode, int64_t n, int64_t k, const PPParams &p, const PPEnvState &st, const PPPolicy &pa,
                    const PPPolicy &pb, uint64_t seed, int64_t step_base, const PPServeSource &src, int32_t quota,
                    int64_t env_id_base, const PPRolloutOut &out, const PPReplayRing *ring, cudaStream_t stream);
int qnet_act_tc_launch(int64_t n, const float *obs, const PPPolicy &pol, uint64_t seed, int64_t step_index,
                       int64_t env_id_base, int32_t stream_id, uint8_t *actions, float *q_out, cudaStream_t stream);
int selfplay_tc_launch(int mode, int64_t n, int64_t k, const PPParams &p, const PPEnvState &st, const PPPolicy &pa,
                       const PPPolicy &pb, uint64_t seed, int64_t step_base, const PPServeSource &src, int32_t quota,
                       int64_t env_id_base, const PPRolloutOut &out, const PPReplayRing *ring, cudaStream_t stream);
int selfplay_rnn_launch(int mode, int64_t n, int64_t k, const PPParams &p, const PPEnvState &st, const PPPolicy &pa,
                        const PPPolicy &pb, uint64_t seed, int64_t step_base, const PPServeSource &src, int32_t quota,
                        int64_t env_id_base, const PPRolloutOut &out, const PPReplayRing *ring, cudaStream_t stream);
int qnetrnn_act_tc_launch(int64_t n, const float *obs, const PPPolicy &pol, const uint8_t *reset_mask, uint64_t seed,
                          int64_t step_index, int64_t env_id_base, int32_t stream_id, uint8_t *actions, float *q_out,
                          cudaStream_t stream);
int selfplay_rnn_tc_launch(int mode, int64_t n, int64_t k, const PPParams &p, const PPEnvState &st, const PPPolicy &pa,
                           const PPPolicy &pb, uint64_t seed, int64_t step_base, const PPServeSource &src, int32_t quota,
                           int64_t env_id_base, const PPRolloutOut &out, const PPReplayRing *ring, cudaStream_t stream);
int replay_scatter_launch(int64_t n, const PPReplayRing &ring, const float *obs, const uint8_t *act, const float *rew,
                          const float *next_obs, const uint8_t *done, const uint8_t *valid, cudaStream_t stream);

int dqn_head_grads_launch(const PPReplayRing &ring, const int64_t *idx, const float *iw, int32_t batch,
                          const float *w1, const float *b1, const float *w2, const float *b2, const PPNoisyLayer &on_v, const PPNoisyLayer &on_a,
                          const PPNoisyLayer &tg_v, const PPNoisyLayer &tg_a, int noisy_online, int noisy_target, float gamma,
                          float *td_out, float *loss_out, float *prios, float *max_prio, float *workspace, cudaStream_t stream);
int64_t dqn_workspace_floats(int32_t batch);
int64_t per_chunk(int64_t capacity);
int per_sample_launch(const float *prios, int64_t capacity, float alpha, const float *beta, const float *size, uint64_t seed,
                      unsigned long long *counter, int32_t batch, float *chunk_sums, int64_t *idx_out, float *w_out,
                      cudaStream_t stream);
int adam_step_launch(const PPAdamParam *params, int32_t count, double lr, double beta1, double beta2, double eps, cudaStream_t stream);
int adam_allreduce_launch(const PPAdamParam *params, int32_t count, float *flat_grad, int64_t numel, const PPPeerBlocks &peers,
                          unsigned long long *epoch, double lr, double beta1, double beta2, double eps, cudaStream_t stream);
int noisy_reset_launch(const PPNoisyLayer *layers, int32_t count, uint64_t seed, unsigned long long *counter, cudaStream_t stream);
int pack_qnet_launch(const float *w1, const float *b1, const float *w2, const float *b2, const PPNoisyLayer &v,
                     const PPNoisyLayer &a, int noisy, float *blob, cudaStream_t stream);

int drqn_grads_launch(const PPReplayRing &ring, const int64_t *rows, int32_t batch, int32_t trace, const PPQNetRNNParams &on,
                      const PPQNetRNNParams &tg, int noisy_on, int noisy_tg, float gamma, const PPQNetRNNGrads &gr,
                      float *loss_out, float *td_out, float *ws, cudaStream_t stream);
int64_t drqn_workspace_floats(int32_t batch, int32_t trace);
int clip_grad_norm_launch(float *flat, int64_t numel, float max_norm, float *norm_out, float *scratch, cudaStream_t stream);
int seq_window_weights_launch(const uint8_t *done, int64_t n, int64_t T, int64_t steps_written, int trace, int starts_fresh,
                              float *weights, unsigned long long *episodes, cudaStream_t stream);
int seq_expand_rows_launch(const int64_t *end_slots, int batch, int trace, int64_t n, int64_t T, int64_t *rows, cudaStream_t stream);
int pack_qnetrnn_tc_launch(const PPQNetRNNParams &p, int noisy, void *image, cudaStream_t stream);
int adam_multi_launch(const PPAdamParam *params, int32_t count, double lr, double beta1, double beta2, double eps, cudaStream_t stream);

}  // namespace pp
