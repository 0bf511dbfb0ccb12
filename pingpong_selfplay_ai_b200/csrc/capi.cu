// capi.cu — the extern "C" surface declared in include/pong_b200.h: argument validation, error text,
// and the HOST-buffer evaluation entry (pp_host_selfplay_eval).  No kernel code lives here.
#include <cstdio>
#include <cstring>
#include <mutex>
#include <vector>

#include "pp_host.h"

namespace {

thread_local char g_err[256] = "";

int fail(int code, const char *what) {
    if (code > 0) snprintf(g_err, sizeof g_err, "%s: CUDA error %d (%s)", what, code, cudaGetErrorString((cudaError_t)code));
    else snprintf(g_err, sizeof g_err, "%s: bad argument (%d)", what, code);
    return code;
}
int ok_or(int cuda_code, const char *what) { return cuda_code == 0 ? 0 : fail(cuda_code, what); }

bool mode_ok(int mode) { return mode == PP_MODE_F64 || mode == PP_MODE_F32; }
bool state_ok(const PPEnvState *s, bool need_ep) {
    if (!s) return false;
    const void *p[] = {s->ball_x, s->ball_y, s->ball_vx, s->ball_vy, s->spin, s->top_paddle_x, s->bottom_paddle_x,
                       s->score_a, s->score_b, s->bounce_count};
    for (const void *q : p) if (!q) return false;
    return !need_ep || (s->ep_idx && s->ep_len);
}
bool params_ok(const PPParams *p) { return p && p->speed_scale_every > 0 && p->max_score > 0; }
bool serve_ok(const PPServeSource *s) {
    if (!s) return false;
    if (s->kind == PP_SERVE_POOL) return s->depth > 0 && s->pool_vx && s->pool_vy && s->pool_spin;
    if (s->kind == PP_SERVE_QUEUE) {                 // with a pool, or (all three NULL) Philox serves keyed by `seed`
        const int have = (s->pool_vx != nullptr) + (s->pool_vy != nullptr) + (s->pool_spin != nullptr);
        return s->queue_total > 0 && s->queue_total < 0x7fffffff && s->queue_head && (have == 0 || have == 3);
    }
    return s->kind == PP_SERVE_PHILOX;
}
bool policy_ok(const PPPolicy *p, bool allow_rnn) {
    if (!p) return false;
    switch (p->kind) {
        case PP_POLICY_QNET: return p->weights && (reinterpret_cast<uintptr_t>(p->weights) & 15u) == 0;
        case PP_POLICY_QNETRNN: return allow_rnn && p->weights && p->h && p->c && (reinterpret_cast<uintptr_t>(p->weights) & 15u) == 0;
        case PP_POLICY_FOLLOWER:
        case PP_POLICY_RANDOM: return true;
        default: return false;
    }
}
bool prec_ok(const PPPolicy *p) { return p->precision == PP_PREC_F32 || p->precision == PP_PREC_F16; }
bool uses_tc(const PPPolicy *p) { return p->kind == PP_POLICY_QNET && p->precision == PP_PREC_F16; }
bool out_ok(const PPRolloutOut *o) {
    if (!o) return false;
    if (o->ep_log && (!o->ep_log_count || o->ep_log_cap < 0 || (reinterpret_cast<uintptr_t>(o->ep_log) & 15u))) return false;
    if (o->actions_out && (reinterpret_cast<uintptr_t>(o->actions_out) & 1u)) return false;
    return true;
}
bool ring_ok(const PPReplayRing *r) {
    return r && r->obs && r->act && r->rew && r->next_obs && r->done && r->head && r->capacity > 0;
}

}  // namespace

extern "C" {

int pp_version(void) { return PP_ABI_VERSION; }
const char *pp_last_error(void) { return g_err; }

int pp_env_step(int mode, int64_t n, const PPParams *params, const PPEnvState *state, const uint8_t *action_a,
                const uint8_t *action_b, float *obs_a, float *obs_b, float *reward_a, float *reward_b, uint8_t *done,
                void *stream) {
    if (!mode_ok(mode)) return fail(PP_E_MODE, "pp_env_step");
    if (n < 0 || n > (int64_t)1 << 40) return fail(PP_E_SIZE, "pp_env_step");
    if (!params_ok(params)) return fail(PP_E_PARAM, "pp_env_step");
    if (!state_ok(state, false) || !action_a || !action_b || !obs_a || !obs_b || !reward_a || !reward_b || !done)
        return fail(PP_E_NULL, "pp_env_step");
    if (n == 0) return 0;
    return ok_or(pp::env_step_launch(mode, n, *params, *state, action_a, action_b, obs_a, obs_b, reward_a, reward_b, done,
                                     (cudaStream_t)stream), "pp_env_step");
}

int pp_env_observe(int mode, int64_t n, const PPEnvState *state, float *obs_a, float *obs_b, void *stream) {
    if (!mode_ok(mode)) return fail(PP_E_MODE, "pp_env_observe");
    if (n < 0) return fail(PP_E_SIZE, "pp_env_observe");
    if (!state_ok(state, false) || !obs_a || !obs_b) return fail(PP_E_NULL, "pp_env_observe");
    if (n == 0) return 0;
    return ok_or(pp::env_observe_launch(mode, n, *state, obs_a, obs_b, (cudaStream_t)stream), "pp_env_observe");
}

int pp_env_serve(int mode, int64_t n, const PPEnvState *state, const uint8_t *mask, const void *vx, const void *vy,
                 const void *spin, void *stream) {
    if (!mode_ok(mode)) return fail(PP_E_MODE, "pp_env_serve");
    if (n < 0) return fail(PP_E_SIZE, "pp_env_serve");
    if (!state_ok(state, false) || !vx || !vy || !spin) return fail(PP_E_NULL, "pp_env_serve");
    if (n == 0) return 0;
    return ok_or(pp::env_serve_launch(mode, n, *state, mask, vx, vy, spin, (cudaStream_t)stream), "pp_env_serve");
}

int pp_env_reset(int mode, int64_t n, const PPParams *params, const PPEnvState *state, const uint8_t *mask,
                 const PPServeSource *serve, int64_t env_id_base, int advance, void *stream) {
    if (!mode_ok(mode)) return fail(PP_E_MODE, "pp_env_reset");
    if (n < 0) return fail(PP_E_SIZE, "pp_env_reset");
    if (!params_ok(params)) return fail(PP_E_PARAM, "pp_env_reset");
    if (!state_ok(state, true) || !serve_ok(serve)) return fail(PP_E_NULL, "pp_env_reset");
    if (n == 0) return 0;
    return ok_or(pp::env_reset_launch(mode, n, *params, *state, mask, *serve, env_id_base, advance, (cudaStream_t)stream),
                 "pp_env_reset");
}

int pp_collide(int mode, int64_t n, const PPParams *params, const void *vn, const void *vt, const void *u, const void *omega,
               void *vn_out, void *vt_out, void *omega_out, void *stream) {
    if (!mode_ok(mode)) return fail(PP_E_MODE, "pp_collide");
    if (n < 0) return fail(PP_E_SIZE, "pp_collide");
    if (!params) return fail(PP_E_PARAM, "pp_collide");
    if (!vn || !vt || !u || !omega || !vn_out || !vt_out || !omega_out) return fail(PP_E_NULL, "pp_collide");
    if (n == 0) return 0;
    return ok_or(pp::collide_launch(mode, n, *params, vn, vt, u, omega, vn_out, vt_out, omega_out, (cudaStream_t)stream),
                 "pp_collide");
}

int pp_env_rollout(int mode, int64_t n, int64_t k, const PPParams *params, const PPEnvState *state, const uint8_t *actions,
                   const PPServeSource *serve, int32_t quota, int64_t env_id_base, const PPRolloutOut *out, void *stream) {
    if (!mode_ok(mode)) return fail(PP_E_MODE, "pp_env_rollout");
    if (n < 0 || k < 0) return fail(PP_E_SIZE, "pp_env_rollout");
    if (!params_ok(params)) return fail(PP_E_PARAM, "pp_env_rollout");
    if (!state_ok(state, true) || !actions || !serve_ok(serve) || !out_ok(out)) return fail(PP_E_NULL, "pp_env_rollout");
    if (serve->kind == PP_SERVE_QUEUE) return fail(PP_E_MODE, "pp_env_rollout");      // the queue is an evaluation mode
    if (reinterpret_cast<uintptr_t>(actions) & 1u) return fail(PP_E_ALIGN, "pp_env_rollout");
    if (n == 0 || k == 0) return 0;
    return ok_or(pp::env_rollout_launch(mode, n, k, *params, *state, actions, *serve, quota, env_id_base, *out,
                                        (cudaStream_t)stream), "pp_env_rollout");
}

int pp_qnet_act(int64_t n, const float *obs, const PPPolicy *policy, uint64_t seed, int64_t step_index,
                int64_t env_id_base, int32_t stream_id, uint8_t *actions, float *q_out, void *stream) {
    if (n < 0) return fail(PP_E_SIZE, "pp_qnet_act");
    if (!obs || !actions) return fail(PP_E_NULL, "pp_qnet_act");
    if (!policy_ok(policy, false)) return fail(PP_E_MODE, "pp_qnet_act");
    if (!prec_ok(policy)) return fail(PP_E_MODE, "pp_qnet_act");
    if (n == 0) return 0;
    if (uses_tc(policy))
        return ok_or(pp::qnet_act_tc_launch(n, obs, *policy, seed, step_index, env_id_base, stream_id, actions, q_out,
                                            (cudaStream_t)stream), "pp_qnet_act");
    return ok_or(pp::qnet_act_launch(n, obs, *policy, seed, step_index, env_id_base, stream_id, actions, q_out,
                                     (cudaStream_t)stream), "pp_qnet_act");
}

int pp_qnetrnn_act(int64_t n, const float *obs, const PPPolicy *policy, const uint8_t *reset_mask, uint64_t seed,
                   int64_t step_index, int64_t env_id_base, int32_t stream_id, uint8_t *actions, float *q_out, void *stream) {
    if (n < 0) return fail(PP_E_SIZE, "pp_qnetrnn_act");
    if (!obs || !actions) return fail(PP_E_NULL, "pp_qnetrnn_act");
    if (!policy_ok(policy, true) || policy->kind != PP_POLICY_QNETRNN || !prec_ok(policy))
        return fail(PP_E_MODE, "pp_qnetrnn_act");
    if (n == 0) return 0;
    if (policy->precision == PP_PREC_F16)        // weights = the fp16 stage image PP_RNNTC_*
        return ok_or(pp::qnetrnn_act_tc_launch(n, obs, *policy, reset_mask, seed, step_index, env_id_base, stream_id,
                                               actions, q_out, (cudaStream_t)stream), "pp_qnetrnn_act");
    return ok_or(pp::qnetrnn_act_launch(n, obs, *policy, reset_mask, seed, step_index, env_id_base, stream_id, actions,
                                        q_out, (cudaStream_t)stream), "pp_qnetrnn_act");
}

int pp_selfplay_rollout(int mode, int64_t n, int64_t k, const PPParams *params, const PPEnvState *state,
                        const PPPolicy *policy_a, const PPPolicy *policy_b, uint64_t seed, int64_t step_base,
                        const PPServeSource *serve, int32_t quota, int64_t env_id_base, const PPRolloutOut *out,
                        const PPReplayRing *ring, void *stream) {
    if (!mode_ok(mode)) return fail(PP_E_MODE, "pp_selfplay_rollout");
    if (n < 0 || k < 0 || k > 0x7fffffff) return fail(PP_E_SIZE, "pp_selfplay_rollout");
    if (!params_ok(params)) return fail(PP_E_PARAM, "pp_selfplay_rollout");
    if (!state_ok(state, true) || !serve_ok(serve) || !out_ok(out)) return fail(PP_E_NULL, "pp_selfplay_rollout");
    if (!policy_ok(policy_a, true) || !policy_ok(policy_b, true)) return fail(PP_E_MODE, "pp_selfplay_rollout");
    const bool rnn = policy_a->kind == PP_POLICY_QNETRNN || policy_b->kind == PP_POLICY_QNETRNN;
    // a recurrent player meets any other player (tests/arena.py pairings); a QNet that meets a recurrent player runs in
    // fp32 on the CUDA cores of the recurrent kernel whatever its `precision` says (exact fp32 meets both tolerances)
    if (!prec_ok(policy_a) || !prec_ok(policy_b)) return fail(PP_E_MODE, "pp_selfplay_rollout");
    const bool rnn_tc = rnn && ((policy_a->kind == PP_POLICY_QNETRNN && policy_a->precision == PP_PREC_F16) ||
                                (policy_b->kind == PP_POLICY_QNETRNN && policy_b->precision == PP_PREC_F16));
    if (rnn_tc && ((policy_a->kind == PP_POLICY_QNETRNN && policy_a->precision != PP_PREC_F16) ||
                   (policy_b->kind == PP_POLICY_QNETRNN && policy_b->precision != PP_PREC_F16)))
        return fail(PP_E_MODE, "pp_selfplay_rollout");                 // both recurrent players on the same path
    const bool tc = !rnn && (uses_tc(policy_a) || uses_tc(policy_b));
    if (tc && ((policy_a->kind == PP_POLICY_QNET && !uses_tc(policy_a)) || (policy_b->kind == PP_POLICY_QNET && !uses_tc(policy_b))))
        return fail(PP_E_MODE, "pp_selfplay_rollout");                 // both QNet players on the same path
    if (ring && !ring_ok(ring)) return fail(PP_E_NULL, "pp_selfplay_rollout");
    if (ring && ring->capacity < n) return fail(PP_E_SIZE, "pp_selfplay_rollout");     // one lock-step step must fit
    if (ring && ring->lockstep_envs != 0 &&
        (ring->lockstep_envs != n || ring->capacity % n != 0 || ring->lockstep_step0 < 0))
        return fail(PP_E_SIZE, "pp_selfplay_rollout");                 // [capacity / n][n] layout of exactly this slab
    if (serve->kind == PP_SERVE_QUEUE && (int64_t)quota != serve->queue_total) return fail(PP_E_SIZE, "pp_selfplay_rollout");
    if (n == 0 || k == 0) return 0;
    if (rnn_tc)
        return ok_or(pp::selfplay_rnn_tc_launch(mode, n, k, *params, *state, *policy_a, *policy_b, seed, step_base, *serve,
                                                quota, env_id_base, *out, ring, (cudaStream_t)stream), "pp_selfplay_rollout");
    if (rnn)
        return ok_or(pp::selfplay_rnn_launch(mode, n, k, *params, *state, *policy_a, *policy_b, seed, step_base, *serve,
                                             quota, env_id_base, *out, ring, (cudaStream_t)stream), "pp_selfplay_rollout");
    if (tc)
        return ok_or(pp::selfplay_tc_launch(mode, n, k, *params, *state, *policy_a, *policy_b, seed, step_base, *serve,
                                            quota, env_id_base, *out, ring, (cudaStream_t)stream), "pp_selfplay_rollout");
    return ok_or(pp::selfplay_launch(mode, n, k, *params, *state, *policy_a, *policy_b, seed, step_base, *serve, quota,
                                     env_id_base, *out, ring, (cudaStream_t)stream), "pp_selfplay_rollout");
}

int pp_replay_scatter(int64_t n, const PPReplayRing *ring, const float *obs, const uint8_t *act, const float *rew,
                      const float *next_obs, const uint8_t *done, const uint8_t *valid, void *stream) {
    if (n < 0) return fail(PP_E_SIZE, "pp_replay_scatter");
    if (!ring_ok(ring) || !obs || !act || !rew || !next_obs || !done) return fail(PP_E_NULL, "pp_replay_scatter");
    if (ring->lockstep_envs != 0) return fail(PP_E_MODE, "pp_replay_scatter");       // appends have no (step, env) address
    if (n == 0) return 0;
    return ok_or(pp::replay_scatter_launch(n, *ring, obs, act, rew, next_obs, done, valid, (cudaStream_t)stream),
                 "pp_replay_scatter");
}

// ---------------------------------------------------------------------------------------------
// Training mode.
namespace {
bool noisy_ok(const PPNoisyLayer *l, int in, int out) {
    return l && l->in_features == in && l->out_features == out && l->weight_mu && l->weight_sigma && l->weight_epsilon &&
           l->bias_mu && l->bias_sigma && l->bias_epsilon;
}
}  // namespace

int pp_noisy_reset(const PPNoisyLayer *layers, int32_t count, uint64_t seed, unsigned long long *counter, void *stream) {
    if (count < 0 || count > 8) return fail(PP_E_SIZE, "pp_noisy_reset");
    if (!counter || (count > 0 && !layers)) return fail(PP_E_NULL, "pp_noisy_reset");
    for (int i = 0; i < count; ++i) {
        const PPNoisyLayer &l = layers[i];
        if (l.in_features <= 0 || l.out_features <= 0 || l.in_features + l.out_features > 1024) return fail(PP_E_SIZE, "pp_noisy_reset");
        if (!l.weight_epsilon || !l.bias_epsilon) return fail(PP_E_NULL, "pp_noisy_reset");
    }
    if (count == 0) return 0;
    return ok_or(pp::noisy_reset_launch(layers, count, seed, counter, (cudaStream_t)stream), "pp_noisy_reset");
}

int pp_pack_qnet(const float *w1, const float *b1, const float *w2, const float *b2, const PPNoisyLayer *fc_v,
                 const PPNoisyLayer *fc_a, int32_t noisy, float *blob, void *stream) {
    if (!w1 || !b1 || !w2 || !b2 || !blob) return fail(PP_E_NULL, "pp_pack_qnet");
    if (!noisy_ok(fc_v, 64, 1) || !noisy_ok(fc_a, 64, 3)) return fail(PP_E_PARAM, "pp_pack_qnet");
    return ok_or(pp::pack_qnet_launch(w1, b1, w2, b2, *fc_v, *fc_a, noisy, blob, (cudaStream_t)stream), "pp_pack_qnet");
}

int pp_dqn_head_grads(const PPReplayRing *ring, const int64_t *idx, const float *iw, int32_t batch,
                      const float *features0_weight, const float *features0_bias, const float *features2_weight,
                      const float *features2_bias, const PPNoisyLayer *online_v, const PPNoisyLayer *online_a,
                      const PPNoisyLayer *target_v, const PPNoisyLayer *target_a, int32_t noisy_online,
                      int32_t noisy_target, float gamma, float *td_out, float *loss_out, float *prios, float *max_prio,
                      float *workspace, void *stream) {
    const char *fn = "pp_dqn_head_grads";
    if (batch <= 0 || batch > 4096) return fail(PP_E_SIZE, fn);
    if (!workspace) return fail(PP_E_NULL, fn);
    if (!ring_ok(ring) || !idx || !iw || !features0_weight || !features0_bias || !features2_weight || !features2_bias)
        return fail(PP_E_NULL, fn);
    if (!noisy_ok(online_v, 64, 1) || !noisy_ok(online_a, 64, 3) || !noisy_ok(target_v, 64, 1) || !noisy_ok(target_a, 64, 3))
        return fail(PP_E_PARAM, fn);
    return ok_or(pp::dqn_head_grads_launch(*ring, idx, iw, batch, features0_weight, features0_bias, features2_weight,
                                           features2_bias, *online_v, *online_a, *target_v, *target_a,
                                           noisy_online, noisy_target, gamma, td_out, loss_out, prios, max_prio, workspace, (cudaStream_t)stream), fn);
}

int64_t pp_dqn_workspace_floats(int32_t batch) { return batch > 0 ? pp::dqn_workspace_floats(batch) : 0; }

int pp_per_sample(const float *prios, int64_t capacity, float alpha, const float *beta, const float *size, uint64_t seed,
                  unsigned long long *counter, int32_t batch, float *chunk_sums, int64_t *idx_out, float *weights_out,
                  void *stream) {
    if (capacity <= 0 || batch <= 0 || batch > 4096) return fail(PP_E_SIZE, "pp_per_sample");
    if (!prios || !beta || !size || !counter || !chunk_sums || !idx_out || !weights_out) return fail(PP_E_NULL, "pp_per_sample");
    if (!(alpha >= 0.f)) return fail(PP_E_PARAM, "pp_per_sample");
    return ok_or(pp::per_sample_launch(prios, capacity, alpha, beta, size, seed, counter, batch, chunk_sums, idx_out,
                                       weights_out, (cudaStream_t)stream), "pp_per_sample");
}

int64_t pp_per_sample_scratch_floats(int64_t capacity) {
    if (capacity <= 0) return 0;
    const int64_t chunk = pp::per_chunk(capacity);
    return (capacity + chunk - 1) / chunk;
}

int pp_adam_step(const PPAdamParam *params, int32_t count, double lr, double beta1, double beta2, double eps, void *stream) {
    if (count < 0 || count > 16) return fail(PP_E_SIZE, "pp_adam_step");
    if (count > 0 && !params) return fail(PP_E_NULL, "pp_adam_step");
    for (int i = 0; i < count; ++i) {
        const PPAdamParam &a = params[i];
        if (!a.param || !a.grad || !a.exp_avg || !a.exp_avg_sq || !a.step) return fail(PP_E_NULL, "pp_adam_step");
        if (a.numel < 0) return fail(PP_E_SIZE, "pp_adam_step");
    }
    if (!(beta1 >= 0.0 && beta1 < 1.0 && beta2 >= 0.0 && beta2 < 1.0 && eps >= 0.0)) return fail(PP_E_PARAM, "pp_adam_step");
    if (count == 0) return 0;
    return ok_or(pp::adam_step_launch(params, count, lr, beta1, beta2, eps, (cudaStream_t)stream), "pp_adam_step");
}

int pp_adam_step_allreduce(const PPAdamParam *params, int32_t count, float *flat_grad, int64_t numel, const PPPeerBlocks *peers,
                           unsigned long long *epoch, double lr, double beta1, double beta2, double eps, void *stream) {
    const char *fn = "pp_adam_step_allreduce";
    if (count <= 0 || count > 16 || numel <= 0) return fail(PP_E_SIZE, fn);
    if (!params || !flat_grad || !peers || !epoch) return fail(PP_E_NULL, fn);
    if (peers->world < 1 || peers->world > 8 || peers->rank < 0 || peers->rank >= peers->world || numel > peers->capacity_floats)
        return fail(PP_E_SIZE, fn);
    for (int r = 0; r < peers->world; ++r) if (!peers->blocks[r]) return fail(PP_E_NULL, fn);
    for (int i = 0; i < count; ++i) {
        const PPAdamParam &a = params[i];
        if (!a.param || !a.grad || !a.exp_avg || !a.exp_avg_sq || !a.step) return fail(PP_E_NULL, fn);
    }
    if (!(beta1 >= 0.0 && beta1 < 1.0 && beta2 >= 0.0 && beta2 < 1.0 && eps >= 0.0)) return fail(PP_E_PARAM, fn);
    return ok_or(pp::adam_allreduce_launch(params, count, flat_grad, numel, *peers, epoch, lr, beta1, beta2, eps, (cudaStream_t)stream), fn);
}

int64_t pp_peer_block_bytes(int64_t capacity_floats) { return capacity_floats > 0 ? (2 * capacity_floats + 8) * 4 : 0; }

// ---------------------------------------------------------------------------------------------
// DRQN training mode.
namespace {
bool rnn_params_ok(const PPQNetRNNParams *p) {
    return p && p->f0_w && p->f0_b && p->f2_w && p->f2_b && p->w_ih && p->w_hh && p->b_ih && p->b_hh &&
           noisy_ok(&p->shared, 128, 128) && noisy_ok(&p->v, 128, 1) && noisy_ok(&p->a, 128, 3);
}
}  // namespace

int pp_drqn_grads(const PPReplayRing *ring, const int64_t *rows, int32_t batch, int32_t trace, const PPQNetRNNParams *online,
                  const PPQNetRNNParams *target, int32_t noisy_online, int32_t noisy_target, float gamma,
                  const PPQNetRNNGrads *grads, float *loss_out, float *td_out, float *workspace, void *stream) {
    const char *fn = "pp_drqn_grads";
    if (batch <= 0 || batch > 256 || batch % 16 != 0 || trace <= 0 || trace > 16) return fail(PP_E_SIZE, fn);
    if (!ring || !ring->obs || !ring->act || !ring->rew || !ring->next_obs || !ring->done || ring->capacity <= 0 || !rows ||
        !grads || !workspace)
        return fail(PP_E_NULL, fn);
    if (!rnn_params_ok(online) || !rnn_params_ok(target)) return fail(PP_E_PARAM, fn);
    if (!grads->shared.grad_weight_mu) return fail(PP_E_NULL, fn);
    if (reinterpret_cast<uintptr_t>(workspace) & 15u) return fail(PP_E_ALIGN, fn);
    return ok_or(pp::drqn_grads_launch(*ring, rows, batch, trace, *online, *target, noisy_online, noisy_target, gamma, *grads,
                                       loss_out, td_out, workspace, (cudaStream_t)stream), fn);
}

int pp_seq_window_weights(const uint8_t *done, int64_t n, int64_t T, int64_t steps_written, int32_t trace, int32_t starts_fresh,
                          float *weights, unsigned long long *episodes, void *stream) {
    if (n <= 0 || T <= 0 || steps_written < 0 || trace <= 0 || trace > T) return fail(PP_E_SIZE, "pp_seq_window_weights");
    if (!done || !weights || !episodes) return fail(PP_E_NULL, "pp_seq_window_weights");
    return ok_or(pp::seq_window_weights_launch(done, n, T, steps_written, trace, starts_fresh, weights, episodes, (cudaStream_t)stream),
                 "pp_seq_window_weights");
}

int pp_seq_expand_rows(const int64_t *end_slots, int32_t batch, int32_t trace, int64_t n, int64_t T, int64_t *rows, void *stream) {
    if (batch <= 0 || trace <= 0 || n <= 0 || T <= 0 || trace > T) return fail(PP_E_SIZE, "pp_seq_expand_rows");
    if (!end_slots || !rows) return fail(PP_E_NULL, "pp_seq_expand_rows");
    return ok_or(pp::seq_expand_rows_launch(end_slots, batch, trace, n, T, rows, (cudaStream_t)stream), "pp_seq_expand_rows");
}

int pp_pack_qnetrnn_tc(const PPQNetRNNParams *net, int32_t noisy, void *image, void *stream) {
    if (!rnn_params_ok(net)) return fail(PP_E_PARAM, "pp_pack_qnetrnn_tc");
    if (!image) return fail(PP_E_NULL, "pp_pack_qnetrnn_tc");
    if (reinterpret_cast<uintptr_t>(image) & 15u) return fail(PP_E_ALIGN, "pp_pack_qnetrnn_tc");
    return ok_or(pp::pack_qnetrnn_tc_launch(*net, noisy, image, (cudaStream_t)stream), "pp_pack_qnetrnn_tc");
}

int64_t pp_drqn_workspace_floats(int32_t batch, int32_t trace) {
    return (batch > 0 && trace > 0) ? pp::drqn_workspace_floats(batch, trace) : 0;
}

int pp_clip_grad_norm(float *flat_grads, int64_t numel, float max_norm, float *norm_out, float *scratch, void *stream) {
    if (numel < 0) return fail(PP_E_SIZE, "pp_clip_grad_norm");
    if (!flat_grads || !norm_out || !scratch) return fail(PP_E_NULL, "pp_clip_grad_norm");
    if (!(max_norm > 0.f)) return fail(PP_E_PARAM, "pp_clip_grad_norm");
    if (numel == 0) return 0;
    return ok_or(pp::clip_grad_norm_launch(flat_grads, numel, max_norm, norm_out, scratch, (cudaStream_t)stream), "pp_clip_grad_norm");
}

int pp_adam_step_multi(const PPAdamParam *params, int32_t count, double lr, double beta1, double beta2, double eps, void *stream) {
    if (count < 0 || count > 32) return fail(PP_E_SIZE, "pp_adam_step_multi");
    if (count > 0 && !params) return fail(PP_E_NULL, "pp_adam_step_multi");
    for (int i = 0; i < count; ++i) {
        const PPAdamParam &a = params[i];
        if (!a.param || !a.grad || !a.exp_avg || !a.exp_avg_sq || !a.step) return fail(PP_E_NULL, "pp_adam_step_multi");
        if (a.numel < 0) return fail(PP_E_SIZE, "pp_adam_step_multi");
    }
    if (!(beta1 >= 0.0 && beta1 < 1.0 && beta2 >= 0.0 && beta2 < 1.0 && eps >= 0.0)) return fail(PP_E_PARAM, "pp_adam_step_multi");
    if (count == 0) return 0;
    return ok_or(pp::adam_multi_launch(params, count, lr, beta1, beta2, eps, (cudaStream_t)stream), "pp_adam_step_multi");
}

// ---------------------------------------------------------------------------------------------
// Host-buffer evaluation.  One context per device ordinal (stream + device staging + pinned staging, grown on demand,
// released by pp_host_release): calls for different devices run concurrently from different host threads, calls for
// the same device are serialised by the context's mutex.  Nothing is bound to "the current device" of the process.
namespace {
struct HostEvalCtx {
    std::mutex mu;
    cudaStream_t stream = nullptr;
    void *dev = nullptr;
    size_t dev_bytes = 0;
    void *pinned = nullptr;
    size_t pinned_bytes = 0;
};
constexpr int MAX_DEVICES = 64;
std::mutex g_ctx_mu;
HostEvalCtx *g_ctx[MAX_DEVICES] = {};

HostEvalCtx *ctx_for(int device) {
    std::lock_guard<std::mutex> lock(g_ctx_mu);
    if (!g_ctx[device]) g_ctx[device] = new HostEvalCtx();
    return g_ctx[device];
}

int ensure(HostEvalCtx &c, size_t dev_bytes, size_t pinned_bytes) {          // the context's device is current
    cudaError_t e;
    if (!c.stream && (e = cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking)) != cudaSuccess) return (int)e;
    if (dev_bytes > c.dev_bytes) {
        if (c.dev) cudaFree(c.dev);
        c.dev = nullptr; c.dev_bytes = 0;
        if ((e = cudaMalloc(&c.dev, dev_bytes)) != cudaSuccess) return (int)e;
        c.dev_bytes = dev_bytes;
    }
    if (pinned_bytes > c.pinned_bytes) {
        if (c.pinned) cudaFreeHost(c.pinned);
        c.pinned = nullptr; c.pinned_bytes = 0;
        if ((e = cudaMallocHost(&c.pinned, pinned_bytes)) != cudaSuccess) return (int)e;
        c.pinned_bytes = pinned_bytes;
    }
    return 0;
}
void release(HostEvalCtx &c) {
    if (c.stream) { cudaStreamSynchronize(c.stream); cudaStreamDestroy(c.stream); c.stream = nullptr; }
    if (c.dev) { cudaFree(c.dev); c.dev = nullptr; c.dev_bytes = 0; }
    if (c.pinned) { cudaFreeHost(c.pinned); c.pinned = nullptr; c.pinned_bytes = 0; }
}
size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct DeviceGuard {             // make `device` current for the call, restore the caller's device afterwards
    int prev = -1;
    cudaError_t err;
    explicit DeviceGuard(int device) {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != device) err = cudaSetDevice(device);
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
}  // namespace

int pp_host_selfplay_eval(int device, int mode, int64_t n, int32_t quota, const PPParams *params,
                          const void *host_pool_vx, const void *host_pool_vy, const void *host_pool_spin, uint64_t seed,
                          int64_t env_id_base, const float *host_weights_a, const float *host_weights_b, int32_t precision,
                          int64_t max_steps, unsigned long long *host_counters, int32_t *host_ep_log, int64_t ep_log_cap) {
    const char *fn = "pp_host_selfplay_eval";
    if (!mode_ok(mode) || (precision != PP_PREC_F32 && precision != PP_PREC_F16)) return fail(PP_E_MODE, fn);
    if (device < 0 || device >= MAX_DEVICES) return fail(PP_E_SIZE, fn);
    if (n <= 0 || quota <= 0 || max_steps <= 0 || ep_log_cap < 0 || env_id_base < 0) return fail(PP_E_SIZE, fn);
    if (!params_ok(params)) return fail(PP_E_PARAM, fn);
    const int pools = (host_pool_vx != nullptr) + (host_pool_vy != nullptr) + (host_pool_spin != nullptr);
    if ((pools != 0 && pools != 3) || !host_weights_a || !host_weights_b || !host_counters) return fail(PP_E_NULL, fn);
    const int64_t total = (int64_t)n * quota;
    if (total >= 0x7fffffff) return fail(PP_E_SIZE, fn);

    DeviceGuard guard(device);
    if (guard.err != cudaSuccess) return fail((int)guard.err, fn);
    HostEvalCtx &ctx = *ctx_for(device);
    std::lock_guard<std::mutex> lock(ctx.mu);

    const size_t rs = mode == PP_MODE_F64 ? 8 : 4;
    const size_t pool_bytes = pools ? align_up((size_t)total * rs, 256) : 0;
    const size_t real_bytes = align_up((size_t)n * rs, 256), int_bytes = align_up((size_t)n * 4, 256);
    const size_t blob_bytes = align_up(PP_QNET_BLOB_FLOATS * sizeof(float), 256);
    const size_t in_bytes = 2 * blob_bytes + 256;                             // [weights A][weights B][queue head]
    const size_t log_bytes = align_up((size_t)(host_ep_log ? ep_log_cap : 0) * 16, 256);
    const size_t dev_bytes = 3 * pool_bytes + 7 * real_bytes + 5 * int_bytes + in_bytes + 256 + log_bytes;
    int rc = ensure(ctx, dev_bytes, in_bytes + 256);                          // pinned: the inputs + counters[8] + log count
    if (rc) return fail(rc, fn);
    cudaStream_t st = ctx.stream;
    char *d = (char *)ctx.dev;
    auto take = [&](size_t b) { char *p = d; d += b; return (void *)p; };
    char *d_in = (char *)take(in_bytes);
    float *wa = (float *)d_in, *wb = (float *)(d_in + blob_bytes);
    unsigned long long *qhead = (unsigned long long *)(d_in + 2 * blob_bytes);
    unsigned long long *ctr = (unsigned long long *)take(256);                // counters[8] + ep_log_count
    PPEnvState es{};
    es.ball_x = take(real_bytes); es.ball_y = take(real_bytes); es.ball_vx = take(real_bytes); es.ball_vy = take(real_bytes);
    es.spin = take(real_bytes); es.top_paddle_x = take(real_bytes); es.bottom_paddle_x = take(real_bytes);
    es.score_a = (int32_t *)take(int_bytes); es.score_b = (int32_t *)take(int_bytes); es.bounce_count = (int32_t *)take(int_bytes);
    es.ep_idx = (int32_t *)take(int_bytes); es.ep_len = (int32_t *)take(int_bytes);
    int32_t *dlog = host_ep_log ? (int32_t *)take(log_bytes) : nullptr;
    void *pvx = pools ? take(pool_bytes) : nullptr, *pvy = pools ? take(pool_bytes) : nullptr, *psp = pools ? take(pool_bytes) : nullptr;

    cudaError_t e;
#define CK(x) do { if ((e = (x)) != cudaSuccess) return fail((int)e, fn); } while (0)
    // the small inputs travel as ONE copy from the pinned staging block: both weight blobs + the queue cursor
    char *h_in = (char *)ctx.pinned;
    unsigned long long *h_out = (unsigned long long *)(h_in + in_bytes);
    memcpy(h_in, host_weights_a, PP_QNET_BLOB_FLOATS * sizeof(float));
    memcpy(h_in + blob_bytes, host_weights_b, PP_QNET_BLOB_FLOATS * sizeof(float));
    *(unsigned long long *)(h_in + 2 * blob_bytes) = (unsigned long long)(n < total ? n : total);   // serves 0..n-1 are taken at reset
    CK(cudaMemcpyAsync(d_in, h_in, in_bytes, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(ctr, 0, 256, st));
    if (pools) {
        CK(cudaMemcpyAsync(pvx, host_pool_vx, (size_t)total * rs, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(pvy, host_pool_vy, (size_t)total * rs, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(psp, host_pool_spin, (size_t)total * rs, cudaMemcpyHostToDevice, st));
    }
    // The n x quota serves form ONE queue (PP_SERVE_QUEUE): an env that finishes an episode claims the next unplayed
    // serve, so no env idles while others still have episodes to play; outcomes per serve are unchanged.  The whole
    // evaluation is a single launch: every group leaves the step loop once the queue is drained and its envs are done.
    PPServeSource src{PP_SERVE_QUEUE, quota, pvx, pvy, psp, seed, qhead, total};
    rc = pp::env_reset_launch(mode, n, *params, es, nullptr, src, env_id_base, /*advance=*/0, st);   // env i starts on serve i
    if (rc) return fail(rc, fn);
    PPPolicy pa{PP_POLICY_QNET, precision, 0, 0.0, wa, nullptr, nullptr};
    PPPolicy pb{PP_POLICY_QNET, precision, 0, 0.0, wb, nullptr, nullptr};
    PPRolloutOut out{ctr, dlog, host_ep_log ? ep_log_cap : 0, ctr + 8, nullptr, nullptr, nullptr};
    const int64_t k = max_steps < 0x7fffffff ? max_steps : 0x7ffffffe;
    rc = precision == PP_PREC_F16
             ? pp::selfplay_tc_launch(mode, n, k, *params, es, pa, pb, seed, 0, src, (int32_t)total, env_id_base, out, nullptr, st)
             : pp::selfplay_launch(mode, n, k, *params, es, pa, pb, seed, 0, src, (int32_t)total, env_id_base, out, nullptr, st);
    if (rc) return fail(rc, fn);
    CK(cudaMemcpyAsync(h_out, ctr, 9 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    memcpy(host_counters, h_out, 8 * sizeof(unsigned long long));
    if (host_ep_log) {
        const unsigned long long rows = h_out[8] < (unsigned long long)ep_log_cap ? h_out[8] : (unsigned long long)ep_log_cap;
        CK(cudaMemcpyAsync(host_ep_log, dlog, rows * 16, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    }
#undef CK
    return 0;
}

int pp_host_release(int device) {
    if (device < -1 || device >= MAX_DEVICES) return fail(PP_E_SIZE, "pp_host_release");
    for (int dv = 0; dv < MAX_DEVICES; ++dv) {
        if (device != -1 && dv != device) continue;
        HostEvalCtx *c;
        { std::lock_guard<std::mutex> lock(g_ctx_mu); c = g_ctx[dv]; }
        if (!c) continue;
        std::lock_guard<std::mutex> lock(c->mu);
        if (!c->stream && !c->dev && !c->pinned) continue;
        DeviceGuard guard(dv);
        if (guard.err != cudaSuccess) return fail((int)guard.err, "pp_host_release");
        release(*c);
    }
    return 0;
}

}  // extern "C"
