// selfplay_kernels.cu — K2a (obs -> QNet -> epsilon-greedy argmax) and the fused self-play rollout.
//
//   qnet_act_kernel   standalone action selection for one player from materialised obs[n][7]
//                     (the reference's per-step model call, scripts/train_iterative.py:124-130,176-177).
//   selfplay_kernel   k lock-step iterations of {obs, act A, act B, step, replay row, auto-reset} with the
//                     env state, both observations and all bookkeeping in registers and both weight
//                     blobs in shared memory: nothing but the replay rows and the episode log touches
//                     HBM between the first and last step of a launch.
#include "pp_rollout.cuh"
#include "pp_host.h"

namespace pp {

constexpr int ACT_THREADS = 128;

__global__ void __launch_bounds__(ACT_THREADS)
qnet_act_kernel(int64_t n, const float *__restrict__ obs, const PPPolicy pol, uint64_t seed, uint32_t step_index,
                int64_t env_id_base, uint32_t stream_id, uint8_t *__restrict__ actions, float *__restrict__ q_out) {
    __shared__ __align__(16) float s_w[PP_QNET_BLOB_FLOATS];
    __shared__ __align__(16) float s_obs[ACT_THREADS * 7];
    if (pol.kind == PP_POLICY_QNET) stage_blob(s_w, pol.weights, PP_QNET_BLOB_FLOATS);
    const int64_t tile0 = (int64_t)blockIdx.x * ACT_THREADS;
    const int rows = (int)((n - tile0) < ACT_THREADS ? (n - tile0) : ACT_THREADS);
    for (int w = threadIdx.x; w < rows * 7; w += ACT_THREADS) s_obs[w] = obs[tile0 * 7 + w];   // coalesced rows
    __syncthreads();
    const int64_t i = tile0 + threadIdx.x;
    if (i >= n) return;
    float o[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) o[k] = s_obs[threadIdx.x * 7 + k];
    const uint32_t g = (uint32_t)(env_id_base + i);
    int a;
    float q[3] = {0.f, 0.f, 0.f};
    if (pol.kind == PP_POLICY_RANDOM) {
        a = random_action(seed, g, step_index, stream_id);
    } else {
        if (pol.kind == PP_POLICY_QNET) { qnet_forward(s_w, o, q); a = argmax3(q); }
        else a = follower_action(o, pol.follower_tol);
        a = explore(a, pol.eps_threshold, seed, g, step_index, stream_id);
    }
    actions[i] = (uint8_t)a;
    if (q_out) { q_out[i * 3 + 0] = q[0]; q_out[i * 3 + 1] = q[1]; q_out[i * 3 + 2] = q[2]; }
}

constexpr int SP_THREADS = 128;

// Action of one player from its observation; `sw` = that player's blob in shared memory.
__device__ __forceinline__ int select_action(const PPPolicy &pol, const float *sw, const float (&o)[7], uint64_t seed,
                                             uint32_t g, uint32_t step, uint32_t stream_id) {
    if (pol.kind == PP_POLICY_RANDOM) return random_action(seed, g, step, stream_id);
    int a;
    if (pol.kind == PP_POLICY_QNET) {
        float q[3];
        qnet_forward(sw, o, q);
        a = argmax3(q);
    } else {
        a = follower_action(o, pol.follower_tol);
    }
    return explore(a, pol.eps_threshold, seed, g, step, stream_id);
}

template <typename R>
__global__ void __launch_bounds__(SP_THREADS, 4)
selfplay_kernel(const PPParams params, const PPEnvState st, int64_t n, int64_t k_steps, const PPPolicy pol_a,
                const PPPolicy pol_b, uint64_t seed, int64_t step_base, const PPServeSource src, int32_t quota,
                int64_t env_id_base, const PPRolloutOut out, const PPReplayRing ring) {
    __shared__ __align__(16) float s_w[2][PP_QNET_BLOB_FLOATS];
    if (pol_a.kind == PP_POLICY_QNET) stage_blob(s_w[0], pol_a.weights, PP_QNET_BLOB_FLOATS);
    if (pol_b.kind == PP_POLICY_QNET) stage_blob(s_w[1], pol_b.weights, PP_QNET_BLOB_FLOATS);
    __syncthreads();

    const EnvConsts<R> c(params);
    const StatePtrs<R> s(st);
    const int64_t i = (int64_t)blockIdx.x * SP_THREADS + threadIdx.x;
    const bool valid = i < n;
    const int64_t ic = valid ? i : 0;
    Lane<R> L;
    L.e = load_env<R>(s, ic);
    L.ep_idx = s.ep_idx[ic]; L.ep_len = s.ep_len[ic];
    const uint32_t g = (uint32_t)(env_id_base + ic);
    const int64_t ring_t0 = ring_first_step(ring, n, k_steps);

#pragma unroll 1
    for (int64_t t = 0; t < k_steps; ++t) {
        const bool active = valid && !(quota > 0 && L.ep_idx >= quota);
        const uint32_t step = (uint32_t)(step_base + t);
        float oa[7], ob[7];
        observe<R>(L.e, oa, ob);
        int act_a = 1, act_b = 1;
        // a warp whose envs are all frozen by the quota is finished (frozen envs stay frozen within a launch)
        if (!__any_sync(0xffffffffu, active)) break;
        {
#pragma unroll 1
            for (int p = 0; p < 2; ++p) {          // one copy of the MLP code serves both players
                float o[7];
#pragma unroll
                for (int k = 0; k < 7; ++k) o[k] = p ? ob[k] : oa[k];
                const int a = select_action(p ? pol_b : pol_a, s_w[p], o, seed, g, step, p ? STREAM_ACT_B : STREAM_ACT_A);
                if (p) act_b = a; else act_a = a;
            }
        }
        step_and_book<R>(c, L, active, act_a, act_b, ob, t, n, i, env_id_base, quota, out, ring,
                         ring.head != nullptr && t >= ring_t0, src,
                         [&](int ep, R &vx, R &vy, R &sp) { next_serve<R>(params, src, n, i, env_id_base, ep, vx, vy, sp); });
    }
    if (valid) {
        store_env<R>(s, i, L.e);
        s.ep_idx[i] = L.ep_idx;
        s.ep_len[i] = L.ep_len;
    }
    L.tally.flush(out.counters, out.ep_log ? nullptr : out.ep_log_count);
}

int qnet_act_launch(int64_t n, const float *obs, const PPPolicy &pol, uint64_t seed, int64_t step_index,
                    int64_t env_id_base, int32_t stream_id, uint8_t *actions, float *q_out, cudaStream_t stream) {
    const unsigned blocks = (unsigned)((n + ACT_THREADS - 1) / ACT_THREADS);
    qnet_act_kernel<<<blocks, ACT_THREADS, 0, stream>>>(n, obs, pol, seed, (uint32_t)step_index, env_id_base,
                                                        (uint32_t)stream_id, actions, q_out);
    return (int)cudaGetLastError();
}

int selfplay_launch(int mode, int64_t n, int64_t k, const PPParams &p, const PPEnvState &st, const PPPolicy &pa,
                    const PPPolicy &pb, uint64_t seed, int64_t step_base, const PPServeSource &src, int32_t quota,
                    int64_t env_id_base, const PPRolloutOut &out, const PPReplayRing *ring, cudaStream_t stream) {
    const unsigned blocks = (unsigned)((n + SP_THREADS - 1) / SP_THREADS);
    PPReplayRing r{};
    if (ring) r = *ring;
    if (mode == PP_MODE_F64)
        selfplay_kernel<double><<<blocks, SP_THREADS, 0, stream>>>(p, st, n, k, pa, pb, seed, step_base, src, quota,
                                                                   env_id_base, out, r);
    else
        selfplay_kernel<float><<<blocks, SP_THREADS, 0, stream>>>(p, st, n, k, pa, pb, seed, step_base, src, quota,
                                                                  env_id_base, out, r);
    return (int)cudaGetLastError();
}

}  // namespace pp
