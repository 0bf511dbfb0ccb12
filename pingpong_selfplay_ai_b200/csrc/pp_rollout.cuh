// pp_rollout.cuh — the part of a lock-step iteration that follows action selection, shared by the CUDA-core and the
// tensor-core self-play kernels: env step, counters, replay row, episode log, auto-reset.
#pragma once
#include "pp_policy.cuh"

namespace pp {

template <typename R> struct Lane {
    Env<R> e;
    int ep_idx, ep_len;
    Tally tally;
};

// Replay rows of a launch that pushes more than `capacity` rows: only the last capacity / n lock-step steps are
// written (earlier rows would be overwritten before the launch ends; writing them concurrently with their
// successors could tear a slot).  Written rows take consecutive cursor values, so their slots are distinct.
__device__ __forceinline__ int64_t ring_first_step(const PPReplayRing &ring, int64_t n, int64_t k_steps) {
    return ring.head ? k_steps - ring.capacity / n : 0;
}

// Called by ALL lanes of a warp.  `ob` = player B's observation the action was chosen from; `serve(ep, vx, vy, spin)`
// yields the serve of episode `ep` of this env.
template <typename R, typename ServeFn>
__device__ __forceinline__ void step_and_book(const EnvConsts<R> &c, Lane<R> &L, bool active, int act_a, int act_b,
                                              const float (&ob)[7], int64_t t, int64_t n, int64_t i, int64_t env_id_base,
                                              int32_t quota, const PPRolloutOut &out, const PPReplayRing &ring, bool ring_on,
                                              const PPServeSource &src, ServeFn &&serve) {
    const int lane = threadIdx.x & 31;
    int flags = 0;
    if (active) {
        flags = env_step<R>(c, L.e, act_a, act_b);
        L.ep_len += 1;
        L.tally.add_flags(flags);
        if (out.actions_out)
            reinterpret_cast<uchar2 *>(out.actions_out)[t * n + i] = make_uchar2((unsigned char)act_a, (unsigned char)act_b);
    }
    if (ring_on && ring.lockstep_envs) {   // memory.push_step scripts/train_rnn_iterative.py:764: [time][env] layout
        if (active) {
            const int64_t steps = ring.capacity / ring.lockstep_envs;
            const int64_t slot = ((ring.lockstep_step0 + t) % steps) * ring.lockstep_envs + i;
            float na[7], nb[7];
            observe<R>(L.e, na, nb);
#pragma unroll
            for (int k = 0; k < 7; ++k) { ring.obs[slot * 7 + k] = ob[k]; ring.next_obs[slot * 7 + k] = nb[k]; }
            ring.act[slot] = (uint8_t)act_b;
            ring.rew[slot] = (flags & F_POINT_B) ? 1.0f : ((flags & F_POINT_A) ? -1.0f : 0.0f);
            ring.done[slot] = (uint8_t)(flags & F_DONE);
        }
    } else if (ring_on) {   // memory.push((oB, aB, rB, nB, done)) scripts/train_iterative.py:243, rows compacted per warp
        const unsigned m = __ballot_sync(0xffffffffu, active);
        if (m) {
            unsigned long long base = 0;
            const int leader = __ffs(m) - 1;
            if (lane == leader) base = atomicAdd(ring.head, (unsigned long long)__popc(m));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (active) {
                const int64_t slot = (int64_t)((base + __popc(m & ((1u << lane) - 1u))) % (unsigned long long)ring.capacity);
                float na[7], nb[7];
                observe<R>(L.e, na, nb);
#pragma unroll
                for (int k = 0; k < 7; ++k) { ring.obs[slot * 7 + k] = ob[k]; ring.next_obs[slot * 7 + k] = nb[k]; }
                ring.act[slot] = (uint8_t)act_b;
                ring.rew[slot] = (flags & F_POINT_B) ? 1.0f : ((flags & F_POINT_A) ? -1.0f : 0.0f);
                ring.done[slot] = (uint8_t)(flags & F_DONE);
            }
        }
    }
    const bool fin = (flags & F_DONE) != 0;
    const unsigned m_fin = __ballot_sync(0xffffffffu, fin);
    const bool queue = src.kind == PP_SERVE_QUEUE;
    if (queue) log_episode(fin, m_fin, out, (int)(env_id_base + L.ep_idx % n), (int)(L.ep_idx / n), L.e.sa, L.e.sb, L.ep_len);
    else log_episode(fin, m_fin, out, (int)(env_id_base + i), L.ep_idx, L.e.sa, L.e.sb, L.ep_len);
    int claimed = 0;
    if (queue && m_fin) claimed = claim_serves(fin, m_fin, src);               // warp-uniform branch
    if (fin) {
        L.tally.episodes += 1;
        if (L.e.sa > L.e.sb) L.tally.wins_a += 1; else L.tally.wins_b += 1;
        L.tally.len_sum += (unsigned)L.ep_len;
        L.ep_idx = queue ? claimed : L.ep_idx + 1;
        if (!(quota > 0 && L.ep_idx >= quota)) {
            R svx, svy, ssp;
            serve(L.ep_idx, svx, svy, ssp);
            serve_env<R>(L.e, svx, svy, ssp);
            L.ep_len = 0;
        }
    }
}

}  // namespace pp
