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

// One warp's 32 replay rows of 7 floats (896 contiguous bytes when the slots are consecutive) through a shared-memory
// staging buffer: two rounds of 16-byte stores instead of 7 scalar stores that each touch 28 sectors.
__device__ __forceinline__ void store_rows_staged(float *__restrict__ stage, float *__restrict__ dst_rows, const float (&v)[7], int lane) {
#pragma unroll
    for (int k = 0; k < 7; ++k) stage[lane * 7 + k] = v[k];                    // stride 7: conflict-free
    __syncwarp();
    const float4 *s4 = reinterpret_cast<const float4 *>(stage);
    float4 *d4 = reinterpret_cast<float4 *>(dst_rows);
    d4[lane] = s4[lane];
    if (lane < 24) d4[32 + lane] = s4[32 + lane];
    __syncwarp();
}

// Called by ALL lanes of a warp.  `ob` = player B's observation the action was chosen from; `serve(ep, vx, vy, spin)`
// yields the serve of episode `ep` of this env.  `stage`: 224 floats of shared memory private to the warp (16-byte
// aligned) for vectorised replay rows, or nullptr.
template <typename R, bool PRECLAIM = false, typename ServeFn>
__device__ __forceinline__ void step_and_book(const EnvConsts<R> &c, Lane<R> &L, bool active, int act_a, int act_b,
                                              const float (&ob)[7], int64_t t, int64_t n, int64_t i, int64_t env_id_base,
                                              int32_t quota, const PPRolloutOut &out, const PPReplayRing &ring, bool ring_on,
                                              const PPServeSource &src, ServeFn &&serve, float *stage = nullptr,
                                              int preclaimed = -1) {
    const int lane = threadIdx.x & 31;
    int flags = 0;
    if (active) {
        flags = env_step<R>(c, L.e, act_a, act_b);
        L.ep_len += 1;
        L.tally.add_flags(flags);
        if (out.actions_out)
            reinterpret_cast<uchar2 *>(out.actions_out)[t * n + i] = make_uchar2((unsigned char)act_a, (unsigned char)act_b);
    }
    if (ring_on) {       // memory.push((oB, aB, rB, nB, done)) scripts/train_iterative.py:243 / push_step train_rnn_iterative.py:764
        const unsigned m = __ballot_sync(0xffffffffu, active);
        if (m) {
            int64_t slot;
            bool whole = m == 0xffffffffu;                                     // 32 consecutive slots, no wrap inside the warp
            if (ring.lockstep_envs) {                                          // [time][env] layout: no cursor
                const int64_t steps = ring.capacity / ring.lockstep_envs;
                slot = ((ring.lockstep_step0 + t) % steps) * ring.lockstep_envs + i;
            } else {                                                           // rows compacted per warp, one cursor atomic
                unsigned long long base = 0;
                const int leader = __ffs(m) - 1;
                if (lane == leader) base = atomicAdd(ring.head, (unsigned long long)__popc(m));
                base = __shfl_sync(0xffffffffu, base, leader) % (unsigned long long)ring.capacity;
                whole = whole && base + 32 <= (unsigned long long)ring.capacity;
                slot = (int64_t)((base + __popc(m & ((1u << lane) - 1u))) % (unsigned long long)ring.capacity);
            }
            float na[7], nb[7];
            observe<R>(L.e, na, nb);
            const int64_t slot0 = __shfl_sync(0xffffffffu, slot, 0);
            if (stage != nullptr && whole && (slot0 & 3) == 0) {               // warp-uniform; rows start 16-byte aligned
                store_rows_staged(stage, ring.obs + slot0 * 7, ob, lane);
                store_rows_staged(stage, ring.next_obs + slot0 * 7, nb, lane);
            } else if (active) {
#pragma unroll
                for (int k = 0; k < 7; ++k) { ring.obs[slot * 7 + k] = ob[k]; ring.next_obs[slot * 7 + k] = nb[k]; }
            }
            if (active) {
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
    // PP_SERVE_QUEUE: a finishing env takes the serve it claimed ahead of time (`preclaimed` >= 0: its serve is already
    // drawn), the others claim now
    int claimed = 0;
    if (PRECLAIM) {
        claimed = preclaimed;
        const unsigned m_claim = __ballot_sync(0xffffffffu, fin && preclaimed < 0);
        if (queue && m_claim) {                                                // warp-uniform branch
            const int c = claim_serves(fin && preclaimed < 0, m_claim, src);
            if (preclaimed < 0) claimed = c;
        }
    } else if (queue && m_fin) claimed = claim_serves(fin, m_fin, src);        // warp-uniform branch
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
