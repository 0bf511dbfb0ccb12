// pp_policy.cuh — per-env action selection on CUDA cores (PP_PREC_F32).
//
// QNet (models/qnet.py:52-75): 7 -> 64 -> 64 -> dueling (V 1, A 3).  One env per thread; the packed
// k-major weight blob (PP_QNET_* offsets, 19.7 KB) lives in shared memory and every lane of a warp
// reads the same weight row, so each LDS.128 is a 4-weight broadcast feeding 4 FFMAs.
// Accumulation order is DEFINED: acc = bias; acc = fmaf(W[j][k], x[k], acc) for k ascending — the C
// oracle (oracle/pong_oracle.c: dense()) executes the same chain, so Q-values and therefore greedy
// actions are bit-identical to the oracle, which in turn is within 1e-5 of torch.
#pragma once
#include "pp_device.cuh"

namespace pp {

__device__ __forceinline__ float relu(float v) { return v > 0.0f ? v : 0.0f; }

// acc += w * x for four outputs: two packed FFMA2 (sm_100: two IEEE fp32 FMAs per instruction — the same bits as four
// fmaf, half the issue slots; this kernel is issue-bound)
__device__ __forceinline__ void fma4(float4 &acc, const float4 &w, float x) {
    const float2 xx = make_float2(x, x);
    const float2 lo = __ffma2_rn(make_float2(w.x, w.y), xx, make_float2(acc.x, acc.y));
    const float2 hi = __ffma2_rn(make_float2(w.z, w.w), xx, make_float2(acc.z, acc.w));
    acc = make_float4(lo.x, lo.y, hi.x, hi.y);
}

// sw: blob in shared memory (16-byte aligned).  q[3] out.
__device__ __forceinline__ void qnet_forward(const float *__restrict__ sw, const float (&obs)[7], float (&q)[3]) {
    float h1[64];
    {
        const float4 *w1 = reinterpret_cast<const float4 *>(sw + PP_QNET_W1T);
        const float4 *b1 = reinterpret_cast<const float4 *>(sw + PP_QNET_B1);
#pragma unroll
        for (int j4 = 0; j4 < 16; ++j4) {
            float4 acc = b1[j4];
#pragma unroll
            for (int k = 0; k < 7; ++k) {
                fma4(acc, w1[k * 16 + j4], obs[k]);
            }
            h1[j4 * 4 + 0] = relu(acc.x); h1[j4 * 4 + 1] = relu(acc.y);
            h1[j4 * 4 + 2] = relu(acc.z); h1[j4 * 4 + 3] = relu(acc.w);
        }
    }
    float4 head = *reinterpret_cast<const float4 *>(sw + PP_QNET_BH);      // (V, A0, A1, A2) accumulators
    const float4 *w2 = reinterpret_cast<const float4 *>(sw + PP_QNET_W2T);
    const float4 *b2 = reinterpret_cast<const float4 *>(sw + PP_QNET_B2);
    const float4 *wh = reinterpret_cast<const float4 *>(sw + PP_QNET_WHT);
#pragma unroll 1
    for (int jb = 0; jb < 4; ++jb) {              // 16 hidden units per pass keeps the live set ~100 registers
        float4 a0 = b2[jb * 4 + 0], a1 = b2[jb * 4 + 1], a2 = b2[jb * 4 + 2], a3 = b2[jb * 4 + 3];
#pragma unroll
        for (int k = 0; k < 64; ++k) {
            const float x = h1[k];
            const float4 u0 = w2[k * 16 + jb * 4 + 0], u1 = w2[k * 16 + jb * 4 + 1];
            const float4 u2 = w2[k * 16 + jb * 4 + 2], u3 = w2[k * 16 + jb * 4 + 3];
            fma4(a0, u0, x); fma4(a1, u1, x); fma4(a2, u2, x); fma4(a3, u3, x);
        }
        const float h2[16] = {relu(a0.x), relu(a0.y), relu(a0.z), relu(a0.w), relu(a1.x), relu(a1.y), relu(a1.z), relu(a1.w),
                              relu(a2.x), relu(a2.y), relu(a2.z), relu(a2.w), relu(a3.x), relu(a3.y), relu(a3.z), relu(a3.w)};
#pragma unroll
        for (int i = 0; i < 16; ++i) {            // heads accumulate in ascending hidden index, like the oracle
            fma4(head, wh[jb * 16 + i], h2[i]);
        }
    }
    // V + (A - mean(A))                                                       models/qnet.py:75
    const float mean = __fdiv_rn(__fadd_rn(__fadd_rn(head.y, head.z), head.w), 3.0f);
    q[0] = __fadd_rn(head.x, __fsub_rn(head.y, mean));
    q[1] = __fadd_rn(head.x, __fsub_rn(head.z, mean));
    q[2] = __fadd_rn(head.x, __fsub_rn(head.w, mean));
}

__device__ __forceinline__ int argmax3(const float (&q)[3]) {     // first maximum wins (torch.argmax)
    int best = 0;
    float m = q[0];
    if (q[1] > m) { best = 1; m = q[1]; }
    if (q[2] > m) { best = 2; }
    return best;
}

// Greedy QNet action with the blob read straight from GLOBAL memory (all lanes read the same words: broadcast
// loads that hit L1).  Used where a QNet meets a recurrent player (tests/arena.py pairings QNet x QNetRNN): the
// recurrent player's tiles own the shared memory there.  Same fmaf chain as above, so the same bits.
static __device__ __noinline__ int qnet_greedy_global(const float *__restrict__ blob, const float (&obs)[7]) {
    float q[3];
    qnet_forward(blob, obs, q);
    return argmax3(q);
}

// HardcodedBallFollower                                                      tests/arena.py:211-217
// obs are np.float32 scalars and the tolerance a Python float: under the reference's pinned numpy 1.24.3 (value-based
// scalar promotion) `my_paddle_x - tolerance` and both compares are evaluated in float64.
__device__ __forceinline__ int follower_action(const float (&obs)[7], double tol) {
    const double x = (double)obs[0], pad = (double)obs[4];
    const double lo = __dsub_rn(pad, tol), hi = __dadd_rn(pad, tol);
    return x < lo ? 0 : (x > hi ? 2 : 1);
}

// epsilon-greedy overlay                                        scripts/train_iterative.py:124-130
__device__ __forceinline__ int explore(int greedy, uint64_t eps_threshold, uint64_t seed, uint32_t env_id, uint32_t step,
                                       uint32_t stream_id) {
    if (eps_threshold == 0) return greedy;
    const uint4 r = philox4x32_10(env_id, step, stream_id, 0u, (uint32_t)seed, (uint32_t)(seed >> 32));
    return ((uint64_t)r.x < eps_threshold) ? (int)(((uint64_t)r.y * 3u) >> 32) : greedy;
}

__device__ __forceinline__ int random_action(uint64_t seed, uint32_t env_id, uint32_t step, uint32_t stream_id) {
    const uint4 r = philox4x32_10(env_id, step, stream_id, 0u, (uint32_t)seed, (uint32_t)(seed >> 32));
    return (int)(((uint64_t)r.y * 3u) >> 32);
}

// cooperative copy of a blob into shared memory (n_floats % 4 == 0, both 16-byte aligned)
__device__ __forceinline__ void stage_blob(float *dst, const float *__restrict__ src, int n_floats) {
    const float4 *s4 = reinterpret_cast<const float4 *>(src);
    float4 *d4 = reinterpret_cast<float4 *>(dst);
    for (int i = threadIdx.x; i < n_floats / 4; i += blockDim.x) d4[i] = __ldg(s4 + i);
}

}  // namespace pp
