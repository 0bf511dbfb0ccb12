// lstm_kernels.cu — K2b: one QNetRNN step (models/qnet_rnn.py:107-144, seq_len 1) for n envs with carried
// per-env (h, c), PP_PREC_F32 path on CUDA cores.
//
// A CTA owns a tile of 64 envs and runs the whole network for them; activations stay in shared memory
// ([k][env], so a thread's 8 envs are two 128-bit loads) and the three large weight matrices are
// streamed from L2 in 32 KB k-chunks with cp.async double buffering:
//     Wf2t [64][128]   features 64 -> 128
//     Wgt  [256][512]  LSTM gates from [features ; h_prev], columns interleaved as unit*4 + gate so the
//                      thread that accumulated a unit's i,f,g,o also applies the cell update
//     Wst  [128][128]  noisy shared head
// Thread tile: 8 envs x (NCOLS/64) columns, accumulated with fmaf in ascending k from the bias — the
// oracle's chain (oracle/pong_oracle.c: oracle_qnetrnn_forward); expf/tanhf are the only difference.
// (h, c) are stored unit-major [128][n] so that loads and stores are coalesced along the env index.
#include "pp_host.h"
#include "pp_rollout.cuh"

namespace pp {

constexpr int L_TILE = 64;            // envs per CTA
constexpr int L_THREADS = 512;
constexpr int L_CHUNK = 8192;         // floats per streamed weight chunk (32 KB)
constexpr int L_F1 = 64, L_F = 128, L_H = 128, L_S = 128;

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void issue_chunk(float *buf, const float *__restrict__ g) {
#pragma unroll
    for (int r = 0; r < L_CHUNK / 4 / L_THREADS; ++r) {
        const int idx = (r * L_THREADS + threadIdx.x) * 4;
        cp_async16(buf + idx, g + idx);
    }
    cp_async_commit();
}

// acc[e][c] (+)= sum_k A[k][eg*8+e] * W[k][cg*CT+c] over K rows of a k-major weight matrix in global
// memory, streamed through wbuf[2].  acc must be pre-loaded with the bias.
template <int NCOLS, int K>
__device__ __forceinline__ void streamed_gemm(const float *__restrict__ A, const float *__restrict__ gW, float *wbuf,
                                              float (&acc)[8][NCOLS / 64], int eg, int cg) {
    constexpr int CT = NCOLS / 64;
    constexpr int KC = L_CHUNK / NCOLS;          // rows per chunk
    constexpr int NCH = K / KC;
    issue_chunk(wbuf, gW);
#pragma unroll 1
    for (int ch = 0; ch < NCH; ++ch) {
        float *cur = wbuf + (ch & 1) * L_CHUNK;
        if (ch + 1 < NCH) { issue_chunk(wbuf + ((ch + 1) & 1) * L_CHUNK, gW + (size_t)(ch + 1) * L_CHUNK); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncthreads();
        const float *a_row = A + (size_t)ch * KC * L_TILE + eg * 8;
        const float *w_row = cur + cg * CT;
#pragma unroll 4
        for (int kk = 0; kk < KC; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4 *>(a_row + kk * L_TILE);
            const float4 a1 = *reinterpret_cast<const float4 *>(a_row + kk * L_TILE + 4);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float w[CT];
            if constexpr (CT == 8) {
                const float4 w0 = *reinterpret_cast<const float4 *>(w_row + kk * NCOLS);
                const float4 w1 = *reinterpret_cast<const float4 *>(w_row + kk * NCOLS + 4);
                w[0] = w0.x; w[1] = w0.y; w[2] = w0.z; w[3] = w0.w; w[4] = w1.x; w[5] = w1.y; w[6] = w1.z; w[7] = w1.w;
            } else {
                const float2 w0 = *reinterpret_cast<const float2 *>(w_row + kk * NCOLS);
                w[0] = w0.x; w[1] = w0.y;
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {        // packed FFMA2 (sm_100): the same bits as fmaf, half the issue slots
                const float2 aa = make_float2(a[e], a[e]);
#pragma unroll
                for (int c = 0; c < CT; c += 2) {
                    const float2 r = __ffma2_rn(make_float2(w[c], w[c + 1]), aa, make_float2(acc[e][c], acc[e][c + 1]));
                    acc[e][c] = r.x; acc[e][c + 1] = r.y;
                }
            }
        }
        __syncthreads();                         // everyone is done with `cur` before it is refilled
    }
}

__device__ __forceinline__ float sigmoid_f(float v) { return 1.0f / (1.0f + expf(-v)); }

// Shared-memory carve-up of one 64-env tile (dynamic smem, L_SMEM bytes)
struct RnnTile {
    float *A;        // [256][64]: rows 0..127 features (later h_new), 128..255 h_prev (later the shared-head output)
    float *wbuf;     // 2 x 32 KB streamed weight chunks
    float *f1s;      // [64][64]
    float *obs_s;    // [7][64]
    uint8_t *reset;  // [64] episode-start flags: (h, c) = 0 before this step (init_hidden, tests/arena.py:298-299)
    __device__ __forceinline__ explicit RnnTile(float *smem)
        : A(smem), wbuf(smem + 256 * L_TILE), f1s(smem + 256 * L_TILE + 2 * L_CHUNK),
          obs_s(smem + 256 * L_TILE + 2 * L_CHUNK + L_F1 * L_TILE),
          reset(reinterpret_cast<uint8_t *>(smem + 256 * L_TILE + 2 * L_CHUNK + L_F1 * L_TILE + 7 * L_TILE)) {}
};

// One QNetRNN step (models/qnet_rnn.py:107-144, seq_len 1) for the CTA's tile.  In: tile.obs_s[k][env], tile.reset[env],
// (h, c) in global memory (unit-major [128][n], updated in place).  Out: q[3] for thread tid < rows (env = tid).
// ALL threads of the CTA call this (it contains block barriers); the caller synchronises after filling obs_s / reset.
__device__ __forceinline__ void rnn_forward(const RnnTile &tile, const float *__restrict__ W, float *__restrict__ gh,
                                            float *__restrict__ gc, int64_t n, int64_t base, int rows, float (&q)[3]) {
    float *A = tile.A, *wbuf = tile.wbuf, *f1s = tile.f1s, *obs_s = tile.obs_s;
    const uint8_t *s_reset = tile.reset;
    const int tid = threadIdx.x, eg = tid & 7, cg = tid >> 3;        // 8 env groups x 64 column groups
    for (int w = tid; w < L_H * L_TILE; w += L_THREADS) {            // h_prev -> A rows 128..255 (zero on episode start)
        const int u = w / L_TILE, env = w - u * L_TILE;
        float v = 0.0f;
        if (env < rows && !s_reset[env]) v = gh[(size_t)u * n + base + env];
        A[(128 + u) * L_TILE + env] = v;
    }
    {   // features layer 1: 7 -> 64, ReLU.  thread = 8 envs x 1 column
        float acc[8];
        const float b = __ldg(W + PP_RNN_BF1 + cg);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = b;
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            const float w = __ldg(W + PP_RNN_WF1T + k * L_F1 + cg);
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] = fmaf(w, obs_s[k * L_TILE + eg * 8 + e], acc[e]);
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) f1s[cg * L_TILE + eg * 8 + e] = relu(acc[e]);
    }
    __syncthreads();
    {   // features layer 2: 64 -> 128, ReLU -> A rows 0..127
        float acc[8][2];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const float b = __ldg(W + PP_RNN_BF2 + cg * 2 + c);
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e][c] = b;
        }
        streamed_gemm<L_F, L_F1>(f1s, W + PP_RNN_WF2T, wbuf, acc, eg, cg);
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int e = 0; e < 8; ++e) A[(cg * 2 + c) * L_TILE + eg * 8 + e] = relu(acc[e][c]);
    }
    __syncthreads();
    float hn[8][2];
    {   // gates: [features ; h_prev] (K = 256) -> 512, then the cell for units 2cg, 2cg+1
        float acc[8][8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const float b = __ldg(W + PP_RNN_BG + cg * 8 + c);
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e][c] = b;
        }
        streamed_gemm<4 * L_H, L_F + L_H>(A, W + PP_RNN_WGT, wbuf, acc, eg, cg);
#pragma unroll
        for (int uu = 0; uu < 2; ++uu) {
            const int u = cg * 2 + uu;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int env = eg * 8 + e;
                float c_prev = 0.0f;
                if (env < rows && !s_reset[env]) c_prev = gc[(size_t)u * n + base + env];
                const float ig = sigmoid_f(acc[e][uu * 4 + 0]), fg = sigmoid_f(acc[e][uu * 4 + 1]);
                const float gg = tanhf(acc[e][uu * 4 + 2]), og = sigmoid_f(acc[e][uu * 4 + 3]);
                const float cn = __fadd_rn(__fmul_rn(fg, c_prev), __fmul_rn(ig, gg));
                const float h = __fmul_rn(og, tanhf(cn));
                hn[e][uu] = h;
                if (env < rows) {
                    gc[(size_t)u * n + base + env] = cn;
                    gh[(size_t)u * n + base + env] = h;
                }
            }
        }
    }
    // streamed_gemm ended with a barrier: A rows 0..127 (features) are free -> h_new
#pragma unroll
    for (int uu = 0; uu < 2; ++uu)
#pragma unroll
        for (int e = 0; e < 8; ++e) A[(cg * 2 + uu) * L_TILE + eg * 8 + e] = hn[e][uu];
    __syncthreads();
    {   // noisy shared head 128 -> 128, ReLU -> A rows 128..255
        float acc[8][2];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const float b = __ldg(W + PP_RNN_BS + cg * 2 + c);
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e][c] = b;
        }
        streamed_gemm<L_S, L_H>(A, W + PP_RNN_WST, wbuf, acc, eg, cg);
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int e = 0; e < 8; ++e) A[(128 + cg * 2 + c) * L_TILE + eg * 8 + e] = relu(acc[e][c]);
    }
    __syncthreads();
    if (tid < rows) {   // dueling heads, one env per thread
        const float4 *wh = reinterpret_cast<const float4 *>(W + PP_RNN_WHT);
        float4 head = __ldg(reinterpret_cast<const float4 *>(W + PP_RNN_BH));
#pragma unroll 8
        for (int k = 0; k < L_S; ++k) {
            const float x = A[(128 + k) * L_TILE + tid];
            const float4 w = __ldg(wh + k);
            head.x = fmaf(w.x, x, head.x); head.y = fmaf(w.y, x, head.y);
            head.z = fmaf(w.z, x, head.z); head.w = fmaf(w.w, x, head.w);
        }
        const float mean = __fdiv_rn(__fadd_rn(__fadd_rn(head.y, head.z), head.w), 3.0f);
        q[0] = __fadd_rn(head.x, __fsub_rn(head.y, mean));
        q[1] = __fadd_rn(head.x, __fsub_rn(head.z, mean));
        q[2] = __fadd_rn(head.x, __fsub_rn(head.w, mean));
    }
}

__global__ void __launch_bounds__(L_THREADS, 1)
qnetrnn_act_kernel(int64_t n, const float *__restrict__ obs, const PPPolicy pol, const uint8_t *__restrict__ reset_mask,
                   uint64_t seed, uint32_t step_index, int64_t env_id_base, uint32_t stream_id,
                   uint8_t *__restrict__ actions, float *__restrict__ q_out) {
    extern __shared__ __align__(16) float smem[];
    const RnnTile tile(smem);
    const int64_t base = (int64_t)blockIdx.x * L_TILE;
    const int rows = (int)((n - base) < L_TILE ? (n - base) : L_TILE);
    const int tid = threadIdx.x;
    if (tid < L_TILE) tile.reset[tid] = (tid < rows && reset_mask) ? reset_mask[base + tid] : (uint8_t)0;
    for (int w = tid; w < L_TILE * 7; w += L_THREADS) {
        const int env = w / 7, k = w - env * 7;
        tile.obs_s[k * L_TILE + env] = env < rows ? obs[base * 7 + w] : 0.0f;
    }
    __syncthreads();
    float q[3] = {0.f, 0.f, 0.f};
    rnn_forward(tile, pol.weights, pol.h, pol.c, n, base, rows, q);
    if (tid < rows) {
        int a = argmax3(q);
        a = explore(a, pol.eps_threshold, seed, (uint32_t)(env_id_base + base + tid), step_index, stream_id);
        actions[base + tid] = (uint8_t)a;
        if (q_out) { q_out[(base + tid) * 3 + 0] = q[0]; q_out[(base + tid) * 3 + 1] = q[1]; q_out[(base + tid) * 3 + 2] = q[2]; }
    }
}

// k fused lock-step iterations of {obs, QNetRNN (or follower / random) A, B, step, replay row, auto-reset} for a tile of
// 64 envs per CTA: the match loop of tests/arena.py:294-304 and the rollout loop of scripts/train_rnn_iterative.py:
// 732-780 for recurrent players.  Env state lives in the registers of threads 0..63 for all k steps; (h, c) of both
// players stay in global memory (unit-major, L2-resident for the tile) and are zeroed at every episode start.
template <typename R>
__global__ void __launch_bounds__(L_THREADS, 1)
selfplay_rnn_kernel(const PPParams params, const PPEnvState st, int64_t n, int64_t k_steps, const PPPolicy pol_a,
                    const PPPolicy pol_b, uint64_t seed, int64_t step_base, const PPServeSource src, int32_t quota,
                    int64_t env_id_base, const PPRolloutOut out, const PPReplayRing ring) {
    extern __shared__ __align__(16) float smem[];
    const RnnTile tile(smem);
    const int64_t base = (int64_t)blockIdx.x * L_TILE;
    const int rows = (int)((n - base) < L_TILE ? (n - base) : L_TILE);
    const int tid = threadIdx.x;
    const bool env_thread = tid < L_TILE;                       // warps 0 and 1, whole warps
    const int64_t i = base + tid;
    const bool valid = env_thread && tid < rows;
    const int64_t ic = valid ? i : 0;
    const EnvConsts<R> c(params);
    const StatePtrs<R> s(st);
    Lane<R> L;
    L.e = load_env<R>(s, ic);
    L.ep_idx = s.ep_idx[ic]; L.ep_len = s.ep_len[ic];
    const uint32_t gid = (uint32_t)(env_id_base + ic);
    const int64_t ring_t0 = ring_first_step(ring, n, k_steps);
    bool fresh = L.ep_len == 0;                                 // episode start: (h, c) = 0 before the first step

#pragma unroll 1
    for (int64_t t = 0; t < k_steps; ++t) {
        const bool active = valid && !(quota > 0 && L.ep_idx >= quota);
        if (!__syncthreads_or(active)) break;                   // the whole tile is frozen by the quota: for good
        const uint32_t step = (uint32_t)(step_base + t);
        float oa[7], ob[7];
        int act_a = 1, act_b = 1;
        if (env_thread) observe<R>(L.e, oa, ob);
#pragma unroll 1
        for (int p = 0; p < 2; ++p) {
            const PPPolicy &pol = p ? pol_b : pol_a;
            const uint32_t stream_id = p ? STREAM_ACT_B : STREAM_ACT_A;
            int a = 1;
            if (pol.kind == PP_POLICY_QNETRNN) {                // uniform over the CTA
                if (env_thread) {
#pragma unroll
                    for (int k = 0; k < 7; ++k) tile.obs_s[k * L_TILE + tid] = valid ? (p ? ob[k] : oa[k]) : 0.0f;
                    tile.reset[tid] = fresh ? 1 : 0;
                }
                __syncthreads();
                float q[3] = {0.f, 0.f, 0.f};
                rnn_forward(tile, pol.weights, pol.h, pol.c, n, base, rows, q);
                if (valid) a = explore(argmax3(q), pol.eps_threshold, seed, gid, step, stream_id);   // (h, c) advance even when exploring
                __syncthreads();                                // obs_s / A are reused by the other player
            } else if (env_thread) {
                if (pol.kind == PP_POLICY_RANDOM) a = random_action(seed, gid, step, stream_id);
                else if (pol.kind == PP_POLICY_QNET) a = explore(qnet_greedy_global(pol.weights, p ? ob : oa), pol.eps_threshold, seed, gid, step, stream_id);
                else a = explore(follower_action(p ? ob : oa, pol.follower_tol), pol.eps_threshold, seed, gid, step, stream_id);
            }
            if (p) act_b = a; else act_a = a;
        }
        if (env_thread) {                                       // two whole warps: the bookkeeping collectives are safe
            const int ep_before = L.ep_idx;
            step_and_book<R>(c, L, active, act_a, act_b, ob, t, n, i, env_id_base, quota, out, ring,
                             ring.head != nullptr && t >= ring_t0, src,
                             [&](int ep, R &vx, R &vy, R &sp) { next_serve<R>(params, src, n, i, env_id_base, ep, vx, vy, sp); });
            fresh = active ? (L.ep_idx != ep_before) : fresh;   // a new episode starts with zero (h, c)
        }
    }
    if (valid) {
        store_env<R>(s, i, L.e);
        s.ep_idx[i] = L.ep_idx;
        s.ep_len[i] = L.ep_len;
    }
    if (env_thread) L.tally.flush(out.counters, out.ep_log ? nullptr : out.ep_log_count);
}

constexpr size_t L_SMEM = (size_t)(256 * L_TILE + 2 * L_CHUNK + L_F1 * L_TILE + 7 * L_TILE) * sizeof(float) + L_TILE;

int qnetrnn_act_launch(int64_t n, const float *obs, const PPPolicy &pol, const uint8_t *reset_mask, uint64_t seed,
                       int64_t step_index, int64_t env_id_base, int32_t stream_id, uint8_t *actions, float *q_out,
                       cudaStream_t stream) {
    cudaError_t err = cudaFuncSetAttribute(qnetrnn_act_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L_SMEM);
    if (err != cudaSuccess) return (int)err;
    const unsigned blocks = (unsigned)((n + L_TILE - 1) / L_TILE);
    qnetrnn_act_kernel<<<blocks, L_THREADS, L_SMEM, stream>>>(n, obs, pol, reset_mask, seed, (uint32_t)step_index,
                                                              env_id_base, (uint32_t)stream_id, actions, q_out);
    return (int)cudaGetLastError();
}

int selfplay_rnn_launch(int mode, int64_t n, int64_t k, const PPParams &p, const PPEnvState &st, const PPPolicy &pa,
                        const PPPolicy &pb, uint64_t seed, int64_t step_base, const PPServeSource &src, int32_t quota,
                        int64_t env_id_base, const PPRolloutOut &out, const PPReplayRing *ring, cudaStream_t stream) {
    PPReplayRing r{};
    if (ring) r = *ring;
    const unsigned blocks = (unsigned)((n + L_TILE - 1) / L_TILE);
    cudaError_t err;
    if (mode == PP_MODE_F64) {
        if ((err = cudaFuncSetAttribute(selfplay_rnn_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L_SMEM)) != cudaSuccess) return (int)err;
        selfplay_rnn_kernel<double><<<blocks, L_THREADS, L_SMEM, stream>>>(p, st, n, k, pa, pb, seed, step_base, src, quota, env_id_base, out, r);
    } else {
        if ((err = cudaFuncSetAttribute(selfplay_rnn_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L_SMEM)) != cudaSuccess) return (int)err;
        selfplay_rnn_kernel<float><<<blocks, L_THREADS, L_SMEM, stream>>>(p, st, n, k, pa, pb, seed, step_base, src, quota, env_id_base, out, r);
    }
    return (int)cudaGetLastError();
}

}  // namespace pp
