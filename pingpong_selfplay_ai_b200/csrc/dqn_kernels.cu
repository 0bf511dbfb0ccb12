// dqn_kernels.cu — training mode of scripts/train_iterative.py on the device: the Double-DQN head update up to its
// gradients (train_step, :139-164), NoisyLinear.reset_noise (models/qnet.py:33-41) and the weight packing that feeds
// the rollout kernels — three small launches instead of ~130 framework kernels per update.
//
// Only the NoisyNet heads train (features are frozen, scripts/train_iterative.py:97): 520 parameters.  A batch row
// needs two feature forwards (s and s') of 4 544 MACs each and three head evaluations; one CTA of 256 threads does a
// batch of 256 rows in ~6 us: one row per thread through the same broadcast-LDS fmaf chain as the CUDA-core rollout,
// the 64 hidden activations of s parked in shared memory, then a column-per-thread reduction over the rows for the
// 4 x 64 weight gradients.  fp32 throughout (tolerance vs torch autograd: 1e-5).
#include "pp_host.h"
#include "pp_policy.cuh"

namespace pp {

namespace {

constexpr int D_THREADS = 256;
constexpr int D_H2_STRIDE = 65;                         // row stride of the h2 tile: conflict-free column walks
constexpr int D_FEAT_FLOATS = PP_QNET_WHT;              // W1T, B1, W2T, B2 of the packed blob (4 672 floats)

// features(x) of QNet: h2 = relu(W2 relu(W1 x + b1) + b2); consume(k, h2[k]) is called for k = 0..63 ascending.
template <typename F>
__device__ __forceinline__ void qnet_features(const float *__restrict__ sw, const float (&obs)[7], F &&consume) {
    float h1[64];
    {
        const float4 *w1 = reinterpret_cast<const float4 *>(sw + PP_QNET_W1T);
        const float4 *b1 = reinterpret_cast<const float4 *>(sw + PP_QNET_B1);
#pragma unroll
        for (int j4 = 0; j4 < 16; ++j4) {
            float4 acc = b1[j4];
#pragma unroll
            for (int k = 0; k < 7; ++k) {
                const float4 w = w1[k * 16 + j4];
                acc.x = fmaf(w.x, obs[k], acc.x); acc.y = fmaf(w.y, obs[k], acc.y);
                acc.z = fmaf(w.z, obs[k], acc.z); acc.w = fmaf(w.w, obs[k], acc.w);
            }
            h1[j4 * 4 + 0] = relu(acc.x); h1[j4 * 4 + 1] = relu(acc.y);
            h1[j4 * 4 + 2] = relu(acc.z); h1[j4 * 4 + 3] = relu(acc.w);
        }
    }
    const float4 *w2 = reinterpret_cast<const float4 *>(sw + PP_QNET_W2T);
    const float4 *b2 = reinterpret_cast<const float4 *>(sw + PP_QNET_B2);
#pragma unroll 1
    for (int jb = 0; jb < 4; ++jb) {
        float4 a0 = b2[jb * 4 + 0], a1 = b2[jb * 4 + 1], a2 = b2[jb * 4 + 2], a3 = b2[jb * 4 + 3];
#pragma unroll
        for (int k = 0; k < 64; ++k) {
            const float x = h1[k];
            const float4 u0 = w2[k * 16 + jb * 4 + 0], u1 = w2[k * 16 + jb * 4 + 1];
            const float4 u2 = w2[k * 16 + jb * 4 + 2], u3 = w2[k * 16 + jb * 4 + 3];
            a0.x = fmaf(u0.x, x, a0.x); a0.y = fmaf(u0.y, x, a0.y); a0.z = fmaf(u0.z, x, a0.z); a0.w = fmaf(u0.w, x, a0.w);
            a1.x = fmaf(u1.x, x, a1.x); a1.y = fmaf(u1.y, x, a1.y); a1.z = fmaf(u1.z, x, a1.z); a1.w = fmaf(u1.w, x, a1.w);
            a2.x = fmaf(u2.x, x, a2.x); a2.y = fmaf(u2.y, x, a2.y); a2.z = fmaf(u2.z, x, a2.z); a2.w = fmaf(u2.w, x, a2.w);
            a3.x = fmaf(u3.x, x, a3.x); a3.y = fmaf(u3.y, x, a3.y); a3.z = fmaf(u3.z, x, a3.z); a3.w = fmaf(u3.w, x, a3.w);
        }
        const float h2[16] = {relu(a0.x), relu(a0.y), relu(a0.z), relu(a0.w), relu(a1.x), relu(a1.y), relu(a1.z), relu(a1.w),
                              relu(a2.x), relu(a2.y), relu(a2.z), relu(a2.w), relu(a3.x), relu(a3.y), relu(a3.z), relu(a3.w)};
#pragma unroll
        for (int i = 0; i < 16; ++i) consume(jb * 16 + i, h2[i]);
    }
}

__device__ __forceinline__ float eff(const float *mu, const float *sigma, const float *eps, int i, bool noisy) {
    return noisy ? __fadd_rn(mu[i], __fmul_rn(sigma[i], eps[i])) : mu[i];      // models/qnet.py:44-46, torch's two roundings
}

// (V, A0, A1, A2) head as a k-major float4 table [64] + bias float4, from the two NoisyLinear layers
__device__ void stage_head(float4 *tab, const PPNoisyLayer &v, const PPNoisyLayer &a, bool noisy) {
    for (int k = threadIdx.x; k < 64; k += blockDim.x)
        tab[k] = make_float4(eff(v.weight_mu, v.weight_sigma, v.weight_epsilon, k, noisy),
                             eff(a.weight_mu, a.weight_sigma, a.weight_epsilon, k, noisy),
                             eff(a.weight_mu, a.weight_sigma, a.weight_epsilon, 64 + k, noisy),
                             eff(a.weight_mu, a.weight_sigma, a.weight_epsilon, 128 + k, noisy));
    if (threadIdx.x == 0)
        tab[64] = make_float4(eff(v.bias_mu, v.bias_sigma, v.bias_epsilon, 0, noisy), eff(a.bias_mu, a.bias_sigma, a.bias_epsilon, 0, noisy),
                              eff(a.bias_mu, a.bias_sigma, a.bias_epsilon, 1, noisy), eff(a.bias_mu, a.bias_sigma, a.bias_epsilon, 2, noisy));
}

__device__ __forceinline__ void dueling(const float4 &h, float (&q)[3]) {     // V + (A - mean(A))   models/qnet.py:75
    const float mean = __fdiv_rn(__fadd_rn(__fadd_rn(h.y, h.z), h.w), 3.0f);
    q[0] = __fadd_rn(h.x, __fsub_rn(h.y, mean));
    q[1] = __fadd_rn(h.x, __fsub_rn(h.z, mean));
    q[2] = __fadd_rn(h.x, __fsub_rn(h.w, mean));
}

__device__ __forceinline__ void head_acc(float4 &acc, const float4 &w, float x) {
    acc.x = fmaf(w.x, x, acc.x); acc.y = fmaf(w.y, x, acc.y); acc.z = fmaf(w.z, x, acc.z); acc.w = fmaf(w.w, x, acc.w);
}

__device__ __forceinline__ float block_sum(float v, float *scratch) {         // all threads get the sum
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    float s = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += scratch[w];
    return s;
}

// Double-DQN on the heads: forward, TD error, loss and the gradients of the eight head tensors.
__global__ void __launch_bounds__(D_THREADS, 1)
dqn_head_grads_kernel(const PPReplayRing ring, const int64_t *__restrict__ idx, const float *__restrict__ iw, int batch,
                      const float *__restrict__ w1, const float *__restrict__ b1, const float *__restrict__ w2,
                      const float *__restrict__ b2, const PPNoisyLayer on_v, const PPNoisyLayer on_a,
                      const PPNoisyLayer tg_v, const PPNoisyLayer tg_a, int noisy_online, int noisy_target, float gamma,
                      float *__restrict__ td_out, float *__restrict__ loss_out, float *__restrict__ prios) {
    extern __shared__ __align__(16) float smem[];
    float *sw = smem;                                                          // feature weights
    float4 *head_on = reinterpret_cast<float4 *>(smem + D_FEAT_FLOATS);        // [65]
    float4 *head_tg = head_on + 65;                                            // [65]
    float *h2s = reinterpret_cast<float *>(head_tg + 65);                      // [256][65]
    float *g_s = h2s + D_THREADS * D_H2_STRIDE;                                // [256] dL/dQ(s, a)
    int *a_s = reinterpret_cast<int *>(g_s + D_THREADS);                       // [256] action taken
    float *scratch = reinterpret_cast<float *>(a_s + D_THREADS);               // [8]
    long long *slot_s = reinterpret_cast<long long *>(scratch + 8);            // [256] ring slot of each row of the tile
    const int tid = threadIdx.x;
    // feature layers straight from the module's tensors ([out][in]) into the k-major tables of the fmaf chain
    for (int i = tid; i < 64 * 7; i += D_THREADS) sw[PP_QNET_W1T + (i % 7) * 64 + i / 7] = w1[i];
    for (int i = tid; i < 64 * 64; i += D_THREADS) sw[PP_QNET_W2T + (i & 63) * 64 + (i >> 6)] = w2[i];
    if (tid < 64) { sw[PP_QNET_B1 + tid] = b1[tid]; sw[PP_QNET_B2 + tid] = b2[tid]; }
    stage_head(head_on, on_v, on_a, noisy_online != 0);
    stage_head(head_tg, tg_v, tg_a, noisy_target != 0);
    __syncthreads();

    const int hj = tid >> 6, hk = tid & 63;                                    // this thread's gradient column: head row, unit
    float grad_w = 0.f, grad_b = 0.f, loss_part = 0.f;
    const float inv_b = 1.0f / (float)batch;
    for (int base = 0; base < batch; base += D_THREADS) {
        const int r = base + tid;
        float g = 0.f, prio = 0.f;
        int act = 0;
        int64_t slot = -1;
        if (r < batch) {
            slot = idx[r];
            float s[7], ns[7];
#pragma unroll
            for (int k = 0; k < 7; ++k) { s[k] = ring.obs[slot * 7 + k]; ns[k] = ring.next_obs[slot * 7 + k]; }
            act = ring.act[slot];
            act = act > 2 ? 2 : act;
            float4 hs = head_on[64];
            float *row = h2s + tid * D_H2_STRIDE;
            qnet_features(sw, s, [&](int k, float v) { row[k] = v; head_acc(hs, head_on[k], v); });
            float4 hn_on = head_on[64], hn_tg = head_tg[64];
            qnet_features(sw, ns, [&](int k, float v) { head_acc(hn_on, head_on[k], v); head_acc(hn_tg, head_tg[k], v); });
            float q[3], qn_on[3], qn_tg[3];
            dueling(hs, q); dueling(hn_on, qn_on); dueling(hn_tg, qn_tg);
            const int best = argmax3(qn_on);                                   // :154
            const float alive = ring.done[slot] ? 0.0f : 1.0f;
            const float target = ring.rew[slot] + gamma * qn_tg[best] * alive; // :155-156
            const float td = q[act] - target;
            const float w = iw[r];
            loss_part += w * td * td;                                          // :158
            g = 2.0f * w * td * inv_b;
            if (td_out) td_out[r] = td;
            prio = fabsf(td) + 1e-6f;                                          // :163-164, :74-76
        } else {
            float *row = h2s + tid * D_H2_STRIDE;
#pragma unroll 1
            for (int k = 0; k < 64; ++k) row[k] = 0.f;
        }
        g_s[tid] = g; a_s[tid] = act; slot_s[tid] = slot;
        __syncthreads();
        if (prios && slot >= 0) {              // `for idx, err in zip(...)`: the LAST occurrence of a slot in the batch wins
            bool last = true;
            for (int rr = tid + 1; rr < D_THREADS; ++rr) last = last && slot_s[rr] != slot;
            if (last) prios[slot] = prio;      // (a later tile overwrites an earlier one: tiles run in batch order)
        }
        // d/dV = g ; d/dA_j = g * ([j == a] - 1/3)        (Q_a = V + A_a - mean A)
#pragma unroll 4
        for (int rr = 0; rr < D_THREADS; ++rr) {
            const float gr = g_s[rr];
            const float coef = hj == 0 ? gr : gr * ((a_s[rr] == hj - 1 ? 1.0f : 0.0f) - (1.0f / 3.0f));
            grad_w = fmaf(coef, h2s[rr * D_H2_STRIDE + hk], grad_w);
            if (hk == 0) grad_b += coef;
        }
        __syncthreads();
    }
    const float loss = block_sum(loss_part, scratch) * inv_b;
    if (tid == 0 && loss_out) *loss_out = loss;
    // gradients of mu and sigma (weight = mu + sigma * eps)
    const PPNoisyLayer &L = hj == 0 ? on_v : on_a;
    const int wi = hj == 0 ? hk : (hj - 1) * 64 + hk, bi = hj == 0 ? 0 : hj - 1;
    const bool noisy = noisy_online != 0;
    if (L.grad_weight_mu) L.grad_weight_mu[wi] = grad_w;
    if (L.grad_weight_sigma) L.grad_weight_sigma[wi] = noisy ? grad_w * L.weight_epsilon[wi] : 0.f;
    if (hk == 0) {
        if (L.grad_bias_mu) L.grad_bias_mu[bi] = grad_b;
        if (L.grad_bias_sigma) L.grad_bias_sigma[bi] = noisy ? grad_b * L.bias_epsilon[bi] : 0.f;
    }
}

constexpr size_t D_SMEM = (size_t)(D_FEAT_FLOATS + 2 * 65 * 4 + D_THREADS * D_H2_STRIDE + 2 * D_THREADS + 8) * sizeof(float) +
                          D_THREADS * sizeof(long long);

// NoisyLinear.reset_noise for up to 8 layers: factorised Gaussian noise from Philox + Box-Muller.
struct NoisyLayers { PPNoisyLayer l[8]; };

__device__ __forceinline__ float signed_sqrt_normal(uint32_t a, uint32_t b) {  // f(g) = sign(g) sqrt|g|, g ~ N(0, 1)
    const float u1 = ((float)(a >> 8) + 0.5f) * (1.0f / 16777216.0f);          // (0, 1)
    const float u2 = ((float)(b >> 8) + 0.5f) * (1.0f / 16777216.0f);
    const float g = sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
    return copysignf(sqrtf(fabsf(g)), g);
}

__global__ void __launch_bounds__(256, 1)
noisy_reset_kernel(const NoisyLayers layers, int count, uint64_t seed, unsigned long long *counter) {
    __shared__ float e[1024];
    const unsigned long long ctr = *counter;
    for (int li = 0; li < count; ++li) {
        const PPNoisyLayer &L = layers.l[li];
        const int nin = L.in_features, nout = L.out_features;
        __syncthreads();
        for (int i = threadIdx.x; i < nin + nout; i += blockDim.x) {           // e[0, nin) = e_in, e[nin, nin + nout) = e_out
            const uint4 r = philox4x32_10((uint32_t)i, (uint32_t)li, (uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)seed,
                                          (uint32_t)(seed >> 32) ^ 0x6e6f6973u);
            e[i] = signed_sqrt_normal(r.x, r.y);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < nin * nout; i += blockDim.x) L.weight_epsilon[i] = e[nin + i / nin] * e[i % nin];
        for (int i = threadIdx.x; i < nout; i += blockDim.x) L.bias_epsilon[i] = e[nin + i];
    }
    __syncthreads();
    if (threadIdx.x == 0) *counter = ctr + 1;
}

// policy.pack_qnet as one launch: torch-layout tensors -> the k-major PP_QNET_* blob
__global__ void __launch_bounds__(256, 1)
pack_qnet_kernel(const float *__restrict__ w1, const float *__restrict__ b1, const float *__restrict__ w2,
                 const float *__restrict__ b2, const PPNoisyLayer v, const PPNoisyLayer a, int noisy, float *__restrict__ blob) {
    const bool nz = noisy != 0;
    for (int i = threadIdx.x; i < PP_QNET_BLOB_FLOATS; i += blockDim.x) {
        float out;
        if (i < PP_QNET_B1) { const int k = i / 64, j = i % 64; out = w1[j * 7 + k]; }
        else if (i < PP_QNET_W2T) out = b1[i - PP_QNET_B1];
        else if (i < PP_QNET_B2) { const int k = (i - PP_QNET_W2T) / 64, j = (i - PP_QNET_W2T) % 64; out = w2[j * 64 + k]; }
        else if (i < PP_QNET_WHT) out = b2[i - PP_QNET_B2];
        else if (i < PP_QNET_BH) {
            const int k = (i - PP_QNET_WHT) / 4, c = (i - PP_QNET_WHT) % 4;
            out = c == 0 ? eff(v.weight_mu, v.weight_sigma, v.weight_epsilon, k, nz)
                         : eff(a.weight_mu, a.weight_sigma, a.weight_epsilon, (c - 1) * 64 + k, nz);
        } else {
            const int c = i - PP_QNET_BH;
            out = c == 0 ? eff(v.bias_mu, v.bias_sigma, v.bias_epsilon, 0, nz) : eff(a.bias_mu, a.bias_sigma, a.bias_epsilon, c - 1, nz);
        }
        blob[i] = out;
    }
}

}  // namespace

int dqn_head_grads_launch(const PPReplayRing &ring, const int64_t *idx, const float *iw, int32_t batch,
                          const float *w1, const float *b1, const float *w2, const float *b2, const PPNoisyLayer &on_v, const PPNoisyLayer &on_a,
                          const PPNoisyLayer &tg_v, const PPNoisyLayer &tg_a, int noisy_online, int noisy_target, float gamma,
                          float *td_out, float *loss_out, float *prios, cudaStream_t stream) {
    static bool attr_set = false;                      // set once, before any stream capture replays the launch
    if (!attr_set) {
        cudaError_t err = cudaFuncSetAttribute(dqn_head_grads_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)D_SMEM);
        if (err != cudaSuccess) return (int)err;
        attr_set = true;
    }
    dqn_head_grads_kernel<<<1, D_THREADS, D_SMEM, stream>>>(ring, idx, iw, batch, w1, b1, w2, b2, on_v, on_a, tg_v, tg_a,
                                                            noisy_online, noisy_target, gamma, td_out, loss_out, prios);
    return (int)cudaGetLastError();
}

int noisy_reset_launch(const PPNoisyLayer *layers, int32_t count, uint64_t seed, unsigned long long *counter, cudaStream_t stream) {
    NoisyLayers pack{};
    for (int i = 0; i < count; ++i) pack.l[i] = layers[i];
    noisy_reset_kernel<<<1, 256, 0, stream>>>(pack, count, seed, counter);
    return (int)cudaGetLastError();
}

int pack_qnet_launch(const float *w1, const float *b1, const float *w2, const float *b2, const PPNoisyLayer &v,
                     const PPNoisyLayer &a, int noisy, float *blob, cudaStream_t stream) {
    pack_qnet_kernel<<<1, 256, 0, stream>>>(w1, b1, w2, b2, v, a, noisy, blob);
    return (int)cudaGetLastError();
}

}  // namespace pp
