// dqn_kernels.cu — training mode of scripts/train_iterative.py on the device: the Double-DQN head update up to its
// gradients (train_step, :139-164), NoisyLinear.reset_noise (models/qnet.py:33-41) and the weight packing that feeds
// the rollout kernels — three small launches instead of ~130 framework kernels per update.
//
// Only the NoisyNet heads train (features are frozen, scripts/train_iterative.py:97): 520 parameters.  A batch row
// needs two feature forwards (s and s') of 4 544 MACs each and three head evaluations.  The batch is small, so the
// kernel is latency-bound: it is cut into tiles of 16 rows, one CTA each; SIXTEEN lanes share a row (each owns 4 of the
// 64 units of either layer), partial head sums meet through four shuffles, the activations of s are parked in shared
// memory and a column-per-thread pass over the tile's rows gives the CTA's share of the 4 x 64 weight gradients.
// Partials go to a workspace; the CTA that takes the last ticket adds them in tile order (deterministic), writes the
// eight gradient tensors, the loss and the new priorities (last occurrence of a slot in the batch wins, like the
// reference's Python loop).  fp32 throughout (tolerance vs torch autograd: 1e-5).
#include "pp_host.h"
#include "pp_policy.cuh"

namespace pp {

namespace {

constexpr int D_THREADS = 256;
constexpr int D_ROWS = 16;                              // batch rows per CTA: sixteen lanes per row, 4 units each
constexpr int D_MAX_BATCH = 4096;                       // the last CTA stages all slots of the batch in shared memory
constexpr int D_H2_STRIDE = 65;                         // row stride of the h2 tile: conflict-free column walks
constexpr int D_FEAT_FLOATS = PP_QNET_WHT;              // W1T, B1, W2T, B2 of the packed blob (4 672 floats)
constexpr int D_PART = 264;                             // per-CTA partials: 256 weight grads, 4 bias grads, loss, pad

// h2 = relu(W2 relu(W1 x + b1) + b2) for one row shared by SIXTEEN lanes (half a warp): lane part `pt` computes units
// pt*4..+3 of both layers; the row's h1 goes through shared memory (h1row, 64 floats, private to the row).  Rolled
// loops, short dependent chains: the batch is small, so the kernel is latency- not throughput-bound.
__device__ __forceinline__ float4 qnet_features4(const float *__restrict__ sw, const float (&obs)[7], int pt,
                                                 float *__restrict__ h1row) {
    const float4 *w1 = reinterpret_cast<const float4 *>(sw + PP_QNET_W1T) + pt;
    float4 acc = reinterpret_cast<const float4 *>(sw + PP_QNET_B1)[pt];
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        const float4 w = w1[k * 16];
        acc.x = fmaf(w.x, obs[k], acc.x); acc.y = fmaf(w.y, obs[k], acc.y);
        acc.z = fmaf(w.z, obs[k], acc.z); acc.w = fmaf(w.w, obs[k], acc.w);
    }
    __syncwarp();                                                              // the row's previous h1 has been consumed
    h1row[pt * 4 + 0] = relu(acc.x); h1row[pt * 4 + 1] = relu(acc.y);
    h1row[pt * 4 + 2] = relu(acc.z); h1row[pt * 4 + 3] = relu(acc.w);
    __syncwarp();                                                              // the sixteen lanes of a row sit in one warp
    const float4 *w2 = reinterpret_cast<const float4 *>(sw + PP_QNET_W2T) + pt;
    float4 a = reinterpret_cast<const float4 *>(sw + PP_QNET_B2)[pt];
#pragma unroll 8
    for (int k = 0; k < 64; ++k) {
        const float x = h1row[k];
        const float4 u = w2[k * 16];
        a.x = fmaf(u.x, x, a.x); a.y = fmaf(u.y, x, a.y); a.z = fmaf(u.z, x, a.z); a.w = fmaf(u.w, x, a.w);
    }
    return make_float4(relu(a.x), relu(a.y), relu(a.z), relu(a.w));
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

__device__ __forceinline__ float4 row_sum(float4 v) {                          // over the sixteen lanes of a row
#pragma unroll
    for (int o = 1; o <= 8; o <<= 1) {
        v.x += __shfl_xor_sync(0xffffffffu, v.x, o); v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
        v.z += __shfl_xor_sync(0xffffffffu, v.z, o); v.w += __shfl_xor_sync(0xffffffffu, v.w, o);
    }
    return v;
}

__device__ __forceinline__ float4 f4_add(const float4 &a, const float4 &b) {
    return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}

// Double-DQN on the heads: forward, TD error, loss and the gradients of the eight head tensors.
// ws: [tiles][D_PART] partials, then [batch] TD errors, then the ticket counter (zero between launches).
__global__ void __launch_bounds__(D_THREADS, 1)
dqn_head_grads_kernel(const PPReplayRing ring, const int64_t *__restrict__ idx, const float *__restrict__ iw, int batch,
                      const float *__restrict__ w1, const float *__restrict__ b1, const float *__restrict__ w2,
                      const float *__restrict__ b2, const PPNoisyLayer on_v, const PPNoisyLayer on_a,
                      const PPNoisyLayer tg_v, const PPNoisyLayer tg_a, int noisy_online, int noisy_target, float gamma,
                      float *__restrict__ td_out, float *__restrict__ loss_out, float *__restrict__ prios,
                      float *__restrict__ max_prio, float *ws) {
    extern __shared__ __align__(16) float smem[];
    float *sw = smem;                                                          // feature weights
    float4 *head_on = reinterpret_cast<float4 *>(smem + D_FEAT_FLOATS);        // [65]
    float4 *head_tg = head_on + 65;                                            // [65]
    float *h2s = reinterpret_cast<float *>(head_tg + 65);                      // [D_ROWS][65] hidden activations of s
    float *g_s = h2s + D_ROWS * D_H2_STRIDE;                                   // [D_ROWS] dL/dQ(s, a)
    int *a_s = reinterpret_cast<int *>(g_s + D_ROWS);                          // [D_ROWS] action taken
    float *scratch = reinterpret_cast<float *>(a_s + D_ROWS);                  // [8]
    float *h1s = scratch + 8;                                                  // [D_ROWS][65] first-layer activations per row
    __shared__ int is_last;
    const int tid = threadIdx.x, tiles = gridDim.x;
    float *ws_td = ws + (size_t)tiles * D_PART;
    unsigned int *ticket = reinterpret_cast<unsigned int *>(ws_td + batch);
    // feature layers straight from the module's tensors ([out][in]) into the k-major tables of the fmaf chain
    // (consecutive lanes write consecutive words; the strided global reads of a warp share their sectors in L1)
    for (int o = tid; o < 64 * 7; o += D_THREADS) sw[PP_QNET_W1T + o] = w1[(o & 63) * 7 + (o >> 6)];
    for (int o = tid; o < 64 * 64; o += D_THREADS) sw[PP_QNET_W2T + o] = w2[(o & 63) * 64 + (o >> 6)];
    if (tid < 64) { sw[PP_QNET_B1 + tid] = b1[tid]; sw[PP_QNET_B2 + tid] = b2[tid]; }
    stage_head(head_on, on_v, on_a, noisy_online != 0);
    stage_head(head_tg, tg_v, tg_a, noisy_target != 0);
    __syncthreads();

    const int row = tid >> 4, pt = tid & 15;                                   // sixteen consecutive lanes share a row
    const int r = blockIdx.x * D_ROWS + row;
    const float inv_b = 1.0f / (float)batch;
    float loss_part = 0.f;
    {
        const bool live = r < batch;
        const int64_t slot = live ? idx[r] : 0;
        float s[7], ns[7];
#pragma unroll
        for (int k = 0; k < 7; ++k) { s[k] = ring.obs[slot * 7 + k]; ns[k] = ring.next_obs[slot * 7 + k]; }
        float *h1row = h1s + row * D_H2_STRIDE;
        const float4 h2 = qnet_features4(sw, s, pt, h1row);
        const float h2v[4] = {h2.x, h2.y, h2.z, h2.w};
        float4 hs = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            h2s[row * D_H2_STRIDE + pt * 4 + i] = live ? h2v[i] : 0.f;
            head_acc(hs, head_on[pt * 4 + i], h2v[i]);
        }
        const float4 n2 = qnet_features4(sw, ns, pt, h1row);
        const float n2v[4] = {n2.x, n2.y, n2.z, n2.w};
        float4 hn_on = make_float4(0.f, 0.f, 0.f, 0.f), hn_tg = hn_on;
#pragma unroll
        for (int i = 0; i < 4; ++i) { head_acc(hn_on, head_on[pt * 4 + i], n2v[i]); head_acc(hn_tg, head_tg[pt * 4 + i], n2v[i]); }
        hs = f4_add(row_sum(hs), head_on[64]);
        hn_on = f4_add(row_sum(hn_on), head_on[64]);
        hn_tg = f4_add(row_sum(hn_tg), head_tg[64]);
        if (pt == 0) {
            float g = 0.f;
            int act = 0;
            if (live) {
                float q[3], qn_on[3], qn_tg[3];
                dueling(hs, q); dueling(hn_on, qn_on); dueling(hn_tg, qn_tg);
                act = ring.act[slot];
                act = act > 2 ? 2 : act;
                const int best = argmax3(qn_on);                               // :154
                const float nq = best == 0 ? qn_tg[0] : (best == 1 ? qn_tg[1] : qn_tg[2]);
                const float qa = act == 0 ? q[0] : (act == 1 ? q[1] : q[2]);
                const float alive = ring.done[slot] ? 0.0f : 1.0f;
                const float td = qa - (ring.rew[slot] + gamma * nq * alive);   // :152-156
                const float w = iw[r];
                loss_part = w * td * td;                                       // :158
                g = 2.0f * w * td * inv_b;
                ws_td[r] = td;
                if (td_out) td_out[r] = td;
            }
            g_s[row] = g; a_s[row] = act;
        }
    }
    __syncthreads();
    // this tile's share of the gradients:  d/dV = g ;  d/dA_j = g * ([j == a] - 1/3)      (Q_a = V + A_a - mean A)
    const int hj = tid >> 6, hk = tid & 63;                                    // gradient column: head row, hidden unit
    float grad_w = 0.f, grad_b = 0.f;
#pragma unroll 4
    for (int rr = 0; rr < D_ROWS; ++rr) {
        const float gr = g_s[rr];
        const float coef = hj == 0 ? gr : gr * ((a_s[rr] == hj - 1 ? 1.0f : 0.0f) - (1.0f / 3.0f));
        grad_w = fmaf(coef, h2s[rr * D_H2_STRIDE + hk], grad_w);
        grad_b += coef;
    }
    const float loss_tile = block_sum(loss_part, scratch);
    float *part = ws + (size_t)blockIdx.x * D_PART;
    part[tid] = grad_w;
    if (hk == 0) part[256 + hj] = grad_b;
    if (tid == 0) part[260] = loss_tile;
    __threadfence();
    __syncthreads();
    if (tid == 0) is_last = atomicAdd(ticket, 1u) == (unsigned)(tiles - 1);
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // ---- the last CTA: partials in tile order, then the outputs
    float gw = 0.f, gb = 0.f, loss = 0.f;
    for (int c = 0; c < tiles; ++c) {
        const float *pc = ws + (size_t)c * D_PART;
        gw += __ldcg(pc + tid);
        gb += __ldcg(pc + 256 + hj);
        loss += __ldcg(pc + 260);
    }
    if (tid == 0 && loss_out) *loss_out = loss * inv_b;
    const PPNoisyLayer &L = hj == 0 ? on_v : on_a;                             // weight = mu + sigma * eps
    const int wi = hj == 0 ? hk : (hj - 1) * 64 + hk, bi = hj == 0 ? 0 : hj - 1;
    const bool noisy = noisy_online != 0;
    if (L.grad_weight_mu) L.grad_weight_mu[wi] = gw;
    if (L.grad_weight_sigma) L.grad_weight_sigma[wi] = noisy ? gw * L.weight_epsilon[wi] : 0.f;
    if (hk == 0) {
        if (L.grad_bias_mu) L.grad_bias_mu[bi] = gb;
        if (L.grad_bias_sigma) L.grad_bias_sigma[bi] = noisy ? gb * L.bias_epsilon[bi] : 0.f;
    }
    if (prios) {               // `for idx, err in zip(idxs, errors)` :74-76 — the LAST occurrence of a slot in the batch wins
        long long *slots = reinterpret_cast<long long *>(smem);                // the weights are no longer needed
        __syncthreads();
        for (int rr = tid; rr < batch; rr += D_THREADS) slots[rr] = idx[rr];
        __syncthreads();
        for (int rr = tid; rr < batch; rr += D_THREADS) {
            const long long slot = slots[rr];
            bool last = true;
            for (int q = rr + 1; q < batch; ++q) last = last && slots[q] != slot;
            if (last) {
                const float pnew = fabsf(__ldcg(ws_td + rr)) + 1e-6f;
                prios[slot] = pnew;
                // running maximum of every priority ever assigned (positive floats order like their bit patterns): what
                // new rows get (:57,62) without a pass over the whole priority array per push
                if (max_prio) atomicMax(reinterpret_cast<int *>(max_prio), __float_as_int(pnew));
            }
        }
    }
    if (tid == 0) *ticket = 0u;                                                // ready for the next launch
}

constexpr size_t D_SMEM_USED = (size_t)(D_FEAT_FLOATS + 2 * 65 * 4 + 2 * D_ROWS * D_H2_STRIDE + 2 * D_ROWS + 8) * sizeof(float);
constexpr size_t D_SMEM = D_SMEM_USED > D_MAX_BATCH * sizeof(long long) ? D_SMEM_USED : D_MAX_BATCH * sizeof(long long);

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

// torch.optim.Adam (no amsgrad, no weight decay) for a handful of small tensors in one launch, on the optimiser's own
// state tensors (exp_avg, exp_avg_sq, step), so that optimizer.state_dict() keeps working for checkpoints.
struct AdamParams { PPAdamParam p[16]; };

__global__ void __launch_bounds__(256, 1)
adam_step_kernel(const AdamParams ps, int count, double lr, double beta1, double beta2, double eps) {
    __shared__ float s_step_size[16], s_bc2_sqrt[16], s_step[16];
    const float w1 = (float)(1.0 - beta1), b2 = (float)beta2, w2 = (float)(1.0 - beta2), epsf = (float)eps;
    if ((int)threadIdx.x < count) {            // one thread per tensor: the scalars in double, like the Python floats of
        const double step = (double)*ps.p[threadIdx.x].step + 1.0;      // torch's Adam (1 - 0.999^step cancels in fp32)
        s_step[threadIdx.x] = (float)step;
        s_step_size[threadIdx.x] = (float)(lr / (1.0 - pow(beta1, step)));
        s_bc2_sqrt[threadIdx.x] = (float)sqrt(1.0 - pow(beta2, step));
    }
    __syncthreads();
    int64_t total = 0;
    for (int t = 0; t < count; ++t) total += ps.p[t].numel;
    for (int64_t f = threadIdx.x; f < total; f += blockDim.x) {               // all tensors as one flat range: the small
        int t = 0;                                                             // ones do not serialise behind each other
        int64_t i = f;
        while (i >= ps.p[t].numel) { i -= ps.p[t].numel; ++t; }
        const PPAdamParam &a = ps.p[t];
        const float g = a.grad[i];
        const float m = a.exp_avg[i] + w1 * (g - a.exp_avg[i]);                                 // lerp_(grad, 1 - beta1)
        const float v = a.exp_avg_sq[i] * b2 + w2 * (g * g);                                    // mul_(beta2).addcmul_
        a.exp_avg[i] = m; a.exp_avg_sq[i] = v;
        a.param[i] = a.param[i] - s_step_size[t] * (m / (sqrtf(v) / s_bc2_sqrt[t] + epsf));    // addcdiv_
    }
    if ((int)threadIdx.x < count) *ps.p[threadIdx.x].step = s_step[threadIdx.x];
}

// ---------------------------------------------------------------------------------------------
// The same Adam step with the gradient ALL-REDUCE over NVLink fused in front of it (several ranks, one process per GPU):
// no NCCL call, no second launch.  Every rank owns one block of peer-mapped ("symmetric") memory
//     float staging[2][capacity]   its gradients of the current / previous epoch
//     uint32 flags[8]              flags[r] = last epoch rank r has published
// and sees all blocks through its own address space (PPPeerBlocks.blocks[r]).  Per update (epoch e = *epoch + 1):
//   1. copy my flat gradient buffer into my staging[e & 1], fence (system scope), then write e into flags[me] of EVERY
//      rank's block (remote stores over NVLink: a waiting rank polls its own memory);
//   2. wait until my flags[r] >= e for all r;
//   3. read every rank's staging[e & 1] (peer loads), add them in rank order — the same order on every rank, so all
//      replicas compute bit-identical averages —, write the mean back into the flat gradient buffer;
//   4. Adam on the averaged gradients.
// Double buffering is enough: a rank overwrites staging[e & 1] at epoch e + 2, i.e. after it has passed the wait of epoch
// e + 1, which no rank can satisfy before it has finished its own epoch-e kernel.
struct PeerBlocks { float *blocks[8]; int rank, world; long long capacity; };

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__global__ void __launch_bounds__(256, 1)
adam_allreduce_kernel(const AdamParams ps, int count, float *__restrict__ flat_grad, int numel, const PeerBlocks peers,
                      unsigned long long *epoch, double lr, double beta1, double beta2, double eps) {
    __shared__ float s_step_size[16], s_bc2_sqrt[16], s_step[16];
    const int tid = threadIdx.x;
    const unsigned e = (unsigned)(*epoch + 1ull);
    const int slot = (int)(e & 1u);
    float *mine = peers.blocks[peers.rank] + (long long)slot * peers.capacity;
    for (int f = tid; f < numel; f += blockDim.x) mine[f] = flat_grad[f];
    __threadfence_system();
    __syncthreads();
    if (tid < peers.world) {
        unsigned *flags = reinterpret_cast<unsigned *>(peers.blocks[tid] + 2 * peers.capacity);
        st_release_sys(flags + peers.rank, e);
    }
    if (tid < peers.world) {
        const unsigned *my_flags = reinterpret_cast<const unsigned *>(peers.blocks[peers.rank] + 2 * peers.capacity);
        while ((int)(ld_acquire_sys(my_flags + tid) - e) < 0) {}
    }
    __syncthreads();
    const float inv = 1.0f / (float)peers.world;
    for (int f = tid; f < numel; f += blockDim.x) {
        float acc = 0.0f;
        for (int r = 0; r < peers.world; ++r)
            acc += *reinterpret_cast<const volatile float *>(peers.blocks[r] + (long long)slot * peers.capacity + f);
        flat_grad[f] = acc * inv;
    }
    __syncthreads();
    // ---- adam_step_kernel's body on the averaged gradients
    const float w1 = (float)(1.0 - beta1), b2 = (float)beta2, w2 = (float)(1.0 - beta2), epsf = (float)eps;
    if (tid < count) {
        const double step = (double)*ps.p[tid].step + 1.0;
        s_step[tid] = (float)step;
        s_step_size[tid] = (float)(lr / (1.0 - pow(beta1, step)));
        s_bc2_sqrt[tid] = (float)sqrt(1.0 - pow(beta2, step));
    }
    __syncthreads();
    int64_t total = 0;
    for (int t = 0; t < count; ++t) total += ps.p[t].numel;
    for (int64_t f = tid; f < total; f += blockDim.x) {
        int t = 0;
        int64_t i = f;
        while (i >= ps.p[t].numel) { i -= ps.p[t].numel; ++t; }
        const PPAdamParam &a = ps.p[t];
        const float g = a.grad[i];
        const float m = a.exp_avg[i] + w1 * (g - a.exp_avg[i]);
        const float v = a.exp_avg_sq[i] * b2 + w2 * (g * g);
        a.exp_avg[i] = m; a.exp_avg_sq[i] = v;
        a.param[i] = a.param[i] - s_step_size[t] * (m / (sqrtf(v) / s_bc2_sqrt[t] + epsf));
    }
    if (tid < count) *ps.p[tid].step = s_step[tid];
    if (tid == 0) *epoch = (unsigned long long)e;
}

// ---------------------------------------------------------------------------------------------
// PrioritizedReplay.sample (scripts/train_iterative.py:64-73): np.random.choice(len, batch, p = prios^alpha / sum) and
// the importance weights (N p)^-beta / max, as a two-level inverse-CDF draw:
//   per_chunk_sums_kernel : sum of prios^alpha over chunks of the priority array (one pass, fixed summation order)
//   per_draw_kernel       : every CTA scans the chunk sums in shared memory (double); one WARP per sample picks the chunk
//                           by binary search and walks it with a warp scan; writes the slot and (N p)^-beta
//   per_normalise_kernel  : divides by the batch maximum
// Deterministic for a given (seed, counter), unlike a device-wide floating-point cumsum.
constexpr int S_THREADS = 256;
constexpr int S_MAX_CHUNKS = 4096;                      // chunk sums every draw CTA keeps in shared memory (as double)

__global__ void __launch_bounds__(S_THREADS)
per_chunk_sums_kernel(const float *__restrict__ prios, int64_t capacity, int64_t chunk, float alpha, float *__restrict__ sums) {
    __shared__ float scratch[S_THREADS / 32];
    const int64_t lo = (int64_t)blockIdx.x * chunk, hi = lo + chunk < capacity ? lo + chunk : capacity;
    float acc = 0.f;
    if (hi - lo == chunk && (reinterpret_cast<uintptr_t>(prios + lo) & 15u) == 0) {       // whole chunk: 16-byte loads
        const float4 *p4 = reinterpret_cast<const float4 *>(prios + lo);
        for (int64_t i = threadIdx.x; i < chunk / 4; i += S_THREADS) {
            const float4 p = p4[i];
            acc += (p.x > 0.f ? __powf(p.x, alpha) : 0.f) + (p.y > 0.f ? __powf(p.y, alpha) : 0.f) +
                   (p.z > 0.f ? __powf(p.z, alpha) : 0.f) + (p.w > 0.f ? __powf(p.w, alpha) : 0.f);
        }
    } else {
        for (int64_t i = lo + threadIdx.x; i < hi; i += S_THREADS) {
            const float p = prios[i];
            acc += p > 0.f ? __powf(p, alpha) : 0.f;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < S_THREADS / 32; ++w) t += scratch[w];
        sums[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(S_THREADS)
per_draw_kernel(const float *__restrict__ prios, int64_t capacity, int64_t chunk, int n_chunks, float alpha,
                const float *__restrict__ sums, const float *__restrict__ beta, const float *__restrict__ size,
                uint64_t seed, const unsigned long long *__restrict__ counter, int batch, int64_t *__restrict__ idx_out,
                float *__restrict__ w_out) {
    extern __shared__ double prefix[];                     // inclusive prefix of the chunk sums, one pad per 32 entries:
    auto px = [](int i) { return i + (i >> 5); };          // a lane's segment of the scan starts in its own bank pair
    for (int i = threadIdx.x; i < n_chunks; i += S_THREADS) prefix[px(i)] = (double)sums[i];
    __syncthreads();
    if (threadIdx.x < 32) {                                                    // one warp scans: n_chunks <= 4096
        const int lane = threadIdx.x, per = (n_chunks + 31) / 32;
        double run = 0.0;
        for (int j = 0; j < per; ++j) { const int i = lane * per + j; if (i < n_chunks) { run += prefix[px(i)]; prefix[px(i)] = run; } }
        double off = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const double v = __shfl_up_sync(0xffffffffu, off, o); if (lane >= o) off += v; }
        off -= run;                                                            // exclusive offset of this lane's segment
        for (int j = 0; j < per; ++j) { const int i = lane * per + j; if (i < n_chunks) prefix[px(i)] += off; }
    }
    __syncthreads();
    const double total = prefix[px(n_chunks - 1)];
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * (S_THREADS / 32) + (threadIdx.x >> 5);          // one warp per sample
    if (r >= batch) return;
    const unsigned long long ctr = *counter;
    const uint4 rnd = philox4x32_10((uint32_t)r, 0x70657273u, (uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)seed, (uint32_t)(seed >> 32));
    const double u = fmin(u53(rnd.x, rnd.y) * total, total * (1.0 - 1e-15));   // strictly inside the last live chunk
    int lo = 0, hi = n_chunks - 1;                                             // first chunk with prefix > u
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (prefix[px(mid)] > u) hi = mid; else lo = mid + 1; }
    const float rem = (float)(u - (lo > 0 ? prefix[px(lo - 1)] : 0.0));
    const int64_t c_lo = (int64_t)lo * chunk, c_hi = c_lo + chunk < capacity ? c_lo + chunk : capacity;
    float run = 0.f, pa_sel = 0.f;
    int64_t pick = -1, last_pos = -1;
    float last_pa = 0.f;
    // 512 slots per step: every lane takes 16 consecutive ones (four independent 16-byte loads in flight), the lane sums
    // go through a warp scan, the lane that crosses `rem` resolves the slot among its 16
    for (int64_t base = c_lo; base < c_hi && pick < 0; base += 512) {
        const int64_t l0 = base + lane * 16;
        float pa[16];
        if (l0 + 16 <= c_hi && (reinterpret_cast<uintptr_t>(prios + l0) & 15u) == 0) {
            const float4 *p4 = reinterpret_cast<const float4 *>(prios + l0);
            const float4 q0 = p4[0], q1 = p4[1], q2 = p4[2], q3 = p4[3];
            const float raw[16] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, q3.x, q3.y, q3.z, q3.w};
#pragma unroll
            for (int e = 0; e < 16; ++e) pa[e] = raw[e] > 0.f ? __powf(raw[e], alpha) : 0.f;
        } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const float p = l0 + e < c_hi ? prios[l0 + e] : 0.f;
                pa[e] = p > 0.f ? __powf(p, alpha) : 0.f;
            }
        }
        float lane_sum = 0.f;
        int last_e = -1;
#pragma unroll
        for (int e = 0; e < 16; ++e) { lane_sum += pa[e]; if (pa[e] > 0.f) last_e = e; }
        float inc = lane_sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const float v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
        // this lane's candidate: the first of its slots whose running sum passes rem (its last live slot if rounding hides it)
        float cum = run + (inc - lane_sum);
        int cand = -1;
#pragma unroll
        for (int e = 0; e < 16; ++e) { cum += pa[e]; if (cand < 0 && pa[e] > 0.f && cum > rem) cand = e; }
        if (cand < 0) cand = last_e;
        float cand_pa = 0.f, last_lane_pa = 0.f;
#pragma unroll
        for (int e = 0; e < 16; ++e) { if (e == cand) cand_pa = pa[e]; if (e == last_e) last_lane_pa = pa[e]; }
        const unsigned hit = __ballot_sync(0xffffffffu, lane_sum > 0.f && run + inc > rem);
        const unsigned pos = __ballot_sync(0xffffffffu, lane_sum > 0.f);
        if (pos) {
            const int l = 31 - __clz(pos);
            last_pos = base + l * 16 + __shfl_sync(0xffffffffu, last_e, l);
            last_pa = __shfl_sync(0xffffffffu, last_lane_pa, l);
        }
        if (hit) {
            const int l = __ffs(hit) - 1;
            pick = base + l * 16 + __shfl_sync(0xffffffffu, cand, l);
            pa_sel = __shfl_sync(0xffffffffu, cand_pa, l);
        }
        run += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (pick < 0) { pick = last_pos; pa_sel = last_pa; }                       // rounding at the chunk's end: its last live slot
    if (pick < 0) { pick = 0; pa_sel = 0.f; }                                  // (an all-zero chunk cannot be chosen: its sum is 0)
    if (lane == 0) {
        idx_out[r] = pick;
        const float prob = (float)((double)pa_sel / total);
        w_out[r] = powf(*size * prob, -*beta);                                 // (N * P(i))^-beta  :71
    }
}

__global__ void __launch_bounds__(S_THREADS)
per_normalise_kernel(float *__restrict__ w, int batch, unsigned long long *counter) {
    __shared__ float scratch[S_THREADS / 32];
    float m = 0.f;
    for (int i = threadIdx.x; i < batch; i += S_THREADS) m = fmaxf(m, w[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = m;
    __syncthreads();
    m = 0.f;
    for (int k = 0; k < S_THREADS / 32; ++k) m = fmaxf(m, scratch[k]);
    for (int i = threadIdx.x; i < batch; i += S_THREADS) w[i] = w[i] / m;      // weights /= weights.max()  :72
    if (threadIdx.x == 0) *counter = *counter + 1;                             // the next launch draws a fresh batch
}

}  // namespace

int64_t per_chunk(int64_t capacity) {                  // chunk length: a multiple of 1024, at most S_MAX_CHUNKS chunks
    int64_t chunk = 4096;
    while ((capacity + chunk - 1) / chunk > S_MAX_CHUNKS) chunk *= 2;
    return chunk;
}

int per_sample_launch(const float *prios, int64_t capacity, float alpha, const float *beta, const float *size, uint64_t seed,
                      unsigned long long *counter, int32_t batch, float *chunk_sums, int64_t *idx_out, float *w_out,
                      cudaStream_t stream) {
    const int64_t chunk = per_chunk(capacity);
    const int n_chunks = (int)((capacity + chunk - 1) / chunk);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t err = cudaFuncSetAttribute(per_draw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((S_MAX_CHUNKS + S_MAX_CHUNKS / 32 + 1) * sizeof(double)));
        if (err != cudaSuccess) return (int)err;
        attr_set = true;
    }
    per_chunk_sums_kernel<<<n_chunks, S_THREADS, 0, stream>>>(prios, capacity, chunk, alpha, chunk_sums);
    per_draw_kernel<<<(batch + S_THREADS / 32 - 1) / (S_THREADS / 32), S_THREADS, (size_t)(n_chunks + n_chunks / 32 + 1) * sizeof(double), stream>>>(
        prios, capacity, chunk, n_chunks, alpha, chunk_sums, beta, size, seed, counter, batch, idx_out, w_out);
    per_normalise_kernel<<<1, S_THREADS, 0, stream>>>(w_out, batch, counter);
    return (int)cudaGetLastError();
}

int adam_step_launch(const PPAdamParam *params, int32_t count, double lr, double beta1, double beta2, double eps, cudaStream_t stream) {
    AdamParams pack{};
    for (int i = 0; i < count; ++i) pack.p[i] = params[i];
    adam_step_kernel<<<1, 256, 0, stream>>>(pack, count, lr, beta1, beta2, eps);
    return (int)cudaGetLastError();
}

int adam_allreduce_launch(const PPAdamParam *params, int32_t count, float *flat_grad, int64_t numel, const PPPeerBlocks &peers,
                          unsigned long long *epoch, double lr, double beta1, double beta2, double eps, cudaStream_t stream) {
    AdamParams pack{};
    for (int i = 0; i < count; ++i) pack.p[i] = params[i];
    PeerBlocks pb{};
    for (int r = 0; r < peers.world; ++r) pb.blocks[r] = reinterpret_cast<float *>(peers.blocks[r]);
    pb.rank = peers.rank; pb.world = peers.world; pb.capacity = peers.capacity_floats;
    adam_allreduce_kernel<<<1, 256, 0, stream>>>(pack, count, flat_grad, (int)numel, pb, epoch, lr, beta1, beta2, eps);
    return (int)cudaGetLastError();
}

int64_t dqn_workspace_floats(int32_t batch) { return (int64_t)((batch + D_ROWS - 1) / D_ROWS) * D_PART + batch + 8; }

int dqn_head_grads_launch(const PPReplayRing &ring, const int64_t *idx, const float *iw, int32_t batch,
                          const float *w1, const float *b1, const float *w2, const float *b2, const PPNoisyLayer &on_v,
                          const PPNoisyLayer &on_a, const PPNoisyLayer &tg_v, const PPNoisyLayer &tg_a, int noisy_online,
                          int noisy_target, float gamma, float *td_out, float *loss_out, float *prios, float *max_prio,
                          float *workspace, cudaStream_t stream) {
    static bool attr_set = false;                      // set once, before any stream capture replays the launch
    if (!attr_set) {
        cudaError_t err = cudaFuncSetAttribute(dqn_head_grads_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)D_SMEM);
        if (err != cudaSuccess) return (int)err;
        attr_set = true;
    }
    const unsigned tiles = (unsigned)((batch + D_ROWS - 1) / D_ROWS);
    dqn_head_grads_kernel<<<tiles, D_THREADS, D_SMEM, stream>>>(ring, idx, iw, batch, w1, b1, w2, b2, on_v, on_a, tg_v, tg_a,
                                                                noisy_online, noisy_target, gamma, td_out, loss_out, prios,
                                                                max_prio, workspace);
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
