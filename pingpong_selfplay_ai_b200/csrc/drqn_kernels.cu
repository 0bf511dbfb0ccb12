// drqn_kernels.cu — the DRQN update of scripts/train_rnn_iterative.py:400-531 (train_step_rnn), hand-written:
// forward of QNetRNN (models/qnet_rnn.py:107-144) over [batch, trace, 7] windows for three streams — online(obs) with
// everything backward needs saved, online(next_obs) and target(next_obs) — last-step Double-DQN Huber loss, backward
// through the heads, the LSTM (BPTT over the trace) and the feature layers, gradient-norm clipping and Adam.
//
// Everything is fp32 on the CUDA cores (the batch is 64 windows x 8 steps: ~0.8 GFLOP per update, latency- not
// math-bound; fp32 keeps the gradients within 1e-5 of torch autograd).  One update is ~20 launches in a CUDA graph:
//   drqn_prep_kernel       effective NoisyNet head weights (mu + sigma * eps) of the online / target net
//   drqn_features_kernel   gather windows from the replay ring, features.0 + ReLU, features.2 + ReLU       (3 streams)
//   sgemm_kernel           Gx = F2 W_ih^T + b_ih + b_hh for all time steps at once                          (3 streams)
//   lstm_fwd_kernel        the recurrence: one CLUSTER of 8 CTAs per (stream, 16 windows); each CTA owns 16 hidden units
//                          (its 64 gate columns of W_hh stay in shared memory for all steps) and h_t is exchanged
//                          through DISTRIBUTED SHARED MEMORY, one cluster barrier per time step
//   sgemm_kernel           shared head S = ReLU(h_T Ws^T + bs)                                               (3 streams)
//   drqn_loss_kernel       dueling Q, Double-DQN target, Huber loss, dS and the gradients of fc_V / fc_A
//   sgemm_kernel x2        dWs = dS^T h_T ;  dh_T = dS Ws
//   lstm_bwd_kernel        BPTT with the same cluster layout: gate derivatives, dW_hh accumulated in registers over the
//                          trace, dh_{t-1} = dGates_t W_hh reduced across the cluster through distributed shared memory
//   sgemm_kernel x5        dW_ih = dGx^T F2 ; dF2 = dGx W_ih (ReLU mask) ; dW2 = dF2^T F1 ; dF1 = dF2 W2 (mask) ; dW1 = dF1^T X
//   drqn_finalize_kernel   bias gradients (column sums), dW_hh summed over the window tiles, sigma gradients (x eps)
//   grad_sqnorm_kernel + grad_scale_kernel   torch.nn.utils.clip_grad_norm_ on the flat gradient buffer
//   adam_multi_kernel + adam_bump_kernel     torch.optim.Adam on the optimiser's own state tensors
#include <cooperative_groups.h>
#include <cuda_fp16.h>

#include "pp_host.h"

namespace cg = cooperative_groups;

namespace pp {

namespace {

constexpr int HID = 128, GATES = 512, FEAT = 128, F1D = 64, OBS = 7, OBSP = 8, SH = 128;
constexpr int CL = 8;                 // CTAs per cluster
constexpr int UPC = HID / CL;         // hidden units per CTA: 16
constexpr int CPC = 4 * UPC;          // gate columns per CTA: 64 (local column c = gate * 16 + unit)
constexpr int BT = 16;                // windows (batch rows) per cluster
constexpr int LSTM_THREADS = 256;
constexpr int MAX_TRACE = 16;

__device__ __forceinline__ float sigmoidf_(float v) { return 1.0f / (1.0f + expf(-v)); }

// ------------------------------------------------------------------------------------------ workspace layout (floats)
struct WsMap {
    int64_t x0, f1, f2, gx, gact, cs, hs, z, s, ds, dz, dgx, df2, df1, dwhh_part, eff, misc, total;
    __host__ __device__ WsMap(int B, int L) {
        const int64_t R = (int64_t)B * L;
        int64_t o = 0;
        auto take = [&](int64_t n) { int64_t at = o; o += (n + 3) / 4 * 4; return at; };
        x0 = take(R * OBSP); f1 = take(R * F1D); f2 = take(3 * R * FEAT); gx = take(3 * R * GATES);
        gact = take((int64_t)L * B * GATES); cs = take((int64_t)L * B * HID); hs = take((int64_t)L * B * HID);
        z = take(3 * (int64_t)B * HID); s = take(3 * (int64_t)B * SH); ds = take((int64_t)B * SH); dz = take((int64_t)B * HID);
        dgx = take(R * GATES); df2 = take(R * FEAT); df1 = take(R * F1D);
        dwhh_part = take((int64_t)(B / BT) * GATES * HID);
        eff = take(2 * (SH * HID + SH + 4 * SH + 4));           // per net: Ws [128][128], bs [128], Wh [4][128] (V, A0..2), bh [4]
        misc = take(64);
        total = o;
    }
};
constexpr int64_t EFF_NET = SH * HID + SH + 4 * SH + 4;
constexpr int64_t EFF_BS = SH * HID, EFF_WH = EFF_BS + SH, EFF_BH = EFF_WH + 4 * SH;

// ------------------------------------------------------------------------------------------ effective head weights
__device__ __forceinline__ float eff(const float *mu, const float *sigma, const float *eps, int64_t i, int noisy) {
    return noisy ? mu[i] + sigma[i] * eps[i] : mu[i];                        // models/qnet_rnn.py:44-46
}
__global__ void drqn_prep_kernel(const PPQNetRNNParams on, const PPQNetRNNParams tg, int noisy_on, int noisy_tg, float *effw) {
    for (int net = 0; net < 2; ++net) {
        const PPQNetRNNParams &p = net ? tg : on;
        const int noisy = net ? noisy_tg : noisy_on;
        float *e = effw + net * EFF_NET;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < EFF_NET; i += gridDim.x * blockDim.x) {
            float v;
            if (i < EFF_BS) v = eff(p.shared.weight_mu, p.shared.weight_sigma, p.shared.weight_epsilon, i, noisy);
            else if (i < EFF_WH) v = eff(p.shared.bias_mu, p.shared.bias_sigma, p.shared.bias_epsilon, i - EFF_BS, noisy);
            else if (i < EFF_BH) {
                const int j = (int)(i - EFF_WH), row = j / SH, k = j % SH;       // row 0 = V, 1..3 = A
                v = row == 0 ? eff(p.v.weight_mu, p.v.weight_sigma, p.v.weight_epsilon, k, noisy)
                             : eff(p.a.weight_mu, p.a.weight_sigma, p.a.weight_epsilon, (row - 1) * SH + k, noisy);
            } else {
                const int j = (int)(i - EFF_BH);
                v = j == 0 ? eff(p.v.bias_mu, p.v.bias_sigma, p.v.bias_epsilon, 0, noisy)
                           : eff(p.a.bias_mu, p.a.bias_sigma, p.a.bias_epsilon, j - 1, noisy);
            }
            e[i] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------ features
// stream 0 = online(obs), 1 = online(next_obs), 2 = target(next_obs).  16 window-steps per CTA, thread j = output unit j.
constexpr int FT_ROWS = 16;
__global__ void __launch_bounds__(FEAT)
drqn_features_kernel(const PPReplayRing ring, const int64_t *__restrict__ rows, int R, const PPQNetRNNParams on,
                     const PPQNetRNNParams tg, float *__restrict__ ws_x0, float *__restrict__ ws_f1, float *__restrict__ ws_f2) {
    __shared__ float w2t[F1D][FEAT];          // features.2 weight, k-major: conflict-free across j
    __shared__ float w1[F1D][OBSP];
    __shared__ float x[FT_ROWS][OBSP], f1[FT_ROWS][F1D];
    const int stream = blockIdx.y, j = threadIdx.x, r0 = blockIdx.x * FT_ROWS;
    const PPQNetRNNParams &p = stream == 2 ? tg : on;
    for (int i = j; i < FEAT * F1D; i += FEAT) w2t[i % F1D][i / F1D] = p.f2_w[i];
    for (int i = j; i < F1D * OBSP; i += FEAT) w1[i / OBSP][i % OBSP] = (i % OBSP) < OBS ? p.f0_w[(i / OBSP) * OBS + (i % OBSP)] : 0.0f;
    const float *src = stream == 0 ? ring.obs : ring.next_obs;
    for (int i = j; i < FT_ROWS * OBSP; i += FEAT) {
        const int r = r0 + i / OBSP, k = i % OBSP;
        x[i / OBSP][k] = (r < R && k < OBS) ? src[rows[r] * OBS + k] : 0.0f;
    }
    __syncthreads();
    if (stream == 0)
        for (int i = j; i < FT_ROWS * OBSP; i += FEAT)
            if (r0 + i / OBSP < R) ws_x0[(int64_t)(r0 + i / OBSP) * OBSP + i % OBSP] = x[i / OBSP][i % OBSP];
    if (j < F1D) {
        const float b = p.f0_b[j];
        for (int r = 0; r < FT_ROWS; ++r) {
            float acc = b;
#pragma unroll
            for (int k = 0; k < OBS; ++k) acc = fmaf(w1[j][k], x[r][k], acc);
            acc = acc > 0.0f ? acc : 0.0f;
            f1[r][j] = acc;
            if (stream == 0 && r0 + r < R) ws_f1[(int64_t)(r0 + r) * F1D + j] = acc;
        }
    }
    __syncthreads();
    const float b2 = p.f2_b[j];
    for (int r = 0; r < FT_ROWS; ++r) {
        float acc = b2;
#pragma unroll 16
        for (int k = 0; k < F1D; ++k) acc = fmaf(w2t[k][j], f1[r][k], acc);
        if (r0 + r < R) ws_f2[((int64_t)stream * R + r0 + r) * FEAT + j] = acc > 0.0f ? acc : 0.0f;
    }
}

// ------------------------------------------------------------------------------------------ generic small SGEMM
// C[m][n] = sum_k A(m, k) B(k, n)  (+ bias[n] + bias2[n]) (ReLU) (x [mask[m][n] > 0]); element strides make every
// transpose case one kernel.  Up to 3 independent problems per launch (blockIdx.z).  64 x 64 x 16 tiles, 4 x 4 per thread.
struct Gemm {
    const float *a, *b, *bias, *bias2, *mask;
    float *c;
    int m, n, k;
    int64_t sam, sak, sbk, sbn, ldc, ldmask;
    int relu;
};
struct GemmBatch { Gemm g[3]; };
constexpr int GM = 64, GN = 64, GK = 32, GLD = GM * GK / 256;     // 8 elements of A and of B per thread and K step

// The K loop is software-pipelined: the next K step's operands are fetched from global memory into registers while the
// current one is multiplied out of shared memory (two shared-memory buffers, one barrier per step).  The weight-gradient
// products reduce over all 512 window-steps with only a handful of output tiles: without the prefetch every step paid a
// full global-memory round trip and those three launches were 60 % of the update.
__global__ void __launch_bounds__(256)
sgemm_kernel(const GemmBatch batch) {
    const Gemm g = batch.g[blockIdx.z];
    const int m0 = blockIdx.y * GM, n0 = blockIdx.x * GN;
    if (m0 >= g.m || n0 >= g.n) return;
    __shared__ __align__(16) float as[2][GK][GM + 4], bs[2][GK][GN + 4];
    const int tid = threadIdx.x, ty = tid / 16, tx = tid % 16;
    float acc[4][4] = {};
    const bool a_k_fast = g.sak == 1, b_n_fast = g.sbn == 1;
    float ra[GLD], rb[GLD];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int i = 0; i < GLD; ++i) {
            const int idx = tid + i * 256;
            const int am = a_k_fast ? idx / GK : idx % GM, ak = a_k_fast ? idx % GK : idx / GM;
            ra[i] = (m0 + am < g.m && k0 + ak < g.k) ? g.a[(int64_t)(m0 + am) * g.sam + (int64_t)(k0 + ak) * g.sak] : 0.0f;
            const int bn = b_n_fast ? idx % GN : idx / GK, bk = b_n_fast ? idx / GN : idx % GK;
            rb[i] = (n0 + bn < g.n && k0 + bk < g.k) ? g.b[(int64_t)(k0 + bk) * g.sbk + (int64_t)(n0 + bn) * g.sbn] : 0.0f;
        }
    };
    auto stash = [&](int buf) {
#pragma unroll
        for (int i = 0; i < GLD; ++i) {
            const int idx = tid + i * 256;
            const int am = a_k_fast ? idx / GK : idx % GM, ak = a_k_fast ? idx % GK : idx / GM;
            as[buf][ak][am] = ra[i];
            const int bn = b_n_fast ? idx % GN : idx / GK, bk = b_n_fast ? idx / GN : idx % GK;
            bs[buf][bk][bn] = rb[i];
        }
    };
    fetch(0);
    stash(0);
    __syncthreads();
    int buf = 0;
    for (int k0 = 0; k0 < g.k; k0 += GK) {
        const bool more = k0 + GK < g.k;
        if (more) fetch(k0 + GK);
#pragma unroll
        for (int kk = 0; kk < GK; ++kk) {
            const float4 av = *reinterpret_cast<const float4 *>(&as[buf][kk][ty * 4]);
            const float4 bv = *reinterpret_cast<const float4 *>(&bs[buf][kk][tx * 4]);
            const float a4[4] = {av.x, av.y, av.z, av.w}, b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a4[i], b4[j], acc[i][j]);
        }
        if (more) stash(buf ^ 1);
        __syncthreads();
        buf ^= 1;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= g.m) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= g.n) continue;
            float v = acc[i][j];
            if (g.bias) v += g.bias[n];
            if (g.bias2) v += g.bias2[n];
            if (g.relu) v = v > 0.0f ? v : 0.0f;
            if (g.mask) v = g.mask[(int64_t)m * g.ldmask + n] > 0.0f ? v : 0.0f;
            g.c[(int64_t)m * g.ldc + n] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------ LSTM forward (cluster)
// grid (CL, B / BT, streams); cluster dims (CL, 1, 1).  CTA rank r owns hidden units [16 r, 16 r + 16): local gate column
// c = gate * 16 + u  <->  global column gate * 128 + 16 r + u (torch's i, f, g, o blocks of W_hh / W_ih).
struct LstmSmem {
    float wt[HID][CPC];                // W_hh slice, k-major:  wt[k][c] = W_hh[col(c)][k]                   32 KB
    float h[2][BT][HID];               // h_{t-1} / h_t of the cluster's 16 windows, all 128 units (replicated)  16 KB
    float gates[BT][CPC];              // pre-activations of this CTA's columns                                  4 KB
};

__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(LSTM_THREADS)
lstm_fwd_kernel(const PPQNetRNNParams on, const PPQNetRNNParams tg, int B, int L, const float *__restrict__ ws_gx,
                float *__restrict__ ws_gact, float *__restrict__ ws_cs, float *__restrict__ ws_hs, float *__restrict__ ws_z) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    LstmSmem &sm = *reinterpret_cast<LstmSmem *>(smem_raw);
    cg::cluster_group cluster = cg::this_cluster();
    const int r = (int)cluster.block_rank(), tile = blockIdx.y, stream = blockIdx.z, tid = threadIdx.x;
    const float *whh = (stream == 2 ? tg : on).w_hh;
    for (int i = tid; i < CPC * HID; i += LSTM_THREADS) {
        const int c = i / HID, k = i % HID;                                     // coalesced along k
        sm.wt[k][c] = whh[(int64_t)((c / UPC) * HID + r * UPC + (c % UPC)) * HID + k];
    }
    for (int i = tid; i < 2 * BT * HID; i += LSTM_THREADS) (&sm.h[0][0][0])[i] = 0.0f;      // zero initial (h, c): :432-433
    const int row = tid / 16, cb = (tid % 16) * 4;                               // GEMV part: 4 local columns of one window
    const int gate = cb / UPC, u0 = cb % UPC;
    const int u = tid % 16;                                                      // cell part: (window row, unit u)
    const int b = tile * BT + row;
    const int64_t R = (int64_t)B * L;
    float c_state = 0.0f;
    cluster.sync();
    for (int t = 0; t < L; ++t) {
        const float *hc = &sm.h[t & 1][row][0];
        float4 acc = *reinterpret_cast<const float4 *>(ws_gx + ((int64_t)stream * R + (int64_t)b * L + t) * GATES + gate * HID + r * UPC + u0);
#pragma unroll 8
        for (int k = 0; k < HID; ++k) {
            const float hv = hc[k];
            const float4 w = *reinterpret_cast<const float4 *>(&sm.wt[k][cb]);
            acc.x = fmaf(hv, w.x, acc.x); acc.y = fmaf(hv, w.y, acc.y); acc.z = fmaf(hv, w.z, acc.z); acc.w = fmaf(hv, w.w, acc.w);
        }
        *reinterpret_cast<float4 *>(&sm.gates[row][cb]) = acc;
        __syncthreads();
        const float ig = sigmoidf_(sm.gates[row][u]), fg = sigmoidf_(sm.gates[row][UPC + u]);
        const float gg = tanhf(sm.gates[row][2 * UPC + u]), og = sigmoidf_(sm.gates[row][3 * UPC + u]);
        c_state = fg * c_state + ig * gg;
        const float hn = og * tanhf(c_state);
        const int unit = r * UPC + u;
        if (stream == 0) {                                                       // what BPTT needs
            float *ga = ws_gact + ((int64_t)t * B + b) * GATES;
            ga[unit] = ig; ga[HID + unit] = fg; ga[2 * HID + unit] = gg; ga[3 * HID + unit] = og;
            ws_cs[((int64_t)t * B + b) * HID + unit] = c_state;
            ws_hs[((int64_t)t * B + b) * HID + unit] = hn;
        }
        if (t == L - 1) ws_z[((int64_t)stream * B + b) * HID + unit] = hn;
#pragma unroll
        for (int dst = 0; dst < CL; ++dst) {                                     // h_t to every CTA of the cluster (DSMEM)
            float *remote = cluster.map_shared_rank(&sm.h[(t + 1) & 1][0][0], dst);
            remote[row * HID + unit] = hn;
        }
        cluster.sync();                                                          // h_t complete everywhere; gates[] reusable
    }
}

// ------------------------------------------------------------------------------------------ loss + head gradients
// One CTA.  S[3][B][128] = ReLU(shared head) of the three streams; writes dS (of stream 0), the gradients of fc_V / fc_A,
// the loss and the TD errors.                                          scripts/train_rnn_iterative.py:468-509
__global__ void __launch_bounds__(256)
drqn_loss_kernel(const PPReplayRing ring, const int64_t *__restrict__ rows, int B, int L, float gamma, const float *__restrict__ effw,
                 const float *__restrict__ ws_s, float *__restrict__ ws_ds, const PPQNetRNNParams on, const PPQNetRNNGrads gr,
                 int noisy_on, float *__restrict__ loss_out, float *__restrict__ td_out) {
    __shared__ float q[3][256][3], dv[256], da[256][3], red[256];
    const int tid = threadIdx.x;
    for (int i = tid; i < 3 * B; i += 256) {                                     // dueling Q of every stream and window
        const int s = i / B, b = i % B;
        const float *e = effw + (s == 2 ? EFF_NET : 0);
        const float *sv = ws_s + ((int64_t)s * B + b) * SH;
        float h4[4] = {e[EFF_BH], e[EFF_BH + 1], e[EFF_BH + 2], e[EFF_BH + 3]};
        for (int k = 0; k < SH; ++k) {
            const float x = sv[k];
#pragma unroll
            for (int o = 0; o < 4; ++o) h4[o] = fmaf(e[EFF_WH + o * SH + k], x, h4[o]);
        }
        const float mean = (h4[1] + h4[2] + h4[3]) / 3.0f;                       // V + (A - mean(A))  models/qnet_rnn.py:142
#pragma unroll
        for (int o = 0; o < 3; ++o) q[s][b][o] = h4[0] + (h4[1 + o] - mean);
    }
    __syncthreads();
    float loss_b = 0.0f;
    if (tid < B) {
        const int64_t last = rows[(int64_t)tid * L + L - 1];
        const int a = ring.act[last];
        const float rew = ring.rew[last];
        const bool done = ring.done[last] != 0;
        int best = 0;                                                            // argmax of the ONLINE net on next_obs (:489-490)
        if (q[1][tid][1] > q[1][tid][best]) best = 1;
        if (q[1][tid][2] > q[1][tid][best]) best = 2;
        const float target = rew + gamma * q[2][tid][best] * (done ? 0.0f : 1.0f);     // :505
        const float td = q[0][tid][a < 3 ? a : 0] - target;
        const float ad = fabsf(td);
        loss_b = ad < 1.0f ? 0.5f * td * td : ad - 0.5f;                         // smooth_l1_loss, beta = 1 (:509)
        const float dq = (ad < 1.0f ? td : (td > 0.0f ? 1.0f : -1.0f)) / (float)B;
        dv[tid] = dq;
#pragma unroll
        for (int o = 0; o < 3; ++o) da[tid][o] = dq * ((o == a ? 1.0f : 0.0f) - 1.0f / 3.0f);
        if (td_out) td_out[tid] = td;
    }
    red[tid] = loss_b;
    __syncthreads();
    if (tid == 0) {
        float t = 0.0f;
        for (int i = 0; i < B; ++i) t += red[i];
        if (loss_out) *loss_out = t / (float)B;
    }
    const float *e = effw;                                                       // online head weights
    for (int i = tid; i < B * SH; i += 256) {                                    // dS = (dV wv + dA Wa) x [S > 0]
        const int b = i / SH, k = i % SH;
        const float sv = ws_s[(int64_t)b * SH + k];
        float g = dv[b] * e[EFF_WH + k];
#pragma unroll
        for (int o = 0; o < 3; ++o) g = fmaf(da[b][o], e[EFF_WH + (1 + o) * SH + k], g);
        ws_ds[(int64_t)b * SH + k] = sv > 0.0f ? g : 0.0f;
    }
    for (int i = tid; i < 4 * SH + 4; i += 256) {                                // gradients of fc_V / fc_A (weights, then biases)
        float g = 0.0f;
        const bool is_bias = i >= 4 * SH;
        const int o = is_bias ? i - 4 * SH : i / SH, k = is_bias ? 0 : i % SH;
        for (int b = 0; b < B; ++b) {
            const float d = o == 0 ? dv[b] : da[b][o - 1];
            g = fmaf(d, is_bias ? 1.0f : ws_s[(int64_t)b * SH + k], g);
        }
        const PPNoisyLayer &lay = o == 0 ? on.v : on.a;
        const PPNoisyLayer &glay = o == 0 ? gr.v : gr.a;
        const int64_t at = o == 0 ? k : (int64_t)(o - 1) * SH + k, bat = o == 0 ? 0 : o - 1;
        if (!is_bias) {
            if (glay.grad_weight_mu) glay.grad_weight_mu[at] = g;
            if (glay.grad_weight_sigma) glay.grad_weight_sigma[at] = noisy_on ? g * lay.weight_epsilon[at] : 0.0f;
        } else {
            if (glay.grad_bias_mu) glay.grad_bias_mu[bat] = g;
            if (glay.grad_bias_sigma) glay.grad_bias_sigma[bat] = noisy_on ? g * lay.bias_epsilon[bat] : 0.0f;
        }
    }
}

// ------------------------------------------------------------------------------------------ LSTM backward (cluster)
// grid (CL, B / BT); stream 0 only.  Per step t = L-1 .. 0: cell derivatives for the CTA's 16 units, dGx row written for
// the batched GEMMs, dW_hh slice accumulated in registers, partial dh_{t-1} = dA_t W_hh[slice] reduced over the cluster.
struct BwdSmem {
    float w[CPC][HID];                 // W_hh slice, column-major here:  w[c][k] = W_hh[col(c)][k]             32 KB
    float part[2][BT][HID];            // this CTA's partial dh_{t-1} (double-buffered across steps)          16 KB
    float da[BT][CPC];                 // gate pre-activation derivatives of this CTA's columns                 4 KB
    float hprev[BT][HID];              // h_{t-1} of the cluster's windows                                       8 KB
};

__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(LSTM_THREADS)
lstm_bwd_kernel(const PPQNetRNNParams on, int B, int L, const float *__restrict__ ws_gact, const float *__restrict__ ws_cs,
                const float *__restrict__ ws_hs, const float *__restrict__ ws_dz, float *__restrict__ ws_dgx,
                float *__restrict__ ws_dwhh_part) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BwdSmem &sm = *reinterpret_cast<BwdSmem *>(smem_raw);
    cg::cluster_group cluster = cg::this_cluster();
    const int r = (int)cluster.block_rank(), tile = blockIdx.y, tid = threadIdx.x;
    for (int i = tid; i < CPC * HID; i += LSTM_THREADS) {
        const int c = i / HID, k = i % HID;
        sm.w[c][k] = on.w_hh[(int64_t)((c / UPC) * HID + r * UPC + (c % UPC)) * HID + k];
    }
    const int row = tid / 16, u = tid % 16, unit = r * UPC + u, b = tile * BT + row;       // cell part
    // dW_hh part: column wc, 32 k's as eight float4 groups INTERLEAVED over the four threads of a column (k = 16 i + 4 wq ..
    // + 3): the four threads of a quad read 64 contiguous bytes of h_{t-1} (a 128-byte stride put them all on one bank)
    const int wc = tid / 4, wq = tid % 4;
    // partial-dh part: window row, k = 4 pj .. + 3 and 64 + 4 pj .. + 3: 16 threads read 256 contiguous bytes of a W_hh row
    const int pj = tid % 16;
    float dwhh[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) dwhh[i] = 0.0f;
    float dh = ws_dz[(int64_t)b * HID + unit], dc_carry = 0.0f;
    cluster.sync();
    for (int t = L - 1; t >= 0; --t) {
        const float *ga = ws_gact + ((int64_t)t * B + b) * GATES;
        const float ig = ga[unit], fg = ga[HID + unit], gg = ga[2 * HID + unit], og = ga[3 * HID + unit];
        const float ct = ws_cs[((int64_t)t * B + b) * HID + unit];
        const float cprev = t > 0 ? ws_cs[((int64_t)(t - 1) * B + b) * HID + unit] : 0.0f;
        const float tc = tanhf(ct);
        const float d_o = dh * tc;
        const float dc = dh * og * (1.0f - tc * tc) + dc_carry;
        const float a_i = dc * gg * ig * (1.0f - ig), a_f = dc * cprev * fg * (1.0f - fg);
        const float a_g = dc * ig * (1.0f - gg * gg), a_o = d_o * og * (1.0f - og);
        dc_carry = dc * fg;
        sm.da[row][u] = a_i; sm.da[row][UPC + u] = a_f; sm.da[row][2 * UPC + u] = a_g; sm.da[row][3 * UPC + u] = a_o;
        float *dg = ws_dgx + ((int64_t)b * L + t) * GATES;
        dg[unit] = a_i; dg[HID + unit] = a_f; dg[2 * HID + unit] = a_g; dg[3 * HID + unit] = a_o;
        for (int i = tid; i < BT * HID; i += LSTM_THREADS) {                      // h_{t-1} of the 16 windows (zero at t = 0)
            const int rr = i / HID, k = i % HID;
            sm.hprev[rr][k] = t > 0 ? ws_hs[((int64_t)(t - 1) * B + tile * BT + rr) * HID + k] : 0.0f;
        }
        __syncthreads();
        if (t > 0) {                                                             // dW_hh[col(wc)][k] += sum_rows dA[row][wc] h_{t-1}[row][k]
#pragma unroll 4
            for (int rr = 0; rr < BT; ++rr) {
                const float a = sm.da[rr][wc];
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const float4 h4 = *reinterpret_cast<const float4 *>(&sm.hprev[rr][4 * i + 4 * wq]);
                    dwhh[i] = fmaf(a, h4.x, dwhh[i]); dwhh[i + 1] = fmaf(a, h4.y, dwhh[i + 1]);
                    dwhh[i + 2] = fmaf(a, h4.z, dwhh[i + 2]); dwhh[i + 3] = fmaf(a, h4.w, dwhh[i + 3]);
                }
            }
            float p[8] = {};                                                     // partial dh_{t-1}[row][this thread's 8 k's] over this CTA's columns
#pragma unroll 4
            for (int c = 0; c < CPC; ++c) {
                const float a = sm.da[row][c];
                const float4 w0 = *reinterpret_cast<const float4 *>(&sm.w[c][4 * pj]), w1 = *reinterpret_cast<const float4 *>(&sm.w[c][64 + 4 * pj]);
                p[0] = fmaf(a, w0.x, p[0]); p[1] = fmaf(a, w0.y, p[1]); p[2] = fmaf(a, w0.z, p[2]); p[3] = fmaf(a, w0.w, p[3]);
                p[4] = fmaf(a, w1.x, p[4]); p[5] = fmaf(a, w1.y, p[5]); p[6] = fmaf(a, w1.z, p[6]); p[7] = fmaf(a, w1.w, p[7]);
            }
            float *mine = &sm.part[t & 1][row][4 * pj];
            *reinterpret_cast<float4 *>(mine) = make_float4(p[0], p[1], p[2], p[3]);
            *reinterpret_cast<float4 *>(mine + 64) = make_float4(p[4], p[5], p[6], p[7]);
            cluster.sync();                                                      // every CTA's partial of this step is in place
            float acc = 0.0f;
#pragma unroll
            for (int src = 0; src < CL; ++src) {                                 // fixed order: deterministic
                const float *remote = cluster.map_shared_rank(&sm.part[t & 1][0][0], src);
                acc += remote[row * HID + unit];
            }
            dh = acc;                                                            // only the last step's Q carries a loss: no direct term
        }
        __syncthreads();                                                         // da / hprev are rewritten by the next step
    }
    float *out = ws_dwhh_part + ((int64_t)tile * GATES + (wc / UPC) * HID + r * UPC + (wc % UPC)) * HID;
#pragma unroll
    for (int i = 0; i < 32; i += 4)
        *reinterpret_cast<float4 *>(out + 4 * i + 4 * wq) = make_float4(dwhh[i], dwhh[i + 1], dwhh[i + 2], dwhh[i + 3]);
    cluster.sync();                                                              // nobody leaves while peers may still read its partials
}

// ------------------------------------------------------------------------------------------ bias / sigma / dW_hh finalisation
__global__ void __launch_bounds__(256)
drqn_finalize_kernel(int B, int L, const float *__restrict__ ws_dgx, const float *__restrict__ ws_df2, const float *__restrict__ ws_df1,
                     const float *__restrict__ ws_ds, const float *__restrict__ ws_dwhh_part, const PPQNetRNNParams on,
                     const PPQNetRNNGrads gr, int noisy_on) {
    const int64_t R = (int64_t)B * L;
    const int tiles = B / BT;
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
    // column sums: b_ih = b_hh (512), features.2 bias (128), features.0 bias (64), shared-head bias (128).  A CTA takes 32
    // columns at a time: 8 row lanes per column walk the rows, then add up in a fixed order through shared memory.
    __shared__ float colpart[8][33];
    const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
    for (int c0 = blockIdx.x * 32; c0 < GATES + FEAT + F1D + SH; c0 += gridDim.x * 32) {
        const int i = c0 + cl;                                       // all four segments are multiples of 32 columns wide
        const float *src; int64_t ld; int64_t nrows; int j;
        if (i < GATES) { src = ws_dgx; ld = GATES; nrows = R; j = i; }
        else if (i < GATES + FEAT) { src = ws_df2; ld = FEAT; nrows = R; j = i - GATES; }
        else if (i < GATES + FEAT + F1D) { src = ws_df1; ld = F1D; nrows = R; j = i - GATES - FEAT; }
        else { src = ws_ds; ld = SH; nrows = B; j = i - GATES - FEAT - F1D; }
        float acc = 0.0f;
        for (int64_t rr = rl; rr < nrows; rr += 8) acc += src[rr * ld + j];
        colpart[rl][cl] = acc;
        __syncthreads();
        if (rl == 0) {
            float t = 0.0f;
#pragma unroll
            for (int q = 0; q < 8; ++q) t += colpart[q][cl];
            if (i < GATES) { if (gr.b_ih) gr.b_ih[j] = t; if (gr.b_hh) gr.b_hh[j] = t; }
            else if (i < GATES + FEAT) { if (gr.f2_b) gr.f2_b[j] = t; }
            else if (i < GATES + FEAT + F1D) { if (gr.f0_b) gr.f0_b[j] = t; }
            else {
                if (gr.shared.grad_bias_mu) gr.shared.grad_bias_mu[j] = t;
                if (gr.shared.grad_bias_sigma) gr.shared.grad_bias_sigma[j] = noisy_on ? t * on.shared.bias_epsilon[j] : 0.0f;
            }
        }
        __syncthreads();
    }
    for (int64_t i = gtid; i < (int64_t)GATES * HID; i += stride) {              // dW_hh: the window tiles in order
        float acc = 0.0f;
        for (int t = 0; t < tiles; ++t) acc += ws_dwhh_part[(int64_t)t * GATES * HID + i];
        if (gr.w_hh) gr.w_hh[i] = acc;
    }
    for (int64_t i = gtid; i < (int64_t)SH * HID; i += stride)                   // sigma gradient of the shared head: dW x eps
        if (gr.shared.grad_weight_sigma)
            gr.shared.grad_weight_sigma[i] = noisy_on ? gr.shared.grad_weight_mu[i] * on.shared.weight_epsilon[i] : 0.0f;
}

// ------------------------------------------------------------------------------------------ clip_grad_norm_ + Adam
constexpr int NORM_THREADS = 256, NORM_MAX_BLOCKS = 256;
// scratch: [0 .. NORM_MAX_BLOCKS) partial sums of squares, then the ticket (as float bits).  out[0] = total norm,
// out[1] = clip coefficient = min(1, max_norm / (norm + 1e-6))             torch.nn.utils.clip_grad_norm_
__global__ void __launch_bounds__(NORM_THREADS)
grad_sqnorm_kernel(const float *__restrict__ g, int64_t n, float max_norm, float *__restrict__ scratch, float *__restrict__ out) {
    __shared__ float warp_sums[NORM_THREADS / 32];
    __shared__ bool last;
    float acc = 0.0f;
    for (int64_t i = (int64_t)blockIdx.x * NORM_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * NORM_THREADS) acc = fmaf(g[i], g[i], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.0f;
        for (int w = 0; w < NORM_THREADS / 32; ++w) t += warp_sums[w];
        scratch[blockIdx.x] = t;
        __threadfence();
        unsigned *ticket = reinterpret_cast<unsigned *>(scratch + NORM_MAX_BLOCKS);
        last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {                                              // the last CTA adds the partials in order
        __threadfence();
        double t = 0.0;
        for (unsigned bidx = 0; bidx < gridDim.x; ++bidx) t += (double)reinterpret_cast<volatile float *>(scratch)[bidx];
        const float norm = (float)sqrt(t);
        const float coef = max_norm / (norm + 1e-6f);
        out[0] = norm;
        out[1] = coef < 1.0f ? coef : 1.0f;
        *reinterpret_cast<unsigned *>(scratch + NORM_MAX_BLOCKS) = 0u;          // ready for the next launch
    }
}
__global__ void grad_scale_kernel(float *__restrict__ g, int64_t n, const float *__restrict__ norm_out) {
    const float coef = norm_out[1];
    if (coef >= 1.0f) return;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) g[i] *= coef;
}

struct AdamMulti { PPAdamParam p[32]; };
// torch.optim.Adam over many tensors, many CTAs: every CTA derives the step scalars from the OLD step counters (which
// adam_bump_kernel advances afterwards, in stream order).
__global__ void __launch_bounds__(256)
adam_multi_kernel(const AdamMulti ps, int count, double lr, double beta1, double beta2, double eps) {
    __shared__ float s_step_size[32], s_bc2_sqrt[32];
    __shared__ int64_t s_start[33];
    const float w1 = (float)(1.0 - beta1), b2 = (float)beta2, w2 = (float)(1.0 - beta2), epsf = (float)eps;
    if ((int)threadIdx.x < count) {
        const double step = (double)*ps.p[threadIdx.x].step + 1.0;
        s_step_size[threadIdx.x] = (float)(lr / (1.0 - pow(beta1, step)));
        s_bc2_sqrt[threadIdx.x] = (float)sqrt(1.0 - pow(beta2, step));
    }
    if (threadIdx.x == 0) {
        int64_t at = 0;
        for (int t = 0; t < count; ++t) { s_start[t] = at; at += ps.p[t].numel; }
        s_start[count] = at;
    }
    __syncthreads();
    const int64_t total = s_start[count];
    for (int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; f < total; f += (int64_t)gridDim.x * blockDim.x) {
        int lo = 0, hi = count - 1;                                              // the tensor that holds flat index f
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (s_start[mid] <= f) lo = mid; else hi = mid - 1; }
        const PPAdamParam &a = ps.p[lo];
        const int64_t i = f - s_start[lo];
        const float g = a.grad[i];
        const float m = a.exp_avg[i] + w1 * (g - a.exp_avg[i]);
        const float v = a.exp_avg_sq[i] * b2 + w2 * (g * g);
        a.exp_avg[i] = m; a.exp_avg_sq[i] = v;
        a.param[i] = a.param[i] - s_step_size[lo] * (m / (sqrtf(v) / s_bc2_sqrt[lo] + epsf));
    }
}
__global__ void adam_bump_kernel(const AdamMulti ps, int count) {
    if ((int)threadIdx.x < count) *ps.p[threadIdx.x].step += 1.0f;
}

}  // namespace

// ------------------------------------------------------------------------------------------ launchers
int64_t drqn_workspace_floats(int32_t batch, int32_t trace) { return WsMap(batch, trace).total; }

static Gemm gemm(const float *a, int64_t sam, int64_t sak, const float *b, int64_t sbk, int64_t sbn, float *c, int64_t ldc,
                 int m, int n, int k, const float *bias = nullptr, const float *bias2 = nullptr, int relu = 0,
                 const float *mask = nullptr, int64_t ldmask = 0) {
    Gemm g{};
    g.a = a; g.b = b; g.c = c; g.bias = bias; g.bias2 = bias2; g.mask = mask;
    g.m = m; g.n = n; g.k = k; g.sam = sam; g.sak = sak; g.sbk = sbk; g.sbn = sbn; g.ldc = ldc; g.ldmask = ldmask; g.relu = relu;
    return g;
}
static void launch_gemms(const Gemm *gs, int count, cudaStream_t stream) {
    GemmBatch batch{};
    int mm = 0, nn = 0;
    for (int i = 0; i < count; ++i) {
        batch.g[i] = gs[i];
        mm = gs[i].m > mm ? gs[i].m : mm;
        nn = gs[i].n > nn ? gs[i].n : nn;
    }
    sgemm_kernel<<<dim3((nn + GN - 1) / GN, (mm + GM - 1) / GM, count), 256, 0, stream>>>(batch);
}

int drqn_grads_launch(const PPReplayRing &ring, const int64_t *rows, int32_t B, int32_t L, const PPQNetRNNParams &on,
                      const PPQNetRNNParams &tg, int noisy_on, int noisy_tg, float gamma, const PPQNetRNNGrads &gr,
                      float *loss_out, float *td_out, float *ws, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e1 = cudaFuncSetAttribute(lstm_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LstmSmem));
        cudaError_t e2 = cudaFuncSetAttribute(lstm_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BwdSmem));
        if (e1 != cudaSuccess) return (int)e1;
        if (e2 != cudaSuccess) return (int)e2;
        attr_set = true;
    }
    const WsMap m(B, L);
    const int R = B * L;
    float *effw = ws + m.eff;
    drqn_prep_kernel<<<32, 256, 0, stream>>>(on, tg, noisy_on, noisy_tg, effw);
    drqn_features_kernel<<<dim3((R + FT_ROWS - 1) / FT_ROWS, 3), FEAT, 0, stream>>>(ring, rows, R, on, tg, ws + m.x0, ws + m.f1, ws + m.f2);
    {   // Gx[s] = F2[s] W_ih(net)^T + b_ih + b_hh                                 models/qnet_rnn.py:130 (the input half of the LSTM)
        Gemm gs[3];
        for (int s = 0; s < 3; ++s) {
            const PPQNetRNNParams &p = s == 2 ? tg : on;
            gs[s] = gemm(ws + m.f2 + (int64_t)s * R * FEAT, FEAT, 1, p.w_ih, 1, FEAT, ws + m.gx + (int64_t)s * R * GATES, GATES, R, GATES, FEAT, p.b_ih, p.b_hh);
        }
        launch_gemms(gs, 3, stream);
    }
    lstm_fwd_kernel<<<dim3(CL, B / BT, 3), LSTM_THREADS, sizeof(LstmSmem), stream>>>(on, tg, B, L, ws + m.gx, ws + m.gact, ws + m.cs, ws + m.hs, ws + m.z);
    {   // S[s] = ReLU(h_T Ws_eff^T + bs_eff)                                      models/qnet_rnn.py:135-136
        Gemm gs[3];
        for (int s = 0; s < 3; ++s) {
            const float *e = effw + (s == 2 ? EFF_NET : 0);
            gs[s] = gemm(ws + m.z + (int64_t)s * B * HID, HID, 1, e, 1, HID, ws + m.s + (int64_t)s * B * SH, SH, B, SH, HID, e + EFF_BS, nullptr, 1);
        }
        launch_gemms(gs, 3, stream);
    }
    drqn_loss_kernel<<<1, 256, 0, stream>>>(ring, rows, B, L, gamma, effw, ws + m.s, ws + m.ds, on, gr, noisy_on, loss_out, td_out);
    {   // dWs[j][k] = sum_b dS[b][j] h_T[b][k]  ;  dh_T[b][k] = sum_j dS[b][j] Ws_eff[j][k]
        Gemm gs[2];
        gs[0] = gemm(ws + m.ds, 1, SH, ws + m.z, HID, 1, gr.shared.grad_weight_mu, HID, SH, HID, B);
        gs[1] = gemm(ws + m.ds, SH, 1, effw, HID, 1, ws + m.dz, HID, B, HID, SH);
        launch_gemms(gs, 2, stream);
    }
    lstm_bwd_kernel<<<dim3(CL, B / BT), LSTM_THREADS, sizeof(BwdSmem), stream>>>(on, B, L, ws + m.gact, ws + m.cs, ws + m.hs, ws + m.dz, ws + m.dgx, ws + m.dwhh_part);
    {   // dW_ih[c][k] = sum_r dGx[r][c] F2[r][k]  ;  dF2[r][k] = sum_c dGx[r][c] W_ih[c][k]  x [F2 > 0]
        Gemm gs[2];
        gs[0] = gemm(ws + m.dgx, 1, GATES, ws + m.f2, FEAT, 1, gr.w_ih, FEAT, GATES, FEAT, R);
        gs[1] = gemm(ws + m.dgx, GATES, 1, on.w_ih, FEAT, 1, ws + m.df2, FEAT, R, FEAT, GATES, nullptr, nullptr, 0, ws + m.f2, FEAT);
        launch_gemms(gs, 2, stream);
    }
    {   // dW2[j][k] = sum_r dF2[r][j] F1[r][k]  ;  dF1[r][k] = sum_j dF2[r][j] W2[j][k]  x [F1 > 0]
        Gemm gs[2];
        gs[0] = gemm(ws + m.df2, 1, FEAT, ws + m.f1, F1D, 1, gr.f2_w, F1D, FEAT, F1D, R);
        gs[1] = gemm(ws + m.df2, FEAT, 1, on.f2_w, F1D, 1, ws + m.df1, F1D, R, F1D, FEAT, nullptr, nullptr, 0, ws + m.f1, F1D);
        launch_gemms(gs, 2, stream);
    }
    {   // dW1[j][k] = sum_r dF1[r][j] X[r][k]
        Gemm g1 = gemm(ws + m.df1, 1, F1D, ws + m.x0, OBSP, 1, gr.f0_w, OBS, F1D, OBS, R);
        launch_gemms(&g1, 1, stream);
    }
    drqn_finalize_kernel<<<64, 256, 0, stream>>>(B, L, ws + m.dgx, ws + m.df2, ws + m.df1, ws + m.ds, ws + m.dwhh_part, on, gr, noisy_on);
    return (int)cudaGetLastError();
}

int clip_grad_norm_launch(float *flat, int64_t numel, float max_norm, float *norm_out, float *scratch, cudaStream_t stream) {
    int64_t blocks = (numel + NORM_THREADS * 4 - 1) / (NORM_THREADS * 4);
    blocks = blocks < 1 ? 1 : (blocks > NORM_MAX_BLOCKS ? NORM_MAX_BLOCKS : blocks);
    grad_sqnorm_kernel<<<(unsigned)blocks, NORM_THREADS, 0, stream>>>(flat, numel, max_norm, scratch, norm_out);
    grad_scale_kernel<<<(unsigned)blocks, 256, 0, stream>>>(flat, numel, norm_out);
    return (int)cudaGetLastError();
}

int adam_multi_launch(const PPAdamParam *params, int32_t count, double lr, double beta1, double beta2, double eps, cudaStream_t stream) {
    AdamMulti pack{};
    int64_t total = 0;
    for (int i = 0; i < count; ++i) { pack.p[i] = params[i]; total += params[i].numel; }
    int64_t blocks = (total + 1023) / 1024;
    blocks = blocks < 1 ? 1 : (blocks > 296 ? 296 : blocks);
    adam_multi_kernel<<<(unsigned)blocks, 256, 0, stream>>>(pack, count, lr, beta1, beta2, eps);
    adam_bump_kernel<<<1, 32, 0, stream>>>(pack, count);
    return (int)cudaGetLastError();
}

}  // namespace pp

// ------------------------------------------------------------------------------------------ weight image for the tensor-core rollout
// QNetRNN in torch's layout -> the fp16 stage image PP_RNNTC_* (include/pong_b200.h) that selfplay_rnn_tc_kernel streams
// with TMA: what policy.pack_qnetrnn_tc builds with ~100 framework kernels (2.6 ms per training chunk), in one launch.
// Every tile is K-major, no swizzle: element (k, n) of a [K][N] tile sits at half index ((k / 8) * N + n) * 8 + k % 8.
namespace pp {
namespace {

__device__ __forceinline__ float noisy_w(const PPNoisyLayer &l, int64_t i, int noisy) {
    return noisy ? l.weight_mu[i] + l.weight_sigma[i] * l.weight_epsilon[i] : l.weight_mu[i];
}
__device__ __forceinline__ float noisy_b(const PPNoisyLayer &l, int64_t i, int noisy) {
    return noisy ? l.bias_mu[i] + l.bias_sigma[i] * l.bias_epsilon[i] : l.bias_mu[i];
}
__device__ __forceinline__ __half hi_of(float v) { return __float2half_rn(v); }
__device__ __forceinline__ __half lo_of(float v) { return __float2half_rn(v - __half2float(__float2half_rn(v))); }

// value of logical matrix `mat` at (k, n):  0 features.0^T (7 x 64)  1 features.2^T (64 x 128)  2 [W_ih | W_hh]^T (256 x 512, torch
// gate-major columns)  3 shared head^T (128 x 128)  4 heads^T (128 x 4: V, A0, A1, A2)
__device__ __forceinline__ float mat_at(const PPQNetRNNParams &p, int noisy, int mat, int k, int n) {
    switch (mat) {
        case 0: return p.f0_w[n * OBS + k];
        case 1: return p.f2_w[n * F1D + k];
        case 2: return k < FEAT ? p.w_ih[(int64_t)n * FEAT + k] : p.w_hh[(int64_t)n * HID + (k - FEAT)];
        case 3: return noisy_w(p.shared, (int64_t)n * HID + k, noisy);
        default: return n == 0 ? noisy_w(p.v, k, noisy) : noisy_w(p.a, (int64_t)(n - 1) * SH + k, noisy);
    }
}
__device__ __forceinline__ float bias_at(const PPQNetRNNParams &p, int noisy, int mat, int n) {
    switch (mat) {
        case 0: return p.f0_b[n];
        case 1: return p.f2_b[n];
        case 2: return p.b_ih[n] + p.b_hh[n];
        case 3: return noisy_b(p.shared, n, noisy);
        default: return n == 0 ? noisy_b(p.v, 0, noisy) : noisy_b(p.a, n - 1, noisy);
    }
}

// One tile of the image: `rows` K-rows x `cols` columns at byte offset `off`.  kind 0 = hi(w), 1 = lo(w), 2 = bias tile
// (16 x cols: bias hi in row 7, lo in row 15), 3 / 4 = the K = 16 first-layer tiles W1h' / W1l'.  Logical element (k, n) of
// the tile is matrix element (k0 + k, column map(n)); for the gate matrix column n of quarter q is gate n / 32, unit
// 32 q + n % 32, i.e. torch column (n / 32) * 128 + 32 q + n % 32.
struct PackTile { int32_t off, kind, mat, rows, cols, k0, quarter; };
struct PackPlan { PackTile t[96]; int count; };

__global__ void __launch_bounds__(256)
pack_qnetrnn_tc_kernel(const PPQNetRNNParams p, int noisy, const PackPlan plan, __half *__restrict__ img) {
    for (int ti = blockIdx.y; ti < plan.count; ti += gridDim.y) {
        const PackTile &t = plan.t[ti];
        __half *out = img + t.off / 2;
        const int total = t.rows * t.cols;
        for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
            const int kk = e & 7, n = (e >> 3) % t.cols, kc = e / (8 * t.cols), k = kc * 8 + kk;     // e is the half index in the tile
            const int col = t.mat == 2 ? (n / 32) * HID + 32 * t.quarter + (n % 32) : n;
            __half v = __float2half_rn(0.0f);
            const bool real = t.mat != 4 || n < 4;             // heads: columns 4..15 are padding (nothing to read there)
            if (t.kind == 0) { if (real) v = hi_of(mat_at(p, noisy, t.mat, t.k0 + k, col)); }
            else if (t.kind == 1) { if (real) v = lo_of(mat_at(p, noisy, t.mat, t.k0 + k, col)); }
            else if (t.kind == 2) {
                if (k == 7 && real) v = hi_of(bias_at(p, noisy, t.mat, col));
                else if (k == 15 && real) v = lo_of(bias_at(p, noisy, t.mat, col));
            } else if (t.kind == 3) {                          // rows 0..6 and 8..14: hi(W1^T); row 7: hi(b1); row 15: lo(b1)
                if ((k & 7) < 7) v = hi_of(mat_at(p, noisy, 0, k & 7, n));
                else v = k == 7 ? hi_of(bias_at(p, noisy, 0, n)) : lo_of(bias_at(p, noisy, 0, n));
            } else {                                           // rows 0..6: lo(W1^T)
                if (k < 7) v = lo_of(mat_at(p, noisy, 0, k, n));
            }
            out[e] = v;
        }
    }
}

}  // namespace

int pack_qnetrnn_tc_launch(const PPQNetRNNParams &p, int noisy, void *image, cudaStream_t stream) {
    static PackPlan plan = [] {
        PackPlan pl{};
        int off = 0;
        auto add = [&](int kind, int mat, int rows, int cols, int k0, int quarter) {
            pl.t[pl.count++] = PackTile{off, kind, mat, rows, cols, k0, quarter};
            off += rows * cols * 2;
        };
        add(3, 0, 16, 64, 0, 0); add(4, 0, 16, 64, 0, 0);                                     // S0
        add(0, 1, 64, 128, 0, 0); add(2, 1, 16, 128, 0, 0); add(1, 1, 64, 128, 0, 0);         // S1, S2
        for (int q = 0; q < 4; ++q) {
            for (int c = 0; c < 4; ++c) {
                add(0, 2, 64, 128, 64 * c, q);
                if (c == 0) add(2, 2, 16, 128, 0, q);
            }
            for (int c = 0; c < 4; ++c) add(1, 2, 64, 128, 64 * c, q);
        }
        add(0, 3, 64, 128, 0, 0); add(2, 3, 16, 128, 0, 0); add(0, 3, 64, 128, 64, 0);        // shared head
        add(1, 3, 64, 128, 0, 0); add(1, 3, 64, 128, 64, 0);
        add(0, 4, 128, 16, 0, 0); add(1, 4, 128, 16, 0, 0); add(2, 4, 16, 16, 0, 0);          // dueling heads
        return pl;
    }();
    pack_qnetrnn_tc_kernel<<<dim3(4, plan.count), 256, 0, stream>>>(p, noisy, plan, reinterpret_cast<__half *>(image));
    return (int)cudaGetLastError();
}

}  // namespace pp

// ------------------------------------------------------------------------------------------ sequence replay on a lock-step ring
// SequenceReplayBuffer.sample (scripts/train_rnn_iterative.py:126-141) draws a stored episode uniformly, then a window of
// `trace` consecutive steps uniformly inside it.  On the time-major ring [T][n] that is: a window ENDING at row (t, i) is
// eligible iff its episode is complete inside the ring, has len >= trace and the window lies inside it, with weight
// 1 / (len - trace + 1) — every stored episode carries total weight 1.  One thread per env walks its column once.
namespace pp {
namespace {

__global__ void __launch_bounds__(256)
seq_window_weights_kernel(const uint8_t *__restrict__ done, int64_t n, int64_t T, int64_t steps_written, int trace, int starts_fresh,
                          float *__restrict__ weights, unsigned long long *__restrict__ episodes) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool wrapped = steps_written > T;
    const int64_t shift = wrapped ? steps_written % T : 0;           // physical row of the oldest step
    unsigned count = 0;
    if (i < n) {
        auto at = [&](int64_t t) { int64_t r = t + shift; if (r >= T) r -= T; return r * n + i; };
        int64_t begin = 0;                                           // first row of the episode being walked
        bool cut = !(starts_fresh && !wrapped);                      // its start lies before the ring's oldest row
        for (int64_t t = 0; t < T; ++t) {
            if (done[at(t)]) {
                const int64_t len = t - begin + 1;
                const bool stored = !cut && len >= trace;
                const float wv = stored ? 1.0f / (float)(len - trace + 1) : 0.0f;
                for (int64_t e = begin; e <= t; ++e) weights[at(e)] = (e - begin + 1 >= trace) ? wv : 0.0f;
                count += stored ? 1u : 0u;
                begin = t + 1;
                cut = false;
            }
        }
        for (int64_t e = begin; e < T; ++e) weights[at(e)] = 0.0f;  // the episode still running (or rows not written yet)
    }
    count = __reduce_add_sync(0xffffffffu, count);
    if ((threadIdx.x & 31) == 0 && count) atomicAdd(episodes, (unsigned long long)count);
}

// sampled window ends (ring slots) -> the `trace` slots of each window in time order
__global__ void seq_expand_rows_kernel(const int64_t *__restrict__ end_slots, int batch, int trace, int64_t n, int64_t T,
                                       int64_t *__restrict__ rows) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= batch * trace) return;
    const int b = j / trace, k = j % trace;
    const int64_t slot = end_slots[b], env = slot % n;
    int64_t r = slot / n - (trace - 1) + k;
    if (r < 0) r += T;
    rows[j] = r * n + env;
}

}  // namespace

int seq_window_weights_launch(const uint8_t *done, int64_t n, int64_t T, int64_t steps_written, int trace, int starts_fresh,
                              float *weights, unsigned long long *episodes, cudaStream_t stream) {
    cudaError_t e = cudaMemsetAsync(episodes, 0, sizeof(unsigned long long), stream);
    if (e != cudaSuccess) return (int)e;
    seq_window_weights_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(done, n, T, steps_written, trace, starts_fresh, weights, episodes);
    return (int)cudaGetLastError();
}

int seq_expand_rows_launch(const int64_t *end_slots, int batch, int trace, int64_t n, int64_t T, int64_t *rows, cudaStream_t stream) {
    seq_expand_rows_kernel<<<(batch * trace + 255) / 256, 256, 0, stream>>>(end_slots, batch, trace, n, T, rows);
    return (int)cudaGetLastError();
}

}  // namespace pp
