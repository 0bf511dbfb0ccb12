// selfplay_tc_kernels.cu — K2a on the 5th-generation tensor cores (PP_PREC_F16) and the fused self-play rollout
// built on it.
//
// A *group* is 128 threads = 128 envs = the 128 rows (TMEM lanes) of one UMMA tile; a CTA holds four groups that
// run independently (named barriers, two mbarriers each), so one group's epilogue overlaps the others' MMA round
// trips, and there is one CTA per SM.  The QNet (models/qnet.py:71-75) of one player is two tcgen05.mma batches
// with fp16 operands and fp32 accumulation in TMEM, and the dueling heads in fp32 on the CUDA cores.  Activations AND
// weights are split x = x_hi + x_lo into two fp16 numbers (22 significant bits) and every product is taken as
// hi*hi + lo*hi + hi*lo, so Q-values carry ~fp32 accuracy (measured ~1e-6 relative) instead of fp16's 2^-12 — trained
// heads have |W| > 20, where a single fp16 pass is off by 3e-2:
//     L1  D[128x64] = X * W1h'^T + X * W1l'^T      X = [obs_hi(7) 1 | obs_lo(7) 1]   (2 MMAs, K = 16)
//                                                  W1h' = [W1_hi ; b1_hi | W1_hi ; b1_lo], W1l' = [W1_lo ; 0 | 0]
//     L2  D[128x64] = H1h*W2h^T + H1l*W2h^T + H1h*W2l^T (3 x 4 MMAs, K = 16 each) + X * B2'^T (bias via X's ones)
//     heads: (V, A0, A1, A2) = ReLU(D) . Wh + bh as packed fp32 FFMA2 per thread, then V + (A - mean A)
// Between the layers each thread reads ITS row of the accumulator (tcgen05.ld 32x32b), applies ReLU, splits and writes
// the row back to TENSOR MEMORY IN PLACE (tcgen05.st over the accumulator's own columns) as the next A operand: hidden
// activations never touch shared memory, the MMAs read A from TMEM (two fp16 per 32-bit column) and only the small
// weight tiles from shared memory.  A group owns 128 TMEM columns = two 64-column regions:
//     L1 -> R0;  epilogue: H1 in place over R0;  L2 (A = R0) -> R1;  heads read R1
// and the second player's L1 goes to R0 as soon as the first player's L2 has completed, under the first player's head
// epilogue (group_forward_both).
// The shared-memory operands (X rows, weight tiles) use the no-swizzle K-major canonical layout stored as
// [K/8][rows][8 halves]: a thread's 16-byte chunk stores are contiguous across the warp (conflict-free) and the
// descriptor strides are LBO = rows*16 B (next K chunk), SBO = 128 B (next 8-row core matrix).
// Both players' fp32 weight blobs arrive by one TMA bulk copy each (cp.async.bulk + mbarrier: the 1-D bulk form, no
// tensor map — the blobs are contiguous) into a staging area and are converted once per launch into fp16 B-operand
// tiles that stay in shared memory for all k steps.
// Env state, observations and bookkeeping never leave registers (same step_and_book as the CUDA-core kernel).
#include <cuda_fp16.h>

#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "pp_host.h"
#include "pp_rollout.cuh"
#include "tc_tiles.cuh"

namespace pp {

namespace {

constexpr int G_ROWS = 128;                    // envs per group
constexpr int CTA_GROUPS = 4;

// per-player weight tiles (bytes); every B tile is [K/8][N][16 B]; *H = fp16(w), *L = fp16(w - fp16(w))
constexpr uint32_t W1_BYTES = 2 * 64 * 16, W2_BYTES = 8 * 64 * 16, B2_BYTES = 2 * 64 * 16, W3_BYTES = 8 * 16 * 16,
                   B3_BYTES = 2 * 16 * 16;
constexpr uint32_t W1H_OFF = 0, W1L_OFF = W1H_OFF + W1_BYTES, W2H_OFF = W1L_OFF + W1_BYTES, W2L_OFF = W2H_OFF + W2_BYTES,
                   B2_OFF = W2L_OFF + W2_BYTES, W3H_OFF = B2_OFF + B2_BYTES, W3L_OFF = W3H_OFF + W3_BYTES,
                   B3_OFF = W3L_OFF + W3_BYTES, PLAYER_W_BYTES = B3_OFF + B3_BYTES;              // 27136
constexpr uint32_t ADV_TABLE_OFF = 66 * 16;     // second fp32 head table (advantages only) behind the 65 float4 of the first
static_assert(ADV_TABLE_OFF + 33 * 16 <= 2 * W3_BYTES + B3_BYTES, "head tables");
constexpr uint32_t X_BYTES = 2 * G_ROWS * 16;                                                    // 4096
constexpr uint32_t GROUP_BYTES = 2 * X_BYTES;                                                    // X rows of both players
constexpr uint32_t TM_R1 = 64;                  // TMEM columns of a group: region R0 at +0, R1 at +64 (64 columns each)
// A hidden-activation operand written IN PLACE over the accumulator it was computed from: accumulator columns
// [32 h, 32 h + 32) become packed hi pairs [32 h, +16) and packed lo pairs [32 h + 16, +16).  K step j (16 units) of the
// hi part therefore starts at column tm_hi(j), of the lo part at tm_hi(j) + 16.
__device__ __forceinline__ constexpr uint32_t tm_hi(int j) { return (uint32_t)((j >> 1) * 32 + (j & 1) * 8); }
constexpr uint32_t BLOB_BYTES = PP_QNET_BLOB_FLOATS * 4;                                         // 19728
constexpr uint32_t CTRL_BYTES = 128;

template <int GROUPS> struct SmemMap {
    static constexpr uint32_t W = 0;                                        // [2 players][PLAYER_W_BYTES]
    static constexpr uint32_t GROUPS_OFF = 2 * PLAYER_W_BYTES;              // [GROUPS][GROUP_BYTES]; start: blob staging
    static constexpr uint32_t CTRL = GROUPS_OFF + (GROUPS * GROUP_BYTES > 2 * BLOB_BYTES ? GROUPS * GROUP_BYTES : 2 * BLOB_BYTES);
    static constexpr uint32_t TOTAL = CTRL + CTRL_BYTES;                    // mbarriers + TMEM base
};
// control block (CTRL_BYTES): mbarriers [0] weights, [1 + 2 g + w] = barrier w of group g; the TMEM base slot follows them.
// A group alternates between TWO completion barriers: consecutive MMA batches of a group are not always separated by a
// group barrier (L1_B is committed right after L2_A has been observed), and a parity wait on ONE barrier cannot tell
// "phase k not yet complete" from "phases k and k + 1 both complete".  With two barriers a barrier's next phase is only
// committed after a group barrier that every thread reaches after having observed its previous phase.
template <int GROUPS> struct CtrlMap {
    static constexpr uint32_t TMEM_SLOT = (8 * (1 + 2 * GROUPS) + 15) / 16 * 16;
    static_assert(8 * (1 + 2 * GROUPS) <= TMEM_SLOT && TMEM_SLOT + 4 <= CTRL_BYTES, "TMEM base slot must not overlap the mbarriers");
};

struct PlayerTiles {     // shared-memory (generic) pointers of one player's operands
    uint8_t *w, *x;
};

// ---- the rollout's head table in CONSTANT memory -------------------------------------------------------------------
// Every thread of a warp needs the same 32 float4 of difference weights per player and step.  From shared memory each
// warp-wide broadcast load still writes 512 bytes of registers (two wavefronts of the shared-memory data pipe per
// instruction); from constant memory the values arrive through the UNIFORM datapath (LDCU into uniform registers, which
// FFMA2 takes as an operand): no register-file write, no load/store unit.  The tables are built per launch by
// head_diff_kernel into a device staging array and copied into the __constant__ bank with a stream-ordered
// device-to-device cudaMemcpyToSymbolAsync.  Two launches in flight on DIFFERENT streams must not share a table, so
// every (device, stream) gets its own slot the first time it launches; when the slots run out a launch uses the
// shared-memory table instead (HEADS_IN_CONST = false).  A launch captured into a CUDA graph keeps the slot of the
// stream it was captured on.
constexpr int HEAD_SLOTS = 16, HEAD_ENTRIES = 33;           // per player: 32 unit pairs + the bias entry
__constant__ float4 c_head_diff[HEAD_SLOTS][2][HEAD_ENTRIES];
__device__ float4 g_head_diff_staging[HEAD_SLOTS][2][HEAD_ENTRIES];

// entry kk < 32: (D01[2 kk], D01[2 kk + 1], D12[2 kk], D12[2 kk + 1]) with D01 = A0 - A1, D12 = A1 - A2; entry 32: biases
__device__ __forceinline__ float4 head_diff_entry(const float *blob, int idx) {
    if (idx < 32) {
        const float *w = blob + PP_QNET_WHT + (2 * idx) * 4;             // w[k * 4 + 1 + a] = A_a[2 idx + k]
        return make_float4(__fsub_rn(w[1], w[2]), __fsub_rn(w[5], w[6]), __fsub_rn(w[2], w[3]), __fsub_rn(w[6], w[7]));
    }
    const float *b = blob + PP_QNET_BH;
    return make_float4(__fsub_rn(b[1], b[2]), __fsub_rn(b[2], b[3]), 0.0f, 0.0f);
}
__global__ void head_diff_kernel(const float *blob_a, const float *blob_b, int slot) {
    const int player = threadIdx.x / 64, idx = threadIdx.x % 64;
    const float *blob = player ? blob_b : blob_a;
    if (idx < HEAD_ENTRIES && blob) g_head_diff_staging[slot][player][idx] = head_diff_entry(blob, idx);
}

// fp32 blob (staging) -> fp16 hi / lo B-operand tiles of one player.  All threads of the CTA.
__device__ void build_weight_tiles(uint8_t *wt, const float *blob, int tid, int nthreads) {
    auto H = [&](uint32_t off) { return reinterpret_cast<__half *>(wt + off); };
    auto hi = [](float v) { return __float2half_rn(v); };
    auto lo = [](float v) { return __float2half_rn(v - __half2float(__float2half_rn(v))); };
    const __half zero = __float2half_rn(0.0f);
    for (int idx = tid; idx < 2 * 64 * 8; idx += nthreads) {                 // [2][64][8]: k = 0..15
        const int k = (idx >> 9) * 8 + (idx & 7), nn = (idx >> 3) & 63;
        const bool bias_row = (k & 7) == 7;
        const float w1 = bias_row ? 0.0f : blob[PP_QNET_W1T + (k & 7) * 64 + nn];
        const float b1 = blob[PP_QNET_B1 + nn], b2v = blob[PP_QNET_B2 + nn];
        H(W1H_OFF)[idx] = bias_row ? (k == 7 ? hi(b1) : lo(b1)) : hi(w1);      // obs_hi and obs_lo both meet W1_hi
        H(W1L_OFF)[idx] = (!bias_row && k < 7) ? lo(w1) : zero;                // obs_hi * W1_lo
        H(B2_OFF)[idx] = bias_row ? (k == 7 ? hi(b2v) : lo(b2v)) : zero;       // rides on X's ones columns 7, 15
    }
    for (int idx = tid; idx < 8 * 64 * 8; idx += nthreads) {                 // W2 [8][64][8]
        const int k = (idx >> 9) * 8 + (idx & 7), nn = (idx >> 3) & 63;
        const float w = blob[PP_QNET_W2T + k * 64 + nn];
        H(W2H_OFF)[idx] = hi(w);
        H(W2L_OFF)[idx] = lo(w);
    }
    // dueling heads stay fp32: [64] x (V, A0, A1, A2) k-major + the bias row, read as broadcast float4 by heads_epilogue
    // (for the packed FFMA2 of sm_100: per pair of hidden units k0 = 2 kk, k1 = k0 + 1 two float4
    //  (V[k0], V[k1], A0[k0], A0[k1]) and (A1[k0], A1[k1], A2[k0], A2[k1]); entry 64 = the bias row)
    float4 *ht = reinterpret_cast<float4 *>(wt + W3H_OFF);
    for (int idx = tid; idx < 65; idx += nthreads) {
        if (idx < 64) {
            const int kk = idx >> 1, o = (idx & 1) * 2;                      // o: first of the two outputs in this float4
            const float *w0 = blob + PP_QNET_WHT + (2 * kk) * 4, *w1 = w0 + 4;
            ht[idx] = make_float4(w0[o], w1[o], w0[o + 1], w1[o + 1]);
        } else {
            const float *b = blob + PP_QNET_BH;
            ht[idx] = make_float4(b[0], b[1], b[2], b[3]);
        }
    }
    // For the fused rollout, which only needs argmax_a Q_a: Q_a = V + (A_a - mean(A)) orders the actions like A_a, and
    // three numbers are ordered by two differences, so the table holds  D01 = A0 - A1  and  D12 = A1 - A2  only: per four
    // hidden units k0..k3 two float4 (D01[k0], D01[k1], D12[k0], D12[k1]) (D01[k2], D01[k3], D12[k2], D12[k3]); entry 32 =
    // (bA0 - bA1, bA1 - bA2, 0, 0)
    float4 *dt = reinterpret_cast<float4 *>(wt + W3H_OFF + ADV_TABLE_OFF);
    for (int idx = tid; idx < 33; idx += nthreads) dt[idx] = head_diff_entry(blob, idx);
}

// 32 accumulator values -> ReLU -> packed fp16 hi pairs and lo pairs.  hi = rz(relu(x)) <= relu(x), so for x >= 0 the
// residual x - hi is >= 0 and for x < 0 it is x < 0: one more ReLU-convert yields lo = relu(x) - hi.
__device__ __forceinline__ void split32(const uint32_t (&r0)[16], const uint32_t (&r1)[16], uint32_t (&hi)[16], uint32_t (&lo)[16]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {                                             // residuals as one packed FADD2 per pair
        const float a = __uint_as_float(r0[2 * j]), b = __uint_as_float(r0[2 * j + 1]);
        hi[j] = tc::pack_f16x2_rz_relu(a, b);
        const float2 ra = __fadd2_rn(make_float2(a, b), make_float2(-h2f(hi[j], 0), -h2f(hi[j], 1)));
        lo[j] = tc::pack_f16x2<true>(ra.x, ra.y);
        const float c = __uint_as_float(r1[2 * j]), d = __uint_as_float(r1[2 * j + 1]);
        hi[8 + j] = tc::pack_f16x2_rz_relu(c, d);
        const float2 rc = __fadd2_rn(make_float2(c, d), make_float2(-h2f(hi[8 + j], 0), -h2f(hi[8 + j], 1)));
        lo[8 + j] = tc::pack_f16x2<true>(rc.x, rc.y);
    }
}

// Accumulator row (64 fp32 columns at `acc`) -> ReLU -> hi/lo fp16 -> written IN PLACE over the same columns as this
// thread's row of the next A operand (layout: tm_hi()).  Hidden activations never touch shared memory.
__device__ __forceinline__ void hidden_epilogue_inplace(uint32_t acc) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        uint32_t r0[16], r1[16], hi[16], lo[16];
        tc::tmem_ld16(acc + half * 32, r0);
        tc::tmem_ld16(acc + half * 32 + 16, r1);
        tc::tmem_ld_wait();
        split32(r0, r1, hi, lo);
        tc::tmem_st16(acc + half * 32, hi);
        tc::tmem_st16(acc + half * 32 + 16, lo);
    }
    tc::tmem_st_wait();
}

// Second hidden layer's accumulator row (64 fp32 columns at `src`) -> ReLU -> the four head outputs in fp32 on the CUDA
// cores (256 FMAs against a broadcast float4 table) -> dueling Q.  This replaces a third MMA batch (13 small MMAs at the
// 45-cycle instruction floor, one more accumulator round trip and group barrier per player) and keeps the heads exact.
__device__ __forceinline__ void heads_epilogue(uint32_t src, const uint8_t *table, float (&q)[3]) {
    const float4 *ht = reinterpret_cast<const float4 *>(table);
    // (even-k, odd-k) partial sums per output, in TWO independent sets (pairs j even / j odd): eight FFMA2 chains in
    // flight instead of four — the epilogue is bound by the latency of its dependent FMAs, not by their number
    float2 v[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)}, a0[2] = {v[0], v[0]}, a1[2] = {v[0], v[0]}, a2[2] = {v[0], v[0]};
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        uint32_t r0[16], r1[16];
        tc::tmem_ld16(src + half * 32, r0);
        tc::tmem_ld16(src + half * 32 + 16, r1);
        tc::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) {                                          // 16 pairs of hidden units
            const uint32_t *r = j < 8 ? r0 + 2 * j : r1 + 2 * (j - 8);
            const float2 h = make_float2(fmaxf(__uint_as_float(r[0]), 0.0f), fmaxf(__uint_as_float(r[1]), 0.0f));
            const float4 w01 = ht[(half * 16 + j) * 2], w23 = ht[(half * 16 + j) * 2 + 1];
            const int s = j & 1;
            v[s] = __ffma2_rn(make_float2(w01.x, w01.y), h, v[s]);
            a0[s] = __ffma2_rn(make_float2(w01.z, w01.w), h, a0[s]);
            a1[s] = __ffma2_rn(make_float2(w23.x, w23.y), h, a1[s]);
            a2[s] = __ffma2_rn(make_float2(w23.z, w23.w), h, a2[s]);
        }
    }
    const float4 bias = ht[64];
    const float2 vs = __fadd2_rn(v[0], v[1]), a0s = __fadd2_rn(a0[0], a0[1]), a1s = __fadd2_rn(a1[0], a1[1]), a2s = __fadd2_rn(a2[0], a2[1]);
    const float V = __fadd_rn(__fadd_rn(vs.x, vs.y), bias.x), A0 = __fadd_rn(__fadd_rn(a0s.x, a0s.y), bias.y);
    const float A1 = __fadd_rn(__fadd_rn(a1s.x, a1s.y), bias.z), A2 = __fadd_rn(__fadd_rn(a2s.x, a2s.y), bias.w);
    const float mean = __fdiv_rn(__fadd_rn(__fadd_rn(A0, A1), A2), 3.0f);               // V + (A - mean(A))  models/qnet.py:75
    q[0] = __fadd_rn(V, __fsub_rn(A0, mean));
    q[1] = __fadd_rn(V, __fsub_rn(A1, mean));
    q[2] = __fadd_rn(V, __fsub_rn(A2, mean));
}

// The fused rollout only ever takes argmax_a Q_a.  Q_a = V + (A_a - mean(A)) orders the actions like A_a, and the order of
// three numbers follows from two differences, so the rollout evaluates  d01 = A0 - A1  and  d12 = A1 - A2  only (difference
// weights precomputed per launch): 128 instead of 256 FMAs and, what matters more, 32 instead of 64 broadcast LDS.128
// per player — the head tables are the largest consumer of the shared-memory data pipe, which ncu shows 52 % busy with
// LSU wavefronts on top of the tensor core's own 23 %.  Returned as a one-hot triple whose first-argmax is the action.
// In fp32 this can differ from argmax Q only where two Q values are within rounding of each other.
template <bool HEADS_IN_CONST>
__device__ __forceinline__ void advantages_epilogue(uint32_t src, const uint8_t *table, int cslot, int player, float (&adv)[3]) {
    const float4 *dt = reinterpret_cast<const float4 *>(table + ADV_TABLE_OFF);
    const float2 z = make_float2(0.f, 0.f);
    float2 d01[4] = {z, z, z, z}, d12[4] = {z, z, z, z};                        // eight independent FFMA2 chains
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        uint32_t r0[16], r1[16];
        tc::tmem_ld16(src + half * 32, r0);
        tc::tmem_ld16(src + half * 32 + 16, r1);
        tc::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) {                                          // 16 pairs of hidden units
            const uint32_t *r = j < 8 ? r0 + 2 * j : r1 + 2 * (j - 8);
            const float2 h = make_float2(fmaxf(__uint_as_float(r[0]), 0.0f), fmaxf(__uint_as_float(r[1]), 0.0f));
#ifdef PP_TC_NOLDS
            const float4 w = make_float4(0.25f + j, 0.5f - j, 0.125f * half, 1.0f);   // experiment (wrong results): no table reads
#else
            const float4 w = HEADS_IN_CONST ? c_head_diff[cslot][player][half * 16 + j] : dt[half * 16 + j];
#endif
            d01[j & 3] = __ffma2_rn(make_float2(w.x, w.y), h, d01[j & 3]);
            d12[j & 3] = __ffma2_rn(make_float2(w.z, w.w), h, d12[j & 3]);
        }
    }
    const float4 bias = HEADS_IN_CONST ? c_head_diff[cslot][player][32] : dt[32];
    const float2 s01 = __fadd2_rn(__fadd2_rn(d01[0], d01[1]), __fadd2_rn(d01[2], d01[3]));
    const float2 s12 = __fadd2_rn(__fadd2_rn(d12[0], d12[1]), __fadd2_rn(d12[2], d12[3]));
    const float a01 = __fadd_rn(__fadd_rn(s01.x, s01.y), bias.x), a12 = __fadd_rn(__fadd_rn(s12.x, s12.y), bias.y);
    // first maximum, decided on the differences themselves (a sum like (A0 - A2, A1 - A2, 0) could absorb a tiny d01)
    const float a02 = __fadd_rn(a01, a12);
    const int pick = (a01 >= 0.0f && a02 >= 0.0f) ? 0 : (a12 >= 0.0f ? 1 : 2);
    adv[0] = pick == 0 ? 1.0f : 0.0f;
    adv[1] = pick == 1 ? 1.0f : 0.0f;
    adv[2] = pick == 2 ? 1.0f : 0.0f;
}
// HEADS: which head evaluation a kernel uses
constexpr int HEADS_FULL_Q = 0,        // the dueling Q values themselves (standalone action kernel, which can return them)
              HEADS_DIFF_SMEM = 1,     // action from the two advantage differences, table in shared memory
              HEADS_DIFF_CONST = 2;    // ... table in constant memory, slot `cslot`
template <int HEADS> __device__ __forceinline__ void heads(uint32_t src, const uint8_t *table, int cslot, int player, float (&q)[3]) {
    if (HEADS == HEADS_FULL_Q) heads_epilogue(src, table, q);
    else advantages_epilogue<HEADS == HEADS_DIFF_CONST>(src, table, cslot, player, q);
}

// MMA batches, issued by one thread per group.  d / a_tm are TMEM addresses with lane 0.
__device__ __forceinline__ void issue_l1(uint32_t d, const PlayerTiles &p) {
    const uint64_t x = tc::smem_desc(tc::smem_u32(p.x), A_LBO, SBO);
    tc::umma_f16(d, x, tc::smem_desc(tc::smem_u32(p.w + W1H_OFF), 64 * 16, SBO), tc::idesc_f16(128, 64), false);
    tc::umma_f16(d, x, tc::smem_desc(tc::smem_u32(p.w + W1L_OFF), 64 * 16, SBO), tc::idesc_f16(128, 64), true);
}
// D = Hh*Wh + Hl*Wh + Hh*Wl + X*B'  (second layer, N = 64): H = the in-place operand in TMEM at a_tm
__device__ __forceinline__ void issue_l2(uint32_t d, uint32_t a_tm, const PlayerTiles &p) {
    const uint32_t wh = tc::smem_u32(p.w + W2H_OFF), wl = tc::smem_u32(p.w + W2L_OFF);
    constexpr uint32_t B_LBO = 64 * 16;
#pragma unroll
    for (int pass = 0; pass < 3; ++pass) {
        const uint32_t b = pass == 2 ? wl : wh;
#pragma unroll
        for (int j = 0; j < 4; ++j)         // K = 16 per MMA = 8 TMEM columns of A, 2 shared-memory chunks of B
            tc::umma_f16_ts(d, a_tm + tm_hi(j) + (pass == 1 ? 16u : 0u), tc::smem_desc(b + j * 2 * B_LBO, B_LBO, SBO),
                            tc::idesc_f16(128, 64), (pass | j) != 0);
    }
    tc::umma_f16(d, tc::smem_desc(tc::smem_u32(p.x), A_LBO, SBO), tc::smem_desc(tc::smem_u32(p.w + B2_OFF), B_LBO, SBO),
                 tc::idesc_f16(128, 64), true);
}

// Shared prologue: barriers, TMEM, weights.  Returns the TMEM base of the CTA.
template <int GROUPS, typename M>
__device__ __forceinline__ uint32_t tc_prologue(uint8_t *smem, const PPPolicy &pol_a, const PPPolicy &pol_b) {
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + M::CTRL);           // [0] weights, [1 + 2 g + w] group g
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + M::CTRL + CtrlMap<GROUPS>::TMEM_SLOT);
    const int tid = threadIdx.x;
    const bool qa = pol_a.kind == PP_POLICY_QNET, qb = pol_b.kind == PP_POLICY_QNET;
    if (tid == 0) {
        for (int b = 0; b < 1 + 2 * GROUPS; ++b) tc::mbar_init(bars + b, 1);
        tc::fence_mbar_init();
    }
    if (tid < 32) tc::tmem_alloc<GROUPS * 128>(tmem_slot);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    uint8_t *stage = smem + M::GROUPS_OFF;
    if (tid == 0 && (qa || qb)) {
        tc::mbar_expect_tx(bars, (qa ? BLOB_BYTES : 0) + (qb ? BLOB_BYTES : 0));
        if (qa) tc::tma_bulk_g2s(stage, pol_a.weights, BLOB_BYTES, bars);
        if (qb) tc::tma_bulk_g2s(stage + BLOB_BYTES, pol_b.weights, BLOB_BYTES, bars);
    }
    if (qa || qb) tc::mbar_wait(bars, 0);
    if (qa) build_weight_tiles(smem + M::W, reinterpret_cast<const float *>(stage), tid, GROUPS * G_ROWS);
    if (qb) build_weight_tiles(smem + M::W + PLAYER_W_BYTES, reinterpret_cast<const float *>(stage + BLOB_BYTES), tid, GROUPS * G_ROWS);
    tc::fence_proxy_async();
    __syncthreads();                              // staging (aliases the group tiles) is dead from here on
    return tmem;
}

template <int GROUPS> __device__ __forceinline__ void tc_epilogue(uint32_t tmem) {
    tc::tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tc::tmem_dealloc<GROUPS * 128>(tmem);
}

// Phase timers of a -DPP_TC_TIMING build (diagnostics only): cycles per phase of a lock-step step, accumulated by every
// thread and printed for lane 0 of the first two warps of group 0 of CTA 0.
#ifdef PP_TC_TIMING
struct PhaseTimer {
    long long t[12] = {}, last = 0;
    __device__ __forceinline__ void start() { last = clock64(); }
    __device__ __forceinline__ void tick(int i) { const long long now = clock64(); t[i] += now - last; last = now; }
};
#define PP_TICK(i) g.timer.tick(i)
#else
#define PP_TICK(i) do {} while (0)
#endif

// One policy evaluation round for a group: X rows are already written and published by a group barrier.
struct GroupCtx {
#ifdef PP_TC_TIMING
    PhaseTimer timer;
#endif
    PlayerTiles pa, pb;
    uint64_t *bar;                               // two completion barriers, used alternately: bar[0] = L1 batches, bar[1] = L2
    uint32_t parity[2], r0, bar_id, lane_addr;   // r0: the group's first TMEM column (lane 0); R1 = r0 + TM_R1
    int row, cslot;                              // cslot: this launch's table in c_head_diff (HEADS_DIFF_CONST)
    bool qa, qb, issuer_warp;
};

template <int W> __device__ __forceinline__ void group_wait(GroupCtx &g) {
    tc::mbar_wait(g.bar + W, g.parity[W]);
    g.parity[W] ^= 1u;
    tc::tc_fence_after();
}

// an elected lane of the group's first warp runs `issue` and commits to completion barrier W of the group
template <int W, typename F> __device__ __forceinline__ void group_issue(GroupCtx &g, F &&issue) {
    if (g.issuer_warp) {                      // warp-uniform branch
        if (tc::elect_one()) {
            tc::tc_fence_after();
            issue();
            tc::umma_commit(g.bar + W);
        }
        __syncwarp();
    }
}

// ONE QNet player p: L1 -> R0; H in place; L2 -> R1; (ReLU, fp32 heads on the CUDA cores) -> Q
template <int HEADS>
__device__ __forceinline__ void group_forward_one(GroupCtx &g, const PlayerTiles &p, int player, float (&q)[3]) {
    group_issue<0>(g, [&] { issue_l1(g.r0, p); });
    group_wait<0>(g);
    hidden_epilogue_inplace(g.r0 + g.lane_addr);
    tc::tc_fence_before();
    tc::bar_sync(g.bar_id, G_ROWS);
    group_issue<1>(g, [&] { issue_l2(g.r0 + TM_R1, g.r0, p); });
    group_wait<1>(g);
    heads<HEADS>(g.r0 + TM_R1 + g.lane_addr, p.w + W3H_OFF, g.cslot, player, q);
}

// BOTH players (the self-play hot path), two 64-column TMEM regions R0, R1 per group:
//   L1_A -> R0;  H_A IN PLACE over R0;  L2_A (A = R0) -> R1
//   as soon as L2_A has completed R0 is dead:  L1_B -> R0 is issued BEFORE heads_A (reads R1), whose ~260 instructions
//   cover its flight;  H_B in place over R0;  L2_B (A = R0) -> R1;  heads_B
// Three exposed MMA round trips and three group barriers per lock-step step (the serial chain with separate operand
// regions had four and four).  [A variant that also freed R1 early by sending H_B through a SHARED-MEMORY A tile, so
// that L2_B could follow L2_A without a drain, was 7 % SLOWER: 12 SS-mode MMAs read 48 KB of A operand per group-step
// from shared memory and the 32 KB of epilogue stores compete with the head table's LDS traffic.]
template <int HEADS>
__device__ __forceinline__ void group_forward_both(GroupCtx &g, float (&q_a)[3], float (&q_b)[3]) {
    group_issue<0>(g, [&] { issue_l1(g.r0, g.pa); });
    PP_TICK(1);
    group_wait<0>(g);
    PP_TICK(2);
    hidden_epilogue_inplace(g.r0 + g.lane_addr);
    tc::tc_fence_before();
    PP_TICK(3);
    tc::bar_sync(g.bar_id, G_ROWS);
    PP_TICK(4);
    group_issue<1>(g, [&] { issue_l2(g.r0 + TM_R1, g.r0, g.pa); });
    PP_TICK(5);
    group_wait<1>(g);
    PP_TICK(6);
    group_issue<0>(g, [&] { issue_l1(g.r0, g.pb); });     // no group barrier since the last commit: the OTHER completion barrier
    PP_TICK(1);
    heads<HEADS>(g.r0 + TM_R1 + g.lane_addr, g.pa.w + W3H_OFF, g.cslot, 0, q_a);
    PP_TICK(7);
    group_wait<0>(g);
    PP_TICK(2);
    hidden_epilogue_inplace(g.r0 + g.lane_addr);
    tc::tc_fence_before();
    PP_TICK(3);
    tc::bar_sync(g.bar_id, G_ROWS);                 // also: every thread has read its heads_A row of R1
    PP_TICK(4);
    group_issue<1>(g, [&] { issue_l2(g.r0 + TM_R1, g.r0, g.pb); });
    PP_TICK(5);
    group_wait<1>(g);
    PP_TICK(6);
    heads<HEADS>(g.r0 + TM_R1 + g.lane_addr, g.pb.w + W3H_OFF, g.cslot, 1, q_b);
    PP_TICK(7);
}

template <int HEADS>
__device__ __forceinline__ void group_forward(GroupCtx &g, float (&q_a)[3], float (&q_b)[3]) {
    if (g.qa && g.qb) group_forward_both<HEADS>(g, q_a, q_b);
    else {
        if (g.qa) group_forward_one<HEADS>(g, g.pa, 0, q_a);
        if (g.qa && g.qb) {                    // R0 / R1 still hold the first player's operands until everyone has read them
            tc::tc_fence_before();
            tc::bar_sync(g.bar_id, G_ROWS);
        }
        if (g.qb) group_forward_one<HEADS>(g, g.pb, 1, q_b);
    }
}

// `grp` and `tmem` must be warp-uniform VALUES THE COMPILER CAN SEE as uniform (shfl broadcasts), so that the UMMA
// descriptors derived from them live in uniform registers instead of being re-broadcast before every MMA.
__device__ __forceinline__ GroupCtx make_group(uint8_t *smem, uint32_t groups_off, uint32_t ctrl_off, uint32_t tmem, int grp,
                                               int row, bool qa, bool qb) {
    GroupCtx g;
    uint8_t *gb = smem + groups_off + grp * GROUP_BYTES;
    g.pa = PlayerTiles{smem, gb};
    g.pb = PlayerTiles{smem + PLAYER_W_BYTES, gb + X_BYTES};
    g.row = row;
    g.bar = reinterpret_cast<uint64_t *>(smem + ctrl_off) + 1 + 2 * grp;
    g.parity[0] = g.parity[1] = 0;
    g.r0 = tmem + grp * 128;
    g.bar_id = 1 + grp;
    g.lane_addr = (uint32_t)((row >> 5) * 32) << 16;
    g.qa = qa; g.qb = qb;
    g.issuer_warp = __shfl_sync(0xffffffffu, (row >> 5) == 0 ? 1 : 0, 0) != 0;
    return g;
}

}  // namespace

// --------------------------------------------------------------------------------- standalone action selection
// obs[n][7] -> Q -> action for ONE player (as player A of a one-group CTA); the debugging and Q-parity vehicle.
__global__ void __launch_bounds__(G_ROWS)
qnet_act_tc_kernel(int64_t n, const float *__restrict__ obs, const PPPolicy pol, uint64_t seed, uint32_t step_index,
                   int64_t env_id_base, uint32_t stream_id, uint8_t *__restrict__ actions, float *__restrict__ q_out) {
    extern __shared__ __align__(128) uint8_t smem[];
    using M = SmemMap<1>;
    PPPolicy none = pol;
    none.kind = PP_POLICY_RANDOM;
    const uint32_t tmem = __shfl_sync(0xffffffffu, tc_prologue<1, SmemMap<1>>(smem, pol, none), 0);
    const int row = threadIdx.x;
    GroupCtx g = make_group(smem, M::GROUPS_OFF, M::CTRL, tmem, 0, row, true, false);
    for (int64_t tile0 = (int64_t)blockIdx.x * G_ROWS; tile0 < n; tile0 += (int64_t)gridDim.x * G_ROWS) {
        const int64_t i = tile0 + row;
        float o[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (i < n) {
#pragma unroll
            for (int k = 0; k < 7; ++k) o[k] = obs[i * 7 + k];
        }
        write_x_row(g.pa.x, row, o);
        tc::fence_proxy_async();
        tc::tc_fence_before();
        tc::bar_sync(g.bar_id, G_ROWS);
        float q[3], unused[3];
        group_forward<HEADS_FULL_Q>(g, q, unused);
        if (i < n) {
            int a = argmax3(q);
            a = explore(a, pol.eps_threshold, seed, (uint32_t)(env_id_base + i), step_index, stream_id);
            actions[i] = (uint8_t)a;
            if (q_out) { q_out[i * 3 + 0] = q[0]; q_out[i * 3 + 1] = q[1]; q_out[i * 3 + 2] = q[2]; }
        }
    }
    tc_epilogue<1>(tmem);
}

// --------------------------------------------------------------------------------- fused self-play rollout
// 16 warps = 4 groups per CTA, one CTA per SM; an elected lane of each group's first warp issues that group's MMAs.
// (A dedicated issuer warp fed through ready / done mbarriers was measured 27% SLOWER: one thread then serialises
// the MMA issue of all four groups, and tcgen05.mma issue is paced by the operand reads, ~1000 cycles per batch.)
// Work is split in units of warps (32 envs): `n_chunks` chunks of at most 16 warps, balanced to within one warp, so
// every SM issues the same number of warp-steps whatever n is.  A CTA walks chunks blockIdx.x, + gridDim.x, ...
constexpr int TC_FUSED_THREADS = G_ROWS * CTA_GROUPS;
struct FusedMap {
    static constexpr uint32_t W = 0, GROUPS_OFF = 2 * PLAYER_W_BYTES;
    static constexpr uint32_t SERVE_OFF = GROUPS_OFF + CTA_GROUPS * GROUP_BYTES;      // next serve per thread: 3 doubles
    static constexpr uint32_t CTRL = SERVE_OFF + TC_FUSED_THREADS * 24;                // mbarriers + TMEM base
    static constexpr uint32_t STAGE_OFF = (CTRL + CTRL_BYTES + 127) / 128 * 128;       // replay-row staging: 896 B per warp
    static constexpr uint32_t TOTAL = STAGE_OFF + (TC_FUSED_THREADS / 32) * 896;
};
static_assert(FusedMap::TOTAL <= 232448, "shared memory of the fused tensor-core kernel exceeds 227 KB");
static_assert(2 * BLOB_BYTES <= CTA_GROUPS * GROUP_BYTES + TC_FUSED_THREADS * 24, "weight staging aliases the X rows and serve slots");

template <typename R, bool PRECLAIM, bool CHEADS>
__global__ void __launch_bounds__(TC_FUSED_THREADS, 1)
selfplay_tc_kernel(const PPParams params, const PPEnvState st, int64_t n, int64_t k_steps, const PPPolicy pol_a,
                   const PPPolicy pol_b, uint64_t seed, int64_t step_base, const PPServeSource src, int32_t quota,
                   int64_t env_id_base, const PPRolloutOut out, const PPReplayRing ring, int64_t n_chunks, int cslot) {
    extern __shared__ __align__(128) uint8_t smem[];
    using M = FusedMap;
    const uint32_t tmem = __shfl_sync(0xffffffffu, tc_prologue<CTA_GROUPS, M>(smem, pol_a, pol_b), 0);
    const int warp_id = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);      // warp-uniform for the compiler too
    const int grp = warp_id >> 2, gw = warp_id & 3, row = threadIdx.x & 127, lane = threadIdx.x & 31;
    const bool qa = pol_a.kind == PP_POLICY_QNET, qb = pol_b.kind == PP_POLICY_QNET;
    GroupCtx g = make_group(smem, M::GROUPS_OFF, M::CTRL, tmem, grp, row, qa, qb);
    g.cslot = cslot;
    double *serve_slot = reinterpret_cast<double *>(smem + M::SERVE_OFF) + threadIdx.x * 3;
    float *row_stage = reinterpret_cast<float *>(smem + M::STAGE_OFF) + warp_id * 224;
    const EnvConsts<R> c(params);
    const StatePtrs<R> s(st);
    const int64_t total_warps = (n + 31) / 32;
    const int64_t ring_t0 = ring_first_step(ring, n, k_steps);
    const bool prefetch_serves = src.kind == PP_SERVE_PHILOX;
    // a serve QUEUE drawn from Philox (the host-buffer evaluation): envs also draw their next serve ahead of time, all
    // lanes together — but they have to CLAIM it first, and an env that holds a claimed serve will play it, so claiming
    // ahead stops 2 n serves before the end of the queue: the last serves go to whichever envs finish first (no tail of
    // envs that sit on a second episode while others idle)
    constexpr bool preclaim_serves = PRECLAIM;           // its own instantiation: the main path does not carry the code
    int next_q = -1;
    Tally total;

#pragma unroll 1
    for (int64_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
        const int64_t w_lo = chunk * total_warps / n_chunks, w_hi = (chunk + 1) * total_warps / n_chunks;
        const int nw = (int)(w_hi - w_lo), base = nw / CTA_GROUPS, rem = nw % CTA_GROUPS;
        const int my_warps = base + (grp < rem ? 1 : 0);                       // groups differ by at most one warp
        if (my_warps == 0) continue;                                           // uniform per group
        const int64_t i = (w_lo + grp * base + (grp < rem ? grp : rem) + gw) * 32 + lane;
        const bool valid = gw < my_warps && i < n;
        const int64_t ic = valid ? i : 0;
        Lane<R> L;
        L.e = load_env<R>(s, ic);
        L.ep_idx = s.ep_idx[ic]; L.ep_len = s.ep_len[ic];
        const uint32_t gid = (uint32_t)(env_id_base + ic);
        bool have_next = false;

#pragma unroll 1
        for (int64_t t = 0; t < k_steps; ++t) {
            const bool active = valid && !(quota > 0 && L.ep_idx >= quota);
            const uint32_t step = (uint32_t)(step_base + t);
            // Every 8 steps each thread that used up its prefetched serve draws the next one NOW, all lanes and warps of
            // the group together, instead of alone inside the divergent episode-end branch (where one finishing lane
            // would hold up its warp, and the warp its group, on almost every step).
            if ((t & 7) == 0 && prefetch_serves && active && !have_next && !(quota > 0 && L.ep_idx + 1 >= quota)) {
                philox_serve(params, src.seed, gid, (uint32_t)(L.ep_idx + 1), serve_slot[0], serve_slot[1], serve_slot[2]);
                have_next = true;
            }
            if ((t & 7) == 0 && preclaim_serves) {
                const bool want = active && !have_next && L.ep_idx != 0x7fffffff;
                const unsigned m_want = __ballot_sync(0xffffffffu, want);
                if (m_want && (long long)*reinterpret_cast<volatile unsigned long long *>(src.queue_head) + 2 * n < src.queue_total) {
                    const int q = claim_serves(want, m_want, src);
                    if (want && q != 0x7fffffff) {
                        philox_serve(params, src.seed, (uint32_t)(env_id_base + q % n), (uint32_t)(q / n), serve_slot[0], serve_slot[1], serve_slot[2]);
                        have_next = true;
                        next_q = q;
                    }
                }
            }
#ifdef PP_TC_TIMING
            if (t == 0) g.timer.start();
#endif
            float oa[7], ob[7];
            observe<R>(L.e, oa, ob);
            if (qa) write_x_row(g.pa.x, row, oa);
            if (qb) write_x_row(g.pb.x, row, ob);
            tc::fence_proxy_async();
            tc::tc_fence_before();
            PP_TICK(0);
            if (!tc::bar_red_or(g.bar_id, G_ROWS, active)) break;             // whole group frozen by the quota: for good
            PP_TICK(10);
            float q_a[3] = {0.f, 0.f, 0.f}, q_b[3] = {0.f, 0.f, 0.f};
            if (qa || qb) group_forward<CHEADS ? HEADS_DIFF_CONST : HEADS_DIFF_SMEM>(g, q_a, q_b);
            int act_a, act_b;
            if (pol_a.kind == PP_POLICY_RANDOM) act_a = random_action(seed, gid, step, STREAM_ACT_A);
            else act_a = explore(qa ? argmax3(q_a) : follower_action(oa, pol_a.follower_tol), pol_a.eps_threshold, seed, gid, step, STREAM_ACT_A);
            if (pol_b.kind == PP_POLICY_RANDOM) act_b = random_action(seed, gid, step, STREAM_ACT_B);
            else act_b = explore(qb ? argmax3(q_b) : follower_action(ob, pol_b.follower_tol), pol_b.eps_threshold, seed, gid, step, STREAM_ACT_B);
            PP_TICK(8);
            if (gw < my_warps) {      // warp-uniform: warps without envs skip the bookkeeping collectives entirely
                auto serve = [&](int ep, R &vx, R &vy, R &sp) {
                    if (have_next) { vx = (R)serve_slot[0]; vy = (R)serve_slot[1]; sp = (R)serve_slot[2]; have_next = false; }
                    else next_serve<R>(params, src, n, i, env_id_base, ep, vx, vy, sp);
                };
                const int pre = (preclaim_serves && have_next) ? next_q : -1;
                step_and_book<R, PRECLAIM>(c, L, active, act_a, act_b, ob, t, n, i, env_id_base, quota, out, ring,
                                           ring.head != nullptr && t >= ring_t0, src, serve, row_stage, pre);
            }
            PP_TICK(9);
        }
#ifdef PP_TC_TIMING
        if (blockIdx.x == 0 && grp == 0 && gw < 2 && lane == 0) {
            const char *names[11] = {"serve prefetch + obs + X rows", "MMA issue L1", "wait L1", "hidden epilogue", "group barrier",
                                     "MMA issue L2", "wait L2", "heads epilogue", "argmax / explore", "env step + bookkeeping",
                                     "step barrier (skew of the group)"};
            long long tot = 0;
            for (int q = 0; q < 11; ++q) tot += g.timer.t[q];
            for (int q = 0; q < 11; ++q)
                printf("warp %d  %-34s %8.0f cycles / step\n", gw, names[q], (double)g.timer.t[q] / (double)k_steps);
            printf("warp %d  %-34s %8.0f cycles / step\n", gw, "STEP TOTAL", (double)tot / (double)k_steps);
        }
#endif
        if (valid) {
            store_env<R>(s, i, L.e);
            s.ep_idx[i] = L.ep_idx;
            s.ep_len[i] = L.ep_len;
        }
        total.add(L.tally);
    }
    total.flush(out.counters, out.ep_log ? nullptr : out.ep_log_count);
    tc_epilogue<CTA_GROUPS>(tmem);
}

// --------------------------------------------------------------------------------- launchers
static int sm_count() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
            sms = 148;
    }
    return sms;
}

int qnet_act_tc_launch(int64_t n, const float *obs, const PPPolicy &pol, uint64_t seed, int64_t step_index,
                       int64_t env_id_base, int32_t stream_id, uint8_t *actions, float *q_out, cudaStream_t stream) {
    constexpr int smem = (int)SmemMap<1>::TOTAL;
    cudaError_t err = cudaFuncSetAttribute(qnet_act_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (err != cudaSuccess) return (int)err;
    const int64_t tiles = (n + G_ROWS - 1) / G_ROWS;
    const int64_t cap = (int64_t)sm_count() * 2;
    const unsigned blocks = (unsigned)(tiles < cap ? tiles : cap);
    qnet_act_tc_kernel<<<blocks, G_ROWS, smem, stream>>>(n, obs, pol, seed, (uint32_t)step_index, env_id_base,
                                                        (uint32_t)stream_id, actions, q_out);
    return (int)cudaGetLastError();
}

// this device's copy of g_head_diff_staging, resolved once per device (no runtime query inside a stream capture later)
static float4 *head_staging_of_current_device() {
    static std::mutex mu;
    static float4 *by_device[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lock(mu);
    if (!by_device[dev] && cudaGetSymbolAddress(reinterpret_cast<void **>(&by_device[dev]), g_head_diff_staging) != cudaSuccess)
        by_device[dev] = nullptr;
    return by_device[dev];
}

// slot of c_head_diff for launches on `stream` of the current device; -1 when all slots are taken by other streams or
// PP_CONST_HEADS=0 asks for the shared-memory table
static int head_slot_for(cudaStream_t stream) {
    static std::mutex mu;
    static struct { int device; cudaStream_t stream; } owner[HEAD_SLOTS];
    static int used = 0;
    static const bool off = [] { const char *e = getenv("PP_CONST_HEADS"); return e && e[0] == '0'; }();
    if (off) return -1;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    std::lock_guard<std::mutex> lock(mu);
    for (int i = 0; i < used; ++i)
        if (owner[i].device == dev && owner[i].stream == stream) return i;
    if (used == HEAD_SLOTS) return -1;
    owner[used] = {dev, stream};
    return used++;
}

int selfplay_tc_launch(int mode, int64_t n, int64_t k, const PPParams &p, const PPEnvState &st, const PPPolicy &pa,
                       const PPPolicy &pb, uint64_t seed, int64_t step_base, const PPServeSource &src, int32_t quota,
                       int64_t env_id_base, const PPRolloutOut &out, const PPReplayRing *ring, cudaStream_t stream) {
    constexpr int smem = (int)FusedMap::TOTAL;
    PPReplayRing r{};
    if (ring) r = *ring;
    const int64_t total_warps = (n + 31) / 32;
    const int64_t slots = (int64_t)sm_count();                                   // resident CTAs: 1 per SM
    // chunks of <= 16 warps; a whole number of rounds over the resident CTAs once there is more than one round
    int64_t n_chunks = (total_warps + CTA_GROUPS - 1) / CTA_GROUPS;              // small n: about 1 warp per group
    if (n_chunks > slots) {
        const int64_t rounds = (total_warps + slots * 16 - 1) / (slots * 16);
        n_chunks = rounds * slots;
    }
    const unsigned blocks = (unsigned)(n_chunks < slots ? n_chunks : slots);
    const bool preclaim = src.kind == PP_SERVE_QUEUE && src.pool_vx == nullptr;   // a Philox-drawn serve queue (host-buffer evaluation)
    // the head tables of this launch: built on the device, copied into this (device, stream)'s slot of constant memory
    const bool qa = pa.kind == PP_POLICY_QNET, qb = pb.kind == PP_POLICY_QNET;
    int cslot = (qa || qb) ? head_slot_for(stream) : -1;
    if (cslot >= 0) {
        cudaError_t err;
        float4 *mine = head_staging_of_current_device();                         // nullptr: could not be resolved
        if (!mine) return (int)cudaErrorInvalidSymbol;
        head_diff_kernel<<<1, 128, 0, stream>>>(qa ? pa.weights : nullptr, qb ? pb.weights : nullptr, cslot);
        constexpr size_t bytes = 2 * HEAD_ENTRIES * sizeof(float4);
        err = cudaMemcpyToSymbolAsync(c_head_diff, mine + (size_t)cslot * 2 * HEAD_ENTRIES, bytes, (size_t)cslot * bytes,
                                      cudaMemcpyDeviceToDevice, stream);
        if (err != cudaSuccess) return (int)err;
    }
    auto launch = [&](auto kernel) -> int {
        cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (err != cudaSuccess) return (int)err;
        kernel<<<blocks, TC_FUSED_THREADS, smem, stream>>>(p, st, n, k, pa, pb, seed, step_base, src, quota, env_id_base, out, r,
                                                        n_chunks, cslot);
        return (int)cudaGetLastError();
    };
    if (mode == PP_MODE_F64) {
        if (cslot >= 0) return preclaim ? launch(selfplay_tc_kernel<double, true, true>) : launch(selfplay_tc_kernel<double, false, true>);
        return preclaim ? launch(selfplay_tc_kernel<double, true, false>) : launch(selfplay_tc_kernel<double, false, false>);
    }
    if (cslot >= 0) return preclaim ? launch(selfplay_tc_kernel<float, true, true>) : launch(selfplay_tc_kernel<float, false, true>);
    return preclaim ? launch(selfplay_tc_kernel<float, true, false>) : launch(selfplay_tc_kernel<float, false, false>);
}

}  // namespace pp
