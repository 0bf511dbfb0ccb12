// lstm_tc_kernels.cu — K2b on the tensor cores: one QNetRNN step (models/qnet_rnn.py:107-144, seq_len 1) for a
// 128-env tile per CTA, every layer a tcgen05.mma batch with fp32 accumulation in TMEM (PP_PREC_F16).
//
// Roles: 8 compute warps (two threads per env = TMEM lane) + 1 issuer warp whose elected lane issues every MMA + 1 producer
// warp that streams the weight image PP_RNNTC_* from L2 through a 4-slot shared-memory ring with TMA bulk copies (in the
// paired form two CTAs of a cluster load half a stage each and multicast it).  The sides meet only at mbarriers:  ready
// (256 arrivals: "my operand rows are written"), done[2] (tcgen05.commit: "accumulator buffer b is complete"), dfree[2]
// (256 arrivals: "buffer b has been drained"), full / empty per ring slot.
//
// Tensor memory (all 512 columns of the SM):
//     [  0,128) A_hi   gate operand [f2 | h_prev] (fp16 pairs), later the shared-head output s      (K = 256)
//     [128,256) A_lo   low halves of the same
//     [256,384) D0     accumulator buffer 0          [384,512) D1   accumulator buffer 1 (its first 64 columns also
//                                                                    hold F1 hi/lo, the operand of features.2)
// Per player-step:  L1 (X*W1 -> D0) -> F1 -> features.2 (-> D0, N = 128) -> f2 into A | h_prev staged into A ->
// gates in four quarters of 32 units (N = 128 = 4 gates x 32, K = 256), double-buffered D0/D1 so the LSTM cell of
// quarter q runs on the CUDA cores while the tensor core computes quarter q+1 -> h_new as an fp16 hi/lo A tile in
// shared memory -> shared head (-> D0) -> s into A -> dueling heads (-> D1[0..15]) -> Q.
// Activations and weights are split hi/lo everywhere and every product is taken as hi*hi + lo*hi + hi*lo (with fp16-
// rounded gate weights alone the reference's trained checkpoint drifts to 1.3e-3 relative after 12 carried steps).
// The 640 KB weight image is what L2 has to deliver to every SM on every player-step.  sigmoid / tanh use ex2.approx +
// rcp.approx (error ~1e-6, far inside the 1e-3 budget).
#include <cstdio>
#include <cstdlib>
#include "pp_host.h"
#include "pp_rollout.cuh"
#include "tc_tiles.cuh"

namespace pp {

#ifdef PP_TC_TIMING
// debug build: cycles per phase, summed over the launch.  issuer: 0 player-steps 1 wait_ready 2 fill(empty) 3 full wait
// 4 wait_dfree 5 total;  worker (thread 0): 8 wait_done 9 h staging 10 cell 12 total 13 env step 14 cell ld 15 cell math
// 16.. wait_done by hand-off (L1, features.2, quarters 0..3, shared head, dueling heads).  Two more experiment switches
// (wrong results, timing only): PP_RT_NOCELL drops the cell arithmetic, PP_RT_HALFBYTES fetches half of every weight stage.
__device__ unsigned long long g_rt_timing[32];
#define RT_T0(v) long long v = clock64()
#define RT_ADD(slot, since) atomicAdd(&g_rt_timing[slot], (unsigned long long)(clock64() - (since)))
#else
#define RT_T0(v)
#define RT_ADD(slot, since)
#endif

namespace {

// RT_HALVES threads share one env row (same TMEM lane, different columns): more warps per scheduler for the epilogues
constexpr int RT_ROWS = 128, RT_HALVES = 2, RT_WORKERS = RT_ROWS * RT_HALVES, RT_THREADS = RT_WORKERS + 64,
              RT_ISSUER_WARP = RT_WORKERS / 32, RT_PRODUCER_WARP = RT_ISSUER_WARP + 1;
// The image's 40 tiles are fetched as 21 STAGES of up to two weight tiles (36 KB): every stage hand-off (a try_wait, a commit)
// idles the tensor pipe for ~85 cycles, so fewer, larger stages (tools/mma_rate.cu: 74 cycles per MMA in stages of 8, 88 in
// stages of 4, 64 back to back).
constexpr uint32_t RT_SLOTS = 4, RT_SLOT = 2 * PP_RNNTC_TILE + PP_RNNTC_BIAS;
constexpr int RT_STAGES = 21;
constexpr uint32_t SM_RING = 0, SM_HNEW = SM_RING + RT_SLOTS * RT_SLOT, SM_HNEW_LO = SM_HNEW + 32768,
                   SM_X = SM_HNEW + 65536, SM_CTRL = SM_X + 4096, SM_FLAGS = SM_CTRL + 128, SM_TOTAL = SM_FLAGS + 2 * RT_ROWS;
// control block: full[] empty[] ready done[2] dfree[2] step (mbarriers), TMEM base, stop flag; then per-row flags
constexpr uint32_t B_FULL = 0, B_EMPTY = RT_SLOTS, B_READY = 2 * RT_SLOTS, B_DONE = B_READY + 1, B_DFREE = B_DONE + 2,
                   B_STEP = B_DFREE + 2, CTRL_TMEM = (B_STEP + 1) * 8, CTRL_STOP = CTRL_TMEM + 4;
static_assert(CTRL_STOP + 4 <= 128, "control block");
// TMEM columns
constexpr uint32_t T_AHI = 0, T_ALO = 128, T_D0 = 256, T_D1 = 384, T_FHI = 384, T_FLO = 416;

__device__ __forceinline__ void stage_info(int i, uint32_t &off, uint32_t &bytes) {
    constexpr uint32_t PAIR_B = 2 * PP_RNNTC_TILE + PP_RNNTC_BIAS, PAIR_T = 2 * PP_RNNTC_TILE;   // tile + bias + tile / two tiles
    if (i == 0) { off = PP_RNNTC_S0; bytes = PP_RNNTC_S0_BYTES; }                      // L1
    else if (i == 1) { off = PP_RNNTC_S1; bytes = PP_RNNTC_S1_BYTES + PP_RNNTC_S2_BYTES; }   // features.2: hi tile + bias, lo tile
    else if (i < 18) {                            // per quarter: hi (c 0, 1 with the bias tile between), hi (c 2, 3), lo (0, 1), lo (2, 3)
        const int q = (i - 2) >> 2, j = (i - 2) & 3;
        off = PP_RNNTC_G + q * PP_RNNTC_GQ_BYTES + (j == 0 ? 0 : PAIR_B + (j - 1) * PAIR_T);
        bytes = j == 0 ? PAIR_B : PAIR_T;
    } else if (i < 20) {                          // shared head: hi (c 0 + bias, c 1), lo (c 0, 1)
        off = PP_RNNTC_WS + (i == 18 ? 0 : PAIR_B);
        bytes = i == 18 ? PAIR_B : PAIR_T;
    } else { off = PP_RNNTC_HD; bytes = PP_RNNTC_HD_BYTES; }
}
static_assert(PP_RNNTC_S1 + PP_RNNTC_S1_BYTES == PP_RNNTC_S2 && PP_RNNTC_S1_BYTES == PP_RNNTC_TILE + PP_RNNTC_BIAS, "image layout");
static_assert(PP_RNNTC_GQ_BYTES == 8 * PP_RNNTC_TILE + PP_RNNTC_BIAS && PP_RNNTC_WS_BYTES == 4 * PP_RNNTC_TILE + PP_RNNTC_BIAS, "image layout");

// ------------------------------------------------------------------------------------------ issuer side
struct Issuer {
    uint8_t *smem;
    uint64_t *bars;
    uint32_t tm;                   // TMEM base
    int c_slot, c_round;           // ring slot to consume next and its round parity
    uint32_t ready_par, dfree_par;
    bool leader;                   // ALL lanes of the issuer warp run the program (uniform control flow keeps counters and
                                   // descriptors in uniform registers); only the leader lane executes the TMA / MMA / commit
    bool pair;                     // the CTA shares its weight stream with its cluster peer: a slot is free once BOTH have read it

    __device__ __forceinline__ uint32_t acquire() {     // shared-memory address of the next stage, landed (Producer sent it)
        { RT_T0(t_); tc::mbar_wait(bars + B_FULL + c_slot, (uint32_t)c_round); if (leader) RT_ADD(3, t_); }
        return tc::smem_u32(smem + SM_RING + c_slot * RT_SLOT);
    }
    __device__ __forceinline__ void release() {         // the slot is free once the MMAs issued so far have read it
        if (leader) {
            if (pair) tc::umma_commit_multicast(bars + B_EMPTY + c_slot, 3);
            else tc::umma_commit(bars + B_EMPTY + c_slot);
        }
        if (++c_slot == (int)RT_SLOTS) { c_slot = 0; c_round ^= 1; }
    }
    __device__ __forceinline__ void wait_ready() {
        { RT_T0(t_); tc::mbar_wait(bars + B_READY, ready_par); if (leader) RT_ADD(1, t_); }
        ready_par ^= 1u;
        tc::tc_fence_after();
    }
    __device__ __forceinline__ void wait_dfree(int b) {
        { RT_T0(t_); tc::mbar_wait(bars + B_DFREE + b, (dfree_par >> b) & 1u); if (leader) RT_ADD(4, t_); }
        dfree_par ^= 1u << b;
        tc::tc_fence_after();
    }
    __device__ __forceinline__ void done(int b) { if (leader) tc::umma_commit(bars + B_DONE + b); }
};

// The producer warp: streams the weight image of every player-step, stage by stage, into the ring (TMA bulk copies).
// It only ever blocks on a slot's `empty` barrier, i.e. on the MMAs that still read that slot.
struct Producer {
    uint8_t *smem;
    uint64_t *bars;
    int slot, round;
    bool leader;
    int pair_rank;                 // < 0: alone; 0 / 1: this CTA loads that HALF of every stage and multicasts it to both CTAs of
                                   // the cluster (each SM then pulls half the image from L2 per player-step)
    __device__ __forceinline__ void player_step(const uint8_t *img) {
#pragma unroll 1
        for (int st = 0; st < RT_STAGES; ++st) {
            tc::mbar_wait(bars + B_EMPTY + slot, (uint32_t)(round ^ 1));       // first round passes at once
            uint32_t off, bytes;
            stage_info(st, off, bytes);
#ifdef PP_RT_HALFBYTES
            bytes >>= 1;
#endif
            if (leader) {
                tc::mbar_expect_tx(bars + B_FULL + slot, bytes);        // both halves land on this barrier
                if (pair_rank < 0) {
                    tc::tma_bulk_g2s(smem + SM_RING + slot * RT_SLOT, img + off, bytes, bars + B_FULL + slot);
                } else {
                    const uint32_t half = bytes >> 1, at = (uint32_t)pair_rank * half;
                    tc::tma_bulk_g2s_multicast(smem + SM_RING + slot * RT_SLOT + at, img + off + at, half, bars + B_FULL + slot, 3);
                }
            }
            if (++slot == (int)RT_SLOTS) { slot = 0; round ^= 1; }
        }
    }
};

// MMA issue helpers.  Descriptors are built once per operand and ADVANCED by constants (one uniform add per MMA): the
// issue loop must run well ahead of the 64-cycle-per-MMA tensor pipe.
template <int N> __device__ __forceinline__ uint64_t bdesc(uint32_t b_sm) { return tc::smem_desc(b_sm, N * 16, SBO); }
__device__ __forceinline__ uint64_t adesc(uint32_t a_sm) { return tc::smem_desc(a_sm, A_LBO, SBO); }
constexpr uint64_t A_KSTEP = (2 * A_LBO) >> 4;                    // descriptor increment per K = 16 step of a 128-row A tile

// COUNT K-steps: D (+)= A[tmem at a_tm + 8 j] * B[bd + j * kstep].  ONE branch on the leader predicate around the whole
// run (a branch per MMA costs more than the MMA's issue slot); the operands are warp-uniform values computed by all lanes.
template <int N, int COUNT> __device__ __forceinline__ void mma_ts_run(bool leader, uint32_t d, uint32_t a_tm, uint64_t bd, bool acc_first) {
    if (leader) {
#pragma unroll
        for (int j = 0; j < COUNT; ++j)
            tc::umma_f16_ts(d, a_tm + 8 * j, bd + (uint64_t)j * ((2 * N * 16) >> 4), tc::idesc_f16(128, N), acc_first || j > 0);
    }
}
template <int N, int COUNT> __device__ __forceinline__ void mma_ss_run(bool leader, uint32_t d, uint64_t ad, uint64_t bd, bool acc_first) {
    if (leader) {
#pragma unroll
        for (int j = 0; j < COUNT; ++j)
            tc::umma_f16(d, ad + (uint64_t)j * A_KSTEP, bd + (uint64_t)j * ((2 * N * 16) >> 4), tc::idesc_f16(128, N), acc_first || j > 0);
    }
}

// one player-step of MMA work (called by ALL lanes of the issuer warp; see Issuer::leader)
__device__ __forceinline__ void issue_player_step(Issuer &is) {
    const bool ld = is.leader;
    const uint32_t tm = is.tm;
    const uint64_t x = adesc(tc::smem_u32(is.smem + SM_X));
    const uint64_t hh = adesc(tc::smem_u32(is.smem + SM_HNEW)), hl = adesc(tc::smem_u32(is.smem + SM_HNEW_LO));
    uint32_t a;
    // ---- L1: D0[0..63] = X * W1h' + X * W1l'
    is.wait_ready();
    a = is.acquire();
    mma_ss_run<64, 1>(ld, tm + T_D0, x, bdesc<64>(a), false);
    mma_ss_run<64, 1>(ld, tm + T_D0, x, bdesc<64>(a + 2048), true);
    is.release();
    is.done(0);
    // ---- features.2: D0[0..127] = F1h*Wh + F1l*Wh + X*B + F1h*Wl
    is.wait_ready();
    a = is.acquire();
    mma_ts_run<128, 4>(ld, tm + T_D0, tm + T_FHI, bdesc<128>(a), false);
    mma_ts_run<128, 4>(ld, tm + T_D0, tm + T_FLO, bdesc<128>(a), true);
    mma_ss_run<128, 1>(ld, tm + T_D0, x, bdesc<128>(a + PP_RNNTC_TILE), true);
    mma_ts_run<128, 4>(ld, tm + T_D0, tm + T_FHI, bdesc<128>(a + PP_RNNTC_TILE + PP_RNNTC_BIAS), true);
    is.release();
    is.done(0);
    // ---- gates, four quarters of 32 units, K = 256 in four stages of 64 (hi weights), then four more (lo weights)
    is.wait_ready();
#pragma unroll 1
    for (int q = 0; q < 4; ++q) {
        if (q >= 2) is.wait_dfree(q & 1);
        const uint32_t d = tm + ((q & 1) ? T_D1 : T_D0);
#pragma unroll 1
        for (int c = 0; c < 4; c += 2) {                // two K chunks of 64 per stage
            a = is.acquire();
            const uint64_t bd0 = bdesc<128>(a), bd1 = bdesc<128>(a + PP_RNNTC_TILE + (c == 0 ? PP_RNNTC_BIAS : 0));
            mma_ts_run<128, 4>(ld, d, tm + T_AHI + c * 32, bd0, c != 0);
            mma_ts_run<128, 4>(ld, d, tm + T_ALO + c * 32, bd0, true);
            if (c == 0) mma_ss_run<128, 1>(ld, d, x, bdesc<128>(a + PP_RNNTC_TILE), true);
            mma_ts_run<128, 4>(ld, d, tm + T_AHI + (c + 1) * 32, bd1, true);
            mma_ts_run<128, 4>(ld, d, tm + T_ALO + (c + 1) * 32, bd1, true);
            is.release();
        }
#pragma unroll 1
        for (int c = 0; c < 4; c += 2) {                // A_hi * W_lo
            a = is.acquire();
            mma_ts_run<128, 4>(ld, d, tm + T_AHI + c * 32, bdesc<128>(a), true);
            mma_ts_run<128, 4>(ld, d, tm + T_AHI + (c + 1) * 32, bdesc<128>(a + PP_RNNTC_TILE), true);
            is.release();
        }
        is.done(q & 1);
    }
    // ---- shared head: D0[0..127] = Hh*Wsh + Hl*Wsh + X*B + Hh*Wsl   (A = h_new tile in shared memory, K = 128)
    is.wait_ready();
    {
        a = is.acquire();
        const uint64_t bd0 = bdesc<128>(a), bd1 = bdesc<128>(a + PP_RNNTC_TILE + PP_RNNTC_BIAS);
        mma_ss_run<128, 4>(ld, tm + T_D0, hh, bd0, false);
        mma_ss_run<128, 4>(ld, tm + T_D0, hl, bd0, true);
        mma_ss_run<128, 1>(ld, tm + T_D0, x, bdesc<128>(a + PP_RNNTC_TILE), true);
        mma_ss_run<128, 4>(ld, tm + T_D0, hh + 4 * A_KSTEP, bd1, true);
        mma_ss_run<128, 4>(ld, tm + T_D0, hl + 4 * A_KSTEP, bd1, true);
        is.release();
        a = is.acquire();
        mma_ss_run<128, 4>(ld, tm + T_D0, hh, bdesc<128>(a), true);
        mma_ss_run<128, 4>(ld, tm + T_D0, hh + 4 * A_KSTEP, bdesc<128>(a + PP_RNNTC_TILE), true);
        is.release();
    }
    is.done(0);
    // ---- dueling heads: D1[0..15] = Sh*Whh + Sl*Whh + Sh*Whl + X*B     (K = 128, N = 16)
    is.wait_ready();
    a = is.acquire();
    mma_ts_run<16, 8>(ld, tm + T_D1, tm + T_AHI, bdesc<16>(a), false);
    mma_ts_run<16, 8>(ld, tm + T_D1, tm + T_ALO, bdesc<16>(a), true);
    mma_ts_run<16, 8>(ld, tm + T_D1, tm + T_AHI, bdesc<16>(a + 4096), true);
    mma_ss_run<16, 1>(ld, tm + T_D1, x, bdesc<16>(a + 8192), true);
    is.release();
    is.done(1);
}

// ------------------------------------------------------------------------------------------ compute side
struct Worker {
    uint8_t *smem;
    uint64_t *bars;
    uint32_t tm;             // TMEM base + this warp's lane offset
    uint32_t done_par;       // bit b = parity of done[b]
    int row, half;           // env row (= TMEM lane) and which share of its columns this thread works on

    __device__ __forceinline__ void publish(bool smem_rows) {       // my operand rows are written
        if (smem_rows) tc::fence_proxy_async();
        tc::tc_fence_before();
        tc::mbar_arrive(bars + B_READY);
    }
    __device__ __forceinline__ void wait_done(int b, int site = 0) {     // site: which hand-off (phase timers only)
        { RT_T0(t_); tc::mbar_wait(bars + B_DONE + b, (done_par >> b) & 1u); if (row == 0 && half == 0) { RT_ADD(8, t_); RT_ADD(16 + site, t_); } }
        done_par ^= 1u << b;
        tc::tc_fence_after();
    }
    __device__ __forceinline__ void drained(int b) {
        tc::tc_fence_before();
        tc::mbar_arrive(bars + B_DFREE + b);
    }
};

// NCOLS accumulator columns at src -> ReLU -> hi/lo fp16 pairs -> TMEM at dst_hi / dst_lo (NCOLS / 2 columns each)
template <int NCOLS> __device__ __forceinline__ void relu_split_to_tmem(uint32_t src, uint32_t dst_hi, uint32_t dst_lo, int half) {
#pragma unroll 1
    for (int blk = half; blk < NCOLS / 32; blk += RT_HALVES) {
        uint32_t r0[16], r1[16], hi[16], lo[16];
        tc::tmem_ld16(src + blk * 32, r0);
        tc::tmem_ld16(src + blk * 32 + 16, r1);
        tc::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float a = __uint_as_float(r0[2 * j]), b = __uint_as_float(r0[2 * j + 1]);
            hi[j] = tc::pack_f16x2_rz_relu(a, b);
            lo[j] = tc::pack_f16x2<true>(a - h2f(hi[j], 0), b - h2f(hi[j], 1));
            const float c = __uint_as_float(r1[2 * j]), d = __uint_as_float(r1[2 * j + 1]);
            hi[8 + j] = tc::pack_f16x2_rz_relu(c, d);
            lo[8 + j] = tc::pack_f16x2<true>(c - h2f(hi[8 + j], 0), d - h2f(hi[8 + j], 1));
        }
        tc::tmem_st16(dst_hi + blk * 16, hi);
        tc::tmem_st16(dst_lo + blk * 16, lo);
    }
    tc::tmem_st_wait();
}

// LSTM cell of one unit from its four gate pre-activations (models/qnet_rnn.py:130, torch gate order i, f, g, o):
// c' = s(f) c + s(i) tanh(g), h' = s(o) tanh(c').  34 instructions, 8 of them MUFU:
//   * ex2.approx.ftz / rcp.approx.ftz directly (one MUFU each): __expf and __fdividef wrap them in denormal / range fix-ups
//     (5 - 6 instructions per call, 45 of the old cell's ~85) that a logistic function never needs;
//   * with a_x = 1 + e^-x the logistic terms share reciprocals PAIRWISE: s(i) = a_f / (a_i a_f), s(f) = a_i / (a_i a_f) and
//     s(o), 1 / a_g likewise (2 MUFU.RCP + 6 FMUL instead of 4 MUFU.RCP: the cell is MUFU-bound once it is this short);
//   * the exponent is clamped from above only (e^-x <= 2^60: a product of two terms stays finite; e^-x -> 0 needs no guard).
// ex2.approx: 2^-22 relative, rcp.approx: 1 ulp — the cell stays ~1e-6 from the fp32 reference (budget 1e-3).
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ void lstm_cell(float gi, float gf, float gg, float go, float c_prev, float &c_new, float &h_new) {
#ifdef PP_RT_NOCELL
    c_new = gi + gf + c_prev; h_new = gg + go; return;
#endif
    constexpr float NL2E = -1.4426950408889634f;                       // e^-x = 2^(-x log2 e)
    auto a1 = [](float scaled) { return 1.0f + ex2_approx(fminf(scaled, 60.0f)); };
    const float ai = a1(gi * NL2E), af = a1(gf * NL2E), ao = a1(go * NL2E), ag = a1(gg * (2.0f * NL2E));
    const float r1 = rcp_approx(ai * af), r2 = rcp_approx(ao * ag);
    const float si = r1 * af, sf = r1 * ai, so = r2 * ag;
    const float tg = fmaf(2.0f, r2 * ao, -1.0f);                       // tanh(g) = 2 s(2 g) - 1
    c_new = fmaf(sf, c_prev, si * tg);
    const float tc_ = fmaf(2.0f, rcp_approx(a1(c_new * (2.0f * NL2E))), -1.0f);
    h_new = so * tc_;
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
}

// One player-step for this thread's env.  obs: its observation; fresh: episode start ((h, c) = 0 before the step);
// gh / gc: the player's (h, c) in global memory, fp32, blocked by warp (see below); live: the env exists and is not frozen.
__device__ __forceinline__ void compute_player_step(Worker &w, const float (&obs)[7], bool fresh, bool live, float *__restrict__ gh,
                                                    float *__restrict__ gc, int64_t env, float (&q)[3]) {
    const uint32_t tm = w.tm;
    const int row = w.row;
    const bool carry = live && !fresh;
    // (h, c) of the tensor-core path are BLOCKED by warp: [n / 32][32 unit groups of 4][32 envs][4 floats].  A thread still
    // walks its own env's 128 units, but the 32 lanes of a warp now touch one contiguous 512-byte run per access instead of
    // 32 rows 512 bytes apart: 8x fewer memory transactions for the 2 KB of recurrent state per env-step (the env-major
    // [n][128] layout cost ~8 000 of the 36 000 cycles of a player-step in the load / store unit).
    const size_t blk4 = ((size_t)env >> 5) * (32 * 32) + ((size_t)env & 31);           // in float4 units; + 32 per unit group
    auto at4 = [blk4](int j4) { return blk4 + (size_t)j4 * 32; };
    const float4 *h4 = reinterpret_cast<const float4 *>(gh);
    float4 *hs4 = reinterpret_cast<float4 *>(gh), *cs4 = reinterpret_cast<float4 *>(gc);
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const int half = w.half;
    constexpr int SUBS = 4 / RT_HALVES;
    // ---- X row; h_prev -> A (columns 64..127 of the hi / lo halves)
    if (half == 0) write_x_row(w.smem + SM_X, row, obs);
    RT_T0(th_);
    float4 hv_[4 / RT_HALVES][8];                        // ALL of this thread's h_prev loads are in flight before the first use
#pragma unroll
    for (int bi = 0; bi < 4 / RT_HALVES; ++bi)
#pragma unroll
        for (int j = 0; j < 8; ++j) hv_[bi][j] = carry ? h4[at4((half + bi * RT_HALVES) * 8 + j)] : zero4;
#pragma unroll
    for (int bi = 0; bi < 4 / RT_HALVES; ++bi) {        // 32 units per block -> 16 packed pairs
        const int blk = half + bi * RT_HALVES;
        const float4 (&v)[8] = hv_[bi];
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            hi[2 * j] = tc::pack_f16x2<false>(v[j].x, v[j].y);
            lo[2 * j] = tc::pack_f16x2<false>(v[j].x - h2f(hi[2 * j], 0), v[j].y - h2f(hi[2 * j], 1));
            hi[2 * j + 1] = tc::pack_f16x2<false>(v[j].z, v[j].w);
            lo[2 * j + 1] = tc::pack_f16x2<false>(v[j].z - h2f(hi[2 * j + 1], 0), v[j].w - h2f(hi[2 * j + 1], 1));
        }
        tc::tmem_st16(tm + T_AHI + 64 + blk * 16, hi);
        tc::tmem_st16(tm + T_ALO + 64 + blk * 16, lo);
    }
    tc::tmem_st_wait();
    if (row == 0 && half == 0) RT_ADD(9, th_);
    w.publish(true);
    // ---- F1
    w.wait_done(0, 0);
    relu_split_to_tmem<64>(tm + T_D0, tm + T_FHI, tm + T_FLO, half);
    w.publish(false);
    // ---- f2 -> A columns 0..63
    w.wait_done(0, 1);
    relu_split_to_tmem<128>(tm + T_D0, tm + T_AHI, tm + T_ALO, half);
    w.publish(false);
    // ---- LSTM cell, quarter by quarter (i, f, g, o at columns 0, 32, 64, 96 of the buffer; 8 units per sub-block, the
    // sub-blocks of a quarter are shared out over the row's threads).
    // The quarter's c_prev values are requested BEFORE waiting for its gates, so the loads ride under the MMAs; the
    // quarter loop is not unrolled (the cell code is large and the kernel was instruction-cache bound when it was).
    uint8_t *hn = w.smem + SM_HNEW;
#pragma unroll 1
    for (int qt = 0; qt < 4; ++qt) {
        float4 cp[2 * SUBS];
#pragma unroll
        for (int sb = 0; sb < SUBS; ++sb) {
            const int b = half * SUBS + sb;
            cp[2 * sb] = carry ? cs4[at4(qt * 8 + 2 * b)] : zero4;
            cp[2 * sb + 1] = carry ? cs4[at4(qt * 8 + 2 * b + 1)] : zero4;
        }
        w.wait_done(qt & 1, 2 + qt);
        const uint32_t d = tm + ((qt & 1) ? T_D1 : T_D0);
        RT_T0(tc_);
#pragma unroll
        for (int sb = 0; sb < SUBS; ++sb) {
            const int b = half * SUBS + sb;
            uint32_t gi[8], gf[8], gg[8], go[8];
            RT_T0(tld_);
            tmem_ld8(d + 8 * b, gi);
            tmem_ld8(d + 32 + 8 * b, gf);
            tmem_ld8(d + 64 + 8 * b, gg);
            tmem_ld8(d + 96 + 8 * b, go);
            tc::tmem_ld_wait();
            if (row == 0 && half == 0) RT_ADD(14, tld_);
            RT_T0(tmath_);
            const float cprev[8] = {cp[2 * sb].x, cp[2 * sb].y, cp[2 * sb].z, cp[2 * sb].w,
                                    cp[2 * sb + 1].x, cp[2 * sb + 1].y, cp[2 * sb + 1].z, cp[2 * sb + 1].w};
            float hv[8], cv[8];
#pragma unroll
            for (int e = 0; e < 8; ++e)
                lstm_cell(__uint_as_float(gi[e]), __uint_as_float(gf[e]), __uint_as_float(gg[e]), __uint_as_float(go[e]),
                          cprev[e], cv[e], hv[e]);
            if (row == 0 && half == 0) RT_ADD(15, tmath_);
            if (live) {
                const int v4 = qt * 8 + 2 * b;
                cs4[at4(v4)] = make_float4(cv[0], cv[1], cv[2], cv[3]); cs4[at4(v4 + 1)] = make_float4(cv[4], cv[5], cv[6], cv[7]);
                hs4[at4(v4)] = make_float4(hv[0], hv[1], hv[2], hv[3]); hs4[at4(v4 + 1)] = make_float4(hv[4], hv[5], hv[6], hv[7]);
            }
            uint32_t ph[4], pl[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                ph[j] = tc::pack_f16x2<false>(hv[2 * j], hv[2 * j + 1]);
                pl[j] = tc::pack_f16x2<false>(hv[2 * j] - h2f(ph[j], 0), hv[2 * j + 1] - h2f(ph[j], 1));
            }
            const int chunk = qt * 4 + b;
            *reinterpret_cast<uint4 *>(hn + chunk * A_LBO + row * 16) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
            *reinterpret_cast<uint4 *>(hn + 32768 + chunk * A_LBO + row * 16) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
        }
        if (row == 0 && half == 0) RT_ADD(10, tc_);
        if (qt < 2) w.drained(qt & 1);
    }
    w.publish(true);
    // ---- shared head -> s into A columns 0..63
    w.wait_done(0, 6);
    relu_split_to_tmem<128>(tm + T_D0, tm + T_AHI, tm + T_ALO, half);
    w.publish(false);
    // ---- Q
    w.wait_done(1, 7);
    dueling_q(tm + T_D1, q);
}

__device__ __forceinline__ uint32_t rnn_tc_prologue(uint8_t *smem, int warp_id, bool pair = false) {
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + SM_CTRL);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + SM_CTRL + CTRL_TMEM);
    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < RT_SLOTS; ++s) { tc::mbar_init(bars + B_FULL + s, 1); tc::mbar_init(bars + B_EMPTY + s, pair ? 2 : 1); }
        tc::mbar_init(bars + B_READY, RT_WORKERS);
        tc::mbar_init(bars + B_DONE, 1); tc::mbar_init(bars + B_DONE + 1, 1);
        tc::mbar_init(bars + B_DFREE, RT_WORKERS); tc::mbar_init(bars + B_DFREE + 1, RT_WORKERS);
        tc::mbar_init(bars + B_STEP, RT_WORKERS);
        tc::fence_mbar_init();
    }
    if (warp_id == RT_ISSUER_WARP) tc::tmem_alloc<512>(tmem_slot);
    tc::tc_fence_before();
    __syncthreads();
    if (pair) tc::cluster_sync_all();                   // the peer's barriers exist before anything is multicast to them
    tc::tc_fence_after();
    return __shfl_sync(0xffffffffu, *tmem_slot, 0);
}

}  // namespace

// standalone: obs[n][7] -> Q -> action for one recurrent player, (h, c) carried in global memory
__global__ void __launch_bounds__(RT_THREADS, 1)
qnetrnn_act_tc_kernel(int64_t n, const float *__restrict__ obs, const PPPolicy pol, const uint8_t *__restrict__ reset_mask,
                      uint64_t seed, uint32_t step_index, int64_t env_id_base, uint32_t stream_id,
                      uint8_t *__restrict__ actions, float *__restrict__ q_out) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp_id = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const uint32_t tm = rnn_tc_prologue(smem, warp_id);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + SM_CTRL);
    const int64_t tiles = (n + RT_ROWS - 1) / RT_ROWS;
    int64_t my_tiles = 0;
    for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) ++my_tiles;
    if (warp_id == RT_ISSUER_WARP) {
        Issuer is{smem, bars, tm, 0, 0, 0u, 0u, tc::elect_one(), false};
        for (int64_t t = 0; t < my_tiles; ++t) {
            RT_T0(t_);
            issue_player_step(is);
            if (is.leader) RT_ADD(5, t_);
        }
        __syncwarp();
    } else if (warp_id == RT_PRODUCER_WARP) {
        Producer pr{smem, bars, 0, 0, tc::elect_one(), -1};
        for (int64_t t = 0; t < my_tiles; ++t) pr.player_step(reinterpret_cast<const uint8_t *>(pol.weights));
        __syncwarp();
    } else {
        const int row = threadIdx.x & (RT_ROWS - 1), half = threadIdx.x / RT_ROWS;
        Worker w{smem, bars, tm + ((uint32_t)((warp_id & 3) * 32) << 16), 0u, row, half};
        for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
            const int64_t i = t * RT_ROWS + row;
            const bool live = i < n;
            const int64_t ic = live ? i : 0;
            float o[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (live) {
#pragma unroll
                for (int k = 0; k < 7; ++k) o[k] = obs[i * 7 + k];
            }
            const bool fresh = live && reset_mask && reset_mask[i];
            float q[3];
            RT_T0(t_);
            compute_player_step(w, o, fresh, live, pol.h, pol.c, ic, q);
            if (threadIdx.x == 0) RT_ADD(12, t_);
            if (live && half == 0) {
                int a = argmax3(q);
                a = explore(a, pol.eps_threshold, seed, (uint32_t)(env_id_base + i), step_index, stream_id);
                actions[i] = (uint8_t)a;
                if (q_out) { q_out[i * 3 + 0] = q[0]; q_out[i * 3 + 1] = q[1]; q_out[i * 3 + 2] = q[2]; }
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp_id == RT_ISSUER_WARP) tc::tmem_dealloc<512>(tm);
}

// ------------------------------------------------------------------------------------------ fused self-play rollout
// k lock-step iterations of {obs, QNetRNN (or follower / random) A, B, env step, replay row, auto-reset} for tiles of
// up to 128 envs per CTA (one CTA per SM): the recurrent-player form of tests/arena.py:294-304 and of the rollout loop
// of scripts/train_rnn_iterative.py:732-780 on the tensor cores.  Env state lives in the registers of the row's first
// thread; (h, c) are env-major in global memory and zeroed at every episode start.  The launch is cut into chunks of at
// most 4 warps of envs, balanced to within one warp, a whole number of rounds over the SMs.
// PAIR: launched as clusters of two CTAs that run the same stage sequence (no quota, equal chunk counts) and share ONE weight
// stream: each loads half of every stage and multicasts it to both.
template <typename R, bool PAIR>
__global__ void __launch_bounds__(RT_THREADS, 1)
selfplay_rnn_tc_kernel(const PPParams params, const PPEnvState st, int64_t n, int64_t k_steps, const PPPolicy pol_a,
                       const PPPolicy pol_b, uint64_t seed, int64_t step_base, const PPServeSource src, int32_t quota,
                       int64_t env_id_base, const PPRolloutOut out, const PPReplayRing ring, int64_t n_chunks) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp_id = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const uint32_t tm = rnn_tc_prologue(smem, warp_id, PAIR);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + SM_CTRL);
    volatile uint32_t *stop_flag = reinterpret_cast<volatile uint32_t *>(smem + SM_CTRL + CTRL_STOP);
    volatile uint8_t *row_fresh = smem + SM_FLAGS, *row_live = smem + SM_FLAGS + RT_ROWS;
    const bool ra = pol_a.kind == PP_POLICY_QNETRNN, rb = pol_b.kind == PP_POLICY_QNETRNN;
    const int64_t total_warps = (n + 31) / 32;

    if (warp_id == RT_ISSUER_WARP || warp_id == RT_PRODUCER_WARP) {
        const bool producer = warp_id == RT_PRODUCER_WARP;                      // both follow the same step protocol
        Issuer is{smem, bars, tm, 0, 0, 0u, 0u, tc::elect_one(), PAIR};
        Producer pr{smem, bars, 0, 0, is.leader, PAIR ? (int)tc::cluster_cta_rank() : -1};
        uint32_t step_par = 0;
#pragma unroll 1
        for (int64_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
            if ((chunk + 1) * total_warps / n_chunks == chunk * total_warps / n_chunks) continue;     // empty chunk
#pragma unroll 1
            for (int64_t t = 0; t < k_steps; ++t) {
                tc::mbar_wait(bars + B_STEP, step_par);
                step_par ^= 1u;
                if (*stop_flag) break;                                         // the tile is frozen by the quota
                if (producer) {
                    if (ra) pr.player_step(reinterpret_cast<const uint8_t *>(pol_a.weights));
                    if (rb) pr.player_step(reinterpret_cast<const uint8_t *>(pol_b.weights));
                } else {
                    if (ra) issue_player_step(is);
                    if (rb) issue_player_step(is);
                }
            }
        }
        __syncwarp();
    } else {
        const int tid = threadIdx.x, row = tid & (RT_ROWS - 1), half = tid / RT_ROWS, lane = tid & 31, gw = row >> 5;
        Worker w{smem, bars, tm + ((uint32_t)((warp_id & 3) * 32) << 16), 0u, row, half};
        const EnvConsts<R> c(params);
        const StatePtrs<R> s(st);
        const int64_t ring_t0 = ring_first_step(ring, n, k_steps);
        Tally total;
#pragma unroll 1
        for (int64_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
            const int64_t w_lo = chunk * total_warps / n_chunks, w_hi = (chunk + 1) * total_warps / n_chunks;
            const int my_warps = (int)(w_hi - w_lo);                           // <= 4
            if (my_warps == 0) continue;
            const int64_t i = (w_lo + gw) * 32 + lane;
            const bool valid = gw < my_warps && i < n;
            const int64_t ic = valid ? i : 0;
            Lane<R> L;
            L.e = load_env<R>(s, ic);
            L.ep_idx = s.ep_idx[ic]; L.ep_len = s.ep_len[ic];
            const uint32_t gid = (uint32_t)(env_id_base + ic);
            bool fresh = L.ep_len == 0;                                        // episode start: (h, c) = 0 before the first step
#pragma unroll 1
            for (int64_t t = 0; t < k_steps; ++t) {
                const bool active = valid && !(quota > 0 && L.ep_idx >= quota);
                const uint32_t step = (uint32_t)(step_base + t);
                if (half == 0) { row_fresh[row] = fresh ? 1 : 0; row_live[row] = active ? 1 : 0; }
                const bool any = tc::bar_red_or(1, RT_WORKERS, active && half == 0);   // also publishes the row flags
                if (tid == 0) *stop_flag = any ? 0u : 1u;
                tc::mbar_arrive(bars + B_STEP);                                // release: the issuer reads the flag after this
                if (!any) break;
                const bool r_fresh = row_fresh[row] != 0, r_live = row_live[row] != 0;
                float oa[7], ob[7];
                observe<R>(L.e, oa, ob);
                int act_a = 1, act_b = 1;
#pragma unroll 1
                for (int p = 0; p < 2; ++p) {
                    const PPPolicy &pol = p ? pol_b : pol_a;
                    const uint32_t stream_id = p ? STREAM_ACT_B : STREAM_ACT_A;
                    int a = 1;
                    if (pol.kind == PP_POLICY_QNETRNN) {
                        float q[3];
                        RT_T0(tps_);
                        compute_player_step(w, p ? ob : oa, r_fresh, r_live, pol.h, pol.c, ic, q);
                        if (tid == 0) { RT_ADD(12, tps_); RT_ADD(0, clock64() - 1); }
                        a = explore(argmax3(q), pol.eps_threshold, seed, gid, step, stream_id);   // (h, c) advance even when exploring
                    } else if (pol.kind == PP_POLICY_RANDOM) {
                        a = random_action(seed, gid, step, stream_id);
                    } else if (pol.kind == PP_POLICY_QNET) {                   // fp32 on the CUDA cores of the env's own thread
                        if (half == 0 && gw < my_warps)
                            a = explore(qnet_greedy_global(pol.weights, p ? ob : oa), pol.eps_threshold, seed, gid, step, stream_id);
                    } else {
                        a = explore(follower_action(p ? ob : oa, pol.follower_tol), pol.eps_threshold, seed, gid, step, stream_id);
                    }
                    if (p) act_b = a; else act_a = a;
                }
                RT_T0(tenv_);
                if (half == 0 && gw < my_warps) {                              // whole warps: the bookkeeping collectives are safe
                    const int ep_before = L.ep_idx;
                    step_and_book<R>(c, L, active, act_a, act_b, ob, t, n, i, env_id_base, quota, out, ring,
                                     ring.head != nullptr && t >= ring_t0, src,
                                     [&](int ep, R &vx, R &vy, R &sp) { next_serve<R>(params, src, n, i, env_id_base, ep, vx, vy, sp); });
                    fresh = active ? (L.ep_idx != ep_before) : fresh;
                }
                if (tid == 0) RT_ADD(13, tenv_);
            }
            if (valid && half == 0) {
                store_env<R>(s, i, L.e);
                s.ep_idx[i] = L.ep_idx;
                s.ep_len[i] = L.ep_len;
            }
            if (half == 0) total.add(L.tally);
        }
        total.flush(out.counters, out.ep_log ? nullptr : out.ep_log_count);
    }
    tc::tc_fence_before();
    __syncthreads();
    if (PAIR) tc::cluster_sync_all();                   // no CTA leaves while its peer may still signal its barriers
    if (warp_id == RT_ISSUER_WARP) tc::tmem_dealloc<512>(tm);
}

#ifdef PP_TC_TIMING
extern "C" int pp_debug_rt_timing(unsigned long long *host32, int reset) {
    cudaDeviceSynchronize();
    cudaError_t e = cudaMemcpyFromSymbol(host32, g_rt_timing, sizeof(g_rt_timing));
    if (reset) { unsigned long long z[32] = {0}; cudaMemcpyToSymbol(g_rt_timing, z, sizeof z); }
    return (int)e;
}
#endif

static int rt_sm_count() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
            sms = 148;
    }
    return sms;
}

int qnetrnn_act_tc_launch(int64_t n, const float *obs, const PPPolicy &pol, const uint8_t *reset_mask, uint64_t seed,
                          int64_t step_index, int64_t env_id_base, int32_t stream_id, uint8_t *actions, float *q_out,
                          cudaStream_t stream) {
    cudaError_t err = cudaFuncSetAttribute(qnetrnn_act_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_TOTAL);
    if (err != cudaSuccess) return (int)err;
    const int64_t tiles = (n + RT_ROWS - 1) / RT_ROWS;
    const unsigned blocks = (unsigned)(tiles < rt_sm_count() ? tiles : rt_sm_count());
    qnetrnn_act_tc_kernel<<<blocks, RT_THREADS, SM_TOTAL, stream>>>(n, obs, pol, reset_mask, seed, (uint32_t)step_index,
                                                                    env_id_base, (uint32_t)stream_id, actions, q_out);
    return (int)cudaGetLastError();
}

// CTAs that can be co-resident as clusters of two (one CTA per SM; a GPC with an odd SM count leaves one out); 0 = no pairs
template <typename R> static int rt_pair_slots() {
    static int slots = -1;
    if (slots < 0) {
        slots = 0;
        const char *off = getenv("PP_RNN_PAIR");
        if (!(off && off[0] == '0') &&
            cudaFuncSetAttribute(selfplay_rnn_tc_kernel<R, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_TOTAL) == cudaSuccess) {
            cudaLaunchConfig_t cfg{};
            cudaLaunchAttribute at{};
            at.id = cudaLaunchAttributeClusterDimension;
            at.val.clusterDim.x = 2; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
            cfg.gridDim = dim3((unsigned)(rt_sm_count() & ~1)); cfg.blockDim = dim3(RT_THREADS); cfg.dynamicSmemBytes = SM_TOTAL;
            cfg.attrs = &at; cfg.numAttrs = 1;
            int clusters = 0;
            if (cudaOccupancyMaxActiveClusters(&clusters, selfplay_rnn_tc_kernel<R, true>, &cfg) == cudaSuccess) slots = 2 * clusters;
            else (void)cudaGetLastError();
        }
#ifdef PP_TC_TIMING
        printf("[rnn-tc] paired slots: %d of %d SMs\n", slots, rt_sm_count());
#endif
    }
    return slots;
}

template <typename R>
static int selfplay_rnn_tc_launch_t(int64_t n, int64_t k, const PPParams &p, const PPEnvState &st, const PPPolicy &pa, const PPPolicy &pb,
                                    uint64_t seed, int64_t step_base, const PPServeSource &src, int32_t quota, int64_t env_id_base,
                                    const PPRolloutOut &out, const PPReplayRing &r, cudaStream_t stream) {
    const int64_t total_warps = (n + 31) / 32;
    cudaError_t err;
    // paired form: every CTA must run the same number of full chunks (no quota freezes a tile, no ragged last round), and
    // leaving a few SMs unpaired must cost less than the shared stream gains
    const int64_t pslots = rt_pair_slots<R>();
    if (quota <= 0 && pslots >= rt_sm_count() - 8 && total_warps >= pslots * 4) {
        const int64_t n_chunks = ((total_warps + pslots * 4 - 1) / (pslots * 4)) * pslots;
        cudaLaunchConfig_t cfg{};
        cudaLaunchAttribute at{};
        at.id = cudaLaunchAttributeClusterDimension;
        at.val.clusterDim.x = 2; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
        cfg.gridDim = dim3((unsigned)pslots); cfg.blockDim = dim3(RT_THREADS); cfg.dynamicSmemBytes = SM_TOTAL; cfg.stream = stream;
        cfg.attrs = &at; cfg.numAttrs = 1;
        return (int)cudaLaunchKernelEx(&cfg, selfplay_rnn_tc_kernel<R, true>, p, st, n, k, pa, pb, seed, step_base, src, quota,
                                       env_id_base, out, r, n_chunks);
    }
    const int64_t slots = rt_sm_count();
    int64_t n_chunks = (total_warps + 3) / 4;                                   // tiles of <= 4 warps ...
    if (n_chunks > slots) n_chunks = ((total_warps + slots * 4 - 1) / (slots * 4)) * slots;   // ... a whole number of rounds
    const unsigned blocks = (unsigned)(n_chunks < slots ? n_chunks : slots);
    if ((err = cudaFuncSetAttribute(selfplay_rnn_tc_kernel<R, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_TOTAL)) != cudaSuccess) return (int)err;
    selfplay_rnn_tc_kernel<R, false><<<blocks, RT_THREADS, SM_TOTAL, stream>>>(p, st, n, k, pa, pb, seed, step_base, src, quota, env_id_base, out, r, n_chunks);
    return (int)cudaGetLastError();
}

int selfplay_rnn_tc_launch(int mode, int64_t n, int64_t k, const PPParams &p, const PPEnvState &st, const PPPolicy &pa,
                           const PPPolicy &pb, uint64_t seed, int64_t step_base, const PPServeSource &src, int32_t quota,
                           int64_t env_id_base, const PPRolloutOut &out, const PPReplayRing *ring, cudaStream_t stream) {
    PPReplayRing r{};
    if (ring) r = *ring;
    return mode == PP_MODE_F64 ? selfplay_rnn_tc_launch_t<double>(n, k, p, st, pa, pb, seed, step_base, src, quota, env_id_base, out, r, stream)
                               : selfplay_rnn_tc_launch_t<float>(n, k, p, st, pa, pb, seed, step_base, src, quota, env_id_base, out, r, stream);
}

}  // namespace pp
