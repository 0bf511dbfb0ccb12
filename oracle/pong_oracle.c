/*
 * pong_oracle.c — plain-C CPU restatement of the reference self-play hot path.
 *
 * TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference leg may load this library; it is the checker for the
 * CUDA kernels in pingpong_selfplay_ai_b200/csrc, never a product path.
 *
 * Parity status: PINNED — checked bit-for-bit against the unmodified Python reference in
 * the build container (tests/test_oracle_vs_reference.py) and against the committed
 * golden vectors under tests/golden/ that oracle/gen_golden.py generated from it.
 *
 * Restated reference code (all under /root/reference):
 *   step   envs/my_pong_env_2p.py:116-225     speed scaling  :227-232
 *   obs    envs/my_pong_env_2p.py:235-263     reset          :83-114 (serve injected)
 *   impact envs/physics.py:3-23
 *   QNet forward      models/qnet.py:43-50,71-75
 *   QNetRNN forward   models/qnet_rnn.py:107-144 (seq_len 1, carried h/c)
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -shared -fPIC (oracle/Makefile).  No FMA
 * contraction: CPython rounds every binary op separately and so must this file.  The
 * policy nets are the exception: they DEFINE their accumulation as an fmaf() chain in
 * ascending k starting from the bias, which is what the fp32 CUDA path executes, so
 * actions (argmax) can be compared exactly; that definition is itself checked against
 * torch within 1e-5 (tests/test_oracle_policy.py).
 *
 * The step is written once over REAL and instantiated for double (bit-exact mode) and
 * float (fast mode: identical op order in binary32).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

typedef struct OracleParams {       /* host-precomputed "the Python way" (oracle/pong_oracle.py) */
    double paddle_speed;            /* ps                                   */
    double half_width;              /* paddle_width / 2                     */
    double magnus_factor;
    double neg_e;                   /* -restitution                         */
    double m_1pe;                   /* m * (1 + e)                          */
    double inertia;                 /* (2/5) * m * R**2                     */
    double two_m_over_7;            /* 2*m/7.0                              */
    double mu;
    double mass;
    double radius;
    double speed_scale;             /* 1.0 + speed_increment                */
    int32_t enable_spin;
    int32_t max_score;
    int32_t speed_scale_every;
    int32_t pad_;
} OracleParams;

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)

#define REAL double
#define SUF f64
#include "pong_oracle_step.inc"
#undef REAL
#undef SUF

#define REAL float
#define SUF f32
#include "pong_oracle_step.inc"
#undef REAL
#undef SUF

/* ------------------------------------------------------------------ policy nets (fp32) */

static inline float relu(float v) { return v > 0.0f ? v : 0.0f; }

/* Function multiversioning: on a CPU with FMA3 the clone compiled for it turns fmaf() into one vfmadd instruction;
 * elsewhere the default clone calls libm's fmaf().  Both are the correctly rounded fused multiply-add: same bits. */
#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
#define ORACLE_CLONES __attribute__((target_clones("arch=x86-64-v3", "default"), optimize("no-tree-vectorize", "no-tree-slp-vectorize")))
#else
#define ORACLE_CLONES
#endif

/* y[j] = chain_j = init[j] (+ init2[j]); then acc = fmaf(W1[j][k], x1[k], acc) for k ascending, then the same over
 * (W2, x2) if W2 != NULL.  Eight outputs are advanced together: their chains are independent, so they overlap in
 * the FMA pipeline (one chain alone is bound by the FMA latency) — every output still sees exactly its own chain. */
ORACLE_CLONES
static void chains(const float *W1, const float *x1, int n1, const float *W2, const float *x2, int n2,
                   const float *init, const float *init2, float *y, int n_out, int do_relu)
{
    enum { B = 8 };
    for (int j0 = 0; j0 < n_out; j0 += B) {
        const int nb = n_out - j0 < B ? n_out - j0 : B;
        float acc[B];
        for (int jj = 0; jj < nb; ++jj) acc[jj] = init2 ? init[j0 + jj] + init2[j0 + jj] : init[j0 + jj];
        if (nb == B) {
            float a0 = acc[0], a1 = acc[1], a2 = acc[2], a3 = acc[3], a4 = acc[4], a5 = acc[5], a6 = acc[6], a7 = acc[7];
            for (int pass = 0; pass < (W2 ? 2 : 1); ++pass) {
                const int nk = pass ? n2 : n1;
                const float *xx = pass ? x2 : x1;
                const float *w0 = (pass ? W2 : W1) + (size_t)j0 * nk;
                const float *w1 = w0 + nk, *w2 = w1 + nk, *w3 = w2 + nk, *w4 = w3 + nk, *w5 = w4 + nk, *w6 = w5 + nk,
                            *w7 = w6 + nk;
                for (int k = 0; k < nk; ++k) {
                    const float xk = xx[k];
                    a0 = fmaf(w0[k], xk, a0); a1 = fmaf(w1[k], xk, a1); a2 = fmaf(w2[k], xk, a2); a3 = fmaf(w3[k], xk, a3);
                    a4 = fmaf(w4[k], xk, a4); a5 = fmaf(w5[k], xk, a5); a6 = fmaf(w6[k], xk, a6); a7 = fmaf(w7[k], xk, a7);
                }
            }
            acc[0] = a0; acc[1] = a1; acc[2] = a2; acc[3] = a3; acc[4] = a4; acc[5] = a5; acc[6] = a6; acc[7] = a7;
        } else {
            for (int jj = 0; jj < nb; ++jj) {
                for (int k = 0; k < n1; ++k) acc[jj] = fmaf(W1[(size_t)(j0 + jj) * n1 + k], x1[k], acc[jj]);
                if (W2) for (int k = 0; k < n2; ++k) acc[jj] = fmaf(W2[(size_t)(j0 + jj) * n2 + k], x2[k], acc[jj]);
            }
        }
        for (int jj = 0; jj < nb; ++jj) y[j0 + jj] = do_relu ? relu(acc[jj]) : acc[jj];
    }
}

/* dense layer, row-major W[out][in] as in a torch state_dict; acc = b; acc = fmaf(W[j][k], x[k], acc) */
static void dense(const float *W, const float *b, const float *x, float *y, int n_out, int n_in, int do_relu)
{
    chains(W, x, n_in, NULL, NULL, 0, b, NULL, y, n_out, do_relu);
}

/* dueling combine exactly as the CUDA path: mean = ((a0 + a1) + a2) / 3 ; q_i = v + (a_i - mean) */
static void dueling(float v, const float *a, float *q)
{
    float mean = ((a[0] + a[1]) + a[2]) / 3.0f;
    for (int i = 0; i < 3; ++i) q[i] = v + (a[i] - mean);
}

static int argmax3(const float *q)     /* first maximum wins, like torch.argmax on CPU */
{
    int best = 0;
    if (q[1] > q[best]) best = 1;
    if (q[2] > q[best]) best = 2;
    return best;
}

/* models/qnet.py:71-75.  Weights: effective (mu or mu + sigma*eps) matrices in state_dict layout. */
void oracle_qnet_forward(int64_t n, const float *obs /*[n,7]*/,
                         const float *W1, const float *b1,     /* [64,7]  [64] */
                         const float *W2, const float *b2,     /* [64,64] [64] */
                         const float *Wv, const float *bv,     /* [1,64]  [1]  */
                         const float *Wa, const float *ba,     /* [3,64]  [3]  */
                         float *q_out /*[n,3]*/, uint8_t *greedy /*[n] or NULL*/)
{
    for (int64_t i = 0; i < n; ++i) {
        float h1[64], h2[64], v, a[3], q[3];
        dense(W1, b1, obs + i * 7, h1, 64, 7, 1);
        dense(W2, b2, h1, h2, 64, 64, 1);
        dense(Wv, bv, h2, &v, 1, 64, 0);
        dense(Wa, ba, h2, a, 3, 64, 0);
        dueling(v, a, q);
        if (q_out) memcpy(q_out + i * 3, q, sizeof q);
        if (greedy) greedy[i] = (uint8_t)argmax3(q);
    }
}

static inline float sigmoidf_(float v) { return 1.0f / (1.0f + expf(-v)); }

/* models/qnet_rnn.py:107-144 with seq_len == 1.  Gate order i,f,g,o (torch nn.LSTM).  The gate
 * pre-activation is ONE fmaf chain over the concatenated input [feat(F) ; h(H)] starting from
 * (b_ih + b_hh): that is the CUDA kernel's definition; torch differs in the last ulps only. */
void oracle_qnetrnn_forward(int64_t n, int F, int H, int S /* head hidden, >0 */,
                            const float *obs /*[n,7]*/,
                            const float *Wf1, const float *bf1,   /* [F/2,7] */
                            const float *Wf2, const float *bf2,   /* [F,F/2] */
                            const float *Wih, const float *bih,   /* [4H,F]  */
                            const float *Whh, const float *bhh,   /* [4H,H]  */
                            const float *Ws, const float *bs,     /* [S,H]   */
                            const float *Wv, const float *bv,     /* [1,S]   */
                            const float *Wa, const float *ba,     /* [3,S]   */
                            float *h /*[n,H] in/out*/, float *c /*[n,H] in/out*/,
                            float *q_out /*[n,3]*/, uint8_t *greedy)
{
    enum { MAXD = 512 };
    float f1[MAXD], f2[MAXD], gates[4 * MAXD], hn[MAXD], s[MAXD];
    if (F > MAXD || H > MAXD || S > MAXD) return;
    for (int64_t e = 0; e < n; ++e) {
        float *he = h + e * H, *ce = c + e * H;
        dense(Wf1, bf1, obs + e * 7, f1, F / 2, 7, 1);
        dense(Wf2, bf2, f1, f2, F, F / 2, 1);
        /* gates[j] = (bih[j] + bhh[j]) then fmaf over Wih[j] . f2, then over Whh[j] . h */
        chains(Wih, f2, F, Whh, he, H, bih, bhh, gates, 4 * H, 0);
        for (int j = 0; j < H; ++j) {
            float ig = sigmoidf_(gates[j]);
            float fg = sigmoidf_(gates[H + j]);
            float gg = tanhf(gates[2 * H + j]);
            float og = sigmoidf_(gates[3 * H + j]);
            float cn = fg * ce[j] + ig * gg;
            ce[j] = cn;
            hn[j] = og * tanhf(cn);
        }
        memcpy(he, hn, sizeof(float) * H);
        dense(Ws, bs, hn, s, S, H, 1);
        float v, a[3], q[3];
        dense(Wv, bv, s, &v, 1, S, 0);
        dense(Wa, ba, s, a, 3, S, 0);
        dueling(v, a, q);
        if (q_out) memcpy(q_out + e * 3, q, sizeof q);
        if (greedy) greedy[e] = (uint8_t)argmax3(q);
    }
}


/* ------------------------------------------------------------------ device RNG restated */

/* Philox4x32-10 (Salmon et al., SC'11) — the counter-based generator the CUDA kernels use for
 * epsilon-greedy exploration, uniform-random players and (throughput mode) serves.
 * counter = {global env id, index (lock-step step or episode), stream, sub}, key = {seed lo, hi}. */
void oracle_philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                       uint32_t *out4)
{
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out4[0] = c0; out4[1] = c1; out4[2] = c2; out4[3] = c3;
}

enum { STREAM_SERVE = 0, STREAM_ACT_A = 1, STREAM_ACT_B = 2 };
enum { POLICY_QNET = 0, POLICY_QNETRNN = 1, POLICY_FOLLOWER = 2, POLICY_RANDOM = 3 };

typedef struct OraclePolicy {
    int32_t kind;
    int32_t pad_;
    uint64_t eps_threshold;     /* explore iff (uint64)r0 < eps_threshold ; floor(eps * 2^32) */
    double follower_tol;        /* tests/arena.py:213 (0.02) / tests/test_round_robin.py:224 (0.01): a Python float */
    const float *W1, *b1, *W2, *b2, *Wv, *bv, *Wa, *ba;   /* QNet effective weights */
} OraclePolicy;

/* action selection for one player: scripts/train_iterative.py:124-130 (eps-greedy),
 * tests/arena.py:199-219 (greedy / ball follower) */
static int policy_act(const OraclePolicy *pi, const float *obs, uint32_t env_id, uint32_t step,
                      uint32_t stream, uint64_t seed)
{
    uint32_t r[4];
    int a = 1;
    if (pi->kind == POLICY_RANDOM) {
        oracle_philox4x32(env_id, step, stream, 0, (uint32_t)seed, (uint32_t)(seed >> 32), r);
        return (int)(((uint64_t)r[1] * 3u) >> 32);
    }
    if (pi->kind == POLICY_FOLLOWER) {
        /* np.float32 - python float -> float64 under the pinned numpy 1.24.3 (requirements.txt:2); compares in double */
        double x = (double)obs[0], lo = (double)obs[4] - pi->follower_tol, hi = (double)obs[4] + pi->follower_tol;
        a = x < lo ? 0 : (x > hi ? 2 : 1);
    } else {
        uint8_t g;
        oracle_qnet_forward(1, obs, pi->W1, pi->b1, pi->W2, pi->b2, pi->Wv, pi->bv, pi->Wa, pi->ba, 0, &g);
        a = g;
    }
    if (pi->eps_threshold) {
        oracle_philox4x32(env_id, step, stream, 0, (uint32_t)seed, (uint32_t)(seed >> 32), r);
        if ((uint64_t)r[0] < pi->eps_threshold) a = (int)(((uint64_t)r[1] * 3u) >> 32);
    }
    return a;
}

/* Closed-loop self-play, K lock-step steps: {act A, act B, step, [replay row], auto-reset}.
 * Same bookkeeping as oracle_rollout_*; actions come from the two policies instead of a stream.
 * Optional outputs: actions_out[K][n][2]; replay rows of player B's view in (t, env) order
 * (scripts/train_iterative.py:243): rp_obs[.][7], rp_act, rp_rew, rp_next[.][7] (terminal obs, not
 * the post-reset one), rp_done; *n_rp counts rows. */
#define SELFPLAY_IMPL(REAL, SUF)                                                                      \
void oracle_selfplay_##SUF(const OracleParams *p, int64_t n, int64_t K,                               \
        REAL *x, REAL *y, REAL *vx, REAL *vy, REAL *spin, REAL *top, REAL *bot,                       \
        int32_t *sa, int32_t *sb, int32_t *bounce, int32_t *ep_idx, int32_t *ep_len,                  \
        const OraclePolicy *polA, const OraclePolicy *polB, uint64_t seed, int64_t step_base,         \
        const REAL *pool_vx, const REAL *pool_vy, const REAL *pool_spin, int32_t depth,               \
        int32_t quota, int64_t env_id_base, uint8_t *actions_out,                                     \
        int64_t *counters, int32_t *ep_log, int64_t log_cap, int64_t *n_log,                          \
        float *rp_obs, uint8_t *rp_act, float *rp_rew, float *rp_next, uint8_t *rp_done,              \
        int64_t rp_cap, int64_t *n_rp)                                                                \
{                                                                                                     \
    for (int64_t t = 0; t < K; ++t) {                                                                 \
        for (int64_t i = 0; i < n; ++i) {                                                             \
            if (quota > 0 && ep_idx[i] >= quota) continue;                                            \
            Env_##SUF s;                                                                              \
            load_##SUF(&s, i, x, y, vx, vy, spin, top, bot, sa, sb, bounce);                          \
            float oa[7], ob[7], na[7], nb[7];                                                         \
            observe_one_##SUF(&s, oa, ob);                                                            \
            uint32_t g = (uint32_t)(env_id_base + i), st = (uint32_t)(step_base + t);                 \
            int aA = policy_act(polA, oa, g, st, STREAM_ACT_A, seed);                                 \
            int aB = policy_act(polB, ob, g, st, STREAM_ACT_B, seed);                                 \
            if (actions_out) { actions_out[(t * n + i) * 2] = (uint8_t)aA;                            \
                               actions_out[(t * n + i) * 2 + 1] = (uint8_t)aB; }                      \
            int flags = step_one_##SUF(p, &s, aA, aB);                                                \
            ep_len[i] += 1;                                                                           \
            counters[0] += 1;                                                                         \
            if (flags & 2) counters[4] += 1;                                                          \
            if (flags & 4) counters[5] += 1;                                                          \
            if (flags & 8) counters[6] += 1;                                                          \
            if (rp_obs) {                                                                             \
                if (*n_rp < rp_cap) {                                                                 \
                    int64_t r_ = *n_rp;                                                               \
                    observe_one_##SUF(&s, na, nb);                                                    \
                    memcpy(rp_obs + r_ * 7, ob, sizeof ob);                                           \
                    memcpy(rp_next + r_ * 7, nb, sizeof nb);                                          \
                    rp_act[r_] = (uint8_t)aB;                                                         \
                    rp_rew[r_] = (flags & 4) ? 1.0f : ((flags & 2) ? -1.0f : 0.0f);                   \
                    rp_done[r_] = (uint8_t)(flags & 1);                                               \
                }                                                                                     \
                *n_rp += 1;                                                                           \
            }                                                                                         \
            if (flags & 1) {                                                                          \
                counters[1] += 1;                                                                     \
                if (s.sa > s.sb) counters[2] += 1; else counters[3] += 1;                             \
                counters[7] += ep_len[i];                                                             \
                if (ep_log && *n_log < log_cap) {                                                     \
                    int32_t *r = ep_log + (*n_log) * 4;                                               \
                    r[0] = (int32_t)(env_id_base + i); r[1] = ep_idx[i];                              \
                    r[2] = (s.sa << 16) | s.sb; r[3] = ep_len[i];                                     \
                }                                                                                     \
                if (n_log) *n_log += 1;                                                               \
                ep_idx[i] += 1;                                                                       \
                if (!(quota > 0 && ep_idx[i] >= quota)) {                                             \
                    size_t j = (size_t)(ep_idx[i] % depth) * n + i;                                   \
                    serve_one_##SUF(&s, pool_vx[j], pool_vy[j], pool_spin[j]);                        \
                    ep_len[i] = 0;                                                                    \
                }                                                                                     \
            }                                                                                         \
            store_##SUF(&s, i, x, y, vx, vy, spin, top, bot, sa, sb, bounce);                         \
        }                                                                                             \
    }                                                                                                 \
}
SELFPLAY_IMPL(double, f64)
SELFPLAY_IMPL(float, f32)

/* 53-bit uniform in [0,1) from two 32-bit words, as CPython's random.random() builds it. */
static double u53(uint32_t a, uint32_t b) { return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) / 9007199254740992.0; }

/* Throughput-mode serve: the reference formula (envs/my_pong_env_2p.py:98-111) driven by Philox
 * instead of MT19937; distribution-equal, not bit-equal, to the reference's RNG. */
void oracle_philox_serve(uint64_t seed, uint32_t env_id, uint32_t ep_idx,
                         double speed_lo, double speed_hi, const double *angles4 /* a0lo a0hi a1lo a1hi */,
                         double spin_lo, double spin_hi, double *out3)
{
    uint32_t r[4], q[4];
    oracle_philox4x32(env_id, ep_idx, STREAM_SERVE, 0, (uint32_t)seed, (uint32_t)(seed >> 32), r);
    oracle_philox4x32(env_id, ep_idx, STREAM_SERVE, 1, (uint32_t)seed, (uint32_t)(seed >> 32), q);
    double speed = speed_lo + (speed_hi - speed_lo) * u53(r[0], r[1]);
    int which = u53(r[2], r[3]) < 0.5 ? 0 : 1;
    double lo = angles4[2 * which], hi = angles4[2 * which + 1];
    double deg = lo + (hi - lo) * u53(q[0], q[1]);
    double rad = deg * (3.14159265358979323846 / 180.0);          /* math.radians */
    out3[0] = speed * cos(rad);
    out3[1] = speed * sin(rad);
    out3[2] = spin_lo + (spin_hi - spin_lo) * u53(q[2], q[3]);
}

int oracle_version(void) { return 1; }
