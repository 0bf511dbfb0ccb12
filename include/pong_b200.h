/*
 * pong_b200.h — C ABI of libpong_b200.so, the B200 (sm_100a) engine for the self-play hot path of
 * MaxChen228/pingpong-selfplay-ai: N lock-step PongEnv2P environments stepped on the device with both
 * paddles' actions chosen by QNet / QNetRNN on the device.
 *
 * The reference has no FFI (it is pure Python); each entry point below names the reference
 * interface it replaces (file:line under /root/reference).  INTEGRATION.md shows the ctypes stub a
 * maintainer of the reference adds to call them.
 *
 * Conventions
 *   - plain C: raw pointers + sizes + a cudaStream_t passed as void*; no torch types.
 *   - every function returns int: 0 ok, <0 bad argument (PP_E_*), >0 a cudaError_t value.
 *     pp_last_error() gives a text for the calling thread's last non-zero return.
 *   - pointers are DEVICE pointers unless the parameter is named host_* or the function is pp_host_*.
 *   - functions never allocate and never synchronise (pp_host_* excepted: they own their staging
 *     buffers and return after the result is in the host buffers).
 *   - `mode` selects the arithmetic of the env state: PP_MODE_F64 reproduces the reference's
 *     IEEE double arithmetic bit for bit (one rounding per Python operation, no FMA contraction);
 *     PP_MODE_F32 is the same operation order in binary32.
 *   - env state is SoA: one array of n reals per field, owned by the caller (torch tensors).
 *   - there is NO CPU fallback: without a CUDA device every compute entry returns a cudaError_t.
 */
#ifndef PONG_B200_H_
#define PONG_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PP_ABI_VERSION 6

enum { PP_MODE_F64 = 0, PP_MODE_F32 = 1 };

enum {
    PP_OK = 0,
    PP_E_NULL = -1,       /* required pointer is NULL         */
    PP_E_SIZE = -2,       /* n / k / capacity out of range    */
    PP_E_MODE = -3,       /* unknown mode / policy / precision */
    PP_E_ALIGN = -4,      /* pointer not aligned as required  */
    PP_E_PARAM = -5       /* inconsistent PPParams            */
};

/* Environment constants, precomputed on the host "the Python way" so that YAML ints
 * (restitution: 1) and libm pow (R**2) behave as in the reference.
 * Replaces the attributes stored by PongEnv2P.__init__  envs/my_pong_env_2p.py:19-62
 * and the per-call constants of collide_sphere_with_moving_plane  envs/physics.py:7-11. */
typedef struct PPParams {
    double paddle_speed;        /* my_pong_env_2p.py:43                                   */
    double half_width;          /* paddle_width / 2              :152-153,190-191        */
    double magnus_factor;       /*                               :136                     */
    double neg_e;               /* -restitution                  physics.py:7             */
    double m_1pe;               /* m * (1 + e)                   physics.py:8             */
    double inertia;             /* (2/5) * m * R**2              physics.py:9             */
    double two_m_over_7;        /* 2*m/7.0                       physics.py:10            */
    double mu;                  /* friction                      physics.py:11            */
    double mass;                /* ball_mass                     physics.py:20            */
    double radius;              /* world_ball_radius             physics.py:10,16,21      */
    double speed_scale;         /* 1.0 + speed_increment         my_pong_env_2p.py:230    */
    /* serve distribution (device-generated serves only)        my_pong_env_2p.py:98-111 */
    double speed_lo, speed_hi;
    double angle_lo[2], angle_hi[2];    /* degrees; interval 0 is taken when u < 0.5      */
    double spin_lo, spin_hi;
    int32_t enable_spin;
    int32_t max_score;
    int32_t speed_scale_every;
    int32_t reserved;
} PPParams;

/* SoA state of n environments.  `real` = double (PP_MODE_F64) or float (PP_MODE_F32).
 * Fields mirror the public attributes of PongEnv2P  envs/my_pong_env_2p.py:85-111. */
typedef struct PPEnvState {
    void *ball_x, *ball_y, *ball_vx, *ball_vy, *spin, *top_paddle_x, *bottom_paddle_x;   /* real[n]  */
    int32_t *score_a, *score_b, *bounce_count;                                          /* int32[n] */
    int32_t *ep_idx;    /* episodes this env has finished = index of its current serve; may be NULL for pp_env_step */
    int32_t *ep_len;    /* steps taken in the current episode;                          may be NULL for pp_env_step */
} PPEnvState;

/* Where reset() takes (vx, vy, spin) from.  The reference draws them from CPython's global
 * MT19937 (envs/my_pong_env_2p.py:98-111), which cannot be reproduced on the device:
 *   PP_SERVE_POOL    host-generated queue, real[depth][n]; env i's j-th episode uses row j % depth
 *   PP_SERVE_PHILOX  Philox4x32-10 keyed by (seed; global env id, episode index): same formula,
 *                    distribution-equal, and independent of how envs are sharded over GPUs.
 *   PP_SERVE_QUEUE   evaluation mode: the pool is ONE queue of queue_total serves (flat index q = j * n + i, the
 *                    same memory as a [depth][n] pool; or, with all three pool pointers NULL, serve q is the Philox
 *                    serve (seed; env q % n, episode q / n)) and an env that finishes an episode claims the next
 *                    unplayed serve with an atomic on *queue_head instead of waiting for its own next one.  Every
 *                    serve is still played exactly once and an episode depends on nothing but its serve and the
 *                    (greedy) players, so counters and episode-log rows (env = q % n, episode = q / n) equal
 *                    those of the fixed per-env quota — without the lock-step tail in which finished envs
 *                    idle.  ep_idx holds q; pass quota = queue_total; *queue_head starts at n. */
enum { PP_SERVE_POOL = 0, PP_SERVE_PHILOX = 1, PP_SERVE_QUEUE = 2 };
typedef struct PPServeSource {
    int32_t kind;
    int32_t depth;
    const void *pool_vx, *pool_vy, *pool_spin;
    uint64_t seed;
    unsigned long long *queue_head;
    int64_t queue_total;
} PPServeSource;

/* A player.  Replaces the per-step `model(torch.tensor(obs).unsqueeze(0)).argmax(1).item()` of
 * scripts/train_iterative.py:124-130,176-177,240 and select_action_universal tests/arena.py:199-219. */
enum { PP_POLICY_QNET = 0, PP_POLICY_QNETRNN = 1, PP_POLICY_FOLLOWER = 2, PP_POLICY_RANDOM = 3 };
enum { PP_PREC_F32 = 0,     /* CUDA-core fp32, fmaf chain in ascending k: bit-identical to the oracle          */
       PP_PREC_F16 = 1 };   /* tcgen05 tensor cores: fp16 operands (obs and biases split hi/lo), fp32 accumulation
                             * in TMEM; Q within 1e-3 of the fp32 reference.  fp16 rather than bf16: same tensor
                             * rate, 8x smaller rounding error, range handled by saturating conversions.        */
typedef struct PPPolicy {
    int32_t kind;
    int32_t precision;
    uint64_t eps_threshold;     /* explore iff (uint64)philox.x < eps_threshold; floor(eps * 2^32), 0 = greedy */
    double follower_tol;        /* tests/arena.py:213 — a Python float: the reference (numpy 1.24.3, requirements.txt:2)
                                 * evaluates `np.float32 - 0.02` and the compares in float64, and so does the engine   */
    const float *weights;       /* packed blob, see PP_QNET_* / PP_RNN_* offsets (QNetRNN with PP_PREC_F16: the
                                 * fp16 image PP_RNNTC_*)                                                     */
    float *h, *c;               /* QNetRNN only, 128 floats per env each, zeroed by the engine at episode start.  PP_PREC_F32:
                                 * unit-major [128][n].  PP_PREC_F16: blocked by warp, [ceil(n / 32)][32][32][4] — unit u of
                                 * env i at ((i / 32 * 32 + u / 4) * 32 + i % 32) * 4 + u % 4                            */
} PPPolicy;

/* QNet blob (floats): effective weights (eval: mu, train: mu + sigma*eps — models/qnet.py:43-50),
 * transposed to k-major so a warp reads one weight row with broadcast 128-bit loads.
 *   W1t[7][64]  b1[64]  W2t[64][64]  b2[64]  Wht[64][4] (col 0 = V, 1..3 = A)  bh[4] */
enum {
    PP_QNET_W1T = 0, PP_QNET_B1 = 448, PP_QNET_W2T = 512, PP_QNET_B2 = 4608,
    PP_QNET_WHT = 4672, PP_QNET_BH = 4928, PP_QNET_BLOB_FLOATS = 4932
};

/* QNetRNN blob (floats), default dims 7-64-128 / LSTM 128 / head 128 (config_rnn.yaml:39-42):
 *   Wf1t[7][64] bf1[64] Wf2t[64][128] bf2[128] Wgt[256][512] (rows 0..127 = W_ih^T, 128..255 = W_hh^T;
 *   column = unit*4 + gate, gates i,f,g,o) bg[512] (= b_ih + b_hh) Wst[128][128] bs[128] Wht[128][4] bh[4] */
enum {
    PP_RNN_WF1T = 0, PP_RNN_BF1 = 448, PP_RNN_WF2T = 512, PP_RNN_BF2 = 8704, PP_RNN_WGT = 8832,
    PP_RNN_BG = 139904, PP_RNN_WST = 140416, PP_RNN_BS = 156800, PP_RNN_WHT = 156928, PP_RNN_BH = 157440,
    PP_RNN_BLOB_FLOATS = 157444
};

/* QNetRNN image for the tensor-core path (precision PP_PREC_F16): fp16 B-operand tiles in the order the kernel
 * streams them through shared memory with TMA, one "stage" per bulk copy.  Every tile is K-major, no swizzle,
 * [K/8][N][8 halves]; *H = fp16(w), *L = fp16(w - fp16(w)); a bias tile is [2][N][8] with the bias split hi / lo in
 * rows k = 7 / 15 (it multiplies the ones columns of the observation tile).  Offsets in BYTES:
 *   S0   W1H[2][64][8] W1L[2][64][8]                       features.0 (obs hi/lo, bias in rows 7/15)
 *   S1   WF2H[8][128][8] BF2[2][128][8]     S2  WF2L[8][128][8]        features.2   64 -> 128
 *   per quarter q = 0..3 (units 32q..32q+31), K rows 64c..64c+63 of [W_ih^T ; W_hh^T], column = gate*32 + unit%32:
 *        GH(q,c) c = 0..3: WGH[8][128][8] (GH(q,0) is followed by its bias tile BG[2][128][8] = b_ih + b_hh),
 *        then GL(q,c) c = 0..3: WGL[8][128][8]
 *   WS0  WSH k<64 + BS[2][128][8]   WS1  WSH k>=64   WS2  WSL k<64   WS3  WSL k>=64      fc_shared_head 128 -> 128
 *   HD   WHH[16][16][8] WHL[16][16][8] BH[2][16][8]        dueling heads, columns 0..3 = V, A0, A1, A2 */
enum {
    PP_RNNTC_TILE = 16384, PP_RNNTC_BIAS = 4096,
    PP_RNNTC_S0 = 0, PP_RNNTC_S0_BYTES = 4096,
    PP_RNNTC_S1 = 4096, PP_RNNTC_S1_BYTES = 20480,
    PP_RNNTC_S2 = 24576, PP_RNNTC_S2_BYTES = 16384,
    PP_RNNTC_G = 40960, PP_RNNTC_GQ_BYTES = 135168,       /* per quarter: 20480 + 7 * 16384 */
    PP_RNNTC_WS = 581632, PP_RNNTC_WS_BYTES = 69632,
    PP_RNNTC_HD = 651264, PP_RNNTC_HD_BYTES = 8704,
    PP_RNNTC_BLOB_BYTES = 659968, PP_RNNTC_STAGES = 40, PP_RNNTC_SLOT_BYTES = 20480
};

/* Per-call outputs of the multi-step kernels.  counters[8] (accumulated with atomics, never reset
 * by the engine): 0 env-steps, 1 episodes, 2 wins A, 3 wins B, 4 points A, 5 points B,
 * 6 paddle hits, 7 sum of finished-episode lengths.  ep_log[cap][4] = {global env id, ep_idx,
 * scoreA<<16|scoreB, ep_len}; *ep_log_count counts every finished episode, also those beyond cap. */
typedef struct PPRolloutOut {
    unsigned long long *counters;
    int32_t *ep_log;
    int64_t ep_log_cap;
    unsigned long long *ep_log_count;
    uint8_t *actions_out;       /* [k][n][2] or NULL                                           */
    void *trace_real;           /* real[k][7][n] post-step pre-reset state, or NULL            */
    int32_t *trace_int;         /* int32[k][4][n] = scoreA, scoreB, bounce, flags, or NULL     */
} PPRolloutOut;

/* Replay ring of player B's transitions (oB, aB, rB, nB, done)  scripts/train_iterative.py:243;
 * 62 bytes per row over five arrays; *head is a monotonically increasing write cursor (slot = cursor %
 * capacity).  A call that pushes more rows than `capacity` writes only the rows that sequential pushes
 * would leave in the ring (the last `capacity` rows of pp_replay_scatter, the last capacity / n lock-step
 * steps of pp_selfplay_rollout, which needs capacity >= n); *head counts written rows.
 *
 * Lock-step layout (lockstep_envs = n > 0, capacity % n == 0; pp_selfplay_rollout only): the row of env i at
 * lock-step step t of the launch goes to slot ((lockstep_step0 + t) % (capacity / n)) * n + i, i.e. the ring is a
 * [capacity / n][n] array in time order per env — what episode-SEQUENCE replay needs (SequenceReplayBuffer,
 * scripts/train_rnn_iterative.py:100-171).  The caller keeps lockstep_step0 (steps written so far); *head is not
 * touched by the kernel in this layout.  Envs frozen by a quota write nothing. */
typedef struct PPReplayRing {
    float *obs;                 /* [capacity][7] */
    uint8_t *act;               /* [capacity]    */
    float *rew;                 /* [capacity]    */
    float *next_obs;            /* [capacity][7] terminal obs on done, not the post-reset one */
    uint8_t *done;              /* [capacity]    */
    int64_t capacity;
    unsigned long long *head;
    int64_t lockstep_envs;      /* 0 = append layout (compacted rows, cursor *head)           */
    int64_t lockstep_step0;     /* lock-step layout: steps written before this launch          */
} PPReplayRing;

/* One NoisyLinear layer (models/qnet.py:6-50): device tensors in torch's layout.  The grad_* pointers are written by
 * pp_dqn_head_grads (each may be NULL) and ignored elsewhere. */
typedef struct PPNoisyLayer {
    int32_t in_features, out_features;
    float *weight_mu, *weight_sigma, *weight_epsilon;       /* [out][in] */
    float *bias_mu, *bias_sigma, *bias_epsilon;             /* [out]     */
    float *grad_weight_mu, *grad_weight_sigma;              /* [out][in] */
    float *grad_bias_mu, *grad_bias_sigma;                  /* [out]     */
} PPNoisyLayer;

/* One parameter tensor of torch.optim.Adam with its optimiser state (optimizer.state[p]): all device pointers.
 * `step` is the 0-d float32 step counter of a capturable optimiser. */
typedef struct PPAdamParam {
    float *param, *grad, *exp_avg, *exp_avg_sq, *step;
    int64_t numel;
} PPAdamParam;

/* One QNetRNN in torch's layout (models/qnet_rnn.py:53-105; default dims 7-64-128 / LSTM 128 / head 128): device
 * pointers to the module's own parameter tensors.  The grad_* pointers inside the three PPNoisyLayer are ignored here. */
typedef struct PPQNetRNNParams {
    const float *f0_w, *f0_b;            /* features_extractor.0  [64][7], [64]                          */
    const float *f2_w, *f2_b;            /* features_extractor.2  [128][64], [128]                       */
    const float *w_ih, *w_hh;            /* lstm.weight_ih_l0 / weight_hh_l0  [512][128] (gates i, f, g, o) */
    const float *b_ih, *b_hh;            /* lstm.bias_ih_l0 / bias_hh_l0  [512]                          */
    PPNoisyLayer shared, v, a;           /* fc_shared_head.0 [128][128], fc_V [1][128], fc_A [3][128]    */
} PPQNetRNNParams;

/* Where pp_drqn_grads writes d loss / d parameter (written, not accumulated; any pointer may be NULL = skipped): plain
 * tensors for the feature layers and the LSTM, the grad_* pointers of the PPNoisyLayer for the three noisy layers (their
 * weight / bias / epsilon pointers are ignored here; shared.grad_weight_mu is required). */
typedef struct PPQNetRNNGrads {
    float *f0_w, *f0_b, *f2_w, *f2_b, *w_ih, *w_hh, *b_ih, *b_hh;
    PPNoisyLayer shared, v, a;
} PPQNetRNNGrads;

int pp_version(void);
const char *pp_last_error(void);

/* PongEnv2P.step for n envs, all outputs materialised          envs/my_pong_env_2p.py:116-225,235-263
 * obs [n][7] fp32, rewards fp32, done u8.  No reset. */
int pp_env_step(int mode, int64_t n, const PPParams *params, const PPEnvState *state,
                const uint8_t *action_a, const uint8_t *action_b,
                float *obs_a, float *obs_b, float *reward_a, float *reward_b, uint8_t *done, void *stream);

/* PongEnv2P._get_obs                                           envs/my_pong_env_2p.py:235-263 */
int pp_env_observe(int mode, int64_t n, const PPEnvState *state, float *obs_a, float *obs_b, void *stream);

/* PongEnv2P.reset with injected serves (vx, vy, spin real[n]) for envs whose mask byte is non-zero
 * (mask NULL = all)                                            envs/my_pong_env_2p.py:83-114 */
int pp_env_serve(int mode, int64_t n, const PPEnvState *state, const uint8_t *mask,
                 const void *vx, const void *vy, const void *spin, void *stream);

/* PongEnv2P.reset drawing from a PPServeSource; `advance` != 0 first increments ep_idx of the
 * reset envs (the reference consumes one serve per reset() call). */
int pp_env_reset(int mode, int64_t n, const PPParams *params, const PPEnvState *state, const uint8_t *mask,
                 const PPServeSource *serve, int64_t env_id_base, int advance, void *stream);

/* collide_sphere_with_moving_plane(vn, vt, u, omega, e, mu, m, R) for n independent impacts   envs/physics.py:3-23
 * all arrays real[n]; of `params` only neg_e, m_1pe, inertia, two_m_over_7, mu, mass and radius are read (the
 * constants the reference computes per call at physics.py:7-11, precomputed "the Python way" by the host). */
int pp_collide(int mode, int64_t n, const PPParams *params, const void *vn, const void *vt, const void *u,
               const void *omega, void *vn_out, void *vt_out, void *omega_out, void *stream);

/* k lock-step steps with an injected action stream actions[k][n][2] and auto-reset; state stays in
 * registers between steps.  An env whose ep_idx reached `quota` (> 0) is frozen.
 * Replaces the `step(); if done: reset()` loop of scripts/train_iterative.py:174-179. */
int pp_env_rollout(int mode, int64_t n, int64_t k, const PPParams *params, const PPEnvState *state,
                   const uint8_t *actions, const PPServeSource *serve, int32_t quota, int64_t env_id_base,
                   const PPRolloutOut *out, void *stream);

/* obs[n][7] -> Q -> (epsilon-)greedy action for one player       models/qnet.py:71-75 +
 * scripts/train_iterative.py:124-130.  q_out [n][3] may be NULL.  `stream_id` = 1 for player A, 2 for B. */
int pp_qnet_act(int64_t n, const float *obs, const PPPolicy *policy, uint64_t seed, int64_t step_index,
                int64_t env_id_base, int32_t stream_id, uint8_t *actions, float *q_out, void *stream);

/* one QNetRNN step (seq_len 1) with carried per-env (h, c)        models/qnet_rnn.py:107-144.
 * reset_mask (may be NULL): envs whose byte is non-zero get (h, c) = 0 BEFORE the step
 * (init_hidden at episode start, tests/arena.py:298-299). */
int pp_qnetrnn_act(int64_t n, const float *obs, const PPPolicy *policy, const uint8_t *reset_mask,
                   uint64_t seed, int64_t step_index, int64_t env_id_base, int32_t stream_id,
                   uint8_t *actions, float *q_out, void *stream);

/* k fused lock-step iterations of {act A, act B, step, replay row, auto-reset}: the inner loops of
 * scripts/train_iterative.py:171-196,238-245 and tests/arena.py:294-304.  Policies of kind QNET,
 * FOLLOWER and RANDOM; observations never leave registers.  ring may be NULL.
 * With PP_PREC_F16 QNet players the launch is preceded, on the same stream, by a one-block kernel and a 1 KB
 * device-to-device copy that place the players' head table in constant memory; every (device, stream) owns a table
 * slot (16 per process; further streams read the table from shared memory instead), so launches on different
 * streams never share one.  A launch captured into a CUDA graph keeps the slot of the stream it was captured on:
 * do not replay it on another stream while the capture stream runs rollouts of other weights.
 * PP_CONST_HEADS=0 in the environment selects the shared-memory table everywhere. */
int pp_selfplay_rollout(int mode, int64_t n, int64_t k, const PPParams *params, const PPEnvState *state,
                        const PPPolicy *policy_a, const PPPolicy *policy_b, uint64_t seed, int64_t step_base,
                        const PPServeSource *serve, int32_t quota, int64_t env_id_base,
                        const PPRolloutOut *out, const PPReplayRing *ring, void *stream);

/* memory.push for a batch: append rows whose valid byte is non-zero (NULL = all) to the ring,
 * compacted with warp ballot + scan, one cursor atomic per warp   scripts/train_iterative.py:56-63,243 */
int pp_replay_scatter(int64_t n, const PPReplayRing *ring, const float *obs, const uint8_t *act,
                      const float *rew, const float *next_obs, const uint8_t *done, const uint8_t *valid,
                      void *stream);

/* Whole evaluation from HOST buffers (what a reference caller holds): eval_vs_model of
 * scripts/train_iterative.py:171-181 for n envs x quota episodes each, QNet A vs QNet B, on GPU `device`.
 * Serves: host_pool_* = real[quota][n] arrays drawn by the caller (the reference's env.reset() formula, any RNG), or
 * all three NULL = drawn on the device from Philox keyed by (seed; env_id_base + i, episode j) — the counterpart of
 * random.seed(seed) before the reference's loop; nothing but the two weight blobs (39 KB) then crosses PCIe.
 * Copies inputs to the device, plays the n x quota serves as ONE queue (PP_SERVE_QUEUE: an env that finishes claims the
 * next unplayed serve) in a single launch of at most max_steps lock-step steps, copies counters[8] and the per-episode
 * records back (host_ep_log may be NULL).  Returns after the results are in the host buffers.
 * Sharding: one call per GPU with its own `device`, slab size n and env_id_base (results do not depend on the split);
 * calls for different devices may run concurrently from different host threads or processes, calls for one device
 * are serialised.  The staging buffers of a device are kept between calls; pp_host_release(device) frees them
 * (device = -1: all).  The caller's current CUDA device is left unchanged. */
int pp_host_selfplay_eval(int device, int mode, int64_t n, int32_t quota, const PPParams *params,
                          const void *host_pool_vx, const void *host_pool_vy, const void *host_pool_spin,
                          uint64_t seed, int64_t env_id_base,
                          const float *host_weights_a, const float *host_weights_b, int32_t precision,
                          int64_t max_steps,
                          unsigned long long *host_counters, int32_t *host_ep_log, int64_t ep_log_cap);
int pp_host_release(int device);

/* ---- training mode (scripts/train_iterative.py): three small launches per update instead of ~130 framework kernels */

/* NoisyLinear.reset_noise (models/qnet.py:33-41) for `count` <= 8 layers in one launch: e = sign(g) sqrt|g| with
 * g ~ N(0, 1) (Philox keyed by seed, *counter, layer, index; Box-Muller), weight_epsilon = outer(e_out, e_in),
 * bias_epsilon = e_out.  *counter (device memory) is read and incremented by the kernel, so a CUDA graph that
 * contains the launch draws fresh noise at every replay.  in_features + out_features <= 1024 per layer. */
int pp_noisy_reset(const PPNoisyLayer *layers, int32_t count, uint64_t seed, unsigned long long *counter, void *stream);

/* The reference QNet (features.0 [64][7], features.2 [64][64], NoisyLinear fc_V [1][64] and fc_A [3][64]) in torch's
 * layout -> the k-major PP_QNET_* blob the act / rollout kernels take.  noisy != 0: the train-mode forward
 * mu + sigma * epsilon (models/qnet.py:44-46), else mu. */
int pp_pack_qnet(const float *features0_weight, const float *features0_bias, const float *features2_weight,
                 const float *features2_bias, const PPNoisyLayer *fc_v, const PPNoisyLayer *fc_a, int32_t noisy,
                 float *blob, void *stream);

/* train_step() of scripts/train_iterative.py:139-164 up to the gradients, for a sampled batch of replay rows:
 *   q = Q(s)[a];  a* = argmax Q(s');  target = r + gamma * Q_target(s')[a*] * (1 - done);  td = q - target;
 *   loss = mean(iw * td^2)
 * and d loss / d {weight_mu, weight_sigma, bias_mu, bias_sigma} of the online heads fc_V / fc_A written (not
 * accumulated) to the grad_* pointers of online_v / online_a; the feature layers (features.0 [64][7], features.2
 * [64][64], torch layout) are frozen (:97) and shared by the online and the target net.  idx[batch] are ring slots,
 * iw[batch] the importance weights.  td_out[batch], loss_out[1] and prios (the PER priority array indexed by ring
 * slot, receives |td| + 1e-6, :74-76,163-164) may be NULL; *max_prio (may be NULL) is raised to the largest priority written
 * (a running maximum: what new rows get at :57,62 without an O(capacity) pass per push).  noisy_online / noisy_target select the train- or
 * eval-mode forward of each net (the reference: online train mode, target eval mode :100).  batch <= 4096.
 * `workspace`: pp_dqn_workspace_floats(batch) floats of device memory, ZERO before the first launch that uses it (the
 * kernel leaves it ready for the next one): per-tile partial sums + the ticket of the CTA that finishes last. */
int pp_dqn_head_grads(const PPReplayRing *ring, const int64_t *idx, const float *iw, int32_t batch,
                      const float *features0_weight, const float *features0_bias, const float *features2_weight,
                      const float *features2_bias, const PPNoisyLayer *online_v, const PPNoisyLayer *online_a,
                      const PPNoisyLayer *target_v, const PPNoisyLayer *target_a, int32_t noisy_online,
                      int32_t noisy_target, float gamma, float *td_out, float *loss_out, float *prios, float *max_prio,
                      float *workspace, void *stream);
int64_t pp_dqn_workspace_floats(int32_t batch);

/* PrioritizedReplay.sample (scripts/train_iterative.py:64-73) on the device: `batch` slots drawn i.i.d. with
 * probability prios[i]^alpha / sum (np.random.choice(p=probs), :68) by a two-level inverse-CDF search, and the importance
 * weights (N P(i))^-beta / max (:71-72).  prios[capacity]: 0 = empty slot (never drawn).  *beta and *size (N = len(buffer),
 * as a float) are read from device memory; *counter (device) is read and incremented, so a CUDA graph that contains the
 * call draws a fresh batch at every replay.  chunk_sums: pp_per_sample_scratch_floats(capacity) floats of scratch.
 * batch <= 4096.  Deterministic for a given (seed, *counter, prios). */
int pp_per_sample(const float *prios, int64_t capacity, float alpha, const float *beta, const float *size, uint64_t seed,
                  unsigned long long *counter, int32_t batch, float *chunk_sums, int64_t *idx_out, float *weights_out,
                  void *stream);
int64_t pp_per_sample_scratch_floats(int64_t capacity);

/* optimizerB.step() of scripts/train_iterative.py:162 — torch.optim.Adam (betas, eps as given; no amsgrad, no weight
 * decay) for `count` <= 16 small tensors in one launch, IN PLACE on the parameter and on the optimiser's own state:
 *   step += 1;  m = lerp(m, g, 1 - beta1);  v = beta2 v + (1 - beta2) g^2;
 *   p -= lr / (1 - beta1^step) * m / (sqrt(v) / sqrt(1 - beta2^step) + eps) */
int pp_adam_step(const PPAdamParam *params, int32_t count, double lr, double beta1, double beta2, double eps, void *stream);

/* Several ranks (one process per GPU): optimizerB.step() with the gradient all-reduce of scripts/train_iterative.py's
 * data-parallel form FUSED in front of it, over peer-mapped memory (NVLink), in ONE launch and without NCCL.
 * blocks[r] = rank r's block as mapped into THIS process (e.g. torch.distributed._symmetric_memory rendezvous:
 * buffer_ptrs), each pp_peer_block_bytes(capacity_floats) bytes, zeroed before the first use on every rank:
 * float staging[2][capacity_floats]; uint32 flags[8].  flat_grad[numel] (numel <= capacity_floats) is the flat buffer the
 * grad pointers of `params` view; on return it holds the MEAN over ranks (summed in rank order: bit-identical on every
 * rank) and the parameters have taken one Adam step.  *epoch (device memory, starts at 0) counts the calls; every rank
 * must make the same calls in the same order — the kernel waits for its peers.  world <= 8, count <= 16. */
typedef struct PPPeerBlocks {
    void *blocks[8];
    int32_t rank, world;
    int64_t capacity_floats;
} PPPeerBlocks;
int pp_adam_step_allreduce(const PPAdamParam *params, int32_t count, float *flat_grad, int64_t numel, const PPPeerBlocks *peers,
                           unsigned long long *epoch, double lr, double beta1, double beta2, double eps, void *stream);
int64_t pp_peer_block_bytes(int64_t capacity_floats);

/* ---- DRQN training mode (scripts/train_rnn_iterative.py)
 *
 * train_step_rnn() of scripts/train_rnn_iterative.py:400-531 up to the gradients, for `batch` sampled windows of `trace`
 * consecutive transitions: rows[batch][trace] are replay-ring slots in time order (SequenceReplayBuffer.sample, :126-165).
 *   q      = Q_online(obs window, zero initial (h, c))[last step][action of the last step]                    :470-478
 *   a*     = argmax Q_online(next_obs window)[last step];   target = r + gamma Q_target(next_obs window)[a*] (1 - done)  :489-505
 *   loss   = smooth_l1(q, target)  (mean over the batch)                                                      :509
 * and d loss / d every parameter of the online net (BPTT through the LSTM over the whole window).  noisy_online /
 * noisy_target select the train- (mu + sigma * epsilon) or eval-mode (mu) forward of the NoisyLinear layers (the
 * reference: online train mode :729, target eval mode :337-338).  batch: a multiple of 16, <= 256; trace <= 16.
 * td_out[batch], loss_out[1] may be NULL.  workspace: pp_drqn_workspace_floats(batch, trace) floats. */
int pp_drqn_grads(const PPReplayRing *ring, const int64_t *rows, int32_t batch, int32_t trace,
                  const PPQNetRNNParams *online, const PPQNetRNNParams *target, int32_t noisy_online, int32_t noisy_target,
                  float gamma, const PPQNetRNNGrads *grads, float *loss_out, float *td_out, float *workspace, void *stream);
int64_t pp_drqn_workspace_floats(int32_t batch, int32_t trace);

/* SequenceReplayBuffer.sample (scripts/train_rnn_iterative.py:126-141) on a lock-step ring ([T][n], PPReplayRing with
 * lockstep_envs = n, `steps_written` steps so far): weights[T * n] (indexed by ring slot) = the probability weight of
 * the window of `trace` steps that ENDS at that slot — 1 / (len - trace + 1) if its episode is complete inside the ring,
 * at least `trace` long and contains the window, else 0 (every stored episode has total weight 1: episode first, window
 * second); *episodes = number of stored episodes (len(memory), :170-171).  starts_fresh != 0: step 0 of the ring is the
 * first step of an episode.  Draw window ends with pp_per_sample(weights, T * n, alpha = 1, ...), then
 * pp_seq_expand_rows turns the ends into rows[batch][trace] (time ascending) for pp_drqn_grads. */
int pp_seq_window_weights(const uint8_t *done, int64_t n, int64_t T, int64_t steps_written, int32_t trace, int32_t starts_fresh,
                          float *weights, unsigned long long *episodes, void *stream);
int pp_seq_expand_rows(const int64_t *end_slots, int32_t batch, int32_t trace, int64_t n, int64_t T, int64_t *rows, void *stream);

/* QNetRNN in torch's layout -> the fp16 stage image PP_RNNTC_* (PP_RNNTC_BLOB_BYTES bytes, 16-byte aligned) the tensor-core
 * kernels take as PPPolicy.weights with PP_PREC_F16; noisy != 0: the train-mode weights mu + sigma * epsilon of the three
 * NoisyLinear layers (models/qnet_rnn.py:44-46).  One launch; the host-side equivalent is policy.pack_qnetrnn_tc. */
int pp_pack_qnetrnn_tc(const PPQNetRNNParams *net, int32_t noisy, void *image, void *stream);

/* torch.nn.utils.clip_grad_norm_(parameters, max_norm) (:516) on ONE flat gradient buffer (the .grad tensors are views of
 * it): norm_out[0] = total L2 norm, norm_out[1] = the clip coefficient min(1, max_norm / (norm + 1e-6)) the buffer was
 * scaled by.  scratch: 257 floats, ZERO before the first use (left ready for the next call). */
int pp_clip_grad_norm(float *flat_grads, int64_t numel, float max_norm, float *norm_out, float *scratch, void *stream);

/* pp_adam_step for up to 32 tensors of any size, spread over the whole GPU (optimizerB.step(), :517). */
int pp_adam_step_multi(const PPAdamParam *params, int32_t count, double lr, double beta1, double beta2, double eps, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PONG_B200_H_ */
