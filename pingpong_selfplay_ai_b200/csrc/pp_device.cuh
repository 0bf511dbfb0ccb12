// pp_device.cuh — device-side building blocks shared by every kernel of libpong_b200.
//
//  * Exact<R>: IEEE arithmetic with one rounding per operation.  The reference is CPython float
//    arithmetic (envs/my_pong_env_2p.py, envs/physics.py): every binary op is rounded separately and
//    nvcc's default FMA contraction would change results bitwise within a few dozen steps
//    (SURVEY.md section 7).  __dmul_rn/__dadd_rn/... are never contracted.
//  * env_step<R>(): one PongEnv2P.step() on register-resident state.
//  * Philox4x32-10, the counter-based RNG behind exploration, random players and device serves.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/pong_b200.h"

namespace pp {

template <typename R> struct Exact;
template <> struct Exact<double> {
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
    static __device__ __forceinline__ double abs(double a) { return fabs(a); }
    static __device__ __forceinline__ double sign1(double a) { return copysign(1.0, a); }
    static __device__ __forceinline__ float to_f32(double a) { return __double2float_rn(a); }
};
template <> struct Exact<float> {
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
    static __device__ __forceinline__ float abs(float a) { return fabsf(a); }
    static __device__ __forceinline__ float sign1(float a) { return copysignf(1.0f, a); }
    static __device__ __forceinline__ float to_f32(float a) { return a; }
};

// Constants of one launch in the arithmetic type of the state.
template <typename R> struct EnvConsts {
    R ps, hw, magnus, neg_e, m_1pe, inertia, two_m_7, mu, mass, radius, scale;
    int enable_spin, max_score, scale_every;
    __device__ __forceinline__ explicit EnvConsts(const PPParams &p)
        : ps((R)p.paddle_speed), hw((R)p.half_width), magnus((R)p.magnus_factor), neg_e((R)p.neg_e),
          m_1pe((R)p.m_1pe), inertia((R)p.inertia), two_m_7((R)p.two_m_over_7), mu((R)p.mu),
          mass((R)p.mass), radius((R)p.radius), scale((R)p.speed_scale), enable_spin(p.enable_spin),
          max_score(p.max_score), scale_every(p.speed_scale_every) {}
};

template <typename R> struct Env {
    R x, y, vx, vy, spin, top, bot;
    int sa, sb, bounce;
};

enum : int { F_DONE = 1, F_POINT_A = 2, F_POINT_B = 4, F_HIT = 8 };

// Typed view of PPEnvState.
template <typename R> struct StatePtrs {
    R *x, *y, *vx, *vy, *spin, *top, *bot;
    int32_t *sa, *sb, *bounce, *ep_idx, *ep_len;
    __host__ __device__ explicit StatePtrs(const PPEnvState &s)
        : x((R *)s.ball_x), y((R *)s.ball_y), vx((R *)s.ball_vx), vy((R *)s.ball_vy), spin((R *)s.spin),
          top((R *)s.top_paddle_x), bot((R *)s.bottom_paddle_x), sa(s.score_a), sb(s.score_b),
          bounce(s.bounce_count), ep_idx(s.ep_idx), ep_len(s.ep_len) {}
};

template <typename R> __device__ __forceinline__ Env<R> load_env(const StatePtrs<R> &s, int64_t i) {
    Env<R> e;
    e.x = s.x[i]; e.y = s.y[i]; e.vx = s.vx[i]; e.vy = s.vy[i]; e.spin = s.spin[i];
    e.top = s.top[i]; e.bot = s.bot[i]; e.sa = s.sa[i]; e.sb = s.sb[i]; e.bounce = s.bounce[i];
    return e;
}
template <typename R> __device__ __forceinline__ void store_env(const StatePtrs<R> &s, int64_t i, const Env<R> &e) {
    s.x[i] = e.x; s.y[i] = e.y; s.vx[i] = e.vx; s.vy[i] = e.vy; s.spin[i] = e.spin;
    s.top[i] = e.top; s.bot[i] = e.bot; s.sa[i] = e.sa; s.sb[i] = e.sb; s.bounce[i] = e.bounce;
}

// Paddle move + np.clip(pos, 0, 1)                              envs/my_pong_env_2p.py:118-128
template <typename R> __device__ __forceinline__ R move_paddle(R pos, int act, R ps) {
    using X = Exact<R>;
    if (act == 0) pos = X::sub(pos, ps);
    else if (act == 2) pos = X::add(pos, ps);
    pos = pos < (R)0 ? (R)0 : pos;
    pos = pos > (R)1 ? (R)1 : pos;
    return pos;
}

// collide_sphere_with_moving_plane(vn, vt, u, omega, e, mu, m, R)            envs/physics.py:3-23
template <typename R>
__device__ __forceinline__ void impact(const EnvConsts<R> &c, R vn, R vt, R u, R om, R &vn_post, R &vt_post, R &om_post) {
    using X = Exact<R>;
    vn_post = X::mul(c.neg_e, vn);                                                // (-e)*vn
    const R jn = X::mul(c.m_1pe, X::abs(vn));                                     // (m*(1+e))*|vn|
    R jt = X::mul(c.two_m_7, X::sub(X::add(u, X::mul(c.radius, om)), vt));        // ((2m)/7)*((u+R*w)-vt)
    const R cap = X::mul(c.mu, jn);
    if (!(X::abs(jt) <= cap)) {                                                   // slip
        const R vrel = X::sub(X::sub(vt, u), X::mul(c.radius, om));
        jt = X::mul(-cap, X::sign1(vrel));
    }
    vt_post = X::add(vt, c.mass == (R)1 ? jt : X::div(jt, c.mass));              // jt / 1.0 is jt, bit for bit (uniform branch)
    om_post = X::sub(om, X::div(X::mul(c.radius, jt), c.inertia));
}

// Ball beyond the top (bottom_side = false) or bottom line.  Hit: rigid impact
// (envs/physics.py:3-23), snap to the line, bounce count, speed scaling (:227-232); returns true.
template <typename R>
__device__ __forceinline__ bool paddle_event(const EnvConsts<R> &c, Env<R> &e, R pad, int act, bool bottom_side) {
    using X = Exact<R>;
    const R lo = X::sub(pad, c.hw), hi = X::add(pad, c.hw);
    if (!(lo <= e.x && e.x <= hi)) return false;
    const R u = act == 0 ? -c.ps : (act == 2 ? c.ps : (R)0);
    const R vn = bottom_side ? -e.vy : e.vy;
    R vn_post, vt_post, om_post;
    impact<R>(c, vn, e.vx, u, e.spin, vn_post, vt_post, om_post);
    e.vy = bottom_side ? -vn_post : vn_post;
    e.vx = vt_post;
    e.spin = om_post;
    e.y = bottom_side ? (R)1 : (R)0;
    e.bounce += 1;
    if (e.bounce % c.scale_every == 0) {
        e.vx = X::mul(e.vx, c.scale);
        e.vy = X::mul(e.vy, c.scale);
    }
    return true;
}

// One PongEnv2P.step()                                          envs/my_pong_env_2p.py:116-225
template <typename R> __device__ __forceinline__ int env_step(const EnvConsts<R> &c, Env<R> &e, int aA, int aB) {
    using X = Exact<R>;
    e.top = move_paddle<R>(e.top, aA, c.ps);
    e.bot = move_paddle<R>(e.bot, aB, c.ps);
    if (c.enable_spin) e.vx = X::add(e.vx, X::mul(X::mul(c.magnus, e.spin), e.vy));
    e.x = X::add(e.x, e.vx);
    e.y = X::add(e.y, e.vy);
    if (e.x < (R)0) { e.x = -e.x; e.vx = -e.vx; }
    else if (e.x > (R)1) { e.x = X::sub((R)2, e.x); e.vx = -e.vx; }
    // ONE copy of the paddle event for both lines (:151-186 top, :189-223 bottom): a warp in which some env is beyond the
    // top line and another beyond the bottom line walks the ~150-instruction impact code once, not twice.
    int flags = 0;
    const bool top = e.y < (R)0, bottom = !top && e.y > (R)1;
    if (top || bottom) {
        if (paddle_event<R>(c, e, bottom ? e.bot : e.top, bottom ? aB : aA, bottom)) flags = F_HIT;
        else if (bottom) { e.sa += 1; flags = F_POINT_A | (e.sa >= c.max_score ? F_DONE : 0); }
        else { e.sb += 1; flags = F_POINT_B | (e.sb >= c.max_score ? F_DONE : 0); }
    }
    return flags;
}

// The two 7-D fp32 observations                                 envs/my_pong_env_2p.py:235-257
template <typename R> __device__ __forceinline__ void observe(const Env<R> &e, float (&oa)[7], float (&ob)[7]) {
    using X = Exact<R>;
    const float fx = X::to_f32(e.x), fvx = X::to_f32(e.vx), ft = X::to_f32(e.top), fb = X::to_f32(e.bot),
                fs = X::to_f32(e.spin);
    oa[0] = fx; oa[1] = X::to_f32(X::sub((R)1, e.y)); oa[2] = fvx; oa[3] = X::to_f32(-e.vy);
    oa[4] = ft; oa[5] = fb; oa[6] = fs;
    ob[0] = fx; ob[1] = X::to_f32(e.y); ob[2] = fvx; ob[3] = X::to_f32(e.vy);
    ob[4] = fb; ob[5] = ft; ob[6] = fs;
}

template <typename R> __device__ __forceinline__ void serve_env(Env<R> &e, R vx, R vy, R spin) {   // reset(): :85-111
    e.sa = 0; e.sb = 0; e.bounce = 0;
    e.top = (R)0.5; e.bot = (R)0.5; e.x = (R)0.5; e.y = (R)0.5;
    e.vx = vx; e.vy = vy; e.spin = spin;
}

// ------------------------------------------------------------------------------------ Philox
enum : uint32_t { STREAM_SERVE = 0, STREAM_ACT_A = 1, STREAM_ACT_B = 2 };

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

__device__ __forceinline__ double u53(uint32_t a, uint32_t b) {   // CPython random.random() construction
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}

// Serve from Philox: the reference formula (envs/my_pong_env_2p.py:98-111) in double.
static __device__ __noinline__ void philox_serve(const PPParams &p, uint64_t seed, uint32_t env_id, uint32_t ep,
                                             double &vx, double &vy, double &spin) {
    const uint4 r = philox4x32_10(env_id, ep, STREAM_SERVE, 0u, (uint32_t)seed, (uint32_t)(seed >> 32));
    const uint4 q = philox4x32_10(env_id, ep, STREAM_SERVE, 1u, (uint32_t)seed, (uint32_t)(seed >> 32));
    const double speed = __dadd_rn(p.speed_lo, __dmul_rn(__dsub_rn(p.speed_hi, p.speed_lo), u53(r.x, r.y)));
    const int w = u53(r.z, r.w) < 0.5 ? 0 : 1;
    const double deg = __dadd_rn(p.angle_lo[w], __dmul_rn(__dsub_rn(p.angle_hi[w], p.angle_lo[w]), u53(q.x, q.y)));
    const double rad = __dmul_rn(deg, 3.14159265358979323846 / 180.0);
    double s, c;
    sincos(rad, &s, &c);
    vx = __dmul_rn(speed, c);
    vy = __dmul_rn(speed, s);
    spin = __dadd_rn(p.spin_lo, __dmul_rn(__dsub_rn(p.spin_hi, p.spin_lo), u53(q.z, q.w)));
}

// Serve j of env i from the source (pool row j % depth, or Philox).
template <typename R>
__device__ __forceinline__ void next_serve(const PPParams &p, const PPServeSource &src, int64_t n, int64_t i,
                                           int64_t env_id_base, int ep, R &vx, R &vy, R &spin) {
    const bool queue = src.kind == PP_SERVE_QUEUE;
    if (src.kind == PP_SERVE_POOL || (queue && src.pool_vx != nullptr)) {
        const int64_t j = queue ? (int64_t)ep : (int64_t)(ep % src.depth) * n + i;
        vx = ((const R *)src.pool_vx)[j]; vy = ((const R *)src.pool_vy)[j]; spin = ((const R *)src.pool_spin)[j];
    } else {                     // Philox; a queue without a pool: serve q is (env q % n, episode q / n) of the same stream
        const int64_t env = queue ? (int64_t)ep % n : i;
        const uint32_t episode = queue ? (uint32_t)(ep / n) : (uint32_t)ep;
        double dvx, dvy, ds;
        philox_serve(p, src.seed, (uint32_t)(env_id_base + env), episode, dvx, dvy, ds);
        vx = (R)dvx; vy = (R)dvy; spin = (R)ds;
    }
}

// PP_SERVE_QUEUE: the finishing lanes of a warp (ballot m) claim the next unplayed serves with one atomic.
// Returns the claimed queue index, or INT32_MAX when the queue is exhausted (the env is frozen for good).
__device__ __forceinline__ int claim_serves(bool fin, unsigned m, const PPServeSource &src) {
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(src.queue_head, (unsigned long long)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    const unsigned long long q = base + __popc(m & ((1u << lane) - 1u));
    return (fin && q < (unsigned long long)src.queue_total) ? (int)q : 0x7fffffff;
}

// Warp-aggregated bookkeeping of a lock-step step: per-thread counters stay in registers and are
// flushed once per launch; finished episodes are appended to the log with ballot + popc rank and one
// cursor atomic per warp.
struct Tally {
    unsigned steps = 0, episodes = 0, wins_a = 0, wins_b = 0, pts_a = 0, pts_b = 0, hits = 0;
    unsigned long long len_sum = 0;
    __device__ __forceinline__ void add_flags(int flags) {
        steps += 1;
        pts_a += (flags & F_POINT_A) ? 1u : 0u;
        pts_b += (flags & F_POINT_B) ? 1u : 0u;
        hits += (flags & F_HIT) ? 1u : 0u;
    }
    __device__ __forceinline__ void add(const Tally &o) {
        steps += o.steps; episodes += o.episodes; wins_a += o.wins_a; wins_b += o.wins_b;
        pts_a += o.pts_a; pts_b += o.pts_b; hits += o.hits; len_sum += o.len_sum;
    }
    // `unlogged_episodes`: the episode cursor of a launch WITHOUT a log buffer (log_episode then skips its per-step atomic)
    __device__ __forceinline__ void flush(unsigned long long *counters, unsigned long long *unlogged_episodes = nullptr) {
        unsigned v[7] = {steps, episodes, wins_a, wins_b, pts_a, pts_b, hits};
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            unsigned s = __reduce_add_sync(0xffffffffu, v[k]);
            if ((threadIdx.x & 31) == 0 && s) {
                if (counters) atomicAdd(counters + k, (unsigned long long)s);
                if (k == 1 && unlogged_episodes) atomicAdd(unlogged_episodes, (unsigned long long)s);
            }
        }
        if (!counters) return;
        unsigned long long l = len_sum;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
        if ((threadIdx.x & 31) == 0 && l) atomicAdd(counters + 7, l);
    }
};

// Called by ALL lanes of a warp (done = this lane finished an episode this step, m = ballot of done).
__device__ __forceinline__ void log_episode(bool done, unsigned m, const PPRolloutOut &out, int env_id, int ep_idx, int sa,
                                            int sb, int ep_len) {
    // without a log buffer there is nothing to rank: the cursor is advanced once per launch by Tally::flush instead of by a
    // returning global atomic (a ~500-cycle round trip) on every warp-step in which some episode ends
    if (m == 0 || out.ep_log_count == nullptr || out.ep_log == nullptr) return;
    const int lane = threadIdx.x & 31;
    unsigned long long base = 0;
    if (lane == (__ffs(m) - 1)) base = atomicAdd(out.ep_log_count, (unsigned long long)__popc(m));
    base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
    if (done && out.ep_log) {
        const unsigned long long slot = base + __popc(m & ((1u << lane) - 1u));
        if (slot < (unsigned long long)out.ep_log_cap)
            reinterpret_cast<int4 *>(out.ep_log)[slot] = make_int4(env_id, ep_idx, (sa << 16) | sb, ep_len);
    }
}

}  // namespace pp
