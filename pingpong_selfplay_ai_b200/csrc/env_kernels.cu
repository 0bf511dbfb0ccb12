// env_kernels.cu — K1: PongEnv2P.step / reset / observe for n lock-step environments.
//
//   step_kernel      one step, everything materialised (the reference's step() signature): HBM-bound,
//                    203 B/env-step in fp64 mode, 147 B in fp32 mode.  SoA state is
//                    read and written with 16-byte vector accesses (2 fp64 / 4 fp32 envs per thread);
//                    the [n][7] fp32 observation rows are staged through shared memory so that the
//                    global stores are contiguous 16-byte vectors.
//   rollout_kernel   k steps per launch from an injected action stream; state lives in registers,
//                    finished episodes are compacted into the episode log (ballot + rank, one atomic
//                    per warp) and the env is re-served in place.
#include "pp_device.cuh"
#include "pp_host.h"

namespace pp {

// ------------------------------------------------------------------ vector helpers
template <typename T, int V> struct Vec;
template <> struct Vec<double, 2> { using type = double2; };
template <> struct Vec<double, 1> { using type = double; };
template <> struct Vec<float, 4> { using type = float4; };
template <> struct Vec<float, 2> { using type = float2; };
template <> struct Vec<float, 1> { using type = float; };
template <> struct Vec<int32_t, 4> { using type = int4; };
template <> struct Vec<int32_t, 2> { using type = int2; };
template <> struct Vec<int32_t, 1> { using type = int32_t; };
template <> struct Vec<uint8_t, 4> { using type = uchar4; };
template <> struct Vec<uint8_t, 2> { using type = uchar2; };
template <> struct Vec<uint8_t, 1> { using type = uint8_t; };

template <typename T, int V> __device__ __forceinline__ void vload(T (&dst)[V], const T *src) {
    using VT = typename Vec<T, V>::type;
    union { VT v; T a[V]; } u;
    u.v = *reinterpret_cast<const VT *>(src);
#pragma unroll
    for (int i = 0; i < V; ++i) dst[i] = u.a[i];
}
template <typename T, int V> __device__ __forceinline__ void vstore(T *dst, const T (&src)[V]) {
    using VT = typename Vec<T, V>::type;
    union { VT v; T a[V]; } u;
#pragma unroll
    for (int i = 0; i < V; ++i) u.a[i] = src[i];
    *reinterpret_cast<VT *>(dst) = u.v;
}

constexpr int STEP_TILE = 512;          // envs per CTA: 2 x 14 KB of staged observations

// One reference step() per env, V envs per thread, STEP_TILE / V threads per CTA.
template <typename R, int V>
__global__ void __launch_bounds__(STEP_TILE / V)
step_kernel(const PPParams params, const PPEnvState st, int64_t n, const uint8_t *__restrict__ act_a,
            const uint8_t *__restrict__ act_b, float *__restrict__ obs_a, float *__restrict__ obs_b,
            float *__restrict__ rew_a, float *__restrict__ rew_b, uint8_t *__restrict__ done) {
    constexpr int TILE = STEP_TILE, STEP_THREADS = STEP_TILE / V;
    __shared__ __align__(16) float s_obs[2][TILE * 7];
    const EnvConsts<R> c(params);
    const StatePtrs<R> s(st);
    const int64_t tile0 = (int64_t)blockIdx.x * TILE;
    const int64_t i0 = tile0 + (int64_t)threadIdx.x * V;
    const bool full = i0 + V <= n;          // V > 1 is only launched with n % V == 0, so a thread is all-in or all-out

    if (full) {
        R x[V], y[V], vx[V], vy[V], sp[V], top[V], bot[V];
        int32_t sa[V], sb[V], bc[V];
        uint8_t aa[V], ab[V];
        vload<R, V>(x, s.x + i0); vload<R, V>(y, s.y + i0); vload<R, V>(vx, s.vx + i0); vload<R, V>(vy, s.vy + i0);
        vload<R, V>(sp, s.spin + i0); vload<R, V>(top, s.top + i0); vload<R, V>(bot, s.bot + i0);
        vload<int32_t, V>(sa, s.sa + i0); vload<int32_t, V>(sb, s.sb + i0); vload<int32_t, V>(bc, s.bounce + i0);
        vload<uint8_t, V>(aa, act_a + i0); vload<uint8_t, V>(ab, act_b + i0);
        float ra[V], rb[V];
        uint8_t dn[V];
#pragma unroll
        for (int v = 0; v < V; ++v) {
            Env<R> e{x[v], y[v], vx[v], vy[v], sp[v], top[v], bot[v], sa[v], sb[v], bc[v]};
            const int f = env_step<R>(c, e, aa[v], ab[v]);
            x[v] = e.x; y[v] = e.y; vx[v] = e.vx; vy[v] = e.vy; sp[v] = e.spin; top[v] = e.top; bot[v] = e.bot;
            sa[v] = e.sa; sb[v] = e.sb; bc[v] = e.bounce;
            ra[v] = (f & F_POINT_A) ? 1.0f : ((f & F_POINT_B) ? -1.0f : 0.0f);
            rb[v] = -ra[v] + 0.0f;                       // (-1,+1) / (+1,-1) / (0,0); "+0.0f" turns -0 into +0
            dn[v] = (uint8_t)(f & F_DONE);
            float oa[7], ob[7];
            observe<R>(e, oa, ob);
            const int l = (threadIdx.x * V + v) * 7;     // stride 7 words: conflict-free across a warp for V = 1
#pragma unroll
            for (int k = 0; k < 7; ++k) { s_obs[0][l + k] = oa[k]; s_obs[1][l + k] = ob[k]; }
        }
        vstore<R, V>(s.x + i0, x); vstore<R, V>(s.y + i0, y); vstore<R, V>(s.vx + i0, vx); vstore<R, V>(s.vy + i0, vy);
        vstore<R, V>(s.spin + i0, sp); vstore<R, V>(s.top + i0, top); vstore<R, V>(s.bot + i0, bot);
        vstore<int32_t, V>(s.sa + i0, sa); vstore<int32_t, V>(s.sb + i0, sb); vstore<int32_t, V>(s.bounce + i0, bc);
        vstore<float, V>(rew_a + i0, ra); vstore<float, V>(rew_b + i0, rb);
        vstore<uint8_t, V>(done + i0, dn);
    }
    __syncthreads();
    // obs rows of this tile are contiguous in global memory: [tile0*7, tile0*7 + rows*7)
    const int64_t rows = (n - tile0) < TILE ? (n - tile0) : TILE;
    const int words = (int)rows * 7;
    float *ga = obs_a + tile0 * 7, *gb = obs_b + tile0 * 7;
    if (rows == TILE) {                       // TILE*7 words is a multiple of 4 and tile0*28 B of 16 B
        const float4 *sa4 = reinterpret_cast<const float4 *>(s_obs[0]);
        const float4 *sb4 = reinterpret_cast<const float4 *>(s_obs[1]);
#pragma unroll
        for (int w = threadIdx.x; w < TILE * 7 / 4; w += STEP_THREADS) {
            reinterpret_cast<float4 *>(ga)[w] = sa4[w];
            reinterpret_cast<float4 *>(gb)[w] = sb4[w];
        }
    } else {
        for (int w = threadIdx.x; w < words; w += STEP_THREADS) { ga[w] = s_obs[0][w]; gb[w] = s_obs[1][w]; }
    }
}

template <typename R>
__global__ void __launch_bounds__(256)
observe_kernel(const PPEnvState st, int64_t n, float *__restrict__ obs_a, float *__restrict__ obs_b) {
    __shared__ __align__(16) float s_obs[2][256 * 7];
    const StatePtrs<R> s(st);
    const int64_t tile0 = (int64_t)blockIdx.x * 256;
    const int64_t i = tile0 + threadIdx.x;
    if (i < n) {
        Env<R> e;
        e.x = s.x[i]; e.y = s.y[i]; e.vx = s.vx[i]; e.vy = s.vy[i]; e.spin = s.spin[i]; e.top = s.top[i]; e.bot = s.bot[i];
        float oa[7], ob[7];
        observe<R>(e, oa, ob);
#pragma unroll
        for (int k = 0; k < 7; ++k) { s_obs[0][threadIdx.x * 7 + k] = oa[k]; s_obs[1][threadIdx.x * 7 + k] = ob[k]; }
    }
    __syncthreads();
    const int64_t rows = (n - tile0) < 256 ? (n - tile0) : 256;
    const int words = (int)rows * 7;
    for (int w = threadIdx.x; w < words; w += 256) {
        obs_a[tile0 * 7 + w] = s_obs[0][w];
        obs_b[tile0 * 7 + w] = s_obs[1][w];
    }
}

// reset() for masked envs: explicit serves (SERVE_ARRAYS) or a PPServeSource.
template <typename R, bool FROM_SOURCE>
__global__ void __launch_bounds__(256)
reset_kernel(const PPParams params, const PPEnvState st, int64_t n, const uint8_t *__restrict__ mask,
             const R *__restrict__ vx, const R *__restrict__ vy, const R *__restrict__ spin,
             const PPServeSource src, int64_t env_id_base, int advance) {
    const StatePtrs<R> s(st);
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n || (mask && !mask[i])) return;
    Env<R> e;
    R svx, svy, ssp;
    if (FROM_SOURCE) {
        int ep = s.ep_idx[i] + (advance ? 1 : 0);
        if (src.kind == PP_SERVE_QUEUE) ep = i < src.queue_total ? (int)i : 0x7fffffff;   // env i starts on serve i
        s.ep_idx[i] = ep;
        if (ep == 0x7fffffff) return;
        next_serve<R>(params, src, n, i, env_id_base, ep, svx, svy, ssp);
    } else {
        svx = vx[i]; svy = vy[i]; ssp = spin[i];
    }
    serve_env<R>(e, svx, svy, ssp);
    store_env<R>(s, i, e);
    if (s.ep_len) s.ep_len[i] = 0;
}

constexpr int ROLL_THREADS = 128;
constexpr int ROLL_PREFETCH = 8;

// k lock-step steps from an action stream actions[k][n][2]; one env per thread, state in registers.
template <typename R>
__global__ void __launch_bounds__(ROLL_THREADS)
rollout_kernel(const PPParams params, const PPEnvState st, int64_t n, int64_t k_steps,
               const uint8_t *__restrict__ actions, const PPServeSource src, int32_t quota, int64_t env_id_base,
               const PPRolloutOut out) {
    const EnvConsts<R> c(params);
    const StatePtrs<R> s(st);
    const int64_t i = (int64_t)blockIdx.x * ROLL_THREADS + threadIdx.x;
    const bool valid = i < n;
    const int64_t ic = valid ? i : 0;
    Env<R> e = load_env<R>(s, ic);
    int ep_idx = s.ep_idx[ic], ep_len = s.ep_len[ic];
    Tally tally;
    const uchar2 *acts = reinterpret_cast<const uchar2 *>(actions);

    for (int64_t t0 = 0; t0 < k_steps; t0 += ROLL_PREFETCH) {
        // all action loads of the chunk are in flight before the first step; packed 16 bits per step
        unsigned long long pk[2] = {0ull, 0ull};
#pragma unroll
        for (int u = 0; u < ROLL_PREFETCH; ++u) {
            const uchar2 a = (t0 + u < k_steps) ? __ldg(acts + (t0 + u) * n + ic) : make_uchar2(1, 1);
            pk[u >> 2] |= (unsigned long long)(a.x | (a.y << 8)) << ((u & 3) * 16);
        }
#pragma unroll 1
        for (int u = 0; u < ROLL_PREFETCH; ++u) {
            const int64_t t = t0 + u;
            if (t >= k_steps) break;
            const unsigned aw = (unsigned)((u < 4 ? pk[0] : pk[1]) >> ((u & 3) * 16));
            const int act_a = aw & 0xff, act_b = (aw >> 8) & 0xff;
            const bool active = valid && !(quota > 0 && ep_idx >= quota);
            int flags = 0;
            if (active) {
                flags = env_step<R>(c, e, act_a, act_b);
                ep_len += 1;
                tally.add_flags(flags);
            }
            if (valid && out.trace_real) {
                R *tr = (R *)out.trace_real + t * 7 * n;
                tr[0 * n + i] = e.x; tr[1 * n + i] = e.y; tr[2 * n + i] = e.vx; tr[3 * n + i] = e.vy;
                tr[4 * n + i] = e.spin; tr[5 * n + i] = e.top; tr[6 * n + i] = e.bot;
            }
            if (valid && out.trace_int) {
                int32_t *ti = out.trace_int + t * 4 * n;
                ti[0 * n + i] = e.sa; ti[1 * n + i] = e.sb; ti[2 * n + i] = e.bounce; ti[3 * n + i] = flags;
            }
            const bool fin = (flags & F_DONE) != 0;
            log_episode(fin, __ballot_sync(0xffffffffu, fin), out, (int)(env_id_base + i), ep_idx, e.sa, e.sb, ep_len);
            if (fin) {
                tally.episodes += 1;
                if (e.sa > e.sb) tally.wins_a += 1; else tally.wins_b += 1;
                tally.len_sum += (unsigned)ep_len;
                ep_idx += 1;
                if (!(quota > 0 && ep_idx >= quota)) {
                    R svx, svy, ssp;
                    next_serve<R>(params, src, n, i, env_id_base, ep_idx, svx, svy, ssp);
                    serve_env<R>(e, svx, svy, ssp);
                    ep_len = 0;
                }
            }
        }
    }
    if (valid) {
        store_env<R>(s, i, e);
        s.ep_idx[i] = ep_idx;
        s.ep_len[i] = ep_len;
    }
    tally.flush(out.counters, out.ep_log ? nullptr : out.ep_log_count);
}

// collide_sphere_with_moving_plane for n independent impacts (envs/physics.py:3-23): the arithmetic of paddle_event,
// exposed so that the reference's `envs.physics` import has a device-backed drop-in and the known answers can be
// checked through the C ABI.
template <typename R>
__global__ void collide_kernel(const PPParams params, int64_t n, const R *__restrict__ vn, const R *__restrict__ vt,
                               const R *__restrict__ u, const R *__restrict__ om, R *__restrict__ vn_out,
                               R *__restrict__ vt_out, R *__restrict__ om_out) {
    const EnvConsts<R> c(params);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        impact<R>(c, vn[i], vt[i], u[i], om[i], vn_out[i], vt_out[i], om_out[i]);
}

// ------------------------------------------------------------------ host launchers
static bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <typename R, int V>
static int launch_step(const PPParams &p, const PPEnvState &st, int64_t n, const uint8_t *aa, const uint8_t *ab,
                       float *oa, float *ob, float *ra, float *rb, uint8_t *dn, cudaStream_t stream) {
    const int64_t blocks = (n + STEP_TILE - 1) / STEP_TILE;
    step_kernel<R, V><<<(unsigned)blocks, STEP_TILE / V, 0, stream>>>(p, st, n, aa, ab, oa, ob, ra, rb, dn);
    return (int)cudaGetLastError();
}

int env_step_launch(int mode, int64_t n, const PPParams &p, const PPEnvState &st, const uint8_t *aa, const uint8_t *ab,
                    float *oa, float *ob, float *ra, float *rb, uint8_t *dn, cudaStream_t stream) {
    const void *ptrs[] = {st.ball_x, st.ball_y, st.ball_vx, st.ball_vy, st.spin, st.top_paddle_x, st.bottom_paddle_x,
                          st.score_a, st.score_b, st.bounce_count, oa, ob, ra, rb};
    bool vec_ok = true;
    for (const void *q : ptrs) vec_ok = vec_ok && aligned16(q);
    if (mode == PP_MODE_F64) {
        vec_ok = vec_ok && (n % 2 == 0) && ((reinterpret_cast<uintptr_t>(aa) | reinterpret_cast<uintptr_t>(ab) |
                                             reinterpret_cast<uintptr_t>(dn)) % 2 == 0);
        return vec_ok ? launch_step<double, 2>(p, st, n, aa, ab, oa, ob, ra, rb, dn, stream)
                      : launch_step<double, 1>(p, st, n, aa, ab, oa, ob, ra, rb, dn, stream);
    }
    vec_ok = vec_ok && (n % 4 == 0) && ((reinterpret_cast<uintptr_t>(aa) | reinterpret_cast<uintptr_t>(ab) |
                                         reinterpret_cast<uintptr_t>(dn)) % 4 == 0);
    return vec_ok ? launch_step<float, 4>(p, st, n, aa, ab, oa, ob, ra, rb, dn, stream)
                  : launch_step<float, 1>(p, st, n, aa, ab, oa, ob, ra, rb, dn, stream);
}

int env_observe_launch(int mode, int64_t n, const PPEnvState &st, float *oa, float *ob, cudaStream_t stream) {
    const unsigned blocks = (unsigned)((n + 255) / 256);
    if (mode == PP_MODE_F64) observe_kernel<double><<<blocks, 256, 0, stream>>>(st, n, oa, ob);
    else observe_kernel<float><<<blocks, 256, 0, stream>>>(st, n, oa, ob);
    return (int)cudaGetLastError();
}

int env_serve_launch(int mode, int64_t n, const PPEnvState &st, const uint8_t *mask, const void *vx, const void *vy,
                     const void *spin, cudaStream_t stream) {
    const unsigned blocks = (unsigned)((n + 255) / 256);
    PPParams p{};
    PPServeSource src{};
    if (mode == PP_MODE_F64)
        reset_kernel<double, false><<<blocks, 256, 0, stream>>>(p, st, n, mask, (const double *)vx, (const double *)vy,
                                                                (const double *)spin, src, 0, 0);
    else
        reset_kernel<float, false><<<blocks, 256, 0, stream>>>(p, st, n, mask, (const float *)vx, (const float *)vy,
                                                               (const float *)spin, src, 0, 0);
    return (int)cudaGetLastError();
}

int env_reset_launch(int mode, int64_t n, const PPParams &p, const PPEnvState &st, const uint8_t *mask,
                     const PPServeSource &src, int64_t env_id_base, int advance, cudaStream_t stream) {
    const unsigned blocks = (unsigned)((n + 255) / 256);
    if (mode == PP_MODE_F64)
        reset_kernel<double, true><<<blocks, 256, 0, stream>>>(p, st, n, mask, nullptr, nullptr, nullptr, src, env_id_base, advance);
    else
        reset_kernel<float, true><<<blocks, 256, 0, stream>>>(p, st, n, mask, nullptr, nullptr, nullptr, src, env_id_base, advance);
    return (int)cudaGetLastError();
}

int collide_launch(int mode, int64_t n, const PPParams &p, const void *vn, const void *vt, const void *u, const void *om,
                   void *vn_out, void *vt_out, void *om_out, cudaStream_t stream) {
    const int64_t want = (n + 255) / 256;
    const unsigned blocks = (unsigned)(want < 148 * 8 ? want : 148 * 8);
    if (mode == PP_MODE_F64)
        collide_kernel<double><<<blocks, 256, 0, stream>>>(p, n, (const double *)vn, (const double *)vt, (const double *)u,
                                                           (const double *)om, (double *)vn_out, (double *)vt_out, (double *)om_out);
    else
        collide_kernel<float><<<blocks, 256, 0, stream>>>(p, n, (const float *)vn, (const float *)vt, (const float *)u,
                                                          (const float *)om, (float *)vn_out, (float *)vt_out, (float *)om_out);
    return (int)cudaGetLastError();
}

int env_rollout_launch(int mode, int64_t n, int64_t k, const PPParams &p, const PPEnvState &st, const uint8_t *actions,
                       const PPServeSource &src, int32_t quota, int64_t env_id_base, const PPRolloutOut &out,
                       cudaStream_t stream) {
    const unsigned blocks = (unsigned)((n + ROLL_THREADS - 1) / ROLL_THREADS);
    if (mode == PP_MODE_F64)
        rollout_kernel<double><<<blocks, ROLL_THREADS, 0, stream>>>(p, st, n, k, actions, src, quota, env_id_base, out);
    else
        rollout_kernel<float><<<blocks, ROLL_THREADS, 0, stream>>>(p, st, n, k, actions, src, quota, env_id_base, out);
    return (int)cudaGetLastError();
}

}  // namespace pp
