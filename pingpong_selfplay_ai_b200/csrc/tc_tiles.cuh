// tc_tiles.cuh — operand-tile helpers shared by the tensor-core kernels (QNet and QNetRNN): the observation tile X,
// fp16 pair unpacking, and the dueling read-out.  A tiles are K-major, no swizzle, stored [K/8][128 rows][8 halves].
#pragma once
#include <cuda_fp16.h>

#include "pp_policy.cuh"
#include "tc_ptx.cuh"

namespace pp {

constexpr uint32_t A_LBO = 128 * 16;           // next K chunk of a 128-row A tile
constexpr uint32_t SBO = 128;                  // next 8-row core matrix

__device__ __forceinline__ float h2f(uint32_t packed, int hi) {
    const __half2 v = *reinterpret_cast<const __half2 *>(&packed);
    return hi ? __high2float(v) : __low2float(v);
}

// this thread's row of X = [obs_hi(7) 1 | obs_lo(7) 1]
__device__ __forceinline__ void write_x_row(uint8_t *x, int row, const float (&o)[7]) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float a = o[2 * j], b = j < 3 ? o[2 * j + 1] : 1.0f;
        hi[j] = tc::pack_f16x2<false>(a, b);
        const float ra = a - h2f(hi[j], 0), rb = j < 3 ? b - h2f(hi[j], 1) : 1.0f;
        lo[j] = tc::pack_f16x2<false>(ra, rb);
    }
    *reinterpret_cast<uint4 *>(x + row * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4 *>(x + A_LBO + row * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

__device__ __forceinline__ void dueling_q(uint32_t taddr, float (&q)[3]) {      // V + (A - mean(A))  models/qnet.py:75
    uint32_t r[4];
    tc::tmem_ld4(taddr, r);
    tc::tmem_ld_wait();
    const float v = __uint_as_float(r[0]), a0 = __uint_as_float(r[1]), a1 = __uint_as_float(r[2]), a2 = __uint_as_float(r[3]);
    const float mean = __fdiv_rn(__fadd_rn(__fadd_rn(a0, a1), a2), 3.0f);
    q[0] = __fadd_rn(v, __fsub_rn(a0, mean));
    q[1] = __fadd_rn(v, __fsub_rn(a1, mean));
    q[2] = __fadd_rn(v, __fsub_rn(a2, mean));
}

}  // namespace pp
