// replay_kernels.cu — K3: append transitions (oB, aB, rB, nB, done) to the replay ring
// (memory.push of scripts/train_iterative.py:56-63,243), 62 bytes per row over five arrays.
//
// *head advances by the number of rows WRITTEN.  Rows are compacted per warp: ballot of the valid lanes, one atomicAdd on the ring cursor per warp,
// rank by popc.  A fully valid warp whose 32 slots do not wrap copies its 2 x 896 B of observations
// with lane-contiguous (coalesced) accesses; otherwise each lane copies its own row.
#include "pp_device.cuh"
#include "pp_host.h"

namespace pp {

__global__ void __launch_bounds__(256)
replay_scatter_kernel(int64_t n, const PPReplayRing ring, const float *__restrict__ obs, const uint8_t *__restrict__ act,
                      const float *__restrict__ rew, const float *__restrict__ next_obs, const uint8_t *__restrict__ done,
                      const uint8_t *__restrict__ valid) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int lane = threadIdx.x & 31;
    // a batch larger than the ring: only its last `capacity` rows can survive sequential pushes, and skipping
    // the others keeps concurrently written slots distinct (no torn rows)
    const bool v = i < n && i >= n - ring.capacity && (valid == nullptr || valid[i] != 0);
    const unsigned m = __ballot_sync(0xffffffffu, v);
    if (m == 0) return;
    const int leader = __ffs(m) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(ring.head, (unsigned long long)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    const unsigned long long cap = (unsigned long long)ring.capacity;
    const int64_t first = (int64_t)(base % cap);
    if (m == 0xffffffffu && first + 32 <= ring.capacity) {
        const int64_t src0 = (i - lane) * 7, dst0 = first * 7;
#pragma unroll
        for (int r = 0; r < 7; ++r) {
            ring.obs[dst0 + r * 32 + lane] = obs[src0 + r * 32 + lane];
            ring.next_obs[dst0 + r * 32 + lane] = next_obs[src0 + r * 32 + lane];
        }
        ring.act[first + lane] = act[i];
        ring.rew[first + lane] = rew[i];
        ring.done[first + lane] = done[i];
        return;
    }
    if (!v) return;
    const int64_t slot = (int64_t)((base + __popc(m & ((1u << lane) - 1u))) % cap);
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        ring.obs[slot * 7 + k] = obs[i * 7 + k];
        ring.next_obs[slot * 7 + k] = next_obs[i * 7 + k];
    }
    ring.act[slot] = act[i];
    ring.rew[slot] = rew[i];
    ring.done[slot] = done[i];
}

int replay_scatter_launch(int64_t n, const PPReplayRing &ring, const float *obs, const uint8_t *act, const float *rew,
                          const float *next_obs, const uint8_t *done, const uint8_t *valid, cudaStream_t stream) {
    const unsigned blocks = (unsigned)((n + 255) / 256);
    replay_scatter_kernel<<<blocks, 256, 0, stream>>>(n, ring, obs, act, rew, next_obs, done, valid);
    return (int)cudaGetLastError();
}

}  // namespace pp
