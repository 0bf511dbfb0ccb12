// tools/tmem_layout_probe.cu — which (lane, column) does register r of thread t receive from the 16-lane tcgen05.ld shapes?
// One warp writes value = lane * 1000 + column into 32 lanes x 64 columns with the 32x32b store, then reads it back with
// 16x256b.x1 / 16x128b.x1 / 16x64b.x1 at lane offsets 0 and 16 and prints the mapping.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I pingpong_selfplay_ai_b200/csrc -o /tmp/tmem_probe tools/tmem_layout_probe.cu
#include <cstdio>
#include "tc_ptx.cuh"
using namespace pp;

__global__ void probe(uint32_t *out) {
    __shared__ uint32_t slot;
    if (threadIdx.x < 32) tc::tmem_alloc<64>(&slot);
    tc::tc_fence_before(); __syncthreads(); tc::tc_fence_after();
    const uint32_t tm = slot;
    const int t = threadIdx.x;
    for (int c0 = 0; c0 < 64; c0 += 16) {
        uint32_t v[16];
        for (int j = 0; j < 16; ++j) v[j] = (uint32_t)(t * 1000 + c0 + j);
        tc::tmem_st16(tm + c0, v);
    }
    tc::tmem_st_wait();
    tc::tc_fence_before(); __syncwarp(); tc::tc_fence_after();
    for (int lo = 0; lo < 2; ++lo) {
        const uint32_t a = tm + ((uint32_t)(lo * 16) << 16);
        uint32_t r[4];
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a) : "memory");
        tc::tmem_ld_wait();
        for (int j = 0; j < 4; ++j) out[((0 * 2 + lo) * 32 + t) * 4 + j] = r[j];
        asm volatile("tcgen05.ld.sync.aligned.16x128b.x1.b32 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(a) : "memory");
        tc::tmem_ld_wait();
        r[2] = r[3] = 0xffffffffu;
        for (int j = 0; j < 4; ++j) out[((1 * 2 + lo) * 32 + t) * 4 + j] = r[j];
        asm volatile("tcgen05.ld.sync.aligned.16x64b.x1.b32 {%0}, [%1];" : "=r"(r[0]) : "r"(a) : "memory");
        tc::tmem_ld_wait();
        r[1] = 0xffffffffu;
        for (int j = 0; j < 4; ++j) out[((2 * 2 + lo) * 32 + t) * 4 + j] = r[j];
        // x2 of 16x256b: where do the second four registers come from?
        uint32_t q[8];
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]) : "r"(a) : "memory");
        tc::tmem_ld_wait();
        for (int j = 0; j < 4; ++j) out[((3 * 2 + lo) * 32 + t) * 4 + j] = q[4 + j];
    }
    tc::tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) tc::tmem_dealloc<64>(tm);
}

int main() {
    uint32_t *d, h[4 * 2 * 32 * 4];
    cudaMalloc(&d, sizeof h);
    probe<<<1, 32>>>(d);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
    cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    const char *names[4] = {"16x256b.x1", "16x128b.x1", "16x64b.x1", "16x256b.x2 regs 4..7"};
    for (int s = 0; s < 4; ++s)
        for (int lo = 0; lo < 2; ++lo) {
            printf("%s lane offset %d: thread -> (lane,col) per register\n", names[s], lo * 16);
            for (int t = 0; t < 32; ++t) {
                printf("  t%2d:", t);
                for (int j = 0; j < 4; ++j) {
                    const uint32_t v = h[((s * 2 + lo) * 32 + t) * 4 + j];
                    if (v == 0xffffffffu) continue;
                    printf(" (%2u,%2u)", v / 1000, v % 1000);
                }
                printf("\n");
            }
        }
    return 0;
}
