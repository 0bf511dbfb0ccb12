// tools/mma_rate.cu — micro-benchmark: cycles per tcgen05.mma (M = 128, kind::f16, K = 16) when ONE thread issues a long
// accumulate chain, for N in {16, 64, 128, 256}, A from shared memory (SS) or tensor memory (TS).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I pingpong_selfplay_ai_b200/csrc -o /tmp/mma_rate tools/mma_rate.cu
#include <cstdio>
#include "tc_ptx.cuh"
using namespace pp;

// STAGE > 0: like a weight-ring stage of STAGE MMAs: a try_wait on a barrier whose phase completed long ago, the MMAs, a commit
template <int N, bool TS, bool COMMIT_EACH, int NOISE = 0, int STAGE = 0>      // NOISE: 1 = the other warps hammer TMEM loads, 2 = MUFU
__global__ void rate_kernel(long long *out, int iters) {
    __shared__ volatile int stop;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar, bar2, bar3;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0;
    if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::mbar_init(&bar2, 1); tc::mbar_init(&bar3, 1); tc::fence_mbar_init(); }
    if (threadIdx.x < 32) tc::tmem_alloc<512>(&slot);
    tc::tc_fence_before(); __syncthreads(); tc::tc_fence_after();
    tc::fence_proxy_async();
    const uint32_t tm = slot;
    if (threadIdx.x == 0) stop = 0;
    __syncthreads();
    if (NOISE && threadIdx.x >= 32) {
        const uint32_t lane_addr = (uint32_t)(((threadIdx.x >> 5) & 3) * 32) << 16;
        float acc = 0.f;
        while (!stop) {
            if (NOISE == 1) {
                uint32_t r[16];
                tc::tmem_ld16(tm + lane_addr + 384, r);
                tc::tmem_ld_wait();
                acc += __uint_as_float(r[3]);
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) acc = __expf(acc) + 1.0f;
            }
        }
        if (acc == 123.456f) out[1] = 0;
    }
    if (threadIdx.x == 0) {
        const uint32_t a_sm = tc::smem_u32(smem), b_sm = tc::smem_u32(smem + 16384);
        const uint64_t ad = tc::smem_desc(a_sm, 128 * 16, 128), bd = tc::smem_desc(b_sm, N * 16, 128);
        const uint32_t idesc = tc::idesc_f16(128, N);
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            if (STAGE > 0) {
#pragma unroll
                for (int st = 0; st < 8 / STAGE; ++st) {
                    tc::mbar_wait(&bar3, 1);               // the phase before the first: complete since initialisation
#pragma unroll
                    for (int j = 0; j < STAGE; ++j) tc::umma_f16_ts(tm + 256, tm + (st * STAGE + j) * 8, bd, idesc, true);
                    tc::umma_commit(&bar2);
                }
                continue;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (TS) tc::umma_f16_ts(tm + 256, tm + j * 8, bd, idesc, true);
                else tc::umma_f16(tm + 256, ad, bd, idesc, true);
            }
            if (COMMIT_EACH) tc::umma_commit(&bar2);      // like releasing a ring slot after every 8 MMAs
        }
        long long t1 = clock64();
        tc::umma_commit(&bar);
        tc::mbar_wait(&bar, 0);
        long long t2 = clock64();
        out[0] = t1 - t0; out[1] = t2 - t0;
        stop = 1;
    }
    tc::tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) tc::tmem_dealloc<512>(tm);
}

template <int N, bool TS, bool CE = false, int NOISE = 0, int THREADS = 128, int STAGE = 0> void run(const char *name) {
    long long *d, h[2];
    cudaMalloc(&d, 16);
    cudaFuncSetAttribute(rate_kernel<N, TS, CE, NOISE, STAGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int iters = 200;
    rate_kernel<N, TS, CE, NOISE, STAGE><<<1, THREADS, 64 * 1024>>>(d, iters);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("%-10s issue %.1f cyc/MMA, complete %.1f cyc/MMA (%s)\n", name, (double)h[0] / (iters * 8), (double)h[1] / (iters * 8), cudaGetErrorString(e));
    cudaFree(d);
}

int main() {
    run<16, false>("N16 SS"); run<16, true>("N16 TS");
    run<64, false>("N64 SS"); run<64, true>("N64 TS");
    run<128, false>("N128 SS"); run<128, true>("N128 TS");
    run<256, false>("N256 SS"); run<256, true>("N256 TS");
    run<128, true, true>("N128 TS + commit/8"); run<64, true, true>("N64 TS + commit/8");
    run<128, true, false, 0, 128, 8>("N128 TS, stages of 8 (try_wait + 8 MMAs + commit)");
    run<128, true, false, 0, 128, 4>("N128 TS, stages of 4");
    run<128, true, false, 1, 288, 4>("N128 TS, stages of 4 + 8 warps tcgen05.ld");
    run<128, true, false, 2, 288, 4>("N128 TS, stages of 4 + 8 warps MUFU");
    run<128, false, false, 1, 288>("N128 SS + 8 warps tcgen05.ld");
    run<128, true, false, 1, 288>("N128 TS + 8 warps tcgen05.ld"); run<128, true, false, 2, 288>("N128 TS + 8 warps MUFU");
    run<64, true, false, 1, 512>("N64 TS + 15 warps tcgen05.ld"); run<64, true, false, 2, 512>("N64 TS + 15 warps MUFU");
    return 0;
}
