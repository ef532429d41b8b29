// tma_bench.cu — how fast can one SM's TMA unit retire cp.async.bulk copies as a function of copy size?
// One CTA per SM, one warp; a ring of S stages of STAGE bytes each; a stage is filled by STAGE / sz copies of sz bytes
// (issued by the 32 lanes in parallel, like the step kernel's producer) and refilled as soon as it has landed.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_bench tma_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(128, 1) k(const unsigned char *src, size_t bytes_per_cta, uint32_t sz, uint32_t stage, uint32_t S, int lanes, uint32_t W) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *full = (uint64_t *)(smem + (size_t)S * stage);
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < S; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&full[s])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp >= W) return;
    const unsigned char *base = src + (size_t)blockIdx.x * bytes_per_cta;
    const uint32_t n_fill = (uint32_t)(bytes_per_cta / stage), per = stage / sz;
    for (uint32_t f = warp; f < n_fill + S; f += W) {
        const uint32_t s = f % S, ph = (f / S) & 1u;
        if (f >= S) {   // wait for the previous fill of this stage
            uint32_t done = 0;
            while (!done) asm volatile("{ .reg .pred P; mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2; selp.u32 %0,1,0,P; }" : "=r"(done) : "r"(s32(&full[s])), "r"(ph ^ 1u) : "memory");
        }
        if (f < n_fill) {
            if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[s])), "r"(stage) : "memory");
            __syncwarp();
            for (uint32_t c = lane; c < per; c += lanes) {
                if ((int)lane < lanes)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(smem + (size_t)s * stage + (size_t)c * sz)),
                                 "l"(base + (size_t)f * stage + (size_t)c * sz), "r"(sz), "r"(s32(&full[s])) : "memory");
            }
        }
    }
}
int main() {
    const int sms = 148;
    const uint32_t stage = 40960, S = 4;
    const size_t per_cta = (size_t)stage * 256;   // 10 MB per CTA, 1.55 GB total
    unsigned char *d;
    cudaMalloc(&d, per_cta * sms);
    cudaMemset(d, 1, per_cta * sms);
    const size_t smem = (size_t)S * stage + 64;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (uint32_t W : {1u, 2u, 4u})
      for (int lanes : {32, 8})
        for (uint32_t sz : {512u, 1024u, 2048u, 4096u}) {
            k<<<sms, 128, smem>>>(d, per_cta, sz, stage, S, lanes, W);
            cudaEventRecord(e0);
            for (int r = 0; r < 5; ++r) k<<<sms, 128, smem>>>(d, per_cta, sz, stage, S, lanes, W);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
            printf("warps %u lanes %2d copy %6u B: %7.1f us, %7.1f GB/s, %5.1f ns per copy per SM (%s)\n", W, lanes, sz, ms * 1e3, per_cta * sms / ms / 1e6,
                   ms * 1e6 / (per_cta / sz), cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
