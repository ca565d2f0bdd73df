// Microbenchmark: cycles per tcgen05.mma.kind::f16 (M=128, K=16) as a function of N and of where A lives
// (TMEM = "TS" form used by the fused kernel, shared memory = "SS" form).  One CTA, one issuing thread,
// batches of 42 MMAs followed by a commit; data are irrelevant (whatever is in TMEM / zeroed smem).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_rate tools/mma_rate.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return uint32_t(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo) {
    return uint64_t((saddr >> 4) & 0x3fff) | (uint64_t((lbo >> 4) & 0x3fff) << 16) | (uint64_t(128 >> 4) << 32) | (uint64_t(1) << 46);
}

template <int N, bool TS, int ND = 1>
__device__ void run(uint32_t tmem, uint8_t *smem, uint64_t *bar, uint32_t &parity, long long *out) {
    const uint32_t idesc = (1u << 4) | (uint32_t(N >> 3) << 17) | (uint32_t(128 >> 4) << 24);
    const uint32_t lbo_b = (N / 8) * 128, lbo_a = (128 / 8) * 128;
    const uint32_t sB = smem_u32(smem), sA = smem_u32(smem + 65536);
    long long t_issue = 0, t_total = 0;
    constexpr int REPS = 100, BATCH = 42;
    for (int rep = 0; rep < REPS; rep++) {
        const long long c0 = clock64();
#pragma unroll
        for (int j = 0; j < BATCH; j++) {
            const uint64_t bd = desc(sB + uint32_t(2 * (j % 7)) * lbo_b, lbo_b);
            if (TS) {
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem + (j % ND) * N),
                             "r"(tmem + 256 + (j % 7) * 8), "l"(bd), "r"(idesc), "r"(1));
            } else {
                const uint64_t ad = desc(sA + uint32_t(2 * (j % 7)) * lbo_a, lbo_a);
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem), "l"(ad),
                             "l"(bd), "r"(idesc), "r"(1));
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)));
        const long long c1 = clock64();
        uint32_t done = 0;
        int spins = 0;
        while (!done && ++spins < (1 << 22))
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                         : "=r"(done)
                         : "r"(smem_u32(bar)), "r"(parity));
        parity ^= 1;
        const long long c2 = clock64();
        t_issue += c1 - c0;
        t_total += c2 - c0;
    }
    out[0] = t_issue / REPS;
    out[1] = t_total / REPS;
}

__global__ void __launch_bounds__(128, 1) rate_kernel(long long *out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (65536 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base_s;
    // zero the A columns so that the operands are ordinary numbers
    for (int c = 0; c < 64; c += 8)
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(
                         tmem + 256 + c + (uint32_t(warp * 32) << 16)),
                     "r"(0));
    asm volatile("tcgen05.wait::st.sync.aligned;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (tid == 0) {
        uint32_t parity = 0;
        run<16, true>(tmem, smem, &bar, parity, out + 0);
        run<32, true>(tmem, smem, &bar, parity, out + 2);
        run<64, true>(tmem, smem, &bar, parity, out + 4);
        run<128, true>(tmem, smem, &bar, parity, out + 6);
        run<256, true>(tmem, smem, &bar, parity, out + 8);
        run<16, false>(tmem, smem, &bar, parity, out + 10);
        run<32, false>(tmem, smem, &bar, parity, out + 12);
        run<64, false>(tmem, smem, &bar, parity, out + 14);
        run<128, false>(tmem, smem, &bar, parity, out + 16);
        run<256, false>(tmem, smem, &bar, parity, out + 18);
        run<16, true, 2>(tmem, smem, &bar, parity, out + 20);
        run<64, true, 2>(tmem, smem, &bar, parity, out + 22);
        run<64, true, 4>(tmem, smem, &bar, parity, out + 24);
        run<128, true, 2>(tmem, smem, &bar, parity, out + 26);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

int main() {
    long long *d, h[28];
    cudaMalloc(&d, sizeof(h));
    cudaMemset(d, 0, sizeof(h));
    const int smem = 65536 + 32768;
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    rate_kernel<<<1, 128, smem>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("status: %s\n", cudaGetErrorString(e));
    const int Ns[5] = {16, 32, 64, 128, 256};
    for (int m = 0; m < 2; m++)
        for (int i = 0; i < 5; i++)
            printf("%s M=128 N=%3d K=16 f16: issue %6.1f cycles/MMA, issue->complete %6.1f cycles/MMA (ideal %5.1f)\n",
                   m == 0 ? "A in TMEM" : "A in smem", Ns[i], h[(m * 5 + i) * 2] / 42.0, h[(m * 5 + i) * 2 + 1] / 42.0,
                   128.0 * Ns[i] * 16 * 2 / 8192.0);
    const char *extra[4] = {"N= 16, 2 independent accumulators", "N= 64, 2 independent accumulators",
                            "N= 64, 4 independent accumulators", "N=128, 2 independent accumulators"};
    for (int i = 0; i < 4; i++)
        printf("A in TMEM %s: issue %6.1f, issue->complete %6.1f cycles/MMA\n", extra[i], h[20 + 2 * i] / 42.0, h[21 + 2 * i] / 42.0);
    return e != cudaSuccess;
}
