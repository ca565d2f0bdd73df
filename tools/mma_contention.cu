// Microbenchmark: how much do the other warps of the fused kernel slow the tcgen05.mma stream down?
// One thread issues the per-stage MMA pattern of vnet_decode_tc_kernel (7 x N=128, 7 x N=64, 8 x N=32, all M=128 K=16,
// A from TMEM) while 12 other warps run one kind of background work until it is done:
//   0 nothing   1 tcgen05.st / tcgen05.ld on other TMEM columns   2 FFMA2 + MUFU arithmetic   3 shared-memory LDS.128
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_contention tools/mma_contention.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo) {
    return uint64_t((saddr >> 4) & 0x3fff) | (uint64_t((lbo >> 4) & 0x3fff) << 16) | (uint64_t(128 >> 4) << 32) | (uint64_t(1) << 46);
}
__device__ __forceinline__ void mma(uint32_t d, uint32_t a, uint64_t bd, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a), "l"(bd), "r"(idesc), "r"(1));
}
__host__ __device__ constexpr uint32_t idesc_n(int N) { return (1u << 4) | (uint32_t(N >> 3) << 17) | (uint32_t(128 >> 4) << 24); }

__global__ void __launch_bounds__(416, 1) contention_kernel(int mode, long long *out, float *sink) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(8) uint64_t bar;
    __shared__ volatile int stop;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0;
    if (tid == 0) {
        stop = 0;
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
    const uint32_t lane_base = uint32_t((warp & 3) * 32) << 16;
    if (warp < 4) {
        for (int c = 0; c < 64; c += 8)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(tmem + 128 + c + lane_base), "r"(0));
        asm volatile("tcgen05.wait::st.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (warp == 12) {
        if (lane == 0) {
            const uint32_t sB = smem_u32(smem);
            uint32_t parity = 0;
            long long t_total = 0;
            const int REPS = 200;
            for (int rep = 0; rep < REPS; rep++) {
                const long long c0 = clock64();
#pragma unroll
                for (int j = 0; j < 7; j++) mma(tmem, tmem + 128 + j * 8, desc(sB + 2 * j * 2048, 2048), idesc_n(128));
#pragma unroll
                for (int j = 0; j < 7; j++) mma(tmem + 64, tmem + 184 + j * 8 % 56, desc(sB + 2 * j * 2048, 2048), idesc_n(64));
#pragma unroll
                for (int j = 0; j < 8; j++) mma(tmem + 256, tmem + 256 + 128 + (j % 4) * 8, desc(sB + 32768 + 2 * (j % 4) * 512, 512), idesc_n(32));
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)));
                uint32_t done = 0;
                int spins = 0;
                while (!done && ++spins < (1 << 22))
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                                 : "=r"(done)
                                 : "r"(smem_u32(&bar)), "r"(parity));
                parity ^= 1;
                t_total += clock64() - c0;
            }
            out[mode] = t_total / REPS;
            stop = 1;
        }
        __syncwarp();
    } else {
        float acc = float(tid);
        unsigned long long pk = 0x3f8000003f800000ull;
        int it = 0;
        while (!stop && ++it < (1 << 20)) {
            if (mode == 1) {
                if (warp < 8) {
                    const uint32_t a = tmem + 400 + (warp >> 2) * 16 + lane_base;
                    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(a), "r"(it));
                    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(a + 8), "r"(it));
                    asm volatile("tcgen05.wait::st.sync.aligned;");
                } else {
                    uint32_t u[16];
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                                 : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
                                   "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
                                 : "r"(tmem + 440 + lane_base));
                    asm volatile("tcgen05.wait::ld.sync.aligned;");
                    acc += __uint_as_float(u[3]);
                }
            } else if (mode == 2) {
#pragma unroll
                for (int q = 0; q < 16; q++) {
                    asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(pk) : "l"(0x3f7fff003f7fff00ull));
                    float e;
                    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(acc));
                    acc = e * 0.5f;
                }
            } else if (mode == 3) {
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    unsigned long long a, b;
                    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(smem_u32(smem) + 40960 + 16 * ((q + it) & 63)));
                    pk += a + b;
                }
            } else {
                __nanosleep(200);
            }
        }
        if (acc == 123.456f && pk == 77) sink[tid] = acc;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

int main() {
    long long *d, h[4] = {0, 0, 0, 0};
    float *sink;
    cudaMalloc(&d, sizeof(h));
    cudaMalloc(&sink, 4096);
    cudaMemset(d, 0, sizeof(h));
    cudaFuncSetAttribute(contention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    const char *names[4] = {"other warps idle", "tcgen05.st / tcgen05.ld traffic", "FFMA2 + MUFU arithmetic", "shared-memory LDS.128"};
    for (int mode = 0; mode < 4; mode++) {
        contention_kernel<<<1, 416, 65536>>>(mode, d, sink);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            printf("mode %d: %s\n", mode, cudaGetErrorString(e));
            return 1;
        }
    }
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    for (int mode = 0; mode < 4; mode++)
        printf("%-34s: %6lld cycles per stage pattern of 22 MMAs (%5.1f per MMA; floor 7x66 + 15x48 = 1182)\n", names[mode], h[mode], h[mode] / 22.0);
    return 0;
}
