// Micro-benchmark: how should the warp-uniform weight stream of MLP layer 2 reach the packed FMAs?
// Loop shape of the ViterbiNet priors net: per stage 100 hidden units (sigmoid each) x 50 outputs
// (25 fp32x2 pairs) x M frames per lane.  Variants:
//   smem    weights in shared memory, broadcast LDS.128 per 2 pairs
//   const   weights in the constant bank: LDCU.64 -> uniform register operand of FFMA2
//   const2p constant bank, TWO passes over k with half of the outputs each (13 pairs): halves the
//           accumulator registers per frame, so M can double; sigmoids are recomputed per pass
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_weights microbench_weights.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__constant__ float cW[100 * 52];
__constant__ float cW1B1[204];
__device__ __forceinline__ void ffma2_acc(u64 &c, u64 a, u64 b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c) : "l"(a), "l"(b)); }
__device__ __forceinline__ u64 pack2(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void lds128(unsigned addr, u64 &a, u64 &b) { asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(addr)); }
__device__ __forceinline__ float ex2a(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float rcpa(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

enum { SMEM = 0, CONST = 1, CONST2P = 2 };

template <int M, int MODE, int NT, int UNR>
__global__ void __launch_bounds__(NT, 1) k(float *out, const float *w, int reps) {
    __shared__ __align__(16) float sw[100 * 52];
    for (int i = threadIdx.x; i < 100 * 52; i += blockDim.x) sw[i] = w[i];
    __syncthreads();
    const unsigned sa = (unsigned)__cvta_generic_to_shared(sw);
    constexpr int NP = (MODE == CONST2P) ? 13 : 25;
    constexpr int PASSES = (MODE == CONST2P) ? 2 : 1;
    float y[M];
#pragma unroll
    for (int m = 0; m < M; m++) y[m] = out[threadIdx.x + 32 * m];
    float s = 0;
    for (int r = 0; r < reps; r++) {
#pragma unroll 1
        for (int pass = 0; pass < PASSES; pass++) {
            u64 acc[M][NP];
#pragma unroll
            for (int i = 0; i < NP; i++)
#pragma unroll
                for (int m = 0; m < M; m++) acc[m][i] = 0;
            float hc[M], hn[M];
#pragma unroll
            for (int m = 0; m < M; m++) hc[m] = rcpa(1.f + ex2a(fmaf(y[m], cW1B1[0], cW1B1[1])));
#pragma unroll UNR
            for (int kk = 0; kk < 100; kk++) {
#pragma unroll
                for (int m = 0; m < M; m++) hn[m] = rcpa(1.f + ex2a(fmaf(y[m], cW1B1[2 * kk + 2], cW1B1[2 * kk + 3])));
                u64 hh[M];
#pragma unroll
                for (int m = 0; m < M; m++) hh[m] = pack2(hc[m], hc[m]);
                if (MODE == SMEM) {
#pragma unroll
                    for (int q = 0; q < 12; q++) {
                        u64 wx, wy;
                        lds128(sa + 4 * (kk * 52) + 16 * q, wx, wy);
#pragma unroll
                        for (int m = 0; m < M; m++) { ffma2_acc(acc[m][2 * q], hh[m], wx); ffma2_acc(acc[m][2 * q + 1], hh[m], wy); }
                    }
                    u64 wx, wy;
                    lds128(sa + 4 * (kk * 52) + 16 * 12, wx, wy);
#pragma unroll
                    for (int m = 0; m < M; m++) ffma2_acc(acc[m][24], hh[m], wx);
                } else {
#pragma unroll
                    for (int i = 0; i < NP; i++) {
                        const int col = (MODE == CONST2P ? 26 * pass : 0) + 2 * i;
                        const u64 wv = pack2(cW[kk * 52 + col], cW[kk * 52 + col + 1]);
#pragma unroll
                        for (int m = 0; m < M; m++) ffma2_acc(acc[m][i], hh[m], wv);
                    }
                }
#pragma unroll
                for (int m = 0; m < M; m++) hc[m] = hn[m];
            }
#pragma unroll
            for (int m = 0; m < M; m++)
#pragma unroll
                for (int i = 0; i < NP; i++) { float a, b; asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(acc[m][i])); s += fmaxf(a, 0.f) + fmaxf(b, 0.f); }
        }
#pragma unroll
        for (int m = 0; m < M; m++) y[m] += 0.01f;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int M, int MODE, int NT, int UNR>
void run(const char *name, float *out, float *w, int sms) {
    const int reps = 100;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k<M, MODE, NT, UNR><<<sms, NT>>>(out, w, reps);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    k<M, MODE, NT, UNR><<<sms, NT>>>(out, w, reps);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double fma = double(sms) * NT * reps * 100.0 * 50 * M;   // useful MACs (pad pair of const2p not counted)
    printf("%-10s M=%d threads/SM=%d unroll=%d  %.3f ms  %.2f TFLOP/s useful (err=%s)\n", name, M, NT, UNR, ms,
           2 * fma / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float *out, *w;
    cudaMalloc(&out, sms * 512 * sizeof(float) + 4096); cudaMemset(out, 0, sms * 512 * sizeof(float) + 4096);
    cudaMalloc(&w, 5200 * sizeof(float));
    float hw[5200]; for (int i = 0; i < 5200; i++) hw[i] = 1e-3f * (i % 7);
    cudaMemcpy(w, hw, sizeof(hw), cudaMemcpyHostToDevice);
    cudaMemcpyToSymbol(cW, hw, sizeof(hw));
    float h1[204]; for (int i = 0; i < 204; i++) h1[i] = 0.01f * (i % 11) - 0.05f;
    cudaMemcpyToSymbol(cW1B1, h1, sizeof(h1));
    run<2, SMEM, 256, 4>("smem", out, w, sms);
    run<2, CONST, 256, 4>("const", out, w, sms);
    run<3, SMEM, 256, 4>("smem", out, w, sms);
    run<3, CONST, 256, 4>("const", out, w, sms);
    run<3, CONST, 256, 2>("const", out, w, sms);
    run<3, CONST, 256, 10>("const", out, w, sms);
    run<4, CONST, 256, 4>("const", out, w, sms);
    run<4, CONST2P, 256, 4>("const2p", out, w, sms);
    run<4, CONST2P, 384, 4>("const2p", out, w, sms);
    run<6, CONST2P, 256, 4>("const2p", out, w, sms);
    run<8, CONST2P, 256, 4>("const2p", out, w, sms);
    run<2, CONST, 512, 4>("const", out, w, sms);
    return 0;
}
