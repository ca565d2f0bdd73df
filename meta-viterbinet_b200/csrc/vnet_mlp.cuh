// Shared device code of the ViterbiNet kernels: packed-FMA helpers, weight sources (shared memory /
// constant bank), the staged weight layout and the MLP layers.  Included by vnet_kernels.cu, which also
// includes the tcgen05 variant (vnet_tc_kernel.cuh), so all of it is one translation unit.
#pragma once
#include "mvn_common.cuh"

// Constant-bank weight slots.  C linkage so that inline PTX can name the symbol.
constexpr int kConstSlotFloats = 7168;
// (defined here: this header is included by exactly one translation unit, vnet_kernels.cu)
extern "C" {
__constant__ float mvn_cParams[2 * kConstSlotFloats];
__device__ float mvn_gStage[2 * kConstSlotFloats];
}

namespace mvn {

typedef unsigned long long u64;

__device__ __forceinline__ u64 pack2(float a, float b) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void unpack2(u64 v, float &a, float &b) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
// d = a*b + c on both halves, round-to-nearest-even (same rounding as two scalar fmaf).
// Accumulating form (c += a*b) with a read-write operand so that ptxas keeps the accumulator in place.
__device__ __forceinline__ void ffma2_acc(u64 &c, u64 a, u64 b) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c) : "l"(a), "l"(b));
}
__device__ __forceinline__ float ex2_approx(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// Shared-memory weight loads as volatile asm: the weights are loop-invariant across stages, and
// without this NVVM hoists them out of the stage loop into (spilled) registers.
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void lds128(uint32_t addr, u64 &a, u64 &b) {
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(addr));
}
__device__ __forceinline__ u64 lds64(uint32_t addr) {
    u64 a;
    asm volatile("ld.shared.b64 %0, [%1];" : "=l"(a) : "r"(addr));
    return a;
}
__device__ __forceinline__ float2 lds64f(uint32_t addr) {
    float2 a;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(a.x), "=f"(a.y) : "r"(addr));
    return a;
}

// ------------------------------------------------------------------ weight sources
// kSmem : staged layout in shared memory, warp-uniform (broadcast) LDS.64/.128.
// kConst: staged layout in the constant bank; warp-uniform addresses make ptxas fetch the weights
//         with LDCU into UNIFORM registers that FFMA2 takes directly as an operand, so the weight
//         stream costs no vector registers, no LSU wavefronts and no RF write ports.
//         (memory_length <= 5: two 28 KB slots of the 64 KB bank, used round-robin per call.)
enum WeightSrc { kSmem = 0, kConst = 1 };

// Volatile asm like the shared-memory loads (no hoisting out of the stage loop); the address is
// warp-uniform (const-space address of cParams + slot + immediate), which is what lets ptxas pick LDCU.
__device__ __forceinline__ u64 ldc64(uint32_t addr) {
    u64 a;
    asm volatile("ld.const.b64 %0, [%1];" : "=l"(a) : "r"(addr));
    return a;
}
__device__ __forceinline__ float2 ldc64f(uint32_t addr) {
    float2 a;
    asm volatile("ld.const.v2.f32 {%0, %1}, [%2];" : "=f"(a.x), "=f"(a.y) : "r"(addr));
    return a;
}
__device__ __forceinline__ uint32_t const_params_addr() {
    uint32_t a;
    asm("mov.u32 %0, mvn_cParams;" : "=r"(a));
    return a;
}

template <int WS>
struct Wt {
    uint32_t base;  // byte address of the staged block in shared (kSmem) or constant (kConst) space
    __device__ __forceinline__ u64 pair(int off) const {
        if constexpr (WS == kSmem) {
            return lds64(base + 4 * off);
        } else {
            return ldc64(base + 4 * off);
        }
    }
    __device__ __forceinline__ void quad(int off, u64 &a, u64 &b) const {
        if constexpr (WS == kSmem) {
            lds128(base + 4 * off, a, b);
        } else {
            a = pair(off);
            b = pair(off + 2);
        }
    }
    __device__ __forceinline__ float2 f2(int off) const {
        if constexpr (WS == kSmem) {
            return lds64f(base + 4 * off);
        } else {
            return ldc64f(base + 4 * off);
        }
    }
};

// ------------------------------------------------------------------ staged weights
#ifndef MVN_K_UNROLL
#define MVN_K_UNROLL 5
#endif
constexpr int kKUnroll = MVN_K_UNROLL;
constexpr int kW2Ld = 52;  // 50 outputs padded to 13 x 128-bit
template <int L>
struct VnetSmem {
    static constexpr int S = 1 << L;
    static constexpr int S4 = (S + 3) / 4 * 4;
    static constexpr int oW1B1 = 0;                     // [101][2]: (-log2e*w1, -log2e*b1), one pad pair
    static constexpr int oB2 = 204;                     // [52]
    static constexpr int oW2T = oB2 + kW2Ld;            // [100][52]   W2T[k][o] = w2[o][k]
    static constexpr int oB3 = oW2T + kH1 * kW2Ld;      // [S4]
    static constexpr int oW3T = oB3 + S4;               // [50][S]     W3T[j][s] = w3[s][j]
    static constexpr int kFloats = (oW3T + kH2 * S + 3) / 4 * 4;
};

struct VnetWeights {
    const float *w1, *b1, *w2, *b2, *w3, *b3;
};

template <int L>
__device__ void stage_weights(float *sm, const VnetWeights &w, int tid, int nt) {
    using W = VnetSmem<L>;
    constexpr int S = W::S;
    const float kNegLog2e = -1.4426950408889634f;
    for (int i = tid; i < 102; i += nt) {
        sm[W::oW1B1 + 2 * i] = i < kH1 ? w.w1[i] * kNegLog2e : 0.f;
        sm[W::oW1B1 + 2 * i + 1] = i < kH1 ? w.b1[i] * kNegLog2e : 0.f;
    }
    for (int i = tid; i < kW2Ld; i += nt) sm[W::oB2 + i] = i < kH2 ? w.b2[i] : 0.f;
    for (int i = tid; i < kH2 * kH1; i += nt) {
        const int o = i / kH1, k = i % kH1;
        sm[W::oW2T + k * kW2Ld + o] = w.w2[i];
    }
    for (int i = tid; i < kH1 * 2; i += nt) sm[W::oW2T + (i >> 1) * kW2Ld + kH2 + (i & 1)] = 0.f;
    for (int i = tid; i < W::S4; i += nt) sm[W::oB3 + i] = i < S ? w.b3[i] : 0.f;
    for (int i = tid; i < S * kH2; i += nt) {
        const int s = i / kH2, j = i % kH2;
        sm[W::oW3T + j * S + s] = w.w3[i];
    }
}

// ------------------------------------------------------------------ the MLP, M samples per lane
template <int M, int WS>
__device__ __forceinline__ void sigmoid_unit(const Wt<WS> &wt, int k, const float (&y)[M], float (&h)[M]) {
    const float2 wb = wt.f2(2 * k);  // oW1B1 == 0
#pragma unroll
    for (int m = 0; m < M; m++) h[m] = rcp_approx(1.f + ex2_approx(fmaf(y[m], wb.x, wb.y)));
}

// layers 1+2 (+ReLU): y[M] -> h2[M][50]
template <int L, int M, int WS, int KU>
__device__ __forceinline__ void mlp_hidden(const Wt<WS> &wt, const float (&y)[M], float (&h2)[M][kH2]) {
    using W = VnetSmem<L>;
    static_assert(W::oW1B1 == 0, "sigmoid_unit assumes the (w1,b1) pairs lead the staged block");
    u64 acc[M][25];
#pragma unroll
    for (int i = 0; i < 25; i++) {
        const u64 b = wt.pair(W::oB2 + 2 * i);
#pragma unroll
        for (int m = 0; m < M; m++) acc[m][i] = b;
    }
    float hc[M], hn[M];
    sigmoid_unit<M, WS>(wt, 0, y, hc);
#pragma unroll KU
    for (int k = 0; k < kH1; k++) {
        sigmoid_unit<M, WS>(wt, k + 1, y, hn);  // entry 100 is a zero pad
        const int wr = W::oW2T + k * kW2Ld;
        u64 hh[M];
#pragma unroll
        for (int m = 0; m < M; m++) hh[m] = pack2(hc[m], hc[m]);
#pragma unroll
        for (int q = 0; q < 12; q++) {
            u64 wx, wy;
            wt.quad(wr + 4 * q, wx, wy);
#pragma unroll
            for (int m = 0; m < M; m++) {
                ffma2_acc(acc[m][2 * q], hh[m], wx);
                ffma2_acc(acc[m][2 * q + 1], hh[m], wy);
            }
        }
        const u64 wl = wt.pair(wr + 48);
#pragma unroll
        for (int m = 0; m < M; m++) ffma2_acc(acc[m][24], hh[m], wl);
#pragma unroll
        for (int m = 0; m < M; m++) hc[m] = hn[m];
    }
#pragma unroll
    for (int m = 0; m < M; m++)
#pragma unroll
        for (int i = 0; i < 25; i++) {
            float a, b;
            unpack2(acc[m][i], a, b);
            h2[m][2 * i] = fmaxf(a, 0.f);
            h2[m][2 * i + 1] = fmaxf(b, 0.f);
        }
}

// layer 3 for output states [c*C, c*C + C): p[m][i] = b3 + sum_j h2[m][j] W3T[j][c*C+i]
template <int L, int M, int WS>
__device__ __forceinline__ void mlp_out_chunk(const Wt<WS> &wt, int c, const float (&h2)[M][kH2],
                                              float (&p)[M][TrellisDims<L>::C]) {
    using W = VnetSmem<L>;
    constexpr int S = W::S, C = TrellisDims<L>::C, P = C / 2;
    u64 acc[M][P];
#pragma unroll
    for (int i = 0; i < P; i++) {
        const u64 b = wt.pair(W::oB3 + c * C + 2 * i);
#pragma unroll
        for (int m = 0; m < M; m++) acc[m][i] = b;
    }
    const int w3 = W::oW3T + c * C;
#pragma unroll
    for (int j = 0; j < kH2; j++) {
        u64 w[P];
        if constexpr (C >= 4) {
#pragma unroll
            for (int i = 0; i < C / 4; i++) wt.quad(w3 + j * S + 4 * i, w[2 * i], w[2 * i + 1]);
        } else {
            w[0] = wt.pair(w3 + j * S);
        }
#pragma unroll
        for (int m = 0; m < M; m++) {
            const u64 hh = pack2(h2[m][j], h2[m][j]);
#pragma unroll
            for (int i = 0; i < P; i++) ffma2_acc(acc[m][i], hh, w[i]);
        }
    }
#pragma unroll
    for (int m = 0; m < M; m++)
#pragma unroll
        for (int i = 0; i < P; i++) unpack2(acc[m][i], p[m][2 * i], p[m][2 * i + 1]);
}


}  // namespace mvn
