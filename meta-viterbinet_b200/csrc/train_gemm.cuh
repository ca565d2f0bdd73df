// Batched (meta-)training step of the priors net as a sequence of small register-tiled GEMMs out of shared memory
// (a9-a11: trainers/trainer.py:425-453 meta_train_loop, :492-505 run_train_loop, metavnet_trainer.py:41-50 loss).
//
// One CTA (MVN_TRAIN_THREADS = 512 threads) owns one realisation.  Everything of a step lives in shared memory:
//   Wset   the weights in a padded torch layout  [w1 100][b1 100][W2 52x100][b2 52][W3 SPx52][b3 SP]   (rows 50, 51 of
//          W2 / b2 and columns 50, 51 of W3 are zero, SP = max(S, 4))
//   Vset   the tangent direction in the same layout (MAML's Hessian-vector product only)
//   G      the gradient being accumulated, same layout
//   acts   FEATURE-MAJOR activations of one chunk of symbols:  y[n], H1T[100][ldn], H2T[52][ldn], ZT[SP][ldn]
//          (+ the tangents RH1T, RH2T, RZT), ldn = 4 (mod 8) so that strided row sets are bank-conflict free.
// A pass over a chunk is
//   E1  H1 = sigmoid(w1 y + b1)                                   (elementwise, all threads)
//   G1  H2T = relu(W2 H1T + b2)          G2  ZT = W3 H2T + b3      (GEMMs, K = 100 / 52)
//   E2  softmax-CE per symbol: loss, dZ over ZT, db3
//   G3  dW3 += dZT H2T^T   (K = symbols)  ||  G4  DA2T = [H2>0] W3^T dZT  (+ db2)          one phase, one block list
//       (the second half of G3's K range runs in the next phase: MVN_TRAIN_SPLIT_DW3)
//   G5  dW2 += DA2T H1T^T  (K = symbols)  ||  G6  dA1 = (W2^T DA2T) h(1-h) -> dW1, db1      one phase, never stored
// and the forward-over-reverse (tangent) pass runs the same schedule with a second accumulator set
//   R{A B} = Av B + A Rb,  which is what MAML's  g_q - meta_lr H_s(theta) g_q  needs.
// The GEMM micro-kernel: a thread owns a 4x4 output tile and walks K in blocks of 4; each operand block is four
// LDS.128 whichever way the operand is laid out (contiguous along the tile dimension -> 4 rows of the k-block;
// contiguous along K -> the tile's 4 rows, which are then an INTERLEAVED row set r, r+4, r+8, r+12 so that the eight /
// four lanes that read different rows hit different banks).  A warp covers a 16 x 32 block of outputs (4 x 8 lanes):
// per k-block 8 LDS.128 (each one shared-memory wavefront, the lanes of a row / column read the same address) feed 64
// FFMA per lane.  Sums run in a different order than torch's: results agree to fp32 rounding (tests: 1e-5).
#pragma once
#include "mvn_common.cuh"

namespace mvn {
namespace tg {

#ifndef MVN_TRAIN_THREADS
#define MVN_TRAIN_THREADS 512   // 16 warps: measured 795k MAML steps/s vs 700k at 256 threads and 770k at 384 (profiles/r02_train_bench.txt)
#endif
#ifndef MVN_TRAIN_SPLIT_DW3
#define MVN_TRAIN_SPLIT_DW3 1
#endif
constexpr int kThreads = MVN_TRAIN_THREADS, kWarps = kThreads / 32;
constexpr int kH2P = 52;   // hidden-2 width padded to a multiple of 4

template <int S>
struct Lay {   // padded parameter layout (floats)
    static constexpr int SP = S < 4 ? 4 : S;
    static constexpr int w1 = 0, b1 = kH1, w2 = 2 * kH1, b2 = w2 + kH2P * kH1, w3 = b2 + kH2P, b3 = w3 + SP * kH2P;
    static constexpr int P = b3 + SP;
    static constexpr int PP = (P + 3) / 4 * 4;
    // torch packing (what theta / Adam state use in HBM): w1[100] b1[100] w2[50][100] b2[50] w3[S][50] b3[S]
    static constexpr int tw2 = 2 * kH1, tb2 = tw2 + kH2 * kH1, tw3 = tb2 + kH2, tb3 = tw3 + S * kH2, TP = tb3 + S;
    __device__ static int to_padded(int i) {
        if (i < tb2) return i;                                   // w1, b1, w2 (same offsets: rows 0..49 of the padded W2)
        if (i < tw3) return b2 + (i - tb2);
        if (i < tb3) {
            const int s = (i - tw3) / kH2, o = (i - tw3) % kH2;
            return w3 + s * kH2P + o;
        }
        return b3 + (i - tb3);
    }
};

enum { AK = 0, AM = 1 };   // A operand: K-contiguous rows (interleaved row set)  |  M-contiguous (k-major storage)
enum { BK = 0, BN = 1 };   // B operand: K-contiguous columns (interleaved set)     |  N-contiguous (k-major storage)
enum { PRIMAL = 0, BOTH = 1, TANGENT = 2 };

__device__ __forceinline__ float4 lds4(const float *p) { return *reinterpret_cast<const float4 *>(p); }

struct Tile {   // what the epilogue gets: the 4 rows / 4 columns of this thread's tile (-1 = outside the matrix)
    int r[4], c[4];
};

// C[M x N] = sum_{k < 4 K4} A(m,k) B(k,n)  (+ the tangent  Av B + A Rb  into a second accumulator set).
// AMODE AK: A[m * lda + k], AM: A[k * lda + m];  BMODE BK: B[n * ldb + k], BN: B[k * ldb + n].
// M, N are the valid extents; rows / columns outside are clamped for the loads (their products are discarded).
// epi(tile, acc, racc) is called once per thread tile; with SKIP tiles that lie entirely outside the matrix are not
// computed at all (without it every lane reaches the epilogue, which may then use warp shuffles).
struct GemmArgs {
    const float *A, *Av;
    int lda;
    const float *B, *Rb;
    int ldb, M, N, K4;
};
__device__ __forceinline__ int num_blocks(const GemmArgs &g) { return ((g.M + 15) / 16) * ((g.N + 31) / 32); }

template <int AMODE, int BMODE, int MODE, bool SKIP, class Epi>
__device__ __forceinline__ void gemm_block(const GemmArgs &g, int blk, Epi epi) {
    const int lane = threadIdx.x & 31, lm = lane >> 3, ln = lane & 7;
    const int nbn = (g.N + 31) / 32;
    const int bm = blk / nbn, bn = blk % nbn;
    const int M = g.M, N = g.N, lda = g.lda, ldb = g.ldb;
    const float *__restrict__ A = g.A, *__restrict__ Av = g.Av, *__restrict__ B = g.B, *__restrict__ Rb = g.Rb;
    Tile t;
    int ra[4], cb[4];   // clamped
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int r = (AMODE == AK) ? bm * 16 + lm + 4 * i : bm * 16 + 4 * lm + i;
        const int c = (BMODE == BK) ? bn * 32 + ln + 8 * i : bn * 32 + 4 * ln + i;
        t.r[i] = r < M ? r : -1;
        t.c[i] = c < N ? c : -1;
        ra[i] = min(r, M - 1);
        cb[i] = min(c, N - 1);
    }
    // AM / BN read 4 consecutive rows / columns with one LDS.128: the padded extents are multiples of 4, so a group
    // that starts inside the matrix lies inside the allocation; a group that starts outside is clamped as a whole
    const int m0 = min(bm * 16 + 4 * lm, ((M + 3) / 4 - 1) * 4), n0 = min(bn * 32 + 4 * ln, ((N + 3) / 4 - 1) * 4);
    if (SKIP && (t.r[0] < 0 || t.c[0] < 0)) return;   // tile entirely outside (the first row / column of the set is the smallest)
    float acc[4][4], racc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = racc[i][j] = 0.f;
#pragma unroll 2
    for (int kb = 0; kb < g.K4; kb++) {
        float a[4][4], b[4][4], av[4][4], rb[4][4];   // a[mi][kk], b[kk][nj]
#pragma unroll
        for (int q = 0; q < 4; q++) {
            if (AMODE == AK) {
                const float4 v = lds4(A + ra[q] * lda + 4 * kb);
                a[q][0] = v.x, a[q][1] = v.y, a[q][2] = v.z, a[q][3] = v.w;
                if (MODE != PRIMAL) {
                    const float4 u = lds4(Av + ra[q] * lda + 4 * kb);
                    av[q][0] = u.x, av[q][1] = u.y, av[q][2] = u.z, av[q][3] = u.w;
                }
            } else {
                const float4 v = lds4(A + (4 * kb + q) * lda + m0);
                a[0][q] = v.x, a[1][q] = v.y, a[2][q] = v.z, a[3][q] = v.w;
                if (MODE != PRIMAL) {
                    const float4 u = lds4(Av + (4 * kb + q) * lda + m0);
                    av[0][q] = u.x, av[1][q] = u.y, av[2][q] = u.z, av[3][q] = u.w;
                }
            }
            if (BMODE == BK) {
                const float4 v = lds4(B + cb[q] * ldb + 4 * kb);
                b[0][q] = v.x, b[1][q] = v.y, b[2][q] = v.z, b[3][q] = v.w;
                if (MODE != PRIMAL) {
                    const float4 u = lds4(Rb + cb[q] * ldb + 4 * kb);
                    rb[0][q] = u.x, rb[1][q] = u.y, rb[2][q] = u.z, rb[3][q] = u.w;
                }
            } else {
                const float4 v = lds4(B + (4 * kb + q) * ldb + n0);
                b[q][0] = v.x, b[q][1] = v.y, b[q][2] = v.z, b[q][3] = v.w;
                if (MODE != PRIMAL) {
                    const float4 u = lds4(Rb + (4 * kb + q) * ldb + n0);
                    rb[q][0] = u.x, rb[q][1] = u.y, rb[q][2] = u.z, rb[q][3] = u.w;
                }
            }
        }
#pragma unroll
        for (int kk = 0; kk < 4; kk++)
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    if (MODE != TANGENT) acc[i][j] = fmaf(a[i][kk], b[kk][j], acc[i][j]);
                    if (MODE != PRIMAL) racc[i][j] = fmaf(av[i][kk], b[kk][j], fmaf(a[i][kk], rb[kk][j], racc[i][j]));
                }
    }
    if (AMODE == AM) {   // contiguous rows: report the unclamped group
#pragma unroll
        for (int i = 0; i < 4; i++) t.r[i] = (bm * 16 + 4 * lm + i < M) ? bm * 16 + 4 * lm + i : -1;
    }
    if (BMODE == BN) {
#pragma unroll
        for (int i = 0; i < 4; i++) t.c[i] = (bn * 32 + 4 * ln + i < N) ? bn * 32 + 4 * ln + i : -1;
    }
    epi(t, acc, racc);
}

// warps take the 16 x 32 blocks of one GEMM round-robin
template <int AMODE, int BMODE, int MODE, bool SKIP = true, class Epi>
__device__ __forceinline__ void cta_gemm(const GemmArgs &g, Epi epi) {
    const int nb = num_blocks(g);
    for (int blk = threadIdx.x >> 5; blk < nb; blk += kWarps) gemm_block<AMODE, BMODE, MODE, SKIP>(g, blk, epi);
}

// 1 / (1 + e^-a) with the fast reciprocal (2 ulp): well inside the 1e-5 parity bar of the training path
__device__ __forceinline__ float sigmoid_acc(float a) { return __fdividef(1.f, 1.f + __expf(-a)); }

// shared-memory views of one step
template <int S>
struct Smem {
    float *W, *V, *G;           // Lay<S>::PP floats each (V only when tangents are used)
    float *y;                   // [ldn]
    float *H1T, *H2T, *ZT;      // [100][ldn], [52][ldn], [SP][ldn]
    float *DA2T;                // [52][ldn]  dL/d(pre-activation 2), kept apart from H2T so that dW3 and dA2 share a phase
    float *RH1T, *RH2T, *RZT, *RDA2T;   // tangents
    float *red;                 // [32]
    int ldn;
};

// activation floats of one layout
template <int S>
__host__ __device__ constexpr size_t act_floats(int ldn, bool tangent) {
    return size_t(kH1 + 2 * kH2P + Lay<S>::SP) * ldn * (tangent ? 2 : 1) + ldn;
}

// G[base + r[i] * ld + c[j]] += scale * v[i][j] for the valid rows / columns of a tile: all 16 loads, then the 16 stores
// (a read-modify-write per element in one loop serialises: each store may alias the next load)
__device__ __forceinline__ void accumulate_tile(float *__restrict__ Gbase, int ld, const Tile &t, const float (&v)[4][4], float scale) {
    float g[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) g[i][j] = (t.r[i] >= 0 && t.c[j] >= 0) ? Gbase[t.r[i] * ld + t.c[j]] : 0.f;
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (t.r[i] >= 0 && t.c[j] >= 0) Gbase[t.r[i] * ld + t.c[j]] = fmaf(scale, v[i][j], g[i][j]);
}

// sum of v over the 8 lanes that share lm (the lanes of one tile row), result in every lane
__device__ __forceinline__ float sum_ln(float v) {
    v += __shfl_xor_sync(kFull, v, 1);
    v += __shfl_xor_sync(kFull, v, 2);
    v += __shfl_xor_sync(kFull, v, 4);
    return v;
}

// One pass over the symbols of one word set: adds  scale * dL/dtheta  (TAN: scale * H(theta) v) into sm.G, returns this
// thread's share of the summed (not yet averaged) loss.  inv_n = 1 / (total symbols of the loss).
// Phases per chunk of symbols (one __syncthreads between them):
//   y -> E1 (sigmoid) -> G1 (layer 2 + ReLU) -> G2 (layer 3) -> E2 (softmax-CE, dZ, db3)
//     -> G3 (dW3) || G4 (dA2, db2)  -> G5 (dW2) || G6 (dA1 -> dW1, db1 folded into the epilogue)
// Bias / first-layer sums that several warps contribute to go through shared-memory atomics (a handful per row; the
// order of those few additions is the only run-to-run nondeterminism, at the 1e-7 level).
template <int S, bool TAN, bool MERGE>
__device__ __forceinline__ float pass(const Smem<S> &sm, const float *__restrict__ y, const int *__restrict__ lab, int n,
                                   float inv_n, float scale, int chunk_cap) {
    using LY = Lay<S>;
    constexpr int SP = LY::SP;
    constexpr int GM = TAN ? BOTH : PRIMAL, WM = TAN ? TANGENT : PRIMAL;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, ldn = sm.ldn;
    const float *W = sm.W, *V = sm.V;
    float *G = sm.G;
    float loss = 0.f;
    for (int base = 0; base < n; base += chunk_cap) {
        const int nv = min(chunk_cap, n - base);
        const int n4 = (nv + 3) / 4 * 4, ng = n4 / 4;   // symbols of this chunk padded to the k-block
        __syncthreads();                                // previous chunk / previous pass done with the activations
        for (int i = tid; i < n4; i += kThreads) sm.y[i] = i < nv ? y[base + i] : 0.f;
        __syncthreads();
        // ---- E1: H1T[k][n] = sigmoid(w1[k] y[n] + b1[k]);  RH1 = h (1 - h) (v1[k] y[n] + c1[k])
        {
            const unsigned inv_ng = (1u << 20) / unsigned(ng) + 1u;   // e / ng == (e * inv_ng) >> 20 for e * ng < 2^20
            for (int e = tid; e < kH1 * ng; e += kThreads) {
                const int k = int((unsigned(e) * inv_ng) >> 20), g = e - k * ng;
                const float w = W[LY::w1 + k], b = W[LY::b1 + k];
                const float4 yv = lds4(sm.y + 4 * g);
                float4 h;
                h.x = sigmoid_acc(fmaf(w, yv.x, b)), h.y = sigmoid_acc(fmaf(w, yv.y, b));
                h.z = sigmoid_acc(fmaf(w, yv.z, b)), h.w = sigmoid_acc(fmaf(w, yv.w, b));
                *reinterpret_cast<float4 *>(sm.H1T + k * ldn + 4 * g) = h;
                if (TAN) {
                    const float vw = V[LY::w1 + k], vb = V[LY::b1 + k];
                    float4 r;
                    r.x = h.x * (1.f - h.x) * fmaf(vw, yv.x, vb), r.y = h.y * (1.f - h.y) * fmaf(vw, yv.y, vb);
                    r.z = h.z * (1.f - h.z) * fmaf(vw, yv.z, vb), r.w = h.w * (1.f - h.w) * fmaf(vw, yv.w, vb);
                    *reinterpret_cast<float4 *>(sm.RH1T + k * ldn + 4 * g) = r;
                }
            }
        }
        __syncthreads();
        // ---- G1: H2T[o][n] = relu(sum_k W2[o][k] H1T[k][n] + b2[o]);  RH2 = [a2 > 0] (V2 H1 + W2 RH1 + c2)
        cta_gemm<AK, BN, GM>(GemmArgs{W + LY::w2, V + LY::w2, kH1, sm.H1T, sm.RH1T, ldn, kH2P, n4, kH1 / 4},
            [&](const Tile &t, float (&acc)[4][4], float (&racc)[4][4]) {
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    if (t.r[i] < 0) continue;
                    const float b = W[LY::b2 + t.r[i]];
                    float4 h, r;
                    const float a0 = acc[i][0] + b, a1 = acc[i][1] + b, a2 = acc[i][2] + b, a3 = acc[i][3] + b;
                    // (a < 0 ? 0 : a) instead of fmaxf: a NaN sample must reach the loss like torch's relu lets it, so that
                    // run_train_loop's NaN rule can fire (fmaxf would return 0 for NaN)
                    h.x = a0 < 0.f ? 0.f : a0, h.y = a1 < 0.f ? 0.f : a1, h.z = a2 < 0.f ? 0.f : a2, h.w = a3 < 0.f ? 0.f : a3;
                    *reinterpret_cast<float4 *>(sm.H2T + t.r[i] * ldn + t.c[0]) = h;
                    if (TAN) {
                        const float c = V[LY::b2 + t.r[i]];
                        r.x = a0 > 0.f ? racc[i][0] + c : 0.f, r.y = a1 > 0.f ? racc[i][1] + c : 0.f;
                        r.z = a2 > 0.f ? racc[i][2] + c : 0.f, r.w = a3 > 0.f ? racc[i][3] + c : 0.f;
                        *reinterpret_cast<float4 *>(sm.RH2T + t.r[i] * ldn + t.c[0]) = r;
                    }
                }
            });
        __syncthreads();
        // ---- G2: ZT[s][n] = sum_o W3[s][o] H2T[o][n] + b3[s];  RZ = V3 H2 + W3 RH2 + c3
        cta_gemm<AK, BN, GM>(GemmArgs{W + LY::w3, V + LY::w3, kH2P, sm.H2T, sm.RH2T, ldn, SP, n4, kH2P / 4},
            [&](const Tile &t, float (&acc)[4][4], float (&racc)[4][4]) {
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    if (t.r[i] < 0) continue;
                    const float b = W[LY::b3 + t.r[i]];
                    *reinterpret_cast<float4 *>(sm.ZT + t.r[i] * ldn + t.c[0]) =
                        make_float4(acc[i][0] + b, acc[i][1] + b, acc[i][2] + b, acc[i][3] + b);
                    if (TAN) {
                        const float c = V[LY::b3 + t.r[i]];
                        *reinterpret_cast<float4 *>(sm.RZT + t.r[i] * ldn + t.c[0]) =
                            make_float4(racc[i][0] + c, racc[i][1] + c, racc[i][2] + c, racc[i][3] + c);
                    }
                }
            });
        __syncthreads();
        // ---- E2: softmax cross-entropy per symbol (torch CrossEntropyLoss, mean reduction: 1/N folded into dz);
        //          dZ over ZT, R{dZ} = p (Rz - <p, Rz>) / N over RZT; padded symbols and padded states get 0;
        //          db3[s] += sum_n dZ[s][n] (TAN: of R{dZ}) by warp shuffles + one atomic per warp and state
        for (int i0 = warp * 32; i0 < n4; i0 += kThreads) {   // warp-uniform trip count: whole warps stay for the shuffles
            const int i = i0 + lane;
            float z[S], rz[TAN ? S : 1];
#pragma unroll
            for (int s = 0; s < S; s++) {
                z[s] = 0.f;
                if (TAN) rz[s] = 0.f;
            }
            if (i < nv) {
#pragma unroll
                for (int s = 0; s < S; s++) {
                    z[s] = sm.ZT[s * ldn + i];
                    if (TAN) rz[s] = sm.RZT[s * ldn + i];
                }
                const int label = lab[base + i];
                float m = z[0];
#pragma unroll
                for (int s = 1; s < S; s++) m = fmaxf(m, z[s]);
                float sum = 0.f, zl = 0.f;
#pragma unroll
                for (int s = 0; s < S; s++) {
                    zl = (s == label) ? z[s] : zl;
                    z[s] = expf(z[s] - m);
                    sum += z[s];
                }
                loss += logf(sum) + m - zl;
                const float inv = 1.f / sum;
                float dot = 0.f;
#pragma unroll
                for (int s = 0; s < S; s++) {
                    z[s] *= inv;   // p
                    if (TAN) dot = fmaf(z[s], rz[s], dot);
                }
#pragma unroll
                for (int s = 0; s < S; s++) {
                    if (TAN) rz[s] = z[s] * (rz[s] - dot) * inv_n;
                    z[s] = (z[s] - ((s == label) ? 1.f : 0.f)) * inv_n;
                }
            }
            if (i < n4) {
#pragma unroll
                for (int s = 0; s < SP; s++) {
                    sm.ZT[s * ldn + i] = s < S ? z[s < S ? s : 0] : 0.f;
                    if (TAN) sm.RZT[s * ldn + i] = s < S ? rz[s < S ? s : 0] : 0.f;
                }
            }
#pragma unroll
            for (int s = 0; s < S; s++) {
                float v = TAN ? rz[s] : z[s];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
                if (lane == 0) atomicAdd(G + LY::b3 + s, scale * v);
            }
        }
        __syncthreads();
        // ---- G3: dW3[s][o] += sum_n dZT[s][n] H2T[o][n]   (TAN: RdZ H2^T + dZ RH2^T)       — in the same phase as
        // ---- G4: DA2T[o][n] = [H2 > 0] sum_s W3[s][o] dZT[s][n]  (TAN: R = V3^T dZ + W3^T RdZ -> RDA2T), db2 folded in
        {
            const GemmArgs g3{sm.ZT, sm.RZT, ldn, sm.H2T, sm.RH2T, ldn, S, kH2, ng};
            const GemmArgs g4{W + LY::w3, V + LY::w3, kH2P, sm.ZT, sm.RZT, ldn, kH2P, n4, SP / 4};
            // MERGE: DA2T is a buffer of its own and both GEMMs share one block list; otherwise DA2T aliases H2T (in place:
            // a thread reads the H2 values of its own tile before it overwrites them) and dW3 has to finish first
            const int nb3 = num_blocks(g3), nb4 = num_blocks(g4);
            if (!MERGE) {
                for (int blk = warp; blk < nb3; blk += kWarps)
                    gemm_block<AK, BK, WM, true>(g3, blk, [&](const Tile &t, float (&acc)[4][4], float (&racc)[4][4]) {
                        accumulate_tile(G + LY::w3, kH2P, t, TAN ? racc : acc, scale);
                    });
                __syncthreads();
            }
            // MERGE: dW3 has only nb3 = 2 blocks with the whole symbol range as K, against nb4 = 20 short blocks of dA2: the
            // first half of its K range runs here (those two units first: longest first), the second half in the next phase
            // next to dW2 / dA1 (H2T and dZT stay intact there), where it accumulates into the same G tiles after the barrier.
            const int ng_a = MERGE ? (MVN_TRAIN_SPLIT_DW3 ? (ng + 1) / 2 : ng) : 0;
            const GemmArgs g3a{g3.A, g3.Av, ldn, g3.B, g3.Rb, ldn, S, kH2, ng_a};
            for (int u = warp; u < (MERGE ? nb3 : 0) + nb4; u += kWarps) {
                const int blk = u - (MERGE ? nb3 : 0);
                if (blk >= 0) {
                    gemm_block<AM, BN, GM, false>(g4, blk, [&](const Tile &t, float (&acc)[4][4], float (&racc)[4][4]) {
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            const bool ok = t.r[i] >= 0 && t.c[0] >= 0;
                            float4 d = make_float4(0.f, 0.f, 0.f, 0.f), r = d;
                            if (ok) {
                                const float4 h = lds4(sm.H2T + t.r[i] * ldn + t.c[0]);
                                d = make_float4(h.x > 0.f ? acc[i][0] : 0.f, h.y > 0.f ? acc[i][1] : 0.f,
                                                h.z > 0.f ? acc[i][2] : 0.f, h.w > 0.f ? acc[i][3] : 0.f);
                                *reinterpret_cast<float4 *>(sm.DA2T + t.r[i] * ldn + t.c[0]) = d;
                                if (TAN) {
                                    r = make_float4(h.x > 0.f ? racc[i][0] : 0.f, h.y > 0.f ? racc[i][1] : 0.f,
                                                    h.z > 0.f ? racc[i][2] : 0.f, h.w > 0.f ? racc[i][3] : 0.f);
                                    *reinterpret_cast<float4 *>(sm.RDA2T + t.r[i] * ldn + t.c[0]) = r;
                                }
                            }
                            const float4 q = TAN ? r : d;
                            const float rowsum = sum_ln((q.x + q.y) + (q.z + q.w));
                            if ((lane & 7) == 0 && t.r[i] >= 0 && t.r[i] < kH2) atomicAdd(G + LY::b2 + t.r[i], scale * rowsum);
                        }
                    });
                } else {
                    gemm_block<AK, BK, WM, true>(g3a, u, [&](const Tile &t, float (&acc)[4][4], float (&racc)[4][4]) {
                        accumulate_tile(G + LY::w3, kH2P, t, TAN ? racc : acc, scale);
                    });
                }
            }
        }
        __syncthreads();
        // ---- G5: dW2[o][k] += sum_n DA2T[o][n] H1T[k][n]   (TAN: RdA2 H1^T + dA2 RH1^T)    — in the same phase as
        // ---- G6: dA1[k][n] = (sum_o W2[o][k] DA2T[o][n]) h (1 - h)   (TAN: R = RdH1 s1 + dH1 s1 (1 - 2h) Ra1,
        //          Ra1 = v1[k] y[n] + c1[k]); never stored: dW1[k] += sum_n dA1 y[n], db1[k] += sum_n dA1 in the epilogue
        {
            const GemmArgs g5{sm.DA2T, sm.RDA2T, ldn, sm.H1T, sm.RH1T, ldn, kH2, kH1, ng};
            const GemmArgs g6{W + LY::w2, V + LY::w2, kH1, sm.DA2T, sm.RDA2T, ldn, kH1, n4, kH2P / 4};
            const int nb5 = num_blocks(g5), nb6 = num_blocks(g6);
            // second half of dW3's K range (MERGE only, see above)
            const int ng_a = (MERGE && MVN_TRAIN_SPLIT_DW3) ? (ng + 1) / 2 : ng;
            const GemmArgs g3b{sm.ZT + 4 * ng_a, sm.RZT + 4 * ng_a, ldn, sm.H2T + 4 * ng_a, sm.RH2T + 4 * ng_a, ldn, S, kH2, ng - ng_a};
            const int nb3b = (ng > ng_a) ? num_blocks(g3b) : 0;
            for (int blk = warp; blk < nb5 + nb6 + nb3b; blk += kWarps) {
                if (blk >= nb5 + nb6) {
                    gemm_block<AK, BK, WM, true>(g3b, blk - nb5 - nb6, [&](const Tile &t, float (&acc)[4][4], float (&racc)[4][4]) {
                        accumulate_tile(G + LY::w3, kH2P, t, TAN ? racc : acc, scale);
                    });
                } else if (blk < nb6) {
                    gemm_block<AM, BN, GM, false>(g6, blk, [&](const Tile &t, float (&acc)[4][4], float (&racc)[4][4]) {
                        const bool cok = t.c[0] >= 0;
                        const float4 yv = cok ? lds4(sm.y + t.c[0]) : make_float4(0.f, 0.f, 0.f, 0.f);
                        const float yy[4] = {yv.x, yv.y, yv.z, yv.w};
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            float sw = 0.f, sb = 0.f;
                            if (cok && t.r[i] >= 0) {
                                const float4 h4 = lds4(sm.H1T + t.r[i] * ldn + t.c[0]);
                                const float h[4] = {h4.x, h4.y, h4.z, h4.w};
#pragma unroll
                                for (int j = 0; j < 4; j++) {
                                    const float s1 = h[j] * (1.f - h[j]);
                                    float d = acc[i][j] * s1;
                                    if (TAN) {
                                        const float ra1 = fmaf(V[LY::w1 + t.r[i]], yy[j], V[LY::b1 + t.r[i]]);
                                        d = fmaf(racc[i][j], s1, d * (1.f - 2.f * h[j]) * ra1);
                                    }
                                    sw = fmaf(d, yy[j], sw);
                                    sb += d;
                                }
                            }
                            sw = sum_ln(sw);
                            sb = sum_ln(sb);
                            if ((lane & 7) == 0 && t.r[i] >= 0) {
                                atomicAdd(G + LY::w1 + t.r[i], scale * sw);
                                atomicAdd(G + LY::b1 + t.r[i], scale * sb);
                            }
                        }
                    });
                } else {
                    gemm_block<AK, BK, WM, true>(g5, blk - nb6, [&](const Tile &t, float (&acc)[4][4], float (&racc)[4][4]) {
                        accumulate_tile(G + LY::w2, kH1, t, TAN ? racc : acc, scale);
                    });
                }
            }
        }
    }
    __syncthreads();
    return loss;
}

}  // namespace tg
}  // namespace mvn
