// tcgen05 variant of the fused ViterbiNet kernel (memory_length <= 4): layer 2 of the priors MLP
// ([symbols,100] x [100,50], 85 % of the flops) runs on the 5th-generation tensor cores, everything
// else (sigmoid, ReLU, layer 3, ACS, decision) stays on the CUDA cores of the same CTA.
//
// fp32 parity on bf16 tensor cores.  h1 and W2 are each split EXACTLY into three bf16 pieces
// (x = p1 + p2 + p3, 8 significant bits each, by mantissa truncation), and
//     h1 . w  =  p1q1 + p1q2 + p2q1 + p1q3 + p2q2 + p3q1  + O(2^-24 |h1||w|)
// is accumulated in fp32 in TMEM by six chains of kind::f16 MMAs (bf16 products are exact in fp32).
// Measured against fp64 (tools/tc_test.cu): 2e-7 of the row maximum, the same class as the FMA path.
// The bias b2 rides along as column k=100 of B against a constant-1 column of A.
//
// Data flow per trellis stage, per group of 128 frames (one frame per thread = one TMEM lane):
//   CUDA cores : 100 sigmoids -> 3-way split -> tcgen05.st   A pieces [128 x 112] bf16 in TMEM (168 columns)
//   one thread : 6 x 7 tcgen05.mma.kind::f16 (M=128, N=64, K=16), A from TMEM, B (W2 pieces) from shared memory
//                in the canonical K-major no-swizzle layout, D [128 x 64] fp32 in TMEM; tcgen05.commit -> mbarrier
//   CUDA cores : tcgen05.ld of the thread's D row -> ReLU -> layer 3 (FFMA2, weights through the constant
//                bank) -> ACS on the thread's 8 private path metrics -> decision bit
// Two groups (warps 0-3 and 4-7) share the SM: while the tensor core works for one, the CUDA cores work for
// the other.  TMEM: 2 x (64 + 168) columns -> the whole 512-column allocation, one CTA per SM.
#pragma once
#include "vnet_mlp.cuh"

namespace mvn {

namespace tc {
constexpr int kM = 128, kN = 64, kK = 112;    // MMA tile: frames x padded outputs x padded hidden units (+ bias column)
constexpr int kKSteps = kK / 16;              // UMMA_K = 16 for bf16
constexpr int kACols = kK / 2;                // 32-bit TMEM columns per A piece (two bf16 per column)
constexpr int kGroupCols = 256;               // D (64) + 3 x 56 A columns, rounded to the allocation granularity
constexpr uint32_t kLBO = (kN / 8) * 128;     // bytes between consecutive 16-byte K chunks (k-chunk stride)
constexpr uint32_t kSBO = 128;                // bytes between 8-row groups along N
constexpr int kBPieceBytes = (kK / 8) * (kN / 8) * 128;
constexpr int kThreads = 256;

__device__ __forceinline__ uint32_t hi16(float x) { return __float_as_uint(x) & 0xffff0000u; }

// exact split x = p1 + p2 + p3 (each representable in bf16); returns the three bf16 bit patterns << 16
__device__ __forceinline__ void split3(float x, uint32_t &b1, uint32_t &b2, uint32_t &b3) {
    b1 = hi16(x);
    const float r1 = x - __uint_as_float(b1);
    b2 = hi16(r1);
    b3 = __float_as_uint(r1 - __uint_as_float(b2));  // <= 8 significant bits left: exact in bf16
}
// (lo, hi) bf16 pair from two fp32 patterns whose low 16 bits are zero / to be dropped
__device__ __forceinline__ uint32_t pack_hi16(uint32_t lo, uint32_t hi) { return __byte_perm(lo, hi, 0x7632); }

__device__ __forceinline__ uint64_t b_desc(uint32_t saddr) {
    return uint64_t((saddr >> 4) & 0x3fff) | (uint64_t((kLBO >> 4) & 0x3fff) << 16) |
           (uint64_t((kSBO >> 4) & 0x3fff) << 32) | (uint64_t(1) << 46);  // version 1, SWIZZLE_NONE
}
__device__ __forceinline__ void mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc));
}
__device__ __forceinline__ void tmem_st8(uint32_t addr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(addr), "r"(v[0]),
                 "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]));
}
__device__ __forceinline__ void tmem_ld16(uint32_t addr, float *r) {
    uint32_t u[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
          "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
        : "r"(addr));
#pragma unroll
    for (int i = 0; i < 16; i++) r[i] = __uint_as_float(u[i]);
}
__device__ __forceinline__ void tmem_ld2(uint32_t addr, float *r) {
    uint32_t u0, u1;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(u0), "=r"(u1) : "r"(addr));
    r[0] = __uint_as_float(u0);
    r[1] = __uint_as_float(u1);
}
}  // namespace tc

template <int L>
__global__ void __launch_bounds__(tc::kThreads, 1) vnet_decode_tc_kernel(VnetParams p, int *timeout_flag) {
    static_assert(L <= 4, "tcgen05 variant: register trellis, one layer-3 chunk");
    using D = TrellisDims<L>;
    using W = VnetSmem<L>;
    constexpr int S = D::S, C = D::C, WARPS = tc::kThreads / 32;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(8) uint64_t bars[2];
    uint8_t *sB = smem_raw;                                                      // 3 x kBPieceBytes
    float *tiles = reinterpret_cast<float *>(smem_raw + 3 * tc::kBPieceBytes);   // one 32x32 tile per warp
    float *sW = tiles + WARPS * kTileFloats;                                     // staged (w1,b1), W3T, b3 (VnetSmem layout)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, group = warp >> 2;
    float *tile = tiles + warp * kTileFloats;

    // ---- W2 (and b2 as column k=100) -> three bf16 pieces in the canonical K-major layout
    for (int idx = tid; idx < tc::kN * tc::kK; idx += tc::kThreads) {
        const int n = idx / tc::kK, k = idx % tc::kK;
        float w = 0.f;
        if (n < kH2) w = k < kH1 ? p.w.w2[n * kH1 + k] : (k == kH1 ? p.w.b2[n] : 0.f);
        uint32_t q[3];
        tc::split3(w, q[0], q[1], q[2]);
        const int off = (k / 8) * (tc::kN / 8) * 128 + (n / 8) * 128 + (n % 8) * 16 + (k % 8) * 2;
#pragma unroll
        for (int i = 0; i < 3; i++) *reinterpret_cast<uint16_t *>(sB + i * tc::kBPieceBytes + off) = uint16_t(q[i] >> 16);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&bars[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&bars[1])));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&tmem_base_s)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;");  // generic-proxy writes of B -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base_s;
    const uint32_t lane_base = uint32_t((warp & 3) * 32) << 16;
    const uint32_t tD = tmem + group * tc::kGroupCols, tA = tD + tc::kN;
    const uint32_t bar = smem_addr(&bars[group]);
    const uint32_t sB_addr = smem_addr(sB);
    // D=F32, A=B=BF16, both K-major, N>>3 at bit 17, M>>4 at bit 24
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(tc::kN >> 3) << 17) | (uint32_t(tc::kM >> 4) << 24);

    // (w1,b1), W3 and b3 are read with warp-broadcast LDS: with one frame per lane ptxas serves the
    // constant-bank alternative with per-lane LDC (measured 4x slower than the whole FMA kernel).
    stage_weights<L>(sW, p.w, tid, tc::kThreads);
    __syncthreads();
    const Wt<kSmem> wt{smem_addr(sW)};
    RegTrellis<L> tr;
    const bool vec_in = is_vec_ok(p.y, p.T, p.T);
    const bool vec_out = p.out_format == MVN_OUT_F32 && is_vec_ok(p.decoded, p.T, p.T);
    const bool vec_tgt = p.target && is_vec_ok(p.target, p.target_T, p.target_T);
    const int n_words = (p.T + 31) / 32;
    ErrAcc acc;
    uint32_t phase = 0;

    const int64_t n_cta_tiles = (p.n_warp_tiles + WARPS - 1) / WARPS;  // warp tile = 32 frames
    for (int64_t ct = blockIdx.x; ct < n_cta_tiles; ct += gridDim.x) {
        const int64_t row0 = (ct * WARPS + warp) * 32;
        const int64_t b = row0 + lane;
        tr.reset();
        unsigned frame_bit_errs = 0;
        for (int t0 = 0; t0 < p.T; t0 += 32) {
            uint32_t bits = 0;
            const int t_end = min(32, p.n_stages - t0);
            if (t_end > 0) {
                warp_load_tile(p.y, p.B, p.T, p.T, row0, t0, tile, lane, vec_in);
#pragma unroll 1
                for (int tt = 0; tt < t_end; tt++) {
                    const float yv = tile[lane * kTileLd + tt];
                    // ---- layer 1: sigmoid, exact 3-way bf16 split, pieces -> TMEM (this thread's lane)
#pragma unroll
                    for (int c0 = 0; c0 < tc::kKSteps; c0++) {
                        uint32_t v[3][8];
#pragma unroll
                        for (int c = 0; c < 8; c++) {
                            uint32_t q0[3], q1[3];
#pragma unroll
                            for (int h = 0; h < 2; h++) {
                                const int k = 16 * c0 + 2 * c + h;
                                float x = 0.f;
                                if (k < kH1) {
                                    const float2 wb = wt.f2(2 * k);
                                    x = rcp_approx(1.f + ex2_approx(fmaf(yv, wb.x, wb.y)));
                                } else if (k == kH1) {
                                    x = 1.f;  // bias column
                                }
                                if (h == 0) tc::split3(x, q0[0], q0[1], q0[2]);
                                else tc::split3(x, q1[0], q1[1], q1[2]);
                            }
#pragma unroll
                            for (int i = 0; i < 3; i++) v[i][c] = tc::pack_hi16(q0[i], q1[i]);
                        }
#pragma unroll
                        for (int i = 0; i < 3; i++) tc::tmem_st8(tA + i * tc::kACols + c0 * 8 + lane_base, v[i]);
                    }
                    asm volatile("tcgen05.wait::st.sync.aligned;");
                    asm volatile("tcgen05.fence::before_thread_sync;");
                    asm volatile("bar.sync %0, 128;" ::"r"(1 + group));
                    // ---- layer 2 on the tensor core: one thread of the group issues 42 MMAs
                    if ((warp & 3) == 0 && lane == 0) {
                        asm volatile("tcgen05.fence::after_thread_sync;");
                        uint32_t accum = 0;
#pragma unroll
                        for (int t = 5; t >= 0; t--) {  // smallest terms first
                            const int pa = (t == 2 || t == 4) ? 1 : (t == 5 ? 2 : 0);
                            const int pb = (t == 1 || t == 4) ? 1 : (t == 3 ? 2 : 0);
#pragma unroll
                            for (int j = 0; j < tc::kKSteps; j++) {
                                tc::mma_f16_ts(tD, tA + pa * tc::kACols + j * 8,
                                               tc::b_desc(sB_addr + pb * tc::kBPieceBytes + uint32_t(2 * j) * tc::kLBO), idesc, accum);
                                accum = 1;
                            }
                        }
                        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar));
                    }
                    __syncwarp();
                    bits |= tr.decide() << tt;  // uses the metrics entering this stage; overlaps the MMAs
                    {
                        uint32_t done = 0;
                        int spins = 0;
                        while (!done) {
                            asm volatile(
                                "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                                : "=r"(done)
                                : "r"(bar), "r"(phase));
                            if (!done && ++spins > (1 << 22)) {  // never expected; keeps a bug from hanging the GPU
                                if (timeout_flag) *timeout_flag = 1;
                                break;
                            }
                        }
                        phase ^= 1;
                    }
                    asm volatile("tcgen05.fence::after_thread_sync;");
                    // ---- D row -> ReLU -> layer 3 -> ACS
                    float h2[1][kH2];
                    {
                        float d[64];
                        tc::tmem_ld16(tD + 0 + lane_base, d);
                        tc::tmem_ld16(tD + 16 + lane_base, d + 16);
                        tc::tmem_ld16(tD + 32 + lane_base, d + 32);
                        tc::tmem_ld2(tD + 48 + lane_base, d + 48);
                        asm volatile("tcgen05.wait::ld.sync.aligned;");
#pragma unroll
                        for (int o = 0; o < kH2; o++) h2[0][o] = fmaxf(d[o], 0.f);
                    }
                    float pr[1][C];
                    mlp_out_chunk<L, 1, kSmem>(wt, 0, h2, pr);
                    float cost[C];
#pragma unroll
                    for (int i = 0; i < C; i++) cost[i] = -pr[0][i];  // vnet_detector.py:57
                    tr.template step_chunk<0>(cost);
                    tr.commit();
                    if (p.priors_out && b < p.B) {
                        float *dst = p.priors_out + (b * p.T + t0 + tt) * S;
#pragma unroll
                        for (int i = 0; i < C; i++) dst[i] = pr[0][i];
                    }
                }
                __syncwarp();
            }
            if (p.decoded) {
                if (p.out_format == MVN_OUT_F32)
                    warp_store_bits_f32(static_cast<float *>(p.decoded), p.B, p.T, p.T, row0, t0, bits, lane, vec_out);
                else if (b < p.B)
                    static_cast<uint32_t *>(p.decoded)[b * n_words + t0 / 32] = bits;
            }
            if (p.target && t0 < p.target_T) {
                warp_load_tile(p.target, p.B, p.target_T, p.target_T, row0, t0, tile, lane, vec_tgt);
                frame_bit_errs += tile_bit_errors(tile + lane * kTileLd, bits, p.target_T - t0);
                __syncwarp();
            }
        }
        if (p.target) {
            const bool counted = b < p.B && !(p.pilot_period > 0 && b % p.pilot_period == 0);
            if (counted) {
                acc.bit_errs += frame_bit_errs;
                acc.frame_errs += frame_bit_errs ? 1u : 0u;
                acc.bits += unsigned(p.target_T);
                acc.frames += 1u;
            }
            acc.flush(p.counters);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

template <int L>
constexpr size_t tc_smem_bytes() {
    return size_t(3) * tc::kBPieceBytes + (size_t(tc::kThreads / 32) * kTileFloats + VnetSmem<L>::kFloats) * sizeof(float);
}

}  // namespace mvn
