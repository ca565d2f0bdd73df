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
// Data flow per trellis stage for the 128 frames of a CTA tile (one frame = one TMEM lane):
//   producers  : 100 sigmoids -> 3-way split -> tcgen05.st   A pieces [128 x 112] bf16 in TMEM (168 columns)
//   MMA warp   : 6 x 7 tcgen05.mma.kind::f16 (M=128, N=64, K=16), A from TMEM, B (W2 pieces) from shared memory
//                in the canonical K-major no-swizzle layout, D [128 x 64] fp32 in TMEM; tcgen05.commit -> mbarrier
//   consumers  : tcgen05.ld of the frame's D row -> ReLU -> layer 3 (FFMA2) -> ACS on the frame's 8 private
//                path metrics -> decision bit, outputs, BER
// TMEM: a two-slot ring of (64 D + 168 A) columns -> the whole 512-column allocation, one CTA per SM.
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

// ---------------------------------------------------------------------------------------------
// Warp-specialised pipeline.  The sigmoid/split/store work of stage t+1 does not depend on the ACS
// result of stage t (only the decision chain is sequential), so the CTA is split into
//   8 producer warps : y -> 100 sigmoids -> bf16x3 split -> tcgen05.st into A[slot]; one of them issues the MMAs
//   4 consumer warps : tcgen05.ld D[slot] -> ReLU -> layer 3 -> ACS -> decision, outputs, BER
// over a two-slot TMEM ring (2 x (64 D + 168 A) columns).  Producer warps w and w+4 serve the same 32
// frames (TMEM lane quadrant w) and split the hidden units between them; consumer warp 8+w owns those
// frames' path metrics.  mbarriers: d_full[slot] (tcgen05.commit) and slot_free[slot] (128 consumer arrivals).
// ---------------------------------------------------------------------------------------------
namespace tc {
constexpr int kProdWarps = 8, kConsWarps = 4, kThreadsTc = 32 * (kProdWarps + kConsWarps + 1);  // + one MMA-issue warp
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int *timeout_flag) {
    uint32_t done = 0;
    int spins = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (!done && ++spins > (1 << 22)) {  // never expected; keeps a bug from hanging the GPU
            if (timeout_flag) *timeout_flag = 1;
            break;
        }
    }
}
// ---- producer math on fp32x2 pairs of hidden units (k, k+1) --------------------------------------
// staged pair table: sP[k/2] = (w1'[k], w1'[k+1], b1'[k], b1'[k+1]) with w1' = -log2(e) w1 (one LDS.128)
__device__ __forceinline__ u64 sub2(u64 a, u64 b) {  // a - b on both halves: b * (-1) + a, the product is exact
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(b), "l"(0xbf800000bf800000ull), "l"(a));
    return d;
}
__device__ __forceinline__ u64 hi16x2(u64 v) { return v & 0xffff0000ffff0000ull; }
__device__ __forceinline__ uint32_t pack_hi16x2(u64 v) { return __byte_perm(uint32_t(v), uint32_t(v >> 32), 0x7632); }

// (h_k, h_{k+1}) -> the three packed bf16 pair words
__device__ __forceinline__ void split3x2(u64 h, uint32_t &w1, uint32_t &w2, uint32_t &w3) {
    const u64 r1 = sub2(h, hi16x2(h));
    const u64 r2 = sub2(r1, hi16x2(r1));
    w1 = pack_hi16x2(h);
    w2 = pack_hi16x2(r1);
    w3 = pack_hi16x2(r2);  // <= 8 significant bits left in each half: exact
}
__device__ __forceinline__ u64 sigmoid2(uint32_t sP_addr, int pair, u64 yy) {
    u64 w, b;
    lds128(sP_addr + 16 * pair, w, b);
    float x0, x1;
    unpack2(fma2(yy, w, b), x0, x1);
    float d0, d1;
    unpack2(add2(pack2(ex2_approx(x0), ex2_approx(x1)), 0x3f8000003f800000ull), d0, d1);
    return pack2(rcp_approx(d0), rcp_approx(d1));
}
// one k-step (16 hidden units = 8 TMEM columns) of this thread's frame: sigmoid -> split -> 3 x tcgen05.st
template <bool LAST>
__device__ __forceinline__ void produce_chunk(uint32_t sP_addr, int c0, u64 yy, uint32_t tA_lane) {
    uint32_t v[3][8];
#pragma unroll
    for (int c = 0; c < 8; c++) {
        if (!LAST || c < 2) {
            split3x2(sigmoid2(sP_addr, 8 * c0 + c, yy), v[0][c], v[1][c], v[2][c]);
        } else {  // k = 100 is the bias column (1.0 = bf16 0x3F80 in the low half), k > 100 is zero padding
            v[0][c] = (c == 2) ? 0x00003f80u : 0u;
            v[1][c] = 0u;
            v[2][c] = 0u;
        }
    }
#pragma unroll
    for (int i = 0; i < 3; i++) tmem_st8(tA_lane + i * kACols + c0 * 8, v[i]);
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
}  // namespace tc

template <int L>
__global__ void __launch_bounds__(tc::kThreadsTc, 1) vnet_decode_tc_kernel(VnetParams p, int *timeout_flag) {
    static_assert(L <= 4, "tcgen05 variant: register trellis, one layer-3 chunk");
    using D = TrellisDims<L>;
    using W = VnetSmem<L>;
    constexpr int S = D::S, C = D::C, NW = tc::kProdWarps + tc::kConsWarps;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(8) uint64_t d_full[2], slot_free[2], a_full[2];
    uint8_t *sB = smem_raw;
    float *tiles = reinterpret_cast<float *>(smem_raw + 3 * tc::kBPieceBytes);  // one 32x32 tile per warp
    float *sW = tiles + NW * kTileFloats;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, quad = warp & 3;
    const bool producer = warp < tc::kProdWarps;
    const bool mma_warp = warp == NW;
    float *tile = tiles + warp * kTileFloats;

    for (int idx = tid; idx < tc::kN * tc::kK; idx += tc::kThreadsTc) {
        const int n = idx / tc::kK, k = idx % tc::kK;
        float w = 0.f;
        if (n < kH2) w = k < kH1 ? p.w.w2[n * kH1 + k] : (k == kH1 ? p.w.b2[n] : 0.f);
        uint32_t q[3];
        tc::split3(w, q[0], q[1], q[2]);
        const int off = (k / 8) * (tc::kN / 8) * 128 + (n / 8) * 128 + (n % 8) * 16 + (k % 8) * 2;
#pragma unroll
        for (int i = 0; i < 3; i++) *reinterpret_cast<uint16_t *>(sB + i * tc::kBPieceBytes + off) = uint16_t(q[i] >> 16);
    }
    stage_weights<L>(sW, p.w, tid, tc::kThreadsTc);
    float *sP = sW + W::kFloats;  // [56][4] pair table for the producers' packed sigmoid
    for (int i = tid; i < tc::kK / 2; i += tc::kThreadsTc) {
        const float kNegLog2e = -1.4426950408889634f;
        const int k = 2 * i;
        sP[4 * i + 0] = k < kH1 ? p.w.w1[k] * kNegLog2e : 0.f;
        sP[4 * i + 1] = k + 1 < kH1 ? p.w.w1[k + 1] * kNegLog2e : 0.f;
        sP[4 * i + 2] = k < kH1 ? p.w.b1[k] * kNegLog2e : 0.f;
        sP[4 * i + 3] = k + 1 < kH1 ? p.w.b1[k + 1] * kNegLog2e : 0.f;
    }
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < 2; s++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&d_full[s])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(&slot_free[s])), "n"(32 * tc::kConsWarps));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(&a_full[s])), "n"(32 * tc::kProdWarps));
        }
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&tmem_base_s)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base_s;
    const uint32_t lane_base = uint32_t(quad * 32) << 16;
    const uint32_t sB_addr = smem_addr(sB);
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(tc::kN >> 3) << 17) | (uint32_t(tc::kM >> 4) << 24);
    const Wt<kSmem> wt{smem_addr(sW)};
    const uint32_t sP_addr = smem_addr(sP);
    const bool vec_in = is_vec_ok(p.y, p.T, p.T);
    const bool vec_out = p.out_format == MVN_OUT_F32 && is_vec_ok(p.decoded, p.T, p.T);
    const bool vec_tgt = p.target && is_vec_ok(p.target, p.target_T, p.target_T);
    const int n_words = (p.T + 31) / 32;
    const int64_t n_cta_tiles = (p.n_warp_tiles + 3) / 4;  // CTA tile = 4 warp tiles of 32 frames
    uint32_t n = 0;                                        // running stage counter: slot = n & 1, use = n >> 1

    if (producer) {
        // hidden units of this producer warp: warps 0-3 take k-steps 0..3, warps 4-7 take k-steps 4..6
        for (int64_t ct = blockIdx.x; ct < n_cta_tiles; ct += gridDim.x) {
            const int64_t row0 = (ct * 4 + quad) * 32;
            for (int t0 = 0; t0 < p.T; t0 += 32) {
                const int t_end = min(32, p.n_stages - t0);
                if (t_end <= 0) continue;
                warp_load_tile(p.y, p.B, p.T, p.T, row0, t0, tile, lane, vec_in);
#pragma unroll 1
                for (int tt = 0; tt < t_end; tt++, n++) {
                    const uint32_t slot = n & 1, use = n >> 1;
                    const uint32_t tA = tmem + slot * tc::kGroupCols + tc::kN;
                    const float yv = tile[lane * kTileLd + tt];
                    tc::mbar_wait(smem_addr(&slot_free[slot]), (use & 1) ^ 1, timeout_flag);
                    asm volatile("tcgen05.fence::after_thread_sync;");
                    {
                        const u64 yy = pack2(yv, yv);
                        if (warp < 4) {
#pragma unroll 1
                            for (int c0 = 0; c0 < 4; c0++) tc::produce_chunk<false>(sP_addr, c0, yy, tA + lane_base);
                        } else {
#pragma unroll 1
                            for (int c0 = 4; c0 < tc::kKSteps - 1; c0++) tc::produce_chunk<false>(sP_addr, c0, yy, tA + lane_base);
                            tc::produce_chunk<true>(sP_addr, tc::kKSteps - 1, yy, tA + lane_base);
                        }
                    }
                    asm volatile("tcgen05.wait::st.sync.aligned;");
                    asm volatile("tcgen05.fence::before_thread_sync;");
                    tc::mbar_arrive(smem_addr(&a_full[slot]));  // 256 arrivals release the MMA warp
                    __syncwarp();
                }
                __syncwarp();
            }
        }
    } else if (mma_warp) {
        // one warp does nothing but issue: per stage 6 x 7 MMAs (about 2 000 cycles of issue time, which a
        // producer warp would otherwise spend blocked), then tcgen05.commit -> d_full[slot]
        for (int64_t ct = blockIdx.x; ct < n_cta_tiles; ct += gridDim.x) {
            for (int t0 = 0; t0 < p.T; t0 += 32) {
                const int t_end = min(32, p.n_stages - t0);
#pragma unroll 1
                for (int tt = 0; tt < t_end; tt++, n++) {
                    const uint32_t slot = n & 1, use = n >> 1;
                    tc::mbar_wait(smem_addr(&a_full[slot]), use & 1, timeout_flag);
                    asm volatile("tcgen05.fence::after_thread_sync;");
                    if (lane == 0) {
                        const uint32_t tD = tmem + slot * tc::kGroupCols, tA = tD + tc::kN;
                        uint32_t accum = 0;
#pragma unroll
                        for (int t = 5; t >= 0; t--) {
                            const int pa = (t == 2 || t == 4) ? 1 : (t == 5 ? 2 : 0);
                            const int pb = (t == 1 || t == 4) ? 1 : (t == 3 ? 2 : 0);
#pragma unroll
                            for (int j = 0; j < tc::kKSteps; j++) {
                                tc::mma_f16_ts(tD, tA + pa * tc::kACols + j * 8,
                                               tc::b_desc(sB_addr + pb * tc::kBPieceBytes + uint32_t(2 * j) * tc::kLBO), idesc, accum);
                                accum = 1;
                            }
                        }
                        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                            smem_addr(&d_full[slot])));
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        RegTrellis<L> tr;
        ErrAcc acc;
        for (int64_t ct = blockIdx.x; ct < n_cta_tiles; ct += gridDim.x) {
            const int64_t row0 = (ct * 4 + quad) * 32;
            const int64_t b = row0 + lane;
            tr.reset();
            unsigned frame_bit_errs = 0;
            for (int t0 = 0; t0 < p.T; t0 += 32) {
                uint32_t bits = 0;
                const int t_end = min(32, p.n_stages - t0);
#pragma unroll 1
                for (int tt = 0; tt < t_end; tt++, n++) {
                    const uint32_t slot = n & 1, use = n >> 1;
                    const uint32_t tD = tmem + slot * tc::kGroupCols;
                    bits |= tr.decide() << tt;
                    tc::mbar_wait(smem_addr(&d_full[slot]), use & 1, timeout_flag);
                    asm volatile("tcgen05.fence::after_thread_sync;");
                    float h2[1][kH2];
                    {
                        float d[64];
                        tc::tmem_ld16(tD + 0 + lane_base, d);
                        tc::tmem_ld16(tD + 16 + lane_base, d + 16);
                        tc::tmem_ld16(tD + 32 + lane_base, d + 32);
                        tc::tmem_ld2(tD + 48 + lane_base, d + 48);
                        asm volatile("tcgen05.wait::ld.sync.aligned;");
                        asm volatile("tcgen05.fence::before_thread_sync;");
                        tc::mbar_arrive(smem_addr(&slot_free[slot]));  // D and A of this slot may be refilled
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
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

template <int L>
constexpr size_t tc_smem_bytes() {
    return size_t(3) * tc::kBPieceBytes +
           (size_t(tc::kProdWarps + tc::kConsWarps) * kTileFloats + VnetSmem<L>::kFloats + 4 * (tc::kK / 2)) * sizeof(float);
}

}  // namespace mvn
