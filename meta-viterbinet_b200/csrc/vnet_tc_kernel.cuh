// tcgen05 variant of the fused ViterbiNet kernel (memory_length 1..8, the default): layers 2 and 3 of the priors MLP
// ([symbols,100] x [100,50] and [symbols,50] x [50,S], 99 % of the flops) run on the 5th-generation tensor cores;
// the sigmoid, the ReLU / operand splits, the ACS recursion and the decision stay on the CUDA cores of the same CTA.
//
// fp32 parity on fp16 tensor cores.  Each operand is split into two fp16 pieces with a scaled remainder,
//     x = hi + lo / 2048,   hi = fp16(x),   lo = fp16((x - hi) * 2048)      (x - hi is exact in fp32),
// and   h1 . w = hi_h hi_w + (hi_h lo_w + lo_h hi_w) / 2048 + O(2^-22 |h1||w|)   is accumulated in fp32 in TMEM;
// fp16 products are exact in fp32 and the 2^11 scaling keeps the remainders normal.  The correction chains either get
// an accumulator of their own (D_corr, combined with D_main by the reader) or run first and are folded into the main
// chain by the tensor core (scale-input-d: D = A B + D 2^-11).  On the reference's fixtures the priors are 2.4e-7 of the
// row maximum away from fp64 — closer than a plain fp32 dot product (4.4e-7), see DESIGN.md §5.1.
// (A first version used an exact three-way bf16 split with six chains: same accuracy, twice the tensor work.)
// The biases ride along as a column of B against a constant column of A.
// Layer 2 takes its A operand PRE-SCALED: the producers compute g = 2048 h1 directly (see denom2), so the remainder
// g - fp16(g) is a normal fp16 number without a scaling multiply; the layer-2 result is then 2048 (W2 h1 + b2), which
// is what the ReLU / split of layer 3 wants as input.
//
// Warp-specialised pipeline over a two-slot TMEM ring; one frame = one TMEM lane, 128 frames per CTA tile:
//  12 producer warps : y -> 100 sigmoids -> fp16 hi/lo split -> tcgen05.st  A_hi, A_lo [128 x 112] (2 x 56 columns)
//                      (three warps per TMEM lane quadrant serve the same 32 frames and split the hidden units; the
//                       bias / padding columns are written once per launch)
//   4 MMA warps      : (one per scheduler, taking the stages round-robin) the layer-2 tcgen05.mma.kind::f16 of a stage (M=128, K=16, N=64: A_hi x W2_lo, A_hi x W2_hi with
//                      scale-input-d, A_lo x W2_hi = 21 MMAs; L >= 7: 7 of N=128 + 7 of N=64 into D_main | D_corr),
//                      A from TMEM, B from shared memory in the canonical K-major no-swizzle layout; tcgen05.commit -> d_full
//   4 converter warps: tcgen05.ld of the layer-2 result -> ReLU -> the same hi/lo split -> tcgen05.st as the A operand
//                      of LAYER 3, which also runs on the tensor core (8-12 MMAs, W3 pieces and b3 in shared memory,
//                      issued by one converter thread, the four warps taking turns; 256 states: two passes)
//   4 consumer warps : tcgen05.ld of the priors -> ACS on the frame's private path metrics, decision bit (or survivor
//                      masks + in-kernel traceback), outputs, BER.  Converters and consumers pipeline across the two
//                      slots: stage n+1 is converted while layer 3 of stage n runs and its ACS is done.
// The sigmoid/split/MMA work of later stages does not depend on the ACS result of earlier ones (only the
// decision chain is sequential), which is what lets producers and the tensor core run ahead.
// TMEM: 2 x 240 columns of the 512-column allocation, one CTA per SM; the slot layouts are described at tc::Lay.
#pragma once
#include <cuda_fp16.h>

#include "vnet_mlp.cuh"

namespace mvn {

#ifndef MVN_PROD_PARK
#define MVN_PROD_PARK false   // producers reach this wait with their stage computed: polling wakes them ~200 cycles sooner (+1 %)
#endif
#ifndef MVN_TC_MIXED_FMA
#define MVN_TC_MIXED_FMA 0   // FHFMA (fma.f32.f16) in the hi/lo splits: one instruction less per pair, but measured 3 % SLOWER (L=4 17.1 -> 16.5 G sym/s)
#endif
#ifndef MVN_TC_EARLY_PROBE
#define MVN_TC_EARLY_PROBE 1
#endif
#ifndef MVN_TC_LATE_PUBLISH
#define MVN_TC_LATE_PUBLISH 1
#endif
#ifndef MVN_TC_CLAMP_FREE
#define MVN_TC_CLAMP_FREE 0   // clamp-free producer path behind a per-stage |y| vote: L=4 17.6 vs 18.0 G sym/s (second copy of the stage body, spills)
#endif
#ifndef MVN_TC_RCP_SHARE
#define MVN_TC_RCP_SHARE 4
#endif
#ifndef MVN_TC_CONV_PREFETCH
#define MVN_TC_CONV_PREFETCH 0   // converter keeps the next chunk's tcgen05.ld in flight: 973 -> 914 cycles per stage, throughput unchanged (17.7 vs 18.0)
#endif
#ifndef MVN_TC_PARK_MINL
#define MVN_TC_PARK_MINL 7   // producers park (instead of polling) from this memory length on
#endif
#ifndef MVN_TC_CONS_PAIRED
#define MVN_TC_CONS_PAIRED 1
#endif
#ifndef MVN_TC_EXPERIMENT
#define MVN_TC_EXPERIMENT 0   // 1, 2: bound-finding builds of the producers, never shipped
#endif
#ifndef MVN_CONS_PARK
#define MVN_CONS_PARK true
#endif
namespace tc {
constexpr int kM = 128, kN = 64, kK = 112;    // MMA tile: frames x padded outputs x padded hidden units (+ bias column)
constexpr int kKSteps = kK / 16;              // UMMA_K = 16 for fp16
constexpr int kACols = kK / 2;                // 32-bit TMEM columns per A piece (two fp16 per column)
constexpr int kSlotCols = 256;                // slot stride; 240 used: D_main | D_corr | A_hi | A_lo
constexpr int oDm = 0, oDc = kN, oAh = 2 * kN, oAl = 2 * kN + kACols;   // round-1 layout (L >= 7): D_main | D_corr | A_hi | A_lo
// TMEM layouts of one slot (240 of its 256 columns):
//   mode 0 (L == 7)      D_main (64) | D_corr (64) | A_hi (56) | A_lo (56); h2 reuses the first 32 columns of A_hi / A_lo,
//                        so the producers may refill A only after the consumers have released the slot
//   mode 1 ("merged")    D (64) | A_hi (56) | A_lo (56) | H2_hi (32) | H2_lo (32): ONE layer-2 accumulator (the correction
//                        chain is folded in on the tensor core with scale-input-d, 21 MMAs of N = 64)
//   mode 2 ("in place")  D_main | D_corr | A_hi | A_lo as mode 0 (7 MMAs of N = 128 + 7 of N = 64), but the converter writes
//                        the h2 pieces of k-step j over the 16 D_corr columns it has just read (hi: 8 columns, lo: 8)
//   mode 3 (L == 8)      A_hi (56) | A_lo (56) ONCE, then per slot D_main (64) | D_corr (64) | H2_hi (32) | H2_lo (32) = 496 columns:
//                        128 / 256 states need all 128 accumulator columns for the priors, so h2 gets columns of its own by
//                        giving up the second A buffer — the producers hold the next stage in registers anyway and only
//                        need the layer-2 MMAs of the PREVIOUS stage to have read A
// In modes 1, 2 and 3 the A columns are free again as soon as the layer-2 MMAs have read them: the producers wait for a
// d_full, not for the consumers (mode 0: for the consumers of the slot's previous use).
#ifndef MVN_TC_LAYOUT
#define MVN_TC_LAYOUT 1
#endif
#ifndef MVN_TC_LAYOUT_L7
#define MVN_TC_LAYOUT_L7 0   // 128 states: 0 (round-1 layout) or 3
#endif
#ifndef MVN_TC_LAYOUT_L8
#define MVN_TC_LAYOUT_L8 3   // 256 states: 0 or 3
#endif
template <int MODE>
struct Lay {
    // TMEM column of the slot's accumulators (D) and of the layer-2 A operand (A_hi; A_lo follows kACols later).
    // Mode 3 has ONE A buffer for both slots.
    __host__ __device__ static constexpr int d_col(int slot) { return MODE == 3 ? 2 * kACols + slot * (2 * kN + 64) : slot * kSlotCols; }
    __host__ __device__ static constexpr int a_col(int slot) { return MODE == 3 ? 0 : slot * kSlotCols + (MODE == 1 ? kN : 2 * kN); }
    // layer-3 A operand (h2): column of the hi / lo piece of k-step j (16 hidden units = 8 columns), relative to d_col
    __host__ __device__ static constexpr int hh(int j) {
        return MODE == 0 ? 2 * kN + 8 * j : MODE == 1 ? kN + 2 * kACols + 8 * j : MODE == 2 ? kN + 16 * j : 2 * kN + 8 * j;
    }
    __host__ __device__ static constexpr int hl(int j) {
        return MODE == 0 ? 2 * kN + kACols + 8 * j : MODE == 1 ? kN + 2 * kACols + 32 + 8 * j : MODE == 2 ? kN + 16 * j + 8 : 2 * kN + 32 + 8 * j;
    }
};
constexpr float kScale = 2048.f, kInvScale = 1.f / 2048.f;
// B operand of layer 2: the hi and lo pieces of W2 stacked along N (rows 0..63 hi, 64..127 lo), so ONE N=128 MMA per
// k-step produces D_main | D_corr side by side; the A_lo x B_hi chain reads rows 0..63 of the same buffer (N=64).
// Every tcgen05.mma costs >= 48 cycles whatever its N (tools/mma_rate.cu: N<=64 48, N=128 66, N=256 131), so
// stacking saves a third of the tensor-pipe time of this small-N problem.
constexpr uint32_t kLBO = (2 * kN / 8) * 128; // bytes between consecutive 16-byte K chunks (k-chunk stride)
constexpr uint32_t kSBO = 128;                // bytes between 8-row groups along N
constexpr int kBBytes = (kK / 8) * (2 * kN / 8) * 128;
#ifndef MVN_TC_PROD_WARPS
#define MVN_TC_PROD_WARPS 12   // 16 (four per quadrant, 72 registers) measured the same: the schedulers are issue-bound, not latency-bound
#endif
constexpr int kProdWarps = MVN_TC_PROD_WARPS, kConvWarps = 4, kConsWarps = 4;   // producers | h2 converters | ACS consumers
constexpr int kProdParts = kProdWarps / 4;    // producer warps per TMEM lane quadrant (3 or 4)
static_assert(kProdWarps == 12 || kProdWarps == 16, "24 double-pairs of hidden units split 8|8|8 or 6|6|6|6");
// MMA-issue warps: stage n's layer-2 MMAs are issued by warp (n mod kMmaWarps) of them, its layer-3 MMAs by converter warp
// ((n + 2) mod 4), so every scheduler (warp w runs on scheduler w mod 4, and serves TMEM lane quadrant w mod 4) carries
// the same share of tcgen05.mma issue.  With ONE issue warp the pipeline trace showed the three producer warps on its
// scheduler a full stage behind the other nine: a tcgen05.mma waiting for the tensor pipe holds up its scheduler.
#ifndef MVN_TC_MMA_WARPS
#define MVN_TC_MMA_WARPS 4   // 1: L=4 17.4, 4: 17.8 G sym/s (once the producers' barrier probe was off the critical path; before that 17.10 vs 17.13)
#endif
constexpr int kMmaWarps = MVN_TC_MMA_WARPS;
constexpr int kThreadsTc = 32 * (kProdWarps + kConvWarps + kConsWarps + kMmaWarps);
// Warp roles per memory length.  At 128 / 256 states one consumer warp per 32 frames is busy 70-90 % of a stage (pipeline
// traces), so MVN_TC_DUAL_CONS=1 gives every TMEM lane quadrant TWO consumer warps that split the source-state chunks of a
// stage (= its destination states; the metrics are in shared memory, only the even / odd minima of the decision are
// exchanged, one named barrier per pair and stage).  Bit-exact, but NOT faster — 12 producers + 8 consumers (28 warps, 72
// registers): 11.4 / 5.64 G sym/s at 128 / 256 states against 11.8 / 5.64 with one consumer warp; 8 + 8 (24 warps): 11.1 /
// 5.04.  The two warps of a pair sit on the same scheduler (a warp reads the TMEM lanes of quadrant warp % 4) and the time
// of a chunk is the latency of its tcgen05.ld / LDS / STS through that scheduler's MIO queue, which they share.  Off.
#ifndef MVN_TC_DUAL_CONS
#define MVN_TC_DUAL_CONS 0
#endif
#ifndef MVN_TC_DUAL_PROD
#define MVN_TC_DUAL_PROD 12
#endif
template <int L>
struct Roles {
    static constexpr bool DUAL = MVN_TC_DUAL_CONS && L >= 7 && kProdWarps == 12;
    // MVN_TC_DUAL_PROD: producer warps of the DUAL instances.  8 keeps the block at 24 warps of 80 registers, but a warp's 12
    // double-pairs then go through the registers in two halves and the second half is computed after the slot has been
    // acquired — on the critical loop at 128 states (measured: slower than one consumer warp).  12 = 28 warps of 72 registers.
    static constexpr int PW = DUAL ? MVN_TC_DUAL_PROD : kProdWarps, KW = DUAL ? 8 : kConsWarps, PARTS = PW / 4;
    static constexpr int kWarps = PW + kConvWarps + KW + kMmaWarps, kThreads = 32 * kWarps;
    static constexpr int kMaxReg = (kWarps > 24) ? 72 : 80;   // 7 warps per scheduler need <= 73 registers
};
// layer 3 on the tensor core as well: D2[128 x 16] = h2[128 x 64] W3^T, K2 = 50 hidden units + bias column, padded
// (N2 = max(16, n_states) output columns, so up to 64 states fit the slot's 64-column D regions)
constexpr int kK2 = 64, kK2Steps = kK2 / 16;
__host__ __device__ constexpr int n2_of(int S) { return S < 16 ? 16 : S; }
__host__ __device__ constexpr int b2_bytes(int S) { return (kK2 / 8) * (2 * n2_of(S) / 8) * 128; }  // hi | lo stacked along N

__device__ __forceinline__ void split_f16(float x, uint16_t &hi, uint16_t &lo) {
    const __half h = __float2half_rn(x);
    const __half l = __float2half_rn((x - __half2float(h)) * kScale);
    hi = __half_as_ushort(h);
    lo = __half_as_ushort(l);
}
__device__ __forceinline__ uint64_t b_desc(uint32_t saddr, uint32_t lbo = kLBO) {
    return uint64_t((saddr >> 4) & 0x3fff) | (uint64_t((lbo >> 4) & 0x3fff) << 16) |
           (uint64_t((kSBO >> 4) & 0x3fff) << 32) | (uint64_t(1) << 46);  // version 1, SWIZZLE_NONE
}
__device__ __forceinline__ void mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc));
}
// same with the accumulator read as D 2^-11 (scale-input-d): merges the scaled correction chain into the main chain on the
// tensor core, D = A B + D / 2048
__device__ __forceinline__ void mma_f16_ts_scale11(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p, 11;\n\t}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc));
}
// compile-time loop i = 0 .. N-1: f(integral_constant<int, i>)
template <int I, int N, class F>
__device__ __forceinline__ void static_for(F &&f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(f);
    }
}
__device__ __forceinline__ void tmem_st8(uint32_t addr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(addr), "r"(v[0]),
                 "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]));
}
__device__ __forceinline__ void tmem_st4(uint32_t addr, const uint32_t *v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]));
}
__device__ __forceinline__ void tmem_st2(uint32_t addr, uint32_t v0, uint32_t v1) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(v0), "r"(v1));
}
__device__ __forceinline__ void tmem_st8p(uint32_t addr, const uint32_t *v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(addr), "r"(v[0]),
                 "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]));
}
__device__ __forceinline__ void tmem_st16p(uint32_t addr, const uint32_t *v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(addr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
        "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]));
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
// tcgen05.wait::ld that also "produces" the 16 registers of an earlier tcgen05.ld: with a load kept in flight across
// other work (h2_to_tmem prefetches the next chunk) the compiler must not move a use of those registers above the wait
__device__ __forceinline__ void tmem_wait_ld16(float *r) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+f"(r[0]), "+f"(r[1]), "+f"(r[2]), "+f"(r[3]), "+f"(r[4]), "+f"(r[5]), "+f"(r[6]), "+f"(r[7]), "+f"(r[8]),
                   "+f"(r[9]), "+f"(r[10]), "+f"(r[11]), "+f"(r[12]), "+f"(r[13]), "+f"(r[14]), "+f"(r[15]));
}
// try_wait with a suspend-time hint: the waiting warp is parked by the hardware until the phase completes (or the
// hint expires) instead of spinning.  A spinning warp competes for the issue slots of its scheduler; the MMA warp
// waits most of the time and made the producers that share its scheduler the slowest of the CTA (pipeline trace).
// Pipeline watchdog: a wait that has not completed after kWaitTimeoutNs of wall clock (%globaltimer) raises the
// launch's flag — a DEVICE word that every other wait of the grid polls (an L2 hit) so the launch drains quickly —
// and reports once to the mapped host word the library checks before the next launch and after its own syncs.
// Never expected (it would mean a protocol bug or a wedged producer); keeps a bug from hanging the GPU.
struct TcWatch {
    int *dev;    // __device__ flag of this GPU
    int *host;   // mapped pinned host word of this GPU
};
constexpr unsigned long long kWaitTimeoutNs = 2000000000ull;
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
template <bool PARK = true>
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, TcWatch watch) {
    uint32_t done = 0;
    int spins = 0;
    unsigned long long t_start = 0;
    while (!done) {
        if (PARK)
            asm volatile(
                "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                : "=r"(done)
                : "r"(bar), "r"(parity), "r"(100000)   // suspend-time hint in ns: an upper bound of one nap, not a latency
                : "memory");
        else
            asm volatile(
                "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                : "=r"(done)
                : "r"(bar), "r"(parity)
                : "memory");
        if (!done && watch.dev && (++spins & (PARK ? 7 : 1023)) == 0) {
            if (*reinterpret_cast<volatile int *>(watch.dev)) break;
            const unsigned long long now = globaltimer_ns();
            if (t_start == 0) t_start = now;
            if (now - t_start > kWaitTimeoutNs) {
                *reinterpret_cast<volatile int *>(watch.dev) = 1;
                if (watch.host) *reinterpret_cast<volatile int *>(watch.host) = 1;
                __threadfence_system();
                break;
            }
        }
    }
}
// non-blocking probe of a phase (true = complete)
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done)
                 : "r"(bar), "r"(parity)
                 : "memory");
    return done != 0;
}
// One thread of a converged warp.  The compiler recognises the elect.sync predicate as "a single thread": MMA
// operands then go to uniform registers once, instead of the per-instruction ELECT / R2UR / BRA.U.ANY loop it emits
// for an ordinary divergent branch such as `if (lane == 0)`.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// ---- producer math on fp32x2 pairs of hidden units (k, k+1) --------------------------------------
// staged pair table: sP[k/2] = (w1'[k], w1'[k+1], b1'[k], b1'[k+1]) with w1' = -log2(e) w1, b1' = -log2(e) b1 - 11
// (one LDS.128), so that 2^(w1' y + b1') = e^-a / 2048 and the reciprocal of  d' = e^-a / 2048 + 2^-11  is
// g = 2048 sigmoid(a): the hidden activations are produced PRE-SCALED by 2^11.  Then the fp16 remainder g - fp16(g)
// (< 1) needs no scaling of its own, it accumulates into D_main next to the hi piece, and the converter's
// x = D_main + D_corr / 2048 is 2048 (W2 h1 + b2), exactly the scaled value its own split wants.
// The producers are bound by issue slots and by the XU (MUFU) pipe (16 lanes per clock per SM; pipeline trace in
// profiles/): EX2 + RCP per unit is 200 MUFU per symbol.  Four reciprocals share ONE MUFU.RCP (Montgomery's trick:
// 1/a = b c d / (a b c d), arranged on packed pairs so that it costs 7 instructions per four values); the exponent
// is clamped so that the product of four d' stays finite (sigmoid floor 2^-41, far below fp32 resolution of the sums).
template <bool CLAMP = true>
__device__ __forceinline__ u64 denom2(uint32_t sP_addr, int pair, u64 yy) {
    u64 w, b;
    lds128(sP_addr + 16 * pair, w, b);
    float x0, x1;
    unpack2(fma2(yy, w, b), x0, x1);
    // CLAMP = false: the caller has checked |y| against the launch's safe bound (every exponent <= 29.5), same bits
    if (!CLAMP) return add2(pack2(ex2_approx(x0), ex2_approx(x1)), 0x3a0000003a000000ull);
#if MVN_TC_EXPERIMENT == 3   // bound-finding build (no overflow guard): no FMNMX clamps
    return add2(pack2(ex2_approx(x0), ex2_approx(x1)), 0x3a0000003a000000ull);
#elif MVN_TC_EXPERIMENT == 2   // bound-finding build (wrong results): no MUFU.EX2
    return add2(pack2(fminf(x0, 30.f) + 1.5f, fminf(x1, 30.f) + 1.5f), 0x3a0000003a000000ull);
#else
    return add2(pack2(ex2_approx(fminf(x0, 30.f)), ex2_approx(fminf(x1, 30.f))), 0x3a0000003a000000ull);  // + 2^-11
#endif
}
// (1/p.x, 1/p.y) and (1/q.x, 1/q.y) from one reciprocal
__device__ __forceinline__ void recip4(u64 p, u64 q, u64 &rp, u64 &rq) {
    float m0, m1;
    unpack2(mul2(p, q), m0, m1);                   // (p.x q.x, p.y q.y)
#if MVN_TC_RCP_SHARE == 2   // two reciprocals per four values: 5 instructions instead of 7, 2 MUFU instead of 1
    const u64 inv2 = pack2(rcp_approx(m0), rcp_approx(m1));
    rp = mul2(inv2, q);
    rq = mul2(inv2, p);
    return;
#endif
#if MVN_TC_EXPERIMENT == 1   // bound-finding build (wrong results): no MUFU.RCP
    const float r = m0 * m1 + 3.f;
#else
    const float r = rcp_approx(m0 * m1);
#endif
    const u64 inv = pack2(r * m1, r * m0);         // (1 / (p.x q.x), 1 / (p.y q.y))
    rp = mul2(inv, q);
    rq = mul2(inv, p);
}
// (g_k, g_{k+1}) -> fp16x2 words of the hi pieces and of the (unscaled) remainders (low half = even k)
// (c0 + m h.lo, c1 + m h.hi) for a packed fp16 pair h and an fp16 constant m, one mixed-precision FMA (FHFMA) per value:
// the fp16 piece goes back to fp32 inside the instruction (PTX ISA 8.6 fma.f32.f16), which saves the two unpack
// instructions of a split.
template <uint16_t M_BITS>
__device__ __forceinline__ void fma_f32_f16x2(uint32_t h, float c0, float c1, float &r0, float &r1) {
#if MVN_TC_MIXED_FMA
    asm("{\n\t.reg .f16 lo, hi, m;\n\tmov.b32 {lo, hi}, %2;\n\tmov.b16 m, %5;\n\t"
        "fma.rn.f32.f16 %0, lo, m, %3;\n\tfma.rn.f32.f16 %1, hi, m, %4;\n\t}\n"
        : "=f"(r0), "=f"(r1)
        : "r"(h), "f"(c0), "f"(c1), "n"(M_BITS));
#else
    // unpack (two HADD2.F32) + one packed FFMA2
    const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&h));
    const float m = __half2float(__ushort_as_half(M_BITS));
    unpack2(fma2(pack2(f.x, f.y), pack2(m, m), pack2(c0, c1)), r0, r1);
#endif
}
__device__ __forceinline__ void split2_pre(u64 g, uint32_t &whi, uint32_t &wlo) {
    float g0, g1;
    unpack2(g, g0, g1);
    const __half2 hi = __floats2half2_rn(g0, g1);
    whi = *reinterpret_cast<const uint32_t *>(&hi);
    float r0, r1;
    fma_f32_f16x2<0xbc00>(whi, g0, g1, r0, r1);   // g - hi, exact
    const __half2 lo = __floats2half2_rn(r0, r1);
    wlo = *reinterpret_cast<const uint32_t *>(&lo);
}
// X = 2048 a2 (converter warps, BEFORE the ReLU) -> fp16x2 words of the A operand of layer 3: hi = fp16(relu(a2)) rounded
// toward zero, lo = fp16((relu(a2) - hi) 2048).  The ReLU rides on the two conversions (cvt.relu): with hi rounded
// toward zero the remainder X - 2048 hi is >= 0 whenever X >= 0, and equals X < 0 otherwise, so clamping both
// conversions at zero is exactly relu on the value and costs no instruction.
__device__ __forceinline__ uint32_t cvt_f16x2_relu_rz(float hi_half, float lo_half) {
    uint32_t d;
    asm("cvt.rz.relu.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi_half), "f"(lo_half));
    return d;
}
__device__ __forceinline__ uint32_t cvt_f16x2_relu_rn(float hi_half, float lo_half) {
    uint32_t d;
    asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi_half), "f"(lo_half));
    return d;
}
__device__ __forceinline__ void split2_relu(u64 X, uint32_t &whi, uint32_t &wlo) {
    float h0, h1, x0, x1;
    unpack2(mul2(X, 0x3a0000003a000000ull), h0, h1);   // a2 = X / 2048
    whi = cvt_f16x2_relu_rz(h1, h0);                    // low half = even unit
    unpack2(X, x0, x1);
    float r0, r1;
    fma_f32_f16x2<0xe800>(whi, x0, x1, r0, r1);        // X - 2048 hi, exact
    wlo = cvt_f16x2_relu_rn(r1, r0);
}
// one double-pair (4 hidden units = 2 TMEM columns per piece) of this thread's frame: sigmoid -> split into registers
template <bool CLAMP = true>
__device__ __forceinline__ void compute_dpair(uint32_t sP_addr, int dp, u64 yy, uint32_t *vh, uint32_t *vl) {
    u64 g0, g1;
    recip4(denom2<CLAMP>(sP_addr, 2 * dp, yy), denom2<CLAMP>(sP_addr, 2 * dp + 1, yy), g0, g1);
    split2_pre(g0, vh[0], vl[0]);
    split2_pre(g1, vh[1], vl[1]);
}
// converter warps: h2 = relu(D_main + D_corr / 2048) of this thread's frame, split into fp16 hi / scaled lo,
// stored as the A operand of layer 3 (columns 0..31 of the slot's A_hi / A_lo regions, free after the MMAs of
// layer 2); column k2 = 50 is the constant 1 that multiplies the b3 row of B2.
template <int MODE>
__device__ __forceinline__ void h2_to_tmem(uint32_t slot_lane) {
    using LY = Lay<MODE>;
    constexpr bool DEC = (MODE == 1);
    if constexpr (DEC && MVN_TC_CONV_PREFETCH) {
        // one accumulator: the tcgen05.ld of chunk c0 + 1 is in flight while chunk c0 is converted (TMEM loads queue behind
        // the MUFU backlog of the producers that share the scheduler)
        float m[2][16];
        tmem_ld16(slot_lane + oDm, m[0]);
#pragma unroll
        for (int c0 = 0; c0 < kK2Steps; c0++) {
            tmem_wait_ld16(m[c0 & 1]);
            if (c0 + 1 < kK2Steps) tmem_ld16(slot_lane + oDm + 16 * (c0 + 1), m[(c0 + 1) & 1]);
            uint32_t vh[8], vl[8];
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const int k = 16 * c0 + 2 * q;
                if (k + 1 < kH2) {
                    split2_relu(pack2(m[c0 & 1][2 * q], m[c0 & 1][2 * q + 1]), vh[q], vl[q]);
                } else {  // k2 = 50: bias column (1.0 in the low half); beyond: zero padding
                    vh[q] = (k == kH2) ? 0x00003c00u : 0u;
                    vl[q] = 0u;
                }
            }
            tmem_st8(slot_lane + LY::hh(c0), vh);
            tmem_st8(slot_lane + LY::hl(c0), vl);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;");
        return;
    }
#pragma unroll
    for (int c0 = 0; c0 < kK2Steps; c0++) {   // 16 hidden units = 8 columns per step
        float m[16], c[16];
        tmem_ld16(slot_lane + oDm + 16 * c0, m);
        if (!DEC) tmem_ld16(slot_lane + oDc + 16 * c0, c);   // (DEC: the accumulator already holds 2048 (W2 h1 + b2))
        asm volatile("tcgen05.wait::ld.sync.aligned;");
        uint32_t vh[8], vl[8];
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const int k = 16 * c0 + 2 * q;
            if (k + 1 < kH2) {
                // D_main + D_corr / 2048 = 2048 (W2 h1 + b2), both units of the pair in one packed FMA; ReLU and the
                // hi / lo split in split2_relu
                const u64 X = DEC ? pack2(m[2 * q], m[2 * q + 1])
                                  : fma2(pack2(c[2 * q], c[2 * q + 1]), 0x3a0000003a000000ull, pack2(m[2 * q], m[2 * q + 1]));
                split2_relu(X, vh[q], vl[q]);
            } else {  // k2 = 50: bias column (1.0 in the low half); beyond: zero padding
                vh[q] = (k == kH2) ? 0x00003c00u : 0u;
                vl[q] = 0u;
            }
        }
        tmem_st8(slot_lane + LY::hh(c0), vh);
        tmem_st8(slot_lane + LY::hl(c0), vl);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;");
}
}  // namespace tc

// Optional pipeline trace (debug builds only, -DMVN_TC_TRACE): CTA 0 records clock64() at the key points of the
// first 64 stages, 32 event slots per stage, into trace[stage*32 + event]; slots 16+w: A stored by producer warp w (12 warps), 28+w: start of producer warps 0..3.
#ifdef MVN_TC_TRACE
#define TC_TRACE(ev, cond)                                                                               \
    do {                                                                                                 \
        if (trace && blockIdx.x == 0 && (cond) && n < 64) {                                                 \
            trace[n * 32 + (ev)] = clock64();                                                            \
            if ((ev) == 0) {  /* wall clock next to the cycle counter: effective SM clock under this load */ \
                unsigned long long gt;                                                                   \
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));                                    \
                trace[n * 32 + 14] = (long long)gt;                                                      \
            }                                                                                            \
        }                                                                                                \
    } while (0)
#else
#define TC_TRACE(ev, cond) do {} while (0)
#endif

template <int L, bool MLSE>
__global__ void __maxnreg__(tc::Roles<L>::kMaxReg) vnet_decode_tc_kernel(VnetParams p, tc::TcWatch timeout_flag, long long *trace) {
    static_assert(L <= 8, "tcgen05 variant: memory_length 1..8");
    // L <= 5: priors_main | priors_corr side by side inside the slot's 64 accumulator columns (one N = 2 N2 MMA per k-step).
    // L == 6: 64 priors: the correction chain is folded into the main chain (scale-input-d), as for L == 7.
    // L == 7: 128 priors fill both regions, so the correction chain is computed first and folded into the main chain by
    //         the tensor core itself (scale-input-d: D = A B + D 2^-11): ONE 128-column accumulator.
    // L == 8: 256 priors go through the same 128 columns in TWO layer-3 passes per stage (states 0..127, then 128..255:
    //         the stage loop consumes the source states in that order anyway); the second pass is issued by a consumer
    //         thread once all consumer warps have read the first half.  W3's fp16 pieces (64 KB) and the 256-state path
    //         metrics of the 128 frames (128 KB) fill the shared memory, so this instance stages no tiles at all: the
    //         producers read their sample straight from global memory (one L1-resident line per frame and 32 stages),
    //         the consumers their targets.
    // TMEM layout of a slot (tc::Lay).  Measured: mode 3 at 128 states 10.9 vs 11.8 G sym/s (mode 0), at 256 states 5.64 vs 5.45
    constexpr int LAYM = (L <= 6) ? MVN_TC_LAYOUT : (L == 7) ? MVN_TC_LAYOUT_L7 : MVN_TC_LAYOUT_L8;
    constexpr bool SINGLE_A = (LAYM == 3);       // one A buffer: the producers wait for the layer-2 MMAs of the previous STAGE
    constexpr bool DEC = (LAYM != 0);            // h2 outside the A columns: producers wait for the layer-2 MMAs, not the consumers
    constexpr bool L2MERGED = (LAYM == 1);       // one layer-2 accumulator (scale-input-d)
    using LY = tc::Lay<LAYM>;
    constexpr bool MERGED = (L >= 6);            // layer-3 correction chain folded into the main accumulator (64..128 columns)
    constexpr int NPASS = (L == 8) ? 2 : 1;      // layer-3 passes per stage
    constexpr int kQ = 4;                        // active TMEM lane quadrants = 32-frame warp tiles per CTA tile
    constexpr bool DIRECT = (L == 8);            // no staged tiles (see above)
    using D = TrellisDims<L>;
    using RL = tc::Roles<L>;
    constexpr int PW = RL::PW, KW = RL::KW, PARTS = RL::PARTS;   // producer / consumer warps, producer warps per quadrant
    constexpr bool DUAL = RL::DUAL;
    constexpr int S = D::S, C = D::C, NCH = D::NCH, NW = PW + tc::kConvWarps + KW;
    constexpr int N2 = tc::n2_of(S), kB2Bytes = tc::b2_bytes(S);
    constexpr int N3 = N2 / NPASS;               // output columns of one layer-3 pass
    constexpr uint32_t kLBO2 = (2 * N2 / 8) * 128;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint32_t tmem_base_s;
    __shared__ float cons_xchg[tc::Roles<L>::DUAL ? 2 * 2 * 32 * 4 : 1];   // DUAL: [stage parity][even | odd][frame] minima of the second consumer warp
    __shared__ float y_safe_s;   // |y| up to which no sigmoid exponent exceeds 29.5 (the producers' clamp-free path)
    __shared__ __align__(8) uint64_t d_full[2], slot_free[2], a_full[4], d2_full[2];
    // a_full: one barrier per issue warp (stage n -> a_full[n mod kAF], phase (n / kAF) & 1), so that every waiter sees every
    // phase of its barrier (two warps alternating on one barrier could not tell phase u-1 from u+1)
    constexpr uint32_t kAF = tc::kMmaWarps > 2 ? 4 : 2;
    constexpr int NT_TILES = DIRECT ? 0 : (PARTS + 1) * kQ;             // producers and consumers stage tiles
    uint8_t *sB = smem_raw;                                                      // W2 pieces, hi rows | lo rows
    float *tiles = reinterpret_cast<float *>(smem_raw + tc::kBBytes);            // one 32x32 tile per such warp
    float *sP = tiles + NT_TILES * kTileFloats;                                  // [56][4] pair table for the packed sigmoid
    uint8_t *sB2 = reinterpret_cast<uint8_t *>(sP + 4 * (tc::kK / 2));           // W3 (+ b3 column) pieces: hi | lo
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, quad = warp & 3;
    const bool producer = warp < PW;
    const bool converter = !producer && warp < PW + tc::kConvWarps;
    const bool mma_warp = warp >= NW;
    // tiles: one per producer warp (y) and one per consumer warp (targets); converters and the MMA warp use none
    const bool active = quad < kQ;
    float *tile = tiles + (DIRECT ? 0 : (producer ? (warp >> 2) * kQ + (active ? quad : 0) : PARTS * kQ + (active ? quad : 0)) * kTileFloats);

    // ---- W2 (and b2 as column k=100) -> fp16 hi / scaled-lo pieces in the canonical K-major layout
    for (int idx = tid; idx < tc::kN * tc::kK; idx += RL::kThreads) {
        const int n = idx / tc::kK, k = idx % tc::kK;
        float w = 0.f;
        if (n < kH2) w = k < kH1 ? p.w.w2[n * kH1 + k] : (k == kH1 ? p.w.b2[n] : 0.f);
        uint16_t hi, lo;
        tc::split_f16(w, hi, lo);
        const int off = (k / 8) * (2 * tc::kN / 8) * 128 + (n / 8) * 128 + (n % 8) * 16 + (k % 8) * 2;
        *reinterpret_cast<uint16_t *>(sB + off) = hi;
        *reinterpret_cast<uint16_t *>(sB + off + (tc::kN / 8) * 128) = lo;
    }
    for (int i = tid; i < tc::kK / 2; i += RL::kThreads) {
        const float kNegLog2e = -1.4426950408889634f;
        const int k = 2 * i;
        sP[4 * i + 0] = k < kH1 ? p.w.w1[k] * kNegLog2e : 0.f;
        sP[4 * i + 1] = k + 1 < kH1 ? p.w.w1[k + 1] * kNegLog2e : 0.f;
        sP[4 * i + 2] = k < kH1 ? fmaf(p.w.b1[k], kNegLog2e, -11.f) : 0.f;      // 2^(...) = e^-a / 2048
        sP[4 * i + 3] = k + 1 < kH1 ? fmaf(p.w.b1[k + 1], kNegLog2e, -11.f) : 0.f;
    }
    // ---- W3 (and b3 as column k2=50) -> fp16 pieces, canonical K-major layout with N2 rows (states)
    for (int idx = tid; idx < N2 * tc::kK2; idx += RL::kThreads) {
        const int n2 = idx / tc::kK2, k = idx % tc::kK2;
        float w = 0.f;
        if (n2 < S) w = k < kH2 ? p.w.w3[n2 * kH2 + k] : (k == kH2 ? p.w.b3[n2] : 0.f);
        uint16_t hi, lo;
        tc::split_f16(w, hi, lo);
        const int off = (k / 8) * (2 * N2 / 8) * 128 + (n2 / 8) * 128 + (n2 % 8) * 16 + (k % 8) * 2;
        *reinterpret_cast<uint16_t *>(sB2 + off) = hi;
        *reinterpret_cast<uint16_t *>(sB2 + off + (N2 / 8) * 128) = lo;
    }
    if (tid == 32) {
        float ys = 3.0e38f;
        for (int k = 0; k < kH1; k++) {
            const float kNegLog2e = -1.4426950408889634f;
            const float w = fabsf(p.w.w1[k] * kNegLog2e), b = fmaf(p.w.b1[k], kNegLog2e, -11.f);
            // |w| Y + b <= 29.5;  NaN weights fail every comparison below and leave the clamped path
            const float lim = (b <= 29.5f) ? (w > 0.f ? (29.5f - b) / w : 3.0e38f) : -1.f;
            ys = (lim < ys) ? lim : ((lim >= ys) ? ys : -1.f);
        }
        y_safe_s = ys * 0.999f;
    }
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < 2; s++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&d_full[s])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&d2_full[s])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(&slot_free[s])), "n"(KW));
        }
#pragma unroll
        for (int s = 0; s < 4; s++)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(&a_full[s])), "n"(PW));
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
    const uint32_t lane_base = uint32_t(quad * 32) << 16;
    if (warp < 4) {
        // constant columns of the layer-2 A operand, both slots: k = 100 multiplies the b2 row of B (2048 = fp16 0x6800 in the
        // low half of column 50, the activations are pre-scaled by 2^11), k = 101..111 is zero padding.  Ordered before the
        // first MMA by this warp's own first a_full arrival (wait::st + fence below).
        const uint32_t z[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int sl = 0; sl < 2; sl++) {
            const uint32_t base = tmem + LY::a_col(sl) + lane_base;
            tc::tmem_st2(base + 50, 0x00006800u, 0u);
            tc::tmem_st4(base + 52, z);
            tc::tmem_st2(base + tc::kACols + 50, 0u, 0u);
            tc::tmem_st4(base + tc::kACols + 52, z);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;");
    }
    const uint32_t sB_addr = smem_addr(sB), sB2_addr = smem_addr(sB2);
    // D=F32, A=B=F16, both K-major, N>>3 at bit 17, M>>4 at bit 24; "w" = hi and lo pieces of B side by side
    constexpr uint32_t idesc2 = (1u << 4) | (uint32_t(N3 >> 3) << 17) | (uint32_t(tc::kM >> 4) << 24);
    constexpr uint32_t idesc2w = (1u << 4) | (uint32_t(2 * N2 >> 3) << 17) | (uint32_t(tc::kM >> 4) << 24);
    constexpr uint32_t idesc = (1u << 4) | (uint32_t(tc::kN >> 3) << 17) | (uint32_t(tc::kM >> 4) << 24);
    constexpr uint32_t idescw = (1u << 4) | (uint32_t(2 * tc::kN >> 3) << 17) | (uint32_t(tc::kM >> 4) << 24);
    const uint32_t sP_addr = smem_addr(sP);
    const bool vec_in = is_vec_ok(p.y, p.T, p.T);
    const bool vec_out = p.out_format == MVN_OUT_F32 && is_vec_ok(p.decoded, p.T, p.T);
    const bool vec_tgt = p.target && is_vec_ok(p.target, p.target_T, p.target_T);
    const int n_words = (p.T + 31) / 32;
    const int64_t n_cta_tiles = (p.n_warp_tiles + kQ - 1) / kQ;  // CTA tile = kQ warp tiles of 32 frames
    uint32_t n = 0;                                        // running stage counter: slot = n & 1, use = n >> 1
    // layer 3 with the correction chain folded in by scale-input-d (MERGED): pass `half` covers states [N3 half, N3 half + N3)
    auto issue_layer3_merged = [&](uint32_t ts, int half) {
        constexpr uint32_t kLoRows = (N2 / 8) * 128;       // byte offset of the W3_lo rows inside a k-chunk
        const uint32_t rows = uint32_t(half) * (N3 / 8) * 128;
#pragma unroll
        for (int j = 0; j < tc::kK2Steps; j++)   // corr = h2_hi W3_lo
            tc::mma_f16_ts(ts + tc::oDm, ts + LY::hh(j), tc::b_desc(sB2_addr + uint32_t(2 * j) * kLBO2 + kLoRows + rows, kLBO2),
                           idesc2, j > 0);
#pragma unroll
        for (int j = 0; j < tc::kK2Steps; j++)   // corr += h2_lo W3_hi
            tc::mma_f16_ts(ts + tc::oDm, ts + LY::hl(j), tc::b_desc(sB2_addr + uint32_t(2 * j) * kLBO2 + rows, kLBO2), idesc2, 1);
        // priors = h2_hi W3_hi + corr / 2048: the first MMA of the main chain reads D scaled by 2^-11
        tc::mma_f16_ts_scale11(ts + tc::oDm, ts + LY::hh(0), tc::b_desc(sB2_addr + rows, kLBO2), idesc2);
#pragma unroll
        for (int j = 1; j < tc::kK2Steps; j++)
            tc::mma_f16_ts(ts + tc::oDm, ts + LY::hh(j), tc::b_desc(sB2_addr + uint32_t(2 * j) * kLBO2 + rows, kLBO2), idesc2, 1);
    };

    if (producer) {
        int rot = 0;  // which of the quadrant's producer warps takes the 25th double-pair in this stage
        const float y_safe = y_safe_s;
        // MVN_TC_LATE_PUBLISH: the tcgen05.st of stage n are published (wait::st, fence, a_full arrival) after the first
        // double-pair of stage n+1 has been computed into spare registers, so their latency overlaps arithmetic
        bool pending = false;
        uint32_t pending_n = 0;
        auto publish = [&]() {
            asm volatile("tcgen05.wait::st.sync.aligned;");
            asm volatile("tcgen05.fence::before_thread_sync;");
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(smem_addr(&a_full[pending_n % kAF]));  // one arrival per producer warp
            __syncwarp();
            pending = false;
        };
        for (int64_t ct = blockIdx.x; ct < n_cta_tiles; ct += gridDim.x) {
            const int64_t row0 = (ct * kQ + quad) * 32;
            for (int t0 = 0; t0 < p.T; t0 += 32) {
                const int t_end = min(32, p.n_stages - t0);
                if (t_end <= 0) continue;
                if (!active) {   // (L == 8) this quadrant carries no frames: keep a_full's arrival count in step
                    for (int tt = 0; tt < t_end; tt++, n++) {
                        tc::mbar_wait<true>(smem_addr(&slot_free[n & 1]), ((n >> 1) & 1) ^ 1, timeout_flag);
                        __syncwarp();
                        if (lane == 0) tc::mbar_arrive(smem_addr(&a_full[n % kAF]));
                        __syncwarp();
                    }
                    continue;
                }
                if (pending) publish();   // not behind the global loads of the next tile
                if constexpr (!DIRECT) warp_load_tile(p.y, p.B, p.T, p.T, row0, t0, tile, lane, vec_in);
                const float *yrow = p.y + (row0 + lane < p.B ? row0 + lane : 0) * int64_t(p.T) + t0;
#pragma unroll 1
                for (int tt = 0; tt < t_end; tt++, n++) {
                    const uint32_t slot = n & 1, use = n >> 1;
                    const uint32_t a_lane = tmem + LY::a_col(slot) + lane_base;
                    const float yv = DIRECT ? __ldg(yrow + tt) : tile[lane * kTileLd + tt];
                    const u64 yy = pack2(yv, yv);
                    // Compute this stage's pieces into registers BEFORE waiting for the slot: the sigmoid/split work
                    // then overlaps the MMAs and the consumer of the stage that still owns the slot.
                    TC_TRACE(0, tid == 0);
                    // Hidden units per producer warp (kProdParts warps per TMEM lane quadrant): the 50 pairs of units are 25
                    // double-pairs (four sigmoids share one reciprocal); warp `part` owns double-pairs [DP part, DP part + DP)
                    // = columns [2 DP part, ...) of A_hi / A_lo, and the 25th (columns 48, 49) rotates among the quadrant's
                    // warps from stage to stage (a fixed owner made that warp the pace-setter of the whole CTA: pipeline
                    // trace).  The bias column (50) and the zero padding (51..55) are written once per launch.
                    // With two producer warps per quadrant (DUAL) a warp's 12 double-pairs go through the registers in two halves.
                    constexpr int DP = 24 / PARTS, NHALF = (PARTS == 2) ? 2 : 1, DPH = DP / NHALF, NC = 2 * DPH;   // per half: 8 double-pairs = 16 columns, or 6 = 12
                    const int part = warp >> 2;
                    const bool extra = rot == part;
                    uint32_t vh[NC + 2], vl[NC + 2];
                    // The slot's barrier is probed in the MIDDLE of the stage's arithmetic: at its end the probe would queue
                    // behind this scheduler's MUFU backlog (pipeline trace: 270 cycles for a wait that succeeds at once).
                    // (SINGLE_A: the previous stage's layer-2 MMAs, i.e. the OTHER slot's d_full; a fresh barrier passes a wait
                    //  for the phase before its first, which covers n = 0)
                    const uint32_t slot_bar = smem_addr(SINGLE_A ? &d_full[(n - 1) & 1] : DEC ? &d_full[slot] : &slot_free[slot]);
                    const uint32_t slot_par = SINGLE_A ? (((n - 1) >> 1) & 1) : ((use & 1) ^ 1);
                    bool slot_ready = false;
                    // (the overflow clamps of the 100 exponents are skipped when every lane's |y| is below the launch's bound)
                    auto compute_half = [&](auto clamp_c, auto hf_c) {
                    constexpr bool CL = decltype(clamp_c)::value;
                    constexpr int HF = decltype(hf_c)::value;
#pragma unroll
                    for (int i = 0; i < DPH; i++) {
#if MVN_TC_EXPERIMENT == 4   // bound-finding build (wrong results): half of the sigmoids
                        if (i & 1) { vh[2 * i] = vh[2 * i - 2]; vh[2 * i + 1] = vh[2 * i - 1]; vl[2 * i] = vl[2 * i - 2]; vl[2 * i + 1] = vl[2 * i - 1]; continue; }
#endif
                        // (fetching the next stage's sample here as well measured the same and cost a spill)
                        if (MVN_TC_EARLY_PROBE && HF == 0 && i == DPH / 2) slot_ready = tc::mbar_test(slot_bar, slot_par);
                        if (MVN_TC_LATE_PUBLISH && HF == 0 && i == 0) {
                            uint32_t th[2], tl[2];   // vh / vl may still be read by the stores of the previous stage
                            tc::compute_dpair<CL>(sP_addr, DP * part, yy, th, tl);
                            if (pending) publish();
                            vh[0] = th[0], vh[1] = th[1], vl[0] = tl[0], vl[1] = tl[1];
                            continue;
                        }
                        tc::compute_dpair<CL>(sP_addr, DP * part + DPH * HF + i, yy, vh + 2 * i, vl + 2 * i);
                    }
                    if (HF == NHALF - 1 && extra) tc::compute_dpair<CL>(sP_addr, 24, yy, vh + NC, vl + NC);
                    };
                    auto store_half = [&](int hf) {
                        const uint32_t ah = a_lane + 2 * DP * part + NC * hf, al = ah + tc::kACols;
                        if constexpr (NC == 16) {
                            tc::tmem_st16p(ah, vh);
                            tc::tmem_st16p(al, vl);
                        } else {
                            tc::tmem_st8p(ah, vh);
                            tc::tmem_st8p(al, vl);
                            tc::tmem_st4(ah + 8, vh + 8);
                            tc::tmem_st4(al + 8, vl + 8);
                        }
                        if (hf == NHALF - 1 && extra) {
                            tc::tmem_st2(a_lane + 48, vh[NC], vh[NC + 1]);
                            tc::tmem_st2(a_lane + tc::kACols + 48, vl[NC], vl[NC + 1]);
                        }
                    };
                    const bool clamp_free = MVN_TC_CLAMP_FREE && __all_sync(0xffffffffu, fabsf(yv) <= y_safe);
                    if (clamp_free) compute_half(std::false_type{}, std::integral_constant<int, 0>{});
                    else compute_half(std::true_type{}, std::integral_constant<int, 0>{});
                    TC_TRACE(1, tid == 0);
                    // (128 / 256 states are bound by the consumers: there the producers park instead of polling, which
                    //  would take issue slots from the consumer warp on their scheduler)
                    // DEC: the A columns are free once the layer-2 MMAs of the slot's previous use have read them (d_full);
                    // round-1 layout: once the consumers have released the slot (h2 lives in the A columns)
                    if (!slot_ready) tc::mbar_wait<(MVN_PROD_PARK || L >= MVN_TC_PARK_MINL)>(slot_bar, slot_par, timeout_flag);
                    TC_TRACE(28, tid == 0);
                    asm volatile("tcgen05.fence::after_thread_sync;");
                    TC_TRACE(2, tid == 0);
                    store_half(0);
                    if constexpr (NHALF == 2) {
                        asm volatile("tcgen05.wait::st.sync.aligned;");   // the registers of the first half are free again
                        if (clamp_free) compute_half(std::false_type{}, std::integral_constant<int, NHALF - 1>{});
                        else compute_half(std::true_type{}, std::integral_constant<int, NHALF - 1>{});
                        store_half(1);
                    }
                    rot = rot == PARTS - 1 ? 0 : rot + 1;
                    TC_TRACE(29, tid == 0);
                    if constexpr (MVN_TC_LATE_PUBLISH) {
                        pending = true;
                        pending_n = n;
                        TC_TRACE(3, tid == 0);            // (stores issued; published during the next stage)
                        TC_TRACE(16 + warp, lane == 0);
                    } else {
                        asm volatile("tcgen05.wait::st.sync.aligned;");
                        TC_TRACE(30, tid == 0);
                        asm volatile("tcgen05.fence::before_thread_sync;");
                        TC_TRACE(3, tid == 0);
                        TC_TRACE(16 + warp, lane == 0);
                        __syncwarp();
                        if (lane == 0) tc::mbar_arrive(smem_addr(&a_full[n % kAF]));  // one arrival per producer warp
                        __syncwarp();
                        TC_TRACE(31, tid == 0);
                    }
                }
                __syncwarp();
            }
        }
        if (pending) publish();
    } else if (mma_warp) {
        // one warp does nothing but issue: per stage 7 + 7 MMAs, then tcgen05.commit -> d_full[slot]
        for (int64_t ct = blockIdx.x; ct < n_cta_tiles; ct += gridDim.x) {
            for (int t0 = 0; t0 < p.T; t0 += 32) {
                const int t_end = min(32, p.n_stages - t0);
#pragma unroll 1
                for (int tt = 0; tt < t_end; tt++, n++) {
                    if (tc::kMmaWarps > 1 && int(n % tc::kMmaWarps) != warp - NW) continue;   // another issue warp's stage
                    const uint32_t slot = n & 1, use = n >> 1;
                    tc::mbar_wait(smem_addr(&a_full[n % kAF]), (n / kAF) & 1, timeout_flag);
                    asm volatile("tcgen05.fence::after_thread_sync;");
                    TC_TRACE(4, lane == 0);
                    if constexpr (DEC) {   // the accumulator columns also carry the priors of the slot's previous use: wait for its consumers
                        tc::mbar_wait(smem_addr(&slot_free[slot]), (use & 1) ^ 1, timeout_flag);
                        asm volatile("tcgen05.fence::after_thread_sync;");
                    }
                    TC_TRACE(15, lane == 0);
                    if (tc::elect_one()) {
                        const uint32_t ts = tmem + LY::d_col(slot), ah = tmem + LY::a_col(slot), al = ah + tc::kACols;
                        if constexpr (!L2MERGED) {
#pragma unroll
                            for (int j = 0; j < tc::kKSteps; j++)   // D_main | D_corr = A_hi [B_hi | B_lo]
                                tc::mma_f16_ts(ts + tc::oDm, ah + j * 8, tc::b_desc(sB_addr + uint32_t(2 * j) * tc::kLBO), idescw,
                                               j > 0);
#pragma unroll
                            for (int j = 0; j < tc::kKSteps; j++)   // D_main += A_lo B_hi (the remainder of the pre-scaled A is unscaled)
                                tc::mma_f16_ts(ts + tc::oDm, al + j * 8, tc::b_desc(sB_addr + uint32_t(2 * j) * tc::kLBO), idesc, 1);
                        } else {
                            constexpr uint32_t kLoRowsB = (tc::kN / 8) * 128;   // W2_lo rows inside a k-chunk of the stacked B
#pragma unroll
                            for (int j = 0; j < tc::kKSteps; j++)   // D = A_hi B_lo (the 2048-scaled correction)
                                tc::mma_f16_ts(ts + tc::oDm, ah + j * 8, tc::b_desc(sB_addr + uint32_t(2 * j) * tc::kLBO + kLoRowsB), idesc,
                                               j > 0);
                            // D = A_hi B_hi + D / 2048: scale-input-d on the first MMA of the main chain
                            tc::mma_f16_ts_scale11(ts + tc::oDm, ah, tc::b_desc(sB_addr), idesc);
#pragma unroll
                            for (int j = 1; j < tc::kKSteps; j++)
                                tc::mma_f16_ts(ts + tc::oDm, ah + j * 8, tc::b_desc(sB_addr + uint32_t(2 * j) * tc::kLBO), idesc, 1);
#pragma unroll
                            for (int j = 0; j < tc::kKSteps; j++)   // D += A_lo B_hi (the remainder of the pre-scaled A is unscaled)
                                tc::mma_f16_ts(ts + tc::oDm, al + j * 8, tc::b_desc(sB_addr + uint32_t(2 * j) * tc::kLBO), idesc, 1);
                        }
                        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                            smem_addr(&d_full[slot])));
                    }
                    TC_TRACE(5, lane == 0);
                    __syncwarp();
                }
            }
        }
    } else if (converter) {
        // h2 converters: layer-2 result of stage n -> ReLU -> hi/lo split -> A operand of layer 3 in the slot's A columns, then
        // the layer-3 MMAs.  They work on stage n+1 (other slot) while layer 3 of stage n runs and the consumers finish it.
        for (int64_t ct = blockIdx.x; ct < n_cta_tiles; ct += gridDim.x) {
            for (int t0 = 0; t0 < p.T; t0 += 32) {
                const int t_end = min(32, p.n_stages - t0);
#pragma unroll 1
                for (int tt = 0; tt < t_end; tt++, n++) {
                    const uint32_t slot = n & 1, use = n >> 1;
                    const uint32_t ts = tmem + LY::d_col(slot), slot_lane = ts + lane_base;
                    TC_TRACE(6, warp == PW && lane == 0);
                    tc::mbar_wait<MVN_CONS_PARK>(smem_addr(&d_full[slot]), use & 1, timeout_flag);
                    asm volatile("tcgen05.fence::after_thread_sync;");
                    TC_TRACE(7, warp == PW && lane == 0);
                    if (active) tc::h2_to_tmem<LAYM>(slot_lane);
                    asm volatile("tcgen05.fence::before_thread_sync;");
                    TC_TRACE(8, warp == PW && lane == 0);
                    asm volatile("bar.sync 2, %0;" ::"n"(32 * tc::kConvWarps));
                    TC_TRACE(9, warp == PW && lane == 0);
                    if (warp == PW + (tc::kMmaWarps > 1 ? int((n + 2) & 3) : 1) && tc::elect_one()) {  // one thread issues layer 3
                        asm volatile("tcgen05.fence::after_thread_sync;");
                        if constexpr (!MERGED) {   // 4 + 4 MMAs
#pragma unroll
                            for (int j = 0; j < tc::kK2Steps; j++)   // priors_main | priors_corr = h2_hi [W3_hi | W3_lo]
                                tc::mma_f16_ts(ts + tc::oDm, ts + LY::hh(j), tc::b_desc(sB2_addr + uint32_t(2 * j) * kLBO2, kLBO2),
                                               idesc2w, j > 0);
#pragma unroll
                            for (int j = 0; j < tc::kK2Steps; j++)   // priors_corr += h2_lo W3_hi
                                tc::mma_f16_ts(ts + tc::oDm + N2, ts + LY::hl(j), tc::b_desc(sB2_addr + uint32_t(2 * j) * kLBO2, kLBO2),
                                               idesc2, 1);
                        } else {                   // 4 + 4 + 4 MMAs of N = 128 into one accumulator (pass 0: states 0..127)
                            issue_layer3_merged(ts, 0);
                        }
                        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                            smem_addr(&d2_full[slot])));
                    }
                    TC_TRACE(10, warp == PW && lane == 0);
                    __syncwarp();
                }
            }
        }
    } else {
        // consumers: priors of stage n -> ACS on the frame's private path metrics -> decision bit, outputs, BER.
        // MLSE: no running decision; the S/2 survivor bits of every stage go to this warp's shared-memory masks and the
        // frame is traced back in the kernel once its last stage is through (the producers are already two stages into the
        // next tile by then, so the traceback overlaps their work).
        typename std::conditional<(L <= 5), RegTrellis<L>, SmemTrellisFixed<L, 32 * kQ>>::type tr;
        float *after_w3 = reinterpret_cast<float *>(sB2 + kB2Bytes);
        if constexpr (L > 5) {  // path metrics of the frames of the tile: [2][H][32 kQ] floats behind the W3 pieces
            tr.init(after_w3, 32 * kQ, (active ? quad : 0) * 32 + lane);
            after_w3 += SmemTrellis<L>::bytes(32 * kQ) / sizeof(float);
        }
        SurvStore<L> surv;
        if constexpr (MLSE) surv.init(reinterpret_cast<uint32_t *>(after_w3) + size_t(quad) * p.surv_words * 32, lane);
        ErrAcc acc;
        constexpr int kConsFirst = PW + tc::kConvWarps;
        // DUAL: warps kConsFirst + q (lead) and kConsFirst + 4 + q serve quadrant q.  The lead decides, emits and counts; both
        // run the add-compare-select of their half of every layer-3 pass's chunks on the shared metrics, meet at a named
        // barrier of their own once per stage, and the lead folds the other warp's even / odd minima into its own.
        const bool lead = !DUAL || warp < kConsFirst + 4;
        constexpr int NCHp = NCH / NPASS;   // chunks per layer-3 pass
        auto pair_sync = [&]() {
            if constexpr (DUAL) asm volatile("bar.sync %0, 64;" ::"r"(4 + quad) : "memory");
        };
        for (int64_t ct = blockIdx.x; ct < n_cta_tiles; ct += gridDim.x) {
            const int64_t row0 = (ct * kQ + quad) * 32;
            const int64_t b = row0 + lane;
            if constexpr (DUAL) {
                pair_sync();                 // the other warp is done with the previous tile's metrics
                if (lead) tr.reset();
                else tr.ne = tr.no = 3.0e38f;
                pair_sync();
            } else {
                if (active) tr.reset();
            }
            unsigned frame_bit_errs = 0;
            auto emit = [&](int t0, uint32_t bits) {
                if (p.decoded) {
                    if (p.out_format == MVN_OUT_F32)
                        warp_store_bits_f32(static_cast<float *>(p.decoded), p.B, p.T, p.T, row0, t0, bits, lane, vec_out);
                    else if (b < p.B)
                        static_cast<uint32_t *>(p.decoded)[b * n_words + t0 / 32] = bits;
                }
                if (p.target && t0 < p.target_T) {
                    if constexpr (DIRECT) {
                        if (b < p.B) frame_bit_errs += row_bit_errors_global(p.target + b * p.target_T + t0, bits, p.target_T - t0);
                    } else {
                        warp_load_tile(p.target, p.B, p.target_T, p.target_T, row0, t0, tile, lane, vec_tgt);
                        frame_bit_errs += tile_bit_errors(tile + lane * kTileLd, bits, p.target_T - t0);
                        __syncwarp();
                    }
                }
            };
            for (int t0 = 0; t0 < p.T; t0 += 32) {
                uint32_t bits = 0;
                const int t_end = min(32, p.n_stages - t0);
#pragma unroll 1
                for (int tt = 0; tt < t_end; tt++, n++) {
                    const uint32_t slot = n & 1, use = n >> 1;
                    const uint32_t ts = tmem + LY::d_col(slot), slot_lane = ts + lane_base;
                    if constexpr (!MLSE) {
                        if (active && lead) bits |= tr.decide() << tt;   // metrics entering this stage; overlaps the wait
                    }
                    tc::mbar_wait<MVN_CONS_PARK>(smem_addr(&d2_full[slot]), (NPASS * use) & 1, timeout_flag);
                    asm volatile("tcgen05.fence::after_thread_sync;");
                    TC_TRACE(11, warp == kConsFirst && lane == 0);
                    float *dst = (p.priors_out && b < p.B) ? p.priors_out + (b * p.T + t0 + tt) * S : nullptr;
                    uint32_t sv = 0;
                    // 16 source states per chunk: priors = D_main + D_corr / 2048, cost = -prior (vnet_detector.py:57)
                    auto chunk = [&](auto cc) {
                        constexpr int c = decltype(cc)::value;
                        constexpr bool last = (c == NCH - 1) || (DUAL && c == NCH - NCHp / 2 - 1);   // this warp's last tcgen05.ld of the stage
                        constexpr int col = 16 * (c % (NCH / NPASS));   // a pass refills the same D columns
                        float pm_[16], pc_[16];
                        if (active) {
                            tc::tmem_ld16(slot_lane + tc::oDm + col, pm_);
                            if constexpr (!MERGED) tc::tmem_ld16(slot_lane + tc::oDm + N2 + col, pc_);
                            asm volatile("tcgen05.wait::ld.sync.aligned;");
                        }
                        if (last) {
                            asm volatile("tcgen05.fence::before_thread_sync;");
                            TC_TRACE(12, warp == kConsFirst && lane == 0);
                            __syncwarp();
                            if (lane == 0) tc::mbar_arrive(smem_addr(&slot_free[slot]));  // the slot's A and D columns may be refilled
                        }
                        if (!active) return;
                        float pr[C], cost[C];
#pragma unroll
                        for (int i = 0; i < C; i++) {
                            pr[i] = MERGED ? pm_[i] : fmaf(pc_[i], tc::kInvScale, pm_[i]);
                            cost[i] = -pr[i];
                        }
                        uint32_t s8;
                        s8 = tr.template step_chunk<c>(cost);
                        if constexpr (MLSE) sv |= s8 << ((c * (C / 2)) & 31);
                        if (dst) {
#pragma unroll
                            for (int i = 0; i < C; i++) dst[c * C + i] = pr[i];
                        }
                    };
                    // MERGED (>= 64 states): two chunks per tcgen05.wait::ld — half as many TMEM round trips through the MIO queue
                    auto chunk2 = [&](auto cc) {
                        constexpr int c = 2 * decltype(cc)::value;
                        constexpr bool last = (c + 1 == NCH - 1);
                        constexpr int col = 16 * (c % NCHp);
                        float pa[16], pb[16];
                        tc::tmem_ld16(slot_lane + tc::oDm + col, pa);
                        tc::tmem_ld16(slot_lane + tc::oDm + col + 16, pb);
                        asm volatile("tcgen05.wait::ld.sync.aligned;");
                        if (last) {
                            asm volatile("tcgen05.fence::before_thread_sync;");
                            TC_TRACE(12, warp == kConsFirst && lane == 0);
                            __syncwarp();
                            if (lane == 0) tc::mbar_arrive(smem_addr(&slot_free[slot]));
                        }
                        float cost[C];
#pragma unroll
                        for (int i = 0; i < C; i++) cost[i] = -pa[i];
                        tr.template step_chunk<c>(cost);
#pragma unroll
                        for (int i = 0; i < C; i++) cost[i] = -pb[i];
                        tr.template step_chunk<c + 1>(cost);
                        if (dst) {
#pragma unroll
                            for (int i = 0; i < C; i++) dst[c * C + i] = pa[i], dst[(c + 1) * C + i] = pb[i];
                        }
                    };
                    // (measured: 128 states 11.9 -> 12.4 G sym/s, 256 states 5.65 -> 5.98; at 64 states 14.65 -> 14.45, so only from 128 on)
                    constexpr bool PAIRED = MVN_TC_CONS_PAIRED && L >= 7 && MERGED && !MLSE && !DUAL && C == 16 && (NCHp % 2 == 0);
                    if constexpr (PAIRED) {
                        tc::static_for<0, NCHp / 2>(chunk2);
                    } else if constexpr (!DUAL) {
                        tc::static_for<0, NCHp>(chunk);
                    } else {
                        if (lead) tc::static_for<0, NCHp / 2>(chunk);
                        else tc::static_for<NCHp / 2, NCHp>(chunk);
                    }
                    if constexpr (NPASS == 2) {
                        // every consumer warp has read the first half of the priors: one thread issues the second layer-3 pass
                        // into the same columns (h2 is still in the slot's A columns), then all wait for it
                        asm volatile("tcgen05.fence::before_thread_sync;");
                        asm volatile("bar.sync 3, %0;" ::"n"(32 * KW));
                        if (warp == kConsFirst && tc::elect_one()) {
                            asm volatile("tcgen05.fence::after_thread_sync;");
                            issue_layer3_merged(ts, 1);
                            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                                smem_addr(&d2_full[slot])));
                        }
                        __syncwarp();
                        tc::mbar_wait<MVN_CONS_PARK>(smem_addr(&d2_full[slot]), (NPASS * use + 1) & 1, timeout_flag);
                        asm volatile("tcgen05.fence::after_thread_sync;");
                        if constexpr (PAIRED) {
                            tc::static_for<NCHp / 2, NCH / 2>(chunk2);
                        } else if constexpr (!DUAL) {
                            tc::static_for<NCHp, NCH>(chunk);
                        } else {
                            if (lead) tc::static_for<NCHp, NCHp + NCHp / 2>(chunk);
                            else tc::static_for<NCHp + NCHp / 2, NCH>(chunk);
                        }
                    }
                    if constexpr (DUAL) {
                        float *xc = cons_xchg + (n & 1) * (2 * 32 * kQ) + quad * 32 + lane;
                        if (!lead) {
                            xc[0] = tr.ne;
                            xc[32 * kQ] = tr.no;
                        }
                        pair_sync();             // both halves of the new metrics are written, the old ones read
                        if (lead) {
                            tr.ne = fminf(tr.ne, xc[0]);
                            tr.no = fminf(tr.no, xc[32 * kQ]);
                        }
                    }
                    if (active) tr.commit();
                    if constexpr (MLSE) surv.put(t0 + tt, sv, t0 + tt == p.n_stages - 1);
                    TC_TRACE(13, warp == kConsFirst && lane == 0);
                }
                __syncwarp();
                if constexpr (!MLSE) {
                    if (active && lead) emit(t0, bits);
                }
            }
            if constexpr (MLSE) {
                __syncwarp();
                const int start = p.decision == MVN_DECIDE_MLSE_TERMINATED ? 0 : best_final_state<D::H>(tr);
                for (int t0 = ((p.T - 1) / 32) * 32; t0 >= ((p.n_stages + 31) / 32) * 32; t0 -= 32) emit(t0, 0u);
                traceback_frame<L>(surv, p.n_stages, start, [&](int tile_idx, uint32_t bits) { emit(tile_idx * 32, bits); });
                __syncwarp();
            }
            if (p.target) {
                const bool counted = active && lead && b < p.B && !(p.pilot_period > 0 && b % p.pilot_period == 0);
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
constexpr size_t tc_smem_bytes() {   // + the consumers' survivor masks in MLSE mode (launch_tc)
    constexpr int kQ = 4;
    return size_t(tc::kBBytes) + tc::b2_bytes(1 << L) + (size_t(L == 8 ? 0 : (tc::Roles<L>::PARTS + 1) * kQ) * kTileFloats + 4 * (tc::kK / 2)) * sizeof(float) +
           (L > 5 ? SmemTrellis<L>::bytes(32 * kQ) : 0);
}

}  // namespace mvn
