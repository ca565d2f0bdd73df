// a6 / a7 + a3 — ViterbiNet priors MLP (1 -> 100 sigmoid -> 50 relu -> S) and the fused
// priors + ACS + decision kernel.  Reference semantics: detectors/VNET/vnet_detector.py:27-61,
// detectors/META_VNET/meta_vnet_detector.py:24-45.
//
// FP32-pipe bound (11 832 flop and 8 HBM bytes per symbol at S=16), so the design goal is to
// keep the FMA pipe busy:
//   * one lane owns M frames; per stage it evaluates the MLP for its M samples entirely in
//     registers (100 layer-2 accumulators for M=2) and then runs the ACS on its private metrics;
//   * weights live in shared memory, transposed so that one warp-uniform (broadcast) 128-bit load
//     feeds two packed fma.rn.f32x2 (SASS FFMA2) per frame: the packed form needs one issue slot
//     per two FMAs, which leaves slots for the LDS/MUFU/ALU work of the same warp;
//   * sigmoid(a) = 1/(1+2^(-a log2 e)) with -log2 e folded into W1/b1 when the weights are staged:
//     one FFMA + MUFU.EX2 + FADD + MUFU.RCP per hidden unit.
#include <algorithm>
#include <mutex>
#include <type_traits>

#include "vnet_mlp.cuh"

namespace mvn {

// =====================================================================================
// a6: priors only, symbol-parallel ('train' phase forward, parity export, small batches)
// =====================================================================================
template <int L, int NT>
__global__ void __launch_bounds__(NT) vnet_priors_kernel(const float *__restrict__ y, int64_t N, VnetWeights w,
                                                          float *__restrict__ priors) {
    using D = TrellisDims<L>;
    constexpr int S = D::S, C = D::C, NCH = D::NCH, M = 2;
    extern __shared__ __align__(16) float smem[];
    stage_weights<L>(smem, w, threadIdx.x, NT);
    __syncthreads();
    const Wt<kSmem> wt{smem_addr(smem)};
    const int lane = threadIdx.x & 31;
    const int64_t n_warps = (int64_t(gridDim.x) * NT) >> 5;
    for (int64_t base = ((int64_t(blockIdx.x) * NT + threadIdx.x) >> 5) * 64; base < N; base += n_warps * 64) {
        int64_t n[M];
        float yv[M];
#pragma unroll
        for (int m = 0; m < M; m++) {
            n[m] = base + 32 * m + lane;
            yv[m] = n[m] < N ? y[n[m]] : 0.f;
        }
        float h2[M][kH2];
        mlp_hidden<L, M, kSmem, kKUnroll>(wt, yv, h2);
        for (int c = 0; c < NCH; c++) {
            float p[M][C];
            mlp_out_chunk<L, M, kSmem>(wt, c, h2, p);
#pragma unroll
            for (int m = 0; m < M; m++) {
                if (n[m] < N) {
                    float *dst = priors + n[m] * S + c * C;
                    if constexpr (C >= 4) {
#pragma unroll
                        for (int i = 0; i < C; i += 4)
                            *reinterpret_cast<float4 *>(dst + i) = make_float4(p[m][i], p[m][i + 1], p[m][i + 2], p[m][i + 3]);
                    } else {
#pragma unroll
                        for (int i = 0; i < C; i++) dst[i] = p[m][i];
                    }
                }
            }
        }
    }
}

// writes the staged layout into the global staging slot (constant-bank path)
template <int L>
__global__ void stage_weights_kernel(VnetWeights w, int slot) {
    stage_weights<L>(mvn_gStage + slot * kConstSlotFloats, w, threadIdx.x, blockDim.x);
}

// =====================================================================================
// a6+a3 fused
// =====================================================================================
struct VnetParams {
    const float *y;
    int64_t B;
    int T, n_stages;
    VnetWeights w;
    int out_format;
    void *decoded;
    float *priors_out;
    const float *target;
    int target_T, pilot_period;
    unsigned long long *counters;
    int64_t n_warp_tiles;  // tiles of 32*M frames
    int const_slot;
    int variant;           // MVN_VARIANT_*: which implementation runs (per call, no global state)
    int decision;          // MVN_DECIDE_*
    int surv_words;        // MLSE: survivor words per frame (SurvStore<L>::words(n_stages))
};

// Variant = (frames per lane, weight source, threads per CTA, layer-2 unroll).  One CTA per SM.
template <int L, int M_, int WS_, int NT_, int KU_>
struct FusedVariant {
    static constexpr int M = M_, WS = WS_, NT = NT_, KU = KU_;
    static constexpr bool kRegPm = (L <= 4);
    static constexpr size_t smem_bytes() {
        size_t fl = (WS == kSmem ? VnetSmem<L>::kFloats : 0) + size_t(NT / 32) * M * kTileFloats;
        size_t bytes = fl * sizeof(float);
        if (!kRegPm) bytes += SmemTrellis<L>::bytes(NT * M);
        return bytes;
    }
};

template <int L, class V>
__global__ void __launch_bounds__(V::NT, 1) vnet_decode_kernel(VnetParams p) {
    using D = TrellisDims<L>;
    using W = VnetSmem<L>;
    constexpr int S = D::S, C = D::C, NCH = D::NCH, M = V::M, NT = V::NT, WS = V::WS;
    constexpr int WARPS = NT / 32;
    constexpr int kWFloats = (WS == kSmem) ? W::kFloats : 0;
    using Tr = typename std::conditional<V::kRegPm, RegTrellis<L>, SmemTrellis<L>>::type;
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *tiles = smem + kWFloats + warp * (M * kTileFloats);

    Wt<WS> wt;
    if constexpr (WS == kSmem) {
        stage_weights<L>(smem, p.w, threadIdx.x, NT);
        __syncthreads();
        wt.base = smem_addr(smem);
    } else {
        wt.base = const_params_addr() + uint32_t(p.const_slot * kConstSlotFloats * 4);
    }

    Tr tr[M];
    if constexpr (!V::kRegPm) {
#pragma unroll
        for (int m = 0; m < M; m++) tr[m].init(smem + kWFloats + WARPS * M * kTileFloats, NT * M, m * NT + threadIdx.x);
    }

    const bool vec_in = is_vec_ok(p.y, p.T, p.T);
    const bool vec_out = p.out_format == MVN_OUT_F32 && is_vec_ok(p.decoded, p.T, p.T);
    const bool vec_tgt = p.target && is_vec_ok(p.target, p.target_T, p.target_T);
    const int n_words = (p.T + 31) / 32;
    ErrAcc acc;

    // CTA-uniform trip count (all warps of a CTA run the same number of tiles, out-of-range frames are
    // masked): keeps the control flow provably warp-uniform, which the uniform datapath (LDCU, UR
    // operands) of the constant-bank variants requires.
    const int64_t n_cta_tiles = (p.n_warp_tiles + WARPS - 1) / WARPS;
    for (int64_t ct = blockIdx.x; ct < n_cta_tiles; ct += gridDim.x) {
        const int64_t row0 = (ct * WARPS + warp) * (32 * M);
        unsigned frame_bit_errs[M];
#pragma unroll
        for (int m = 0; m < M; m++) {
            tr[m].reset();
            frame_bit_errs[m] = 0;
        }
        for (int t0 = 0; t0 < p.T; t0 += 32) {
            uint32_t bits[M];
#pragma unroll
            for (int m = 0; m < M; m++) bits[m] = 0;
            const int t_end = min(32, p.n_stages - t0);
            if (t_end > 0) {
#pragma unroll
                for (int m = 0; m < M; m++)
                    warp_load_tile(p.y, p.B, p.T, p.T, row0 + 32 * m, t0, tiles + m * kTileFloats, lane, vec_in);
#pragma unroll 1
                for (int tt = 0; tt < t_end; tt++) {
                    float yv[M];
#pragma unroll
                    for (int m = 0; m < M; m++) yv[m] = tiles[m * kTileFloats + lane * kTileLd + tt];
                    float h2[M][kH2];
                    mlp_hidden<L, M, WS, V::KU>(wt, yv, h2);
#pragma unroll
                    for (int m = 0; m < M; m++) bits[m] |= tr[m].decide() << tt;
                    if constexpr (V::kRegPm) {
                        static_assert(!V::kRegPm || NCH == 1, "register trellis in the fused kernel: S <= 16");
                        float pr[M][C];
                        mlp_out_chunk<L, M, WS>(wt, 0, h2, pr);
#pragma unroll
                        for (int m = 0; m < M; m++) {
                            float cost[C];
#pragma unroll
                            for (int i = 0; i < C; i++) cost[i] = -pr[m][i];  // vnet_detector.py:57
                            tr[m].template step_chunk<0>(cost);
                            tr[m].commit();
                            const int64_t b = row0 + 32 * m + lane;
                            if (p.priors_out && b < p.B) {
                                float *dst = p.priors_out + (b * p.T + t0 + tt) * S;
#pragma unroll
                                for (int i = 0; i < C; i++) dst[i] = pr[m][i];
                            }
                        }
                    } else {
                        for (int c = 0; c < NCH; c++) {
                            float pr[M][C];
                            mlp_out_chunk<L, M, WS>(wt, c, h2, pr);
#pragma unroll
                            for (int m = 0; m < M; m++) {
                                float cost[C];
#pragma unroll
                                for (int i = 0; i < C; i++) cost[i] = -pr[m][i];
                                tr[m].step_chunk_rt(c, cost);
                                const int64_t b = row0 + 32 * m + lane;
                                if (p.priors_out && b < p.B) {
                                    float *dst = p.priors_out + (b * p.T + t0 + tt) * S + c * C;
#pragma unroll
                                    for (int i = 0; i < C; i++) dst[i] = pr[m][i];
                                }
                            }
                        }
#pragma unroll
                        for (int m = 0; m < M; m++) tr[m].commit();
                    }
                }
                __syncwarp();
            }
#pragma unroll
            for (int m = 0; m < M; m++) {
                const int64_t b = row0 + 32 * m + lane;
                if (p.decoded) {
                    if (p.out_format == MVN_OUT_F32)
                        warp_store_bits_f32(static_cast<float *>(p.decoded), p.B, p.T, p.T, row0 + 32 * m, t0, bits[m],
                                            lane, vec_out);
                    else if (b < p.B)
                        static_cast<uint32_t *>(p.decoded)[b * n_words + t0 / 32] = bits[m];
                }
                if (p.target && t0 < p.target_T) {
                    float *tt_tile = tiles + m * kTileFloats;  // the y tile of this block is consumed
                    warp_load_tile(p.target, p.B, p.target_T, p.target_T, row0 + 32 * m, t0, tt_tile, lane, vec_tgt);
                    frame_bit_errs[m] += tile_bit_errors(tt_tile + lane * kTileLd, bits[m], p.target_T - t0);
                    __syncwarp();
                }
            }
        }
        if (p.target) {
#pragma unroll
            for (int m = 0; m < M; m++) {
                const int64_t b = row0 + 32 * m + lane;
                const bool counted = b < p.B && !(p.pilot_period > 0 && b % p.pilot_period == 0);
                if (counted) {
                    acc.bit_errs += frame_bit_errs[m];
                    acc.frame_errs += frame_bit_errs[m] ? 1u : 0u;
                    acc.bits += unsigned(p.target_T);
                    acc.frames += 1u;
                }
            }
            acc.flush(p.counters);
        }
    }
}

}  // namespace mvn
#include "vnet_tc_kernel.cuh"
namespace mvn {

// The two constant-bank slots are the only state shared between calls.  Each slot carries an event
// recorded after the kernel that read it; the next call that wants the slot makes its stream wait
// on that event before overwriting it, so calls on different streams (and host threads) stay safe.
struct ConstSlots {
    cudaEvent_t ev[2] = {nullptr, nullptr};
    int next = 0;
};
static ConstSlots g_slots[64];
static std::mutex g_slot_mu;

template <int L, class V>
static int launch_variant(VnetParams p, cudaStream_t st) {
    const size_t smem = V::smem_bytes();
    auto kern = vnet_decode_kernel<L, V>;
    MVN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    p.n_warp_tiles = (p.B + 32 * V::M - 1) / (32 * V::M);
    const int warps = V::NT / 32;
    const int64_t need = (p.n_warp_tiles + warps - 1) / warps;
    const int grid = int(std::min<int64_t>(need, sm_count()));
    if (V::WS == kConst) {
        static_assert(V::WS != kConst || VnetSmem<L>::kFloats <= kConstSlotFloats, "constant slot too small");
        int dev = 0;
        MVN_CUDA(cudaGetDevice(&dev));
        std::lock_guard<std::mutex> lock(g_slot_mu);
        ConstSlots &cs = g_slots[dev & 63];
        for (int i = 0; i < 2; i++)
            if (!cs.ev[i]) MVN_CUDA(cudaEventCreateWithFlags(&cs.ev[i], cudaEventDisableTiming));
        const int slot = cs.next;
        cs.next ^= 1;
        p.const_slot = slot;
        MVN_CUDA(cudaStreamWaitEvent(st, cs.ev[slot], 0));
        stage_weights_kernel<L><<<1, 256, 0, st>>>(p.w, slot);
        note_launch();
        void *src = nullptr;
        MVN_CUDA(cudaGetSymbolAddress(&src, mvn_gStage));
        MVN_CUDA(cudaMemcpyToSymbolAsync(mvn_cParams, static_cast<float *>(src) + slot * kConstSlotFloats,
                                         VnetSmem<L>::kFloats * sizeof(float),
                                         size_t(slot) * kConstSlotFloats * sizeof(float), cudaMemcpyDeviceToDevice, st));
        kern<<<grid, V::NT, smem, st>>>(p);
        note_launch();
        MVN_CUDA(cudaGetLastError());
        MVN_CUDA(cudaEventRecord(cs.ev[slot], st));
        return MVN_OK;
    }
    kern<<<grid, V::NT, smem, st>>>(p);
    note_launch();
    MVN_CUDA(cudaGetLastError());
    return MVN_OK;
}

// Pipeline watchdog of the tcgen05 kernel (tc::TcWatch): one __device__ flag per GPU that the kernel's waits poll, and one
// word per GPU in mapped pinned host memory that the kernel writes once if a wait gives up.  The host reads that word
// without a synchronisation: before a launch (a set flag refuses the launch loudly until mvn_reset_tc_timeout()) and, in
// the host-buffer entry points, after their own stream syncs — so a call that timed out does not return MVN_OK.
__device__ int g_tc_watch_dev = 0;
static int *g_tc_watch_host = nullptr, *g_tc_watch_host_dev = nullptr;   // [64] words, one per device ordinal
static std::once_flag g_tc_watch_once;
static bool tc_watch_init() {
    std::call_once(g_tc_watch_once, [] {
        if (cudaHostAlloc(reinterpret_cast<void **>(&g_tc_watch_host), 64 * sizeof(int),
                          cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) {
            g_tc_watch_host = nullptr;
            return;
        }
        for (int i = 0; i < 64; i++) g_tc_watch_host[i] = 0;
        if (cudaHostGetDevicePointer(reinterpret_cast<void **>(&g_tc_watch_host_dev), g_tc_watch_host, 0) != cudaSuccess)
            g_tc_watch_host_dev = nullptr;
    });
    return g_tc_watch_host && g_tc_watch_host_dev;
}
int tc_timeout_seen() {
    int dev = 0;
    if (!g_tc_watch_host || cudaGetDevice(&dev) != cudaSuccess) return 0;
    return g_tc_watch_host[dev & 63];
}
#ifdef MVN_TC_TRACE
__device__ long long g_tc_trace[64 * 32];
#endif

// tcgen05 variant (memory_length 1..8): every CTA stages its own weights (W2/b2 and W3/b3 as fp16 pieces in shared memory).
template <int L, bool MLSE>
static int launch_tc_mode(VnetParams p, cudaStream_t st) {
    size_t smem = tc_smem_bytes<L>();
    if (MLSE) {
        p.surv_words = SurvStore<L>::words(p.n_stages);
        smem += size_t(tc::kConsWarps) * SurvStore<L>::bytes_per_warp(p.n_stages);
        if (smem > 227 * 1024) {
            set_error("mvn_vnet_decode (MLSE): %d stages of %d survivor bits do not fit the shared memory of one CTA",
                      p.n_stages, SurvStore<L>::H);
            return MVN_ERR_UNSUPPORTED;
        }
    }
    auto kern = vnet_decode_tc_kernel<L, MLSE>;
    MVN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    p.n_warp_tiles = (p.B + 31) / 32;
    constexpr int kQ = 4;                  // 32-frame warp tiles per CTA tile
    const int64_t need = (p.n_warp_tiles + kQ - 1) / kQ;
    const int grid = int(std::min<int64_t>(need, sm_count()));
    int dev = 0;
    MVN_CUDA(cudaGetDevice(&dev));
    if (!tc_watch_init()) {
        set_error("vnet_decode (tcgen05): cannot allocate the mapped pipeline-watchdog word");
        return MVN_ERR_CUDA;
    }
    if (g_tc_watch_host[dev & 63]) {
        set_error("vnet_decode (tcgen05): a pipeline wait timed out in an earlier launch on this device; its results "
                  "are invalid.  Call mvn_reset_tc_timeout() to re-arm, or select MVN_VARIANT_FMA");
        return MVN_ERR_CUDA;
    }
    tc::TcWatch flag{nullptr, g_tc_watch_host_dev + (dev & 63)};
    MVN_CUDA(cudaGetSymbolAddress(reinterpret_cast<void **>(&flag.dev), g_tc_watch_dev));
    void *trace = nullptr;
#ifdef MVN_TC_TRACE
    MVN_CUDA(cudaGetSymbolAddress(&trace, g_tc_trace));
#endif
    kern<<<grid, tc::Roles<L>::kThreads, smem, st>>>(p, flag, static_cast<long long *>(trace));
    note_launch();
    MVN_CUDA(cudaGetLastError());
    return MVN_OK;
}

// reference decision rule for every L; the fused traceback (MLSE) instance is built for memory_length <= 6
template <int L>
static int launch_tc(const VnetParams &p, cudaStream_t st) {
    if (p.decision == MVN_DECIDE_REFERENCE) return launch_tc_mode<L, false>(p, st);
    if constexpr (L <= 6) {
        return launch_tc_mode<L, true>(p, st);
    } else {
        set_error("vnet_decode (tcgen05): fused traceback is built for memory_length <= 6");
        return MVN_ERR_UNSUPPORTED;
    }
}

// Default variants, picked by tools/tune_fused.py on a B200 (profiles/r01_tune_fused.txt):
//   L 1..6: tcgen05 variant (vnet_tc_kernel.cuh): layers 2-3 on the tensor cores with an fp16 two-piece split;
//           the FP32-FMA variants below stay selectable (mvn_debug_set_variant / ops.set_fused_variant)
//   L 4..5: (FMA) constant-bank weights, 2 frames per lane, 384 threads (3 warps per scheduler), unroll 10
//   L 1..3: (FMA) same with 448 threads.  ptxas only routes the weight stream through uniform registers
//           (LDCU) when vector registers are scarce; with the smaller layer-3 tile of these trellises
//           the 384-thread build falls back to per-lane LDC, which is 2x slower.
//           tools/check_sass.py asserts LDCU in the hot loop of every default variant.
//   L 6..7: staged weights in shared memory (the constant bank is too small for W3), 128 threads
//   L = 8 : as above with one frame per lane (the path metrics take 131 KB of shared memory)
template <int L>
static int launch_fused(const VnetParams &p, cudaStream_t st) {
    const bool want_fma = p.variant == MVN_VARIANT_FMA_SMEM || p.variant == MVN_VARIANT_FMA_CONST320 || p.variant == MVN_VARIANT_FMA;
    if (!want_fma) return launch_tc<L>(p, st);  // auto and tcgen05
    if constexpr (L <= 4) {
        switch (p.variant) {
            case 1: return launch_variant<L, FusedVariant<L, 2, kSmem, 256, 10>>(p, st);
            case 2: return launch_variant<L, FusedVariant<L, 2, kConst, 320, 10>>(p, st);
            default: return launch_variant<L, FusedVariant<L, 2, kConst, (L <= 3 ? 448 : 384), 10>>(p, st);
        }
    } else if constexpr (L == 5) {
        return launch_variant<L, FusedVariant<L, 2, kConst, 384, 10>>(p, st);
    } else if constexpr (L <= 7) {
        return launch_variant<L, FusedVariant<L, 2, kSmem, 128, 5>>(p, st);
    } else {
        return launch_variant<L, FusedVariant<L, 1, kSmem, 128, 5>>(p, st);
    }
}

// frames decoded by one full wave of CTAs (host pipeline chunk sizing)
int vnet_frames_per_wave(int L, int variant) {
    const int per_cta = !(variant == 1 || variant == 2 || variant == 4) ? 128
                        : L <= 3 ? 448 * 2 : L <= 5 ? 384 * 2 : L <= 7 ? 128 * 2 : 128;
    return per_cta * sm_count();
}

template <int L>
static int launch_priors(const float *y, int64_t N, const VnetWeights &w, float *priors, cudaStream_t st) {
    constexpr int NT = 128;
    const size_t smem = VnetSmem<L>::kFloats * sizeof(float);
    auto kern = vnet_priors_kernel<L, NT>;
    MVN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    int per_sm = 1;
    MVN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem));
    if (per_sm < 1) per_sm = 1;
    const int64_t need = (N + (NT / 32) * 64 - 1) / ((NT / 32) * 64);
    const int grid = int(std::min<int64_t>(need, int64_t(sm_count()) * per_sm));
    kern<<<grid, NT, smem, st>>>(y, N, w, priors);
    note_launch();
    MVN_CUDA(cudaGetLastError());
    return MVN_OK;
}

#define MVN_DISPATCH_L(L, FN, ...)                               \
    switch (L) {                                                 \
        case 1: return FN<1>(__VA_ARGS__);                       \
        case 2: return FN<2>(__VA_ARGS__);                       \
        case 3: return FN<3>(__VA_ARGS__);                       \
        case 4: return FN<4>(__VA_ARGS__);                       \
        case 5: return FN<5>(__VA_ARGS__);                       \
        case 6: return FN<6>(__VA_ARGS__);                       \
        case 7: return FN<7>(__VA_ARGS__);                       \
        case 8: return FN<8>(__VA_ARGS__);                       \
        default: set_error("memory_length %d outside [1,8]", L); \
                 return MVN_ERR_ARG;                             \
    }

int vnet_decode_impl(const VnetParams &p, int L, cudaStream_t st) { MVN_DISPATCH_L(L, launch_fused, p, st) }
int vnet_priors_impl(const float *y, int64_t N, int L, const VnetWeights &w, float *priors, cudaStream_t st) {
    MVN_DISPATCH_L(L, launch_priors, y, N, w, priors, st)
}

}  // namespace mvn

using namespace mvn;

static bool weights_ok(const float *w1, const float *b1, const float *w2, const float *b2, const float *w3,
                       const float *b3) {
    return w1 && b1 && w2 && b2 && w3 && b3;
}

extern "C" int mvn_vnet_priors(const float *y, int64_t N, int L, const float *w1, const float *b1, const float *w2,
                               const float *b2, const float *w3, const float *b3, float *priors, void *stream) {
    if (L < 1 || L > 8) {
        set_error("memory_length %d outside [1,8]", L);
        return MVN_ERR_ARG;
    }
    if (N < 0 || !weights_ok(w1, b1, w2, b2, w3, b3) || (N > 0 && (!y || !priors))) {
        set_error("mvn_vnet_priors: bad argument");
        return MVN_ERR_ARG;
    }
    if (N == 0) return MVN_OK;
    VnetWeights w{w1, b1, w2, b2, w3, b3};
    return vnet_priors_impl(y, N, L, w, priors, static_cast<cudaStream_t>(stream));
}

extern "C" int mvn_vnet_decode_ex(const float *y, int64_t B, int T, int L, int n_stages, const float *w1,
                                  const float *b1, const float *w2, const float *b2, const float *w3, const float *b3,
                                  int out_format, void *decoded, float *priors_out, const float *target, int target_T,
                                  int pilot_period, uint64_t *counters, int variant, int decision, void *stream) {
    if (L < 1 || L > 8) {
        set_error("memory_length %d outside [1,8]", L);
        return MVN_ERR_ARG;
    }
    if (B < 0 || T < 0 || n_stages < 0 || n_stages > T || !weights_ok(w1, b1, w2, b2, w3, b3) ||
        (B > 0 && T > 0 && !y) || (out_format != MVN_OUT_F32 && out_format != MVN_OUT_BITS)) {
        set_error("mvn_vnet_decode: bad argument (B=%lld T=%d n_stages=%d)", (long long)B, T, n_stages);
        return MVN_ERR_ARG;
    }
    if (variant < MVN_VARIANT_AUTO || variant > MVN_VARIANT_FMA) {
        set_error("mvn_vnet_decode: unknown kernel variant %d", variant);
        return MVN_ERR_ARG;
    }
    if (decision < MVN_DECIDE_REFERENCE || decision > MVN_DECIDE_MLSE_TERMINATED) {
        set_error("mvn_vnet_decode: unknown decision mode %d", decision);
        return MVN_ERR_ARG;
    }
    if (decision != MVN_DECIDE_REFERENCE && (L > 6 || variant == MVN_VARIANT_FMA_SMEM || variant == MVN_VARIANT_FMA_CONST320 ||
                                             variant == MVN_VARIANT_FMA)) {
        set_error("mvn_vnet_decode: the fused in-kernel traceback runs in the tensor-core kernel (memory_length <= 6, "
                  "MVN_VARIANT_AUTO / MVN_VARIANT_TCGEN05); use mvn_vnet_priors + mvn_acs_decode + mvn_traceback otherwise");
        return MVN_ERR_UNSUPPORTED;
    }
    if (target && (!counters || target_T < 1 || target_T > T)) {
        set_error("mvn_vnet_decode: target needs counters and 1 <= target_T <= T");
        return MVN_ERR_ARG;
    }
    if (B == 0 || T == 0) return MVN_OK;
    VnetParams p{y, B, T, n_stages, VnetWeights{w1, b1, w2, b2, w3, b3}, out_format, decoded, priors_out, target,
                 target_T, pilot_period, reinterpret_cast<unsigned long long *>(counters), 0, 0, variant, decision, 0};
    return vnet_decode_impl(p, L, static_cast<cudaStream_t>(stream));
}

extern "C" int mvn_vnet_decode(const float *y, int64_t B, int T, int L, int n_stages, const float *w1,
                               const float *b1, const float *w2, const float *b2, const float *w3, const float *b3,
                               int out_format, void *decoded, float *priors_out, const float *target, int target_T,
                               int pilot_period, uint64_t *counters, void *stream) {
    return mvn_vnet_decode_ex(y, B, T, L, n_stages, w1, b1, w2, b2, w3, b3, out_format, decoded, priors_out, target,
                              target_T, pilot_period, counters, MVN_VARIANT_AUTO, MVN_DECIDE_REFERENCE, stream);
}

/* 1 if a tcgen05 pipeline wait timed out on the current device since the last reset */
extern "C" int mvn_tc_timeout_status(void) { return mvn::tc_timeout_seen(); }

extern "C" int mvn_reset_tc_timeout(void) {
    int dev = 0;
    MVN_CUDA(cudaGetDevice(&dev));
    const int zero = 0;
    MVN_CUDA(cudaMemcpyToSymbol(mvn::g_tc_watch_dev, &zero, sizeof(int)));
    if (mvn::g_tc_watch_host) mvn::g_tc_watch_host[dev & 63] = 0;
    return MVN_OK;
}

#ifdef MVN_TC_TRACE
extern "C" int mvn_debug_tc_trace(long long *host_out) {
    return int(cudaMemcpyFromSymbol(host_out, mvn::g_tc_trace, sizeof(long long) * 64 * 32));
}
#endif
