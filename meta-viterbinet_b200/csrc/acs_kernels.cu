// a2 / a3 / a5 — ACS stage, stage loop on a cost tensor, and the fused full-CSI Viterbi.
// Reference semantics: python_code/utils/trellis_utils.py:16-30, detectors/VA/va_detector.py:52-98.
#include <algorithm>
#include <type_traits>

#include "mvn_common.cuh"

namespace mvn {

// =====================================================================================
// a2: one stage on [B,S] tensors (drop-in for acs_block; not a hot kernel)
// =====================================================================================
__global__ void acs_block_kernel(const float *__restrict__ in_prob, const float *__restrict__ llrs,
                                 int llrs_stride, int64_t B, int S, float *__restrict__ out_prob,
                                 int64_t *__restrict__ out_idx) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= B * S) return;
    const int64_t b = i / S;
    const int j = int(i % S);
    const int s0 = (2 * j) % S, s1 = (2 * j + 1) % S;
    const float l0 = llrs_stride == 1 ? llrs[b] : llrs[b * llrs_stride + s0];
    const float l1 = llrs_stride == 1 ? llrs[b] : llrs[b * llrs_stride + s1];
    const float a = in_prob[b * S + s0] + l0;
    const float c = in_prob[b * S + s1] + l1;
    out_prob[i] = fminf(a, c);
    if (out_idx) out_idx[i] = (c < a) ? 1 : 0;
}

// =====================================================================================
// a3: stage loop on cost[B,T,S].  HBM-bound: S*4 bytes in, 4 bytes out per symbol.
// One lane per frame; the warp stages [32 frames x 32 floats] tiles of the [B, T*S] matrix.
// =====================================================================================
struct AcsParams {
    int layout;   // MVN_LAYOUT_*
    const float *cost;
    int64_t B;
    int T, n_stages;
    int out_format;
    void *decoded;
    float *final_pm;
    uint32_t *survivors;
    int64_t n_warp_tiles;
};

// path metrics in registers up to 64 states (32 distinct metrics + 32 new ones per frame), in shared memory beyond
constexpr int kRegTrellisMaxL = 6;
template <int L>
using TrellisFor = typename std::conditional<(L <= kRegTrellisMaxL), RegTrellis<L>, SmemTrellis<L>>::type;

// chunk c of a stage on either kind of trellis; for the register trellis c must fold to a constant (unrolled callers)
template <int L>
__device__ __forceinline__ uint32_t step_chunk_any(TrellisFor<L> &tr, int c, const float (&cc)[TrellisDims<L>::C]) {
    if constexpr (L <= kRegTrellisMaxL) {
        switch (c) {
            case 0: return tr.template step_chunk<0>(cc);
            case 1: return tr.template step_chunk<(TrellisDims<L>::NCH > 1 ? 1 : 0)>(cc);
            case 2: return tr.template step_chunk<(TrellisDims<L>::NCH > 2 ? 2 : 0)>(cc);
            default: return tr.template step_chunk<(TrellisDims<L>::NCH > 3 ? 3 : 0)>(cc);
        }
    } else {
        return tr.step_chunk_rt(c, cc);
    }
}

// Tiles of the [B, T*S] cost matrix stream through a per-warp ring of kAcsBuf buffers filled with cp.async: a warp keeps
// kAcsBuf - 1 tiles (4 KB each) in flight while it consumes the oldest, across the boundary between its frame tiles, so the
// kernel is bound by HBM and not by the latency of one tile at a time (which is what limited the 64..256-state trellises,
// where the path metrics in shared memory allow only four warps per SM).
template <int L> struct AcsCfg { static constexpr int NT = (L <= 5) ? 256 : 128, NBUF = (L <= 5) ? 2 : 4; };

template <int L, int NT>
__global__ void __launch_bounds__(NT, (L <= 5) ? 3 : (L == 6) ? 3 : 1) acs_decode_kernel(AcsParams p) {
    constexpr int kAcsBuf = AcsCfg<L>::NBUF;
    using D = TrellisDims<L>;
    constexpr int S = D::S, H = D::H, C = D::C, NCH = D::NCH;
    constexpr int WARPS = NT / 32;
    constexpr int SW = (H + 31) / 32;  // survivor words per stage
    constexpr int SPT = (S <= 32) ? 32 / S : 1;   // stages per tile
    constexpr int TPS = (S <= 32) ? 1 : S / 32;   // tiles per stage
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *ring = smem + warp * (kAcsBuf * kTileFloats);

    TrellisFor<L> tr;
    if constexpr (L > kRegTrellisMaxL) tr.init(smem + WARPS * kAcsBuf * kTileFloats, NT, threadIdx.x);

    const int64_t ld = int64_t(p.T) * S;
    const bool vec_in = is_vec_ok(p.cost, ld, ld);
    const bool vec_out = p.out_format == MVN_OUT_F32 && is_vec_ok(p.decoded, p.T, p.T);
    const int n_words = (p.T + 31) / 32;
    const int KT = (S <= 32) ? (p.n_stages + SPT - 1) / SPT : p.n_stages * TPS;   // tiles per frame tile
    const int64_t wt0 = int64_t(blockIdx.x) * WARPS + warp, wt_step = int64_t(gridDim.x) * WARPS;
    const int64_t my_tiles = wt0 < p.n_warp_tiles ? (p.n_warp_tiles - wt0 + wt_step - 1) / wt_step : 0;
    const int64_t total = my_tiles * KT;          // flat sequence of this warp's cost tiles

    auto issue = [&](int64_t idx) {               // tile idx of the flat sequence -> ring slot idx % kAcsBuf
        if (idx < total) {
            const int64_t wt = wt0 + (idx / KT) * wt_step;
            const int k = int(idx % KT);
            warp_load_tile_async(p.cost, p.B, ld, ld, wt * 32, int64_t(k) * 32, ring + (idx % kAcsBuf) * kTileFloats, lane);
        } else {
            cp_async_commit_empty();
        }
    };
    if (vec_in)
        for (int i = 0; i < kAcsBuf - 1; i++) issue(i);

    int64_t idx = 0;
    for (int64_t j = 0; j < my_tiles; j++) {
        const int64_t row0 = (wt0 + j * wt_step) * 32;
        const int64_t b = row0 + lane;
        tr.reset();
        uint32_t bits = 0;
        uint32_t sv = 0;
        for (int k = 0; k < KT; k++, idx++) {
            float *tile = ring + (idx % kAcsBuf) * kTileFloats;
            if (vec_in) {
                issue(idx + kAcsBuf - 1);          // refills the slot consumed in the previous iteration
                cp_async_wait<kAcsBuf - 1>();
                __syncwarp();
            } else {
                warp_load_tile(p.cost, p.B, ld, ld, row0, int64_t(k) * 32, tile, lane, false);
            }
            const float *row = tile + lane * kTileLd;
            if constexpr (S <= 32) {
#pragma unroll
                for (int u = 0; u < SPT; u++) {
                    const int t = k * SPT + u;
                    if (t < p.n_stages) {
                        bits |= tr.decide() << (t & 31);
                        float c0[C];
#pragma unroll
                        for (int i = 0; i < C; i++) c0[i] = row[u * S + i];
                        uint32_t s1 = tr.template step_chunk<0>(c0);
                        if constexpr (NCH == 2) {
                            float c1[C];
#pragma unroll
                            for (int i = 0; i < C; i++) c1[i] = row[u * S + C + i];
                            s1 |= tr.template step_chunk<1>(c1) << (C / 2);
                        }
                        tr.commit();
                        if (p.survivors && b < p.B) p.survivors[(b * p.n_stages + t) * SW] = s1;
                        if ((t & 31) == 31 || t == p.n_stages - 1) {
                            // (the store below contains shuffles: t is warp-uniform)
                            if (p.decoded) {
                                if (p.out_format == MVN_OUT_F32)
                                    warp_store_bits_f32(static_cast<float *>(p.decoded), p.B, p.T, p.T, row0, (t >> 5) * 32, bits,
                                                        lane, vec_out);
                                else if (b < p.B)
                                    static_cast<uint32_t *>(p.decoded)[b * n_words + (t >> 5)] = bits;
                            }
                            bits = 0;
                        }
                    }
                }
            } else {
                const int t = k / TPS, part = k % TPS;
                if (part == 0) bits |= tr.decide() << (t & 31);
#pragma unroll
                for (int u = 0; u < 2; u++) {
                    float cc[C];
#pragma unroll
                    for (int i = 0; i < C; i++) cc[i] = row[u * C + i];
                    const int c = 2 * part + u;
                    const uint32_t s8 = step_chunk_any<L>(tr, c, cc);
                    sv |= s8 << ((c * 8) & 31);
                    if ((c & 3) == 3 || c == NCH - 1) {
                        if (p.survivors && b < p.B) p.survivors[(b * p.n_stages + t) * SW + (c * 8) / 32] = sv;
                        sv = 0;
                    }
                }
                if (part == TPS - 1) {
                    tr.commit();
                    if ((t & 31) == 31 || t == p.n_stages - 1) {
                        if (p.decoded) {
                            if (p.out_format == MVN_OUT_F32)
                                warp_store_bits_f32(static_cast<float *>(p.decoded), p.B, p.T, p.T, row0, (t >> 5) * 32, bits, lane,
                                                    vec_out);
                            else if (b < p.B)
                                static_cast<uint32_t *>(p.decoded)[b * n_words + (t >> 5)] = bits;
                        }
                        bits = 0;
                    }
                }
            }
            __syncwarp();   // every lane is done with this ring slot before it is refilled
        }
        // output words past the stage loop stay 0 (decoded_word = zeros(y.shape), va_detector.py:89)
        if (p.decoded)
            for (int t0 = ((p.n_stages + 31) / 32) * 32; t0 < p.T; t0 += 32) {
                if (p.out_format == MVN_OUT_F32)
                    warp_store_bits_f32(static_cast<float *>(p.decoded), p.B, p.T, p.T, row0, t0, 0u, lane, vec_out);
                else if (b < p.B)
                    static_cast<uint32_t *>(p.decoded)[b * n_words + t0 / 32] = 0u;
            }
        if (p.final_pm && b < p.B) {
            for (int h = 0; h < H; h++) {
                const float v = tr.metric(h);
                p.final_pm[b * S + h] = v;
                if (S > 1) p.final_pm[b * S + h + H] = v;
            }
        }
    }
    if (vec_in) cp_async_wait<0>();
}

// =====================================================================================
// a3 for 64..256 states, STATES ON LANES: one warp owns one frame, lane l keeps the R = S/64 distinct path metrics
// pm[R l .. R l + R) in registers.  The reference trellis makes this layout shuffle-light: the two sources of new state j
// are 2j and 2j+1, i.e. ADJACENT metrics, and source s and s + S/2 share a metric, so a lane adds its R metrics to the 2R
// branch costs it loaded straight from HBM (two coalesced vector loads per lane and stage, no shared-memory staging), takes
// the pairwise minima locally (R >= 2) and the warp only has to re-deal the results: new[R l' + i] comes from the low halves
// of lanes 2l', 2l'+1 for l' < 16 and from the high halves of lanes 2(l'-16), 2(l'-16)+1 above (2R shuffles per stage).
// The decision (lowest state attaining the minimum, & 1) is one REDUX.MIN on order-preserving integer keys + a ballot.
// No per-frame shared memory at all, so the SMs run 48+ warps each and the kernel is bound by HBM, not by the latency of
// four lone warps walking shared-memory path metrics (what limited the lane-per-frame form at 128 / 256 states).
// =====================================================================================
__device__ __forceinline__ uint32_t order_key(float v) {   // monotone float -> uint map (+0 and -0 collapse first)
    const uint32_t u = __float_as_uint(v + 0.f);
    return u ^ ((u >> 31) ? 0xffffffffu : 0x80000000u);
}

template <int L>
__global__ void __launch_bounds__(256) acs_decode_warp_kernel(AcsParams p) {
    static_assert(L >= 6 && L <= 8, "states-on-lanes layout: 64..256 states");
    constexpr int S = 1 << L, H = S / 2, R = H / 32;
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = (int64_t(gridDim.x) * blockDim.x) >> 5;
    const int n_words = (p.T + 31) / 32;
    const bool lower = lane < 16;
    const int src = lower ? 2 * lane : 2 * (lane - 16);   // lanes whose results this lane collects
    for (int64_t b = warp_global; b < p.B; b += n_warps) {
        const float *cost = p.cost + b * int64_t(p.T) * S;
        float pm[R];
#pragma unroll
        for (int i = 0; i < R; i++) pm[i] = 0.f;
        uint32_t bits = 0;
        auto load = [&](int t, float (&lo)[R], float (&hi)[R]) {   // branch costs of sources R l + i and H + R l + i
            const float *c = cost + int64_t(t) * S + R * lane;
            if constexpr (R == 4) {
                const float4 a = ldg_stream4(c), d = ldg_stream4(c + H);
                lo[0] = a.x, lo[1] = a.y, lo[2] = a.z, lo[3] = a.w, hi[0] = d.x, hi[1] = d.y, hi[2] = d.z, hi[3] = d.w;
            } else if constexpr (R == 2) {
                const float2 a = ldg_stream2(c), d = ldg_stream2(c + H);
                lo[0] = a.x, lo[1] = a.y, hi[0] = d.x, hi[1] = d.y;
            } else {
                lo[0] = ldg_stream1(c);
                hi[0] = ldg_stream1(c + H);
            }
        };
        constexpr int U = 4;   // stages in flight per warp (8 was slower at 128 states: registers cost resident warps)
        float clo[U][R], chi[U][R];
#pragma unroll
        for (int u = 0; u < U; u++)
            if (u < p.n_stages) load(u, clo[u], chi[u]);
        for (int t0 = 0; t0 < p.n_stages; t0 += U) {
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int t = t0 + u;
                if (t < p.n_stages) {
                    // ---- decision on the metrics entering stage t
                    uint32_t kmin = order_key(pm[0]), par = (R == 1) ? uint32_t(lane & 1) : 0u;
#pragma unroll
                    for (int i = 1; i < R; i++) {
                        const uint32_t k = order_key(pm[i]);
                        if (k < kmin) {
                            kmin = k;
                            par = uint32_t(i & 1);
                        }
                    }
                    const uint32_t m = __reduce_min_sync(kFull, kmin);
                    const uint32_t who = __ballot_sync(kFull, kmin == m);
                    bits |= __shfl_sync(kFull, par, __ffs(who) - 1) << (t & 31);
                    // ---- ACS
                    float nlo[R >= 2 ? R / 2 : 1], nhi[R >= 2 ? R / 2 : 1];
                    if constexpr (R >= 2) {
#pragma unroll
                        for (int k = 0; k < R / 2; k++) {
                            nlo[k] = fminf(pm[2 * k] + clo[u][2 * k], pm[2 * k + 1] + clo[u][2 * k + 1]);
                            nhi[k] = fminf(pm[2 * k] + chi[u][2 * k], pm[2 * k + 1] + chi[u][2 * k + 1]);
                        }
#pragma unroll
                        for (int i = 0; i < R; i++) {   // new[R l' + i]: element i % (R/2) of lane src + i / (R/2)
                            const float a = __shfl_sync(kFull, nlo[i % (R / 2)], src + i / (R / 2));
                            const float c = __shfl_sync(kFull, nhi[i % (R / 2)], src + i / (R / 2));
                            pm[i] = lower ? a : c;
                        }
                    } else {
                        const float tl = pm[0] + clo[u][0], th = pm[0] + chi[u][0];
                        const float a = fminf(__shfl_sync(kFull, tl, src), __shfl_sync(kFull, tl, src + 1));
                        const float c = fminf(__shfl_sync(kFull, th, src), __shfl_sync(kFull, th, src + 1));
                        pm[0] = lower ? a : c;
                    }
                    if (t + U < p.n_stages) load(t + U, clo[u], chi[u]);
                    if ((t & 31) == 31 || t == p.n_stages - 1) {
                        if (p.decoded) {
                            if (p.out_format == MVN_OUT_F32) {
                                const int col = (t & ~31) + lane;
                                if (col < p.T) static_cast<float *>(p.decoded)[b * p.T + col] = float((bits >> lane) & 1u);
                            } else if (lane == 0) {
                                static_cast<uint32_t *>(p.decoded)[b * n_words + (t >> 5)] = bits;
                            }
                        }
                        bits = 0;
                    }
                }
            }
        }
        if (p.decoded)   // columns past the stage loop stay 0
            for (int t0 = ((p.n_stages + 31) / 32) * 32; t0 < p.T; t0 += 32) {
                if (p.out_format == MVN_OUT_F32) {
                    if (t0 + lane < p.T) static_cast<float *>(p.decoded)[b * p.T + t0 + lane] = 0.f;
                } else if (lane == 0) {
                    static_cast<uint32_t *>(p.decoded)[b * n_words + t0 / 32] = 0u;
                }
            }
        if (p.final_pm) {
#pragma unroll
            for (int i = 0; i < R; i++) {
                p.final_pm[b * S + R * lane + i] = pm[i];
                p.final_pm[b * S + H + R * lane + i] = pm[i];
            }
        }
    }
}

// States on lanes for the SMALL trellises (4..32 states) — the layout BASELINE.json's north star sketches ("the 16 states
// mapped to lanes, ACS butterflies with __shfl_sync").  G = S/2 lanes carry the distinct metrics of one frame (one each),
// a warp holds 32/G frames.  Built for the A/B against lane-per-frame; mvn_acs_decode_ex selects it.
template <int L>
__global__ void __launch_bounds__(256) acs_decode_group_kernel(AcsParams p) {
    static_assert(L >= 2 && L <= 5, "group layout: 4..32 states");
    constexpr int S = 1 << L, G = S / 2, FPW = 32 / G;   // lanes per frame, frames per warp
    const int lane = threadIdx.x & 31, sub = lane / G, j = lane % G, base = sub * G;
    const unsigned gmask = (G == 32) ? kFull : (((1u << G) - 1u) << base);
    const int64_t warp_global = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = (int64_t(gridDim.x) * blockDim.x) >> 5;
    const int n_words = (p.T + 31) / 32;
    const bool lower = j < G / 2;
    const int src = base + (lower ? 2 * j : 2 * (j - G / 2));
    for (int64_t f0 = warp_global * FPW; f0 < p.B; f0 += n_warps * FPW) {
        const int64_t b = f0 + sub;
        const bool live = b < p.B;
        const float *cost = p.cost + (live ? b : 0) * int64_t(p.T) * S;
        float pm = 0.f;
        uint32_t bits = 0;
        constexpr int U = 4;
        float clo[U], chi[U];
#pragma unroll
        for (int u = 0; u < U; u++)
            if (u < p.n_stages) {
                clo[u] = ldg_stream1(cost + int64_t(u) * S + j);
                chi[u] = ldg_stream1(cost + int64_t(u) * S + G + j);
            }
        for (int t0 = 0; t0 < p.n_stages; t0 += U) {
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int t = t0 + u;
                if (t < p.n_stages) {
                    const uint32_t k = order_key(pm);
                    const uint32_t m = __reduce_min_sync(gmask, k);
                    const uint32_t who = __ballot_sync(kFull, k == m) & gmask;
                    bits |= uint32_t((__ffs(who) - 1 - base) & 1) << (t & 31);   // lowest state attaining the minimum, & 1
                    const float tl = pm + clo[u], th = pm + chi[u];
                    const float a = fminf(__shfl_sync(kFull, tl, src), __shfl_sync(kFull, tl, src + 1));
                    const float c = fminf(__shfl_sync(kFull, th, src), __shfl_sync(kFull, th, src + 1));
                    pm = lower ? a : c;
                    if (t + U < p.n_stages) {
                        clo[u] = ldg_stream1(cost + int64_t(t + U) * S + j);
                        chi[u] = ldg_stream1(cost + int64_t(t + U) * S + G + j);
                    }
                    if ((t & 31) == 31 || t == p.n_stages - 1) {
                        if (p.decoded && live) {
                            if (p.out_format == MVN_OUT_F32) {
                                for (int col = (t & ~31) + j; col < min(p.T, (t & ~31) + 32); col += G)
                                    static_cast<float *>(p.decoded)[b * p.T + col] = float((bits >> (col & 31)) & 1u);
                            } else if (j == 0) {
                                static_cast<uint32_t *>(p.decoded)[b * n_words + (t >> 5)] = bits;
                            }
                        }
                        bits = 0;
                    }
                }
            }
        }
        if (p.decoded && live)
            for (int t0 = ((p.n_stages + 31) / 32) * 32; t0 < p.T; t0 += 32) {
                if (p.out_format == MVN_OUT_F32) {
                    for (int col = t0 + j; col < min(p.T, t0 + 32); col += G) static_cast<float *>(p.decoded)[b * p.T + col] = 0.f;
                } else if (j == 0) {
                    static_cast<uint32_t *>(p.decoded)[b * n_words + t0 / 32] = 0u;
                }
            }
        if (p.final_pm && live) {
            p.final_pm[b * S + j] = pm;
            p.final_pm[b * S + G + j] = pm;
        }
    }
}

// =====================================================================================
// a5+a3 fused: y[B,T] + state_priors[n_h,S] -> bits.  cost = (y-sp)^2/2 - ln sqrt(2 pi) in
// separately rounded fp32 ops (no contraction of d*d into the following op: the halving is exact,
// so fma(d*d, 0.5, -c) with d*d rounded first reproduces torch's three roundings bit for bit).
// =====================================================================================
struct VaParams {
    const float *y;
    int64_t B;
    int T, n_stages;
    const float *sp;  // [n_h, S]
    int n_h;
    int out_format;
    void *decoded;
    const float *target;
    int target_T, pilot_period;
    unsigned long long *counters;
    int64_t n_warp_tiles;
    int decision;      // MVN_DECIDE_*
    int surv_words;    // MLSE: survivor words per frame (SurvStore<L>::words(n_stages))
};

__device__ __forceinline__ float va_cost(float y, float sp) {
    const float d = __fsub_rn(y, sp);
    const float sq = __fmul_rn(d, d);
    return __fmaf_rn(sq, 0.5f, -kLogSqrt2Pi);
}

template <int L, int NT, bool MLSE>
__global__ void __launch_bounds__(NT, (L <= 5) ? (MLSE ? 2 : 4) : 1) va_decode_kernel(VaParams p) {
    using D = TrellisDims<L>;
    constexpr int S = D::S, H = D::H, C = D::C, NCH = D::NCH;
    constexpr int WARPS = NT / 32;
    constexpr bool SP_REGS = (S <= 32);
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *tile = smem + warp * kTileFloats;
    float *ttile = tile;   // the target tile of a block is staged after its y tile has been consumed
    const float *row = tile + lane * kTileLd;

    constexpr bool PACKED = (L >= 2 && L <= 5);
    using Tr = typename std::conditional<PACKED, PackedTrellis<PACKED ? L : 2>, TrellisFor<L>>::type;
    Tr tr;
    float *after_tiles = smem + WARPS * kTileFloats;
    if constexpr (L > kRegTrellisMaxL) {
        tr.init(after_tiles, NT, threadIdx.x);
        after_tiles += SmemTrellis<L>::bytes(NT) / sizeof(float);
    }
    SurvStore<L> surv;
    if constexpr (MLSE) surv.init(reinterpret_cast<uint32_t *>(after_tiles) + size_t(warp) * p.surv_words * 32, lane);

    const bool vec_in = is_vec_ok(p.y, p.T, p.T);
    const bool vec_out = p.out_format == MVN_OUT_F32 && is_vec_ok(p.decoded, p.T, p.T);
    const bool vec_tgt = p.target && is_vec_ok(p.target, p.target_T, p.target_T);
    const int n_words = (p.T + 31) / 32;
    ErrAcc acc;

    for (int64_t wt = int64_t(blockIdx.x) * WARPS + warp; wt < p.n_warp_tiles; wt += int64_t(gridDim.x) * WARPS) {
        const int64_t row0 = wt * 32;
        const int64_t b = row0 + lane;
        const float *sp_row = p.sp + int64_t((b < p.B ? b : 0) % p.n_h) * S;
        float sp[(SP_REGS && !PACKED) ? S : 1];
        u64_t nsp2[PACKED ? S / 2 : 1];  // (-sp[2i], -sp[2i+1]): y - sp == y + (-sp) exactly
        if constexpr (PACKED) {
#pragma unroll
            for (int i = 0; i < S / 2; i++) nsp2[i] = pk2(-__ldg(sp_row + 2 * i), -__ldg(sp_row + 2 * i + 1));
        } else if constexpr (SP_REGS) {
#pragma unroll
            for (int s = 0; s < S; s++) sp[s] = __ldg(sp_row + s);
        }
        tr.reset();
        unsigned frame_bit_errs = 0;
        // one tile of 32 decided bits: words out (full-line stores), optional BER against the staged target tile
        auto emit = [&](int t0, uint32_t bits) {
            if (p.decoded) {
                if (p.out_format == MVN_OUT_F32)
                    warp_store_bits_f32(static_cast<float *>(p.decoded), p.B, p.T, p.T, row0, t0, bits, lane, vec_out);
                else if (b < p.B)
                    static_cast<uint32_t *>(p.decoded)[b * n_words + t0 / 32] = bits;
            }
            if (p.target && t0 < p.target_T) {
                warp_load_tile(p.target, p.B, p.target_T, p.target_T, row0, t0, ttile, lane, vec_tgt);
                frame_bit_errs += tile_bit_errors(ttile + lane * kTileLd, bits, p.target_T - t0);
                __syncwarp();
            }
        };
        for (int t0 = 0; t0 < p.T; t0 += 32) {
            uint32_t bits = 0;
            const int t_end = min(32, p.n_stages - t0);
            if (t_end > 0) {
                // (prefetching the next tile into registers was measured slower: 271 vs 300 G sym/s — the extra
                //  32 registers cost a resident CTA and the kernel is issue-bound, not latency-bound)
                warp_load_tile(p.y, p.B, p.T, p.T, row0, t0, tile, lane, vec_in);
                for (int tt = 0; tt < t_end; tt++) {
                    const float yv = row[tt];
                    const int t = t0 + tt;
                    if constexpr (!MLSE) bits |= tr.decide() << tt;
                    if constexpr (PACKED) {
                        // d = y*1 + (-sp) (one rounding, == y - sp); sq = d*d; cost = sq*0.5 - ln sqrt(2 pi)
                        const u64_t yy = pk2(yv, yv), one2 = pk2(1.f, 1.f), half2 = pk2(0.5f, 0.5f);
                        const u64_t negk2 = pk2(-kLogSqrt2Pi, -kLogSqrt2Pi);
                        u64_t cost2[S / 2];
#pragma unroll
                        for (int i = 0; i < S / 2; i++) {
                            const u64_t d2 = fma2(yy, one2, nsp2[i]);
                            cost2[i] = fma2(mul2(d2, d2), half2, negk2);
                        }
                        const uint32_t sv = tr.template step<MLSE>(cost2);
                        if constexpr (MLSE) surv.put(t, sv, t == p.n_stages - 1);
                    } else if constexpr (SP_REGS) {
                        float c0[C];
#pragma unroll
                        for (int i = 0; i < C; i++) c0[i] = va_cost(yv, sp[i]);
                        uint32_t sv = tr.template step_chunk<0>(c0);
                        if constexpr (NCH == 2) {
                            float c1[C];
#pragma unroll
                            for (int i = 0; i < C; i++) c1[i] = va_cost(yv, sp[C + i]);
                            sv |= tr.template step_chunk<1>(c1) << (C / 2);
                        }
                        if constexpr (MLSE) surv.put(t, sv, t == p.n_stages - 1);
                    } else {
                        uint32_t sv = 0;
#pragma unroll(L <= kRegTrellisMaxL ? NCH : 1)
                        for (int c = 0; c < NCH; c++) {
                            float cc[C];
#pragma unroll
                            for (int i = 0; i < C; i += 4) {
                                const float4 s4 = __ldg(reinterpret_cast<const float4 *>(sp_row + c * C + i));
                                cc[i] = va_cost(yv, s4.x);
                                cc[i + 1] = va_cost(yv, s4.y);
                                cc[i + 2] = va_cost(yv, s4.z);
                                cc[i + 3] = va_cost(yv, s4.w);
                            }
                            const uint32_t s8 = step_chunk_any<L>(tr, c, cc);
                            if constexpr (MLSE) {
                                sv |= s8 << ((c * 8) & 31);
                                if ((c & 3) == 3) {
                                    if constexpr (H <= 32) surv.put(t, sv, t == p.n_stages - 1);
                                    else surv.put_word(t, c >> 2, sv);
                                    sv = 0;
                                }
                            }
                        }
                    }
                    if constexpr (!PACKED) tr.commit();
                }
                __syncwarp();
            }
            if constexpr (!MLSE) emit(t0, bits);
        }
        if constexpr (MLSE) {
            __syncwarp();
            const int start = p.decision == MVN_DECIDE_MLSE_TERMINATED ? 0 : best_final_state<H>(tr);
            for (int t0 = ((p.T - 1) / 32) * 32; t0 >= ((p.n_stages + 31) / 32) * 32; t0 -= 32) emit(t0, 0u);  // columns past the loop stay 0
            traceback_frame<L>(surv, p.n_stages, start, [&](int tile_idx, uint32_t bits) { emit(tile_idx * 32, bits); });
            __syncwarp();
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

// =====================================================================================
// a5+a3 for 128 / 256 states, STATES ON LANES (see acs_decode_warp_kernel): one warp per frame, lane l owns the metrics
// pm[R l .. R l + R) and the 2R noiseless outputs sp[R l + i], sp[S/2 + R l + i] of its sources; y comes in as one
// coalesced 32-sample load per 32 stages and is broadcast by shuffle.  Branch metrics with the reference's three
// separately rounded operations.  Reference rule only (the fused traceback stays with the lane-per-frame kernel).
// =====================================================================================
template <int L>
__global__ void __launch_bounds__(256) va_decode_warp_kernel(VaParams p) {
    static_assert(L >= 7 && L <= 8, "states-on-lanes VA: 128 / 256 states");
    constexpr int S = 1 << L, H = S / 2, R = H / 32;
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = (int64_t(gridDim.x) * blockDim.x) >> 5;
    const int n_words = (p.T + 31) / 32;
    const bool lower = lane < 16;
    const int src = lower ? 2 * lane : 2 * (lane - 16);
    unsigned bit_errs = 0, frame_errs = 0, frames = 0;
    for (int64_t b = warp_global; b < p.B; b += n_warps) {
        const float *sp_row = p.sp + int64_t(b % p.n_h) * S;
        float slo[R], shi[R], pm[R];
#pragma unroll
        for (int i = 0; i < R; i++) {
            slo[i] = __ldg(sp_row + R * lane + i);
            shi[i] = __ldg(sp_row + H + R * lane + i);
            pm[i] = 0.f;
        }
        unsigned frame_bit_errs = 0;
        for (int t0 = 0; t0 < p.T; t0 += 32) {
            uint32_t bits = 0;
            const int t_end = min(32, p.n_stages - t0);
            const float ytile = (t_end > 0 && t0 + lane < p.T) ? ldg_stream1(p.y + b * p.T + t0 + lane) : 0.f;
            for (int tt = 0; tt < t_end; tt++) {
                const float yv = __shfl_sync(kFull, ytile, tt);
                uint32_t kmin = order_key(pm[0]), par = 0u;
#pragma unroll
                for (int i = 1; i < R; i++) {
                    const uint32_t k = order_key(pm[i]);
                    if (k < kmin) {
                        kmin = k;
                        par = uint32_t(i & 1);
                    }
                }
                const uint32_t m = __reduce_min_sync(kFull, kmin);
                const uint32_t who = __ballot_sync(kFull, kmin == m);
                bits |= __shfl_sync(kFull, par, __ffs(who) - 1) << tt;
                float nlo[R / 2], nhi[R / 2];
#pragma unroll
                for (int k = 0; k < R / 2; k++) {
                    nlo[k] = fminf(pm[2 * k] + va_cost(yv, slo[2 * k]), pm[2 * k + 1] + va_cost(yv, slo[2 * k + 1]));
                    nhi[k] = fminf(pm[2 * k] + va_cost(yv, shi[2 * k]), pm[2 * k + 1] + va_cost(yv, shi[2 * k + 1]));
                }
#pragma unroll
                for (int i = 0; i < R; i++) {
                    const float a = __shfl_sync(kFull, nlo[i % (R / 2)], src + i / (R / 2));
                    const float c = __shfl_sync(kFull, nhi[i % (R / 2)], src + i / (R / 2));
                    pm[i] = lower ? a : c;
                }
            }
            if (p.decoded) {
                if (p.out_format == MVN_OUT_F32) {
                    if (t0 + lane < p.T) static_cast<float *>(p.decoded)[b * p.T + t0 + lane] = float((bits >> lane) & 1u);
                } else if (lane == 0) {
                    static_cast<uint32_t *>(p.decoded)[b * n_words + t0 / 32] = bits;
                }
            }
            if (p.target && t0 < p.target_T) {
                const bool in = t0 + lane < p.target_T;
                const int tgt = in ? int(ldg_stream1(p.target + b * p.target_T + t0 + lane)) : 0;
                frame_bit_errs += __popc(__ballot_sync(kFull, in && tgt != int((bits >> lane) & 1u)));
            }
        }
        if (p.target && !(p.pilot_period > 0 && b % p.pilot_period == 0)) {
            bit_errs += frame_bit_errs;
            frame_errs += frame_bit_errs ? 1u : 0u;
            frames += 1u;
        }
    }
    if (p.target && lane == 0 && frames) {
        atomicAdd(p.counters + MVN_CNT_BIT_ERRORS, (unsigned long long)bit_errs);
        atomicAdd(p.counters + MVN_CNT_FRAME_ERRORS, (unsigned long long)frame_errs);
        atomicAdd(p.counters + MVN_CNT_BITS, (unsigned long long)frames * unsigned(p.target_T));
        atomicAdd(p.counters + MVN_CNT_FRAMES, (unsigned long long)frames);
    }
}

// =====================================================================================
// host wrappers
// =====================================================================================
template <int L>
static int launch_acs(const AcsParams &p, cudaStream_t st) {
    if constexpr (L >= 2 && L <= 5) {
        if (p.layout == MVN_LAYOUT_STATES_ON_LANES && !p.survivors) {   // only on request: lane-per-frame is faster here
            auto kern = acs_decode_group_kernel<L>;
            int per_sm = 1;
            MVN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, 0));
            const int64_t fpw = 32 / ((1 << L) / 2);
            const int grid = int(std::min<int64_t>((p.B + 8 * fpw - 1) / (8 * fpw), int64_t(sm_count()) * std::max(per_sm, 1)));
            kern<<<grid, 256, 0, st>>>(p);
            note_launch();
            MVN_CUDA(cudaGetLastError());
            return MVN_OK;
        }
    }
    if constexpr (L >= 6) {
        // states on lanes (one warp per frame): the default at 128 / 256 states (at 64 the lane-per-frame kernel with its
        // register trellis is faster: 65 % vs 39 % of HBM), unless the survivor export is wanted (that stays with the
        // lane-per-frame kernel, which assembles the bit masks per lane); needs 16-byte aligned rows for the vector loads
        const bool want = p.layout == MVN_LAYOUT_STATES_ON_LANES || (p.layout == MVN_LAYOUT_AUTO && L >= 7);
        if (want && !p.survivors && (reinterpret_cast<uintptr_t>(p.cost) & 15u) == 0) {
            auto kern = acs_decode_warp_kernel<L>;
            int per_sm = 1;
            MVN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, 0));
            if (per_sm < 1) per_sm = 1;
            const int grid = int(std::min<int64_t>((p.B + 7) / 8, int64_t(sm_count()) * per_sm));
            kern<<<grid, 256, 0, st>>>(p);
            note_launch();
            MVN_CUDA(cudaGetLastError());
            return MVN_OK;
        }
    }
    constexpr int NT = AcsCfg<L>::NT;
    size_t smem = size_t(NT / 32) * AcsCfg<L>::NBUF * kTileFloats * sizeof(float);
    if (L > kRegTrellisMaxL) smem += SmemTrellis<L>::bytes(NT);
    auto kern = acs_decode_kernel<L, NT>;
    MVN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    int per_sm = 1;
    MVN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem));
    if (per_sm < 1) per_sm = 1;
    const int64_t need = (p.n_warp_tiles + NT / 32 - 1) / (NT / 32);
    const int grid = int(std::min<int64_t>(need, int64_t(sm_count()) * per_sm));
    kern<<<grid, NT, smem, st>>>(p);
    note_launch();
    MVN_CUDA(cudaGetLastError());
    return MVN_OK;
}

template <int L, bool MLSE>
static int launch_va_mode(VaParams p, cudaStream_t st) {
    // MLSE at 128 / 256 states: the survivor masks (T x S/2 bits per frame) share the CTA's shared memory with the
    // path metrics, so fewer frames are resident per CTA
    constexpr int NT = (L <= 5) ? 256 : (MLSE && L == 7) ? 64 : (MLSE && L == 8) ? 32 : 128;
    size_t smem = size_t(NT / 32) * kTileFloats * sizeof(float);
    if (L > kRegTrellisMaxL) smem += SmemTrellis<L>::bytes(NT);
    if (MLSE) {
        p.surv_words = SurvStore<L>::words(p.n_stages);
        smem += size_t(NT / 32) * SurvStore<L>::bytes_per_warp(p.n_stages);
        if (smem > 227 * 1024) {
            set_error("mvn_va_decode (MLSE): %d stages of %d survivor bits do not fit the shared memory of one CTA", p.n_stages,
                      SurvStore<L>::H);
            return MVN_ERR_UNSUPPORTED;
        }
    }
    auto kern = va_decode_kernel<L, NT, MLSE>;
    MVN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    int per_sm = 1;
    MVN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem));
    if (per_sm < 1) per_sm = 1;
    const int64_t need = (p.n_warp_tiles + NT / 32 - 1) / (NT / 32);
    const int grid = int(std::min<int64_t>(need, int64_t(sm_count()) * per_sm));
    kern<<<grid, NT, smem, st>>>(p);
    note_launch();
    MVN_CUDA(cudaGetLastError());
    return MVN_OK;
}

template <int L>
static int launch_va(const VaParams &p, cudaStream_t st) {
    if constexpr (L >= 7) {
        if (p.decision == MVN_DECIDE_REFERENCE) {   // 128 / 256 states: one warp per frame, states on lanes
            auto kern = va_decode_warp_kernel<L>;
            int per_sm = 1;
            MVN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, 0));
            const int grid = int(std::min<int64_t>((p.B + 7) / 8, int64_t(sm_count()) * std::max(per_sm, 1)));
            kern<<<grid, 256, 0, st>>>(p);
            note_launch();
            MVN_CUDA(cudaGetLastError());
            return MVN_OK;
        }
    }
    return p.decision == MVN_DECIDE_REFERENCE ? launch_va_mode<L, false>(p, st) : launch_va_mode<L, true>(p, st);
}

#define MVN_DISPATCH_L(L, FN, ...)                                   \
    switch (L) {                                                     \
        case 1: return FN<1>(__VA_ARGS__);                           \
        case 2: return FN<2>(__VA_ARGS__);                           \
        case 3: return FN<3>(__VA_ARGS__);                           \
        case 4: return FN<4>(__VA_ARGS__);                           \
        case 5: return FN<5>(__VA_ARGS__);                           \
        case 6: return FN<6>(__VA_ARGS__);                           \
        case 7: return FN<7>(__VA_ARGS__);                           \
        case 8: return FN<8>(__VA_ARGS__);                           \
        default: set_error("memory_length %d outside [1,8]", L);     \
                 return MVN_ERR_ARG;                                 \
    }

int acs_decode_impl(const AcsParams &p, int L, cudaStream_t st) { MVN_DISPATCH_L(L, launch_acs, p, st) }
int va_decode_impl(const VaParams &p, int L, cudaStream_t st) { MVN_DISPATCH_L(L, launch_va, p, st) }

}  // namespace mvn

using namespace mvn;

extern "C" int mvn_acs_block(const float *in_prob, const float *llrs, int llrs_stride, int64_t B, int L,
                             float *out_prob, int64_t *out_idx, void *stream) {
    if (!in_prob || !llrs || !out_prob || L < 1 || L > 8 || B < 0) {
        set_error("mvn_acs_block: bad argument");
        return MVN_ERR_ARG;
    }
    const int S = 1 << L;
    if (llrs_stride != 1 && llrs_stride != S) {
        set_error("mvn_acs_block: llrs_stride must be 1 or n_states");
        return MVN_ERR_ARG;
    }
    if (B == 0) return MVN_OK;
    const int64_t n = B * S;
    acs_block_kernel<<<unsigned((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        in_prob, llrs, llrs_stride, B, S, out_prob, out_idx);
    note_launch();
    MVN_CUDA(cudaGetLastError());
    return MVN_OK;
}

extern "C" int mvn_acs_decode_ex(const float *cost, int64_t B, int T, int L, int n_stages, int out_format, void *decoded,
                                 float *final_pm, uint32_t *survivors, int layout, void *stream) {
    if (L < 1 || L > 8) {
        set_error("memory_length %d outside [1,8]", L);
        return MVN_ERR_ARG;
    }
    if (B < 0 || T < 0 || n_stages < 0 || n_stages > T || (B > 0 && T > 0 && !cost) ||
        (out_format != MVN_OUT_F32 && out_format != MVN_OUT_BITS) || layout < MVN_LAYOUT_AUTO || layout > MVN_LAYOUT_STATES_ON_LANES) {
        set_error("mvn_acs_decode: bad argument (B=%lld T=%d n_stages=%d layout=%d)", (long long)B, T, n_stages, layout);
        return MVN_ERR_ARG;
    }
    if (B == 0 || T == 0) return MVN_OK;
    AcsParams p{layout, cost, B, T, n_stages, out_format, decoded, final_pm, survivors, (B + 31) / 32};
    return acs_decode_impl(p, L, static_cast<cudaStream_t>(stream));
}

extern "C" int mvn_acs_decode(const float *cost, int64_t B, int T, int L, int n_stages, int out_format,
                              void *decoded, float *final_pm, uint32_t *survivors, void *stream) {
    return mvn_acs_decode_ex(cost, B, T, L, n_stages, out_format, decoded, final_pm, survivors, MVN_LAYOUT_AUTO, stream);
}

extern "C" int mvn_va_decode_ex(const float *y, int64_t B, int T, int L, int n_stages, const float *state_priors,
                                int n_h, int out_format, void *decoded, const float *target, int target_T,
                                int pilot_period, uint64_t *counters, int decision, void *stream) {
    if (L < 1 || L > 8) {
        set_error("memory_length %d outside [1,8]", L);
        return MVN_ERR_ARG;
    }
    if (B < 0 || T < 0 || n_stages < 0 || n_stages > T || !state_priors || n_h < 1 || (B > 0 && T > 0 && !y) ||
        (out_format != MVN_OUT_F32 && out_format != MVN_OUT_BITS) || decision < MVN_DECIDE_REFERENCE ||
        decision > MVN_DECIDE_MLSE_TERMINATED) {
        set_error("mvn_va_decode: bad argument (B=%lld T=%d n_stages=%d n_h=%d decision=%d)", (long long)B, T, n_stages, n_h,
                  decision);
        return MVN_ERR_ARG;
    }
    if (B % n_h != 0) {  // the reference's tiling of the table fails to broadcast (va_detector.py:64-66)
        set_error("mvn_va_decode: batch %lld is not a multiple of the %d tap blocks", (long long)B, n_h);
        return MVN_ERR_ARG;
    }
    if (target && (!counters || target_T < 1 || target_T > T)) {
        set_error("mvn_va_decode: target needs counters and 1 <= target_T <= T");
        return MVN_ERR_ARG;
    }
    if (B == 0 || T == 0) return MVN_OK;
    VaParams p{y, B, T, n_stages, state_priors, n_h, out_format, decoded, target, target_T, pilot_period,
               reinterpret_cast<unsigned long long *>(counters), (B + 31) / 32, decision, 0};
    return va_decode_impl(p, L, static_cast<cudaStream_t>(stream));
}

extern "C" int mvn_va_decode(const float *y, int64_t B, int T, int L, int n_stages, const float *state_priors,
                             int n_h, int out_format, void *decoded, const float *target, int target_T,
                             int pilot_period, uint64_t *counters, void *stream) {
    return mvn_va_decode_ex(y, B, T, L, n_stages, state_priors, n_h, out_format, decoded, target, target_T, pilot_period,
                            counters, MVN_DECIDE_REFERENCE, stream);
}
