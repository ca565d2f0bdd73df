// Common device/host helpers for the sm_100a detection kernels.
//
// Thread mapping used by every decode kernel (DESIGN.md "thread owns frames"):
//   one LANE owns one frame (or M frames) for its whole length; the S/2 distinct path
//   metrics of the reference trellis (pm[j] == pm[j+S/2]) live in that lane's
//   registers (L<=5) or in a private shared-memory column (L>=6), so a stage is S adds and
//   S/2 mins with no cross-lane traffic.  The 32 lanes of a warp move data between HBM and
//   their frames through 32x32 fp32 tiles staged in shared memory, so every global access is a
//   full 128-byte line even though each lane walks its own row.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/mvn_b200.h"

namespace mvn {

constexpr int kH1 = MVN_HIDDEN1;
constexpr int kH2 = MVN_HIDDEN2;
// fp32(ln sqrt(2 pi)); va_detector.py:68 subtracts the double constant as an fp32 scalar.
constexpr float kLogSqrt2Pi = 0.9189385175704956f;
constexpr unsigned kFull = 0xffffffffu;

// ---------------------------------------------------------------- host-side error plumbing
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);
void note_launch();
#define MVN_CUDA(call)                                               \
    do {                                                             \
        cudaError_t e__ = (call);                                    \
        if (e__ != cudaSuccess) return ::mvn::cuda_fail(e__, #call); \
    } while (0)

int sm_count();

// ---------------------------------------------------------------- tiles
// A warp-private tile holds 32 rows x 32 fp32; rows padded to 36 floats: 16-byte aligned rows,
// and both the row-wise 128-bit fills and the "lane reads its own row" accesses are
// bank-conflict free ((36*lane + c) mod 32 = 4*lane + c).
constexpr int kTileLd = 36;
constexpr int kTileFloats = 32 * kTileLd;

// Plain read-only loads: with L1::no_allocate the lines take the L2 evict-first class, and the second
// 32-byte sector of each 64-byte DRAM fetch (rows are 480 B, tiles 128 B) is dropped before the next
// tile of the same rows wants it — ncu showed 1.75x the algorithmic DRAM reads (profiles/r01).
__device__ __forceinline__ float4 ldg_stream4(const float *p) {
    float4 v;
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ float2 ldg_stream2(const float *p) {
    float2 v;
    asm volatile("ld.global.nc.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ float ldg_stream1(const float *p) {
    float v;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

// Rows [row0,row0+32) x columns [c0,c0+32) of a row-major matrix (leading dimension ld floats,
// `ncols` valid columns, `nrows` valid rows) -> tile.  Out-of-range elements read as 0.
// vec: ld % 4 == 0, ncols % 4 == 0 and 16-byte aligned base (checked by the host wrapper).
__device__ __forceinline__ void warp_load_tile(const float *__restrict__ src, int64_t nrows, int64_t ld,
                                               int64_t ncols, int64_t row0, int64_t c0, float *tile,
                                               int lane, bool vec) {
    if (vec) {
        float4 v[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int r = 4 * i + (lane >> 3);
            const int c = (lane & 7) * 4;
            const int64_t row = row0 + r;
            v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row < nrows && c0 + c < ncols) v[i] = ldg_stream4(src + row * ld + c0 + c);
        }
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int r = 4 * i + (lane >> 3);
            const int c = (lane & 7) * 4;
            *reinterpret_cast<float4 *>(tile + r * kTileLd + c) = v[i];
        }
    } else {
#pragma unroll 8
        for (int r = 0; r < 32; r++) {
            const int64_t row = row0 + r;
            float v = 0.f;
            if (row < nrows && c0 + lane < ncols) v = ldg_stream1(src + row * ld + c0 + lane);
            tile[r * kTileLd + lane] = v;
        }
    }
    __syncwarp();
}

// Asynchronous form of the vectorised tile load (cp.async, SASS LDGSTS): the copies of one tile form one commit group, so
// a warp can keep several tiles in flight and wait for the oldest only (cp_async_wait<N>, then __syncwarp).  Needs the
// vec conditions of warp_load_tile; out-of-range elements are zero-filled (src-size 0).
__device__ __forceinline__ void warp_load_tile_async(const float *__restrict__ src, int64_t nrows, int64_t ld, int64_t ncols,
                                                     int64_t row0, int64_t c0, float *tile, int lane) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int r = 4 * i + (lane >> 3);
        const int c = (lane & 7) * 4;
        const int64_t row = row0 + r;
        const bool ok = row < nrows && c0 + c < ncols;
        const float *g = ok ? src + row * ld + c0 + c : src;
        const uint32_t dst = uint32_t(__cvta_generic_to_shared(tile + r * kTileLd + c));
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(g), "r"(ok ? 16 : 0) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void cp_async_commit_empty() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// Each lane holds a 32-step decision mask for row row0+lane; the warp writes them as fp32 0/1
// into dst[row][c0 .. c0+32) with full-line stores.
__device__ __forceinline__ void warp_store_bits_f32(float *__restrict__ dst, int64_t nrows, int64_t ld,
                                                    int64_t ncols, int64_t row0, int64_t c0, uint32_t bits,
                                                    int lane, bool vec) {
    if (vec) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int r = 4 * i + (lane >> 3);
            const int c = (lane & 7) * 4;
            const uint32_t m = __shfl_sync(kFull, bits, r) >> c;
            const int64_t row = row0 + r;
            if (row < nrows && c0 + c < ncols) {
                float4 v = make_float4(float(m & 1u), float((m >> 1) & 1u), float((m >> 2) & 1u),
                                       float((m >> 3) & 1u));
                *reinterpret_cast<float4 *>(dst + row * ld + c0 + c) = v;
            }
        }
    } else {
#pragma unroll 8
        for (int r = 0; r < 32; r++) {
            const uint32_t m = __shfl_sync(kFull, bits, r);
            const int64_t row = row0 + r;
            if (row < nrows && c0 + lane < ncols) dst[row * ld + c0 + lane] = float((m >> lane) & 1u);
        }
    }
}

__device__ __forceinline__ bool is_vec_ok(const void *p, int64_t ld, int64_t ncols) {
    return ((reinterpret_cast<uintptr_t>(p) & 15u) == 0) && (ld % 4 == 0) && (ncols % 4 == 0);
}

// ---------------------------------------------------------------- trellis engines
// Both engines implement Appendix A of SURVEY.md exactly:
//   decide(): (lowest s attaining min_s pm[s]) & 1  — the argmin always lies in [0,S/2)
//   step   : tmp[s] = pm[s mod H] + cost[s];  pm'[j] = min(tmp[2j], tmp[2j+1]), j < H = S/2
// fminf == torch.min for the non-NaN values that occur; ties keep the lower index because the
// scan only replaces on a strict '<'.
template <int L>
struct TrellisDims {
    static constexpr int S = 1 << L;
    static constexpr int H = (S >= 2) ? S / 2 : 1;
    static constexpr int C = (S < 16) ? S : 16;  // source states consumed per chunk
    static constexpr int NCH = S / C;
};

// Register-resident path metrics (L <= 5: at most 16 distinct metrics per frame).
template <int L>
struct RegTrellis {
    static constexpr int S = TrellisDims<L>::S, H = TrellisDims<L>::H, C = TrellisDims<L>::C;
    float pm[H];
    float nw[H];
    __device__ __forceinline__ void reset() {
#pragma unroll
        for (int h = 0; h < H; h++) pm[h] = 0.f;
    }
    __device__ __forceinline__ uint32_t decide() const {
        float best = pm[0];
        uint32_t bit = 0;
#pragma unroll
        for (int h = 1; h < H; h++) {
            const bool lt = pm[h] < best;
            best = fminf(best, pm[h]);
            bit = lt ? uint32_t(h & 1) : bit;
        }
        return bit;
    }
    // Chunk c covers source states [c*C, c*C+C) and produces new states [c*C/2, c*C/2 + C/2).
    // cost[i] is the branch cost of source state c*C+i.  Returns the C/2 survivor bits.
    template <int c>
    __device__ __forceinline__ uint32_t step_chunk(const float (&cost)[C]) {
        uint32_t surv = 0;
#pragma unroll
        for (int i = 0; i < C / 2; i++) {
            const float a = pm[(c * C + 2 * i) % H] + cost[2 * i];
            const float b = pm[(c * C + 2 * i + 1) % H] + cost[2 * i + 1];
            nw[c * (C / 2) + i] = fminf(a, b);
            surv |= uint32_t(b < a) << i;
        }
        return surv;
    }
    __device__ __forceinline__ void commit() {
#pragma unroll
        for (int h = 0; h < H; h++) pm[h] = nw[h];
    }
    __device__ __forceinline__ float metric(int h) const { return pm[h]; }
};

// Shared-memory path metrics (L >= 6): column `col` of a [2][H][ncols] fp32 array, one column
// per frame, consecutive threads -> consecutive columns (conflict-free).
template <int L>
struct SmemTrellis {
    static constexpr int S = TrellisDims<L>::S, H = TrellisDims<L>::H, C = TrellisDims<L>::C;
    float *base;  // &array[0][0][col]
    int ncols;
    int cur;
    __device__ __forceinline__ void init(float *array, int ncols_, int col) {
        base = array + col;
        ncols = ncols_;
        cur = 0;
    }
    static __host__ __device__ constexpr size_t bytes(int ncols_) { return size_t(2) * H * ncols_ * sizeof(float); }
    __device__ __forceinline__ float &at(int buf, int h) const { return base[(size_t(buf) * H + h) * ncols]; }
    // minima over the even / odd states of the current metrics (e, o) and of the metrics being built (ne, no): the
    // decision needs only their order; an exact tie falls back to the ordered scan (lowest index wins)
    float e, o, ne, no;
    __device__ __forceinline__ void reset() {
        for (int h = 0; h < H; h++) at(0, h) = 0.f;
        cur = 0;
        e = o = 0.f;
        ne = no = 3.0e38f;
    }
    __device__ __forceinline__ uint32_t decide() const {
        if (o != e) return o < e ? 1u : 0u;
        float best = at(cur, 0);
        uint32_t bit = 0;
#pragma unroll 8
        for (int h = 1; h < H; h++) {
            const float v = at(cur, h);
            const bool lt = v < best;
            best = fminf(best, v);
            bit = lt ? uint32_t(h & 1) : bit;
        }
        return bit;
    }
    __device__ __forceinline__ uint32_t step_chunk_rt(int c, const float (&cost)[C]) {
        // all loads of the chunk first, then the arithmetic, then the stores: the compiler must assume that a store to the
        // next-metrics buffer aliases the following loads of the current one, so an interleaved loop serialises on the
        // shared-memory latency of every pair (measured: ~400 cycles per 16-state chunk)
        float old_[C];
#pragma unroll
        for (int i = 0; i < C; i++) old_[i] = at(cur, (c * C + i) % H);
        uint32_t surv = 0;
        float v[C / 2];
#pragma unroll
        for (int i = 0; i < C / 2; i++) {
            const float a = old_[2 * i] + cost[2 * i];
            const float b = old_[2 * i + 1] + cost[2 * i + 1];
            v[i] = fminf(a, b);
            if (i & 1) no = fminf(no, v[i]);   // new state c*(C/2)+i: C/2 is even for every trellis kept in shared memory
            else ne = fminf(ne, v[i]);
            surv |= uint32_t(b < a) << i;
        }
#pragma unroll
        for (int i = 0; i < C / 2; i++) at(cur ^ 1, c * (C / 2) + i) = v[i];
        return surv;
    }
    __device__ __forceinline__ void commit() {
        cur ^= 1;
        e = ne, o = no;
        ne = no = 3.0e38f;
    }
    __device__ __forceinline__ float metric(int h) const { return at(cur, h); }
};

// Same trellis with the column count known at compile time and two base pointers swapped at commit: every access of a
// fully unrolled stage is one LDS / STS with an immediate offset (the generic form above pays an address computation per
// access).  Used by the consumer warps of the tensor-core kernel.
template <int L, int NCOLS>
struct SmemTrellisFixed {
    static constexpr int S = TrellisDims<L>::S, H = TrellisDims<L>::H, C = TrellisDims<L>::C;
    float *cur_p, *nxt_p;   // &array[buf][0][col]
    float e, o, ne, no;
    __device__ __forceinline__ void init(float *array, int /*ncols*/, int col) {
        cur_p = array + col;
        nxt_p = array + size_t(H) * NCOLS + col;
    }
    static __host__ __device__ constexpr size_t bytes(int) { return size_t(2) * H * NCOLS * sizeof(float); }
    __device__ __forceinline__ void reset() {
#pragma unroll 8
        for (int h = 0; h < H; h++) cur_p[h * NCOLS] = 0.f;
        e = o = 0.f;
        ne = no = 3.0e38f;
    }
    __device__ __forceinline__ uint32_t decide() const {
        if (o != e) return o < e ? 1u : 0u;
        float best = cur_p[0];
        uint32_t bit = 0;
#pragma unroll 8
        for (int h = 1; h < H; h++) {
            const float v = cur_p[h * NCOLS];
            const bool lt = v < best;
            best = fminf(best, v);
            bit = lt ? uint32_t(h & 1) : bit;
        }
        return bit;
    }
    template <int c>
    __device__ __forceinline__ uint32_t step_chunk(const float (&cost)[C]) {
        float old_[C];   // loads, then arithmetic, then stores (see SmemTrellis::step_chunk_rt)
#pragma unroll
        for (int i = 0; i < C; i++) old_[i] = cur_p[((c * C + i) % H) * NCOLS];
        uint32_t surv = 0;
        float v[C / 2];
#pragma unroll
        for (int i = 0; i < C / 2; i++) {
            const float a = old_[2 * i] + cost[2 * i];
            const float b = old_[2 * i + 1] + cost[2 * i + 1];
            v[i] = fminf(a, b);
            if (i & 1) no = fminf(no, v[i]);
            else ne = fminf(ne, v[i]);
            surv |= uint32_t(b < a) << i;
        }
#pragma unroll
        for (int i = 0; i < C / 2; i++) nxt_p[(c * (C / 2) + i) * NCOLS] = v[i];
        return surv;
    }
    __device__ __forceinline__ void commit() {
        float *t = cur_p;
        cur_p = nxt_p;
        nxt_p = t;
        e = ne, o = no;
        ne = no = 3.0e38f;
    }
    __device__ __forceinline__ float metric(int h) const { return cur_p[h * NCOLS]; }
};

// ---------------------------------------------------------------- packed fp32x2 helpers (sm_100)
// add/mul/fma.rn.f32x2 round each half exactly like the scalar instruction, so bit-exactness with
// the reference's fp32 arithmetic is preserved while the instruction count halves.
typedef unsigned long long u64_t;
__device__ __forceinline__ u64_t pk2(float a, float b) {
    u64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void upk2(u64_t v, float &a, float &b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64_t fma2(u64_t a, u64_t b, u64_t c) {
    u64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ u64_t mul2(u64_t a, u64_t b) {
    u64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ u64_t add2(u64_t a, u64_t b) {
    u64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

// Register trellis on fp32x2 pairs (2 <= L <= 5).  Source states 2i and 2i+1 are exactly the pair
// (pm[2i mod H], pm[2i+1 mod H]) = pm2[i mod H/2], so one packed add serves both candidates of new
// state i.  decide() uses the even/odd minima; only an exact tie between them (common at the first
// stages and with integer costs, rare otherwise) takes the ordered scan.
template <int L>
struct PackedTrellis {
    static constexpr int S = 1 << L, H = S / 2, HP = H / 2;
    static_assert(L >= 2 && L <= 5, "packed trellis: 4..32 states");
    u64_t pm2[HP];
    __device__ __forceinline__ void reset() {
#pragma unroll
        for (int i = 0; i < HP; i++) pm2[i] = 0ull;
    }
    __device__ __forceinline__ uint32_t decide() const {
        float v[H];
#pragma unroll
        for (int i = 0; i < HP; i++) upk2(pm2[i], v[2 * i], v[2 * i + 1]);
        float e = v[0], o = v[1];
#pragma unroll
        for (int i = 1; i < HP; i++) {
            e = fminf(e, v[2 * i]);
            o = fminf(o, v[2 * i + 1]);
        }
        uint32_t bit = (o < e) ? 1u : 0u;
        if (o == e) {  // lowest index among equal minima decides
            float best = v[0];
            bit = 0;
#pragma unroll
            for (int h = 1; h < H; h++) {
                const bool lt = v[h] < best;
                best = fminf(best, v[h]);
                bit = lt ? uint32_t(h & 1) : bit;
            }
        }
        return bit;
    }
    // cost2[i] = packed branch costs of source states (2i, 2i+1), i < H.  SURV: also return the H survivor bits
    // (bit i = 1 iff the odd predecessor 2i+1 won strictly; ties keep predecessor 2i like torch.min, trellis_utils.py:30)
    template <bool SURV = false>
    __device__ __forceinline__ uint32_t step(const u64_t (&cost2)[H]) {
        float nw[H];
        uint32_t surv = 0;
#pragma unroll
        for (int i = 0; i < H; i++) {
            float a, b;
            upk2(add2(pm2[i % HP], cost2[i]), a, b);
            nw[i] = fminf(a, b);
            if (SURV) surv |= (b < a) ? (1u << i) : 0u;
        }
#pragma unroll
        for (int i = 0; i < HP; i++) pm2[i] = pk2(nw[2 * i], nw[2 * i + 1]);
        return surv;
    }
    __device__ __forceinline__ float metric(int h) const {
        float v[H];
#pragma unroll
        for (int i = 0; i < HP; i++) upk2(pm2[i], v[2 * i], v[2 * i + 1]);
        float r = v[0];
#pragma unroll
        for (int i = 1; i < H; i++) r = (i == h) ? v[i] : r;
        return r;
    }
};

// ---------------------------------------------------------------- in-kernel survivor store + traceback (true MLSE)
// The H = S/2 survivor bits of every stage (what acs_block returns as indices, trellis_utils.py:30, and the
// reference's detectors discard) are kept per frame as bit masks in shared memory: a warp owns words[w][32 lanes]
// (lane-contiguous: conflict-free), SPW stages per 32-bit word for H <= 32, H/32 words per stage above.  At the end of
// the frame each lane walks its own masks backwards: state(t) = sum_i b[t+i] 2^i, the predecessor of j is
// (2j + sigma) mod S and sigma IS the transmitted bit b[t].
template <int L>
struct SurvStore {
    static constexpr int S = 1 << L, H = (S >= 2) ? S / 2 : 1;
    static constexpr int SPW = (H <= 32) ? 32 / H : 1;   // stages per word
    static constexpr int WPS = (H <= 32) ? 1 : H / 32;   // words per stage
    __host__ __device__ static constexpr int words(int n_stages) {
        return (H <= 32) ? (n_stages + SPW - 1) / SPW : n_stages * WPS;
    }
    __host__ __device__ static constexpr size_t bytes_per_warp(int n_stages) { return size_t(words(n_stages)) * 32 * sizeof(uint32_t); }
    uint32_t *base;   // &words[0][lane]
    uint32_t acc;
    __device__ __forceinline__ void init(uint32_t *warp_words, int lane) {
        base = warp_words + lane;
        acc = 0;
    }
    // all H bits of stage t at once (H <= 32)
    __device__ __forceinline__ void put(int t, uint32_t sv, bool last_stage) {
        static_assert(H <= 32, "put(): whole-stage form");
        if constexpr (SPW == 1) {
            base[t * 32] = sv;
        } else {
            const int k = t % SPW;
            acc |= sv << (k * H);
            if (k == SPW - 1 || last_stage) {
                base[(t / SPW) * 32] = acc;
                acc = 0;
            }
        }
    }
    // word w of stage t (H > 32)
    __device__ __forceinline__ void put_word(int t, int w, uint32_t bits) { base[(t * WPS + w) * 32] = bits; }
    __device__ __forceinline__ uint32_t get(int t, int jj) const {
        if constexpr (H <= 32) {
            return (base[(t / SPW) * 32] >> ((t % SPW) * H + jj)) & 1u;
        } else {
            return (base[(t * WPS + (jj >> 5)) * 32] >> (jj & 31)) & 1u;
        }
    }
};

// Lane-private traceback over stages [0, n_stages): calls emit(tile, bits) (warp-uniform control flow) with the 32
// decided bits of stages [32 tile, 32 tile + 32), highest tile first.
template <int L, class Emit>
__device__ __forceinline__ void traceback_frame(const SurvStore<L> &sv, int n_stages, int start_state, Emit emit) {
    constexpr int S = 1 << L, H = SurvStore<L>::H;
    int j = start_state;
    uint32_t bits = 0;
    for (int t = n_stages - 1; t >= 0; t--) {
        const uint32_t sigma = sv.get(t, j & (H - 1));
        j = (2 * j + int(sigma)) & (S - 1);
        bits |= sigma << (t & 31);
        if ((t & 31) == 0) {
            emit(t >> 5, bits);
            bits = 0;
        }
    }
}

// lowest state index attaining the minimum final metric (always < H: states j and j + H are identical)
template <int H, class Tr>
__device__ __forceinline__ int best_final_state(const Tr &tr) {
    float best = tr.metric(0);
    int j = 0;
#pragma unroll(H <= 32 ? H : 8)
    for (int h = 1; h < H; h++) {
        const float v = tr.metric(h);
        if (v < best) {
            best = v;
            j = h;
        }
    }
    return j;
}

// ---------------------------------------------------------------- fused BER/FER accounting
struct ErrAcc {
    unsigned bit_errs = 0, frame_errs = 0, bits = 0, frames = 0;
    __device__ __forceinline__ void flush(unsigned long long *counters) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            bit_errs += __shfl_xor_sync(kFull, bit_errs, o);
            frame_errs += __shfl_xor_sync(kFull, frame_errs, o);
            bits += __shfl_xor_sync(kFull, bits, o);
            frames += __shfl_xor_sync(kFull, frames, o);
        }
        if ((threadIdx.x & 31) == 0 && frames) {
            atomicAdd(counters + MVN_CNT_BIT_ERRORS, (unsigned long long)bit_errs);
            atomicAdd(counters + MVN_CNT_FRAME_ERRORS, (unsigned long long)frame_errs);
            atomicAdd(counters + MVN_CNT_BITS, (unsigned long long)bits);
            atomicAdd(counters + MVN_CNT_FRAMES, (unsigned long long)frames);
        }
        bit_errs = frame_errs = bits = frames = 0;
    }
};

// Count mismatches between this lane's 32 decided bits and its row of a staged target tile.
// target.long() truncates toward zero (metrics.py:11-12).
__device__ __forceinline__ unsigned tile_bit_errors(const float *tile_row, uint32_t bits, int valid_cols) {
    unsigned e = 0;
#pragma unroll 8
    for (int c = 0; c < 32; c++) {
        const int tgt = int(tile_row[c]);
        const int bit = int((bits >> c) & 1u);
        e += (c < valid_cols && tgt != bit) ? 1u : 0u;
    }
    return e;
}

// same straight from global memory (kernels that stage no target tile): loads beyond the row are not issued
__device__ __forceinline__ unsigned row_bit_errors_global(const float *__restrict__ row, uint32_t bits, int valid_cols) {
    unsigned e = 0;
#pragma unroll 8
    for (int c = 0; c < 32; c++) {
        if (c < valid_cols) e += (int(__ldg(row + c)) != int((bits >> c) & 1u)) ? 1u : 0u;
    }
    return e;
}

}  // namespace mvn
