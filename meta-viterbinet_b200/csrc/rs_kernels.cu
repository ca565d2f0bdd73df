// f2 of SURVEY.md §8(f): Reed-Solomon over GF(2^8) (primitive polynomial 0x11d, generator roots 2^0 .. 2^(nsym-1)),
// the codec that sits between detection and BER in every coded evaluation of the reference
// (trainers/trainer.py:234-236, :286-315; ecc/rs_main.py:9-37, rs_decoder.py:37-218, rs_encoder.py:7-37).
//
// One WARP per word (<= 255 bytes), words are independent: a grid of warps strides over the batch.
//   bits -> bytes          lane j packs byte j, j+32, ... (most significant bit first, numpy packbits)
//   syndromes              S_i = sum_j word[j] 2^(i (n-1-j)): lanes own bytes, one butterfly XOR per syndrome
//   all S_i = 0            (the common case at working SNR) -> the message bits are copied, status 0
//   Berlekamp-Massey       lane 0, the reference's update rule incl. its list-LENGTH test (rs_decoder.py:150-203)
//   locator roots          lanes own candidate positions i < n (ballot keeps the reference's order)
//   Forney                 lane d owns coefficient d of the evaluator, lane i the magnitude of error i
// Integer work: results are bit-exact with the reference, including what it returns when a word has more errors than
// the code repairs (unchanged message when the locator is too long, partial correction when roots are missing).
#include <algorithm>

#include "mvn_common.cuh"
#include "../../include/mvn_b200_next.h"

namespace mvn {

constexpr int kRsWarps = 4;
constexpr int kRsMaxSym = 32;          // parity bytes
constexpr int kRsPoly = kRsMaxSym + 4; // locator / evaluator scratch length

struct GfTables {
    uint8_t exp[512];
    uint8_t log[256];
};
constexpr GfTables make_gf_tables() {
    GfTables t{};
    int x = 1;
    for (int i = 0; i < 255; i++) {
        t.exp[i] = uint8_t(x);
        t.log[x] = uint8_t(i);
        x <<= 1;
        if (x & 0x100) x ^= 0x11d;
    }
    for (int i = 255; i < 512; i++) t.exp[i] = t.exp[i - 255];
    return t;
}
__constant__ GfTables c_gf = make_gf_tables();

struct Gf {  // tables staged in shared memory: the lookups are data dependent (per-lane addresses)
    const uint8_t *exp, *log;
    __device__ __forceinline__ int mul(int a, int b) const { return (a && b) ? exp[log[a] + log[b]] : 0; }
    __device__ __forceinline__ int inv(int a) const { return exp[255 - log[a]]; }
    __device__ __forceinline__ int alpha(int k) const { return exp[k % 255]; }
};

struct RsScratch {
    uint8_t word[256];
    uint8_t synd[kRsMaxSym + 4];
    uint8_t cur[kRsPoly + 4], old[kRsPoly + 4];   // Berlekamp-Massey polynomials, index = degree
    uint8_t loc[kRsPoly + 4], omega[kRsPoly + 4];
    uint8_t pos[kRsMaxSym], xs[kRsMaxSym];
    int cur_len;  // -1: "too many errors"
};

__device__ __forceinline__ void stage_tables(uint8_t *s_exp, uint8_t *s_log) {
    for (int i = threadIdx.x; i < 512; i += blockDim.x) s_exp[i] = c_gf.exp[i];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_log[i] = c_gf.log[i];
    __syncthreads();
}

// byte j of row `src` (8 floats, nonzero = 1, most significant bit first)
__device__ __forceinline__ int pack_byte(const float *src, int j, bool vec) {
    float v[8];
    if (vec) {
        const float4 a = *reinterpret_cast<const float4 *>(src + 8 * j), b = *reinterpret_cast<const float4 *>(src + 8 * j + 4);
        v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
    } else {
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = src[8 * j + i];
    }
    int byte = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) byte |= (v[i] != 0.f ? 1 : 0) << (7 - i);
    return byte;
}
__device__ __forceinline__ void unpack_byte(float *dst, int j, int byte, bool vec) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = float((byte >> (7 - i)) & 1);
    if (vec) {
        *reinterpret_cast<float4 *>(dst + 8 * j) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4 *>(dst + 8 * j + 4) = make_float4(v[4], v[5], v[6], v[7]);
    } else {
#pragma unroll
        for (int i = 0; i < 8; i++) dst[8 * j + i] = v[i];
    }
}
__device__ __forceinline__ int warp_xor(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v ^= __shfl_xor_sync(kFull, v, o);
    return v;
}

// ---------------------------------------------------------------------------------------------
// decode: rs_main.py:21-37
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32 * kRsWarps) rs_decode_kernel(const float *__restrict__ in, int64_t B, int ld_in, int n,
                                                                  int nsym, float *__restrict__ out, int ld_out,
                                                                  int32_t *__restrict__ status, bool vec_in, bool vec_out) {
    __shared__ uint8_t s_exp[512], s_log[256];
    __shared__ RsScratch scratch[kRsWarps];
    stage_tables(s_exp, s_log);
    const Gf gf{s_exp, s_log};
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    RsScratch &s = scratch[warp];
    const int k = n - nsym;
    for (int64_t w = int64_t(blockIdx.x) * kRsWarps + warp; w < B; w += int64_t(gridDim.x) * kRsWarps) {
        const float *src = in + w * ld_in;
        float *dst = out + w * ld_out;
        // ---- bits -> bytes; per lane: log of its bytes and the exponent n-1-j of their positions
        int lg[8], ex[8];
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const int j = lane + 32 * q;
            int byte = 0;
            if (j < n) {
                byte = pack_byte(src, j, vec_in);
                s.word[j] = uint8_t(byte);
            }
            lg[q] = byte ? int(s_log[byte]) : -1;
            ex[q] = (n - 1 - j) % 255;
        }
        // ---- syndromes S_i = r(2^i) (rs_decoder.py:37-48).  The exponent of byte j in S_i is log(byte) + i (n-1-j) mod 255:
        // kept as a running index per byte (one add and one conditional subtract per syndrome instead of a multiply
        // and a modulo); S_0 is the plain XOR of the bytes.
        int any = 0;
        int idx[8];
#pragma unroll
        for (int q = 0; q < 8; q++) idx[q] = lg[q];
        for (int i = 0; i < nsym; i++) {
            int part = 0;
#pragma unroll
            for (int q = 0; q < 8; q++) {
                if (lg[q] >= 0) {
                    part ^= s_exp[idx[q]];
                    const int nx = idx[q] + ex[q];
                    idx[q] = nx >= 255 ? nx - 255 : nx;
                }
            }
            part = warp_xor(part);
            if (lane == 0) s.synd[i] = uint8_t(part);
            any |= part;
        }
        __syncwarp();
        int st = 0;
        if (any) {
            // ---- Berlekamp-Massey (rs_decoder.py:150-203); cur/old = err_loc/old_loc by degree, lengths = list lengths
            if (lane == 0) {
                int cur_len = 1, old_len = 1;
                for (int d = 0; d < kRsPoly; d++) s.cur[d] = s.old[d] = 0;
                s.cur[0] = s.old[0] = 1;
                for (int i = 0; i < nsym; i++) {
                    int delta = s.synd[i];
                    for (int j = 1; j < cur_len && j <= i; j++) delta ^= gf.mul(s.cur[j], s.synd[i - j]);
                    for (int d = old_len; d > 0; d--) s.old[d] = s.old[d - 1];   // old(x) <- x old(x)
                    s.old[0] = 0;
                    old_len++;
                    if (delta) {
                        if (old_len > cur_len) {
                            const int dinv = gf.inv(delta);
                            for (int d = 0; d < old_len || d < cur_len; d++) {
                                const int nw = gf.mul(s.old[d], delta), od = gf.mul(s.cur[d], dinv);
                                s.cur[d] = uint8_t(nw);
                                s.old[d] = uint8_t(od);
                            }
                            const int t = cur_len;
                            cur_len = old_len;
                            old_len = t;
                        }
                        for (int d = 0; d < old_len; d++) s.cur[d] ^= uint8_t(gf.mul(s.old[d], delta));
                        cur_len = max(cur_len, old_len);
                    }
                }
                while (cur_len > 0 && s.cur[cur_len - 1] == 0) cur_len--;
                s.cur_len = ((cur_len - 1) * 2 > nsym) ? -1 : cur_len;
            }
            __syncwarp();
            const int cur_len = s.cur_len;
            if (cur_len < 0) {
                st = 2;  // too many errors: the received message bytes are returned (rs_main.py:31-32)
            } else {
                // ---- roots of the reversed locator among 2^0 .. 2^(n-1) (rs_decoder.py:206-218), reference order
                const int deg = cur_len - 1;
                int n_pos = 0;
                for (int base = 0; base < n; base += 32) {
                    const int i = base + lane;
                    int val = 1;  // non-root for lanes past the end
                    if (i < n) {
                        val = 0;
                        for (int d = 0; d <= deg; d++) {
                            const int c = s.cur[d];
                            if (c) val ^= s_exp[(s_log[c] + i * (deg - d)) % 255];
                        }
                    }
                    const unsigned m = __ballot_sync(kFull, val == 0);
                    if (val == 0) {
                        const int slot = n_pos + __popc(m & ((1u << lane) - 1u));
                        if (slot < kRsMaxSym) {
                            s.pos[slot] = uint8_t(n - 1 - i);
                            s.xs[slot] = s_exp[i % 255];      // X = 2^(coefficient degree) (rs_decoder.py:101-104)
                        }
                    }
                    n_pos += __popc(m);
                }
                n_pos = min(n_pos, kRsMaxSym);
                __syncwarp();
                st = (n_pos == deg) ? 1 : 3;
                // ---- Forney (rs_decoder.py:88-147): errata locator from the positions found, evaluator mod x^(n_pos+1)
                if (lane == 0) {
                    for (int d = 0; d < kRsPoly; d++) s.loc[d] = 0;
                    s.loc[0] = 1;
                    for (int e = 0; e < n_pos; e++)          // loc(x) <- loc(x) (1 + X_e x)
                        for (int d = e + 1; d > 0; d--) s.loc[d] ^= uint8_t(gf.mul(s.loc[d - 1], s.xs[e]));
                }
                __syncwarp();
                if (lane <= n_pos) {                          // omega[d] = sum_{a+b=d} S'[a] loc[b], S'[a] = S[a-1], S'[0] = 0
                    int acc = 0;
                    for (int a = 1; a <= lane && a <= nsym; a++) acc ^= gf.mul(s.synd[a - 1], s.loc[lane - a]);
                    s.omega[lane] = uint8_t(acc);
                }
                __syncwarp();
                if (lane < n_pos) {
                    const int xi = s.xs[lane], xi_inv = gf.inv(xi);
                    int den = 1;
                    for (int j = 0; j < n_pos; j++)
                        if (j != lane) den = gf.mul(den, 1 ^ gf.mul(xi_inv, s.xs[j]));
                    int num = 0;
                    for (int d = n_pos; d >= 0; d--) num = gf.mul(num, xi_inv) ^ s.omega[d];
                    num = gf.mul(xi, num);
                    // den != 0: the X_j are distinct (positions are distinct and n <= 255)
                    s.word[s.pos[lane]] ^= uint8_t(gf.mul(num, gf.inv(den)));
                }
                __syncwarp();
            }
        }
        // ---- message bytes -> bits
        for (int j = lane; j < k; j += 32) unpack_byte(dst, j, s.word[j], vec_out);
        if (status && lane == 0) status[w] = st;
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// encode: rs_main.py:9-18, rs_encoder.py:7-37.  Lane j < nsym holds remainder coefficient j of the LFSR division of
// message(x) x^nsym by g(x); g is built once per CTA.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32 * kRsWarps) rs_encode_kernel(const float *__restrict__ in, int64_t B, int ld_in, int k,
                                                                  int nsym, float *__restrict__ out, int ld_out,
                                                                  bool vec_in, bool vec_out) {
    __shared__ uint8_t s_exp[512], s_log[256];
    __shared__ uint8_t gen[kRsMaxSym + 4];
    __shared__ uint8_t words[kRsWarps][256];
    stage_tables(s_exp, s_log);
    const Gf gf{s_exp, s_log};
    if (threadIdx.x == 0) {  // g(x) = prod_{i<nsym} (x + 2^i), index = degree (polynomials_manipulation.py:8-13)
        for (int d = 0; d <= nsym; d++) gen[d] = 0;
        gen[0] = 1;
        for (int i = 0; i < nsym; i++)
            for (int d = i + 1; d >= 0; d--) gen[d] = uint8_t((d > 0 ? gen[d - 1] : 0) ^ gf.mul(gen[d], gf.alpha(i)));
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g_lane = lane < nsym ? gen[lane] : 0;
    for (int64_t w = int64_t(blockIdx.x) * kRsWarps + warp; w < B; w += int64_t(gridDim.x) * kRsWarps) {
        const float *src = in + w * ld_in;
        float *dst = out + w * ld_out;
        for (int j = lane; j < k; j += 32) {
            const int byte = pack_byte(src, j, vec_in);
            words[warp][j] = uint8_t(byte);
            unpack_byte(dst, j, byte, vec_out);  // systematic: the message is copied (normalised to 0/1)
        }
        __syncwarp();
        int rem = 0;
        for (int j = 0; j < k; j++) {
            const int top = __shfl_sync(kFull, rem, nsym - 1);
            const int below = __shfl_up_sync(kFull, rem, 1);
            const int fb = words[warp][j] ^ top;
            rem = (lane == 0 ? 0 : below) ^ gf.mul(g_lane, fb);
        }
        if (lane < nsym) unpack_byte(dst, k + (nsym - 1 - lane), rem, vec_out);  // highest degree first
        __syncwarp();
    }
}

static bool rows_vec_ok(const void *p, int ld) { return (reinterpret_cast<uintptr_t>(p) % 16 == 0) && (ld % 4 == 0); }

static int rs_grid(int64_t B) {
    return int(std::max<int64_t>(1, std::min<int64_t>((B + kRsWarps - 1) / kRsWarps, int64_t(sm_count()) * 16)));
}

}  // namespace mvn

using namespace mvn;

static int rs_check(const char *fn, int64_t B, int ld_in, int ld_out, int n, int k, int nsym, const void *in, const void *out) {
    if (nsym < 1 || nsym > kRsMaxSym) {
        set_error("%s: nsym must be 1..%d parity bytes, got %d", fn, kRsMaxSym, nsym);
        return MVN_ERR_ARG;
    }
    if (k < 1 || n > 255) {
        set_error("%s: Message is too long (%d when max is 255) or empty", fn, n);
        return MVN_ERR_ARG;
    }
    if (B < 0 || (B > 0 && (!in || !out)) || ld_in < 0 || ld_out < 0) {
        set_error("%s: bad argument", fn);
        return MVN_ERR_ARG;
    }
    return MVN_OK;
}

extern "C" int mvn_rs_decode(const float *rx_bits, int64_t B, int ld_in, int n_bytes, int nsym, float *msg_bits, int ld_out,
                             int32_t *status, void *stream) {
    const int k = n_bytes - nsym;
    if (int rc = rs_check("mvn_rs_decode", B, ld_in, ld_out, n_bytes, k, nsym, rx_bits, msg_bits)) return rc;
    if (ld_in < 8 * n_bytes || ld_out < 8 * k) {
        set_error("mvn_rs_decode: rows too short (%d bits in for %d bytes, %d bits out for %d bytes)", ld_in, n_bytes, ld_out, k);
        return MVN_ERR_ARG;
    }
    if (B == 0) return MVN_OK;
    rs_decode_kernel<<<rs_grid(B), 32 * kRsWarps, 0, static_cast<cudaStream_t>(stream)>>>(
        rx_bits, B, ld_in, n_bytes, nsym, msg_bits, ld_out, status, rows_vec_ok(rx_bits, ld_in), rows_vec_ok(msg_bits, ld_out));
    note_launch();
    MVN_CUDA(cudaGetLastError());
    return MVN_OK;
}

extern "C" int mvn_rs_encode(const float *msg_bits, int64_t B, int ld_in, int k_bytes, int nsym, float *cw_bits, int ld_out,
                             void *stream) {
    const int n = k_bytes + nsym;
    if (int rc = rs_check("mvn_rs_encode", B, ld_in, ld_out, n, k_bytes, nsym, msg_bits, cw_bits)) return rc;
    if (ld_in < 8 * k_bytes || ld_out < 8 * n) {
        set_error("mvn_rs_encode: rows too short (%d bits in for %d bytes, %d bits out for %d bytes)", ld_in, k_bytes, ld_out, n);
        return MVN_ERR_ARG;
    }
    if (B == 0) return MVN_OK;
    rs_encode_kernel<<<rs_grid(B), 32 * kRsWarps, 0, static_cast<cudaStream_t>(stream)>>>(
        msg_bits, B, ld_in, k_bytes, nsym, cw_bits, ld_out, rows_vec_ok(msg_bits, ld_in), rows_vec_ok(cw_bits, ld_out));
    note_launch();
    MVN_CUDA(cudaGetLastError());
    return MVN_OK;
}
