// "Next" rows of SURVEY.md §8(f), built to the same parity bar:
//   f1  on-device channel simulator: bits -> pad L zeros -> BPSK -> ISI convolution -> AWGN
//       (reference: python_code/channel/channel_dataset.py:71,87-95, channel.py:12-35, modulator.py:12)
//   f3  true-MLSE mode: survivor traceback over the reference trellis (survivors are what
//       trellis_utils.py:30 returns and the reference detectors discard)
#include <curand_kernel.h>

#include <algorithm>

#include "mvn_common.cuh"
#include "../../include/mvn_b200_next.h"

namespace mvn {

// ---------------------------------------------------------------------------------------------
// f1.  One thread per (frame, symbol).  The noiseless part is evaluated in float64 in the reference's
// order (conv = sum_i h[L-1-i] * s[t+i], i ascending; numpy float64) so that with the SAME noise samples
// the fp32 result is bit-identical to the reference's `torch.Tensor(conv + w)`; with noise == NULL the
// samples come from Philox4x32-10 (seed, one subsequence per thread).
// ---------------------------------------------------------------------------------------------
__global__ void channel_kernel(const float *__restrict__ bits, int64_t B, int T, int L, const double *__restrict__ taps,
                               int n_h, double sigma, const double *__restrict__ noise, unsigned long long seed,
                               float *__restrict__ y) {
    const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= B * T) return;
    const int64_t b = idx / T;
    const int t = int(idx % T);
    const double *h = taps + (b % n_h) * L;
    double conv = 0.0;
    for (int i = 0; i < L; i++) {
        const int k = t + i;                                  // codeword padded with L zero bits -> symbol +1
        const double s = (k < T) ? 1.0 - 2.0 * double(bits[b * T + k]) : 1.0;
        conv += h[L - 1 - i] * s;
    }
    double w;
    if (noise) {
        w = sigma * noise[idx];
    } else {
        curandStatePhilox4_32_10_t st;
        curand_init(seed, (unsigned long long)idx, 0, &st);
        w = sigma * curand_normal_double(&st);
    }
    y[idx] = float(conv + w);
}

// Bernoulli(1/2) words for the on-device Monte-Carlo source (the role of word_rand_gen.randint(0, 2, ...),
// channel_dataset.py:67): one Philox4x32-10 call yields 128 bits = 128 consecutive symbols of the flattened [B*T] array.
__global__ void random_bits_kernel(float *__restrict__ bits, int64_t n, unsigned long long seed) {
    const int64_t g = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;   // group of 128 symbols
    if (g * 128 >= n) return;
    curandStatePhilox4_32_10_t st;
    curand_init(seed, (unsigned long long)g, 0, &st);
    const uint4 r = curand4(&st);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
    const int64_t base = g * 128;
#pragma unroll
    for (int k = 0; k < 4; k++) {
#pragma unroll 8
        for (int i = 0; i < 32; i += 4) {
            const int64_t o = base + 32 * k + i;
            const float4 v = make_float4(float((w[k] >> i) & 1u), float((w[k] >> (i + 1)) & 1u), float((w[k] >> (i + 2)) & 1u),
                                         float((w[k] >> (i + 3)) & 1u));
            if (o + 3 < n) {
                *reinterpret_cast<float4 *>(bits + o) = v;
            } else {
                if (o < n) bits[o] = v.x;
                if (o + 1 < n) bits[o + 1] = v.y;
                if (o + 2 < n) bits[o + 2] = v.z;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// f3.  Traceback.  survivors[b][t][W] are the bit-packed decisions exported by mvn_acs_decode (bit j of the
// word group = predecessor choice of new state j < S/2; states j and j+S/2 share it).  State at time t is
// sum_i b[t+i] 2^i, so the predecessor of j is (2j + sigma) mod S and sigma IS the transmitted bit b[t].
// start_state < 0: start from the best final state (lowest index on ties); otherwise from that state
// (0 for the reference's zero-padded, i.e. terminated, words).
// ---------------------------------------------------------------------------------------------
__global__ void traceback_kernel(const uint32_t *__restrict__ surv, const float *__restrict__ final_pm, int64_t B, int T,
                                 int n_stages, int L, int start_state, int out_format, void *__restrict__ decoded) {
    const int64_t b = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int S = 1 << L, H = S > 1 ? S / 2 : 1, W = (H + 31) / 32;
    int j = start_state;
    if (j < 0) {
        float best = final_pm[b * S];
        j = 0;
        for (int s = 1; s < S; s++) {
            const float v = final_pm[b * S + s];
            if (v < best) {
                best = v;
                j = s;
            }
        }
    }
    const int n_words = (T + 31) / 32;
    float *outf = out_format == MVN_OUT_F32 ? static_cast<float *>(decoded) + b * T : nullptr;
    uint32_t *outw = out_format == MVN_OUT_BITS ? static_cast<uint32_t *>(decoded) + b * n_words : nullptr;
    if (outf)
        for (int t = n_stages; t < T; t++) outf[t] = 0.f;
    uint32_t word = 0;
    for (int t = T - 1; t >= n_stages; t--)
        if (outw && (t & 31) == 0) outw[t >> 5] = 0;
    for (int t = n_stages - 1; t >= 0; t--) {
        const int jj = j % H;
        const uint32_t sigma = (surv[(b * n_stages + t) * W + (jj >> 5)] >> (jj & 31)) & 1u;
        j = (2 * j + int(sigma)) % S;
        if (outf) outf[t] = float(sigma);
        word |= sigma << (t & 31);
        if ((t & 31) == 0) {
            if (outw) outw[t >> 5] = word;
            word = 0;
        }
    }
}

// branch metrics of the full-CSI Viterbi as a tensor (only the MLSE path materialises them)
__global__ void va_cost_kernel(const float *__restrict__ y, int64_t B, int T, int S, const float *__restrict__ sp, int n_h,
                               float *__restrict__ cost) {
    const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= B * T * S) return;
    const int s = int(idx % S);
    const int64_t bt = idx / S;
    const int64_t b = bt / T;
    const float d = __fsub_rn(y[bt], sp[(b % n_h) * S + s]);
    cost[idx] = __fmaf_rn(__fmul_rn(d, d), 0.5f, -kLogSqrt2Pi);
}

}  // namespace mvn

using namespace mvn;

extern "C" int mvn_channel_transmit(const float *bits, int64_t B, int T, int L, const double *taps, int n_h, double snr_db,
                                    const double *noise, uint64_t seed, float *y, void *stream) {
    if (B < 0 || T < 1 || L < 1 || L > 16 || n_h < 1 || !taps || (B > 0 && (!bits || !y))) {
        set_error("mvn_channel_transmit: bad argument");
        return MVN_ERR_ARG;
    }
    if (B == 0) return MVN_OK;
    const double sigma = pow(pow(10.0, snr_db / 10.0), -0.5);   // channel.py:19,27
    const int64_t n = B * T;
    channel_kernel<<<unsigned((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(bits, B, T, L, taps, n_h, sigma, noise,
                                                                                             seed, y);
    note_launch();
    MVN_CUDA(cudaGetLastError());
    return MVN_OK;
}

extern "C" int mvn_random_bits(float *bits, int64_t B, int T, uint64_t seed, void *stream) {
    if (B < 0 || T < 1 || (B > 0 && !bits) || (reinterpret_cast<uintptr_t>(bits) & 15u)) {
        set_error("mvn_random_bits: bad argument (bits must be 16-byte aligned)");
        return MVN_ERR_ARG;
    }
    const int64_t n = B * T;
    if (n == 0) return MVN_OK;
    const int64_t groups = (n + 127) / 128;
    random_bits_kernel<<<unsigned((groups + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(bits, n, seed);
    note_launch();
    MVN_CUDA(cudaGetLastError());
    return MVN_OK;
}

extern "C" int mvn_traceback(const uint32_t *survivors, const float *final_pm, int64_t B, int T, int n_stages, int L,
                             int start_state, int out_format, void *decoded, void *stream) {
    if (B < 0 || T < 0 || n_stages < 0 || n_stages > T || L < 1 || L > 8 || start_state >= (1 << L) ||
        (B > 0 && (!survivors || !decoded || (start_state < 0 && !final_pm))) ||
        (out_format != MVN_OUT_F32 && out_format != MVN_OUT_BITS)) {
        set_error("mvn_traceback: bad argument");
        return MVN_ERR_ARG;
    }
    if (B == 0 || T == 0) return MVN_OK;
    traceback_kernel<<<unsigned((B + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(survivors, final_pm, B, T, n_stages,
                                                                                               L, start_state, out_format, decoded);
    note_launch();
    MVN_CUDA(cudaGetLastError());
    return MVN_OK;
}

extern "C" int mvn_va_cost(const float *y, int64_t B, int T, int L, const float *state_priors, int n_h, float *cost,
                           void *stream) {
    if (B < 0 || T < 0 || L < 1 || L > 8 || n_h < 1 || !state_priors || (B * T > 0 && (!y || !cost))) {
        set_error("mvn_va_cost: bad argument");
        return MVN_ERR_ARG;
    }
    if (B % n_h != 0) {
        set_error("mvn_va_cost: batch %lld is not a multiple of the %d tap blocks", (long long)B, n_h);
        return MVN_ERR_ARG;
    }
    const int S = 1 << L;
    const int64_t n = B * T * S;
    if (n == 0) return MVN_OK;
    va_cost_kernel<<<unsigned((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(y, B, T, S, state_priors, n_h, cost);
    note_launch();
    MVN_CUDA(cudaGetLastError());
    return MVN_OK;
}
