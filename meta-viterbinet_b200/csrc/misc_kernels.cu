// a8 (state labels), a12 (BER/FER counters), library plumbing (errors, device info, launch
// counter) and the FP32 micro-benchmark used as the measured roofline denominator.
#include <algorithm>
#include <cstdarg>
#include <cstdio>

#include "mvn_common.cuh"

namespace mvn {

static thread_local char g_err[512] = "";
static thread_local int64_t g_launches = 0;

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int cuda_fail(cudaError_t e, const char *what) {
    set_error("CUDA error %d (%s) in %s", int(e), cudaGetErrorString(e), what);
    return MVN_ERR_CUDA;
}
void note_launch() { g_launches++; }

int sm_count() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1) n = 148;
        cached = n;
        cached_dev = dev;
    }
    return cached;
}

// ---------------------------------------------------------------- a8
// trellis_utils.py:33-46: the reference forms sum_i tx[b,t+i] * 2^i in fp32 and casts to long.
__global__ void states_kernel(const float *__restrict__ tx, int64_t B, int T, int L, int64_t *__restrict__ states) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= B * T) return;
    const int64_t b = i / T;
    const int t = int(i % T);
    float s = 0.f;
    for (int k = 0; k < L; k++) {
        const float v = (t + k < T) ? tx[b * T + t + k] : 0.f;
        s = __fadd_rn(s, __fmul_rn(v, float(1 << k)));
    }
    states[i] = (long long)s;
}

// ---------------------------------------------------------------- a12
// One warp per row (grid-stride): lanes stride the row with full-line reads; integer compare
// after truncation (.long()), exact 64-bit totals.
__global__ void __launch_bounds__(256) error_count_kernel(const float *__restrict__ pred, int pred_ld,
                                                          const float *__restrict__ tgt, int tgt_ld, int64_t B, int T,
                                                          int pilot_period, unsigned long long *counters,
                                                          uint8_t *__restrict__ row_errors) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = (int64_t(gridDim.x) * blockDim.x) >> 5;
    unsigned long long bit_errs = 0, frame_errs = 0, frames = 0;
    for (int64_t b = warp0; b < B; b += n_warps) {
        if (pilot_period > 0 && b % pilot_period == 0) {
            if (row_errors && lane == 0) row_errors[b] = 0;
            continue;
        }
        unsigned e = 0;
        for (int t = lane; t < T; t += 32) {
            const long long p = (long long)pred[b * pred_ld + t];
            const long long q = (long long)tgt[b * tgt_ld + t];
            e += (p != q) ? 1u : 0u;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(kFull, e, o);
        bit_errs += e;
        frame_errs += e ? 1 : 0;
        frames += 1;
        if (row_errors && lane == 0) row_errors[b] = e ? 1 : 0;
    }
    if (lane == 0 && frames) {
        atomicAdd(counters + MVN_CNT_BIT_ERRORS, bit_errs);
        atomicAdd(counters + MVN_CNT_FRAME_ERRORS, frame_errs);
        atomicAdd(counters + MVN_CNT_BITS, frames * (unsigned long long)T);
        atomicAdd(counters + MVN_CNT_FRAMES, frames);
    }
}

// ---------------------------------------------------------------- FP32 peak micro-benchmark
// Register-only dependent chains, 16 independent accumulators (pairs) per thread.
template <int MODE>
__global__ void __launch_bounds__(256) fp32_peak_kernel(float *out, int iters, float a, float b) {
    if constexpr (MODE == 0) {
        float acc[16];
#pragma unroll
        for (int i = 0; i < 16; i++) acc[i] = float(threadIdx.x + i);
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int r = 0; r < 8; r++)
#pragma unroll
                for (int i = 0; i < 16; i++) acc[i] = fmaf(acc[i], a, b);
        }
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 16; i++) s += acc[i];
        if (s == 12345.678f) out[0] = s;
    } else {
        unsigned long long acc[8], aa, bb;
        asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
        asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const float v = float(threadIdx.x + i);
            asm("mov.b64 %0, {%1, %1};" : "=l"(acc[i]) : "f"(v));
        }
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int r = 0; r < 8; r++)
#pragma unroll
                for (int i = 0; i < 8; i++) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[i]) : "l"(aa), "l"(bb));
        }
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            float x, y;
            asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(acc[i]));
            s += x + y;
        }
        if (s == 12345.678f) out[0] = s;
    }
}

}  // namespace mvn

using namespace mvn;

extern "C" const char *mvn_last_error(void) { return g_err; }
extern "C" int mvn_version(void) { return 100; }
extern "C" int64_t mvn_launch_count(int reset) {
    const int64_t v = g_launches;
    if (reset) g_launches = 0;
    return v;
}

extern "C" int mvn_device_info(int *sms, int *sm_clock_khz, int *cc_major, int *cc_minor) {
    int dev = 0;
    MVN_CUDA(cudaGetDevice(&dev));
    int v = 0;
    if (sms) {
        MVN_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
        *sms = v;
    }
    if (sm_clock_khz) {
        MVN_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrClockRate, dev));
        *sm_clock_khz = v;
    }
    if (cc_major) {
        MVN_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev));
        *cc_major = v;
    }
    if (cc_minor) {
        MVN_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev));
        *cc_minor = v;
    }
    return MVN_OK;
}

extern "C" int mvn_calculate_states(const float *tx, int64_t B, int T, int L, int64_t *states, void *stream) {
    if (B < 0 || T < 0 || L < 1 || L > 30 || (B * T > 0 && (!tx || !states))) {
        set_error("mvn_calculate_states: bad argument");
        return MVN_ERR_ARG;
    }
    const int64_t n = B * T;
    if (n == 0) return MVN_OK;
    states_kernel<<<unsigned((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(tx, B, T, L, states);
    note_launch();
    MVN_CUDA(cudaGetLastError());
    return MVN_OK;
}

extern "C" int mvn_error_counts(const float *prediction, int pred_stride, const float *target, int target_stride,
                                int64_t B, int T, int pilot_period, uint64_t *counters, uint8_t *row_errors,
                                void *stream) {
    if (B < 0 || T < 0 || !counters || pred_stride < T || target_stride < T || (B * T > 0 && (!prediction || !target))) {
        set_error("mvn_error_counts: bad argument");
        return MVN_ERR_ARG;
    }
    if (B == 0) return MVN_OK;
    const int64_t blocks = std::min<int64_t>((B + 7) / 8, int64_t(sm_count()) * 8);
    error_count_kernel<<<unsigned(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        prediction, pred_stride, target, target_stride, B, T, pilot_period,
        reinterpret_cast<unsigned long long *>(counters), row_errors);
    note_launch();
    MVN_CUDA(cudaGetLastError());
    return MVN_OK;
}

extern "C" int mvn_fp32_peak(int mode, int iters, double *fma_per_s, double *ms_out, void *stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float *d = nullptr;
    MVN_CUDA(cudaMalloc(&d, sizeof(float)));
    cudaEvent_t e0, e1;
    MVN_CUDA(cudaEventCreate(&e0));
    MVN_CUDA(cudaEventCreate(&e1));
    const int blocks = sm_count() * 8;
    for (int rep = 0; rep < 2; rep++) {  // first pass warms up
        MVN_CUDA(cudaEventRecord(e0, st));
        if (mode == 0)
            fp32_peak_kernel<0><<<blocks, 256, 0, st>>>(d, iters, 0.999f, 0.001f);
        else
            fp32_peak_kernel<1><<<blocks, 256, 0, st>>>(d, iters, 0.999f, 0.001f);
        note_launch();
        MVN_CUDA(cudaEventRecord(e1, st));
        MVN_CUDA(cudaEventSynchronize(e1));
    }
    float ms = 0.f;
    MVN_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    const double fmas = double(blocks) * 256.0 * double(iters) * 128.0;
    if (fma_per_s) *fma_per_s = fmas / (double(ms) * 1e-3);
    if (ms_out) *ms_out = ms;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    return MVN_OK;
}
