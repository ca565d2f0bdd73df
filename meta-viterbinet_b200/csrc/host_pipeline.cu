// End-to-end entry points with HOST buffers: a chunk pipeline
//   H2D(y chunk i+1)  ||  fused decode(chunk i)  ||  D2H(bits chunk i-1)
// over a ring of private streams, so PCIe copies in both directions overlap the kernel.  This is the call bench.py
// times for `e2e`; it is what a caller holding numpy/CPU tensors (the reference's dataset output,
// channel_dataset.py:97-104) would use.  Also here: the evaluation forms that bring back only the 32 bytes of error
// counters (targets uploaded next to y, or the whole Monte-Carlo point generated on the device), pinned-buffer
// allocation for callers, and the raw copy-rate measurement the e2e number is compared with.
#include <algorithm>
#include <chrono>
#include <cstring>
#include <new>

#include "mvn_common.cuh"
#include "../../include/mvn_b200_next.h"

struct mvn_ctx {
    int device = 0;
    int64_t chunk = 0;
    int T_max = 0;
    int L = 0;
    int S = 0;
    int variant = MVN_VARIANT_AUTO;
    int decision = MVN_DECIDE_REFERENCE;
    static constexpr int kSlots = 6;   // H2D of chunk i+1, kernel of chunk i and D2H of chunk i-1 in flight together (+ slack: more chunks in flight keep both copy engines busy)
    cudaStream_t st[kSlots] = {};
    float *d_y[kSlots] = {};
    void *d_out[kSlots] = {};   // decoded words, or the targets of the evaluation forms
    static constexpr int kWSlots = 8;   // weight sets in flight (one per SNR point of a sweep): a ring, see set_vnet_weights
    float *d_w = nullptr;       // kWSlots x packed w1,b1,w2,b2,w3,b3
    int w_cur = -1;             // slot the next decode call reads
    cudaStream_t st_w = nullptr;         // weight uploads (never queued behind a chunk)
    cudaEvent_t w_ready[kWSlots] = {};   // upload of the slot finished
    bool w_inflight[kWSlots] = {};       // launches enqueued since the last drain read the slot
    int next_slot = 0;          // stream-ring position carried across asynchronous calls
    bool pending = false;       // asynchronous work enqueued since the last mvn_ctx_synchronize
    float *d_sp = nullptr;      // state priors table (VA), up to sp_cap floats
    int64_t sp_cap = 0;
    double *d_taps = nullptr;   // taps of the on-device Monte-Carlo source
    int64_t taps_cap = 0;
    unsigned long long *d_cnt = nullptr;  // kSlots x 4 counters (one row per stream: no atomics across streams needed)
    bool have_w = false;
};

namespace mvn {
int vnet_frames_per_wave(int L, int variant);
int tc_timeout_seen();
}
using namespace mvn;

static int param_count(int S) { return kH1 + kH1 + kH2 * kH1 + kH2 + S * kH2 + S; }

extern "C" int mvn_ctx_create(mvn_ctx **out, int device, int64_t chunk_frames, int T_max, int L) {
    if (!out || T_max < 1 || L < 1 || L > 8) {
        set_error("mvn_ctx_create: bad argument");
        return MVN_ERR_ARG;
    }
    if (chunk_frames <= 0) {  // auto: two full waves of the fused kernel per chunk
        cudaError_t e0 = cudaSetDevice(device);
        if (e0 != cudaSuccess) return cuda_fail(e0, "cudaSetDevice");
        chunk_frames = 2 * int64_t(vnet_frames_per_wave(L, MVN_VARIANT_AUTO));
    }
    mvn_ctx *c = new (std::nothrow) mvn_ctx();
    if (!c) {
        set_error("mvn_ctx_create: out of host memory");
        return MVN_ERR_ARG;
    }
    c->device = device;
    c->chunk = chunk_frames;
    c->T_max = T_max;
    c->L = L;
    c->S = 1 << L;
    cudaError_t e = cudaSetDevice(device);
    for (int i = 0; i < mvn_ctx::kSlots && e == cudaSuccess; i++) {
        e = cudaStreamCreateWithFlags(&c->st[i], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaMalloc(&c->d_y[i], size_t(chunk_frames) * T_max * sizeof(float));
        if (e == cudaSuccess) e = cudaMalloc(&c->d_out[i], size_t(chunk_frames) * T_max * sizeof(float));
    }
    if (e == cudaSuccess) e = cudaMalloc(&c->d_w, size_t(mvn_ctx::kWSlots) * param_count(c->S) * sizeof(float));
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->st_w, cudaStreamNonBlocking);
    for (int i = 0; i < mvn_ctx::kWSlots && e == cudaSuccess; i++) e = cudaEventCreateWithFlags(&c->w_ready[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMalloc(&c->d_cnt, sizeof(unsigned long long) * 4 * mvn_ctx::kSlots);
    if (e != cudaSuccess) {
        mvn_ctx_destroy(c);
        return cuda_fail(e, "mvn_ctx_create");
    }
    *out = c;
    return MVN_OK;
}

extern "C" void mvn_ctx_destroy(mvn_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    for (int i = 0; i < mvn_ctx::kSlots; i++) {
        if (c->st[i]) cudaStreamSynchronize(c->st[i]);
        if (c->d_y[i]) cudaFree(c->d_y[i]);
        if (c->d_out[i]) cudaFree(c->d_out[i]);
        if (c->st[i]) cudaStreamDestroy(c->st[i]);
    }
    if (c->d_w) cudaFree(c->d_w);
    for (int i = 0; i < mvn_ctx::kWSlots; i++)
        if (c->w_ready[i]) cudaEventDestroy(c->w_ready[i]);
    if (c->st_w) {
        cudaStreamSynchronize(c->st_w);
        cudaStreamDestroy(c->st_w);
    }
    if (c->d_sp) cudaFree(c->d_sp);
    if (c->d_taps) cudaFree(c->d_taps);
    if (c->d_cnt) cudaFree(c->d_cnt);
    delete c;
}

extern "C" int mvn_ctx_set_variant(mvn_ctx *c, int variant) {
    if (!c || variant < MVN_VARIANT_AUTO || variant > MVN_VARIANT_FMA) {
        set_error("mvn_ctx_set_variant: bad argument");
        return MVN_ERR_ARG;
    }
    c->variant = variant;
    return MVN_OK;
}

extern "C" int mvn_ctx_set_decision(mvn_ctx *c, int decision) {
    if (!c || decision < MVN_DECIDE_REFERENCE || decision > MVN_DECIDE_MLSE_TERMINATED) {
        set_error("mvn_ctx_set_decision: bad argument");
        return MVN_ERR_ARG;
    }
    c->decision = decision;
    return MVN_OK;
}

static int drain(mvn_ctx *c, int rc);

// The weights go to the NEXT slot of a ring, so that decode calls already enqueued (mvn_ctx_vnet_decode_host_async) keep
// reading theirs: a sweep can stream point after point with different checkpoints without draining the pipeline in
// between.  A slot that enqueued launches may still read is never overwritten: reaching it again (8 weight sets
// without a synchronize in between) drains the pipeline first.
extern "C" int mvn_ctx_set_vnet_weights_host(mvn_ctx *c, const float *w1, const float *b1, const float *w2,
                                             const float *b2, const float *w3, const float *b3) {
    if (!c || !w1 || !b1 || !w2 || !b2 || !w3 || !b3) {
        set_error("mvn_ctx_set_vnet_weights_host: bad argument");
        return MVN_ERR_ARG;
    }
    MVN_CUDA(cudaSetDevice(c->device));
    const int slot = (c->w_cur + 1) % mvn_ctx::kWSlots;
    if (c->w_inflight[slot]) {
        const int rc = drain(c, MVN_OK);
        if (rc != MVN_OK) return rc;
    }
    const float *src[6] = {w1, b1, w2, b2, w3, b3};
    const int n[6] = {kH1, kH1, kH2 * kH1, kH2, c->S * kH2, c->S};
    float *dst = c->d_w + size_t(slot) * param_count(c->S);
    // pageable sources are staged by the runtime before the call returns; pinned ones must stay valid until the next sync
    cudaStream_t st = c->st_w;
    for (int i = 0; i < 6; i++) {
        MVN_CUDA(cudaMemcpyAsync(dst, src[i], size_t(n[i]) * sizeof(float), cudaMemcpyHostToDevice, st));
        dst += n[i];
    }
    MVN_CUDA(cudaEventRecord(c->w_ready[slot], st));
    c->w_cur = slot;
    c->have_w = true;
    return MVN_OK;
}

// Drains every stream of the context (also on the error paths: earlier chunks may still be copying into the caller's
// buffers) and reports the tcgen05 watchdog, so that a call whose pipeline timed out does not return MVN_OK.
static int drain(mvn_ctx *c, int rc) {
    cudaError_t first = cudaSuccess;
    for (int i = 0; i < mvn_ctx::kWSlots; i++) c->w_inflight[i] = false;
    for (int i = 0; i < mvn_ctx::kSlots; i++) {
        const cudaError_t e = cudaStreamSynchronize(c->st[i]);
        if (e != cudaSuccess && first == cudaSuccess) first = e;
    }
    if (rc != MVN_OK) return rc;
    if (first != cudaSuccess) return cuda_fail(first, "cudaStreamSynchronize (host pipeline)");
    if (tc_timeout_seen()) {
        set_error("host decode: the tcgen05 pipeline watchdog fired during this call; the output is invalid "
                  "(mvn_reset_tc_timeout() re-arms the kernel)");
        return MVN_ERR_CUDA;
    }
    return MVN_OK;
}

#define MVN_PIPE(call)                                                    \
    do {                                                                  \
        cudaError_t e__ = (call);                                         \
        if (e__ != cudaSuccess) return drain(c, cuda_fail(e__, #call));   \
    } while (0)

// y_host [B,T] -> chunks -> launch(...) -> optional decoded_host.  aux_host: a second [B,aux_T] fp32 input uploaded next
// to y into the slot's d_out buffer (the targets of the evaluation form; then nothing is copied back per chunk).
template <class Launch>
static int run_pipeline(mvn_ctx *c, const float *y_host, const float *aux_host, int aux_T, int64_t B, int T, int out_format,
                        void *decoded_host, bool drain_at_end, Launch launch) {
    if (!c || !y_host || B < 0 || T < 1 || T > c->T_max) {
        set_error("host decode: bad argument (T=%d, T_max=%d)", T, c ? c->T_max : -1);
        return MVN_ERR_ARG;
    }
    MVN_CUDA(cudaSetDevice(c->device));
    const int n_words = (T + 31) / 32;
    const size_t out_row = out_format == MVN_OUT_F32 ? size_t(T) * sizeof(float) : size_t(n_words) * sizeof(uint32_t);
    int slot = c->next_slot;
    c->pending = true;
    for (int64_t b0 = 0; b0 < B; b0 += c->chunk, slot = (slot + 1) % mvn_ctx::kSlots) {
        const int64_t nb = std::min<int64_t>(c->chunk, B - b0);
        cudaStream_t st = c->st[slot];
        MVN_PIPE(cudaMemcpyAsync(c->d_y[slot], y_host + b0 * T, size_t(nb) * T * sizeof(float), cudaMemcpyHostToDevice, st));
        if (aux_host)
            MVN_PIPE(cudaMemcpyAsync(c->d_out[slot], aux_host + b0 * aux_T, size_t(nb) * aux_T * sizeof(float),
                                     cudaMemcpyHostToDevice, st));
        const int rc = launch(c->d_y[slot], nb, c->d_out[slot], b0, slot, st);
        if (rc != MVN_OK) return drain(c, rc);
        if (decoded_host && !aux_host)
            MVN_PIPE(cudaMemcpyAsync(static_cast<char *>(decoded_host) + size_t(b0) * out_row, c->d_out[slot],
                                     size_t(nb) * out_row, cudaMemcpyDeviceToHost, st));
    }
    c->next_slot = slot;
    return drain_at_end ? drain(c, MVN_OK) : MVN_OK;
}

// every stream of the ring waits for the current weight slot's upload; afterwards the slot's last-use event is the join
// of all ring streams (recorded on st[0] after it has waited for the others)
static int weights_begin(mvn_ctx *c) {
    for (int i = 0; i < mvn_ctx::kSlots; i++) MVN_CUDA(cudaStreamWaitEvent(c->st[i], c->w_ready[c->w_cur], 0));
    return MVN_OK;
}
static int weights_end(mvn_ctx *c) {
    c->w_inflight[c->w_cur] = true;
    return MVN_OK;
}

static int vnet_decode_host_impl(mvn_ctx *c, const float *y_host, int64_t B, int T, int n_stages, int out_format,
                                 void *decoded_host, bool drain_at_end) {
    if (!c || !c->have_w || !decoded_host) {
        set_error("mvn_ctx_vnet_decode_host: weights not set or no output buffer");
        return MVN_ERR_ARG;
    }
    const int S = c->S;
    const float *w1 = c->d_w + size_t(c->w_cur) * param_count(S), *b1 = w1 + kH1, *w2 = b1 + kH1, *b2 = w2 + kH2 * kH1,
                *w3 = b2 + kH2, *b3 = w3 + S * kH2;
    const int L = c->L, variant = c->variant, decision = c->decision;
    MVN_CUDA(cudaSetDevice(c->device));
    int rc = weights_begin(c);
    if (rc != MVN_OK) return rc;
    rc = run_pipeline(c, y_host, nullptr, 0, B, T, out_format, decoded_host, false,
                      [=](const float *dy, int64_t nb, void *dout, int64_t, int, cudaStream_t st) {
                          return mvn_vnet_decode_ex(dy, nb, T, L, n_stages, w1, b1, w2, b2, w3, b3, out_format, dout,
                                                    nullptr, nullptr, 0, 0, nullptr, variant, decision, st);
                      });
    if (rc != MVN_OK) return rc;
    rc = weights_end(c);
    if (rc != MVN_OK) return drain(c, rc);
    return drain_at_end ? drain(c, MVN_OK) : MVN_OK;
}

extern "C" int mvn_ctx_vnet_decode_host(mvn_ctx *c, const float *y_host, int64_t B, int T, int n_stages,
                                        int out_format, void *decoded_host) {
    return vnet_decode_host_impl(c, y_host, B, T, n_stages, out_format, decoded_host, true);
}

extern "C" int mvn_ctx_vnet_decode_host_async(mvn_ctx *c, const float *y_host, int64_t B, int T, int n_stages,
                                              int out_format, void *decoded_host) {
    return vnet_decode_host_impl(c, y_host, B, T, n_stages, out_format, decoded_host, false);
}

extern "C" int mvn_ctx_synchronize(mvn_ctx *c) {
    if (!c) {
        set_error("mvn_ctx_synchronize: bad argument");
        return MVN_ERR_ARG;
    }
    MVN_CUDA(cudaSetDevice(c->device));
    c->pending = false;
    return drain(c, MVN_OK);
}

// sum the per-stream counter rows on the host: 6 x 32 bytes, one copy
static int fetch_counters(mvn_ctx *c, uint64_t *counters_host) {
    unsigned long long rows[mvn_ctx::kSlots * 4];
    MVN_CUDA(cudaMemcpy(rows, c->d_cnt, sizeof(rows), cudaMemcpyDeviceToHost));
    for (int k = 0; k < 4; k++) {
        uint64_t s = 0;
        for (int i = 0; i < mvn_ctx::kSlots; i++) s += rows[i * 4 + k];
        counters_host[k] = s;
    }
    return MVN_OK;
}

extern "C" int mvn_ctx_vnet_eval_host(mvn_ctx *c, const float *y_host, const float *target_host, int64_t B, int T,
                                      int n_stages, int target_T, int pilot_period, int out_format, void *decoded_host,
                                      uint64_t *counters_host) {
    if (!c || !c->have_w || !target_host || !counters_host || target_T < 1 || target_T > T) {
        set_error("mvn_ctx_vnet_eval_host: weights not set, or bad target / counters argument");
        return MVN_ERR_ARG;
    }
    if (decoded_host) {
        set_error("mvn_ctx_vnet_eval_host: decoded_host must be NULL (the slot's output buffer holds the targets); "
                  "use mvn_ctx_vnet_decode_host for the words");
        return MVN_ERR_UNSUPPORTED;
    }
    if (pilot_period > 0 && c->chunk % pilot_period != 0) {
        set_error("mvn_ctx_vnet_eval_host: the context's chunk (%lld frames) must be a multiple of pilot_period %d",
                  (long long)c->chunk, pilot_period);
        return MVN_ERR_ARG;
    }
    MVN_CUDA(cudaSetDevice(c->device));
    MVN_CUDA(cudaMemsetAsync(c->d_cnt, 0, sizeof(unsigned long long) * 4 * mvn_ctx::kSlots, c->st[0]));
    MVN_CUDA(cudaStreamSynchronize(c->st[0]));
    const int S = c->S;
    const float *w1 = c->d_w + size_t(c->w_cur) * param_count(S), *b1 = w1 + kH1, *w2 = b1 + kH1, *b2 = w2 + kH2 * kH1,
                *w3 = b2 + kH2, *b3 = w3 + S * kH2;
    const int L = c->L, variant = c->variant, decision = c->decision;
    unsigned long long *cnt = c->d_cnt;
    {
        const int rcw = weights_begin(c);
        if (rcw != MVN_OK) return rcw;
    }
    const int rc = run_pipeline(c, y_host, target_host, target_T, B, T, out_format, nullptr, true,
                                [=](const float *dy, int64_t nb, void *dtgt, int64_t, int slot, cudaStream_t st) {
                                    return mvn_vnet_decode_ex(dy, nb, T, L, n_stages, w1, b1, w2, b2, w3, b3, out_format,
                                                              nullptr, nullptr, static_cast<const float *>(dtgt), target_T,
                                                              pilot_period, reinterpret_cast<uint64_t *>(cnt + 4 * slot),
                                                              variant, decision, st);
                                });
    if (rc != MVN_OK) return rc;
    return fetch_counters(c, counters_host);
}

extern "C" int mvn_ctx_vnet_sweep_point(mvn_ctx *c, int64_t B, int T, int n_stages, const double *taps_host, int n_h,
                                        double snr_db, uint64_t seed, int pilot_period, uint64_t *counters_host) {
    if (!c || !c->have_w || !taps_host || n_h < 1 || !counters_host || B < 0 || T < 1 || T > c->T_max) {
        set_error("mvn_ctx_vnet_sweep_point: bad argument");
        return MVN_ERR_ARG;
    }
    if ((pilot_period > 0 && c->chunk % pilot_period != 0) || (n_h > 1 && c->chunk % n_h != 0)) {
        set_error("mvn_ctx_vnet_sweep_point: the context's chunk (%lld frames) must be a multiple of pilot_period and n_h",
                  (long long)c->chunk);
        return MVN_ERR_ARG;
    }
    MVN_CUDA(cudaSetDevice(c->device));
    const int64_t need = int64_t(n_h) * c->L;
    if (need > c->taps_cap) {
        for (int i = 0; i < mvn_ctx::kSlots; i++) MVN_CUDA(cudaStreamSynchronize(c->st[i]));
        if (c->d_taps) cudaFree(c->d_taps);
        c->d_taps = nullptr;
        c->taps_cap = 0;
        MVN_CUDA(cudaMalloc(&c->d_taps, size_t(need) * sizeof(double)));
        c->taps_cap = need;
    }
    MVN_CUDA(cudaMemcpyAsync(c->d_taps, taps_host, size_t(need) * sizeof(double), cudaMemcpyHostToDevice, c->st[0]));
    MVN_CUDA(cudaMemsetAsync(c->d_cnt, 0, sizeof(unsigned long long) * 4 * mvn_ctx::kSlots, c->st[0]));
    MVN_CUDA(cudaStreamSynchronize(c->st[0]));
    const int S = c->S, L = c->L;
    const float *w1 = c->d_w + size_t(c->w_cur) * param_count(S), *b1 = w1 + kH1, *w2 = b1 + kH1, *b2 = w2 + kH2 * kH1,
                *w3 = b2 + kH2, *b3 = w3 + S * kH2;
    {
        const int rcw = weights_begin(c);
        if (rcw != MVN_OK) return rcw;
    }
    int slot = 0;
    for (int64_t b0 = 0; b0 < B; b0 += c->chunk, slot = (slot + 1) % mvn_ctx::kSlots) {
        const int64_t nb = std::min<int64_t>(c->chunk, B - b0);
        cudaStream_t st = c->st[slot];
        float *bits = static_cast<float *>(c->d_out[slot]);
        // chunk-local seeds: every chunk draws from its own Philox subsequences (seed, first frame of the chunk)
        int rc = mvn_random_bits(bits, nb, T, seed * 0x9E3779B97F4A7C15ull + uint64_t(b0), st);
        if (rc == MVN_OK)
            rc = mvn_channel_transmit(bits, nb, T, L, c->d_taps, n_h, snr_db, nullptr, seed ^ (uint64_t(b0) << 20), c->d_y[slot], st);
        if (rc == MVN_OK)
            rc = mvn_vnet_decode_ex(c->d_y[slot], nb, T, L, n_stages, w1, b1, w2, b2, w3, b3, MVN_OUT_BITS, nullptr, nullptr,
                                    bits, T, pilot_period, reinterpret_cast<uint64_t *>(c->d_cnt + 4 * slot), c->variant,
                                    c->decision, st);
        if (rc != MVN_OK) return drain(c, rc);
    }
    const int rc = drain(c, MVN_OK);
    if (rc != MVN_OK) return rc;
    return fetch_counters(c, counters_host);
}

extern "C" int mvn_ctx_va_decode_host(mvn_ctx *c, const float *y_host, int64_t B, int T, int n_stages,
                                      const float *sp_host, int n_h, int out_format, void *decoded_host) {
    if (!c || !sp_host || n_h < 1 || !decoded_host) {
        set_error("mvn_ctx_va_decode_host: bad argument");
        return MVN_ERR_ARG;
    }
    if (B % n_h != 0 || (n_h > 1 && c->chunk % n_h != 0)) {
        set_error("mvn_ctx_va_decode_host: batch and chunk must be multiples of the %d tap blocks", n_h);
        return MVN_ERR_ARG;
    }
    MVN_CUDA(cudaSetDevice(c->device));
    const int64_t need = int64_t(n_h) * c->S;
    if (need > c->sp_cap) {
        if (c->d_sp) cudaFree(c->d_sp);
        c->d_sp = nullptr;
        c->sp_cap = 0;
        MVN_CUDA(cudaMalloc(&c->d_sp, size_t(need) * sizeof(float)));
        c->sp_cap = need;
    }
    // the table is read by kernels on every stream of the ring: upload it on one of them and wait for it before
    // anything is launched (the streams are non-blocking, the legacy stream would not order against them)
    MVN_CUDA(cudaMemcpyAsync(c->d_sp, sp_host, size_t(need) * sizeof(float), cudaMemcpyHostToDevice, c->st[0]));
    MVN_CUDA(cudaStreamSynchronize(c->st[0]));
    const float *dsp = c->d_sp;
    const int L = c->L, decision = c->decision;
    return run_pipeline(c, y_host, nullptr, 0, B, T, out_format, decoded_host, true,
                        [=](const float *dy, int64_t nb, void *dout, int64_t, int, cudaStream_t st) {
                            return mvn_va_decode_ex(dy, nb, T, L, n_stages, dsp, n_h, out_format, dout, nullptr, 0, 0,
                                                    nullptr, decision, st);
                        });
}

// ---------------------------------------------------------------- pinned buffers and the raw copy ceiling
extern "C" int mvn_host_alloc(void **ptr, size_t bytes, int write_combined) {
    if (!ptr || bytes == 0) {
        set_error("mvn_host_alloc: bad argument");
        return MVN_ERR_ARG;
    }
    MVN_CUDA(cudaHostAlloc(ptr, bytes, cudaHostAllocPortable | (write_combined ? cudaHostAllocWriteCombined : 0)));
    return MVN_OK;
}

extern "C" int mvn_host_free(void *ptr) {
    if (ptr) MVN_CUDA(cudaFreeHost(ptr));
    return MVN_OK;
}

extern "C" int mvn_copy_ceiling(int device, void *host_in, void *host_out, size_t bytes, size_t chunk_bytes, int h2d, int d2h,
                                int reps, double *seconds) {
    if ((!h2d && !d2h) || (h2d && !host_in) || (d2h && !host_out) || bytes == 0 || reps < 1 || !seconds) {
        set_error("mvn_copy_ceiling: bad argument");
        return MVN_ERR_ARG;
    }
    if (chunk_bytes == 0 || chunk_bytes > bytes) chunk_bytes = bytes;
    MVN_CUDA(cudaSetDevice(device));
    constexpr int kSlots = mvn_ctx::kSlots;
    cudaStream_t st[kSlots] = {};
    char *d_in[kSlots] = {}, *d_out[kSlots] = {};
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < kSlots && e == cudaSuccess; i++) {
        e = cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaMalloc(&d_in[i], chunk_bytes);
        if (e == cudaSuccess) e = cudaMalloc(&d_out[i], chunk_bytes);
        if (e == cudaSuccess) e = cudaMemsetAsync(d_out[i], 0, chunk_bytes, st[i]);
    }
    auto pass = [&]() {
        int slot = 0;
        for (size_t o = 0; o < bytes && e == cudaSuccess; o += chunk_bytes, slot = (slot + 1) % kSlots) {
            const size_t n = std::min(chunk_bytes, bytes - o);
            if (h2d) e = cudaMemcpyAsync(d_in[slot], static_cast<char *>(host_in) + o, n, cudaMemcpyHostToDevice, st[slot]);
            if (d2h && e == cudaSuccess)
                e = cudaMemcpyAsync(static_cast<char *>(host_out) + o, d_out[slot], n, cudaMemcpyDeviceToHost, st[slot]);
        }
        for (int i = 0; i < kSlots; i++) {
            const cudaError_t s = cudaStreamSynchronize(st[i]);
            if (s != cudaSuccess && e == cudaSuccess) e = s;
        }
    };
    if (e == cudaSuccess) pass();  // warm-up
    const auto t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < reps && e == cudaSuccess; r++) pass();
    *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    for (int i = 0; i < kSlots; i++) {
        if (d_in[i]) cudaFree(d_in[i]);
        if (d_out[i]) cudaFree(d_out[i]);
        if (st[i]) cudaStreamDestroy(st[i]);
    }
    if (e != cudaSuccess) return cuda_fail(e, "mvn_copy_ceiling");
    return MVN_OK;
}
