// End-to-end entry points with HOST buffers: a double-buffered chunk pipeline
//   H2D(y chunk i+1)  ||  fused decode(chunk i)  ||  D2H(bits chunk i-1)
// on three private streams, so PCIe copies overlap the kernel.  This is the call bench.py times
// for `e2e`; it is what a caller holding numpy/CPU tensors (the reference's dataset output,
// channel_dataset.py:97-104) would use.
#include <algorithm>
#include <cstring>
#include <new>

#include "mvn_common.cuh"

struct mvn_ctx {
    int device = 0;
    int64_t chunk = 0;
    int T_max = 0;
    int L = 0;
    int S = 0;
    static constexpr int kSlots = 6;   // H2D of chunk i+1, kernel of chunk i and D2H of chunk i-1 in flight together (+ slack: more chunks in flight keep both copy engines busy)
    cudaStream_t st[kSlots] = {};
    float *d_y[kSlots] = {};
    void *d_out[kSlots] = {};
    float *d_w = nullptr;   // packed w1,b1,w2,b2,w3,b3
    float *d_sp = nullptr;  // state priors table (VA), up to sp_cap floats
    int64_t sp_cap = 0;
    bool have_w = false;
};

namespace mvn {
int vnet_frames_per_wave(int L);
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
        chunk_frames = 2 * int64_t(vnet_frames_per_wave(L));
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
    if (e == cudaSuccess) e = cudaMalloc(&c->d_w, size_t(param_count(c->S)) * sizeof(float));
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
    if (c->d_sp) cudaFree(c->d_sp);
    delete c;
}

extern "C" int mvn_ctx_set_vnet_weights_host(mvn_ctx *c, const float *w1, const float *b1, const float *w2,
                                             const float *b2, const float *w3, const float *b3) {
    if (!c || !w1 || !b1 || !w2 || !b2 || !w3 || !b3) {
        set_error("mvn_ctx_set_vnet_weights_host: bad argument");
        return MVN_ERR_ARG;
    }
    MVN_CUDA(cudaSetDevice(c->device));
    const float *src[6] = {w1, b1, w2, b2, w3, b3};
    const int n[6] = {kH1, kH1, kH2 * kH1, kH2, c->S * kH2, c->S};
    float *dst = c->d_w;
    for (int i = 0; i < 6; i++) {
        MVN_CUDA(cudaMemcpyAsync(dst, src[i], size_t(n[i]) * sizeof(float), cudaMemcpyHostToDevice, c->st[0]));
        dst += n[i];
    }
    MVN_CUDA(cudaStreamSynchronize(c->st[0]));
    c->have_w = true;
    return MVN_OK;
}

template <class Launch>
static int run_pipeline(mvn_ctx *c, const float *y_host, int64_t B, int T, int out_format, void *decoded_host,
                        Launch launch) {
    if (!c || !y_host || !decoded_host || B < 0 || T < 1 || T > c->T_max) {
        set_error("host decode: bad argument (T=%d, T_max=%d)", T, c ? c->T_max : -1);
        return MVN_ERR_ARG;
    }
    MVN_CUDA(cudaSetDevice(c->device));
    const int n_words = (T + 31) / 32;
    const size_t out_row = out_format == MVN_OUT_F32 ? size_t(T) * sizeof(float) : size_t(n_words) * sizeof(uint32_t);
    int slot = 0;
    for (int64_t b0 = 0; b0 < B; b0 += c->chunk, slot = (slot + 1) % mvn_ctx::kSlots) {
        const int64_t nb = std::min<int64_t>(c->chunk, B - b0);
        cudaStream_t st = c->st[slot];
        MVN_CUDA(cudaMemcpyAsync(c->d_y[slot], y_host + b0 * T, size_t(nb) * T * sizeof(float), cudaMemcpyHostToDevice, st));
        const int rc = launch(c->d_y[slot], nb, c->d_out[slot], b0, st);
        if (rc != MVN_OK) return rc;
        MVN_CUDA(cudaMemcpyAsync(static_cast<char *>(decoded_host) + size_t(b0) * out_row, c->d_out[slot],
                                 size_t(nb) * out_row, cudaMemcpyDeviceToHost, st));
    }
    for (int i = 0; i < mvn_ctx::kSlots; i++) MVN_CUDA(cudaStreamSynchronize(c->st[i]));
    return MVN_OK;
}

extern "C" int mvn_ctx_vnet_decode_host(mvn_ctx *c, const float *y_host, int64_t B, int T, int n_stages,
                                        int out_format, void *decoded_host) {
    if (!c || !c->have_w) {
        set_error("mvn_ctx_vnet_decode_host: weights not set");
        return MVN_ERR_ARG;
    }
    const int S = c->S;
    const float *w1 = c->d_w, *b1 = w1 + kH1, *w2 = b1 + kH1, *b2 = w2 + kH2 * kH1, *w3 = b2 + kH2, *b3 = w3 + S * kH2;
    const int L = c->L;
    return run_pipeline(c, y_host, B, T, out_format, decoded_host,
                        [=](const float *dy, int64_t nb, void *dout, int64_t, cudaStream_t st) {
                            return mvn_vnet_decode(dy, nb, T, L, n_stages, w1, b1, w2, b2, w3, b3, out_format, dout,
                                                   nullptr, nullptr, 0, 0, nullptr, st);
                        });
}

extern "C" int mvn_ctx_va_decode_host(mvn_ctx *c, const float *y_host, int64_t B, int T, int n_stages,
                                      const float *sp_host, int n_h, int out_format, void *decoded_host) {
    if (!c || !sp_host || n_h < 1) {
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
    MVN_CUDA(cudaMemcpy(c->d_sp, sp_host, size_t(need) * sizeof(float), cudaMemcpyHostToDevice));
    const float *dsp = c->d_sp;
    const int L = c->L;
    return run_pipeline(c, y_host, B, T, out_format, decoded_host,
                        [=](const float *dy, int64_t nb, void *dout, int64_t, cudaStream_t st) {
                            return mvn_va_decode(dy, nb, T, L, n_stages, dsp, n_h, out_format, dout, nullptr, 0, 0,
                                                 nullptr, st);
                        });
}
