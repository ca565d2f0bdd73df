// a9 / a10 / a11 — training of the priors net: CE loss over state labels, hand-derived backward,
// Hessian-vector product for second-order MAML, torch-compatible Adam.  One CTA per realisation.
// Reference: trainers/trainer.py:425-453 (meta_train_loop), :492-505 (run_train_loop),
// trainers/META_VNET/metavnet_trainer.py:41-50 (loss over all symbols), torch.optim.Adam defaults.
//
// Every step of the reference depends on the previous one through the weights and the Adam state,
// so the only parallelism is ACROSS independent runs (SNR points, channel realisations, seeds):
// R realisations -> R CTAs, each with its own theta / Adam state in HBM and a private scratch slab.
//
// Inside a CTA a step is two phases:
//   phase 1  one thread per symbol: forward, softmax-CE, backward to the pre-activations (and, for
//            the MAML Hessian-vector product, the forward-over-reverse tangents of all of them);
//            rows of h1, h2, dz, da2, da1 (and their tangents) go to the scratch slab;
//   phase 2  one thread per parameter: the sum over symbols of the outer products
//            (dW2[o][k] = sum_n da2[n][o] h1[n][k], ...), sequential in n => deterministic.
#include <algorithm>

#include "mvn_common.cuh"
#include "train_gemm.cuh"
#include "../../include/mvn_b200_train.h"

namespace mvn {

constexpr int kTrainThreads = 256;

__host__ __device__ constexpr int param_count_s(int S) { return kH1 + kH1 + kH2 * kH1 + kH2 + S * kH2 + S; }

template <int S>
struct ThetaView {  // offsets into the packed parameter vector (torch parameter order)
    static constexpr int w1 = 0, b1 = kH1, w2 = 2 * kH1, b2 = w2 + kH2 * kH1, w3 = b2 + kH2, b3 = w3 + S * kH2;
    static constexpr int P = b3 + S;
};

// scratch rows per symbol
template <int S>
struct Slab {
    static constexpr int SP = (S + 3) / 4 * 4;
    static constexpr int oH1 = 0, oH2 = oH1 + kH1, oDZ = oH2 + kH2 + 2, oDA2 = oDZ + SP, oDA1 = oDA2 + kH2 + 2;
    static constexpr int kHalf = (oDA1 + kH1 + 3) / 4 * 4;  // one set (values); tangents follow
    static constexpr int kRow = 2 * kHalf;
    static_assert(oDZ % 4 == 0 && oDA2 % 4 == 0 && oDA1 % 4 == 0, "rows are float4 aligned");
};

__device__ __forceinline__ float sigmoidf_acc(float a) { return 1.f / (1.f + expf(-a)); }

// ---------------------------------------------------------------------------------------------
// phase 1.  th: theta in shared memory; tv: tangent direction (TAN only).  If dz_ext != nullptr the
// upstream gradient w.r.t. the priors is taken from it (autograd backward) instead of the CE loss.
// ---------------------------------------------------------------------------------------------
template <int S, bool TAN>
__device__ __forceinline__ float symbol_pass(const float *__restrict__ th, const float *__restrict__ tv, float y,
                                             int label, const float *__restrict__ dz_ext, float inv_n,
                                             float *__restrict__ row, float *__restrict__ rz_out = nullptr) {
    using TV = ThetaView<S>;
    using SL = Slab<S>;
    float *rt = row + SL::kHalf;  // tangent half of the row
    float h1[kH1];
    float rh1[TAN ? kH1 : 1];
#pragma unroll
    for (int k = 0; k < kH1; k++) {
        const float h = sigmoidf_acc(fmaf(th[TV::w1 + k], y, th[TV::b1 + k]));
        h1[k] = h;
        if (TAN) rh1[k] = h * (1.f - h) * fmaf(tv[TV::w1 + k], y, tv[TV::b1 + k]);
    }
#pragma unroll
    for (int k = 0; k < kH1; k += 4) {
        *reinterpret_cast<float4 *>(row + SL::oH1 + k) = make_float4(h1[k], h1[k + 1], h1[k + 2], h1[k + 3]);
        if (TAN) *reinterpret_cast<float4 *>(rt + SL::oH1 + k) = make_float4(rh1[k], rh1[k + 1], rh1[k + 2], rh1[k + 3]);
    }
    float z[S], rz[TAN ? S : 1];
#pragma unroll
    for (int s = 0; s < S; s++) {
        z[s] = th[TV::b3 + s];
        if (TAN) rz[s] = tv[TV::b3 + s];
    }
#pragma unroll 1
    for (int o = 0; o < kH2; o++) {
        const float *w2 = th + TV::w2 + o * kH1;
        const float *v2 = tv + TV::w2 + o * kH1;
        float a0 = th[TV::b2 + o], a1 = 0.f, r0 = TAN ? tv[TV::b2 + o] : 0.f, r1 = 0.f;
#pragma unroll
        for (int k = 0; k < kH1; k += 2) {
            a0 = fmaf(w2[k], h1[k], a0);
            a1 = fmaf(w2[k + 1], h1[k + 1], a1);
            if (TAN) {
                r0 = fmaf(v2[k], h1[k], fmaf(w2[k], rh1[k], r0));
                r1 = fmaf(v2[k + 1], h1[k + 1], fmaf(w2[k + 1], rh1[k + 1], r1));
            }
        }
        const float a2 = a0 + a1;
        const float h2 = fmaxf(a2, 0.f);
        const float rh2 = (TAN && a2 > 0.f) ? (r0 + r1) : 0.f;
        row[SL::oH2 + o] = h2;
        if (TAN) rt[SL::oH2 + o] = rh2;
#pragma unroll
        for (int s = 0; s < S; s++) {
            const float w3 = th[TV::w3 + s * kH2 + o];
            z[s] = fmaf(w3, h2, z[s]);
            if (TAN) rz[s] = fmaf(tv[TV::w3 + s * kH2 + o], h2, fmaf(w3, rh2, rz[s]));
        }
    }
    // softmax cross-entropy (torch CrossEntropyLoss, mean reduction -> 1/N folded into dz)
    float dz[S], rdz[TAN ? S : 1];
    float loss = 0.f;
    if (dz_ext) {  // upstream gradient is an independent input: its tangent is zero
#pragma unroll
        for (int s = 0; s < S; s++) {
            dz[s] = dz_ext[s];
            if (TAN) {
                rdz[s] = 0.f;
                if (rz_out) rz_out[s] = rz[s];
            }
        }
    } else {
        float m = z[0];
#pragma unroll
        for (int s = 1; s < S; s++) m = fmaxf(m, z[s]);
        float sum = 0.f, zl = 0.f;
#pragma unroll
        for (int s = 0; s < S; s++) {
            dz[s] = expf(z[s] - m);
            sum += dz[s];
            zl = (s == label) ? z[s] : zl;
        }
        loss = logf(sum) + m - zl;
        const float inv = 1.f / sum;
        float dot = 0.f;
#pragma unroll
        for (int s = 0; s < S; s++) {
            dz[s] *= inv;  // p
            if (TAN) dot = fmaf(dz[s], rz[s], dot);
        }
#pragma unroll
        for (int s = 0; s < S; s++) {
            if (TAN) rdz[s] = dz[s] * (rz[s] - dot) * inv_n;
            dz[s] = (dz[s] - ((s == label) ? 1.f : 0.f)) * inv_n;
        }
    }
#pragma unroll
    for (int s = 0; s < S; s++) {
        row[SL::oDZ + s] = dz[s];
        if (TAN) rt[SL::oDZ + s] = rdz[s];
    }
    // backward to the hidden pre-activations
    float dh1[kH1], rdh1[TAN ? kH1 : 1];
#pragma unroll
    for (int k = 0; k < kH1; k++) {
        dh1[k] = 0.f;
        if (TAN) rdh1[k] = 0.f;
    }
#pragma unroll 1
    for (int o = 0; o < kH2; o++) {
        const bool on = row[SL::oH2 + o] > 0.f;
        float dh2 = 0.f, rdh2 = 0.f;
#pragma unroll
        for (int s = 0; s < S; s++) {
            const float w3 = th[TV::w3 + s * kH2 + o];
            dh2 = fmaf(w3, dz[s], dh2);
            if (TAN) rdh2 = fmaf(tv[TV::w3 + s * kH2 + o], dz[s], fmaf(w3, rdz[s], rdh2));
        }
        const float da2 = on ? dh2 : 0.f;
        const float rda2 = (TAN && on) ? rdh2 : 0.f;
        row[SL::oDA2 + o] = da2;
        if (TAN) rt[SL::oDA2 + o] = rda2;
        const float *w2 = th + TV::w2 + o * kH1;
        const float *v2 = tv + TV::w2 + o * kH1;
#pragma unroll
        for (int k = 0; k < kH1; k++) {
            dh1[k] = fmaf(w2[k], da2, dh1[k]);
            if (TAN) rdh1[k] = fmaf(v2[k], da2, fmaf(w2[k], rda2, rdh1[k]));
        }
    }
#pragma unroll
    for (int k = 0; k < kH1; k++) {
        const float h = row[SL::oH1 + k];
        const float s1 = h * (1.f - h);
        row[SL::oDA1 + k] = dh1[k] * s1;
        if (TAN) {
            const float ra1 = fmaf(tv[TV::w1 + k], y, tv[TV::b1 + k]);
            rt[SL::oDA1 + k] = fmaf(rdh1[k], s1, dh1[k] * s1 * (1.f - 2.f * h) * ra1);
        }
    }
    return loss;
}

// ---------------------------------------------------------------------------------------------
// phase 2: out[idx] = sum_n (outer products).  TAN: the tangent (Hessian-vector) combination.
// ---------------------------------------------------------------------------------------------
template <int S, bool TAN>
__device__ __forceinline__ float param_grad(int idx, const float *__restrict__ slab, const float *__restrict__ ys, int n) {
    using TV = ThetaView<S>;
    using SL = Slab<S>;
    constexpr int R = SL::kRow, Hf = SL::kHalf;
    float acc = 0.f;
    if (idx < TV::b1) {  // w1[k]
        const int k = idx;
        for (int i = 0; i < n; i++) acc = fmaf(slab[i * R + (TAN ? Hf : 0) + SL::oDA1 + k], ys[i], acc);
    } else if (idx < TV::w2) {  // b1[k]
        const int k = idx - TV::b1;
        for (int i = 0; i < n; i++) acc += slab[i * R + (TAN ? Hf : 0) + SL::oDA1 + k];
    } else if (idx < TV::b2) {  // w2[o][k]
        const int o = (idx - TV::w2) / kH1, k = (idx - TV::w2) % kH1;
        for (int i = 0; i < n; i++) {
            const float *r = slab + i * R;
            if (TAN)
                acc = fmaf(r[Hf + SL::oDA2 + o], r[SL::oH1 + k], fmaf(r[SL::oDA2 + o], r[Hf + SL::oH1 + k], acc));
            else
                acc = fmaf(r[SL::oDA2 + o], r[SL::oH1 + k], acc);
        }
    } else if (idx < TV::w3) {  // b2[o]
        const int o = idx - TV::b2;
        for (int i = 0; i < n; i++) acc += slab[i * R + (TAN ? Hf : 0) + SL::oDA2 + o];
    } else if (idx < TV::b3) {  // w3[s][o]
        const int s = (idx - TV::w3) / kH2, o = (idx - TV::w3) % kH2;
        for (int i = 0; i < n; i++) {
            const float *r = slab + i * R;
            if (TAN)
                acc = fmaf(r[Hf + SL::oDZ + s], r[SL::oH2 + o], fmaf(r[SL::oDZ + s], r[Hf + SL::oH2 + o], acc));
            else
                acc = fmaf(r[SL::oDZ + s], r[SL::oH2 + o], acc);
        }
    } else {  // b3[s]
        const int s = idx - TV::b3;
        for (int i = 0; i < n; i++) acc += slab[i * R + (TAN ? Hf : 0) + SL::oDZ + s];
    }
    return acc;
}

struct AdamCfg {
    float lr, beta1, beta2, eps;
};

// torch.optim.Adam (single-tensor path): m.lerp_(g, 1-b1); v = b2 v + (1-b2) g^2;
// denom = sqrt(v)/sqrt(1-b2^t) + eps; theta -= (lr / (1-b1^t)) * m / denom.
__device__ __forceinline__ void adam_update(float *theta, float *m, float *v, const float *g, int P, int step,
                                            AdamCfg c) {
    const float bc1 = float(1.0 - pow(double(c.beta1), double(step)));
    const float bc2_sqrt = float(sqrt(1.0 - pow(double(c.beta2), double(step))));
    const float step_size = c.lr / bc1;
    for (int i = threadIdx.x; i < P; i += kTrainThreads) {
        const float gi = g[i];
        const float mi = m[i] + (1.f - c.beta1) * (gi - m[i]);
        const float vi = c.beta2 * v[i] + (1.f - c.beta2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + c.eps;
        theta[i] = theta[i] - step_size * (mi / denom);
    }
}

// same update with the gradient read from the padded shared-memory layout; bc = (1 - b1^t, sqrt(1 - b2^t)) computed once
// per CTA in double like torch does on the host
template <int S>
__device__ __forceinline__ void adam_update_padded(float *__restrict__ theta, float *__restrict__ m, float *__restrict__ v,
                                                   const float *__restrict__ g, const float *__restrict__ bc, AdamCfg c) {
    using LY = tg::Lay<S>;
    const float step_size = c.lr / bc[0], bc2_sqrt = bc[1];
#pragma unroll 4
    for (int i = threadIdx.x; i < LY::TP; i += tg::kThreads) {
        const int ip = i < LY::tb2 ? i : LY::to_padded(i);
        const float gi = g[ip];
        const float mi = m[i] + (1.f - c.beta1) * (gi - m[i]);
        const float vi = c.beta2 * v[i] + (1.f - c.beta2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + c.eps;
        theta[i] = theta[i] - step_size * (mi / denom);
    }
}

struct TrainParams {
    float *theta, *adam_m, *adam_v;
    int32_t *adam_step;
    int R;
    const float *y_s;
    const int32_t *lab_s;
    int Ns;
    const float *y_q;
    const int32_t *lab_q;
    int Nq;
    float meta_lr;
    AdamCfg adam;
    int second_order;
    float *loss_out, *grad_out;
    float *workspace;
    int update;  // 0: only loss/grad
};

// ---------------------------------------------------------------------------------------------
// Batched step kernels (mvn_train_step_batched / mvn_meta_step_batched): one CTA per realisation, the GEMM-structured
// passes of train_gemm.cuh.  Shared memory: Wset | Vset (second-order MAML only) | G | activations | red.
// ---------------------------------------------------------------------------------------------
constexpr int kCapWide = 136, kLdnWide = 140;   // one word of the reference's coded block per chunk (non-tangent passes)
// tangent passes hold two activation sets: half-word chunks (a third of a word at 32 states, where the sets are larger)
template <int S> struct TanCap { static constexpr int cap = S <= 16 ? 72 : 56, ldn = cap + 4; };
constexpr int kCapPlain = 72, kLdnPlain = 76;   // plain steps: narrow layout, two CTAs per SM

template <int S, bool META>
struct StepSmem {
    using LY = tg::Lay<S>;
    // META: the wide layout for the support / query gradient passes and the two-set narrow layout for the tangent pass
    // share one region
    static constexpr size_t wide = tg::act_floats<S>(kLdnWide, false), tan = tg::act_floats<S>(TanCap<S>::ldn, true);
    static constexpr size_t acts = META ? (wide > tan ? wide : tan) : tg::act_floats<S>(kLdnPlain, false) - size_t(tg::kH2P) * kLdnPlain;
    static constexpr size_t floats = size_t(META ? 3 : 2) * LY::PP + acts + 32;
    static constexpr size_t bytes = floats * sizeof(float);
    static_assert(bytes <= 227 * 1024, "step kernel: shared memory of one CTA");
};

template <int S>
__device__ __forceinline__ tg::Smem<S> carve(float *sm, bool with_v, int ldn, bool separate_da2 = true) {
    using LY = tg::Lay<S>;
    tg::Smem<S> v;
    v.W = sm;
    v.V = with_v ? sm + LY::PP : sm;
    v.G = sm + (with_v ? 2 : 1) * LY::PP;
    float *a = v.G + LY::PP;
    v.red = a;
    a += 32;
    v.ldn = ldn;
    v.y = a;
    v.H1T = v.y + ldn;
    v.H2T = v.H1T + kH1 * ldn;
    v.ZT = v.H2T + tg::kH2P * ldn;
    v.DA2T = separate_da2 ? v.ZT + LY::SP * ldn : v.H2T;
    v.RH1T = v.ZT + LY::SP * ldn + (separate_da2 ? tg::kH2P * ldn : 0);
    v.RH2T = v.RH1T + kH1 * ldn;
    v.RZT = v.RH2T + tg::kH2P * ldn;
    v.RDA2T = v.RZT + LY::SP * ldn;
    return v;
}

// theta (torch packing, HBM) -> padded layout in shared memory (padding zeroed)
template <int S>
__device__ __forceinline__ void load_theta(float *dst, const float *__restrict__ theta) {
    using LY = tg::Lay<S>;
    const int tid = threadIdx.x;
    for (int i = tid; i < LY::tb2; i += tg::kThreads) dst[i] = theta[i];                       // w1, b1, W2 rows 0..49
    for (int i = tid; i < 2 * kH1; i += tg::kThreads) dst[LY::w2 + kH2 * kH1 + i] = 0.f;        // W2 rows 50, 51
    for (int i = tid; i < tg::kH2P; i += tg::kThreads) dst[LY::b2 + i] = i < kH2 ? theta[LY::tb2 + i] : 0.f;
    for (int i = tid; i < LY::SP * tg::kH2P; i += tg::kThreads) {
        const int s = i / tg::kH2P, o = i - s * tg::kH2P;
        dst[LY::w3 + i] = (s < S && o < kH2) ? theta[LY::tw3 + s * kH2 + o] : 0.f;
    }
    for (int i = tid; i < LY::SP; i += tg::kThreads) dst[LY::b3 + i] = i < S ? theta[LY::tb3 + i] : 0.f;
    for (int i = LY::P + tid; i < LY::PP; i += tg::kThreads) dst[i] = 0.f;
    __syncthreads();
}

__device__ __forceinline__ float block_sum256(float v, float *red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
    for (int i = 0; i < tg::kWarps; i++) t += red[i];
    __syncthreads();
    return t;
}

template <int S, bool META>
__global__ void __launch_bounds__(tg::kThreads, META ? 1 : 2) train_step_kernel(TrainParams p) {
    using LY = tg::Lay<S>;
    extern __shared__ __align__(16) float sm[];
    const int tid = threadIdx.x;
    for (int r = blockIdx.x; r < p.R; r += gridDim.x) {
        float *theta = p.theta + size_t(r) * LY::TP;
        float loss;
        tg::Smem<S> v = carve<S>(sm, META, META ? kLdnWide : kLdnPlain, META);
        load_theta<S>(v.W, theta);
        for (int i = tid; i < LY::PP; i += tg::kThreads) v.G[i] = 0.f;
        __syncthreads();
        if constexpr (!META) {
            loss = block_sum256(tg::pass<S, false, false>(v, p.y_s + size_t(r) * p.Ns, p.lab_s + size_t(r) * p.Ns, p.Ns,
                                                   1.f / float(p.Ns), 1.f, kCapPlain), v.red) / float(p.Ns);
        } else {
            const float *ys = p.y_s + size_t(r) * p.Ns, *yq = p.y_q + size_t(r) * p.Nq;
            const int *ls = p.lab_s + size_t(r) * p.Ns, *lq = p.lab_q + size_t(r) * p.Nq;
            // gradient of the support loss at theta, inner step theta' = theta - meta_lr * grad   (trainer.py:433-439),
            // then the query loss and its gradient at theta'                                    (trainer.py:442-444)
            float part = 0.f;
            tg::pass<S, false, true>(v, ys, ls, p.Ns, 1.f / float(p.Ns), 1.f, kCapWide);
            for (int i = tid; i < LY::PP; i += tg::kThreads) {
                v.W[i] -= p.meta_lr * v.G[i];
                v.G[i] = 0.f;
            }
            __syncthreads();
            part = tg::pass<S, false, true>(v, yq, lq, p.Nq, 1.f / float(p.Nq), 1.f, kCapWide);
            loss = block_sum256(part, v.red) / float(p.Nq);
            if (p.second_order) {
                // d/dtheta L_q(theta - a grad L_s(theta)) = g_q - a H_s(theta) g_q: forward-over-reverse pass of the
                // support loss at theta along v = g_q, accumulated into G with scale -a
                tg::Smem<S> t = carve<S>(sm, true, TanCap<S>::ldn);
                for (int i = tid; i < LY::PP; i += tg::kThreads) t.V[i] = t.G[i];
                load_theta<S>(t.W, theta);
                tg::pass<S, true, true>(t, ys, ls, p.Ns, 1.f / float(p.Ns), -p.meta_lr, TanCap<S>::cap);
            }
        }
        __syncthreads();
        const float *g = v.G;
        if (p.grad_out)
            for (int i = tid; i < LY::TP; i += tg::kThreads) p.grad_out[size_t(r) * LY::TP + i] = g[LY::to_padded(i)];
        if (p.loss_out && tid == 0) p.loss_out[r] = loss;
        // run_train_loop returns before backward / optimizer.step when the loss is NaN (trainer.py:495-498); the
        // meta loop has no such test (trainer.py:425-453)
        const bool skip = !META && isnan(loss);
        if (p.update && !skip) {
            const int step = p.adam_step[r] + 1;
            if (tid == 0) {
                v.red[0] = float(1.0 - pow(double(p.adam.beta1), double(step)));
                v.red[1] = float(sqrt(1.0 - pow(double(p.adam.beta2), double(step))));
            }
            __syncthreads();
            adam_update_padded<S>(theta, p.adam_m + size_t(r) * LY::TP, p.adam_v + size_t(r) * LY::TP, g, v.red, p.adam);
            __syncthreads();
            if (tid == 0) p.adam_step[r] = step;
        }
        __syncthreads();
    }
}

// Backward of the priors w.r.t. the six weight tensors for an arbitrary upstream gradient
// (autograd of the 'train' phase).  Grid over blocks of 256 symbols; partial sums are added to
// grad_theta with atomics (zeroed by the wrapper).
template <int S>
__global__ void __launch_bounds__(kTrainThreads, 1) priors_backward_kernel(const float *__restrict__ y, int64_t N,
                                                                          const float *__restrict__ theta,
                                                                          const float *__restrict__ grad_priors,
                                                                          float *grad_theta, float *workspace) {
    using TV = ThetaView<S>;
    constexpr int P = TV::P;
    extern __shared__ __align__(16) float sm[];
    float *th = sm;
    float *ysm = sm + (P + 3) / 4 * 4;
    float *slab = workspace + size_t(blockIdx.x) * kTrainThreads * Slab<S>::kRow;
    for (int i = threadIdx.x; i < P; i += kTrainThreads) th[i] = theta[i];
    __syncthreads();
    for (int64_t base = int64_t(blockIdx.x) * kTrainThreads; base < N; base += int64_t(gridDim.x) * kTrainThreads) {
        const int cnt = int(min((long long)kTrainThreads, (long long)(N - base)));
        if (int(threadIdx.x) < cnt) {
            const float yv = y[base + threadIdx.x];
            ysm[threadIdx.x] = yv;
            symbol_pass<S, false>(th, th, yv, 0, grad_priors + (base + threadIdx.x) * S, 1.f,
                                  slab + size_t(threadIdx.x) * Slab<S>::kRow);
        }
        __syncthreads();
        for (int idx = threadIdx.x; idx < P; idx += kTrainThreads)
            atomicAdd(grad_theta + idx, param_grad<S, false>(idx, slab, ysm, cnt));
        __syncthreads();
    }
}

// double backward: for F(theta, g) = J(theta)^T g (the kernel above) and an upstream u [P]:
//   d<u,F>/dg     = J u                      -> grad_gp [N,S]   (forward tangent of the priors along u)
//   d<u,F>/dtheta = R_u{ J(theta)^T g }      -> grad_theta2 [P] (forward-over-reverse, g held fixed)
// This is what torch.autograd.grad(..., create_graph=True) in trainer.py:437 differentiates through.
template <int S>
__global__ void __launch_bounds__(kTrainThreads, 1) priors_backward2_kernel(const float *__restrict__ y, int64_t N,
                                                                           const float *__restrict__ theta,
                                                                           const float *__restrict__ grad_priors,
                                                                           const float *__restrict__ u, float *grad_theta2,
                                                                           float *__restrict__ grad_gp, float *workspace) {
    using TV = ThetaView<S>;
    constexpr int P = TV::P, PP = (P + 3) / 4 * 4;
    extern __shared__ __align__(16) float sm[];
    float *th = sm;
    float *tu = sm + PP;
    float *ysm = tu + PP;
    float *slab = workspace + size_t(blockIdx.x) * kTrainThreads * Slab<S>::kRow;
    for (int i = threadIdx.x; i < P; i += kTrainThreads) {
        th[i] = theta[i];
        tu[i] = u[i];
    }
    __syncthreads();
    for (int64_t base = int64_t(blockIdx.x) * kTrainThreads; base < N; base += int64_t(gridDim.x) * kTrainThreads) {
        const int cnt = int(min((long long)kTrainThreads, (long long)(N - base)));
        if (int(threadIdx.x) < cnt) {
            const float yv = y[base + threadIdx.x];
            ysm[threadIdx.x] = yv;
            symbol_pass<S, true>(th, tu, yv, 0, grad_priors + (base + threadIdx.x) * S, 1.f,
                                 slab + size_t(threadIdx.x) * Slab<S>::kRow, grad_gp + (base + threadIdx.x) * S);
        }
        __syncthreads();
        for (int idx = threadIdx.x; idx < P; idx += kTrainThreads)
            atomicAdd(grad_theta2 + idx, param_grad<S, true>(idx, slab, ysm, cnt));
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// Detection with PER-REALISATION weights (the eval_by_word shape, trainer.py:267-298: one word per call and weights
// that change between calls).  One CTA per realisation: thread per symbol evaluates the priors with that realisation's
// theta (blocks of 256 symbols into shared memory), then one thread runs the reference's ACS / decision recursion on
// its register trellis — the same RegTrellis code as the batch kernels, so decisions are bit-exact functions of the
// priors this kernel computes (and optionally exports).
// ---------------------------------------------------------------------------------------------
template <int L>
__global__ void __launch_bounds__(kTrainThreads, 1) detect_batched_kernel(const float *__restrict__ theta_all, int theta_stride,
                                                                         int R, const float *__restrict__ y, int T, int n_stages,
                                                                         float *__restrict__ decoded,
                                                                         float *__restrict__ priors_out) {
    constexpr int S = 1 << L;
    using TV = ThetaView<S>;
    using D = TrellisDims<L>;
    constexpr int P = TV::P, PP = (P + 3) / 4 * 4, C = D::C;
    extern __shared__ __align__(16) float sm[];
    float *th = sm, *pri = sm + PP;  // pri[256][S]
    for (int r = blockIdx.x; r < R; r += gridDim.x) {
        for (int i = threadIdx.x; i < P; i += kTrainThreads) th[i] = theta_all[size_t(r) * theta_stride + i];
        __syncthreads();
        RegTrellis<L> tr;
        tr.reset();
        for (int base = 0; base < T; base += kTrainThreads) {
            const int t = base + threadIdx.x;
            if (t < n_stages) {
                const float yv = y[size_t(r) * T + t];
                float h1[kH1];
#pragma unroll
                for (int k = 0; k < kH1; k++) h1[k] = sigmoidf_acc(fmaf(th[TV::w1 + k], yv, th[TV::b1 + k]));
                float z[S];
#pragma unroll
                for (int s = 0; s < S; s++) z[s] = th[TV::b3 + s];
#pragma unroll 1
                for (int o = 0; o < kH2; o++) {
                    const float *w2 = th + TV::w2 + o * kH1;
                    float a0 = th[TV::b2 + o], a1 = 0.f;
#pragma unroll
                    for (int k = 0; k < kH1; k += 2) {
                        a0 = fmaf(w2[k], h1[k], a0);
                        a1 = fmaf(w2[k + 1], h1[k + 1], a1);
                    }
                    const float h2 = fmaxf(a0 + a1, 0.f);
#pragma unroll
                    for (int s = 0; s < S; s++) z[s] = fmaf(th[TV::w3 + s * kH2 + o], h2, z[s]);
                }
#pragma unroll
                for (int s = 0; s < S; s++) {
                    pri[threadIdx.x * S + s] = z[s];
                    if (priors_out) priors_out[(size_t(r) * T + t) * S + s] = z[s];
                }
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                const int t_end = min(kTrainThreads, n_stages - base);
                for (int tt = 0; tt < t_end; tt++) {
                    decoded[size_t(r) * T + base + tt] = float(tr.decide());   // before this stage's ACS (vnet_detector.py:55-61)
                    float cost[C];
                    if constexpr (D::NCH == 1) {
#pragma unroll
                        for (int i = 0; i < C; i++) cost[i] = -pri[tt * S + i];
                        tr.template step_chunk<0>(cost);
                    } else {
#pragma unroll
                        for (int i = 0; i < C; i++) cost[i] = -pri[tt * S + i];
                        tr.template step_chunk<0>(cost);
#pragma unroll
                        for (int i = 0; i < C; i++) cost[i] = -pri[tt * S + C + i];
                        tr.template step_chunk<1>(cost);
                    }
                    tr.commit();
                }
            }
            __syncthreads();
        }
        for (int t = max(n_stages, 0) + threadIdx.x; t < T; t += kTrainThreads) decoded[size_t(r) * T + t] = 0.f;
        __syncthreads();
    }
}

template <int L>
static int launch_detect_batched(const float *theta, int shared_theta, int R, const float *y, int T, int n_stages, float *decoded,
                                 float *priors, cudaStream_t st) {
    constexpr int S = 1 << L;
    constexpr int P = ThetaView<S>::P;
    const size_t smem = (size_t((P + 3) / 4 * 4) + size_t(kTrainThreads) * S) * sizeof(float);
    auto kern = detect_batched_kernel<L>;
    MVN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    kern<<<std::max(1, std::min(R, 2 * sm_count())), kTrainThreads, smem, st>>>(theta, shared_theta ? 0 : P, R, y, T, n_stages, decoded,
                                                                                priors);
    note_launch();
    MVN_CUDA(cudaGetLastError());
    return MVN_OK;
}

static int train_grid(int R) { return std::max(1, std::min(R, 2 * sm_count())); }

template <int L>
static int launch_train(const TrainParams &p, bool meta, cudaStream_t st) {
    constexpr int S = 1 << L;
    if (meta) {
        auto kern = train_step_kernel<S, true>;
        const size_t smem = StepSmem<S, true>::bytes;
        MVN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        kern<<<std::max(1, std::min(p.R, sm_count())), tg::kThreads, smem, st>>>(p);
    } else {
        auto kern = train_step_kernel<S, false>;
        const size_t smem = StepSmem<S, false>::bytes;
        MVN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        kern<<<train_grid(p.R), tg::kThreads, smem, st>>>(p);
    }
    note_launch();
    MVN_CUDA(cudaGetLastError());
    return MVN_OK;
}

template <int L>
static int launch_priors_bwd(const float *y, int64_t N, const float *theta, const float *gp, float *gt, float *ws,
                             cudaStream_t st) {
    constexpr int S = 1 << L;
    constexpr int P = ThetaView<S>::P;
    const size_t smem = (size_t((P + 3) / 4 * 4) + kTrainThreads) * sizeof(float);
    auto kern = priors_backward_kernel<S>;
    MVN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    MVN_CUDA(cudaMemsetAsync(gt, 0, P * sizeof(float), st));
    const int grid = int(std::max<int64_t>(1, std::min<int64_t>((N + kTrainThreads - 1) / kTrainThreads, 2 * sm_count())));
    kern<<<grid, kTrainThreads, smem, st>>>(y, N, theta, gp, gt, ws);
    note_launch();
    MVN_CUDA(cudaGetLastError());
    return MVN_OK;
}

template <int L>
static int launch_priors_bwd2(const float *y, int64_t N, const float *theta, const float *gp, const float *u, float *gt2,
                              float *ggp, float *ws, cudaStream_t st) {
    constexpr int S = 1 << L;
    constexpr int P = ThetaView<S>::P;
    const size_t smem = (size_t(2) * ((P + 3) / 4 * 4) + kTrainThreads) * sizeof(float);
    auto kern = priors_backward2_kernel<S>;
    MVN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    MVN_CUDA(cudaMemsetAsync(gt2, 0, P * sizeof(float), st));
    const int grid = int(std::max<int64_t>(1, std::min<int64_t>((N + kTrainThreads - 1) / kTrainThreads, 2 * sm_count())));
    kern<<<grid, kTrainThreads, smem, st>>>(y, N, theta, gp, u, gt2, ggp, ws);
    note_launch();
    MVN_CUDA(cudaGetLastError());
    return MVN_OK;
}

static size_t slab_row_floats(int L) {
    switch (L) {
        case 1: return Slab<2>::kRow;
        case 2: return Slab<4>::kRow;
        case 3: return Slab<8>::kRow;
        case 4: return Slab<16>::kRow;
        case 5: return Slab<32>::kRow;
        default: return 0;
    }
}

}  // namespace mvn

using namespace mvn;

#define MVN_TRAIN_DISPATCH(L, FN, ...)                                                          \
    switch (L) {                                                                                \
        case 1: return FN<1>(__VA_ARGS__);                                                      \
        case 2: return FN<2>(__VA_ARGS__);                                                      \
        case 3: return FN<3>(__VA_ARGS__);                                                      \
        case 4: return FN<4>(__VA_ARGS__);                                                      \
        case 5: return FN<5>(__VA_ARGS__);                                                      \
        default:                                                                                \
            set_error("training kernels support memory_length 1..5 (reference: 'tested <= 4'), got %d", L); \
            return MVN_ERR_UNSUPPORTED;                                                         \
    }

extern "C" int mvn_param_count(int L) { return (L < 1 || L > 8) ? -1 : param_count_s(1 << L); }

extern "C" int64_t mvn_meta_workspace_bytes(int L, int R, int n_max) {
    (void)n_max;  // the batched step kernels keep everything in shared memory; the argument survives for ABI stability
    if (!slab_row_floats(L) || R < 1) return -1;
    return 256;
}

extern "C" int64_t mvn_priors_backward_workspace_bytes(int L, int64_t N) {
    const size_t row = slab_row_floats(L);
    if (!row || N < 0) return -1;
    const int64_t grid = std::max<int64_t>(1, std::min<int64_t>((N + kTrainThreads - 1) / kTrainThreads, 2 * sm_count()));
    return grid * kTrainThreads * row * sizeof(float);
}

static int train_common(TrainParams &p, int L, bool meta, void *stream) {
    if (!p.theta || p.R < 0 || !p.y_s || !p.lab_s || p.Ns < 1 ||
        (p.update && (!p.adam_m || !p.adam_v || !p.adam_step)) || (meta && (!p.y_q || !p.lab_q || p.Nq < 1))) {
        set_error("batched training step: bad argument");
        return MVN_ERR_ARG;
    }
    if (p.R == 0) return MVN_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    MVN_TRAIN_DISPATCH(L, launch_train, p, meta, st)
}

extern "C" int mvn_meta_step_batched(float *theta, float *adam_m, float *adam_v, int32_t *adam_step, int R, int L,
                                     const float *y_s, const int32_t *lab_s, int Ns, const float *y_q,
                                     const int32_t *lab_q, int Nq, float meta_lr, float lr, int second_order,
                                     float *loss_out, float *grad_out, void *workspace, void *stream) {
    TrainParams p{theta, adam_m, adam_v, adam_step, R, y_s, lab_s, Ns, y_q, lab_q, Nq, meta_lr,
                  AdamCfg{lr, 0.9f, 0.999f, 1e-8f}, second_order, loss_out, grad_out,
                  static_cast<float *>(workspace), adam_m != nullptr};
    return train_common(p, L, true, stream);
}

extern "C" int mvn_train_step_batched(float *theta, float *adam_m, float *adam_v, int32_t *adam_step, int R, int L,
                                      const float *y, const int32_t *lab, int N, float lr, float *loss_out,
                                      float *grad_out, void *workspace, void *stream) {
    TrainParams p{theta, adam_m, adam_v, adam_step, R, y, lab, N, nullptr, nullptr, 0, 0.f,
                  AdamCfg{lr, 0.9f, 0.999f, 1e-8f}, 0, loss_out, grad_out, static_cast<float *>(workspace),
                  adam_m != nullptr};
    return train_common(p, L, false, stream);
}

extern "C" int mvn_vnet_priors_backward(const float *y, int64_t N, int L, const float *theta, const float *grad_priors,
                                        float *grad_theta, void *workspace, void *stream) {
    if (N < 0 || !theta || !grad_theta || !workspace || (N > 0 && (!y || !grad_priors))) {
        set_error("mvn_vnet_priors_backward: bad argument");
        return MVN_ERR_ARG;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    MVN_TRAIN_DISPATCH(L, launch_priors_bwd, y, N, theta, grad_priors, grad_theta, static_cast<float *>(workspace), st)
}

extern "C" int mvn_vnet_priors_backward2(const float *y, int64_t N, int L, const float *theta, const float *grad_priors,
                                         const float *u, float *grad_theta2, float *grad_grad_priors, void *workspace,
                                         void *stream) {
    if (N < 0 || !theta || !u || !grad_theta2 || !workspace || (N > 0 && (!y || !grad_priors || !grad_grad_priors))) {
        set_error("mvn_vnet_priors_backward2: bad argument");
        return MVN_ERR_ARG;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    MVN_TRAIN_DISPATCH(L, launch_priors_bwd2, y, N, theta, grad_priors, u, grad_theta2, grad_grad_priors,
                       static_cast<float *>(workspace), st)
}

extern "C" int mvn_vnet_detect_batched(const float *theta, int R, int L, const float *y, int T, int n_stages, float *decoded,
                                       float *priors_out, void *stream) {
    if (R < 0 || T < 0 || n_stages < 0 || n_stages > T || (R > 0 && T > 0 && (!theta || !y || !decoded))) {
        set_error("mvn_vnet_detect_batched: bad argument (n_stages must be 0..T)");
        return MVN_ERR_ARG;
    }
    if (R == 0 || T == 0) return MVN_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    MVN_TRAIN_DISPATCH(L, launch_detect_batched, theta, 0, R, y, T, n_stages, decoded, priors_out, st)
}

extern "C" int mvn_vnet_detect_small(const float *theta, int64_t B, int L, const float *y, int T, int n_stages, float *decoded,
                                     float *priors_out, void *stream) {
    if (B < 0 || B > (1 << 20) || T < 0 || n_stages < 0 || n_stages > T || (B > 0 && T > 0 && (!theta || !y || !decoded))) {
        set_error("mvn_vnet_detect_small: bad argument (n_stages must be 0..T)");
        return MVN_ERR_ARG;
    }
    if (B == 0 || T == 0) return MVN_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    MVN_TRAIN_DISPATCH(L, launch_detect_batched, theta, 1, int(B), y, T, n_stages, decoded, priors_out, st)
}
