/*
 * mvn_b200.h — C ABI of the B200-native Meta-ViterbiNet detection hot path.
 *
 * The reference (tomerraviv95/meta-viterbinet) has no FFI of its own: its seam is the Python
 * detector API.  This header is the boundary a maintainer binds instead
 * (ctypes stub in INTEGRATION.md); every entry point names the reference code it replaces
 * (paths relative to the reference checkout).
 *
 * Conventions
 *   - All pointers are DEVICE pointers unless the name ends in `_host`.  fp32 = IEEE binary32,
 *     tensors contiguous row-major.  `stream` is a cudaStream_t passed as void* (NULL = default).
 *   - Calls are asynchronous on `stream` and allocate no device memory.  Library state shared between calls is
 *     limited to: (1) the pair of constant-bank weight slots of the FP32-FMA variants (handed over between streams
 *     with events), (2) the tcgen05 pipeline-watchdog word per device (mvn_tc_timeout_status / mvn_reset_tc_timeout),
 *     (3) lazily created CUDA events for (1).  Which kernel variant runs is a per-call / per-context argument,
 *     never a process-wide switch.  Thread-compatible: concurrent calls from several host threads are safe.
 *   - Return 0 on success, non-zero MVN_ERR_* otherwise; mvn_last_error() describes the last
 *     failure of the calling thread.
 *   - memory_length L in [1,8] (n_states S = 2^L); the trellis is the reference's:
 *     predecessors of state j are (2j) mod S and (2j+1) mod S, the branch cost is indexed by the
 *     SOURCE state (python_code/utils/trellis_utils.py:7-30).
 *   - "decode" = the reference's stage loop (va_detector.py:83-98, vnet_detector.py:46-61):
 *     pm = 0; for t < n_stages: bit[t] = (lowest s attaining min pm) & 1, then one ACS stage.
 *     Output columns t >= n_stages are written as 0.
 */
#ifndef MVN_B200_H
#define MVN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MVN_OK 0
#define MVN_ERR_ARG 1       /* bad argument (NULL pointer, L out of range, n_stages > T, ...) */
#define MVN_ERR_CUDA 2      /* a CUDA runtime call failed */
#define MVN_ERR_UNSUPPORTED 3

#define MVN_HIDDEN1 100     /* vnet_detector.py:7  */
#define MVN_HIDDEN2 50      /* vnet_detector.py:8  */

/* Output formats for decoded bits. */
#define MVN_OUT_F32 0       /* [B,T] fp32 0.0/1.0 — the reference's dtype                     */
#define MVN_OUT_BITS 1      /* [B,ceil(T/32)] uint32, bit (t%32) of word t/32 = decoded[b,t]  */

/* Implementation of the fused ViterbiNet kernel (argument `variant` of mvn_vnet_decode_ex / mvn_ctx_set_variant). */
#define MVN_VARIANT_AUTO 0          /* library default: tcgen05 tensor cores for every memory_length */
#define MVN_VARIANT_FMA_SMEM 1      /* FP32 FMA pipe, weights staged in shared memory              */
#define MVN_VARIANT_FMA_CONST320 2  /* FP32 FMA pipe, constant-bank weights, 320 threads (tuning)  */
#define MVN_VARIANT_TCGEN05 3       /* tensor cores explicitly                                     */
#define MVN_VARIANT_FMA 4           /* FP32 FMA pipe, constant-bank weights (default FMA form)     */

/* Decision rule (argument `decision` of the *_ex entry points). */
#define MVN_DECIDE_REFERENCE 0      /* the reference's rule: bit t = (lowest argmin of pm) & 1 BEFORE stage t's ACS  */
#define MVN_DECIDE_MLSE 1           /* true MLSE: in-kernel survivor traceback from the best final state             */
#define MVN_DECIDE_MLSE_TERMINATED 2 /* traceback from state 0 (the reference pads every word with L zero bits)      */

const char *mvn_last_error(void);
int mvn_version(void);
/* Device properties needed by callers to size grids/rooflines: fills sm_count, clock kHz. */
int mvn_device_info(int *sm_count, int *sm_clock_khz, int *cc_major, int *cc_minor);

/* Error counters accumulated by the kernels: 4 x uint64. */
#define MVN_CNT_BIT_ERRORS 0
#define MVN_CNT_FRAME_ERRORS 1
#define MVN_CNT_BITS 2
#define MVN_CNT_FRAMES 3

/* ---- a2: one ACS stage.  Replaces trellis_utils.py:16-30 (acs_block).
 * in_prob [B,S]; llrs [B,S] (llrs_stride = S) or [B,1] broadcast (llrs_stride = 1).
 * out_prob [B,S] fp32; out_idx [B,S] int64 in {0,1} (tie -> 0), may be NULL. */
int mvn_acs_block(const float *in_prob, const float *llrs, int llrs_stride, int64_t B, int L,
                  float *out_prob, int64_t *out_idx, void *stream);

/* ---- a3: stage loop on an arbitrary cost tensor.  Replaces the loop va_detector.py:91-97 /
 * vnet_detector.py:53-59 / meta_vnet_detector.py:37-43 for cost [B,T,S].
 * decoded: MVN_OUT_F32 or MVN_OUT_BITS buffer (may be NULL if only counters are wanted).
 * final_pm [B,S] fp32 (optional).  survivors [B,n_stages,max(1,S/64)] uint32 (optional):
 * bit j of the word group = survivor (0/1) chosen for new state j < S/2 (states j and j+S/2 are
 * identical); this is trellis_utils.py:30's index output, bit-packed. */
int mvn_acs_decode(const float *cost, int64_t B, int T, int L, int n_stages, int out_format,
                   void *decoded, float *final_pm, uint32_t *survivors, void *stream);

/* Same with the thread layout selectable (for measurements; AUTO is what mvn_acs_decode uses):
 * LANE_PER_FRAME  a lane owns a frame, metrics in registers (<= 64 states) / shared memory, no cross-lane traffic;
 * STATES_ON_LANES the distinct metrics of a frame sit on S/2 lanes (4..64 states, 32/(S/2) frames per warp) or S/64 per
 *                 lane (128 / 256 states, one frame per warp), butterflies and decision with warp shuffles / REDUX.
 * AUTO = lane per frame up to 64 states, states on lanes at 128 / 256.  The survivor export always runs lane-per-frame. */
#define MVN_LAYOUT_AUTO 0
#define MVN_LAYOUT_LANE_PER_FRAME 1
#define MVN_LAYOUT_STATES_ON_LANES 2
int mvn_acs_decode_ex(const float *cost, int64_t B, int T, int L, int n_stages, int out_format, void *decoded,
                      float *final_pm, uint32_t *survivors, int layout, void *stream);

/* ---- a5+a3 fused: classical Viterbi with full CSI.  Replaces VADetector.forward
 * (va_detector.py:52-98) given the host-built table of va_detector.py:42-50.
 * y [B,T]; state_priors [n_h,S] fp32 (row k = noiseless outputs of tap block k; NOTE this is
 * the transpose of compute_state_priors' [S,n_h]); frame b uses block (b mod n_h).
 * cost = (y-sp)^2/2 - fp32(ln sqrt(2 pi)), each op rounded separately, never materialised.
 * target/counters: optional fused BER/FER accumulation (see mvn_error_counts). */
int mvn_va_decode(const float *y, int64_t B, int T, int L, int n_stages, const float *state_priors,
                  int n_h, int out_format, void *decoded, const float *target, int target_T,
                  int pilot_period, uint64_t *counters, void *stream);

/* Same with the decision rule selectable: MVN_DECIDE_MLSE* keep the S/2 survivor bits of every stage
 * (trellis_utils.py:30's indices, which the reference discards) as bit masks in shared memory and trace back inside
 * the kernel at the end of the frame — nothing but y and the decoded bits touches HBM (8 B per symbol). */
int mvn_va_decode_ex(const float *y, int64_t B, int T, int L, int n_stages, const float *state_priors, int n_h,
                     int out_format, void *decoded, const float *target, int target_T, int pilot_period,
                     uint64_t *counters, int decision, void *stream);

/* ---- a6: ViterbiNet priors.  Replaces VNETDetector.net / META_VNETDetector's F.linear chain
 * (vnet_detector.py:27-33,49; meta_vnet_detector.py:27-33).  Weights in torch nn.Linear layout:
 * w1 [100,1], b1 [100], w2 [50,100], b2 [50], w3 [S,50], b3 [S].  y flattened [N]; priors [N,S]. */
int mvn_vnet_priors(const float *y, int64_t N, int L, const float *w1, const float *b1, const float *w2,
                    const float *b2, const float *w3, const float *b3, float *priors, void *stream);

/* ---- a6+a3 fused: priors MLP + stage loop + decision, nothing but y and bits touch HBM.
 * Replaces VNETDetector.forward(y,'val') (vnet_detector.py:35-61) and the 'val' branch of
 * META_VNETDetector.forward (meta_vnet_detector.py:24-45).
 * priors_out [B,T,S] optional (parity export; the ACS consumes exactly these values). */
int mvn_vnet_decode(const float *y, int64_t B, int T, int L, int n_stages, const float *w1, const float *b1,
                    const float *w2, const float *b2, const float *w3, const float *b3, int out_format,
                    void *decoded, float *priors_out, const float *target, int target_T, int pilot_period,
                    uint64_t *counters, void *stream);

/* Same with the kernel variant (MVN_VARIANT_*) and the decision rule (MVN_DECIDE_*) chosen per call. */
int mvn_vnet_decode_ex(const float *y, int64_t B, int T, int L, int n_stages, const float *w1, const float *b1,
                       const float *w2, const float *b2, const float *w3, const float *b3, int out_format,
                       void *decoded, float *priors_out, const float *target, int target_T, int pilot_period,
                       uint64_t *counters, int variant, int decision, void *stream);
/* tcgen05 pipeline watchdog of the current device: status = 1 if an mbarrier wait of the tensor-core kernel gave up
 * (2 s of wall clock; never expected) since the last reset; such a launch's output is invalid and further tcgen05
 * launches on that device are refused until mvn_reset_tc_timeout(). */
int mvn_tc_timeout_status(void);
int mvn_reset_tc_timeout(void);

/* ---- a8: ground-truth state labels.  Replaces trellis_utils.py:33-46 (calculate_states).
 * tx [B,T] fp32 0/1 -> states [B*T] int64, state[b,t] = sum_{i<L} tx[b,t+i] 2^i (zero past T). */
int mvn_calculate_states(const float *tx, int64_t B, int T, int L, int64_t *states, void *stream);

/* ---- a12: BER/FER.  Replaces metrics.py:7-17 (calculate_error_rates) with exact integers.
 * prediction [B,Tp] and target [B,Tt] fp32 (compared on the first min(Tp,Tt)=T columns after
 * truncation to integer, like .long()); rows with (b % pilot_period == 0) are skipped when
 * pilot_period > 0 (trainer.py:100-102).  counters[4] uint64 are ADDED to (zero them first).
 * row_errors [B] uint8 optional: 1 where the row has any error (torch.nonzero source). */
int mvn_error_counts(const float *prediction, int pred_stride, const float *target, int target_stride,
                     int64_t B, int T, int pilot_period, uint64_t *counters, uint8_t *row_errors, void *stream);

/* ---- end-to-end with HOST buffers (pinned or pageable): chunked H2D -> fused decode -> D2H
 * pipeline on the context's own streams and device buffers.  This is what bench.py's `e2e`
 * times.  y_host [B,T] fp32 -> decoded_host (out_format).  Weights are host pointers too. */
typedef struct mvn_ctx mvn_ctx;
/* chunk_frames <= 0 selects two full waves of the fused kernel per chunk. */
int mvn_ctx_create(mvn_ctx **ctx, int device, int64_t chunk_frames, int T_max, int L);
void mvn_ctx_destroy(mvn_ctx *ctx);
/* (pageable source arrays may be reused as soon as the call returns; pinned ones after the next synchronising call) */
int mvn_ctx_set_vnet_weights_host(mvn_ctx *ctx, const float *w1, const float *b1, const float *w2,
                                  const float *b2, const float *w3, const float *b3);
int mvn_ctx_vnet_decode_host(mvn_ctx *ctx, const float *y_host, int64_t B, int T, int n_stages,
                             int out_format, void *decoded_host);
int mvn_ctx_va_decode_host(mvn_ctx *ctx, const float *y_host, int64_t B, int T, int n_stages,
                           const float *state_priors_host, int n_h, int out_format, void *decoded_host);
/* Asynchronous form for streaming several batches (e.g. the points of an SNR sweep, each after its own
 * mvn_ctx_set_vnet_weights_host: the context keeps a ring of 8 weight sets) through ONE pipeline without draining it
 * between them: returns once everything is enqueued; the host buffers must stay valid, and the results are complete,
 * after mvn_ctx_synchronize (which also reports CUDA errors and the tcgen05 watchdog of the enqueued work). */
int mvn_ctx_vnet_decode_host_async(mvn_ctx *ctx, const float *y_host, int64_t B, int T, int n_stages, int out_format,
                                   void *decoded_host);
int mvn_ctx_synchronize(mvn_ctx *ctx);
/* kernel variant (MVN_VARIANT_*) and decision rule (MVN_DECIDE_*) used by this context's decode calls */
int mvn_ctx_set_variant(mvn_ctx *ctx, int variant);
int mvn_ctx_set_decision(mvn_ctx *ctx, int decision);
/* Same pipeline with the error counters fused: target_host [B,target_T] fp32 words (the transmitted bits the
 * reference's dataset returns next to y) are uploaded chunk by chunk next to y, BER / FER are counted in the decode
 * kernel (metrics.py:7-17, pilot rows b % pilot_period == 0 skipped) and ONLY counters_host[4] (32 bytes) come back;
 * decoded_host may be NULL (or a MVN_OUT_* buffer to also get the words). */
int mvn_ctx_vnet_eval_host(mvn_ctx *ctx, const float *y_host, const float *target_host, int64_t B, int T, int n_stages,
                           int target_T, int pilot_period, int out_format, void *decoded_host, uint64_t *counters_host);
/* Monte-Carlo point evaluated entirely on the device (the sweep shape of trainer.py:222-241 / plotter_main.py:117-122
 * at scale): Bernoulli(1/2) words and AWGN from Philox (seed), BPSK + ISI with taps_host [n_h,L] float64 (row b mod n_h)
 * at snr_db (mvn_channel_transmit), fused ViterbiNet decode with in-kernel BER / FER; counters_host[4] is the only
 * device->host traffic.  Frames are processed in the context's chunks. */
int mvn_ctx_vnet_sweep_point(mvn_ctx *ctx, int64_t B, int T, int n_stages, const double *taps_host, int n_h,
                             double snr_db, uint64_t seed, int pilot_period, uint64_t *counters_host);
/* Pinned (page-locked) host buffers for the *_host entry points; write_combined != 0 allocates the buffer
 * write-combined (fast for the device to read over PCIe, slow for the host to read back: inputs only). */
int mvn_host_alloc(void **ptr, size_t bytes, int write_combined);
int mvn_host_free(void *ptr);
/* Measurement helper: raw pinned host <-> device copy rate of `device` with the same chunking and stream count as the
 * pipeline above but no kernel (the e2e ceiling).  h2d / d2h: which directions run (both = concurrently);
 * seconds = wall time of `reps` passes over `bytes`. */
int mvn_copy_ceiling(int device, void *host_in, void *host_out, size_t bytes, size_t chunk_bytes, int h2d, int d2h,
                     int reps, double *seconds);
/* number of kernel launches issued by this library on this thread since the last reset
 * (bench.py's gpu_launches claim). */
int64_t mvn_launch_count(int reset);

/* ---- measurement helper: register-only FP32 FMA micro-benchmark (SURVEY.md §8d asks for a
 * measured FP32 peak).  mode 0 = scalar FFMA, 1 = packed FFMA2.  Returns achieved FMA lane-ops/s
 * through *fma_per_s (each FMA = 2 flop). */
int mvn_fp32_peak(int mode, int iters, double *fma_per_s, double *ms, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MVN_B200_H */
