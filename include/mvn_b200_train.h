/*
 * mvn_b200_train.h — C ABI of the batched (meta-)training of the ViterbiNet priors network
 * (SURVEY.md §8a rows a9-a11).  Same conventions as mvn_b200.h (device pointers, fp32, async on
 * `stream`, 0 = ok, mvn_last_error()).
 *
 * Parameters are PACKED per realisation in torch parameter order
 *   theta = [ w1 (100x1) | b1 (100) | w2 (50x100) | b2 (50) | w3 (Sx50) | b3 (S) ],  P = 5350 + 51 S
 * (the order of VNETDetector.parameters(), vnet_detector.py:27-33).
 * Loss: torch CrossEntropyLoss (mean) of the priors against the state labels of
 * trellis_utils.py:33-46 over ALL symbols (metavnet_trainer.py:41-50).
 * Optimiser: torch.optim.Adam defaults betas=(0.9,0.999), eps=1e-8 (trainer.py:167-169).
 * memory_length 1..5 (the reference states "tested with values <= 4", config.yaml:9).
 */
#ifndef MVN_B200_TRAIN_H
#define MVN_B200_TRAIN_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* P for a given memory_length (-1 if out of range). */
int mvn_param_count(int L);

/* Device scratch needed by the two batched steps for R realisations (bytes, -1 on bad args). */
int64_t mvn_meta_workspace_bytes(int L, int R, int n_max);

/* ---- a10: R independent MAML / FO-MAML steps.  Replaces Trainer.meta_train_loop
 * (trainer.py:425-453) run once per realisation:
 *   g_s = grad L(theta; support);  theta' = theta - meta_lr g_s;
 *   g_q = grad L(theta'; query);   meta_grad = second_order ? g_q - meta_lr H_s(theta) g_q : g_q;
 *   Adam step on theta with meta_grad.
 * theta/adam_m/adam_v [R,P]; adam_step [R] int32 = steps already taken (incremented);
 * y_s [R,Ns] / lab_s [R,Ns] int32 labels (support, Ns = window_size * T symbols),
 * y_q [R,Nq] / lab_q [R,Nq] (query).  loss_out [R] = query loss (optional), grad_out [R,P] =
 * meta-gradient (optional).  adam_m == NULL: compute loss/gradient only, no update. */
int mvn_meta_step_batched(float *theta, float *adam_m, float *adam_v, int32_t *adam_step, int R, int L,
                          const float *y_s, const int32_t *lab_s, int Ns, const float *y_q,
                          const int32_t *lab_q, int Nq, float meta_lr, float lr, int second_order,
                          float *loss_out, float *grad_out, void *workspace, void *stream);

/* ---- a11: R independent plain steps.  Replaces run_train_loop (trainer.py:492-505) /
 * online_training's inner iteration (metavnet_trainer.py:52-64): CE over the N symbols, Adam. */
int mvn_train_step_batched(float *theta, float *adam_m, float *adam_v, int32_t *adam_step, int R, int L,
                           const float *y, const int32_t *lab, int N, float lr, float *loss_out,
                           float *grad_out, void *workspace, void *stream);

/* ---- backward of the 'train'-phase priors (autograd of VNETDetector.forward(y,'train'),
 * vnet_detector.py:49,63): grad_theta [P] = d<grad_priors, priors(y; theta)>/d theta for ONE
 * parameter set, y [N], grad_priors [N,S].  workspace: mvn_priors_backward_workspace_bytes. */
int64_t mvn_priors_backward_workspace_bytes(int L, int64_t N);
int mvn_vnet_priors_backward(const float *y, int64_t N, int L, const float *theta, const float *grad_priors,
                             float *grad_theta, void *workspace, void *stream);

/* ---- double backward (torch.autograd.grad(..., create_graph=True) of the support loss, trainer.py:437;
 * the query loss is then differentiated through the fast weights, trainer.py:441-449).
 * With F(theta, g) = mvn_vnet_priors_backward's result and an upstream u [P]:
 *   grad_theta2 [P]        = d<u,F>/d theta   (g held fixed)
 *   grad_grad_priors [N,S] = d<u,F>/d g       (= Jacobian of the priors applied to u)
 * Same workspace size as the first backward. */
int mvn_vnet_priors_backward2(const float *y, int64_t N, int L, const float *theta, const float *grad_priors,
                              const float *u, float *grad_theta2, float *grad_grad_priors, void *workspace,
                              void *stream);

/* ---- detection with per-realisation weights: VNETDetector.forward(y, 'val') (vnet_detector.py:46-61) for R
 * independent runs at once, run r using theta[r] on its own word y[r] — the shape of the word-by-word online
 * evaluation (trainer.py:267-298), where every run's weights differ and change between blocks.
 * theta [R,P], y [R,T], decoded [R,T] fp32 0/1 (columns >= n_stages are 0), priors_out [R,T,S] optional. */
int mvn_vnet_detect_batched(const float *theta, int R, int L, const float *y, int T, int n_stages, float *decoded,
                            float *priors_out, void *stream);

/* Same kernel with ONE weight set shared by all B words: the single-launch form of VNETDetector.forward(y, 'val') for
 * batches too small to fill the GPU with the frame-per-lane kernels (B = 1 in eval_by_word, trainer.py:295; B = 300 in the
 * aggregated evaluation of plotter_main.py:96-111).  theta [P] packed in torch parameter order; memory_length 1..5. */
int mvn_vnet_detect_small(const float *theta, int64_t B, int L, const float *y, int T, int n_stages, float *decoded,
                          float *priors_out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MVN_B200_TRAIN_H */
