/*
 * mvn_b200_next.h — C ABI of the rows SURVEY.md §8(f) marks "next": the callers / data formats on either
 * side of the detection path.  Conventions as in mvn_b200.h.
 */
#ifndef MVN_B200_NEXT_H
#define MVN_B200_NEXT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- f1: channel simulator.  Replaces the per-word numpy path ChannelModelDataset.transmit ->
 * BPSKModulator.modulate -> ISIAWGNChannel.transmit (channel_dataset.py:71,87-95; modulator.py:12; channel.py:12-35):
 * codeword bits [B,T] fp32 0/1 (already RS-encoded if ECC is on), padded with L zero bits, s = 1-2c,
 * y[t] = sum_i h[L-1-i] s[t+i] + 10^(-snr/20) n[t], computed in float64 like numpy and stored as fp32.
 * taps [n_h, L] float64 (frame b uses row b mod n_h: n_h = B for per-word fading, 1 for a static channel).
 * noise [B,T] float64 standard-normal samples for bit-exact parity with a reference run, or NULL to draw them on the
 * device (Philox4x32-10, `seed`, one subsequence per symbol). */
int mvn_channel_transmit(const float *bits, int64_t B, int T, int L, const double *taps, int n_h, double snr_db,
                         const double *noise, uint64_t seed, float *y, void *stream);

/* Bernoulli(1/2) words [B,T] fp32 0/1 from Philox4x32-10 (the role of word_rand_gen.randint, channel_dataset.py:67, for
 * Monte-Carlo runs generated on the device). */
int mvn_random_bits(float *bits, int64_t B, int T, uint64_t seed, void *stream);

/* ---- f3: true-MLSE decoding by survivor traceback.  survivors [B,n_stages,max(1,S/64)] uint32 and final_pm [B,S] are the
 * optional outputs of mvn_acs_decode (trellis_utils.py:30's indices, which the reference detectors discard).
 * start_state < 0: trace back from the best final state (lowest index on ties); >= 0: from that state (0 for the
 * reference's zero-padded words).  decoded as in mvn_acs_decode (columns >= n_stages are 0). */
int mvn_traceback(const uint32_t *survivors, const float *final_pm, int64_t B, int T, int n_stages, int L, int start_state,
                  int out_format, void *decoded, void *stream);

/* Branch metrics of the full-CSI Viterbi as a tensor cost [B,T,S] (va_detector.py:62-68), for the MLSE path. */
int mvn_va_cost(const float *y, int64_t B, int T, int L, const float *state_priors, int n_h, float *cost, void *stream);

/* ---- f2: Reed-Solomon over GF(2^8), primitive polynomial 0x11d, generator roots 2^0 .. 2^(nsym-1)
 * (ecc/rs_main.py:9-37, rs_encoder.py:7-37, rs_decoder.py:37-218).  Words are rows of bits stored as fp32 0/1 (the
 * detectors' output format; any nonzero value counts as 1), 8 bits per byte, most significant bit first
 * (numpy packbits, polynomials_manipulation.py:119-125).  nsym = 1..32 parity bytes, n_bytes = k_bytes + nsym <= 255.
 *
 * mvn_rs_decode replaces rs_main.decode per detected word (trainer.py:234-236, :298): rx_bits [B, ld_in] with
 * 8*n_bytes used columns -> msg_bits [B, ld_out], 8*(n_bytes-nsym) columns written.  Bit-exact with the reference also
 * beyond the code's capacity: when Berlekamp-Massey claims more than nsym/2 errors the received message bytes are
 * returned unchanged, when the locator has fewer roots than its degree the partial correction is applied.
 * status [B] (optional): 0 clean, 1 corrected, 2 too many errors (unchanged), 3 missing roots (partial correction).
 *
 * mvn_rs_encode replaces rs_main.encode (channel_dataset.py:51, trainer.py:286,304,315): msg_bits [B, ld_in] with
 * 8*k_bytes used columns -> cw_bits [B, ld_out], 8*(k_bytes+nsym) columns: message then parity. */
int mvn_rs_decode(const float *rx_bits, int64_t B, int ld_in, int n_bytes, int nsym, float *msg_bits, int ld_out,
                  int32_t *status, void *stream);
int mvn_rs_encode(const float *msg_bits, int64_t B, int ld_in, int k_bytes, int nsym, float *cw_bits, int ld_out,
                  void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MVN_B200_NEXT_H */
