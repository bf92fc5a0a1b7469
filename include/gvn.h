/* gvn.h -- C ABI of libgvn.so: the B200 (sm_100a) implementation of the MCEM-NMF
 * speech-enhancement hot path of sp-uhh/guided-vae-nmf.
 *
 * This is the drop-in boundary.  The reference is pure Python/torch, so the binding a
 * maintainer adds is a ctypes stub; every entry point below names the
 * reference code it replaces (file:line relative to the reference tree).
 *
 * Conventions
 *  - extern "C", plain pointers and sizes; no torch / C++ types.
 *  - every `*` argument not marked HOST is a DEVICE pointer owned by the caller
 *    (torch-allocated); the library never allocates or frees user-visible memory.
 *  - `stream` is a cudaStream_t passed as void*; all calls are asynchronous on it and never
 *    synchronise the host.
 *  - return value: 0 = OK, <0 = error (GVN_E_*); the message is in gvn_last_error()
 *    (thread-local).  The Python host maps GVN_E_UNSUPPORTED_MODEL to NameError and
 *    GVN_E_BAD_WINDOW to ValueError, the two exceptions the reference raises on this path
 *    (python/models/mcem.py:208-209, python/processing/stft.py:37-38).
 *
 * Batch layout in HBM ("frame-minor", the reference's own (F,N) orientation)
 *  B utterances are laid side by side on one global frame axis of length NP (a multiple
 *  of GVN_FRAME_ALIGN).  Utterance b owns global frames [frame_off[b], frame_off[b] +
 *  n_frames[b]); frame_off[b] is a multiple of GVN_FRAME_ALIGN; the gap up to
 *  frame_off[b+1] is padding that kernels skip (frame_utt[n] == -1).
 *    X2   [F][NP]  f32   |X|^2                     (mcem.py:47  X_abs_2)
 *    Xc   [F][NP]  c64   mixture STFT              (mcem.py:46  X)
 *    W    [B][F][K] f32  normalised dictionary     (mcem.py:48,131)
 *    Wun  [B][F][K] f32  scratch: W before the column normalisation (mcem.py:110)
 *    H    [K][NP]  f32   activations               (mcem.py:49,133)
 *    g    [NP]     f32   gain                      (mcem.py:51,142)
 *    Vb   [F][NP]  f32   noise variance W@H        (mcem.py:82; NOT refreshed after the
 *                                                   normalisation, as in mcem.py:124-133)
 *    Z    [L][NP]  f32   current latent state      (mcem.py:215, 319)
 *    Vs   [R][NP/8][F][8] f32 speech variance of the kept samples (mcem.py:307) in column-tile
 *                        order: element (r, f, n) lives at ((r*(NP/8) + n/8)*F + f)*8 + n%8, so the
 *                        F x 8 block of one 8-frame column tile is contiguous (16 KB at F=513) --
 *                        what the M-step stages per tile -- and a row of it is one 32-byte sector.
 *                        Slot form: slot r
 *                        holds the decoder output of the PROPOSAL of kept step r (slot 0: the state
 *                        after the burn-in) and
 *    Vs_w [R][NP]    f32 its multiplicity: 0 when that proposal was rejected, 1 + (number of
 *                        following rejected steps) otherwise.  A Metropolis-Hastings reject repeats
 *                        the previous sample (mcem.py:280-289), so sum_r phi(Vs_r) of the reference
 *                        equals sum_slot Vs_w[slot] * phi(Vs[slot]); nothing is ever copied.
 *                        sum_slot Vs_w[slot][n] == R for every frame.  Every stored value is finite (<= GVN_VS_MAX),
 *                        also in slots of multiplicity 0.
 *    XV   [F][NP]    u32 per-bin constants of the tensor-core chain, one word per (f, n): low half =
 *                        bf16(Vb), high half such that the whole word read as f32 is nearest to X2.
 *                        Written by gvn_estep (unless GVN_PREC_XV_CURRENT) and kept current by gvn_mstep.
 *    X2t  [NP/8][F][8] f32 X2 in column-tile order (built by gvn_init_nmf; read by gvn_mstep)
 *    yproj[HID][NP] f32  b1 + W1[:, L:] @ y  -- the label part of the decoder's first layer,
 *                        constant per utterance (mcem.py:242 concatenates y every step)
 */
#ifndef GVN_H_
#define GVN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GVN_VERSION 100
#define GVN_FRAME_ALIGN 32      /* utterances start on multiples of this many frames   */
#define GVN_VS_TILE 8          /* frames per column block of Vs / X2t                    */
#define GVN_COST_TILE 8        /* frames per partial cost sum written by gvn_mstep (= GVN_VS_TILE) */
#define GVN_HIDDEN 128          /* decoder hidden width (h_dim=[128,128] in every script) */
#define GVN_MAX_K 32            /* NMF rank limit */
#define GVN_MAX_L 64            /* latent dimension limit */
#define GVN_MAX_R_SLOTS 128     /* sample-slot limit of gvn_mstep_gain */
#define GVN_VS_MAX 1e18f        /* gvn_estep clamps every speech variance it stores to this: a slot whose proposal was
                                 * rejected because its decoder output overflowed (Vs_w = 0) must not put inf * 0 = NaN
                                 * into the multiplicity-weighted sums of gvn_mstep / gvn_wiener */

enum {
  GVN_OK = 0,
  GVN_E_INVALID = -1,           /* bad argument (null pointer, size out of range)      */
  GVN_E_UNSUPPORTED_SHAPE = -2, /* hidden != 128, K > 32, n_fft not a power of two ... */
  GVN_E_CUDA = -3,              /* a CUDA runtime call failed; message has the code    */
  GVN_E_UNSUPPORTED_MODEL = -4, /* recurrent VAE (mcem.py:208-209 raises NameError)     */
  GVN_E_BAD_WINDOW = -5         /* wlen_sec*fs not integral (stft.py:37-38 ValueError)  */
};

/* arithmetic of the decoder contraction inside gvn_estep */
enum {
  GVN_PREC_FP32 = 0,   /* CUDA-core fp32 FMA (bit-faithful mode, parity rtol 1e-4)           */
  /* 1 is not assigned (a hi/lo-split fp32-parity tensor-core mode was planned under this value and never built:
   * W3 as an f16 hi + lo pair is 270 KB, more than one SM's shared memory) */
  GVN_PREC_F16 = 2,    /* tcgen05 f16 operands, fp32 accumulate (11-bit mantissa, = TF32)    */
  /* flag, OR-ed into `precision`: batch->XV already matches (X2, Vb).  gvn_mstep rewrites XV next to
   * every Vb it writes, so inside the EM loop only the first gvn_estep has to pack it (the packing
   * pass reads X2 and Vb and writes XV: 100 MB per launch at 64 utterances). */
  GVN_PREC_XV_CURRENT = 0x100,
  /* flag for GVN_PREC_FP32 (diagnostic): the CUDA-core chain reads X2 and Vb through the bf16 rounding of XV, i.e.
   * exactly the two constants the tensor-core chain sees, everything else in fp32 -- the A/B that isolates the
   * effect of that rounding (tests/test_gpu_fullshape.py). */
  GVN_PREC_XV_BF16 = 0x200
};

typedef struct gvn_batch {
  int32_t B, F, K, L, NP, R_cap;  /* R_cap: number of sample slots allocated in Vs          */
  const int32_t* frame_off;       /* [B+1]                                                  */
  const int32_t* n_frames;        /* [B]                                                    */
  const int32_t* frame_utt;       /* [NP] utterance of each global frame, -1 = padding      */
  float* X2;
  float* Xc;                      /* interleaved re,im                                      */
  float* W;
  float* Wun;
  float* H;
  float* g;
  float* Vb;
  float* Z;
  float* Vs;
  const float* yproj;
  float* Vs_w;
  uint32_t* XV;                   /* may be NULL when only GVN_PREC_FP32 is used            */
  float* X2t;
} gvn_batch;

/* random input of one Metropolis-Hastings chain (mcem.py:257 randn(L,N), :271 rand(N)).
 * Replay mode: eps/u point at recorded draws (parity tests).  Otherwise both are NULL and
 * the kernel draws from Philox4x32-10 keyed by (seed, chain) with counter (frame, step). */
typedef struct gvn_noise {
  const float* eps;               /* [n_steps][L][NP] or NULL                               */
  const float* u;                 /* [n_steps][NP]    or NULL                               */
  const uint8_t* forced_accept;   /* [n_steps][NP] or NULL: override the decision (tests)   */
  uint64_t seed;
  uint64_t chain;
} gvn_noise;

/* optional per-step outputs of a chain (NULL to skip) */
typedef struct gvn_trace {
  float* acc_prob;                /* [n_steps][NP] log acceptance ratio (mcem.py:266-268)   */
  uint8_t* accepted;              /* [n_steps][NP] decision taken       (mcem.py:271)       */
  int32_t* n_accepted;            /* [NP] running count per frame (mcem.py:276-277)         */
  float* z_samples;               /* [R][L][NP] kept latent samples (mcem.py:287 Z_sampled) */
} gvn_trace;

int32_t gvn_version(void);
const char* gvn_last_error(void);

/* Decoder weights (models.py:107-121; state-dict keys decoder.hidden.{0,1}, decoder.
 * reconstruction).  W1 is (HID, L+y_dim) row-major as nn.Linear stores it.  The packed
 * image holds the transposed fp32 copies for the CUDA-core path and the f16 UMMA operand
 * images for the tensor-core path. */
size_t gvn_decoder_packed_bytes(int32_t L, int32_t y_dim, int32_t F, int32_t hidden);
int32_t gvn_pack_decoder(const float* W1, const float* b1, const float* W2, const float* b2,
                         const float* W3, const float* b3, int32_t L, int32_t y_dim, int32_t F,
                         int32_t hidden, void* packed, void* stream);

/* yproj = b1 + W1[:, L:] @ y   (y: [y_dim][NP], may be NULL when y_dim == 0). */
int32_t gvn_label_projection(const void* packed, const float* y, int32_t L, int32_t y_dim, int32_t F,
                             int32_t NP, float* yproj, void* stream);

/* One MH chain for every frame of the batch: replaces sample_posterior + compute_Vs
 * (mcem.py:218-307 / :371-454).  n_steps = burnin + R.  On return Z holds the last kept
 * sample (mcem.py:319) and Vs[0..R) / Vs_w[0..R) the speech variance of the kept samples in
 * slot form (see the layout comment above). */
int32_t gvn_estep(const gvn_batch* batch /*HOST*/, const void* packed, int32_t burnin, int32_t R,
                  float var_RW, const gvn_noise* noise /*HOST*/, const gvn_trace* trace /*HOST or NULL*/,
                  int32_t precision, void* stream);

/* NMF / gain M-step: replaces EM.M_step + compute_expected_neg_log_like
 * (mcem.py:90-152, :68-70).  Updates W, H, g, Vb in place; cost_part receives one partial
 * sum per GVN_COST_TILE frames ([NP/8]); gvn_cost_reduce turns niter of them into cost[niter][B].
 * variant 0: straightforward schedule (IEEE divisions, logf), any shape.
 * variant 1 (default): bulk-copy W sweep + column sweep with the (R+1) x F x 8 tile staged once in
 * shared memory (K <= 12 at R = 10), or the generic L2-resident column sweep (K <= 32, R <= ~90);
 * falls back to 0 when neither fits. */
size_t gvn_mstep_workspace_bytes(const gvn_batch* batch /*HOST*/);
int32_t gvn_mstep(const gvn_batch* batch /*HOST*/, int32_t R, float* cost_part, void* workspace,
                  int32_t variant, void* stream);
int32_t gvn_cost_reduce(const gvn_batch* batch /*HOST*/, int32_t R, int32_t niter,
                        const float* cost_part /*[niter][NP/8]*/, double* cost /*[niter][B]*/,
                        void* stream);

/* Gain-only M-step of the models without an NMF noise dictionary: replaces EM_noNMF.M_step +
 * compute_expected_neg_log_like (mcem.py:551-588, :530-532; used by MCEM_M2_noNMF :609-760).  batch->Vb is an
 * input that stays fixed; updates g in place and writes the cost partials like gvn_mstep (same gvn_cost_reduce).
 * Reads X2t, Vs, Vs_w, Vb, g; W / H are not touched. */
int32_t gvn_mstep_gain(const gvn_batch* batch /*HOST*/, int32_t R, float* cost_part, void* stream);

/* Wiener filter from the R kept samples of the final chain: replaces the tail of
 * compute_WF and EM.run (mcem.py:341-343, :175-176).  S_hat/N_hat are [F][NP] c64;
 * WFs/WFn ([F][NP] f32) are optional. */
int32_t gvn_wiener(const gvn_batch* batch /*HOST*/, int32_t R, float* S_hat, float* N_hat,
                   float* WFs, float* WFn, void* stream);

/* STFT of B zero-padded waveforms wav[B][T_stride] (f32) with true lengths T[b]:
 * replaces python/processing/stft.py:16-63 (periodic Hann, reflect centring, the end-pad
 * rule of :48-53 applied when end_pad[b] != 0).  Writes Xc and X2 columns of the batch. */
int32_t gvn_stft_power(const gvn_batch* batch /*HOST*/, const float* wav, int32_t T_stride,
                       const int32_t* T, const int32_t* end_pad, int32_t n_fft, int32_t hop,
                       void* stream);

/* ISTFT: replaces python/processing/stft.py:66-102 (overlap-add, window-sum-square
 * normalisation, drop n_fft/2, pad/trim to out_len[b]).  S is [F][NP] c64; out is
 * out[B][T_stride] f32; workspace holds the windowed inverse frames. */
size_t gvn_istft_workspace_bytes(const gvn_batch* batch /*HOST*/, int32_t n_fft);
int32_t gvn_istft(const gvn_batch* batch /*HOST*/, const float* S, int32_t n_fft, int32_t hop,
                  const int32_t* out_len, float* out, int32_t T_stride, void* workspace, void* stream);

/* One dense layer on feature-major activations: out[j][n] = act(b[j] + sum_i W[j][i] *
 * in[i][n]) where the input rows are the concatenation [in0 (D0 rows); in1 (D1 rows)] and
 * in0 may be standardised on load ((x-mean)/(std+eps), scripts/evaluate_M2_ibm.py:121-125).
 * Replaces Encoder / Classifier forward (models.py:90-104, :41-62).  act: 0 none, 1 tanh,
 * 2 relu, 3 sigmoid, 4 sigmoid>0.5 (hard label). */
int32_t gvn_dense(const float* W, const float* b, const float* in0, int32_t D0, const float* in1,
                  int32_t D1, const float* mean, const float* std_, float eps, int32_t D_out,
                  int32_t NP, int32_t act, float* out, void* stream);

/* "timo" guide labels: the speech presence probability of every time-frequency bin of batch->X2 from the
 * SPP-based noise tracker.  Replaces timo_mask_estimation / SPPNoiseEstimator.update
 * (python/models/spp_estimation.py:198-218, :84-141; call site scripts/evaluate_M2_ibm.py:136-141).  The reference's
 * defaults are fixed_smooth 0.8, prob_smooth 0.9, prior 0.5, snr_opt_db 15, n_init 10 (:10-14).  soft / hard are
 * [F][NP] f32 (either may be NULL): the mask and its threshold at 0.5. */
int32_t gvn_spp_mask(const gvn_batch* batch /*HOST*/, float fixed_smooth, float prob_smooth, float prior,
                     float snr_opt_db, int32_t n_init, float* soft, float* hard, void* stream);

/* Oracle guide labels from the clean-speech STFT: replaces clean_speech_IBM / clean_speech_VAD
 * (python/processing/target.py:7-27, :29-50; call sites scripts/evaluate_M2_ibm.py:132-134).  Per utterance: power
 * |S conj(S)| (vad != 0: summed over frequency per frame), sorted descending, threshold = the last value whose share
 * cumsum / sum is below quantile_fraction, label = power > threshold.  The float32 arithmetic follows numpy's order
 * (sequential cumsum, pairwise sum, FMA product), so the labels equal the reference's bit for bit on the same S.
 * S: [F][NP] c64 (from_power == 0) or the power itself, [F][NP] f32 (from_power != 0); y: [F][NP] f32 (vad == 0) or
 * [NP] f32 (vad != 0), 0 on padding frames.  quantile_weight only softens the mask before it is rounded back
 * (target.py:22-24): any value in (0, 1] gives the same labels; others are rejected. */
size_t gvn_speech_labels_workspace_bytes(const gvn_batch* batch /*HOST*/);
int32_t gvn_speech_labels(const gvn_batch* batch /*HOST*/, const float* S, int32_t from_power, int32_t vad,
                          float quantile_fraction, float quantile_weight, float* y, void* workspace, void* stream);

/* NMF initialisation W = max(rand, eps), H = max(rand, eps), g = 1, Vb = W@H, from caller
 * supplied uniforms (mcem.py:36-57); also fills padding frames with benign values. */
int32_t gvn_init_nmf(const gvn_batch* batch /*HOST*/, const float* rand_W /*[B][F][K]*/,
                     const float* rand_H /*[K][NP]*/, float eps, void* stream);

/* Quality metrics of B enhanced signals against their clean-speech and noise references:
 * replaces energy_ratios / si_sdr_components (python/metrics.py:12-60).  est, s, n are
 * [B][T_stride] f32 with true lengths T[b]; out is [B][3] f64 = (SI-SDR, SI-SIR, SI-SAR) in dB. */
int32_t gvn_energy_ratios(const float* est, const float* s, const float* n, int32_t B, int32_t T_stride,
                          const int32_t* T, double* out, void* stream);

/* Hardware self test of the tensor-core plumbing used by gvn_estep in the f16 modes:
 * D[128][N] = A[128][K] @ W[N][K]^T through tcgen05.st (A -> TMEM), a packed shared-memory
 * image of W, tcgen05.mma and tcgen05.ld.  variant bit2 selects the 3-term hi/lo split. */
int32_t gvn_selftest_umma(const float* A, const float* W, int32_t N, int32_t K, int32_t variant,
                          float* D, void* stream);

/* Debug: when set to a device buffer of [n_tiles][18][16] u64, the tensor-core chain kernel writes
 * per-warp cycle counters of its phases (tools/estep_phases.py); NULL switches it off. */
void gvn_debug_profile_buffer(void* dev_u64);

/* Number of kernels this library has launched in the calling process so far. */
uint64_t gvn_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* GVN_H_ */
