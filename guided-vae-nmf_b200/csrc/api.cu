// extern "C" surface of libgvn.so (declared in include/gvn.h): argument validation, error
// strings, dispatch to the kernel launchers.  No torch types, no allocation, no host sync.
#include "gvn_common.cuh"

#include <string.h>

namespace gvn {

char* error_buffer() {
  static thread_local char buf[512] = "";
  return buf;
}

int32_t fail(int32_t code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(error_buffer(), 512, fmt, ap);
  va_end(ap);
  return code;
}

static unsigned long long g_launches = 0;      // kernels launched by this library (one check_launch per launch)

int32_t check_launch(const char* what) {
  ++g_launches;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(GVN_E_CUDA, "%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
  return GVN_OK;
}

// launchers implemented in the other translation units
int32_t launch_estep_simt(const gvn_batch*, const void*, int, int, float, const gvn_noise*, const gvn_trace*, int, cudaStream_t);
int32_t launch_estep_tc(const gvn_batch*, const void*, int, int, float, const gvn_noise*, const gvn_trace*, int, cudaStream_t);
size_t mstep_workspace_bytes(const gvn_batch*);
int32_t launch_mstep(const gvn_batch*, int, float*, void*, int, cudaStream_t);
int32_t launch_mstep_gain(const gvn_batch*, int, float*, cudaStream_t);
int32_t launch_cost_reduce(const gvn_batch*, int, int, const float*, double*, cudaStream_t);
int32_t launch_spp_mask(const gvn_batch*, float, float, float, float, int, float*, float*, cudaStream_t);
int32_t launch_wiener(const gvn_batch*, int, float*, float*, float*, float*, cudaStream_t);
int32_t launch_init_nmf(const gvn_batch*, const float*, const float*, float, cudaStream_t);
int32_t launch_stft_power(const gvn_batch*, const float*, int, const int32_t*, const int32_t*, int, int, cudaStream_t);
size_t istft_workspace_bytes(const gvn_batch*, int);
int32_t launch_istft(const gvn_batch*, const float*, int, int, const int32_t*, float*, int, void*, cudaStream_t);
int32_t launch_dense(const float*, const float*, const float*, int, const float*, int, const float*, const float*, float,
                     int, int, int, float*, cudaStream_t);
void set_profile_buffer(void*);
size_t speech_labels_workspace_bytes(const gvn_batch*);
int32_t launch_speech_labels(const gvn_batch*, const float*, int, int, float, float*, void*, cudaStream_t);
int32_t launch_energy_ratios(const float*, const float*, const float*, int, int, const int32_t*, double*, cudaStream_t);
int32_t launch_selftest_umma(const float*, const float*, int, int, int, float*, cudaStream_t);
int32_t launch_pack_decoder(const float*, const float*, const float*, const float*, const float*, const float*, int, int,
                            int, void*, cudaStream_t);

static int32_t check_batch(const gvn_batch* b, bool need_state) {
  GVN_REQUIRE(b != nullptr, GVN_E_INVALID, "batch is NULL");
  GVN_REQUIRE(b->B > 0 && b->F > 0 && b->NP > 0, GVN_E_INVALID, "batch dims B=%d F=%d NP=%d", b->B, b->F, b->NP);
  GVN_REQUIRE(b->NP % GVN_FRAME_ALIGN == 0, GVN_E_INVALID, "NP=%d is not a multiple of %d", b->NP, GVN_FRAME_ALIGN);
  GVN_REQUIRE(b->frame_off && b->n_frames && b->frame_utt, GVN_E_INVALID, "frame index arrays are NULL");
  if (need_state) {
    GVN_REQUIRE(b->K >= 1 && b->K <= GVN_MAX_K, GVN_E_UNSUPPORTED_SHAPE, "NMF rank K=%d outside [1,%d]", b->K, GVN_MAX_K);
    GVN_REQUIRE(b->X2 && b->W && b->Wun && b->H && b->g && b->Vb && b->Vs && b->Vs_w, GVN_E_INVALID, "batch state pointer is NULL");
  }
  return GVN_OK;
}

}  // namespace gvn

using namespace gvn;

extern "C" {

int32_t gvn_version(void) { return GVN_VERSION; }
const char* gvn_last_error(void) { return error_buffer(); }

size_t gvn_decoder_packed_bytes(int32_t L, int32_t y_dim, int32_t F, int32_t hidden) {
  if (hidden != GVN_HIDDEN || L < 1 || L > GVN_MAX_L || y_dim < 0 || F < 1) return 0;
  return decoder_layout(L, y_dim, F).total_bytes;
}

int32_t gvn_pack_decoder(const float* W1, const float* b1, const float* W2, const float* b2, const float* W3,
                         const float* b3, int32_t L, int32_t y_dim, int32_t F, int32_t hidden, void* packed,
                         void* stream) {
  GVN_REQUIRE(hidden == GVN_HIDDEN, GVN_E_UNSUPPORTED_SHAPE, "decoder hidden width %d, only %d is supported", hidden, GVN_HIDDEN);
  GVN_REQUIRE(L >= 1 && L <= GVN_MAX_L, GVN_E_UNSUPPORTED_SHAPE, "latent dim L=%d outside [1,%d]", L, GVN_MAX_L);
  GVN_REQUIRE(y_dim >= 0 && F >= 1, GVN_E_INVALID, "y_dim=%d F=%d", y_dim, F);
  GVN_REQUIRE(W1 && b1 && W2 && b2 && W3 && b3 && packed, GVN_E_INVALID, "NULL decoder pointer");
  return launch_pack_decoder(W1, b1, W2, b2, W3, b3, L, y_dim, F, packed, (cudaStream_t)stream);
}

int32_t gvn_label_projection(const void* packed, const float* y, int32_t L, int32_t y_dim, int32_t F, int32_t NP,
                             float* yproj, void* stream) {
  GVN_REQUIRE(packed && yproj, GVN_E_INVALID, "NULL pointer");
  GVN_REQUIRE(y_dim == 0 || y != nullptr, GVN_E_INVALID, "y is NULL with y_dim=%d", y_dim);
  DecoderLayout d = decoder_layout(L, y_dim, F);
  const float* p = reinterpret_cast<const float*>(packed);
  return launch_dense(p + d.w1y, p + d.b1, y, y_dim, nullptr, 0, nullptr, nullptr, 0.f, GVN_HIDDEN, NP, 0, yproj,
                      (cudaStream_t)stream);
}

int32_t gvn_estep(const gvn_batch* batch, const void* packed, int32_t burnin, int32_t R, float var_RW,
                  const gvn_noise* noise, const gvn_trace* trace, int32_t precision, void* stream) {
  int32_t rc = check_batch(batch, true);
  if (rc) return rc;
  GVN_REQUIRE(packed && noise && batch->Z && batch->yproj, GVN_E_INVALID, "NULL pointer");
  GVN_REQUIRE(batch->L >= 1 && batch->L <= GVN_MAX_L, GVN_E_UNSUPPORTED_SHAPE, "latent dim L=%d outside [1,%d]", batch->L, GVN_MAX_L);
  GVN_REQUIRE(burnin >= 0 && R >= 1 && R <= batch->R_cap, GVN_E_INVALID, "burnin=%d R=%d R_cap=%d", burnin, R, batch->R_cap);
  GVN_REQUIRE((noise->eps == nullptr) == (noise->u == nullptr), GVN_E_INVALID, "eps and u must both be given or both NULL");
  GVN_REQUIRE(var_RW > 0.f, GVN_E_INVALID, "var_RW=%g", (double)var_RW);
  const int32_t prec = precision & ~(GVN_PREC_XV_CURRENT | GVN_PREC_XV_BF16);
  if (prec == GVN_PREC_FP32)
    return launch_estep_simt(batch, packed, burnin, R, var_RW, noise, trace, precision, (cudaStream_t)stream);
  if (prec == GVN_PREC_F16)
    return launch_estep_tc(batch, packed, burnin, R, var_RW, noise, trace, precision, (cudaStream_t)stream);
  return fail(GVN_E_INVALID, "unknown precision %d", precision);
}

size_t gvn_mstep_workspace_bytes(const gvn_batch* batch) { return batch ? mstep_workspace_bytes(batch) : 0; }

int32_t gvn_mstep(const gvn_batch* batch, int32_t R, float* cost_part, void* workspace, int32_t variant, void* stream) {
  int32_t rc = check_batch(batch, true);
  if (rc) return rc;
  GVN_REQUIRE(cost_part && workspace, GVN_E_INVALID, "NULL pointer");
  GVN_REQUIRE(R >= 1 && R <= batch->R_cap, GVN_E_INVALID, "R=%d R_cap=%d", R, batch->R_cap);
  return launch_mstep(batch, R, cost_part, workspace, variant, (cudaStream_t)stream);
}

int32_t gvn_mstep_gain(const gvn_batch* batch, int32_t R, float* cost_part, void* stream) {
  int32_t rc = check_batch(batch, false);
  if (rc) return rc;
  GVN_REQUIRE(cost_part && batch->X2t && batch->Vs && batch->Vs_w && batch->Vb && batch->g, GVN_E_INVALID, "NULL pointer");
  GVN_REQUIRE(R >= 1 && R <= batch->R_cap, GVN_E_INVALID, "R=%d R_cap=%d", R, batch->R_cap);
  return launch_mstep_gain(batch, R, cost_part, (cudaStream_t)stream);
}

int32_t gvn_cost_reduce(const gvn_batch* batch, int32_t R, int32_t niter, const float* cost_part, double* cost,
                        void* stream) {
  int32_t rc = check_batch(batch, false);
  if (rc) return rc;
  GVN_REQUIRE(cost_part && cost && niter >= 1 && R >= 1, GVN_E_INVALID, "bad argument");
  return launch_cost_reduce(batch, R, niter, cost_part, cost, (cudaStream_t)stream);
}

int32_t gvn_wiener(const gvn_batch* batch, int32_t R, float* S_hat, float* N_hat, float* WFs, float* WFn, void* stream) {
  int32_t rc = check_batch(batch, true);
  if (rc) return rc;
  GVN_REQUIRE(S_hat && N_hat && batch->Xc, GVN_E_INVALID, "NULL pointer");
  GVN_REQUIRE(R >= 1 && R <= batch->R_cap, GVN_E_INVALID, "R=%d R_cap=%d", R, batch->R_cap);
  return launch_wiener(batch, R, S_hat, N_hat, WFs, WFn, (cudaStream_t)stream);
}

int32_t gvn_stft_power(const gvn_batch* batch, const float* wav, int32_t T_stride, const int32_t* T,
                       const int32_t* end_pad, int32_t n_fft, int32_t hop, void* stream) {
  int32_t rc = check_batch(batch, false);
  if (rc) return rc;
  GVN_REQUIRE(wav && T && end_pad && batch->X2 && batch->Xc, GVN_E_INVALID, "NULL pointer");
  GVN_REQUIRE(hop >= 1 && hop <= n_fft, GVN_E_INVALID, "hop=%d n_fft=%d", hop, n_fft);
  return launch_stft_power(batch, wav, T_stride, T, end_pad, n_fft, hop, (cudaStream_t)stream);
}

size_t gvn_istft_workspace_bytes(const gvn_batch* batch, int32_t n_fft) {
  return batch ? istft_workspace_bytes(batch, n_fft) : 0;
}

int32_t gvn_istft(const gvn_batch* batch, const float* S, int32_t n_fft, int32_t hop, const int32_t* out_len,
                  float* out, int32_t T_stride, void* workspace, void* stream) {
  int32_t rc = check_batch(batch, false);
  if (rc) return rc;
  GVN_REQUIRE(S && out_len && out && workspace, GVN_E_INVALID, "NULL pointer");
  GVN_REQUIRE(hop >= 1 && hop <= n_fft, GVN_E_INVALID, "hop=%d n_fft=%d", hop, n_fft);
  return launch_istft(batch, S, n_fft, hop, out_len, out, T_stride, workspace, (cudaStream_t)stream);
}

int32_t gvn_dense(const float* W, const float* b, const float* in0, int32_t D0, const float* in1, int32_t D1,
                  const float* mean, const float* std_, float eps, int32_t D_out, int32_t NP, int32_t act, float* out,
                  void* stream) {
  GVN_REQUIRE(W && out && D_out >= 1 && NP >= 1 && D0 >= 0 && D1 >= 0, GVN_E_INVALID, "bad argument");
  GVN_REQUIRE((D0 == 0 || in0) && (D1 == 0 || in1), GVN_E_INVALID, "NULL input");
  GVN_REQUIRE((mean == nullptr) == (std_ == nullptr), GVN_E_INVALID, "mean and std must both be given or both NULL");
  GVN_REQUIRE(act >= 0 && act <= 4, GVN_E_INVALID, "act=%d", act);
  return launch_dense(W, b, in0, D0, in1, D1, mean, std_, eps, D_out, NP, act, out, (cudaStream_t)stream);
}

int32_t gvn_energy_ratios(const float* est, const float* s, const float* n, int32_t B, int32_t T_stride, const int32_t* T,
                          double* out, void* stream) {
  GVN_REQUIRE(est && s && n && T && out, GVN_E_INVALID, "NULL pointer");
  GVN_REQUIRE(B > 0 && T_stride > 0, GVN_E_INVALID, "B=%d T_stride=%d", B, T_stride);
  return launch_energy_ratios(est, s, n, B, T_stride, T, out, (cudaStream_t)stream);
}

void gvn_debug_profile_buffer(void* dev_u64) { set_profile_buffer(dev_u64); }
uint64_t gvn_launch_count(void) { return g_launches; }

int32_t gvn_selftest_umma(const float* A, const float* W, int32_t N, int32_t K, int32_t variant, float* D, void* stream) {
  GVN_REQUIRE(A && W && D, GVN_E_INVALID, "NULL pointer");
  return launch_selftest_umma(A, W, N, K, variant, D, (cudaStream_t)stream);
}

int32_t gvn_spp_mask(const gvn_batch* batch, float fixed_smooth, float prob_smooth, float prior, float snr_opt_db,
                     int32_t n_init, float* soft, float* hard, void* stream) {
  int32_t rc = check_batch(batch, false);
  if (rc) return rc;
  GVN_REQUIRE(batch->X2 && (soft || hard), GVN_E_INVALID, "NULL pointer");
  GVN_REQUIRE(prior > 0.f && prior < 1.f && n_init >= 0, GVN_E_INVALID, "prior=%g n_init=%d", (double)prior, n_init);
  return launch_spp_mask(batch, fixed_smooth, prob_smooth, prior, snr_opt_db, n_init, soft, hard, (cudaStream_t)stream);
}

size_t gvn_speech_labels_workspace_bytes(const gvn_batch* batch) { return batch ? speech_labels_workspace_bytes(batch) : 0; }

int32_t gvn_speech_labels(const gvn_batch* batch, const float* S, int32_t from_power, int32_t vad, float quantile_fraction,
                          float quantile_weight, float* y, void* workspace, void* stream) {
  int32_t rc = check_batch(batch, false);
  if (rc) return rc;
  GVN_REQUIRE(S && y && workspace, GVN_E_INVALID, "NULL pointer");
  GVN_REQUIRE(quantile_fraction > 0.f && quantile_fraction <= 1.f, GVN_E_INVALID, "quantile_fraction=%g outside (0,1]", (double)quantile_fraction);
  GVN_REQUIRE(quantile_weight > 0.f && quantile_weight <= 1.f, GVN_E_INVALID, "quantile_weight=%g outside (0,1]", (double)quantile_weight);
  return launch_speech_labels(batch, S, from_power, vad, quantile_fraction, y, workspace, (cudaStream_t)stream);
}

int32_t gvn_init_nmf(const gvn_batch* batch, const float* rand_W, const float* rand_H, float eps, void* stream) {
  int32_t rc = check_batch(batch, true);
  if (rc) return rc;
  GVN_REQUIRE(rand_W && rand_H, GVN_E_INVALID, "NULL pointer");
  return launch_init_nmf(batch, rand_W, rand_H, eps, (cudaStream_t)stream);
}

}  // extern "C"
