// Speech-presence-probability guide labels ("timo" label source): replaces timo_mask_estimation on top of
// SPPNoiseEstimator.update (reference python/models/spp_estimation.py:198-218, :84-141; call site
// scripts/evaluate_M2_ibm.py:136-141).  Per frequency bin the estimator is a first-order recursion over the
// frames of one utterance -- noise PSD tracking with a fixed a-priori SNR, the SPP of the frame from the
// generalised likelihood ratio, a smoothed SPP as stuck protection -- so the parallel axes are (utterance,
// bin) and the frame loop stays sequential.  A warp owns 32 bins; it loads 32 frames x 32 bins through shared
// memory so that the global reads run along the frame axis (coalesced) while each thread consumes its own bin.
// The reference computes in float64 (its state vectors are float64), so does this kernel: the hard label is a
// threshold at 0.5 and must not depend on rounding.
#include "gvn_common.cuh"

namespace gvn {

namespace {

struct SppParams { double fixed_smooth, prob_smooth, inv_glr_factor, inv_glr_exp_factor; int n_init; };

__global__ void __launch_bounds__(32) k_spp_mask(int F, int NP, const int32_t* __restrict__ frame_off,
                                                 const int32_t* __restrict__ n_frames, const float* __restrict__ X2,
                                                 SppParams q, float* __restrict__ soft, float* __restrict__ hard) {
  __shared__ float tile[32][33];
  const int lane = threadIdx.x, b = blockIdx.y, f0 = blockIdx.x * 32, f = f0 + lane;
  const int n0 = frame_off[b], N = n_frames[b];
  double old_psd = 0.0, smooth = 0.0;
  for (int c0 = 0; c0 < N; c0 += 32) {
    const int cn = min(32, N - c0);
    for (int r = 0; r < 32; ++r)                                  // row r = bin f0+r, lane = frame
      tile[r][lane] = (f0 + r < F && lane < cn) ? X2[(size_t)(f0 + r) * NP + n0 + c0 + lane] : 0.f;
    __syncwarp();
#pragma unroll 1
    for (int i = 0; i < cn; ++i) {
      const double per = (double)tile[lane][i];
      double spp;
      if (c0 + i < q.n_init) {                                    // spp_estimation.py:98-108: average of the first frames
        old_psd = old_psd + (double)(tile[lane][i] / (float)q.n_init);      // float32 periodogram / int stays float32 in numpy
        spp = 0.0;
      } else {                                                    // :110-133
        const double inv_glr = q.inv_glr_factor * exp(-per / (old_psd + 1e-8) * q.inv_glr_exp_factor);
        spp = 1.0 / (1.0 + inv_glr);
        smooth = (1.0 - q.prob_smooth) * spp + q.prob_smooth * smooth;
        if (smooth > 0.99) spp = fmin(spp, 0.99);
        const double noise_per = (1.0 - spp) * per + spp * old_psd;
        old_psd = (1.0 - q.fixed_smooth) * noise_per + q.fixed_smooth * old_psd;
      }
      tile[lane][i] = (float)spp;                                 // reuse the tile for the transposed write-back
    }
    __syncwarp();
    for (int r = 0; r < 32; ++r) {
      if (f0 + r < F && lane < cn) {
        const float v = tile[r][lane];
        const size_t o = (size_t)(f0 + r) * NP + n0 + c0 + lane;
        if (soft != nullptr) soft[o] = v;
        if (hard != nullptr) hard[o] = v > 0.5f ? 1.f : 0.f;        // evaluate_M2_ibm.py:139 (on the float32 mask)
      }
    }
    __syncwarp();
  }
}

}  // namespace

int32_t launch_spp_mask(const gvn_batch* b, float fixed_smooth, float prob_smooth, float prior, float snr_opt_db,
                        int n_init, float* soft, float* hard, cudaStream_t st) {
  SppParams q;
  const double snr = pow(10.0, (double)snr_opt_db / 10.0), pr = (double)prior;
  q.fixed_smooth = (double)fixed_smooth; q.prob_smooth = (double)prob_smooth;
  q.inv_glr_factor = (1.0 - pr) / pr * (1.0 + snr);               // spp_estimation.py:80-81
  q.inv_glr_exp_factor = snr / (1.0 + snr);
  q.n_init = n_init;
  dim3 grid((b->F + 31) / 32, b->B);
  k_spp_mask<<<grid, 32, 0, st>>>(b->F, b->NP, b->frame_off, b->n_frames, b->X2, q, soft, hard);
  return check_launch("k_spp_mask");
}

}  // namespace gvn
