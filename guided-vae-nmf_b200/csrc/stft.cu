// STFT power spectrum and ISTFT, framing / window / |.|^2 / overlap-add fused around a
// shared-memory radix-2 FFT (n_fft a power of two, <= 2048).
//
// Replaces python/processing/stft.py:16-63 and :66-102 of the reference, i.e. the librosa
// semantics those wrappers select: reflect centring by n_fft/2, periodic Hann, rFFT without
// scaling; inverse: irFFT * window, overlap-add, division by the overlap-added squared
// window where it exceeds `tiny`, drop n_fft/2 samples, pad / trim to the requested length.
#include "gvn_common.cuh"

namespace gvn {

namespace {

constexpr int FFT_THREADS = 256;
constexpr int FPB = 8;                  // frames per CTA pass: stores cover whole 32-byte sectors of the frame-minor arrays

__device__ __forceinline__ float hann_periodic(int i, int n) { return 0.5f - 0.5f * cospif(2.0f * (float)i / (float)n); }

__device__ __forceinline__ int bit_reverse(int x, int bits) { return (int)(__brev((unsigned)x) >> (32 - bits)); }

// twiddle table tw[k] = exp(-2 pi i k / n), k < n/2, once per CTA
__device__ void fft_twiddles(float2* tw, int n) {
  for (int k = threadIdx.x; k < n / 2; k += FFT_THREADS) {
    float sn, cs;
    sincospif(-2.0f * (float)k / (float)n, &sn, &cs);
    tw[k] = make_float2(cs, sn);
  }
}

// in-place radix-2 decimation-in-time FFTs of FPB bit-reversed sequences s[q][n] side by side, forward transform
__device__ void fft_inplace(float2* s, const float2* __restrict__ tw, int n, int bits) {
  const int nb = n / 2 * FPB;                                   // butterflies per stage
  for (int st = 1; st <= bits; ++st) {
    const int half = 1 << (st - 1), tstep = n >> st;
    for (int idx = threadIdx.x; idx < nb; idx += FFT_THREADS) {
      const int q = idx / (n / 2), bi = idx - q * (n / 2);
      const int grp = bi >> (st - 1), pos = bi & (half - 1);
      float2* sq = s + (size_t)q * n;
      const int i = (grp << st) + pos, j = i + half;
      const float2 w = tw[pos * tstep];
      const float2 a = sq[i], b = sq[j];
      const float2 t = make_float2(w.x * b.x - w.y * b.y, w.x * b.y + w.y * b.x);
      sq[i] = make_float2(a.x + t.x, a.y + t.y);
      sq[j] = make_float2(a.x - t.x, a.y - t.y);
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(FFT_THREADS) k_stft_power(int F, int NP, int n_fft, int bits, int hop,
                                                           const int32_t* __restrict__ frame_utt,
                                                           const int32_t* __restrict__ frame_off,
                                                           const float* __restrict__ wav, int T_stride,
                                                           const int32_t* __restrict__ T, const int32_t* __restrict__ end_pad,
                                                           float2* __restrict__ Xc, float* __restrict__ X2) {
  extern __shared__ float2 s_fft[];                             // [FPB][n_fft] | twiddles [n_fft/2]
  float2* tw = s_fft + (size_t)FPB * n_fft;
  fft_twiddles(tw, n_fft);
  for (int g0 = blockIdx.x * FPB; g0 < NP; g0 += gridDim.x * FPB) {
    const int b = frame_utt[g0];                                // groups of 8 frames never straddle utterances
    if (b < 0) continue;                                        // uniform per CTA
    const int Tb = T[b], Tx = Tb + (end_pad[b] ? hop : 0);
    const float* x = wav + (size_t)b * T_stride;
    for (int e = threadIdx.x; e < FPB * n_fft; e += FFT_THREADS) {
      const int q = e / n_fft, i = e - q * n_fft, j = g0 + q - frame_off[b];
      int idx = j * hop + i - n_fft / 2;
      if (idx < 0) idx = -idx;                                  // numpy 'reflect' (no edge repeat)
      if (idx >= Tx) idx = 2 * (Tx - 1) - idx;
      const float v = (idx >= 0 && idx < Tb) ? x[idx] : 0.f;
      s_fft[(size_t)q * n_fft + bit_reverse(i, bits)] = make_float2(v * hann_periodic(i, n_fft), 0.f);
    }
    __syncthreads();
    fft_inplace(s_fft, tw, n_fft, bits);
    for (int e = threadIdx.x; e < F * FPB; e += FFT_THREADS) {
      const int f = e / FPB, q = e - f * FPB;                   // consecutive threads -> consecutive frames
      if (frame_utt[g0 + q] < 0) continue;
      const float2 v = s_fft[(size_t)q * n_fft + f];
      const size_t o = (size_t)f * NP + g0 + q;
      Xc[o] = v;
      X2[o] = v.x * v.x + v.y * v.y;
    }
    __syncthreads();
  }
}

// inverse frames: ws[gn][i] = window[i] * irfft(S[:, gn])[i]
__global__ void __launch_bounds__(FFT_THREADS) k_istft_frames(int F, int NP, int n_fft, int bits,
                                                             const int32_t* __restrict__ frame_utt,
                                                             const float2* __restrict__ S, float* __restrict__ ws) {
  extern __shared__ float2 s_fft[];
  float2* tw = s_fft + (size_t)FPB * n_fft;
  fft_twiddles(tw, n_fft);
  for (int g0 = blockIdx.x * FPB; g0 < NP; g0 += gridDim.x * FPB) {
    if (frame_utt[g0] < 0) continue;
    // ifft(Y) = conj(fft(conj(Y)))/n with Y the Hermitian extension; the imaginary parts of
    // the DC and Nyquist bins are ignored, as numpy's irfft does
    for (int e = threadIdx.x; e < n_fft * FPB; e += FFT_THREADS) {
      const int k = e / FPB, q = e - k * FPB;                   // consecutive threads -> consecutive frames
      float2 v;
      if (k <= n_fft / 2) {
        v = S[(size_t)k * NP + g0 + q];
        v.y = (k == 0 || k == n_fft / 2) ? 0.f : -v.y;          // conj(Y[k])
      } else {
        v = S[(size_t)(n_fft - k) * NP + g0 + q];               // conj(conj(S)) = S
      }
      s_fft[(size_t)q * n_fft + bit_reverse(k, bits)] = v;
    }
    __syncthreads();
    fft_inplace(s_fft, tw, n_fft, bits);
    const float inv_n = 1.0f / (float)n_fft;
    for (int e = threadIdx.x; e < FPB * n_fft; e += FFT_THREADS) {
      const int q = e / n_fft, i = e - q * n_fft;
      if (frame_utt[g0 + q] >= 0) ws[(size_t)(g0 + q) * n_fft + i] = s_fft[e].x * inv_n * hann_periodic(i, n_fft);
    }
    __syncthreads();
  }
}

// overlap-add + window-sum-square normalisation, one thread per output sample
__global__ void k_istft_ola(int n_fft, int hop, int T_stride, const int32_t* __restrict__ frame_off,
                            const int32_t* __restrict__ n_frames, const int32_t* __restrict__ out_len,
                            const float* __restrict__ ws, float* __restrict__ out) {
  const int b = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T_stride) return;
  float y = 0.f;
  if (t < out_len[b]) {
    const int pos = t + n_fft / 2, N = n_frames[b], off = frame_off[b];
    int j_lo = pos - n_fft + 1;
    j_lo = j_lo <= 0 ? 0 : (j_lo + hop - 1) / hop;
    int j_hi = pos / hop;
    if (j_hi > N - 1) j_hi = N - 1;
    float wss = 0.f;
    for (int j = j_lo; j <= j_hi; ++j) {
      const int i = pos - j * hop;
      const float w = hann_periodic(i, n_fft);
      y += ws[(size_t)(off + j) * n_fft + i];
      wss = fmaf(w, w, wss);
    }
    if (wss > 1.1754944e-38f) y /= wss;
  }
  out[(size_t)b * T_stride + t] = y;
}

int log2_exact(int n) {
  int b = 0;
  while ((1 << b) < n) ++b;
  return (1 << b) == n ? b : -1;
}

// ------------------------------------------------------------------------------------------
// SI-SDR / SI-SIR / SI-SAR of B enhanced signals (reference python/metrics.py:12-60): the three
// projections need six inner products; one CTA per utterance, fp64 accumulation.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_energy_ratios(const float* __restrict__ est, const float* __restrict__ s,
                                                       const float* __restrict__ n, int T_stride, const int32_t* __restrict__ T,
                                                       double* __restrict__ out) {
  __shared__ double red[6][8];
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const size_t o = (size_t)b * T_stride;
  double a[6] = {0, 0, 0, 0, 0, 0};                         // ee, es, en, ss, nn, sn
  for (int t = tid; t < T[b]; t += 256) {
    const double e = est[o + t], sv = s[o + t], nv = n[o + t];
    a[0] += e * e; a[1] += e * sv; a[2] += e * nv; a[3] += sv * sv; a[4] += nv * nv; a[5] += sv * nv;
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) a[i] += __shfl_xor_sync(0xffffffffu, a[i], d);
    if (lane == 0) red[i][warp] = a[i];
  }
  __syncthreads();
  if (tid == 0) {
    double v[6];
    for (int i = 0; i < 6; ++i) { v[i] = 0; for (int w = 0; w < 8; ++w) v[i] += red[i][w]; }
    const double ee = v[0], es = v[1], en = v[2], ss = v[3], nn = v[4], sn = v[5];
    const double as = es / ss, an = en / nn;                // s_hat = as*s + an*n + e_art
    const double p_target = as * as * ss, p_noise = an * an * nn;
    const double p_art = ee + p_target + p_noise - 2 * as * es - 2 * an * en + 2 * as * an * sn;
    const double p_res = ee - 2 * as * es + p_target;       // |e_noise + e_art|^2 = |s_hat - as*s|^2
    out[b * 3 + 0] = 10.0 * log10(p_target / p_res);
    out[b * 3 + 1] = 10.0 * log10(p_target / p_noise);
    out[b * 3 + 2] = 10.0 * log10(p_target / p_art);
  }
}

}  // namespace

int32_t launch_energy_ratios(const float* est, const float* s, const float* n, int B, int T_stride, const int32_t* T, double* out,
                             cudaStream_t st) {
  k_energy_ratios<<<B, 256, 0, st>>>(est, s, n, T_stride, T, out);
  return check_launch("k_energy_ratios");
}

int32_t launch_stft_power(const gvn_batch* b, const float* wav, int T_stride, const int32_t* T, const int32_t* end_pad,
                          int n_fft, int hop, cudaStream_t st) {
  int bits = log2_exact(n_fft);
  GVN_REQUIRE(bits >= 4 && bits <= 11, GVN_E_UNSUPPORTED_SHAPE, "n_fft=%d must be a power of two in [16,2048]", n_fft);
  GVN_REQUIRE(b->F == n_fft / 2 + 1, GVN_E_INVALID, "F=%d does not match n_fft=%d", b->F, n_fft);
  const size_t smem = ((size_t)FPB * n_fft + n_fft / 2) * sizeof(float2);
  static size_t smem_tab[GVN_MAX_DEVICES] = {0};
  size_t& smem_set = *per_device_slot(smem_tab);
  if (smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(k_stft_power, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(GVN_E_CUDA, "k_stft_power smem attr: %s", cudaGetErrorString(e));
    smem_set = smem;
  }
  int grid = b->NP / FPB < 148 * 3 ? b->NP / FPB : 148 * 3;
  k_stft_power<<<grid, FFT_THREADS, smem, st>>>(
      b->F, b->NP, n_fft, bits, hop, b->frame_utt, b->frame_off, wav, T_stride, T, end_pad,
      reinterpret_cast<float2*>(b->Xc), b->X2);
  return check_launch("k_stft_power");
}

size_t istft_workspace_bytes(const gvn_batch* b, int n_fft) { return (size_t)b->NP * n_fft * sizeof(float); }

int32_t launch_istft(const gvn_batch* b, const float* S, int n_fft, int hop, const int32_t* out_len, float* out,
                     int T_stride, void* workspace, cudaStream_t st) {
  int bits = log2_exact(n_fft);
  GVN_REQUIRE(bits >= 4 && bits <= 11, GVN_E_UNSUPPORTED_SHAPE, "n_fft=%d must be a power of two in [16,2048]", n_fft);
  GVN_REQUIRE(b->F == n_fft / 2 + 1, GVN_E_INVALID, "F=%d does not match n_fft=%d", b->F, n_fft);
  float* ws = reinterpret_cast<float*>(workspace);
  const size_t smem = ((size_t)FPB * n_fft + n_fft / 2) * sizeof(float2);
  static size_t smem_tab[GVN_MAX_DEVICES] = {0};
  size_t& smem_set = *per_device_slot(smem_tab);
  if (smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(k_istft_frames, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(GVN_E_CUDA, "k_istft_frames smem attr: %s", cudaGetErrorString(e));
    smem_set = smem;
  }
  int grid = b->NP / FPB < 148 * 3 ? b->NP / FPB : 148 * 3;
  k_istft_frames<<<grid, FFT_THREADS, smem, st>>>(b->F, b->NP, n_fft, bits, b->frame_utt,
                                                                           reinterpret_cast<const float2*>(S), ws);
  int32_t rc = check_launch("k_istft_frames");
  if (rc) return rc;
  dim3 g2((T_stride + 255) / 256, b->B);
  k_istft_ola<<<g2, 256, 0, st>>>(n_fft, hop, T_stride, b->frame_off, b->n_frames, out_len, ws, out);
  return check_launch("k_istft_ola");
}

}  // namespace gvn
