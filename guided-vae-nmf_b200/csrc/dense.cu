// Small dense layers on feature-major activations and the decoder weight packing.
//
// gvn_dense replaces the per-utterance MLPs that bracket the MCEM loop in the reference:
// Encoder.forward mean head (python/models/models.py:90-104, called at mcem.py:214 / :367),
// Classifier.forward with the input standardisation of scripts/evaluate_M2_ibm.py:121-130
// (models.py:41-62), and the label half of the decoder's first layer (gvn_label_projection).
// They run once per utterance, so this is a plain shared-memory tiled SGEMM.
#include "gvn_common.cuh"

#include <cuda_fp16.h>

namespace gvn {

namespace {

constexpr int DT = 64;   // output tile (features x frames)
constexpr int DK = 16;   // reduction tile

__device__ __forceinline__ float activate(float v, int act) {
  switch (act) {
    case 1: return tanhf(v);
    case 2: return fmaxf(v, 0.f);
    case 3: return 1.0f / (1.0f + expf(-v));
    case 4: return (1.0f / (1.0f + expf(-v))) > 0.5f ? 1.f : 0.f;
    default: return v;
  }
}

// out[j][n] = act(b[j] + sum_i W[j][i] * in[i][n]),  in = [in0 (D0 rows); in1 (D1 rows)]
__global__ void __launch_bounds__(256) k_dense(const float* __restrict__ W, const float* __restrict__ bias,
                                               const float* __restrict__ in0, int D0, const float* __restrict__ in1,
                                               int D1, const float* __restrict__ mean, const float* __restrict__ std_,
                                               float eps, int D_out, int NP, int act, float* __restrict__ out) {
  __shared__ float sW[DK][DT + 1];
  __shared__ float sI[DK][DT];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int j0 = blockIdx.y * DT, n0 = blockIdx.x * DT;
  const int D = D0 + D1;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[a][c] = 0.f;
  for (int i0 = 0; i0 < D; i0 += DK) {
    for (int idx = tid; idx < DK * DT; idx += 256) {
      // weights: consecutive threads walk the reduction index (contiguous in memory)
      int ii = idx % DK, jj = idx / DK;
      int i = i0 + ii, j = j0 + jj;
      sW[ii][jj] = (i < D && j < D_out) ? W[(size_t)j * D + i] : 0.f;
      // activations: consecutive threads walk the frame index
      int kk = idx / DT, nn = idx % DT;
      int k = i0 + kk, n = n0 + nn;
      float v = 0.f;
      if (k < D && n < NP) {
        if (k < D0) {
          v = in0[(size_t)k * NP + n];
          if (mean != nullptr) v = (v - mean[k]) / (std_[k] + eps);
        } else {
          v = in1[(size_t)(k - D0) * NP + n];
        }
      }
      sI[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < DK; ++kk) {
      float w[4], x[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) w[a] = sW[kk][ty * 4 + a];
#pragma unroll
      for (int c = 0; c < 4; ++c) x[c] = sI[kk][tx * 4 + c];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[a][c] = fmaf(w[a], x[c], acc[a][c]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    int j = j0 + ty * 4 + a;
    if (j >= D_out) continue;
    float bj = bias ? bias[j] : 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      int n = n0 + tx * 4 + c;
      if (n < NP) out[(size_t)j * NP + n] = activate(acc[a][c] + bj, act);
    }
  }
}

// ---- decoder packing: transposed fp32 copies (CUDA-core path) -------------------------------
__global__ void k_pack_f32(const float* __restrict__ W1, const float* __restrict__ b1, const float* __restrict__ W2,
                           const float* __restrict__ b2, const float* __restrict__ W3, const float* __restrict__ b3,
                           int L, int y_dim, int F, DecoderLayout d, float* __restrict__ p) {
  const int H = GVN_HIDDEN, D1 = L + y_dim;
  const size_t stride = (size_t)gridDim.x * blockDim.x, t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (size_t i = t0; i < (size_t)L * H; i += stride) { int l = (int)(i / H), j = (int)(i % H); p[d.w1zT + i] = W1[(size_t)j * D1 + l]; }
  for (size_t i = t0; i < (size_t)H * y_dim; i += stride) { int j = (int)(i / y_dim), c = (int)(i % y_dim); p[d.w1y + i] = W1[(size_t)j * D1 + L + c]; }
  for (size_t i = t0; i < (size_t)H; i += stride) { p[d.b1 + i] = b1[i]; p[d.b2 + i] = b2[i]; }
  for (size_t i = t0; i < (size_t)H * H; i += stride) { int k = (int)(i / H), j = (int)(i % H); p[d.w2T + i] = W2[(size_t)j * H + k]; }
  for (size_t i = t0; i < (size_t)H * d.FS; i += stride) {
    int k = (int)(i / d.FS), f = (int)(i % d.FS);
    p[d.w3T + i] = f < F ? W3[(size_t)f * H + k] : 0.f;
  }
  for (size_t i = t0; i < round_up((size_t)F, 4); i += stride) p[d.b3 + i] = i < (size_t)F ? b3[i] : 0.f;
}

}  // namespace

int32_t launch_dense(const float* W, const float* b, const float* in0, int D0, const float* in1, int D1,
                     const float* mean, const float* std_, float eps, int D_out, int NP, int act, float* out,
                     cudaStream_t st) {
  dim3 grid((NP + DT - 1) / DT, (D_out + DT - 1) / DT);
  k_dense<<<grid, 256, 0, st>>>(W, b, in0, D0, in1, D1, mean, std_, eps, D_out, NP, act, out);
  return check_launch("k_dense");
}

int32_t launch_pack_tc(const float* W1, const float* W2, const float* W3, const float* b3, int L, int y_dim, int F,
                       unsigned char* image, cudaStream_t st);   // estep_tc.cu

int32_t launch_pack_decoder(const float* W1, const float* b1, const float* W2, const float* b2, const float* W3,
                            const float* b3, int L, int y_dim, int F, void* packed, cudaStream_t st) {
  DecoderLayout d = decoder_layout(L, y_dim, F);
  k_pack_f32<<<148, 256, 0, st>>>(W1, b1, W2, b2, W3, b3, L, y_dim, F, d, reinterpret_cast<float*>(packed));
  int32_t rc = check_launch("k_pack_f32");
  if (rc) return rc;
  return launch_pack_tc(W1, W2, W3, b3, L, y_dim, F, reinterpret_cast<unsigned char*>(packed) + d.tc_image, st);
}

}  // namespace gvn
