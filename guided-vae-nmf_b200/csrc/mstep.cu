// NMF / gain M-step, cost and Wiener filter: the memory-bound half of the MCEM loop.
//
// Replaces EM.M_step (reference python/models/mcem.py:90-152), compute_expected_neg_log_like
// (:68-70) and the Wiener tail of compute_WF / run (:341-343, :175-176).
//
// Dependencies inside one M-step: the W update reduces over frames,
// the H update reduces over frequency and needs the complete new W, the g update needs the
// new H, the cost needs the new g.  Variant 0 below is the straightforward schedule:
//   k_mstep_w     one warp per (utterance, frequency row): streams Vs[r][f][:] once
//   k_colnorm     per utterance: column sums of |W|, normalised copy of W
//   k_mstep_cols  one CTA per 32-frame tile: three passes (H, g, cost) over its Vs columns
// Algorithmic traffic is 2*(R+1)*F*N*4 bytes per utterance and iteration (Vs and X2 each read
// for the W sweep and for the column sweep); variant 0 re-reads the column block from L2/HBM
// for the g and cost passes.
#include "gvn_common.cuh"

namespace gvn {

namespace {

constexpr int MS_TILE = 32;     // frames per column tile == GVN_FRAME_ALIGN
constexpr int MS_WARPS = 8;

// ------------------------------------------------------------------ W update (mcem.py:105-110)
template <int KMAX>
__global__ void __launch_bounds__(256) k_mstep_w(int F, int K, int NP, int R, const int32_t* __restrict__ frame_off,
                                                 const int32_t* __restrict__ n_frames, const float* __restrict__ X2,
                                                 const float* __restrict__ Vs, const float* __restrict__ Vs_w,
                                                 const float* __restrict__ Vb, const float* __restrict__ g,
                                                 const float* __restrict__ H, const float* __restrict__ W,
                                                 float* __restrict__ Wun) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y, f = blockIdx.x * MS_WARPS + warp;
  if (f >= F) return;
  const int n_begin = frame_off[b], N = n_frames[b];
  float num[KMAX], den[KMAX];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) num[k] = den[k] = 0.f;
  const size_t row = (size_t)f * NP;
  const size_t slab = (size_t)F * NP;
  for (int n = lane; n < N; n += 32) {
    const int gn = n_begin + n;
    const float gg = g[gn], vb = Vb[row + gn], x2 = X2[row + gn];
    float s1 = 0.f, s2 = 0.f;
    const float* vsp = Vs + tile_off(f, gn, F);
#pragma unroll 5
    for (int r = 0; r < R; ++r) {
      float vx = fmaf(gg, vsp[(size_t)r * slab], vb);
      float inv = 1.0f / vx, wi = Vs_w[(size_t)r * NP + gn] * inv;     // slot multiplicity (gvn.h)
      s1 += wi;
      s2 = fmaf(wi, inv, s2);
    }
    const float a = x2 * s2;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      if (k < K) {
        float h = H[(size_t)k * NP + gn];
        num[k] = fmaf(a, h, num[k]);
        den[k] = fmaf(s1, h, den[k]);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    if (k < K) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        num[k] += __shfl_xor_sync(0xffffffffu, num[k], o);
        den[k] += __shfl_xor_sync(0xffffffffu, den[k], o);
      }
    }
  }
  // lane k writes column k (static indexing keeps num/den in registers)
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    if (k < K && lane == k) {
      size_t o = ((size_t)b * F + f) * K + k;
      Wun[o] = W[o] * sqrtf(num[k] / den[k]);
    }
  }
}

// ---------------------------------------------------- column normalisation (mcem.py:128-131)
__global__ void __launch_bounds__(256) k_colnorm(int F, int K, const float* __restrict__ Wun, float* __restrict__ W,
                                                 float* __restrict__ cnorm) {
  __shared__ float red[256];
  __shared__ float c_s[GVN_MAX_K];
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* w = Wun + (size_t)b * F * K;
  for (int k = 0; k < K; ++k) {
    float s = 0.f;
    for (int f = tid; f < F; f += 256) s += fabsf(w[(size_t)f * K + k]);
    red[tid] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (tid < o) red[tid] += red[tid + o];
      __syncthreads();
    }
    if (tid == 0) { c_s[k] = red[0]; cnorm[b * K + k] = red[0]; }
    __syncthreads();
  }
  for (int i = tid; i < F * K; i += 256) W[(size_t)b * F * K + i] = w[i] / c_s[i % K];
}

// -------------------------------- H, g, cost for one 32-frame tile (mcem.py:113-152, :68-70)
template <int KMAX>
__global__ void __launch_bounds__(256) k_mstep_cols(int F, int K, int NP, int R, const int32_t* __restrict__ frame_utt,
                                                    const float* __restrict__ X2, const float* __restrict__ Vs,
                                                    const float* __restrict__ Vs_w, float* __restrict__ Vb, float* __restrict__ g, float* __restrict__ H,
                                                    const float* __restrict__ Wun, const float* __restrict__ cnorm,
                                                    float* __restrict__ cost_part, uint32_t* __restrict__ XV) {
  __shared__ float red[MS_WARPS][KMAX][MS_TILE];
  __shared__ float red1[MS_WARPS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x, gn = tile * MS_TILE + lane;
  const int b = frame_utt[tile * MS_TILE];
  constexpr int CPT = MS_TILE / GVN_COST_TILE;                      // cost_part entries per 32-frame tile
  if (threadIdx.x < CPT) cost_part[tile * CPT + threadIdx.x] = 0.f;
  if (b < 0) return;
  const bool valid = frame_utt[gn] >= 0;
  const float* wb = Wun + (size_t)b * F * K;
  const size_t slab = (size_t)F * NP;
  const float gg = g[gn];

  float hk[KMAX], num[KMAX], den[KMAX];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) { hk[k] = (k < K) ? H[(size_t)k * NP + gn] : 0.f; num[k] = den[k] = 0.f; }

  // pass 1: H update with Vb = Wun @ H_old   (mcem.py:113-121)
  for (int f = warp; f < F; f += MS_WARPS) {
    float wr[KMAX];
    float vb = 0.f;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) { wr[k] = (k < K) ? __ldg(wb + (size_t)f * K + k) : 0.f; vb = fmaf(wr[k], hk[k], vb); }
    const size_t o = (size_t)f * NP + gn;
    const float x2 = X2[o];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll 5
    for (int r = 0; r < R; ++r) {
      float inv = 1.0f / fmaf(gg, Vs[tile_off(f, gn, F) + (size_t)r * slab], vb), wi = Vs_w[(size_t)r * NP + gn] * inv;
      s1 += wi;
      s2 = fmaf(wi, inv, s2);
    }
    const float a = x2 * s2;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) { num[k] = fmaf(wr[k], a, num[k]); den[k] = fmaf(wr[k], s1, den[k]); }
  }
  // cross-warp sums in two rounds (numerators, then denominators) to keep smem at 32 KB
#pragma unroll
  for (int k = 0; k < KMAX; ++k) red[warp][k][lane] = num[k];
  __syncthreads();
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < MS_WARPS; ++w) s += red[w][k][lane];
    num[k] = s;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < KMAX; ++k) red[warp][k][lane] = den[k];
  __syncthreads();
  float hn[KMAX];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < MS_WARPS; ++w) s += red[w][k][lane];
    hn[k] = (k < K) ? hk[k] * sqrtf(num[k] / s) : 0.f;
  }
  __syncthreads();

  // pass 2: Vb = Wun @ H_new (kept for the next E-step, mcem.py:124), g update (mcem.py:138-142)
  float ng = 0.f, dg = 0.f;
  for (int f = warp; f < F; f += MS_WARPS) {
    float vb = 0.f;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) { if (k < K) vb = fmaf(__ldg(wb + (size_t)f * K + k), hn[k], vb); }
    const size_t o = (size_t)f * NP + gn;
    const float x2 = X2[o];
    if (valid) { Vb[o] = vb; if (XV != nullptr) XV[o] = pack_xv_word(x2, vb); }
    float t2 = 0.f, t1 = 0.f;
#pragma unroll 5
    for (int r = 0; r < R; ++r) {
      float vs = Vs[tile_off(f, gn, F) + (size_t)r * slab];
      float inv = 1.0f / fmaf(gg, vs, vb), wv = Vs_w[(size_t)r * NP + gn] * vs;
      t1 = fmaf(wv, inv, t1);
      t2 = fmaf(wv * inv, inv, t2);
    }
    ng = fmaf(x2, t2, ng);
    dg += t1;
  }
  red[warp][0][lane] = ng;
  red[warp][1][lane] = dg;
  __syncthreads();
  float sn = 0.f, sd = 0.f;
#pragma unroll
  for (int w = 0; w < MS_WARPS; ++w) { sn += red[w][0][lane]; sd += red[w][1][lane]; }
  const float gnew = gg * sqrtf(sn / sd);
  __syncthreads();

  // pass 3: cost with the new g (mcem.py:151-152, :68-70)
  float cs = 0.f;
  if (valid) {
    for (int f = warp; f < F; f += MS_WARPS) {
      const size_t o = (size_t)f * NP + gn;
      const float vb = Vb[o], x2 = X2[o];
      float c = 0.f;
#pragma unroll 5
      for (int r = 0; r < R; ++r) {
        float vx = fmaf(gnew, Vs[tile_off(f, gn, F) + (size_t)r * slab], vb);
        c = fmaf(Vs_w[(size_t)r * NP + gn], logf(vx) + x2 / vx, c);
      }
      cs += c;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cs += __shfl_xor_sync(0xffffffffu, cs, o);
  if (lane == 0) red1[warp] = cs;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < MS_WARPS; ++w) s += red1[w];
    cost_part[tile * CPT] = s;
  }
  // normalised H (mcem.py:133) and new g
  if (warp == 0 && valid) {
#pragma unroll
    for (int k = 0; k < KMAX; ++k) { if (k < K) H[(size_t)k * NP + gn] = hn[k] * __ldg(cnorm + b * K + k); }
    g[gn] = gnew;
  }
}

__global__ void k_cost_reduce(int B, int F, int R, int niter, int ntiles, const int32_t* __restrict__ frame_off,
                              const int32_t* __restrict__ n_frames, const float* __restrict__ part,
                              double* __restrict__ cost) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= niter * B) return;
  int it = i / B, b = i % B;
  int t0 = frame_off[b] / GVN_COST_TILE, t1 = (frame_off[b] + n_frames[b] + MS_TILE - 1) / MS_TILE * (MS_TILE / GVN_COST_TILE);
  double s = 0.0;
  for (int t = t0; t < t1; ++t) s += (double)part[(size_t)it * ntiles + t];
  cost[i] = s / ((double)R * F * n_frames[b]);
}

// ------------------------------------------------------------ Wiener (mcem.py:341-343, :175-176)
// Threads walk the points in the column-tile order of Vs (tile, bin, frame-in-tile): the R slot reads of a warp -- the
// bulk of the traffic, R x F x NP floats -- are 128 contiguous bytes each; Vb, X and the outputs ([F][NP] order) are
// touched in whole 32- / 64-byte row segments.
__global__ void __launch_bounds__(256) k_wiener(int F, int NP, int R, const int32_t* __restrict__ frame_utt,
                                                const float* __restrict__ Vs, const float* __restrict__ Vs_w,
                                                const float* __restrict__ Vb, const float* __restrict__ g,
                                                const float2* __restrict__ Xc,
                                                float2* __restrict__ S_hat, float2* __restrict__ N_hat,
                                                float* __restrict__ WFs, float* __restrict__ WFn) {
  const size_t total = (size_t)F * NP, slab = total;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int j = (int)(t % GVN_VS_TILE);
    const size_t q = t / GVN_VS_TILE;
    const int f = (int)(q % F), gn = (int)(q / F) * GVN_VS_TILE + j;
    const size_t i = (size_t)f * NP + gn;
    float ws = 0.f, wn = 0.f;
    if (frame_utt[gn] >= 0) {
      const float gg = g[gn], vb = Vb[i];
      // the slot reads are the traffic of this kernel (R x F x NP floats, read once): WU of them are requested before
      // the first is used -- one load in flight per thread keeps ~8 KB per SM on the wire, a fifth of what the HBM latency needs
      constexpr int WU = 5;
      for (int r0 = 0; r0 < R; r0 += WU) {
        float v[WU], w[WU];
#pragma unroll
        for (int u = 0; u < WU; ++u) {
          const int r = r0 + u;
          w[u] = r < R ? Vs_w[(size_t)r * NP + gn] : 0.f;
          v[u] = r < R ? __ldcs(Vs + t + (size_t)r * slab) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < WU; ++u) {
          if (w[u] == 0.f) continue;                            // dead slot (rejected proposal)
          const float sc = gg * v[u];
          const float wi = w[u] * __frcp_rn(sc + vb);
          ws = fmaf(wi, sc, ws);
          wn = fmaf(wi, vb, wn);
        }
      }
      ws /= (float)R;
      wn /= (float)R;
    }
    const float2 x = Xc[i];
    S_hat[i] = make_float2(ws * x.x, ws * x.y);
    N_hat[i] = make_float2(wn * x.x, wn * x.y);
    if (WFs != nullptr) WFs[i] = ws;
    if (WFn != nullptr) WFn[i] = wn;
  }
}

// ------------------------------------------------------------------ init (mcem.py:36-57)
__global__ void k_init_nmf_w(size_t n, const float* __restrict__ rnd, float eps, float* __restrict__ W) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) W[i] = fmaxf(rnd[i], eps);
}
// one CTA per 32-frame group: H, g and the padding constants by the first warp, then Vb = W @ H and the
// column-tile copy of X2 with threads over (frequency, frame) -- coalesced along the frame axis
__global__ void __launch_bounds__(256) k_init_nmf_cols(int F, int K, int NP, const int32_t* __restrict__ frame_utt,
                                                       const float* __restrict__ rndH, float eps, const float* __restrict__ W,
                                                       float* __restrict__ H, float* __restrict__ g, float* __restrict__ Vb,
                                                       float* __restrict__ X2, float2* __restrict__ Xc, float* __restrict__ X2t) {
  __shared__ float h_s[GVN_MAX_K][32];
  const int lane = threadIdx.x & 31, fw = threadIdx.x >> 5;
  const int gn = blockIdx.x * 32 + lane;
  const int b = frame_utt[gn];
  if (fw == 0) {
    g[gn] = 1.f;
    for (int k = 0; k < K; ++k) {
      const float h = b < 0 ? 1.f : fmaxf(rndH[(size_t)k * NP + gn], eps);
      H[(size_t)k * NP + gn] = h;
      h_s[k][lane] = h;
    }
  }
  __syncthreads();
  const float* w = W + (size_t)(b < 0 ? 0 : b) * F * K;
  for (int f = fw; f < F; f += 8) {
    const size_t o = (size_t)f * NP + gn;
    if (b < 0) {   // padding frame: benign constants, never updated
      Vb[o] = 1.f;
      X2[o] = 1.f;
      if (X2t != nullptr) X2t[tile_off(f, gn, F)] = 1.f;
      if (Xc != nullptr) Xc[o] = make_float2(0.f, 0.f);
    } else {
      float vb = 0.f;
      for (int k = 0; k < K; ++k) vb = fmaf(__ldg(w + (size_t)f * K + k), h_s[k][lane], vb);
      Vb[o] = vb;
      if (X2t != nullptr) X2t[tile_off(f, gn, F)] = X2[o];   // column-tile copy for the M-step
    }
  }
}

}  // namespace

size_t mstep_v1_workspace_bytes(const gvn_batch*);
size_t mstep_workspace_bytes(const gvn_batch* b) { return round_up((size_t)b->B * b->K, 64) * sizeof(float) + mstep_v1_workspace_bytes(b) + 256; }

bool mstep_v1_supported(const gvn_batch*, int);
size_t mstep_v1_workspace_bytes(const gvn_batch*);
int32_t launch_mstep_v1(const gvn_batch*, int, float*, float*, cudaStream_t);

int32_t launch_mstep(const gvn_batch* b, int R, float* cost_part, void* workspace, int variant, cudaStream_t st) {
  // variant 0: reference schedule (any shape); variant 1 (default when the tile fits in shared memory): mstep_v1.cu
  if (variant != 0 && mstep_v1_supported(b, R))
    return launch_mstep_v1(b, R, cost_part, reinterpret_cast<float*>(workspace) + round_up((size_t)b->B * b->K, 64), st);   // 256-byte aligned
  float* cnorm = reinterpret_cast<float*>(workspace);
  dim3 gw((b->F + MS_WARPS - 1) / MS_WARPS, b->B);
  const int ntiles = b->NP / MS_TILE;
  if (b->K <= 16) {
    k_mstep_w<16><<<gw, 256, 0, st>>>(b->F, b->K, b->NP, R, b->frame_off, b->n_frames, b->X2, b->Vs, b->Vs_w, b->Vb, b->g,
                                       b->H, b->W, b->Wun);
  } else {
    k_mstep_w<32><<<gw, 256, 0, st>>>(b->F, b->K, b->NP, R, b->frame_off, b->n_frames, b->X2, b->Vs, b->Vs_w, b->Vb, b->g,
                                       b->H, b->W, b->Wun);
  }
  int32_t rc = check_launch("k_mstep_w");
  if (rc) return rc;
  k_colnorm<<<b->B, 256, 0, st>>>(b->F, b->K, b->Wun, b->W, cnorm);
  if ((rc = check_launch("k_colnorm"))) return rc;
  if (b->K <= 16) {
    k_mstep_cols<16><<<ntiles, 256, 0, st>>>(b->F, b->K, b->NP, R, b->frame_utt, b->X2, b->Vs, b->Vs_w, b->Vb, b->g, b->H,
                                             b->Wun, cnorm, cost_part, b->XV);
  } else {
    k_mstep_cols<32><<<ntiles, 256, 0, st>>>(b->F, b->K, b->NP, R, b->frame_utt, b->X2, b->Vs, b->Vs_w, b->Vb, b->g, b->H,
                                             b->Wun, cnorm, cost_part, b->XV);
  }
  return check_launch("k_mstep_cols");
}

int32_t launch_cost_reduce(const gvn_batch* b, int R, int niter, const float* part, double* cost, cudaStream_t st) {
  int n = niter * b->B;
  k_cost_reduce<<<(n + 127) / 128, 128, 0, st>>>(b->B, b->F, R, niter, b->NP / GVN_COST_TILE, b->frame_off, b->n_frames,
                                                 part, cost);
  return check_launch("k_cost_reduce");
}

int32_t launch_wiener(const gvn_batch* b, int R, float* S_hat, float* N_hat, float* WFs, float* WFn, cudaStream_t st) {
  size_t total = (size_t)b->F * b->NP;
  int grid = (int)((total + 255) / 256);
  if (grid > 148 * 16) grid = 148 * 16;
  k_wiener<<<grid, 256, 0, st>>>(b->F, b->NP, R, b->frame_utt, b->Vs, b->Vs_w, b->Vb, b->g,
                                 reinterpret_cast<const float2*>(b->Xc), reinterpret_cast<float2*>(S_hat),
                                 reinterpret_cast<float2*>(N_hat), WFs, WFn);
  return check_launch("k_wiener");
}

int32_t launch_init_nmf(const gvn_batch* b, const float* rand_W, const float* rand_H, float eps, cudaStream_t st) {
  size_t n = (size_t)b->B * b->F * b->K;
  k_init_nmf_w<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, rand_W, eps, b->W);
  int32_t rc = check_launch("k_init_nmf_w");
  if (rc) return rc;
  k_init_nmf_cols<<<b->NP / 32, 256, 0, st>>>(b->F, b->K, b->NP, b->frame_utt, rand_H, eps, b->W, b->H, b->g,
                                                       b->Vb, b->X2, reinterpret_cast<float2*>(b->Xc), b->X2t);
  return check_launch("k_init_nmf_cols");
}

}  // namespace gvn
