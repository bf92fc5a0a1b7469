// Gain-only M-step for the models without an NMF noise dictionary: replaces EM_noNMF.M_step and
// compute_expected_neg_log_like (reference python/models/mcem.py:551-588, :530-532) of
// MCEM_M2_noNMF (:609-760).  The noise variance Vb is an input that never changes; per frame
//     g <- g * sqrt( sum_f X2 sum_r Vs/Vx^2  /  sum_f sum_r Vs/Vx ),   Vx = g*Vs + Vb   (old g)
// and the cost of the refreshed Vx.  Both reduce over frequency only, so one CTA owns an 8-frame column
// tile (the contiguous F x 8 block of the column-tile layout, include/gvn.h): the g pass streams the R
// sample slots from HBM -- a warp covers 4 bins x 8 frames = 128 contiguous bytes per slot -- and the
// cost pass finds them in L2.  Algorithmic traffic: (R+2)*F*N*4 bytes per utterance and iteration.
#include "gvn_common.cuh"

namespace gvn {

namespace {

constexpr int NB = GVN_COST_TILE;
constexpr int GT = 256;               // GFL frequency lanes x NB frames
constexpr int GFL = GT / NB;

__device__ __forceinline__ float rcp_fast(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_fast(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

__global__ void __launch_bounds__(GT) k_gain_cols(int F, int NP, int R, int ntiles, const int32_t* __restrict__ frame_utt,
                                                  const float* __restrict__ X2t, const float* __restrict__ Vs,
                                                  const float* __restrict__ Vs_w, const float* __restrict__ Vb,
                                                  float* __restrict__ g, float* __restrict__ cost_part) {
  __shared__ float red[8][2][NB];
  __shared__ float wts[GVN_MAX_R_SLOTS][NB];
  __shared__ float misc[8];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = tid & (NB - 1), fl = tid / NB;
  const size_t slab = (size_t)(NP / NB) * F * NB;           // one sample slot
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int gn = t * NB + n;
    if (frame_utt[(size_t)t * NB] < 0) { if (tid == 0) cost_part[t] = 0.f; continue; }   // tiles never straddle utterances
    __syncthreads();                                        // previous tile is done with the shared arrays
    for (int i = tid; i < R * NB; i += GT) wts[i / NB][i % NB] = Vs_w[(size_t)(i / NB) * NP + t * NB + (i % NB)];
    __syncthreads();
    const bool valid = frame_utt[gn] >= 0;
    const float gg = g[gn];
    const float* vs0 = Vs + (size_t)t * F * NB + n;         // + f*NB + r*slab
    const float* x2p = X2t + (size_t)t * F * NB + n;

    // ---------------- gain update (mcem.py:566-576)
    float ng = 0.f, dg = 0.f;
    for (int f = fl; f < F; f += GFL) {
      const float vb = Vb[(size_t)f * NP + gn];
      const float* vs = vs0 + (size_t)f * NB;
      float t1 = 0.f, t2 = 0.f;
      int r = 0;
      for (; r + 1 < R; r += 2) {                           // 1/a and 1/c from one reciprocal of the product
        const float va = __ldg(vs + r * slab), vc = __ldg(vs + (r + 1) * slab);
        const float a = fmaf(gg, va, vb), c = fmaf(gg, vc, vb);
        const float ip = rcp_fast(a * c);
        const float ia = c * ip, ic = a * ip;
        const float ua = wts[r][n] * va * ia, uc = wts[r + 1][n] * vc * ic;
        t1 += ua + uc;
        t2 = fmaf(ua, ia, fmaf(uc, ic, t2));
      }
      if (r < R) {
        const float va = __ldg(vs + r * slab);
        const float ia = rcp_fast(fmaf(gg, va, vb)), ua = wts[r][n] * va * ia;
        t1 += ua;
        t2 = fmaf(ua, ia, t2);
      }
      ng = fmaf(__ldg(x2p + (size_t)f * NB), t2, ng);
      dg += t1;
    }
#pragma unroll
    for (int o = NB; o < 32; o <<= 1) {                     // lanes that share the frame n
      ng += __shfl_xor_sync(0xffffffffu, ng, o);
      dg += __shfl_xor_sync(0xffffffffu, dg, o);
    }
    if (lane < NB) { red[warp][0][lane] = ng; red[warp][1][lane] = dg; }
    __syncthreads();
    float sn = 0.f, sd = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) { sn += red[w8][0][n]; sd += red[w8][1][n]; }
    const float gnew = gg * sqrtf(sn / sd);
    if (tid < NB && valid) g[gn] = gnew;

    // ---------------- cost with the new gain (mcem.py:585-588, :530-532)
    float cl = 0.f, cr = 0.f;
    for (int f = fl; f < F; f += GFL) {
      const float vb = Vb[(size_t)f * NP + gn];
      const float* vs = vs0 + (size_t)f * NB;
      float sl = 0.f, sr = 0.f;
      int r = 0;
      for (; r + 1 < R; r += 2) {
        const float a = fmaf(gnew, __ldg(vs + r * slab), vb), c = fmaf(gnew, __ldg(vs + (r + 1) * slab), vb);
        const float wa = wts[r][n], wc = wts[r + 1][n];
        sl = fmaf(wa, lg2_fast(a), fmaf(wc, lg2_fast(c), sl));
        sr = fmaf(fmaf(wa, c, wc * a), rcp_fast(a * c), sr);
      }
      if (r < R) {
        const float a = fmaf(gnew, __ldg(vs + r * slab), vb), wa = wts[r][n];
        sl = fmaf(wa, lg2_fast(a), sl);
        sr = fmaf(wa, rcp_fast(a), sr);
      }
      cl += sl;
      cr = fmaf(__ldg(x2p + (size_t)f * NB), sr, cr);
    }
    float cs = valid ? fmaf(cl, 0.6931471805599453f, cr) : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cs += __shfl_xor_sync(0xffffffffu, cs, o);
    if (lane == 0) misc[warp] = cs;
    __syncthreads();
    if (tid == 0) {
      float sum = 0.f;
      for (int w8 = 0; w8 < 8; ++w8) sum += misc[w8];
      cost_part[t] = sum;
    }
  }
}

}  // namespace

int32_t launch_mstep_gain(const gvn_batch* b, int R, float* cost_part, cudaStream_t st) {
  GVN_REQUIRE(R <= GVN_MAX_R_SLOTS, GVN_E_UNSUPPORTED_SHAPE, "gain M-step: R=%d sample slots, at most %d", R, GVN_MAX_R_SLOTS);
  const int ntiles = b->NP / NB;
  const int grid = ntiles < 148 * 8 ? ntiles : 148 * 8;
  k_gain_cols<<<grid, GT, 0, st>>>(b->F, b->NP, R, ntiles, b->frame_utt, b->X2t, b->Vs, b->Vs_w, b->Vb, b->g, cost_part);
  return check_launch("k_gain_cols");
}

}  // namespace gvn
