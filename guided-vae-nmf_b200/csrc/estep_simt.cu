// E-step, CUDA-core fp32 variant (GVN_PREC_FP32): one Metropolis-Hastings chain per frame,
// the whole chain inside one persistent CTA that owns a tile of TN frames.
//
// Replaces MCEM_M{1,2}.sample_posterior + compute_Vs (reference python/models/mcem.py:218-307,
// :371-454).  Differences in *how*, not *what*:
//   * the decoder is evaluated once per proposal; the reference's second sweep after the
//     accept (mcem.py:283) and its re-decode of the kept samples (mcem.py:300) are replaced
//     by the values already computed, kept in slot form (Vs + multiplicities Vs_w, gvn.h);
//   * the current-state part of the log acceptance ratio (mcem.py:266-268) is cached per
//     frame as C_t = sum_f log Vx_t + X2/Vx_t and only refreshed on accept;
//   * the label columns of the first layer are hoisted into `yproj` (gvn_label_projection).
// This is also the on-device reference the tensor-core variant (estep_tc.cu) is tested against.
#include "gvn_common.cuh"

namespace gvn {

namespace {

constexpr int TN = 64;        // frames per CTA
constexpr int NT = 256;       // threads per CTA
constexpr int KT = 32;        // weight rows staged per k-tile
constexpr int HID = GVN_HIDDEN;

struct EstepArgs {
  int F, L, NP, burnin, R;
  float sd;                    // sqrt(var_RW) in fp32, as mcem.py:231,257
  const int32_t* frame_utt;
  const float* X2; const float* g; const float* Vb; const float* yproj;
  int xv_bf16;
  float* Z; float* Vs; float* Vs_w;
  // decoder (fp32 section of the packed image)
  const float* w1zT; const float* w2T; const float* b2; const float* w3T; const float* b3; int FS;
  // noise
  const float* eps; const float* u; const uint8_t* forced; uint64_t seed, chain;
  // trace
  float* t_acc; uint8_t* t_dec; int32_t* t_cnt; float* t_zs;
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// acc[i][j] = sum_k sA[k][4*tg+i] * gW[k*ldw + c0 + 8*fg + j]   (k < Kdim)
__device__ __forceinline__ void gemm_tile(float (&acc)[4][8], const float* sA, const float* __restrict__ gW,
                                          int ldw, int c0, int Kdim, float* sW, int tid, int tg, int fg,
                                          bool active) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  const int ntile = (Kdim + KT - 1) / KT;
  auto load = [&](int t, int buf) {
    for (int idx = tid; idx < KT * 32; idx += NT) {
      int kk = idx >> 5, q = idx & 31, k = t * KT + kk;
      if (k < Kdim) cp_async16(sW + (buf * KT + kk) * HID + 4 * q, gW + (size_t)k * ldw + c0 + 4 * q);
    }
    cp_async_commit();
  };
  load(0, 0);
  for (int t = 0; t < ntile; ++t) {
    if (t + 1 < ntile) { load(t + 1, (t + 1) & 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    __syncthreads();
    if (active) {
      const int kmax = min(KT, Kdim - t * KT);
      const float* a = sA + (size_t)(t * KT) * TN + 4 * tg;
      const float* w = sW + ((t & 1) * KT) * HID + 8 * fg;
#pragma unroll 8
      for (int kk = 0; kk < kmax; ++kk) {
        float4 av = *reinterpret_cast<const float4*>(a + kk * TN);
        float4 w0 = *reinterpret_cast<const float4*>(w + kk * HID);
        float4 w1 = *reinterpret_cast<const float4*>(w + kk * HID + 4);
        float aa[4] = {av.x, av.y, av.z, av.w};
        float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(aa[i], ww[j], acc[i][j]);
      }
    }
    __syncthreads();
  }
}

enum { MODE_INIT = 0, MODE_PROP = 1, MODE_WRITE = 2 };

__global__ void __launch_bounds__(NT, 2) k_estep_simt(EstepArgs p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int L = p.L, F = p.F, NP = p.NP;
  float* sZ = reinterpret_cast<float*>(smem_raw);      // [L][TN] current state
  float* sZp = sZ + L * TN;                             // [L][TN] proposal
  float* sA1 = sZp + L * TN;                            // [HID][TN]
  float* sA2 = sA1 + HID * TN;                          // [HID][TN]
  float* sW = sA2 + HID * TN;                           // [2][KT][HID]
  double* sPart = reinterpret_cast<double*>(sW + 2 * KT * HID);   // [8][TN]
  double* sCt = sPart + 8 * TN;                         // [TN] cached current-state energy
  double* sCp = sCt + TN;                               // [TN] proposal energy
  float* sG = reinterpret_cast<float*>(sCp + TN);       // [TN]
  int* sAcc = reinterpret_cast<int*>(sG + TN);          // [TN] decision of this step
  int* sCnt = sAcc + TN;                                // [TN] accepted count
  int* sValid = sCnt + TN;                              // [TN]
  int* sCur = sValid + TN;                              // [TN] slot holding the current state
  float* sMul = reinterpret_cast<float*>(sCur + TN);    // [TN] its multiplicity so far

  const int tid = threadIdx.x, tg = tid & 15, fg = tid >> 4;
  const int warp = tid >> 5, lane = tid & 31;
  const int n0 = blockIdx.x * TN;

  if (tid < TN) {
    int n = n0 + tid;
    bool v = n < NP && p.frame_utt[n] >= 0;
    sValid[tid] = v;
    sG[tid] = v ? p.g[n] : 1.f;
    sCnt[tid] = 0;
    sAcc[tid] = 0;
    sCur[tid] = 0;
    sMul[tid] = 0.f;
    if (v) for (int s = 0; s < p.R; ++s) p.Vs_w[(size_t)s * NP + n] = 0.f;
  }
  for (int idx = tid; idx < L * TN; idx += NT) {
    int l = idx / TN, c = idx % TN, n = n0 + c;
    sZ[idx] = (n < NP) ? p.Z[(size_t)l * NP + n] : 0.f;
  }
  __syncthreads();

  // frames of this thread in the register tile and their validity for global reads
  const int nb = n0 + 4 * tg;
  const bool in_range = nb < NP;      // NP is a multiple of 32 and nb a multiple of 4

  // ---- one decoder evaluation of the tile; MODE selects what the output layer does ----
  auto decode = [&](const float* zsrc, int mode, float* vs_out) {
    float acc[4][8];
    // layer 1: tanh(yproj + W1z z)
    gemm_tile(acc, zsrc, p.w1zT, HID, 0, L, sW, tid, tg, fg, true);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int jj = 8 * fg + j;
      float4 yp = in_range ? *reinterpret_cast<const float4*>(p.yproj + (size_t)jj * NP + nb)
                           : make_float4(0.f, 0.f, 0.f, 0.f);
      float4 o = make_float4(tanhf(acc[0][j] + yp.x), tanhf(acc[1][j] + yp.y), tanhf(acc[2][j] + yp.z),
                             tanhf(acc[3][j] + yp.w));
      *reinterpret_cast<float4*>(sA1 + jj * TN + 4 * tg) = o;
    }
    // layer 2: tanh(b2 + W2 a1)
    gemm_tile(acc, sA1, p.w2T, HID, 0, HID, sW, tid, tg, fg, true);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int jj = 8 * fg + j;
      float b = p.b2[jj];
      float4 o = make_float4(tanhf(acc[0][j] + b), tanhf(acc[1][j] + b), tanhf(acc[2][j] + b),
                             tanhf(acc[3][j] + b));
      *reinterpret_cast<float4*>(sA2 + jj * TN + 4 * tg) = o;
    }
    // layer 3 in chunks of 128 output features, fused epilogue
    double cs[4] = {0.0, 0.0, 0.0, 0.0};
    float gq[4] = {sG[4 * tg], sG[4 * tg + 1], sG[4 * tg + 2], sG[4 * tg + 3]};
    for (int c0 = 0; c0 < F; c0 += HID) {
      const int cw = min(HID, F - c0);
      const bool active = 8 * fg < cw;
      gemm_tile(acc, sA2, p.w3T, p.FS, c0, HID, sW, tid, tg, fg, active);
      if (active && in_range) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          int f = c0 + 8 * fg + j;
          if (f < F) {
            float b = p.b3[f];
            float vs[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) vs[i] = expf(acc[i][j] + b);
            size_t off = (size_t)f * NP + nb;
            if (mode != MODE_WRITE) {
              float4 vb = *reinterpret_cast<const float4*>(p.Vb + off);
              float4 x2 = *reinterpret_cast<const float4*>(p.X2 + off);
              float vbv[4] = {vb.x, vb.y, vb.z, vb.w}, x2v[4] = {x2.x, x2.y, x2.z, x2.w};
              if (p.xv_bf16) {                               // GVN_PREC_XV_BF16: the tensor-core chain's view of the two constants
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const uint32_t w = pack_xv_word(x2v[i], vbv[i]);
                  vbv[i] = __uint_as_float(w << 16);
                  x2v[i] = __uint_as_float(w);
                }
              }
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                float vx = fmaf(gq[i], vs[i], vbv[i]);
                cs[i] += (double)(logf(vx) + x2v[i] / vx);
              }
            }
            if (vs_out != nullptr)    // column-tile order; nb is a multiple of 4, so the four frames stay adjacent
              *reinterpret_cast<float4*>(vs_out + tile_off(f, nb, F)) =      // GVN_VS_MAX: a slot never holds inf (include/gvn.h)
                  make_float4(fminf(vs[0], GVN_VS_MAX), fminf(vs[1], GVN_VS_MAX), fminf(vs[2], GVN_VS_MAX), fminf(vs[3], GVN_VS_MAX));
          }
        }
      }
    }
    if (mode != MODE_WRITE) {
      // reduce over the 16 feature groups: lanes l and l^16 share tg, then across the 8 warps
#pragma unroll
      for (int i = 0; i < 4; ++i) cs[i] += __shfl_xor_sync(0xffffffffu, cs[i], 16);
      if (lane < 16) {
#pragma unroll
        for (int i = 0; i < 4; ++i) sPart[warp * TN + 4 * tg + i] = cs[i];
      }
      __syncthreads();
      if (tid < TN) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += sPart[w * TN + tid];
        (mode == MODE_INIT ? sCt : sCp)[tid] = s;
      }
      __syncthreads();
    }
  };

  decode(sZ, MODE_INIT, nullptr);

  const int n_steps = p.burnin + p.R;
  for (int m = 0; m < n_steps; ++m) {
    // ---- proposal  Z' = Z + sd * eps   (mcem.py:257) ----
    if (p.eps != nullptr) {
      for (int idx = tid; idx < L * TN; idx += NT) {
        int l = idx / TN, c = idx % TN, n = n0 + c;
        float e = (n < NP) ? p.eps[((size_t)m * L + l) * NP + n] : 0.f;
        sZp[idx] = sZ[idx] + p.sd * e;
      }
    } else {
      const int LQ = (L + 3) / 4;
      for (int idx = tid; idx < LQ * TN; idx += NT) {
        int lq = idx / TN, c = idx % TN, n = n0 + c;
        uint4 r = philox4x32_10(make_uint4((uint32_t)n, (uint32_t)m, (uint32_t)lq, (uint32_t)p.chain),
                                make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32)));
        float2 a = box_muller(r.x, r.y), b = box_muller(r.z, r.w);
        float e[4] = {a.x, a.y, b.x, b.y};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          int l = 4 * lq + i;
          if (l < L) sZp[l * TN + c] = sZ[l * TN + c] + p.sd * e[i];
        }
      }
    }
    __syncthreads();

    const int r = m - p.burnin;                         // kept-sample index (mcem.py:286)
    float* spec = (r >= 1) ? p.Vs + (size_t)r * F * NP : nullptr;
    decode(sZp, MODE_PROP, spec);

    // ---- accept / reject  (mcem.py:266-280) ----
    if (tid < TN) {
      int n = n0 + tid;
      float prior = 0.f;
      for (int l = 0; l < L; ++l) {
        float a = sZ[l * TN + tid], b = sZp[l * TN + tid];
        prior += a * a - b * b;
      }
      float acc_prob = (float)(sCt[tid] - sCp[tid]) + 0.5f * prior;
      int ok = 0;
      if (sValid[tid]) {
        float uu;
        if (p.u != nullptr) {
          uu = p.u[(size_t)m * NP + n];
        } else {
          uint4 rr = philox4x32_10(make_uint4((uint32_t)n, (uint32_t)m, 0xffffffffu, (uint32_t)p.chain),
                                   make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32)));
          uu = u01(rr.x);
        }
        ok = logf(uu) < acc_prob;
        if (p.forced != nullptr) ok = p.forced[(size_t)m * NP + n] != 0;
        if (p.t_acc != nullptr) p.t_acc[(size_t)m * NP + n] = acc_prob;
        if (p.t_dec != nullptr) p.t_dec[(size_t)m * NP + n] = (uint8_t)ok;
      }
      sAcc[tid] = ok;
      if (ok) { sCt[tid] = sCp[tid]; sCnt[tid] += 1; }
      // multiplicities of the kept samples: a rejected step repeats the state (mcem.py:280-289)
      if (r == 0) { sCur[tid] = 0; sMul[tid] = 1.f; }
      else if (r >= 1) {
        if (ok) { if (sValid[tid]) p.Vs_w[(size_t)sCur[tid] * NP + n] = sMul[tid]; sCur[tid] = r; sMul[tid] = 1.f; }
        else sMul[tid] += 1.f;
      }
    }
    __syncthreads();
    for (int idx = tid; idx < L * TN; idx += NT)
      if (sAcc[idx % TN]) sZ[idx] = sZp[idx];
    __syncthreads();

    // ---- emit the kept sample (mcem.py:286-289 + compute_Vs) ----
    if (r >= 0 && p.t_zs != nullptr) {
      for (int idx = tid; idx < L * TN; idx += NT) {
        int l = idx / TN, c = idx % TN, n = n0 + c;
        if (n < NP && sValid[c]) p.t_zs[((size_t)r * L + l) * NP + n] = sZ[idx];
      }
    }
    if (r == 0) decode(sZ, MODE_WRITE, p.Vs);
  }
  if (tid < TN && sValid[tid]) p.Vs_w[(size_t)sCur[tid] * NP + n0 + tid] = sMul[tid];

  for (int idx = tid; idx < L * TN; idx += NT) {
    int l = idx / TN, c = idx % TN, n = n0 + c;
    if (n < NP && sValid[c]) p.Z[(size_t)l * NP + n] = sZ[idx];
  }
  if (p.t_cnt != nullptr && tid < TN && sValid[tid]) p.t_cnt[n0 + tid] += sCnt[tid];
}

size_t estep_simt_smem(int L) {
  return (size_t)(2 * L * TN + 2 * HID * TN + 2 * KT * HID) * 4 + (size_t)(8 * TN + 2 * TN) * 8 +
         (size_t)TN * 4 * 6;
}

}  // namespace

int32_t launch_estep_simt(const gvn_batch* b, const void* packed, int burnin, int R,
                          float var_RW, const gvn_noise* nz, const gvn_trace* tr, int precision, cudaStream_t stream) {
  DecoderLayout d = decoder_layout(b->L, 0, b->F);
  const float* base = reinterpret_cast<const float*>(packed);
  EstepArgs a;
  a.F = b->F; a.L = b->L; a.NP = b->NP; a.burnin = burnin; a.R = R;
  a.sd = sqrtf(var_RW);
  a.frame_utt = b->frame_utt; a.X2 = b->X2; a.g = b->g; a.Vb = b->Vb; a.yproj = b->yproj;
  a.xv_bf16 = (precision & GVN_PREC_XV_BF16) != 0;
  a.Z = b->Z; a.Vs = b->Vs; a.Vs_w = b->Vs_w;
  a.w1zT = base + d.w1zT; a.w2T = base + d.w2T; a.b2 = base + d.b2; a.w3T = base + d.w3T;
  a.b3 = base + d.b3; a.FS = d.FS;
  a.eps = nz->eps; a.u = nz->u; a.forced = nz->forced_accept; a.seed = nz->seed; a.chain = nz->chain;
  a.t_acc = tr ? tr->acc_prob : nullptr; a.t_dec = tr ? tr->accepted : nullptr;
  a.t_cnt = tr ? tr->n_accepted : nullptr; a.t_zs = tr ? tr->z_samples : nullptr;
  size_t smem = estep_simt_smem(b->L);
  cudaError_t e = cudaFuncSetAttribute(k_estep_simt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return fail(GVN_E_CUDA, "estep_simt smem attr: %s", cudaGetErrorString(e));
  int grid = (b->NP + TN - 1) / TN;
  k_estep_simt<<<grid, NT, smem, stream>>>(a);
  return check_launch("k_estep_simt");
}

}  // namespace gvn
