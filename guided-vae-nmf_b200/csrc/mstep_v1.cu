// NMF / gain M-step, variant 1: the HBM-bound schedule.
//
// Same arithmetic as variant 0 in mstep.cu (reference python/models/mcem.py:90-152 and the cost
// of :68-70); what changes is how the (R,F,N) block of speech variances moves:
//
//   k_w_v1     W update (mcem.py:105-110).  One warp per (utterance, frequency row), float4
//              loads of all R sample slots + X2 + Vb per lane in flight, H through L1, one pass.
//   k_cols_v1  H update, refresh, normalisation, g update and cost (mcem.py:113-152, :68-70)
//              as ONE sweep: a persistent CTA owns a contiguous range of 8-frame column tiles;
//              the whole F x (R+1) x 8 block of a tile (Vs slots + X2) is staged ONCE in shared
//              memory by 16-byte async copies and the three dependent passes (H | g | cost) run
//              out of shared memory.  The copy of the next tile is issued chunk by chunk behind
//              the cost pass of the current one and the H pass of the next tile chases the
//              arriving chunks, so loads are in flight during two of the three passes.
//
// The samples arrive in slot form (Vs, Vs_w: include/gvn.h): every sum over the R samples of the
// reference is a multiplicity-weighted sum over the R slots.
// Transcendental budget: 1/a and 1/c of a pair of slots come from ONE reciprocal of the product
// (1/a = c * rcp(ac)); Vx >= K*eps^2 ~ 1e-15, so the product stays far inside the fp32 range.
#include "gvn_common.cuh"

namespace gvn {

namespace {

constexpr int NB = GVN_COST_TILE;     // frames per column tile
constexpr int CT = 256;               // threads per CTA
constexpr int MAXCH = 8;              // load chunks per tile

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_wait_dyn(int pending) {
  switch (pending) {
    case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
    case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
    case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
    case 3: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
    case 4: asm volatile("cp.async.wait_group 4;" ::: "memory"); break;
    case 5: asm volatile("cp.async.wait_group 5;" ::: "memory"); break;
    case 6: asm volatile("cp.async.wait_group 6;" ::: "memory"); break;
    default: asm volatile("cp.async.wait_group 7;" ::: "memory"); break;
  }
}
__device__ __forceinline__ float rcp_fast(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_fast(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float4 ldg_stream4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// weighted sums  s1 += wa/a + wc/c,  s2 += wa/a^2 + wc/c^2  with a = g*va + vb, c = g*vc + vb
__device__ __forceinline__ void pair_acc(float g, float vb, float va, float vc, float wa, float wc, float& s1, float& s2) {
  const float a = fmaf(g, va, vb), c = fmaf(g, vc, vb);
  const float ip = rcp_fast(a * c);
  const float ia = c * ip, ic = a * ip;          // 1/a, 1/c
  const float ta = wa * ia, tc = wc * ic;
  s1 += ta;
  s1 += tc;
  s2 = fmaf(ta, ia, s2);
  s2 = fmaf(tc, ic, s2);
}
__device__ __forceinline__ void single_acc(float g, float vb, float va, float wa, float& s1, float& s2) {
  const float ia = rcp_fast(fmaf(g, va, vb)), ta = wa * ia;
  s1 += ta;
  s2 = fmaf(ta, ia, s2);
}

// ------------------------------------------------------------------ W update (mcem.py:105-110)
template <int KMAX, int RT>
__global__ void __launch_bounds__(CT) k_w_v1(int F, int K, int NP, int R_rt, const int32_t* __restrict__ frame_off,
                                             const int32_t* __restrict__ n_frames, const float* __restrict__ X2,
                                             const float* __restrict__ Vs, const float* __restrict__ Vs_w,
                                             const float* __restrict__ Vb, const float* __restrict__ g,
                                             const float* __restrict__ H, const float* __restrict__ W,
                                             float* __restrict__ Wun) {
  const int R = RT > 0 ? RT : R_rt;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y, f = blockIdx.x * (CT / 32) + warp;
  if (f >= F) return;
  const int n_begin = frame_off[b], N = n_frames[b];
  const int NPAD = (N + GVN_FRAME_ALIGN - 1) / GVN_FRAME_ALIGN * GVN_FRAME_ALIGN;
  float num[KMAX], den[KMAX];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) num[k] = den[k] = 0.f;
  const size_t row = (size_t)f * NP, slab = (size_t)F * NP;
  for (int n4 = lane * 4; n4 < NPAD; n4 += 128) {
    const size_t o = row + n_begin + n4;
    const float* vsp = Vs + o;
    const float* wp = Vs_w + n_begin + n4;
    const float4 vb = ldg_stream4(Vb + o), x2 = ldg_stream4(X2 + o);
    const float4 gg = __ldg(reinterpret_cast<const float4*>(g + n_begin + n4));
    float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
    if (RT > 0) {
      float4 v[RT > 0 ? RT : 1];
#pragma unroll
      for (int r = 0; r < RT; ++r) v[r] = ldg_stream4(vsp + (size_t)r * slab);     // all slots in flight
#pragma unroll
      for (int r = 0; r + 1 < RT; r += 2) {
        const float4 wa = __ldg(reinterpret_cast<const float4*>(wp + (size_t)r * NP));
        const float4 wc = __ldg(reinterpret_cast<const float4*>(wp + (size_t)(r + 1) * NP));
        pair_acc(gg.x, vb.x, v[r].x, v[r + 1].x, wa.x, wc.x, s1.x, s2.x);
        pair_acc(gg.y, vb.y, v[r].y, v[r + 1].y, wa.y, wc.y, s1.y, s2.y);
        pair_acc(gg.z, vb.z, v[r].z, v[r + 1].z, wa.z, wc.z, s1.z, s2.z);
        pair_acc(gg.w, vb.w, v[r].w, v[r + 1].w, wa.w, wc.w, s1.w, s2.w);
      }
      if (RT & 1) {
        const float4 wa = __ldg(reinterpret_cast<const float4*>(wp + (size_t)(RT - 1) * NP));
        single_acc(gg.x, vb.x, v[RT - 1].x, wa.x, s1.x, s2.x);
        single_acc(gg.y, vb.y, v[RT - 1].y, wa.y, s1.y, s2.y);
        single_acc(gg.z, vb.z, v[RT - 1].z, wa.z, s1.z, s2.z);
        single_acc(gg.w, vb.w, v[RT - 1].w, wa.w, s1.w, s2.w);
      }
    } else {
      int r = 0;
#pragma unroll 2
      for (; r + 1 < R; r += 2) {
        const float4 va = ldg_stream4(vsp + (size_t)r * slab), vc = ldg_stream4(vsp + (size_t)(r + 1) * slab);
        const float4 wa = __ldg(reinterpret_cast<const float4*>(wp + (size_t)r * NP));
        const float4 wc = __ldg(reinterpret_cast<const float4*>(wp + (size_t)(r + 1) * NP));
        pair_acc(gg.x, vb.x, va.x, vc.x, wa.x, wc.x, s1.x, s2.x);
        pair_acc(gg.y, vb.y, va.y, vc.y, wa.y, wc.y, s1.y, s2.y);
        pair_acc(gg.z, vb.z, va.z, vc.z, wa.z, wc.z, s1.z, s2.z);
        pair_acc(gg.w, vb.w, va.w, vc.w, wa.w, wc.w, s1.w, s2.w);
      }
      if (r < R) {
        const float4 va = ldg_stream4(vsp + (size_t)r * slab);
        const float4 wa = __ldg(reinterpret_cast<const float4*>(wp + (size_t)r * NP));
        single_acc(gg.x, vb.x, va.x, wa.x, s1.x, s2.x);
        single_acc(gg.y, vb.y, va.y, wa.y, s1.y, s2.y);
        single_acc(gg.z, vb.z, va.z, wa.z, s1.z, s2.z);
        single_acc(gg.w, vb.w, va.w, wa.w, s1.w, s2.w);
      }
    }
    float4 a = make_float4(x2.x * s2.x, x2.y * s2.y, x2.z * s2.z, x2.w * s2.w);
    // frames beyond the utterance (padding up to the 32-frame boundary) contribute nothing
    if (n4 + 0 >= N) { a.x = 0.f; s1.x = 0.f; }
    if (n4 + 1 >= N) { a.y = 0.f; s1.y = 0.f; }
    if (n4 + 2 >= N) { a.z = 0.f; s1.z = 0.f; }
    if (n4 + 3 >= N) { a.w = 0.f; s1.w = 0.f; }
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      if (k < K) {
        const float4 h = __ldg(reinterpret_cast<const float4*>(H + (size_t)k * NP + n_begin + n4));
        num[k] = fmaf(a.x, h.x, fmaf(a.y, h.y, fmaf(a.z, h.z, fmaf(a.w, h.w, num[k]))));
        den[k] = fmaf(s1.x, h.x, fmaf(s1.y, h.y, fmaf(s1.z, h.z, fmaf(s1.w, h.w, den[k]))));
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
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    if (k < K && lane == k) {
      const size_t o = ((size_t)b * F + f) * K + k;
      Wun[o] = W[o] * sqrtf(num[k] / den[k]);
    }
  }
}

// ------------------------------------------- column sweep: H, Vb, normalisation, g, cost
struct ColsArgs {
  int F, K, KS, NP, R, RS, B, ntiles, nchunk;
  const int32_t* frame_utt; const int32_t* frame_off;
  const float* X2; const float* Vs; const float* Vs_w;
  float* Vb; float* g; float* H; const float* Wun; float* W; float* cost_part;
};

// shared memory (floats): data[F][RS][NB] (slot r of bin f at [f][r], X2 at [f][R]) | W_s[F][KS] |
// red[8][2*KMAX][NB] | red2[2*KMAX][NB] | Hn_s[KMAX][NB] | meta[(KMAX+2+R)][NB] | wts_s[R][NB] | cn_s[KMAX] | misc[32]
__host__ __device__ inline size_t cols_smem_floats(int F, int KS, int R, int RS, int KMAX) {
  return (size_t)F * RS * NB + (size_t)F * KS + (size_t)8 * 2 * KMAX * NB + (size_t)2 * KMAX * NB + (size_t)KMAX * NB +
         (size_t)(KMAX + 2 + R) * NB + (size_t)R * NB + KMAX + 32;
}

template <int KMAX, int RT>
__global__ void __launch_bounds__(CT, 1) k_cols_v1(ColsArgs p) {
  extern __shared__ __align__(16) float sm[];
  const int F = p.F, K = p.K, KS = p.KS, NP = p.NP;
  const int R = RT > 0 ? RT : p.R;
  const int RS = RT > 0 ? ((RT + 1) | 1) : p.RS;            // row stride in 8-float groups (odd: conflict-free)
  const int FS = RS * NB;                                   // floats per frequency row
  float* data = sm;
  float* W_s = data + (size_t)F * FS;
  float* red = W_s + (size_t)F * KS;
  float* red2 = red + 8 * 2 * KMAX * NB;
  float* Hn_s = red2 + 2 * KMAX * NB;
  float* meta = Hn_s + KMAX * NB;                           // H_old [KMAX][NB] | g [NB] | frame_utt [NB] | Vs_w [R][NB]
  float* wts_s = meta + (KMAX + 2 + R) * NB;                // [R][NB] multiplicities of the current tile
  float* cn_s = wts_s + R * NB;
  float* misc = cn_s + KMAX;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = tid & (NB - 1), fl = tid >> 3;              // column of the tile, frequency lane (0..31)
  const int NI = (F + 31) / 32, NCH = p.nchunk;

  // contiguous tile range of this CTA, walked from the END of the batch: the W sweep that ran
  // just before leaves the tail of Vs in L2
  const int per = (p.ntiles + gridDim.x - 1) / gridDim.x;
  const int c_rev = gridDim.x - 1 - blockIdx.x;
  int t_hi = p.ntiles - c_rev * per, t_lo = t_hi - per;
  if (t_lo < 0) t_lo = 0;
  if (t_hi <= 0) return;

  auto issue_chunk = [&](int t, int j) {
    const int i0 = j * NI / NCH, i1 = (j + 1) * NI / NCH;
    const int r0 = i0 * 32, r1 = min(i1 * 32, F);
    const int rows2 = (r1 - r0) * 2;
    const size_t col = (size_t)t * NB;
    const int total = rows2 * (R + 1);
    const bool pow2 = rows2 == 128;
    for (int q = tid; q < total; q += CT) {
      const int plane = pow2 ? (q >> 7) : q / rows2;
      const int rem = pow2 ? (q & 127) : q - plane * rows2;
      const int rowi = r0 + (rem >> 1), half = rem & 1;
      const float* src = (plane < R ? p.Vs + ((size_t)plane * F + rowi) * NP : p.X2 + (size_t)rowi * NP) + col + 4 * half;
      cp16(data + (size_t)rowi * FS + plane * NB + 4 * half, src);
    }
    if (j == 0) {   // tile meta data: H_old rows, g, frame_utt, slot multiplicities
      const int nH = 2 * K, nW = 2 * R;
      if (tid < nH) cp16(meta + (tid >> 1) * NB + 4 * (tid & 1), p.H + (size_t)(tid >> 1) * NP + col + 4 * (tid & 1));
      else if (tid < nH + 2) cp16(meta + KMAX * NB + 4 * (tid - nH), p.g + col + 4 * (tid - nH));
      else if (tid < nH + 4) cp16(meta + (KMAX + 1) * NB + 4 * (tid - nH - 2), p.frame_utt + col + 4 * (tid - nH - 2));
      else if (tid < nH + 4 + nW) {
        const int q = tid - nH - 4;
        cp16(meta + (KMAX + 2 + (q >> 1)) * NB + 4 * (q & 1), p.Vs_w + (size_t)(q >> 1) * NP + col + 4 * (q & 1));
      }
    }
    cp_commit();
  };

  // walk the range downwards
  for (int j = 0; j < NCH; ++j) issue_chunk(t_hi - 1, j);
  int cur_b = -1, cur_fo = -1;
  for (int t = t_hi - 1; t >= t_lo; --t) {
    const bool has_next = t - 1 >= t_lo;
    cp_wait_dyn(NCH - 1);
    __syncthreads();
    const int b = __float_as_int(meta[(KMAX + 1) * NB]);
    if (b < 0) {                                            // tile entirely in padding
      cp_wait_dyn(0);
      __syncthreads();
      if (tid == 0) p.cost_part[t] = 0.f;
      if (has_next) for (int j = 0; j < NCH; ++j) issue_chunk(t - 1, j);
      continue;
    }
    const bool valid = __float_as_int(meta[(KMAX + 1) * NB + n]) >= 0;
    const float gg = meta[KMAX * NB + n];
    float hk[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) hk[k] = (k < K) ? meta[k * NB + n] : 0.f;
    // multiplicities of the slots of this column: registers when R is a compile-time constant,
    // a private shared copy otherwise (`meta` is overwritten by the next tile's copies)
    float wr[RT > 0 ? RT : 1];
    if (RT > 0) {
#pragma unroll
      for (int r = 0; r < RT; ++r) wr[r] = meta[(KMAX + 2 + r) * NB + n];
    } else {
      for (int i = tid; i < R * NB; i += CT) wts_s[i] = meta[(KMAX + 2) * NB + i];
    }
    const float* wsm = wts_s + n;
    auto wgt = [&](int r) -> float { return RT > 0 ? wr[RT > 0 ? r : 0] : wsm[r * NB]; };

    if (b != cur_b) {                                       // new utterance: dictionary -> smem, column norms
      cur_b = b;
      cur_fo = p.frame_off[b];
      const float* wsrc = p.Wun + (size_t)b * F * K;
      for (int i = tid; i < F * K; i += CT) { const int ff = i / K; W_s[ff * KS + (i - ff * K)] = __ldg(wsrc + i); }
      if (KS > K) for (int i = tid; i < F * (KS - K); i += CT) { const int ff = i / (KS - K); W_s[ff * KS + K + (i - ff * (KS - K))] = 0.f; }
      __syncthreads();
      for (int k = warp; k < K; k += CT / 32) {             // c_k = sum_f |W_fk|  (mcem.py:128)
        float s = 0.f;
        for (int ff = lane; ff < F; ff += 32) s += fabsf(W_s[ff * KS + k]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) cn_s[k] = s;
      }
    }
    __syncthreads();                                        // W_s, cn_s, wts_s visible
    if (t * NB == cur_fo) {                                 // the first tile of an utterance writes W / c (mcem.py:131)
      float* wdst = p.W + (size_t)b * F * K;
      for (int i = tid; i < F * K; i += CT) { const int ff = i / K, k = i - ff * K; wdst[i] = W_s[ff * KS + k] / cn_s[k]; }
    }

    // ---------------- pass A: H update with Vb = Wun @ H_old (mcem.py:113-121), chasing the copies
    float num[KMAX], den[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) num[k] = den[k] = 0.f;
    for (int j = 0; j < NCH; ++j) {
      if (j > 0) { cp_wait_dyn(NCH - 1 - j); __syncthreads(); }
      const int i0 = j * NI / NCH, i1 = (j + 1) * NI / NCH;
      const float* vs = data + (size_t)(i0 * 32 + fl) * FS + n;
      const float* wrow = W_s + (i0 * 32 + fl) * KS;
      for (int i = i0; i < i1; ++i, vs += 32 * FS, wrow += 32 * KS) {
        if (i * 32 + fl < F) {
          float w[KMAX];
#pragma unroll
          for (int k4 = 0; k4 < KMAX; k4 += 4) {
            const float4 t4 = *reinterpret_cast<const float4*>(wrow + k4);
            w[k4] = t4.x; w[k4 + 1] = t4.y; w[k4 + 2] = t4.z; w[k4 + 3] = t4.w;
          }
          float vb = 0.f;
#pragma unroll
          for (int k = 0; k < KMAX; ++k) vb = fmaf(w[k], hk[k], vb);
          float s1 = 0.f, s2 = 0.f;
          if (RT > 0) {
#pragma unroll
            for (int r = 0; r + 1 < RT; r += 2) pair_acc(gg, vb, vs[r * NB], vs[(r + 1) * NB], wgt(r), wgt(r + 1), s1, s2);
            if (RT & 1) single_acc(gg, vb, vs[(RT - 1) * NB], wgt(RT - 1), s1, s2);
          } else {
            int r = 0;
            for (; r + 1 < R; r += 2) pair_acc(gg, vb, vs[r * NB], vs[(r + 1) * NB], wgt(r), wgt(r + 1), s1, s2);
            if (r < R) single_acc(gg, vb, vs[r * NB], wgt(r), s1, s2);
          }
          const float a = vs[R * NB] * s2;
#pragma unroll
          for (int k = 0; k < KMAX; ++k) { num[k] = fmaf(w[k], a, num[k]); den[k] = fmaf(w[k], s1, den[k]); }
        }
      }
    }
    // reduce over the 32 frequency lanes: lanes (8,16) in the warp, then the 8 warps
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      num[k] += __shfl_xor_sync(0xffffffffu, num[k], 8);
      num[k] += __shfl_xor_sync(0xffffffffu, num[k], 16);
      den[k] += __shfl_xor_sync(0xffffffffu, den[k], 8);
      den[k] += __shfl_xor_sync(0xffffffffu, den[k], 16);
    }
    if (lane < NB) {
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        red[(warp * 2 * KMAX + k) * NB + lane] = num[k];
        red[(warp * 2 * KMAX + KMAX + k) * NB + lane] = den[k];
      }
    }
    __syncthreads();
    if (tid < 2 * KMAX * NB) {
      float s = 0.f;
#pragma unroll
      for (int w8 = 0; w8 < 8; ++w8) s += red[w8 * 2 * KMAX * NB + tid];
      red2[tid] = s;
    }
    __syncthreads();
    if (tid < KMAX * NB) {
      const int k = tid >> 3;
      Hn_s[tid] = (k < K) ? meta[tid] * sqrtf(red2[tid] / red2[KMAX * NB + tid]) : 0.f;
    }
    __syncthreads();
    float hn[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) hn[k] = Hn_s[k * NB + n];

    // ---------------- pass B: Vb = Wun @ H_new (kept for the next E-step, mcem.py:124); g update (:138-142)
    float ng = 0.f, dg = 0.f;
    {
      const float* vs = data + (size_t)fl * FS + n;
      const float* wrow = W_s + fl * KS;
      float* vbo = p.Vb + (size_t)fl * NP + (size_t)t * NB + n;
      for (int i = 0; i < NI; ++i, vs += 32 * FS, wrow += 32 * KS, vbo += (size_t)32 * NP) {
        if (i * 32 + fl < F) {
          float vb = 0.f;
#pragma unroll
          for (int k4 = 0; k4 < KMAX; k4 += 4) {
            const float4 t4 = *reinterpret_cast<const float4*>(wrow + k4);
            vb = fmaf(t4.x, hn[k4], fmaf(t4.y, hn[k4 + 1], fmaf(t4.z, hn[k4 + 2], fmaf(t4.w, hn[k4 + 3], vb))));
          }
          if (valid) *vbo = vb;
          float t1 = 0.f, t2 = 0.f;
          auto pairB = [&](float va, float vc, float wa, float wc) {
            const float a = fmaf(gg, va, vb), c = fmaf(gg, vc, vb);
            const float ip = rcp_fast(a * c);
            const float ia = c * ip, ic = a * ip;
            const float ua = wa * va * ia, uc = wc * vc * ic;
            t1 += ua;
            t1 += uc;
            t2 = fmaf(ua, ia, t2);
            t2 = fmaf(uc, ic, t2);
          };
          auto singleB = [&](float va, float wa) {
            const float ia = rcp_fast(fmaf(gg, va, vb)), ua = wa * va * ia;
            t1 += ua;
            t2 = fmaf(ua, ia, t2);
          };
          if (RT > 0) {
#pragma unroll
            for (int r = 0; r + 1 < RT; r += 2) pairB(vs[r * NB], vs[(r + 1) * NB], wgt(r), wgt(r + 1));
            if (RT & 1) singleB(vs[(RT - 1) * NB], wgt(RT - 1));
          } else {
            int r = 0;
            for (; r + 1 < R; r += 2) pairB(vs[r * NB], vs[(r + 1) * NB], wgt(r), wgt(r + 1));
            if (r < R) singleB(vs[r * NB], wgt(r));
          }
          ng = fmaf(vs[R * NB], t2, ng);
          dg += t1;
        }
      }
    }
    ng += __shfl_xor_sync(0xffffffffu, ng, 8);
    ng += __shfl_xor_sync(0xffffffffu, ng, 16);
    dg += __shfl_xor_sync(0xffffffffu, dg, 8);
    dg += __shfl_xor_sync(0xffffffffu, dg, 16);
    if (lane < NB) { red[(warp * 2) * NB + lane] = ng; red[(warp * 2 + 1) * NB + lane] = dg; }
    __syncthreads();
    float sn = 0.f, sd = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) { sn += red[(w8 * 2) * NB + n]; sd += red[(w8 * 2 + 1) * NB + n]; }
    const float gnew = gg * sqrtf(sn / sd);
    // outputs of the tile (registers hold everything still needed from `meta`)
    if (tid < NB && valid) p.g[(size_t)t * NB + n] = gnew;
    if (tid < KMAX * NB) {
      const int k = tid >> 3;
      if (k < K && valid) p.H[(size_t)k * NP + (size_t)t * NB + n] = Hn_s[tid] * cn_s[k];   // mcem.py:133
    }

    // ---------------- pass C: cost with the new g (mcem.py:151-152, :68-70); refill behind it
    float cl = 0.f, cr = 0.f;
    for (int j = 0; j < NCH; ++j) {
      const int i0 = j * NI / NCH, i1 = (j + 1) * NI / NCH;
      const float* vs = data + (size_t)(i0 * 32 + fl) * FS + n;
      const float* wrow = W_s + (i0 * 32 + fl) * KS;
      for (int i = i0; i < i1; ++i, vs += 32 * FS, wrow += 32 * KS) {
        if (i * 32 + fl < F) {
          float vb = 0.f;
#pragma unroll
          for (int k4 = 0; k4 < KMAX; k4 += 4) {
            const float4 t4 = *reinterpret_cast<const float4*>(wrow + k4);
            vb = fmaf(t4.x, hn[k4], fmaf(t4.y, hn[k4 + 1], fmaf(t4.z, hn[k4 + 2], fmaf(t4.w, hn[k4 + 3], vb))));
          }
          float sl = 0.f, sr = 0.f;
          auto pairC = [&](float va, float vc, float wa, float wc) {
            const float a = fmaf(gnew, va, vb), c = fmaf(gnew, vc, vb);
            sl = fmaf(wa, lg2_fast(a), sl);
            sl = fmaf(wc, lg2_fast(c), sl);
            sr = fmaf(fmaf(wa, c, wc * a), rcp_fast(a * c), sr);
          };
          auto singleC = [&](float va, float wa) {
            const float a = fmaf(gnew, va, vb);
            sl = fmaf(wa, lg2_fast(a), sl);
            sr = fmaf(wa, rcp_fast(a), sr);
          };
          if (RT > 0) {
#pragma unroll
            for (int r = 0; r + 1 < RT; r += 2) pairC(vs[r * NB], vs[(r + 1) * NB], wgt(r), wgt(r + 1));
            if (RT & 1) singleC(vs[(RT - 1) * NB], wgt(RT - 1));
          } else {
            int r = 0;
            for (; r + 1 < R; r += 2) pairC(vs[r * NB], vs[(r + 1) * NB], wgt(r), wgt(r + 1));
            if (r < R) singleC(vs[r * NB], wgt(r));
          }
          cl += sl;
          cr = fmaf(vs[R * NB], sr, cr);
        }
      }
      __syncthreads();                                      // chunk j is free (and `red` of pass B is consumed)
      if (has_next) issue_chunk(t - 1, j);
    }
    float cs = valid ? fmaf(cl, 0.6931471805599453f, cr) : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cs += __shfl_xor_sync(0xffffffffu, cs, o);
    if (lane == 0) misc[warp] = cs;
    __syncthreads();
    if (tid == 0) {
      float s = 0.f;
      for (int w8 = 0; w8 < 8; ++w8) s += misc[w8];
      p.cost_part[t] = s;
    }
  }
  cp_wait_dyn(0);
}

template <int KMAX, int RT>
int32_t launch_cols(const ColsArgs& a, size_t smem, int grid, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(k_cols_v1<KMAX, RT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return fail(GVN_E_CUDA, "k_cols_v1 smem attr (%zu B): %s", smem, cudaGetErrorString(e));
  k_cols_v1<KMAX, RT><<<grid, CT, smem, st>>>(a);
  return check_launch("k_cols_v1");
}

inline int kmax_of(int K) { return (K + 3) / 4 * 4; }
inline int ks_of(int K) { int ks = kmax_of(K); return (ks % 16 == 0) ? ks + 4 : ks; }
inline int rs_of(int R) { return (R + 1) | 1; }

template <int KMAX>
int32_t launch_v1_k(const gvn_batch* b, int R, const ColsArgs& a, size_t smem, int grid, cudaStream_t st) {
  dim3 gw((b->F + CT / 32 - 1) / (CT / 32), b->B);
  if (R == 10)
    k_w_v1<KMAX, 10><<<gw, CT, 0, st>>>(b->F, b->K, b->NP, R, b->frame_off, b->n_frames, b->X2, b->Vs, b->Vs_w, b->Vb, b->g, b->H, b->W, b->Wun);
  else
    k_w_v1<KMAX, 0><<<gw, CT, 0, st>>>(b->F, b->K, b->NP, R, b->frame_off, b->n_frames, b->X2, b->Vs, b->Vs_w, b->Vb, b->g, b->H, b->W, b->Wun);
  int32_t rc = check_launch("k_w_v1");
  if (rc) return rc;
  return R == 10 ? launch_cols<KMAX, 10>(a, smem, grid, st) : launch_cols<KMAX, 0>(a, smem, grid, st);
}

}  // namespace

// true when variant 1 can run this shape (the tile block must fit in shared memory)
bool mstep_v1_supported(const gvn_batch* b, int R) {
  if (b->K > 16) return false;
  const size_t bytes = cols_smem_floats(b->F, ks_of(b->K), R, rs_of(R), kmax_of(b->K)) * 4;
  return bytes <= 227 * 1024;
}

int32_t launch_mstep_v1(const gvn_batch* b, int R, float* cost_part, cudaStream_t st) {
  const int KMAX = kmax_of(b->K);
  ColsArgs a;
  a.F = b->F; a.K = b->K; a.KS = ks_of(b->K); a.NP = b->NP; a.R = R; a.RS = rs_of(R); a.B = b->B;
  a.ntiles = b->NP / NB;
  const int NI = (b->F + 31) / 32;
  a.nchunk = NI < MAXCH ? NI : MAXCH;
  a.frame_utt = b->frame_utt; a.frame_off = b->frame_off; a.X2 = b->X2; a.Vs = b->Vs; a.Vs_w = b->Vs_w;
  a.Vb = b->Vb; a.g = b->g; a.H = b->H; a.Wun = b->Wun; a.W = b->W; a.cost_part = cost_part;
  const size_t smem = cols_smem_floats(b->F, a.KS, R, a.RS, KMAX) * 4;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = a.ntiles < sms ? a.ntiles : sms;
  switch (KMAX) {
    case 4: return launch_v1_k<4>(b, R, a, smem, grid, st);
    case 8: return launch_v1_k<8>(b, R, a, smem, grid, st);
    case 12: return launch_v1_k<12>(b, R, a, smem, grid, st);
    default: return launch_v1_k<16>(b, R, a, smem, grid, st);
  }
}

}  // namespace gvn
