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
#include <stdlib.h>

#include "gvn_common.cuh"
#include "tc_common.cuh"

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

// ---- packed fp32x2 arithmetic (FFMA2 / FMUL2 / FADD2 of sm_100): the column sweep is bound by the
// FMA pipe, where a 3-register FFMA sustains ~0.6 warp-instructions per clock and scheduler
// (register-bank limit, tools/ffma2_bench.cu) while FFMA2 carries two per instruction at ~0.45
typedef float2 f2;
__device__ __forceinline__ f2 F2(float x, float y) { return make_float2(x, y); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ f2 rcp2(f2 a) { return make_float2(rcp_fast(a.x), rcp_fast(a.y)); }
__device__ __forceinline__ f2 lg2_2(f2 a) { return make_float2(lg2_fast(a.x), lg2_fast(a.y)); }

// four slots at once: slots (r0, r1) play `a`, slots (r2, r3) play `c` of two independent pairs
__device__ __forceinline__ void quad_acc(f2 g2, f2 vb2, f2 va, f2 vc, f2 wa, f2 wc, f2& s1, f2& s2) {
  const f2 a = fma2(g2, va, vb2), c = fma2(g2, vc, vb2);
  const f2 ip = rcp2(mul2(a, c));
  const f2 t1 = mul2(wa, c), t2 = mul2(wc, a);            // w/a = w * c * 1/(ac)
  s1 = fma2(add2(t1, t2), ip, s1);
  s2 = fma2(fma2(t1, c, mul2(t2, a)), mul2(ip, ip), s2);
}
__device__ __forceinline__ void quad_accB(f2 g2, f2 vb2, f2 va, f2 vc, f2 wa, f2 wc, f2& t1, f2& t2) {
  const f2 a = fma2(g2, va, vb2), c = fma2(g2, vc, vb2);
  const f2 ip = rcp2(mul2(a, c));
  const f2 ia = mul2(c, ip), ic = mul2(a, ip);
  const f2 ua = mul2(mul2(wa, va), ia), uc = mul2(mul2(wc, vc), ic);
  t1 = add2(t1, add2(ua, uc));
  t2 = fma2(ua, ia, fma2(uc, ic, t2));
}
__device__ __forceinline__ void quad_accC(f2 g2, f2 vb2, f2 va, f2 vc, f2 wa, f2 wc, f2& sl, f2& sr) {
  const f2 a = fma2(g2, va, vb2), c = fma2(g2, vc, vb2);
  sl = fma2(wa, lg2_2(a), fma2(wc, lg2_2(c), sl));
  sr = fma2(fma2(wa, c, mul2(wc, a)), rcp2(mul2(a, c)), sr);
}

// ---- bulk async copies (TMA engine, no tensor map) + mbarrier -------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(s_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) { tc::mbar_wait(bar, parity); }   // time-bounded spin: tc_common.cuh
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s_u32(dst)),
               "l"(src), "r"(bytes), "r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// ------------------------------------------------------------------ column data in tile order
// Mt[t][n][MS]: one record of MS floats per frame n of column tile t -- H[0..K), g, the R slot multiplicities, padding --
// so a stage needs ONE bulk copy for the column data of a tile and a thread reads its frame's record with 16-byte loads.
// MS = 4 * (odd number): the records of the 8 frames of a tile start in different bank groups.
__host__ __device__ inline int meta_stride(int K, int R) {
  int m = (K + 1 + R + 3) / 4;
  if ((m & 1) == 0) ++m;
  return 4 * m;
}

__global__ void __launch_bounds__(256) k_tile_meta(int K, int R, int NP, const float* __restrict__ H, const float* __restrict__ g,
                                                   const float* __restrict__ Vs_w, float* __restrict__ Mt) {
  __shared__ float rec[256 * 29 + 32];                        // records of 256 columns, staged so that the writes are coalesced
  const int MS = meta_stride(K, R), MR = K + 1 + R;
  const int c0 = blockIdx.x * 256, col = c0 + threadIdx.x;
  const bool staged = MS <= 29;
  if (col < NP) {
    float* dst = staged ? rec + threadIdx.x * MS : Mt + (size_t)col * MS;
    for (int m = 0; m < K; ++m) dst[m] = H[(size_t)m * NP + col];
    dst[K] = g[col];
    for (int r = 0; r < R; ++r) dst[K + 1 + r] = Vs_w[(size_t)r * NP + col];
    for (int m = MR; m < MS; ++m) dst[m] = 0.f;
  }
  if (staged) {
    __syncthreads();
    const int ncol = min(256, NP - c0);
    for (int i = threadIdx.x; i < ncol * MS; i += 256) Mt[(size_t)c0 * MS + i] = rec[i];
  }
}

// ------------------------------------------------------------------ W update (mcem.py:105-110)
// One CTA per (utterance, block of 64 frequency rows); compute thread = (row, column n of an
// 8-frame tile).  The [rows][8] blocks of the R sample slots and of X2 (column-tile order:
// contiguous rows*32 bytes each) plus the tile's column data stream through a ring of WS stages
// filled by bulk async copies from a dedicated producer warp, so the bytes in flight live in
// shared memory, not in registers, and the compute warps never meet at a CTA-wide barrier.
// Vb of the reference is W @ H here (the stored Vb is the same product before the
// normalisation, mcem.py:124-133).
constexpr int WROWS = 64;             // frequency rows per CTA
constexpr int WCT = WROWS * NB;       // 512 compute threads
constexpr int WT = WCT + 32;          // + producer warp
constexpr int WS = 4;                 // ring stages at most (two CTAs per SM: 8 stages of ~23 KB in flight at R = 10)

// floats of one ring stage: (R+1) planes of [WROWS][NB] + column data (H rows, g, multiplicities)
__host__ __device__ inline int w_stage_floats(int K, int R) { return ((R + 1) * WROWS * NB + meta_stride(K, R) * NB + 31) / 32 * 32; }

// KX > 0: the rank is exactly KX (10 in every evaluate script): no guards, no padded columns in the rank-K loops
template <int KMAX, int RT, int KX = 0>
__global__ void __launch_bounds__(WT, (KMAX > 16 ? 1 : 2)) k_w_v2(int F, int K, int NP, int R_rt, int ws, const int32_t* __restrict__ frame_off,
                                                const int32_t* __restrict__ n_frames, const float* __restrict__ X2t,
                                                const float* __restrict__ Vs, const float* __restrict__ Mt,
                                                const float* __restrict__ W, float* __restrict__ Wun, float* __restrict__ Wpart) {
  extern __shared__ __align__(128) float smw[];             // [WS][stage]
  __shared__ __align__(8) uint64_t full[WS], empty[WS];
  const int R = RT > 0 ? RT : R_rt;
  const int tid = threadIdx.x, lane = tid & 31;
  const int b = blockIdx.y, f0 = blockIdx.x * WROWS;
  const int rows = min(WROWS, F - f0);
  const int n_begin = frame_off[b], N = n_frames[b];
  // blockIdx.z splits the frames of the utterance (few long utterances would otherwise leave most SMs idle): this
  // CTA sweeps tiles [tb, te) and, when split, leaves its partial sums in Wpart for k_w_finish
  const int ntile_all = (N + NB - 1) / NB, nsplit = gridDim.z;
  const int per_split = (ntile_all + nsplit - 1) / nsplit;
  const int tb = min((int)blockIdx.z * per_split, ntile_all), te = min(tb + per_split, ntile_all);
  const int ntile = te - tb, t0 = n_begin / NB + tb, T8 = NP / NB;
  constexpr int PSt = WROWS * NB;                           // plane stride inside a stage (floats)
  const int MO = (R + 1) * PSt;                             // offset of the column data inside a stage
  const int MS = meta_stride(K, R);                         // floats per frame record of the column data
  const int SSt = w_stage_floats(K, R);

  if (tid == 0) {
    for (int s = 0; s < ws; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, WCT / 32); }
    fence_mbar_init();
  }
  __syncthreads();

  if (tid >= WCT) {
    // ================= producer warp: lane p copies plane p, lane R+1 the column data =================
    const uint32_t blk = (uint32_t)rows * NB * 4;
    for (int ti = 0; ti < ntile; ++ti) {
      const int s = ti % ws;
      if (ti >= ws) mbar_wait(empty + s, ((ti / ws) - 1) & 1);
      float* dst = smw + (size_t)s * SSt;
      if (lane == 0) mbar_expect_tx(full + s, blk * (R + 1) + (uint32_t)MS * NB * 4);
      __syncwarp();
      for (int c = lane; c <= R + 1; c += 32) {
        if (c < R) bulk_g2s(dst + c * PSt, Vs + (((size_t)c * T8 + t0 + ti) * F + f0) * NB, blk, full + s);
        else if (c == R) bulk_g2s(dst + R * PSt, X2t + ((size_t)(t0 + ti) * F + f0) * NB, blk, full + s);
        else bulk_g2s(dst + MO, Mt + (size_t)(t0 + ti) * MS * NB, (uint32_t)MS * NB * 4, full + s);
      }
    }
    return;
  }

  // ================= compute warps =================
  const int n = tid & (NB - 1), rl = tid >> 3, f = f0 + rl;
  const bool rowok = rl < rows;
  if constexpr (KMAX > 16) {
    // Large ranks: a thread cannot carry 2*K accumulators.  Phase 1, thread = (row, frame): a = X2 * s2 and s1 of its
    // point; the four rows of a warp exchange them through 256 bytes of shared memory.  Phase 2, the same warp
    // regrouped as (row, group of KMAX/8 dictionary columns): num/den of those columns summed over the 8 frames of
    // the stage -- four (KMAX = 32) accumulator pairs per thread, already reduced over frames.
    __shared__ __align__(16) float xch[WCT / 32][2][4][NB];
    constexpr int KG = KMAX / 8;                            // dictionary columns per thread in phase 2
    const int wi = tid >> 5, rw = lane >> 3, kg = lane & 7;
    float w[KMAX], num[KG], den[KG];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) w[k] = (k < K && rowok) ? W[((size_t)b * F + f) * K + k] : 0.f;
#pragma unroll
    for (int i = 0; i < KG; ++i) num[i] = den[i] = 0.f;
    for (int ti = 0; ti < ntile; ++ti) {
      const int s = ti % ws;
      mbar_wait(full + s, (ti / ws) & 1);
      const float* st = smw + (size_t)s * SSt;
      const float* mt0 = st + MO;
      float a = 0.f, s1 = 0.f;
      if (rowok) {
        const float* vs = st + rl * NB + n;
        const float* mt = mt0 + n * MS;                       // record of frame n
        float vb0 = 0.f, vb1 = 0.f;
#pragma unroll
        for (int k = 0; k < KMAX; k += 2) {                   // entries beyond K of the record are not H
          if (k < K) vb0 = fmaf(w[k], mt[k], vb0);
          if (k + 1 < K) vb1 = fmaf(w[k + 1], mt[k + 1], vb1);
        }
        const float vb = vb0 + vb1, gg = mt[K];
        const float* wt = mt + K + 1;
        float s2 = 0.f;
        int r = 0;
        for (; r + 1 < R; r += 2) pair_acc(gg, vb, vs[r * PSt], vs[(r + 1) * PSt], wt[r], wt[r + 1], s1, s2);
        if (r < R) single_acc(gg, vb, vs[r * PSt], wt[r], s1, s2);
        a = vs[R * PSt] * s2;
        if ((tb + ti) * NB + n >= N) { a = 0.f; s1 = 0.f; }
      }
      __syncwarp();                                         // phase 2 of the previous stage has read the exchange area
      xch[wi][0][rw][n] = a;
      xch[wi][1][rw][n] = s1;
      __syncwarp();
      float av[NB], sv[NB];
#pragma unroll
      for (int j4 = 0; j4 < NB; j4 += 4) {
        const float4 t0_ = *reinterpret_cast<const float4*>(&xch[wi][0][rw][j4]), t1_ = *reinterpret_cast<const float4*>(&xch[wi][1][rw][j4]);
        av[j4] = t0_.x; av[j4 + 1] = t0_.y; av[j4 + 2] = t0_.z; av[j4 + 3] = t0_.w;
        sv[j4] = t1_.x; sv[j4 + 1] = t1_.y; sv[j4 + 2] = t1_.z; sv[j4 + 3] = t1_.w;
      }
      static_assert(KG == 4, "phase 2 reads its dictionary columns as one float4 per frame");
#pragma unroll
      if (kg * KG < K)                                      // groups entirely beyond the rank have nothing to do (and would read past the record)
      for (int j = 0; j < NB; ++j) {                        // H[4 kg .. 4 kg + 3] of frame j (entries beyond K are finite and never written out)
        const float4 h = *reinterpret_cast<const float4*>(mt0 + j * MS + kg * KG);
        num[0] = fmaf(av[j], h.x, num[0]); den[0] = fmaf(sv[j], h.x, den[0]);
        num[1] = fmaf(av[j], h.y, num[1]); den[1] = fmaf(sv[j], h.y, den[1]);
        num[2] = fmaf(av[j], h.z, num[2]); den[2] = fmaf(sv[j], h.z, den[2]);
        num[3] = fmaf(av[j], h.w, num[3]); den[3] = fmaf(sv[j], h.w, den[3]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + s);                // this warp is done with stage s
    }
    const int f2_ = f0 + (wi * 4 + rw);                     // the row this lane holds sums for
    if (wi * 4 + rw < rows) {
#pragma unroll
      for (int i = 0; i < KG; ++i) {
        const int k = kg * KG + i;
        if (k < K) {
          if (nsplit == 1) {
            Wun[((size_t)b * F + f2_) * K + k] = W[((size_t)b * F + f2_) * K + k] * sqrtf(num[i] / den[i]);
          } else {
            float* dst = Wpart + ((((size_t)b * nsplit + blockIdx.z) * F + f2_) * K) * 2;
            dst[2 * k] = num[i]; dst[2 * k + 1] = den[i];
          }
        }
      }
    }
    return;
  }
  constexpr int KE = KX > 0 ? KX : KMAX;                    // columns the rank-K loops run over
  float w[KE], num[KE], den[KE];
#pragma unroll
  for (int k = 0; k < KE; ++k) { w[k] = (k < K && rowok) ? W[((size_t)b * F + f) * K + k] : 0.f; num[k] = den[k] = 0.f; }

  for (int ti = 0; ti < ntile; ++ti) {
    const int s = ti % ws;
    mbar_wait(full + s, (ti / ws) & 1);
    if (rowok) {
      const float* st = smw + (size_t)s * SSt;
      const float* vs = st + rl * NB + n;
      const float* mt = st + MO + n * MS;                   // record of frame n: H | g | multiplicities
      constexpr bool VEC = KX > 0 && RT > 0;                // all offsets known: the record is read with 16-byte loads
      constexpr int NV = VEC ? (KX + 1 + RT + 3) / 4 * 4 : 4;
      float md[NV];
      if (VEC) {
#pragma unroll
        for (int q = 0; q < NV; q += 4) {
          const float4 t4 = *reinterpret_cast<const float4*>(mt + q);
          md[q] = t4.x; md[q + 1] = t4.y; md[q + 2] = t4.z; md[q + 3] = t4.w;
        }
      }
      auto hval = [&](int k) -> float { return VEC ? md[VEC ? k : 0] : mt[k]; };
      auto wval = [&](int r) -> float { return VEC ? md[VEC ? KX + 1 + r : 0] : mt[K + 1 + r]; };
      float vb = 0.f;
#pragma unroll
      for (int k = 0; k < KE; ++k) if (KX > 0 || k < K) vb = fmaf(w[k], hval(k), vb);
      const float gg = VEC ? md[VEC ? KX : 0] : mt[K];
      float s1 = 0.f, s2 = 0.f;
      if (RT > 0) {
#pragma unroll
        for (int r = 0; r + 1 < RT; r += 2) pair_acc(gg, vb, vs[r * PSt], vs[(r + 1) * PSt], wval(r), wval(r + 1), s1, s2);
        if (RT & 1) single_acc(gg, vb, vs[(RT - 1) * PSt], wval(RT - 1), s1, s2);
      } else {
        int r = 0;
        for (; r + 1 < R; r += 2) pair_acc(gg, vb, vs[r * PSt], vs[(r + 1) * PSt], wval(r), wval(r + 1), s1, s2);
        if (r < R) single_acc(gg, vb, vs[r * PSt], wval(r), s1, s2);
      }
      float a = vs[R * PSt] * s2;
      if ((tb + ti) * NB + n >= N) { a = 0.f; s1 = 0.f; }   // padding frames of the last tile contribute nothing
#pragma unroll
      for (int k = 0; k < KE; ++k) if (KX > 0 || k < K) { const float h = hval(k); num[k] = fmaf(a, h, num[k]); den[k] = fmaf(s1, h, den[k]); }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty + s);                  // this warp is done with stage s
  }
  // reduce over the 8 columns of the tile (adjacent lanes), lane n == 0 writes the row
#pragma unroll
  for (int k = 0; k < KE; ++k) {
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      num[k] += __shfl_xor_sync(0xffffffffu, num[k], o);
      den[k] += __shfl_xor_sync(0xffffffffu, den[k], o);
    }
  }
  if (n == 0 && rowok) {
    if (nsplit == 1) {
#pragma unroll
      for (int k = 0; k < KE; ++k)
        if (k < K) Wun[((size_t)b * F + f) * K + k] = w[k] * sqrtf(num[k] / den[k]);
    } else {
      float* dst = Wpart + ((((size_t)b * nsplit + blockIdx.z) * F + f) * K) * 2;
#pragma unroll
      for (int k = 0; k < KE; ++k)
        if (k < K) { dst[2 * k] = num[k]; dst[2 * k + 1] = den[k]; }
    }
  }
}

// W <- W * sqrt(sum_z num / sum_z den) from the partial sums of a frame-split W sweep (fixed summation order)
__global__ void __launch_bounds__(256) k_w_finish(int B, int F, int K, int nsplit, const float* __restrict__ W,
                                                  const float* __restrict__ Wpart, float* __restrict__ Wun) {
  const size_t total = (size_t)B * F * K;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t b = i / ((size_t)F * K), fk = i - b * F * K;
    float nu = 0.f, de = 0.f;
    for (int z = 0; z < nsplit; ++z) {
      const float2 v = *reinterpret_cast<const float2*>(Wpart + (((b * nsplit + z) * F * K) + fk) * 2);
      nu += v.x; de += v.y;
    }
    Wun[i] = W[i] * sqrtf(nu / de);
  }
}

// ------------------------------------------- column sweep: H, Vb, normalisation, g, cost
struct ColsArgs {
  int F, K, KS, NP, R, B, ntiles, nchunk;
  const int32_t* frame_utt; const int32_t* frame_off;
  const float* X2t; const float* Vs; const float* Mt;
  float* Vb; float* g; float* H; const float* Wun; float* W; float* cost_part; uint32_t* XV;
};

constexpr int CCT = 384;              // compute threads of the column sweep: 12 warps (+1 producer) = at most 4 per scheduler
constexpr int CFL = CCT / NB;         // frequency lanes (rows handled per thread-loop iteration)
constexpr int CW = CCT / 32;          // compute warps
constexpr int CTT = CCT + 32;         // + producer warp

// shared memory (floats): data[(R+1)][F][NB] (plane R = X2) | W_s[F][KS] | red[8][2*KMAX][NB] |
// red2[2*KMAX][NB] | Hn_s[KMAX][NB] | meta[NB][meta_stride] | cn_s[KMAX] | misc[32] | bars
__host__ __device__ inline size_t cols_smem_floats(int F, int KS, int K, int R, int KMAX) {
  return (size_t)(R + 1) * F * NB + (size_t)F * KS + (size_t)CW * 2 * KMAX * NB + (size_t)2 * KMAX * NB + (size_t)KMAX * NB +
         (size_t)meta_stride(K, R) * NB + KMAX + 32 + 2 * (2 * MAXCH + 2) + 4 + 2 * (MAXCH + 1);
}

__device__ __forceinline__ void bar_compute() { asm volatile("bar.sync 1, %0;" ::"n"(CCT) : "memory"); }

// FT > 0: the number of frequency bins is a compile-time constant (513 in every evaluate script), so the slot offsets
// r * F * 8 of the inner loops become immediates instead of one integer multiply-add per slot and row
template <int KMAX, int RT, int FT = 0>
__global__ void __launch_bounds__(CTT, 1) k_cols_v1(ColsArgs p) {
  extern __shared__ __align__(128) float sm[];
  const int F = FT > 0 ? FT : p.F, K = p.K, NP = p.NP;
  constexpr int KS = (KMAX % 16 == 0) ? KMAX + 4 : KMAX;     // dictionary row stride in shared memory (= ks_of(K), the launcher sizes with it)
  const int R = RT > 0 ? RT : p.R;
  const int MS = meta_stride(K, R);
  const int PS = F * NB;                                    // plane stride (floats); a plane is one bulk-contiguous block
  float* data = sm;
  float* W_s = data + (size_t)(R + 1) * PS;
  float* red = W_s + (size_t)F * KS;
  float* red2 = red + CW * 2 * KMAX * NB;
  float* Hn_s = red2 + 2 * KMAX * NB;
  float* meta = Hn_s + KMAX * NB;                           // H_old [K][NB] | g [NB] | Vs_w [R][NB]   (one bulk copy)
  float* cn_s = meta + MS * NB;
  float* misc = cn_s + KMAX;
  uint64_t* bars = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(misc + 32) + 7) & ~(uintptr_t)7);
  uint64_t* full = bars;                                    // [MAXCH] chunk j of the tile has landed
  uint64_t* empty = bars + MAXCH;                           // [MAXCH] chunk j has been consumed by the cost pass
  uint64_t* mfull = bars + 2 * MAXCH;                       // column data landed
  uint64_t* mempty = mfull + 1;                             // column data consumed
  int* cb = reinterpret_cast<int*>(bars + 2 * MAXCH + 2);   // [NCH+1] chunk boundaries in units of 32 rows

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int NI = (F + CFL - 1) / CFL, NCH = p.nchunk, T8 = NP / NB;

  // contiguous tile range of this CTA, walked from the END of the batch: the W sweep that ran
  // just before leaves the tail of Vs in L2
  const int per = (p.ntiles + gridDim.x - 1) / gridDim.x;
  const int c_rev = gridDim.x - 1 - blockIdx.x;
  int t_hi = p.ntiles - c_rev * per, t_lo = t_hi - per;
  if (t_lo < 0) t_lo = 0;
  if (t_hi <= 0) return;

  if (tid == 0) {
    for (int j = 0; j < NCH; ++j) { mbar_init(full + j, 1); mbar_init(empty + j, CCT / 32); }
    mbar_init(mfull, 1);
    mbar_init(mempty, CCT / 32);
    for (int j = 0; j <= NCH; ++j) cb[j] = j * NI / NCH;
    fence_mbar_init();
  }
  __syncthreads();

  if (tid >= CCT) {
    // ================= producer warp: lane = plane; one bulk copy per plane and chunk =================
    uint32_t it = 0;
    for (int t = t_hi - 1; t >= t_lo; --t, ++it) {
      if (it > 0) mbar_wait(mempty, (it - 1) & 1);
      if (lane == 0) {
        mbar_expect_tx(mfull, (uint32_t)MS * NB * 4);
        bulk_g2s(meta, p.Mt + (size_t)t * MS * NB, (uint32_t)MS * NB * 4, mfull);
      }
      for (int j = 0; j < NCH; ++j) {
        if (it > 0) mbar_wait(empty + j, (it - 1) & 1);
        const int r0 = cb[j] * CFL, r1 = min(cb[j + 1] * CFL, F);
        const uint32_t bytes = (uint32_t)(r1 - r0) * NB * 4;
        if (lane == 0) mbar_expect_tx(full + j, bytes * (R + 1));
        __syncwarp();
        for (int pl = lane; pl <= R; pl += 32) {
          const float* src = pl < R ? p.Vs + (((size_t)pl * T8 + t) * F + r0) * NB : p.X2t + ((size_t)t * F + r0) * NB;
          bulk_g2s(data + (size_t)pl * PS + r0 * NB, src, bytes, full + j);
        }
      }
    }
    return;
  }

  // ================= compute warps =================
  const int n = tid & (NB - 1), fl = tid >> 3;              // column of the tile, frequency lane
  const int my_ni = (F - fl + CFL - 1) / CFL;               // rows of this thread: fl, fl + CFL, ... (no bound check inside the loops)
  int cur_b = -1, cur_fo = -1;
  uint32_t par = 0;                                         // phase parity of the barriers for this tile
  // utterance of the tile and validity of this thread's frame, fetched one tile ahead (an L2 round trip otherwise sits at
  // the head of every tile)
  int b_nxt = p.frame_utt[(size_t)(t_hi - 1) * NB], v_nxt = p.frame_utt[(size_t)(t_hi - 1) * NB + n];
  for (int t = t_hi - 1; t >= t_lo; --t, par ^= 1) {
    const int b = b_nxt;                                    // tiles never straddle utterances (32-frame alignment)
    const bool valid = v_nxt >= 0;
    if (t > t_lo) { b_nxt = p.frame_utt[(size_t)(t - 1) * NB]; v_nxt = p.frame_utt[(size_t)(t - 1) * NB + n]; }
    mbar_wait(mfull, par);
    if (b < 0) {                                            // tile entirely in padding: drain its copies, move on
      for (int j = 0; j < NCH; ++j) mbar_wait(full + j, par);
      if (tid == 0) p.cost_part[t] = 0.f;
      __syncwarp();
      if (lane == 0) { mbar_arrive(mempty); for (int j = 0; j < NCH; ++j) mbar_arrive(empty + j); }
      continue;
    }
    const float* mrec = meta + n * MS;                      // record of this thread's frame: H_old | g | multiplicities
    const float gg = mrec[K];
    float hk[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) hk[k] = (k < K) ? mrec[k] : 0.f;
    // multiplicities of the slots of this column: registers when R is a compile-time constant
    float wr[RT > 0 ? RT : 1];
    if (RT > 0) {
#pragma unroll
      for (int r = 0; r < RT; ++r) wr[r] = mrec[K + 1 + r];
    }
    const float* wsm = mrec + K + 1;
    auto wgt = [&](int r) -> float { return RT > 0 ? wr[RT > 0 ? r : 0] : wsm[r]; };

    if (b != cur_b) {                                       // new utterance: dictionary -> smem, column norms
      cur_b = b;
      cur_fo = p.frame_off[b];
      bar_compute();                                        // everybody is done with the previous W_s
      const float* wsrc = p.Wun + (size_t)b * F * K;
      for (int i = tid; i < F * K; i += CCT) { const int ff = i / K; W_s[ff * KS + (i - ff * K)] = __ldg(wsrc + i); }
      if (KS > K) for (int i = tid; i < F * (KS - K); i += CCT) { const int ff = i / (KS - K); W_s[ff * KS + K + (i - ff * (KS - K))] = 0.f; }
      bar_compute();
      for (int k = warp; k < K; k += CCT / 32) {            // c_k = sum_f |W_fk|  (mcem.py:128)
        float s = 0.f;
        for (int ff = lane; ff < F; ff += 32) s += fabsf(W_s[ff * KS + k]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) cn_s[k] = s;
      }
      bar_compute();
    }
    if (t * NB == cur_fo) {                                 // the first tile of an utterance writes W / c (mcem.py:131)
      float* wdst = p.W + (size_t)b * F * K;
      for (int i = tid; i < F * K; i += CCT) { const int ff = i / K, k = i - ff * K; wdst[i] = W_s[ff * KS + k] / cn_s[k]; }
    }

    // ---------------- pass A: H update with Vb = Wun @ H_old (mcem.py:113-121), chasing the copies
    float num[KMAX], den[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) num[k] = den[k] = 0.f;
    for (int j = 0; j < NCH; ++j) {
      mbar_wait(full + j, par);
      const int i0 = cb[j], i1 = min(cb[j + 1], my_ni);
      const float* vs = data + (size_t)(i0 * CFL + fl) * NB + n;
      const float* wrow = W_s + (i0 * CFL + fl) * KS;
      for (int i = i0; i < i1; ++i, vs += CFL * NB, wrow += CFL * KS) {
        {
          float w[KMAX];
#pragma unroll
          for (int k4 = 0; k4 < KMAX; k4 += 4) {
            const float4 t4 = *reinterpret_cast<const float4*>(wrow + k4);
            w[k4] = t4.x; w[k4 + 1] = t4.y; w[k4 + 2] = t4.z; w[k4 + 3] = t4.w;
          }
          f2 vbp = F2(0.f, 0.f);
#pragma unroll
          for (int k = 0; k < KMAX; k += 2) vbp = fma2(F2(w[k], w[k + 1]), F2(hk[k], hk[k + 1]), vbp);
          const float vb = vbp.x + vbp.y;
          float s1 = 0.f, s2 = 0.f;
          if (RT > 0) {
            f2 s1v = F2(0.f, 0.f), s2v = F2(0.f, 0.f);
            const f2 g2 = F2(gg, gg), vb2 = F2(vb, vb);
#pragma unroll
            for (int r = 0; r + 3 < RT; r += 4)
              quad_acc(g2, vb2, F2(vs[r * PS], vs[(r + 1) * PS]), F2(vs[(r + 2) * PS], vs[(r + 3) * PS]), F2(wgt(r), wgt(r + 1)),
                       F2(wgt(r + 2), wgt(r + 3)), s1v, s2v);
            s1 = s1v.x + s1v.y;
            s2 = s2v.x + s2v.y;
            constexpr int R4 = RT / 4 * 4;
            if (RT - R4 >= 2) pair_acc(gg, vb, vs[R4 * PS], vs[(R4 + 1) * PS], wgt(R4), wgt(R4 + 1), s1, s2);
            if (RT & 1) single_acc(gg, vb, vs[(RT - 1) * PS], wgt(RT - 1), s1, s2);
          } else {
            int r = 0;
            for (; r + 1 < R; r += 2) pair_acc(gg, vb, vs[r * PS], vs[(r + 1) * PS], wgt(r), wgt(r + 1), s1, s2);
            if (r < R) single_acc(gg, vb, vs[r * PS], wgt(r), s1, s2);
          }
          const float a = vs[R * PS] * s2;
          const f2 a2 = F2(a, a), d2 = F2(s1, s1);
#pragma unroll
          for (int k = 0; k < KMAX; k += 2) {
            const f2 w2 = F2(w[k], w[k + 1]);
            const f2 nn = fma2(w2, a2, F2(num[k], num[k + 1])), dd = fma2(w2, d2, F2(den[k], den[k + 1]));
            num[k] = nn.x; num[k + 1] = nn.y; den[k] = dd.x; den[k + 1] = dd.y;
          }
        }
      }
    }
    // reduce over the frequency lanes: lanes (8,16) in the warp, then the warps
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
    bar_compute();
    if (tid < KMAX * NB) {                                  // thread (k, n): both sums over the warps, then the update (one barrier less)
      const int k = tid >> 3;
      float nu = 0.f, de = 0.f;
#pragma unroll
      for (int w8 = 0; w8 < CW; ++w8) { nu += red[w8 * 2 * KMAX * NB + tid]; de += red[w8 * 2 * KMAX * NB + KMAX * NB + tid]; }
      Hn_s[tid] = (k < K) ? meta[(tid & (NB - 1)) * MS + k] * sqrtf(nu / de) : 0.f;
    }
    bar_compute();
    float hn[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) hn[k] = Hn_s[k * NB + n];
    if (RT > 0) {                                           // column data is in registers: release it to the producer
      __syncwarp();
      if (lane == 0) mbar_arrive(mempty);
    }

    // ---------------- pass B: Vb = Wun @ H_new (kept for the next E-step, mcem.py:124); g update (:138-142)
    float ng = 0.f, dg = 0.f;
    {
      const float* vs = data + (size_t)fl * NB + n;
      const float* wrow = W_s + fl * KS;
      const size_t vstep = (size_t)CFL * NP;
      float* vbo = p.Vb + (size_t)fl * NP + (size_t)t * NB + n;
      uint32_t* xvo = p.XV != nullptr ? p.XV + (size_t)fl * NP + (size_t)t * NB + n : nullptr;
      for (int i = 0; i < my_ni; ++i, vs += CFL * NB, wrow += CFL * KS, vbo += vstep, xvo += vstep) {
        {
          f2 vbp = F2(0.f, 0.f);
#pragma unroll
          for (int k4 = 0; k4 < KMAX; k4 += 4) {
            const float4 t4 = *reinterpret_cast<const float4*>(wrow + k4);
            vbp = fma2(F2(t4.x, t4.y), F2(hn[k4], hn[k4 + 1]), fma2(F2(t4.z, t4.w), F2(hn[k4 + 2], hn[k4 + 3]), vbp));
          }
          const float vb = vbp.x + vbp.y;
          if (valid) { *vbo = vb; if (p.XV != nullptr) *xvo = pack_xv_word(vs[R * PS], vb); }
          float t1 = 0.f, t2 = 0.f;
          f2 t1v = F2(0.f, 0.f), t2v = F2(0.f, 0.f);
          const f2 g2 = F2(gg, gg), vb2 = F2(vb, vb);
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
            for (int r = 0; r + 3 < RT; r += 4)
              quad_accB(g2, vb2, F2(vs[r * PS], vs[(r + 1) * PS]), F2(vs[(r + 2) * PS], vs[(r + 3) * PS]), F2(wgt(r), wgt(r + 1)),
                        F2(wgt(r + 2), wgt(r + 3)), t1v, t2v);
            t1 = t1v.x + t1v.y;
            t2 = t2v.x + t2v.y;
            constexpr int R4 = RT / 4 * 4;
            if (RT - R4 >= 2) pairB(vs[R4 * PS], vs[(R4 + 1) * PS], wgt(R4), wgt(R4 + 1));
            if (RT & 1) singleB(vs[(RT - 1) * PS], wgt(RT - 1));
          } else {
            int r = 0;
            for (; r + 1 < R; r += 2) pairB(vs[r * PS], vs[(r + 1) * PS], wgt(r), wgt(r + 1));
            if (r < R) singleB(vs[r * PS], wgt(r));
          }
          ng = fmaf(vs[R * PS], t2, ng);
          dg += t1;
        }
      }
    }
    ng += __shfl_xor_sync(0xffffffffu, ng, 8);
    ng += __shfl_xor_sync(0xffffffffu, ng, 16);
    dg += __shfl_xor_sync(0xffffffffu, dg, 8);
    dg += __shfl_xor_sync(0xffffffffu, dg, 16);
    if (lane < NB) { red[(warp * 2) * NB + lane] = ng; red[(warp * 2 + 1) * NB + lane] = dg; }
    bar_compute();
    float sn = 0.f, sd = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < CW; ++w8) { sn += red[(w8 * 2) * NB + n]; sd += red[(w8 * 2 + 1) * NB + n]; }
    const float gnew = gg * sqrtf(sn / sd);
    // outputs of the tile
    if (tid < NB && valid) p.g[(size_t)t * NB + n] = gnew;
    if (tid < KMAX * NB) {
      const int k = tid >> 3;
      if (k < K && valid) p.H[(size_t)k * NP + (size_t)t * NB + n] = Hn_s[tid] * cn_s[k];   // mcem.py:133
    }

    // ---------------- pass C: cost with the new g (mcem.py:151-152, :68-70); chunks are released behind it
    float cl = 0.f, cr = 0.f;
    for (int j = 0; j < NCH; ++j) {
      const int i0 = cb[j], i1 = min(cb[j + 1], my_ni);
      const float* vs = data + (size_t)(i0 * CFL + fl) * NB + n;
      const float* wrow = W_s + (i0 * CFL + fl) * KS;
      for (int i = i0; i < i1; ++i, vs += CFL * NB, wrow += CFL * KS) {
        {
          f2 vbp = F2(0.f, 0.f);
#pragma unroll
          for (int k4 = 0; k4 < KMAX; k4 += 4) {
            const float4 t4 = *reinterpret_cast<const float4*>(wrow + k4);
            vbp = fma2(F2(t4.x, t4.y), F2(hn[k4], hn[k4 + 1]), fma2(F2(t4.z, t4.w), F2(hn[k4 + 2], hn[k4 + 3]), vbp));
          }
          const float vb = vbp.x + vbp.y;
          float sl = 0.f, sr = 0.f;
          f2 slv = F2(0.f, 0.f), srv = F2(0.f, 0.f);
          const f2 g2 = F2(gnew, gnew), vb2 = F2(vb, vb);
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
            for (int r = 0; r + 3 < RT; r += 4)
              quad_accC(g2, vb2, F2(vs[r * PS], vs[(r + 1) * PS]), F2(vs[(r + 2) * PS], vs[(r + 3) * PS]), F2(wgt(r), wgt(r + 1)),
                        F2(wgt(r + 2), wgt(r + 3)), slv, srv);
            sl = slv.x + slv.y;
            sr = srv.x + srv.y;
            constexpr int R4 = RT / 4 * 4;
            if (RT - R4 >= 2) pairC(vs[R4 * PS], vs[(R4 + 1) * PS], wgt(R4), wgt(R4 + 1));
            if (RT & 1) singleC(vs[(RT - 1) * PS], wgt(RT - 1));
          } else {
            int r = 0;
            for (; r + 1 < R; r += 2) pairC(vs[r * PS], vs[(r + 1) * PS], wgt(r), wgt(r + 1));
            if (r < R) singleC(vs[r * PS], wgt(r));
          }
          cl += sl;
          cr = fmaf(vs[R * PS], sr, cr);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + j);                // this warp is done with chunk j of this tile
    }
    if (RT == 0) {                                          // runtime-R path kept reading the column data until here
      __syncwarp();
      if (lane == 0) mbar_arrive(mempty);
    }
    float cs = valid ? fmaf(cl, 0.6931471805599453f, cr) : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cs += __shfl_xor_sync(0xffffffffu, cs, o);
    if (lane == 0) misc[warp] = cs;
    bar_compute();
    if (tid == 0) {
      float s = 0.f;
      for (int w8 = 0; w8 < CW; ++w8) s += misc[w8];
      p.cost_part[t] = s;
    }
  }
}

// ------------------------------------------- generic column sweep (any K <= 32, any R)
// Same three passes as k_cols_v1 for the shapes whose (R+1) x F x 8 tile does not fit in shared memory
// next to the dictionary (K > 12, or the R = 30 chains of MCEM_M1).  The sample slots are read straight
// from HBM in the H pass -- a warp covers 4 bins x 8 frames = 128 contiguous bytes of the column-tile
// layout -- and from L2 in the g and cost passes (the tiles in flight on the chip stay far below its
// capacity).  The rank-K part is taken out of the element loop: the H pass leaves a = X2 * s2 and s1 per
// (f, n) in shared memory and a (k, n)-threaded product with the dictionary finishes the update, so no
// thread carries K-sized arrays besides its column of H.  ~110 KB of shared memory: two CTAs per SM.
struct GenArgs {
  int F, K, KS, NP, R, ntiles;
  const int32_t* frame_utt; const int32_t* frame_off;
  const float* X2t; const float* Vs; const float* Mt;
  float* Vb; float* g; float* H; const float* Wun; float* W; float* cost_part; uint32_t* XV;
};
constexpr int GT = 256;
__host__ __device__ inline size_t gen_smem_floats(int F, int KS, int K, int R) {
  return (size_t)F * KS + (size_t)2 * F * NB + (size_t)meta_stride(K, R) * NB + (size_t)K * NB + 32 + 8 * 2 * NB + 32 + 16;
}

template <int KMAX>
__global__ void __launch_bounds__(GT, 2) k_cols_gen(GenArgs p) {
  extern __shared__ __align__(16) float sg[];
  const int F = p.F, K = p.K, KS = p.KS, NP = p.NP, R = p.R, MS = meta_stride(K, R), T8 = NP / NB;
  float* W_s = sg;                                          // [F][KS]
  float* a_s = W_s + (size_t)F * KS;                        // [F][NB]  X2 * s2 of the H pass
  float* s1_s = a_s + (size_t)F * NB;                       // [F][NB]
  float* meta = s1_s + (size_t)F * NB;                      // H_old [K][NB] | g [NB] | multiplicities [R][NB]
  float* Hn_s = meta + (size_t)MS * NB;                     // [K][NB]
  float* cn_s = Hn_s + (size_t)K * NB;                      // [32]
  float* red = cn_s + 32;                                   // [8 warps][2][NB]
  float* misc = red + 8 * 2 * NB;                           // [32]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = tid & (NB - 1), fl = tid >> 3;              // 32 frequency lanes
  const size_t slab = (size_t)T8 * F * NB;                  // one sample slot
  int cur_b = -1, cur_fo = -1;
  for (int t = p.ntiles - 1 - (int)blockIdx.x; t >= 0; t -= gridDim.x) {
    const int b = p.frame_utt[(size_t)t * NB];
    if (b < 0) { if (tid == 0) p.cost_part[t] = 0.f; continue; }
    __syncthreads();                                        // previous tile is done with the shared arrays
    for (int i = tid; i < MS * NB; i += GT) meta[i] = __ldg(p.Mt + (size_t)t * MS * NB + i);
    if (b != cur_b) {
      cur_b = b;
      cur_fo = p.frame_off[b];
      const float* wsrc = p.Wun + (size_t)b * F * K;
      for (int i = tid; i < F * K; i += GT) { const int ff = i / K; W_s[ff * KS + (i - ff * K)] = __ldg(wsrc + i); }
      if (KS > K) for (int i = tid; i < F * (KS - K); i += GT) { const int ff = i / (KS - K); W_s[ff * KS + K + (i - ff * (KS - K))] = 0.f; }
      __syncthreads();
      for (int k = warp; k < K; k += GT / 32) {             // c_k = sum_f |W_fk|  (mcem.py:128)
        float sum = 0.f;
        for (int ff = lane; ff < F; ff += 32) sum += fabsf(W_s[ff * KS + k]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (lane == 0) cn_s[k] = sum;
      }
    }
    __syncthreads();
    if (t * NB == cur_fo) {                                 // the first tile of an utterance writes W / c (mcem.py:131)
      float* wdst = p.W + (size_t)b * F * K;
      for (int i = tid; i < F * K; i += GT) { const int ff = i / K, k = i - ff * K; wdst[i] = W_s[ff * KS + k] / cn_s[k]; }
    }
    const bool valid = p.frame_utt[(size_t)t * NB + n] >= 0;
    const float* mrec = meta + n * MS;                      // record of this thread's frame
    const float gg = mrec[K];
    const float* wts = mrec + K + 1;
    const float* vs0 = p.Vs + (size_t)t * F * NB + n;       // + f*NB + r*slab
    const float* x2p = p.X2t + (size_t)t * F * NB + n;
    float hk[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) hk[k] = (k < K) ? mrec[k] : 0.f;
    auto dot_w = [&](const float* wrow, const float (&h)[KMAX]) -> float {
      float vb0 = 0.f, vb1 = 0.f;
#pragma unroll
      for (int k4 = 0; k4 < KMAX; k4 += 4) {
        const float4 t4 = *reinterpret_cast<const float4*>(wrow + k4);
        vb0 = fmaf(t4.x, h[k4], fmaf(t4.z, h[k4 + 2], vb0));
        vb1 = fmaf(t4.y, h[k4 + 1], fmaf(t4.w, h[k4 + 3], vb1));
      }
      return vb0 + vb1;
    };

    // ---------------- pass A: s1, s2 with Vb = Wun @ H_old (mcem.py:113-121)
    for (int f = fl; f < F; f += 32) {
      const float vb = dot_w(W_s + f * KS, hk);
      const float* vs = vs0 + (size_t)f * NB;
      float s1 = 0.f, s2 = 0.f;
      int r = 0;
      for (; r + 1 < R; r += 2) pair_acc(gg, vb, __ldg(vs + r * slab), __ldg(vs + (r + 1) * slab), wts[r], wts[r + 1], s1, s2);
      if (r < R) single_acc(gg, vb, __ldg(vs + r * slab), wts[r], s1, s2);
      a_s[f * NB + n] = __ldg(x2p + (size_t)f * NB) * s2;
      s1_s[f * NB + n] = s1;
    }
    __syncthreads();
    // H update  H <- H * sqrt( (W^T (X2 s2)) / (W^T s1) ): thread = (slice of the frequency rows, group of four dictionary
    // columns, frame) with eight accumulators -- one 16-byte load of W and two scalar loads per eight multiply-adds --
    // then the slices are summed through the (now free) a_s array
    {
      const int KQ = (K + 3) / 4, combos = KQ * NB, nfs = GT / combos;
      const int fs = tid / combos, rem = tid - fs * combos, kq = rem / NB, nn = rem & (NB - 1);
      float nu[4] = {0.f, 0.f, 0.f, 0.f}, de[4] = {0.f, 0.f, 0.f, 0.f};
      if (fs < nfs) {
        for (int f = fs; f < F; f += nfs) {
          const float4 w4 = *reinterpret_cast<const float4*>(W_s + f * KS + 4 * kq);     // columns beyond K are zero (KS >= KMAX)
          const float av = a_s[f * NB + nn], sv = s1_s[f * NB + nn];
          nu[0] = fmaf(w4.x, av, nu[0]); de[0] = fmaf(w4.x, sv, de[0]);
          nu[1] = fmaf(w4.y, av, nu[1]); de[1] = fmaf(w4.y, sv, de[1]);
          nu[2] = fmaf(w4.z, av, nu[2]); de[2] = fmaf(w4.z, sv, de[2]);
          nu[3] = fmaf(w4.w, av, nu[3]); de[3] = fmaf(w4.w, sv, de[3]);
        }
      }
      __syncthreads();                                      // everybody has read a_s / s1_s
      float* part = a_s;                                    // [nfs][4 KQ][NB][2]
      if (fs < nfs) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float* d = part + (((size_t)fs * 4 * KQ + 4 * kq + i) * NB + nn) * 2;
          d[0] = nu[i]; d[1] = de[i];
        }
      }
      __syncthreads();
      if (tid < K * NB) {
        const int k = tid / NB;
        float sn = 0.f, sd_ = 0.f;
        for (int q = 0; q < nfs; ++q) { const float* d = part + (((size_t)q * 4 * KQ + k) * NB + n) * 2; sn += d[0]; sd_ += d[1]; }
        Hn_s[tid] = meta[n * MS + k] * sqrtf(sn / sd_);
      }
    }
    __syncthreads();
    float hn[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) hn[k] = (k < K) ? Hn_s[k * NB + n] : 0.f;

    // ---------------- pass B: Vb = Wun @ H_new (mcem.py:124), g update (:138-142)
    float ng = 0.f, dg = 0.f;
    for (int f = fl; f < F; f += 32) {
      const float vb = dot_w(W_s + f * KS, hn);
      if (valid) {
        p.Vb[(size_t)f * NP + (size_t)t * NB + n] = vb;
        if (p.XV != nullptr) p.XV[(size_t)f * NP + (size_t)t * NB + n] = pack_xv_word(__ldg(x2p + (size_t)f * NB), vb);
      }
      a_s[f * NB + n] = vb;                                 // kept for the cost pass
      const float* vs = vs0 + (size_t)f * NB;
      float t1 = 0.f, t2 = 0.f;
      int r = 0;
      for (; r + 1 < R; r += 2) {
        const float va = __ldg(vs + r * slab), vc = __ldg(vs + (r + 1) * slab);
        const float a = fmaf(gg, va, vb), c = fmaf(gg, vc, vb);
        const float ip = rcp_fast(a * c);
        const float ia = c * ip, ic = a * ip;
        const float ua = wts[r] * va * ia, uc = wts[r + 1] * vc * ic;
        t1 += ua + uc;
        t2 = fmaf(ua, ia, fmaf(uc, ic, t2));
      }
      if (r < R) {
        const float va = __ldg(vs + r * slab);
        const float ia = rcp_fast(fmaf(gg, va, vb)), ua = wts[r] * va * ia;
        t1 += ua;
        t2 = fmaf(ua, ia, t2);
      }
      ng = fmaf(__ldg(x2p + (size_t)f * NB), t2, ng);
      dg += t1;
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
    if (tid < NB && valid) p.g[(size_t)t * NB + n] = gnew;
    if (tid < K * NB && valid) p.H[(size_t)(tid >> 3) * NP + (size_t)t * NB + n] = Hn_s[tid] * cn_s[tid >> 3];   // mcem.py:133

    // ---------------- pass C: cost with the new g (mcem.py:151-152, :68-70)
    float cl = 0.f, cr = 0.f;
    for (int f = fl; f < F; f += 32) {
      const float vb = a_s[f * NB + n];
      const float* vs = vs0 + (size_t)f * NB;
      float sl = 0.f, sr = 0.f;
      int r = 0;
      for (; r + 1 < R; r += 2) {
        const float a = fmaf(gnew, __ldg(vs + r * slab), vb), c = fmaf(gnew, __ldg(vs + (r + 1) * slab), vb);
        const float wa = wts[r], wc = wts[r + 1];
        sl = fmaf(wa, lg2_fast(a), fmaf(wc, lg2_fast(c), sl));
        sr = fmaf(fmaf(wa, c, wc * a), rcp_fast(a * c), sr);
      }
      if (r < R) {
        const float a = fmaf(gnew, __ldg(vs + r * slab), vb), wa = wts[r];
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
      p.cost_part[t] = sum;
    }
  }
}

// ------------------------------------------- streamed column sweep (any K <= 32, any R)
// The generic sweep above loads every sample slot with scalar loads, three times per tile, and is bound by the latency
// of those loads (C4: 0.58 ms per iteration, 9 % of the HBM peak).  This one moves the same bytes the way the W sweep
// does: a producer warp streams the tile -- (R+1) planes, chunk after chunk of SFL frequency rows -- through a ring of
// shared-memory stages with bulk async copies, ONCE PER PASS (the first pass from HBM, the other two from L2, where the
// 180 KB a tile occupies are still resident), and the compute warps never touch global memory for the slots.  One row
// per thread and chunk; the arithmetic, the reductions and the outputs are those of k_cols_gen.
constexpr int SW_ = 12;                // compute warps (+ 1 producer warp: 13 warps, 128 registers)
constexpr int SCT = SW_ * 32;
constexpr int SFL = SCT / NB;          // 48 frequency rows per chunk
constexpr int STT = SCT + 32;
constexpr int SMAXST = 8;
__host__ __device__ inline int stream_stage_floats(int R) { return (R + 1) * SFL * NB; }
// a_s | s1_s ([F][NB] each) double as the scratch of the H product: SCT threads x 16 partial sums
__host__ __device__ inline size_t stream_as_floats(int F) { const size_t a = (size_t)2 * F * NB, b = (size_t)SCT * 16; return a > b ? a : b; }
__host__ __device__ inline size_t stream_fixed_floats(int F, int KS, int K, int R) {
  return (size_t)F * KS + stream_as_floats(F) + (size_t)meta_stride(K, R) * NB + (size_t)K * NB + 32 + SW_ * 2 * NB + 32 + 2 * (2 * SMAXST + 2) + 16;
}
__device__ __forceinline__ void bar_stream() { asm volatile("bar.sync 1, %0;" ::"n"(SCT) : "memory"); }

// Tensor maps of the sample slots and of X2 in column-tile order: Vs = [R][NP/8][F][8] f32 as a 4-D tensor (8, F, NP/8, R)
// with box (8, SFL rows, 1 tile, R slots) -- ONE instruction fetches a chunk of all R planes (eleven separate bulk
// copies of 1.5 KB cost the producer ~1 k cycles per chunk to issue) -- and X2t = [NP/8][F][8] as (8, F, NP/8) with box
// (8, SFL, 1).  Rows beyond F read as zeros.
struct StreamMaps { CUtensorMap vs, x2; };

// RT > 0: the number of sample slots is a compile-time constant (10 in the M2 scripts and in BASELINE config 4): the slot
// loops unroll, their offsets become immediates and the multiplicities of the thread's frame live in registers.
template <int KMAX, int RT>
__global__ void __launch_bounds__(STT, 1) k_cols_stream(const __grid_constant__ StreamMaps maps, GenArgs p, int nstage) {
  extern __shared__ __align__(128) float ss[];
  const int F = p.F, K = p.K, KS = p.KS, NP = p.NP, R = RT > 0 ? RT : p.R, MS = meta_stride(K, R), T8 = NP / NB;
  const int SSt = stream_stage_floats(R);
  constexpr int PSt = SFL * NB;   // stage, plane inside a stage (floats)
  float* ring = ss;                                         // [nstage][(R+1)][SFL][NB]
  float* W_s = ring + (size_t)nstage * SSt;                 // [F][KS]
  float* a_s = W_s + (size_t)F * KS;                        // [F][NB]  X2 * s2 of the H pass, then Vb of the g pass
  float* s1_s = a_s + (size_t)F * NB;                       // [F][NB]
  float* meta = a_s + stream_as_floats(F);                  // records of the tile's 8 frames: H_old | g | multiplicities
  float* Hn_s = meta + (size_t)MS * NB;                     // [K][NB]
  float* cn_s = Hn_s + (size_t)K * NB;                      // [32]
  float* red = cn_s + 32;                                   // [SW_][2][NB]
  float* misc = red + SW_ * 2 * NB;                         // [32]
  uint64_t* bars = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(misc + 32) + 7) & ~(uintptr_t)7);
  uint64_t* full = bars;                                    // [nstage]
  uint64_t* empty = bars + SMAXST;                          // [nstage]
  uint64_t* mfull = bars + 2 * SMAXST;                      // column data of the tile landed
  uint64_t* mempty = mfull + 1;                             // ... and may be overwritten
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int NCHK = (F + SFL - 1) / SFL;                     // chunks per pass
  if (tid == 0) {
    for (int s_ = 0; s_ < nstage; ++s_) { mbar_init(full + s_, 1); mbar_init(empty + s_, SW_); }
    mbar_init(mfull, 1);
    mbar_init(mempty, SW_);
    fence_mbar_init();
  }
  __syncthreads();
  const int t_first = p.ntiles - 1 - (int)blockIdx.x;

  if (tid >= SCT) {
    // ================= producer warp: lane = plane; one bulk copy per plane and chunk, three passes per tile =================
    uint32_t q = 0, tiles = 0;                              // chunks issued so far, tiles started
    for (int t = t_first; t >= 0; t -= gridDim.x) {
      if (p.frame_utt[(size_t)t * NB] < 0) continue;        // tile entirely in padding: the compute warps skip it too
      if (tiles > 0) mbar_wait(mempty, (tiles - 1) & 1);
      if (lane == 0) {
        mbar_expect_tx(mfull, (uint32_t)MS * NB * 4);
        bulk_g2s(meta, p.Mt + (size_t)t * MS * NB, (uint32_t)MS * NB * 4, mfull);
      }
      ++tiles;
      for (int pass = 0; pass < 3; ++pass) {
        for (int c = 0; c < NCHK; ++c, ++q) {
          const int st_ = q % nstage;
          if (q >= (uint32_t)nstage) mbar_wait(empty + st_, ((q / nstage) - 1) & 1);
          float* dst = ring + (size_t)st_ * SSt;
          if (lane == 0) {
            mbar_expect_tx(full + st_, (uint32_t)SSt * 4);                // whole boxes: rows beyond F arrive as zeros
            tc::tma_load_4d(dst, &maps.vs, 0, c * SFL, t, 0, full + st_);
            tc::tma_load_3d(dst + (size_t)R * PSt, &maps.x2, 0, c * SFL, t, full + st_);
          }
        }
      }
    }
    return;
  }

  // ================= compute warps =================
  const int n = tid & (NB - 1), fl = tid >> 3;              // column of the tile, frequency lane (row of a chunk)
  int cur_b = -1, cur_fo = -1;
  uint32_t q = 0, tiles = 0;
  for (int t = t_first; t >= 0; t -= gridDim.x) {
    const int b = p.frame_utt[(size_t)t * NB];
    if (b < 0) { if (tid == 0) p.cost_part[t] = 0.f; continue; }
    bar_stream();                                           // previous tile is done with the shared arrays
    if (b != cur_b) {
      cur_b = b;
      cur_fo = p.frame_off[b];
      const float* wsrc = p.Wun + (size_t)b * F * K;
      for (int i = tid; i < F * K; i += SCT) { const int ff = i / K; W_s[ff * KS + (i - ff * K)] = __ldg(wsrc + i); }
      if (KS > K) for (int i = tid; i < F * (KS - K); i += SCT) { const int ff = i / (KS - K); W_s[ff * KS + K + (i - ff * (KS - K))] = 0.f; }
      bar_stream();
      for (int k = warp; k < K; k += SW_) {                 // c_k = sum_f |W_fk|  (mcem.py:128)
        float sum = 0.f;
        for (int ff = lane; ff < F; ff += 32) sum += fabsf(W_s[ff * KS + k]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (lane == 0) cn_s[k] = sum;
      }
    }
    bar_stream();
    if (t * NB == cur_fo) {                                 // the first tile of an utterance writes W / c (mcem.py:131)
      float* wdst = p.W + (size_t)b * F * K;
      for (int i = tid; i < F * K; i += SCT) { const int ff = i / K, k = i - ff * K; wdst[i] = W_s[ff * KS + k] / cn_s[k]; }
    }
    mbar_wait(mfull, tiles & 1);
    ++tiles;
    const bool valid = p.frame_utt[(size_t)t * NB + n] >= 0;
    const float* mrec = meta + n * MS;                      // record of this thread's frame
    const float gg = mrec[K];
    const float* wsm = mrec + K + 1;
    float wr[RT > 0 ? RT : 1];
    if (RT > 0) {
#pragma unroll
      for (int r = 0; r < RT; ++r) wr[RT > 0 ? r : 0] = wsm[r];
    }
    auto wgt = [&](int r) -> float { return RT > 0 ? wr[RT > 0 ? r : 0] : wsm[r]; };
    float hk[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) hk[k] = (k < K) ? mrec[k] : 0.f;
    auto dot_w = [&](const float* wrow, const float (&h)[KMAX]) -> float {
      f2 acc = F2(0.f, 0.f);
#pragma unroll
      for (int k4 = 0; k4 < KMAX; k4 += 4) {
        const float4 t4 = *reinterpret_cast<const float4*>(wrow + k4);
        acc = fma2(F2(t4.x, t4.y), F2(h[k4], h[k4 + 1]), fma2(F2(t4.z, t4.w), F2(h[k4 + 2], h[k4 + 3]), acc));
      }
      return acc.x + acc.y;
    };
    // next chunk of the stream: waits for its stage; returns the stage's base and this thread's row (or -1 past the spectrum)
    auto next_chunk = [&](int c, int& f) -> const float* {
      const int st_ = q % nstage;
      mbar_wait(full + st_, (q / nstage) & 1);
      f = c * SFL + fl;
      if (f >= F) f = -1;
      return ring + (size_t)st_ * SSt + fl * NB + n;
    };
    auto done_chunk = [&]() {
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + (q % nstage));
      ++q;
    };

    // ---------------- pass A: s1, s2 with Vb = Wun @ H_old (mcem.py:113-121)
    for (int c = 0; c < NCHK; ++c) {
      int f;
      const float* vs = next_chunk(c, f);
      if (f >= 0) {
        const float vb = dot_w(W_s + f * KS, hk);
        float s1 = 0.f, s2 = 0.f;
        if (RT > 0) {                                       // four slots at a time in packed f32x2 arithmetic (as k_cols_v1)
          f2 s1v = F2(0.f, 0.f), s2v = F2(0.f, 0.f);
          const f2 g2 = F2(gg, gg), vb2 = F2(vb, vb);
#pragma unroll
          for (int r = 0; r + 3 < RT; r += 4)
            quad_acc(g2, vb2, F2(vs[r * PSt], vs[(r + 1) * PSt]), F2(vs[(r + 2) * PSt], vs[(r + 3) * PSt]), F2(wgt(r), wgt(r + 1)),
                     F2(wgt(r + 2), wgt(r + 3)), s1v, s2v);
          s1 = s1v.x + s1v.y;
          s2 = s2v.x + s2v.y;
          constexpr int R4 = RT / 4 * 4;
          if (RT - R4 >= 2) pair_acc(gg, vb, vs[R4 * PSt], vs[(R4 + 1) * PSt], wgt(R4), wgt(R4 + 1), s1, s2);
          if (RT & 1) single_acc(gg, vb, vs[(RT - 1) * PSt], wgt(RT - 1), s1, s2);
        } else {
          int r = 0;
          for (; r + 1 < R; r += 2) pair_acc(gg, vb, vs[r * PSt], vs[(r + 1) * PSt], wgt(r), wgt(r + 1), s1, s2);
          if (r < R) single_acc(gg, vb, vs[r * PSt], wgt(r), s1, s2);
        }
        a_s[f * NB + n] = vs[R * PSt] * s2;
        s1_s[f * NB + n] = s1;
      }
      done_chunk();
    }
    bar_stream();
    // H update  H <- H * sqrt( (W^T (X2 s2)) / (W^T s1) ): thread = (slice of the frequency rows, group of four dictionary
    // columns, frame) with eight accumulators, then the slices are summed through the (now free) a_s array
    {
      // thread = (slice of the frequency rows, group of four dictionary columns, PAIR of frames): sixteen accumulators as
      // eight packed pairs per 16-byte load of W and two 8-byte loads of a / s1
      const int KQ = (K + 3) / 4, combos = KQ * (NB / 2), nfs = SCT / combos;
      const int fs = tid / combos, rem = tid - fs * combos, kq = rem / (NB / 2), np2 = (rem % (NB / 2)) * 2;
      f2 nu[4], de[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) nu[i] = de[i] = F2(0.f, 0.f);
      if (fs < nfs) {
        const float* wp = W_s + (size_t)fs * KS + 4 * kq;
        const float* ap = a_s + fs * NB + np2;
        const float* sp = s1_s + fs * NB + np2;
        for (int f = fs; f < F; f += nfs, wp += (size_t)nfs * KS, ap += nfs * NB, sp += nfs * NB) {
          const float4 w4 = *reinterpret_cast<const float4*>(wp);                          // columns beyond K are zero (KS >= KMAX)
          const f2 av = *reinterpret_cast<const float2*>(ap), sv = *reinterpret_cast<const float2*>(sp);
          nu[0] = fma2(F2(w4.x, w4.x), av, nu[0]); de[0] = fma2(F2(w4.x, w4.x), sv, de[0]);
          nu[1] = fma2(F2(w4.y, w4.y), av, nu[1]); de[1] = fma2(F2(w4.y, w4.y), sv, de[1]);
          nu[2] = fma2(F2(w4.z, w4.z), av, nu[2]); de[2] = fma2(F2(w4.z, w4.z), sv, de[2]);
          nu[3] = fma2(F2(w4.w, w4.w), av, nu[3]); de[3] = fma2(F2(w4.w, w4.w), sv, de[3]);
        }
      }
      bar_stream();                                         // everybody has read a_s / s1_s
      float* part = a_s;                                    // [nfs][4 KQ][NB][2]
      if (fs < nfs) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float* d = part + (((size_t)fs * 4 * KQ + 4 * kq + i) * NB + np2) * 2;
          *reinterpret_cast<float4*>(d) = make_float4(nu[i].x, de[i].x, nu[i].y, de[i].y);
        }
      }
      bar_stream();
      if (tid < K * NB) {
        const int k = tid / NB;
        float sn = 0.f, sd_ = 0.f;
        for (int qq = 0; qq < nfs; ++qq) { const float* d = part + (((size_t)qq * 4 * KQ + k) * NB + n) * 2; sn += d[0]; sd_ += d[1]; }
        Hn_s[tid] = meta[n * MS + k] * sqrtf(sn / sd_);
      }
    }
    bar_stream();
#pragma unroll
    for (int k = 0; k < KMAX; ++k) hk[k] = (k < K) ? Hn_s[k * NB + n] : 0.f;      // hk is H_new from here on

    // ---------------- pass B: Vb = Wun @ H_new (mcem.py:124), g update (:138-142)
    float ng = 0.f, dg = 0.f;
    for (int c = 0; c < NCHK; ++c) {
      int f;
      const float* vs = next_chunk(c, f);
      if (f >= 0) {
        const float vb = dot_w(W_s + f * KS, hk);
        const float x2 = vs[R * PSt];
        if (valid) {
          p.Vb[(size_t)f * NP + (size_t)t * NB + n] = vb;
          if (p.XV != nullptr) p.XV[(size_t)f * NP + (size_t)t * NB + n] = pack_xv_word(x2, vb);
        }
        a_s[f * NB + n] = vb;                               // kept for the cost pass
        float t1 = 0.f, t2 = 0.f;
        auto pairB = [&](int r) {
          const float va = vs[r * PSt], vc = vs[(r + 1) * PSt];
          const float a = fmaf(gg, va, vb), cc = fmaf(gg, vc, vb);
          const float ip = rcp_fast(a * cc);
          const float ia = cc * ip, ic = a * ip;
          const float ua = wgt(r) * va * ia, uc = wgt(r + 1) * vc * ic;
          t1 += ua + uc;
          t2 = fmaf(ua, ia, fmaf(uc, ic, t2));
        };
        auto singleB = [&](int r) {
          const float va = vs[r * PSt];
          const float ia = rcp_fast(fmaf(gg, va, vb)), ua = wgt(r) * va * ia;
          t1 += ua;
          t2 = fmaf(ua, ia, t2);
        };
        if (RT > 0) {
          f2 t1v = F2(0.f, 0.f), t2v = F2(0.f, 0.f);
          const f2 g2 = F2(gg, gg), vb2 = F2(vb, vb);
#pragma unroll
          for (int r = 0; r + 3 < RT; r += 4)
            quad_accB(g2, vb2, F2(vs[r * PSt], vs[(r + 1) * PSt]), F2(vs[(r + 2) * PSt], vs[(r + 3) * PSt]), F2(wgt(r), wgt(r + 1)),
                      F2(wgt(r + 2), wgt(r + 3)), t1v, t2v);
          t1 = t1v.x + t1v.y;
          t2 = t2v.x + t2v.y;
          constexpr int R4 = RT / 4 * 4;
          if (RT - R4 >= 2) pairB(R4);
          if (RT & 1) singleB(RT - 1);
        } else {
          int r = 0;
          for (; r + 1 < R; r += 2) pairB(r);
          if (r < R) singleB(r);
        }
        ng = fmaf(x2, t2, ng);
        dg += t1;
      }
      done_chunk();
    }
    ng += __shfl_xor_sync(0xffffffffu, ng, 8);
    ng += __shfl_xor_sync(0xffffffffu, ng, 16);
    dg += __shfl_xor_sync(0xffffffffu, dg, 8);
    dg += __shfl_xor_sync(0xffffffffu, dg, 16);
    if (lane < NB) { red[(warp * 2) * NB + lane] = ng; red[(warp * 2 + 1) * NB + lane] = dg; }
    bar_stream();
    float sn = 0.f, sd = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < SW_; ++w8) { sn += red[(w8 * 2) * NB + n]; sd += red[(w8 * 2 + 1) * NB + n]; }
    const float gnew = gg * sqrtf(sn / sd);
    if (tid < NB && valid) p.g[(size_t)t * NB + n] = gnew;
    if (tid < K * NB && valid) p.H[(size_t)(tid >> 3) * NP + (size_t)t * NB + n] = Hn_s[tid] * cn_s[tid >> 3];   // mcem.py:133

    // ---------------- pass C: cost with the new g (mcem.py:151-152, :68-70)
    float cl = 0.f, cr = 0.f;
    for (int c = 0; c < NCHK; ++c) {
      int f;
      const float* vs = next_chunk(c, f);
      if (f >= 0) {
        const float vb = a_s[f * NB + n];
        float sl = 0.f, sr = 0.f;
        auto pairC = [&](int r) {
          const float a = fmaf(gnew, vs[r * PSt], vb), cc = fmaf(gnew, vs[(r + 1) * PSt], vb);
          const float wa = wgt(r), wc = wgt(r + 1);
          sl = fmaf(wa, lg2_fast(a), fmaf(wc, lg2_fast(cc), sl));
          sr = fmaf(fmaf(wa, cc, wc * a), rcp_fast(a * cc), sr);
        };
        auto singleC = [&](int r) {
          const float a = fmaf(gnew, vs[r * PSt], vb), wa = wgt(r);
          sl = fmaf(wa, lg2_fast(a), sl);
          sr = fmaf(wa, rcp_fast(a), sr);
        };
        if (RT > 0) {
          f2 slv = F2(0.f, 0.f), srv = F2(0.f, 0.f);
          const f2 g2 = F2(gnew, gnew), vb2 = F2(vb, vb);
#pragma unroll
          for (int r = 0; r + 3 < RT; r += 4)
            quad_accC(g2, vb2, F2(vs[r * PSt], vs[(r + 1) * PSt]), F2(vs[(r + 2) * PSt], vs[(r + 3) * PSt]), F2(wgt(r), wgt(r + 1)),
                      F2(wgt(r + 2), wgt(r + 3)), slv, srv);
          sl = slv.x + slv.y;
          sr = srv.x + srv.y;
          constexpr int R4 = RT / 4 * 4;
          if (RT - R4 >= 2) pairC(R4);
          if (RT & 1) singleC(RT - 1);
        } else {
          int r = 0;
          for (; r + 1 < R; r += 2) pairC(r);
          if (r < R) singleC(r);
        }
        cl += sl;
        cr = fmaf(vs[R * PSt], sr, cr);
      }
      done_chunk();
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(mempty);                     // the column data of this tile (wts) is not read any more
    float cs = valid ? fmaf(cl, 0.6931471805599453f, cr) : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cs += __shfl_xor_sync(0xffffffffu, cs, o);
    if (lane == 0) misc[warp] = cs;
    bar_stream();
    if (tid == 0) {
      float sum = 0.f;
      for (int w8 = 0; w8 < SW_; ++w8) sum += misc[w8];
      p.cost_part[t] = sum;
    }
  }
}

template <int KMAX, int RT>
int32_t launch_cols_stream(const GenArgs& a, int nstage, size_t smem, int grid, cudaStream_t st) {
  static size_t smem_tab[GVN_MAX_DEVICES] = {0};
  size_t& smem_set = *per_device_slot(smem_tab);
  if (smem_set != smem) {
    cudaError_t e = cudaFuncSetAttribute(k_cols_stream<KMAX, RT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(GVN_E_CUDA, "k_cols_stream smem attr (%zu B): %s", smem, cudaGetErrorString(e));
    smem_set = smem;
  }
  static thread_local StreamMaps maps;                  // encoded once per (pointers, shape)
  static thread_local const void* m_vs = nullptr; static thread_local const void* m_x2 = nullptr;
  static thread_local int m_F = 0, m_NP = 0, m_R = 0;
  if (m_vs != a.Vs || m_x2 != a.X2t || m_F != a.F || m_NP != a.NP || m_R != a.R) {
    const uint64_t T8 = (uint64_t)a.NP / NB;
    const uint64_t d4[4] = {(uint64_t)NB, (uint64_t)a.F, T8, (uint64_t)a.R};
    const uint64_t s4[3] = {(uint64_t)NB * 4, (uint64_t)a.F * NB * 4, T8 * a.F * NB * 4};
    const uint32_t b4[4] = {(uint32_t)NB, (uint32_t)SFL, 1u, (uint32_t)a.R};
    int rc = tc::encode_f32_map(&maps.vs, a.Vs, 4, d4, s4, b4);
    if (rc == 0) rc = tc::encode_f32_map(&maps.x2, a.X2t, 3, d4, s4, b4);
    if (rc != 0) return fail(GVN_E_CUDA, "cuTensorMapEncodeTiled failed with %d (column sweep)", rc);
    m_vs = a.Vs; m_x2 = a.X2t; m_F = a.F; m_NP = a.NP; m_R = a.R;
  }
  k_cols_stream<KMAX, RT><<<grid, STT, smem, st>>>(maps, a, nstage);
  return check_launch("k_cols_stream");
}

template <int KMAX>
int32_t launch_cols_gen(const GenArgs& a, size_t smem, int grid, cudaStream_t st) {
  static size_t smem_tab[GVN_MAX_DEVICES] = {0};
  size_t& smem_set = *per_device_slot(smem_tab);
  if (smem_set != smem) {
    cudaError_t e = cudaFuncSetAttribute(k_cols_gen<KMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(GVN_E_CUDA, "k_cols_gen smem attr (%zu B): %s", smem, cudaGetErrorString(e));
    smem_set = smem;
  }
  k_cols_gen<KMAX><<<grid, GT, smem, st>>>(a);
  return check_launch("k_cols_gen");
}

template <int KMAX, int RT, int FT = 0>
int32_t launch_cols(const ColsArgs& a, size_t smem, int grid, cudaStream_t st) {
  static size_t smem_tab[GVN_MAX_DEVICES] = {0};
  size_t& smem_set = *per_device_slot(smem_tab);
  if (smem_set != smem) {
    cudaError_t e = cudaFuncSetAttribute(k_cols_v1<KMAX, RT, FT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(GVN_E_CUDA, "k_cols_v1 smem attr (%zu B): %s", smem, cudaGetErrorString(e));
    smem_set = smem;
  }
  k_cols_v1<KMAX, RT, FT><<<grid, CTT, smem, st>>>(a);
  return check_launch("k_cols_v1");
}

// Frame split of the W sweep: as many CTAs as the chip holds at once (one per SM for K > 16, else two) when the
// batch has few, long utterances; every split keeps at least 8 tiles.  Depends only on the batch shape, so the
// workspace can be sized from it.
inline int w_nsplit(const gvn_batch* b) {
  const int row_blocks = (b->F + WROWS - 1) / WROWS;
  const int slots = 148 * (b->K > 16 ? 1 : 2);
  const int tiles = b->NP / NB / b->B;                   // average tiles per utterance
  int ns = slots / (row_blocks * b->B);
  if (ns > tiles / 8) ns = tiles / 8;
  if (ns > 16) ns = 16;
  return ns < 1 ? 1 : ns;
}

template <int KMAX, int RT, int KX = 0>
int32_t launch_w(const gvn_batch* b, int R, const float* Mt, float* Wpart, cudaStream_t st) {
  const size_t stage = (size_t)w_stage_floats(b->K, R) * 4;
  const size_t budget = (KMAX > 16 || 2 * stage > 100 * 1024 ? 200 : 100) * 1024;    // one or two CTAs per SM
  int ws = (int)(budget / stage);
  if (ws > WS) ws = WS;
  if (ws < 2) return fail(GVN_E_UNSUPPORTED_SHAPE, "M-step W sweep: a ring stage of %zu bytes does not fit (R=%d)", stage, R);
  const size_t smem = (size_t)ws * stage;
  static size_t smem_tab[GVN_MAX_DEVICES] = {0};
  size_t& smem_set = *per_device_slot(smem_tab);
  if (smem_set != smem) {
    cudaError_t e = cudaFuncSetAttribute(k_w_v2<KMAX, RT, KX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(GVN_E_CUDA, "k_w_v2 smem attr (%zu B): %s", smem, cudaGetErrorString(e));
    smem_set = smem;
  }
  const int nsplit = w_nsplit(b);
  dim3 gw((b->F + WROWS - 1) / WROWS, b->B, nsplit);
  k_w_v2<KMAX, RT, KX><<<gw, WT, smem, st>>>(b->F, b->K, b->NP, R, ws, b->frame_off, b->n_frames, b->X2t, b->Vs, Mt, b->W, b->Wun, Wpart);
  int32_t rc = check_launch("k_w_v2");
  if (rc || nsplit == 1) return rc;
  const size_t total = (size_t)b->B * b->F * b->K;
  k_w_finish<<<(unsigned)((total + 255) / 256 < 592 ? (total + 255) / 256 : 592), 256, 0, st>>>(b->B, b->F, b->K, nsplit, b->W, Wpart, b->Wun);
  return check_launch("k_w_finish");
}

inline int kmax_of(int K) { return (K + 3) / 4 * 4; }
inline int ks_of(int K) { int ks = kmax_of(K); return (ks % 16 == 0) ? ks + 4 : ks; }

template <int KMAX>
int32_t launch_v1_k(const gvn_batch* b, int R, const ColsArgs& a, size_t smem, int grid, cudaStream_t st) {
  float* Wpart = const_cast<float*>(a.Mt) + (size_t)meta_stride(b->K, b->R_cap) * b->NP;     // behind the column data (mstep_v1_workspace_bytes)
  int32_t rc;
  if (R == 10) rc = (KMAX == 12 && b->K == 10) ? launch_w<KMAX, 10, (KMAX == 12 ? 10 : 0)>(b, R, a.Mt, Wpart, st) : launch_w<KMAX, 10>(b, R, a.Mt, Wpart, st);
  else rc = launch_w<KMAX, 0>(b, R, a.Mt, Wpart, st);
  if (rc) return rc;
  if (R == 10) return b->F == 513 ? launch_cols<KMAX, 10, 513>(a, smem, grid, st) : launch_cols<KMAX, 10>(a, smem, grid, st);
  return launch_cols<KMAX, 0>(a, smem, grid, st);
}

}  // namespace

inline int kmax_gen(int K) { return K <= 16 ? 16 : 32; }
inline int ks_gen(int K) { return kmax_gen(K) + 4; }     // dictionary row stride of the generic sweep: float4 reads cover KMAX columns

// variant 1 proper: the tile block fits in shared memory next to the dictionary
static bool v1_staged(const gvn_batch* b, int R) {
  if (b->K > 16) return false;
  return cols_smem_floats(b->F, ks_of(b->K), b->K, R, kmax_of(b->K)) * 4 <= 227 * 1024 &&
         (size_t)2 * w_stage_floats(b->K, R) * 4 <= 100 * 1024;
}

// true when the bulk-copy W sweep + one of the two column sweeps can run this shape
bool mstep_v1_supported(const gvn_batch* b, int R) {
  if (b->X2t == nullptr || b->K > 32) return false;
  if (v1_staged(b, R)) return true;
  const size_t wstage = (size_t)w_stage_floats(b->K, R) * 4;
  // (the generic sweep sums its frequency slices through 2048 floats of the a_s / s1_s arrays: 2 * F * 8 floats)
  return b->F >= 128 && 2 * wstage <= 200 * 1024 && gen_smem_floats(b->F, ks_gen(b->K), b->K, R) * 4 <= 227 * 1024;
}

// workspace of variant 1: the column data in tile order + the partial sums of a frame-split W sweep
size_t mstep_v1_workspace_bytes(const gvn_batch* b) {
  const int ns = w_nsplit(b);
  return ((size_t)meta_stride(b->K, b->R_cap) * b->NP + (ns > 1 ? (size_t)b->B * ns * b->F * b->K * 2 : 0)) * sizeof(float);
}

int32_t launch_mstep_v1(const gvn_batch* b, int R, float* cost_part, float* Mt, cudaStream_t st) {
  k_tile_meta<<<(b->NP + 255) / 256, 256, 0, st>>>(b->K, R, b->NP, b->H, b->g, b->Vs_w, Mt);
  int32_t rc = check_launch("k_tile_meta");
  if (rc) return rc;
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int ntiles = b->NP / NB;
  static int force_gen = -1;                            // GVN_MSTEP_GEN=1: take the generic column sweep everywhere (experiments)
  if (force_gen < 0) { const char* e = getenv("GVN_MSTEP_GEN"); force_gen = e ? atoi(e) : 0; }
  if (v1_staged(b, R) && !force_gen) {
    const int KMAX = kmax_of(b->K);
    ColsArgs a;
    a.F = b->F; a.K = b->K; a.KS = ks_of(b->K); a.NP = b->NP; a.R = R; a.B = b->B;
    a.ntiles = ntiles;
    const int NI = (b->F + CFL - 1) / CFL;
    a.nchunk = NI < MAXCH ? NI : MAXCH;
    a.frame_utt = b->frame_utt; a.frame_off = b->frame_off; a.X2t = b->X2t; a.Vs = b->Vs; a.Mt = Mt;
    a.Vb = b->Vb; a.g = b->g; a.H = b->H; a.Wun = b->Wun; a.W = b->W; a.cost_part = cost_part; a.XV = b->XV;
    const size_t smem = cols_smem_floats(b->F, a.KS, b->K, R, KMAX) * 4;
    const int grid = a.ntiles < sms ? a.ntiles : sms;
    switch (KMAX) {
      case 4: return launch_v1_k<4>(b, R, a, smem, grid, st);
      case 8: return launch_v1_k<8>(b, R, a, smem, grid, st);
      case 12: return launch_v1_k<12>(b, R, a, smem, grid, st);
      default: return launch_v1_k<16>(b, R, a, smem, grid, st);
    }
  }
  // generic shapes: bulk-copy W sweep (runtime R) + L2-resident column sweep
  float* Wpart = Mt + (size_t)meta_stride(b->K, b->R_cap) * b->NP;
  rc = b->K <= 16 ? launch_w<16, 0>(b, R, Mt, Wpart, st) : launch_w<32, 0>(b, R, Mt, Wpart, st);
  if (rc) return rc;
  GenArgs a;
  a.F = b->F; a.K = b->K; a.KS = ks_gen(b->K); a.NP = b->NP; a.R = R; a.ntiles = ntiles;
  a.frame_utt = b->frame_utt; a.frame_off = b->frame_off; a.X2t = b->X2t; a.Vs = b->Vs; a.Mt = Mt;
  a.Vb = b->Vb; a.g = b->g; a.H = b->H; a.Wun = b->Wun; a.W = b->W; a.cost_part = cost_part; a.XV = b->XV;
  static int force_scalar = -1;                         // GVN_MSTEP_GEN=2: the scalar-load generic sweep instead of the streamed one (experiments)
  if (force_scalar < 0) { const char* e = getenv("GVN_MSTEP_GEN"); force_scalar = e ? atoi(e) == 2 : 0; }
  {
    const size_t fixed = stream_fixed_floats(b->F, a.KS, b->K, R) * 4, stage = (size_t)stream_stage_floats(R) * 4;
    int nstage = fixed < 227 * 1024 ? (int)((227 * 1024 - fixed) / stage) : 0;
    if (nstage > SMAXST) nstage = SMAXST;
    if (nstage >= 3 && !force_scalar) {
      const int grid_s = ntiles < sms ? ntiles : sms;
      const size_t smem_s = fixed + (size_t)nstage * stage;
      if (R == 10) return b->K <= 16 ? launch_cols_stream<16, 10>(a, nstage, smem_s, grid_s, st) : launch_cols_stream<32, 10>(a, nstage, smem_s, grid_s, st);
      return b->K <= 16 ? launch_cols_stream<16, 0>(a, nstage, smem_s, grid_s, st) : launch_cols_stream<32, 0>(a, nstage, smem_s, grid_s, st);
    }
  }
  const size_t smem = gen_smem_floats(b->F, a.KS, b->K, R) * 4;
  const int per_sm = smem <= 110 * 1024 ? 2 : 1;
  const int grid = ntiles < sms * per_sm ? ntiles : sms * per_sm;
  return b->K <= 16 ? launch_cols_gen<16>(a, smem, grid, st) : launch_cols_gen<32>(a, smem, grid, st);
}

}  // namespace gvn
