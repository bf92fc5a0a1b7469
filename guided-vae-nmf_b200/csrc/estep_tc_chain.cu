// E-step on the Blackwell tensor cores (GVN_PREC_F16): one Metropolis-Hastings chain per frame,
// a persistent warp-specialised CTA per tile of 128 frames.
//
// Same algorithm and same outputs as estep_simt.cu (reference python/models/mcem.py:218-307 /
// :371-454); what changes is where the arithmetic runs:
//   * the 128 frames of the tile are the M dimension of tcgen05.mma: 128 TMEM lanes = 128 frames, so
//     after tcgen05.ld every epilogue thread owns one frame row and the sum over frequency of
//     the log acceptance ratio (mcem.py:266-268) is a private register accumulation;
//   * activations never touch shared memory: the epilogue threads write tanh(.) as packed f16
//     straight into TMEM (tcgen05.st) and the next layer's MMA reads its A operand from there;
//     the latent z goes in as an f16 hi+lo pair against a duplicated W1 (K = 2*L), so the
//     random walk itself keeps fp32 resolution;
//   * weights sit in shared memory for the whole chain as UMMA operand images (packed once by
//     gvn_pack_decoder; f16 because activations are in [-1,1] and weights O(0.05): 11 mantissa
//     bits = TF32 precision at twice the MMA rate and half the footprint);
//   * the output layer (F=513 -> 4 chunks of 128 columns + one of 16) is double-buffered in TMEM:
//     the MMA of chunk c+1 runs while the epilogue consumes chunk c;
//   * the per-bin constants of the chain, X2 and Vb, are packed once per launch into one 32-bit
//     word per (f, frame) (bf16 pair; both appear on the two sides of the acceptance ratio, so the
//     rounding is a fixed perturbation of the target density, not noise) and stream through a TMA
//     ring of [32 frequency rows][128 frames] boxes, produced by a dedicated warp;
//   * the epilogue lives on the special-function pipe (exp per bin, log and reciprocal of Vx, 8 cycles
//     per warp instruction and scheduler): bins are processed in pairs, log a + log b = log(ab) and
//     x/a + y/b = (xb + ya)/(ab); the pair products are split into exponent and mantissa (integer sums / a running
//     product < 2^8) and the quotients carried as one fraction, so a stage of 8 pairs needs ONE lg2 and ONE rcp;
//     tanh as tanh.approx.f16x2;
//   * the per-frame bias of the first layer (label projection + b1) sits in TMEM for the whole chain;
//     the proposal noise and log u of step m+1 are drawn by the warps that idle while the owners
//     accept/propose in step m and handed over through TMEM (L <= 16).
// Kept samples: slot r of Vs receives the proposal of kept step r (speculative store from the
// epilogue); Vs_w[r][n] is the multiplicity of slot r (0 when that proposal was rejected, k+1 when
// the k following proposals were) -- see include/gvn.h.
// Warp roles: 0-7 epilogue (warp w and w+4 share TMEM lane quarter w%4 and split the columns),
// 8 = MMA issuer (one elected lane), 9 = TMA producer (one elected lane).
#include <cuda.h>
#include <cuda_bf16.h>

#include <stdlib.h>

#include <type_traits>

#include "tc_chain_common.cuh"

namespace gvn {

namespace {

constexpr int NE = 256;                 // epilogue threads
constexpr int NEW = NE / 32;            // epilogue warps
constexpr int NTHREADS = 320;
constexpr int STAGE_BYTES = SROWS * TM * 4;
#ifndef GVN_TC_DEFAULT_VARIANT
#define GVN_TC_DEFAULT_VARIANT 84
#endif

// TMEM column map (512 columns allocated)
// The hidden-layer accumulator aliases output buffer 0 (the two are never live together), which
// leaves columns 256..383 for the per-frame bias of the first layer (label projection + b1, fp32).
constexpr uint32_t COL_ACC0 = 0, COL_ACC1 = 128, COL_ACCH = 0, COL_YP = 256, COL_A = 384, COL_Z = 448;
constexpr uint32_t COL_LOGU = 464, COL_EPS = 480;     // L16 == 16 only: noise of the next step, double-buffered (see PRE)

__device__ __forceinline__ void bar_epilogue() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

template <int L16, bool PROF_ON, int VAR>
__global__ void __launch_bounds__(NTHREADS, 1) k_estep_tc(const __grid_constant__ CUtensorMap tm_xv, TcArgs p) {
  constexpr int K1 = 2 * L16;                             // z hi | z lo against [W1 | W1]
  // Hidden-layer hand-over in NBLK blocks: the epilogue warps signal every 64/NBLK columns of tanh output, the issuer
  // starts the next layer's MMA k-steps block by block, so the MMA runs under the rest of the tanh work.  The
  // second hidden layer then accumulates in output buffer 1 (the first layer's accumulator, aliased to buffer 0,
  // is still being read when its first k-steps start).
  constexpr int NBLK = (VAR & 16) ? 4 : ((VAR & 8) ? 2 : 1);
  constexpr int CPB = 64 / NBLK;                          // columns per block and half
  constexpr uint32_t COL_ACCH2 = NBLK > 1 ? COL_ACC1 : COL_ACCH;
  extern __shared__ __align__(1024) unsigned char smem[];
  const int F = p.F, FN = p.FN, NP = p.NP, L = p.L;
  const int NCH = (FN + 127) / 128;                       // output-layer chunks (last one may be 16 wide)
  // ---- shared memory carve-up ----
  unsigned char* sW1 = smem;                              // [128][K1] f16 image
  unsigned char* sW2 = sW1 + HID * K1 * 2;                // [128][128]
  unsigned char* sW3 = sW2 + HID * HID * 2;               // [FN][128]
  unsigned char* sRing = sW3 + (size_t)FN * HID * 2;      // nstage x [32][128] u32 (bf16 X2 | bf16 Vb)
  float* sB3 = reinterpret_cast<float*>(sRing + (size_t)p.nstage * STAGE_BYTES);   // [FN + 16]  b3*log2e
  float* sB2 = sB3 + FN + 16;                             // [128]
  double* sPart = reinterpret_cast<double*>(sB2 + HID);   // [2][256]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sPart + 2 * NE);
  uint64_t* bar_z = bars + 0;                             // E(0-3) -> M : Z operand in TMEM
  uint64_t* bar_h = bars + 1;                             // E -> M : hidden activations in TMEM, ACCH drained
  uint64_t* bar_hfull = bars + 2;                         // M -> E : hidden-layer accumulator full
  uint64_t* bar_d = bars + 3;                             // [2] M -> E : accumulator buffer full
  uint64_t* bar_free = bars + 5;                          // [2] E -> M : accumulator buffer drained
  uint64_t* bar_full = bars + 7;                          // [nstage] TMA -> E
  uint64_t* bar_empty = bars + 7 + MAX_STAGES;            // [nstage] E -> P
  uint64_t* bar_eps = bars + 7 + 2 * MAX_STAGES;          // E(owners) -> E(noise warps): hand-over buffer has been read
  uint64_t* bar_hb = bars + 8 + 2 * MAX_STAGES;           // [4] E -> M : one 16/32-column block of the hidden activations is in TMEM (NBLK > 1)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12 + 2 * MAX_STAGES);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n0 = blockIdx.x * TM;

  // ---- one-time setup: weights -> smem, barriers, TMEM ----
  {
    const uint4* src; uint4* dst; int n16;
    src = reinterpret_cast<const uint4*>(p.img + p.off_w1d); dst = reinterpret_cast<uint4*>(sW1); n16 = HID * K1 * 2 / 16;
    for (int i = tid; i < n16; i += NTHREADS) dst[i] = src[i];
    src = reinterpret_cast<const uint4*>(p.img + p.off_w2); dst = reinterpret_cast<uint4*>(sW2); n16 = HID * HID * 2 / 16;
    for (int i = tid; i < n16; i += NTHREADS) dst[i] = src[i];
    src = reinterpret_cast<const uint4*>(p.img + p.off_w3); dst = reinterpret_cast<uint4*>(sW3); n16 = FN * HID * 2 / 16;
    for (int i = tid; i < n16; i += NTHREADS) dst[i] = src[i];
    const float* b3s = reinterpret_cast<const float*>(p.img + p.off_b3s);
    for (int i = tid; i < FN + 16; i += NTHREADS) sB3[i] = i < FN ? b3s[i] : 0.f;
    for (int i = tid; i < HID; i += NTHREADS) sB2[i] = p.b2[i];
  }
  if (tid == 0) {
    mbar_init(bar_z, 4);
    mbar_init(bar_h, NEW);
    mbar_init(bar_hfull, 1);
    mbar_init(bar_eps, 4);
    for (int i = 0; i < 4; ++i) mbar_init(bar_hb + i, NEW);
    for (int i = 0; i < 2; ++i) { mbar_init(bar_d + i, 1); mbar_init(bar_free + i, NEW); }
    for (int i = 0; i < p.nstage; ++i) { mbar_init(bar_full + i, 1); mbar_init(bar_empty + i, NEW); }
    mbar_init_fence();
  }
  if (warp == 8) tmem_alloc(tmem_slot, 512);
  fence_proxy_async();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tbase = *tmem_slot;

  const int n_steps = p.burnin + p.R;
  const int NST = (F + SROWS - 1) / SROWS;                // ring stages per streaming pass
  // pass schedule, identical for all roles: INIT, then per step m: PROP, and WRITE after the step m == burnin

  if (warp == 9) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      uint32_t slot = 0, ph = 0;
      const int n_load_passes = 1 + n_steps;
      for (int ps = 0; ps < n_load_passes; ++ps) {
        for (int s = 0; s < NST; ++s) {
          mbar_wait_idle(bar_empty + slot, ph ^ 1);
          mbar_expect_tx(bar_full + slot, STAGE_BYTES);
          tma_load_2d(sRing + (size_t)slot * STAGE_BYTES, &tm_xv, n0, s * SROWS, bar_full + slot);
          if (++slot == (uint32_t)p.nstage) { slot = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 8) {
    // =============================== MMA issuer ===============================
    if (lane == 0) {
      uint32_t n_z = 0, n_h = 0, n_free0 = 0, n_free1 = 0;
      const uint32_t sbo1 = img_sbo(K1), sbo = img_sbo(HID), lbo = img_lbo();
      const int n_pass = 1 + n_steps + (p.R > 0 ? 1 : 0);
      for (int ps = 0; ps < n_pass; ++ps) {
        // layer 1: ACCH = [z_hi | z_lo](128 x K1) * [W1 | W1]^T
        mbar_wait_idle(bar_z, n_z++ & 1);
        fence_after();
#pragma unroll
        for (int k0 = 0; k0 < K1; k0 += 16)
          mma_ts(tbase + COL_ACCH, tbase + COL_Z + k0 / 2, smem_desc(smem_u32(sW1) + (k0 / 8) * 128, lbo, sbo1),
                 idesc_f16(TM, HID), k0 > 0);
        mma_commit(bar_hfull);
        // layer 2: ACCH2 = H1(128 x 128) * W2^T
        auto issue_blocks = [&](uint32_t d_col, uint32_t b_base, uint32_t sbo_, uint32_t idesc_) {
#pragma unroll
          for (int blk = 0; blk < NBLK; ++blk) {
            mbar_wait_idle(bar_hb + blk, n_h & 1);
            fence_after();
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
              for (int kk = 0; kk < CPB; kk += 16) {
                const int k0 = 64 * h + blk * CPB + kk;
                mma_ts(tbase + d_col, tbase + COL_A + k0 / 2, smem_desc(b_base + (k0 / 8) * 128, lbo, sbo_), idesc_,
                       !(blk == 0 && h == 0 && kk == 0));
              }
          }
          ++n_h;
        };
        if constexpr (NBLK > 1) {
          issue_blocks(COL_ACCH2, smem_u32(sW2), sbo, idesc_f16(TM, HID));
        } else {
          mbar_wait_idle(bar_h, n_h++ & 1);
          fence_after();
#pragma unroll
          for (int k0 = 0; k0 < HID; k0 += 16)
            mma_ts(tbase + COL_ACCH, tbase + COL_A + k0 / 2, smem_desc(smem_u32(sW2) + (k0 / 8) * 128, lbo, sbo),
                   idesc_f16(TM, HID), k0 > 0);
        }
        mma_commit(bar_hfull);
        // layer 3: chunks of the output features, alternating accumulator buffers
        if constexpr (NBLK == 1) {
          mbar_wait_idle(bar_h, n_h++ & 1);
          fence_after();
        }
        for (int c = 0; c < NCH; ++c) {
          const int buf = c & 1, ncol = min(128, FN - c * 128);
          if (buf) mbar_wait_idle(bar_free + 1, (n_free1++ & 1) ^ 1); else mbar_wait_idle(bar_free + 0, (n_free0++ & 1) ^ 1);
          fence_after();
          const uint32_t b0 = smem_u32(sW3) + (uint32_t)(c * 16) * sbo;
          const uint32_t idesc = idesc_f16(TM, ncol);
          if (NBLK > 1 && c == 0) {                          // first chunk: k-steps follow the blocks of the second tanh layer
            issue_blocks(COL_ACC0, b0, sbo, idesc);
          } else {
#pragma unroll
            for (int k0 = 0; k0 < HID; k0 += 16)
              mma_ts(tbase + (buf ? COL_ACC1 : COL_ACC0), tbase + COL_A + k0 / 2, smem_desc(b0 + (k0 / 8) * 128, lbo, sbo),
                     idesc, k0 > 0);
          }
          mma_commit(bar_d + buf);
        }
      }
    }
  } else {
    // =============================== epilogue warps ===============================
    const int row = tid & 127, half = tid >> 7, q = warp & 3;
    const uint32_t tlane = tbase + ((uint32_t)(q * 32) << 16);
    const int n = n0 + row;
    const bool in_range = n < NP;
    const bool valid = in_range && p.frame_utt[n] >= 0;
    const float g = valid ? p.g[n] : 1.f;
    uint32_t n_d0 = 0, n_d1 = 0, n_hf = 0, n_dec = 0;
    uint32_t slot = 0, ph = 0;                            // ring position of the next stage
    uint32_t ring_ok = 0;                                 // early probe result for that stage (V_PROBE)
    unsigned long long pc[PROF_ON ? 12 : 1] = {0}, pt = PROF_ON ? clock64() : 0ull;
#define PROF(i) do { if constexpr (PROF_ON) { const unsigned long long t_ = clock64(); pc[i] += t_ - pt; pt = t_; } } while (0)
    constexpr bool PRE = (L16 == 16);
    float z[L16];
    float zp[PRE ? 1 : L16];                              // without the TMEM noise hand-over the proposal is kept
    // label projection + b1 of this thread's 64 hidden units -> TMEM (constant for the whole chain)
#pragma unroll
    for (int c0 = 0; c0 < 64; c0 += 16) {
      uint32_t w[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) w[i] = in_range ? __float_as_uint(p.yproj[(size_t)(64 * half + c0 + i) * NP + n]) : 0u;
      tmem_st16(tlane + COL_YP + 64 * half + c0, w);
    }
    tmem_st_wait();
    double Ct = 0.0;
    int n_acc = 0;
    if (half == 0) {
#pragma unroll
      for (int l = 0; l < L16; ++l) z[l] = (l < L && in_range) ? p.Z[(size_t)l * NP + n] : 0.f;
      if (valid) for (int s = 0; s < p.R; ++s) p.Vs_w[(size_t)s * NP + n] = 0.f;
    }

    // Noise of a step does not depend on the chain state.  With L16 == 16 there is room in TMEM to
    // take it off the critical path: warps 4-7, which idle while the owners build a proposal, draw
    // eps and log u of step m+1 at the start of step m and hand them over through TMEM columns
    // (double-buffered on the step parity; ordered by the energy barrier that ends every step).
    auto draw_eps = [&](int m, float (&e)[L16]) {
      if (p.eps != nullptr) {
#pragma unroll
        for (int l = 0; l < L16; ++l) e[l] = (l < L && in_range) ? p.eps[((size_t)m * L + l) * NP + n] : 0.f;
      } else {
#pragma unroll
        for (int lq = 0; lq < L16 / 4; ++lq) {
          uint4 rr = philox4x32_10(make_uint4((uint32_t)n, (uint32_t)m, (uint32_t)lq, (uint32_t)p.chain),
                                   make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32)));
          float2 a = box_muller(rr.x, rr.y), b = box_muller(rr.z, rr.w);
          const float e4[4] = {a.x, a.y, b.x, b.y};
#pragma unroll
          for (int i = 0; i < 4; ++i) e[4 * lq + i] = (4 * lq + i < L) ? e4[i] : 0.f;
        }
      }
    };
    auto draw_logu = [&](int m) -> float {
      float uu = 0.5f;
      if (valid) {
        if (p.u != nullptr) {
          uu = p.u[(size_t)m * NP + n];
        } else {
          uint4 rr = philox4x32_10(make_uint4((uint32_t)n, (uint32_t)m, 0xffffffffu, (uint32_t)p.chain),
                                   make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32)));
          uu = u01(rr.x);
        }
      }
      return logf(uu);
    };
    auto stage_noise = [&](int m) {                   // warps 4-7 (PRE): noise of step m -> TMEM
      float e[L16];
      draw_eps(m, e);
      uint32_t w[16];
#pragma unroll
      for (int l = 0; l < 16; ++l) w[l] = __float_as_uint(e[l % L16]);
      // buffer (m & 1) held the noise of step m-2; the owners signal when they have read it for the last time
      // (long before this point: the draw above takes thousands of cycles, their read a few hundred)
      if (m >= 2) mbar_wait(bar_eps, (uint32_t)(m - 2) & 1);
      __syncwarp();
      tmem_st16(tlane + COL_EPS + 16 * (m & 1), w);
      tmem_st1(tlane + COL_LOGU + (m & 1), __float_as_uint(draw_logu(m)));
      tmem_st_wait();
      fence_before();
    };

    // V_SPLIT: the draw of a step takes ~3.4 k cycles on warps 4-7, the owners' accept + propose + first-layer wait only ~2 k
    // -- the noise warps reach the first hidden layer late and the whole tile waits for their half of its columns (in-kernel
    // counters: the owners wait 1.8 k cycles for the second layer's accumulator, the noise warps 0.4 k).  So only log u and
    // the first two quads of eps are drawn at the start of the step; the other two quads follow where every epilogue warp
    // idles anyway: while the first output chunk's MMA completes (nz_m carries the step number to that point of decode()).
    constexpr bool V_SPLIT = PRE && (VAR & 64) != 0;
    int nz_m = -1;
    auto draw_quads = [&](int m, int lq0, uint32_t (&w)[8]) {     // quads lq0, lq0 + 1 of the noise of step m
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int lq = lq0 + j;
        if (p.eps != nullptr) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            w[4 * j + i] = __float_as_uint((4 * lq + i < L && in_range) ? p.eps[((size_t)m * L + 4 * lq + i) * NP + n] : 0.f);
        } else {
          uint4 rr = philox4x32_10(make_uint4((uint32_t)n, (uint32_t)m, (uint32_t)lq, (uint32_t)p.chain),
                                   make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32)));
          float2 a = box_muller(rr.x, rr.y), b = box_muller(rr.z, rr.w);
          const float e4[4] = {a.x, a.y, b.x, b.y};
#pragma unroll
          for (int i = 0; i < 4; ++i) w[4 * j + i] = __float_as_uint((4 * lq + i < L) ? e4[i] : 0.f);
        }
      }
    };
    auto stage_noise_a = [&](int m) {                 // warps 4-7: log u and eps[0..8) of step m -> TMEM
      uint32_t w[8];
      draw_quads(m, 0, w);
      const float lu = draw_logu(m);
      if (m >= 2) mbar_wait(bar_eps, (uint32_t)(m - 2) & 1);     // as in stage_noise
      __syncwarp();
      tmem_st8(tlane + COL_EPS + 16 * (m & 1), w);
      tmem_st1(tlane + COL_LOGU + (m & 1), __float_as_uint(lu));
      tmem_st_wait();
      fence_before();
      nz_m = m;
    };
    auto stage_noise_b = [&]() {                      // warps 4-7: eps[8..16) of step nz_m -> TMEM
      uint32_t w[8];
      draw_quads(nz_m, 2, w);
      __syncwarp();
      tmem_st8(tlane + COL_EPS + 16 * (nz_m & 1) + 8, w);
      tmem_st_wait();
      fence_before();
      nz_m = -1;
    };

    auto put_z = [&](const float (&zz)[L16]) {       // warps 0-3: Z operand (hi | lo) -> TMEM, signal the issuer
#pragma unroll
      for (int k0 = 0; k0 < L16; k0 += 16) {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float a0 = zz[k0 + 2 * j], a1 = zz[k0 + 2 * j + 1];
          const __half h0 = __float2half_rn(a0), h1 = __float2half_rn(a1);
          hi[j] = pack_f16(__half2float(h0), __half2float(h1));
          lo[j] = pack_f16(a0 - __half2float(h0), a1 - __half2float(h1));
        }
        tmem_st8(tlane + COL_Z + k0 / 2, hi);
        tmem_st8(tlane + COL_Z + (L16 + k0) / 2, lo);
      }
      tmem_st_wait();
      fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_z);
    };

    // hidden layer epilogue: tanh(acc/256 + bias) of this thread's 64 columns -> f16 A operand in TMEM
    auto hidden = [&](auto layer0) {
      PROF(0);
      mbar_wait(bar_hfull, n_hf++ & 1);
      PROF(decltype(layer0)::value ? 1 : 3);
      fence_after();
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 16) {
        uint32_t r[16], bq[16];
        tmem_ld16(tlane + (decltype(layer0)::value ? COL_ACCH : COL_ACCH2) + 64 * half + c0, r);
        if constexpr (decltype(layer0)::value) tmem_ld16(tlane + COL_YP + 64 * half + c0, bq);
        tmem_ld_wait();
        uint32_t o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float b0, b1;
          if constexpr (decltype(layer0)::value) {
            b0 = __uint_as_float(bq[2 * j]); b1 = __uint_as_float(bq[2 * j + 1]);
          } else {
            const float2 t = *reinterpret_cast<const float2*>(sB2 + 64 * half + c0 + 2 * j);
            b0 = t.x; b1 = t.y;
          }
          o[j] = tanh_f16x2(fmaf(__uint_as_float(r[2 * j]), W_SCALE_INV, b0), fmaf(__uint_as_float(r[2 * j + 1]), W_SCALE_INV, b1));
        }
        tmem_st8(tlane + COL_A + (64 * half + c0) / 2, o);
        if (NBLK > 1 && (c0 + 16) % CPB == 0) {             // this block of the A operand is complete
          tmem_st_wait();
          fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_hb + (c0 + 16) / CPB - 1);
        }
      }
      if (NBLK == 1) {
        tmem_st_wait();
        fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_h);
      }
      PROF(decltype(layer0)::value ? 2 : 4);
    };

    // one decoder evaluation.  ENERGY: stream X2/Vb and return sum_f log Vx + X2/Vx of the row
    // (valid in half 0); STORE: write the speech variance of the row to vs_out[f*NP] (f = 0..F)
    auto decode = [&](auto energy_t, auto store_t, float* vs_out) -> double {
      constexpr bool ENERGY = decltype(energy_t)::value, STORE = decltype(store_t)::value;
      hidden(std::true_type{});
      hidden(std::false_type{});
      if constexpr (V_SPLIT) { if (nz_m >= 0) stage_noise_b(); }     // under the wait for the first output chunk
      double dsum = 0.0;
      float* vo = STORE ? vs_out + tile_off(16 * half, n, F) : nullptr;   // column-tile order: bin stride = GVN_VS_TILE floats
#pragma unroll 1
      for (int c = 0; c < NCH; ++c) {
        const int buf = c & 1, ncol = min(128, FN - c * 128);
        if (buf) mbar_wait(bar_d + 1, n_d1++ & 1); else mbar_wait(bar_d + 0, n_d0++ & 1);
        PROF(5);
        fence_after();
        const uint32_t acc_col = (buf ? COL_ACC1 : COL_ACC0) + 16 * half;
        const int nst = (ncol + SROWS - 1) / SROWS;
        float sl = 0.f, sr = 0.f;                          // fp32 inside a chunk (<= 64 pairs), double across chunks
        int es = 0;                                        // V_FRAC: sum of the unbiased exponents of the pair products
        // one stage = this thread's 16 bins of a 32-row ring box; the accumulator columns of stage s+1
        // are requested from TMEM before the math of stage s (two register sets, statically indexed)
        constexpr bool V_PF = (VAR & 1) != 0;             // prefetch the next stage's accumulator columns
        constexpr bool V_PROBE = (VAR & 2) != 0;          // probe the next stage's ring barrier before this stage's math
        constexpr bool V_FRAC = (VAR & 4) != 0;           // one lg2 + one rcp per 8 pairs: exponent/mantissa split of the pair products
        auto stage = [&](uint32_t (&r)[16], uint32_t (&rn)[16], int s_) {
          const int f0 = c * 128 + s_ * SROWS + 16 * half;
          const uint32_t* xv = nullptr;
          if (!V_PF) tmem_ld16(tlane + acc_col + s_ * SROWS, r);
          if (ENERGY) {
            PROF(7);
            if (!(V_PROBE && ring_ok)) mbar_wait(bar_full + slot, ph);
            PROF(6);
            xv = reinterpret_cast<const uint32_t*>(sRing + (size_t)slot * STAGE_BYTES) + (16 * half) * TM + row;
            if (V_PROBE) {                                 // is the stage after this one already there?
              uint32_t ns = slot + 1, np_ = ph;
              if (ns == (uint32_t)p.nstage) { ns = 0; np_ ^= 1; }
              ring_ok = mbar_test(bar_full + ns, np_);
            }
          }
          if (V_PF && s_ + 1 < nst) tmem_ld16(tlane + acc_col + (s_ + 1) * SROWS, rn);   // prefetch
          float* vo = STORE ? vs_out + tile_off(f0, n, F) : nullptr;              // column-tile order: bin stride = GVN_VS_TILE floats
          if (f0 + 16 <= F) {                              // full group: 8 pairs of bins
            float b3v[16];
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const float4 t = *reinterpret_cast<const float4*>(sB3 + f0 + 4 * j4);
              b3v[4 * j4] = t.x; b3v[4 * j4 + 1] = t.y; b3v[4 * j4 + 2] = t.z; b3v[4 * j4 + 3] = t.w;
            }
            if (!V_PF) tmem_ld_wait();
            if (V_FRAC && ENERGY) {
              // sum_j log(pr_j) = ln2 * (sum_j e_j + lg2(prod_j m_j)) and sum_j qn_j/pr_j = N/P with pr_j = m_j * 2^e_j,
              // m_j in [1,2): the exponents are summed as integers, the mantissas multiplied (P < 2^8), the quotients
              // (scaled by 2^-e_j, exact) carried as one fraction N/P -- 2 special-function operations per 8 pairs
              // instead of 16, paid with integer/FMA-pipe instructions
              float Pm = 1.f, Nn = 0.f;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float t0 = fmaf(__uint_as_float(r[2 * j]), SC3, b3v[2 * j]), t1 = fmaf(__uint_as_float(r[2 * j + 1]), SC3, b3v[2 * j + 1]);
                const float v0 = ex2_approx(t0);
                const float v1 = ex2_approx(t1);
                const uint32_t w0 = xv[(2 * j) * TM], w1 = xv[(2 * j + 1) * TM];
                const float a = fmaf(g, v0, __uint_as_float(w0 << 16)), b = fmaf(g, v1, __uint_as_float(w1 << 16));
                const float pr = a * b;
                const float qn = fmaf(__uint_as_float(w0), b, __uint_as_float(w1) * a);
                const uint32_t pb = __float_as_uint(pr);
                const float m = __uint_as_float((pb & 0x007fffffu) | 0x3f800000u);
                es += (int)(pb >> 23);
                const float q = qn * __uint_as_float(0x7f000000u - (pb & 0x7f800000u));       // qn * 2^-e
                Nn = fmaf(q, Pm, Nn * m);
                Pm *= m;
                if (STORE && valid) { st_stream(vo + (2 * j) * GVN_VS_TILE, fminf(v0, GVN_VS_MAX)); st_stream(vo + (2 * j + 1) * GVN_VS_TILE, fminf(v1, GVN_VS_MAX)); }
              }
              es -= 127 * 8;
              sl += lg2_approx(Pm);
              sr = fmaf(Nn, rcp_approx(Pm), sr);
            } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float t0 = fmaf(__uint_as_float(r[2 * j]), SC3, b3v[2 * j]), t1 = fmaf(__uint_as_float(r[2 * j + 1]), SC3, b3v[2 * j + 1]);
              const float v0 = ex2_approx(t0);
              const float v1 = ex2_approx(t1);
              if (ENERGY) {
                const uint32_t w0 = xv[(2 * j) * TM], w1 = xv[(2 * j + 1) * TM];
                const float a = fmaf(g, v0, __uint_as_float(w0 << 16)), b = fmaf(g, v1, __uint_as_float(w1 << 16));
                const float pr = a * b;
                sl += lg2_approx(pr);
                const float qn = fmaf(__uint_as_float(w0), b, __uint_as_float(w1) * a);      // X2 = the whole word (k_pack_xv)
                sr = fmaf(qn, rcp_approx(pr), sr);
              }
              if (STORE && valid) { st_stream(vo + (2 * j) * GVN_VS_TILE, fminf(v0, GVN_VS_MAX)); st_stream(vo + (2 * j + 1) * GVN_VS_TILE, fminf(v1, GVN_VS_MAX)); }
            }
            }
          } else {                                         // ragged tail of the spectrum (F = 513: one bin)
            const int nv = F - f0;
            if (!V_PF) tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              if (j >= nv) break;
              const float v0 = ex2_approx(fmaf(__uint_as_float(r[j]), SC3, sB3[f0 + j]));
              if (ENERGY) {
                const uint32_t w0 = xv[j * TM];
                const float a = fmaf(g, v0, __uint_as_float(w0 << 16));
                sl += lg2_approx(a);
                sr = fmaf(__uint_as_float(w0), rcp_approx(a), sr);
              }
              if (STORE && valid) st_stream(vo + j * GVN_VS_TILE, fminf(v0, GVN_VS_MAX));
            }
          }
          if (ENERGY) {
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_empty + slot);
            if (++slot == (uint32_t)p.nstage) { slot = 0; ph ^= 1; }
          }
          if (V_PF) tmem_ld_wait();                        // the prefetched columns are in registers
        };
        uint32_t ra[16], rb[16];
        if (V_PF) { tmem_ld16(tlane + acc_col, ra); tmem_ld_wait(); }
        if constexpr (!V_PF && (VAR & 32) != 0) {            // one stage per loop iteration (half the code of the output loop)
#pragma unroll 1
          for (int s = 0; s < nst; ++s) stage(ra, rb, s);
        } else {
#pragma unroll 1
          for (int s = 0; s < nst; s += 2) {
            stage(ra, rb, s);
            if (s + 1 < nst) stage(V_PF ? rb : ra, ra, s + 1);
          }
        }
        if (ENERGY) dsum += (double)fmaf(V_FRAC ? sl + (float)es : sl, LN2, sr);
        fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_free + buf);
        PROF(7);
      }
      if (!ENERGY) return 0.0;
      double* part = sPart + (n_dec & 1) * NE;
      ++n_dec;
      part[tid] = dsum;
      bar_epilogue();
      PROF(8);
      return part[row] + part[row + 128];
    };
    const std::true_type T{};
    const std::false_type N{};

    // ---- chain ----
    if (half == 0) put_z(z);
    if (PRE && half == 1 && n_steps > 0) { if constexpr (V_SPLIT) stage_noise_a(0); else stage_noise(0); }
    Ct = decode(T, N, nullptr);
    int cur = 0;
    float cnt = 0.f;

#pragma unroll 1
    for (int m = 0; m < n_steps; ++m) {
      float prior = 0.f, logu = 0.f;
      if (half == 0) {
        // proposal Z' = Z + sd * eps  (mcem.py:257)
        float e[L16];
        if (PRE) {
          uint32_t w[16];
          fence_after();
          __syncwarp();
          tmem_ld16(tlane + COL_EPS + 16 * (m & 1), w);
          logu = __uint_as_float(tmem_ld1(tlane + COL_LOGU + (m & 1)));
          tmem_ld_wait();
#pragma unroll
          for (int l = 0; l < L16; ++l) e[l] = __uint_as_float(w[l % 16]);
        } else {
          draw_eps(m, e);
        }
        float zq[L16];
#pragma unroll
        for (int l = 0; l < L16; ++l) zq[l] = fmaf(p.sd, e[l], z[l]);
        if (!PRE) {
#pragma unroll
          for (int l = 0; l < L16; ++l) zp[l % (PRE ? 1 : L16)] = zq[l];
        }
        put_z(zq);                                          // the issuer waits for this; the prior is not needed before the accept
#pragma unroll
        for (int l = 0; l < L16; ++l) prior += z[l] * z[l] - zq[l] * zq[l];
      } else if (PRE && m + 1 < n_steps) {
        if constexpr (V_SPLIT) stage_noise_a(m + 1); else stage_noise(m + 1);
      }
      const int r = m - p.burnin;
      PROF(9);
      const double Cp = (r >= 1) ? decode(T, T, p.Vs + (size_t)r * F * NP) : decode(T, N, nullptr);

      // accept / reject (mcem.py:266-280)
      if (half == 0) {
        const float acc_prob = (float)(Ct - Cp) + 0.5f * prior;
        int ok = 0;
        if (valid) {
          if (!PRE) logu = draw_logu(m);
          ok = logu < acc_prob;
          if (p.forced != nullptr) ok = p.forced[(size_t)m * NP + n] != 0;
          if (p.t_acc != nullptr) p.t_acc[(size_t)m * NP + n] = acc_prob;
          if (p.t_dec != nullptr) p.t_dec[(size_t)m * NP + n] = (uint8_t)ok;
        }
        if (PRE) {                                          // z += sd * eps with the same noise, re-read from TMEM
          uint32_t w[16];                                   // (aligned TMEM load: executed by the whole warp)
          __syncwarp();
          tmem_ld16(tlane + COL_EPS + 16 * (m & 1), w);
          tmem_ld_wait();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_eps);              // phase m: buffer (m & 1) may be overwritten
          if (ok) {
#pragma unroll
            for (int l = 0; l < L16; ++l) z[l] = fmaf(p.sd, __uint_as_float(w[l % 16]), z[l]);
          }
        } else if (ok) {
#pragma unroll
          for (int l = 0; l < L16; ++l) z[l] = zp[l % (PRE ? 1 : L16)];
        }
        if (ok) { Ct = Cp; ++n_acc; }
        if (r >= 0 && p.t_zs != nullptr && valid) {
#pragma unroll
          for (int l = 0; l < L16; ++l) if (l < L) p.t_zs[((size_t)r * L + l) * NP + n] = z[l];
        }
        // multiplicities of the kept samples (mcem.py:286-289: a rejected step repeats the state)
        if (r == 0) { cur = 0; cnt = 1.f; }
        else if (r >= 1) {
          if (ok) { if (valid) p.Vs_w[(size_t)cur * NP + n] = cnt; cur = r; cnt = 1.f; }
          else cnt += 1.f;
        }
      }
      // slot 0 holds the state after the burn-in: decode it once (mcem.py:286-289 + compute_Vs)
      if (r == 0) {
        if (half == 0) put_z(z);
        decode(N, T, p.Vs);
      }
    }
    PROF(9);
    if constexpr (PROF_ON) {
      if (p.prof != nullptr && lane == 0)
        for (int i = 0; i < 12; ++i) p.prof[((size_t)blockIdx.x * 10 + warp) * 16 + i] = pc[i];
    }
    if (half == 0 && valid) {
#pragma unroll
      for (int l = 0; l < L16; ++l) if (l < L) p.Z[(size_t)l * NP + n] = z[l];
      p.Vs_w[(size_t)cur * NP + n] = cnt;
      if (p.t_cnt != nullptr) p.t_cnt[n] += n_acc;
    }
  }

  fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tbase, 512);
}

// ---- host side -------------------------------------------------------------------------------
template <int L16, bool PROF_ON, int VAR>
int32_t launch_tc(const CUtensorMap& mx, const TcArgs& a, int grid, cudaStream_t st) {
  const size_t fixed = (size_t)HID * 2 * L16 * 2 + (size_t)HID * HID * 2 + (size_t)a.FN * HID * 2 + (size_t)(a.FN + 16 + HID) * 4 +
                       (size_t)2 * NE * 8 + (12 + 2 * MAX_STAGES) * 8 + 16;
  TcArgs args = a;
  const size_t cap = 227 * 1024;
  int nstage = (int)((cap - fixed) / STAGE_BYTES);
  if (nstage > MAX_STAGES) nstage = MAX_STAGES;
  if (nstage < 2) return fail(GVN_E_UNSUPPORTED_SHAPE, "tensor-core E-step: shared memory does not fit (F=%d L=%d)", a.F, a.L);
  args.nstage = nstage;
  const size_t smem = fixed + (size_t)nstage * STAGE_BYTES;
  static size_t smem_tab[GVN_MAX_DEVICES] = {0};
  size_t& smem_set = *per_device_slot(smem_tab);                           // the attribute is sticky: set it once per size
  if (smem_set != smem) {
    cudaError_t e = cudaFuncSetAttribute(k_estep_tc<L16, PROF_ON, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(GVN_E_CUDA, "estep_tc smem attr (%zu B): %s", smem, cudaGetErrorString(e));
    smem_set = smem;
  }
  k_estep_tc<L16, PROF_ON, VAR><<<grid, NTHREADS, smem, st>>>(mx, args);
  return check_launch("k_estep_tc");
}

}  // namespace

static unsigned long long* g_prof = nullptr;
void set_profile_buffer(void* p) { g_prof = reinterpret_cast<unsigned long long*>(p); }

int32_t launch_estep_tc(const gvn_batch* b, const void* packed, int burnin, int R, float var_RW, const gvn_noise* nz,
                        const gvn_trace* tr, int precision, cudaStream_t st) {
  const bool xv_current = (precision & GVN_PREC_XV_CURRENT) != 0;
  precision &= ~GVN_PREC_XV_CURRENT;
  GVN_REQUIRE(precision == GVN_PREC_F16, GVN_E_INVALID, "precision %d is not a tensor-core mode", precision);
  GVN_REQUIRE(b->XV != nullptr && b->Vs_w != nullptr, GVN_E_INVALID, "batch.XV / batch.Vs_w is NULL");
  DecoderLayout d = decoder_layout(b->L, 0, b->F);
  TcLayout t = tc_layout(b->L, b->F);
  const unsigned char* base = reinterpret_cast<const unsigned char*>(packed);
  // per-launch constants of the chain: X2 and Vb as one bf16 pair per (f, frame)
  // (skipped when the caller vouches that gvn_mstep, which refreshes XV next to Vb, was the last writer of Vb)
  int32_t rc = GVN_OK;
  if (!xv_current) {
    const size_t n4 = (size_t)b->F * b->NP / 4;
    k_pack_xv<<<148 * 8, 256, 0, st>>>(n4, reinterpret_cast<const float4*>(b->X2), reinterpret_cast<const float4*>(b->Vb),
                                       reinterpret_cast<uint4*>(b->XV));
    if ((rc = check_launch("k_pack_xv"))) return rc;
  }
  static thread_local CUtensorMap mx;                   // encoded once per (pointer, shape)
  static thread_local const void* mx_ptr = nullptr;
  static thread_local int mx_F = 0, mx_NP = 0;
  if (mx_ptr != b->XV || mx_F != b->F || mx_NP != b->NP) {
    if ((rc = make_tile_map(&mx, b->XV, b->F, b->NP, TM))) return rc;
    mx_ptr = b->XV; mx_F = b->F; mx_NP = b->NP;
  }
  TcArgs a;
  a.F = b->F; a.FN = t.FN; a.L = b->L; a.NP = b->NP; a.burnin = burnin; a.R = R; a.nstage = 0;
  a.sd = sqrtf(var_RW);
  a.frame_utt = b->frame_utt; a.g = b->g; a.yproj = b->yproj; a.Z = b->Z; a.Vs = b->Vs; a.Vs_w = b->Vs_w;
  a.img = base + d.tc_image; a.off_w1d = t.w1d; a.off_w2 = t.w2; a.off_w3 = t.w3; a.off_b3s = t.b3s;
  a.b2 = reinterpret_cast<const float*>(packed) + d.b2;
  a.eps = nz->eps; a.u = nz->u; a.forced = nz->forced_accept; a.seed = nz->seed; a.chain = nz->chain;
  a.t_acc = tr ? tr->acc_prob : nullptr; a.t_dec = tr ? tr->accepted : nullptr;
  a.t_cnt = tr ? tr->n_accepted : nullptr; a.t_zs = tr ? tr->z_samples : nullptr;
  a.prof = g_prof;
  const int grid = (b->NP + TM - 1) / TM;
  static int var = -1;                                  // GVN_TC_VARIANT: epilogue scheduling experiments (see `stage`)
  if (var < 0) { const char* e = getenv("GVN_TC_VARIANT"); var = e ? atoi(e) & 127 : GVN_TC_DEFAULT_VARIANT; }
  if (t.L16 == 16) {
    if (a.prof != nullptr) return launch_tc<16, true, GVN_TC_DEFAULT_VARIANT>(mx, a, grid, st);
    switch (var) {
      case 22: return launch_tc<16, false, 22>(mx, a, grid, st);
      case 86: return launch_tc<16, false, 86>(mx, a, grid, st);
      case 118: return launch_tc<16, false, 118>(mx, a, grid, st);
      default: return launch_tc<16, false, GVN_TC_DEFAULT_VARIANT>(mx, a, grid, st);
    }
  }
  if (t.L16 == 32) return launch_tc<32, false, GVN_TC_DEFAULT_VARIANT>(mx, a, grid, st);
  return fail(GVN_E_UNSUPPORTED_SHAPE, "tensor-core E-step supports L <= 32 (got %d)", b->L);
}

}  // namespace gvn
