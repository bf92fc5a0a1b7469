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
//   * weights sit in shared memory for the whole chain as UMMA operand images (packed once by
//     gvn_pack_decoder; f16 because activations are in [-1,1] and weights O(0.05): 11 mantissa
//     bits = TF32 precision at twice the MMA rate and half the footprint);
//   * the output layer (F=513 -> 4 chunks of 128 columns + one of 16) is double-buffered in TMEM:
//     the MMA of chunk c+1 runs while the epilogue consumes chunk c;
//   * X2 and Vb stream through a TMA ring ([16 frequency rows][128 frames] boxes of the
//     frame-minor arrays), produced by a dedicated warp.
// Warp roles: 0-7 epilogue (warp w and w+4 share TMEM lane quarter w%4 and split the columns),
// 8 = MMA issuer (one elected lane), 9 = TMA producer (one elected lane).
#include <cuda.h>

#include "gvn_common.cuh"
#include "tc_common.cuh"

namespace gvn {

using namespace tc;

namespace {

constexpr int TM = 128;                 // frames per tile (MMA M)
constexpr int HID = GVN_HIDDEN;
constexpr int NE = 256;                 // epilogue threads
constexpr int NTHREADS = 320;
constexpr int SUB = 16;                 // frequency rows per TMA box
constexpr int STAGE_BYTES = 2 * SUB * TM * 4;
constexpr float W_SCALE_INV = 1.0f / 256.0f;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

// TMEM column map (512 columns allocated)
constexpr uint32_t COL_ACC0 = 0, COL_ACC1 = 128, COL_A_HI = 256, COL_A_LO = 320, COL_Z_HI = 384, COL_Z_LO = 416;

struct TcArgs {
  int F, FN, L, NP, burnin, R, nstage;
  float sd;
  const int32_t* frame_utt;
  const float* g; const float* yproj;
  float* Z; float* Vs;
  const unsigned char* img;            // tensor-core operand image
  size_t off_w1, off_w2, off_w3, off_b3s;
  const float* b2;
  const float* eps; const float* u; const uint8_t* forced; uint64_t seed, chain;
  float* t_acc; uint8_t* t_dec; int32_t* t_cnt; float* t_zs;
};

__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_approx(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float tanh_approx(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

__device__ __forceinline__ void bar_epilogue() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

enum { MODE_INIT = 0, MODE_PROP = 1, MODE_WRITE = 2 };

template <int L16>
__global__ void __launch_bounds__(NTHREADS, 1) k_estep_tc(const __grid_constant__ CUtensorMap tm_x2,
                                                          const __grid_constant__ CUtensorMap tm_vb, TcArgs p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int F = p.F, FN = p.FN, NP = p.NP, L = p.L;
  const int NCH = (FN + 127) / 128;                       // output-layer chunks (last one may be 16 wide)
  // ---- shared memory carve-up ----
  unsigned char* sW1 = smem;                              // [128][L16] f16 image
  unsigned char* sW2 = sW1 + HID * L16 * 2;               // [128][128]
  unsigned char* sW3 = sW2 + HID * HID * 2;               // [FN][128]
  unsigned char* sRing = sW3 + (size_t)FN * HID * 2;      // nstage x {X2 [16][128] f32, Vb [16][128] f32}
  float* sB3 = reinterpret_cast<float*>(sRing + (size_t)p.nstage * STAGE_BYTES);   // [FN]  b3*log2e
  float* sB2 = sB3 + FN;                                  // [128]
  double* sPart = reinterpret_cast<double*>(sB2 + HID);   // [256]
  int* sAcc = reinterpret_cast<int*>(sPart + NE);         // [128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sAcc + TM);
  uint64_t* bar_z = bars + 0;                             // E(0-3) -> M : Z operand in TMEM
  uint64_t* bar_h = bars + 1;                             // E -> M : hidden activations in TMEM
  uint64_t* bar_d = bars + 2;                             // [2] M -> E : accumulator buffer full
  uint64_t* bar_free = bars + 4;                          // [2] E -> M : accumulator buffer drained
  uint64_t* bar_full = bars + 6;                          // [nstage] TMA -> E
  uint64_t* bar_empty = bars + 6 + 8;                     // [nstage] E -> P
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6 + 16);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n0 = blockIdx.x * TM;

  // ---- one-time setup: weights -> smem, barriers, TMEM ----
  {
    const uint4* src; uint4* dst; int n16;
    src = reinterpret_cast<const uint4*>(p.img + p.off_w1); dst = reinterpret_cast<uint4*>(sW1); n16 = HID * L16 * 2 / 16;
    for (int i = tid; i < n16; i += NTHREADS) dst[i] = src[i];
    src = reinterpret_cast<const uint4*>(p.img + p.off_w2); dst = reinterpret_cast<uint4*>(sW2); n16 = HID * HID * 2 / 16;
    for (int i = tid; i < n16; i += NTHREADS) dst[i] = src[i];
    src = reinterpret_cast<const uint4*>(p.img + p.off_w3); dst = reinterpret_cast<uint4*>(sW3); n16 = FN * HID * 2 / 16;
    for (int i = tid; i < n16; i += NTHREADS) dst[i] = src[i];
    const float* b3s = reinterpret_cast<const float*>(p.img + p.off_b3s);
    for (int i = tid; i < FN; i += NTHREADS) sB3[i] = b3s[i];
    for (int i = tid; i < HID; i += NTHREADS) sB2[i] = p.b2[i];
  }
  if (tid == 0) {
    mbar_init(bar_z, 128);
    mbar_init(bar_h, NE);
    for (int i = 0; i < 2; ++i) { mbar_init(bar_d + i, 1); mbar_init(bar_free + i, NE); }
    for (int i = 0; i < p.nstage; ++i) { mbar_init(bar_full + i, 1); mbar_init(bar_empty + i, NE); }
    mbar_init_fence();
  }
  if (warp == 8) tmem_alloc(tmem_slot, 512);
  fence_proxy_async();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tbase = *tmem_slot;

  const int n_steps = p.burnin + p.R;
  // number of decoder passes with X2/Vb streaming and the pass schedule are the same for all roles:
  //   INIT, then per step m: PROP, and WRITE after the step with m == burnin.

  if (warp == 9) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      uint32_t it = 0;
      const int n_load_passes = 1 + n_steps;
      for (int ps = 0; ps < n_load_passes; ++ps) {
        for (int f0 = 0; f0 < FN; f0 += SUB, ++it) {
          const uint32_t slot = it % p.nstage, par = ((it / p.nstage) & 1) ^ 1;
          mbar_wait(bar_empty + slot, par);
          mbar_expect_tx(bar_full + slot, STAGE_BYTES);
          unsigned char* dst = sRing + (size_t)slot * STAGE_BYTES;
          tma_load_2d(dst, &tm_x2, n0, f0, bar_full + slot);
          tma_load_2d(dst + SUB * TM * 4, &tm_vb, n0, f0, bar_full + slot);
        }
      }
    }
  } else if (warp == 8) {
    // =============================== MMA issuer ===============================
    if (lane == 0) {
      uint32_t n_z = 0, n_h = 0, n_free[2] = {0, 0};
      const uint32_t sbo1 = img_sbo(L16), sbo = img_sbo(HID), lbo = img_lbo();
      const int n_pass = 1 + n_steps + (p.R > 0 ? 1 : 0);
      for (int ps = 0; ps < n_pass; ++ps) {
        // layer 1: ACC0[:, 0:128] = Z(128 x L16) * W1^T
        mbar_wait(bar_free + 0, (n_free[0]++ & 1) ^ 1);
        mbar_wait(bar_z, n_z++ & 1);
        fence_after();
        for (int k0 = 0; k0 < L16; k0 += 16)
          mma_ts(tbase + COL_ACC0, tbase + COL_Z_HI + k0 / 2, smem_desc(smem_u32(sW1) + (k0 / 8) * 128, lbo, sbo1),
                 idesc_f16(TM, HID), k0 > 0);
        mma_commit(bar_d + 0);
        // layer 2: ACC1[:, 0:128] = H1(128 x 128) * W2^T
        mbar_wait(bar_free + 1, (n_free[1]++ & 1) ^ 1);
        mbar_wait(bar_h, n_h++ & 1);
        fence_after();
        for (int k0 = 0; k0 < HID; k0 += 16)
          mma_ts(tbase + COL_ACC1, tbase + COL_A_HI + k0 / 2, smem_desc(smem_u32(sW2) + (k0 / 8) * 128, lbo, sbo),
                 idesc_f16(TM, HID), k0 > 0);
        mma_commit(bar_d + 1);
        // layer 3: chunks of the output features, alternating accumulator buffers
        mbar_wait(bar_h, n_h++ & 1);
        fence_after();
        for (int c = 0; c < NCH; ++c) {
          const int buf = c & 1, ncol = min(128, FN - c * 128);
          mbar_wait(bar_free + buf, (n_free[buf]++ & 1) ^ 1);
          fence_after();
          const uint32_t b0 = smem_u32(sW3) + (uint32_t)(c * 16) * sbo;
          for (int k0 = 0; k0 < HID; k0 += 16)
            mma_ts(tbase + (buf ? COL_ACC1 : COL_ACC0), tbase + COL_A_HI + k0 / 2, smem_desc(b0 + (k0 / 8) * 128, lbo, sbo),
                   idesc_f16(TM, ncol), k0 > 0);
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
    uint32_t n_d[2] = {0, 0}, it = 0;
    float z[L16], zp[L16];
    double Ct = 0.0;
    int n_acc = 0;
    if (half == 0) {
#pragma unroll
      for (int l = 0; l < L16; ++l) z[l] = (l < L && in_range) ? p.Z[(size_t)l * NP + n] : 0.f;
    }

    auto put_z = [&](const float (&zz)[L16]) {       // warps 0-3: Z operand -> TMEM, signal the issuer
#pragma unroll
      for (int k0 = 0; k0 < L16; k0 += 16) {
        uint32_t hi[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) hi[j] = pack_f16(zz[k0 + 2 * j], zz[k0 + 2 * j + 1]);
        tmem_st8(tlane + COL_Z_HI + k0 / 2, hi);
      }
      tmem_st_wait();
      fence_before();
      mbar_arrive(bar_z);
    };

    // one decoder evaluation; returns the energy sum_f log Vx + X2/Vx of this thread's row (half 0)
    auto decode = [&](int mode, float* vs_out) -> double {
      // ---- hidden layers: 64 columns per thread ----
#pragma unroll 1
      for (int layer = 0; layer < 2; ++layer) {
        mbar_wait(bar_d + layer, n_d[layer]++ & 1);
        fence_after();
        const uint32_t acc_col = (layer ? COL_ACC1 : COL_ACC0) + 64 * half;
#pragma unroll 1
        for (int c0 = 0; c0 < 64; c0 += 16) {
          uint32_t r[16];
          tmem_ld16(tlane + acc_col + c0, r);
          tmem_ld_wait();
          uint32_t hi[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int j0 = 64 * half + c0 + 2 * j;
            float b0, b1;
            if (layer == 0) {
              b0 = in_range ? p.yproj[(size_t)j0 * NP + n] : 0.f;
              b1 = in_range ? p.yproj[(size_t)(j0 + 1) * NP + n] : 0.f;
            } else {
              b0 = sB2[j0]; b1 = sB2[j0 + 1];
            }
            float h0 = tanh_approx(fmaf(__uint_as_float(r[2 * j]), W_SCALE_INV, b0));
            float h1 = tanh_approx(fmaf(__uint_as_float(r[2 * j + 1]), W_SCALE_INV, b1));
            hi[j] = pack_f16(h0, h1);
          }
          tmem_st8(tlane + COL_A_HI + (64 * half + c0) / 2, hi);
        }
        tmem_st_wait();
        fence_before();
        mbar_arrive(bar_free + layer);
        mbar_arrive(bar_h);
      }
      // ---- output layer: chunks, fused epilogue ----
      double dl = 0.0, dr = 0.0;
#pragma unroll 1
      for (int c = 0; c < NCH; ++c) {
        const int buf = c & 1, ncol = min(128, FN - c * 128);
        mbar_wait(bar_d + buf, n_d[buf]++ & 1);
        fence_after();
        const uint32_t acc_col = buf ? COL_ACC1 : COL_ACC0;
#pragma unroll 1
        for (int s0 = 0; s0 < ncol; s0 += SUB) {
          const int fbase = c * 128 + s0 + 8 * half;
          uint32_t r[8];
          tmem_ld8(tlane + acc_col + s0 + 8 * half, r);
          const float* sx = nullptr;
          uint32_t slot = 0;
          if (mode != MODE_WRITE) {
            slot = it % p.nstage;
            mbar_wait(bar_full + slot, (it / p.nstage) & 1);
            ++it;
            sx = reinterpret_cast<const float*>(sRing + (size_t)slot * STAGE_BYTES) + (8 * half) * TM + row;
          }
          tmem_ld_wait();
          float ls = 0.f, rs = 0.f;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int f = fbase + j;
            if (f < F) {
              const float vs = ex2_approx(fmaf(__uint_as_float(r[j]), W_SCALE_INV * LOG2E, sB3[f]));
              if (mode != MODE_WRITE) {
                const float x2 = sx[j * TM], vb = sx[(SUB + j) * TM];
                const float vx = fmaf(g, vs, vb);
                ls += lg2_approx(vx);
                rs = fmaf(x2, rcp_approx(vx), rs);
              }
              if (vs_out != nullptr && valid) vs_out[(size_t)f * NP + n] = vs;
            }
          }
          dl += (double)ls;
          dr += (double)rs;
          if (mode != MODE_WRITE) mbar_arrive(bar_empty + slot);
        }
        fence_before();
        mbar_arrive(bar_free + buf);
      }
      if (mode == MODE_WRITE) return 0.0;
      sPart[tid] = dl * (double)LN2 + dr;
      bar_epilogue();
      const double e = sPart[row] + sPart[row + 128];
      bar_epilogue();
      return e;
    };

    // ---- chain ----
    if (half == 0) put_z(z);
    Ct = decode(MODE_INIT, nullptr);

#pragma unroll 1
    for (int m = 0; m < n_steps; ++m) {
      float prior = 0.f;
      if (half == 0) {
        // proposal Z' = Z + sd * eps  (mcem.py:257)
        if (p.eps != nullptr) {
#pragma unroll
          for (int l = 0; l < L16; ++l) {
            float e = (l < L && in_range) ? p.eps[((size_t)m * L + l) * NP + n] : 0.f;
            zp[l] = z[l] + p.sd * e;
          }
        } else {
#pragma unroll
          for (int lq = 0; lq < L16 / 4; ++lq) {
            uint4 rr = philox4x32_10(make_uint4((uint32_t)n, (uint32_t)m, (uint32_t)lq, (uint32_t)p.chain),
                                     make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32)));
            float2 a = box_muller(rr.x, rr.y), b = box_muller(rr.z, rr.w);
            float e[4] = {a.x, a.y, b.x, b.y};
#pragma unroll
            for (int i = 0; i < 4; ++i) { const int l = 4 * lq + i; zp[l] = (l < L) ? z[l] + p.sd * e[i] : 0.f; }
          }
        }
#pragma unroll
        for (int l = 0; l < L16; ++l) prior += z[l] * z[l] - zp[l] * zp[l];
        put_z(zp);
      }
      const int r = m - p.burnin;
      float* spec = (r >= 1) ? p.Vs + (size_t)r * F * NP : nullptr;
      const double Cp = decode(MODE_PROP, spec);

      // accept / reject (mcem.py:266-280)
      if (half == 0) {
        const float acc_prob = (float)(Ct - Cp) + 0.5f * prior;
        int ok = 0;
        if (valid) {
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
        if (ok) {
          Ct = Cp; ++n_acc;
#pragma unroll
          for (int l = 0; l < L16; ++l) z[l] = zp[l];
        }
        if (r >= 0 && p.t_zs != nullptr && valid) {
#pragma unroll
          for (int l = 0; l < L16; ++l) if (l < L) p.t_zs[((size_t)r * L + l) * NP + n] = z[l];
        }
        if (r >= 1) sAcc[row] = ok;
      }
      // emit the kept sample (mcem.py:286-289 + compute_Vs)
      if (r == 0) {
        if (half == 0) put_z(z);
        decode(MODE_WRITE, p.Vs);
      } else if (r >= 1) {
        bar_epilogue();
        if (valid && !sAcc[row]) {                 // rejected: the sample repeats the previous one
          const float* prev = p.Vs + (size_t)(r - 1) * F * NP;
          for (int f0 = 8 * half; f0 < F; f0 += 16) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int f = f0 + j;
              if (f < F) spec[(size_t)f * NP + n] = prev[(size_t)f * NP + n];
            }
          }
        }
        bar_epilogue();
      }
    }
    if (half == 0 && valid) {
#pragma unroll
      for (int l = 0; l < L16; ++l) if (l < L) p.Z[(size_t)l * NP + n] = z[l];
      if (p.t_cnt != nullptr) p.t_cnt[n] += n_acc;
    }
  }

  fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tbase, 512);
}

// ---- host side -------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

// [F][NP] f32 array, box = [SUB rows][128 frames]
int32_t make_tile_map(CUtensorMap* m, const float* base, int F, int NP) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (fn == nullptr) return fail(GVN_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[2] = {(cuuint64_t)NP, (cuuint64_t)F};
  cuuint64_t gstride[1] = {(cuuint64_t)NP * 4};
  cuuint32_t box[2] = {(cuuint32_t)TM, (cuuint32_t)SUB};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(GVN_E_CUDA, "cuTensorMapEncodeTiled failed with %d", (int)r);
  return GVN_OK;
}

template <int L16>
int32_t launch_tc(const CUtensorMap& mx, const CUtensorMap& mv, const TcArgs& a, int grid, cudaStream_t st) {
  const size_t fixed = (size_t)HID * L16 * 2 + (size_t)HID * HID * 2 + (size_t)a.FN * HID * 2 + (size_t)(a.FN + HID) * 4 +
                       (size_t)NE * 8 + TM * 4 + (6 + 16) * 8 + 16;
  TcArgs args = a;
  const size_t cap = 227 * 1024;
  int nstage = (int)((cap - fixed) / STAGE_BYTES);
  if (nstage > 8) nstage = 8;
  if (nstage < 2) return fail(GVN_E_UNSUPPORTED_SHAPE, "tensor-core E-step: shared memory does not fit (F=%d L=%d)", a.F, a.L);
  args.nstage = nstage;
  const size_t smem = fixed + (size_t)nstage * STAGE_BYTES;
  cudaError_t e = cudaFuncSetAttribute(k_estep_tc<L16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return fail(GVN_E_CUDA, "estep_tc smem attr (%zu B): %s", smem, cudaGetErrorString(e));
  k_estep_tc<L16><<<grid, NTHREADS, smem, st>>>(mx, mv, args);
  return check_launch("k_estep_tc");
}

}  // namespace

int32_t launch_estep_tc(const gvn_batch* b, const void* packed, int burnin, int R, float var_RW, const gvn_noise* nz,
                        const gvn_trace* tr, int precision, cudaStream_t st) {
  if (precision != GVN_PREC_F16)
    return fail(GVN_E_UNSUPPORTED_SHAPE, "precision %d: the hi/lo-split tensor-core chain is not built yet", precision);
  DecoderLayout d = decoder_layout(b->L, 0, b->F);
  TcLayout t = tc_layout(b->L, b->F);
  const unsigned char* base = reinterpret_cast<const unsigned char*>(packed);
  CUtensorMap mx, mv;
  int32_t rc = make_tile_map(&mx, b->X2, b->F, b->NP);
  if (rc) return rc;
  if ((rc = make_tile_map(&mv, b->Vb, b->F, b->NP))) return rc;
  TcArgs a;
  a.F = b->F; a.FN = t.FN; a.L = b->L; a.NP = b->NP; a.burnin = burnin; a.R = R; a.nstage = 0;
  a.sd = sqrtf(var_RW);
  a.frame_utt = b->frame_utt; a.g = b->g; a.yproj = b->yproj; a.Z = b->Z; a.Vs = b->Vs;
  a.img = base + d.tc_image; a.off_w1 = t.w1; a.off_w2 = t.w2; a.off_w3 = t.w3; a.off_b3s = t.b3s;
  a.b2 = reinterpret_cast<const float*>(packed) + d.b2;
  a.eps = nz->eps; a.u = nz->u; a.forced = nz->forced_accept; a.seed = nz->seed; a.chain = nz->chain;
  a.t_acc = tr ? tr->acc_prob : nullptr; a.t_dec = tr ? tr->accepted : nullptr;
  a.t_cnt = tr ? tr->n_accepted : nullptr; a.t_zs = tr ? tr->z_samples : nullptr;
  const int grid = (b->NP + TM - 1) / TM;
  if (t.L16 == 16) return launch_tc<16>(mx, mv, a, grid, st);
  if (t.L16 == 32) return launch_tc<32>(mx, mv, a, grid, st);
  return fail(GVN_E_UNSUPPORTED_SHAPE, "tensor-core E-step supports L <= 32 (got %d)", b->L);
}

}  // namespace gvn
