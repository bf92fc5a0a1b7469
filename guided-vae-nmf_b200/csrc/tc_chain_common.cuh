// Pieces of the tensor-core chain kernel (estep_tc_chain.cu: one 128-frame tile per CTA): argument block, special-function wrappers,
// the XV packing pass and the TMA tensor map of XV.  Included inside an anonymous namespace of each file.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

#include "gvn_common.cuh"
#include "tc_common.cuh"

namespace gvn {
namespace {

using namespace tc;

constexpr int TM = 128;                 // frames per CTA (MMA M)
constexpr int HID = GVN_HIDDEN;
constexpr int SROWS = 32;               // frequency rows per TMA box
constexpr int MAX_STAGES = 8;
constexpr float W_SCALE_INV = 1.0f / 256.0f;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;
constexpr float SC3 = W_SCALE_INV * LOG2E;

struct TcArgs {
  int F, FN, L, NP, burnin, R, nstage;
  float sd;
  const int32_t* frame_utt;
  const float* g; const float* yproj;
  float* Z; float* Vs; float* Vs_w;
  const unsigned char* img;            // tensor-core operand image
  size_t off_w1d, off_w2, off_w3, off_b3s;
  const float* b2;
  const float* eps; const float* u; const uint8_t* forced; uint64_t seed, chain;
  float* t_acc; uint8_t* t_dec; int32_t* t_cnt; float* t_zs;
  unsigned long long* prof;            // optional [grid][10 warps][16] cycle counters (gvn_debug_profile_buffer)
};

__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_approx(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float tanh_approx(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// two hidden units per MUFU instruction: the result is needed as an f16 pair (the next layer's A operand) anyway
__device__ __forceinline__ uint32_t tanh_f16x2(float lo, float hi) {
  uint32_t x = pack_f16(lo, hi), y;
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}

__device__ __forceinline__ void st_stream(float* p, float v) { asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }

// XV[f][n] = pack_xv_word(X2, Vb) (gvn_common.cuh) for the whole batch: the chain's per-bin constants
__global__ void __launch_bounds__(256) k_pack_xv(size_t n4, const float4* __restrict__ X2, const float4* __restrict__ Vb,
                                                 uint4* __restrict__ XV) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 x = X2[i], v = Vb[i];
    uint4 o;
    o.x = pack_xv_word(x.x, v.x); o.y = pack_xv_word(x.y, v.y); o.z = pack_xv_word(x.z, v.z); o.w = pack_xv_word(x.w, v.w);
    XV[i] = o;
  }
}

// parity wait for the two single-lane service warps: suspends in hardware between polls so that
// their spinning does not take issue slots from the epilogue warps of the same scheduler
__device__ __forceinline__ void mbar_wait_idle(uint64_t* bar, uint32_t parity) {
#ifdef GVN_DEBUG_WATCHDOG
  uint32_t ok = 0;
  unsigned long long t0 = 0ull;
  for (uint32_t spins = 0; !ok; ++spins) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(20000u) : "memory");
    if (!ok) watchdog_poll(spins | 1023u, t0);       // every failed poll already slept up to ~20 us (tc_common.cuh)
  }
#else
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "GVN_WAITI_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@!p bra GVN_WAITI_%=;\n\t}"
      :: "r"(smem_u32(bar)), "r"(parity), "r"(20000u) : "memory");
#endif
}

// [F][NP] u32 array, box = [SROWS rows][box_frames frames]
int32_t make_tile_map(CUtensorMap* m, const uint32_t* base, int F, int NP, int box_frames) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (fn == nullptr) return fail(GVN_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[2] = {(cuuint64_t)NP, (cuuint64_t)F};
  cuuint64_t gstride[1] = {(cuuint64_t)NP * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_frames, (cuuint32_t)SROWS};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<uint32_t*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(GVN_E_CUDA, "cuTensorMapEncodeTiled failed with %d", (int)r);
  return GVN_OK;
}

}  // namespace
}  // namespace gvn
