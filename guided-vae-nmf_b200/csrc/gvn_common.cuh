// Shared host/device helpers of libgvn.so (B200 / sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/gvn.h"

namespace gvn {

// ---------------------------------------------------------------------------------------
// error reporting (thread-local message, negative status)
// ---------------------------------------------------------------------------------------
char* error_buffer();
int32_t fail(int32_t code, const char* fmt, ...);
int32_t check_launch(const char* what);

#define GVN_REQUIRE(cond, code, ...) \
  do { if (!(cond)) return ::gvn::fail((code), __VA_ARGS__); } while (0)

// Function attributes (cudaFuncAttributeMaxDynamicSharedMemorySize) are per DEVICE and sticky: launchers remember the
// value they have set in a table indexed by the current device, so a process that drives several GPUs sets it on each.
constexpr int GVN_MAX_DEVICES = 64;
inline size_t* per_device_slot(size_t (&table)[GVN_MAX_DEVICES]) {
  int dev = 0;
  cudaGetDevice(&dev);
  return &table[dev & (GVN_MAX_DEVICES - 1)];
}

// ---------------------------------------------------------------------------------------
// packed decoder image: offsets (in floats) of the fp32 section, computed from the dims
// ---------------------------------------------------------------------------------------
struct DecoderLayout {
  int L, y_dim, F, FS;        // FS: row stride of W3T, F rounded up to a multiple of 128
  size_t w1zT, w1y, b1, w2T, b2, w3T, b3, end_f32;   // float offsets
  size_t tc_image;            // byte offset of the tensor-core operand image (16 B aligned)
  size_t tc_bytes;
  size_t total_bytes;
};

__host__ __device__ inline size_t round_up(size_t x, size_t m) { return (x + m - 1) / m * m; }

// column-tile order of Vs / X2t (include/gvn.h): offset of (f, n) inside one [NP/8][F][8] plane (GVN_VS_TILE = 8)
__host__ __device__ inline size_t tile_off(int f, int n, int F) {
  return ((size_t)(n / GVN_VS_TILE) * F + f) * GVN_VS_TILE + (n % GVN_VS_TILE);
}

// Tensor-core operand image (estep_tc.cu): f16 hi and lo planes of W1z (K padded to 16),
// W2 and W3 (N padded to a multiple of 16) in the canonical no-swizzle K-major UMMA layout.
struct TcLayout {
  int L16;                    // L rounded up to 16
  int FN;                     // F rounded up to 16
  size_t w1, w2, w3;          // byte offsets (relative to tc_image) of the hi planes
  size_t plane_w1, plane_w2, plane_w3;   // bytes of one plane; lo plane follows hi plane
  size_t w1d, plane_w1d;      // [128][2*L16] image of [W1 | W1] (A operand = z hi | z lo)
  size_t b1s, b3s;            // f32 vectors: (unused) / b3 * log2(e)
  size_t bytes;
};

inline TcLayout tc_layout(int L, int F) {
  TcLayout t;
  t.L16 = (int)round_up((size_t)L, 16);
  t.FN = (int)round_up((size_t)F, 16);
  t.plane_w1 = (size_t)GVN_HIDDEN * t.L16 * 2;
  t.plane_w2 = (size_t)GVN_HIDDEN * GVN_HIDDEN * 2;
  t.plane_w3 = (size_t)t.FN * GVN_HIDDEN * 2;
  size_t o = 0;
  t.w1 = o; o += 2 * t.plane_w1;
  t.w2 = o; o += 2 * t.plane_w2;
  t.w3 = o; o += 2 * t.plane_w3;
  t.plane_w1d = (size_t)GVN_HIDDEN * 2 * t.L16 * 2;
  t.w1d = o; o += t.plane_w1d;
  o = round_up(o, 16);
  t.b1s = o; o += GVN_HIDDEN * 4;
  t.b3s = o; o += (size_t)t.FN * 4;
  t.bytes = round_up(o, 16);
  return t;
}

inline DecoderLayout decoder_layout(int L, int y_dim, int F) {
  // Everything the chain kernels read comes first and does not depend on y_dim; the label
  // columns of the first layer (only read by gvn_label_projection) sit at the very end.
  DecoderLayout d;
  d.L = L; d.y_dim = y_dim; d.F = F;
  d.FS = (int)round_up((size_t)F, 128);
  size_t o = 0;
  d.w1zT = o; o += (size_t)L * GVN_HIDDEN;
  d.b1 = o;   o += GVN_HIDDEN;
  d.w2T = o;  o += (size_t)GVN_HIDDEN * GVN_HIDDEN;
  d.b2 = o;   o += GVN_HIDDEN;
  d.w3T = o;  o += (size_t)GVN_HIDDEN * d.FS;
  d.b3 = o;   o += round_up((size_t)F, 4);
  d.end_f32 = o;
  d.tc_image = round_up(o * 4, 128);
  d.tc_bytes = tc_layout(L, F).bytes;
  d.w1y = (d.tc_image + d.tc_bytes) / 4;
  d.total_bytes = d.w1y * 4 + round_up((size_t)GVN_HIDDEN * y_dim, 4) * 4;
  return d;
}

// ---------------------------------------------------------------------------------------
// Philox4x32-10 counter-based generator (throughput mode; parity mode replays a tape)
// ---------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0; key.y += W1;
  }
  return ctr;
}
// uniform in (0,1): 24 random bits + half an ulp
__device__ __forceinline__ float u01(uint32_t x) { return ((x >> 8) + 0.5f) * (1.0f / 16777216.0f); }
// One word of XV (include/gvn.h): low half = bf16(Vb) (round to nearest); high half chosen so that the WHOLE word,
// read as an f32, is the value nearest to X2 -- the low half then acts as extra mantissa bits of X2 (same 2^-9
// relative bound as a bf16 rounding) and the chain epilogue needs no mask to take X2 out of the word.
__device__ __forceinline__ uint32_t pack_xv_word(float x2, float vb) {
  uint32_t vbits = __float_as_uint(fmaxf(vb, 0.f));
  vbits += 0x7fffu + ((vbits >> 16) & 1u);                 // bf16 round to nearest even (finite, non-negative input)
  const uint32_t lo = vbits >> 16;
  const uint32_t xb = __float_as_uint(fmaxf(x2, 0.f));
  const uint32_t hi = xb > lo ? (xb - lo + 0x8000u) >> 16 : 0u;
  return (hi << 16) | lo;
}
// two standard normals from two 32-bit words (Box-Muller)
__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b) {
  float r = sqrtf(-2.0f * __logf(u01(a)));
  float s, c;
  __sincosf(6.283185307179586f * u01(b), &s, &c);
  return make_float2(r * c, r * s);
}
#endif

}  // namespace gvn
