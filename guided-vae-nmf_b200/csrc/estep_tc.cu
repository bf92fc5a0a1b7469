// E-step, tensor-core variant (GVN_PREC_F16): tcgen05.mma with the accumulators
// and the activation operand in TMEM, weights in shared memory.  See DESIGN.md section 5.
//
// This file currently holds: the weight packing into UMMA operand images, and
// gvn_selftest_umma -- a one-tile GEMM through exactly the instruction forms the chain kernel
// uses (A written to TMEM by tcgen05.st, B from a packed shared-memory image, D read back with
// tcgen05.ld), which pins the descriptor conventions on hardware.
#include "gvn_common.cuh"
#include "tc_common.cuh"

namespace gvn {

using namespace tc;

namespace {

constexpr float W_SCALE = 256.0f;    // weights are scaled by 2^8 so that the f16 lo plane stays normal

// [N][K] fp32 row-major (K contiguous, as nn.Linear stores weights; rows n >= N_valid or k >= K_valid
// read as zero) -> hi and lo f16 planes in the packed image layout (tc::img_offset)
__global__ void k_pack_plane(const float* __restrict__ W, int ldw, int col0, int N_valid, int K_valid, int N, int K,
                             float scale, unsigned char* __restrict__ hi, unsigned char* __restrict__ lo) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N * K; i += gridDim.x * blockDim.x) {
    int n = i / K, k = i % K;
    float w = (n < N_valid && k < K_valid) ? W[(size_t)n * ldw + col0 + k] * scale : 0.f;
    __half h = __float2half_rn(w);
    __half l = __float2half_rn(w - __half2float(h));
    size_t o = img_offset(n, k, K);
    *reinterpret_cast<__half*>(hi + o) = h;
    *reinterpret_cast<__half*>(lo + o) = l;
  }
}

// [128][2*L16] image of [W1z | W1z] (hi plane only): the A operand carries z as an f16 hi | lo pair
__global__ void k_pack_w1dup(const float* __restrict__ W1, int ldw, int L, int L16, float scale, unsigned char* __restrict__ img) {
  const int K = 2 * L16;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < GVN_HIDDEN * K; i += gridDim.x * blockDim.x) {
    const int n = i / K, k = i % K, l = k % L16;
    const float w = (l < L) ? W1[(size_t)n * ldw + l] * scale : 0.f;
    *reinterpret_cast<__half*>(img + img_offset(n, k, K)) = __float2half_rn(w);
  }
}

__global__ void k_pack_bias(const float* __restrict__ b3, int F, int FN, float* __restrict__ b3s) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < FN) b3s[i] = i < F ? b3[i] * 1.4426950408889634f : 0.f;
}

// ------------------------------------------------------------------------------------------
// self test: D[128][N] = A[128][K] * W[N][K]^T with f16 operands (optionally hi/lo split)
// variant bit0: swap the two halves when packing A into TMEM; bit1: swap LBO/SBO in the
// descriptor; bit2: use the 3-term hi/lo split.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) k_selftest_umma(const float* __restrict__ A, const float* __restrict__ W, int N,
                                                          int K, int variant, float* __restrict__ D) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  unsigned char* sHi = smem;
  unsigned char* sLo = smem + (size_t)N * K * 2;
  for (int i = tid; i < N * K; i += 128) {
    int n = i / K, k = i % K;
    float w = W[(size_t)n * K + k];
    __half h = __float2half_rn(w);
    __half l = __float2half_rn(w - __half2float(h));
    *reinterpret_cast<__half*>(sHi + img_offset(n, k, K)) = h;
    *reinterpret_cast<__half*>(sLo + img_offset(n, k, K)) = l;
  }
  if (tid == 0) { mbar_init(&bar, 1); mbar_init_fence(); }
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  fence_proxy_async();          // generic-proxy smem writes -> visible to the tensor core (async proxy)
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tbase = tmem_slot;
  const uint32_t lane_base = tbase + ((uint32_t)(warp * 32) << 16);
  const uint32_t A_HI = 256, A_LO = 320;   // column offsets of the A operand planes
  // A row of this thread -> packed f16 pairs in TMEM
  for (int k0 = 0; k0 < K; k0 += 16) {
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float a0 = A[(size_t)tid * K + k0 + 2 * j], a1 = A[(size_t)tid * K + k0 + 2 * j + 1];
      __half h0 = __float2half_rn(a0), h1 = __float2half_rn(a1);
      float l0 = a0 - __half2float(h0), l1 = a1 - __half2float(h1);
      if (variant & 1) { hi[j] = pack_f16(__half2float(h1), __half2float(h0)); lo[j] = pack_f16(l1, l0); }
      else             { hi[j] = pack_f16(__half2float(h0), __half2float(h1)); lo[j] = pack_f16(l0, l1); }
    }
    tmem_st8(lane_base + A_HI + k0 / 2, hi);
    tmem_st8(lane_base + A_LO + k0 / 2, lo);
  }
  tmem_st_wait();
  fence_before();
  __syncthreads();
  if (tid == 0) {
    fence_after();
    const uint32_t idesc = idesc_f16(128, N);
    const uint32_t lbo = (variant & 2) ? img_sbo(K) : img_lbo(), sbo = (variant & 2) ? img_lbo() : img_sbo(K);
    uint32_t acc = 0;
    const int terms = (variant & 4) ? 3 : 1;
    for (int t = 0; t < terms; ++t) {
      // t=0: hi*hi, t=1: hi*lo, t=2: lo*hi
      const uint32_t a_col = (t == 2) ? A_LO : A_HI;
      const unsigned char* bimg = (t == 1) ? sLo : sHi;
      for (int k0 = 0; k0 < K; k0 += 16) {
        uint64_t bd = smem_desc(smem_u32(bimg) + (uint32_t)(k0 / 8) * 128, lbo, sbo);
        mma_ts(tbase, tbase + a_col + k0 / 2, bd, idesc, acc);
        acc = 1;
      }
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  fence_after();
  for (int c0 = 0; c0 < N; c0 += 8) {
    uint32_t r[8];
    tmem_ld8(lane_base + c0, r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 8; ++j) D[(size_t)tid * N + c0 + j] = __uint_as_float(r[j]);
  }
  fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 512);
}

}  // namespace

int32_t launch_pack_tc(const float* W1, const float* W2, const float* W3, const float* b3, int L, int y_dim, int F,
                       unsigned char* image, cudaStream_t st) {
  TcLayout t = tc_layout(L, F);
  const int H = GVN_HIDDEN;
  k_pack_plane<<<32, 256, 0, st>>>(W1, L + y_dim, 0, H, L, H, t.L16, W_SCALE, image + t.w1, image + t.w1 + t.plane_w1);
  k_pack_plane<<<64, 256, 0, st>>>(W2, H, 0, H, H, H, H, W_SCALE, image + t.w2, image + t.w2 + t.plane_w2);
  k_pack_plane<<<148, 256, 0, st>>>(W3, H, 0, F, H, t.FN, H, W_SCALE, image + t.w3, image + t.w3 + t.plane_w3);
  k_pack_w1dup<<<32, 256, 0, st>>>(W1, L + y_dim, L, t.L16, W_SCALE, image + t.w1d);
  k_pack_bias<<<(t.FN + 255) / 256, 256, 0, st>>>(b3, F, t.FN, reinterpret_cast<float*>(image + t.b3s));
  return check_launch("k_pack_tc");
}

// launch_estep_tc: see estep_tc_chain.cu

int32_t launch_selftest_umma(const float* A, const float* W, int N, int K, int variant, float* D, cudaStream_t st) {
  GVN_REQUIRE(N % 16 == 0 && N >= 16 && N <= 256 && K % 16 == 0 && K >= 16 && K <= 128, GVN_E_INVALID,
              "selftest shape N=%d K=%d", N, K);
  size_t smem = (size_t)N * K * 2 * 2;
  cudaError_t e = cudaFuncSetAttribute(k_selftest_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return fail(GVN_E_CUDA, "selftest smem attr: %s", cudaGetErrorString(e));
  k_selftest_umma<<<1, 128, smem, st>>>(A, W, N, K, variant, D);
  return check_launch("k_selftest_umma");
}

}  // namespace gvn
