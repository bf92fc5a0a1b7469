// Oracle guide labels from the clean-speech STFT on the device: replaces clean_speech_IBM / clean_speech_VAD
// (reference python/processing/target.py:7-27, :29-50; call sites scripts/evaluate_M2_ibm.py:132-134).
//
//   power = |S conj(S)|            (VAD: summed over frequency per frame)
//   sorted = sort(power, descending);  lorenz = cumsum(sorted) / sum(sorted)
//   threshold = sorted[last k with lorenz[k] < quantile_fraction];   label = power > threshold
//
// The labels are discrete, so the device path reproduces the ARITHMETIC ORDER of the numpy statements, per utterance:
//   * power as numpy's complex64 product gives it on an FMA machine: fma(re, re, fl(im * im));
//   * the VAD row sum over frequency is sequential in float32 (numpy reduces a C-ordered array along axis 0 row by row);
//   * np.sum is numpy's pairwise summation (blocks of 128 with 8 accumulators, halves rounded down to a multiple of 8),
//     np.cumsum is sequential -- both restated below and run by ONE thread per utterance (a quarter of a million
//     dependent additions, ~0.3 ms, once per utterance, all utterances in parallel);
//   * the sort is a stable LSD radix sort of the float bit patterns (powers are >= 0, so the patterns order like the
//     values), one CTA per utterance, 8-bit digits, per-warp histograms.
#include "gvn_common.cuh"

namespace gvn {

namespace {

constexpr int LT = 1024;               // threads per CTA of the sort
constexpr int LW = LT / 32;

__device__ __forceinline__ float power_of(float2 s) { return __fmaf_rn(s.x, s.x, __fmul_rn(s.y, s.y)); }

// keys[b][i] = float bits of the power: IBM i = f * N_b + n (any order will do), VAD i = n
__global__ void __launch_bounds__(256) k_label_keys(int F, int NP, int vad, int from_power, const int32_t* __restrict__ frame_utt,
                                                    const int32_t* __restrict__ frame_off, const int32_t* __restrict__ n_frames,
                                                    const float* __restrict__ S, const size_t* __restrict__ seg_off,
                                                    uint32_t* __restrict__ keys, float* __restrict__ pw_out) {
  const int gn = blockIdx.x * 32 + (threadIdx.x & 31), fw = threadIdx.x >> 5;
  if (gn >= NP) return;
  const int b = frame_utt[gn];
  if (b < 0) return;
  const int nn = gn - frame_off[b], N = n_frames[b];
  uint32_t* k = keys + seg_off[b];
  auto pw = [&](int f) -> float {
    const size_t o = (size_t)f * NP + gn;
    return from_power ? S[o] : power_of(reinterpret_cast<const float2*>(S)[o]);
  };
  if (vad) {
    if (fw != 0) return;
    float acc = 0.f;
    for (int f = 0; f < F; ++f) acc = __fadd_rn(acc, pw(f));
    k[nn] = __float_as_uint(acc);
    pw_out[gn] = acc;
  } else {
    for (int f = fw; f < F; f += 8) {
      const float v = pw(f);
      k[(size_t)f * N + nn] = __float_as_uint(v);
      pw_out[(size_t)f * NP + gn] = v;
    }
  }
}

// stable LSD radix sort (descending) of one utterance's keys, in place via the second buffer; one CTA per utterance
__global__ void __launch_bounds__(LT) k_label_sort(int F, int vad, const int32_t* __restrict__ n_frames, const size_t* __restrict__ seg_off,
                                                   uint32_t* __restrict__ keys, uint32_t* __restrict__ tmp) {
  __shared__ uint32_t hist[LW][256];
  __shared__ uint32_t dig_base[256];
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const size_t n = (size_t)n_frames[b] * (vad ? 1 : F);
  uint32_t* src = keys + seg_off[b];
  uint32_t* dst = tmp + seg_off[b];
  // contiguous sub-range of this warp, a multiple of 32 long (except the last warp's tail)
  const size_t per = ((n + LW - 1) / LW + 31) / 32 * 32;
  const size_t lo = min(n, (size_t)warp * per), hi = min(n, lo + per);
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 8 * pass;
    for (int i = lane; i < 256; i += 32) hist[warp][i] = 0;
    __syncwarp();
    for (size_t i = lo + lane; i < hi; i += 32) atomicAdd(&hist[warp][255 - ((src[i] >> shift) & 255)], 1u);   // inverted digit: descending
    __syncthreads();
    if (tid < 256) {                                        // total of digit tid over the warps
      uint32_t t = 0;
      for (int w = 0; w < LW; ++w) t += hist[w][tid];
      dig_base[tid] = t;
    }
    __syncthreads();
    if (tid == 0) {                                         // exclusive scan over the 256 digits
      uint32_t run = 0;
      for (int d = 0; d < 256; ++d) { const uint32_t t = dig_base[d]; dig_base[d] = run; run += t; }
    }
    __syncthreads();
    if (tid < 256) {                                        // start of (digit, warp): digits first, then warps in order (stability)
      uint32_t run = dig_base[tid];
      for (int w = 0; w < LW; ++w) { const uint32_t t = hist[w][tid]; hist[w][tid] = run; run += t; }
    }
    __syncthreads();
    for (size_t i0 = lo; i0 < hi; i0 += 32) {               // this warp's keys in order
      const size_t i = i0 + lane;
      const bool act = i < hi;
      const uint32_t key = act ? src[i] : 0u;
      const uint32_t d = act ? 255 - ((key >> shift) & 255) : 256 + lane;   // inactive lanes match nobody
      const uint32_t peers = __match_any_sync(0xffffffffu, d);
      const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
      uint32_t base = 0;
      if (act) base = hist[warp][d];
      __syncwarp();
      if (act && rank == 0) hist[warp][d] = base + __popc(peers);
      __syncwarp();
      if (act) dst[base + rank] = key;
    }
    __syncthreads();
    uint32_t* t = src; src = dst; dst = t;                  // four passes: the result ends in `keys`
  }
}

// numpy's pairwise summation (numpy/_core/src/umath/loops_utils.h.src: pairwise_sum), float32, over a[0..n) in global
// memory, by one warp.  The recursion (halves rounded down to a multiple of 8, leaves of <= 128 values summed with 8
// interleaved accumulators) is an explicit stack that every lane walks in lockstep.  The leaf sums do not depend on one
// another: the walk runs twice -- first it only collects the leaves, 32 at a time, and each lane sums one of them
// (into `leaf`, global scratch); then it combines the stored leaf sums in the reference's order.
constexpr int TCH = 2048;
struct PwFrame { size_t off, len; float left; int stage; };

__device__ __forceinline__ float np_leaf_sum(const float* __restrict__ p, size_t len) {
  if (len < 8) {
    float res = 0.f;
    for (size_t i = 0; i < len; ++i) res = __fadd_rn(res, p[i]);
    return res;
  }
  float r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = p[j];
  size_t i = 8;
  for (; i < len - (len % 8); i += 8) {
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], p[i + j]);
  }
  float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])), __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
  for (; i < len; ++i) res = __fadd_rn(res, p[i]);
  return res;
}

__device__ float np_pairwise_sum(const float* __restrict__ a, size_t n, float* __restrict__ leaf, int lane) {
  float ret = 0.f;
  for (int phase = 0; phase < 2; ++phase) {
    PwFrame st[48];
    int sp = 0;
    st[0] = {0, n, 0.f, 0};
    size_t n_leaf = 0;
    size_t my_off = 0, my_len = 0;                          // phase 0: the leaf this lane sums in the current group of 32
    while (sp >= 0) {
      PwFrame& f = st[sp];
      if (f.len <= 128) {
        if (phase == 0) {
          if ((int)(n_leaf & 31) == lane) { my_off = f.off; my_len = f.len; }
          if ((n_leaf & 31) == 31) { leaf[n_leaf - 31 + lane] = np_leaf_sum(a + my_off, my_len); my_len = 0; }
        } else {
          ret = leaf[n_leaf];
        }
        ++n_leaf;
        --sp;
        continue;
      }
      size_t n2 = f.len / 2;
      n2 -= n2 % 8;
      if (f.stage == 0) {                                   // descend into the left half
        f.stage = 1;
        st[sp + 1] = {f.off, n2, 0.f, 0};
        ++sp;
      } else if (f.stage == 1) {                            // left half done: keep it, descend into the right half
        f.left = ret;
        f.stage = 2;
        st[sp + 1] = {f.off + n2, f.len - n2, 0.f, 0};
        ++sp;
      } else {
        ret = __fadd_rn(f.left, ret);
        --sp;
      }
    }
    if (phase == 0) {
      if ((n_leaf & 31) != 0 && lane < (int)(n_leaf & 31)) leaf[(n_leaf & ~(size_t)31) + lane] = np_leaf_sum(a + my_off, my_len);
      __syncwarp();
      __threadfence_block();
    }
  }
  return ret;
}

// threshold of one utterance: the last sorted value whose Lorenz share is below the quantile (target.py:19-21).
// One warp per utterance: lane 0 carries the sequential float32 cumsum, the warp stages the sorted values through shared
// memory in coalesced chunks.  The shares fl(cumsum / total) never decrease, so the first failure ends the prefix.
__global__ void __launch_bounds__(32) k_label_threshold(int F, int vad, float qf, const int32_t* __restrict__ n_frames,
                                                        const size_t* __restrict__ seg_off, const uint32_t* __restrict__ keys,
                                                        uint32_t* __restrict__ tmp, float* __restrict__ thr) {
  __shared__ __align__(16) float buf[TCH];
  const int b = blockIdx.x, lane = threadIdx.x;
  const size_t n = (size_t)n_frames[b] * (vad ? 1 : F);
  const float* a = reinterpret_cast<const float*>(keys + seg_off[b]);
  const float total = np_pairwise_sum(a, n, reinterpret_cast<float*>(tmp + seg_off[b]), lane);   // the sort's second buffer is free now
  // The share test fl(acc / total) < qf is monotone in acc: the largest float `amax` that passes it is found once by
  // bisection on the bit pattern (31 exact divisions), and the scan itself only compares acc <= amax -- the same
  // decisions as the reference's division per element, without a division per element.
  float amax = __int_as_float(0x7fc00000);
  if (__fdiv_rn(0.f, total) < qf) {
    uint32_t lo = 0u, hi = 0x7f800000u;                     // passes | fails (inf / total is not below qf <= 1)
    while (hi - lo > 1u) {
      const uint32_t mid = lo + (hi - lo) / 2u;
      if (__fdiv_rn(__uint_as_float(mid), total) < qf) lo = mid; else hi = mid;
    }
    amax = __uint_as_float(lo);
  }
  float acc = 0.f, t = __int_as_float(0x7fc00000);          // no share below the quantile: numpy raises IndexError; here nothing is flagged
  int done = 0;
  for (size_t c0 = 0; c0 < n && !done; c0 += TCH) {
    const int m = (int)min((size_t)TCH, n - c0);
    __syncwarp();
    for (int i = lane; i < m; i += 32) buf[i] = a[c0 + i];
    __syncwarp();
    if (lane == 0) {
      int i = 0;
      // groups of 8 with two 16-byte loads and ONE test: the running sums only grow, so a group whose last sum passes
      // lies entirely below the quantile (same additions in the same order as the element-wise loop)
      float4 u = make_float4(0.f, 0.f, 0.f, 0.f), v = u;
      if (m >= 8) { u = *reinterpret_cast<const float4*>(buf); v = *reinterpret_cast<const float4*>(buf + 4); }
      for (; i + 8 <= m; i += 8) {
        const float4 cu = u, cv = v;
        if (i + 16 <= m) { u = *reinterpret_cast<const float4*>(buf + i + 8); v = *reinterpret_cast<const float4*>(buf + i + 12); }   // next group, ahead of the adds
        float a8 = __fadd_rn(acc, cu.x);
        a8 = __fadd_rn(a8, cu.y); a8 = __fadd_rn(a8, cu.z); a8 = __fadd_rn(a8, cu.w);
        a8 = __fadd_rn(a8, cv.x); a8 = __fadd_rn(a8, cv.y); a8 = __fadd_rn(a8, cv.z); a8 = __fadd_rn(a8, cv.w);
        if (!(a8 <= amax)) break;
        acc = a8;
        t = cv.w;
      }
      for (; i < m; ++i) {                                  // the group at the crossing (or the chunk's tail), element by element
        acc = __fadd_rn(acc, buf[i]);
        if (!(acc <= amax)) { done = 1; break; }
        t = buf[i];
      }
    }
    done = __shfl_sync(0xffffffffu, done, 0);
  }
  if (lane == 0) thr[b] = t;
}

__global__ void __launch_bounds__(256) k_label_mask(int F, int NP, int vad, const int32_t* __restrict__ frame_utt, const float* __restrict__ pw,
                                                    const float* __restrict__ thr, float* __restrict__ y) {
  const size_t total = (size_t)(vad ? 1 : F) * NP;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int gn = (int)(i % NP), b = frame_utt[gn];
    y[i] = (b >= 0 && pw[i] > thr[b]) ? 1.f : 0.f;          // round(0.5 + w * (m - 0.5)) = m for 0 < w <= 1 (target.py:22-24)
  }
}

}  // namespace

// workspace: seg_off[B+1] (size_t) | thr[B] | power [F or 1][NP] | keys | tmp
static size_t label_elems(const gvn_batch* b) { return (size_t)b->F * b->NP; }
size_t speech_labels_workspace_bytes(const gvn_batch* b) {
  return round_up((size_t)(b->B + 1) * sizeof(size_t) + (size_t)b->B * 4, 256) + 3 * round_up(label_elems(b) * 4, 256);
}

// seg_off is filled on the device from n_frames (no host copy): utterance b's keys start at F * (frames before b)
__global__ void k_label_offsets(int B, int F, int vad, const int32_t* __restrict__ n_frames, size_t* __restrict__ seg_off) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    size_t run = 0;
    for (int b = 0; b < B; ++b) { seg_off[b] = run; run += (size_t)n_frames[b] * (vad ? 1 : F); }
    seg_off[B] = run;
  }
}

int32_t launch_speech_labels(const gvn_batch* b, const float* S, int from_power, int vad, float qf, float* y, void* workspace,
                             cudaStream_t st) {
  unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
  size_t* seg_off = reinterpret_cast<size_t*>(ws);
  float* thr = reinterpret_cast<float*>(ws + (size_t)(b->B + 1) * sizeof(size_t));
  const size_t head = round_up((size_t)(b->B + 1) * sizeof(size_t) + (size_t)b->B * 4, 256), plane = round_up(label_elems(b) * 4, 256);
  float* pw = reinterpret_cast<float*>(ws + head);
  uint32_t* keys = reinterpret_cast<uint32_t*>(ws + head + plane);
  uint32_t* tmp = reinterpret_cast<uint32_t*>(ws + head + 2 * plane);
  k_label_offsets<<<1, 32, 0, st>>>(b->B, b->F, vad, b->n_frames, seg_off);
  int32_t rc = check_launch("k_label_offsets");
  if (rc) return rc;
  k_label_keys<<<(b->NP + 31) / 32, 256, 0, st>>>(b->F, b->NP, vad, from_power, b->frame_utt, b->frame_off, b->n_frames, S, seg_off, keys, pw);
  if ((rc = check_launch("k_label_keys"))) return rc;
  k_label_sort<<<b->B, LT, 0, st>>>(b->F, vad, b->n_frames, seg_off, keys, tmp);
  if ((rc = check_launch("k_label_sort"))) return rc;
  k_label_threshold<<<b->B, 32, 0, st>>>(b->F, vad, qf, b->n_frames, seg_off, keys, tmp, thr);
  if ((rc = check_launch("k_label_threshold"))) return rc;
  const size_t total = (size_t)(vad ? 1 : b->F) * b->NP;
  int grid = (int)((total + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  k_label_mask<<<grid, 256, 0, st>>>(b->F, b->NP, vad, b->frame_utt, pw, thr, y);
  return check_launch("k_label_mask");
}

}  // namespace gvn
