// Multi-scale deformable attention, forward + backward, for B200 (sm_100a).
//
// Replaces mmcv-full 1.4.0 `_ext.ms_deform_attn_{forward,backward}` as called from
// projects/mmdet3d_plugin/bevformer/modules/multi_scale_deformable_attn_function.py:118-124,150-160.
// Semantics: pixel = loc * (W,H) - 0.5, bilinear with per-corner zero
// padding, a sample contributes only if -1 < pixel < size; out = sum_{l,p} w * bilerp.
//
// Mapping.  The op is a sparse gather: per (b, q, h) it reads L*P samples x 4 corners x C
// floats (one 128-byte line per corner for C = 32).  One warp owns 32 samples at a time:
//   * lane j decodes sample j once (coordinates, corner offsets, validity) -- the
//     reference re-derives this once per *channel* thread (32x redundant);
//   * the warp is then split into NG = 32/CV groups of CV lanes, a lane holding 4
//     consecutive channels (float4): every corner fetch is a 16-byte vector load and a
//     group of CV lanes covers one head vector, so a warp-level load instruction
//     touches NG lines instead of 1 (4x fewer LSU instructions than lane = channel);
//   * sample parameters travel from the decoding lane to its group by warp shuffle;
//   * backward reduces grad_loc / grad_attn over channels with a 4-value transposed
//     shuffle reduction (no shared memory, no __syncthreads -- the reference does a
//     smem write + 2 barriers + a serial 32-way sum per sample), and scatters
//     grad_value with 16-byte vector reductions (red.global.add.v4.f32).
// Work order: for L*P >= 32 a block is 8 consecutive queries of one (b, h) and
// neighbouring blocks are the other heads of the same queries, so the lines a block
// touches are neighbours in the image (L1/L2 locality).
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace vidar {
namespace {

constexpr int kWarpsPerBlock = 8;

struct MsdaParams {
  const float* value;
  const int64_t* shapes;
  const int64_t* lsi;
  const float* loc;
  const float* attn;
  int B, K, H, C, L, Q, P, LP;
  int ipw;            // items (b,q,h) per warp (vector kernels)
  int lp_shift;       // log2(LP) when ipw > 1
  int pix_stride;     // H*C floats between neighbouring pixels
  long long items;    // B*Q*H
  // fused sampling-location / softmax epilogue (EPI kernels): `loc` then holds the raw offsets
  // of the sampling_offsets Linear, `attn` the logits of the attention_weights Linear
  const float* ref;   // [B, Q, D, 2] projected Z-anchors (reference_points_cam), normalised
  int D;
  // Row-indirect mode (IDX kernels): a row of batch element n = b*ncl + cl is entry j of camera
  // (cam0 + cl)'s visible-pillar list; results go to / gradients come from the dense BEV slot grid.
  const int* idx;      // [cams, Qd] pillar of row j, or nullptr (row j is pillar j)
  const int* count;    // [cams] live rows per camera, or nullptr (all Q rows are live)
  const float* inv;    // [bs, Qd] output scale 1 / #cameras seeing the pillar, or nullptr
  int ncl, cam0, bs, Qd;   // cameras in this launch, first camera, batch size, pillars per batch element
  int S, s_lo, s_hi;   // the launch covers rows with (j / 64) % S in [s_lo, s_hi)   (rank sub-slices)
  unsigned na_bytes;  // levels whose per-head slab (H_l*W_l*C*4 bytes) exceeds this stream past L1 (no_allocate)
};

// Decode one sample: pixel coordinates -> clamped base pixel, corner mask, fractions.
// meta = mask(4b) | dx << 4 | dyW << 5 ; mask == 0 <=> sample contributes nothing.
__device__ __forceinline__ void decode_sample(float lx, float ly, int Hl, int Wl, int start,
                                              int& base, int& meta, float& lh, float& lw,
                                              int scale = 1) {
  // mmcv: h_im = loc_h * spatial_h - 0.5 (no fused multiply-add there either)
  const float h_im = __fsub_rn(__fmul_rn(ly, (float)Hl), 0.5f);
  const float w_im = __fsub_rn(__fmul_rn(lx, (float)Wl), 0.5f);
  base = 0;
  meta = 0;
  lh = 0.f;
  lw = 0.f;
  if (h_im > -1.f && w_im > -1.f && h_im < (float)Hl && w_im < (float)Wl) {
    const float hf = floorf(h_im), wf = floorf(w_im);
    const int h_low = (int)hf, w_low = (int)wf;
    lh = h_im - hf;
    lw = w_im - wf;
    const bool h0 = h_low >= 0, w0 = w_low >= 0;
    const bool h1 = h_low + 1 <= Hl - 1, w1 = w_low + 1 <= Wl - 1;
    const int mask = (int)(h0 && w0) | ((int)(h0 && w1) << 1) | ((int)(h1 && w0) << 2) |
                     ((int)(h1 && w1) << 3);
    const int hl = max(h_low, 0), wl = max(w_low, 0);
    const int hh = min(h_low + 1, Hl - 1), wh = min(w_low + 1, Wl - 1);
    // scale = 1: pixel units (generic kernels); scale = H*C: float offsets (vector kernels)
    base = (start + hl * Wl + wl) * scale;
    meta = mask | ((wh - wl) << 4) | (((hh - hl) * Wl * scale) << 5);
  }
}

// Blackwell packed fp32: one FFMA2 does two fused multiply-adds (fma.rn.f32x2, sm_100+), so
// a float4 of channels costs two issue slots instead of four.  Same IEEE rounding as FFMA.
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpk(u64 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) {
  u64 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ u64 fmul2(u64 a, u64 b) {
  u64 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// Cache policy of the gather.  The fine pyramid levels have (almost) no reuse inside an SM -- a
// 24x24-pixel window per anchor and two samples in it -- while one head's coarse levels (15x25: 48 KB,
// 29x50: 186 KB) are re-read by every query: fine-level lines are loaded with L1::no_allocate so they
// do not evict the coarse slabs (kHintBit of a sample's meta word selects the policy).
constexpr int kHintBit = 1 << 30;
__device__ __forceinline__ float4 ldg4_na(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float4 ldg4_h(const float* p, bool na) { return na ? ldg4_na(p) : ldg4(p); }

__device__ __forceinline__ float4 f4_zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ void f4_fma(float4& a, float s, const float4& v) {
  const u64 ss = pk(s, s);
  unpk(ffma2(ss, pk(v.x, v.y), pk(a.x, a.y)), a.x, a.y);
  unpk(ffma2(ss, pk(v.z, v.w), pk(a.z, a.w)), a.z, a.w);
}
__device__ __forceinline__ float f4_dot(const float4& a, const float4& b) {
  float lo, hi;
  unpk(ffma2(pk(a.x, a.y), pk(b.x, b.y), fmul2(pk(a.z, a.w), pk(b.z, b.w))), lo, hi);
  return lo + hi;
}
__device__ __forceinline__ float4 f4_scale(float s, const float4& v) {
  float4 r;
  const u64 ss = pk(s, s);
  unpk(fmul2(ss, pk(v.x, v.y)), r.x, r.y);
  unpk(fmul2(ss, pk(v.z, v.w)), r.z, r.w);
  return r;
}


// EPI (MSDeformableAttention3D, spatial_cross_attention.py:339-371 folded into the kernel): lane s
// holds sample s = l*P + p of one (b,q,h); the attention weight is the softmax of the 32 logits
// (`attention_weights.softmax(-1)`, :342) and the location is
//   loc = offsets / (W_l, H_l) + reference_points_cam[b, q, p % D]      (:356-371)
// with the reference's operation order (a true division, then an add; no FMA contraction).
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ void epi_decode(const MsdaParams& p, size_t refrow, int s, float Wl, float Hl,
                                           const float* loc0, const float* att0, float2& xy, float& aw) {
  const float logit = __ldg(att0 + s);
  const float e = expf(logit - warp_max(logit));
  aw = __fdiv_rn(e, warp_sum(e));
  const float2 off = __ldg(reinterpret_cast<const float2*>(loc0) + s);
  const int z = (s % p.P) % p.D;
  const float2 r = __ldg(reinterpret_cast<const float2*>(p.ref) + refrow + z);   // refrow: anchor 0 of this row
  xy.x = __fadd_rn(__fdiv_rn(off.x, Wl), r.x);
  xy.y = __fadd_rn(__fdiv_rn(off.y, Hl), r.y);
}

// Which (b,q,h) items does this warp own?  Returns false if none.
//  ipw == 1 : block = 8 consecutive queries of one (b,h); blocks ordered (b, h, qtile).
//  ipw  > 1 : ipw consecutive items in memory order (heads of the same query first).
// All divisions happen here, once per warp -- never inside the sample loop.
__device__ __forceinline__ bool warp_items(const MsdaParams& p, long long& item0) {
  const int warp = threadIdx.x >> 5;
  if (p.ipw == 1) {
    // blocks ordered (b, h, qtile): at any moment the whole chip works on ONE (b, h) slab, so
    // an SM's L1 holds a single head's coarse levels (15x25: 48 KB, 29x50: 186 KB) instead of
    // thrashing between eight.
    const unsigned blk = blockIdx.x;
    const unsigned nqt = (p.Q + kWarpsPerBlock - 1) / kWarpsPerBlock;
    const unsigned qt = blk % nqt;
    const unsigned t = blk / nqt;
    const unsigned h = t % p.H;
    const unsigned b = t / p.H;
    const int q = qt * kWarpsPerBlock + warp;
    if (q >= p.Q) return false;
    item0 = ((long long)b * p.Q + q) * p.H + h;
    return true;
  }
  item0 = ((long long)blockIdx.x * kWarpsPerBlock + warp) * p.ipw;
  return item0 < p.items;
}

// base pointer offset (floats) of item's (b, h) slab inside value / grad_value
__device__ __forceinline__ size_t slab_offset(const MsdaParams& p, long long item64) {
  const unsigned item = (unsigned)item64;            // items < 2^31 on the vector path
  const unsigned qh = (unsigned)p.Q * (unsigned)p.H;
  const unsigned b = item / qh;
  const unsigned h = item % (unsigned)p.H;
  return (size_t)b * p.K * p.pix_stride + (size_t)h * p.C;
}

// Where a row's per-query data lives.  Plain: everything is indexed by the item (rebatched layout).
// IDX: the row is (camera c, list entry j) -> pillar q; `dense` indexes per-(b, pillar, head) arrays
// (BEV slots, and -- in the fused-prologue kernels -- the offsets / logits computed once per pillar).
struct Row {
  long long dense;   // ((b*Qd + q)*H + h)
  size_t refrow;     // float2 index of the row's Z-anchor 0 in p.ref
  float inv;         // output scale
};
template <bool IDX>
__device__ __forceinline__ bool resolve_row(const MsdaParams& p, long long item0, Row& r) {
  if (!IDX) {
    r.dense = item0;
    r.refrow = (size_t)((unsigned)item0 / (unsigned)p.H) * p.D;      // ref [B, Q, D, 2]
    r.inv = 1.f;
    return true;
  }
  const unsigned item = (unsigned)item0;
  const unsigned h = item % (unsigned)p.H;
  const unsigned nj = item / (unsigned)p.H;
  const unsigned j = nj % (unsigned)p.Q, n = nj / (unsigned)p.Q;
  const unsigned cl = n % (unsigned)p.ncl, b = n / (unsigned)p.ncl;
  const unsigned c = (unsigned)p.cam0 + cl;
  if (p.count && (int)j >= __ldg(p.count + c)) return false;
  if (p.S > 1) {
    const int sl = (int)((j >> 6) % (unsigned)p.S);
    if (sl < p.s_lo || sl >= p.s_hi) return false;
  }
  const unsigned q = p.idx ? (unsigned)__ldg(p.idx + (size_t)c * p.Qd + j) : j;
  const size_t bq = (size_t)b * p.Qd + q;
  r.dense = (long long)(bq * p.H + h);
  r.refrow = (((size_t)c * p.bs + b) * p.Qd + q) * p.D;             // reference_points_cam [cams, bs, Qd, D, 2]
  r.inv = p.inv ? __ldg(p.inv + bq) : 1.f;
  return true;
}

__device__ __forceinline__ void red_add_v2(float* p, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}

template <int CV>
__device__ __forceinline__ void store_item(const MsdaParams& p, float* __restrict__ out,
                                           long long item, float4 acc, int g, int cl) {
#pragma unroll
  for (int off = CV; off < 32; off <<= 1) {
    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, off);
    acc.y += __shfl_xor_sync(0xffffffffu, acc.y, off);
    acc.z += __shfl_xor_sync(0xffffffffu, acc.z, off);
    acc.w += __shfl_xor_sync(0xffffffffu, acc.w, off);
  }
  if (g == 0 && item < p.items)
    *reinterpret_cast<float4*>(out + (size_t)item * p.C + cl * 4) = acc;
}

// MULTI = several items per warp (L*P a power of two < 32, one 32-sample chunk).
// EPI = fused softmax / sampling-location epilogue (L*P == 32, !MULTI only).
// 6 blocks (48 warps) per SM: the register cap (42) is what the tuned round-1 kernel used (40)
template <int CV, bool MULTI, bool EPI = false, bool IDX = false>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, 6)
msda_forward_kernel(const MsdaParams p, float* __restrict__ out) {
  constexpr int NG = 32 / CV;       // sample groups per warp
  constexpr int ITERS = 32 / NG;    // == CV
  const int lane = threadIdx.x & 31;
  const int g = lane / CV;          // group = which sample of the current NG
  const int cl = lane % CV;         // which float4 of the head vector
  long long item0;
  if (!warp_items(p, item0)) return;
  Row row;
  if (!resolve_row<IDX>(p, item0, row)) return;

  const int span = MULTI ? 32 : p.LP;                  // samples owned by this warp
  const int nchunks = MULTI ? 1 : (span + 31) >> 5;
  // fused prologue + IDX: offsets / logits exist once per pillar (dense); otherwise per row
  const long long prow = (EPI && IDX) ? row.dense : item0;
  const float* loc0 = p.loc + (size_t)prow * p.LP * 2;
  const float* att0 = p.attn + (size_t)prow * p.LP;
  const float* vb = p.value + slab_offset(p, item0) + cl * 4;   // moves with the item (MULTI)
  const unsigned pix = (unsigned)p.pix_stride;

  float4 acc = f4_zero();
  for (int chunk = 0; chunk < nchunks; ++chunk) {
    // ---- lane j decodes sample j of the chunk; weights are pre-multiplied by attention
    const int s = chunk * 32 + lane;
    int base = 0, meta = 0;
    float w1 = 0.f, w2 = 0.f, w3 = 0.f, w4 = 0.f;
    const bool live = MULTI ? (item0 + (s >> p.lp_shift) < p.items) : (s < span);
    if (live) {
      const int sl = MULTI ? (s & (p.LP - 1)) : s;
      const int l = sl / p.P;
      const int Hl = (int)__ldg(p.shapes + 2 * l), Wl = (int)__ldg(p.shapes + 2 * l + 1);
      float2 xy;
      float aw;
      if (EPI) {            // all 32 lanes are live here (L*P == 32)
        epi_decode(p, row.refrow, s, (float)Wl, (float)Hl, loc0, att0, xy, aw);
      } else {
        xy = __ldg(reinterpret_cast<const float2*>(loc0) + s);
        aw = __ldg(att0 + s);
      }
      float lh, lw;
      decode_sample(xy.x, xy.y, Hl, Wl, (int)__ldg(p.lsi + l), base, meta, lh, lw, p.pix_stride);
      if (meta && (unsigned)(Hl * Wl) * (unsigned)p.C * 4u > p.na_bytes) meta |= kHintBit;
      const float hh = 1.f - lh, hw = 1.f - lw;
      // invalid corners get weight 0; their (clamped) address aliases a valid corner's pixel
      w1 = (meta & 1) ? aw * (hh * hw) : 0.f;
      w2 = (meta & 2) ? aw * (hh * lw) : 0.f;
      w3 = (meta & 4) ? aw * (lh * hw) : 0.f;
      w4 = (meta & 8) ? aw * (lh * lw) : 0.f;
    }
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      if (!MULTI && chunk * 32 + it * NG >= span) break;
      const int src = it * NG + g;
      const int sbase = __shfl_sync(0xffffffffu, base, src);
      const int smeta = __shfl_sync(0xffffffffu, meta, src);
      const float a1 = __shfl_sync(0xffffffffu, w1, src);
      const float a2 = __shfl_sync(0xffffffffu, w2, src);
      const float a3 = __shfl_sync(0xffffffffu, w3, src);
      const float a4 = __shfl_sync(0xffffffffu, w4, src);
      if (MULTI && it > 0 && ((it * NG) & (p.LP - 1)) == 0) {
        // item changes every LP/NG iterations: flush the finished one, move the slab pointer
        const int il = (it * NG) >> p.lp_shift;
        store_item<CV>(p, out, item0 + il - 1, acc, g, cl);
        acc = f4_zero();
        vb = p.value + slab_offset(p, item0 + il) + cl * 4;
      }
      const float* vbi = vb;
      asm volatile("" : "+l"(vbi));   // keep the slab base in a 64-bit register pair
      if (smeta & 15) {
        // four independent 16-byte loads, issued back to back (32-bit offsets, 64-bit base)
        const unsigned o1 = (unsigned)sbase;                       // float offsets inside the slab
        const unsigned o2 = o1 + ((smeta & 16) ? pix : 0u);
        const unsigned o3 = o1 + (unsigned)((smeta & ~kHintBit) >> 5);
        const unsigned o4 = o3 + (o2 - o1);
        const bool na = (smeta & kHintBit) != 0;      // warp-uniform when a level's points fill whole iterations
        const float4 v1 = ldg4_h(vbi + o1, na);
        const float4 v2 = ldg4_h(vbi + o2, na);
        const float4 v3 = ldg4_h(vbi + o3, na);
        const float4 v4 = ldg4_h(vbi + o4, na);
        f4_fma(acc, a1, v1);
        f4_fma(acc, a2, v2);
        f4_fma(acc, a3, v3);
        f4_fma(acc, a4, v4);
      }
    }
  }
  if (IDX) {
    // SpatialCrossAttention's scatter-add + count normalisation (spatial_cross_attention.py:164-171) as
    // the epilogue: the head vector goes straight into the BEV slot of its pillar.
#pragma unroll
    for (int off = CV; off < 32; off <<= 1) {
      acc.x += __shfl_xor_sync(0xffffffffu, acc.x, off);
      acc.y += __shfl_xor_sync(0xffffffffu, acc.y, off);
      acc.z += __shfl_xor_sync(0xffffffffu, acc.z, off);
      acc.w += __shfl_xor_sync(0xffffffffu, acc.w, off);
    }
    if (g == 0) red_add_v4(out + (size_t)row.dense * p.C + cl * 4, f4_scale(row.inv, acc));
    return;
  }
  store_item<CV>(p, out, MULTI ? item0 + p.ipw - 1 : item0, acc, g, cl);
}

// 4 blocks (32 warps) per SM = 64 registers, the round-1 operating point
template <int CV, bool MULTI, bool EPI = false, bool IDX = false>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, 4)
msda_backward_kernel(const MsdaParams p, const float* __restrict__ grad_out,
                     float* __restrict__ grad_value, float* __restrict__ grad_loc,
                     float* __restrict__ grad_attn) {
  constexpr int NG = 32 / CV;
  constexpr int ITERS = 32 / NG;
  const int lane = threadIdx.x & 31;
  const int g = lane / CV;
  const int cl = lane % CV;
  long long item0;
  if (!warp_items(p, item0)) return;
  Row row;
  if (!resolve_row<IDX>(p, item0, row)) return;

  const int span = MULTI ? 32 : p.LP;
  const int nchunks = MULTI ? 1 : (span + 31) >> 5;
  const long long prow = (EPI && IDX) ? row.dense : item0;
  const float* loc0 = p.loc + (size_t)prow * p.LP * 2;
  const float* att0 = p.attn + (size_t)prow * p.LP;
  float* gloc0 = grad_loc + (size_t)prow * p.LP * 2;
  float* gatt0 = grad_attn + (size_t)prow * p.LP;
  const unsigned pix = (unsigned)p.pix_stride;
  size_t slab = slab_offset(p, item0) + cl * 4;
  float4 go = f4_zero();
  // IDX: grad_out is the gradient of the BEV slot grid; the row's share is slot[pillar] / count
  if (!MULTI) go = IDX ? f4_scale(row.inv, ldg4(grad_out + (size_t)row.dense * p.C + cl * 4))
                       : ldg4(grad_out + (size_t)item0 * p.C + cl * 4);

  for (int chunk = 0; chunk < nchunks; ++chunk) {
    const int s = chunk * 32 + lane;
    int base = 0, meta = 0;
    float lh = 0.f, lw = 0.f, aw = 0.f, fH = 0.f, fW = 0.f;
    float w1 = 0.f, w2 = 0.f, w3 = 0.f, w4 = 0.f;
    const bool live = MULTI ? (item0 + (s >> p.lp_shift) < p.items) : (s < span);
    if (live) {
      const int sl = MULTI ? (s & (p.LP - 1)) : s;
      const int l = sl / p.P;
      const int Hl = (int)__ldg(p.shapes + 2 * l), Wl = (int)__ldg(p.shapes + 2 * l + 1);
      float2 xy;
      if (EPI) {
        epi_decode(p, row.refrow, s, (float)Wl, (float)Hl, loc0, att0, xy, aw);
      } else {
        xy = __ldg(reinterpret_cast<const float2*>(loc0) + s);
        aw = __ldg(att0 + s);
      }
      decode_sample(xy.x, xy.y, Hl, Wl, (int)__ldg(p.lsi + l), base, meta, lh, lw, p.pix_stride);
      if (meta && (unsigned)(Hl * Wl) * (unsigned)p.C * 4u > p.na_bytes) meta |= kHintBit;
      fH = (float)Hl;
      fW = (float)Wl;
      const float hh = 1.f - lh, hw = 1.f - lw;
      w1 = (meta & 1) ? aw * (hh * hw) : 0.f;
      w2 = (meta & 2) ? aw * (hh * lw) : 0.f;
      w3 = (meta & 4) ? aw * (lh * hw) : 0.f;
      w4 = (meta & 8) ? aw * (lh * lw) : 0.f;
    }
    float r1 = 0.f, r2 = 0.f, r3 = 0.f, r4 = 0.f;   // this lane's sample: sum_c g*v_k
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      if (!MULTI && chunk * 32 + it * NG >= span) break;
      const int src = it * NG + g;
      const int sbase = __shfl_sync(0xffffffffu, base, src);
      const int smeta = __shfl_sync(0xffffffffu, meta, src);
      const float a1 = __shfl_sync(0xffffffffu, w1, src);
      const float a2 = __shfl_sync(0xffffffffu, w2, src);
      const float a3 = __shfl_sync(0xffffffffu, w3, src);
      const float a4 = __shfl_sync(0xffffffffu, w4, src);
      if (MULTI && ((it * NG) & (p.LP - 1)) == 0) {
        const long long item = item0 + ((it * NG) >> p.lp_shift);
        if (item < p.items) {
          slab = slab_offset(p, item) + cl * 4;
          go = ldg4(grad_out + (size_t)item * p.C + cl * 4);
        } else {
          go = f4_zero();
        }
      }
      float d1 = 0.f, d2 = 0.f, d3 = 0.f, d4 = 0.f;
      if (smeta & 15) {
        const unsigned o1 = (unsigned)sbase;                       // float offsets inside the slab
        const unsigned o2 = o1 + ((smeta & 16) ? pix : 0u);
        const unsigned o3 = o1 + (unsigned)((smeta & ~kHintBit) >> 5);
        const unsigned o4 = o3 + (o2 - o1);
        const float* vs = p.value + slab;
        float* gs = grad_value + slab;
        asm volatile("" : "+l"(vs), "+l"(gs));
        const bool na = (smeta & kHintBit) != 0;
        const float4 v1 = ldg4_h(vs + o1, na);
        const float4 v2 = ldg4_h(vs + o2, na);
        const float4 v3 = ldg4_h(vs + o3, na);
        const float4 v4 = ldg4_h(vs + o4, na);
        // grad_value: top_grad * attn * corner weight as 16-byte vector reductions.  An
        // out-of-image corner has weight 0 and aliases a valid pixel: its reduction adds 0.
        red_add_v4(gs + o1, f4_scale(a1, go));
        red_add_v4(gs + o2, f4_scale(a2, go));
        red_add_v4(gs + o3, f4_scale(a3, go));
        red_add_v4(gs + o4, f4_scale(a4, go));
        d1 = f4_dot(go, v1); d2 = f4_dot(go, v2); d3 = f4_dot(go, v3); d4 = f4_dot(go, v4);
      }
      // ---- transposed reduction of (d1..d4) over the CV lanes of the group: 4 shuffles;
      //      afterwards lane class (cl & CV/2, cl & CV/4) = (0,0):d1 (0,1):d2 (1,0):d3 (1,1):d4
      const bool up1 = (cl & (CV / 2)) != 0;
      const float e0 = __shfl_xor_sync(0xffffffffu, up1 ? d1 : d3, CV / 2);
      const float e1 = __shfl_xor_sync(0xffffffffu, up1 ? d2 : d4, CV / 2);
      const float k0 = (up1 ? d3 : d1) + e0;    // lower half: d1   upper half: d3
      const float k1 = (up1 ? d4 : d2) + e1;    // lower half: d2   upper half: d4
      const bool up2 = (cl & (CV / 4)) != 0;
      const float eb = __shfl_xor_sync(0xffffffffu, up2 ? k0 : k1, CV / 4);
      float k = (up2 ? k1 : k0) + eb;
#pragma unroll
      for (int off = CV / 8; off >= 1; off >>= 1) k += __shfl_xor_sync(0xffffffffu, k, off);
      // ---- hand the four sums back to the lane that decoded the sample (lane it*NG + g')
      const int back = (lane & (NG - 1)) * CV;
      const float t1 = __shfl_sync(0xffffffffu, k, back);
      const float t2 = __shfl_sync(0xffffffffu, k, back + CV / 4);
      const float t3 = __shfl_sync(0xffffffffu, k, back + CV / 2);
      const float t4 = __shfl_sync(0xffffffffu, k, back + CV / 2 + CV / 4);
      if ((lane / NG) == it) { r1 = t1; r2 = t2; r3 = t3; r4 = t4; }
    }
    // ---- every lane finishes its own sample; coalesced stores
    if (live) {
      const float hh = 1.f - lh, hw = 1.f - lw;
      // an out-of-image corner contributes v = 0 to every gradient (mmcv col2im_bilinear)
      if (!(meta & 1)) r1 = 0.f;
      if (!(meta & 2)) r2 = 0.f;
      if (!(meta & 4)) r3 = 0.f;
      if (!(meta & 8)) r4 = 0.f;
      const bool ok = (meta & 15) != 0;
      const float ga = ok ? (hh * hw) * r1 + (hh * lw) * r2 + (lh * hw) * r3 + (lh * lw) * r4 : 0.f;
      const float gx = ok ? fW * aw * (hh * (r2 - r1) + lh * (r4 - r3)) : 0.f;
      const float gy = ok ? fH * aw * (hw * (r3 - r1) + lw * (r4 - r2)) : 0.f;
      if (EPI) {
        // through loc = off / (W,H) + ref and w = softmax(logits): grad_off = grad_loc / (W,H),
        // grad_logit = w * (grad_w - sum_j w_j grad_w_j)   (all 32 lanes take part in the sum)
        const float dot = warp_sum(aw * ga);
        if (IDX) {
          // offsets / logits are shared by every camera that sees the pillar: accumulate (caller zeroes)
          red_add_f32(gatt0 + s, aw * (ga - dot));
          red_add_v2(gloc0 + 2 * s, __fdiv_rn(gx, fW), __fdiv_rn(gy, fH));
        } else {
          gatt0[s] = aw * (ga - dot);
          reinterpret_cast<float2*>(gloc0)[s] = make_float2(__fdiv_rn(gx, fW), __fdiv_rn(gy, fH));
        }
      } else {
        gatt0[s] = ga;
        reinterpret_cast<float2*>(gloc0)[s] = make_float2(gx, gy);
      }
    }
  }
}

// ---- small L*P (temporal self-attention, future decoder: L*P = 4 or 8) -------------------
// When a 32-sample chunk holds at least NG whole items, give each group of CV lanes its OWN
// item: the group walks that item's L*P samples and finishes the head vector by itself, so the
// cross-group shuffle reduction + flush of the generic MULTI path (8 shuffles per item, i.e.
// 2 shuffles per sample at L*P = 4) disappears and every group stores 128 contiguous bytes.
template <int CV>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
msda_forward_small_kernel(const MsdaParams p, float* __restrict__ out) {
  constexpr int NG = 32 / CV;
  const int lane = threadIdx.x & 31;
  const int g = lane / CV, cl = lane % CV;
  long long item0;
  if (!warp_items(p, item0)) return;
  const float* loc0 = p.loc + (size_t)item0 * p.LP * 2;
  const float* att0 = p.attn + (size_t)item0 * p.LP;
  const unsigned pix = (unsigned)p.pix_stride;
  // lane j decodes sample j of the chunk (32 samples = ipw items)
  int base = 0, meta = 0;
  float w1 = 0.f, w2 = 0.f, w3 = 0.f, w4 = 0.f;
  if (item0 + (lane >> p.lp_shift) < p.items) {
    const float2 xy = __ldg(reinterpret_cast<const float2*>(loc0) + lane);
    const float aw = __ldg(att0 + lane);
    const int l = (lane & (p.LP - 1)) / p.P;
    const int Hl = (int)__ldg(p.shapes + 2 * l), Wl = (int)__ldg(p.shapes + 2 * l + 1);
    float lh, lw;
    decode_sample(xy.x, xy.y, Hl, Wl, (int)__ldg(p.lsi + l), base, meta, lh, lw, p.pix_stride);
    const float hh = 1.f - lh, hw = 1.f - lw;
    w1 = (meta & 1) ? aw * (hh * hw) : 0.f;
    w2 = (meta & 2) ? aw * (hh * lw) : 0.f;
    w3 = (meta & 4) ? aw * (lh * hw) : 0.f;
    w4 = (meta & 8) ? aw * (lh * lw) : 0.f;
  }
  const int rounds = p.ipw / NG;
  for (int r = 0; r < rounds; ++r) {
    const int il = r * NG + g;                       // this group's item in the chunk
    const long long item = item0 + il;
    const bool live = item < p.items;
    const float* vb = p.value + (live ? slab_offset(p, item) : 0) + cl * 4;
    asm volatile("" : "+l"(vb));
    float4 acc = f4_zero();
    for (int sidx = 0; sidx < p.LP; ++sidx) {
      const int src = (il << p.lp_shift) + sidx;
      const int sbase = __shfl_sync(0xffffffffu, base, src);
      const int smeta = __shfl_sync(0xffffffffu, meta, src);
      const float a1 = __shfl_sync(0xffffffffu, w1, src);
      const float a2 = __shfl_sync(0xffffffffu, w2, src);
      const float a3 = __shfl_sync(0xffffffffu, w3, src);
      const float a4 = __shfl_sync(0xffffffffu, w4, src);
      if (smeta & 15) {
        const unsigned o1 = (unsigned)sbase;
        const unsigned o2 = o1 + ((smeta & 16) ? pix : 0u);
        const unsigned o3 = o1 + (unsigned)(smeta >> 5);
        const unsigned o4 = o3 + (o2 - o1);
        const float4 v1 = ldg4(vb + o1);
        const float4 v2 = ldg4(vb + o2);
        const float4 v3 = ldg4(vb + o3);
        const float4 v4 = ldg4(vb + o4);
        f4_fma(acc, a1, v1);
        f4_fma(acc, a2, v2);
        f4_fma(acc, a3, v3);
        f4_fma(acc, a4, v4);
      }
    }
    if (live) *reinterpret_cast<float4*>(out + (size_t)item * p.C + cl * 4) = acc;
  }
}

template <int CV>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
msda_backward_small_kernel(const MsdaParams p, const float* __restrict__ grad_out,
                           float* __restrict__ grad_value, float* __restrict__ grad_loc,
                           float* __restrict__ grad_attn) {
  constexpr int NG = 32 / CV;
  const int lane = threadIdx.x & 31;
  const int g = lane / CV, cl = lane % CV;
  long long item0;
  if (!warp_items(p, item0)) return;
  const float* loc0 = p.loc + (size_t)item0 * p.LP * 2;
  const float* att0 = p.attn + (size_t)item0 * p.LP;
  const unsigned pix = (unsigned)p.pix_stride;
  int base = 0, meta = 0;
  float lh = 0.f, lw = 0.f, aw = 0.f, fH = 0.f, fW = 0.f;
  float w1 = 0.f, w2 = 0.f, w3 = 0.f, w4 = 0.f;
  const bool mine = item0 + (lane >> p.lp_shift) < p.items;
  if (mine) {
    const float2 xy = __ldg(reinterpret_cast<const float2*>(loc0) + lane);
    aw = __ldg(att0 + lane);
    const int l = (lane & (p.LP - 1)) / p.P;
    const int Hl = (int)__ldg(p.shapes + 2 * l), Wl = (int)__ldg(p.shapes + 2 * l + 1);
    decode_sample(xy.x, xy.y, Hl, Wl, (int)__ldg(p.lsi + l), base, meta, lh, lw, p.pix_stride);
    fH = (float)Hl;
    fW = (float)Wl;
    const float hh = 1.f - lh, hw = 1.f - lw;
    w1 = (meta & 1) ? aw * (hh * hw) : 0.f;
    w2 = (meta & 2) ? aw * (hh * lw) : 0.f;
    w3 = (meta & 4) ? aw * (lh * hw) : 0.f;
    w4 = (meta & 8) ? aw * (lh * lw) : 0.f;
  }
  float r1 = 0.f, r2 = 0.f, r3 = 0.f, r4 = 0.f;
  const int rounds = p.ipw / NG;
  const int my_il = lane >> p.lp_shift;              // item (in the chunk) of the sample this lane decoded
  for (int r = 0; r < rounds; ++r) {
    const int il = r * NG + g;
    const long long item = item0 + il;
    const bool live = item < p.items;
    const size_t slab = (live ? slab_offset(p, item) : 0) + cl * 4;
    const float4 go = live ? ldg4(grad_out + (size_t)item * p.C + cl * 4) : f4_zero();
    const float* vs = p.value + slab;
    float* gs = grad_value + slab;
    asm volatile("" : "+l"(vs), "+l"(gs));
    for (int sidx = 0; sidx < p.LP; ++sidx) {
      const int src = (il << p.lp_shift) + sidx;
      const int sbase = __shfl_sync(0xffffffffu, base, src);
      const int smeta = __shfl_sync(0xffffffffu, meta, src);
      const float a1 = __shfl_sync(0xffffffffu, w1, src);
      const float a2 = __shfl_sync(0xffffffffu, w2, src);
      const float a3 = __shfl_sync(0xffffffffu, w3, src);
      const float a4 = __shfl_sync(0xffffffffu, w4, src);
      float d1 = 0.f, d2 = 0.f, d3 = 0.f, d4 = 0.f;
      if (smeta & 15) {
        const unsigned o1 = (unsigned)sbase;
        const unsigned o2 = o1 + ((smeta & 16) ? pix : 0u);
        const unsigned o3 = o1 + (unsigned)(smeta >> 5);
        const unsigned o4 = o3 + (o2 - o1);
        const float4 v1 = ldg4(vs + o1);
        const float4 v2 = ldg4(vs + o2);
        const float4 v3 = ldg4(vs + o3);
        const float4 v4 = ldg4(vs + o4);
        red_add_v4(gs + o1, f4_scale(a1, go));
        red_add_v4(gs + o2, f4_scale(a2, go));
        red_add_v4(gs + o3, f4_scale(a3, go));
        red_add_v4(gs + o4, f4_scale(a4, go));
        d1 = f4_dot(go, v1); d2 = f4_dot(go, v2); d3 = f4_dot(go, v3); d4 = f4_dot(go, v4);
      }
      const bool up1 = (cl & (CV / 2)) != 0;
      const float e0 = __shfl_xor_sync(0xffffffffu, up1 ? d1 : d3, CV / 2);
      const float e1 = __shfl_xor_sync(0xffffffffu, up1 ? d2 : d4, CV / 2);
      const float k0 = (up1 ? d3 : d1) + e0;
      const float k1 = (up1 ? d4 : d2) + e1;
      const bool up2 = (cl & (CV / 4)) != 0;
      const float eb = __shfl_xor_sync(0xffffffffu, up2 ? k0 : k1, CV / 4);
      float k = (up2 ? k1 : k0) + eb;
#pragma unroll
      for (int off = CV / 8; off >= 1; off >>= 1) k += __shfl_xor_sync(0xffffffffu, k, off);
      // hand back to the decoding lane: its sample is processed by group (my_il % NG) in round my_il / NG
      const int back = (my_il & (NG - 1)) * CV;
      const float t1 = __shfl_sync(0xffffffffu, k, back);
      const float t2 = __shfl_sync(0xffffffffu, k, back + CV / 4);
      const float t3 = __shfl_sync(0xffffffffu, k, back + CV / 2);
      const float t4 = __shfl_sync(0xffffffffu, k, back + CV / 2 + CV / 4);
      if ((my_il / NG) == r && (lane & (p.LP - 1)) == sidx) { r1 = t1; r2 = t2; r3 = t3; r4 = t4; }
    }
  }
  if (mine) {
    const float hh = 1.f - lh, hw = 1.f - lw;
    if (!(meta & 1)) r1 = 0.f;
    if (!(meta & 2)) r2 = 0.f;
    if (!(meta & 4)) r3 = 0.f;
    if (!(meta & 8)) r4 = 0.f;
    const bool ok = (meta & 15) != 0;
    const float ga = ok ? (hh * hw) * r1 + (hh * lw) * r2 + (lh * hw) * r3 + (lh * lw) * r4 : 0.f;
    const float gx = ok ? fW * aw * (hh * (r2 - r1) + lh * (r4 - r3)) : 0.f;
    const float gy = ok ? fH * aw * (hw * (r3 - r1) + lw * (r4 - r2)) : 0.f;
    grad_attn[(size_t)item0 * p.LP + lane] = ga;
    reinterpret_cast<float2*>(grad_loc)[(size_t)item0 * p.LP + lane] = make_float2(gx, gy);
  }
}

#include "msda_slab.cuh"

// ---- generic fallback (any C): one thread per (item, channel); scalar atomics.
__global__ void msda_forward_generic_kernel(const MsdaParams p, float* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= p.items * p.C) return;
  const int c = (int)(idx % p.C);
  const long long item = idx / p.C;
  const int h = (int)(item % p.H);
  const int b = (int)(item / ((long long)p.Q * p.H));
  const size_t pix_stride = (size_t)p.H * p.C;
  const float* vb = p.value + ((size_t)b * p.K * p.H + h) * p.C + c;
  float acc = 0.f;
  for (int s = 0; s < p.LP; ++s) {
    const int l = s / p.P;
    const int Hl = (int)p.shapes[2 * l], Wl = (int)p.shapes[2 * l + 1];
    int base, meta;
    float lh, lw;
    decode_sample(p.loc[((size_t)item * p.LP + s) * 2], p.loc[((size_t)item * p.LP + s) * 2 + 1],
                  Hl, Wl, (int)p.lsi[l], base, meta, lh, lw);
    const int mask = meta & 15;
    if (!mask) continue;
    const float aw = p.attn[(size_t)item * p.LP + s];
    const size_t dx = (size_t)((meta >> 4) & 1) * pix_stride, dy = (size_t)(meta >> 5) * pix_stride;
    const float* p1 = vb + (size_t)base * pix_stride;
    const float v1 = (mask & 1) ? p1[0] : 0.f, v2 = (mask & 2) ? p1[dx] : 0.f;
    const float v3 = (mask & 4) ? p1[dy] : 0.f, v4 = (mask & 8) ? p1[dy + dx] : 0.f;
    const float hh = 1.f - lh, hw = 1.f - lw;
    acc += aw * (hh * hw * v1 + hh * lw * v2 + lh * hw * v3 + lh * lw * v4);
  }
  out[idx] = acc;
}

__global__ void msda_backward_generic_kernel(const MsdaParams p, const float* __restrict__ grad_out,
                                             float* __restrict__ grad_value,
                                             float* __restrict__ grad_loc,
                                             float* __restrict__ grad_attn) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= p.items * p.C) return;
  const int c = (int)(idx % p.C);
  const long long item = idx / p.C;
  const int h = (int)(item % p.H);
  const int b = (int)(item / ((long long)p.Q * p.H));
  const size_t pix_stride = (size_t)p.H * p.C;
  const size_t voff = ((size_t)b * p.K * p.H + h) * p.C + c;
  const float go = grad_out[idx];
  for (int s = 0; s < p.LP; ++s) {
    const int l = s / p.P;
    const int Hl = (int)p.shapes[2 * l], Wl = (int)p.shapes[2 * l + 1];
    int base, meta;
    float lh, lw;
    const size_t si = (size_t)item * p.LP + s;
    decode_sample(p.loc[si * 2], p.loc[si * 2 + 1], Hl, Wl, (int)p.lsi[l], base, meta, lh, lw);
    const int mask = meta & 15;
    if (!mask) continue;
    const float aw = p.attn[si];
    const size_t dx = (size_t)((meta >> 4) & 1) * pix_stride, dy = (size_t)(meta >> 5) * pix_stride;
    const size_t o1 = voff + (size_t)base * pix_stride;
    const float v1 = (mask & 1) ? p.value[o1] : 0.f, v2 = (mask & 2) ? p.value[o1 + dx] : 0.f;
    const float v3 = (mask & 4) ? p.value[o1 + dy] : 0.f, v4 = (mask & 8) ? p.value[o1 + dy + dx] : 0.f;
    const float hh = 1.f - lh, hw = 1.f - lw;
    const float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
    const float top = go * aw;
    if (mask & 1) red_add_f32(grad_value + o1, w1 * top);
    if (mask & 2) red_add_f32(grad_value + o1 + dx, w2 * top);
    if (mask & 4) red_add_f32(grad_value + o1 + dy, w3 * top);
    if (mask & 8) red_add_f32(grad_value + o1 + dy + dx, w4 * top);
    red_add_f32(grad_attn + si, go * (w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4));
    red_add_f32(grad_loc + si * 2, (float)Wl * top * (hh * (v2 - v1) + lh * (v4 - v3)));
    red_add_f32(grad_loc + si * 2 + 1, (float)Hl * top * (hw * (v3 - v1) + lw * (v4 - v2)));
  }
}

// FORWARD only: per-head level slabs larger than this are read with L1::no_allocate (see ldg4_na).
// Measured on B200, 6 cameras x 40000 queries (tools/exp_msda.py, profiles/r02_msda_cache_hints.json):
//   off 2.44 ms | > 1 MB (level 0) 2.42 | > 256 KB (levels 0-1) 2.32 | > 100 KB (levels 0-2) 2.26
// so everything but the 48 KB 15x25 level streams.  The backward gets slower with any of these
// (4.60 -> 4.68 / 4.74 ms: its loads share the LSU -> XBAR request path with the reductions, and a
// no-allocate miss cannot merge with a neighbouring warp's pending miss), so it keeps plain loads.
// VIDAR_MSDA_STREAM_BYTES overrides the forward threshold (experiments; 0xffffffff = off).
inline unsigned stream_threshold() {
  static const unsigned v = [] {
    const char* e = getenv("VIDAR_MSDA_STREAM_BYTES");
    return e ? (unsigned)strtoul(e, nullptr, 0) : 100000u;
  }();
  return v;
}

int fill_params(MsdaParams& p, const float* value, const int64_t* shapes, const int64_t* lsi,
                const float* loc, const float* attn, int B, int K, int H, int C, int L, int Q,
                int P, int im2col_step, const char* who) {
  VIDAR_REQUIRE(value && shapes && lsi && loc && attn, "%s: null pointer argument", who);
  VIDAR_REQUIRE(B > 0 && K > 0 && H > 0 && C > 0 && L > 0 && Q > 0 && P > 0,
                "%s: all of B,K,H,C,L,Q,P must be positive (got %d,%d,%d,%d,%d,%d,%d)", who, B, K,
                H, C, L, Q, P);
  VIDAR_REQUIRE(im2col_step > 0, "%s: im2col_step must be positive", who);
  const int step = im2col_step < B ? im2col_step : B;
  // mmcv: AT_ASSERTM(batch % im2col_step_ == 0, "batch(%d) must divide im2col_step(%d)")
  VIDAR_REQUIRE(B % step == 0, "%s: batch(%d) must divide im2col_step(%d)", who, B, step);
  p.value = value;
  p.shapes = shapes;
  p.lsi = lsi;
  p.loc = loc;
  p.attn = attn;
  p.B = B; p.K = K; p.H = H; p.C = C; p.L = L; p.Q = Q; p.P = P;
  p.LP = L * P;
  p.items = (long long)B * Q * H;
  p.ipw = 1;
  p.lp_shift = 0;
  p.pix_stride = H * C;
  p.ref = nullptr;
  p.D = 1;
  p.idx = nullptr; p.count = nullptr; p.inv = nullptr;
  p.ncl = B; p.cam0 = 0; p.bs = 1; p.Qd = Q;
  p.S = 1; p.s_lo = 0; p.s_hi = 1;
  p.na_bytes = stream_threshold();
  return VIDAR_OK;
}

inline int pick_ipw(int LP, int CV);
// vector kernels: C in {16,32,64}, 32-bit intra-batch offsets and item counts
inline bool vec_ok(const MsdaParams& p) {
  return (p.C == 16 || p.C == 32 || p.C == 64) && (long long)p.K * p.H * p.C < (1LL << 25) &&   /* offsets ride in meta bits 5..29 */
         p.items < (1LL << 31) && (long long)p.Q * p.H < (1LL << 31);
}
inline void set_ipw(MsdaParams& p, int CV) {
  p.ipw = pick_ipw(p.LP, CV);
  p.lp_shift = 0;
  while ((1 << p.lp_shift) < p.LP) ++p.lp_shift;
}

// items per warp for the vector kernels: pack several (b,q,h) into one 32-sample chunk
// when L*P is small (temporal self-attention: L*P = 4).
inline int pick_ipw(int LP, int CV) {
  const int NG = 32 / CV;
  if (LP >= 2 && LP < 32 && 32 % LP == 0 && LP % NG == 0) return 32 / LP;
  return 1;
}

inline long long num_blocks(const MsdaParams& p) {
  if (p.ipw == 1) {
    const long long nqt = (p.Q + kWarpsPerBlock - 1) / kWarpsPerBlock;
    return (long long)p.B * nqt * p.H;
  }
  const long long warps = (p.items + p.ipw - 1) / p.ipw;
  return (warps + kWarpsPerBlock - 1) / kWarpsPerBlock;
}


// ---- slab (persistent, shared-memory) variants: host side --------------------------------------
// VIDAR_MSDA_SLAB: 0 = off (default), 1 = slab kernels with plain staging / flush, 2 = with TMA.  Measured on
// B200 (profiles/r02_msda_slab_experiment.json): backward 4.48 -> 5.25 ms, forward 2.06 -> 2.19 ms per 6
// cameras -- the shared-memory hand-over costs more than the 25 % of reduction lines it removes (a barrier
// per 16-query tile, ~10 issue slots per record, half the L1), with or without TMA.  The variants stay
// selectable (and parity-tested, tests/test_msda_gpu.py::test_slab_variants) but off.
inline int slab_mode() {
  static const int v = [] {
    const char* e = getenv("VIDAR_MSDA_SLAB");
    return e ? atoi(e) : 0;
  }();
  return v;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn encode_tiled() {
  static const EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      f = nullptr;
    }
    return (EncodeTiledFn)f;
  }();
  return fn;
}

// 2-D map over a [rows = B*K, cols = H*C] fp32 matrix (value or grad_value) with a [kSlabBoxRows x 32] box:
// a head's 128-byte column slice of 128 consecutive pixel rows.
inline bool make_slab_map(CUtensorMap* m, const float* base, long long rows, int cols) {
  memset(m, 0, sizeof(*m));
  EncodeTiledFn enc = encode_tiled();
  if (!enc || ((uintptr_t)base & 15u) || (cols & 3)) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)cols * sizeof(float)};
  const cuuint32_t box[2] = {32u, (cuuint32_t)kSlabBoxRows};
  const cuuint32_t estr[2] = {1u, 1u};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

inline bool slab_ok(const MsdaParams& p) {
  return slab_mode() > 0 && p.C == 32 && p.LP == 32 && p.P % 4 == 0 && p.L >= 1 && (p.pix_stride & (p.pix_stride - 1)) == 0 &&
         vec_ok(p) && (long long)p.B * p.H < (1 << 20);
}

inline void slab_geo(const MsdaParams& p, SlabGeo& sg, int& grid) {
  const int slots = 2 * kNumSMs;
  const int nh = p.B * p.H;
  sg.first_it = (p.LP - p.P) / 4;
  sg.pix_shift = 0;
  while ((1 << sg.pix_shift) < p.pix_stride) ++sg.pix_shift;
  sg.tiles = (p.Q + kSlabWarps - 1) / kSlabWarps;
  sg.parts = slots / nh > 1 ? slots / nh : 1;
  if (sg.parts > sg.tiles) sg.parts = sg.tiles > 0 ? sg.tiles : 1;
  sg.use_tma = 0;
  const long long units = (long long)nh * sg.parts;
  grid = (int)(units < slots ? units : slots);
}

constexpr size_t kSlabBwdSmem = (size_t)kSlabMaxRows * 128 + 2 * kSlabRecords * 4 + 2 * kSlabRecords * 4 + 2 * kSlabRecords * 2;
constexpr size_t kSlabFwdSmem = (size_t)kSlabMaxRows * 128;

template <bool EPI, bool IDX>
int launch_slab_backward(MsdaParams& p, const float* grad_out, float* grad_value, float* grad_loc, float* grad_attn,
                         cudaStream_t st, const char* who) {
  SlabGeo sg;
  int grid;
  slab_geo(p, sg, grid);
  CUtensorMap map;
  sg.use_tma = (slab_mode() >= 2 && make_slab_map(&map, grad_value, (long long)p.B * p.K, p.pix_stride)) ? 1 : 0;
  if (!sg.use_tma) memset(&map, 0, sizeof(map));
  auto k = msda_backward_slab_kernel<EPI, IDX>;
  if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSlabBwdSmem) != cudaSuccess)
    return set_error(VIDAR_E_CUDA, "%s: cannot opt in to %zu bytes of shared memory", who, kSlabBwdSmem);
  k<<<grid, kSlabWarps * 32, kSlabBwdSmem, st>>>(p, sg, map, grad_out, grad_value, grad_loc, grad_attn);
  return check_launch(who);
}

template <bool EPI, bool IDX>
int launch_slab_forward(MsdaParams& p, float* out, cudaStream_t st, const char* who) {
  SlabGeo sg;
  int grid;
  slab_geo(p, sg, grid);
  CUtensorMap map;
  sg.use_tma = (slab_mode() >= 2 && make_slab_map(&map, p.value, (long long)p.B * p.K, p.pix_stride)) ? 1 : 0;
  if (!sg.use_tma) memset(&map, 0, sizeof(map));
  auto k = msda_forward_slab_kernel<EPI, IDX>;
  if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSlabFwdSmem) != cudaSuccess)
    return set_error(VIDAR_E_CUDA, "%s: cannot opt in to %zu bytes of shared memory", who, kSlabFwdSmem);
  k<<<grid, kSlabWarps * 32, kSlabFwdSmem, st>>>(p, sg, map, out);
  return check_launch(who);
}

// VIDAR_MSDA_SLAB_FWD=1: the forward also runs its slab variant (off by default)
inline bool slab_forward_on() {
  static const bool v = [] {
    const char* e = getenv("VIDAR_MSDA_SLAB_FWD");
    return e && atoi(e) != 0;
  }();
  return v;
}

}  // namespace
}  // namespace vidar

using namespace vidar;

extern "C" int vidar_msda_forward(const float* value, const int64_t* spatial_shapes,
                                  const int64_t* level_start, const float* sampling_loc,
                                  const float* attn_weight, float* out, int B, int K, int H, int C,
                                  int L, int Q, int P, int im2col_step, void* stream) {
  MsdaParams p;
  int rc = fill_params(p, value, spatial_shapes, level_start, sampling_loc, attn_weight, B, K, H,
                       C, L, Q, P, im2col_step, "ms_deform_attn_forward");
  if (rc) return rc;
  VIDAR_REQUIRE(out, "ms_deform_attn_forward: null output");
  cudaStream_t st = (cudaStream_t)stream;
  if (slab_ok(p) && slab_forward_on()) return launch_slab_forward<false, false>(p, out, st, "ms_deform_attn_forward");
  if (vec_ok(p)) {
    const int CV = C / 4;
    set_ipw(p, CV);
    const long long nb = num_blocks(p);
    VIDAR_REQUIRE(nb < 2147483647LL, "ms_deform_attn_forward: problem too large");
    const dim3 grid((unsigned)nb), block(kWarpsPerBlock * 32);
#define VIDAR_FWD(CVV)                                                                     \
  if (p.ipw >= 32 / CVV) msda_forward_small_kernel<CVV><<<grid, block, 0, st>>>(p, out);  \
  else if (p.ipw > 1) msda_forward_kernel<CVV, true><<<grid, block, 0, st>>>(p, out);     \
  else msda_forward_kernel<CVV, false><<<grid, block, 0, st>>>(p, out)
    if (CV == 8) { VIDAR_FWD(8); }
    else if (CV == 4) { VIDAR_FWD(4); }
    else { VIDAR_FWD(16); }
#undef VIDAR_FWD
  } else {
    const long long n = p.items * C;
    const long long nb = (n + 255) / 256;
    VIDAR_REQUIRE(nb < 2147483647LL, "ms_deform_attn_forward: problem too large");
    msda_forward_generic_kernel<<<(unsigned)nb, 256, 0, st>>>(p, out);
  }
  return check_launch("ms_deform_attn_forward");
}

extern "C" int vidar_msda_backward(const float* value, const int64_t* spatial_shapes,
                                   const int64_t* level_start, const float* sampling_loc,
                                   const float* attn_weight, const float* grad_out,
                                   float* grad_value, float* grad_sampling_loc,
                                   float* grad_attn_weight, int B, int K, int H, int C, int L,
                                   int Q, int P, int im2col_step, void* stream) {
  MsdaParams p;
  int rc = fill_params(p, value, spatial_shapes, level_start, sampling_loc, attn_weight, B, K, H,
                       C, L, Q, P, im2col_step, "ms_deform_attn_backward");
  if (rc) return rc;
  VIDAR_REQUIRE(grad_out && grad_value && grad_sampling_loc && grad_attn_weight,
                "ms_deform_attn_backward: null gradient pointer");
  p.na_bytes = 0xffffffffu;
  cudaStream_t st = (cudaStream_t)stream;
  if (slab_ok(p))
    return launch_slab_backward<false, false>(p, grad_out, grad_value, grad_sampling_loc, grad_attn_weight, st,
                                              "ms_deform_attn_backward");
  if (vec_ok(p)) {
    const int CV = C / 4;
    set_ipw(p, CV);
    const long long nb = num_blocks(p);
    VIDAR_REQUIRE(nb < 2147483647LL, "ms_deform_attn_backward: problem too large");
    const dim3 grid((unsigned)nb), block(kWarpsPerBlock * 32);
#define VIDAR_BWD(CVV)                                                                        \
  if (p.ipw >= 32 / CVV)                                                                      \
    msda_backward_small_kernel<CVV><<<grid, block, 0, st>>>(p, grad_out, grad_value,         \
                                                            grad_sampling_loc, grad_attn_weight); \
  else if (p.ipw > 1)                                                                         \
    msda_backward_kernel<CVV, true><<<grid, block, 0, st>>>(p, grad_out, grad_value,          \
                                                            grad_sampling_loc, grad_attn_weight); \
  else                                                                                        \
    msda_backward_kernel<CVV, false><<<grid, block, 0, st>>>(p, grad_out, grad_value,         \
                                                             grad_sampling_loc, grad_attn_weight)
    if (CV == 8) { VIDAR_BWD(8); }
    else if (CV == 4) { VIDAR_BWD(4); }
    else { VIDAR_BWD(16); }
#undef VIDAR_BWD
  } else {
    cudaError_t e = cudaMemsetAsync(grad_sampling_loc, 0, sizeof(float) * 2 * (size_t)p.items * p.LP, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(grad_attn_weight, 0, sizeof(float) * (size_t)p.items * p.LP, st);
    if (e != cudaSuccess) return set_error(VIDAR_E_CUDA, "ms_deform_attn_backward: memset: %s", cudaGetErrorString(e));
    const long long n = p.items * C;
    const long long nb = (n + 255) / 256;
    VIDAR_REQUIRE(nb < 2147483647LL, "ms_deform_attn_backward: problem too large");
    msda_backward_generic_kernel<<<(unsigned)nb, 256, 0, st>>>(p, grad_out, grad_value, grad_sampling_loc, grad_attn_weight);
  }
  return check_launch("ms_deform_attn_backward");
}

// ---- MSDeformableAttention3D with the softmax and the sampling-location arithmetic folded in
static int check_sca(MsdaParams& p, const float* ref, int D, const char* who) {
  VIDAR_REQUIRE(ref, "%s: null reference points", who);
  VIDAR_REQUIRE(D > 0 && p.P % D == 0, "%s: num_points (%d) must be a multiple of the number of Z anchors (%d)", who, p.P, D);
  VIDAR_REQUIRE(p.LP == 32, "%s: the fused epilogue needs num_levels * num_points == 32 (got %d)", who, p.LP);
  VIDAR_REQUIRE(vec_ok(p) && (p.C == 16 || p.C == 32 || p.C == 64), "%s: head dim must be 16, 32 or 64 (got %d)", who, p.C);
  p.ref = ref;
  p.D = D;
  return VIDAR_OK;
}

extern "C" int vidar_msda_sca_forward(const float* value, const int64_t* spatial_shapes,
                                      const int64_t* level_start, const float* ref_points,
                                      const float* offsets, const float* logits, float* out, int B,
                                      int K, int H, int C, int L, int Q, int P, int D, void* stream) {
  const char* who = "msda_sca_forward";
  MsdaParams p;
  int rc = fill_params(p, value, spatial_shapes, level_start, offsets, logits, B, K, H, C, L, Q, P, B, who);
  if (rc) return rc;
  rc = check_sca(p, ref_points, D, who);
  if (rc) return rc;
  VIDAR_REQUIRE(out, "%s: null output", who);
  if (slab_ok(p) && slab_forward_on()) return launch_slab_forward<true, false>(p, out, (cudaStream_t)stream, who);
  set_ipw(p, C / 4);
  const long long nb = num_blocks(p);
  VIDAR_REQUIRE(nb < 2147483647LL, "%s: problem too large", who);
  const dim3 grid((unsigned)nb), block(kWarpsPerBlock * 32);
  cudaStream_t st = (cudaStream_t)stream;
  if (C == 32) msda_forward_kernel<8, false, true><<<grid, block, 0, st>>>(p, out);
  else if (C == 16) msda_forward_kernel<4, false, true><<<grid, block, 0, st>>>(p, out);
  else msda_forward_kernel<16, false, true><<<grid, block, 0, st>>>(p, out);
  return check_launch(who);
}

extern "C" int vidar_msda_sca_backward(const float* value, const int64_t* spatial_shapes,
                                       const int64_t* level_start, const float* ref_points,
                                       const float* offsets, const float* logits, const float* grad_out,
                                       float* grad_value, float* grad_offsets, float* grad_logits, int B,
                                       int K, int H, int C, int L, int Q, int P, int D, void* stream) {
  const char* who = "msda_sca_backward";
  MsdaParams p;
  int rc = fill_params(p, value, spatial_shapes, level_start, offsets, logits, B, K, H, C, L, Q, P, B, who);
  if (rc) return rc;
  rc = check_sca(p, ref_points, D, who);
  if (rc) return rc;
  VIDAR_REQUIRE(grad_out && grad_value && grad_offsets && grad_logits, "%s: null gradient pointer", who);
  p.na_bytes = 0xffffffffu;
  if (slab_ok(p)) return launch_slab_backward<true, false>(p, grad_out, grad_value, grad_offsets, grad_logits, (cudaStream_t)stream, who);
  set_ipw(p, C / 4);
  const long long nb = num_blocks(p);
  VIDAR_REQUIRE(nb < 2147483647LL, "%s: problem too large", who);
  const dim3 grid((unsigned)nb), block(kWarpsPerBlock * 32);
  cudaStream_t st = (cudaStream_t)stream;
  if (C == 32) msda_backward_kernel<8, false, true><<<grid, block, 0, st>>>(p, grad_out, grad_value, grad_offsets, grad_logits);
  else if (C == 16) msda_backward_kernel<4, false, true><<<grid, block, 0, st>>>(p, grad_out, grad_value, grad_offsets, grad_logits);
  else msda_backward_kernel<16, false, true><<<grid, block, 0, st>>>(p, grad_out, grad_value, grad_offsets, grad_logits);
  return check_launch(who);
}

// ---- row-indirect entry points: SpatialCrossAttention's rebatch / scatter-add fused into the op ------
// Rows are (camera, j) entries of per-camera visible-pillar lists (vidar_sca_compact); the forward adds
// each row's head vectors, scaled by 1/#cameras, straight into the BEV slot grid and the backward reads
// the slot gradient back through the same map (spatial_cross_attention.py:136-171 without the
// [cams, max_len, C] intermediates, the index_add_ and the host sync on max_len).
static int fill_rows(MsdaParams& p, const int32_t* idx, const int32_t* count, const float* inv, int bs, int ncl,
                     int cam0, int Qd, int S, int s_lo, int s_hi, const char* who) {
  VIDAR_REQUIRE(bs > 0 && ncl > 0 && cam0 >= 0 && Qd > 0, "%s: bad sizes bs=%d cameras=%d cam0=%d pillars=%d", who, bs, ncl, cam0, Qd);
  VIDAR_REQUIRE(p.B == bs * ncl, "%s: value batch (%d) must be bs * cameras (%d x %d)", who, p.B, bs, ncl);
  VIDAR_REQUIRE(p.Q <= Qd, "%s: rows per camera (%d) exceed the pillar count (%d)", who, p.Q, Qd);
  VIDAR_REQUIRE(idx || p.Q == Qd, "%s: without an index list every pillar is a row (rows %d != pillars %d)", who, p.Q, Qd);
  VIDAR_REQUIRE(S >= 1 && s_lo >= 0 && s_lo <= s_hi && s_hi <= S, "%s: bad sub-slice [%d, %d) of %d", who, s_lo, s_hi, S);
  VIDAR_REQUIRE(vec_ok(p), "%s: head dim must be 16, 32 or 64 (got %d)", who, p.C);
  p.idx = idx; p.count = count; p.inv = inv;
  p.bs = bs; p.ncl = ncl; p.cam0 = cam0; p.Qd = Qd;
  p.S = S; p.s_lo = s_lo; p.s_hi = s_hi;
  p.ipw = 1;
  p.lp_shift = 0;
  return VIDAR_OK;
}

#define VIDAR_ROWS_LAUNCH(KERNEL, EPI_, ...)                                                     \
  do {                                                                                           \
    const long long nb = num_blocks(p);                                                          \
    VIDAR_REQUIRE(nb < 2147483647LL, "%s: problem too large", who);                              \
    const dim3 grid((unsigned)nb), block(kWarpsPerBlock * 32);                                   \
    cudaStream_t st = (cudaStream_t)stream;                                                      \
    if (p.C == 32) KERNEL<8, false, EPI_, true><<<grid, block, 0, st>>>(__VA_ARGS__);            \
    else if (p.C == 16) KERNEL<4, false, EPI_, true><<<grid, block, 0, st>>>(__VA_ARGS__);       \
    else KERNEL<16, false, EPI_, true><<<grid, block, 0, st>>>(__VA_ARGS__);                     \
  } while (0)

extern "C" int vidar_msda_rows_forward(const float* value, const int64_t* spatial_shapes, const int64_t* level_start,
                                       const float* sampling_loc, const float* attn_weight, const int32_t* idx,
                                       const int32_t* count, const float* inv_count, float* slots, int bs, int ncl,
                                       int cam0, int K, int H, int C, int L, int Qrows, int Qd, int P, int S, int s_lo,
                                       int s_hi, void* stream) {
  const char* who = "msda_rows_forward";
  MsdaParams p;
  int rc = fill_params(p, value, spatial_shapes, level_start, sampling_loc, attn_weight, bs * ncl, K, H, C, L, Qrows, P,
                       bs * ncl, who);
  if (rc) return rc;
  rc = fill_rows(p, idx, count, inv_count, bs, ncl, cam0, Qd, S, s_lo, s_hi, who);
  if (rc) return rc;
  VIDAR_REQUIRE(slots, "%s: null output", who);
  // the streaming hint helps the plain forward (2.24 -> 2.06 ms) but not this one, whose epilogue also sends 245 MB
  // of slot reductions through L2 (2.25 -> 2.35 ms, gpurun r2g): plain loads here
  p.na_bytes = 0xffffffffu;
  if (slab_ok(p) && slab_forward_on()) return launch_slab_forward<false, true>(p, slots, (cudaStream_t)stream, who);
  VIDAR_ROWS_LAUNCH(msda_forward_kernel, false, p, slots);
  return check_launch(who);
}

extern "C" int vidar_msda_rows_backward(const float* value, const int64_t* spatial_shapes, const int64_t* level_start,
                                        const float* sampling_loc, const float* attn_weight, const int32_t* idx,
                                        const int32_t* count, const float* inv_count, const float* grad_slots,
                                        float* grad_value, float* grad_sampling_loc, float* grad_attn_weight, int bs,
                                        int ncl, int cam0, int K, int H, int C, int L, int Qrows, int Qd, int P, int S,
                                        int s_lo, int s_hi, void* stream) {
  const char* who = "msda_rows_backward";
  MsdaParams p;
  int rc = fill_params(p, value, spatial_shapes, level_start, sampling_loc, attn_weight, bs * ncl, K, H, C, L, Qrows, P,
                       bs * ncl, who);
  if (rc) return rc;
  rc = fill_rows(p, idx, count, inv_count, bs, ncl, cam0, Qd, S, s_lo, s_hi, who);
  if (rc) return rc;
  VIDAR_REQUIRE(grad_slots && grad_value && grad_sampling_loc && grad_attn_weight, "%s: null gradient pointer", who);
  p.na_bytes = 0xffffffffu;
  if (slab_ok(p))
    return launch_slab_backward<false, true>(p, grad_slots, grad_value, grad_sampling_loc, grad_attn_weight, (cudaStream_t)stream, who);
  VIDAR_ROWS_LAUNCH(msda_backward_kernel, false, p, grad_slots, grad_value, grad_sampling_loc, grad_attn_weight);
  return check_launch(who);
}

extern "C" int vidar_msda_sca_rows_forward(const float* value, const int64_t* spatial_shapes, const int64_t* level_start,
                                           const float* ref_cam, const float* offsets, const float* logits,
                                           const int32_t* idx, const int32_t* count, const float* inv_count, float* slots,
                                           int bs, int ncl, int cam0, int K, int H, int C, int L, int Qrows, int Qd, int P,
                                           int D, int S, int s_lo, int s_hi, void* stream) {
  const char* who = "msda_sca_rows_forward";
  MsdaParams p;
  int rc = fill_params(p, value, spatial_shapes, level_start, offsets, logits, bs * ncl, K, H, C, L, Qrows, P, bs * ncl, who);
  if (rc) return rc;
  rc = check_sca(p, ref_cam, D, who);
  if (rc) return rc;
  rc = fill_rows(p, idx, count, inv_count, bs, ncl, cam0, Qd, S, s_lo, s_hi, who);
  if (rc) return rc;
  VIDAR_REQUIRE(slots, "%s: null output", who);
  p.na_bytes = 0xffffffffu;
  if (slab_ok(p) && slab_forward_on()) return launch_slab_forward<true, true>(p, slots, (cudaStream_t)stream, who);
  VIDAR_ROWS_LAUNCH(msda_forward_kernel, true, p, slots);
  return check_launch(who);
}

extern "C" int vidar_msda_sca_rows_backward(const float* value, const int64_t* spatial_shapes, const int64_t* level_start,
                                            const float* ref_cam, const float* offsets, const float* logits,
                                            const int32_t* idx, const int32_t* count, const float* inv_count,
                                            const float* grad_slots, float* grad_value, float* grad_offsets,
                                            float* grad_logits, int bs, int ncl, int cam0, int K, int H, int C, int L,
                                            int Qrows, int Qd, int P, int D, int S, int s_lo, int s_hi, void* stream) {
  const char* who = "msda_sca_rows_backward";
  MsdaParams p;
  int rc = fill_params(p, value, spatial_shapes, level_start, offsets, logits, bs * ncl, K, H, C, L, Qrows, P, bs * ncl, who);
  if (rc) return rc;
  rc = check_sca(p, ref_cam, D, who);
  if (rc) return rc;
  rc = fill_rows(p, idx, count, inv_count, bs, ncl, cam0, Qd, S, s_lo, s_hi, who);
  if (rc) return rc;
  VIDAR_REQUIRE(grad_slots && grad_value && grad_offsets && grad_logits, "%s: null gradient pointer", who);
  p.na_bytes = 0xffffffffu;
  if (slab_ok(p))
    return launch_slab_backward<true, true>(p, grad_slots, grad_value, grad_offsets, grad_logits, (cudaStream_t)stream, who);
  VIDAR_ROWS_LAUNCH(msda_backward_kernel, true, p, grad_slots, grad_value, grad_offsets, grad_logits);
  return check_launch(who);
}
#undef VIDAR_ROWS_LAUNCH
