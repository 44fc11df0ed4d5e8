// Same-box CUDA baseline for multi-scale deformable attention with the REFERENCE-LINEAGE thread mapping
// (VERDICT r01 item 9b).  mmcv-full 1.4.0's kernel is not in the reference tree; its in-tree twin
// projects/mmdet3d_plugin/bevformer/backbones/ops_dcnv3/src/cuda/dcnv3_im2col_cuda.cuh:216-370 shows the
// mapping, which is what this file re-implements for the MSDA semantics of SURVEY.md A.1:
//   forward : one thread per output scalar (b, q, h, c); the thread walks all L*P samples, re-deriving the
//             sample geometry per channel, scalar loads, stride-C coalescing across a warp;
//   backward: one block of C threads per (b, q, h); per sample every thread writes its channel's partial
//             grad_loc / grad_attn to shared memory, __syncthreads, thread 0 sums the C partials serially,
//             __syncthreads; grad_value through 4 scalar atomicAdd per sample and channel.
// It is a measuring stick, not product code: built by tools/bench_ref_msda.py, never loaded by vidar_b200.
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

__device__ __forceinline__ float corner(const float* v, int H_, int W_, int stride, int y, int x) {
  return (y >= 0 && x >= 0 && y < H_ && x < W_) ? v[(size_t)(y * W_ + x) * stride] : 0.f;
}

__global__ void lineage_forward(const float* __restrict__ value, const int64_t* __restrict__ shapes,
                                const int64_t* __restrict__ lsi, const float* __restrict__ loc,
                                const float* __restrict__ attn, float* __restrict__ out, int B, int K, int H, int C,
                                int L, int Q, int P) {
  const long long n = (long long)B * Q * H * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int h = (int)((i / C) % H);
    const int q = (int)((i / ((long long)C * H)) % Q);
    const int b = (int)(i / ((long long)C * H * Q));
    const size_t item = ((size_t)b * Q + q) * H + h;
    float acc = 0.f;
    for (int l = 0; l < L; ++l) {
      const int Hl = (int)shapes[2 * l], Wl = (int)shapes[2 * l + 1];
      const float* v = value + ((size_t)b * K + lsi[l]) * H * C + (size_t)h * C + c;
      for (int p = 0; p < P; ++p) {
        const size_t s = (item * L + l) * P + p;
        const float x = loc[2 * s] * Wl - 0.5f, y = loc[2 * s + 1] * Hl - 0.5f;
        if (y > -1 && x > -1 && y < Hl && x < Wl) {
          const int y0 = (int)floorf(y), x0 = (int)floorf(x);
          const float ly = y - y0, lx = x - x0;
          const float val = (1 - ly) * (1 - lx) * corner(v, Hl, Wl, H * C, y0, x0) + (1 - ly) * lx * corner(v, Hl, Wl, H * C, y0, x0 + 1) +
                            ly * (1 - lx) * corner(v, Hl, Wl, H * C, y0 + 1, x0) + ly * lx * corner(v, Hl, Wl, H * C, y0 + 1, x0 + 1);
          acc += attn[s] * val;
        }
      }
    }
    out[i] = acc;
  }
}

__global__ void lineage_backward(const float* __restrict__ value, const int64_t* __restrict__ shapes,
                                 const int64_t* __restrict__ lsi, const float* __restrict__ loc,
                                 const float* __restrict__ attn, const float* __restrict__ grad_out,
                                 float* __restrict__ grad_value, float* __restrict__ grad_loc,
                                 float* __restrict__ grad_attn, int B, int K, int H, int C, int L, int Q, int P) {
  extern __shared__ float cache[];               // [3][C]: d/dx, d/dy, d/dattn partials of every channel
  const int c = threadIdx.x;
  const size_t item = blockIdx.x;                // (b, q, h)
  const int h = (int)(item % H);
  const int b = (int)(item / ((size_t)H * Q));
  const float go = grad_out[item * C + c];
  for (int l = 0; l < L; ++l) {
    const int Hl = (int)shapes[2 * l], Wl = (int)shapes[2 * l + 1];
    const size_t base = ((size_t)b * K + lsi[l]) * H * C + (size_t)h * C + c;
    const float* v = value + base;
    float* gv = grad_value + base;
    for (int p = 0; p < P; ++p) {
      const size_t s = (item * L + l) * P + p;
      const float x = loc[2 * s] * Wl - 0.5f, y = loc[2 * s + 1] * Hl - 0.5f;
      float gx = 0.f, gy = 0.f, ga = 0.f;
      if (y > -1 && x > -1 && y < Hl && x < Wl) {
        const int y0 = (int)floorf(y), x0 = (int)floorf(x);
        const float ly = y - y0, lx = x - x0, hy = 1 - ly, hx = 1 - lx;
        const float v1 = corner(v, Hl, Wl, H * C, y0, x0), v2 = corner(v, Hl, Wl, H * C, y0, x0 + 1);
        const float v3 = corner(v, Hl, Wl, H * C, y0 + 1, x0), v4 = corner(v, Hl, Wl, H * C, y0 + 1, x0 + 1);
        const float top = go * attn[s];
        const int st = H * C;
        if (y0 >= 0 && x0 >= 0) atomicAdd(gv + (size_t)(y0 * Wl + x0) * st, hy * hx * top);
        if (y0 >= 0 && x0 + 1 < Wl) atomicAdd(gv + (size_t)(y0 * Wl + x0 + 1) * st, hy * lx * top);
        if (y0 + 1 < Hl && x0 >= 0) atomicAdd(gv + (size_t)((y0 + 1) * Wl + x0) * st, ly * hx * top);
        if (y0 + 1 < Hl && x0 + 1 < Wl) atomicAdd(gv + (size_t)((y0 + 1) * Wl + x0 + 1) * st, ly * lx * top);
        gx = Wl * top * (hy * (v2 - v1) + ly * (v4 - v3));
        gy = Hl * top * (hx * (v3 - v1) + lx * (v4 - v2));
        ga = go * (hy * hx * v1 + hy * lx * v2 + ly * hx * v3 + ly * lx * v4);
      }
      cache[c] = gx;
      cache[C + c] = gy;
      cache[2 * C + c] = ga;
      __syncthreads();
      if (c == 0) {                              // the serial C-way sum of the lineage kernel
        float sx = 0.f, sy = 0.f, sa = 0.f;
        for (int t = 0; t < C; ++t) { sx += cache[t]; sy += cache[C + t]; sa += cache[2 * C + t]; }
        grad_loc[2 * s] = sx;
        grad_loc[2 * s + 1] = sy;
        grad_attn[s] = sa;
      }
      __syncthreads();
    }
  }
}

}  // namespace

extern "C" int lineage_msda_forward(const float* value, const int64_t* shapes, const int64_t* lsi, const float* loc,
                                    const float* attn, float* out, int B, int K, int H, int C, int L, int Q, int P,
                                    void* stream) {
  const long long n = (long long)B * Q * H * C;
  const int threads = 1024;                      // mmcv: CUDA_NUM_THREADS
  const long long blocks = (n + threads - 1) / threads;
  lineage_forward<<<(unsigned)(blocks < 2147483647LL ? blocks : 2147483647LL), threads, 0, (cudaStream_t)stream>>>(
      value, shapes, lsi, loc, attn, out, B, K, H, C, L, Q, P);
  return (int)cudaGetLastError();
}

extern "C" int lineage_msda_backward(const float* value, const int64_t* shapes, const int64_t* lsi, const float* loc,
                                     const float* attn, const float* grad_out, float* grad_value, float* grad_loc,
                                     float* grad_attn, int B, int K, int H, int C, int L, int Q, int P, void* stream) {
  const long long items = (long long)B * Q * H;
  lineage_backward<<<(unsigned)items, C, 3 * C * sizeof(float), (cudaStream_t)stream>>>(
      value, shapes, lsi, loc, attn, grad_out, grad_value, grad_loc, grad_attn, B, K, H, C, L, Q, P);
  return (int)cudaGetLastError();
}
