// Warp-level LayerNorm, prompt splice, token assembly, im2col, transposes and small fp32 linears
// for the MaPLe towers (HBM-bound kernels; SURVEY.md §2.2 K2,K3,K9,K10,K11,K12,K13).
#include "mfk_common.cuh"
#include "../../include/mfk.h"

namespace {
using namespace mfk;

constexpr int kLnWarps = 8;
#ifndef MFK_LNF_WARPS
#define MFK_LNF_WARPS 8
#endif
constexpr int kLnFwdWarps = MFK_LNF_WARPS;  // rows (= warps) per CTA of the forward kernel

// ============================================================================ LayerNorm forward
// clip/model.py:153-159: fp32 LayerNorm (eps 1e-5, biased variance). One warp per row,
// VEC float4 per lane (D = 128*VEC), two-pass statistics held in registers.
template <int VEC>
__device__ __forceinline__ void ln_row_stats(const float4 (&v)[VEC], float& mean, float& rstd, float eps) {
  constexpr float invD = 1.0f / (128.0f * VEC);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  mean = warp_sum(s) * invD;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  rstd = rsqrtf(warp_sum(q) * invD + eps);
}

template <int VEC>
__device__ __forceinline__ void ln_row_write(const float4 (&v)[VEC], float mean, float rstd, const float* gamma,
                                             const float* beta, bf16* y16, float* y32, int lane) {
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const int c4 = lane + 32 * i;
    float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c4);
    float4 b = __ldg(reinterpret_cast<const float4*>(beta) + c4);
    float4 o;
    o.x = (v[i].x - mean) * rstd * g.x + b.x;
    o.y = (v[i].y - mean) * rstd * g.y + b.y;
    o.z = (v[i].z - mean) * rstd * g.z + b.z;
    o.w = (v[i].w - mean) * rstd * g.w + b.w;
    if (y32) reinterpret_cast<float4*>(y32)[c4] = o;
    if (y16) reinterpret_cast<uint2*>(y16)[c4] = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
  }
}

template <int VEC>
__global__ void __launch_bounds__(kLnFwdWarps * 32)
ln_fwd_kernel(float* __restrict__ x, const int* __restrict__ rowidx, const float* __restrict__ gamma,
              const float* __restrict__ beta, bf16* __restrict__ y16, float* __restrict__ y32,
              float* __restrict__ xsave, float* __restrict__ mean_o, float* __restrict__ rstd_o, int M, float eps,
              const float* __restrict__ prompt, int T, int row0, int n_ctx) {
  constexpr int D = 128 * VEC;
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * kLnFwdWarps + (threadIdx.x >> 5);
  if (row >= M) return;
  const size_t src = rowidx ? (size_t)rowidx[row] : (size_t)row;
  float4* xr = reinterpret_cast<float4*>(x + src * D);
  float4 v[VEC];
  // fused deep-prompt splice (clip/model.py:320-349): the n_ctx prompt rows of every sequence are overwritten with
  // q16(prompt) on the way in, and written back so that the residual stream holds them too
  const int tpos = prompt ? row % T - row0 : -1;
  if (tpos >= 0 && tpos < n_ctx) {
    const float4* pr = reinterpret_cast<const float4*>(prompt + (size_t)tpos * D);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const float4 a = __ldg(pr + lane + 32 * i);
      v[i] = make_float4(q16(a.x), q16(a.y), q16(a.z), q16(a.w));
      xr[lane + 32 * i] = v[i];
    }
  } else {
#pragma unroll
    for (int i = 0; i < VEC; ++i) v[i] = xr[lane + 32 * i];
  }
  if (xsave) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) reinterpret_cast<float4*>(xsave + (size_t)row * D)[lane + 32 * i] = v[i];
  }
  float mean, rstd;
  ln_row_stats<VEC>(v, mean, rstd, eps);
  if (lane == 0) {
    if (mean_o) mean_o[row] = mean;
    if (rstd_o) rstd_o[row] = rstd;
  }
  ln_row_write<VEC>(v, mean, rstd, gamma, beta, y16 ? y16 + (size_t)row * D : nullptr,
                    y32 ? y32 + (size_t)row * D : nullptr, lane);
}

// ============================================================================ LayerNorm backward
// dx = rstd * (dy*g - mean_D(dy*g) - xhat * mean_D(dy*g*xhat));  g_out = g_in + dx.
// dgamma/dbeta: per-CTA partial column sums (deterministic two-stage reduction).
template <int VEC, bool DY_BF16>
__global__ void __launch_bounds__(kLnWarps * 32, 2)
ln_bwd_kernel(const void* __restrict__ dy_, const float* __restrict__ x, const float* __restrict__ mean_i,
              const float* __restrict__ rstd_i, const float* __restrict__ gamma, const float* g_in,
              float* g_out, bf16* __restrict__ g16, float* __restrict__ partial, int M,
              float* __restrict__ gprompt, int T, int row0, int n_ctx) {
  constexpr int D = 128 * VEC;
  // per-warp dgamma / dbeta accumulators live in shared memory (lane-private entries, no synchronisation inside the
  // row loop): the 48 registers they used to take now hold the prefetched residual gradient, so one row costs ONE
  // memory round trip (x, dy, g_in issued together) instead of two.
  __shared__ __align__(16) float acc[kLnWarps][2][D];
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  float4* mydg = reinterpret_cast<float4*>(&acc[warp][0][0]);
  float4* mydb = reinterpret_cast<float4*>(&acc[warp][1][0]);
  if (partial) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) mydg[lane + 32 * i] = mydb[lane + 32 * i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float4* gam4 = reinterpret_cast<const float4*>(gamma);

  for (int row = blockIdx.x * kLnWarps + warp; row < M; row += gridDim.x * kLnWarps) {
    float4 xh[VEC], d[VEC], r[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const int c4 = lane + 32 * i;
      xh[i] = reinterpret_cast<const float4*>(x + (size_t)row * D)[c4];
      if (DY_BF16) {
        uint2 u = reinterpret_cast<const uint2*>(static_cast<const bf16*>(dy_) + (size_t)row * D)[c4];
        float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y);
        d[i] = make_float4(a.x, a.y, b.x, b.y);
      } else {
        d[i] = reinterpret_cast<const float4*>(static_cast<const float*>(dy_) + (size_t)row * D)[c4];
      }
      r[i] = g_in ? reinterpret_cast<const float4*>(g_in + (size_t)row * D)[c4] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float mean = mean_i[row], rstd = rstd_i[row];
    int prow = -1;  // row of gprompt [N, n_ctx, D] if this is a spliced prompt position
    if (gprompt) {
      const int tp = row % T - row0;
      if (tp >= 0 && tp < n_ctx) prow = (row / T) * n_ctx + tp;
    }
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const int c4 = lane + 32 * i;
      xh[i] = make_float4((xh[i].x - mean) * rstd, (xh[i].y - mean) * rstd, (xh[i].z - mean) * rstd,
                          (xh[i].w - mean) * rstd);
      if (partial) {
        float4 a = mydg[c4], b = mydb[c4];
        a.x += d[i].x * xh[i].x; a.y += d[i].y * xh[i].y; a.z += d[i].z * xh[i].z; a.w += d[i].w * xh[i].w;
        b.x += d[i].x; b.y += d[i].y; b.z += d[i].z; b.w += d[i].w;
        mydg[c4] = a;
        mydb[c4] = b;
      }
      const float4 gm = __ldg(gam4 + c4);
      d[i].x *= gm.x; d[i].y *= gm.y; d[i].z *= gm.z; d[i].w *= gm.w;
      s1 += (d[i].x + d[i].y) + (d[i].z + d[i].w);
      s2 += (d[i].x * xh[i].x + d[i].y * xh[i].y) + (d[i].z * xh[i].z + d[i].w * xh[i].w);
    }
    s1 = warp_sum(s1) * (1.0f / D);
    s2 = warp_sum(s2) * (1.0f / D);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const int c4 = lane + 32 * i;
      float4 o;
      o.x = (d[i].x - s1 - xh[i].x * s2) * rstd;
      o.y = (d[i].y - s1 - xh[i].y * s2) * rstd;
      o.z = (d[i].z - s1 - xh[i].z * s2) * rstd;
      o.w = (d[i].w - s1 - xh[i].w * s2) * rstd;
      if (g_in) { o.x += r[i].x; o.y += r[i].y; o.z += r[i].z; o.w += r[i].w; }
      if (prow >= 0) {
        // fused backward of the deep-prompt splice: this row was overwritten by the prompt in the forward pass, so
        // its gradient belongs to the prompt (summed over the batch later, in batch order) and nothing flows on
        reinterpret_cast<float4*>(gprompt + (size_t)prow * D)[c4] = o;
        o = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      reinterpret_cast<float4*>(g_out + (size_t)row * D)[c4] = o;
      if (g16) reinterpret_cast<uint2*>(g16 + (size_t)row * D)[c4] = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
    }
  }
  if (!partial) return;
  // cross-warp reduction of the per-warp column sums (fixed order over warps)
  __syncthreads();
  float* pg = partial + (size_t)blockIdx.x * 2 * D;
  for (int c = threadIdx.x; c < 2 * D; c += kLnWarps * 32) {
    const int half = c / D, col = c % D;
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < kLnWarps; ++w) a += acc[w][half][col];
    pg[c] = a;
  }
}

// out[c] (+)= sum_p partial[p, c]. block (32, 32): 32 row groups sum P/32 partials each (strided, fixed order),
// then a fixed-order 32-way combine. Two outputs (dgamma | dbeta) are handled by blockIdx.y.
__global__ void partial_reduce_kernel(const float* __restrict__ partial, int P, int N, long long pstride,
                                      long long ystride, float* __restrict__ out0, float* __restrict__ out1,
                                      int accumulate) {
  __shared__ float red[32][33];
  pdl_trigger();
  pdl_wait();
  const int c = blockIdx.x * 32 + threadIdx.x;
  float* out = blockIdx.y == 0 ? out0 : out1;
  const float* src = partial + (size_t)blockIdx.y * ystride;
  float s = 0.f;
  if (c < N && out)
    for (int p = threadIdx.y; p < P; p += 32) s += src[(size_t)p * pstride + c];
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < N && out) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 32; ++w) t += red[w][threadIdx.x];
    out[c] = accumulate ? out[c] + t : t;
  }
}

// ============================================================================ column sums (bias grads)
template <bool IN_BF16>
__global__ void colsum_partial_kernel(const void* __restrict__ x_, long long ld, int M, int N,
                                      float* __restrict__ partial) {
  pdl_trigger();
  pdl_wait();
  // block (32, 8): 64 columns per block.x, rows strided by 8*gridDim.y
  const int c = (blockIdx.x * 32 + threadIdx.x) * 2;
  __shared__ float red[8][64];
  float a = 0.f, b = 0.f;
  if (c < N) {
    for (int r = blockIdx.y * 8 + threadIdx.y; r < M; r += gridDim.y * 8) {
      if (IN_BF16) {
        uint32_t u = *reinterpret_cast<const uint32_t*>(static_cast<const bf16*>(x_) + (size_t)r * ld + c);
        float2 f = unpack_bf16(u);
        a += f.x; b += f.y;
      } else {
        float2 f = *reinterpret_cast<const float2*>(static_cast<const float*>(x_) + (size_t)r * ld + c);
        a += f.x; b += f.y;
      }
    }
  }
  red[threadIdx.y][threadIdx.x * 2] = a;
  red[threadIdx.y][threadIdx.x * 2 + 1] = b;
  __syncthreads();
  if (threadIdx.y == 0 && c < N) {
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) { s0 += red[w][threadIdx.x * 2]; s1 += red[w][threadIdx.x * 2 + 1]; }
    partial[(size_t)blockIdx.y * N + c] = s0;
    partial[(size_t)blockIdx.y * N + c + 1] = s1;
  }
}

// ============================================================================ im2col for conv1 16x16/s16
// clip/model.py:514-518: conv(3->768, k=16, s=16, no bias) == [B*196, 768] x [768, 768]^T with
// K index = c*256 + ky*16 + kx, patch p = gy*14 + gx.
__global__ void im2col16_kernel(const float* __restrict__ img, bf16* __restrict__ out, int B, int S) {
  const int G = S / 16;
  const int patch = blockIdx.x;             // b*G*G + gy*G + gx
  const int b = patch / (G * G), pp = patch % (G * G), gy = pp / G, gx = pp % G;
  // 768 K-values = 3 channels x 16 rows x 16 px; thread handles 4 consecutive px
  for (int t = threadIdx.x; t < 192; t += blockDim.x) {
    const int c = t / 64, ky = (t % 64) / 4, kx = (t % 4) * 4;
    const float4 v = *reinterpret_cast<const float4*>(img + (((size_t)b * 3 + c) * S + gy * 16 + ky) * S + gx * 16 + kx);
    *reinterpret_cast<uint2*>(out + (size_t)patch * 768 + c * 256 + ky * 16 + kx) =
        make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
  }
}

// ============================================================================ vision token assembly + ln_pre
// clip/model.py:522-544: [cls+pos0 ; patches+pos ; q16(shared_ctx)] -> ln_pre. One warp per token row.
template <int VEC>
__global__ void __launch_bounds__(kLnWarps * 32)
vis_assemble_kernel(const float* __restrict__ tok, const float* __restrict__ cls, const float* __restrict__ pos,
                    const float* __restrict__ shared_ctx, const float* __restrict__ gamma,
                    const float* __restrict__ beta, float* __restrict__ x0, float* __restrict__ x,
                    float* __restrict__ mean_o, float* __restrict__ rstd_o, int B, int T, int n_ctx, float eps) {
  constexpr int D = 128 * VEC;
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * kLnWarps + (threadIdx.x >> 5);
  if (row >= B * T) return;
  const int b = row / T, t = row % T;
  const int P = T - n_ctx - 1;  // patches
  float4 v[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const int c4 = lane + 32 * i;
    if (t == 0) {
      float4 a = __ldg(reinterpret_cast<const float4*>(cls) + c4), p = __ldg(reinterpret_cast<const float4*>(pos) + c4);
      v[i] = make_float4(a.x + p.x, a.y + p.y, a.z + p.z, a.w + p.w);
    } else if (t <= P) {
      float4 a = reinterpret_cast<const float4*>(tok + ((size_t)b * P + (t - 1)) * D)[c4];
      float4 p = __ldg(reinterpret_cast<const float4*>(pos + (size_t)t * D) + c4);
      v[i] = make_float4(a.x + p.x, a.y + p.y, a.z + p.z, a.w + p.w);
    } else {
      float4 a = __ldg(reinterpret_cast<const float4*>(shared_ctx + (size_t)(t - P - 1) * D) + c4);
      v[i] = make_float4(q16(a.x), q16(a.y), q16(a.z), q16(a.w));
    }
    if (x0) reinterpret_cast<float4*>(x0 + (size_t)row * D)[c4] = v[i];
  }
  float mean, rstd;
  ln_row_stats<VEC>(v, mean, rstd, eps);
  if (lane == 0) {
    if (mean_o) mean_o[row] = mean;
    if (rstd_o) rstd_o[row] = rstd;
  }
  ln_row_write<VEC>(v, mean, rstd, gamma, beta, nullptr, x + (size_t)row * D, lane);
}

// ============================================================================ text prompt assembly
// trainers/maple.py:157-166,181-187,54: cat(prefix, ctx, suffix) + positional_embedding, truncated to Te rows.
__global__ void text_assemble_kernel(const float* __restrict__ prefix, const float* __restrict__ ctx,
                                     const float* __restrict__ suffix, const float* __restrict__ pos,
                                     float* __restrict__ x, int C, int Te, int n_ctx, int Tfull, int D) {
  const int row = blockIdx.x;  // c*Te + t
  const int c = row / Te, t = row % Te;
  const float* src;
  if (t == 0) src = prefix + (size_t)c * D;
  else if (t <= n_ctx) src = ctx + (size_t)(t - 1) * D;
  else src = suffix + ((size_t)c * (Tfull - 1 - n_ctx) + (t - 1 - n_ctx)) * D;
  for (int i = threadIdx.x; i < D / 4; i += blockDim.x) {
    float4 a = __ldg(reinterpret_cast<const float4*>(src) + i);
    float4 p = __ldg(reinterpret_cast<const float4*>(pos + (size_t)t * D) + i);
    reinterpret_cast<float4*>(x + (size_t)row * D)[i] = make_float4(a.x + p.x, a.y + p.y, a.z + p.z, a.w + p.w);
  }
}

// ============================================================================ deep prompt splice
// clip/model.py:320-349: overwrite n_ctx rows of every sequence with q16(prompt).
__global__ void splice_fwd_kernel(float* __restrict__ x, const float* __restrict__ prompt, int T, int row0, int n_ctx,
                                  int D) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.x / n_ctx, j = blockIdx.x % n_ctx;
  float* dst = x + ((size_t)b * T + row0 + j) * D;
  const float* src = prompt + (size_t)j * D;
  for (int i = threadIdx.x; i < D / 4; i += blockDim.x) {
    float4 a = __ldg(reinterpret_cast<const float4*>(src) + i);
    reinterpret_cast<float4*>(dst)[i] = make_float4(q16(a.x), q16(a.y), q16(a.z), q16(a.w));
  }
}
// backward: dprompt[j,:] = sum_b q16?(g[b,row0+j,:]) in batch order; optionally zero those rows of g / g16.
// block (32 columns, 8 batch groups): group y sums sequences y, y+8, ... in order; groups are combined in order
// 0..7 (for N <= 8 this is the plain sequential batch order of autograd's expand-backward).
__global__ void splice_bwd_kernel(float* __restrict__ g, bf16* __restrict__ g16, float* __restrict__ dprompt, int N,
                                  int T, int row0, int n_ctx, int D, int round16, int zero, long long g_stride,
                                  long long dp_stride) {
  pdl_trigger();
  pdl_wait();
  g += (size_t)blockIdx.z * g_stride;  // batched over layers (blockIdx.z): one launch for every spliced layer
  dprompt += (size_t)blockIdx.z * dp_stride;
  __shared__ float red[8][33];
  const int j = blockIdx.y;
  const int c = blockIdx.x * 32 + threadIdx.x;
  float s = 0.f;
  if (c < D) {
    for (int b = threadIdx.y; b < N; b += 8) {
      const size_t idx = ((size_t)b * T + row0 + j) * D + c;
      const float v = g[idx];
      s += round16 ? q16(v) : v;
      if (zero) {
        g[idx] = 0.f;
        if (g16) g16[idx] = __float2bfloat16_rn(0.f);
      }
    }
  }
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < D) {
    float t = red[0][threadIdx.x];
#pragma unroll
    for (int w = 1; w < 8; ++w) t += red[w][threadIdx.x];
    dprompt[(size_t)j * D + c] = t;
  }
}

// g[rowidx[r], :] = dx[r, :] (+ bf16 copy); g must have been zero-filled.
__global__ void scatter_rows_kernel(const float* __restrict__ dx, const int* __restrict__ rowidx,
                                    float* __restrict__ g, bf16* __restrict__ g16, int D) {
  pdl_trigger();
  pdl_wait();
  const int r = blockIdx.x;
  const size_t dst = (size_t)rowidx[r] * D;
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    const float v = dx[(size_t)r * D + i];
    g[dst + i] = v;
    if (g16) g16[dst + i] = __float2bfloat16_rn(v);
  }
}

// Dense form: g[m, :] = dx[n, :] if m == rowidx[n] (n = m / T: exactly one consumed row per sequence) else 0, for ALL
// N*T rows in one pass of float4 stores — replaces two full-buffer fills + the sparse scatter at the last block.
__global__ void scatter_rows_dense_kernel(const float4* __restrict__ dx, const int* __restrict__ rowidx,
                                          float4* __restrict__ g, int rows, int T, int D4) {
  pdl_trigger();
  pdl_wait();
  for (int m = blockIdx.x; m < rows; m += gridDim.x) {
    const int n = m / T;
    const bool hit = rowidx[n] == m;
    for (int i = threadIdx.x; i < D4; i += blockDim.x)
      g[(size_t)m * D4 + i] = hit ? dx[(size_t)n * D4 + i] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// dst[r, :] = src[rowidx[r], :] in 16-byte chunks (row_bytes % 16 == 0)
__global__ void gather_rows_kernel(const uint4* __restrict__ src, const int* __restrict__ rowidx,
                                   uint4* __restrict__ dst, int chunks_per_row) {
  pdl_trigger();
  pdl_wait();
  const int r = blockIdx.x;
  const size_t s0 = (size_t)rowidx[r] * chunks_per_row, d0 = (size_t)r * chunks_per_row;
  for (int i = threadIdx.x; i < chunks_per_row; i += blockDim.x) dst[d0 + i] = src[s0 + i];
}
// dst[rowidx[r], :] = src[r, :] (bf16 rows); dst must have been zero-filled
__global__ void scatter_rows_bf16_kernel(const uint4* __restrict__ src, const int* __restrict__ rowidx,
                                         uint4* __restrict__ dst, int chunks_per_row) {
  pdl_trigger();
  pdl_wait();
  const int r = blockIdx.x;
  const size_t d0 = (size_t)rowidx[r] * chunks_per_row, s0 = (size_t)r * chunks_per_row;
  for (int i = threadIdx.x; i < chunks_per_row; i += blockDim.x) dst[d0 + i] = src[s0 + i];
}

// ============================================================================ transposes / packing
template <typename TIN>
__global__ void transpose_to_bf16_kernel(const TIN* __restrict__ in, long long ldi, bf16* __restrict__ out,
                                         long long ldo, bf16* __restrict__ copy, long long ldc, int M, int N) {
  __shared__ float tile[32][33];
  const int n0 = blockIdx.x * 32, m0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int m = m0 + i, n = n0 + threadIdx.x;
    float v = 0.f;
    if (m < M && n < N) {
      v = (float)in[(size_t)m * ldi + n];
      if (copy) copy[(size_t)m * ldc + n] = __float2bfloat16_rn(v);
    }
    tile[i][threadIdx.x] = v;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int n = n0 + i, m = m0 + threadIdx.x;
    if (n < N && m < M) out[(size_t)n * ldo + m] = __float2bfloat16_rn(tile[threadIdx.x][i]);
  }
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ in, bf16* __restrict__ out, long long n) {
  long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    float4 v = *reinterpret_cast<const float4*>(in + i);
    *reinterpret_cast<uint2*>(out + i) = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
  } else {
    for (; i < n; ++i) out[i] = __float2bfloat16_rn(in[i]);
  }
}

// x = hi + lo with hi = bf16(x), lo = bf16(x - hi): out[r] = [hi | lo | hi] (K-concatenated A operand). Against a
// weight packed as [hi | hi | lo] one bf16 GEMM of depth 3K yields hi*hi + lo*hi + hi*lo, i.e. ~16 mantissa bits —
// used for the two feature heads, whose rounding error would otherwise dominate the logit error.
__global__ void split_bf16x3_kernel(const float* __restrict__ x, bf16* __restrict__ out, int rows, int D) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)rows * D) return;
  const int r = (int)(i / D), c = (int)(i % D);
  const float v = x[i];
  const bf16 hi = __float2bfloat16_rn(v);
  const bf16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
  bf16* o = out + (size_t)r * 3 * D;
  o[c] = hi;
  o[D + c] = lo;
  o[2 * D + c] = hi;
}

// QuickGELU in fp32 (exact sigmoid) followed by the hi | lo | hi split: the activation of the fp32 mode.
__global__ void quickgelu_split_bf16x3_kernel(const float* __restrict__ u, bf16* __restrict__ out, int rows, int D) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)rows * D) return;
  const int r = (int)(i / D), c = (int)(i % D);
  const float x = u[i];
  const float v = x / (1.0f + expf(-1.702f * x));
  const bf16 hi = __float2bfloat16_rn(v);
  const bf16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
  bf16* o = out + (size_t)r * 3 * D;
  o[c] = hi;
  o[D + c] = lo;
  o[2 * D + c] = hi;
}
// The B-side packing of the split-precision product for ACTIVATIONS: out[r] = [hi | hi | lo]. A wgrad
// dW = dY^T X of the fp32 training mode contracts over the rows: a [rows, 3D] split buffer read as [3 rows, D]
// (row 3r + j = part j of row r) pairs [hi | lo | hi] of one operand with [hi | hi | lo] of the other.
__global__ void split_bf16x3_rhs_kernel(const float* __restrict__ x, bf16* __restrict__ out, int rows, int D) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)rows * D) return;
  const int r = (int)(i / D), c = (int)(i % D);
  const float v = x[i];
  const bf16 hi = __float2bfloat16_rn(v);
  const bf16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
  bf16* o = out + (size_t)r * 3 * D;
  o[c] = hi;
  o[D + c] = hi;
  o[2 * D + c] = lo;
}
// Backward of QuickGELU in fp32 (exact sigmoid): du = dact * s (1 + 1.702 u (1 - s)), s = sigmoid(1.702 u).
__global__ void dquickgelu_mul_f32_kernel(const float* __restrict__ dact, const float* __restrict__ u,
                                          float* __restrict__ du, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = u[i];
  const float s = 1.0f / (1.0f + expf(-1.702f * x));
  du[i] = dact[i] * (s * (1.0f + 1.702f * x * (1.0f - s)));
}
// im2col of the 16x16 / stride-16 patch embedding kept in fp32 (fp32 mode: split into bf16x3 afterwards)
__global__ void im2col16_f32_kernel(const float* __restrict__ img, float* __restrict__ out, int B, int S) {
  const int G = S / 16;
  const int patch = blockIdx.x;
  const int b = patch / (G * G), pp = patch % (G * G), gy = pp / G, gx = pp % G;
  for (int t = threadIdx.x; t < 192; t += blockDim.x) {
    const int c = t / 64, ky = (t % 64) / 4, kx = (t % 4) * 4;
    const float4 v = *reinterpret_cast<const float4*>(img + (((size_t)b * 3 + c) * S + gy * 16 + ky) * S + gx * 16 + kx);
    *reinterpret_cast<float4*>(out + (size_t)patch * 768 + c * 256 + ky * 16 + kx) = v;
  }
}

// ============================================================================ small fp32 linears (prompt learner)
// trainers/maple.py:194-215: y[m,N] = x[m,K] W[N,K]^T + b, m = n_ctx (tiny). One warp per output column.
__global__ void linear_small_fwd_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                        const float* __restrict__ b, float* __restrict__ y, int m, int N, int K) {
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  for (int r = 0; r < m; ++r) {
    float s = 0.f;
    for (int k = lane; k < K; k += 32) s += x[(size_t)r * K + k] * W[(size_t)n * K + k];
    s = warp_sum(s);
    if (lane == 0) y[(size_t)r * N + n] = s + (b ? b[n] : 0.f);
  }
}
// dW[N,K] = dy^T x ; db[N] = sum_r dy ; dx[m,K] (+)= dy W
__global__ void linear_small_bwd_w_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                          float* __restrict__ dW, float* __restrict__ db, int m, int N, int K) {
  const int n = blockIdx.x;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    float s = 0.f;
    for (int r = 0; r < m; ++r) s += dy[(size_t)r * N + n] * x[(size_t)r * K + k];
    dW[(size_t)n * K + k] = s;
  }
  if (threadIdx.x == 0 && db) {
    float s = 0.f;
    for (int r = 0; r < m; ++r) s += dy[(size_t)r * N + n];
    db[n] = s;
  }
}
// block (32, 32): 32 consecutive k per block, 32 groups stride over n; fixed-order combine. grid (K/32, m).
__global__ void linear_small_bwd_x_kernel(const float* __restrict__ W, const float* __restrict__ dy,
                                          const float* __restrict__ add, float* __restrict__ dx, int m, int N, int K) {
  __shared__ float red[32][33];
  const int k = blockIdx.x * 32 + threadIdx.x, r = blockIdx.y;
  float s = 0.f;
  if (k < K)
    for (int n = threadIdx.y; n < N; n += 32) s += dy[(size_t)r * N + n] * W[(size_t)n * K + k];
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && k < K) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 32; ++w) t += red[w][threadIdx.x];
    dx[(size_t)r * K + k] = t + (add ? add[(size_t)r * K + k] : 0.f);
  }
}


// ---- grouped forms: ONE launch for all prompt-learner projections of a step (9 forward problems, 9 backward
// problems with m = n_ctx rows): the per-launch latency of 27 tiny kernels was a visible slice of the step's serial
// head and tail. The problem table lives in device memory (pointers into the parameter / gradient arenas are stable).
struct SmallLinearProblem {  // mirrors mfk.h
  const float* x; const float* W; const float* b; float* y;
  const float* dy; float* dW; float* db; const float* dx_add; float* dx;
  int m, N, K, pad;
};
__global__ void linear_small_fwd_grouped_kernel(const SmallLinearProblem* __restrict__ tab) {
  const SmallLinearProblem pr = tab[blockIdx.y];
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (n >= pr.N) return;
  for (int r = 0; r < pr.m; ++r) {
    float s = 0.f;
    for (int k = lane; k < pr.K; k += 32) s += pr.x[(size_t)r * pr.K + k] * pr.W[(size_t)n * pr.K + k];
    s = warp_sum(s);
    if (lane == 0) pr.y[(size_t)r * pr.N + n] = s + (pr.b ? pr.b[n] : 0.f);
  }
}
__global__ void linear_small_bwd_w_grouped_kernel(const SmallLinearProblem* __restrict__ tab) {
  const SmallLinearProblem pr = tab[blockIdx.y];
  const int n = blockIdx.x;
  if (n >= pr.N || !pr.dW) return;
  for (int k = threadIdx.x; k < pr.K; k += blockDim.x) {
    float s = 0.f;
    for (int r = 0; r < pr.m; ++r) s += pr.dy[(size_t)r * pr.N + n] * pr.x[(size_t)r * pr.K + k];
    pr.dW[(size_t)n * pr.K + k] = s;
  }
  if (threadIdx.x == 0 && pr.db) {
    float s = 0.f;
    for (int r = 0; r < pr.m; ++r) s += pr.dy[(size_t)r * pr.N + n];
    pr.db[n] = s;
  }
}
// grid (ceil(maxK/32), max m, problems), block (32, 32): same reduction order as linear_small_bwd_x_kernel
__global__ void linear_small_bwd_x_grouped_kernel(const SmallLinearProblem* __restrict__ tab) {
  __shared__ float red[32][33];
  const SmallLinearProblem pr = tab[blockIdx.z];
  const int k = blockIdx.x * 32 + threadIdx.x, r = blockIdx.y;
  if (r >= pr.m || !pr.dx || blockIdx.x * 32 >= pr.K) return;  // uniform per block
  float s = 0.f;
  if (k < pr.K)
    for (int n = threadIdx.y; n < pr.N; n += 32) s += pr.dy[(size_t)r * pr.N + n] * pr.W[(size_t)n * pr.K + k];
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && k < pr.K) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 32; ++w) t += red[w][threadIdx.x];
    pr.dx[(size_t)r * pr.K + k] = t + (pr.dx_add ? pr.dx_add[(size_t)r * pr.K + k] : 0.f);
  }
}

// ---- grouped second stage of the LayerNorm dgamma / dbeta reductions: the per-CTA partials of every LayerNorm
// backward of a tower are reduced by ONE launch at the end of the tower (the 4.5 us dependent launch after each
// of the 26 LayerNorm backwards sat on the critical path). Same summation order as partial_reduce_kernel.
struct PartialReduceProblem {  // mirrors mfk.h
  const float* partial; int P, N; float* out0; float* out1; int accumulate, pad;
};
__global__ void partial_reduce_grouped_kernel(const PartialReduceProblem* __restrict__ tab) {
  __shared__ float red[32][33];
  pdl_trigger();
  pdl_wait();  // launched with the PDL attribute: the LayerNorm backwards before it must have completed
  const PartialReduceProblem pr = tab[blockIdx.z];
  const int c = blockIdx.x * 32 + threadIdx.x;
  if (blockIdx.x * 32 >= pr.N) return;  // uniform per block
  float* out = blockIdx.y == 0 ? pr.out0 : pr.out1;
  const float* src = pr.partial + (size_t)blockIdx.y * pr.N;
  float s = 0.f;
  if (c < pr.N && out)
    for (int p = threadIdx.y; p < pr.P; p += 32) s += src[(size_t)p * 2 * pr.N + c];
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < pr.N && out) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 32; ++w) t += red[w][threadIdx.x];
    out[c] = pr.accumulate ? out[c] + t : t;
  }
}

// ---- grouped fp32 [M,N] -> bf16 copy + bf16 transpose (refresh of the trainable block weights after an update)
struct RepackProblem {  // mirrors mfk.h
  const float* in; bf16* out_t; bf16* copy; int M, N;
  int tile0, tiles_n;  // first global 64x64 tile index of this problem; its tiles per row
};
// Rounding of the trainable masters to bf16 is DITHERED with a fixed per-element threshold (hash of the element index)
// instead of round-to-nearest-even: the reference's weights sit on the fp16 grid, i.e. on only 8 positions inside a
// bf16 interval, 1/8 of them exactly on the tie. Updates at the yaml's LR (2.6e-3 x clipped gradients) are far below
// one bf16 ulp, so under RNE the copy seen by the forward GEMMs would not move at all for 7/8 of the elements and
// jump half an ulp IN THE SIGN OF THE UPDATE for the tie elements — a coherent over-shoot of the whole update
// (measured: eval logits 2.1e-2 from the fp32 mode at the same parameters after three lr=0.0026 steps, image-feature
// error 33 % aligned with the text features, against 2e-3 / 6 % at initialisation). A fixed uniform threshold makes
// the expected rounded value equal to the master for any grid, so small updates are carried by the right fraction
// of elements. Deterministic (no RNG state), same value in the copy and in the transpose.
__device__ __forceinline__ float dither_bf16(float v, uint32_t idx) {
  uint32_t h = idx * 2654435761u;
  h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
  const uint32_t b = __float_as_uint(v);
  if ((b & 0x7F800000u) == 0x7F800000u) return v;  // Inf / NaN pass through
  return __uint_as_float((b + (h >> 16)) & 0xFFFF0000u);
}
// grid.x = total number of 64x64 tiles over all problems: every block finds its problem by a short scan.
// block (32, 8): float2 loads / bf16x2 stores (128-byte warp rows both for the copy and for the transpose).
__global__ void repack_grouped_kernel(const RepackProblem* __restrict__ tab, int n_problems) {
  __shared__ float tile[64][65];
  int pi = 0;
  while (pi + 1 < n_problems && (int)blockIdx.x >= tab[pi + 1].tile0) ++pi;
  const RepackProblem pr = tab[pi];
  const int t = (int)blockIdx.x - pr.tile0;
  const int n0 = (t % pr.tiles_n) * 64, m0 = (t / pr.tiles_n) * 64;
  if (m0 >= pr.M) return;  // uniform per block
  const bool even = (pr.N & 1) == 0 && (pr.M & 1) == 0;
  for (int i = threadIdx.y; i < 64; i += blockDim.y) {
    const int m = m0 + i, n = n0 + 2 * threadIdx.x;
    float v0 = 0.f, v1 = 0.f;
    if (m < pr.M) {
      const uint32_t e = (uint32_t)pi * 0x9E3779B9u + (uint32_t)m * (uint32_t)pr.N + (uint32_t)n;
      if (even && n + 1 < pr.N) {
        const float2 v = *reinterpret_cast<const float2*>(pr.in + (size_t)m * pr.N + n);
        v0 = dither_bf16(v.x, e); v1 = dither_bf16(v.y, e + 1);   // exactly bf16-representable from here on
        if (pr.copy) *reinterpret_cast<uint32_t*>(pr.copy + (size_t)m * pr.N + n) = pack_bf16(v0, v1);
      } else {
        if (n < pr.N) { v0 = dither_bf16(pr.in[(size_t)m * pr.N + n], e); if (pr.copy) pr.copy[(size_t)m * pr.N + n] = __float2bfloat16_rn(v0); }
        if (n + 1 < pr.N) { v1 = dither_bf16(pr.in[(size_t)m * pr.N + n + 1], e + 1); if (pr.copy) pr.copy[(size_t)m * pr.N + n + 1] = __float2bfloat16_rn(v1); }
      }
    }
    tile[i][2 * threadIdx.x] = v0;
    tile[i][2 * threadIdx.x + 1] = v1;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 64; i += blockDim.y) {
    const int n = n0 + i, m = m0 + 2 * threadIdx.x;
    if (n >= pr.N) continue;
    const float v0 = tile[2 * threadIdx.x][i], v1 = tile[2 * threadIdx.x + 1][i];
    if (even && m + 1 < pr.M) {
      *reinterpret_cast<uint32_t*>(pr.out_t + (size_t)n * pr.M + m) = pack_bf16(v0, v1);
    } else {
      if (m < pr.M) pr.out_t[(size_t)n * pr.M + m] = __float2bfloat16_rn(v0);
      if (m + 1 < pr.M) pr.out_t[(size_t)n * pr.M + m + 1] = __float2bfloat16_rn(v1);
    }
  }
}


// ============================================================================ training-time augmentation on the GPU
// Dassl's build_transform for the reference's yaml (configs/trainers/MaPLe/*.yaml:8-13): RandomResizedCrop(224,
// bicubic) -> RandomHorizontalFlip -> ToTensor -> Normalize(mean, std). The random draws (crop box, flip) stay on the
// host (cheap, torch RNG: trainers/client_datamanager.py); this kernel does the pixel work for a whole batch: antialiased
// bicubic resampling (a = -0.5, PIL / torch `antialias=True` weights) of the crop box to S x S, rounding to the
// uint8 grid like the PIL / uint8-tensor pipeline does, mirror, /255 and normalisation. src uint8 [B,3,H,W].
__device__ __forceinline__ float cubic_aa(float x) {
  constexpr float a = -0.5f;
  x = fabsf(x);
  if (x < 1.f) return ((a + 2.f) * x - (a + 3.f)) * x * x + 1.f;
  if (x < 2.f) return (((x - 5.f) * x + 8.f) * x - 4.f) * a;
  return 0.f;
}
__device__ __forceinline__ void aa_window(float scale, int o, int in_size, int& lo, int& n, float& center,
                                          float& invscale, float& total) {
  const float support = scale >= 1.f ? 2.f * scale : 2.f;
  center = scale * (o + 0.5f);
  invscale = scale >= 1.f ? 1.f / scale : 1.f;
  lo = max((int)(center - support + 0.5f), 0);
  n = min((int)(center + support + 0.5f), in_size) - lo;
  total = 0.f;
  for (int j = 0; j < n; ++j) total += cubic_aa((j + lo - center + 0.5f) * invscale);
}
__global__ void rrc_flip_normalize_kernel(const uint8_t* __restrict__ src, int H, int W, const int* __restrict__ boxes,
                                          const uint8_t* __restrict__ flip, const float* __restrict__ mean,
                                          const float* __restrict__ stdv, float* __restrict__ out, int S,
                                          int round_u8) {
  const int b = blockIdx.z;
  const int ox = blockIdx.x * blockDim.x + threadIdx.x, oy = blockIdx.y * blockDim.y + threadIdx.y;
  if (ox >= S || oy >= S) return;
  const int top = boxes[4 * b], left = boxes[4 * b + 1], bh = boxes[4 * b + 2], bw = boxes[4 * b + 3];
  const int sx = flip[b] ? S - 1 - ox : ox;  // the flip follows the resize: output column ox shows resized column sx
  int x0, nx, y0, ny;
  float cx, ix, tx, cy, iy, ty;
  aa_window((float)bw / (float)S, sx, bw, x0, nx, cx, ix, tx);
  aa_window((float)bh / (float)S, oy, bh, y0, ny, cy, iy, ty);
  const float inv_tx = 1.f / tx, inv_ty = 1.f / ty;
  for (int c = 0; c < 3; ++c) {
    const uint8_t* base = src + (((size_t)b * 3 + c) * H + top) * W + left;
    float acc = 0.f;
    for (int j = 0; j < ny; ++j) {
      const float wy = cubic_aa((j + y0 - cy + 0.5f) * iy) * inv_ty;
      const uint8_t* rowp = base + (size_t)(y0 + j) * W + x0;
      float racc = 0.f;
      for (int i = 0; i < nx; ++i) racc += cubic_aa((i + x0 - cx + 0.5f) * ix) * (float)rowp[i];
      acc += wy * (racc * inv_tx);
    }
    if (round_u8) acc = fminf(fmaxf(rintf(acc), 0.f), 255.f);
    out[(((size_t)b * 3 + c) * S + oy) * S + ox] = (acc * (1.f / 255.f) - mean[c]) / stdv[c];
  }
}

}  // namespace

// ================================================================================ C ABI
#define ST(s) static_cast<cudaStream_t>(s)

extern "C" int mfk_layernorm_fwd_splice(float* x, const int* rowidx, const float* gamma, const float* beta,
                                        void* y_bf16, float* y_f32, float* x_save, float* mean, float* rstd, int M,
                                        int D, float eps, const float* prompt, int T, int row0, int n_ctx,
                                        void* stream);
extern "C" int mfk_layernorm_fwd(const float* x, const int* rowidx, const float* gamma, const float* beta,
                                 void* y_bf16, float* y_f32, float* x_save, float* mean, float* rstd, int M, int D,
                                 float eps, void* stream) {
  return mfk_layernorm_fwd_splice(const_cast<float*>(x), rowidx, gamma, beta, y_bf16, y_f32, x_save, mean, rstd, M, D,
                                  eps, nullptr, 1, 0, 0, stream);
}

extern "C" int mfk_layernorm_fwd_splice(float* x, const int* rowidx, const float* gamma, const float* beta,
                                        void* y_bf16, float* y_f32, float* x_save, float* mean, float* rstd, int M,
                                        int D, float eps, const float* prompt, int T, int row0, int n_ctx,
                                        void* stream) {
  if (!x || !gamma || !beta || M <= 0) return MFK_EARG;
  if (!y_bf16 && !y_f32) return MFK_EARG;
  if (prompt && (rowidx || T <= 0 || row0 < 0 || n_ctx <= 0 || row0 + n_ctx > T)) return MFK_EARG;
  const int grid = (M + kLnFwdWarps - 1) / kLnFwdWarps;
  bf16* y16 = static_cast<bf16*>(y_bf16);
  if (D == 768) launch_pdl(ln_fwd_kernel<6>, dim3(grid), dim3(kLnFwdWarps * 32), 0, ST(stream), x, rowidx, gamma, beta, y16, y_f32, x_save, mean, rstd, M, eps, prompt, T, row0, n_ctx);
  else if (D == 512) launch_pdl(ln_fwd_kernel<4>, dim3(grid), dim3(kLnFwdWarps * 32), 0, ST(stream), x, rowidx, gamma, beta, y16, y_f32, x_save, mean, rstd, M, eps, prompt, T, row0, n_ctx);
  else if (D == 128) launch_pdl(ln_fwd_kernel<1>, dim3(grid), dim3(kLnFwdWarps * 32), 0, ST(stream), x, rowidx, gamma, beta, y16, y_f32, x_save, mean, rstd, M, eps, prompt, T, row0, n_ctx);
  else return MFK_ESHAPE;
  MFK_CHECK_LAUNCH();
  return MFK_OK;
}

extern "C" int mfk_ln_bwd_ctas(int M) {
  int g = (M + kLnWarps - 1) / kLnWarps;
  return g < 296 ? g : 296;  // 2 resident CTAs per SM on 148 SMs
}

extern "C" int mfk_layernorm_bwd_splice(const void* dy, int dy_is_bf16, const float* x, const float* mean,
                                        const float* rstd, const float* gamma, const float* g_in, float* g_out,
                                        void* g_out_bf16, float* dgamma, float* dbeta, float* partial_ws,
                                        int accumulate, int M, int D, float* gprompt, int T, int row0, int n_ctx,
                                        void* stream);
extern "C" int mfk_layernorm_bwd(const void* dy, int dy_is_bf16, const float* x, const float* mean, const float* rstd,
                                 const float* gamma, const float* g_in, float* g_out, void* g_out_bf16,
                                 float* dgamma, float* dbeta, float* partial_ws, int accumulate, int M, int D,
                                 void* stream) {
  return mfk_layernorm_bwd_splice(dy, dy_is_bf16, x, mean, rstd, gamma, g_in, g_out, g_out_bf16, dgamma, dbeta,
                                  partial_ws, accumulate, M, D, nullptr, 1, 0, 0, stream);
}

extern "C" int mfk_layernorm_bwd_splice(const void* dy, int dy_is_bf16, const float* x, const float* mean,
                                        const float* rstd, const float* gamma, const float* g_in, float* g_out,
                                        void* g_out_bf16, float* dgamma, float* dbeta, float* partial_ws,
                                        int accumulate, int M, int D, float* gprompt, int T, int row0, int n_ctx,
                                        void* stream) {
  if (!dy || !x || !mean || !rstd || !gamma || !g_out || M <= 0) return MFK_EARG;
  if (gprompt && (T <= 0 || row0 < 0 || n_ctx <= 0 || row0 + n_ctx > T)) return MFK_EARG;
  if ((dgamma || dbeta) && !partial_ws) return MFK_EARG;
  const int grid = mfk_ln_bwd_ctas(M);
  bf16* g16 = static_cast<bf16*>(g_out_bf16);
  float* part = (dgamma || dbeta) ? partial_ws : nullptr;
#define LNB(V)                                                                                                     \
  do {                                                                                                             \
    if (dy_is_bf16) launch_pdl(ln_bwd_kernel<V, true>, dim3(grid), dim3(kLnWarps * 32), 0, ST(stream), dy, x, mean, rstd, gamma, g_in, g_out, g16, part, M, gprompt, T, row0, n_ctx); \
    else launch_pdl(ln_bwd_kernel<V, false>, dim3(grid), dim3(kLnWarps * 32), 0, ST(stream), dy, x, mean, rstd, gamma, g_in, g_out, g16, part, M, gprompt, T, row0, n_ctx);           \
  } while (0)
  if (D == 768) LNB(6);
  else if (D == 512) LNB(4);
  else if (D == 128) LNB(1);
  else return MFK_ESHAPE;
#undef LNB
  MFK_CHECK_LAUNCH();
  if (part && !(accumulate & 2)) {  // bit 1: the caller reduces the partials later (mfk_partial_reduce_grouped)
    launch_pdl(partial_reduce_kernel, dim3((D + 31) / 32, 2), dim3(32, 32), 0, ST(stream), (const float*)part, grid, D,
               2LL * D, (long long)D, dgamma, dbeta, accumulate & 1);
    MFK_CHECK_LAUNCH();
  }
  return MFK_OK;
}

extern "C" int mfk_colsum(const void* x, int is_bf16, long long ld, int M, int N, float* out, float* partial_ws,
                          int accumulate, void* stream) {
  if (!x || !out || !partial_ws || M <= 0 || N <= 0 || (N & 1)) return MFK_EARG;
  const int P = 32;
  dim3 grid((N + 63) / 64, P), block(32, 8);
  if (is_bf16) launch_pdl(colsum_partial_kernel<true>, grid, block, 0, ST(stream), x, ld, M, N, partial_ws);
  else launch_pdl(colsum_partial_kernel<false>, grid, block, 0, ST(stream), x, ld, M, N, partial_ws);
  launch_pdl(partial_reduce_kernel, dim3((N + 31) / 32, 1), dim3(32, 32), 0, ST(stream), (const float*)partial_ws, P, N,
             (long long)N, 0LL, out, (float*)nullptr, accumulate);
  MFK_CHECK_LAUNCH();
  return MFK_OK;
}

extern "C" int mfk_patch_im2col(const float* img, void* out_bf16, int B, int S, void* stream) {
  if (!img || !out_bf16 || B <= 0 || S % 16) return MFK_EARG;
  const int G = S / 16;
  im2col16_kernel<<<B * G * G, 192, 0, ST(stream)>>>(img, static_cast<bf16*>(out_bf16), B, S);
  MFK_CHECK_LAUNCH();
  return MFK_OK;
}

extern "C" int mfk_vis_assemble_lnpre(const float* tok, const float* cls, const float* pos, const float* shared_ctx,
                                      const float* gamma, const float* beta, float* x0_save, float* x, float* mean,
                                      float* rstd, int B, int T, int n_ctx, int D, float eps, void* stream) {
  if (!tok || !cls || !pos || !shared_ctx || !x || D != 768) return MFK_EARG;
  const int M = B * T;
  vis_assemble_kernel<6><<<(M + kLnWarps - 1) / kLnWarps, kLnWarps * 32, 0, ST(stream)>>>(
      tok, cls, pos, shared_ctx, gamma, beta, x0_save, x, mean, rstd, B, T, n_ctx, eps);
  MFK_CHECK_LAUNCH();
  return MFK_OK;
}

extern "C" int mfk_text_assemble(const float* prefix, const float* ctx, const float* suffix, const float* pos,
                                 float* x, int C, int Te, int n_ctx, int Tfull, int D, void* stream) {
  if (!prefix || !ctx || !suffix || !pos || !x || Te > Tfull || D % 4) return MFK_EARG;
  text_assemble_kernel<<<C * Te, 128, 0, ST(stream)>>>(prefix, ctx, suffix, pos, x, C, Te, n_ctx, Tfull, D);
  MFK_CHECK_LAUNCH();
  return MFK_OK;
}

extern "C" int mfk_prompt_splice_fwd(float* x, const float* prompt, int N, int T, int row0, int n_ctx, int D,
                                     void* stream) {
  if (!x || !prompt || row0 < 0 || row0 + n_ctx > T || D % 4) return MFK_EARG;
  launch_pdl(splice_fwd_kernel, dim3(N * n_ctx), dim3(128), 0, ST(stream), x, prompt, T, row0, n_ctx, D);
  MFK_CHECK_LAUNCH();
  return MFK_OK;
}

extern "C" int mfk_prompt_splice_bwd(float* g, void* g_bf16, float* dprompt, int N, int T, int row0, int n_ctx, int D,
                                     int round_fp16, int zero_rows, void* stream) {
  if (!g || !dprompt || row0 < 0 || row0 + n_ctx > T) return MFK_EARG;
  dim3 grid((D + 31) / 32, n_ctx);
  launch_pdl(splice_bwd_kernel, grid, dim3(32, 8), 0, ST(stream), g, static_cast<bf16*>(g_bf16), dprompt, N, T, row0,
             n_ctx, D, round_fp16, zero_rows, 0LL, 0LL);
  MFK_CHECK_LAUNCH();
  return MFK_OK;
}

extern "C" int mfk_prompt_splice_bwd_batched(float* g, long long g_stride, float* dprompt, long long dp_stride,
                                             int layers, int N, int T, int row0, int n_ctx, int D, int round_fp16,
                                             void* stream) {
  if (!g || !dprompt || layers <= 0 || row0 < 0 || row0 + n_ctx > T) return MFK_EARG;
  dim3 grid((D + 31) / 32, n_ctx, layers);
  launch_pdl(splice_bwd_kernel, grid, dim3(32, 8), 0, ST(stream), g, static_cast<bf16*>(nullptr), dprompt, N, T, row0,
             n_ctx, D, round_fp16, 0, g_stride, dp_stride);
  MFK_CHECK_LAUNCH();
  return MFK_OK;
}

extern "C" int mfk_scatter_rows(const float* dx, const int* rowidx, float* g, void* g_bf16, int R, int D,
                                void* stream) {
  if (!dx || !rowidx || !g || R <= 0) return MFK_EARG;
  launch_pdl(scatter_rows_kernel, dim3(R), dim3(128), 0, ST(stream), dx, rowidx, g, static_cast<bf16*>(g_bf16), D);
  MFK_CHECK_LAUNCH();
  return MFK_OK;
}

extern "C" int mfk_scatter_rows_dense(const float* dx, const int* rowidx, float* g, int N, int T, int D, void* stream) {
  if (!dx || !rowidx || !g || N <= 0 || T <= 0 || D <= 0 || D % 4) return MFK_EARG;
  if ((reinterpret_cast<uintptr_t>(dx) | reinterpret_cast<uintptr_t>(g)) & 15) return MFK_EALIGN;
  const int rows = N * T, D4 = D / 4;
  const int threads = D4 >= 192 ? 192 : (D4 >= 128 ? 128 : 64);
  launch_pdl(scatter_rows_dense_kernel, dim3(rows < 148 * 8 ? rows : 148 * 8), dim3(threads), 0, ST(stream),
             reinterpret_cast<const float4*>(dx), rowidx, reinterpret_cast<float4*>(g), rows, T, D4);
  MFK_CHECK_LAUNCH();
  return MFK_OK;
}

extern "C" int mfk_gather_rows(const void* src, const int* rowidx, void* dst, int R, long long row_bytes, int scatter,
                               void* stream) {
  if (!src || !rowidx || !dst || R <= 0 || row_bytes <= 0 || row_bytes % 16) return MFK_EARG;
  const int chunks = (int)(row_bytes / 16);
  if (scatter)
    launch_pdl(scatter_rows_bf16_kernel, dim3(R), dim3(128), 0, ST(stream), static_cast<const uint4*>(src), rowidx,
               static_cast<uint4*>(dst), chunks);
  else
    launch_pdl(gather_rows_kernel, dim3(R), dim3(128), 0, ST(stream), static_cast<const uint4*>(src), rowidx,
               static_cast<uint4*>(dst), chunks);
  MFK_CHECK_LAUNCH();
  return MFK_OK;
}

extern "C" int mfk_transpose_bf16(const void* in, int in_is_f32, long long ldi, void* out, long long ldo, void* copy,
                                  long long ldc, int M, int N, void* stream) {
  if (!in || !out || M <= 0 || N <= 0) return MFK_EARG;
  dim3 grid((N + 31) / 32, (M + 31) / 32), block(32, 8);
  if (in_is_f32)
    transpose_to_bf16_kernel<float><<<grid, block, 0, ST(stream)>>>(static_cast<const float*>(in), ldi, static_cast<bf16*>(out), ldo, static_cast<bf16*>(copy), ldc, M, N);
  else
    transpose_to_bf16_kernel<bf16><<<grid, block, 0, ST(stream)>>>(static_cast<const bf16*>(in), ldi, static_cast<bf16*>(out), ldo, static_cast<bf16*>(copy), ldc, M, N);
  MFK_CHECK_LAUNCH();
  return MFK_OK;
}

extern "C" int mfk_cast_f32_bf16(const float* in, void* out, long long n, void* stream) {
  if (!in || !out || n <= 0) return MFK_EARG;
  const long long thr = (n + 3) / 4;
  cast_f32_bf16_kernel<<<(unsigned)((thr + 255) / 256), 256, 0, ST(stream)>>>(in, static_cast<bf16*>(out), n);
  MFK_CHECK_LAUNCH();
  return MFK_OK;
}

extern "C" int mfk_split_bf16x3(const float* x, void* out_bf16, int rows, int D, void* stream) {
  if (!x || !out_bf16 || rows <= 0 || D <= 0) return MFK_EARG;
  const long long n = (long long)rows * D;
  split_bf16x3_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ST(stream)>>>(x, static_cast<bf16*>(out_bf16), rows, D);
  MFK_CHECK_LAUNCH();
  return MFK_OK;
}

extern "C" int mfk_linear_small_fwd(const float* x, const float* W, const float* b, float* y, int m, int N, int K,
                                    void* stream) {
  if (!x || !W || !y || m <= 0) return MFK_EARG;
  linear_small_fwd_kernel<<<(N + 7) / 8, 256, 0, ST(stream)>>>(x, W, b, y, m, N, K);
  MFK_CHECK_LAUNCH();
  return MFK_OK;
}

extern "C" int mfk_linear_small_bwd(const float* x, const float* W, const float* dy, float* dW, float* db,
                                    const float* dx_add, float* dx, int m, int N, int K, void* stream) {
  if (!x || !W || !dy || m <= 0) return MFK_EARG;
  if (dW) linear_small_bwd_w_kernel<<<N, 128, 0, ST(stream)>>>(x, dy, dW, db, m, N, K);
  if (dx) linear_small_bwd_x_kernel<<<dim3((K + 31) / 32, m), dim3(32, 32), 0, ST(stream)>>>(W, dy, dx_add, dx, m, N, K);
  MFK_CHECK_LAUNCH();
  return MFK_OK;
}

extern "C" int mfk_linear_small_fwd_grouped(const void* problems_dev, int n_problems, int max_N, void* stream) {
  if (!problems_dev || n_problems <= 0 || max_N <= 0) return MFK_EARG;
  linear_small_fwd_grouped_kernel<<<dim3((max_N + 7) / 8, n_problems), 256, 0, ST(stream)>>>(
      static_cast<const SmallLinearProblem*>(problems_dev));
  MFK_CHECK_LAUNCH();
  return MFK_OK;
}

extern "C" int mfk_linear_small_bwd_grouped(const void* problems_dev, int n_problems, int max_m, int max_N, int max_K,
                                            void* stream) {
  if (!problems_dev || n_problems <= 0 || max_m <= 0 || max_N <= 0 || max_K <= 0) return MFK_EARG;
  const SmallLinearProblem* tab = static_cast<const SmallLinearProblem*>(problems_dev);
  linear_small_bwd_w_grouped_kernel<<<dim3(max_N, n_problems), 128, 0, ST(stream)>>>(tab);
  linear_small_bwd_x_grouped_kernel<<<dim3((max_K + 31) / 32, max_m, n_problems), dim3(32, 32), 0, ST(stream)>>>(tab);
  MFK_CHECK_LAUNCH();
  return MFK_OK;
}

extern "C" int mfk_repack_grouped(const void* problems_dev, int n_problems, int total_tiles, void* stream) {
  if (!problems_dev || n_problems <= 0 || total_tiles <= 0) return MFK_EARG;
  repack_grouped_kernel<<<total_tiles, dim3(32, 8), 0, ST(stream)>>>(static_cast<const RepackProblem*>(problems_dev),
                                                                      n_problems);
  MFK_CHECK_LAUNCH();
  return MFK_OK;
}

extern "C" int mfk_partial_reduce_grouped(const void* problems_dev, int n_problems, int max_N, void* stream) {
  if (!problems_dev || n_problems <= 0 || max_N <= 0) return MFK_EARG;
  launch_pdl(partial_reduce_grouped_kernel, dim3((max_N + 31) / 32, 2, n_problems), dim3(32, 32), 0, ST(stream),
             static_cast<const PartialReduceProblem*>(problems_dev));
  MFK_CHECK_LAUNCH();
  return MFK_OK;
}

extern "C" int mfk_quickgelu_split_bf16x3(const float* u, void* out_bf16, int rows, int D, void* stream) {
  if (!u || !out_bf16 || rows <= 0 || D <= 0) return MFK_EARG;
  const long long n = (long long)rows * D;
  quickgelu_split_bf16x3_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ST(stream)>>>(u, static_cast<bf16*>(out_bf16),
                                                                                   rows, D);
  MFK_CHECK_LAUNCH();
  return MFK_OK;
}

extern "C" int mfk_split_bf16x3_rhs(const float* x, void* out_bf16, int rows, int D, void* stream) {
  if (!x || !out_bf16 || rows <= 0 || D <= 0) return MFK_EARG;
  const long long n = (long long)rows * D;
  split_bf16x3_rhs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ST(stream)>>>(x, static_cast<bf16*>(out_bf16), rows, D);
  MFK_CHECK_LAUNCH();
  return MFK_OK;
}

extern "C" int mfk_dquickgelu_mul_f32(const float* dact, const float* u, float* du, long long n, void* stream) {
  if (!dact || !u || !du || n <= 0) return MFK_EARG;
  dquickgelu_mul_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ST(stream)>>>(dact, u, du, n);
  MFK_CHECK_LAUNCH();
  return MFK_OK;
}

extern "C" int mfk_patch_im2col_f32(const float* img, float* out, int B, int S, void* stream) {
  if (!img || !out || B <= 0 || S % 16) return MFK_EARG;
  const int G = S / 16;
  im2col16_f32_kernel<<<B * G * G, 192, 0, ST(stream)>>>(img, out, B, S);
  MFK_CHECK_LAUNCH();
  return MFK_OK;
}

extern "C" int mfk_rrc_flip_normalize(const void* src_u8, int B, int H, int W, const int* boxes, const void* flip_u8,
                                      const float* mean, const float* stdv, float* out, int S, int round_u8,
                                      void* stream) {
  if (!src_u8 || !boxes || !flip_u8 || !mean || !stdv || !out || B <= 0 || H <= 0 || W <= 0 || S <= 0) return MFK_EARG;
  dim3 block(32, 8), grid((S + 31) / 32, (S + 7) / 8, B);
  rrc_flip_normalize_kernel<<<grid, block, 0, ST(stream)>>>(static_cast<const uint8_t*>(src_u8), H, W, boxes,
                                                           static_cast<const uint8_t*>(flip_u8), mean, stdv, out, S,
                                                           round_u8);
  MFK_CHECK_LAUNCH();
  return MFK_OK;
}
