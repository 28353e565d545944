// FedAvg reduction, validity scan and the fused clip + SGD step (HBM-bound).
//
// mfk_fedavg_reduce restates MaPLeFederated.safe_average_weights (trainers/maple_fed.py:309-315):
//   per element: fp32 cast -> nan_to_num(nan=0, posinf=1e4, neginf=-1e4) -> sum over clients in the
//   FIXED order of torch's CPU cascade sum (sequential inside chunks of 16 clients, chunk sums added
//   sequentially, remainder last) -> true division by K -> fp32 result + fp16-rounded result (`.half()`).
// Weighted mode (north_star extension): each row is multiplied by float(n_k) first and the divisor is
// float(sum n_k). The order does not depend on the number of GPUs, so every rank that holds the gathered
// client tensors computes bit-identical results. A NaN/Inf flag word per client replaces
// check_weights_valid (trainers/maple_fed.py:317-325).
#include "mfk_common.cuh"
#include "../../include/mfk.h"

namespace {
using namespace mfk;

__device__ __forceinline__ float sanitize(float v, int* flag, bool& bad_nan, bool& bad_inf) {
  if (isnan(v)) { bad_nan = true; return 0.f; }
  if (isinf(v)) { bad_inf = true; return v > 0.f ? 1e4f : -1e4f; }
  return v;
}

// Pointer-table entries are caller data the entry point cannot inspect (device memory), so the kernels test the
// alignment of every vector access themselves and fall back to element loads / stores (rows at
// base + j*n*4 with n % 4 != 0, or arbitrary state_dict data_ptr()s).
template <typename T>
__device__ __forceinline__ bool vec_ok(const T* p, long long i) {
  return (reinterpret_cast<uintptr_t>(p + i) & (4 * sizeof(T) - 1)) == 0;
}
template <typename TIN>
__device__ __forceinline__ float4 load4(const TIN* p, long long i);
template <>
__device__ __forceinline__ float4 load4<float>(const float* p, long long i) {
  return *reinterpret_cast<const float4*>(p + i);
}
template <>
__device__ __forceinline__ float4 load4<__half>(const __half* p, long long i) {
  uint2 u = *reinterpret_cast<const uint2*>(p + i);
  float2 a = __half22float2(*reinterpret_cast<__half2*>(&u.x)), b = __half22float2(*reinterpret_cast<__half2*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}

// ---- fixed-order reduction of one float4 column group over the K client rows.
// The pointer table and the weights are staged in shared memory once per block (a pointer fetched from global memory
// in front of every row load serialised two DRAM round trips per client), and the rows are fetched EIGHT AT A TIME
// — independent 16-byte loads issued back to back — before the (order-preserving) additions: the reduction is a
// pure streaming pass and needs ~40 KB in flight per SM to reach the HBM rate (ncu, v16: 3.3 TB/s of 6.5, 41 % ALU:
// dependent loads and four branchy isnan / isinf tests per element). NaN / Inf are detected on the raw exponent
// bits of the whole group; the per-element nan_to_num path runs only when one is present.
constexpr int FED_TABLE = 64;   // clients whose pointers / weights are staged in shared memory
constexpr int FED_BATCH = 8;    // rows in flight per thread

template <typename TIN>
__device__ __forceinline__ float4 fed_load_row(const TIN* p, long long i, long long n) {
  if (i + 3 < n && vec_ok(p, i)) return load4<TIN>(p, i);
  float4 v;
  v.x = (float)p[i];
  v.y = i + 1 < n ? (float)p[i + 1] : 0.f;
  v.z = i + 2 < n ? (float)p[i + 2] : 0.f;
  v.w = i + 3 < n ? (float)p[i + 3] : 0.f;
  return v;
}
__device__ __forceinline__ bool nonfinite_bits(float v) { return (__float_as_uint(v) & 0x7F800000u) == 0x7F800000u; }

template <typename TIN>
__device__ __forceinline__ float4 fed_reduce4(const void* const* ptrs, const float* weights, int K, long long i,
                                              long long n, int* flags) {
  float4 total = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 chunk = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int k0 = 0; k0 < K; k0 += FED_BATCH) {
    float4 v[FED_BATCH];
#pragma unroll
    for (int j = 0; j < FED_BATCH; ++j)
      if (k0 + j < K) v[j] = fed_load_row<TIN>(static_cast<const TIN*>(ptrs[k0 + j]), i, n);
#pragma unroll
    for (int j = 0; j < FED_BATCH; ++j) {
      const int k = k0 + j;
      if (k < K) {
        float4 x = v[j];
        if (nonfinite_bits(x.x) | nonfinite_bits(x.y) | nonfinite_bits(x.z) | nonfinite_bits(x.w)) {
          bool bn = false, bi = false;
          x.x = sanitize(x.x, flags, bn, bi); x.y = sanitize(x.y, flags, bn, bi);
          x.z = sanitize(x.z, flags, bn, bi); x.w = sanitize(x.w, flags, bn, bi);
          if (flags) atomicOr(&flags[k], (bn ? 1 : 0) | (bi ? 2 : 0));
        }
        if (weights) {
          const float w = weights[k];
          x.x = __fmul_rn(x.x, w); x.y = __fmul_rn(x.y, w); x.z = __fmul_rn(x.z, w); x.w = __fmul_rn(x.w, w);
        }
        // explicit __fadd_rn: no FMA contraction, the order is the contract (sequential inside chunks of 16 clients,
        // chunk sums added sequentially, remainder last)
        const int pos = k & 15;
        if (pos == 0) chunk = x;
        else { chunk.x = __fadd_rn(chunk.x, x.x); chunk.y = __fadd_rn(chunk.y, x.y); chunk.z = __fadd_rn(chunk.z, x.z); chunk.w = __fadd_rn(chunk.w, x.w); }
        if (pos == 15 || k == K - 1) {
          if (k < 16) total = chunk;
          else { total.x = __fadd_rn(total.x, chunk.x); total.y = __fadd_rn(total.y, chunk.y); total.z = __fadd_rn(total.z, chunk.z); total.w = __fadd_rn(total.w, chunk.w); }
        }
      }
    }
  }
  return total;
}

// stage the first min(K, FED_TABLE) table entries in shared memory; returns the tables to read from
struct FedTables { const void* const* ptrs; const float* weights; };
__device__ __forceinline__ FedTables fed_stage_tables(const void* const* ptrs, const float* weights, int K,
                                                      const void** s_ptrs, float* s_w) {
  if (K > FED_TABLE) return {ptrs, weights};
  if ((int)threadIdx.x < K) {
    s_ptrs[threadIdx.x] = ptrs[threadIdx.x];
    if (weights) s_w[threadIdx.x] = weights[threadIdx.x];
  }
  __syncthreads();
  return {s_ptrs, weights ? s_w : nullptr};
}

template <typename TIN>
__global__ void __launch_bounds__(256)
fedavg_kernel(const void* const* __restrict__ ptrs, const float* __restrict__ weights, int K, long long n,
              float divisor, float* __restrict__ out32, __half* __restrict__ out16, int* __restrict__ flags) {
  __shared__ const void* s_ptrs[FED_TABLE];
  __shared__ float s_w[FED_TABLE];
  const FedTables tb = fed_stage_tables(ptrs, weights, K, s_ptrs, s_w);
  const long long stride = (long long)gridDim.x * blockDim.x * 4;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    const float4 total = fed_reduce4<TIN>(tb.ptrs, tb.weights, K, i, n, flags);
    float4 m;
    m.x = __fdiv_rn(total.x, divisor); m.y = __fdiv_rn(total.y, divisor);
    m.z = __fdiv_rn(total.z, divisor); m.w = __fdiv_rn(total.w, divisor);
    if (i + 3 < n && (!out32 || vec_ok(out32, i)) && (!out16 || vec_ok(out16, i))) {
      if (out32) *reinterpret_cast<float4*>(out32 + i) = m;
      if (out16) {
        __half2 h0 = __floats2half2_rn(m.x, m.y), h1 = __floats2half2_rn(m.z, m.w);
        *reinterpret_cast<uint2*>(out16 + i) = make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
      }
    } else {
      const float mm[4] = {m.x, m.y, m.z, m.w};
      for (int e = 0; e < 4 && i + e < n; ++e) {
        if (out32) out32[i + e] = mm[e];
        if (out16) out16[i + e] = __float2half_rn(mm[e]);
      }
    }
  }
}

// Sharded form of the same reduction for W ranks (one 8xB200 box): this rank reduces elements [lo, hi) only — same
// per-element order, so the values are those of fedavg_kernel bit for bit — and PUSHES the result into every rank's
// output buffers (peer pointers over NVLink). Per rank (W-1)/W * n * (4 + 4 + 2) bytes cross NVLink instead of
// (W-1) * n * 4 for the gather form. fp32 inputs; lo is a multiple of 4.
__global__ void __launch_bounds__(256)
fedavg_scatter_kernel(const void* const* __restrict__ ptrs, const float* __restrict__ weights, int K, long long n,
                      long long lo, long long hi, float divisor, float* const* __restrict__ out32,
                      __half* const* __restrict__ out16, int W) {
  __shared__ const void* s_ptrs[FED_TABLE];
  __shared__ float s_w[FED_TABLE];
  const FedTables tb = fed_stage_tables(ptrs, weights, K, s_ptrs, s_w);
  const long long stride = (long long)gridDim.x * blockDim.x * 4;
  for (long long i = lo + ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < hi; i += stride) {
    const float4 total = fed_reduce4<float>(tb.ptrs, tb.weights, K, i, n, nullptr);
    float4 m;
    m.x = __fdiv_rn(total.x, divisor); m.y = __fdiv_rn(total.y, divisor);
    m.z = __fdiv_rn(total.z, divisor); m.w = __fdiv_rn(total.w, divisor);
    const __half2 h0 = __floats2half2_rn(m.x, m.y), h1 = __floats2half2_rn(m.z, m.w);
    const float mm[4] = {m.x, m.y, m.z, m.w};
    for (int r = 0; r < W; ++r) {
      if (i + 3 < n && vec_ok(out32[r], i) && vec_ok(out16[r], i)) {
        *reinterpret_cast<float4*>(out32[r] + i) = m;
        *reinterpret_cast<uint2*>(out16[r] + i) =
            make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
      } else {
        for (int e = 0; e < 4 && i + e < n; ++e) {
          out32[r][i + e] = mm[e];
          out16[r][i + e] = __float2half_rn(mm[e]);
        }
      }
    }
  }
}

template <typename TIN>
__global__ void check_finite_kernel(const TIN* __restrict__ p, long long n, int* __restrict__ flag, bool vec) {
  bool bn = false, bi = false;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nthr = (long long)gridDim.x * blockDim.x;
  // 16-byte loads over the aligned body (the C ABI requires a 16-byte aligned base), scalar tail
  constexpr int V = 16 / (int)sizeof(TIN);
  const long long nv = vec ? n / V : 0;  // base not 16-byte aligned: all scalar
  const uint4* p4 = reinterpret_cast<const uint4*>(p);
  for (long long i = tid; i < nv; i += nthr) {
    const uint4 u = p4[i];
    const TIN* e = reinterpret_cast<const TIN*>(&u);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const float v = (float)e[j];
      bn |= isnan(v);
      bi |= isinf(v);
    }
  }
  for (long long i = nv * V + tid; i < n; i += nthr) {
    const float v = (float)p[i];
    bn |= isnan(v);
    bi |= isinf(v);
  }
  if (bn || bi) atomicOr(flag, (bn ? 1 : 0) | (bi ? 2 : 0));
}

// ---------------------------------------------------------------------------- clip_grad_norm_ + SGD
__global__ void sumsq_partial_kernel(const float* __restrict__ g, long long n, float* __restrict__ partial) {
  __shared__ float red[8];
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  const long long n4 = n >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  // four independent 16-byte loads in flight per thread (a pure streaming pass; one load per iteration ran at 48 %
  // of the HBM rate); the accumulation order per thread is fixed, so the result does not depend on timing
  for (; i + 3 * stride < n4; i += 4 * stride) {
    const float4 a = g4[i], b = g4[i + stride], c = g4[i + 2 * stride], d = g4[i + 3 * stride];
    s0 += a.x * a.x; s1 += a.y * a.y; s2 += a.z * a.z; s3 += a.w * a.w;
    s0 += b.x * b.x; s1 += b.y * b.y; s2 += b.z * b.z; s3 += b.w * b.w;
    s0 += c.x * c.x; s1 += c.y * c.y; s2 += c.z * c.z; s3 += c.w * c.w;
    s0 += d.x * d.x; s1 += d.y * d.y; s2 += d.z * d.z; s3 += d.w * d.w;
  }
  for (; i < n4; i += stride) {
    const float4 v = g4[i];
    s0 += v.x * v.x; s1 += v.y * v.y; s2 += v.z * v.z; s3 += v.w * v.w;
  }
  float s = (s0 + s1) + (s2 + s3);
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (long long k = n4 << 2; k < n; ++k) s += g[k] * g[k];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    partial[blockIdx.x] = t;
  }
}
// one warp: lane l adds partials l, l + 32, ... (loads in parallel instead of P dependent ones), then a fixed butterfly
__global__ void sumsq_final_kernel(const float* __restrict__ partial, int P, float* __restrict__ norm_out) {
  pdl_trigger();
  pdl_wait();
  float t = 0.f;
  for (int p = threadIdx.x; p < P; p += 32) t += partial[p];
  t = warp_sum(t);
  if (threadIdx.x == 0) norm_out[0] = sqrtf(t);
}
// torch.nn.utils.clip_grad_norm_(max_norm) followed by torch.optim.SGD.step (trainers/maple.py:592-598).
// hp = {lr, momentum, dampening, weight_decay, max_norm, nesterov, first_step}
// Guard: the reference raises on a NaN/Inf input image (trainers/maple.py:556-557) or total loss
// (375-376) BEFORE optim.step(), so its parameters and momentum stay intact; the fused step therefore skips the
// whole update when *loss_dev is non-finite or *flag_dev != 0. A non-finite gradient norm with a finite loss does
// reach optim.step() in the reference (error_if_nonfinite=False): torch.clamp propagates the NaN coefficient
// (fminf would return 1), reproduced here.
__global__ void sgd_step_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ mom, long long n,
                                const float* __restrict__ hp, const float* __restrict__ total_norm,
                                const float* __restrict__ loss_dev, const int* __restrict__ flag_dev) {
  if (loss_dev && !isfinite(loss_dev[0])) return;
  if (flag_dev && flag_dev[0] != 0) return;
  const float lr = hp[0], mu = hp[1], damp = hp[2], wd = hp[3], max_norm = hp[4];
  const bool nesterov = hp[5] != 0.f, first = hp[6] != 0.f;
  float coef = 1.f;
  if (max_norm > 0.f) {
    const float c = max_norm / (total_norm[0] + 1e-6f);
    coef = isnan(c) ? c : fminf(c, 1.f);
  }
  auto upd = [&](float& pv, float& gv, float& mv) {
    const float gc = gv * coef;
    gv = gc;  // grads are clipped in place, as clip_grad_norm_ does
    float d = gc + wd * pv;
    if (mu != 0.f) {
      const float b = first ? d : mu * mv + (1.f - damp) * d;
      mv = b;
      d = nesterov ? d + mu * b : b;
    }
    pv -= lr * d;
  };
  const long long n4 = n >> 2;
  float4* p4 = reinterpret_cast<float4*>(p);
  float4* g4 = reinterpret_cast<float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(mom);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 pv = p4[i], gv = g4[i], mv = (mu != 0.f && !first) ? m4[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    upd(pv.x, gv.x, mv.x); upd(pv.y, gv.y, mv.y); upd(pv.z, gv.z, mv.z); upd(pv.w, gv.w, mv.w);
    p4[i] = pv; g4[i] = gv;
    if (mu != 0.f) m4[i] = mv;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (long long i = n4 << 2; i < n; ++i) upd(p[i], g[i], mom[i]);
}

}  // namespace

#define ST(s) static_cast<cudaStream_t>(s)

extern "C" int mfk_fedavg_reduce(const void* const* client_ptrs_dev, const float* weights_dev, float divisor, int K,
                                 long long n, int in_is_fp16, float* out_f32, void* out_f16, int* flags_dev,
                                 void* stream) {
  if (!client_ptrs_dev || K <= 0 || n <= 0 || (!out_f32 && !out_f16) || !(divisor > 0.f)) return MFK_EARG;
  long long thr = (n + 3) / 4;
  long long blocks = (thr + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (in_is_fp16)
    fedavg_kernel<__half><<<(unsigned)blocks, 256, 0, ST(stream)>>>(client_ptrs_dev, weights_dev, K, n, divisor, out_f32, static_cast<__half*>(out_f16), flags_dev);
  else
    fedavg_kernel<float><<<(unsigned)blocks, 256, 0, ST(stream)>>>(client_ptrs_dev, weights_dev, K, n, divisor, out_f32, static_cast<__half*>(out_f16), flags_dev);
  MFK_CHECK_LAUNCH();
  return MFK_OK;
}

extern "C" int mfk_fedavg_reduce_scatter(const void* const* client_ptrs_dev, const float* weights_dev, float divisor,
                                         int K, long long n, long long lo, long long hi,
                                         float* const* out_f32_ptrs_dev, void* const* out_f16_ptrs_dev, int W,
                                         void* stream) {
  if (!client_ptrs_dev || !out_f32_ptrs_dev || !out_f16_ptrs_dev || K <= 0 || W <= 0 || n <= 0 || !(divisor > 0.f))
    return MFK_EARG;
  if (lo < 0 || hi > n) return MFK_ESHAPE;
  if (hi <= lo) return MFK_OK;  // empty shard (ranks beyond the end of a short tensor)
  if (lo & 3) return MFK_ESHAPE;
  long long blocks = ((hi - lo + 3) / 4 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  fedavg_scatter_kernel<<<(unsigned)blocks, 256, 0, ST(stream)>>>(client_ptrs_dev, weights_dev, K, n, lo, hi, divisor,
                                                                  out_f32_ptrs_dev,
                                                                  reinterpret_cast<__half* const*>(out_f16_ptrs_dev), W);
  MFK_CHECK_LAUNCH();
  return MFK_OK;
}

extern "C" int mfk_check_finite(const void* p, long long n, int dtype, int* flag_dev, void* stream) {
  if (!p || n <= 0 || !flag_dev) return MFK_EARG;
  long long blocks = (n + 4095) / 4096;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (dtype == 0) check_finite_kernel<float><<<(unsigned)blocks, 256, 0, ST(stream)>>>(static_cast<const float*>(p), n, flag_dev, mfk_aligned16(p));
  else if (dtype == 1) check_finite_kernel<__half><<<(unsigned)blocks, 256, 0, ST(stream)>>>(static_cast<const __half*>(p), n, flag_dev, mfk_aligned16(p));
  else if (dtype == 2) check_finite_kernel<bf16><<<(unsigned)blocks, 256, 0, ST(stream)>>>(static_cast<const bf16*>(p), n, flag_dev, mfk_aligned16(p));
  else return MFK_EARG;
  MFK_CHECK_LAUNCH();
  return MFK_OK;
}

extern "C" int mfk_grad_norm(const float* g, long long n, float* partial_ws, float* norm_out, void* stream) {
  if (!g || n <= 0 || !partial_ws || !norm_out) return MFK_EARG;
  if (!mfk_aligned16(g)) return MFK_EALIGN;
  const int P = 296;
  sumsq_partial_kernel<<<P, 256, 0, ST(stream)>>>(g, n, partial_ws);
  launch_pdl(sumsq_final_kernel, dim3(1), dim3(32), 0, ST(stream), static_cast<const float*>(partial_ws), P, norm_out);
  MFK_CHECK_LAUNCH();
  return MFK_OK;
}

extern "C" int mfk_sgd_step(float* p, float* g, float* mom, long long n, const float* hyper_dev,
                            const float* total_norm_dev, const float* loss_dev, const int* flag_dev, void* stream) {
  if (!p || !g || !mom || n <= 0 || !hyper_dev || !total_norm_dev) return MFK_EARG;
  if (!mfk_aligned16(p) || !mfk_aligned16(g) || !mfk_aligned16(mom)) return MFK_EALIGN;
  long long blocks = (n / 4 + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 8) blocks = 148 * 8;
  sgd_step_kernel<<<(unsigned)blocks, 256, 0, ST(stream)>>>(p, g, mom, n, hyper_dev, total_norm_dev, loss_dev,
                                                            flag_dev);
  MFK_CHECK_LAUNCH();
  return MFK_OK;
}

extern "C" int mfk_version(void) { return 100; }

extern "C" const char* mfk_error_string(int code) {
  switch (code) {
    case MFK_OK: return "ok";
    case MFK_EARG: return "mfk: invalid argument";
    case MFK_ESHAPE: return "mfk: unsupported shape";
    case MFK_EALIGN: return "mfk: pointer or leading dimension not 16-byte aligned";
    case MFK_EDRIVER: return "mfk: CUDA driver entry point unavailable / tensor-map encode failed";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "mfk: unknown error";
  }
}
