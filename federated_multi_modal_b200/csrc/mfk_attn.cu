// Flash-style multi-head attention forward/backward for the MaPLe towers (head dim 64).
//   vision: T = 197 + n_ctx tokens, no mask; text: T <= 77, causal (clip/model.py:303-305, 679-685;
//   nn.MultiheadAttention == softmax(Q K^T / 8 [+ causal mask]) V per head).
// Input is the fused in_proj output qkv[N*T, 3D] (q | k | v, heads are contiguous 64-wide slices,
// SURVEY.md Appendix A); output o[N*T, D] in the same token-major layout, so no permutes are needed.
//
// v1 math path: warp-level mma.sync.m16n8k16 (bf16 in, fp32 accumulate) with ldmatrix from swizzled
// shared memory; whole K/V of one (sequence, head) resident in smem (T <= 256). Softmax statistics
// in fp32, exp2 with pre-scaled log2(e). Backward is split in two deterministic kernels (dQ by query
// tile, dK/dV by key tile) — no atomics.
#include "mfk_common.cuh"
#include "../../include/mfk.h"

namespace {
using namespace mfk;

constexpr int HD = 64;          // head dim
constexpr int kWarps = 4;       // 4 warps x 16 rows = 64-row tiles
constexpr int TILE = 64;
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ uint32_t sw_off(int r, int c8) { return (uint32_t)(r * 128 + ((c8 ^ (r & 7)) << 4)); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool pred) {
  const int sz = pred ? 16 : 0;  // src-size 0 => zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Load `rows` x 64 bf16 (row stride ld elements) into swizzled smem; rows >= valid are zero-filled.
__device__ __forceinline__ void load_tile(uint32_t sbase, const bf16* g, long long ld, int rows, int valid, int tid,
                                          int nthreads) {
  for (int i = tid; i < rows * 8; i += nthreads) {
    const int r = i >> 3, c8 = i & 7;
    const bool ok = r < valid;
    cp_async16(sbase + sw_off(r, c8), g + (size_t)(ok ? r : 0) * ld + c8 * 8, ok);
  }
}

// A fragments (16 rows x 64 cols = 4 k16 steps) of this warp's rows from a swizzled tile
__device__ __forceinline__ void load_a_frags(uint32_t sbase, int row0, int lane, uint32_t (&a)[4][4]) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int r = row0 + (lane & 15), c8 = 2 * k + (lane >> 4);
    ldsm_x4(sbase + sw_off(r, c8), a[k][0], a[k][1], a[k][2], a[k][3]);
  }
}

// acc[j] (16 x 8, j = 0..7 column tiles of 8 rows of `sB`) += A(16x64) * B^T where B rows start at brow0
// (B stored [n][k], k contiguous): used for Q K^T, dO V^T, K Q^T, V dO^T.
__device__ __forceinline__ void gemm_nt_64(float (&acc)[8][4], const uint32_t (&a)[4][4], uint32_t sB, int brow0,
                                           int lane, int ntiles) {
#pragma unroll
  for (int jp = 0; jp < 4; ++jp) {
    if (2 * jp < ntiles) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        uint32_t b0, b1, b2, b3;
        const int r = brow0 + jp * 16 + (lane & 7) + ((lane >> 4) << 3), c8 = 2 * k + ((lane >> 3) & 1);
        ldsm_x4(sB + sw_off(r, c8), b0, b1, b2, b3);
        mma16816(acc[2 * jp], a[k], b0, b1);
        mma16816(acc[2 * jp + 1], a[k], b2, b3);
      }
    }
  }
}

// out[jd] (16 x 8, jd = 0..7 tiles over the 64 columns of sB) += P(16 x 64 rows-of-sB) * B where
// B stored [k][n] (n contiguous), rows brow0..brow0+63: used for P V, dS K, P^T dO, dS^T Q.
__device__ __forceinline__ void gemm_nn_64(float (&out)[8][4], const uint32_t (&p)[4][4], uint32_t sB, int brow0,
                                           int lane, int k16s) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    if (kk < k16s) {
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) {
        uint32_t b0, b1, b2, b3;
        const int r = brow0 + kk * 16 + (lane & 7) + (((lane >> 3) & 1) << 3), c8 = 2 * jp + (lane >> 4);
        ldsm_x4_t(sB + sw_off(r, c8), b0, b1, b2, b3);
        mma16816(out[2 * jp], p[kk], b0, b1);
        mma16816(out[2 * jp + 1], p[kk], b2, b3);
      }
    }
  }
}

__device__ __forceinline__ void zero_acc(float (&a)[8][4]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j][0] = a[j][1] = a[j][2] = a[j][3] = 0.f;
}
// C fragments (16 x 64 fp32) -> A fragments (bf16) for the next GEMM
__device__ __forceinline__ void acc_to_a(const float (&s)[8][4], uint32_t (&p)[4][4]) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    p[kk][0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
    p[kk][1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
    p[kk][2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
    p[kk][3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
  }
}

// store a warp's 16 x 64 fp32 fragment tile as bf16 rows (row stride ld), rows >= T skipped
__device__ __forceinline__ void store_frag_bf16(bf16* g, long long ld, int row0, int T, const float (&o)[8][4],
                                                float s0, float s1, int lane) {
  const int r0 = row0 + (lane >> 2), r1 = r0 + 8, c = (lane & 3) * 2;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (r0 < T) *reinterpret_cast<uint32_t*>(g + (size_t)r0 * ld + j * 8 + c) = pack_bf16(o[j][0] * s0, o[j][1] * s0);
    if (r1 < T) *reinterpret_cast<uint32_t*>(g + (size_t)r1 * ld + j * 8 + c) = pack_bf16(o[j][2] * s1, o[j][3] * s1);
  }
}

// ============================================================================ forward
template <bool CAUSAL>
__global__ void __launch_bounds__(kWarps * 32)
attn_fwd_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, float* __restrict__ lse, int T, int Tp,
                int heads, float scale) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int D = heads * HD;
  const long long ld = 3LL * D;
  const int qt = blockIdx.x, h = blockIdx.y, n = blockIdx.z;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t sQ = smem_u32(smem), sK = sQ + TILE * 128, sV = sK + Tp * 128;
  const bf16* base = qkv + (size_t)n * T * ld + h * HD;
  const int q0 = qt * TILE;
  const int kend = CAUSAL ? min(T, q0 + TILE) : T;  // keys needed by this query tile
  load_tile(sQ, base + (size_t)q0 * ld, ld, TILE, T - q0, tid, kWarps * 32);
  load_tile(sK, base + D, ld, Tp, kend, tid, kWarps * 32);
  load_tile(sV, base + 2 * D, ld, Tp, kend, tid, kWarps * 32);
  cp_async_wait_all();
  __syncthreads();

  uint32_t qa[4][4];
  load_a_frags(sQ, warp * 16, lane, qa);
  float o[8][4];
  zero_acc(o);
  const float c = scale * kLog2e;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  const int r0 = q0 + warp * 16 + (lane >> 2), r1 = r0 + 8;

  for (int kb = 0; kb < kend; kb += TILE) {
    const int nvalid = min(TILE, kend - kb);
    const int ntiles = (nvalid + 7) >> 3;
    float s[8][4];
    zero_acc(s);
    gemm_nt_64(s, qa, sK, kb, lane, ntiles);
    float mx0 = m0, mx1 = m1;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int key = kb + j * 8 + (lane & 3) * 2 + (e & 1);
        const int row = (e < 2) ? r0 : r1;
        const bool ok = key < kend && (!CAUSAL || key <= row);
        s[j][e] = ok ? s[j][e] * c : -INFINITY;
      }
      mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
      mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float sub0 = (mx0 == -INFINITY) ? 0.f : mx0, sub1 = (mx1 == -INFINITY) ? 0.f : mx1;
    const float corr0 = exp2f(m0 - sub0), corr1 = exp2f(m1 - sub1);
    m0 = mx0; m1 = mx1;
    float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j][0] = exp2f(s[j][0] - sub0); s[j][1] = exp2f(s[j][1] - sub0);
      s[j][2] = exp2f(s[j][2] - sub1); s[j][3] = exp2f(s[j][3] - sub1);
      rs0 += s[j][0] + s[j][1]; rs1 += s[j][2] + s[j][3];
    }
    l0 = l0 * corr0 + rs0; l1 = l1 * corr1 + rs1;
#pragma unroll
    for (int j = 0; j < 8; ++j) { o[j][0] *= corr0; o[j][1] *= corr0; o[j][2] *= corr1; o[j][3] *= corr1; }
    uint32_t p[4][4];
    acc_to_a(s, p);
    gemm_nn_64(o, p, sV, kb, lane, (nvalid + 15) >> 4);
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = l0 > 0.f ? 1.f / l0 : 0.f, i1 = l1 > 0.f ? 1.f / l1 : 0.f;
  store_frag_bf16(out + (size_t)n * T * D + h * HD, D, q0 + warp * 16, T, o, i0, i1, lane);
  if (lse && (lane & 3) == 0) {
    float* L = lse + ((size_t)n * heads + h) * T;
    if (r0 < T) L[r0] = m0 + log2f(l0);  // log2-domain log-sum-exp of the scaled scores
    if (r1 < T) L[r1] = m1 + log2f(l1);
  }
}

// ============================================================================ backward: delta = rowsum(dO * O)
__global__ void attn_delta_kernel(const bf16* __restrict__ o, const bf16* __restrict__ d_o, float* __restrict__ delta,
                                  int T, int heads, long long rows) {
  const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);  // (row, head)
  const int lane = threadIdx.x & 31;
  if (w >= rows * heads) return;
  const long long row = w / heads;
  const int h = (int)(w % heads);
  const size_t off = (size_t)row * heads * HD + h * HD + lane * 2;
  float2 a = unpack_bf16(*reinterpret_cast<const uint32_t*>(o + off));
  float2 b = unpack_bf16(*reinterpret_cast<const uint32_t*>(d_o + off));
  float s = warp_sum(a.x * b.x + a.y * b.y);
  if (lane == 0) {
    const long long n = row / T, t = row % T;
    delta[((size_t)n * heads + h) * T + t] = s;
  }
}

// ============================================================================ backward: dQ (per query tile)
template <bool CAUSAL>
__global__ void __launch_bounds__(kWarps * 32)
attn_bwd_dq_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ d_o, const float* __restrict__ lse,
                   const float* __restrict__ delta, bf16* __restrict__ dqkv, int T, int Tp, int heads, float scale) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int D = heads * HD;
  const long long ld = 3LL * D;
  const int qt = blockIdx.x, h = blockIdx.y, n = blockIdx.z;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t sQ = smem_u32(smem), sdO = sQ + TILE * 128, sK = sdO + TILE * 128, sV = sK + Tp * 128;
  const bf16* base = qkv + (size_t)n * T * ld + h * HD;
  const int q0 = qt * TILE;
  const int kend = CAUSAL ? min(T, q0 + TILE) : T;
  load_tile(sQ, base + (size_t)q0 * ld, ld, TILE, T - q0, tid, kWarps * 32);
  load_tile(sdO, d_o + ((size_t)n * T + q0) * D + h * HD, D, TILE, T - q0, tid, kWarps * 32);
  load_tile(sK, base + D, ld, Tp, kend, tid, kWarps * 32);
  load_tile(sV, base + 2 * D, ld, Tp, kend, tid, kWarps * 32);
  cp_async_wait_all();
  __syncthreads();

  uint32_t qa[4][4], da[4][4];
  load_a_frags(sQ, warp * 16, lane, qa);
  load_a_frags(sdO, warp * 16, lane, da);
  const float c = scale * kLog2e;
  const int r0 = q0 + warp * 16 + (lane >> 2), r1 = r0 + 8;
  const size_t sidx = ((size_t)n * heads + h) * T;
  const float L0 = r0 < T ? lse[sidx + r0] : 0.f, L1 = r1 < T ? lse[sidx + r1] : 0.f;
  const float D0 = r0 < T ? delta[sidx + r0] : 0.f, D1 = r1 < T ? delta[sidx + r1] : 0.f;
  float dq[8][4];
  zero_acc(dq);
  for (int kb = 0; kb < kend; kb += TILE) {
    const int nvalid = min(TILE, kend - kb);
    const int ntiles = (nvalid + 7) >> 3;
    float s[8][4], dp[8][4];
    zero_acc(s);
    zero_acc(dp);
    gemm_nt_64(s, qa, sK, kb, lane, ntiles);
    gemm_nt_64(dp, da, sV, kb, lane, ntiles);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int key = kb + j * 8 + (lane & 3) * 2 + (e & 1);
        const int row = (e < 2) ? r0 : r1;
        const bool ok = key < kend && row < T && (!CAUSAL || key <= row);
        const float p = ok ? exp2f(s[j][e] * c - ((e < 2) ? L0 : L1)) : 0.f;
        s[j][e] = p * (dp[j][e] - ((e < 2) ? D0 : D1));  // dS
      }
    }
    uint32_t ds[4][4];
    acc_to_a(s, ds);
    gemm_nn_64(dq, ds, sK, kb, lane, (nvalid + 15) >> 4);
  }
  store_frag_bf16(dqkv + (size_t)n * T * ld + h * HD, ld, q0 + warp * 16, T, dq, scale, scale, lane);
}

// ============================================================================ backward: dK, dV (per key tile)
template <bool CAUSAL>
__global__ void __launch_bounds__(kWarps * 32)
attn_bwd_dkv_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ d_o, const float* __restrict__ lse,
                    const float* __restrict__ delta, bf16* __restrict__ dqkv, int T, int Tp, int heads, float scale) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int D = heads * HD;
  const long long ld = 3LL * D;
  const int kt = blockIdx.x, h = blockIdx.y, n = blockIdx.z;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t sK = smem_u32(smem), sV = sK + TILE * 128, sQ = sV + TILE * 128, sdO = sQ + Tp * 128;
  float* sL = reinterpret_cast<float*>(smem + 2 * TILE * 128 + 2 * (size_t)Tp * 128);
  float* sDl = sL + Tp;
  const bf16* base = qkv + (size_t)n * T * ld + h * HD;
  const int k0 = kt * TILE;
  const int qbeg = CAUSAL ? (k0 / TILE) * TILE : 0;  // queries < k0 never see these keys
  load_tile(sK, base + D + (size_t)k0 * ld, ld, TILE, T - k0, tid, kWarps * 32);
  load_tile(sV, base + 2 * D + (size_t)k0 * ld, ld, TILE, T - k0, tid, kWarps * 32);
  load_tile(sQ, base, ld, Tp, T, tid, kWarps * 32);
  load_tile(sdO, d_o + (size_t)n * T * D + h * HD, D, Tp, T, tid, kWarps * 32);
  const size_t sidx = ((size_t)n * heads + h) * T;
  for (int i = tid; i < Tp; i += kWarps * 32) {
    sL[i] = i < T ? lse[sidx + i] : 0.f;
    sDl[i] = i < T ? delta[sidx + i] : 0.f;
  }
  cp_async_wait_all();
  __syncthreads();

  uint32_t ka[4][4], va[4][4];
  load_a_frags(sK, warp * 16, lane, ka);
  load_a_frags(sV, warp * 16, lane, va);
  const float c = scale * kLog2e;
  const int key0 = k0 + warp * 16 + (lane >> 2), key1 = key0 + 8;
  float dk[8][4], dv[8][4];
  zero_acc(dk);
  zero_acc(dv);
  for (int qb = qbeg; qb < T; qb += TILE) {
    const int nvalid = min(TILE, T - qb);
    const int ntiles = (nvalid + 7) >> 3;
    float st[8][4], dpt[8][4];
    zero_acc(st);
    zero_acc(dpt);
    gemm_nt_64(st, ka, sQ, qb, lane, ntiles);    // S^T[key, query]
    gemm_nt_64(dpt, va, sdO, qb, lane, ntiles);  // dP^T[key, query]
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int qi = qb + j * 8 + (lane & 3) * 2 + (e & 1);
        const int key = (e < 2) ? key0 : key1;
        const bool ok = qi < T && key < T && (!CAUSAL || key <= qi);
        const int qs = qi < Tp ? qi : 0;
        const float p = ok ? exp2f(st[j][e] * c - sL[qs]) : 0.f;
        st[j][e] = p;                               // P^T
        dpt[j][e] = p * (dpt[j][e] - sDl[qs]);      // dS^T
      }
    }
    uint32_t pa[4][4], dsa[4][4];
    acc_to_a(st, pa);
    acc_to_a(dpt, dsa);
    const int k16s = (nvalid + 15) >> 4;
    gemm_nn_64(dv, pa, sdO, qb, lane, k16s);   // dV += P^T dO
    gemm_nn_64(dk, dsa, sQ, qb, lane, k16s);   // dK += dS^T Q
  }
  bf16* gk = dqkv + (size_t)n * T * ld + D + h * HD;
  store_frag_bf16(gk, ld, k0 + warp * 16, T, dk, scale, scale, lane);
  store_frag_bf16(gk + D, ld, k0 + warp * 16, T, dv, 1.f, 1.f, lane);
}

template <typename K>
int set_smem(K kernel, size_t bytes) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  return e == cudaSuccess ? MFK_OK : (int)e;
}

}  // namespace

#define ST(s) static_cast<cudaStream_t>(s)

extern "C" int mfk_attn_fwd(const void* qkv, void* out, float* lse, int N, int T, int heads, int causal,
                            void* stream) {
  if (!qkv || !out || N <= 0 || T <= 0 || T > 256 || heads <= 0) return MFK_EARG;
  const int Tp = ((T + TILE - 1) / TILE) * TILE;
  const size_t smem = (size_t)TILE * 128 + 2 * (size_t)Tp * 128;
  dim3 grid(Tp / TILE, heads, N);
  const float scale = 0.125f;  // 1/sqrt(64)
  int rc;
  if (causal) {
    if ((rc = set_smem(attn_fwd_kernel<true>, smem)) != MFK_OK) return rc;
    attn_fwd_kernel<true><<<grid, kWarps * 32, smem, ST(stream)>>>(static_cast<const bf16*>(qkv), static_cast<bf16*>(out), lse, T, Tp, heads, scale);
  } else {
    if ((rc = set_smem(attn_fwd_kernel<false>, smem)) != MFK_OK) return rc;
    attn_fwd_kernel<false><<<grid, kWarps * 32, smem, ST(stream)>>>(static_cast<const bf16*>(qkv), static_cast<bf16*>(out), lse, T, Tp, heads, scale);
  }
  MFK_CHECK_LAUNCH();
  return MFK_OK;
}

extern "C" int mfk_attn_bwd(const void* qkv, const void* out, const void* d_out, const float* lse, float* delta_ws,
                            void* dqkv, int N, int T, int heads, int causal, void* stream) {
  if (!qkv || !out || !d_out || !lse || !delta_ws || !dqkv || N <= 0 || T <= 0 || T > 256) return MFK_EARG;
  const int Tp = ((T + TILE - 1) / TILE) * TILE;
  const long long rows = (long long)N * T;
  const long long warps = rows * heads;
  attn_delta_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, ST(stream)>>>(static_cast<const bf16*>(out), static_cast<const bf16*>(d_out), delta_ws, T, heads, rows);
  const size_t smem_dq = 2 * (size_t)TILE * 128 + 2 * (size_t)Tp * 128;
  const size_t smem_dkv = smem_dq + 2 * (size_t)Tp * sizeof(float);
  dim3 grid(Tp / TILE, heads, N);
  const float scale = 0.125f;
  const bf16* q = static_cast<const bf16*>(qkv);
  const bf16* d = static_cast<const bf16*>(d_out);
  bf16* g = static_cast<bf16*>(dqkv);
  int rc;
  if (causal) {
    if ((rc = set_smem(attn_bwd_dq_kernel<true>, smem_dq)) != MFK_OK) return rc;
    if ((rc = set_smem(attn_bwd_dkv_kernel<true>, smem_dkv)) != MFK_OK) return rc;
    attn_bwd_dq_kernel<true><<<grid, kWarps * 32, smem_dq, ST(stream)>>>(q, d, lse, delta_ws, g, T, Tp, heads, scale);
    attn_bwd_dkv_kernel<true><<<grid, kWarps * 32, smem_dkv, ST(stream)>>>(q, d, lse, delta_ws, g, T, Tp, heads, scale);
  } else {
    if ((rc = set_smem(attn_bwd_dq_kernel<false>, smem_dq)) != MFK_OK) return rc;
    if ((rc = set_smem(attn_bwd_dkv_kernel<false>, smem_dkv)) != MFK_OK) return rc;
    attn_bwd_dq_kernel<false><<<grid, kWarps * 32, smem_dq, ST(stream)>>>(q, d, lse, delta_ws, g, T, Tp, heads, scale);
    attn_bwd_dkv_kernel<false><<<grid, kWarps * 32, smem_dkv, ST(stream)>>>(q, d, lse, delta_ws, g, T, Tp, heads, scale);
  }
  MFK_CHECK_LAUNCH();
  return MFK_OK;
}
