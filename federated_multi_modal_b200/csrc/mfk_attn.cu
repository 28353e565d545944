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
  // one warp per token row: 16-byte loads, 8 lanes cover one head (64 elements), shuffle-reduce inside the octet
  pdl_trigger();
  pdl_wait();
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int chunks = heads * 8;  // 16-byte chunks per row
  const uint4* po = reinterpret_cast<const uint4*>(o + (size_t)row * heads * HD);
  const uint4* pd = reinterpret_cast<const uint4*>(d_o + (size_t)row * heads * HD);
  const long long n = row / T, t = row % T;
  for (int c = lane; c < ((chunks + 31) & ~31); c += 32) {
    float s = 0.f;
    if (c < chunks) {
      const uint4 a = po[c], b = pd[c];
      const float2 a0 = unpack_bf16(a.x), a1 = unpack_bf16(a.y), a2 = unpack_bf16(a.z), a3 = unpack_bf16(a.w);
      const float2 b0 = unpack_bf16(b.x), b1 = unpack_bf16(b.y), b2 = unpack_bf16(b.z), b3 = unpack_bf16(b.w);
      s = (a0.x * b0.x + a0.y * b0.y) + (a1.x * b1.x + a1.y * b1.y) + (a2.x * b2.x + a2.y * b2.y) +
          (a3.x * b3.x + a3.y * b3.y);
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if ((lane & 7) == 0 && c < chunks) delta[((size_t)n * heads + (c >> 3)) * T + t] = s;
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


// ============================================================================ backward, short sequences (T <= 32)
// The text tower runs on T_eff ~ 10 positions: delta + dQ + dK/dV as three launches of 64-row tensor-core tiles are
// three dependent launches of almost-empty tiles on a latency-bound side stream. One CTA per (sequence, head) does
// the whole backward in fp32 out of shared memory instead: S = Q K^T, P = exp2(c S - lse), dP = dO V^T,
// delta = rowsum(dO * O), dS = P * (dP - delta), dQ = scale dS K, dK = scale dS^T Q, dV = P^T dO.
constexpr int SMALL_T = 32, SMALL_THREADS = 128;
template <bool CAUSAL>
__global__ void __launch_bounds__(SMALL_THREADS)
attn_bwd_small_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ out, const bf16* __restrict__ d_out,
                      const float* __restrict__ lse, bf16* __restrict__ dqkv, int T, int heads, float scale) {
  extern __shared__ float sm_small[];
  const int D = heads * HD, D3 = 3 * D;
  const int h = blockIdx.x, n = blockIdx.y, tid = threadIdx.x;
  float* sQ = sm_small;                // [T][HD + 1] (padded: a thread walks a row while its neighbours walk other rows)
  float* sK = sQ + T * (HD + 1);
  float* sV = sK + T * (HD + 1);
  float* sdO = sV + T * (HD + 1);
  float* sP = sdO + T * (HD + 1);      // [T][T]
  float* sdS = sP + T * T;             // [T][T]
  float* sDelta = sdS + T * T;         // [T]
  float* sL = sDelta + T;              // [T]
  pdl_trigger();
  pdl_wait();
  const bf16* base = qkv + (size_t)n * T * D3 + h * HD;
  const bf16* ob = out + (size_t)n * T * D + h * HD;
  const bf16* db = d_out + (size_t)n * T * D + h * HD;
  for (int i = tid; i < T * (HD / 2); i += SMALL_THREADS) {
    const int r = i / (HD / 2), c2 = (i % (HD / 2)) * 2;
    const float2 q = unpack_bf16(*reinterpret_cast<const uint32_t*>(base + (size_t)r * D3 + c2));
    const float2 k = unpack_bf16(*reinterpret_cast<const uint32_t*>(base + (size_t)r * D3 + D + c2));
    const float2 v = unpack_bf16(*reinterpret_cast<const uint32_t*>(base + (size_t)r * D3 + 2 * D + c2));
    const float2 g = unpack_bf16(*reinterpret_cast<const uint32_t*>(db + (size_t)r * D + c2));
    float* d = sQ + r * (HD + 1) + c2; d[0] = q.x; d[1] = q.y;
    d = sK + r * (HD + 1) + c2; d[0] = k.x; d[1] = k.y;
    d = sV + r * (HD + 1) + c2; d[0] = v.x; d[1] = v.y;
    d = sdO + r * (HD + 1) + c2; d[0] = g.x; d[1] = g.y;
  }
  // delta_i = <dO_i, O_i>: 4 lanes per row would be enough; a warp per row keeps it simple (T <= 32 rows, 4 warps)
  for (int r = tid >> 5; r < T; r += SMALL_THREADS / 32) {
    const int lane = tid & 31;
    const float2 o = unpack_bf16(*reinterpret_cast<const uint32_t*>(ob + (size_t)r * D + 2 * lane));
    const float2 g = unpack_bf16(*reinterpret_cast<const uint32_t*>(db + (size_t)r * D + 2 * lane));
    const float s = warp_sum(o.x * g.x + o.y * g.y);
    if (lane == 0) { sDelta[r] = s; sL[r] = lse[((size_t)n * heads + h) * T + r]; }
  }
  __syncthreads();
  const float c = scale * kLog2e;
  for (int e = tid; e < T * T; e += SMALL_THREADS) {
    const int i = e / T, j = e % T;
    const float* qi = sQ + i * (HD + 1);
    const float* kj = sK + j * (HD + 1);
    const float* gi = sdO + i * (HD + 1);
    const float* vj = sV + j * (HD + 1);
    float s = 0.f, dp = 0.f;
#pragma unroll 16
    for (int d = 0; d < HD; ++d) { s = fmaf(qi[d], kj[d], s); dp = fmaf(gi[d], vj[d], dp); }
    const bool ok = !CAUSAL || j <= i;
    const float pr = ok ? exp2f(s * c - sL[i]) : 0.f;
    sP[e] = pr;
    sdS[e] = pr * (dp - sDelta[i]);
  }
  __syncthreads();
  bf16* gbase = dqkv + (size_t)n * T * D3 + h * HD;
  for (int e = tid; e < T * HD; e += SMALL_THREADS) {
    const int r = e / HD, d = e % HD;
    float dq = 0.f, dk = 0.f, dv = 0.f;
    for (int j = 0; j < T; ++j) {
      dq = fmaf(sdS[r * T + j], sK[j * (HD + 1) + d], dq);     // dQ_r += dS[r][j] K_j
      dk = fmaf(sdS[j * T + r], sQ[j * (HD + 1) + d], dk);     // dK_r += dS[j][r] Q_j
      dv = fmaf(sP[j * T + r], sdO[j * (HD + 1) + d], dv);     // dV_r += P[j][r] dO_j
    }
    gbase[(size_t)r * D3 + d] = __float2bfloat16_rn(dq * scale);
    gbase[(size_t)r * D3 + D + d] = __float2bfloat16_rn(dk * scale);
    gbase[(size_t)r * D3 + 2 * D + d] = __float2bfloat16_rn(dv);
  }
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
  if (T <= SMALL_T) {  // one launch: the whole backward of a (sequence, head) in one CTA (delta_ws is not used)
    const size_t smem = sizeof(float) * (4 * (size_t)T * (HD + 1) + 2 * (size_t)T * T + 2 * (size_t)T);
    cudaError_t e;
    if (causal)
      e = launch_pdl(attn_bwd_small_kernel<true>, dim3(heads, N), dim3(SMALL_THREADS), smem, ST(stream),
                     static_cast<const bf16*>(qkv), static_cast<const bf16*>(out), static_cast<const bf16*>(d_out), lse,
                     static_cast<bf16*>(dqkv), T, heads, 0.125f);
    else
      e = launch_pdl(attn_bwd_small_kernel<false>, dim3(heads, N), dim3(SMALL_THREADS), smem, ST(stream),
                     static_cast<const bf16*>(qkv), static_cast<const bf16*>(out), static_cast<const bf16*>(d_out), lse,
                     static_cast<bf16*>(dqkv), T, heads, 0.125f);
    if (e != cudaSuccess) return (int)e;
    MFK_CHECK_LAUNCH();
    return MFK_OK;
  }
  const int Tp = ((T + TILE - 1) / TILE) * TILE;
  const long long rows = (long long)N * T;
  const long long warps = rows * heads;
  attn_delta_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, ST(stream)>>>(static_cast<const bf16*>(out), static_cast<const bf16*>(d_out), delta_ws, T, heads, rows);
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

// =====================================================================================================
// tcgen05 / TMEM / TMA flash attention FORWARD (T <= 256: the whole key range of a sequence is one UMMA N tile).
//   S = Q K^T   : tcgen05.mma  M=128 (queries) x N=Tk (keys, multiple of 16) x K=64, A/B K-major from TMA tiles
//   P = softmax : 4 warps, one TMEM lane (= query row) per thread, two passes over S in TMEM (max, then exp2/sum),
//                 P written as bf16 into 128B-swizzled smem in the K-major A-operand layout
//   O = P V     : tcgen05.mma  M=128 x N=64 x K=Tk, B = V tile used as an MN-major operand (no transpose)
// Persistent CTAs; Q/K/V smem and the S/O TMEM accumulator are double-buffered so the TMA loads and the
// QK^T MMA of work item i+1 overlap the softmax of item i (the kernel is MUFU/issue bound, not tensor bound).
namespace {
using namespace mfk;

constexpr int TC_SOFTMAX_WARPS = 8;  // two warps per TMEM lane quarter, each takes alternate 32-column chunks
constexpr int TC_THREADS = (TC_SOFTMAX_WARPS + 1) * 32;

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// named barrier shared by the two warps that own the same TMEM lane quarter
__device__ __forceinline__ void pair_sync(int q) { asm volatile("bar.sync %0, 64;" ::"r"(q + 1) : "memory"); }

struct AttnTcParams {
  const bf16* qkv;
  bf16* out;
  float* lse;
  int T, Tk, heads, m_tiles, total_items;
  float c1;  // softmax scale * log2(e)
  long long* trace;  // optional debug: clock64 stamps of CTA 0, softmax warp 0 (tools/attn_trace.py fwd); NULL in production
  unsigned smem_bytes;  // dynamic shared memory given to the launch
};
#define FTRC()                                                                                  \
  do {                                                                                          \
    if (p.trace && blockIdx.x == 0 && warp == 0 && lane == 0 && ftrc_n < 60) p.trace[ftrc_n++] = clock64(); \
  } while (0)

// Work unit = one (sequence, head): K and V are loaded once and shared by its (<= 2) 128-query tiles, whose S/O
// accumulators use the two 256-column TMEM buffers. Q+K of the next unit are fetched as soon as this unit's QK^T
// MMAs have retired and V is double-buffered, so TMA latency is hidden behind a whole unit of softmax work.
// Two softmax warp groups (8 warps each), one per query tile of the unit: while one group is in its MUFU-bound
// exp pass the other runs its TMEM-latency-bound max pass, waits for its P V, or stores its O tile, so the phases
// of the two tiles of a (sequence, head) overlap instead of running back to back. Each group has its own MMA-issuing
// thread (S = Q K^T on tmem_free, O = P V on p_full) and a third thread refills Q/K and V as soon as both groups'
// MMAs on them have retired, so a group never waits on the other group's progress except through those buffers.
// 16 + 3 warps = 608 threads: 5 warps share one scheduler's register file -> at most 96 registers per thread.
constexpr int FWD_GROUPS = 2;
constexpr int FWD_THREADS = (FWD_GROUPS * TC_SOFTMAX_WARPS + 3) * 32;  // + MMA issuer per group + TMA loader

template <bool CAUSAL>
__global__ void __maxnreg__(96)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                   const AttnTcParams p) {
  extern __shared__ uint8_t smem_raw_tc[];
  // barriers and the max / sum exchange slots come first, the swizzled tiles after them at the next 1024-byte
  // boundary: for T = 256 the tiles alone take 224 KB and the total must stay within the 227 KB CTA limit
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw_tc);
  uint64_t* qk_full = bars;         // [1]
  uint64_t* v_full = bars + 1;      // [1]
  uint64_t* s_full = bars + 3;      // [2] per query tile / group
  uint64_t* o_full = bars + 5;      // [2]
  uint64_t* tmem_free = bars + 7;   // [2]
  uint64_t* p_full = bars + 9;      // [2][4] per group and 64-key block of P: P V starts on block 0 while the
                                    //        softmax warps are still exponentiating the later blocks
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);
  float* sx = reinterpret_cast<float*>(smem_raw_tc + 256);  // [2 groups][2 halves][128 rows] max / sum exchange
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_tc) + 256 + 2048 + 1023) &
                                             ~uintptr_t(1023));
  const int Tk = p.Tk, D = p.heads * HD, NT = p.m_tiles;
  const uint32_t kvBytes = (uint32_t)Tk * 128u, qBytes = (uint32_t)NT * 16384u;
  const int nkb = (Tk + 63) / 64;  // 64-key column blocks of P
  const uint32_t pBytes = (uint32_t)nkb * 16384u;
  uint8_t* sQ = smem;                       // [NT x 128 rows x 128 B]
  uint8_t* sK = sQ + qBytes;                // [Tk x 128 B]
  uint8_t* sV = sK + kvBytes;               // [Tk x 128 B]
  uint8_t* sP = sV + kvBytes;               // [2 groups][nkb][128 x 128 B]: P of the group's tile (also stages its O)
  if (threadIdx.x == 0 && (sP + 2 * (size_t)pBytes) - smem_raw_tc > (ptrdiff_t)p.smem_bytes) __trap();  // layout fits

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;  // provably warp-uniform
  constexpr int kCtl = FWD_GROUPS * TC_SOFTMAX_WARPS;  // control warp
  const int n_units = (p.total_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == kCtl) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmKV);
      mbar_init(qk_full, 1);
      mbar_init(v_full, 1);
      for (int b = 0; b < 2; ++b) {
        mbar_init(&s_full[b], 1);
        mbar_init(&o_full[b], 1);
        mbar_init(&tmem_free[b], TC_SOFTMAX_WARPS);
        for (int k = 0; k < 4; ++k) mbar_init(&p_full[b * 4 + k], TC_SOFTMAX_WARPS);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();

  auto unit_coords = [&](int u, int& n, int& h) {
    const int w = (int)blockIdx.x + u * (int)gridDim.x;
    h = w % p.heads;
    n = w / p.heads;
  };

  if (warp == kCtl + 2) {
    // ============================ loader: TMA loads of Q/K and V (one thread) ============================
    if (elect_one()) {
      auto issue_qk = [&](int u) {
        int n, h;
        unit_coords(u, n, h);
        mbar_arrive_expect_tx(qk_full, qBytes + kvBytes);
        tma_load_2d(&tmQ, qk_full, sQ, h * HD, n * p.T);
        tma_load_2d(&tmKV, qk_full, sK, D + h * HD, n * p.T);
      };
      auto issue_v = [&](int u) {
        int n, h;
        unit_coords(u, n, h);
        mbar_arrive_expect_tx(v_full, kvBytes);
        tma_load_2d(&tmKV, v_full, sV, 2 * D + h * HD, n * p.T);
      };
      if (n_units > 0) {
        issue_qk(0);
        issue_v(0);
      }
      for (int u = 0; u + 1 < n_units; ++u) {
        const uint32_t par = (uint32_t)u & 1u;
        for (int mt = 0; mt < NT; ++mt) mbar_wait(&s_full[mt], par);   // every S MMA on Q/K of unit u retired
        issue_qk(u + 1);
        for (int mt = 0; mt < NT; ++mt) mbar_wait(&o_full[mt], par);   // every P V MMA on V of unit u retired
        issue_v(u + 1);
      }
    }
    __syncwarp();
  } else if (warp >= kCtl) {
    // ============================ MMA issuer of query tile mt (one thread per group) ============================
    const int mt = warp - kCtl;
    if (mt < NT && elect_one()) {  // (elect_one, not a lane test: UMMA descriptors stay in uniform registers)
      const uint32_t idesc_s = umma_idesc_bf16(128, Tk, 0, 0);
      const uint32_t idesc_o = umma_idesc_bf16(128, HD, 0, 1);
      const int ksteps = Tk / 16;
      const uint32_t vbase = smem_u32(sV), pbase = smem_u32(sP) + (uint32_t)mt * pBytes;
      const uint32_t d_tmem = tmem_base + (uint32_t)mt * 256u;
      const uint64_t adesc = umma_desc_k_sw128(smem_u32(sQ) + (uint32_t)mt * 16384u);
      const uint64_t bdesc = umma_desc_k_sw128(smem_u32(sK));
      for (int u = 0; u < n_units; ++u) {
        const uint32_t par = (uint32_t)u & 1u;
        mbar_wait(qk_full, par);
        mbar_wait(&tmem_free[mt], par ^ 1u);  // the group's epilogue of the previous unit drained this buffer
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, adesc + 2ull * k, bdesc + 2ull * k, idesc_s, k > 0 ? 1u : 0u);
        umma_commit(&s_full[mt]);
        mbar_wait(v_full, par);
        // O accumulates into columns [0, 64) of the S buffer: block 0 of S has been read by every softmax warp once
        // p_full[0] completes, and no later block touches those columns, so the P V MMAs of block k are issued as soon
        // as block k of P is staged (they run under the exp pass of the later blocks instead of after it).
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&p_full[mt * 4 + kb], par);
          tc_fence_after();
          const int ks_end = min(ksteps, kb * 4 + 4);
          for (int ks = kb * 4; ks < ks_end; ++ks) {
            const uint64_t pa = umma_desc_k_sw128(pbase + (uint32_t)kb * 16384u) + 2ull * (ks & 3);
            const uint64_t vb = umma_desc_mn_sw128(vbase + (uint32_t)ks * 2048u, 1024);
            umma_bf16(d_tmem, pa, vb, idesc_o, ks > 0 ? 1u : 0u);
          }
        }
        umma_commit(&o_full[mt]);
      }
    }
    __syncwarp();
  } else if ((warp >> 3) < NT) {
    // ============================ softmax + epilogue warp groups (group mt owns query tile mt) ============================
    const int mt = warp >> 3, gw = warp & 7;          // group, warp inside the group
    const int q = gw & 3, half = gw >> 2;
    const int rl = q * 32 + lane;  // row inside the 128-row tile == TMEM lane
    const uint32_t x7 = (uint32_t)rl & 7u;
    float* gsx = sx + mt * 256;
    const int pair_bar = 1 + mt * 4 + q;
    auto pair_sync_g = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory"); };
    const uint32_t pbuf = smem_u32(sP) + (uint32_t)mt * pBytes;
    const uint32_t prow = pbuf + (uint32_t)rl * 128u;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)mt * 256u;
    const int row = mt * 128 + rl;  // query index inside the sequence
    const int kmax = CAUSAL ? min(p.T, row + 1) : p.T;  // keys [0, kmax) are visible
    // a warp whose 32 query rows all lie beyond T (T = 199: the last quarter of the second tile) only keeps the
    // barrier protocol going: no TMEM loads, no exponentials, no stores
    const bool live = mt * 128 + q * 32 < p.T;
    int ftrc_n = 0;
    for (int u = 0; u < n_units; ++u) {
      int n, h;
      unit_coords(u, n, h);
      const uint32_t par = (uint32_t)u & 1u;
      FTRC();  // about to wait for S
      mbar_wait(&s_full[mt], par);
      tc_fence_after();
      FTRC();  // S ready
      float mx = -INFINITY;
      if (live) {
        for (int c = half * 32; c < Tk; c += 64) {
          uint32_t r[32];
          tmem_ld32(taddr + (uint32_t)c, r);
          tc_wait_ld();
          float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};  // independent chains: ILP instead of a 32-deep dependency
          if (!CAUSAL && c + 32 <= p.T) {
#pragma unroll
            for (int j = 0; j < 32; ++j) m4[j & 3] = fmaxf(m4[j & 3], __uint_as_float(r[j]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              m4[j & 3] = fmaxf(m4[j & 3], (c + j < kmax) ? __uint_as_float(r[j]) : -INFINITY);
          }
          mx = fmaxf(mx, fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])));
        }
      }
      gsx[half * 128 + rl] = mx;
      pair_sync_g();
      FTRC();  // pass 1 (row max) done
      mx = fmaxf(mx, gsx[(half ^ 1) * 128 + rl]);
      const float mc = (mx == -INFINITY) ? 0.f : mx * p.c1;
      float l = 0.f;
      for (int kb = 0; kb < nkb; ++kb) {
        const int c = kb * 64 + half * 32;
        if (live && c < Tk) {
          uint32_t r[32];
          tmem_ld32(taddr + (uint32_t)c, r);
          tc_wait_ld();
          float l4[4] = {0.f, 0.f, 0.f, 0.f};
          uint32_t pk[16];
          if (!CAUSAL && c + 32 <= p.T) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float e0 = fast_exp2(__uint_as_float(r[2 * j]) * p.c1 - mc);
              const float e1 = fast_exp2(__uint_as_float(r[2 * j + 1]) * p.c1 - mc);
              l4[j & 3] += e0 + e1;
              pk[j] = pack_bf16(e0, e1);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float e0 = (c + 2 * j < kmax) ? fast_exp2(__uint_as_float(r[2 * j]) * p.c1 - mc) : 0.f;
              const float e1 = (c + 2 * j + 1 < kmax) ? fast_exp2(__uint_as_float(r[2 * j + 1]) * p.c1 - mc) : 0.f;
              l4[j & 3] += e0 + e1;
              pk[j] = pack_bf16(e0, e1);
            }
          }
          l += (l4[0] + l4[1]) + (l4[2] + l4[3]);
          const uint32_t blk = prow + (uint32_t)kb * 16384u;
          const uint32_t ch0 = (uint32_t)half * 4u;  // first 16-byte chunk of this 32-column group: 0 or 4
#pragma unroll
          for (int j = 0; j < 4; ++j)
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(blk + (((ch0 + j) ^ x7) << 4)), "r"(pk[4 * j]),
                         "r"(pk[4 * j + 1]), "r"(pk[4 * j + 2]), "r"(pk[4 * j + 3])
                         : "memory");
        }
        // block kb of P is staged by this warp (or is not its business): its P V MMAs may go
        tc_fence_before();
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[mt * 4 + kb]);
      }
      pair_sync_g();                       // both warps have read the partner's max: the slot can be reused
      gsx[half * 128 + rl] = l;
      FTRC();  // pass 2 (exp, P staged) done
      pair_sync_g();
      l += gsx[(half ^ 1) * 128 + rl];
      // ---- epilogue: O = (P V) / l -> bf16, staged by THIS warp in its own 32 rows x 64 bytes of the group's (now
      // idle) P buffer and copied out by the same warp as 64-byte row segments (8 rows per instruction): no
      // group-wide barrier, the other warps of the group are already on their way to the next unit
      mbar_wait(&o_full[mt], par);
      tc_fence_after();
      FTRC();  // O = P V ready
      const float inv = l > 0.f ? 1.f / l : 0.f;
      if (live) {
        uint32_t r[32];
        tmem_ld32(taddr + (uint32_t)(half * 32), r);
        tc_wait_ld();
#pragma unroll
        for (int j = 0; j < 4; ++j)
          asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(prow + ((((uint32_t)half * 4 + j) ^ x7) << 4)),
                       "r"(pack_bf16(__uint_as_float(r[8 * j]) * inv, __uint_as_float(r[8 * j + 1]) * inv)),
                       "r"(pack_bf16(__uint_as_float(r[8 * j + 2]) * inv, __uint_as_float(r[8 * j + 3]) * inv)),
                       "r"(pack_bf16(__uint_as_float(r[8 * j + 4]) * inv, __uint_as_float(r[8 * j + 5]) * inv)),
                       "r"(pack_bf16(__uint_as_float(r[8 * j + 6]) * inv, __uint_as_float(r[8 * j + 7]) * inv))
                       : "memory");
        if (p.lse && half == 0 && row < p.T) p.lse[((size_t)n * p.heads + h) * p.T + row] = mc + log2f(l);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_free[mt]);
      if (live) {
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const int idx = lane + m * 32;
          const int r_ = q * 32 + (idx >> 2), ch = half * 4 + (idx & 3);
          if (mt * 128 + r_ < p.T) {
            uint4 v;
            asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];"
                         : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                         : "r"(pbuf + (uint32_t)r_ * 128u + (((uint32_t)ch ^ ((uint32_t)r_ & 7u)) << 4)));
            *reinterpret_cast<uint4*>(p.out + ((size_t)n * p.T + mt * 128 + r_) * D + h * HD + ch * 8) = v;
          }
        }
        __syncwarp();  // the staged rows are overwritten by this warp's next P chunk: every lane has read its part
      }
      FTRC();  // epilogue done
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kCtl) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int g_attn_sms = 0;
}  // namespace
static long long* g_attn_trace = nullptr;

extern "C" int mfk_attn_fwd_tc(const void* qkv, void* out, float* lse, int N, int T, int heads, int causal,
                               void* stream) {
  if (!qkv || !out || N <= 0 || T <= 0 || T > 256 || heads <= 0) return MFK_EARG;
  if (g_attn_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_attn_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_attn_sms <= 0) g_attn_sms = 148;
  }
  const int D = heads * HD;
  AttnTcParams p;
  p.qkv = static_cast<const bf16*>(qkv);
  p.out = static_cast<bf16*>(out);
  p.lse = lse;
  p.T = T;
  p.Tk = (T + 15) / 16 * 16;
  p.heads = heads;
  p.m_tiles = (T + 127) / 128;
  p.total_items = N * heads;  // work units = (sequence, head)
  p.c1 = 0.125f * kLog2e;
  p.trace = g_attn_trace;
  CUtensorMap tmQ, tmKV;
  int rc = mfk_make_tmap_2d(&tmQ, qkv, 2, (uint64_t)N * T, (uint64_t)3 * D, (uint64_t)3 * D,
                            (uint32_t)(128 * p.m_tiles), 64, 128);
  if (rc != MFK_OK) return rc;
  if ((rc = mfk_make_tmap_2d(&tmKV, qkv, 2, (uint64_t)N * T, (uint64_t)3 * D, (uint64_t)3 * D, (uint32_t)p.Tk, 64,
                             128)) != MFK_OK)
    return rc;
  const int nkb = (p.Tk + 63) / 64;
  // 128 (barriers) + 2048 (exchange slots) rounded up to the tiles' 1024-byte alignment, then Q, K, V, 2 x P
  const size_t smem = 3072 + (size_t)p.m_tiles * 16384 + 2 * (size_t)p.Tk * 128 + 2 * (size_t)nkb * 16384;
  p.smem_bytes = (unsigned)smem;
  // every CTA runs the same number of units (384 units on 148 SMs = 3 rounds either way): the balanced grid (128) leaves
  // the other SMs to the concurrent text-tower stream instead of parking CTAs that finish a round early
  const int rounds = (p.total_items + g_attn_sms - 1) / g_attn_sms;
  const int grid = (p.total_items + rounds - 1) / rounds;
  cudaError_t e;
  if (causal) {
    e = cudaFuncSetAttribute(attn_fwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    e = launch_pdl(attn_fwd_tc_kernel<true>, dim3(grid), dim3(FWD_THREADS), smem, static_cast<cudaStream_t>(stream), tmQ, tmKV, p);
    if (e != cudaSuccess) return (int)e;
  } else {
    e = cudaFuncSetAttribute(attn_fwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    e = launch_pdl(attn_fwd_tc_kernel<false>, dim3(grid), dim3(FWD_THREADS), smem, static_cast<cudaStream_t>(stream), tmQ, tmKV, p);
    if (e != cudaSuccess) return (int)e;
  }
  MFK_CHECK_LAUNCH();
  return MFK_OK;
}

// =====================================================================================================
// tcgen05 / TMEM / TMA flash attention BACKWARD (T <= 256), two deterministic kernels (no atomics):
//   dq kernel,  item = (sequence, head, 128-query tile):   S = Q_i K^T, dP = dO_i V^T  ->  dS  ->  dQ_i = dS K
//   dkv kernel, item = (sequence, head, 128-key tile):      S^T = K_j Q^T, dP^T = V_j dO^T -> P^T, dS^T
//                                                           dV_j = P^T dO,  dK_j = dS^T Q
// One TMEM lane (= query row resp. key row) per thread for the elementwise part; P / dS are staged as bf16 in
// 128B-swizzled smem in the K-major A-operand layout; K / Q / dO are re-used in place as MN-major B operands
// (no transposed copies). The accumulators of the second GEMMs alias the TMEM columns of S / dP.
namespace {
using namespace mfk;

struct AttnBwdTcParams {
  const float* lse;
  const float* delta;
  bf16* dqkv;
  int T, Tk, heads, m_tiles, total_items;
  float c1, scale;
};

// MODE 0: dq kernel, MODE 1: dkv kernel
template <int MODE, bool CAUSAL>
__global__ void __launch_bounds__(TC_THREADS, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmTile,   // 128-row boxes of qkv (Q_i | K_j, V_j)
                   const __grid_constant__ CUtensorMap tmFull,   // Tk-row boxes of qkv (K, V | Q)
                   const __grid_constant__ CUtensorMap tmDoTile, // 128-row boxes of dO
                   const __grid_constant__ CUtensorMap tmDoFull, // Tk-row boxes of dO
                   const AttnBwdTcParams p) {
  extern __shared__ uint8_t smem_raw_tcb[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_tcb) + 1023) & ~uintptr_t(1023));
  const int Tk = p.Tk, D = p.heads * HD;
  const uint32_t fullBytes = (uint32_t)Tk * 128u;
  const int nkb = (Tk + 63) / 64;
  // MODE 0: tA = Q_i, tB = dO_i, fA = K, fB = V, st0 = dS
  // MODE 1: tA = K_j, tB = V_j,  fA = Q, fB = dO, st0 = P^T, st1 = dS^T
  uint8_t* tA = smem;
  uint8_t* tB = tA + 16384;
  uint8_t* fA = tB + 16384;
  uint8_t* fB = fA + fullBytes;
  uint8_t* st0 = fB + fullBytes;
  uint8_t* st1 = st0 + (size_t)nkb * 16384;
  uint8_t* after = MODE == 1 ? st1 + (size_t)nkb * 16384 : st1;
  float* sL = reinterpret_cast<float*>(after);  // [256] lse   (MODE 1)
  float* sD = sL + 256;                         // [256] delta (MODE 1)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sD + 256);
  uint64_t* ld_full = bars;        // TMA loads landed
  uint64_t* sd_full = bars + 1;    // S and dP accumulators complete
  uint64_t* ds_full = bars + 2;    // staged bf16 operands written (count 4)
  uint64_t* out_full = bars + 3;   // second GEMM(s) complete
  uint64_t* item_free = bars + 4;  // epilogue finished with TMEM + smem (count 4)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_items = (p.total_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == TC_SOFTMAX_WARPS) {
    if (lane == 0) {
      tma_prefetch_desc(&tmTile);
      tma_prefetch_desc(&tmFull);
      tma_prefetch_desc(&tmDoTile);
      tma_prefetch_desc(&tmDoFull);
      mbar_init(ld_full, 1);
      mbar_init(sd_full, 1);
      mbar_init(ds_full, TC_SOFTMAX_WARPS);
      mbar_init(out_full, 1);
      mbar_init(item_free, TC_SOFTMAX_WARPS);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t colS = 0, colDP = 256;  // accumulators of the second GEMMs alias these

  auto item_coords = [&](int i, int& n, int& h, int& mt) {
    const int w = (int)blockIdx.x + i * (int)gridDim.x;
    mt = w % p.m_tiles;
    h = (w / p.m_tiles) % p.heads;
    n = w / (p.m_tiles * p.heads);
  };

  if (warp == TC_SOFTMAX_WARPS) {
    if (lane == 0) {
      for (int i = 0; i < n_items; ++i) {
        int n, h, mt;
        item_coords(i, n, h, mt);
        const uint32_t par = (uint32_t)i & 1u;
        mbar_wait(item_free, par ^ 1u);  // previous item fully drained (passes for i = 0)
        mbar_arrive_expect_tx(ld_full, 2u * 16384u + 2u * fullBytes);
        const int row_t = n * p.T + mt * 128, row_f = n * p.T;
        if (MODE == 0) {
          tma_load_2d(&tmTile, ld_full, tA, h * HD, row_t);             // Q_i
          tma_load_2d(&tmDoTile, ld_full, tB, h * HD, row_t);           // dO_i
          tma_load_2d(&tmFull, ld_full, fA, D + h * HD, row_f);         // K
          tma_load_2d(&tmFull, ld_full, fB, 2 * D + h * HD, row_f);     // V
        } else {
          tma_load_2d(&tmTile, ld_full, tA, D + h * HD, row_t);         // K_j
          tma_load_2d(&tmTile, ld_full, tB, 2 * D + h * HD, row_t);     // V_j
          tma_load_2d(&tmFull, ld_full, fA, h * HD, row_f);             // Q
          tma_load_2d(&tmDoFull, ld_full, fB, h * HD, row_f);           // dO
        }
        mbar_wait(ld_full, par);
        tc_fence_after();
        {  // S (or S^T) and dP (or dP^T): M=128, N=Tk, K=64, all operands K-major
          const uint32_t idesc = umma_idesc_bf16(128, Tk, 0, 0);
          const uint64_t a0 = umma_desc_k_sw128(smem_u32(tA)), b0 = umma_desc_k_sw128(smem_u32(fA));
          const uint64_t a1 = umma_desc_k_sw128(smem_u32(tB)), b1 = umma_desc_k_sw128(smem_u32(fB));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + colS, a0 + 2ull * k, b0 + 2ull * k, idesc, k > 0);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + colDP, a1 + 2ull * k, b1 + 2ull * k, idesc, k > 0);
          umma_commit(sd_full);
        }
        mbar_wait(ds_full, par);
        tc_fence_after();
        {  // second GEMM(s): M=128, N=64, K=Tk; A = staged bf16 (K-major), B = full tile as MN-major
          const uint32_t idesc = umma_idesc_bf16(128, HD, 0, 1);
          const int ksteps = Tk / 16;
          const uint32_t s0 = smem_u32(st0), s1 = smem_u32(st1);
          const uint32_t bq = smem_u32(fA), bdo = smem_u32(fB);
          for (int ks = 0; ks < ksteps; ++ks) {
            const uint64_t a = umma_desc_k_sw128(s0 + (uint32_t)(ks >> 2) * 16384u) + 2ull * (ks & 3);
            // MODE 0: dQ = dS K (B = K = fA).  MODE 1: dV = P^T dO (B = dO = fB)
            const uint64_t b = umma_desc_mn_sw128((MODE == 0 ? bq : bdo) + (uint32_t)ks * 2048u, 1024);
            umma_bf16(tmem_base + colS, a, b, idesc, ks > 0);
          }
          if (MODE == 1) {
            for (int ks = 0; ks < ksteps; ++ks) {  // dK = dS^T Q (B = Q = fA)
              const uint64_t a = umma_desc_k_sw128(s1 + (uint32_t)(ks >> 2) * 16384u) + 2ull * (ks & 3);
              const uint64_t b = umma_desc_mn_sw128(bq + (uint32_t)ks * 2048u, 1024);
              umma_bf16(tmem_base + colDP, a, b, idesc, ks > 0);
            }
          }
          umma_commit(out_full);
        }
      }
    }
    __syncwarp();
  } else {
    const int q = warp & 3, half = warp >> 2;
    const int rl = q * 32 + lane;
    const uint32_t x7 = (uint32_t)rl & 7u;
    const uint32_t row0_s = smem_u32(st0) + (uint32_t)rl * 128u, row1_s = smem_u32(st1) + (uint32_t)rl * 128u;
    for (int i = 0; i < n_items; ++i) {
      int n, h, mt;
      item_coords(i, n, h, mt);
      const uint32_t par = (uint32_t)i & 1u;
      const int row = mt * 128 + rl;  // MODE 0: query index; MODE 1: key index
      const size_t sidx = ((size_t)n * p.heads + h) * p.T;
      const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
      float Lr = 0.f, Dr = 0.f;
      if (MODE == 0) {
        if (row < p.T) { Lr = p.lse[sidx + row]; Dr = p.delta[sidx + row]; }
      } else {
        // the previous item's reads of sL/sD are ordered before this point by its ds_full arrive + the
        // epilogue; a named barrier makes the new values visible to all 128 threads
        for (int t = threadIdx.x; t < 256; t += TC_SOFTMAX_WARPS * 32) {
          sL[t] = t < p.T ? p.lse[sidx + t] : 0.f;
          sD[t] = t < p.T ? p.delta[sidx + t] : 0.f;
        }
        asm volatile("bar.sync 5, %0;" ::"n"(TC_SOFTMAX_WARPS * 32) : "memory");
      }
      mbar_wait(sd_full, par);
      tc_fence_after();
      for (int c = half * 32; c < Tk; c += 64) {
        uint32_t rs[32], rd[32];
        tmem_ld32(tlane + colS + (uint32_t)c, rs);
        tmem_ld32(tlane + colDP + (uint32_t)c, rd);
        tc_wait_ld();
        float pv[32], dv[32];
        const bool full = !CAUSAL && c + 32 <= p.T;  // whole chunk visible (rows >= T produce unused values)
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int col = c + j;  // MODE 0: key index; MODE 1: query index
          bool ok = true;
          float L, Dl;
          if (MODE == 0) {
            if (!full) ok = col < p.T && (!CAUSAL || col <= row);
            L = Lr; Dl = Dr;
          } else {
            if (!full) ok = col < p.T && (!CAUSAL || row <= col);
            L = sL[col & 255]; Dl = sD[col & 255];
          }
          const float e = fast_exp2(__uint_as_float(rs[j]) * p.c1 - L);
          const float pr = (ok && row < p.T) ? e : 0.f;
          pv[j] = pr;
          dv[j] = pr * (__uint_as_float(rd[j]) - Dl);
        }
        const uint32_t boff = (uint32_t)(c >> 6) * 16384u;
        const uint32_t ch0 = (uint32_t)(c & 63) >> 3;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t off = boff + (((ch0 + j) ^ x7) << 4);
          if (MODE == 0) {
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(row0_s + off),
                         "r"(pack_bf16(dv[8 * j], dv[8 * j + 1])), "r"(pack_bf16(dv[8 * j + 2], dv[8 * j + 3])),
                         "r"(pack_bf16(dv[8 * j + 4], dv[8 * j + 5])), "r"(pack_bf16(dv[8 * j + 6], dv[8 * j + 7]))
                         : "memory");
          } else {
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(row0_s + off),
                         "r"(pack_bf16(pv[8 * j], pv[8 * j + 1])), "r"(pack_bf16(pv[8 * j + 2], pv[8 * j + 3])),
                         "r"(pack_bf16(pv[8 * j + 4], pv[8 * j + 5])), "r"(pack_bf16(pv[8 * j + 6], pv[8 * j + 7]))
                         : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(row1_s + off),
                         "r"(pack_bf16(dv[8 * j], dv[8 * j + 1])), "r"(pack_bf16(dv[8 * j + 2], dv[8 * j + 3])),
                         "r"(pack_bf16(dv[8 * j + 4], dv[8 * j + 5])), "r"(pack_bf16(dv[8 * j + 6], dv[8 * j + 7]))
                         : "memory");
          }
        }
      }
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(ds_full);
      // ---- epilogue
      mbar_wait(out_full, par);
      tc_fence_after();
      const size_t grow = ((size_t)n * p.T + row) * (3 * (size_t)D) + h * HD;
#pragma unroll
      for (int part = 0; part < (MODE == 0 ? 1 : 2); ++part) {
        // MODE 0: dQ (scaled) -> q part.  MODE 1: part 0 = dV -> v part, part 1 = dK (scaled) -> k part
        const uint32_t colbase = part == 0 ? colS : colDP;
        const float sc = (MODE == 0 || part == 1) ? p.scale : 1.f;
        bf16* dst = p.dqkv + grow + (MODE == 0 ? 0 : (part == 0 ? 2 * D : D));
        {
          const int c = half * 32;  // each warp of the pair writes 32 of the 64 head-dim columns
          uint32_t r[32];
          tmem_ld32(tlane + colbase + (uint32_t)c, r);
          tc_wait_ld();
          if (row < p.T) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 v;
              v.x = pack_bf16(__uint_as_float(r[8 * j]) * sc, __uint_as_float(r[8 * j + 1]) * sc);
              v.y = pack_bf16(__uint_as_float(r[8 * j + 2]) * sc, __uint_as_float(r[8 * j + 3]) * sc);
              v.z = pack_bf16(__uint_as_float(r[8 * j + 4]) * sc, __uint_as_float(r[8 * j + 5]) * sc);
              v.w = pack_bf16(__uint_as_float(r[8 * j + 6]) * sc, __uint_as_float(r[8 * j + 7]) * sc);
              *reinterpret_cast<uint4*>(dst + c + 8 * j) = v;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(item_free);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == TC_SOFTMAX_WARPS) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int MODE, bool CAUSAL>
int launch_bwd_tc(const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& c, const CUtensorMap& d,
                  const AttnBwdTcParams& p, cudaStream_t st) {
  const int nkb = (p.Tk + 63) / 64;
  const size_t smem = 2 * 16384 + 2 * (size_t)p.Tk * 128 + (size_t)(MODE == 1 ? 2 : 1) * nkb * 16384 + 2048 + 128 + 1024;
  cudaError_t e = cudaFuncSetAttribute(attn_bwd_tc_kernel<MODE, CAUSAL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem);
  if (e != cudaSuccess) return (int)e;
  const int grid = p.total_items < g_attn_sms ? p.total_items : g_attn_sms;
  attn_bwd_tc_kernel<MODE, CAUSAL><<<grid, TC_THREADS, smem, st>>>(a, b, c, d, p);
  return MFK_OK;
}
}  // namespace

extern "C" int mfk_attn_bwd_tc(const void* qkv, const void* out, const void* d_out, const float* lse, float* delta_ws,
                               void* dqkv, int N, int T, int heads, int causal, void* stream) {
  if (!qkv || !out || !d_out || !lse || !delta_ws || !dqkv || N <= 0 || T <= 0) return MFK_EARG;
  if (T > 240) return MFK_ESHAPE;  // staged P^T + dS^T + Q + dO exceed 227 KB of smem beyond 240 keys
  if (g_attn_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_attn_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_attn_sms <= 0) g_attn_sms = 148;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int D = heads * HD;
  const long long rows = (long long)N * T;
  const long long warps = rows * heads;
  attn_delta_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(static_cast<const bf16*>(out),
                                                                  static_cast<const bf16*>(d_out), delta_ws, T, heads, rows);
  AttnBwdTcParams p;
  p.lse = lse; p.delta = delta_ws; p.dqkv = static_cast<bf16*>(dqkv);
  p.T = T; p.Tk = (T + 15) / 16 * 16; p.heads = heads;
  p.m_tiles = (T + 127) / 128;
  p.total_items = N * heads * p.m_tiles;
  p.c1 = 0.125f * kLog2e; p.scale = 0.125f;
  CUtensorMap tmTile, tmFull, tmDoTile, tmDoFull;
  int rc;
  if ((rc = mfk_make_tmap_2d(&tmTile, qkv, 2, (uint64_t)rows, 3ull * D, 3ull * D, 128, 64, 128)) != MFK_OK) return rc;
  if ((rc = mfk_make_tmap_2d(&tmFull, qkv, 2, (uint64_t)rows, 3ull * D, 3ull * D, (uint32_t)p.Tk, 64, 128)) != MFK_OK) return rc;
  if ((rc = mfk_make_tmap_2d(&tmDoTile, d_out, 2, (uint64_t)rows, (uint64_t)D, (uint64_t)D, 128, 64, 128)) != MFK_OK) return rc;
  if ((rc = mfk_make_tmap_2d(&tmDoFull, d_out, 2, (uint64_t)rows, (uint64_t)D, (uint64_t)D, (uint32_t)p.Tk, 64, 128)) != MFK_OK) return rc;
  if (causal) {
    if ((rc = launch_bwd_tc<0, true>(tmTile, tmFull, tmDoTile, tmDoFull, p, st)) != MFK_OK) return rc;
    if ((rc = launch_bwd_tc<1, true>(tmTile, tmFull, tmDoTile, tmDoFull, p, st)) != MFK_OK) return rc;
  } else {
    if ((rc = launch_bwd_tc<0, false>(tmTile, tmFull, tmDoTile, tmDoFull, p, st)) != MFK_OK) return rc;
    if ((rc = launch_bwd_tc<1, false>(tmTile, tmFull, tmDoTile, tmDoFull, p, st)) != MFK_OK) return rc;
  }
  MFK_CHECK_LAUNCH();
  return MFK_OK;
}

// =====================================================================================================
// Fused tcgen05 attention BACKWARD: one work unit = one (sequence, head); Q, K, V, dO are loaded ONCE into smem
// and every 128x128 (query tile i, key tile j) block computes S, dP once:
//     S_ij = Q_i K_j^T, dP_ij = dO_i V_j^T   (K-major operands)        -> TMEM cols [0,128) / [128,256)
//     P, dS (bf16) staged in swizzled smem by 8 warps (one TMEM lane = one query row per thread)
//     dV_j += P^T dO_i, dK_j += dS^T Q_i      (staged tile as MN-major A, dO_i / Q_i as MN-major B)
//     dQ_i += dS K_j                           (staged tile as K-major A,  K_j as MN-major B)
// Accumulators: dQ_0, dQ_1, dK_j, dV_j live in TMEM cols [256,512). tcgen05.commit ordering is used so that the
// S/dP MMAs of block k+1 are queued right behind the second-stage MMAs of block k. Deterministic (no atomics).
namespace {
using namespace mfk;

// 8 compute warps + two single-thread MMA issuers: tcgen05.mma issue costs ~70 cycles per instruction from one
// thread, and a block needs 32 of them, so S/dP + dQ and dV + dK are issued from two warps in parallel.
// + 4 drain warps (one per TMEM lane quarter): dK_j / dV_j / dQ_i leave TMEM -> bf16 -> their own staging rows ->
// coalesced global stores while the compute warps and the tensor core are already on the next key tile / the next
// unit (the drains were 7.8 k of a unit's 23 k cycles when the compute warps did them between blocks).
constexpr int FUSED_DRAIN_WARPS = 4;
constexpr int FUSED_THREADS = (TC_SOFTMAX_WARPS + 2 + FUSED_DRAIN_WARPS) * 32;

struct AttnBwdFusedParams {
  const float* lse;
  const bf16* out;    // forward output O  [rows, D]
  const bf16* d_out;  // dO                 [rows, D]   (delta = rowsum(dO * O) is computed in-kernel)
  bf16* dqkv;
  int T, R, tiles, heads, total_units;
  float c1, scale;
  long long* trace;  // optional debug: clock64 stamps of CTA 0 (see tools/attn_trace.py); NULL in production
};

#define TRC(slot)                                                             \
  do {                                                                        \
    if (p.trace && blockIdx.x == 0 && trc_n < 60) p.trace[(slot) * 64 + trc_n++] = clock64(); \
  } while (0)

__global__ void __launch_bounds__(FUSED_THREADS, 1)
attn_bwd_fused_tc_kernel(const __grid_constant__ CUtensorMap tmQkv, const __grid_constant__ CUtensorMap tmDo,
                         const __grid_constant__ CUtensorMap tmQkv2, const __grid_constant__ CUtensorMap tmDo2,
                         const __grid_constant__ CUtensorMap tmOut, const AttnBwdFusedParams p) {
  extern __shared__ uint8_t smem_raw_f[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_f) + 1023) & ~uintptr_t(1023));
  const int D = p.heads * HD, R = p.R, NT = p.tiles;
  const uint32_t fullBytes = (uint32_t)R * 128u;
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + fullBytes;
  uint8_t* sV = sK + fullBytes;
  uint8_t* sdO = sV + fullBytes;
  uint8_t* sP = sdO + fullBytes;   // [2 column blocks][128 rows x 128 B]
  uint8_t* sdS = sP + 32768;
  uint8_t* sEpi = sdS + 32768;     // [4 drain warps][32 rows x 128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sEpi + 16384);
  uint64_t* ld_full = bars;        // first 128 rows of Q, K, V, dO (two arrivals: K + V, Q + dO)
  uint64_t* ld2_full = bars + 9;   // rows [128, T) of the four operands (T > 128)
  uint64_t* sd_full = bars + 1;
  uint64_t* ds_full = bars + 2;
  uint64_t* mma2_done = bars + 3;
  uint64_t* dkv_free = bars + 4;
  uint64_t* unit_free = bars + 5;
  uint64_t* acc_done = bars + 6;   // every MMA into dK_j / dV_j (and, on the last key tile, dQ) has retired
  uint64_t* sdp_issued = bars + 7;  // issuer A has queued S / dP of the next block: issuer B's dV / dK may follow
  uint64_t* stats_full = bars + 10;  // [2] lse / delta of a unit's query rows are in smem (one arrival per drain warp);
                                     //     barrier u & 1 for unit u, so that the statistics may run a whole unit ahead
  uint64_t* sdp_consumed = bars + 12;  // every compute warp has read S / dP of the block out of TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  float* s_lse = reinterpret_cast<float*>(bars + 14);   // [2][256] log-sum-exp of the unit's query rows (unit parity)
  float* s_delta = s_lse + 512;                         // [2][256] delta = rowsum(dO * O)

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;  // provably warp-uniform
  const int n_units = (p.total_units - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int nblk = NT * NT;

  if (warp == TC_SOFTMAX_WARPS) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQkv);
      tma_prefetch_desc(&tmDo);
      tma_prefetch_desc(&tmOut);
      mbar_init(ld_full, 2);
      mbar_init(ld2_full, 1);
      mbar_init(sd_full, 1);
      mbar_init(ds_full, TC_SOFTMAX_WARPS);
      mbar_init(mma2_done, 2);  // one commit per issuing thread
      mbar_init(acc_done, 2);
      mbar_init(sdp_issued, 1);
      mbar_init(&stats_full[0], FUSED_DRAIN_WARPS);
      mbar_init(&stats_full[1], FUSED_DRAIN_WARPS);
      mbar_init(sdp_consumed, TC_SOFTMAX_WARPS);
      mbar_init(dkv_free, FUSED_DRAIN_WARPS);
      mbar_init(unit_free, FUSED_DRAIN_WARPS);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  // The second tile of each operand is loaded as a box of R2 = 16 * ceil((T - 128) / 16) rows only (80 of 128 at
  // T = 199): rows [R2, 128) of it are never written by TMA and are zeroed once here — they feed MMAs as M / N rows whose
  // results are masked, so they must be finite.
  const int R2 = NT == 2 ? (p.T - 128 + 15) / 16 * 16 : 0;
  if (NT == 2 && R2 < 128) {
    const int per = (128 - R2) * 8;  // 16-byte chunks per buffer
    for (int idx = (int)threadIdx.x; idx < 4 * per; idx += (int)blockDim.x) {
      uint8_t* base = (idx / per == 0 ? sQ : idx / per == 1 ? sK : idx / per == 2 ? sV : sdO) + 16384 + R2 * 128;
      reinterpret_cast<uint4*>(base)[idx % per] = make_uint4(0u, 0u, 0u, 0u);
    }
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t colS = 0, colDP = 128, colDQ = 256, colDK = 384, colDV = 448;
  pdl_trigger();
  pdl_wait();

  if (warp >= TC_SOFTMAX_WARPS + 2) {
    // ============================ drain warps: accumulators -> global ============================
    // A thread owns one TMEM lane (= row) of its warp's lane quarter; 32 lanes storing 128 bytes of 32 different
    // rows each would cost 32 memory wavefronts per instruction, so the bf16 rows are staged in the warp's own 4 KB
    // (16-byte chunks XOR-swizzled by row = the 128-byte TMA swizzle) and leave as one TMA store. Warp-local: no barrier
    // with the other drain warps.
    const int q = warp & 3;
    const uint32_t x7 = (uint32_t)lane & 7u;
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t my = smem_u32(sEpi) + (uint32_t)q * 4096u;
    uint32_t kt_ctr = 0;
    int trc_n = 0;
    (void)trc_n;
    // lse and delta = rowsum(dO * O) of the unit's query rows -> smem. Coalesced: a call covers 32 of the (<= 256)
    // rows; 8 lanes read the 128 bytes of one (row, head) of O and of dO, 4 rows per instruction, and the row dot
    // product is finished with 3 shuffles (one lane per row reading its own 128 bytes would cost 32 memory
    // wavefronts per instruction instead of 4).
    auto row_stats_32 = [&](int u, int r0) {
      const int w = (int)blockIdx.x + u * (int)gridDim.x;
      const int h = w % p.heads, n = w / p.heads;
      const size_t sidx = ((size_t)n * p.heads + h) * p.T;
      float* s_lse_u = s_lse + (u & 1) * 256;
      float* s_delta_u = s_delta + (u & 1) * 256;
      {
        const int qrow = r0 + lane;
        s_lse_u[qrow] = qrow < p.T ? p.lse[sidx + qrow] : 0.f;
      }
      // all 16 loads are issued before the first use (one memory round trip instead of eight)
      uint4 va[8], vb[8];
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const int qrow = min(r0 + g * 4 + (lane >> 3), p.T - 1);  // clamped: rows >= T are discarded below
        const size_t off = ((size_t)n * p.T + qrow) * D + h * HD + (lane & 7) * 8;
        va[g] = __ldg(reinterpret_cast<const uint4*>(p.out + off));
        vb[g] = __ldg(reinterpret_cast<const uint4*>(p.d_out + off));
      }
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const int qrow = r0 + g * 4 + (lane >> 3);
        const uint4 a = va[g], b = vb[g];
        const float2 a0 = unpack_bf16(a.x), a1 = unpack_bf16(a.y), a2 = unpack_bf16(a.z), a3 = unpack_bf16(a.w);
        const float2 b0 = unpack_bf16(b.x), b1 = unpack_bf16(b.y), b2 = unpack_bf16(b.z), b3 = unpack_bf16(b.w);
        float dsum = (a0.x * b0.x + a0.y * b0.y) + (a1.x * b1.x + a1.y * b1.y) + (a2.x * b2.x + a2.y * b2.y) +
                     (a3.x * b3.x + a3.y * b3.y);
        dsum += __shfl_xor_sync(0xffffffffu, dsum, 1);
        dsum += __shfl_xor_sync(0xffffffffu, dsum, 2);
        dsum += __shfl_xor_sync(0xffffffffu, dsum, 4);
        if ((lane & 7) == 0) s_delta_u[qrow] = qrow < p.T ? dsum : 0.f;
      }
    };
    // Statistics of unit u -> buffer and barrier u & 1, 64 rows per drain warp. Unit 0 up front, unit u + 1 during this
    // warp's pass over unit u: the drains of unit u - 1 are behind it, so the compute warps are past their wait on
    // barrier (u + 1) & 1 for unit u - 1 (a barrier must never run two phases ahead of a waiter) and done with that
    // buffer.
    // (two passes of 32 rows; `last` publishes. Each pass is a global round trip: they are placed in the two gaps in
    // which this warp waits for a key tile anyway.)
    auto row_stats = [&](int u, int pass, bool last) {
      if (u >= n_units) return;
      const int r0 = q * 64 + pass * 32;
      if (r0 < R) row_stats_32(u, r0);
      if (last) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&stats_full[u & 1]);
      }
    };
    row_stats(0, 0, false);
    row_stats(0, 1, true);
    for (int u = 0; u < n_units; ++u) {
      const int w = (int)blockIdx.x + u * (int)gridDim.x;
      const int h = w % p.heads, n = w / p.heads;
      row_stats(u + 1, 0, false);
      if (NT == 1) row_stats(u + 1, 1, true);
      // rows of the tile start at sequence position seq0; column offset 0 q | D k | 2D v. The 32 staged rows (128-byte
      // swizzle, = the tensor map's) leave as ONE TMA store per warp and tile; rows beyond the sequence are clipped by
      // the map. Lane 0 issues, commits and later waits for the store's shared-memory reads (bulk groups are per thread).
      auto drain = [&](uint32_t col, float sc, int seq0, int coff, uint64_t* release) {
        const int r0 = seq0 + q * 32;
        if (r0 < p.T) {
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // previous tile left the rows
          __syncwarp();
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            uint32_t r[32];
            tmem_ld32(tlane + col + (uint32_t)(hh * 32), r);
            tc_wait_ld();
#pragma unroll
            for (int t = 0; t < 4; ++t)
              asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(my + (uint32_t)lane * 128u +
                                                                        ((((uint32_t)hh * 4 + t) ^ x7) << 4)),
                           "r"(pack_bf16(__uint_as_float(r[8 * t]) * sc, __uint_as_float(r[8 * t + 1]) * sc)),
                           "r"(pack_bf16(__uint_as_float(r[8 * t + 2]) * sc, __uint_as_float(r[8 * t + 3]) * sc)),
                           "r"(pack_bf16(__uint_as_float(r[8 * t + 4]) * sc, __uint_as_float(r[8 * t + 5]) * sc)),
                           "r"(pack_bf16(__uint_as_float(r[8 * t + 6]) * sc, __uint_as_float(r[8 * t + 7]) * sc))
                           : "memory");
          }
          fence_proxy_async();  // the staged rows are read by the TMA unit
        }
        if (release) tc_fence_before();  // the accumulator has been read: the MMAs that overwrite it may go
        __syncwarp();
        if (lane == 0) {
          if (release) mbar_arrive(release);
          if (r0 < p.T) {
            tma_store_3d(&tmOut, my, coff + h * HD, r0, n);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
      };
      for (int j = 0; j < NT; ++j) {
#ifdef MFK_TRACE2
        if (q == 0 && lane == 0) TRC(3);
#endif
        mbar_wait(acc_done, kt_ctr & 1u);
#ifdef MFK_TRACE2
        if (q == 0 && lane == 0) TRC(3);
#endif
        ++kt_ctr;
        tc_fence_after();
        drain(colDK, p.scale, j * 128, D, nullptr);
        drain(colDV, 1.f, j * 128, 2 * D, dkv_free);
        if (j == NT - 1)
          for (int ii = 0; ii < NT; ++ii)
            drain(colDQ + (uint32_t)ii * 64u, p.scale, ii * 128, 0, ii == NT - 1 ? unit_free : nullptr);
        if (j == 0 && NT == 2) row_stats(u + 1, 1, true);  // in the gap before the second key tile completes
#ifdef MFK_TRACE2
        if (q == 0 && lane == 0) TRC(3);
#endif
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // stores complete before the CTA exits
  } else if (warp >= TC_SOFTMAX_WARPS) {
    if (elect_one()) {
      const bool issuer_a = warp == TC_SOFTMAX_WARPS;  // A: TMA loads, S / dP, dQ.   B: dV, dK.
      uint32_t blk_ctr = 0, kt_ctr = 0;
      const uint32_t idesc1 = umma_idesc_bf16(128, 128, 0, 0);
      const uint32_t idesc_tt = umma_idesc_bf16(128, HD, 1, 1);  // A, B MN-major (dV, dK)
      const uint32_t idesc_kt = umma_idesc_bf16(128, HD, 0, 1);  // A K-major, B MN-major (dQ)
      const uint32_t uQ = smem_u32(sQ), uK = smem_u32(sK), uV = smem_u32(sV), uDo = smem_u32(sdO);
      const uint32_t uP = smem_u32(sP), uDs = smem_u32(sdS);
      // Descriptors are built once; per MMA only a 64-bit add remains (the issuing thread is on the critical path).
      // Strides in 16-byte units: a 128-row tile = 1024, a K-major k-step = 2, an MN-major k-step (16 rows) = 128.
      const uint64_t dQk = umma_desc_k_sw128(uQ), dKk = umma_desc_k_sw128(uK), dVk = umma_desc_k_sw128(uV),
                     dDok = umma_desc_k_sw128(uDo), dDsk = umma_desc_k_sw128(uDs);
      const uint64_t dPm = umma_desc_mn_sw128(uP, 16384), dDsm = umma_desc_mn_sw128(uDs, 16384);
      const uint64_t dDom = umma_desc_mn_sw128(uDo, 1024), dQm = umma_desc_mn_sw128(uQ, 1024),
                     dKm = umma_desc_mn_sw128(uK, 1024);
      auto issue_sdp = [&](int i, int j) {
        const uint64_t aq = dQk + 1024ull * i, bk = dKk + 1024ull * j, ad = dDok + 1024ull * i, bv = dVk + 1024ull * j;
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + colS, aq + 2ull * k, bk + 2ull * k, idesc1, k > 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + colDP, ad + 2ull * k, bv + 2ull * k, idesc1, k > 0);
        umma_commit(sd_full);
      };
      // reduction depth in 16-row steps: the last query / key tile holds only T - 128 (NT - 1) valid rows (its staged
      // rows / columns beyond T are zero), so the MMAs over the all-zero tail are not issued. Plain runtime loops:
      // `#pragma unroll` + `if (ks < n)` around the tcgen05.mma asm miscompiled (nvcc 12.9: illegal address at run
      // time even with the predicate always true).
      const int ks_last = (p.T - (NT - 1) * 128 + 15) / 16;
      int trc_n = 0;
      // Operand loads of unit u, in three groups that follow the order in which the previous unit releases the
      // buffers (T > 128: blocks run (i0,j0) (i1,j0) (i0,j1) (i1,j1)): K_0 / V_0 are last read by block 1, Q_0 / dO_0
      // by block 2, the second tiles by block 3 — so only the (short) second tiles are fetched after the unit's last
      // MMA, and they are not needed before block 1 of the next unit.
      auto issue_kv0 = [&](int u) {
        const int w = (int)blockIdx.x + u * (int)gridDim.x;
        const int h = w % p.heads, row0 = (w / p.heads) * p.T;
        mbar_arrive_expect_tx(ld_full, 32768u);
        tma_load_2d(&tmQkv, ld_full, sK, D + h * HD, row0);
        tma_load_2d(&tmQkv, ld_full, sV, 2 * D + h * HD, row0);
      };
      auto issue_qdo0 = [&](int u) {
        const int w = (int)blockIdx.x + u * (int)gridDim.x;
        const int h = w % p.heads, row0 = (w / p.heads) * p.T;
        mbar_arrive_expect_tx(ld_full, 32768u);
        tma_load_2d(&tmQkv, ld_full, sQ, h * HD, row0);
        tma_load_2d(&tmDo, ld_full, sdO, h * HD, row0);
      };
      auto issue_second = [&](int u) {
        if (NT < 2) return;
        const int w = (int)blockIdx.x + u * (int)gridDim.x;
        const int h = w % p.heads, row1 = (w / p.heads) * p.T + 128;
        mbar_arrive_expect_tx(ld2_full, 4u * (uint32_t)R2 * 128u);
        tma_load_2d(&tmQkv2, ld2_full, sQ + 16384, h * HD, row1);
        tma_load_2d(&tmDo2, ld2_full, sdO + 16384, h * HD, row1);
        tma_load_2d(&tmQkv2, ld2_full, sK + 16384, D + h * HD, row1);
        tma_load_2d(&tmQkv2, ld2_full, sV + 16384, 2 * D + h * HD, row1);
      };
      if (issuer_a) {
        if (n_units > 0) {
          issue_kv0(0);
          issue_qdo0(0);
          issue_second(0);
        }
        for (int u = 0; u < n_units; ++u) {
          // this unit's loads were issued early (below); S / dP of its first block only need the S / dP columns, which
          // the compute warps have left (ds_full of the previous unit's last block) — the accumulators of the previous
          // unit may still be draining and are waited for before the first dQ MMA
          TRC(0);  // unit start
          if (u == 0 || NT != 2) {  // (T > 128: issued inside the previous unit's last block, below)
            mbar_wait(ld_full, (uint32_t)u & 1u);
            tc_fence_after();
            issue_sdp(0, 0);
          }
          TRC(0);  // loads landed
          TRC(0);  // S/dP issued
          for (int k = 0; k < nblk; ++k, ++blk_ctr) {
            const int j = k / NT, i = k % NT;
            // S / dP of the NEXT block go out as soon as the compute warps have read this block's S / dP out of TMEM
            // (about half-way through their pass): they are complete when the warps come back for them, and their exp /
            // dS arithmetic then overlaps the second-stage MMAs of this block. After the unit's last block the next
            // block is the first one of the next unit (T > 128: its first operand tiles were requested during blocks 1
            // and 2, below).
            mbar_wait(sdp_consumed, blk_ctr & 1u);
            tc_fence_after();
            if (k + 1 < nblk) {
              if (k == 0) {  // block 1 is the first to touch the second tiles
                mbar_wait(ld2_full, (uint32_t)u & 1u);
                tc_fence_after();
              }
              issue_sdp((k + 1) % NT, (k + 1) / NT);
            } else if (NT == 2 && u + 1 < n_units) {
              mbar_wait(ld_full, (uint32_t)(u + 1) & 1u);
              tc_fence_after();
              issue_sdp(0, 0);
            }
            // the compute warps wait for exactly these MMAs next: issuer B holds its 16 second-stage MMAs of block k
            // back until they are queued
            mbar_arrive(sdp_issued);
            mbar_wait(ds_full, blk_ctr & 1u);
            TRC(0);  // staged operands ready
            tc_fence_after();
            const uint64_t bk = dKm + 1024ull * j;
            const uint32_t acc_j = j > 0;
            const int ks_j = j == NT - 1 ? ks_last : 8;
            if (k == 0) {  // dQ of the previous unit drained
              mbar_wait(unit_free, ((uint32_t)u & 1u) ^ 1u);
              tc_fence_after();
            }
            for (int ks = 0; ks < ks_j; ++ks)  // dQ_i += dS K_j        (K = keys of tile j: 2 column blocks x 4 k-steps)
              umma_bf16(tmem_base + colDQ + (uint32_t)i * 64u, dDsk + 1024ull * (ks >> 2) + 2ull * (ks & 3),
                        bk + 128ull * ks, idesc_kt, acc_j | (ks > 0));
            umma_commit(mma2_done);
            if (i == NT - 1) umma_commit(acc_done);
            TRC(0);  // second-stage (+ next S/dP) issued
            // T > 128, blocks (i0,j0) (i1,j0) (i0,j1) (i1,j1): K_0 / V_0 are last read by block 1, Q_0 / dO_0 by block 2.
            // Their refill with the next unit's rows starts as soon as those MMAs have retired (this thread has nothing
            // else to do until the compute warps are half-way through the next block).
            if (NT == 2 && u + 1 < n_units && (k == 1 || k == 2)) {
              mbar_wait(mma2_done, blk_ctr & 1u);
              if (k == 1) issue_kv0(u + 1);
              else issue_qdo0(u + 1);
            }
          }
          // the second tiles (T > 128) resp. all operands are free once the last block's MMAs of both issuers have retired
          mbar_wait(mma2_done, (blk_ctr - 1u) & 1u);
          if (u + 1 < n_units) {
            if (NT == 2) {
              issue_second(u + 1);
            } else {
              issue_kv0(u + 1);
              issue_qdo0(u + 1);
            }
          }
        }
      } else {
        for (int u = 0; u < n_units; ++u) {
          for (int k = 0; k < nblk; ++k, ++blk_ctr) {
            const int j = k / NT, i = k % NT;
            mbar_wait(ds_full, blk_ctr & 1u);
            mbar_wait(sdp_issued, blk_ctr & 1u);
#ifdef MFK_TRACE2
            TRC(2);
#endif
            if (i == 0 && j > 0) {  // accumulators of the previous key tile must have been drained
              mbar_wait(dkv_free, kt_ctr & 1u);
              ++kt_ctr;
            }
            tc_fence_after();
            const uint64_t bdo = dDom + 1024ull * i, bq = dQm + 1024ull * i;
            const uint32_t acc_i = i > 0;
            const int ks_i = i == NT - 1 ? ks_last : 8;
            for (int ks = 0; ks < ks_i; ++ks)  // dV_j += P^T dO_i   (K = query rows of tile i, 16 per MMA)
              umma_bf16(tmem_base + colDV, dPm + 128ull * ks, bdo + 128ull * ks, idesc_tt, acc_i | (ks > 0));
            for (int ks = 0; ks < ks_i; ++ks)  // dK_j += dS^T Q_i
              umma_bf16(tmem_base + colDK, dDsm + 128ull * ks, bq + 128ull * ks, idesc_tt, acc_i | (ks > 0));
            umma_commit(mma2_done);
            if (i == NT - 1) umma_commit(acc_done);
#ifdef MFK_TRACE2
            TRC(2);
#endif
          }
          // the last key tile's dkv_free arrive is consumed here so the phase counters stay in step
          mbar_wait(dkv_free, kt_ctr & 1u);
          ++kt_ctr;
        }
      }
    }
    __syncwarp();
  } else {
    const int q = warp & 3, half = warp >> 2;
    const int rl = q * 32 + lane;
    const uint32_t x7 = (uint32_t)rl & 7u;
    const uint32_t prow = smem_u32(sP) + (uint32_t)rl * 128u, dsrow = smem_u32(sdS) + (uint32_t)rl * 128u;
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t blk_ctr = 0;
    int trc_n = 0;
    const bool trc_me = warp == 0 && lane == 0;
#ifdef MFK_TRACE2
    int trc4 = 0;
#endif
    for (int u = 0; u < n_units; ++u) {
      mbar_wait(&stats_full[u & 1], ((uint32_t)u >> 1) & 1u);  // row statistics of this unit are in smem (drain warps)
#ifdef MFK_TRACE2
      if (p.trace && trc_me && trc4 < 60) p.trace[4 * 64 + trc4++] = clock64();
#endif
      const float* u_lse = s_lse + (u & 1) * 256;
      const float* u_delta = s_delta + (u & 1) * 256;
      const float Lc[2] = {u_lse[rl], u_lse[128 + rl]}, dcache[2] = {u_delta[rl], u_delta[128 + rl]};
      for (int k = 0; k < nblk; ++k, ++blk_ctr) {
        const int j = k / NT, i = k % NT;
        const int qrow = i * 128 + rl;
        const bool qok = qrow < p.T;
        const float L = qok ? (i == 0 ? Lc[0] : Lc[1]) : INFINITY;  // dead rows: exp2(-inf) = 0
        const float Dl = i == 0 ? dcache[0] : dcache[1];
        if (trc_me) TRC(1);  // about to wait for S/dP
        mbar_wait(sd_full, blk_ctr & 1u);
        if (trc_me) TRC(1);  // S/dP ready
        tc_fence_after();
        // A warp whose 32 query rows all lie beyond T (T = 199: the last quarter of the second query tile) has nothing to
        // stage: those rows of P / dS are only ever read as the K tail that the second-stage MMAs skip (ks_last) or as
        // M rows whose results are discarded.
        const bool rows_live = i * 128 + q * 32 < p.T;
#pragma unroll 1  // (one copy of the two 32-column bodies: the unrolled pair cost instruction-fetch stalls, ncu stall_no_inst)
        for (int cc = 0; cc < 2; ++cc) {
          const int c = half * 32 + cc * 64;
          const int key0 = j * 128 + c;
          // columns beyond T likewise (T = 199: keys 224.. of the second key tile): never read as K, discarded as M
          const bool live = rows_live && key0 < p.T;
          uint32_t pk[16], dk[16];
          if (live) {
            uint32_t rs[32], rd[32];
            tmem_ld32(tlane + colS + (uint32_t)c, rs);
            tmem_ld32(tlane + colDP + (uint32_t)c, rd);
            tc_wait_ld();
            if (cc == 1) {  // this warp is done with the block's S / dP columns
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(sdp_consumed);
            }
            if (key0 + 32 <= p.T) {
              // no masks: L = +inf for query rows beyond T makes their exponentials (and with them dS) exactly 0
#pragma unroll
              for (int jj = 0; jj < 16; ++jj) {
                const float e0 = fast_exp2(__uint_as_float(rs[2 * jj]) * p.c1 - L);
                const float e1 = fast_exp2(__uint_as_float(rs[2 * jj + 1]) * p.c1 - L);
                pk[jj] = pack_bf16(e0, e1);
                dk[jj] = pack_bf16(e0 * (__uint_as_float(rd[2 * jj]) - Dl), e1 * (__uint_as_float(rd[2 * jj + 1]) - Dl));
              }
            } else {
#pragma unroll
              for (int jj = 0; jj < 16; ++jj) {
                const float e0 = key0 + 2 * jj < p.T ? fast_exp2(__uint_as_float(rs[2 * jj]) * p.c1 - L) : 0.f;
                const float e1 = key0 + 2 * jj + 1 < p.T ? fast_exp2(__uint_as_float(rs[2 * jj + 1]) * p.c1 - L) : 0.f;
                pk[jj] = pack_bf16(e0, e1);
                dk[jj] = pack_bf16(e0 * (__uint_as_float(rd[2 * jj]) - Dl), e1 * (__uint_as_float(rd[2 * jj + 1]) - Dl));
              }
            }
          }
          if (cc == 1 && !live) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(sdp_consumed);
          }
          const uint32_t boff = (uint32_t)(c >> 6) * 16384u;
          const uint32_t ch0 = (uint32_t)(c & 63) >> 3;
          // the staged P / dS tiles are still being read by the previous block's second-stage MMAs (which now run
          // concurrently with the arithmetic above): wait for them before the first store of this block
          if (cc == 0 && blk_ctr > 0) {
            if (trc_me) TRC(1);  // first half of the arithmetic done
            mbar_wait(mma2_done, (blk_ctr - 1u) & 1u);
            if (trc_me) TRC(1);  // staged tiles of the previous block released
          }
          if (live) {
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const uint32_t off = boff + (((ch0 + t) ^ x7) << 4);
              asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(prow + off), "r"(pk[4 * t]), "r"(pk[4 * t + 1]),
                           "r"(pk[4 * t + 2]), "r"(pk[4 * t + 3])
                           : "memory");
              asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dsrow + off), "r"(dk[4 * t]), "r"(dk[4 * t + 1]),
                           "r"(dk[4 * t + 2]), "r"(dk[4 * t + 3])
                           : "memory");
            }
          }
        }
        tc_fence_before();
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(ds_full);
        if (trc_me) TRC(1);  // staged
#ifdef MFK_TRACE2
        if (p.trace && k == nblk - 1 && trc_me && trc4 < 60) p.trace[4 * 64 + trc4++] = clock64();
#endif
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == TC_SOFTMAX_WARPS) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}
}  // namespace

// debug hook: device buffer of 2*64 int64 receiving clock64 stamps of CTA 0 (NULL disables)
extern "C" int mfk_debug_set_attn_trace(void* dev_buf) {
  g_attn_trace = static_cast<long long*>(dev_buf);
  return MFK_OK;
}

extern "C" int mfk_attn_bwd_fused(const void* qkv, const void* out, const void* d_out, const float* lse,
                                  float* delta_ws, void* dqkv, int N, int T, int heads, void* stream) {
  if (!qkv || !out || !d_out || !lse || !delta_ws || !dqkv || N <= 0 || T <= 0 || T > 256) return MFK_EARG;
  if (g_attn_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_attn_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_attn_sms <= 0) g_attn_sms = 148;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int D = heads * HD;
  const long long rows = (long long)N * T;
  cudaError_t e;
  (void)delta_ws;  // kept in the signature for symmetry with mfk_attn_bwd; delta is computed inside the kernel
  AttnBwdFusedParams p;
  p.lse = lse; p.out = static_cast<const bf16*>(out); p.d_out = static_cast<const bf16*>(d_out);
  p.dqkv = static_cast<bf16*>(dqkv);
  p.T = T; p.tiles = (T + 127) / 128; p.R = p.tiles * 128; p.heads = heads;
  p.total_units = N * heads;
  p.c1 = 0.125f * kLog2e; p.scale = 0.125f;
  p.trace = g_attn_trace;
  // boxes: 128 rows for the first tile of an operand, 16 * ceil((T - 128) / 16) rows for the second
  CUtensorMap tmQkv, tmDo, tmQkv2, tmDo2;
  int rc;
  if ((rc = mfk_make_tmap_2d(&tmQkv, qkv, 2, (uint64_t)rows, 3ull * D, 3ull * D, 128, 64, 128)) != MFK_OK) return rc;
  if ((rc = mfk_make_tmap_2d(&tmDo, d_out, 2, (uint64_t)rows, (uint64_t)D, (uint64_t)D, 128, 64, 128)) != MFK_OK) return rc;
  tmQkv2 = tmQkv; tmDo2 = tmDo;
  if (p.tiles == 2) {
    const uint32_t r2 = (uint32_t)((T - 128 + 15) / 16 * 16);
    if ((rc = mfk_make_tmap_2d(&tmQkv2, qkv, 2, (uint64_t)rows, 3ull * D, 3ull * D, r2, 64, 128)) != MFK_OK) return rc;
    if ((rc = mfk_make_tmap_2d(&tmDo2, d_out, 2, (uint64_t)rows, (uint64_t)D, (uint64_t)D, r2, 64, 128)) != MFK_OK) return rc;
  }
  CUtensorMap tmOut;
  if ((rc = mfk_make_tmap_bf16_3d(&tmOut, dqkv, (uint64_t)N, (uint64_t)T, 3ull * D, 3ull * D, 32, 64)) != MFK_OK) return rc;
  const size_t smem = 4 * (size_t)p.R * 128 + 65536 + 16384 + 128 + 4096 + 1024;
  e = cudaFuncSetAttribute(attn_bwd_fused_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  const int rounds = (p.total_units + g_attn_sms - 1) / g_attn_sms;  // balanced grid, see mfk_attn_fwd_tc
  const int grid = (p.total_units + rounds - 1) / rounds;
  e = launch_pdl(attn_bwd_fused_tc_kernel, dim3(grid), dim3(FUSED_THREADS), smem, st, tmQkv, tmDo, tmQkv2, tmDo2, tmOut, p);
  if (e != cudaSuccess) return (int)e;
  return MFK_OK;
}

// =====================================================================================================
// fp32 reference-precision attention forward (the "fp32 mode" of the parity contract: logits within 1e-3 of the
// reference's fp32 path). Plain SIMT: one CTA per (sequence, head, 32-query block), K and V of the head in shared
// memory as fp32, one warp per query row at a time (lanes over keys for the scores, lanes over the 64 head dims for
// P V). Throughput is irrelevant here; every product and sum is fp32.
namespace {
constexpr int F32_QBLK = 32, F32_WARPS = 8;
template <bool CAUSAL>
__global__ void __launch_bounds__(F32_WARPS * 32)
attn_fwd_f32_kernel(const float* __restrict__ qkv, float* __restrict__ out, int T, int heads) {
  extern __shared__ float sm_f32[];
  const int D = heads * HD;
  float* sK = sm_f32;                     // [T][65] (padded: conflict-free column reads)
  float* sV = sK + (size_t)T * 65;        // [T][64]
  float* sP = sV + (size_t)T * 64;        // [F32_WARPS][T] probabilities of the row a warp is working on
  const int n = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * F32_QBLK;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* base = qkv + (size_t)n * T * 3 * D + h * HD;
  const int kend = CAUSAL ? min(T, q0 + F32_QBLK) : T;  // keys any row of this block can see
  for (int i = threadIdx.x; i < kend * HD; i += blockDim.x) {
    const int r = i / HD, c = i % HD;
    sK[r * 65 + c] = base[(size_t)r * 3 * D + D + c];
    sV[r * 64 + c] = base[(size_t)r * 3 * D + 2 * D + c];
  }
  __syncthreads();
  float* myP = sP + (size_t)warp * T;
  for (int qi = q0 + warp; qi < min(T, q0 + F32_QBLK); qi += F32_WARPS) {
    const float* qrow = base + (size_t)qi * 3 * D;
    const float q_lo = qrow[lane], q_hi = qrow[lane + 32];
    const int kmax = CAUSAL ? qi + 1 : T;
    float mx = -INFINITY;
    for (int k0 = 0; k0 < kmax; k0 += 32) {
      const int k = k0 + lane;
      const int kk = min(k, kmax - 1);  // clamped: every lane runs the same shuffles, the result is masked below
      float a = 0.f;
#pragma unroll 16
      for (int c = 0; c < 32; ++c) {
        a = fmaf(__shfl_sync(0xffffffffu, q_lo, c), sK[kk * 65 + c], a);
        a = fmaf(__shfl_sync(0xffffffffu, q_hi, c), sK[kk * 65 + 32 + c], a);
      }
      const float sc = k < kmax ? a * 0.125f : -INFINITY;
      if (k < kmax) myP[k] = sc;
      mx = fmaxf(mx, sc);
    }
    mx = warp_max(mx);
    __syncwarp();
    float l = 0.f;
    for (int k = lane; k < kmax; k += 32) {
      const float e = expf(myP[k] - mx);
      myP[k] = e;
      l += e;
    }
    l = warp_sum(l);
    __syncwarp();
    float o_lo = 0.f, o_hi = 0.f;
    for (int k = 0; k < kmax; ++k) {
      const float pk = myP[k];
      o_lo = fmaf(pk, sV[k * 64 + lane], o_lo);
      o_hi = fmaf(pk, sV[k * 64 + 32 + lane], o_hi);
    }
    const float inv = 1.f / l;
    float* orow = out + ((size_t)n * T + qi) * D + h * HD;
    orow[lane] = o_lo * inv;
    orow[lane + 32] = o_hi * inv;
    __syncwarp();
  }
}
}  // namespace

extern "C" int mfk_attn_fwd_f32(const float* qkv, float* out, int N, int T, int heads, int causal, void* stream) {
  if (!qkv || !out || N <= 0 || T <= 0 || T > 512 || heads <= 0) return MFK_EARG;
  const size_t smem = ((size_t)T * 65 + (size_t)T * 64 + (size_t)F32_WARPS * T) * sizeof(float);
  dim3 grid((T + F32_QBLK - 1) / F32_QBLK, heads, N);
  cudaError_t e;
  if (causal) {
    e = cudaFuncSetAttribute(attn_fwd_f32_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    attn_fwd_f32_kernel<true><<<grid, F32_WARPS * 32, smem, static_cast<cudaStream_t>(stream)>>>(qkv, out, T, heads);
  } else {
    e = cudaFuncSetAttribute(attn_fwd_f32_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    attn_fwd_f32_kernel<false><<<grid, F32_WARPS * 32, smem, static_cast<cudaStream_t>(stream)>>>(qkv, out, T, heads);
  }
  MFK_CHECK_LAUNCH();
  return MFK_OK;
}

// =====================================================================================================
// fp32 attention BACKWARD of the fp32 training mode (cfg PREC = "fp32": the reference trains its fp32 model,
// trainers/maple.py:438-439, 590). Same SIMT organisation as the forward above, two deterministic kernels:
//   dq  kernel, CTA = (sequence, head, 32-query block), K and V of the head in smem: per query row the softmax is
//       recomputed (row max, sum), dP_j = dO . V_j, delta = sum_j P_j dP_j, dS_j = P_j (dP_j - delta),
//       dQ = sum_j dS_j K_j / 8; it also leaves lse (natural log) and delta of every row in scratch;
//   dkv kernel, CTA = (sequence, head, 32-key block), Q and dO of the head in smem: per key row
//       P_i = exp(q_i . k / 8 - lse_i), dS_i = P_i (dO_i . v - delta_i), dV = sum_i P_i dO_i, dK = sum_i dS_i q_i / 8.
// Sums run in index order inside one warp: bit-reproducible.
namespace {
template <bool CAUSAL>
__global__ void __launch_bounds__(F32_WARPS * 32)
attn_bwd_dq_f32_kernel(const float* __restrict__ qkv, const float* __restrict__ d_out, float* __restrict__ dqkv,
                       float* __restrict__ lse, float* __restrict__ delta, int T, int heads) {
  extern __shared__ float sm_f32[];
  const int D = heads * HD;
  float* sK = sm_f32;                      // [T][65]
  float* sV = sK + (size_t)T * 65;         // [T][65]
  float* sP = sV + (size_t)T * 65;         // [F32_WARPS][2][T]: probabilities and dP of the row a warp works on
  const int n = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * F32_QBLK;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* base = qkv + (size_t)n * T * 3 * D + h * HD;
  const int kend = CAUSAL ? min(T, q0 + F32_QBLK) : T;
  for (int i = threadIdx.x; i < kend * HD; i += blockDim.x) {
    const int r = i / HD, c = i % HD;
    sK[r * 65 + c] = base[(size_t)r * 3 * D + D + c];
    sV[r * 65 + c] = base[(size_t)r * 3 * D + 2 * D + c];
  }
  __syncthreads();
  float* myP = sP + (size_t)warp * 2 * T;
  float* myDP = myP + T;
  for (int qi = q0 + warp; qi < min(T, q0 + F32_QBLK); qi += F32_WARPS) {
    const float* qrow = base + (size_t)qi * 3 * D;
    const float q_lo = qrow[lane], q_hi = qrow[lane + 32];
    const float* dorow = d_out + ((size_t)n * T + qi) * D + h * HD;
    const float do_lo = dorow[lane], do_hi = dorow[lane + 32];
    const int kmax = CAUSAL ? qi + 1 : T;
    float mx = -INFINITY;
    for (int k0 = 0; k0 < kmax; k0 += 32) {
      const int k = k0 + lane;
      const int kk = min(k, kmax - 1);
      float a = 0.f, b = 0.f;
#pragma unroll 16
      for (int c = 0; c < 32; ++c) {
        a = fmaf(__shfl_sync(0xffffffffu, q_lo, c), sK[kk * 65 + c], a);
        a = fmaf(__shfl_sync(0xffffffffu, q_hi, c), sK[kk * 65 + 32 + c], a);
        b = fmaf(__shfl_sync(0xffffffffu, do_lo, c), sV[kk * 65 + c], b);
        b = fmaf(__shfl_sync(0xffffffffu, do_hi, c), sV[kk * 65 + 32 + c], b);
      }
      const float sc = k < kmax ? a * 0.125f : -INFINITY;
      if (k < kmax) {
        myP[k] = sc;
        myDP[k] = b;
      }
      mx = fmaxf(mx, sc);
    }
    mx = warp_max(mx);
    __syncwarp();
    float l = 0.f;
    for (int k = lane; k < kmax; k += 32) {
      const float e = expf(myP[k] - mx);
      myP[k] = e;
      l += e;
    }
    l = warp_sum(l);
    const float inv = 1.f / l;
    float dl = 0.f;
    for (int k = lane; k < kmax; k += 32) {  // each lane revisits the entries it wrote itself
      const float pn = myP[k] * inv;
      myP[k] = pn;
      dl = fmaf(pn, myDP[k], dl);
    }
    dl = warp_sum(dl);
    __syncwarp();
    float dq_lo = 0.f, dq_hi = 0.f;
    for (int k = 0; k < kmax; ++k) {
      const float ds = myP[k] * (myDP[k] - dl);
      dq_lo = fmaf(ds, sK[k * 65 + lane], dq_lo);
      dq_hi = fmaf(ds, sK[k * 65 + 32 + lane], dq_hi);
    }
    float* dqrow = dqkv + ((size_t)n * T + qi) * 3 * D + h * HD;
    dqrow[lane] = dq_lo * 0.125f;
    dqrow[lane + 32] = dq_hi * 0.125f;
    if (lane == 0) {
      lse[((size_t)n * heads + h) * T + qi] = mx + logf(l);
      delta[((size_t)n * heads + h) * T + qi] = dl;
    }
    __syncwarp();
  }
}

template <bool CAUSAL>
__global__ void __launch_bounds__(F32_WARPS * 32)
attn_bwd_dkv_f32_kernel(const float* __restrict__ qkv, const float* __restrict__ d_out, float* __restrict__ dqkv,
                        const float* __restrict__ lse, const float* __restrict__ delta, int T, int heads) {
  extern __shared__ float sm_f32[];
  const int D = heads * HD;
  float* sQ = sm_f32;                      // [T][65]
  float* sdO = sQ + (size_t)T * 65;        // [T][65]
  float* sL = sdO + (size_t)T * 65;        // [T] lse
  float* sDl = sL + T;                     // [T] delta
  float* sP = sDl + T;                     // [F32_WARPS][2][T]
  const int n = blockIdx.z, h = blockIdx.y, j0 = blockIdx.x * F32_QBLK;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* base = qkv + (size_t)n * T * 3 * D + h * HD;
  const int qbeg = CAUSAL ? j0 : 0;  // queries before the block's first key see none of its keys
  for (int i = threadIdx.x + qbeg * HD; i < T * HD; i += blockDim.x) {
    const int r = i / HD, c = i % HD;
    sQ[r * 65 + c] = base[(size_t)r * 3 * D + c];
    sdO[r * 65 + c] = d_out[((size_t)n * T + r) * D + h * HD + c];
  }
  for (int i = threadIdx.x + qbeg; i < T; i += blockDim.x) {
    sL[i] = lse[((size_t)n * heads + h) * T + i];
    sDl[i] = delta[((size_t)n * heads + h) * T + i];
  }
  __syncthreads();
  float* myP = sP + (size_t)warp * 2 * T;
  float* myDS = myP + T;
  for (int kj = j0 + warp; kj < min(T, j0 + F32_QBLK); kj += F32_WARPS) {
    const float* krow = base + (size_t)kj * 3 * D + D;
    const float k_lo = krow[lane], k_hi = krow[lane + 32];
    const float v_lo = krow[D + lane], v_hi = krow[D + lane + 32];
    const int imin = CAUSAL ? kj : 0;
    for (int i0 = imin; i0 < T; i0 += 32) {
      const int i = i0 + lane;
      const int ii = min(i, T - 1);
      float a = 0.f, b = 0.f;
#pragma unroll 16
      for (int c = 0; c < 32; ++c) {
        a = fmaf(__shfl_sync(0xffffffffu, k_lo, c), sQ[ii * 65 + c], a);
        a = fmaf(__shfl_sync(0xffffffffu, k_hi, c), sQ[ii * 65 + 32 + c], a);
        b = fmaf(__shfl_sync(0xffffffffu, v_lo, c), sdO[ii * 65 + c], b);
        b = fmaf(__shfl_sync(0xffffffffu, v_hi, c), sdO[ii * 65 + 32 + c], b);
      }
      if (i < T) {
        const float pr = expf(a * 0.125f - sL[i]);
        myP[i] = pr;
        myDS[i] = pr * (b - sDl[i]);
      }
    }
    __syncwarp();
    float dv_lo = 0.f, dv_hi = 0.f, dk_lo = 0.f, dk_hi = 0.f;
    for (int i = imin; i < T; ++i) {
      const float pr = myP[i], ds = myDS[i];
      dv_lo = fmaf(pr, sdO[i * 65 + lane], dv_lo);
      dv_hi = fmaf(pr, sdO[i * 65 + 32 + lane], dv_hi);
      dk_lo = fmaf(ds, sQ[i * 65 + lane], dk_lo);
      dk_hi = fmaf(ds, sQ[i * 65 + 32 + lane], dk_hi);
    }
    float* dkrow = dqkv + ((size_t)n * T + kj) * 3 * D + D + h * HD;
    dkrow[lane] = dk_lo * 0.125f;
    dkrow[lane + 32] = dk_hi * 0.125f;
    dkrow[D + lane] = dv_lo;
    dkrow[D + lane + 32] = dv_hi;
    __syncwarp();
  }
}
}  // namespace

extern "C" int mfk_attn_bwd_f32(const float* qkv, const float* d_out, float* dqkv, float* stat_ws, int N, int T,
                                int heads, int causal, void* stream) {
  if (!qkv || !d_out || !dqkv || !stat_ws || N <= 0 || T <= 0 || T > 256 || heads <= 0) return MFK_EARG;
  const size_t smem_dq = ((size_t)2 * T * 65 + (size_t)F32_WARPS * 2 * T) * sizeof(float);
  const size_t smem_dkv = ((size_t)2 * T * 65 + 2 * (size_t)T + (size_t)F32_WARPS * 2 * T) * sizeof(float);
  float* lse = stat_ws;
  float* delta = stat_ws + (size_t)N * heads * T;
  dim3 grid((T + F32_QBLK - 1) / F32_QBLK, heads, N);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e;
#define MFK_ATTN_BWD_F32(C_)                                                                                          \
  e = cudaFuncSetAttribute(attn_bwd_dq_f32_kernel<C_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_dq);    \
  if (e != cudaSuccess) return (int)e;                                                                                \
  e = cudaFuncSetAttribute(attn_bwd_dkv_f32_kernel<C_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_dkv);  \
  if (e != cudaSuccess) return (int)e;                                                                                \
  attn_bwd_dq_f32_kernel<C_><<<grid, F32_WARPS * 32, smem_dq, st>>>(qkv, d_out, dqkv, lse, delta, T, heads);          \
  attn_bwd_dkv_f32_kernel<C_><<<grid, F32_WARPS * 32, smem_dkv, st>>>(qkv, d_out, dqkv, lse, delta, T, heads);
  if (causal) {
    MFK_ATTN_BWD_F32(true)
  } else {
    MFK_ATTN_BWD_F32(false)
  }
#undef MFK_ATTN_BWD_F32
  MFK_CHECK_LAUNCH();
  return MFK_OK;
}

// =====================================================================================================
// Single-query attention for the LAST block of a tower: only one row per sequence (CLS / EOT) of the block output
// is consumed (clip/model.py:567, trainers/maple.py:72-76), so the attention core is needed for that query row
// only — Q of one row against K, V of the whole sequence — and so is its backward: dQ lives on that row, dK / dV are
// rank-1 updates. One CTA per (head, sequence), 4 warps, fp32 arithmetic, fixed reduction orders (deterministic).
//   forward : out_r[n, h*64..] = softmax(q_r K^T / 8 [causal: keys <= r]) V ;  lse_r[n, h] = log-sum-exp (natural)
//   backward: dqkv[n*T + j] for every j: dV_j = p_j dO, dK_j = dS_j q_r / 8, dQ_j = 0 except row r: sum_j dS_j k_j / 8
namespace {
constexpr int ROWS_WARPS = 8;   // 8 warps: the phases are latency-bound serial loops over the keys (4 -> 8 warps: 28.7 -> ~17 us backward)
constexpr int ROWS_PARTS = ROWS_WARPS * 32 / HD;  // key partitions of rows_wsum

__device__ __forceinline__ float block_reduce_rows(float v, float* red, bool is_max) {
  v = is_max ? warp_max(v) : warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = red[0];
  for (int w = 1; w < ROWS_WARPS; ++w) r = is_max ? fmaxf(r, red[w]) : r + red[w];
  return r;
}

// K and V slices (64 bf16 = 128 bytes per row) of this (sequence, head) -> shared memory, coalesced (8 lanes per row,
// 16 rows per block-wide instruction, every load independent: one memory round trip instead of one per key)
__device__ __forceinline__ void rows_stage_kv(const bf16* seq_base, int D3, int D, int T, uint4* sK, uint4* sV) {
  for (int i = threadIdx.x; i < T * 8; i += blockDim.x) {
    const int j = i >> 3, ch = i & 7;
    sK[i] = __ldg(reinterpret_cast<const uint4*>(seq_base + (size_t)j * D3 + D) + ch);
    sV[i] = __ldg(reinterpret_cast<const uint4*>(seq_base + (size_t)j * D3 + 2 * D) + ch);
  }
}
// s[j] = scale * <vec, M_j> for j < kmax, M = staged [T][64] bf16; a warp takes keys w, w + 4, ... (2 dims per lane)
__device__ __forceinline__ void rows_dots(const bf16* M, const float* vec, float* s, int kmax, float scale) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float v0 = vec[2 * lane], v1 = vec[2 * lane + 1];
  for (int j = warp; j < kmax; j += ROWS_WARPS) {
    const float2 k2 = unpack_bf16(*reinterpret_cast<const uint32_t*>(M + (size_t)j * HD + 2 * lane));
    const float d = warp_sum(v0 * k2.x + v1 * k2.y);
    if (lane == 0) s[j] = d * scale;
  }
}
// out[d] = sum_{j < kmax} w[j] * M_j[d]: thread t owns dim t & 63 and the keys j = t >> 6 (mod ROWS_PARTS); fixed-order combine
__device__ __forceinline__ float rows_wsum(const bf16* M, const float* w, int kmax, float* comb) {
  const int d = threadIdx.x & 63, part = threadIdx.x >> 6;
  float acc = 0.f;
  for (int j = part; j < kmax; j += ROWS_PARTS) acc = fmaf(w[j], __bfloat162float(M[(size_t)j * HD + d]), acc);
  __syncthreads();
  comb[threadIdx.x] = acc;
  __syncthreads();
  float r = comb[d];
#pragma unroll
  for (int q = 1; q < ROWS_PARTS; ++q) r += comb[q * HD + d];   // fixed order
  return r;
}

__global__ void __launch_bounds__(ROWS_WARPS * 32)
attn_rows_fwd_kernel(const bf16* __restrict__ qkv, const int* __restrict__ rows, bf16* __restrict__ out_r,
                     float* __restrict__ lse_r, int T, int heads, int causal) {
  extern __shared__ uint4 rows_smem[];
  __shared__ float q[HD], s[256], red[ROWS_WARPS], comb[ROWS_WARPS * 32];
  uint4* sK = rows_smem;
  uint4* sV = rows_smem + (size_t)T * 8;
  pdl_trigger();
  pdl_wait();
  const int h = blockIdx.x, n = blockIdx.y, D = heads * HD, D3 = 3 * D;
  const int r = rows[n] - n * T;
  const int kmax = causal ? r + 1 : T;
  const bf16* seq = qkv + (size_t)n * T * D3 + h * HD;
  rows_stage_kv(seq, D3, D, kmax, sK, sV);
  if (threadIdx.x < HD) q[threadIdx.x] = __bfloat162float(seq[(size_t)r * D3 + threadIdx.x]);
  __syncthreads();
  rows_dots(reinterpret_cast<const bf16*>(sK), q, s, kmax, 0.125f);
  __syncthreads();
  float mx = -INFINITY;
  for (int j = threadIdx.x; j < kmax; j += blockDim.x) mx = fmaxf(mx, s[j]);
  mx = block_reduce_rows(mx, red, true);
  float l = 0.f;
  for (int j = threadIdx.x; j < kmax; j += blockDim.x) {
    const float e = expf(s[j] - mx);
    s[j] = e;
    l += e;
  }
  l = block_reduce_rows(l, red, false);
  const float inv = 1.f / l;
  for (int j = threadIdx.x; j < kmax; j += blockDim.x) s[j] *= inv;
  __syncthreads();
  const float o = rows_wsum(reinterpret_cast<const bf16*>(sV), s, kmax, comb);
  if (threadIdx.x < HD) out_r[(size_t)n * D + h * HD + threadIdx.x] = __float2bfloat16_rn(o);
  if (threadIdx.x == 0 && lse_r) lse_r[(size_t)n * heads + h] = mx + logf(l);
}

__global__ void __launch_bounds__(ROWS_WARPS * 32)
attn_rows_bwd_kernel(const bf16* __restrict__ qkv, const int* __restrict__ rows, const bf16* __restrict__ d_out_r,
                     const float* __restrict__ lse_r, bf16* __restrict__ dqkv, int T, int heads, int causal) {
  extern __shared__ uint4 rows_smem[];
  __shared__ float q[HD], da[HD], pr[256], ds[256], red[ROWS_WARPS], comb[ROWS_WARPS * 32];
  uint4* sK = rows_smem;
  uint4* sV = rows_smem + (size_t)T * 8;
  pdl_trigger();
  pdl_wait();
  const int h = blockIdx.x, n = blockIdx.y, D = heads * HD, D3 = 3 * D;
  const int r = rows[n] - n * T;
  const int kmax = causal ? r + 1 : T;
  const bf16* seq = qkv + (size_t)n * T * D3 + h * HD;
  bf16* dseq = dqkv + (size_t)n * T * D3 + h * HD;
  rows_stage_kv(seq, D3, D, kmax, sK, sV);
  if (threadIdx.x < HD) {
    q[threadIdx.x] = __bfloat162float(seq[(size_t)r * D3 + threadIdx.x]);
    da[threadIdx.x] = __bfloat162float(d_out_r[(size_t)n * D + h * HD + threadIdx.x]);
  }
  __syncthreads();
  rows_dots(reinterpret_cast<const bf16*>(sK), q, pr, kmax, 0.125f);   // scores
  rows_dots(reinterpret_cast<const bf16*>(sV), da, ds, kmax, 1.0f);    // dP_j = <dO, v_j>
  __syncthreads();
  const float lse = lse_r[(size_t)n * heads + h];
  float dl = 0.f;
  for (int j = threadIdx.x; j < kmax; j += blockDim.x) {
    const float p = expf(pr[j] - lse);
    pr[j] = p;
    dl += p * ds[j];
  }
  const float delta = block_reduce_rows(dl, red, false);
  for (int j = threadIdx.x; j < kmax; j += blockDim.x) ds[j] = pr[j] * (ds[j] - delta);
  __syncthreads();
  const float dq = 0.125f * rows_wsum(reinterpret_cast<const bf16*>(sK), ds, kmax, comb);   // dQ_r[d]
  __syncthreads();
  if (threadIdx.x < HD) comb[threadIdx.x] = dq;
  __syncthreads();
  // every row of this (sequence, head): dQ (zero except row r), dK_j, dV_j — a warp per key, 2 dims per lane
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float q0 = q[2 * lane] * 0.125f, q1 = q[2 * lane + 1] * 0.125f, a0 = da[2 * lane], a1 = da[2 * lane + 1];
  for (int j = warp; j < T; j += ROWS_WARPS) {
    const float p = j < kmax ? pr[j] : 0.f, g = j < kmax ? ds[j] : 0.f;
    bf16* row = dseq + (size_t)j * D3;
    *reinterpret_cast<uint32_t*>(row + 2 * lane) = j == r ? pack_bf16(comb[2 * lane], comb[2 * lane + 1]) : 0u;
    *reinterpret_cast<uint32_t*>(row + D + 2 * lane) = pack_bf16(g * q0, g * q1);
    *reinterpret_cast<uint32_t*>(row + 2 * D + 2 * lane) = pack_bf16(p * a0, p * a1);
  }
}
}  // namespace

static int rows_smem_attr(const void* fn, size_t smem) {
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  return e == cudaSuccess ? MFK_OK : (int)e;
}

extern "C" int mfk_attn_rows_fwd(const void* qkv, const int* rows, void* out_rows, float* lse_rows, int N, int T,
                                 int heads, int causal, void* stream) {
  if (!qkv || !rows || !out_rows || N <= 0 || T <= 0 || T > 256 || heads <= 0) return MFK_EARG;
  const size_t smem = (size_t)T * 256;  // K and V slices of one (sequence, head)
  if (int rc = rows_smem_attr(reinterpret_cast<const void*>(attn_rows_fwd_kernel), smem)) return rc;
  launch_pdl(attn_rows_fwd_kernel, dim3(heads, N), dim3(ROWS_WARPS * 32), smem, static_cast<cudaStream_t>(stream),
             static_cast<const bf16*>(qkv), rows, static_cast<bf16*>(out_rows), lse_rows, T, heads, causal);
  MFK_CHECK_LAUNCH();
  return MFK_OK;
}

extern "C" int mfk_attn_rows_bwd(const void* qkv, const int* rows, const void* d_out_rows, const float* lse_rows,
                                 void* dqkv, int N, int T, int heads, int causal, void* stream) {
  if (!qkv || !rows || !d_out_rows || !lse_rows || !dqkv || N <= 0 || T <= 0 || T > 256 || heads <= 0) return MFK_EARG;
  const size_t smem = (size_t)T * 256;
  if (int rc = rows_smem_attr(reinterpret_cast<const void*>(attn_rows_bwd_kernel), smem)) return rc;
  launch_pdl(attn_rows_bwd_kernel, dim3(heads, N), dim3(ROWS_WARPS * 32), smem, static_cast<cudaStream_t>(stream),
             static_cast<const bf16*>(qkv), rows, static_cast<const bf16*>(d_out_rows), lse_rows,
             static_cast<bf16*>(dqkv), T, heads, causal);
  MFK_CHECK_LAUNCH();
  return MFK_OK;
}
