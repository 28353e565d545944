// tcgen05 / TMEM / TMA GEMM with fused epilogues for the MaPLe towers (sm_100a).
//
//   C[M,N] = epilogue( A[M,K] * B[N,K]^T )       A, B bf16 row-major (K contiguous), fp32 accumulate
//
// One kernel covers the reference's call sites (SURVEY.md §2.2 K1,K4,K6,K7,K8,K12,K13 and their
// dgrad / wgrad forms, K15): nn.MultiheadAttention in_proj / out_proj (clip/model.py:274,303-305,350),
// mlp.c_fc + QuickGELU and mlp.c_proj (clip/model.py:276-280,351), conv1-as-GEMM (clip/model.py:514),
// `@ proj` / `@ text_projection` (clip/model.py:569-570, trainers/maple.py:76).
//
// Design: persistent, warp-specialised. warp0 = TMA producer (one thread), warp1 = tcgen05.mma issuer
// (one thread), warp2 = TMEM allocator, warps4-11 = epilogue (one TMEM lane == one output row per thread; two
// warps per lane quarter split the columns). Epilogue inputs (bias / residual / QuickGELU aux) are fetched while
// the accumulator is still in flight; outputs are staged per warp in swizzled smem and written with TMA stores
// (coalesced, M/N tails clipped by the tensor map).
// smem ring of kStages x {A 128x64, B BNx64} bf16 tiles in 128-byte swizzle, filled by TMA and consumed
// straight by UMMA descriptors; two 256-column TMEM accumulators so the epilogue of tile i overlaps the
// MMAs of tile i+1. The last partial wave of tiles is split into narrower tiles (runtime UMMA N) so the
// tail spreads over all SMs (50 M-tiles x a few N-tiles quantises badly on 148 SMs otherwise).
#include "mfk_common.cuh"
#include "../../include/mfk.h"

namespace {

using namespace mfk;

constexpr int BM = 128;       // UMMA M (cta_group::1)
constexpr int BK = 64;        // 64 bf16 = 128 B = one swizzle row
constexpr int UK = 16;        // UMMA K for 16-bit inputs
constexpr int BOXN = 64;      // B rows per TMA box (tiles are 64/128/256 wide)
constexpr int kThreads = 384;
constexpr int kEpiWarps = 8;
constexpr int kStageBufBytes = 4096;  // per epilogue warp: 32 rows x 128 B
constexpr int kTmemCols = 512;
constexpr int kSplitKCounterBytes = 4096;  // head of the split-K workspace: 2 x rem int counters

struct GemmParams {
  int M, N, K;
  int n_big;        // number of full-width N tiles
  int bn;           // full tile width (128 or 256)
  int full_tiles;   // tiles [0, full_tiles) are full width
  int split;        // each remaining big tile is split into `split` tiles of width bn/split
  int total_tiles;
  int cluster;      // 2: CTA-pair kernel (cta_group::2); tile indices then count PAIRS of M tiles. 0: single CTAs
  int splitk;       // > 0: each remaining big tile is split along K into `splitk` slices (units), one per CTA; the
  int kb_per;       //      slices (kb_per k-blocks each) leave fp32 partials in `ws` and are summed in slice order
  float* ws;        // split-K partials: [unit][128][256] fp32
  int* ws_cnt;      // split-K counters: [0, rem) partials written, [rem, 2 rem) fix-ups done (all zero between launches)
  const float* bias;
  int act;          // 0 none, 1 QuickGELU, 2 multiply by QuickGELU'(aux)
  const bf16* aux;
  long long ldaux;
  const float* res;
  long long ldres;
  float* out32;
  long long ld32;
  bf16* out16;
  long long ld16;
  bf16* outpre;
  long long ldpre;
  long long* trace;  // optional debug: per-CTA clock64 stamps of the three roles (tools/gemm_trace.py); NULL in production
};
// trace layout: [CTA][role 0 producer | 1 mma | 2 epilogue warp 4][64]
#ifndef MFK_GTRACE  // phase tracing is compiled in only for tools/gemm_trace*.py (build with MFK_DEFS=-DMFK_GTRACE)
#define GTRACE(role, n) do { } while (0)
#else
#define GTRACE(role, n)                                                                         \
  do {                                                                                          \
    if (p.trace && (n) < 64) p.trace[((size_t)blockIdx.x * 3 + (role)) * 64 + (n)] = clock64(); \
  } while (0)
#endif


// split-K unit of tile index t: slice number (or -1), remainder-tile index and k-block range
__device__ __forceinline__ int decode_slice(const GemmParams& p, int t, int num_kb, int& rem_idx, int& kb0, int& kb1) {
  kb0 = 0;
  kb1 = num_kb;
  rem_idx = 0;
  if (p.splitk == 0 || t < p.full_tiles) return -1;
  const int u = t - p.full_tiles;
  rem_idx = u / p.splitk;
  const int slice = u % p.splitk;
  kb0 = slice * p.kb_per;
  kb1 = min(num_kb, kb0 + p.kb_per);
  return slice;
}

__device__ __forceinline__ void decode_tile(const GemmParams& p, int t, int rank, int& m0, int& n0, int& w) {
  int big, sub = 0;
  if (t < p.full_tiles) {
    big = t;
    w = p.bn;
  } else if (p.splitk) {
    big = p.full_tiles + (t - p.full_tiles) / p.splitk;
    w = p.bn;
  } else {
    int r = t - p.full_tiles;
    big = p.full_tiles + r / p.split;
    sub = r % p.split;
    w = p.bn / p.split;
  }
  m0 = ((big / p.n_big) * (p.cluster ? 2 : 1) + rank) * BM;
  n0 = (big % p.n_big) * p.bn + sub * w;
}

// Per-warp epilogue context: output tensor maps, this warp's staging buffer and the swizzled offsets of this
// lane's row inside it.
struct EpiCtx {
  const GemmParams& p;
  const CUtensorMap* tmO32;
  const CUtensorMap* tmO16;
  const CUtensorMap* tmPre;
  uint8_t* sbuf;
  uint32_t row128, x128, row64a, x64, row64b;
  bool has_pre;
  int lane;
};

// Epilogue operands of one 32-row x 32-column chunk, fetched COALESCED: a lane that reads "its own row" touches 32
// different 128-byte lines per load instruction (32 memory wavefronts each; the fp32-residual epilogue spent most
// of its time there). Instead lane l loads, for j = 0..7, the float4 at (row 4j + l/8, columns 4(l%8)..) — 4 whole
// rows per instruction — and finish_chunk transposes through the warp's staging buffer. Same idea for the bf16
// QuickGELU aux (row 8j + l/4, 16-byte chunk l%4, j = 0..3).
__device__ __forceinline__ void load_res_chunk(const GemmParams& p, int row0, int col, int lane, float4 (&rv)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int r = row0 + 4 * j + (lane >> 3);
    rv[j] = r < p.M ? *reinterpret_cast<const float4*>(p.res + (size_t)r * p.ldres + col + 4 * (lane & 7))
                    : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
__device__ __forceinline__ void load_aux_chunk(const GemmParams& p, int row0, int col, int lane, uint4 (&av)[4]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int r = row0 + 8 * j + (lane >> 2);
    av[j] = r < p.M ? __ldg(reinterpret_cast<const uint4*>(p.aux + (size_t)r * p.ldaux + col + 8 * (lane & 3)))
                    : make_uint4(0u, 0u, 0u, 0u);
  }
}

// Epilogue math + stores of one 32-row x 32-column chunk held as v[32] (this lane's row): bias, QuickGELU
// (+ pre-activation), * QuickGELU'(aux), + residual; outputs staged in swizzled smem and written by TMA.
// row0 = first row of the 32-row group, col = first column, ok = this lane's element range is inside [M, N).
template <int EPI>
__device__ __forceinline__ void finish_chunk(const EpiCtx& c, float (&v)[32], const float4 (&rv)[EPI == 3 ? 8 : 1],
                                             const uint4 (&av)[EPI == 2 ? 4 : 1], int row0, int col, bool ok,
                                             uint32_t pp, const float4* bpre = nullptr) {
  if (EPI != 2 && c.p.bias && col < c.p.N) {
    // bpre: the chunk's 32 bias values, fetched by the caller ahead of the TMEM load (their L2 latency then hides
    // behind it instead of following it)
    const float4* b4 = reinterpret_cast<const float4*>(c.p.bias + col);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 bv = bpre ? bpre[j] : __ldg(b4 + j);
      v[4 * j] += bv.x; v[4 * j + 1] += bv.y; v[4 * j + 2] += bv.z; v[4 * j + 3] += bv.w;
    }
  }
  // A single bf16 output needs only half of the 4 KB staging buffer per chunk: the two halves are used in turn
  // (pp = chunk parity) and only the store before the previous one must have finished READING its half, so the TMA
  // store of chunk i overlaps the arithmetic of chunk i + 1. Otherwise the whole buffer is reused every chunk.
  const GemmParams& p = c.p;
  const bool trc = (threadIdx.x == 128) && pp == 4;  // warp 4 lane 0, first chunk of its second tile
  if (trc) GTRACE(2, 32);
  const bool pingpong = EPI != 1 && EPI != 3 && c.p.out16 && !c.p.out32;
  const uint32_t poff = pingpong ? (pp & 1u) * 2048u : 0u;
  if (pingpong) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
  else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  __syncwarp();
  if (trc) GTRACE(2, 33);
  if (EPI == 1) {
    // bf16-round the pre-activation ONCE with the packed (ALU-pipe) conversion: the same words are stored as the
    // saved pre-activation and unpacked by shifts as the activation's input. The scalar round trip
    // (F2F.BF16.F32 per element) runs on the quarter-rate XU pipe that the 32 tanh of the chunk already saturate.
    uint32_t pk[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) pk[j] = pack_bf16(v[2 * j], v[2 * j + 1]);
    if (c.has_pre) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(c.row64b + ((j ^ c.x64) << 4)), "r"(pk[4 * j]),
                     "r"(pk[4 * j + 1]), "r"(pk[4 * j + 2]), "r"(pk[4 * j + 3])
                     : "memory");
    }
    // the activation is applied to the bf16-rounded pre-activation that backward will re-read
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      v[2 * j] = quickgelu(__uint_as_float(pk[j] << 16));
      v[2 * j + 1] = quickgelu(__uint_as_float(pk[j] & 0xffff0000u));
    }
  }
  const uint32_t sb = c.row128 - (uint32_t)c.lane * 128u;  // this warp's staging buffer
  if constexpr (EPI == 2) {
    // aux arrives as (row 8j + lane/4, 16-byte chunk lane%4): park it in the staging half this chunk will use for its
    // output (64-byte rows, SWIZZLE_64B pattern), then every lane picks up its own row
    const uint32_t ab = sb + poff;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t rr = 8u * j + ((uint32_t)c.lane >> 2);
      asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(ab + rr * 64u + ((((uint32_t)c.lane & 3u) ^ ((rr >> 1) & 3u)) << 4)),
                   "r"(av[j].x), "r"(av[j].y), "r"(av[j].z), "r"(av[j].w)
                   : "memory");
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint4 a;
      asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];"
                   : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w)
                   : "r"(c.row64a + poff + ((j ^ c.x64) << 4)));
      float2 u0 = unpack_bf16(a.x), u1 = unpack_bf16(a.y), u2 = unpack_bf16(a.z), u3 = unpack_bf16(a.w);
      v[8 * j] *= dquickgelu(u0.x); v[8 * j + 1] *= dquickgelu(u0.y);
      v[8 * j + 2] *= dquickgelu(u1.x); v[8 * j + 3] *= dquickgelu(u1.y);
      v[8 * j + 4] *= dquickgelu(u2.x); v[8 * j + 5] *= dquickgelu(u2.y);
      v[8 * j + 6] *= dquickgelu(u3.x); v[8 * j + 7] *= dquickgelu(u3.y);
    }
  }
  if constexpr (EPI == 3) {
    // residual arrives as (row 4j + lane/8, 16-byte chunk lane%8): through the staging buffer (128-byte rows,
    // SWIZZLE_128B pattern = the layout of the fp32 output image written below), then every lane adds its own row
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t rr = 4u * j + ((uint32_t)c.lane >> 3);
      asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(sb + rr * 128u + ((((uint32_t)c.lane & 7u) ^ (rr & 7u)) << 4)),
                   "f"(rv[j].x), "f"(rv[j].y), "f"(rv[j].z), "f"(rv[j].w)
                   : "memory");
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float4 x;
      asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                   : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w)
                   : "r"(c.row128 + ((j ^ c.x128) << 4)));
      v[4 * j] += x.x; v[4 * j + 1] += x.y; v[4 * j + 2] += x.z; v[4 * j + 3] += x.w;
    }
  }
  if (c.p.out16 || c.has_pre) {
    if (c.p.out16) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(c.row64a + poff + ((j ^ c.x64) << 4)),
                     "r"(pack_bf16(v[8 * j], v[8 * j + 1])), "r"(pack_bf16(v[8 * j + 2], v[8 * j + 3])),
                     "r"(pack_bf16(v[8 * j + 4], v[8 * j + 5])), "r"(pack_bf16(v[8 * j + 6], v[8 * j + 7]))
                     : "memory");
    }
    if (trc) GTRACE(2, 34);
    fence_proxy_async();
    __syncwarp();
    if (trc) GTRACE(2, 35);
    // the two stores are issued (and their bulk groups tracked) by two different lanes: one lane takes ~200 cycles
    // per UTMASTG + commit, and every lane executes the wait_group above before the staging buffer is reused
    if (c.lane == 0 && c.p.out16) {
      tma_store_2d(c.tmO16, c.sbuf + poff, col, row0);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    if (c.lane == 1 && c.has_pre) {
      tma_store_2d_hint(c.tmPre, c.sbuf + 2048, col, row0, l2_policy_evict_first());
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    if (trc) GTRACE(2, 36);
  }
  if (c.p.out32) {
    if (c.p.out16 || c.has_pre) {  // the fp32 image needs the whole buffer: wait for the bf16 stores to drain it
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncwarp();
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(c.row128 + ((j ^ c.x128) << 4)), "f"(v[4 * j]),
                   "f"(v[4 * j + 1]), "f"(v[4 * j + 2]), "f"(v[4 * j + 3])
                   : "memory");
    fence_proxy_async();
    __syncwarp();
    if (c.lane == 0) {
      tma_store_2d(c.tmO32, c.sbuf, col, row0);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
}

// One split-K unit in the epilogue warps (accumulator complete): raw fp32 partial -> workspace, publish, wait for the
// sibling slices, then this CTA's share of the ordered reduction + epilogue, and re-arm the counters. Only the
// kSplitK instantiations of the kernel contain it (its registers would weigh on the plain epilogue loop otherwise).
template <int EPI>
__device__ __forceinline__ void splitk_unit(const EpiCtx& c, uint32_t taddr, int w, int half, int q, int ew, int m0,
                                         int n0, int rem_idx, int slice, uint64_t* tempty, uint32_t& pp) {
  const GemmParams& p = c.p;
  const int lane = c.lane;
    // partial layout: [unit][row group rq 0..3][chunk cc 0..7][j 0..7][lane] float4 — one warp store/load
    // instruction covers 512 contiguous bytes (lane = row inside the 32-row group, j = float4 of its 32 columns)
    float4* wsu = reinterpret_cast<float4*>(p.ws) + (size_t)(rem_idx * p.splitk + slice) * (BM * 256 / 4);
    for (int c = half * 32; c < w; c += 64) {
      uint32_t r[32];
      tmem_ld32(taddr + (uint32_t)c, r);
      tc_wait_ld();
      float4* dst = wsu + ((q * 8 + (c >> 5)) * 8) * 32 + lane;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        __stcg(dst + j * 32, make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                         __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3])));
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(tempty);
    if (ew == 0 && lane == 0) GTRACE(2, 40);
    // publish: every epilogue thread's partial is visible device-wide before the unit is counted
    __threadfence();
    asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
    if (ew == 0 && lane == 0) {
      GTRACE(2, 41);
      atomicAdd(p.ws_cnt + rem_idx, 1);
      
    }
    // The epilogue warps have nothing else to do until the next (full) tile's accumulator is complete, a whole
    // main loop away: the ordered reduction of this CTA's share runs now, hidden behind that main loop. All
    // sibling units were started at kernel launch too, so the wait is short.
    if (ew == 0 && lane == 0) {
      int seen;
      do {
        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(p.ws_cnt + rem_idx) : "memory");
      } while (seen < p.splitk);
      GTRACE(2, 42);
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
    // Fix-up. The 128 x 256 tile is 32 pieces of 32 rows x 32 columns; the CTA of slice s owns pieces s, s + S, ...
    // (S = splitk is a power of two). S <= 4: a warp sums all S slices of a piece. S >= 8: the 8/(32/S) warps
    // that share a piece sum 4 consecutive slices each, park their sub-sums in their staging buffers and the
    // first of them adds those up in warp order. Either way the association is fixed: bit-reproducible.
    const int S = p.splitk;
    const int npieces = 32 / S;                          // pieces owned by this CTA
    const int wpp = npieces >= kEpiWarps ? 1 : kEpiWarps / npieces;  // warps per piece
    const int nsl = S / wpp;                             // slices summed by one warp (<= 4)
    const float4* ws4 = reinterpret_cast<const float4*>(p.ws) + (size_t)(rem_idx * S) * (BM * 256 / 4);
    for (int j0 = 0; j0 < npieces; j0 += kEpiWarps / wpp) {
      const int jp = j0 + ew / wpp;                      // this warp's piece number in this round
      const int sub = ew % wpp;
      const bool active = jp < npieces;
      const int pc = slice + jp * S;
      const int rq = pc >> 3, cc = pc & 7;
      const int prow = m0 + rq * 32 + lane, col = n0 + cc * 32;
      const bool inside = active && col < p.N && m0 + rq * 32 < p.M;
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = 0.f;
      if (inside) {
        const float4* src = ws4 + (size_t)(sub * nsl) * (BM * 256 / 4) + ((rq * 8 + cc) * 8) * 32 + lane;
        for (int sl = 0; sl < nsl; ++sl, src += BM * 256 / 4) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 x = __ldcg(src + j * 32);
            v[4 * j] += x.x; v[4 * j + 1] += x.y; v[4 * j + 2] += x.z; v[4 * j + 3] += x.w;
          }
        }
      }
      if (ew == 0 && lane == 0) GTRACE(2, 43);
      if (wpp > 1) {
        if (sub != 0 && inside) {
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // earlier TMA stores left this buffer
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j)
            asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(c.row128 + ((j ^ c.x128) << 4)), "f"(v[4 * j]),
                         "f"(v[4 * j + 1]), "f"(v[4 * j + 2]), "f"(v[4 * j + 3])
                         : "memory");
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
        if (sub == 0 && inside) {
          for (int o = 1; o < wpp; ++o) {
            const uint32_t orow = c.row128 + o * kStageBufBytes;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float4 x;
              asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                           : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w)
                           : "r"(orow + ((j ^ c.x128) << 4)));
              v[4 * j] += x.x; v[4 * j + 1] += x.y; v[4 * j + 2] += x.z; v[4 * j + 3] += x.w;
            }
          }
        }
      }
      if (ew == 0 && lane == 0) GTRACE(2, 44);
      if (sub == 0 && inside) {
        const bool ok = prow < p.M;
        float4 rv[EPI == 3 ? 8 : 1];
        uint4 av[EPI == 2 ? 4 : 1];
        if constexpr (EPI == 3) load_res_chunk(p, m0 + rq * 32, col, lane, rv);
        if constexpr (EPI == 2) load_aux_chunk(p, m0 + rq * 32, col, lane, av);
        finish_chunk<EPI>(c, v, rv, av, m0 + rq * 32, col, ok, pp++);
      }
      if (ew == 0 && lane == 0) GTRACE(2, 45);
      if (wpp > 1 && j0 + kEpiWarps / wpp < npieces)  // buffers are reused by the next round
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
    }
    // the last CTA to finish its share re-arms the counters for the next launch
    asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
    if (ew == 0 && lane == 0) {
      const int rem = (p.total_tiles - p.full_tiles) / p.splitk;
      if (atomicAdd(p.ws_cnt + rem + rem_idx, 1) == p.splitk - 1) {
        p.ws_cnt[rem_idx] = 0;
        p.ws_cnt[rem + rem_idx] = 0;
      }
      GTRACE(2, 46);
    }
}

// kMN = false: A[M,K], B[N,K] (K contiguous; "TN").  kMN = true: A given as [K,M], B as [K,N] (M / N contiguous:
// MN-major UMMA operands) — the wgrad form dW = dY^T X straight from the row-major activations, no transposes.
// EPI selects the epilogue at compile time (keeps each variant's register footprint small):
//   0 plain (+bias)   1 bias + QuickGELU (+pre-activation store)   2 * QuickGELU'(aux)   3 (+bias) + fp32 residual
// kPair: launched as clusters of 2 CTAs; tcgen05 cta_group::2 — one UMMA of M = 256 spans the pair, each CTA keeps its
// own 128 A rows and HALF of the B rows in smem (the tensor cores of the two SMs exchange the halves), which halves
// the shared-memory traffic per SM; the leader CTA issues all MMAs, both CTAs run producers and epilogues.
template <int BN, int kStages, bool kMN, int EPI, bool kPair, bool kSplitK>
__global__ void __launch_bounds__(kThreads, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmO32, const __grid_constant__ CUtensorMap tmO16,
                    const __grid_constant__ CUtensorMap tmPre, const GemmParams p) {
  constexpr uint32_t kABytes = BM * BK * 2;
  constexpr uint32_t kBBytes = (kPair ? BN / 2 : BN) * BK * 2;  // a pair CTA keeps only its half of every B tile
  constexpr uint32_t kStageBytes = kABytes + kBBytes;
  constexpr uint32_t kSubBytes = kABytes + 64 * BK * 2;  // one k-block of a 64-wide tile: A + 8 KB of B (1024-B multiple)
  constexpr bool kDoubleNarrow = !kMN && !kPair && kStageBytes >= 2 * kSubBytes;  // the 256-wide instantiations

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_out = smem + kStages * kStageBytes;  // kEpiWarps x 4 KB output staging (1024-B aligned)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stage_out + kEpiWarps * kStageBufBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tfull_bar = empty_bar + kStages;   // [2] accumulator ready
  uint64_t* tempty_bar = tfull_bar + 2;        // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // provably warp-uniform
  const int lane = threadIdx.x & 31;
  const int num_kb = (p.K + BK - 1) / BK;
  // pair mode: the two CTAs of a cluster work on vertically adjacent M tiles of the same N range
  const int crank = kPair ? (int)cluster_ctarank() : 0;
  const int tile0 = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int tstep = kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  // This CTA's tiles are tile0, tile0 + tstep, ... (n_mine of them). With a split-K tail the last one may be a
  // split-K unit: it is then processed FIRST, so that writing / publishing its partial and the ordered reduction
  // are hidden behind the main loop of the following full tile.
  const int n_mine = p.total_tiles > tile0 ? (p.total_tiles - tile0 + tstep - 1) / tstep : 0;
  const bool unit_first = kSplitK && p.splitk > 0 && n_mine > 0 && tile0 + (n_mine - 1) * tstep >= p.full_tiles;
  auto tile_of = [&](int it) {
    return unit_first ? (it == 0 ? tile0 + (n_mine - 1) * tstep : tile0 + (it - 1) * tstep) : tile0 + it * tstep;
  };

  if (warp == 0 && lane == 0) {
    GTRACE(0, 0);
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.out32) tma_prefetch_desc(&tmO32);
    if (p.out16) tma_prefetch_desc(&tmO16);
    if (p.outpre) tma_prefetch_desc(&tmPre);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], kPair ? 2 * kEpiWarps : kEpiWarps);  // one arrive per epilogue warp (of both CTAs)
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if (kPair) {
      tmem_alloc_2sm(tmem_slot, kTmemCols);
      tmem_relinquish_2sm();
    } else {
      tmem_alloc(tmem_slot, kTmemCols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (kPair) cluster_sync_all();  // peer barriers are initialised before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // everything above overlapped the tail of the previous kernel (PDL); from here on global memory is touched
  pdl_trigger();
  pdl_wait();

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int trc = 2;
      GTRACE(0, 1);
      if (p.trace) p.trace[((size_t)blockIdx.x * 3) * 64 + 63] = (long long)globaltimer_ns();
      for (int it = 0; it < n_mine; ++it) {
        const int t = tile_of(it);
        int m0, n0, w;
        decode_tile(p, t, crank, m0, n0, w);
        const uint32_t tx = kABytes + (uint32_t)w * BK * 2;
        int rem_idx = 0, kb0 = 0, kb1 = num_kb;
        if (kSplitK) decode_slice(p, t, num_kb, rem_idx, kb0, kb1);
        // Narrow tail tiles (w <= 64) are bound by the TMA round trip, not by the tensor core: a 24 KB k-block per
        // 48 KB stage leaves only 4 x 24 KB in flight (~480 cycles per k-block). They carry TWO k-blocks per stage.
        if (kDoubleNarrow && w == 64 && !(kSplitK && p.splitk > 0 && t >= p.full_tiles)) {
          for (int kb = kb0; kb < kb1; kb += 2) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            const int nsub = min(2, kb1 - kb);
            uint8_t* sa = smem + stage * kStageBytes;
            mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)nsub * tx);
            for (int sub = 0; sub < nsub; ++sub) {
              tma_load_2d(&tmA, &full_bar[stage], sa + sub * kSubBytes, (kb + sub) * BK, m0);
              tma_load_2d(&tmB, &full_bar[stage], sa + sub * kSubBytes + kABytes, (kb + sub) * BK, n0);
            }
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
          continue;
        }
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (kb == kb0 || kb == kb1 - 1) { GTRACE(0, trc); ++trc; }
          uint8_t* sa = smem + stage * kStageBytes;
          uint8_t* sb = sa + kABytes;
          if (!kMN && kPair) {
            // pair mode: this CTA fetches its own A rows and its half of the B rows; all bytes of both CTAs are
            // accounted on the LEADER's full barrier (the leader alone issues the MMAs)
            if (crank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2u * kABytes + (uint32_t)w * BK * 2);
            const uint32_t lead_bar = map_to_cta(smem_u32(&full_bar[stage]), 0);
            tma_load_2d_2sm(&tmA, lead_bar, sa, kb * BK, m0);
            const int hw = w >> 1;
            for (int j = 0; j < hw; j += 32)
              tma_load_2d_2sm(&tmB, lead_bar, sb + j * (BK * 2), kb * BK, n0 + crank * hw + j);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
            continue;
          }
          mbar_arrive_expect_tx(&full_bar[stage], tx);
          if (!kMN) {
            tma_load_2d(&tmA, &full_bar[stage], sa, kb * BK, m0);
            for (int j = 0; j < w; j += BOXN)
              tma_load_2d(&tmB, &full_bar[stage], sb + j * (BK * 2), kb * BK, n0 + j);
          } else {
            // boxes of 64 (M or N, contiguous) x 64 K-rows = 8 KB, one per 64-wide MN group
            for (int j = 0; j < BM; j += 64)
              tma_load_2d(&tmA, &full_bar[stage], sa + j * 128, m0 + j, kb * BK);
            for (int j = 0; j < w; j += 64)
              tma_load_2d(&tmB, &full_bar[stage], sb + j * 128, n0 + j, kb * BK);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    if (!(kPair && crank != 0) && elect_one()) {  // (elect_one: no per-MMA elect / branch loop around tcgen05.mma)
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < n_mine; ++it) {
        const int t = tile_of(it);
        int m0, n0, w;
        decode_tile(p, t, crank, m0, n0, w);
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);  // epilogue drained this accumulator
        tc_fence_after();
        GTRACE(1, 3 * it);
        const uint32_t idesc = umma_idesc_bf16(kPair ? 2 * BM : BM, w, kMN ? 1 : 0, kMN ? 1 : 0);
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * 256u;
        int rem_idx = 0, kb0 = 0, kb1 = num_kb;
        if (kSplitK) decode_slice(p, t, num_kb, rem_idx, kb0, kb1);
        if (kDoubleNarrow && w == 64 && !(kSplitK && p.splitk > 0 && t >= p.full_tiles)) {
          for (int kb = kb0; kb < kb1; kb += 2) {  // two k-blocks per stage (see the producer)
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            if (kb == kb0) GTRACE(1, 3 * it + 1);
            const int nsub = min(2, kb1 - kb);
            for (int sub = 0; sub < nsub; ++sub) {
              const uint32_t sa = smem_u32(smem + stage * kStageBytes + sub * kSubBytes);
              const uint64_t adesc = umma_desc_k_sw128(sa), bdesc = umma_desc_k_sw128(sa + kABytes);
#pragma unroll
              for (int k = 0; k < BK / UK; ++k)
                umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb > kb0 || sub > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit(&empty_bar[stage]);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
          umma_commit(&tfull_bar[acc]);
          GTRACE(1, 3 * it + 2);
          continue;
        }
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (kb == kb0) GTRACE(1, 3 * it + 1);
          const uint32_t sa = smem_u32(smem + stage * kStageBytes);
          const uint64_t adesc = kMN ? umma_desc_mn_sw128(sa, 8192) : umma_desc_k_sw128(sa);
          const uint64_t bdesc = kMN ? umma_desc_mn_sw128(sa + kABytes, 8192) : umma_desc_k_sw128(sa + kABytes);
          // per UMMA_K = 16: K-major advances 32 B inside the swizzle row (+2 in 16-byte units); MN-major
          // advances 16 K-rows of 128 B (+128 in 16-byte units)
          constexpr uint64_t kStep = kMN ? 128 : 2;
          if (kPair) {
#pragma unroll
            for (int k = 0; k < BK / UK; ++k)
              umma_bf16_2sm(d_tmem, adesc + kStep * k, bdesc + kStep * k, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            umma_commit_2sm_mc(&empty_bar[stage], (uint16_t)3);
          } else {
#pragma unroll
            for (int k = 0; k < BK / UK; ++k)
              umma_bf16(d_tmem, adesc + kStep * k, bdesc + kStep * k, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            umma_commit(&empty_bar[stage]);  // smem slot free once these MMAs retire
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (kPair) umma_commit_2sm_mc(&tfull_bar[acc], (uint16_t)3);  // both CTAs' epilogues wake
        else umma_commit(&tfull_bar[acc]);                                    // accumulator complete
        GTRACE(1, 3 * it + 2);
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ================================ epilogue ====================================
    const int q = warp & 3;          // TMEM lane quarter this warp may access
    const int half = (warp - 4) >> 2;  // which interleaved set of 32-column chunks
    uint8_t* sbuf = stage_out + (warp - 4) * kStageBufBytes;
    const uint32_t sbuf_u32 = smem_u32(sbuf);
    // swizzled 16-byte chunk offsets of this lane's row inside the staging buffer
    const uint32_t row128 = sbuf_u32 + lane * 128, x128 = lane & 7;            // 128-B rows, SWIZZLE_128B
    const uint32_t row64a = sbuf_u32 + lane * 64, x64 = (lane >> 1) & 3;       // 64-B rows,  SWIZZLE_64B
    const uint32_t row64b = row64a + 2048;                                     // second bf16 output
    const bool has_pre = EPI == 1 && p.outpre != nullptr;
    const EpiCtx ctx{p, &tmO32, &tmO16, &tmPre, sbuf, row128, x128, row64a, x64, row64b, has_pre, lane};

    // Epilogue operands that do not depend on the accumulator (fp32 residual / QuickGELU aux) are fetched one
    // chunk ahead — across tile boundaries too — so their DRAM latency hides behind the previous chunk.
    float4 rv_nx[EPI == 3 ? 8 : 1];
    uint4 av_nx[EPI == 2 ? 4 : 1];
    auto fetch_ops = [&](int t, int c) {
      if (t >= p.total_tiles) return;
      if (kSplitK && p.splitk && t >= p.full_tiles) return;  // split-K units fetch their operands in the fix-up pass
      int m0, n0, w;
      decode_tile(p, t, crank, m0, n0, w);
      const int col = n0 + c;
      if (c >= w || col >= p.N) return;
      if constexpr (EPI == 3) load_res_chunk(p, m0 + q * 32, col, lane, rv_nx);
      if constexpr (EPI == 2) load_aux_chunk(p, m0 + q * 32, col, lane, av_nx);
    };
    if (n_mine > 0) fetch_ops(tile_of(0), half * 32);
    uint32_t nchunk = 0;  // chunks finished by this warp (ping-pong parity of the staging buffer)
    for (int it = 0; it < n_mine; ++it) {
      const int t = tile_of(it);
      int m0, n0, w;
      decode_tile(p, t, crank, m0, n0, w);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int row = m0 + q * 32 + lane;
      const bool row_ok = row < p.M;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * 256u;
      int rem_idx = 0, kb0 = 0, kb1 = num_kb;
      const int slice = kSplitK ? decode_slice(p, t, num_kb, rem_idx, kb0, kb1) : -1;
      if (kSplitK && slice >= 0) {
        // ---------------- split-K unit: raw fp32 partial -> workspace, then a share of the ordered reduction
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
        if (warp == 4 && lane == 0) GTRACE(2, 2 * it);
        splitk_unit<EPI>(ctx, taddr, w, half, q, warp - 4, m0, n0, rem_idx, slice, &tempty_bar[acc], nchunk);
        if (EPI >= 2 && it + 1 < n_mine) fetch_ops(tile_of(it + 1), half * 32);
        continue;
      }
      bool waited = false;
      for (int c = half * 32; c < w; c += 64) {
        const int col = n0 + c;
        const bool ok = row_ok && col < p.N;
        float4 rv[EPI == 3 ? 8 : 1];
        uint4 av[EPI == 2 ? 4 : 1];
        if (EPI == 3) {
#pragma unroll
          for (int j = 0; j < 8; ++j) rv[j] = rv_nx[j];
        }
        if (EPI == 2) {
#pragma unroll
          for (int j = 0; j < 4; ++j) av[j] = av_nx[j];
        }
        if (EPI >= 2) {  // next chunk of this tile, else the first chunk of this CTA's next tile
          if (c + 64 < w) fetch_ops(t, c + 64);
          else if (it + 1 < n_mine) fetch_ops(tile_of(it + 1), half * 32);
        }
        if (!waited) {
          mbar_wait(&tfull_bar[acc], acc_phase);
          tc_fence_after();
          waited = true;
          if (warp == 4 && lane == 0) GTRACE(2, 2 * it);
        }
        if (it == 1 && warp == 4 && lane == 0) GTRACE(2, 48 + 3 * (c >> 6));
        float4 bv[8];
        if (EPI != 2 && p.bias && col < p.N) {
#pragma unroll
          for (int j = 0; j < 8; ++j) bv[j] = __ldg(reinterpret_cast<const float4*>(p.bias + col) + j);
        }
        uint32_t r[32];
        tmem_ld32(taddr + (uint32_t)c, r);
        tc_wait_ld();
        if (it == 1 && warp == 4 && lane == 0) GTRACE(2, 49 + 3 * (c >> 6));
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        finish_chunk<EPI>(ctx, v, rv, av, m0 + q * 32, col, ok, nchunk++, bv);
        if (it == 1 && warp == 4 && lane == 0) GTRACE(2, 50 + 3 * (c >> 6));
      }
      if (!waited) {  // narrow tile: this warp had no chunk, but it still takes part in the hand-shake
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (kPair && crank != 0) mbar_arrive_remote(map_to_cta(smem_u32(&tempty_bar[acc]), 0));
        else mbar_arrive(&tempty_bar[acc]);
        if (warp == 4) GTRACE(2, 2 * it + 1);
      }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // all output stores complete before exit
    if (warp == 4 && lane == 0) GTRACE(2, 62);
  }

  tc_fence_before();
  __syncthreads();
  if (kPair) cluster_sync_all();  // the peer may still arrive on this CTA's barriers until it is done too
  if (warp == 2) {
    tc_fence_after();
    if (kPair) tmem_dealloc_2sm(tmem_base, kTmemCols);
    else tmem_dealloc(tmem_base, kTmemCols);
  }
}

template <int BN, int kStages, bool kPair = false>
constexpr size_t gemm_smem_bytes() {
  return (size_t)kStages * (BM * BK * 2 + (kPair ? BN / 2 : BN) * BK * 2) + kEpiWarps * kStageBufBytes +
         (2 * kStages + 4) * 8 + 16 + 1024;
}

int g_num_sms = 0;
long long* g_gemm_trace = nullptr;

}  // namespace

// ----------------------------------------------------------------------------- host: tensor maps
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encode = nullptr;

static int load_encode() {
  if (g_encode) return MFK_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || fn == nullptr || qres != cudaDriverEntryPointSuccess) return MFK_EDRIVER;
  g_encode = reinterpret_cast<PFN_encodeTiled>(fn);
  return MFK_OK;
}

int mfk_make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                          uint32_t box_rows, uint32_t box_cols) {
  int rc = load_encode();
  if (rc != MFK_OK) return rc;
  if (!mfk_aligned16(base) || (ld_elems * 2) % 16 != 0) return MFK_EALIGN;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld_elems * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MFK_OK : MFK_EDRIVER;
}

int mfk_make_tmap_2d(CUtensorMap* out, const void* base, int elem_bytes, uint64_t rows, uint64_t cols,
                     uint64_t ld_elems, uint32_t box_rows, uint32_t box_cols, int swizzle_bytes) {
  int rc = load_encode();
  if (rc != MFK_OK) return rc;
  if (!mfk_aligned16(base) || (ld_elems * elem_bytes) % 16 != 0) return MFK_EALIGN;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld_elems * (uint64_t)elem_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(out, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                        const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MFK_OK : MFK_EDRIVER;
}

int mfk_make_tmap_bf16_3d(CUtensorMap* out, const void* base, uint64_t seqs, uint64_t rows, uint64_t cols,
                          uint64_t ld_elems, uint32_t box_rows, uint32_t box_cols) {
  int rc = load_encode();
  if (rc != MFK_OK) return rc;
  if (!mfk_aligned16(base) || (ld_elems * 2) % 16 != 0) return MFK_EALIGN;
  cuuint64_t gdim[3] = {cols, rows, seqs};
  cuuint64_t gstride[2] = {ld_elems * 2ull, rows * ld_elems * 2ull};
  cuuint32_t box[3] = {box_cols, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MFK_OK : MFK_EDRIVER;
}

static int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

template <int BN, int kStages, bool kMN, int EPI, bool kPair = false, bool kSplitK = false>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO32,
                       const CUtensorMap& tmO16, const CUtensorMap& tmPre, const GemmParams& p, int grid,
                       cudaStream_t st) {
  static bool configured = false;
  constexpr size_t smem = gemm_smem_bytes<BN, kStages, kPair>();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_tn_kernel<BN, kStages, kMN, EPI, kPair, kSplitK>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  cudaError_t le = launch_pdl_cluster(gemm_bf16_tn_kernel<BN, kStages, kMN, EPI, kPair, kSplitK>, dim3(grid), dim3(kThreads), smem,
                                      st, kPair ? 2 : 1, tmA, tmB, tmO32, tmO16, tmPre, p);
  if (le != cudaSuccess) return (int)le;
  return MFK_OK;
}

extern "C" int mfk_gemm_bf16(const void* A, long long lda, const void* B, long long ldb, int M, int N, int K,
                             const float* bias, int act, const void* aux, long long ldaux, const float* residual,
                             long long ldres, float* out_f32, long long ld32, void* out_bf16, long long ld16,
                             void* out_pre_bf16, long long ldpre, int tile_n, void* splitk_ws,
                             long long splitk_ws_bytes, void* stream) {
  if (!A || !B || M <= 0 || N <= 0 || K <= 0) return MFK_EARG;
  if (splitk_ws && (!mfk_aligned16(splitk_ws) || splitk_ws_bytes < 0)) return MFK_EALIGN;
  if (N % 32 != 0 || lda % 8 != 0 || ldb % 8 != 0 || lda < K || ldb < K) return MFK_ESHAPE;
  if (act < 0 || act > 2 || (act == 2 && !aux)) return MFK_EARG;
  if (!out_f32 && !out_bf16) return MFK_EARG;
  if ((out_f32 && (ld32 % 4 || !mfk_aligned16(out_f32))) || (out_bf16 && (ld16 % 8 || !mfk_aligned16(out_bf16))) ||
      (out_pre_bf16 && (ldpre % 8 || !mfk_aligned16(out_pre_bf16))) || (aux && (ldaux % 8 || !mfk_aligned16(aux))) ||
      (residual && (ldres % 4 || !mfk_aligned16(residual))) || (bias && !mfk_aligned16(bias)))
    return MFK_EALIGN;

  const int sms = num_sms();
  const int m_tiles = (M + BM - 1) / BM;
  // full tile width: 256 unless N is small or the caller forces 128
  int bn = (tile_n == 128) ? 128 : 256;  // tile_n == 2 selects the CTA-pair (cta_group::2) kernel with 256-wide tiles
  if (tile_n == 0 && N < 256) bn = 128;
  GemmParams p;
  p.M = M; p.N = N; p.K = K;
  p.bn = bn;
  p.n_big = (N + bn - 1) / bn;
  // CTA-pair kernel (tcgen05 cta_group::2, UMMA M = 256) on request. It halves per-SM shared-memory traffic, which
  // pays for long K loops (8192^3: +3 %); the MaPLe shapes (K <= 3072, <= 4 tiles per SM) are bound by per-launch
  // fixed costs instead and are as fast with single CTAs, so `auto` keeps those.
  p.cluster = ((tile_n == 2 || tile_n == 3) && m_tiles >= 2) ? 2 : 0;
  const int units = p.cluster ? sms / 2 : sms;                      // schedulable CTAs or CTA pairs
  const int big = (p.cluster ? (m_tiles + 1) / 2 : m_tiles) * p.n_big;
  p.full_tiles = (big / units) * units;
  const int rem = big - p.full_tiles;
  p.split = 1;
  if (rem > 0) {
    const int max_split = bn / 64;
    while (p.split * 2 <= max_split && rem * p.split * 2 <= units) p.split *= 2;
  }
  p.total_tiles = p.full_tiles + rem * p.split;
  // Split-K tail (needs a caller workspace): when the remainder tiles are few and K is long, N-narrowed tiles are a
  // poor tail — a 64-wide tile still streams the whole 128 x K operand A through ONE SM (L2->SM bandwidth bound,
  // ~480 cycles per k-block), e.g. 2 of 150 tiles cost 2/3 of a wave at M=6368, N=768, K=3072. Instead every
  // remainder tile is cut along K into S slices on S different SMs; slices leave fp32 partials in the workspace
  // and each of the S CTAs then reduces (fixed slice order) and finishes 1/S of the tile.
  p.splitk = 0; p.kb_per = 0; p.ws = nullptr; p.ws_cnt = nullptr;
  const int num_kb = (K + BK - 1) / BK;
  if (splitk_ws && rem > 0 && !p.cluster && bn == 256 && tile_n == 0) {
    const long long avail = (splitk_ws_bytes - kSplitKCounterBytes) / ((long long)BM * 256 * 4);
    int S = 32;  // power of two: <= one unit per SM, >= 2 k-blocks per slice, no empty slice, fits the workspace
    while (S >= 2 && (rem * S > sms || S * 2 > num_kb || (long long)rem * S > avail ||
                      (S - 1) * ((num_kb + S - 1) / S) >= num_kb))
      S >>= 1;
    if (S >= 2 && 2 * rem <= kSplitKCounterBytes / 4) {
      const int kb_per = (num_kb + S - 1) / S;
      const int w_narrow = bn / p.split;
      const long long narrow_cost = (long long)num_kb * (w_narrow <= 64 ? 480 : w_narrow <= 128 ? 580 : 700);
      // exposed cost of a unit = its few k-blocks (it runs first; publishing and the reduction hide behind the next
      // main loop). Measured on the training step: pays for long K (>= 24 k-blocks: the N=768 dgrads and c_proj of
      // ViT-B/16); on short K the epilogue warps have no slack to hide the reduction in. MFK_SPLITK_MIN_KB tunes it.
      const long long splitk_cost = 1500 + 700LL * kb_per;
      static const int min_kb = getenv("MFK_SPLITK_MIN_KB") ? atoi(getenv("MFK_SPLITK_MIN_KB")) : 24;
      if (S >= 2 && splitk_cost < narrow_cost && num_kb >= min_kb) {
        p.splitk = S; p.kb_per = kb_per; p.split = 1;
        p.total_tiles = p.full_tiles + rem * S;
        p.ws_cnt = static_cast<int*>(splitk_ws);
        p.ws = reinterpret_cast<float*>(static_cast<char*>(splitk_ws) + kSplitKCounterBytes);
      }
    }
  }
  p.bias = bias; p.act = act;
  p.aux = static_cast<const bf16*>(aux); p.ldaux = ldaux;
  p.res = residual; p.ldres = ldres;
  p.out32 = out_f32; p.ld32 = ld32;
  p.out16 = static_cast<bf16*>(out_bf16); p.ld16 = ld16;
  p.outpre = static_cast<bf16*>(out_pre_bf16); p.ldpre = ldpre;
  p.trace = g_gemm_trace;

  CUtensorMap tmA, tmB;
  int rc = mfk_make_tmap_bf16_2d(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, BM, BK);
  if (rc != MFK_OK) return rc;
  rc = mfk_make_tmap_bf16_2d(&tmB, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, p.cluster ? 32 : BOXN, BK);
  if (rc != MFK_OK) return rc;
  // output tensor maps: 32-row x 32-column boxes (one epilogue warp's chunk); absent outputs get a dummy map
  CUtensorMap tmO32 = tmA, tmO16 = tmA, tmPre = tmA;
  if (out_f32 && (rc = mfk_make_tmap_2d(&tmO32, out_f32, 4, (uint64_t)M, (uint64_t)N, (uint64_t)ld32, 32, 32, 128)))
    return rc;
  if (out_bf16 && (rc = mfk_make_tmap_2d(&tmO16, out_bf16, 2, (uint64_t)M, (uint64_t)N, (uint64_t)ld16, 32, 32, 64)))
    return rc;
  if (out_pre_bf16 &&
      (rc = mfk_make_tmap_2d(&tmPre, out_pre_bf16, 2, (uint64_t)M, (uint64_t)N, (uint64_t)ldpre, 32, 32, 64)))
    return rc;
  const int grid = p.cluster ? 2 * (p.total_tiles < units ? p.total_tiles : units)
                             : (p.total_tiles < sms ? p.total_tiles : sms);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int epi = act == 1 ? 1 : act == 2 ? 2 : residual ? 3 : 0;
#define MFK_GEMM_DISPATCH(BN_, ST_)                                                                           \
  switch (epi) {                                                                                              \
    case 0: return launch_gemm<BN_, ST_, false, 0>(tmA, tmB, tmO32, tmO16, tmPre, p, grid, st);               \
    case 1: return launch_gemm<BN_, ST_, false, 1>(tmA, tmB, tmO32, tmO16, tmPre, p, grid, st);               \
    case 2: return launch_gemm<BN_, ST_, false, 2>(tmA, tmB, tmO32, tmO16, tmPre, p, grid, st);               \
    default: return launch_gemm<BN_, ST_, false, 3>(tmA, tmB, tmO32, tmO16, tmPre, p, grid, st);              \
  }
  if (p.cluster == 2 && tile_n == 3) {  // experiment: 6 stages of 32 KB (deeper TMA queue)
    switch (epi) {
      case 0: return launch_gemm<256, 6, false, 0, true>(tmA, tmB, tmO32, tmO16, tmPre, p, grid, st);
      case 1: return launch_gemm<256, 6, false, 1, true>(tmA, tmB, tmO32, tmO16, tmPre, p, grid, st);
      case 2: return launch_gemm<256, 6, false, 2, true>(tmA, tmB, tmO32, tmO16, tmPre, p, grid, st);
      default: return launch_gemm<256, 6, false, 3, true>(tmA, tmB, tmO32, tmO16, tmPre, p, grid, st);
    }
  }
  if (p.cluster == 2) {
    switch (epi) {
      case 0: return launch_gemm<256, 4, false, 0, true>(tmA, tmB, tmO32, tmO16, tmPre, p, grid, st);
      case 1: return launch_gemm<256, 4, false, 1, true>(tmA, tmB, tmO32, tmO16, tmPre, p, grid, st);
      case 2: return launch_gemm<256, 4, false, 2, true>(tmA, tmB, tmO32, tmO16, tmPre, p, grid, st);
      default: return launch_gemm<256, 4, false, 3, true>(tmA, tmB, tmO32, tmO16, tmPre, p, grid, st);
    }
  }
  if (p.splitk) {  // bn == 256, single CTAs
    switch (epi) {
      case 0: return launch_gemm<256, 4, false, 0, false, true>(tmA, tmB, tmO32, tmO16, tmPre, p, grid, st);
      case 1: return launch_gemm<256, 4, false, 1, false, true>(tmA, tmB, tmO32, tmO16, tmPre, p, grid, st);
      case 2: return launch_gemm<256, 4, false, 2, false, true>(tmA, tmB, tmO32, tmO16, tmPre, p, grid, st);
      default: return launch_gemm<256, 4, false, 3, false, true>(tmA, tmB, tmO32, tmO16, tmPre, p, grid, st);
    }
  }
  if (bn == 256) { MFK_GEMM_DISPATCH(256, 4) }
  MFK_GEMM_DISPATCH(128, 6)
#undef MFK_GEMM_DISPATCH
}

// debug hook: device buffer of grid x 3 x 64 int64 receiving clock64 stamps of every CTA (NULL disables)
extern "C" int mfk_debug_set_gemm_trace(void* dev_buf) {
  g_gemm_trace = static_cast<long long*>(dev_buf);
  return MFK_OK;
}

// out[M,N] (fp32) = At^T * Bt with At[K,M], Bt[K,N] bf16 row-major (leading dimensions lda, ldb >= M, N).
extern "C" int mfk_gemm_bf16_at_b(const void* At, long long lda, const void* Bt, long long ldb, int M, int N, int K,
                                  float* out_f32, long long ld32, void* stream) {
  if (!At || !Bt || !out_f32 || M <= 0 || N <= 0 || K <= 0) return MFK_EARG;
  if (N % 32 != 0 || lda % 8 != 0 || ldb % 8 != 0 || lda < M || ldb < N) return MFK_ESHAPE;
  if (ld32 % 4 || !mfk_aligned16(out_f32)) return MFK_EALIGN;
  const int sms = num_sms();
  const int m_tiles = (M + BM - 1) / BM;
  GemmParams p{};
  p.M = M; p.N = N; p.K = K;
  p.bn = 256;
  p.n_big = (N + 255) / 256;
  const int big = m_tiles * p.n_big;
  p.full_tiles = (big / sms) * sms;
  const int rem = big - p.full_tiles;
  p.split = 1;
  if (rem > 0)
    while (p.split * 2 <= 4 && rem * p.split * 2 <= sms) p.split *= 2;
  p.total_tiles = p.full_tiles + rem * p.split;
  p.out32 = out_f32; p.ld32 = ld32;
  p.cluster = 0;
  CUtensorMap tmA, tmB, tmO32;
  int rc = mfk_make_tmap_2d(&tmA, At, 2, (uint64_t)K, (uint64_t)M, (uint64_t)lda, 64, 64, 128);
  if (rc != MFK_OK) return rc;
  if ((rc = mfk_make_tmap_2d(&tmB, Bt, 2, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, 64, 64, 128)) != MFK_OK) return rc;
  if ((rc = mfk_make_tmap_2d(&tmO32, out_f32, 4, (uint64_t)M, (uint64_t)N, (uint64_t)ld32, 32, 32, 128)) != MFK_OK)
    return rc;
  const int grid = p.total_tiles < sms ? p.total_tiles : sms;
  return launch_gemm<256, 4, true, 0>(tmA, tmB, tmO32, tmA, tmA, p, grid, static_cast<cudaStream_t>(stream));
}
