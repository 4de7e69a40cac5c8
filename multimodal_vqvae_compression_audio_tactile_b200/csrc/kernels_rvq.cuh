// ResidualVQEMA.forward (Evaluation/dac_vcpwq_proposed6_latency.py:421-435) on tcgen05: ALL books in ONE launch.
//
// A CTA owns 128 token rows for the whole residual loop.  Per book:
//   scores S[128 x K] = R . E_b^T  as a bf16x3 GEMM (R = current residual, hi/lo bf16 planes written by the row threads
//   straight into the K-major swizzled layout UMMA reads; E_b = the book's hi/lo planes, TMA-streamed in tiles of BN
//   codes); the whole [128 x K] score matrix of a book lives in TMEM (K <= 512 columns) and never reaches memory;
//   epilogue, one thread per row (its TMEM lane):  arg-max of S - 0.5|e|^2 with the runner-up; a code can only be the
//   FP32 kernel's arg-max if its tensor-core score is within tol = 2 x (bf16x3 error bound) of the maximum, so a row
//   whose top two are >= tol apart is decided, and an ambiguous row re-scans its TMEM lane and re-scores the few
//   candidates with the FP32 kernel's arithmetic (one fmaf chain over d, minus 0.5|e|^2, first maximum wins) --
//   the indices equal rvq_books_f32's bit for bit;
//   then the codeword gather and  q_sum = q_sum + (q - r) + r ;  r = r - q  (op order of :433-434), the new residual
//   goes back to shared memory as fp32 (for the exact re-scores) and as bf16 planes (the next book's A operand).
// Warp roles: warp 0 = TMA producer (code tiles; it runs ahead across books, the books do not depend on the residual),
// warp 1 = MMA issuer + TMEM owner, warps 2..5 = the 128 row threads.
#pragma once
#include "kernels_tc.cuh"

namespace b2c {

constexpr int RVQ_TC_THREADS = 192;
constexpr int RVQ_TC_MAX_BOOKS = 16;

struct RvqTcParams {
  const float* x;         // [N, D] fp32 rows
  const float* books;     // [n_books][K][D] fp32 (codeword gather + exact re-scores)
  const float* half_n;    // [n_books][K]
  float* qsum;            // [N, D]
  int* idx;               // [B, books_use, Tl] (or flat [N] when idx_flat)
  int N, D, K, books_use;
  int row_mode, B, Tl, chunk, nfix, idx_flat;
  int BK, n_kblk, BN, n_ntiles, stages, tmem_cols;
  uint32_t a_blk_bytes, a_plane_bytes, b_blk_bytes, b_plane_bytes, b_stage_bytes, sbo, layout_type;
  const float* emax2;     // [n_books] device: max_k |e_k|^2 per book (error bound of the tensor-core scores)
};

// row `r`, elements [c0, c0 + 8) of a [128 x D] operand tile -> its 16-byte unit in the swizzled K-major layout
__device__ __forceinline__ uint32_t rvq_unit_off(const RvqTcParams& p, int r, int c0) {
  const int kb = c0 / p.BK, cin = c0 - kb * p.BK;
  const uint32_t sw = p.BK == 64 ? ((uint32_t)r & 7u) : (((uint32_t)r >> 1) & 3u);
  return (uint32_t)kb * p.a_blk_bytes + (uint32_t)r * ((uint32_t)p.BK * 2u) + ((((uint32_t)cin >> 3)) ^ sw) * 16u;
}

__global__ void __launch_bounds__(RVQ_TC_THREADS, 1)
rvq_tc_kernel(const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
              const __grid_constant__ RvqTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[TC_MAX_STAGES];
  __shared__ __align__(8) uint64_t bar_empty[TC_MAX_STAGES];
  __shared__ __align__(8) uint64_t bar_aready;     // the residual planes of the next book are in shared memory
  __shared__ __align__(8) uint64_t bar_sfull;      // all score tiles of the current book are in TMEM
  __shared__ uint32_t tmem_base_s;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm0 = smem_raw + (smem0 - smem_u32(smem_raw));
  const uint32_t a_bytes = 2u * p.a_plane_bytes;
  const uint32_t smemB = smem0 + a_bytes;
  uint8_t* abuf = sm0;
  float* rs = reinterpret_cast<float*>(sm0 + a_bytes + (size_t)p.stages * p.b_stage_bytes);   // [128][D + 1] fp32 residual
  float* hn_s = rs + TC_BM * (p.D + 1);                                                       // [K] 0.5|e|^2 of the book
  const int DP = p.D + 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmB_hi); tma_prefetch_desc(&tmB_lo);
    for (int s = 0; s < p.stages; ++s) { mbar_init(smem_u32(&bar_full[s]), 1); mbar_init(smem_u32(&bar_empty[s]), 1); }
    mbar_init(smem_u32(&bar_aready), 1);
    mbar_init(smem_u32(&bar_sfull), 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  pdl_sync();
  const int row0 = blockIdx.x * TC_BM;

  if (warp == 0) {
    // ===================== TMA producer: code tiles of every book, in order =====================
    if (lane == 0) {
      Ring rg;
      for (int bk = 0; bk < p.books_use; ++bk)
        for (int nt = 0; nt < p.n_ntiles; ++nt, rg.next(p.stages)) {
          const uint32_t s = rg.s;
          mbar_wait(smem_u32(&bar_empty[s]), rg.par ^ 1u, 1);
          const uint32_t full = smem_u32(&bar_full[s]);
          mbar_expect_tx(full, p.b_stage_bytes);
          const uint32_t dst = smemB + s * p.b_stage_bytes;
          for (int kb = 0; kb < p.n_kblk; ++kb) {
            tma_load_2d(dst + kb * p.b_blk_bytes, &tmB_hi, full, kb * p.BK, bk * p.K + nt * p.BN);
            tma_load_2d(dst + p.b_plane_bytes + kb * p.b_blk_bytes, &tmB_lo, full, kb * p.BK, bk * p.K + nt * p.BN);
          }
        }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = umma_idesc_bf16(TC_BM, p.BN);
    const int ksteps = p.BK / 16;
    const uint32_t desc_hi = (uint32_t)(umma_desc_base(p.sbo, p.layout_type) >> 32);
    const uint32_t a_base = (smem0 & 0x3FFFFu) >> 4;
    Ring rg;
    for (int bk = 0; bk < p.books_use; ++bk) {
      mbar_wait(smem_u32(&bar_aready), (uint32_t)bk & 1u, 2);     // also: the row threads are done with book bk-1's scores
      tc_fence_after();
      for (int nt = 0; nt < p.n_ntiles; ++nt, rg.next(p.stages)) {
        const uint32_t s = rg.s;
        mbar_wait(smem_u32(&bar_full[s]), rg.par, 3);
        tc_fence_after();
        const uint32_t b_base = ((smemB + s * p.b_stage_bytes) & 0x3FFFFu) >> 4;
        const uint32_t d = tmem_base + (uint32_t)(nt * p.BN);
        for (int kb = 0; kb < p.n_kblk; ++kb) {
          const uint32_t a_lo = a_base + ((uint32_t)kb * p.a_blk_bytes >> 4), b_lo = b_base + ((uint32_t)kb * p.b_blk_bytes >> 4);
          if (ksteps == 4) umma_ksteps<1, 4>(d, a_lo, b_lo, p.a_plane_bytes >> 4, p.b_plane_bytes >> 4, desc_hi, idesc, kb != 0);
          else umma_ksteps<1, 2>(d, a_lo, b_lo, p.a_plane_bytes >> 4, p.b_plane_bytes >> 4, desc_hi, idesc, kb != 0);
        }
        umma_commit_w(smem_u32(&bar_empty[s]));
      }
      umma_commit_w(smem_u32(&bar_sfull));
    }
  } else {
    // ===================== row threads: thread = token row = TMEM lane =====================
    const int row = (warp & 3) * 32 + lane;               // warp w may read TMEM lanes [32 (w % 4), 32 (w % 4) + 32)
    const int n = row0 + row;
    const bool live = n < p.N;
    float* rr = rs + row * DP;
    const int rt = threadIdx.x - 64;                      // 0..127
    // residual <- x ; planes of x
    float xn2 = 0.f;
    for (int c0 = 0; c0 < p.D; c0 += 8) {
      float v[8];
      if (live) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(p.x + (size_t)n * p.D + c0));
        const float4 b = __ldg(reinterpret_cast<const float4*>(p.x + (size_t)n * p.D + c0 + 4));
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.f;
      }
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
        const float2 f = __bfloat1622float2(h);
        const __nv_bfloat162 l = __floats2bfloat162_rn(v[2 * e] - f.x, v[2 * e + 1] - f.y);
        hi[e] = *reinterpret_cast<const uint32_t*>(&h);
        lo[e] = *reinterpret_cast<const uint32_t*>(&l);
      }
      const uint32_t off = rvq_unit_off(p, row, c0);
      *reinterpret_cast<uint4*>(abuf + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(abuf + p.a_plane_bytes + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
#pragma unroll
      for (int i = 0; i < 8; ++i) { rr[c0 + i] = v[i]; xn2 = fmaf(v[i], v[i], xn2); }
    }
    fence_proxy_async();                                  // generic-proxy smem writes -> async proxy (UMMA reads)
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (rt == 0) mbar_arrive(smem_u32(&bar_aready));

    for (int bk = 0; bk < p.books_use; ++bk) {
      const float* hn_g = p.half_n + (size_t)bk * p.K;
      for (int k = rt; k < p.K; k += 128) hn_s[k] = __ldg(hn_g + k);
      asm volatile("bar.sync 1, 128;" ::: "memory");
      mbar_wait(smem_u32(&bar_sfull), (uint32_t)bk & 1u, 4);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
      // ---- pass 1: tensor-core arg-max with runner-up
      float best = -INFINITY, second = -INFINITY;
      int bidx = 0;
      for (int c = 0; c < p.K; c += 16) {
        float v[16];
        tmem_ld16(t_row + c, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float sc = __fsub_rn(v[i], hn_s[c + i]);
          second = fmaxf(second, fminf(best, sc));
          if (sc > best) { best = sc; bidx = c + i; }
        }
      }
      // error bound of a bf16x3 score: 3 * 2^-16 |x||e| (hi.lo, lo.hi rounding, dropped lo.lo) -- doubled and rounded
      // up to 2^-13 |x| max|e|, the tolerance nearest_finalize_rows uses
      const float tol = 1.220703125e-4f * sqrtf(xn2 * 1.0001f * __ldg(p.emax2 + bk)) + 1e-30f;
      const bool amb = live && (best - second < tol);
      if (__any_sync(0xffffffffu, amb)) {
        // ---- pass 2 (rare): every candidate within tol of the maximum, re-scored exactly, first maximum wins
        const float* book = p.books + (size_t)bk * p.K * p.D;
        float ebest = -INFINITY;
        int eidx = bidx;
        for (int c = 0; c < p.K; c += 16) {
          float v[16];
          tmem_ld16(t_row + c, v);                        // warp-collective: all lanes load, ambiguous lanes act
          if (!amb) continue;
#pragma unroll 1
          for (int i = 0; i < 16; ++i) {
            if (__fsub_rn(v[i], hn_s[c + i]) < best - tol) continue;
            const float* e = book + (size_t)(c + i) * p.D;
            float acc = 0.f;
            for (int d = 0; d < p.D; d += 4) {
              const float4 e4 = __ldg(reinterpret_cast<const float4*>(e + d));
              acc = fmaf(rr[d], e4.x, acc); acc = fmaf(rr[d + 1], e4.y, acc);
              acc = fmaf(rr[d + 2], e4.z, acc); acc = fmaf(rr[d + 3], e4.w, acc);
            }
            const float sc = __fsub_rn(acc, hn_s[c + i]);
            if (sc > ebest) { ebest = sc; eidx = c + i; }
          }
        }
        if (amb) bidx = eidx;
      }
      tc_fence_before();
      // ---- codeword gather, q_sum = q_sum + (q - r) + r, r = r - q, planes of the new residual
      xn2 = 0.f;
      if (live) {
        const float* q_row = p.books + ((size_t)bk * p.K + bidx) * p.D;
        float* qs_row = p.qsum + (size_t)n * p.D;
        const bool last = bk + 1 == p.books_use;
        for (int c0 = 0; c0 < p.D; c0 += 8) {
          const float4 qa = __ldg(reinterpret_cast<const float4*>(q_row + c0)), qb = __ldg(reinterpret_cast<const float4*>(q_row + c0 + 4));
          const float q[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
          float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          if (bk > 0) {
            const float4 sa = *reinterpret_cast<const float4*>(qs_row + c0), sb = *reinterpret_cast<const float4*>(qs_row + c0 + 4);
            s[0] = sa.x; s[1] = sa.y; s[2] = sa.z; s[3] = sa.w; s[4] = sb.x; s[5] = sb.y; s[6] = sb.z; s[7] = sb.w;
          }
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float r = rr[c0 + i];
            s[i] = __fadd_rn(__fadd_rn(s[i], __fsub_rn(q[i], r)), r);
            v[i] = __fsub_rn(r, q[i]);
            rr[c0 + i] = v[i];
            xn2 = fmaf(v[i], v[i], xn2);
          }
          *reinterpret_cast<float4*>(qs_row + c0) = make_float4(s[0], s[1], s[2], s[3]);
          *reinterpret_cast<float4*>(qs_row + c0 + 4) = make_float4(s[4], s[5], s[6], s[7]);
          if (!last) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
              const float2 f = __bfloat1622float2(h);
              const __nv_bfloat162 l = __floats2bfloat162_rn(v[2 * e] - f.x, v[2 * e + 1] - f.y);
              hi[e] = *reinterpret_cast<const uint32_t*>(&h);
              lo[e] = *reinterpret_cast<const uint32_t*>(&l);
            }
            const uint32_t off = rvq_unit_off(p, row, c0);
            *reinterpret_cast<uint4*>(abuf + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4*>(abuf + p.a_plane_bytes + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          }
        }
        if (p.idx_flat) p.idx[n] = bidx;
        else {
          int b, tt;
          if (p.row_mode == ROWS_DENSE) { b = n / p.Tl; tt = n - b * p.Tl; }
          else { b = n / p.nfix; tt = p.chunk * (n - b * p.nfix + 1); }
          p.idx[((long)b * p.books_use + bk) * p.Tl + tt] = bidx;
        }
      }
      if (bk + 1 < p.books_use) {
        fence_proxy_async();
        asm volatile("bar.sync 1, 128;" ::: "memory");   // all rows: scores read, planes + hn_s no longer needed / written
        if (rt == 0) mbar_arrive(smem_u32(&bar_aready));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols)
                 : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct TcBooks {
  __nv_bfloat16* hi = nullptr;   // [n_books * K][D], K-major
  __nv_bfloat16* lo = nullptr;
  int n_books = 0, K = 0, D = 0;
};

struct RvqTcPlan {
  RvqTcParams q;
  size_t smem = 0;
  int grid = 0;
  CUtensorMap mB_hi, mB_lo;
  bool ready = false;
};

inline bool rvq_tc_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B2C_RVQ_TC");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

// 0 = planned; > 0 = shape not served by this kernel (the FP32 kernels take it)
inline int rvq_tc_plan(int N, int D, int K, int n_books, int books_use, RvqTcPlan* plan) {
  if (!rvq_tc_enabled()) return 9;
  if (D % 32 != 0 || D < 32 || D > 128) return 1;
  if (K % 64 != 0 || K < 64 || K > 512) return 2;
  if (n_books > RVQ_TC_MAX_BOOKS || books_use < 1 || books_use > n_books) return 3;
  if (!tc_encode_fn()) return 4;
  RvqTcParams& p = plan->q;
  memset(&p, 0, sizeof(p));
  p.N = N; p.D = D; p.K = K; p.books_use = books_use;
  p.BK = D % 64 == 0 ? 64 : 32;
  p.n_kblk = D / p.BK;
  p.a_blk_bytes = TC_BM * p.BK * 2;
  p.a_plane_bytes = (uint32_t)p.n_kblk * p.a_blk_bytes;
  p.sbo = 8 * p.BK * 2;
  p.layout_type = p.BK == 64 ? 2u : 4u;
  p.tmem_cols = K <= 64 ? 64 : (K <= 128 ? 128 : (K <= 256 ? 256 : 512));
  const long fixed = 2L * p.a_plane_bytes + (long)TC_BM * (D + 1) * 4 + (long)K * 4 + 1024 + 256;
  bool ok = false;
  for (int bn = K >= 128 ? 128 : 64; bn >= 64 && !ok; bn >>= 1) {
    if (K % bn) continue;
    p.BN = bn;
    p.n_ntiles = K / bn;
    p.b_blk_bytes = (uint32_t)bn * p.BK * 2;
    p.b_plane_bytes = (uint32_t)p.n_kblk * p.b_blk_bytes;
    p.b_stage_bytes = 2u * p.b_plane_bytes;
    long st = (232448L - 2048 - fixed) / (long)p.b_stage_bytes;
    if (st >= 2) {
      p.stages = (int)(st > TC_MAX_STAGES ? TC_MAX_STAGES : st);
      ok = true;
    }
  }
  if (!ok) return 5;
  plan->smem = (size_t)fixed + (size_t)p.stages * p.b_stage_bytes;
  plan->grid = (N + TC_BM - 1) / TC_BM;
  plan->ready = false;
  return 0;
}

inline int rvq_tc_launch(RvqTcPlan& plan, const RvqTcParams& args, const TcBooks& tb, cudaStream_t st) {
  RvqTcParams q = plan.q;
  q.x = args.x; q.books = args.books; q.half_n = args.half_n; q.qsum = args.qsum; q.idx = args.idx;
  q.row_mode = args.row_mode; q.B = args.B; q.Tl = args.Tl; q.chunk = args.chunk; q.nfix = args.nfix; q.idx_flat = args.idx_flat;
  q.emax2 = args.emax2;
  if (!plan.ready) {
    cuuint64_t dims[2] = {(cuuint64_t)q.D, (cuuint64_t)tb.n_books * tb.K};
    cuuint64_t str[1] = {(cuuint64_t)q.D * 2};
    cuuint32_t box[2] = {(cuuint32_t)q.BK, (cuuint32_t)q.BN};
    int rc = tc_encode(&plan.mB_hi, tb.hi, 2, dims, str, box, q.BK);
    if (!rc) rc = tc_encode(&plan.mB_lo, tb.lo, 2, dims, str, box, q.BK);
    if (rc) return rc;
    plan.ready = true;
  }
  if (cudaFuncSetAttribute(rvq_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem) != cudaSuccess) return -2;
  tc_launch(rvq_tc_kernel, plan.grid, RVQ_TC_THREADS, plan.smem, st, plan.mB_hi, plan.mB_lo, q);
  return 0;
}

}  // namespace b2c
