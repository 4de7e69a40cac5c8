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
constexpr int RVQ_TC_MAXC = 32;          // exact re-score candidates per ambiguous row handled by the warp (one per lane)

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
  int* cand_s = reinterpret_cast<int*>(hn_s + p.K);                                            // [128][RVQ_TC_MAXC] candidate codes
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
      // ---- pass 1: tensor-core arg-max with runner-up.  32 columns per TMEM load, two independent (best, runner-up)
      // chains (even / odd columns) so the compare-select dependences overlap; hn_s read as float4.
      float b0 = -INFINITY, s0 = -INFINITY, b1 = -INFINITY, s1 = -INFINITY;
      int i0 = 0, i1 = 1;
      for (int c = 0; c < p.K; c += 32) {
        uint32_t raw[32];
        tmem_ld32_issue(t_row + c, raw);
        float hn[32];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float4 h4 = *reinterpret_cast<const float4*>(hn_s + c + 4 * u);
          hn[4 * u] = h4.x; hn[4 * u + 1] = h4.y; hn[4 * u + 2] = h4.z; hn[4 * u + 3] = h4.w;
        }
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float sa = __fsub_rn(__uint_as_float(raw[i]), hn[i]);
          const float sb = __fsub_rn(__uint_as_float(raw[i + 1]), hn[i + 1]);
          s0 = fmaxf(s0, fminf(b0, sa));
          if (sa > b0) { b0 = sa; i0 = c + i; }
          s1 = fmaxf(s1, fminf(b1, sb));
          if (sb > b1) { b1 = sb; i1 = c + i + 1; }
        }
      }
      const bool take1 = b1 > b0 || (b1 == b0 && i1 < i0);
      const float best = take1 ? b1 : b0;
      const float second = fmaxf(fmaxf(s0, s1), take1 ? b0 : b1);
      int bidx = take1 ? i1 : i0;
      // error bound of a bf16x3 score: 3 * 2^-16 |x||e| (hi.lo, lo.hi rounding, dropped lo.lo) -- doubled and rounded
      // up to 2^-13 |x| max|e|, the tolerance nearest_finalize_rows uses
      const float tol = 1.220703125e-4f * sqrtf(xn2 * 1.0001f * __ldg(p.emax2 + bk)) + 1e-30f;
      const bool amb = live && (best - second < tol);
      unsigned amb_mask = __ballot_sync(0xffffffffu, amb);
      if (amb_mask) {
        // ---- pass 2 (a few rows per warp and book): every candidate within tol of the maximum is re-scored exactly.
        // One more sweep over the warp's TMEM lanes collects each ambiguous row's candidates (ascending code order) in
        // shared memory; then the WARP re-scores one row at a time, lane j = candidate j: the FP32 kernel's arithmetic
        // (one fmaf chain over d from 0, minus 0.5|e|^2) with the 24 codeword loads of a chain independent of each
        // other, and a shuffle arg-max with the smaller code winning ties (= first maximum).  A row with more than
        // RVQ_TC_MAXC candidates falls back to re-scoring in its own lane.
        const float* book = p.books + (size_t)bk * p.K * p.D;
        int* my_c = cand_s + row * RVQ_TC_MAXC;
        int nc = 0;
        for (int c = 0; c < p.K; c += 32) {
          uint32_t raw[32];
          tmem_ld32_issue(t_row + c, raw);                // warp-collective: all lanes load, ambiguous lanes act
          tmem_ld_wait();
          if (!amb) continue;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (__fsub_rn(__uint_as_float(raw[i]), hn_s[c + i]) >= best - tol) {
              if (nc < RVQ_TC_MAXC) my_c[nc] = c + i;
              ++nc;
            }
          }
        }
        __syncwarp();
        const bool overflow = amb && nc > RVQ_TC_MAXC;
        if (overflow) {                                   // pathological: exact arg-max over the whole book in this lane
          float ebest = -INFINITY;
          int eidx = bidx;
          for (int k = 0; k < p.K; ++k) {
            const float* e = book + (size_t)k * p.D;
            float acc = 0.f;
            for (int d = 0; d < p.D; d += 4) {
              const float4 e4 = __ldg(reinterpret_cast<const float4*>(e + d));
              acc = fmaf(rr[d], e4.x, acc); acc = fmaf(rr[d + 1], e4.y, acc);
              acc = fmaf(rr[d + 2], e4.z, acc); acc = fmaf(rr[d + 3], e4.w, acc);
            }
            const float sc = __fsub_rn(acc, hn_s[k]);
            if (sc > ebest) { ebest = sc; eidx = k; }
          }
          bidx = eidx;
        }
        amb_mask = __ballot_sync(0xffffffffu, amb && !overflow);
        while (amb_mask) {
          const int src = __ffs(amb_mask) - 1;
          amb_mask &= amb_mask - 1;
          const int n_c = __shfl_sync(0xffffffffu, nc, src);
          const int srow = (warp & 3) * 32 + src;
          const float* rsrc = rs + srow * DP;
          float sc = -INFINITY;
          int code = 0x7fffffff;
          if (lane < n_c) {
            code = cand_s[srow * RVQ_TC_MAXC + lane];
            const float4* e = reinterpret_cast<const float4*>(book + (size_t)code * p.D);
            float acc = 0.f;
            for (int d = 0; d < p.D; d += 16) {
              const float4 e0 = __ldg(e + (d >> 2)), e1 = __ldg(e + (d >> 2) + 1), e2 = __ldg(e + (d >> 2) + 2), e3 = __ldg(e + (d >> 2) + 3);
              acc = fmaf(rsrc[d], e0.x, acc); acc = fmaf(rsrc[d + 1], e0.y, acc); acc = fmaf(rsrc[d + 2], e0.z, acc); acc = fmaf(rsrc[d + 3], e0.w, acc);
              acc = fmaf(rsrc[d + 4], e1.x, acc); acc = fmaf(rsrc[d + 5], e1.y, acc); acc = fmaf(rsrc[d + 6], e1.z, acc); acc = fmaf(rsrc[d + 7], e1.w, acc);
              acc = fmaf(rsrc[d + 8], e2.x, acc); acc = fmaf(rsrc[d + 9], e2.y, acc); acc = fmaf(rsrc[d + 10], e2.z, acc); acc = fmaf(rsrc[d + 11], e2.w, acc);
              acc = fmaf(rsrc[d + 12], e3.x, acc); acc = fmaf(rsrc[d + 13], e3.y, acc); acc = fmaf(rsrc[d + 14], e3.z, acc); acc = fmaf(rsrc[d + 15], e3.w, acc);
            }
            sc = __fsub_rn(acc, hn_s[code]);
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const float so = __shfl_xor_sync(0xffffffffu, sc, o);
            const int co = __shfl_xor_sync(0xffffffffu, code, o);
            if (so > sc || (so == sc && co < code)) { sc = so; code = co; }
          }
          if (lane == src) bidx = code;
        }
      }
      tc_fence_before();
      // ---- codeword gather, q_sum = q_sum + (q - r) + r, r = r - q, planes of the new residual
      xn2 = 0.f;
      if (live) {
        const float* q_row = p.books + ((size_t)bk * p.K + bidx) * p.D;
        float* qs_row = p.qsum + (size_t)n * p.D;
        const bool last = bk + 1 == p.books_use;
#pragma unroll 4                                       // D / 8 is a multiple of 4: the codeword / q_sum loads of 4 units are in flight together
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
  const long fixed = 2L * p.a_plane_bytes + (long)TC_BM * (D + 1) * 4 + (long)K * 4 + (long)TC_BM * RVQ_TC_MAXC * 4 + 1024 + 256;
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
