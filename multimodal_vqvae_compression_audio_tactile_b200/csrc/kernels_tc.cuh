// tcgen05 / TMEM / TMA implicit-GEMM Conv1d for sm_100a.
//
//   D[128 positions x BN out-channels] (fp32, TMEM) += A[128 x BK] (activations, bf16, smem via TMA)
//                                                      * B[BN x BK]^T (weights, bf16, smem via TMA)
// over (tap, input-channel block).  Activations are channel-last so both operands are K-major and a
// tap is a row offset of the TMA box; zero padding is the TMA out-of-bounds fill.  Warp roles:
// warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane) + TMEM allocator, warps 2..9 = epilogue
// (tcgen05.ld -> bias / residual / snake -> global).  Persistent CTAs, static tile round-robin,
// two TMEM accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// Precision modes: single-pass bf16, or "bf16x3": x = hi + lo (two bf16 planes, 16 mantissa bits),
// D += A_hi*B_hi + A_hi*B_lo + A_lo*B_hi, fp32 accumulate.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "kernels_f32.cuh"

namespace b2c {

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P1;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait: a protocol bug traps (the launch fails) instead of hanging the GPU
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P1;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int tag) {
#ifdef B2C_MBAR_POLL
  // experiment: non-suspending poll
  for (uint32_t spin = 0; spin < (1u << 26); ++spin)
    if (mbar_test_wait(bar, parity)) return;
  printf("b2c conv_tc: mbarrier poll timeout tag=%d block=%d thread=%d\n", tag, blockIdx.x, threadIdx.x);
  __trap();
#endif
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (clock64() - t0 < 3000000000LL) { // ~1.5 s at 1.9 GHz
#ifdef B2C_WAIT_NAP
    __nanosleep(B2C_WAIT_NAP);            // experiment: sleep between polls (power of the waiting warps vs wake-up latency)
#endif
    if (mbar_try_wait(bar, parity)) return;
  }
  printf("b2c conv_tc: mbarrier timeout tag=%d block=%d thread=%d\n", tag, blockIdx.x, threadIdx.x);
  __trap();
}
// position in a ring of n mbarrier-guarded slots: slot index + phase parity, advanced without integer division
// (two runtime divisions per pipeline step cost more than the MMAs of a narrow layer)
struct Ring {
  uint32_t s = 0, par = 0;
  __device__ __forceinline__ void next(uint32_t n) {
    if (++s == n) { s = 0; par ^= 1u; }
  }
};

// one lane polls, the rest of the warp waits at __syncwarp (32 lanes polling the same mbarrier slow every
// other mbarrier operation of the CTA down)
__device__ __forceinline__ void mbar_wait_warp(uint32_t bar, uint32_t parity, int tag, int lane) {
  (void)lane;
  mbar_wait(bar, parity, tag);   // measured: every lane polling is slightly FASTER than one lane + __syncwarp
}
// one thread of the 16 epilogue warps polls, the other 511 sleep in a hardware named barrier
__device__ __forceinline__ void mbar_wait_epilogue(uint32_t bar, uint32_t parity, int tag) {
  mbar_wait(bar, parity, tag);   // measured: all 512 threads polling beats one poller + a named barrier
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// Programmatic dependent launch: a kernel launched with the programmatic-serialization attribute may start while its
// predecessor in the stream still runs.  Everything before pdl_sync() (mbarrier init, TMEM allocation, tensor-map
// prefetch -- nothing that touches the predecessor's output) then overlaps the predecessor's tail; pdl_sync() waits
// for the predecessor's completion and memory flush, and releases this grid's own successor (whose CTAs take an SM
// as soon as one of ours exits, and wait in turn).  Both instructions are no-ops in a normally launched grid.
__device__ __forceinline__ void pdl_sync() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// TMA store (shared -> global, bulk async-group completion): the epilogues that stage a whole output tile in shared
// memory hand it to the copy engine instead of storing from 512 threads
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// K-major, swizzled shared-memory matrix descriptor (sm_100 UMMA): one swizzle atom along K
// (BK * 2 bytes == swizzle span), 8-row groups SBO bytes apart.  Advancing along K inside the
// atom = adding bytes to the start address (the swizzle is a function of the address bits).
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);        // start address, bits [0,14)
  d |= (uint64_t)0 << 16;                              // leading byte offset (unused: one atom along K)
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;   // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                              // descriptor version (sm_100)
  d |= (uint64_t)(layout_type & 7u) << 61;             // 2 = SWIZZLE_128B, 4 = SWIZZLE_64B
  return d;
}
// The constant high part of the descriptor; the low word is (smem byte address >> 4), so moving the
// start address by n bytes is an integer add of n >> 4 (shared memory addresses are < 2^18).
__device__ __forceinline__ uint64_t umma_desc_base(uint32_t sbo_bytes, uint32_t layout_type) {
  return ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | ((uint64_t)1 << 46) | ((uint64_t)(layout_type & 7u) << 61);
}
// instruction descriptor, kind::f16: D fp32, A/B bf16, both K-major, M = 128
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Converged-warp variants: every lane executes them, elect.sync picks the one lane that issues.  ptxas then
// emits a predicated UTCHMMA with uniform-register operands instead of a per-active-thread serialisation loop.
__device__ __forceinline__ void umma_bf16_w(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred pe, p;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// same, descriptors given as (low word, shared constant high word): the issue loop then only does 32-bit adds
__device__ __forceinline__ void umma_bf16_w32(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi,
                                              uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred pe, p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(desc_hi)
      : "memory");
}
// KS K-steps (16 elements = 32 bytes = 2 descriptor units each) of one operand pair, fully unrolled
template <int X3, int KS>
__device__ __forceinline__ void umma_ksteps(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t a_plane, uint32_t b_plane,
                                            uint32_t desc_hi, uint32_t idesc, uint32_t acc_first) {
#pragma unroll
  for (int k = 0; k < KS; ++k) {
    umma_bf16_w32(d_tmem, a_lo + 2 * k, b_lo + 2 * k, desc_hi, idesc, k == 0 ? acc_first : 1u);
    if (X3) {
      umma_bf16_w32(d_tmem, a_lo + 2 * k, b_lo + b_plane + 2 * k, desc_hi, idesc, 1u);
      umma_bf16_w32(d_tmem, a_lo + a_plane + 2 * k, b_lo + 2 * k, desc_hi, idesc, 1u);
    }
  }
}
// bf16x3 with the two weight planes stacked along N: B' = [B_hi ; B_lo] is ONE K-major tile of 2N rows (the lo plane
// follows the hi plane in shared memory), so  D[:, 0:2N] += A_hi . B'^T  gives hi.hi in columns [0, N) and hi.lo in
// [N, 2N) with A_hi fetched once, and  D[:, 0:N] += A_lo . B_hi^T  completes the first half; the epilogue adds the two
// halves.  Operand fetch per K step: (4 + 4N/32.. ) -- 14 KB instead of 18 KB at N = 64, where the MMAs are paced by the
// shared-memory fetch, not by the tensor pipe (tools/micro/umma_rate.cu).
template <int KS>
__device__ __forceinline__ void umma_ksteps_stacked(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t a_plane,
                                                    uint32_t desc_hi, uint32_t idesc_2n, uint32_t idesc_n, uint32_t acc_first) {
#pragma unroll
  for (int k = 0; k < KS; ++k) {
    umma_bf16_w32(d_tmem, a_lo + 2 * k, b_lo + 2 * k, desc_hi, idesc_2n, k == 0 ? acc_first : 1u);
    umma_bf16_w32(d_tmem, a_lo + a_plane + 2 * k, b_lo + 2 * k, desc_hi, idesc_n, 1u);
  }
}
__device__ __forceinline__ void umma_commit_w(uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
// 32 consecutive columns of this thread's TMEM lane (issue only; pair with tmem_ld_wait before reading v)
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]),
        "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]),
        "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ---------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------
struct TcConvParams {
  const float* bias;
  const float* res;
  const float* dmul;    // backward-data pass: (acc + bias) * snake'(dmul[out position]; alpha), then + res
  float* out_raw;
  void* out_act;        // FMT_F32: float*; FMT_PLANES / FMT_HI: bf16 hi plane, lo plane = hi + act_plane_elems
  const float* alpha;
  const float* inv_alpha;   // 1 / (alpha + 1e-9), fp32, precomputed at pack time
  long act_plane_elems;
  int B, Lin, Cin, Cout, KT, in_step, dil, n_phase, Lj, out_step, Lout;
  int in_off[8], out_off[8];
  int act, res_mode, Tl, chunk, out_fmt;
  int BN, BK, n_kblk, stages, tiles_j, n_ntiles, total_tiles;
  int tmem_cols, acc_stride;
  uint32_t a_bytes, b_bytes, sbo, layout_type;
  // nearest-code search epilogue (EPI = 1): scores = acc - half_norm[code]; per (row, code tile) the best
  // code, its score and the runner-up score go to cand[row * n_ntiles + nt] = {best, idx, second, -}
  const float* half_norm;
  float4* cand;
  int n_rows, n_codes;
  // slab kernel (conv_tc2_kernel): MT m-tiles per work item share one activation slab and every weight tile
  int MT, groups_j, slab_rows, box_rows, n_aloads, SA, SB;
  int epi_groups; // 2: the 16 epilogue warps work as two independent groups of 8, one per TMEM accumulator buffer
  int kgroup;     // K blocks per pipeline stage of the generic kernel (one mbarrier hand-off per kgroup blocks)
  int dbg;        // B2C_TC_DEBUG bit mask (timing experiments only): 1 skip epilogue work, 2 no TMA, 4 no MMA
  int pair_off;   // > 0: the accumulator is the sum of two column ranges pair_off apart (N-stacked bf16x3, fused unit)
  int stg_bufs;   // epilogue staging tiles: 2 (one barrier per chunk) or 1 (two barriers, frees 18 KB for the rings)
  uint32_t row_bytes, a_plane_bytes;
  // frame-packed tiles (generic kernel, short sequences): an M tile is JB consecutive positions of BB consecutive frames
  // (JB * BB = 128, both powers of two) instead of 128 positions of one frame -- 75 positions per frame fill 59 % of a
  // 128-row tile, 4 positions x 32 frames fill 99 %.  Row m of a tile = frame (m >> jb_shift), position (m & (JB - 1)):
  // that is the order in which the TMA box {BK, JB, BB} lands in shared memory, and per-frame zero padding is still the
  // tensor map's out-of-bounds fill.  JB == 0: the legacy tile (JB = 128, BB = 1).  Rows are independent in an MMA, so
  // a frame's bits do not depend on how tiles are packed.
  int JB, jb_shift, BB;
  int stg_lock;   // fused unit: two epilogue groups share one staging tile under a shared-memory lock
};

// tile (frame-group index bg, position-tile index jt), row m of the tile -> frame b and position j
__device__ __forceinline__ void tc_row_coords(const TcConvParams& p, int bg, int jt, int m, int& b, int& j) {
  if (p.JB == 0) { b = bg; j = jt * 128 + m; }
  else { b = bg * p.BB + (m >> p.jb_shift); j = jt * p.JB + (m & (p.JB - 1)); }
}

constexpr int TC_BM = 128;
constexpr int TC_MAX_STAGES = 8;
constexpr int TC_EPI_WARPS = 16;
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;   // TMA warp + MMA warp + epilogue warps
constexpr int TC_STG_LD = 36;                        // floats per staging row: 32 columns + 4 pad (bank-conflict free)
constexpr int TC_STG_BYTES = 2 * TC_BM * TC_STG_LD * 4;


// snake'(x) with the per-channel reciprocal 1 / (alpha + 1e-9) precomputed (backward-data epilogues)
__device__ __forceinline__ float dsnake_pre(float v, float alpha, float inv) {
  return fmaf(sinf(2.0f * alpha * v), alpha * inv, 1.0f);
}

// L2 prefetch of the residual rows of one output tile (issued one tile ahead by the 512 epilogue threads): the
// epilogue's residual loads are latency-bound otherwise -- only one 32-column chunk (16 KB per SM) is in flight.
__device__ __forceinline__ void tc_prefetch_res(const TcConvParams& p, int b, int ph, int jt, int nt) {
  if (!p.res || p.res_mode == 1) return;
  const int et = threadIdx.x - 64;
  const int lines_per_row = p.BN >> 5;                       // 128-byte lines of one tile row
  for (int idx = et; idx < TC_BM * lines_per_row; idx += 32 * TC_EPI_WARPS) {
    const int row = idx / lines_per_row, seg = idx - row * lines_per_row;
    int br, j;
    tc_row_coords(p, b, jt, row, br, j);
    const int lo = j * p.out_step + p.out_off[ph];
    if (j < p.Lj && br < p.B && lo >= 0 && lo < p.Lout) {
      const float* a = p.res + ((size_t)br * p.Lout + lo) * p.Cout + nt * p.BN + seg * 32;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
    }
  }
}

// Epilogue of one 128 x BN accumulator (16 warps).  Per 32-column chunk: (1) every warp copies its TMEM
// quadrant (thread = row, 8 columns) into a padded fp32 staging tile in shared memory; (2) after a named
// barrier the 512 threads re-read the tile row-major (8 lanes x float4 = one 128-byte row segment), so the
// bias / residual loads and the raw / activated stores are coalesced.  Two staging tiles alternate: one
// barrier per chunk.  When tempty_bar != 0 it is arrived on once the accumulator has been drained.
template <int SFU, int DM = 0>
__device__ __forceinline__ void tc_epilogue_tile(const TcConvParams& p, float* stg, uint32_t& chunk_ctr, uint32_t t_acc,
                                                 int b, int ph, int jt, int nt, uint32_t tempty_bar, int warp, int lane) {
  const int ew = warp - 2;
  const int quad = warp & 3;            // TMEM lanes [32*quad, 32*quad+32) are the ones this warp may read
  const int cg8 = ew >> 2;              // which 8 columns of the chunk this warp copies
  const int et = threadIdx.x - 64;      // 0..511
  const int cq = et & 7;                // float4 column group of the chunk in the row-major phase
  const int r0 = et >> 3;               // rows r0 and r0 + 64
  const int nchunks = p.BN >> 5;
  bool valid[2];
  size_t orow[2], rrow[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    int br, j;
    tc_row_coords(p, b, jt, r0 + 64 * i, br, j);
    const int lo = j * p.out_step + p.out_off[ph];
    valid[i] = (j < p.Lj) && (br < p.B) && (lo >= 0) && (lo < p.Lout);
    const int los = valid[i] ? lo : 0;
    orow[i] = ((size_t)(valid[i] ? br : 0) * p.Lout + los) * p.Cout;
    rrow[i] = p.res_mode == 1 ? (size_t)((los % p.Tl) % p.chunk) * p.Cout : orow[i];
  }
  const uint32_t t_src = t_acc + ((uint32_t)(quad * 32) << 16) + cg8 * 8;
  for (int c = 0; c < nchunks; ++c, ++chunk_ctr) {
    float* sb = stg + (p.stg_bufs == 2 ? (chunk_ctr & 1u) * (TC_BM * TC_STG_LD) : 0u);
    const int co = nt * p.BN + c * 32 + cq * 4;
    float4 rr[2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
      rr[i] = (p.res && valid[i]) ? __ldg(reinterpret_cast<const float4*>(p.res + rrow[i] + co))
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 bb = make_float4(0.f, 0.f, 0.f, 0.f), al = bb, ia = bb;
    if (p.bias) bb = __ldg(reinterpret_cast<const float4*>(p.bias + co));
    if ((p.out_act && p.act == ACT_SNAKE) || DM) {
      al = __ldg(reinterpret_cast<const float4*>(p.alpha + co));
      ia = __ldg(reinterpret_cast<const float4*>(p.inv_alpha + co));
    }
    float4 dm[2];
    if (DM) {
#pragma unroll
      for (int i = 0; i < 2; ++i)
        dm[i] = valid[i] ? __ldg(reinterpret_cast<const float4*>(p.dmul + orow[i] + co)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    {
      float v[8];
      tmem_ld8(t_src + c * 32, v);
      if (p.pair_off) {
        float v2[8];
        tmem_ld8(t_src + p.pair_off + c * 32, v2);
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] += v2[u];
      }
      float* dst = sb + (quad * 32 + lane) * TC_STG_LD + cg8 * 8;
      *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(dst + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
    const bool last = (c == nchunks - 1) && tempty_bar != 0;
    if (last) tc_fence_before();
    asm volatile("bar.sync 1, 512;" ::: "memory");
    if (last && et == 0) mbar_arrive(tempty_bar);   // accumulator (set) drained
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      if (!valid[i]) continue;
      float4 a = *reinterpret_cast<const float4*>(sb + (r0 + 64 * i) * TC_STG_LD + cq * 4);
      a.x += bb.x; a.y += bb.y; a.z += bb.z; a.w += bb.w;
      if (DM) {
        a.x *= dsnake_pre(dm[i].x, al.x, ia.x); a.y *= dsnake_pre(dm[i].y, al.y, ia.y);
        a.z *= dsnake_pre(dm[i].z, al.z, ia.z); a.w *= dsnake_pre(dm[i].w, al.w, ia.w);
      }
      if (p.res) { a.x += rr[i].x; a.y += rr[i].y; a.z += rr[i].z; a.w += rr[i].w; }
      if (p.out_raw) *reinterpret_cast<float4*>(p.out_raw + orow[i] + co) = a;
      if (p.out_act) {
        float4 w;
        if (p.act == ACT_SNAKE) {
          w.x = snake_sel<SFU>(a.x, al.x, ia.x); w.y = snake_sel<SFU>(a.y, al.y, ia.y);
          w.z = snake_sel<SFU>(a.z, al.z, ia.z); w.w = snake_sel<SFU>(a.w, al.w, ia.w);
        } else {
          w.x = apply_act(a.x, p.act, 0.f); w.y = apply_act(a.y, p.act, 0.f);
          w.z = apply_act(a.z, p.act, 0.f); w.w = apply_act(a.w, p.act, 0.f);
        }
        if (p.out_fmt == FMT_F32) {
          *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out_act) + orow[i] + co) = w;
        } else {
          const __nv_bfloat162 h01 = __floats2bfloat162_rn(w.x, w.y), h23 = __floats2bfloat162_rn(w.z, w.w);
          __nv_bfloat16* oh = reinterpret_cast<__nv_bfloat16*>(p.out_act) + orow[i] + co;
          *reinterpret_cast<uint2*>(oh) = make_uint2(*reinterpret_cast<const uint32_t*>(&h01),
                                                     *reinterpret_cast<const uint32_t*>(&h23));
          if (p.out_fmt == FMT_PLANES) {
            const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
            const __nv_bfloat162 l01 = __floats2bfloat162_rn(w.x - f01.x, w.y - f01.y);
            const __nv_bfloat162 l23 = __floats2bfloat162_rn(w.z - f23.x, w.w - f23.y);
            *reinterpret_cast<uint2*>(oh + p.act_plane_elems) =
                make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
          }
        }
      }
    }
    if (p.stg_bufs != 2) asm volatile("bar.sync 1, 512;" ::: "memory");   // the single tile is rewritten next chunk
  }
}

// ---------------------------------------------------------------------------------------------
// Two-group epilogue.  The 16 epilogue warps form two groups of 8 (256 threads); group g owns the tiles whose
// accumulator is TMEM buffer g, its own staging tile and its own named barrier (id 1 + g), so one group can be
// in its TMEM-load / barrier phase while the other computes and stores: with all 16 warps in lock step
// (tc_epilogue_tile) nothing overlapped the latency of either phase.
// Mapping inside a group: TMEM -> staging: warp = (lane quadrant, 16-column half); row-major phase: thread =
// (float4 column group cq, row r0 + 32 i, i < 4).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void epi_group_sync(int g) {
  asm volatile("bar.sync %0, 256;" ::"r"(1 + g) : "memory");
}
__device__ __forceinline__ void tc_prefetch_res_g(const TcConvParams& p, int b, int ph, int jt, int nt) {
  if (!p.res || p.res_mode == 1) return;
  const int et = (threadIdx.x - 64) & 255;
  const int lines_per_row = p.BN >> 5;
  for (int idx = et; idx < TC_BM * lines_per_row; idx += 256) {
    const int row = idx / lines_per_row, seg = idx - row * lines_per_row;
    int br, j;
    tc_row_coords(p, b, jt, row, br, j);
    const int lo = j * p.out_step + p.out_off[ph];
    if (j < p.Lj && br < p.B && lo >= 0 && lo < p.Lout) {
      const float* a = p.res + ((size_t)br * p.Lout + lo) * p.Cout + nt * p.BN + seg * 32;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
    }
  }
}

// `lock` != nullptr: both groups share ONE staging tile (shared memory has no room for two): a group owns it only from
// its TMEM registers -> staging write to the row-major re-read (a few hundred cycles per chunk); the loads before and
// the arithmetic and stores after run from registers, outside the critical section.
template <int SFU, int DM = 0>
__device__ __forceinline__ void tc_epilogue_tile_g(const TcConvParams& p, float* stg_g, int g, uint32_t t_acc, int b, int ph,
                                                   int jt, int nt, uint32_t tempty_bar, int warp, int lane, int* lock = nullptr) {
  const int wg = (warp - 2) & 7;
  const int quad = warp & 3;
  const int half = wg >> 2;
  const int et = (threadIdx.x - 64) & 255;
  const int cq = et & 7;
  const int r0 = et >> 3;               // rows r0 + 32 i
  const int nchunks = p.BN >> 5;
  bool valid[4];
  size_t orow[4], rrow[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int br, j;
    tc_row_coords(p, b, jt, r0 + 32 * i, br, j);
    const int lo = j * p.out_step + p.out_off[ph];
    valid[i] = (j < p.Lj) && (br < p.B) && (lo >= 0) && (lo < p.Lout);
    const int los = valid[i] ? lo : 0;
    orow[i] = ((size_t)(valid[i] ? br : 0) * p.Lout + los) * p.Cout;
    rrow[i] = p.res_mode == 1 ? (size_t)((los % p.Tl) % p.chunk) * p.Cout : orow[i];
  }
  const uint32_t t_src = t_acc + ((uint32_t)(quad * 32) << 16) + half * 16;
  for (int c = 0; c < nchunks; ++c) {
    const int co = nt * p.BN + c * 32 + cq * 4;
    float4 rr[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
      rr[i] = (p.res && valid[i]) ? __ldg(reinterpret_cast<const float4*>(p.res + rrow[i] + co))
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 bb = make_float4(0.f, 0.f, 0.f, 0.f), al = bb, ia = bb;
    if (p.bias) bb = __ldg(reinterpret_cast<const float4*>(p.bias + co));
    if ((p.out_act && p.act == ACT_SNAKE) || DM) {
      al = __ldg(reinterpret_cast<const float4*>(p.alpha + co));
      ia = __ldg(reinterpret_cast<const float4*>(p.inv_alpha + co));
    }
    float4 dm[4];
    if (DM) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        dm[i] = valid[i] ? __ldg(reinterpret_cast<const float4*>(p.dmul + orow[i] + co)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    {
      float v[16];
      tmem_ld16(t_src + c * 32, v);
      if (p.pair_off) {
        float v2[16];
        tmem_ld16(t_src + p.pair_off + c * 32, v2);
#pragma unroll
        for (int u = 0; u < 16; ++u) v[u] += v2[u];
      }
      if (lock) {                                         // acquire the shared staging tile
        if (et == 0) {
          while (atomicCAS(lock, 0, 1) != 0) {}
          __threadfence_block();
        }
        epi_group_sync(g);
      }
      float* dst = stg_g + (quad * 32 + lane) * TC_STG_LD + half * 16;
#pragma unroll
      for (int u = 0; u < 16; u += 4) *reinterpret_cast<float4*>(dst + u) = make_float4(v[u], v[u + 1], v[u + 2], v[u + 3]);
    }
    const bool last = (c == nchunks - 1) && tempty_bar != 0;
    if (last) tc_fence_before();
    epi_group_sync(g);
    if (last && et == 0) mbar_arrive(tempty_bar);
    float4 acc4[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) acc4[i] = *reinterpret_cast<const float4*>(stg_g + (r0 + 32 * i) * TC_STG_LD + cq * 4);
    if (lock) {                                           // every thread has its rows in registers: release
      epi_group_sync(g);
      if (et == 0) {
        __threadfence_block();
        atomicExch(lock, 0);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (!valid[i]) continue;
      float4 a = acc4[i];
      a.x += bb.x; a.y += bb.y; a.z += bb.z; a.w += bb.w;
      if (DM) {
        a.x *= dsnake_pre(dm[i].x, al.x, ia.x); a.y *= dsnake_pre(dm[i].y, al.y, ia.y);
        a.z *= dsnake_pre(dm[i].z, al.z, ia.z); a.w *= dsnake_pre(dm[i].w, al.w, ia.w);
      }
      if (p.res) { a.x += rr[i].x; a.y += rr[i].y; a.z += rr[i].z; a.w += rr[i].w; }
      if (p.out_raw) *reinterpret_cast<float4*>(p.out_raw + orow[i] + co) = a;
      if (p.out_act) {
        float4 w;
        if (p.act == ACT_SNAKE) {
          w.x = snake_sel<SFU>(a.x, al.x, ia.x); w.y = snake_sel<SFU>(a.y, al.y, ia.y);
          w.z = snake_sel<SFU>(a.z, al.z, ia.z); w.w = snake_sel<SFU>(a.w, al.w, ia.w);
        } else {
          w.x = apply_act(a.x, p.act, 0.f); w.y = apply_act(a.y, p.act, 0.f);
          w.z = apply_act(a.z, p.act, 0.f); w.w = apply_act(a.w, p.act, 0.f);
        }
        if (p.out_fmt == FMT_F32) {
          *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out_act) + orow[i] + co) = w;
        } else {
          const __nv_bfloat162 h01 = __floats2bfloat162_rn(w.x, w.y), h23 = __floats2bfloat162_rn(w.z, w.w);
          __nv_bfloat16* oh = reinterpret_cast<__nv_bfloat16*>(p.out_act) + orow[i] + co;
          *reinterpret_cast<uint2*>(oh) = make_uint2(*reinterpret_cast<const uint32_t*>(&h01),
                                                     *reinterpret_cast<const uint32_t*>(&h23));
          if (p.out_fmt == FMT_PLANES) {
            const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
            const __nv_bfloat162 l01 = __floats2bfloat162_rn(w.x - f01.x, w.y - f01.y);
            const __nv_bfloat162 l23 = __floats2bfloat162_rn(w.z - f23.x, w.w - f23.y);
            *reinterpret_cast<uint2*>(oh + p.act_plane_elems) =
                make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
          }
        }
      }
    }
    if (!lock) epi_group_sync(g);   // the group's staging tile is rewritten by the next chunk (lock mode: the acquire orders it)
  }
}

// Nearest-code epilogue of one 128-row x BN-code score tile (16 warps: 4 TMEM lane quadrants x 4 column
// quarters; thread = row).  Each thread scans its quarter of the codes keeping (best, first index of best,
// runner-up); the four quarters of a row are merged through shared memory.  The [N, K] score matrix never
// leaves TMEM.
__device__ __forceinline__ void tc_epilogue_argmax(const TcConvParams& p, float* stg, uint32_t tcount, uint32_t t_acc,
                                                   int jt, int nt, uint32_t tempty_bar, int warp, int lane) {
  const int quad = warp & 3;
  const int part = (warp - 2) >> 2;
  const int et = threadIdx.x - 64;
  const int row = quad * 32 + lane;
  const int cpp = p.BN >> 2;                       // codes per quarter (multiple of 16)
  const int col0 = nt * p.BN + part * cpp;
  const uint32_t t_src = t_acc + ((uint32_t)(quad * 32) << 16) + part * cpp;
  // 0.5|e|^2 of this tile's codes -> shared memory once (a global load per column stalled every compare: with
  // ~210 KB of dynamic shared memory the L1 is a few KB and each load went to L2)
  float* hn_s = stg + 4096 + (tcount & 1u) * 256;      // after the two 8 KB merge tiles
  if (et < p.BN) {
    const int col = nt * p.BN + et;
    hn_s[et] = col < p.n_codes ? __ldg(p.half_norm + col) : 0.f;
  }
  asm volatile("bar.sync 1, 512;" ::: "memory");
  float best = -INFINITY, second = -INFINITY;
  int bidx = 0x7fffffff;
  const int ncol = min(cpp, p.n_codes - col0);           // valid codes of this thread's quarter
  for (int c = 0; c < cpp; c += 16) {
    float v[16];
    tmem_ld16(t_src + c, v);
    if (c + 16 <= ncol) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float sc = __fsub_rn(v[i], hn_s[part * cpp + c + i]);
        second = fmaxf(second, fminf(best, sc));         // runner-up = max over all but the (first) best
        if (sc > best) { best = sc; bidx = col0 + c + i; }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if (c + i < ncol) {
          const float sc = __fsub_rn(v[i], hn_s[part * cpp + c + i]);
          second = fmaxf(second, fminf(best, sc));
          if (sc > best) { best = sc; bidx = col0 + c + i; }
        }
      }
    }
  }
  float4* sm = reinterpret_cast<float4*>(stg) + (tcount & 1u) * (4 * TC_BM);
  sm[part * TC_BM + row] = make_float4(best, __int_as_float(bidx), second, 0.f);
  tc_fence_before();
  asm volatile("bar.sync 1, 512;" ::: "memory");
  if (et == 0) mbar_arrive(tempty_bar);
  if (part == 0) {
    const int grow = jt * TC_BM + row;
#pragma unroll
    for (int q = 1; q < 4; ++q) {
      const float4 o = sm[q * TC_BM + row];
      const int oi = __float_as_int(o.y);
      if (o.x > best || (o.x == best && oi < bidx)) {
        second = fmaxf(best, o.z);
        best = o.x; bidx = oi;
      } else {
        second = fmaxf(second, o.x);
      }
    }
    if (grow < p.n_rows) p.cand[(size_t)grow * p.n_ntiles + nt] = make_float4(best, __int_as_float(bidx), second, 0.f);
  }
}

template <int X3, int EPI>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
               const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
               const __grid_constant__ TcConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[TC_MAX_STAGES];
  __shared__ __align__(8) uint64_t bar_empty[TC_MAX_STAGES];
  __shared__ __align__(8) uint64_t bar_tfull[2];
  __shared__ __align__(8) uint64_t bar_tempty[2];
  __shared__ uint32_t tmem_base_s;

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;  // provably warp-uniform: ptxas keeps the role code on the uniform datapath
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sub_bytes = (p.a_bytes + p.b_bytes) * (X3 ? 2u : 1u);   // one K block: A (+lo) | B (+lo)
  const uint32_t stage_bytes = sub_bytes * (uint32_t)p.kgroup;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA_hi);
    tma_prefetch_desc(&tmB_hi);
    if (X3) { tma_prefetch_desc(&tmA_lo); tma_prefetch_desc(&tmB_lo); }
    for (int s = 0; s < p.stages; ++s) { mbar_init(smem_u32(&bar_full[s]), 1); mbar_init(smem_u32(&bar_empty[s]), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(smem_u32(&bar_tfull[a]), 1); mbar_init(smem_u32(&bar_tempty[a]), 1); }
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
  const int n_kiter = p.KT * p.n_kblk;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      Ring rg;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int nt = tile % p.n_ntiles;
        int mt = tile / p.n_ntiles;
        const int jt = mt % p.tiles_j;
        mt /= p.tiles_j;
        const int ph = mt % p.n_phase;
        const int b = (mt / p.n_phase) * (p.JB ? p.BB : 1);        // first frame of the tile
        const int j0 = jt * (p.JB ? p.JB : TC_BM);
        int tap = 0, cb = 0;
        for (int k0 = 0; k0 < n_kiter; k0 += p.kgroup, rg.next(p.stages)) {
          const int cnt = min(p.kgroup, n_kiter - k0);
          const uint32_t s = rg.s, par = rg.par;
          mbar_wait(smem_u32(&bar_empty[s]), par ^ 1u, 1);
          const uint32_t full = smem_u32(&bar_full[s]);
          if (p.dbg & 2) { mbar_arrive(full); continue; }
          mbar_expect_tx(full, sub_bytes * (uint32_t)cnt);
          for (int g = 0; g < cnt; ++g) {
            const uint32_t sa = smem0 + s * stage_bytes + g * sub_bytes;
            const uint32_t sb = sa + p.a_bytes * (X3 ? 2u : 1u);
            const int c0 = cb * p.BK;
            const int brow = (ph * p.KT + tap) * p.Cout + nt * p.BN;
            if (p.in_step == 1) {
              const int row = j0 + tap * p.dil + p.in_off[ph];
              tma_load_3d(sa, &tmA_hi, full, c0, row, b);
              if (X3) tma_load_3d(sa + p.a_bytes, &tmA_lo, full, c0, row, b);
            } else {
              const int kk = tap + p.in_off[ph];                 // dil == 1 for strided convs
              int q = kk / p.in_step, r = kk - q * p.in_step;
              if (r < 0) { r += p.in_step; q -= 1; }
              tma_load_4d(sa, &tmA_hi, full, c0, r, j0 + q, b);
              if (X3) tma_load_4d(sa + p.a_bytes, &tmA_lo, full, c0, r, j0 + q, b);
            }
            tma_load_2d(sb, &tmB_hi, full, c0, brow);
            if (X3) tma_load_2d(sb + p.b_bytes, &tmB_lo, full, c0, brow);
            if (++cb == p.n_kblk) { cb = 0; ++tap; }   // tap outer, channels ascending: the accumulation order does
          }                                            // not depend on BK, so a frame's bits do not depend on the batch
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp runs the (warp-uniform) loop; elect.sync inside the MMA / commit wrappers picks the lane
    // that issues.  K-steps are unrolled at compile time and descriptors advance by 32-bit adds, so the issue
    // loop costs a few instructions per tcgen05.mma (it was the bottleneck of the narrow layers).
    {
      const uint32_t idesc = umma_idesc_bf16(TC_BM, p.BN);
      const int ksteps = p.BK / 16;
      const uint32_t desc_hi = (uint32_t)(umma_desc_base(p.sbo, p.layout_type) >> 32);
      const uint32_t a_plane = p.a_bytes >> 4, b_plane = p.b_bytes >> 4;
      uint32_t tcount = 0;
      Ring rg;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++tcount) {
        const uint32_t acc = tcount & 1u, apar = (tcount >> 1) & 1u;
        mbar_wait_warp(smem_u32(&bar_tempty[acc]), apar ^ 1u, 2, lane);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * p.acc_stride;
        for (int k0 = 0; k0 < n_kiter; k0 += p.kgroup, rg.next(p.stages)) {
          const int cnt = min(p.kgroup, n_kiter - k0);
          const uint32_t s = rg.s, par = rg.par;
          mbar_wait_warp(smem_u32(&bar_full[s]), par, 3, lane);
          if (!(p.dbg & 16)) tc_fence_after();
          for (int g = 0; g < cnt; ++g) {
            const uint32_t sa = smem0 + s * stage_bytes + g * sub_bytes;
            const uint32_t a_lo = (sa & 0x3FFFFu) >> 4;
            const uint32_t b_lo = ((sa + p.a_bytes * (X3 ? 2u : 1u)) & 0x3FFFFu) >> 4;
            const uint32_t first = (k0 + g) != 0;
            if (p.dbg & 4) {}
            else if (ksteps == 4) umma_ksteps<X3, 4>(d_tmem, a_lo, b_lo, a_plane, b_plane, desc_hi, idesc, first);
            else if (ksteps == 2) umma_ksteps<X3, 2>(d_tmem, a_lo, b_lo, a_plane, b_plane, desc_hi, idesc, first);
            else umma_ksteps<X3, 1>(d_tmem, a_lo, b_lo, a_plane, b_plane, desc_hi, idesc, first);
          }
          if (p.dbg & 8) { if (lane == 0) mbar_arrive(smem_u32(&bar_empty[s])); __syncwarp(); }
          else umma_commit_w(smem_u32(&bar_empty[s]));   // frees the smem stage when these MMAs retire
        }
        umma_commit_w(smem_u32(&bar_tfull[acc]));   // accumulator complete -> epilogue
      }
    }
  } else {
    // ===================== epilogue: 16 warps =====================
    // Per 32-column chunk of the accumulator: (1) every warp copies its TMEM quadrant (thread = row, 8
    // columns) into a padded fp32 staging tile in shared memory; (2) after a named barrier the 512 threads
    // re-read the tile row-major (8 lanes x float4 = one 128-byte row segment), so the bias / residual loads
    // and the raw / activated stores are coalesced.  Two staging tiles alternate: one barrier per chunk.
    float* stg = reinterpret_cast<float*>(smem_raw + (smem0 - smem_u32(smem_raw)) + (size_t)p.stages * stage_bytes);
    if (EPI != 1 && p.epi_groups == 2) {
      // two independent groups of 8 warps: group g drains TMEM buffer g (tiles g, g + 2, ... of this CTA)
      const int g = (warp - 2) >> 3;
      float* stg_g = stg + g * (TC_BM * TC_STG_LD);
      auto coords = [&](int tile, int& b, int& ph, int& jt, int& nt) {
        nt = tile % p.n_ntiles;
        int mt = tile / p.n_ntiles;
        jt = mt % p.tiles_j;
        mt /= p.tiles_j;
        ph = mt % p.n_phase;
        b = mt / p.n_phase;
      };
      uint32_t use = 0;   // how many times this group's TMEM buffer has been used
      for (int tile = blockIdx.x + g * gridDim.x; tile < p.total_tiles; tile += 2 * gridDim.x, ++use) {
        int b, ph, jt, nt;
        coords(tile, b, ph, jt, nt);
        if (use == 0) tc_prefetch_res_g(p, b, ph, jt, nt);
        if (tile + 2 * (int)gridDim.x < p.total_tiles) {
          int b2, ph2, jt2, nt2;
          coords(tile + 2 * gridDim.x, b2, ph2, jt2, nt2);
          tc_prefetch_res_g(p, b2, ph2, jt2, nt2);
        }
        mbar_wait(smem_u32(&bar_tfull[g]), use & 1u, 4);
        tc_fence_after();
        if (p.dbg & 1) {
          tc_fence_before();
          epi_group_sync(g);
          if (((threadIdx.x - 64) & 255) == 0) mbar_arrive(smem_u32(&bar_tempty[g]));
        } else {
          tc_epilogue_tile_g<!X3, EPI == 2>(p, stg_g, g, tmem_base + g * p.acc_stride, b, ph, jt, nt, smem_u32(&bar_tempty[g]), warp, lane);
        }
      }
    } else {
      uint32_t tcount = 0, chunk_ctr = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++tcount) {
        const int nt = tile % p.n_ntiles;
        int mt = tile / p.n_ntiles;
        const int jt = mt % p.tiles_j;
        mt /= p.tiles_j;
        const int ph = mt % p.n_phase;
        const int b = mt / p.n_phase;
        const uint32_t acc = tcount & 1u, apar = (tcount >> 1) & 1u;
        if (tcount == 0) tc_prefetch_res(p, b, ph, jt, nt);
        {
          const int nx = tile + gridDim.x;
          if (nx < p.total_tiles) {
            const int nt2 = nx % p.n_ntiles;
            int m2 = nx / p.n_ntiles;
            const int jt2 = m2 % p.tiles_j;
            m2 /= p.tiles_j;
            tc_prefetch_res(p, m2 / p.n_phase, m2 % p.n_phase, jt2, nt2);
          }
        }
        mbar_wait_epilogue(smem_u32(&bar_tfull[acc]), apar, 4);
        tc_fence_after();
        if (p.dbg & 1) {
          tc_fence_before();
          asm volatile("bar.sync 1, 512;" ::: "memory");
          if (threadIdx.x == 64) mbar_arrive(smem_u32(&bar_tempty[acc]));
        } else if (EPI != 1)
          tc_epilogue_tile<!X3, EPI == 2>(p, stg, chunk_ctr, tmem_base + acc * p.acc_stride, b, ph, jt, nt, smem_u32(&bar_tempty[acc]),
                           warp, lane);
        else
          tc_epilogue_argmax(p, stg, tcount, tmem_base + acc * p.acc_stride, jt, nt, smem_u32(&bar_tempty[acc]), warp, lane);
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
// Slab kernel (stride-1 convs, ConvTranspose1d phases, linears).  The generic kernel above re-reads the
// activation tile once per tap and the weight tile once per 128 positions, which makes the long, narrow
// layers L2-bandwidth bound.  Here a work item is MT consecutive 128-position tiles x one BN channel
// block:
//   * per input-channel block ONE activation slab of MT*128 + (KT-1)*dil rows is loaded; tap t of m-tile m
//     is the same shared memory at a row offset (m*128 + t*dil): only the UMMA descriptor start address
//     moves (the 128B/64B swizzle is a function of the absolute smem address, which TMA used when writing);
//   * every (tap, channel block) weight tile is loaded once and feeds the MT accumulators (MT*BN TMEM
//     columns, double buffered).
// Two rings: activation slabs (SA slots) and weight tiles (SB slots).
// ---------------------------------------------------------------------------------------------
constexpr int TC2_MAX_SA = 4;
constexpr int TC2_MAX_SB = 8;

template <int X3>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                const __grid_constant__ TcConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_afull[TC2_MAX_SA];
  __shared__ __align__(8) uint64_t bar_aempty[TC2_MAX_SA];
  __shared__ __align__(8) uint64_t bar_bfull[TC2_MAX_SB];
  __shared__ __align__(8) uint64_t bar_bempty[TC2_MAX_SB];
  __shared__ __align__(8) uint64_t bar_tfull[2];
  __shared__ __align__(8) uint64_t bar_tempty[2];
  __shared__ uint32_t tmem_base_s;

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;  // provably warp-uniform: ptxas keeps the role code on the uniform datapath
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t planes = X3 ? 2u : 1u;
  const uint32_t a_slot = p.a_plane_bytes * planes;
  const uint32_t b_slot = p.b_bytes * planes;
  const uint32_t smemB = smem0 + p.SA * a_slot;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA_hi);
    tma_prefetch_desc(&tmB_hi);
    if (X3) { tma_prefetch_desc(&tmA_lo); tma_prefetch_desc(&tmB_lo); }
    for (int i = 0; i < p.SA; ++i) { mbar_init(smem_u32(&bar_afull[i]), 1); mbar_init(smem_u32(&bar_aempty[i]), 1); }
    for (int i = 0; i < p.SB; ++i) { mbar_init(smem_u32(&bar_bfull[i]), 1); mbar_init(smem_u32(&bar_bempty[i]), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&bar_tfull[i]), 1); mbar_init(smem_u32(&bar_tempty[i]), 1); }
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
  const uint32_t set_cols = (uint32_t)(p.MT * p.acc_stride);

  if (warp == 0) {
    // ===================== TMA producer =====================
    // Steps = (work item, input-channel block) in order.  The slab of step s+1 is requested BEFORE the weight
    // tiles of step s, so the (large, possibly DRAM-resident) activation slab has a whole step of lead time
    // instead of the few weight tiles the B ring would otherwise allow.
    if (lane == 0) {
      Ring ra, rb;
      auto issue_slab = [&](int item, int cb) {
        int r = item / p.n_ntiles;
        const int jg = r % p.groups_j;
        r /= p.groups_j;
        const int ph = r % p.n_phase;
        const int b = r / p.n_phase;
        const int row0 = jg * p.MT * TC_BM + p.in_off[ph];
        const uint32_t sa = ra.s, par = ra.par;
        mbar_wait(smem_u32(&bar_aempty[sa]), par ^ 1u, 1);
        const uint32_t full = smem_u32(&bar_afull[sa]);
        mbar_expect_tx(full, a_slot);
        const uint32_t dst = smem0 + sa * a_slot;
        for (int ld = 0; ld < p.n_aloads; ++ld) {
          const uint32_t off = (uint32_t)(ld * p.box_rows) * p.row_bytes;
          tma_load_3d(dst + off, &tmA_hi, full, cb * p.BK, row0 + ld * p.box_rows, b);
          if (X3) tma_load_3d(dst + p.a_plane_bytes + off, &tmA_lo, full, cb * p.BK, row0 + ld * p.box_rows, b);
        }
        ra.next(p.SA);
      };
      if ((int)blockIdx.x < p.total_tiles) issue_slab(blockIdx.x, 0);
      for (int item = blockIdx.x; item < p.total_tiles; item += gridDim.x) {
        const int nt = item % p.n_ntiles;
        const int ph = (item / (p.n_ntiles * p.groups_j)) % p.n_phase;
        for (int cb = 0; cb < p.n_kblk; ++cb) {
          if (cb + 1 < p.n_kblk) issue_slab(item, cb + 1);
          else if (item + (int)gridDim.x < p.total_tiles) issue_slab(item + gridDim.x, 0);
          const int c0 = cb * p.BK;
          for (int tap = 0; tap < p.KT; ++tap, rb.next(p.SB)) {
            const uint32_t sb = rb.s, par = rb.par;
            mbar_wait(smem_u32(&bar_bempty[sb]), par ^ 1u, 2);
            const uint32_t full = smem_u32(&bar_bfull[sb]);
            mbar_expect_tx(full, b_slot);
            const int brow = (ph * p.KT + tap) * p.Cout + nt * p.BN;
            const uint32_t dst = smemB + sb * b_slot;
            tma_load_2d(dst, &tmB_hi, full, c0, brow);
            if (X3) tma_load_2d(dst + p.b_bytes, &tmB_lo, full, c0, brow);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform loop, elect.sync issues) =====================
    {
      const uint32_t idesc = umma_idesc_bf16(TC_BM, p.BN);
      const int ksteps = p.BK / 16;
      const uint32_t desc_hi = (uint32_t)(umma_desc_base(p.sbo, p.layout_type) >> 32);
      const uint32_t a_plane = p.a_plane_bytes >> 4, b_plane = p.b_bytes >> 4;
      const uint32_t tap_step = ((uint32_t)p.dil * p.row_bytes) >> 4;        // descriptor units per tap
      const uint32_t mtile_step = ((uint32_t)TC_BM * p.row_bytes) >> 4;      // descriptor units per m-tile
      uint32_t tcount = 0;
      Ring ra, rb;
      for (int item = blockIdx.x; item < p.total_tiles; item += gridDim.x, ++tcount) {
        const int jg = (item / p.n_ntiles) % p.groups_j;
        const int n_m = min(p.MT, p.tiles_j - jg * p.MT);
        const uint32_t acc = tcount & 1u, apar = (tcount >> 1) & 1u;
        mbar_wait_warp(smem_u32(&bar_tempty[acc]), apar ^ 1u, 3, lane);
        tc_fence_after();
        const uint32_t d_set = tmem_base + acc * set_cols;
        for (int cb = 0; cb < p.n_kblk; ++cb, ra.next(p.SA)) {
          const uint32_t sa = ra.s, apr = ra.par;
          mbar_wait_warp(smem_u32(&bar_afull[sa]), apr, 4, lane);
          tc_fence_after();
          uint32_t a_tap = ((smem0 + sa * a_slot) & 0x3FFFFu) >> 4;
          for (int tap = 0; tap < p.KT; ++tap, rb.next(p.SB), a_tap += tap_step) {
            const uint32_t sb = rb.s, bpr = rb.par;
            mbar_wait_warp(smem_u32(&bar_bfull[sb]), bpr, 5, lane);
            tc_fence_after();
            const uint32_t b_lo = ((smemB + sb * b_slot) & 0x3FFFFu) >> 4;
            const uint32_t first = (cb | tap) != 0;
            uint32_t a_lo = a_tap, d_tmem = d_set;
            for (int m = 0; m < n_m; ++m, a_lo += mtile_step, d_tmem += p.acc_stride) {
              if (ksteps == 4) umma_ksteps<X3, 4>(d_tmem, a_lo, b_lo, a_plane, b_plane, desc_hi, idesc, first);
              else if (ksteps == 2) umma_ksteps<X3, 2>(d_tmem, a_lo, b_lo, a_plane, b_plane, desc_hi, idesc, first);
              else umma_ksteps<X3, 1>(d_tmem, a_lo, b_lo, a_plane, b_plane, desc_hi, idesc, first);
            }
            umma_commit_w(smem_u32(&bar_bempty[sb]));
          }
          umma_commit_w(smem_u32(&bar_aempty[sa]));
        }
        umma_commit_w(smem_u32(&bar_tfull[acc]));
      }
    }
  } else {
    // ===================== epilogue =====================
    float* stg = reinterpret_cast<float*>(smem_raw + (smem0 - smem_u32(smem_raw)) + (size_t)p.SA * a_slot +
                                          (size_t)p.SB * b_slot);
    uint32_t tcount = 0, chunk_ctr = 0;
    for (int item = blockIdx.x; item < p.total_tiles; item += gridDim.x, ++tcount) {
      const int nt = item % p.n_ntiles;
      int r = item / p.n_ntiles;
      const int jg = r % p.groups_j;
      r /= p.groups_j;
      const int ph = r % p.n_phase;
      const int b = r / p.n_phase;
      const int n_m = min(p.MT, p.tiles_j - jg * p.MT);
      const uint32_t acc = tcount & 1u, apar = (tcount >> 1) & 1u;
      mbar_wait_epilogue(smem_u32(&bar_tfull[acc]), apar, 6);
      tc_fence_after();
      for (int m = 0; m < n_m; ++m)
        tc_epilogue_tile<!X3>(p, stg, chunk_ctr, tmem_base + acc * set_cols + m * p.acc_stride, b, ph, jg * p.MT + m, nt,
                         m == n_m - 1 ? smem_u32(&bar_tempty[acc]) : 0u, warp, lane);
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

// fp32 [n] -> bf16 hi plane [n] | lo plane [n]
__global__ void split_planes_f32(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi,
                                 __nv_bfloat16* __restrict__ lo, size_t n) {
  size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    float4 v = *reinterpret_cast<const float4*>(x + i);
    __nv_bfloat16 h[4], l[4];
    split_bf16(v.x, h[0], l[0]); split_bf16(v.y, h[1], l[1]); split_bf16(v.z, h[2], l[2]); split_bf16(v.w, h[3], l[3]);
    *reinterpret_cast<uint2*>(hi + i) = make_uint2(pack_bf16(h[0], h[1]), pack_bf16(h[2], h[3]));
    if (lo) *reinterpret_cast<uint2*>(lo + i) = make_uint2(pack_bf16(l[0], l[1]), pack_bf16(l[2], l[3]));
  } else {
    for (; i < n; ++i) {
      __nv_bfloat16 h, l;
      split_bf16(x[i], h, l);
      hi[i] = h;
      if (lo) lo[i] = l;
    }
  }
}

// bf16 hi plane (+ lo plane) -> fp32
__global__ void merge_planes_f32(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo,
                                 float* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = __bfloat162float(hi[i]);
  if (lo) v += __bfloat162float(lo[i]);
  out[i] = v;
}


// Launch with the programmatic-serialization attribute (see pdl_sync()); B2C_PDL=0 launches normally.
inline bool tc_pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B2C_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}
template <typename... KArgs, typename... Args>
inline cudaError_t tc_launch(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = tc_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct TcWeight {
  __nv_bfloat16* hi = nullptr;   // [n_phase*KT*Cout][Cin], K-major
  __nv_bfloat16* lo = nullptr;
  int rows = 0, cin = 0;
};

inline void tc_weight_free(TcWeight& w) {
  if (w.hi) cudaFree(w.hi);
  if (w.lo) cudaFree(w.lo);
  w.hi = w.lo = nullptr;
}

inline uint16_t f32_to_bf16_rn_host(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7F800000u) == 0x7F800000u) return (uint16_t)(u >> 16);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
inline float bf16_to_f32_host(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

// packed: fp32 [n_phase][KT][Cin][Cout] (the FP32 kernel's layout) -> bf16 hi/lo [n_phase][KT][Cout][Cin]
inline int tc_weight_pack(const float* packed, int n_phase, int kt, int cin, int cout, TcWeight* out, size_t* bytes) {
  if (cin % 32 != 0 || cout % 32 != 0) return 0;   // not a tensor-core layer (stem, head, tiny projections)
  size_t n = (size_t)n_phase * kt * cin * cout;
  std::vector<uint16_t> hi(n), lo(n);
  for (int pt = 0; pt < n_phase * kt; ++pt)
    for (int ci = 0; ci < cin; ++ci)
      for (int co = 0; co < cout; ++co) {
        float w = packed[((size_t)pt * cin + ci) * cout + co];
        uint16_t h = f32_to_bf16_rn_host(w);
        uint16_t l = f32_to_bf16_rn_host(w - bf16_to_f32_host(h));
        size_t o = ((size_t)pt * cout + co) * cin + ci;
        hi[o] = h;
        lo[o] = l;
      }
  if (cudaMalloc((void**)&out->hi, n * 2) != cudaSuccess) return -1;
  if (cudaMalloc((void**)&out->lo, n * 2) != cudaSuccess) return -1;
  if (cudaMemcpy(out->hi, hi.data(), n * 2, cudaMemcpyHostToDevice) != cudaSuccess) return -1;
  if (cudaMemcpy(out->lo, lo.data(), n * 2, cudaMemcpyHostToDevice) != cudaSuccess) return -1;
  out->rows = n_phase * kt * cout;
  out->cin = cin;
  if (bytes) *bytes += n * 4;
  return 0;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled tc_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

struct TcConvPlan {
  TcConvParams p;
  int slab = 0;   // 1: conv_tc2_kernel (activation slab + MT m-tiles), 0: conv_tc_kernel
  int x3 = 0;
  int in_fmt = FMT_PLANES;
  int grid = 0;
  size_t smem = 0;
  // tensor-map cache (re-encoded when the activation pointer changes)
  const void* cached_x = nullptr;
  CUtensorMap mA_hi, mA_lo, mB_hi, mB_lo;
  bool b_ready = false;
};

// B2C_TC_SLAB=0 in the environment keeps every layer on the generic kernel (a debugging / A-B switch)
inline bool tc_slab_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B2C_TC_SLAB");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

// B2C_TC_SLAB=2 forces the slab kernel on every stride-1 layer it can take
inline bool tc_slab_forced() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B2C_TC_SLAB");
    v = (e && e[0] == '2') ? 1 : 0;
  }
  return v == 1;
}

inline int largest_bn(int cout) {
  for (int bn = 256; bn >= 32; bn -= 32)
    if (cout % bn == 0) return bn;
  return 0;
}

inline bool tc_conv_eligible(const TcWeight& w, int cin, int cout, int stride, int dilation, int Lin) {
  if (!w.hi || cin % 32 != 0 || !largest_bn(cout)) return false;
  if (stride > 1 && (dilation != 1 || Lin % stride != 0)) return false;
  return true;
}

// a narrower channel block is taken when its modelled cost is below this percentage of the wider one's (B2C_TC_BNWAVE_PCT)
inline long tc_bnwave_pct() {
  static long v = -1;
  if (v < 0) {
    const char* e = getenv("B2C_TC_BNWAVE_PCT");
    v = e ? atol(e) : 90;
    if (v < 1 || v > 100) v = 90;
  }
  return v;
}

// fewest M tiles of the generic kernel for Lj positions x B frames: 128 positions of one frame per tile (jb = 0) or
// frame-packed tiles of jb positions x 128 / jb frames (TcConvParams::JB)
inline long tc_best_pack(int Lj, int B, int* jb_out) {
  long best = (long)B * ((Lj + TC_BM - 1) / TC_BM);
  *jb_out = 0;
  const char* e = getenv("B2C_TC_PACK");
  if (B < 2 || (e && e[0] == '0')) return best;
  for (int sh = 6; sh >= 0; --sh) {
    const int jb = 1 << sh, bb = TC_BM >> sh;
    const long t = (long)((Lj + jb - 1) / jb) * ((B + bb - 1) / bb);
    if (t < best) { best = t; *jb_out = jb; }
  }
  return best;
}

// returns 0 = planned; >0 = shape not eligible; <0 = error
inline int tc_conv_plan(const ConvArgs& a, const TcWeight& w, int precision, int out_fmt, int sm_count, TcConvPlan* plan) {
  if (!w.hi) return 1;
  if (a.Cin % 32 != 0) return 2;
  if (a.in_step > 1 && (a.dil != 1 || a.Lin % a.in_step != 0)) return 3;
  int bn = largest_bn(a.Cout);
  if (!bn) return 4;
  // Channel-block width against wave quantisation: a layer with few tiles (short sequences, batch-1 streaming) runs
  // ceil(tiles / SMs) rounds of tile time, so a narrower block can win although its MMAs are less efficient
  // (75 positions x 64 frames, 1024 channels: 152 tiles of N = 256 take 2 rounds of 128 cycles per K step, 304 tiles of
  // N = 128 take 3 rounds of 64).  Per K step an MMA costs max(N / 2, 32 + N / 4) cycles (tools/micro/umma_rate.cu);
  // ~1500 cycles per tile for hand-offs and the epilogue's tail.  The bits of an output do not depend on N.
  {
    int jb_unused = 0;
    const long mtiles = tc_best_pack(a.Lj, a.B, &jb_unused) * a.n_phase;
    const long ksteps = (long)a.KT * (a.Cin / 16) * (precision == 1 ? 3 : 1);
    long best_cost = -1;
    int best_bn = bn;
    for (int c = bn; c >= 64; c /= 2) {
      if (c % 32 != 0 || a.Cout % c != 0) break;
      const long tiles = mtiles * (a.Cout / c);
      const long waves = (tiles + sm_count - 1) / sm_count;
      const long cyc = c / 2 > 32 + c / 4 ? c / 2 : 32 + c / 4;
      const long cost = waves * (ksteps * cyc + 1500);
      // a narrower block must pay clearly (>= 10 %): it re-reads every activation tile once more per halving
      if (best_cost < 0 || cost * 100 < best_cost * tc_bnwave_pct()) { best_cost = cost; best_bn = c; }
    }
    const char* e = getenv("B2C_TC_BNWAVE");
    if (e && e[0] == '0') {                       // the round-1 rule: narrower blocks only while the GPU is not full
      const long mt = (long)a.B * a.n_phase * ((a.Lj + TC_BM - 1) / TC_BM);
      while (bn > 64 && (bn / 2) % 32 == 0 && a.Cout % (bn / 2) == 0 && mt * (a.Cout / bn) * 2 <= sm_count) bn /= 2;
    } else {
      bn = best_bn;
    }
  }
  if (!tc_encode_fn()) return -10;
  TcConvParams& p = plan->p;
  memset(&p, 0, sizeof(p));
  plan->x3 = precision == 1 ? 1 : 0;
  plan->in_fmt = plan->x3 ? FMT_PLANES : FMT_HI;
  p.B = a.B; p.Lin = a.Lin; p.Cin = a.Cin; p.Cout = a.Cout; p.KT = a.KT; p.in_step = a.in_step; p.dil = a.dil;
  p.n_phase = a.n_phase; p.Lj = a.Lj; p.out_step = a.out_step; p.Lout = a.Lout;
  for (int i = 0; i < 8; ++i) { p.in_off[i] = a.in_off[i]; p.out_off[i] = a.out_off[i]; }
  p.act = a.act; p.res_mode = a.res_mode; p.Tl = a.Tl; p.chunk = a.chunk; p.out_fmt = out_fmt;
  {
    const char* e = getenv("B2C_TC_DEBUG");
    p.dbg = e ? atoi(e) : 0;
  }
  p.act_plane_elems = (long)a.B * a.Lout * a.Cout;
  p.BN = bn;
  // shared memory: pipeline stages + the epilogue's two staging tiles + 1 KB of alignment slack, under the
  // 227 KB per-CTA limit (the kernel's static shared memory is ~1.3 KB).  Prefer 128-byte K slices (BK = 64)
  // when at least 3 stages fit, else 64-byte slices (BK = 32).
  const int budget = 232448 - 2048 - TC_STG_BYTES - 1024;        // with two staging tiles
  const int budget1 = budget + TC_STG_BYTES / 2;                  // with one
  uint32_t stage = 0;
  int stages = 0;
  p.stg_bufs = 2;
  {
    // Two independent epilogue groups (one per TMEM buffer) overlap the load and store phases of consecutive
    // tiles -- a gain where the epilogue dominates (k = 1 convs, up-convs: -12..-16 %), a loss where a tile's MMAs
    // are long, because a 256-thread group takes twice as long to release its buffer (128-wide bf16x3 k = 7:
    // +16 %).  Rule from the B200 A/B run: groups when the MMAs of a tile take <= ~1000 tensor-pipe cycles per
    // 32-column chunk of epilogue work.
    const long mma_cyc = (long)a.KT * (a.Cin / 16) * (plan->x3 ? 3 : 1) * (bn / 2);
    p.epi_groups = mma_cyc <= 1000L * (bn / 32) ? 2 : 1;
    const char* e = getenv("B2C_TC_EPI2");
    if (e) p.epi_groups = e[0] == '0' ? 1 : 2;
  }
  for (int bk = (a.Cin % 64 == 0) ? 64 : 32; bk >= 32; bk -= 32) {
    p.BK = bk;
    p.a_bytes = TC_BM * bk * 2;
    p.b_bytes = bn * bk * 2;
    stage = (p.a_bytes + p.b_bytes) * (plan->x3 ? 2 : 1);
    stages = budget / (int)stage;
    // give up the second staging tile when that buys one more K block in flight
    if (p.epi_groups == 1 && stages < 4 && budget1 / (int)stage > stages) { stages = budget1 / (int)stage; p.stg_bufs = 1; }
    else p.stg_bufs = 2;
    if (stages >= 3 || bk == 32) break;
  }
  // K blocks per stage: every mbarrier hand-off costs ~400 cycles on both the TMA and the MMA warp (measured
  // with the MMAs and copies switched off).  K blocks whose MMAs take less than ~512 tensor-pipe cycles are
  // grouped under one hand-off while at least two stages remain; longer blocks lose more from the shallower
  // ring than they gain (measured: grouping helps 64..384-channel layers, hurts 256-wide bf16x3 and 768 bf16).
  p.kgroup = 1;
  {
    const int cyc = (p.BK / 16) * (plan->x3 ? 3 : 1) * (bn / 2);   // tensor-pipe cycles of one K block (M = 128)
    int want = (512 + cyc - 1) / cyc;
    const char* e = getenv("B2C_TC_KGROUP");
    if (e) want = atoi(e);
    if (want < 1) want = 1;
    if (want > 8) want = 8;
    while (want > 1 && stages / want < 2) --want;
    while (want > 1 && a.KT * (a.Cin / p.BK) < 2 * want) --want;
    p.kgroup = want;
    stages /= want;
    stage *= want;
  }
  p.n_kblk = a.Cin / p.BK;
  p.sbo = 8 * p.BK * 2;
  p.layout_type = p.BK == 64 ? 2u : 4u;
  if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
  if (stages < 2) return 5;
  p.stages = stages;
  p.tiles_j = (a.Lj + TC_BM - 1) / TC_BM;
  p.n_ntiles = a.Cout / bn;
  long total = (long)a.B * a.n_phase * p.tiles_j * p.n_ntiles;
  if (total > 0x7fffffffL) return 6;
  p.total_tiles = (int)total;
  p.acc_stride = bn <= 64 ? 64 : (bn <= 128 ? 128 : 256);
  p.tmem_cols = 2 * p.acc_stride;
  plan->grid = (int)(total < sm_count ? total : sm_count);
  plan->smem = (size_t)stages * stage + (p.stg_bufs == 2 ? TC_STG_BYTES : TC_STG_BYTES / 2) + 1024;
  plan->cached_x = nullptr;
  plan->b_ready = false;
  plan->slab = 0;
  // Measured on B200 (tools/tc_selftest.py, batch 32): the slab kernel wins where the generic kernel's MMA is
  // narrow (N <= 128) and the contraction is long enough to amortise the slab (k = 7 at 96 / 128 channels);
  // wider layers are MMA-issue bound and prefer the generic kernel's N = 192 / 256 instructions.
  // ... that was before K blocks were grouped per hand-off in the generic kernel; with grouping the generic
  // kernel is as fast or faster on every layer of this model (96 ch: 0.240 vs 0.263 ms, 128 ch: 0.258 vs 0.256),
  // so the slab kernel is kept for experiments (B2C_TC_SLAB=2) and not selected by default.
  // One exception, measured at micro-batch 64: single-pass bf16 k = 7 layers with wide channel blocks (the decoder's
  // 768- and 384-channel units) gain 8-14 % from the slab kernel when it keeps the generic kernel's N (one m-tile per
  // item): their A re-reads were a third of the L2 traffic of an L2-bound loop.
  // (decided from the layer alone -- never from the batch -- and with a fixed K slice, because the slab kernel
  // accumulates channel-block-major: a layer must use the same kernel and order at every batch size.)
  const bool slab_wide = !plan->x3 && a.KT >= 3 && largest_bn(a.Cout) >= 192;
  const bool slab_pays = slab_wide;
  if (a.in_step == 1 && !a.dmul && tc_slab_enabled() && (slab_pays || tc_slab_forced())) {
    // slab kernel: narrower channel blocks, several m-tiles per work item
    int bn2 = a.Cout % 128 == 0 ? 128 : (a.Cout % 96 == 0 ? 96 : (a.Cout % 64 == 0 ? 64 : bn));
    {
      const char* e = getenv("B2C_TC_SLAB_WIDE");   // keep the generic kernel's wide N, one m-tile per item
      if (slab_wide || (e && e[0] == '1')) bn2 = bn;
    }
    const int stride2 = bn2 <= 64 ? 64 : (bn2 <= 128 ? 128 : 256);
    int mt = 256 / stride2;
    if (mt > 4) mt = 4;
    while (mt > 1 && mt > p.tiles_j) mt >>= 1;
    const int planes = plan->x3 ? 2 : 1;
    for (; mt >= 1 && !plan->slab; mt >>= 1) {
      for (int bk = (a.Cin % 64 == 0) ? 64 : 32; bk >= (slab_wide ? ((a.Cin % 64 == 0) ? 64 : 32) : 32) && !plan->slab; bk -= 32) {
        int rows = mt * TC_BM + (a.KT - 1) * a.dil;
        const int n_loads = (rows + 255) / 256;
        const int q = 16 * n_loads;
        rows = (rows + q - 1) / q * q;
        const int box_rows = rows / n_loads;
        if (box_rows > 256) continue;
        const uint32_t a_plane = (uint32_t)rows * bk * 2;
        const uint32_t a_slot = a_plane * planes;
        const uint32_t b_bytes = (uint32_t)bn2 * bk * 2;
        const uint32_t b_slot = b_bytes * planes;
        // three slab slots (one in use, one landed, one in flight) when they leave room for >= 4 weight tiles
        int SA = 3;
        if ((long)budget - 3L * a_slot < 4L * b_slot) {
          if (mt > 1) continue;   // fewer m-tiles per item rather than a two-slot slab ring
          SA = 2;
        }
        const long rest = (long)budget - (long)SA * a_slot;
        if (rest < 3L * b_slot) continue;
        int SB = (int)(rest / b_slot);
        if (SB > TC2_MAX_SB) SB = TC2_MAX_SB;
        p.BN = bn2; p.BK = bk; p.n_kblk = a.Cin / bk;
        p.a_bytes = TC_BM * bk * 2; p.b_bytes = b_bytes;
        p.sbo = 8 * bk * 2; p.layout_type = bk == 64 ? 2u : 4u;
        p.n_ntiles = a.Cout / bn2;
        p.acc_stride = stride2;
        p.MT = mt; p.groups_j = (p.tiles_j + mt - 1) / mt;
        p.slab_rows = rows; p.box_rows = box_rows; p.n_aloads = n_loads; p.SA = SA; p.SB = SB;
        p.row_bytes = bk * 2; p.a_plane_bytes = a_plane; p.stg_bufs = 2;
        p.tmem_cols = 2 * mt * stride2;
        if (p.tmem_cols < 32) p.tmem_cols = 32;
        long items = (long)a.B * a.n_phase * p.groups_j * p.n_ntiles;
        if (items > 0x7fffffffL) continue;
        p.total_tiles = (int)items;
        p.stages = 0;
        plan->grid = (int)(items < sm_count ? items : sm_count);
        plan->smem = (size_t)SA * a_slot + (size_t)SB * b_slot + TC_STG_BYTES + 1024;
        plan->slab = 1;
      }
    }
  }
  // frame-packed tiles (TcConvParams::JB): the generic kernel, more than one frame, and fewer tiles than 128 positions of
  // one frame per tile give.  B2C_TC_PACK=0 off.
  if (!plan->slab) {
    int best_jb = 0;
    const long best = tc_best_pack(a.Lj, a.B, &best_jb);
    if (best_jb) {
      p.JB = best_jb; p.BB = TC_BM / best_jb;
      p.jb_shift = 0;
      while ((1 << p.jb_shift) < best_jb) ++p.jb_shift;
      p.tiles_j = (a.Lj + p.JB - 1) / p.JB;
      const long tot = best * a.n_phase * p.n_ntiles;
      p.total_tiles = (int)tot;
      plan->grid = (int)(tot < sm_count ? tot : sm_count);
    }
  }
  return 0;
}

inline int tc_encode(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_b,
                     const cuuint32_t* box, int bk) {
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = tc_encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims,
                              strides_b, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -(int)r - 100;
}

// [B, L, C] channel-last tensor (fp32 or bf16) as a 3-D map with a {32 channels, 128 rows, 1} box: 128-byte rows
// (fp32, SWIZZLE_128B) or 64-byte rows (bf16, SWIZZLE_64B).  Rows past L are clipped on store and zero-filled on load.
inline int tc_encode_rows32(CUtensorMap* m, const void* base, bool f32, int B, int L, int C) {
  const cuuint64_t es_b = f32 ? 4 : 2;
  cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)L, (cuuint64_t)B};
  cuuint64_t str[2] = {(cuuint64_t)C * es_b, (cuuint64_t)L * C * es_b};
  cuuint32_t box[3] = {32, 128, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = tc_encode_fn()(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
                              const_cast<void*>(base), dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              f32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -(int)r - 100;
}

// x: bf16 hi plane of the input activations [B, Lin, Cin]; the lo plane follows at + B*Lin*Cin elements.
inline int tc_conv_launch(TcConvPlan& plan, const ConvArgs& a, const float* inv_alpha, const void* x_planes,
                          void* out_act, const TcWeight& w, cudaStream_t st) {
  TcConvParams p = plan.p;
  p.bias = a.bias; p.res = a.res; p.out_raw = a.out_raw; p.out_act = out_act; p.alpha = a.alpha;
  p.inv_alpha = inv_alpha; p.dmul = a.dmul;
  if (plan.cached_x != x_planes) {
    const __nv_bfloat16* xh = reinterpret_cast<const __nv_bfloat16*>(x_planes);
    const __nv_bfloat16* xl = xh + (size_t)p.B * p.Lin * p.Cin;
    int rc;
    if (p.in_step == 1) {
      cuuint64_t dims[3] = {(cuuint64_t)p.Cin, (cuuint64_t)p.Lin, (cuuint64_t)p.B};
      cuuint64_t str[2] = {(cuuint64_t)p.Cin * 2, (cuuint64_t)p.Lin * p.Cin * 2};
      cuuint32_t box[3] = {(cuuint32_t)p.BK, (cuuint32_t)(plan.slab ? p.box_rows : (p.JB ? p.JB : TC_BM)), (cuuint32_t)(p.JB ? p.BB : 1)};
      rc = tc_encode(&plan.mA_hi, xh, 3, dims, str, box, p.BK);
      if (!rc) rc = tc_encode(&plan.mA_lo, plan.x3 ? xl : xh, 3, dims, str, box, p.BK);
    } else {
      const int S = p.in_step;
      cuuint64_t dims[4] = {(cuuint64_t)p.Cin, (cuuint64_t)S, (cuuint64_t)(p.Lin / S), (cuuint64_t)p.B};
      cuuint64_t str[3] = {(cuuint64_t)p.Cin * 2, (cuuint64_t)S * p.Cin * 2, (cuuint64_t)p.Lin * p.Cin * 2};
      cuuint32_t box[4] = {(cuuint32_t)p.BK, 1, (cuuint32_t)(p.JB ? p.JB : TC_BM), (cuuint32_t)(p.JB ? p.BB : 1)};
      rc = tc_encode(&plan.mA_hi, xh, 4, dims, str, box, p.BK);
      if (!rc) rc = tc_encode(&plan.mA_lo, plan.x3 ? xl : xh, 4, dims, str, box, p.BK);
    }
    if (rc) return rc;
    plan.cached_x = x_planes;
  }
  if (!plan.b_ready) {
    cuuint64_t dims[2] = {(cuuint64_t)p.Cin, (cuuint64_t)w.rows};
    cuuint64_t str[1] = {(cuuint64_t)p.Cin * 2};
    cuuint32_t box[2] = {(cuuint32_t)p.BK, (cuuint32_t)p.BN};
    int rc = tc_encode(&plan.mB_hi, w.hi, 2, dims, str, box, p.BK);
    if (!rc) rc = tc_encode(&plan.mB_lo, w.lo, 2, dims, str, box, p.BK);
    if (rc) return rc;
    plan.b_ready = true;
  }
  cudaError_t e;
  if (plan.slab) {
    if (plan.x3) {
      e = cudaFuncSetAttribute(conv_tc2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem);
      if (e != cudaSuccess) return -2;
      tc_launch(conv_tc2_kernel<1>, plan.grid, TC_THREADS, plan.smem, st, plan.mA_hi, plan.mA_lo, plan.mB_hi, plan.mB_lo, p);
    } else {
      e = cudaFuncSetAttribute(conv_tc2_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem);
      if (e != cudaSuccess) return -2;
      tc_launch(conv_tc2_kernel<0>, plan.grid, TC_THREADS, plan.smem, st, plan.mA_hi, plan.mA_lo, plan.mB_hi, plan.mB_lo, p);
    }
    return 0;
  }
  if (p.dmul) {
    // backward-data epilogue (EPI = 2): a separate instantiation, so the forward kernels' code and registers are untouched
    if (plan.x3) {
      e = cudaFuncSetAttribute(conv_tc_kernel<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem);
      if (e != cudaSuccess) return -2;
      tc_launch(conv_tc_kernel<1, 2>, plan.grid, TC_THREADS, plan.smem, st, plan.mA_hi, plan.mA_lo, plan.mB_hi, plan.mB_lo, p);
    } else {
      e = cudaFuncSetAttribute(conv_tc_kernel<0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem);
      if (e != cudaSuccess) return -2;
      tc_launch(conv_tc_kernel<0, 2>, plan.grid, TC_THREADS, plan.smem, st, plan.mA_hi, plan.mA_lo, plan.mB_hi, plan.mB_lo, p);
    }
    return 0;
  }
  if (plan.x3) {
    e = cudaFuncSetAttribute(conv_tc_kernel<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem);
    if (e != cudaSuccess) return -2;
    tc_launch(conv_tc_kernel<1, 0>, plan.grid, TC_THREADS, plan.smem, st, plan.mA_hi, plan.mA_lo, plan.mB_hi, plan.mB_lo, p);
  } else {
    e = cudaFuncSetAttribute(conv_tc_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem);
    if (e != cudaSuccess) return -2;
    tc_launch(conv_tc_kernel<0, 0>, plan.grid, TC_THREADS, plan.smem, st, plan.mA_hi, plan.mA_lo, plan.mB_hi, plan.mB_lo, p);
  }
  return 0;
}


// ---------------------------------------------------------------------------------------------
// Nearest-code search on tensor cores: prep (bf16 hi/lo planes, norms) -> score GEMM with the argmax
// epilogue above -> exact fp32 finalisation of the candidates.
// ---------------------------------------------------------------------------------------------
// rows [n, d] fp32 -> bf16 planes; optional 0.5*|row|^2 (fmaf chain, the FP32 kernel's arithmetic), |row|^2
// and the running maximum of |row|^2 (as ordered int bits; the values are non-negative).
__global__ void nearest_prep(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                             float* __restrict__ half_norm, float* __restrict__ norm2, int* __restrict__ max_norm2_bits,
                             int n, int d) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  float s = 0.f;
  for (int i = 0; i < d; ++i) {
    const float v = __ldg(x + (size_t)r * d + i);
    s = fmaf(v, v, s);
    __nv_bfloat16 h, l;
    split_bf16(v, h, l);
    hi[(size_t)r * d + i] = h;
    lo[(size_t)r * d + i] = l;
  }
  if (half_norm) half_norm[r] = 0.5f * s;
  if (norm2) norm2[r] = s;
  if (max_norm2_bits) atomicMax(max_norm2_bits, __float_as_int(s));
}

// query rows: one warp per row, coalesced loads and plane stores; |row|^2 only feeds the error bound
// nearest_prep with coalesced traffic: 32 rows per CTA staged in shared memory (pitch d + 1), the squared norm of a row
// is still ONE thread's fmaf chain over d = 0 .. d-1 (the FP32 kernel's half norms, bit for bit); the bf16 hi / lo
// planes are written element-strided.  (Thread-per-row global access took 38 us for an 8192 x 96 codebook.)
__global__ void __launch_bounds__(128) nearest_prep_tiled(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi,
                                                          __nv_bfloat16* __restrict__ lo, float* __restrict__ half_norm,
                                                          float* __restrict__ norm2, int* __restrict__ max_norm2_bits,
                                                          int n, int d) {
  extern __shared__ float pt[];                    // [32][d + 1]
  const int r0 = blockIdx.x * 32, rows = min(32, n - r0), dp = d + 1;
  const size_t base = (size_t)r0 * d;
  for (int i = threadIdx.x; i < rows * d; i += 128) {
    const int rr = i / d, c = i - rr * d;
    const float v = __ldg(x + base + i);
    pt[rr * dp + c] = v;
    __nv_bfloat16 h, l;
    split_bf16(v, h, l);
    hi[base + i] = h;
    lo[base + i] = l;
  }
  __syncthreads();
  if ((int)threadIdx.x < rows) {
    const float* row = pt + threadIdx.x * dp;
    float s = 0.f;
    for (int i = 0; i < d; ++i) s = fmaf(row[i], row[i], s);
    const int r = r0 + threadIdx.x;
    if (half_norm) half_norm[r] = 0.5f * s;
    if (norm2) norm2[r] = s;
    if (max_norm2_bits) atomicMax(max_norm2_bits, __float_as_int(s));
  }
}
__global__ void __launch_bounds__(256) nearest_prep_rows(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi,
                                                         __nv_bfloat16* __restrict__ lo, float* __restrict__ norm2, int n, int d) {
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;  // provably warp-uniform: ptxas keeps the role code on the uniform datapath
  const int r = blockIdx.x * 8 + warp;
  if (r >= n) return;
  float s = 0.f;
  for (int i = lane; i < d; i += 32) {
    const float v = __ldg(x + (size_t)r * d + i);
    s = fmaf(v, v, s);
    __nv_bfloat16 h, l;
    split_bf16(v, h, l);
    hi[(size_t)r * d + i] = h;
    lo[(size_t)r * d + i] = l;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) norm2[r] = s * 1.0001f;   // a hair above any summation order's result: it scales a tolerance
}

__device__ __forceinline__ float exact_score(const float* __restrict__ xr, const float* __restrict__ e,
                                             const float* __restrict__ hn, int k, int D) {
  const float* ek = e + (size_t)k * D;
  float acc = 0.f;
  if ((D & 15) == 0 && ((reinterpret_cast<uintptr_t>(ek) | reinterpret_cast<uintptr_t>(xr)) & 15) == 0) {
    // same fmaf chain, operands fetched 16 elements at a time (four independent 16-byte loads in flight): the scalar
    // loop pays one L2 round trip per element when a thread walks a code row of its own
    const float4* e4 = reinterpret_cast<const float4*>(ek);
    const float4* x4 = reinterpret_cast<const float4*>(xr);
    for (int d4 = 0; d4 < (D >> 2); d4 += 4) {
      const float4 ea = __ldg(e4 + d4), eb = __ldg(e4 + d4 + 1), ec = __ldg(e4 + d4 + 2), ed = __ldg(e4 + d4 + 3);
      const float4 xa = __ldg(x4 + d4), xb = __ldg(x4 + d4 + 1), xc = __ldg(x4 + d4 + 2), xd = __ldg(x4 + d4 + 3);
      acc = fmaf(xa.x, ea.x, acc); acc = fmaf(xa.y, ea.y, acc); acc = fmaf(xa.z, ea.z, acc); acc = fmaf(xa.w, ea.w, acc);
      acc = fmaf(xb.x, eb.x, acc); acc = fmaf(xb.y, eb.y, acc); acc = fmaf(xb.z, eb.z, acc); acc = fmaf(xb.w, eb.w, acc);
      acc = fmaf(xc.x, ec.x, acc); acc = fmaf(xc.y, ec.y, acc); acc = fmaf(xc.z, ec.z, acc); acc = fmaf(xc.w, ec.w, acc);
      acc = fmaf(xd.x, ed.x, acc); acc = fmaf(xd.y, ed.y, acc); acc = fmaf(xd.z, ed.z, acc); acc = fmaf(xd.w, ed.w, acc);
    }
  } else {
    for (int d = 0; d < D; ++d) acc = fmaf(__ldg(xr + d), __ldg(ek + d), acc);
  }
  return __fsub_rn(acc, __ldg(hn + k));
}

// One warp per row.  A code can only be the fp32 arg-max if its tensor-core score is within tol = 2 * (error
// bound of the bf16x3 contraction) of the tensor-core maximum; those few candidates are re-scored with the FP32
// kernel's exact arithmetic (sequential fmaf over d, minus 0.5|e|^2) and the first maximum wins -- so the result
// equals the FP32 kernel's bit for bit.  A tile whose top two tensor-core scores are closer than tol is
// re-scanned completely.
__global__ void __launch_bounds__(256) nearest_finalize(const float4* __restrict__ cand, int n_tiles, int BN,
                                                        const float* __restrict__ x, const float* __restrict__ emb,
                                                        const float* __restrict__ hn, const float* __restrict__ xnorm2,
                                                        const int* __restrict__ emax2_bits, int N, int D, int K,
                                                        int* __restrict__ idx, int* __restrict__ defer_count = nullptr,
                                                        int* __restrict__ defer_list = nullptr,
                                                        unsigned long long* __restrict__ defer_keys = nullptr) {
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;  // provably warp-uniform: ptxas keeps the role code on the uniform datapath
  const int row = blockIdx.x * 8 + warp;
  if (row >= N) return;
  const float4* cr = cand + (size_t)row * n_tiles;
  const float* xr = x + (size_t)row * D;
  float M = -INFINITY;
  for (int t = lane; t < n_tiles; t += 32) M = fmaxf(M, cr[t].x);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, o));
  const float tol = 1.220703125e-4f * sqrtf(__ldg(xnorm2 + row) * __int_as_float(__ldg(emax2_bits))) + 1e-30f;
  float ebest = -INFINITY;
  int eidx = 0x7fffffff;
  for (int t0 = 0; t0 < n_tiles; t0 += 32) {
    const int t = t0 + lane;
    bool rescan = false;
    if (t < n_tiles) {
      const float4 c = cr[t];
      if (c.x >= M - tol) {
        if (c.x - c.z >= tol) {
          const int k = __float_as_int(c.y);
          const float sc = exact_score(xr, emb, hn, k, D);
          if (sc > ebest || (sc == ebest && k < eidx)) { ebest = sc; eidx = k; }
        } else {
          rescan = true;
        }
      }
    }
    unsigned m = __ballot_sync(0xffffffffu, rescan);
    if (m && defer_list && BN > 1024) {
      // a wide code range to re-scan (one candidate per work item): one warp would walk it for milliseconds; the row
      // goes to nearest_rescan, which gives it a whole CTA
      if (lane == 0) {
        const int pos = atomicAdd(defer_count, 1);
        defer_list[pos] = row;
        defer_keys[pos] = 0ull;
      }
      return;
    }
    while (m) {
      const int tt = t0 + __ffs(m) - 1;
      m &= m - 1;
      const int k1 = min(K, (tt + 1) * BN);
      for (int k = tt * BN + lane; k < k1; k += 32) {
        const float sc = exact_score(xr, emb, hn, k, D);
        if (sc > ebest || (sc == ebest && k < eidx)) { ebest = sc; eidx = k; }
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float os = __shfl_xor_sync(0xffffffffu, ebest, o);
    const int oi = __shfl_xor_sync(0xffffffffu, eidx, o);
    if (os > ebest || (os == ebest && oi < eidx)) { ebest = os; eidx = oi; }
  }
  if (lane == 0) idx[row] = eidx < K ? eidx : 0;
}

// Finalisation for one candidate per (row, work item), ONE THREAD per row: a code can only be the fp32 arg-max if its
// tensor-core score is within tol of the row's maximum M; an item whose top two are at least tol apart contributes its
// best code, re-scored with the FP32 kernel's exact arithmetic.  A row with an ambiguous item (0.7 % of the rows at
// K = 8192) is deferred to nearest_rescan.  (One warp per row with one active lane took 68 us for 65536 rows.)
__global__ void __launch_bounds__(256) nearest_finalize_rows(const float4* __restrict__ cand, int n_items,
                                                             const float* __restrict__ x, const float* __restrict__ emb,
                                                             const float* __restrict__ hn, const float* __restrict__ xnorm2,
                                                             const int* __restrict__ emax2_bits, int N, int D, int K,
                                                             int* __restrict__ idx, int* __restrict__ defer_count,
                                                             int* __restrict__ defer_list,
                                                             unsigned long long* __restrict__ defer_keys) {
  const int row = blockIdx.x * 256 + threadIdx.x;
  if (row >= N) return;
  const float4* cr = cand + (size_t)row * n_items;
  const float* xr = x + (size_t)row * D;
  float M = -INFINITY;
  for (int t = 0; t < n_items; ++t) M = fmaxf(M, cr[t].x);
  const float tol = 1.220703125e-4f * sqrtf(__ldg(xnorm2 + row) * __int_as_float(__ldg(emax2_bits))) + 1e-30f;
  float ebest = -INFINITY;
  int eidx = 0x7fffffff;
  bool ambiguous = false;
  for (int t = 0; t < n_items; ++t) {
    const float4 c = cr[t];
    if (c.x >= M - tol) {
      if (c.x - c.z >= tol) {
        const int k = __float_as_int(c.y);
        const float sc = exact_score(xr, emb, hn, k, D);
        if (sc > ebest || (sc == ebest && k < eidx)) { ebest = sc; eidx = k; }
      } else {
        ambiguous = true;
      }
    }
  }
  if (ambiguous) {
    const int pos = atomicAdd(defer_count, 1);
    defer_list[pos] = row;
    defer_keys[pos] = 0ull;
    return;
  }
  idx[row] = eidx < K ? eidx : 0;
}

// The deferred rows, 16 at a time: work unit = (group of <= 16 deferred rows, work item, 256-code chunk of the item).
// The chunk's code rows are staged once in shared memory (coalesced 16-byte copies, pitch D + 4) and serve every row
// of the group that needs this item re-scanned; thread = code, the score is the FP32 kernel's fmaf chain over
// d = 0 .. D-1.  An item that is a clear candidate of a deferred row contributes its best code (chunk 0).  The
// units of a row are combined by a 64-bit atomicMax on (score bits, ~index) = first maximum; nearest_rescan_apply
// writes the indices.  D % 4 == 0.
constexpr int RESCAN_ROWS = 16;
__global__ void __launch_bounds__(256) nearest_rescan(const float4* __restrict__ cand, int n_items, int range,
                                                      const float* __restrict__ x, const float* __restrict__ emb,
                                                      const float* __restrict__ hn, const float* __restrict__ xnorm2,
                                                      const int* __restrict__ emax2_bits, int D, int K,
                                                      const int* __restrict__ defer_count, const int* __restrict__ defer_list,
                                                      unsigned long long* __restrict__ defer_keys, int CH) {
  extern __shared__ __align__(16) float rs[];            // [CH][D + 4] codes (CH = 256, or 128 for wide D), then [16][D] rows
  __shared__ unsigned long long s_key[RESCAN_ROWS];
  __shared__ int s_row[RESCAN_ROWS], s_mode[RESCAN_ROWS];   // mode: 0 skip, 1 best code only, 2 re-scan
  const int lane = threadIdx.x & 31;
  const int DP = D + 4, vec_row = D >> 2;
  float* xs = rs + CH * DP;
  const int n = *defer_count;
  const int groups = (n + RESCAN_ROWS - 1) / RESCAN_ROWS;
  const int chunks = (range + CH - 1) / CH;
  const long units = (long)groups * n_items * chunks;
  for (long u = blockIdx.x; u < units; u += gridDim.x) {
    const int ch = (int)(u % chunks), t = (int)((u / chunks) % n_items), g = (int)(u / ((long)chunks * n_items));
    const int nr = min(RESCAN_ROWS, n - g * RESCAN_ROWS);
    __syncthreads();                                    // previous unit's shared state has been consumed
    if ((int)threadIdx.x < RESCAN_ROWS) {
      int mode = 0, row = 0;
      if ((int)threadIdx.x < nr) {
        row = defer_list[g * RESCAN_ROWS + threadIdx.x];
        const float4* cr = cand + (size_t)row * n_items;
        float M = -INFINITY;
        for (int tt = 0; tt < n_items; ++tt) M = fmaxf(M, cr[tt].x);
        const float tol = 1.220703125e-4f * sqrtf(__ldg(xnorm2 + row) * __int_as_float(__ldg(emax2_bits))) + 1e-30f;
        const float4 c = cr[t];
        if (c.x >= M - tol) mode = (c.x - c.z >= tol) ? 1 : 2;
      }
      s_row[threadIdx.x] = row;
      s_mode[threadIdx.x] = mode;
      s_key[threadIdx.x] = 0ull;
    }
    __syncthreads();
    bool any_scan = false;
    for (int j = 0; j < nr; ++j) any_scan |= s_mode[j] == 2;
    const int k0 = t * range + ch * CH, k1 = min(min(K, (t + 1) * range), k0 + CH);
    if (ch == 0 && (int)threadIdx.x < nr && s_mode[threadIdx.x] == 1) {
      // a clear candidate item of a deferred row: its best code, exactly
      const int row = s_row[threadIdx.x];
      const int k = __float_as_int(cand[(size_t)row * n_items + t].y);
      const float sc = exact_score(x + (size_t)row * D, emb, hn, k, D);
      if (k < K) atomicMax(&s_key[threadIdx.x], rvq_key(sc, k));
    }
    if (any_scan && k0 < k1) {
      const int nc = k1 - k0;
      for (int i = threadIdx.x; i < nc * vec_row; i += 256) {
        const int r = i / vec_row, v = i - r * vec_row;
        cp_async16(rs + r * DP + 4 * v, emb + (size_t)(k0 + r) * D + 4 * v);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      for (int i = threadIdx.x; i < nr * D; i += 256) {
        const int j = i / D, d = i - j * D;
        xs[i] = __ldg(x + (size_t)s_row[j] * D + d);
      }
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();
      const bool live = (int)threadIdx.x < nc;
      const int k = k0 + threadIdx.x;
      const float* er = rs + threadIdx.x * DP;
      const float hk = live ? __ldg(hn + k) : 0.f;
      // four rows per pass: four independent fmaf chains hide the chain latency (each is still d = 0 .. D-1 in order)
      for (int j0 = 0; j0 < nr; j0 += 4) {
        bool want[4];
        bool any = false;
#pragma unroll
        for (int u = 0; u < 4; ++u) { want[u] = j0 + u < nr && s_mode[j0 + u] == 2; any |= want[u]; }
        if (!any) continue;                             // block-uniform
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        if (live) {
          for (int d = 0; d < D; d += 4) {
            const float4 ev = *reinterpret_cast<const float4*>(er + d);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float4 xv = *reinterpret_cast<const float4*>(xs + min(j0 + u, nr - 1) * D + d);
              float a = acc[u];
              a = fmaf(xv.x, ev.x, a); a = fmaf(xv.y, ev.y, a); a = fmaf(xv.z, ev.z, a); a = fmaf(xv.w, ev.w, a);
              acc[u] = a;
            }
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (!want[u]) continue;                       // block-uniform
          unsigned long long key = live ? rvq_key(__fsub_rn(acc[u], hk), k) : 0ull;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long ok = __shfl_xor_sync(0xffffffffu, key, o);
            key = ok > key ? ok : key;
          }
          if (lane == 0 && key) atomicMax(&s_key[j0 + u], key);
        }
      }
    }
    __syncthreads();
    if ((int)threadIdx.x < nr && s_key[threadIdx.x])
      atomicMax(defer_keys + g * RESCAN_ROWS + threadIdx.x, s_key[threadIdx.x]);
  }
}
__global__ void __launch_bounds__(256) nearest_rescan_apply(const int* __restrict__ defer_count, const int* __restrict__ defer_list,
                                                            const unsigned long long* __restrict__ defer_keys, int K,
                                                            int* __restrict__ idx) {
  const int n = *defer_count;
  for (int pos = blockIdx.x * 256 + threadIdx.x; pos < n; pos += gridDim.x * 256) {
    const unsigned long long key = defer_keys[pos];
    const int bi = (int)(0xFFFFFFFFu - (unsigned int)(key & 0xFFFFFFFFull));
    idx[defer_list[pos]] = (key != 0ull && bi < K) ? bi : 0;
  }
}

// ---------------------------------------------------------------------------------------------
// Dedicated nearest-code score kernel ("rows resident").  A work item = one block of 128 rows x a contiguous
// range of code tiles.  The rows' bf16 hi/lo planes (all of D) are loaded ONCE per item and stay in shared memory;
// the TMA ring then carries one whole code tile (BN codes x D, both planes) per entry, i.e. ONE mbarrier hand-off
// per 128 x BN score tile instead of one per K block, and only the codebook streams from L2.  MMA / TMEM /
// arg-max epilogue as in conv_tc_kernel<1, 1>.
// ---------------------------------------------------------------------------------------------
struct TcSearchParams {
  TcConvParams e;             // BN, BK, n_kblk, acc_stride, tmem_cols, sbo, layout_type, half_norm, cand, n_rows, n_codes, n_ntiles
  int row_blocks, splits, tiles_per_split, n_items;
  int a_bufs, b_stages;
  uint32_t a_blk_bytes;       // 128 * BK * 2 (one K block of one plane)
  uint32_t a_buf_bytes;       // n_kblk * a_blk_bytes * 2 planes
  uint32_t b_blk_bytes;       // BN * BK * 2
  uint32_t b_stage_bytes;     // n_kblk * b_blk_bytes * 2 planes
};

__global__ void __launch_bounds__(TC_THREADS, 1)
nearest_tc_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                  const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                  const __grid_constant__ TcSearchParams q) {
  const TcConvParams& p = q.e;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_afull[2];
  __shared__ __align__(8) uint64_t bar_aempty[2];
  __shared__ __align__(8) uint64_t bar_bfull[TC_MAX_STAGES];
  __shared__ __align__(8) uint64_t bar_bempty[TC_MAX_STAGES];
  __shared__ __align__(8) uint64_t bar_tfull[2];
  __shared__ __align__(8) uint64_t bar_tempty[2];
  __shared__ uint32_t tmem_base_s;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;  // provably warp-uniform: ptxas keeps the role code on the uniform datapath
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smemB = smem0 + (uint32_t)q.a_bufs * q.a_buf_bytes;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA_hi); tma_prefetch_desc(&tmA_lo); tma_prefetch_desc(&tmB_hi); tma_prefetch_desc(&tmB_lo);
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&bar_afull[i]), 1); mbar_init(smem_u32(&bar_aempty[i]), 1);
      mbar_init(smem_u32(&bar_tfull[i]), 1); mbar_init(smem_u32(&bar_tempty[i]), TC_EPI_WARPS);   // one arrive per epilogue warp
    }
    for (int i = 0; i < q.b_stages; ++i) { mbar_init(smem_u32(&bar_bfull[i]), 1); mbar_init(smem_u32(&bar_bempty[i]), 1); }
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
  const uint32_t a_plane = (uint32_t)p.n_kblk * q.a_blk_bytes;   // bytes from the hi to the lo plane of the rows
  const uint32_t b_plane = (uint32_t)p.n_kblk * q.b_blk_bytes;

  if (warp == 0) {
    if (lane == 0) {
      Ring ra, rb;
      for (int item = blockIdx.x; item < q.n_items; item += gridDim.x) {
        const int rbk = item / q.splits, sp = item - rbk * q.splits;
        const int nt0 = sp * q.tiles_per_split, nt1 = min(p.n_ntiles, nt0 + q.tiles_per_split);
        {
          const uint32_t sa = ra.s;
          mbar_wait(smem_u32(&bar_aempty[sa]), ra.par ^ 1u, 1);
          const uint32_t full = smem_u32(&bar_afull[sa]);
          mbar_expect_tx(full, q.a_buf_bytes);
          const uint32_t dst = smem0 + sa * q.a_buf_bytes;
          for (int kb = 0; kb < p.n_kblk; ++kb) {
            tma_load_3d(dst + kb * q.a_blk_bytes, &tmA_hi, full, kb * p.BK, rbk * TC_BM, 0);
            tma_load_3d(dst + a_plane + kb * q.a_blk_bytes, &tmA_lo, full, kb * p.BK, rbk * TC_BM, 0);
          }
          ra.next(q.a_bufs);
        }
        for (int nt = nt0; nt < nt1; ++nt, rb.next(q.b_stages)) {
          const uint32_t sb = rb.s;
          mbar_wait(smem_u32(&bar_bempty[sb]), rb.par ^ 1u, 2);
          const uint32_t full = smem_u32(&bar_bfull[sb]);
          mbar_expect_tx(full, q.b_stage_bytes);
          const uint32_t dst = smemB + sb * q.b_stage_bytes;
          for (int kb = 0; kb < p.n_kblk; ++kb) {
            tma_load_2d(dst + kb * q.b_blk_bytes, &tmB_hi, full, kb * p.BK, nt * p.BN);
            tma_load_2d(dst + b_plane + kb * q.b_blk_bytes, &tmB_lo, full, kb * p.BK, nt * p.BN);
          }
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = umma_idesc_bf16(TC_BM, p.BN);
    const int ksteps = p.BK / 16;
    const uint32_t desc_hi = (uint32_t)(umma_desc_base(p.sbo, p.layout_type) >> 32);
    Ring ra, rb;
    uint32_t tcount = 0;
    for (int item = blockIdx.x; item < q.n_items; item += gridDim.x) {
      const int rbk = item / q.splits, sp = item - rbk * q.splits;
      const int nt0 = sp * q.tiles_per_split, nt1 = min(p.n_ntiles, nt0 + q.tiles_per_split);
      const uint32_t sa = ra.s;
      mbar_wait(smem_u32(&bar_afull[sa]), ra.par, 3);
      tc_fence_after();
      const uint32_t a_base = ((smem0 + sa * q.a_buf_bytes) & 0x3FFFFu) >> 4;
      for (int nt = nt0; nt < nt1; ++nt, rb.next(q.b_stages), ++tcount) {
        const uint32_t acc = tcount & 1u, apar = (tcount >> 1) & 1u;
        const uint32_t sb = rb.s;
        mbar_wait(smem_u32(&bar_bfull[sb]), rb.par, 4);
        mbar_wait(smem_u32(&bar_tempty[acc]), apar ^ 1u, 5);
        tc_fence_after();
        const uint32_t b_base = ((smemB + sb * q.b_stage_bytes) & 0x3FFFFu) >> 4;
        const uint32_t d = tmem_base + acc * p.acc_stride;
        for (int kb = 0; kb < p.n_kblk; ++kb) {
          const uint32_t a_lo = a_base + ((uint32_t)kb * q.a_blk_bytes >> 4), b_lo = b_base + ((uint32_t)kb * q.b_blk_bytes >> 4);
          if (ksteps == 4) umma_ksteps<1, 4>(d, a_lo, b_lo, a_plane >> 4, b_plane >> 4, desc_hi, idesc, kb != 0);
          else if (ksteps == 2) umma_ksteps<1, 2>(d, a_lo, b_lo, a_plane >> 4, b_plane >> 4, desc_hi, idesc, kb != 0);
          else umma_ksteps<1, 1>(d, a_lo, b_lo, a_plane >> 4, b_plane >> 4, desc_hi, idesc, kb != 0);
        }
        umma_commit_w(smem_u32(&bar_bempty[sb]));
        umma_commit_w(smem_u32(&bar_tfull[acc]));
      }
      umma_commit_w(smem_u32(&bar_aempty[sa]));
      ra.next(q.a_bufs);
    }
  } else {
    float* stg = reinterpret_cast<float*>(smem_raw + (smem0 - smem_u32(smem_raw)) + (size_t)q.a_bufs * q.a_buf_bytes +
                                          (size_t)q.b_stages * q.b_stage_bytes);
    // Arg-max over ALL code tiles of a work item in registers: a thread keeps (best, first index of best, runner-up)
    // for its row and its quarter of every tile; nothing is merged, synchronised CTA-wide or written per tile -- a
    // warp releases the accumulator with its own arrive (bar_tempty counts the 16 epilogue warps) and stages the
    // 0.5|e|^2 of its quarter in a warp-private slot (requested one tile ahead, so its L2 latency hides behind the
    // scan).  One merge + one candidate per (row, item): cand[row * splits + sp] = {best, idx, runner-up, -};
    // nearest_finalize treats an item's code range as one "tile".  (Per-tile candidates cost two 512-thread barriers,
    // an exposed global load and a 16-byte scattered store per row and tile: 3000 cycles per 128 x 128 tile against
    // 1152 cycles of MMAs.)
    const int quad = warp & 3, part = (warp - 2) >> 2;
    const int row = quad * 32 + lane;
    const int cpp = p.BN >> 2;                            // codes per quarter: 16 .. 64
    float* hn_w = stg + 2048 + (warp - 2) * 128;           // [2][64] per warp, after the 8 KB merge tile (16 of the 20 KB)
    float4* sm = reinterpret_cast<float4*>(stg);           // [4][128] merge tile
    uint32_t tcount = 0;
    auto load_hn = [&](int nt, float& h0, float& h1) {
      const int c0 = nt * p.BN + part * cpp + lane, c1 = c0 + 32;
      h0 = (lane < cpp && c0 < p.n_codes) ? __ldg(p.half_norm + c0) : 0.f;
      h1 = (lane + 32 < cpp && c1 < p.n_codes) ? __ldg(p.half_norm + c1) : 0.f;
    };
    float h0 = 0.f, h1 = 0.f;
    if ((int)blockIdx.x < q.n_items) load_hn(((int)blockIdx.x % q.splits) * q.tiles_per_split, h0, h1);
    for (int item = blockIdx.x; item < q.n_items; item += gridDim.x) {
      const int rbk = item / q.splits, sp = item - rbk * q.splits;
      const int nt0 = sp * q.tiles_per_split, nt1 = min(p.n_ntiles, nt0 + q.tiles_per_split);
      float best = -INFINITY, second = -INFINITY;
      int bidx = 0x7fffffff;
      for (int nt = nt0; nt < nt1; ++nt, ++tcount) {
        const uint32_t acc = tcount & 1u, apar = (tcount >> 1) & 1u;
        float* hw = hn_w + (tcount & 1u) * 64;
        hw[lane] = h0; hw[lane + 32] = h1;
        __syncwarp();
        int nt_next = nt + 1;
        if (nt_next == nt1) {
          const int item2 = item + (int)gridDim.x;
          nt_next = item2 < q.n_items ? (item2 % q.splits) * q.tiles_per_split : -1;
        }
        if (nt_next >= 0) load_hn(nt_next, h0, h1);
        mbar_wait(smem_u32(&bar_tfull[acc]), apar, 6);
        tc_fence_after();
        const int col0 = nt * p.BN + part * cpp;
        const int ncol = min(cpp, p.n_codes - col0);
        const uint32_t t_src = tmem_base + acc * p.acc_stride + ((uint32_t)(quad * 32) << 16) + part * cpp;
        for (int c = 0; c < cpp; c += 16) {
          float v[16];
          tmem_ld16(t_src + c, v);
          if (c + 16 <= ncol) {
            float hv[16];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float4 h4 = *reinterpret_cast<const float4*>(hw + c + 4 * u);
              hv[4 * u] = h4.x; hv[4 * u + 1] = h4.y; hv[4 * u + 2] = h4.z; hv[4 * u + 3] = h4.w;
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float sc = __fsub_rn(v[i], hv[i]);
              second = fmaxf(second, fminf(best, sc));     // runner-up = max over all but the (first) best
              if (sc > best) { best = sc; bidx = col0 + c + i; }
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              if (c + i < ncol) {
                const float sc = __fsub_rn(v[i], hw[c + i]);
                second = fmaxf(second, fminf(best, sc));
                if (sc > best) { best = sc; bidx = col0 + c + i; }
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bar_tempty[acc]));
      }
      // merge the four quarters of every row (a quarter's codes ascend inside a tile but interleave with the other
      // quarters across tiles: equal scores go to the smaller index explicitly)
      sm[part * TC_BM + row] = make_float4(best, __int_as_float(bidx), second, 0.f);
      asm volatile("bar.sync 1, 512;" ::: "memory");
      if (part == 0) {
        const int grow = rbk * TC_BM + row;
#pragma unroll
        for (int qq = 1; qq < 4; ++qq) {
          const float4 o = sm[qq * TC_BM + row];
          const int oi = __float_as_int(o.y);
          if (o.x > best || (o.x == best && oi < bidx)) {
            second = fmaxf(best, fmaxf(second, o.z));
            best = o.x; bidx = oi;
          } else {
            second = fmaxf(second, o.x);
          }
        }
        if (grow < p.n_rows) p.cand[(size_t)grow * q.splits + sp] = make_float4(best, __int_as_float(bidx), second, 0.f);
      }
      asm volatile("bar.sync 1, 512;" ::: "memory");      // the merge tile is rewritten at the end of the next item
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

// rows-resident plan: returns false when the rows of one block (all of D, both planes) plus two code-tile stages
// do not fit in shared memory (D = 256): the caller then uses conv_tc_kernel<1, 1>
inline bool tc_search_plan(int N, int D, int K, int sm_count, TcSearchParams* q, size_t* smem, int* grid) {
  if (D % 8 != 0 || D < 8 || D > 256 || K < 1 || N < 1) return false;
  memset(q, 0, sizeof(*q));
  TcConvParams& p = q->e;
  p.BK = D % 64 == 0 ? 64 : (D % 32 == 0 ? 32 : 16);
  p.n_kblk = (D + p.BK - 1) / p.BK;
  q->a_blk_bytes = TC_BM * p.BK * 2;
  q->a_buf_bytes = (uint32_t)p.n_kblk * q->a_blk_bytes * 2;
  const int budget = 232448 - 2048 - 1024 - 20480;    // 20 KB: the arg-max epilogue's two merge tiles + 0.5|e|^2 tiles
  int bn = K >= 256 ? 256 : (K + 63) / 64 * 64;
  for (;; bn >>= 1) {
    q->b_blk_bytes = (uint32_t)bn * p.BK * 2;
    q->b_stage_bytes = (uint32_t)p.n_kblk * q->b_blk_bytes * 2;
    const long rest2 = (long)budget - 2L * q->a_buf_bytes, rest1 = (long)budget - (long)q->a_buf_bytes;
    if (rest2 >= 2L * q->b_stage_bytes) { q->a_bufs = 2; q->b_stages = (int)(rest2 / q->b_stage_bytes); break; }
    if (rest1 >= 2L * q->b_stage_bytes) { q->a_bufs = 1; q->b_stages = (int)(rest1 / q->b_stage_bytes); break; }
    if (bn <= 64) return false;
  }
  if (q->b_stages > TC_MAX_STAGES) q->b_stages = TC_MAX_STAGES;
  p.BN = bn;
  p.n_ntiles = (K + bn - 1) / bn;
  p.acc_stride = bn <= 64 ? 64 : (bn <= 128 ? 128 : 256);
  p.tmem_cols = 2 * p.acc_stride;
  p.sbo = 8 * p.BK * 2;
  p.layout_type = p.BK == 64 ? 2u : (p.BK == 32 ? 4u : 6u);
  p.stg_bufs = 2;
  q->row_blocks = (N + TC_BM - 1) / TC_BM;
  int splits = (2 * sm_count + q->row_blocks - 1) / q->row_blocks;      // aim at >= 2 items per SM
  if (splits > p.n_ntiles) splits = p.n_ntiles;
  if (splits < 1) splits = 1;
  q->tiles_per_split = (p.n_ntiles + splits - 1) / splits;
  q->splits = (p.n_ntiles + q->tiles_per_split - 1) / q->tiles_per_split;
  const long items = (long)q->row_blocks * q->splits;
  if (items > 0x7fffffffL) return false;
  q->n_items = (int)items;
  *grid = (int)(items < sm_count ? items : sm_count);
  *smem = (size_t)q->a_bufs * q->a_buf_bytes + (size_t)q->b_stages * q->b_stage_bytes + 20480 + 1024;
  return true;
}

inline CUtensorMapSwizzle tc_swizzle_of(int bk) {
  return bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (bk == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

struct TcNearestDims {
  int BN, BK, n_kblk, stages, tiles_j, n_ntiles;
  size_t smem;
};
inline bool tc_nearest_dims(int N, int D, int K, TcNearestDims* o) {
  if (D % 8 != 0 || D < 8 || D > 256 || K < 1 || N < 1) return false;
  o->BK = D % 64 == 0 ? 64 : (D % 32 == 0 ? 32 : 16);
  o->n_kblk = (D + o->BK - 1) / o->BK;
  o->BN = K >= 256 ? 256 : (K + 63) / 64 * 64;
  uint32_t stage = 0;
  int stages = 0;
  for (;; o->BN >>= 1) {   // narrower code tiles until three pipeline stages fit (D = 256: BN = 128)
    stage = (uint32_t)(TC_BM + o->BN) * o->BK * 2 * 2;
    stages = (232448 - 2048 - TC_STG_BYTES - 1024) / (int)stage;
    if (stages >= 3 || o->BN <= 64) break;
  }
  o->tiles_j = (N + TC_BM - 1) / TC_BM;
  o->n_ntiles = (K + o->BN - 1) / o->BN;
  if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
  if (stages < 2) return false;
  o->stages = stages;
  o->smem = (size_t)stages * stage + TC_STG_BYTES + 1024;
  return (long)o->tiles_j * o->n_ntiles < 0x7fffffffL;
}
// scratch layout (bytes, 256-aligned pieces): x planes | emb planes | half_norm[K] | xnorm2[N] | emax2 | cand
inline size_t tc_nearest_scratch_bytes(int N, int D, int K) {
  TcNearestDims d;
  if (!tc_nearest_dims(N, D, K, &d)) return (size_t)K * 4;
  auto al = [](size_t v) { return (v + 255) / 256 * 256; };
  const size_t max_tiles = (size_t)(K + 63) / 64;   // the narrowest code tile either kernel uses
  return al((size_t)N * D * 4) + al((size_t)K * D * 4) + al((size_t)K * 4) + al((size_t)N * 4) + 256 +
         al((size_t)N * max_tiles * 16) + al((size_t)N * 4) + al((size_t)N * 8);   // + rows deferred to nearest_rescan, their keys
}

// returns 0 = launched; 1 = shape not eligible (caller uses the FP32 kernel); < 0 = error
inline int tc_nearest_launch(const float* x, const float* emb, void* scratch, int* idx, int N, int D, int K,
                             int sm_count, cudaStream_t st) {
  TcNearestDims d;
  if (!tc_nearest_dims(N, D, K, &d)) return 1;
  if (!tc_encode_fn()) return -10;
  auto al = [](size_t v) { return (v + 255) / 256 * 256; };
  char* s = reinterpret_cast<char*>(scratch);
  __nv_bfloat16* xh = reinterpret_cast<__nv_bfloat16*>(s);
  __nv_bfloat16* xl = xh + (size_t)N * D;
  s += al((size_t)N * D * 4);
  __nv_bfloat16* eh = reinterpret_cast<__nv_bfloat16*>(s);
  __nv_bfloat16* el = eh + (size_t)K * D;
  s += al((size_t)K * D * 4);
  float* hn = reinterpret_cast<float*>(s);
  s += al((size_t)K * 4);
  float* xn2 = reinterpret_cast<float*>(s);
  s += al((size_t)N * 4);
  int* emax = reinterpret_cast<int*>(s);
  s += 256;
  float4* cand = reinterpret_cast<float4*>(s);
  s += al((size_t)N * ((size_t)(K + 63) / 64) * 16);
  int* defer_list = reinterpret_cast<int*>(s);
  s += al((size_t)N * 4);
  unsigned long long* defer_keys = reinterpret_cast<unsigned long long*>(s);
  int* defer_count = emax + 1;                                  // second word of the 256-byte slot
  if (cudaMemsetAsync(emax, 0, 8, st) != cudaSuccess) return -3;
  nearest_prep_rows<<<(N + 7) / 8, 256, 0, st>>>(x, xh, xl, xn2, N, D);
  {
    const size_t pt_smem = (size_t)32 * (D + 1) * sizeof(float);
    if (pt_smem > 48 * 1024 &&
        cudaFuncSetAttribute(nearest_prep_tiled, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pt_smem) != cudaSuccess)
      return -2;
    nearest_prep_tiled<<<(K + 31) / 32, 128, pt_smem, st>>>(emb, eh, el, hn, nullptr, emax, K, D);
  }

  {
    // preferred: rows resident in shared memory, one hand-off per score tile
    TcSearchParams q;
    size_t smem = 0;
    int grid = 0;
    static int use_rows = -1;
    if (use_rows < 0) {
      const char* e = getenv("B2C_SEARCH_ROWS");
      use_rows = (e && e[0] == '0') ? 0 : 1;
    }
    if (use_rows && tc_search_plan(N, D, K, sm_count, &q, &smem, &grid)) {
      TcConvParams& p = q.e;
      p.half_norm = hn; p.cand = cand; p.n_rows = N; p.n_codes = K;
      CUtensorMap mA_hi, mA_lo, mB_hi, mB_lo;
      cuuint32_t es[3] = {1, 1, 1};
      cuuint64_t adims[3] = {(cuuint64_t)D, (cuuint64_t)N, 1};
      cuuint64_t astr[2] = {(cuuint64_t)D * 2, (cuuint64_t)N * D * 2};
      cuuint32_t abox[3] = {(cuuint32_t)p.BK, TC_BM, 1};
      cuuint64_t bdims[2] = {(cuuint64_t)D, (cuuint64_t)K};
      cuuint64_t bstr[1] = {(cuuint64_t)D * 2};
      cuuint32_t bbox[2] = {(cuuint32_t)p.BK, (cuuint32_t)p.BN};
      for (int pl = 0; pl < 2; ++pl) {
        if (tc_encode_fn()(pl ? &mA_lo : &mA_hi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, pl ? (void*)xl : (void*)xh, adims, astr,
                           abox, es, CU_TENSOR_MAP_INTERLEAVE_NONE, tc_swizzle_of(p.BK), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
          return -13;
        if (tc_encode_fn()(pl ? &mB_lo : &mB_hi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, pl ? (void*)el : (void*)eh, bdims, bstr,
                           bbox, es, CU_TENSOR_MAP_INTERLEAVE_NONE, tc_swizzle_of(p.BK), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
          return -14;
      }
      if (cudaFuncSetAttribute(nearest_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return -2;
      nearest_tc_kernel<<<grid, TC_THREADS, smem, st>>>(mA_hi, mA_lo, mB_hi, mB_lo, q);
      // one candidate per (row, work item): an item's code range is one "tile" of nearest_finalize
      // one candidate per (row, work item); rows with an ambiguous item go through nearest_rescan
      const int range = q.tiles_per_split * p.BN;
      nearest_finalize_rows<<<(N + 255) / 256, 256, 0, st>>>(cand, q.splits, x, emb, hn, xn2, emax, N, D, K, idx, defer_count,
                                                             defer_list, defer_keys);
      const int rs_codes = D <= 160 ? 256 : 128;
      const size_t rs_smem = ((size_t)rs_codes * (D + 4) + (size_t)RESCAN_ROWS * D) * sizeof(float);
      if (cudaFuncSetAttribute(nearest_rescan, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs_smem) != cudaSuccess) return -2;
      nearest_rescan<<<4 * sm_count, 256, rs_smem, st>>>(cand, q.splits, range, x, emb, hn, xn2, emax, D, K, defer_count,
                                                         defer_list, defer_keys, rs_codes);
      nearest_rescan_apply<<<sm_count, 256, 0, st>>>(defer_count, defer_list, defer_keys, K, idx);
      return 0;
    }
  }

  TcConvParams p;
  memset(&p, 0, sizeof(p));
  p.B = 1; p.Lin = N; p.Cin = D; p.Cout = K; p.KT = 1; p.in_step = 1; p.dil = 1; p.n_phase = 1;
  p.Lj = N; p.out_step = 1; p.Lout = N;
  p.BN = d.BN; p.BK = d.BK; p.n_kblk = d.n_kblk; p.stages = d.stages; p.tiles_j = d.tiles_j; p.n_ntiles = d.n_ntiles;
  p.total_tiles = d.tiles_j * d.n_ntiles;
  p.acc_stride = d.BN <= 64 ? 64 : (d.BN <= 128 ? 128 : 256);
  p.tmem_cols = 2 * p.acc_stride;
  p.a_bytes = TC_BM * d.BK * 2; p.b_bytes = d.BN * d.BK * 2;
  p.sbo = 8 * d.BK * 2;
  p.layout_type = d.BK == 64 ? 2u : (d.BK == 32 ? 4u : 6u);
  p.half_norm = hn; p.cand = cand; p.n_rows = N; p.n_codes = K; p.stg_bufs = 2; p.kgroup = 1;
  CUtensorMap mA_hi, mA_lo, mB_hi, mB_lo;
  cuuint32_t es[3] = {1, 1, 1};
  {
    cuuint64_t dims[3] = {(cuuint64_t)D, (cuuint64_t)N, 1};
    cuuint64_t str[2] = {(cuuint64_t)D * 2, (cuuint64_t)N * D * 2};
    cuuint32_t box[3] = {(cuuint32_t)d.BK, TC_BM, 1};
    for (int pl = 0; pl < 2; ++pl)
      if (tc_encode_fn()(pl ? &mA_lo : &mA_hi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, pl ? (void*)xl : (void*)xh, dims, str,
                         box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, tc_swizzle_of(d.BK), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return -11;
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)D, (cuuint64_t)K};
    cuuint64_t str[1] = {(cuuint64_t)D * 2};
    cuuint32_t box[2] = {(cuuint32_t)d.BK, (cuuint32_t)d.BN};
    for (int pl = 0; pl < 2; ++pl)
      if (tc_encode_fn()(pl ? &mB_lo : &mB_hi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, pl ? (void*)el : (void*)eh, dims, str,
                         box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, tc_swizzle_of(d.BK), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return -12;
  }
  if (cudaFuncSetAttribute(conv_tc_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d.smem) != cudaSuccess)
    return -2;
  const int grid = p.total_tiles < sm_count ? p.total_tiles : sm_count;
  conv_tc_kernel<1, 1><<<grid, TC_THREADS, d.smem, st>>>(mA_hi, mA_lo, mB_hi, mB_lo, p);
  nearest_finalize<<<(N + 7) / 8, 256, 0, st>>>(cand, d.n_ntiles, d.BN, x, emb, hn, xn2, emax, N, D, K, idx);
  return 0;
}

}  // namespace b2c
