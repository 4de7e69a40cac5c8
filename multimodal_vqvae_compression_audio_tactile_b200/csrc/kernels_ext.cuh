// Kernels of the callers either side of the codec path:
//   * full-length multi-head cross-attention of the packet-loss-concealment forward
//     (AllPredPLC.forward_step, PLC/PLC1_eval.py:467-520: ONE CrossPredictor call over all T_lat tokens),
//   * token selection  z_filled = where(mask, z_pred, zt_in)  (:497),
//   * ResidualVQEMA.ema_step (Training/compare_dacvsproposal_3.py:264-276): per-code counts, per-code sums in row
//     order (bit-equal to index_add_ on the CPU), EMA blend of the codebook rows that were hit.
// FP32 CUDA-core arithmetic: none of these decides a code index of the codec path and none is on the throughput
// path BASELINE.json times.
#pragma once
#include "kernels_f32.cuh"

namespace b2c {

// ---------------------------------------------------------------------------------------------
// softmax(Q K^T / sqrt(dh)) V over ALL keys of a frame (no chunking, no causal mask), head dim 128.
// Flash-style: a CTA owns 64 queries of one (batch, head) and walks the keys in tiles of 64 with an
// online softmax; the [T, T] score matrix is never written anywhere.
//   q [B*T, C], kv [B*T, 2C] (K | V), out [B*T, C], C = heads * 128, channel-last fp32.
// Thread (ty, tx) = (tid / 16, tid % 16): score rows 4ty..4ty+3, score columns tx + 16j; output rows
// 4ty..4ty+3, output columns 4tx..4tx+3 and 64 + 4tx..  Row statistics live in registers; the 16 threads of
// a row are the 16 lanes of a half-warp (shuffle reductions).
// ---------------------------------------------------------------------------------------------
struct AttnFullArgs {
  const float* q;
  const float* kv;
  float* out;
  int B, T, heads;
};

constexpr int AF_BQ = 64, AF_BK = 64, AF_DH = 128, AF_LD = 132, AF_PLD = 68;
constexpr size_t AF_SMEM = ((size_t)3 * AF_BQ * AF_LD + (size_t)AF_BQ * AF_PLD) * sizeof(float);

__global__ void __launch_bounds__(256) attention_full_f32(const AttnFullArgs p) {
  extern __shared__ float af_sm[];
  float* Qs = af_sm;
  float* Ks = Qs + AF_BQ * AF_LD;
  float* Vs = Ks + AF_BK * AF_LD;
  float* Ps = Vs + AF_BK * AF_LD;
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int q0 = blockIdx.x * AF_BQ, h = blockIdx.y, b = blockIdx.z;
  const int C = p.heads * AF_DH;
  const long row0 = (long)b * p.T;

  // Q tile (rows past T are zero; they are never stored)
  for (int i = tid; i < AF_BQ * (AF_DH / 4); i += 256) {
    const int r = i >> 5, c4 = i & 31;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q0 + r < p.T) v = __ldg(reinterpret_cast<const float4*>(p.q + (row0 + q0 + r) * C + h * AF_DH) + c4);
    *reinterpret_cast<float4*>(Qs + r * AF_LD + 4 * c4) = v;
  }
  float m[4], l[4], o[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m[i] = -INFINITY; l[i] = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) o[i][c] = 0.f;
  }
  const float sqrt_dh = 11.313708498984761f;   // the reference divides the scores by sqrt(dh) (PLC1_eval.py:411)

  for (int k0 = 0; k0 < p.T; k0 += AF_BK) {
    __syncthreads();                           // the previous tile's K / V / P are no longer read
    for (int i = tid; i < AF_BK * (AF_DH / 4); i += 256) {
      const int r = i >> 5, c4 = i & 31;
      float4 kk = make_float4(0.f, 0.f, 0.f, 0.f), vv = kk;
      if (k0 + r < p.T) {
        const float* row = p.kv + (row0 + k0 + r) * (2L * C) + h * AF_DH;
        kk = __ldg(reinterpret_cast<const float4*>(row) + c4);
        vv = __ldg(reinterpret_cast<const float4*>(row + C) + c4);
      }
      *reinterpret_cast<float4*>(Ks + r * AF_LD + 4 * c4) = kk;
      *reinterpret_cast<float4*>(Vs + r * AF_LD + 4 * c4) = vv;
    }
    __syncthreads();
    // ---- S = Q K^T (4 x 4 per thread)
    float s[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll 4
    for (int d = 0; d < AF_DH; d += 4) {
      float4 qv[4], kv4[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) qv[i] = *reinterpret_cast<const float4*>(Qs + (4 * ty + i) * AF_LD + d);
#pragma unroll
      for (int j = 0; j < 4; ++j) kv4[j] = *reinterpret_cast<const float4*>(Ks + (tx + 16 * j) * AF_LD + d);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          s[i][j] = fmaf(qv[i].x, kv4[j].x, s[i][j]);
          s[i][j] = fmaf(qv[i].y, kv4[j].y, s[i][j]);
          s[i][j] = fmaf(qv[i].z, kv4[j].z, s[i][j]);
          s[i][j] = fmaf(qv[i].w, kv4[j].w, s[i][j]);
        }
    }
    // ---- online softmax
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        s[i][j] = (k0 + tx + 16 * j < p.T) ? __fdiv_rn(s[i][j], sqrt_dh) : -INFINITY;
        mx = fmaxf(mx, s[i][j]);
      }
#pragma unroll
      for (int w = 8; w > 0; w >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, w));
      const float mn = fmaxf(m[i], mx);              // finite: every tile holds at least one valid key
      const float corr = expf(m[i] - mn);            // exp(-inf) = 0 on the first tile
      float rs = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float e = expf(s[i][j] - mn);
        Ps[(4 * ty + i) * AF_PLD + tx + 16 * j] = e;
        rs += e;
      }
#pragma unroll
      for (int w = 8; w > 0; w >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, w);
      l[i] = fmaf(l[i], corr, rs);
      m[i] = mn;
#pragma unroll
      for (int c = 0; c < 8; ++c) o[i][c] *= corr;
    }
    __syncthreads();
    // ---- O += P V (4 rows x 8 columns per thread)
#pragma unroll 2
    for (int k = 0; k < AF_BK; k += 4) {
      float4 pv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) pv[i] = *reinterpret_cast<const float4*>(Ps + (4 * ty + i) * AF_PLD + k);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const float4 v0 = *reinterpret_cast<const float4*>(Vs + (k + kk) * AF_LD + 4 * tx);
        const float4 v1 = *reinterpret_cast<const float4*>(Vs + (k + kk) * AF_LD + 64 + 4 * tx);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float pk = kk == 0 ? pv[i].x : kk == 1 ? pv[i].y : kk == 2 ? pv[i].z : pv[i].w;
          o[i][0] = fmaf(pk, v0.x, o[i][0]); o[i][1] = fmaf(pk, v0.y, o[i][1]);
          o[i][2] = fmaf(pk, v0.z, o[i][2]); o[i][3] = fmaf(pk, v0.w, o[i][3]);
          o[i][4] = fmaf(pk, v1.x, o[i][4]); o[i][5] = fmaf(pk, v1.y, o[i][5]);
          o[i][6] = fmaf(pk, v1.z, o[i][6]); o[i][7] = fmaf(pk, v1.w, o[i][7]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = q0 + 4 * ty + i;
    if (r >= p.T) continue;
    const float inv = 1.0f / l[i];
    float* dst = p.out + (row0 + r) * C + h * AF_DH;
    *reinterpret_cast<float4*>(dst + 4 * tx) = make_float4(o[i][0] * inv, o[i][1] * inv, o[i][2] * inv, o[i][3] * inv);
    *reinterpret_cast<float4*>(dst + 64 + 4 * tx) = make_float4(o[i][4] * inv, o[i][5] * inv, o[i][6] * inv, o[i][7] * inv);
  }
}

// out[n, :] = mask[n] ? a[n, :] : b[n, :]     (torch.where over tokens, PLC1_eval.py:503)
__global__ void __launch_bounds__(256) select_rows_f32(const unsigned char* __restrict__ mask, const float* __restrict__ a,
                                                       const float* __restrict__ b, float* __restrict__ out, long total4,
                                                       int C4) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const long n = i / C4;
  const float4 v = mask[n] ? __ldg(reinterpret_cast<const float4*>(a) + i) : __ldg(reinterpret_cast<const float4*>(b) + i);
  reinterpret_cast<float4*>(out)[i] = v;
}

// ---------------------------------------------------------------------------------------------
// ResidualVQEMA.ema_step for one book, given the nearest-code indices of the N token rows:
//   counts = bincount(idx); sums.index_add_(0, idx, X); means = sums / (counts + 1e-9);
//   emb[hit] = decay * emb[hit] + (1 - decay) * means[hit]                       (:268-276)
// One CTA per code, thread = dimension: the rows of a code are added in increasing row order, i.e. in the order the
// CPU's index_add_ adds them, so with equal indices the updated codebook is bit-equal to the reference's.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) ema_update_f32(const float* __restrict__ x, const int* __restrict__ idx, float* __restrict__ emb,
                                                      int* __restrict__ counts_out, int N, int D, float decay,
                                                      float one_minus_decay) {
  __shared__ int s_idx[1024];
  const int k = blockIdx.x;
  float sum[2] = {0.f, 0.f};                   // D <= 256: two dimensions per thread
  int count = 0;
  for (int n0 = 0; n0 < N; n0 += 1024) {
    const int cnt = min(1024, N - n0);
    __syncthreads();
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) s_idx[i] = __ldg(idx + n0 + i);
    __syncthreads();
    for (int i = 0; i < cnt; ++i) {
      if (s_idx[i] != k) continue;             // uniform branch: every thread of the CTA sees the same index
      ++count;
      const float* row = x + (long)(n0 + i) * D;
      if ((int)threadIdx.x < D) sum[0] = __fadd_rn(sum[0], __ldg(row + threadIdx.x));
      if ((int)threadIdx.x + 128 < D) sum[1] = __fadd_rn(sum[1], __ldg(row + threadIdx.x + 128));
    }
  }
  if (threadIdx.x == 0 && counts_out) counts_out[k] = count;
  if (count == 0) return;
  const float den = __fadd_rn((float)count, 1e-9f);
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int d = threadIdx.x + 128 * u;
    if (d >= D) continue;
    const float mean = __fdiv_rn(sum[u], den);
    float* e = emb + (long)k * D + d;
    *e = __fadd_rn(__fmul_rn(decay, *e), __fmul_rn(one_minus_decay, mean));
  }
}

// ---------------------------------------------------------------------------------------------
// Backward-data of the decoder's tail (the forward is Snake1d -> Conv1d(C, 1, 7, p=3) -> tanh,
// SURVEY App. A): with v the conv output and y = tanh(v),
//   g_v[l] = g_y[l] (1 - y[l]^2);   g_s[q][c] = sum_t w[t][c] g_v[q + 3 - t];   g_x[q][c] = g_s[q][c] snake'(x[q][c]; alpha[c])
// x is the pre-activation input of the last snake.  g_x is written as fp32 (the residual stream of the gradient)
// and in the activation format the next backward contraction reads.  HBM-bound: x read once, g_x written once.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) head_k7_bwd_f32(const float* __restrict__ g_y, const float* __restrict__ y,
                                                       const float* __restrict__ x_raw, const float* __restrict__ w,
                                                       const float* __restrict__ alpha, float* __restrict__ g_raw,
                                                       void* __restrict__ g_act, int act_fmt, size_t act_n, int L, int C) {
  constexpr int TP = 64;
  extern __shared__ float hb_sm[];
  float* gv = hb_sm;            // TP + 6: g_v[l0 - 3 .. l0 + TP + 3)
  float* ws = hb_sm + TP + 8;   // 7 * C
  const int b = blockIdx.y, l0 = blockIdx.x * TP;
  for (int i = threadIdx.x; i < TP + 6; i += blockDim.x) {
    const int l = l0 + i - 3;
    float g = 0.f;
    if (l >= 0 && l < L) {
      const float yy = __ldg(y + (size_t)b * L + l);
      g = __ldg(g_y + (size_t)b * L + l) * (1.0f - yy * yy);
    }
    gv[i] = g;
  }
  for (int i = threadIdx.x; i < 7 * C; i += blockDim.x) ws[i] = __ldg(w + i);
  __syncthreads();
  for (int idx = threadIdx.x; idx < TP * C; idx += blockDim.x) {
    const int pp = idx / C, c = idx - pp * C;
    const int l = l0 + pp;
    if (l >= L) break;
    float acc = 0.f;
#pragma unroll
    for (int t = 0; t < 7; ++t) acc = fmaf(ws[t * C + c], gv[pp + 6 - t], acc);
    const size_t o = ((size_t)b * L + l) * C + c;
    const float g = acc * dsnake_f(__ldg(x_raw + o), __ldg(alpha + c));
    if (g_raw) g_raw[o] = g;
    if (g_act) store_fmt(g_act, act_fmt, act_n, o, g);
  }
}

}  // namespace b2c
