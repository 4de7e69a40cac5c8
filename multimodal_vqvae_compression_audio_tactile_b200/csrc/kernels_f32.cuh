// FP32 CUDA-core kernels of the encode -> quantize -> decode path (sm_100a).
// These are the bit-faithful arm (fp32 everywhere, only the summation order differs from
// the reference's CPU fp32) and the on-GPU cross-check for the tcgen05 kernels.
// Activations are channel-last: [batch][position][channel].
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace b2c {

enum { ACT_NONE = 0, ACT_SNAKE = 1, ACT_GELU = 2, ACT_TANH = 3 };
// activation storage: fp32, two bf16 planes (hi | lo, lo = hi + n elements; x = hi + lo to 16 mantissa
// bits), or the bf16 hi plane alone
enum { FMT_F32 = 0, FMT_PLANES = 1, FMT_HI = 2 };

__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}
__device__ __forceinline__ uint32_t pack_bf16(__nv_bfloat16 a, __nv_bfloat16 b) {
  return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}
// store one activation value at element offset o of a buffer of n elements in format fmt
__device__ __forceinline__ void store_fmt(void* base, int fmt, size_t n, size_t o, float v) {
  if (fmt == FMT_F32) {
    reinterpret_cast<float*>(base)[o] = v;
  } else {
    __nv_bfloat16 h, l;
    split_bf16(v, h, l);
    reinterpret_cast<__nv_bfloat16*>(base)[o] = h;
    if (fmt == FMT_PLANES) reinterpret_cast<__nv_bfloat16*>(base)[n + o] = l;
  }
}
__device__ __forceinline__ float load_fmt(const void* base, int fmt, size_t n, size_t o) {
  if (fmt == FMT_F32) return __ldg(reinterpret_cast<const float*>(base) + o);
  const __nv_bfloat16* b = reinterpret_cast<const __nv_bfloat16*>(base);
  float v = __bfloat162float(b[o]);
  if (fmt == FMT_PLANES) v += __bfloat162float(b[n + o]);
  return v;
}
enum { ROWS_DENSE = 0, ROWS_HEAD = 1, ROWS_HEAD_PREV = 2, ROWS_ZERO = 3 };
enum { PE_NONE = 0, PE_CHUNK_POS = 1, PE_ROW0 = 2, PE_ROW_N = 3 };

// ---------------------------------------------------------------------------------------------
// epilogue activations.  __f*_rn keep the reference's op order (no FMA contraction).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float snake_f(float v, float alpha) {
  // dac Snake1d: x + (alpha + 1e-9)^-1 * sin(alpha x)^2
  float inv = __frcp_rn(__fadd_rn(alpha, 1e-9f));
  float s = sinf(__fmul_rn(alpha, v));
  return __fadd_rn(v, __fmul_rn(inv, __fmul_rn(s, s)));
}
// sin(t)^2 for the tensor-core epilogues: t - k*pi by a 3-term Cody-Waite split, then an odd degree-9
// polynomial on [-pi/2, pi/2] (the sign of sin does not matter).  Max abs error 2.6e-7 for |t| <= 60
// (fp32 sinf squared: 1.2e-7), ~14 instructions instead of ~45.
__device__ __forceinline__ float sin2_fast(float t) {
  const float kf = fmaf(t, 0.31830987334251404f, 12582912.f) - 12582912.f;   // rint(t / pi)
  float r = fmaf(kf, -3.140625f, t);
  r = fmaf(kf, -9.67502593994140625e-4f, r);
  r = fmaf(kf, -1.509957990978376e-07f, r);
  const float u = r * r;
  float q = fmaf(u, 2.600054131107754e-06f, -0.00019806614727713168f);
  q = fmaf(u, q, 0.008333017118275166f);
  q = fmaf(u, q, -0.16666656732559204f);
  const float sn = fmaf(u * r, q, r);
  return sn * sn;
}
// snake with the reciprocal 1/(alpha + 1e-9) precomputed per channel
__device__ __forceinline__ float snake_fast(float v, float alpha, float inv) { return fmaf(inv, sin2_fast(alpha * v), v); }
// the same on the SFU (MUFU.SIN, abs error ~1e-6): for epilogues whose output is rounded to ONE bf16 plane (8 mantissa
// bits) anyway -- the single-pass bf16 decoder, downstream of the quantizer.  5 instructions instead of 14.
__device__ __forceinline__ float snake_sfu(float v, float alpha, float inv) {
  const float s = __sinf(alpha * v);
  return fmaf(inv, s * s, v);
}
// SFU = 1 only where the activation is stored as a single bf16 plane
#ifndef B2C_SNAKE_SFU_X3
#define B2C_SNAKE_SFU_X3 1     // 1 = MUFU.SIN snake in the bf16x3 epilogues too (0: the polynomial)
#endif
template <int SFU>
__device__ __forceinline__ float snake_sel(float v, float alpha, float inv) {
  return (SFU || B2C_SNAKE_SFU_X3) ? snake_sfu(v, alpha, inv) : snake_fast(v, alpha, inv);
}

// d/dx snake(x) = 1 + sin(2 alpha x) * alpha / (alpha + 1e-9): the factor of the decoder's backward-data pass
__device__ __forceinline__ float dsnake_f(float v, float alpha) {
  return fmaf(sinf(2.0f * alpha * v), alpha * __frcp_rn(__fadd_rn(alpha, 1e-9f)), 1.0f);
}

__device__ __forceinline__ float gelu_f(float v) {
  // nn.GELU() (erf form): 0.5 x (1 + erf(x / sqrt 2))
  return __fmul_rn(__fmul_rn(v, 0.5f), __fadd_rn(1.0f, erff(__fmul_rn(v, 0.70710678118654752440f))));
}
__device__ __forceinline__ float apply_act(float v, int act, float alpha) {
  if (act == ACT_SNAKE) return snake_f(v, alpha);
  if (act == ACT_GELU) return gelu_f(v);
  if (act == ACT_TANH) return tanhf(v);
  return v;
}

// ---------------------------------------------------------------------------------------------
// Generic Conv1d / Linear / ConvTranspose1d-phase implicit GEMM, FP32 FFMA.
//   out[b, j*out_step + out_off[ph], co] = bias[co] + sum_{t<KT} sum_ci
//        x[b, j*in_step + t*dil + in_off[ph], ci] * w[ph][t][ci][co]      (+ res)
// M = positions (128 per CTA), N = output channels (16*TN per CTA), K = KT*Cin in slices of 16.
// ---------------------------------------------------------------------------------------------
struct ConvArgs {
  const float* x;
  const float* w;
  const float* bias;
  const float* res;
  float* out_raw;
  float* out_act;
  const float* alpha;
  int B, Lin, Cin, Cout, KT;
  int in_step, dil;
  int n_phase, Lj;
  int out_step, Lout;
  int in_off[8];
  int out_off[8];
  int act, res_mode, Tl, chunk;
  // backward-data pass: the contraction (+ bias) is multiplied by snake'(dmul[out position]; alpha) BEFORE res is added
  const float* dmul;
};

template <int TN>
__global__ void __launch_bounds__(256, 2) conv_gemm_f32(const ConvArgs a) {
  constexpr int BM = 128, BK = 16, BN = 16 * TN;
  constexpr int VW = (TN == 6) ? 2 : 4;   // vector width of the per-thread channel groups
  constexpr int NG = TN / VW;             // channel groups per thread
  constexpr int NB4 = (4 * BN + 255) / 256;  // float4 loads of the weight tile per thread
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int b = blockIdx.z / a.n_phase, ph = blockIdx.z % a.n_phase;
  const int j0 = blockIdx.x * BM, co0 = blockIdx.y * BN;
  const float* __restrict__ wp = a.w + (size_t)ph * a.KT * a.Cin * a.Cout;
  const float* __restrict__ xb = a.x + (size_t)b * a.Lin * a.Cin;
  const int in_off = a.in_off[ph];

  // A-tile load mapping: two float4 per thread
  int a_m[2], a_kq[2], a_l0[2];
  bool a_ok[2];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    int f = tid + u * 256;
    a_m[u] = f >> 2;
    a_kq[u] = f & 3;
    int j = j0 + a_m[u];
    a_ok[u] = j < a.Lj;
    a_l0[u] = j * a.in_step + in_off;
  }
  const int nK = a.Cin / BK;
  const int nIt = a.KT * nK;

  float4 ra[2];
  float4 rb[NB4];

  auto load_g = [&](int it) {
    int tap = it / nK;
    int ci0 = (it - tap * nK) * BK;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      int l = a_l0[u] + tap * a.dil;
      if (a_ok[u] && l >= 0 && l < a.Lin)
        ra[u] = __ldg(reinterpret_cast<const float4*>(xb + (size_t)l * a.Cin + ci0 + a_kq[u] * 4));
      else
        ra[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < NB4; ++u) {
      int f = tid + u * 256;
      int kk = f / (BN / 4), nq = f % (BN / 4);
      int co = co0 + nq * 4;
      if (f < 4 * BN && co < a.Cout)
        rb[u] = __ldg(reinterpret_cast<const float4*>(wp + ((size_t)tap * a.Cin + ci0 + kk) * a.Cout + co));
      else
        rb[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto store_s = [&](int buf) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      As[buf][a_kq[u] * 4 + 0][a_m[u]] = ra[u].x;
      As[buf][a_kq[u] * 4 + 1][a_m[u]] = ra[u].y;
      As[buf][a_kq[u] * 4 + 2][a_m[u]] = ra[u].z;
      As[buf][a_kq[u] * 4 + 3][a_m[u]] = ra[u].w;
    }
#pragma unroll
    for (int u = 0; u < NB4; ++u) {
      int f = tid + u * 256;
      if (f < 4 * BN) {
        int kk = f / (BN / 4), nq = f % (BN / 4);
        *reinterpret_cast<float4*>(&Bs[buf][kk][nq * 4]) = rb[u];
      }
    }
  };

  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int n = 0; n < TN; ++n) acc[i][n] = 0.f;

  load_g(0);
  store_s(0);
  __syncthreads();
  for (int it = 0; it < nIt; ++it) {
    const int buf = it & 1;
    if (it + 1 < nIt) load_g(it + 1);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float av[8], bv[TN];
      *reinterpret_cast<float4*>(&av[0]) = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8]);
      *reinterpret_cast<float4*>(&av[4]) = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8 + 4]);
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        if constexpr (VW == 4)
          *reinterpret_cast<float4*>(&bv[g * 4]) =
              *reinterpret_cast<const float4*>(&Bs[buf][k][g * 64 + tx * 4]);
        else
          *reinterpret_cast<float2*>(&bv[g * 2]) =
              *reinterpret_cast<const float2*>(&Bs[buf][k][g * 32 + tx * 2]);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int n = 0; n < TN; ++n) acc[i][n] = fmaf(av[i], bv[n], acc[i][n]);
    }
    if (it + 1 < nIt) store_s(buf ^ 1);
    __syncthreads();
  }

  // epilogue
  const int out_off = a.out_off[ph];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int j = j0 + ty * 8 + i;
    if (j >= a.Lj) continue;
    int lo = j * a.out_step + out_off;
    if (lo < 0 || lo >= a.Lout) continue;
    size_t orow = ((size_t)b * a.Lout + lo) * a.Cout;
    size_t rrow = orow;
    if (a.res_mode == 1) rrow = (size_t)((lo % a.Tl) % a.chunk) * a.Cout;
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      int co = co0 + g * (16 * VW) + tx * VW;
      if (co >= a.Cout) continue;
      float v[VW], w[VW];
#pragma unroll
      for (int e = 0; e < VW; ++e) {
        float t = acc[i][g * VW + e];
        if (a.bias) t = __fadd_rn(t, __ldg(a.bias + co + e));
        if (a.dmul) t = __fmul_rn(t, dsnake_f(__ldg(a.dmul + orow + co + e), __ldg(a.alpha + co + e)));
        if (a.res) t = __fadd_rn(t, __ldg(a.res + rrow + co + e));
        v[e] = t;
      }
      if (a.out_raw) {
        if constexpr (VW == 4)
          *reinterpret_cast<float4*>(a.out_raw + orow + co) = make_float4(v[0], v[1], v[2], v[3]);
        else
          *reinterpret_cast<float2*>(a.out_raw + orow + co) = make_float2(v[0], v[1]);
      }
      if (a.out_act) {
#pragma unroll
        for (int e = 0; e < VW; ++e)
          w[e] = apply_act(v[e], a.act, a.act == ACT_SNAKE ? __ldg(a.alpha + co + e) : 0.f);
        if constexpr (VW == 4)
          *reinterpret_cast<float4*>(a.out_act + orow + co) = make_float4(w[0], w[1], w[2], w[3]);
        else
          *reinterpret_cast<float2*>(a.out_act + orow + co) = make_float2(w[0], w[1]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Encoder stem: Conv1d(1, Cout, k=7, p=3).  HBM-bound on the store (Cout floats per sample).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) stem_k7_f32(const float* __restrict__ x, const float* __restrict__ w,
                                                    const float* __restrict__ bias, float* __restrict__ out_raw,
                                                    void* __restrict__ out_act, const float* __restrict__ alpha,
                                                    int L, int Cout, int act, int act_fmt, size_t act_n) {
  constexpr int TP = 64;
  extern __shared__ float sm[];
  float* xs = sm;             // TP + 6
  float* ws = sm + TP + 8;    // 7 * Cout, then bias, then alpha
  const int b = blockIdx.y, l0 = blockIdx.x * TP;
  for (int i = threadIdx.x; i < TP + 6; i += blockDim.x) {
    int l = l0 + i - 3;
    xs[i] = (l >= 0 && l < L) ? __ldg(x + (size_t)b * L + l) : 0.f;
  }
  for (int i = threadIdx.x; i < 7 * Cout; i += blockDim.x) ws[i] = __ldg(w + i);
  __syncthreads();
  for (int idx = threadIdx.x; idx < TP * Cout; idx += blockDim.x) {
    int p = idx / Cout, co = idx - p * Cout;
    int l = l0 + p;
    if (l >= L) break;
    float acc = 0.f;
#pragma unroll
    for (int t = 0; t < 7; ++t) acc = fmaf(xs[p + t], ws[t * Cout + co], acc);
    if (bias) acc = __fadd_rn(acc, __ldg(bias + co));
    size_t o = ((size_t)b * L + l) * Cout + co;
    if (out_raw) out_raw[o] = acc;
    if (out_act) store_fmt(out_act, act_fmt, act_n, o, apply_act(acc, act, act == ACT_SNAKE ? __ldg(alpha + co) : 0.f));
  }
}

// Same stem for the tensor-core plans: 4 channels per thread, float4 raw stores, bf16x2-packed plane stores,
// polynomial snake.  Cout % 4 == 0.
__global__ void __launch_bounds__(256) stem_k7_planes(const float* __restrict__ x, const float* __restrict__ w,
                                                       const float* __restrict__ bias, float* __restrict__ out_raw,
                                                       __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo,
                                                       const float* __restrict__ alpha, const float* __restrict__ inv_alpha,
                                                       int L, int Cout) {
  constexpr int TP = 64;
  extern __shared__ float sm[];
  float* xs = sm;             // TP + 6
  float* ws = sm + TP + 8;    // 7 * Cout
  const int b = blockIdx.y, l0 = blockIdx.x * TP;
  for (int i = threadIdx.x; i < TP + 6; i += blockDim.x) {
    int l = l0 + i - 3;
    xs[i] = (l >= 0 && l < L) ? __ldg(x + (size_t)b * L + l) : 0.f;
  }
  for (int i = threadIdx.x; i < 7 * Cout; i += blockDim.x) ws[i] = __ldg(w + i);
  __syncthreads();
  const int ng = Cout >> 2;
  for (int idx = threadIdx.x; idx < TP * ng; idx += blockDim.x) {
    const int pp = idx / ng, co = (idx - pp * ng) * 4;
    const int l = l0 + pp;
    if (l >= L) break;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int t = 0; t < 7; ++t) {
      const float xv = xs[pp + t];
      const float4 wv = *reinterpret_cast<const float4*>(ws + t * Cout + co);
      a.x = fmaf(xv, wv.x, a.x); a.y = fmaf(xv, wv.y, a.y); a.z = fmaf(xv, wv.z, a.z); a.w = fmaf(xv, wv.w, a.w);
    }
    if (bias) {
      const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + co));
      a.x += bb.x; a.y += bb.y; a.z += bb.z; a.w += bb.w;
    }
    const size_t o = ((size_t)b * L + l) * Cout + co;
    if (out_raw) *reinterpret_cast<float4*>(out_raw + o) = a;
    if (out_hi) {
      const float4 al = __ldg(reinterpret_cast<const float4*>(alpha + co));
      const float4 ia = __ldg(reinterpret_cast<const float4*>(inv_alpha + co));
      float4 v;
      v.x = snake_fast(a.x, al.x, ia.x); v.y = snake_fast(a.y, al.y, ia.y);
      v.z = snake_fast(a.z, al.z, ia.z); v.w = snake_fast(a.w, al.w, ia.w);
      const __nv_bfloat162 h01 = __floats2bfloat162_rn(v.x, v.y), h23 = __floats2bfloat162_rn(v.z, v.w);
      *reinterpret_cast<uint2*>(out_hi + o) =
          make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
      if (out_lo) {
        const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
        const __nv_bfloat162 l01 = __floats2bfloat162_rn(v.x - f01.x, v.y - f01.y);
        const __nv_bfloat162 l23 = __floats2bfloat162_rn(v.z - f23.x, v.w - f23.y);
        *reinterpret_cast<uint2*>(out_lo + o) =
            make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
      }
    }
  }
}

// Decoder head, sliding window (tensor-core plans, Cin = 32 * UPT): the CTA stages 128 + 6 channel-last rows in shared
// memory as fp32; thread = (16 segments of 8 positions) x (8 channel groups).  A thread keeps the 7 x (4 * UPT) weights
// of its channel group in registers and walks the 14 rows of its segment: each row is loaded once (UPT float4 reads,
// the 8 lanes of a quarter-warp read 8 different 16-byte units: no bank conflicts) and feeds up to 7 outputs, so
// shared-memory reads drop 8x against the one-position-per-thread form, which was bound by them (0.37 ms per 64
// frames for 2 x 6.1 GFLOP).  The 8 channel-group partials of a position are lanes g = 0..7: three shuffles.
template <int FMT, int UPT>
__global__ void __launch_bounds__(128) head_k7_tanh_sw(const void* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ bias, float* __restrict__ y, int L,
                                                        size_t x_n) {
  constexpr int TP = 128, Cin = 32 * UPT, ld = Cin + 4;
  extern __shared__ __align__(16) float hsm[];
  float* xs = hsm;                    // [TP + 6][ld]
  const int b = blockIdx.y, l0 = blockIdx.x * TP;
  const int g = threadIdx.x & 7, seg = threadIdx.x >> 3;
  float wr[7][4 * UPT];
#pragma unroll
  for (int r = 0; r < 7; ++r)
#pragma unroll
    for (int u = 0; u < UPT; ++u) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(w + r * Cin + 4 * (g + 8 * u)));
      wr[r][4 * u] = v.x; wr[r][4 * u + 1] = v.y; wr[r][4 * u + 2] = v.z; wr[r][4 * u + 3] = v.w;
    }
  constexpr int vec_per_row = Cin >> 3;
  for (int v = threadIdx.x; v < (TP + 6) * vec_per_row; v += 128) {
    const int r = v / vec_per_row, c = (v - r * vec_per_row) * 8;
    const int l = l0 + r - 3;
    float f[8];
    if (l >= 0 && l < L) {
      const size_t e = ((size_t)b * L + l) * Cin + c;
      const __nv_bfloat16* xb = reinterpret_cast<const __nv_bfloat16*>(x);
      const uint4 h = __ldg(reinterpret_cast<const uint4*>(xb + e));
      const __nv_bfloat162* hp = reinterpret_cast<const __nv_bfloat162*>(&h);
#pragma unroll
      for (int q = 0; q < 4; ++q) { const float2 t2 = __bfloat1622float2(hp[q]); f[2 * q] = t2.x; f[2 * q + 1] = t2.y; }
      if (FMT == FMT_PLANES) {
        const uint4 lo = __ldg(reinterpret_cast<const uint4*>(xb + x_n + e));
        const __nv_bfloat162* lp = reinterpret_cast<const __nv_bfloat162*>(&lo);
#pragma unroll
        for (int q = 0; q < 4; ++q) { const float2 t2 = __bfloat1622float2(lp[q]); f[2 * q] += t2.x; f[2 * q + 1] += t2.y; }
      }
    } else {
#pragma unroll
      for (int q = 0; q < 8; ++q) f[q] = 0.f;
    }
    *reinterpret_cast<float4*>(xs + r * ld + c) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4*>(xs + r * ld + c + 4) = make_float4(f[4], f[5], f[6], f[7]);
  }
  __syncthreads();
  float acc[8];
#pragma unroll
  for (int o = 0; o < 8; ++o) acc[o] = 0.f;
#pragma unroll
  for (int j = 0; j < 14; ++j) {            // staged row seg * 8 + j = input position l0 + seg * 8 + j - 3
    const float* xr = xs + (seg * 8 + j) * ld;
    float xv[4 * UPT];
#pragma unroll
    for (int u = 0; u < UPT; ++u) {
      const float4 v = *reinterpret_cast<const float4*>(xr + 4 * (g + 8 * u));
      xv[4 * u] = v.x; xv[4 * u + 1] = v.y; xv[4 * u + 2] = v.z; xv[4 * u + 3] = v.w;
    }
#pragma unroll
    for (int r = 0; r < 7; ++r) {
      const int o = j - r;                  // output seg * 8 + o uses this row with tap r
      if (o >= 0 && o < 8) {
#pragma unroll
        for (int c = 0; c < 4 * UPT; ++c) acc[o] = fmaf(xv[c], wr[r][c], acc[o]);
      }
    }
  }
  const float bv = bias ? __ldg(bias) : 0.f;
#pragma unroll
  for (int o = 0; o < 8; ++o) {
    float v = acc[o];
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    acc[o] = v;
  }
  // lane g of each group of 8 writes output o = g: one coalesced 32-byte store per segment
  float mine = acc[0];
#pragma unroll
  for (int o = 1; o < 8; ++o) mine = (g == o) ? acc[o] : mine;
  const int l = l0 + seg * 8 + g;
  if (l < L) y[(size_t)b * L + l] = tanhf(__fadd_rn(mine, bv));
}

// Decoder head, tiled: a CTA stages 128 + 6 rows of the channel-last input (contiguous in memory) in shared
// memory as fp32 with a padded row (Cin + 4 words: conflict-free float4 reads), thread t computes position t.
// Cin % 8 == 0.  Reads each input element once from HBM.
template <int FMT>
__global__ void __launch_bounds__(128) head_k7_tanh_tiled(const void* __restrict__ x, const float* __restrict__ w,
                                                           const float* __restrict__ bias, float* __restrict__ y, int L,
                                                           int Cin, size_t x_n) {
  constexpr int TP = 128;
  extern __shared__ __align__(16) float hsm[];
  const int ld = Cin + 4;
  float* xs = hsm;                    // [TP + 6][ld]
  float* ws = hsm + (TP + 6) * ld;    // [7][Cin]
  const int b = blockIdx.y, l0 = blockIdx.x * TP;
  for (int i = threadIdx.x; i < 7 * Cin; i += blockDim.x) ws[i] = __ldg(w + i);
  const int vec_per_row = Cin >> 3;
  for (int v = threadIdx.x; v < (TP + 6) * vec_per_row; v += blockDim.x) {
    const int r = v / vec_per_row, c = (v - r * vec_per_row) * 8;
    const int l = l0 + r - 3;
    float f[8];
    if (l >= 0 && l < L) {
      const size_t e = ((size_t)b * L + l) * Cin + c;
      if (FMT == FMT_F32) {
        const float4 a0 = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(x) + e));
        const float4 a1 = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(x) + e + 4));
        f[0] = a0.x; f[1] = a0.y; f[2] = a0.z; f[3] = a0.w; f[4] = a1.x; f[5] = a1.y; f[6] = a1.z; f[7] = a1.w;
      } else {
        const __nv_bfloat16* xb = reinterpret_cast<const __nv_bfloat16*>(x);
        const uint4 h = __ldg(reinterpret_cast<const uint4*>(xb + e));
        const __nv_bfloat162* hp = reinterpret_cast<const __nv_bfloat162*>(&h);
#pragma unroll
        for (int q = 0; q < 4; ++q) { const float2 t2 = __bfloat1622float2(hp[q]); f[2 * q] = t2.x; f[2 * q + 1] = t2.y; }
        if (FMT == FMT_PLANES) {
          const uint4 lo = __ldg(reinterpret_cast<const uint4*>(xb + x_n + e));
          const __nv_bfloat162* lp = reinterpret_cast<const __nv_bfloat162*>(&lo);
#pragma unroll
          for (int q = 0; q < 4; ++q) { const float2 t2 = __bfloat1622float2(lp[q]); f[2 * q] += t2.x; f[2 * q + 1] += t2.y; }
        }
      }
    } else {
#pragma unroll
      for (int q = 0; q < 8; ++q) f[q] = 0.f;
    }
    *reinterpret_cast<float4*>(xs + r * ld + c) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4*>(xs + r * ld + c + 4) = make_float4(f[4], f[5], f[6], f[7]);
  }
  __syncthreads();
  const int l = l0 + threadIdx.x;
  if (l >= L) return;
  // same summation order as the one-warp-per-output kernel is not required: the head is downstream of the
  // quantizer; four partial sums keep the FMA chains short
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  for (int r = 0; r < 7; ++r) {
    const float* xr = xs + (threadIdx.x + r) * ld;
    const float* wr = ws + r * Cin;
    for (int c = 0; c < Cin; c += 4) {
      const float4 xv = *reinterpret_cast<const float4*>(xr + c);
      const float4 wv = *reinterpret_cast<const float4*>(wr + c);
      a0 = fmaf(xv.x, wv.x, a0); a1 = fmaf(xv.y, wv.y, a1); a2 = fmaf(xv.z, wv.z, a2); a3 = fmaf(xv.w, wv.w, a3);
    }
  }
  y[(size_t)b * L + l] = tanhf(__fadd_rn((a0 + a1) + (a2 + a3), bias ? __ldg(bias) : 0.f));
}

// ---------------------------------------------------------------------------------------------
// Decoder head: Conv1d(Cin, 1, k=7, p=3) + tanh.  In channel-last layout the 7*Cin window of one
// output is CONTIGUOUS, so y[l] = tanh(b + <w_flat, x_flat[(l-3)*Cin ...]>): one warp per output,
// coalesced reads, HBM/L1-bound.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) head_k7_tanh_f32(const void* __restrict__ x, const float* __restrict__ w,
                                                         const float* __restrict__ bias, float* __restrict__ y, int L,
                                                         int Cin, int x_fmt, size_t x_n) {
  extern __shared__ float wsm[];
  const int n = 7 * Cin;
  for (int i = threadIdx.x; i < n; i += blockDim.x) wsm[i] = __ldg(w + i);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const size_t xoff = (size_t)b * L * Cin;
  const long total = (long)L * Cin;
  constexpr int PPW = 8;
  const int l_base = (blockIdx.x * 8 + warp) * PPW;
  for (int p = 0; p < PPW; ++p) {
    int l = l_base + p;
    if (l >= L) break;
    long base = (long)(l - 3) * Cin;
    float acc = 0.f;
    for (int f = lane; f < n; f += 32) {
      long g = base + f;
      float xv = (g >= 0 && g < total) ? load_fmt(x, x_fmt, x_n, xoff + g) : 0.f;
      acc = fmaf(xv, wsm[f], acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) y[(size_t)b * L + l] = tanhf(__fadd_rn(acc, bias ? __ldg(bias) : 0.f));
  }
}

// ---------------------------------------------------------------------------------------------
// Row LayerNorm with fused input composition: v = a[row(n)] - sub[n] + pe[pos(n)];
// out[n] = LN(v) * gamma + beta; optional out = post_scale * tanh(out).   One warp per row.
// ---------------------------------------------------------------------------------------------
struct LnArgs {
  const float* a;
  const float* sub;
  const float* pe;
  const float* gamma;
  const float* beta;
  void* out;
  int N, C, Tl, chunk, nfix;
  int a_mode, pe_mode, tanh_post;
  float post_scale;
  int out_fmt;
  const unsigned char* rmask;   // optional [N]: rows with a non-zero byte read a as zeros (PLC: zt * ~mask, PLC1_eval.py:491)
};

__device__ __forceinline__ long gather_row(int n, int mode, int Tl, int chunk, int nfix) {
  if (mode == ROWS_DENSE) return n;
  int b = n / nfix, j = n - b * nfix;
  long r = (long)b * Tl + (long)chunk * (j + 1);
  return mode == ROWS_HEAD ? r : r - 1;
}

__global__ void __launch_bounds__(256) layernorm_rows_f32(const LnArgs p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * 8 + warp;
  if (n >= p.N) return;
  const int C = p.C;
  const float* ar = nullptr;
  if (p.a_mode != ROWS_ZERO && !(p.rmask && p.rmask[n])) ar = p.a + gather_row(n, p.a_mode, p.Tl, p.chunk, p.nfix) * (long)C;
  const float* sr = p.sub ? p.sub + (long)n * C : nullptr;
  const float* pr = nullptr;
  if (p.pe_mode == PE_CHUNK_POS) pr = p.pe + (long)((n % p.Tl) % p.chunk) * C;
  else if (p.pe_mode == PE_ROW0) pr = p.pe;
  else if (p.pe_mode == PE_ROW_N) pr = p.pe + (long)n * C;

  auto val = [&](int c) -> float {
    float v = ar ? __ldg(ar + c) : 0.f;
    if (sr) v = __fsub_rn(v, __ldg(sr + c));
    if (pr) v = __fadd_rn(v, __ldg(pr + c));
    return v;
  };
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += val(c);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)C;
  float q = 0.f;
  for (int c = lane; c < C; c += 32) {
    float d = val(c) - mean;
    q = fmaf(d, d, q);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = 1.0f / sqrtf(q / (float)C + 1e-5f);
  const float nb = -rstd * mean;
  for (int c = lane; c < C; c += 32) {
    float v = __fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(val(c), rstd), nb), __ldg(p.gamma + c)), __ldg(p.beta + c));
    if (p.tanh_post) v = __fmul_rn(p.post_scale, tanhf(v));
    store_fmt(p.out, p.out_fmt, (size_t)p.N * C, (size_t)n * C + c, v);
  }
}

// ---------------------------------------------------------------------------------------------
// Chunk-local multi-head attention (head dim 128, <= 16 keys per chunk).  One warp per head.
// ---------------------------------------------------------------------------------------------
struct AttnArgs {
  const float* q;   // q_mode 0: [chunk, C] table; 1: [B*nfix, C]; 2: dense [B*Tl, C]
  const float* kv;  // [B*Tl, 2C]
  float* out;       // q_mode 0: [B*Tl, C]; 1: [B*nfix, C]
  int B, Tl, chunk, heads, nchunks, nfix, q_mode;
};

__global__ void __launch_bounds__(256) attention_chunk_f32(const AttnArgs p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp >= p.heads) return;
  const int C = p.heads * 128;
  // gridDim.y > 1 (q_mode != 1, large batches): the queries of a chunk are split over gridDim.y CTAs -- a warp walks
  // its queries serially (16 dot products x 5 shuffles each), so shorter walks are what shortens the kernel
  int b, ck;
  if (p.q_mode != 1) { b = blockIdx.x / p.nchunks; ck = blockIdx.x % p.nchunks; }
  else { b = blockIdx.x / p.nfix; ck = blockIdx.x % p.nfix + 1; }
  const int s = ck * p.chunk;
  const int tk = min(p.Tl, s + p.chunk) - s;
  const int hoff = warp * 128 + lane * 4;
  float4 kr[16], vr[16];
#pragma unroll
  for (int l = 0; l < 16; ++l) {
    if (l < tk) {
      const float* row = p.kv + ((long)b * p.Tl + s + l) * (2 * C);
      kr[l] = __ldg(reinterpret_cast<const float4*>(row + hoff));
      vr[l] = __ldg(reinterpret_cast<const float4*>(row + C + hoff));
    } else {
      kr[l] = make_float4(0, 0, 0, 0);
      vr[l] = make_float4(0, 0, 0, 0);
    }
  }
  const int nq = p.q_mode != 1 ? tk : 1;
  const float inv_sqrt_dh = 11.313708498984761f;  // sqrt(128): the reference divides by it (:401)
  for (int i = blockIdx.y; i < nq; i += gridDim.y) {
    const float* qrow = p.q_mode == 0 ? p.q + (long)i * C
                        : p.q_mode == 1 ? p.q + (long)blockIdx.x * C
                                        : p.q + ((long)b * p.Tl + s + i) * C;
    float4 q4 = __ldg(reinterpret_cast<const float4*>(qrow + hoff));
    float sc[16];
    float mx = -INFINITY;
#pragma unroll
    for (int l = 0; l < 16; ++l) {
      float d = q4.x * kr[l].x;
      d = fmaf(q4.y, kr[l].y, d);
      d = fmaf(q4.z, kr[l].z, d);
      d = fmaf(q4.w, kr[l].w, d);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
      sc[l] = __fdiv_rn(d, inv_sqrt_dh);
      if (l < tk) mx = fmaxf(mx, sc[l]);
    }
    float sum = 0.f;
#pragma unroll
    for (int l = 0; l < 16; ++l) {
      sc[l] = (l < tk) ? expf(sc[l] - mx) : 0.f;
      sum += sc[l];
    }
    float4 o4 = make_float4(0, 0, 0, 0);
#pragma unroll
    for (int l = 0; l < 16; ++l) {
      float pl = __fdiv_rn(sc[l], sum);
      o4.x = fmaf(pl, vr[l].x, o4.x);
      o4.y = fmaf(pl, vr[l].y, o4.y);
      o4.z = fmaf(pl, vr[l].z, o4.z);
      o4.w = fmaf(pl, vr[l].w, o4.w);
    }
    long orow = p.q_mode != 1 ? ((long)b * p.Tl + s + i) : (long)blockIdx.x;
    *reinterpret_cast<float4*>(p.out + orow * C + hoff) = o4;
  }
}

// ---------------------------------------------------------------------------------------------
// Residual VQ (ResidualVQEMA.forward) -- FP32 CUDA-core version.  32 tokens per CTA, codebooks
// streamed through shared memory 64 codes at a time; score = x.e - 0.5|e|^2, first maximum wins.
// ---------------------------------------------------------------------------------------------
struct RvqArgs {
  const float* x;       // [N, D]
  const float* books;   // [n_books][K][D]
  const float* half_n;  // [n_books][K]   0.5*|e|^2
  float* qsum;          // [N, D] or null
  int* idx;             // [B, books_use, Tl] or (single-book nearest) [N]
  int N, D, K, books_use;
  int row_mode, B, Tl, chunk, nfix;
  int idx_flat;         // 1: idx[n] (nearest op)
  int lookup;           // 1: receiver side -- idx is an INPUT, qsum = sum over books of book[idx] (b2c_prog_rvq_lookup)
};

// TPW tokens per warp (8 * TPW per CTA): 4 for large N, 1 when that would leave most SMs idle.
template <int TPW>
__global__ void __launch_bounds__(256) rvq_f32(const RvqArgs p) {
  constexpr int CH = 64, TOK = 8 * TPW;
  extern __shared__ float sm[];
  const int D = p.D, DP = D + 1;
  float* xs = sm;                   // [TOK][D]  residual
  float* qs = xs + TOK * D;         // [TOK][D]  q_sum
  float* es = qs + TOK * D;         // [CH][D+1]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * TOK;
  for (int i = threadIdx.x; i < TOK * D; i += 256) {
    int t = i / D;
    int n = n0 + t;
    xs[i] = n < p.N ? __ldg(p.x + (long)n0 * D + i) : 0.f;
    qs[i] = 0.f;
  }
  __syncthreads();
  for (int bk = 0; bk < p.books_use; ++bk) {
    const float* book = p.books + (long)bk * p.K * D;
    const float* hn = p.half_n + (long)bk * p.K;
    float best[TPW];
    int bidx[TPW];
#pragma unroll
    for (int t = 0; t < TPW; ++t) { best[t] = -INFINITY; bidx[t] = 0x7fffffff; }
    for (int c0 = 0; c0 < p.K; c0 += CH) {
      __syncthreads();
      for (int i = threadIdx.x; i < CH * D; i += 256) {
        int r = i / D, d = i - r * D;
        es[r * DP + d] = (c0 + r < p.K) ? __ldg(book + (long)c0 * D + i) : 0.f;
      }
      __syncthreads();
      float acc[TPW][2];
#pragma unroll
      for (int t = 0; t < TPW; ++t) acc[t][0] = acc[t][1] = 0.f;
      const float* e0 = es + lane * DP;
      const float* e1 = es + (lane + 32) * DP;
      const float* xw = xs + warp * TPW * D;
      for (int d = 0; d < D; ++d) {
        float ev0 = e0[d], ev1 = e1[d];
#pragma unroll
        for (int t = 0; t < TPW; ++t) {
          float xv = xw[t * D + d];
          acc[t][0] = fmaf(xv, ev0, acc[t][0]);
          acc[t][1] = fmaf(xv, ev1, acc[t][1]);
        }
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        int code = c0 + lane + 32 * h;
        if (code < p.K) {
          float hv = __ldg(hn + code);
#pragma unroll
          for (int t = 0; t < TPW; ++t) {
            float s = __fsub_rn(acc[t][h], hv);
            if (s > best[t]) { best[t] = s; bidx[t] = code; }
          }
        }
      }
    }
    // warp arg-max, lowest index wins ties (torch.argmax)
#pragma unroll
    for (int t = 0; t < TPW; ++t) {
      float bs = best[t];
      int bi = bidx[t];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        float os = __shfl_xor_sync(0xffffffffu, bs, o);
        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (os > bs || (os == bs && oi < bi)) { bs = os; bi = oi; }
      }
      const int tl = warp * TPW + t;
      const int n = n0 + tl;
      if (n < p.N) {
        if (bi >= p.K) bi = 0;
        const float* e = book + (long)bi * D;
        for (int d = lane; d < D; d += 32) {
          float q = __ldg(e + d);
          float r = xs[tl * D + d];
          // q_sum = q_sum + (q - residual) + residual ; residual = residual - q   (:433-434)
          qs[tl * D + d] = __fadd_rn(__fadd_rn(qs[tl * D + d], __fsub_rn(q, r)), r);
          xs[tl * D + d] = __fsub_rn(r, q);
        }
        if (lane == 0) {
          if (p.idx_flat) p.idx[n] = bi;
          else {
            int b, tt;
            if (p.row_mode == ROWS_DENSE) { b = n / p.Tl; tt = n - b * p.Tl; }
            else { b = n / p.nfix; tt = p.chunk * (n - b * p.nfix + 1); }
            p.idx[((long)b * p.books_use + bk) * p.Tl + tt] = bi;
          }
        }
      }
    }
    __syncwarp();
  }
  __syncthreads();
  if (p.qsum)
    for (int i = threadIdx.x; i < TOK * D; i += 256)
      if (n0 + i / D < p.N) p.qsum[(long)n0 * D + i] = qs[i];
}

// ---------------------------------------------------------------------------------------------
// Residual VQ, split form: per book one score launch over (32-token block) x (64-code slice) CTAs -- every SM works
// even for a handful of tokens -- and one apply launch.  A token's winner is reduced across the code slices with
// a 64-bit atomicMax on (order-preserving score bits << 32 | ~index): highest score, ties -> lowest index, i.e.
// torch.argmax's first maximum.  Scores use rvq_f32's arithmetic (sequential fmaf over d, minus 0.5|e|^2), so the
// indices are bit-identical to the fused kernel's.  The batch-1 streaming path drops from ~0.6 ms to < 0.1 ms.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
               "l"(gmem_src)
               : "memory");
}

__device__ __forceinline__ unsigned long long rvq_key(float score, int idx) {
  unsigned int u = __float_as_uint(score);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);          // monotone map float -> uint
  return ((unsigned long long)u << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned int)idx);
}

// CPT codes per lane (CH = 32 * CPT codes per CTA) x 4 tokens per warp.  With D % 4 == 0 the shared-memory tiles are
// read as float4 along d: 4 + CPT 16-byte loads feed 16 * CPT FMAs (the 2-code scalar form was bound by its 6 loads
// per 8 FMAs).  Every (token, code) dot product is one fmaf chain over d = 0 .. D-1 in either form: identical bits.
template <int CPT>
__global__ void __launch_bounds__(256) rvq_scores_f32(const float* __restrict__ res, const float* __restrict__ book,
                                                      const float* __restrict__ hn, unsigned long long* __restrict__ keys,
                                                      int N, int D, int K) {
  constexpr int TPW = 4, CH = 32 * CPT, TOK = 32;
  extern __shared__ __align__(16) float sm[];
  const bool vec = (D & 3) == 0;
  const int DP = vec ? D + 4 : D + 1;   // row pitch of the code tile: conflict-free for float4 / scalar reads
  float* xs = sm;               // [TOK][D]
  float* es = sm + TOK * D;     // [CH][DP]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * TOK, c0 = blockIdx.y * CH;
  for (int i = threadIdx.x; i < TOK * D; i += 256) xs[i] = (n0 + i / D < N) ? __ldg(res + (long)n0 * D + i) : 0.f;
  for (int i = threadIdx.x; i < CH * D; i += 256) {
    const int r = i / D, d = i - r * D;
    es[r * DP + d] = (c0 + r < K) ? __ldg(book + (long)c0 * D + i) : 0.f;
  }
  __syncthreads();
  float acc[TPW][CPT];
#pragma unroll
  for (int t = 0; t < TPW; ++t)
#pragma unroll
    for (int h = 0; h < CPT; ++h) acc[t][h] = 0.f;
  const float* xw = xs + warp * TPW * D;
  if (vec) {
    for (int d = 0; d < D; d += 4) {
      float4 ev[CPT], xv[TPW];
#pragma unroll
      for (int h = 0; h < CPT; ++h) ev[h] = *reinterpret_cast<const float4*>(es + (lane + 32 * h) * DP + d);
#pragma unroll
      for (int t = 0; t < TPW; ++t) xv[t] = *reinterpret_cast<const float4*>(xw + t * D + d);
#pragma unroll
      for (int t = 0; t < TPW; ++t)
#pragma unroll
        for (int h = 0; h < CPT; ++h) {
          float a = acc[t][h];
          a = fmaf(xv[t].x, ev[h].x, a); a = fmaf(xv[t].y, ev[h].y, a);
          a = fmaf(xv[t].z, ev[h].z, a); a = fmaf(xv[t].w, ev[h].w, a);
          acc[t][h] = a;
        }
    }
  } else {
    for (int d = 0; d < D; ++d) {
      float ev[CPT];
#pragma unroll
      for (int h = 0; h < CPT; ++h) ev[h] = es[(lane + 32 * h) * DP + d];
#pragma unroll
      for (int t = 0; t < TPW; ++t) {
        const float xv = xw[t * D + d];
#pragma unroll
        for (int h = 0; h < CPT; ++h) acc[t][h] = fmaf(xv, ev[h], acc[t][h]);
      }
    }
  }
  float best[TPW];
  int bidx[TPW];
#pragma unroll
  for (int t = 0; t < TPW; ++t) { best[t] = -INFINITY; bidx[t] = 0x7fffffff; }
#pragma unroll
  for (int h = 0; h < CPT; ++h) {
    const int code = c0 + lane + 32 * h;
    if (code < K) {
      const float hv = __ldg(hn + code);
#pragma unroll
      for (int t = 0; t < TPW; ++t) {
        const float sc = __fsub_rn(acc[t][h], hv);
        if (sc > best[t]) { best[t] = sc; bidx[t] = code; }
      }
    }
  }
#pragma unroll
  for (int t = 0; t < TPW; ++t) {
    float bs = best[t];
    int bi = bidx[t];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float os = __shfl_xor_sync(0xffffffffu, bs, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (os > bs || (os == bs && oi < bi)) { bs = os; bi = oi; }
    }
    const int n = n0 + warp * TPW + t;
    if (lane == 0 && n < N && bi < K) atomicMax(keys + n, rvq_key(bs, bi));
  }
}
// All books in ONE launch (large batches): a CTA owns 32 tokens (4 per warp) for the whole residual loop.  Per book the
// code slices (128 codes) stream through a double-buffered shared-memory tile (cp.async), every lane keeps the first
// maximum over its codes across the slices, one warp reduction per book gives the index, and the warp that owns a
// token applies q_sum / residual itself -- no grid-wide hand-off between the score and apply steps, 1 launch instead
// of 2 per book.  Scores, tie-breaking and the q_sum / residual op order are those of rvq_scores_f32 + rvq_apply_f32
// (identical bits); D % 4 == 0, D <= 128.  TPW tokens per warp: 4, or 2 when that is what gives every SM two CTAs.
template <int TPW>
__global__ void __launch_bounds__(256) rvq_books_f32(const RvqArgs p) {
  constexpr int CPT = 4, CH = 128, TOK = 8 * TPW;
  extern __shared__ __align__(16) float sm[];
  const int D = p.D, DP = D + 4, K = p.K;
  float* xs = sm;                       // [TOK][D] residual (rows of warp w are touched by warp w only)
  float* es0 = sm + TOK * D;            // 2 x [CH][DP]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * TOK;
  const int n_slices = (K + CH - 1) / CH;
  const int vec_row = D >> 2;           // 16-byte units per code row
  auto prefetch = [&](int bk, int sl, int buf) {
    const float* src = p.books + ((size_t)bk * K + (size_t)sl * CH) * D;
    float* dst = es0 + (size_t)buf * CH * DP;
    const int rows = min(CH, K - sl * CH);
    for (int i = threadIdx.x; i < rows * vec_row; i += 256) {
      const int r = i / vec_row, u = i - r * vec_row;
      cp_async16(dst + r * DP + 4 * u, src + (size_t)r * D + 4 * u);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  prefetch(0, 0, 0);
  for (int i = threadIdx.x; i < TOK * D; i += 256) xs[i] = (n0 + i / D < p.N) ? __ldg(p.x + (size_t)n0 * D + i) : 0.f;
  float qs[TPW][4];
#pragma unroll
  for (int t = 0; t < TPW; ++t)
#pragma unroll
    for (int j = 0; j < 4; ++j) qs[t][j] = 0.f;
  const float* xw = xs + warp * TPW * D;
  int it = 0;
  for (int bk = 0; bk < p.books_use; ++bk) {
    float best[TPW];
    int bidx[TPW];
#pragma unroll
    for (int t = 0; t < TPW; ++t) { best[t] = -INFINITY; bidx[t] = 0x7fffffff; }
    for (int sl = 0; sl < n_slices; ++sl, ++it) {
      // next slice (possibly of the next book) goes into the other buffer while this one is scored
      int nb = bk, ns = sl + 1;
      if (ns == n_slices) { ns = 0; ++nb; }
      if (nb < p.books_use) {
        prefetch(nb, ns, (it + 1) & 1);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
      } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
      __syncthreads();
      const float* es = es0 + (size_t)(it & 1) * CH * DP;
      float acc[TPW][CPT];
#pragma unroll
      for (int t = 0; t < TPW; ++t)
#pragma unroll
        for (int h = 0; h < CPT; ++h) acc[t][h] = 0.f;
      for (int d = 0; d < D; d += 4) {
        float4 ev[CPT], xv[TPW];
#pragma unroll
        for (int h = 0; h < CPT; ++h) ev[h] = *reinterpret_cast<const float4*>(es + (lane + 32 * h) * DP + d);
#pragma unroll
        for (int t = 0; t < TPW; ++t) xv[t] = *reinterpret_cast<const float4*>(xw + t * D + d);
#pragma unroll
        for (int t = 0; t < TPW; ++t)
#pragma unroll
          for (int h = 0; h < CPT; ++h) {
            float a = acc[t][h];
            a = fmaf(xv[t].x, ev[h].x, a); a = fmaf(xv[t].y, ev[h].y, a);
            a = fmaf(xv[t].z, ev[h].z, a); a = fmaf(xv[t].w, ev[h].w, a);
            acc[t][h] = a;
          }
      }
#pragma unroll
      for (int h = 0; h < CPT; ++h) {
        const int code = sl * CH + lane + 32 * h;
        if (code < K) {
          const float hv = __ldg(p.half_n + (size_t)bk * K + code);
#pragma unroll
          for (int t = 0; t < TPW; ++t) {
            const float sc = __fsub_rn(acc[t][h], hv);
            if (sc > best[t]) { best[t] = sc; bidx[t] = code; }
          }
        }
      }
      __syncthreads();      // the buffer scored here is refilled by the prefetch of the next iteration
    }
    // index of the first maximum, then q_sum = q_sum + (q - r) + r ; r -= q for this warp's tokens (:433-434)
    const float* book = p.books + (size_t)bk * K * D;
#pragma unroll
    for (int t = 0; t < TPW; ++t) {
      float bs = best[t];
      int bi = bidx[t];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float os = __shfl_xor_sync(0xffffffffu, bs, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (os > bs || (os == bs && oi < bi)) { bs = os; bi = oi; }
      }
      if (bi >= K) bi = 0;
      const int n = n0 + warp * TPW + t;
      if (n >= p.N) continue;
      float* xr = xs + (warp * TPW + t) * D;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int d = lane + 32 * j;
        if (d < D) {
          const float q = __ldg(book + (size_t)bi * D + d);
          const float r = xr[d];
          qs[t][j] = __fadd_rn(__fadd_rn(qs[t][j], __fsub_rn(q, r)), r);
          xr[d] = __fsub_rn(r, q);
        }
      }
      if (lane == 0) {
        if (p.idx_flat) p.idx[n] = bi;
        else {
          int b, tt;
          if (p.row_mode == ROWS_DENSE) { b = n / p.Tl; tt = n - b * p.Tl; }
          else { b = n / p.nfix; tt = p.chunk * (n - b * p.nfix + 1); }
          p.idx[((long)b * p.books_use + bk) * p.Tl + tt] = bi;
        }
      }
    }
    __syncwarp();           // the residual rows written above are read by the other lanes of this warp next book
  }
#pragma unroll
  for (int t = 0; t < TPW; ++t) {
    const int n = n0 + warp * TPW + t;
    if (n >= p.N) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int d = lane + 32 * j;
      if (d < D) p.qsum[(size_t)n * D + d] = qs[t][j];
    }
  }
}
// Residual VQ for a handful of tokens (batch-1 streaming: 75 tokens, then 4 chunk heads): ONE CTA PER TOKEN walks all
// the books.  The codes stream through shared memory in tiles of 256 (coalesced cp.async, double buffered across tiles
// and books; thread-per-row reads straight from global memory cost one L1 tag look-up per lane and load: 9 us per
// book); thread = code: one fmaf chain over d per code from a padded, conflict-free row, block arg-max (first maximum)
// at the end of a book, then the D threads that own a channel apply q_sum + (q - r) + r and r - q.  Same scores,
// tie-breaking and op order as rvq_scores_f32 + rvq_apply_f32: the indices and q_sum are the same bits.
constexpr int RVQT_TILE = 256;
__global__ void __launch_bounds__(256) rvq_token_f32(const RvqArgs p, int nbuf) {
  extern __shared__ __align__(16) float rvqt_sm[];
  __shared__ __align__(16) float r_s[256];
  __shared__ float red_s[8];
  __shared__ int red_i[8];
  const int n = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int D = p.D, K = p.K;
  const int DP = D + 4, D4 = D >> 2;                    // staged path: D % 4 == 0
  float r = 0.f, qs = 0.f;
  if (tid < D) { r = __ldg(p.x + (long)n * D + tid); r_s[tid] = r; }
  int b, tt;
  if (p.row_mode == ROWS_DENSE) { b = n / p.Tl; tt = n - b * p.Tl; }
  else { b = n / p.nfix; tt = p.chunk * (n - b * p.nfix + 1); }
  const int tiles_per_book = (K + RVQT_TILE - 1) / RVQT_TILE;
  const int n_tiles = p.books_use * tiles_per_book;
  auto stage = [&](int t) {          // tile t of the (book, tile) sequence -> buffer t % nbuf
    const int bk = t / tiles_per_book, c0 = (t - bk * tiles_per_book) * RVQT_TILE;
    const int rows = min(RVQT_TILE, K - c0);
    const float* src = p.books + ((size_t)bk * K + c0) * D;
    float* dst = rvqt_sm + (size_t)(t % nbuf) * RVQT_TILE * DP;
    for (int i = tid; i < rows * D4; i += 256) {
      const int row = i / D4, u = i - row * D4;
      cp_async16(dst + row * DP + 4 * u, src + (size_t)row * D + 4 * u);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  stage(0);
  float best = -INFINITY;
  int bidx = 0x7fffffff;
  for (int t = 0; t < n_tiles; ++t) {
    const int bk = t / tiles_per_book, ti = t - bk * tiles_per_book;
    if (nbuf == 2 && t + 1 < n_tiles) {
      stage(t + 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();                 // tile t landed; r_s holds this book's residual
    const int code = ti * RVQT_TILE + tid;
    if (code < K) {
      const float* e = rvqt_sm + (size_t)(t % nbuf) * RVQT_TILE * DP + (size_t)tid * DP;
      float acc = 0.f;
#pragma unroll 4
      for (int d = 0; d < D; d += 4) {
        const float4 ev = *reinterpret_cast<const float4*>(e + d);
        const float4 xv = *reinterpret_cast<const float4*>(r_s + d);
        acc = fmaf(xv.x, ev.x, acc); acc = fmaf(xv.y, ev.y, acc);
        acc = fmaf(xv.z, ev.z, acc); acc = fmaf(xv.w, ev.w, acc);
      }
      const float sc = __fsub_rn(acc, __ldg(p.half_n + (size_t)bk * K + code));
      if (sc > best) { best = sc; bidx = code; }
    }
    if (ti == tiles_per_book - 1) {  // end of a book: arg-max over the CTA, apply
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float os = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
        if (os > best || (os == best && oi < bidx)) { best = os; bidx = oi; }
      }
      if (lane == 0) { red_s[warp] = best; red_i[warp] = bidx; }
      __syncthreads();               // also: every thread is done reading r_s and this tile
      best = red_s[0]; bidx = red_i[0];
#pragma unroll
      for (int w = 1; w < 8; ++w) {
        const float os = red_s[w];
        const int oi = red_i[w];
        if (os > best || (os == best && oi < bidx)) { best = os; bidx = oi; }
      }
      if (bidx >= K) bidx = 0;
      if (tid < D) {
        const float q = __ldg(p.books + ((size_t)bk * K + bidx) * D + tid);
        qs = __fadd_rn(__fadd_rn(qs, __fsub_rn(q, r)), r);
        r = __fsub_rn(r, q);
        r_s[tid] = r;
      }
      if (tid == 0) {
        if (p.idx_flat) p.idx[n] = bidx;
        else p.idx[((long)b * p.books_use + bk) * p.Tl + tt] = bidx;
      }
      best = -INFINITY; bidx = 0x7fffffff;
    }
    __syncthreads();                 // tile t is free (the next stage() call of either scheme overwrites it); r_s updated
    if (nbuf == 1 && t + 1 < n_tiles) stage(t + 1);
  }
  if (tid < D) p.qsum[(long)n * D + tid] = qs;
}

// Receiver side of the residual VQ: the code indices are given, qsum[n] = sum_b book_b[idx[b][n]] (books in order, plain
// fp32 adds).  One warp per token; the index layout is the one rvq_apply_f32 / rvq_books_f32 write.
__global__ void __launch_bounds__(256) rvq_lookup_f32(const RvqArgs p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * 8 + warp;
  if (n >= p.N) return;
  int b, tt;
  if (p.row_mode == ROWS_DENSE) { b = n / p.Tl; tt = n - b * p.Tl; }
  else { b = n / p.nfix; tt = p.chunk * (n - b * p.nfix + 1); }
  for (int d = lane; d < p.D; d += 32) {
    float qs = 0.f;
    for (int bk = 0; bk < p.books_use; ++bk) {
      int bi = p.idx[((long)b * p.books_use + bk) * p.Tl + tt];
      bi = min(max(bi, 0), p.K - 1);                       // a corrupt index must not read outside the codebook
      qs = __fadd_rn(qs, __ldg(p.books + ((size_t)bk * p.K + bi) * p.D + d));
    }
    p.qsum[(size_t)n * p.D + d] = qs;
  }
}
__global__ void __launch_bounds__(256) rvq_apply_f32(const RvqArgs p, const float* __restrict__ book, float* __restrict__ res,
                                                     unsigned long long* __restrict__ keys, int bk) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * 8 + warp;
  if (n >= p.N) return;
  const unsigned long long key = keys[n];
  int bi = (int)(0xFFFFFFFFu - (unsigned int)(key & 0xFFFFFFFFull));
  if (key == 0ull || bi >= p.K) bi = 0;
  const float* e = book + (long)bi * p.D;
  for (int d = lane; d < p.D; d += 32) {
    const float q = __ldg(e + d);
    const float r = bk == 0 ? __ldg(p.x + (long)n * p.D + d) : res[(long)n * p.D + d];
    const float qs = bk == 0 ? 0.f : p.qsum[(long)n * p.D + d];
    p.qsum[(long)n * p.D + d] = __fadd_rn(__fadd_rn(qs, __fsub_rn(q, r)), r);
    res[(long)n * p.D + d] = __fsub_rn(r, q);
  }
  __syncwarp();
  if (lane == 0) {
    keys[n] = 0ull;
    if (p.idx_flat) p.idx[n] = bi;
    else {
      int b, tt;
      if (p.row_mode == ROWS_DENSE) { b = n / p.Tl; tt = n - b * p.Tl; }
      else { b = n / p.nfix; tt = p.chunk * (n - b * p.nfix + 1); }
      p.idx[((long)b * p.books_use + bk) * p.Tl + tt] = bi;
    }
  }
}

// 0.5 * |e_k|^2 for caller-provided codebooks (nearest op)
__global__ void half_sqnorm_f32(const float* __restrict__ emb, float* __restrict__ out, int K, int D) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  float s = 0.f;
  for (int d = 0; d < D; ++d) { float v = __ldg(emb + (long)k * D + d); s = fmaf(v, v, s); }
  out[k] = 0.5f * s;
}

// ---------------------------------------------------------------------------------------------
// dac ResidualVectorQuantize (eval): all stages fused.  One warp per token (8 tokens per CTA), the residual
// and the running z_q in registers (C = 32 * CPL channels, codebook dim 8).  The weights of a stage are
// staged ONCE per CTA in shared memory (cp.async, double buffered: stage s+1 streams in while stage s is
// computed) in layouts that make every warp access conflict-free:
//   Win[8][C] | bin[8] | cbnT[8][K] | c2[K] | WoutT[8][C] | bout[C]        (staged, 17C + 9K + 8 floats)
//   cb[K][8]                                                                 (global: one row gathered per token)
// cbnT = L2-normalised codebook, transposed; c2 = |cbn_k|^2; WoutT = out_proj weight, transposed.
// ---------------------------------------------------------------------------------------------
struct DacRvqArgs {
  const float* z;   // [N, C]
  const float* w;   // packed stages
  float* zq;        // [N, C]
  int* codes;       // [B, n_q, Tl]
  int N, C, K, n_q, Tl;
  long stage_stride;
};

constexpr int DACRVQ_WARPS = 8;    // warps per CTA of the one-token-per-warp form (16 was measured slower on B200)
constexpr int DACRVQ_MAX_WARPS = 10;

// T tokens per warp: every shared-memory read of a stage weight (in_proj row, normalised codebook column, out_proj
// row) then serves T tokens -- the kernel is bound by those reads (96 KB per token and stage at T = 1) and by the
// 106 KB of stage weights each CTA streams from L2 per stage, which T x more tokens per CTA amortise.  Per-token
// arithmetic and its order are identical for every T (bit-identical results); blockDim.x / 32 warps per CTA.
template <int CPL, int T>
__global__ void __launch_bounds__(32 * DACRVQ_MAX_WARPS, 1) dac_rvq_f32(const DacRvqArgs p) {
  extern __shared__ __align__(16) float wsm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;
  const int n0 = (blockIdx.x * nwarps + warp) * T;
  const int C = p.C, K = p.K;
  const int staged = 17 * C + 9 * K + 8;
  float r[T][CPL], zq[T][CPL];
  bool live[T];
  int bb[T], tt[T];
#pragma unroll
  for (int u = 0; u < T; ++u) {
    const int n = n0 + u;
    live[u] = n < p.N;
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
      r[u][i] = live[u] ? __ldg(p.z + (long)n * C + lane + 32 * i) : 0.f;
      zq[u][i] = 0.f;
    }
    bb[u] = live[u] ? n / p.Tl : 0;
    tt[u] = live[u] ? n - bb[u] * p.Tl : 0;
  }
  auto prefetch = [&](int st, int buf) {
    const float4* src = reinterpret_cast<const float4*>(p.w + (long)st * p.stage_stride);
    float4* dst = reinterpret_cast<float4*>(wsm + (long)buf * staged);
    for (int i = threadIdx.x; i < staged / 4; i += blockDim.x) cp_async16(dst + i, src + i);
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  prefetch(0, 0);
  for (int st = 0; st < p.n_q; ++st) {
    if (st + 1 < p.n_q) {
      prefetch(st + 1, (st + 1) & 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const float* Win = wsm + (long)(st & 1) * staged;
    const float* bin = Win + 8 * C;
    const float* cbnT = bin + 8;
    const float* c2 = cbnT + 8 * K;
    const float* WoutT = c2 + K;
    const float* bout = WoutT + 8 * C;
    const float* cb = p.w + (long)st * p.stage_stride + staged;
    if (live[0]) {
      // in_proj (1x1 conv C -> 8)
      float ze[T][8];
#pragma unroll
      for (int d = 0; d < 8; ++d) {
        float s[T];
#pragma unroll
        for (int u = 0; u < T; ++u) s[u] = 0.f;
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
          const float wv = Win[d * C + lane + 32 * i];
#pragma unroll
          for (int u = 0; u < T; ++u) s[u] = fmaf(wv, r[u][i], s[u]);
        }
#pragma unroll
        for (int u = 0; u < T; ++u) {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) s[u] += __shfl_xor_sync(0xffffffffu, s[u], o);
          ze[u][d] = __fadd_rn(s[u], bin[d]);
        }
      }
      // F.normalize(encodings): x / max(|x|, 1e-12)
      float en2[T][8], e2[T];
#pragma unroll
      for (int u = 0; u < T; ++u) {
        float nn = 0.f;
#pragma unroll
        for (int d = 0; d < 8; ++d) nn = fmaf(ze[u][d], ze[u][d], nn);
        const float den = fmaxf(sqrtf(nn), 1e-12f);
        float en[8];
#pragma unroll
        for (int d = 0; d < 8; ++d) en[d] = __fdiv_rn(ze[u][d], den);
        e2[u] = 0.f;
#pragma unroll
        for (int d = 0; d < 8; ++d) e2[u] = __fadd_rn(e2[u], __fmul_rn(en[d], en[d]));
#pragma unroll
        for (int d = 0; d < 8; ++d) en2[u][d] = 2.f * en[d];
      }
      // dist = |enc|^2 - 2 enc.cb + |cb|^2 ; first minimum of dist (== first maximum of -dist)
      float best[T];
      int bi[T];
#pragma unroll
      for (int u = 0; u < T; ++u) { best[u] = INFINITY; bi[u] = 0x7fffffff; }
      for (int k = lane; k < K; k += 32) {
        float cv[8];
#pragma unroll
        for (int d = 0; d < 8; ++d) cv[d] = cbnT[d * K + k];
        const float ck = c2[k];
#pragma unroll
        for (int u = 0; u < T; ++u) {
          float dot = en2[u][0] * cv[0];
#pragma unroll
          for (int d = 1; d < 8; ++d) dot = fmaf(en2[u][d], cv[d], dot);
          const float dist = __fadd_rn(__fsub_rn(e2[u], dot), ck);
          if (dist < best[u]) { best[u] = dist; bi[u] = k; }
        }
      }
      float stv[T][8];
#pragma unroll
      for (int u = 0; u < T; ++u) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float os = __shfl_xor_sync(0xffffffffu, best[u], o);
          const int oi = __shfl_xor_sync(0xffffffffu, bi[u], o);
          if (os < best[u] || (os == best[u] && oi < bi[u])) { best[u] = os; bi[u] = oi; }
        }
        if (bi[u] >= K) bi[u] = 0;
        if (lane == 0 && live[u]) p.codes[((long)bb[u] * p.n_q + st) * p.Tl + tt[u]] = bi[u];
        // straight-through value z_e + (z_q - z_e), then out_proj (1x1 conv 8 -> C)
        const float4 q0 = __ldg(reinterpret_cast<const float4*>(cb + (long)bi[u] * 8));
        const float4 q1 = __ldg(reinterpret_cast<const float4*>(cb + (long)bi[u] * 8 + 4));
        const float qv[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
        for (int d = 0; d < 8; ++d) stv[u][d] = __fadd_rn(ze[u][d], __fsub_rn(qv[d], ze[u][d]));
      }
#pragma unroll
      for (int i = 0; i < CPL; ++i) {
        const int c = lane + 32 * i;
        float wv[8];
#pragma unroll
        for (int d = 0; d < 8; ++d) wv[d] = WoutT[d * C + c];
        const float bo = bout[c];
#pragma unroll
        for (int u = 0; u < T; ++u) {
          float o = wv[0] * stv[u][0];
#pragma unroll
          for (int d = 1; d < 8; ++d) o = fmaf(wv[d], stv[u][d], o);
          o = __fadd_rn(o, bo);
          zq[u][i] = __fadd_rn(zq[u][i], o);
          r[u][i] = __fsub_rn(r[u][i], o);
        }
      }
    }
    __syncthreads();   // this buffer is refilled by the prefetch issued at the top of iteration st + 1
  }
#pragma unroll
  for (int u = 0; u < T; ++u)
    if (live[u]) {
#pragma unroll
      for (int i = 0; i < CPL; ++i) p.zq[(long)(n0 + u) * C + lane + 32 * i] = zq[u][i];
    }
}

// The same 32 stages for a handful of tokens (batch-1 streaming: 75): ONE CTA (8 warps) PER TOKEN, so the serial chain of
// a stage is split eight ways -- warp = output dimension of in_proj (the same 32-term per-lane sums and shuffle tree as
// above), thread = 4 codes of the search, thread = C/256 channels of out_proj -- instead of one warp walking ~1600
// dependent instructions per stage.  Stage weights stream through the same double-buffered shared-memory tile.  Every
// value is computed with the arithmetic and order of dac_rvq_f32: same codes, same z_q bits (0.19 -> ~0.05 ms at 75 tokens).
template <int CPL>
__global__ void __launch_bounds__(256, 1) dac_rvq_token_f32(const DacRvqArgs p) {
  extern __shared__ __align__(16) float wsm[];
  __shared__ float r_s[32 * CPL];
  __shared__ float ze_s[8];
  __shared__ float red_s[8];
  __shared__ int red_i[8];
  constexpr int CPT = CPL / 8;            // channels per thread of out_proj (C / 256)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = blockIdx.x;
  const int C = p.C, K = p.K;
  const int staged = 17 * C + 9 * K + 8;
  const int bb = n / p.Tl, tt = n - bb * p.Tl;
  float r[CPT], zq[CPT];
#pragma unroll
  for (int j = 0; j < CPT; ++j) {
    r[j] = __ldg(p.z + (long)n * C + tid + 256 * j);
    zq[j] = 0.f;
    r_s[tid + 256 * j] = r[j];
  }
  auto prefetch = [&](int st, int buf) {
    const float4* src = reinterpret_cast<const float4*>(p.w + (long)st * p.stage_stride);
    float4* dst = reinterpret_cast<float4*>(wsm + (long)buf * staged);
    for (int i = tid; i < staged / 4; i += 256) cp_async16(dst + i, src + i);
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  prefetch(0, 0);
  for (int st = 0; st < p.n_q; ++st) {
    if (st + 1 < p.n_q) {
      prefetch(st + 1, (st + 1) & 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();                       // stage weights landed; r_s of the previous stage complete
    const float* Win = wsm + (long)(st & 1) * staged;
    const float* bin = Win + 8 * C;
    const float* cbnT = bin + 8;
    const float* c2 = cbnT + 8 * K;
    const float* WoutT = c2 + K;
    const float* bout = WoutT + 8 * C;
    const float* cb = p.w + (long)st * p.stage_stride + staged;
    {  // in_proj: warp = dimension
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < CPL; ++i) s = fmaf(Win[warp * C + lane + 32 * i], r_s[lane + 32 * i], s);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) ze_s[warp] = __fadd_rn(s, bin[warp]);
    }
    __syncthreads();
    float ze[8], en2[8], e2;
    {
      float nn = 0.f;
#pragma unroll
      for (int d = 0; d < 8; ++d) { ze[d] = ze_s[d]; nn = fmaf(ze[d], ze[d], nn); }
      const float den = fmaxf(sqrtf(nn), 1e-12f);
      float en[8];
#pragma unroll
      for (int d = 0; d < 8; ++d) en[d] = __fdiv_rn(ze[d], den);
      e2 = 0.f;
#pragma unroll
      for (int d = 0; d < 8; ++d) e2 = __fadd_rn(e2, __fmul_rn(en[d], en[d]));
#pragma unroll
      for (int d = 0; d < 8; ++d) en2[d] = 2.f * en[d];
    }
    float best = INFINITY;
    int bi = 0x7fffffff;
    for (int k = tid; k < K; k += 256) {
      float dot = en2[0] * cbnT[k];
#pragma unroll
      for (int d = 1; d < 8; ++d) dot = fmaf(en2[d], cbnT[d * K + k], dot);
      const float dist = __fadd_rn(__fsub_rn(e2, dot), c2[k]);
      if (dist < best) { best = dist; bi = k; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float os = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (os < best || (os == best && oi < bi)) { best = os; bi = oi; }
    }
    if (lane == 0) { red_s[warp] = best; red_i[warp] = bi; }
    __syncthreads();
    best = red_s[0]; bi = red_i[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) {
      const float os = red_s[w];
      const int oi = red_i[w];
      if (os < best || (os == best && oi < bi)) { best = os; bi = oi; }
    }
    if (bi >= K) bi = 0;
    if (tid == 0) p.codes[((long)bb * p.n_q + st) * p.Tl + tt] = bi;
    float stv[8];
    {
      const float4 q0 = __ldg(reinterpret_cast<const float4*>(cb + (long)bi * 8));
      const float4 q1 = __ldg(reinterpret_cast<const float4*>(cb + (long)bi * 8 + 4));
      const float qv[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
      for (int d = 0; d < 8; ++d) stv[d] = __fadd_rn(ze[d], __fsub_rn(qv[d], ze[d]));
    }
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
      const int c = tid + 256 * j;
      float o = WoutT[c] * stv[0];
#pragma unroll
      for (int d = 1; d < 8; ++d) o = fmaf(WoutT[d * C + c], stv[d], o);
      o = __fadd_rn(o, bout[c]);
      zq[j] = __fadd_rn(zq[j], o);
      r[j] = __fsub_rn(r[j], o);
      r_s[c] = r[j];
    }
    __syncthreads();   // r_s complete for the next stage; this weight buffer may be refilled
  }
#pragma unroll
  for (int j = 0; j < CPT; ++j) p.zq[(long)n * C + tid + 256 * j] = zq[j];
}

// ---------------------------------------------------------------------------------------------
// small data-movement kernels
// ---------------------------------------------------------------------------------------------
__global__ void transpose_brc_f32(const float* __restrict__ in, float* __restrict__ out, int R, int C) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const float* ib = in + (size_t)b * R * C;
  float* ob = out + (size_t)b * R * C;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < R && c < C) ? ib[(size_t)r * C + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = c0 + i, r = r0 + threadIdx.x;
    if (r < R && c < C) ob[(size_t)c * R + r] = tile[threadIdx.x][i];
  }
}

__global__ void scatter_heads_f32(const float* __restrict__ src, float* __restrict__ dst, int nfix, int Tl, int chunk,
                                  int C, long total) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  long n = i / C;
  int c = (int)(i - n * C);
  int b = (int)(n / nfix), j = (int)(n - (long)b * nfix);
  dst[((long)b * Tl + (long)chunk * (j + 1)) * C + c] = src[i];
}

__global__ void widen_i32_i64(const int* __restrict__ in, long long* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i];
}

}  // namespace b2c
