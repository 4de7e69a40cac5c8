// Evaluation metrics of the codec's callers on the GPU: after the codec itself is fast the
// evaluation wall time is the per-sample Python loops of Evaluation/compare_dacvsproposal_5_eval.py:
//   * align_pair_24k (:188-211): 401 cross-correlation shifts per frame, one torch.sum each, in a Python loop,
//   * resample_f32 24 kHz -> 3 kHz (:91-97, torchaudio sinc_interp_hann polyphase FIR) + psnr_batch (:180-185),
//   * stsim_batch (:166-177): STFT(512, hop 128, hann, centre/reflect) -> |.| -> 64 HTK mel bands -> per-frame cosine.
// All are HBM/L2-bound streaming kernels: FP32 CUDA-core arithmetic, every signal read from HBM once per kernel,
// reductions in a fixed order (thread stride, then a shared-memory tree in double) so results are reproducible.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace b2c {

// block-wide sum of one double per thread (blockDim.x a multiple of 32, <= 1024); result valid in thread 0
__device__ __forceinline__ double block_sum_f64(double v, double* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sh[w] = v;
  __syncthreads();
  if (w == 0) {
    v = l < (int)(blockDim.x >> 5) ? sh[l] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  }
  __syncthreads();
  return v;
}

// ---------------------------------------------------------------------------------------------
// c[b][s + S] = sum_j ref[b][j] * est[b][j + s]  over 0 <= j < L, 0 <= j + s < L,  s = -S .. S
// (align_pair_24k :193-201: s < 0 pairs r[-s:] with e[:n], s > 0 pairs r[:-s] with e[s:]).
// A CTA owns XC_SPB consecutive shifts of one frame: every thread walks j with the block stride, reads ref[j] once
// and the XC_SPB neighbouring est values (consecutive addresses: L1 hits), fp32 partial per thread, double across
// the block.
// ---------------------------------------------------------------------------------------------
constexpr int XC_SPB = 8;
__global__ void __launch_bounds__(256) xcorr_shifts_f32(const float* __restrict__ ref, const float* __restrict__ est,
                                                        float* __restrict__ corr, int L, int S) {
  __shared__ double sh[8];
  const int b = blockIdx.y, s0 = (int)blockIdx.x * XC_SPB - S;
  const float* r = ref + (size_t)b * L;
  const float* e = est + (size_t)b * L;
  float acc[XC_SPB];
#pragma unroll
  for (int u = 0; u < XC_SPB; ++u) acc[u] = 0.f;
  for (int j = threadIdx.x; j < L; j += 256) {
    const float rv = __ldg(r + j);
#pragma unroll
    for (int u = 0; u < XC_SPB; ++u) {
      const int k = j + s0 + u;
      if (k >= 0 && k < L) acc[u] = fmaf(rv, __ldg(e + k), acc[u]);
    }
  }
#pragma unroll
  for (int u = 0; u < XC_SPB; ++u) {
    const double t = block_sum_f64((double)acc[u], sh);
    if (threadIdx.x == 0 && s0 + u <= S) corr[(size_t)b * (2 * S + 1) + (s0 + u + S)] = (float)t;
  }
}

// best_shift[b] = first s (ascending) with the strictly largest correlation (:202 `if c > best_corr`)
__global__ void xcorr_pick_first_max(const float* __restrict__ corr, int* __restrict__ best, int B, int S) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* c = corr + (size_t)b * (2 * S + 1);
  float bc = -1e18f;
  int bs = 0;
  for (int i = 0; i <= 2 * S; ++i)
    if (c[i] > bc) { bc = c[i]; bs = i - S; }
  best[b] = bs;
}

// ---------------------------------------------------------------------------------------------
// torchaudio sinc resampling (functional._apply_sinc_resample_kernel): the waveform is zero-padded by (width,
// width + orig), y[j*nw + p] = sum_k kern[p][k] * xpad[j*orig + k], k < 2*width + orig, cropped to ceil(nw*n/orig).
// `x` rows may start at a per-row offset and have a per-row length (the aligned segments of align_pair_24k).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float resample_point(const float* __restrict__ x, int n, const float* __restrict__ kern,
                                                int orig, int nw, int width, int t) {
  const int j = t / nw, p = t - j * nw;
  const int kw = 2 * width + orig;
  const float* kp = kern + (size_t)p * kw;
  const int m0 = j * orig - width;          // xpad index j*orig + k  <->  x index m0 + k
  const int k_lo = max(0, -m0), k_hi = min(kw, n - m0);
  float acc = 0.f;
  for (int k = k_lo; k < k_hi; ++k) acc = fmaf(kp[k], __ldg(x + m0 + k), acc);
  return acc;
}

__global__ void __launch_bounds__(256) resample_sinc_f32(const float* __restrict__ x, float* __restrict__ y,
                                                         const float* __restrict__ kern, int L, int Lout, int orig,
                                                         int nw, int width) {
  extern __shared__ float ksm[];
  const int kw = 2 * width + orig;
  for (int i = threadIdx.x; i < nw * kw; i += 256) ksm[i] = kern[i];
  __syncthreads();
  const int b = blockIdx.y;
  const int t = blockIdx.x * 256 + threadIdx.x;
  if (t < Lout) y[(size_t)b * Lout + t] = resample_point(x + (size_t)b * L, L, ksm, orig, nw, width, t);
}

// psnr_3k_aligned_batch (:213-223) for the whole batch in one launch: per frame the aligned segments
// (r[-s:], e[:n]) / (r[:-s], e[s:]) are resampled on the fly and only sum (r3 - e3)^2 leaves the CTA:
// psnr = 10 log10(1 / max(mse, eps)), peak 1.0 (psnr_batch :180-185).  shifts == nullptr: no alignment.
__global__ void __launch_bounds__(256) psnr_resampled_f32(const float* __restrict__ ref, const float* __restrict__ est,
                                                          const int* __restrict__ shifts, float* __restrict__ out,
                                                          const float* __restrict__ kern, int L, int orig, int nw,
                                                          int width, float eps) {
  extern __shared__ float ksm[];
  __shared__ double sh[8];
  const int kw = 2 * width + orig;
  for (int i = threadIdx.x; i < nw * kw; i += 256) ksm[i] = kern[i];
  __syncthreads();
  const int b = blockIdx.x;
  const int s = shifts ? shifts[b] : 0;
  const int n = L - abs(s);
  const float* r = ref + (size_t)b * L + (s < 0 ? -s : 0);
  const float* e = est + (size_t)b * L + (s > 0 ? s : 0);
  const int n_out = (int)(((long)nw * n + orig - 1) / orig);
  double acc = 0.0;
  for (int t = threadIdx.x; t < n_out; t += 256) {
    const float d = resample_point(r, n, ksm, orig, nw, width, t) - resample_point(e, n, ksm, orig, nw, width, t);
    acc += (double)(d * d);
  }
  acc = block_sum_f64(acc, sh);
  if (threadIdx.x == 0) {
    const float mse = n_out > 0 ? fmaxf((float)(acc / n_out), eps) : eps;
    out[b] = 10.0f * log10f(1.0f / mse);
  }
}

// psnr_batch (:180-185) on n samples per row
__global__ void __launch_bounds__(256) psnr_rows_f32(const float* __restrict__ ref, const float* __restrict__ est,
                                                     float* __restrict__ out, int n, float eps) {
  __shared__ double sh[8];
  const int b = blockIdx.x;
  const float* r = ref + (size_t)b * n;
  const float* e = est + (size_t)b * n;
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) {
    const float d = __ldg(r + i) - __ldg(e + i);
    acc += (double)(d * d);
  }
  acc = block_sum_f64(acc, sh);
  if (threadIdx.x == 0) out[b] = 10.0f * log10f(1.0f / fmaxf((float)(acc / n), eps));
}

// ---------------------------------------------------------------------------------------------
// _mel_mag (:142-163) for ref and est together: one CTA per (frame, batch row).  The two real frames are packed
// into one complex 512-point FFT (z = w*(r + i e); R[k] = (Z[k] + conj Z[N-k]) / 2, E[k] = (Z[k] - conj Z[N-k]) / 2i),
// radix-2 in shared memory, magnitudes clamped at 1e-8 (:152), 64 mel bands (dense [257, 64] filter bank as
// torchaudio's MelScale applies it), mel[b][sig][frame][band], running max per (b, sig) for the normalisation (:162).
// center=True: the signal is reflect-padded by 256 on both sides (torch.stft default pad_mode).
// ---------------------------------------------------------------------------------------------
constexpr int ST_NFFT = 512, ST_HOP = 128, ST_BINS = 257;

__device__ __forceinline__ int reflect_idx(int i, int L) {
  if (i < 0) i = -i;
  if (i >= L) i = 2 * (L - 1) - i;
  return i;
}

__global__ void __launch_bounds__(256) stft_mel_pair_f32(const float* __restrict__ ref, const float* __restrict__ est,
                                                         const float* __restrict__ fb, const int* __restrict__ mel_range,
                                                         float* __restrict__ mel, float* __restrict__ amax, int L, int frames,
                                                         int n_mels) {
  __shared__ float2 z[ST_NFFT];
  __shared__ float2 tw[ST_NFFT / 2];
  __shared__ float mag[2][ST_BINS + 3];
  const int f = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const float* r = ref + (size_t)b * L;
  const float* e = est + (size_t)b * L;
  {
    float sn, cs;
    sincospif(-(float)tid / 256.0f, &sn, &cs);      // exp(-2 pi i tid / 512)
    tw[tid] = make_float2(cs, sn);
  }
  for (int n = tid; n < ST_NFFT; n += 256) {
    const int src = reflect_idx(f * ST_HOP + n - ST_NFFT / 2, L);
    const float sw = sinpif((float)n / (float)ST_NFFT);
    const float w = sw * sw;                        // periodic hann: 0.5 - 0.5 cos(2 pi n / N)
    z[__brev((unsigned)n) >> 23] = make_float2(w * __ldg(r + src), w * __ldg(e + src));
  }
  __syncthreads();
#pragma unroll
  for (int st = 0; st < 9; ++st) {
    const int half = 1 << st;
    const int grp = tid >> st, pos = tid & (half - 1);
    const int i0 = (grp << (st + 1)) + pos, i1 = i0 + half;
    const float2 w = tw[pos << (8 - st)];
    const float2 a = z[i0], c = z[i1];
    const float2 t = make_float2(c.x * w.x - c.y * w.y, c.x * w.y + c.y * w.x);
    z[i0] = make_float2(a.x + t.x, a.y + t.y);
    z[i1] = make_float2(a.x - t.x, a.y - t.y);
    __syncthreads();
  }
  for (int k = tid; k < ST_BINS; k += 256) {
    const float2 a = z[k], c = z[(ST_NFFT - k) & (ST_NFFT - 1)];
    const float rr = 0.5f * (a.x + c.x), ri = 0.5f * (a.y - c.y);     // R = (Z[k] + conj Z[N-k]) / 2
    const float er = 0.5f * (a.y + c.y), ei = -0.5f * (a.x - c.x);    // E = (Z[k] - conj Z[N-k]) / (2i)
    mag[0][k] = fmaxf(sqrtf(rr * rr + ri * ri), 1e-8f);
    mag[1][k] = fmaxf(sqrtf(er * er + ei * ei), 1e-8f);
  }
  __syncthreads();
  // thread = (signal, band); the triangular filter of a band is non-zero on bins [lo, hi) only (mel_range, optional)
  __shared__ float wmax[2][4];
  float acc = 0.f;
  const int sig = tid >> 7, m = tid & 127;
  if (m < n_mels) {
    const int k_lo = mel_range ? mel_range[2 * m] : 0, k_hi = mel_range ? mel_range[2 * m + 1] : ST_BINS;
    for (int k = k_lo; k < k_hi; ++k) acc = fmaf(mag[sig][k], __ldg(fb + (size_t)k * n_mels + m), acc);
    mel[(((size_t)b * 2 + sig) * frames + f) * n_mels + m] = acc;
  }
  // one atomic per (CTA, signal): 128 per CTA on two addresses serialised in L2 (0.95 -> 0.1 ms per 64 frames)
  float v = acc;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((tid & 31) == 0) wmax[sig][(tid >> 5) & 3] = v;
  __syncthreads();
  if (tid < 2) {
    const float t = fmaxf(fmaxf(wmax[tid][0], wmax[tid][1]), fmaxf(wmax[tid][2], wmax[tid][3]));
    atomicMax(reinterpret_cast<int*>(amax + b * 2 + tid), __float_as_int(t));     // t >= 0: int order == float order
  }
}

// stsim_batch (:166-177): M / max(amax, 1e-8), per-frame cosine with the denominator clamped at 1e-8, clamp to
// [-1, 1], mean over frames, 0.5 * (mean + 1).  One CTA per batch row, a warp per frame.
__global__ void __launch_bounds__(256) stsim_from_mel_f32(const float* __restrict__ mel, const float* __restrict__ amax,
                                                          float* __restrict__ out, int frames, int n_mels) {
  __shared__ double sh[8];
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float ir = 1.0f / fmaxf(amax[b * 2], 1e-8f), ie = 1.0f / fmaxf(amax[b * 2 + 1], 1e-8f);
  const float* mr = mel + (size_t)b * 2 * frames * n_mels;
  const float* me = mr + (size_t)frames * n_mels;
  double acc = 0.0;
  for (int f = warp; f < frames; f += 8) {
    float num = 0.f, nr = 0.f, ne = 0.f;
    for (int m = lane; m < n_mels; m += 32) {
      const float a = mr[(size_t)f * n_mels + m] * ir, c = me[(size_t)f * n_mels + m] * ie;
      num = fmaf(a, c, num); nr = fmaf(a, a, nr); ne = fmaf(c, c, ne);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      num += __shfl_xor_sync(0xffffffffu, num, o);
      nr += __shfl_xor_sync(0xffffffffu, nr, o);
      ne += __shfl_xor_sync(0xffffffffu, ne, o);
    }
    const float den = fmaxf(sqrtf(nr) * sqrtf(ne), 1e-8f);
    const float c = fminf(fmaxf(num / den, -1.0f), 1.0f);
    if (lane == 0) acc += (double)c;
  }
  acc = block_sum_f64(acc, sh);
  if (threadIdx.x == 0) out[b] = 0.5f * ((float)(acc / frames) + 1.0f);
}

}  // namespace b2c
