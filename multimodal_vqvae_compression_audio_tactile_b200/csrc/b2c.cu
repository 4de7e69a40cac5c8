// libb2c.so -- C ABI (include/b2c.h): packed weights, straight-line programs, launchers.
#include "../../include/b2c.h"

#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "kernels_f32.cuh"
#include "kernels_tc.cuh"
#include "kernels_ru.cuh"
#include "kernels_ext.cuh"
#include "kernels_rvq.cuh"
#include "kernels_metrics.cuh"

using namespace b2c;

// ------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define CUDA_TRY(expr)                                                                       \
  do {                                                                                       \
    cudaError_t e__ = (expr);                                                                \
    if (e__ != cudaSuccess) return fail(B2C_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e__)); \
  } while (0)

// Every entry point that touches the device selects the context's device for the duration of the call and puts the
// caller's device back on return: the library never changes the process-wide current device (torch's included), and a
// program built for device d launches on d whatever device is current.
struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int device) {
    if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; ok = false; return; }
    if (prev != device && cudaSetDevice(device) != cudaSuccess) ok = false;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};
#define DEVICE_GUARD(dev)                                                                          \
  DeviceGuard guard__(dev);                                                                        \
  if (!guard__.ok) return fail(B2C_ERR_CUDA, "cannot select CUDA device %d", (int)(dev))

extern "C" const char* b2c_last_error(void) { return g_err; }
extern "C" int b2c_abi_version(void) { return B2C_ABI_VERSION; }

// timing experiments only (B2C_TC_DEBUG bit 8): copy out and reset the fused-unit pipeline trace of CTA 0
extern "C" int b2c_debug_ru_trace(unsigned long long* dst, int cap) {
  if (cap < 8192) return fail(B2C_ERR_ARG, "b2c_debug_ru_trace: cap < 8192");
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(cudaMemcpyFromSymbol(dst, b2c::g_ru_trace, 8192 * sizeof(unsigned long long)));
  void* sym = nullptr;
  CUDA_TRY(cudaGetSymbolAddress(&sym, b2c::g_ru_trace));
  CUDA_TRY(cudaMemset(sym, 0, 8192 * sizeof(unsigned long long)));
  return 8192;
}

// ------------------------------------------------------------------------------------------
// context + weights
// ------------------------------------------------------------------------------------------
enum WKind { W_CONV = 1, W_VEC = 2, W_BOOKS = 3, W_DACRVQ = 4 };

struct Weight {
  int kind = 0;
  float* dev = nullptr;       // main blob
  float* bias = nullptr;      // conv bias or null
  float* aux = nullptr;       // codebooks: half norms
  TcWeight tc;                // bf16 hi/lo operand planes for the tcgen05 path (lazily absent)
  TcBooks tcb;                // codebooks: bf16 hi/lo planes of every book + max |e|^2 per book (tcgen05 residual VQ)
  float* emax2 = nullptr;     // device copy of tcb.emax2 [n_books]
  size_t n = 0;
  int cout = 0, cin = 0, k = 0, transposed = 0, stride = 1, padding = 0;
  int n_phase = 1, kt = 0;
  int in_off[8] = {0}, out_off[8] = {0};
  int n_books = 0, K = 0, D = 0;     // codebooks / dac rvq
  int n_q = 0, c = 0;
  long stage_stride = 0;
};

struct b2c_ctx {
  int device = 0;
  std::vector<Weight> w;
  size_t bytes = 0;
  int sm_count = 148;
  // copy engine side of b2c_prog_run_host_pipelined (created on first use)
  cudaStream_t copy_stream = nullptr;   // H2D
  cudaStream_t out_stream = nullptr;    // D2H
  cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
  // second launch queue of two-lane programs (b2c_prog_set_lane): created on first use
  cudaStream_t lane_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
};

static int upload(b2c_ctx* ctx, const float* host, size_t n, float** dev) {
  CUDA_TRY(cudaMalloc((void**)dev, n * sizeof(float)));
  CUDA_TRY(cudaMemcpy(*dev, host, n * sizeof(float), cudaMemcpyHostToDevice));
  ctx->bytes += n * sizeof(float);
  return B2C_OK;
}

extern "C" int b2c_ctx_create(int device, b2c_ctx** out) {
  if (!out) return fail(B2C_ERR_ARG, "b2c_ctx_create: out is NULL");
  int count = 0;
  CUDA_TRY(cudaGetDeviceCount(&count));
  if (device < 0 || device >= count) return fail(B2C_ERR_ARG, "b2c_ctx_create: device %d of %d", device, count);
  DEVICE_GUARD(device);
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(B2C_ERR_UNSUPPORTED, "b2c: built for sm_100a only, device %d is sm_%d%d (no fallback path)", device,
                prop.major, prop.minor);
  b2c_ctx* c = new b2c_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  *out = c;
  return B2C_OK;
}

extern "C" int b2c_ctx_destroy(b2c_ctx* ctx) {
  if (!ctx) return B2C_OK;
  DeviceGuard guard__(ctx->device);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->out_stream) cudaStreamDestroy(ctx->out_stream);
  if (ctx->lane_stream) cudaStreamDestroy(ctx->lane_stream);
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
  for (int i = 0; i < 2; ++i) {
    if (ctx->ev_in[i]) cudaEventDestroy(ctx->ev_in[i]);
    if (ctx->ev_done[i]) cudaEventDestroy(ctx->ev_done[i]);
    if (ctx->ev_out[i]) cudaEventDestroy(ctx->ev_out[i]);
  }
  for (auto& w : ctx->w) {
    if (w.dev) cudaFree(w.dev);
    if (w.bias) cudaFree(w.bias);
    if (w.aux) cudaFree(w.aux);
    if (w.tcb.hi) cudaFree(w.tcb.hi);
    if (w.tcb.lo) cudaFree(w.tcb.lo);
    if (w.emax2) cudaFree(w.emax2);
    tc_weight_free(w.tc);
  }
  delete ctx;
  return B2C_OK;
}

extern "C" size_t b2c_ctx_weight_bytes(const b2c_ctx* ctx) { return ctx ? ctx->bytes : 0; }

// w = g * v / ||v||_2 over all dims but 0 (torch._weight_norm, dim = 0); norm in double.
static void fold_weight_norm(const float* v, const float* g, size_t n0, size_t inner, std::vector<float>& out) {
  out.resize(n0 * inner);
  for (size_t i = 0; i < n0; ++i) {
    const float* row = v + i * inner;
    if (!g) {
      memcpy(&out[i * inner], row, inner * sizeof(float));
      continue;
    }
    double s = 0.0;
    for (size_t j = 0; j < inner; ++j) s += (double)row[j] * (double)row[j];
    float scale = (float)((double)g[i] / sqrt(s));
    for (size_t j = 0; j < inner; ++j) out[i * inner + j] = row[j] * scale;
  }
}

extern "C" int b2c_pack_conv(b2c_ctx* ctx, const float* v, const float* g, const float* bias, int cout, int cin,
                             int k, int transposed, int stride, int padding) {
  if (!ctx || !v || cout <= 0 || cin <= 0 || k <= 0) return fail(B2C_ERR_ARG, "b2c_pack_conv: bad argument");
  DEVICE_GUARD(ctx->device);
  Weight w;
  w.kind = W_CONV;
  w.cout = cout; w.cin = cin; w.k = k; w.transposed = transposed; w.stride = stride; w.padding = padding;
  std::vector<float> folded, packed;
  if (!transposed) {
    fold_weight_norm(v, g, cout, (size_t)cin * k, folded);  // [co][ci][t]
    w.n_phase = 1; w.kt = k;
    packed.resize((size_t)k * cin * cout);
    for (int co = 0; co < cout; ++co)
      for (int ci = 0; ci < cin; ++ci)
        for (int t = 0; t < k; ++t)
          packed[((size_t)t * cin + ci) * cout + co] = folded[((size_t)co * cin + ci) * k + t];
  } else {
    if (k != 2 * stride || stride > 8 || stride < 1)
      return fail(B2C_ERR_UNSUPPORTED, "b2c_pack_conv: ConvTranspose1d needs k == 2*stride <= 16 (got k=%d s=%d)", k, stride);
    fold_weight_norm(v, g, cin, (size_t)cout * k, folded);  // [ci][co][kk]
    w.n_phase = stride; w.kt = 2;
    packed.resize((size_t)stride * 2 * cin * cout);
    for (int r = 0; r < stride; ++r) {
      int q = (r + padding) / stride, rem = (r + padding) % stride;
      w.in_off[r] = q - 1;
      w.out_off[r] = r;
      for (int t = 0; t < 2; ++t) {
        int kk = t == 0 ? rem + stride : rem;  // tap 0 reads x[j+q-1], tap 1 reads x[j+q]
        for (int ci = 0; ci < cin; ++ci)
          for (int co = 0; co < cout; ++co)
            packed[(((size_t)r * 2 + t) * cin + ci) * cout + co] = folded[((size_t)ci * cout + co) * k + kk];
      }
    }
  }
  w.n = packed.size();
  int rc = upload(ctx, packed.data(), packed.size(), &w.dev);
  if (rc) return rc;
  if (bias) {
    rc = upload(ctx, bias, cout, &w.bias);
    if (rc) return rc;
  }
  // tensor-core operand planes (bf16 hi / lo, K-major) for layers the tcgen05 kernel can take
  rc = tc_weight_pack(packed.data(), w.n_phase, w.kt, cin, cout, &w.tc, &ctx->bytes);
  if (rc) return fail(B2C_ERR_CUDA, "b2c_pack_conv: tensor-core operand upload failed (%d)", rc);
  ctx->w.push_back(w);
  return (int)ctx->w.size() - 1;
}

extern "C" int b2c_pack_vector(b2c_ctx* ctx, const float* data, size_t n) {
  if (!ctx || !data || n == 0) return fail(B2C_ERR_ARG, "b2c_pack_vector: bad argument");
  DEVICE_GUARD(ctx->device);
  Weight w;
  w.kind = W_VEC;
  w.n = n;
  int rc = upload(ctx, data, n, &w.dev);
  if (rc) return rc;
  {  // reciprocal 1 / (v + 1e-9) in fp32, read by the snake epilogue of the tensor-core kernel
    std::vector<float> inv(n);
    for (size_t i = 0; i < n; ++i) inv[i] = 1.0f / (data[i] + 1e-9f);
    rc = upload(ctx, inv.data(), n, &w.aux);
    if (rc) return rc;
  }
  ctx->w.push_back(w);
  return (int)ctx->w.size() - 1;
}

extern "C" int b2c_pack_codebooks(b2c_ctx* ctx, const float* const* books, int n_books, int K, int D) {
  if (!ctx || !books || n_books <= 0 || K <= 0 || D <= 0) return fail(B2C_ERR_ARG, "b2c_pack_codebooks: bad argument");
  DEVICE_GUARD(ctx->device);
  std::vector<float> all((size_t)n_books * K * D), hn((size_t)n_books * K);
  for (int b = 0; b < n_books; ++b) {
    if (!books[b]) return fail(B2C_ERR_ARG, "b2c_pack_codebooks: book %d is NULL", b);
    memcpy(&all[(size_t)b * K * D], books[b], (size_t)K * D * sizeof(float));
    for (int k = 0; k < K; ++k) {
      // 0.5 * (emb*emb).sum(1): fp32 accumulation like the reference's tensor op
      float s = 0.f;
      for (int d = 0; d < D; ++d) { float e = books[b][(size_t)k * D + d]; s += e * e; }
      hn[(size_t)b * K + k] = 0.5f * s;
    }
  }
  Weight w;
  w.kind = W_BOOKS; w.n_books = n_books; w.K = K; w.D = D; w.n = all.size();
  int rc = upload(ctx, all.data(), all.size(), &w.dev);
  if (rc) return rc;
  rc = upload(ctx, hn.data(), hn.size(), &w.aux);
  if (rc) return rc;
  {  // bf16 hi / lo planes (K-major) and max |e|^2 per book for the tcgen05 residual VQ
    std::vector<__nv_bfloat16> hi(all.size()), lo(all.size());
    std::vector<float> emax((size_t)n_books, 0.f);
    for (size_t i = 0; i < all.size(); ++i) {
      hi[i] = __float2bfloat16_rn(all[i]);
      lo[i] = __float2bfloat16_rn(all[i] - __bfloat162float(hi[i]));
    }
    for (int b = 0; b < n_books; ++b)
      for (int k = 0; k < K; ++k) emax[b] = fmaxf(emax[b], 2.0f * hn[(size_t)b * K + k] * 1.0001f);
    CUDA_TRY(cudaMalloc((void**)&w.tcb.hi, all.size() * 2));
    CUDA_TRY(cudaMalloc((void**)&w.tcb.lo, all.size() * 2));
    CUDA_TRY(cudaMemcpy(w.tcb.hi, hi.data(), all.size() * 2, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(w.tcb.lo, lo.data(), all.size() * 2, cudaMemcpyHostToDevice));
    rc = upload(ctx, emax.data(), emax.size(), &w.emax2);
    if (rc) return rc;
    ctx->bytes += all.size() * 4;
    w.tcb.n_books = n_books; w.tcb.K = K; w.tcb.D = D;
  }
  ctx->w.push_back(w);
  return (int)ctx->w.size() - 1;
}

// after an in-place codebook change: bf16 planes of one book and max |e|^2 (ordered-int atomicMax; values >= 0)
static __global__ void book_planes_f32(const float* __restrict__ emb, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                       const float* __restrict__ half_n, float* __restrict__ emax2, int K, int D) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < K * D) {
    __nv_bfloat16 h, l;
    split_bf16(emb[i], h, l);
    hi[i] = h; lo[i] = l;
  }
  if (i < K) atomicMax(reinterpret_cast<int*>(emax2), __float_as_int(2.0f * half_n[i] * 1.0001f));
}

// 0.5 * |e_k|^2 with the arithmetic of b2c_pack_codebooks (separate multiply and add, ascending d)
static __global__ void half_sqnorm_seq_f32(const float* __restrict__ emb, float* __restrict__ out, int K, int D) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  float s = 0.f;
  for (int d = 0; d < D; ++d) { const float e = emb[(long)k * D + d]; s = __fadd_rn(s, __fmul_rn(e, e)); }
  out[k] = 0.5f * s;
}

extern "C" int b2c_codebooks_refresh(b2c_ctx* ctx, int wid, int book, const float* dev_book, void* stream) {
  if (!ctx || wid < 0 || wid >= (int)ctx->w.size() || ctx->w[wid].kind != W_BOOKS || !dev_book)
    return fail(B2C_ERR_ARG, "b2c_codebooks_refresh: bad argument");
  Weight& w = ctx->w[wid];
  if (book < 0 || book >= w.n_books) return fail(B2C_ERR_ARG, "b2c_codebooks_refresh: book %d of %d", book, w.n_books);
  DEVICE_GUARD(ctx->device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float* dst = w.dev + (size_t)book * w.K * w.D;
  CUDA_TRY(cudaMemcpyAsync(dst, dev_book, (size_t)w.K * w.D * sizeof(float), cudaMemcpyDeviceToDevice, st));
  half_sqnorm_seq_f32<<<(w.K + 127) / 128, 128, 0, st>>>(dst, w.aux + (size_t)book * w.K, w.K, w.D);
  if (w.tcb.hi) {
    CUDA_TRY(cudaMemsetAsync(w.emax2 + book, 0, sizeof(float), st));
    const size_t off = (size_t)book * w.K * w.D;
    book_planes_f32<<<(w.K * w.D + 255) / 256, 256, 0, st>>>(dst, w.tcb.hi + off, w.tcb.lo + off, w.aux + (size_t)book * w.K,
                                                             w.emax2 + book, w.K, w.D);
  }
  CUDA_TRY(cudaPeekAtLastError());
  return B2C_OK;
}

extern "C" int b2c_pack_dac_rvq(b2c_ctx* ctx, int n_q, int c, int d, int K, const float* const* in_v,
                                const float* const* in_g, const float* const* in_b, const float* const* out_v,
                                const float* const* out_g, const float* const* out_b,
                                const float* const* codebook) {
  if (!ctx || n_q <= 0 || !in_v || !out_v || !codebook) return fail(B2C_ERR_ARG, "b2c_pack_dac_rvq: bad argument");
  if (d != 8) return fail(B2C_ERR_UNSUPPORTED, "b2c_pack_dac_rvq: codebook_dim must be 8 (got %d)", d);
  if (c % 32 != 0 || c > 1024) return fail(B2C_ERR_UNSUPPORTED, "b2c_pack_dac_rvq: latent dim %d (need multiple of 32, <= 1024)", c);
  DEVICE_GUARD(ctx->device);
  if (K % 4 != 0) return fail(B2C_ERR_UNSUPPORTED, "b2c_pack_dac_rvq: codebook size %d must be a multiple of 4", K);
  // per stage: Win[8][c] | bin[8] | cbnT[8][K] | c2[K] | WoutT[8][c] | bout[c]  (staged in smem) | cb[K][8]
  const long stride = 8L * c + 8 + 8L * K + K + 8L * c + c + 8L * K;
  std::vector<float> blob((size_t)stride * n_q, 0.f), tmp;
  for (int s = 0; s < n_q; ++s) {
    float* Win = &blob[(size_t)s * stride];
    float* bin = Win + 8L * c;
    float* cbnT = bin + 8;
    float* c2 = cbnT + 8L * K;
    float* WoutT = c2 + K;
    float* bout = WoutT + 8L * c;
    float* cb = bout + c;
    fold_weight_norm(in_v[s], in_g ? in_g[s] : nullptr, 8, c, tmp);  // [d][c]
    memcpy(Win, tmp.data(), 8L * c * sizeof(float));
    if (in_b && in_b[s]) memcpy(bin, in_b[s], 8 * sizeof(float));
    for (int k = 0; k < K; ++k) {
      const float* e = codebook[s] + 8L * k;
      float nn = 0.f;
      for (int j = 0; j < 8; ++j) nn += e[j] * e[j];
      float den = fmaxf(sqrtf(nn), 1e-12f);  // F.normalize
      float cc = 0.f;
      for (int j = 0; j < 8; ++j) {
        float v = e[j] / den;
        cbnT[(long)j * K + k] = v;
        cb[8L * k + j] = e[j];
        cc += v * v;
      }
      c2[k] = cc;
    }
    fold_weight_norm(out_v[s], out_g ? out_g[s] : nullptr, c, 8, tmp);  // [c][d]
    for (int ch = 0; ch < c; ++ch)
      for (int j = 0; j < 8; ++j) WoutT[(long)j * c + ch] = tmp[8L * ch + j];
    if (out_b && out_b[s]) memcpy(bout, out_b[s], c * sizeof(float));
  }
  Weight w;
  w.kind = W_DACRVQ; w.n_q = n_q; w.c = c; w.D = d; w.K = K; w.stage_stride = stride; w.n = blob.size();
  int rc = upload(ctx, blob.data(), blob.size(), &w.dev);
  if (rc) return rc;
  ctx->w.push_back(w);
  return (int)ctx->w.size() - 1;
}

// ------------------------------------------------------------------------------------------
// programs
// ------------------------------------------------------------------------------------------
enum OpType { OP_STEM, OP_CONV, OP_HEAD, OP_LN, OP_ATTN, OP_RVQ, OP_NEAREST, OP_DACRVQ, OP_SCATTER, OP_TRANSPOSE,
              OP_WIDEN, OP_CONV_TC, OP_CONVERT, OP_RU_TC, OP_ATTN_FULL, OP_SELECT, OP_EMA, OP_HEAD_BWD, OP_LANE, OP_JOIN };

struct Op {
  OpType type;
  b2c_ref r[6];
  int wid = -1, wid2 = -1, wid3 = -1;
  ConvArgs conv;
  LnArgs ln;
  AttnArgs attn;
  RvqArgs rvq;
  DacRvqArgs dac;
  TcConvPlan tc;
  TcRuPlan ru;
  RvqTcPlan rvq_tc;
  bool use_rvq_tc = false;
  int i[8] = {0};
  float f[2] = {0.f, 0.f};
  size_t n = 0;
  int precision = 0;
  int x_fmt = 0, act_fmt = 0;
};

struct b2c_prog {
  b2c_ctx* ctx;
  std::vector<Op> ops;
};

static const Weight* get_w(const b2c_prog* p, int wid, int kind, const char* who) {
  if (wid < 0 || wid >= (int)p->ctx->w.size() || p->ctx->w[wid].kind != kind) {
    fail(B2C_ERR_ARG, "%s: weight id %d is not of the expected kind", who, wid);
    return nullptr;
  }
  return &p->ctx->w[wid];
}

extern "C" int b2c_prog_create(b2c_ctx* ctx, b2c_prog** out) {
  if (!ctx || !out) return fail(B2C_ERR_ARG, "b2c_prog_create: NULL argument");
  *out = new b2c_prog{ctx, {}};
  return B2C_OK;
}
extern "C" int b2c_prog_destroy(b2c_prog* p) {
  delete p;
  return B2C_OK;
}
// residual VQ over all books in ONE launch (rvq_books_f32): when the 32-token blocks alone occupy the GPU
static bool rvq_one_launch(const b2c_ctx* ctx, const Op& op) {
  static int fused = -1;
  if (fused < 0) {
    const char* e = getenv("B2C_RVQ_FUSED");
    fused = (e && e[0] == '0') ? 0 : 1;
  }
  const RvqArgs& r = op.rvq;
  return fused && op.type == OP_RVQ && !r.lookup && r.books_use > 0 && op.r[1] != B2C_NULL_REF && (r.D & 3) == 0 && r.D <= 128 &&
         (long)((r.N + 31) / 32) * 2 >= ctx->sm_count;
}
// a handful of tokens (batch-1 streaming): one CTA per token walks all the books (rvq_token_f32)
static bool rvq_per_token(const b2c_ctx* ctx, const Op& op) {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("B2C_RVQ_TOKEN");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  const RvqArgs& r = op.rvq;
  return on && op.type == OP_RVQ && !r.lookup && r.books_use > 0 && op.r[1] != B2C_NULL_REF && (r.D & 3) == 0 && r.D <= 196 && r.N <= 2 * ctx->sm_count;
}
// kernel launches one run of the program enqueues (an op can be several launches)
extern "C" int b2c_prog_num_launches(const b2c_prog* p) {
  if (!p) return 0;
  int n = 0;
  for (const auto& op : p->ops) {
    if (op.type == OP_LANE || op.type == OP_JOIN) continue;   // launch-queue markers
    if (rvq_per_token(p->ctx, op)) n += 1;
    else if (op.type == OP_RVQ && op.use_rvq_tc) n += 1;
    else if (rvq_one_launch(p->ctx, op)) n += 1;
    else if (op.type == OP_RVQ && op.r[3] != B2C_NULL_REF && op.rvq.books_use > 0) n += 2 * op.rvq.books_use;   // scores + apply per book
    else if (op.type == OP_NEAREST) {
      if (op.precision == B2C_PREC_F32) n += 2;                       // half norms, scores
      else {
        // tcgen05 search: prep x2, scores, finalise (+ re-scan, apply on the rows-resident path)
        TcSearchParams q;
        size_t smem = 0;
        int grid = 0;
        n += tc_search_plan(op.rvq.N, op.rvq.D, op.rvq.K, p->ctx->sm_count, &q, &smem, &grid) ? 6 : 4;
      }
    }
    else n += 1;
  }
  return n;
}
extern "C" int b2c_prog_num_ops(const b2c_prog* p) { return p ? (int)p->ops.size() : 0; }

// Two launch queues inside one program: ops emitted after b2c_prog_set_lane(p, 1) are enqueued on a second stream
// owned by the context (forked from the caller's stream at the first of them), ops after b2c_prog_set_lane(p, 0) on the
// caller's stream again; b2c_prog_join makes the caller's stream wait for the second queue.  Used by the batch-1
// codec program: the two encoders are independent until the predictor and neither fills the GPU alone.
extern "C" int b2c_prog_set_lane(b2c_prog* p, int lane) {
  if (!p || lane < 0 || lane > 1) return fail(B2C_ERR_ARG, "b2c_prog_set_lane: lane 0 or 1");
  Op op;
  op.type = OP_LANE;
  for (auto& r : op.r) r = B2C_NULL_REF;
  op.i[0] = lane;
  p->ops.push_back(op);
  return B2C_OK;
}
extern "C" int b2c_prog_join(b2c_prog* p) {
  if (!p) return fail(B2C_ERR_ARG, "b2c_prog_join: NULL program");
  Op op;
  op.type = OP_JOIN;
  for (auto& r : op.r) r = B2C_NULL_REF;
  p->ops.push_back(op);
  return B2C_OK;
}

static void blank_refs(Op& op) {
  for (auto& r : op.r) r = B2C_NULL_REF;
}

static bool fmt_ok(int f) { return f == B2C_FMT_F32 || f == B2C_FMT_BF16X2 || f == B2C_FMT_BF16; }

extern "C" int b2c_conv_tc_eligible(const b2c_ctx* ctx, int wid, int Lin, int stride, int dilation) {
  if (!ctx || wid < 0 || wid >= (int)ctx->w.size() || ctx->w[wid].kind != W_CONV)
    return fail(B2C_ERR_ARG, "b2c_conv_tc_eligible: bad weight id %d", wid);
  const Weight& w = ctx->w[wid];
  return tc_conv_eligible(w.tc, w.cin, w.cout, w.transposed ? 1 : stride, w.transposed ? 1 : dilation, Lin) ? 1 : 0;
}

extern "C" int b2c_prog_stem(b2c_prog* p, int wid, b2c_ref x, b2c_ref out_raw, b2c_ref out_act, int act,
                             int alpha_wid, int B, int L, int act_fmt) {
  if (!p) return fail(B2C_ERR_ARG, "b2c_prog_stem: NULL program");
  const Weight* w = get_w(p, wid, W_CONV, "b2c_prog_stem");
  if (!w) return B2C_ERR_ARG;
  if (w->cin != 1 || w->k != 7 || w->transposed) return fail(B2C_ERR_ARG, "b2c_prog_stem: expects Conv1d(1, C, 7)");
  if (act == B2C_ACT_SNAKE && !get_w(p, alpha_wid, W_VEC, "b2c_prog_stem(alpha)")) return B2C_ERR_ARG;
  if (B <= 0 || L <= 0) return fail(B2C_ERR_ARG, "b2c_prog_stem: empty batch or length");
  if (!fmt_ok(act_fmt)) return fail(B2C_ERR_ARG, "b2c_prog_stem: bad activation format %d", act_fmt);
  Op op;
  op.act_fmt = act_fmt;
  op.type = OP_STEM;
  blank_refs(op);
  op.r[0] = x; op.r[1] = out_raw; op.r[2] = out_act;
  op.wid = wid; op.wid2 = alpha_wid;
  op.i[0] = B; op.i[1] = L; op.i[2] = act;
  p->ops.push_back(op);
  return B2C_OK;
}

static int add_conv(b2c_prog* p, const char* who, int wid, b2c_ref x, b2c_ref res, b2c_ref out_raw, b2c_ref out_act,
                    int act, int alpha_wid, int B, int Lin, int stride, int dilation, int padding, int res_mode,
                    int Tl, int chunk, int precision, int x_fmt, int act_fmt, b2c_ref dmul = B2C_NULL_REF) {
  if (!p) return fail(B2C_ERR_ARG, "%s: NULL program", who);
  const Weight* w = get_w(p, wid, W_CONV, who);
  if (!w) return B2C_ERR_ARG;
  if ((act == B2C_ACT_SNAKE || dmul != B2C_NULL_REF) && !get_w(p, alpha_wid, W_VEC, who)) return B2C_ERR_ARG;
  if (B <= 0 || Lin <= 0) return fail(B2C_ERR_ARG, "%s: empty batch or length", who);
  if (w->cin % 16 != 0) return fail(B2C_ERR_UNSUPPORTED, "%s: Cin=%d must be a multiple of 16", who, w->cin);
  if (w->cout % 2 != 0) return fail(B2C_ERR_UNSUPPORTED, "%s: Cout=%d must be even", who, w->cout);
  if (res_mode == 1 && (Tl <= 0 || chunk <= 0)) return fail(B2C_ERR_ARG, "%s: table residual needs Tl and chunk", who);
  if (!fmt_ok(x_fmt) || !fmt_ok(act_fmt)) return fail(B2C_ERR_ARG, "%s: bad activation format", who);
  if (precision == B2C_PREC_F32 && (x_fmt != B2C_FMT_F32 || act_fmt != B2C_FMT_F32))
    return fail(B2C_ERR_ARG, "%s: the FP32 kernel reads and writes fp32 activations (got x_fmt %d, act_fmt %d)", who, x_fmt, act_fmt);
  if (precision == B2C_PREC_BF16X3 && x_fmt != B2C_FMT_BF16X2)
    return fail(B2C_ERR_ARG, "%s: precision bf16x3 needs x as two bf16 planes (got x_fmt %d)", who, x_fmt);
  if (precision == B2C_PREC_BF16 && x_fmt == B2C_FMT_F32)
    return fail(B2C_ERR_ARG, "%s: precision bf16 needs x as bf16 plane(s)", who);
  if (precision < 0 || precision > B2C_PREC_BF16) return fail(B2C_ERR_ARG, "%s: bad precision %d", who, precision);
  Op op;
  op.type = OP_CONV;
  op.x_fmt = x_fmt; op.act_fmt = act_fmt;
  blank_refs(op);
  op.r[0] = x; op.r[1] = res; op.r[2] = out_raw; op.r[3] = out_act; op.r[4] = dmul;
  op.wid = wid; op.wid2 = alpha_wid;
  op.precision = precision;
  ConvArgs& a = op.conv;
  memset(&a, 0, sizeof(a));
  if (dmul != B2C_NULL_REF) a.dmul = reinterpret_cast<const float*>(1);   // "present" for the planner; resolved at run time
  a.B = B; a.Lin = Lin; a.Cin = w->cin; a.Cout = w->cout; a.KT = w->kt;
  a.n_phase = w->n_phase;
  a.act = act; a.res_mode = res_mode; a.Tl = Tl > 0 ? Tl : 1; a.chunk = chunk > 0 ? chunk : 1;
  if (!w->transposed) {
    if (stride < 1 || dilation < 1 || padding < 0) return fail(B2C_ERR_ARG, "%s: bad stride/dilation/padding", who);
    int Lout = (Lin + 2 * padding - dilation * (w->k - 1) - 1) / stride + 1;
    if (Lout <= 0) return fail(B2C_ERR_ARG, "%s: input of length %d is too short", who, Lin);
    a.in_step = stride; a.dil = dilation; a.Lj = Lout; a.out_step = 1; a.Lout = Lout;
    a.in_off[0] = -padding; a.out_off[0] = 0;
  } else {
    int s = w->stride;
    int Lout = (Lin - 1) * s - 2 * w->padding + w->k;
    a.in_step = 1; a.dil = 1; a.out_step = s; a.Lout = Lout; a.Lj = (Lout + s - 1) / s;
    for (int r = 0; r < s; ++r) { a.in_off[r] = w->in_off[r]; a.out_off[r] = w->out_off[r]; }
  }
  if ((long)B * a.n_phase > 65535) return fail(B2C_ERR_UNSUPPORTED, "%s: batch*phases %ld exceeds grid.z", who, (long)B * a.n_phase);
  if (precision != B2C_PREC_F32) {
    int rc = tc_conv_plan(a, w->tc, precision, act_fmt, p->ctx->sm_count, &op.tc);
    if (rc != 0)
      return fail(B2C_ERR_UNSUPPORTED, "%s: layer is not eligible for the tcgen05 kernel (plan code %d); ask "
                  "b2c_conv_tc_eligible() and use B2C_PREC_F32 for it", who, rc);
    op.type = OP_CONV_TC;
  }
  p->ops.push_back(op);
  return B2C_OK;
}

extern "C" int b2c_prog_conv(b2c_prog* p, int wid, b2c_ref x, b2c_ref res, b2c_ref out_raw, b2c_ref out_act,
                             int act, int alpha_wid, int B, int Lin, int stride, int dilation, int padding,
                             int res_mode, int Tl, int chunk, int precision, int x_fmt, int act_fmt) {
  if (p) {
    const Weight* w = get_w(p, wid, W_CONV, "b2c_prog_conv");
    if (!w) return B2C_ERR_ARG;
    if (w->transposed) return fail(B2C_ERR_ARG, "b2c_prog_conv: weight %d is a ConvTranspose1d", wid);
  }
  return add_conv(p, "b2c_prog_conv", wid, x, res, out_raw, out_act, act, alpha_wid, B, Lin, stride, dilation,
                  padding, res_mode, Tl, chunk, precision, x_fmt, act_fmt);
}

extern "C" int b2c_prog_convT(b2c_prog* p, int wid, b2c_ref x, b2c_ref out_raw, b2c_ref out_act, int act,
                              int alpha_wid, int B, int Lin, int precision, int x_fmt, int act_fmt) {
  if (p) {
    const Weight* w = get_w(p, wid, W_CONV, "b2c_prog_convT");
    if (!w) return B2C_ERR_ARG;
    if (!w->transposed) return fail(B2C_ERR_ARG, "b2c_prog_convT: weight %d is not a ConvTranspose1d", wid);
  }
  return add_conv(p, "b2c_prog_convT", wid, x, B2C_NULL_REF, out_raw, out_act, act, alpha_wid, B, Lin, 1, 1, 0, 0, 0,
                  0, precision, x_fmt, act_fmt);
}

extern "C" int b2c_prog_conv_dsnake(b2c_prog* p, int wid, b2c_ref x, b2c_ref pre, int alpha_wid, b2c_ref res, b2c_ref out_raw,
                                    b2c_ref out_act, int B, int Lin, int stride, int dilation, int padding, int precision,
                                    int x_fmt, int act_fmt) {
  if (p) {
    const Weight* w = get_w(p, wid, W_CONV, "b2c_prog_conv_dsnake");
    if (!w) return B2C_ERR_ARG;
    if (w->transposed) return fail(B2C_ERR_ARG, "b2c_prog_conv_dsnake: weight %d is a ConvTranspose1d (pack its backward form as a Conv1d)", wid);
    if (w->bias) return fail(B2C_ERR_ARG, "b2c_prog_conv_dsnake: backward-data weights carry no bias");
  }
  if (pre == B2C_NULL_REF) return fail(B2C_ERR_ARG, "b2c_prog_conv_dsnake: the pre-activation tensor is required");
  return add_conv(p, "b2c_prog_conv_dsnake", wid, x, res, out_raw, out_act, B2C_ACT_NONE, alpha_wid, B, Lin, stride, dilation,
                  padding, 0, 0, 0, precision, x_fmt, act_fmt, pre);
}

extern "C" int b2c_prog_head_bwd(b2c_prog* p, int wid, int alpha_wid, b2c_ref g_y, b2c_ref y, b2c_ref x_raw, b2c_ref g_raw,
                                 b2c_ref g_act, int B, int L, int act_fmt) {
  if (!p) return fail(B2C_ERR_ARG, "b2c_prog_head_bwd: NULL program");
  const Weight* w = get_w(p, wid, W_CONV, "b2c_prog_head_bwd");
  if (!w || !get_w(p, alpha_wid, W_VEC, "b2c_prog_head_bwd(alpha)")) return B2C_ERR_ARG;
  if (w->cout != 1 || w->k != 7 || w->transposed) return fail(B2C_ERR_ARG, "b2c_prog_head_bwd: expects Conv1d(C, 1, 7)");
  if (B <= 0 || L <= 0 || B > 65535) return fail(B2C_ERR_ARG, "b2c_prog_head_bwd: empty batch or length");
  if (!fmt_ok(act_fmt)) return fail(B2C_ERR_ARG, "b2c_prog_head_bwd: bad activation format %d", act_fmt);
  if (g_y == B2C_NULL_REF || y == B2C_NULL_REF || x_raw == B2C_NULL_REF || (g_raw == B2C_NULL_REF && g_act == B2C_NULL_REF))
    return fail(B2C_ERR_ARG, "b2c_prog_head_bwd: missing buffer");
  Op op;
  op.type = OP_HEAD_BWD;
  blank_refs(op);
  op.r[0] = g_y; op.r[1] = y; op.r[2] = x_raw; op.r[3] = g_raw; op.r[4] = g_act;
  op.wid = wid; op.wid2 = alpha_wid;
  op.i[0] = B; op.i[1] = L;
  op.act_fmt = act_fmt;
  p->ops.push_back(op);
  return B2C_OK;
}

static int ru_weights_ok(const b2c_ctx* ctx, int wid7, int wid1, int precision) {
  if (!ctx || wid7 < 0 || wid1 < 0 || wid7 >= (int)ctx->w.size() || wid1 >= (int)ctx->w.size()) return 0;
  const Weight& a = ctx->w[wid7];
  const Weight& b = ctx->w[wid1];
  if (a.kind != W_CONV || b.kind != W_CONV || a.transposed || b.transposed) return 0;
  if (a.k != 7 || b.k != 1 || a.cin != a.cout || b.cin != b.cout || a.cin != b.cin || a.stride != 1 || b.stride != 1) return 0;
  const int C = a.cin;
  if (!(C == 64 || C == 96 || C == 128 || C == 192)) return 0;
  if (!(a.tc.hi && b.tc.hi && tc_ru_enabled())) return 0;
  if (precision != B2C_PREC_BF16X3 && precision != B2C_PREC_BF16) return 0;
  TcRuPlan probe;   // shared-memory feasibility depends on the precision (two planes of h for bf16x3)
  return tc_ru_plan(1, 128, C, 1, a.tc, b.tc, precision, FMT_HI, ctx->sm_count, &probe) == 0 ? 1 : 0;
}

extern "C" int b2c_ru_tc_eligible(const b2c_ctx* ctx, int wid7, int wid1, int precision) {
  return ru_weights_ok(ctx, wid7, wid1, precision);
}

extern "C" int b2c_prog_ru(b2c_prog* p, int wid7, int alpha2_wid, int wid1, b2c_ref x_act, b2c_ref x_raw, b2c_ref out_raw,
                           b2c_ref out_act, int alpha_next_wid, int B, int L, int dilation, int precision, int act_fmt) {
  if (!p) return fail(B2C_ERR_ARG, "b2c_prog_ru: NULL program");
  if (!ru_weights_ok(p->ctx, wid7, wid1, precision))
    return fail(B2C_ERR_UNSUPPORTED, "b2c_prog_ru: weights %d / %d are not a fusable ResidualUnit (ask b2c_ru_tc_eligible)", wid7, wid1);
  if (precision != B2C_PREC_BF16X3 && precision != B2C_PREC_BF16) return fail(B2C_ERR_ARG, "b2c_prog_ru: tensor-core precisions only");
  if (!get_w(p, alpha2_wid, W_VEC, "b2c_prog_ru(alpha2)") || !get_w(p, alpha_next_wid, W_VEC, "b2c_prog_ru(alpha_next)")) return B2C_ERR_ARG;
  if (B <= 0 || L <= 0 || dilation < 1) return fail(B2C_ERR_ARG, "b2c_prog_ru: bad sizes");
  if (act_fmt != B2C_FMT_BF16X2 && act_fmt != B2C_FMT_BF16) return fail(B2C_ERR_ARG, "b2c_prog_ru: out_act is stored as bf16 plane(s)");
  const Weight& w7 = p->ctx->w[wid7];
  if (w7.padding != 3 * dilation) return fail(B2C_ERR_UNSUPPORTED, "b2c_prog_ru: 'same' padding (3*dilation) only");
  Op op;
  op.type = OP_RU_TC;
  blank_refs(op);
  op.r[0] = x_act; op.r[1] = x_raw; op.r[2] = out_raw; op.r[3] = out_act;
  op.wid = wid7; op.wid2 = alpha2_wid; op.wid3 = wid1;
  op.i[0] = alpha_next_wid; op.i[1] = B; op.i[2] = L; op.i[3] = dilation;
  op.precision = precision; op.act_fmt = act_fmt;
  int rc = tc_ru_plan(B, L, w7.cin, dilation, w7.tc, p->ctx->w[wid1].tc, precision, act_fmt, p->ctx->sm_count, &op.ru);
  if (rc) return fail(B2C_ERR_UNSUPPORTED, "b2c_prog_ru: plan failed (%d)", rc);
  p->ops.push_back(op);
  return B2C_OK;
}

extern "C" int b2c_prog_head(b2c_prog* p, int wid, b2c_ref x, b2c_ref y, int B, int L, int x_fmt) {
  if (!p) return fail(B2C_ERR_ARG, "b2c_prog_head: NULL program");
  const Weight* w = get_w(p, wid, W_CONV, "b2c_prog_head");
  if (!w) return B2C_ERR_ARG;
  if (w->cout != 1 || w->k != 7 || w->transposed) return fail(B2C_ERR_ARG, "b2c_prog_head: expects Conv1d(C, 1, 7)");
  if (B <= 0 || L <= 0) return fail(B2C_ERR_ARG, "b2c_prog_head: empty batch or length");
  if (!fmt_ok(x_fmt)) return fail(B2C_ERR_ARG, "b2c_prog_head: bad activation format %d", x_fmt);
  Op op;
  op.x_fmt = x_fmt;
  op.type = OP_HEAD;
  blank_refs(op);
  op.r[0] = x; op.r[1] = y;
  op.wid = wid;
  op.i[0] = B; op.i[1] = L;
  p->ops.push_back(op);
  return B2C_OK;
}

static int nfix_of(int Tl, int chunk) { return (Tl + chunk - 1) / chunk - 1; }

static int add_layernorm(b2c_prog* p, int gamma_wid, int beta_wid, b2c_ref a, int a_mode, b2c_ref sub, int pe_wid,
                         int pe_mode, int tanh_post, float post_scale, b2c_ref out, int N, int C, int Tl, int chunk,
                         int out_fmt, b2c_ref row_mask);

extern "C" int b2c_prog_layernorm(b2c_prog* p, int gamma_wid, int beta_wid, b2c_ref a, int a_mode, b2c_ref sub,
                                  int pe_wid, int pe_mode, int tanh_post, float post_scale, b2c_ref out, int N, int C,
                                  int Tl, int chunk, int out_fmt) {
  return add_layernorm(p, gamma_wid, beta_wid, a, a_mode, sub, pe_wid, pe_mode, tanh_post, post_scale, out, N, C, Tl,
                       chunk, out_fmt, B2C_NULL_REF);
}
extern "C" int b2c_prog_layernorm_masked(b2c_prog* p, int gamma_wid, int beta_wid, b2c_ref a, b2c_ref row_mask, int pe_wid,
                                         int pe_mode, b2c_ref out, int N, int C, int Tl, int chunk, int out_fmt) {
  if (row_mask == B2C_NULL_REF) return fail(B2C_ERR_ARG, "b2c_prog_layernorm_masked: row_mask is required");
  return add_layernorm(p, gamma_wid, beta_wid, a, B2C_ROWS_DENSE, B2C_NULL_REF, pe_wid, pe_mode, 0, 1.0f, out, N, C, Tl,
                       chunk, out_fmt, row_mask);
}

static int add_layernorm(b2c_prog* p, int gamma_wid, int beta_wid, b2c_ref a, int a_mode, b2c_ref sub, int pe_wid,
                         int pe_mode, int tanh_post, float post_scale, b2c_ref out, int N, int C, int Tl, int chunk,
                         int out_fmt, b2c_ref row_mask) {
  if (!p) return fail(B2C_ERR_ARG, "b2c_prog_layernorm: NULL program");
  if (!get_w(p, gamma_wid, W_VEC, "b2c_prog_layernorm(gamma)") || !get_w(p, beta_wid, W_VEC, "b2c_prog_layernorm(beta)"))
    return B2C_ERR_ARG;
  if (pe_mode != B2C_PE_NONE && !get_w(p, pe_wid, W_VEC, "b2c_prog_layernorm(pe)")) return B2C_ERR_ARG;
  if (N <= 0 || C <= 0 || Tl <= 0 || chunk <= 0) return fail(B2C_ERR_ARG, "b2c_prog_layernorm: bad sizes");
  if (!fmt_ok(out_fmt)) return fail(B2C_ERR_ARG, "b2c_prog_layernorm: bad activation format %d", out_fmt);
  Op op;
  op.type = OP_LN;
  blank_refs(op);
  op.r[0] = a; op.r[1] = sub; op.r[2] = out; op.r[3] = row_mask;
  op.wid = gamma_wid; op.wid2 = beta_wid; op.wid3 = pe_wid;
  LnArgs& l = op.ln;
  memset(&l, 0, sizeof(l));
  l.N = N; l.C = C; l.Tl = Tl; l.chunk = chunk; l.nfix = nfix_of(Tl, chunk) > 0 ? nfix_of(Tl, chunk) : 1;
  l.a_mode = a_mode; l.pe_mode = pe_mode; l.tanh_post = tanh_post; l.post_scale = post_scale; l.out_fmt = out_fmt;
  p->ops.push_back(op);
  return B2C_OK;
}

extern "C" int b2c_prog_attention(b2c_prog* p, b2c_ref q, int q_mode, b2c_ref kv, b2c_ref out, int B, int Tl,
                                  int chunk, int heads, int dh) {
  if (!p) return fail(B2C_ERR_ARG, "b2c_prog_attention: NULL program");
  if (dh != 128 || heads < 1 || heads > 8) return fail(B2C_ERR_UNSUPPORTED, "b2c_prog_attention: heads<=8, head dim 128 only (got %d x %d)", heads, dh);
  if (chunk < 1 || chunk > 16) return fail(B2C_ERR_UNSUPPORTED, "b2c_prog_attention: chunk %d > 16", chunk);
  if (B <= 0 || Tl <= 0) return fail(B2C_ERR_ARG, "b2c_prog_attention: empty");
  Op op;
  op.type = OP_ATTN;
  blank_refs(op);
  op.r[0] = q; op.r[1] = kv; op.r[2] = out;
  AttnArgs& a = op.attn;
  memset(&a, 0, sizeof(a));
  a.B = B; a.Tl = Tl; a.chunk = chunk; a.heads = heads; a.q_mode = q_mode;
  a.nchunks = (Tl + chunk - 1) / chunk;
  a.nfix = a.nchunks - 1;
  if (q_mode == 1 && a.nfix <= 0) return B2C_OK;  // single chunk: nothing to re-do
  p->ops.push_back(op);
  return B2C_OK;
}

extern "C" int b2c_prog_attention_full(b2c_prog* p, b2c_ref q, b2c_ref kv, b2c_ref out, int B, int T, int heads, int dh) {
  if (!p) return fail(B2C_ERR_ARG, "b2c_prog_attention_full: NULL program");
  if (dh != 128 || heads < 1) return fail(B2C_ERR_UNSUPPORTED, "b2c_prog_attention_full: head dim 128 only (got %d x %d)", heads, dh);
  if (B <= 0 || T <= 0 || B > 65535 || heads > 65535) return fail(B2C_ERR_ARG, "b2c_prog_attention_full: bad sizes");
  Op op;
  op.type = OP_ATTN_FULL;
  blank_refs(op);
  op.r[0] = q; op.r[1] = kv; op.r[2] = out;
  op.i[0] = B; op.i[1] = T; op.i[2] = heads;
  p->ops.push_back(op);
  return B2C_OK;
}

extern "C" int b2c_prog_select_rows(b2c_prog* p, b2c_ref row_mask, b2c_ref a, b2c_ref b, b2c_ref out, int N, int C) {
  if (!p) return fail(B2C_ERR_ARG, "b2c_prog_select_rows: NULL program");
  if (N <= 0 || C <= 0 || (C & 3)) return fail(B2C_ERR_ARG, "b2c_prog_select_rows: N > 0 and C a multiple of 4");
  Op op;
  op.type = OP_SELECT;
  blank_refs(op);
  op.r[0] = row_mask; op.r[1] = a; op.r[2] = b; op.r[3] = out;
  op.i[0] = N; op.i[1] = C;
  p->ops.push_back(op);
  return B2C_OK;
}

extern "C" int b2c_prog_ema_update(b2c_prog* p, b2c_ref x, b2c_ref idx, b2c_ref emb, b2c_ref counts, int N, int D, int K,
                                   float decay, float one_minus_decay) {
  if (!p) return fail(B2C_ERR_ARG, "b2c_prog_ema_update: NULL program");
  if (N <= 0 || D <= 0 || K <= 0 || D > 256) return fail(B2C_ERR_ARG, "b2c_prog_ema_update: bad sizes (D <= 256)");
  Op op;
  op.type = OP_EMA;
  blank_refs(op);
  op.r[0] = x; op.r[1] = idx; op.r[2] = emb; op.r[3] = counts;
  op.i[0] = N; op.i[1] = D; op.i[2] = K;
  op.f[0] = decay; op.f[1] = one_minus_decay;
  p->ops.push_back(op);
  return B2C_OK;
}

extern "C" size_t b2c_rvq_scratch_bytes(int N, int D) {
  if (N <= 0 || D <= 0) return 0;
  // split residual VQ (FP32 kernels, small batches): residual [N, D] fp32 + one 64-bit arg-max key per row
  return ((size_t)N * D * 4 + 255) / 256 * 256 + (size_t)N * 8;
}

extern "C" int b2c_prog_rvq(b2c_prog* p, int books_wid, int books_use, b2c_ref x, b2c_ref qsum, b2c_ref idx, b2c_ref scratch,
                            int N, int row_mode, int B, int Tl, int chunk, int precision) {
  if (!p) return fail(B2C_ERR_ARG, "b2c_prog_rvq: NULL program");
  const Weight* w = get_w(p, books_wid, W_BOOKS, "b2c_prog_rvq");
  if (!w) return B2C_ERR_ARG;
  if (books_use < 0 || books_use > w->n_books) return fail(B2C_ERR_ARG, "b2c_prog_rvq: books_use %d of %d", books_use, w->n_books);
  if (N <= 0) return fail(B2C_ERR_ARG, "b2c_prog_rvq: empty");
  if (w->D > 256) return fail(B2C_ERR_UNSUPPORTED, "b2c_prog_rvq: code dim %d > 256", w->D);
  if (precision < 0 || precision > B2C_PREC_BF16) return fail(B2C_ERR_ARG, "b2c_prog_rvq: bad precision %d", precision);
  Op op;
  op.type = OP_RVQ;
  blank_refs(op);
  op.r[0] = x; op.r[1] = qsum; op.r[2] = idx; op.r[3] = scratch;
  op.wid = books_wid;
  op.precision = precision;
  RvqArgs& r = op.rvq;
  memset(&r, 0, sizeof(r));
  r.N = N; r.D = w->D; r.K = w->K; r.books_use = books_use; r.row_mode = row_mode; r.B = B; r.Tl = Tl > 0 ? Tl : 1;
  r.chunk = chunk > 0 ? chunk : 1;
  r.nfix = nfix_of(r.Tl, r.chunk) > 0 ? nfix_of(r.Tl, r.chunk) : 1;
  r.idx_flat = 0;
  // tensor-core precisions: the one-launch tcgen05 kernel when the shape is in its range; its indices are the FP32
  // kernels' bit for bit (exact re-score of every candidate), so falling back to them changes time, not results
  if (precision != B2C_PREC_F32 && books_use > 0 && qsum != B2C_NULL_REF && w->tcb.hi)
    op.use_rvq_tc = rvq_tc_plan(N, w->D, w->K, w->n_books, books_use, &op.rvq_tc) == 0;
  p->ops.push_back(op);
  return B2C_OK;
}

extern "C" int b2c_prog_rvq_lookup(b2c_prog* p, int books_wid, int books_use, b2c_ref idx, b2c_ref qsum, int N,
                                   int row_mode, int B, int Tl, int chunk) {
  if (!p) return fail(B2C_ERR_ARG, "b2c_prog_rvq_lookup: NULL program");
  const Weight* w = get_w(p, books_wid, W_BOOKS, "b2c_prog_rvq_lookup");
  if (!w) return B2C_ERR_ARG;
  if (books_use < 0 || books_use > w->n_books)
    return fail(B2C_ERR_ARG, "b2c_prog_rvq_lookup: books_use %d of %d", books_use, w->n_books);
  if (N <= 0) return fail(B2C_ERR_ARG, "b2c_prog_rvq_lookup: empty");
  Op op;
  op.type = OP_RVQ;
  blank_refs(op);
  op.r[1] = qsum; op.r[2] = idx;
  op.wid = books_wid;
  RvqArgs& r = op.rvq;
  memset(&r, 0, sizeof(r));
  r.N = N; r.D = w->D; r.K = w->K; r.books_use = books_use; r.row_mode = row_mode; r.B = B; r.Tl = Tl > 0 ? Tl : 1;
  r.chunk = chunk > 0 ? chunk : 1;
  r.nfix = nfix_of(r.Tl, r.chunk) > 0 ? nfix_of(r.Tl, r.chunk) : 1;
  r.lookup = 1;
  p->ops.push_back(op);
  return B2C_OK;
}

extern "C" size_t b2c_nearest_scratch_bytes(int N, int D, int K, int precision) {
  if (N <= 0 || D <= 0 || K <= 0) return 0;
  if (precision == B2C_PREC_F32) return (size_t)K * sizeof(float);
  return tc_nearest_scratch_bytes(N, D, K);
}
extern "C" int b2c_nearest_tc_eligible(int N, int D, int K) {
  TcNearestDims d;
  return tc_nearest_dims(N, D, K, &d) ? 1 : 0;
}

extern "C" int b2c_prog_nearest(b2c_prog* p, b2c_ref x, b2c_ref emb, b2c_ref scratch, b2c_ref idx, int N, int D,
                                int K, int precision) {
  if (!p) return fail(B2C_ERR_ARG, "b2c_prog_nearest: NULL program");
  if (N <= 0 || D <= 0 || K <= 0) return fail(B2C_ERR_ARG, "b2c_prog_nearest: empty input (N=%d D=%d K=%d)", N, D, K);
  if (D > 256) return fail(B2C_ERR_UNSUPPORTED, "b2c_prog_nearest: D=%d > 256", D);
  if (precision != B2C_PREC_F32 && !b2c_nearest_tc_eligible(N, D, K))
    return fail(B2C_ERR_UNSUPPORTED, "b2c_prog_nearest: N=%d D=%d K=%d is not eligible for the tcgen05 search "
                "(D must be a multiple of 8); use B2C_PREC_F32", N, D, K);
  Op op;
  op.type = OP_NEAREST;
  blank_refs(op);
  op.r[0] = x; op.r[1] = emb; op.r[2] = scratch; op.r[3] = idx;
  op.precision = precision;
  RvqArgs& r = op.rvq;
  memset(&r, 0, sizeof(r));
  r.N = N; r.D = D; r.K = K; r.books_use = 1; r.idx_flat = 1; r.Tl = 1; r.chunk = 1; r.nfix = 1;
  p->ops.push_back(op);
  return B2C_OK;
}

extern "C" int b2c_prog_dac_rvq(b2c_prog* p, int wid, int n_q, b2c_ref z, b2c_ref zq, b2c_ref codes, int B, int Tl) {
  if (!p) return fail(B2C_ERR_ARG, "b2c_prog_dac_rvq: NULL program");
  const Weight* w = get_w(p, wid, W_DACRVQ, "b2c_prog_dac_rvq");
  if (!w) return B2C_ERR_ARG;
  if (n_q < 1 || n_q > w->n_q) return fail(B2C_ERR_ARG, "b2c_prog_dac_rvq: n_q %d of %d", n_q, w->n_q);
  if (B <= 0 || Tl <= 0) return fail(B2C_ERR_ARG, "b2c_prog_dac_rvq: empty");
  Op op;
  op.type = OP_DACRVQ;
  blank_refs(op);
  op.r[0] = z; op.r[1] = zq; op.r[2] = codes;
  op.wid = wid;
  DacRvqArgs& d = op.dac;
  memset(&d, 0, sizeof(d));
  d.N = B * Tl; d.C = w->c; d.K = w->K; d.n_q = n_q; d.Tl = Tl; d.stage_stride = w->stage_stride;
  p->ops.push_back(op);
  return B2C_OK;
}

extern "C" int b2c_prog_scatter_heads(b2c_prog* p, b2c_ref src, b2c_ref dst, int B, int Tl, int chunk, int C) {
  if (!p) return fail(B2C_ERR_ARG, "b2c_prog_scatter_heads: NULL program");
  int nfix = nfix_of(Tl, chunk);
  if (nfix <= 0) return B2C_OK;
  Op op;
  op.type = OP_SCATTER;
  blank_refs(op);
  op.r[0] = src; op.r[1] = dst;
  op.i[0] = B; op.i[1] = Tl; op.i[2] = chunk; op.i[3] = C; op.i[4] = nfix;
  p->ops.push_back(op);
  return B2C_OK;
}

extern "C" int b2c_prog_transpose(b2c_prog* p, b2c_ref in, b2c_ref out, int B, int R, int C) {
  if (!p) return fail(B2C_ERR_ARG, "b2c_prog_transpose: NULL program");
  if (B <= 0 || R <= 0 || C <= 0 || B > 65535) return fail(B2C_ERR_ARG, "b2c_prog_transpose: bad sizes");
  Op op;
  op.type = OP_TRANSPOSE;
  blank_refs(op);
  op.r[0] = in; op.r[1] = out;
  op.i[0] = B; op.i[1] = R; op.i[2] = C;
  p->ops.push_back(op);
  return B2C_OK;
}

extern "C" int b2c_prog_convert(b2c_prog* p, b2c_ref src, int src_fmt, b2c_ref dst, int dst_fmt, size_t n) {
  if (!p) return fail(B2C_ERR_ARG, "b2c_prog_convert: NULL program");
  if (!fmt_ok(src_fmt) || !fmt_ok(dst_fmt) || n == 0) return fail(B2C_ERR_ARG, "b2c_prog_convert: bad argument");
  if ((src_fmt == B2C_FMT_F32) == (dst_fmt == B2C_FMT_F32))
    return fail(B2C_ERR_UNSUPPORTED, "b2c_prog_convert: only fp32 <-> bf16 plane conversions (got %d -> %d)", src_fmt, dst_fmt);
  Op op;
  op.type = OP_CONVERT;
  blank_refs(op);
  op.r[0] = src; op.r[1] = dst;
  op.x_fmt = src_fmt; op.act_fmt = dst_fmt;
  op.n = n;
  p->ops.push_back(op);
  return B2C_OK;
}

extern "C" int b2c_prog_i32_to_i64(b2c_prog* p, b2c_ref in, b2c_ref out, size_t n) {
  if (!p) return fail(B2C_ERR_ARG, "b2c_prog_i32_to_i64: NULL program");
  Op op;
  op.type = OP_WIDEN;
  blank_refs(op);
  op.r[0] = in; op.r[1] = out;
  op.n = n;
  p->ops.push_back(op);
  return B2C_OK;
}

// ------------------------------------------------------------------------------------------
// run
// ------------------------------------------------------------------------------------------
struct Resolver {
  char* ws;
  size_t ws_bytes;
  void* const* ext;
  int n_ext;
  bool bad = false;
  template <class T>
  T* get(b2c_ref r) {
    if (r == B2C_NULL_REF) return nullptr;
    int slot = (int)(r >> 56);
    size_t off = (size_t)(r & 0x00FFFFFFFFFFFFFFull);
    if (slot == 0) {
      if (!ws || off >= ws_bytes) { bad = true; return nullptr; }
      return reinterpret_cast<T*>(ws + off);
    }
    if (slot > n_ext || !ext || !ext[slot - 1]) { bad = true; return nullptr; }
    return reinterpret_cast<T*>(reinterpret_cast<char*>(ext[slot - 1]) + off);
  }
};

static int launch_conv_f32(const ConvArgs& a, cudaStream_t st, int sm_count) {
  long tiles_m = (a.Lj + 127) / 128;
  long z = (long)a.B * a.n_phase;
  int tn;
  if (a.Cout % 128 == 0 && tiles_m * z * (a.Cout / 128) >= 2L * sm_count) tn = 8;
  else if (a.Cout % 64 == 0) tn = 4;
  else if (a.Cout % 96 == 0) tn = 6;
  else tn = 4;
  if (tn != 6 && a.Cout % 4 != 0) return fail(B2C_ERR_UNSUPPORTED, "conv: Cout=%d needs a multiple of 4", a.Cout);
  dim3 grid((unsigned)tiles_m, (unsigned)((a.Cout + 16 * tn - 1) / (16 * tn)), (unsigned)z);
  if (tn == 8) conv_gemm_f32<8><<<grid, 256, 0, st>>>(a);
  else if (tn == 6) conv_gemm_f32<6><<<grid, 256, 0, st>>>(a);
  else conv_gemm_f32<4><<<grid, 256, 0, st>>>(a);
  return B2C_OK;
}

static int run_ops(b2c_prog* p, cudaStream_t main_st, Resolver& R, cudaEvent_t* ev = nullptr) {
  b2c_ctx* ctx = p->ctx;
  cudaStream_t st = main_st;
  bool forked = false;
  auto join = [&]() -> int {
    if (!forked) return B2C_OK;
    CUDA_TRY(cudaEventRecord(ctx->ev_join, ctx->lane_stream));
    CUDA_TRY(cudaStreamWaitEvent(main_st, ctx->ev_join, 0));
    forked = false;
    return B2C_OK;
  };
  for (size_t oi = 0; oi < p->ops.size(); ++oi) {
    Op& op = p->ops[oi];
    if (op.type == OP_LANE || op.type == OP_JOIN) {
      if (ev) {   // per-op profile: one queue, the markers cost nothing
        cudaEventRecord(ev[2 * oi], st);
        cudaEventRecord(ev[2 * oi + 1], st);
        continue;
      }
      if (op.type == OP_JOIN) {
        int rc = join();
        if (rc) return rc;
        st = main_st;
      } else if (op.i[0] == 1) {
        if (!ctx->lane_stream) {
          CUDA_TRY(cudaStreamCreateWithFlags(&ctx->lane_stream, cudaStreamNonBlocking));
          CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
          CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
        }
        if (!forked) {
          CUDA_TRY(cudaEventRecord(ctx->ev_fork, main_st));
          CUDA_TRY(cudaStreamWaitEvent(ctx->lane_stream, ctx->ev_fork, 0));
          forked = true;
        }
        st = ctx->lane_stream;
      } else {
        st = main_st;
      }
      continue;
    }
    if (ev) cudaEventRecord(ev[2 * oi], st);
    switch (op.type) {
      case OP_STEM: {
        const Weight& w = ctx->w[op.wid];
        const float* x = R.get<const float>(op.r[0]);
        float* o_raw = R.get<float>(op.r[1]);
        void* o_act = R.get<char>(op.r[2]);
        const float* alpha = op.i[2] == ACT_SNAKE ? ctx->w[op.wid2].dev : nullptr;
        if (R.bad || !x) return fail(B2C_ERR_WORKSPACE, "op %zu (stem): unresolved buffer", oi);
        int B = op.i[0], L = op.i[1];
        dim3 grid((L + 63) / 64, B);
        size_t sm = (64 + 8 + 7 * w.cout) * sizeof(float);
        if (op.act_fmt != B2C_FMT_F32 && op.i[2] == ACT_SNAKE && w.cout % 4 == 0 && o_act) {
          __nv_bfloat16* hi = reinterpret_cast<__nv_bfloat16*>(o_act);
          stem_k7_planes<<<grid, 256, sm, st>>>(x, w.dev, w.bias, o_raw, hi,
                                                op.act_fmt == B2C_FMT_BF16X2 ? hi + (size_t)B * L * w.cout : nullptr, alpha,
                                                ctx->w[op.wid2].aux, L, w.cout);
        } else {
          stem_k7_f32<<<grid, 256, sm, st>>>(x, w.dev, w.bias, o_raw, o_act, alpha, L, w.cout, op.i[2], op.act_fmt,
                                             (size_t)B * L * w.cout);
        }
        break;
      }
      case OP_CONV:
      case OP_CONV_TC: {
        const Weight& w = ctx->w[op.wid];
        ConvArgs a = op.conv;
        const void* x_any = R.get<const char>(op.r[0]);
        void* act_any = R.get<char>(op.r[3]);
        a.x = reinterpret_cast<const float*>(x_any);
        a.res = R.get<const float>(op.r[1]);
        a.out_raw = R.get<float>(op.r[2]);
        a.out_act = reinterpret_cast<float*>(act_any);
        a.w = w.dev;
        a.bias = w.bias;
        a.dmul = R.get<const float>(op.r[4]);
        a.alpha = (a.act == ACT_SNAKE || a.dmul) ? ctx->w[op.wid2].dev : nullptr;
        if (R.bad || !a.x || (!a.out_raw && !a.out_act)) return fail(B2C_ERR_WORKSPACE, "op %zu (conv): unresolved buffer", oi);
        if (op.type == OP_CONV_TC) {
          const float* inv_alpha = (a.act == ACT_SNAKE || a.dmul) ? ctx->w[op.wid2].aux : nullptr;
          int rc = tc_conv_launch(op.tc, a, inv_alpha, x_any, act_any, w.tc, st);
          if (rc) return fail(B2C_ERR_CUDA, "op %zu (conv, tcgen05): launch failed (%d)", oi, rc);
        } else {
          int rc = launch_conv_f32(a, st, ctx->sm_count);
          if (rc) return rc;
        }
        break;
      }
      case OP_RU_TC: {
        const Weight& w7 = ctx->w[op.wid];
        const Weight& w1 = ctx->w[op.wid3];
        const Weight& a2 = ctx->w[op.wid2];
        const Weight& an = ctx->w[op.i[0]];
        TcRuArgs ra;
        ra.x_planes = R.get<const char>(op.r[0]);
        ra.x_raw = R.get<const float>(op.r[1]);
        ra.out_raw = R.get<float>(op.r[2]);
        ra.out_act = R.get<char>(op.r[3]);
        ra.bias7 = w7.bias; ra.alpha2 = a2.dev; ra.inv_alpha2 = a2.aux;
        ra.bias1 = w1.bias; ra.alpha_next = an.dev; ra.inv_alpha_next = an.aux;
        if (R.bad || !ra.x_planes || !ra.x_raw || !ra.out_act) return fail(B2C_ERR_WORKSPACE, "op %zu (residual unit): unresolved buffer", oi);
        int rc = tc_ru_launch(op.ru, ra, w7.tc, w1.tc, st);
        if (rc) return fail(B2C_ERR_CUDA, "op %zu (residual unit, tcgen05): launch failed (%d)", oi, rc);
        break;
      }
      case OP_HEAD: {
        const Weight& w = ctx->w[op.wid];
        const void* x = R.get<const char>(op.r[0]);
        float* y = R.get<float>(op.r[1]);
        if (R.bad || !x || !y) return fail(B2C_ERR_WORKSPACE, "op %zu (head): unresolved buffer", oi);
        int B = op.i[0], L = op.i[1];
        const size_t tiled_sm = ((size_t)(128 + 6) * (w.cin + 4) + 7 * w.cin) * sizeof(float);
        if (op.x_fmt != B2C_FMT_F32 && w.cin % 8 == 0 && tiled_sm <= 100 * 1024) {
          // tensor-core plans: tiled head (each input element read once); the f32 plan keeps the
          // one-warp-per-output kernel and its summation order
          dim3 grid((L + 127) / 128, B);
          const size_t xn = (size_t)B * L * w.cin;
          cudaError_t e;
          if (w.cin % 32 == 0 && w.cin <= 128) {
            // sliding-window head: weights in registers, each staged row read once per channel group
            const size_t sw_sm = (size_t)(128 + 6) * (w.cin + 4) * sizeof(float);
            const bool planes = op.x_fmt == B2C_FMT_BF16X2;
#define B2C_HEAD_SW(UPT)                                                                                                   \
  {                                                                                                                        \
    e = planes ? cudaFuncSetAttribute(head_k7_tanh_sw<FMT_PLANES, UPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sw_sm) \
               : cudaFuncSetAttribute(head_k7_tanh_sw<FMT_HI, UPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sw_sm);    \
    if (e != cudaSuccess) return fail(B2C_ERR_CUDA, "head smem: %s", cudaGetErrorString(e));                              \
    if (planes) head_k7_tanh_sw<FMT_PLANES, UPT><<<grid, 128, sw_sm, st>>>(x, w.dev, w.bias, y, L, xn);                   \
    else head_k7_tanh_sw<FMT_HI, UPT><<<grid, 128, sw_sm, st>>>(x, w.dev, w.bias, y, L, xn);                              \
  }
            switch (w.cin / 32) {
              case 1: B2C_HEAD_SW(1) break;
              case 2: B2C_HEAD_SW(2) break;
              case 3: B2C_HEAD_SW(3) break;
              default: B2C_HEAD_SW(4) break;
            }
#undef B2C_HEAD_SW
          } else if (op.x_fmt == B2C_FMT_BF16X2) {
            e = cudaFuncSetAttribute(head_k7_tanh_tiled<FMT_PLANES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tiled_sm);
            if (e != cudaSuccess) return fail(B2C_ERR_CUDA, "head smem: %s", cudaGetErrorString(e));
            head_k7_tanh_tiled<FMT_PLANES><<<grid, 128, tiled_sm, st>>>(x, w.dev, w.bias, y, L, w.cin, xn);
          } else {
            e = cudaFuncSetAttribute(head_k7_tanh_tiled<FMT_HI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tiled_sm);
            if (e != cudaSuccess) return fail(B2C_ERR_CUDA, "head smem: %s", cudaGetErrorString(e));
            head_k7_tanh_tiled<FMT_HI><<<grid, 128, tiled_sm, st>>>(x, w.dev, w.bias, y, L, w.cin, xn);
          }
        } else {
          dim3 grid((L + 63) / 64, B);
          head_k7_tanh_f32<<<grid, 256, 7 * w.cin * sizeof(float), st>>>(x, w.dev, w.bias, y, L, w.cin, op.x_fmt,
                                                                         (size_t)B * L * w.cin);
        }
        break;
      }
      case OP_LN: {
        LnArgs l = op.ln;
        l.a = R.get<const float>(op.r[0]);
        l.sub = R.get<const float>(op.r[1]);
        l.out = R.get<char>(op.r[2]);
        l.gamma = ctx->w[op.wid].dev;
        l.beta = ctx->w[op.wid2].dev;
        l.pe = l.pe_mode != PE_NONE ? ctx->w[op.wid3].dev : nullptr;
        l.rmask = R.get<const unsigned char>(op.r[3]);
        if (R.bad || !l.out || (l.a_mode != ROWS_ZERO && !l.a)) return fail(B2C_ERR_WORKSPACE, "op %zu (layernorm): unresolved buffer", oi);
        layernorm_rows_f32<<<(l.N + 7) / 8, 256, 0, st>>>(l);
        break;
      }
      case OP_ATTN: {
        AttnArgs a = op.attn;
        a.q = R.get<const float>(op.r[0]);
        a.kv = R.get<const float>(op.r[1]);
        a.out = R.get<float>(op.r[2]);
        if (R.bad || !a.q || !a.kv || !a.out) return fail(B2C_ERR_WORKSPACE, "op %zu (attention): unresolved buffer", oi);
        int blocks = a.q_mode != 1 ? a.B * a.nchunks : a.B * a.nfix;
        // a warp walks its queries serially: split the queries of a chunk over 4 CTAs (large batches) or one CTA per
        // query (small batches, where the walk is the whole latency); the K / V rows are re-read from L2
        const int qsplit = a.q_mode == 1 ? 1 : (blocks >= ctx->sm_count ? 4 : (a.chunk < 16 ? a.chunk : 16));
        attention_chunk_f32<<<dim3(blocks, qsplit), 32 * a.heads, 0, st>>>(a);
        break;
      }
      case OP_RVQ:
      case OP_NEAREST: {
        RvqArgs r = op.rvq;
        if (op.type == OP_RVQ) {
          const Weight& w = ctx->w[op.wid];
          r.x = R.get<const float>(op.r[0]);
          r.qsum = R.get<float>(op.r[1]);
          r.idx = r.books_use > 0 ? R.get<int>(op.r[2]) : nullptr;
          r.books = w.dev;
          r.half_n = w.aux;
          if (r.lookup) {
            if (R.bad || !r.qsum || (r.books_use > 0 && !r.idx))
              return fail(B2C_ERR_WORKSPACE, "op %zu (rvq lookup): unresolved buffer", oi);
            rvq_lookup_f32<<<(r.N + 7) / 8, 256, 0, st>>>(r);
            break;
          }
        } else {
          r.x = R.get<const float>(op.r[0]);
          r.books = R.get<const float>(op.r[1]);
          float* scratch = R.get<float>(op.r[2]);
          r.idx = R.get<int>(op.r[3]);
          r.qsum = nullptr;
          if (R.bad || !scratch || !r.books) return fail(B2C_ERR_WORKSPACE, "op %zu (nearest): unresolved buffer", oi);
          if (op.precision == B2C_PREC_F32) half_sqnorm_f32<<<(r.K + 127) / 128, 128, 0, st>>>(r.books, scratch, r.K, r.D);
          r.half_n = scratch;
          if (op.precision != B2C_PREC_F32) {
            if (R.bad || !r.x || !r.idx) return fail(B2C_ERR_WORKSPACE, "op %zu (nearest): unresolved buffer", oi);
            int rc = tc_nearest_launch(r.x, r.books, scratch, r.idx, r.N, r.D, r.K, ctx->sm_count, st);
            if (rc == 0) break;
            return fail(rc < 0 ? B2C_ERR_CUDA : B2C_ERR_UNSUPPORTED,
                        "op %zu (nearest, tcgen05): %s (%d)", oi, rc < 0 ? "launch failed" : "shape not eligible", rc);
          }
        }
        if (R.bad || !r.x || (!r.idx && r.books_use > 0)) return fail(B2C_ERR_WORKSPACE, "op %zu (rvq): unresolved buffer", oi);
        static int rvq_split = -1;
        if (rvq_split < 0) {
          const char* e = getenv("B2C_RVQ_SPLIT");
          rvq_split = (e && e[0] == '0') ? 0 : 1;
        }
        if (rvq_per_token(ctx, op)) {
          const size_t tile = (size_t)RVQT_TILE * (r.D + 4) * sizeof(float);
          const int nbuf = 2 * tile <= 200 * 1024 ? 2 : 1;
          cudaError_t e = cudaFuncSetAttribute(rvq_token_f32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(nbuf * tile));
          if (e != cudaSuccess) return fail(B2C_ERR_CUDA, "rvq smem: %s", cudaGetErrorString(e));
          rvq_token_f32<<<r.N, 256, nbuf * tile, st>>>(r, nbuf);
        } else if (op.type == OP_RVQ && op.use_rvq_tc) {
          const Weight& w = ctx->w[op.wid];
          RvqTcParams ra;
          memset(&ra, 0, sizeof(ra));
          ra.x = r.x; ra.books = r.books; ra.half_n = r.half_n; ra.qsum = r.qsum; ra.idx = r.idx;
          ra.row_mode = r.row_mode; ra.B = r.B; ra.Tl = r.Tl; ra.chunk = r.chunk; ra.nfix = r.nfix; ra.idx_flat = r.idx_flat;
          ra.emax2 = w.emax2;
          int rc = rvq_tc_launch(op.rvq_tc, ra, w.tcb, st);
          if (rc) return fail(B2C_ERR_CUDA, "op %zu (residual VQ, tcgen05): launch failed (%d)", oi, rc);
        } else if (r.qsum && rvq_one_launch(ctx, op)) {
          // enough 32-token blocks to occupy the GPU: one launch for all books.  (16-token CTAs, two per SM, measured
          // slower: 0.37 vs 0.28 ms at 4800 tokens -- the code tiles are then streamed twice as often.)
          const size_t sm = ((size_t)32 * r.D + (size_t)2 * 128 * (r.D + 4)) * sizeof(float);
          cudaError_t e = cudaFuncSetAttribute(rvq_books_f32<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
          if (e != cudaSuccess) return fail(B2C_ERR_CUDA, "rvq smem: %s", cudaGetErrorString(e));
          rvq_books_f32<4><<<(r.N + 31) / 32, 256, sm, st>>>(r);
        } else if (r.books_use > 0 && op.type == OP_RVQ && op.r[3] != B2C_NULL_REF && rvq_split && r.qsum) {
          // per book: scores over (token block x code slice) CTAs with an atomic arg-max, then apply
          char* scratch = R.get<char>(op.r[3]);
          if (R.bad || !scratch) return fail(B2C_ERR_WORKSPACE, "op %zu (rvq): unresolved scratch", oi);
          float* resid = reinterpret_cast<float*>(scratch);
          unsigned long long* keys = reinterpret_cast<unsigned long long*>(scratch + ((size_t)r.N * r.D * 4 + 255) / 256 * 256);
          cudaError_t me = cudaMemsetAsync(keys, 0, (size_t)r.N * 8, st);      // the scratch is arena memory: keys start empty
          if (me != cudaSuccess) return fail(B2C_ERR_CUDA, "rvq keys: %s", cudaGetErrorString(me));
          // 128-code slices (4 codes per lane) once the token blocks alone fill the GPU, else 64-code slices (more CTAs)
          const bool wide_codes = (long)((r.N + 31) / 32) * ((r.K + 127) / 128) >= 2L * ctx->sm_count;
          const int ch = wide_codes ? 128 : 64;
          const size_t sm = ((size_t)32 * r.D + (size_t)ch * (r.D + 4)) * sizeof(float);
          cudaError_t e = wide_codes ? cudaFuncSetAttribute(rvq_scores_f32<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)
                                     : cudaFuncSetAttribute(rvq_scores_f32<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
          if (e != cudaSuccess) return fail(B2C_ERR_CUDA, "rvq smem: %s", cudaGetErrorString(e));
          dim3 grid((r.N + 31) / 32, (r.K + ch - 1) / ch);
          for (int bk = 0; bk < r.books_use; ++bk) {
            const float* src = bk == 0 ? r.x : resid;
            const float* book = r.books + (size_t)bk * r.K * r.D;
            if (wide_codes) rvq_scores_f32<4><<<grid, 256, sm, st>>>(src, book, r.half_n + (size_t)bk * r.K, keys, r.N, r.D, r.K);
            else rvq_scores_f32<2><<<grid, 256, sm, st>>>(src, book, r.half_n + (size_t)bk * r.K, keys, r.N, r.D, r.K);
            rvq_apply_f32<<<(r.N + 7) / 8, 256, 0, st>>>(r, book, resid, keys, bk);
          }
        } else if (r.books_use > 0) {
          // 32 tokens per CTA when that still gives every SM two CTAs, else 8 (one per warp)
          const bool wide = (long)r.N >= 64L * ctx->sm_count;
          const int tok = wide ? 32 : 8;
          size_t sm = ((size_t)2 * tok * r.D + 64 * (r.D + 1)) * sizeof(float);
          cudaError_t e = wide ? cudaFuncSetAttribute(rvq_f32<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)
                               : cudaFuncSetAttribute(rvq_f32<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
          if (e != cudaSuccess) return fail(B2C_ERR_CUDA, "rvq smem: %s", cudaGetErrorString(e));
          if (wide) rvq_f32<4><<<(r.N + tok - 1) / tok, 256, sm, st>>>(r);
          else rvq_f32<1><<<(r.N + tok - 1) / tok, 256, sm, st>>>(r);
        } else if (r.qsum) {
          cudaError_t e = cudaMemsetAsync(r.qsum, 0, (size_t)r.N * r.D * sizeof(float), st);
          if (e != cudaSuccess) return fail(B2C_ERR_CUDA, "rvq memset: %s", cudaGetErrorString(e));
        }
        break;
      }
      case OP_DACRVQ: {
        DacRvqArgs d = op.dac;
        d.z = R.get<const float>(op.r[0]);
        d.zq = R.get<float>(op.r[1]);
        d.codes = R.get<int>(op.r[2]);
        d.w = ctx->w[op.wid].dev;
        if (R.bad || !d.z || !d.zq || !d.codes) return fail(B2C_ERR_WORKSPACE, "op %zu (dac rvq): unresolved buffer", oi);
        const size_t sm = 2 * (size_t)(17 * d.C + 9 * d.K + 8) * sizeof(float);
        if (sm > 227 * 1024) return fail(B2C_ERR_UNSUPPORTED, "dac rvq: C=%d K=%d needs %zu bytes of shared memory", d.C, d.K, sm);
        // Large batches: two tokens per warp (each weight read from shared memory serves both) and the warp count
        // that minimises (waves of CTAs) x (per-CTA time ~ fixed stage streaming + warps); small batches keep one
        // token per warp and 8 warps (more CTAs in flight, shortest latency).  Results do not depend on the choice.
        int tpw = 1, warps = DACRVQ_WARPS;
        if (d.C / 32 <= 32 && (long)d.N >= (long)DACRVQ_WARPS * ctx->sm_count) {
          tpw = 2;
          long best_cost = -1;
          for (int w = 4; w <= DACRVQ_MAX_WARPS; ++w) {
            const long ctas = (d.N + 2L * w - 1) / (2L * w);
            const long waves = (ctas + ctx->sm_count - 1) / ctx->sm_count;
            const long cost = waves * (6 + w);
            if (best_cost < 0 || cost < best_cost) { best_cost = cost; warps = w; }
          }
        }
        // a handful of tokens (batch-1 streaming): one CTA per token, the stage's serial chain split over 8 warps
        static int dac_token = -1;
        if (dac_token < 0) {
          const char* e = getenv("B2C_DACRVQ_TOKEN");
          dac_token = (e && e[0] == '0') ? 0 : 1;
        }
        if (dac_token && d.C % 256 == 0 && d.N <= 2 * ctx->sm_count) {
          cudaError_t e = cudaSuccess;
          switch (d.C / 32) {
            case 32:
              e = cudaFuncSetAttribute(dac_rvq_token_f32<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
              if (e == cudaSuccess) dac_rvq_token_f32<32><<<d.N, 256, sm, st>>>(d);
              break;
            case 16:
              e = cudaFuncSetAttribute(dac_rvq_token_f32<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
              if (e == cudaSuccess) dac_rvq_token_f32<16><<<d.N, 256, sm, st>>>(d);
              break;
            case 8:
              e = cudaFuncSetAttribute(dac_rvq_token_f32<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
              if (e == cudaSuccess) dac_rvq_token_f32<8><<<d.N, 256, sm, st>>>(d);
              break;
            default: return fail(B2C_ERR_UNSUPPORTED, "dac rvq: latent dim %d not in {256,512,1024}", d.C);
          }
          if (e != cudaSuccess) return fail(B2C_ERR_CUDA, "dac rvq smem: %s", cudaGetErrorString(e));
          break;
        }
        const int blocks = (d.N + tpw * warps - 1) / (tpw * warps);
#define B2C_DACRVQ(CPL)                                                                                             \
  {                                                                                                                 \
    cudaError_t e = tpw == 2 ? cudaFuncSetAttribute(dac_rvq_f32<CPL, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm) \
                             : cudaFuncSetAttribute(dac_rvq_f32<CPL, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); \
    if (e != cudaSuccess) return fail(B2C_ERR_CUDA, "dac rvq smem: %s", cudaGetErrorString(e));                     \
    if (tpw == 2) dac_rvq_f32<CPL, 2><<<blocks, 32 * warps, sm, st>>>(d);                                           \
    else dac_rvq_f32<CPL, 1><<<blocks, 32 * warps, sm, st>>>(d);                                                    \
  }
        switch (d.C / 32) {
          case 32: B2C_DACRVQ(32) break;
          case 16: B2C_DACRVQ(16) break;
          case 8: B2C_DACRVQ(8) break;
          case 4: B2C_DACRVQ(4) break;
          case 2: B2C_DACRVQ(2) break;
          default: return fail(B2C_ERR_UNSUPPORTED, "dac rvq: latent dim %d not in {64,128,256,512,1024}", d.C);
        }
#undef B2C_DACRVQ
        break;
      }
      case OP_SCATTER: {
        const float* src = R.get<const float>(op.r[0]);
        float* dst = R.get<float>(op.r[1]);
        if (R.bad || !src || !dst) return fail(B2C_ERR_WORKSPACE, "op %zu (scatter): unresolved buffer", oi);
        long total = (long)op.i[0] * op.i[4] * op.i[3];
        scatter_heads_f32<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(src, dst, op.i[4], op.i[1], op.i[2], op.i[3], total);
        break;
      }
      case OP_TRANSPOSE: {
        const float* in = R.get<const float>(op.r[0]);
        float* out = R.get<float>(op.r[1]);
        if (R.bad || !in || !out) return fail(B2C_ERR_WORKSPACE, "op %zu (transpose): unresolved buffer", oi);
        dim3 grid((op.i[2] + 31) / 32, (op.i[1] + 31) / 32, op.i[0]);
        transpose_brc_f32<<<grid, dim3(32, 8), 0, st>>>(in, out, op.i[1], op.i[2]);
        break;
      }
      case OP_CONVERT: {
        const void* src = R.get<const char>(op.r[0]);
        void* dst = R.get<char>(op.r[1]);
        if (R.bad || !src || !dst) return fail(B2C_ERR_WORKSPACE, "op %zu (convert): unresolved buffer", oi);
        if (op.x_fmt == B2C_FMT_F32) {
          __nv_bfloat16* hi = reinterpret_cast<__nv_bfloat16*>(dst);
          split_planes_f32<<<(unsigned)((op.n / 4 + 256) / 256), 256, 0, st>>>(
              reinterpret_cast<const float*>(src), hi, op.act_fmt == B2C_FMT_BF16X2 ? hi + op.n : nullptr, op.n);
        } else {
          const __nv_bfloat16* hi = reinterpret_cast<const __nv_bfloat16*>(src);
          merge_planes_f32<<<(unsigned)((op.n + 255) / 256), 256, 0, st>>>(
              hi, op.x_fmt == B2C_FMT_BF16X2 ? hi + op.n : nullptr, reinterpret_cast<float*>(dst), op.n);
        }
        break;
      }
      case OP_ATTN_FULL: {
        AttnFullArgs a;
        a.q = R.get<const float>(op.r[0]);
        a.kv = R.get<const float>(op.r[1]);
        a.out = R.get<float>(op.r[2]);
        a.B = op.i[0]; a.T = op.i[1]; a.heads = op.i[2];
        if (R.bad || !a.q || !a.kv || !a.out) return fail(B2C_ERR_WORKSPACE, "op %zu (full attention): unresolved buffer", oi);
        cudaError_t e = cudaFuncSetAttribute(attention_full_f32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AF_SMEM);
        if (e != cudaSuccess) return fail(B2C_ERR_CUDA, "full attention smem: %s", cudaGetErrorString(e));
        attention_full_f32<<<dim3((a.T + AF_BQ - 1) / AF_BQ, a.heads, a.B), 256, AF_SMEM, st>>>(a);
        break;
      }
      case OP_SELECT: {
        const unsigned char* mk = R.get<const unsigned char>(op.r[0]);
        const float* a = R.get<const float>(op.r[1]);
        const float* b = R.get<const float>(op.r[2]);
        float* out = R.get<float>(op.r[3]);
        if (R.bad || !mk || !a || !b || !out) return fail(B2C_ERR_WORKSPACE, "op %zu (select rows): unresolved buffer", oi);
        const long total4 = (long)op.i[0] * (op.i[1] / 4);
        select_rows_f32<<<(unsigned)((total4 + 255) / 256), 256, 0, st>>>(mk, a, b, out, total4, op.i[1] / 4);
        break;
      }
      case OP_EMA: {
        const float* x = R.get<const float>(op.r[0]);
        const int* idx = R.get<const int>(op.r[1]);
        float* emb = R.get<float>(op.r[2]);
        int* counts = R.get<int>(op.r[3]);
        if (R.bad || !x || !idx || !emb) return fail(B2C_ERR_WORKSPACE, "op %zu (ema update): unresolved buffer", oi);
        ema_update_f32<<<op.i[2], 128, 0, st>>>(x, idx, emb, counts, op.i[0], op.i[1], op.f[0], op.f[1]);
        break;
      }
      case OP_HEAD_BWD: {
        const Weight& w = ctx->w[op.wid];
        const float* gy = R.get<const float>(op.r[0]);
        const float* yy = R.get<const float>(op.r[1]);
        const float* xr = R.get<const float>(op.r[2]);
        float* graw = R.get<float>(op.r[3]);
        void* gact = R.get<char>(op.r[4]);
        if (R.bad || !gy || !yy || !xr) return fail(B2C_ERR_WORKSPACE, "op %zu (head backward): unresolved buffer", oi);
        const int B = op.i[0], L = op.i[1];
        const size_t sm = (size_t)(64 + 8 + 7 * w.cin) * sizeof(float);
        head_k7_bwd_f32<<<dim3((L + 63) / 64, B), 256, sm, st>>>(gy, yy, xr, w.dev, ctx->w[op.wid2].dev, graw, gact, op.act_fmt,
                                                                  (size_t)B * L * w.cin, L, w.cin);
        break;
      }
      case OP_WIDEN: {
        const int* in = R.get<const int>(op.r[0]);
        long long* out = R.get<long long>(op.r[1]);
        if (R.bad || !in || !out) return fail(B2C_ERR_WORKSPACE, "op %zu (widen): unresolved buffer", oi);
        widen_i32_i64<<<(unsigned)((op.n + 255) / 256), 256, 0, st>>>(in, out, op.n);
        break;
      }
      case OP_LANE:
      case OP_JOIN: break;   // handled above
    }
    if (ev) cudaEventRecord(ev[2 * oi + 1], st);
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
      join();
      return fail(B2C_ERR_CUDA, "op %zu (type %d): %s", oi, (int)op.type, cudaGetErrorString(e));
    }
  }
  return join();   // a program never returns with work on the second queue the caller's stream does not wait for
}

// algorithmic work of one launch: flops of the contraction, minimal global bytes
static void op_work(const b2c_ctx* ctx, const Op& op, int* kind, double* flops, double* bytes) {
  *flops = 0; *bytes = 0; *kind = B2C_KIND_MOVE;
  switch (op.type) {
    case OP_CONV:
    case OP_CONV_TC: {
      const ConvArgs& a = op.conv;
      double rows = (double)a.B * a.Lout;
      *kind = op.type == OP_CONV ? B2C_KIND_CONV_F32 : (op.precision == B2C_PREC_BF16X3 ? B2C_KIND_CONV_TC_X3 : B2C_KIND_CONV_TC);
      *flops = 2.0 * rows * a.Cout * a.Cin * a.KT;
      double outs = (op.r[2] != B2C_NULL_REF ? 1 : 0) + (op.r[3] != B2C_NULL_REF ? 1 : 0) + (op.r[1] != B2C_NULL_REF ? 1 : 0);
      *bytes = 4.0 * ((double)a.B * a.Lin * a.Cin + rows * a.Cout * outs + (double)a.n_phase * a.KT * a.Cin * a.Cout);
      break;
    }
    case OP_RU_TC: {
      const Weight& w = ctx->w[op.wid];
      double rows = (double)op.i[1] * op.i[2], C = w.cin;
      *kind = op.precision == B2C_PREC_BF16X3 ? B2C_KIND_CONV_TC_X3 : B2C_KIND_CONV_TC;
      *flops = 2.0 * rows * C * C * 8.0;                       // k = 7 conv + k = 1 conv
      *bytes = 4.0 * (rows * C * (3.0 + (op.r[2] != B2C_NULL_REF ? 1 : 0)) + 8.0 * C * C);   // x_act, x_raw, out_act (+ out_raw)
      break;
    }
    case OP_STEM: {
      const Weight& w = ctx->w[op.wid];
      double n = (double)op.i[0] * op.i[1];
      *kind = B2C_KIND_STEM; *flops = 2.0 * n * 7 * w.cout;
      *bytes = 4.0 * (n + n * w.cout * ((op.r[1] != B2C_NULL_REF) + (op.r[2] != B2C_NULL_REF)));
      break;
    }
    case OP_HEAD_BWD: {
      const Weight& w = ctx->w[op.wid];
      double n = (double)op.i[0] * op.i[1];
      *kind = B2C_KIND_HEAD; *flops = 2.0 * n * 7 * w.cin; *bytes = 4.0 * (2.0 * n * w.cin + 2.0 * n) + 2.0 * n * w.cin;
      break;
    }
    case OP_HEAD: {
      const Weight& w = ctx->w[op.wid];
      double n = (double)op.i[0] * op.i[1];
      *kind = B2C_KIND_HEAD; *flops = 2.0 * n * 7 * w.cin; *bytes = 4.0 * (n * w.cin + n);
      break;
    }
    case OP_LN: *kind = B2C_KIND_LAYERNORM; *flops = 8.0 * op.ln.N * op.ln.C; *bytes = 4.0 * 2 * (double)op.ln.N * op.ln.C; break;
    case OP_ATTN: {
      const AttnArgs& a = op.attn;
      double nq = a.q_mode == 1 ? (double)a.B * a.nfix : (double)a.B * a.Tl;
      *kind = B2C_KIND_ATTENTION; *flops = 4.0 * nq * a.chunk * a.heads * 128;
      *bytes = 4.0 * ((double)a.B * a.Tl * 2 * a.heads * 128 + 2 * nq * a.heads * 128);
      break;
    }
    case OP_RVQ:
    case OP_NEAREST: {
      const RvqArgs& r = op.rvq;
      *kind = op.type == OP_RVQ ? B2C_KIND_RVQ : B2C_KIND_NEAREST;
      *flops = 2.0 * r.N * r.D * r.K * r.books_use;
      // 4*(3*N*D + K*D) + 2*N per book for the residual VQ; the bare search reads x and emb, writes idx
      *bytes = op.type == OP_RVQ ? r.books_use * (4.0 * (3.0 * r.N * r.D + (double)r.K * r.D) + 2.0 * r.N)
                                 : 4.0 * ((double)r.N * r.D + (double)r.K * r.D + r.N);
      break;
    }
    case OP_DACRVQ: {
      const DacRvqArgs& d = op.dac;
      *kind = B2C_KIND_DAC_RVQ; *flops = 2.0 * d.N * d.n_q * (16.0 * d.C + 8.0 * d.K);
      *bytes = 4.0 * (2.0 * d.N * d.C + (double)d.n_q * d.stage_stride + (double)d.N * d.n_q);
      break;
    }
    case OP_ATTN_FULL: {
      const double B = op.i[0], T = op.i[1], H = op.i[2];
      *kind = B2C_KIND_ATTENTION; *flops = 4.0 * B * H * T * T * 128; *bytes = 4.0 * 4.0 * B * T * H * 128;
      break;
    }
    case OP_SELECT: *bytes = 8.0 * op.i[0] * op.i[1]; break;
    case OP_EMA: *bytes = 4.0 * ((double)op.i[0] * op.i[1] + 2.0 * op.i[2] * op.i[1] + op.i[0]); break;
    case OP_SCATTER: *bytes = 8.0 * op.i[0] * op.i[4] * op.i[3]; break;
    case OP_TRANSPOSE: *bytes = 8.0 * op.i[0] * op.i[1] * op.i[2]; break;
    case OP_WIDEN: *bytes = 12.0 * op.n; break;
    case OP_CONVERT: *bytes = 8.0 * op.n; break;
  }
}

extern "C" int b2c_prog_profile(b2c_prog* p, void* stream, void* workspace, size_t workspace_bytes,
                                void* const* ext, int n_ext, float* ms, int* kind, double* flops, double* bytes,
                                int cap) {
  if (!p || !ms || !kind || !flops || !bytes) return fail(B2C_ERR_ARG, "b2c_prog_profile: NULL argument");
  DEVICE_GUARD(p->ctx->device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  size_t n = p->ops.size();
  std::vector<cudaEvent_t> ev(2 * n);
  for (auto& e : ev) CUDA_TRY(cudaEventCreate(&e));
  Resolver R{reinterpret_cast<char*>(workspace), workspace_bytes, ext, n_ext};
  int rc = run_ops(p, st, R, ev.data());
  cudaError_t se = cudaStreamSynchronize(st);
  if (rc == B2C_OK && se != cudaSuccess) rc = fail(B2C_ERR_CUDA, "b2c_prog_profile: %s", cudaGetErrorString(se));
  for (size_t i = 0; i < n && (int)i < cap && rc == B2C_OK; ++i) {
    cudaEventElapsedTime(&ms[i], ev[2 * i], ev[2 * i + 1]);
    op_work(p->ctx, p->ops[i], &kind[i], &flops[i], &bytes[i]);
  }
  for (auto& e : ev) cudaEventDestroy(e);
  return rc == B2C_OK ? (int)n : rc;
}

extern "C" int b2c_prog_run(b2c_prog* p, void* stream, void* workspace, size_t workspace_bytes, void* const* ext,
                            int n_ext) {
  if (!p) return fail(B2C_ERR_ARG, "b2c_prog_run: NULL program");
  DEVICE_GUARD(p->ctx->device);
  Resolver R{reinterpret_cast<char*>(workspace), workspace_bytes, ext, n_ext};
  return run_ops(p, reinterpret_cast<cudaStream_t>(stream), R);
}

extern "C" int b2c_prog_run_host(b2c_prog* p, void* stream, void* workspace, size_t workspace_bytes,
                                 void* const* ext, int n_ext, const b2c_hostcopy* h2d, int n_h2d,
                                 const b2c_hostcopy* d2h, int n_d2h) {
  if (!p) return fail(B2C_ERR_ARG, "b2c_prog_run_host: NULL program");
  DEVICE_GUARD(p->ctx->device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  for (int i = 0; i < n_h2d; ++i) {
    if (h2d[i].slot < 1 || h2d[i].slot > n_ext || !h2d[i].host) return fail(B2C_ERR_ARG, "b2c_prog_run_host: bad h2d[%d]", i);
    CUDA_TRY(cudaMemcpyAsync(ext[h2d[i].slot - 1], h2d[i].host, h2d[i].bytes, cudaMemcpyHostToDevice, st));
  }
  int rc = b2c_prog_run(p, stream, workspace, workspace_bytes, ext, n_ext);
  if (rc) return rc;
  for (int i = 0; i < n_d2h; ++i) {
    if (d2h[i].slot < 1 || d2h[i].slot > n_ext || !d2h[i].host) return fail(B2C_ERR_ARG, "b2c_prog_run_host: bad d2h[%d]", i);
    CUDA_TRY(cudaMemcpyAsync(d2h[i].host, ext[d2h[i].slot - 1], d2h[i].bytes, cudaMemcpyDeviceToHost, st));
  }
  CUDA_TRY(cudaStreamSynchronize(st));
  return B2C_OK;
}

extern "C" int b2c_prog_run_host_pipelined(b2c_prog* p, void* stream, void* workspace, size_t workspace_bytes,
                                           void* const* ext_sets, int n_ext, const b2c_hostcopy* h2d, int n_h2d,
                                           const b2c_hostcopy* d2h, int n_d2h, int n_micro) {
  if (!p || !ext_sets || n_micro <= 0) return fail(B2C_ERR_ARG, "b2c_prog_run_host_pipelined: bad argument");
  DEVICE_GUARD(p->ctx->device);
  b2c_ctx* ctx = p->ctx;
  cudaStream_t cs = reinterpret_cast<cudaStream_t>(stream);
  if (!ctx->copy_stream) {
    CUDA_TRY(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&ctx->out_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_in[i], cudaEventDisableTiming));
      CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_done[i], cudaEventDisableTiming));
      CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_out[i], cudaEventDisableTiming));
    }
  }
  // Three in-order queues: xs carries only H2D copies, cs the programs, os only D2H copies.  H2D(k+1) therefore waits
  // for nothing but "the program of k-1 has finished with staging set (k+1)&1" and runs under the program of k; D2H(k)
  // runs under the program of k+1.  (One copy queue would order H2D(k+1) behind D2H(k), i.e. behind the program of k.)
  cudaStream_t xs = ctx->copy_stream, os = ctx->out_stream;
  for (int i = 0; i < n_h2d; ++i)
    if (h2d[i].slot < 1 || h2d[i].slot > n_ext || !h2d[i].host) return fail(B2C_ERR_ARG, "b2c_prog_run_host_pipelined: bad h2d[%d]", i);
  for (int i = 0; i < n_d2h; ++i)
    if (d2h[i].slot < 1 || d2h[i].slot > n_ext || !d2h[i].host) return fail(B2C_ERR_ARG, "b2c_prog_run_host_pipelined: bad d2h[%d]", i);
  for (int k = 0; k < n_micro; ++k) {
    const int set = k & 1;
    void* const* ext = ext_sets + (size_t)set * n_ext;
    // H2D queue: inputs of micro-batch k (the program of k-2 was the last reader of this staging set)
    if (k >= 2) CUDA_TRY(cudaStreamWaitEvent(xs, ctx->ev_done[set], 0));
    for (int i = 0; i < n_h2d; ++i)
      CUDA_TRY(cudaMemcpyAsync(ext[h2d[i].slot - 1], (const char*)h2d[i].host + (size_t)k * h2d[i].bytes, h2d[i].bytes,
                               cudaMemcpyHostToDevice, xs));
    CUDA_TRY(cudaEventRecord(ctx->ev_in[set], xs));
    // compute queue: wait for the inputs, and for the D2H of k-2 that last read this set's outputs
    CUDA_TRY(cudaStreamWaitEvent(cs, ctx->ev_in[set], 0));
    if (k >= 2) CUDA_TRY(cudaStreamWaitEvent(cs, ctx->ev_out[set], 0));
    int rc = b2c_prog_run(p, stream, workspace, workspace_bytes, ext, n_ext);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(ctx->ev_done[set], cs));
    // D2H queue: results of micro-batch k
    CUDA_TRY(cudaStreamWaitEvent(os, ctx->ev_done[set], 0));
    for (int i = 0; i < n_d2h; ++i)
      CUDA_TRY(cudaMemcpyAsync((char*)d2h[i].host + (size_t)k * d2h[i].bytes, ext[d2h[i].slot - 1], d2h[i].bytes,
                               cudaMemcpyDeviceToHost, os));
    CUDA_TRY(cudaEventRecord(ctx->ev_out[set], os));
  }
  CUDA_TRY(cudaStreamSynchronize(os));
  CUDA_TRY(cudaStreamSynchronize(xs));
  CUDA_TRY(cudaStreamSynchronize(cs));
  return B2C_OK;
}

// ------------------------------------------------------------------------------------------
// evaluation metrics of the callers: plain entry points on caller tensors, no program needed
// ------------------------------------------------------------------------------------------
extern "C" int b2c_metric_xcorr_align(int device, void* stream, const float* ref, const float* est, int B, int L,
                                      int max_shift, float* corr, int* best_shift) {
  if (!ref || !est || !corr || !best_shift || B <= 0 || L <= 0 || max_shift < 0 || B > 65535)
    return fail(B2C_ERR_ARG, "b2c_metric_xcorr_align: bad argument");
  DEVICE_GUARD(device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int n_shift = 2 * max_shift + 1;
  xcorr_shifts_f32<<<dim3((n_shift + XC_SPB - 1) / XC_SPB, B), 256, 0, st>>>(ref, est, corr, L, max_shift);
  xcorr_pick_first_max<<<(B + 127) / 128, 128, 0, st>>>(corr, best_shift, B, max_shift);
  CUDA_TRY(cudaPeekAtLastError());
  return B2C_OK;
}

static int resample_args_ok(const char* who, int orig, int nw, int width) {
  if (orig < 1 || nw < 1 || width < 1) return fail(B2C_ERR_ARG, "%s: bad resampling ratio %d -> %d, width %d", who, orig, nw, width);
  if ((size_t)nw * (2 * width + orig) * sizeof(float) > 96 * 1024)
    return fail(B2C_ERR_UNSUPPORTED, "%s: filter bank of %d x %d taps exceeds 96 KB of shared memory", who, nw, 2 * width + orig);
  return B2C_OK;
}

extern "C" int b2c_metric_resample(int device, void* stream, const float* x, float* y, const float* kern, int B, int L,
                                   int Lout, int orig, int nw, int width) {
  if (!x || !y || !kern || B <= 0 || L <= 0 || Lout <= 0 || B > 65535) return fail(B2C_ERR_ARG, "b2c_metric_resample: bad argument");
  int rc = resample_args_ok("b2c_metric_resample", orig, nw, width);
  if (rc) return rc;
  DEVICE_GUARD(device);
  const size_t sm = (size_t)nw * (2 * width + orig) * sizeof(float);
  CUDA_TRY(cudaFuncSetAttribute(resample_sinc_f32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  resample_sinc_f32<<<dim3((Lout + 255) / 256, B), 256, sm, reinterpret_cast<cudaStream_t>(stream)>>>(x, y, kern, L, Lout, orig, nw, width);
  CUDA_TRY(cudaPeekAtLastError());
  return B2C_OK;
}

extern "C" int b2c_metric_psnr(int device, void* stream, const float* ref, const float* est, float* out, int B, int n, float eps) {
  if (!ref || !est || !out || B <= 0 || n <= 0) return fail(B2C_ERR_ARG, "b2c_metric_psnr: bad argument");
  DEVICE_GUARD(device);
  psnr_rows_f32<<<B, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(ref, est, out, n, eps);
  CUDA_TRY(cudaPeekAtLastError());
  return B2C_OK;
}

extern "C" int b2c_metric_psnr_resampled(int device, void* stream, const float* ref, const float* est, const int* shifts,
                                         float* out, const float* kern, int B, int L, int orig, int nw, int width, float eps) {
  if (!ref || !est || !out || !kern || B <= 0 || L <= 0) return fail(B2C_ERR_ARG, "b2c_metric_psnr_resampled: bad argument");
  int rc = resample_args_ok("b2c_metric_psnr_resampled", orig, nw, width);
  if (rc) return rc;
  DEVICE_GUARD(device);
  const size_t sm = (size_t)nw * (2 * width + orig) * sizeof(float);
  CUDA_TRY(cudaFuncSetAttribute(psnr_resampled_f32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  psnr_resampled_f32<<<B, 256, sm, reinterpret_cast<cudaStream_t>(stream)>>>(ref, est, shifts, out, kern, L, orig, nw, width, eps);
  CUDA_TRY(cudaPeekAtLastError());
  return B2C_OK;
}

extern "C" size_t b2c_metric_stsim_scratch_bytes(int B, int L, int n_mels) {
  if (B <= 0 || L <= 0 || n_mels <= 0) return 0;
  const size_t frames = (size_t)L / ST_HOP + 1;
  return ((size_t)B * 2 * frames * n_mels + (size_t)B * 2) * sizeof(float);
}

extern "C" int b2c_metric_stsim(int device, void* stream, const float* ref, const float* est, const float* mel_fb,
                                const int* mel_range, float* scratch, float* out, int B, int L, int n_mels) {
  if (!ref || !est || !mel_fb || !scratch || !out || B <= 0 || n_mels <= 0 || n_mels > 128 || B > 65535)
    return fail(B2C_ERR_ARG, "b2c_metric_stsim: bad argument (n_mels <= 128)");
  if (L <= ST_NFFT / 2) return fail(B2C_ERR_ARG, "b2c_metric_stsim: reflect padding needs more than %d samples (got %d)", ST_NFFT / 2, L);
  DEVICE_GUARD(device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int frames = L / ST_HOP + 1;
  float* mel = scratch;
  float* amax = scratch + (size_t)B * 2 * frames * n_mels;
  CUDA_TRY(cudaMemsetAsync(amax, 0, (size_t)B * 2 * sizeof(float), st));
  stft_mel_pair_f32<<<dim3(frames, B), 256, 0, st>>>(ref, est, mel_fb, mel_range, mel, amax, L, frames, n_mels);
  stsim_from_mel_f32<<<B, 256, 0, st>>>(mel, amax, out, frames, n_mels);
  CUDA_TRY(cudaPeekAtLastError());
  return B2C_OK;
}
