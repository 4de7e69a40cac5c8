/*
 * b2c.h -- C ABI of libb2c.so: the B200 (sm_100a) encode -> quantize -> decode path of
 * the proposed audio/vibrotactile VQ-VAE codec.
 *
 * The reference (aymenboudhina/Multimodal_VQVAE_compression_audio_tactile) has no FFI:
 * its "plugin interface" is a set of PyTorch nn.Module slots
 * (ProposedEval(A_ENC, A_QUANT, T_ENC, T_DEC, ...),
 * Evaluation/dac_vcpwq_proposed6_latency.py:437-449).  The host-side mirror of those
 * modules lives in multimodal_vqvae_compression_audio_tactile_b200/modules.py and binds
 * this header with ctypes; INTEGRATION.md shows the binding.  Each entry point below
 * names the reference computation it replaces.
 *
 * Conventions
 *  - plain pointers and sizes only; no torch types; nothing throws; every function
 *    returns B2C_OK (0) or a negative error code, b2c_last_error() gives the text.
 *  - device tensors are fp32 unless stated.  Activations inside a program are
 *    CHANNEL-LAST: [batch][position][channel]; the module boundary tensors of the
 *    reference ([B, C, L]) are converted by b2c_prog_transpose.
 *  - a *program* is a straight-line list of kernel launches over one workspace;
 *    it is built once per (batch, length) and run many times (b2c_prog_run enqueues
 *    on the caller's stream and never synchronises).
 *  - a buffer reference (b2c_ref) is (slot << 56) | byte_offset.  slot 0 is the
 *    workspace passed to b2c_prog_run, slots 1.. are the caller's external tensors
 *    (ext[slot-1]).  B2C_NULL_REF means "absent".
 *  - packed weights are owned by the context and freed by b2c_ctx_destroy.
 */
#ifndef B2C_H
#define B2C_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b2c_ctx b2c_ctx;
typedef struct b2c_prog b2c_prog;
typedef uint64_t b2c_ref;

#define B2C_ABI_VERSION 4
#define B2C_NULL_REF ((b2c_ref)0xFFFFFFFFFFFFFFFFull)
#define B2C_REF(slot, off) ((((b2c_ref)(slot)) << 56) | (b2c_ref)(off))

#define B2C_OK 0
#define B2C_ERR_ARG (-1)
#define B2C_ERR_CUDA (-2)
#define B2C_ERR_WORKSPACE (-3)
#define B2C_ERR_STATE (-4)
#define B2C_ERR_UNSUPPORTED (-5)

/* epilogue activation written to out_act (out_raw always gets the pre-activation) */
#define B2C_ACT_NONE 0
#define B2C_ACT_SNAKE 1 /* x + sin(alpha x)^2 / (alpha + 1e-9)   (dac Snake1d) */
#define B2C_ACT_GELU 2  /* exact erf GELU (nn.GELU default, :379) */
#define B2C_ACT_TANH 3

/* arithmetic of the contraction */
#define B2C_PREC_F32 0    /* FP32 FFMA on CUDA cores (bit-faithful to fp32 up to summation order) */
#define B2C_PREC_BF16X3 1 /* tcgen05 bf16 hi/lo split, 3 MMAs, fp32 accumulate (>= 16 mantissa bits) */
#define B2C_PREC_BF16 2   /* tcgen05 single-pass bf16, fp32 accumulate */

/* storage of an ACTIVATION buffer (out_act of a producer == x of the consuming contraction).
 * Residual / boundary tensors (out_raw, res, module inputs and outputs) are always fp32. */
#define B2C_FMT_F32 0
#define B2C_FMT_BF16X2 1 /* two bf16 planes: hi at the reference, lo at + n_elements; x = hi + lo (16 mantissa bits) */
#define B2C_FMT_BF16 2   /* one bf16 plane */

/* row addressing modes of the token-wise ops (predictor two-pass schedule) */
#define B2C_ROWS_DENSE 0     /* row n                                               */
#define B2C_ROWS_HEAD 1      /* n = (b, j): row b*Tl + chunk*(j+1)      (chunk heads) */
#define B2C_ROWS_HEAD_PREV 2 /* n = (b, j): row b*Tl + chunk*(j+1) - 1              */
#define B2C_ROWS_ZERO 3      /* no input: zeros                                     */

/* positional-encoding row selection */
#define B2C_PE_NONE 0
#define B2C_PE_CHUNK_POS 1 /* pe[(n % Tl) % chunk]  (PosEnc1D on a chunk slice, :349-351) */
#define B2C_PE_ROW0 2      /* pe[0] */
#define B2C_PE_ROW_N 3     /* pe[n] */

const char* b2c_last_error(void);
int b2c_abi_version(void);
/* developer aid (timing experiments, B2C_TC_DEBUG bit 8): copies out and resets the pipeline-event trace CTA 0 of the
 * fused residual-unit kernel recorded: entries (tag << 56) | (tile << 44) | SM clock.  Returns the entry count. */
int b2c_debug_ru_trace(unsigned long long* dst, int cap);

/* ---------------- context + packed weights ---------------- */
int b2c_ctx_create(int device, b2c_ctx** out);
int b2c_ctx_destroy(b2c_ctx* ctx);
/* bytes of device memory held by packed weights */
size_t b2c_ctx_weight_bytes(const b2c_ctx* ctx);

/* Conv1d / ConvTranspose1d / Linear weights, HOST fp32 pointers in PyTorch layout.
 *   v    : Conv1d [cout, cin, k]; ConvTranspose1d [cin, cout, k]; Linear [cout, cin] (k = 1)
 *   g    : old-style weight-norm gain (weight_g, one per dim-0 slice) or NULL for a plain weight;
 *          the fold w = g * v / ||v|| replaces torch._weight_norm (dac WNConv1d)
 *   bias : [cout] or NULL
 *   transposed : 0 Conv1d, 1 ConvTranspose1d (k = 2*stride, phase-decomposed at pack time)
 * Returns a weight id >= 0. */
int b2c_pack_conv(b2c_ctx* ctx, const float* v, const float* g, const float* bias, int cout, int cin, int k,
                  int transposed, int stride, int padding);
/* a plain fp32 vector / matrix (snake alpha, LayerNorm gamma/beta, positional table ...) */
int b2c_pack_vector(b2c_ctx* ctx, const float* data, size_t n);
/* ResidualVQEMA.books (:412-415): n_books host pointers to [K, D]; also stores 0.5*|e|^2 */
int b2c_pack_codebooks(b2c_ctx* ctx, const float* const* books, int n_books, int K, int D);
/* After an in-place change of ResidualVQEMA.books[book] on the device (ema_step, Training/compare_dacvsproposal_3.py:
 * 264-276): copy the new rows (device pointer, [K, D]) into the packed codebooks and recompute 0.5*|e|^2 on `stream`. */
int b2c_codebooks_refresh(b2c_ctx* ctx, int wid, int book, const float* dev_book, void* stream);
/* dac ResidualVectorQuantize: per stage in_proj (v,g,bias) [d,c,1], out_proj (v,g,bias) [c,d,1],
 * codebook [K,d]; arrays of n_q host pointers each. */
int b2c_pack_dac_rvq(b2c_ctx* ctx, int n_q, int c, int d, int K, const float* const* in_v, const float* const* in_g,
                     const float* const* in_b, const float* const* out_v, const float* const* out_g,
                     const float* const* out_b, const float* const* codebook);

/* ---------------- programs ---------------- */
int b2c_prog_create(b2c_ctx* ctx, b2c_prog** out);
int b2c_prog_destroy(b2c_prog* prog);
int b2c_prog_num_launches(const b2c_prog* prog); /* kernel launches per run */
int b2c_prog_num_ops(const b2c_prog* prog);      /* ops (what b2c_prog_profile times) */

/* 1 when the tcgen05 implicit-GEMM kernel takes this layer (precision BF16X3 / BF16), 0 when only the FP32
 * CUDA-core kernel does, negative on a bad argument.  Callers pick activation formats with it. */
int b2c_conv_tc_eligible(const b2c_ctx* ctx, int wid, int Lin, int stride, int dilation);

/* dac Encoder stem: Conv1d(1, cout, k=7, p=3) on x [B, L] -> [B, L, cout] (out_act stored as act_fmt). */
int b2c_prog_stem(b2c_prog* p, int wid, b2c_ref x, b2c_ref out_raw, b2c_ref out_act, int act, int alpha_wid, int B,
                  int L, int act_fmt);
/* Conv1d (any k, stride, dilation, zero padding) or Linear (k = 1) as an implicit GEMM:
 *   x [B, Lin, cin] -> [B, Lout, cout];  v = conv + bias (+ res);  out_raw = v;  out_act = act(v).
 *   res_mode 0: res has the output's shape; 1: res is a [chunk, cout] table indexed by (lo % Tl) % chunk. */
int b2c_prog_conv(b2c_prog* p, int wid, b2c_ref x, b2c_ref res, b2c_ref out_raw, b2c_ref out_act, int act,
                  int alpha_wid, int B, int Lin, int stride, int dilation, int padding, int res_mode, int Tl,
                  int chunk, int precision, int x_fmt, int act_fmt);
/*   precision F32: x_fmt and act_fmt must be B2C_FMT_F32.  BF16X3: x must be BF16X2.  BF16: x BF16 or BF16X2
 *   (the hi plane is read).  A tensor-core precision on a layer b2c_conv_tc_eligible() rejects is an error:
 *   there is no silent change of arithmetic. */
/* ConvTranspose1d (k = 2*stride, padding = ceil(stride/2)), x [B, Lin, cin] -> [B, Lout, cout]. */
int b2c_prog_convT(b2c_prog* p, int wid, b2c_ref x, b2c_ref out_raw, b2c_ref out_act, int act, int alpha_wid, int B,
                   int Lin, int precision, int x_fmt, int act_fmt);
/* One fused launch for a dac ResidualUnit (snake -> Conv1d k=7 dilated -> snake -> Conv1d k=1 -> + x):
 *   x_act = snake1(x) as bf16 plane(s) (precision BF16X3: two planes, BF16: one), x_raw = x (fp32 residual),
 *   out_raw = y (fp32, may be B2C_NULL_REF), out_act = snake_next(y) stored as act_fmt (bf16 plane(s)).
 * The intermediate h = snake2(conv7(.)) stays in shared memory.  b2c_ru_tc_eligible(): 1 when the weight pair is
 * a k=7 / k=1 pair of equal width C in {64, 96, 128, 192} with 'same' padding. */
int b2c_ru_tc_eligible(const b2c_ctx* ctx, int wid7, int wid1, int precision);
int b2c_prog_ru(b2c_prog* p, int wid7, int alpha2_wid, int wid1, b2c_ref x_act, b2c_ref x_raw, b2c_ref out_raw,
                b2c_ref out_act, int alpha_next_wid, int B, int L, int dilation, int precision, int act_fmt);
/* dac Decoder head: Conv1d(cin, 1, k=7, p=3) + tanh on x [B, L, cin] -> y [B, L]. */
int b2c_prog_head(b2c_prog* p, int wid, b2c_ref x, b2c_ref y, int B, int L, int x_fmt);
/* Backward-data pass of the decoder (Training/compare_dacvsproposal_3.py:386-409 back-propagates the
 * loss through the frozen T_DEC into predict / proj_*).  One op covers "snake -> conv" of the forward:
 *   out = conv(x; w) * snake'(pre; alpha) + res,   snake'(v) = 1 + sin(2 alpha v) alpha / (alpha + 1e-9)
 * x = gradient w.r.t. the forward conv's output (activation format x_fmt), w = that conv's backward-data form packed as
 * a bias-free Conv1d (Conv1d: channels transposed and taps flipped, padding dil*(k-1) - p; ConvTranspose1d: the same
 * tensor read as Conv1d [cin, cout, k] with the forward stride and padding), pre = the forward input of the snake
 * (fp32, shape of the output), res = gradient arriving over the skip connection (optional).  out_raw fp32 and / or
 * out_act (act_fmt: what the next backward contraction reads). */
int b2c_prog_conv_dsnake(b2c_prog* p, int wid, b2c_ref x, b2c_ref pre, int alpha_wid, b2c_ref res, b2c_ref out_raw,
                         b2c_ref out_act, int B, int Lin, int stride, int dilation, int padding, int precision, int x_fmt,
                         int act_fmt);
/* Backward-data of the decoder tail snake -> Conv1d(C, 1, 7, p=3) -> tanh: g_y, y [B, L]; x_raw [B, L, C] = the
 * forward input of the last snake; g_raw fp32 / g_act (act_fmt) [B, L, C] = gradient w.r.t. x_raw. */
int b2c_prog_head_bwd(b2c_prog* p, int wid, int alpha_wid, b2c_ref g_y, b2c_ref y, b2c_ref x_raw, b2c_ref g_raw,
                      b2c_ref g_act, int B, int L, int act_fmt);
/* LayerNorm over C of rows gathered by a_mode from a (minus sub, plus pe row), optional scale*tanh.
 * nn.LayerNorm eps = 1e-5.  (CrossPredictor.ln_q/ln_kv/ffn[0] :394-395,:377; TokenNorm :357-360 with
 * tanh and the clamped scalar, :472-474) */
int b2c_prog_layernorm(b2c_prog* p, int gamma_wid, int beta_wid, b2c_ref a, int a_mode, b2c_ref sub, int pe_wid,
                       int pe_mode, int tanh_post, float post_scale, b2c_ref out, int N, int C, int Tl, int chunk,
                       int out_fmt);
/* The same LayerNorm on dense rows with a token mask: rows whose row_mask byte is non-zero read `a` as zeros
 * (AllPredPLC.forward_step: zt_in = zt_full * ~mask, PLC/PLC1_eval.py:497, then CrossPredictor's pos + ln_q :402-405). */
int b2c_prog_layernorm_masked(b2c_prog* p, int gamma_wid, int beta_wid, b2c_ref a, b2c_ref row_mask, int pe_wid,
                              int pe_mode, b2c_ref out, int N, int C, int Tl, int chunk, int out_fmt);
/* softmax(Q K^T / sqrt(dh)) V over ALL T keys of a frame: the packet-loss-concealment forward calls CrossPredictor
 * once over the whole file (PLC/PLC1_eval.py:500, attention :411-412; PosEnc1D max_len 8192).  q [B*T, heads*dh],
 * kv [B*T, 2*heads*dh] (K | V), out [B*T, heads*dh].  Flash-style online softmax: the [T, T] scores never exist. */
int b2c_prog_attention_full(b2c_prog* p, b2c_ref q, b2c_ref kv, b2c_ref out, int B, int T, int heads, int dh);
/* out[n, :] = row_mask[n] ? a[n, :] : b[n, :]   (z_filled = torch.where(mask, z_pred, zt_in), PLC/PLC1_eval.py:503) */
int b2c_prog_select_rows(b2c_prog* p, b2c_ref row_mask, b2c_ref a, b2c_ref b, b2c_ref out, int N, int C);
/* One book of ResidualVQEMA.ema_step (Training/compare_dacvsproposal_3.py:264-276) given idx = nearest code of every
 * row (b2c_prog_nearest): emb[k] = decay*emb[k] + one_minus_decay * mean(x[idx == k]) for every code that was hit; the
 * rows of a code are added in row order like index_add_ on the CPU.  x [N, D], idx int32 [N], emb [K, D] updated IN
 * PLACE, counts int32 [K] (optional: bincount). */
int b2c_prog_ema_update(b2c_prog* p, b2c_ref x, b2c_ref idx, b2c_ref emb, b2c_ref counts, int N, int D, int K, float decay,
                        float one_minus_decay);
/* softmax(Q K^T / sqrt(dh)) V inside each chunk (:401-402).
 *   kv [B*Tl, 2*heads*dh] (K | V).  q_mode 0: q is a [chunk, heads*dh] table (row = position in chunk),
 *   out rows dense [B*Tl].  q_mode 1: q is [B*nfix, heads*dh], one query per chunk head, out [B*nfix]. */
int b2c_prog_attention(b2c_prog* p, b2c_ref q, int q_mode, b2c_ref kv, b2c_ref out, int B, int Tl, int chunk,
                       int heads, int dh);
/* ResidualVQEMA.forward (:421-435) on rows x [N, D]; qsum [N, D]; idx int32 laid out [B, books_use, Tl]
 * (row_mode DENSE: n = b*Tl + t; HEAD: n = (b, j) -> t = chunk*(j+1)).
 *   precision B2C_PREC_F32: FP32 CUDA-core kernels (FFMA scores, first maximum).  BF16X3 / BF16: ONE tcgen05 launch
 *   for all books (rvq_tc_kernel: bf16x3 score GEMM per book with the scores in TMEM, arg-max + codeword gather +
 *   q_sum / residual update in the epilogue, candidates within the contraction's error bound re-scored with the FP32
 *   kernels' arithmetic) for D in {32, 64, 96, 128}, K a multiple of 64 up to 512; other shapes run the FP32
 *   kernels -- the indices and q_sum are the same bits either way.
 *   scratch: b2c_rvq_scratch_bytes(N, D) bytes of workspace (used by the small-batch FP32 kernels). */
size_t b2c_rvq_scratch_bytes(int N, int D);
int b2c_prog_rvq(b2c_prog* p, int books_wid, int books_use, b2c_ref x, b2c_ref qsum, b2c_ref idx, b2c_ref scratch, int N,
                 int row_mode, int B, int Tl, int chunk, int precision);
/* Receiver side of ResidualVQEMA.forward (:421-435): the code indices are an INPUT (int32, the layout b2c_prog_rvq
 * writes), qsum[n] = sum over the first books_use books of book[idx] (plain fp32 adds in book order).  Not in the
 * reference, which never decodes from indices; it is what "indices out, reconstruction in" needs. */
int b2c_prog_rvq_lookup(b2c_prog* p, int books_wid, int books_use, b2c_ref idx, b2c_ref qsum, int N, int row_mode,
                        int B, int Tl, int chunk);
/* ResidualVQEMA._nearest_l2 (:417-419) with caller tensors: x [N, D], emb [K, D] -> idx int32 [N].
 * scratch: b2c_nearest_scratch_bytes(N, D, K, precision) bytes.
 *   B2C_PREC_F32: FFMA scores, 32 rows per CTA.  B2C_PREC_BF16X3 / BF16: tcgen05 score GEMM (bf16 hi/lo split,
 *   3 MMAs, fp32 accumulate in TMEM) with the arg-max in the epilogue -- the [N, K] score matrix never exists
 *   in memory -- then the candidates within the contraction's error bound of the maximum are re-scored with
 *   the FP32 kernel's arithmetic, so both precisions return the same indices. */
size_t b2c_nearest_scratch_bytes(int N, int D, int K, int precision);
int b2c_nearest_tc_eligible(int N, int D, int K);
int b2c_prog_nearest(b2c_prog* p, b2c_ref x, b2c_ref emb, b2c_ref scratch, b2c_ref idx, int N, int D, int K,
                     int precision);
/* dac ResidualVectorQuantize.forward (eval), z [B*Tl, c] -> zq [B*Tl, c], codes int32 [B, n_q, Tl]. */
int b2c_prog_dac_rvq(b2c_prog* p, int wid, int n_q, b2c_ref z, b2c_ref zq, b2c_ref codes, int B, int Tl);
/* out[b*Tl + chunk*(j+1)] = src[b*nfix + j] for rows of width C (second pass write-back). */
int b2c_prog_scatter_heads(b2c_prog* p, b2c_ref src, b2c_ref dst, int B, int Tl, int chunk, int C);
/* [B, R, C] -> [B, C, R] */
int b2c_prog_transpose(b2c_prog* p, b2c_ref in, b2c_ref out, int B, int R, int C);
/* change the storage format of an n-element activation buffer (fp32 <-> bf16 planes) */
int b2c_prog_convert(b2c_prog* p, b2c_ref src, int src_fmt, b2c_ref dst, int dst_fmt, size_t n);
/* widen int32 -> int64 (PyTorch index dtype at the module boundary) */
int b2c_prog_i32_to_i64(b2c_prog* p, b2c_ref in, b2c_ref out, size_t n);

/* Two launch queues inside one program (batch-1 streaming: the two encoders are independent until the predictor and
 * neither fills the GPU alone).  Ops added after b2c_prog_set_lane(p, 1) are enqueued on a second stream owned by the
 * context, forked from the caller's stream at the first of them; after b2c_prog_set_lane(p, 0) on the caller's stream
 * again; b2c_prog_join makes the caller's stream wait for the second queue (b2c_prog_run joins at the end in any case).
 * Buffers written on one lane and read on the other must not be touched in between; arithmetic is unaffected. */
int b2c_prog_set_lane(b2c_prog* p, int lane);
int b2c_prog_join(b2c_prog* p);

/* Enqueue the program on `stream` (a cudaStream_t).  ext[i] is the device pointer of slot i+1. */
int b2c_prog_run(b2c_prog* p, void* stream, void* workspace, size_t workspace_bytes, void* const* ext, int n_ext);

/* Run the program once with a cudaEvent pair around every launch (a measuring aid for bench.py and
 * profiles/: never used on the timed path).  Fills up to cap entries:
 *   ms[i] device time of launch i; kind[i] one of B2C_KIND_*; flops[i] / bytes[i] the ALGORITHMIC
 *   work of that launch (2*M*N*K for contractions; minimal global reads+writes).  Returns the number
 *   of launches.  Synchronises the stream. */
#define B2C_KIND_CONV_F32 1
#define B2C_KIND_CONV_TC 2
#define B2C_KIND_STEM 3
#define B2C_KIND_HEAD 4
#define B2C_KIND_LAYERNORM 5
#define B2C_KIND_ATTENTION 6
#define B2C_KIND_RVQ 7
#define B2C_KIND_NEAREST 8
#define B2C_KIND_DAC_RVQ 9
#define B2C_KIND_MOVE 10
#define B2C_KIND_CONV_TC_X3 11 /* tcgen05 contraction in bf16x3: 3 MMA FLOPs per algorithmic FLOP */
int b2c_prog_profile(b2c_prog* p, void* stream, void* workspace, size_t workspace_bytes, void* const* ext, int n_ext,
                     float* ms, int* kind, double* flops, double* bytes, int cap);

typedef struct {
  void* host;   /* pinned or pageable host buffer */
  int slot;     /* external slot (>= 1) whose device buffer is the other end */
  size_t bytes;
} b2c_hostcopy;
/* Host-buffer entry: H2D copies, the program, D2H copies, then a stream synchronise.
 * This is the call bench.py's e2e leg and a non-PyTorch host would make. */
int b2c_prog_run_host(b2c_prog* p, void* stream, void* workspace, size_t workspace_bytes, void* const* ext, int n_ext,
                      const b2c_hostcopy* h2d, int n_h2d, const b2c_hostcopy* d2h, int n_d2h);

/* The same for n_micro equal micro-batches with copies and compute overlapped: ext_sets holds TWO sets of n_ext
 * device pointers (double-buffered staging, [2][n_ext]); host pointers of h2d / d2h advance by their `bytes` per
 * micro-batch.  H2D of micro-batch k+1 and D2H of k-1 run on a copy stream owned by the context while the program
 * of micro-batch k runs on `stream`.  Synchronises both streams before returning. */
int b2c_prog_run_host_pipelined(b2c_prog* p, void* stream, void* workspace, size_t workspace_bytes, void* const* ext_sets,
                                int n_ext, const b2c_hostcopy* h2d, int n_h2d, const b2c_hostcopy* d2h, int n_d2h,
                                int n_micro);

/* ---------------- evaluation metrics of the callers (Evaluation/compare_dacvsproposal_5_eval.py) ----------------
 * Plain entry points on caller tensors (device fp32 pointers, rows of L samples), enqueued on `stream`, no sync. */
/* align_pair_24k (:188-211) for B frames: corr [B, 2*max_shift+1] = sum_j ref[j]*est[j+s], s = -max_shift..max_shift,
 * best_shift int32 [B] = the first s with the strictly largest correlation (the reference's Python loop order). */
int b2c_metric_xcorr_align(int device, void* stream, const float* ref, const float* est, int B, int L, int max_shift,
                           float* corr, int* best_shift);
/* resample_f32 (:91-97) = torchaudio sinc_interp_hann: kern [nw, 2*width+orig] is the polyphase filter bank
 * (orig, nw = the two rates divided by their gcd), y [B, Lout], Lout = ceil(nw*L/orig). */
int b2c_metric_resample(int device, void* stream, const float* x, float* y, const float* kern, int B, int L, int Lout,
                        int orig, int nw, int width);
/* psnr_batch (:180-185): out[b] = 10 log10(1 / max(mean((ref-est)^2), eps)) over n samples per row. */
int b2c_metric_psnr(int device, void* stream, const float* ref, const float* est, float* out, int B, int n, float eps);
/* psnr_3k_aligned_batch (:213-223) in one launch: rows aligned by shifts[b] (NULL: none) as align_pair_24k does,
 * both resampled on the fly, only the PSNR leaves the kernel. */
int b2c_metric_psnr_resampled(int device, void* stream, const float* ref, const float* est, const int* shifts, float* out,
                              const float* kern, int B, int L, int orig, int nw, int width, float eps);
/* stsim_batch (:166-177) with _mel_mag (:142-163): STFT n_fft 512 / hop 128 / periodic hann / centre + reflect,
 * |.| clamped at 1e-8, mel_fb [257, n_mels], normalised by the per-signal maximum, per-frame cosine, 0.5*(mean+1).
 * mel_range (optional, int32 [n_mels][2]): bins [lo, hi) outside which a band's filter is zero.
 * scratch: b2c_metric_stsim_scratch_bytes(B, L, n_mels) bytes. */
size_t b2c_metric_stsim_scratch_bytes(int B, int L, int n_mels);
int b2c_metric_stsim(int device, void* stream, const float* ref, const float* est, const float* mel_fb,
                     const int* mel_range, float* scratch, float* out, int B, int L, int n_mels);

#ifdef __cplusplus
}
#endif
#endif /* B2C_H */
