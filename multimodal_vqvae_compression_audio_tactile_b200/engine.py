"""Host-side engine: packs module weights into a libb2c context and emits the
straight-line programs (lists of CUDA kernel launches) for the encoder, the DAC
quantizer, the predictor/residual-VQ two-pass schedule and the decoder.

PyTorch is used here for device memory and streams only; all arithmetic runs in
libb2c.so (hand-written sm_100a CUDA).  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import warnings
from dataclasses import dataclass, field

import numpy as np
import torch

from . import _lib as L

ALIGN = 256


def _np32(t: torch.Tensor) -> np.ndarray:
    return np.ascontiguousarray(t.detach().to("cpu", torch.float32).numpy())


def _fp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class Arena:
    """First-fit allocator over the program workspace (offsets only; lifetimes are known
    while the program is being emitted)."""

    def __init__(self):
        self.free_list = []   # (off, size)
        self.top = 0
        self.peak = 0
        self.live = {}

    #: extra bytes added to every allocation (experiment knob: relative placement of the streams a kernel reads
    #: and writes decides DRAM bank conflicts)
    PAD = int(os.environ.get("B2C_ARENA_PAD", "0"))

    def alloc(self, nbytes: int) -> int:
        n = (max(int(nbytes), 1) + self.PAD + ALIGN - 1) // ALIGN * ALIGN
        for i, (off, size) in enumerate(self.free_list):
            if size >= n:
                if size == n:
                    self.free_list.pop(i)
                else:
                    self.free_list[i] = (off + n, size - n)
                self.live[off] = n
                return off
        off = self.top
        self.top += n
        self.peak = max(self.peak, self.top)
        self.live[off] = n
        return off

    def free(self, off):
        if off is None:
            return
        n = self.live.pop(off)
        self.free_list.append((off, n))
        self.free_list.sort()
        merged = []
        for o, s in self.free_list:
            if merged and merged[-1][0] + merged[-1][1] == o:
                merged[-1] = (merged[-1][0], merged[-1][1] + s)
            else:
                merged.append((o, s))
        if merged and merged[-1][0] + merged[-1][1] == self.top:
            self.top = merged[-1][0]
            merged.pop()
        self.free_list = merged


@dataclass
class ConvW:
    wid: int
    cin: int
    cout: int
    k: int
    stride: int = 1
    dilation: int = 1
    padding: int = 0
    transposed: bool = False


@dataclass
class Program:
    handle: C.c_void_p
    ws_bytes: int
    n_ext: int
    info: dict = field(default_factory=dict)
    pinned: bool = False      # captured in a CUDA graph: never evicted


class ProgramCache:
    """key -> Program, least-recently-used eviction: callers with ever-changing shapes (per-file PLC / latency
    evaluation) would otherwise grow the set of built programs without bound.  Evicted programs are destroyed."""

    def __init__(self, lib, capacity: int = int(os.environ.get("B2C_PROGRAM_CACHE", "48"))):
        self.lib, self.capacity = lib, max(1, capacity)
        self._d = {}

    def get(self, key, default=None):
        p = self._d.get(key)
        if p is None:
            return default
        self._d[key] = self._d.pop(key)      # most recently used last
        return p

    def __setitem__(self, key, prog: Program):
        old = self._d.pop(key, None)
        if old is not None and old is not prog:
            self.lib.b2c_prog_destroy(old.handle)
        self._d[key] = prog
        if len(self._d) > self.capacity:
            for k in [k for k, v in self._d.items() if not v.pinned][: len(self._d) - self.capacity]:
                self.lib.b2c_prog_destroy(self._d.pop(k).handle)

    def __len__(self):
        return len(self._d)

    def values(self):
        return self._d.values()

    def clear(self):
        while self._d:
            _, p = self._d.popitem()
            self.lib.b2c_prog_destroy(p.handle)


class Engine:
    """One libb2c context (= one set of packed weights on one device)."""

    def __init__(self, device: torch.device):
        if device.type != "cuda":
            raise L.B2CError("b200 codec modules only run on CUDA tensors (sm_100a); there is no CPU path")
        self.lib = L.load()
        self.device = device
        idx = device.index if device.index is not None else torch.cuda.current_device()
        h = C.c_void_p()
        L.check(self.lib.b2c_ctx_create(idx, C.byref(h)), "b2c_ctx_create")
        self.ctx = h
        self.programs = ProgramCache(self.lib)
        #: records that live as long as the packed weights but are not programs (CUDA-graph captures, host staging)
        self.aux = {}
        self._ws = None
        self._keep = []  # host arrays kept alive during packing
        #: layers a tensor-core plan had to run on the FP32 kernel (see Emitter._contract)
        self.fp32_reroutes = []

    def close(self):
        """Free the programs, the packed weights and the context (idempotent)."""
        ctx, self.ctx = self.ctx, None
        if ctx is None:
            return
        self.aux.clear()
        try:
            self.programs.clear()
        finally:
            self.lib.b2c_ctx_destroy(ctx)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------ packing ------------------------------
    def pack_wnconv(self, m) -> ConvW:
        """m: modules.WNConv1d / WNConvTranspose1d (weight_g, weight_v, bias)."""
        v, g = _np32(m.weight_v), _np32(m.weight_g).reshape(-1)
        b = _np32(m.bias) if m.bias is not None else None
        tr = bool(getattr(m, "transposed", False))
        if tr:
            cin, cout, k = v.shape
        else:
            cout, cin, k = v.shape
        wid = L.check(self.lib.b2c_pack_conv(self.ctx, _fp(v), _fp(g), _fp(b) if b is not None else None,
                                             cout, cin, k, int(tr), m.stride, m.padding), "b2c_pack_conv")
        return ConvW(wid, cin, cout, k, m.stride, m.dilation, m.padding, tr)

    def pack_plain(self, weight: torch.Tensor, bias=None) -> ConvW:
        """nn.Linear [cout, cin] or nn.Conv1d(k=1) [cout, cin, 1] weights."""
        v = _np32(weight)
        if v.ndim == 2:
            v = v[:, :, None]
        cout, cin, k = v.shape
        v = np.ascontiguousarray(v)
        b = _np32(bias) if bias is not None else None
        wid = L.check(self.lib.b2c_pack_conv(self.ctx, _fp(v), None, _fp(b) if b is not None else None,
                                             cout, cin, k, 0, 1, 0), "b2c_pack_conv")
        return ConvW(wid, cin, cout, k)

    def pack_vec(self, t: torch.Tensor) -> int:
        a = _np32(t).reshape(-1)
        return L.check(self.lib.b2c_pack_vector(self.ctx, _fp(a), a.size), "b2c_pack_vector")

    def pack_books(self, books) -> int:
        arrs = [_np32(b) for b in books]
        k, d = arrs[0].shape
        ptrs = (C.POINTER(C.c_float) * len(arrs))(*[_fp(a) for a in arrs])
        return L.check(self.lib.b2c_pack_codebooks(self.ctx, ptrs, len(arrs), k, d), "b2c_pack_codebooks")

    def pack_dac_rvq(self, quantizers) -> int:
        keep = []

        def col(fn):
            arrs = [fn(q) for q in quantizers]
            keep.append(arrs)
            return (C.POINTER(C.c_float) * len(arrs))(*[_fp(a) for a in arrs])

        n_q = len(quantizers)
        q0 = quantizers[0]
        d, c, _ = q0.in_proj.weight_v.shape
        k = q0.codebook.weight.shape[0]
        return L.check(self.lib.b2c_pack_dac_rvq(
            self.ctx, n_q, c, d, k,
            col(lambda q: _np32(q.in_proj.weight_v)), col(lambda q: _np32(q.in_proj.weight_g).reshape(-1)),
            col(lambda q: _np32(q.in_proj.bias)),
            col(lambda q: _np32(q.out_proj.weight_v)), col(lambda q: _np32(q.out_proj.weight_g).reshape(-1)),
            col(lambda q: _np32(q.out_proj.bias)),
            col(lambda q: _np32(q.codebook.weight))), "b2c_pack_dac_rvq")

    def weight_bytes(self) -> int:
        return int(self.lib.b2c_ctx_weight_bytes(self.ctx))

    # ------------------------------ running ------------------------------
    def workspace(self, nbytes: int) -> torch.Tensor:
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = None
            self._ws = torch.empty(max(nbytes, ALIGN), dtype=torch.uint8, device=self.device)
        return self._ws

    def workspace_side(self, nbytes: int) -> torch.Tensor:
        """Second workspace of two-lane programs (the side lane's activations: nothing the main lane allocates while
        both run can alias them)."""
        ws = getattr(self, "_ws_side", None)
        if ws is None or ws.numel() < nbytes:
            self._ws_side = None
            self._ws_side = ws = torch.empty(max(nbytes, ALIGN), dtype=torch.uint8, device=self.device)
        return ws

    def _ext(self, prog: Program, ext_ptrs):
        """external pointers of a run: a two-lane program takes the side workspace as its last slot"""
        side = prog.info.get("side_bytes", 0)
        if side:
            return list(ext_ptrs) + [self.workspace_side(side).data_ptr()]
        return ext_ptrs

    def run(self, prog: Program, ext_ptrs):
        ext_ptrs = self._ext(prog, ext_ptrs)
        ws = self.workspace(prog.ws_bytes)
        arr = (C.c_void_p * max(len(ext_ptrs), 1))(*ext_ptrs)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        L.check(self.lib.b2c_prog_run(prog.handle, C.c_void_p(stream), C.c_void_p(ws.data_ptr()), ws.numel(), arr,
                                      len(ext_ptrs)), "b2c_prog_run")

    def profile(self, prog: Program, ext_ptrs):
        """Per-launch device times with cudaEvent pairs (measuring aid, not the timed path).
        -> list of dicts {kind, ms, flops, bytes}."""
        ext_ptrs = self._ext(prog, ext_ptrs)
        ws = self.workspace(prog.ws_bytes)
        arr = (C.c_void_p * max(len(ext_ptrs), 1))(*ext_ptrs)
        cap = prog.info["ops"] + 8
        ms = (C.c_float * cap)()
        kind = (C.c_int * cap)()
        fl = (C.c_double * cap)()
        by = (C.c_double * cap)()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        n = L.check(self.lib.b2c_prog_profile(prog.handle, C.c_void_p(stream), C.c_void_p(ws.data_ptr()), ws.numel(),
                                              arr, len(ext_ptrs), ms, kind, fl, by, cap), "b2c_prog_profile")
        return [dict(kind=L.KIND_NAMES.get(kind[i], "?"), ms=ms[i], flops=fl[i], bytes=by[i]) for i in range(min(n, cap))]

    def run_host(self, prog: Program, ext_ptrs, h2d, d2h):
        """h2d / d2h: lists of (host_ptr, slot, nbytes).  Synchronises the stream."""
        ext_ptrs = self._ext(prog, ext_ptrs)
        ws = self.workspace(prog.ws_bytes)
        arr = (C.c_void_p * max(len(ext_ptrs), 1))(*ext_ptrs)
        hin = (L.HostCopy * max(len(h2d), 1))(*[L.HostCopy(p, s, n) for p, s, n in h2d])
        hout = (L.HostCopy * max(len(d2h), 1))(*[L.HostCopy(p, s, n) for p, s, n in d2h])
        stream = torch.cuda.current_stream(self.device).cuda_stream
        L.check(self.lib.b2c_prog_run_host(prog.handle, C.c_void_p(stream), C.c_void_p(ws.data_ptr()), ws.numel(),
                                           arr, len(ext_ptrs), hin, len(h2d), hout, len(d2h)), "b2c_prog_run_host")


    def run_host_pipelined(self, prog: Program, ext_sets, h2d, d2h, n_micro: int):
        """n_micro equal micro-batches, copies overlapped with compute (b2c_prog_run_host_pipelined).
        ext_sets: two lists of device pointers; h2d / d2h: (host_ptr of micro-batch 0, slot, bytes per micro-batch)."""
        ext_sets = [self._ext(prog, ext_sets[0]), self._ext(prog, ext_sets[1])]
        ws = self.workspace(prog.ws_bytes)
        n_ext = len(ext_sets[0])
        arr = (C.c_void_p * (2 * n_ext))(*(list(ext_sets[0]) + list(ext_sets[1])))
        hin = (L.HostCopy * max(len(h2d), 1))(*[L.HostCopy(p, s, n) for p, s, n in h2d])
        hout = (L.HostCopy * max(len(d2h), 1))(*[L.HostCopy(p, s, n) for p, s, n in d2h])
        stream = torch.cuda.current_stream(self.device).cuda_stream
        L.check(self.lib.b2c_prog_run_host_pipelined(prog.handle, C.c_void_p(stream), C.c_void_p(ws.data_ptr()), ws.numel(),
                                                     arr, n_ext, hin, len(h2d), hout, len(d2h), n_micro),
                "b2c_prog_run_host_pipelined")


class Emitter:
    """Builds one program.  Buffers are workspace offsets (ints) or ('ext', slot, off)."""

    def __init__(self, eng: Engine):
        self.eng = eng
        self.lib = eng.lib
        h = C.c_void_p()
        L.check(self.lib.b2c_prog_create(eng.ctx, C.byref(h)), "b2c_prog_create")
        self.h = h
        self.arena = Arena()
        self.fp32_reroutes = 0
        # second launch queue (b2c_prog_set_lane): its buffers live in a second workspace, passed as external slot
        # `side_slot`, with its own arena -- see Engine.workspace_side
        self.arena_side = Arena()
        self.side_slot = None
        self._lane = 0

    def lane(self, lane: int, side_slot: int = None):
        """Ops (and buffers) emitted from here on go to launch queue `lane` (0 = the caller's stream)."""
        if lane == 1:
            if side_slot is None and self.side_slot is None:
                raise L.B2CError("the side lane needs an external slot for its workspace")
            self.side_slot = side_slot if side_slot is not None else self.side_slot
        L.check(self.lib.b2c_prog_set_lane(self.h, lane), "b2c_prog_set_lane")
        self._lane = lane

    def join(self):
        L.check(self.lib.b2c_prog_join(self.h), "b2c_prog_join")
        self._lane = 0

    # buffers
    def new(self, nfloats: int):
        if self._lane == 1:
            return ("ext", self.side_slot, self.arena_side.alloc(4 * nfloats))
        return self.arena.alloc(4 * nfloats)

    def drop(self, *offs):
        for o in offs:
            if isinstance(o, int):
                self.arena.free(o)
            elif isinstance(o, tuple) and self.side_slot is not None and o[1] == self.side_slot:
                self.arena_side.free(o[2])

    @staticmethod
    def ext(slot: int, off_bytes: int = 0):
        return ("ext", slot, off_bytes)

    @staticmethod
    def _r(buf):
        if buf is None:
            return L.NULL_REF
        if isinstance(buf, tuple):
            return L.ref(buf[1], buf[2])
        return L.ref(0, buf)

    def finish(self, n_ext: int, **info) -> Program:
        if self.side_slot is not None:
            info["side_bytes"] = self.arena_side.peak + ALIGN
        return Program(self.h, self.arena.peak + ALIGN, n_ext,
                       dict(info, launches=self.lib.b2c_prog_num_launches(self.h), ops=self.lib.b2c_prog_num_ops(self.h),
                            fp32_reroutes=self.fp32_reroutes))

    # ops
    def stem(self, w: ConvW, x, out_raw, out_act, alpha, B, Lx, act_fmt=L.FMT_F32):
        L.check(self.lib.b2c_prog_stem(self.h, w.wid, self._r(x), self._r(out_raw), self._r(out_act),
                                       L.ACT_SNAKE if alpha is not None else L.ACT_NONE,
                                       alpha if alpha is not None else -1, B, Lx, act_fmt), "b2c_prog_stem")

    def convert(self, src, src_fmt, dst, dst_fmt, n):
        L.check(self.lib.b2c_prog_convert(self.h, self._r(src), src_fmt, self._r(dst), dst_fmt, n),
                "b2c_prog_convert")

    def tc_ok(self, w: ConvW, Lin, prec) -> bool:
        """Does the tcgen05 kernel take this layer at this precision?  (else: the FP32 CUDA-core kernel)"""
        if prec == L.PREC_F32:
            return False
        return L.check(self.lib.b2c_conv_tc_eligible(self.eng.ctx, w.wid, Lin, w.stride, w.dilation),
                       "b2c_conv_tc_eligible") == 1

    def _contract(self, w: ConvW, x, B, Lin, Lout, x_fmt, act_fmt, out_act, prec, emit, quiet=False):
        """Shared by conv / convT: run the contraction at `prec` when the tensor-core kernel takes the
        layer, else on the FP32 kernel, converting activation storage on either side when the caller's
        formats differ from what that kernel reads / writes."""
        if prec != L.PREC_F32 and not self.tc_ok(w, Lin, prec):
            # Not silent: the layer keeps its results (the FP32 kernel is the more exact arithmetic) but costs several
            # times the tcgen05 time.  Recorded on the engine and in the program's info, warned once per layer shape;
            # B2C_STRICT_PRECISION=1 turns it into an error.
            what = (f"conv {w.cin}->{w.cout} k={w.k} stride={w.stride} dil={w.dilation} at Lin={Lin}: not eligible for "
                    f"the tcgen05 kernel (precision {prec}); running on the FP32 CUDA-core kernel")
            if os.environ.get("B2C_STRICT_PRECISION", "0") == "1" and not quiet:
                raise L.B2CError(what)
            if what not in self.eng.fp32_reroutes and not quiet:
                self.eng.fp32_reroutes.append(what)
                warnings.warn("b200 codec: " + what, RuntimeWarning, stacklevel=3)
            self.fp32_reroutes += 1
            prec = L.PREC_F32
        need = L.FMT_OF_PREC[prec]
        n_in, n_out = B * Lin * w.cin, B * Lout * w.cout
        tmp_in = tmp_out = None
        x_ok = x_fmt == need or (prec == L.PREC_BF16 and x_fmt == L.FMT_BF16X2)
        if not x_ok:
            if L.FMT_F32 not in (x_fmt, need):
                raise L.B2CError(f"activation format {x_fmt} cannot feed a precision-{prec} contraction")
            tmp_in = self.new(n_in)
            self.convert(x, x_fmt, tmp_in, need, n_in)
            x, x_fmt = tmp_in, need
        k_act_fmt, k_out_act = act_fmt, out_act
        if out_act is not None and prec == L.PREC_F32 and act_fmt != L.FMT_F32:
            tmp_out = self.new(n_out)
            k_act_fmt, k_out_act = L.FMT_F32, tmp_out
        if out_act is None:
            k_act_fmt = L.FMT_F32
        emit(x, x_fmt, k_out_act, k_act_fmt, prec)
        if tmp_out is not None:
            self.convert(tmp_out, L.FMT_F32, out_act, act_fmt, n_out)
        self.drop(tmp_in, tmp_out)

    def conv(self, w: ConvW, x, B, Lin, *, res=None, out_raw=None, out_act=None, act=L.ACT_NONE, alpha=None,
             res_mode=0, Tl=0, chunk=0, prec=L.PREC_F32, x_fmt=L.FMT_F32, act_fmt=L.FMT_F32):
        if alpha is not None:
            act = L.ACT_SNAKE
        Lout = (Lin + 2 * w.padding - w.dilation * (w.k - 1) - 1) // w.stride + 1

        def emit(xb, xf, ob, of, pr):
            L.check(self.lib.b2c_prog_conv(self.h, w.wid, self._r(xb), self._r(res), self._r(out_raw), self._r(ob),
                                           act, alpha if alpha is not None else -1, B, Lin, w.stride, w.dilation,
                                           w.padding, res_mode, Tl, chunk, pr, xf, of), "b2c_prog_conv")

        self._contract(w, x, B, Lin, Lout, x_fmt, act_fmt, out_act, prec, emit)

    def convT(self, w: ConvW, x, B, Lin, *, out_raw=None, out_act=None, alpha=None, prec=L.PREC_F32,
              x_fmt=L.FMT_F32, act_fmt=L.FMT_F32):
        Lout = (Lin - 1) * w.stride - 2 * w.padding + w.k

        def emit(xb, xf, ob, of, pr):
            L.check(self.lib.b2c_prog_convT(self.h, w.wid, self._r(xb), self._r(out_raw), self._r(ob),
                                            L.ACT_SNAKE if alpha is not None else L.ACT_NONE,
                                            alpha if alpha is not None else -1, B, Lin, pr, xf, of), "b2c_prog_convT")

        self._contract(w, x, B, Lin, Lout, x_fmt, act_fmt, out_act, prec, emit)

    def conv_dsnake(self, w: ConvW, x, pre, alpha, B, Lin, *, res=None, out_raw=None, out_act=None, prec=L.PREC_F32,
                    x_fmt=L.FMT_F32, act_fmt=L.FMT_F32):
        """Backward-data of "snake -> conv": out = conv(x; w) * snake'(pre; alpha) + res (b2c_prog_conv_dsnake).
        w is the conv's backward-data form (PackedDecoderBwd)."""
        Lout = (Lin + 2 * w.padding - w.dilation * (w.k - 1) - 1) // w.stride + 1

        def emit(xb, xf, ob, of, pr):
            L.check(self.lib.b2c_prog_conv_dsnake(self.h, w.wid, self._r(xb), self._r(pre), alpha, self._r(res),
                                                  self._r(out_raw), self._r(ob), B, Lin, w.stride, w.dilation, w.padding,
                                                  pr, xf, of), "b2c_prog_conv_dsnake")

        # a strided contraction whose input length is not a multiple of the stride (the 2999-sample stage of the
        # decoder) has no TMA view: it runs on the FP32 kernel by construction, not as a surprise
        self._contract(w, x, B, Lin, Lout, x_fmt, act_fmt, out_act, prec, emit, quiet=w.stride > 1 and Lin % w.stride != 0)

    def head_bwd(self, w: ConvW, alpha, g_y, y, x_raw, g_raw, g_act, B, Lx, act_fmt=L.FMT_F32):
        L.check(self.lib.b2c_prog_head_bwd(self.h, w.wid, alpha, self._r(g_y), self._r(y), self._r(x_raw), self._r(g_raw),
                                           self._r(g_act), B, Lx, act_fmt), "b2c_prog_head_bwd")

    def head(self, w: ConvW, x, y, B, Lx, x_fmt=L.FMT_F32):
        L.check(self.lib.b2c_prog_head(self.h, w.wid, self._r(x), self._r(y), B, Lx, x_fmt), "b2c_prog_head")

    def layernorm(self, gamma, beta, a, a_mode, out, N, Cc, Tl, chunk, *, sub=None, pe=-1, pe_mode=L.PE_NONE,
                  tanh_post=0, post_scale=1.0, out_fmt=L.FMT_F32):
        L.check(self.lib.b2c_prog_layernorm(self.h, gamma, beta, self._r(a), a_mode, self._r(sub), pe, pe_mode,
                                            tanh_post, post_scale, self._r(out), N, Cc, Tl, chunk, out_fmt),
                "b2c_prog_layernorm")

    def layernorm_masked(self, gamma, beta, a, row_mask, out, N, Cc, Tl, chunk, *, pe=-1, pe_mode=L.PE_NONE,
                         out_fmt=L.FMT_F32):
        L.check(self.lib.b2c_prog_layernorm_masked(self.h, gamma, beta, self._r(a), self._r(row_mask), pe, pe_mode,
                                                   self._r(out), N, Cc, Tl, chunk, out_fmt), "b2c_prog_layernorm_masked")

    def attention_full(self, q, kv, out, B, T, heads, dh):
        L.check(self.lib.b2c_prog_attention_full(self.h, self._r(q), self._r(kv), self._r(out), B, T, heads, dh),
                "b2c_prog_attention_full")

    def select_rows(self, row_mask, a, b, out, N, Cc):
        L.check(self.lib.b2c_prog_select_rows(self.h, self._r(row_mask), self._r(a), self._r(b), self._r(out), N, Cc),
                "b2c_prog_select_rows")

    def ema_update(self, x, idx, emb, counts, N, D, K, decay, one_minus_decay):
        L.check(self.lib.b2c_prog_ema_update(self.h, self._r(x), self._r(idx), self._r(emb), self._r(counts), N, D, K,
                                             decay, one_minus_decay), "b2c_prog_ema_update")

    def attention(self, q, q_mode, kv, out, B, Tl, chunk, heads, dh):
        L.check(self.lib.b2c_prog_attention(self.h, self._r(q), q_mode, self._r(kv), self._r(out), B, Tl, chunk,
                                            heads, dh), "b2c_prog_attention")

    def rvq(self, books_wid, books_use, x, qsum, idx, N, row_mode, B, Tl, chunk, D, prec=L.PREC_F32):
        scratch = self.arena.alloc(int(self.lib.b2c_rvq_scratch_bytes(N, D)))
        L.check(self.lib.b2c_prog_rvq(self.h, books_wid, books_use, self._r(x), self._r(qsum), self._r(idx),
                                      self._r(scratch), N, row_mode, B, Tl, chunk, prec), "b2c_prog_rvq")
        self.arena.free(scratch)

    def rvq_lookup(self, books_wid, books_use, idx, qsum, N, row_mode, B, Tl, chunk):
        L.check(self.lib.b2c_prog_rvq_lookup(self.h, books_wid, books_use, self._r(idx), self._r(qsum), N, row_mode, B,
                                             Tl, chunk), "b2c_prog_rvq_lookup")

    def nearest(self, x, emb, scratch, idx, N, D, K, prec=L.PREC_F32):
        L.check(self.lib.b2c_prog_nearest(self.h, self._r(x), self._r(emb), self._r(scratch), self._r(idx), N, D, K,
                                          prec), "b2c_prog_nearest")

    def dac_rvq(self, wid, n_q, z, zq, codes, B, Tl):
        L.check(self.lib.b2c_prog_dac_rvq(self.h, wid, n_q, self._r(z), self._r(zq), self._r(codes), B, Tl),
                "b2c_prog_dac_rvq")

    def scatter_heads(self, src, dst, B, Tl, chunk, Cc):
        L.check(self.lib.b2c_prog_scatter_heads(self.h, self._r(src), self._r(dst), B, Tl, chunk, Cc),
                "b2c_prog_scatter_heads")

    def transpose(self, src, dst, B, R, Cc):
        L.check(self.lib.b2c_prog_transpose(self.h, self._r(src), self._r(dst), B, R, Cc), "b2c_prog_transpose")

    def widen(self, src, dst, n):
        L.check(self.lib.b2c_prog_i32_to_i64(self.h, self._r(src), self._r(dst), n), "b2c_prog_i32_to_i64")


# ---------------------------------------------------------------------------------------------
# packed sub-graphs + emitters
# ---------------------------------------------------------------------------------------------
@dataclass
class PackedRU:
    a1: int
    c7: ConvW
    a2: int
    c1: ConvW


def _pack_ru(eng: Engine, ru) -> PackedRU:
    s1, c7, s2, c1 = ru.block[0], ru.block[1], ru.block[2], ru.block[3]
    return PackedRU(eng.pack_vec(s1.alpha), eng.pack_wnconv(c7), eng.pack_vec(s2.alpha), eng.pack_wnconv(c1))


@dataclass
class PackedEncoder:
    stem: ConvW
    blocks: list      # [(ru0, ru1, ru2, alpha_down, down ConvW)]
    a_final: int
    head: ConvW
    strides: list

    @staticmethod
    def pack(eng: Engine, enc) -> "PackedEncoder":
        layers = list(enc.block)
        stem = eng.pack_wnconv(layers[0])
        blocks, strides = [], []
        for eb in layers[1:-2]:
            b = eb.block
            blocks.append((_pack_ru(eng, b[0]), _pack_ru(eng, b[1]), _pack_ru(eng, b[2]), eng.pack_vec(b[3].alpha),
                           eng.pack_wnconv(b[4])))
            strides.append(b[4].stride)
        return PackedEncoder(stem, blocks, eng.pack_vec(layers[-2].alpha), eng.pack_wnconv(layers[-1]), strides)

    def out_len(self, T: int) -> int:
        Lx = T
        for (_, _, _, _, dn) in self.blocks:
            Lx = (Lx + 2 * dn.padding - (dn.k - 1) - 1) // dn.stride + 1
        hd = self.head
        return (Lx + 2 * hd.padding - (hd.k - 1) - 1) // hd.stride + 1


def _emit_ru(em: Emitter, ru: PackedRU, x_raw, x_act, B, Lx, next_alpha, need_raw, prec):
    """ResidualUnit: y = x + conv1(snake(conv7(snake(x)))).  x_act = snake1(x_raw) already exists.
    Returns (y_raw or None, y_act = snake_next(y))."""
    C_ = ru.c7.cout
    f = L.FMT_OF_PREC[prec]
    if prec != L.PREC_F32 and em.lib.b2c_ru_tc_eligible(em.eng.ctx, ru.c7.wid, ru.c1.wid, prec) == 1:
        # one fused tcgen05 launch: h = snake2(conv7(x_act)) never leaves shared memory
        y_raw = em.new(B * Lx * C_) if need_raw else None
        y_act = em.new(B * Lx * C_)
        L.check(em.lib.b2c_prog_ru(em.h, ru.c7.wid, ru.a2, ru.c1.wid, em._r(x_act), em._r(x_raw), em._r(y_raw),
                                   em._r(y_act), next_alpha, B, Lx, ru.c7.dilation, prec, f), "b2c_prog_ru")
        em.drop(x_act, x_raw)
        return y_raw, y_act
    h_act = em.new(B * Lx * C_)
    em.conv(ru.c7, x_act, B, Lx, out_act=h_act, alpha=ru.a2, prec=prec, x_fmt=f, act_fmt=f)
    em.drop(x_act)
    y_raw = em.new(B * Lx * C_) if need_raw else None
    y_act = em.new(B * Lx * C_)
    em.conv(ru.c1, h_act, B, Lx, res=x_raw, out_raw=y_raw, out_act=y_act, alpha=next_alpha, prec=prec, x_fmt=f,
            act_fmt=f)
    em.drop(h_act, x_raw)
    return y_raw, y_act


def emit_encoder(em: Emitter, pe: PackedEncoder, x, B, T, prec):
    """x: [B, T] -> returns (z buffer [B, Tl, C], Tl)."""
    c0 = pe.stem.cout
    Lx = T
    x_raw = em.new(B * Lx * c0)
    x_act = em.new(B * Lx * c0)
    f = L.FMT_OF_PREC[prec]
    em.stem(pe.stem, x, x_raw, x_act, pe.blocks[0][0].a1, B, Lx, act_fmt=f)
    for bi, (r0, r1, r2, a_dn, dn) in enumerate(pe.blocks):
        x_raw, x_act = _emit_ru(em, r0, x_raw, x_act, B, Lx, r1.a1, True, prec)
        x_raw, x_act = _emit_ru(em, r1, x_raw, x_act, B, Lx, r2.a1, True, prec)
        x_raw, x_act = _emit_ru(em, r2, x_raw, x_act, B, Lx, a_dn, False, prec)
        Lo = (Lx + 2 * dn.padding - (dn.k - 1) - 1) // dn.stride + 1
        last = bi == len(pe.blocks) - 1
        n_raw = None if last else em.new(B * Lo * dn.cout)
        n_act = em.new(B * Lo * dn.cout)
        em.conv(dn, x_act, B, Lx, out_raw=n_raw, out_act=n_act,
                alpha=pe.a_final if last else pe.blocks[bi + 1][0].a1, prec=prec, x_fmt=f, act_fmt=f)
        em.drop(x_act)
        x_raw, x_act, Lx = n_raw, n_act, Lo
    hd = pe.head
    Lo = (Lx + 2 * hd.padding - (hd.k - 1) - 1) // hd.stride + 1
    z = em.new(B * Lo * hd.cout)
    em.conv(hd, x_act, B, Lx, out_raw=z, prec=prec, x_fmt=f)
    em.drop(x_act)
    return z, Lo


@dataclass
class PackedDecoder:
    stem: ConvW
    blocks: list   # [(alpha_up, up ConvW, ru0, ru1, ru2)]
    a_final: int
    head: ConvW

    @staticmethod
    def pack(eng: Engine, dec) -> "PackedDecoder":
        layers = list(dec.model)
        stem = eng.pack_wnconv(layers[0])
        blocks = []
        for db in layers[1:-3]:
            b = db.block
            blocks.append((eng.pack_vec(b[0].alpha), eng.pack_wnconv(b[1]), _pack_ru(eng, b[2]), _pack_ru(eng, b[3]),
                           _pack_ru(eng, b[4])))
        return PackedDecoder(stem, blocks, eng.pack_vec(layers[-3].alpha), eng.pack_wnconv(layers[-2]))

    def out_len(self, Tl: int) -> int:
        Lx = Tl + 2 * self.stem.padding - (self.stem.k - 1)
        for (_, up, *_r) in self.blocks:
            Lx = (Lx - 1) * up.stride - 2 * up.padding + up.k
        return Lx + 2 * self.head.padding - (self.head.k - 1)


def emit_decoder(em: Emitter, pd: PackedDecoder, z, y, B, Tl, prec, free_input=True):
    """z: [B, Tl, C] buffer -> y: [B, Lout] buffer."""
    Lx = Tl
    f = L.FMT_OF_PREC[prec]
    x_act = em.new(B * Lx * pd.stem.cout)
    em.conv(pd.stem, z, B, Lx, out_act=x_act, alpha=pd.blocks[0][0], prec=prec, x_fmt=L.FMT_F32, act_fmt=f)
    if free_input:
        em.drop(z)
    for bi, (a_up, up, r0, r1, r2) in enumerate(pd.blocks):
        Lo = (Lx - 1) * up.stride - 2 * up.padding + up.k
        x_raw = em.new(B * Lo * up.cout)
        n_act = em.new(B * Lo * up.cout)
        em.convT(up, x_act, B, Lx, out_raw=x_raw, out_act=n_act, alpha=r0.a1, prec=prec, x_fmt=f, act_fmt=f)
        em.drop(x_act)
        x_act, Lx = n_act, Lo
        nxt = pd.blocks[bi + 1][0] if bi + 1 < len(pd.blocks) else pd.a_final
        x_raw, x_act = _emit_ru(em, r0, x_raw, x_act, B, Lx, r1.a1, True, prec)
        x_raw, x_act = _emit_ru(em, r1, x_raw, x_act, B, Lx, r2.a1, True, prec)
        x_raw, x_act = _emit_ru(em, r2, x_raw, x_act, B, Lx, nxt, False, prec)
    em.head(pd.head, x_act, y, B, Lx, x_fmt=f)
    em.drop(x_act)
    return Lx


# ---------------------------------------------------------------------------------------------
# decoder with gradients: forward that keeps every snake's input, and the backward-data pass
# ---------------------------------------------------------------------------------------------
def _folded_weight(m) -> torch.Tensor:
    """w = g * v / ||v|| of a weight-normed conv (norm over all dims but 0, in double), fp32 CPU."""
    v = m.weight_v.detach().to("cpu", torch.float64)
    g = m.weight_g.detach().to("cpu", torch.float64).reshape(-1, 1, 1)
    return (v * (g / v.flatten(1).norm(dim=1).view(-1, 1, 1))).to(torch.float32)


def _pack_conv_bwd(eng: Engine, m) -> ConvW:
    """Backward-data form of a conv as a bias-free Conv1d.  Conv1d [co, ci, k] (stride 1): channels transposed, taps
    flipped, padding dil*(k-1) - p.  ConvTranspose1d [ci, co, k]: d/dx is the strided Conv1d whose weight [cout'=ci,
    cin'=co, k] is the same tensor, with the forward stride and padding."""
    w = _folded_weight(m)
    if getattr(m, "transposed", False):
        cw = eng.pack_plain(w.contiguous())
        cw.stride, cw.padding = m.stride, m.padding
        return cw
    if m.stride != 1:
        raise L.B2CError("backward-data of a strided Conv1d is not needed by the decoder and not built")
    cw = eng.pack_plain(w.permute(1, 0, 2).flip(2).contiguous())
    cw.dilation, cw.padding = m.dilation, m.dilation * (m.kernel_size - 1) - m.padding
    return cw


@dataclass
class PackedDecoderBwd:
    stem: ConvW
    blocks: list   # [(up backward ConvW, [(c7 backward, c1 backward)] * 3)]

    @staticmethod
    def pack(eng: Engine, dec) -> "PackedDecoderBwd":
        layers = list(dec.model)
        blocks = []
        for db in layers[1:-3]:
            b = db.block
            blocks.append((_pack_conv_bwd(eng, b[1]),
                           [(_pack_conv_bwd(eng, ru.block[1]), _pack_conv_bwd(eng, ru.block[3])) for ru in (b[2], b[3], b[4])]))
        return PackedDecoderBwd(_pack_conv_bwd(eng, layers[0]), blocks)


class SavedLayout:
    """Where the forward leaves the inputs of every snake (fp32, channel-last) inside the buffer the autograd node
    owns (external slot `slot`): x0 = stem output, per block u = up-conv output, h[r] = k=7 conv output and y[r] = unit
    output of the three residual units.  Forward and backward programs are emitted against the same layout."""

    def __init__(self, pd: PackedDecoder, B: int, Tl: int, slot: int):
        self.slot, self.top = slot, 0
        self.lens = [Tl + 2 * pd.stem.padding - (pd.stem.k - 1)]
        self.x0 = self._new(B * self.lens[0] * pd.stem.cout)
        self.u, self.h, self.y = [], [], []
        for (_, up, *_r) in pd.blocks:
            Lo = (self.lens[-1] - 1) * up.stride - 2 * up.padding + up.k
            self.lens.append(Lo)
            n = B * Lo * up.cout
            self.u.append(self._new(n))
            self.h.append([self._new(n) for _ in range(3)])
            self.y.append([self._new(n) for _ in range(3)])
        self.nbytes = self.top

    def _new(self, nfloats):
        off = self.top
        self.top += (4 * nfloats + ALIGN - 1) // ALIGN * ALIGN
        return ("ext", self.slot, off)


def emit_decoder_train(em: Emitter, pd: PackedDecoder, z, y, B, Tl, prec, lay: SavedLayout):
    """emit_decoder with every residual unit as two launches whose pre-activations go to `lay` (the fused unit keeps
    h in shared memory; the backward needs it)."""
    f = L.FMT_OF_PREC[prec]
    Lx = Tl
    x_act = em.new(B * Lx * pd.stem.cout)
    em.conv(pd.stem, z, B, Lx, out_raw=lay.x0, out_act=x_act, alpha=pd.blocks[0][0], prec=prec, x_fmt=L.FMT_F32, act_fmt=f)
    for bi, (a_up, up, r0, r1, r2) in enumerate(pd.blocks):
        Lo = lay.lens[bi + 1]
        n = B * Lo * up.cout
        n_act = em.new(n)
        em.convT(up, x_act, B, Lx, out_raw=lay.u[bi], out_act=n_act, alpha=r0.a1, prec=prec, x_fmt=f, act_fmt=f)
        em.drop(x_act)
        x_raw, x_act, Lx = lay.u[bi], n_act, Lo
        nxt = pd.blocks[bi + 1][0] if bi + 1 < len(pd.blocks) else pd.a_final
        for ri, ru in enumerate((r0, r1, r2)):
            h_act = em.new(n)
            em.conv(ru.c7, x_act, B, Lx, out_raw=lay.h[bi][ri], out_act=h_act, alpha=ru.a2, prec=prec, x_fmt=f, act_fmt=f)
            em.drop(x_act)
            y_act = em.new(n)
            em.conv(ru.c1, h_act, B, Lx, res=x_raw, out_raw=lay.y[bi][ri], out_act=y_act,
                    alpha=(r1.a1, r2.a1, nxt)[ri], prec=prec, x_fmt=f, act_fmt=f)
            em.drop(h_act)
            x_raw, x_act = lay.y[bi][ri], y_act
    em.head(pd.head, x_act, y, B, Lx, x_fmt=f)
    em.drop(x_act)


def emit_decoder_bwd(em: Emitter, pd: PackedDecoder, pb: PackedDecoderBwd, lay: SavedLayout, g_y, y, g_z, B, Tl, prec):
    """dL/dz of the decoder: g_y, y [B, Lout] -> g_z [B, C, Tl] (the reference's layout).  Reverse walk of
    emit_decoder_train; per forward "snake -> conv" one b2c_prog_conv_dsnake launch, the skip connection of a residual
    unit is its `res` operand."""
    f = L.FMT_OF_PREC[prec]
    nb = len(pd.blocks)
    Lx, C = lay.lens[nb], pd.head.cin
    g_raw, g_act = em.new(B * Lx * C), em.new(B * Lx * C)
    em.head_bwd(pd.head, pd.a_final, g_y, y, lay.y[nb - 1][2], g_raw, g_act, B, Lx, act_fmt=f)
    for bi in reversed(range(nb)):
        a_up, up, r0, r1, r2 = pd.blocks[bi]
        up_b, ru_b = pb.blocks[bi]
        Lx, C = lay.lens[bi + 1], up.cout
        n = B * Lx * C
        for ri in (2, 1, 0):
            ru = (r0, r1, r2)[ri]
            c7b, c1b = ru_b[ri]
            x_in = lay.u[bi] if ri == 0 else lay.y[bi][ri - 1]
            gh = em.new(n)
            em.conv_dsnake(c1b, g_act, lay.h[bi][ri], ru.a2, B, Lx, out_act=gh, prec=prec, x_fmt=f, act_fmt=f)
            em.drop(g_act)
            n_raw, n_act = em.new(n), em.new(n)
            em.conv_dsnake(c7b, gh, x_in, ru.a1, B, Lx, res=g_raw, out_raw=n_raw, out_act=n_act, prec=prec, x_fmt=f,
                           act_fmt=f)
            em.drop(gh, g_raw)
            g_raw, g_act = n_raw, n_act
        Lp, Cp = lay.lens[bi], up.cin
        pre = lay.x0 if bi == 0 else lay.y[bi - 1][2]
        n_raw = em.new(B * Lp * Cp) if bi > 0 else None
        n_act = em.new(B * Lp * Cp)
        em.conv_dsnake(up_b, g_act, pre, a_up, B, Lx, out_raw=n_raw, out_act=n_act, prec=prec, x_fmt=f, act_fmt=f)
        em.drop(g_act, g_raw)
        g_raw, g_act = n_raw, n_act
    gz = em.new(B * Tl * pd.stem.cin)
    em.conv(pb.stem, g_act, B, lay.lens[0], out_raw=gz, prec=prec, x_fmt=f)
    em.drop(g_act)
    em.transpose(gz, g_z, B, Tl, pd.stem.cin)
    em.drop(gz)


@dataclass
class PackedPredictor:
    """CrossPredictor + TokenNorm + scale + proj_down/up + ResidualVQEMA of ProposedEval."""
    pe: int
    lnq_g: int
    lnq_b: int
    lnkv_g: int
    lnkv_b: int
    wq: ConvW
    wkv: ConvW
    wo: ConvW
    lnf_g: int
    lnf_b: int
    w1: ConvW
    w2: ConvW
    heads: int
    dh: int
    c: int
    # codec-level
    tn_g: int = -1
    tn_b: int = -1
    scale: float = 0.08
    down: ConvW = None
    up: ConvW = None
    books: int = -1
    n_books: int = 0
    code_dim: int = 0

    @staticmethod
    def pack_predictor(eng: Engine, pr, max_chunk=64) -> "PackedPredictor":
        c = pr.q_proj.weight.shape[0]
        pe = eng.pack_vec(pr.pos.pe[:max_chunk].contiguous())
        wkv = torch.cat([pr.k_proj.weight.detach(), pr.v_proj.weight.detach()], dim=0)
        return PackedPredictor(
            pe=pe,
            lnq_g=eng.pack_vec(pr.ln_q.weight), lnq_b=eng.pack_vec(pr.ln_q.bias),
            lnkv_g=eng.pack_vec(pr.ln_kv.weight), lnkv_b=eng.pack_vec(pr.ln_kv.bias),
            wq=eng.pack_plain(pr.q_proj.weight), wkv=eng.pack_plain(wkv), wo=eng.pack_plain(pr.out.weight),
            lnf_g=eng.pack_vec(pr.ffn[0].weight), lnf_b=eng.pack_vec(pr.ffn[0].bias),
            w1=eng.pack_plain(pr.ffn[1].weight, pr.ffn[1].bias), w2=eng.pack_plain(pr.ffn[3].weight, pr.ffn[3].bias),
            heads=pr.h, dh=pr.dh, c=c)


def emit_predict_rows(em: Emitter, pp: PackedPredictor, qn, ctx, N, Tl, chunk, qn_is_table, prec):
    """Everything after attention for N rows: y = out(ctx) + qn; z_pred = y + ffn(y).
    qn: LayerNorm-ed queries ([chunk, C] table when qn_is_table else [N, C]).  Frees ctx."""
    c = pp.c
    f = L.FMT_OF_PREC[prec]
    y1 = em.new(N * c)
    em.conv(pp.wo, ctx, 1, N, res=qn, out_raw=y1, res_mode=1 if qn_is_table else 0, Tl=Tl, chunk=chunk, prec=prec)
    em.drop(ctx)
    h = em.new(N * c)
    em.layernorm(pp.lnf_g, pp.lnf_b, y1, L.ROWS_DENSE, h, N, c, Tl, chunk, out_fmt=f)
    f1 = em.new(N * pp.w1.cout)
    em.conv(pp.w1, h, 1, N, out_act=f1, act=L.ACT_GELU, prec=prec, x_fmt=f, act_fmt=f)
    em.drop(h)
    z_pred = em.new(N * c)
    em.conv(pp.w2, f1, 1, N, res=y1, out_raw=z_pred, prec=prec, x_fmt=f)
    em.drop(f1, y1)
    return z_pred


def emit_predict_full(em: Emitter, pp: PackedPredictor, zq, row_mask, za, B, T, pe_wid, prec):
    """CrossPredictor.forward over ALL T tokens of every frame in one pass (the packet-loss-concealment use,
    PLC/PLC1_eval.py:400-415): q = ln_q(pos(zq * ~mask)), kv = ln_kv(pos(za)), full multi-head attention, out + q,
    + ffn.  zq, za: [B*T, C] channel-last fp32; row_mask: [B*T] bytes or None.  Returns z_pred [B*T, C]."""
    c, N = pp.c, B * T
    f = L.FMT_OF_PREC[prec]
    qn = em.new(N * c)
    if row_mask is None:
        em.layernorm(pp.lnq_g, pp.lnq_b, zq, L.ROWS_DENSE, qn, N, c, T, T, pe=pe_wid, pe_mode=L.PE_CHUNK_POS)
    else:
        em.layernorm_masked(pp.lnq_g, pp.lnq_b, zq, row_mask, qn, N, c, T, T, pe=pe_wid, pe_mode=L.PE_CHUNK_POS)
    kvn = em.new(N * c)
    em.layernorm(pp.lnkv_g, pp.lnkv_b, za, L.ROWS_DENSE, kvn, N, c, T, T, pe=pe_wid, pe_mode=L.PE_CHUNK_POS, out_fmt=f)
    q, kv = em.new(N * c), em.new(N * 2 * c)
    em.conv(pp.wq, qn, 1, N, out_raw=q, prec=prec)
    em.conv(pp.wkv, kvn, 1, N, out_raw=kv, prec=prec, x_fmt=f)
    em.drop(kvn)
    ctx = em.new(N * c)
    em.attention_full(q, kv, ctx, B, T, pp.heads, pp.dh)
    em.drop(q, kv)
    z_pred = emit_predict_rows(em, pp, qn, ctx, N, T, T, False, prec)
    em.drop(qn)
    return z_pred


def emit_residual_code(em: Emitter, pp: PackedPredictor, zt, zt_mode, z_pred, N, B, Tl, chunk, books_use, idx,
                       row_mode, z_hat, prec):
    """r = zt - z_pred; rD = proj_down(scale*tanh(LN(r))); qD = RVQ(rD); z_hat = proj_up(qD) + z_pred."""
    c = pp.c
    f = L.FMT_OF_PREC[prec]
    rn = em.new(N * c)
    em.layernorm(pp.tn_g, pp.tn_b, zt, zt_mode, rn, N, c, Tl, chunk, sub=z_pred, tanh_post=1, post_scale=pp.scale,
                 out_fmt=f)
    rd = em.new(N * pp.code_dim)
    em.conv(pp.down, rn, 1, N, out_raw=rd, prec=prec, x_fmt=f)
    em.drop(rn)
    qd = em.new(N * pp.code_dim)
    em.rvq(pp.books, books_use, rd, qd, idx, N, row_mode, B, Tl, chunk, pp.code_dim, prec)
    em.drop(rd)
    em.conv(pp.up, qd, 1, N, res=z_pred, out_raw=z_hat, prec=prec)
    em.drop(qd)


def emit_residual_decode(em: Emitter, pp: PackedPredictor, z_pred, N, B, Tl, chunk, books_use, idx, row_mode, z_hat,
                         prec):
    """Receiver side of emit_residual_code: qD = sum of the indexed code vectors; z_hat = proj_up(qD) + z_pred."""
    qd = em.new(N * pp.code_dim)
    em.rvq_lookup(pp.books, books_use, idx, qd, N, row_mode, B, Tl, chunk)
    em.conv(pp.up, qd, 1, N, res=z_pred, out_raw=z_hat, prec=prec)
    em.drop(qd)


def emit_latent_coder(em: Emitter, pp: PackedPredictor, qa, zt, z_run, idx, B, Tl, chunk, books_use, prec):
    """The reference's 5-chunk AR loop (Evaluation/dac_vcpwq_proposed6_latency.py:462-477) as two
    dependent passes: pass 1 = all tokens with a zero query input, pass 2 = the first
    token of chunks 1.. with the previous chunk's last reconstructed latent as query input.
    ``zt is None`` emits the RECEIVER: ``idx`` is then an input and the latents are rebuilt from it (pass 1 places
    every token's code vectors on its prediction, pass 2 redoes the chunk heads with their true query input)."""
    c = pp.c
    N = B * Tl
    # K/V for every token
    f = L.FMT_OF_PREC[prec]
    kvn = em.new(N * c)
    em.layernorm(pp.lnkv_g, pp.lnkv_b, qa, L.ROWS_DENSE, kvn, N, c, Tl, chunk, pe=pp.pe, pe_mode=L.PE_CHUNK_POS,
                 out_fmt=f)
    kv = em.new(N * 2 * c)
    em.conv(pp.wkv, kvn, 1, N, out_raw=kv, prec=prec, x_fmt=f)
    em.drop(kvn)
    # pass 1: queries are LN_q(pe[pos]) -- a [chunk, C] table
    qn_tab = em.new(chunk * c)
    em.layernorm(pp.lnq_g, pp.lnq_b, None, L.ROWS_ZERO, qn_tab, chunk, c, Tl, chunk, pe=pp.pe, pe_mode=L.PE_ROW_N)
    q_tab = em.new(chunk * c)
    em.conv(pp.wq, qn_tab, 1, chunk, out_raw=q_tab, prec=prec)
    ctx = em.new(N * c)
    em.attention(q_tab, 0, kv, ctx, B, Tl, chunk, pp.heads, pp.dh)
    em.drop(q_tab)
    z_pred = emit_predict_rows(em, pp, qn_tab, ctx, N, Tl, chunk, True, prec)
    em.drop(qn_tab)
    if zt is None:      # receiver: idx is an input
        emit_residual_decode(em, pp, z_pred, N, B, Tl, chunk, books_use, idx, L.ROWS_DENSE, z_run, prec)
    else:
        emit_residual_code(em, pp, zt, L.ROWS_DENSE, z_pred, N, B, Tl, chunk, books_use, idx, L.ROWS_DENSE, z_run, prec)
    em.drop(z_pred)
    # pass 2: chunk heads
    nfix = (Tl + chunk - 1) // chunk - 1
    if nfix > 0:
        N2 = B * nfix
        qn2 = em.new(N2 * c)
        em.layernorm(pp.lnq_g, pp.lnq_b, z_run, L.ROWS_HEAD_PREV, qn2, N2, c, Tl, chunk, pe=pp.pe,
                     pe_mode=L.PE_ROW0)
        q2 = em.new(N2 * c)
        em.conv(pp.wq, qn2, 1, N2, out_raw=q2, prec=prec)
        ctx2 = em.new(N2 * c)
        em.attention(q2, 1, kv, ctx2, B, Tl, chunk, pp.heads, pp.dh)
        em.drop(q2)
        z_pred2 = emit_predict_rows(em, pp, qn2, ctx2, N2, Tl, chunk, False, prec)
        em.drop(qn2)
        z_hat2 = em.new(N2 * c)
        if zt is None:
            emit_residual_decode(em, pp, z_pred2, N2, B, Tl, chunk, books_use, idx, L.ROWS_HEAD, z_hat2, prec)
        else:
            emit_residual_code(em, pp, zt, L.ROWS_HEAD, z_pred2, N2, B, Tl, chunk, books_use, idx, L.ROWS_HEAD, z_hat2,
                               prec)
        em.drop(z_pred2)
        em.scatter_heads(z_hat2, z_run, B, Tl, chunk, c)
        em.drop(z_hat2)
    em.drop(kv)
