"""autograd.Functions that drive the sm_100a kernels through the C ABI (ekaid_b200.lib).

Four stages make up the hot path; each has a hand-written forward and backward:

  QuestionFn   word embedding gather -> GRU (20 steps) -> tanh-MLP attention with the batch-axis softmax (Q4)
  LinearFn     y = x W^T + b                              (ROI projection `img`, modules.py:195-196)
  RelationFn   one relation encoder step, in closed form (Q1-Q3, Q5-Q7, Q9, Q10, Q13):
               self_feat GEMM (question half broadcast per sample) -> ONE GEMM for [query | key | Z_h] ->
               fused edge kernels -> X + relu(2 out)
  FusionFn     graph combine + difference -> gated fusion -> embed/att -> attention pooling (modules.py:233-308)

Tensors that feed GEMMs are kept in the operand type T of the precision mode: bf16 (tcgen05 path) or fp32
(SIMT parity path).  Residual stream, softmax, reductions and all gradients of parameters are fp32.
"""
from __future__ import annotations

import contextlib
import ctypes
import os
from typing import Optional

import torch

from . import lib
from .lib import Epilogue, call, ptr

ACT_NONE, ACT_RELU, ACT_TANH, ACT_SIGMOID = 0, 1, 2, 3

# tests only: when set to a list, the ReLU active sets chosen by the kernels are appended (relation steps, then the
# embed ReLU) so the oracle can evaluate gradients on the same piecewise-linear branch
DEBUG_SINK = None


# Gradient slots (step.FlatAdam): {parameter data_ptr: (preallocated fp32 gradient view in the flat gradient buffer,
# weak reference to the parameter)}.  When a parameter has a slot, the backward kernels write its gradient straight into
# the slot and hand that tensor to autograd, whose AccumulateGrad adopts it without a copy -- no per-parameter
# "grad += new" kernels and no gradient memset.  That is only valid for the FIRST gradient a parameter receives after
# its .grad was reset: a second contribution (gradient accumulation without zero_grad, or a module applied twice in one
# graph, like the reference's `semantic_relation.forward(bef)` then `.forward(aft)`) gets a fresh tensor, which autograd
# adds onto the first one.  Empty dict = plain autograd behaviour.
GRAD_SLOTS = {}
_SLOTS_WRITTEN = set()      # slots handed out since the owner's last zero_grad (FlatAdam.zero_grad clears it)


def slots_reset() -> None:
    _SLOTS_WRITTEN.clear()


def _dst(key_ptr, shape, dev):
    """Destination of a parameter gradient: its slot on first use after a gradient reset, else a fresh tensor."""
    ent = GRAD_SLOTS.get(key_ptr) if GRAD_SLOTS else None
    if ent is not None:
        slot, ref = ent
        p = ref()
        if (p is not None and p.grad is None and key_ptr not in _SLOTS_WRITTEN and tuple(slot.shape) == tuple(shape)):
            _SLOTS_WRITTEN.add(key_ptr)
            # a FRESH view object: AccumulateGrad only adopts a gradient nobody else references (use_count == 1)
            return slot.view(slot.shape)
    return torch.empty(shape, dtype=torch.float32, device=dev)


_rng_state = {}      # {device: [seed tensor, torch seed it was derived from]}


def _mix64(x: int) -> int:
    """splitmix64 finaliser (host side): decorrelates (torch seed, rank, device) into the 63-bit device seed."""
    x &= 0xFFFFFFFFFFFFFFFF
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return (x ^ (x >> 31)) & 0x7FFFFFFFFFFFFFFF


def rng_state(dev) -> torch.Tensor:
    """Device-resident 64-bit seed of the dropout masks (one per device).  Derived from torch's seed, the data-parallel
    rank and the device index, so ranks draw different masks; re-derived when `torch.manual_seed` has been called since
    (never during a CUDA-graph capture: the tensor's address is what captured kernels read)."""
    dev = torch.device(dev)
    key = str(dev)
    seed_now = torch.initial_seed()
    st = _rng_state.get(key)
    if st is not None and (st[1] == seed_now or torch.cuda.is_current_stream_capturing()):
        return st[0]
    rank = int(os.environ.get("RANK", "0"))
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    val = _mix64(seed_now + 0x9E3779B97F4A7C15 * (rank + 1) + 0xD1B54A32D192ED03 * (idx + 1))
    if st is None:
        st = [torch.tensor([val], dtype=torch.int64, device=dev), seed_now]
        _rng_state[key] = st
    else:
        st[0].copy_(torch.tensor([val], dtype=torch.int64))
        st[1] = seed_now
    return st[0]


def rng_advance(dev) -> None:
    """Draw fresh dropout masks for the next forward (a kernel: stays valid inside a captured CUDA graph).  Called by
    every train-mode forward entry point (ChangeDetector.forward, the stand-alone relation encoders), like nn.Dropout
    draws fresh masks on every call."""
    call("rng_advance", rng_state(dev).data_ptr())


class Drop:
    """Train-mode dropout description handed to the stage functions: probabilities per site group, or off."""

    def __init__(self, dev, on: bool, p_fc=0.2, p_gat=0.2, p_qv=0.2, p_fuse=0.5, p_embed=0.5):
        self.on = on
        self.seed = rng_state(dev).data_ptr() if on else None
        self.p_fc, self.p_gat, self.p_qv, self.p_fuse, self.p_embed = p_fc, p_gat, p_qv, p_fuse, p_embed

    def a(self, site, p):
        """(seed pointer, site, p) triple of one dropout site for the C ABI."""
        return (self.seed, site, float(p) if self.on else 0.0)


def drop_combine(ins, sites, M, C, outf=None, accumulate=0, outT=None):
    """out = sum_k dropout_k(in_k); `sites` = [(seed, site, p)] per input."""
    nin = len(ins)
    i0 = ins[0]
    pad = [None] * (3 - nin)
    ptrs = [t.data_ptr() for t in ins] + pad
    sp = list(sites) + [(None, 0, 0.0)] * (3 - nin)
    seed = next((x[0] for x in sp if x[0] is not None and x[2] > 0), None)
    call("drop_combine", _fmt(i0), _fmt(outT) if outT is not None else 0,
         nin, ptrs[0], ptrs[1], ptrs[2], i0.stride(0), seed, sp[0][1], sp[0][2], sp[1][1], sp[1][2], sp[2][1], sp[2][2],
         M, C, ptr(outf), outf.stride(0) if outf is not None else 0, accumulate, ptr(outT),
         outT.stride(0) if outT is not None else 0)


FWD16 = os.environ.get("EKAID_B200_FWD16", "fp16")     # "bf16": pure-bf16 forward (the round-1 numerics)


def _fmt(t) -> int:
    """format code of a tensor for the C ABI: 0 = fp32, 1 = bf16, 2 = fp16"""
    return {torch.bfloat16: 1, torch.float16: 2}.get(t.dtype, 0)


class PC:
    """precision config: 'bf16' (16-bit tensor-core path) or 'fp32' (SIMT, 1e-4 parity mode).

    On the tensor-core path two 16-bit formats are in use, both with fp32 accumulation in TMEM:
      T  = bf16: gradients (range matters) and the fusion stage;
      TF = fp16: weights and activations of the question path and the relation encoders in the FORWARD pass.  They are
           range-bounded here (GRU state and attention weights in [0, 1], weights O(0.1), node features O(10)) and the
           conversion saturates; the three extra mantissa bits are what brings `input_attended` -- a 40x cancelling
           difference -- inside the 2e-2 bar (scripts/bf16_error_budget.py).  One MMA takes ONE format for both
           operands (a mixed pair raises an illegal-instruction error on the B200), so the backward pass, whose other
           operand is always a bf16 gradient, reads bf16 copies of the tensors it needs (`dual`)."""

    def __init__(self, precision: str):
        if precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32', got %r" % (precision,))
        self.bf16 = precision == "bf16"
        self.T = torch.bfloat16 if self.bf16 else torch.float32
        self.f = 1 if self.bf16 else 0
        self.TF = torch.float16 if (self.bf16 and FWD16 == "fp16") else self.T
        self.ff = 2 if self.TF == torch.float16 else self.f      # format code of forward tensors for the C ABI
        self.dual = self.TF != self.T


# ------------------------------------------------------------------------------------------------
# thin wrappers
# ------------------------------------------------------------------------------------------------
def gemm(A, B, M, N, K, transA=0, transB=0, *, bias=None, addend=None, rowb=None, rowb_div=1, rowb_mod=1,
         rowflag=None, rowb_alt=None, act=ACT_NONE, C=None, Cb=None, splits=0, force_bn=0, drop=None,
         Cb2=None, cb_n1=0, cb2_n0=0, C2=None, c_n1=0, add_n1=0):
    """C[M,N] = op(A) op(B) with the fused epilogue; dtype of A selects tcgen05 (16-bit operands: bf16 or fp16, chosen
    per operand) or SIMT (fp32).  Cb / Cb2: 16-bit outputs (bf16 or fp16 by their dtype): Cb takes columns < cb_n1
    (0 = all), Cb2 columns >= cb2_n0."""
    ep = Epilogue()
    ep.bias = ptr(bias)
    ep.addend = ptr(addend)
    ep.ldadd = addend.stride(0) if addend is not None else 0
    ep.rowb = ptr(rowb)
    ep.ldrowb = rowb.stride(0) if rowb is not None else 0
    ep.rowb_div, ep.rowb_mod = rowb_div, rowb_mod
    ep.rowflag = ptr(rowflag)
    ep.rowb_alt = ptr(rowb_alt)
    ep.act = act
    if drop is not None and drop[2] > 0.0:
        # drop = (seed, site, p[, row pitch of the mask index, column offset])
        ep.drop_seed, ep.drop_site, ep.drop_p = drop[0], drop[1], drop[2]
        ep.drop_n = drop[3] if len(drop) > 3 else N
        ep.drop_off = drop[4] if len(drop) > 4 else 0
    ep.C = ptr(C)
    ep.ldc = C.stride(0) if C is not None else 0
    ep.Cb = ptr(Cb)
    ep.ldcb = Cb.stride(0) if Cb is not None else 0
    ep.cb_fmt = 1 if (Cb is not None and Cb.dtype == torch.float16) else 0
    ep.cb_n1 = cb_n1
    ep.Cb2 = ptr(Cb2)
    ep.ldcb2 = Cb2.stride(0) if Cb2 is not None else 0
    ep.cb2_fmt = 1 if (Cb2 is not None and Cb2.dtype == torch.float16) else 0
    ep.cb2_n0 = cb2_n0
    # fp32 output split by column (C2 takes columns >= c_n1), addend limited to columns < add_n1
    ep.C2 = ptr(C2)
    ep.ldc2 = C2.stride(0) if C2 is not None else 0
    ep.c_n1 = c_n1 if C2 is not None else 0
    ep.add_n1 = add_n1
    assert A.stride(-1) == 1 and B.stride(-1) == 1
    # algorithmic bytes: both operands once, every output once, the epilogue addend once
    nbytes = A.element_size() * (M * K + K * N) + M * N * ((4 if C is not None else 0) + (2 if Cb is not None else 0) +
                                                          (addend.element_size() if addend is not None else 0))
    if Cb2 is not None:
        nbytes += M * (N - cb2_n0) * 2 - (M * (N - cb_n1) * 2 if cb_n1 else 0)
    info = {"flops": 2.0 * M * N * K, "bytes": float(nbytes), "shape": (M, N, K, transA, transB)}
    if A.dtype in (torch.bfloat16, torch.float16):
        assert B.dtype in (torch.bfloat16, torch.float16)
        call("gemm_tc", transA, transB, M, N, K, A.data_ptr(), A.stride(0), 1 if A.dtype == torch.float16 else 0,
             B.data_ptr(), B.stride(0), 1 if B.dtype == torch.float16 else 0, ctypes.addressof(ep), force_bn, splits,
             info=info)
    else:
        assert A.dtype == torch.float32 and B.dtype == torch.float32 and Cb is None and Cb2 is None
        if _split3_ok(A, B, M, N, K, transA, transB):
            # fp32 parity path on the tensor cores: each operand as bf16 (hi, lo) planes, three products accumulated in
            # the same fp32 TMEM accumulator = ONE tcgen05 GEMM over a 3x longer contraction axis (split3_bf16_kernel)
            A3 = _split3(A, K if transA else M, M if transA else K, 0, transA)
            B3 = _split3(B, K if transB else N, N if transB else K, 1, transB)
            info3 = dict(info, split3=True)
            # no automatic split-K here: its partial sums meet in fp32 atomics whose order varies from run to run, and the
            # fp32 path is the reproducible one (same bits every run, like the SIMT kernel it replaces)
            call("gemm_tc", transA, transB, M, N, 3 * K, A3.data_ptr(), A3.stride(0), 0, B3.data_ptr(), B3.stride(0), 0,
                 ctypes.addressof(ep), force_bn, splits if splits > 0 else 1, info=info3)
            return
        call("gemm_f32", transA, transB, M, N, K, A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0),
             ctypes.addressof(ep), info=info)


# one GEMM for both halves of d[v | q] in the relation backward (EKAID_B200_FUSE_DVQ=0: two launches, for A/B measurements)
FUSE_DVQ = os.environ.get("EKAID_B200_FUSE_DVQ", "1") != "0"

# fp32 GEMMs: "split" (default) = 3 x bf16 split-precision products on tcgen05 for every product big enough to fill the
# machine, "simt" = the register-tiled fp32 FMA kernel everywhere (the test oracle of the split path)
FP32_GEMM = os.environ.get("EKAID_B200_FP32_GEMM", "split")
SPLIT3_MIN_MACS = 1 << 27


def _split3_ok(A, B, M, N, K, transA, transB) -> bool:
    if FP32_GEMM != "split" or M * N * K < SPLIT3_MIN_MACS:
        return False
    # TMA: 16-byte row pitches and base addresses of the bf16 planes; float4 reads of the fp32 source
    a_cols, b_cols = (M if transA else K), (N if transB else K)
    return (a_cols % 8 == 0 and b_cols % 8 == 0 and A.stride(0) % 4 == 0 and B.stride(0) % 4 == 0
            and A.data_ptr() % 16 == 0 and B.data_ptr() % 16 == 0)


def _split3(X, rows, cols, pattern, along_rows):
    """fp32 [rows, cols] view -> bf16 planes: [rows, 3*cols] (contraction axis = columns) or [3*rows, cols] (= rows)."""
    out = torch.empty((3 * rows, cols) if along_rows else (rows, 3 * cols), dtype=torch.bfloat16, device=X.device)
    call("split3_bf16", X.data_ptr(), X.stride(0), out.data_ptr(), out.stride(0), rows, cols, pattern, 1 if along_rows else 0)
    return out


def gemm_T(pc: PC, A, B, M, N, K, transA=0, transB=0, *, want_f32=False, **kw):
    """GEMM whose result is needed in the operand type T (and optionally also in fp32)."""
    dev = A.device
    if pc.bf16:
        Cb = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
        C = torch.empty(M, N, dtype=torch.float32, device=dev) if want_f32 else None
        gemm(A, B, M, N, K, transA, transB, C=C, Cb=Cb, **kw)
        return Cb, C
    C = torch.empty(M, N, dtype=torch.float32, device=dev)
    gemm(A, B, M, N, K, transA, transB, C=C, **kw)
    return C, C


def gemm_f32out(A, B, M, N, K, transA=0, transB=0, out=None, **kw):
    C = out if out is not None else torch.empty(M, N, dtype=torch.float32, device=A.device)
    gemm(A, B, M, N, K, transA, transB, C=C, **kw)
    return C


def _to16(w: torch.Tensor, dtype) -> torch.Tensor:
    w = w.detach()
    if w.dim() == 1:
        w = w.view(1, -1)
    if not w.is_contiguous():
        w = w.contiguous()
    if dtype == torch.float32:
        return w
    out = torch.empty(w.shape, dtype=dtype, device=w.device)
    call("cast_f32_f16" if dtype == torch.float16 else "cast_f32_bf16", w.data_ptr(), w.stride(0), out.data_ptr(),
         out.stride(0), w.shape[0], w.shape[1])
    return out


def to_T(pc: PC, w: torch.Tensor) -> torch.Tensor:
    """fp32 2-D tensor -> gradient-side operand type (own cast kernel for bf16; identity for fp32)."""
    return _to16(w, pc.T)


def to_TF(pc: PC, w: torch.Tensor) -> torch.Tensor:
    """fp32 2-D tensor -> forward operand type (fp16 on the tensor-core path unless EKAID_B200_FWD16=bf16)."""
    return _to16(w, pc.TF)


def bcopy(pairs) -> None:
    """[(src fp16 2-D view, dst bf16 view)]: the backward pass's bf16 copies of forward tensors, ONE launch."""
    for lo in range(0, len(pairs), 16):
        chunk = pairs[lo:lo + 16]
        n = len(chunk)
        srcs = [(a if a.dim() == 2 else a.view(1, -1)) for a, _ in chunk]
        dsts = [(b if b.dim() == 2 else b.view(1, -1)) for _, b in chunk]
        for a, b in zip(srcs, dsts):
            assert a.dtype == torch.float16 and b.dtype == torch.bfloat16 and a.shape == b.shape
        ps = (ctypes.c_void_p * n)(*[t.data_ptr() for t in srcs])
        pd = (ctypes.c_void_p * n)(*[t.data_ptr() for t in dsts])
        ls = (ctypes.c_int64 * n)(*[t.stride(0) for t in srcs])
        ld = (ctypes.c_int64 * n)(*[t.stride(0) for t in dsts])
        rows = (ctypes.c_int64 * n)(*[t.shape[0] for t in srcs])
        cols = (ctypes.c_int32 * n)(*[t.shape[1] for t in srcs])
        mode = (ctypes.c_int32 * n)(*([4] * n))
        call("cast_many", n, ctypes.addressof(ps), ctypes.addressof(ls), ctypes.addressof(pd), ctypes.addressof(ld),
             ctypes.addressof(rows), ctypes.addressof(cols), ctypes.addressof(mode))


def cast_many(pc: PC, pairs) -> None:
    """[(src fp32 2-D view, dst view)] -> ONE launch: each dst receives a cast to ITS dtype (bf16 / fp16) or, for an fp32
    dst, a copy.  Views may have any row pitch."""
    pairs = [(s_.detach(), d) for s_, d in pairs]
    for lo in range(0, len(pairs), 16):
        chunk = pairs[lo:lo + 16]
        n = len(chunk)
        srcs = [(s_ if s_.dim() == 2 else s_.view(1, -1)) for s_, _ in chunk]
        dsts = [(d if d.dim() == 2 else d.view(1, -1)) for _, d in chunk]
        for a, b in zip(srcs, dsts):
            assert a.dtype == torch.float32 and a.stride(-1) == 1 and b.stride(-1) == 1 and a.shape == b.shape
        ps = (ctypes.c_void_p * n)(*[t.data_ptr() for t in srcs])
        pd = (ctypes.c_void_p * n)(*[t.data_ptr() for t in dsts])
        ls = (ctypes.c_int64 * n)(*[t.stride(0) for t in srcs])
        ld = (ctypes.c_int64 * n)(*[t.stride(0) for t in dsts])
        rows = (ctypes.c_int64 * n)(*[t.shape[0] for t in srcs])
        cols = (ctypes.c_int32 * n)(*[t.shape[1] for t in srcs])
        mode = (ctypes.c_int32 * n)(*[{torch.bfloat16: 0, torch.float16: 3}.get(t.dtype, 1) for t in dsts])
        call("cast_many", n, ctypes.addressof(ps), ctypes.addressof(ls), ctypes.addressof(pd), ctypes.addressof(ld),
             ctypes.addressof(rows), ctypes.addressof(cols), ctypes.addressof(mode))


def copy_many_bytes(pairs) -> None:
    """[(src, dst)] contiguous device tensors of equal byte size (multiples of 16) -> one launch of raw copies."""
    for lo in range(0, len(pairs), 16):
        chunk = pairs[lo:lo + 16]
        n = len(chunk)
        nb = [s_.numel() * s_.element_size() for s_, _ in chunk]
        ps = (ctypes.c_void_p * n)(*[s_.data_ptr() for s_, _ in chunk])
        pd = (ctypes.c_void_p * n)(*[d.data_ptr() for _, d in chunk])
        ls = (ctypes.c_int64 * n)(*nb)
        rows = (ctypes.c_int64 * n)(*([1] * n))
        cols = (ctypes.c_int32 * n)(*nb)
        mode = (ctypes.c_int32 * n)(*([2] * n))
        call("cast_many", n, ctypes.addressof(ps), ctypes.addressof(ls), ctypes.addressof(pd), ctypes.addressof(ls),
             ctypes.addressof(rows), ctypes.addressof(cols), ctypes.addressof(mode))


_ws_cache = {}


def _workspace(dev, n):
    key = (dev, torch.cuda.current_stream().cuda_stream)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < n:
        ws = torch.empty(max(n, 64 * 8192), dtype=torch.float32, device=dev)
        _ws_cache[key] = ws
    return ws


_colsum_bufs = {}


def _colsum_ws(dev, N):
    """Workspace of ekaid_colsum per (device, stream): ticket counters (zero between calls) + partial sums."""
    key = (str(dev), torch.cuda.current_stream().cuda_stream)
    need = 1024 + 64 * N
    ws = _colsum_bufs.get(key)
    if ws is None or ws.numel() < need:
        if ws is not None:
            _colsum_bufs.setdefault("retired", []).append(ws)     # a captured CUDA graph may still point at it
        ws = torch.zeros(max(need, 1024 + 64 * 6144), dtype=torch.float32, device=dev)
        _colsum_bufs[key] = ws
    return ws


def colsum_many(jobs, M) -> None:
    """[(src 2-D view [M, N_j], out fp32 [N_j])]: all column sums in ONE launch (deterministic)."""
    n = len(jobs)
    dev = jobs[0][0].device
    nmax = max(s_.shape[1] for s_, _ in jobs)
    key = ("many", str(dev), torch.cuda.current_stream().cuda_stream)
    need = n * (1024 + 64 * nmax)
    ws = _colsum_bufs.get(key)
    if ws is None or ws.numel() < need:
        if ws is not None:
            _colsum_bufs.setdefault("retired", []).append(ws)
        ws = torch.zeros(max(need, 8 * (1024 + 64 * 6144)), dtype=torch.float32, device=dev)
        _colsum_bufs[key] = ws
    isb = (ctypes.c_int32 * n)(*[1 if s_.dtype == torch.bfloat16 else 0 for s_, _ in jobs])
    ps = (ctypes.c_void_p * n)(*[s_.data_ptr() for s_, _ in jobs])
    ld = (ctypes.c_int64 * n)(*[s_.stride(0) for s_, _ in jobs])
    ns = (ctypes.c_int32 * n)(*[s_.shape[1] for s_, _ in jobs])
    po = (ctypes.c_void_p * n)(*[o.data_ptr() for _, o in jobs])
    call("colsum_many", n, ctypes.addressof(isb), ctypes.addressof(ps), ctypes.addressof(ld), M, ctypes.addressof(ns),
         ctypes.addressof(po), ws.data_ptr())


def colsum(src, M, N, rowscale=None, out=None):
    """out[n] = sum_m rowscale[m] * src[m, n]  (fp32 result)."""
    if out is None:
        out = torch.empty(N, dtype=torch.float32, device=src.device)
    ws = _colsum_ws(src.device, N)
    call("colsum", 1 if src.dtype == torch.bfloat16 else 0, src.data_ptr(), src.stride(0) if src.dim() == 2 else 1,
         M, N, ptr(rowscale), out.data_ptr(), ws.data_ptr())
    return out


def _f32c(t):
    t = t.detach()
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


# ------------------------------------------------------------------------------------------------
# legacy weight_norm(dim=None)
# ------------------------------------------------------------------------------------------------
class WNormFn(torch.autograd.Function):
    """w = v * g / ||v||_F  (models/fc.py:33-34; torch's `_weight_norm` with dim=None), two launches each way instead
    of the ~10 element-wise / reduction ATen kernels of the composite implementation."""

    @staticmethod
    def forward(ctx, v, g):
        lib.require_device()
        vc, gc = _f32c(v), _f32c(g).reshape(1)
        w = torch.empty_like(vc)
        norm = torch.empty(1, dtype=torch.float32, device=vc.device)
        ws = torch.empty(128, dtype=torch.float32, device=vc.device)
        call("wn_fwd", vc.data_ptr(), gc.data_ptr(), vc.numel(), w.data_ptr(), norm.data_ptr(), ws.data_ptr())
        ctx.saved = (vc, gc, norm)
        ctx.gshape = g.shape
        ctx.keys = (v.data_ptr(), g.data_ptr())
        return w

    @staticmethod
    def backward(ctx, dw):
        vc, gc, norm = ctx.saved
        dwc = _f32c(dw)
        dv = _dst(ctx.keys[0], vc.shape, vc.device)
        dg = _dst(ctx.keys[1], ctx.gshape, vc.device)
        ws = torch.empty(128, dtype=torch.float32, device=vc.device)
        call("wn_bwd", dwc.data_ptr(), vc.data_ptr(), gc.data_ptr(), norm.data_ptr(), vc.numel(), dv.data_ptr(),
             dg.data_ptr(), ws.data_ptr())
        return dv, dg


def _ptr_array(tensors):
    return (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


def wn_many_compute(vg):
    """The kernels of WNormManyFn.forward without autograd: (v0, g0, v1, g1, ...) -> (vs, gs, [w0, w1, ...], norms)."""
    lib.require_device()
    vs = [_f32c(t) for t in vg[0::2]]
    gs = [_f32c(t).reshape(1) for t in vg[1::2]]
    cnt = len(vs)
    dev = vs[0].device
    ws_out = [torch.empty_like(v) for v in vs]
    norms = torch.empty(cnt, dtype=torch.float32, device=dev)
    work = torch.empty(cnt * 128, dtype=torch.float32, device=dev)
    ns = (ctypes.c_int64 * cnt)(*[v.numel() for v in vs])
    pv, pg, pw = _ptr_array(vs), _ptr_array(gs), _ptr_array(ws_out)      # host arrays, read during the call
    call("wn_fwd_many", cnt, ctypes.addressof(pv), ctypes.addressof(pg), ctypes.addressof(ns), ctypes.addressof(pw),
         norms.data_ptr(), work.data_ptr())
    return vs, gs, ws_out, norms


class WNormManyFn(torch.autograd.Function):
    """WNormFn for up to 16 (v, g) pairs at once: two launches forward and two backward for all of them.
    Arguments: pre, v0, g0, v1, g1, ...; returns (w0, w1, ...).

    `pre` = wn_many_compute(...) of the same arguments, or None.  ChangeDetector computes the effective weights early
    (they feed the activation-independent preparation that overlaps the question path) but creates THIS autograd node
    only right before the encoder runs: autograd executes backward nodes newest-first, so the weight-norm backward of an
    encoder then runs right after that encoder's backward -- its v / g gradients are complete early and can be
    exchanged between ranks while the rest of backward is still running."""

    @staticmethod
    def forward(ctx, pre, *vg):
        vs, gs, ws_out, norms = pre if pre is not None else wn_many_compute(vg)
        ctx.saved = (vs, gs, norms)
        ctx.gshapes = [t.shape for t in vg[1::2]]
        ctx.keys = [(v.data_ptr(), g.data_ptr()) for v, g in zip(vg[0::2], vg[1::2])]
        return tuple(w.view(w.shape) for w in ws_out) if pre is not None else tuple(ws_out)

    @staticmethod
    def backward(ctx, *dws):
        vs, gs, norms = ctx.saved
        cnt = len(vs)
        dev = vs[0].device
        dwc = [_f32c(d) if d is not None else torch.zeros_like(v) for d, v in zip(dws, vs)]
        dvs = [_dst(k[0], v.shape, dev) for k, v in zip(ctx.keys, vs)]
        dgs = [_dst(k[1], sh, dev) for k, sh in zip(ctx.keys, ctx.gshapes)]
        work = torch.empty(cnt * 128, dtype=torch.float32, device=dev)
        ns = (ctypes.c_int64 * cnt)(*[v.numel() for v in vs])
        pdw, pv, pg, pdv, pdg = (_ptr_array(x) for x in (dwc, vs, gs, dvs, dgs))
        call("wn_bwd_many", cnt, ctypes.addressof(pdw), ctypes.addressof(pv), ctypes.addressof(pg), norms.data_ptr(),
             ctypes.addressof(ns), ctypes.addressof(pdv), ctypes.addressof(pdg), work.data_ptr())
        out = [None]
        for dv, dg in zip(dvs, dgs):
            out += [dv, dg]
        return tuple(out)


class SmallLinearFn(torch.autograd.Function):
    """nn.Linear with a handful of outputs (fc1, modules.py:312): one warp per output element instead of a GEMM
    launch.  The reference never uses `pred` in the loss (Q11); the backward is plain torch for whoever does."""

    @staticmethod
    def forward(ctx, x, W, b):
        lib.require_device()
        xc, Wc = _f32c(x), _f32c(W)
        y = torch.empty(xc.shape[0], Wc.shape[0], dtype=torch.float32, device=xc.device)
        call("small_linear", xc.data_ptr(), xc.stride(0), xc.shape[0], xc.shape[1], Wc.data_ptr(),
             _f32c(b).data_ptr() if b is not None else None, Wc.shape[0], y.data_ptr())
        ctx.save_for_backward(xc, Wc)
        ctx.has_bias = b is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        xc, Wc = ctx.saved_tensors
        return dy @ Wc, dy.t() @ xc, dy.sum(0) if ctx.has_bias else None


class WeightedSumsFn(torch.autograd.Function):
    """loss = sum_k coef_k <a_k, w_k>  (w_k None: sum of a_k) in one launch; the gradient of a_k is coef_k * w_k."""

    @staticmethod
    def forward(ctx, coefs, weights, *tensors):
        lib.require_device()
        ts = [_f32c(t) for t in tensors]
        dev = ts[0].device
        out = torch.empty(1, dtype=torch.float32, device=dev)
        cnt = len(ts)
        pa = _ptr_array(ts)
        pw = (ctypes.c_void_p * cnt)(*[w.data_ptr() if w is not None else None for w in weights])
        ns = (ctypes.c_int64 * cnt)(*[t.numel() for t in ts])
        cf = (ctypes.c_float * cnt)(*[float(c) for c in coefs])
        key = ("wsum", str(dev), torch.cuda.current_stream().cuda_stream)
        if key not in _colsum_bufs:
            _colsum_bufs[key] = torch.zeros(128, dtype=torch.float32, device=dev)
        call("weighted_sums", cnt, ctypes.addressof(pa), ctypes.addressof(pw), ctypes.addressof(ns),
             ctypes.addressof(cf), out.data_ptr(), _colsum_bufs[key].data_ptr())
        ctx.coefs, ctx.weights = coefs, weights
        ctx.shapes = [t.shape for t in tensors]
        return out.view(())

    @staticmethod
    def backward(ctx, g):
        # the gradient of every term is a constant times g: all of them in ONE launch (the composite form is ~25 tiny
        # ATen kernels, 50 us of launch latency in the captured step)
        cnt = len(ctx.shapes)
        dev = g.device
        gc = _f32c(g).reshape(1)
        outs = [torch.empty(sh, dtype=torch.float32, device=dev) for sh in ctx.shapes]
        po = _ptr_array(outs)
        pw = (ctypes.c_void_p * cnt)(*[w.data_ptr() if w is not None else None for w in ctx.weights])
        ns = (ctypes.c_int64 * cnt)(*[o.numel() for o in outs])
        cf = (ctypes.c_float * cnt)(*[float(c) for c in ctx.coefs])
        call("weighted_sums_bwd", cnt, ctypes.addressof(po), ctypes.addressof(pw), ctypes.addressof(ns),
             ctypes.addressof(cf), gc.data_ptr())
        return (None, None, *outs)


def cast_into(pc: PC, src: torch.Tensor, dst: torch.Tensor) -> None:
    """fp32 2-D `src` (any row pitch) -> operand-type view `dst` (any row pitch), own kernels."""
    src = src.detach()
    if src.dim() == 1:
        src = src.view(1, -1)
    assert src.stride(-1) == 1 and dst.stride(-1) == 1 and src.shape == dst.shape
    call("cast_f32_bf16" if pc.bf16 else "copy_f32", src.data_ptr(), src.stride(0), dst.data_ptr(), dst.stride(0),
         src.shape[0], src.shape[1])


# ------------------------------------------------------------------------------------------------
# y = x W^T + b
# ------------------------------------------------------------------------------------------------
class LinearFn(torch.autograd.Function):
    """nn.Linear (modules.py:93,195-196).  `x2` (optional) is stacked under `x` along the row axis so the main and
    the reference image of every pair go through ONE GEMM.  Returns (y fp32 [rows, N], y in operand type or None)."""

    @staticmethod
    def forward(ctx, pc: PC, x, x2, W, b):
        lib.require_device()
        dev = x.device
        K = x.shape[-1]
        xa = _f32c(x).view(-1, K)
        Ma = xa.shape[0]
        Mb = 0
        if x2 is not None:
            xb = _f32c(x2).view(-1, K)
            Mb = xb.shape[0]
        M = Ma + Mb
        N = W.shape[0]
        xT = torch.empty(M, K, dtype=pc.T, device=dev)
        fn = "cast_f32_bf16" if pc.bf16 else "copy_f32"
        call(fn, xa.data_ptr(), K, xT.data_ptr(), K, Ma, K)
        if Mb:
            call(fn, xb.data_ptr(), K, xT[Ma:].data_ptr(), K, Mb, K)
        WT = to_T(pc, W)
        bb = _f32c(b) if b is not None else None
        y = torch.empty(M, N, dtype=torch.float32, device=dev)
        yT = torch.empty(M, N, dtype=pc.TF, device=dev) if pc.bf16 else None      # operand copy for the next stage's GEMM
        gemm(xT, WT, M, N, K, bias=bb, C=y, Cb=yT)
        ctx.pc, ctx.xT, ctx.WT, ctx.has_b = pc, xT, WT, b is not None
        ctx.keys = (W.data_ptr(), b.data_ptr() if b is not None else 0)
        ctx.need = (x.requires_grad, x2 is not None and x2.requires_grad)
        ctx.shapes = (x.shape, x2.shape if x2 is not None else None, Ma)
        if yT is not None:
            ctx.mark_non_differentiable(yT)
        return y, yT

    @staticmethod
    def backward(ctx, dy, _dyT=None):
        pc = ctx.pc
        dy2 = _f32c(dy).view(-1, dy.shape[-1])
        M, N = dy2.shape
        K = ctx.xT.shape[1]
        dyT = to_T(pc, dy2)
        dW = _dst(ctx.keys[0], (N, K), dy2.device)
        db = _dst(ctx.keys[1], (N,), dy2.device) if ctx.has_b else None
        dx = dx2 = None
        need_dx = ctx.need[0] or ctx.need[1]
        dfull = torch.empty(M, K, dtype=torch.float32, device=dy2.device) if need_dx else None
        fk = Fork(dy2.device, 1)
        with fk.branch(0) if need_dx else contextlib.nullcontext():
            gemm_f32out(dyT, ctx.xT, N, K, M, transA=1, transB=1, out=dW)
            if ctx.has_b:
                colsum(dy2, M, N, out=db)
        if need_dx:
            gemm_f32out(dyT, ctx.WT, M, K, N, transB=1, out=dfull)
        fk.join()
        if need_dx:
            sa, sb, Ma = ctx.shapes
            if ctx.need[0]:
                dx = dfull[:Ma].view(sa)
            if ctx.need[1]:
                dx2 = dfull[Ma:].view(sb)
        return None, dx, dx2, dW, db


# ------------------------------------------------------------------------------------------------
# question path
# ------------------------------------------------------------------------------------------------
# step.GraphFusionStep sets this: called (no arguments) right before the BPTT recurrence is launched, i.e. when the
# question path starts its longest serial stretch -- the moment to put independent work next to it (the optimizer update
# of the image-path parameters, whose gradients are all enqueued by then).
BPTT_HOOK = None
CHECK_TOKEN_RANGE = os.environ.get("EKAID_B200_CHECK_TOKENS", "1") != "0"
GRU_SEQ = os.environ.get("EKAID_B200_GRU_SEQ", "1") != "0"     # 0: per-step GEMM + cell kernels on the bf16 path too
GRU_SEQ_MAX_BATCH = 64
_barrier_bufs = {}


def _barrier_ws(dev):
    """Two zeroed words of device memory per (device, stream) for the persistent GRU kernels' grid barrier; every launch
    leaves them zero again (gru_seq.cu: grid_barrier_retire), so launches ordered on one stream share them."""
    key = (str(dev), torch.cuda.current_stream().cuda_stream)
    if key not in _barrier_bufs:
        _barrier_bufs[key] = torch.zeros(4, dtype=torch.int32, device=dev)
    return _barrier_bufs[key]


_branch = {}


def _branch_streams(dev):
    """Two extra streams per device for independent gradient groups of the question path."""
    if os.environ.get("EKAID_B200_QBRANCH", "1") == "0":
        cur = torch.cuda.current_stream(dev)
        return cur, cur
    key = str(dev)
    if key not in _branch:
        _branch[key] = (torch.cuda.Stream(dev), torch.cuda.Stream(dev))
    return _branch[key]


def _gru_seq_ok(pc, dev, B, H):
    """The one-launch recurrence covers the bf16 path when all H/16 CTAs are co-resident (see gru_seq.cu).  It works
    in passes of 64 batch rows, so for large batches the per-step tcgen05 GEMMs win."""
    if not (GRU_SEQ and pc.bf16 and H % 1024 == 0 and B <= GRU_SEQ_MAX_BATCH):
        return False
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    return H // 8 <= sms and B * 64 + 192 * 1024 <= 227 * 1024


_fork_pool = {}
FORK_ENABLED = os.environ.get("EKAID_B200_FORK", "1") != "0"      # bench.py serialises the branches for per-kernel timing


class Fork:
    """Run independent kernels of one stage on parallel streams.  The GEMM kernels are persistent (one CTA per SM): when
    two independent ones are in flight, the second one's CTAs move onto SMs as the first one's CTAs retire, so the
    tail of one launch (last-tile epilogue, idle SMs of a partial wave) overlaps the head of the next.  Inside a captured
    CUDA graph this becomes parallel branches.  All tensors are allocated on the calling stream BEFORE the fork and the
    caller joins before using any result, which keeps the caching allocator's stream bookkeeping trivial.

        f = Fork(dev, 2); with f.branch(0): ...; with f.branch(1): ...; (work on the calling stream); f.join()
    """

    def __init__(self, dev, n: int, pool: str = "stage"):
        key = (str(dev), pool)          # separate pools: work queued on one pool never sits behind another pool's work
        pool = _fork_pool.setdefault(key, [])
        while len(pool) < n:
            pool.append(torch.cuda.Stream(dev))
        self.cur = torch.cuda.current_stream(dev)
        self.on = FORK_ENABLED
        self.streams = pool[:n] if self.on else [self.cur] * n
        if self.on:
            for st in self.streams:
                st.wait_stream(self.cur)

    def branch(self, i: int, resync: bool = False):
        """Context of branch i; resync=True makes the branch wait for what the calling stream has enqueued since the fork."""
        if resync and self.on:
            self.streams[i].wait_stream(self.cur)
        return torch.cuda.stream(self.streams[i])

    def join(self):
        if self.on:
            for st in self.streams:
                self.cur.wait_stream(st)


class GRUFn(torch.autograd.Function):
    """nn.GRU(batch_first, 1 layer, unidirectional, h0 = 0) on x [B,L,in] -> all outputs [B,L,H]
    (QuestionEmbedding.forward_all stand-alone, language_model.py:106-115).  Same kernels as QuestionFn's middle part."""

    @staticmethod
    def forward(ctx, pc: PC, x, Wih, Whh, bih, bhh):
        lib.require_device()
        dev = x.device
        B, L, I = x.shape
        H = Whh.shape[1]
        # time-major rows t*B + b, in the operand type
        E = to_T(pc, _f32c(x).transpose(0, 1).reshape(L * B, I))
        WihT, WhhT = to_T(pc, Wih), to_T(pc, Whh)
        bihc, bhhc = _f32c(bih), _f32c(bhh)
        gi = gemm_f32out(E, WihT, L * B, 3 * H, I, bias=bihc)
        Hs = torch.empty(L * B, H, dtype=torch.float32, device=dev)
        HsT = torch.zeros((L + 1) * B, H, dtype=pc.T, device=dev)
        gates = torch.empty(L, B, 4 * H, dtype=torch.float32, device=dev)
        if _gru_seq_ok(pc, dev, B, H):
            call("gru_seq_fwd", gi.data_ptr(), WhhT.data_ptr(), bhhc.data_ptr(), B, H, L, Hs.data_ptr(), HsT.data_ptr(),
                 gates.data_ptr(), _barrier_ws(dev).data_ptr(), 0, None)
        else:
            gh = torch.empty(B, 3 * H, dtype=torch.float32, device=dev)
            call("copy_f32", bhhc.data_ptr(), 0, gh.data_ptr(), 3 * H, B, 3 * H)
            for t in range(L):
                gemm(HsT[t * B:(t + 1) * B], WhhT, B, 3 * H, H, addend=gh, C=gh)
                hprev = Hs[(t - 1) * B:t * B] if t > 0 else None
                call("gru_cell_fwd", pc.f, gi[t * B:(t + 1) * B].data_ptr(), gh.data_ptr(), ptr(hprev), B, H,
                     Hs[t * B:(t + 1) * B].data_ptr(), HsT[(t + 1) * B:(t + 2) * B].data_ptr(), gates[t].data_ptr(),
                     bhhc.data_ptr())
        ctx.pc, ctx.dims = pc, (B, L, I, H)
        ctx.saved = (E, WihT, WhhT, Hs, HsT, gates)
        return Hs.view(L, B, H).transpose(0, 1).contiguous()

    @staticmethod
    def backward(ctx, dout):
        pc = ctx.pc
        B, L, I, H = ctx.dims
        E, WihT, WhhT, Hs, HsT, gates = ctx.saved
        dev = Hs.device
        dHs = _f32c(dout).transpose(0, 1).reshape(L * B, H).contiguous()
        dgi = torch.empty(L * B, 3 * H, dtype=torch.float32, device=dev)
        dgh = torch.empty(L * B, 3 * H, dtype=torch.float32, device=dev)
        dgiT = torch.empty(L * B, 3 * H, dtype=pc.T, device=dev) if pc.bf16 else dgi
        dghT = torch.empty(L * B, 3 * H, dtype=pc.T, device=dev) if pc.bf16 else dgh
        if _gru_seq_ok(pc, dev, B, H):
            call("gru_seq_bwd", dHs.data_ptr(), gates.data_ptr(), Hs.data_ptr(), WhhT.data_ptr(), B, H, L, dgi.data_ptr(),
                 dgh.data_ptr(), dgiT.data_ptr(), dghT.data_ptr(), _barrier_ws(dev).data_ptr())
        else:
            carry = torch.empty(B, H, dtype=torch.float32, device=dev)
            for t in range(L - 1, -1, -1):
                sl = slice(t * B, (t + 1) * B)
                if t < L - 1:
                    call("add_inplace", dHs[sl].data_ptr(), carry.data_ptr(), B * H)
                hprev = Hs[(t - 1) * B:t * B] if t > 0 else None
                call("gru_cell_bwd", pc.f, dHs[sl].data_ptr(), gates[t].data_ptr(), ptr(hprev), B, H, dgi[sl].data_ptr(),
                     dgh[sl].data_ptr(), dgiT[sl].data_ptr() if pc.bf16 else None,
                     dghT[sl].data_ptr() if pc.bf16 else None, carry.data_ptr())
                if t > 0:
                    gemm(dghT[sl], WhhT, B, H, 3 * H, transB=1, addend=carry, C=carry)
        dWih = gemm_f32out(dgiT, E, 3 * H, I, L * B, transA=1, transB=1)
        dbih = colsum(dgi, L * B, 3 * H)
        dWhh = gemm_f32out(dghT, HsT[:L * B], 3 * H, H, L * B, transA=1, transB=1)
        dbhh = colsum(dgh, L * B, 3 * H)
        dx = gemm_f32out(dgiT, WihT, L * B, I, 3 * H, transB=1)
        return None, dx.view(L, B, I).transpose(0, 1).contiguous(), dWih, dWhh, dbih, dbhh


class QuestionFn(torch.autograd.Function):
    """w_emb -> q_emb.forward_all -> q_att   (modules.py:200-206; language_model.py:48-53,106-115,127-156).

    Forward operands are in pc.TF (fp16 on the tensor-core path: embeddings, GRU state and tanh outputs are bounded);
    the backward pairs bf16 gradients with bf16 copies of the weights / saved activations (suffix B below)."""

    @staticmethod
    def forward(ctx, pc: PC, drop, question, emb, emb2, Wih, Whh, bih, bhh, W1, b1, w2, b2, padding_idx=-1):
        lib.require_device()
        dev = emb.device
        B, L = question.shape
        ed = emb.shape[1]
        H = Whh.shape[1]
        q = question.detach().to(torch.int64).contiguous()
        if CHECK_TOKEN_RANGE and not torch.cuda.is_current_stream_capturing():
            # nn.Embedding raises on ids outside the table; the gather kernel would read out of bounds (one host sync,
            # skipped inside a CUDA-graph capture; EKAID_B200_CHECK_TOKENS=0 turns it off)
            if bool(((q < 0) | (q >= emb.shape[0])).any()):
                raise IndexError("question holds token ids outside [0, %d)" % emb.shape[0])
        need_bwd = any(ctx.needs_input_grad)
        dual = pc.dual and need_bwd
        embc, emb2c = _f32c(emb), _f32c(emb2)
        E = torch.empty(L * B, 2 * ed, dtype=pc.TF, device=dev)
        call("embed_gather", pc.ff, q.data_ptr(), embc.data_ptr(), emb2c.data_ptr(), B, L, ed, E.data_ptr())
        srcs = [_f32c(Wih), _f32c(Whh), _f32c(W1)]
        if pc.bf16:
            WihT, WhhT, W1T = (torch.empty(t.shape, dtype=pc.TF, device=dev) for t in srcs)
            jobs = list(zip(srcs, (WihT, WhhT, W1T)))
            if dual:
                WihB, WhhB, W1B = (torch.empty(t.shape, dtype=pc.T, device=dev) for t in srcs)
                jobs += list(zip(srcs, (WihB, WhhB, W1B)))
            else:
                WihB, WhhB, W1B = WihT, WhhT, W1T
            # (measured: casting only the recurrence's two copies ahead of it and the rest on a branch stream changes
            # nothing -- 3.2565 vs 3.2536 ms/step, profiles/r02_notes.md -- so all copies go out in one launch)
            cast_many(pc, jobs)
        else:
            WihT, WhhT, W1T = srcs
            WihB, WhhB, W1B = srcs
        bihc, bhhc, b1c, w2c, b2c = _f32c(bih), _f32c(bhh), _f32c(b1), _f32c(w2).view(-1), _f32c(b2).view(-1)
        gi = gemm_f32out(E, WihT, L * B, 3 * H, 2 * ed, bias=bihc)
        Hs = torch.empty(L * B, H, dtype=torch.float32, device=dev)
        # operand-type copy with one leading zero block: HsT[t] = h_{t-1}, so "previous h of every step" is a view
        # (the recurrence writes every other block; only block 0 is cleared)
        HsT = torch.empty((L + 1) * B, H, dtype=pc.TF, device=dev)
        HsT[:B].zero_()
        if dual:
            HsB = torch.empty((L + 1) * B, H, dtype=pc.T, device=dev)
            HsB[:B].zero_()
        else:
            HsB = HsT
        gates = torch.empty(L, B, 4 * H, dtype=torch.float32, device=dev)
        if _gru_seq_ok(pc, dev, B, H):
            # the whole recurrence in one persistent launch (gru_seq.cu)
            call("gru_seq_fwd", gi.data_ptr(), WhhT.data_ptr(), bhhc.data_ptr(), B, H, L, Hs.data_ptr(), HsT.data_ptr(),
                 gates.data_ptr(), _barrier_ws(dev).data_ptr(), 1 if pc.ff == 2 else 0, HsB.data_ptr() if dual else None)
        else:
            # gh accumulator: armed with b_hh (broadcast copy), "gh += h W_hh^T" as a split-K GEMM (M = B is tiny, so
            # the K loop is what can be spread over the SMs), re-armed by the cell kernel
            gh = torch.empty(B, 3 * H, dtype=torch.float32, device=dev)
            call("copy_f32", bhhc.data_ptr(), 0, gh.data_ptr(), 3 * H, B, 3 * H)
            for t in range(L):
                gemm(HsT[t * B:(t + 1) * B], WhhT, B, 3 * H, H, addend=gh, C=gh)
                hprev = Hs[(t - 1) * B:t * B] if t > 0 else None
                call("gru_cell_fwd", pc.ff, gi[t * B:(t + 1) * B].data_ptr(), gh.data_ptr(), ptr(hprev), B, H,
                     Hs[t * B:(t + 1) * B].data_ptr(), HsT[(t + 1) * B:(t + 2) * B].data_ptr(), gates[t].data_ptr(),
                     bhhc.data_ptr())
            if dual:
                bcopy([(HsT[B:], HsB[B:])])
        HsT_cur = HsT[B:]
        if drop is not None and drop.on:          # Dropout(0.2) on the input of W1 (language_model.py:123-124)
            Hd = torch.empty(L * B, H, dtype=pc.TF, device=dev)
            drop_combine([HsT_cur], [drop.a(10, drop.p_fc)], L * B, H, outT=Hd)
            if dual:
                HdB = torch.empty(L * B, H, dtype=pc.T, device=dev)
                bcopy([(Hd, HdB)])
            else:
                HdB = Hd
        else:
            Hd, HdB = HsT_cur, HsB[B:]
        # a1 = tanh(Hd W1^T + b1) stays in fp32 (5 MB): it is no GEMM operand, and its backward factor 1 - a1^2 is
        # ill-conditioned under 16-bit storage wherever the tanh saturates
        a1 = torch.empty(L * B, H, dtype=torch.float32, device=dev)
        gemm(Hd, W1T, L * B, H, H, bias=b1c, act=ACT_TANH, C=a1)
        a1B = a1
        a = torch.empty(L * B, dtype=torch.float32, device=dev)
        call("rowdot", 0, a1.data_ptr(), a1.stride(0), L * B, H, w2c.data_ptr(), b2c.data_ptr(), a.data_ptr())
        S = torch.empty(L * B, dtype=torch.float32, device=dev)
        qv = torch.empty(B, H, dtype=torch.float32, device=dev)
        call("qpool_fwd", a.data_ptr(), Hs.data_ptr(), B, L, H, S.data_ptr(), qv.data_ptr())
        if drop is not None and drop.on:          # self.drop on the pooled vector (language_model.py:155)
            drop_combine([qv], [drop.a(11, drop.p_qv)], B, H, outf=qv)
        ctx.pc, ctx.drop = pc, drop
        ctx.keys = {k: t.data_ptr() for k, t in (("emb", emb), ("Wih", Wih), ("Whh", Whh), ("bih", bih), ("bhh", bhh),
                                                 ("b1", b1), ("b2", b2))}
        ctx.dims = (B, L, ed, H, emb.shape[0])
        ctx.padding_idx = int(padding_idx)
        if need_bwd:
            if dual:
                EB = torch.empty(L * B, 2 * ed, dtype=pc.T, device=dev)
                bcopy([(E, EB)])
            else:
                EB = E
            ctx.saved = (q, EB, WihB, WhhB, W1B, w2c, Hs, HsB, gates, a1B, S, HdB)
        return qv

    @staticmethod
    def backward(ctx, dqv):
        pc = ctx.pc
        B, L, ed, H, V = ctx.dims
        q, E, WihT, WhhT, W1T, w2c, Hs, HsT, gates, a1, S, Hd = ctx.saved      # all 16-bit tensors here are in pc.T
        drop = ctx.drop
        don = drop is not None and drop.on
        dev = Hs.device
        dqv = _f32c(dqv)
        if don:
            dq2 = torch.empty_like(dqv)
            drop_combine([dqv], [drop.a(11, drop.p_qv)], B, H, outf=dq2)
            dqv = dq2
        dS = torch.empty(L * B, dtype=torch.float32, device=dev)
        da = torch.empty(L * B, dtype=torch.float32, device=dev)
        dHs = torch.empty(L * B, H, dtype=torch.float32, device=dev)
        call("qpool_bwd", dqv.data_ptr(), S.data_ptr(), Hs.data_ptr(), B, L, H, dS.data_ptr(), da.data_ptr(),
             dHs.data_ptr())
        dpre = torch.empty(L * B, H, dtype=pc.T, device=dev)
        # (unrounded fp32 copy of dpre for the bias gradient: a column sum that cancels -- sum_b da[l, b] = 0 under the
        # batch-axis softmax -- so bf16 rounding noise of the terms would dominate it)
        dpre32 = torch.empty(L * B, H, dtype=torch.float32, device=dev) if pc.bf16 else None
        call("qatt_tanh_bwd", 3 if pc.bf16 else 0, da.data_ptr(), w2c.data_ptr(), a1.data_ptr(), L * B, H, dpre.data_ptr(),
             ptr(dpre32))
        kk = ctx.keys
        # Everything below that does not feed BPTT (weight / bias gradients of the attention MLP) goes to a branch stream
        # and runs next to the recurrence; the serial chain of this stream is only dHs -> BPTT.  Temporaries are
        # allocated here, on this stream, so the caching allocator's stream bookkeeping stays simple.
        cur = torch.cuda.current_stream(dev)
        br1, br2 = _branch_streams(dev)
        dw2 = torch.empty(H, dtype=torch.float32, device=dev)
        dW1 = torch.empty(H, H, dtype=torch.float32, device=dev)
        db2 = _dst(kk["b2"], (1,), dev)
        db1 = _dst(kk["b1"], (H,), dev)
        br1.wait_stream(cur)
        with torch.cuda.stream(br1):
            colsum(a1, L * B, H, rowscale=da, out=dw2)
            colsum(da.view(-1, 1), L * B, 1, out=db2)
            gemm_f32out(dpre, Hd, H, H, L * B, transA=1, transB=1, out=dW1)
            colsum(dpre32 if dpre32 is not None else dpre, L * B, H, out=db1)
        dw2 = dw2.view(1, H)
        if don:
            # dHs += mask * (dpre W1): dropout mask of W1's input and the accumulation, both in the GEMM epilogue
            gemm(dpre, W1T, L * B, H, H, transB=1, addend=dHs, C=dHs, drop=drop.a(10, drop.p_fc))
        else:
            gemm(dpre, W1T, L * B, H, H, transB=1, addend=dHs, C=dHs)     # dHs += dpre W1
        # BPTT
        dgi = torch.empty(L * B, 3 * H, dtype=torch.float32, device=dev)
        dgh = torch.empty(L * B, 3 * H, dtype=torch.float32, device=dev)
        dgiT = torch.empty(L * B, 3 * H, dtype=pc.T, device=dev) if pc.bf16 else dgi
        dghT = torch.empty(L * B, 3 * H, dtype=pc.T, device=dev) if pc.bf16 else dgh
        if BPTT_HOOK is not None:
            BPTT_HOOK()
        if _gru_seq_ok(pc, dev, B, H):
            call("gru_seq_bwd", dHs.data_ptr(), gates.data_ptr(), Hs.data_ptr(), WhhT.data_ptr(), B, H, L, dgi.data_ptr(),
                 dgh.data_ptr(), dgiT.data_ptr(), dghT.data_ptr(), _barrier_ws(dev).data_ptr())
        else:
            carry = torch.empty(B, H, dtype=torch.float32, device=dev)
            for t in range(L - 1, -1, -1):
                sl = slice(t * B, (t + 1) * B)
                if t < L - 1:
                    call("add_inplace", dHs[sl].data_ptr(), carry.data_ptr(), B * H)
                hprev = Hs[(t - 1) * B:t * B] if t > 0 else None
                # the cell kernel writes dh*z into `carry`; the split-K GEMM then adds dgh W_hh onto it
                call("gru_cell_bwd", pc.f, dHs[sl].data_ptr(), gates[t].data_ptr(), ptr(hprev), B, H, dgi[sl].data_ptr(),
                     dgh[sl].data_ptr(), dgiT[sl].data_ptr() if pc.bf16 else None,
                     dghT[sl].data_ptr() if pc.bf16 else None, carry.data_ptr())
                if t > 0:
                    gemm(dghT[sl], WhhT, B, H, 3 * H, transB=1, addend=carry, C=carry)   # carry = dh*z + dgh W_hh
        # three independent groups of parameter gradients: this stream and the two branches take one each
        dWih = _dst(kk["Wih"], (3 * H, 2 * ed), dev)
        dbih = _dst(kk["bih"], (3 * H,), dev)
        dWhh = _dst(kk["Whh"], (3 * H, H), dev)
        dbhh = _dst(kk["bhh"], (3 * H,), dev)
        demb = _dst(kk["emb"], (V, ed), dev)
        dE = torch.empty(L * B, ed, dtype=torch.float32, device=dev)
        br1.wait_stream(cur)
        br2.wait_stream(cur)
        with torch.cuda.stream(br1):
            gemm_f32out(dghT, HsT[:L * B], 3 * H, H, L * B, transA=1, transB=1, out=dWhh)
            colsum(dgh, L * B, 3 * H, out=dbhh)
        with torch.cuda.stream(br2):
            gemm_f32out(dgiT, E, 3 * H, 2 * ed, L * B, transA=1, transB=1, out=dWih)
            colsum(dgi, L * B, 3 * H, out=dbih)
        # the two-kernel chain (dE, then the table gradient) stays on this stream: it starts without a cross-stream hop,
        # and its second kernel runs next to the other two products instead of after them
        gemm_f32out(dgiT, WihT, L * B, ed, 3 * H, transB=1, out=dE)           # only the trainable table's columns
        call("embed_gather_bwd", q.data_ptr(), dE.data_ptr(), dE.stride(0), B, L, ed, V, demb.data_ptr(), ctx.padding_idx)
        cur.wait_stream(br1)
        cur.wait_stream(br2)
        return None, None, None, demb, None, dWih, dWhh, dbih, dbhh, dW1, db1, dw2, db2, None


# ------------------------------------------------------------------------------------------------
# the edge part alone (stand-alone GraphSelfAttentionLayer / GAttNet forward)
# ------------------------------------------------------------------------------------------------
class EdgeAttentionFn(torch.autograd.Function):
    """out = sum_h softmax(mask(Q_h K_h^T / sqrt(d_h) + gbias_h) + lbias) Z_h + b_out   (graph_att_layer.py:105-178 with
    the out-projection re-associated per Q3).  QKZ32 [G*N, (2+H)*D] fp32 = [query | key | Z_0..Z_{H-1}] (rows j >= Kn of
    the key / Z parts are never read).  cond / lbias [G,N,Kn], gbias [G,N,Kn,H] may be None.  Returns (out, P)."""

    @staticmethod
    def forward(ctx, pc: PC, dims, QKZ32, cond, lbias, gbias, bout):
        lib.require_device()
        G, N, Kn, D, H = dims
        dev = QKZ32.device
        M, W = G * N, (2 + H) * D
        q32 = _f32c(QKZ32).view(M, W)
        QKZ = to_T(pc, q32)
        condc = _f32c(cond).contiguous() if cond is not None else None
        lbc = _f32c(lbias).contiguous() if lbias is not None else None
        gbc = _f32c(gbias).contiguous() if gbias is not None else None
        P = torch.empty(G, N, H, Kn, dtype=torch.float32, device=dev)
        Phl = None
        if pc.bf16 and (H * Kn) % 8 == 0 and N <= 128 and (D // H) % 16 == 0:
            Phl = torch.empty(2, G, N, H * Kn, dtype=torch.bfloat16, device=dev)
        call("edge_softmax_fwd", pc.f, QKZ.data_ptr(), QKZ.stride(0), D, ptr(condc), ptr(lbc), ptr(gbc), G, N, Kn, H,
             P.data_ptr(), ptr(Phl))
        out = torch.empty(M, D, dtype=torch.float32, device=dev)
        call("edge_aggregate_fwd", pc.f, P.data_ptr(), QKZ.data_ptr(), QKZ.stride(0), D, _f32c(bout).data_ptr(), None,
             G, N, Kn, H, out.data_ptr(), None, D, None, None, 0, 0.0, ptr(Phl), None, 0)
        ctx.pc, ctx.dims = pc, dims
        ctx.flags = (lbias is not None, gbias is not None)
        ctx.saved = (QKZ, condc, P, Phl)
        ctx.mark_non_differentiable(P)
        return out, P

    @staticmethod
    def backward(ctx, dout, _dP):
        pc = ctx.pc
        G, N, Kn, D, H = ctx.dims
        QKZ, cond, P, Phl = ctx.saved
        dev = P.device
        M, W = G * N, (2 + H) * D
        dout = _f32c(dout).view(M, D)
        ns = lib.load().ekaid_edge_bwd_slices(pc.f, D, N, Kn, H, 1 if Phl is not None else 0)
        dQKZ = torch.zeros(M, W, dtype=pc.T, device=dev)
        dOut = torch.empty(M, D, dtype=torch.float32, device=dev)
        dPpart = torch.empty(ns, G, N, H, Kn, dtype=torch.float32, device=dev)
        ones = torch.ones(M, D, dtype=torch.uint8, device=dev)          # plain output: no ReLU mask, gradient scale 1
        call("edge_aggregate_bwd", pc.f, dout.data_ptr(), ones.data_ptr(), P.data_ptr(), QKZ.data_ptr(), QKZ.stride(0), D,
             G, N, Kn, H, dQKZ.data_ptr(), dOut.data_ptr(), dPpart.data_ptr(), 1.0, ptr(Phl))
        dbout = colsum(dOut, M, D)
        has_lb, has_gb = ctx.flags
        dlb = torch.empty(H, G, N, Kn, dtype=torch.float32, device=dev) if has_lb else None
        dgb = torch.empty(G, N, Kn, H, dtype=torch.float32, device=dev) if has_gb else None
        call("edge_softmax_bwd", pc.f, P.data_ptr(), dPpart.data_ptr(), ns, QKZ.data_ptr(), QKZ.stride(0), D, ptr(cond),
             G, N, Kn, H, dQKZ.data_ptr(), ptr(dlb), ptr(dgb))
        return None, None, dQKZ.float(), None, (dlb.sum(0) if has_lb else None), dgb, dbout


# ------------------------------------------------------------------------------------------------
# one relation encoder step
# ------------------------------------------------------------------------------------------------
_dim_t_cache = {}


def _dim_t(dev, feat_dim=64, wave_length=1000.0):
    """The wave lengths exactly as utils/mimic_utils.py:195-197 computes them (fp32)."""
    key = (str(dev), feat_dim)
    if key not in _dim_t_cache:
        feat_range = torch.arange(0, feat_dim / 8)
        dim_mat = torch.pow(torch.ones((1,)) * wave_length, (8.0 / feat_dim) * feat_range)
        _dim_t_cache[key] = dim_mat.float().contiguous().to(dev)
    return _dim_t_cache[key]


def _labels_i8(t: torch.Tensor) -> torch.Tensor:
    t = t.detach()
    if t.dtype != torch.int8:
        t = t.to(torch.int8)
    return t.contiguous()


def relation_prepare(pc: PC, drop, site0, kind: str, dims, Wsw, Wq, bq, Wk, bk, Wo2, p0, p1, adj0, adj1, g_split,
                     need_bwd: bool = True):
    """Everything of a relation step that does not depend on the activations or the question vector: operand-type
    copies of the weights ([Wq; Wk; Z-blocks] stacked for the single QKZ GEMM), the adjacency condition / label bias
    (explicit) or the geometry bias (implicit).  ChangeDetector runs this for all encoders while the question path is
    still busy on its own stream; RelationFn.forward falls back to it when no `prep` is handed in."""
    G, B, N, Kn, D, H = dims
    dev = Wsw.device
    don = drop is not None and drop.on
    Wsw32 = _f32c(Wsw)
    dual = pc.dual and need_bwd
    WswT = torch.empty(Wsw32.shape, dtype=pc.TF, device=dev) if pc.bf16 else Wsw32
    # [Wq ; Wk ; Z-blocks] operand: ONE GEMM yields query, key and Z_h = self_feat W_out2[:, hD:(h+1)D]^T (Q3)
    WqkzT = torch.empty((2 + H) * D, D, dtype=pc.TF, device=dev)
    Wo2c = _f32c(Wo2)
    bqkzc = torch.zeros((2 + H) * D, dtype=torch.float32, device=dev)

    def wjobs(Wq_dst, Wsw_dst):
        j = [(_f32c(Wq), Wq_dst[0:D]), (_f32c(Wk), Wq_dst[D:2 * D])]
        j += [(Wo2c[:, h * D:(h + 1) * D], Wq_dst[(2 + h) * D:(3 + h) * D]) for h in range(H)]
        if pc.bf16:
            j.append((Wsw32, Wsw_dst))
        return j

    jobs = wjobs(WqkzT, WswT)
    jobs += [(_f32c(bq).view(1, D), bqkzc[0:D].view(1, D)), (_f32c(bk).view(1, D), bqkzc[D:2 * D].view(1, D))]
    if dual:
        # bf16 copies for the backward's dgrad GEMMs (their other operand is a bf16 gradient), same launch
        WswB = torch.empty(Wsw32.shape, dtype=pc.T, device=dev)
        WqkzB = torch.empty((2 + H) * D, D, dtype=pc.T, device=dev)
        jobs += wjobs(WqkzB, WswB)
    else:
        WswB, WqkzB = WswT, WqkzT
    cast_many(pc, jobs)          # one launch instead of up to sixteen
    cond = lbias = gbias = None
    if kind == "explicit":
        wb = _f32c(p0).view(-1)
        cond = torch.empty(G, N, Kn, dtype=torch.float32, device=dev)
        lbias = torch.empty(G, N, Kn, dtype=torch.float32, device=dev)
        if adj0.dim() == 3:
            # the loader's integer label matrices [*, S, S] (0 = no edge, c + 1 = plane c of process_matrix): consumed
            # as they are, the fp32 one-hot planes are never built
            a0 = _labels_i8(adj0)
            a1 = _labels_i8(adj1) if adj1 is not None else None
            Lb, S = wb.numel(), a0.shape[1]
            if a0.shape[2] != S or S < N or (a1 is not None and a1.shape[1:] != a0.shape[1:]):
                raise ValueError("label matrices must be [*, S, S] with S >= %d, got %s" % (N, tuple(a0.shape)))
            call("adj_labels_fwd", a0.data_ptr(), ptr(a1), g_split, S, wb.data_ptr(), G, N, Kn, Lb, cond.data_ptr(),
                 lbias.data_ptr())
            geo = (a0, a1, Lb, S)
        else:
            a0 = _f32c(adj0)
            a1 = _f32c(adj1) if adj1 is not None else None
            Lb = a0.shape[-1]
            if a0.shape[1] != N or a0.shape[2] != N or (a1 is not None and a1.shape[1:] != a0.shape[1:]):
                raise ValueError("adjacency must be [*, %d, %d, labels], got %s" % (N, N, tuple(a0.shape)))
            call("adj_prep_fwd", a0.data_ptr(), ptr(a1), g_split, wb.data_ptr(), G, N, Kn, Lb, cond.data_ptr(),
                 lbias.data_ptr())
            geo = (a0, a1, Lb, 0)
    else:
        a0 = adj0.detach().to(device=dev, dtype=torch.float64).contiguous()
        a1 = adj1.detach().to(device=dev, dtype=torch.float64).contiguous() if adj1 is not None else None
        Wp, bp = _f32c(p0), _f32c(p1)
        gbias = torch.empty(G, N, Kn, H, dtype=torch.float32, device=dev)
        dgeo = drop.a(site0 + 4, drop.p_fc) if don else (None, 0, 0.0)
        need_bwd = torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in (p0, p1))
        emb_cache = torch.empty(G, N * Kn, 64, dtype=torch.float32, device=dev) if need_bwd else None
        call("geom_bias_fwd", a0.data_ptr(), ptr(a1), g_split, Wp.data_ptr(), bp.data_ptr(),
             _dim_t(dev).data_ptr(), G, N, Kn, H, gbias.data_ptr(), *dgeo, ptr(emb_cache), pc.f)
        geo = (a0, a1, Wp, bp, emb_cache)
    return {"WswT": WswT, "Wsw32": Wsw32, "WqkzT": WqkzT, "WswB": WswB, "WqkzB": WqkzB, "bqkzc": bqkzc, "cond": cond,
            "lbias": lbias, "gbias": gbias, "geo": geo, "dual": dual}


class RelationFn(torch.autograd.Function):
    """X <- X + relu(2 * GAT_dir1(cat(X, q)))  for G stacked images (G = S*B; image g uses question g % B).

    kind 'explicit': adj0/adj1 are one-hot/float adjacency [*, N, N, L]; wb = effective label-bias table [L].
    kind 'implicit': adj0/adj1 are fp64 boxes [*, N, 4] (device); Wp [H,64], bp [H] = effective pair_pos_fc1.
    Images [0, g_split) read adj0, the rest adj1."""

    @staticmethod
    def forward(ctx, pc: PC, drop, site0, kind: str, dims, X, XT, qv, Wsw, bsw, Wq, bq, Wk, bk, Wo2, bout, p0, p1,
                adj0, adj1, g_split, prep=None):
        lib.require_device()
        G, B, N, Kn, D, H = dims
        dev = X.device
        M = G * N
        X = _f32c(X).view(M, D)
        qv = _f32c(qv)
        don = drop is not None and drop.on
        need_bwd = any(ctx.needs_input_grad)
        if prep is None or (need_bwd and pc.dual and not prep["dual"]):
            prep = relation_prepare(pc, drop, site0, kind, dims, Wsw, Wq, bq, Wk, bk, Wo2, p0, p1, adj0, adj1, g_split,
                                    need_bwd=need_bwd)
        dual = pc.dual and need_bwd       # bf16 copies of what the backward's GEMMs / edge kernels read (suffix B)
        WswT, Wsw32, WqkzT, bqkzc = prep["WswT"], prep["Wsw32"], prep["WqkzT"], prep["bqkzc"]
        bswc, boutc = _f32c(bsw), _f32c(bout)
        flags = torch.empty(M, dtype=torch.uint8, device=dev)
        call("row_zero_flags", X.data_ptr(), M, D, flags.data_ptr())
        W = (2 + H) * D
        Dq = qv.shape[1]
        qvT = to_TF(pc, qv)
        qvB = to_T(pc, qv) if dual else qvT
        # Z in the forward format for the aggregation (fp16: 11 significant bits, one MMA per product); Q, K -- and, when
        # a backward follows, Z -- in bf16 for the score kernels and the backward edge kernels
        Z16 = torch.empty(M, H * D, dtype=pc.TF, device=dev) if pc.dual else None
        wq = W if (need_bwd or not pc.dual) else 2 * D        # inference: no bf16 Z at all
        QKZ = torch.empty(M, wq, dtype=pc.T, device=dev)
        SfB = SqB = SkB = XB = None
        if not don:
            if XT is None or XT.dtype != pc.TF:
                XT = to_TF(pc, X)
            # question half of self_weights, once per sample (M = B rows), broadcast per row in the GEMM epilogue
            qpart = gemm_f32out(qvT, WswT[:, D:], B, D, Dq, bias=bswc)
            Sf = torch.empty(M, D, dtype=pc.TF, device=dev)
            if dual:
                SfB = torch.empty(M, D, dtype=pc.T, device=dev)
                XB = torch.empty(M, D, dtype=pc.T, device=dev)
            kw = dict(rowb=qpart, rowb_div=N, rowb_mod=B, rowflag=flags, rowb_alt=bswc)
            if pc.bf16:
                gemm(XT, WswT[:, :D], M, D, D, Cb=Sf, Cb2=SfB, **kw)
                if pc.dual:
                    gemm(Sf, WqkzT, M, W, D, bias=bqkzc, Cb=QKZ, cb_n1=wq if wq < W else 0, Cb2=Z16, cb2_n0=2 * D)
                else:
                    gemm(Sf, WqkzT, M, W, D, bias=bqkzc, Cb=QKZ)
                if dual:
                    bcopy([(XT, XB)])
            else:
                gemm(XT, WswT[:, :D], M, D, D, C=Sf, **kw)
                gemm(Sf, WqkzT, M, W, D, bias=bqkzc, C=QKZ)
            Sq = Sk = Sf
            if not dual:
                SfB, XB = Sf, XT
            SqB = SkB = SfB
        else:
            # train mode: Dropout(0.2) hits the concatenated [v | q] element-wise, so the question half is no longer
            # the same for every node -> K = D + Dq GEMM on the dropped concat; query / key see two more masks
            XT = torch.empty(M, D + Dq, dtype=pc.TF, device=dev)          # the dropped [v | q] operand
            Sf = torch.empty(M, D, dtype=pc.TF, device=dev)
            Sq = torch.empty(M, D, dtype=pc.TF, device=dev)
            Sk = torch.empty(M, D, dtype=pc.TF, device=dev)
            if dual:
                # bf16 copies for the backward's wgrad GEMMs, written by the same kernels that produce the fp16 ones
                SfB = torch.empty(M, D, dtype=pc.T, device=dev)
                SqB = torch.empty(M, D, dtype=pc.T, device=dev)
                SkB = torch.empty(M, D, dtype=pc.T, device=dev)
                XB = torch.empty(M, D + Dq, dtype=pc.T, device=dev)
            call("build_vq", pc.ff, X.data_ptr(), qv.data_ptr(), flags.data_ptr(), M, N, B, D, Dq, XT.data_ptr(),
                 *drop.a(site0 + 1, drop.p_fc), ptr(XB))
            if pc.bf16:
                gemm(XT, WswT, M, D, D + Dq, bias=bswc, Cb=Sf, Cb2=SfB)
            else:
                gemm(XT, WswT, M, D, D + Dq, bias=bswc, C=Sf)
            call("drop_fanout", pc.ff, Sf.data_ptr(), Sf.stride(0), drop.seed, site0 + 2, float(drop.p_fc), site0 + 3,
                 float(drop.p_fc), M, D, Sq.data_ptr(), Sk.data_ptr(), D, ptr(SqB), ptr(SkB))
            # three independent projections: the big Z GEMM on this stream, query / key next to it
            fk = Fork(dev, 2)
            for bi, (src, lo, hi) in enumerate(((Sq, 0, D), (Sk, D, 2 * D), (Sf, 2 * D, W))):
                out = QKZ[:, lo:hi]
                with (fk.branch(bi) if bi < 2 else contextlib.nullcontext()):
                    if not pc.bf16:
                        gemm(src, WqkzT[lo:hi], M, hi - lo, D, bias=bqkzc[lo:hi], C=out)
                    elif bi < 2:
                        gemm(src, WqkzT[lo:hi], M, hi - lo, D, bias=bqkzc[lo:hi], Cb=out)
                    elif pc.dual:
                        gemm(src, WqkzT[lo:hi], M, hi - lo, D, bias=bqkzc[lo:hi], Cb=out if need_bwd else None, Cb2=Z16)
                    else:
                        gemm(src, WqkzT[lo:hi], M, hi - lo, D, bias=bqkzc[lo:hi], Cb=out)
            fk.join()
            if not dual:
                SfB, SqB, SkB, XB = Sf, Sq, Sk, XT
        cond, lbias, gbias = prep["cond"], prep["lbias"], prep["gbias"]
        ctx.geo = prep["geo"]
        P = torch.empty(G, N, H, Kn, dtype=torch.float32, device=dev)
        es = 2 if pc.bf16 else 4
        # bf16 path: P also as 16-bit planes, staged by the aggregation kernels with async 16-byte copies: bf16 hi + lo
        # (backward) and, on the dual-format path, fp16 (forward)
        Phl = None
        if pc.bf16 and (H * Kn) % 8 == 0 and N <= 128 and (D // H) % 16 == 0:
            Phl = torch.empty(3 if pc.dual else 2, G, N, H * Kn, dtype=torch.bfloat16, device=dev)
        use16 = pc.dual and Phl is not None
        if pc.dual and not use16:
            raise lib.EkaidError("the fp16 forward aggregation needs H*K % 8 == 0, N <= 128, head dim % 16 == 0 "
                                 "(set EKAID_B200_FWD16=bf16 for other shapes)")
        call("edge_softmax_fwd", 3 if use16 else pc.f, QKZ.data_ptr(), QKZ.stride(0), D, ptr(cond), ptr(lbias), ptr(gbias),
             G, N, Kn, H, P.data_ptr(), ptr(Phl),
             info={"bytes": G * ((N + Kn) * D * es + N * H * Kn * 4 + N * Kn * 4 * (2 if cond is not None else H))})
        Xn = torch.empty(M, D, dtype=torch.float32, device=dev)
        XnT = torch.empty(M, D, dtype=pc.TF, device=dev) if pc.bf16 else None
        mask = torch.empty(M, D, dtype=torch.uint8, device=dev)
        call("edge_aggregate_fwd", pc.f, P.data_ptr(), QKZ.data_ptr(), QKZ.stride(0), D, boutc.data_ptr(), X.data_ptr(),
             G, N, Kn, H, Xn.data_ptr(), ptr(XnT), D, mask.data_ptr(),
             *(drop.a(site0 + 5, drop.p_gat) if don else (None, 0, 0.0)), ptr(Phl),
             Z16.data_ptr() if use16 else None, Z16.stride(0) if use16 else 0,
             info={"bytes": G * (N * H * Kn * (2 if use16 else 4) + Kn * H * D * es + N * D * (4 + 4 + 1 + (2 if pc.bf16 else 0)))})
        ctx.pc, ctx.kind, ctx.dims, ctx.g_split = pc, kind, dims, g_split
        ctx.drop, ctx.site0 = drop, site0
        ctx.keys = {k: (t.data_ptr() if t is not None else 0) for k, t in (("bsw", bsw), ("bq", bq), ("bk", bk),
                                                                          ("Wo2", Wo2), ("bout", bout), ("p1", p1))}
        if need_bwd:
            ctx.saved = (XB, qvB, prep["WswB"], Wsw32, prep["WqkzB"], flags, SfB, SqB, SkB, QKZ, cond, P, mask, Phl)
        if DEBUG_SINK is not None:
            DEBUG_SINK.append(mask.bool().cpu())
        if XnT is not None:
            ctx.mark_non_differentiable(P, XnT)
        else:
            ctx.mark_non_differentiable(P)
        return Xn, XnT, P

    @staticmethod
    def backward(ctx, dXn, _dXnT, _dP):
        pc, kind = ctx.pc, ctx.kind
        G, B, N, Kn, D, H = ctx.dims
        XT, qvT, WswT, Wsw32, WqkzT, flags, Sf, Sq, Sk, QKZ, cond, P, mask, Phl = ctx.saved
        drop, site0 = ctx.drop, ctx.site0
        don = drop is not None and drop.on
        dev = P.device
        M = G * N
        W = (2 + H) * D
        dXn = _f32c(dXn).view(M, D)
        ns = lib.load().ekaid_edge_bwd_slices(pc.f, D, N, Kn, H, 1 if Phl is not None else 0)
        alloc = torch.zeros if N > Kn else torch.empty
        dQKZ = alloc(M, W, dtype=pc.T, device=dev)
        dOut = torch.empty(M, D, dtype=torch.float32, device=dev)
        dPpart = torch.empty(ns, G, N, H, Kn, dtype=torch.float32, device=dev)
        es = 2 if pc.bf16 else 4
        call("edge_aggregate_bwd", pc.f, dXn.data_ptr(), mask.data_ptr(), P.data_ptr(), QKZ.data_ptr(), QKZ.stride(0), D,
             G, N, Kn, H, dQKZ.data_ptr(), dOut.data_ptr(), dPpart.data_ptr(),
             2.0 / (1.0 - drop.p_gat) if don else 2.0, ptr(Phl),
             info={"bytes": G * (N * D * (4 + 1 + 4) + N * H * Kn * 4 * (1 + ns) + 2 * Kn * H * D * es)})
        kk = ctx.keys
        dbout = _dst(kk["bout"], (D,), dev)          # filled with the other bias gradients of the layer (colsum_many)
        dlb = dgb = None
        if kind == "explicit":
            dlb = torch.empty(H, G, N, Kn, dtype=torch.float32, device=dev)
        else:
            dgb = torch.empty(G, N, Kn, H, dtype=torch.float32, device=dev)
        call("edge_softmax_bwd", pc.f, P.data_ptr(), dPpart.data_ptr(), ns, QKZ.data_ptr(), QKZ.stride(0), D, ptr(cond),
             G, N, Kn, H, dQKZ.data_ptr(), ptr(dlb), ptr(dgb),
             info={"bytes": G * (N * H * Kn * 4 * (2 + ns) + 2 * (N + Kn) * D * es)})
        # gradient of the label-bias table / the geometry FC: leaves of the backward graph (they feed no other gradient),
        # so they run on a side stream next to the dgrad / wgrad GEMMs below instead of in front of them
        dp0 = dp1 = None
        fkp = Fork(dev, 1, pool="leaf")
        if kind == "explicit":
            a0, a1, Lb, S = ctx.geo
            part = torch.empty(G, Lb, dtype=torch.float32, device=dev)
            dp0 = torch.empty(Lb, dtype=torch.float32, device=dev)
            with fkp.branch(0):
                if S:
                    call("adj_labels_bwd", a0.data_ptr(), ptr(a1), ctx.g_split, S, dlb.data_ptr(), H, G, N, Kn, Lb,
                         part.data_ptr())
                else:
                    call("adj_prep_bwd", a0.data_ptr(), ptr(a1), ctx.g_split, dlb.data_ptr(), H, G, N, Kn, Lb,
                         part.data_ptr())
                colsum(part, G, Lb, out=dp0)
            dp0 = dp0.view(1, Lb)
        else:
            a0, a1, Wp, bp, emb_cache = ctx.geo
            nparts = G * lib.load().ekaid_geom_bias_bwd_parts()
            part = torch.empty(nparts, H * 65, dtype=torch.float32, device=dev)
            tot = torch.empty(H * 65, dtype=torch.float32, device=dev)
            dp0 = torch.empty(H, 64, dtype=torch.float32, device=dev)
            dp1 = _dst(kk["p1"], (H,), dev)
            with fkp.branch(0):
                call("geom_bias_bwd", a0.data_ptr(), ptr(a1), ctx.g_split, Wp.data_ptr(), bp.data_ptr(),
                     _dim_t(dev).data_ptr(), G, N, Kn, H, dgb.data_ptr(), part.data_ptr(),
                     *(drop.a(site0 + 4, drop.p_fc) if don else (None, 0, 0.0)), ptr(emb_cache), pc.f)
                colsum(part, nparts, H * 65, out=tot)
                t2 = tot.view(H, 65)
                dp0.copy_(t2[:, :64])
                dp1.copy_(t2[:, 64])
        Dq = qvT.shape[1]
        dbq = _dst(kk["bq"], (D,), dev)
        dbk = _dst(kk["bk"], (D,), dev)
        dbsw = _dst(kk["bsw"], (D,), dev)
        dWo2 = _dst(kk["Wo2"], (D, H * D), dev)
        if don:
            dWq = torch.empty(D, D, dtype=torch.float32, device=dev)
            dWk = torch.empty(D, D, dtype=torch.float32, device=dev)
        dWsw = torch.empty(D, D + Dq, dtype=torch.float32, device=dev)
        dX = torch.empty(M, D, dtype=torch.float32, device=dev)
        if not don:
            # [query | key | Z] projection: ONE wgrad GEMM, then the Z blocks are copied into linear_out_2's layout
            dWqkz = gemm_f32out(dQKZ, Sf, W, D, M, transA=1, transB=1)
            dWq, dWk = dWqkz[:D], dWqkz[D:2 * D]
            cast_many(pc, [(dWqkz[(2 + h) * D:(3 + h) * D], dWo2[:, h * D:(h + 1) * D]) for h in range(H)])
            dSf, _ = gemm_T(pc, dQKZ, WqkzT, M, D, W, transB=1)
            # self_feat = X Wv^T + (flag ? b_sw : q Wq^T + b_sw)
            gemm(dSf, XT, D, D, M, transA=1, transB=1, C=dWsw[:, :D])
            colsum_many([(dOut, dbout), (dQKZ[:, :D], dbq), (dQKZ[:, D:2 * D], dbk), (dSf, dbsw)], M)   # one launch
            dqpart = torch.empty(B, D, dtype=torch.float32, device=dev)
            call("group_rowsum", pc.f, dSf.data_ptr(), dSf.stride(0), N, B, G // B, D, flags.data_ptr(),
                 dqpart.data_ptr())
            dqpT = to_T(pc, dqpart)
            gemm(dqpT, qvT, D, Dq, B, transA=1, transB=1, C=dWsw[:, D:], splits=1)   # K = B
            dqv = gemm_f32out(dqpT, WswT[:, D:], B, Dq, D, transB=1)
            gemm(dSf, WswT[:, :D], M, D, D, transB=1, addend=dXn, C=dX)         # residual + dSf Wv
        else:
            dWz = torch.empty(H * D, D, dtype=torch.float32, device=dev)
            acc = torch.empty(M, D, dtype=torch.float32, device=dev)
            dSf = torch.empty(M, D, dtype=pc.T, device=dev)
            dVq = torch.empty(M, Dq, dtype=torch.float32, device=dev)
            dqv = torch.empty(B, Dq, dtype=torch.float32, device=dev)
            # the weight gradients are leaves of the backward graph: they run on two side streams next to the dgrad chain
            fk = Fork(dev, 2)
            with fk.branch(0):
                gemm(dQKZ[:, 2 * D:W], Sf, W - 2 * D, D, M, transA=1, transB=1, C=dWz)
                cast_many(pc, [(dWz[h * D:(h + 1) * D], dWo2[:, h * D:(h + 1) * D]) for h in range(H)])   # one launch
            with fk.branch(1):
                gemm(dQKZ[:, 0:D], Sq, D, D, M, transA=1, transB=1, C=dWq)
                gemm(dQKZ[:, D:2 * D], Sk, D, D, M, transA=1, transB=1, C=dWk)
            # dSf = mask_q * (dQ Wq) + mask_k * (dK Wk) + dZ Wz: the three products are chained through the GEMM epilogue
            # (dropout mask of the forward's query / key inputs, then "+ addend"), the last one writes the operand type
            gemm(dQKZ[:, 0:D], WqkzT[0:D], M, D, D, transB=1, C=acc, drop=drop.a(site0 + 2, drop.p_fc))
            gemm(dQKZ[:, D:2 * D], WqkzT[D:2 * D], M, D, D, transB=1, addend=acc, C=acc, drop=drop.a(site0 + 3, drop.p_fc))
            gemm(dQKZ[:, 2 * D:W], WqkzT[2 * D:W], M, D, W - 2 * D, transB=1, addend=acc,
                 C=None if pc.bf16 else dSf, Cb=dSf if pc.bf16 else None)
            fk.join()
            # everything below needs dSf: wgrad of self_weights + the bias sums on one side stream, the question half of
            # the input gradient on the other, the node half (+ residual) here
            s1 = drop.a(site0 + 1, drop.p_fc)
            fk = Fork(dev, 2)
            with fk.branch(0):
                gemm(dSf, XT, D, D + Dq, M, transA=1, transB=1, C=dWsw)            # XT = dropped [v | q]
                colsum_many([(dOut, dbout), (dQKZ[:, :D], dbq), (dQKZ[:, D:2 * D], dbk), (dSf, dbsw)], M)   # one launch
            # d[v | q] = dropout-mask * (dSf W_sw): the mask of the forward concat is re-applied in the GEMM epilogue
            # (index = row * (D + Dq) + column).  Tensor-core path: ONE product over all D + Dq columns -- the node half plus
            # the residual gradient goes to dX, the question half to dVq (fp32 output split by column): twice the tiles per
            # launch instead of two 1.4-wave launches.
            if pc.bf16 and FUSE_DVQ:
                gemm(dSf, WswT, M, D + Dq, D, transB=1, addend=dXn, add_n1=D, C=dX, C2=dVq, c_n1=D, drop=s1 + (D + Dq, 0))
                with fk.branch(1, resync=True):          # (the row sum waits for the product just enqueued on this stream)
                    call("group_rowsum", 0, dVq.data_ptr(), dVq.stride(0), N, B, G // B, Dq, flags.data_ptr(),
                         dqv.data_ptr())
            else:
                with fk.branch(1):
                    gemm(dSf, WswT[:, D:], M, Dq, D, transB=1, C=dVq, drop=s1 + (D + Dq, D))
                    call("group_rowsum", 0, dVq.data_ptr(), dVq.stride(0), N, B, G // B, Dq, flags.data_ptr(),
                         dqv.data_ptr())
                gemm(dSf, WswT[:, :D], M, D, D, transB=1, addend=dXn, C=dX, drop=s1 + (D + Dq, 0))   # + residual gradient
            fk.join()
        fkp.join()
        return (None, None, None, None, None, dX, None, dqv, dWsw, dbsw, dWq, dbq, dWk, dbk, dWo2, dbout, dp0, dp1,
                None, None, None, None)


# ------------------------------------------------------------------------------------------------
# difference + gated fusion + attention pooling
# ------------------------------------------------------------------------------------------------
class FusionFn(torch.autograd.Function):
    """modules.py:233-308 for the stacked [bef; aft] residual stream X3 [2*B*N, D].

    Wcg [2D, 2D] = [[context2 | context1], [gate2 | gate1]], bcg = [b_context2 ; b_gate2]; We [dim, 3D], be;
    wa [1, dim], ba [1].  Returns att [2BN] and attended [2B, D]."""

    @staticmethod
    def forward(ctx, pc: PC, drop, dims, mode, coefs, X3, C1, C2, bC2, G1, G2, bG2, We, be, wa, ba):
        lib.require_device()
        B, N, D, dim = dims
        dev = X3.device
        BN = B * N
        M = 2 * BN
        X3 = _f32c(X3).view(M, D)
        c1, c2, c3 = coefs
        Xc = torch.empty(M, D, dtype=torch.float32, device=dev)
        CAT = torch.empty(M, 3 * D, dtype=pc.T, device=dev)
        call("combine_diff_fwd", pc.f, X3.data_ptr(), BN, D, mode, c1, c2, c3, Xc.data_ptr(), CAT.data_ptr())
        # [[context2 | context1], [gate2 | gate1]]: one GEMM over CAT[:, :2D] = [X | diff] gives both pre-activations
        WcgT = torch.empty(2 * D, 2 * D, dtype=pc.T, device=dev)
        bcgc = torch.empty(2 * D, dtype=torch.float32, device=dev)
        We32 = _f32c(We)
        WeT = torch.empty(We32.shape, dtype=pc.T, device=dev) if pc.bf16 else We32
        jobs = [(_f32c(C2), WcgT[:D, :D]), (_f32c(C1), WcgT[:D, D:]), (_f32c(G2), WcgT[D:, :D]), (_f32c(G1), WcgT[D:, D:]),
                (_f32c(bC2).view(1, D), bcgc[:D].view(1, D)), (_f32c(bG2).view(1, D), bcgc[D:].view(1, D))]
        if pc.bf16:
            jobs.append((We32, WeT))
        cast_many(pc, jobs)      # one launch instead of seven
        bec, wac, bac = _f32c(be), _f32c(wa).view(-1), _f32c(ba).view(-1)
        pre = gemm_f32out(CAT[:, :2 * D], WcgT, M, 2 * D, 2 * D, bias=bcgc)
        cx = torch.empty(M, D, dtype=pc.T, device=dev)
        gt = torch.empty(M, D, dtype=pc.T, device=dev)
        don = drop is not None and drop.on
        pf = drop.p_fuse if don else 0.0
        call("gate_fwd", pc.f, pre.data_ptr(), M, D, cx.data_ptr(), gt.data_ptr(), CAT.data_ptr(),
             drop.seed if don else None, 20, 21, pf)
        # embed = Linear -> Dropout(0.5) -> ReLU (modules.py:105-111): dropout fused before the ReLU in the epilogue
        E = gemm_f32out(CAT, WeT, M, dim, 3 * D, bias=bec, act=ACT_RELU,
                        drop=drop.a(22, drop.p_embed) if don else None)
        att = torch.empty(M, dtype=torch.float32, device=dev)
        attended = torch.empty(2 * B, D, dtype=torch.float32, device=dev)
        call("att_pool_fwd", E.data_ptr(), M, N, D, dim, wac.data_ptr(), bac.data_ptr(), Xc.data_ptr(), att.data_ptr(),
             attended.data_ptr())
        ctx.pc, ctx.dims, ctx.mode, ctx.coefs, ctx.drop = pc, dims, mode, coefs, drop
        ctx.keys = {k: t.data_ptr() for k, t in (("C1", C1), ("C2", C2), ("bC2", bC2), ("G1", G1), ("G2", G2),
                                                 ("bG2", bG2), ("We", We), ("be", be), ("wa", wa), ("ba", ba))}
        ctx.saved = (Xc, CAT, WcgT, WeT, wac, cx, gt, E, att)
        if DEBUG_SINK is not None:
            DEBUG_SINK.append((E > 0).cpu())
        # the module's outputs (modules.py:305-313) as views of the two buffers + the difference vector: their five
        # gradients come back together and are folded into (d_att, d_attended) by one kernel
        ia = torch.empty(B, D, dtype=torch.float32, device=dev)
        call("head_fwd", attended.data_ptr(), B * D, ia.data_ptr())
        return att[:BN].view(B, 1, N), att[BN:].view(B, 1, N), attended[:B], attended[B:], ia

    @staticmethod
    def backward(ctx, d_att_bef, d_att_aft, d_a1, d_a2, d_ia):
        pc = ctx.pc
        B, N, D, dim = ctx.dims
        Xc, CAT, WcgT, WeT, wac, cx, gt, E, att = ctx.saved
        dev = Xc.device
        drop = ctx.drop
        don = drop is not None and drop.on
        BN = B * N
        M = 2 * BN
        c1, c2, c3 = ctx.coefs
        gs = [(_f32c(t) if t is not None else None) for t in (d_att_bef, d_att_aft, d_a1, d_a2, d_ia)]
        dA = torch.empty(2 * B, D, dtype=torch.float32, device=dev)
        dw_ = torch.empty(M, dtype=torch.float32, device=dev)
        call("head_bwd", ptr(gs[0]), ptr(gs[1]), ptr(gs[2]), ptr(gs[3]), ptr(gs[4]), BN, B * D, dw_.data_ptr(), dA.data_ptr())
        dXc = torch.empty(M, D, dtype=torch.float32, device=dev)
        dE = torch.empty(M, dim, dtype=pc.T, device=dev)
        dpa = torch.empty(M, dtype=torch.float32, device=dev)
        call("att_pool_bwd", pc.f, dA.data_ptr(), ptr(dw_), att.data_ptr(), Xc.data_ptr(), E.data_ptr(), wac.data_ptr(),
             M, N, D, dim, dXc.data_ptr(), dE.data_ptr(), dpa.data_ptr(),
             1.0 / (1.0 - drop.p_embed) if don else 1.0)
        kk = ctx.keys
        # every buffer first (on this stream), then the weight / bias gradients -- leaves of the backward graph -- go to
        # side streams and run next to the dgrad chain
        dwa = _dst(kk["wa"], (1, dim), dev)
        dba = _dst(kk["ba"], (1,), dev)
        dWe = _dst(kk["We"], (dim, 3 * D), dev)
        dbe = _dst(kk["be"], (dim,), dev)
        dCAT = torch.empty(M, 3 * D, dtype=torch.float32, device=dev)
        dpre = torch.empty(M, 2 * D, dtype=pc.T, device=dev)
        dWcg = torch.empty(2 * D, 2 * D, dtype=torch.float32, device=dev)
        blocks = {name: _dst(kk[name], (D, D), dev) for name in ("C2", "C1", "G2", "G1")}
        dbC2 = _dst(kk["bC2"], (D,), dev)
        dbG2 = _dst(kk["bG2"], (D,), dev)
        dX3 = torch.empty(M, D, dtype=torch.float32, device=dev)
        fk = Fork(dev, 2)
        with fk.branch(0):
            colsum(E, M, dim, rowscale=dpa, out=dwa.view(dim))
            colsum(dpa.view(-1, 1), M, 1, out=dba)
            gemm_f32out(dE, CAT, dim, 3 * D, M, transA=1, transB=1, out=dWe)
            colsum(dE, M, dim, out=dbe)
        gemm_f32out(dE, WeT, M, 3 * D, dim, transB=1, out=dCAT)
        call("gate_bwd", pc.f, dCAT.data_ptr(), cx.data_ptr(), gt.data_ptr(), M, D, dpre.data_ptr(),
             drop.seed if don else None, 20, 21, drop.p_fuse if don else 0.0)
        with fk.branch(1, resync=True):
            gemm_f32out(dpre, CAT[:, :2 * D], 2 * D, 2 * D, M, transA=1, transB=1, out=dWcg)
            # blocks of the fused [[context2 | context1], [gate2 | gate1]] gradient -> parameter-shaped (contiguous) tensors
            cast_many(pc, [(dWcg[r0:r0 + D, c0:c0 + D], blocks[name])
                           for name, r0, c0 in (("C2", 0, 0), ("C1", 0, D), ("G2", D, 0), ("G1", D, D))])     # one launch
            colsum(dpre[:, :D], M, D, out=dbC2)
            colsum(dpre[:, D:], M, D, out=dbG2)
        gemm(dpre, WcgT, M, 2 * D, 2 * D, transB=1, addend=dCAT[:, :2 * D], C=dCAT[:, :2 * D])
        call("combine_diff_bwd", dXc.data_ptr(), dCAT.data_ptr(), BN, D, ctx.mode, c1, c2, c3, dX3.data_ptr())
        fk.join()
        return (None, None, None, None, None, dX3, blocks["C1"], blocks["C2"], dbC2, blocks["G1"], blocks["G2"], dbG2, dWe,
                dbe, dwa, dba)


# ------------------------------------------------------------------------------------------------
# step glue
# ------------------------------------------------------------------------------------------------
def spatial_labels(boxes: torch.Tensor, size: int = 100, lx: float = 1024.0, ly: float = 1024.0) -> torch.Tensor:
    """get_adj_matrix / bbox_relation_type ("feature extraction/ana_bbox_generator.py":266-302,320-335) as one kernel:
    boxes [B,N,4] (xmin,ymin,xmax,ymax, any real dtype) on the device -> float64 labels [B,S,S], S = max(size, N),
    the layout the loader hands to process_matrix / `onehot_adj`."""
    lib.require_device()
    if boxes.dim() != 3 or boxes.shape[-1] != 4:
        raise ValueError("spatial_labels: boxes must be [B, N, 4], got %s" % (tuple(boxes.shape),))
    bb = boxes.detach()
    if bb.dtype != torch.float64:
        bb = bb.double()
    bb = bb.contiguous()
    Bn, N = bb.shape[0], bb.shape[1]
    S = max(int(size), N)
    out = torch.empty(Bn, S, S, dtype=torch.float64, device=bb.device)
    call("spatial_labels", bb.data_ptr(), Bn, N, S, float(lx), float(ly), out.data_ptr())
    return out


def semantic_tables(ana_classes, di_classes, kg_ana: dict, small_name2index: dict, small_adj, device):
    """The reference's dictionaries ("feature extraction/combine_dicts.py":106-151: class-name lists, knowledge-graph organ
    of every class, the co-occurrence matrix of the CheXpert diseases and its name index) as the per-class device tables
    `semantic_labels` takes."""
    names = list(ana_classes) + list(di_classes)
    organs = {o: k for k, o in enumerate(sorted({kg_ana[n] for n in names}))}
    ana_set, di_set = set(ana_classes), set(di_classes)
    t = lambda x, dt: torch.tensor(x, dtype=dt, device=device)      # noqa: E731
    return {"ncls": len(names), "group": t([organs[kg_ana[n]] for n in names], torch.int32),
            "in_ana": t([n in ana_set for n in names], torch.uint8), "in_di": t([n in di_set for n in names], torch.uint8),
            "small_idx": t([small_name2index.get(n.lower(), -1) for n in names], torch.int32),
            "small_adj": torch.as_tensor(small_adj).to(device=device, dtype=torch.int32).contiguous()}


def semantic_labels(classes: torch.Tensor, tables: dict, size: int = 100) -> torch.Tensor:
    """get_semantic_adj ("feature extraction/combine_dicts.py":106-151) as one kernel: detected classes [B,T] (anatomy ids,
    then disease ids offset by the anatomy count; tables['ncls'] = background) -> int8 labels [B,S,S], S = max(size, T)."""
    lib.require_device()
    cls = classes.detach().to(torch.int32).contiguous()
    Bn, T = cls.shape
    S = max(int(size), T)
    out = torch.empty(Bn, S, S, dtype=torch.int8, device=cls.device)
    sa = tables["small_adj"]
    call("semantic_labels", cls.data_ptr(), Bn, T, S, int(tables["ncls"]), tables["group"].data_ptr(),
         tables["in_ana"].data_ptr(), tables["in_di"].data_ptr(), tables["small_idx"].data_ptr(), sa.data_ptr(),
         int(sa.shape[0]), out.data_ptr())
    return out


def onehot_adj(labels: torch.Tensor, num_objects: int, label_num: int) -> torch.Tensor:
    """process_matrix (utils/mimic_utils.py:141-149) as one kernel: labels [B,S,S] (any real dtype) on the device
    -> fp32 one-hot [B,N,N,L]."""
    lib.require_device()
    lab = labels.detach()
    if lab.dtype == torch.int8:
        lab = lab.contiguous()
        Bn, S = lab.shape[0], lab.shape[1]
        out = torch.empty(Bn, num_objects, num_objects, label_num, dtype=torch.float32, device=lab.device)
        call("onehot_adj_i8", lab.data_ptr(), Bn, S, num_objects, label_num, out.data_ptr())
        return out
    if lab.dtype != torch.float64:
        lab = lab.double()
    lab = lab.contiguous()
    Bn, S = lab.shape[0], lab.shape[1]
    out = torch.empty(Bn, num_objects, num_objects, label_num, dtype=torch.float32, device=lab.device)
    call("onehot_adj", lab.data_ptr(), Bn, S, num_objects, label_num, out.data_ptr())
    return out
