"""ctypes binding of libekaid_b200.so (the C ABI declared in include/ekaid_b200.h).

The prototypes are parsed from the header so the binding can never drift from the declared ABI.
There is no CPU fallback: if the shared library is missing the import of a kernel fails loudly.
"""
from __future__ import annotations

import contextlib
import ctypes
import gc
import os
import re
from typing import Dict, List, Tuple

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(_HERE), "include", "ekaid_b200.h")
LIB_PATH = os.path.join(_HERE, "libekaid_b200.so")


class Epilogue(ctypes.Structure):
    """struct ekaid_epilogue (include/ekaid_b200.h)."""
    _fields_ = [
        ("bias", ctypes.c_void_p), ("addend", ctypes.c_void_p), ("ldadd", ctypes.c_int64),
        ("rowb", ctypes.c_void_p), ("ldrowb", ctypes.c_int64), ("rowb_div", ctypes.c_int32),
        ("rowb_mod", ctypes.c_int32), ("rowflag", ctypes.c_void_p), ("rowb_alt", ctypes.c_void_p),
        ("act", ctypes.c_int32), ("drop_seed", ctypes.c_void_p), ("drop_site", ctypes.c_uint32),
        ("drop_p", ctypes.c_float), ("drop_n", ctypes.c_int32), ("drop_off", ctypes.c_int32),
        ("C", ctypes.c_void_p), ("ldc", ctypes.c_int64),
        ("Cb", ctypes.c_void_p), ("ldcb", ctypes.c_int64),
        ("cb_fmt", ctypes.c_int32), ("cb_n1", ctypes.c_int32),
        ("Cb2", ctypes.c_void_p), ("ldcb2", ctypes.c_int64),
        ("cb2_fmt", ctypes.c_int32), ("cb2_n0", ctypes.c_int32),
        ("C2", ctypes.c_void_p), ("ldc2", ctypes.c_int64), ("c_n1", ctypes.c_int32), ("add_n1", ctypes.c_int32),
    ]


def parse_header(path: str = HEADER) -> Dict[str, Tuple[str, List[str]]]:
    """{name: (return type, [param types])} for every `ekaid_*` prototype in the header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(int|const char\*)\s+(ekaid_\w+)\s*\(([^)]*)\)\s*;", src):
        ret, name, params = m.group(1), m.group(2), m.group(3).strip()
        types = []
        if params and params != "void":
            for p in params.split(","):
                p = " ".join(p.split())
                if "*" in p:
                    types.append("ptr")
                else:
                    types.append(p.rsplit(" ", 1)[0])
        protos[name] = (ret, types)
    return protos


_CT = {"int": ctypes.c_int, "int32_t": ctypes.c_int32, "int64_t": ctypes.c_int64, "uint32_t": ctypes.c_uint32,
       "float": ctypes.c_float, "double": ctypes.c_double,
       "ptr": ctypes.c_void_p}

_lib = None


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m ekaid_b200.build` "
            "(there is deliberately no CPU / PyTorch fallback for the CUDA path)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (ret, types) in parse_header().items():
        fn = getattr(lib, name)
        fn.restype = ctypes.c_char_p if ret != "int" else ctypes.c_int
        fn.argtypes = [_CT[t] for t in types]
    _lib = lib
    return lib


class EkaidError(RuntimeError):
    pass


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().ekaid_last_error().decode()
        if rc in (-1, -2, -5):
            raise ValueError(f"{what}: {msg} (code {rc})")
        raise EkaidError(f"{what}: {msg} (code {rc})")


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t) -> int | None:
    return None if t is None else t.data_ptr()


# kernels launched per C-ABI call (for the bench's `gpu_launches` count)
KERNELS_PER_CALL = {"att_pool_fwd": 2, "qpool_fwd": 2, "qpool_bwd": 2, "wn_fwd": 2, "wn_bwd": 2, "wn_fwd_many": 2,
                    "wn_bwd_many": 2}
LAUNCHES = 0
# bench only: when a list, every call is bracketed by CUDA events on the launching stream:
# (name, start_event, end_event, info)
PROFILE = None
PROFILE_EXTERNAL = False      # events that may be recorded inside a CUDA-graph capture (event-record nodes)


def call(name: str, *args, info=None) -> None:
    """Call `ekaid_<name>` on the current torch CUDA stream (appended as last argument)."""
    global LAUNCHES
    lib = load()
    fn = getattr(lib, "ekaid_" + name)
    if PROFILE is None:
        rc = fn(*args, stream_ptr())
    else:
        e0 = torch.cuda.Event(enable_timing=True, external=PROFILE_EXTERNAL)
        e1 = torch.cuda.Event(enable_timing=True, external=PROFILE_EXTERNAL)
        e0.record()
        rc = fn(*args, stream_ptr())
        e1.record()
        PROFILE.append((name, e0, e1, info))
    LAUNCHES += KERNELS_PER_CALL.get(name, 1)
    check(rc, name)


@contextlib.contextmanager
def graph_capture(graph: "torch.cuda.CUDAGraph"):
    """`with torch.cuda.graph(graph)` with Python's cyclic garbage collector held off for the duration of the capture.
    torch captures in the `global` error mode, where a prohibited runtime call from anywhere in the process invalidates the
    capture -- and torch.cuda.graph no longer collects garbage before it starts (torch/cuda/graphs.py, `force_cudagraph_gc`).
    A collection that happens to run while the step is being captured and frees an OLD CUDAGraph (a runner of another
    batch size, a previous GraphFusionStep: they sit in reference cycles) calls cudaGraphExecDestroy mid-capture, and every
    later launch fails with cudaErrorStreamCaptureInvalidated.  So: collect first, then keep the collector off inside."""
    gc.collect()
    was_enabled = gc.isenabled()
    gc.disable()
    try:
        with torch.cuda.graph(graph):
            yield
    finally:
        if was_enabled:
            gc.enable()


_device_ok = False


def require_device() -> None:
    global _device_ok
    if _device_ok:
        return
    if not torch.cuda.is_available():
        raise EkaidError("ekaid_b200 needs a CUDA device (sm_100a); no CPU fallback exists")
    check(load().ekaid_check_device(), "check_device")
    _device_ok = True
