"""ctypes binding of ``libaudiomps.so`` (include/audiomps.h).

There is NO fallback: if the shared library is missing or a symbol cannot be bound, importing
the compute path raises.  ``build()`` compiles it in-tree with nvcc for sm_100a.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# AMPS_LIB: alternative build of the same library (A/B timing of kernel variants, profiles/build_variants.sh)
LIB_PATH = os.environ.get("AMPS_LIB") or os.path.join(_HERE, "libaudiomps.so")
CSRC = os.path.join(_HERE, "csrc")
HEADER = os.path.join(os.path.dirname(_HERE), "include", "audiomps.h")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
]

AMPS_OK = 0
ERRORS = {-1: "AMPS_E_INVALID", -2: "AMPS_E_UNSUPPORTED", -3: "AMPS_E_WORKSPACE",
          -4: "AMPS_E_CUDA", -5: "AMPS_E_STATE"}


class AmpsError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{ERRORS.get(code, code)}: {msg}")
        self.code = code


class AmpsParams(C.Structure):
    _fields_ = [("D", C.c_int32), ("reserved", C.c_int32),
                ("R_dev", C.c_void_p), ("freqs_dev", C.c_void_p),
                ("psi0_dev", C.c_void_p), ("rho0_dev", C.c_void_p),
                ("A", C.c_float), ("sigma", C.c_float), ("delta_t", C.c_double),
                ("A_dev", C.c_void_p)]


class AmpsHostParams(C.Structure):
    _fields_ = [("D", C.c_int32), ("reserved", C.c_int32),
                ("R", C.c_void_p), ("freqs", C.c_void_p), ("psi0", C.c_void_p),
                ("A", C.c_float), ("sigma", C.c_float), ("delta_t", C.c_double)]


# name -> (restype, argtypes); every symbol include/audiomps.h declares
SYMBOLS = {
    "amps_version": (C.c_int, []),
    "amps_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "amps_destroy": (C.c_int, [C.c_void_p]),
    "amps_last_error": (C.c_char_p, [C.c_void_p]),
    "amps_launch_count": (C.c_int64, [C.c_void_p]),
    "amps_set_profiling": (C.c_int, [C.c_void_p, C.c_int]),
    "amps_get_kernel_ms": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_float)]),
    "amps_fma_peak_tflops": (C.c_double, [C.c_void_p]),
    "amps_fma_peak_tflops2": (C.c_double, [C.c_void_p, C.c_int]),
    "amps_psi_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "amps_psi_grad_count": (C.c_size_t, [C.c_int]),
    "amps_psi_loss_fwd": (C.c_int, [C.c_void_p, C.POINTER(AmpsParams), C.c_void_p, C.c_int, C.c_int,
                                    C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]),
    "amps_psi_loss_bwd": (C.c_int, [C.c_void_p, C.POINTER(AmpsParams), C.c_void_p, C.c_int, C.c_int,
                                    C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "amps_psi_ckpt_interval": (C.c_int, [C.c_int, C.c_int]),
    "amps_psi_workspace_bytes_k": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "amps_psi_loss_fwd_k": (C.c_int, [C.c_void_p, C.POINTER(AmpsParams), C.c_void_p, C.c_int, C.c_int, C.c_int,
                                      C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "amps_psi_loss_bwd_k": (C.c_int, [C.c_void_p, C.POINTER(AmpsParams), C.c_void_p, C.c_int, C.c_int, C.c_int,
                                      C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "amps_psi_scan_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "amps_psi_loss_fwd_scan": (C.c_int, [C.c_void_p, C.POINTER(AmpsParams), C.c_void_p, C.c_int, C.c_int,
                                         C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]),
    "amps_psi_loss_bwd_scan": (C.c_int, [C.c_void_p, C.POINTER(AmpsParams), C.c_void_p, C.c_int, C.c_int,
                                         C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "amps_psi_params_fwd": (C.c_int, [C.c_void_p, C.c_int] + [C.c_void_p] * 5 + [C.c_float] * 4 + [C.c_void_p] * 5),
    "amps_psi_params_bwd": (C.c_int, [C.c_void_p, C.c_int] + [C.c_void_p] * 5 + [C.c_float] * 4 + [C.c_void_p] * 11),
    "amps_comm_unique_id": (C.c_int, [C.c_void_p]),
    "amps_comm_init": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "amps_allreduce_grads": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "amps_comm_destroy": (C.c_int, [C.c_void_p]),
    "amps_time_table_host": (C.c_int, [C.c_double, C.c_int, C.c_void_p]),
    "amps_psi_sample_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "amps_psi_sample": (C.c_int, [C.c_void_p, C.POINTER(AmpsParams), C.c_void_p, C.c_int, C.c_int,
                                  C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "amps_psi_evolve": (C.c_int, [C.c_void_p, C.POINTER(AmpsParams), C.c_void_p, C.c_int, C.c_int,
                                  C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "amps_psi_loss_grad_host": (C.c_int, [C.c_void_p, C.POINTER(AmpsHostParams), C.c_void_p, C.c_int,
                                          C.c_int, C.c_float, C.c_void_p, C.c_void_p]),
    "amps_rho_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "amps_rho_grad_count": (C.c_size_t, [C.c_int]),
    "amps_rho_loss_fwd": (C.c_int, [C.c_void_p, C.POINTER(AmpsParams), C.c_void_p, C.c_int, C.c_int,
                                    C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]),
    "amps_rho_loss_bwd": (C.c_int, [C.c_void_p, C.POINTER(AmpsParams), C.c_void_p, C.c_int, C.c_int,
                                    C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "amps_rho_evolve": (C.c_int, [C.c_void_p, C.POINTER(AmpsParams), C.c_void_p, C.c_int, C.c_int,
                                  C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "amps_rho_sample": (C.c_int, [C.c_void_p, C.POINTER(AmpsParams), C.c_void_p, C.c_int, C.c_int,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                  C.c_void_p]),
}


def build(verbose: bool = False) -> str:
    """Compile libaudiomps.so in-tree with nvcc for sm_100a (no GPU needed)."""
    src = os.path.join(CSRC, "amps_api.cu")
    deps = [src, HEADER] + [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    if os.path.exists(LIB_PATH):
        newest = max(os.path.getmtime(d) for d in deps)
        if os.path.getmtime(LIB_PATH) >= newest:
            return LIB_PATH
    cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH, src]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB_PATH


_lock = threading.Lock()
_lib = None


def load() -> C.CDLL:
    """Load the library and bind every declared symbol; raises if anything is missing."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: the CUDA extension has not been built "
                "(run `python -c 'import __graft_entry__ as g; g.build()'`). "
                "There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SYMBOLS.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is not exported
            fn.restype = restype
            fn.argtypes = argtypes
        if lib.amps_version() < 200:
            raise ImportError("libaudiomps.so is older than this Python package")
        _lib = lib
        return lib


_ctx = {}


def context(device_index: int) -> C.c_void_p:
    """One amps_ctx per device per process."""
    lib = load()
    with _lock:
        h = _ctx.get(device_index)
        if h is None:
            h = C.c_void_p()
            rc = lib.amps_create(int(device_index), C.byref(h))
            if rc != AMPS_OK:
                raise AmpsError(rc, f"amps_create(device={device_index}) failed (is a CUDA GPU visible?)")
            _ctx[device_index] = h
        return h


def check(ctx, rc: int):
    if rc != AMPS_OK:
        raise AmpsError(rc, load().amps_last_error(ctx).decode(errors="replace"))


def launch_count(device_index: int = 0) -> int:
    return int(load().amps_launch_count(context(device_index)))


def set_profiling(device_index: int, enable: bool):
    check(context(device_index), load().amps_set_profiling(context(device_index), 1 if enable else 0))


def kernel_ms(device_index: int, which: int) -> float:
    ms = C.c_float()
    h = context(device_index)
    check(h, load().amps_get_kernel_ms(h, which, C.byref(ms)))
    return float(ms.value)
