"""ctypes binding of libfsmg.so (the C-ABI declared in include/fsmg.h).

The product path fails loudly when the CUDA library is missing: there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("FSMG_LIB", _HERE / "libfsmg.so"))

FSMG_OK = 0
FSMG_FLAG_SIMT_GEMM = 1
FSMG_FLAG_SIMT_RECURRENT = 2
FSMG_GRAD_EXTRA = 8
FSMG_PROF_PHASES = 11


class FsmgError(RuntimeError):
    pass


class fsmg_config(C.Structure):
    _fields_ = [
        ("vocab", C.c_int32), ("embed", C.c_int32), ("hidden", C.c_int32), ("layers", C.c_int32),
        ("max_len", C.c_int32), ("max_seqs", C.c_int32), ("n_decay", C.c_int32), ("flags", C.c_int32),
        ("lr", C.c_float), ("max_grad_norm", C.c_float),
        ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float), ("reserved", C.c_float),
    ]


class fsmg_param_info(C.Structure):
    _fields_ = [("name", C.c_char * 96), ("offset", C.c_int64), ("rows", C.c_int32), ("cols", C.c_int32)]


# every symbol include/fsmg.h declares: (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "fsmg_last_error": (C.c_char_p, []),
    "fsmg_abi_version": (C.c_int, []),
    "fsmg_create": (C.c_int, [C.POINTER(fsmg_config), C.c_char_p, C.POINTER(_P)]),
    "fsmg_destroy": (None, [_P]),
    "fsmg_param_count": (C.c_int64, [_P]),
    "fsmg_grad_count": (C.c_int64, [_P]),
    "fsmg_workspace_bytes": (C.c_int64, [_P]),
    "fsmg_num_params": (C.c_int, [_P]),
    "fsmg_param_info_at": (C.c_int, [_P, C.c_int, C.POINTER(fsmg_param_info)]),
    "fsmg_bind": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int64]),
    "fsmg_refresh_weights": (C.c_int, [_P, _P]),
    "fsmg_forward_nll": (C.c_int, [_P, _P, C.c_int32, _P, _P, _P]),
    "fsmg_forward_backward": (C.c_int, [_P, _P, C.c_int32, C.c_float, _P, _P]),
    "fsmg_set_stage_events": (C.c_int, [_P, _P, _P, C.c_int32]),
    "fsmg_set_loss_event": (C.c_int, [_P, _P]),
    "fsmg_param_range": (C.c_int, [_P, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "fsmg_apply_update": (C.c_int, [_P, C.c_int64, _P, _P]),
    "fsmg_sample_greedy": (C.c_int, [_P, C.c_int32, C.c_int32, _P, _P]),
    "fsmg_eval_host": (C.c_int, [_P, _P, C.c_int32, C.POINTER(C.c_float), _P, _P]),
    "fsmg_train_host": (C.c_int, [_P, _P, C.c_int32, C.c_int64, C.POINTER(C.c_float), _P]),
    "fsmg_sample_host": (C.c_int, [_P, C.c_int32, C.c_int32, _P, _P]),
    "fsmg_set_profile": (C.c_int, [_P, C.c_int]),
    "fsmg_read_profile": (C.c_int, [_P, _P, _P, _P]),
    "fsmg_profile_phase_name": (C.c_char_p, [C.c_int]),
    "fsmg_last_launch_count": (C.c_int64, [_P]),
    "fsmg_token_range_errors": (C.c_int, [_P, C.POINTER(C.c_int64), _P]),
    "fsmg_debug_prep_tokens": (C.c_int, [_P, _P, C.c_int32, _P, _P, _P]),
    "fsmg_debug_mma_probe": (C.c_int, [C.c_int32, C.c_int32, _P, _P]),
    "fsmg_debug_gemm": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, _P]),
    "fsmg_gather_token_rows": (C.c_int, [_P, C.c_int64, C.c_int32, _P, C.c_int32, _P, _P]),
    "fsmg_unigram_step": (C.c_int, [_P, C.c_int32, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P]),
    "fsmg_unigram_argmax": (C.c_int, [_P, C.c_int32, _P, _P]),
    "fsmg_debug_plan": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32)]),
    "fsmg_debug_gemm_xf": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, _P, C.c_int64, _P, C.c_int64, _P, C.c_int32, _P, C.c_int64, _P, _P,
                                    C.c_float, _P, _P]),
    "fsmg_debug_softmax_grad": (C.c_int, [C.c_int32, C.c_int32, C.c_int64, _P, _P, _P, C.c_float, _P, C.c_int32, C.c_int32, C.c_int32, _P]),
}

_lib = None


def load() -> C.CDLL:
    """Load libfsmg.so and type every entry point.  Raises FsmgError if the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise FsmgError(
            f"{LIB_PATH} not found: build it with `python __graft_entry__.py build` "
            "(nvcc -gencode arch=compute_100a,code=sm_100a). There is no CPU fallback.")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != FSMG_OK:
        msg = load().fsmg_last_error()
        raise FsmgError(f"libfsmg error {rc}: {msg.decode() if msg else '?'}")
