"""ctypes binding of libtdb.so (the C ABI declared in include/tubedetr_b200.h).

There is no fallback: if the library is missing or a call fails, a RuntimeError is raised.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libtdb.so")
MAX_TAPS = 9
OUT_BF16, OUT_F32 = 0, 1
REMAP_NONE, REMAP_C2P, REMAP_P2C, REMAP_C2S, REMAP_S2C, REMAP_C2P1 = 0, 1, 2, 3, 4, 5


class GemmDesc(C.Structure):
    _fields_ = [
        ("A", C.c_void_p), ("a_rows", C.c_int64), ("a_cols", C.c_int64), ("lda", C.c_int64), ("a_major", C.c_int32),
        ("B", C.c_void_p), ("b_rows", C.c_int64), ("b_cols", C.c_int64), ("ldb", C.c_int64), ("b_major", C.c_int32),
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32), ("ntaps", C.c_int32),
        ("a_off0", C.c_int32 * MAX_TAPS), ("a_off1", C.c_int32 * MAX_TAPS),
        ("b_off0", C.c_int32 * MAX_TAPS), ("b_off1", C.c_int32 * MAX_TAPS),
        ("nz", C.c_int32), ("z_b_off1", C.c_int32 * MAX_TAPS), ("z_out_col", C.c_int32 * MAX_TAPS),
        ("splits", C.c_int32),
        ("scale", C.c_void_p), ("bias", C.c_void_p),
        ("residual", C.c_void_p), ("ldr", C.c_int64),
        ("mask", C.c_void_p), ("ldmask", C.c_int64),
        ("relu", C.c_int32),
        ("out", C.c_void_p), ("out_dtype", C.c_int32), ("ldo", C.c_int64),
        ("remap", C.c_int32), ("img_h", C.c_int32), ("img_w", C.c_int32),
        ("block_n", C.c_int32), ("max_ctas", C.c_int32), ("debug_flags", C.c_int32),
    ]


LOSS_MAX_LAYERS = 8


class LossDesc(C.Structure):
    _fields_ = [
        ("nlayers", C.c_int32), ("K", C.c_int32), ("B", C.c_int32), ("T", C.c_int32),
        ("pred_boxes", C.c_void_p * LOSS_MAX_LAYERS), ("pred_sted", C.c_void_p * LOSS_MAX_LAYERS),
        ("weights", C.c_void_p * LOSS_MAX_LAYERS),
        ("tgt_boxes", C.c_void_p), ("num_boxes", C.c_void_p), ("gauss", C.c_void_p), ("time_mask", C.c_void_p),
        ("neg", C.c_void_p), ("nneg", C.c_void_p),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"tubedetr_b200: {LIB_PATH} is missing -- run `python -m tubedetr_b200.build` "
                               "(there is no CPU/PyTorch fallback for the CUDA path)")
        _lib = C.CDLL(LIB_PATH)
        _lib.tdb_last_error_string.restype = C.c_char_p
        _lib.tdb_launch_count.restype = C.c_int64
    return _lib


def check(rc, what=""):
    if rc != 0:
        raise RuntimeError(f"tubedetr_b200 {what} failed (rc={rc}): {lib().tdb_last_error_string().decode()}")


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def launch_count():
    return int(lib().tdb_launch_count())
