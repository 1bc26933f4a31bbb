"""ctypes binding of ``libpaos_b200.so`` (the C ABI declared in ``include/paos_b200.h``).

There is deliberately no fallback: if the shared library has not been built (``python paos_b200/build.py``)
importing this module raises, and every entry point fails with ``PaosCudaError`` when no sm_100 device is
usable.  Nothing in this package computes a wavefront on the CPU.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# PAOS_LIB: an experiment variant built by `PAOS_BUILD_TAG=... python paos_b200/build.py` (profiling tools only)
LIB_PATH = os.environ.get("PAOS_LIB") or os.path.join(HERE, "libpaos_b200.so")

PAOS_OK = 0
PAOS_ERR_ARG, PAOS_ERR_CUDA, PAOS_ERR_STATE, PAOS_ERR_UNSUPPORTED = -1, -2, -3, -4
PAOS_C128, PAOS_C64 = 0, 1
READ_WFO, READ_AMPLITUDE, READ_PHASE, READ_PSF = 0, 1, 2, 3
SHAPE_ELLIPSE, SHAPE_RECT = 0, 1
ABI_VERSION = 3
MAX_CHAINED_FFTS = 16


class PaosError(RuntimeError):
    """Base class of errors reported by libpaos_b200."""


class PaosCudaError(PaosError):
    """CUDA runtime failure or no usable sm_100 device (there is no CPU fallback)."""


class PaosStats(C.Structure):
    _fields_ = [
        ("kernel_launches", C.c_uint64),
        ("pass_launches", C.c_uint64),
        ("passes_planned", C.c_uint64),
        ("fft2_recorded", C.c_uint64),
        ("line_ffts_run", C.c_uint64),
        ("lines_transformed", C.c_uint64),
        ("lines_tabled", C.c_uint64),
        ("lines_swept", C.c_uint64),
        ("host_plan_us", C.c_uint64),
        ("last_flush_ms", C.c_double),
    ]


_vp, _i, _d = C.c_void_p, C.c_int, C.c_double
_dp, _ip = C.POINTER(C.c_double), C.POINTER(C.c_int)

# name -> (restype, argtypes); one entry per function declared in include/paos_b200.h
SIGNATURES = {
    "paos_abi_version": (_i, []),
    "paos_last_error": (C.c_char_p, []),
    "paos_build_info": (C.c_char_p, []),
    "paos_device_count": (_i, []),
    "paos_wfo_create": (_i, [C.POINTER(_vp), _i, _i, _i, _vp, _vp]),
    "paos_wfo_destroy": (_i, [_vp]),
    "paos_wfo_reset": (_i, [_vp]),
    "paos_wfo_fill_ones": (_i, [_vp]),
    "paos_wfo_flush": (_i, [_vp]),
    "paos_wfo_sync": (_i, [_vp]),
    "paos_wfo_upload": (_i, [_vp, _vp]),
    "paos_wfo_upload_device": (_i, [_vp, _vp]),
    "paos_wfo_read": (_i, [_vp, _i, _vp]),
    "paos_wfo_read_device": (_i, [_vp, _i, _vp]),
    "paos_wfo_aperture": (_i, [_vp, _i, _d, _d, _d, _d, _d, _i]),
    "paos_wfo_make_stop": (_i, [_vp]),
    "paos_wfo_quadphase": (_i, [_vp, _d, _d, _d, _d]),
    "paos_wfo_phase_screen": (_i, [_vp, _vp, _d]),
    "paos_wfo_phase_screen_device": (_i, [_vp, _vp, _d]),
    "paos_wfo_zernike": (_i, [_vp, _i, _ip, _ip, _dp, _d, _d, _d, _d, _i, _d, _vp]),
    "paos_wfo_zernike_masked": (_i, [_vp, _i, _ip, _ip, _dp, _d, _d, _d, _d, _i, _d, _vp, _vp]),
    "paos_zernike_cov": (_i, [_vp, _i, _ip, _ip, _vp, _d, _d, _d, _d, _i, _vp, _vp]),
    "paos_wfo_grid_sag": (_i, [_vp, _vp, _vp, _i, _i, _d, _d, _d, _d, _d, _d, _d, _vp, _vp]),
    "paos_fourier_shift_kernel": (_i, [_i, _d, _vp, _vp]),
    "paos_grid_sag_cache_clear": (_i, []),
    "paos_wfo_psd": (_i, [_vp, _d, _d, _d, _d, _d, _d, _d, _d, _d, _d, _d, _vp, _vp, C.c_uint64, _vp]),
    "paos_wfo_ptp": (_i, [_vp, _d, _d, _d, _d]),
    "paos_wfo_stw": (_i, [_vp, _d, _d, _d, _d]),
    "paos_wfo_wts": (_i, [_vp, _d, _d, _d, _d]),
    "paos_wfo_fft2": (_i, [_vp, _i]),
    "paos_abi_struct_size": (C.c_long, [_i]),
    "paos_wfo_materialize": (_i, [_vp]),
    "paos_wfo_read_device_final": (_i, [_vp, _i, _vp]),
    "paos_chain_run": (_i, [_vp, _d, _d, _d, _d, _d, _vp, _i, _vp, _i, C.POINTER(C.c_int), _vp]),
    "paos_zernike_points": (_i, [_i, _i, _ip, _ip, _vp, _vp, _vp, _vp, _vp, C.c_size_t, _vp, _vp]),
    "paos_encircled_energy": (_i, [_vp, _vp, _d, _d, _d, _d, _d, _d, _i, _vp]),
    "paos_psf_peak": (_i, [_vp, _vp, _vp]),
    "paos_screen_stats": (_i, [_vp, _vp, _d, _d, _d, _vp]),
    "paos_crop_convert": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "paos_comm_last_error": (C.c_char_p, []),
    "paos_comm_unique_id": (_i, [_vp]),
    "paos_comm_create": (_i, [C.POINTER(_vp), _vp, _i, _i, _i]),
    "paos_comm_destroy": (_i, [_vp]),
    "paos_gather_psf": (_i, [_vp, _vp, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), _vp, _i, _vp]),
    "paos_wfo_stats": (_i, [_vp, C.POINTER(PaosStats)]),
    "paos_wfo_enable_timing": (_i, [_vp, _i]),
    "paos_wfo_timing": (_i, [_vp, _dp, C.POINTER(C.c_uint64)]),
    "paos_wfo_timing_detail": (_i, [_vp, _i, _i, _dp, C.POINTER(C.c_uint64), _i]),
    "paos_wfo_timing_totals": (_i, [_vp, _dp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), _i]),
    "paos_batch_capacity": (_i, []),
    "paos_wfo_begin_record": (_i, [_vp]),
    "paos_batch_execute": (_i, [C.POINTER(_vp), _i]),
    "paos_batch_chain_run": (_i, [C.POINTER(_vp), _i, _vp]),
}

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python paos_b200/build.py` (nvcc, sm_100a). "
        "paos_b200 has no CPU fallback."
    )

lib = C.CDLL(LIB_PATH)
for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)
    _fn.restype = _res
    _fn.argtypes = _args

if lib.paos_abi_version() != ABI_VERSION:
    raise ImportError(f"libpaos_b200 ABI {lib.paos_abi_version()} != binding ABI {ABI_VERSION}; rebuild the library")


def last_error():
    msg = lib.paos_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc):
    """Raise the Python exception matching a non-zero status."""
    if rc == PAOS_OK:
        return
    msg = last_error()
    if rc == PAOS_ERR_CUDA:
        raise PaosCudaError(msg)
    if rc == PAOS_ERR_ARG:
        raise ValueError(msg)
    if rc == PAOS_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise PaosError(msg)


def check_comm(rc):
    """Same for the communicator entry points (their messages come from paos_comm_last_error)."""
    if rc == PAOS_OK:
        return
    msg = lib.paos_comm_last_error()
    msg = msg.decode("utf-8", "replace") if msg else ""
    if rc == PAOS_ERR_ARG:
        raise ValueError(msg)
    if rc == PAOS_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise PaosCudaError(msg)


def device_count():
    return lib.paos_device_count()
