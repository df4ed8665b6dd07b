"""ctypes binding of libgdr.so (the C ABI declared in include/gdr.h).

There is no fallback: if the CUDA library cannot be loaded (and cannot be built
because nvcc is absent) importing the rasterizer raises.  The product path never
routes through a CPU implementation.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

_vp = C.c_void_p
_i = C.c_int
_i64 = C.c_int64
_f = C.c_float
_pi64 = C.POINTER(C.c_int64)

GDR_OK = 0
GRAD_MEANS2D, GRAD_MEANS3D, GRAD_COLOR, GRAD_OPACITY, GRAD_COV, GRAD_ALL = 1, 2, 4, 8, 16, 31
GRAD_RAW_PARAMS = 32
GRAD_HWC_COLOR = 64
GRAD_SCRATCH_CLEAN = 128
FLAG_NO_TILE_CULL = 1
FLAG_RAW_PARAMS = 2
FLAG_RERUN = 4
FLAG_FUSED_EPILOGUE = 8
COUNT_RENDERED, COUNT_FLAGS, COUNT_MAX_TILE, COUNT_READY = 0, 1, 2, 3
COUNT_FLAG_PREFILTERED = 2
CAMERA_FLOATS = 48

# name -> (restype, argtypes); must list every symbol include/gdr.h declares
SIGNATURES = {
    "gdr_abi_version": (_i, []),
    "gdr_last_error": (C.c_char_p, []),
    "gdr_geom_state_bytes": (_i, [_i, _pi64]),
    "gdr_image_state_bytes": (_i, [_i, _i, _pi64]),
    "gdr_splat_stream_bytes": (_i, [_i64, _pi64]),
    "gdr_sort_scratch_bytes": (_i, [_i, _i, _i64, _pi64]),
    "gdr_sort_scratch_exact_bytes": (_i, [_i64, _pi64]),
    "gdr_backward_scratch_bytes": (_i, [_i, _pi64]),
    "gdr_tile_offsets": (_i, [_i, _i, _i, _vp, _vp, _vp]),
    "gdr_forward_project": (_i, [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp, _f, _f, _i,
                                 _vp, _vp, _vp, _vp, _i64, _vp, _vp, _i, _vp]),
    "gdr_forward_render": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _i, _vp]),
    "gdr_backward": (_i, [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp, _f, _f, _vp, _vp,
                          _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "gdr_views_forward_project": (_i, [_i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _f, _vp, _vp, _vp, _i, _vp, _vp,
                                       _vp, _vp, _i64, _vp, _vp, _i, _vp]),
    "gdr_views_forward_render": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _i,
                                      _vp]),
    "gdr_views_backward": (_i, [_i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64,
                                _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "gdr_mse_grad": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "gdr_views_densify_scores": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "gdr_topk_select": (_i, [_i, _vp, _i, _vp, _vp, _vp, _vp, _vp]),
    "gdr_mark_visible": (_i, [_i, _vp, _vp, _vp, _vp, _vp]),
    "gdr_surfel_state_bytes": (_i, [_i, _pi64]),
    "gdr_surfel_stream_bytes": (_i, [_i64, _pi64]),
    "gdr_surfel_aux_bytes": (_i, [_i, _i, _pi64]),
    "gdr_surfel_backward_scratch_bytes": (_i, [_i, _pi64]),
    "gdr_surfel_forward_project": (_i, [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _f, _vp, _vp, _vp, _vp, _vp,
                                        _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    "gdr_surfel_forward_render": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _i,
                                       _vp]),
    "gdr_surfel_backward": (_i, [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _f, _vp, _vp, _vp, _vp, _vp, _vp,
                                 _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp,
                                 _vp, _vp, _vp]),
    "gdr_knn3_mean_dist2": (_i, [_i, _vp, _vp, _vp]),
    "gdr_debug_unpack_geom": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "gdr_debug_unpack_bins": (_i, [_i, _i, _vp, _vp, _i64, _vp, _vp, _vp, _vp]),
    "gdr_profile_enable": (_i, [_i]),
    "gdr_profile_read": (_i, [C.POINTER(C.c_double), _pi64]),
}

STAGES = ("project", "tile_sort", "blend_fwd", "blend_bwd", "gauss_bwd")


def profile_enable(on: bool) -> None:
    check(load().gdr_profile_enable(int(on)), "gdr_profile_enable")


def profile_read():
    """{stage: (total_ms, launches)} since the last read; synchronises on the recorded events."""
    ms = (C.c_double * len(STAGES))()
    n = (C.c_int64 * len(STAGES))()
    check(load().gdr_profile_read(ms, n), "gdr_profile_read")
    return {s: (float(ms[i]), int(n[i])) for i, s in enumerate(STAGES)}

_lib = None


class GdrError(RuntimeError):
    pass


def lib_path() -> str:
    return _build.LIB


def load():
    """Load (building in-tree first if the sources are newer) and type the library."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("GDR_LIB")  # experiments: a pre-built variant of the library, loaded as it is
    if path:
        if not os.path.isfile(path):
            raise GdrError(f"GDR_LIB={path} does not exist")
    else:
        path = _build.LIB
    if path == _build.LIB and (not os.path.isfile(path) or (os.path.isfile(_build.NVCC) and not _build.is_fresh())):
        try:
            _build.build()
        except Exception as e:  # no nvcc, or a compile error: fail loudly
            if not os.path.isfile(path):
                raise GdrError(
                    "libgdr.so (the CUDA rasterizer) is not built and could not be built: "
                    f"{e}. Run `python -m generativedensification_b200.build`.") from e
            raise
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.gdr_abi_version() != 2:
        raise GdrError(f"libgdr.so ABI version {lib.gdr_abi_version()} != 2")
    _lib = lib
    return lib


def check(status: int, what: str) -> None:
    if status != GDR_OK:
        msg = load().gdr_last_error()
        raise GdrError(f"{what} failed ({status}): {msg.decode() if msg else ''}")


_query_cache = {}


def query_bytes(fn_name: str, *args) -> int:
    """Size of a caller-owned buffer (pure host functions of their arguments: memoised, the drop-in path asks for
    the same five sizes on every call)."""
    key = (fn_name, args)
    v = _query_cache.get(key)
    if v is None:
        out = C.c_int64(0)
        check(getattr(load(), fn_name)(*args, C.byref(out)), fn_name)
        v = int(out.value)
        if len(_query_cache) > 4096:
            _query_cache.clear()
        _query_cache[key] = v
    return v
