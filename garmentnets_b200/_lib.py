"""ctypes binding of the C-ABI declared in ``include/garmentnets_b200.h``.

The prototypes are parsed from the header itself, so the Python argtypes can never drift from the C
declarations.  There is NO fallback: if the shared library is missing (not built) every call raises
``GarmentNetsB200Error`` -- the product path never routes around the CUDA extension.
"""
from __future__ import annotations

import ctypes
import os
import re
from typing import Dict, List, Tuple

_ROOT = os.path.dirname(os.path.abspath(__file__))
HEADER_PATH = os.path.join(os.path.dirname(_ROOT), "include", "garmentnets_b200.h")
# GNB_B200_LIBRARY (read once, at import): development override used by tools/ to load the profiling build
# (`make PROFILE_KNOBS=1` -> lib/libgarmentnets_b200_prof.so); never a CPU fallback -- it must be another build of this library
LIB_PATH = os.environ.get("GNB_B200_LIBRARY") or os.path.join(_ROOT, "lib", "libgarmentnets_b200.so")


class GarmentNetsB200Error(RuntimeError):
    """Raised for every non-zero status of the native library (and when it cannot be loaded)."""


_CTYPES = {
    "int32_t": ctypes.c_int32,
    "int64_t": ctypes.c_int64,
    "float": ctypes.c_float,
    "double": ctypes.c_double,
}


def parse_header(path: str = HEADER_PATH) -> Dict[str, Tuple[object, List[object]]]:
    """Return {name: (restype, [argtypes])} for every ``gnb_*`` prototype in the header."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    protos: Dict[str, Tuple[object, List[object]]] = {}
    for m in re.finditer(r"(const\s+char\s*\*|int32_t|int64_t)\s+(gnb_\w+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        restype = ctypes.c_char_p if "char" in ret else _CTYPES[ret]
        argtypes: List[object] = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                if "*" in a:
                    argtypes.append(ctypes.c_void_p)
                else:
                    base = a.replace("const ", "").split(" ")[0]
                    argtypes.append(_CTYPES[base])
        protos[name] = (restype, argtypes)
    return protos


_lib = None
_protos = None


def load() -> ctypes.CDLL:
    """Load the shared library once; fail loudly when it is absent."""
    global _lib, _protos
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GarmentNetsB200Error(
            f"native library not built: {LIB_PATH} (run `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C garmentnets_b200/csrc`); there is no CPU fallback")
    lib = ctypes.CDLL(LIB_PATH)
    _protos = parse_header()
    for name, (restype, argtypes) in _protos.items():
        fn = getattr(lib, name)  # AttributeError if the header declares something the library lacks
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def last_error() -> str:
    return load().gnb_last_error().decode("utf-8", "replace")


# kernels launched per entry point (for the launch counter bench.py reports); entry points not listed launch one
KERNELS_PER_CALL = {"gnb_scatter_reduce": 2, "gnb_gaussian_gradient_magnitude": 1,
                    "gnb_gaussian_gradient_magnitude_batched": 1, "gnb_conv3d_tc_supported": 0, "gnb_conv3d_tc_dx_supported": 0, "gnb_mc_totals_offset": 0, "gnb_decode_lattice_set_mode": 0, "gnb_decode_query_set_mode": 0, "gnb_mc_cell_tiling_host": 0,
                    "gnb_mc_count": 5, "gnb_mc_emit": 3, "gnb_mc_count_batch": 5, "gnb_mc_emit_batch": 3,
                    "gnb_groupnorm_stats": 2, "gnb_groupnorm_stats_cat": 3, "gnb_decode_tc": 2, "gnb_decode_tc_query": 2, "gnb_decode_tc_query_fused": 3, "gnb_decode_lattice": 2, "gnb_version": 0, "gnb_last_error": 0, "gnb_device_sm_count": 0,
                    "gnb_mc_workspace_bytes": 0, "gnb_f16_overflow_fetch": 0, "gnb_f16_overflow_fetch_async": 1, "gnb_copy_to_pinned_host": 1, "gnb_conv_tc_set_cross_precision": 0, "gnb_conv_tc_cross_precision": 0, "gnb_pointconv_mlp_supported": 0, "gnb_pointconv_mlp_packed_bytes": 0, "gnb_pointconv_mlp_pack": 3, "gnb_pointconv_mlp_max": 3, "gnb_linear_tc_packed_bytes": 0,
                    "gnb_linear_tc_padded_cols": 0, "gnb_mesh_cleanup_workspace_bytes": 0, "gnb_mesh_cleanup_count": 8,
                    "gnb_mesh_cleanup_emit": 2, "gnb_linear_tc_segmax": 1, "gnb_mesh_components": 4,
                    "gnb_mesh_sample_barycentric": 5}
launch_count = 0          # kernels launched through this binding since import (monotonic)
_tag = None               # current profiling tag (see garmentnets_b200.profiling)
_tag_sink = None          # callable(tag, name) -> context manager, installed by profiling.KernelTimer


def call(name: str, *args):
    """Invoke ``name`` and raise on a negative status.  Pointers are passed as ints (``tensor.data_ptr()``)."""
    global launch_count
    lib = load()
    fn = getattr(lib, name)
    launch_count += KERNELS_PER_CALL.get(name, 1)
    if _tag_sink is not None and _tag is not None:
        with _tag_sink(_tag, name):
            status = fn(*args)
    else:
        status = fn(*args)
    if isinstance(status, int) and status < 0:
        msg = last_error()
        if status == -4:
            raise ValueError(f"{name}: {msg}")  # mirrors skimage's ValueError (ref predict.py:188)
        raise GarmentNetsB200Error(f"{name} failed with status {status}: {msg}")
    return status
