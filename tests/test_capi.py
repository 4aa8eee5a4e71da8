"""C-ABI surface: the library loads and exports every symbol the header declares (no compute, CPU-only)."""
import ctypes
import os
import subprocess

from garmentnets_b200 import _lib


def test_header_parses_all_entry_points():
    protos = _lib.parse_header()
    for name in ["gnb_fps", "gnb_ball_query", "gnb_knn", "gnb_knn_interpolate", "gnb_linear", "gnb_scatter_reduce",
                 "gnb_conv3d_k3", "gnb_groupnorm_stats", "gnb_trilinear_sample", "gnb_gaussian_gradient_magnitude",
                 "gnb_mc_count", "gnb_mc_emit", "gnb_version", "gnb_last_error"]:
        assert name in protos, name
    # pointer / scalar classification
    restype, argtypes = protos["gnb_ball_query"]
    assert restype is ctypes.c_int32
    assert argtypes[:4] == [ctypes.c_void_p] * 4 and argtypes[6] is ctypes.c_double and argtypes[7] is ctypes.c_int32


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"
    lib = _lib.load()
    for name in _lib.parse_header():
        assert hasattr(lib, name), f"{name} declared in include/garmentnets_b200.h but not exported"
    assert lib.gnb_version() >= 100
    assert lib.gnb_last_error() is not None


def test_no_undeclared_exports():
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l and l.split()[-1].startswith("gnb_")}
    assert exported == set(_lib.parse_header()), exported ^ set(_lib.parse_header())


def test_product_does_not_import_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dirpath, _, files in os.walk(os.path.join(root, "garmentnets_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, os.path.join(dirpath, f)
