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


def test_host_side_layout_helpers():
    """Entry points that only compute sizes / supported shapes run on the host (no GPU, no launch)."""
    lib = _lib.load()
    # packed weight images: n_blocks x nchunk x {hi, lo} x Npad rows x 128 B
    assert lib.gnb_linear_tc_packed_bytes(64, 6) == 1 * 1 * 2 * 64 * 128
    assert lib.gnb_linear_tc_packed_bytes(137, 137) == 1 * 3 * 2 * 160 * 128          # 137 columns pad to 160
    assert lib.gnb_linear_tc_packed_bytes(256, 1280) == 2 * 20 * 2 * 128 * 128          # long K: two 128-column blocks
    assert lib.gnb_linear_tc_padded_cols(1024, 512) == 1024 and lib.gnb_linear_tc_padded_cols(137, 137) == 160
    assert lib.gnb_linear_tc_packed_bytes(0, 5) == 0
    # tensor-core convolution: power-of-two spatial sizes, Cout multiple of 32 up to 128, any batch that fills whole tiles
    assert lib.gnb_conv3d_tc_supported(32, 32, 32, 32, 128, 128) == 1
    assert lib.gnb_conv3d_tc_supported(5, 8, 8, 8, 64, 64) == 1            # 512 voxels per sample: any batch size
    assert lib.gnb_conv3d_tc_supported(5, 4, 4, 4, 128, 128) == 0          # 64 voxels per sample: batch must be even
    assert lib.gnb_conv3d_tc_supported(6, 4, 4, 4, 128, 128) == 1
    assert lib.gnb_conv3d_tc_supported(1, 8, 8, 8, 64, 256) == 0           # Cout 256 runs on the fp32 kernel
    assert lib.gnb_conv3d_tc_supported(1, 12, 8, 8, 64, 64) == 0
    assert lib.gnb_conv3d_tc_dx_supported(32, 32, 32, 32, 128, 32) == 1 and lib.gnb_conv3d_tc_dx_supported(32, 32, 32, 32, 128, 128) == 0
    assert lib.gnb_conv3d_tc_dx_supported(2, 8, 8, 8, 64, 64) == 1 and lib.gnb_conv3d_tc_dx_supported(2, 8, 8, 4, 64, 64) == 0
    # workspaces grow with the problem and cover the scans' tile arrays
    small, big = lib.gnb_mesh_cleanup_workspace_bytes(10, 10), lib.gnb_mesh_cleanup_workspace_bytes(6000000, 12000000)
    assert 0 < small < big and big > 6000000 * 12 + 12000000 * 12
    assert lib.gnb_mc_workspace_bytes(128, 128, 128) > 128 ** 3 * 2
