"""CPU tests of the C-ABI library: it loads, exports every symbol include/vxp.h declares,
the ctypes table matches the header, host-only helpers work, and compute entry points fail
loudly without a CUDA device (there is no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "vxp.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vxp_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    import voxelphile_b200._lib as L
    if not os.path.exists(L.LIB_PATH):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "voxelphile_b200", "csrc")],
                              stdout=subprocess.DEVNULL)
    return L.load()


def test_every_declared_symbol_is_exported_and_bound(lib):
    import voxelphile_b200._lib as L
    names = declared_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in vxp.h but not exported"
    assert set(names) == set(L.SIGNATURES), set(names) ^ set(L.SIGNATURES)


def test_library_is_sm100a_only():
    import voxelphile_b200._lib as L
    out = subprocess.run(["cuobjdump", "-lelf", L.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_status_strings_and_version(lib):
    assert lib.vxp_spec_version() == 1
    assert lib.vxp_status_str(0) == b"ok"
    assert b"no CPU fallback" in lib.vxp_status_str(2)


def test_nbr6_from_coords(lib):
    from voxelphile_b200 import grid_coords, neighbour_table, MeshError
    coords = grid_coords((0, 0, 0), (3, 2, 2))
    nbr = neighbour_table(coords)
    assert nbr.shape == (12, 6)
    assert nbr[0].tolist() == [-1, 1, -1, 3, -1, 6]
    assert nbr[4].tolist() == [3, 5, 1, -1, -1, 10]
    # negative coordinates and a sparse set
    sparse = np.array([[-5, 0, 0], [-4, 0, 0], [7, 7, 7]], dtype=np.int32)
    assert neighbour_table(sparse).tolist() == [[-1, 1, -1, -1, -1, -1], [0, -1, -1, -1, -1, -1], [-1] * 6]
    with pytest.raises(MeshError) as e:
        neighbour_table(np.array([[0, 0, 0], [0, 0, 0]], dtype=np.int32))
    assert e.value.kind == "BadArg"
    assert neighbour_table(np.zeros((0, 3), dtype=np.int32)).shape == (0, 6)


def test_no_device_is_an_error_not_a_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from voxelphile_b200 import Mesher, MeshError
    with pytest.raises(MeshError) as e:
        Mesher(0)
    assert e.value.kind == "NoDevice"
    h = C.c_void_p()
    assert lib.vxp_ctx_create(0, C.byref(h)) == 2 and not h.value
    # NULL context is rejected by every compute entry point
    assert lib.vxp_mesh_batch(None, None, None, 0, 0, None) == 1
    assert lib.vxp_terrain_fill_batch(None, None, 0, 0, None) == 1


def test_product_does_not_reference_the_oracle():
    """The shipped package and csrc never import, link or call anything under oracle/."""
    pkg = os.path.join(ROOT, "voxelphile_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "vxo_" not in text and "vxp_oracle" not in text and "oracle_lib" not in text, f
    import voxelphile_b200._lib as L
    deps = subprocess.run(["ldd", L.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in deps
