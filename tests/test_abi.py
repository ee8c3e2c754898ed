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


def test_exports_are_exactly_the_header(lib):
    """No undeclared vxp_* entry point ships: the dynamic symbols of libvxp.so are the functions of include/vxp.h
    (debug / profiling exports exist only in the `make prof` / `make variant` builds)."""
    import voxelphile_b200._lib as L
    out = subprocess.run(["nm", "-D", "--defined-only", L.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = sorted(line.split()[-1] for line in out.splitlines() if line.split()[-1].startswith("vxp_"))
    assert exported == declared_functions(), set(exported) ^ set(declared_functions())


def test_library_has_no_runtime_switches():
    """No environment variable changes what the shipped library runs (no multi-path dispatch): tuning variants are
    separate `make variant` builds."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "voxelphile_b200", "csrc")):
        for f in files:
            assert "getenv" not in open(os.path.join(dirpath, f), errors="ignore").read(), f


def build_c_consumer(tmpdir):
    """tests/c_consumer/consumer.c compiled as strict C11 against include/vxp.h and linked (not dlopen'd) with libvxp.so."""
    import voxelphile_b200._lib as L
    exe = os.path.join(str(tmpdir), "vxp_consumer")
    libdir = os.path.dirname(L.LIB_PATH)
    link = os.path.join(str(tmpdir), "libvxp.so")
    if not os.path.exists(link):
        os.symlink(L.LIB_PATH, link)
    subprocess.run(["gcc", "-std=c11", "-pedantic", "-Wall", "-Wextra", "-Werror", "-O1", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "c_consumer", "consumer.c"), "-L", str(tmpdir), "-lvxp",
                    f"-Wl,-rpath,{libdir}", "-o", exe], check=True, capture_output=True, text=True)
    return exe


def test_c11_consumer_compiles_links_and_fails_loudly_without_a_device(lib, tmp_path):
    import torch
    exe = build_c_consumer(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present: the run is checked by the gpu test")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 2 and "no CUDA device" in r.stderr, (r.returncode, r.stderr)


@pytest.mark.gpu
def test_c11_consumer_matches_the_oracle(lib, tmp_path):
    """The linked C program's meshes (stored and fused, both modes) against the oracle on the same coordinates."""
    import oracle_lib as O
    exe = build_c_consumer(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    coords = np.array([[0, 0, 0], [0, 8, 0], [1, 8, 0]], dtype=np.int32)
    from voxelphile_b200 import neighbour_table
    nbr = neighbour_table(coords)
    vox = O.terrain_fill_batch(42, coords)
    want = {}
    for mode in (0, 1):
        h, c = O.mesh_batch_hash(vox, nbr, mode)
        for i in range(3):
            want[("stored", mode, i)] = (int(c[i]), int(h[i]))
            lo = tuple(int(v) for v in coords[i])
            fh, fc = O.terrain_world_hash(42, lo, tuple(v + 1 for v in lo), mode)
            want[("fused", mode, i)] = (int(fc[0]), int(fh[0]))
    got = {}
    for line in r.stdout.split("\n"):
        if line.strip():
            call, mode, i, cnt, h = line.split()
            got[(call, int(mode), int(i))] = (int(cnt), int(h))
    assert got == want
    assert sum(c for c, _ in got.values()) > 0


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


def test_rust_crates_and_integration_guide_agree_with_the_header():
    """The Rust crates cannot be compiled here (no toolchain), so consistency of the text is the check there is:
    vxp-sys declares exactly the functions of include/vxp.h with the same number of parameters, and every
    `mesher.<method>(` the integration guide shows exists in the safe wrapper."""
    hdr = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    c_decl = {m.group(1): m.group(2) for m in re.finditer(r"\b(vxp_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.S)}
    sys_src = open(os.path.join(ROOT, "rust", "vxp-sys", "src", "lib.rs")).read()
    r_decl = {m.group(1): m.group(2) for m in re.finditer(r"pub fn (vxp_[a-z0-9_]+)\s*\((.*?)\)\s*(?:->[^;]*)?;", sys_src, flags=re.S)}
    assert set(r_decl) == set(c_decl) == set(declared_functions()), set(r_decl) ^ set(c_decl)

    def arity(params):
        params = params.strip()
        return 0 if params in ("", "void") else params.count(",") + 1
    for name in c_decl:
        assert arity(c_decl[name]) == arity(r_decl[name]), name
    wrapper = open(os.path.join(ROOT, "rust", "vxp", "src", "lib.rs")).read()
    methods = set(re.findall(r"pub fn ([a-z_]+)\s*[(<]", wrapper))
    guide = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    used = set(re.findall(r"\bmesher\.([a-z_]+)\(", guide)) | set(re.findall(r"\bMesher::([a-z_]+)\(", guide))
    assert used and used <= methods, used - methods
