"""CPU-side checks of the drop-in boundary: the library loads and exports every declared symbol."""
import ctypes as C
import os
import re

from conftest import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "plmatch.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(plm_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_path():
    syms = _declared_symbols()
    for must in ("plm_hamming256", "plm_knn2", "plm_match_nnr", "plm_match", "plm_match_grid_points",
                 "plm_match_grid_lines", "plm_stereo_filter_points", "plm_stereo_filter_lines", "plm_db_knn2",
                 "plm_dev_top2_merge", "plm_batch_run"):
        assert must in syms


def test_library_exports_every_declared_symbol(plm_lib):
    from pl_inertial_slam_b200 import _lib
    syms = _declared_symbols()
    assert syms, "no symbols parsed from include/plmatch.h"
    raw = C.CDLL(_lib.LIB_PATH)
    missing = [s for s in syms if not hasattr(raw, s)]
    assert not missing, missing
    # the Python binding covers the same set
    assert sorted(_lib.SIGNATURES) == syms


def test_status_strings_and_version(plm_lib):
    assert plm_lib.plm_version() == 100
    assert plm_lib.plm_status_string(0) == b"ok"
    assert b"GridStructure" in plm_lib.plm_status_string(-4)


def test_argument_validation_needs_no_gpu(plm_lib):
    """Precondition failures are reported before any CUDA call."""
    import numpy as np
    from pl_inertial_slam_b200 import _lib as L
    d = np.zeros((4, 32), np.uint8)
    m = np.full(4, -1, np.int32)
    n = C.c_int(0)
    p = d.ctypes.data_as(L.u8p)
    # empty train set with a non-empty query: the reference throws (matching.cpp:50-51); an empty query matches nothing
    assert plm_lib.plm_match_nnr(None, p, 4, 32, p, 0, 32, C.c_float(0.9), m.ctypes.data_as(L.i32p), C.byref(n)) == L.PLM_E_TRAIN
    assert plm_lib.plm_match(None, p, 4, 32, p, 0, 32, C.c_float(0.9), 1, m.ctypes.data_as(L.i32p), C.byref(n)) == L.PLM_E_TRAIN
    n.value = 7
    assert plm_lib.plm_match(None, p, 0, 32, p, 4, 32, C.c_float(0.9), 1, m.ctypes.data_as(L.i32p), C.byref(n)) == L.PLM_OK
    assert n.value == 0
    # step smaller than a descriptor
    assert plm_lib.plm_match_nnr(None, p, 4, 16, p, 4, 32, C.c_float(0.9), m.ctypes.data_as(L.i32p), C.byref(n)) == L.PLM_E_INVALID
    # invalid grid dimension / ratio > 1
    cs = np.zeros(2, np.int32)
    win = np.zeros(4, np.int32)
    xy = np.zeros((4, 2), np.int32)
    args = (xy.ctypes.data_as(L.i32p), p, 4, 32, cs.ctypes.data_as(L.i32p), cs.ctypes.data_as(L.i32p))
    assert plm_lib.plm_match_grid_points(None, *args, 0, 1, p, 4, 32, win.ctypes.data_as(L.i32p), 0.9, 1,
                                         m.ctypes.data_as(L.i32p), C.byref(n)) == L.PLM_E_GRID
    assert plm_lib.plm_match_grid_points(None, *args, 1, 1, p, 4, 32, win.ctypes.data_as(L.i32p), 1.5, 1,
                                         m.ctypes.data_as(L.i32p), C.byref(n)) == L.PLM_E_RATIO
    assert b"ratio" in plm_lib.plm_last_error()


def test_host_mirror_raises_reference_messages(plm_lib):
    import numpy as np
    import pytest
    from pl_inertial_slam_b200 import matching as M
    from pl_inertial_slam_b200.grid import GridStructure, GridWindow
    d = np.zeros((4, 32), np.uint8)
    with pytest.raises(RuntimeError, match=r"\[matchNNR\] Different size"):
        M.matchNNR(d, d[:0], 0.9, [])
    with pytest.raises(RuntimeError, match=r"\[GridStructure\] invalid dimension"):
        GridStructure(0, 4)
    g = GridStructure(4, 4)
    with pytest.raises(RuntimeError, match=r"\[matchGrid\] Each point needs"):
        M.matchGrid(np.zeros((3, 2), np.int32), d, g, d, GridWindow(), [])
    with pytest.raises(RuntimeError, match=r"\[matchGrid\] Each line needs"):
        M.matchGrid(np.zeros((3, 4), np.int32), d, g, d, np.zeros((4, 2)), GridWindow(), [])


def test_product_package_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under pl_inertial_slam_b200/ may import or load it."""
    pkg = os.path.join(ROOT, "pl_inertial_slam_b200")
    offenders = []
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inl", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                if re.search(r"^\s*(import oracle|from oracle)\b", text, flags=re.M) or "libploracle" in text or "libplref" in text:
                    offenders.append(f)
    assert not offenders, offenders


def test_ctypes_struct_layouts_match_the_header(tmp_path):
    """Every struct of include/plmatch.h that the Python binding mirrors has the same size and field offsets when
    compiled by the C compiler."""
    import ctypes as C
    import subprocess
    from pl_inertial_slam_b200 import _lib as L
    pairs = [("plm_pair_job", L.PairJob), ("plm_grid_job", L.GridJob), ("plm_dev_grid_args", L.DevGridArgs),
             ("plm_frame_rec", L.FrameRec), ("plm_frame_config", L.FrameConfig), ("plm_frames_out", L.FramesOut),
             ("plm_peer_group", L.PeerGroup), ("plm_map_view", L.MapView)]
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "plmatch.h"', 'int main(void) {']
    for cname, cls in pairs:
        lines.append(f'  printf("{cname} %zu", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf(" %zu", offsetof({cname}, {fname}));')
        lines.append('  printf("\\n");')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.strip().splitlines()
    for (cname, cls), line in zip(pairs, out):
        nums = [int(x) for x in line.split()[1:]]
        assert nums[0] == C.sizeof(cls), cname
        assert nums[1:] == [getattr(cls, f).offset for f, _ in cls._fields_], cname
