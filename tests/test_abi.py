"""The C-ABI library loads and exports every symbol include/instance_stixels_b200.h declares
(no compute calls: this runs without a GPU)."""
import ctypes as C
import os
import re

import pytest

from instance_stixels_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "instance_stixels_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(isx_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = L.load()
    declared = header_functions()
    assert len(declared) >= 35
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(L.EXPORTS) == declared


def test_abi_version_and_struct_sizes():
    lib = L.load()
    assert lib.isx_abi_version() == 2
    assert L.PACKED_FRAME_DTYPE.itemsize == 32  # isx_packed_frame
    assert C.sizeof(L.Config) == 41 * 4
    assert L.SECTION_DTYPE.itemsize == 32  # Section, types.h:186-194
    assert C.sizeof(L.FrameMeta) == 36 and C.sizeof(L.Road) == 16


def test_config_defaults_match_reference_struct():
    """isx_config_init == default member initialisers of StixelConfig (types.h:30-141)."""
    lib = L.load()
    c = L.Config()
    lib.isx_config_init(C.byref(c))
    for n in ("rows", "cols", "max_dis", "eps", "min_pts", "size_filter", "n_semantic_classes",
              "n_offset_channels", "prior_weight", "segmentation_weight", "instance_weight",
              "disparity_weight", "column_step", "focal", "baseline", "camera_center_x", "camera_center_y"):
        assert getattr(c, n) == -1, n
    assert c.invalid_disparity == -1.0 and c.pairwise == 0 and c.median_join == 0 and c.width_margin == 0
    f32 = lambda x: C.c_float(x).value
    expect = dict(sigma_disparity_object=1.0, sigma_disparity_ground=2.0, sigma_sky=0.1, pout=0.15, pout_sky=0.4,
                  pord=0.2, pgrav=0.1, pblg=0.04, pground_given_nexist=0.28, pobject_given_nexist=0.44,
                  psky_given_nexist=0.28, pnexist_dis=0.25, sigma_camera_tilt=0.05, sigma_camera_height=0.05,
                  epsilon=3.0, range_objects_z=10.20, road_vdisparity_threshold=0.2)
    for k, v in expect.items():
        assert getattr(c, k) == f32(v), k
    third = f32(f32(1.0) / f32(3.0))
    assert c.pground == third and c.pobject == third and c.psky == third
    # the oracle's python-side defaults are the same numbers
    from oracle import cpubind
    d = cpubind.default_config()
    for n, _ in L.Config._fields_:
        assert getattr(c, n) == getattr(d, n), n


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    lib = L.load()
    h = C.c_void_p()
    rc = lib.isx_create(C.byref(h), 0)
    assert rc == -3  # ISX_ERR_CUDA: fails loudly, nothing is computed on the host
    assert b"no CPU fallback" in lib.isx_last_error(None)
