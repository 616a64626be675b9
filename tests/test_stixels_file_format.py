"""The `.stixels` text format: files written by the drop-in Stixels::SaveStixels
(Stixels.cu:889-926) are read back with the REFERENCE's own reader, `read_stixel_file`
(tools/visualization/clustering_visualization.py:73-112), taken from the reference tree when it is present (this
container) -- the consumer defines the format.  CPU only: the Sections come from the CPU oracle."""
import ast
import os
import subprocess

import numpy as np
import pytest

from instance_stixels_b200 import _lib as L, synth
from oracle import cpubind

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_READER = "/root/reference/tools/visualization/clustering_visualization.py"
EXE = os.path.join(ROOT, "tests", "cpp", "save_stixels_check")


def reference_reader():
    """`read_stixel_file` alone (the module itself imports h5py, matplotlib, ...)."""
    tree = ast.parse(open(REF_READER).read())
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "read_stixel_file"][0]
    ns = {}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), REF_READER, "exec"), ns)
    return ns["read_stixel_file"]


@pytest.mark.skipif(not os.path.exists(REF_READER), reason="reference tree not present")
@pytest.mark.parametrize("mode", ["unary", "pairwise"])
def test_reference_reader_reads_our_files(tmp_path, mode):
    libdir = os.path.join(ROOT, "instance_stixels_b200")
    subprocess.run(["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include", "InstanceStixels"),
                    os.path.join(ROOT, "tests", "cpp", "save_stixels_check.cpp"), "-L" + libdir,
                    "-linstance_stixels_b200", "-Wl,-rpath," + libdir, "-o", EXE], check=True, capture_output=True)
    rows, cols = 128, 256
    pre = synth.preset(mode, rows, cols, 8)
    fr = synth.make_frame(4, rows=rows, cols=cols)
    sec, inst, _ = cpubind.compute_frame(cpubind.default_config(**pre), mode == "pairwise", fr.disparity,
                                         fr.segmentation, fr.road)
    assert len(inst) > 0
    sec.tofile(tmp_path / "sections.bin")
    np.stack([inst["column"], inst["index"], inst["label"]], axis=1).astype(np.int32).tofile(tmp_path / "inst.i32")
    out = tmp_path / "frame.stixels"
    vhor = rows - 1 - fr.road["vhor"]
    subprocess.run([EXE, str(tmp_path / "sections.bin"), str(cols // 8), "200", str(tmp_path / "inst.i32"),
                    repr(fr.road["alpha_ground"]), str(vhor), str(out)], check=True)
    stixels, groundplane = reference_reader()(str(out))
    assert len(stixels) == cols // 8
    assert groundplane[1] == vhor and np.isclose(groundplane[0], fr.road["alpha_ground"], rtol=1e-5)
    labels = {(int(r["column"]), int(r["index"])): int(r["label"]) for r in inst}
    n_labelled = 0
    for c, column in enumerate(stixels):
        n = int(np.argmax(sec[c]["type"] == -1))
        assert len(column) == n
        for j, s in enumerate(column):
            w = sec[c, j]
            assert (s["type"], s["vB"], s["vT"], s["class"]) == (w["type"], w["vB"], w["vT"], w["semantic_class"])
            for key, field in (("disparity", "disparity"), ("cost", "cost"), ("instance_mean_x", "instance_meanx"),
                               ("instance_mean_y", "instance_meany")):
                assert np.isclose(s[key], w[field], rtol=1e-5, atol=1e-6), (c, j, key)   # 6 significant digits
            if (c, j) in labels:
                lab = labels[(c, j)]   # "convert to cityscapes style" (:104-111)
                assert s["instance_label"] == (lab + w["semantic_class"] * 1000 if 0 <= lab < 1000 else -1)
                n_labelled += 1
            else:
                assert "instance_label" not in s
    assert n_labelled == len(inst)
