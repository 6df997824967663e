"""Frame output: fsg_write_point_mesh against golden files written by the reference's
UNMODIFIED VisIt writer (tests/golden/make_vtk.py).  Host code: no GPU needed."""
import importlib.util
import pathlib

import numpy as np
import pytest

GOLD = pathlib.Path(__file__).parent / "golden"
spec = importlib.util.spec_from_file_location("make_vtk", GOLD / "make_vtk.py")
make_vtk = importlib.util.module_from_spec(spec)
spec.loader.exec_module(make_vtk)


@pytest.mark.parametrize("binary", [False, True])
def test_frames_are_byte_identical_to_the_reference_writer(fsg, tmp_path, binary):
    for name, pts, variables in make_vtk.cases():
        gold = GOLD / f"vtk_{name}_{'bin' if binary else 'ascii'}.vtk"
        out = tmp_path / f"{name}.vtk"
        fsg.write_point_mesh(out, pts, variables, binary=binary)
        assert out.read_bytes() == gold.read_bytes(), (name, binary)


def test_filename_extension_and_errors(fsg, tmp_path):
    pts = np.zeros((2, 3), np.float32)
    fsg.write_point_mesh(tmp_path / "noext", pts, {})                   # ".vtk" is appended (visit_writer.cpp:136-143)
    assert (tmp_path / "noext.vtk").exists()
    with pytest.raises(fsg.FsgError):                                    # the reference crashes here (no NULL check, :145)
        fsg.write_point_mesh(tmp_path / "missing_dir" / "f.vtk", pts, {})


@pytest.mark.gpu
def test_write_frame_after_a_step(fsg, tmp_path):
    cfg = fsg.FluidSolver.unidyn_config()
    with fsg.FluidSolver(cfg) as s:
        s.upload(fsg.scenes.unidyn_default_scene())
        s.step(2)
        s.write_frame(tmp_path / "anim_s_GPU0_2.vtk")
        spts, a3, b3 = s.export_viz()
    ref = tmp_path / "ref.vtk"
    fsg.write_point_mesh(ref, spts.reshape(-1, 3), {"mass": a3, "surface_level": b3})
    text = (tmp_path / "anim_s_GPU0_2.vtk").read_bytes()
    assert text == ref.read_bytes() and text.startswith(b"# vtk DataFile Version 2.0\nWritten using VisIt writer\nASCII\n")
    assert b"POINTS 14040 float" in text and b"SCALARS mass float" in text and b"surface_level 1 14040 float" in text
