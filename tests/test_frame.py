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


def _expected_ascii(pts, variables):
    """The legacy-VTK text of visit_writer.cpp:673-719 restated in Python (every number "%20.12e ", nine per line)."""
    def block(a):
        a = np.asarray(a, np.float32).reshape(-1)
        out, col = [], 0
        for v in a:
            out.append("%20.12e " % float(v))
            col += 1
            if col == 9:
                out.append("\n")
                col = 0
        return "".join(out), col
    n = pts.shape[0]
    t = "# vtk DataFile Version 2.0\nWritten using VisIt writer\nASCII\nDATASET UNSTRUCTURED_GRID\n"
    b, col = block(pts)
    t += f"POINTS {n} float\n" + b + ("\n" if col else "")
    t += f"CELLS {n} {2 * n}\n" + "".join(f"1 {i} \n" for i in range(n))
    t += f"CELL_TYPES {n}\n" + "1 \n" * n
    t += f"CELL_DATA {n}\nPOINT_DATA {n}\n"
    names = list(variables)
    b, col = block(variables[names[0]])
    t += f"SCALARS {names[0]} float\nLOOKUP_TABLE default\n" + b + "\n"
    t += f"FIELD FieldData {len(names) - 1}\n"
    for k in names[1:]:
        b, col = block(variables[k])
        t += f"{k} 1 {n} float\n" + b + "\n"
    return (t + "\n").encode()


@pytest.mark.parametrize("n", [7, 9, 40000, 150001])
def test_large_frames_formatted_by_several_threads_have_the_same_bytes(fsg, tmp_path, n):
    """Arrays above 32 k numbers are formatted by several threads, each at the file offset the fixed 21-byte number width gives it
    (fsg_frame.cu put_floats): the file must be what the one-number-at-a-time writer produces, whatever the line phase (n % 9)."""
    rng = np.random.default_rng(n)
    pts = (rng.standard_normal((n, 3)) * 10.0 ** rng.integers(-30, 30, (n, 1))).astype(np.float32)
    variables = {"dens": rng.uniform(9000, 9900, n).astype(np.float32), "cellnumber": rng.integers(0, 1 << 24, n).astype(np.float32),
                 "extra": (-rng.uniform(0, 1, n)).astype(np.float32)}
    out = tmp_path / "big.vtk"
    fsg.write_point_mesh(out, pts, variables)
    assert out.read_bytes() == _expected_ascii(pts, variables)


@pytest.mark.gpu
def test_asynchronous_frames_are_the_synchronous_frames(fsg, tmp_path):
    """fsg_write_frame_async: export kernel on the solver's stream, copy on a second stream, formatting + I/O on a writer thread, two
    staging slots with back-pressure.  The solver keeps stepping while frames are written; every file must equal the synchronous
    fsg_write_frame of the same step."""
    cfg = fsg.scenes.plume_config(32)
    cfg.pair_mode = 1                               # the deterministic pair kernel: two contexts give the same bits
    state = fsg.scenes.plume_scene(cfg)
    cfg.capacity = state["pos"].shape[0]
    with fsg.FluidSolver(cfg) as a, fsg.FluidSolver(cfg) as b:
        a.upload(state)
        b.upload(state)
        for k in range(1, 7):                       # six frames through two slots: the third call onwards waits for a free slot
            a.step(2)
            b.step(2)
            a.write_frame_async(tmp_path / f"async_{k}.vtk", binary=(k % 2 == 0))
            b.write_frame(tmp_path / f"sync_{k}.vtk", binary=(k % 2 == 0))
        a.step(3)                                   # the solver is not held back by the writer
        assert a.frame_wait() == 6
        for k in range(1, 7):
            assert (tmp_path / f"async_{k}.vtk").read_bytes() == (tmp_path / f"sync_{k}.vtk").read_bytes(), k
        with pytest.raises(fsg.FsgError):           # an unwritable path surfaces at the next wait
            a.write_frame_async(tmp_path / "no_such_dir" / "f.vtk")
            a.frame_wait()
        assert a.frame_wait() == 6                  # the error was reported once and cleared
