"""Regenerates tests/golden/kat_base.json from the reference's own host-compiled functions.

Needs /root/reference (build container only): `make -C oracle _ref/ref_kat` compiles the
UNMODIFIED /root/reference/FluidGPU.cu and links oracle/ref_kat.cu against it.
"""
import pathlib
import subprocess

root = pathlib.Path(__file__).resolve().parents[2]
subprocess.check_call(["make", "-C", str(root / "oracle"), "_ref/ref_kat"])
out = subprocess.check_output([str(root / "oracle/_ref/ref_kat"), "257"])
(root / "tests/golden/kat_base.json").write_bytes(out)
print("wrote", len(out), "bytes")
