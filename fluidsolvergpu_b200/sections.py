"""Section files: the tiny container format oracle/ref_harness_*.cu reads and writes
(char name[16]; int32 dtype (0=f32, 1=i32, 2=u8); int64 count; payload)."""
from __future__ import annotations

import struct

import numpy as np

_DT = {0: np.float32, 1: np.int32, 2: np.uint8}
_CODE = {np.dtype(np.float32): 0, np.dtype(np.int32): 1, np.dtype(np.uint8): 2}


def read_sections(path) -> dict:
    out = {}
    with open(path, "rb") as f:
        while True:
            head = f.read(28)
            if len(head) < 28:
                break
            name = head[:16].split(b"\0", 1)[0].decode()
            dtype, count = struct.unpack("<iq", head[16:28])
            dt = np.dtype(_DT[dtype])
            out[name] = np.frombuffer(f.read(count * dt.itemsize), dtype=dt).copy()
    return out


def write_sections(path, arrays: dict) -> None:
    with open(path, "wb") as f:
        for name, a in arrays.items():
            a = np.ascontiguousarray(a)
            f.write(name.encode()[:15].ljust(16, b"\0"))
            f.write(struct.pack("<iq", _CODE[a.dtype], a.size))
            f.write(a.tobytes())
