#!/usr/bin/env python
"""Dev tool (CPU only, no GPU): build and run tools/dev/layout_prof.cpp on the fixture label lists."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import CONFIG_NAMES, FixtureInputs  # noqa: E402
from osm_renderer_b200.wire import LABEL_STYLE_DTYPE  # noqa: E402

so = "/tmp/layout_prof.so"
subprocess.check_call(["g++", "-O2", "-g", "-std=c++17", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"), "-I",
                       os.path.join(ROOT, "osm_renderer_b200", "csrc"), os.path.join(ROOT, "tools", "dev", "layout_prof.cpp"), "-o", so,
                       "-lpthread"] + sys.argv[2:])
L = C.CDLL(so)
L.layout_prof.restype = C.c_double
fx = FixtureInputs()
lt, font, per = fx.labels()
styles = np.array(lt.rows, dtype=LABEL_STYLE_DTYPE)
strings = bytes(lt.strings)
wh = np.array([[w, h] for (w, h, _) in lt.icons], dtype=np.uint32).reshape(-1)
binb = np.frombuffer(fx.bin, dtype=np.uint8) if isinstance(fx.bin, (bytes, bytearray)) else np.ascontiguousarray(fx.bin)
fontb = np.frombuffer(font, dtype=np.uint8)
names = sys.argv[1].split(",") if len(sys.argv) > 1 and sys.argv[1] != "all" else CONFIG_NAMES
for name in names:
    tiles, _, _ = fx.batches[name]
    lb, labels = per[name]
    lb = np.ascontiguousarray(lb, dtype=np.uint32)
    labels = np.ascontiguousarray(labels)
    counts = (C.c_uint64 * 2)()
    ms = L.layout_prof(C.c_void_p(binb.ctypes.data), C.c_size_t(binb.size), C.c_void_p(fontb.ctypes.data), C.c_size_t(fontb.size),
                       C.c_void_p(styles.ctypes.data), C.c_uint32(len(styles)), strings, C.c_void_p(wh.ctypes.data), C.c_uint32(len(lt.icons)),
                       C.c_void_p(tiles.ctypes.data), C.c_uint32(len(tiles)), C.c_void_p(lb.ctypes.data), C.c_void_p(labels.ctypes.data),
                       C.c_int(3), counts)
    print(f"{name:6s} {len(tiles):3d} tiles {len(labels):6d} labels -> {counts[0]:6d} recs {counts[1]:8d} segs   {ms:8.2f} ms  ({ms / len(tiles):6.2f} ms/tile)")
