"""CPU: the C-ABI library loads and exports every symbol include/osmr.h declares (no compute without a GPU),
the wire structs match the header, and the product package never touches the oracle."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def _declared_functions():
    hdr = open(os.path.join(ROOT, "include", "osmr.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(osmr_[a-z_0-9]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    from osm_renderer_b200 import _lib, build

    if not os.path.exists(_lib.LIB_PATH):
        build.build_lib()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), n
    assert set(names) == set(_lib.EXPORTS)
    assert lib.osmr_abi_version() == 5


def test_context_creation_fails_loudly_without_a_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from osm_renderer_b200._lib import OsmrError
    from osm_renderer_b200.drawer import GpuContext

    with pytest.raises(OsmrError):
        GpuContext(0)


def test_wire_struct_layouts_match_header():
    from osm_renderer_b200.wire import AREA_DTYPE, STYLE_DTYPE, TILE_DTYPE

    assert TILE_DTYPE.itemsize == 16 and AREA_DTYPE.itemsize == 8
    assert STYLE_DTYPE.itemsize == 72
    off = {n: STYLE_DTYPE.fields[n][1] for n in STYLE_DTYPE.names}
    assert (off["flags"], off["color"], off["line_cap"], off["fill_color"], off["casing_color"], off["fill_image"]) == (0, 4, 7, 8, 12, 16)
    assert (off["width"], off["opacity"], off["fill_opacity"], off["casing_width"], off["dashes_off"], off["casing_dashes_len"]) == (24, 32, 40, 48, 56, 68)


def test_c99_client_compiles_against_the_header_and_runs(tmp_path):
    """tests/abi_client.c: gcc -std=c99 against include/osmr.h (static assertions on every sizeof / offsetof that crosses the
    boundary), linked to libosmr_b200.so, one real call sequence (context creation fails loudly without a device)."""
    import shutil
    import subprocess

    from osm_renderer_b200 import _lib, build

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    if not os.path.exists(_lib.LIB_PATH):
        build.build_lib()
    exe = str(tmp_path / "abi_client")
    libdir = os.path.dirname(_lib.LIB_PATH)
    cmd = [gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "abi_client.c"), "-o", exe, "-L", libdir, "-losmr_b200", f"-Wl,-rpath,{libdir}"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    run = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert run.returncode == 0, run.stdout + run.stderr
    assert "abi_version" in run.stdout


def test_product_package_does_not_reference_the_oracle():
    pkg = os.path.join(ROOT, "osm_renderer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "libosmr_oracle" not in txt, f
