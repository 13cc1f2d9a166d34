import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
CONFIG_NAMES = ["14", "15", "16", "17", "18", "18_2x"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # OSMR_TEST_EMU=1 python -m pytest tests -m gpu: run the GPU-marked parity tests against the kernels compiled for
    # the host SIMT emulator (tests/emu/): a check of kernel LOGIC in the GPU-less container before a GPU box is
    # spent.  Test infrastructure only; never set on the GPU box, where the same tests load libosmr_b200.so.
    if os.environ.get("OSMR_TEST_EMU") == "1":
        sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
        import build_emu

        from osm_renderer_b200 import _lib

        _lib.LIB_PATH = build_emu.build()


class FixtureInputs:
    """tests/golden/fixture_inputs.npz + nano_moscow.bin (made by tools/make_fixtures.py)."""

    def __init__(self):
        from osm_renderer_b200.wire import StyleTable

        with open(os.path.join(GOLDEN, "nano_moscow.bin"), "rb") as f:
            self.bin = f.read()
        z = np.load(os.path.join(GOLDEN, "fixture_inputs.npz"))
        self.canvas_rgb = tuple(int(v) for v in z["canvas_rgb"])
        self.use_caps_for_dashes = bool(z["use_caps_for_dashes"])
        self.table = StyleTable(None)
        self.table.rows = list(z["styles"])
        self.table.dashes = list(z["dashes"])
        for i in range(int(z["n_icons"])):
            self.table.add_raw_icon(f"icon_{i}", z[f"icon_{i}"])
        self.batches = {n: (z[f"tiles_{n}"], z[f"area_begin_{n}"], z[f"areas_{n}"]) for n in CONFIG_NAMES}

    def labels(self):
        """tests/golden/label_inputs.npz -> (LabelStyleTable, font bytes, {config: (label_begin, labels)})"""
        from osm_renderer_b200.wire import LabelStyleTable

        if getattr(self, "_labels", None) is None:
            z = np.load(os.path.join(GOLDEN, "label_inputs.npz"))
            lt = LabelStyleTable(None)
            lt.rows = list(z["label_styles"])
            lt.strings = bytearray(z["label_strings"].tobytes())
            lt.icons = []
            for i in range(int(z["n_label_icons"])):
                px = np.ascontiguousarray(z[f"label_icon_{i}"])
                lt.icons.append((px.shape[1], px.shape[0], px))
            per = {}
            for n in CONFIG_NAMES:
                src = "18" if n == "18_2x" else n
                per[n] = (z[f"label_begin_{src}"], z[f"labels_{src}"])
            self._labels = (lt, z["font"].tobytes(), per)
        return self._labels

    def icon_loader(self):
        """name -> (w, h, rgba) from the committed fixtures (fill patterns + label icons); None for unknown names,
        which is what a failed Icon::load looks like."""
        z = np.load(os.path.join(GOLDEN, "fixture_inputs.npz"))
        lz = np.load(os.path.join(GOLDEN, "label_inputs.npz"))
        icons = {}
        for i, n in enumerate(z["icon_names"]):
            px = np.ascontiguousarray(z[f"icon_{i}"])
            icons[str(n)] = (px.shape[1], px.shape[0], px)
        for i, n in enumerate(lz["label_icon_names"]):
            px = np.ascontiguousarray(lz[f"label_icon_{i}"])
            icons[str(n)] = (px.shape[1], px.shape[0], px)
        return lambda name: icons.get(name)

    def golden(self, name):
        g = np.load(os.path.join(GOLDEN, f"golden_{name}.npz"))
        d = int(g["dim"])
        mask = np.unpackbits(g["label_mask"], axis=-1)[..., :d].astype(bool)
        return g["golden"], mask


@pytest.fixture(scope="session")
def fx():
    return FixtureInputs()


@pytest.fixture(scope="session")
def gpu_ctx(fx):
    from osm_renderer_b200.drawer import GpuContext

    ctx = GpuContext(0)
    ctx.set_geodata(fx.bin)
    ctx.set_table(fx.table)
    yield ctx
    ctx.close()
