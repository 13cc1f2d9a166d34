#!/usr/bin/env python
"""Generate tests/golden/* from the reference's own fixtures (run in the build container only).

Inputs (read-only, /root/reference): tests/osm/nano_moscow.osm, tests/mapcss/mapnik.mapcss (+ symbols/),
tests/rendered/{14,15,16,17,18,18_2x}_expected.png and the tile sets of tests/test_rendering.rs:147-176.
Outputs (committed, travel to the GPU box where /root/reference does not exist):
  tests/golden/nano_moscow.bin        geodata image written by the restated importer (saver.rs format)
  tests/golden/fixture_inputs.npz     style table, dashes, icons, canvas colour, per-config tile lists and
                                      ordered styled-area lists (the styler output = C-ABI input)
  tests/golden/label_inputs.npz       label generations per tile, label-style table, label icons, font bytes (GPU label tests)
  tests/golden/*_rules.json.gz        parsed rule lists of tests/mapcss/mapnik.mapcss and mapcss/osmosnimki-minimal.mapcss
  tests/golden/golden_<cfg>.npz       reference golden pixels per tile + `label_mask`: pixels the reference's
                                      label pass (drawer.rs:106-126, not restated yet) or the red test grid
                                      (test_rendering.rs:109-114) touched.  The mask is *derived*: it is the set of
                                      pixels where the golden differs from the area-only oracle render at the
                                      time of generation (inspected: glyphs and icons only, 0.4-2.5 % of pixels).
"""
import os
import sys
import time

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from osm_renderer_b200.upstream import geodata, mapcss, pipeline, styler as st  # noqa: E402
from osm_renderer_b200.wire import LABEL_DTYPE, LabelStyleTable, StyleTable  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")

# tests/test_rendering.rs:147-176
CONFIGS = {
    "14": (14, 9903, 9904, 5121, 5122, 1),
    "15": (15, 19807, 19808, 10243, 10244, 1),
    "16": (16, 39614, 39616, 20486, 20488, 1),
    "17": (17, 79228, 79232, 40973, 40976, 1),
    "18": (18, 158457, 158465, 81946, 81953, 1),
    "18_2x": (18, 158457, 158465, 81946, 81953, 2),
}


def main():
    os.makedirs(OUT, exist_ok=True)
    t0 = time.time()
    data = geodata.import_osm(os.path.join(REF, "tests/osm/nano_moscow.osm"))
    with open(os.path.join(OUT, "nano_moscow.bin"), "wb") as f:
        f.write(data)
    rd = geodata.GeodataReader(data)
    S = st.Styler(mapcss.parse_file(os.path.join(REF, "tests/mapcss"), "mapnik.mapcss"), "josm", None)
    table = StyleTable(os.path.join(REF, "tests/mapcss"))
    # parsed stylesheets for the benchmark's synthetic data (the .mapcss files do not exist on the GPU box)
    mapcss.save_rules_json(S.rules, os.path.join(OUT, "mapnik_rules.json.gz"))
    mapcss.save_rules_json(
        mapcss.parse_file(os.path.join(REF, "mapcss"), "osmosnimki-minimal.mapcss"), os.path.join(OUT, "osmosnimki_rules.json.gz")
    )
    ts = pipeline.TileStyler(rd, S, table)

    batches = {}
    for name, (z, x0, x1, y0, y1, s) in CONFIGS.items():
        tiles = [(z, x, y, s) for y in range(y0, y1 + 1) for x in range(x0, x1 + 1)]
        batches[name] = pipeline.build_batch(ts, tiles)

    save = {
        "styles": table.styles_array(),
        "dashes": table.dashes_array(),
        "canvas_rgb": np.asarray(S.canvas_fill_color, dtype=np.uint8),
        "use_caps_for_dashes": np.asarray(S.use_caps_for_dashes),
        "n_icons": np.asarray(len(table.icons)),
    }
    for i, (w, h, px) in enumerate(table.icons):
        save[f"icon_{i}"] = px
    save["icon_names"] = np.array(table.icon_names)
    for name, (tarr, begins, areas) in batches.items():
        save[f"tiles_{name}"] = tarr
        save[f"area_begin_{name}"] = begins
        save[f"areas_{name}"] = areas
    np.savez_compressed(os.path.join(OUT, "fixture_inputs.npz"), **save)

    # label pass inputs for the GPU tests: label generations per tile (styler order), label-style table, label icons
    # and the font the reference embeds (src/draw/font/NotoSans-Regular.ttf, SIL OFL; stored as a byte array)
    ltable = LabelStyleTable(os.path.join(REF, "tests/mapcss"))
    lb = pipeline.LabelListBuilder(ts, os.path.join(REF, "tests/mapcss"))
    lsave = {}
    for name, (z, x0, x1, y0, y1, s) in CONFIGS.items():
        if name == "18_2x":
            continue  # same labels as "18"
        parts, lbeg = [], [0]
        for y in range(y0, y1 + 1):
            for x in range(x0, x1 + 1):
                parts.append(lb.labels_abi(z, x, y, ltable))
                lbeg.append(lbeg[-1] + len(parts[-1]))
        lsave[f"labels_{name}"] = np.concatenate(parts) if parts else np.zeros(0, dtype=LABEL_DTYPE)
        lsave[f"label_begin_{name}"] = np.asarray(lbeg, dtype=np.uint32)
    lsave["label_styles"] = ltable.styles_array()
    lsave["label_strings"] = np.frombuffer(bytes(ltable.strings), dtype=np.uint8)
    lsave["n_label_icons"] = np.asarray(len(ltable.icons))
    for i, (w, h, px) in enumerate(ltable.icons):
        lsave[f"label_icon_{i}"] = px
    lsave["label_icon_names"] = np.array(ltable.icon_names)
    with open(os.path.join(REF, "src/draw/font/NotoSans-Regular.ttf"), "rb") as f:
        lsave["font"] = np.frombuffer(f.read(), dtype=np.uint8)
    np.savez_compressed(os.path.join(OUT, "label_inputs.npz"), **lsave)
    print(f"label fixtures: {len(ltable.rows)} label styles, {len(ltable.icons)} icons")

    for name, (z, x0, x1, y0, y1, s) in CONFIGS.items():
        tarr, begins, areas = batches[name]
        imgs = oracle.draw_tiles(data, table, tarr, begins, areas, S.canvas_fill_color, S.use_caps_for_dashes, n_threads=8)
        gold = np.asarray(Image.open(os.path.join(REF, f"tests/rendered/{name}_expected.png")).convert("RGB"))
        D = 256 * s
        nx = x1 - x0 + 1
        g_tiles = np.empty((len(imgs), D, D, 3), dtype=np.uint8)
        masks = np.empty((len(imgs), D, D), dtype=bool)
        for i in range(len(imgs)):
            r, c = divmod(i, nx)
            g = gold[r * D : (r + 1) * D, c * D : (c + 1) * D]
            g_tiles[i] = g
            m = (imgs[i] != g).any(axis=2)
            m[0, :] = True       # red grid: row 0 ...
            m[:, D - 1] = True   # ... and the last column of every tile (test_rendering.rs:109-114)
            masks[i] = m
        frac = masks.mean()
        print(f"{name}: {len(imgs)} tiles, label+grid mask {100 * frac:.3f} % of pixels")
        np.savez_compressed(
            os.path.join(OUT, f"golden_{name}.npz"), golden=g_tiles, label_mask=np.packbits(masks, axis=-1), dim=np.asarray(D)
        )
    print("done in %.1fs" % (time.time() - t0))


if __name__ == "__main__":
    main()
