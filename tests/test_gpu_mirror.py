"""-m gpu: the host-side mirror of the reference interface, driven the way the reference's own integration test
drives the reference (tests/test_rendering.rs:69-100): load geodata -> Styler -> Drawer -> per tile
get_entities_in_tile_with_neighbors + draw_to_pixels -> compare with the golden render (every pixel, labels included)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def test_drawer_mirror_reproduces_golden_z16(fx):
    from osm_renderer_b200.drawer import Drawer, OsmEntities, Tile, TilePixels
    from osm_renderer_b200.upstream import geodata, mapcss, styler as st

    reader = geodata.GeodataReader(fx.bin)
    styler = st.Styler(mapcss.load_rules_json(os.path.join(GOLDEN, "mapnik_rules.json.gz")), "josm", None)
    _, font, _ = fx.labels()
    drawer = Drawer(None, font=font, icon_loader=fx.icon_loader())
    pixels = TilePixels(1)
    golden, _ = fx.golden("16")
    i = 0
    for y in range(20486, 20489):
        for x in range(39614, 39617):
            nodes, ways, mps = reader.get_entities_in_tile_with_neighbors(16, x, y)
            rendered = drawer.draw_to_pixels(OsmEntities(reader, nodes, ways, mps), Tile(16, x, y), pixels, 1, styler)
            assert rendered.dimension == 256
            diff = (rendered.triples != golden[i]).any(axis=-1)
            diff[0, :] = False
            diff[:, 255] = False
            assert diff.sum() == 0, (x, y, int(diff.sum()))
            i += 1
    pixels.ctx.close()
