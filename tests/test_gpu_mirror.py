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


def test_one_drawer_serves_two_tile_pixels_and_a_second_drawer_replaces_the_tables(fx):
    """The reference shares one Drawer between worker threads, each with its own TilePixels (http_server.rs:42-48,69-72):
    every context must receive the style / label tables itself; and a context that has served one Drawer must take the
    tables of another one even when they are equally long."""
    from osm_renderer_b200.drawer import Drawer, OsmEntities, Tile, TilePixels
    from osm_renderer_b200.upstream import geodata, mapcss, styler as st

    reader = geodata.GeodataReader(fx.bin)
    styler = st.Styler(mapcss.load_rules_json(os.path.join(GOLDEN, "mapnik_rules.json.gz")), "josm", None)
    _, font, _ = fx.labels()
    golden, _ = fx.golden("16")
    x, y, i = 39615, 20487, 4
    nodes, ways, mps = reader.get_entities_in_tile_with_neighbors(16, x, y)
    ents = OsmEntities(reader, nodes, ways, mps)

    def check(rendered):
        diff = (rendered.triples != golden[i]).any(axis=-1)
        diff[0, :] = False
        diff[:, 255] = False
        assert diff.sum() == 0, int(diff.sum())

    drawer = Drawer(None, font=font, icon_loader=fx.icon_loader())
    p1, p2 = TilePixels(1), TilePixels(1)
    check(drawer.draw_to_pixels(ents, Tile(16, x, y), p1, 1, styler))
    check(drawer.draw_to_pixels(ents, Tile(16, x, y), p2, 1, styler))  # second context, same Drawer
    drawer2 = Drawer(None, font=font, icon_loader=fx.icon_loader())  # same table sizes, different objects
    check(drawer2.draw_to_pixels(ents, Tile(16, x, y), p1, 1, styler))
    p1.ctx.close()
    p2.ctx.close()
