"""CPU: the upstream restatement that feeds the fixtures (OUT of the accelerated scope, but it decides what the draw
path is asked to draw, so it is pinned to the reference's own test expectations where they exist)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

REF = "/root/reference"
needs_ref = pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout only exists in the build container")


def _styler():
    from osm_renderer_b200.upstream import mapcss, styler as st

    return st.Styler(mapcss.load_rules_json(os.path.join(GOLDEN, "mapnik_rules.json.gz")), "josm", None)


@needs_ref
def test_parser_reproduces_reference_canonical_dump():
    # reference tests/test_mapcss_parser.rs:13-46
    from osm_renderer_b200.upstream import mapcss

    rules = mapcss.parse_file(os.path.join(REF, "tests/mapcss"), "mapnik.mapcss")
    canonical = open(os.path.join(REF, "tests/mapcss/mapnik.parsed.canonical"), newline="").read().replace("\r\n", "\n")
    assert mapcss.rules_to_string(rules) == canonical
    again = mapcss.parse_file(os.path.join(REF, "tests/mapcss"), "mapnik.parsed.canonical")
    assert mapcss.rules_to_string(again) == canonical
    committed = mapcss.load_rules_json(os.path.join(GOLDEN, "mapnik_rules.json.gz"))
    assert mapcss.rules_to_string(committed) == canonical


def test_styler_matches_josm_cascade_strings_of_the_reference_test(fx):
    # reference tests/test_mapcss_styler.rs:43-96
    from osm_renderer_b200.upstream import geodata, styler as st

    rd = geodata.GeodataReader(fx.bin)
    S = _styler()
    _, ways, _ = rd.get_entities_in_tile_with_neighbors(18, 158458, 81948)
    ents = []
    for w in ways:
        tags = rd.way_tags(int(w))
        if "name" in tags:
            ents.append((st.KIND_WAY_CLOSED if rd.way_is_closed(int(w)) else st.KIND_WAY_OPEN, int(w), int(rd.ways[int(w)]["id"]), tags))
    res = S.style_entities(ents, 18, False)

    def get(gid, name):
        return [s for e, s in res if e[2] == gid and e[3].get("name") == name]

    s1 = get(23369934, "Романов переулок")
    want1 = [
        (-1.0, (187, 187, 187), 16.0, None, st.CAP_ROUND),
        (3.0, (255, 255, 255), 13.0, [4.0, 2.0], st.CAP_ROUND),
        (15.0, (108, 112, 213), 1.0, [0.0, 12.0, 10.0, 152.0], st.CAP_BUTT),
        (15.1, (108, 112, 213), 2.0, [0.0, 12.0, 9.0, 153.0], st.CAP_BUTT),
        (15.2, (108, 112, 213), 3.0, [0.0, 18.0, 2.0, 154.0], st.CAP_BUTT),
        (15.3, (108, 112, 213), 4.0, [0.0, 18.0, 1.0, 155.0], st.CAP_BUTT),
    ]
    assert [(s.z_index, s.color, s.width, s.dashes, s.line_cap) for s in s1] == want1
    s2 = get(373569473, "Аллея Романов")
    assert [(s.z_index, s.color, s.width, s.dashes, s.line_cap) for s in s2] == [
        (-1.0, (128, 128, 128), 9.0, None, st.CAP_ROUND),
        (3.0, (237, 237, 237), 8.0, None, st.CAP_ROUND),
    ]
    for gid, name in [(31497212, "Бизнес-центр «Романов двор»"), (31482164, "Факультет искусств МГУ"), (44642919, "Факультет журналистики МГУ")]:
        s = get(gid, name)[0]
        assert (s.z_index, s.color, s.fill_color, s.fill_opacity, s.width, s.line_cap) == (-900.0, (51, 0, 102), (188, 169, 169), 0.9, 0.2, st.CAP_BUTT)


@needs_ref
def test_committed_fixture_inputs_are_reproducible(fx):
    """tools/make_fixtures.py must regenerate the committed geodata image and styled-area lists bit for bit."""
    from osm_renderer_b200.upstream import geodata, pipeline
    from osm_renderer_b200.wire import StyleTable

    data = geodata.import_osm(os.path.join(REF, "tests/osm/nano_moscow.osm"))
    assert data == fx.bin
    rd = geodata.GeodataReader(data)
    ts = pipeline.TileStyler(rd, _styler(), StyleTable(os.path.join(REF, "tests/mapcss")))
    tiles, begins, areas = fx.batches["17"]
    for i in (0, 9, 19):
        z, x, y, s = (int(v) for v in tiles[i])
        a = ts.areas_array(z, x, y)
        want = areas[begins[i] : begins[i + 1]]
        assert len(a) == len(want) and (a["entity"] == want["entity"]).all()


def test_tile_index_query_matches_brute_force(fx):
    """reader.rs:60-180 semantics: entity listed in every z18 tile of its node-tile bbox; 3x3 neighbourhood union."""
    from osm_renderer_b200.upstream import geodata

    rd = geodata.GeodataReader(fx.bin)
    z, x, y = 16, 39615, 20487
    _, ways, _ = rd.get_entities_in_tile_with_neighbors(z, x, y)
    mul = 1 << (18 - z)
    xa, xb, ya, yb = (x - 1) * mul, (x + 2) * mul - 1, (y - 1) * mul, (y + 2) * mul - 1
    tx = np.array([geodata.coords_to_max_zoom_tile(float(n["lat"]), float(n["lon"])) for n in rd.nodes])
    brute = []
    for w in range(len(rd.ways)):
        ids = rd.way_node_ids(w)
        if len(ids) == 0:
            continue
        t = tx[ids]
        if t[:, 0].min() <= xb and t[:, 0].max() >= xa and t[:, 1].min() <= yb and t[:, 1].max() >= ya:
            brute.append(w)
    assert list(ways) == brute


def test_vectorised_batch_builder_equals_per_tile_styler():
    from osm_renderer_b200.upstream import geodata, pipeline, synth
    from osm_renderer_b200.wire import StyleTable

    data = synth.make_metro(n=2)  # 2x2 z14 tiles: small but with every kind of synthetic feature
    assert data == synth.make_metro(n=2), "the generator must be deterministic"
    rd = geodata.GeodataReader(data)
    S = _styler()
    table = StyleTable(None)
    fast = pipeline.FastBatchBuilder(rd, S, table)
    slow = pipeline.TileStyler(rd, S, table)
    for (z, x, y) in [(14, 9888, 5104), (14, 9889, 5105), (15, 19777, 10209), (17, 79108, 40836)]:
        a, b = fast.areas_array(z, x, y), slow.areas_array(z, x, y)
        assert len(a) == len(b) and (a == b).all(), (z, x, y)
