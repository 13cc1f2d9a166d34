"""MapCSS styler restatement: rules x entity tags x zoom -> ordered (entity, Style) list.

Upstream of the draw path (string matching on the host; NOT accelerated).  It is the *producer* of the
boundary the C ABI consumes: the ordered styled-area list of `Drawer::draw_to_pixels`
(/root/reference/src/draw/drawer.rs:75-78).

Follows /root/reference/src/mapcss/styler.rs:
  Style ............................ :49-72        Styler::new ............... :94-113
  style_entities ................... :115-166      style_areas (merge) ....... :168-203
  style_area (layer cascade) ....... :205-242      compare_styled_entities ... :246-272
  property_map_to_style ............ :277-429      canvas colour ............. :431-448
  matches_by_tags / area_matches ... :450-529      object types / z-index .... :531-579
and the memo key of src/mapcss/style_cache.rs:25-87 (a pure memo: same key => same styles).
Pinned by the JOSM cascade strings of /root/reference/tests/test_mapcss_styler.rs:43-96.
"""
from __future__ import annotations

from dataclasses import dataclass

from .mapcss import Rule

CAP_NONE, CAP_BUTT, CAP_ROUND, CAP_SQUARE = 0, 1, 2, 3

_COLOR_NAMES = {
    "white": (255, 255, 255),
    "black": (0, 0, 0),
    "blue": (0, 0, 255),
    "brown": (165, 42, 42),
    "green": (0, 255, 0),
    "grey": (128, 128, 128),
    "pink": (255, 192, 203),
    "purple": (128, 0, 128),
    "red": (255, 0, 0),
    "salmon": (250, 128, 114),
}

BASE_LAYER_NAME = "default"

# entity kinds == style-cache slots (styler.rs:557-579)
KIND_NODE, KIND_WAY_CLOSED, KIND_WAY_OPEN, KIND_MULTIPOLYGON = 0, 1, 2, 3


@dataclass(eq=False)
class TextStyle:
    text: str
    text_color: tuple | None
    text_position: str | None  # "center" | "line"
    font_size: float | None


@dataclass(eq=False)
class Style:
    layer: int | None
    z_index: float
    color: tuple | None
    fill_color: tuple | None
    is_foreground_fill: bool
    background_color: tuple | None
    opacity: float | None
    fill_opacity: float | None
    width: float | None
    dashes: list | None
    line_cap: int  # CAP_*
    casing_color: tuple | None
    casing_width: float | None
    casing_dashes: list | None
    casing_line_cap: int
    icon_image: str | None
    fill_image: str | None
    text_style: TextStyle | None


def _parse_i64(s: str):
    """Rust `str::parse::<i64>`: optional sign, ASCII digits only, no whitespace."""
    if not s:
        return None
    body = s[1:] if s[0] in "+-" else s
    if not body or not all("0" <= c <= "9" for c in body):
        return None
    v = int(s)
    if v < -(1 << 63) or v >= (1 << 63):
        return None
    return v


def _parse_f64(s: str):
    """Rust `str::parse::<f64>` accepts decimal/exponent forms, 'inf', 'infinity', 'nan' (case-insens.)."""
    if not s or s != s.strip() or "_" in s:
        return None
    try:
        return float(s)
    except ValueError:
        return None


def _is_true_value(v: str) -> bool:
    return v == "yes" or v == "true" or v == "1"


class Styler:
    def __init__(self, rules: list[Rule], style_type: str = "josm", font_size_multiplier: float | None = None):
        self.rules = rules
        self.use_caps_for_dashes = style_type == "josm"
        self.casing_width_multiplier = 1.0 if style_type == "mapsme" else 2.0
        self.font_size_multiplier = font_size_multiplier
        self.canvas_fill_color = self._extract_canvas_fill_color(style_type)
        # style_cache.rs:25-56
        tvm = {"layer": True}
        for r in rules:
            for sel in r.selectors:
                for t in sel.tests:
                    matters = not (t.kind == "unary" and t.op in ("exists", "not_exists"))
                    tvm[t.tag] = tvm.get(t.tag, False) or matters
        self.tag_value_matters = tvm
        self._cache: dict = {}
        self._sel_index: dict = {}

    def _extract_canvas_fill_color(self, style_type):
        prop = "fill-color" if style_type == "josm" else "background-color"
        for r in self.rules:
            for sel in r.selectors:
                if sel.object_type == "canvas":
                    for p in r.properties:
                        if p.name == prop and p.kind == "color":
                            return p.value
        return None

    # ------------------------------------------------------------------------------------------
    def _selectors_for(self, zoom: int, kind: int):
        key = (zoom, kind)
        lst = self._sel_index.get(key)
        if lst is None:
            lst = []
            for r in self.rules:
                for sel in r.selectors:
                    if sel.min_zoom is not None and zoom < sel.min_zoom:
                        continue
                    if sel.max_zoom is not None and zoom > sel.max_zoom:
                        continue
                    ot = sel.object_type
                    if kind == KIND_NODE:
                        ok = ot == "node"
                    elif kind == KIND_WAY_OPEN:
                        ok = ot == "way"
                    else:  # closed way / multipolygon
                        ok = ot in ("way", "area")
                    if ok:
                        lst.append((r, sel))
            self._sel_index[key] = lst
        return lst

    @staticmethod
    def _test_matches(tags: dict, t) -> bool:
        v = tags.get(t.tag)
        if t.kind == "unary":
            if t.op == "exists":
                return v is not None
            if t.op == "not_exists":
                return v is None
            tv = v is not None and _is_true_value(v)
            return tv if t.op == "true" else not tv
        if t.kind == "str":
            return (v == t.value) if t.op == "=" else (v != t.value)
        if v is None:
            return False
        f = _parse_f64(v)
        if f is None:
            return False
        if t.op == "<":
            return f < t.value
        if t.op == "<=":
            return f <= t.value
        if t.op == ">":
            return f > t.value
        return f >= t.value

    def _style_area(self, tags: dict, zoom: int, kind: int) -> dict:
        """styler.rs:205-242: layer id -> property map (insertion ordered)."""
        result: dict = {}
        match = self._test_matches
        for rule, sel in self._selectors_for(zoom, kind):
            if not all(match(tags, t) for t in sel.tests):
                continue
            layer_id = sel.layer_id if sel.layer_id is not None else BASE_LAYER_NAME
            if layer_id not in result:
                result[layer_id] = dict(result.get("*", {}))
            layer = result[layer_id]
            for p in rule.properties:
                layer[p.name] = p
            if layer_id == "*":
                for k, v in result.items():
                    if k != "*":
                        for p in rule.properties:
                            v[p.name] = p
        return result

    def _to_style(self, cur: dict, base: dict | None, default_z: float, tags: dict) -> Style:
        """styler.rs:277-429."""

        def get_color(name):
            p = cur.get(name)
            if p is None:
                return None
            if p.kind == "color":
                return tuple(p.value)
            if p.kind == "ident":
                return _COLOR_NAMES.get(p.value)
            return None

        def get_num(m, name):
            p = m.get(name)
            if p is not None and p.kind == "numbers" and len(p.value) == 1:
                return p.value[0]
            return None

        def get_id(name):
            p = cur.get(name)
            return p.value if p is not None and p.kind == "ident" else None

        def get_string(name):
            p = cur.get(name)
            return p.value if p is not None and p.kind in ("ident", "string") else None

        def get_cap(name):
            v = get_id(name)
            if v in ("none", "butt"):
                return CAP_BUTT
            if v == "round":
                return CAP_ROUND
            if v == "square":
                return CAP_SQUARE
            return CAP_NONE

        def get_dashes(name):
            p = cur.get(name)
            return list(p.value) if p is not None and p.kind == "numbers" else None

        layer_tag = tags.get("layer")
        layer = _parse_i64(layer_tag) if layer_tag is not None else None
        z = get_num(cur, "z-index")
        z_index = z if z is not None else default_z
        fp = cur.get("fill-position")
        is_fg = not (fp is not None and fp.kind == "ident" and fp.value == "background")
        width = get_num(cur, "width")
        base_w = width
        if base_w is None and base is not None:
            base_w = get_num(base, "width")
        if base_w is None:
            base_w = 0.0
        cw = cur.get("casing-width")
        casing_only = None
        if cw is not None:
            if cw.kind == "numbers" and len(cw.value) == 1:
                casing_only = cw.value[0]
            elif cw.kind == "width_delta":
                casing_only = base_w + cw.value
        full_casing = None if casing_only is None else base_w + self.casing_width_multiplier * casing_only
        text = get_string("text")
        fs = get_num(cur, "font-size")
        if fs is not None:
            fs = fs * (self.font_size_multiplier if self.font_size_multiplier is not None else 1.0)
        text_style = None
        if text is not None:
            tp = get_id("text-position")
            text_style = TextStyle(text, get_color("text-color"), tp if tp in ("center", "line") else None, fs)
        return Style(
            layer=layer,
            z_index=z_index,
            color=get_color("color"),
            fill_color=get_color("fill-color"),
            is_foreground_fill=is_fg,
            background_color=get_color("background-color"),
            opacity=get_num(cur, "opacity"),
            fill_opacity=get_num(cur, "fill-opacity"),
            width=width,
            dashes=get_dashes("dashes"),
            line_cap=get_cap("linecap"),
            casing_color=get_color("casing-color"),
            casing_width=full_casing,
            casing_dashes=get_dashes("casing-dashes"),
            casing_line_cap=get_cap("casing-linecap"),
            icon_image=get_string("icon-image"),
            fill_image=get_string("fill-image"),
            text_style=text_style,
        )

    def styles_for(self, tags: dict, zoom: int, kind: int) -> list:
        """Memoised per style_cache.rs key: (slot, relevant tags[+values], zoom)."""
        tvm = self.tag_value_matters
        key_tags = []
        for k in sorted(tags.keys()):
            m = tvm.get(k)
            if m is not None:
                key_tags.append((k, tags[k]) if m else (k,))
        key = (kind, tuple(key_tags), zoom)
        styles = self._cache.get(key)
        if styles is None:
            default_z = 4.0 if kind == KIND_NODE else (3.0 if kind == KIND_WAY_OPEN else 1.0)
            maps = self._style_area(tags, zoom, kind)
            base = maps.get(BASE_LAYER_NAME)
            styles = [self._to_style(m, base, default_z, tags) for layer, m in maps.items() if layer != "*"]
            self._cache[key] = styles
        return styles

    # ------------------------------------------------------------------------------------------
    @staticmethod
    def _sort_key(for_labels: bool):
        if for_labels:
            return lambda e: (e[1].layer or 0, e[1].z_index, e[0][2])
        return lambda e: (e[1].layer or 0, e[1].is_foreground_fill, e[1].z_index, e[0][2])

    def style_entities(self, entities, zoom: int, for_labels: bool):
        """styler.rs:115-166.  `entities`: iterable of (kind, local_id, global_id, tags).

        Returns [(entity, Style)] stably sorted by (layer, [is_foreground_fill], z_index, global id).
        """
        out = []
        for ent in entities:
            for s in self.styles_for(ent[3], zoom, ent[0]):
                out.append((ent, s))
        out.sort(key=self._sort_key(for_labels))
        return out

    def style_areas(self, ways, multipolygons, zoom: int, for_labels: bool):
        """styler.rs:168-203: both lists sorted, then merged, multipolygon first on ties."""
        sw = self.style_entities(ways, zoom, for_labels)
        sm = self.style_entities(multipolygons, zoom, for_labels)
        key = self._sort_key(for_labels)
        res = []
        i = j = 0
        while i < len(sm) or j < len(sw):
            if j >= len(sw):
                take_mp = True
            elif i >= len(sm):
                take_mp = False
            else:
                take_mp = not (key(sm[i]) > key(sw[j]))
            if take_mp:
                res.append(sm[i])
                i += 1
            else:
                res.append(sw[j])
                j += 1
        return res
