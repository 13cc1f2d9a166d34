"""MapCSS tokenizer + parser restatement (upstream of the draw path; NOT accelerated).

Follows /root/reference/src/mapcss/token.rs:127-480 (tokenizer) and parser.rs:223-699 (parser),
including the `Display` impls (parser.rs:25-221) so that the reference's own canonical dump
`tests/mapcss/mapnik.parsed.canonical` pins this restatement (tests/test_mapcss_parser.rs:13-46).
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

# token kinds
IMPORT, IDENT, STRING, NUMBER, ZOOM, COLORREF, COLOR, SIMPLE = range(8)

_TWO = {"!=": "!=", "<=": "<=", ">=": ">=", "=~": "=~", "::": "::"}
_ONE = set("()[]{}=<>!?:;,")


class MapcssError(Exception):
    pass


@dataclass
class Token:
    kind: int
    value: object
    pos: tuple


def _can_start_ident(ch: str) -> bool:
    return ch == "_" or ("a" <= ch <= "z") or ("A" <= ch <= "Z")


def _can_continue_ident(ch: str) -> bool:
    return ch in "-./" or ("0" <= ch <= "9") or _can_start_ident(ch)


def _in_at_directive(ch: str) -> bool:
    return ch == "_" or ("a" <= ch <= "z") or ("0" <= ch <= "9")


def _is_ascii_digit(ch: str) -> bool:
    return "0" <= ch <= "9"


class Tokenizer:
    def __init__(self, text: str):
        self.t = text
        self.i = 0
        self.line = 1
        self.col = 0
        self.had_newline = False

    # -- char helpers (token.rs:370-400) ---------------------------------------------------------
    def _next(self):
        if self.i >= len(self.t):
            ch = None
        else:
            ch = self.t[self.i]
            self.i += 1
        if self.had_newline:
            self.line += 1
            self.col = 0
            self.had_newline = False
        self.col += 1
        self.had_newline = ch == "\n"
        return ch

    def _peek(self):
        return self.t[self.i] if self.i < len(self.t) else None

    def _err(self, msg):
        raise MapcssError(f"lexer error: {msg} (at line {self.line}, col {self.col})")

    # -- main --------------------------------------------------------------------------------------
    def next_token(self):
        while True:
            ch = self._next()
            if ch is None:
                return None
            if ch.isspace():
                continue
            if ch == "/":
                p = self._peek()
                if p == "/":
                    self._next()
                    while True:
                        c = self._next()
                        if c is None or c == "\n":
                            break
                    continue
                if p == "*":
                    self._next()
                    while True:
                        c = self._next()
                        if c is None:
                            self._err("Unterminated block comment")
                        if c == "*" and self._peek() == "/":
                            self._next()
                            break
                    continue
            pos = (self.line, self.col)
            start = self.i - 1
            return Token(*self._read_token(start, ch), pos)

    def _read_token(self, idx, ch):
        nxt = self._peek()
        if nxt is not None and (ch + nxt) in _TWO:
            self._next()
            return SIMPLE, ch + nxt
        if ch in _ONE:
            return SIMPLE, ch
        if ch == "@":
            return self._read_at()
        if ch == "*":
            return IDENT, "*"
        if _can_start_ident(ch):
            return self._read_ident(idx)
        if ch == '"':
            return self._read_string()
        if _is_ascii_digit(ch) or ch in "+.":
            return self._read_number(ch)
        if ch == "-":
            if nxt is not None and _is_ascii_digit(nxt):
                return self._read_number(ch)
            if nxt is not None and _can_continue_ident(nxt):
                return self._read_ident(idx)
            self._err("Expected a valid number or identifier after '-'")
        if ch == "|":
            return self._read_zoom()
        if ch == "#":
            return self._read_color()
        self._err(f"Unexpected symbol: '{ch}'")

    def _read_at(self):
        c = self._next()
        if c is None or not _in_at_directive(c):
            self._err("Expected a letter or underscore after @")
        start = self.i - 1
        while self._peek() is not None and _in_at_directive(self._peek()):
            self._next()
        text = self.t[start : self.i]
        if text == "import":
            p = self._peek()
            if p is not None and (p.isspace() or p == "("):
                self._next()
            c = self._next()
            if c != '"':
                self._err("Expected a string")
            _, s = self._read_string()
            p = self._peek()
            if p is not None and (p.isspace() or p == ")"):
                self._next()
            return IMPORT, s
        return COLORREF, text

    def _read_ident(self, start):
        while self._peek() is not None and _can_continue_ident(self._peek()):
            self._next()
        return IDENT, self.t[start : self.i]

    def _read_string(self):
        start = self.i
        while True:
            c = self._next()
            if c is None:
                self._err("Unterminated string")
            if c == '"':
                return STRING, self.t[start : self.i - 1]

    def _read_number(self, first):
        sign = 1.0
        if first in "+-":
            c = self._next()
            if c is None:
                self._err("Expected a digit after '-' or '+'")
            sign = -1.0 if first == "-" else 1.0
            first = c
        had_dot = False
        if _is_unicode_digit(first):
            number = float(int(first))
        elif first == ".":
            had_dot = True
            number = 0.0
        else:
            self._err(f"Expected a digit or '.' instead of '{first}'")
        after = 0.0
        n_after = 0
        while True:
            p = self._peek()
            if p is None:
                break
            if _is_unicode_digit(p):
                if had_dot:
                    n_after += 1
                    after = 10.0 * after + float(int(p))
                else:
                    number = 10.0 * number + float(int(p))
                self._next()
            elif p == "." and not had_dot:
                had_dot = True
                self._next()
            else:
                break
        if had_dot and n_after == 0:
            self._err("Expected a digit after '.'")
        if n_after > 0:
            number += after / _powi10(n_after)
        return NUMBER, sign * number

    def _read_digit(self, radix):
        p = self._peek()
        if p is None:
            return None
        try:
            v = int(p, radix)
        except ValueError:
            return None
        if not (p.isascii() and p.isalnum()):
            return None
        self._next()
        return v

    def _read_color(self):
        digits = []
        while True:
            d = self._read_digit(16)
            if d is None:
                break
            digits.append(d)
        if len(digits) == 6:
            c = (digits[0] * 16 + digits[1], digits[2] * 16 + digits[3], digits[4] * 16 + digits[5])
        elif len(digits) == 3:
            c = (digits[0] * 17, digits[1] * 17, digits[2] * 17)
        else:
            self._err("Invalid hex color (expected #RGB or #RRGGBB)")
        return COLOR, c

    def _read_zoom_level(self):
        a = self._read_digit(10)
        if a is None:
            return None
        b = self._read_digit(10)
        return a if b is None else 10 * a + b

    def _read_zoom(self):
        if self._next() != "z":
            self._err("Expected 'z' character")
        mn = self._read_zoom_level()
        had_hyphen = False
        if self._peek() == "-":
            self._next()
            had_hyphen = True
        mx = self._read_zoom_level()
        if mn is None and mx is None:
            self._err("A zoom range should have either minumum or maximum level")
        return ZOOM, (mn, mx if had_hyphen else mn)


def _is_unicode_digit(ch: str) -> bool:
    # Rust char::to_digit(10) accepts ASCII digits only.
    return "0" <= ch <= "9"


def _powi10(n: int) -> float:
    # 10.0f64.powi(n): repeated multiplication is exact for the n (< 23) that occur in stylesheets.
    r = 1.0
    for _ in range(n):
        r *= 10.0
    return r


# ------------------------------------------------------------------------------------------------
# AST (parser.rs:12-221)
# ------------------------------------------------------------------------------------------------
@dataclass
class Test:
    kind: str  # "unary" | "str" | "num"
    tag: str
    op: str  # unary: exists/not_exists/true/false ; str: "=" "!=" ; num: < <= > >=
    value: object = None

    def __str__(self):
        q = f'"{self.tag}"' if ":" in self.tag else self.tag
        if self.kind == "unary":
            body = {"exists": q, "not_exists": "!" + q, "true": q + "?", "false": "!" + q + "?"}[self.op]
        elif self.kind == "str":
            body = f"{q}{self.op}{self.value}"
        else:
            body = f"{q}{self.op}{fmt_f64(self.value)}"
        return f"[{body}]"


@dataclass
class Selector:
    object_type: str  # "*", canvas, meta, node, way, area
    min_zoom: int | None = None
    max_zoom: int | None = None
    tests: list = field(default_factory=list)
    layer_id: str | None = None

    def __str__(self):
        mn, mx = self.min_zoom, self.max_zoom
        if mn is None and mx is None:
            z = ""
        elif mx is None:
            z = f"{mn}-"
        elif mn is None:
            z = f"-{mx}"
        else:
            z = f"{mn}-{mx}" if mn != mx else f"{mn}"
        layer = f"::{self.layer_id}" if self.layer_id is not None else ""
        return f"{self.object_type}{'|z' if z else ''}{z}{''.join(str(t) for t in self.tests)}{layer}"


@dataclass
class Property:
    name: str
    kind: str  # ident | string | color | numbers | width_delta
    value: object

    def __str__(self):
        if self.kind == "color":
            r, g, b = self.value
            v = f"#{r:02x}{g:02x}{b:02x}"
        elif self.kind == "ident":
            v = self.value
        elif self.kind == "string":
            v = f'"{self.value}"'
        elif self.kind == "numbers":
            v = ",".join(fmt_f64(n) for n in self.value)
        else:
            v = f'eval(prop("width")) + {fmt_f64(self.value)}'
        return f"{self.name}: {v};"


@dataclass
class Rule:
    selectors: list
    properties: list

    def __str__(self):
        return "{} {{\n{}\n}}".format(
            ",\n".join(str(s) for s in self.selectors), "\n".join(str(p) for p in self.properties)
        )


def fmt_f64(v: float) -> str:
    """Rust `{}` for f64: shortest round-trip digits, never scientific, no trailing '.0'."""
    if v != v:
        return "NaN"
    if v in (float("inf"), float("-inf")):
        return "inf" if v > 0 else "-inf"
    if v == int(v) and abs(v) < 1e16:
        s = str(int(v))
        if s == "0" and str(v).startswith("-"):
            return "-0"
        return s
    r = repr(v)
    if "e" in r or "E" in r:
        from decimal import Decimal

        r = format(Decimal(r), "f")
    return r


_OBJECT_TYPES = {"*": "*", "canvas": "canvas", "meta": "meta", "node": "node", "way": "way", "line": "way", "area": "area"}


class Parser:
    def __init__(self, base_path: str, file_name: str, color_defs: dict | None = None):
        self.base_path = base_path
        self.file_name = file_name
        with open(os.path.join(base_path, file_name), "r", encoding="utf-8", newline="") as f:
            self.tok = Tokenizer(f.read())
        self.color_defs = dict(color_defs or {})

    def _perr(self, msg, pos):
        return MapcssError(f"parse error: {msg} ({self.file_name} at line {pos[0]}, col {pos[1]})")

    def _opt(self):
        return self.tok.next_token()

    def _must(self) -> Token:
        t = self.tok.next_token()
        if t is None:
            raise self._perr("Unexpected end of file", (self.tok.line, self.tok.col))
        return t

    def _expect(self, simple: str):
        t = self._must()
        if not (t.kind == SIMPLE and t.value == simple):
            raise self._perr(f"Expected '{simple}', found '{t.value}' instead", t.pos)

    def _unexpected(self, t: Token):
        return self._perr(f"Unexpected token: '{t.value}'", t.pos)

    def _ident(self) -> str:
        t = self._must()
        if t.kind != IDENT:
            raise self._unexpected(t)
        return t.value

    def parse(self) -> list:
        rules = []
        while True:
            t = self._opt()
            if t is None:
                break
            if t.kind == IMPORT:
                self._expect(";")
                sub = Parser(self.base_path, t.value, self.color_defs)
                rules.extend(sub.parse())
                self.color_defs.update(sub.color_defs)
            elif t.kind == COLORREF:
                self._expect(":")
                v = self._must()
                self._expect(";")
                if v.kind == COLOR:
                    self.color_defs[t.value] = v.value
            else:
                rules.append(self._rule(t))
        return rules

    def _rule(self, start: Token) -> Rule:
        rule = Rule([], [])
        while True:
            if start.kind == SIMPLE and start.value == "{":
                break
            if start.kind == IDENT and start.value == "colors":
                while True:
                    t = self._must()
                    if t.kind == SIMPLE and t.value == "}":
                        break
                return rule
            sel, more = self._selector(start)
            rule.selectors.append(sel)
            if not more:
                break
            start = self._must()
        rule.properties = self._properties()
        return rule

    def _selector(self, first: Token):
        if first.kind != IDENT:
            raise self._unexpected(first)
        ot = _OBJECT_TYPES.get(first.value)
        if ot is None:
            raise self._perr(f"Unknown object type: {first.value}", first.pos)
        sel = Selector(ot)
        while True:
            t = self._must()
            if t.kind == SIMPLE and t.value == "{":
                return sel, False
            if t.kind == SIMPLE and t.value == ",":
                return sel, True
            if t.kind == ZOOM:
                sel.min_zoom, sel.max_zoom = t.value
            elif t.kind == SIMPLE and t.value == "[":
                sel.tests.append(self._test())
            elif t.kind == SIMPLE and t.value == ":":
                self._ident()  # pseudo-class: parsed and ignored (parser.rs:340-344)
            elif t.kind == SIMPLE and t.value == "::":
                sel.layer_id = self._ident()
            else:
                raise self._unexpected(t)

    def _test(self) -> Test:
        bang = False
        t = self._must()
        if t.kind in (IDENT, STRING):
            lhs = t.value
        elif t.kind == SIMPLE and t.value == "!":
            bang = True
            lhs = self._ident()
        else:
            raise self._unexpected(t)
        t = self._must()
        if t.kind == SIMPLE and t.value == ":":
            lhs = lhs + ":" + self._ident()
            t = self._must()
        if not bang:
            if t.kind == SIMPLE and t.value in ("=", "!="):
                op = t.value
                t = self._must()
                if t.kind == IDENT:
                    rhs = t.value
                elif t.kind == NUMBER:
                    rhs = fmt_f64(t.value)
                else:
                    raise self._unexpected(t)
                self._expect("]")
                return Test("str", lhs, op, rhs)
            if t.kind == SIMPLE and t.value in ("<", "<=", ">", ">="):
                op = t.value
                t = self._must()
                if t.kind != NUMBER:
                    raise self._unexpected(t)
                self._expect("]")
                return Test("num", lhs, op, t.value)
        if t.kind == SIMPLE and t.value == "]":
            return Test("unary", lhs, "not_exists" if bang else "exists")
        if t.kind == SIMPLE and t.value == "?":
            t = self._must()
            if t.kind == SIMPLE and t.value == "]":
                return Test("unary", lhs, "false" if bang else "true")
            if t.kind == SIMPLE and t.value == "!" and not bang:
                self._expect("]")
                return Test("unary", lhs, "false")
            raise self._unexpected(t)
        raise self._unexpected(t)

    def _properties(self) -> list:
        props = []
        while True:
            t = self._must()
            if t.kind == IDENT:
                self._expect(":")
                kind, value = self._property_value()
                props.append(Property(t.value, kind, value))
            elif t.kind == SIMPLE and t.value == "}":
                return props
            else:
                raise self._unexpected(t)

    def _property_value(self):
        t = self._must()
        if t.kind == IDENT:
            if t.value == "eval":
                return self._simple_eval(t.pos)
            full = t.value
            n = self._must()
            if n.kind == SIMPLE and n.value == ":":
                full = full + ":" + self._ident()
                self._expect(";")
            elif n.kind == SIMPLE and n.value == ";":
                pass
            else:
                raise self._unexpected(n)
            return "ident", full
        if t.kind == STRING:
            self._expect(";")
            return "string", t.value
        if t.kind == COLOR:
            self._expect(";")
            return "color", t.value
        if t.kind == COLORREF:
            c = self.color_defs.get(t.value)
            if c is None:
                raise self._perr(f"Unknown color reference: {t.value}", (self.tok.line, self.tok.col))
            self._expect(";")
            return "color", c
        if t.kind == NUMBER:
            nums = [t.value]
            consumed = True
            while True:
                n = self._must()
                if n.kind == SIMPLE and n.value == "," and consumed:
                    consumed = False
                elif n.kind == SIMPLE and n.value == ";" and consumed:
                    break
                elif n.kind == NUMBER and not consumed:
                    consumed = True
                    nums.append(n.value)
                else:
                    raise self._unexpected(n)
            return "numbers", nums
        raise self._unexpected(t)

    def _simple_eval(self, pos):
        toks = []
        while True:
            t = self._must()
            if t.kind == SIMPLE and t.value == ";":
                break
            toks.append((t.kind, t.value))
        prefix = [(SIMPLE, "("), (IDENT, "prop"), (SIMPLE, "("), (STRING, "width"), (SIMPLE, ")")]
        inc = None
        if toks[: len(prefix)] == prefix:
            suffix = toks[len(prefix) :]
            if suffix and suffix[-1] == (SIMPLE, ")"):
                if len(suffix) == 1:
                    inc = 0.0
                elif len(suffix) == 2 and suffix[0][0] == NUMBER:
                    inc = suffix[0][1]
        if inc is None:
            raise self._perr("Unknown eval(...) form", pos)
        return "width_delta", inc


def parse_file(base_path: str, file_name: str) -> list:
    """parser.rs:223-232."""
    return Parser(base_path, file_name).parse()


def rules_to_string(rules: list) -> str:
    return "\n\n".join(str(r) for r in rules)


# ------------------------------------------------------------------------------------------------
# (de)serialisation of parsed rules: lets the benchmark style synthetic data on a machine that does not have
# the reference's stylesheet files (tests/golden/*_rules.json.gz are produced by tools/make_fixtures.py).
# ------------------------------------------------------------------------------------------------
def rules_to_json(rules: list) -> list:
    out = []
    for r in rules:
        out.append(
            {
                "selectors": [
                    {
                        "object_type": s.object_type,
                        "min_zoom": s.min_zoom,
                        "max_zoom": s.max_zoom,
                        "layer_id": s.layer_id,
                        "tests": [{"kind": t.kind, "tag": t.tag, "op": t.op, "value": t.value} for t in s.tests],
                    }
                    for s in r.selectors
                ],
                "properties": [
                    {"name": p.name, "kind": p.kind, "value": list(p.value) if isinstance(p.value, (tuple, list)) else p.value}
                    for p in r.properties
                ],
            }
        )
    return out


def rules_from_json(data: list) -> list:
    rules = []
    for r in data:
        sels = [
            Selector(s["object_type"], s["min_zoom"], s["max_zoom"], [Test(t["kind"], t["tag"], t["op"], t["value"]) for t in s["tests"]], s["layer_id"])
            for s in r["selectors"]
        ]
        props = []
        for p in r["properties"]:
            v = p["value"]
            if p["kind"] == "color":
                v = tuple(v)
            props.append(Property(p["name"], p["kind"], v))
        rules.append(Rule(sels, props))
    return rules


def load_rules_json(path: str) -> list:
    import gzip
    import json

    with gzip.open(path, "rt", encoding="utf-8") as f:
        return rules_from_json(json.load(f))


def save_rules_json(rules: list, path: str):
    import gzip
    import json

    with gzip.open(path, "wt", encoding="utf-8") as f:
        json.dump(rules_to_json(rules), f)
