"""CPU: the closed forms the CUDA kernels use instead of the reference's step-by-step walks, restated in Python and
pinned against (a) the oracle's step-by-step fill edge walk and (b) a literal transcription of line.rs:65-158.

These are the formulas of osm_renderer_b200/csrc/osmr_device.cuh (fill_edge_row_span, ncorr) and the order-free
even-odd rule of fill_rows_kernel."""
import random

import numpy as np

import oracle


def fdiv(a, b):
    return a // b


def cdiv(a, b):
    return -((-a) // b)


def fill_edge_row_span(x1, y1, x2, y2, y):
    a, b = abs(x2 - x1), abs(y2 - y1)
    sx = 1 if x1 < x2 else -1
    sy = 1 if y1 < y2 else -1
    j = (y - y1) * sy
    if j < 0 or j > b:
        return None
    if b == 0:
        return (min(x1, x2), max(x1, x2), True)

    def hi(jj):
        if jj == b:
            return a
        return min(a, max(cdiv(a - 2 * b + 2 * jj * a, 2 * b), 0))

    if j == 0:
        lo = 0
    else:
        h = hi(j - 1)
        X = fdiv(2 * a - b + 2 * (j - 1) * a, 2 * b)
        lo = min(a, h + (1 if h <= X else 0))
    e = max(lo, hi(j))
    xa, xb = x1 + sx * lo, x1 + sx * e
    return (min(xa, xb), max(xa, xb), (j == 0 and y1 <= y2) or (j == b and y2 <= y1))


def test_fill_row_span_closed_form_equals_bresenham_walk():
    rng = random.Random(7)
    edges = [(3, 5, 3 + dx, 5 + dy) for dx in range(-24, 25) for dy in range(-24, 25)]
    edges += [tuple(rng.randint(-9000, 9000) for _ in range(4)) for _ in range(150)]
    for (x1, y1, x2, y2) in edges:
        lo, hi = min(y1, y2) - 2, max(y1, y2) + 2
        rows = oracle.fill_edge_rows(x1, y1, x2, y2, lo, hi)
        for y in range(lo, hi + 1):
            want = rows[y - lo]
            got = fill_edge_row_span(x1, y1, x2, y2, y)
            if want[2] == 0:
                assert got is None
            else:
                assert got == (int(want[0]), int(want[1]), want[2] == 2), (x1, y1, x2, y2, y)


def test_order_free_even_odd_equals_sorted_pairing():
    """covered(x) = [c(x) odd and c(x) < m] or [x in span(e), rank(e) odd]  ==  fill.rs:23-46 pairing."""
    rng = random.Random(11)
    D = 64
    for _ in range(3000):
        m = rng.randint(0, 9)
        spans = []
        for _ in range(m):
            a = rng.randint(-10, D + 10)
            spans.append((a, a + rng.choice([0, 0, 1, 2, 5, 20])))
        order = sorted(range(m), key=lambda i: (spans[i][0], i))
        want = np.zeros(D, dtype=bool)
        for q in range(0, m - 1, 2):
            f, t = max(spans[order[q]][0], 0), min(spans[order[q + 1]][1], D - 1)
            if f <= t:
                want[f : t + 1] = True
        rank = {e: r for r, e in enumerate(order)}
        got = np.zeros(D, dtype=bool)
        for x in range(D):
            c = sum(1 for s in spans if s[0] <= x)
            if (c % 2 == 1 and c < m) or any(rank[i] % 2 == 1 and spans[i][0] <= x <= spans[i][1] for i in range(m)):
                got[x] = True
        assert (got == want).all(), spans


def _line_walks_reference(p1, p2, T):
    (x1, y1), (x2, y2) = p1, p2
    inc = lambda f, t: 1 if f <= t else -1  # noqa: E731
    dx, dy = abs(x2 - x1), abs(y2 - y1)
    swap = dx > dy
    sw = lambda a, b: (b, a) if swap else (a, b)  # noqa: E731
    mn, mx = sw(x1, y1)
    mn_last, mx_last = sw(x2, y2)
    mn_d, mx_d = sw(dx, dy)
    mn_inc, mx_inc = sw(inc(x1, x2), inc(y1, y2))

    def upd(e):
        c = False
        if e + 2 * mn_d > mx_d:
            e -= 2 * mx_d
            c = True
        return e + 2 * mn_d, c

    out = []

    def perps(mn, mx, pe):
        for mul in (1, -1):
            p_mn, p_mx, e, pix = mx, mn, mul * pe, []
            for _ in range(T):
                pix.append(sw(p_mx, p_mn))
                e, c = upd(e)
                if c:
                    p_mn -= mul * mx_inc
                p_mx += mul * mn_inc
            out.append((mul, tuple(pix)))

    error = p_error = 0
    while True:
        perps(mn, mx, p_error)
        if mn == mn_last and mx == mx_last:
            break
        error, c = upd(error)
        if c:
            mn += mn_inc
            p_error, c2 = upd(p_error)
            if c2:
                perps(mn, mx, p_error)
        mx += mx_inc
    return out


def _line_walks_closed(p1, p2, T):
    (x1, y1), (x2, y2) = p1, p2
    inc = lambda f, t: 1 if f <= t else -1  # noqa: E731
    dx, dy = abs(x2 - x1), abs(y2 - y1)
    swap = dx > dy
    sw = lambda a, b: (b, a) if swap else (a, b)  # noqa: E731
    mn0, mx0 = sw(x1, y1)
    mn_d, mx_d = sw(dx, dy)
    mn_inc, mx_inc = sw(inc(x1, x2), inc(y1, y2))
    magic = (2**64 - 1) // (2 * mx_d) + 1

    def ncorr(e0, n):
        num = e0 + 2 * mn_d * n - mx_d
        if num <= 0:
            return 0
        q = ((num + 2 * mx_d - 1) * magic) >> 64  # __umul64hi with the per-segment magic
        assert q == -((-num) // (2 * mx_d))
        return q

    out = []

    def walk(mn, mx, pe):
        for mul in (1, -1):
            pix = []
            e, p_mn, p_mx = mul * pe, mx, mn
            for _ in range(T):
                pix.append(sw(p_mx, p_mn))
                if e + 2 * mn_d > mx_d:
                    e -= 2 * mx_d
                    p_mn -= mul * mx_inc
                e += 2 * mn_d
                p_mx += mul * mn_inc
            out.append((mul, tuple(pix)))

    for k in range(mx_d + 1):
        c = ncorr(0, k)
        pc = ncorr(0, c)
        pe = 2 * mn_d * c - 2 * mx_d * pc
        walk(mn0 + mn_inc * c, mx0 + mx_inc * k, pe)
        e_main = 2 * mn_d * k - 2 * mx_d * c
        if k < mx_d and e_main + 2 * mn_d > mx_d and pe + 2 * mn_d > mx_d:
            walk(mn0 + mn_inc * (c + 1), mx0 + mx_inc * k, pe - 2 * mx_d + 2 * mn_d)
    return out


def test_thick_line_random_access_equals_sequential_traversal():
    rng = random.Random(5)
    cases = [((2, -3), (2 + dx, -3 + dy)) for dx in range(-20, 21) for dy in range(-20, 21) if dx or dy]
    cases += [((rng.randint(-3000, 3000), rng.randint(-3000, 3000)), (rng.randint(-3000, 3000), rng.randint(-3000, 3000))) for _ in range(120)]
    for p1, p2 in cases:
        if p1 == p2:
            continue
        assert _line_walks_reference(p1, p2, 7) == _line_walks_closed(p1, p2, 7), (p1, p2)
