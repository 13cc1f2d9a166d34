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


# ------------------------------------------------------------------------------------------------------------------
# Properties behind the per-dataset entity boxes, the device-side dedup and the parallel CRC (round 1 additions)
# ------------------------------------------------------------------------------------------------------------------
def _project(m, dim, t256, scale):
    """project_point of osmr_device.cuh for one coordinate (numpy float64 = the device's IEEE operations)."""
    x = m * dim - t256
    r = np.round(np.abs(x * scale)) * np.sign(x * scale)  # f64::round: half away from zero (np.round is half to even)
    frac = np.abs(x * scale) - np.floor(np.abs(x * scale))
    r = np.where(frac == 0.5, (np.floor(np.abs(x * scale)) + 1.0) * np.sign(x * scale), r)
    r = np.where(np.isnan(r), 0.0, r)
    return np.clip(r, -2147483648.0, 2147483647.0).astype(np.int64)


def test_pixel_bbox_of_an_entity_is_the_projection_of_its_mercator_box():
    """entity_pixel_bbox: project_point is monotone per coordinate, so min / max commute with it -- for every zoom, tile
    and scale, including coordinates that land exactly on .5 ties and far outside the tile."""
    rng = np.random.default_rng(11)
    for _ in range(300):
        zoom = int(rng.integers(0, 19))
        scale = float(rng.integers(1, 9))
        dim = float(256 * (1 << zoom))
        tx = int(rng.integers(0, 1 << zoom))
        t256 = float((tx * 256) & 0xFFFFFFFF)
        k = int(rng.integers(1, 40))
        centre = (tx + rng.random()) / (1 << zoom)
        m = centre + rng.normal(0.0, 2.0 ** -rng.integers(8, 30), size=k)
        # put some coordinates exactly on half pixels
        ties = (np.floor(m[: k // 3] * dim * scale) + 0.5) / (dim * scale)
        m[: k // 3] = ties
        px = _project(m, dim, t256, scale)
        lo = _project(np.array([m.min()]), dim, t256, scale)[0]
        hi = _project(np.array([m.max()]), dim, t256, scale)[0]
        assert px.min() == lo and px.max() == hi


def test_ownership_rule_emits_every_listed_entity_exactly_once():
    """osmr_auto.cuh: an entity listed in every index tile of its rectangle is emitted by exactly one record of a query
    rectangle -- the one at (max(ent.min_x, rect.min_x), max(ent.min_y, rect.min_y)) -- iff the rectangles intersect."""
    rng = random.Random(5)
    for _ in range(2000):
        ex0, ey0 = rng.randrange(0, 40), rng.randrange(0, 40)
        ex1, ey1 = ex0 + rng.randrange(0, 6), ey0 + rng.randrange(0, 6)
        xa, ya = rng.randrange(0, 40), rng.randrange(0, 40)
        xb, yb = xa + rng.randrange(0, 12), ya + rng.randrange(0, 12)
        owners = 0
        for x in range(max(ex0, xa), min(ex1, xb) + 1):  # the index records inside the query that list the entity
            for y in range(max(ey0, ya), min(ey1, yb) + 1):
                if x == max(ex0, xa) and y == max(ey0, ya):
                    owners += 1
        intersects = ex0 <= xb and ex1 >= xa and ey0 <= yb and ey1 >= ya
        assert owners == (1 if intersects else 0)


def test_crc32_slices_combine_like_png_finish_kernel():
    """png_finish_kernel: a raw CRC (zero start, no final inversion) ignores leading zeros and combines by multiplication
    with x^(8n) mod P; the standard CRC-32 is raw ^ (all-ones pushed through n bytes) ^ all-ones (zlib's crc32_combine
    arithmetic in the reflected domain)."""
    import zlib

    POLY = 0xEDB88320

    def multmodp(a, b):
        m, p = 1 << 31, 0
        while True:
            if a & m:
                p ^= b
                if (a & (m - 1)) == 0:
                    break
            m >>= 1
            b = (b >> 1) ^ POLY if b & 1 else b >> 1
        return p

    x2n = [1 << 30]
    for _ in range(31):
        x2n.append(multmodp(x2n[-1], x2n[-1]))

    def x2nmodp(n, k):
        p = 1 << 31
        while n:
            if n & 1:
                p = multmodp(x2n[k & 31], p)
            n >>= 1
            k += 1
        return p

    table = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ POLY if c & 1 else c >> 1
        table.append(c)

    def raw(data):
        c = 0
        for byte in data:
            c = table[(c ^ byte) & 0xFF] ^ (c >> 8)
        return c

    rng = random.Random(3)
    for n in (1, 5, 255, 256, 257, 1000, 4099):
        msg = bytes(rng.randrange(256) for _ in range(n))
        threads = 16
        lc = (n + threads - 1) // threads
        pad = lc * threads - n
        framed = b"\0" * pad + msg  # right-aligned frame: the leading zeros do not change a raw CRC
        part = [raw(framed[t * lc : (t + 1) * lc]) for t in range(threads)]
        mul = x2nmodp(lc, 3)
        stride = 1
        while stride < threads:  # the combine tree of the kernel
            for t in range(0, threads, 2 * stride):
                part[t] = multmodp(mul, part[t]) ^ part[t + stride]
            mul = multmodp(mul, mul)
            stride *= 2
        crc = part[0] ^ multmodp(x2nmodp(n, 3), 0xFFFFFFFF) ^ 0xFFFFFFFF
        assert crc == zlib.crc32(msg)
