"""Host-side model of the GICP dense sorted grid with a CLAMPED region (geoflowslam_b200/csrc/gicp.cu: struct DGrid, dense_coord,
make_shell_query(DGrid), visit_shell(DGrid), visit_ball(DGrid), grid_knn).

The CUDA kernels are checked against the hash grid on the GPU (tests/test_gpu_gicp.py::test_dense_grid_with_a_clamped_region).  This
file checks the ARGUMENT they rest on, without a GPU: when points and queries outside the region are clamped into its border cells,
the cell lower bounds (axis_gap2), the shell termination bound (r * cell + margin) and the ball's clamped bounding block still never
prune a point that matters -- the searches stay exact.  The model follows the kernels statement for statement (same shrink factors),
the truth is brute force under the kernels' total order (distance, index).
"""
import numpy as np

OFF = 1 << 20


def dense_coord(v, inv, lo, hi):
    t = min(max(v * inv, -2.0e6), 2.0e6)
    c = int(np.floor(t)) + OFF
    return min(max(c, lo), hi)


class DenseGrid:
    def __init__(self, pts, cell, axis_cap):
        self.pts, self.cell = pts, cell
        inv = 1.0 / cell
        cells = np.array([[dense_coord(p[k], inv, 0, (1 << 21) - 1) for k in range(3)] for p in pts])
        lo, hi = cells.min(0), cells.max(0)
        mean = cells.sum(0) // len(pts)
        for k in range(3):                                   # k_dense_region
            if hi[k] - lo[k] + 1 > axis_cap:
                l = min(max(mean[k] - axis_cap // 2, lo[k]), hi[k] - axis_cap + 1)
                lo[k], hi[k] = l, l + axis_cap - 1
        self.lo, self.hi = lo, hi
        self.cells = {}
        for i, p in enumerate(pts):                          # k_dense_count / fill (records of a cell: any order)
            key = tuple(dense_coord(p[k], inv, lo[k], hi[k]) for k in range(3))
            self.cells.setdefault(key, []).append(i)

    def query(self, q):                                      # make_shell_query(DGrid)
        inv = 1.0 / self.cell
        c = [dense_coord(q[k], inv, self.lo[k], self.hi[k]) for k in range(3)]
        f = [q[k] - (c[k] - OFF) * self.cell for k in range(3)]
        m = min(min(f[k], self.cell - f[k]) for k in range(3))
        return c, f, max(m, 0.0) * 0.999999

    def run(self, y, z, xa, xb):                             # dense_run: the records of cells xa..xb of row (y, z)
        out = []
        for x in range(xa, xb + 1):
            out += self.cells.get((x, y, z), [])
        return out


def axis_gap2(d, f, cell):
    if d == 0:
        return 0.0
    g = d * cell - f if d > 0 else f - (d + 1) * cell
    gg = max(g * 0.999999 - 1e-12, 0.0)
    return gg * gg


def knn(grid, q, k, max_shell=8):
    """grid_knn<K, DGrid>: shells until the k-th distance is provably final; None = the kernel would brute-force."""
    c, f, margin = grid.query(q)
    best = []   # sorted list of (d2, index)

    def limit():
        return best[k - 1][0] if len(best) >= k else np.inf

    def visit(ids):
        nonlocal best
        for i in ids:
            d = grid.pts[i] - q
            best.append((float((d[0] * d[0] + d[2] * d[2]) + d[1] * d[1]), i))
        best = sorted(best)[:k]

    for r in range(max_shell + 1):
        x0, x1 = max(c[0] - r, grid.lo[0]), min(c[0] + r, grid.hi[0])
        y0, y1 = max(c[1] - r, grid.lo[1]), min(c[1] + r, grid.hi[1])
        z0, z1 = max(c[2] - r, grid.lo[2]), min(c[2] + r, grid.hi[2])
        for z in range(z0, z1 + 1):
            gz2 = axis_gap2(z - c[2], f[2], grid.cell)
            if gz2 > limit():
                continue
            for y in range(y0, y1 + 1):
                gyz2 = gz2 + axis_gap2(y - c[1], f[1], grid.cell)
                if gyz2 > limit():
                    continue
                face = z in (c[2] - r, c[2] + r) or y in (c[1] - r, c[1] + r)
                if face:
                    xa, xb = x0, x1
                    while xa < c[0] and gyz2 + axis_gap2(xa - c[0], f[0], grid.cell) > limit():
                        xa += 1
                    while xb > c[0] and gyz2 + axis_gap2(xb - c[0], f[0], grid.cell) > limit():
                        xb -= 1
                    visit(grid.run(y, z, xa, xb))
                else:
                    if c[0] - r >= x0 and not gyz2 + axis_gap2(-r, f[0], grid.cell) > limit():
                        visit(grid.run(y, z, c[0] - r, c[0] - r))
                    if r > 0 and c[0] + r <= x1 and not gyz2 + axis_gap2(r, f[0], grid.cell) > limit():
                        visit(grid.run(y, z, c[0] + r, c[0] + r))
        bound = r * grid.cell + margin
        if len(best) >= k and best[k - 1][0] <= bound * bound:
            return best
        if all(c[a] - r <= grid.lo[a] and c[a] + r >= grid.hi[a] for a in range(3)):
            return best
    return None


def nn1_ball(grid, q, cap2):
    """nn1_ball_search over visit_ball<false>(DGrid): home cell, then the rows of the ball's clamped bounding block."""
    c, f, margin = grid.query(q)
    best = (np.inf, 0x7fffffff)

    def push(ids):
        nonlocal best
        for i in ids:
            d = grid.pts[i] - q
            best = min(best, (float((d[0] * d[0] + d[2] * d[2]) + d[1] * d[1]), i))

    push(grid.run(c[1], c[2], c[0], c[0]))
    lim0 = min(best[0], cap2)
    if lim0 <= margin * margin:
        return best
    R = np.sqrt(lim0) * 1.000001 + 1e-12
    inv = 1.0 / grid.cell

    def span(a):
        lo = c[a] + int(np.floor(min(max((f[a] - R) * inv, -3e6), 3e6)))
        hi = c[a] + int(np.floor(min(max((f[a] + R) * inv, -3e6), 3e6)))
        return min(max(lo, grid.lo[a]), grid.hi[a]), min(max(hi, grid.lo[a]), grid.hi[a])

    (x0, x1), (y0, y1), (z0, z1) = span(0), span(1), span(2)
    for z in range(z0, z1 + 1):
        gz2 = axis_gap2(z - c[2], f[2], grid.cell)
        if gz2 > min(best[0], cap2):
            continue
        for y in range(y0, y1 + 1):
            gyz2 = gz2 + axis_gap2(y - c[1], f[1], grid.cell)
            if gyz2 > min(best[0], cap2):
                continue
            xa, xb = x0, x1
            while xa < c[0] and gyz2 + axis_gap2(xa - c[0], f[0], grid.cell) > min(best[0], cap2):
                xa += 1
            while xb > c[0] and gyz2 + axis_gap2(xb - c[0], f[0], grid.cell) > min(best[0], cap2):
                xb -= 1
            if y == c[1] and z == c[2]:
                if xa < c[0]:
                    push(grid.run(y, z, xa, min(c[0] - 1, xb)))
                if xb > c[0]:
                    push(grid.run(y, z, max(c[0] + 1, xa), xb))
            else:
                push(grid.run(y, z, xa, xb))
    return best


def _cloud(seed, n=700, n_far=40):
    rng = np.random.default_rng(seed)
    # two planes of a room corner, 2.5 cm spacing with jitter, plus outliers tens of metres away (some of them in tight clusters)
    u, v = rng.uniform(0, 1.2, n // 2), rng.uniform(0, 1.0, n // 2)
    a = np.c_[u, v, 2.0 + 0.002 * rng.standard_normal(n // 2)]
    b = np.c_[1.2 + 0.002 * rng.standard_normal(n // 2), v, 2.0 - u]
    far = rng.uniform(-40, 40, (n_far, 3))
    far[n_far // 2:] = far[: n_far - n_far // 2] + rng.uniform(-0.03, 0.03, (n_far - n_far // 2, 3))
    return np.vstack([a, b, far])


def _brute_knn(pts, q, k):
    d = pts - q
    d2 = (d[:, 0] * d[:, 0] + d[:, 2] * d[:, 2]) + d[:, 1] * d[:, 1]
    order = np.lexsort((np.arange(len(pts)), d2))[:k]
    return [(float(d2[i]), int(i)) for i in order]


def test_knn_on_a_clamped_grid_is_exact():
    for seed, axis_cap in ((1, 6), (2, 4), (3, 1000)):      # 6 / 4 cells per axis: most of the room is clamped; 1000: nothing is
        pts = _cloud(seed)
        grid = DenseGrid(pts, 0.11, axis_cap)
        assert np.all(grid.hi - grid.lo + 1 <= max(axis_cap, 1)) or axis_cap == 1000
        finished = 0
        for i in range(0, len(pts), 3):
            got = knn(grid, pts[i], 10)
            if got is None:                                  # shells ran out: the kernel hands the query to the brute-force path
                continue
            finished += 1
            assert got == _brute_knn(pts, pts[i], 10), (seed, axis_cap, i)
        assert finished > 50


def test_bounded_nn_on_a_clamped_grid_is_exact():
    rng = np.random.default_rng(9)
    for seed, axis_cap in ((4, 5), (5, 3), (6, 1000)):
        pts = _cloud(seed)
        grid = DenseGrid(pts, 0.11, axis_cap)
        # queries: the cloud moved by a few centimetres (the optimiser's situation), points far outside every region, points next to
        # the far clusters
        qs = np.vstack([pts[::4] + rng.uniform(-0.04, 0.04, (len(pts[::4]), 3)), rng.uniform(-60, 60, (40, 3)),
                        pts[-20:] + rng.uniform(-0.05, 0.05, (20, 3))])
        cap2 = 0.1 * 0.1 * 1.0000001
        for q in qs:
            d2, idx = nn1_ball(grid, q, cap2)
            t2, tidx = _brute_knn(pts, q, 1)[0]
            if t2 > 0.1 * 0.1:                               # DistanceRejector: no correspondence -- the search may stop anywhere beyond
                assert not (d2 <= 0.1 * 0.1), (seed, axis_cap)
            else:
                assert (d2, idx) == (t2, tidx), (seed, axis_cap)
