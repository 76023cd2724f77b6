"""CPU model of the q8 candidate tier's proof obligation (csrc/scan_topk.cu: stb_q8_build_kernel,
stb_scan_q8): the score the int8 scan ranks by must be an UPPER BOUND of the exact cosine,
u >= c - 1e-5, for every row and query -- that is what lets the completeness check compare the k-th
exact distance with the best dropped score.  The kernels' arithmetic is restated here
in numpy float32 / int32 (same operation order where it matters) and checked against the f64 cosine on
random and adversarial inputs.  No GPU, no oracle: this pins the BOUND, the GPU tests pin the kernels."""
import numpy as np

F = np.float32


def build_q8(rows):
    """stb_q8_build_kernel: x^ = x * rsqrt(sum x^2) in f32, s = max|x^| / 127, codes = rint(x^ * (127 / max))."""
    rows = rows.astype(F)
    ss = (rows * rows).sum(axis=1, dtype=F)
    with np.errstate(divide="ignore"):
        inv = np.where(ss > 0, F(1) / np.sqrt(ss, dtype=F), F(0)).astype(F)
    xh = (rows * inv[:, None]).astype(F)
    am = np.abs(xh).max(axis=1).astype(F)
    s = (am * F(1.0 / 127.0)).astype(F)
    with np.errstate(divide="ignore"):
        inv_s = np.where(am > 0, F(127.0) / am, F(0)).astype(F)
    codes = np.clip(np.rint((xh * inv_s[:, None]).astype(F)), -127, 127).astype(np.int32)
    return codes, s


def q8_scores(codes, s, q):
    """stb_scan_q8: quantised query (16 bits per component), exact int32 dot, upper-bound score in f32."""
    q = q.astype(F)
    b2 = (q * q).sum(dtype=F)
    rq = F(1) / np.sqrt(b2, dtype=F)
    amax = F(np.abs(q).max() * rq)
    S = F(32639.0) / amax
    qs = F(rq * S)
    q16 = np.clip(np.rint((q * qs).astype(F)), -32639, 32639).astype(np.int64)
    l1 = int(np.abs(q16).sum())
    inv_S = F(1.0) / S
    h_l1 = F(F(0.50025) * F(l1) * inv_S)
    e_q = F(F(9.7) * inv_S)
    dot = codes.astype(np.int64) @ q16
    assert np.abs(dot).max() < 2 ** 31                       # the kernel accumulates in int32
    inner = (dot.astype(F) * inv_S + h_l1).astype(F)
    return (s * inner + e_q).astype(F), q16, float(S)


def exact_cos(rows, q):
    r, qq = rows.astype(np.float64), q.astype(np.float64)
    n = np.sqrt((r * r).sum(axis=1)) * np.sqrt((qq * qq).sum())
    return np.divide(r @ qq, n, out=np.zeros(len(r)), where=n > 0)


def check(rows, q):
    codes, s = build_q8(rows)
    u, _, _ = q8_scores(codes, s, q)
    c = exact_cos(rows, q)
    slack = u.astype(np.float64) - c
    assert slack.min() >= -1e-5, (slack.min(), int(slack.argmin()))
    return slack


def unit(rng, n):
    x = rng.standard_normal((n, 256)).astype(F)
    return (x / np.linalg.norm(x, axis=1, keepdims=True)).astype(F)


def test_upper_bound_holds_on_random_unit_rows_and_is_not_wasteful():
    rng = np.random.default_rng(1)
    rows = unit(rng, 20000)
    for _ in range(8):
        slack = check(rows, unit(rng, 1)[0])
        # typical slack ~ s/2 * ||q~||_1 ~ 0.011: the bound must stay useful (K' = 128 needs ~0.03 of room)
        assert 0.005 < np.median(slack) < 0.02 and slack.max() < 0.04


def test_upper_bound_holds_on_scaled_rows_and_scaled_queries():
    rng = np.random.default_rng(2)
    rows = (unit(rng, 5000) * rng.uniform(1e-3, 1e3, (5000, 1))).astype(F)
    for scale in (1e-4, 1.0, 37.5, 1e4):
        check(rows, (unit(rng, 1)[0] * F(scale)).astype(F))


def test_upper_bound_holds_on_adversarial_rows():
    rng = np.random.default_rng(3)
    n = 4000
    rows = unit(rng, n)
    rows[:500, 0] += F(3.0)                                   # one dominant component: large per-row scale
    # components parked just below / above the rounding boundaries of the row's own grid
    base = unit(rng, 500)
    am = np.abs(base).max(axis=1, keepdims=True)
    grid = am / 127.0
    rows[500:1000] = ((np.rint(base / grid) + rng.choice([-0.4999, 0.4999], base.shape)) * grid).astype(F)
    rows[1000:1100] = 0.0                                     # zero rows: scale 0, bound = the query's slack
    rows[1100:1200] *= F(1e-12)
    sparse = np.zeros((300, 256), dtype=F)                    # one-hot and two-hot rows
    sparse[np.arange(300), rng.integers(0, 256, 300)] = 1.0
    sparse[np.arange(300), rng.integers(0, 256, 300)] += F(0.5)
    rows[1200:1500] = sparse
    queries = [unit(rng, 1)[0] for _ in range(4)]
    spike = unit(rng, 1)[0]; spike[7] = 40.0                  # dominant query component: coarse grid for the others
    queries.append(spike.astype(F))
    queries.append(np.sign(unit(rng, 1)[0]).astype(F))        # all components +-1: ||q^||_1 = 16, the maximum
    onehot = np.zeros(256, dtype=F); onehot[3] = 1.0
    queries.append(onehot)
    queries.append(rows[0].copy())
    queries.append((-rows[0]).astype(F))
    for q in queries:
        check(rows, q)


def test_query_quantisation_splits_into_two_signed_bytes_exactly():
    """q16 = 256 * hi + lo with hi, lo in [-128, 127] (two dp4a chains): exact for the whole range."""
    v = np.arange(-32639, 32640, dtype=np.int64)
    lo = ((v + 128) & 255) - 128
    hi = (v - lo) >> 8
    assert np.all(256 * hi + lo == v) and lo.min() >= -128 and lo.max() <= 127 and hi.min() >= -128 and hi.max() <= 127
    # worst-case |dot| fits int32: 256 components x 127 x 32639
    assert 256 * 127 * 32639 < 2 ** 31
