"""NumPy float32 model of the walk kernel's FP32 fate filter (csrc/cmt_device.cuh: filter_fate,
filter_input, draw_f32; plane table as cmt_beamline_create builds it in csrc/cmt_api.cu).

Test infrastructure only: it lets the CPU suite check the filter's error model against the oracle
(every fate the filter decides must be the oracle's fate, with the oracle's row count) without a GPU.
NumPy rounds every operation separately where the device fuses multiply-adds; the error bounds hold
for either, which is the point of the check."""
from __future__ import annotations

import numpy as np

from oracle import oracle

F = np.float32
U = F(2.0 ** -24)
MAX_PLANES = 32
CIRCLE, BOX = 0, 1


def filter_planes(flat, g=oracle.G):
    """-> (planes, covers_all) with planes = list of dicts, or ([], False) when the filter does not apply."""
    planes, usable = [], np.isfinite(g) and abs(g) < 1e6

    def ordinary(v):
        return np.isfinite(v) and abs(v) < 1e15

    def edge_ok(v):
        return not np.isnan(v) and (np.isinf(v) or abs(v) < 1e15)

    def circle(z, R, fate):
        nonlocal usable
        T = R * R
        if not (ordinary(z) and np.isfinite(T) and 1e-30 <= T <= 1e30):
            usable = False
            return
        a = F(T)
        planes.append(dict(z=F(z), kind=CIRCLE, a=a, b=F(0), c=F(0), d=F(0), fate=fate,
                           tol=F(np.ldexp(abs(a), -21)) + F(1e-37)))

    def box(z, x1, x2, y1, y2, fate):
        nonlocal usable
        if not (ordinary(z) and all(edge_ok(v) for v in (x1, x2, y1, y2))):
            usable = False
            return
        e = [F(x1), F(x2), F(y1), F(y2)]
        m = max([abs(v) for v in e if np.isfinite(v)] + [F(0)])
        planes.append(dict(z=F(z), kind=BOX, a=e[0], b=e[1], c=e[2], d=e[3], fate=fate,
                           tol=F(np.ldexp(m, -21)) + F(1e-37)))

    els = flat.elements
    e = 0
    while usable and e < len(els) and len(planes) + 2 <= MAX_PLANES:
        t = els[e]
        if t["type"] == oracle.CIRCULAR:
            circle(t["z0"], t["R"], int(t["fate"]))
            if usable:
                circle(t["z1"], t["R"], int(t["fate"]))
        elif t["type"] == oracle.RECTANGULAR:
            for z in (t["z0"], t["z1"]):
                if usable:
                    box(z, t["x1"], t["x2"], t["y1"], t["y2"], int(t["fate"]))
        elif t["type"] == oracle.FIELDPLATES:
            for z in (t["z0"], t["z1"]):
                if usable:
                    box(z, t["x1"], t["x2"], -np.inf, np.inf, int(t["fate"]))
        else:
            break
        e += 1
    lens = [i for i, t in enumerate(els) if t["type"] == oracle.LENS]
    first_lens = lens[0] if lens else len(els)
    covers_all = False
    if usable and e < len(els) and e == first_lens and len(planes) + 1 <= MAX_PLANES:
        circle(els[e]["z0"], els[e]["R"], int(els[e]["fate"]))
    elif usable and e == len(els):
        covers_all = True
    if not usable:
        return [], False
    return planes, covers_all


def filter_input(ic):
    with np.errstate(over="ignore"):
        q = {k: ic[i].astype(F) for i, k in enumerate(("x0", "y0", "z0", "vx", "vy", "vz"))}
    for k, src in (("ex0", "x0"), ("ey0", "y0"), ("evx", "vx"), ("evy", "vy"), ("evz", "vz")):
        q[k] = U * np.abs(q[src])
    return q


def filter_fate(planes, covers_all, fate_detected, q, g=oracle.G):
    """-> (fate, rows): fate -1 where undecided."""
    n = q["x0"].shape[0]
    fate = np.full(n, -1, dtype=np.int32)
    rows = np.zeros(n, dtype=np.int32)
    hg, g_abs = F(0.5 * g), F(abs(g))
    with np.errstate(all="ignore"):
        inv = F(1) / q["vz"]
        ainv = np.abs(inv)
        relv = q["evz"] * ainv + F(4) * U
        c0 = F(2) * (q["ex0"] + q["ey0"] + F(2) * U * (np.abs(q["x0"]) + np.abs(q["y0"])))
        c1 = F(2) * ((np.abs(q["vx"]) + np.abs(q["vy"])) * (relv + F(6) * U) + (q["evx"] + q["evy"]))
        c2 = F(4) * g_abs * (relv + F(3) * U)
        az0 = np.abs(q["z0"])
        big = np.abs(q["x0"]) + np.abs(q["y0"]) + az0 + np.abs(q["vx"]) + np.abs(q["vy"]) + np.abs(q["vz"])
        live = (big < F(1e15)) & (ainv < F(1e15))        # still passing every plane so far
        for p, pl in enumerate(planes):
            dt = (pl["z"] - q["z0"]) * inv
            dtE = ainv * (np.abs(pl["z"]) + az0)
            x = q["vx"] * dt + q["x0"]
            y = (-hg * dt + q["vy"]) * dt + q["y0"]
            E = (c2 * dtE + c1) * dtE + c0
            if pl["kind"] == CIRCLE:
                s = x * x + y * y
                tol = F(2) * (np.abs(x) + np.abs(y) + E) * E + (s * (F(8) * U) + pl["tol"])
                d = s - pl["a"]
                clear = np.abs(d) > tol
                dead, ok = clear & (d > 0), clear & (d < 0)
            else:
                Eb = E + pl["tol"]
                mx = np.fmin(x - pl["a"], pl["b"] - x)
                my = np.fmin(y - pl["c"], pl["d"] - y)
                sane = (x == x) & (y == y)
                ok = sane & (mx > Eb) & (my > Eb)
                dead = sane & ((-mx > Eb) | (-my > Eb))
            hit = live & dead
            fate[hit] = pl["fate"]
            rows[hit] = p + 1
            live = live & ok & ~dead
        if covers_all:
            fate[live] = fate_detected
            rows[live] = len(planes)
    return fate, rows


# ---- the source in single precision (draw_f32) ----
def _philox(index, block, seed):
    """Philox4x32-10, vectorised: counter (index_lo, index_hi, block, 0), key = seed."""
    M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    c0 = (index & np.uint64(0xFFFFFFFF)).astype(np.uint64)
    c1 = (index >> np.uint64(32)).astype(np.uint64)
    c2 = np.full_like(c0, block)
    c3 = np.zeros_like(c0)
    k0, k1 = np.uint64(seed & 0xFFFFFFFF), np.uint64(seed >> 32)
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        h0, l0, h1, l1 = p0 >> np.uint64(32), p0 & mask, p1 >> np.uint64(32), p1 & mask
        c0, c1, c2, c3 = h1 ^ c1 ^ k0, l1, h0 ^ c3 ^ k1, l0
        k0 = (k0 + np.uint64(0x9E3779B9)) & mask
        k1 = (k1 + np.uint64(0xBB67AE85)) & mask
    return [c.astype(np.uint32) for c in (c0, c1, c2, c3)]


def _unit(lo, hi):
    return hi.astype(F) * F(2.0 ** -32) + ((lo & np.uint32(0xFFFFF800)).astype(F) * F(2.0 ** -64) + F(2.0 ** -54))


def _cossin(hi, noise):
    a = F(6.2831855) * ((hi ^ np.uint32(0x80000000)).view(np.int32).astype(F) * F(2.0 ** -32))
    # the device uses __cosf/__sinf (2^-21.41 absolute on [-pi, pi]); model that error explicitly
    e = F(2.0 ** -21.41)
    return -(np.cos(a) + noise[0] * e), -(np.sin(a) + noise[1] * e)


def _box_muller(w, noise):
    u = _unit(w[0], w[1])
    lnu = np.log(u)
    # __logf: 2^-21.41 absolute on [0.5, 2], 3 ulp elsewhere
    lnu = lnu + noise[2] * np.where(u >= F(0.5), F(2.0 ** -21.41), F(3) * np.spacing(np.abs(lnu)))
    with np.errstate(all="ignore"):
        L = F(-2) * lnu
        rs = (F(1) / np.sqrt(L)) * (F(1) + noise[3] * F(2.0 ** -22))     # rsqrt.approx: 2 ulp
        R = L * rs
        c, s = _cossin(w[3], noise)
    return R * c, R * s, F(2.0 ** -18) * (F(2) * R + rs)


def draw_f32(source, seed, first, n, rng=None):
    """-> filter input dict.  rng: worst-case-magnitude noise on the fast intrinsics (None = accurate)."""
    s = source[0]
    index = np.arange(first, first + n, dtype=np.uint64)

    def noise():
        if rng is None:
            return [F(0)] * 4
        return [rng.choice(np.array([-1, 1], dtype=F), n) for _ in range(4)]

    q = {}
    sx, sy, sz = (F(v) for v in s["vsigma"])
    mx, my, mz = (F(v) for v in s["vmean"])
    n0, n1, en = _box_muller(_philox(index, 0, seed), noise())
    q["vx"], q["vy"] = sx * n0 + mx, sy * n1 + my
    q["evx"] = abs(sx) * en + F(2.0 ** -22) * (abs(mx) + np.abs(sx * n0))
    q["evy"] = abs(sy) * en + F(2.0 ** -22) * (abs(my) + np.abs(sy * n1))
    n0, n1, en = _box_muller(_philox(index, 1, seed), noise())
    q["vz"] = sz * n0 + mz
    q["evz"] = abs(sz) * en + F(2.0 ** -22) * (abs(mz) + np.abs(sz * n0))
    w = _philox(index, 2, seed)
    if int(s["pos_kind"]) == 0:
        c, sn = _cossin(w[1], noise())
        r = np.sqrt(_unit(w[2], w[3])) * F(s["p0"])
        q["x0"], q["y0"] = r * c, r * sn
        q["ex0"] = q["ey0"] = F(2.0 ** -18) * np.abs(r)
    else:
        n0, n1, en = _box_muller(w, noise())
        p0, p1 = F(s["p0"]), F(s["p1"])
        q["x0"], q["y0"] = p0 * n0, p1 * n1
        q["ex0"] = abs(p0) * en + F(2.0 ** -22) * np.abs(q["x0"])
        q["ey0"] = abs(p1) * en + F(2.0 ** -22) * np.abs(q["y0"])
    q["z0"] = np.full(n, F(s["z"]), dtype=F)
    return q


# ---- adversarial inputs: molecules aimed at the edges ----
def aimed_ics(flat, n, rng, scale, base):
    """Initial conditions `base` (6, n) with the transverse velocity re-aimed so that the exact
    parabola meets a random filter plane at relative distance ~N(0, scale) from its edge."""
    planes = []           # (z, kind, R, x1, x2, y1, y2) in binary64
    for t in flat.elements:
        if t["type"] == oracle.CIRCULAR:
            planes += [(t["z0"], 0, t["R"], 0, 0, 0, 0), (t["z1"], 0, t["R"], 0, 0, 0, 0)]
        elif t["type"] == oracle.RECTANGULAR:
            planes += [(z, 1, 0, t["x1"], t["x2"], t["y1"], t["y2"]) for z in (t["z0"], t["z1"])]
        elif t["type"] == oracle.FIELDPLATES:
            planes += [(z, 2, 0, t["x1"], t["x2"], 0, 0) for z in (t["z0"], t["z1"])]
        else:
            if t["type"] == oracle.LENS:
                planes.append((t["z0"], 0, t["R"], 0, 0, 0, 0))
            break
    pl = np.array(planes, dtype=np.float64)
    ic = np.array(base, dtype=np.float64, copy=True)
    p = rng.integers(0, len(pl), n)
    dt = (pl[p, 0] - ic[2]) / ic[5]
    delta = 1 + scale * rng.standard_normal(n)
    th = rng.uniform(0, 2 * np.pi, n)
    side = rng.integers(0, 2, n)
    kind = pl[p, 1]
    free = rng.uniform(-0.004, 0.004, n)
    on_y = (kind == 1) & (rng.integers(0, 2, n) == 1)
    xt = np.where(kind == 0, pl[p, 2] * delta * np.cos(th),
                  np.where(on_y, free, np.where(side == 0, pl[p, 3], pl[p, 4]) * delta))
    yt = np.where(kind == 0, pl[p, 2] * delta * np.sin(th),
                  np.where(on_y, np.where(side == 0, pl[p, 5], pl[p, 6]) * delta, free))
    ic[3] = (xt - ic[0]) / dt
    ic[4] = (yt - ic[1] + 0.5 * oracle.G * dt * dt) / dt
    return ic


# ---- the filter with constant thresholds (quick_fate / build_quick) ----
def _down(x):
    f = F(x)
    return np.nextafter(f, F(-np.inf)) if float(f) > x else f


def _up(x):
    f = F(x)
    return np.nextafter(f, F(np.inf)) if float(f) < x else f


def quick_table(flat, source=None, g=oracle.G):
    """-> dict(guards..., planes=[...]) or None when the thresholds cannot be used (cmt_api.cu: build_quick)."""
    planes, _ = filter_planes(flat, g)
    if not planes:
        return None
    # binary64 plane data, in the order filter_planes lists them
    pd = []
    for t in flat.elements:
        if t["type"] == oracle.CIRCULAR:
            pd += [(t["z0"], CIRCLE, t["R"] ** 2, None), (t["z1"], CIRCLE, t["R"] ** 2, None)]
        elif t["type"] == oracle.RECTANGULAR:
            pd += [(z, BOX, 0.0, (t["x1"], t["x2"], t["y1"], t["y2"])) for z in (t["z0"], t["z1"])]
        elif t["type"] == oracle.FIELDPLATES:
            pd += [(z, BOX, 0.0, (t["x1"], t["x2"], -np.inf, np.inf)) for z in (t["z0"], t["z1"])]
        else:
            if t["type"] == oracle.LENS:
                pd.append((t["z0"], CIRCLE, t["R"] ** 2, None))
            break
    pd = pd[:len(planes)]
    u = 2.0 ** -24
    lxy = max([np.sqrt(T) if k == CIRCLE else max(abs(v) for v in e if np.isfinite(v)) for _, k, T, e in pd])
    zmin, zmax = min(abs(z) for z, *_ in pd), max(abs(z) for z, *_ in pd)
    pos_g, ang_g = 4 * lxy, 2.0
    if source is None:
        z0_g, ex_g, eva_g, relvz_g = 2 * zmin, 1.01 * u * pos_g, 1.01 * u * ang_g, 1.01 * u
    else:
        s = source[0]
        en, k, r22 = 2.0 ** -18 * (2 * 1.25 + 0.8), 4.0, 2.0 ** -22
        sxy, sz, mz = abs(s["vsigma"][0]) + abs(s["vsigma"][1]), abs(s["vsigma"][2]), abs(s["vmean"][2])
        vref = max(mz - 2 * sz, 0.5 * mz)
        if not vref > 0:
            return None
        z0_g = abs(s["z"]) * (1 + 4 * u) + 1e-30
        eva_g = k * (sxy * en + r22 * (abs(s["vmean"][0]) + abs(s["vmean"][1]) + 1.25 * sxy)) / vref
        relvz_g = k * (sz * en + r22 * (mz + 1.25 * sz)) / vref
        ex_g = k * 2 * 2.0 ** -18 * abs(s["p0"]) if int(s["pos_kind"]) == 0 else k * (abs(s["p0"]) + abs(s["p1"])) * (en + 1.25 * r22)
    zsum = zmax + z0_g
    vmin = 8 * np.sqrt(2 * abs(g) * zsum)
    ainv_g = 1 / vmin if vmin > 1e-12 else 1e12
    relv = relvz_g * 1.001 + 4 * u
    k0, k1 = 2 * (ex_g + 2 * u * pos_g), 2 * (ang_g * (relv + 8 * u) + eva_g)
    k2 = 4 * abs(g) * (relv + 4 * u) * ainv_g ** 2
    out = []
    for (z, kind, T, e), pl in zip(pd, planes):
        Z = abs(z) + z0_g
        eps = (k0 + k1 * Z + k2 * Z * Z) * (1 + 2.0 ** -10)
        if kind == CIRCLE:
            def B(s_):
                return 2 * (np.sqrt(2 * s_ * (1 + 4 * u)) + eps) * eps + 8 * u * s_ + 1e-37
            T_lo, T_hi = T - B(T), T + B(T)
            for _ in range(6):
                T_hi = T + B(T_hi)
            T_hi *= 1 + 2.0 ** -20
            if not (T_lo >= 0.5 * T and T_hi >= 16 * eps * eps and T_hi <= 2 * T):
                return None
            v = [_down(T_lo), _up(T_hi)] + [F(0)] * 6
        else:
            m = max(abs(x) for x in e if np.isfinite(x))
            eb = eps + np.ldexp(m, -21) + 1e-37
            v = [_up(e[0] + eb), _down(e[1] - eb), _up(e[2] + eb), _down(e[3] - eb),
                 _down(e[0] - eb), _up(e[1] + eb), _down(e[2] - eb), _up(e[3] + eb)]
            if not (v[0] < v[1] and v[2] < v[3]):
                return None
        out.append(dict(z=F(z), kind=kind, fate=pl["fate"], v=v))
    return dict(pos_g=_down(pos_g), z0_g=_down(z0_g), ainv_g=_down(ainv_g), ang_g=_down(ang_g), eva_g=_down(eva_g),
                relvz_g=_down(relvz_g), ex_g=_down(ex_g), planes=out)


def quick_fate(table, covers_all, fate_detected, q, g=oracle.G):
    n = q["x0"].shape[0]
    fate = np.full(n, -1, dtype=np.int32)
    rows = np.zeros(n, dtype=np.int32)
    hg = F(0.5 * g)
    with np.errstate(all="ignore"):
        inv = F(1) / q["vz"]
        ainv = np.abs(inv)
        live = ((np.abs(q["x0"]) + np.abs(q["y0"]) <= table["pos_g"]) & (np.abs(q["z0"]) <= table["z0_g"])
                & (ainv <= table["ainv_g"]) & ((np.abs(q["vx"]) + np.abs(q["vy"])) * ainv <= table["ang_g"])
                & ((q["evx"] + q["evy"]) * ainv <= table["eva_g"]) & (q["evz"] * ainv <= table["relvz_g"])
                & (q["ex0"] + q["ey0"] <= table["ex_g"]))
        guarded = live.copy()
        sx, sy, qg = q["vx"] * inv, q["vy"] * inv, hg * inv * inv
        bx = -sx * q["z0"] + q["x0"]
        B = F(2) * qg * q["z0"] + sy
        A = -q["z0"] * (qg * q["z0"] + sy) + q["y0"]
        for p, pl in enumerate(table["planes"]):
            x = sx * pl["z"] + bx
            y = (-qg * pl["z"] + B) * pl["z"] + A
            v = pl["v"]
            if pl["kind"] == CIRCLE:
                s = x * x + y * y
                dead, ok = s > v[1], s < v[0]
            else:
                ok = (x > v[0]) & (x < v[1]) & (y > v[2]) & (y < v[3])
                dead = (x < v[4]) | (x > v[5]) | (y < v[6]) | (y > v[7])
            hit = live & dead
            fate[hit] = pl["fate"]
            rows[hit] = p + 1
            live = live & ok & ~dead
        if covers_all:
            fate[live] = fate_detected
            rows[live] = len(table["planes"])
    return fate, rows, guarded
