"""TEST INFRASTRUCTURE ONLY -- CPU oracle (numpy) for the Remhos RK-stage hot path.

Problem definitions restated from remhos.cpp: velocity_function (:2001-2120),
u0_function (:2201-2355), inflow_function (:2363-2381).  Vectorised over points:
x is [npts, dim]; bb_min/bb_max is the mesh bounding box (remhos.cpp:457).
"""
import numpy as np
from scipy.special import erfc


def _normalise(x, bb_min, bb_max):
    # remhos.cpp:2005-2011
    center = (bb_min + bb_max) * 0.5
    return 2.0 * (x - center) / (bb_max - bb_min)


def velocity(problem, x, bb_min, bb_max):
    x = np.asarray(x, dtype=np.float64)
    n, dim = x.shape
    X = _normalise(x, bb_min, bb_max)
    v = np.zeros_like(x)
    pe = problem % 20
    if pe == 0:                                              # :2018-2028
        if dim == 1:
            v[:, 0] = 1.0
        elif dim == 2:
            v[:, 0] = np.sqrt(2. / 3.); v[:, 1] = np.sqrt(1. / 3.)
        else:
            v[:, 0] = np.sqrt(3. / 6.); v[:, 1] = np.sqrt(2. / 6.); v[:, 2] = np.sqrt(1. / 6.)
    elif pe in (1, 2, 4):                                    # :2029-2042
        w = np.pi / 2
        if dim == 1:
            v[:, 0] = 1.0
        else:
            v[:, 0] = -w * X[:, 1]; v[:, 1] = w * X[:, 0]
    elif pe == 3:                                            # :2043-2056
        w = np.pi / 2
        d = np.maximum((X[:, 0] + 1.) * (1. - X[:, 0]), 0.) * \
            np.maximum((X[:, 1] + 1.) * (1. - X[:, 1]), 0.)
        d = d * d
        if dim == 1:
            v[:, 0] = 1.0
        else:
            v[:, 0] = d * w * X[:, 1]; v[:, 1] = -d * w * X[:, 0]
    elif pe == 5:                                            # :2057-2066
        v[:, :] = 1.0
    elif pe in (6, 7):                                       # :2067-2077
        if dim == 1:
            v[:, 0] = 1.0
        else:
            v[:, 0] = x[:, 1]; v[:, 1] = -x[:, 0]
    elif pe == 11:                                           # :2078-2094 (Gresho)
        r = np.sqrt(x[:, 0] ** 2 + x[:, 1] ** 2)
        with np.errstate(divide='ignore', invalid='ignore'):
            a0 = np.where(r < 0.2, 5.0 * x[:, 1],
                          np.where(r < 0.4, 2.0 * x[:, 1] / r - 5.0 * x[:, 1], 0.0))
            a1 = np.where(r < 0.2, -5.0 * x[:, 0],
                          np.where(r < 0.4, -2.0 * x[:, 0] / r + 5.0 * x[:, 0], 0.0))
        v[:, 0] = a0; v[:, 1] = a1
    elif pe in (10, 12, 13, 14, 15, 16, 17):                 # :2095-2117 (Taylor-Green)
        Y = X * 0.5 + 0.5
        v[:, 0] = np.sin(np.pi * Y[:, 0]) * np.cos(np.pi * Y[:, 1])
        v[:, 1] = -np.cos(np.pi * Y[:, 0]) * np.sin(np.pi * Y[:, 1])
        if dim == 3:
            v[:, 0] *= np.cos(np.pi * Y[:, 2])
            v[:, 1] *= np.cos(np.pi * Y[:, 2])
            v[:, 2] = 0.0
    else:
        raise ValueError('unknown problem %d' % problem)
    return v


def _box(p1, p2, theta, origin, x, y):                       # :2122-2148
    s = np.sin(theta * np.pi / 180); c = np.cos(theta * np.pi / 180)
    ox, oy = origin
    xn = c * (x - ox) - s * (y - oy) + ox
    yn = s * (x - ox) + c * (y - oy) + oy
    return ((xn > p1[0]) & (xn < p2[0]) & (yn > p1[1]) & (yn < p2[1])).astype(np.float64)


def _box3d(xmin, xmax, ymin, ymax, zmin, zmax, theta, ox, oy, x, y, z):   # :2150-2169
    s = np.sin(theta * np.pi / 180); c = np.cos(theta * np.pi / 180)
    xn = c * (x - ox) - s * (y - oy) + ox
    yn = s * (x - ox) + c * (y - oy) + oy
    return ((xn > xmin) & (xn < xmax) & (yn > ymin) & (yn < ymax) &
            (z > zmin) & (z < zmax)).astype(np.float64)


def _cross(r1, r2):                                          # :2171-2175
    return r1 + r2 - r1 * r2


def _ring(rin, rout, c, y):                                  # :2177-2198
    r = np.sqrt(((y - np.asarray(c)) ** 2).sum(axis=1))
    return ((r > rin) & (r < rout)).astype(np.float64)


def u0(problem, x, bb_min, bb_max):
    x = np.asarray(x, dtype=np.float64)
    n, dim = x.shape
    X = _normalise(x, bb_min, bb_max)
    pe = problem % 10
    if pe in (0, 1):                                         # :2217-2238
        if dim == 1:
            return np.exp(-40. * (X[:, 0] - 0.5) ** 2)
        rx, ry, cx, cy, w = 0.45, 0.25, 0., -0.2, 10.
        if dim == 3:
            s = (1. + 0.25 * np.cos(2 * np.pi * X[:, 2]))
            rx = rx * s
            ry = ry * s
        return (erfc(w * (X[:, 0] - cx - rx)) * erfc(-w * (X[:, 0] - cx + rx)) *
                erfc(w * (X[:, 1] - cy - ry)) * erfc(-w * (X[:, 1] - cy + ry))) / 16
    if pe == 2:                                              # :2239-2245
        rho = np.hypot(X[:, 0], X[:, 1]); phi = np.arctan2(X[:, 1], X[:, 0])
        return np.sin(np.pi * rho) ** 2 * np.sin(3 * phi)
    if pe == 3:                                              # :2246-2249
        return .5 * (np.sin(np.pi * X[:, 0]) * np.sin(np.pi * X[:, 1]) + 1.)
    if pe == 4:                                              # :2250-2262 (operator precedence
        scale = 0.0225                                       #  of ?: kept literally)
        coef = 0.5 / np.sqrt(scale)
        slit = (X[:, 0] <= -0.05) | (X[:, 0] >= 0.05) | (X[:, 1] >= 0.7)
        cone = coef * np.sqrt(X[:, 0] ** 2 + (X[:, 1] + 0.5) ** 2)
        hump = coef * np.sqrt((X[:, 0] + 0.5) ** 2 + X[:, 1] ** 2)
        cond = slit & ((X[:, 0] ** 2 + (X[:, 1] - .5) ** 2) <= 4. * scale)
        other = (0. + (1. - cone) * ((X[:, 0] ** 2 + (X[:, 1] + .5) ** 2) <= 4. * scale)
                 + .25 * (1. + np.cos(np.pi * hump))
                 * (((X[:, 0] + .5) ** 2 + X[:, 1] ** 2) <= 4. * scale))
        return np.where(cond, 1., other)
    if pe == 5:                                              # :2263-2338
        y = 50. * (x + 1.)
        if dim == 2:
            origin = (15.5, 11.5)
            rect1 = _box((14., 3.), (17., 26.), -45., origin, y[:, 0], y[:, 1])
            rect2 = _box((7., 10.), (32., 13.), -45., origin, y[:, 0], y[:, 1])
            cross = _cross(rect1, rect2)
            ring1 = _ring(7., 10., [40., 40.], y)
            ring2 = _ring(3., 7., [40., 20.], y)
            return cross + ring1 + ring2
        rect1 = _box3d(7., 32., 10., 13., 10., 13., -45., 15.5, 11.5, y[:, 0], y[:, 1], y[:, 2])
        rect2 = _box3d(14., 17., 3., 26., 10., 13., -45., 15.5, 11.5, y[:, 0], y[:, 1], y[:, 2])
        rect3 = _box3d(14., 17., 10., 13., 3., 26., -45., 15.5, 11.5, y[:, 0], y[:, 1], y[:, 2])
        cross = _cross(_cross(rect1, rect2), rect3)
        c1 = [40., 40., 40.]; c2 = [40., 20., 20.]
        dom2 = cross + _ring(7., 10., c1, y) + _ring(3., 7., c2, y)
        rect1 = _box3d(2., 27., 30., 33., 30., 33., 0., 0., 0., y[:, 0], y[:, 1], y[:, 2])
        rect2 = _box3d(9., 12., 23., 46., 30., 33., 0., 0., 0., y[:, 0], y[:, 1], y[:, 2])
        rect3 = _box3d(9., 12., 30., 33., 23., 46., 0., 0., 0., y[:, 0], y[:, 1], y[:, 2])
        cross = _cross(_cross(rect1, rect2), rect3)
        dom3 = cross + _ring(0., 7., c1, y) + _ring(0., 3., c2, y) + _ring(7., 10., c2, y)
        dom1 = 1. - _cross(dom2, dom3)
        return dom1 + 2. * dom2 + 3. * dom3
    if pe == 6:                                              # :2339-2348
        r = np.sqrt((x ** 2).sum(axis=1))
        return np.where((r >= 0.15) & (r < 0.45), 1.,
                        np.where((r >= 0.55) & (r < 0.85),
                                 np.cos(10. * np.pi * (r - 0.7) / 3.) ** 2, 0.))
    if pe == 7:                                              # :2349-2354
        r = np.sqrt((x ** 2).sum(axis=1))
        a, b, c = 0.5, 3.e-2, 0.1
        return 0.25 * (1. + np.tanh((r + c - a) / b)) * (1. - np.tanh((r - c - a) / b))
    return np.zeros(n)


def inflow(problem, x):                                      # :2363-2381
    x = np.asarray(x, dtype=np.float64)
    r = np.sqrt((x ** 2).sum(axis=1))
    if (problem % 10) == 6 and x.shape[1] == 2:
        return np.where((r >= 0.15) & (r < 0.45), 1.,
                        np.where((r >= 0.55) & (r < 0.85),
                                 np.cos(10. * np.pi * (r - 0.7) / 3.) ** 2, 0.))
    if (problem % 10) == 7:
        a, b, c = 0.5, 3.e-2, 0.1
        return 0.25 * (1. + np.tanh((r + c - a) / b)) * (1. - np.tanh((r - c - a) / b))
    return np.zeros(x.shape[0])
