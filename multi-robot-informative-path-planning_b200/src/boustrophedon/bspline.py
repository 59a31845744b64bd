"""Clamped uniform B-spline sampling (mirror of the reference's src/boustrophedon/bspline.py:8-32)."""
import numpy as np
import scipy.interpolate as si


def bspline(cv, n=100, degree=3):
    """n samples of the degree-`degree` clamped B-spline through control vertices `cv`."""
    cv = np.asarray(cv)
    count = cv.shape[0]
    degree = np.clip(degree, 1, count - 1)
    knots = np.array([0] * degree + list(range(count - degree + 1)) + [count - degree] * degree, dtype="int")
    u = np.linspace(0, (count - degree), n)
    return np.array(si.splev(u, (knots, cv.T, degree))).T
