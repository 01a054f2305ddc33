"""Synthetic FITS-like radio maps for the benchmark and the tests (SURVEY.md §8d
config #2): per image i, numpy.random.default_rng(1234 + i); S x S float32; Gaussian noise
sigma = 3e-4 Jy; 20-60 point sources with log-uniform peak 1e-3..5e-2 Jy convolved with an
elliptical Gaussian beam (FWHM 3-6 px); 0-3 extended double-lobe sources; an 8-px NaN strip on 25 %
of the images (exercises the NaN -> min fill)."""
import numpy as np


def _gauss2d(S, y0, x0, fwhm_y, fwhm_x, theta, peak, yy, xx):
    sy, sx = fwhm_y / 2.3548, fwhm_x / 2.3548
    ct, st = np.cos(theta), np.sin(theta)
    dy, dx = yy - y0, xx - x0
    u = ct * dx + st * dy
    v = -st * dx + ct * dy
    return peak * np.exp(-0.5 * ((u / sx) ** 2 + (v / sy) ** 2))


def radio_map(i, S=256):
    rng = np.random.default_rng(1234 + i)
    img = rng.normal(0.0, 3e-4, size=(S, S))
    yy, xx = np.mgrid[0:S, 0:S].astype(np.float64)
    bmaj, bmin, bpa = rng.uniform(3, 6), rng.uniform(3, 6), rng.uniform(0, np.pi)
    for _ in range(int(rng.integers(20, 61))):
        peak = 10 ** rng.uniform(-3, np.log10(5e-2))
        y0, x0 = rng.uniform(0, S, 2)
        r = int(4 * max(bmaj, bmin))
        ys, ye = max(0, int(y0) - r), min(S, int(y0) + r + 1)
        xs, xe = max(0, int(x0) - r), min(S, int(x0) + r + 1)
        img[ys:ye, xs:xe] += _gauss2d(S, y0, x0, bmaj, bmin, bpa, peak, yy[ys:ye, xs:xe], xx[ys:ye, xs:xe])
    for _ in range(int(rng.integers(0, 4))):
        y0, x0 = rng.uniform(0.15 * S, 0.85 * S, 2)
        sep, ang = rng.uniform(6, 0.12 * S), rng.uniform(0, np.pi)
        peak = 10 ** rng.uniform(-2.5, -1.5)
        for sgn in (-1, 1):
            cy, cx = y0 + sgn * sep * np.sin(ang), x0 + sgn * sep * np.cos(ang)
            img += _gauss2d(S, cy, cx, rng.uniform(6, 14), rng.uniform(4, 8), ang, peak, yy, xx)
    img = img.astype(np.float32)
    if rng.random() < 0.25:
        k = int(rng.integers(0, 4))
        if k == 0:
            img[:8, :] = np.nan
        elif k == 1:
            img[-8:, :] = np.nan
        elif k == 2:
            img[:, :8] = np.nan
        else:
            img[:, -8:] = np.nan
    return img


def radio_maps(n, S=256, start=0):
    return np.stack([radio_map(start + i, S) for i in range(n)])
