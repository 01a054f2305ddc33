"""`vertexes` of the catalogue (mrcnn/analyze.py:1908-1927, mrcnn/sfinder.py:885-910 -> skimage.measure.find_contours):
the product's host-only C++ routine (mrcnn_host_contours through mrcnn.analyze.contours_of_pixel_lists) against the
oracle's pure-Python restatement of the scikit-image 0.15 algorithm (oracle/contours.py).  Both are restatements of an
absent third-party package: this pins them to each other and to the algorithm's invariants, not to scikit-image."""
import numpy as np

from mrcnn import analyze as P
from oracle import contours as C


def _pixels(mask, ymin=0, xmin=0):
    return (np.argwhere(mask == 1) + np.array([ymin, xmin])).astype(np.int32)


def test_single_pixel_and_diagonal_pixels_known_answers():
    m = np.zeros((5, 6), np.uint8)
    m[2, 3] = 1
    got = P.contours_of_pixel_lists([_pixels(m)])[0]
    assert got == C.mask_vertexes(m)
    assert len(got) == 1 and got[0][0] == got[0][-1] and len(got[0]) == 5          # a closed diamond around the pixel
    assert sorted(map(tuple, got[0][:-1])) == sorted([(3.0, 1.5), (2.5, 2.0), (3.0, 2.5), (3.5, 2.0)])
    # two pixels touching at a corner: fully_connected='low' keeps them apart -> two diamonds, in raster order of creation
    m = np.zeros((4, 4), np.uint8)
    m[1, 1] = m[2, 2] = 1
    got = P.contours_of_pixel_lists([_pixels(m)])[0]
    assert got == C.mask_vertexes(m) and len(got) == 2 and all(c[0] == c[-1] and len(c) == 5 for c in got)
    m = np.zeros((4, 4), np.uint8)
    m[1, 2] = m[2, 1] = 1
    got = P.contours_of_pixel_lists([_pixels(m)])[0]
    assert got == C.mask_vertexes(m) and len(got) == 2


def test_random_masks_match_oracle_and_invariants():
    rng = np.random.default_rng(4)
    cases = []
    for shape, density in (((12, 17), 0.5), ((20, 20), 0.15), ((9, 31), 0.8), ((16, 16), 1.0), ((1, 7), 0.6), ((7, 1), 1.0)):
        for _ in range(4):
            cases.append((rng.uniform(0, 1, shape) < density).astype(np.uint8))
    ring = np.zeros((11, 11), np.uint8)
    ring[2:9, 2:9] = 1
    ring[4:7, 4:7] = 0                                   # a hole: outer and inner contour
    cases.append(ring)
    blobs = np.zeros((30, 40), np.uint8)
    yy, xx = np.mgrid[0:30, 0:40]
    blobs[(yy - 10) ** 2 + (xx - 12) ** 2 < 40] = 1
    blobs[(yy - 20) ** 2 + (xx - 30) ** 2 < 30] = 1
    cases.append(blobs)
    cases = [m for m in cases if m.any()]
    origin = (100, 2000)                                  # tile origin (ymin, xmin) added to the pixel lists
    got_all = P.contours_of_pixel_lists([_pixels(m, *origin) for m in cases])      # one batched call
    for m, got in zip(cases, got_all):
        want = C.mask_vertexes(m, xmin=origin[1], ymin=origin[0])
        assert got == want
        assert got == P.contours_of_pixel_lists([_pixels(m, *origin)])[0]
        on = {(int(y) + origin[0], int(x) + origin[1]) for y, x in np.argwhere(m == 1)}
        for contour in got:
            assert contour[0] == contour[-1], "contours of a zero-padded mask are closed"
            for x, y in contour:
                assert (2 * x) % 1 == 0 and (2 * y) % 1 == 0 and ((x % 1 == 0.5) != (y % 1 == 0.5))
                # every vertex sits on the edge between a mask pixel and a background pixel
                if x % 1 == 0.5:
                    a, b = (int(y), int(x - 0.5)), (int(y), int(x + 0.5))
                else:
                    a, b = (int(y - 0.5), int(x)), (int(y + 0.5), int(x))
                assert (a in on) != (b in on)
    assert P.contours_of_pixel_lists([]) == [] and P.contours_of_pixel_lists([np.zeros((0, 2), np.int32)]) == [[]]


def test_general_find_contours_oracle_on_float_field():
    """the oracle's general (non-binary) path: interpolated crossings and orientation flags"""
    yy, xx = np.mgrid[0:8, 0:9].astype(float)
    f = np.exp(-((yy - 3.3) ** 2 + (xx - 4.1) ** 2) / 6.0)
    low = C.find_contours(f, 0.4)
    high = C.find_contours(f, 0.4, positive_orientation="high")
    assert len(low) == 1 and np.allclose(low[0][0], low[0][-1]) and np.array_equal(high[0], low[0][::-1])
    r, c = low[0][:, 0], low[0][:, 1]
    val = np.exp(-((r - 3.3) ** 2 + (c - 4.1) ** 2) / 6.0)
    assert np.abs(val - 0.4).max() < 0.03                 # linear interpolation lands close to the level set
