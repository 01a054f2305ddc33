"""ORACLE (test infrastructure) — restatement of skimage.measure.find_contours as scikit-image 0.15 computes it
(third-party, absent from /root/reference and from this image; call sites: mrcnn/analyze.py:1914 and mrcnn/sfinder.py:897,
`find_contours(padded_mask, 0.5)` with the defaults fully_connected='low', positive_orientation='low').

parity: UNPINNED — scikit-image is not installable here, so this follows the published algorithm
(skimage/measure/_find_contours.py + _find_contours_cy.pyx, v0.15): a 2x2 square marched over the array in raster order,
one or two oriented segments per square case, linear interpolation of the crossing along each square edge, then
`_assemble_contours`: segments are joined through `starts` / `ends` dictionaries in the order they were produced; when two
open contours meet the one created FIRST survives; the result is ordered by creation index.
"""
from collections import deque

import numpy as np


def _fraction(from_value, to_value, level):
    if to_value == from_value:
        return 0
    return (level - from_value) / (to_value - from_value)


def marching_segments(array, level, vertex_connect_high=False):
    """iterate_and_store of _find_contours_cy.pyx: flat list of points (from, to, from, to, ...)"""
    array = np.asarray(array, dtype=np.double)
    if array.shape[0] < 2 or array.shape[1] < 2:
        raise ValueError("Input array must be at least 2x2.")
    arc = []
    for r0 in range(array.shape[0] - 1):
        for c0 in range(array.shape[1] - 1):
            r1, c1 = r0 + 1, c0 + 1
            ul, ur, ll, lr = array[r0, c0], array[r0, c1], array[r1, c0], array[r1, c1]
            case = (1 if ul > level else 0) + (2 if ur > level else 0) + (4 if ll > level else 0) + (8 if lr > level else 0)
            if case in (0, 15):
                continue
            top = (r0, c0 + _fraction(ul, ur, level))
            bottom = (r1, c0 + _fraction(ll, lr, level))
            left = (r0 + _fraction(ul, ll, level), c0)
            right = (r0 + _fraction(ur, lr, level), c1)
            if case == 1:
                arc += [top, left]
            elif case == 2:
                arc += [right, top]
            elif case == 3:
                arc += [right, left]
            elif case == 4:
                arc += [left, bottom]
            elif case == 5:
                arc += [top, bottom]
            elif case == 6:
                arc += [left, top, right, bottom] if vertex_connect_high else [right, top, left, bottom]
            elif case == 7:
                arc += [right, bottom]
            elif case == 8:
                arc += [bottom, right]
            elif case == 9:
                arc += [top, right, bottom, left] if vertex_connect_high else [top, left, bottom, right]
            elif case == 10:
                arc += [bottom, top]
            elif case == 11:
                arc += [bottom, left]
            elif case == 12:
                arc += [left, right]
            elif case == 13:
                arc += [top, right]
            elif case == 14:
                arc += [left, top]
    return arc


def assemble_contours(points):
    """_assemble_contours of _find_contours.py over consecutive (from, to) pairs"""
    current_index = 0
    contours, starts, ends = {}, {}, {}
    for k in range(0, len(points), 2):
        from_point, to_point = points[k], points[k + 1]
        if from_point == to_point:
            continue
        tail_data = starts.get(to_point)
        head_data = ends.get(from_point)
        if tail_data is not None and head_data is not None:
            tail, tail_num = tail_data
            head, head_num = head_data
            if tail is head:                      # close the contour
                head.append(to_point)
                del starts[to_point]
                del ends[from_point]
            elif tail_num > head_num:             # tail was created second: append it to head
                head.extend(tail)
                del starts[to_point]
                try:
                    del ends[tail[-1]]
                except KeyError:
                    pass
                contours.pop(tail_num, None)
                del ends[from_point]
                ends[head[-1]] = (head, head_num)
            else:                                 # head was created second: prepend it to tail
                tail.extendleft(reversed(head))
                del starts[head[0]]
                del ends[from_point]
                contours.pop(head_num, None)
                del starts[to_point]
                starts[tail[0]] = (tail, tail_num)
        elif tail_data is None and head_data is None:
            current_index += 1
            new_contour = deque((from_point, to_point))
            contours[current_index] = new_contour
            starts[from_point] = (new_contour, current_index)
            ends[to_point] = (new_contour, current_index)
        elif head_data is None:                   # prepend to the contour that starts at to_point
            tail, tail_num = tail_data
            tail.appendleft(from_point)
            del starts[to_point]
            starts[from_point] = (tail, tail_num)
        else:                                     # append to the contour that ends at from_point
            head, head_num = head_data
            head.append(to_point)
            del ends[from_point]
            ends[to_point] = (head, head_num)
    return [np.array(contour, dtype=np.float64) for (num, contour) in sorted(contours.items())]


def find_contours(array, level, fully_connected="low", positive_orientation="low"):
    """list of [n,2] float64 (row, column) arrays"""
    if fully_connected not in ("high", "low") or positive_orientation not in ("high", "low"):
        raise ValueError("parameters must be 'high' or 'low'")
    contours = assemble_contours(marching_segments(array, level, fully_connected == "high"))
    if positive_orientation == "high":
        contours = [c[::-1] for c in contours]
    return contours


def mask_vertexes(mask, xmin=0, ymin=0):
    """the `vertexes` entry of make_json_results (mrcnn/analyze.py:1908-1921): contours of the zero-padded mask at 0.5,
    as lists of [x, y] with the padding and the tile origin removed / added"""
    mask = np.asarray(mask)
    padded = np.zeros((mask.shape[0] + 2, mask.shape[1] + 2), dtype=np.uint8)
    padded[1:-1, 1:-1] = mask
    out = []
    for verts in find_contours(padded, 0.5):
        verts = np.fliplr(verts) - 1
        if xmin != 0 or ymin != 0:
            verts = verts + np.array([xmin, ymin])
        out.append(verts.tolist())
    return out
