"""CPU oracle of the tile driver's bookkeeping and edge merging. TEST INFRASTRUCTURE ONLY.

Only tests/ may import this module; the product (caesar-mrcnn_b200/mrcnn/sfinder.py + csrc/analyze.cu) never does.

Restates:
  utils.generate_tiles                 /root/reference/mrcnn/utils.py:1254-1328
  TileTask.is_task_tile_adjacent / _overlapping / _neighbor     /root/reference/mrcnn/sfinder.py:119-159
  SFinder.create_tile_tasks (round-robin assignment, neighbour lists)   sfinder.py:1216-1289
  SFinder.find_sources_at_edge         sfinder.py:643-706
  SFinder.merge_edge_sources           sfinder.py:711-935 (all keys except "vertexes": skimage.measure.find_contours
                                       is absent from this image and not restated — PARITY UNPINNED for that key)

Pinned against the reference itself: tests/golden/make_golden_sfinder.py imports /root/reference/mrcnn/sfinder.py in
the build container (astropy / skimage / cv2 / regions / numpyencoder stubbed, find_contours -> no contours) and
stores inputs and outputs in tests/golden/sfinder_golden.json; tests/test_oracle_sfinder.py replays them here.

Reference quirks kept on purpose (they change results):
  * merge_edge_sources takes class / score of the LAST member of a merged group, not of the largest one: the lookup
    uses the loop variable `index` instead of `index_largest` (sfinder.py:862-864);
  * pixel adjacency is 8-connected (|dx| <= 1 and |dy| <= 1, sfinder.py:799), unlike the 4-connected mask test of
    the Analyzer;
  * merged bounding boxes are inclusive pixel extremes (x2 = max x), while tile sources carry exclusive x2 / y2.
"""
import numpy as np


def generate_tiles(img_xmin, img_xmax, img_ymin, img_ymax, tile_size_x, tile_size_y, grid_step_x, grid_step_y):
    """utils.py:1254-1328 -> list of (xmin, xmax, ymin, ymax) with exclusive maxima, row-major; None on bad input."""
    if img_xmax <= img_xmin or img_ymax <= img_ymin:
        return None
    if tile_size_x <= 0 or tile_size_y <= 0:
        return None
    if grid_step_x <= 0 or grid_step_y <= 0 or grid_step_x > 1 or grid_step_y > 1:
        return None
    nx = img_xmax - img_xmin + 1
    ny = img_ymax - img_ymin + 1
    if tile_size_x > nx or tile_size_y > ny:
        return None
    step_x = int(np.round(grid_step_x * tile_size_x))
    step_y = int(np.round(grid_step_y * tile_size_y))

    def axis(n, size, step):
        lo, hi, index = [], [], 0
        while index <= n:
            offset = min(size, n - index)
            if index >= n or offset == 0:
                break
            lo.append(index)
            hi.append(index + offset)
            index += step
        return lo, hi

    iy_min, iy_max = axis(ny, tile_size_y, step_y)
    ix_min, ix_max = axis(nx, tile_size_x, step_x)
    return [(img_xmin + ix_min[i], img_xmin + ix_max[i], img_ymin + iy_min[j], img_ymin + iy_max[j])
            for j in range(len(iy_min)) for i in range(len(ix_min))]


def tiles_are_neighbors(a, b):
    """sfinder.py:119-159 on (xmin, xmax, ymin, ymax) tuples."""
    adj_x = a[1] == b[0] - 1 or a[0] == b[1] + 1 or (a[0] == b[0] and a[1] == b[1])
    adj_y = a[3] == b[2] - 1 or a[2] == b[3] + 1 or (a[2] == b[2] and a[3] == b[3])
    overlapping = not (a[1] < b[0] or a[0] > b[1] or a[3] < b[2] or a[2] > b[3])
    return (adj_x and adj_y) or overlapping


def create_tile_tasks(tile_grid, nproc):
    """sfinder.py:1230-1289 -> tasks_per_worker: list (per worker) of dicts(tid, coords, neighborTaskId,
    neighborTaskIndex, neighborWorkerId), neighbour lists in the reference's insertion order."""
    workers = [[] for _ in range(nproc)]
    w = 0
    for tid, coords in enumerate(tile_grid):
        workers[w].append(dict(tid=tid, wid=w, coords=tuple(coords), neighborTaskId=[], neighborTaskIndex=[], neighborWorkerId=[]))
        w = 0 if w >= nproc - 1 else w + 1

    def link(t, tid, tindex, wid):
        t["neighborTaskId"].append(tid)
        t["neighborTaskIndex"].append(tindex)
        t["neighborWorkerId"].append(wid)

    for i in range(nproc):
        for j, task in enumerate(workers[i]):
            for k in range(j + 1, len(workers[i])):
                other = workers[i][k]
                if tiles_are_neighbors(task["coords"], other["coords"]):
                    link(task, other["tid"], k, i)
                    link(other, task["tid"], j, i)
            for s in range(i + 1, nproc):
                for t, other in enumerate(workers[s]):
                    if tiles_are_neighbors(task["coords"], other["coords"]):
                        link(task, other["tid"], t, s)
                        link(other, task["tid"], j, i)
    return workers


def find_sources_at_edge(objs, tile, neighbor_tiles):
    """sfinder.py:643-706: sets obj["edge"] = True for sources on the tile border or inside a neighbour tile's range.
    tile / neighbor_tiles: (xmin, xmax, ymin, ymax). Sources not matched keep the flag they already carry."""
    xmin, xmax, ymin, ymax = tile
    for src in objs:
        if src["x1"] == xmin or src["x2"] == xmax or src["y1"] == ymin or src["y2"] == ymax:
            src["edge"] = True
            continue
        for n in neighbor_tiles:
            if src["x2"] < n[0] or src["x1"] > n[1] or src["y2"] < n[2] or src["y1"] > n[3]:
                continue
            src["edge"] = True
            break


def pixels_adjacent(pixels_a, pixels_b):
    """sfinder.py:787-808: any pixel pair with |dx| <= 1 and |dy| <= 1 (vectorised form of the double loop)."""
    a = np.asarray(pixels_a, dtype=np.int64).reshape(-1, 2)
    b = np.asarray(pixels_b, dtype=np.int64).reshape(-1, 2)
    if not len(a) or not len(b):
        return False
    for lo in range(0, len(a), 512):
        d = np.abs(a[lo:lo + 512, None, :] - b[None, :, :])
        if np.any((d[:, :, 0] <= 1) & (d[:, :, 1] <= 1)):
            return True
    return False


def merge_edge_sources(tile_sources):
    """sfinder.py:711-935. tile_sources: list of per-tile dicts (objs, workerId, tileId, neighborTileIds, ...), the
    content of SFinder.tile_sources["sources"]. Returns the final source list (SFinder.sources["sources"])."""
    final, to_merge = [], []
    for tindex, tile in enumerate(tile_sources):
        for sindex, src in enumerate(tile["objs"]):
            if not src["edge"]:
                src["merged"] = False
                final.append(src)
            else:
                to_merge.append((sindex, tindex))

    n = len(to_merge)
    adj = [[] for _ in range(n)]
    for i in range(n):
        si, ti = to_merge[i]
        a = tile_sources[ti]["objs"][si]
        neighbors = tile_sources[ti]["neighborTileIds"]
        for j in range(i + 1, n):
            sj, tj = to_merge[j]
            b = tile_sources[tj]["objs"][sj]
            if tile_sources[tj]["tileId"] not in neighbors:
                continue
            if a["x2"] < b["x1"] or a["x1"] > b["x2"] or a["y2"] < b["y1"] or a["y1"] > b["y2"]:
                continue
            if not pixels_adjacent(a["pixels"], b["pixels"]):
                continue
            adj[i].append(j)
            adj[j].append(i)

    seen = [False] * n
    components = []

    def visit(v, acc):
        seen[v] = True
        acc.append(v)
        for u in adj[v]:
            if not seen[u]:
                visit(u, acc)

    for v in range(n):
        if not seen[v]:
            acc = []
            visit(v, acc)
            components.append(acc)

    for i, comp in enumerate(components):
        name = "S" + str(i + 1) + "_merged"
        if len(comp) == 1:
            si, ti = to_merge[comp[0]]
            src = tile_sources[ti]["objs"][si]
            src["name"] = name
            src["merged"] = False
            final.append(src)
            continue
        pixels_merged, npix_largest = [], -1
        for index in comp:
            si, ti = to_merge[index]
            pixels = tile_sources[ti]["objs"][si]["pixels"]
            if len(pixels) > npix_largest:
                npix_largest = len(pixels)
            have = set(map(tuple, pixels_merged))
            pixels_merged = pixels_merged + [x for x in pixels if tuple(x) not in have]
        si, ti = to_merge[comp[-1]]                       # the reference's `index` (last member), not index_largest
        last = tile_sources[ti]["objs"][si]
        pix_min = np.min(pixels_merged, axis=0)
        pix_max = np.max(pixels_merged, axis=0)
        final.append({"name": name, "x1": pix_min[1], "x2": pix_max[1], "y1": pix_min[0], "y2": pix_max[0], "edge": True,
                      "merged": True, "score": last["score"], "class_name": last["class_name"], "class_id": last["class_id"],
                      "pixels": pixels_merged, "vertexes": []})
    for i, src in enumerate(final):
        src["name"] = "S" + str(i + 1)
    return final
