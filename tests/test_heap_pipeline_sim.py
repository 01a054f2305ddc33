"""CPU model of the ProposalLayer's pipelined heap popper (csrc/box_ops.cuh: popper_warp) against libstdc++'s
std::priority_queue (oracle/nms_ref.cpp: oracle_heap_pop_order, the TF-1.13 NMS pop order).

The model is a statement-by-statement transliteration of the device routine for a 32-lane warp in lock step: one
loop iteration = one round, every lane's loads happen before any lane's store of the round, warp votes / min
reductions are taken over the lanes' values.  It proves nothing about the CUDA compiler, but it does check the
algorithm the kernel implements — top-down pops two levels apart, v re-read, freeze on an uncertain stop — on heaps
with every kind of tie pattern, and that the hazard fallback is not what makes it right (hazards are counted)."""
import numpy as np
import pytest

from oracle import _native as oracle_native

INF = float("inf")
POP_TAIL = 32
INT_MAX = 2 ** 31 - 1


def _seq_pop(h, ln):
    """std::pop_heap on the 1-indexed list h[1..ln] (bottom-up __adjust_heap); returns (top id, new length)."""
    top = h[1][1]
    v = h[ln]
    ln -= 1
    if ln == 0:
        return top, ln
    c = 1
    while 2 * c + 1 <= ln:
        l, r = h[2 * c], h[2 * c + 1]
        right = not (r[0] < l[0])
        h[c] = r if right else l
        c = 2 * c + (1 if right else 0)
    if 2 * c == ln:
        h[c] = h[2 * c]
        c = 2 * c
    while c > 1 and h[c >> 1][0] < v[0]:
        h[c] = h[c >> 1]
        c >>= 1
    h[c] = v
    return top, ln


def _ffs(x):
    return (x & -x).bit_length()


def popper_warp_model(sorted_scores, stats=None, ref=None):
    n = len(sorted_scores)
    h = [(0.0, -7)] + [(float(s), q) for q, s in enumerate(sorted_scores)] + [(0.0, -7)] * 4
    order = [None] * n
    npipe = n - POP_TAIL if n > POP_TAIL else 0
    L = [dict(c=1, len=0, vsrc=1, vid=-1, pop=-1, vs=0.0, xlast=INF, active=False, hazard=False) for _ in range(32)]
    jnext, since, rounds, published, act = 0, 2, 0, 0, 0
    fast_rounds = 0
    if npipe > 0:
        order[0] = h[1][1]
        L[0].update(len=n - 1, vsrc=n, pop=0, active=True)
        # ---- fast mode: fixed schedule, no freeze logic, hands over at the first stop that comes too early
        r, handover, jo = 0, False, 0
        mystart = [2 * (8 if lane == 0 else lane) if lane < 8 else INT_MAX for lane in range(32)]
        while True:
            if (r & 15) == 0 and r > 0:
                if any(x["hazard"] for x in L):
                    break
                fin = min(npipe, (r - 12) // 2 + 1 if r >= 12 else 0)
                published = max(published, fin)
                if ref is not None:
                    assert order[:published] == ref[:published]
                if r >= 2 * npipe + 16:
                    break
            for lane in range(32):
                if r == mystart[lane]:
                    x = L[lane]
                    x["hazard"] |= x["active"]
                    x["pop"] = r >> 1
                    if x["pop"] < npipe:
                        x.update(c=1, len=n - x["pop"] - 1, vsrc=n - x["pop"], vid=-1, xlast=INF, active=True)
                    mystart[lane] += 16
            T, sb = [], 0
            for lane in range(32):
                x = L[lane]
                l = 2 * x["c"]
                inn = x["active"] and l <= x["len"]
                kl, kr = (h[l], h[l + 1]) if inn else (h[2], h[3])
                u = h[x["vsrc"]]
                if x["active"] and u[1] != x["vid"]:
                    x["vs"], x["vid"] = u
                    x["hazard"] |= x["xlast"] < x["vs"]
                right = l < x["len"] and not (kr[0] < kl[0])
                xs, xid = kr if right else kl
                stop = (not inn) or xs < x["vs"]
                if x["active"] and stop:
                    sb |= 1 << lane
                T.append((l, right, xs, xid, stop))
            ob = 1 << (jo & 7)
            if sb & ~ob:
                handover = True
                break
            jo += 1 if (sb & ob) else 0
            for lane in range(32):
                l, right, xs, xid, stop = T[lane]
                x = L[lane]
                if x["active"]:
                    ss, sid = (x["vs"], x["vid"]) if stop else (xs, xid)
                    h[x["c"]] = (ss, sid)
                    if x["c"] == 1:
                        order[x["pop"] + 1] = sid
                    if not stop:
                        x["xlast"] = xs
                        x["c"] = l + (1 if right else 0)
                    x["active"] = not stop
            r += 1
        fast_rounds = r
        act = sum(1 << i for i in range(32) if L[i]["active"])
        jnext = min(npipe, (r >> 1) + 1 if handover else npipe)
        since = (r & 1) if handover else 2
    while act != 0 or jnext < npipe:
        rounds += 1
        if (rounds & 15) == 0:
            if any(x["hazard"] for x in L):
                break
            rot = ((act | (act << 8)) >> (jnext & 7)) & 0xFF
            fin = jnext - 8 + _ffs(rot) if rot else jnext
            published = max(published, fin)
            if ref is not None:                      # what the consumers may read now must already be final
                assert order[:published] == ref[:published]
        rot = ((act | (act << 8)) >> (jnext & 7)) & 0xFF
        minpop = jnext - 9 + _ffs(rot)
        T = []
        for lane in range(32):                      # loads + decisions of the round (no stores yet)
            x = L[lane]
            l = 2 * x["c"]
            inn = x["active"] and l <= x["len"]
            kl, kr = (h[l], h[l + 1]) if inn else (h[2], h[3])
            u = h[x["vsrc"]]
            if x["active"] and u[1] != x["vid"]:
                x["vs"], x["vid"] = u
                x["hazard"] |= x["xlast"] < x["vs"]
            right = l < x["len"] and not (kr[0] < kl[0])
            xs, xid = kr if right else kl
            stop = (not inn) or xs < x["vs"]
            T.append((l, right, xs, xid, stop, x["active"] and stop and x["pop"] != minpop))
        ub = sum(1 << i for i in range(32) if T[i][5])
        stall_pop = INT_MAX
        if ub:
            urot = ((ub | (ub << 8)) >> (jnext & 7)) & 0xFF
            stall_pop = jnext - 9 + _ffs(urot)
            if stats is not None:
                stats["stall_rounds"] = stats.get("stall_rounds", 0) + 1
        for lane in range(32):                      # stores
            l, right, xs, xid, stop, _ = T[lane]
            x = L[lane]
            if x["active"] and x["pop"] < stall_pop:
                ss, sid = (x["vs"], x["vid"]) if stop else (xs, xid)
                h[x["c"]] = (ss, sid)
                if x["c"] == 1:
                    order[x["pop"] + 1] = sid
                if not stop:
                    x["xlast"] = xs
                    x["c"] = l + (1 if right else 0)
                x["active"] = not stop
        act = sum(1 << i for i in range(32) if L[i]["active"])
        if ub == 0:
            since += 1
        sl = jnext & 7
        if jnext < npipe and since >= 2 and not ((act >> sl) & 1):
            L[sl].update(c=1, len=n - jnext - 1, vsrc=n - jnext, vid=-1, xlast=INF, pop=jnext, active=True)
            act |= 1 << sl
            jnext += 1
            since = 0
    hazard = any(x["hazard"] for x in L)
    k0, hl = npipe, n - npipe
    if hazard:
        h = [(0.0, -7)] + [(float(s), q) for q, s in enumerate(sorted_scores)] + [(0.0, -7)] * 4
        k0, hl = 0, n
    elif npipe > published:
        published = npipe
        if ref is not None:
            assert order[:published] == ref[:published]
    for k in range(k0, n):
        idx, hl = _seq_pop(h, hl)
        if k >= published:
            order[k] = idx
    if stats is not None:
        stats.update(rounds=rounds + fast_rounds, npipe=npipe, hazard=hazard, fast_rounds=fast_rounds)
    return order


def _scores(rng, n, kind):
    if kind == "distinct":
        s = rng.random(n)
    elif kind == "few_values":
        s = rng.integers(0, 5, n)
    elif kind == "many_pairs":
        s = rng.integers(0, max(2, n // 3), n)
    elif kind == "all_equal":
        s = np.ones(n)
    elif kind == "quantised":
        s = np.round(rng.random(n), 2)
    elif kind == "sparse_pairs":                    # what real score vectors look like: a handful of equal pairs
        s = rng.random(n)
        if n > 8:
            idx = rng.integers(0, n - 1, 4)
            s[idx] = s[idx + 1]
    elif kind == "saturated_head":                  # trained networks: a plateau of exact 1.0 at the top
        s = rng.random(n)
        s[: max(1, n // 5)] = 1.0
    else:
        raise ValueError(kind)
    return -np.sort(-s.astype(np.float32))


KINDS = ["distinct", "few_values", "many_pairs", "all_equal", "quantised", "sparse_pairs", "saturated_head"]


@pytest.mark.parametrize("kind", KINDS)
def test_pipelined_pops_equal_std_pop_heap_small(kind):
    rng = np.random.default_rng(KINDS.index(kind))
    hazards = 0
    for n in list(range(1, 80)) + [127, 128, 129, 255, 256, 257, 511, 640, 1023, 1024, 1025]:
        s = _scores(rng, n, kind)
        st = {}
        ref = oracle_native.heap_pop_order(s).tolist()
        got = popper_warp_model(s, st, ref)
        assert got == ref, (kind, n)
        hazards += int(st["hazard"])
    assert hazards == 0


@pytest.mark.parametrize("kind", KINDS)
def test_pipelined_pops_equal_std_pop_heap_reference_sizes(kind):
    rng = np.random.default_rng(100 + KINDS.index(kind))
    for n in (6000, 4097, 6144):
        s = _scores(rng, n, kind)
        st = {}
        ref = oracle_native.heap_pop_order(s).tolist()
        got = popper_warp_model(s, st, ref)
        assert got == ref, (kind, n)
        assert not st["hazard"]
        assert st["rounds"] <= 2.2 * st["npipe"] + 32       # freezes are rare: about two rounds per pop


# ---------------------------------------------------------------------------------------------------------------
# closed-form pop order (csrc/proposal.cu step 5): no heap at all for runs of equal scores that end at or below rank K/2
# ---------------------------------------------------------------------------------------------------------------

def pop_key(i):
    """same integer as the device function pop_key: pre-order 'node, right, left' rank of heap index i"""
    d = i.bit_length() - 1
    path = i - (1 << d)
    inv = (~path) & ((1 << d) - 1)
    return ((inv << (16 - d)) << 5) | d


def closed_form_order(sorted_scores):
    """-> (order, covered): order[k] for k < covered is the k-th pop; from `covered` on the heap emulation decides"""
    n = len(sorted_scores)
    order = list(range(n))
    covered, q = n, 0
    while q < n:
        e = q
        while e + 1 < n and sorted_scores[e + 1] == sorted_scores[q]:
            e += 1
        if e > q:
            if e + 1 > n // 2:
                covered = min(covered, q)
            else:
                order[q:e + 1] = sorted(range(q, e + 1), key=lambda r: pop_key(r + 1))
        q = e + 1
    return order, covered


@pytest.mark.parametrize("kind", KINDS)
def test_closed_form_pop_order_equals_std_priority_queue(kind):
    rng = np.random.default_rng(200 + KINDS.index(kind))
    checked = 0
    for n in list(range(1, 70)) + [127, 128, 129, 255, 256, 257, 1000, 4097, 6000, 6144]:
        for _ in range(3 if n < 1000 else 1):
            s = _scores(rng, n, kind)
            ref = oracle_native.heap_pop_order(s).tolist()
            order, covered = closed_form_order(s.tolist())
            assert order[:covered] == ref[:covered], (kind, n, covered)
            checked += covered
    assert checked > 0
    if kind in ("sparse_pairs", "many_pairs", "quantised"):         # runs that really get reordered were part of it
        s = _scores(rng, 6000, kind)
        order, covered = closed_form_order(s.tolist())
        assert covered > 2500 and order[:covered] != list(range(covered))
