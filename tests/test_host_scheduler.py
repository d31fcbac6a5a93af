"""The host scheduler's pair bookkeeping (InterpolateBetweenTwo, X:1274-1446) without a GPU: the C++ in
libmci_b200.so against an independent Python restatement on random pair sets, and against hand-derived cases
(including the `++rest` quirk of X:1387-1388)."""
import ctypes
import random

import pytest

from itkcontourinterpolation_b200 import _lib as L


def schedule_cpp(pairs, use_extrapolation=True):
    lib = L.lib()
    arr = (L.Pair * max(len(pairs), 1))(*[L.Pair(0, a, b) for a, b in pairs])
    jobs = (L.Job * 4096)()
    ids = (ctypes.c_int32 * 65536)()
    nj, ni = ctypes.c_int32(0), ctypes.c_int32(0)
    rc = lib.mci_filter_schedule_pairs(arr, len(pairs), int(use_extrapolation), jobs, 4096, ctypes.byref(nj), ids, 65536,
                                       ctypes.byref(ni))
    assert rc == 0
    out = []
    for k in range(nj.value):
        j = jobs[k]
        out.append((j.kind, j.slice_a, j.a_id, tuple(ids[j.b_ids_offset:j.b_ids_offset + j.n_b]), j.pos_a, j.pos_b))
    return out


def schedule_py(pairs, use_extrapolation=True):
    """Literal restatement of X:1274-1446 on Python sets/lists (sorted list = std::set iteration order)."""
    unclean = set(pairs)
    unwanted = set()
    for a, b in unclean:
        if a and b:
            unwanted.add((0, b))
            unwanted.add((a, 0))
    ps = sorted(unclean - unwanted)
    jobs = []
    rest = []
    for a, b in ps:                                   # extrapolation first (X:1280-1304)
        if b == 0:
            if use_extrapolation:
                jobs.append((L.JOB_EXTRAPOLATE, 0, a, (), 0, 1))
        elif a == 0:
            if use_extrapolation:
                jobs.append((L.JOB_EXTRAPOLATE, 1, b, (), 1, 0))
        else:
            rest.append((a, b))
    ps = rest
    ic, jc = {}, {}
    for a, b in ps:
        ic[a] = ic.get(a, 0) + 1
        jc[b] = jc.get(b, 0) + 1
    keep = []
    for a, b in ps:                                   # 1-to-1 (X:1315-1333)
        if ic.get(a, 0) == 1 and jc.get(b, 0) == 1:
            jobs.append((L.JOB_ONE_TO_ONE, 0, a, (b,), 0, 1))
            ic.pop(a)
            jc.pop(b)
        else:
            keep.append((a, b))
    ps = keep
    k = 0                                             # M-to-1 / 1-to-N with in-place erasure (X:1335-1411)
    while k < len(ps):
        a, b = ps[k]
        if ic.get(a, 0) == 1:
            ids = tuple(x for x, y in ps if y == b)
            jobs.append((L.JOB_ONE_TO_N, 1, b, ids, 1, 0))
            p = ps[k]
            new = []
            for q in ps:
                if q is not p and q[1] == b:
                    ic[q[0]] = ic.get(q[0], 0) - 1
                    jc[q[1]] = jc.get(q[1], 0) - 1
                else:
                    new.append(q)
            k = new.index(p)
            ic[a] = ic.get(a, 0) - 1
            jc[b] = jc.get(b, 0) - 1
            new.pop(k)
            ps = new
        elif jc.get(b, 0) == 1:
            ids = tuple(y for x, y in ps if x == a)
            jobs.append((L.JOB_ONE_TO_N, 0, a, ids, 0, 1))
            p = ps[k]
            new = [ps[0]] if ps else []               # the erase scan starts at the SECOND element (X:1387-1388)
            for q in ps[1:]:
                if q is not p and q[0] == a:
                    ic[q[0]] = ic.get(q[0], 0) - 1
                    jc[q[1]] = jc.get(q[1], 0) - 1
                else:
                    new.append(q)
            k = [i for i, q in enumerate(new) if q is p][0]
            ic[a] = ic.get(a, 0) - 1
            jc[b] = jc.get(b, 0) - 1
            new.pop(k)
            ps = new
        else:
            k += 1
    while ps:                                         # M-to-N (X:1413-1446)
        a = ps[0][0]
        ids = tuple(y for x, y in ps if x == a)
        jobs.append((L.JOB_ONE_TO_N, 0, a, ids, 0, 1))
        ps = [q for q in ps if q[0] != a]
    return jobs


def test_hand_cases():
    E, O1, ON = L.JOB_EXTRAPOLATE, L.JOB_ONE_TO_ONE, L.JOB_ONE_TO_N
    assert schedule_cpp([]) == []
    # one component on each side, overlapping: (1,1) plus background contacts
    assert schedule_cpp([(1, 1), (1, 0), (0, 1)]) == [(O1, 0, 1, (1,), 0, 1)]
    # appearing / disappearing components extrapolate with swapped roles
    assert schedule_cpp([(1, 0), (0, 2)]) == [(E, 1, 2, (), 1, 0), (E, 0, 1, (), 0, 1)]
    assert schedule_cpp([(1, 0), (0, 2)], use_extrapolation=False) == []
    # 1-to-3
    assert schedule_cpp([(1, 1), (1, 2), (1, 3), (1, 0)]) == [(ON, 0, 1, (1, 2, 3), 0, 1)]
    # 3-to-1: roles swap, the single j component is split
    assert schedule_cpp([(1, 1), (2, 1), (3, 1)]) == [(ON, 1, 1, (1, 2, 3), 1, 0)]
    # ManyToMany.nrrd: all nine (i, j) pairs -> three 1-to-3 jobs
    nine = [(a, b) for a in (1, 2, 3) for b in (1, 2, 3)]
    assert schedule_cpp(nine) == [(ON, 0, a, (1, 2, 3), 0, 1) for a in (1, 2, 3)]


def test_random_pair_sets_match_python_restatement():
    rng = random.Random(7)
    for trial in range(400):
        ni, nj = rng.randint(0, 6), rng.randint(0, 6)
        pairs = set()
        for a in range(0, ni + 1):
            for b in range(0, nj + 1):
                if (a or b) and rng.random() < 0.3:
                    pairs.add((a, b))
        pairs = sorted(pairs)
        for ext in (True, False):
            assert schedule_cpp(pairs, ext) == schedule_py(pairs, ext), (pairs, ext)


def test_quirk_second_element_scan():
    # X:1387-1388: after a 1-to-N job the erase scan starts at the second element, so a first element sharing the
    # same i component survives and is scheduled again in the M-to-N pass
    pairs = [(1, 1), (1, 2), (2, 2), (2, 3)]
    got = schedule_cpp(pairs)
    assert got == schedule_py(pairs)
    assert len(got) >= 2


def test_allocation_free_path_equals_the_literal_set_algebra():
    """The scheduler handles pair sets that hold only extrapolations and 1-to-1 correspondences without the std::set /
    std::map algebra; bit 1 of the hook's flag forces the literal path, and the two must emit the same jobs in the same
    order on every pair set (simple or not)."""
    rng = random.Random(20260421)
    simple = 0
    for trial in range(3000):
        n_i, n_j = rng.randint(1, 5), rng.randint(1, 5)
        pairs = set()
        for _ in range(rng.randint(1, 8)):
            a, b = rng.randint(0, n_i), rng.randint(0, n_j)
            if a or b:
                pairs.add((a, b))
        pairs = sorted(pairs)
        rng.shuffle(pairs)
        for extrap in (0, 1):
            fast = schedule_cpp(pairs, extrap)
            literal = schedule_cpp(pairs, extrap | 2)
            assert fast == literal == schedule_py(pairs, bool(extrap)), (pairs, extrap)
        if all(k != L.JOB_ONE_TO_N for k, *_ in fast):
            simple += 1
    assert simple > 300  # the family the fast path exists for is well represented
