"""Search for the two upstream baselines nobody reproduces (SevenLabels_C / SevenLabels_T) inside a family of
SYSTEMATIC variants of the algorithm -- the same alternative reading applied to every call of every label -- as an
older revision of the reference could have had (tools/bisect_sevenlabels.py covers the other family: today's
algorithm with single calls deviating).  Every variant is run through the literal restatement
(tests/literal_mci.py) on SevenLabels, written through ITK's NRRD byte layout and hashed; SevenLabels_B must stay
reproduced by a variant for it to count as a candidate (printed separately).

Usage: python tools/variant_search.py C|T [nproc [knob=value]]
"""
import itertools
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import literal_mci as L  # noqa: E402
import bisect_sevenlabels as B  # noqa: E402

TARGETS = {m: B.PINS["baselines"][m]["sha512"] for m in "BCT"}

STRUCT = dict(  # knobs shared by both medians
    align=["default", "zero", "centroid", "first_max_exhaustive"],
    carry=["default", "trunc", "floor", "away"],
    rectrans=["itrans", "zero"],
    write=["max", "first", "last"],
    extrap=[True, False],
    cc=["se", "swapped"],
    clip=["default", "keep"],
)
DIL = dict(
    begin=[False, True],
    rev=["i", "none", "j", "both"],
    swap=["lt", "never"],
    ratio=["float", "m1", "round", "min"],
    tie=["first", "last"],
    pick=["score", "mid", "mid_lo"],
)
DT = dict(
    op=["le", "lt"],
    squared=[False, True],
    frac=[10, 1, 2, 5, 100],
    rnd=["trunc", "round"],
    alg=["default", "raya", "raya_lt"],
    tie=["first", "last"],
    ball=[True, False],
)


class Variant(L.Literal):
    def __init__(self, vol, k, **kw):
        super().__init__(vol, **kw)
        self.k = k
        self.extrap = k.get("extrap", True)

    def bounding_box(self, image):
        if not image.a.any():  # a variant translation moved the shape out of the union region
            return image.region
        return super().bounding_box(image)

    def regioned_cc(self, axis, label, s):
        if self.k.get("cc", "se") == "se":
            return super().regioned_cc(axis, label, s)
        ball = self.ball
        self.ball = not ball
        try:
            return super().regioned_cc(axis, label, s)
        finally:
            self.ball = ball

    def align(self, iconn, iid, jconn, jids):
        mode = self.k.get("align", "default")
        if mode == "default":
            return super().align(iconn, iid, jconn, jids)
        if mode == "zero":
            return (0, 0)
        if mode == "centroid":
            ic = self.centroid(iconn, [iid])
            jc = self.centroid(jconn, jids)
            return (jc[0] - ic[0], jc[1] - ic[1])
        h = self.heuristic
        self.heuristic = False
        try:
            return super().align(iconn, iid, jconn, jids)
        finally:
            self.heuristic = h

    def dilation_sequence(self, begin, end):
        seq = super().dilation_sequence(begin, end)
        if self.k.get("begin", False):
            seq.insert(0, begin & end)
        return seq

    def median_dilations(self, inter, imask, jmask):
        k = self.k
        iseq = self.dilation_sequence(inter, imask)
        jseq = self.dilation_sequence(inter, jmask)
        rev = k.get("rev", "i")
        if rev in ("i", "both"):
            iseq.reverse()
        if rev in ("j", "both"):
            jseq.reverse()
        if k.get("swap", "lt") == "lt" and len(iseq) < len(jseq):
            iseq, jseq = jseq, iseq
        ni, nj = len(iseq), len(jseq)
        ratio = np.float32(nj) / np.float32(ni)
        rm = k.get("ratio", "float")
        seq = []
        for x in range(ni):
            if rm == "float":
                xj = int(np.float32(ratio * np.float32(x)))
            elif rm == "m1":
                xj = (x * (nj - 1)) // (ni - 1) if ni > 1 else 0
            elif rm == "round":
                xj = int(float(ratio) * x + 0.5)
            else:
                xj = x
            xj = min(xj, nj - 1)
            seq.append(iseq[x] | jseq[xj])
        pick = k.get("pick", "score")
        if pick == "mid":
            return seq[len(seq) // 2]
        if pick == "mid_lo":
            return seq[(len(seq) - 1) // 2]
        last = k.get("tie", "first") == "last"
        best, mn = 0, imask.size
        for x in range(len(seq)):
            s = abs(int(np.count_nonzero(seq[x] != imask)) - int(np.count_nonzero(seq[x] != jmask)))
            if s < mn or (last and s == mn):
                mn, best = s, x
        return seq[best]

    def median_distances(self, inter, imask, jmask):
        k = self.k
        alg = k.get("alg", "default")
        if alg != "default":  # shape-based interpolation: mean of the two signed distance maps
            si, sj = L.signed_maurer(imask), L.signed_maurer(jmask)
            s = si + sj
            return (s < 0) if alg == "raya_lt" else (s <= 0)
        L.KNOBS["squared"] = k.get("squared", False)
        sdf = L.signed_maurer(inter)
        L.KNOBS.pop("squared", None)
        frac = k.get("frac", 10)
        xor = imask != jmask
        orimg = imask | jmask
        if not xor.any():
            return inter
        v = np.float32(frac) * sdf[xor]
        with np.errstate(all="ignore"):
            d = (np.floor(v + np.float32(0.5)) if k.get("rnd", "trunc") == "round" else np.trunc(v))
            d = np.where(np.abs(d) < 2 ** 31, d, 0).astype(np.int64) & 0xFFFF
        d = np.where(d >= 32768, d - 65536, d)
        if (d < 0).any():
            return inter
        n = int(d.max()) + 1
        ih = np.bincount(d[imask[xor]], minlength=n)
        jh = np.bincount(d[jmask[xor]], minlength=n)
        isum, jsum = np.cumsum(ih), np.cumsum(jh)
        it, jt = int(isum[-1]), int(jsum[-1])
        diff = np.abs(np.abs(it - isum + jsum) - np.abs(jt - jsum + isum))
        mn = diff.min()
        idx = np.flatnonzero(diff == mn)
        best = int(idx[-1] if k.get("tie", "first") == "last" else idx[0])
        upper = np.float32(best) / np.float32(frac)
        return ((sdf < upper) if k.get("op", "le") == "lt" else (sdf <= upper)) & orimg

    def interpolate_1to1(self, axis, label, i, j, iconn, iid, jconn, jid, translation, recursive):
        k = self.k
        cm = k.get("carry", "default")
        itr, jtr = [0, 0], [0, 0]
        ireg, jreg = list(iconn.region), list(jconn.region)
        jbottom = [0, 0]
        carry = False
        for d in range(2):
            t = translation[d]
            if cm == "default":
                if not carry:
                    itr[d] = L.cdiv(t, 2)
                    carry = L.cmod(t, 2) != 0
                elif L.cmod(t, 2) == 0:
                    itr[d] = L.cdiv(t, 2)
                else:
                    itr[d] = L.cdiv(t, 2) + (1 if t > 0 else -1)
                    carry = False
            elif cm == "trunc":
                itr[d] = L.cdiv(t, 2)
            elif cm == "floor":
                itr[d] = t // 2
            else:
                itr[d] = L.cdiv(t + (1 if t > 0 else -1), 2) if t % 2 else t // 2
            jtr[d] = itr[d] - t
            ireg[d] += itr[d]
            jreg[d] += jtr[d]
            jbottom[d] = jreg[d] + jreg[2 + d] - 1
        new_region = L.expand_region(tuple(ireg), (jreg[0], jreg[1]))
        new_region = L.expand_region(new_region, jbottom)
        mid = L.cdiv(i + j + (1 if carry else 0), 2)
        iconn_t = self.translate_image(iconn, itr, new_region)
        jconn_t = self.translate_image(jconn, jtr, new_region)
        if not recursive:
            new_region = self.bounding_box(iconn_t)
            jbb = self.bounding_box(jconn_t)
            new_region = L.expand_region(new_region, (jbb[0], jbb[1]))
            new_region = L.expand_region(new_region, (jbb[0] + jbb[2] - 1, jbb[1] + jbb[3] - 1))
        islice = iconn_t.sub(new_region) == iid
        jslice = jconn_t.sub(new_region) == jid
        inter = islice & jslice
        median = self.median_distances(inter, islice, jslice) if self.dt else self.median_dilations(inter, islice, jslice)
        median = L.Img((new_region[0], new_region[1]), median)
        full_region = new_region
        dimsv = self.dims()
        dd = [a for a in range(3) if a != axis]
        slice_region = (0, 0, dimsv[dd[0]], dimsv[dd[1]])
        _, new_region = L.intersection_regions((0, 0), slice_region, new_region)
        mid_conn = L.zeros(new_region, np.int64)
        if new_region[2] > 0 and new_region[3] > 0:
            m = median.sub(new_region)
            mid_conn.a[m] = 1
            osl = self.slice2d(self.out, axis, mid)
            tgt = osl[new_region[1]:new_region[1] + new_region[3], new_region[0]:new_region[0] + new_region[2]]
            w = k.get("write", "max")
            if w == "max":
                tgt[m & (tgt < label)] = label
            elif w == "first":
                tgt[m & (tgt == 0)] = label
            else:
                tgt[m] = label
        if k.get("clip", "default") == "keep":
            mid_conn = L.Img((full_region[0], full_region[1]), median.a.astype(np.int64))
        if abs(i - j) > 2:
            zero = k.get("rectrans", "itrans") == "zero"
            if abs(i - mid) > 1:
                self.interpolate_1to1(axis, label, i, mid, iconn, iid, mid_conn, 1, (0, 0) if zero else tuple(itr), True)
            if abs(j - mid) > 1:
                self.interpolate_1to1(axis, label, j, mid, jconn, jid, mid_conn, 1, (0, 0) if zero else tuple(jtr), True)


def run_one(args):
    mode, k = args
    dt, ball = B.MODES[mode]
    if mode == "T" and "ball" in k:
        ball = k["ball"]
    try:
        f = Variant(B.VOL, k, dt=dt, ball=ball)
        f.run()
        return k, B.file_hash(f.out)
    except Exception as e:  # a variant may index outside an image; that is a miss
        return k, "error: %r" % (e,)


def product(*spaces):
    keys, vals = [], []
    for s in spaces:
        for kk, vv in s.items():
            keys.append(kk)
            vals.append(vv)
    for combo in itertools.product(*vals):
        yield dict(zip(keys, combo))


def main():
    mode = sys.argv[1].upper()
    nproc = int(sys.argv[2]) if len(sys.argv) > 2 else os.cpu_count()
    spaces = (STRUCT, DT) if mode == "T" else (STRUCT, DIL)
    if len(sys.argv) > 3:  # restrict one knob: name=value
        name, val = sys.argv[3].split("=")
        spaces = tuple({k: ([x for x in v if str(x) == val] if k == name else v) for k, v in sp.items()} for sp in spaces)
    # sanity: the default reading must give today's oracle output
    k0, h0 = run_one((mode, {}))
    print("default reading:", h0[:16], "target:", TARGETS[mode][:16], flush=True)
    todo = ((mode, k) for k in product(*spaces))
    n = 0
    hits = []
    errors = 0
    distinct = set()
    with mp.Pool(nproc) as pool:
        for k, h in pool.imap_unordered(run_one, todo, chunksize=64):
            n += 1
            if h.startswith("error"):
                errors += 1
            else:
                distinct.add(h)
            if h == TARGETS[mode]:
                hits.append(k)
                print("HIT", k, flush=True)
            if n % 20000 == 0:
                print(n, "variants,", len(distinct), "distinct outputs,", errors, "errors,", len(hits), "hits", flush=True)
    print("done:", n, "variants,", len(distinct), "distinct outputs,", errors, "errors,", len(hits), "hits")


if __name__ == "__main__":
    main()
