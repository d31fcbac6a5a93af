"""A second, independent restatement of itk::MorphologicalContourInterpolator -- TEST INFRASTRUCTURE ONLY.

Purpose ("next round" item 1a): the C++ oracle (oracle/mci_oracle.cpp) reformulates several steps
(arrival maps instead of image sequences, integer d^2 instead of a float distance image).  This module does NOT: it
follows /root/reference/include/itkMorphologicalContourInterpolator.hxx ("X") statement by statement on whole
images, builds every dilation-sequence image (X:322-388), a float32 signed distance IMAGE and the float threshold
(X:390-516), and takes its primitives from scipy (ndimage.label, binary_dilation, distance_transform_edt) rather
than from the oracle's own code.  Pure Python loops: small cases only.  tests/test_literal_restatement.py compares
the two restatements call type by call type; agreement of two independently written readings of X is the
substitute for the two upstream baselines (SevenLabels_C / _T) that neither reproduces.

`KNOBS` holds switches for alternative readings tried while bisecting those two baselines
(tools/bisect_sevenlabels.py); every default is the reading of X as it stands.
"""
from collections import deque

import numpy as np
from scipy import ndimage

KNOBS = {}


def knob(name, default):
    return KNOBS.get(name, default)


CHOOSER = None  # hook used by tools/bisect_sevenlabels.py: CHOOSER(kind, options, default) -> option

TRACE = None  # list collecting (kind, payload) records of the calls made, when not None


def trace(kind, **kw):
    if TRACE is not None:
        TRACE.append((kind, kw))


class Img:
    """itk::Image<P,2> stand-in: region index (d0, d1), array indexed [d1, d0] (d0 fastest)."""

    def __init__(self, idx, arr):
        self.idx = (int(idx[0]), int(idx[1]))
        self.a = arr

    @property
    def size(self):
        return (self.a.shape[1], self.a.shape[0])

    @property
    def region(self):
        return (self.idx[0], self.idx[1], self.a.shape[1], self.a.shape[0])

    def get(self, p):
        return self.a[p[1] - self.idx[1], p[0] - self.idx[0]]

    def sub(self, reg):
        """view over region reg = (i0, i1, s0, s1), which must lie inside"""
        x0, y0 = reg[0] - self.idx[0], reg[1] - self.idx[1]
        assert x0 >= 0 and y0 >= 0 and x0 + reg[2] <= self.a.shape[1] and y0 + reg[3] <= self.a.shape[0], (self.region, reg)
        return self.a[y0:y0 + reg[3], x0:x0 + reg[2]]


def zeros(reg, dtype):
    return Img((reg[0], reg[1]), np.zeros((max(reg[3], 0), max(reg[2], 0)), dtype))


def expand_region(reg, p):  # X:107-126
    idx, sz = [reg[0], reg[1]], [reg[2], reg[3]]
    for a in range(2):
        if idx[a] > p[a]:
            sz[a] = sz[a] + idx[a] - p[a]
            idx[a] = p[a]
        elif idx[a] + sz[a] <= p[a]:
            sz[a] = p[a] - idx[a] + 1
    return (idx[0], idx[1], sz[0], sz[1])


def intersection_regions(t, ri, rj):  # X:1011-1030
    iidx, isz, jidx = [ri[0], ri[1]], [ri[2], ri[3]], [rj[0], rj[1]]
    jsz = [rj[2], rj[3]]
    for d in range(2):
        ib = iidx[d] + t[d]
        tt = max(ib, jidx[d])
        isz[d] = min(isz[d] - (tt - ib), jsz[d] - (tt - jidx[d]))
        iidx[d] = tt - t[d]
        jidx[d] = tt
    return (iidx[0], iidx[1], isz[0], isz[1]), (jidx[0], jidx[1], isz[0], isz[1])


def cdiv(a, b):  # C++ truncating division
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


def cmod(a, b):
    return a - b * cdiv(a, b)


# ------------------------------------------------------------------ ITK primitives from scipy
def connected_components(mask, fully):
    st = np.ones((3, 3), int) if fully else ndimage.generate_binary_structure(2, 1)
    lab, n = ndimage.label(mask, structure=st)
    return lab, n


def dilate_r1(seed, ball):
    st = np.ones((3, 3), bool) if ball else ndimage.generate_binary_structure(2, 1)
    return ndimage.binary_dilation(seed, structure=st, border_value=0)


def signed_maurer(mask):
    """SignedMaurerDistanceMapImageFilter<bool,float>, UseImageSpacing off, SquaredDistance off, inside negative:
    distance to the object's border pixels (BinaryContour, fully connected), float32."""
    if not mask.any():
        return np.full(mask.shape, np.finfo(np.float32).max, np.float32)
    pad = np.pad(mask, 1, constant_values=True)  # outside the image does not make a border pixel
    er = ndimage.binary_erosion(pad, structure=np.ones((3, 3), bool))[1:-1, 1:-1]
    contour = mask & ~er
    if not contour.any():
        return np.full(mask.shape, -np.finfo(np.float32).max, np.float32)
    # exact integer squared distance to the nearest contour pixel
    ys, xs = np.nonzero(contour)
    yy, xx = np.mgrid[0:mask.shape[0], 0:mask.shape[1]]
    d2 = ((yy[..., None] - ys) ** 2 + (xx[..., None] - xs) ** 2).min(axis=-1)
    if knob("squared", False):
        d = d2.astype(np.float32)
    else:
        d = np.sqrt(d2.astype(np.float32)).astype(np.float32)
    return np.where(mask, -d, d).astype(np.float32)


# ------------------------------------------------------------------ the filter
class Literal:
    def __init__(self, vol, label=0, axis=-1, heuristic=True, dt=True, ball=False, extrapolation=True):
        self.inp = vol
        self.out = np.zeros_like(vol)
        self.ptype = vol.dtype
        self.label, self.axis, self.heuristic, self.dt, self.ball, self.extrap = label, axis, heuristic, dt, ball, extrapolation
        self.min_iters = int(2 ** 3)
        self.max_iters = int(6 ** 3)

    # volume index p = (x, y, z); numpy is [z, y, x]
    def vget(self, vol, p):
        return vol[p[2], p[1], p[0]]

    def dims(self):
        return (self.inp.shape[2], self.inp.shape[1], self.inp.shape[0])

    def determine_slice_orientations(self):  # X:128-201
        self.slices = [dict(), dict(), dict()]
        self.bboxes = {}
        X, Y, Z = self.dims()
        nz = np.argwhere(self.inp != 0)
        for z, y, x in nz:
            ind = (int(x), int(y), int(z))
            val = int(self.inp[z, y, x])
            bb = self.bboxes.get(val)
            if bb is None:
                self.bboxes[val] = [list(ind), list(ind)]
            else:
                for a in range(3):
                    bb[0][a] = min(bb[0][a], ind[a])
                    bb[1][a] = max(bb[1][a], ind[a])
            ctrue = cadj = 0
            ax = 0
            for a in range(3):
                prev = nxt = 0
                if ind[a] - 1 >= 0:
                    q = list(ind); q[a] -= 1
                    prev = int(self.vget(self.inp, q))
                if ind[a] + 1 < (X, Y, Z)[a]:
                    q = list(ind); q[a] += 1
                    nxt = int(self.vget(self.inp, q))
                if prev == 0 and nxt == 0:
                    ax = a
                    ctrue += 1
                elif prev == val and nxt == val:
                    cadj += 1
            if ctrue == 1 and cadj == 2 and (self.axis == -1 or self.axis == ax):
                self.slices[ax].setdefault(val, set()).add(ind[ax])

    def slice2d(self, vol, axis, s):
        if axis == 0:
            return vol[:, :, s]
        if axis == 1:
            return vol[:, s, :]
        return vol[s, :, :]

    def regioned_cc(self, axis, label, s):  # X:1224-1238 over the label's bounding box
        lo, hi = self.bboxes[label]
        d = [a for a in range(3) if a != axis]
        full = self.slice2d(self.inp, axis, s)
        sub = full[lo[d[1]]:hi[d[1]] + 1, lo[d[0]]:hi[d[0]] + 1] == label
        lab, n = connected_components(sub, self.ball)
        return Img((lo[d[0]], lo[d[1]]), lab.astype(np.int64)), n

    # ---------------------------------------------------------------- X:1101-1134
    def centroid(self, conn, ids):
        m = np.isin(conn.a, list(ids)) & (conn.a != 0)
        ys, xs = np.nonzero(m)
        cnt = len(xs)
        sx = int(xs.sum()) + cnt * conn.idx[0]
        sy = int(ys.sum()) + cnt * conn.idx[1]
        return (cdiv(sx, cnt), cdiv(sy, cnt))

    # ---------------------------------------------------------------- X:1032-1077
    def intersection(self, iconn, iid, jconn, jids, t):
        ri, rj = intersection_regions(t, iconn.region, jconn.region)
        ia, ja = iconn.sub(ri), jconn.sub(rj)
        sel = ia == iid
        total = 0
        for j in jids:
            c = int(np.count_nonzero(sel & (ja == j)))
            if c == 0:
                return 0
            total += c
        return total

    # ---------------------------------------------------------------- X:1136-1222
    def align(self, iconn, iid, jconn, jids):
        ic = self.centroid(iconn, [iid])
        jc = self.centroid(jconn, jids)
        ind = (jc[0] - ic[0], jc[1] - ic[1])
        ir, jr = iconn.region, jconn.region
        sidx = [jr[d] - ir[d] - ir[2 + d] + 1 for d in range(2)]
        ssz = [ir[2 + d] + jr[2 + d] - 1 for d in range(2)]
        npx = ssz[0] * ssz[1]

        def inside(p):
            return all(sidx[d] <= p[d] < sidx[d] + ssz[d] for d in range(2))

        searched = set()
        q = deque()
        t0 = (0, 0)
        q.append(t0)
        q.append(ind)
        searched.add(t0)
        searched.add(ind)
        max_score = 0
        best = None
        it = 0
        min_iter = min(self.min_iters, npx)
        max_iter = max(self.max_iters, int(np.sqrt(float(npx))))
        evals = 0
        while q:
            ind = q.popleft()
            score = self.intersection(iconn, iid, jconn, jids, ind) if inside(ind) else 0
            evals += 1
            if score > max_score:
                max_score = score
                best = ind
            if (not self.heuristic) or max_score == 0 or it <= min_iter or (score > max_score * 0.9 and it <= max_iter):
                p = list(ind)
                for d in range(2):
                    p[d] -= 1
                    if inside(p) and tuple(p) not in searched:
                        q.append(tuple(p)); searched.add(tuple(p)); it += 1
                    p[d] += 2
                    if inside(p) and tuple(p) not in searched:
                        q.append(tuple(p)); searched.add(tuple(p)); it += 1
                    p[d] -= 1
        trace("align", iid=iid, jids=list(jids), t=best, evals=evals)
        return best

    # ---------------------------------------------------------------- X:994-1009
    def translate_image(self, image, t, new_region):
        res = zeros(new_region, image.a.dtype)
        inr, outr = intersection_regions(t, image.region, new_region)
        if inr[2] > 0 and inr[3] > 0:
            res.sub(outr)[...] = image.sub(inr)
        return res

    # ---------------------------------------------------------------- X:518-554
    def bounding_box(self, image):
        ys, xs = np.nonzero(image.a)
        return (image.idx[0] + int(xs.min()), image.idx[1] + int(ys.min()), int(xs.max() - xs.min()) + 1, int(ys.max() - ys.min()) + 1)

    # ---------------------------------------------------------------- X:322-338
    def dilation_sequence(self, begin, end):
        seq = [dilate_r1(begin, self.ball) & end]
        while True:
            seq.append(dilate_r1(seq[-1], self.ball) & end)
            if np.array_equal(seq[-1], seq[-2]):
                break
        seq.pop()
        return seq

    # ---------------------------------------------------------------- X:340-388
    def median_dilations(self, inter, imask, jmask):
        iseq = self.dilation_sequence(inter, imask)
        jseq = self.dilation_sequence(inter, jmask)
        iseq.reverse()
        if len(iseq) < len(jseq):
            iseq, jseq = jseq, iseq
        ratio = np.float32(len(jseq)) / np.float32(len(iseq))
        seq = []
        for x in range(len(iseq)):
            xj = int(np.float32(ratio * np.float32(x)))
            seq.append(iseq[x] | jseq[xj])
        min_index = 0
        mn = imask.size
        for x in range(len(iseq)):
            i_s = int(np.count_nonzero(seq[x] != imask))
            j_s = int(np.count_nonzero(seq[x] != jmask))
            score = abs(i_s - j_s)
            if score < mn:
                mn = score
                min_index = x
        trace("median_dil", ki=len(iseq), kj=len(jseq), pick=min_index)
        if CHOOSER is not None:
            min_index = CHOOSER("dil", list(range(len(seq))), min_index)
        return seq[min_index]

    # ---------------------------------------------------------------- X:405-516
    def to_pixel(self, f):
        """PixelType dist = short(10) * float: float multiply, truncating conversion, narrowing"""
        v = int(np.float32(10) * np.float32(f))  # trunc toward zero
        bits = self.ptype.itemsize * 8
        v &= (1 << bits) - 1
        if self.ptype.kind == "i" and v >= 1 << (bits - 1):
            v -= 1 << bits
        return v

    def median_distances(self, inter, imask, jmask):
        sdf = signed_maurer(inter)
        ih, jh = [], []
        orimg = np.zeros(imask.shape, bool)
        H, W = imask.shape
        for y in range(H):
            for x in range(W):
                im, jm = bool(imask[y, x]), bool(jmask[y, x])
                if im == jm:
                    if im:
                        orimg[y, x] = True
                    continue
                dist = self.to_pixel(sdf[y, x])
                assert dist >= 0
                h = ih if im else jh
                if dist >= len(h):
                    h.extend([0] * (dist + 1 - len(h)))
                h[dist] += 1
                orimg[y, x] = True
        max_size = max(len(ih), len(jh))
        if max_size == 0:
            trace("median_dt", best=None)
            return inter
        ih += [0] * (max_size - len(ih))
        jh += [0] * (max_size - len(jh))
        isum, jsum = np.cumsum(ih), np.cumsum(jh)
        itot, jtot = int(isum[-1]), int(jsum[-1])
        best_bin, best_diff = 0, None
        for b in range(max_size):
            i_s = abs(itot - int(isum[b]) + int(jsum[b]))
            j_s = abs(jtot - int(jsum[b]) + int(isum[b]))
            diff = abs(i_s - j_s)
            if best_diff is None or diff < best_diff:
                best_diff, best_bin = diff, b
        if CHOOSER is not None:
            best_bin = CHOOSER("dt", sorted(set([0] + [b for b in range(max_size) if ih[b] or jh[b]])), best_bin)
        upper = np.float32(best_bin) / np.float32(10)
        trace("median_dt", best=best_bin)
        if knob("strict_threshold", False):
            return (sdf < upper) & orimg
        return (sdf <= upper) & orimg

    # ---------------------------------------------------------------- X:556-793
    def interpolate_1to1(self, axis, label, i, j, iconn, iid, jconn, jid, translation, recursive):
        itr, jtr = [0, 0], [0, 0]
        ireg, jreg = list(iconn.region), list(jconn.region)
        jbottom = [0, 0]
        carry = False
        for d in range(2):
            t = translation[d]
            if not carry:
                itr[d] = cdiv(t, 2)
                carry = cmod(t, 2) != 0
            elif cmod(t, 2) == 0:
                itr[d] = cdiv(t, 2)
            else:
                itr[d] = cdiv(t, 2) + (1 if t > 0 else -1)
                carry = False
            jtr[d] = itr[d] - t
            ireg[d] += itr[d]
            jreg[d] += jtr[d]
            jbottom[d] = jreg[d] + jreg[2 + d] - 1
        new_region = expand_region(tuple(ireg), (jreg[0], jreg[1]))
        new_region = expand_region(new_region, jbottom)
        mid = cdiv(i + j + (1 if carry else 0), 2)
        trace("one2one", label=label, i=i, j=j, mid=mid, iid=iid, jid=jid, t=tuple(translation), recursive=recursive)

        iconn_t = self.translate_image(iconn, itr, new_region)
        jconn_t = self.translate_image(jconn, jtr, new_region)
        if not recursive:
            new_region = self.bounding_box(iconn_t)
            jbb = self.bounding_box(jconn_t)
            new_region = expand_region(new_region, (jbb[0], jbb[1]))
            new_region = expand_region(new_region, (jbb[0] + jbb[2] - 1, jbb[1] + jbb[3] - 1))
        islice = iconn_t.sub(new_region) == iid
        jslice = jconn_t.sub(new_region) == jid
        inter = islice & jslice
        if self.dt:
            median = self.median_distances(inter, islice, jslice)
        else:
            median = self.median_dilations(inter, islice, jslice)
        median = Img((new_region[0], new_region[1]), median)

        # clip to the output slice (X:679-712)
        dimsv = self.dims()
        d = [a for a in range(3) if a != axis]
        slice_region = (0, 0, dimsv[d[0]], dimsv[d[1]])
        _, new_region = intersection_regions((0, 0), slice_region, new_region)
        mid_conn = zeros(new_region, np.int64)
        if new_region[2] > 0 and new_region[3] > 0:
            m = median.sub(new_region)
            mid_conn.a[m] = 1
            # write rule X:754-763
            osl = self.slice2d(self.out, axis, mid)
            tgt = osl[new_region[1]:new_region[1] + new_region[3], new_region[0]:new_region[0] + new_region[2]]
            tgt[m & (tgt < label)] = label
        if abs(i - j) > 2:
            if abs(i - mid) > 1:
                self.interpolate_1to1(axis, label, i, mid, iconn, iid, mid_conn, 1, tuple(itr), True)
            if abs(j - mid) > 1:
                self.interpolate_1to1(axis, label, j, mid, jconn, jid, mid_conn, 1, tuple(jtr), True)

    # ---------------------------------------------------------------- X:824-992
    def interpolate_1toN(self, axis, label, i, j, iconn, iid, jconn, jids, translation):
        mask = iconn.a == iid
        ireg_full = iconn.region
        ri, rj = intersection_regions(translation, ireg_full, jconn.region)
        n = len(jids)
        blobs = [np.zeros(mask.shape, bool) for _ in range(n)]
        belongs = np.zeros(mask.shape, np.int64)
        # seeds (X:871-892)
        ox, oy = ri[0] - iconn.idx[0], ri[1] - iconn.idx[1]
        jsub = jconn.sub(rj)
        for y in range(ri[3]):
            for x in range(ri[2]):
                if mask[oy + y, ox + x]:
                    jv = jsub[y, x]
                    if jv in jids:
                        k = list(jids).index(jv)
                        blobs[k][oy + y, ox + x] = True
                        belongs[oy + y, ox + x] = k + 1
        steps = 0
        while True:
            steps += 1
            assert steps < 10000
            for k in range(n):
                blobs[k] = dilate_r1(blobs[k], self.ball) & mask
            hollowed_empty = True
            H, W = mask.shape
            for y in range(H):
                for x in range(W):
                    if not mask[y, x]:
                        continue
                    if not belongs[y, x]:
                        k = 0
                        while k < n and not blobs[k][y, x]:
                            k += 1
                        if k < n:
                            belongs[y, x] = k + 1
                            for k2 in range(k + 1, n):
                                blobs[k2][y, x] = False
                        else:
                            hollowed_empty = False
                    else:
                        for k in range(n):
                            if belongs[y, x] != k + 1:
                                blobs[k][y, x] = False
            if hollowed_empty:
                break
        trace("one2N", label=label, i=i, j=j, iid=iid, jids=list(jids), t=tuple(translation), steps=steps)
        for k in range(n):
            c = zeros(ireg_full, np.int64)
            c.a[belongs == k + 1] = iid
            self.interpolate_1to1(axis, label, i, j, c, iid, jconn, jids[k], translation, False)

    # ---------------------------------------------------------------- X:203-251
    def extrapolate(self, axis, label, i, j, iconn, iid):
        c = self.centroid(iconn, [iid])
        ph = zeros((c[0] - 1, c[1] - 1, 3, 3), np.int64)
        ph.a[1, 1] = iid
        t = self.align(iconn, iid, ph, [iid])
        ph = Img((c[0] - 1 - t[0], c[1] - 1 - t[1]), ph.a)
        trace("extrapolate", label=label, i=i, j=j, iid=iid, t=t)
        self.interpolate_1to1(axis, label, i, j, iconn, iid, ph, iid, (0, 0), False)

    # ---------------------------------------------------------------- X:1240-1447
    def interpolate_between_two(self, axis, label, i, j, iconn, jconn):
        unclean, unwanted = set(), set()
        ia, ja = iconn.a, jconn.a
        assert ia.shape == ja.shape
        for a, b in zip(ia.ravel().tolist(), ja.ravel().tolist()):
            if a != 0 or b != 0:
                unclean.add((a, b))
                if a != 0 and b != 0:
                    unwanted.add((0, b))
                    unwanted.add((a, 0))
        pairs = sorted(unclean - unwanted)

        def next_after(v):
            for k, q in enumerate(pairs):
                if q > v:
                    return k
            return len(pairs)

        # extrapolation (X:1280-1304)
        k = 0
        while k < len(pairs):
            p = pairs[k]
            if p[1] == 0:
                if self.extrap:
                    self.extrapolate(axis, label, i, j, iconn, p[0])
                pairs.pop(k)
            elif p[0] == 0:
                if self.extrap:
                    self.extrapolate(axis, label, j, i, jconn, p[1])
                pairs.pop(k)
            else:
                k += 1
        icounts, jcounts = {}, {}
        for p in pairs:
            icounts[p[0]] = icounts.get(p[0], 0) + 1
            jcounts[p[1]] = jcounts.get(p[1], 0) + 1
        # 1-to-1 (X:1315-1333)
        k = 0
        while k < len(pairs):
            p = pairs[k]
            if icounts.get(p[0], 0) == 1 and jcounts.get(p[1], 0) == 1:
                t = self.align(iconn, p[0], jconn, [p[1]])
                self.interpolate_1to1(axis, label, i, j, iconn, p[0], jconn, p[1], t, False)
                icounts.pop(p[0], None)
                jcounts.pop(p[1], None)
                pairs.pop(k)
            else:
                k += 1
        # 1-to-N and M-to-1 (X:1335-1411)
        k = 0
        while k < len(pairs):
            p = pairs[k]
            if icounts.get(p[0], 0) == 1:  # M-to-1
                ids = [r[0] for r in pairs if r[1] == p[1]]
                t = self.align(jconn, p[1], iconn, ids)
                self.interpolate_1toN(axis, label, j, i, jconn, p[1], iconn, ids, t)
                for r in list(pairs):
                    if r != p and r[1] == p[1]:
                        icounts[r[0]] = icounts.get(r[0], 0) - 1
                        jcounts[r[1]] = jcounts.get(r[1], 0) - 1
                        pairs.remove(r)
                icounts[p[0]] = icounts.get(p[0], 0) - 1
                jcounts[p[1]] = jcounts.get(p[1], 0) - 1
                pairs.remove(p)
                k = next_after(p)
            elif jcounts.get(p[1], 0) == 1:  # 1-to-N
                ids = [r[1] for r in pairs if r[0] == p[0]]
                t = self.align(iconn, p[0], jconn, ids)
                self.interpolate_1toN(axis, label, i, j, iconn, p[0], jconn, ids, t)
                for r in list(pairs)[1:]:  # the erase loop starts at ++pairs.begin() (X:1387-1388)
                    if r != p and r[0] == p[0]:
                        icounts[r[0]] = icounts.get(r[0], 0) - 1
                        jcounts[r[1]] = jcounts.get(r[1], 0) - 1
                        pairs.remove(r)
                icounts[p[0]] = icounts.get(p[0], 0) - 1
                jcounts[p[1]] = jcounts.get(p[1], 0) - 1
                pairs.remove(p)
                k = next_after(p)
            else:
                k += 1
        # M-to-N (X:1413-1446)
        while pairs:
            p = pairs[0]
            ids = [r[1] for r in pairs if r[0] == p[0]]
            t = self.align(iconn, p[0], jconn, ids)
            self.interpolate_1toN(axis, label, i, j, iconn, p[0], jconn, ids, t)
            pairs[:] = [r for r in pairs if r[0] != p[0]]

    # ---------------------------------------------------------------- X:1449-1539
    def interpolate_along(self, axis, keep=None):
        """keep = (lo, hi): only the segments with a mid slice in [lo, hi) (slab-sharded runs)"""
        for label, sl in self.slices[axis].items():
            if self.label not in (0, label) or not sl or label not in self.bboxes:
                continue
            order = sorted(sl)
            iconn, _ = self.regioned_cc(axis, label, order[0])
            prev = order[0]
            for nxt in order[1:]:
                jconn, _ = self.regioned_cc(axis, label, nxt)
                if prev + 1 < nxt and (keep is None or max(prev + 1, keep[0]) <= min(nxt - 1, keep[1] - 1)):
                    trace("segment", axis=axis, label=label, i=prev, j=nxt)
                    self.interpolate_between_two(axis, label, prev, nxt, iconn, jconn)
                iconn, prev = jconn, nxt

    # ---------------------------------------------------------------- X:1559-1637
    def run(self):
        self.determine_slice_orientations()
        if not self.bboxes:
            return self.inp.copy()
        if self.axis == -1:
            for a in range(3):
                if self.label == 0:
                    agg = any(len(s) > 1 for s in self.slices[a].values())
                else:
                    agg = len(self.slices[a].get(self.label, ())) > 1
                if agg:
                    self.interpolate_along(a)
        else:
            self.interpolate_along(self.axis)
        nz = self.inp != 0
        self.out[nz] = self.inp[nz]
        return self.out


def interpolate(vol, **kw):
    return Literal(np.ascontiguousarray(vol), **kw).run()


def interpolate_with_detection(vol, bboxes, slices, axis, keep, **kw):
    """The filter after slice detection, on tables supplied by the caller (the slab-sharded multi-GPU path:
    include/mci_b200.h, mci_filter_run_with_detection).  bboxes {label: (lo[3], hi[3])}, slices {label: positions}."""
    f = Literal(np.ascontiguousarray(vol), axis=axis, **kw)
    f.bboxes = {int(l): [list(lo), list(hi)] for l, (lo, hi) in bboxes.items()}
    f.slices = [dict(), dict(), dict()]
    f.slices[axis] = {int(l): set(int(z) for z in zs) for l, zs in slices.items()}
    f.interpolate_along(axis, keep)
    nz = f.inp != 0
    f.out[nz] = f.inp[nz]
    return f.out
