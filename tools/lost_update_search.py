"""Search for the two unreproduced upstream baselines (SevenLabels_C / _T) inside the family "a racy write lost an
update": before the reference serialised its volume writes (the mutex of X:740-742), two threads applying the max rule
`if (out < label) out = label` (X:754-763) to the same voxel could leave the LOWER label.  Every voxel that several
labels cover may have kept any lower covering label; all combinations of K such voxels are written through ITK's NRRD
byte layout and hashed.  Usage: python tools/lost_update_search.py C|T K
Result (round 2): no hit for K <= 5 (C, 632 M volumes) and K <= 7 (T, 62 M)."""
import sys, itertools, multiprocessing as mp
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo/tools')
import numpy as np
import bisect_sevenlabels as B
import literal_mci as L
mode = sys.argv[1]; K = int(sys.argv[2])
dt, ball = B.MODES[mode]
full = L.Literal(B.VOL, dt=dt, ball=ball); full.run()
cover = {}
for lab in range(1, 8):
    f = L.Literal(B.VOL, label=lab, dt=dt, ball=ball); f.run()
    cover[lab] = (f.out == lab) & (B.VOL == 0)
stack = np.stack([cover[l] for l in range(1, 8)])
cnt = stack.sum(0)
alts = []
for z, y, x in np.argwhere(cnt > 1):
    labs = [l for l in range(1, 8) if cover[l][z, y, x]]
    for l in labs[:-1]:
        alts.append(((int(z), int(y), int(x)), l))
target = B.PINS["baselines"][mode]["sha512"]
base = full.out.copy()
def work(first):
    n = 0
    for rest in itertools.combinations(range(first + 1, len(alts)), K - 1):
        combo = (first,) + rest
        ps = [alts[c][0] for c in combo]
        if len(set(ps)) < K:
            continue
        v = base.copy()
        for c in combo:
            (z, y, x), l = alts[c]
            v[z, y, x] = l
        n += 1
        if B.file_hash(v) == target:
            return ("HIT", [alts[c] for c in combo])
    return ("ok", n)
if __name__ == "__main__":
    tot = 0
    with mp.Pool(8) as pool:
        for r in pool.imap_unordered(work, range(len(alts))):
            if r[0] == "HIT":
                print("HIT", mode, r[1], flush=True)
            else:
                tot += r[1]
    print(mode, "k =", K, "tried", tot, "no hit" )
