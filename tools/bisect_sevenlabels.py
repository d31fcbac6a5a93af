"""Search for the two upstream baselines nobody reproduces (SevenLabels_C / SevenLabels_T) inside the family
"every median call may have picked ANY element of its sequence / ANY populated distance bin".

For each label the literal restatement (tests/literal_mci.py) is replayed over every combination of choices its
median calls offer (children's options depend on the parent's choice: dynamic-radix odometer); the distinct label
coverages are combined over the seven labels with the max rule, written through ITK's NRRD byte layout and hashed.
A hit would tell which call deviates; exhausting the family without a hit says the baselines are not a
median-selection variant of the current algorithm.
Usage: python tools/bisect_sevenlabels.py C|T [nproc [kmax]]   (kmax: at most that many labels deviate; without it
the full product over the labels is searched, 2.4e8 combinations for C)
"""
import hashlib
import itertools
import json
import multiprocessing as mp
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import literal_mci as L  # noqa: E402
from itkcontourinterpolation_b200.nrrd import nrrd_bytes  # noqa: E402

PINS = json.load(open(os.path.join(ROOT, "tests", "golden", "upstream_baselines.json")))["SevenLabels"]
VOL = np.load(os.path.join(ROOT, "tests", "golden", "SevenLabels.npy")).astype(np.int16)
MODES = {"C": (False, False), "T": (True, True), "B": (False, True)}


def coverages(label, dt, ball, limit=200000):
    """distinct coverage masks of `label` over all choice vectors"""
    seen = {}
    prefix = []
    runs = 0
    while True:
        counts = []
        pos = [0]

        def chooser(kind, options, default):
            n = pos[0]
            pos[0] += 1
            counts.append(len(options))
            if n < len(prefix):
                return options[prefix[n]]
            prefix.append(0)
            return options[0]

        L.CHOOSER = chooser
        f = L.Literal(VOL, label=label, dt=dt, ball=ball)
        f.run()
        L.CHOOSER = None
        runs += 1
        cov = (f.out == label) & (VOL == 0)
        seen.setdefault(cov.tobytes(), cov)
        # next vector
        del prefix[len(counts):]
        k = len(prefix) - 1
        while k >= 0 and prefix[k] + 1 >= counts[k]:
            k -= 1
        if k < 0 or runs >= limit:
            break
        prefix[k] += 1
        del prefix[k + 1:]
    return list(seen.values()), runs


_REF = nrrd_bytes(VOL, PINS["space_directions"], PINS["space_origin"], **PINS["writer"])
_HEAD = _REF[:_REF.find(b"\n\n") + 2 + 10]  # NRRD header + teem's 10-byte gzip header
_W = PINS["writer"]


def file_hash(vol):
    raw = vol.tobytes()
    c = zlib.compressobj(_W.get("zlib_level", -1), zlib.DEFLATED, -15, 8, 0)
    body = c.compress(raw) + c.flush()
    return hashlib.sha512(_HEAD + body + zlib.crc32(raw).to_bytes(4, "little") + len(raw).to_bytes(4, "little")).hexdigest()


assert file_hash(VOL) == hashlib.sha512(_REF).hexdigest()


def worker(args):
    target, base, labels, covs, first = args
    hits = []
    n = 0
    rest = [covs[l] for l in labels[1:]]
    b0 = np.where(first, np.int16(labels[0]), base)
    for combo in itertools.product(*rest):
        v = b0
        for l, c in zip(labels[1:], combo):
            v = np.where(c & (v < l), np.int16(l), v)
        v = np.where(VOL != 0, VOL, v)
        n += 1
        if file_hash(v) == target:
            hits.append(v.copy())
    return n, hits


def worker_k(args):
    """labels in `dev` take every alternative coverage, all others their default (index 0 = the restatement's own)"""
    target, covs, dev = args
    order = sorted(covs)
    n = 0
    for combo in itertools.product(*[range(1, len(covs[l])) for l in dev]):
        pick = {l: 0 for l in order}
        pick.update(dict(zip(dev, combo)))
        v = np.zeros_like(VOL)
        for l in order:
            c = covs[l][pick[l]]
            v = np.where(c & (v < l), np.int16(l), v)
        v = np.where(VOL != 0, VOL, v)
        n += 1
        if file_hash(v) == target:
            return n, [v.copy()], (dev, combo)
    return n, [], None


def main_k(mode, kmax, nproc):
    """at most kmax labels deviate from the restatement's own choices (the full product is 10^8 and more)"""
    dt, ball = MODES[mode]
    target = PINS["baselines"][mode]["sha512"]
    covs = {}
    for label in range(1, 8):
        cv, runs = coverages(label, dt, ball)
        L.CHOOSER = None
        f = L.Literal(VOL, label=label, dt=dt, ball=ball)
        f.run()
        own = ((f.out == label) & (VOL == 0)).tobytes()
        cv.sort(key=lambda c: c.tobytes() != own)  # the restatement's own coverage first
        assert cv[0].tobytes() == own
        covs[label] = cv
        print("label", label, "distinct coverages", len(cv), flush=True)
    jobs = []
    for k in range(0, kmax + 1):
        for dev in itertools.combinations(sorted(covs), k):
            if all(len(covs[l]) > 1 for l in dev):
                jobs.append((target, covs, dev))
    done = 0
    with mp.Pool(nproc) as pool:
        for n, hits, what in pool.imap_unordered(worker_k, jobs):
            done += n
            if hits:
                print("HIT", what, flush=True)
                np.save(os.path.join(ROOT, "gpurun_out", "sevenlabels_%s_hit.npy" % mode), hits[0])
                pool.terminate()
                return 0
    print("no hit in", done, "combinations with at most", kmax, "deviating labels")
    return 1


def main():
    mode = sys.argv[1]
    nproc = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    if len(sys.argv) > 3:
        return main_k(mode, int(sys.argv[3]), nproc)
    dt, ball = MODES[mode]
    target = PINS["baselines"][mode]["sha512"]
    covs = {}
    total = 1
    for label in range(1, 8):
        covs[label], runs = coverages(label, dt, ball)
        total *= len(covs[label])
        print("label", label, "choice vectors", runs, "distinct coverages", len(covs[label]), flush=True)
    print("combinations:", total, flush=True)
    labels = sorted(covs, key=lambda l: -len(covs[l]))
    base = np.zeros_like(VOL)
    jobs = [(target, base, labels, covs, c) for c in covs[labels[0]]]
    done = 0
    with mp.Pool(nproc) as pool:
        for k, (n, hits) in enumerate(pool.imap_unordered(worker, jobs)):
            done += n
            if k % 16 == 0:
                print("  ...", done, flush=True)
            if hits:
                print("HIT", flush=True)
                np.save(os.path.join(ROOT, "gpurun_out", "sevenlabels_%s_hit.npy" % mode), hits[0])
                pool.terminate()
                return 0
    print("no hit in", done, "combinations")
    return 1


if __name__ == "__main__":
    sys.exit(main())
