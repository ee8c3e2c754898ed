"""Generate tests/golden/spec_v1.npz from the pure-Python restatement of SPEC.md
(tests/spec_py.py).  Run once, by hand, in the build container:

    python tests/make_golden.py

The reference snapshot holds no golden vectors for this path (it has no tests at
all), so these fixtures pin the spec text itself: inputs are stored
verbatim next to the expected canonical keys, heights and terrain voxels.  The C
oracle (oracle/) and the CUDA path are both checked against them.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import spec_py as S  # noqa: E402

OUT = os.path.join(HERE, "golden", "spec_v1.npz")
DIRS = S.DIRS


def idx(x, y, z):
    return x + 32 * y + 1024 * z


def nbr_from_coords(coords):
    table = {tuple(c): i for i, c in enumerate(coords)}
    nbr = np.full((len(coords), 6), -1, dtype=np.int32)
    for i, c in enumerate(coords):
        for f, d in enumerate(DIRS):
            nbr[i, f] = table.get((c[0] + d[0], c[1] + d[1], c[2] + d[2]), -1)
    return nbr


def build_cases():
    rng = np.random.default_rng(20261019)
    cases = {}

    def add(name, vox, nbr=None):
        vox = np.asarray(vox, dtype=np.uint16).reshape(-1, 32768)
        if nbr is None:
            nbr = np.full((vox.shape[0], 6), -1, dtype=np.int32)
        cases[name] = (vox, np.asarray(nbr, dtype=np.int32))

    add("empty", np.zeros(32768))
    add("solid", np.full(32768, 7))
    v = np.zeros(32768); v[idx(5, 6, 7)] = 9
    add("single", v)
    v = np.zeros(32768); v[idx(3, 4, 5)] = 2; v[idx(4, 4, 5)] = 2
    add("bar2", v)
    v = np.zeros(32768); v[idx(3, 4, 5)] = 2; v[idx(4, 4, 5)] = 3
    add("two_ids", v)
    v = np.zeros(32768); v[idx(0, 0, 0)] = 1; v[idx(31, 31, 31)] = 65535; v[idx(31, 0, 0)] = 0x8000
    add("corners_highid", v)
    # a slab of one id with a different-id stripe: merges must stop at the stripe
    v = np.zeros((32, 32, 32), dtype=np.uint16); v[:, 10, :] = 5; v[:, 10, 13] = 6; v[7, 10, :] = 6
    add("slab_stripes", v.reshape(-1))
    add("checker_alone", S.pattern_chunk("checker", 0, 0, 0))
    # centre chunk with all six neighbours continuing the world-space pattern
    cc = [(1, 1, 1)] + [(1 + d[0], 1 + d[1], 1 + d[2]) for d in DIRS]
    add("checker_nbrs", np.stack([S.pattern_chunk("checker", *c) for c in cc]), nbr_from_coords(cc))
    add("solid4_nbrs", np.stack([S.pattern_chunk("solid4", *c) for c in cc[:4]]), nbr_from_coords(cc[:4]))
    for p in (0.02, 0.5, 0.9):
        solid = rng.random(32768) < p
        ids = rng.integers(1, 5, size=32768)
        add(f"random_p{int(p * 100):02d}", np.where(solid, ids, 0))
    # two random chunks that are each other's +X/-X neighbours plus a self-referencing Y link
    a = np.where(rng.random(32768) < 0.6, rng.integers(1, 3, size=32768), 0)
    b = np.where(rng.random(32768) < 0.4, rng.integers(1, 3, size=32768), 0)
    nbr = np.array([[-1, 1, 0, -1, -1, 1], [0, -1, -1, -1, 0, -1]], dtype=np.int32)
    add("random_pair_asym", np.stack([a, b]), nbr)
    # terrain: a 2x3x2 block of chunks straddling the seed-42 surface near the origin
    tc = [(x, y, z) for z in (0, 1) for y in (7, 8, 9) for x in (0, 1)]
    add("terrain42", np.stack([S.terrain_chunk(42, *c) for c in tc]), nbr_from_coords(tc))
    return cases


def main():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    blob = {}
    # heights
    seeds = np.array([0, 1, 42, 0xC0FFEE, 2**63 + 12345], dtype=np.uint64)
    rng = np.random.default_rng(7)
    xz = np.concatenate([rng.integers(-100000, 100000, size=(200, 2)),
                         np.array([[0, 0], [-1, -1], [127, 128], [-128, 127], [2**20, -2**20]])]).astype(np.int32)
    hs = np.array([[S.height(int(s), int(x), int(z)) for (x, z) in xz] for s in seeds], dtype=np.int32)
    blob["height_seeds"], blob["height_xz"], blob["height_h"] = seeds, xz, hs
    # terrain voxels of a few chunks, verbatim
    tcoords = np.array([[0, 8, 0], [-3, 7, 5], [10, 9, -7], [0, 0, 0], [0, 20, 0]], dtype=np.int32)
    blob["terrain_coords"] = tcoords
    blob["terrain_seed"] = np.array([42], dtype=np.uint64)
    blob["terrain_vox"] = np.stack([S.terrain_chunk(42, *map(int, c)) for c in tcoords])
    # mesh cases
    cases = build_cases()
    blob["case_names"] = np.array(sorted(cases))
    for name, (vox, nbr) in cases.items():
        blob[f"{name}/vox"], blob[f"{name}/nbr"] = vox, nbr
        for mode, greedy in (("culled", False), ("greedy", True)):
            keys, offs = [], [0]
            for i in range(vox.shape[0]):
                nbs = [None if nbr[i, f] < 0 else vox[nbr[i, f]] for f in range(6)]
                k = S.mesh_chunk_keys(vox[i], nbs, greedy)
                keys.extend(k)
                offs.append(len(keys))
            blob[f"{name}/{mode}/keys"] = np.array(keys, dtype=np.uint64)
            blob[f"{name}/{mode}/offs"] = np.array(offs, dtype=np.int64)
            print(name, mode, offs[-1], flush=True)
    # vertex expansion of a few keys (SPEC §5)
    sample = blob["slab_stripes/greedy/keys"][:64]
    blob["expand_keys"] = sample
    ev, ei = [], []
    for q, k in enumerate(sample):
        v, i = S.expand_quad(int(k), q)
        ev.append(v)
        ei.append(i)
    blob["expand_verts"] = np.array(ev, dtype=np.uint64)
    blob["expand_idx"] = np.array(ei, dtype=np.uint32)
    np.savez_compressed(OUT, **blob)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
