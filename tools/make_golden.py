#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REFERENCE'S OWN code (oracle/_ref: the
reference's src/esdf_map.cpp + src/base_map.cpp compiled against shim headers) on small
seeded inputs.  Only runs where /root/reference is mounted (this container); the vectors it
writes are committed so the GPU box can check against them without the reference tree.

    python tools/make_golden.py
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ctopprm_learn_b200 import synth  # noqa: E402
from oracle import orc  # noqa: E402

CLR, CDC = synth.MIN_CLEARANCE, synth.COLLISION_DISTANCE_CHECK


def small_map(seed, n=32, extent=3.2, n_col=6, rect=False):
    vox, c, e = synth.random_columns_esdf(n, extent, n_col, seed)
    rng = np.random.default_rng(seed + 1000)
    vox = (vox + 0.02 * rng.standard_normal(vox.shape)).astype(np.float32)  # break z-invariance
    if rect:  # per-axis extents differ from the cubic index scaling (esdf_map.cpp:131-137 quirk)
        e = e * np.array([1.0, 0.8, 0.6])
    return vox, c, e


def polyline(rng, c, e, k):
    p = c + (rng.random((k, 3)) - 0.5) * (e * 0.8)
    return p


def main():
    orc.build()
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name, seed, rect in (("cube", 11, False), ("rect", 12, True)):
        vox, c, e = small_map(seed, rect=rect)
        fn = tempfile.mktemp(suffix=".npy")
        synth.write_map(fn, vox, c, e)
        ref = orc.RefMap(fn)
        rng = np.random.default_rng(seed)
        # points: inside, slightly outside, exact integer index coordinates (0/0 quirk), NaN
        pts = c + (rng.random((4000, 3)) - 0.5) * (e + 0.4)
        n = vox.shape[0]
        ints = c - e.max() / 2 + rng.integers(0, n, (64, 3)) * (e.max() / n)
        pts = np.concatenate([pts, ints, [[np.nan, 0, 0]], [c], [c - e / 2], [c + e / 2]])
        clear = ref.clearance(pts)
        grad, centre = ref.gradient(pts[:1500])
        coll = ref.in_collision(pts, CLR)
        a, b = synth.random_edges(c, e, 3000, seed, lo=0.01, hi=2.5)
        # degenerate: zero length, shorter than cdc, leaving the map
        a[:4] = b[:4]
        b[4:8] = a[4:8] + 0.01
        b[8:12] = a[8:12] + e
        free_nodes, hit = ref.edge_nodes(a, b, CLR, CDC)
        free_guards = ref.edge_guards(a, b, CLR, CDC)
        free_nodes2, _ = ref.edge_nodes(a, b, 0.05, 0.11, want_hit=False)
        # polylines
        polys = [polyline(rng, c, e, k) for k in (2, 3, 5, 9, 4, 7)]
        lens = [float(orc.path_length(p)) for p in polys]
        nums = [float(np.ceil(L / CDC) + 1) for L in lens]
        sampled = [ref.sample_path(p, L, k) for p, L, k in zip(polys, lens, nums)]
        pairs = [(i, j) for i in range(len(polys)) for j in range(i + 1, len(polys))]
        deform = []
        for i, j in pairs:
            k = float(np.ceil(max(lens[i], lens[j]) / CDC) + 1)
            deform.append(ref.deformable(polys[i], lens[i], polys[j], lens[j], k, 0.02, CDC))
        # polylines that hug free space (so some pairs ARE deformable): jittered copies
        base = polyline(rng, c, e * 0.3, 4)
        near = [base + 0.01 * rng.standard_normal(base.shape) for _ in range(3)]
        nlen = [float(orc.path_length(p)) for p in near]
        ndef = []
        for i in range(3):
            for j in range(i + 1, 3):
                k = float(np.ceil(max(nlen[i], nlen[j]) / CDC) + 1)
                ndef.append(ref.deformable(near[i], nlen[i], near[j], nlen[j], k, -10.0, CDC))
        pcf = [ref.path_collision_free(p, CLR, CDC) for p in polys]
        quad = polyline(rng, c, e * 0.5, 4)
        dbt = ref.deformable_between(quad[0], quad[1], quad[2], quad[3], -10.0, CDC)
        dbt2 = ref.deformable_between(quad[0], quad[1], quad[2], quad[3], CLR, CDC)
        u = 0.37
        g = np.array([0.3, -1.2, 0.8])
        ell = ref.fill_random_state_in_ellipse(quad[0], quad[1], 1.2, False, u, g)
        ell_planar = ref.fill_random_state_in_ellipse(quad[0], quad[1], 1.2, True, u, g)
        np.savez_compressed(
            os.path.join(out_dir, f"ref_{name}.npz"),
            vox=vox, centroid=c, extents=e, clr=CLR, cdc=CDC,
            min_pos=ref.min_pos(), max_pos=ref.max_pos(),
            pts=pts, clearance=clear, grad=grad, centre=centre, in_collision=coll,
            a=a, b=b, free_nodes=free_nodes, hit=hit, free_guards=free_guards, free_nodes_alt=free_nodes2,
            poly_pts=np.concatenate(polys), poly_off=np.cumsum([0] + [len(p) for p in polys]),
            poly_len=np.array(lens), poly_num=np.array(nums), sampled=np.concatenate(sampled),
            pair_i=np.array([p[0] for p in pairs]), pair_j=np.array([p[1] for p in pairs]),
            deformable=np.array(deform), near_pts=np.concatenate(near), near_len=np.array(nlen),
            near_deformable=np.array(ndef), path_free=np.array([r[0] for r in pcf]),
            path_free_hit=np.array([r[1] for r in pcf]), quad=quad, deformable_between=np.array([dbt, dbt2]),
            ell_u=u, ell_g=g, ell=ell, ell_planar=ell_planar)
        ref.close()
        os.unlink(fn)
        print("wrote", name, os.path.getsize(os.path.join(out_dir, f"ref_{name}.npz")), "bytes",
              "free frac", free_nodes.mean(), "nan frac", np.isnan(clear).mean())


if __name__ == "__main__":
    main()
