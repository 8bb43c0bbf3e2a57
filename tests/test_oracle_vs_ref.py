"""Larger randomized cross-check of the restatement against the reference's own code.
Runs only where oracle/_ref was built (this container; the GPU box gets the prebuilt .so)."""
import os
import tempfile

import numpy as np
import pytest

from ctopprm_learn_b200 import synth


@pytest.fixture(scope="module")
def pair(orc, c1_map):
    if not orc.ref_available():
        pytest.skip("oracle/_ref not built (reference tree not mounted)")
    vox, c, e = c1_map
    fn = tempfile.mktemp(suffix=".npy")
    synth.write_map(fn, vox, c, e)
    ref = orc.RefMap(fn)
    yield orc.OracleMap(vox, c, e), ref, c, e
    ref.close()
    os.unlink(fn)


def test_clearance_and_gradient(pair):
    om, ref, c, e = pair
    p = c + (np.random.default_rng(5).random((100000, 3)) - 0.5) * (e + 0.3)
    assert np.array_equal(om.clearance(p), ref.clearance(p), equal_nan=True)
    g1, c1 = om.gradient(p[:20000])
    g2, c2 = ref.gradient(p[:20000])
    assert np.array_equal(g1, g2, equal_nan=True) and np.array_equal(c1, c2, equal_nan=True)


@pytest.mark.parametrize("clr,cdc", [(0.2, 0.05), (0.35, 0.02), (0.0, 0.3)])
def test_edges(pair, clr, cdc):
    om, ref, c, e = pair
    a, b = synth.random_edges(c, e, 40000, 9, lo=0.0, hi=3.0)
    f1, h1, _, _, _ = om.edge_nodes(a, b, clr, cdc)
    f2, h2 = ref.edge_nodes(a, b, clr, cdc)
    assert np.array_equal(f1, f2) and np.array_equal(h1, h2, equal_nan=True)
    assert np.array_equal(om.edge_guards(a, b, clr, cdc)[0], ref.edge_guards(a, b, clr, cdc))


def test_map_file_roundtrip_f64(orc):
    # fill_data<double> narrows to float (include/esdf_map.hpp:43,55-66)
    vox, c, e = synth.random_columns_esdf(16, 1.6, 3, 4)
    fn = tempfile.mktemp(suffix=".npy")
    synth.write_map(fn, vox.astype(np.float64) * (1 + 1e-9), c, e)
    ref = orc.RefMap(fn)
    om = orc.OracleMap((vox.astype(np.float64) * (1 + 1e-9)).astype(np.float32), c, e)
    p = c + (np.random.default_rng(1).random((2000, 3)) - 0.5) * e
    assert np.array_equal(om.clearance(p), ref.clearance(p), equal_nan=True)
    ref.close()
    os.unlink(fn)
