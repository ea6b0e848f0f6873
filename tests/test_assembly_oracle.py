"""Pins oracle/assembly_oracle.py against golden vectors produced by the reference's own Python
(tools/make_golden_assembly.py; reference furniture.py:1057-1153, transform_utils.py:633-664)."""
import os

import numpy as np

from oracle import assembly_oracle as A

G = os.path.join(os.path.dirname(__file__), "golden")


def test_is_aligned_matches_reference_bit_exact():
    z = np.load(os.path.join(G, "is_aligned.npz"))
    n = len(z["aligned"])
    assert n >= 5000 and 0.2 < z["aligned"].mean() < 0.8
    for i in range(0, n, 4):  # every 4th case keeps the CPU suite short; the GPU test runs all of them
        ang = z["angles"][i][: z["nangles"][i]]
        ok, tq = A.is_aligned(z["p1"][i], z["m1"][i], z["p2"][i], z["m2"][i], ang, z["thr"][i])
        assert ok == bool(z["aligned"][i]), i
        assert (tq is not None) == bool(z["tq_set"][i]), i
        if tq is not None:
            assert np.array_equal(tq, z["tq"][i]), i  # same numpy ops in the same order: bit-identical


def test_connect_geometry_matches_reference():
    z = np.load(os.path.join(G, "connect_geom.npz"))
    for i in range(0, len(z["qb"]), 4):
        p, q = A.transform_to_target_quat(z["qb"][i], z["q"][i], z["tq"][i])
        assert np.allclose(p, z["new_pos"][i], rtol=0, atol=1e-14)
        assert np.allclose(q, z["new_quat"][i], rtol=0, atol=1e-14)
        assert np.allclose(A.rel_pose(z["qb"][i], z["q"][i]), z["rel"][i], rtol=0, atol=1e-14)
        assert np.allclose(A.euler_to_quat(z["eul"][i], z["tq"][i]), z["eq"][i], rtol=0, atol=1e-14)


def test_connect_masks():
    ct, ca = A.connect_masks(0)
    assert ct == (1 << 30) - 1 - 2 and ca == 2
    ct3, ca3 = A.connect_masks(3)
    assert (ct & ca) == 0 and (ct3 & ca) != 0 and (ct & ca3) != 0  # merged group never self-collides; collides with others
    assert (ct & 1) != 0  # still collides with default conaffinity=1 geoms (floor, robot)
