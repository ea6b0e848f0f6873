"""Mesh geoms that only carry mass (non-colliding, density > 0: 4 furniture models incl. toy_table): the MJCF compiler
integrates the STL solid exactly (signed tetrahedra).  Known answer: a box mesh must give the box formulas."""
import numpy as np

from furniture_b200 import mjcf


def _box_mesh(a, b, c, centre):
    V = np.array([[x, y, z] for x in (-a, a) for y in (-b, b) for z in (-c, c)]) + np.asarray(centre)
    faces = [(0, 1, 3), (0, 3, 2), (4, 6, 7), (4, 7, 5), (0, 4, 5), (0, 5, 1), (2, 3, 7), (2, 7, 6), (0, 2, 6), (0, 6, 4), (1, 5, 7), (1, 7, 3)]
    return np.array([[V[i], V[j], V[k]] for i, j, k in faces])


def test_mesh_mass_properties_of_a_box():
    a, b, c = 0.5, 0.3, 0.2
    for tri in (_box_mesh(a, b, c, [1.0, 2.0, 3.0]), _box_mesh(a, b, c, [1.0, 2.0, 3.0])[:, ::-1]):  # both orientations
        m, com, I = mjcf.mesh_mass_properties(tri, 1000.0)
        assert abs(m - 1000 * 8 * a * b * c) < 1e-9
        assert np.abs(com - [1, 2, 3]).max() < 1e-12
        assert np.abs(I - np.diag(m / 3 * np.array([b * b + c * c, a * a + c * c, a * a + b * b]))).max() < 1e-9


def test_toy_table_compiles_with_its_mesh_inertia():
    m = mjcf.load_scene("Sawyer", "toy_table")  # BASELINE.json config 1 furniture; part 1 carries a density-1 mesh geom
    assert m.nv == 39 and len(m.meta["part_names"]) == 5
    b = m.names["body"].index("1_part1")
    assert 0.01 < m.body_mass[b] < 0.2 and (m.body_inertia[b] > 0).all()
