"""Edge tags on the device: b200_unique_edges against the numpy restatement of the reference's first-appearance
order (feng_b200/numbering.py:build_edges, itself pinned on the compiled reference by tests/test_host_tables.py) -- bit-identical
edge lists and cell->edge tables on Kuhn meshes (boundary triangles swept first), 2-D meshes and unstructured tetrahedra; and the
complete numbering / element->DOF tables built on top of it."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("which", ["t2d", "t3d", "delaunay", "t3d_big"])
def test_device_edge_tags_are_bit_identical(which):
    from feng_b200 import mesh as M, numbering as NB
    m = {"t2d": lambda: M.square_mesh(37), "t3d": lambda: M.cube_mesh(7), "delaunay": lambda: M.unstructured_tet_mesh(6, seed=9),
         "t3d_big": lambda: M.cube_mesh(40)}[which]()
    e_h, c_h = NB.build_edges(m, device=None)
    e_d, c_d = NB.build_edges(m, device=0)
    assert e_h.shape == e_d.shape and np.array_equal(e_h, e_d)
    assert np.array_equal(c_h, c_d)


def test_numbering_on_device_edges_gives_the_same_tables(monkeypatch):
    from feng_b200 import mesh as M, numbering as NB
    m = M.cube_mesh(9)
    sp = NB.taylor_hood_spaces(3, False, True)
    monkeypatch.setenv("B200_EDGES", "host")
    a = NB.build_numbering(m, sp)
    monkeypatch.setenv("B200_EDGES", "device")
    b = NB.build_numbering(m, sp)
    assert a.n_inc == b.n_inc and a.n_dof == b.n_dof
    assert np.array_equal(a.adr(m, "U", 2), b.adr(m, "U", 2)) and np.array_equal(a.adr(m, "P", 1), b.adr(m, "P", 1))
