"""Observation-operator generation on the device (oak_b200/csrc/hgen.cu: batched cinterp, ndgrid.F90:1183-1257) against
the oracle's restatement (oracle/oak_ndgrid.c, pinned on test/test_ndgrid.F90 in tests/test_ndgrid_oracle.py), through
the C ABI (oakb200_cinterp)."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ob():
    import oak_b200
    return oak_b200


def _axes(rng, gshape, descending=()):
    axes = []
    for k, g in enumerate(gshape):
        a = np.cumsum(rng.uniform(0.5, 1.5, g)) + 10.0 * k      # rectilinear, unequal spacing
        if k in descending:
            a = a[::-1].copy()
        axes.append(a)
    return axes


def _points(rng, axes, m, frac_out=0.05):
    n = len(axes)
    lo = np.array([a.min() for a in axes]); hi = np.array([a.max() for a in axes])
    xi = lo + (hi - lo) * rng.uniform(0, 1, (m, n))
    k = int(m * frac_out)
    xi[:k] += (hi - lo) * rng.choice([-1.0, 1.0], (k, n)) * rng.uniform(0.0, 0.6, (k, n))   # some outside
    return xi


@pytest.mark.parametrize("gshape,descending", [((40,), ()), ((17, 13), ()), ((17, 13), (1,)), ((9, 8, 7), ()),
                                               ((9, 8, 7), (2,)), ((5, 4, 3, 6), ()), ((2, 2, 2, 2), ())])
def test_cinterp_matches_the_oracle(ob, gshape, descending):
    rng = np.random.default_rng(sum(gshape) + len(descending))
    axes = _axes(rng, gshape, descending)
    n = len(gshape)
    m = 3000
    xi = _points(rng, axes, m)
    masked = (rng.uniform(size=int(np.prod(gshape))) < 0.05).astype(np.uint8) if np.prod(gshape) > 100 else None
    h = ob.Handle(0)
    idx, co, nbp = h.cinterp(gshape, axes, xi, masked)
    h.close()
    coord = oracle.ndgrid_full_coords(gshape, axes=axes)
    idx0, co0, nbp0 = oracle.cinterp(gshape, coord, xi, masked=masked)
    assert np.array_equal(nbp, nbp0)
    assert (nbp == 0).sum() > 0 and (nbp == 2 ** n).sum() > m // 4
    ins = nbp > 0
    # interior points: the same cell (1-based corner subscripts) and the same weights
    assert np.array_equal(idx[ins], idx0[ins])
    assert np.abs(co[ins] - co0[ins]).max() < 1e-12
    assert np.abs(co[ins].sum(axis=1) - 1.0).max() < 1e-12
    # the weights reproduce a linear field exactly (test/test_ndgrid.F90: fun_nd)
    ioff = np.concatenate([[1], np.cumprod(np.array(gshape[:-1], dtype=np.int64))])
    lin = ((idx[ins].astype(np.int64) - 1) * ioff).sum(axis=2)
    f = sum(2 * (k + 1) * coord[k] for k in range(n))
    ref = sum(2 * (k + 1) * xi[ins][:, k] for k in range(n))
    assert np.abs((co[ins] * f[lin]).sum(axis=1) - ref).max() < 1e-9 * max(1.0, np.abs(ref).max())


def test_cinterp_points_on_nodes_and_faces(ob):
    # a point on a shared face lies in several cells: the reference's tree visits upper halves first, i.e. finds the cell
    # with the highest subscripts; on the last node of an axis that is the last cell.  Same cell as the oracle's tree,
    # same interpolated value whatever the cell.
    gshape = (6, 5, 4)
    axes = [np.arange(6.0), 10.0 + 2.0 * np.arange(5.0), np.array([0.0, -5.0, -20.0, -100.0])]   # depth descending
    g = np.stack(np.meshgrid(*axes, indexing="ij"), axis=-1).reshape(-1, 3)                     # every node
    rng = np.random.default_rng(3)
    faces = g.copy()
    faces[:, 0] = np.clip(faces[:, 0] + rng.uniform(-0.4, 0.4, len(faces)), 0, 5)              # on faces in y and z
    xi = np.concatenate([g, faces])
    h = ob.Handle(0)
    idx, co, nbp = h.cinterp(gshape, axes, xi)
    h.close()
    coord = oracle.ndgrid_full_coords(gshape, axes=axes)
    idx0, co0, nbp0 = oracle.cinterp(gshape, coord, xi)
    assert (nbp == 8).all() and np.array_equal(nbp, nbp0)
    assert np.array_equal(idx, idx0)
    assert np.abs(co - co0).max() < 1e-12


def test_cinterp_degenerate_cells_fail_loudly(ob):
    # singleton dimension: the reference goes through the SVD branch of interp_tetrahedron (ndgrid.F90:527-627), which
    # is not on the device: status -6, nbp = -1 for those observations, the count is reported
    h = ob.Handle(0)
    with pytest.raises(ob.OakB200Error) as e:
        h.cinterp((4, 1), [np.arange(4.0), np.array([2.0])], np.array([[1.5, 2.0], [2.5, 2.0], [1.5, 2.5]]))
    assert e.value.code == -6
    idx, co, nbp, ndeg = h.last_cinterp
    assert ndeg == 2 and list(nbp) == [-1, -1, 0]
    h.close()


def test_gen_observation_oper_rows(ob):
    # the COO triplets genObservationOper builds from cinterp (assimilation.F90:2587-2611): 2^n entries per observation
    # inside the grid, one zero entry with model index -1 otherwise; H f reproduces a linear field at the observations
    import oak_b200
    gshape = (12, 10)
    axes = [np.linspace(0, 11, 12), np.linspace(40, 49, 10)]
    rng = np.random.default_rng(5)
    xi = np.column_stack([rng.uniform(-1, 12, 200), rng.uniform(39, 50, 200)])
    h = ob.Handle(0)
    Hi, Hj, Hs = oak_b200.gen_observation_oper(h, gshape, axes, xi)
    h.close()
    inside = (xi[:, 0] >= 0) & (xi[:, 0] <= 11) & (xi[:, 1] >= 40) & (xi[:, 1] <= 49)
    assert (Hj == -1).sum() == (~inside).sum() and len(Hi) == 4 * inside.sum() + (~inside).sum()
    X, Y = np.meshgrid(*axes, indexing="ij")
    f = (3 * X - 2 * Y).ravel(order="F")
    Hf = np.zeros(200)
    ok = Hj > 0
    np.add.at(Hf, Hi[ok] - 1, Hs[ok] * f[Hj[ok] - 1])
    assert np.abs(Hf[inside] - (3 * xi[inside, 0] - 2 * xi[inside, 1])).max() < 1e-10
    assert (Hf[~inside] == 0).all()
