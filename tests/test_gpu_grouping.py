"""Instance grouping kernel on its own (csrc/group.cu behind isx_dbscan_fit_host), against the working
definition of the reference's ML::dbscanFit call (oracle/dbscan_def.h, Stixels.cu:660-666): labels must be
IDENTICAL, not just the partition -- both number clusters by their lowest-index core point and attach border
points to their lowest-index core neighbour."""
import numpy as np
import pytest

from instance_stixels_b200 import api
from oracle import cpubind

pytestmark = pytest.mark.gpu


def _blobs(rng, n, k, spread, extent=2000.0):
    centres = rng.uniform(0, extent, size=(k, 2))
    xy = centres[rng.integers(0, k, n)] + rng.normal(0, spread, size=(n, 2))
    return xy.astype(np.float32)


@pytest.mark.parametrize("n,k,spread,eps,min_pts,p_cand", [
    (1, 1, 1.0, 10.0, 1, 1.0),          # single point
    (2, 1, 1.0, 10.0, 3, 1.0),          # fewer points than min_pts: all noise
    (37, 3, 5.0, 18.8, 3, 0.7),
    (500, 12, 8.0, 18.82232269133926, 3, 0.6),      # pairwise preset
    (1500, 20, 10.0, 23.89408062110343, 4, 0.5),    # unary preset
    (3000, 2, 6.0, 23.9, 4, 0.9),       # two dense blobs: nearly every pair is a neighbour
    (4096, 40, 15.0, 18.8, 3, 0.5),     # exactly the shared-memory staging capacity
    (4097, 40, 15.0, 18.8, 3, 0.5),     # one more: the global-memory path
    (6000, 5, 30.0, 12.0, 5, 0.3),      # global-memory path, chains of border points
    (800, 800, 0.0, 5.0, 2, 1.0),       # isolated points: everything is noise
])
@pytest.mark.parametrize("threads", ["256", "1024"])   # the batch and the single-frame launch shape
def test_labels_equal_the_definition(n, k, spread, eps, min_pts, p_cand, threads, monkeypatch):
    monkeypatch.setenv("ISX_GROUP_THREADS", threads)
    rng = np.random.default_rng(n * 7919 + k)
    xy = _blobs(rng, n, k, spread)
    cand = (rng.random(n) < p_cand).astype(np.uint8)
    want = cpubind.dbscan(xy, eps, min_pts, cand)
    got = api.dbscan_fit(xy, eps, min_pts, cand)
    assert np.array_equal(got, want), (int((got != want).sum()), n)


def test_chain_is_one_cluster_and_duplicates_are_neighbours():
    # a 2000-point chain with spacing just under eps: one component whose union needs the whole forest
    n = 2000
    xy = np.stack([np.arange(n, dtype=np.float32) * 9.5, np.zeros(n, np.float32)], axis=1)
    perm = np.random.default_rng(5).permutation(n)
    xy = xy[perm]
    cand = np.ones(n, np.uint8)
    got = api.dbscan_fit(xy, 10.0, 3, cand)
    assert np.array_equal(got, cpubind.dbscan(xy, 10.0, 3, cand))
    assert got.max() == 0 and (got == 0).sum() >= n - 2
    # identical points (stixels of one object voting for the same centre)
    xy = np.tile(np.array([[100.0, 50.0]], np.float32), (300, 1))
    got = api.dbscan_fit(xy, 1.0, 4, np.ones(300, np.uint8))
    assert np.all(got == 0)
    # the size filter: no candidate -> no core point -> all noise
    assert np.all(api.dbscan_fit(xy, 1.0, 4, np.zeros(300, np.uint8)) == -1)


def test_empty_and_bad_arguments():
    assert len(api.dbscan_fit(np.zeros((0, 2), np.float32), 10.0, 3, np.zeros(0, np.uint8))) == 0
    with pytest.raises(api.InvalidArgument):
        api.dbscan_fit(np.zeros((4, 2), np.float32), 10.0, 3, np.zeros(3, np.uint8))
