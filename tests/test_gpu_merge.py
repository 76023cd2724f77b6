"""GPU: K4 hits merge == numpy lexsort by (distance,row)."""
import numpy as np
import pytest

from semtools_b200 import capi

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n_lists,per,k", [(1, 1, 1), (2, 10, 10), (8, 10, 10), (8, 10, 3), (3, 7, 30), (64, 64, 96)])
def test_merge_matches_lexsort(ctx, n_lists, per, k):
    rng = np.random.default_rng(n_lists * 100 + per)
    lists = np.zeros((n_lists, per), dtype=capi.HIT_DTYPE)
    lists["distance"] = rng.choice(np.linspace(0, 2, 50), (n_lists, per))     # many ties
    lists["row"] = rng.permutation(n_lists * per).reshape(n_lists, per)
    lists["distance"][0, -1] = np.inf                                          # padding entry
    lists["row"][0, -1] = np.uint64(0xFFFFFFFFFFFFFFFF)
    flat = lists.reshape(-1)
    valid = flat[flat["row"] != np.uint64(0xFFFFFFFFFFFFFFFF)]
    order = np.lexsort((valid["row"], valid["distance"]))
    exp = valid[order][:k]
    got = ctx.hits_merge(lists, k)
    assert np.array_equal(got, exp)
