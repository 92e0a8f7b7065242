"""Pins the oracle's restatement of lap.lapjv(extend_cost=True, cost_limit=thresh) (lap 0.4.0 is
not under /root/reference): brute force over all partial matchings on small cases, and
scipy.optimize.linear_sum_assignment on lap's extended matrix on larger ones."""
import itertools

import numpy as np
import pytest

from oracle import oracle_np as O


def _brute(cost, thresh):
    n, m = cost.shape
    best, best_x = 0.0, tuple([-1] * n)
    cols = list(range(m)) + [-1] * n
    seen = set()
    for perm in itertools.permutations(cols, n):
        if perm in seen:
            continue
        seen.add(perm)
        val = sum(cost[i, j] - thresh for i, j in enumerate(perm) if j >= 0)
        if val < best - 1e-15:
            best, best_x = val, perm
    return best, np.array(best_x)


@pytest.mark.parametrize("seed", range(40))
def test_objective_equals_brute_force(seed):
    rng = np.random.default_rng(seed)
    n, m = int(rng.integers(1, 5)), int(rng.integers(1, 5))
    cost = rng.uniform(0, 1.2, (n, m))
    thresh = float(rng.choice([0.5, 0.7, 0.8]))
    best, bx = _brute(cost, thresh)
    for solver in ("jv", "scipy"):
        x, y = O.lapjv_extended(cost, thresh, solver)
        assert abs(O.assignment_objective(cost, thresh, x) - best) <= 1e-12
        np.testing.assert_array_equal(x, bx)
        for i, j in enumerate(x):
            if j >= 0:
                assert y[j] == i and cost[i, j] < thresh


@pytest.mark.parametrize("n,m,density", [(30, 30, 1.0), (64, 48, 0.3), (200, 220, 0.05), (400, 400, 1.0)])
def test_jv_port_equals_scipy_on_extended_matrix(n, m, density):
    rng = np.random.default_rng(n + m)
    cost = rng.uniform(0, 1, (n, m))
    cost[rng.uniform(size=(n, m)) > density] = 1.0
    for thresh in (0.8, 0.5):
        x, y = O.lapjv_extended(cost, thresh, "jv")
        xs, ys = O.lapjv_extended(cost, thresh, "scipy")
        np.testing.assert_array_equal(x, xs)
        np.testing.assert_array_equal(y, ys)


def test_linear_assignment_quirks():
    m, ua, ub = O.linear_assignment(np.zeros((0, 4)), 0.8)
    assert m.shape == (0, 2) and ua == () and ub == (0, 1, 2, 3)
    m, ua, ub = O.linear_assignment(np.ones((2, 2)), 0.8)
    assert m.shape == (0,) and list(ua) == [0, 1] and list(ub) == [0, 1]
