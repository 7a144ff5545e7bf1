"""Host loops of R/optimWrapper.R / R/GIST.R without a GPU: the BFGS minimiser and the lasso helpers."""
import numpy as np
import pytest

from psydiff_b200 import optim


def test_vmmin_rosenbrock_and_quadratic():
    f = lambda x: (1 - x[0]) ** 2 + 100 * (x[1] - x[0] ** 2) ** 2
    g = lambda x: np.array([-2 * (1 - x[0]) - 400 * x[0] * (x[1] - x[0] ** 2), 200 * (x[1] - x[0] ** 2)])
    out = optim.vmmin([-1.2, 1.0], f, g, maxit=200)
    assert out["convergence"] == 0 and np.allclose(out["par"], [1, 1], atol=1e-3)   # optim(c(-1.2,1), fr, grr, method="BFGS") reaches ~9.6e-18
    A = np.array([[3.0, 1.0], [1.0, 2.0]])
    out = optim.vmmin([5.0, -3.0], lambda x: 0.5 * x @ A @ x, lambda x: A @ x)
    assert np.allclose(out["par"], 0, atol=1e-5)
    with pytest.raises(Exception, match="not finite"):
        optim.vmmin([0.0], lambda x: np.nan, lambda x: np.zeros(1))


def test_reg_value_and_subgradients():
    theta = {"a": 0.5, "b": -2.0, "c": 0.0}
    w = {"a": 1.0, "b": 2.0, "c": 1.0}
    assert optim.getRegValue(10, theta, ["b", "c"], w) == pytest.approx(40.0)
    assert optim.getRegValue(10, theta, ["a"]) == pytest.approx(5.0)
    sub = optim.getSubgradients(theta, {"a": 1.0, "b": 3.0, "c": 4.0}, ["b", "c"], 10, w)
    assert sub["a"] == 1.0 and sub["b"] == pytest.approx(3.0 - 20.0) and sub["c"] == 0.0      # |4| < 10 -> inside the interval
    sub = optim.getSubgradients(theta, {"a": 1.0, "b": 3.0, "c": 14.0}, ["c"], 10, w)
    assert sub["c"] == pytest.approx(4.0)


def test_data_template():
    d = optim.getDataTemplate(3, 2, 4, 0.5)
    assert d["observations"].shape == (12, 2) and list(d["dt"][:5]) == [0, .5, .5, .5, 0] and list(d["person"][:5]) == [1, 1, 1, 1, 2]
    with pytest.raises(Exception, match="length of dts"):
        optim.getDataTemplate(3, 2, 4, [1, 2])
