"""The robust-loss helpers of the single-fit path (lsqfit_b200/dense.py: _rho, DenseFit._cost, DenseFit._robust_rows)
against scipy's own functions (scipy/optimize/_lsq/least_squares.py: construct_loss_function; common.py:
scale_for_robust_loss_function) -- the arithmetic lsqfit.scipy_least_squares(loss=, f_scale=) runs
(reference src/lsqfit/_scipy.py:77, 156-161).  torch CPU tensors; no GPU needed."""
import numpy as np
import pytest
import torch


@pytest.mark.parametrize("loss", ["huber", "soft_l1", "cauchy", "arctan"])
@pytest.mark.parametrize("f_scale", [0.7, 2.0])
def test_robust_rows_and_cost_match_scipy(loss, f_scale):
    from scipy.optimize._lsq.common import scale_for_robust_loss_function
    from scipy.optimize._lsq.least_squares import construct_loss_function
    from lsqfit_b200 import dense
    rng = np.random.default_rng(3)
    f = 3.0 * rng.standard_normal(60)
    f[:4] = [0.0, 1e-9, f_scale, -f_scale]                    # the kinks of huber, and z = 0
    J = rng.standard_normal((60, 4))
    lf = construct_loss_function(60, loss, f_scale)
    rho = lf(f.copy())
    cost = lf(f.copy(), cost_only=True)
    Js, fs = scale_for_robust_loss_function(J.copy(), f.copy(), rho)

    class Stub(object):
        pass
    d = Stub()
    d.loss, d.f_scale, d.robust = loss, f_scale, True
    tf = torch.as_tensor(f)
    js, fsc = dense.DenseFit._robust_rows(d, tf)
    # split the residuals into a "data" and a "prior" part: the cost is the sum over both
    c = dense.DenseFit._cost(d, tf[:45], tf[45:])
    assert abs(c - cost) <= 1e-14 * abs(cost)
    # huber beyond its kink: rho' + 2 rho'' z vanishes identically, the computed value is rounding noise around the EPS
    # clamp -- rows scaled by ~1.5e-8 either way (atol); everything else agrees to rounding
    np.testing.assert_allclose(js.numpy()[:, None] * J, Js, rtol=1e-13, atol=1e-7)
    noise = (js.numpy() < 1e-7)
    np.testing.assert_allclose(fsc.numpy()[~noise], fs[~noise], rtol=1e-13, atol=1e-300)
    np.testing.assert_allclose(js.numpy()[noise] * fsc.numpy()[noise], (Js[noise, 0] / J[noise, 0]) * fs[noise], rtol=1e-12)
    d.robust = False
    assert abs(dense.DenseFit._cost(d, tf[:45], tf[45:]) - 0.5 * f @ f) <= 1e-14 * (f @ f)


def test_unknown_loss_is_refused():
    from lsqfit_b200 import dense
    with pytest.raises(ValueError):
        dense._rho("nonsense", torch.zeros(3, dtype=torch.float64))
