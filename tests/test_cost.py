"""StaticGoalQuadraticCost, mirroring the reference's tests/trajopt/test_cost.py:10-55: the cost
against an explicit for-loop, closed-form grad / hess against autodiff of the base class."""
import torch

from ambersim_b200.trajopt.base import CostFunction, CostFunctionParams
from ambersim_b200.trajopt.cost import StaticGoalQuadraticCost


def test_sgqc():
    nx, nu, N = 16, 4, 10  # the Barrett hand's sizes
    cost_function = StaticGoalQuadraticCost(Q=torch.eye(nx), Qf=10.0 * torch.eye(nx), R=0.01 * torch.eye(nu), xg=torch.zeros(nx))
    g = torch.Generator().manual_seed(0)
    xs = torch.randn((N + 1, nx), generator=g)
    us = torch.randn((N, nu), generator=g)

    val_test, _ = cost_function.cost(xs, us, params=CostFunctionParams())
    val_gt = 0.0
    for i in range(N):
        xs_err = xs[i] - cost_function.xg
        val_gt += 0.5 * (xs_err @ cost_function.Q @ xs_err + us[i] @ cost_function.R @ us[i])
    xs_err = xs[N] - cost_function.xg
    val_gt += 0.5 * (xs_err @ cost_function.Qf @ xs_err)
    assert torch.allclose(val_test, val_gt)

    gx, gu, _, _ = cost_function.grad(xs, us, params=CostFunctionParams())
    gx_gt, gu_gt, _, _ = CostFunction.grad(cost_function, xs, us, CostFunctionParams())
    assert torch.allclose(gx, gx_gt, atol=1e-5) and torch.allclose(gu, gu_gt, atol=1e-6)

    hxx, hxu, _, huu, _, _, _ = cost_function.hess(xs, us, params=CostFunctionParams())
    hxx_gt, hxu_gt, _, huu_gt, _, _, _ = CostFunction.hess(cost_function, xs, us, CostFunctionParams())
    assert torch.allclose(hxx, hxx_gt, atol=1e-5)
    assert torch.allclose(hxu, hxu_gt, atol=1e-5)
    assert torch.allclose(huu, huu_gt, atol=1e-5)


def test_sgqc_dense_weights_and_batch():
    g = torch.Generator().manual_seed(1)
    nx, nu, N = 5, 2, 4
    A = torch.randn((nx, nx), generator=g)
    cf = StaticGoalQuadraticCost(Q=A @ A.T, Qf=2 * A @ A.T, R=torch.eye(nu), xg=torch.randn(nx, generator=g))
    xs, us = torch.randn((3, N + 1, nx), generator=g), torch.randn((3, N, nu), generator=g)
    batched, _ = cf.cost(xs, us, None)
    single = torch.stack([cf.cost(xs[i], us[i], None)[0] for i in range(3)])
    assert torch.allclose(batched, single)
