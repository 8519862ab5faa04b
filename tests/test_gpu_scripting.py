"""TorchScript survives the drop-in (compressai tests/test_scripting.py:37-59 scripts GDN): the scripted
modules are one call of a registered ``torch.library`` op (mmcodec/library.py) on the same Parameters, equal to the eager kernel path
in value and in gradient."""
import pytest
import torch

pytestmark = pytest.mark.gpu

import mmcodec  # noqa: E402


def dev():
    return torch.device("cuda", 0)


@pytest.mark.parametrize("inverse", [False, True])
def test_scripted_gdn_equals_eager(inverse):
    # the reference's own case: GDN(128) on a (1, 128, 1, 1) input, plus a spatial one
    torch.manual_seed(0)
    g = mmcodec.GDN(128, inverse=inverse).to(dev())
    with torch.no_grad():
        g.gamma.add_(0.02 * torch.rand_like(g.gamma))
    m = torch.jit.script(g)
    for shape in ((1, 128, 1, 1), (2, 128, 9, 7)):
        x = torch.rand(shape, device=dev())
        y0 = g(x)
        y1 = m(x)
        assert y1.shape == y0.shape and torch.equal(y0, y1)
    # closed form of the reference's test_layers.py:145-146 for the default initialisation
    g0 = torch.jit.script(mmcodec.GDN(32, inverse=inverse).to(dev()))
    x = torch.rand(1, 32, 4, 4, device=dev())
    ref = x * torch.sqrt(1 + 0.1 * x ** 2) if inverse else x / torch.sqrt(1 + 0.1 * x ** 2)
    assert torch.allclose(g0(x), ref, atol=1e-5, rtol=1e-5)


def test_scripted_gdn_shares_parameters_and_trains():
    torch.manual_seed(1)
    g = mmcodec.GDN(64).to(dev())
    m = torch.jit.script(g)
    assert list(m.state_dict().keys()) == list(g.state_dict().keys())
    x = torch.rand(2, 64, 16, 16, device=dev()) * 2 - 1
    xe = x.clone().requires_grad_(True)
    xs = x.clone().requires_grad_(True)
    w = torch.rand(2, 64, 16, 16, device=dev())
    (g(xe) * w).sum().backward()
    ge = (xe.grad.clone(), g.beta.grad.clone(), g.gamma.grad.clone())
    g.beta.grad = g.gamma.grad = None
    (m(xs) * w).sum().backward()                      # the scripted module updates the SAME Parameter objects
    assert g.beta.grad is not None and g.gamma.grad is not None
    for a, b in zip(ge, (xs.grad, g.beta.grad, g.gamma.grad)):
        assert torch.allclose(a, b, rtol=1e-3, atol=1e-5 * float(a.abs().max()))
    # an update through the shared Parameter is seen by the scripted module
    with torch.no_grad():
        g.beta.mul_(1.5)
    assert torch.equal(g(x), m(x))


def test_scripted_lower_bound_and_parametrizer():
    lb = mmcodec.LowerBound(0.11).to(dev())
    s = torch.jit.script(lb)
    x = torch.tensor([-1.0, 0.05, 0.11, 0.2, 3.0], device=dev(), requires_grad=True)
    y = s(x)
    assert torch.equal(y, torch.max(x.detach(), torch.tensor(0.11, device=dev())))
    # pass-through gradient towards the bound (bound_ops.py:40-56): d/dx = 1 where x >= bound or grad < 0
    y.backward(torch.tensor([1.0, -1.0, 1.0, 1.0, -1.0], device=dev()))
    assert torch.equal(x.grad, torch.tensor([0.0, -1.0, 1.0, 1.0, -1.0], device=dev()))
    p = mmcodec.NonNegativeParametrizer(minimum=1e-6).to(dev())
    sp = torch.jit.script(p)
    v = torch.rand(16, device=dev())
    assert torch.allclose(sp(v), p(v), rtol=0, atol=1e-7)


def test_scripted_op_has_no_cpu_path():
    m = torch.jit.script(mmcodec.GDN(16))
    with pytest.raises((RuntimeError, NotImplementedError)):
        m(torch.rand(1, 16, 2, 2))
