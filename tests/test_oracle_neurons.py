"""Known-answer vectors of SURVEY.md section 8(c), hand-derived from the Row 6 equations
(NOT produced by running SpikingJelly) -- they pin oracle/sj_compat.py."""
import math

import pytest
import torch

from oracle import sj_compat as sj


def run(node, xs):
    hs, ss, vs = [], [], []
    for x in xs:
        x = torch.tensor([x])
        node.neuronal_charge(x)
        hs.append(float(node.v))
        node.neuronal_fire()
        ss.append(float(node.spike))
        node.neuronal_reset()
        vs.append(float(node.v))
    return hs, ss, vs


def test_if_vector():
    hs, ss, vs = run(sj.IFNode(detach_reset=True), [0.6, 0.6, 0.6])
    assert hs == pytest.approx([0.6, 1.2, 0.6]) and ss == [0, 1, 0] and vs == pytest.approx([0.6, 0, 0.6])


def test_if_threshold_is_inclusive():
    _, ss, _ = run(sj.IFNode(), [1.0])
    assert ss == [1]


def test_lif_vector():
    hs, ss, vs = run(sj.LIFNode(tau=2.0, detach_reset=True), [1.2] * 4)
    assert hs == pytest.approx([0.6, 0.9, 1.05, 0.6])
    assert ss == [0, 0, 1, 0]
    assert vs == pytest.approx([0.6, 0.9, 0.0, 0.6])


def test_plif_vector():
    n = sj.ParametricLIFNode(init_tau=3.0)
    assert float(n.w) == pytest.approx(-math.log(2.0))
    assert float(n.w.sigmoid()) == pytest.approx(1 / 3)
    hs, ss, _ = run(n, [3.3])
    assert hs == pytest.approx([1.1]) and ss == [1]
    n.reset()
    hs, ss, _ = run(n, [2.7])
    assert hs == pytest.approx([0.9]) and ss == [0]


def test_i_neuron_never_fires_and_integrates():
    n = sj.IFNode(v_threshold=float('inf'), v_reset=0.0, surrogate_function=sj.ATan())
    tot = 0.0
    for x in (0.5, 2.0, -1.0, 100.0):
        s = n(torch.tensor([x]))
        tot += x
        assert float(s) == 0.0 and float(n.v) == pytest.approx(tot)
    n.reset()
    assert n.v == 0.0


@pytest.mark.parametrize('cls,alpha_half', [(sj.ATan, 1 / (1 + (math.pi / 2) ** 2)),
                                            (sj.Sigmoid, 4 * torch.sigmoid(torch.tensor(2.0)).item()
                                             * (1 - torch.sigmoid(torch.tensor(2.0)).item()))])
def test_surrogate_gradients(cls, alpha_half):
    for u, want in ((0.0, 1.0), (0.5, alpha_half), (-0.5, alpha_half)):
        x = torch.tensor([u], requires_grad=True)
        y = cls()(x)
        assert float(y) == (1.0 if u >= 0 else 0.0)
        y.backward()
        assert float(x.grad) == pytest.approx(want, rel=1e-5)
    assert alpha_half == pytest.approx(0.288400 if cls is sj.ATan else 0.419974, abs=1e-6)


def test_bptt_lif_tau2_T2():
    """g_h(1)=b*d1, g_x(1)=g_h(1)/2, g_v(0)=g_h(1)/2, g_h(0)=a*d0+g_v(0)*(1-s0), g_x(0)=g_h(0)/2."""
    a, b = 0.7, -1.3
    for x0 in (0.8, 2.4):                       # no spike at t=0 / spike at t=0
        n = sj.LIFNode(tau=2.0, surrogate_function=sj.ATan(), detach_reset=True)
        x = torch.tensor([x0, 0.6], requires_grad=True)
        s0 = n(x[0:1])
        h0 = x0 / 2
        s1 = n(x[1:2])
        v0 = 0.0 if h0 >= 1 else h0
        h1 = v0 + (0.6 - v0) / 2
        (a * s0 + b * s1).sum().backward()
        d = lambda h: 1.0 / (1 + (math.pi * (h - 1.0)) ** 2)
        gh1 = b * d(h1)
        gv0 = gh1 / 2
        gh0 = a * d(h0) + gv0 * (1 - float(s0))
        assert float(x.grad[1]) == pytest.approx(gh1 / 2, rel=1e-5)
        assert float(x.grad[0]) == pytest.approx(gh0 / 2, rel=1e-5)


def test_bptt_plif_tau3_T2_with_decay_gradient():
    """Row 6, PLIF: r = sigmoid(w) = 1/3; g_h(1) = b d1, g_v(0) = g_h(1) (1 - r) (1 - s0), g_h(0) = a d0 + g_v(0), g_x(t) = g_h(t) r,
    g_w = r (1 - r) sum_t g_h(t) (x_t - v_{t-1}); Sigmoid(alpha = 4) surrogate: d(h) = 4 sig(4 (h - 1)) (1 - sig(4 (h - 1)))."""
    a, b = 0.7, -1.3
    sig = lambda z: 1.0 / (1.0 + math.exp(-z))
    d = lambda h: 4.0 * sig(4.0 * (h - 1.0)) * (1.0 - sig(4.0 * (h - 1.0)))
    for x0 in (0.9, 3.6):                       # no spike at t=0 (h0 = 0.3) / spike at t=0 (h0 = 1.2)
        n = sj.ParametricLIFNode(init_tau=3.0, surrogate_function=sj.Sigmoid(), detach_reset=True)
        x = torch.tensor([x0, 0.9], requires_grad=True)
        s0 = n(x[0:1])
        s1 = n(x[1:2])
        r = 1.0 / 3.0
        h0 = x0 * r
        v0 = 0.0 if h0 >= 1 else h0
        h1 = v0 + (0.9 - v0) * r
        assert float(s0) == (1.0 if h0 >= 1 else 0.0) and float(s1) == 0.0
        (a * s0 + b * s1).sum().backward()
        gh1 = b * d(h1)
        gv0 = gh1 * (1 - r) * (1 - float(s0))
        gh0 = a * d(h0) + gv0
        assert float(x.grad[1]) == pytest.approx(gh1 * r, rel=1e-5)
        assert float(x.grad[0]) == pytest.approx(gh0 * r, rel=1e-5)
        assert float(n.w.grad) == pytest.approx(r * (1 - r) * (gh1 * (0.9 - v0) + gh0 * x0), rel=1e-5)


def test_reset_net_and_shim():
    sj.install_shim()
    from spikingjelly.clock_driven import neuron, functional, surrogate  # noqa: F401
    m = torch.nn.Sequential(neuron.IFNode(), neuron.LIFNode())
    m[0](torch.ones(3))
    assert isinstance(m[0].v, torch.Tensor)
    functional.reset_net(m)
    assert m[0].v == 0.0
