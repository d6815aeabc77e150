"""``spikingjelly.clock_driven.functional`` replacement (reference train.py:221, test.py)."""


def reset_net(net):
    """Calls ``reset()`` on every sub-module that has one (membrane potentials back to ``v_reset``)."""
    for m in net.modules():
        if hasattr(m, 'reset'):
            m.reset()
