import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def make_net(name="giga", sd=None, device="cuda:0", frozen=True):
    """Build a giga_b200 network with the oracle's seeded parameters (subset by key).  frozen (default): requires_grad off, i.e.
    every call is an inference call whatever the gradient mode (with trainable parameters and gradient mode on, forward() is the
    differentiable native training step, like any nn.Module)."""
    import giga_b200
    from oracle import giga_oracle as O

    sd = sd if sd is not None else O.seeded_state_dict(seed=1)
    net = giga_b200.get_network(name)
    net.load_state_dict({k: v for k, v in sd.items() if k in net.state_dict()})
    net = net.to(device)
    return net.requires_grad_(False) if frozen else net
