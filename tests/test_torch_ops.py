import numpy as np
import torch

from pytensor_federated_b200.models import LinregShards, make_demo_data
from pytensor_federated_b200.parallel import FederatedEngine
from pytensor_federated_b200.torch_ops import FederatedLogp, arrays_to_arrays, federated_logp, logp_only


def test_autograd_twin_uses_federated_gradients_once():
    x, y, sigma = make_demo_data()
    eng = FederatedEngine(LinregShards([x], [y], [sigma]), backend="collective")
    a = torch.tensor(0.4, dtype=torch.float64, requires_grad=True)
    b = torch.tensor(1.2, dtype=torch.float64, requires_grad=True)
    before = eng.n_evals
    loss = 2.0 * federated_logp(eng.logp_grad, a, b) + a * a
    loss.backward()
    assert eng.n_evals == before + 1
    logp, (da, db) = eng.logp_grad(np.array(0.4), np.array(1.2))
    np.testing.assert_allclose(loss.item(), 2 * logp + 0.16)
    np.testing.assert_allclose(a.grad.item(), 2 * da + 0.8)
    np.testing.assert_allclose(b.grad.item(), 2 * db)
    # torch.optim can drive the federated model: MAP by Adam moves towards the MLE
    mod = FederatedLogp(eng.logp_grad)
    p = [torch.zeros((), dtype=torch.float64, requires_grad=True) for _ in range(2)]
    opt = torch.optim.Adam(p, lr=0.1)
    first = None
    for _ in range(60):
        opt.zero_grad()
        nll = -mod(*p)
        first = first if first is not None else nll.item()
        nll.backward()
        opt.step()
    assert nll.item() < first
    outs = arrays_to_arrays(lambda u, v: (u + v, u * v), torch.tensor([1.0, 2.0]), torch.tensor([3.0, 4.0]))
    assert torch.equal(outs[0], torch.tensor([4.0, 6.0]))
    assert logp_only(lambda u: np.asarray(-float(u) ** 2), torch.tensor(3.0)).item() == -9.0
