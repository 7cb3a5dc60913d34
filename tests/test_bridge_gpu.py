"""copo_b200.torch_bridge: the nn.Module / autograd face of the learner behaves like `learn_on_batch` when a host drives
it the way RLlib's TorchPolicy does (loss -> backward -> torch optimiser)."""
import pytest
import torch

from test_learner_gpu import _copo_batch, close

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("algo", ["copo", "ippo"])
def test_backward_and_adam_step_equal_learn_on_batch(algo):
    from copo_b200 import policy as P
    from copo_b200.torch_bridge import KernelModule, kernel_loss
    D, B = 92, 1024
    cls, cfg = (P.CoPOPolicy, P.copo_config) if algo == "copo" else (P.IPPOPolicy, P.ippo_config)
    host, own = cls(D, 2, cfg(seed=4)), cls(D, 2, cfg(seed=4))
    module = KernelModule(host.model)
    params = list(module.parameters())
    assert len(params) == 1 and params[0].data_ptr() == host.model.flat.data_ptr()
    assert set(module.state_dict()) == set(host.model.state_dict())
    lr = 5.0                                                     # large steps: stale weight operands would show
    opt = torch.optim.SGD(module.parameters(), lr=lr)
    for step in range(3):
        batch = {k: v.cuda() for k, v in _copo_batch(B, D, 30 + step).items()}
        before = host.model.flat.clone()
        opt.zero_grad()
        loss = kernel_loss(host, module, batch)
        assert loss.requires_grad and loss.shape == ()
        (2.0 * loss).backward()                                   # the incoming gradient scales the kernels' gradient
        own.model.zero_grad()
        want = own.loss(own.model, None, batch)                   # the twin holds the same weights: direct call
        scale = float(own.model.grad.abs().max())
        close(loss.detach(), want, 1e-5, 1e-6)
        close(module.kernel_parameters.grad, 2.0 * own.model.grad, 2e-5, 2e-6 * scale)
        module.kernel_parameters.grad.mul_(0.5)
        opt.step()                                                # steps the kernels' flat weight buffer in place
        close(host.model.flat, before - lr * own.model.grad, 1e-6, lr * 4e-6 * scale)
        assert float((host.model.flat - before).abs().max()) > 0
        own.model.flat.copy_(host.model.flat)
        own.model.mark_weights_changed()
        # inference through the module sees the stepped weights (the next loss call re-derives the weight operands)
        logits = module({"obs": batch["obs"]}, [], None)[0]
        assert float((logits - own.model.forward(batch["obs"])).abs().max()) < 1e-4
        assert float((logits - batch["action_dist_inputs"]).abs().max()) > 1e-2
    with pytest.raises(ValueError):
        module.value_function()
