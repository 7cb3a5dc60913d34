"""Host mechanics of copo_b200.torch_bridge with a stand-in model (no GPU): the module's one parameter aliases the flat
weight buffer, the loss carries an autograd edge whose backward is the buffer the "kernels" filled, a torch optimiser
steps the shared storage, unknown attributes fall through to the model."""
import importlib.util
import os

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bridge():
    spec = importlib.util.spec_from_file_location("b2c_torch_bridge", os.path.join(ROOT, "copo_b200", "torch_bridge.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class _Model:
    def __init__(self):
        self.flat, self.grad = torch.arange(5.0), torch.zeros(5)
        self.lcf_parameters, self.tower_stats, self.versions = torch.zeros(2), {}, 0

    def forward(self, d, state, seq_lens):
        return d["obs"] * self.flat.sum(), state

    def value_function(self):
        raise ValueError("Centralized Value Function should not be called directly!")

    def state_dict(self):
        return {"_logits._model.0.weight": self.flat.clone()}

    def mark_weights_changed(self):
        self.versions += 1

    def zero_grad(self):
        self.grad.zero_()


class _Policy:
    def __init__(self, model):
        self.model = model

    def loss(self, model, dist_class, batch, global_rows=None):
        model.grad.add_(batch["g"])                              # what the kernels do: loss and gradient in one pass
        return (batch["g"] * model.flat).sum().to(torch.float64)


def test_module_parameter_aliases_the_flat_buffer_and_backward_hands_out_the_kernel_gradient():
    tb = _bridge()
    model = _Model()
    module, policy = tb.KernelModule(model), _Policy(model)
    params = list(module.parameters())
    assert len(params) == 1 and params[0].data_ptr() == model.flat.data_ptr()
    assert list(module.state_dict()) == ["_logits._model.0.weight"]
    assert module.lcf_parameters is model.lcf_parameters and module.tower_stats is model.tower_stats
    opt = torch.optim.SGD(module.parameters(), lr=0.5)
    for step in range(2):
        g = torch.tensor([1.0, -2.0, 0.0, 3.0, 0.5]) * (step + 1)
        before = model.flat.clone()
        opt.zero_grad()
        loss = tb.kernel_loss(policy, module, {"g": g})
        assert loss.requires_grad and loss.dtype == torch.float32 and float(loss.detach()) == float((g * before).sum())
        (3.0 * loss).backward()
        assert torch.equal(module.kernel_parameters.grad, 3.0 * g)
        opt.step()
        assert torch.equal(model.flat, before - 1.5 * g)         # the optimiser stepped the kernels' weights in place
        seen = model.versions
        tb.kernel_loss(policy, module, {"g": 0 * g})
        assert model.versions == seen + 1                        # weights written through the parameter: operands invalidated
        tb.kernel_loss(policy, module, {"g": 0 * g})
        assert model.versions == seen + 1                        # ... once, not again while nobody writes them
    seen = model.versions
    with torch.no_grad():
        module.kernel_parameters.mul_(2.0)
    out, state = module({"obs": torch.ones(2)}, [], None)
    assert model.versions == seen + 1                            # the forward pass notices in-place writes too
    assert torch.equal(out, torch.ones(2) * model.flat.sum()) and state == [] and not out.requires_grad
    with pytest.raises(ValueError):
        module.value_function()
