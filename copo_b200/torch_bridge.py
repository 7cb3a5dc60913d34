"""torch.nn.Module + autograd face of the hand-written learner (Model / Policy interface).

RLlib's TorchPolicy drives a policy as `loss = policy.loss(model, dist_class, train_batch); loss.backward();
optimizer.step()` over `model.parameters()` (the loop the reference's policies inherit: algo_ippo.py:78-79,
algo_ccppo.py:314, algo_copo.py:207).  The kernels here compute the loss AND its gradient in one pass and keep all
parameters of a model in one flat buffer, so the bridge is thin:

  * `KernelModule(model)` is an `nn.Module` whose single parameter shares storage with `model.flat` (a torch optimiser
    stepping it updates the kernels' weights in place) and which answers `forward(input_dict, state, seq_lens)`,
    `value_function()`, `central_value_function(...)`, `state_dict()` (RLlib key names) like the reference's models;
  * `kernel_loss(policy, module, train_batch)` returns the total loss as a scalar with an autograd edge to that
    parameter: its backward hands out the gradient the kernels wrote (scaled by the incoming gradient), so
    `.backward()`, gradient clipping and any `torch.optim` optimiser work unchanged.

Nothing on the product path needs this module (the trainers call `learn_on_batch`); it exists for hosts that own the
optimisation loop.  tests/test_bridge_gpu.py checks gradients and an Adam step against `learn_on_batch`.
"""
import torch


class _KernelLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, flat, module, policy, train_batch, global_rows):
        model = module.model
        assert flat.data_ptr() == model.flat.data_ptr(), "the module's parameter must alias model.flat"
        module.sync()                                    # an optimiser may have stepped the shared storage in place
        model.zero_grad()
        total = policy.loss(model, None, train_batch, global_rows=global_rows)
        ctx.save_for_backward(model.grad.clone())        # model.grad is rewritten by the next loss call
        return total.detach().to(torch.float32).reshape(())

    @staticmethod
    def backward(ctx, grad_out):
        (grad,) = ctx.saved_tensors
        return grad * grad_out, None, None, None, None


class KernelModule(torch.nn.Module):
    def __init__(self, model):
        super().__init__()
        self.__dict__["model"] = model                   # a plain attribute: CCModel / CoPOModel are not Modules
        # (named apart from the model's own attributes: unknown names fall through to the model below; CoPO's
        # lcf_parameters stay the model's - meta_update steps them, not the loss, algo_copo.py:228)
        self.kernel_parameters = torch.nn.Parameter(model.flat, requires_grad=True)

    def sync(self):
        """Tells the model when its weights were written through the parameter (torch counts in-place writes per tensor):
        the cached tensor-core weight operands are then rebuilt at the next use."""
        v = self.kernel_parameters._version
        if v != self.__dict__.get("_seen_version"):
            self.model.mark_weights_changed()
            self.__dict__["_seen_version"] = v

    def forward(self, input_dict, state=None, seq_lens=None):
        self.sync()
        with torch.no_grad():
            return self.model.forward(input_dict, [] if state is None else state, seq_lens)

    def value_function(self):
        return self.model.value_function()

    def __getattr__(self, name):                         # central_value_function, get_nei_value, lcf_mean, tower_stats ...
        try:
            return super().__getattr__(name)
        except AttributeError:
            return getattr(self.__dict__["model"], name)

    def state_dict(self, *args, **kwargs):
        return self.model.state_dict()

    def load_state_dict(self, state_dict, strict=True):
        return self.model.load_state_dict(state_dict, strict)


def kernel_loss(policy, module, train_batch, global_rows=None):
    """Total loss of `policy` on `train_batch` with an autograd edge to `module.kernel_parameters`; statistics land in
    `policy.model.tower_stats` as with `policy.loss`."""
    assert module.model is policy.model
    return _KernelLoss.apply(module.kernel_parameters, module, policy, train_batch, global_rows)
