"""Training utilities on the hot path -- mirrors the used part of the reference's training/utils.py
(GlobalStep :15-33, DataParallelMock :36-43, get_device :48-60, num_parameters :103-112, compute_grad2 :115-124,
adjust_batch_size :167-171, get_latest_ckpt :160-164)."""
import os

import torch
from torch import autograd

from . import ops


class GlobalStep(object):
    def __init__(self, gs=-1):
        self._gs = gs

    def step(self):
        self._gs += 1

    def get(self):
        return self._gs

    def set(self, gs):
        self._gs = gs

    def state_dict(self):
        return {"global_step": self._gs}

    def load_state_dict(self, d):
        self.set(d["global_step"])


class DataParallelMock:
    def __init__(self, module):
        self.module = module

    def forward(self, *inputs, **kwargs):
        return self.module.forward(*inputs, **kwargs)


def get_device(device_type, device_ids, verbose=True):
    """The hot path is CUDA only: asking for anything else is an error, not a fallback."""
    if device_type != 'cuda' or not torch.cuda.is_available():
        raise RuntimeError("the B200 GIM path needs a CUDA device (requested %r, cuda available: %s)" % (device_type, torch.cuda.is_available()))
    name = "cuda:{}".format(min(device_ids)) if device_ids else "cuda"
    if verbose:
        print('Using device {}'.format(name))
    device = torch.device(name)
    # one process drives ONE GPU: every kernel of libgim_b200 is launched on the current device's current stream, so the device the
    # tensors live on must be the current one (the reference's nn.DataParallel over device_ids is replaced by one process per GPU)
    torch.cuda.set_device(device)
    return device


def num_parameters(parameter_list):
    return float(sum(t.numel() for t in parameter_list))


def compute_grad2(out, x_in):
    """R1 penalty term: sum over the inputs of the squared input-gradient norm, per episode (reference :115-124).
    The gradient graph is built through the kernels' own differentiable backward operators (ops.py)."""
    with ops.input_grads_only():
        grad_out = autograd.grad(outputs=out.sum(), inputs=x_in, create_graph=True, retain_graph=True, only_inputs=True)
    reg = None
    for g in grad_out:
        r = ops.RowsSqSumFn.apply(g)
        reg = r if reg is None else reg + r
    return reg


def adjust_batch_size(batch_size, n_devices):
    if (n_devices > 0) and (batch_size % n_devices != 0):
        batch_size = (batch_size // n_devices) * n_devices
    return batch_size


def get_latest_ckpt(ckpt_dir):
    files = [f for f in os.listdir(ckpt_dir) if f.endswith('.pt')]
    return os.path.join(ckpt_dir, max(files)) if files else None
