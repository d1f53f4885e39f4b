"""torch.optim.Adam drop-in whose step() is ONE fused multi-tensor kernel launch (gim_adam_multi).

It subclasses torch.optim.Adam so that param_groups, state_dict()/load_state_dict() and LR schedulers behave exactly like
the reference's optimizers (gim_img_trainer.py:50-61, gim_gaussian_trainer.py:48-49): per-parameter state keys `step`,
`exp_avg`, `exp_avg_sq`; parameters whose grad is None are skipped and get no state.
"""
import ctypes

import torch

from . import _cabi as C


class _AdamTensor(ctypes.Structure):
    _fields_ = [("p", ctypes.c_void_p), ("g", ctypes.c_void_p), ("m", ctypes.c_void_p), ("v", ctypes.c_void_p),
                ("numel", ctypes.c_longlong), ("group", ctypes.c_int), ("pad", ctypes.c_int)]


class FusedAdam(torch.optim.Adam):
    FOLDS_GRAD_SCALE = True        # ddp.attach: the 1/world of the gradient mean is applied inside the update kernel (self.grad_scale)

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(params, lr=lr, betas=betas, eps=eps)
        self._table_key = None
        self._table_dev = None
        self._lrs_dev = None
        self._lrs_host = None
        self._step_dev = None
        self._n = 0
        self._max_numel = 0
        self._steps_done = None
        self.grad_scale = 1.0          # set to 1/world_size by the data-parallel wrapper (allreduce(sum))

    def _build(self, active, device):
        entries = (_AdamTensor * len(active))()
        max_numel = 0
        for i, (gi, p) in enumerate(active):
            st = self.state[p]
            if len(st) == 0:
                st["step"] = torch.tensor(0.0, dtype=torch.float32)
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            for t in (p, p.grad, st["exp_avg"], st["exp_avg_sq"]):
                if not (t.is_cuda and t.is_contiguous() and t.dtype == torch.float32):
                    raise RuntimeError("FusedAdam needs contiguous float32 CUDA parameters, gradients and state (no CPU fallback)")
            entries[i] = _AdamTensor(p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.numel(), gi, 0)
            max_numel = max(max_numel, p.numel())
        raw = torch.frombuffer(bytearray(bytes(entries)), dtype=torch.uint8)
        self._table_dev = raw.to(device)
        self._n = len(active)
        self._max_numel = max_numel
        if self._steps_done is None:          # first build, or just after load_state_dict: resume torch's count
            self._steps_done = max(int(float(self.state[p]["step"])) for _, p in active)
        self._step_dev = torch.full((1,), self._steps_done, dtype=torch.int64, device=device)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        active = [(gi, p) for gi, g in enumerate(self.param_groups) for p in g["params"] if p.grad is not None]
        if not active:
            return loss
        betas = {g["betas"] for g in self.param_groups}
        epss = {g["eps"] for g in self.param_groups}
        if len(betas) != 1 or len(epss) != 1:
            raise RuntimeError("FusedAdam: all groups must share betas/eps (they do on the GIM path)")
        device = active[0][1].device
        key = tuple((p.data_ptr(), p.grad.data_ptr()) for _, p in active)
        if key != self._table_key:
            self._build(active, device)
            self._table_key = key
        self._device = device
        if not torch.cuda.is_current_stream_capturing():
            self.sync_lrs()
        (b1, b2), eps = next(iter(betas)), next(iter(epss))
        C.call("gim_adam_multi", self._table_dev.data_ptr(), self._n, self._max_numel, self._lrs_dev.data_ptr(),
               self._step_dev.data_ptr(), b1, b2, eps, float(self.grad_scale))
        if not torch.cuda.is_current_stream_capturing():
            self._steps_done += 1
        return loss

    def sync_lrs(self):
        """Upload the per-group learning rates if a scheduler changed them (in place: the device array is referenced by
        captured CUDA graphs)."""
        lrs = [float(g["lr"]) for g in self.param_groups]
        if lrs != self._lrs_host and getattr(self, "_device", None) is not None:
            if self._lrs_dev is None:
                self._lrs_dev = torch.tensor(lrs, dtype=torch.float32, device=self._device)
            else:
                self._lrs_dev.copy_(torch.tensor(lrs, dtype=torch.float32))
            self._lrs_host = lrs

    def note_graph_steps(self, n):
        """Account for optimizer steps executed by CUDA-graph replays (the host-side counter is only used for state_dict)."""
        if self._steps_done is not None:
            self._steps_done += n

    def state_dict(self):
        # publish the device-side step count in torch's per-parameter `step` entries
        if self._table_key is not None:
            for st in self.state.values():
                if "step" in st:
                    st["step"] = torch.tensor(float(self._steps_done), dtype=torch.float32)
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._table_key = None      # rebuild the pointer table (and the device step) from the loaded state
        self._steps_done = None

    def zero_grad(self, set_to_none=False):
        """Gradients are zeroed in place by default so their addresses (the fused kernel's pointer table, CUDA graphs,
        the data-parallel flat buckets) stay valid; parameters that never received a gradient keep grad None."""
        if not set_to_none and self._table_key is not None:
            active = [p for g in self.param_groups for p in g["params"] if p.grad is not None]
            if tuple((p.data_ptr(), p.grad.data_ptr()) for p in active) == self._table_key:
                C.call("gim_zero_grads_multi", self._table_dev.data_ptr(), self._n, self._max_numel)      # one launch for all gradients
                return
        super().zero_grad(set_to_none=set_to_none)
