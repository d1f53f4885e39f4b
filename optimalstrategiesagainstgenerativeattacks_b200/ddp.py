"""Data parallelism over independent episodes: one process per GPU, episodes sharded evenly, ONE gradient all-reduce(sum)
per optimizer step over a flat fp32 bucket (NCCL over NVLink/NVSwitch on the GPU box, gloo in the CPU tests), the 1/world
scale folded into the fused Adam kernel.  Replaces the reference's nn.DataParallel (training/gim_img_training.py:406-411),
whose math it reproduces: the mean over the global batch.  Nothing else crosses GPUs: spectral-norm
u/v stay identical on all ranks because they depend only on the (replicated) weights and the call count.
"""
import torch
import torch.distributed as dist


def shard_range(global_batch, rank, world):
    """Rank r owns episodes [r*B/W, (r+1)*B/W); B % W == 0 as the reference enforces (gim_img_training.py:377)."""
    if global_batch % world != 0:
        raise ValueError("global batch %d is not divisible by world size %d" % (global_batch, world))
    per = global_batch // world
    return rank * per, (rank + 1) * per


class FlatGradBucket:
    """Re-homes the .grad of every parameter that receives gradients into one contiguous fp32 buffer (so the collective is a
    single call and gradient addresses are stable for the fused Adam pointer table / CUDA graphs)."""

    def __init__(self, params):
        self.params = [p for p in params if p.grad is not None]      # unused parameters (img_att, out_mlp) are not waited on
        total = sum(p.numel() for p in self.params)
        ref = self.params[0]
        self.flat = torch.zeros(total, dtype=torch.float32, device=ref.device)
        off = 0
        for p in self.params:
            view = self.flat[off:off + p.numel()].view_as(p)
            view.copy_(p.grad)
            p.grad = view
            off += p.numel()

    def all_reduce(self, group=None):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)

    def covers(self, params):
        """True if exactly the parameters re-homed here carry a gradient (a parameter whose first gradient appears later would
        otherwise be stepped with an un-reduced gradient and the replicas would drift apart silently)."""
        with_grad = [p for p in params if p.grad is not None]
        return len(with_grad) == len(self.params) and all(a is b for a, b in zip(with_grad, self.params))


def broadcast_module_state(module, src=0, group=None):
    """Every rank adopts rank `src`'s parameters and buffers (spectral-norm u/v included): replica consistency no longer depends on
    the caller seeding every process identically before building the networks."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) <= 1:
        return
    with torch.no_grad():
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t.data, src=src, group=group)


_comm = {"stream": None, "busy": False}


def _comm_stream(device):
    """The communication stream: HIGH priority, so that the collective's few CTAs are placed as soon as any SM frees up -- with the
    weight-gradient branches filling every tail of the persistent convolution kernels there are no idle SMs left to wait for."""
    if _comm["stream"] is None or _comm["stream"].device != device:
        _comm["stream"] = torch.cuda.Stream(device=device, priority=-1)
    return _comm["stream"]


def _order_after_deferred(device):
    """Two collectives on one communicator must execute in the same order on every rank.  A deferred all-reduce lives on the
    communication stream; a later all-reduce issued from the compute stream has no stream dependency on it, so a rank whose deferred
    collective has not been scheduled yet could start the later one first -- and the ranks would wait for each other forever.  The
    compute stream therefore waits for the communication stream before it enqueues its own collective."""
    if _comm["busy"] and _comm["stream"] is not None:
        torch.cuda.current_stream(device).wait_stream(_comm["stream"])
        _comm["busy"] = False


def attach(optimizer, group=None, defer=False):
    """Make `optimizer.step()` all-reduce its gradients first and apply the mean (grad_scale = 1/world).

    defer=True -- overlap: `step()` only LAUNCHES the all-reduce, on a communication stream, and returns; the Adam update runs at
    `optimizer.flush()` (also triggered by the next zero_grad / step / state_dict), after that stream has finished.  The training
    loops attach the attacker's optimizer this way and flush it at the end of the iteration: the authenticator step that follows the
    attacker step never reads the attacker's parameters (it consumes the already generated `fake_sample`), so the 246 MB all-reduce
    of G's gradients travels over NVLink underneath D's forward and backward and the result is bit-identical to stepping immediately.
    Inside a captured CUDA graph the communication stream becomes a parallel branch of the graph."""
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    in_kernel = getattr(type(optimizer), "FOLDS_GRAD_SCALE", False)      # FusedAdam multiplies by 1/world inside its update kernel
    if in_kernel:
        optimizer.grad_scale = 1.0 / world
    if world > 1:                                    # replicas start from rank 0's parameters
        with torch.no_grad():
            for g in optimizer.param_groups:
                for p in g["params"]:
                    dist.broadcast(p.data, src=0, group=group)
    inner_step, inner_zero, inner_sd = optimizer.step, optimizer.zero_grad, optimizer.state_dict
    state = {"bucket": None, "pending": False}

    def reduce_now():
        params = [p for g in optimizer.param_groups for p in g["params"]]
        if state["bucket"] is None:
            state["bucket"] = FlatGradBucket(params)
        elif not (torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()) and not state["bucket"].covers(params):
            raise RuntimeError("ddp: the set of parameters that receive gradients changed after the first optimizer step "
                               "(flat gradient bucket is frozen) -- re-attach the optimizer")
        state["bucket"].all_reduce(group)
        if not in_kernel and world > 1:
            state["bucket"].flat.mul_(1.0 / world)   # a stock optimizer sees the mean gradient

    def flush():
        """Finish a deferred step: wait for the communication stream, then run the (fused) Adam update."""
        if not state["pending"]:
            return None
        state["pending"] = False
        torch.cuda.current_stream().wait_stream(_comm_stream(state["bucket"].flat.device))
        _comm["busy"] = False
        return inner_step()

    def step(closure=None):
        if closure is not None:
            raise RuntimeError("ddp: closures are not supported")
        flush()
        if world <= 1:
            return inner_step()
        if not (defer and state["bucket"] is not None and state["bucket"].flat.is_cuda):
            if torch.cuda.is_available():
                _order_after_deferred(torch.cuda.current_device())
            reduce_now()                             # (the first step builds the bucket: gradients move into it)
            return inner_step()
        comm, cur = _comm_stream(state["bucket"].flat.device), torch.cuda.current_stream()
        comm.wait_stream(cur)                        # gradients are complete on the compute stream
        with torch.cuda.stream(comm):
            reduce_now()
        state["pending"] = True
        _comm["busy"] = True
        return None

    def zero_grad(*a, **k):
        flush()
        return inner_zero(*a, **k)

    def state_dict(*a, **k):
        flush()
        return inner_sd(*a, **k)

    optimizer.step, optimizer.flush, optimizer.zero_grad, optimizer.state_dict = step, flush, zero_grad, state_dict
    return optimizer
