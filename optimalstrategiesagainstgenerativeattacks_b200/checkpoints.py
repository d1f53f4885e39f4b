"""Checkpoint container for the GIM trainers.

On-disk schema = the reference's (training/checkpoints.py:9-44): ONE `torch.save`d dict
    {'global_step': int, 'last_epoch': int, <registered name>: <that object's state_dict()>, ...}
so files written by either implementation load in the other.  What differs is how the file gets there:

  * the state dicts are snapshotted to host memory first, so the training stream is only blocked for the device->host copies;
  * the file is written to `<name>.tmp` and renamed, so a crash never leaves a truncated checkpoint behind;
  * with `async_save=True` the serialisation + write run on a background thread (`wait()` joins it; `load` / the next `save`
    wait for it on their own) -- at > 2000 episodes/s a synchronous 1 GB `torch.save` would cost hundreds of iterations.
"""
import os
import threading

import torch


def _to_host(obj):
    """Deep copy of a state_dict-like structure with every tensor moved to host memory."""
    if torch.is_tensor(obj):
        return obj.detach().to("cpu", copy=True)
    if isinstance(obj, dict):
        out = type(obj)((k, _to_host(v)) for k, v in obj.items())
        if hasattr(obj, "_metadata"):                  # nn.Module.state_dict() versions (e.g. spectral_norm's weight.version)
            out._metadata = obj._metadata
        return out
    if isinstance(obj, (list, tuple)):
        return type(obj)(_to_host(v) for v in obj)
    return obj


class CheckpointIO:
    def __init__(self, checkpoint_dir, async_save=False, **named_objects):
        self.checkpoint_dir = checkpoint_dir
        self.module_dict = dict(named_objects)          # name -> anything with state_dict() / load_state_dict()
        self.async_save = async_save
        self._writer = None
        self._writer_error = None
        os.makedirs(checkpoint_dir, exist_ok=True)

    def register_modules(self, **named_objects):
        # NB the trainers register their GlobalStep under the name 'global_step': in the file that entry ({'global_step': n}) then
        # replaces the plain integer written first -- the reference's files look exactly like this, so it is kept
        self.module_dict.update(named_objects)

    # -- writing -------------------------------------------------------------------------------------------------
    def _write(self, payload, path):
        tmp = "%s.tmp.%d" % (path, os.getpid())        # unique per process: two writers never share a temporary file
        torch.save(payload, tmp)
        os.replace(tmp, path)

    def _write_bg(self, payload, path):
        try:
            self._write(payload, path)
        except BaseException as e:                     # surfaced by wait() on the training thread
            self._writer_error = e

    def wait(self):
        """Block until a background save (if any) has reached the disk; re-raises the writer's exception if it failed."""
        writer, self._writer = self._writer, None
        if writer is not None:
            writer.join()
        err, self._writer_error = getattr(self, "_writer_error", None), None
        if err is not None:
            raise RuntimeError("background checkpoint write failed: %r" % (err,)) from err

    def save(self, global_step, last_epoch, filename):
        self.wait()
        payload = {"global_step": global_step, "last_epoch": last_epoch}
        for name, obj in self.module_dict.items():
            payload[name] = _to_host(obj.state_dict())
        path = os.path.join(self.checkpoint_dir, filename)
        if self.async_save:
            self._writer = threading.Thread(target=self._write_bg, args=(payload, path), daemon=False)
            self._writer.start()
        else:
            self._write(payload, path)
        return path

    # -- reading -------------------------------------------------------------------------------------------------
    def load(self, filepath):
        """-> (global_step, last_epoch); (-1, -1) when the file does not exist (a fresh run), like the reference."""
        self.wait()
        if not os.path.exists(filepath):
            return -1, -1
        print('=> Loading checkpoint...')
        stored = torch.load(filepath, map_location='cpu')
        for name, obj in self.module_dict.items():
            if name not in stored:
                print('Warning: Could not find %s in checkpoint!' % name)
                continue
            obj.load_state_dict(stored[name])
        return stored["global_step"], stored["last_epoch"]
