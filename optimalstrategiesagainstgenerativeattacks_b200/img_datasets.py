"""Device-resident episode source for the GIM image path -- the B200 replacement of the reference's DataLoader + PIL pipeline
(data_handling/img_datasets.py:68-103 ImgGIMDataSet.__getitem__, :153-187 OmniglotGIMDataSet.__getitem__).

At > 1 k episodes/s per GPU a host pipeline (15 image decodes per episode) is the bottleneck, so the decoded images of every
class live in HBM once (Omniglot is already RAM-resident in the reference, `_load_data` :189-211) and an episode is an index
gather on the device.  The *index algebra* is the reference's, integer for integer:

    cls    = index // example_cnt_per_class
    idx    = rng.sample(range(n_imgs_in_class), m + n + k)        # python `random`, here an explicitly seeded random.Random
    leaked = idx[:m];  real = idx[m:m+n];  si = idx[m+n:]

so that, given the same generator state, the same images land in the same slots (tests/test_host_cpu.py checks this against
the oracle's statement of the reference).
"""
import random

import torch


def episode_indices(index, example_cnt_per_class, n_imgs_in_class, m, n, k, rng):
    """-> (class index, leaked image indices, real image indices, si image indices); reference img_datasets.py:79-91, 164-172."""
    cls = index // example_cnt_per_class
    idx = rng.sample(list(range(n_imgs_in_class)), m + n + k)
    return cls, idx[:m], idx[m:m + n], idx[m + n:]


class ResidentGIMDataSet:
    """Episodes over classes whose images are resident on `device`.

    `class_images`: list of tensors [n_imgs_c, C, S, S] (float32, already normalised like the reference's `load_image`), or one tensor
    [n_classes, n_imgs, C, S, S].  Classes with fewer than m+n+k images are dropped (reference :60-63).  `__getitem__` returns the
    reference's example dict; `batch(indices)` returns the same dict batched, built by ONE gather per sample kind."""

    def __init__(self, class_images, m, n, k, example_cnt_per_class=50, device=None, seed=None, class_names=None):
        if torch.is_tensor(class_images):
            class_images = list(class_images.unbind(0))
        keep = [i for i, t in enumerate(class_images) if t.shape[0] >= m + n + k]
        self.m, self.n, self.si = m, n, k
        self.example_cnt_per_class = example_cnt_per_class
        self.device = torch.device(device) if device is not None else class_images[0].device
        self.class_names = [str(class_names[i]) if class_names is not None else str(i) for i in keep]
        self.counts = [int(class_images[i].shape[0]) for i in keep]
        offs, total = [], 0
        for c in self.counts:
            offs.append(total)
            total += c
        self.offsets = offs
        self.images = torch.cat([class_images[i] for i in keep], dim=0).to(self.device).contiguous()      # [sum n_imgs, C, S, S]
        self.n_classes = len(keep)
        self.rng = random.Random(seed) if seed is not None else random       # the reference draws from the global `random` module

    def __len__(self):
        return self.n_classes * self.example_cnt_per_class

    def _draw(self, index):
        cls, leaked, real, si = episode_indices(index, self.example_cnt_per_class, self.counts[index // self.example_cnt_per_class],
                                                self.m, self.n, self.si, self.rng)
        off = self.offsets[cls]
        return cls, [off + i for i in leaked], [off + i for i in real], [off + i for i in si]

    def __getitem__(self, index):
        cls, leaked, real, si = self._draw(index)
        take = lambda ids: self.images.index_select(0, torch.tensor(ids, device=self.device))
        return {"real_sample": take(real), "leaked_sample": take(leaked), "si_sample": take(si), "class": cls, "class_name": self.class_names[cls]}

    def batch(self, indices):
        """Episodes `indices` (drawn in order, like a DataLoader's collate of __getitem__ calls) -> batched example dict on the device."""
        rows = [self._draw(int(i)) for i in indices]
        flat = lambda j: torch.tensor([r[j] for r in rows], device=self.device).reshape(-1)
        out = {}
        for name, j, s in (("real_sample", 2, self.n), ("leaked_sample", 1, self.m), ("si_sample", 3, self.si)):
            out[name] = self.images.index_select(0, flat(j)).reshape(len(rows), s, *self.images.shape[1:])
        out["class"] = torch.tensor([r[0] for r in rows])
        out["class_name"] = [self.class_names[r[0]] for r in rows]
        return out

    def iter_batches(self, batch_size, shuffle=True, drop_last=False, generator=None):
        """The DataLoader(ds, batch_size, shuffle) loop of the reference's training / eval drivers, without worker processes."""
        order = torch.randperm(len(self), generator=generator).tolist() if shuffle else list(range(len(self)))
        for i in range(0, len(order), batch_size):
            chunk = order[i:i + batch_size]
            if drop_last and len(chunk) < batch_size:
                return
            yield self.batch(chunk)


def synthetic_classes(n_classes, n_imgs, channels, size, seed=1234, device="cpu"):
    """U(-1, 1) images standing in for a decoded dataset (synthetic Omniglot- / VoxCeleb2-shaped episodes)."""
    g = torch.Generator().manual_seed(seed)
    return (torch.rand((n_classes, n_imgs, channels, size, size), generator=g) * 2 - 1).to(device)
