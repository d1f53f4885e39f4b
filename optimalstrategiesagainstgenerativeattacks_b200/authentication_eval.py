"""Authentication evaluation (GIM vs GIM / replay / random-source attackers) -- mirrors the reference's
authentication_eval/agents.py:16-62, authentication_score.py:21-125 and the GIM parts of eval_gim_on_authentication.py:25-106,
running the networks on the libgim_b200 kernels (inference only).

Kept: class / function names, argument names, return values, and the reference's quirk that the networks stay in train mode
(so the spectral-norm power iteration runs on every encoder call, exactly as often as in the reference: `act` encodes si and
test per call).  Changed: batches can come from a device-resident `ResidentGIMDataSet` (no DataLoader workers), the per-batch
outputs stay on the device until the end, and the ROC-AUC is computed here (rank statistic with average ranks for ties, equal to
sklearn.metrics.roc_auc_score) so that scikit-learn is not needed on the GPU box.
"""
import itertools
import os
import random

import numpy as np
import torch

from .gim_img_models import get_au, get_im


class Authenticator:
    """Reference agents.py:16-26."""

    def __init__(self, au_model_func, th=0.):
        self.au_model_func = au_model_func
        self.th = th

    def act(self, test_sample, si_sample):
        out = self.au_model_func(test_sample=test_sample, si_sample=si_sample)
        pred = torch.ge(out, self.th).to(torch.long)
        return out, pred


class Impersonator:
    """Reference agents.py:32-41."""

    def __init__(self, im_model_func):
        self.im_model_func = im_model_func

    def act(self, leaked_sample, n):
        return self.im_model_func(leaked_sample=leaked_sample, n=n)


def replay_impersonator(leaked_sample, n):
    """Reference agents.py:47-52: n draws (python `random`) among the m leaked images, the same draw for the whole batch."""
    m = leaked_sample.size(1)
    picks = [random.randrange(m) for _ in range(n)]
    return leaked_sample.index_select(1, torch.tensor(picks, device=leaked_sample.device))


def rand_source_impersonator(leaked_sample, n, gim_ds):
    """Reference agents.py:55-65: the real sample of a random other episode."""
    batch_size = leaked_sample.size(0)
    fake_sample = torch.stack([gim_ds[random.randint(0, len(gim_ds) - 1)]["real_sample"] for _ in range(batch_size)], dim=0)
    assert fake_sample.size(1) == n
    return fake_sample.to(leaked_sample.device)


def get_au_function(au):
    """Reference eval_gim_on_authentication.py:25-43 (encode order si-src, si-env, test-src, test-env; no mode change)."""
    def au_model_func(test_sample, si_sample):
        with torch.no_grad():
            au_si_src = au.src_encode_sample(si_sample)
            au_si_env = au.env_encode_sample(si_sample)
            au_test_src = au.src_encode_sample(test_sample)
            au_test_env = au.env_encode_sample(test_sample)
            out = au.dis(test_src=au_test_src, test_env=au_test_env, si_src=au_si_src, si_env=au_si_env)
        return out.detach()
    return au_model_func


def get_im_function(im, args_dict):
    """Reference eval_gim_on_authentication.py:76-81."""
    def im_model_func(leaked_sample, n):
        with torch.no_grad():
            fake_sample = im.forward(leaked_sample=leaked_sample, n=n, remove_noise_mean=args_dict['remove_noise_mean'])
        return fake_sample.detach()
    return im_model_func


def get_gim_authenticator(device, ckpt_path, args_dict):
    """Reference eval_gim_on_authentication.py:84-93."""
    au = get_au(img_size=args_dict['img_size'], img_channels=args_dict['img_channels'], style_dim=args_dict['style_dim'])
    au.load_state_dict(torch.load(ckpt_path, map_location='cpu')['authenticator'])
    return Authenticator(get_au_function(au.to(device)))


def get_gim_impersonator(device, ckpt_path, args_dict):
    """Reference eval_gim_on_authentication.py:96-107."""
    im = get_im(img_size=args_dict['img_size'], img_channels=args_dict['img_channels'], style_dim=args_dict['style_dim'],
                use_img_att=args_dict['use_img_att'], num_env_noise_layers=args_dict['num_env_noise_layers'])
    im.load_state_dict(torch.load(ckpt_path, map_location='cpu')['impersonator'])
    return Impersonator(get_im_function(im.to(device), args_dict))


def write_results(file_path, acc, acc_on_fake, acc_on_real, print_to_stdout=False):
    """Reference authentication_score.py:21-30."""
    s = "accuracy: {}\naccuracy on fake: {}\naccuracy on real: {}\n".format(acc, acc_on_fake, acc_on_real)
    file_dir = os.path.dirname(file_path)
    if file_dir and not os.path.isdir(file_dir):
        os.makedirs(file_dir)
    with open(file_path, 'w') as f:
        f.write(s)
    if print_to_stdout:
        print(s)


def comp_acc(pred_on_real, pred_on_fake):
    """Reference authentication_score.py:33-44."""
    assert len(pred_on_real.size()) == 1 and len(pred_on_fake.size()) == 1
    assert pred_on_real.size(0) == pred_on_fake.size(0)
    acc_on_real = pred_on_real.to(torch.float).mean()
    acc_on_fake = torch.eq(pred_on_fake, 0).to(torch.float).mean()
    acc = 0.5 * (acc_on_real + acc_on_fake)
    return acc, acc_on_fake, acc_on_real


def roc_auc_score(y_true, y_score):
    """Area under the ROC curve = P(score_pos > score_neg) + 0.5 P(equal): Mann-Whitney U with average ranks for ties (what
    sklearn.metrics.roc_auc_score returns for binary labels, authentication_score.py:95)."""
    y_true = np.asarray(y_true).astype(bool).reshape(-1)
    y_score = np.asarray(y_score, dtype=np.float64).reshape(-1)
    n_pos, n_neg = int(y_true.sum()), int((~y_true).sum())
    if n_pos == 0 or n_neg == 0:
        raise ValueError("Only one class present in y_true. ROC AUC score is not defined in that case.")
    order = np.argsort(y_score, kind="mergesort")
    s = y_score[order]
    ranks = np.empty(len(s), dtype=np.float64)
    i = 0
    while i < len(s):
        j = i
        while j + 1 < len(s) and s[j + 1] == s[i]:
            j += 1
        ranks[order[i:j + 1]] = 0.5 * (i + j) + 1.0
        i = j + 1
    return float((ranks[y_true].sum() - n_pos * (n_pos + 1) / 2.0) / (n_pos * n_neg))


def _batches(ds, batch_size, num_workers):
    if hasattr(ds, "iter_batches"):                       # device-resident episode source
        n_batches = (len(ds) + batch_size - 1) // batch_size
        return ds.iter_batches(batch_size, shuffle=True), n_batches
    from torch.utils.data import DataLoader
    loader = DataLoader(ds, batch_size=batch_size, shuffle=True, num_workers=num_workers)
    return loader, len(loader)


def eval_authenticator_and_impersonator(device, ds, batch_size, num_workers, authenticator, impersonator, dbg=False):
    """Reference authentication_score.py:47-97: -> (acc, acc_on_fake, acc_on_real, auc)."""
    pred_on_fake_list, pred_on_real_list, out_on_fake_list, out_on_real_list = [], [], [], []
    batches, n_batches = _batches(ds, batch_size, num_workers)
    num_iters = min(1000, n_batches) if dbg else n_batches
    for data_batch in itertools.islice(batches, num_iters):
        real_sample = data_batch["real_sample"].to(device)
        leaked_sample = data_batch["leaked_sample"].to(device)
        si_sample = data_batch["si_sample"].to(device)
        n = real_sample.size(1)
        out_on_real, pred_on_real = authenticator.act(test_sample=real_sample, si_sample=si_sample)
        fake_sample = impersonator.act(leaked_sample=leaked_sample, n=n)
        out_on_fake, pred_on_fake = authenticator.act(test_sample=fake_sample, si_sample=si_sample)
        out_on_real_list.append(out_on_real.view(-1).detach())
        out_on_fake_list.append(out_on_fake.view(-1).detach())
        pred_on_real_list.append(pred_on_real.view(-1).detach())
        pred_on_fake_list.append(pred_on_fake.view(-1).detach())
    out_on_real, out_on_fake = torch.cat(out_on_real_list), torch.cat(out_on_fake_list)
    pred_on_real, pred_on_fake = torch.cat(pred_on_real_list), torch.cat(pred_on_fake_list)
    acc, acc_on_fake, acc_on_real = comp_acc(pred_on_real=pred_on_real, pred_on_fake=pred_on_fake)
    y_true = torch.cat([torch.ones_like(out_on_real), torch.zeros_like(out_on_fake)]).cpu().numpy()
    y_score = torch.cat([out_on_real, out_on_fake]).cpu().numpy()
    auc = roc_auc_score(y_true=y_true, y_score=y_score)
    return acc, acc_on_fake, acc_on_real, auc


def eval_dis_on_multiple_im(device, ds, batch_size, num_workers, authenticator, impersonator_dict):
    """Reference authentication_score.py:100-125."""
    results = {}
    for im_key in impersonator_dict.keys():
        print("\nEvaluating on impersonator: {}\n".format(im_key))
        acc, acc_on_fake, acc_on_real, auc = eval_authenticator_and_impersonator(
            device=device, ds=ds, batch_size=batch_size, num_workers=num_workers, authenticator=authenticator,
            impersonator=impersonator_dict[im_key])
        results[im_key] = {"acc": acc, "acc_on_fake": acc_on_fake, "acc_on_real": acc_on_real, "auc": auc}
    return results
