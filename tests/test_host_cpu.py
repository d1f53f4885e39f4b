"""CPU: host-side logic of the drop-in path that needs no kernel -- the episode index algebra and the device-resident episode source
(reference data_handling/img_datasets.py:68-103, 153-187), the authentication-score arithmetic (authentication_eval/
authentication_score.py:33-97) and the attackers' RNG consumption (agents.py:47-65)."""
import random

import numpy as np
import torch

from oracle import gim_oracle as O
from optimalstrategiesagainstgenerativeattacks_b200 import authentication_eval as AE
from optimalstrategiesagainstgenerativeattacks_b200 import img_datasets as D


def test_episode_indices_match_oracle_bit_exact():
    for seed in (0, 1, 7, 1234):
        r1, r2 = random.Random(seed), random.Random(seed)
        for index in (0, 3, 49, 50, 999, 12345):
            for (per_cls, n_imgs, m, n, k) in ((50, 20, 5, 5, 5), (50, 20, 1, 5, 5), (7, 100, 5, 5, 10), (1, 15, 5, 5, 5)):
                assert D.episode_indices(index, per_cls, n_imgs, m, n, k, r1) == O.episode_indices(index, per_cls, n_imgs, m, n, k, r2)


def test_resident_dataset_gathers_the_drawn_images():
    imgs = [torch.arange(c * 1 * 2 * 2, dtype=torch.float32).reshape(c, 1, 2, 2) + 1000 * i for i, c in enumerate((20, 14, 30, 16))]
    ds = D.ResidentGIMDataSet(imgs, m=5, n=5, k=5, example_cnt_per_class=3, device="cpu", seed=11)
    assert ds.n_classes == 3 and len(ds) == 9                  # the 14-image class cannot fill m+n+k = 15 slots (reference :60-63)
    ref_rng = random.Random(11)
    kept = [imgs[0], imgs[2], imgs[3]]
    batch = ds.batch([4, 0, 8])
    for row, index in enumerate((4, 0, 8)):
        cls, leaked, real, si = O.episode_indices(index, 3, kept[index // 3].shape[0], 5, 5, 5, ref_rng)
        assert int(batch["class"][row]) == cls
        assert torch.equal(batch["leaked_sample"][row], kept[cls][leaked])
        assert torch.equal(batch["real_sample"][row], kept[cls][real])
        assert torch.equal(batch["si_sample"][row], kept[cls][si])
    ex = ds[5]
    assert ex["real_sample"].shape == (5, 1, 2, 2) and ex["class"] == 1 and ex["class_name"] == "2"
    n_seen = sum(b["real_sample"].shape[0] for b in ds.iter_batches(4, shuffle=True, generator=torch.Generator().manual_seed(0)))
    assert n_seen == len(ds)
    # global-`random` mode: the same stream the reference's datasets consume
    ds2 = D.ResidentGIMDataSet(imgs, m=5, n=5, k=5, example_cnt_per_class=3, device="cpu")
    random.seed(5)
    a = ds2[2]["si_sample"]
    random.seed(5)
    _, _, _, si = O.episode_indices(2, 3, 20, 5, 5, 5, random)
    assert torch.equal(a, imgs[0][si])


def test_roc_auc_matches_sklearn_with_ties():
    from sklearn.metrics import roc_auc_score as sk
    g = np.random.default_rng(3)
    for n in (10, 257, 4000):
        y = g.integers(0, 2, n)
        y[0], y[1] = 0, 1
        s = np.round(g.normal(size=n) + 0.7 * y, 1)            # rounding creates many ties
        assert abs(AE.roc_auc_score(y, s) - sk(y, s)) < 1e-12
    assert AE.roc_auc_score([1, 1, 0, 0], [2.0, 3.0, 0.0, 1.0]) == 1.0 and AE.roc_auc_score([1, 0], [0.5, 0.5]) == 0.5


def test_comp_acc_and_agents():
    acc, acc_fake, acc_real = AE.comp_acc(torch.tensor([1, 1, 0, 1]), torch.tensor([0, 1, 0, 0]))
    assert float(acc_real) == 0.75 and float(acc_fake) == 0.75 and float(acc) == 0.75
    out, pred = AE.Authenticator(lambda test_sample, si_sample: test_sample.sum(1) - si_sample.sum(1), th=0.).act(
        torch.tensor([[1., 2.], [0., 0.]]), torch.tensor([[1., 1.], [1., 0.]]))
    assert pred.tolist() == [1, 0] and pred.dtype == torch.long
    leaked = torch.arange(2 * 5 * 3, dtype=torch.float32).reshape(2, 5, 3)
    random.seed(9)
    got = AE.replay_impersonator(leaked, 4)
    random.seed(9)
    want = torch.cat([leaked[:, random.randrange(5)].unsqueeze(1) for _ in range(4)], dim=1)     # the reference expression (agents.py:50)
    assert torch.equal(got, want)
    ds = D.ResidentGIMDataSet(D.synthetic_classes(4, 16, 1, 4, seed=1), m=5, n=5, k=5, example_cnt_per_class=2, device="cpu", seed=2)
    fake = AE.rand_source_impersonator(leaked, 5, ds)
    assert fake.shape == (2, 5, 1, 4, 4)


def test_checkpoint_io_schema_async_and_atomic(tmp_path):
    """CheckpointIO keeps the reference's file schema (training/checkpoints.py:20-44), writes atomically and can save in the background."""
    from optimalstrategiesagainstgenerativeattacks_b200.checkpoints import CheckpointIO
    from optimalstrategiesagainstgenerativeattacks_b200.utils import GlobalStep
    net = torch.nn.Linear(3, 2)
    gs = GlobalStep()
    gs.set(41)
    for async_save in (False, True):
        io = CheckpointIO(str(tmp_path / ("a%d" % async_save)), async_save=async_save, net=net)
        io.register_modules(global_step=gs)
        path = io.save(global_step=41, last_epoch=3, filename="model_00000041.pt")
        io.wait()
        stored = torch.load(path, map_location="cpu")
        assert set(stored) == {"global_step", "last_epoch", "net"} and stored["last_epoch"] == 3
        assert stored["global_step"] == {"global_step": 41}            # the registered GlobalStep replaces the integer, as in the reference
        assert not [f for f in __import__("os").listdir(io.checkpoint_dir) if f.endswith(".tmp")]
        net2, gs2 = torch.nn.Linear(3, 2), GlobalStep()
        io2 = CheckpointIO(io.checkpoint_dir, net=net2, global_step=gs2)
        step, epoch = io2.load(path)
        assert epoch == 3 and gs2.get() == 41 and torch.equal(net2.weight, net.weight)
        assert io2.load(str(tmp_path / "missing.pt")) == (-1, -1)


def test_cat_params_direct_accumulation_and_merged_rows():
    """Host-side autograd plumbing of the merged projections (no kernel involved): CatParamsFn hands out row blocks of the gradient, or --
    inside deferred_weight_grads() with allocated .grad -- adds them straight into the leaves; _MergedRowsFn is a view forward and a split
    backward."""
    import torch
    from optimalstrategiesagainstgenerativeattacks_b200 import ops
    torch.manual_seed(0)
    ws = [torch.randn(r, 4, requires_grad=True) for r in (2, 3, 5)]
    pad = torch.zeros(6, 4)
    probe = torch.randn(16, 4)
    ref = torch.autograd.grad((torch.cat(ws + [pad], 0) * probe).sum(), ws)
    got = torch.autograd.grad((ops.cat_params(ws + [pad]) * probe).sum(), ws)
    for a, b in zip(got, ref):
        assert torch.equal(a, b)
    for w in ws:
        w.grad = torch.ones_like(w)
    with ops.deferred_weight_grads():
        (ops.cat_params(ws + [pad]) * probe).sum().backward()
    for w, r in zip(ws, ref):
        assert torch.allclose(w.grad, r + 1.0)
    # without the context the ordinary AccumulateGrad route gives the same result
    for w in ws:
        w.grad = torch.ones_like(w)
    (ops.cat_params(ws + [pad]) * probe).sum().backward()
    for w, r in zip(ws, ref):
        assert torch.allclose(w.grad, r + 1.0)

    buf = torch.arange(40, dtype=torch.float32)
    parts = [buf[0:8].view(1, 2, 4).clone().requires_grad_(), buf[8:20].view(1, 3, 4).clone().requires_grad_()]
    merged = ops._MergedRowsFn.apply((buf, 0), *parts)
    assert merged.shape == (1, 5, 4) and merged.data_ptr() == buf.data_ptr()
    g = torch.randn(1, 5, 4)
    ga, gb = torch.autograd.grad((merged * g).sum(), parts)
    assert torch.equal(ga, g[:, :2]) and torch.equal(gb, g[:, 2:])


def test_checkpoint_keeps_spectral_norm_metadata_and_cross_loads(tmp_path):
    """State dicts keep nn.Module's `_metadata` through the host snapshot, SNConv2d publishes spectral_norm's `weight.version`, and
    its state dict loads (strict) into torch.nn.utils.spectral_norm(nn.Conv2d) -- the module the reference builds
    (models/model_blocks.py:492-495) -- with the same effective weight state."""
    from optimalstrategiesagainstgenerativeattacks_b200 import model_blocks as mb
    from optimalstrategiesagainstgenerativeattacks_b200.checkpoints import CheckpointIO, _to_host
    torch.manual_seed(3)
    blk = torch.nn.Sequential(mb.SNConv2d(4, 6, 3, padding=1))
    sd = blk.state_dict()
    assert sd._metadata["0"]["spectral_norm"] == {"weight.version": 1}
    host = _to_host(sd)
    assert host._metadata["0"]["spectral_norm"] == {"weight.version": 1}
    io = CheckpointIO(str(tmp_path), net=blk)
    path = io.save(global_step=0, last_epoch=0, filename="m.pt")
    stored = torch.load(path, map_location="cpu", weights_only=False)
    ref = torch.nn.Sequential(torch.nn.utils.spectral_norm(torch.nn.Conv2d(4, 6, 3, padding=1)))
    ref.load_state_dict(stored["net"], strict=True)
    for k in ("0.weight_orig", "0.weight_u", "0.weight_v", "0.bias"):
        assert torch.equal(ref.state_dict()[k], sd[k]), k
    # and the other way round: a state dict written by the reference's module loads here
    blk2 = torch.nn.Sequential(mb.SNConv2d(4, 6, 3, padding=1))
    blk2.load_state_dict(ref.state_dict(), strict=True)
    assert torch.equal(blk2[0].weight_orig, blk[0].weight_orig)


def test_checkpoint_background_writer_errors_surface(tmp_path):
    from optimalstrategiesagainstgenerativeattacks_b200.checkpoints import CheckpointIO

    class Unpicklable:
        def state_dict(self):
            return {"f": lambda: None}                   # torch.save cannot pickle a lambda

        def load_state_dict(self, d):
            pass
    io = CheckpointIO(str(tmp_path), async_save=True, bad=Unpicklable())
    io.save(global_step=0, last_epoch=0, filename="m.pt")
    import pytest
    with pytest.raises(RuntimeError, match="background checkpoint write failed"):
        io.wait()
    io.wait()                                            # the error is reported once


def test_reference_archive_is_importable_when_built():
    """oracle/build_ref.py packs the unmodified reference into oracle/_ref/reference.zip (git-ignored); zipimport finds the packages."""
    import os
    import zipfile
    from oracle import build_ref
    path = build_ref.build()
    if path is None:
        import pytest
        pytest.skip("neither /root/reference nor a prebuilt archive is present")
    names = zipfile.ZipFile(path).namelist()
    assert "training/gim_img_trainer.py" in names and "models/gim_img_models.py" in names
    assert not [n for n in names if not n.endswith(".py")]
    assert os.path.getsize(path) < 1 << 20


def test_image_grid_and_png_writer(tmp_path):
    """sample_and_save_imgs' host side (reference training/gim_img_training.py:21-73, logger.py:43-52): grid layout and a valid PNG."""
    import struct
    import zlib
    from optimalstrategiesagainstgenerativeattacks_b200 import gim_img_training as T
    imgs = torch.rand(7, 3, 6, 5)
    grid = T.make_grid(imgs, nrow=5, padding=2)
    assert grid.shape == (3, 2 * 8 + 2, 5 * 7 + 2)
    assert torch.equal(grid[:, 2:8, 2:7], imgs[0]) and torch.equal(grid[:, 10:16, 9:14], imgs[6]) and float(grid[:, :2].abs().max()) == 0
    log = T.ScalarLog(img_dir=str(tmp_path))
    T.save_imgs(log, (imgs * 2 - 1).unsqueeze(0), "train imgs_0003", "leaked", 12)
    path = tmp_path / "train imgs_0003" / "leaked" / "00000012.png"
    raw = path.read_bytes()
    assert raw[:8] == b"\x89PNG\r\n\x1a\n"
    pos, chunks = 8, {}
    while pos < len(raw):
        n, tag = struct.unpack(">I4s", raw[pos:pos + 8])
        body = raw[pos + 8:pos + 8 + n]
        assert struct.unpack(">I", raw[pos + 8 + n:pos + 12 + n])[0] == zlib.crc32(tag + body) & 0xffffffff
        chunks[tag] = body
        pos += 12 + n
    w, h, depth, ctype = struct.unpack(">IIBB", chunks[b"IHDR"][:10])
    assert (w, h, depth, ctype) == (37, 18, 8, 2)
    rows = zlib.decompress(chunks[b"IDAT"])
    pix = np.frombuffer(rows, dtype=np.uint8).reshape(h, 1 + 3 * w)[:, 1:].reshape(h, w, 3)
    want = (log.images[("train imgs_0003", "leaked")][1].clamp(0, 1) * 255 + 0.5).to(torch.uint8).permute(1, 2, 0).numpy()
    assert np.array_equal(pix, want) and np.abs(want[2:8, 2:7].astype(float) / 255 - imgs[0].permute(1, 2, 0).numpy()).max() < 3e-3


def test_eval_auc_orientation_matches_reference_labels():
    """authentication_score.py:94-96: y_true = 1 for real, 0 for fake, y_score = the authenticator's logits.  An authenticator that scores
    real samples higher must get AUC 1, the opposite ordering AUC 0 (what an untrained authenticator shows on an untrained attacker's
    out-of-distribution images -- tools/bench_extra.py's `auc_random_init: 0.0` -- is this second case, not a label slip)."""
    ds = D.ResidentGIMDataSet(D.synthetic_classes(4, 16, 1, 4, seed=1), m=2, n=3, k=2, example_cnt_per_class=3, device="cpu", seed=2)
    au = AE.Authenticator(lambda test_sample, si_sample: test_sample.mean(dim=(1, 2, 3, 4)).unsqueeze(1), th=0.)
    for shift, want in ((-10.0, 1.0), (10.0, 0.0)):
        im = AE.Impersonator(lambda leaked_sample, n, s=shift: leaked_sample[:, :1].expand(-1, n, -1, -1, -1) + s)
        acc, acc_fake, acc_real, auc = AE.eval_authenticator_and_impersonator("cpu", ds, 4, 0, au, im)
        assert auc == want, (shift, auc)


def test_weight_gradient_streams_only_for_deferred_consumers():
    """ops._wg_async_ok: a weight gradient may be computed on the side stream only when its consumer is a spectral-norm node (whose
    backward just queues the tensor for the batched flush); plain parameters (AccumulateGrad reads the gradient at once) are refused."""
    import torch
    from optimalstrategiesagainstgenerativeattacks_b200 import ops
    w_orig = torch.nn.Parameter(torch.randn(8, 4, 3, 3))
    packed, aux = torch.randn(9, 8, 4), torch.zeros(8 + 36 + 1)
    w_sn = ops.SpectralNormPreparedFn.apply(w_orig, (packed, aux, None, None))
    assert ops._wg_async_ok(w_sn)
    assert not ops._wg_async_ok(w_orig)
    assert not ops._wg_async_ok(torch.randn(9, 8, 4))
    merged = ops._MergedRowsFn.apply((torch.randn(64), 0), torch.randn(1, 4, 8, requires_grad=True), torch.randn(1, 4, 8, requires_grad=True))
    assert ops._wg_async_ok(merged)
    ops._wg_join()                                   # nothing in flight: a no-op that must not touch CUDA


def _built_cabi():
    """The ctypes stub over a built library (builds it first on a fresh checkout; compiling needs no GPU)."""
    import os
    from optimalstrategiesagainstgenerativeattacks_b200 import _cabi as C
    if not os.path.exists(C.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return C


def test_conv_launch_plans_respect_the_hardware_limits():
    """gim_conv2d_fwd_plan (a dry run of the real launcher, no device needed) over every layer family of the O / V / 105x105 networks, small
    and benchmark-sized batches and every fused epilogue: shared memory <= 227 KB, TMEM <= 512 columns, ring depths, tile geometry, grid
    sizes -- and the choices DESIGN.md section 4 documents for the profiled shapes."""
    from optimalstrategiesagainstgenerativeattacks_b200 import _cabi as C
    EPI_LRELU, EPI_MASK, EPI_ADD, EPI_POOL, EPI_ADDUP, EPI_UNIT = 1, 2, 4, 8, 16, 32
    sizes = (1, 2, 4, 6, 8, 13, 16, 26, 32, 52, 64, 105)
    chans = ((64, 64), (64, 128), (128, 128), (128, 256), (256, 256), (256, 512), (512, 512), (512, 256), (256, 128), (128, 64), (1536, 1024),
             (16, 128), (128, 16), (88, 16), (320, 256), (32, 64), (72, 1000))
    checked = 0
    for hw in sizes:
        for ci, co in chans:
            for k in (1, 3, 9):
                for n in (5, 80, 640):
                    if n * hw * hw > 640 * 64 * 64:
                        continue
                    if not C.conv_tc_supported(n, hw, hw, ci, co, k, C.BF16):
                        continue
                    variants = [(C.F32, 0), (C.BF16, 0)]
                    if co % 32 == 0:
                        variants += [(C.BF16, EPI_LRELU), (C.F32, EPI_MASK), (C.F32, EPI_MASK | EPI_ADD)]
                        if hw % 2 == 0:
                            variants += [(C.F32, EPI_POOL | EPI_ADD), (C.F32, EPI_POOL | EPI_UNIT), (C.F32, EPI_MASK | EPI_ADDUP), (C.F32, EPI_ADDUP | EPI_UNIT)]
                    for out_dt, epi in variants:
                        try:
                            pl = C.conv_fwd_plan(n, hw, hw, ci, co, k, out_dt, epi)
                        except RuntimeError as e:            # the launcher may refuse (fused epilogues need cout >= 32 ...), never mis-size
                            assert "conv_fwd_tc" in str(e) or "conv2d" in str(e), str(e)
                            continue
                        tag = (n, hw, ci, co, k, out_dt, epi, pl)
                        assert pl["smem_bytes"] <= 227 * 1024, tag
                        t = pl["tmem_cols"]
                        assert 32 <= t <= 512 and (t & (t - 1)) == 0, tag
                        assert pl["bw"] * pl["bh"] * pl["bn"] == 128, tag
                        tiles = -(-hw // pl["bw"]) * -(-hw // pl["bh"]) * -(-n // pl["bn"])
                        assert pl["pixel_tiles"] == tiles, tag
                        assert pl["block_n"] in (16, 32, 64, 128, 256) and pl["block_k"] in (16, 64), tag
                        if pl["persistent"]:
                            assert pl["threads"] == 384 and 1 <= pl["grid_x"] <= 148 and pl["grid_y"] == 1, tag
                            assert pl["m_sub"] * pl["k_chains"] * pl["block_n"] * 2 <= 512, tag          # two accumulator buffers
                            if pl["pair"]:
                                assert pl["grid_x"] % 2 == 0 and pl["block_n"] in (128, 256) and co % pl["block_n"] == 0, tag
                            if pl["halo"]:
                                assert k == 3 and hw >= 8 and pl["bw"] == 8 and pl["block_k"] == 64, tag
                                assert 1 <= pl["a_stages"] <= 4 and 2 <= pl["b_stages"] <= 8 and pl["halo_bytes"] % 1024 == 0, tag
                                assert pl["halo_bytes"] >= 128 * (pl["bw"] + 2) * (pl["bh"] + 2) * pl["bn"], tag
                            else:
                                assert 2 <= pl["stages"] <= 8, tag
                        else:
                            assert pl["threads"] == 192 and epi & ~EPI_UNIT == 0 and 2 <= pl["stages"] <= 8, tag
                            assert pl["grid_x"] == tiles and pl["grid_y"] == -(-co // pl["block_n"]), tag
                        if epi & EPI_POOL:
                            assert pl["persistent"] and pl["bw"] <= 16, tag
                        checked += 1
    assert checked > 3000
    # the documented choices for the profiled layers (640 images = 128 episodes x 5 samples)
    for hw, c in ((32, 128), (16, 256), (8, 512)):
        pl = C.conv_fwd_plan(640, hw, hw, c, c, 3)
        assert (pl["persistent"], pl["halo"], pl["pair"], pl["block_n"], pl["m_sub"]) == (1, 1, 1, 128, 2), pl
    pl = C.conv_fwd_plan(640, 4, 4, 512, 512, 3)
    assert (pl["halo"], pl["pair"], pl["block_n"], pl["m_sub"]) == (0, 1, 256, 1), pl
    pl = C.conv_fwd_plan(640, 32, 32, 128, 128, 9)
    assert (pl["halo"], pl["pair"], pl["block_n"], pl["m_sub"]) == (0, 1, 128, 2), pl
    pl = C.conv_fwd_plan(640, 1, 1, 1536, 1024, 1)
    assert pl["persistent"] == 0 and pl["grid_x"] * pl["grid_y"] >= 148, pl


def test_weight_gradient_launch_plans():
    """gim_conv2d_wgrad_plan (dry run of the real launchers): shared memory, TMEM, split-K coverage, and the wave rule -- a grid never spills
    a few CTAs into a third wave (every CTA owns an SM: 297 CTAs on 148 SMs cost three waves, 294 cost two)."""
    C = _built_cabi()
    checked = 0
    for hw in (1, 2, 4, 8, 13, 16, 26, 32, 64, 105):
        for ci, co in ((64, 64), (64, 128), (128, 128), (128, 256), (256, 256), (256, 512), (512, 512), (512, 256), (256, 128), (1536, 1024), (16, 128),
                       (128, 16), (320, 256), (32, 64), (72, 1000)):
            for k in (1, 3, 9):
                for n in (5, 80, 640):
                    if n * hw * hw > 640 * 64 * 64 or not C.wgrad_tc_supported(n, hw, hw, ci, co, k, C.BF16):
                        continue
                    pl = C.conv_wgrad_plan(n, hw, hw, ci, co, k)
                    tag = (n, hw, ci, co, k, pl)
                    assert pl["smem_bytes"] <= 227 * 1024 and pl["stages"] >= 2 and pl["threads"] == 192, tag
                    t = pl["tmem_cols"]
                    assert 32 <= t <= 512 and (t & (t - 1)) == 0, tag
                    assert pl["grid_y"] >= 1 and pl["tiles_per_split"] * pl["grid_y"] >= pl["pixel_tiles"], tag            # split-K covers every pixel tile
                    assert pl["tiles_per_split"] * (pl["grid_y"] - 1) < pl["pixel_tiles"], tag                              # and no split is empty
                    ctas = pl["grid_x"] * pl["grid_y"]
                    if pl["grid_y"] > 1:
                        assert ctas <= 2 * 148, tag                                                                         # at most two waves
                    if pl["kernel"] == 3:
                        assert k == 3 and co <= 128 and pl["taps_per_cta"] == 3, tag
                    if pl["kernel"] == 2:
                        assert co % 256 == 0 and ci % 128 == 0 and pl["grid_x"] % 2 == 0, tag
                    checked += 1
    assert checked > 500
    hot = {(32, 128, 128): 294, (16, 256, 256): 288, (8, 512, 512): 288}
    for (hw, ci, co), ctas in hot.items():
        pl = C.conv_wgrad_plan(640, hw, hw, ci, co, 3)
        assert pl["grid_x"] * pl["grid_y"] == ctas, pl
