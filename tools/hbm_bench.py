"""HBM-bound kernels of the GIM path against the measured HBM bandwidth (MEASURED_PEAKS.json: 6550 GB/s), one launch each at the
shapes the O workload (B = 128 -> 640 images per sample set) and the d = 1000 Gaussian workload actually run:

  python tools/hbm_bench.py                       # CUDA-event timing, L2 flushed between launches -> JSON lines (achieved GB/s)
  ncu --set full --profile-from-start off -o gpurun_out/hbm_r02 python tools/hbm_bench.py --ncu   # every case once inside the range
  python tools/hbm_bench.py --merge gpurun_out/hbm_raw.csv   # `ncu -i .. --page raw --csv` -> profiles/ncu_hbm_r02.json

"algorithmic bytes" = every input read once + every output written once, at the dtypes the kernel uses.
"""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = []


def case(name, kernels):
    def deco(fn):
        CASES.append((name, kernels, fn))
        return fn
    return deco


def build_cases():
    import torch
    from optimalstrategiesagainstgenerativeattacks_b200 import _cabi as C
    from optimalstrategiesagainstgenerativeattacks_b200 import fused_adam, model_blocks as mb, ops
    dev = torch.device("cuda", 0)
    f32, bf = torch.float32, torch.bfloat16
    N = 640

    def rnd(*s, dtype=f32):
        return torch.randn(s, device=dev, dtype=f32).to(dtype)

    # InstanceNorm / ada_in passes on the 128-channel 32x32 activations of the attacker's last up block
    x = rnd(N, 32, 32, 128)
    st = torch.empty((4, N, 128), device=dev)
    y = torch.empty_like(x)
    gy = rnd(N, 32, 32, 128)
    red = torch.empty((5, N, 128), device=dev)
    ones = torch.ones(128, device=dev)
    nb = x.numel() * 4

    @case("norm_stats [640,32,32,128] fp32", ["norm_stats_kernel"])
    def _():
        C.call("gim_norm_stats", C.ptr(x), st[0].data_ptr(), st[1].data_ptr(), N, 1024, 128, C.F32)
        return nb                                  # one pass (pivot-shifted sums)

    C.call("gim_norm_stats", C.ptr(x), st[0].data_ptr(), st[1].data_ptr(), N, 1024, 128, C.F32)
    C.call("gim_norm_coeffs", 0, st[0].data_ptr(), st[1].data_ptr(), C.ptr(ones), C.ptr(ones), st[2].data_ptr(), st[3].data_ptr(), N, 1024, 128, 1e-5)

    @case("affine_act (normalise + LeakyReLU) [640,32,32,128] fp32", ["affine_act_kernel", "affine_act_vec4_kernel"])
    def _():
        C.call("gim_affine_act_fwd", C.ptr(x), st[0].data_ptr(), st[2].data_ptr(), st[3].data_ptr(), C.ptr(y), N, 1024, 128, 0.2, C.F32)
        return 2 * nb

    @case("norm_bwd_reduce [640,32,32,128] fp32", ["norm_bwd_reduce_kernel"])
    def _():
        C.call("gim_norm_bwd_reduce", C.ptr(gy), C.ptr(x), C.ptr(y), st[0].data_ptr(), None, None, red[0].data_ptr(), red[1].data_ptr(), N, 1024, 128, 0.2, C.F32)
        return 3 * nb

    @case("norm_bwd_apply [640,32,32,128] fp32", ["norm_bwd_apply_kernel", "norm_bwd_apply_vec4_kernel"])
    def _():
        C.call("gim_norm_bwd_apply", C.ptr(gy), C.ptr(x), C.ptr(y), st[0].data_ptr(), None, None, red[2].data_ptr(), red[3].data_ptr(), red[4].data_ptr(), C.ptr(y), N, 1024,
               128, 0.2, C.F32)
        return 4 * nb

    # Conv2d backward helper: bf16 operand copy of the output gradient + bias gradient in one pass
    op = torch.empty(x.shape, device=dev, dtype=bf)
    sums = torch.zeros(128, device=dev)

    @case("cast_colsum (fp32 -> bf16 operand + bias grad) [655360,128]", ["cast_colsum_kernel"])
    def _():
        C.call("gim_cast_colsum", C.ptr(gy), C.ptr(op), C.ptr(sums), N * 1024, 128, 0)
        return nb + nb // 2

    gp = rnd(N, 16, 16, 128)

    @case("unpool2_cast (AvgPool backward -> bf16 operand) [640,32,32,128]", ["unpool2_cast_kernel", "unpool2_cast_even_kernel"])
    def _():
        C.call("gim_unpool2_cast", C.ptr(gp), C.ptr(op), N, 32, 32, 128, 0.25)
        return gp.numel() * 4 + op.numel() * 2

    y32 = torch.empty((N, 16, 16, 128), device=dev)
    yb, yl = torch.empty_like(y32, dtype=bf), torch.empty_like(y32, dtype=bf)

    @case("pool2_multi (AvgPool(a + b) -> fp32 + bf16 + bf16 LeakyReLU) [640,32,32,128]", ["pool2_multi_kernel"])
    def _():
        C.call("gim_pool2_multi", C.ptr(x), C.ptr(gy), C.ptr(y32), C.ptr(yb), C.ptr(yl), N, 32, 32, 128, 0.25, 0.2)
        return 2 * nb + y32.numel() * 8

    # image-side first block (1 -> 128 channels, 3x3) forward and weight gradient
    img = rnd(N, 32, 32, 1)
    w1, wl = rnd(9, 128, 1), rnd(1, 128, 1)
    b1, bl = rnd(128), rnd(128)
    tl = torch.empty((N, 32, 32, 128), device=dev, dtype=bf)
    res = torch.empty((N, 16, 16, 128), device=dev)

    @case("first_block_fwd (1 -> 128 ch, both input convs) [640,32,32]", ["first_block_mma_kernel", "first_block_kernel"])
    def _():
        C.call("gim_first_block_fwd", C.ptr(img), C.ptr(w1), C.ptr(b1), C.ptr(wl), C.ptr(bl), C.ptr(tl), C.ptr(res), N, 32, 32, 1, 128, 3, 0.2)
        return img.numel() * 4 + tl.numel() * 2 + res.numel() * 4

    gw1, gwl = torch.empty((9, 128, 1), device=dev), torch.empty((1, 128, 1), device=dev)
    scratch = torch.empty(32 * 10 * 128, device=dev)

    @case("first_block_wgrad (both image-side weight gradients) [640,32,32] x 128 ch", ["first_block_wgrad_mma_kernel", "first_block_wgrad_kernel"])
    def _():
        C.call("gim_first_block_wgrad", C.ptr(img), C.ptr(tl), C.ptr(res), C.ptr(gw1), C.ptr(gwl), C.ptr(scratch), scratch.numel(), N, 32, 32, 1, 128, 3, 0.2)
        return img.numel() * 4 + tl.numel() * 2 + res.numel() * 4

    # encoder tail and set statistics
    feat = rnd(N, 4, 4, 512)
    gm, gi = torch.empty((N, 512), device=dev), torch.empty((N, 512), device=dev, dtype=torch.int32)

    @case("global max + LeakyReLU [640,4,4,512]", ["gmax_fwd_kernel"])
    def _():
        C.call("gim_gmax_fwd", C.ptr(feat), C.ptr(gm), C.ptr(gi), N, 16, 512, 0.2, C.F32)
        return feat.numel() * 4 + gm.numel() * 8

    gs = rnd(4096, 15, 1000)
    gout = torch.empty((4096, 2000), device=dev)

    @case("set statistics mean|std, Gaussian d=1000 [4096,15,1000]", ["set_stats_fwd_kernel"])
    def _():
        C.call("gim_set_stats_fwd", C.ptr(gs), C.ptr(gout), C.ptr(gout) + 4000, 2000, 4096, 15, 1000, 1.0 / 15, 1e-8)
        return gs.numel() * 4 + gout.numel() * 4

    ggs = rnd(4096, 2000)
    gxs = torch.empty_like(gs)

    @case("set statistics backward [4096,15,1000]", ["set_stats_bwd_kernel"])
    def _():
        C.call("gim_set_stats_bwd", C.ptr(ggs), C.ptr(ggs) + 4000, 2000, C.ptr(gs), C.ptr(gxs), 4096, 15, 1000, 1.0 / 15, 1e-8)
        return 2 * gs.numel() * 4 + ggs.numel() * 4

    # spectral norm, batched over the twelve 512 -> 512 3x3 convolutions of the AdaIN residual stack + up block (28.3 M parameters)
    convs = [mb.SNConv2d(512, 512, 3, padding=1).to(dev) for _ in range(12)]
    n_par = sum(m.weight_orig.numel() for m in convs)
    ops.set_precision("bf16")

    @case("spectral norm forward, 12 x (512->512 3x3): power iteration + sigma + fp32/bf16/flipped-bf16 packs", ["sn_wtu_multi_kernel", "sn_vnorm_multi_kernel", "sn_wv_multi_kernel", "sn_unorm_multi_kernel", "sn_pack_multi_kernel", "sn_pack_multi_vec_kernel"])
    def _():
        ops.sn_prepare(convs, True, 1e-12)
        return n_par * (3 * 4 + 4 + 2 + 2)          # W read by W^T u, W v and the pack; W/sigma written in fp32 and twice in bf16

    grads = [rnd(9, 512, 512) for _ in convs]
    for m in convs:
        m.weight_orig.grad = torch.zeros_like(m.weight_orig)
    ops.sn_prepare(convs, True, 1e-12)
    preps = [m._prepared for m in convs]

    @case("spectral norm backward, 12 x (512->512 3x3): sum(G.W) then G/sigma - c u v^T accumulated into .grad", ["sn_bwd_dot_multi_kernel", "sn_bwd_apply_multi_kernel", "sn_bwd_dot_multi_vec_kernel", "sn_bwd_apply_multi_vec_kernel"])
    def _():
        table = (C.SnBwdLayer * len(convs))()
        sc = torch.empty(len(convs), device=dev)
        for i, (m, g, p) in enumerate(zip(convs, grads, preps)):
            aux = p[1]
            e = table[i]
            e.g, e.w, e.u, e.v, e.sigma = g.data_ptr(), m.weight_orig.data_ptr(), aux.data_ptr(), aux[512:].data_ptr(), aux[512 + 4608:].data_ptr()
            e.grad, e.scratch = m.weight_orig.grad.data_ptr(), sc[i:].data_ptr()
            e.cout, e.cin, e.ksize, e.accumulate = 512, 512, 3, 1
        C.call("gim_sn_backward_multi", ctypes.cast(table, ctypes.c_void_p), len(convs))
        return n_par * (2 * 4 + 3 * 4)              # dot: G, W; apply: G, grad in, grad out

    # fused Adam over the same parameters (one launch)
    opt = fused_adam.FusedAdam([m.weight_orig for m in convs], lr=1e-5, betas=(0.0, 0.99))
    opt.step()

    @case("fused Adam, 28.3 M parameters in one launch", ["adam_multi_kernel"])
    def _():
        opt.step()
        return n_par * 28                           # p, g, m, v read; p, m, v written
    return dev


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ncu", action="store_true")
    ap.add_argument("--merge", default=None)
    a = ap.parse_args()
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    peak = float(peaks.get("hbm_gbs", 6550.0))
    if a.merge:
        return merge(a.merge, peak)
    import torch
    torch.cuda.set_device(0)
    build_cases()
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")        # > 126 MB L2
    out = []
    if a.ncu:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        for name, kernels, fn in CASES:
            flush.zero_()
            fn()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    for name, kernels, fn in CASES:
        for _ in range(2):
            fn()
        ms = []
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            nbytes = fn()
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        t = sorted(ms)[len(ms) // 2]
        rec = {"case": name, "kernels": kernels, "algorithmic_bytes": int(nbytes), "ms": t, "achieved_gbs": nbytes / t / 1e6, "peak_gbs": peak,
               "frac": nbytes / t / 1e6 / peak}
        out.append(rec)
        print(json.dumps(rec), flush=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "hbm_events_r02.json"), "w"), indent=1)


def merge(raw_csv, peak):
    """Join the event timings (gpurun_out/hbm_events_r02.json) with the ncu raw page (dram bytes, duration per kernel) by kernel name."""
    import csv
    ev = json.load(open(os.path.join(ROOT, "gpurun_out", "hbm_events_r02.json")))
    rows = list(csv.reader(l for l in open(raw_csv) if l.startswith('"')))
    hdr = rows[0]
    col = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) == len(hdr)]
    per_kernel = {}
    for r in data:
        nm = r[col["Kernel Name"]]
        def num(key):
            try:
                return float(r[col[key]].replace(",", ""))
            except Exception:
                return None
        per_kernel.setdefault(nm, []).append({"dram_read": num("dram__bytes_read.sum"), "dram_write": num("dram__bytes_write.sum"), "ns": num("gpu__time_duration.sum"),
                                              "dram_pct": num("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")})
    units = {h: rows[1][i] for i, h in enumerate(hdr)}
    def to_bytes(v, key):
        u = units.get(key, "byte").lower()
        return None if v is None else v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    def to_ns(v):
        u = units.get("gpu__time_duration.sum", "ns").lower()
        return None if v is None else v * {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1, "second": 1e9}.get(u, 1)
    for rec in ev:
        tot_r = tot_w = tot_ns = 0.0
        found = []
        for k in rec["kernels"]:
            for nm, launches in per_kernel.items():
                if k in nm:
                    for l in launches:
                        tot_r += to_bytes(l["dram_read"], "dram__bytes_read.sum") or 0
                        tot_w += to_bytes(l["dram_write"], "dram__bytes_write.sum") or 0
                        tot_ns += to_ns(l["ns"]) or 0
                    found.append("%s x%d" % (nm.split("(")[0][-48:], len(launches)))
        rec["ncu"] = {"dram_bytes_read": tot_r, "dram_bytes_write": tot_w, "dram_bytes": tot_r + tot_w, "duration_us_under_ncu": tot_ns / 1e3,
                      "traffic_over_algorithmic": (tot_r + tot_w) / rec["algorithmic_bytes"] if rec["algorithmic_bytes"] else None,
                      "gbs_under_ncu": (tot_r + tot_w) / tot_ns if tot_ns else None, "launches": found}
    json.dump({"peak_hbm_gbs": peak, "peak_source": "MEASURED_PEAKS.json hbm_gbs", "note": "ms / achieved_gbs: CUDA events, median of 5, L2 flushed before each launch; "
               "ncu: one launch per case under `ncu --set full --clock-control none` (cold, serialised)", "cases": ev},
              open(os.path.join(ROOT, "profiles", "ncu_hbm_r02.json"), "w"), indent=1)
    for rec in ev:
        print("%-100s %7.1f us %6.0f GB/s (%.2f of peak)  dram/algorithmic %.2f" % (rec["case"][:100], rec["ms"] * 1e3, rec["achieved_gbs"], rec["frac"],
                                                                                  rec["ncu"]["traffic_over_algorithmic"] or 0))


if __name__ == "__main__":
    main()
