"""Summarise `ncu --set full` captures of tools/conv_bench.py into profiles/ (run here, no GPU needed):

  python tools/ncu_summarize.py gpurun_out/prof_conv_fwd_tc2_shapes_r01.ncu-rep gpurun_out/prof_conv_wgrad_tc_shapes_r01.ncu-rep

Writes profiles/ncu_conv_shapes_<round>.txt (table) and .json (read by bench.py for `roofline.traffic`); also accepts the raw-page CSV
exported on the GPU box (tools/capture_profiles.sh).
The captured command is `python tools/conv_bench.py --only fwd|wgrad --shapes 0,1,2 --no-check --reps 1` (640 images = B 128 x 5):
four launches per shape (three warm-up + one), the last of each shape is reported.
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHAPES = [(32, 32, 128, 128, 3), (16, 16, 256, 256, 3), (8, 8, 512, 512, 3)]
N_IMG = 640
M = {"dur": "gpu__time_duration.sum", "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum",
     "tensor": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "lts": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
     "dram": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "regs": "launch__registers_per_thread",
     "tma_ld": "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "grid": "Grid Size", "name": "Kernel Name"}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}


def load(rep):
    if rep.endswith(".csv"):          # already exported on the GPU box (tools/capture_profiles.sh): `ncu -i x.ncu-rep --page raw --csv`
        out = open(rep).read()
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    res = []
    for r in data:
        d = {}
        for k, name in M.items():
            i = hdr.index(name)
            v = r[i]
            try:
                v = float(v.replace(",", "")) * SCALE.get(units[i], 1.0)
            except ValueError:
                pass
            d[k] = v
        res.append(d)
    return res


def main():
    out_txt, out_json = [], {}
    for rep in sys.argv[1:]:
        rows = load(rep)
        kind = "wgrad" if "wgrad" in rows[0]["name"] else "fwd"
        per_shape = len(rows) // len(SHAPES)
        out_txt.append("%s kernels   (%s; last of %d launches per shape; <1> = cta_group::2 two-CTA cluster variant)" % (kind, os.path.basename(rep), per_shape))
        out_txt.append("  shape (640 images)            grid        us    TFLOP/s  tensor-pipe%  DRAM rd / wr MB   algorithmic MB  DRAM%  L2%   TMA-load GB (L2->SM TB/s)  regs")
        for si, (h, w, ci, co, k) in enumerate(SHAPES):
            r = rows[(si + 1) * per_shape - 1]
            flops = 2.0 * N_IMG * h * w * ci * co * k * k
            alg = N_IMG * h * w * (ci * 2 + (co * 4 if kind == "fwd" else co * 2)) + k * k * ci * co * (2 if kind == "fwd" else 4)
            variant = r["name"].split("(")[0].split("::")[-1].replace("void ", "")
            out_txt.append("  %2dx%-2d %3d->%-3d k%d   %14s  %7.1f  %8.1f  %8.1f      %7.1f / %-7.1f   %8.1f     %5.1f  %5.1f   %6.2f (%.1f)            %d   %s" % (
                h, w, ci, co, k, r["grid"], r["dur"], flops / r["dur"] / 1e6, r["tensor"], r["rd"] / 1e6, r["wr"] / 1e6, alg / 1e6, r["dram"], r["lts"],
                r["tma_ld"] / 1e9, r["tma_ld"] / r["dur"] / 1e6, int(r["regs"]), variant))
            out_json["%s_%dx%d_%d_%d_k%d" % (kind, h, w, ci, co, k)] = {
                "duration_us": r["dur"], "dram_bytes_read": r["rd"], "dram_bytes_write": r["wr"], "algorithmic_bytes": alg, "flops": flops,
                "tensor_pipe_active_pct": r["tensor"], "tma_load_bytes": r["tma_ld"]}
        out_txt.append("")
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    tag = os.environ.get("GIM_PROFILE_ROUND", "r02")
    open(os.path.join(ROOT, "profiles", "ncu_conv_shapes_%s.txt" % tag), "w").write("\n".join(out_txt) + "\n")
    json.dump(out_json, open(os.path.join(ROOT, "profiles", "ncu_conv_shapes_%s.json" % tag), "w"), indent=1)
    print("\n".join(out_txt))


if __name__ == "__main__":
    main()
