"""Pack the UNMODIFIED reference (pure Python) into oracle/_ref/reference.zip so that it can travel to the GPU box.

TEST / BASELINE INFRASTRUCTURE ONLY.  The reference has no native code, so "building" it is packaging: the python packages on
the GIM path are stored byte-for-byte in ONE zip archive (a binary artefact, git-ignored like our own .so files, never
committed) and imported from there with zipimport (`oracle.ref_shim.install()` puts the archive on sys.path).  Used by
  * bench.py's `cpu_baseline` leg and `--impl reference` arm (kind = "reference": the reference's own trainer on the host CPU),
  * bench.py's `gpu_eager_baseline` leg (the same code with device='cuda': eager PyTorch/cuDNN, the GPU bar of BASELINE.md 5.4),
  * oracle/make_golden.py in the build container.
Run by `__graft_entry__.build()` whenever /root/reference exists (the GPU box only uses the prebuilt archive).

    python oracle/build_ref.py [reference_root]
"""
import os
import sys
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_ref")
ARCHIVE = os.path.join(OUT_DIR, "reference.zip")
PACKAGES = ("models", "training", "authentication_eval", "data_handling", "theory")


def build(reference_root="/root/reference", archive=ARCHIVE):
    """-> path of the archive (rebuilt when any source is newer), or None when the reference tree is absent."""
    if not os.path.isdir(reference_root):
        return archive if os.path.exists(archive) else None
    files = []
    for pkg in PACKAGES:
        base = os.path.join(reference_root, pkg)
        for d, _, names in os.walk(base):
            for nm in sorted(names):
                if nm.endswith(".py"):
                    files.append(os.path.join(d, nm))
    if not files:
        raise RuntimeError("no python sources under %s" % reference_root)
    if os.path.exists(archive) and all(os.path.getmtime(f) <= os.path.getmtime(archive) for f in files):
        return archive
    os.makedirs(os.path.dirname(archive), exist_ok=True)
    tmp = archive + ".tmp"
    with zipfile.ZipFile(tmp, "w", zipfile.ZIP_DEFLATED) as z:
        for f in sorted(files):
            z.write(f, os.path.relpath(f, reference_root))
    os.replace(tmp, archive)
    return archive


if __name__ == "__main__":
    print(build(*(sys.argv[1:2] or ["/root/reference"])))
