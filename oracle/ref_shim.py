"""Import shim that makes the UNMODIFIED reference importable under torch 2.11 in THIS container.

TEST INFRASTRUCTURE ONLY (used by oracle/make_golden.py and the cpu_baseline leg when /root/reference
exists).  Nothing here is copied from the reference; it only stubs two absent third-party modules and
restores torch-1.2 behaviour for InstanceNorm2d on a 1x1 map.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("GIM_REFERENCE_ROOT", "/root/reference")
ARCHIVE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "reference.zip")


def available():
    """Where the unmodified reference can be imported from: its source tree (build container) or the archive packed by
    oracle/build_ref.py (GPU box), else None."""
    if os.path.isdir(REFERENCE_ROOT):
        return REFERENCE_ROOT
    if os.path.exists(ARCHIVE):
        return ARCHIVE
    return None


def install(reference_root=None):
    reference_root = reference_root or available()
    if reference_root is None or not os.path.exists(reference_root):
        raise RuntimeError("reference not present (neither %s nor %s)" % (REFERENCE_ROOT, ARCHIVE))
    if "colorama" not in sys.modules:
        c = types.ModuleType("colorama")

        class _Fore:
            YELLOW = ""
            RESET = ""
        c.Fore = _Fore
        sys.modules["colorama"] = c
    if "tensorboardX" not in sys.modules:
        t = types.ModuleType("tensorboardX")

        class SummaryWriter:
            def __init__(self, *a, **k):
                pass

            def __getattr__(self, name):
                return lambda *a, **k: None
        t.SummaryWriter = SummaryWriter
        sys.modules["tensorboardX"] = t
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    sys.dont_write_bytecode = True
    import torch.nn.functional as F
    F._verify_spatial_size = lambda size: None      # torch-1.2 semantics for the 1x1 InstanceNorm2d
    return reference_root
