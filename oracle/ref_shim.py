"""Import shim that makes the UNMODIFIED reference importable under torch 2.11 in THIS container.

TEST INFRASTRUCTURE ONLY (used by oracle/make_golden.py and the cpu_baseline leg when /root/reference
exists).  Nothing here is copied from the reference; it only stubs two absent third-party modules and
restores torch-1.2 behaviour for InstanceNorm2d on a 1x1 map.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("GIM_REFERENCE_ROOT", "/root/reference")


def install(reference_root=REFERENCE_ROOT):
    if not os.path.isdir(reference_root):
        raise RuntimeError("reference tree not present at %s" % reference_root)
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
