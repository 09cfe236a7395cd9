"""Import the UNMODIFIED reference from /root/reference (build container only).

The reference's adapters import pip modules that are absent here and carry no
arithmetic (wget, h5py, skimage, matplotlib); they are stubbed in sys.modules
before import, exactly as SURVEY.md §8(c) describes.  Nothing under
/root/reference is copied or modified.  This module must never be imported by
anything that runs on the GPU box (the tree does not exist there).
"""
import os
import sys
import types

REF = os.environ.get("CVD_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF, "loss"))


def setup():
    if not available():
        raise RuntimeError(f"reference tree not found at {REF}")
    for name in ("wget", "h5py", "skimage", "skimage.io", "matplotlib", "matplotlib.cm"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                m = types.ModuleType(name)
                if name == "skimage.io":
                    m.imsave = lambda *a, **k: None
                if name == "matplotlib.cm":
                    class _CM:
                        colors = [[i / 255.0] * 3 for i in range(256)]
                    m.get_cmap = lambda *a, **k: _CM()
                sys.modules[name] = m
    if "skimage" in sys.modules and "skimage.io" in sys.modules:
        setattr(sys.modules["skimage"], "io", sys.modules["skimage.io"])
    if "matplotlib" in sys.modules and "matplotlib.cm" in sys.modules:
        setattr(sys.modules["matplotlib"], "cm", sys.modules["matplotlib.cm"])
    if REF not in sys.path:
        sys.path.insert(0, REF)
