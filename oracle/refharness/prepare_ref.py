"""TEST INFRASTRUCTURE -- not part of the product path.

Makes the *reference's own* Python importable in this container so that its
numpy builders can be executed to produce golden vectors (tests/golden/).

The reference (/root/reference, read-only) is pure Python but targets
numpy<1.24 / COMPASS; three things stop it importing here:

  1. ``ndarray.itemset`` (removed in numpy 2) used by
     shesha/util/iterkolmo.py:56-66,85-94   -> source patch in a /tmp copy
  2. ``np.int / np.float / np.math``          -> runtime aliases on the numpy module
  3. missing third-party modules (h5py, astropy.io.fits, matplotlib, gym,
     tensorboardX, docopt, carmaWrap, sutraWrap) -> stub modules in sys.modules

Nothing from the reference is copied into this repository: the scratch copy
lives under /tmp and is rebuilt on every call.
"""
import math
import os
import re
import shutil
import sys
import types

REF_ROOT = os.environ.get("AOMARL_REFERENCE", "/root/reference")
SCRATCH = os.environ.get("AOMARL_REF_SCRATCH", "/tmp/aomarl_ref")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "shesha"))


def _patch_itemset(text: str) -> str:
    # X.itemset(idx, val)  ->  X.flat[idx] = val     (single-line calls only)
    return re.sub(r"(\w+)\.itemset\((.+),\s*([^,()]+)\)\s*$", r"\1.flat[\2] = \3", text,
                  flags=re.M)


def build_scratch_copy() -> str:
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    if os.path.isdir(SCRATCH):
        shutil.rmtree(SCRATCH)
    os.makedirs(SCRATCH)
    for sub in ("shesha", "src", "data/par"):
        shutil.copytree(os.path.join(REF_ROOT, sub), os.path.join(SCRATCH, sub))
    p = os.path.join(SCRATCH, "shesha/util/iterkolmo.py")
    with open(p) as f:
        txt = f.read()
    with open(p, "w") as f:
        f.write(_patch_itemset(txt))
    return SCRATCH


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _Anything:
    """Attribute sink used for plotting / logging stubs."""

    def __init__(self, *a, **k):
        pass

    def __getattr__(self, item):
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()


def install_stubs():
    import numpy as np
    np.int = int
    np.float = float
    np.math = types.SimpleNamespace(factorial=lambda x: math.factorial(int(x)))

    _stub("h5py", File=_Anything)
    astropy = _stub("astropy")
    astropy_io = _stub("astropy.io")
    fits = _stub("astropy.io.fits", open=_Anything, writeto=_Anything, getdata=_Anything)
    astropy.io = astropy_io
    astropy_io.fits = fits
    mpl = _stub("matplotlib", use=lambda *a, **k: None)
    plt = _stub("matplotlib.pyplot")

    def _plt_getattr(name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()

    plt.__getattr__ = _plt_getattr
    gs = _stub("matplotlib.gridspec", GridSpec=_Anything)
    mpl.pyplot = plt
    mpl.gridspec = gs

    class _Env:
        pass

    class _Box:
        def __init__(self, low=None, high=None, shape=None, dtype=None):
            self.shape = tuple(shape)
            self.dtype = dtype

    spaces = _stub("gym.spaces", Box=_Box)
    _stub("gym", Env=_Env, spaces=spaces)
    _stub("tensorboardX", SummaryWriter=_Anything)
    _stub("docopt", docopt=_Anything)
    _stub("torchvision")
    _stub("torchvision.transforms")


def activate(fake_sutra=True):
    """Build the scratch copy, install stubs, put it first on sys.path."""
    root = build_scratch_copy()
    install_stubs()
    if fake_sutra:
        from . import fake_sutra as fs
        fs.install()
    os.environ["SHESHA_ROOT"] = root
    if root not in sys.path:
        sys.path.insert(0, root)
    return root
