"""
TEST INFRASTRUCTURE - not part of the product path.

Makes the unmodified reference (pure Python, mounted read-only at /root/reference) importable in
THIS container so that golden vectors can be generated and the numpy restatement in
``oracle/oracle.py`` can be pinned against it. The GPU box has no /root/reference: nothing that
runs there may import this module (tests guard with ``reference_available()``).

Three shims, none of which touches the arithmetic of the path:
  1. the package uses absolute imports ``from pyjjasim....`` but has no installer -> expose the
     directory under the name ``pyjjasim`` through a symlink in ``baseline/_ref`` (git-ignored);
  2. matplotlib is imported at module top but is not installed -> permissive stub modules;
  3. ``np.asscalar`` (time_evolution.py:322) no longer exists in numpy 2 -> one-line replacement.
"""
import os
import sys
import types
import warnings

import numpy as np

REFERENCE_DIR = "/root/reference"
_REF_PARENT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_DIR, "time_evolution.py"))


class _Stub(types.ModuleType):
    def __getattr__(self, key):
        if key.startswith("__"):
            raise AttributeError(key)
        return _Stub(self.__name__ + "." + key)

    def __call__(self, *args, **kwargs):
        return _Stub("stub")


def import_reference():
    """Return the reference package as module ``pyjjasim``."""
    if "pyjjasim" in sys.modules:
        return sys.modules["pyjjasim"]
    if not reference_available():
        raise RuntimeError("reference source tree not present (only available in the build container)")
    os.makedirs(_REF_PARENT, exist_ok=True)
    link = os.path.join(_REF_PARENT, "pyjjasim")
    if not os.path.islink(link):
        os.symlink(REFERENCE_DIR, link)
    for name in ["matplotlib", "matplotlib.pyplot", "matplotlib.collections", "matplotlib.colors",
                 "matplotlib.animation", "matplotlib.cm", "matplotlib.figure", "matplotlib.axes",
                 "matplotlib.patches", "matplotlib.lines", "matplotlib.path", "matplotlib.transforms"]:
        sys.modules.setdefault(name, _Stub(name))
    if not hasattr(np, "asscalar"):
        np.asscalar = lambda a: np.asarray(a).item()
    sys.path.insert(0, _REF_PARENT)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import pyjjasim
    return pyjjasim


def replay_noise(Nj, W, Nt, seed):
    """The reference's exact Gaussian draw sequence (time_evolution.py:533-537) as a
    (Nt, Nj, W) array, for feeding into the oracle / the device path's noise-injection hook."""
    rs = np.random.RandomState(seed)
    out = np.empty((Nt, Nj, W))
    rand = None
    for i in range(Nt):
        if Nj > 500:
            rand = rs.randn(Nj, W) if i % 3 == 0 else rand[rs.permutation(Nj), :]
        else:
            rand = rs.randn(Nj, W)
        out[i] = rand
    return out
