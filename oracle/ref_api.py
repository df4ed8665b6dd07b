"""Load the compiled reference rasterizer (oracle/_ref) without shadowing our drop-in module.

TEST / BENCH INFRASTRUCTURE ONLY.  The reference's package is also called
`diff_gaussian_rasterization`; it is imported here under the private name
`_gdr_reference_rasterizer` from oracle/_ref/diff_gaussian_rasterization/
(built by oracle/build_ref.py).  Needs a GPU to *run*; importing works anywhere.
"""
from __future__ import annotations

import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(HERE, "_ref", "diff_gaussian_rasterization")
_NAME = "_gdr_reference_rasterizer"


def available() -> bool:
    return os.path.isfile(os.path.join(PKG_DIR, "_C.so")) and os.path.isfile(os.path.join(PKG_DIR, "__init__.py"))


def load():
    """Returns the reference's module (GaussianRasterizationSettings, GaussianRasterizer, _C, ...)."""
    if _NAME in sys.modules:
        return sys.modules[_NAME]
    if not available():
        raise RuntimeError("oracle/_ref is not built; run `python oracle/build_ref.py` where /root/reference exists")
    import torch  # noqa: F401  (the extension links against libtorch)

    spec = importlib.util.spec_from_file_location(_NAME, os.path.join(PKG_DIR, "__init__.py"),
                                                  submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[_NAME] = mod
    try:
        spec.loader.exec_module(mod)
    except Exception:
        del sys.modules[_NAME]
        raise
    return mod
