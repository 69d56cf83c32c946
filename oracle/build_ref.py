"""TEST INFRASTRUCTURE ONLY -- builds the parts of the unmodified reference that compile from their own few source files into oracle/_ref/
(git-ignored; travels to the GPU box with the snapshot).  Nothing is copied into the repository: the sources are read where they lie
under /root/reference, generated C++ and the extension module go to oracle/_ref/.

    python oracle/build_ref.py        (also called by __graft_entry__.build() when /root/reference is present)

Built here:  mise  <-  src/vgn/ConvONets/utils/libmise/mise.pyx   (Cython -> C++ -> CPython extension; pins oracle/mise_oracle.py)
The network path itself is Python/PyTorch (imported in this container by tests/golden/make_golden.py, not buildable into _ref).
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "oracle", "_ref")
REF = "/root/reference/src/vgn/ConvONets/utils/libmise/mise.pyx"


def build(force: bool = False) -> str | None:
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    so = os.path.join(OUT, "mise" + ext)
    if os.path.exists(so) and not force:
        return so
    if not os.path.exists(REF):
        return None                      # no reference tree here (e.g. the GPU box): use the prebuilt module if it travelled
    try:
        import Cython  # noqa: F401
        import numpy
    except ImportError:
        return None
    os.makedirs(OUT, exist_ok=True)
    cpp = os.path.join(OUT, "mise.cpp")
    subprocess.run([sys.executable, "-m", "cython", "--cplus", "-3", REF, "-o", cpp], check=True, capture_output=True)
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-w", "-I", sysconfig.get_paths()["include"], "-I", numpy.get_include(), cpp, "-o", so], check=True,
                   capture_output=True)
    return so


def load_mise():
    """The compiled reference MISE module, or None when it is not available."""
    so = build()
    if not so:
        return None
    import importlib.util
    spec = importlib.util.spec_from_file_location("mise", so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
