"""Build recipe for oracle/_ref: the reference's ONLY native component, compiled unmodified.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is imported by the product
package (ecord_b200/); only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may use it, and only as the checker.

What it does: cythonizes /root/reference/ecord/environment/_tradjectory_step_cy.pyx
*where it lies* (the reference tree is read-only; nothing is copied into the repo)
and compiles the generated C with gcc into

    oracle/_ref/_tradjectory_step_cy.<abi>.so

mirroring the reference's own recipe `python setup_cy.py build_ext --inplace`
(/root/reference/setup_cy.py:7,12-13,19-25: a plain Extension, no OpenMP, so
`prange` runs serially).  The intermediate .c file is deleted after the build:
oracle/_ref/ holds binaries only, is git-ignored, and travels to the GPU box.

If /root/reference is absent (GPU box) this is a no-op: the prebuilt .so is used.
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF_PYX = "/root/reference/ecord/environment/_tradjectory_step_cy.pyx"
OUT_DIR = os.path.join(HERE, "_ref")
MODNAME = "_tradjectory_step_cy"


def so_path():
    return os.path.join(OUT_DIR, MODNAME + sysconfig.get_config_var("EXT_SUFFIX"))


def build(force=False, quiet=True):
    """Returns the path of the built .so, or None if it cannot be built and is not prebuilt."""
    out = so_path()
    if os.path.exists(out) and not force:
        return out
    if not os.path.exists(REF_PYX):
        return out if os.path.exists(out) else None
    import numpy as np
    os.makedirs(OUT_DIR, exist_ok=True)
    c_file = os.path.join(OUT_DIR, MODNAME + ".c")
    # cython reads the .pyx in place and writes only the generated C into oracle/_ref/
    subprocess.check_call([sys.executable, "-m", "cython", "-3", "--module-name",
                           "ecord.environment." + MODNAME, REF_PYX, "-o", c_file],
                          stdout=subprocess.DEVNULL if quiet else None)
    inc = [sysconfig.get_paths()["include"], np.get_include()]
    cmd = ["gcc", "-shared", "-fPIC", "-O2", "-fwrapv", "-fno-strict-aliasing",
           "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION", "-w"]
    for i in inc:
        cmd += ["-I", i]
    cmd += [c_file, "-o", out]
    subprocess.check_call(cmd)
    os.remove(c_file)
    return out


def load_step_cy():
    """Import the compiled reference module under the name the reference package expects
    and return its `step_cy`.  Returns None if the .so is unavailable."""
    path = build()
    if path is None or not os.path.exists(path):
        return None
    import importlib.machinery
    import importlib.util
    full = "ecord.environment." + MODNAME
    if full in sys.modules:
        return sys.modules[full].step_cy
    loader = importlib.machinery.ExtensionFileLoader(full, path)
    spec = importlib.util.spec_from_loader(full, loader, origin=path)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    sys.modules[full] = mod
    return mod.step_cy


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, quiet=False)
    print("oracle/_ref:", p)
