"""baseline/install_ref.py -- put the UNMODIFIED reference under baseline/_ref (git-ignored; it travels to the GPU box).

    python baseline/install_ref.py            # needs /root/reference (build container only)

Step 1 is the install the task prescribes,
    pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target baseline/_ref <copy of /root/reference>
(from a copy under /tmp: the build writes the cythonized C file into the source tree, and /root/reference is read-only).
It succeeds, but the reference's setup.py lists `find_packages(include=["ecord"])`, i.e. ONLY the top-level package:
the wheel holds ecord/__init__.py, ecord/utils.py and the compiled Cython step, none of the sub-packages
(environment, models, solvers, validate, ...).  Step 2 therefore copies the missing sub-packages and `experiments/`
verbatim next to them.  Nothing is edited; nothing lands in the git history (baseline/_ref/ is ignored).

The reference still cannot be imported on this stack by itself: torch_scatter / torch_sparse / torch_geometric are
not installed.  tests/test_reference_driver_gpu.py puts oracle/shims/ (stand-ins for exactly the symbols
the validate path touches) in front of it on sys.path.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
SRC = "/root/reference"


def main():
    if not os.path.isdir(SRC):
        print("no /root/reference here: baseline/_ref is used as shipped")
        return 0
    tmp = "/tmp/ecord_ref_copy"
    shutil.rmtree(tmp, ignore_errors=True)
    shutil.copytree(SRC, tmp, ignore=shutil.ignore_patterns("graphs", "imgs", ".git"))
    shutil.rmtree(DST, ignore_errors=True)
    rc = subprocess.call([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
                          "--find-links", "/opt/wheelhouse", "--target", DST, tmp])
    print("pip install rc =", rc)
    for sub in ("ecord", "experiments"):
        for root, dirs, files in os.walk(os.path.join(SRC, sub)):
            rel = os.path.relpath(root, SRC)
            os.makedirs(os.path.join(DST, rel), exist_ok=True)
            for f in files:
                if f.endswith((".py", ".pyx")) and not os.path.exists(os.path.join(DST, rel, f)):
                    shutil.copy2(os.path.join(root, f), os.path.join(DST, rel, f))
    so = [f for f in os.listdir(os.path.join(DST, "ecord", "environment")) if f.endswith(".so")]
    print("installed:", DST, "compiled step:", so)
    return 0 if so else 1


if __name__ == "__main__":
    sys.exit(main())
