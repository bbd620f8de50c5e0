#!/usr/bin/env python
"""Stage the UNMODIFIED reference env package for bench.py's reference arm (test / measurement infrastructure).

    python baseline/install_ref.py            (runs where /root/reference exists: the build container)

The reference (xuecy22/NeuralPlane) is plain Python with no setup.py / pyproject, so there is nothing to pip-install:
this copies its `envs/` package (1.2 MB: env classes, tasks, reward / termination functions, the F16 / UAV models and
the 43 .pth coefficient nets) byte for byte into baseline/_ref/envs, plus the two 15-line stubs of the third-party
modules it imports but this image lacks (gym, torchdiffeq: SURVEY.md App. F; the same stubs the golden fixtures were
generated with, tests/golden/_shims).  baseline/_ref/ is git-ignored (never part of the repo's history) but travels to
the GPU box with the working tree, where `bench.py --impl reference` and its `cpu_baseline` leg import it:
ControlEnv(device='cpu') on every host core, and ControlEnv(device='cuda:0') as the labelled same-silicon eager run.
Nothing under neuralplane_b200/ imports it.
"""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("NPLANE_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")


def install(verbose=True):
    src = os.path.join(REF, "envs")
    if not os.path.isdir(src):
        if verbose:
            print(f"install_ref: {src} not found (not the build container) -- keeping whatever is in {DST}")
        return False
    if os.path.isdir(os.path.join(DST, "envs")):
        cmp = filecmp.dircmp(src, os.path.join(DST, "envs"), ignore=["__pycache__"])
        if not (cmp.left_only or cmp.right_only or cmp.diff_files):
            return True
    shutil.rmtree(DST, ignore_errors=True)
    shutil.copytree(src, os.path.join(DST, "envs"), ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    shutil.copytree(os.path.join(HERE, "..", "tests", "golden", "_shims"), os.path.join(DST, "_shims"),
                    ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    if verbose:
        print("install_ref: staged", os.path.join(DST, "envs"))
    return True


if __name__ == "__main__":
    sys.exit(0 if install() else 1)
