"""Install the UNMODIFIED reference sources the CPU baseline needs into git-ignored baseline/_ref/ (BASELINE.md section 3).

The GPU box only receives /root/repo, and the reference is a collection of scripts (no setup.py / pyproject), so the
"install" is a byte-for-byte copy of the files of the hot path; nothing is edited, and nothing under baseline/_ref
enters git history (.gitignore) or the product path (only bench.py's reference arm / cpu_baseline leg load it, through
baseline/ref_bench.py).  Run in the build container:  python baseline/install_ref.py
"""
import hashlib
import json
import os
import shutil
import sys

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
FILES = ["DeepBSDE.py", "1d_BSPDE_case.py", "nd_BSPDE_case.py", "with_corr_high_dimension_pde.py", "hjb_implement.py",
         "numerics/multidimensional_mc_pricer.py", "Functions/Sine.py", "Functions/naisnet.py", "Functions/networks.py"]


def install(verbose=True) -> bool:
    if not os.path.isdir(REF):
        if verbose:
            print(f"{REF} not present: keeping whatever baseline/_ref holds")
        return os.path.isdir(DST)
    manifest = {}
    for f in FILES:
        src, dst = os.path.join(REF, f), os.path.join(DST, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as fh:
            manifest[f] = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": REF, "sha256": manifest}, fh, indent=1)
    if verbose:
        print(f"installed {len(FILES)} reference files into {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if install() else 1)
