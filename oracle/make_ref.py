"""TEST / MEASUREMENT INFRASTRUCTURE -- stages the UNMODIFIED reference for the `bench.py --impl reference` arm.

The reference is plain Python (no setup.py / pyproject: `pip install /root/reference` has nothing to build), so "installing"
it is making its own source files importable on the GPU box, which has no /root/reference.  This recipe copies the files
the stage-1 / stage-2 renderers import -- byte for byte, nothing edited -- from /root/reference into oracle/_ref/, which is
git-ignored (it never enters the repository's history) but travels with the working tree to the GPU box like the built
.so files.  oracle/ref_harness.py then imports from there (NUNERF_REFERENCE_ROOT) behind the shim layer of SURVEY 8(c).

    python oracle/make_ref.py            (run by __graft_entry__.build() when /root/reference exists)
"""
import filecmp
import os
import shutil
import sys

SRC = os.environ.get("NUNERF_REFERENCE_SRC", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
# directories / files the renderer modules import at module top or open at construction (renderer_zerothick.py:1-18,
# field.py:583, utils/base_utils.load_cfg)
TREES = ["network", "utils", "dataset", "colmap"]
FILES = ["assets/bsdf_256_256.bin", "configs/shape/nerf/spherepot.yaml", "configs/stage2/nerf/spherepot.yaml", "LICENSE"]


def stage():
    if not os.path.isdir(os.path.join(SRC, "network")):
        return False
    for t in TREES:
        for root, _, files in os.walk(os.path.join(SRC, t)):
            for f in files:
                if f.endswith((".py", ".yaml", ".txt", ".bin", ".json")):
                    s = os.path.join(root, f)
                    d = os.path.join(DST, os.path.relpath(s, SRC))
                    os.makedirs(os.path.dirname(d), exist_ok=True)
                    if not (os.path.exists(d) and filecmp.cmp(s, d, shallow=False)):
                        shutil.copyfile(s, d)
    for f in FILES:
        s, d = os.path.join(SRC, f), os.path.join(DST, f)
        if os.path.exists(s):
            os.makedirs(os.path.dirname(d), exist_ok=True)
            if not (os.path.exists(d) and filecmp.cmp(s, d, shallow=False)):
                shutil.copyfile(s, d)
    return True


def verify():
    """Every staged file is byte-identical to its source (when the source tree is present)."""
    bad = []
    for root, _, files in os.walk(DST):
        for f in files:
            d = os.path.join(root, f)
            s = os.path.join(SRC, os.path.relpath(d, DST))
            if os.path.exists(s) and not filecmp.cmp(s, d, shallow=False):
                bad.append(d)
    return bad


if __name__ == "__main__":
    ok = stage()
    print("staged" if ok else f"{SRC} not present: nothing staged", DST)
    sys.exit(1 if verify() else 0)
