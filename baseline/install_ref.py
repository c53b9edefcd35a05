#!/usr/bin/env python
"""Place an UNMODIFIED copy of the reference where the GPU box can find it.

    python baseline/install_ref.py [--src /root/reference]

The reference (jensbrodersen/mc-trade-sizing) is not a package: no setup.py / pyproject.toml, nothing for pip to
install, no compilable sources.  Its hot path is `src/trading_models.py`, run either in process or as the CLI that
`dps.py` spawns.  `/root/reference` exists only in the build container, so this script copies the files the
measurements and tests need, byte for byte, into `baseline/_ref/` -- git-ignored (reference sources never enter the
history) but shipped to the GPU box by gpurun.  Users of the copy:
    bench.py                  cpu_baseline.python_reference: the real Python reference timed on the box's host cores
    tests/test_dps_overlay_*  the unchanged dps.py run end to end against the B200 drop-in
`__graft_entry__.build()` calls this whenever the source directory is present.
"""
import argparse
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
FILES = ("dps.py", "dps_config.yaml", "requirements.txt", "LICENSE", os.path.join("src", "trading_models.py"),
         os.path.join("src", "output_handler.py"), os.path.join("src", "influx_handler.py"), os.path.join("src", "api_handler.py"))


def install(src: str = "/root/reference", quiet: bool = False) -> bool:
    """Copy the reference files; returns False (and leaves any existing copy alone) when `src` is absent."""
    if not os.path.isdir(src):
        return False
    for rel in FILES:
        a, b = os.path.join(src, rel), os.path.join(DEST, rel)
        if not os.path.exists(a):
            continue
        os.makedirs(os.path.dirname(b), exist_ok=True)
        if not (os.path.exists(b) and filecmp.cmp(a, b, shallow=False)):
            shutil.copyfile(a, b)
            os.chmod(b, 0o644)
    if not quiet:
        print("reference copied to", DEST)
    return True


def ref_path(*parts: str):
    """Path inside the shipped copy, or None when it has not been installed."""
    p = os.path.join(DEST, *parts)
    return p if os.path.exists(p) else None


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    if not install(ap.parse_args().src):
        sys.exit("reference source directory not found")
