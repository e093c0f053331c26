#!/usr/bin/env python
"""Copy the unmodified reference's Python sources into the git-ignored ``baseline/_ref/`` so that
``bench.py --impl reference`` (and the ``cpu_baseline`` leg) can time the reference's own classes on the GPU box's
host cores - /root/reference does not exist there, the snapshot of this repo does.  Nothing is edited; figures,
READMEs and the label encoder are skipped.  Called by ``__graft_entry__.build()`` when /root/reference is present."""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")


def install(src: str = SRC, dst: str = DST) -> int:
    if not os.path.isdir(src):
        print(f"install_reference: {src} not present (GPU box?) - keeping whatever is in {dst}")
        return 0
    n = 0
    for base, dirs, files in os.walk(src):
        dirs[:] = [d for d in dirs if d not in (".git", "figure", "__pycache__")]
        for f in files:
            if not f.endswith(".py"):
                continue
            rel = os.path.relpath(os.path.join(base, f), src)
            out = os.path.join(dst, rel)
            os.makedirs(os.path.dirname(out), exist_ok=True)
            shutil.copyfile(os.path.join(base, f), out)
            n += 1
    print(f"install_reference: {n} files -> {dst}")
    return n


if __name__ == "__main__":
    sys.exit(0 if install() >= 0 else 1)
