#!/usr/bin/env python
"""Delete stale artefacts from mpc-code_b200/_build and oracle/_build: everything whose name does not carry the digest of
a library / harness / oracle module that the CURRENT sources produce (`__graft_entry__.build()` is run first, so the
current ones exist).  Keeps the gpurun snapshot small.  Also removes what only the tests build (`ref_*` libraries of the
reference's example files, temporary examples, harnesses): the next `pytest -m "not gpu"` run rebuilds those, which takes
minutes - so run the CPU suite once after pruning.   python tools/prune_build.py [--dry-run]"""
import os
import re
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402


def main():
    dry = "--dry-run" in sys.argv
    import io
    import contextlib
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        entry.build()
    keep = set(re.findall(r"_([0-9a-f]{16})", buf.getvalue()))
    freed = 0
    for d in (os.path.join(ROOT, "mpc-code_b200", "_build"), os.path.join(ROOT, "oracle", "_build")):
        for name in sorted(os.listdir(d)):
            digests = re.findall(r"_([0-9a-f]{16})", name)
            if not digests or any(g in keep for g in digests):
                continue
            path = os.path.join(d, name)
            size = sum(os.path.getsize(os.path.join(r, f)) for r, _, fs in os.walk(path) for f in fs) if os.path.isdir(path) \
                else os.path.getsize(path)
            freed += size
            if not dry:
                shutil.rmtree(path) if os.path.isdir(path) else os.remove(path)
    print("%s %.1f MB; %d current digests kept" % ("would free" if dry else "freed", freed / 1e6, len(keep)))


if __name__ == "__main__":
    main()
