#!/usr/bin/env python
"""Development aid: recompile only the named translation units (substring match), relink the library and
refresh the source-hash stamp.  usage: rebuild.py sde_mma ccvm_abi"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g  # noqa: E402

t0 = time.time()
g.compile_library(only=tuple(sys.argv[1:]) or None)
with open(os.path.join(g.BUILD_DIR, "libccvm_b200.sha256"), "w") as fh:
    fh.write(g._source_hash())
print(f"rebuilt {sys.argv[1:]} in {time.time() - t0:.0f} s")
