#!/usr/bin/env python3
"""Build experimental variants of the library into tools/_build/lib_<name>.so:  name=DEF1,DEF2=val ...
   python tools/build_variants.py p0=LA_POLY_MASK=0x00u p2=LA_POLY_MASK=0x90u"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "liteattention_b200", "csrc"))
import build as B
os.makedirs(os.path.join(ROOT, "tools", "_build"), exist_ok=True)
for spec in sys.argv[1:]:
    name, _, defs = spec.partition("=")
    out = os.path.join(ROOT, "tools", "_build", f"lib_{name}.so")
    B.build(force=True, out=out, defines=[d for d in defs.split(",") if d])
    print(out)
