#!/usr/bin/env python
"""Build the current tree into build/lib_<name>.so with extra nvcc flags (A/B of library variants in one GPU call: bench.py and the
tests load $FLOCK_LIB when it is set).   usage: python tools/mkvariant.py <name> [-DFLAG=1 ...]"""
import os, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gym_flock_b200 import _lib

os.makedirs(os.path.join(ROOT, "build"), exist_ok=True)
print(_lib.build(force=True, extra_flags=sys.argv[2:], out=os.path.join(ROOT, "build", f"lib_{sys.argv[1]}.so")))
