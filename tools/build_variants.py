#!/usr/bin/env python3
"""Build experiment variants of libacq_b200.so:  python tools/build_variants.py name=DEF1,DEF2 name2=...  (name= alone: no defines)"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flydog_sdr_gps_b200 import _build
d = os.path.join(_build.CSRC, "variants")
os.makedirs(d, exist_ok=True)
for a in sys.argv[1:]:
    name, _, defs = a.partition("=")
    print(_build.build(defines=[x for x in defs.split(",") if x], out=os.path.join(d, "libacq_b200_%s.so" % name)))
