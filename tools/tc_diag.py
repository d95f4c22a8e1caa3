"""Builds tuning variants of the tcgen05 layer kernel (loads / MMAs / epilogue knocked out one at
a time, 3-plane 6-term split) -- build here, run with SELFRECON_B200_LIB=<variant> on the GPU box."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from selfreconcode_b200 import build
for tag, defs in [("noepi", ["SR_TC_DBG_NOEPI"]), ("nomma", ["SR_TC_DBG_NOMMA"]), ("noload", ["SR_TC_DBG_NOLOAD"]),
                  ("nopfull", ["SR_TC_DBG_NOPFULL"]), ("noload_noepi", ["SR_TC_DBG_NOLOAD", "SR_TC_DBG_NOEPI"])]:
    print(tag, build.build_variant(tag, defs))
