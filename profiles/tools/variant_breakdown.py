"""mcts_breakdown.py against an alternative build of the library: python variant_breakdown.py <lib.so> <n> <sims>"""
import os, sys, runpy
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from nypc_yacht_auction_b200 import _lib
_lib.LIB_PATH = os.path.abspath(sys.argv[1])
sys.argv = [sys.argv[0]] + sys.argv[2:]
runpy.run_path(os.path.join(ROOT, "profiles", "tools", "mcts_breakdown.py"), run_name="__main__")
