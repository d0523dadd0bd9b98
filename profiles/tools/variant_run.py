"""Runs another tool against an alternative build of the library: python variant_run.py <lib.so> <tool.py> [args]"""
import os, sys, runpy
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from nypc_yacht_auction_b200 import _lib
_lib.LIB_PATH = os.path.abspath(sys.argv[1])
tool = sys.argv[2]
sys.argv = [tool] + sys.argv[3:]
runpy.run_path(tool, run_name="__main__")
