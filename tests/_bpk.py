"""Import helper: the package directory name contains hyphens, so it is imported by string."""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

bpk = importlib.import_module("baby-plonk-rust_b200")
