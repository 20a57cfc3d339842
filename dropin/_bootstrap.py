"""Makes `lr2ppo_b200` importable from a checkout (dropin/ sits next to the package) and, on images without h5py,
puts the read-only shim on sys.path.  Imported first by every drop-in module."""
import importlib.util
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
for p in (_ROOT, _HERE, os.path.join(_HERE, "finetune")):
    if p not in sys.path:
        sys.path.insert(0, p)
if importlib.util.find_spec("h5py") is None:
    sys.path.append(os.path.join(_HERE, "_shims"))          # appended: a real h5py always wins
