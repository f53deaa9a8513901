"""Import shim: the package directory is named ``rl-6-nimmt_b200`` (not a Python identifier);
``import rl_6_nimmt_b200`` resolves to it."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("rl-6-nimmt_b200")
sys.modules[__name__] = _pkg
# make `import rl_6_nimmt_b200.env` style submodule imports resolve as well
for _name, _mod in list(sys.modules.items()):
    if _name.startswith("rl-6-nimmt_b200."):
        sys.modules["rl_6_nimmt_b200." + _name.split(".", 1)[1]] = _mod
