"""Import shim: the package directory is named ``pixlzr-rust_b200`` (not a valid Python
identifier), so ``import pixlzr_b200`` loads it from there under this name."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pixlzr-rust_b200")
_spec = importlib.util.spec_from_file_location(
    "pixlzr_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["pixlzr_b200"] = _mod
_spec.loader.exec_module(_mod)

if __name__ == "__main__":  # `python pixlzr_b200.py -i ... -o ...`: the reference CLI's arguments (src/bin/main.rs)
    sys.exit(_mod.cli.main())
