"""Import shim: `import fidm_b200` loads the package that lives in the (non-identifier)
directory `face-inpainting-diffusion-models_b200/` next to this file."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                        "face-inpainting-diffusion-models_b200")


def _load():
    spec = importlib.util.spec_from_file_location(
        "fidm_b200", os.path.join(_PKG_DIR, "__init__.py"),
        submodule_search_locations=[_PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["fidm_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


_load()
