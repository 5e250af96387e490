"""CPU ORACLE support (test infrastructure only): import the reference's OWN implementation of the path.

`oracle/_ref/` holds an unmodified copy of the five reference files the path lives in (`nn.py`, `unet.py`,
`gaussian_diffusion.py`, `losses.py`, `utils/schedules.py`), made by `__graft_entry__.build_reference_copy()` in the
build container; it is git-ignored and travels to the GPU box with the snapshot.  `load()` imports those modules
(top-level names, as the reference imports them itself: `from nn import ...`, unet.py:8) and returns them, or None
when no copy exists -- callers then fall back to the restatement in `oracle/*.py` and say so (`kind = "port"`).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU legs may use this.
"""
import importlib
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
_CANDIDATES = (os.path.join(_HERE, "_ref"), "/root/reference/code")
_cache = None


def available():
    return any(os.path.exists(os.path.join(c, "gaussian_diffusion.py")) for c in _CANDIDATES)


def load():
    """SimpleNamespace(UNetModel, DiffusionInpaintingModel, GaussianDiffusion, create_gaussian_diffusion,
    get_named_beta_schedule, path) of the reference, or None."""
    global _cache
    if _cache is not None:
        return _cache
    root = next((c for c in _CANDIDATES if os.path.exists(os.path.join(c, "gaussian_diffusion.py"))), None)
    if root is None:
        return None
    sys.dont_write_bytecode = True
    saved = {k: sys.modules.get(k) for k in ("nn", "unet", "losses", "gaussian_diffusion", "utils", "utils.schedules")}
    sys.path.insert(0, root)
    try:
        for k in saved:
            sys.modules.pop(k, None)
        mods = {k: importlib.import_module(k) for k in ("nn", "losses", "unet", "gaussian_diffusion", "utils.schedules")}
    finally:
        sys.path.remove(root)
        # keep the reference's modules importable by name for its own lazy imports (schedules.py:84 imports
        # gaussian_diffusion inside the factory); restore anything we displaced that was not the reference's
        for k, v in saved.items():
            if v is not None and getattr(v, "__file__", "") and not str(v.__file__).startswith(root):
                sys.modules[k] = v
    # create_gaussian_diffusion does `from gaussian_diffusion import GaussianDiffusion` at call time
    sys.modules.setdefault("gaussian_diffusion", mods["gaussian_diffusion"])
    sys.modules.setdefault("losses", mods["losses"])
    _cache = types.SimpleNamespace(
        UNetModel=mods["unet"].UNetModel, DiffusionInpaintingModel=mods["unet"].DiffusionInpaintingModel,
        GaussianDiffusion=mods["gaussian_diffusion"].GaussianDiffusion,
        create_gaussian_diffusion=mods["utils.schedules"].create_gaussian_diffusion,
        get_named_beta_schedule=mods["utils.schedules"].get_named_beta_schedule, path=root)
    return _cache


def build_model(cfg, state_dict):
    """The reference's DiffusionInpaintingModel(UNetModel(**cfg with in_channels=3)) carrying `state_dict`."""
    ref = load()
    model = ref.DiffusionInpaintingModel(ref.UNetModel(**dict(cfg, in_channels=3)), in_channels=9).eval()
    model.load_state_dict(state_dict, strict=True)
    return model
