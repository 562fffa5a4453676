"""TEST INFRASTRUCTURE ONLY -- import the UNMODIFIED reference modules in this container.

`/root/reference` exists only in the build container (never on the GPU box), so this helper
is used by `oracle/make_golden.py` and by the `not gpu` tests that cross-check the standalone
restatement (`oracle/odevit_oracle.py`) against the reference itself.  It puts the shim
directory (torchdiffeq restatement, empty `utils`, empty `turtle`) and the reference root on
`sys.path`, then imports `models.ode_transformer_gpt`, `models.macaron`, `models.time_emb`,
`models.utils` and (optionally) `loss_trainer`.

Nothing under odevit_b200/ may import this file.
"""
import importlib
import os
import sys

REFERENCE_ROOT = os.environ.get("ODEVIT_REFERENCE_ROOT", "/root/reference")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "ode_transformer_gpt.py"))


def import_reference(with_loss_trainer: bool = False):
    """Returns a dict of the reference's modules, imported from where they lie."""
    if not reference_available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    # shims first so `import utils` / `import torchdiffeq` / `from turtle import pd` resolve
    # to the stand-ins; the reference root after, so `models.*` resolves to the real files.
    for p in (REFERENCE_ROOT, _SHIMS):
        if p in sys.path:
            sys.path.remove(p)
    sys.path.insert(0, REFERENCE_ROOT)
    sys.path.insert(0, _SHIMS)
    for name in ("utils", "turtle", "torchdiffeq"):
        mod = sys.modules.get(name)
        if mod is not None and not getattr(mod, "__file__", "").startswith(_SHIMS):
            del sys.modules[name]
    mods = {
        "ode": importlib.import_module("models.ode_transformer_gpt"),
        "macaron": importlib.import_module("models.macaron"),
        "time_emb": importlib.import_module("models.time_emb"),
        "model_utils": importlib.import_module("models.utils"),
    }
    if with_loss_trainer:
        mods["loss_trainer"] = importlib.import_module("loss_trainer")
    return mods
