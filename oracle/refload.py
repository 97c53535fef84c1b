"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Imports the UNMODIFIED reference (`/root/reference`) on CPU so that it can be used
as the parity oracle and to generate golden vectors.

The reference's `models/__init__.py` pulls in the whole viz/IO stack through
`utils.py:3-13` (laspy, plotly, dash, open3d, pykeops) and the PAConv tree, whose
`lib/pointops/functions/pointops.py:7-44` tries a broken JIT build of
`pointops_cuda`.  None of those are needed by the hot path, so they are replaced
by empty stub modules *before* `import models`.

`/root/reference` exists only in the authoring container; everything that must run
on the GPU box uses `oracle/port.py` + the committed fixtures in `tests/golden/`.
"""
import os
import sys
import types
import warnings

REFERENCE_ROOT = os.environ.get("FLOWCOMPARE_REFERENCE", "/root/reference")

_STUBS = {
    "laspy": {},
    "laspy.file": {"File": object},
    "plotly": {},
    "plotly.graph_objects": {},
    "dash": {},
    "dash_core_components": {},
    "dash_html_components": {},
    "open3d": {},
    "pykeops": {},
    "pykeops.torch": {"Vi": object, "Vj": object},
    "pointops_cuda": {},
    "wandb": {},
}


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "models"))


def load():
    """Returns (models_module, model_initialization_module) of the reference."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    for name, attrs in _STUBS.items():
        if name in sys.modules:
            continue
        try:
            __import__(name)
            continue
        except Exception:
            pass
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        if "." in name:
            parent, child = name.rsplit(".", 1)
            setattr(sys.modules[parent], child, m)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    warnings.filterwarnings("ignore", message=".*use_reentrant.*")
    warnings.filterwarnings("ignore", message=".*requires_grad=True.*")
    import models  # noqa: F401  (the reference's package)
    import model_initialization
    return models, model_initialization


def load_yaml_config(name: str) -> dict:
    """Reference `config/<name>.yaml` flattened like `utils.config_loader` (utils.py:373-377)."""
    import yaml
    with open(os.path.join(REFERENCE_ROOT, "config", f"{name}.yaml")) as f:
        raw = yaml.safe_load(f)
    return {k: v["value"] for k, v in raw.items() if isinstance(v, dict) and "value" in v}


def load_test_flow():
    """The reference's evaluation script `test_flow.py` as a module (for `log_prob_to_change` / `clamp_infs`,
    test_flow.py:241-275).  Its data-set and Dash imports (`dataloaders/__init__.py:3` imports a file that does not
    exist; `visualize_change_map.py` needs dash) are replaced by empty stubs; nothing of the functions under test
    touches them."""
    load()
    stubs = {"dataloaders": {"ChallengeDataset": object, "AmsVoxelLoader": object, "FullSceneLoader": object},
             "visualize_change_map": {"visualize_change": None}}
    for name, attrs in stubs.items():
        if name not in sys.modules:
            m = types.ModuleType(name)
            for k, v in attrs.items():
                setattr(m, k, v)
            sys.modules[name] = m
    import test_flow
    return test_flow
