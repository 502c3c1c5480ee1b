"""Registers the hyphen-named package directory `videollm-online_b200/` as the importable module
`videollm_online_b200` (a hyphen is not legal in a Python identifier).  `import vlo_bootstrap` first."""
import importlib.util
import pathlib
import sys

_ROOT = pathlib.Path(__file__).resolve().parent
_PKG = _ROOT / "videollm-online_b200"
_NAME = "videollm_online_b200"

if _NAME not in sys.modules:
    _spec = importlib.util.spec_from_file_location(_NAME, _PKG / "__init__.py", submodule_search_locations=[str(_PKG)])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_NAME] = _mod
    _spec.loader.exec_module(_mod)
if str(_ROOT) not in sys.path:
    sys.path.insert(0, str(_ROOT))
