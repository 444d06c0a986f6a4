"""Import shim.  The package directory is `font-ocr_b200/` (the name the layout contract gives it);
a hyphen cannot be imported by name, so `import font_ocr_b200` lands here and this file swaps
itself for the real package."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "font-ocr_b200")
_spec = importlib.util.spec_from_file_location(
    "font_ocr_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir]
)
_mod = importlib.util.module_from_spec(_spec)
sys.modules["font_ocr_b200"] = _mod
_spec.loader.exec_module(_mod)
