"""font-ocr hot path, B200-native: NCC template scan + focr least-squares glyph search.

Layout (DESIGN.md has the full map):
  csrc/      hand-written sm_100a CUDA kernels + the C ABI (include/focr_b200.h) -> libfocr_b200.so
  host/      C++ mirror of the reference's host-side operators (Searcher, process_hits, decode_line)
  ncc.py     ctypes binding + Python mirror of the reference interface for the NCC path
  focr.py    same for the focr path
  shard.py   page sharding across GPUs + host-side gather (no data-path collective)
  raster.py  FreeType template/glyph-bank producer (the (glyph, subpixel shift) raster cache)
  pages.py   synthetic page generator for the BASELINE configs
  cli.py     flag-compatible `ncc` / `focr` front-ends (text, --csv, --raw)

There is NO CPU fallback: anything that computes calls into libfocr_b200.so and raises if the
library is missing.  The CPU restatement lives in oracle/ and is test infrastructure only.
"""
from . import raster, pages  # noqa: F401  (pure host-side; importable without the CUDA library)

__all__ = ["raster", "pages", "shard", "ncc", "focr", "native"]


def __getattr__(name):
    if name in ("ncc", "focr", "native", "shard"):
        import importlib

        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
