"""Glyph rasterisation and template-bank production (host side, input producer).

Restates, over FreeType through ctypes, the calls the reference makes into
font-kit 0.14.3 (feature `freetype`) and pathfinder_geometry 0.5.1 on the hot
path's input side:

  * ``Font.glyph_for_char / typographic_bounds / advance / metrics /
    raster_bounds / rasterize_glyph``   -- call sites ncc.rs:154-194,605-618,
    main.rs:49-51,98-106,127,136-144,176
  * ``offset_grid``                     -- ncc.rs:563-573 (x-major)
  * ``alphabet_box`` / ``font_box``     -- ncc.rs:588-629
  * ``render``                          -- ncc.rs:143-196
  * ``TemplateBank``                    -- the (glyph, subpixel shift) raster cache the
    north star adds: the reference re-rasterises every template for every page
    (ncc.rs:561,631); here the bank is rendered once and uploaded once.

PARITY UNPINNED for the rasters themselves: font-kit and pathfinder are crates.io
dependencies that are not under /root/reference and there is no Rust toolchain,
so the behaviour below follows their published source from memory (FT_Set_Char_Size
at 72 dpi, FT_Set_Transform with a 26.6 delta whose y is negated, FT_LOAD_NO_HINTING,
FT_RENDER_MODE_NORMAL, copy-blit at (bitmap_left, -bitmap_top) clipped to the canvas;
raster_bounds = typographic bounds * size/upem, y flipped, transformed, round_out).
The hot path's contract is taken at the byte boundary: whatever bytes this module
produces are fed identically to the oracle and to the CUDA path.

FreeType comes from Pillow's bundled libfreetype (2.14.3 in this image); there are
no FreeType headers or system library here.
"""
from __future__ import annotations

import ctypes as C
import glob
import math
import os
from dataclasses import dataclass

import numpy as np

f32 = np.float32

# ncc.rs:28-29 and main.rs:13-14
NCC_DEFAULT_ALPHABET = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789=+<>(){};:/-"
FOCR_DEFAULT_ALPHABET = "> =ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/"

_FONT_CANDIDATES = [
    # a monospace OTF/TTF is preferred when the box has one (BASELINE.json configs)
    "/usr/share/fonts/truetype/dejavu/DejaVuSansMono.ttf",
    "/usr/share/fonts/truetype/liberation/LiberationMono-Regular.ttf",
    "/usr/share/fonts/dejavu/DejaVuSansMono.ttf",
    # what this image actually has: proportional web fonts in doc trees
    "/usr/local/cuda-12.9/compute-sanitizer/docs/_static/css/fonts/lato-normal.woff",
    "/usr/local/cuda/compute-sanitizer/docs/_static/css/fonts/lato-normal.woff",
    "/usr/local/cuda-12.9/extras/CUPTI/doc/html/_static/css/fonts/lato-normal.woff",
]


def find_font() -> str:
    for p in _FONT_CANDIDATES:
        if os.path.exists(p):
            return p
    hits = sorted(glob.glob("/usr/local/cuda*/**/lato-normal.woff", recursive=True)) + sorted(
        glob.glob("/opt/nvidia/**/lato-normal.woff", recursive=True)
    )
    if hits:
        return hits[0]
    raise FileNotFoundError("no usable font found (looked for DejaVu/Liberation Mono and Lato)")


# --------------------------------------------------------------------------- FreeType binding
FT_LOAD_DEFAULT = 0
FT_LOAD_NO_HINTING = 1 << 1
FT_LOAD_RENDER = 1 << 2
FT_PIXEL_MODE_GRAY = 2


class _FT_Vector(C.Structure):
    _fields_ = [("x", C.c_long), ("y", C.c_long)]


class _FT_Matrix(C.Structure):
    _fields_ = [("xx", C.c_long), ("xy", C.c_long), ("yx", C.c_long), ("yy", C.c_long)]


class _FT_BBox(C.Structure):
    _fields_ = [("xMin", C.c_long), ("yMin", C.c_long), ("xMax", C.c_long), ("yMax", C.c_long)]


class _FT_Generic(C.Structure):
    _fields_ = [("data", C.c_void_p), ("finalizer", C.c_void_p)]


class _FT_Glyph_Metrics(C.Structure):
    _fields_ = [
        ("width", C.c_long), ("height", C.c_long),
        ("horiBearingX", C.c_long), ("horiBearingY", C.c_long), ("horiAdvance", C.c_long),
        ("vertBearingX", C.c_long), ("vertBearingY", C.c_long), ("vertAdvance", C.c_long),
    ]


class _FT_Bitmap(C.Structure):
    _fields_ = [
        ("rows", C.c_uint), ("width", C.c_uint), ("pitch", C.c_int),
        ("buffer", C.POINTER(C.c_ubyte)), ("num_grays", C.c_ushort),
        ("pixel_mode", C.c_ubyte), ("palette_mode", C.c_ubyte), ("palette", C.c_void_p),
    ]


class _FT_GlyphSlotRec(C.Structure):
    _fields_ = [
        ("library", C.c_void_p), ("face", C.c_void_p), ("next", C.c_void_p),
        ("glyph_index", C.c_uint), ("generic", _FT_Generic),
        ("metrics", _FT_Glyph_Metrics),
        ("linearHoriAdvance", C.c_long), ("linearVertAdvance", C.c_long),
        ("advance", _FT_Vector), ("format", C.c_int),
        ("bitmap", _FT_Bitmap), ("bitmap_left", C.c_int), ("bitmap_top", C.c_int),
    ]


class _FT_FaceRec(C.Structure):
    _fields_ = [
        ("num_faces", C.c_long), ("face_index", C.c_long), ("face_flags", C.c_long),
        ("style_flags", C.c_long), ("num_glyphs", C.c_long),
        ("family_name", C.c_char_p), ("style_name", C.c_char_p),
        ("num_fixed_sizes", C.c_int), ("available_sizes", C.c_void_p),
        ("num_charmaps", C.c_int), ("charmaps", C.c_void_p),
        ("generic", _FT_Generic), ("bbox", _FT_BBox),
        ("units_per_EM", C.c_ushort), ("ascender", C.c_short), ("descender", C.c_short),
        ("height", C.c_short), ("max_advance_width", C.c_short), ("max_advance_height", C.c_short),
        ("underline_position", C.c_short), ("underline_thickness", C.c_short),
        ("glyph", C.POINTER(_FT_GlyphSlotRec)),
    ]


_ft = None
_ft_lib = None


def _freetype():
    global _ft, _ft_lib
    if _ft is not None:
        return _ft, _ft_lib
    import PIL

    cands = glob.glob(os.path.join(os.path.dirname(PIL.__file__), "..", "pillow.libs", "libfreetype*.so*"))
    if not cands:
        raise OSError("Pillow's bundled libfreetype not found")
    import PIL._imagingft  # noqa: F401  (pulls in libfreetype's own dependencies)

    ft = C.CDLL(os.path.realpath(cands[0]))
    ft.FT_Init_FreeType.argtypes = [C.POINTER(C.c_void_p)]
    ft.FT_New_Face.argtypes = [C.c_void_p, C.c_char_p, C.c_long, C.POINTER(C.POINTER(_FT_FaceRec))]
    ft.FT_Set_Char_Size.argtypes = [C.POINTER(_FT_FaceRec), C.c_long, C.c_long, C.c_uint, C.c_uint]
    ft.FT_Set_Transform.argtypes = [C.POINTER(_FT_FaceRec), C.POINTER(_FT_Matrix), C.POINTER(_FT_Vector)]
    ft.FT_Load_Glyph.argtypes = [C.POINTER(_FT_FaceRec), C.c_uint, C.c_int32]
    ft.FT_Get_Char_Index.argtypes = [C.POINTER(_FT_FaceRec), C.c_ulong]
    ft.FT_Get_Char_Index.restype = C.c_uint
    lib = C.c_void_p()
    if ft.FT_Init_FreeType(C.byref(lib)) != 0:
        raise OSError("FT_Init_FreeType failed")
    _ft, _ft_lib = ft, lib
    return _ft, _ft_lib


# --------------------------------------------------------------------------- geometry (pathfinder)
@dataclass(frozen=True)
class RectF:
    """pathfinder_geometry::rect::RectF stored as origin + lower_right, f32 lanes."""

    x0: np.float32 = f32(0)
    y0: np.float32 = f32(0)
    x1: np.float32 = f32(0)
    y1: np.float32 = f32(0)

    @staticmethod
    def from_origin_size(ox, oy, w, h) -> "RectF":
        return RectF(f32(ox), f32(oy), f32(f32(ox) + f32(w)), f32(f32(oy) + f32(h)))

    def width(self):
        return f32(self.x1 - self.x0)

    def height(self):
        return f32(self.y1 - self.y0)

    def scale(self, s) -> "RectF":
        s = f32(s)
        return RectF(f32(self.x0 * s), f32(self.y0 * s), f32(self.x1 * s), f32(self.y1 * s))

    def translate(self, tx, ty) -> "RectF":
        tx, ty = f32(tx), f32(ty)
        return RectF(f32(self.x0 + tx), f32(self.y0 + ty), f32(self.x1 + tx), f32(self.y1 + ty))

    def union_rect(self, o: "RectF") -> "RectF":
        return RectF(min(self.x0, o.x0), min(self.y0, o.y0), max(self.x1, o.x1), max(self.y1, o.y1))

    def round_out_i32(self):
        """(x0, y0, x1, y1) as ints: floor the origin, ceil the lower-right."""
        return (int(math.floor(self.x0)), int(math.floor(self.y0)),
                int(math.ceil(self.x1)), int(math.ceil(self.y1)))


def _trunc_26_6(v) -> int:
    """font-kit `f32_to_ft_fixed_26_6`: `(x * 64.0) as i64` (Rust `as` truncates toward zero)."""
    return int(f32(f32(v) * f32(64.0)))


class Font:
    """The subset of font_kit::loaders::freetype::Font the hot path calls."""

    def __init__(self, path: str | None = None, face_index: int = 0, hinting: bool = False):
        """hinting: the reference's --hinting = HintingOptions::Full(size) (ncc.rs:547-551, main.rs:394-398): font-kit then
        rasterises with FT_LOAD_TARGET_NORMAL instead of FT_LOAD_NO_HINTING (metrics stay unhinted font units)."""
        ft, lib = _freetype()
        self._ft = ft
        self.hinting = bool(hinting)
        self.path = path or find_font()
        face = C.POINTER(_FT_FaceRec)()
        if ft.FT_New_Face(lib, self.path.encode(), face_index, C.byref(face)) != 0:
            raise OSError(f"FT_New_Face failed for {self.path}")
        self._face = face
        self.units_per_em = int(face.contents.units_per_EM)
        self.ascent = float(face.contents.ascender)
        self.descent = float(face.contents.descender)
        self.line_gap = float(face.contents.height - (face.contents.ascender - face.contents.descender))
        bb = face.contents.bbox
        self.bounding_box = RectF(f32(bb.xMin), f32(bb.yMin), f32(bb.xMax), f32(bb.yMax))
        self._reset_size()
        self._tb_cache: dict[int, RectF] = {}
        self._adv_cache: dict[int, tuple] = {}

    def _reset_size(self):
        # font-kit keeps the face at ppem == units_per_em so that 26.6 metrics / 64 = font units
        self._ft.FT_Set_Char_Size(self._face, self.units_per_em << 6, 0, 0, 0)

    def glyph_for_char(self, ch: str) -> int:
        gid = self._ft.FT_Get_Char_Index(self._face, ord(ch))
        if gid == 0:
            raise KeyError(f"no glyph for {ch!r}")  # the reference .unwrap()s a None here
        return gid

    def typographic_bounds(self, gid: int) -> RectF:
        if gid not in self._tb_cache:
            if self._ft.FT_Load_Glyph(self._face, gid, FT_LOAD_DEFAULT | FT_LOAD_NO_HINTING) != 0:
                raise OSError("FT_Load_Glyph failed")
            m = self._face.contents.glyph.contents.metrics
            if m.width == 0 or m.height == 0:
                r = RectF()
            else:
                r = RectF.from_origin_size(f32(m.horiBearingX) / f32(64), f32(m.horiBearingY - m.height) / f32(64),
                                           f32(m.width) / f32(64), f32(m.height) / f32(64))
            self._tb_cache[gid] = r
        return self._tb_cache[gid]

    def advance(self, gid: int):
        """font units, (x, y) as f32."""
        if gid not in self._adv_cache:
            if self._ft.FT_Load_Glyph(self._face, gid, FT_LOAD_DEFAULT | FT_LOAD_NO_HINTING) != 0:
                raise OSError("FT_Load_Glyph failed")
            a = self._face.contents.glyph.contents.advance
            self._adv_cache[gid] = (f32(a.x) / f32(64), f32(a.y) / f32(64))
        return self._adv_cache[gid]

    def raster_bounds(self, gid: int, point_size: float, tx: float = 0.0, ty: float = 0.0):
        """font-kit Loader::raster_bounds with a pure translation; returns int (x0,y0,x1,y1)."""
        tb = self.typographic_bounds(gid).scale(f32(point_size) / f32(self.units_per_em))
        flipped = RectF.from_origin_size(tb.x0, f32(-tb.y0) - tb.height(), tb.width(), tb.height())
        return flipped.translate(tx, ty).round_out_i32()

    def rasterize_glyph(self, canvas: np.ndarray, gid: int, point_size: float, tx: float, ty: float):
        """Rasterise into `canvas` (u8 [h, w], A8) at translation (tx, ty); copy-blit, clipped."""
        bmp, left, top = self.glyph_bitmap(gid, point_size, _trunc_26_6(tx), -_trunc_26_6(ty))
        _blit(canvas, bmp, left, -top)

    def glyph_bitmap(self, gid: int, point_size: float, dx_26_6: int, dy_26_6: int):
        """FT bitmap for a 26.6 pen delta: (u8 [rows, width], bitmap_left, bitmap_top)."""
        ft = self._ft
        ft.FT_Set_Char_Size(self._face, _trunc_26_6(point_size), 0, 0, 0)
        mat = _FT_Matrix(0x10000, 0, 0, 0x10000)
        delta = _FT_Vector(dx_26_6, dy_26_6)
        ft.FT_Set_Transform(self._face, C.byref(mat), C.byref(delta))
        try:
            if ft.FT_Load_Glyph(self._face, gid, FT_LOAD_DEFAULT | FT_LOAD_RENDER | (0 if self.hinting else FT_LOAD_NO_HINTING)) != 0:
                raise OSError("FT_Load_Glyph(render) failed")
            slot = self._face.contents.glyph.contents
            b = slot.bitmap
            if b.rows and b.width:
                if b.pixel_mode != FT_PIXEL_MODE_GRAY:
                    raise OSError(f"unexpected pixel mode {b.pixel_mode}")
                raw = np.ctypeslib.as_array(b.buffer, shape=(b.rows, abs(b.pitch)))
                bmp = np.array(raw[:, : b.width], dtype=np.uint8, copy=True)
            else:
                bmp = np.zeros((0, 0), np.uint8)
            return bmp, int(slot.bitmap_left), int(slot.bitmap_top)
        finally:
            ft.FT_Set_Transform(self._face, None, None)
            self._reset_size()


def _blit(canvas: np.ndarray, bmp: np.ndarray, dx: int, dy: int):
    """font-kit Canvas::blit_from for A8 -> A8: row memcpy of the overlap (overwrites)."""
    if bmp.size == 0:
        return
    h, w = canvas.shape
    bh, bw = bmp.shape
    x0, y0 = max(dx, 0), max(dy, 0)
    x1, y1 = min(dx + bw, w), min(dy + bh, h)
    if x1 <= x0 or y1 <= y0:
        return
    canvas[y0:y1, x0:x1] = bmp[y0 - dy : y1 - dy, x0 - dx : x1 - dx]


# --------------------------------------------------------------------------- ncc template producer
def offset_grid(x_bits: int, y_bits: int):
    """ncc.rs:563-573: x-major list of [x/2^xb, y/2^yb] as f32."""
    xd = f32(1.0) / f32(2 ** x_bits)
    yd = f32(1.0) / f32(2 ** y_bits)
    return [(f32(f32(x) * xd), f32(f32(y) * yd)) for x in range(2 ** x_bits) for y in range(2 ** y_bits)]


def alphabet_box(font: Font, alphabet: str, size: float, offset):
    """ncc.rs:600-626 BoxSize::Alphabet -> (y_offset f32, (w, h))."""
    to_px = f32(f32(1.0) / f32(font.units_per_em)) * f32(size)
    y_offset = f32(0.0)
    bounds = RectF()  # RectF::default(): the union is seeded with the point (0,0)
    for c in alphabet:
        gid = font.glyph_for_char(c)
        gb = font.typographic_bounds(gid).scale(to_px)
        bearing_y = f32(gb.y0 + gb.height())
        x0, y0, x1, y1 = font.raster_bounds(gid, size, offset[0], offset[1])
        y_offset = max(y_offset, f32(math.ceil(bearing_y)))
        bounds = bounds.union_rect(RectF(f32(x0), f32(y0), f32(x1), f32(y1)))
    bx0, by0, bx1, by1 = bounds.round_out_i32()
    return y_offset, (bx1 - bx0, by1 - by0)


def font_box(font: Font, size: float):
    """ncc.rs:589-599 BoxSize::Font."""
    to_px = f32(f32(1.0) / f32(font.units_per_em)) * f32(size)
    x0, y0, x1, y1 = font.bounding_box.scale(to_px).round_out_i32()
    return f32(math.ceil(f32(font.ascent) * to_px)), (x1 - x0, y1 - y0)


def render(font: Font, ch: str, offset, size: float, canvas_size=None, padding=(0, 0)) -> np.ndarray:
    """ncc.rs:143-196: one A8 template canvas (u8 [h, w])."""
    gid = font.glyph_for_char(ch)
    rb = font.raster_bounds(gid, size, offset[0], offset[1])
    if canvas_size is not None:
        w, h = canvas_size
        ox, oy = f32(0), f32(0)
    else:  # BoxSize::Char: tight box, origin = -raster_bounds.origin
        w, h = rb[2] - rb[0], rb[3] - rb[1]
        ox, oy = f32(-rb[0]), f32(-rb[1])
    w, h = w + 2 * padding[0], h + 2 * padding[1]
    canvas = np.zeros((h, w), np.uint8)
    tx = f32(f32(ox + f32(padding[0])) + f32(offset[0]))
    ty = f32(f32(oy + f32(padding[1])) + f32(offset[1]))
    font.rasterize_glyph(canvas, gid, size, tx, ty)
    return canvas


@dataclass
class Template:
    letter: str
    offset_index: int
    offset: tuple          # the (x, y) subpixel offset BEFORE the y_offset correction
    pixels: np.ndarray     # u8 [n_h, n_w], A8 coverage
    corrected_y: float = 0.0  # offset y + y_offset (ncc.rs:629), what --raw prints


class TemplateBank:
    """The (glyph, subpixel shift) raster cache: every template `get_hits` would render for one
    page (ncc.rs:587-640), in the reference's iteration order (offset index, alphabet index)."""

    def __init__(self, font: Font, size: float, alphabet: str = NCC_DEFAULT_ALPHABET, x_bits: int = 0,
                 y_bits: int = 0, box_size: str = "alphabet", padding=(0, 0)):
        self.font, self.size, self.alphabet = font, size, alphabet
        self.templates: list[Template] = []
        for oi, off in enumerate(offset_grid(x_bits, y_bits)):
            if box_size == "alphabet":
                y_off, csize = alphabet_box(font, alphabet, size, off)
            elif box_size == "font":
                y_off, csize = font_box(font, size)
            elif box_size == "char":
                y_off, csize = f32(0), None
            else:
                raise ValueError(box_size)  # the reference .unwrap()s the TryFrom error (ncc.rs:559)
            corrected = (off[0], f32(off[1] + y_off))  # ncc.rs:629
            for ch in alphabet:
                self.templates.append(Template(ch, oi, off, render(font, ch, corrected, size, csize, padding), corrected[1]))

    def __len__(self):
        return len(self.templates)

    def sizes(self):
        return sorted({t.pixels.shape[::-1] for t in self.templates})

    def letters(self):
        return [t.letter for t in self.templates]


# --------------------------------------------------------------------------- the C++ driver (host/focr_raster.cpp)
def freetype_library_path() -> str:
    """Path of the libfreetype the Python binding uses (Pillow's bundled one): what the C++ driver dlopens."""
    import PIL

    cands = glob.glob(os.path.join(os.path.dirname(PIL.__file__), "..", "pillow.libs", "libfreetype*.so*"))
    if not cands:
        raise OSError("Pillow's bundled libfreetype not found")
    import PIL._imagingft  # noqa: F401  (pulls in libfreetype's own dependencies)

    return os.path.realpath(cands[0])


class NativeFont:
    """focr_host_font: the C++ FreeType driver of libfocr_b200.so (same producers as this module, in C++)."""

    def __init__(self, path: str | None = None, hinting: bool = False):
        from . import native

        self.path = path or find_font()
        self._h = C.c_void_p()
        native.check(native.lib().focr_host_font_open(freetype_library_path().encode(), self.path.encode(), C.byref(self._h)))
        native.lib().focr_host_font_set_hinting(self._h, int(bool(hinting)))

    def template_bank(self, size: float, alphabet: str = NCC_DEFAULT_ALPHABET, x_bits: int = 0, y_bits: int = 0,
                      box_size: str = "alphabet", padding=(0, 0)):
        """ncc.rs:587-640 in C++: ([u8 [n_h, n_w]] in (offset, letter) order, letters, corrected y offsets)."""
        from . import native

        lib = native.lib()
        a = np.array([ord(c) for c in alphabet], np.uint32)
        h = C.c_void_p()
        native.check(lib.focr_host_tbank_render(self._h, C.c_float(size), native.ptr(a), len(a), x_bits, y_bits,
                                                {"alphabet": 0, "font": 1, "char": 2}[box_size], padding[0], padding[1], C.byref(h)))
        try:
            n = int(lib.focr_host_tbank_count(h))
            px = np.zeros(int(lib.focr_host_tbank_pixel_bytes(h)), np.uint8)
            off, nw, nh = np.zeros(n, np.uint64), np.zeros(n, np.uint16), np.zeros(n, np.uint16)
            let, cy = np.zeros(n, np.uint32), np.zeros(n, np.float32)
            native.check(lib.focr_host_tbank_get(h, native.ptr(px), native.ptr(off), native.ptr(nw), native.ptr(nh), native.ptr(let),
                                                 native.ptr(cy)))
        finally:
            lib.focr_host_tbank_free(h)
        tpls = [px[int(off[i]):int(off[i]) + int(nw[i]) * int(nh[i])].reshape(int(nh[i]), int(nw[i])).copy() for i in range(n)]
        return tpls, [chr(c) for c in let], cy

    def glyph_bank(self, size: float, alphabet: str = FOCR_DEFAULT_ALPHABET, kern_x: float = 1.0):
        """focr's raster cache in C++: (pixels u8, rasters [n, 64] focr_glyph_raster, advance_px f32 [n], (origin_x, origin_y))."""
        from . import native

        lib = native.lib()
        a = np.array([ord(c) for c in alphabet], np.uint32)
        h = C.c_void_p()
        native.check(lib.focr_host_gbank_render(self._h, C.c_float(size), native.ptr(a), len(a), C.c_float(kern_x), C.byref(h)))
        try:
            px = np.zeros(int(lib.focr_host_gbank_pixel_bytes(h)), np.uint8)
            ras = np.zeros((len(a), 64), native.RASTER_DTYPE)
            adv, org = np.zeros(len(a), np.float32), np.zeros(2, np.int32)
            native.check(lib.focr_host_gbank_get(h, native.ptr(px), native.ptr(ras), native.ptr(adv), native.ptr(org)))
        finally:
            lib.focr_host_gbank_free(h)
        return px, ras, adv, (int(org[0]), int(org[1]))

    def close(self):
        if self._h:
            from . import native

            native.lib().focr_host_font_close(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
