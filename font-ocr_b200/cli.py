"""Flag-compatible front-ends for the reference's two binaries over the C ABI.

    python -m font_ocr_b200.cli ncc  --img p0.png p1.png --font F.ttf -t 13 --x-bits 2 [--csv | --raw]
    python -m font_ocr_b200.cli focr --img p0.png --font F.ttf -t 13 -x 45 -y 39 -w 608 --line-height 12 --line-advance 15

`ncc_main` follows ncc.rs:486-542 (flags) and ncc.rs:788-878 (text / --csv / --raw output); `focr_main` follows
main.rs:342-385 and main.rs:387-470.  Differences, all deliberate:
  * the template bank / glyph bank is rendered ONCE per run (the (glyph, shift) cache), not per page (ncc.rs:561,631);
  * pages of one run are scanned as batches on the GPU instead of one rayon task each; output order is by page index
    either way (ncc.rs:847, main.rs:468);
  * a page without any anchor line prints nothing, where the reference panics (ncc.rs:1040);
  * `--rust`, `--test`, `--verify` are refused: the scalar fallback (numerically different,
    SURVEY K9) and the diagnostics images are outside the hot path.
All compute happens in libfocr_b200.so; there is no CPU path here.
"""
from __future__ import annotations

import argparse
import sys

import numpy as np

from . import raster


def rust_f32(v) -> str:
    """`{}` of an f32 in Rust: shortest digits that round-trip, never an exponent, no trailing `.0`."""
    v = np.float32(v)
    if np.isnan(v):
        return "NaN"
    if np.isinf(v):
        return "inf" if v > 0 else "-inf"
    return np.format_float_positional(v, unique=True, trim="-")


def _narrow16(v: np.ndarray) -> np.ndarray:
    """The image crate's u16 -> u8 sample conversion (`FromPrimitive<u16> for u8`): (c + 128) / 257, not c >> 8."""
    return ((v.astype(np.uint32) + 128) // 257).astype(np.uint8)


def load_luma8(path: str) -> np.ndarray:
    """image::open(path).into_luma8() (ncc.rs:575, main.rs:429): u8 [h, w].  Colour images are reduced with the image
    crate's integer Rec.709 weights, (2126 r + 7152 g + 722 b) / 10000, alpha dropped; 16-bit samples are reduced in
    16 bits first and then narrowed with (c + 128) / 257 like the crate does."""
    from PIL import Image

    try:   # Pillow narrows 16-bit colour on load; OpenCV keeps the samples
        import cv2

        raw = cv2.imread(path, cv2.IMREAD_UNCHANGED)
        if raw is not None and raw.dtype == np.uint16:
            if raw.ndim == 2:
                return _narrow16(raw)
            if raw.shape[2] == 2:                      # gray + alpha
                return _narrow16(raw[..., 0])
            b, g, r = (raw[..., i].astype(np.uint64) for i in range(3))   # BGR[A]
            return _narrow16((2126 * r + 7152 * g + 722 * b) // 10000)
    except ImportError:
        pass
    im = Image.open(path)
    if im.mode in ("L", "1", "P", "I;16", "I", "F", "LA", "PA"):
        if im.mode in ("I;16", "I"):
            return _narrow16(np.asarray(im, np.uint32))
        if im.mode in ("P", "PA"):
            im = im.convert("RGB")
        else:
            return np.ascontiguousarray(np.asarray(im.convert("L"), np.uint8))
    rgb = np.asarray(im.convert("RGB"), np.uint32)
    return ((2126 * rgb[..., 0] + 7152 * rgb[..., 1] + 722 * rgb[..., 2]) // 10000).astype(np.uint8)


def _batches(images):
    """Group page indices by image size, keeping page order inside a group."""
    groups: dict[tuple, list[int]] = {}
    for i, im in enumerate(images):
        groups.setdefault(im.shape, []).append(i)
    return groups


def _ncc_parser():
    ap = argparse.ArgumentParser(prog="ncc", description="NCC template OCR on B200 (flags of ncc.rs:486-542)")
    ap.add_argument("-i", "--img", nargs="+", required=True)
    ap.add_argument("-f", "--font", required=True)
    ap.add_argument("-t", "--text-size", type=float, required=True)
    ap.add_argument("--x-bits", type=int, default=0)
    ap.add_argument("--y-bits", type=int, default=0)
    ap.add_argument("--hinting", action="store_true")
    ap.add_argument("--threshold", type=float, default=0.8)
    ap.add_argument("--anchor-threshold", type=float, default=0.95)
    ap.add_argument("--overlap", type=int, default=5)
    ap.add_argument("-a", "--alphabet", default=raster.NCC_DEFAULT_ALPHABET)
    ap.add_argument("--box-size", default="alphabet")
    ap.add_argument("--x-padding", type=int, default=0)
    ap.add_argument("--y-padding", type=int, default=0)
    ap.add_argument("--save-letters", action="store_true")
    ap.add_argument("--rust", action="store_true")
    ap.add_argument("-v", "--verbose", action="store_true")
    ap.add_argument("--csv", action="store_true")
    ap.add_argument("--raw", action="store_true")
    ap.add_argument("--device", type=int, default=0, help="(extension) CUDA device")
    ap.add_argument("--batch", type=int, default=16, help="(extension) pages per GPU batch")
    ap.add_argument("--spaces", action="store_true",
                    help="(extension, off by default) insert spaces where the gap between two kept hits exceeds the left glyph's "
                         "advance; the reference does not detect spaces (README.md:46)")
    ap.add_argument("--space-advance", type=float, default=None,
                    help="(extension) advance of a space in pixels for --spaces (default: the font's U+0020 advance at -t)")
    ap.add_argument("--max-matches", type=int, default=1024,
                    help="(extension) hits kept per template and page; the reference's MAX_MATCHES is 1024 (ncc.rs:31)")
    return ap


def ncc_raw_lines(bank: raster.TemplateBank, matches: np.ndarray, counts: np.ndarray):
    """--raw (ncc.rs:683-698): one line per hit in scan order (offset, letter, y, x):
    codepoint, centre x, centre y, x, y, w, h, bearing_x, corrected y offset, offset x, offset y."""
    font, size = bank.font, bank.size
    f32 = np.float32
    to_px = f32(f32(1.0) / f32(font.units_per_em)) * f32(size)
    out = []
    for t, tpl in enumerate(bank.templates):
        n_h, n_w = tpl.pixels.shape
        bearing_x = f32(font.typographic_bounds(font.glyph_for_char(tpl.letter)).x0 * to_px)
        tail = f"{n_w},{n_h},{rust_f32(bearing_x)},{rust_f32(tpl.corrected_y)},{rust_f32(tpl.offset[0])},{rust_f32(tpl.offset[1])}"
        for m in matches[t, : counts[t]]:
            x, y = int(m["x"]), int(m["y"])
            out.append(f"{ord(tpl.letter)},{rust_f32(f32(x) + f32(n_w) * f32(0.5))},{rust_f32(f32(y) + f32(n_h) * f32(0.5))},"
                       f"{x},{y},{tail}")
    return out


def ncc_main(argv=None, out=None) -> int:
    args = _ncc_parser().parse_args(argv)
    out = out or sys.stdout
    if args.rust:
        sys.stderr.write("ncc: --rust is not supported by the B200 path (DESIGN.md section 7)\n")
        return 2
    if args.raw and len(args.img) != 1:
        raise AssertionError("--raw takes exactly one image (ncc.rs:833-837)")
    import torch

    from . import ncc

    font = raster.Font(args.font, hinting=args.hinting)
    bank_h = raster.TemplateBank(font, args.text_size, args.alphabet, args.x_bits, args.y_bits, args.box_size,
                                 (args.x_padding, args.y_padding))
    if args.verbose:
        sys.stderr.write(f"templates {len(bank_h)} sizes {bank_h.sizes()}\n")
    if args.save_letters:  # ncc.rs:642-650
        import os

        from PIL import Image

        os.makedirs("letters", exist_ok=True)
        for t in bank_h.templates:
            x, y = int(t.offset[0] * np.float32(1000.0)), int(t.offset[1] * np.float32(1000.0))
            Image.fromarray(t.pixels).save(f"letters/{t.letter}-{x}_{y}.png")
    images = [load_luma8(p) for p in args.img]
    letters = bank_h.letters()
    tsize = [t.pixels.shape[::-1] for t in bank_h.templates]
    T, n_out = len(bank_h), args.max_matches
    ctx = ncc.Context(args.device)
    bank = ncc.Bank(ctx, [t.pixels for t in bank_h.templates])
    dev = torch.device("cuda", args.device)
    page_lines: dict[int, list] = {}
    raw_lines: list[str] = []
    try:
        for (r_h, r_w), idx in _batches(images).items():
            for b0 in range(0, len(idx), args.batch):
                chunk = idx[b0:b0 + args.batch]
                P = len(chunk)
                pages = torch.from_numpy(np.stack([images[i] for i in chunk])).to(dev)
                m_dev = torch.zeros(P * T * n_out * 8, dtype=torch.uint8, device=dev)
                c_dev = torch.zeros(P * T, dtype=torch.int32, device=dev)
                torch.cuda.synchronize(dev)
                ncc.scan_pages_device(ctx, bank, pages.data_ptr(), r_w * r_h, r_w, r_w, r_h, P, args.threshold, n_out,
                                      m_dev.data_ptr(), c_dev.data_ptr())
                if args.raw:
                    ctx.sync()
                    m = m_dev.cpu().numpy().view(ncc.MATCH_DTYPE).reshape(P, T, n_out)
                    c = c_dev.cpu().numpy().view(np.uint32).reshape(P, T)
                    raw_lines += ncc_raw_lines(bank_h, m[0], c[0])
                    continue
                lp, ls, st, sel = ncc.process_hits_device(ctx, m_dev.data_ptr(), c_dev.data_ptr(), T, n_out, P, letters,
                                                          args.anchor_threshold, args.overlap, raw=True)
                for l in range(len(lp)):
                    page_lines.setdefault(chunk[int(lp[l])], []).append((st[ls[l]:ls[l + 1]], sel[ls[l]:ls[l + 1]]))
    finally:
        bank.close()
        ctx.close()
    if args.raw:
        out.write("".join(l + "\n" for l in raw_lines))
        return 0
    f32 = np.float32
    adv_px, space_px = {}, 0.0
    if args.spaces:   # pen advances in pixels, f32 like main.rs:176-178
        upem = f32(font.units_per_em)
        px = lambda ch: float(f32(f32(font.advance(font.glyph_for_char(ch))[0] / upem) * f32(args.text_size)))
        adv_px = {ch: px(ch) for ch in args.alphabet}
        space_px = args.space_advance if args.space_advance is not None else px(" ")
    for i in range(len(images)):                    # pages.sort_by_key(|(i, _)| *i), ncc.rs:847
        for tpl, sel in page_lines.get(i, []):
            if args.csv:                            # ncc.rs:849-867
                for t, m in zip(tpl, sel):
                    w, h = tsize[int(t)]
                    x, y = int(m["x"]), int(m["y"])
                    out.write(f"{i},{ord(letters[int(t)])},{rust_f32(f32(x) + f32(w) * f32(0.5))},"
                              f"{rust_f32(f32(y) + f32(h) * f32(0.5))},{x},{y},{w},{h}\n")
            elif args.spaces:                       # extension: the same line with detected spaces
                line = [(letters[int(t)], int(m["x"]), int(m["y"]), m["similarity"]) for t, m in zip(tpl, sel)]
                out.write(ncc.lines_to_text_with_spaces([line], adv_px, space_px)[0] + "\n")
            else:                                   # ncc.rs:869-876
                out.write("".join(letters[int(t)] for t in tpl) + "\n")
    return 0


def _focr_parser():
    ap = argparse.ArgumentParser(prog="focr", description="least-squares glyph OCR on B200 (flags of main.rs:342-385)")
    ap.add_argument("-i", "--img", nargs="+", required=True)
    ap.add_argument("-f", "--font", required=True)
    ap.add_argument("-a", "--alphabet", default=raster.FOCR_DEFAULT_ALPHABET)
    ap.add_argument("--hinting", action="store_true")
    ap.add_argument("-t", "--text-size", type=float, required=True)
    ap.add_argument("-k", "--kerning", type=float, default=1.0)
    ap.add_argument("-x", type=int, default=0)
    ap.add_argument("-y", type=int, default=0)
    ap.add_argument("-w", "--width", type=int, required=True)
    ap.add_argument("--line-height", type=int, required=True)
    ap.add_argument("--line-advance", type=int, required=True)
    ap.add_argument("--test", default=None)
    ap.add_argument("--verify", default=None)
    ap.add_argument("--device", type=int, default=0, help="(extension) CUDA device")
    ap.add_argument("--batch", type=int, default=16, help="(extension) pages per GPU batch")
    return ap


def focr_main(argv=None, out=None) -> int:
    args = _focr_parser().parse_args(argv)
    out = out or sys.stdout
    if args.test or args.verify:
        sys.stderr.write("focr: --test / --verify are not supported by the B200 path (DESIGN.md section 7)\n")
        return 2
    from . import focr, ncc

    font = raster.Font(args.font, hinting=args.hinting)
    images = [load_luma8(p) for p in args.img]
    ctx = ncc.Context(args.device)
    bank = focr.GlyphBank(ctx, font, args.text_size, args.alphabet, args.kerning)
    texts: dict[int, list] = {}
    try:
        for _, idx in _batches(images).items():
            for b0 in range(0, len(idx), args.batch):
                chunk = idx[b0:b0 + args.batch]
                res = focr.decode_images(ctx, bank, np.stack([images[i] for i in chunk]), args.x, args.y, args.width,
                                         args.line_height, args.line_advance)
                for i, lines in zip(chunk, res):
                    texts[i] = lines
    finally:
        bank.close()
        ctx.close()
    for i in range(len(images)):                    # liness.sort_by_key(|(i, _)| *i), main.rs:468-471
        for text, _y in texts[i]:
            out.write(text + "\n")
    return 0


def main(argv=None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv or argv[0] not in ("ncc", "focr"):
        sys.stderr.write("usage: python -m font_ocr_b200.cli {ncc|focr} <flags of the reference binary>\n")
        return 2
    return ncc_main(argv[1:]) if argv[0] == "ncc" else focr_main(argv[1:])


if __name__ == "__main__":
    sys.exit(main())
