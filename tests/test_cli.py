"""The flag-compatible front-ends (font-ocr_b200/cli.py; ncc.rs:486-542,788-878, main.rs:342-470): formatting and
image loading on the CPU, and on the GPU the printed text / --csv / --raw against the oracle's pipeline."""
import io

import numpy as np
import pytest


def test_rust_f32_display(pkg):
    from font_ocr_b200.cli import rust_f32

    assert [rust_f32(v) for v in (0.0, 7.0, 7.5, 0.25, 1234.5, -3.0, 0.1, 1e-7, 16777216.0)] == \
           ["0", "7", "7.5", "0.25", "1234.5", "-3", "0.1", "0.0000001", "16777216"]


def test_load_luma8_gray_and_rec709(pkg, tmp_path):
    from PIL import Image

    from font_ocr_b200.cli import load_luma8

    rng = np.random.default_rng(3)
    g = rng.integers(0, 256, (9, 13), dtype=np.uint8)
    Image.fromarray(g).save(tmp_path / "g.png")
    assert np.array_equal(load_luma8(str(tmp_path / "g.png")), g)
    Image.fromarray(g).save(tmp_path / "g.pgm")
    assert np.array_equal(load_luma8(str(tmp_path / "g.pgm")), g)
    rgb = rng.integers(0, 256, (9, 13, 3), dtype=np.uint8)
    Image.fromarray(rgb).save(tmp_path / "c.png")
    r32 = rgb.astype(np.uint32)
    exp = ((2126 * r32[..., 0] + 7152 * r32[..., 1] + 722 * r32[..., 2]) // 10000).astype(np.uint8)
    assert np.array_equal(load_luma8(str(tmp_path / "c.png")), exp)
    # 16-bit gray: the image crate narrows with (c + 128) / 257 (129 -> 1, 127 -> 0), not with >> 8
    g16 = rng.integers(0, 65536, (9, 13), dtype=np.uint16)
    g16[0, :4] = (127, 128, 129, 65535)
    Image.fromarray(g16).save(tmp_path / "g16.png")
    got = load_luma8(str(tmp_path / "g16.png"))
    assert np.array_equal(got, ((g16.astype(np.uint32) + 128) // 257).astype(np.uint8))
    assert got[0, :4].tolist() == [0, 0, 1, 255]


def _cpp_cli(built_lib):
    import os

    return os.path.join(os.path.dirname(built_lib), "bin", "focr_cli")


def test_cpp_cli_decodes_images_like_the_python_front_end(built_lib, pkg, tmp_path):
    """host/focr_cli.cpp decodes PNG (through zlib) and PNM itself: `into_luma8()` must give the same bytes as cli.load_luma8
    for gray / RGB / RGBA / palette / 16-bit PNGs and binary PGM / PPM."""
    import io
    import subprocess

    from PIL import Image

    from font_ocr_b200.cli import load_luma8

    rng = np.random.default_rng(8)
    g = rng.integers(0, 256, (37, 53), dtype=np.uint8)
    rgb = rng.integers(0, 256, (37, 53, 3), dtype=np.uint8)
    rgba = rng.integers(0, 256, (37, 53, 4), dtype=np.uint8)
    g16 = rng.integers(0, 65536, (37, 53), dtype=np.uint16)
    files = {}
    Image.fromarray(g).save(tmp_path / "g.png"); files["g.png"] = None
    Image.fromarray(rgb).save(tmp_path / "c.png"); files["c.png"] = None
    Image.fromarray(rgba).save(tmp_path / "a.png"); files["a.png"] = None
    Image.fromarray(g16).save(tmp_path / "g16.png"); files["g16.png"] = None
    Image.fromarray(rgb).convert("P", palette=Image.ADAPTIVE, colors=64).save(tmp_path / "p.png"); files["p.png"] = None
    Image.fromarray((g > 127).astype(np.uint8) * 255).convert("1").save(tmp_path / "b.png"); files["b.png"] = None
    Image.fromarray(g).save(tmp_path / "g.pgm"); files["g.pgm"] = None
    Image.fromarray(rgb).save(tmp_path / "c.ppm"); files["c.ppm"] = None
    for name in files:
        out = subprocess.run([_cpp_cli(built_lib), "luma", str(tmp_path / name)], capture_output=True, check=True).stdout
        got = np.asarray(Image.open(io.BytesIO(out)))
        exp = load_luma8(str(tmp_path / name))
        assert got.shape == exp.shape and np.array_equal(got, exp), name


def test_flags_mirror_the_reference(pkg):
    from font_ocr_b200 import cli

    a = cli._ncc_parser().parse_args(["-i", "a.png", "b.png", "-f", "F", "-t", "13"])
    assert (a.x_bits, a.y_bits, a.threshold, a.anchor_threshold, a.overlap, a.box_size, a.x_padding, a.csv, a.raw) == \
           (0, 0, 0.8, 0.95, 5, "alphabet", 0, False, False)
    assert a.alphabet == pkg.raster.NCC_DEFAULT_ALPHABET and a.img == ["a.png", "b.png"]
    assert a.max_matches == 1024   # ncc.rs:31 unless the caller lifts it explicitly
    f = cli._focr_parser().parse_args(["-i", "a.png", "-f", "F", "-t", "13", "-w", "608", "--line-height", "12",
                                       "--line-advance", "15"])
    assert (f.x, f.y, f.kerning, f.width, f.line_height, f.line_advance) == (0, 0, 1.0, 608, 12, 15)
    assert f.alphabet == pkg.raster.FOCR_DEFAULT_ALPHABET
    assert cli.ncc_main(["-i", "a.png", "-f", "F", "-t", "13", "--rust"]) == 2       # refused, not silently different
    assert cli.focr_main(["-i", "a.png", "-f", "F", "-t", "13", "-w", "1", "--line-height", "1", "--line-advance", "1",
                          "--verify", "d"]) == 2
    assert cli.main(["nope"]) == 2


@pytest.mark.gpu
def test_ncc_cli_text_csv_raw(built_lib, oracle, font, pkg, tmp_path):
    """Two pages of different sizes through `ncc`: printed lines, --csv rows and --raw rows equal what the oracle's
    search + process_hits give for the same page and template bytes."""
    from PIL import Image

    from font_ocr_b200 import cli

    bank_h = pkg.raster.TemplateBank(font, 13, x_bits=1)
    letters = bank_h.letters()
    pages = [pkg.pages.make_ncc_page(bank_h, 608, 300, seed=11, shifts="bank")[0],
             pkg.pages.make_ncc_page(bank_h, 500, 260, seed=12, shifts="bank")[0]]
    paths = []
    for i, p in enumerate(pages):
        paths.append(str(tmp_path / f"p{i}.png"))
        Image.fromarray(p).save(paths[-1])
    base = ["-f", font.path, "-t", "13", "--x-bits", "1"]
    exp_lines = []
    per_tpl0 = None
    for i, p in enumerate(pages):
        per_tpl = oracle.get_hits(p, [t.pixels for t in bank_h.templates], 0.8)
        per_tpl0 = per_tpl if i == 0 else per_tpl0
        # the "letter" slot carries the template index (process_hits treats it as opaque), so rows know their box size
        exp_lines.append(oracle.process_hits(oracle.hits_with_letters(per_tpl, range(len(letters))), 0.95, 5))

    buf = io.StringIO()
    assert cli.ncc_main(["-i"] + paths + base, out=buf) == 0
    exp_text = ["".join(letters[hh[0]] for hh in line) for lines in exp_lines for line in lines]
    assert buf.getvalue().splitlines() == exp_text and len(exp_text) > 4

    buf = io.StringIO()
    assert cli.ncc_main(["-i"] + paths + base + ["--csv"], out=buf) == 0
    wh = [t.pixels.shape[::-1] for t in bank_h.templates]
    exp_csv = [f"{i},{ord(letters[hh[0]])},{cli.rust_f32(hh[1] + wh[hh[0]][0] / 2)},{cli.rust_f32(hh[2] + wh[hh[0]][1] / 2)},"
               f"{hh[1]},{hh[2]},{wh[hh[0]][0]},{wh[hh[0]][1]}" for i, lines in enumerate(exp_lines) for line in lines for hh in line]
    assert buf.getvalue().splitlines() == exp_csv and len({r.split(",")[6] for r in exp_csv}) == 2   # both box widths occur

    buf = io.StringIO()
    assert cli.ncc_main(["-i", paths[0]] + base + ["--raw"], out=buf) == 0
    rows = [r.split(",") for r in buf.getvalue().splitlines()]
    exp_raw = [(ord(letters[t]), int(m["x"]), int(m["y"])) for t, ms in enumerate(per_tpl0) for m in ms]
    assert [(int(r[0]), int(r[3]), int(r[4])) for r in rows] == exp_raw
    exp_wh = [wh[t] for t, ms in enumerate(per_tpl0) for _ in ms]
    assert all(len(r) == 11 for r in rows) and [(int(r[5]), int(r[6])) for r in rows] == exp_wh
    assert {r[9] for r in rows} == {"0", "0.5"} and {r[10] for r in rows} == {"0"}   # offsets of --x-bits 1


@pytest.mark.gpu
def test_focr_cli(built_lib, oracle, font, pkg, tmp_path):
    from PIL import Image

    from font_ocr_b200 import cli

    paths, exp = [], []
    for i, seed in enumerate((21, 22, 23)):
        page, _ = pkg.pages.make_focr_page(font, 13, 700, 39 + 15 * 4 + 20, seed=seed, fill=1.0)
        paths.append(str(tmp_path / f"f{i}.png"))
        Image.fromarray(page).save(paths[-1])
        exp += [t for t, _ in oracle.decode_image(page, font, pkg.raster.FOCR_DEFAULT_ALPHABET, 13, 45, 39, 608, 12, 15)]
    buf = io.StringIO()
    assert cli.focr_main(["-i"] + paths + ["-f", font.path, "-t", "13", "-x", "45", "-y", "39", "-w", "608",
                                            "--line-height", "12", "--line-advance", "15", "--batch", "2"], out=buf) == 0
    assert buf.getvalue().splitlines() == exp and len(exp) >= 9


@pytest.mark.gpu
def test_cpp_cli_prints_what_the_python_front_end_prints(built_lib, font, pkg, tmp_path):
    """host/focr_cli.cpp (C++ FreeType driver, PNG/PNM decoding, C++ process_hits) against font-ocr_b200/cli.py (Python
    producers, device process_hits): identical stdout for `ncc` text / --csv / --raw / --spaces and for `focr`."""
    import subprocess

    from PIL import Image

    from font_ocr_b200 import cli

    exe, ft = _cpp_cli(built_lib), pkg.raster.freetype_library_path()
    bank_h = pkg.raster.TemplateBank(font, 13, x_bits=1)
    paths = []
    for i, (w, h, seed) in enumerate(((608, 300, 11), (500, 260, 12), (608, 300, 13))):
        paths.append(str(tmp_path / (f"p{i}.png" if i != 1 else "p1.pgm")))
        Image.fromarray(pkg.pages.make_ncc_page(bank_h, w, h, seed=seed, shifts="bank")[0]).save(paths[-1])
    base = ["-f", font.path, "-t", "13", "--x-bits", "1"]

    def both(sub, args):
        buf = io.StringIO()
        assert (cli.ncc_main if sub == "ncc" else cli.focr_main)(args, out=buf) == 0
        r = subprocess.run([exe, sub] + args + ["--freetype", ft], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        return buf.getvalue(), r.stdout

    for extra in ([], ["--csv"], ["--spaces"], ["--hinting"], ["--threshold", "0.7", "--anchor-threshold", "0.9", "--overlap", "3"]):
        py, cpp = both("ncc", ["-i"] + paths + base + extra)
        assert py == cpp and (len(py.splitlines()) > 4 or extra == ["--hinting"]), extra   # (the pages were rendered unhinted)
    py, cpp = both("ncc", ["-i", paths[0]] + base + ["--raw"])
    assert py == cpp and len(py.splitlines()) > 100
    fpaths = []
    for i, seed in enumerate((21, 22)):
        fpaths.append(str(tmp_path / f"f{i}.png"))
        Image.fromarray(pkg.pages.make_focr_page(font, 13, 700, 39 + 15 * 4 + 20, seed=seed, fill=1.0)[0]).save(fpaths[-1])
    py, cpp = both("focr", ["-i"] + fpaths + ["-f", font.path, "-t", "13", "-x", "45", "-y", "39", "-w", "608", "--line-height", "12",
                                               "--line-advance", "15"])
    assert py == cpp and len(py.splitlines()) >= 6
    # refused flags are refused by both, never silently different
    assert subprocess.run([exe, "ncc", "-i", paths[0]] + base + ["--rust"], capture_output=True).returncode == 2
