// focr_cli.cpp -- flag-compatible C++ front-ends of the reference binaries over the C ABI:
//
//     focr_cli ncc  <flags of src/ncc.rs:486-542>      text (ncc.rs:869-876), --csv (ncc.rs:849-867), --raw (ncc.rs:683-698)
//     focr_cli focr <flags of src/main.rs:342-385>     one decoded line per row (main.rs:468-471)
//
// The host flow a maintainer would keep, written once in C++: load the pages (`image::open(..).into_luma8()`, ncc.rs:575:
// PNG through zlib and PNM are decoded here), render the (glyph, subpixel shift) raster cache once (focr_raster.cpp),
// scan / decode the pages in GPU batches (focr_ncc_scan / focr_decode_pages), run process_hits (focr_host.cpp) and print in
// the reference's formats.  Same output as the Python front-end font-ocr_b200/cli.py (tests/test_cli.py compares them).
// Extensions are opt-in and named as such: --spaces, --space-advance, --max-matches, --device, --batch, --freetype.
// --rust, --test, --verify are refused, never silently different.
#include <zlib.h>

#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "focr_host.hpp"

using namespace focr_host;

namespace {

struct Image {
    uint32_t w = 0, h = 0;
    std::vector<uint8_t> luma;
};

[[noreturn]] void die(const std::string &msg)
{
    fprintf(stderr, "focr_cli: %s\n", msg.c_str());
    exit(1);
}

// the image crate's conversions: Rec.709 integer weights for colour, (c + 128) / 257 for 16-bit samples
inline uint32_t luma_of(uint32_t r, uint32_t g, uint32_t b) { return (uint32_t)((2126ull * r + 7152ull * g + 722ull * b) / 10000ull); }
inline uint8_t narrow16(uint32_t v) { return (uint8_t)((v + 128) / 257); }

std::vector<uint8_t> read_file(const std::string &path)
{
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) die("cannot open " + path);
    std::vector<uint8_t> d;
    uint8_t buf[1 << 16];
    size_t n;
    while ((n = fread(buf, 1, sizeof(buf), f)) > 0) d.insert(d.end(), buf, buf + n);
    fclose(f);
    return d;
}

// ---- PNM (P5 / P6, 8- or 16-bit, big-endian samples)
bool load_pnm(const std::vector<uint8_t> &d, Image &im)
{
    if (d.size() < 3 || d[0] != 'P' || (d[1] != '5' && d[1] != '6')) return false;
    size_t p = 2;
    auto next_int = [&]() -> uint32_t {
        for (;;) {
            while (p < d.size() && isspace(d[p])) p++;
            if (p < d.size() && d[p] == '#') {
                while (p < d.size() && d[p] != '\n') p++;
                continue;
            }
            break;
        }
        uint32_t v = 0;
        while (p < d.size() && isdigit(d[p])) v = v * 10 + (d[p++] - '0');
        return v;
    };
    const uint32_t w = next_int(), h = next_int(), maxv = next_int();
    p++;  // the single whitespace byte after maxval
    const int ch = d[1] == '6' ? 3 : 1, bps = maxv > 255 ? 2 : 1;
    if (!w || !h || p + (size_t)w * h * ch * bps > d.size()) die("truncated PNM");
    im.w = w, im.h = h;
    im.luma.resize((size_t)w * h);
    for (size_t i = 0; i < (size_t)w * h; i++) {
        uint32_t c[3];
        for (int k = 0; k < ch; k++) {
            const uint8_t *s = &d[p + (i * ch + k) * bps];
            c[k] = bps == 2 ? ((uint32_t)s[0] << 8 | s[1]) : s[0];
        }
        const uint32_t v = ch == 3 ? luma_of(c[0], c[1], c[2]) : c[0];
        im.luma[i] = bps == 2 ? narrow16(v) : (uint8_t)v;
    }
    return true;
}

// ---- PNG (non-interlaced; gray, gray+alpha, RGB, RGBA at 8 / 16 bits, palette and gray at 1..8 bits)
bool load_png(const std::vector<uint8_t> &d, Image &im)
{
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    if (d.size() < 8 || memcmp(d.data(), sig, 8) != 0) return false;
    auto be32 = [&](size_t o) { return (uint32_t)d[o] << 24 | (uint32_t)d[o + 1] << 16 | (uint32_t)d[o + 2] << 8 | d[o + 3]; };
    uint32_t w = 0, h = 0;
    int depth = 0, ctype = 0, interlace = 0;
    std::vector<uint8_t> idat, plte;
    for (size_t p = 8; p + 12 <= d.size();) {
        const uint32_t len = be32(p);
        const std::string type((const char *)&d[p + 4], 4);
        if (p + 12 + len > d.size()) die("truncated PNG");
        const uint8_t *body = &d[p + 8];
        if (type == "IHDR") {
            w = be32(p + 8), h = be32(p + 12);
            depth = body[8], ctype = body[9], interlace = body[12];
        } else if (type == "PLTE") {
            plte.assign(body, body + len);
        } else if (type == "IDAT") {
            idat.insert(idat.end(), body, body + len);
        } else if (type == "IEND") {
            break;
        }
        p += 12 + len;
    }
    if (!w || !h) die("PNG without IHDR");
    if (interlace) die("interlaced PNG is not supported by focr_cli (use the Python front-end)");
    const int channels = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
    if (!channels || (depth != 8 && depth != 16 && !((ctype == 0 || ctype == 3) && (depth == 1 || depth == 2 || depth == 4))))
        die("unsupported PNG colour type / bit depth");
    const size_t bpp = std::max<size_t>(1, (size_t)channels * depth / 8);      // bytes per complete pixel (filter unit)
    const size_t stride = ((size_t)w * channels * depth + 7) / 8;
    std::vector<uint8_t> raw((stride + 1) * h);
    uLongf out_len = raw.size();
    if (uncompress(raw.data(), &out_len, idat.data(), idat.size()) != Z_OK || out_len != raw.size()) die("PNG inflate failed");
    std::vector<uint8_t> prev(stride, 0), cur(stride);
    im.w = w, im.h = h;
    im.luma.resize((size_t)w * h);
    for (uint32_t y = 0; y < h; y++) {
        const uint8_t *row = &raw[(stride + 1) * y];
        const int filter = row[0];
        for (size_t i = 0; i < stride; i++) {
            const int a = i >= bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= bpp ? prev[i - bpp] : 0;
            int pred = 0;
            switch (filter) {
                case 0: pred = 0; break;
                case 1: pred = a; break;
                case 2: pred = b; break;
                case 3: pred = (a + b) / 2; break;
                case 4: {
                    const int pa = std::abs(b - c), pb = std::abs(a - c), pc = std::abs(a + b - 2 * c);
                    pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
                    break;
                }
                default: die("bad PNG filter");
            }
            cur[i] = (uint8_t)(row[1 + i] + pred);
        }
        for (uint32_t x = 0; x < w; x++) {
            uint32_t v;
            if (depth < 8) {   // packed gray / palette indices
                const uint32_t bit = x * depth, s = (cur[bit >> 3] >> (8 - depth - (bit & 7))) & ((1u << depth) - 1);
                if (ctype == 3) {
                    if (3 * s + 2 >= plte.size()) die("PNG palette index out of range");
                    v = luma_of(plte[3 * s], plte[3 * s + 1], plte[3 * s + 2]);
                } else {
                    v = s * 255 / ((1u << depth) - 1);
                }
                im.luma[(size_t)y * w + x] = (uint8_t)v;
                continue;
            }
            const uint8_t *px = &cur[(size_t)x * channels * (depth / 8)];
            auto sample = [&](int k) -> uint32_t { return depth == 16 ? ((uint32_t)px[2 * k] << 8 | px[2 * k + 1]) : px[k]; };
            if (ctype == 3) {
                const uint32_t s = px[0];
                if (3 * s + 2 >= plte.size()) die("PNG palette index out of range");
                v = luma_of(plte[3 * s], plte[3 * s + 1], plte[3 * s + 2]);
            } else if (channels >= 3) {
                v = luma_of(sample(0), sample(1), sample(2));
            } else {
                v = sample(0);   // gray (+ alpha, dropped)
            }
            im.luma[(size_t)y * w + x] = depth == 16 ? narrow16(v) : (uint8_t)v;
        }
        prev.swap(cur);
    }
    return true;
}

Image load_luma8(const std::string &path)
{
    const std::vector<uint8_t> d = read_file(path);
    Image im;
    if (load_png(d, im) || load_pnm(d, im)) return im;
    die(path + ": only PNG and binary PNM (P5/P6) are decoded by focr_cli (use the Python front-end for other formats)");
}

// Rust `{}` of an f32: shortest digits that round-trip, never an exponent, no trailing ".0"
std::string rust_f32(float v)
{
    if (std::isnan(v)) return "NaN";
    if (std::isinf(v)) return v > 0 ? "inf" : "-inf";
    char buf[128];
    auto r = std::to_chars(buf, buf + sizeof(buf), v, std::chars_format::fixed);
    return std::string(buf, r.ptr);
}

std::u32string utf8_to_u32(const std::string &s)
{
    std::u32string out;
    for (size_t i = 0; i < s.size();) {
        const uint8_t c = (uint8_t)s[i];
        uint32_t cp;
        int n;
        if (c < 0x80) cp = c, n = 1;
        else if ((c >> 5) == 6) cp = c & 0x1F, n = 2;
        else if ((c >> 4) == 14) cp = c & 0x0F, n = 3;
        else cp = c & 0x07, n = 4;
        for (int k = 1; k < n && i + k < s.size(); k++) cp = (cp << 6) | ((uint8_t)s[i + k] & 0x3F);
        out.push_back(cp);
        i += n;
    }
    return out;
}
std::string u32_to_utf8(const std::u32string &s)
{
    std::string out;
    for (char32_t cp : s) {
        if (cp < 0x80) out.push_back((char)cp);
        else if (cp < 0x800) out.push_back((char)(0xC0 | cp >> 6)), out.push_back((char)(0x80 | (cp & 0x3F)));
        else if (cp < 0x10000)
            out.push_back((char)(0xE0 | cp >> 12)), out.push_back((char)(0x80 | ((cp >> 6) & 0x3F))), out.push_back((char)(0x80 | (cp & 0x3F)));
        else
            out.push_back((char)(0xF0 | cp >> 18)), out.push_back((char)(0x80 | ((cp >> 12) & 0x3F))),
                out.push_back((char)(0x80 | ((cp >> 6) & 0x3F))), out.push_back((char)(0x80 | (cp & 0x3F)));
    }
    return out;
}

// ---- flags
struct Args {
    std::vector<std::string> img;
    std::map<std::string, std::string> val;
    std::map<std::string, bool> flag;
    const std::string &s(const std::string &k) const { return val.at(k); }
    double f(const std::string &k) const { return atof(val.at(k).c_str()); }
    long i(const std::string &k) const { return atol(val.at(k).c_str()); }
    bool has(const std::string &k) const { return val.count(k) && !val.at(k).empty(); }
    bool on(const std::string &k) const { auto it = flag.find(k); return it != flag.end() && it->second; }
};

// spec: "--long" or "-s/--long"; values map long name -> default ("" = required / absent), flags list long names
Args parse(int argc, char **argv, const std::map<std::string, std::string> &short_of, std::map<std::string, std::string> values,
           const std::vector<std::string> &flags, const std::vector<std::string> &required)
{
    Args a;
    a.val = std::move(values);
    for (int k = 0; k < argc; k++) {
        std::string t = argv[k];
        if (short_of.count(t)) t = short_of.at(t);
        if (t == "--img") {
            while (k + 1 < argc && argv[k + 1][0] != '-') a.img.push_back(argv[++k]);
            continue;
        }
        if (std::find(flags.begin(), flags.end(), t) != flags.end()) {
            a.flag[t] = true;
            continue;
        }
        if (!a.val.count(t)) die("unknown flag " + t);
        if (k + 1 >= argc) die(t + " needs a value");
        a.val[t] = argv[++k];
    }
    for (auto &r : required)
        if (!a.has(r)) die("missing " + r);
    if (a.img.empty()) die("missing -i/--img");
    return a;
}

const char *NCC_ALPHABET = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789=+<>(){};:/-";   // ncc.rs:28-29
const char *FOCR_ALPHABET = "> =ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";          // main.rs:13-14

std::string freetype_path(const Args &a)
{
    if (a.has("--freetype")) return a.s("--freetype");
    if (const char *e = getenv("FOCR_FREETYPE_LIB")) return e;
    return "libfreetype.so.6";
}

#define CHECK(call)                                                            \
    do {                                                                       \
        if ((call) != FOCR_OK) die(std::string(#call) + ": " + focr_last_error()); \
    } while (0)

// group page indices by image size, keeping page order inside a group (one batch = pages of one size)
std::map<std::pair<uint32_t, uint32_t>, std::vector<size_t>> by_size(const std::vector<Image> &images)
{
    std::map<std::pair<uint32_t, uint32_t>, std::vector<size_t>> g;
    for (size_t i = 0; i < images.size(); i++) g[{images[i].h, images[i].w}].push_back(i);
    return g;
}

int ncc_main(int argc, char **argv)
{
    const Args a = parse(argc, argv, {{"-i", "--img"}, {"-f", "--font"}, {"-t", "--text-size"}, {"-a", "--alphabet"}, {"-v", "--verbose"}},
                         {{"--font", ""}, {"--text-size", ""}, {"--x-bits", "0"}, {"--y-bits", "0"}, {"--threshold", "0.8"},
                          {"--anchor-threshold", "0.95"}, {"--overlap", "5"}, {"--alphabet", NCC_ALPHABET}, {"--box-size", "alphabet"},
                          {"--x-padding", "0"}, {"--y-padding", "0"}, {"--device", "0"}, {"--batch", "16"}, {"--max-matches", "1024"},
                          {"--space-advance", ""}, {"--freetype", ""}},
                         {"--hinting", "--rust", "--csv", "--raw", "--verbose", "--spaces", "--save-letters"}, {"--font", "--text-size"});
    if (a.on("--rust")) {
        fprintf(stderr, "ncc: --rust is not supported by the B200 path (DESIGN.md section 7)\n");
        return 2;
    }
    if (a.on("--save-letters")) {   // diagnostics (ncc.rs:642-650): the Python front-end writes the template PNGs
        fprintf(stderr, "ncc: --save-letters is only implemented by the Python front-end (python -m font_ocr_b200.cli ncc ...)\n");
        return 2;
    }
    if (a.on("--raw") && a.img.size() != 1) die("--raw takes exactly one image (ncc.rs:833-837)");
    const std::map<std::string, int> modes = {{"alphabet", 0}, {"font", 1}, {"char", 2}};
    if (!modes.count(a.s("--box-size"))) die("bad --box-size");   // the reference .unwrap()s the TryFrom error (ncc.rs:559)
    const float size = (float)a.f("--text-size");
    const uint32_t xb = (uint32_t)a.i("--x-bits"), yb = (uint32_t)a.i("--y-bits"), n_out = (uint32_t)a.i("--max-matches");
    const std::u32string alphabet = utf8_to_u32(a.s("--alphabet"));
    std::vector<uint32_t> alpha(alphabet.begin(), alphabet.end());

    focr_host_font *font = nullptr;
    CHECK(focr_host_font_open(freetype_path(a).c_str(), a.s("--font").c_str(), &font));
    focr_host_font_set_hinting(font, a.on("--hinting"));   // HintingOptions::Full(text_size), ncc.rs:547-551
    focr_host_tbank *tb = nullptr;
    CHECK(focr_host_tbank_render(font, size, alpha.data(), (uint32_t)alpha.size(), xb, yb, modes.at(a.s("--box-size")),
                                 (int)a.i("--x-padding"), (int)a.i("--y-padding"), &tb));
    const uint32_t T = focr_host_tbank_count(tb);
    std::vector<uint8_t> px(focr_host_tbank_pixel_bytes(tb));
    std::vector<uint64_t> offs(T);
    std::vector<uint16_t> nw(T), nh(T);
    std::vector<uint32_t> letters(T);
    std::vector<float> cor_y(T);
    CHECK(focr_host_tbank_get(tb, px.data(), offs.data(), nw.data(), nh.data(), letters.data(), cor_y.data()));
    if (a.on("--verbose")) fprintf(stderr, "templates %u\n", T);

    std::vector<Image> images;
    for (auto &p : a.img) images.push_back(load_luma8(p));
    focr_ctx *ctx = nullptr;
    CHECK(focr_ctx_create((int)a.i("--device"), &ctx));
    focr_bank *bank = nullptr;
    CHECK(focr_bank_create(ctx, px.data(), offs.data(), nw.data(), nh.data(), T, &bank));

    std::vector<std::vector<std::vector<MatchWithLetter>>> page_lines(images.size());
    const size_t batch = std::max<long>(1, a.i("--batch"));
    for (auto &grp : by_size(images)) {
        const uint32_t r_h = grp.first.first, r_w = grp.first.second;
        for (size_t b0 = 0; b0 < grp.second.size(); b0 += batch) {
            const size_t P = std::min(batch, grp.second.size() - b0);
            std::vector<uint8_t> pages((size_t)P * r_w * r_h);
            for (size_t q = 0; q < P; q++) memcpy(&pages[q * r_w * r_h], images[grp.second[b0 + q]].luma.data(), (size_t)r_w * r_h);
            std::vector<focr_match> out((size_t)P * T * n_out);
            std::vector<uint32_t> counts((size_t)P * T);
            CHECK(focr_ncc_scan(ctx, bank, pages.data(), (size_t)r_w * r_h, r_w, r_h, (uint32_t)P, (float)a.f("--threshold"), n_out,
                                out.data(), counts.data()));
            for (size_t q = 0; q < P; q++) {
                const size_t page = grp.second[b0 + q];
                if (a.on("--raw")) {   // ncc.rs:683-698: one line per hit in scan order (offset, letter, y, x)
                    for (uint32_t t = 0; t < T; t++) {
                        float bearing_x = 0, adv = 0;
                        CHECK(focr_host_font_glyph_metrics(font, letters[t], size, &bearing_x, &adv));
                        const uint32_t oi = t / (uint32_t)alpha.size();
                        const float off_x = (float)(oi >> yb) * (1.0f / (float)(1u << xb)), off_y = (float)(oi & ((1u << yb) - 1)) * (1.0f / (float)(1u << yb));
                        const std::string tail = std::to_string(nw[t]) + "," + std::to_string(nh[t]) + "," + rust_f32(bearing_x) + "," +
                                                 rust_f32(cor_y[t]) + "," + rust_f32(off_x) + "," + rust_f32(off_y);
                        const focr_match *m = &out[(q * T + t) * n_out];
                        for (uint32_t k = 0; k < counts[q * T + t]; k++)
                            printf("%u,%s,%s,%u,%u,%s\n", letters[t], rust_f32((float)m[k].x + (float)nw[t] * 0.5f).c_str(),
                                   rust_f32((float)m[k].y + (float)nh[t] * 0.5f).c_str(), m[k].x, m[k].y, tail.c_str());
                    }
                    continue;
                }
                std::vector<MatchWithLetter> all_hits;   // get_hits order (ncc.rs:675-681): (template, y, x)
                for (uint32_t t = 0; t < T; t++) {
                    const focr_match *m = &out[(q * T + t) * n_out];
                    for (uint32_t k = 0; k < counts[q * T + t]; k++)
                        all_hits.push_back(MatchWithLetter{RectI{m[k].x, m[k].y, nw[t], nh[t]}, m[k].similarity, letters[t], (uint32_t)all_hits.size()});
                }
                try {
                    page_lines[page] = process_hits((float)a.f("--anchor-threshold"), (int32_t)a.i("--overlap"), all_hits);
                } catch (const Panic &) {
                    // no anchor line on this page: the reference panics in partition_by (ncc.rs:1040); print nothing for it
                }
            }
        }
    }
    focr_bank_destroy(bank);
    focr_ctx_destroy(ctx);
    if (a.on("--raw")) return 0;
    float space_px = 0.f;
    std::map<uint32_t, float> adv_px;
    if (a.on("--spaces")) {   // extension (README.md:46): pen advances in pixels, f32 like main.rs:176-178
        for (uint32_t ch : alpha) {
            float bx, adv;
            CHECK(focr_host_font_glyph_metrics(font, ch, size, &bx, &adv));
            adv_px[ch] = adv;
        }
        if (a.has("--space-advance")) {
            space_px = (float)a.f("--space-advance");
        } else {
            float bx;
            CHECK(focr_host_font_glyph_metrics(font, ' ', size, &bx, &space_px));
        }
    }
    for (size_t i = 0; i < images.size(); i++)   // pages.sort_by_key(|(i, _)| *i), ncc.rs:847
        for (auto &line : page_lines[i]) {
            if (a.on("--csv")) {                 // ncc.rs:849-867
                for (auto &m : line)
                    printf("%zu,%u,%s,%s,%d,%d,%d,%d\n", i, m.letter, rust_f32((float)m.rect.x + (float)m.rect.w * 0.5f).c_str(),
                           rust_f32((float)m.rect.y + (float)m.rect.h * 0.5f).c_str(), m.rect.x, m.rect.y, m.rect.w, m.rect.h);
            } else if (a.on("--spaces")) {
                printf("%s\n", u32_to_utf8(line_text_with_spaces(line, [&](uint32_t l) { return adv_px.count(l) ? adv_px[l] : 0.f; }, space_px)).c_str());
            } else {                             // ncc.rs:869-876
                std::u32string s;
                for (auto &m : line) s.push_back(m.letter);
                printf("%s\n", u32_to_utf8(s).c_str());
            }
        }
    focr_host_tbank_free(tb);
    focr_host_font_close(font);
    return 0;
}

int focr_main(int argc, char **argv)
{
    const Args a = parse(argc, argv, {{"-i", "--img"}, {"-f", "--font"}, {"-a", "--alphabet"}, {"-t", "--text-size"}, {"-k", "--kerning"},
                                      {"-x", "--x"}, {"-y", "--y"}, {"-w", "--width"}},
                         {{"--font", ""}, {"--alphabet", FOCR_ALPHABET}, {"--text-size", ""}, {"--kerning", "1.0"}, {"--x", "0"}, {"--y", "0"},
                          {"--width", ""}, {"--line-height", ""}, {"--line-advance", ""}, {"--test", ""}, {"--verify", ""}, {"--device", "0"},
                          {"--batch", "16"}, {"--freetype", ""}},
                         {"--hinting"}, {"--font", "--text-size", "--width", "--line-height", "--line-advance"});
    if (a.has("--test") || a.has("--verify")) {
        fprintf(stderr, "focr: --test / --verify are not supported by the B200 path (DESIGN.md section 7)\n");
        return 2;
    }
    const std::u32string alphabet = utf8_to_u32(a.s("--alphabet"));
    std::vector<uint32_t> alpha(alphabet.begin(), alphabet.end());
    focr_host_font *font = nullptr;
    CHECK(focr_host_font_open(freetype_path(a).c_str(), a.s("--font").c_str(), &font));
    focr_host_font_set_hinting(font, a.on("--hinting"));   // HintingOptions::Full(text_size), main.rs:394-398
    focr_host_gbank *gb = nullptr;
    CHECK(focr_host_gbank_render(font, (float)a.f("--text-size"), alpha.data(), (uint32_t)alpha.size(), (float)a.f("--kerning"), &gb));
    std::vector<uint8_t> px(focr_host_gbank_pixel_bytes(gb));
    std::vector<focr_glyph_raster> ras(alpha.size() * 64);
    std::vector<float> adv(alpha.size());
    int32_t origin[2];
    CHECK(focr_host_gbank_get(gb, px.data(), ras.data(), adv.data(), origin));
    std::vector<Image> images;
    for (auto &p : a.img) images.push_back(load_luma8(p));
    focr_ctx *ctx = nullptr;
    CHECK(focr_ctx_create((int)a.i("--device"), &ctx));
    focr_glyph_bank *bank = nullptr;
    CHECK(focr_glyph_bank_create(ctx, px.data(), px.size(), ras.data(), adv.data(), (uint32_t)alpha.size(), origin[0], &bank));
    std::vector<std::vector<DecodedLine>> texts(images.size());
    const size_t batch = std::max<long>(1, a.i("--batch"));
    for (auto &grp : by_size(images)) {
        const uint32_t r_h = grp.first.first, r_w = grp.first.second;
        for (size_t b0 = 0; b0 < grp.second.size(); b0 += batch) {
            const size_t P = std::min(batch, grp.second.size() - b0);
            std::vector<uint8_t> pages((size_t)P * r_w * r_h);
            for (size_t q = 0; q < P; q++) memcpy(&pages[q * r_w * r_h], images[grp.second[b0 + q]].luma.data(), (size_t)r_w * r_h);
            std::vector<std::vector<DecodedLine>> res;
            try {
                res = decode_image_vec(ctx, bank, alphabet, pages.data(), r_w, r_h, (uint32_t)P, (uint32_t)a.i("--x"), (uint32_t)a.i("--y"),
                                       (uint32_t)a.i("--width"), (uint32_t)a.i("--line-height"), (uint32_t)a.i("--line-advance"));
            } catch (const Panic &e) {
                die(e.what());
            }
            for (size_t q = 0; q < P; q++) texts[grp.second[b0 + q]] = std::move(res[q]);
        }
    }
    for (size_t i = 0; i < images.size(); i++)   // liness.sort_by_key(|(i, _)| *i), main.rs:468-471
        for (auto &l : texts[i]) printf("%s\n", u32_to_utf8(l.text).c_str());
    focr_glyph_bank_destroy(bank);
    focr_ctx_destroy(ctx);
    focr_host_gbank_free(gb);
    focr_host_font_close(font);
    return 0;
}

}  // namespace

int main(int argc, char **argv)
{
    if (argc == 3 && strcmp(argv[1], "luma") == 0) {   // diagnostics: what into_luma8() gives for a file, as a binary PGM on stdout
        const Image im = load_luma8(argv[2]);
        printf("P5\n%u %u\n255\n", im.w, im.h);
        fwrite(im.luma.data(), 1, im.luma.size(), stdout);
        return 0;
    }
    if (argc < 2 || (strcmp(argv[1], "ncc") != 0 && strcmp(argv[1], "focr") != 0)) {
        fprintf(stderr, "usage: focr_cli {ncc|focr} <flags of the reference binary>\n");
        return 2;
    }
    return strcmp(argv[1], "ncc") == 0 ? ncc_main(argc - 2, argv + 2) : focr_main(argc - 2, argv + 2);
}
