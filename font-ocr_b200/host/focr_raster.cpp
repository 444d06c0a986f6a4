// focr_raster.cpp -- C++ FreeType driver: the template / glyph-raster producers of the hot path's input side
// (section 8f rank 2), i.e. the (glyph, subpixel shift) raster cache the north star adds.
//
// What it restates (call sites in /root/reference/src): font.glyph_for_char / typographic_bounds / advance / metrics /
// raster_bounds / rasterize_glyph (ncc.rs:154-194, 605-618; main.rs:49-51, 98-106, 127, 136-144, 176), the offset grid
// ncc.rs:563-573, the box logic ncc.rs:588-629, `render` ncc.rs:143-196 and the per-phase glyph rasters decode_line would
// produce cell by cell (main.rs:98-106).  PARITY UNPINNED for the rasters themselves: font-kit 0.14.3 and
// pathfinder_geometry 0.5.1 are crates.io dependencies that are not under /root/reference and there is no Rust toolchain, so
// this follows their published behaviour from memory (FT_Set_Char_Size at 72 dpi, FT_Set_Transform with a 26.6 delta,
// FT_LOAD_NO_HINTING, FT_RENDER_MODE_NORMAL, copy-blit at (bitmap_left, -bitmap_top) clipped to the canvas; raster_bounds =
// typographic bounds * size/upem, y flipped, translated, round_out).  tests/test_raster_native.py pins this file byte for
// byte to the Python producer (font-ocr_b200/raster.py), which is what every other test and bench.py feed to both the
// oracle and the GPU.
//
// There are no FreeType headers in this image: the library (any libfreetype.so.6; the tests pass Pillow's bundled one) is
// opened with dlopen and the handful of public structs used here are declared below exactly as <freetype/freetype.h>
// declares them for LP64.
#include <dirent.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "focr_host.hpp"

int focr_internal_fail(int code, const std::string &msg);  // api.cu

namespace {

// ---- <freetype/freetype.h>, LP64
struct FT_Vector { long x, y; };
struct FT_Matrix { long xx, xy, yx, yy; };
struct FT_BBox { long xMin, yMin, xMax, yMax; };
struct FT_Generic { void *data; void *finalizer; };
struct FT_Glyph_Metrics { long width, height, horiBearingX, horiBearingY, horiAdvance, vertBearingX, vertBearingY, vertAdvance; };
struct FT_Bitmap {
    unsigned int rows, width;
    int pitch;
    unsigned char *buffer;
    unsigned short num_grays;
    unsigned char pixel_mode, palette_mode;
    void *palette;
};
struct FT_GlyphSlotRec {
    void *library, *face, *next;
    unsigned int glyph_index;
    FT_Generic generic;
    FT_Glyph_Metrics metrics;
    long linearHoriAdvance, linearVertAdvance;
    FT_Vector advance;
    int format;
    FT_Bitmap bitmap;
    int bitmap_left, bitmap_top;
};
struct FT_FaceRec {
    long num_faces, face_index, face_flags, style_flags, num_glyphs;
    char *family_name, *style_name;
    int num_fixed_sizes;
    void *available_sizes;
    int num_charmaps;
    void *charmaps;
    FT_Generic generic;
    FT_BBox bbox;
    unsigned short units_per_EM;
    short ascender, descender, height, max_advance_width, max_advance_height, underline_position, underline_thickness;
    FT_GlyphSlotRec *glyph;
};
constexpr int FT_LOAD_DEFAULT = 0, FT_LOAD_NO_HINTING = 1 << 1, FT_LOAD_RENDER = 1 << 2, FT_PIXEL_MODE_GRAY = 2;

struct FtApi {
    void *so = nullptr;
    void *library = nullptr;
    int (*Init_FreeType)(void **) = nullptr;
    int (*New_Face)(void *, const char *, long, FT_FaceRec **) = nullptr;
    int (*Done_Face)(FT_FaceRec *) = nullptr;
    int (*Set_Char_Size)(FT_FaceRec *, long, long, unsigned, unsigned) = nullptr;
    void (*Set_Transform)(FT_FaceRec *, FT_Matrix *, FT_Vector *) = nullptr;
    int (*Load_Glyph)(FT_FaceRec *, unsigned, int) = nullptr;
    unsigned (*Get_Char_Index)(FT_FaceRec *, unsigned long) = nullptr;
};

// pathfinder_geometry::rect::RectF as origin + lower right, f32 lanes
struct RectF {
    float x0 = 0, y0 = 0, x1 = 0, y1 = 0;
    static RectF from_origin_size(float ox, float oy, float w, float h) { return RectF{ox, oy, ox + w, oy + h}; }
    float width() const { return x1 - x0; }
    float height() const { return y1 - y0; }
    RectF scale(float s) const { return RectF{x0 * s, y0 * s, x1 * s, y1 * s}; }
    RectF translate(float tx, float ty) const { return RectF{x0 + tx, y0 + ty, x1 + tx, y1 + ty}; }
    RectF union_rect(const RectF &o) const { return RectF{std::min(x0, o.x0), std::min(y0, o.y0), std::max(x1, o.x1), std::max(y1, o.y1)}; }
};
struct RectI { int x0, y0, x1, y1; };
RectI round_out(const RectF &r) { return RectI{(int)std::floor(r.x0), (int)std::floor(r.y0), (int)std::ceil(r.x1), (int)std::ceil(r.y1)}; }

// font-kit f32_to_ft_fixed_26_6: `(x * 64.0) as i64` (Rust `as` truncates toward zero)
long trunc_26_6(float v) { return (long)(v * 64.0f); }

}  // namespace

struct focr_host_font {
    FtApi ft;
    FT_FaceRec *face = nullptr;
    unsigned units_per_em = 0;
    bool hinting = false;   // --hinting = HintingOptions::Full: rasterise with FT_LOAD_TARGET_NORMAL instead of FT_LOAD_NO_HINTING
    float ascent = 0;
    RectF bounding_box;
    std::map<unsigned, RectF> tb_cache;
    std::map<unsigned, std::pair<float, float>> adv_cache;

    void reset_size() { ft.Set_Char_Size(face, (long)units_per_em << 6, 0, 0, 0); }   // font-kit keeps the face at ppem == units_per_em
    bool glyph_for_char(uint32_t ch, unsigned &gid) { gid = ft.Get_Char_Index(face, ch); return gid != 0; }
    bool typographic_bounds(unsigned gid, RectF &out)
    {
        auto it = tb_cache.find(gid);
        if (it == tb_cache.end()) {
            if (ft.Load_Glyph(face, gid, FT_LOAD_DEFAULT | FT_LOAD_NO_HINTING) != 0) return false;
            const FT_Glyph_Metrics &m = face->glyph->metrics;
            RectF r;
            if (m.width != 0 && m.height != 0)
                r = RectF::from_origin_size((float)m.horiBearingX / 64.0f, (float)(m.horiBearingY - m.height) / 64.0f,
                                            (float)m.width / 64.0f, (float)m.height / 64.0f);
            it = tb_cache.emplace(gid, r).first;
        }
        out = it->second;
        return true;
    }
    bool advance(unsigned gid, float &ax)
    {
        auto it = adv_cache.find(gid);
        if (it == adv_cache.end()) {
            if (ft.Load_Glyph(face, gid, FT_LOAD_DEFAULT | FT_LOAD_NO_HINTING) != 0) return false;
            const FT_Vector a = face->glyph->advance;
            it = adv_cache.emplace(gid, std::make_pair((float)a.x / 64.0f, (float)a.y / 64.0f)).first;
        }
        ax = it->second.first;
        return true;
    }
    // font-kit Loader::raster_bounds with a pure translation
    bool raster_bounds(unsigned gid, float size, float tx, float ty, RectI &out)
    {
        RectF tb;
        if (!typographic_bounds(gid, tb)) return false;
        tb = tb.scale(size / (float)units_per_em);
        const RectF flipped = RectF::from_origin_size(tb.x0, -tb.y0 - tb.height(), tb.width(), tb.height());
        out = round_out(flipped.translate(tx, ty));
        return true;
    }
    // FT bitmap for a 26.6 pen delta: A8 rows tightly packed + bitmap_left / bitmap_top
    bool glyph_bitmap(unsigned gid, float size, long dx26, long dy26, std::vector<uint8_t> &bmp, int &w, int &h, int &left, int &top)
    {
        ft.Set_Char_Size(face, trunc_26_6(size), 0, 0, 0);
        FT_Matrix mat{0x10000, 0, 0, 0x10000};
        FT_Vector delta{dx26, dy26};
        ft.Set_Transform(face, &mat, &delta);
        const int rc = ft.Load_Glyph(face, gid, FT_LOAD_DEFAULT | FT_LOAD_RENDER | (hinting ? 0 : FT_LOAD_NO_HINTING));
        bool ok = rc == 0;
        if (ok) {
            const FT_GlyphSlotRec *slot = face->glyph;
            const FT_Bitmap &b = slot->bitmap;
            w = (int)b.width, h = (int)b.rows;
            bmp.assign((size_t)w * h, 0);
            if (w && h) {
                if (b.pixel_mode != FT_PIXEL_MODE_GRAY) ok = false;
                for (int y = 0; ok && y < h; y++) memcpy(&bmp[(size_t)y * w], b.buffer + (size_t)y * std::abs(b.pitch), w);
            } else {
                w = h = 0;
            }
            left = slot->bitmap_left, top = slot->bitmap_top;
        }
        ft.Set_Transform(face, nullptr, nullptr);
        reset_size();
        return ok;
    }
    // rasterise into `canvas` (A8, cw x ch) at translation (tx, ty): font-kit Canvas::blit_from, copy-blit, clipped
    bool rasterize_glyph(std::vector<uint8_t> &canvas, int cw, int chh, unsigned gid, float size, float tx, float ty)
    {
        std::vector<uint8_t> bmp;
        int w, h, left, top;
        if (!glyph_bitmap(gid, size, trunc_26_6(tx), -trunc_26_6(ty), bmp, w, h, left, top)) return false;
        const int dx = left, dy = -top;
        const int x0 = std::max(dx, 0), y0 = std::max(dy, 0), x1 = std::min(dx + w, cw), y1 = std::min(dy + h, chh);
        for (int y = y0; y < y1 && x1 > x0; y++) memcpy(&canvas[(size_t)y * cw + x0], &bmp[(size_t)(y - dy) * w + (x0 - dx)], x1 - x0);
        return true;
    }
};

struct focr_host_tbank {
    std::vector<uint8_t> pixels;
    std::vector<uint64_t> offsets;
    std::vector<uint16_t> n_w, n_h;
    std::vector<uint32_t> letters;
    std::vector<float> corrected_y;
};
struct focr_host_gbank {
    std::vector<uint8_t> pixels;
    std::vector<focr_glyph_raster> rasters;   // [n_glyphs][64]
    std::vector<float> advance_px;
    int32_t origin_x = 0, origin_y = 0;
};

extern "C" int focr_host_font_open(const char *freetype_so, const char *font_path, focr_host_font **out)
{
    if (!freetype_so || !font_path || !out) return focr_internal_fail(FOCR_ERR_ARG, "focr_host_font_open: NULL argument");
    focr_host_font *f = new focr_host_font();
    f->ft.so = dlopen(freetype_so, RTLD_NOW | RTLD_LOCAL);
    if (!f->ft.so) {
        // a bundled FreeType (e.g. Pillow's) keeps its own dependencies (libpng, libbrotli) next to it without an rpath: load
        // the directory's other libraries first (a few passes: they depend on each other), then try again
        const std::string path = freetype_so;
        const size_t slash = path.rfind('/');
        if (slash != std::string::npos) {
            const std::string dir = path.substr(0, slash);
            for (int pass = 0; pass < 3; pass++)
                if (DIR *d = opendir(dir.c_str())) {
                    while (dirent *e = readdir(d)) {
                        const std::string name = e->d_name;
                        if (name.rfind("lib", 0) == 0 && name.find(".so") != std::string::npos && dir + "/" + name != path)
                            dlopen((dir + "/" + name).c_str(), RTLD_LAZY | RTLD_GLOBAL);
                    }
                    closedir(d);
                }
            f->ft.so = dlopen(freetype_so, RTLD_NOW | RTLD_LOCAL);
        }
    }
    if (!f->ft.so) {
        const std::string why = dlerror();
        delete f;
        return focr_internal_fail(FOCR_ERR_ARG, "focr_host_font_open: dlopen(" + std::string(freetype_so) + "): " + why);
    }
    auto sym = [&](const char *name) { return dlsym(f->ft.so, name); };
    f->ft.Init_FreeType = (int (*)(void **))sym("FT_Init_FreeType");
    f->ft.New_Face = (int (*)(void *, const char *, long, FT_FaceRec **))sym("FT_New_Face");
    f->ft.Done_Face = (int (*)(FT_FaceRec *))sym("FT_Done_Face");
    f->ft.Set_Char_Size = (int (*)(FT_FaceRec *, long, long, unsigned, unsigned))sym("FT_Set_Char_Size");
    f->ft.Set_Transform = (void (*)(FT_FaceRec *, FT_Matrix *, FT_Vector *))sym("FT_Set_Transform");
    f->ft.Load_Glyph = (int (*)(FT_FaceRec *, unsigned, int))sym("FT_Load_Glyph");
    f->ft.Get_Char_Index = (unsigned (*)(FT_FaceRec *, unsigned long))sym("FT_Get_Char_Index");
    if (!f->ft.Init_FreeType || !f->ft.New_Face || !f->ft.Set_Char_Size || !f->ft.Set_Transform || !f->ft.Load_Glyph ||
        !f->ft.Get_Char_Index) {
        delete f;
        return focr_internal_fail(FOCR_ERR_ARG, "focr_host_font_open: not a FreeType library");
    }
    if (f->ft.Init_FreeType(&f->ft.library) != 0 || f->ft.New_Face(f->ft.library, font_path, 0, &f->face) != 0) {
        delete f;
        return focr_internal_fail(FOCR_ERR_ARG, std::string("focr_host_font_open: cannot load ") + font_path);
    }
    f->units_per_em = f->face->units_per_EM;
    f->ascent = (float)f->face->ascender;
    const FT_BBox bb = f->face->bbox;
    f->bounding_box = RectF{(float)bb.xMin, (float)bb.yMin, (float)bb.xMax, (float)bb.yMax};
    f->reset_size();
    *out = f;
    return FOCR_OK;
}

// per-letter metrics in pixels at `size`: the left bearing --raw prints (ncc.rs:683-698: typographic_bounds().origin_x() * to_px)
// and the pen advance (advance / units_per_em * size, f32 in that order, main.rs:176-178)
extern "C" int focr_host_font_glyph_metrics(focr_host_font *f, uint32_t letter, float size, float *bearing_x_px, float *advance_px)
{
    if (!f || !bearing_x_px || !advance_px) return focr_internal_fail(FOCR_ERR_ARG, "focr_host_font_glyph_metrics: NULL argument");
    unsigned gid;
    if (!f->glyph_for_char(letter, gid)) return focr_internal_fail(FOCR_ERR_ARG, "panic: no glyph for U+" + std::to_string(letter));
    RectF tb;
    float ax;
    if (!f->typographic_bounds(gid, tb) || !f->advance(gid, ax)) return focr_internal_fail(FOCR_ERR_ARG, "FT_Load_Glyph failed");
    const float to_px = (1.0f / (float)f->units_per_em) * size;
    *bearing_x_px = tb.x0 * to_px;
    *advance_px = (ax / (float)f->units_per_em) * size;
    return FOCR_OK;
}

// the reference's --hinting (HintingOptions::Full(size), ncc.rs:547-551, main.rs:394-398): hinted rasters; metrics stay unhinted
extern "C" void focr_host_font_set_hinting(focr_host_font *f, int on)
{
    if (f) f->hinting = on != 0;
}

extern "C" void focr_host_font_close(focr_host_font *f)
{
    if (!f) return;
    if (f->face && f->ft.Done_Face) f->ft.Done_Face(f->face);
    delete f;   // the FreeType library object and the dlopen handle stay for the life of the process
}

// ncc.rs:587-640: every template get_hits renders for one page, in its iteration order (offset index, alphabet index).
// box_mode: 0 = BoxSize::Alphabet (default, ncc.rs:600-626), 1 = Font (ncc.rs:589-599), 2 = Char (ncc.rs:627)
extern "C" int focr_host_tbank_render(focr_host_font *f, float size, const uint32_t *alphabet, uint32_t n_alphabet, uint32_t x_bits,
                                      uint32_t y_bits, int box_mode, int pad_x, int pad_y, focr_host_tbank **out)
{
    if (!f || !alphabet || !n_alphabet || !out || x_bits > 8 || y_bits > 8 || box_mode < 0 || box_mode > 2)
        return focr_internal_fail(FOCR_ERR_ARG, "focr_host_tbank_render: bad argument");
    std::vector<unsigned> gids(n_alphabet);
    for (uint32_t i = 0; i < n_alphabet; i++)
        if (!f->glyph_for_char(alphabet[i], gids[i]))   // the reference .unwrap()s a None here (ncc.rs:154)
            return focr_internal_fail(FOCR_ERR_ARG, "panic: no glyph for U+" + std::to_string(alphabet[i]));
    focr_host_tbank *b = new focr_host_tbank();
    const float to_px = (1.0f / (float)f->units_per_em) * size;
    const float xd = 1.0f / (float)(1u << x_bits), yd = 1.0f / (float)(1u << y_bits);
    uint32_t oi = 0;
    for (uint32_t xi = 0; xi < (1u << x_bits); xi++)          // ncc.rs:563-573: x-major
        for (uint32_t yi = 0; yi < (1u << y_bits); yi++, oi++) {
            const float off_x = (float)xi * xd, off_y = (float)yi * yd;
            float y_offset = 0.f;
            int cw = 0, chh = 0;
            bool fixed_canvas = true;
            if (box_mode == 0) {                               // ncc.rs:600-626
                RectF bounds;                                  // RectF::default(): the union is seeded with the point (0,0)
                for (uint32_t i = 0; i < n_alphabet; i++) {
                    RectF tb;
                    RectI rb;
                    if (!f->typographic_bounds(gids[i], tb) || !f->raster_bounds(gids[i], size, off_x, off_y, rb)) {
                        delete b;
                        return focr_internal_fail(FOCR_ERR_ARG, "focr_host_tbank_render: FT_Load_Glyph failed");
                    }
                    const RectF gb = tb.scale(to_px);
                    const float bearing_y = gb.y0 + gb.height();
                    y_offset = std::max(y_offset, std::ceil(bearing_y));
                    bounds = bounds.union_rect(RectF{(float)rb.x0, (float)rb.y0, (float)rb.x1, (float)rb.y1});
                }
                const RectI bi = round_out(bounds);
                cw = bi.x1 - bi.x0, chh = bi.y1 - bi.y0;
            } else if (box_mode == 1) {                        // ncc.rs:589-599
                const RectI bi = round_out(f->bounding_box.scale(to_px));
                cw = bi.x1 - bi.x0, chh = bi.y1 - bi.y0;
                y_offset = std::ceil(f->ascent * to_px);
            } else {
                fixed_canvas = false;                          // ncc.rs:627: tight per-glyph box
            }
            const float cor_x = off_x, cor_y = off_y + y_offset;   // ncc.rs:629
            for (uint32_t i = 0; i < n_alphabet; i++) {        // `render`, ncc.rs:143-196
                RectI rb;
                if (!f->raster_bounds(gids[i], size, cor_x, cor_y, rb)) {
                    delete b;
                    return focr_internal_fail(FOCR_ERR_ARG, "focr_host_tbank_render: FT_Load_Glyph failed");
                }
                int w = cw, h = chh;
                float ox = 0.f, oy = 0.f;
                if (!fixed_canvas) w = rb.x1 - rb.x0, h = rb.y1 - rb.y0, ox = (float)-rb.x0, oy = (float)-rb.y0;
                w += 2 * pad_x, h += 2 * pad_y;
                if (w <= 0 || h <= 0 || w > 65535 || h > 65535) {
                    delete b;
                    return focr_internal_fail(FOCR_ERR_UNSUPPORTED, "focr_host_tbank_render: empty or oversized canvas");
                }
                std::vector<uint8_t> canvas((size_t)w * h, 0);
                const float tx = (ox + (float)pad_x) + cor_x, ty = (oy + (float)pad_y) + cor_y;
                if (!f->rasterize_glyph(canvas, w, h, gids[i], size, tx, ty)) {
                    delete b;
                    return focr_internal_fail(FOCR_ERR_ARG, "focr_host_tbank_render: rasterisation failed");
                }
                b->offsets.push_back(b->pixels.size());
                b->pixels.insert(b->pixels.end(), canvas.begin(), canvas.end());
                b->n_w.push_back((uint16_t)w);
                b->n_h.push_back((uint16_t)h);
                b->letters.push_back(alphabet[i]);
                b->corrected_y.push_back(cor_y);
            }
        }
    *out = b;
    return FOCR_OK;
}

extern "C" uint32_t focr_host_tbank_count(const focr_host_tbank *b) { return b ? (uint32_t)b->offsets.size() : 0; }
extern "C" uint64_t focr_host_tbank_pixel_bytes(const focr_host_tbank *b) { return b ? b->pixels.size() : 0; }
// the arrays focr_bank_create takes (+ letters and the corrected y offset --raw prints); any pointer may be NULL
extern "C" int focr_host_tbank_get(const focr_host_tbank *b, uint8_t *pixels, uint64_t *offsets, uint16_t *n_w, uint16_t *n_h,
                                   uint32_t *letters, float *corrected_y)
{
    if (!b) return focr_internal_fail(FOCR_ERR_ARG, "focr_host_tbank_get: NULL bank");
    const size_t n = b->offsets.size();
    if (pixels) memcpy(pixels, b->pixels.data(), b->pixels.size());
    if (offsets) memcpy(offsets, b->offsets.data(), n * 8);
    if (n_w) memcpy(n_w, b->n_w.data(), n * 2);
    if (n_h) memcpy(n_h, b->n_h.data(), n * 2);
    if (letters) memcpy(letters, b->letters.data(), n * 4);
    if (corrected_y) memcpy(corrected_y, b->corrected_y.data(), n * 4);
    return FOCR_OK;
}
extern "C" void focr_host_tbank_free(focr_host_tbank *b) { delete b; }

// focr's raster cache (README.md:44): 64 horizontal 26.6 phases of every alphabet glyph at the vertical delta decode_line's
// origin gives them (main.rs:133-147), and the f32 pen advances of main.rs:176-178 -- the arrays focr_glyph_bank_create takes
extern "C" int focr_host_gbank_render(focr_host_font *f, float size, const uint32_t *alphabet, uint32_t n_alphabet, float kern_x,
                                      focr_host_gbank **out)
{
    if (!f || !alphabet || !n_alphabet || !out) return focr_internal_fail(FOCR_ERR_ARG, "focr_host_gbank_render: bad argument");
    std::vector<unsigned> gids(n_alphabet);
    for (uint32_t i = 0; i < n_alphabet; i++)
        if (!f->glyph_for_char(alphabet[i], gids[i]))   // main.rs:127 .unwrap()
            return focr_internal_fail(FOCR_ERR_ARG, "panic: no glyph for U+" + std::to_string(alphabet[i]));
    focr_host_gbank *b = new focr_host_gbank();
    int x0 = 0, y0 = 0;                                 // RectF::default() seeds the union (main.rs:133-146)
    for (unsigned gid : gids) {
        RectI rb;
        if (!f->raster_bounds(gid, size, 0.f, 0.f, rb)) {
            delete b;
            return focr_internal_fail(FOCR_ERR_ARG, "focr_host_gbank_render: FT_Load_Glyph failed");
        }
        x0 = std::min(x0, rb.x0), y0 = std::min(y0, rb.y0);
    }
    b->origin_x = -x0, b->origin_y = -y0;               // main.rs:147
    const float upem = (float)f->units_per_em;
    b->rasters.resize((size_t)n_alphabet * 64);
    for (uint32_t g = 0; g < n_alphabet; g++) {
        float ax;
        if (!f->advance(gids[g], ax)) {
            delete b;
            return focr_internal_fail(FOCR_ERR_ARG, "focr_host_gbank_render: FT_Load_Glyph failed");
        }
        b->advance_px.push_back(((ax / upem) * size) * kern_x);   // main.rs:176-178, f32 in that order
        for (int ph = 0; ph < 64; ph++) {
            std::vector<uint8_t> bmp;
            int w, h, left, top;
            if (!f->glyph_bitmap(gids[g], size, ph, -(long)b->origin_y * 64, bmp, w, h, left, top)) {
                delete b;
                return focr_internal_fail(FOCR_ERR_ARG, "focr_host_gbank_render: rasterisation failed");
            }
            focr_glyph_raster &r = b->rasters[(size_t)g * 64 + ph];
            r.offset = b->pixels.size();
            r.left = (int16_t)left, r.top = (int16_t)-top, r.w = (uint16_t)w, r.h = (uint16_t)h;
            b->pixels.insert(b->pixels.end(), bmp.begin(), bmp.end());
        }
    }
    if (b->pixels.empty()) b->pixels.push_back(0);
    *out = b;
    return FOCR_OK;
}
extern "C" uint64_t focr_host_gbank_pixel_bytes(const focr_host_gbank *b) { return b ? b->pixels.size() : 0; }
extern "C" int focr_host_gbank_get(const focr_host_gbank *b, uint8_t *pixels, focr_glyph_raster *rasters, float *advance_px,
                                   int32_t *origin_xy)
{
    if (!b) return focr_internal_fail(FOCR_ERR_ARG, "focr_host_gbank_get: NULL bank");
    if (pixels) memcpy(pixels, b->pixels.data(), b->pixels.size());
    if (rasters) memcpy(rasters, b->rasters.data(), b->rasters.size() * sizeof(focr_glyph_raster));
    if (advance_px) memcpy(advance_px, b->advance_px.data(), b->advance_px.size() * 4);
    if (origin_xy) origin_xy[0] = b->origin_x, origin_xy[1] = b->origin_y;
    return FOCR_OK;
}
extern "C" void focr_host_gbank_free(focr_host_gbank *b) { delete b; }
