// focr_host.hpp -- C++ mirror of the reference's host-side operators for the hot path.
//
// The reference's host is Rust (src/ncc.rs, src/main.rs) and cannot be built here (no Rust toolchain),
// so the layer a maintainer would keep -- Searcher, get_hits, process_hits, partition_by, decode_image --
// is mirrored in C++ above the C ABI (include/focr_b200.h), with the reference's names, argument
// meaning and error behaviour.  Where the reference panics (`.unwrap()` on None, `panic!`), these
// throw; the C test hooks at the bottom of focr_host.cpp turn that into FOCR_ERR_ARG + message.
#pragma once
#include <cstdint>
#include <functional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/focr_b200.h"

namespace focr_host {

constexpr size_t MAX_MATCHES = 1024;  // ncc.rs:31

struct RectI {  // pathfinder_geometry::rect::RectI as origin + size
    int32_t x, y, w, h;
};
struct Match {  // ncc.rs:60-64
    RectI rect;
    float similarity;
};
struct MatchWithLetter {  // ncc.rs:74-79
    RectI rect;
    float similarity;
    uint32_t letter;  // char (Unicode scalar value)
    uint32_t index;   // position in get_hits' all_hits (not in the reference; lets callers map results back)
};

struct Panic : std::runtime_error {  // what a Rust panic / unwrap-on-None is mapped to
    using std::runtime_error::runtime_error;
};

// ncc.rs:1036-1052.  Groups are anchored to their FIRST element; panics on empty input (ncc.rs:1040).
template <class T, class Pred>
std::vector<std::pair<size_t, size_t>> partition_by(const T *xs, size_t n, Pred pred)
{
    if (n == 0) throw Panic("called `Option::unwrap()` on a `None` value (partition_by on empty input, ncc.rs:1040)");
    std::vector<std::pair<size_t, size_t>> slices;
    size_t i = 0, j = 0;
    const T *last = &xs[0];
    for (size_t k = 1; k < n; k++) {
        j += 1;
        if (!pred(*last, xs[k])) {
            slices.emplace_back(i, j);
            i = j;
            last = &xs[k];
        }
    }
    slices.emplace_back(i, j + 1);
    return slices;
}

// ncc.rs:723-786 (the --verbose prints are diagnostics and are not mirrored)
std::vector<std::vector<MatchWithLetter>> process_hits(float anchor_threshold, int32_t overlap,
                                                       const std::vector<MatchWithLetter> &all_hits);

// EXTENSION (opt-in; the reference "does not currently detect spaces", README.md:46): gaps between the kept hits of a line
// that are wider than the left glyph's advance are filled with round(excess / space_px) spaces.  `line` is one output
// line of process_hits (x ascending); advance_px(letter) is the glyph's pen advance in pixels (what a renderer moved the
// pen by: ncc's template origin is the same for every letter, so hit x differences ARE pen advances), space_px the
// advance of U+0020.  Hits keep their order; the result is the line's text as code points.  With space_px <= 0 nothing
// is inserted (the reference's output).
std::u32string line_text_with_spaces(const std::vector<MatchWithLetter> &line, const std::function<float(uint32_t)> &advance_px,
                                     float space_px);

// ncc.rs:128-141, 230-404: one page's search state.  search_c_u8 marshals exactly like the reference
// and calls the library's ncc_8_u8 / ncc_16_u8 (the compat shim); the window statistics the reference
// keeps in Searcher (SATs, patch_sum, patch_rnorm, start_end) live on the device instead.
class Searcher {
public:
    Searcher(const uint8_t *gray, uint32_t width, uint32_t height);  // Searcher::new, ncc.rs:231-261
    const std::vector<Match> &search_c_u8(const uint8_t *needle, uint32_t n_w, uint32_t n_h, float threshold);
    uint32_t cols() const { return r_w_; }
    uint32_t rows() const { return r_h_; }

private:
    uint32_t r_w_, r_h_;
    std::vector<uint8_t> reference_u8_;  // image_to_u8, ncc.rs:887-892
    std::vector<uint8_t> needle_u8_;
    std::vector<focr_match> matches_c_;
    std::vector<Match> matches_;
};

// ncc.rs:544-721 restricted to the scan, batched: every template of `bank` against every page in one
// library call.  Returns per page the reference's all_hits (order: template, y, x; ncc.rs:675-681).
struct TemplateMeta {
    uint32_t letter;
    uint16_t n_w, n_h;
};
std::vector<std::vector<MatchWithLetter>> get_hits(focr_ctx *ctx, const focr_bank *bank,
                                                   const std::vector<TemplateMeta> &meta, const uint8_t *pages,
                                                   uint32_t r_w, uint32_t r_h, uint32_t n_pages, float threshold);

// main.rs:35-38 / 220-239
struct DecodedLine {
    std::u32string text;  // alphabet code points
    uint32_t y;
};
std::vector<std::vector<DecodedLine>> decode_image_vec(focr_ctx *ctx, const focr_glyph_bank *bank,
                                                       const std::u32string &alphabet, const uint8_t *pages,
                                                       uint32_t r_w, uint32_t r_h, uint32_t n_pages, uint32_t x_start,
                                                       uint32_t y_start, uint32_t width, uint32_t line_height,
                                                       uint32_t line_advance);

}  // namespace focr_host
