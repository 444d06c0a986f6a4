// focr_host.cpp -- see focr_host.hpp.  Compiled into libfocr_b200.so next to the kernels.
#include "focr_host.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <unordered_set>

int focr_internal_fail(int code, const std::string &msg);  // api.cu

namespace focr_host {

// f32::total_cmp (ncc.rs:763): IEEE total order on the bit patterns
static inline int32_t total_key(float f)
{
    int32_t b;
    memcpy(&b, &f, 4);
    b ^= (int32_t)(((uint32_t)(b >> 31)) >> 1);
    return b;
}

std::vector<std::vector<MatchWithLetter>> process_hits(float anchor_threshold, int32_t overlap,
                                                       const std::vector<MatchWithLetter> &all_hits)
{
    std::unordered_set<int32_t> keep_y;  // ncc.rs:726-731
    keep_y.reserve(512);
    for (const auto &h : all_hits)
        if (h.similarity >= anchor_threshold) keep_y.insert(h.rect.y);
    std::vector<MatchWithLetter> hits;  // ncc.rs:732-738
    for (const auto &h : all_hits)
        if (keep_y.count(h.rect.y)) hits.push_back(h);
    std::stable_sort(hits.begin(), hits.end(),  // slice::sort_by_key is stable, ncc.rs:741
                     [](const MatchWithLetter &a, const MatchWithLetter &b) { return a.rect.y < b.rect.y; });
    auto line_slices = partition_by(hits.data(), hits.size(),  // ncc.rs:747: lines are runs of EXACTLY equal y
                                    [](const MatchWithLetter &a, const MatchWithLetter &b) { return a.rect.y == b.rect.y; });
    for (auto &s : line_slices)  // ncc.rs:749-752
        std::stable_sort(hits.begin() + s.first, hits.begin() + s.second,
                         [](const MatchWithLetter &a, const MatchWithLetter &b) { return a.rect.x < b.rect.x; });
    std::vector<std::vector<MatchWithLetter>> lines;
    for (auto &s : line_slices) {
        const MatchWithLetter *slice = hits.data() + s.first;
        const size_t n = s.second - s.first;
        auto dups = partition_by(slice, n, [overlap](const MatchWithLetter &a, const MatchWithLetter &b) {
            return std::abs(a.rect.x - b.rect.x) <= overlap;  // ncc.rs:755-757
        });
        std::vector<MatchWithLetter> dedup;
        for (auto &d : dups) {
            // Iterator::max_by returns the LAST element among equal maxima (ncc.rs:761-764)
            const MatchWithLetter *best = &slice[d.first];
            for (size_t k = d.first + 1; k < d.second; k++)
                if (total_key(slice[k].similarity) >= total_key(best->similarity)) best = &slice[k];
            dedup.push_back(*best);
        }
        lines.push_back(std::move(dedup));
    }
    return lines;
}

std::u32string line_text_with_spaces(const std::vector<MatchWithLetter> &line, const std::function<float(uint32_t)> &advance_px,
                                     float space_px)
{
    std::u32string out;
    for (size_t i = 0; i < line.size(); i++) {
        if (i > 0 && space_px > 0.f) {
            // pen travel between the two hits minus what the left glyph itself advances = room taken by spaces
            const float excess = (float)(line[i].rect.x - line[i - 1].rect.x) - advance_px(line[i - 1].letter);
            const int k = excess > 0.f ? (int)std::floor(excess / space_px + 0.5f) : 0;
            out.append((size_t)k, U' ');
        }
        out.push_back((char32_t)line[i].letter);
    }
    return out;
}

Searcher::Searcher(const uint8_t *gray, uint32_t width, uint32_t height)
    : r_w_(width), r_h_(height), reference_u8_((size_t)width * height), needle_u8_(128), matches_c_(MAX_MATCHES)
{
    for (size_t i = 0; i < reference_u8_.size(); i++) reference_u8_[i] = (uint8_t)(255 - gray[i]);  // ncc.rs:888
    matches_.reserve(1024);
}

const std::vector<Match> &Searcher::search_c_u8(const uint8_t *needle, uint32_t n_w, uint32_t n_h, float threshold)
{
    size_t N;
    if (n_w <= 8)
        N = 8;
    else if (n_w <= 16)
        N = 16;
    else
        throw Panic("not handled");  // ncc.rs:392
    needle_u8_.assign(N * n_h, 0);   // ncc.rs:340,367 + copy_needle_n_u8 ncc.rs:925-935
    for (uint32_t y = 0; y < n_h; y++) memcpy(&needle_u8_[y * N], needle + (size_t)y * n_w, n_w);
    auto fn = N == 8 ? ncc_8_u8 : ncc_16_u8;
    const size_t n_matches = fn(reference_u8_.data(), r_w_, r_h_, needle_u8_.data(), n_w, n_h, nullptr, 0, nullptr,
                                nullptr, nullptr, threshold, matches_c_.data(), matches_c_.size());
    if (n_matches == MAX_MATCHES) fprintf(stderr, "WARN got >= %zu matches\n", n_matches);  // ncc.rs:395-397
    matches_.clear();
    for (size_t i = 0; i < n_matches; i++)  // Match::from_matchc, ncc.rs:81-90
        matches_.push_back(Match{RectI{(int32_t)matches_c_[i].x, (int32_t)matches_c_[i].y, (int32_t)n_w, (int32_t)n_h},
                                 matches_c_[i].similarity});
    return matches_;
}

std::vector<std::vector<MatchWithLetter>> get_hits(focr_ctx *ctx, const focr_bank *bank,
                                                   const std::vector<TemplateMeta> &meta, const uint8_t *pages,
                                                   uint32_t r_w, uint32_t r_h, uint32_t n_pages, float threshold)
{
    const uint32_t T = focr_bank_size(bank);
    if (meta.size() != T) throw Panic("get_hits: one TemplateMeta per bank template expected");
    std::vector<focr_match> out((size_t)n_pages * T * MAX_MATCHES);
    std::vector<uint32_t> counts((size_t)n_pages * T);
    const int rc = focr_ncc_scan(ctx, bank, pages, (size_t)r_w * r_h, r_w, r_h, n_pages, threshold, MAX_MATCHES,
                                 out.data(), counts.data());
    if (rc != FOCR_OK) throw Panic(std::string("focr_ncc_scan: ") + focr_last_error());
    std::vector<std::vector<MatchWithLetter>> res(n_pages);
    for (uint32_t p = 0; p < n_pages; p++) {
        auto &all_hits = res[p];
        for (uint32_t t = 0; t < T; t++) {  // ncc.rs:587,630: offsets outer, letters inner == bank order
            const focr_match *m = &out[((size_t)p * T + t) * MAX_MATCHES];
            for (uint32_t k = 0; k < counts[(size_t)p * T + t]; k++)  // ncc.rs:675-681
                all_hits.push_back(MatchWithLetter{RectI{m[k].x, m[k].y, meta[t].n_w, meta[t].n_h}, m[k].similarity,
                                                   meta[t].letter, (uint32_t)all_hits.size()});
        }
    }
    return res;
}

std::vector<std::vector<DecodedLine>> decode_image_vec(focr_ctx *ctx, const focr_glyph_bank *bank,
                                                       const std::u32string &alphabet, const uint8_t *pages,
                                                       uint32_t r_w, uint32_t r_h, uint32_t n_pages, uint32_t x_start,
                                                       uint32_t y_start, uint32_t width, uint32_t line_height,
                                                       uint32_t line_advance)
{
    const uint32_t max_lines = std::max<uint32_t>(1, y_start >= r_h ? 1 : (r_h - y_start + line_advance - 1) / line_advance);
    const uint32_t max_cells = 1024;
    std::vector<uint16_t> glyphs((size_t)n_pages * max_lines * max_cells);
    std::vector<uint32_t> n_cells((size_t)n_pages * max_lines), line_y((size_t)n_pages * max_lines), n_lines(n_pages);
    const int rc = focr_decode_pages(ctx, bank, pages, (size_t)r_w * r_h, r_w, r_h, n_pages, x_start, y_start, width,
                                     line_height, line_advance, max_lines, max_cells, glyphs.data(), n_cells.data(),
                                     line_y.data(), n_lines.data());
    if (rc != FOCR_OK) throw Panic(std::string("focr_decode_pages: ") + focr_last_error());
    std::vector<std::vector<DecodedLine>> res(n_pages);
    for (uint32_t p = 0; p < n_pages; p++)
        for (uint32_t l = 0; l < n_lines[p]; l++) {
            DecodedLine d;
            d.y = line_y[(size_t)p * max_lines + l];
            const uint16_t *g = &glyphs[((size_t)p * max_lines + l) * max_cells];
            for (uint32_t k = 0; k < n_cells[(size_t)p * max_lines + l]; k++) d.text.push_back(alphabet.at(g[k]));
            res[p].push_back(std::move(d));
        }
    return res;
}

}  // namespace focr_host

// ------------------------------------------------------------------------------------------------
// C hooks (include/focr_b200.h section 4): let the ctypes tests drive the C++ host mirror.
using namespace focr_host;

extern "C" int focr_host_line_text_with_spaces(const int32_t *xs, const uint32_t *letters, uint32_t n, const uint32_t *adv_letters,
                                               const float *adv_px, uint32_t n_adv, float space_px, uint32_t *out, uint32_t out_cap,
                                               uint32_t *n_out)
{
    if ((n && (!xs || !letters)) || (n_adv && (!adv_letters || !adv_px)) || !out || !n_out)
        return focr_internal_fail(FOCR_ERR_ARG, "focr_host_line_text_with_spaces: NULL argument");
    std::vector<MatchWithLetter> line(n);
    for (uint32_t i = 0; i < n; i++) line[i] = MatchWithLetter{RectI{xs[i], 0, 0, 0}, 0.f, letters[i], i};
    auto adv = [&](uint32_t letter) -> float {
        for (uint32_t k = 0; k < n_adv; k++)
            if (adv_letters[k] == letter) return adv_px[k];
        return 0.f;
    };
    const std::u32string t = line_text_with_spaces(line, adv, space_px);
    *n_out = (uint32_t)t.size();
    if (t.size() > out_cap) return focr_internal_fail(FOCR_ERR_NOMEM, "focr_host_line_text_with_spaces: out_cap too small (see n_out)");
    for (size_t i = 0; i < t.size(); i++) out[i] = (uint32_t)t[i];
    return FOCR_OK;
}

extern "C" int focr_host_process_hits(const int32_t *xs, const int32_t *ys, const float *sims, const uint32_t *letters,
                                      uint32_t n, float anchor_threshold, int32_t overlap, uint32_t *out_index,
                                      uint32_t *line_offsets, uint32_t *n_lines)
{
    if ((n && (!xs || !ys || !sims || !letters)) || !out_index || !line_offsets || !n_lines)
        return focr_internal_fail(FOCR_ERR_ARG, "focr_host_process_hits: NULL argument");
    std::vector<MatchWithLetter> all(n);
    for (uint32_t i = 0; i < n; i++) all[i] = MatchWithLetter{RectI{xs[i], ys[i], 0, 0}, sims[i], letters[i], i};
    try {
        auto lines = process_hits(anchor_threshold, overlap, all);
        uint32_t k = 0;
        line_offsets[0] = 0;
        for (size_t l = 0; l < lines.size(); l++) {
            for (auto &m : lines[l]) out_index[k++] = m.index;
            line_offsets[l + 1] = k;
        }
        *n_lines = (uint32_t)lines.size();
    } catch (const Panic &e) {
        return focr_internal_fail(FOCR_ERR_ARG, std::string("panic: ") + e.what());
    }
    return FOCR_OK;
}

extern "C" int focr_host_search_c_u8(const uint8_t *gray, uint32_t r_w, uint32_t r_h, const uint8_t *needle,
                                     uint32_t n_w, uint32_t n_h, float threshold, focr_match *out, uint32_t *n_out)
{
    if (!gray || !needle || !out || !n_out) return focr_internal_fail(FOCR_ERR_ARG, "focr_host_search_c_u8: NULL argument");
    try {
        Searcher s(gray, r_w, r_h);
        const auto &m = s.search_c_u8(needle, n_w, n_h, threshold);
        for (size_t i = 0; i < m.size(); i++)
            out[i] = focr_match{(uint16_t)m[i].rect.x, (uint16_t)m[i].rect.y, m[i].similarity};
        *n_out = (uint32_t)m.size();
    } catch (const Panic &e) {
        return focr_internal_fail(FOCR_ERR_UNSUPPORTED, std::string("panic: ") + e.what());
    }
    return FOCR_OK;
}
