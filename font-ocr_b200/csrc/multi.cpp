// multi.cpp -- one process, N GPUs: the page-parallel driver of the reference (`pages.par_iter().map(..)` followed by
// `sort_by_key(page index)`, ncc.rs:839-847; main.rs:443-468) over the single-device C ABI.
//
// Pages are independent units: every device gets a contiguous block of the batch, the template /
// glyph bank is replicated to every device once, each device runs its own pipeline (its own streams, pinned staging and
// chunk schedule) from its own host thread, and the "gather by page index" is simply that every device writes the match
// lists of its block to the block's place in the caller's arrays.  Nothing is exchanged or reduced between GPUs: no
// collective, no peer traffic.  Only public entries of include/focr_b200.h are used here.
#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "../../include/focr_b200.h"

int focr_internal_fail(int code, const std::string &msg);   // api.cu: sets the calling thread's focr_last_error()
void focr_internal_stage_sharers(int n);                    // api.cu: how many GPUs share the host's staging threads right now

struct focr_multi {
    std::vector<focr_ctx *> ctx;
    std::vector<int> device;
};
struct focr_multi_bank {
    focr_multi *m = nullptr;
    std::vector<focr_bank *> bank;
    uint32_t T = 0;
};
struct focr_multi_glyph_bank {
    focr_multi *m = nullptr;
    std::vector<focr_glyph_bank *> bank;
};

namespace {
// contiguous block of device i: sizes differ by at most one, lower devices first (shard.shard_range in the Python mirror)
void block_of(uint32_t n_pages, uint32_t i, uint32_t n_dev, uint32_t &p0, uint32_t &np)
{
    const uint32_t base = n_pages / n_dev, extra = n_pages % n_dev;
    p0 = i * base + std::min(i, extra);
    np = base + (i < extra ? 1u : 0u);
}

// run fn(i) for every device on its own host thread; the first failure (lowest device index) is reported on the caller's thread
template <class F>
int for_each_device(uint32_t n_dev, F fn)
{
    std::vector<int> rc(n_dev, FOCR_OK);
    std::vector<std::string> err(n_dev);
    auto run = [&](uint32_t i) {
        rc[i] = fn(i);
        if (rc[i] != FOCR_OK) err[i] = focr_last_error();   // thread-local: fetch it on the worker
    };
    std::vector<std::thread> th;
    for (uint32_t i = 1; i < n_dev; i++) th.emplace_back(run, i);
    run(0);
    for (auto &t : th) t.join();
    for (uint32_t i = 0; i < n_dev; i++)
        if (rc[i] != FOCR_OK) return focr_internal_fail(rc[i], "device slot " + std::to_string(i) + ": " + err[i]);
    return FOCR_OK;
}
}  // namespace

extern "C" int focr_multi_create(const int *devices, uint32_t n_devices, focr_multi **out)
{
    if (!out) return focr_internal_fail(FOCR_ERR_ARG, "focr_multi_create: out is NULL");
    std::vector<int> dev;
    if (devices) {
        if (n_devices == 0) return focr_internal_fail(FOCR_ERR_ARG, "focr_multi_create: empty device list");
        dev.assign(devices, devices + n_devices);
    } else {
        for (uint32_t i = 0; n_devices == 0 || i < n_devices; i++) {
            if (n_devices == 0) {   // every visible device: create contexts until the index runs out
                focr_ctx *c = nullptr;
                if (focr_ctx_create((int)i, &c) != FOCR_OK) break;
                focr_ctx_destroy(c);
            }
            dev.push_back((int)i);
        }
        if (dev.empty()) return FOCR_ERR_CUDA;   // focr_last_error() holds focr_ctx_create's message
    }
    focr_multi *m = new focr_multi();
    for (int d : dev) {
        focr_ctx *c = nullptr;
        const int rc = focr_ctx_create(d, &c);
        if (rc != FOCR_OK) {
            for (focr_ctx *x : m->ctx) focr_ctx_destroy(x);
            delete m;
            return rc;
        }
        m->ctx.push_back(c);
        m->device.push_back(d);
    }
    *out = m;
    return FOCR_OK;
}

extern "C" void focr_multi_destroy(focr_multi *m)
{
    if (!m) return;
    for (focr_ctx *c : m->ctx) focr_ctx_destroy(c);
    delete m;
}

extern "C" uint32_t focr_multi_size(const focr_multi *m) { return m ? (uint32_t)m->ctx.size() : 0; }
extern "C" focr_ctx *focr_multi_ctx(focr_multi *m, uint32_t i) { return (m && i < m->ctx.size()) ? m->ctx[i] : nullptr; }

extern "C" void focr_multi_page_block(const focr_multi *m, uint32_t n_pages, uint32_t i, uint32_t *first, uint32_t *count)
{
    uint32_t p0 = 0, np = 0;
    if (m && i < m->ctx.size()) block_of(n_pages, i, (uint32_t)m->ctx.size(), p0, np);
    if (first) *first = p0;
    if (count) *count = np;
}

extern "C" int focr_multi_bank_create(focr_multi *m, const uint8_t *pixels, const uint64_t *offsets, const uint16_t *n_w,
                                      const uint16_t *n_h, uint32_t n_templates, focr_multi_bank **out)
{
    if (!m || !out) return focr_internal_fail(FOCR_ERR_ARG, "focr_multi_bank_create: NULL argument");
    focr_multi_bank *b = new focr_multi_bank();
    b->m = m;
    b->T = n_templates;
    b->bank.assign(m->ctx.size(), nullptr);
    const int rc = for_each_device((uint32_t)m->ctx.size(), [&](uint32_t i) {
        return focr_bank_create(m->ctx[i], pixels, offsets, n_w, n_h, n_templates, &b->bank[i]);
    });
    if (rc != FOCR_OK) {
        for (focr_bank *x : b->bank) focr_bank_destroy(x);
        delete b;
        return rc;
    }
    *out = b;
    return FOCR_OK;
}

extern "C" void focr_multi_bank_destroy(focr_multi_bank *b)
{
    if (!b) return;
    for (focr_bank *x : b->bank) focr_bank_destroy(x);
    delete b;
}

extern "C" int focr_multi_ncc_scan(focr_multi *m, const focr_multi_bank *b, const uint8_t *pages_host, size_t page_stride,
                                   uint32_t r_w, uint32_t r_h, uint32_t n_pages, float threshold, uint32_t n_out,
                                   focr_match *out_host, uint32_t *counts_host)
{
    if (!m || !b || b->m != m || !pages_host || !out_host || !counts_host || n_pages == 0)
        return focr_internal_fail(FOCR_ERR_ARG, "focr_multi_ncc_scan: NULL argument, foreign bank or no pages");
    const uint32_t n_dev = (uint32_t)std::min<size_t>(m->ctx.size(), n_pages);
    const size_t T = b->T;
    focr_internal_stage_sharers((int)n_dev);   // pageable callers: the devices share the host's staging threads
    const int rc = for_each_device(n_dev, [&](uint32_t i) {
        uint32_t p0, np;
        block_of(n_pages, i, n_dev, p0, np);
        return focr_ncc_scan(m->ctx[i], b->bank[i], pages_host + (size_t)p0 * page_stride, page_stride, r_w, r_h, np, threshold,
                             n_out, out_host + (size_t)p0 * T * n_out, counts_host + (size_t)p0 * T);
    });
    focr_internal_stage_sharers(1);
    return rc;
}

extern "C" int focr_multi_glyph_bank_create(focr_multi *m, const uint8_t *pixels, size_t n_pixel_bytes,
                                            const focr_glyph_raster *rasters, const float *advance_px, uint32_t n_glyphs,
                                            int32_t origin_x, focr_multi_glyph_bank **out)
{
    if (!m || !out) return focr_internal_fail(FOCR_ERR_ARG, "focr_multi_glyph_bank_create: NULL argument");
    focr_multi_glyph_bank *b = new focr_multi_glyph_bank();
    b->m = m;
    b->bank.assign(m->ctx.size(), nullptr);
    const int rc = for_each_device((uint32_t)m->ctx.size(), [&](uint32_t i) {
        return focr_glyph_bank_create(m->ctx[i], pixels, n_pixel_bytes, rasters, advance_px, n_glyphs, origin_x, &b->bank[i]);
    });
    if (rc != FOCR_OK) {
        for (focr_glyph_bank *x : b->bank) focr_glyph_bank_destroy(x);
        delete b;
        return rc;
    }
    *out = b;
    return FOCR_OK;
}

extern "C" void focr_multi_glyph_bank_destroy(focr_multi_glyph_bank *b)
{
    if (!b) return;
    for (focr_glyph_bank *x : b->bank) focr_glyph_bank_destroy(x);
    delete b;
}

extern "C" int focr_multi_decode_pages(focr_multi *m, const focr_multi_glyph_bank *b, const uint8_t *pages_host,
                                       size_t page_stride, uint32_t r_w, uint32_t r_h, uint32_t n_pages, uint32_t x_start,
                                       uint32_t y_start, uint32_t width, uint32_t line_height, uint32_t line_advance,
                                       uint32_t max_lines, uint32_t max_cells, uint16_t *glyphs_host, uint32_t *n_cells_host,
                                       uint32_t *line_y_host, uint32_t *n_lines_host)
{
    if (!m || !b || b->m != m || !pages_host || n_pages == 0)
        return focr_internal_fail(FOCR_ERR_ARG, "focr_multi_decode_pages: NULL argument, foreign bank or no pages");
    const uint32_t n_dev = (uint32_t)std::min<size_t>(m->ctx.size(), n_pages);
    return for_each_device(n_dev, [&](uint32_t i) {
        uint32_t p0, np;
        block_of(n_pages, i, n_dev, p0, np);
        const size_t l0 = (size_t)p0 * max_lines;
        return focr_decode_pages(m->ctx[i], b->bank[i], pages_host + (size_t)p0 * page_stride, page_stride, r_w, r_h, np, x_start,
                                 y_start, width, line_height, line_advance, max_lines, max_cells,
                                 glyphs_host ? glyphs_host + l0 * max_cells : nullptr, n_cells_host ? n_cells_host + l0 : nullptr,
                                 line_y_host ? line_y_host + l0 : nullptr, n_lines_host ? n_lines_host + p0 : nullptr);
    });
}
