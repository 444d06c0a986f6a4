// postprocess.cu -- process_hits (ncc.rs:723-786) + partition_by (ncc.rs:1036-1052) on the device, for a whole
// batch of pages whose match lists are still resident in HBM.
//
// The reference, per page:
//   (1) keep_y = { y of every hit with similarity >= anchor_threshold }            ncc.rs:727-731
//   (2) keep the hits whose y is in keep_y, in get_hits order                       ncc.rs:732-738
//   (3) stable sort by y; every distinct y is one output line                       ncc.rs:741-747
//   (4) stable sort each line by x                                                  ncc.rs:749-752
//   (5) partition_by(|a.x - b.x| <= overlap), a = FIRST element of the open group   ncc.rs:755-757, 1036-1052
//   (6) per group the LAST hit of maximal similarity (Iterator::max_by)             ncc.rs:761-764
// Two stable sorts on top of get_hits order (template, y, x) are ONE sort by the key (y, x, position in get_hits
// order); with the page in the top bits a single radix sort orders the whole batch.  Steps (5)-(6) are a
// sequential walk along a line, lines are independent: one thread per line.
//
// Kernels: mark_anchor -> collect_keys -> [cub radix sort] -> line_heads -> [cub scan] -> line_starts ->
// dedup_lines -> [cub scan] -> emit.  The sort and scans are CUB device primitives (CUDA toolkit); everything else
// is hand written.  Latency class (microseconds per page): not on the roofline-critical path.
#include <cub/cub.cuh>

#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"

int focr_internal_fail(int code, const std::string &msg);
int focr_internal_device(const focr_ctx *ctx);
extern "C" void *focr_ctx_stream(focr_ctx *ctx);

namespace focr {

// key = page:10 | y:16 | x:16 | orig:22   (orig = t * n_out + i, the hit's position in get_hits order inside its page)
constexpr int PP_ORIG_BITS = 22, PP_X_SHIFT = 22, PP_Y_SHIFT = 38, PP_PAGE_SHIFT = 54;
constexpr uint32_t PP_MAX_PAGES = 1024;

struct PpArgs {
    const focr_match *m;      // [P][T][n_out]
    const uint32_t *counts;   // [P][T]
    uint32_t T, n_out, n_pages;
    float anchor;
    int overlap;
    unsigned char *flag;      // [P][65536] 1 = some hit of this (page, y) reaches the anchor threshold
    unsigned long long *keys; // kept hits
    unsigned int *n_kept;
};

__global__ void __launch_bounds__(256) pp_mark_anchor(PpArgs a)
{
    const uint32_t list = blockIdx.x, page = list / a.T;
    const uint32_t n = min(a.counts[list], a.n_out);
    const focr_match *m = a.m + (size_t)list * a.n_out;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
        if (m[i].similarity >= a.anchor) a.flag[(size_t)page * 65536 + m[i].y] = 1;  // f32 >=, ncc.rs:728
}

__global__ void __launch_bounds__(256) pp_collect_keys(PpArgs a)
{
    const uint32_t list = blockIdx.x, page = list / a.T, t = list - page * a.T;
    const uint32_t n = min(a.counts[list], a.n_out);
    const focr_match *m = a.m + (size_t)list * a.n_out;
    for (uint32_t i0 = 0; i0 < n; i0 += blockDim.x) {
        const uint32_t i = i0 + threadIdx.x;
        bool keep = false;
        unsigned long long key = 0;
        if (i < n) {
            const focr_match h = m[i];
            keep = a.flag[(size_t)page * 65536 + h.y] != 0;
            key = ((unsigned long long)page << PP_PAGE_SHIFT) | ((unsigned long long)h.y << PP_Y_SHIFT) |
                  ((unsigned long long)h.x << PP_X_SHIFT) | (unsigned long long)(t * a.n_out + i);
        }
        // warp-aggregated append (order is irrelevant: the keys are sorted next)
        const unsigned vote = __ballot_sync(0xffffffffu, keep);
        if (vote) {
            const int lane = threadIdx.x & 31;
            unsigned base = 0;
            if (lane == __ffs(vote) - 1) base = atomicAdd(a.n_kept, (unsigned)__popc(vote));
            base = __shfl_sync(0xffffffffu, base, __ffs(vote) - 1);
            if (keep) a.keys[base + __popc(vote & ((1u << lane) - 1u))] = key;
        }
    }
}

// head[k] = 1 when sorted hit k opens a line: first hit, or another (page, y) than its predecessor
__global__ void __launch_bounds__(256) pp_line_heads(const unsigned long long *keys, uint32_t n, uint32_t *head)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) head[k] = (k == 0 || (keys[k] >> PP_Y_SHIFT) != (keys[k - 1] >> PP_Y_SHIFT)) ? 1u : 0u;
}

__global__ void __launch_bounds__(256) pp_line_starts(const uint32_t *head, const uint32_t *line_of_incl, uint32_t n,
                                                      uint32_t *lstart)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n && head[k]) lstart[line_of_incl[k] - 1] = k;
    if (k == n - 1) lstart[line_of_incl[k]] = n;
}

struct PpSel {  // one surviving hit
    uint32_t tpl;
    focr_match m;
};

// one WARP per line: the lanes fetch 32 hits at a time (coalesced key loads, independent match loads), then the warp
// walks them in order -- every lane runs the same scalar logic on values broadcast by shuffle, lane 0 writes.
// Groups are anchored to their first element (ncc.rs:1040-1048); per group the LAST maximum wins (ncc.rs:761-764).
__global__ void __launch_bounds__(128) pp_dedup_lines(PpArgs a, const unsigned long long *keys, const uint32_t *lstart,
                                                      uint32_t n_lines, PpSel *tmp, uint32_t *gcount, uint32_t *line_page)
{
    const uint32_t l = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (l >= n_lines) return;
    const uint32_t s = lstart[l], e = lstart[l + 1];
    const uint32_t page = (uint32_t)(keys[s] >> PP_PAGE_SHIFT);
    if (lane == 0) line_page[l] = page;
    const focr_match *pm = a.m + (size_t)page * a.T * a.n_out;
    uint32_t g = 0;
    int first_x = 0;
    uint32_t best_t = 0, best_xy = 0;
    float best_sim = 0.f;
    bool open = false;
    for (uint32_t base = s; base < e; base += 32) {
        const uint32_t k = base + lane;
        uint32_t t = 0, xy = 0;
        float sim = 0.f;
        if (k < e) {
            const uint32_t orig = (uint32_t)(keys[k] & ((1ull << PP_ORIG_BITS) - 1ull));
            const focr_match h = pm[orig];
            t = orig / a.n_out, xy = (uint32_t)h.x | ((uint32_t)h.y << 16), sim = h.similarity;
        }
        const uint32_t n = min(32u, e - base);
        for (uint32_t j = 0; j < n; j++) {
            const uint32_t ht = __shfl_sync(0xffffffffu, t, j), hxy = __shfl_sync(0xffffffffu, xy, j);
            const float hs = __shfl_sync(0xffffffffu, sim, j);
            const int hx = (int)(hxy & 0xFFFFu);
            if (open && abs(hx - first_x) <= a.overlap) {          // same group (ncc.rs:756)
                if (hs >= best_sim) best_t = ht, best_xy = hxy, best_sim = hs;  // max_by keeps the LAST maximum
            } else {
                if (open && lane == 0) {
                    PpSel o;
                    o.tpl = best_t, o.m.x = (uint16_t)(best_xy & 0xFFFFu), o.m.y = (uint16_t)(best_xy >> 16), o.m.similarity = best_sim;
                    tmp[s + g] = o;
                }
                g += open ? 1u : 0u;
                open = true;
                first_x = hx;
                best_t = ht, best_xy = hxy, best_sim = hs;
            }
        }
    }
    if (lane == 0) {
        PpSel o;
        o.tpl = best_t, o.m.x = (uint16_t)(best_xy & 0xFFFFu), o.m.y = (uint16_t)(best_xy >> 16), o.m.similarity = best_sim;
        tmp[s + g] = o;
        gcount[l] = g + 1;
    }
}

__global__ void __launch_bounds__(128) pp_emit(const PpSel *tmp, const uint32_t *lstart, const uint32_t *gcount,
                                               const uint32_t *out_start, uint32_t n_lines, uint32_t *sel_tpl,
                                               focr_match *sel)
{
    const uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n_lines) return;
    const uint32_t s = lstart[l], o = out_start[l], n = gcount[l];
    for (uint32_t g = 0; g < n; g++) {
        sel_tpl[o + g] = tmp[s + g].tpl;
        sel[o + g] = tmp[s + g].m;
    }
}

// Scratch: grow-only device buffers cached per device (a cudaMalloc / cudaFree pair per buffer and call costs more
// than the kernels, and cudaFree synchronises the device).  Calls on one device are serialised by the pool's mutex.
struct Pool {
    std::mutex mu;
    struct Buf {
        void *p = nullptr;
        size_t cap = 0;
    } bufs[16];
    ~Pool()
    {
        // device memory is reclaimed at process exit; calling cudaFree during static destruction is not safe
    }
};
static Pool g_pools[64];

struct Scratch {
    Pool &pool;
    int next = 0;
    explicit Scratch(Pool &p_) : pool(p_) {}
    template <class T>
    T *get(size_t n)
    {
        Pool::Buf &b = pool.bufs[next++];
        const size_t bytes = std::max<size_t>(n, 1) * sizeof(T);
        if (b.cap < bytes) {
            if (b.p) cudaFree(b.p);
            b.p = nullptr, b.cap = 0;
            const size_t want = bytes + bytes / 4;
            if (cudaMalloc(&b.p, want) != cudaSuccess) return nullptr;
            b.cap = want;
        }
        return (T *)b.p;
    }
};

}  // namespace focr

#define PP_CU(x)                                                                                           \
    do {                                                                                                   \
        cudaError_t e_ = (x);                                                                              \
        if (e_ != cudaSuccess) return focr_internal_fail(FOCR_ERR_CUDA, std::string(#x ": ") + cudaGetErrorString(e_)); \
    } while (0)

extern "C" int focr_process_hits_device(focr_ctx *ctx, const focr_match *matches_dev, const uint32_t *counts_dev, uint32_t T,
                                        uint32_t n_out, uint32_t n_pages, float anchor_threshold, int32_t overlap,
                                        uint32_t line_cap, uint32_t sel_cap, uint32_t *n_lines_out, uint32_t *n_sel_out,
                                        uint32_t *line_page_host, uint32_t *line_start_host, uint32_t *sel_tpl_host,
                                        focr_match *sel_host)
{
    using namespace focr;
    if (!ctx || !matches_dev || !counts_dev || !n_lines_out || !n_sel_out || T == 0 || n_out == 0 || n_pages == 0)
        return focr_internal_fail(FOCR_ERR_ARG, "focr_process_hits_device: NULL or empty argument");
    if (n_pages > PP_MAX_PAGES) return focr_internal_fail(FOCR_ERR_ARG, "focr_process_hits_device: at most 1024 pages per call");
    if ((unsigned long long)T * n_out > (1ull << PP_ORIG_BITS))
        return focr_internal_fail(FOCR_ERR_UNSUPPORTED, "focr_process_hits_device: T * n_out exceeds 2^22");
    PP_CU(cudaSetDevice(focr_internal_device(ctx)));
    cudaStream_t st = (cudaStream_t)focr_ctx_stream(ctx);
    const size_t cap = (size_t)n_pages * T * n_out;  // upper bound of the kept hits
    const int dev = focr_internal_device(ctx);
    if (dev < 0 || dev >= 64) return focr_internal_fail(FOCR_ERR_ARG, "focr_process_hits_device: device index");
    std::lock_guard<std::mutex> lock(g_pools[dev].mu);
    Scratch sc(g_pools[dev]);
    PpArgs a;
    a.m = matches_dev, a.counts = counts_dev, a.T = T, a.n_out = n_out, a.n_pages = n_pages;
    a.anchor = anchor_threshold, a.overlap = overlap;
    a.flag = sc.get<unsigned char>((size_t)n_pages * 65536);
    a.keys = sc.get<unsigned long long>(cap);
    a.n_kept = sc.get<unsigned int>(4);
    unsigned long long *keys_sorted = sc.get<unsigned long long>(cap);
    if (!a.flag || !a.keys || !a.n_kept || !keys_sorted) return focr_internal_fail(FOCR_ERR_NOMEM, "focr_process_hits_device: cudaMalloc");
    PP_CU(cudaMemsetAsync(a.flag, 0, (size_t)n_pages * 65536, st));
    PP_CU(cudaMemsetAsync(a.n_kept, 0, 16, st));
    pp_mark_anchor<<<n_pages * T, 256, 0, st>>>(a);
    pp_collect_keys<<<n_pages * T, 256, 0, st>>>(a);
    PP_CU(cudaGetLastError());
    unsigned int n_kept = 0;
    PP_CU(cudaMemcpyAsync(&n_kept, a.n_kept, 4, cudaMemcpyDeviceToHost, st));
    PP_CU(cudaStreamSynchronize(st));
    *n_lines_out = 0, *n_sel_out = 0;
    if (n_kept == 0) return FOCR_OK;  // no anchor line anywhere (the reference would panic per page: the caller decides)

    // one sort for the batch: (page, y, x, get_hits order)
    size_t tmp_bytes = 0, b2 = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, a.keys, keys_sorted, (int)n_kept, 0, 64, st);
    uint32_t *head = sc.get<uint32_t>(n_kept), *incl = sc.get<uint32_t>(n_kept);
    cub::DeviceScan::InclusiveSum(nullptr, b2, head, incl, (int)n_kept, st);
    tmp_bytes = std::max(tmp_bytes, b2);
    void *cub_tmp = sc.get<unsigned char>(tmp_bytes);
    if (!head || !incl || !cub_tmp) return focr_internal_fail(FOCR_ERR_NOMEM, "focr_process_hits_device: cudaMalloc");
    PP_CU(cub::DeviceRadixSort::SortKeys(cub_tmp, tmp_bytes, a.keys, keys_sorted, (int)n_kept, 0, 64, st));
    const unsigned nb = (n_kept + 255) / 256;
    pp_line_heads<<<nb, 256, 0, st>>>(keys_sorted, n_kept, head);
    PP_CU(cub::DeviceScan::InclusiveSum(cub_tmp, tmp_bytes, head, incl, (int)n_kept, st));
    unsigned int n_lines = 0;
    PP_CU(cudaMemcpyAsync(&n_lines, incl + (n_kept - 1), 4, cudaMemcpyDeviceToHost, st));
    PP_CU(cudaStreamSynchronize(st));
    uint32_t *lstart = sc.get<uint32_t>((size_t)n_lines + 1), *gcount = sc.get<uint32_t>((size_t)n_lines + 1),
             *ostart = sc.get<uint32_t>((size_t)n_lines + 1), *lpage = sc.get<uint32_t>(n_lines);
    PpSel *tmp = sc.get<PpSel>(n_kept);
    if (!lstart || !gcount || !ostart || !lpage || !tmp) return focr_internal_fail(FOCR_ERR_NOMEM, "focr_process_hits_device: cudaMalloc");
    pp_line_starts<<<nb, 256, 0, st>>>(head, incl, n_kept, lstart);
    PP_CU(cudaMemsetAsync(gcount + n_lines, 0, 4, st));
    const unsigned lb = (n_lines + 127) / 128;
    pp_dedup_lines<<<(n_lines + 3) / 4, 128, 0, st>>>(a, keys_sorted, lstart, n_lines, tmp, gcount, lpage);
    size_t b3 = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, b3, gcount, ostart, (int)n_lines + 1, st);
    void *cub_tmp2 = b3 > tmp_bytes ? sc.get<unsigned char>(b3) : cub_tmp;
    if (!cub_tmp2) return focr_internal_fail(FOCR_ERR_NOMEM, "focr_process_hits_device: cudaMalloc");
    size_t b3cap = std::max(b3, tmp_bytes);
    PP_CU(cub::DeviceScan::ExclusiveSum(cub_tmp2, b3cap, gcount, ostart, (int)n_lines + 1, st));
    unsigned int n_sel = 0;
    PP_CU(cudaMemcpyAsync(&n_sel, ostart + n_lines, 4, cudaMemcpyDeviceToHost, st));
    PP_CU(cudaStreamSynchronize(st));
    *n_lines_out = n_lines, *n_sel_out = n_sel;
    if (n_lines > line_cap || n_sel > sel_cap || !line_page_host || !line_start_host || !sel_tpl_host || !sel_host)
        return focr_internal_fail(FOCR_ERR_NOMEM, "focr_process_hits_device: output capacity too small (see n_lines / n_sel)");
    uint32_t *sel_tpl = sc.get<uint32_t>(n_sel);
    focr_match *sel = sc.get<focr_match>(n_sel);
    if (!sel_tpl || !sel) return focr_internal_fail(FOCR_ERR_NOMEM, "focr_process_hits_device: cudaMalloc");
    pp_emit<<<lb, 128, 0, st>>>(tmp, lstart, gcount, ostart, n_lines, sel_tpl, sel);
    PP_CU(cudaGetLastError());
    PP_CU(cudaMemcpyAsync(line_page_host, lpage, (size_t)n_lines * 4, cudaMemcpyDeviceToHost, st));
    PP_CU(cudaMemcpyAsync(line_start_host, ostart, ((size_t)n_lines + 1) * 4, cudaMemcpyDeviceToHost, st));
    PP_CU(cudaMemcpyAsync(sel_tpl_host, sel_tpl, (size_t)n_sel * 4, cudaMemcpyDeviceToHost, st));
    PP_CU(cudaMemcpyAsync(sel_host, sel, (size_t)n_sel * sizeof(focr_match), cudaMemcpyDeviceToHost, st));
    PP_CU(cudaStreamSynchronize(st));
    return FOCR_OK;
}
