// api.cu -- the C ABI of libfocr_b200.so (include/focr_b200.h): contexts, template banks, the
// batched scan pipeline (H2D -> stage -> stats -> scan -> finalize -> D2H) and the compat shim that
// exports the reference's own `ncc_8_u8` / `ncc_16_u8` symbols (ncc.cpp:48-63, 253-268).
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "common.cuh"
#include "kernels.cuh"
#include "scan_tc.cuh"
#include "focr_decode.cuh"

using namespace focr;

// ------------------------------------------------------------------------------------------ errors
static thread_local std::string g_err;
static int fail(int code, const std::string &msg)
{
    g_err = msg;
    // a CUDA error that is reported here must not surface again in the cudaGetLastError() check of the next launch
    // (the runtime keeps a non-sticky error as "last error" until somebody reads it)
    if (code == FOCR_ERR_CUDA) cudaGetLastError();
    return code;
}
// shared with focr_decode.cu
int focr_internal_fail(int code, const std::string &msg) { return fail(code, msg); }
#define CU(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess)                                                                       \
            return fail(FOCR_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));          \
    } while (0)

extern "C" const char *focr_last_error(void) { return g_err.c_str(); }
extern "C" const char *focr_version(void) { return "focr_b200 0.1 (sm_100a)"; }
extern "C" void focr_get_limits(focr_limits *out)
{
    if (!out) return;
    out->max_template_w = MAX_TPL_W;
    out->max_template_h = MAX_TPL_H;
    out->max_page_w = 15360;  // finalize_sel_cap: n_out + r_w keys must fit the shared-memory sort
    out->max_page_h = 65535 - PAGE_PAD_ROWS;  // stage_invert: grid.y = r_h + PAGE_PAD_ROWS <= 65535
    out->max_n_out = 4096;
}

// ------------------------------------------------------------------------------------------ buffers
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) {
            cudaError_t e = cudaFree(p);
            if (e != cudaSuccess) return e;
            p = nullptr;
            cap = 0;
        }
        size_t want = bytes + bytes / 8;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            p = nullptr;
            return e;
        }
        cap = want;
        return cudaSuccess;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T>
    T *as() const { return (T *)p; }
};

struct PinBuf {  // library-owned pinned staging (grow-only)
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr, cap = 0;
        const size_t want = bytes + bytes / 8;
        cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocDefault);
        if (e != cudaSuccess) {
            p = nullptr;
            return e;
        }
        cap = want;
        return cudaSuccess;
    }
    void release()
    {
        if (p) cudaFreeHost(p);
        p = nullptr, cap = 0;
    }
    template <class T>
    T *as() const { return (T *)p; }
};

constexpr int FLAG_WORDS = 16;    // one 64-byte flag block per chunk
constexpr int FLAG_BLOCKS = 64;   // chunks the device-resident scan may enqueue between two host checks

struct Slot {  // everything one in-flight chunk of pages needs
    DevBuf gray, inv, sp, s2p, pf, rn, sp2, pf2, rowcount, hits, cands, candcnt, sel, ycut, flags, out, counts, acc;
    PinBuf gray_pin, out_pin, counts_pin;   // staging for callers whose host buffers are pageable (ncc.rs:575: a Rust Vec<u8>)
    unsigned int *flags_host = nullptr;  // pinned, FLAG_BLOCKS blocks of: [0] hit_count, [1] overflow, [2] cand_count,
                                         // [3] cand high-water mark, [4..9] scan_tc watchdog (raised, tag, info, CTA, warp, parity)
    cudaEvent_t ev_h2d = nullptr, ev_compute = nullptr, ev_d2h = nullptr;
    uint32_t hits_per_page = 0;
    uint32_t reserve_pages = 0;   // size the buffers for this many pages at once (a ramp of growing chunks would re-allocate)
};

struct focr_ctx {
    int device = 0;
    cudaStream_t stream = nullptr, h2d = nullptr, d2h = nullptr;
    int kernel = FOCR_KERNEL_AUTO;
    uint64_t launches = 0;
    Slot slot[2];
    uint32_t hits_per_page = 2u << 20;
    uint32_t cand_per_warp = 16384;  // capacity of each epilogue warp's private candidate list (grows on overflow)
    int sm_count = 148;
    // per-stage profiling (focr_ctx_profile)
    bool profile = false;
    struct Span { cudaEvent_t a, b; int stage; int launches; };
    std::vector<Span> spans;
    std::vector<cudaEvent_t> event_pool;
    cudaEvent_t get_event()
    {
        if (!event_pool.empty()) { cudaEvent_t e = event_pool.back(); event_pool.pop_back(); return e; }
        cudaEvent_t e = nullptr;
        cudaEventCreate(&e);
        return e;
    }
};

struct StageTimer {  // RAII: events around one stage when profiling is on
    focr_ctx *c; int stage; int launches; cudaEvent_t a = nullptr;
    StageTimer(focr_ctx *c_, int stage_, int launches_ = 1) : c(c_), stage(stage_), launches(launches_)
    {
        if (c->profile) { a = c->get_event(); cudaEventRecord(a, c->stream); }
    }
    ~StageTimer()
    {
        if (a) { cudaEvent_t b = c->get_event(); cudaEventRecord(b, c->stream); c->spans.push_back({a, b, stage, launches}); }
    }
};

void focr_internal_streams(focr_ctx *c, cudaStream_t out[3]) { out[0] = c->stream, out[1] = c->h2d, out[2] = c->d2h; }
// focr_decode.cu times its kernel through these (the StageTimer lives here)
void *focr_internal_stage_begin(focr_ctx *c, int stage) { return new StageTimer(c, stage); }
void focr_internal_stage_end(void *t) { delete (StageTimer *)t; }

struct ExactStage : TcHook {  // the exact pass as its own profiling stage (inside FOCR_STAGE_SCAN)
    focr_ctx *c; StageTimer *t = nullptr;
    explicit ExactStage(focr_ctx *c_) : c(c_) {}
    void exact_begin() override { t = new StageTimer(c, FOCR_STAGE_EXACT); }
    void exact_end() override { delete t; t = nullptr; }
};

struct ClassHost {
    uint32_t n_w, n_h, np;
    std::vector<uint32_t> index;
    DevBuf rows, index_dev;
    int group = -1;       // launch group of the tcgen05 kernel this box size belongs to
    uint32_t group_pos = 0;  // index of the class's first template within the group
};

// launch group of the tcgen05 kernel (scan_tc.cu): one box size, or two box sizes of the same height that share the
// correlation GEMM (their statistics ride in the two K chunks of the normalisation MMA)
struct GroupHost {
    int cls[2] = {-1, -1};
    TcClass tc;
};

struct focr_bank {
    focr_ctx *ctx = nullptr;
    uint32_t T = 0;
    std::vector<ClassHost> classes;
    std::vector<GroupHost> groups;
    std::vector<TplInfo> info;
    DevBuf info_dev;
};

// ------------------------------------------------------------------------------------------ context
extern "C" int focr_ctx_create(int device, focr_ctx **out)
{
    if (!out) return fail(FOCR_ERR_ARG, "focr_ctx_create: out is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(FOCR_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e) +
                                       " (libfocr_b200 has no CPU fallback)");
    if (device < 0 || device >= n) return fail(FOCR_ERR_ARG, "focr_ctx_create: bad device index");
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(FOCR_ERR_UNSUPPORTED, std::string("device is sm_") + std::to_string(prop.major * 10 + prop.minor) +
                                              "; libfocr_b200 is built for sm_100a only");
    focr_ctx *c = new focr_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&c->h2d, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&c->d2h, cudaStreamNonBlocking));
    for (auto &s : c->slot) {
        CU(cudaMallocHost((void **)&s.flags_host, FLAG_BLOCKS * FLAG_WORDS * 4));
        memset(s.flags_host, 0, FLAG_BLOCKS * FLAG_WORDS * 4);
        CU(cudaEventCreateWithFlags(&s.ev_h2d, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&s.ev_compute, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&s.ev_d2h, cudaEventDisableTiming));
    }
    *out = c;
    return FOCR_OK;
}

extern "C" void focr_ctx_destroy(focr_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (auto &s : c->slot) {
        for (DevBuf *b : {&s.gray, &s.inv, &s.sp, &s.s2p, &s.pf, &s.rn, &s.sp2, &s.pf2, &s.rowcount, &s.hits, &s.cands, &s.candcnt, &s.sel, &s.ycut,
                          &s.flags, &s.out, &s.counts, &s.acc})
            b->release();
        for (PinBuf *b : {&s.gray_pin, &s.out_pin, &s.counts_pin}) b->release();
        if (s.flags_host) cudaFreeHost(s.flags_host);
        if (s.ev_h2d) cudaEventDestroy(s.ev_h2d);
        if (s.ev_compute) cudaEventDestroy(s.ev_compute);
        if (s.ev_d2h) cudaEventDestroy(s.ev_d2h);
    }
    cudaStreamDestroy(c->stream);
    cudaStreamDestroy(c->h2d);
    cudaStreamDestroy(c->d2h);
    delete c;
}

extern "C" int focr_ctx_set_kernel(focr_ctx *c, int kernel)
{
    if (!c || kernel < FOCR_KERNEL_AUTO || kernel > FOCR_KERNEL_TCGEN05)
        return fail(FOCR_ERR_ARG, "focr_ctx_set_kernel: bad argument");
    c->kernel = kernel;
    return FOCR_OK;
}
extern "C" void *focr_ctx_stream(focr_ctx *c) { return c ? (void *)c->stream : nullptr; }
extern "C" int focr_ctx_sync(focr_ctx *c)
{
    if (!c) return fail(FOCR_ERR_ARG, "focr_ctx_sync: NULL");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->h2d));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaStreamSynchronize(c->d2h));
    return FOCR_OK;
}
extern "C" uint64_t focr_ctx_launch_count(const focr_ctx *c) { return c ? c->launches : 0; }
int focr_internal_device(const focr_ctx *c) { return c->device; }
void focr_internal_count_launch(focr_ctx *c, int n) { c->launches += n; }
extern "C" int focr_ctx_profile(focr_ctx *c, int enable)
{
    if (!c) return fail(FOCR_ERR_ARG, "focr_ctx_profile: NULL");
    c->profile = enable != 0;
    return FOCR_OK;
}
extern "C" int focr_ctx_profile_read(focr_ctx *c, double *ms_out, uint64_t *launches_out)
{
    if (!c || !ms_out || !launches_out) return fail(FOCR_ERR_ARG, "focr_ctx_profile_read: NULL");
    int rc = focr_ctx_sync(c);
    if (rc) return rc;
    for (int i = 0; i < FOCR_N_STAGES; i++) { ms_out[i] = 0; launches_out[i] = 0; }
    for (auto &sp : c->spans) {
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, sp.a, sp.b));
        ms_out[sp.stage] += ms;
        launches_out[sp.stage] += sp.launches;
        c->event_pool.push_back(sp.a);
        c->event_pool.push_back(sp.b);
    }
    c->spans.clear();
    return FOCR_OK;
}

// ------------------------------------------------------------------------------------------ bank
static void tpl_info(const uint8_t *px, uint32_t n_w, uint32_t n_h, TplInfo &ti)
{
    // ncc.cpp:73-86 (the zero padding adds nothing to either sum)
    uint32_t s_n = 0, s2_n = 0;
    for (uint32_t i = 0; i < n_w * n_h; i++) {
        s_n += px[i];
        s2_n += (uint32_t)px[i] * (uint32_t)px[i];
    }
    const size_t n = (size_t)n_w * n_h;
    volatile double q = (double)((uint64_t)s_n * (uint64_t)s_n) / (double)n;  // volatile: no contraction
    const double norm2_n = (double)s2_n - q;
    ti.rnorm_n = 1. / std::sqrt(norm2_n);
    ti.n_recip = 1. / (double)n;
    ti.s_n = (double)s_n;
    ti.n_w = n_w;
    ti.n_h = n_h;
}

extern "C" int focr_bank_create(focr_ctx *c, const uint8_t *pixels, const uint64_t *offsets, const uint16_t *n_w,
                                const uint16_t *n_h, uint32_t T, focr_bank **out)
{
    if (!c || !pixels || !offsets || !n_w || !n_h || !out || T == 0)
        return fail(FOCR_ERR_ARG, "focr_bank_create: NULL argument or empty bank");
    CU(cudaSetDevice(c->device));
    focr_bank *b = new focr_bank();
    b->ctx = c;
    b->T = T;
    b->info.resize(T);
    std::map<std::pair<uint32_t, uint32_t>, uint32_t> cls_of;
    for (uint32_t t = 0; t < T; t++) {
        const uint32_t w = n_w[t], h = n_h[t];
        if (w == 0 || h == 0 || w > MAX_TPL_W || h > MAX_TPL_H) {
            delete b;
            return fail(FOCR_ERR_UNSUPPORTED, "focr_bank_create: template " + std::to_string(t) + " is " +
                                                  std::to_string(w) + "x" + std::to_string(h) + "; supported 1..32 x 1..64");
        }
        auto key = std::make_pair(w, h);
        auto it = cls_of.find(key);
        if (it == cls_of.end()) {
            it = cls_of.emplace(key, (uint32_t)b->classes.size()).first;
            ClassHost ch;
            ch.n_w = w;
            ch.n_h = h;
            ch.np = w <= 16 ? 16 : 32;
            b->classes.push_back(std::move(ch));
        }
        tpl_info(pixels + offsets[t], w, h, b->info[t]);
        b->info[t].cls = it->second;
        b->info[t].data_off = (uint32_t)b->classes[it->second].index.size() * h * b->classes[it->second].np;
        b->classes[it->second].index.push_back(t);
    }
    std::vector<std::vector<uint8_t>> rows_host(b->classes.size());
    std::vector<std::vector<TplInfo>> cls_info(b->classes.size());
    for (size_t ci = 0; ci < b->classes.size(); ci++) {
        ClassHost &ch = b->classes[ci];
        // copy_needle_n_u8 (ncc.rs:925-935): rows zero-padded to np bytes
        std::vector<uint8_t> &rows = rows_host[ci];
        rows.assign((size_t)ch.index.size() * ch.n_h * ch.np, 0);
        for (size_t i = 0; i < ch.index.size(); i++) {
            const uint8_t *src = pixels + offsets[ch.index[i]];
            for (uint32_t y = 0; y < ch.n_h; y++)
                memcpy(&rows[(i * ch.n_h + y) * ch.np], src + (size_t)y * ch.n_w, ch.n_w);
        }
        CU(ch.rows.ensure(rows.size()));
        CU(cudaMemcpy(ch.rows.p, rows.data(), rows.size(), cudaMemcpyHostToDevice));
        CU(ch.index_dev.ensure(ch.index.size() * 4));
        CU(cudaMemcpy(ch.index_dev.p, ch.index.data(), ch.index.size() * 4, cudaMemcpyHostToDevice));
        for (uint32_t t : ch.index) cls_info[ci].push_back(b->info[t]);
    }
    // launch groups of the tcgen05 kernel: pair box sizes of equal height (and padded row width) whose boxes are small
    // enough for the unscaled screen; everything else is a group of its own
    const bool no_merge = getenv("FOCR_TC_NOMERGE") != nullptr;
    auto src_of = [&](size_t ci) {
        const ClassHost &ch = b->classes[ci];
        return TcClassSrc{rows_host[ci].data(), ch.n_w, (uint32_t)ch.index.size(), ch.index.data(), cls_info[ci].data()};
    };
    auto mergeable = [](const ClassHost &ch) { return ch.np == 16 && ch.n_h <= 16 && ch.n_w * ch.n_h <= 256; };
    for (size_t i = 0; i < b->classes.size(); i++) {
        ClassHost &ci = b->classes[i];
        if (ci.group >= 0) continue;
        b->groups.emplace_back();
        const int gi = (int)b->groups.size() - 1;
        GroupHost &g = b->groups.back();
        g.cls[0] = (int)i;
        ci.group = gi;
        ci.group_pos = 0;
        int rc = 0;
        if (!no_merge && mergeable(ci)) {
            for (size_t j = i + 1; j < b->classes.size(); j++) {
                ClassHost &cj = b->classes[j];
                if (cj.group >= 0 || cj.n_h != ci.n_h || cj.np != ci.np || !mergeable(cj)) continue;
                const TcClassSrc src[2] = {src_of(i), src_of(j)};
                rc = tc_class_build(g.tc, src, 2, ci.n_h, ci.np);
                if (rc == 0 && g.tc.supported) {
                    g.cls[1] = (int)j;
                    cj.group = gi;
                    cj.group_pos = (uint32_t)ci.index.size();
                }
                break;
            }
        }
        if (rc == 0 && g.cls[1] < 0) {
            const TcClassSrc src = src_of(i);
            rc = tc_class_build(g.tc, &src, 1, ci.n_h, ci.np);
        }
        if (rc != 0) {
            focr_bank_destroy(b);
            return fail(FOCR_ERR_CUDA, "focr_bank_create: tcgen05 operand upload failed");
        }
    }
    CU(b->info_dev.ensure(T * sizeof(TplInfo)));
    CU(cudaMemcpy(b->info_dev.p, b->info.data(), T * sizeof(TplInfo), cudaMemcpyHostToDevice));
    *out = b;
    return FOCR_OK;
}

extern "C" void focr_bank_destroy(focr_bank *b)
{
    if (!b) return;
    cudaSetDevice(b->ctx->device);
    cudaDeviceSynchronize();
    for (auto &ch : b->classes) {
        ch.rows.release();
        ch.index_dev.release();
    }
    for (auto &g : b->groups) tc_class_release(g.tc);
    b->info_dev.release();
    delete b;
}
extern "C" uint32_t focr_bank_size(const focr_bank *b) { return b ? b->T : 0; }

// ------------------------------------------------------------------------------------------ scan pipeline
struct Geometry {
    uint32_t r_w, r_h, n_out;
    int pitch;       // inverted page pitch (bytes, multiple of 128)
    int spitch;      // stats plane pitch (elements, multiple of 32)
    size_t inv_page_stride, plane_page_stride;
    uint32_t sel_cap;
};

static int make_geometry(uint32_t r_w, uint32_t r_h, uint32_t n_out, Geometry &g)
{
    if (r_w == 0 || r_h == 0) return fail(FOCR_ERR_ARG, "empty page");
    if (r_w > 65535 || r_h > 65535 - PAGE_PAD_ROWS)   // u16 match coordinates (ncc.cpp:8); the staging kernel's grid.y
        return fail(FOCR_ERR_UNSUPPORTED, "page larger than focr_get_limits() allows (height <= " + std::to_string(65535 - PAGE_PAD_ROWS) + ")");
    if (n_out == 0 || n_out > 4096) return fail(FOCR_ERR_ARG, "n_out must be in 1..4096");
    g.r_w = r_w;
    g.r_h = r_h;
    g.n_out = n_out;
    g.pitch = (int)((r_w + 32 + 127) / 128 * 128);  // >= r_w + 32: any 32-byte window row may over-read
    g.spitch = (int)((r_w + 31) / 32 * 32);
    g.inv_page_stride = (size_t)g.pitch * (r_h + PAGE_PAD_ROWS);
    g.plane_page_stride = (size_t)g.spitch * r_h;
    const size_t sc = finalize_sel_cap(r_w, n_out);
    if (sc == 0) return fail(FOCR_ERR_UNSUPPORTED, "page wider than 15360 px is not supported yet");
    g.sel_cap = (uint32_t)sc;
    return FOCR_OK;
}

static bool use_tc(const focr_ctx *c, const focr_bank *b, const ClassHost &ch)
{
    if (c->kernel == FOCR_KERNEL_SIMT) return false;
    return tc_class_supported(b->groups[ch.group].tc);
}

// enqueue the whole device pipeline for nB pages that sit (gray or inverted) in device memory; the chunk reports through
// flag block `flag_block` of the slot (copied to the pinned mirror right away when copy_flags)
static int enqueue_chunk(focr_ctx *c, const focr_bank *b, Slot &s, const Geometry &g, const uint8_t *pages_dev,
                         size_t page_stride, size_t in_pitch, uint32_t nB, float threshold, int invert,
                         focr_match *out_dev, uint32_t *counts_dev, int flag_block = 0, bool copy_flags = true)
{
    cudaStream_t st = c->stream;
    const uint32_t T = b->T;
    bool any_simt = false, any_tc = false, any_pair = false;
    for (auto &ch : b->classes) {
        if (c->kernel == FOCR_KERNEL_TCGEN05 && !use_tc(c, b, ch))
            return fail(FOCR_ERR_UNSUPPORTED, "tcgen05 kernel does not support box " + std::to_string(ch.n_w) + "x" +
                                                  std::to_string(ch.n_h));
        any_simt |= !use_tc(c, b, ch);
        any_tc |= use_tc(c, b, ch);
    }
    for (auto &gr : b->groups) any_pair |= gr.cls[1] >= 0;
    if (s.hits_per_page < c->hits_per_page) s.hits_per_page = c->hits_per_page;
    const size_t hit_cap = (size_t)s.hits_per_page * nB;
    const size_t PT = (size_t)nB * T;
    const size_t nR = std::max(nB, s.reserve_pages), PTR = nR * T;   // allocation sizes
    CU(s.inv.ensure(g.inv_page_stride * nR));
    // statistics planes: the tcgen05 path reads ONE packed word per window and box size (sp / sp2) for boxes of at most 256
    // pixels, sp + pf otherwise; only the SIMT kernel reads the sum-of-squares and f64 planes
    bool any_unpacked = any_simt;
    for (auto &gr : b->groups) any_unpacked |= use_tc(c, b, b->classes[gr.cls[0]]) && gr.tc.sshift != 0;
    CU(s.sp.ensure(g.plane_page_stride * nR * 4));
    if (any_unpacked) CU(s.pf.ensure(g.plane_page_stride * nR * 4));
    if (any_simt) {
        CU(s.s2p.ensure(g.plane_page_stride * nR * 4));
        CU(s.rn.ensure(g.plane_page_stride * nR * 8));
    }
    if (any_tc && any_pair) CU(s.sp2.ensure(g.plane_page_stride * nR * 4));   // paired groups are always packed (<= 256 pixels)
    // per-row hit counts and, right behind them, the selection counters: ONE memset clears both
    const size_t rowcount_bytes = (PT * g.r_h * 4 + 7) & ~(size_t)7;
    CU(s.rowcount.ensure(((PTR * g.r_h * 4 + 7) & ~(size_t)7) + PTR * 8));
    CU(s.hits.ensure((size_t)s.hits_per_page * nR * sizeof(Hit)));
    const size_t n_lists = (size_t)c->sm_count * TC_LISTS_PER_CTA;
    if (any_tc) {
        CU(s.cands.ensure(n_lists * c->cand_per_warp * sizeof(Hit)));
        CU(s.candcnt.ensure(n_lists * 4));
    }
    CU(s.sel.ensure(PTR * g.sel_cap * 8));
    CU(s.ycut.ensure(PTR * 4));
    CU(s.flags.ensure(FLAG_BLOCKS * FLAG_WORDS * 4));
    unsigned int *const flags = s.flags.as<unsigned int>() + (size_t)flag_block * FLAG_WORDS;

    {
        StageTimer tm(c, FOCR_STAGE_INVERT);
        CU(launch_stage_invert(pages_dev, page_stride, in_pitch, s.inv.as<uint8_t>(), g.inv_page_stride, g.pitch,
                               g.r_w, g.r_h, nB, invert, st));
    }
    c->launches++;
    CU(cudaMemsetAsync(s.rowcount.p, 0, rowcount_bytes + PT * 8, st));
    CU(cudaMemsetAsync(flags, 0, FLAG_WORDS * 4, st));

    HitSink sink;
    sink.hits = s.hits.as<Hit>();
    sink.hit_cap = (uint32_t)std::min<size_t>(hit_cap, 0xFFFFFFFFu);
    sink.hit_count = flags;
    sink.rowcount = s.rowcount.as<unsigned int>();
    sink.T = T;
    sink.r_h = g.r_h;

    // One scan per launch group on the tcgen05 path (its one or two box sizes share a statistics pass: the vertical sums
    // and the row prefix sums do not depend on the width), one per box size on the SIMT path.
    auto scan_unit = [&](const ClassHost &ch, const ClassHost *ch2, const TcClass *tcg) -> int {
        const bool tc = tcg != nullptr;
        StatsArgs sa{};
        sa.inv = s.inv.as<uint8_t>();
        sa.inv_page_stride = g.inv_page_stride;
        sa.pitch = g.pitch;
        sa.r_w = g.r_w;
        sa.r_h = g.r_h;
        sa.n_w = ch.n_w;
        sa.n_h = ch.n_h;
        sa.inv_n_f = 1.0f / (float)(ch.n_w * ch.n_h);
        sa.sp = s.sp.as<uint32_t>();
        sa.s2p = tc ? nullptr : s.s2p.as<uint32_t>();   // the tcgen05 path recomputes s2_p for its few survivors
        sa.pf = s.pf.as<float>();
        sa.rn = tc ? nullptr : s.rn.as<double>();
        sa.spitch = g.spitch;
        sa.plane_page_stride = g.plane_page_stride;
        sa.pack = (tc && tcg->sshift == 0) ? 1 : 0;
        if (ch2) {
            sa.n_w2 = ch2->n_w;
            sa.inv_n_f2 = 1.0f / (float)(ch2->n_w * ch2->n_h);
            sa.sp2 = s.sp2.as<uint32_t>();
            sa.pf2 = nullptr;
        }
        {
            StageTimer tm(c, FOCR_STAGE_STATS);
            CU(launch_window_stats(sa, nB, st));
            c->launches++;
        }
        ScanArgs a;
        a.inv = sa.inv;
        a.inv_page_stride = g.inv_page_stride;
        a.pitch = g.pitch;
        a.r_w = g.r_w;
        a.r_h = g.r_h;
        a.cls.n_w = ch.n_w;
        a.cls.n_h = ch.n_h;
        a.cls.np = ch.np;
        a.cls.n_tpl = (uint32_t)ch.index.size();
        a.cls.tpl_index = ch.index_dev.as<uint32_t>();
        a.cls.rows = ch.rows.as<uint8_t>();
        a.tpl = b->info_dev.as<TplInfo>();
        a.sp = sa.sp;
        a.s2p = sa.s2p;
        a.pf = sa.pf;
        a.rn = sa.rn;
        a.sp2 = sa.sp2;
        a.pf2 = sa.pf2;
        a.pack = sa.pack;
        a.spitch = g.spitch;
        a.plane_page_stride = g.plane_page_stride;
        a.thr_d = (double)threshold;  // ncc.cpp:83
        a.thr_f = threshold;
        a.sink = sink;
        a.cands = s.cands.as<Hit>();
        a.cand_cap = c->cand_per_warp;
        a.cand_count = s.candcnt.as<unsigned int>();
        a.cand_max = flags + 3;
        a.acc_out = nullptr;
        // (no clearing of cand_count: every epilogue warp of the launch writes its list's count when it is through, and the
        // exact pass reads exactly the launch's lists)
        int nl = 0;
        {
            StageTimer tm(c, FOCR_STAGE_SCAN);
            if (tc) {
                ExactStage hook(c);
                CU(launch_scan_tc(*tcg, a, nB, c->sm_count, st, &nl, nullptr, -1, &hook));
            }
            else
                CU(launch_scan_simt(a, nB, st, &nl));
            tm.launches = nl;
        }
        c->launches += nl;
        return FOCR_OK;
    };
    auto fits = [&](const ClassHost &ch) { return ch.n_w <= g.r_w && ch.n_h <= g.r_h; };  // else no window: no hits
    for (auto &gr : b->groups) {
        const ClassHost &c0 = b->classes[gr.cls[0]];
        const ClassHost *c1 = gr.cls[1] >= 0 ? &b->classes[gr.cls[1]] : nullptr;
        if (use_tc(c, b, c0)) {
            // (a page narrower than the wider box: the group still runs, that box's windows are all flagged invalid)
            if (!fits(c0) && !(c1 && fits(*c1))) continue;
            if (int rc = scan_unit(c0, c1, &gr.tc)) return rc;
        } else {
            if (fits(c0))
                if (int rc = scan_unit(c0, nullptr, nullptr)) return rc;
            if (c1 && fits(*c1))
                if (int rc = scan_unit(*c1, nullptr, nullptr)) return rc;
        }
    }

    FinalizeArgs f;
    f.hits = s.hits.as<Hit>();
    f.hit_count = flags;
    f.hit_cap = sink.hit_cap;
    f.rowcount = s.rowcount.as<unsigned int>();
    f.y_cut = s.ycut.as<uint32_t>();
    f.sel_count = (unsigned int *)(s.rowcount.as<uint8_t>() + rowcount_bytes);
    f.sel = s.sel.as<unsigned long long>();
    f.sel_cap = g.sel_cap;
    f.overflow = flags + 1;
    f.T = T;
    f.r_h = g.r_h;
    f.n_pages = nB;
    f.n_out = g.n_out;
    f.out = out_dev;
    f.counts = counts_dev;
    int nl = 0;
    {
        StageTimer tm(c, FOCR_STAGE_FINALIZE, 3);
        CU(launch_finalize(f, st, &nl));
    }
    c->launches += nl;
    if (copy_flags)
        CU(cudaMemcpyAsync(s.flags_host + (size_t)flag_block * FLAG_WORDS, flags, FLAG_WORDS * 4, cudaMemcpyDeviceToHost, st));
    return FOCR_OK;
}

// after the stream has been synchronised: did a wait inside the tcgen05 kernel give up (protocol bug)?
static int chunk_watchdog(const unsigned int *fh)
{
    if (!fh[4]) return FOCR_OK;
    return fail(FOCR_ERR_CUDA, "internal: scan_tc pipeline stalled (wait tag " + std::to_string(fh[5]) + ", info " +
                                   std::to_string(fh[6]) + ", CTA " + std::to_string(fh[7]) + ", warp " +
                                   std::to_string(fh[8]) + ", parity " + std::to_string(fh[9]) + ")");
}

// after the stream has been synchronised: did the chunk's candidate lists or its hit list overflow?  (grows the context's
// capacities; the caller then repeats the scan)
static bool chunk_overflowed(focr_ctx *c, Slot &s, const unsigned int *fh, uint32_t nB)
{
    if (getenv("FOCR_DEBUG_COUNTS"))
        fprintf(stderr, "[focr] chunk of %u pages: hits %u, longest overflowing candidate list %u (cap %u)\n", nB,
                fh[0], fh[3], c->cand_per_warp);
    if (fh[3] > c->cand_per_warp) {  // a warp's private candidate list overflowed
        c->cand_per_warp = (uint32_t)std::min<size_t>((size_t)fh[3] + fh[3] / 2, 1u << 24);
        return true;
    }
    const size_t cap = (size_t)s.hits_per_page * nB;
    const size_t seen = fh[0];
    if (seen > cap) {
        const size_t need = (seen + nB - 1) / nB;
        c->hits_per_page = (uint32_t)std::min<size_t>(need + need / 4 + 1024, 0x7FFFFFFFu);
        return true;
    }
    return false;
}

// max_pages: 16 for host buffers (the pipeline's fill and drain grow with the chunk), 25 for pages already in HBM (fewer
// launches and tail rounds: +1.2 % on config 3)
static uint32_t pick_chunk(const focr_bank *b, const Geometry &g, uint32_t n_pages, uint32_t hits_per_page, size_t max_pages)
{
    const size_t per_page = g.inv_page_stride + (size_t)g.r_w * g.r_h + g.plane_page_stride * 20 +
                            (size_t)b->T * g.r_h * 4 + (size_t)hits_per_page * sizeof(Hit) +
                            (size_t)b->T * g.sel_cap * 8 + (size_t)b->T * g.n_out * 8;
    size_t nb = (size_t)(6ull << 30) / std::max<size_t>(per_page, 1);
    static const size_t cap = [] { const char *e = getenv("FOCR_CHUNK_PAGES"); return e ? (size_t)std::max(1, atoi(e)) : (size_t)0; }();   // experiments
    nb = std::max<size_t>(1, std::min<size_t>(nb, cap ? cap : max_pages));
    return (uint32_t)std::min<size_t>(nb, n_pages);
}

extern "C" int focr_ncc_scan_device(focr_ctx *c, const focr_bank *b, const uint8_t *pages_dev, size_t page_stride,
                                    size_t pitch, uint32_t r_w, uint32_t r_h, uint32_t n_pages, float threshold,
                                    uint32_t n_out, focr_match *out_dev, uint32_t *counts_dev)
{
    if (!c || !b || !pages_dev || !out_dev || !counts_dev || n_pages == 0)
        return fail(FOCR_ERR_ARG, "focr_ncc_scan_device: NULL argument or no pages");
    if (b->ctx != c) return fail(FOCR_ERR_ARG, "bank belongs to another context");
    CU(cudaSetDevice(c->device));
    Geometry g;
    int rc = make_geometry(r_w, r_h, n_out, g);
    if (rc) return rc;
    // All chunks are enqueued back to back (they share one slot's scratch, the stream orders them); each reports through
    // its own flag block and the host looks at the blocks ONCE, after the last chunk (every FLAG_BLOCKS chunks for very
    // large batches).  Only an overflowing candidate / hit list makes the host grow the lists and repeat the scan.
    for (int attempt = 0; attempt < 6; attempt++) {
        const uint32_t B = pick_chunk(b, g, n_pages, c->hits_per_page, 25);
        Slot &s = c->slot[0];
        s.reserve_pages = B;
        bool redo = false;
        for (uint32_t r0 = 0; r0 < n_pages && !redo; r0 += B * FLAG_BLOCKS) {
            const uint32_t r1 = (uint32_t)std::min<uint64_t>((uint64_t)r0 + (uint64_t)B * FLAG_BLOCKS, n_pages);
            int k = 0;
            for (uint32_t p0 = r0; p0 < r1; p0 += B, k++) {
                const uint32_t nB = std::min(B, r1 - p0);
                rc = enqueue_chunk(c, b, s, g, pages_dev + (size_t)p0 * page_stride, page_stride, pitch, nB, threshold, 1,
                                   out_dev + (size_t)p0 * b->T * n_out, counts_dev + (size_t)p0 * b->T, k, false);
                if (rc) return rc;
            }
            CU(cudaMemcpyAsync(s.flags_host, s.flags.p, (size_t)k * FLAG_WORDS * 4, cudaMemcpyDeviceToHost, c->stream));
            CU(cudaStreamSynchronize(c->stream));
            k = 0;
            for (uint32_t p0 = r0; p0 < r1 && !redo; p0 += B, k++) {
                const unsigned int *fh = s.flags_host + (size_t)k * FLAG_WORDS;
                if (fh[1]) return fail(FOCR_ERR_CUDA, "internal: selection list overflow");
                if (int wrc = chunk_watchdog(fh)) return wrc;
                redo = chunk_overflowed(c, s, fh, std::min(B, r1 - p0));
            }
        }
        if (!redo) return FOCR_OK;
    }
    return fail(FOCR_ERR_NOMEM, "hit list kept overflowing");
}

// ---- pageable callers: staged through library-owned pinned buffers with a few host threads
bool focr_internal_host_pinned(const void *p);
void focr_internal_parallel_copy(uint8_t *dst, size_t dst_stride, const uint8_t *src, size_t src_stride, size_t row_bytes, size_t rows);
static bool host_ptr_is_pinned(const void *p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
}

// Host threads per staging copy: the hardware threads, at most 8, shared between the GPUs that stage at the same time --
// the ranks torchrun placed on this host (LOCAL_WORLD_SIZE) or the devices of a focr_multi call in flight.
static std::atomic<int> g_stage_sharers{1};
void focr_internal_stage_sharers(int n) { g_stage_sharers.store(n < 1 ? 1 : n); }
static int stage_threads()
{
    static const int base = [] {
        if (const char *e = getenv("FOCR_STAGE_THREADS")) return -std::max(1, std::min(16, atoi(e)));   // explicit: not divided
        const unsigned hc = std::max(2u, std::thread::hardware_concurrency());
        int ranks = 1;
        if (const char *e = getenv("LOCAL_WORLD_SIZE")) ranks = std::max(1, atoi(e));
        return (int)std::max(4u, std::min(8u, hc / (unsigned)ranks));   // (4 oversubscribed threads still beat 2: the copies wait on memory)
    }();
    if (base < 0) return -base;
    return std::max(std::min(base, 4), base / g_stage_sharers.load());
}

// A small persistent pool for the staging copies: a chunk is staged in 32-MB pieces, and creating and joining a set of
// threads per piece costs about as much as copying a quarter of it.  run(n, fn) executes fn(0..n-1), fn(0) on the caller.
// One job at a time per pool; every calling host thread has its own.
class StagePool {
    std::mutex job_mu, mu;
    std::condition_variable cv_work, cv_done;
    std::vector<std::thread> workers;
    std::function<void(int)> fn;
    int n_parts = 0, next = 0, pending = 0;
    uint64_t epoch = 0;
    bool stop = false;
    void loop()
    {
        uint64_t seen = 0;
        std::unique_lock<std::mutex> lk(mu);
        for (;;) {
            cv_work.wait(lk, [&] { return stop || (epoch != seen && next < n_parts); });
            if (stop) return;
            while (next < n_parts) {
                const int i = next++;
                lk.unlock();
                fn(i);
                lk.lock();
                if (--pending == 0) cv_done.notify_all();
            }
            seen = epoch;
        }
    }
  public:
    ~StagePool()
    {
        {
            std::lock_guard<std::mutex> lk(mu);
            stop = true;
        }
        cv_work.notify_all();
        for (auto &t : workers) t.join();
    }
    void run(int n, const std::function<void(int)> &f)
    {
        if (n <= 1) {
            if (n == 1) f(0);
            return;
        }
        std::lock_guard<std::mutex> job(job_mu);
        std::unique_lock<std::mutex> lk(mu);
        while ((int)workers.size() < n - 1) workers.emplace_back([this] { loop(); });
        fn = f;
        n_parts = n, next = 1, pending = n - 1;
        epoch++;
        lk.unlock();
        cv_work.notify_all();
        f(0);
        lk.lock();
        while (next < n_parts) {   // workers that are slow to wake (an oversubscribed host) do not hold the job up
            const int i = next++;
            lk.unlock();
            f(i);
            lk.lock();
            --pending;
        }
        cv_done.wait(lk, [&] { return pending == 0; });
        n_parts = 0;
    }
};
static StagePool &stage_pool()
{
    // one pool per calling host thread: the devices of a focr_multi call stage from their own threads at the same time
    thread_local StagePool p;
    return p;
}

// rows x row_bytes from src (stride src_stride) to dst (stride dst_stride), split over the staging threads
static void parallel_copy(uint8_t *dst, size_t dst_stride, const uint8_t *src, size_t src_stride, size_t row_bytes, size_t rows)
{
    if (rows == 1 || (dst_stride == row_bytes && src_stride == row_bytes)) {   // one contiguous block: split by bytes
        const size_t total = row_bytes * rows;
        const int nt = (int)std::max<size_t>(1, std::min<size_t>(stage_threads(), total >> 20));
        if (nt == 1) {
            memcpy(dst, src, total);
            return;
        }
        const size_t per = ((total + nt - 1) / nt + 4095) & ~(size_t)4095;
        stage_pool().run(nt, [=](int i) {
            const size_t o = std::min(total, per * (size_t)i), n = std::min(total, per * (size_t)(i + 1)) - o;
            if (n) memcpy(dst + o, src + o, n);
        });
        return;
    }
    const int nt = (int)std::max<size_t>(1, std::min<size_t>(stage_threads(), rows));
    stage_pool().run(nt, [=](int i) {
        for (size_t r = (size_t)i; r < rows; r += (size_t)nt) memcpy(dst + r * dst_stride, src + r * src_stride, row_bytes);
    });
}

bool focr_internal_host_pinned(const void *p) { return host_ptr_is_pinned(p); }
// focr_decode.cu: its row gather runs on the same pool, with the same thread budget
void focr_internal_parallel_for(int n, const std::function<void(int)> &fn) { stage_pool().run(n, fn); }
int focr_internal_stage_threads() { return stage_threads(); }
void focr_internal_parallel_copy(uint8_t *dst, size_t dst_stride, const uint8_t *src, size_t src_stride, size_t row_bytes, size_t rows)
{
    parallel_copy(dst, dst_stride, src, src_stride, row_bytes, rows);
}

static int scan_host_impl(focr_ctx *c, const focr_bank *b, const uint8_t *pages_host, size_t page_stride,
                          uint32_t r_w, uint32_t r_h, uint32_t n_pages, float threshold, uint32_t n_out, int invert,
                          focr_match *out_host, uint32_t *counts_host)
{
    CU(cudaSetDevice(c->device));
    Geometry g;
    int rc = make_geometry(r_w, r_h, n_out, g);
    if (rc) return rc;
    const size_t page_bytes = (size_t)r_w * r_h;
    const uint32_t T = b->T;
    // Caller-owned buffers may be pageable (the reference hands over a Vec<u8>, ncc.rs:575): cudaMemcpyAsync would then
    // stage through the driver's small bounce buffer synchronously.  Such buffers go through the slot's own pinned
    // staging instead, filled / drained by a few host threads while the previous chunk computes.
    const bool stage_in = !host_ptr_is_pinned(pages_host) || getenv("FOCR_FORCE_STAGING");
    const bool stage_out = !host_ptr_is_pinned(out_host) || !host_ptr_is_pinned(counts_host) || getenv("FOCR_FORCE_STAGING");
    for (int attempt = 0; attempt < 6; attempt++) {
        const uint32_t B = pick_chunk(b, g, n_pages, c->hits_per_page, 16);
        // chunk schedule {first page, pages}: the first chunks ramp up 2, 4, 8, .. B so that the kernels start after the
        // H2D of two pages instead of a whole chunk (the copy of chunk i+1, twice the size, still hides behind chunk i)
        std::vector<std::pair<uint32_t, uint32_t>> chunks;
        for (uint32_t p0 = 0, sz = std::min<uint32_t>(2, B); p0 < n_pages; sz = std::min(B, sz * 2)) {
            const uint32_t nB = std::min(sz, n_pages - p0);
            chunks.emplace_back(p0, nB);
            p0 += nB;
        }
        bool redo = false;
        const uint32_t Bres = std::min(B, n_pages);
        for (auto &sl : c->slot) sl.reserve_pages = Bres;
        // a chunk whose D2H has completed: check its flags, hand staged results to the caller
        auto drain = [&](uint32_t k) -> int {
            Slot &s = c->slot[k & 1];
            const uint32_t p0 = chunks[k].first, nB = chunks[k].second;
            if (s.flags_host[1]) return fail(FOCR_ERR_CUDA, "internal: selection list overflow");
            if (int wrc = chunk_watchdog(s.flags_host)) return wrc;
            if (chunk_overflowed(c, s, s.flags_host, nB)) {
                redo = true;
                return FOCR_OK;
            }
            if (stage_out) {
                parallel_copy((uint8_t *)(out_host + (size_t)p0 * T * n_out), 0, s.out_pin.as<uint8_t>(), 0,
                              (size_t)nB * T * n_out * sizeof(focr_match), 1);
                memcpy(counts_host + (size_t)p0 * T, s.counts_pin.p, (size_t)nB * T * 4);
            }
            return FOCR_OK;
        };
        uint32_t ci = 0;
        // software pipeline over chunks: H2D of chunk i+1 overlaps the kernels of chunk i,
        // D2H of chunk i overlaps the kernels of chunk i+1
        for (; ci < chunks.size(); ci++) {
            const uint32_t p0 = chunks[ci].first, nB = chunks[ci].second;
            Slot &s = c->slot[ci & 1];
            if (ci >= 2) {  // the slot's previous chunk must be fully drained before its buffers are reused
                CU(cudaEventSynchronize(s.ev_d2h));
                if (int drc = drain(ci - 2)) return drc;
                if (redo) break;
            }
            CU(s.gray.ensure(page_bytes * Bres));
            CU(s.out.ensure((size_t)Bres * T * n_out * sizeof(focr_match)));
            CU(s.counts.ensure((size_t)Bres * T * 4));
            const uint8_t *src = pages_host + (size_t)p0 * page_stride;
            if (stage_in) {
                // pageable pages: host threads copy a few pages into the slot's pinned staging, their DMA starts at once
                // and runs while the next few are being copied (copy-in and H2D of a chunk overlap, not add up)
                CU(s.gray_pin.ensure(page_bytes * Bres));
                const uint32_t sub = std::max<uint32_t>(1, (uint32_t)((32u << 20) / std::max<size_t>(page_bytes, 1)));
                for (uint32_t q0 = 0; q0 < nB; q0 += sub) {
                    const uint32_t nq = std::min(sub, nB - q0);
                    uint8_t *pin = s.gray_pin.as<uint8_t>() + (size_t)q0 * page_bytes;
                    parallel_copy(pin, page_bytes, src + (size_t)q0 * page_stride, page_stride, page_bytes, nq);
                    CU(cudaMemcpyAsync(s.gray.as<uint8_t>() + (size_t)q0 * page_bytes, pin, page_bytes * nq, cudaMemcpyHostToDevice, c->h2d));
                }
            } else if (page_stride == page_bytes) {
                CU(cudaMemcpyAsync(s.gray.p, src, page_bytes * nB, cudaMemcpyHostToDevice, c->h2d));
            } else {
                CU(cudaMemcpy2DAsync(s.gray.p, page_bytes, src, page_stride, page_bytes, nB, cudaMemcpyHostToDevice, c->h2d));
            }
            CU(cudaEventRecord(s.ev_h2d, c->h2d));
            CU(cudaStreamWaitEvent(c->stream, s.ev_h2d, 0));
            rc = enqueue_chunk(c, b, s, g, s.gray.as<uint8_t>(), page_bytes, r_w, nB, threshold, invert,
                               s.out.as<focr_match>(), s.counts.as<uint32_t>());
            if (rc) return rc;
            CU(cudaEventRecord(s.ev_compute, c->stream));
            CU(cudaStreamWaitEvent(c->d2h, s.ev_compute, 0));
            focr_match *dst_out = out_host + (size_t)p0 * T * n_out;
            uint32_t *dst_counts = counts_host + (size_t)p0 * T;
            if (stage_out) {
                CU(s.out_pin.ensure((size_t)Bres * T * n_out * sizeof(focr_match)));
                CU(s.counts_pin.ensure((size_t)Bres * T * 4));
                dst_out = s.out_pin.as<focr_match>();
                dst_counts = s.counts_pin.as<uint32_t>();
            }
            CU(cudaMemcpyAsync(dst_out, s.out.p, (size_t)nB * T * n_out * sizeof(focr_match), cudaMemcpyDeviceToHost, c->d2h));
            CU(cudaMemcpyAsync(dst_counts, s.counts.p, (size_t)nB * T * 4, cudaMemcpyDeviceToHost, c->d2h));
            CU(cudaEventRecord(s.ev_d2h, c->d2h));
        }
        CU(cudaStreamSynchronize(c->h2d));
        CU(cudaStreamSynchronize(c->stream));
        CU(cudaStreamSynchronize(c->d2h));
        if (!redo) {
            // the last (up to) two chunks have not been checked yet
            const uint32_t n_chunks = ci;
            for (uint32_t k = (n_chunks >= 2 ? n_chunks - 2 : 0); k < n_chunks && !redo; k++)
                if (int drc = drain(k)) return drc;
        }
        if (!redo) return FOCR_OK;
    }
    return fail(FOCR_ERR_NOMEM, "hit list kept overflowing");
}

// ---- page-locking for callers that have no CUDA binding of their own (include/focr_b200.h)
extern "C" int focr_pin_register(focr_ctx *c, void *ptr, size_t bytes)
{
    if (!c || !ptr || bytes == 0) return fail(FOCR_ERR_ARG, "focr_pin_register: NULL argument or no bytes");
    CU(cudaSetDevice(c->device));
    CU(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
    return FOCR_OK;
}
extern "C" int focr_pin_unregister(focr_ctx *c, void *ptr)
{
    if (!c || !ptr) return fail(FOCR_ERR_ARG, "focr_pin_unregister: NULL argument");
    CU(cudaSetDevice(c->device));
    CU(cudaHostUnregister(ptr));
    return FOCR_OK;
}
extern "C" int focr_pin_alloc(focr_ctx *c, size_t bytes, void **out)
{
    if (!c || !out || bytes == 0) return fail(FOCR_ERR_ARG, "focr_pin_alloc: NULL argument or no bytes");
    CU(cudaSetDevice(c->device));
    CU(cudaHostAlloc(out, bytes, cudaHostAllocPortable));
    return FOCR_OK;
}
extern "C" int focr_pin_free(focr_ctx *c, void *ptr)
{
    if (!c || !ptr) return fail(FOCR_ERR_ARG, "focr_pin_free: NULL argument");
    CU(cudaSetDevice(c->device));
    CU(cudaFreeHost(ptr));
    return FOCR_OK;
}

extern "C" int focr_ncc_scan(focr_ctx *c, const focr_bank *b, const uint8_t *pages_host, size_t page_stride,
                             uint32_t r_w, uint32_t r_h, uint32_t n_pages, float threshold, uint32_t n_out,
                             focr_match *out_host, uint32_t *counts_host)
{
    if (!c || !b || !pages_host || !out_host || !counts_host || n_pages == 0)
        return fail(FOCR_ERR_ARG, "focr_ncc_scan: NULL argument or no pages");
    if (b->ctx != c) return fail(FOCR_ERR_ARG, "bank belongs to another context");
    if (page_stride < (size_t)r_w * r_h) return fail(FOCR_ERR_ARG, "page_stride smaller than a page");
    return scan_host_impl(c, b, pages_host, page_stride, r_w, r_h, n_pages, threshold, n_out, 1, out_host, counts_host);
}

// ------------------------------------------------------------------------------------------ parity probes
extern "C" int focr_window_stats(focr_ctx *c, const uint8_t *page_gray_host, uint32_t r_w, uint32_t r_h, uint32_t n_w,
                                 uint32_t n_h, uint32_t *s_p_host, uint64_t *s2_p_host, double *patch_rnorm_host)
{
    if (!c || !page_gray_host || !s_p_host || !s2_p_host || !patch_rnorm_host)
        return fail(FOCR_ERR_ARG, "focr_window_stats: NULL argument");
    if (n_w == 0 || n_h == 0 || n_w > MAX_TPL_W || n_h > MAX_TPL_H || n_w > r_w || n_h > r_h)
        return fail(FOCR_ERR_UNSUPPORTED, "focr_window_stats: unsupported box");
    CU(cudaSetDevice(c->device));
    Geometry g;
    int rc = make_geometry(r_w, r_h, 1024, g);
    if (rc) return rc;
    Slot &s = c->slot[0];
    const size_t page_bytes = (size_t)r_w * r_h;
    CU(s.gray.ensure(page_bytes));
    CU(s.inv.ensure(g.inv_page_stride));
    CU(s.sp.ensure(g.plane_page_stride * 4));
    CU(s.s2p.ensure(g.plane_page_stride * 4));
    CU(s.pf.ensure(g.plane_page_stride * 4));
    CU(s.rn.ensure(g.plane_page_stride * 8));
    CU(cudaMemcpyAsync(s.gray.p, page_gray_host, page_bytes, cudaMemcpyHostToDevice, c->stream));
    CU(launch_stage_invert(s.gray.as<uint8_t>(), page_bytes, r_w, s.inv.as<uint8_t>(), g.inv_page_stride, g.pitch, r_w,
                           r_h, 1, 1, c->stream));
    StatsArgs sa{};
    sa.inv = s.inv.as<uint8_t>();
    sa.inv_page_stride = g.inv_page_stride;
    sa.pitch = g.pitch;
    sa.r_w = r_w;
    sa.r_h = r_h;
    sa.n_w = n_w;
    sa.n_h = n_h;
    sa.inv_n_f = 1.0f / (float)(n_w * n_h);
    sa.sp = s.sp.as<uint32_t>();
    sa.s2p = s.s2p.as<uint32_t>();
    sa.pf = s.pf.as<float>();
    sa.rn = s.rn.as<double>();
    sa.spitch = g.spitch;
    sa.plane_page_stride = g.plane_page_stride;
    CU(launch_window_stats(sa, 1, c->stream));
    c->launches += 2;
    const uint32_t xs = r_w - n_w + 1, ys = r_h - n_h + 1;
    std::vector<uint32_t> s2((size_t)xs * ys);
    CU(cudaMemcpy2DAsync(s_p_host, (size_t)r_w * 4, s.sp.p, (size_t)g.spitch * 4, (size_t)xs * 4, ys,
                         cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpy2DAsync(s2.data(), (size_t)xs * 4, s.s2p.p, (size_t)g.spitch * 4, (size_t)xs * 4, ys,
                         cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpy2DAsync(patch_rnorm_host, (size_t)r_w * 8, s.rn.p, (size_t)g.spitch * 8, (size_t)xs * 8, ys,
                         cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    for (uint32_t y = 0; y < ys; y++)
        for (uint32_t x = 0; x < xs; x++) s2_p_host[(size_t)y * r_w + x] = s2[(size_t)y * xs + x];
    return FOCR_OK;
}

extern "C" int focr_ncc_numerators(focr_ctx *c, const focr_bank *b, uint32_t t, const uint8_t *page_gray_host,
                                   uint32_t r_w, uint32_t r_h, uint32_t *acc_host)
{
    if (!c || !b || !page_gray_host || !acc_host || t >= b->T)
        return fail(FOCR_ERR_ARG, "focr_ncc_numerators: bad argument");
    CU(cudaSetDevice(c->device));
    Geometry g;
    int rc = make_geometry(r_w, r_h, 1024, g);
    if (rc) return rc;
    const TplInfo &ti = b->info[t];
    const ClassHost &ch = b->classes[ti.cls];
    if (ch.n_w > r_w || ch.n_h > r_h) return fail(FOCR_ERR_ARG, "template larger than the page");
    Slot &s = c->slot[0];
    const size_t page_bytes = (size_t)r_w * r_h;
    CU(s.gray.ensure(page_bytes));
    CU(s.inv.ensure(g.inv_page_stride));
    CU(s.sp.ensure(g.plane_page_stride * 4));
    CU(s.s2p.ensure(g.plane_page_stride * 4));
    CU(s.pf.ensure(g.plane_page_stride * 4));
    CU(s.rn.ensure(g.plane_page_stride * 8));
    CU(s.rowcount.ensure((size_t)b->T * r_h * 4));
    CU(s.hits.ensure(1024 * sizeof(Hit)));
    CU(s.flags.ensure(64));
    CU(s.acc.ensure(page_bytes * 4));
    cudaStream_t st = c->stream;
    CU(cudaMemcpyAsync(s.gray.p, page_gray_host, page_bytes, cudaMemcpyHostToDevice, st));
    CU(launch_stage_invert(s.gray.as<uint8_t>(), page_bytes, r_w, s.inv.as<uint8_t>(), g.inv_page_stride, g.pitch, r_w,
                           r_h, 1, 1, st));
    CU(cudaMemsetAsync(s.rowcount.p, 0, (size_t)b->T * r_h * 4, st));
    CU(cudaMemsetAsync(s.flags.p, 0, 64, st));
    CU(cudaMemsetAsync(s.acc.p, 0, page_bytes * 4, st));
    StatsArgs sa{};
    sa.inv = s.inv.as<uint8_t>();
    sa.inv_page_stride = g.inv_page_stride;
    sa.pitch = g.pitch;
    sa.r_w = r_w;
    sa.r_h = r_h;
    sa.n_w = ch.n_w;
    sa.n_h = ch.n_h;
    sa.inv_n_f = 1.0f / (float)(ch.n_w * ch.n_h);
    sa.sp = s.sp.as<uint32_t>();
    sa.s2p = s.s2p.as<uint32_t>();
    sa.pf = s.pf.as<float>();
    sa.rn = s.rn.as<double>();
    sa.spitch = g.spitch;
    sa.plane_page_stride = g.plane_page_stride;
    CU(launch_window_stats(sa, 1, st));
    ScanArgs a;
    a.inv = sa.inv;
    a.inv_page_stride = g.inv_page_stride;
    a.pitch = g.pitch;
    a.r_w = r_w;
    a.r_h = r_h;
    a.cls.n_w = ch.n_w;
    a.cls.n_h = ch.n_h;
    a.cls.np = ch.np;
    a.cls.n_tpl = 1;
    const uint32_t pos = ti.data_off / (ch.n_h * ch.np);
    a.cls.tpl_index = ch.index_dev.as<uint32_t>() + pos;
    a.cls.rows = ch.rows.as<uint8_t>() + ti.data_off;
    a.tpl = b->info_dev.as<TplInfo>();
    a.sp = sa.sp;
    a.s2p = sa.s2p;
    a.pf = sa.pf;
    a.rn = sa.rn;
    a.spitch = g.spitch;
    a.plane_page_stride = g.plane_page_stride;
    a.thr_d = 2.0;  // nothing is a hit: this probe only wants the numerators
    a.thr_f = 2.0f;
    a.sink.hits = s.hits.as<Hit>();
    a.sink.hit_cap = 1024;
    a.sink.hit_count = s.flags.as<unsigned int>();
    a.sink.rowcount = s.rowcount.as<unsigned int>();
    a.sink.T = b->T;
    a.sink.r_h = r_h;
    CU(s.cands.ensure((size_t)c->sm_count * TC_LISTS_PER_CTA * 16 * sizeof(Hit)));
    CU(s.candcnt.ensure((size_t)c->sm_count * TC_LISTS_PER_CTA * 4));
    a.cands = s.cands.as<Hit>();
    a.cand_cap = 16;
    a.cand_count = s.candcnt.as<unsigned int>();
    a.cand_max = s.flags.as<unsigned int>() + 3;
    a.acc_out = s.acc.as<uint32_t>();
    int nl = 0;
    if (c->kernel == FOCR_KERNEL_TCGEN05) {
        // same probe through the tcgen05 kernel: the template's whole launch group runs, its column is dumped from TMEM
        const GroupHost &gr = b->groups[ch.group];
        if (!tc_class_supported(gr.tc)) return fail(FOCR_ERR_UNSUPPORTED, "tcgen05 kernel does not support this box");
        if (gr.cls[1] >= 0) {   // the A2 warps read the planes of both box sizes of the group
            const ClassHost &c0 = b->classes[gr.cls[0]], &c1 = b->classes[gr.cls[1]];
            if (c0.n_w > r_w || c1.n_w > r_w) return fail(FOCR_ERR_ARG, "template larger than the page");
            CU(s.sp2.ensure(g.plane_page_stride * 4));
            CU(s.pf2.ensure(g.plane_page_stride * 4));
            sa.n_w = c0.n_w;
            sa.inv_n_f = 1.0f / (float)(c0.n_w * c0.n_h);
            sa.n_w2 = c1.n_w;
            sa.inv_n_f2 = 1.0f / (float)(c1.n_w * c1.n_h);
            sa.sp2 = s.sp2.as<uint32_t>();
            sa.pf2 = s.pf2.as<float>();
            CU(launch_window_stats(sa, 1, st));
            a.sp2 = sa.sp2;
            a.pf2 = sa.pf2;
        }
        a.acc_out = nullptr;
        CU(launch_scan_tc(gr.tc, a, 1, c->sm_count, st, &nl, s.acc.as<uint32_t>(), (int)(ch.group_pos + pos)));
    } else {
        CU(launch_scan_simt(a, 1, st, &nl));
    }
    c->launches += 2 + nl;
    CU(cudaMemcpyAsync(acc_host, s.acc.p, page_bytes * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return FOCR_OK;
}

// ------------------------------------------------------------------------------------------ compat shim
static std::mutex g_shim_mu;
static focr_ctx *g_shim_ctx = nullptr;
// The reference calls the kernel once per (page, offset, letter) (ncc.rs:587-702): the same needle comes back for every
// page.  The shim keeps the banks of the most recent needles (keyed by their bytes) so that a repeated needle costs no
// operand upload; a 1-template bank is a few KB of device memory.
struct ShimBank {
    std::vector<uint8_t> key;   // n_w, n_h (little-endian u16 each) + tight pixels
    focr_bank *bank;
    uint64_t stamp;
};
static std::vector<ShimBank> g_shim_banks;
static uint64_t g_shim_clock = 0;
constexpr size_t SHIM_BANKS = 4096;   // covers config 5's 3040 templates

static size_t shim_scan(uint8_t *reference, size_t r_w, size_t r_h, uint8_t *needle_u8, size_t N, size_t n_w,
                        size_t n_h, float threshold, focr_match *out, size_t n_out)
{
    std::lock_guard<std::mutex> lk(g_shim_mu);
    auto die = [](const char *what) {
        fprintf(stderr, "libfocr_b200 shim: %s: %s\n", what, focr_last_error());
        abort();
    };
    if (!g_shim_ctx) {
        const char *dev = getenv("FOCR_DEVICE");
        if (focr_ctx_create(dev ? atoi(dev) : 0, &g_shim_ctx) != FOCR_OK) die("focr_ctx_create");
    }
    if (n_out == 0 || n_w == 0 || n_h == 0 || n_w > N || r_w < n_w || r_h < n_h) return 0;
    // "returns n_out when full" (ncc.cpp:225-227) cannot be honoured beyond the library's per-list capacity; the reference's
    // callers use 1024 (ncc.rs:31,240).  Never return a truncated list as if it were complete.
    if (n_out > 4096) {
        fail(FOCR_ERR_UNSUPPORTED, "n_out = " + std::to_string(n_out) + " exceeds the supported 4096 matches per call");
        die("ncc_*_u8");
    }
    // unpad the needle rows (the bank API takes the tight canvas of ncc.rs:640)
    std::vector<uint8_t> key(4 + n_w * n_h);
    key[0] = (uint8_t)n_w, key[1] = (uint8_t)(n_w >> 8), key[2] = (uint8_t)n_h, key[3] = (uint8_t)(n_h >> 8);
    for (size_t y = 0; y < n_h; y++) memcpy(&key[4 + y * n_w], needle_u8 + y * N, n_w);
    focr_bank *bank = nullptr;
    for (auto &e : g_shim_banks)
        if (e.key == key) {
            bank = e.bank;
            e.stamp = ++g_shim_clock;
            break;
        }
    if (!bank) {
        uint64_t off = 0;
        uint16_t w16 = (uint16_t)n_w, h16 = (uint16_t)n_h;
        if (focr_bank_create(g_shim_ctx, key.data() + 4, &off, &w16, &h16, 1, &bank) != FOCR_OK) die("focr_bank_create");
        if (g_shim_banks.size() >= SHIM_BANKS) {   // evict the least recently used
            size_t lru = 0;
            for (size_t i = 1; i < g_shim_banks.size(); i++)
                if (g_shim_banks[i].stamp < g_shim_banks[lru].stamp) lru = i;
            focr_bank_destroy(g_shim_banks[lru].bank);
            g_shim_banks[lru] = ShimBank{std::move(key), bank, ++g_shim_clock};
        } else {
            g_shim_banks.push_back(ShimBank{std::move(key), bank, ++g_shim_clock});
        }
    }
    const uint32_t n_out32 = (uint32_t)n_out;
    std::vector<focr_match> tmp(n_out32);
    uint32_t cnt = 0;
    int rc = scan_host_impl(g_shim_ctx, bank, reference, r_w * r_h, (uint32_t)r_w, (uint32_t)r_h, 1, threshold, n_out32,
                            0 /* already inverted */, tmp.data(), &cnt);
    if (rc != FOCR_OK) die("scan");
    memcpy(out, tmp.data(), cnt * sizeof(focr_match));
    return cnt;
}

extern "C" size_t ncc_8_u8(uint8_t *reference, size_t r_w, size_t r_h, uint8_t *needle_u8, size_t n_w, size_t n_h,
                           uint32_t *, size_t, uint32_t *, double *, uint16_t *, float threshold, focr_match *out,
                           size_t n_out)
{
    return shim_scan(reference, r_w, r_h, needle_u8, 8, n_w, n_h, threshold, out, n_out);
}

extern "C" size_t ncc_16_u8(uint8_t *reference, size_t r_w, size_t r_h, uint8_t *needle_u8, size_t n_w, size_t n_h,
                            uint32_t *, size_t, uint32_t *, double *, uint16_t *, float threshold, focr_match *out,
                            size_t n_out)
{
    return shim_scan(reference, r_w, r_h, needle_u8, 16, n_w, n_h, threshold, out, n_out);
}
