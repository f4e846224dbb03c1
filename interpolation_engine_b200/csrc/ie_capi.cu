// C ABI of the engine (include/ie_b200.h): engine / table lifetime, host- and device-buffer
// batch entry points.  No torch types, no exceptions across the boundary, no CPU fallback.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <atomic>
#include <thread>
#include <vector>

#include "ie_host.hpp"
#include "ie_kernels.h"

namespace {

thread_local std::string g_err;

ie_status_t fail(ie_status_t st, const std::string& msg) { g_err = msg; return st; }
ie_status_t cuda_fail(cudaError_t e, const char* what) {
    return fail(e == cudaErrorMemoryAllocation ? IE_E_NOMEM : IE_E_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return cuda_fail(e_, #call); } while (0)

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes, cudaStream_t stream) {
        if (bytes <= cap) return cudaSuccess;
        if (p) { cudaStreamSynchronize(stream); cudaFree(p); p = nullptr; cap = 0; }
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
struct PinBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) { cudaFreeHost(p); p = nullptr; cap = 0; }
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocDefault);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    // grows to at least `bytes`, keeping the first `keep` bytes
    cudaError_t grow_keep(size_t bytes, size_t keep) {
        if (bytes <= cap) return cudaSuccess;
        void* q = nullptr;
        const size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaHostAlloc(&q, want, cudaHostAllocDefault);
        if (e != cudaSuccess) return e;
        if (p && keep) std::memcpy(q, p, keep);
        if (p) cudaFreeHost(p);
        p = q; cap = want;
        return cudaSuccess;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

constexpr uint32_t kDefaultExpansions = 4096;
constexpr uint32_t kDefaultResultBytes = 64u << 10;

}  // namespace

struct ie_engine {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t s_in = nullptr, s_out = nullptr;  // copy engines of the pipelined host-buffer path
    std::vector<cudaEvent_t> ev_in, ev_done;
    double expand = 2.0;  // output bytes per input byte the pipelined path provisions (adapts upward)
    uint64_t in_bias = 0; // tmpl_offs[0] of the batch in d_in: the arena may be a shard of a larger one, d_in holds bytes [offs[0], offs[n])
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // workspace
    DevBuf ws_zero, ws_list, ws_scratch;
    uint32_t tcap = 0;
    // host-API staging
    DevBuf d_in, d_in_offs, d_out, d_out_offs, d_out_lens, d_status, d_aux, d_info, d_mask, d_misc, d_esc;
    PinBuf h_out, h_out_offs, h_out_lens, h_status, h_aux, h_info;
    // small batches (one task at a time, the reference's interactive shape): one staging block each way
    DevBuf d_small_in, d_small_res;
    PinBuf h_small_in, h_small_res;
};

struct ie_table {
    ie_engine* e = nullptr;
    void* d_base = nullptr;       // every snapshot's image, back to back, followed by the view array
    size_t bytes = 0;
    IeTableView view{};           // snapshot 0 (ie_lookup_batch)
    const IeTableView* d_views = nullptr;
    uint32_t n_states = 1;
    bool has_balanced = false;    // some value holds properly nested groups of its own: rescan rounds can do work
    bool pooled = false;          // small table: allocated from the stream-ordered pool (no cudaMalloc / cudaFree per snapshot)
    cudaEvent_t ready = nullptr;  // upload finished (calls on a caller's stream wait for it)
    // in-place mutation (ie_table_set / ie_table_delete): the bump-allocated arena behind the slots, its header, and
    // the per-snapshot count of non-empty slots
    uint64_t arena_off = 0, arena_bytes = 0;
    IeTableHeader* d_hdr = nullptr;
    uint32_t* d_used = nullptr;
    double build_ms = 0.0;        // device time of the build kernels (device-built tables)
};

namespace {

// Lays the per-batch workspace out for n items; returns the kernel-side view.
ie_status_t prepare_workspace(ie_engine* e, uint64_t n, uint32_t tcap, bool need_general, IeWorkspace* ws, uint64_t tiles,
                              bool need_rounds = false) {
    static_assert(sizeof(IeRoundCtl) <= 64, "round control block has a 64-byte slot");
    const size_t zero_bytes = 128 + (size_t)(tiles + 1) * sizeof(uint64_t);  // [counters 64][round control 64][tile states]
    CU(e->ws_zero.ensure(zero_bytes, e->stream));
    ws->zero_base = (uint8_t*)e->ws_zero.p;
    ws->zero_bytes = zero_bytes;
    ws->tile_counter = (uint32_t*)e->ws_zero.p;
    ws->general_count = ws->tile_counter + 1;
    ws->overflow = ws->tile_counter + 2;
    ws->fix_count = ws->tile_counter + 3;
    ws->retry_count = ws->tile_counter + 4;
    ws->retry_list = nullptr;
    ws->fix_list = nullptr;
    ws->tile_first = nullptr;
    ws->tile_state = (uint64_t*)((uint8_t*)e->ws_zero.p + 128);
    ws->round_ctl = nullptr;
    ws->round_list[0] = ws->round_list[1] = ws->round_list[2] = nullptr;
    ws->round_offs = nullptr;
    ws->general_list = nullptr;
    ws->scratch = nullptr;
    ws->general_workers = IE_GENERAL_WORKERS;
    if (need_general) {
        const uint64_t m = std::max<uint64_t>(n, 1);
        // general list, retry list [, three round lists, round offsets (8-byte aligned: 5 m u32 precede them, m padded even)]
        const uint64_t me = (m + 1) & ~uint64_t(1);
        CU(e->ws_list.ensure((size_t)me * (need_rounds ? 6 : 2) * sizeof(uint32_t) + (need_rounds ? (size_t)(m + 1) * sizeof(uint64_t) : 0), e->stream));
        ws->general_list = (uint32_t*)e->ws_list.p;
        ws->retry_list = ws->general_list + me;
        if (need_rounds) {
            ws->round_ctl = (IeRoundCtl*)((uint8_t*)e->ws_zero.p + 64);
            for (int k = 0; k < 3; ++k) ws->round_list[k] = ws->general_list + me * (2 + k);
            ws->round_offs = (uint64_t*)(ws->general_list + me * 6);  // me even: 24 me bytes, 8-byte aligned
        }
        if (tcap > e->tcap || !e->ws_scratch.p) {
            CU(e->ws_scratch.ensure((size_t)IE_GENERAL_WORKERS * ie_general_worker_bytes(tcap), e->stream));
            e->tcap = tcap;
        }
        ws->scratch = (uint8_t*)e->ws_scratch.p;
    }
    return IE_OK;
}

// Groups per template x 16, estimated from the '{' density of a sample of the caller's text (host-buffer calls): sizes
// the tiles so that brace-dense templates do not overflow a tile's event / segment tables.
uint64_t sample_groups_x16(const uint8_t* tmpl, uint64_t in_bytes, uint64_t n) {
    if (!n || !in_bytes) return 0;
    const uint64_t take = std::min<uint64_t>(in_bytes, 32u << 10);
    uint64_t opens = 0;
    for (uint64_t i = 0; i < take; ++i) opens += tmpl[i] == '{';
    const long double per_byte = (long double)opens / (long double)take;
    return (uint64_t)(per_byte * (long double)in_bytes / (long double)n * 16.0L + 0.5L);
}

// Rescan rounds of the host-buffer calls: two by default (their few extra launches hide behind the PCIe copies);
// the device-buffer call runs exactly limits->rescan_rounds of them (default none: every launch is on its clock).
uint32_t host_rounds(const ie_limits* in) { return (in && in->rescan_rounds) ? in->rescan_rounds : 2u; }

void resolve_limits(const ie_limits* in, uint32_t* max_exp, uint32_t* tcap) {
    *max_exp = (in && in->max_expansions) ? in->max_expansions : kDefaultExpansions;
    uint32_t rb = (in && in->max_result_bytes) ? in->max_result_bytes : kDefaultResultBytes;
    *tcap = (rb + 15u) & ~15u;
}

}  // namespace

extern "C" {

const char* ie_last_error(void) { return g_err.c_str(); }

int ie_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

ie_status_t ie_engine_create(int device, ie_engine** out) {
    if (!out) return fail(IE_E_INVALID, "ie_engine_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t ce = cudaGetDeviceCount(&n);
    if (ce != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(IE_E_CUDA, "ie_engine_create: no CUDA device available (this engine has no CPU fallback)");
    }
    if (device < 0 || device >= n) return fail(IE_E_INVALID, "ie_engine_create: device index out of range");
    CU(cudaSetDevice(device));
    ie_engine* e = new (std::nothrow) ie_engine();
    if (!e) return fail(IE_E_NOMEM, "ie_engine_create: out of host memory");
    e->device = device;
    cudaError_t err = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking);
    if (err == cudaSuccess) err = cudaStreamCreateWithFlags(&e->s_in, cudaStreamNonBlocking);
    if (err == cudaSuccess) err = cudaStreamCreateWithFlags(&e->s_out, cudaStreamNonBlocking);
    if (err == cudaSuccess) err = cudaEventCreate(&e->ev0);
    if (err == cudaSuccess) err = cudaEventCreate(&e->ev1);
    if (err == cudaSuccess) {  // the pool keeps what small tables release instead of returning it to the driver
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t keep = 256ull << 20;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    if (err != cudaSuccess) { ie_engine_destroy(e); return cuda_fail(err, "ie_engine_create"); }
    *out = e;
    return IE_OK;
}

void ie_engine_destroy(ie_engine* e) {
    if (!e) return;
    ie_host::drop_engine(e);
    cudaSetDevice(e->device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    for (DevBuf* b : {&e->ws_zero, &e->ws_list, &e->ws_scratch, &e->d_in, &e->d_in_offs, &e->d_out, &e->d_out_offs, &e->d_out_lens,
                      &e->d_status, &e->d_aux, &e->d_info, &e->d_mask, &e->d_misc, &e->d_esc, &e->d_small_in, &e->d_small_res})
        b->release();
    for (PinBuf* b : {&e->h_out, &e->h_out_offs, &e->h_out_lens, &e->h_status, &e->h_aux, &e->h_info, &e->h_small_in, &e->h_small_res})
        b->release();
    if (e->ev0) cudaEventDestroy(e->ev0);
    if (e->ev1) cudaEventDestroy(e->ev1);
    for (cudaEvent_t ev : e->ev_in) cudaEventDestroy(ev);
    for (cudaEvent_t ev : e->ev_done) cudaEventDestroy(ev);
    if (e->s_in) cudaStreamDestroy(e->s_in);
    if (e->s_out) cudaStreamDestroy(e->s_out);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
}

void* ie_engine_stream(ie_engine* e) { return e ? (void*)e->stream : nullptr; }

ie_status_t ie_engine_sync(ie_engine* e) {
    if (!e) return fail(IE_E_INVALID, "ie_engine_sync: engine is NULL");
    CU(cudaSetDevice(e->device));
    CU(cudaStreamSynchronize(e->stream));
    return IE_OK;
}

// Uploads the images of n_states snapshots (one allocation: images at 256-byte aligned offsets, then the
// IeTableView array the kernels index by snapshot) and fills *out.
static ie_status_t upload_tables(ie_engine* e, const std::vector<std::vector<uint8_t>>& images, const std::vector<uint32_t>& caps,
                                 const std::vector<uint32_t>& counts, ie_table** out, const char* who) {
    const size_t S = images.size();
    std::vector<size_t> at(S + 1, 0);
    for (size_t s = 0; s < S; ++s) at[s + 1] = at[s] + ((images[s].size() + 255) & ~size_t(255));
    // behind the images: slack for values and keys that ie_table_set appends, the view array, the header, the used counts
    size_t slack = 0;
    if (S == 1) slack = std::max<size_t>(4096, (images[0].size() / 4 + 15) & ~size_t(15));
    const size_t arena_at = at[S], views_at = arena_at + slack, hdr_at = views_at + S * sizeof(IeTableView);
    const size_t used_at = hdr_at + sizeof(IeTableHeader), total = used_at + S * sizeof(uint32_t);
    CU(cudaSetDevice(e->device));
    ie_table* t = new (std::nothrow) ie_table();
    if (!t) return fail(IE_E_NOMEM, std::string(who) + ": out of host memory");
    t->e = e;
    t->bytes = total;
    t->n_states = (uint32_t)S;
    // Small tables (one snapshot of an interactive run) come from the device's stream-ordered pool and are uploaded
    // with ONE copy and no host synchronisation: a pageable cudaMemcpyAsync returns once the bytes are staged, and
    // everything that uses the table is ordered behind it (same stream, or the `ready` event).
    // Mid-size tables (a few thousand to 64 k inserts) also come from the pool - a cudaMalloc / cudaFree pair costs about
    // 10 ms per snapshot, which a caller that repacks per task (replace_map: inserts + captures) pays every iteration -
    // but keep their own two copies instead of a second host image.
    const bool pooled = S == 1 && total <= (48u << 20);
    const bool small = pooled && total <= (1u << 20);
    cudaError_t err = pooled ? cudaMallocAsync(&t->d_base, total, e->stream) : cudaMalloc(&t->d_base, total);
    if (err != cudaSuccess) { delete t; return cuda_fail(err, who); }
    t->pooled = pooled;
    std::vector<IeTableView> views(S);
    // trailer: views, header (nothing of the arena used yet), non-empty slots per snapshot
    std::vector<uint8_t> trailer(total - views_at, 0);
    for (size_t s = 0; s < S; ++s) {
        views[s].base = (const uint8_t*)t->d_base + at[s];
        views[s].mask = caps[s] - 1;
        views[s].n_entries = counts[s];
        uint32_t used = 0;
        const IeSlot* sl = reinterpret_cast<const IeSlot*>(images[s].data());
        for (uint32_t k = 0; k < caps[s]; ++k) used += sl[k].key_len != IE_SLOT_EMPTY;
        std::memcpy(trailer.data() + (used_at - views_at) + s * sizeof(uint32_t), &used, sizeof used);
    }
    std::memcpy(trailer.data(), views.data(), S * sizeof(IeTableView));
    if (small || S > 1) {  // images and the trailer in one staged block, one copy
        std::vector<uint8_t> all(total, 0);
        for (size_t s = 0; s < S; ++s) std::memcpy(all.data() + at[s], images[s].data(), images[s].size());
        std::memcpy(all.data() + views_at, trailer.data(), trailer.size());
        err = cudaMemcpyAsync(t->d_base, all.data(), all.size(), cudaMemcpyHostToDevice, e->stream);
        if (err == cudaSuccess && !small) err = cudaStreamSynchronize(e->stream);
    } else {               // one large snapshot: no second host copy of its image
        err = cudaMemcpyAsync(t->d_base, images[0].data(), images[0].size(), cudaMemcpyHostToDevice, e->stream);
        if (err == cudaSuccess) err = cudaMemcpyAsync((uint8_t*)t->d_base + views_at, trailer.data(), trailer.size(), cudaMemcpyHostToDevice, e->stream);
        if (err == cudaSuccess && !pooled) err = cudaStreamSynchronize(e->stream);
    }
    if (err == cudaSuccess) err = cudaEventCreateWithFlags(&t->ready, cudaEventDisableTiming);
    if (err == cudaSuccess) err = cudaEventRecord(t->ready, e->stream);
    if (err != cudaSuccess) { if (pooled) cudaFreeAsync(t->d_base, e->stream); else cudaFree(t->d_base); delete t; return cuda_fail(err, who); }
    t->arena_off = arena_at;
    t->arena_bytes = slack;
    t->d_hdr = (IeTableHeader*)((uint8_t*)t->d_base + hdr_at);
    t->d_used = (uint32_t*)((uint8_t*)t->d_base + used_at);
    t->view = views[0];
    t->d_views = (const IeTableView*)((uint8_t*)t->d_base + views_at);
    *out = t;
    return IE_OK;
}

static ie_status_t build_on_device(ie_engine* e, uint64_t n_states, const uint64_t* state_offs, const uint8_t* keys, const uint64_t* key_offs,
                                   const uint8_t* vals, const uint64_t* val_offs, const uint8_t* tags, const char* hhmm, const char* hhmmss,
                                   bool compact, ie_table** out, const char* who);

ie_status_t ie_table_pack(ie_engine* e, uint64_t n, const uint8_t* keys, const uint64_t* key_offs, const uint8_t* vals,
                          const uint64_t* val_offs, const uint8_t* tags, const char* hhmm, const char* hhmmss, ie_table** out) {
    if (!e || !out || (n && (!keys || !key_offs || !vals || !val_offs || !tags)))
        return fail(IE_E_INVALID, "ie_table_pack: NULL argument");
    *out = nullptr;
    if (n >= 4096) {  // large snapshots are hashed, classified and laid out on the device (ie_table_build.cu)
        const uint64_t so[2] = {0, n};
        return build_on_device(e, 1, so, keys, key_offs, vals, val_offs, tags, hhmm, hhmmss, /*compact=*/false, out, "ie_table_pack");
    }
    std::vector<std::vector<uint8_t>> images(1);
    std::vector<uint32_t> caps(1, 0), counts(1, (uint32_t)n);
    std::string why;
    bool balanced = false;
    if (!ie_host::build_table_image(n, keys, key_offs, vals, val_offs, tags, hhmm, hhmmss, &images[0], &caps[0], &why, false, &balanced))
        return fail(IE_E_INVALID, "ie_table_pack: " + why);
    ie_status_t st = upload_tables(e, images, caps, counts, out, "ie_table_pack");
    if (st == IE_OK) (*out)->has_balanced = balanced;
    return st;
}

// Device-side build (ie_table_build.cu): the raw packed arrays go to the device as they are, the tables of all
// snapshots are built there by one thread per insert.  Host work is O(snapshots): capacities and the layout.
static ie_status_t build_on_device(ie_engine* e, uint64_t n_states, const uint64_t* state_offs, const uint8_t* keys, const uint64_t* key_offs,
                                   const uint8_t* vals, const uint64_t* val_offs, const uint8_t* tags, const char* hhmm, const char* hhmmss,
                                   bool compact, ie_table** out, const char* who) {
    const uint64_t n = state_offs[n_states];
    if (n > 0x3FFFFFFFull) return fail(IE_E_INVALID, std::string(who) + ": too many inserts (max 2^30 - 1)");
    const size_t hm = hhmm ? std::strlen(hhmm) : 0, hs = hhmmss ? std::strlen(hhmmss) : 0;
    if (hm > 64 || hs > 64) return fail(IE_E_INVALID, std::string(who) + ": clock rendering longer than 64 bytes");
    const uint64_t kbytes = n ? key_offs[n] : 0, vbytes = n ? val_offs[n] : 0;
    // layout: slot arrays (load factor <= 0.5 for packed-side-by-side snapshots, <= 0.25 for a single one), then the arena
    std::vector<uint64_t> slot_base(n_states);
    std::vector<uint32_t> slot_cap(n_states);
    uint64_t slots = 0;
    for (uint64_t s = 0; s < n_states; ++s) {
        if (state_offs[s + 1] < state_offs[s]) return fail(IE_E_INVALID, std::string(who) + ": state offsets not monotone");
        const uint64_t cnt = state_offs[s + 1] - state_offs[s] + 2;
        const uint64_t want = (!compact && cnt <= (1ull << 22)) ? cnt * 4 : cnt * 2;
        uint64_t cap = 16;
        while (cap < want) cap <<= 1;
        if (cap > (1ull << 31)) return fail(IE_E_INVALID, std::string(who) + ": table too large");
        slot_base[s] = slots;
        slot_cap[s] = (uint32_t)cap;
        slots += cap;
    }
    const uint64_t items = n + 2 * n_states;
    // every item longer than 16 bytes takes its length rounded up to 16 from the arena; + slack for later ie_table_set calls
    const uint64_t arena_need = kbytes + vbytes + 15 * items + 2 * 64 * n_states;
    const uint64_t arena_bytes = ((arena_need + std::max<uint64_t>(64u << 10, arena_need / 8)) + 15) & ~uint64_t(15);
    const uint64_t slots_bytes = slots * sizeof(IeSlot);
    const uint64_t views_at = slots_bytes + arena_bytes, hdr_at = views_at + n_states * sizeof(IeTableView);
    const uint64_t used_at = hdr_at + sizeof(IeTableHeader), total = used_at + n_states * sizeof(uint32_t);
    if ((total >> 4) > 0xFFFFFFFFull) return fail(IE_E_INVALID, std::string(who) + ": table exceeds 64 GiB");
    CU(cudaSetDevice(e->device));
    cudaStream_t s = e->stream;
    ie_table* t = new (std::nothrow) ie_table();
    if (!t) return fail(IE_E_NOMEM, std::string(who) + ": out of host memory");
    t->e = e;
    t->bytes = total;
    t->n_states = (uint32_t)n_states;
    t->pooled = total <= (48u << 20);
    cudaError_t err = t->pooled ? cudaMallocAsync(&t->d_base, total, s) : cudaMalloc(&t->d_base, total);
    if (err != cudaSuccess) { delete t; return cuda_fail(err, who); }
    // staging area of the inputs (stream-ordered pool; released behind the build kernels)
    auto up16 = [](uint64_t x) { return (x + 15) & ~uint64_t(15); };
    const uint64_t o_so = 0, o_sb = o_so + up16((n_states + 1) * 8), o_sc = o_sb + up16(n_states * 8), o_ko = o_sc + up16(n_states * 4);
    const uint64_t o_vo = o_ko + up16((n + 1) * 8), o_tg = o_vo + up16((n + 1) * 8), o_ck = o_tg + up16(n), o_k = o_ck + 160;
    const uint64_t o_v = o_k + up16(kbytes), o_sl = o_v + up16(vbytes), stage_total = o_sl + up16(items * 4);
    uint8_t* d_stage = nullptr;
    err = cudaMallocAsync((void**)&d_stage, stage_total, s);
    auto bail = [&](cudaError_t ce) {
        if (d_stage) cudaFreeAsync(d_stage, s);
        if (t->pooled) cudaFreeAsync(t->d_base, s); else cudaFree(t->d_base);
        delete t;
        return cuda_fail(ce, who);
    };
    if (err != cudaSuccess) { d_stage = nullptr; return bail(err); }
    uint8_t clock[160] = {0};
    std::memcpy(clock, "HH:MM", 5);
    std::memcpy(clock + 8, "HH:MM:SS", 8);
    if (hhmm) std::memcpy(clock + 16, hhmm, hm);
    if (hhmmss) std::memcpy(clock + 80, hhmmss, hs);
    const uint64_t zero_off[1] = {0};
    cudaEvent_t ev_a = nullptr, ev_b = nullptr;
    cudaEventCreate(&ev_a);
    cudaEventCreate(&ev_b);
#define UP(dst_off, src, bytes) if (err == cudaSuccess && (bytes)) err = cudaMemcpyAsync(d_stage + (dst_off), (src), (bytes), cudaMemcpyHostToDevice, s)
    UP(o_so, state_offs, (n_states + 1) * 8);
    UP(o_sb, slot_base.data(), n_states * 8);
    UP(o_sc, slot_cap.data(), n_states * 4);
    UP(o_ko, n ? key_offs : zero_off, (n + 1) * 8);
    UP(o_vo, n ? val_offs : zero_off, (n + 1) * 8);
    UP(o_tg, tags, n);
    UP(o_ck, clock, sizeof clock);
    UP(o_k, keys, kbytes);
    UP(o_v, vals, vbytes);
#undef UP
    // empty slots are all-ones (key_len = IE_SLOT_EMPTY, claim word = none); header and used counts start at zero
    if (err == cudaSuccess) err = cudaMemsetAsync(t->d_base, 0xFF, slots_bytes, s);
    if (err == cudaSuccess) err = cudaMemsetAsync((uint8_t*)t->d_base + hdr_at, 0, total - hdr_at, s);
    IeBuildArgs a{};
    a.base = (uint8_t*)t->d_base;
    a.arena_off = slots_bytes;
    a.arena_bytes = arena_bytes;
    a.hdr = (IeTableHeader*)((uint8_t*)t->d_base + hdr_at);
    a.used = (uint32_t*)((uint8_t*)t->d_base + used_at);
    a.state_offs = (const uint64_t*)(d_stage + o_so);
    a.slot_base = (const uint64_t*)(d_stage + o_sb);
    a.slot_cap = (const uint32_t*)(d_stage + o_sc);
    a.slot_of = (uint32_t*)(d_stage + o_sl);
    a.keys = d_stage + o_k; a.key_offs = (const uint64_t*)(d_stage + o_ko);
    a.vals = d_stage + o_v; a.val_offs = (const uint64_t*)(d_stage + o_vo);
    a.tags = d_stage + o_tg;
    a.clock = d_stage + o_ck;
    a.hhmm_len = (uint32_t)hm; a.hhmmss_len = (uint32_t)hs;
    a.n = n;
    a.n_states = (uint32_t)n_states;
    a.with_clock = (hhmm ? 1u : 0u) | (hhmmss ? 2u : 0u);
    if (err == cudaSuccess) err = cudaEventRecord(ev_a, s);
    if (err == cudaSuccess) err = ie_launch_table_build(a, (IeTableView*)((uint8_t*)t->d_base + views_at), s);
    if (err == cudaSuccess) err = cudaEventRecord(ev_b, s);
    IeTableHeader hdr{};
    IeTableView v0{};
    if (err == cudaSuccess) err = cudaMemcpyAsync(&hdr, a.hdr, sizeof hdr, cudaMemcpyDeviceToHost, s);
    if (err == cudaSuccess) err = cudaMemcpyAsync(&v0, (uint8_t*)t->d_base + views_at, sizeof v0, cudaMemcpyDeviceToHost, s);
    if (err == cudaSuccess) err = cudaStreamSynchronize(s);
    float ms = 0.f;
    if (err == cudaSuccess) cudaEventElapsedTime(&ms, ev_a, ev_b);
    cudaEventDestroy(ev_a);
    cudaEventDestroy(ev_b);
    if (err != cudaSuccess) return bail(err);
    cudaFreeAsync(d_stage, s);
    d_stage = nullptr;
    if (hdr.error) {
        const uint32_t code = hdr.error;
        bail(cudaSuccess);
        return fail(IE_E_INVALID, std::string(who) + ((code & 1) ? ": offsets not monotone, key longer than 4 GiB, value longer than 32 MiB or bad tag"
                                                                   : ": internal error while building the table (code " + std::to_string(code) + ")"));
    }
    err = cudaEventCreateWithFlags(&t->ready, cudaEventDisableTiming);
    if (err == cudaSuccess) err = cudaEventRecord(t->ready, s);
    if (err != cudaSuccess) return bail(err);
    t->view = v0;
    t->d_views = (const IeTableView*)((uint8_t*)t->d_base + views_at);
    t->has_balanced = (hdr.flags & 1u) != 0;
    t->arena_off = slots_bytes;
    t->arena_bytes = arena_bytes;
    t->d_hdr = a.hdr;
    t->d_used = a.used;
    t->build_ms = ms;
    *out = t;
    return IE_OK;
}

ie_status_t ie_table_pack_many(ie_engine* e, uint64_t n_states, const uint64_t* state_offs, const uint8_t* keys, const uint64_t* key_offs,
                               const uint8_t* vals, const uint64_t* val_offs, const uint8_t* tags, const char* hhmm, const char* hhmmss,
                               ie_table** out) {
    if (!e || !out || !n_states || !state_offs) return fail(IE_E_INVALID, "ie_table_pack_many: NULL argument or no snapshots");
    if (n_states > 0x7FFFFFFFull) return fail(IE_E_INVALID, "ie_table_pack_many: too many snapshots");
    if (state_offs[n_states] && (!keys || !key_offs || !vals || !val_offs || !tags)) return fail(IE_E_INVALID, "ie_table_pack_many: NULL argument");
    *out = nullptr;
    return build_on_device(e, n_states, state_offs, keys, key_offs, vals, val_offs, tags, hhmm, hhmmss, /*compact=*/true, out, "ie_table_pack_many");
}

double ie_table_build_ms(const ie_table* t) { return t ? t->build_ms : 0.0; }

// set_interpdata / delete_interpdata (interp.rs:139-145) on the device table, in place.
static ie_status_t mutate(ie_engine* e, ie_table* t, uint32_t state, uint64_t n, const uint8_t* keys, const uint64_t* key_offs, const uint8_t* vals,
                          const uint64_t* val_offs, const uint8_t* tags, const uint32_t* entries, const char* who) {
    if (!e || !t || (n && (!keys && key_offs && key_offs[n]))) return fail(IE_E_INVALID, std::string(who) + ": NULL argument");
    if (n && !key_offs) return fail(IE_E_INVALID, std::string(who) + ": NULL argument");
    if (t->e != e) return fail(IE_E_INVALID, std::string(who) + ": the table belongs to another engine");
    if (state != IE_ALL_STATES && state >= t->n_states) return fail(IE_E_INVALID, std::string(who) + ": snapshot index out of range");
    if (n > 0xFFFFFFFFull) return fail(IE_E_INVALID, std::string(who) + ": too many operations");
    if (!n) return IE_OK;
    const bool is_set = vals != nullptr || val_offs != nullptr;
    if (is_set && (!val_offs || !tags)) return fail(IE_E_INVALID, std::string(who) + ": NULL argument");
    CU(cudaSetDevice(e->device));
    cudaStream_t s = e->stream;
    const uint64_t kbytes = key_offs[n], vbytes = is_set ? val_offs[n] : 0;
    std::vector<uint8_t> flags(is_set ? n : 0);
    bool balanced = false;
    for (uint64_t i = 0; i < n; ++i) {
        if (key_offs[i + 1] < key_offs[i] || key_offs[i + 1] - key_offs[i] >= IE_SLOT_TOMB) return fail(IE_E_INVALID, std::string(who) + ": bad key offsets");
        if (!is_set) continue;
        if (val_offs[i + 1] < val_offs[i] || val_offs[i + 1] - val_offs[i] > IE_VLEN_MAX || tags[i] > IE_TAG_OBJECT)
            return fail(IE_E_INVALID, std::string(who) + ": bad value offsets, value longer than 32 MiB or bad tag");
        flags[i] = (uint8_t)ie_classify_value(vals + val_offs[i], val_offs[i + 1] - val_offs[i]);
        balanced |= (flags[i] & IE_VF_BALANCED) != 0;
    }
    auto up16 = [](uint64_t x) { return (x + 15) & ~uint64_t(15); };
    const uint64_t o_ko = 0, o_vo = o_ko + up16((n + 1) * 8), o_tg = o_vo + up16((n + 1) * 8), o_fl = o_tg + up16(n), o_en = o_fl + up16(n);
    const uint64_t o_k = o_en + up16(n * 4), o_v = o_k + up16(kbytes), stage_total = o_v + up16(vbytes) + 16;
    // one staged block, one copy
    std::vector<uint8_t> blk(stage_total, 0);
    std::memcpy(blk.data() + o_ko, key_offs, (n + 1) * 8);
    if (kbytes) std::memcpy(blk.data() + o_k, keys, kbytes);
    if (is_set) {
        std::memcpy(blk.data() + o_vo, val_offs, (n + 1) * 8);
        std::memcpy(blk.data() + o_tg, tags, n);
        std::memcpy(blk.data() + o_fl, flags.data(), n);
        if (entries) std::memcpy(blk.data() + o_en, entries, n * 4);
        if (vbytes) std::memcpy(blk.data() + o_v, vals, vbytes);
    }
    uint8_t* d_stage = nullptr;
    CU(cudaMallocAsync((void**)&d_stage, stage_total, s));
    cudaError_t err = cudaMemcpyAsync(d_stage, blk.data(), stage_total, cudaMemcpyHostToDevice, s);
    IeMutateArgs m{};
    m.base = (uint8_t*)t->d_base;
    m.arena_off = t->arena_off;
    m.arena_bytes = t->arena_bytes;
    m.hdr = t->d_hdr;
    m.views = t->d_views;
    m.used_slots = t->d_used;
    m.state = state == IE_ALL_STATES ? 0u : state;
    m.all_states = state == IE_ALL_STATES ? 1u : 0u;
    m.n_ops = (uint32_t)n;
    m.keys = d_stage + o_k; m.key_offs = (const uint64_t*)(d_stage + o_ko);
    m.vals = is_set ? d_stage + o_v : nullptr; m.val_offs = (const uint64_t*)(d_stage + o_vo);
    m.tags = d_stage + o_tg; m.flags = d_stage + o_fl;
    m.entries = (is_set && entries) ? (const uint32_t*)(d_stage + o_en) : nullptr;
    if (err == cudaSuccess) err = ie_launch_table_mutate(m, t->n_states, s);
    IeTableHeader hdr{};
    if (err == cudaSuccess) err = cudaMemcpyAsync(&hdr, t->d_hdr, sizeof hdr, cudaMemcpyDeviceToHost, s);
    if (err == cudaSuccess && hdr.error) {}  // (read below, after the synchronisation)
    if (err == cudaSuccess) err = cudaStreamSynchronize(s);
    cudaFreeAsync(d_stage, s);
    if (err != cudaSuccess) return cuda_fail(err, who);
    if (balanced) t->has_balanced = true;
    if (hdr.error) {
        const uint32_t zero = 0;  // the table stays usable: clear the flag, report which room ran out
        cudaMemcpyAsync(&t->d_hdr->error, &zero, sizeof zero, cudaMemcpyHostToDevice, s);
        cudaStreamSynchronize(s);
        return fail(IE_E_OVERFLOW, std::string(who) + ((hdr.error & 8) ? ": the snapshot's slot array is more than three quarters full"
                                                                       : ": the table's arena is full") +
                                       " - the operations that did not fit were skipped; pack the snapshot again");
    }
    return IE_OK;
}

ie_status_t ie_table_set(ie_engine* e, ie_table* t, uint32_t state, uint64_t n, const uint8_t* keys, const uint64_t* key_offs, const uint8_t* vals,
                         const uint64_t* val_offs, const uint8_t* tags, const uint32_t* entries) {
    static const uint8_t none = 0;
    if (n && !val_offs) return fail(IE_E_INVALID, "ie_table_set: NULL argument");
    return mutate(e, t, state, n, keys, key_offs, vals ? vals : &none, val_offs, tags, entries, "ie_table_set");
}

ie_status_t ie_table_delete(ie_engine* e, ie_table* t, uint32_t state, uint64_t n, const uint8_t* keys, const uint64_t* key_offs) {
    return mutate(e, t, state, n, keys, key_offs, nullptr, nullptr, nullptr, nullptr, "ie_table_delete");
}

uint32_t ie_table_states(const ie_table* t) { return t ? t->n_states : 0; }

void ie_table_free(ie_table* t) {
    if (!t) return;
    cudaSetDevice(t->e->device);
    if (t->ready) cudaEventDestroy(t->ready);
    if (t->pooled) {  // stream-ordered: freed once everything queued on the engine's stream so far has run
        if (t->d_base) cudaFreeAsync(t->d_base, t->e->stream);
    } else {
        cudaStreamSynchronize(t->e->stream);
        if (t->d_base) cudaFree(t->d_base);
    }
    delete t;
}

uint64_t ie_table_device_bytes(const ie_table* t) { return t ? t->bytes : 0; }

static ie_status_t resolve_device(ie_engine* e, const ie_table* t, const uint8_t* d_tmpl, const uint64_t* d_tmpl_offs, uint64_t n,
                                  const ie_limits* limits, uint8_t* d_out, uint64_t out_capacity, uint64_t* d_out_offs,
                                  uint32_t* d_out_lens, int32_t* d_status, uint32_t* d_aux, ie_batch_info* d_info, uint64_t out_bias,
                                  cudaStream_t s, uint64_t avg_bytes = 0, uint32_t rounds = 0, uint64_t groups_x16 = 0) {
    uint32_t max_exp, tcap;
    resolve_limits(limits, &max_exp, &tcap);
    if (!avg_bytes && limits) avg_bytes = limits->avg_template_bytes;
    if (!groups_x16 && limits) groups_x16 = (uint64_t)limits->avg_template_groups * 16;
    const uint32_t tt = ie_pick_tile(avg_bytes, groups_x16);
    IeWorkspace ws;
    if (n * t->n_states >= 0xFFFFFFFFull) return fail(IE_E_INVALID, "resolve: at most 2^32-2 (snapshot, template) pairs per batch");
    if (s != e->stream && t->ready) CU(cudaStreamWaitEvent(s, t->ready, 0));
    if (rounds > 3) rounds = 3;
    if (!t->has_balanced) rounds = 0;  // no value could be spliced: the rounds would be empty launches
    ie_status_t st = prepare_workspace(e, n * t->n_states, tcap, true, &ws, 0, rounds != 0);
    if (st != IE_OK) return st;
    CU(ie_launch_resolve(t->d_views, t->n_states, d_tmpl, d_tmpl_offs, n, d_out, out_capacity, d_out_offs, d_out_lens, d_status, d_aux, ws, d_info,
                         max_exp, tcap, out_bias, tt, rounds, t->bytes, s));
    return IE_OK;
}

ie_status_t ie_resolve_batch_device(ie_engine* e, const ie_table* t, const uint8_t* d_tmpl, const uint64_t* d_tmpl_offs, uint64_t n,
                                    const ie_limits* limits, uint8_t* d_out, uint64_t out_capacity, uint64_t* d_out_offs,
                                    uint32_t* d_out_lens, int32_t* d_status, uint32_t* d_aux, ie_batch_info* d_info, void* stream) {
    if (!e || !t || !d_info || (n && (!d_tmpl_offs || !d_out_offs || !d_out_lens || !d_status || !d_aux)))
        return fail(IE_E_INVALID, "ie_resolve_batch_device: NULL argument");
    if (n >= 0xFFFFFFFFull) return fail(IE_E_INVALID, "ie_resolve_batch_device: at most 2^32-2 templates per batch");
    CU(cudaSetDevice(e->device));
    return resolve_device(e, t, d_tmpl, d_tmpl_offs, n, limits, d_out, out_capacity, d_out_offs, d_out_lens, d_status, d_aux, d_info, 0,
                          stream ? (cudaStream_t)stream : e->stream, 0, limits ? limits->rescan_rounds : 0);
}

// Host-buffer batches of at least 64 Ki templates are cut into chunks (of up to this many templates) whose H2D copy,
// kernels and D2H copies overlap on three streams (PCIe is full duplex: 55 GB/s each way alone, 95 GB/s combined on
// the measured box; the kernels hide behind the copies and never touch a copy engine themselves).
static constexpr uint64_t kPipeChunk = 1u << 17;

// A chunk's ie_batch_info goes to the host through a one-thread kernel that writes pinned (device-mapped) host memory,
// NOT through a device-to-host copy: a copy on the compute stream queues behind the 18 MB arena transfers of the
// previous chunks on the same copy engine and stalls the next chunk's kernels (measured: the compute stream finished
// at 7.6 ms of an 8.0 ms batch although its kernels take 0.4 ms).
__global__ void ie_publish_info_kernel(const ie_batch_info* __restrict__ d_info, ie_batch_info* __restrict__ h_info_mapped) {
    *h_info_mapped = *d_info;
    __threadfence_system();
}

// Returns IE_OK with *done = false when a chunk's provisioned output region was too small (the caller
// then reruns the batch unpipelined with exact sizes, and e->expand has been raised).
static ie_status_t resolve_pipelined(ie_engine* e, const ie_table* t, const uint8_t* tmpl, const uint64_t* tmpl_offs, uint64_t n,
                                     const ie_limits* limits, ie_result* res, bool* done) {
    *done = false;
    // Chunk schedule: two quarter-size chunks and one half-size chunk first (the device-to-host engine, the long pole,
    // starts after 0.15 ms instead of 0.5), then full chunks of kPipeChunk templates (fewer, larger copies).
    std::vector<uint64_t> cut{0};
    for (uint64_t step : {kPipeChunk / 4, kPipeChunk / 4, kPipeChunk / 2})
        if (cut.back() + step < n) cut.push_back(cut.back() + step);
    while (cut.back() < n) cut.push_back(std::min(n, cut.back() + kPipeChunk));
    const uint64_t K = cut.size() - 1;
    const uint64_t base0 = tmpl_offs[0], in_bytes = tmpl_offs[n] - base0;
    e->in_bias = base0;
    cudaStream_t sc = e->stream;
    while (e->ev_in.size() < K) {
        cudaEvent_t a, b;
        CU(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
        e->ev_in.push_back(a);
        CU(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
        e->ev_done.push_back(b);
    }
    // per-chunk output regions at fixed bases: the results need not be contiguous, every template has (offset, length)
    std::vector<uint64_t> base(K + 1, 0);
    for (uint64_t k = 0; k < K; ++k) {
        const uint64_t lo = cut[k], hi = cut[k + 1];
        const uint64_t cap = (uint64_t)((double)(tmpl_offs[hi] - tmpl_offs[lo]) * e->expand) + (64u << 10);
        base[k + 1] = base[k] + ((cap + 255) & ~uint64_t(255));
    }
    CU(e->d_in.ensure(in_bytes + 64, sc));
    CU(e->d_in_offs.ensure((n + 1) * 8, sc));
    CU(e->d_out.ensure(base[K] + 64, sc));
    CU(e->d_out_offs.ensure(n * 8 + 8, sc));
    CU(e->d_out_lens.ensure(n * 4 + 4, sc));
    CU(e->d_status.ensure(n * 4 + 4, sc));
    CU(e->d_aux.ensure(n * 4 + 4, sc));
    CU(e->d_info.ensure(K * sizeof(ie_batch_info), sc));
    CU(e->h_info.ensure(K * sizeof(ie_batch_info)));
    CU(e->h_out.ensure(base[K] + 1));
    CU(e->h_out_offs.ensure(n * 8 + 8));
    CU(e->h_out_lens.ensure(n * 4 + 4));
    CU(e->h_status.ensure(n * 4 + 4));
    CU(e->h_aux.ensure(n * 4 + 4));
    {   // size the workspace once for the largest chunk so that no launch reallocates while others are in flight
        uint32_t max_exp, tcap;
        resolve_limits(limits, &max_exp, &tcap);
        IeWorkspace ws;
        ie_status_t st = prepare_workspace(e, kPipeChunk, tcap, true, &ws, 0, true);
        if (st != IE_OK) return st;
    }
    const uint64_t groups_x16 = sample_groups_x16(tmpl + base0, in_bytes, n);
    ie_batch_info* hinfo = (ie_batch_info*)e->h_info.p;
    ie_batch_info* hinfo_dev = nullptr;  // the same pinned block as the device sees it
    CU(cudaHostGetDevicePointer((void**)&hinfo_dev, hinfo, 0));
    CU(cudaMemcpyAsync(e->d_in_offs.p, tmpl_offs, (n + 1) * 8, cudaMemcpyHostToDevice, e->s_in));
    CU(cudaEventRecord(e->ev0, sc));
    for (uint64_t k = 0; k < K; ++k) {
        const uint64_t lo = cut[k], hi = cut[k + 1];
        const uint64_t b0 = tmpl_offs[lo], b1 = tmpl_offs[hi];
        if (b1 > b0) CU(cudaMemcpyAsync((uint8_t*)e->d_in.p + (b0 - base0), tmpl + b0, b1 - b0, cudaMemcpyHostToDevice, e->s_in));
        CU(cudaEventRecord(e->ev_in[k], e->s_in));
        CU(cudaStreamWaitEvent(sc, e->ev_in[k], 0));
        ie_status_t st = resolve_device(e, t, (const uint8_t*)e->d_in.p - base0, (const uint64_t*)e->d_in_offs.p + lo, hi - lo, limits,
                                        (uint8_t*)e->d_out.p + base[k], base[k + 1] - base[k], (uint64_t*)e->d_out_offs.p + lo,
                                        (uint32_t*)e->d_out_lens.p + lo, (int32_t*)e->d_status.p + lo, (uint32_t*)e->d_aux.p + lo,
                                        (ie_batch_info*)e->d_info.p + k, base[k], sc, in_bytes / n, host_rounds(limits), groups_x16);
        if (st != IE_OK) return st;
        ie_publish_info_kernel<<<1, 1, 0, sc>>>((const ie_batch_info*)e->d_info.p + k, hinfo_dev + k);
        CU(cudaGetLastError());
        CU(cudaEventRecord(e->ev_done[k], sc));
    }
    CU(cudaEventRecord(e->ev1, sc));
    bool overflow = false;
    uint64_t n_general = 0, n_limit = 0;
    double worst = 0.0;
    for (uint64_t k = 0; k < K; ++k) {
        const uint64_t lo = cut[k], hi = cut[k + 1];
        CU(cudaEventSynchronize(e->ev_done[k]));
        const uint64_t ob = hinfo[k].out_bytes;
        const uint64_t ib = tmpl_offs[hi] - tmpl_offs[lo];
        if (ib) worst = std::max(worst, (double)ob / (double)ib);
        n_general += hinfo[k].n_general;
        n_limit += hinfo[k].n_limit;
        if (ob > base[k + 1] - base[k]) { overflow = true; continue; }
        if (overflow) continue;
        if (ob) CU(cudaMemcpyAsync((uint8_t*)e->h_out.p + base[k], (uint8_t*)e->d_out.p + base[k], ob, cudaMemcpyDeviceToHost, e->s_out));
    }
    if (!overflow) {
        // The per-template arrays go over in four copies for the whole batch, behind the arenas: the device-to-host
        // engine is the long pole of the batch, and 64 small copies (16 chunks x 4 arrays) cost it about 1 ms of gaps.
        CU(cudaMemcpyAsync(e->h_out_offs.p, e->d_out_offs.p, n * 8, cudaMemcpyDeviceToHost, e->s_out));
        CU(cudaMemcpyAsync(e->h_out_lens.p, e->d_out_lens.p, n * 4, cudaMemcpyDeviceToHost, e->s_out));
        CU(cudaMemcpyAsync(e->h_status.p, e->d_status.p, n * 4, cudaMemcpyDeviceToHost, e->s_out));
        CU(cudaMemcpyAsync(e->h_aux.p, e->d_aux.p, n * 4, cudaMemcpyDeviceToHost, e->s_out));
    }
    CU(cudaStreamSynchronize(e->s_out));
    CU(cudaStreamSynchronize(sc));
    if (worst * 1.1 > e->expand) e->expand = worst * 1.25;
    if (overflow) return IE_OK;
    if (n_limit && !(limits && limits->max_expansions && limits->max_result_bytes)) return IE_OK;  // a default bound was hit: the unpipelined route escalates
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, e->ev0, e->ev1));
    res->out = (const uint8_t*)e->h_out.p;
    res->out_offs = (const uint64_t*)e->h_out_offs.p;
    res->out_lens = (const uint32_t*)e->h_out_lens.p;
    res->status = (const int32_t*)e->h_status.p;
    res->aux = (const uint32_t*)e->h_aux.p;
    res->info.n = n;
    res->info.out_bytes = base[K];
    res->info.n_general = n_general;
    res->info.n_limit = n_limit;
    res->info.kernel_ms = ms;  // first launch to last kernel end, copies overlapped
    *done = true;
    return IE_OK;
}

// Small batches (the reference's interactive shape: one task, a handful of strings): one copy in, the kernels, one copy
// out, ONE host synchronisation.  Inputs travel as [offsets | text] in one pinned block; results come back as
// [info | out_offs | out_lens | status | aux | first bytes of the out arena] in another.  *done = false: the results did
// not fit the block's arena and the caller takes the general route.
static constexpr uint64_t kSmallTemplates = 4096, kSmallBytes = 128u << 10, kSmallArena = 512u << 10;
static ie_status_t resolve_small(ie_engine* e, const ie_table* t, const uint8_t* tmpl, const uint64_t* tmpl_offs, uint64_t n,
                                 const ie_limits* limits, ie_result* res, bool* done) {
    *done = false;
    cudaStream_t s = e->stream;
    const uint64_t base0 = tmpl_offs[0], in_bytes = tmpl_offs[n] - base0, nr = n * t->n_states;
    const size_t offs_bytes = (n + 1) * 8, in_total = offs_bytes + in_bytes;
    const size_t a_info = 0, a_offs = 64, a_lens = a_offs + nr * 8, a_stat = a_lens + nr * 4, a_aux = a_stat + nr * 4;
    const size_t a_arena = (a_aux + nr * 4 + 15) & ~size_t(15), res_total = a_arena + kSmallArena;
    CU(e->d_small_in.ensure(in_total + 64, s));
    CU(e->d_small_res.ensure(res_total + 64, s));
    CU(e->h_small_in.ensure(in_total + 64));
    CU(e->h_small_res.ensure(res_total + 64));
    uint8_t* hin = (uint8_t*)e->h_small_in.p;
    std::memcpy(hin, tmpl_offs, offs_bytes);
    if (in_bytes) std::memcpy(hin + offs_bytes, tmpl + base0, in_bytes);
    CU(cudaMemcpyAsync(e->d_small_in.p, hin, in_total, cudaMemcpyHostToDevice, s));
    uint8_t* dres = (uint8_t*)e->d_small_res.p;
    CU(cudaEventRecord(e->ev0, s));
    ie_status_t st = resolve_device(e, t, (const uint8_t*)e->d_small_in.p + offs_bytes - base0, (const uint64_t*)e->d_small_in.p, n, limits,
                                    dres + a_arena, kSmallArena, (uint64_t*)(dres + a_offs), (uint32_t*)(dres + a_lens), (int32_t*)(dres + a_stat),
                                    (uint32_t*)(dres + a_aux), (ie_batch_info*)(dres + a_info), 0, s, n ? in_bytes / n : 0, host_rounds(limits),
                                    sample_groups_x16(tmpl + base0, in_bytes, n));
    if (st != IE_OK) return st;
    CU(cudaEventRecord(e->ev1, s));
    // results + the first part of the arena in one copy; the rest of the arena only if the batch produced more
    const size_t first = a_arena + std::min<size_t>(kSmallArena, std::max<size_t>(4 * in_bytes + 4096, 16384));
    uint8_t* hres = (uint8_t*)e->h_small_res.p;
    CU(cudaMemcpyAsync(hres, dres, first, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    const ie_batch_info* hi = (const ie_batch_info*)(hres + a_info);
    if (hi->out_bytes > kSmallArena) return IE_OK;  // too much output for this route
    if (hi->n_limit && !(limits && limits->max_expansions && limits->max_result_bytes)) return IE_OK;  // a default bound was hit: the general route escalates
    if (a_arena + hi->out_bytes > first) {
        CU(cudaMemcpyAsync(hres + first, dres + first, a_arena + hi->out_bytes - first, cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
    }
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, e->ev0, e->ev1));
    res->out = hres + a_arena;
    res->out_offs = (const uint64_t*)(hres + a_offs);
    res->out_lens = (const uint32_t*)(hres + a_lens);
    res->status = (const int32_t*)(hres + a_stat);
    res->aux = (const uint32_t*)(hres + a_aux);
    res->info = *hi;
    res->info.n = nr;
    res->info.kernel_ms = ms;
    *done = true;
    return IE_OK;
}

// Templates that stopped at a DEFAULT bound are run again on the full-size general tier with 8x the bounds, level by
// level, up to the hard caps of the header: "limit" then means "the reference would not finish within the caps", not
// "deeper than this engine's first guess".  Works on the engine-buffer layout of the host routes: templates in e->d_in /
// e->d_in_offs, per-template results in e->d_out_offs ... e->d_aux and their pinned mirrors; escalated result bytes are
// appended to the host arena at *host_bytes.  A bound the caller fixed is never raised.
static ie_status_t escalate_limits(ie_engine* e, const ie_table* t, uint64_t n, uint64_t nr, const ie_limits* limits, uint64_t* host_bytes,
                                   uint64_t* n_limit_left) {
    cudaStream_t s = e->stream;
    const bool exp_fixed = limits && limits->max_expansions, rb_fixed = limits && limits->max_result_bytes;
    uint32_t max_exp, tcap;
    resolve_limits(limits, &max_exp, &tcap);
    int32_t* h_status = (int32_t*)e->h_status.p;
    std::vector<uint32_t> list;
    for (uint64_t r = 0; r < nr; ++r)
        if ((h_status[r] & 0xFF) == IE_RES_LIMIT) list.push_back((uint32_t)r);
    while (!list.empty()) {
        const uint32_t exp2 = exp_fixed ? max_exp : (uint32_t)std::min<uint64_t>((uint64_t)max_exp * 8, IE_HARD_MAX_EXPANSIONS);
        const uint32_t tcap2 = rb_fixed ? tcap : (uint32_t)std::min<uint64_t>((uint64_t)tcap * 8, IE_HARD_MAX_RESULT_BYTES);
        if (exp2 == max_exp && tcap2 == tcap) break;  // nothing left to raise
        max_exp = exp2; tcap = tcap2;
        const uint64_t count = list.size();
        // few workers with a large scratch each: at most 1 GiB of scratch in total
        const size_t per_worker = ie_general_worker_bytes(tcap);
        const uint32_t workers = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>({count, (uint64_t)IE_GENERAL_WORKERS, (1ull << 30) / per_worker}));
        IeWorkspace ws{};
        CU(e->d_esc.ensure((size_t)workers * per_worker + 256, s));
        ws.scratch = (uint8_t*)e->d_esc.p;
        ws.general_workers = workers;
        static_assert(sizeof(ie_batch_info) <= 112, "d_misc layout: [count, overflow | info at 16 | list at 128]");
        CU(e->d_misc.ensure(128 + count * 4, s));
        uint32_t* d_count = (uint32_t*)e->d_misc.p;
        ie_batch_info* d_info2 = (ie_batch_info*)((uint8_t*)e->d_misc.p + 16);
        uint32_t* d_list = (uint32_t*)((uint8_t*)e->d_misc.p + 128);
        ws.overflow = d_count + 1;
        const uint32_t head[2] = {(uint32_t)count, 0u};
        CU(cudaMemcpyAsync(d_count, head, sizeof head, cudaMemcpyHostToDevice, s));
        CU(cudaMemcpyAsync(d_list, list.data(), count * 4, cudaMemcpyHostToDevice, s));
        // the escalated results go to the start of the engine's out arena (its content is on the host already)
        uint64_t cap = e->d_out.cap;
        ie_batch_info info2{};
        for (int attempt = 0;; ++attempt) {
            CU(cudaMemsetAsync(d_info2, 0, sizeof(ie_batch_info), s));
            CU(ie_launch_general_escalate(t->d_views, (const uint8_t*)e->d_in.p - e->in_bias, (const uint64_t*)e->d_in_offs.p, n, (uint8_t*)e->d_out.p, cap,
                                          (uint64_t*)e->d_out_offs.p, (uint32_t*)e->d_out_lens.p, (int32_t*)e->d_status.p, (uint32_t*)e->d_aux.p, ws,
                                          d_info2, max_exp, tcap, *host_bytes, d_list, d_count, s));
            CU(cudaMemcpyAsync(&info2, d_info2, sizeof info2, cudaMemcpyDeviceToHost, s));
            CU(cudaStreamSynchronize(s));
            if (info2.out_bytes <= cap) break;
            if (attempt == 1) return fail(IE_E_OVERFLOW, "ie_resolve_batch: output arena overflow while escalating limits");
            CU(e->d_out.ensure(info2.out_bytes, s));
            cap = e->d_out.cap;
        }
        // append the bytes, refresh the per-template entries of the escalated templates
        CU(e->h_out.grow_keep(*host_bytes + info2.out_bytes + 1, *host_bytes));
        if (info2.out_bytes) CU(cudaMemcpyAsync((uint8_t*)e->h_out.p + *host_bytes, e->d_out.p, info2.out_bytes, cudaMemcpyDeviceToHost, s));
        if (count <= 64) {
            for (uint32_t r : list) {
                CU(cudaMemcpyAsync((uint64_t*)e->h_out_offs.p + r, (uint64_t*)e->d_out_offs.p + r, 8, cudaMemcpyDeviceToHost, s));
                CU(cudaMemcpyAsync((uint32_t*)e->h_out_lens.p + r, (uint32_t*)e->d_out_lens.p + r, 4, cudaMemcpyDeviceToHost, s));
                CU(cudaMemcpyAsync((int32_t*)e->h_status.p + r, (int32_t*)e->d_status.p + r, 4, cudaMemcpyDeviceToHost, s));
                CU(cudaMemcpyAsync((uint32_t*)e->h_aux.p + r, (uint32_t*)e->d_aux.p + r, 4, cudaMemcpyDeviceToHost, s));
            }
        } else {
            CU(cudaMemcpyAsync(e->h_out_offs.p, e->d_out_offs.p, nr * 8, cudaMemcpyDeviceToHost, s));
            CU(cudaMemcpyAsync(e->h_out_lens.p, e->d_out_lens.p, nr * 4, cudaMemcpyDeviceToHost, s));
            CU(cudaMemcpyAsync(e->h_status.p, e->d_status.p, nr * 4, cudaMemcpyDeviceToHost, s));
            CU(cudaMemcpyAsync(e->h_aux.p, e->d_aux.p, nr * 4, cudaMemcpyDeviceToHost, s));
        }
        CU(cudaStreamSynchronize(s));
        *host_bytes += (info2.out_bytes + 15) & ~uint64_t(15);
        h_status = (int32_t*)e->h_status.p;
        std::vector<uint32_t> next;
        for (uint32_t r : list)
            if ((h_status[r] & 0xFF) == IE_RES_LIMIT) next.push_back(r);
        list.swap(next);
    }
    *n_limit_left = list.size();
    return IE_OK;
}

ie_status_t ie_resolve_batch(ie_engine* e, const ie_table* t, const uint8_t* tmpl, const uint64_t* tmpl_offs, uint64_t n,
                             const ie_limits* limits, ie_result* res) {
    if (!e || !t || !res || (n && (!tmpl_offs))) return fail(IE_E_INVALID, "ie_resolve_batch: NULL argument");
    std::memset(res, 0, sizeof *res);
    CU(cudaSetDevice(e->device));
    const uint64_t base0 = n ? tmpl_offs[0] : 0;  // not 0 when the batch is a shard of a larger arena
    if (n && tmpl_offs[n] < base0) return fail(IE_E_INVALID, "ie_resolve_batch: offsets not monotone");
    const uint64_t in_bytes = n ? tmpl_offs[n] - base0 : 0;
    if (in_bytes && !tmpl) return fail(IE_E_INVALID, "ie_resolve_batch: NULL template arena");
    cudaStream_t s = e->stream;
    const uint64_t S = t->n_states, nr = n * S;  // every snapshot resolves all n templates: nr results, index = snapshot * n + template
    if (n && nr <= kSmallTemplates && in_bytes <= kSmallBytes) {
        bool done = false;
        ie_status_t st = resolve_small(e, t, tmpl, tmpl_offs, n, limits, res, &done);
        if (st != IE_OK || done) return st;
    }
    if (S == 1 && n >= (1u << 16)) {
        bool done = false;
        ie_status_t st = resolve_pipelined(e, t, tmpl, tmpl_offs, n, limits, res, &done);
        if (st != IE_OK || done) return st;
    }
    CU(e->d_in.ensure(in_bytes + 16, s));
    CU(e->d_in_offs.ensure((n + 1) * 8, s));
    CU(e->d_out_offs.ensure(nr * 8 + 8, s));
    CU(e->d_out_lens.ensure(nr * 4 + 4, s));
    CU(e->d_status.ensure(nr * 4 + 4, s));
    CU(e->d_aux.ensure(nr * 4 + 4, s));
    CU(e->d_info.ensure(sizeof(ie_batch_info), s));
    CU(e->h_info.ensure(sizeof(ie_batch_info)));
    CU(e->d_out.ensure(std::max<uint64_t>(in_bytes * 2 * S + (1u << 16), 1u << 20), s));
    if (in_bytes) CU(cudaMemcpyAsync(e->d_in.p, tmpl + base0, in_bytes, cudaMemcpyHostToDevice, s));
    if (n) CU(cudaMemcpyAsync(e->d_in_offs.p, tmpl_offs, (n + 1) * 8, cudaMemcpyHostToDevice, s));
    e->in_bias = base0;
    ie_batch_info* hinfo = (ie_batch_info*)e->h_info.p;
    float kernel_ms = 0.f;
    // An overflowing run skips the stages behind the one that overflowed (a rescan round whose gather did not fit
    // reserves nothing for the rounds after it), so the size it reports is a lower bound: regrow to at least twice the
    // arena, rerun, and repeat until a run fits - every attempt gets at least one stage further.
    const int max_attempts = 8;
    for (int attempt = 0;; ++attempt) {
        CU(cudaEventRecord(e->ev0, s));
        ie_status_t st = resolve_device(e, t, (const uint8_t*)e->d_in.p - base0, (const uint64_t*)e->d_in_offs.p, n, limits,
                                        (uint8_t*)e->d_out.p, e->d_out.cap, (uint64_t*)e->d_out_offs.p, (uint32_t*)e->d_out_lens.p,
                                        (int32_t*)e->d_status.p, (uint32_t*)e->d_aux.p, (ie_batch_info*)e->d_info.p, 0, s,
                                        n ? in_bytes / n : 0, host_rounds(limits), sample_groups_x16(tmpl + base0, in_bytes, n));
        if (st != IE_OK) return st;
        CU(cudaEventRecord(e->ev1, s));
        CU(cudaMemcpyAsync(hinfo, e->d_info.p, sizeof(ie_batch_info), cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
        CU(cudaEventElapsedTime(&kernel_ms, e->ev0, e->ev1));
        if (hinfo->out_bytes <= e->d_out.cap) break;
        if (attempt + 1 == max_attempts) return fail(IE_E_OVERFLOW, "ie_resolve_batch: output arena overflow after regrowing it " + std::to_string(max_attempts - 1) + " times");
        CU(e->d_out.ensure(std::max<uint64_t>(hinfo->out_bytes, 2 * (uint64_t)e->d_out.cap), s));
    }
    const uint64_t ob = hinfo->out_bytes;
    CU(e->h_out.ensure(ob + 1));
    CU(e->h_out_offs.ensure(nr * 8 + 8));
    CU(e->h_out_lens.ensure(nr * 4 + 4));
    CU(e->h_status.ensure(nr * 4 + 4));
    CU(e->h_aux.ensure(nr * 4 + 4));
    if (ob) CU(cudaMemcpyAsync(e->h_out.p, e->d_out.p, ob, cudaMemcpyDeviceToHost, s));
    if (n) {
        CU(cudaMemcpyAsync(e->h_out_offs.p, e->d_out_offs.p, nr * 8, cudaMemcpyDeviceToHost, s));
        CU(cudaMemcpyAsync(e->h_out_lens.p, e->d_out_lens.p, nr * 4, cudaMemcpyDeviceToHost, s));
        CU(cudaMemcpyAsync(e->h_status.p, e->d_status.p, nr * 4, cudaMemcpyDeviceToHost, s));
        CU(cudaMemcpyAsync(e->h_aux.p, e->d_aux.p, nr * 4, cudaMemcpyDeviceToHost, s));
    }
    CU(cudaStreamSynchronize(s));
    ie_batch_info info = *hinfo;
    if (info.n_limit) {  // templates that stopped at a default bound: larger bounds, up to the hard caps
        uint64_t host_bytes = (ob + 15) & ~uint64_t(15), left = 0;
        ie_status_t st = escalate_limits(e, t, n, nr, limits, &host_bytes, &left);
        if (st != IE_OK) return st;
        info.out_bytes = host_bytes;
        info.n_limit = left;
    }
    res->out = (const uint8_t*)e->h_out.p;
    res->out_offs = (const uint64_t*)e->h_out_offs.p;
    res->out_lens = (const uint32_t*)e->h_out_lens.p;
    res->status = (const int32_t*)e->h_status.p;
    res->aux = (const uint32_t*)e->h_aux.p;
    res->info = info;
    res->info.n = nr;
    res->info.kernel_ms = kernel_ms;
    return IE_OK;
}

// One host thread per engine: contiguous shards of ONE host batch, each through its own engine's pipelined path.
ie_status_t ie_resolve_batch_multi(ie_engine* const* engines, const ie_table* const* tables, uint32_t n_engines, const uint8_t* tmpl,
                                   const uint64_t* tmpl_offs, uint64_t n, const ie_limits* limits, ie_shard_result* shards) {
    if (!engines || !tables || !n_engines || !shards || (n && !tmpl_offs)) return fail(IE_E_INVALID, "ie_resolve_batch_multi: NULL argument");
    for (uint32_t g = 0; g < n_engines; ++g) {
        if (!engines[g] || !tables[g]) return fail(IE_E_INVALID, "ie_resolve_batch_multi: NULL engine or table");
        if (tables[g]->e != engines[g]) return fail(IE_E_INVALID, "ie_resolve_batch_multi: table " + std::to_string(g) + " belongs to another engine");
        if (tables[g]->n_states != 1) return fail(IE_E_INVALID, "ie_resolve_batch_multi: one snapshot per table");
        for (uint32_t h = 0; h < g; ++h)
            if (engines[h] == engines[g]) return fail(IE_E_INVALID, "ie_resolve_batch_multi: the same engine twice");
    }
    const uint64_t per = (n + n_engines - 1) / n_engines;  // ceil(n / G) templates per shard, the last ones may be short or empty
    auto work = [&](uint32_t g) {
        ie_shard_result& sh = shards[g];
        std::memset(&sh, 0, sizeof sh);
        sh.first = std::min<uint64_t>(n, (uint64_t)g * per);
        sh.n = std::min<uint64_t>(n, (uint64_t)(g + 1) * per) - sh.first;
        sh.status = ie_resolve_batch(engines[g], tables[g], tmpl, tmpl_offs ? tmpl_offs + sh.first : nullptr, sh.n, limits, &sh.res);
        if (sh.status != IE_OK) std::snprintf(sh.error, sizeof sh.error, "%s", ie_last_error());
    };
    std::vector<std::thread> th;
    for (uint32_t g = 1; g < n_engines; ++g) th.emplace_back(work, g);
    work(0);
    for (auto& x : th) x.join();
    for (uint32_t g = 0; g < n_engines; ++g)
        if (shards[g].status != IE_OK) return fail(shards[g].status, "ie_resolve_batch_multi: shard " + std::to_string(g) + ": " + shards[g].error);
    return IE_OK;
}

// The host gather of SURVEY.md §8(e): the shards' results concatenated in template order into the caller's arrays.
ie_status_t ie_shards_gather(const ie_shard_result* shards, uint32_t n_shards, uint8_t* out, uint64_t out_capacity, uint64_t* out_offs,
                             int32_t* status, uint32_t* aux, uint64_t* out_bytes) {
    if (!shards || !n_shards || !out_offs || !status || !aux) return fail(IE_E_INVALID, "ie_shards_gather: NULL argument");
    // shard bases by a prefix over the shards' exact result bytes, then every shard is copied by its own thread
    std::vector<uint64_t> base(n_shards + 1, 0);
    std::vector<uint64_t> bytes(n_shards, 0);
    auto count = [&](uint32_t g) {
        uint64_t b = 0;
        for (uint64_t i = 0; i < shards[g].n; ++i) b += shards[g].res.out_lens[i];
        bytes[g] = b;
    };
    {
        std::vector<std::thread> th;
        for (uint32_t g = 1; g < n_shards; ++g) th.emplace_back(count, g);
        count(0);
        for (auto& x : th) x.join();
    }
    for (uint32_t g = 0; g < n_shards; ++g) base[g + 1] = base[g] + bytes[g];
    if (out_bytes) *out_bytes = base[n_shards];
    if (base[n_shards] > out_capacity || (base[n_shards] && !out)) return fail(IE_E_OVERFLOW, "ie_shards_gather: output arena too small (see out_bytes)");
    auto copy = [&](uint32_t g) {
        const ie_shard_result& sh = shards[g];
        uint64_t at = base[g];
        // results of one tile lie back to back in template order in the shard's arena: copy whole runs, not strings
        for (uint64_t i = 0; i < sh.n;) {
            const uint64_t src0 = sh.res.out_offs[i];
            uint64_t run = 0, j = i;
            for (; j < sh.n && sh.res.out_offs[j] == src0 + run; ++j) {
                out_offs[sh.first + j] = at + run;
                run += sh.res.out_lens[j];
            }
            if (run) std::memcpy(out + at, sh.res.out + src0, run);
            at += run;
            i = j;
        }
        std::memcpy(status + sh.first, sh.res.status, sh.n * sizeof(int32_t));
        std::memcpy(aux + sh.first, sh.res.aux, sh.n * sizeof(uint32_t));
    };
    std::vector<std::thread> th;
    for (uint32_t g = 1; g < n_shards; ++g) th.emplace_back(copy, g);
    copy(0);
    for (auto& x : th) x.join();
    const ie_shard_result& last = shards[n_shards - 1];
    out_offs[last.first + last.n] = base[n_shards];
    return IE_OK;
}

ie_status_t ie_lookup_batch(ie_engine* e, const ie_table* t, const uint8_t* keys, const uint64_t* key_offs, uint64_t n,
                            int32_t* tag_out, uint32_t* entry_out) {
    if (!e || !t || (n && (!key_offs || !tag_out || !entry_out))) return fail(IE_E_INVALID, "ie_lookup_batch: NULL argument");
    if (!n) return IE_OK;
    CU(cudaSetDevice(e->device));
    cudaStream_t s = e->stream;
    const uint64_t bytes = key_offs[n];
    CU(e->d_in.ensure(bytes + 16, s));
    CU(e->d_in_offs.ensure((n + 1) * 8, s));
    CU(e->d_status.ensure(n * 4, s));
    CU(e->d_aux.ensure(n * 4, s));
    if (bytes) CU(cudaMemcpyAsync(e->d_in.p, keys, bytes, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(e->d_in_offs.p, key_offs, (n + 1) * 8, cudaMemcpyHostToDevice, s));
    CU(ie_launch_lookup(t->view, (const uint8_t*)e->d_in.p, (const uint64_t*)e->d_in_offs.p, n, (int32_t*)e->d_status.p,
                        (uint32_t*)e->d_aux.p, s));
    CU(cudaMemcpyAsync(tag_out, e->d_status.p, n * 4, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(entry_out, e->d_aux.p, n * 4, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    return IE_OK;
}

ie_status_t ie_escape_batch_device(ie_engine* e, int mode, const uint8_t* d_in, const uint64_t* d_in_offs, uint64_t n, uint64_t in_bytes,
                                   uint8_t* d_out, uint64_t out_capacity, uint64_t* d_out_offs, void* stream) {
    if (!e || !d_out_offs || (n && !d_in_offs) || (mode != 0 && mode != 1)) return fail(IE_E_INVALID, "ie_escape_batch_device: bad argument");
    CU(cudaSetDevice(e->device));
    IeWorkspace ws;
    ie_status_t st = prepare_workspace(e, n, 0, false, &ws, ie_escape_tiles(in_bytes));
    if (st != IE_OK) return st;
    CU(e->ws_list.ensure(((size_t)IE_ESCAPE_FIX_CAP + ie_escape_tiles(in_bytes) + 1) * sizeof(uint64_t), e->stream));
    ws.fix_list = (uint64_t*)e->ws_list.p;
    ws.tile_first = ws.fix_list + IE_ESCAPE_FIX_CAP;
    cudaStream_t s = stream ? (cudaStream_t)stream : e->stream;
    CU(ie_launch_escape(mode, d_in, d_in_offs, n, in_bytes, d_out, out_capacity, d_out_offs, ws, s));
    return IE_OK;
}

ie_status_t ie_escape_batch(ie_engine* e, int mode, const uint8_t* in, const uint64_t* in_offs, uint64_t n, const uint8_t** out,
                            const uint64_t** out_offs) {
    if (!e || !out || !out_offs || (n && !in_offs)) return fail(IE_E_INVALID, "ie_escape_batch: NULL argument");
    CU(cudaSetDevice(e->device));
    cudaStream_t s = e->stream;
    const uint64_t in_bytes = n ? in_offs[n] : 0;
    const uint64_t cap = mode == 1 ? in_bytes * 2 : in_bytes;  // escape at most doubles, unescape never grows
    CU(e->d_in.ensure(in_bytes + 16, s));
    CU(e->d_in_offs.ensure((n + 1) * 8, s));
    CU(e->d_out.ensure(cap + 16, s));
    CU(e->d_out_offs.ensure((n + 1) * 8, s));
    CU(e->h_out.ensure(cap + 1));
    CU(e->h_out_offs.ensure((n + 1) * 8));
    if (in_bytes) CU(cudaMemcpyAsync(e->d_in.p, in, in_bytes, cudaMemcpyHostToDevice, s));
    if (n) CU(cudaMemcpyAsync(e->d_in_offs.p, in_offs, (n + 1) * 8, cudaMemcpyHostToDevice, s));
    ie_status_t st = ie_escape_batch_device(e, mode, (const uint8_t*)e->d_in.p, (const uint64_t*)e->d_in_offs.p, n, in_bytes,
                                            (uint8_t*)e->d_out.p, cap, (uint64_t*)e->d_out_offs.p, s);
    if (st != IE_OK) return st;
    CU(cudaMemcpyAsync(e->h_out_offs.p, e->d_out_offs.p, (n + 1) * 8, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    const uint64_t ob = ((const uint64_t*)e->h_out_offs.p)[n];
    if (ob) CU(cudaMemcpyAsync(e->h_out.p, e->d_out.p, ob, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    *out = (const uint8_t*)e->h_out.p;
    *out_offs = (const uint64_t*)e->h_out_offs.p;
    return IE_OK;
}

static ie_status_t pack_patterns(const uint8_t* pats, const uint64_t* pat_offs, uint32_t n_pat, int invert, IeGlobPatterns* gp) {
    if (n_pat > IE_MAX_PATTERNS) return fail(IE_E_INVALID, "ie_glob_sweep: more than IE_MAX_PATTERNS patterns");
    if (n_pat && (!pat_offs || pat_offs[n_pat] > sizeof gp->bytes)) return fail(IE_E_INVALID, "ie_glob_sweep: patterns exceed 3584 bytes");
    std::memset(gp, 0, sizeof *gp);
    gp->n_pat = n_pat;
    gp->invert = invert ? 1u : 0u;
    for (uint32_t i = 0; i <= n_pat && n_pat; ++i) gp->off[i] = (uint16_t)pat_offs[i];
    if (n_pat && pat_offs[n_pat]) std::memcpy(gp->bytes, pats, pat_offs[n_pat]);
    ie_glob_compile(gp);
    return IE_OK;
}

ie_status_t ie_glob_sweep_device(ie_engine* e, const uint8_t* d_keys, const uint64_t* d_key_offs, uint64_t n, const uint8_t* pats,
                                 const uint64_t* pat_offs, uint32_t n_pat, int invert, uint32_t* d_mask, uint64_t* d_n_deleted,
                                 void* stream) {
    if (!e || !d_n_deleted || (n && (!d_key_offs || !d_mask))) return fail(IE_E_INVALID, "ie_glob_sweep_device: NULL argument");
    IeGlobPatterns gp;
    ie_status_t st = pack_patterns(pats, pat_offs, n_pat, invert, &gp);
    if (st != IE_OK) return st;
    CU(cudaSetDevice(e->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : e->stream;
    CU(ie_launch_glob(d_keys, d_key_offs, n, gp, d_mask, d_n_deleted, nullptr, s));
    return IE_OK;
}

static const uint64_t kFirstMatchFewKeys = 256;  // up to here ie_glob_first_match runs one CTA per text

ie_status_t ie_glob_first_match(ie_engine* e, const uint8_t* keys, const uint64_t* key_offs, uint64_t n, const uint8_t* pats,
                                const uint64_t* pat_offs, uint32_t n_pat, uint32_t* first) {
    if (!e || (n && (!key_offs || !first))) return fail(IE_E_INVALID, "ie_glob_first_match: NULL argument");
    if (n && n <= kFirstMatchFewKeys) {
        // A handful of texts (replace_map / goto_map test ONE): a CTA per text instead of a thread per key, patterns of
        // any length and number from global memory.
        if (n_pat && !pat_offs) return fail(IE_E_INVALID, "ie_glob_first_match: NULL argument");
        CU(cudaSetDevice(e->device));
        cudaStream_t s = e->stream;
        const uint64_t bytes = key_offs[n], pbytes = n_pat ? pat_offs[n_pat] : 0;
        const uint64_t zero = 0;
        CU(e->d_in.ensure(bytes + 16, s));
        CU(e->d_in_offs.ensure((n + 1) * 8, s));
        CU(e->d_mask.ensure(pbytes + 16, s));
        CU(e->d_misc.ensure(((uint64_t)n_pat + 1) * 8, s));
        CU(e->d_aux.ensure(n * 4 + 4, s));
        if (bytes) CU(cudaMemcpyAsync(e->d_in.p, keys, bytes, cudaMemcpyHostToDevice, s));
        CU(cudaMemcpyAsync(e->d_in_offs.p, key_offs, (n + 1) * 8, cudaMemcpyHostToDevice, s));
        if (pbytes) CU(cudaMemcpyAsync(e->d_mask.p, pats, pbytes, cudaMemcpyHostToDevice, s));
        CU(cudaMemcpyAsync(e->d_misc.p, n_pat ? pat_offs : &zero, ((uint64_t)n_pat + 1) * 8, cudaMemcpyHostToDevice, s));
        CU(ie_launch_glob_first_long((const uint8_t*)e->d_in.p, (const uint64_t*)e->d_in_offs.p, n, (const uint8_t*)e->d_mask.p,
                                     (const uint64_t*)e->d_misc.p, n_pat, (uint32_t*)e->d_aux.p, s));
        CU(cudaMemcpyAsync(first, e->d_aux.p, n * 4, cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
        return IE_OK;
    }
    IeGlobPatterns gp;
    ie_status_t st = pack_patterns(pats, pat_offs, n_pat, 0, &gp);
    if (st != IE_OK) return st;
    CU(cudaSetDevice(e->device));
    cudaStream_t s = e->stream;
    const uint64_t bytes = n ? key_offs[n] : 0;
    CU(e->d_in.ensure(bytes + 16, s));
    CU(e->d_in_offs.ensure((n + 1) * 8, s));
    CU(e->d_mask.ensure((n + 31) / 32 * 4 + 4, s));
    CU(e->d_aux.ensure(n * 4 + 4, s));
    CU(e->d_misc.ensure(64, s));
    if (bytes) CU(cudaMemcpyAsync(e->d_in.p, keys, bytes, cudaMemcpyHostToDevice, s));
    if (n) CU(cudaMemcpyAsync(e->d_in_offs.p, key_offs, (n + 1) * 8, cudaMemcpyHostToDevice, s));
    CU(ie_launch_glob((const uint8_t*)e->d_in.p, (const uint64_t*)e->d_in_offs.p, n, gp, (uint32_t*)e->d_mask.p, (uint64_t*)e->d_misc.p,
                      (uint32_t*)e->d_aux.p, s));
    if (n) CU(cudaMemcpyAsync(first, e->d_aux.p, n * 4, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    return IE_OK;
}

ie_status_t ie_glob_sweep(ie_engine* e, const uint8_t* keys, const uint64_t* key_offs, uint64_t n, const uint8_t* pats,
                          const uint64_t* pat_offs, uint32_t n_pat, int invert, uint32_t* mask, uint64_t* n_deleted) {
    if (!e || (n && (!key_offs || !mask))) return fail(IE_E_INVALID, "ie_glob_sweep: NULL argument");
    CU(cudaSetDevice(e->device));
    cudaStream_t s = e->stream;
    const uint64_t bytes = n ? key_offs[n] : 0;
    const uint64_t words = (n + 31) / 32;
    CU(e->d_in.ensure(bytes + 16, s));
    CU(e->d_in_offs.ensure((n + 1) * 8, s));
    CU(e->d_mask.ensure(words * 4 + 4, s));
    CU(e->d_misc.ensure(64, s));
    if (bytes) CU(cudaMemcpyAsync(e->d_in.p, keys, bytes, cudaMemcpyHostToDevice, s));
    if (n) CU(cudaMemcpyAsync(e->d_in_offs.p, key_offs, (n + 1) * 8, cudaMemcpyHostToDevice, s));
    ie_status_t st = ie_glob_sweep_device(e, (const uint8_t*)e->d_in.p, (const uint64_t*)e->d_in_offs.p, n, pats, pat_offs, n_pat, invert,
                                          (uint32_t*)e->d_mask.p, (uint64_t*)e->d_misc.p, s);
    if (st != IE_OK) return st;
    uint64_t nd = 0;
    if (words) CU(cudaMemcpyAsync(mask, e->d_mask.p, words * 4, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(&nd, e->d_misc.p, 8, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    if (n_deleted) *n_deleted = nd;
    return IE_OK;
}

ie_status_t ie_device_alloc(ie_engine* e, uint64_t bytes, void** d_ptr) {
    if (!e || !d_ptr) return fail(IE_E_INVALID, "ie_device_alloc: NULL argument");
    CU(cudaSetDevice(e->device));
    CU(cudaMalloc(d_ptr, bytes ? bytes : 1));
    return IE_OK;
}
void ie_device_free(ie_engine* e, void* d_ptr) {
    if (!e || !d_ptr) return;
    cudaSetDevice(e->device);
    cudaFree(d_ptr);
}
ie_status_t ie_copy_to_device(ie_engine* e, void* d_dst, const void* h_src, uint64_t bytes) {
    if (!e) return fail(IE_E_INVALID, "ie_copy_to_device: engine is NULL");
    CU(cudaSetDevice(e->device));
    if (bytes) CU(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return IE_OK;
}
ie_status_t ie_copy_to_host(ie_engine* e, void* h_dst, const void* d_src, uint64_t bytes) {
    if (!e) return fail(IE_E_INVALID, "ie_copy_to_host: engine is NULL");
    CU(cudaSetDevice(e->device));
    if (bytes) CU(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return IE_OK;
}
ie_status_t ie_host_alloc(uint64_t bytes, void** h_ptr) {
    if (!h_ptr) return fail(IE_E_INVALID, "ie_host_alloc: NULL argument");
    CU(cudaHostAlloc(h_ptr, bytes ? bytes : 1, cudaHostAllocDefault));
    return IE_OK;
}
ie_status_t ie_host_alloc_wc(uint64_t bytes, void** h_ptr) {
    if (!h_ptr) return fail(IE_E_INVALID, "ie_host_alloc_wc: NULL argument");
    CU(cudaHostAlloc(h_ptr, bytes ? bytes : 1, cudaHostAllocWriteCombined));
    return IE_OK;
}
void ie_host_free(void* h_ptr) { if (h_ptr) cudaFreeHost(h_ptr); }

void ie_free(void* p) { std::free(p); }

ie_status_t ie_call_json(ie_engine* e, const char* args_json, size_t len, char** out_json, size_t* out_len) {
    if (!e || !args_json || !out_json) return fail(IE_E_INVALID, "ie_call_json: NULL argument");
    std::string out;
    std::string why;
    ie_status_t st = ie_host::call_json(e, std::string(args_json, len), &out, &why);
    if (st != IE_OK) return fail(st, why);
    char* p = (char*)std::malloc(out.size() + 1);
    if (!p) return fail(IE_E_NOMEM, "ie_call_json: out of host memory");
    std::memcpy(p, out.data(), out.size());
    p[out.size()] = 0;
    *out_json = p;
    if (out_len) *out_len = out.size();
    return IE_OK;
}

}  // extern "C"
