// Device half of the C-ABI: context (pinned staging / pinned row-group cache / host copy threads), upload, decode,
// and the thin extern "C" wrappers around the post-processing kernels.
#include <cuda_runtime.h>
#include <sched.h>

#include <atomic>
#include <cctype>
#include <cstdlib>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <functional>
#include <memory>
#include <stdexcept>
#include <algorithm>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/pst_b200.h"
#include "host_state.h"
#include "kernels.h"

using namespace pst;

namespace {

struct CudaFail : std::runtime_error {
    using std::runtime_error::runtime_error;
};
inline void ck(cudaError_t e, const char *what) {
    if (e != cudaSuccess) throw CudaFail(std::string(what) + ": " + cudaGetErrorString(e));
}

// CPUs that are local (same NUMA node / PCIe root) to a CUDA device, from sysfs; empty when unknown.  Pinned staging
// memory is allocated and filled from these CPUs so that the H2D DMA does not cross the socket interconnect.
static std::vector<int> local_cpus_of_device(int device) {
    std::vector<int> cpus;
    char bus[64] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) != cudaSuccess) return cpus;
    for (char *c = bus; *c; c++) *c = (char)tolower(*c);
    std::string path = std::string("/sys/bus/pci/devices/") + bus + "/local_cpulist";
    FILE *f = fopen(path.c_str(), "r");
    if (!f) return cpus;
    char line[4096] = {0};
    if (fgets(line, sizeof line, f)) {
        const char *p = line;
        while (*p) {
            char *e;
            long a = strtol(p, &e, 10);
            if (e == p) break;
            long b = a;
            if (*e == '-') { p = e + 1; b = strtol(p, &e, 10); }
            for (long c = a; c <= b && c < CPU_SETSIZE; c++) cpus.push_back((int)c);
            p = (*e == ',') ? e + 1 : e;
            if (*e != ',') break;
        }
    }
    fclose(f);
    return cpus;
}
static bool pin_current_thread(const std::vector<int> &cpus, cpu_set_t *old) {
    if (cpus.empty()) return false;
    if (old && sched_getaffinity(0, sizeof(cpu_set_t), old) != 0) return false;
    cpu_set_t set;
    CPU_ZERO(&set);
    for (int c : cpus) CPU_SET(c, &set);
    return sched_setaffinity(0, sizeof set, &set) == 0;
}

// Minimal fork-join pool for the host-side staging copies (mmap -> pinned).
class CopyPool {
public:
    explicit CopyPool(int n, std::vector<int> cpus = {}) : cpus_(std::move(cpus)) {
        for (int i = 0; i < n; i++) workers_.emplace_back([this] { pin_current_thread(cpus_, nullptr); run(); });
    }
    ~CopyPool() {
        {
            std::lock_guard<std::mutex> g(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto &t : workers_) t.join();
    }
    int size() const { return (int)workers_.size(); }
    // runs fn(i) for i in [0, n) on the pool and the calling thread; returns when all are done
    void parallel_for(int n, const std::function<void(int)> &fn) {
        if (n <= 0) return;
        if (workers_.empty() || n == 1) {
            for (int i = 0; i < n; i++) fn(i);
            return;
        }
        std::unique_lock<std::mutex> g(m_);
        fn_ = &fn;
        next_ = 0;
        total_ = n;
        pending_ = n;
        gen_++;
        g.unlock();
        cv_.notify_all();
        work();  // the caller helps
        g.lock();
        done_cv_.wait(g, [this] { return pending_ == 0; });
        fn_ = nullptr;
    }

private:
    void work() {
        for (;;) {
            int i;
            {
                std::lock_guard<std::mutex> g(m_);
                if (!fn_ || next_ >= total_) return;
                i = next_++;
            }
            (*fn_)(i);
            {
                std::lock_guard<std::mutex> g(m_);
                if (--pending_ == 0) done_cv_.notify_all();
            }
        }
    }
    void run() {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> g(m_);
                cv_.wait(g, [&] { return stop_ || gen_ != seen; });
                if (stop_) return;
                seen = gen_;
            }
            work();
        }
    }
    std::vector<int> cpus_;
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_, done_cv_;
    const std::function<void(int)> *fn_ = nullptr;
    int next_ = 0, total_ = 0, pending_ = 0;
    uint64_t gen_ = 0;
    bool stop_ = false;
};

struct PinnedBuf {
    uint8_t *ptr = nullptr;
    int64_t cap = 0;
    cudaEvent_t ev = nullptr;  // last H2D that read from this buffer
    bool busy = false;
};

}  // namespace

struct pst_ctx {
    int device = 0;
    int sm_count = 148;
    int64_t cache_budget = 0;
    int64_t cache_bytes = 0;
    std::unordered_map<uint64_t, PinnedBuf> cache;
    PinnedBuf ring[3];
    int ring_next = 0;
    std::unique_ptr<CopyPool> pool;
    std::vector<int> local_cpus;
    std::mutex mu;
    // counters
    std::atomic<int64_t> bytes_staged{0}, bytes_h2d{0}, cache_hits{0}, cache_misses{0}, pages_decoded{0},
        rowgroups_decoded{0}, kernels_launched{0};
};

#define PST_TRY try {
#define PST_CATCH(ret)                \
    }                                 \
    catch (const std::exception &e) { \
        pst::set_error(e.what());     \
        return ret;                   \
    }

static void fill_parallel(pst_ctx *c, const pst_plan *p, uint8_t *dst) {
    // split the pages into ~equal byte ranges, one task per range
    int64_t npages = (int64_t)p->pages.size();
    int tasks = c->pool ? c->pool->size() + 1 : 1;
    if (p->payload_bytes < (4 << 20)) tasks = 1;
    std::vector<int64_t> cuts(1, 0);
    int64_t per = p->payload_bytes / tasks + 1, acc = 0;
    for (int64_t i = 0; i < npages; i++) {
        acc += p->pages[i].d.comp_size;
        if (acc >= per && (int)cuts.size() < tasks) {
            cuts.push_back(i + 1);
            acc = 0;
        }
    }
    if (cuts.back() != npages) cuts.push_back(npages);
    int n = (int)cuts.size() - 1;
    std::function<void(int)> fn = [&](int t) { pst_plan_fill_raw(p, dst, cuts[t], cuts[t + 1]); };
    if (n <= 0) {
        pst_plan_fill_raw(p, dst, 0, 0);
        return;
    }
    if (c->pool) c->pool->parallel_for(n, fn);
    else for (int t = 0; t < n; t++) fn(t);
}

// A row-group's raw region goes up in ONE cudaMemcpyAsync: copies from different streams are served in submission
// order, so each takes ~5 ms.  (Chunking them was tried: chunks of two row-groups interleave, every copy then takes
// twice as long for the same PCIe throughput, and the decode behind it starts later.)
static void h2d_copy(uint64_t d_dst, const uint8_t *src, int64_t n, cudaStream_t s) {
    ck(cudaMemcpyAsync((void *)d_dst, src, (size_t)n, cudaMemcpyHostToDevice, s), "cudaMemcpyAsync H2D");
}

// Snappy: fragment index of the multi-fragment pages, every fragment of every page in parallel, serial fallback for
// flagged pages.  Returns the number of launches; `ev` (optional, 4 events) brackets the three launches.
static int launch_snappy_stage(pst_plan *p, uint8_t *arena, int32_t *status, cudaStream_t s, cudaEvent_t *ev) {
    const DevPage *pages = (const DevPage *)(arena + p->pages_off);
    const SnFrag *frags = (const SnFrag *)(arena + p->frag_list_off);
    const int32_t *multi = (const int32_t *)(arena + p->multi_list_off);
    uint32_t *frag_pos = (uint32_t *)(arena + p->frag_pos_off);
    uint32_t *page_flag = (uint32_t *)(arena + p->page_flag_off);
    const int n_frags = (int)p->snappy_frags.size(), n_multi = (int)p->multi_pages.size();
    int nl = 0;
    if (ev) ck(cudaEventRecord(ev[0], s), "record");
    if (!p->index_pages.empty()) {
        // Longest pages first.  PST_IDX_CLUSTER=1 (latency mode, read per call) gives every page of >= 256 KiB a cluster of
        // four SMs: measured on C2, 0.49 ms for the sixteen 1 MiB dictionary pages instead of 0.65 ms, but three times the
        // SM-time, which costs the overlapped decode of several row-groups 14 % of its throughput - hence off by default.
        const char *cl_env = getenv("PST_IDX_CLUSTER");
        const bool use_cluster = cl_env && cl_env[0] == '1';
        const int32_t *list = (const int32_t *)(arena + p->index_list_off);
        const int n_all = (int)p->index_pages.size(), n_big = use_cluster ? (int)p->index_big_count : 0;
        if (n_big > 0) {
            ck(launch_snappy_index_cluster(arena, pages, list, n_big, frag_pos, page_flag, s), "snappy cluster index launch");
            nl++;
        }
        if (n_all > n_big) {
            ck(launch_snappy_index(arena, pages, list + n_big, n_all - n_big, frag_pos, page_flag, s), "snappy index launch");
            nl++;
        }
    }
    if (ev) ck(cudaEventRecord(ev[1], s), "record");
    if (n_frags > 0) {
        ck(launch_snappy(arena, pages, frags, n_frags, multi, n_multi, frag_pos, page_flag, status, 0, s),
           "snappy fragment launch");
        nl++;
    }
    if (ev) ck(cudaEventRecord(ev[2], s), "record");
    if (n_multi > 0) {
        ck(launch_snappy(arena, pages, frags, n_frags, multi, n_multi, frag_pos, page_flag, status, 1, s),
           "snappy fallback launch");
        nl++;
    }
    if (!p->gzip_pages.empty()) {   // (timed together with the Snappy fallback slot: a plan normally has one codec)
        ck(launch_gzip_pages(arena, pages, (const int32_t *)(arena + p->gzip_list_off), (int)p->gzip_pages.size(),
                             status, s), "gzip launch");
        nl++;
    }
    if (ev) ck(cudaEventRecord(ev[3], s), "record");
    return nl;
}

extern "C" {

int pst_has_cuda(void) { return 1; }

int pst_ctx_create(int device, int64_t pinned_cache_bytes, int copy_threads, pst_ctx **out) {
    PST_TRY
    *out = nullptr;
    ck(cudaSetDevice(device), "cudaSetDevice");
    ck(cudaFree(0), "cuda context init");
    ck(configure_decode_kernels(), "configure kernels");
    ck(configure_copy_kernel(), "configure copy kernel");
    std::unique_ptr<pst_ctx> c(new pst_ctx());
    c->device = device;
    ck(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device), "SM count");
    c->cache_budget = pinned_cache_bytes;
    if (copy_threads < 0) {
        // staging copies (page cache -> pinned) of the cold path, 290 MB per C2 row-group: 16 threads 25 GB/s (11.5 ms),
        // 32 threads 40 GB/s (7.3 ms), 48 threads 38 GB/s (r2cold, 128 host cores).  Default: up to 32, of this rank's share
        // of the host cores when a launcher says how many ranks the box runs (LOCAL_WORLD_SIZE); PST_COPY_THREADS overrides.
        unsigned hc = std::thread::hardware_concurrency();
        unsigned ranks = 1;
        if (const char *e = getenv("LOCAL_WORLD_SIZE")) ranks = (unsigned)std::max(1, atoi(e));
        const unsigned share = hc > 2 ? std::max(4u, (hc - 1) / ranks) : 0u;
        copy_threads = hc > 2 ? (int)std::min(share, 32u) : 0;
        if (const char *e = getenv("PST_COPY_THREADS")) copy_threads = std::max(0, atoi(e));
    }
    c->local_cpus = local_cpus_of_device(device);
    if (copy_threads > 0) c->pool.reset(new CopyPool(copy_threads, c->local_cpus));
    *out = c.release();
    return 0;
    PST_CATCH(1)
}

void pst_ctx_destroy(pst_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    for (auto &kv : c->cache)
        if (kv.second.ptr) cudaFreeHost(kv.second.ptr);
    for (auto &b : c->ring) {
        if (b.ev) {
            cudaEventSynchronize(b.ev);
            cudaEventDestroy(b.ev);
        }
        if (b.ptr) cudaFreeHost(b.ptr);
    }
    delete c;
}

int pst_ctx_set_pinned_cache_bytes(pst_ctx *c, int64_t nbytes) {
    PST_TRY
    std::lock_guard<std::mutex> g(c->mu);
    ck(cudaSetDevice(c->device), "cudaSetDevice");
    if (nbytes < c->cache_bytes) {
        // shrinking: in-flight H2D copies may still read the cached buffers
        ck(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
        for (auto &kv : c->cache)
            if (kv.second.ptr) cudaFreeHost(kv.second.ptr);
        c->cache.clear();
        c->cache_bytes = 0;
    }
    c->cache_budget = nbytes;
    return 0;
    PST_CATCH(1)
}

int pst_ctx_stats_json(pst_ctx *c, char *buf, size_t cap) {
    int n = snprintf(buf, cap,
                     "{\"bytes_staged\":%lld,\"bytes_h2d\":%lld,\"pinned_cache_bytes\":%lld,\"pinned_cache_hits\":%lld,"
                     "\"pinned_cache_misses\":%lld,\"pages_decoded\":%lld,\"rowgroups_decoded\":%lld,"
                     "\"kernels_launched\":%lld}",
                     (long long)c->bytes_staged.load(), (long long)c->bytes_h2d.load(), (long long)c->cache_bytes,
                     (long long)c->cache_hits.load(), (long long)c->cache_misses.load(),
                     (long long)c->pages_decoded.load(), (long long)c->rowgroups_decoded.load(),
                     (long long)c->kernels_launched.load());
    return (n < 0 || (size_t)n >= cap) ? 1 : 0;
}

int pst_plan_upload(pst_ctx *c, pst_plan *p, uint64_t d_arena, uint64_t stream) {
    PST_TRY
    cudaStream_t s = (cudaStream_t)stream;
    std::lock_guard<std::mutex> g(c->mu);
    ck(cudaSetDevice(c->device), "cudaSetDevice");
    const int64_t need = p->raw_bytes;
    uint8_t *src = nullptr;
    auto it = c->cache.find(p->cache_key);
    if (it != c->cache.end() && it->second.cap >= need) {
        src = it->second.ptr;
        c->cache_hits++;
    } else {
        c->cache_misses++;
        // allocate + first-touch the pinned pages from a CPU next to the GPU (restored below)
        cpu_set_t old_affinity;
        const bool repinned = pin_current_thread(c->local_cpus, &old_affinity);
        struct Restore {
            bool on; cpu_set_t *old;
            ~Restore() { if (on) sched_setaffinity(0, sizeof(cpu_set_t), old); }
        } restore{repinned, &old_affinity};
        bool cacheable = c->cache_bytes + need <= c->cache_budget;
        PinnedBuf *slot;
        if (cacheable) {
            PinnedBuf nb;
            ck(cudaHostAlloc((void **)&nb.ptr, (size_t)need, cudaHostAllocDefault), "cudaHostAlloc (pinned cache)");
            nb.cap = need;
            c->cache_bytes += need;
            slot = &(c->cache[p->cache_key] = nb);
        } else {
            slot = &c->ring[c->ring_next];
            c->ring_next = (c->ring_next + 1) % 3;
            if (slot->ev) ck(cudaEventSynchronize(slot->ev), "staging ring wait");
            if (slot->cap < need) {
                if (slot->ptr) cudaFreeHost(slot->ptr);
                slot->ptr = nullptr;
                int64_t cap = need + need / 8;
                ck(cudaHostAlloc((void **)&slot->ptr, (size_t)cap, cudaHostAllocDefault), "cudaHostAlloc (staging ring)");
                slot->cap = cap;
            }
            if (!slot->ev) ck(cudaEventCreateWithFlags(&slot->ev, cudaEventDisableTiming), "cudaEventCreate");
        }
        fill_parallel(c, p, slot->ptr);
        c->bytes_staged += p->payload_bytes;
        src = slot->ptr;
        if (!cacheable) {
            h2d_copy(d_arena, src, need, s);
            ck(cudaEventRecord(slot->ev, s), "cudaEventRecord");
            c->bytes_h2d += need;
            return 0;
        }
    }
    h2d_copy(d_arena, src, need, s);
    c->bytes_h2d += need;
    return 0;
    PST_CATCH(1)
}

int pst_plan_decode(pst_ctx *c, pst_plan *p, uint64_t d_arena, uint64_t d_out, uint64_t d_status, uint64_t stream,
                    int *launches) {
    PST_TRY
    cudaStream_t s = (cudaStream_t)stream;
    uint8_t *arena = (uint8_t *)d_arena;
    uint8_t *outp = (uint8_t *)d_out;
    int32_t *status = (int32_t *)d_status;
    const DevCol *cols = (const DevCol *)(arena + p->cols_off);
    const DevPage *pages = (const DevPage *)(arena + p->pages_off);
    const int32_t *data = (const int32_t *)(arena + p->data_list_off);
    const int32_t *dict = (const int32_t *)(arena + p->dict_list_off);
    int nl = 0;
    nl += launch_snappy_stage(p, arena, status, s, nullptr);
    if (!p->ba_dict_pages.empty()) {
        ck(launch_ba_dict_index(arena, pages, cols, dict, (int)p->ba_dict_pages.size(), status, s), "dict index launch");
        nl++;
    }
    if (!p->copy_tiles.empty()) {
        ck(launch_copy_tiles(arena, outp, (const CopyTile *)(arena + p->copy_tiles_off), (int)p->copy_tiles.size(),
                             c ? c->sm_count : 148, s), "copy tiles launch");
        nl++;
    }
    if (!p->data_pages.empty()) {
        ck(launch_decode_pages(arena, outp, cols, pages, data, (int)p->data_pages.size(), status, s), "decode launch");
        nl++;
    }
    if (launches) *launches = nl;
    if (c) {
        c->pages_decoded += (int64_t)p->pages.size();
        c->rowgroups_decoded++;
        c->kernels_launched += nl;
    }
    return 0;
    PST_CATCH(1)
}

// Same launches as pst_plan_decode with a CUDA event between them, on the launching stream; synchronises and reports
// the device time of each kernel: ms[0..5] = Snappy fragment index, Snappy fragments, Snappy serial fallback,
// BYTE_ARRAY dictionary index, value tile copy, page decode.  Measurement aid for bench.py's roofline numbers -- not used by the readers.
int pst_plan_decode_timed(pst_ctx *c, pst_plan *p, uint64_t d_arena, uint64_t d_out, uint64_t d_status, uint64_t stream,
                          float *ms6) {
    PST_TRY
    cudaStream_t s = (cudaStream_t)stream;
    uint8_t *arena = (uint8_t *)d_arena;
    const DevCol *cols = (const DevCol *)(arena + p->cols_off);
    const DevPage *pages = (const DevPage *)(arena + p->pages_off);
    cudaEvent_t ev[7];
    for (auto &e : ev) ck(cudaEventCreate(&e), "cudaEventCreate");
    launch_snappy_stage(p, arena, (int32_t *)d_status, s, ev);
    ck(launch_ba_dict_index(arena, pages, cols, (const int32_t *)(arena + p->dict_list_off),
                            (int)p->ba_dict_pages.size(), (int32_t *)d_status, s), "dict index launch");
    ck(cudaEventRecord(ev[4], s), "record");
    ck(launch_copy_tiles(arena, (uint8_t *)d_out, (const CopyTile *)(arena + p->copy_tiles_off),
                         (int)p->copy_tiles.size(), c ? c->sm_count : 148, s), "copy tiles launch");
    ck(cudaEventRecord(ev[5], s), "record");
    ck(launch_decode_pages(arena, (uint8_t *)d_out, cols, pages, (const int32_t *)(arena + p->data_list_off),
                           (int)p->data_pages.size(), (int32_t *)d_status, s), "decode launch");
    ck(cudaEventRecord(ev[6], s), "record");
    ck(cudaEventSynchronize(ev[6]), "sync");
    for (int i = 0; i < 6; i++) ck(cudaEventElapsedTime(&ms6[i], ev[i], ev[i + 1]), "elapsed");
    for (auto &e : ev) cudaEventDestroy(e);
    return 0;
    PST_CATCH(1)
}

// ---- post-processing wrappers -------------------------------------------------------------------------------
#define WRAP(expr, what)              \
    PST_TRY                           \
    ck((expr), what);                 \
    return 0;                         \
    PST_CATCH(1)

int pst_nullable_to_f64(uint64_t values, uint64_t valid, int64_t n, int physical_type, int bit_width, int is_unsigned,
                        uint64_t out, uint64_t stream) {
    WRAP(launch_nullable_to_f64((const void *)values, (const uint8_t *)valid, n, physical_type, bit_width, is_unsigned,
                                (double *)out, (cudaStream_t)stream), "nullable_to_f64")
}
int pst_narrow_int32(uint64_t src, int64_t n, int bit_width, uint64_t dst, uint64_t stream) {
    WRAP(launch_narrow_int32((const int32_t *)src, n, bit_width, (void *)dst, (cudaStream_t)stream), "narrow_int32")
}
int pst_gather_rows(uint64_t src, uint64_t idx, int64_t n_out, int64_t row_bytes, uint64_t dst, uint64_t stream) {
    WRAP(launch_gather_rows((const uint8_t *)src, (const int64_t *)idx, n_out, row_bytes, (uint8_t *)dst,
                            (cudaStream_t)stream), "gather_rows")
}
int pst_npy_batch(uint64_t base, uint64_t offs, uint64_t lens, uint64_t row_idx, int64_t n, int64_t data_off,
                  int64_t payload_bytes, uint64_t dst, uint64_t d_status, uint64_t stream) {
    WRAP(launch_npy_batch((const uint8_t *)base, (const int64_t *)offs, (const int32_t *)lens, (const int64_t *)row_idx,
                          n, data_off, payload_bytes, (uint8_t *)dst, (int32_t *)d_status, (cudaStream_t)stream),
         "npy_batch")
}
int pst_zip_inflate_batch(uint64_t base, uint64_t offs, uint64_t lens, uint64_t row_idx, int64_t n, int64_t member_bytes,
                          uint64_t dst, uint64_t d_status, uint64_t stream) {
    WRAP(launch_zip_inflate_batch((const uint8_t *)base, (const int64_t *)offs, (const int32_t *)lens,
                                  (const int64_t *)row_idx, n, member_bytes, (uint8_t *)dst, (int32_t *)d_status,
                                  (cudaStream_t)stream), "zip_inflate_batch")
}
int pst_blob_prefix(uint64_t base, uint64_t offs, uint64_t lens, int64_t n, int k, uint64_t dst, uint64_t stream) {
    WRAP(launch_blob_prefix((const uint8_t *)base, (const int64_t *)offs, (const int32_t *)lens, n, k, (uint8_t *)dst,
                            (cudaStream_t)stream), "blob_prefix")
}
int64_t pst_png_work_bytes(int height, int width, int channels, int sample_bytes) {
    return png_work_bytes(height, width, channels, sample_bytes);
}
int pst_png_batch(uint64_t base, uint64_t offs, uint64_t lens, uint64_t row_idx, int64_t n, int height, int width,
                  int channels, int sample_bytes, uint64_t dst, uint64_t d_work, uint64_t d_status, uint64_t stream) {
    WRAP(launch_png_batch((const uint8_t *)base, (const int64_t *)offs, (const int32_t *)lens, (const int64_t *)row_idx,
                          n, height, width, channels, sample_bytes, (uint8_t *)dst, (uint8_t *)d_work,
                          (int32_t *)d_status, (cudaStream_t)stream), "png_batch")
}
int pst_mask_in_set_i64(uint64_t keys, int key_bytes, int key_unsigned, int64_t n, uint64_t set_sorted, int64_t set_n,
                        uint64_t mask, uint64_t stream) {
    WRAP(launch_mask_in_set((const void *)keys, key_bytes, key_unsigned, n, (const int64_t *)set_sorted, set_n,
                            (uint8_t *)mask, (cudaStream_t)stream), "mask_in_set")
}
int pst_mask_md5_split_i64(uint64_t keys, int key_bytes, int key_unsigned, int64_t n, double lo, double hi,
                           uint64_t mask, uint64_t stream) {
    WRAP(launch_mask_md5_split((const void *)keys, key_bytes, key_unsigned, n, lo, hi, (uint8_t *)mask,
                               (cudaStream_t)stream), "mask_md5_split")
}
int64_t pst_compact_tmp_bytes(int64_t n) { return compact_tmp_bytes(n); }
int pst_mask_compact(uint64_t mask, int64_t n, uint64_t out_idx, uint64_t d_count, uint64_t d_tmp, uint64_t stream) {
    WRAP(launch_mask_compact((const uint8_t *)mask, n, (int64_t *)out_idx, (int64_t *)d_count, (void *)d_tmp,
                             (cudaStream_t)stream), "mask_compact")
}
int pst_normalize(uint64_t src, int src_dtype, int64_t n, float mean, float stddev, uint64_t dst, int dst_dtype,
                  uint64_t stream) {
    WRAP(launch_normalize((const void *)src, src_dtype, n, mean, stddev, (void *)dst, dst_dtype, (cudaStream_t)stream),
         "normalize")
}
int pst_ngram_valid_starts(uint64_t ts, int64_t n, int length, int64_t delta, uint64_t ok, uint64_t d_status,
                           uint64_t stream) {
    WRAP(launch_ngram_valid_starts((const int64_t *)ts, n, length, delta, (uint8_t *)ok, (int32_t *)d_status,
                                   (cudaStream_t)stream), "ngram_valid_starts")
}
int pst_ngram_gather(uint64_t src, uint64_t starts, int64_t n_windows, int length, int64_t row_bytes, uint64_t dst,
                     uint64_t stream) {
    WRAP(launch_ngram_gather((const uint8_t *)src, (const int64_t *)starts, n_windows, length, row_bytes,
                             (uint8_t *)dst, (cudaStream_t)stream), "ngram_gather")
}
int pst_sanitize(uint64_t src, int64_t n, int kind, uint64_t dst, uint64_t stream) {
    WRAP(launch_sanitize((const void *)src, n, kind, (void *)dst, (cudaStream_t)stream), "sanitize")
}
int pst_list_uniform(uint64_t rep, uint64_t def, int64_t n, int max_def, int64_t list_len, uint64_t d_flags,
                     uint64_t stream) {
    WRAP(launch_list_uniform((const uint8_t *)rep, (const uint8_t *)def, n, max_def, list_len, (int64_t *)d_flags,
                             (cudaStream_t)stream), "list_uniform")
}

}  // extern "C"
