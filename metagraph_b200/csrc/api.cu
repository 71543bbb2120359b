// C-ABI (include/mgb.h) of the B200-native `metagraph align` hot path: index upload, the three
// sm_100a kernels (query preparation, quad-per-strand exact seeding, warp-per-read
// seed-and-extend) and result unpacking.
//
// When compiled by g++ with -DMGB_HOST_EMU (tests/emu/ only) the same logic runs with
// one-lane "warps" on host memory so that it can be checked against the oracle on a machine
// without a GPU. That build is test infrastructure; the product library is always the nvcc
// build and has no CPU path.
#include <algorithm>
#include <atomic>
#include <thread>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>
#if defined(_OPENMP)
#include <omp.h>
#endif

// Threads per OpenMP team of the result-unpacking loops. They are short, memory-bound loops that run on two
// pipeline lanes at once: beyond a few cores they gain nothing, and with one team per lane as wide as a
// 128-thread host the teams spend their time spinning against each other (measured: a 36 ms call took 470 ms,
// profiles/r2_e2e_probe.txt). mgb_set_host_threads() can only lower the number.
[[maybe_unused]] static inline int host_team() {
#if defined(_OPENMP)
    const int n = omp_get_max_threads();
    return n < 16 ? (n < 1 ? 1 : n) : 16;
#else
    return 1;
#endif
}

#include "../../include/mgb.h"
#if !defined(MGB_HOST_EMU)
#define MGB_NARROW_ONLY 1        // this translation unit's kernels: DNA block layout only (see kernels.cuh)
#define MGB_BASIC_ONLY 1         // ... and BASIC-mode graphs only (api_canonical.cu holds the CANONICAL-mode k_align)
#endif
#define MGB_KERNEL_NS kern_dna
#include "kernels.cuh"
#include "host_common.hpp"
#include "index_build.hpp"

#if !defined(MGB_HOST_EMU)
#include <cuda_runtime.h>
#endif

using namespace mgb;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string &msg) { g_err = msg; return code; }


#if !defined(MGB_HOST_EMU)
} // namespace
// the same kernels compiled for the alphabet-generic layout (api_generic.cu)
namespace kern_any {
cudaError_t launch_radj_bwd(unsigned grid, const mgb::RadjArgs &a);
cudaError_t launch_sfx_extend(unsigned grid, const mgb::SfxArgs &a);
cudaError_t launch_seed(unsigned grid, cudaStream_t s, const mgb::SeedArgs &a);
cudaError_t launch_premap(unsigned grid, cudaStream_t s, const mgb::SeedArgs &a);
cudaError_t launch_subk(unsigned grid, cudaStream_t s, const mgb::SubkArgs &a, uint32_t chunks_per_strand);
cudaError_t launch_align(unsigned grid, size_t smem_block, cudaStream_t s, const mgb::AlignArgs &a);
cudaError_t align_occupancy(size_t smem_limit, size_t smem_block, int *blocks_per_sm);
}
// k_align for CANONICAL-mode DNA graphs (api_canonical.cu)
namespace kern_canon {
cudaError_t launch_align(unsigned grid, size_t smem_block, cudaStream_t s, const mgb::AlignArgs &a);
cudaError_t align_occupancy(size_t smem_limit, size_t smem_block, int *blocks_per_sm);
}
namespace {
#endif

#if !defined(MGB_HOST_EMU)
#define CUDA_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) \
    return fail(MGB_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(e_)); } while (0)
#endif

// ---------------------------------------------------------------------------------------
// device memory helpers
// ---------------------------------------------------------------------------------------
#if defined(MGB_HOST_EMU)
struct Stream { int dummy; };
inline int dev_alloc(void **p, size_t bytes, Stream&) { *p = std::calloc(1, bytes ? bytes : 1); return *p ? 0 : MGB_ERR_CUDA; }
inline void dev_free(void *p, Stream&) { std::free(p); }
inline int h2d(void *d, const void *h, size_t bytes, Stream&) { std::memcpy(d, h, bytes); return 0; }
inline int d2h(void *h, const void *d, size_t bytes, Stream&) { std::memcpy(h, d, bytes); return 0; }
inline int dev_zero(void *d, size_t bytes, Stream&) { std::memset(d, 0, bytes); return 0; }
#else
struct Stream { cudaStream_t s; };
inline int dev_alloc(void **p, size_t bytes, Stream &st) {
    cudaError_t e = cudaMallocAsync(p, bytes ? bytes : 16, st.s);
    if (e != cudaSuccess) { g_err = std::string("cudaMallocAsync: ") + cudaGetErrorString(e); return MGB_ERR_CUDA; }
    return 0;
}
inline void dev_free(void *p, Stream &st) { if (p) cudaFreeAsync(p, st.s); }
inline int h2d(void *d, const void *h, size_t bytes, Stream &st) {
    cudaError_t e = cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, st.s);
    if (e != cudaSuccess) { g_err = std::string("H2D: ") + cudaGetErrorString(e); return MGB_ERR_CUDA; }
    return 0;
}
inline int d2h(void *h, const void *d, size_t bytes, Stream &st) {
    cudaError_t e = cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, st.s);
    if (e != cudaSuccess) { g_err = std::string("D2H: ") + cudaGetErrorString(e); return MGB_ERR_CUDA; }
    return 0;
}
inline int dev_zero(void *d, size_t bytes, Stream &st) {
    cudaError_t e = cudaMemsetAsync(d, 0, bytes, st.s);
    if (e != cudaSuccess) { g_err = std::string("memset: ") + cudaGetErrorString(e); return MGB_ERR_CUDA; }
    return 0;
}
#endif

// Host buffers that receive the packed results. On the device build they are pinned and
// recycled through a process-wide pool (cudaHostAlloc of GB-sized buffers costs 100s of ms), so a
// mgb_results_t owns its buffers zero-copy and hands them back in mgb_results_free().
struct HostBuf { char *p = nullptr; size_t cap = 0; };
#if defined(MGB_HOST_EMU)
inline HostBuf hostbuf_acquire(size_t bytes) { HostBuf b; b.cap = bytes ? bytes : 1; b.p = (char*)std::malloc(b.cap); return b; }
inline void hostbuf_release(HostBuf b) { std::free(b.p); }
#else
struct PinnedPool {
    std::mutex mu;
    std::vector<HostBuf> free_list;
    HostBuf acquire(size_t bytes) {
        {
            std::lock_guard<std::mutex> lk(mu);
            size_t best = free_list.size();
            for (size_t i = 0; i < free_list.size(); ++i)
                if (free_list[i].cap >= bytes && (best == free_list.size() || free_list[i].cap < free_list[best].cap)) best = i;
            if (best != free_list.size()) { HostBuf b = free_list[best]; free_list.erase(free_list.begin() + best); return b; }
        }
        HostBuf b;
        size_t want = bytes + bytes / 8 + 4096;
        if (cudaHostAlloc((void**)&b.p, want, cudaHostAllocDefault) != cudaSuccess) { b.p = nullptr; b.cap = 0; return b; }
        b.cap = want;
        return b;
    }
    void release(HostBuf b) {
        if (!b.p) return;
        std::lock_guard<std::mutex> lk(mu);
        if (free_list.size() >= 32) { cudaFreeHost(b.p); return; }
        free_list.push_back(b);
    }
};
PinnedPool g_pinned_pool;
inline HostBuf hostbuf_acquire(size_t bytes) { return g_pinned_pool.acquire(bytes); }
inline void hostbuf_release(HostBuf b) { g_pinned_pool.release(b); }
#endif

// Device working memory of one call. On the device build the buffers belong to a Workspace that the
// index keeps between calls (slot i = the i-th request of the call): cudaMalloc of GB-sized arenas per
// call is slow, and stream-ordered pools stall when two streams trade memory back and forth.
#if defined(MGB_HOST_EMU)
struct Workspace {};
struct DevBufs {               // frees everything it owns on scope exit
    Stream &st;
    std::vector<void*> ptrs;
    DevBufs(Stream &s, Workspace*, size_t) : st(s) {}
    ~DevBufs() { for (void *p : ptrs) dev_free(p, st); }
    template <class T> int alloc(T **p, size_t count) {
        void *v = nullptr;
        int rc = dev_alloc(&v, count * sizeof(T), st);
        if (rc) return rc;
        ptrs.push_back(v);
        *p = (T*)v;
        return 0;
    }
    size_t capacity_of_next() const { return 0; }
};
#else
struct Workspace {
    std::vector<void*> ptr;
    std::vector<size_t> cap;
    cudaStream_t stream = nullptr;       // created on first use, kept with the buffers
    cudaEvent_t ev[6] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };
    int ready() {
        if (stream) return 0;
        cudaError_t e = cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking);
        for (auto &x : ev) if (e == cudaSuccess) e = cudaEventCreate(&x);
        if (e != cudaSuccess) { g_err = std::string("stream/event creation: ") + cudaGetErrorString(e); return MGB_ERR_CUDA; }
        return 0;
    }
    ~Workspace() {
        for (void *p : ptr) if (p) cudaFree(p);
        for (auto &x : ev) if (x) cudaEventDestroy(x);
        if (stream) cudaStreamDestroy(stream);
    }
    int get(size_t slot, size_t bytes, void **out) {
        if (slot >= ptr.size()) { ptr.resize(slot + 1, nullptr); cap.resize(slot + 1, 0); }
        if (cap[slot] < bytes || !ptr[slot]) {
            if (ptr[slot]) cudaFree(ptr[slot]);
            ptr[slot] = nullptr; cap[slot] = 0;
            size_t want = bytes + bytes / 8 + 256;
            cudaError_t e = cudaMalloc(&ptr[slot], want);
            if (e != cudaSuccess) { cudaGetLastError(); want = bytes ? bytes : 16; e = cudaMalloc(&ptr[slot], want); }
            if (e != cudaSuccess) {
                ptr[slot] = nullptr;
                g_err = std::string("cudaMalloc: ") + cudaGetErrorString(e);
                return MGB_ERR_CUDA;
            }
            cap[slot] = want;
        }
        *out = ptr[slot];
        return 0;
    }
};
struct DevBufs {
    Stream &st;
    Workspace *ws;
    size_t slot;
    DevBufs(Stream &s, Workspace *w, size_t slot_base) : st(s), ws(w), slot(slot_base) {}
    template <class T> int alloc(T **p, size_t count) {
        void *v = nullptr;
        int rc = ws->get(slot++, count * sizeof(T), &v);
        if (rc) return rc;
        *p = (T*)v;
        return 0;
    }
    size_t capacity_of_next() const { return slot < ws->cap.size() ? ws->cap[slot] : 0; }
};
#endif

} // namespace

// ---------------------------------------------------------------------------------------
// opaque types
// ---------------------------------------------------------------------------------------
struct mgb_index {
    // owns its device buffers and cached workspaces: every error path of mgb_index_create releases what was uploaded
    ~mgb_index();
    IndexView view;            // device pointers (host pointers in the emulation build)
    int device = 0;
    uint64_t device_bytes = 0;
    int num_sms = 1;
    std::vector<void*> bufs;
#if defined(MGB_HOST_EMU)
    std::vector<uint32_t> rc_host;   // PRIMARY graphs: rcs / rcp / palin (mgb_index_set_mode)
#endif
    int alphabet = MGB_ALPHABET_DNA;
    AlphabetTables at;
    // working memory recycled between calls (at most 4 sets are kept)
    mutable std::mutex ws_mu;
    mutable std::vector<Workspace*> ws_free;
    Workspace* ws_acquire() const {
        std::lock_guard<std::mutex> lk(ws_mu);
        if (ws_free.empty()) return new Workspace();
        Workspace *w = ws_free.back(); ws_free.pop_back(); return w;
    }
    void ws_release(Workspace *w) const {
        std::lock_guard<std::mutex> lk(ws_mu);
        if (ws_free.size() < 4) ws_free.push_back(w); else delete w;
    }
#if defined(MGB_HOST_EMU)
    HostIndex host;
    std::vector<uint2> radj_host;
#endif
};

mgb_index::~mgb_index() {
#if !defined(MGB_HOST_EMU)
    if (!bufs.empty()) { cudaSetDevice(device); for (void *p : bufs) cudaFree(p); }
#endif
    for (Workspace *w : ws_free) delete w;
}

struct mgb_results {
    uint32_t n_reads = 0;
    std::vector<uint64_t> first;
    std::vector<uint32_t> count;
    mgb_alignment_t *alns = nullptr;         // n_alns records in read order, inside alns_buf
    uint64_t n_alns = 0;
    HostBuf alns_buf;
    std::vector<HostBuf> heaps;              // host copies of the output heaps (recycled on free)
    std::vector<uint64_t> heap_bytes;        // used bytes of each heap
    // where the packed records of read r start: heap src_heap[r], byte src_off[r] (mgb_results_export)
    std::vector<uint64_t> src_off;
    std::vector<uint32_t> src_heap;
    uint32_t result_nodes = 0;
    ~mgb_results() { for (HostBuf b : heaps) hostbuf_release(b); hostbuf_release(alns_buf); }
    mgb_stats_t stats;
};

extern "C" {

const char* mgb_last_error(void) { return g_err.c_str(); }

int mgb_device_count(void) {
#if defined(MGB_HOST_EMU)
    return 1;
#else
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
#endif
}

int mgb_index_create(const uint8_t *W, const uint8_t *last, uint64_t n_plus_1, const uint64_t *F,
                     const uint8_t *valid, uint32_t k, int alphabet, uint32_t suffix_len,
                     int device, mgb_index_t **out) {
    if (!W || !last || !F || !out) return fail(MGB_ERR_INVALID_ARGUMENT, "null argument");
    std::unique_ptr<mgb_index> idx(new mgb_index());
    if (!alphabet_tables(alphabet, &idx->at)) return fail(MGB_ERR_UNSUPPORTED, "unknown alphabet");
    idx->alphabet = alphabet;
    const uint32_t sigma = idx->at.sigma;
    const bool force_wide = std::getenv("MGB_TEST_WIDE") != nullptr;   // test knob: generic layout for DNA
    HostIndex hloc;
#if defined(MGB_HOST_EMU)
    HostIndex &h = idx->host;
#else
    HostIndex &h = hloc;
#endif
    if (suffix_len == 0) {
        // reference default 12 (cli/config/config.cpp:24-25). Here: as long as the table has at most
        // ~12 entries per edge and 2^30 entries (8 GiB): a cold k-mer lookup that misses in the table
        // costs one load instead of a chain of tighten_range steps, and on a strand that does not match
        // every k-mer is a cold lookup (k_seed 7.6 -> 5.5 ms per 200 k reads going from s = 13 to 14 on a
        // 20 M-node graph). HBM capacity is traded for latency.
        uint64_t n = n_plus_1 - 1;
        suffix_len = 1;
        uint64_t entries = sigma - 1;
        while (suffix_len < 15 && suffix_len + 1 <= k - 1 && entries * (sigma - 1) <= 12 * n + 1024
               && entries * (sigma - 1) <= (1ull << 30) + (1ull << 26)) {
            entries *= sigma - 1; ++suffix_len;
        }
    }
    // levels up to 8 are refined on the host, the rest by k_sfx_extend on the device
    const uint32_t sfx_target = suffix_len > k - 1 ? k - 1 : suffix_len;
    // (host levels: at most 4^8 = 65k entries' worth)
    uint32_t sfx_host = 0;
    for (uint64_t e = 1; sfx_host < sfx_target && e * (sigma - 1) <= 65536; e *= sigma - 1) ++sfx_host;
    if (sfx_host == 0 && sfx_target) sfx_host = 1;
    try {
        build_host_index(W, last, n_plus_1, F, valid, k, sfx_host, &h, sigma, force_wide);
    } catch (const std::exception &e) {
        return fail(MGB_ERR_INVALID_ARGUMENT, e.what());
    }
#if defined(MGB_HOST_EMU)
    (void)device; (void)hloc;
    idx->view = h.view();
    {
        uint64_t cur_num = 1;
        for (uint32_t i = 0; i < sfx_host; ++i) cur_num *= (sigma - 1);
        for (uint32_t len = sfx_host + 1; len <= sfx_target && sfx_host >= 1; ++len) {
            const uint64_t nxt_num = cur_num * (sigma - 1);
            std::vector<uint32_t> nxt(2 * nxt_num);
            SfxArgs sa { idx->view, idx->view.sfx, nxt.data(), cur_num };
            sa.ix.sfx = nullptr; sa.ix.sfx_len = 0;
            for (uint64_t o = 0; o < nxt_num; ++o) sfx_extend_item(sa, o);
            h.sfx.swap(nxt);
            cur_num = nxt_num;
            idx->view.sfx = h.sfx.data(); idx->view.sfx_len = len;
        }
    }
    {
        const uint64_t n = idx->view.n;
        std::vector<uint32_t> bwd_arr(n + 1, 0);
        std::vector<uint8_t> c0(n + 1, 0), c1(n + 1, 0), multi(n + 1, 0);
        idx->radj_host.assign(n + 1, uint2{0, 0});
        RadjArgs ra { idx->view, bwd_arr.data(), c0.data(), c1.data(), multi.data(), idx->radj_host.data(), n, nullptr };
        for (uint64_t e = 1; e <= n; ++e) radj_bwd_item(ra, e);
        for (uint32_t r = 0; r + 2 < k; ++r) {
            for (uint64_t e = 1; e <= n; ++e) ra.c_nxt[e] = ra.c_cur[bwd_arr[e]];
            std::swap(ra.c_cur, ra.c_nxt);
        }
        for (uint64_t e = 1; e <= n; ++e)
            idx->radj_host[e] = uint2{ bwd_arr[e], (uint32_t)ra.c_cur[e] | ((uint32_t)multi[e] << radj_multi_shift(idx->view)) };
        idx->view.radj = idx->radj_host.data();
    }
    idx->device_bytes = h.blocks.size() * 4 + h.wW.size();
#else
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(MGB_ERR_NO_DEVICE, "no CUDA device available (the B200 kernels have no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(MGB_ERR_INVALID_ARGUMENT, "bad device ordinal");
    CUDA_TRY(cudaSetDevice(device));
    idx->device = device;
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    idx->num_sms = prop.multiProcessorCount;
    {   // Every access of this path to the index and to the per-read arenas is a single 8 .. 64-byte record at a
        // data-dependent address; L2 fetches 128-byte lines from HBM for them (k_seed: 146 bytes of DRAM traffic per
        // 8-byte adjacency / 32-byte hash-bucket access). MGB_L2_FETCH=32 / 64 asks for smaller fetches (process-wide).
        // (measured: 32-byte fetches cut k_seed's DRAM bytes from 35 to 12 GB per 1 M reads and made it SLOWER, 12.9 ->
        // 15.7 ms: the access rate, not the bytes, is what HBM limits here. Left at the default unless MGB_L2_FETCH is set.)
        size_t gran = 0;
        if (const char *e = std::getenv("MGB_L2_FETCH")) gran = (size_t)std::atoi(e);
        if (gran) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
        cudaGetLastError();
    }
    {   // keep stream-ordered allocations cached between calls
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t thr = UINT64_MAX;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
        }
    }
    IndexView v = h.view();
    auto upload = [&](const std::vector<uint32_t> &src, const uint32_t **dst) -> int {
        if (src.empty()) { *dst = nullptr; return 0; }
        void *p = nullptr;
        cudaError_t e = cudaMalloc(&p, src.size() * 4);
        if (e != cudaSuccess) return fail(MGB_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
        idx->bufs.push_back(p);
        e = cudaMemcpy(p, src.data(), src.size() * 4, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) return fail(MGB_ERR_CUDA, std::string("cudaMemcpy: ") + cudaGetErrorString(e));
        idx->device_bytes += src.size() * 4;
        *dst = (const uint32_t*)p;
        return 0;
    };
    int rc;
    if ((rc = upload(h.blocks, &v.blocks))) return rc;
    if ((rc = upload(h.blk_rank, &v.blk_rank))) return rc;
    if ((rc = upload(h.sel_last, &v.sel_last))) return rc;
    for (int c = 0; c < kSigmaDNA; ++c)
        if ((rc = upload(h.sel_W[c], &v.sel_W[c]))) return rc;
    if ((rc = upload(h.valid, &v.valid))) return rc;
    if ((rc = upload(h.sfx, &v.sfx))) return rc;
    if (h.wide) {
        std::vector<uint32_t> wW32((h.wW.size() + 3) / 4);
        std::memcpy(wW32.data(), h.wW.data(), h.wW.size());
        const uint32_t *p32 = nullptr;
        if ((rc = upload(wW32, &p32))) return rc;
        v.wW = (const uint8_t*)p32;
        if ((rc = upload(h.wl, &v.wl))) return rc;
        if ((rc = upload(h.wrank, &v.wrank))) return rc;
        if ((rc = upload(h.wsel, &v.wsel))) return rc;
        if ((rc = upload(h.wadj, &v.wadj))) return rc;
        v.adj = nullptr;
    } else {
        void *p = nullptr;
        cudaError_t e = cudaMalloc(&p, h.adj.size() * sizeof(uint2));
        if (e != cudaSuccess) return fail(MGB_ERR_CUDA, std::string("cudaMalloc(adj): ") + cudaGetErrorString(e));
        idx->bufs.push_back(p);
        e = cudaMemcpy(p, h.adj.data(), h.adj.size() * sizeof(uint2), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) return fail(MGB_ERR_CUDA, std::string("cudaMemcpy(adj): ") + cudaGetErrorString(e));
        idx->device_bytes += h.adj.size() * sizeof(uint2);
        v.adj = (const uint2*)p;
    }
    idx->view = v;
    // deeper suffix-range levels on the device
    {
        uint64_t cur_num = 1;
        for (uint32_t i = 0; i < sfx_host; ++i) cur_num *= (sigma - 1);
        const uint32_t *cur = v.sfx;
        for (uint32_t len = sfx_host + 1; len <= sfx_target && sfx_host >= 1; ++len) {
            uint32_t *nxt = nullptr;
            const uint64_t nxt_num = cur_num * (sigma - 1);
            cudaError_t e = cudaMalloc((void**)&nxt, nxt_num * 8);
            if (e != cudaSuccess) return fail(MGB_ERR_CUDA, std::string("cudaMalloc(sfx): ") + cudaGetErrorString(e));
            SfxArgs sa { idx->view, cur, nxt, cur_num };
            sa.ix.sfx = nullptr; sa.ix.sfx_len = 0;
            uint64_t blocks = std::min<uint64_t>((nxt_num + 31) / 32, (uint64_t)idx->num_sms * 32);
            e = idx->view.wide ? kern_any::launch_sfx_extend((unsigned)blocks, sa) : kern_dna::launch_sfx_extend((unsigned)blocks, sa);
            if (e == cudaSuccess) e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { cudaFree(nxt); return fail(MGB_ERR_CUDA, std::string("k_sfx_extend: ") + cudaGetErrorString(e)); }   // (idx's destructor frees the rest)
            // the previous level is no longer needed (the host-built one is freed with the index)
            if (len > sfx_host + 1) {
                cudaFree((void*)cur);
                idx->bufs.erase(std::find(idx->bufs.begin(), idx->bufs.end(), (void*)cur));
                idx->device_bytes -= cur_num * 8;
            }
            idx->bufs.push_back(nxt);
            idx->device_bytes += nxt_num * 8;
            cur = nxt; cur_num = nxt_num;
            idx->view.sfx = nxt; idx->view.sfx_len = len;
        }
    }
    // reverse adjacency records
    {
        const uint64_t n = idx->view.n;
        uint32_t *bwd_arr = nullptr; uint8_t *c0 = nullptr, *c1 = nullptr, *multi = nullptr; uint2 *radj = nullptr;
        cudaError_t e = cudaMalloc((void**)&bwd_arr, (n + 1) * 4);
        if (e == cudaSuccess) e = cudaMalloc((void**)&c0, n + 1);
        if (e == cudaSuccess) e = cudaMalloc((void**)&c1, n + 1);
        if (e == cudaSuccess) e = cudaMalloc((void**)&multi, n + 1);
        if (e == cudaSuccess) e = cudaMalloc((void**)&radj, (n + 1) * sizeof(uint2));
        if (e != cudaSuccess) {
            cudaFree(bwd_arr); cudaFree(c0); cudaFree(c1); cudaFree(multi); cudaFree(radj);     // cudaFree(nullptr) is a no-op
            return fail(MGB_ERR_CUDA, std::string("cudaMalloc(radj): ") + cudaGetErrorString(e));
        }
        cudaMemset(bwd_arr, 0, 4); cudaMemset(c0, 0, 1); cudaMemset(c1, 0, 1); cudaMemset(multi, 0, 1);
        cudaMemset(radj, 0, sizeof(uint2));
        // k-mer hash index (index.cuh kh_*): DNA block layout, k <= 31 (62-bit keys). The k-mer of every edge falls out
        // of the gather rounds below; 12 bytes per slot at a load factor of 0.7. Skipped (the seeding kernel then looks
        // k-mers up through the suffix-range table and tighten_range) if the memory is not there or MGB_NO_KMER_HASH is set.
        unsigned long long *kmer = nullptr, *kh_keys = nullptr; uint32_t *kh_vals = nullptr;
        const uint64_t kh_slots = (((uint64_t)((double)(n + 1) / 0.7) + 64) + 3) & ~3ull;     // whole buckets of 4
        if (!idx->view.wide && k >= 3 && k <= 31 && !std::getenv("MGB_NO_KMER_HASH")) {
            size_t free_b = 0, total_b = 0;
            cudaMemGetInfo(&free_b, &total_b);
            if ((double)free_b > 1.2 * (8.0 * (n + 1) + 12.0 * kh_slots) + 16e9
                    && cudaMalloc((void**)&kmer, (n + 1) * 8) == cudaSuccess) {
                if (cudaMalloc((void**)&kh_keys, kh_slots * 8) != cudaSuccess
                        || cudaMalloc((void**)&kh_vals, kh_slots * 4) != cudaSuccess) {
                    cudaGetLastError();
                    cudaFree(kmer); cudaFree(kh_keys); cudaFree(kh_vals);
                    kmer = nullptr; kh_keys = nullptr; kh_vals = nullptr;
                } else {
                    cudaMemset(kmer, 0xFF, 8);                  // position 0: no k-mer
                    cudaMemset(kh_keys, 0, kh_slots * 8);
                }
            } else cudaGetLastError();
        }
        RadjArgs ra { idx->view, bwd_arr, c0, c1, multi, radj, n, kmer };
        const unsigned grid = (unsigned)idx->num_sms * 16;
        e = idx->view.wide ? kern_any::launch_radj_bwd(grid, ra) : kern_dna::launch_radj_bwd(grid, ra);
        for (uint32_t r = 0; r + 2 < k; ++r) {          // k - 2 bwd steps
            kern_dna::k_radj_gather<<<grid, 256>>>(ra, 2 * ((int)k - 3 - (int)r));
            std::swap(ra.c_cur, ra.c_nxt);
        }
        kern_dna::k_radj_pack<<<grid, 256>>>(ra, radj_multi_shift(idx->view));
        if (kmer) kern_dna::k_kmer_insert<<<grid, 256>>>(kmer, n, kh_keys, kh_vals, kh_slots);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        cudaFree(bwd_arr); cudaFree(c0); cudaFree(c1); cudaFree(multi); cudaFree(kmer);
        if (e != cudaSuccess) {
            cudaFree(radj); cudaFree(kh_keys); cudaFree(kh_vals);
            return fail(MGB_ERR_CUDA, std::string("radj / k-mer hash build: ") + cudaGetErrorString(e));
        }
        idx->bufs.push_back(radj);
        idx->device_bytes += (n + 1) * sizeof(uint2);
        idx->view.radj = radj;
        if (kh_keys) {
            idx->bufs.push_back(kh_keys); idx->bufs.push_back(kh_vals);
            idx->device_bytes += kh_slots * 12;
            idx->view.kh_keys = kh_keys; idx->view.kh_vals = kh_vals; idx->view.kh_slots = kh_slots;
        }
    }
#endif
    *out = idx.release();
    return MGB_OK;
}

int mgb_index_set_mode(mgb_index_t *index, int mode) {
    if (!index) return fail(MGB_ERR_INVALID_ARGUMENT, "null argument");
    if (mode < 0 || mode > 2) return fail(MGB_ERR_INVALID_ARGUMENT, "mode must be 0 (BASIC), 1 (CANONICAL) or 2 (PRIMARY)");
    if (mode != 0 && !index->at.has_complement)
        return fail(MGB_ERR_BAD_CONFIG, "CANONICAL / PRIMARY modes need a DNA graph");
    if (mode != 0 && index->view.wide)
        return fail(MGB_ERR_UNSUPPORTED, "CANONICAL / PRIMARY modes are served on the DNA block layout only");
    if (mode == 2 && !index->view.rcs) {
        // PRIMARY graph: the rc-strand jump tables of the CanonicalDBG semantics (IndexView::rcs / rcp / palin)
        if (index->view.k > (uint32_t)kMaxPrimaryK)
            return fail(MGB_ERR_UNSUPPORTED, "PRIMARY graphs: k above 128 is not supported");
        const uint64_t n = index->view.n;
        const bool even_k = index->view.k % 2 == 0;
        const size_t pal_words = (n >> 5) + 1;
#if defined(MGB_HOST_EMU)
        index->rc_host.assign(2 * (n + 1) + (even_k ? pal_words : 0), 0);
        RcArgs ra { index->view, index->rc_host.data(), index->rc_host.data() + (n + 1),
                    even_k ? index->rc_host.data() + 2 * (n + 1) : nullptr, n };
        for (uint64_t e = 1; e <= n; ++e) rc_tables_item(ra, e);
#else
        CUDA_TRY(cudaSetDevice(index->device));
        uint32_t *buf = nullptr;
        const size_t words = 2 * (n + 1) + (even_k ? pal_words : 0);
        cudaError_t e = cudaMalloc((void**)&buf, words * 4);
        if (e != cudaSuccess) return fail(MGB_ERR_CUDA, std::string("cudaMalloc(rc tables): ") + cudaGetErrorString(e));
        cudaMemset(buf, 0, words * 4);
        RcArgs ra { index->view, buf, buf + (n + 1), even_k ? buf + 2 * (n + 1) : nullptr, n };
        kern_dna::k_rc_tables<<<(unsigned)index->num_sms * 16, 128>>>(ra);
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { cudaFree(buf); return fail(MGB_ERR_CUDA, std::string("rc tables: ") + cudaGetErrorString(e)); }
        index->bufs.push_back(buf);
        index->device_bytes += words * 4;
#endif
        index->view.rcs = ra.rcs; index->view.rcp = ra.rcp; index->view.palin = ra.palin;
    }
    index->view.mode = (uint32_t)mode;
    return MGB_OK;
}

void mgb_index_destroy(mgb_index_t *index) { delete index; }

uint64_t mgb_index_num_edges(const mgb_index_t *index) { return index->view.n; }
uint64_t mgb_index_device_bytes(const mgb_index_t *index) { return index->device_bytes; }
uint32_t mgb_index_k(const mgb_index_t *index) { return index->view.k; }

void mgb_config_init(mgb_config_t *c) {
    std::memset(c, 0, sizeof(*c));
    c->num_alternative_paths = 1;
    c->max_num_seeds_per_locus = UINT64_MAX;
    c->min_cell_score = INT32_MIN + 100;
    c->min_path_score = 0;
    c->xdrop = INT32_MAX;
    c->max_nodes_per_seq_char = 1.7976931348623157e308;
    c->max_ram_per_alignment = 1.7976931348623157e308;
    c->gap_opening_penalty = -5; c->gap_extension_penalty = -2;
    c->forward_and_reverse_complement = 1; c->global_xdrop = 1; c->allow_left_trim = 1;
    // dna_scoring_matrix(2, -1, -2) (aligner_config.cpp:164-183)
    std::memset(c->score_matrix, -2, sizeof(c->score_matrix));
    c->score_matrix['A']['G'] = c->score_matrix['G']['A'] = -1;
    c->score_matrix['C']['T'] = c->score_matrix['T']['C'] = -1;
    for (const char *p = "ACGT"; *p; ++p) c->score_matrix[(int)*p][(int)*p] = 2;
}

void mgb_config_init_cli(mgb_config_t *c, uint32_t k, int alphabet) {
    mgb_config_init(c);
    c->min_seed_length = k < 19 ? k : 19;           // cli/align.cpp:42-43
    c->max_seed_length = UINT64_MAX;
    c->max_num_seeds_per_locus = 1000;
    c->xdrop = 27;
    c->min_exact_match = 0.7;
    c->max_nodes_per_seq_char = 5.0;
    c->max_ram_per_alignment = 200;
    c->rel_score_cutoff = 0.95;
    c->gap_opening_penalty = -6; c->gap_extension_penalty = -2;
    c->left_end_bonus = 5; c->right_end_bonus = 5;
    std::memset(c->score_matrix, -3, sizeof(c->score_matrix));
    for (const char *p = "ACGT"; *p; ++p) c->score_matrix[(int)*p][(int)*p] = 2;
    if (alphabet == MGB_ALPHABET_PROTEIN) {              // set_scoring_matrix (aligner_config.cpp:128-163)
        blosum62_matrix(c->score_matrix);
        c->forward_and_reverse_complement = 0;           // no reverse complement (dbg_aligner.cpp:224-229)
    }
}

} // extern "C"

namespace {

struct Batch {                 // device-resident read batch
    uint32_t n_reads = 0; uint64_t total_chars = 0, total_kmers = 0; uint32_t L_max = 0;
    char *seqs = nullptr, *qf = nullptr, *qr = nullptr; uint8_t *cf = nullptr, *cr = nullptr;
    uint64_t *offsets = nullptr, *koff = nullptr, *nodes_f = nullptr, *nodes_r = nullptr;
};

int upload_batch(const mgb_index_t *index, const char *seqs, const uint64_t *offsets, uint32_t n_reads,
                 bool need_rc_nodes, Stream &st, DevBufs &bufs, Batch *b, std::vector<uint64_t> *koff_host,
                 mgb_stats_t *stats) {
    const uint32_t k = index->view.k;
    b->n_reads = n_reads;
    b->total_chars = offsets[n_reads];
    koff_host->assign(n_reads + 1, 0);
    for (uint32_t r = 0; r < n_reads; ++r) {
        if (offsets[r + 1] < offsets[r]) return fail(MGB_ERR_INVALID_ARGUMENT, "offsets must be non-decreasing");
        uint64_t L = offsets[r + 1] - offsets[r];
        if (L > 0x3fffffffull) return fail(MGB_ERR_INVALID_ARGUMENT, "read too long");
        if (L > b->L_max) b->L_max = (uint32_t)L;
        (*koff_host)[r + 1] = (*koff_host)[r] + (L >= k ? L - k + 1 : 0);
    }
    b->total_kmers = (*koff_host)[n_reads];
    int rc;
    const size_t pad = 64;
    if ((rc = bufs.alloc(&b->seqs, b->total_chars + pad))) return rc;
    if ((rc = bufs.alloc(&b->qf, b->total_chars + pad))) return rc;
    if ((rc = bufs.alloc(&b->qr, b->total_chars + pad))) return rc;
    if ((rc = bufs.alloc(&b->cf, b->total_chars + pad))) return rc;
    if ((rc = bufs.alloc(&b->cr, b->total_chars + pad))) return rc;
    if ((rc = bufs.alloc(&b->offsets, (size_t)n_reads + 1))) return rc;
    if ((rc = bufs.alloc(&b->koff, (size_t)n_reads + 1))) return rc;
    if ((rc = bufs.alloc(&b->nodes_f, b->total_kmers + 1))) return rc;
    if (need_rc_nodes && (rc = bufs.alloc(&b->nodes_r, b->total_kmers + 1))) return rc;
    if ((rc = h2d(b->seqs, seqs, b->total_chars, st))) return rc;
    if ((rc = h2d(b->offsets, offsets, ((size_t)n_reads + 1) * 8, st))) return rc;
    if ((rc = h2d(b->koff, koff_host->data(), ((size_t)n_reads + 1) * 8, st))) return rc;
    stats->h2d_bytes += b->total_chars + 2 * ((size_t)n_reads + 1) * 8;
    return 0;
}

int launch_prepare(const mgb_index_t *index, const Batch &b, Stream &st, int num_sms) {
    PrepArgs a { b.seqs, b.offsets, b.n_reads, b.qf, b.qr, b.cf, b.cr, {0} };
    std::memcpy(a.code_of, index->at.code_of, sizeof(a.code_of));
#if defined(MGB_HOST_EMU)
    (void)st; (void)num_sms;
    for (uint32_t r = 0; r < b.n_reads; ++r) prepare_read(a, r);
#else
    if (!b.n_reads) return 0;
    int blocks = std::min<uint64_t>((b.n_reads + 7) / 8, (uint64_t)num_sms * 8);
    kern_dna::k_prepare<<<blocks, 256, 0, st.s>>>(a);
    CUDA_TRY(cudaGetLastError());
#endif
    return 0;
}

int launch_seed(const mgb_index_t *index, const Batch &b, uint32_t n_strands, Stream &st, DevBufs &bufs,
                uint64_t *n_launches = nullptr) {
    SeedArgs a { index->view, b.cf, b.cr, b.offsets, b.koff, b.nodes_f, b.nodes_r, b.n_reads, n_strands, 0, nullptr, nullptr };
    uint64_t items = (uint64_t)b.n_reads * n_strands;
    // first pass (k_premap): rules out the k-mers the suffix-range table has no node for; the node arrays
    // are zero-filled, so the second pass only visits what is left
    const bool premap = index->view.sfx_len && index->view.sfx_len <= index->view.k - 1 && b.total_kmers
                        && !index->view.kh_slots               // (the k-mer hash index answers every k-mer with one load)
                        && !std::getenv("MGB_TEST_NOPREMAP");
    if (premap) {
        const size_t words = (size_t)(b.total_kmers >> 5) + b.n_reads + 2;
        int rc;
        if ((rc = bufs.alloc(&a.hint_f, words))) return rc;
        if (n_strands > 1 && (rc = bufs.alloc(&a.hint_r, words))) return rc;
        a.hinted = 1;
    }
    if (n_launches) *n_launches += premap ? 2 : 1;
#if defined(MGB_HOST_EMU)
    (void)st;
    if (premap)
        for (uint64_t it = 0; it < items; ++it) premap_item(a, (uint32_t)(it / n_strands), (uint32_t)(it % n_strands));
    for (uint64_t it = 0; it < items; ++it) seed_item(a, it);
#else
    if (!items) return 0;
    if (premap) {
        const uint64_t pb = std::min<uint64_t>((items + 7) / 8, (uint64_t)index->num_sms * 16);
        CUDA_TRY(index->view.wide ? kern_any::launch_premap((unsigned)pb, st.s, a) : kern_dna::launch_premap((unsigned)pb, st.s, a));
    }
    // one strand per lane, 128 strands per block; enough blocks to fill the machine, grid-stride beyond that
    uint64_t blocks = std::min<uint64_t>((items + 127) / 128, (uint64_t)index->num_sms * 16);
    CUDA_TRY(index->view.wide ? kern_any::launch_seed((unsigned)blocks, st.s, a) : kern_dna::launch_seed((unsigned)blocks, st.s, a));
#endif
    return 0;
}

} // namespace

extern "C" {

int mgb_map_to_nodes(const mgb_index_t *index, const char *seqs, const uint64_t *offsets,
                     uint32_t n_seqs, uint64_t *out_nodes) {
    if (!index || !seqs || !offsets || !out_nodes) return fail(MGB_ERR_INVALID_ARGUMENT, "null argument");
    Stream st;
    int rc = 0;
    struct WsGuard { const mgb_index_t *ix; Workspace *w; ~WsGuard() { ix->ws_release(w); } } wsg{ index, index->ws_acquire() };
#if !defined(MGB_HOST_EMU)
    CUDA_TRY(cudaSetDevice(index->device));
    if ((rc = wsg.w->ready())) return rc;
    st.s = wsg.w->stream;
#endif
    {
        DevBufs bufs(st, wsg.w, 0);
        Batch b; std::vector<uint64_t> koff; mgb_stats_t stats; std::memset(&stats, 0, sizeof(stats));
        // PRIMARY graph: CanonicalDBG::map_to_nodes_sequentially (canonical_dbg.cpp:55-146) = both strands mapped in
        // the stored graph; a k-mer missing on its own strand takes the id of its reverse complement + n
        const bool primary = index->view.mode == 2;
        std::vector<uint64_t> rev;
        rc = upload_batch(index, seqs, offsets, n_seqs, primary, st, bufs, &b, &koff, &stats);
        if (!rc) rc = dev_zero(b.nodes_f, (b.total_kmers + 1) * 8, st);
        if (!rc && primary) rc = dev_zero(b.nodes_r, (b.total_kmers + 1) * 8, st);
        if (!rc) rc = launch_prepare(index, b, st, index->num_sms);
        if (!rc) rc = launch_seed(index, b, primary ? 2 : 1, st, bufs);
        if (!rc) rc = d2h(out_nodes, b.nodes_f, b.total_kmers * 8, st);
        if (!rc && primary) { rev.resize(b.total_kmers + 1); rc = d2h(rev.data(), b.nodes_r, b.total_kmers * 8, st); }
#if !defined(MGB_HOST_EMU)
        if (!rc) { cudaError_t e = cudaStreamSynchronize(st.s); if (e != cudaSuccess) rc = fail(MGB_ERR_CUDA, cudaGetErrorString(e)); }
#endif
        if (!rc && primary) {
            const uint64_t n = index->view.n;
            for (uint32_t r = 0; r < n_seqs; ++r) {
                const uint64_t lo = koff[r], nk = koff[r + 1] - koff[r];
                for (uint64_t i = 0; i < nk; ++i) {
                    const uint64_t rv = rev[lo + nk - 1 - i];
                    if (!out_nodes[lo + i] && rv) out_nodes[lo + i] = rv + n;
                }
            }
        }
    }
#if !defined(MGB_HOST_EMU)
    cudaStreamSynchronize(st.s);
#endif
    return rc;
}

} // extern "C"

// One contiguous range of reads [lo, lo + n) of a batch, processed start to finish on its own
// stream: upload, prepare + seed, align passes, download, materialisation. Pieces of one batch are
// independent (IDBGAligner::align_batch treats every query separately, dbg_aligner.cpp:251-355), so
// the download and unpacking of one piece overlap the kernels of the next.
struct Piece {
    uint32_t lo = 0, n_reads = 0;
    // views into the arrays of the enclosing mgb_results: entries of reads [lo, lo + n_reads), and a
    // region of n_reads * num_alternative_paths records that receives n_alns of them (piece-local
    // indices in `first`)
    uint64_t *first = nullptr;
    uint32_t *count = nullptr;
    mgb_alignment_t *alns = nullptr;
    uint64_t n_alns = 0;
    std::vector<HostBuf> heaps;
    std::vector<uint64_t> heap_bytes;
    uint64_t *src_off = nullptr; uint32_t *src_heap = nullptr;   // views into the enclosing mgb_results
    mgb_stats_t stats;
    int rc = 0;
    std::string err;
    ~Piece() { for (HostBuf b : heaps) hostbuf_release(b); }
};

static int align_range(const mgb_index_t *index, const DevConfig &dcfg, const char *all_seqs,
                       const uint64_t *all_offsets, uint32_t lo, uint32_t n_reads, Piece *res) {
    // MGB_DEBUG=1: wall-clock split of the call on stderr
    const bool dbg_time = std::getenv("MGB_DEBUG") != nullptr;
    auto t_last = std::chrono::steady_clock::now();
    auto tick = [&](const char *what) {
        if (!dbg_time) return;
        auto now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[mgb] reads %u+%u %-28s %8.2f ms\n", lo, n_reads, what,
                     std::chrono::duration<double, std::milli>(now - t_last).count());
        t_last = now;
    };
    int rc = 0;
    std::memset(&res->stats, 0, sizeof(res->stats));
    res->lo = lo; res->n_reads = n_reads;
    // offsets relative to the first character of the piece
    std::vector<uint64_t> offs_local;
    const uint64_t *offsets = all_offsets + lo;
    const char *seqs = all_seqs ? all_seqs + all_offsets[lo] : nullptr;
    if (all_offsets[lo]) {
        offs_local.resize((size_t)n_reads + 1);
        for (uint32_t r = 0; r <= n_reads; ++r) offs_local[r] = all_offsets[lo + r] - all_offsets[lo];
        offsets = offs_local.data();
    }

    Stream st;
    struct WsGuard { const mgb_index_t *ix; Workspace *w; ~WsGuard() { ix->ws_release(w); } } wsg{ index, index->ws_acquire() };
#if !defined(MGB_HOST_EMU)
    CUDA_TRY(cudaSetDevice(index->device));
    if ((rc = wsg.w->ready())) return rc;
    st.s = wsg.w->stream;
    cudaEvent_t *ev = wsg.w->ev;
#endif
    HostBuf hdr_buf = hostbuf_acquire(((size_t)n_reads + 1) * sizeof(ReadHdr));
    if (!hdr_buf.p) return fail(MGB_ERR_CUDA, "host buffer allocation failed");
    ReadHdr *hdr_host = (ReadHdr*)hdr_buf.p;
    struct HdrGuard { HostBuf b; ~HdrGuard() { hostbuf_release(b); } } hdr_guard{ hdr_buf };
    std::vector<uint32_t> read_heap(n_reads, 0);   // which pass produced the read's alignments
    {
        DevBufs bufs(st, wsg.w, 0);
        Batch b; std::vector<uint64_t> koff;
        const bool both = dcfg.forward_and_reverse_complement;
        const bool map_nodes = dcfg.max_seed_length >= index->view.k;
#if !defined(MGB_HOST_EMU)
        cudaEventRecord(ev[0], st.s);
#endif
        tick("setup");
        rc = upload_batch(index, seqs, offsets, n_reads, both, st, bufs, &b, &koff, &res->stats);
        tick("upload_batch (host part)");
        if (!rc) rc = dev_zero(b.nodes_f, (b.total_kmers + 1) * 8, st);
        if (!rc && both) rc = dev_zero(b.nodes_r, (b.total_kmers + 1) * 8, st);
#if !defined(MGB_HOST_EMU)
        cudaEventRecord(ev[1], st.s);
#endif
        if (!rc) rc = launch_prepare(index, b, st, index->num_sms);
        res->stats.kernel_launches += 1;
        if (!rc && map_nodes) rc = launch_seed(index, b, both ? 2 : 1, st, bufs, &res->stats.kernel_launches);
        // sub-k seeding (min_seed_length < k): the independent index_range lookups run GPU-wide first
        SubkArgs sk;
        std::memset(&sk, 0, sizeof(sk));
        const bool subk = !rc && map_nodes && dcfg.min_seed_length < index->view.k && b.total_chars
                          && !std::getenv("MGB_TEST_NOSUBK");
        if (subk) {
            sk.ix = index->view; sk.cf = b.cf; sk.cr = b.cr; sk.offsets = b.offsets; sk.koff = b.koff;
            sk.nodes_f = b.nodes_f; sk.nodes_r = b.nodes_r; sk.n_reads = n_reads; sk.n_strands = both ? 2 : 1;
            sk.min_seed_length = dcfg.min_seed_length;
            sk.max_len = dcfg.max_seed_length < index->view.k - 1 ? dcfg.max_seed_length : index->view.k - 1;
            const size_t nc = b.total_chars + 64;
            if (!rc) rc = bufs.alloc(&sk.first_f, nc);
            if (!rc) rc = bufs.alloc(&sk.last_f, nc);
            if (!rc) rc = bufs.alloc(&sk.len_f, nc);
            if (!rc && both) rc = bufs.alloc(&sk.first_r, nc);
            if (!rc && both) rc = bufs.alloc(&sk.last_r, nc);
            if (!rc && both) rc = bufs.alloc(&sk.len_r, nc);
            const uint32_t chunks = (b.L_max + kSubkChunk - 1) / kSubkChunk;
#if defined(MGB_HOST_EMU)
            for (uint32_t r = 0; r < n_reads && !rc; ++r)
                for (uint32_t s2 = 0; s2 < sk.n_strands; ++s2)
                    for (uint32_t c = 0; c < chunks; ++c) subk_item(sk, r, s2, c);
#else
            if (!rc) {
                // one warp per 32 query positions of a strand, 4 warps per block
                const uint64_t items = (uint64_t)n_reads * sk.n_strands * ((chunks * kSubkChunk + 31) / 32);
                const uint64_t blocks = std::min<uint64_t>((items + 3) / 4, (uint64_t)index->num_sms * 16);
                CUDA_TRY(index->view.wide ? kern_any::launch_subk((unsigned)blocks, st.s, sk, chunks)
                                          : kern_dna::launch_subk((unsigned)blocks, st.s, sk, chunks));
            }
#endif
            res->stats.kernel_launches += 1;
        }
#if !defined(MGB_HOST_EMU)
        cudaEventRecord(ev[2], st.s);
#endif
        ReadHdr *d_hdr = nullptr;
        if (!rc) rc = bufs.alloc(&d_hdr, (size_t)n_reads + 1);

        // align passes: all reads with scale 1, overflowed reads again with larger arenas
        std::vector<uint32_t> list(n_reads);
        for (uint32_t r = 0; r < n_reads; ++r) list[r] = r;
        uint32_t scale = 1;
        float align_ms = 0, d2h_ms = 0; (void)align_ms; (void)d2h_ms;
        for (int pass = 0; !rc && !list.empty(); ++pass, scale *= 4) {
            if (pass == 6) { rc = fail(MGB_ERR_OVERFLOW, "a read exceeded the largest per-read work arena"); break; }
            Caps caps = choose_caps(b.L_max, dcfg, index->view.k, scale);
            size_t stride = arena_bytes(caps);
            // on-chip working set per lane group: the column buffers hold the widest column the x-drop band
            // allows (choose_caps' estimate; at least 64 cells, the register path of the extender needs 40) of
            // reads up to 248 bp; wider columns spill to the arena scratch
            uint64_t band_est = (uint64_t)b.L_max + 9;
            if (dcfg.gap_ext < 0 && dcfg.xdrop < (1 << 20))
                band_est = std::min<uint64_t>(band_est, 2ull * (dcfg.xdrop / (-dcfg.gap_ext)) + 16 + 8);
            const int bmax = band_est <= 256 ? std::max(64, (int)((band_est + 31) & ~31ull)) : 256;
            const int lq = b.L_max + 1 <= 512 ? (int)((b.L_max + 1 + 15) & ~15u) : 0;
            int hcap = 16;
            // test knobs: force the spill paths (arena scratch, queue migration, unstaged query)
            int bmax_v = bmax, lq_v = lq;
            if (const char *e = std::getenv("MGB_TEST_BMAX")) bmax_v = std::atoi(e);
            if (const char *e = std::getenv("MGB_TEST_LQ")) lq_v = std::atoi(e);
            if (const char *e = std::getenv("MGB_TEST_HCAP")) hcap = std::atoi(e);
            SmemLayout slay;
            const size_t smem_per_warp = slay.carve(bmax_v, lq_v, hcap, !index->view.wide);
            DevBufs pass_bufs(st, wsg.w, 32);          // slot 32: the arena
#if defined(MGB_HOST_EMU)
            uint32_t n_warps = 1;
#else
            int blocks_per_sm = 0;
            const size_t smem_block = smem_per_warp * kern_dna::kGroupsPerBlock;
            {
                // driver queries serialise against running work: ask once per shared-memory size
                static std::mutex occ_mu;
                static std::map<std::pair<int, size_t>, int> occ_cache;
                std::lock_guard<std::mutex> lk(occ_mu);
                const int variant = index->view.wide ? 1 : index->view.mode != 0 ? 2 : 0;    // which k_align
                auto key = std::make_pair(index->device * 4 + variant, smem_block);
                auto it = occ_cache.find(key);
                if (it == occ_cache.end()) {
                    // the shared memory limit is a high-water mark per device and kernel: keep the largest
                    size_t mx = smem_block;
                    for (auto &kv : occ_cache) if (kv.first.first == key.first) mx = std::max(mx, kv.first.second);
                    CUDA_TRY(variant == 1 ? kern_any::align_occupancy(mx, smem_block, &blocks_per_sm)
                           : variant == 2 ? kern_canon::align_occupancy(mx, smem_block, &blocks_per_sm)
                                          : kern_dna::align_occupancy(mx, smem_block, &blocks_per_sm));
                    occ_cache[key] = blocks_per_sm;
                } else blocks_per_sm = it->second;
            }
            if (blocks_per_sm < 1) blocks_per_sm = 1;
            const uint64_t gpb = kern_dna::kGroupsPerBlock;       // lane groups (reads in flight) per block
            uint64_t n_warps64 = (uint64_t)index->num_sms * blocks_per_sm * gpb;
            if (stride * n_warps64 > pass_bufs.capacity_of_next()) {
                size_t free_b = 0, total_b = 0;
                cudaMemGetInfo(&free_b, &total_b);
                uint64_t mem_warps = (uint64_t)(free_b * 0.6 + pass_bufs.capacity_of_next()) / (stride ? stride : 1);
                if (mem_warps < gpb) mem_warps = gpb;
                if (n_warps64 > mem_warps) n_warps64 = mem_warps / gpb * gpb;
            }
            if (n_warps64 > ((uint64_t)list.size() + gpb - 1) / gpb * gpb) n_warps64 = ((uint64_t)list.size() + gpb - 1) / gpb * gpb;
            uint32_t n_warps = (uint32_t)n_warps64;
#endif
            // output heap: generous first guess, doubled on retry passes
            uint64_t per_read = sizeof(OutAln) + 9ull * (b.L_max + 64 + index->view.k) + 4ull * 64;
            uint64_t heap_cap = (uint64_t)list.size() * per_read * dcfg.num_alternative_paths * scale + 4096;
            char *d_arena = nullptr, *d_heap = nullptr; uint32_t *d_list = nullptr;
            unsigned long long *d_used = nullptr; unsigned int *d_next = nullptr;
            if ((rc = pass_bufs.alloc(&d_arena, stride * n_warps))) break;
            if ((rc = pass_bufs.alloc(&d_heap, heap_cap))) break;
            if ((rc = pass_bufs.alloc(&d_list, list.size()))) break;
            if ((rc = pass_bufs.alloc(&d_used, 1))) break;
            if ((rc = pass_bufs.alloc(&d_next, 1))) break;
            tick("pass: arena/heap alloc");
            if ((rc = h2d(d_list, list.data(), list.size() * 4, st))) break;
            if ((rc = dev_zero(d_used, 8, st))) break;
            if ((rc = dev_zero(d_next, 4, st))) break;
            AlignArgs a;
            a.ix = index->view; a.cfg = dcfg; a.caps = caps; a.lay.carve(caps); a.slay = slay;
            a.use_fast = std::getenv("MGB_TEST_NOFAST") ? 0 : 1;
            a.phase_out = nullptr;
#if defined(MGB_PHASE_TIMERS) && !defined(MGB_HOST_EMU)
            unsigned long long *d_phase = nullptr;
            if ((rc = pass_bufs.alloc(&d_phase, 8))) break;
            if ((rc = dev_zero(d_phase, 64, st))) break;
            a.phase_out = d_phase;
#endif
            a.qf = b.qf; a.qr = b.qr; a.cf = b.cf; a.cr = b.cr; a.offsets = b.offsets; a.koff = b.koff;
            a.nodes_f = b.nodes_f; a.nodes_r = b.nodes_r;
            a.sub_first_f = subk ? sk.first_f : nullptr; a.sub_last_f = sk.last_f; a.sub_len_f = subk ? sk.len_f : nullptr;
            a.sub_first_r = sk.first_r; a.sub_last_r = sk.last_r; a.sub_len_r = subk && both ? sk.len_r : nullptr;
            a.read_list = d_list; a.n_list = (uint32_t)list.size();
            a.arena = d_arena; a.arena_stride = stride;
            a.hdr = d_hdr; a.heap = d_heap; a.heap_cap = heap_cap; a.heap_used = d_used; a.next = d_next;
            unsigned long long used = 0;
#if defined(MGB_HOST_EMU)
            init_arena(a, d_arena);
            std::vector<char> smem_emu(smem_per_warp + 64);
            const WarpMem mem_emu { d_arena, &a.lay };
            const WarpSmem sm_emu { smem_emu.data(), &a.slay };
            for (uint32_t t = 0; t < a.n_list; ++t) align_read(a, a.read_list[t], mem_emu, sm_emu, true);
            used = *d_used;
#else
            cudaEventRecord(ev[3], st.s);
            CUDA_TRY(index->view.wide ? kern_any::launch_align(n_warps / (uint32_t)gpb, smem_block, st.s, a)
                   : index->view.mode != 0 ? kern_canon::launch_align(n_warps / (uint32_t)gpb, smem_block, st.s, a)
                                           : kern_dna::launch_align(n_warps / (uint32_t)gpb, smem_block, st.s, a));
            cudaEventRecord(ev[4], st.s);
            if ((rc = d2h(&used, d_used, 8, st))) break;
            CUDA_TRY(cudaStreamSynchronize(st.s));
            float ms = 0; cudaEventElapsedTime(&ms, ev[3], ev[4]); align_ms += ms;
            cudaEventRecord(ev[3], st.s);
#endif
            tick("pass: h2d+seed+align kernels");
            res->stats.kernel_launches += 1;
#if defined(MGB_PHASE_TIMERS) && !defined(MGB_HOST_EMU)
            {
                unsigned long long ph[8];
                cudaMemcpy(ph, a.phase_out, 64, cudaMemcpyDeviceToHost);
                std::fprintf(stderr, "[phase cycles/read] setup %.0f seeds %.0f fwd %.0f backtrack %.0f align_total %.0f; columns/read: register path %.1f, general path %.1f (reads %u)\n",
                             (double)ph[0] / a.n_list, (double)ph[1] / a.n_list, (double)ph[2] / a.n_list,
                             (double)ph[3] / a.n_list, (double)ph[4] / a.n_list, (double)ph[5] / a.n_list,
                             (double)ph[6] / a.n_list, a.n_list);
            }
#endif
            if (used > heap_cap) used = heap_cap;
            HostBuf heap_host = hostbuf_acquire((size_t)used + 16);
            tick("pass: host buffer");
            if (!heap_host.p) { rc = fail(MGB_ERR_CUDA, "host buffer allocation failed"); break; }
            res->heaps.push_back(heap_host);
            res->heap_bytes.push_back(used);
            if (used && (rc = d2h(heap_host.p, d_heap, used, st))) break;
            if ((rc = d2h(hdr_host, d_hdr, (size_t)n_reads * sizeof(ReadHdr), st))) break;
#if !defined(MGB_HOST_EMU)
            CUDA_TRY(cudaStreamSynchronize(st.s));
            cudaEventRecord(ev[4], st.s);
            CUDA_TRY(cudaStreamSynchronize(st.s));
            { float ms = 0; cudaEventElapsedTime(&ms, ev[3], ev[4]); d2h_ms += ms; }
#endif
            res->stats.d2h_bytes += used + (size_t)n_reads * sizeof(ReadHdr);
            tick("pass: d2h");
            // classify this pass; alignments are materialised in read order after the last pass
            std::vector<uint32_t> retry;
            const uint32_t heap_id = (uint32_t)res->heaps.size() - 1;
            for (uint32_t r : list) {
                const ReadHdr &h = hdr_host[r];
                if (h.status == MGB_READ_OVERFLOW) { retry.push_back(r); continue; }
                res->stats.num_seeds += h.stats.num_seeds;
                res->stats.num_extensions += h.stats.num_extensions;
                res->stats.num_explored_nodes += h.stats.num_explored_nodes;
                res->stats.dp_cells += h.stats.dp_cells;
                res->stats.dp_columns += h.stats.dp_columns;
                res->count[r] = h.n_aln;
                res->first[r] = h.heap_off;          // byte offset for now, patched below
                read_heap[r] = heap_id;
            }
            res->stats.num_reads_retried += retry.size();
            list.swap(retry);
            tick("pass: classify");
        }
#if !defined(MGB_HOST_EMU)
        if (!rc) {
            float ms = 0;
            cudaEventElapsedTime(&ms, ev[0], ev[1]); res->stats.h2d_ms = ms;
            cudaEventElapsedTime(&ms, ev[1], ev[2]); res->stats.seed_kernel_ms = ms;
            res->stats.align_kernel_ms = align_ms; res->stats.d2h_ms = d2h_ms;
        }
#endif
    }
#if !defined(MGB_HOST_EMU)
    cudaStreamSynchronize(st.s);
#endif
    if (rc) return rc;
    tick("free device buffers");
    // materialise mgb_alignment_t records in read order
    {
        std::vector<uint64_t> byte_off(res->first, res->first + n_reads);
        for (uint32_t r = 0; r < n_reads; ++r) { res->src_off[r] = byte_off[r]; res->src_heap[r] = read_heap[r]; }
        uint64_t pos = 0;
        for (uint32_t r = 0; r < n_reads; ++r) { res->first[r] = pos; pos += res->count[r]; }
        res->n_alns = pos;
        #pragma omp parallel for schedule(static) num_threads(host_team()) if (n_reads > 20000)
        for (int64_t r = 0; r < (int64_t)n_reads; ++r) {
            const char *p = res->heaps[read_heap[r]].p + byte_off[r];
            for (uint32_t i = 0; i < res->count[r]; ++i) {
                const OutAln *o = (const OutAln*)p;
                mgb_alignment_t al;
                std::memset(&al, 0, sizeof(al));
                al.read_index = lo + (uint32_t)r; al.orientation = (uint8_t)o->orientation; al.score = o->score;
                al.offset = o->offset; al.query_begin = o->query_begin; al.query_len = o->query_len;
                al.num_nodes = o->n_nodes; al.sequence_len = o->seq_len; al.num_cigar_ops = o->n_cigar;
                p += sizeof(OutAln);
                if (dcfg.result_nodes) al.nodes = nullptr;
                else { al.nodes = (const uint64_t*)p; p += 8ull * o->n_nodes; }
                al.cigar = (const uint32_t*)p; p += (4ull * o->n_cigar + 7) & ~7ull;
                al.sequence = p; p += ((uint64_t)o->seq_len + 7) & ~7ull;
                res->alns[res->first[r] + i] = al;
            }
        }
    }
    tick("materialise");
    return MGB_OK;
}

static std::atomic<uint32_t> g_max_pieces{0};      // 0 = automatic
#if defined(MGB_HOST_EMU)
extern "C" { unsigned long long mgb_emu_whole_read_hits = 0; }
#endif

extern "C" {

void mgb_set_pipeline_pieces(uint32_t max_pieces) { g_max_pieces.store(max_pieces); }

void mgb_set_host_threads(int n) {
#if defined(_OPENMP)
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int mgb_config_check(const mgb_index_t *index, const mgb_config_t *config) {
    if (!index || !config) return fail(MGB_ERR_INVALID_ARGUMENT, "null argument");
    DevConfig dcfg;
    std::string err;
    const int rc = lower_config(*config, index->view.k, index->alphabet, &dcfg, &err);
    return rc ? fail(rc, err) : MGB_OK;
}

int mgb_align_batch(const mgb_index_t *index, const mgb_config_t *config, const char *seqs,
                    const uint64_t *offsets, uint32_t n_reads, mgb_results_t **out) {
    if (!index || !config || !offsets || !out || (!seqs && n_reads && offsets[n_reads]))
        return fail(MGB_ERR_INVALID_ARGUMENT, "null argument");
    DevConfig dcfg;
    std::string err;
    int rc = lower_config(*config, index->view.k, index->alphabet, &dcfg, &err);
    if (rc) return fail(rc, err);
    if (index->view.mode != 0) {                 // CANONICAL-mode graph, or a PRIMARY one with CanonicalDBG semantics
        dcfg.canonical = 1;                      // (get_mode() == CANONICAL): both strands, no RCDBG view
        dcfg.forward_and_reverse_complement = 1; // (dbg_aligner.cpp:224-226)
    }


    // pieces of >= 64k reads, at most 8; two host threads keep two pieces in flight
    const uint32_t kMinPiece = 65536, kMaxPieces = 8;
    uint32_t n_pieces = std::min<uint32_t>(kMaxPieces, n_reads / kMinPiece);
#if defined(MGB_HOST_EMU)
    n_pieces = 1;
#endif
    if (uint32_t cap = g_max_pieces.load()) n_pieces = std::min<uint32_t>(std::max<uint32_t>(n_pieces, 1), cap);
    if (const char *e = std::getenv("MGB_TEST_PIECES")) n_pieces = (uint32_t)std::atoi(e);   // test knob
    if (n_pieces < 1) n_pieces = 1;
    if (n_pieces > n_reads) n_pieces = n_reads ? n_reads : 1;

    const bool dbg_time = std::getenv("MGB_DEBUG") != nullptr;
    const auto t_begin = std::chrono::steady_clock::now();
    auto since = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count(); };

    // every piece writes its records straight into the result arrays: a read has at most
    // num_alternative_paths alignments, so piece i owns the record region starting at lo_i * num_alt
    std::unique_ptr<mgb_results> res(new mgb_results());
    std::memset(&res->stats, 0, sizeof(res->stats));
    res->n_reads = n_reads;
    res->first.assign(n_reads, 0);
    res->count.assign(n_reads, 0);
    res->src_off.assign(n_reads, 0);
    res->src_heap.assign(n_reads, 0);
    res->result_nodes = dcfg.result_nodes;
    const uint64_t num_alt = dcfg.num_alternative_paths;
    res->alns_buf = hostbuf_acquire(((size_t)n_reads * num_alt + 1) * sizeof(mgb_alignment_t));
    if (!res->alns_buf.p) return fail(MGB_ERR_CUDA, "host buffer allocation failed");
    res->alns = (mgb_alignment_t*)res->alns_buf.p;

    std::vector<Piece> pieces(n_pieces);
    auto piece_lo = [&](uint32_t i) { return (uint32_t)((uint64_t)n_reads * i / n_pieces); };
    for (uint32_t i = 0; i < n_pieces; ++i) {
        pieces[i].first = res->first.data() + piece_lo(i);
        pieces[i].count = res->count.data() + piece_lo(i);
        pieces[i].alns = res->alns + (uint64_t)piece_lo(i) * num_alt;
        pieces[i].src_off = res->src_off.data() + piece_lo(i);
        pieces[i].src_heap = res->src_heap.data() + piece_lo(i);
    }
    if (n_pieces == 1) {
        rc = align_range(index, dcfg, seqs, offsets, 0, n_reads, &pieces[0]);
        if (rc) return rc;
    } else {
        std::atomic<uint32_t> next{0};
        std::atomic<bool> failed{false};
        auto lane = [&]() {
            for (uint32_t i; !failed.load() && (i = next.fetch_add(1)) < n_pieces; ) {
                Piece &pc = pieces[i];
                pc.rc = align_range(index, dcfg, seqs, offsets, piece_lo(i), piece_lo(i + 1) - piece_lo(i), &pc);
                if (pc.rc) { pc.err = g_err; failed.store(true); }
            }
        };
        // two host threads keep two pieces in flight (MGB_TEST_LANES: measurement knob)
        uint32_t n_lanes = 2;
        if (const char *e = std::getenv("MGB_TEST_LANES")) n_lanes = std::max(1, std::atoi(e));
        n_lanes = std::min(n_lanes, n_pieces);
        std::vector<std::thread> others;
        for (uint32_t t = 1; t < n_lanes; ++t) others.emplace_back(lane);
        lane();
        for (std::thread &t : others) t.join();
        for (Piece &pc : pieces)
            if (pc.rc) return fail(pc.rc, pc.err);
    }
    if (dbg_time) std::fprintf(stderr, "[mgb] %u pieces done at %.2f ms\n", n_pieces, since());

    // close the gaps between the regions (none if every read produced num_alt alignments)
    uint64_t total = 0;
    for (uint32_t i = 0; i < n_pieces; ++i) {
        Piece &pc = pieces[i];
        const uint64_t region = (uint64_t)pc.lo * num_alt;
        if (total != region) {
            std::memmove(res->alns + total, res->alns + region, pc.n_alns * sizeof(mgb_alignment_t));
        }
        if (total) {
            const int64_t n = pc.n_reads;
            #pragma omp parallel for schedule(static) num_threads(host_team()) if (n > 50000)
            for (int64_t r = 0; r < n; ++r) pc.first[r] += total;
        }
        total += pc.n_alns;
        {   // heap ids of the piece become ids in the merged list
            const uint32_t base = (uint32_t)res->heaps.size();
            if (base) for (uint32_t r = 0; r < pc.n_reads; ++r) pc.src_heap[r] += base;
        }
        for (HostBuf b : pc.heaps) res->heaps.push_back(b);
        for (uint64_t b : pc.heap_bytes) res->heap_bytes.push_back(b);
        pc.heaps.clear();
        mgb_stats_t &t = res->stats; const mgb_stats_t &u = pc.stats;
        t.num_seeds += u.num_seeds; t.num_extensions += u.num_extensions;
        t.num_explored_nodes += u.num_explored_nodes; t.dp_cells += u.dp_cells; t.dp_columns += u.dp_columns;
        t.num_reads_retried += u.num_reads_retried; t.seed_kernel_ms += u.seed_kernel_ms;
        t.align_kernel_ms += u.align_kernel_ms; t.h2d_ms += u.h2d_ms; t.d2h_ms += u.d2h_ms;
        t.h2d_bytes += u.h2d_bytes; t.d2h_bytes += u.d2h_bytes; t.kernel_launches += u.kernel_launches;
    }
    res->n_alns = total;
    if (dbg_time) std::fprintf(stderr, "[mgb] merged at %.2f ms\n", since());
    *out = res.release();
    return MGB_OK;
}

// Relocatable form of a result set (what a rank ships to rank 0): header, per-read record counts and
// positions, then the packed records exactly as the kernel wrote them (OutAln | nodes | cigar | sequence).
namespace {
struct ExportHdr { uint64_t magic, n_reads, n_alns, n_heaps, result_nodes, total_bytes; };
const uint64_t kExportMagic = 0x3142474d53455231ull;      // "1RESMGB1"
}
uint64_t mgb_results_export_bytes(const mgb_results_t *r) {
    uint64_t b = sizeof(ExportHdr) + (uint64_t)r->n_reads * (8 + 4 + 4) + 8 * r->heap_bytes.size();
    b = (b + 15) & ~15ull;
    for (uint64_t h : r->heap_bytes) b += (h + 15) & ~15ull;
    return b;
}
int mgb_results_export(const mgb_results_t *r, void *dst, uint64_t capacity) {
    if (!r || !dst) return fail(MGB_ERR_INVALID_ARGUMENT, "null argument");
    const uint64_t total = mgb_results_export_bytes(r);
    if (capacity < total) return fail(MGB_ERR_INVALID_ARGUMENT, "export buffer too small");
    char *p = (char*)dst;
    ExportHdr h { kExportMagic, r->n_reads, r->n_alns, r->heap_bytes.size(), r->result_nodes, total };
    std::memcpy(p, &h, sizeof(h)); p += sizeof(h);
    std::memcpy(p, r->src_off.data(), 8ull * r->n_reads); p += 8ull * r->n_reads;
    std::memcpy(p, r->count.data(), 4ull * r->n_reads); p += 4ull * r->n_reads;
    std::memcpy(p, r->src_heap.data(), 4ull * r->n_reads); p += 4ull * r->n_reads;
    std::memcpy(p, r->heap_bytes.data(), 8 * r->heap_bytes.size()); p += 8 * r->heap_bytes.size();
    p = (char*)dst + (((uint64_t)(p - (char*)dst) + 15) & ~15ull);
    for (size_t i = 0; i < r->heaps.size(); ++i) {
        const uint64_t n = r->heap_bytes[i];
        const int64_t chunks = (int64_t)((n + (1 << 22) - 1) >> 22);
        #pragma omp parallel for schedule(static) num_threads(host_team()) if (chunks > 4)
        for (int64_t c = 0; c < chunks; ++c) {
            const uint64_t lo = (uint64_t)c << 22, len = std::min<uint64_t>(1ull << 22, n - lo);
            std::memcpy(p + lo, r->heaps[i].p + lo, len);
        }
        p += (n + 15) & ~15ull;
    }
    return MGB_OK;
}
int mgb_results_import(const void *blob, uint64_t bytes, uint32_t read_index_base, mgb_results_t **out) {
    if (!blob || !out || bytes < sizeof(ExportHdr)) return fail(MGB_ERR_INVALID_ARGUMENT, "bad export blob");
    ExportHdr h; std::memcpy(&h, blob, sizeof(h));
    if (h.magic != kExportMagic || h.total_bytes > bytes || h.n_reads > 0xffffffffull)
        return fail(MGB_ERR_INVALID_ARGUMENT, "bad export blob");
    std::unique_ptr<mgb_results> res(new mgb_results());
    std::memset(&res->stats, 0, sizeof(res->stats));
    const uint32_t n = (uint32_t)h.n_reads;
    res->n_reads = n; res->n_alns = h.n_alns; res->result_nodes = (uint32_t)h.result_nodes;
    res->first.assign(n, 0); res->count.assign(n, 0); res->src_off.assign(n, 0); res->src_heap.assign(n, 0);
    const char *p = (const char*)blob + sizeof(h);
    std::memcpy(res->src_off.data(), p, 8ull * n); p += 8ull * n;
    std::memcpy(res->count.data(), p, 4ull * n); p += 4ull * n;
    std::memcpy(res->src_heap.data(), p, 4ull * n); p += 4ull * n;
    res->heap_bytes.assign(h.n_heaps, 0);
    std::memcpy(res->heap_bytes.data(), p, 8 * h.n_heaps); p += 8 * h.n_heaps;
    p = (const char*)blob + (((uint64_t)(p - (const char*)blob) + 15) & ~15ull);
    // the records are referenced in place: one host buffer takes over the heaps
    uint64_t heaps_total = 0;
    for (uint64_t b : res->heap_bytes) heaps_total += (b + 15) & ~15ull;
    if ((uint64_t)(p - (const char*)blob) + heaps_total > bytes) return fail(MGB_ERR_INVALID_ARGUMENT, "truncated export blob");
    HostBuf hb = hostbuf_acquire(heaps_total + 16);
    if (!hb.p) return fail(MGB_ERR_CUDA, "host buffer allocation failed");
    res->heaps.push_back(hb);
    std::vector<uint64_t> heap_base(h.n_heaps, 0);
    {
        uint64_t o = 0;
        for (size_t i = 0; i < h.n_heaps; ++i) { heap_base[i] = o; o += (res->heap_bytes[i] + 15) & ~15ull; }
        const int64_t chunks = (int64_t)((heaps_total + (1 << 22) - 1) >> 22);
        #pragma omp parallel for schedule(static) num_threads(host_team()) if (chunks > 4)
        for (int64_t c = 0; c < chunks; ++c) {
            const uint64_t lo = (uint64_t)c << 22, len = std::min<uint64_t>(1ull << 22, heaps_total - lo);
            std::memcpy(hb.p + lo, p + lo, len);
        }
    }
    res->alns_buf = hostbuf_acquire((h.n_alns + 1) * sizeof(mgb_alignment_t));
    if (!res->alns_buf.p) return fail(MGB_ERR_CUDA, "host buffer allocation failed");
    res->alns = (mgb_alignment_t*)res->alns_buf.p;
    uint64_t pos = 0;
    for (uint32_t r = 0; r < n; ++r) { res->first[r] = pos; pos += res->count[r]; }
    if (pos != h.n_alns) return fail(MGB_ERR_INVALID_ARGUMENT, "inconsistent export blob");
    const bool no_nodes = res->result_nodes != 0;
    #pragma omp parallel for schedule(static) num_threads(host_team()) if (n > 20000)
    for (int64_t r = 0; r < (int64_t)n; ++r) {
        if (!res->count[r]) continue;
        const char *q = hb.p + heap_base[res->src_heap[r]] + res->src_off[r];
        for (uint32_t i = 0; i < res->count[r]; ++i) {
            const OutAln *o = (const OutAln*)q;
            mgb_alignment_t al;
            std::memset(&al, 0, sizeof(al));
            al.read_index = read_index_base + (uint32_t)r; al.orientation = (uint8_t)o->orientation; al.score = o->score;
            al.offset = o->offset; al.query_begin = o->query_begin; al.query_len = o->query_len;
            al.num_nodes = o->n_nodes; al.sequence_len = o->seq_len; al.num_cigar_ops = o->n_cigar;
            q += sizeof(OutAln);
            if (no_nodes) al.nodes = nullptr; else { al.nodes = (const uint64_t*)q; q += 8ull * o->n_nodes; }
            al.cigar = (const uint32_t*)q; q += (4ull * o->n_cigar + 7) & ~7ull;
            al.sequence = q; q += ((uint64_t)o->seq_len + 7) & ~7ull;
            res->alns[res->first[r] + i] = al;
        }
    }
    // after the import the records live in heap 0 at the rebased offsets
    for (uint32_t r = 0; r < n; ++r) { res->src_off[r] += heap_base[res->src_heap[r]]; res->src_heap[r] = 0; }
    res->heap_bytes.assign(1, heaps_total);
    *out = res.release();
    return MGB_OK;
}

uint32_t mgb_results_num_reads(const mgb_results_t *r) { return r->n_reads; }
void mgb_results_read_range(const mgb_results_t *r, uint32_t read, uint64_t *first, uint32_t *count) {
    *first = r->first[read]; *count = r->count[read];
}
uint64_t mgb_results_num_alignments(const mgb_results_t *r) { return r->n_alns; }
const mgb_alignment_t* mgb_results_alignments(const mgb_results_t *r) { return r->alns; }
const mgb_stats_t* mgb_results_stats(const mgb_results_t *r) { return &r->stats; }
void mgb_results_free(mgb_results_t *r) { delete r; }

} // extern "C"
