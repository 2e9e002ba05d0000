// C ABI of libsnvknn (include/snvknn.h): index objects, staging of host buffers, kernel dispatch.
#include <algorithm>
#include <map>
#include <mutex>
#include <unordered_map>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "common.cuh"
#include "kernels.cuh"

namespace snv {

static thread_local std::string t_error;
long long g_launch_count = 0;
void set_error(const std::string& msg) { t_error = msg; }

// Measurement hooks are per calling thread (searches on distinct streams from distinct threads do not share them).
// The events belong to the device that is current when they are first used.
thread_local int g_last_hamming_engine = -1;
static thread_local bool g_prof_on = false;
static thread_local cudaEvent_t g_prof_e0 = nullptr, g_prof_e1 = nullptr;
static thread_local int g_prof_dev = -1;
static thread_local bool g_prof_valid = false;
void profile_begin(cudaStream_t stream)
{
    if (!g_prof_on) return;
    int dev = -1;
    cudaGetDevice(&dev);
    if (g_prof_e0 && dev != g_prof_dev) {  // the search moved to another device: events are per device
        cudaEventDestroy(g_prof_e0);
        cudaEventDestroy(g_prof_e1);
        g_prof_e0 = g_prof_e1 = nullptr;
        g_prof_valid = false;
    }
    if (!g_prof_e0) { cudaEventCreate(&g_prof_e0); cudaEventCreate(&g_prof_e1); g_prof_dev = dev; }
    cudaEventRecord(g_prof_e0, stream);
}
void profile_end(cudaStream_t stream)
{
    if (!g_prof_on || !g_prof_e0) return;
    cudaEventRecord(g_prof_e1, stream);
    g_prof_valid = true;
}

namespace {

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; }
        ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// Device blocks of freed indexes / outgrown workspaces are kept (per device, up to SNV_CACHE_MB, default 1024 MiB) and
// handed to the next allocation of a similar size: the reference builds and drops one faiss index per window
// (src/dataset/rag_train_dataset.py:107-131, scripts_test/batch_test_faiss_l2.py:104-111), and a raw cudaFree /
// cudaMalloc pair unmaps and re-maps the pages every time (measured on B200: 2-130 ms spikes per build + search call).
// dev_free has cudaFree's ordering: it waits for the device before the block can be reused.
struct DevCache {
    std::mutex mu;
    std::multimap<size_t, void*> blocks[64];   // free blocks by capacity
    std::unordered_map<void*, size_t> cap_of;  // every block handed out or cached
    size_t cached[64] = {};
    size_t limit = [] {
        const char* e = getenv("SNV_CACHE_MB");
        return (size_t)(e ? std::max(0L, atol(e)) : 1024L) << 20;
    }();
};
DevCache& dev_cache()
{
    static DevCache* c = new DevCache();  // never destroyed: the CUDA context may be gone at exit
    return *c;
}

static size_t round_block(size_t bytes)
{
    const size_t q = bytes <= ((size_t)1 << 20) ? (size_t)4096 : (size_t)1 << 18;
    return (bytes + q - 1) / q * q;
}

static void dev_cache_flush(int dev)
{
    DevCache& c = dev_cache();
    for (auto& kv : c.blocks[dev]) { cudaFree(kv.second); c.cap_of.erase(kv.second); }
    c.blocks[dev].clear();
    c.cached[dev] = 0;
}

cudaError_t dev_alloc(void** out, size_t bytes)
{
    int dev = 0;
    cudaGetDevice(&dev);
    DevCache& c = dev_cache();
    const size_t want = round_block(bytes ? bytes : 1);
    std::lock_guard<std::mutex> lk(c.mu);
    if (dev >= 0 && dev < 64) {
        auto it = c.blocks[dev].lower_bound(want);
        if (it != c.blocks[dev].end() && it->first <= want + want / 4) {
            *out = it->second;
            c.cached[dev] -= it->first;
            c.blocks[dev].erase(it);
            return cudaSuccess;
        }
    }
    cudaError_t e = cudaMalloc(out, want);
    if (e != cudaSuccess && dev >= 0 && dev < 64 && !c.blocks[dev].empty()) {
        cudaGetLastError();
        dev_cache_flush(dev);
        e = cudaMalloc(out, want);
    }
    if (e == cudaSuccess) c.cap_of[*out] = want;
    return e;
}

void dev_free(void* p)
{
    if (!p) return;
    int dev = 0;
    cudaGetDevice(&dev);
    DevCache& c = dev_cache();
    std::lock_guard<std::mutex> lk(c.mu);
    auto it = c.cap_of.find(p);
    if (it == c.cap_of.end() || dev < 0 || dev >= 64 || it->second > c.limit) {
        if (it != c.cap_of.end()) c.cap_of.erase(it);
        cudaFree(p);
        return;
    }
    cudaDeviceSynchronize();  // nothing in flight may still use the block when it is handed out again
    const size_t cap = it->second;
    while (c.cached[dev] + cap > c.limit && !c.blocks[dev].empty()) {  // make room: drop the largest cached blocks first
        auto big = std::prev(c.blocks[dev].end());
        cudaFree(big->second);
        c.cap_of.erase(big->second);
        c.cached[dev] -= big->first;
        c.blocks[dev].erase(big);
    }
    c.blocks[dev].emplace(cap, p);
    c.cached[dev] += cap;
}

// grow-only device scratch buffer
struct Buf {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes)
    {
        if (bytes <= cap) return SNV_OK;
        if (p) { dev_free(p); p = nullptr; cap = 0; }
        size_t want = bytes + bytes / 4;
        cudaError_t e = dev_alloc(&p, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            want = bytes;
            e = dev_alloc(&p, want);
        }
        if (e != cudaSuccess) {
            p = nullptr;
            set_error(std::string("cudaMalloc(") + std::to_string(bytes) + "): " + cudaGetErrorString(e));
            return SNV_ERR_NOMEM;
        }
        cap = want;
        return SNV_OK;
    }
    void release()
    {
        if (p) dev_free(p);
        p = nullptr;
        cap = 0;
    }
};

size_t dtype_row_bytes(int dtype, int64_t d, int stride)
{
    switch (dtype) {
        case SNV_DT_U8: return (size_t)d;
        case SNV_DT_F32: return (size_t)d * 4;
        case SNV_DT_PACKED_U32: return (size_t)stride * 4;
        case SNV_DT_PACKED_U8: return (size_t)((d + 7) / 8);
        case SNV_DT_I64_TOKENS: return (size_t)d * 8;
        case SNV_DT_PACKED_U32_DENSE: return (size_t)((d + 31) / 32) * 4;
        default: return 0;
    }
}

int bucket_stride(int64_t words)
{
    static const int buckets[] = {4, 8, 16, 24, 32, 36, 48, 68};
    for (int b : buckets)
        if (words <= b) return b;
    return (int)round_up(words, 4);
}

__global__ void invert_packed_kernel(const uint32_t* __restrict__ x, int64_t rows, int64_t d, int stride,
                                     uint32_t* __restrict__ out)
{
    const int64_t total = rows * stride;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t w = i % stride;
        const int64_t lo = w * 32;
        uint32_t valid = 0;
        if (lo + 32 <= d) valid = 0xFFFFFFFFu;
        else if (lo < d) valid = (1u << (d - lo)) - 1u;
        out[i] = ~x[i] & valid;
    }
}

}  // namespace
}  // namespace snv

using namespace snv;

struct snv_index {
    int kind = 0;
    int64_t d = 0;
    int n_windows = 1;
    int device = 0;
    int l2_mode = SNV_L2_TF32X3;
    int words = 0, stride = 0;  // HAMMING packed geometry
    int kp = 0;                 // L2 operand depth
    int64_t ntotal = 0, cap = 0;
    uint32_t* panel = nullptr;  // HAMMING [W][cap][stride]
    float* rows = nullptr;      // L2 [W][cap][d]
    float* ops = nullptr;       // L2 [W][cap][kp]
    float* norms = nullptr;     // L2 [W][cap]
    float* mean = nullptr;      // L2 + SNV_L2_CENTER: [W][d] column means of the first add
    Buf ws_in, ws_q, ws_mask, ws_min, ws_partial, ws_di, ws_df, ws_i, ws_qops, ws_qnorm, ws_misc;
    // host-buffer pipeline
    static constexpr int kPipeStreams = 3;
    bool pipe_ready = false;
    cudaStream_t pipe_stream[kPipeStreams] = {};
    cudaEvent_t pipe_done[kPipeStreams] = {};
    cudaEvent_t pipe_start = nullptr;
};

extern "C" {

const char* snv_last_error(void) { return t_error.c_str(); }
int snv_version(void) { return SNVKNN_VERSION; }
int64_t snv_launch_count(void) { return g_launch_count; }

int snv_profile_enable(int on)
{
    g_prof_on = on != 0;
    g_prof_valid = false;
    return SNV_OK;
}

int snv_last_hamming_engine(void) { return g_last_hamming_engine; }

static int hamming_chunk_bounds(const HammingSearchParams& p, bool host_io, std::vector<int>& bounds);

int snv_debug_hamming_plan(int n_windows, int nq, int64_t n, int d, int k, int32_t* plan_out, int64_t* items_out,
                           int64_t cap, int64_t* n_items)
{
    if (n_windows < 0 || nq < 0 || n < 0 || d < 1 || k < 1 || !plan_out || !n_items || (cap > 0 && !items_out)) {
        set_error("snv_debug_hamming_plan: bad arguments");
        return SNV_ERR_INVALID;
    }
    HammingSearchParams p{};
    p.d = d;
    p.words = (d + 31) / 32;
    p.stride = snv_packed_stride(d);
    p.n = n;
    p.nq = nq;
    p.nw = n_windows;
    p.k = k;
    HammingTcPlan plan;
    const size_t ws = hamming_tc_plan(p, plan);
    if (ws == (size_t)-1) return SNV_ERR_INVALID;
    const int32_t f[12] = {plan.engine, plan.kt, plan.kblocks, plan.qtiles, plan.n_tiles, plan.nsplit, plan.tiles_per_split,
                           plan.idx_bits, plan.tail_items, plan.tail_split, plan.tail_tiles, (int32_t)std::min<size_t>(ws >> 10, 0x7fffffff)};
    for (int i = 0; i < 12; ++i) plan_out[i] = f[i];
    *n_items = hamming_tc_debug_items(p, plan, items_out, cap);
    return SNV_OK;
}

int snv_debug_tc_codes(int fp4, uint32_t q_word, uint32_t mask_word, uint32_t panel_word, uint8_t* q_codes, uint8_t* panel_codes)
{
    if (!q_codes || !panel_codes) { set_error("snv_debug_tc_codes: null output"); return SNV_ERR_INVALID; }
    hamming_tc_debug_codes(fp4, q_word, mask_word, panel_word, q_codes, panel_codes);
    return SNV_OK;
}

int snv_debug_hamming_chunks(int n_windows, int nq, int64_t n, int d, int k, int host_io, int32_t* bounds_out, int cap,
                             int* n_bounds)
{
    if (n_windows < 1 || nq < 1 || n < 0 || d < 1 || k < 1 || !n_bounds || (cap > 0 && !bounds_out)) {
        set_error("snv_debug_hamming_chunks: bad arguments");
        return SNV_ERR_INVALID;
    }
    HammingSearchParams p{};
    p.d = d;
    p.words = (d + 31) / 32;
    p.stride = snv_packed_stride(d);
    p.n = n;
    p.nq = nq;
    p.nw = n_windows;
    p.k = k;
    std::vector<int> bounds;
    const int rc = hamming_chunk_bounds(p, host_io != 0, bounds);
    if (rc) return rc;
    *n_bounds = (int)bounds.size();
    for (int i = 0; i < (int)bounds.size() && i < cap; ++i) bounds_out[i] = bounds[i];
    return SNV_OK;
}

int snv_profile_last_ms(float* ms)
{
    if (!ms) { set_error("snv_profile_last_ms: null"); return SNV_ERR_INVALID; }
    if (!g_prof_valid) { set_error("snv_profile_last_ms: no profiled launch"); return SNV_ERR_INVALID; }
    SNV_CUDA_CHECK(cudaEventSynchronize(g_prof_e1));
    SNV_CUDA_CHECK(cudaEventElapsedTime(ms, g_prof_e0, g_prof_e1));
    return SNV_OK;
}

int snv_device_count(int* count)
{
    if (!count) { set_error("snv_device_count: null"); return SNV_ERR_INVALID; }
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) {
        cudaGetLastError();
        *count = 0;
        return SNV_OK;
    }
    if (e != cudaSuccess) {
        set_error(std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
        *count = 0;
        return SNV_ERR_CUDA;
    }
    *count = n;
    return SNV_OK;
}

int64_t snv_packed_words(int64_t d) { return d <= 0 ? 0 : (d + 31) / 32; }
int64_t snv_packed_stride(int64_t d) { return d <= 0 ? 0 : bucket_stride((d + 31) / 32); }

int snv_index_create(int kind, int64_t d, int n_windows, int device, int l2_mode, snv_index** out)
{
    if (!out) { set_error("snv_index_create: out is null"); return SNV_ERR_INVALID; }
    *out = nullptr;
    if (kind != SNV_KIND_HAMMING && kind != SNV_KIND_L2) { set_error("snv_index_create: bad kind"); return SNV_ERR_INVALID; }
    if (d <= 0) { set_error("snv_index_create: d must be positive"); return SNV_ERR_INVALID; }
    if (n_windows < 1) { set_error("snv_index_create: n_windows must be >= 1"); return SNV_ERR_INVALID; }
    if (kind == SNV_KIND_L2 && (l2_mode & 0xF) != SNV_L2_TF32 && (l2_mode & 0xF) != SNV_L2_TF32X3) { set_error("snv_index_create: bad l2_mode"); return SNV_ERR_INVALID; }
    if (kind == SNV_KIND_L2 && (l2_mode & ~0x3F)) { set_error("snv_index_create: bad l2_mode flags"); return SNV_ERR_INVALID; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error("snv_index_create: no CUDA device (this engine has no CPU fallback)");
        return SNV_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) { set_error("snv_index_create: bad device ordinal"); return SNV_ERR_INVALID; }
    snv_index* idx = new (std::nothrow) snv_index();
    if (!idx) { set_error("snv_index_create: out of host memory"); return SNV_ERR_NOMEM; }
    idx->kind = kind;
    idx->d = d;
    idx->n_windows = n_windows;
    idx->device = device;
    idx->l2_mode = l2_mode;
    if (kind == SNV_KIND_HAMMING) {
        idx->words = (int)snv_packed_words(d);
        idx->stride = (int)snv_packed_stride(d);
    } else {
        idx->kp = l2_operand_depth(d, l2_mode);
    }
    *out = idx;
    return SNV_OK;
}

void snv_index_free(snv_index* idx)
{
    if (!idx) return;
    DeviceGuard g(idx->device);
    cudaDeviceSynchronize();
    if (idx->panel) dev_free(idx->panel);
    if (idx->rows) dev_free(idx->rows);
    if (idx->ops) dev_free(idx->ops);
    if (idx->norms) dev_free(idx->norms);
    if (idx->mean) dev_free(idx->mean);
    Buf* bufs[] = {&idx->ws_in, &idx->ws_q, &idx->ws_mask, &idx->ws_min, &idx->ws_partial, &idx->ws_di,
                   &idx->ws_df, &idx->ws_i, &idx->ws_qops, &idx->ws_qnorm, &idx->ws_misc};
    for (Buf* b : bufs) b->release();
    if (idx->pipe_ready) {
        for (int i = 0; i < snv_index::kPipeStreams; ++i) {
            cudaStreamDestroy(idx->pipe_stream[i]);
            cudaEventDestroy(idx->pipe_done[i]);
        }
        cudaEventDestroy(idx->pipe_start);
    }
    delete idx;
}

int64_t snv_index_ntotal(const snv_index* idx) { return idx ? idx->ntotal : -1; }
int64_t snv_index_d(const snv_index* idx) { return idx ? idx->d : -1; }
int snv_index_kind(const snv_index* idx) { return idx ? idx->kind : -1; }
int snv_index_n_windows(const snv_index* idx) { return idx ? idx->n_windows : -1; }
int snv_index_device(const snv_index* idx) { return idx ? idx->device : -1; }

int snv_index_reset(snv_index* idx)
{
    if (!idx) { set_error("snv_index_reset: null index"); return SNV_ERR_INVALID; }
    idx->ntotal = 0;
    if (idx->mean) { dev_free(idx->mean); idx->mean = nullptr; }
    if (idx->l2_mode & SNV_L2_CENTER_AUTO) idx->l2_mode &= ~SNV_L2_CENTER;  // decided again on the next first add
    return SNV_OK;
}

// re-layout [W][cap][row] -> [W][new_cap][row]
// Re-lays out up to three [W][cap][row] arrays for a larger capacity, all or nothing: every new array is allocated before
// anything is copied or freed, so a failed allocation leaves the index exactly as it was (pointers and cap unchanged).
struct GrowSpec {
    void** arr;
    size_t row_bytes;
};
static int grow_arrays(const GrowSpec* specs, int n_specs, int W, int64_t ntotal, int64_t old_cap, int64_t new_cap,
                       cudaStream_t stream)
{
    void* fresh[3] = {nullptr, nullptr, nullptr};
    for (int i = 0; i < n_specs; ++i) {
        cudaError_t e = dev_alloc(&fresh[i], (size_t)W * new_cap * specs[i].row_bytes);
        if (e != cudaSuccess) {
            for (int j = 0; j < i; ++j) dev_free(fresh[j]);
            set_error(std::string("cudaMalloc(panel): ") + cudaGetErrorString(e));
            return SNV_ERR_NOMEM;
        }
    }
    cudaError_t e = cudaSuccess;
    if (ntotal > 0) {
        for (int i = 0; i < n_specs && e == cudaSuccess; ++i)
            if (*specs[i].arr)
                e = cudaMemcpy2DAsync(fresh[i], (size_t)new_cap * specs[i].row_bytes, *specs[i].arr, (size_t)old_cap * specs[i].row_bytes,
                                      (size_t)ntotal * specs[i].row_bytes, W, cudaMemcpyDeviceToDevice, stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    }
    if (e != cudaSuccess) {
        for (int i = 0; i < n_specs; ++i) dev_free(fresh[i]);
        set_error(std::string("index growth copy: ") + cudaGetErrorString(e));
        return SNV_ERR_CUDA;
    }
    for (int i = 0; i < n_specs; ++i) {
        if (*specs[i].arr) dev_free(*specs[i].arr);
        *specs[i].arr = fresh[i];
    }
    return SNV_OK;
}

int snv_index_add(snv_index* idx, const void* x, int64_t n, int dtype, unsigned flags, void* stream_)
{
    if (!idx) { set_error("snv_index_add: null index"); return SNV_ERR_INVALID; }
    if (n < 0) { set_error("snv_index_add: negative n"); return SNV_ERR_INVALID; }
    if (n == 0) return SNV_OK;
    if (!x) { set_error("snv_index_add: x is null"); return SNV_ERR_INVALID; }
    const bool hamming = idx->kind == SNV_KIND_HAMMING;
    if (hamming) {
        if (dtype != SNV_DT_U8 && dtype != SNV_DT_F32 && dtype != SNV_DT_PACKED_U32 && dtype != SNV_DT_PACKED_U32_DENSE &&
            dtype != SNV_DT_PACKED_U8 && dtype != SNV_DT_I64_TOKENS) {
            set_error("snv_index_add: bad dtype for a HAMMING index");
            return SNV_ERR_INVALID;
        }
    } else if (dtype != SNV_DT_F32) {
        set_error("snv_index_add: an L2 index takes float32 rows");
        return SNV_ERR_INVALID;
    }
    DeviceGuard g(idx->device);
    if (!g.ok) { set_error("snv_index_add: cudaSetDevice failed"); return SNV_ERR_CUDA; }
    cudaStream_t stream = (cudaStream_t)stream_;
    const int W = idx->n_windows;
    const bool on_dev = flags & SNV_X_ON_DEVICE;

    // capacity
    const int64_t need = idx->ntotal + n;
    if (need > idx->cap) {
        int64_t new_cap = std::max<int64_t>(need, idx->cap + idx->cap / 2);
        int rc;
        if (hamming) {
            const GrowSpec g[1] = {{(void**)&idx->panel, (size_t)idx->stride * 4}};
            rc = grow_arrays(g, 1, W, idx->ntotal, idx->cap, new_cap, stream);
        } else {
            const GrowSpec g[3] = {{(void**)&idx->rows, (size_t)idx->d * 4}, {(void**)&idx->ops, (size_t)idx->kp * 4}, {(void**)&idx->norms, 4}};
            rc = grow_arrays(g, 3, W, idx->ntotal, idx->cap, new_cap, stream);
        }
        if (rc) return rc;
        idx->cap = new_cap;
    }

    const size_t in_row = dtype_row_bytes(dtype, idx->d, idx->stride);
    const size_t in_bytes = (size_t)W * n * in_row;
    const void* xd = x;
    if (hamming && dtype == SNV_DT_PACKED_U32) {
        // straight 2-D copy into the panel
        const size_t rb = (size_t)idx->stride * 4;
        SNV_CUDA_CHECK(cudaMemcpy2DAsync(idx->panel + idx->ntotal * idx->stride, (size_t)idx->cap * rb, x,
                                         (size_t)n * rb, (size_t)n * rb, W,
                                         on_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, stream));
    } else {
        if (!on_dev) {
            int rc = idx->ws_in.reserve(in_bytes);
            if (rc) return rc;
            SNV_CUDA_CHECK(cudaMemcpyAsync(idx->ws_in.p, x, in_bytes, cudaMemcpyHostToDevice, stream));
            xd = idx->ws_in.p;
        }
        if (hamming) {
            const size_t rb = (size_t)idx->stride * 4;
            int rc = idx->ws_q.reserve((size_t)W * n * rb);
            if (rc) return rc;
            rc = pack_launch(xd, (int64_t)W * n, idx->d, dtype, false, idx->stride, (uint32_t*)idx->ws_q.p, nullptr, stream);
            if (rc) return rc;
            SNV_CUDA_CHECK(cudaMemcpy2DAsync(idx->panel + idx->ntotal * idx->stride, (size_t)idx->cap * rb,
                                             idx->ws_q.p, (size_t)n * rb, (size_t)n * rb, W,
                                             cudaMemcpyDeviceToDevice, stream));
        } else {
            const size_t rb = (size_t)idx->d * 4;
            SNV_CUDA_CHECK(cudaMemcpy2DAsync(idx->rows + idx->ntotal * idx->d, (size_t)idx->cap * rb, xd,
                                             (size_t)n * rb, (size_t)n * rb, W, cudaMemcpyDeviceToDevice, stream));
            if ((idx->l2_mode & SNV_L2_CENTER_AUTO) && idx->ntotal == 0 && !(idx->l2_mode & SNV_L2_CENTER)) {
                // decide once, on the first rows: integer-valued vectors (tokens, genotypes) stay as they are - their
                // products are exact - anything else is centred on its column means (one 4-byte read back)
                int rc = idx->ws_misc.reserve(16);
                if (rc) return rc;
                int one = 1, integral = 1;
                SNV_CUDA_CHECK(cudaMemcpyAsync(idx->ws_misc.p, &one, 4, cudaMemcpyHostToDevice, stream));
                rc = l2_integral_check_launch((const float*)xd, (int64_t)W * n * idx->d, (int*)idx->ws_misc.p, stream);
                if (rc) return rc;
                SNV_CUDA_CHECK(cudaMemcpyAsync(&integral, idx->ws_misc.p, 4, cudaMemcpyDeviceToHost, stream));
                SNV_CUDA_CHECK(cudaStreamSynchronize(stream));
                if (!integral) idx->l2_mode |= SNV_L2_CENTER;
            }
            const bool center = idx->l2_mode & SNV_L2_CENTER;
            const int pchunks = l2_prep_chunks(idx->d);
            if (pchunks > 1) { int rc = idx->ws_qnorm.reserve((size_t)n * pchunks * 4); if (rc) return rc; }
            if (center && !idx->mean) {
                if (dev_alloc((void**)&idx->mean, (size_t)W * idx->d * 4) != cudaSuccess) { cudaGetLastError(); set_error("cudaMalloc(mean)"); return SNV_ERR_NOMEM; }
                { int rc = idx->ws_partial.reserve(l2_colmean_scratch_bytes(n, idx->d)); if (rc) return rc; }
                for (int w = 0; w < W; ++w) {
                    int rc = l2_colmean_launch((const float*)xd + (size_t)w * n * idx->d, n, idx->d, idx->mean + (size_t)w * idx->d,
                                               (float*)idx->ws_partial.p, stream);
                    if (rc) return rc;
                }
            }
            for (int w = 0; w < W; ++w) {
                int rc = l2_prep_launch((const float*)xd + (size_t)w * n * idx->d, center ? idx->mean + (size_t)w * idx->d : nullptr, n, idx->d, idx->l2_mode, false,
                                        idx->kp, idx->ops + ((size_t)w * idx->cap + idx->ntotal) * idx->kp,
                                        idx->norms + (size_t)w * idx->cap + idx->ntotal, pchunks > 1 ? (float*)idx->ws_qnorm.p : nullptr, stream);
                if (rc) return rc;
            }
        }
    }
    if (!on_dev) SNV_CUDA_CHECK(cudaStreamSynchronize(stream));
    idx->ntotal = need;
    return SNV_OK;
}

// Stage `rows` rows starting at `src` (host or device, `dtype`) as packed rows.  `raw_dst` /
// `packed_dst` are pre-reserved device slices (nullptr when not needed).  Returns the device
// pointer of the packed rows in *out.
static int stage_packed(snv_index* idx, const void* src, int64_t rows, int dtype, bool on_dev, bool invert,
                        void* raw_dst, uint32_t* packed_dst, const uint32_t** out, uint32_t* obs_out,
                        cudaStream_t stream)
{
    const size_t in_row = dtype_row_bytes(dtype, idx->d, idx->stride);
    const void* xd = src;
    if (!on_dev) {
        SNV_CUDA_CHECK(cudaMemcpyAsync(raw_dst, src, (size_t)rows * in_row, cudaMemcpyHostToDevice, stream));
        xd = raw_dst;
    }
    if (dtype == SNV_DT_PACKED_U32 && !invert) {
        *out = (const uint32_t*)xd;
        return SNV_OK;
    }
    if (dtype == SNV_DT_PACKED_U32) {
        const int block = 256;
        const int grid = (int)std::min<int64_t>(ceil_div(rows * idx->stride, block), (int64_t)kNumSMs * 16);
        invert_packed_kernel<<<grid, block, 0, stream>>>((const uint32_t*)xd, rows, idx->d, idx->stride, packed_dst);
        SNV_LAUNCH_CHECK();
    } else {
        int rc = pack_launch(xd, rows, idx->d, dtype, invert, idx->stride, packed_dst, obs_out, stream);
        if (rc) return rc;
    }
    *out = packed_dst;
    return SNV_OK;
}

// Internal streams for the host-buffer pipeline (H2D of chunk i+1 | scan of chunk i | D2H of chunk i-1)
static int ensure_pipe(snv_index* idx)
{
    if (idx->pipe_ready) return SNV_OK;
    for (int i = 0; i < snv_index::kPipeStreams; ++i) {
        SNV_CUDA_CHECK(cudaStreamCreateWithFlags(&idx->pipe_stream[i], cudaStreamNonBlocking));
        SNV_CUDA_CHECK(cudaEventCreateWithFlags(&idx->pipe_done[i], cudaEventDisableTiming));
    }
    SNV_CUDA_CHECK(cudaEventCreateWithFlags(&idx->pipe_start, cudaEventDisableTiming));
    idx->pipe_ready = true;
    return SNV_OK;
}

// Window-chunk boundaries of one uniform Hamming search (p: the whole call, p.nw windows).  Host buffers are pipelined
// over internal streams chunk by chunk (H2D of chunk i+1 | scan of chunk i | D2H of chunk i-1); device-resident
// searches run as one chunk.  bounds = 0 = b_0 < b_1 < ... = p.nw.
static int hamming_chunk_bounds(const HammingSearchParams& p, bool host_io, std::vector<int>& bounds)
{
    const int nw = p.nw;
    const int64_t nq = p.nq;
    int chunk_w = nw;
    if (host_io && nw >= 8) {
        const int64_t qtiles = ceil_div(nq, 128);
        HammingTcPlan tprobe;
        if (hamming_tc_plan(p, tprobe) == (size_t)-1) return SNV_ERR_UNSUPPORTED;
        if (tprobe.engine) {
            // tensor-core engine: persistent CTAs (or CTA pairs) take (window, query tile [pair]) items round-robin,
            // so a chunk should hold a whole number of items per SM (pair); ~24 chunks keep the pipeline's fill
            // (first H2D) and drain (last D2H) short.  (Chunking device-resident searches to overlap the query
            // expansion with the previous chunk's scan was measured: < 1 %, not worth the extra launches.)
            const int64_t per_w = tprobe.engine >= 4 ? ceil_div(qtiles, 2) : qtiles;
            const int64_t units = tprobe.engine >= 4 ? kNumSMs / 2 : kNumSMs;
            int64_t a = units, b = per_w;
            while (b) { const int64_t t = a % b; a = b; b = t; }
            const int64_t unit = units / a;  // windows per chunk so that items per chunk is a multiple of the units
            int64_t target = 24;
            if (const char* e = getenv("SNV_PIPE_CHUNKS")) target = std::max(1, atoi(e));  // tuning override
            int64_t cw = std::max<int64_t>(unit, nw / target / unit * unit);
            if (cw * per_w < 2 * units) cw = ceil_div(2 * units, per_w);
            chunk_w = (int)std::min<int64_t>(cw, nw);
            if (const char* e = getenv("SNV_PIPE_CHUNK_W")) chunk_w = (int)std::min<int64_t>(std::max(1, atoi(e)), nw);  // tuning override
        } else {
            // popcount engine: a chunk keeps enough CTAs (>= 16 per SM) that the scan needs no row split
            chunk_w = (int)std::max<int64_t>(ceil_div(nw, 16), ceil_div((int64_t)kNumSMs * 16, qtiles));
            if (chunk_w > nw) chunk_w = nw;
        }
    }
    // uniform chunks of chunk_w windows; with many chunks the first and the last one are cut short (an eighth) so
    // that the pipeline fills (first H2D + expansion) and drains (last scan + D2H) quickly
    bounds.clear();
    bounds.push_back(0);
    if (nw <= 0) { bounds.push_back(0); return SNV_OK; }
    const int nfull = (int)ceil_div(nw, chunk_w);
    const int small = std::max(1, chunk_w / 8);
    if (chunk_w < nw && nfull >= 8 && small < chunk_w) {
        bounds.push_back(small);
        int at = small;
        while (nw - at > chunk_w + small) { at += chunk_w; bounds.push_back(at); }
        if (nw - at > small) { at = nw - small; bounds.push_back(at); }
    } else {
        for (int at = chunk_w; at < nw; at += chunk_w) bounds.push_back(at);
    }
    bounds.push_back(nw);
    return SNV_OK;
}

// c_D16 / c_I32 non-null: compact results (snv_index_search_compact) - the scan writes (int32, int64) rows into the
// index's workspace, a narrowing kernel turns each chunk into (uint16, int32) rows, and only those travel.
static int search_hamming(snv_index* idx, int w0, int nw, const void* q, int64_t nq, int q_dtype,
                          const void* mask, int mask_mode, int k, int64_t id_offset, int32_t* D_i32,
                          float* D_f32, int64_t* I, unsigned flags, cudaStream_t stream, uint16_t* c_D16 = nullptr,
                          int32_t* c_I32 = nullptr)
{
    const bool compact = c_I32 != nullptr;
    const bool q_dev = flags & SNV_Q_ON_DEVICE;
    const bool out_dev = flags & SNV_OUT_ON_DEVICE;
    const bool invert = flags & SNV_MASK_IS_MISSING;
    const int64_t nqt = (int64_t)nw * nq;
    const size_t row_b = (size_t)idx->stride * 4;
    const bool tokens = q_dtype == SNV_DT_I64_TOKENS;
    if (dtype_row_bytes(q_dtype, idx->d, idx->stride) == 0) { set_error("search: bad dtype"); return SNV_ERR_INVALID; }
    if (tokens && mask_mode != SNV_MASK_NONE) { set_error("search: token queries carry their own mask"); return SNV_ERR_INVALID; }
    if (mask_mode != SNV_MASK_NONE) {
        if (!mask) { set_error("search: mask_mode set but mask is null"); return SNV_ERR_INVALID; }
        if (q_dtype == SNV_DT_PACKED_U8) { set_error("search: masks are not supported with byte-packed codes"); return SNV_ERR_INVALID; }
    }
    const size_t in_row = dtype_row_bytes(q_dtype, idx->d, idx->stride);
    const bool q_needs_pack = q_dtype != SNV_DT_PACKED_U32;
    const bool m_needs_pack = mask_mode != SNV_MASK_NONE && (q_dtype != SNV_DT_PACKED_U32 || invert);
    const int64_t mrows_total = mask_mode == SNV_MASK_PER_WINDOW ? nw : (mask_mode == SNV_MASK_PER_QUERY ? nqt : 0);

    // ---- reserve every workspace up front (nothing is reallocated while chunks are in flight)
    int rc;
    if (!q_dev) { rc = idx->ws_in.reserve((size_t)nqt * in_row); if (rc) return rc; }
    if (q_needs_pack) { rc = idx->ws_q.reserve((size_t)nqt * row_b); if (rc) return rc; }
    if (tokens) { rc = idx->ws_mask.reserve((size_t)nqt * row_b); if (rc) return rc; }
    if (mask_mode != SNV_MASK_NONE) {
        if (!q_dev) { rc = idx->ws_min.reserve((size_t)mrows_total * in_row); if (rc) return rc; }
        if (m_needs_pack) { rc = idx->ws_mask.reserve((size_t)mrows_total * row_b); if (rc) return rc; }
    }
    if (compact) {
        rc = idx->ws_di.reserve((size_t)nqt * k * 4); if (rc) return rc;
        rc = idx->ws_i.reserve((size_t)nqt * k * 8); if (rc) return rc;
        if (!out_dev) { rc = idx->ws_df.reserve((size_t)nqt * k * 6); if (rc) return rc; }  // int32 rows, then uint16 rows
    } else if (!out_dev) {
        if (D_i32) { rc = idx->ws_di.reserve((size_t)nqt * k * 4); if (rc) return rc; }
        if (D_f32) { rc = idx->ws_df.reserve((size_t)nqt * k * 4); if (rc) return rc; }
        rc = idx->ws_i.reserve((size_t)nqt * k * 8);
        if (rc) return rc;
    }

    // ---- chunking: host buffers are pipelined over internal streams, window chunk by window chunk;
    // (popcount engine) a chunk keeps enough CTAs (>= 16 per SM) that the scan needs no row split
    auto make_params = [&](int wb, int wc, HammingSearchParams& p) {
        p = HammingSearchParams{};
        p.panel = idx->panel + (int64_t)(w0 + wb) * idx->cap * idx->stride;
        p.panel_win_stride = idx->cap * idx->stride;
        p.words = idx->words;
        p.stride = idx->stride;
        p.d = (int)idx->d;
        p.n = idx->ntotal;
        p.nq = (int)nq;
        p.nw = wc;
        p.k = k;
        p.id_offset = id_offset;
        p.mask = mask_mode != SNV_MASK_NONE || tokens ? (const uint32_t*)1 : nullptr;  // plan only needs null-ness
    };
    std::vector<int> bounds;
    {
        HammingSearchParams probe;
        make_params(0, nw, probe);
        rc = hamming_chunk_bounds(probe, !q_dev || !out_dev, bounds);
        if (rc) return rc;
    }
    const int nchunks = (int)bounds.size() - 1;
    size_t part_chunk = 0;  // partial-key bytes per chunk (row-split plans); every chunk gets its own slice
    size_t tc_chunk = 0;    // tensor-core engine workspace bytes per chunk
    {
        HammingSearchParams probe;
        HammingTcPlan tplan;
        int seen[4] = {0, 0, 0, 0};  // distinct chunk sizes (at most: short, full, remainder)
        int nseen = 0;
        for (int c = 0; c < nchunks; ++c) {
            const int wc = bounds[c + 1] - bounds[c];
            bool dup = false;
            for (int i = 0; i < nseen; ++i) dup = dup || seen[i] == wc;
            if (dup) continue;
            if (nseen < 4) seen[nseen++] = wc;
            make_params(0, wc, probe);
            const size_t part = hamming_plan(probe);
            if (part == (size_t)-1) return SNV_ERR_UNSUPPORTED;
            part_chunk = std::max(part_chunk, part);
            // tensor-core engine (chosen by shape): per-chunk slice for the query operand rows, biases, partial keys
            const size_t tcb = hamming_tc_plan(probe, tplan);
            if (tcb == (size_t)-1) return SNV_ERR_UNSUPPORTED;
            tc_chunk = std::max(tc_chunk, tcb);
        }
        part_chunk = (size_t)round_up((int64_t)part_chunk, 256);
        if (part_chunk) { rc = idx->ws_partial.reserve(part_chunk * (size_t)nchunks); if (rc) return rc; }
        tc_chunk = (size_t)round_up((int64_t)tc_chunk, 1024);
        if (tc_chunk) { rc = idx->ws_qops.reserve(tc_chunk * (size_t)nchunks); if (rc) return rc; }
    }
    const bool piped = nchunks > 1;
    if (piped) {
        rc = ensure_pipe(idx);
        if (rc) return rc;
        SNV_CUDA_CHECK(cudaEventRecord(idx->pipe_start, stream));
        for (int i = 0; i < snv_index::kPipeStreams; ++i) SNV_CUDA_CHECK(cudaStreamWaitEvent(idx->pipe_stream[i], idx->pipe_start, 0));
    }

    // one chunk; an error return leaves the loop, and the internal streams are joined with the caller's stream either way
    auto run_chunk = [&](int c) -> int {
        const int wb = bounds[c];
        const int wc = bounds[c + 1] - wb;
        const int64_t r0 = (int64_t)wb * nq, rows = (int64_t)wc * nq;
        cudaStream_t cs = piped ? idx->pipe_stream[c % snv_index::kPipeStreams] : stream;
        HammingSearchParams p;
        make_params(wb, wc, p);
        p.mask = nullptr;

        const uint32_t* qd = nullptr;
        uint32_t* obs = tokens ? (uint32_t*)idx->ws_mask.p + r0 * idx->stride : nullptr;
        rc = stage_packed(idx, (const char*)q + (size_t)r0 * in_row, rows, q_dtype, q_dev, false,
                          q_dev ? nullptr : (char*)idx->ws_in.p + (size_t)r0 * in_row,
                          q_needs_pack ? (uint32_t*)idx->ws_q.p + r0 * idx->stride : nullptr, &qd, obs, cs);
        if (rc) return rc;
        p.q = qd;
        if (tokens) {
            p.mask = obs;
            p.mask_win_stride = nq * idx->stride;
            p.mask_q_stride = idx->stride;
        } else if (mask_mode != SNV_MASK_NONE) {
            const int64_t mr0 = mask_mode == SNV_MASK_PER_WINDOW ? wb : r0;
            const int64_t mrows = mask_mode == SNV_MASK_PER_WINDOW ? wc : rows;
            const uint32_t* md = nullptr;
            rc = stage_packed(idx, (const char*)mask + (size_t)mr0 * in_row, mrows, q_dtype, q_dev, invert,
                              q_dev ? nullptr : (char*)idx->ws_min.p + (size_t)mr0 * in_row,
                              m_needs_pack ? (uint32_t*)idx->ws_mask.p + mr0 * idx->stride : nullptr, &md, nullptr, cs);
            if (rc) return rc;
            p.mask = md;
            p.mask_win_stride = mask_mode == SNV_MASK_PER_WINDOW ? idx->stride : nq * idx->stride;
            p.mask_q_stride = mask_mode == SNV_MASK_PER_WINDOW ? 0 : idx->stride;
        }
        const int64_t o0 = r0 * k;
        if (compact) {
            p.D_i32 = (int32_t*)idx->ws_di.p + o0;
            p.D_f32 = nullptr;
            p.I = (int64_t*)idx->ws_i.p + o0;
        } else if (out_dev) {
            p.D_i32 = D_i32 ? D_i32 + o0 : nullptr;
            p.D_f32 = D_f32 ? D_f32 + o0 : nullptr;
            p.I = I + o0;
        } else {
            p.D_i32 = D_i32 ? (int32_t*)idx->ws_di.p + o0 : nullptr;
            p.D_f32 = D_f32 ? (float*)idx->ws_df.p + o0 : nullptr;
            p.I = (int64_t*)idx->ws_i.p + o0;
        }
        HammingTcPlan tplan;
        const size_t tc_need = hamming_tc_plan(p, tplan);
        if (tc_need == (size_t)-1) return SNV_ERR_UNSUPPORTED;
        g_last_hamming_engine = tplan.engine;
        if (tplan.engine) {
            if (tc_need > tc_chunk) { set_error("search: internal tensor-core workspace sizing error"); return SNV_ERR_INVALID; }
            rc = hamming_tc_launch(p, tplan, (char*)idx->ws_qops.p + (size_t)c * tc_chunk, cs);
            if (rc) return rc;
        } else {
            const size_t part = hamming_plan(p);
            if (part == (size_t)-1) return SNV_ERR_UNSUPPORTED;
            if (part > part_chunk) { set_error("search: internal partial-buffer sizing error"); return SNV_ERR_INVALID; }
            p.partial = part ? (uint64_t*)((char*)idx->ws_partial.p + (size_t)c * part_chunk) : nullptr;
            rc = hamming_launch(p, cs);
            if (rc) return rc;
        }
        if (compact) {
            const size_t cnt = (size_t)rows * k;
            // staging: the int32 rows first, the uint16 rows behind them (keeps both naturally aligned)
            int32_t* i32 = out_dev ? c_I32 + o0 : (int32_t*)idx->ws_df.p + o0;
            uint16_t* d16 = out_dev ? c_D16 + o0 : (uint16_t*)((int32_t*)idx->ws_df.p + (size_t)nqt * k) + o0;
            rc = narrow_results_launch(p.D_i32, p.I, (int64_t)cnt, d16, i32, cs);
            if (rc) return rc;
            if (!out_dev) {
                SNV_CUDA_CHECK(cudaMemcpyAsync(c_D16 + o0, d16, cnt * 2, cudaMemcpyDeviceToHost, cs));
                SNV_CUDA_CHECK(cudaMemcpyAsync(c_I32 + o0, i32, cnt * 4, cudaMemcpyDeviceToHost, cs));
            }
        } else if (!out_dev) {
            const size_t cnt = (size_t)rows * k;
            if (D_i32) SNV_CUDA_CHECK(cudaMemcpyAsync(D_i32 + o0, p.D_i32, cnt * 4, cudaMemcpyDeviceToHost, cs));
            if (D_f32) SNV_CUDA_CHECK(cudaMemcpyAsync(D_f32 + o0, p.D_f32, cnt * 4, cudaMemcpyDeviceToHost, cs));
            SNV_CUDA_CHECK(cudaMemcpyAsync(I + o0, p.I, cnt * 8, cudaMemcpyDeviceToHost, cs));
        }
        return SNV_OK;
    };
    rc = SNV_OK;
    for (int c = 0; c < nchunks && rc == SNV_OK; ++c) rc = run_chunk(c);
    if (piped) {
        for (int i = 0; i < snv_index::kPipeStreams; ++i) {
            const cudaError_t e1 = cudaEventRecord(idx->pipe_done[i], idx->pipe_stream[i]);
            const cudaError_t e2 = e1 == cudaSuccess ? cudaStreamWaitEvent(stream, idx->pipe_done[i], 0) : e1;
            if (e2 != cudaSuccess && rc == SNV_OK) { set_error(std::string("search: joining the pipeline streams: ") + cudaGetErrorString(e2)); rc = SNV_ERR_CUDA; }
        }
    }
    if (rc != SNV_OK) {
        // chunks already queued still use the shared workspaces and the caller's buffers: let them finish before returning
        cudaStreamSynchronize(stream);
        cudaGetLastError();
        return rc;
    }
    if (!out_dev || !q_dev) SNV_CUDA_CHECK(cudaStreamSynchronize(stream));
    return SNV_OK;
}

static int search_hamming_grouped(snv_index* idx, const void* q, const int32_t* window_ids, int64_t nqt,
                                  int q_dtype, const void* mask, int k, int32_t* D_i32, float* D_f32,
                                  int64_t* I, unsigned flags, cudaStream_t stream)
{
    const bool q_dev = flags & SNV_Q_ON_DEVICE;
    const bool out_dev = flags & SNV_OUT_ON_DEVICE;
    const bool invert = flags & SNV_MASK_IS_MISSING;
    const size_t row_b = (size_t)idx->stride * 4;
    const bool tokens = q_dtype == SNV_DT_I64_TOKENS;
    const size_t in_row = dtype_row_bytes(q_dtype, idx->d, idx->stride);
    if (in_row == 0) { set_error("search_grouped: bad dtype"); return SNV_ERR_INVALID; }
    if (tokens && mask) { set_error("search_grouped: token queries carry their own mask"); return SNV_ERR_INVALID; }
    if (mask && q_dtype == SNV_DT_PACKED_U8) { set_error("search_grouped: masks are not supported with byte-packed codes"); return SNV_ERR_INVALID; }

    // ---- host: stable counting sort of the queries by window -> permutation + work items
    const int W = idx->n_windows;
    std::vector<int32_t> count(W + 1, 0);
    for (int64_t i = 0; i < nqt; ++i) {
        const int32_t w = window_ids[i];
        if (w < 0 || w >= W) { set_error("search_grouped: window id out of range"); return SNV_ERR_INVALID; }
        ++count[w + 1];
    }
    int n_groups = 0;
    for (int w = 0; w < W; ++w) n_groups += count[w + 1] > 0;
    for (int w = 0; w < W; ++w) count[w + 1] += count[w];
    std::vector<int32_t> order(nqt);
    {
        std::vector<int32_t> pos(count.begin(), count.end() - 1);
        for (int64_t i = 0; i < nqt; ++i) order[pos[window_ids[i]]++] = (int32_t)i;
    }
    HammingSearchParams p{};
    p.words = idx->words;
    p.stride = idx->stride;
    p.d = (int)idx->d;
    const int64_t avg = n_groups ? nqt / n_groups : 0;
    int block = avg >= 64 ? 128 : 32;
    {   // wide rows (generic kernel) always use 32-query blocks
        HammingSearchParams probe{};
        probe.words = idx->words; probe.stride = idx->stride; probe.d = (int)idx->d; probe.k = k;
        probe.n = idx->ntotal; probe.nq = 1; probe.nw = 1;
        if (hamming_plan(probe) == (size_t)-1) return SNV_ERR_UNSUPPORTED;
        if (!probe.nw_templ) block = 32;
    }
    std::vector<int32_t> work;
    for (int w = 0; w < W; ++w) {
        for (int32_t s0 = count[w]; s0 < count[w + 1]; s0 += block) {
            work.push_back(w);
            work.push_back(s0);
            work.push_back(std::min<int32_t>(block, count[w + 1] - s0));
        }
    }
    const int n_work = (int)(work.size() / 3);

    int rc;
    rc = idx->ws_misc.reserve((size_t)(nqt + work.size()) * 4);
    if (rc) return rc;
    int32_t* d_order = (int32_t*)idx->ws_misc.p;
    int32_t* d_work = d_order + nqt;
    SNV_CUDA_CHECK(cudaMemcpyAsync(d_order, order.data(), (size_t)nqt * 4, cudaMemcpyHostToDevice, stream));
    SNV_CUDA_CHECK(cudaMemcpyAsync(d_work, work.data(), work.size() * 4, cudaMemcpyHostToDevice, stream));

    const bool q_needs_pack = q_dtype != SNV_DT_PACKED_U32;
    const bool m_needs_pack = mask && (q_dtype != SNV_DT_PACKED_U32 || invert);
    if (!q_dev) { rc = idx->ws_in.reserve((size_t)nqt * in_row); if (rc) return rc; }
    if (q_needs_pack) { rc = idx->ws_q.reserve((size_t)nqt * row_b); if (rc) return rc; }
    if (tokens || m_needs_pack) { rc = idx->ws_mask.reserve((size_t)nqt * row_b); if (rc) return rc; }
    if (mask && !q_dev) { rc = idx->ws_min.reserve((size_t)nqt * in_row); if (rc) return rc; }

    const uint32_t* qd = nullptr;
    uint32_t* obs = tokens ? (uint32_t*)idx->ws_mask.p : nullptr;
    rc = stage_packed(idx, q, nqt, q_dtype, q_dev, false, idx->ws_in.p, (uint32_t*)idx->ws_q.p, &qd, obs, stream);
    if (rc) return rc;
    p.q = qd;
    if (tokens) {
        p.mask = obs;
        p.mask_q_stride = idx->stride;
    } else if (mask) {
        const uint32_t* md = nullptr;
        rc = stage_packed(idx, mask, nqt, q_dtype, q_dev, invert, idx->ws_min.p, (uint32_t*)idx->ws_mask.p, &md, nullptr, stream);
        if (rc) return rc;
        p.mask = md;
        p.mask_q_stride = idx->stride;
    }
    p.panel = idx->panel;
    p.panel_win_stride = idx->cap * idx->stride;
    p.n = idx->ntotal;
    p.nq = (int)std::min<int64_t>(nqt, 0x7fffffff);
    p.nw = W;
    p.k = k;
    p.id_offset = 0;
    p.work = d_work;
    p.order = d_order;
    p.n_work = n_work;
    p.work_block = block;
    p.nq_total = nqt;
    if (out_dev) {
        p.D_i32 = D_i32; p.D_f32 = D_f32; p.I = I;
    } else {
        if (D_i32) { rc = idx->ws_di.reserve((size_t)nqt * k * 4); if (rc) return rc; p.D_i32 = (int32_t*)idx->ws_di.p; }
        if (D_f32) { rc = idx->ws_df.reserve((size_t)nqt * k * 4); if (rc) return rc; p.D_f32 = (float*)idx->ws_df.p; }
        rc = idx->ws_i.reserve((size_t)nqt * k * 8);
        if (rc) return rc;
        p.I = (int64_t*)idx->ws_i.p;
    }
    const size_t part = hamming_plan(p);
    if (part == (size_t)-1) return SNV_ERR_UNSUPPORTED;
    if (part) { rc = idx->ws_partial.reserve(part); if (rc) return rc; p.partial = (uint64_t*)idx->ws_partial.p; }
    g_last_hamming_engine = 0;
    rc = hamming_launch(p, stream);
    if (rc) return rc;
    if (!out_dev) {
        const size_t cnt = (size_t)nqt * k;
        if (D_i32) SNV_CUDA_CHECK(cudaMemcpyAsync(D_i32, p.D_i32, cnt * 4, cudaMemcpyDeviceToHost, stream));
        if (D_f32) SNV_CUDA_CHECK(cudaMemcpyAsync(D_f32, p.D_f32, cnt * 4, cudaMemcpyDeviceToHost, stream));
        SNV_CUDA_CHECK(cudaMemcpyAsync(I, p.I, cnt * 8, cudaMemcpyDeviceToHost, stream));
    }
    // order/work live in host vectors that die with this frame: always wait for their upload
    SNV_CUDA_CHECK(cudaStreamSynchronize(stream));
    return SNV_OK;
}

static int search_l2(snv_index* idx, int w0, int nw, const void* q, int64_t nq, int k, int64_t id_offset,
                     float* D_f32, int64_t* I, unsigned flags, cudaStream_t stream)
{
    const bool q_dev = flags & SNV_Q_ON_DEVICE;
    const bool out_dev = flags & SNV_OUT_ON_DEVICE;
    const int64_t nqt = (int64_t)nw * nq;
    const float* qd = (const float*)q;
    int rc;
    if (!q_dev) {
        rc = idx->ws_in.reserve((size_t)nqt * idx->d * 4);
        if (rc) return rc;
        SNV_CUDA_CHECK(cudaMemcpyAsync(idx->ws_in.p, q, (size_t)nqt * idx->d * 4, cudaMemcpyHostToDevice, stream));
        qd = (const float*)idx->ws_in.p;
    }
    rc = idx->ws_qops.reserve((size_t)nqt * idx->kp * 4);
    if (rc) return rc;
    const int pchunks = l2_prep_chunks(idx->d);
    rc = idx->ws_qnorm.reserve((size_t)nqt * 4 * (pchunks > 1 ? 1 + pchunks : 1));
    if (rc) return rc;
    {
        const bool center = (idx->l2_mode & SNV_L2_CENTER) && idx->mean;
        for (int w = 0; w < (center ? nw : 1); ++w) {
            const int64_t rows = center ? nq : nqt;
            rc = l2_prep_launch(qd + (size_t)w * nq * idx->d, center ? idx->mean + (size_t)(w0 + w) * idx->d : nullptr, rows,
                                idx->d, idx->l2_mode, true, idx->kp, (float*)idx->ws_qops.p + (size_t)w * nq * idx->kp,
                                (float*)idx->ws_qnorm.p + (size_t)w * nq,
                                pchunks > 1 ? (float*)idx->ws_qnorm.p + nqt + (size_t)w * nq * pchunks : nullptr, stream);
            if (rc) return rc;
        }
    }
    float* Dd = D_f32;
    int64_t* Id = I;
    if (!out_dev) {
        rc = idx->ws_df.reserve((size_t)nqt * k * 4);
        if (rc) return rc;
        rc = idx->ws_i.reserve((size_t)nqt * k * 8);
        if (rc) return rc;
        Dd = (float*)idx->ws_df.p;
        Id = (int64_t*)idx->ws_i.p;
    } else if (!Dd) {
        rc = idx->ws_df.reserve((size_t)nqt * k * 4);
        if (rc) return rc;
        Dd = (float*)idx->ws_df.p;
    }
    for (int w = 0; w < nw; ++w) {
        L2SearchParams p{};
        p.ref_ops = idx->ops + (size_t)(w0 + w) * idx->cap * idx->kp;
        p.ref_norm = idx->norms + (size_t)(w0 + w) * idx->cap;
        p.q_ops = (const float*)idx->ws_qops.p + (size_t)w * nq * idx->kp;
        p.q_norm = (const float*)idx->ws_qnorm.p + (size_t)w * nq;
        p.n = idx->ntotal;
        p.nq = nq;
        p.kp = idx->kp;
        p.k = k;
        p.id_offset = id_offset;
        p.D_f32 = Dd + (size_t)w * nq * k;
        p.I = Id + (size_t)w * nq * k;
        p.one = 1;
        const size_t part = l2_plan(p);
        if (part == (size_t)-1) return SNV_ERR_UNSUPPORTED;
        rc = idx->ws_partial.reserve(part ? part : 16);
        if (rc) return rc;
        p.partial = (uint64_t*)idx->ws_partial.p;
        rc = l2_launch(p, stream);
        if (rc) return rc;
    }
    if (!out_dev) {
        if (D_f32) SNV_CUDA_CHECK(cudaMemcpyAsync(D_f32, Dd, (size_t)nqt * k * 4, cudaMemcpyDeviceToHost, stream));
        SNV_CUDA_CHECK(cudaMemcpyAsync(I, Id, (size_t)nqt * k * 8, cudaMemcpyDeviceToHost, stream));
    }
    if (!out_dev || !q_dev) SNV_CUDA_CHECK(cudaStreamSynchronize(stream));
    return SNV_OK;
}

int snv_index_search(snv_index* idx, int w0, int nw, const void* q, int64_t nq, int q_dtype,
                     const void* mask, int mask_mode, int k, int64_t id_offset, int32_t* D_i32,
                     float* D_f32, int64_t* I, unsigned flags, void* stream_)
{
    if (!idx) { set_error("snv_index_search: null index"); return SNV_ERR_INVALID; }
    if (w0 < 0 || nw < 0 || w0 + nw > idx->n_windows) { set_error("snv_index_search: window range out of bounds"); return SNV_ERR_INVALID; }
    if (nq < 0) { set_error("snv_index_search: negative nq"); return SNV_ERR_INVALID; }
    if (k < 1) { set_error("snv_index_search: k must be >= 1"); return SNV_ERR_INVALID; }
    if (nw == 0 || nq == 0) return SNV_OK;
    if (!q || !I) { set_error("snv_index_search: q and I must not be null"); return SNV_ERR_INVALID; }
    if (nq > 0x7fffffff) { set_error("snv_index_search: nq too large"); return SNV_ERR_INVALID; }
    if (mask_mode < SNV_MASK_NONE || mask_mode > SNV_MASK_PER_QUERY) { set_error("snv_index_search: bad mask_mode"); return SNV_ERR_INVALID; }
    DeviceGuard g(idx->device);
    if (!g.ok) { set_error("snv_index_search: cudaSetDevice failed"); return SNV_ERR_CUDA; }
    cudaStream_t stream = (cudaStream_t)stream_;
    if (idx->kind == SNV_KIND_HAMMING) {
        return search_hamming(idx, w0, nw, q, nq, q_dtype, mask, mask_mode, k, id_offset, D_i32, D_f32, I, flags, stream);
    }
    if (q_dtype != SNV_DT_F32) { set_error("snv_index_search: an L2 index takes float32 queries"); return SNV_ERR_INVALID; }
    if (mask_mode != SNV_MASK_NONE) { set_error("snv_index_search: masks apply to HAMMING indexes only"); return SNV_ERR_INVALID; }
    if (D_i32) { set_error("snv_index_search: D_i32 applies to HAMMING indexes only"); return SNV_ERR_INVALID; }
    return search_l2(idx, w0, nw, q, nq, k, id_offset, D_f32, I, flags, stream);
}

int snv_index_search_compact(snv_index* idx, int w0, int nw, const void* q, int64_t nq, int q_dtype, const void* mask,
                             int mask_mode, int k, uint16_t* D_u16, int32_t* I_i32, unsigned flags, void* stream_)
{
    if (!idx || idx->kind != SNV_KIND_HAMMING) { set_error("snv_index_search_compact: needs a HAMMING index"); return SNV_ERR_INVALID; }
    if (w0 < 0 || nw < 0 || w0 + nw > idx->n_windows) { set_error("snv_index_search_compact: window range out of bounds"); return SNV_ERR_INVALID; }
    if (nq < 0 || nq > 0x7fffffff) { set_error("snv_index_search_compact: bad nq"); return SNV_ERR_INVALID; }
    if (k < 1) { set_error("snv_index_search_compact: k must be >= 1"); return SNV_ERR_INVALID; }
    if (idx->ntotal > 0x7fffffffLL || idx->d >= 65535) { set_error("snv_index_search_compact: ids or distances do not fit the compact types"); return SNV_ERR_UNSUPPORTED; }
    if (nw == 0 || nq == 0) return SNV_OK;
    if (!q || !D_u16 || !I_i32) { set_error("snv_index_search_compact: q, D and I must not be null"); return SNV_ERR_INVALID; }
    if (mask_mode < SNV_MASK_NONE || mask_mode > SNV_MASK_PER_QUERY) { set_error("snv_index_search_compact: bad mask_mode"); return SNV_ERR_INVALID; }
    DeviceGuard g(idx->device);
    if (!g.ok) { set_error("snv_index_search_compact: cudaSetDevice failed"); return SNV_ERR_CUDA; }
    return search_hamming(idx, w0, nw, q, nq, q_dtype, mask, mask_mode, k, 0, nullptr, nullptr, nullptr, flags, (cudaStream_t)stream_,
                          D_u16, I_i32);
}

int snv_index_search_grouped(snv_index* idx, const void* q, const int32_t* window_ids, int64_t nq_total,
                             int q_dtype, const void* mask, int k, int32_t* D_i32, float* D_f32, int64_t* I,
                             unsigned flags, void* stream_)
{
    if (!idx || idx->kind != SNV_KIND_HAMMING) { set_error("snv_index_search_grouped: needs a HAMMING index"); return SNV_ERR_INVALID; }
    if (nq_total < 0 || nq_total > 0x7fffffff) { set_error("snv_index_search_grouped: bad nq_total"); return SNV_ERR_INVALID; }
    if (k < 1) { set_error("snv_index_search_grouped: k must be >= 1"); return SNV_ERR_INVALID; }
    if (nq_total == 0) return SNV_OK;
    if (!q || !I || !window_ids) { set_error("snv_index_search_grouped: null buffer"); return SNV_ERR_INVALID; }
    DeviceGuard g(idx->device);
    if (!g.ok) { set_error("snv_index_search_grouped: cudaSetDevice failed"); return SNV_ERR_CUDA; }
    return search_hamming_grouped(idx, q, window_ids, nq_total, q_dtype, mask, k, D_i32, D_f32, I, flags, (cudaStream_t)stream_);
}

int snv_index_gather_tokens_grouped(snv_index* idx, const int64_t* I, const int32_t* window_ids, int64_t nq_total,
                                    int k, const int32_t* n_sites, int seq_len, int64_t* out, unsigned flags,
                                    void* stream_)
{
    if (!idx || idx->kind != SNV_KIND_HAMMING) { set_error("snv_index_gather_tokens_grouped: needs a HAMMING index"); return SNV_ERR_INVALID; }
    if (nq_total < 0 || k < 1 || seq_len < 1) { set_error("snv_index_gather_tokens_grouped: bad arguments"); return SNV_ERR_INVALID; }
    if (nq_total == 0) return SNV_OK;
    if (!I || !out || !window_ids) { set_error("snv_index_gather_tokens_grouped: null buffer"); return SNV_ERR_INVALID; }
    DeviceGuard g(idx->device);
    if (!g.ok) { set_error("snv_index_gather_tokens_grouped: cudaSetDevice failed"); return SNV_ERR_CUDA; }
    cudaStream_t stream = (cudaStream_t)stream_;
    const bool i_dev = flags & SNV_Q_ON_DEVICE;
    const bool out_dev = flags & SNV_OUT_ON_DEVICE;
    const int W = idx->n_windows;
    // per-query (window, n_sites) pairs, uploaded once
    std::vector<int32_t> meta((size_t)nq_total * 2);
    for (int64_t i = 0; i < nq_total; ++i) {
        const int32_t w = window_ids[i];
        if (w < 0 || w >= W) { set_error("snv_index_gather_tokens_grouped: window id out of range"); return SNV_ERR_INVALID; }
        const int32_t ns = n_sites ? n_sites[w] : (int32_t)idx->d;
        if (ns < 0 || ns > idx->d) { set_error("snv_index_gather_tokens_grouped: n_sites out of range"); return SNV_ERR_INVALID; }
        meta[2 * i] = w;
        meta[2 * i + 1] = ns;
    }
    int rc = idx->ws_misc.reserve(meta.size() * 4);
    if (rc) return rc;
    SNV_CUDA_CHECK(cudaMemcpyAsync(idx->ws_misc.p, meta.data(), meta.size() * 4, cudaMemcpyHostToDevice, stream));
    const int64_t rows = nq_total * k;
    const int64_t* Id = I;
    if (!i_dev) {
        rc = idx->ws_i.reserve((size_t)rows * 8);
        if (rc) return rc;
        SNV_CUDA_CHECK(cudaMemcpyAsync(idx->ws_i.p, I, (size_t)rows * 8, cudaMemcpyHostToDevice, stream));
        Id = (const int64_t*)idx->ws_i.p;
    }
    int64_t* od = out;
    if (!out_dev) {
        rc = idx->ws_in.reserve((size_t)rows * seq_len * 8);
        if (rc) return rc;
        od = (int64_t*)idx->ws_in.p;
    }
    rc = gather_tokens_grouped_launch(idx->panel, idx->cap * idx->stride, idx->stride, idx->ntotal, Id,
                                      (const int32_t*)idx->ws_misc.p, nq_total, k, seq_len, od, stream);
    if (rc) return rc;
    if (!out_dev) SNV_CUDA_CHECK(cudaMemcpyAsync(out, od, (size_t)rows * seq_len * 8, cudaMemcpyDeviceToHost, stream));
    SNV_CUDA_CHECK(cudaStreamSynchronize(stream));
    return SNV_OK;
}

int snv_index_gather_tokens(snv_index* idx, int w0, int nw, const int64_t* I, int64_t nq, int k,
                            const int32_t* n_sites, int seq_len, int64_t* out, unsigned flags, void* stream_)
{
    if (!idx || idx->kind != SNV_KIND_HAMMING) { set_error("snv_index_gather_tokens: needs a HAMMING index"); return SNV_ERR_INVALID; }
    if (w0 < 0 || nw < 0 || w0 + nw > idx->n_windows || nq < 0 || k < 1 || seq_len < 1) { set_error("snv_index_gather_tokens: bad arguments"); return SNV_ERR_INVALID; }
    if (nw == 0 || nq == 0) return SNV_OK;
    if (!I || !out) { set_error("snv_index_gather_tokens: null buffer"); return SNV_ERR_INVALID; }
    DeviceGuard g(idx->device);
    if (!g.ok) { set_error("snv_index_gather_tokens: cudaSetDevice failed"); return SNV_ERR_CUDA; }
    cudaStream_t stream = (cudaStream_t)stream_;
    const bool i_dev = flags & SNV_Q_ON_DEVICE;
    const bool out_dev = flags & SNV_OUT_ON_DEVICE;
    const int64_t rows = (int64_t)nw * nq * k;
    const int64_t* Id = I;
    int rc;
    if (!i_dev) {
        rc = idx->ws_i.reserve((size_t)rows * 8);
        if (rc) return rc;
        SNV_CUDA_CHECK(cudaMemcpyAsync(idx->ws_i.p, I, (size_t)rows * 8, cudaMemcpyHostToDevice, stream));
        Id = (const int64_t*)idx->ws_i.p;
    }
    const int32_t* nsd = nullptr;
    if (n_sites) {
        for (int w = 0; w < nw; ++w)
            if (n_sites[w] < 0 || n_sites[w] > idx->d) { set_error("snv_index_gather_tokens: n_sites out of range"); return SNV_ERR_INVALID; }
        rc = idx->ws_misc.reserve((size_t)nw * 4);
        if (rc) return rc;
        SNV_CUDA_CHECK(cudaMemcpyAsync(idx->ws_misc.p, n_sites, (size_t)nw * 4, cudaMemcpyHostToDevice, stream));
        nsd = (const int32_t*)idx->ws_misc.p;
    }
    int64_t* od = out;
    if (!out_dev) {
        rc = idx->ws_in.reserve((size_t)rows * seq_len * 8);
        if (rc) return rc;
        od = (int64_t*)idx->ws_in.p;
    }
    rc = gather_tokens_launch(idx->panel + (int64_t)w0 * idx->cap * idx->stride, idx->cap * idx->stride, idx->stride,
                              idx->ntotal, Id, 0, nw, nq, k, nsd, (int)idx->d, seq_len, od, stream);
    if (rc) return rc;
    if (!out_dev) SNV_CUDA_CHECK(cudaMemcpyAsync(out, od, (size_t)rows * seq_len * 8, cudaMemcpyDeviceToHost, stream));
    if (!out_dev || !i_dev || n_sites) SNV_CUDA_CHECK(cudaStreamSynchronize(stream));
    return SNV_OK;
}

int snv_index_gather_rows(snv_index* idx, int w0, int nw, const int64_t* I, int64_t nq, int k, float* out,
                          unsigned flags, void* stream_)
{
    if (!idx || idx->kind != SNV_KIND_L2) { set_error("snv_index_gather_rows: needs an L2 index"); return SNV_ERR_INVALID; }
    if (w0 < 0 || nw < 0 || w0 + nw > idx->n_windows || nq < 0 || k < 1) { set_error("snv_index_gather_rows: bad arguments"); return SNV_ERR_INVALID; }
    if (nw == 0 || nq == 0) return SNV_OK;
    if (!I || !out) { set_error("snv_index_gather_rows: null buffer"); return SNV_ERR_INVALID; }
    DeviceGuard g(idx->device);
    if (!g.ok) { set_error("snv_index_gather_rows: cudaSetDevice failed"); return SNV_ERR_CUDA; }
    cudaStream_t stream = (cudaStream_t)stream_;
    const bool i_dev = flags & SNV_Q_ON_DEVICE;
    const bool out_dev = flags & SNV_OUT_ON_DEVICE;
    const int64_t rows = (int64_t)nw * nq * k;
    const int64_t* Id = I;
    int rc;
    if (!i_dev) {
        rc = idx->ws_i.reserve((size_t)rows * 8);
        if (rc) return rc;
        SNV_CUDA_CHECK(cudaMemcpyAsync(idx->ws_i.p, I, (size_t)rows * 8, cudaMemcpyHostToDevice, stream));
        Id = (const int64_t*)idx->ws_i.p;
    }
    float* od = out;
    if (!out_dev) {
        rc = idx->ws_in.reserve((size_t)rows * idx->d * 4);
        if (rc) return rc;
        od = (float*)idx->ws_in.p;
    }
    rc = gather_rows_launch(idx->rows + (int64_t)w0 * idx->cap * idx->d, idx->cap * idx->d, idx->d, idx->ntotal, Id,
                            nw, nq, k, od, stream);
    if (rc) return rc;
    if (!out_dev) SNV_CUDA_CHECK(cudaMemcpyAsync(out, od, (size_t)rows * idx->d * 4, cudaMemcpyDeviceToHost, stream));
    if (!out_dev || !i_dev) SNV_CUDA_CHECK(cudaStreamSynchronize(stream));
    return SNV_OK;
}

int snv_index_export(snv_index* idx, int window, void* out)
{
    if (!idx || !out) { set_error("snv_index_export: null argument"); return SNV_ERR_INVALID; }
    if (window < 0 || window >= idx->n_windows) { set_error("snv_index_export: bad window"); return SNV_ERR_INVALID; }
    if (idx->ntotal == 0) return SNV_OK;
    DeviceGuard g(idx->device);
    if (!g.ok) { set_error("snv_index_export: cudaSetDevice failed"); return SNV_ERR_CUDA; }
    SNV_CUDA_CHECK(cudaDeviceSynchronize());
    if (idx->kind == SNV_KIND_HAMMING) {
        const size_t rb = (size_t)idx->stride * 4;
        SNV_CUDA_CHECK(cudaMemcpy(out, idx->panel + (size_t)window * idx->cap * idx->stride, (size_t)idx->ntotal * rb, cudaMemcpyDeviceToHost));
    } else {
        const size_t rb = (size_t)idx->d * 4;
        SNV_CUDA_CHECK(cudaMemcpy(out, idx->rows + (size_t)window * idx->cap * idx->d, (size_t)idx->ntotal * rb, cudaMemcpyDeviceToHost));
    }
    return SNV_OK;
}

int snv_topk_merge(int device, const int32_t* D_i32, const float* D_f32, const int64_t* I, int parts,
                   int64_t nq, int k_in, int k_out, int32_t* Do_i32, float* Do_f32, int64_t* Io, void* stream_)
{
    if ((D_i32 != nullptr) == (D_f32 != nullptr)) { set_error("snv_topk_merge: exactly one of D_i32 / D_f32"); return SNV_ERR_INVALID; }
    if ((D_i32 && !Do_i32) || (D_f32 && !Do_f32) || !I || !Io) { set_error("snv_topk_merge: null buffer"); return SNV_ERR_INVALID; }
    if (parts < 1 || nq < 0 || k_in < 1 || k_out < 1) { set_error("snv_topk_merge: bad sizes"); return SNV_ERR_INVALID; }
    DeviceGuard g(device);
    if (!g.ok) { set_error("snv_topk_merge: cudaSetDevice failed"); return SNV_ERR_CUDA; }
    return merge_results_launch(D_i32, D_f32, I, parts, nq, k_in, k_out, Do_i32, Do_f32, Io, (cudaStream_t)stream_);
}

int snv_exchange_pack(int device, const int32_t* D_i32, const int64_t* I, int nw, int64_t nq, int k, int parts, int64_t* keys,
                      void* stream)
{
    if (!D_i32 || !I || !keys || nw < 0 || nq < 0 || k < 1) { set_error("snv_exchange_pack: bad arguments"); return SNV_ERR_INVALID; }
    DeviceGuard g(device);
    if (!g.ok) { set_error("snv_exchange_pack: cudaSetDevice failed"); return SNV_ERR_CUDA; }
    return exchange_pack_launch(D_i32, I, nw, nq, k, parts, keys, (cudaStream_t)stream);
}

int snv_exchange_merge(int device, const int64_t* keys, int parts, int64_t n, int k_in, int k_out, int32_t* Do_i32, int64_t* Io,
                       void* stream)
{
    if (!keys || !Do_i32 || !Io || parts < 1 || n < 0 || k_in < 1) { set_error("snv_exchange_merge: bad arguments"); return SNV_ERR_INVALID; }
    DeviceGuard g(device);
    if (!g.ok) { set_error("snv_exchange_merge: cudaSetDevice failed"); return SNV_ERR_CUDA; }
    return exchange_merge_launch(keys, parts, n, k_in, k_out, Do_i32, Io, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------ NVLink peer exchange
// One allocation per rank: [0, 4 KiB) flag words (source s at byte 16 s) and the block counter (byte 2048), then two
// receive slots (even / odd epochs).  Peers map it through CUDA IPC; the tables of peer pointers live on the device.
struct snv_peer {
    int device = 0, rank = 0, world = 1;
    size_t slot_bytes = 0;
    char* base = nullptr;
    std::vector<char*> peer_base;     // [world]; [rank] = base
    std::vector<bool> opened;         // mapped through IPC (to be closed)
    int64_t** d_recv[2] = {nullptr, nullptr};  // device tables [world]
    uint64_t** d_flags = nullptr;
    uint64_t epoch = 0;
    bool wired = false;
};
static constexpr size_t kPeerHeader = 4096, kPeerCounterOff = 2048;

static int peer_wire(snv_peer* pe)
{
    std::vector<int64_t*> r0(pe->world), r1(pe->world);
    std::vector<uint64_t*> fl(pe->world);
    for (int g = 0; g < pe->world; ++g) {
        r0[g] = (int64_t*)(pe->peer_base[g] + kPeerHeader);
        r1[g] = (int64_t*)(pe->peer_base[g] + kPeerHeader + pe->slot_bytes);
        fl[g] = (uint64_t*)pe->peer_base[g];
    }
    const size_t tb = (size_t)pe->world * sizeof(void*);
    if (!pe->d_flags) {
        SNV_CUDA_CHECK(cudaMalloc((void**)&pe->d_recv[0], tb));
        SNV_CUDA_CHECK(cudaMalloc((void**)&pe->d_recv[1], tb));
        SNV_CUDA_CHECK(cudaMalloc((void**)&pe->d_flags, tb));
    }
    SNV_CUDA_CHECK(cudaMemcpy(pe->d_recv[0], r0.data(), tb, cudaMemcpyHostToDevice));
    SNV_CUDA_CHECK(cudaMemcpy(pe->d_recv[1], r1.data(), tb, cudaMemcpyHostToDevice));
    SNV_CUDA_CHECK(cudaMemcpy(pe->d_flags, fl.data(), tb, cudaMemcpyHostToDevice));
    pe->wired = true;
    return SNV_OK;
}

int snv_peer_create(int device, int rank, int world, size_t slot_bytes, snv_peer** out, void* handle_out)
{
    if (!out || world < 1 || world > 64 || rank < 0 || rank >= world || slot_bytes == 0) { set_error("snv_peer_create: bad arguments"); return SNV_ERR_INVALID; }
    DeviceGuard g(device);
    if (!g.ok) { set_error("snv_peer_create: cudaSetDevice failed"); return SNV_ERR_CUDA; }
    snv_peer* pe = new (std::nothrow) snv_peer();
    if (!pe) return SNV_ERR_NOMEM;
    pe->device = device; pe->rank = rank; pe->world = world;
    pe->slot_bytes = (size_t)round_up((int64_t)slot_bytes, 256);
    pe->peer_base.assign(world, nullptr);
    pe->opened.assign(world, false);
    const size_t total = kPeerHeader + 2 * pe->slot_bytes;
    cudaError_t e = cudaMalloc((void**)&pe->base, total);
    if (e != cudaSuccess) { cudaGetLastError(); delete pe; set_error("snv_peer_create: out of device memory"); return SNV_ERR_NOMEM; }
    e = cudaMemset(pe->base, 0, kPeerHeader);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e == cudaSuccess && handle_out) {
        cudaIpcMemHandle_t h;
        e = cudaIpcGetMemHandle(&h, pe->base);
        if (e == cudaSuccess) memcpy(handle_out, &h, sizeof(h));
    }
    if (e != cudaSuccess) {
        set_error(std::string("snv_peer_create: ") + cudaGetErrorString(e));
        cudaGetLastError(); cudaFree(pe->base); delete pe;
        return SNV_ERR_CUDA;
    }
    pe->peer_base[rank] = pe->base;
    *out = pe;
    return SNV_OK;
}

int snv_peer_open(snv_peer* pe, const void* handles)
{
    if (!pe || !handles) { set_error("snv_peer_open: null argument"); return SNV_ERR_INVALID; }
    DeviceGuard g(pe->device);
    if (!g.ok) { set_error("snv_peer_open: cudaSetDevice failed"); return SNV_ERR_CUDA; }
    static_assert(sizeof(cudaIpcMemHandle_t) == SNV_PEER_HANDLE_BYTES, "IPC handle size");
    for (int r = 0; r < pe->world; ++r) {
        if (r == pe->rank || pe->peer_base[r]) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char*)handles + (size_t)r * sizeof(h), sizeof(h));
        void* ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            set_error(std::string("snv_peer_open: cannot map rank ") + std::to_string(r) + "'s exchange buffer (no peer access between the GPUs?): " + cudaGetErrorString(e));
            return SNV_ERR_CUDA;
        }
        pe->peer_base[r] = (char*)ptr;
        pe->opened[r] = true;
    }
    return peer_wire(pe);
}

int snv_peer_open_local(snv_peer* pe, snv_peer* const* peers)
{
    if (!pe || !peers) { set_error("snv_peer_open_local: null argument"); return SNV_ERR_INVALID; }
    DeviceGuard g(pe->device);
    if (!g.ok) { set_error("snv_peer_open_local: cudaSetDevice failed"); return SNV_ERR_CUDA; }
    for (int r = 0; r < pe->world; ++r) {
        if (!peers[r] || peers[r]->rank != r || peers[r]->world != pe->world || peers[r]->slot_bytes != pe->slot_bytes) {
            set_error("snv_peer_open_local: peers[r] must be rank r of the same world and slot size");
            return SNV_ERR_INVALID;
        }
        pe->peer_base[r] = peers[r]->base;
    }
    return peer_wire(pe);
}

static int peer_run(snv_peer* pe, const char* what, int phases, uint64_t epoch, const int32_t* D_i32, const int64_t* I, int nw, int64_t nq,
                    int k, int k_out, int32_t* Do_i32, int64_t* Io, void* stream)
{
    if (!pe || nw < 0 || nq < 0) { set_error(std::string(what) + ": bad arguments"); return SNV_ERR_INVALID; }
    if ((phases & 1) && (!D_i32 || !I)) { set_error(std::string(what) + ": null candidates"); return SNV_ERR_INVALID; }
    if ((phases & 2) && (!Do_i32 || !Io)) { set_error(std::string(what) + ": null result buffers"); return SNV_ERR_INVALID; }
    if (!pe->wired) { set_error(std::string(what) + ": call snv_peer_open first"); return SNV_ERR_INVALID; }
    if ((size_t)nw * (size_t)nq * (size_t)(k > 0 ? k : 0) * 8 > pe->slot_bytes) { set_error(std::string(what) + ": batch larger than the exchange slot"); return SNV_ERR_INVALID; }
    DeviceGuard g(pe->device);
    if (!g.ok) { set_error(std::string(what) + ": cudaSetDevice failed"); return SNV_ERR_CUDA; }
    const int slot = (int)(epoch & 1u);
    return peer_exchange_launch(D_i32, I, nw, nq, k, pe->world, pe->rank, pe->d_recv[slot], pe->d_flags,
                                (const int64_t*)(pe->base + kPeerHeader + (size_t)slot * pe->slot_bytes), (const uint64_t*)pe->base,
                                (unsigned*)(pe->base + kPeerCounterOff), epoch, (phases & 2) ? k_out : 1, Do_i32, Io, phases, (cudaStream_t)stream);
}

int snv_peer_exchange(snv_peer* pe, const int32_t* D_i32, const int64_t* I, int nw, int64_t nq, int k, int k_out, int32_t* Do_i32,
                      int64_t* Io, void* stream)
{
    if (!pe) { set_error("snv_peer_exchange: null object"); return SNV_ERR_INVALID; }
    const int rc = peer_run(pe, "snv_peer_exchange", 3, pe->epoch + 1, D_i32, I, nw, nq, k, k_out, Do_i32, Io, stream);
    if (rc == SNV_OK) ++pe->epoch;
    return rc;
}

int snv_peer_push(snv_peer* pe, const int32_t* D_i32, const int64_t* I, int nw, int64_t nq, int k, uint64_t* epoch_out, void* stream)
{
    if (!pe || !epoch_out) { set_error("snv_peer_push: null argument"); return SNV_ERR_INVALID; }
    const int rc = peer_run(pe, "snv_peer_push", 1, pe->epoch + 1, D_i32, I, nw, nq, k, 1, nullptr, nullptr, stream);
    if (rc == SNV_OK) *epoch_out = ++pe->epoch;
    return rc;
}

int snv_peer_merge(snv_peer* pe, uint64_t epoch, int nw, int64_t nq, int k, int k_out, int32_t* Do_i32, int64_t* Io, void* stream)
{
    if (!pe) { set_error("snv_peer_merge: null object"); return SNV_ERR_INVALID; }
    if (epoch == 0 || epoch > pe->epoch || epoch + 1 < pe->epoch) {
        set_error("snv_peer_merge: epoch must be that of the last or the last-but-one push (its slot has been reused otherwise)");
        return SNV_ERR_INVALID;
    }
    return peer_run(pe, "snv_peer_merge", 2, epoch, nullptr, nullptr, nw, nq, k, k_out, Do_i32, Io, stream);
}

int snv_peer_destroy(snv_peer* pe)
{
    if (!pe) return SNV_OK;
    DeviceGuard g(pe->device);
    if (g.ok) {
        cudaDeviceSynchronize();
        for (int r = 0; r < pe->world; ++r)
            if (pe->opened[r] && pe->peer_base[r]) cudaIpcCloseMemHandle(pe->peer_base[r]);
        cudaFree(pe->d_recv[0]); cudaFree(pe->d_recv[1]); cudaFree(pe->d_flags);
        cudaFree(pe->base);
        cudaGetLastError();
    }
    delete pe;
    return SNV_OK;
}

int snv_intersect_masks(int device, const int64_t* ref_pos, int64_t n_ref, const int64_t* tgt_pos, int64_t n_tgt,
                        const int64_t* window_info, int n_windows, int64_t d, int ploidy, uint32_t* out, void* stream_)
{
    if (!ref_pos || !window_info || !out || (!tgt_pos && n_tgt > 0)) { set_error("snv_intersect_masks: null buffer"); return SNV_ERR_INVALID; }
    if (n_ref < 0 || n_tgt < 0 || n_windows < 0 || d <= 0 || (ploidy != 1 && ploidy != 2)) { set_error("snv_intersect_masks: bad arguments"); return SNV_ERR_INVALID; }
    DeviceGuard g(device);
    if (!g.ok) { set_error("snv_intersect_masks: cudaSetDevice failed"); return SNV_ERR_CUDA; }
    return intersect_masks_launch(ref_pos, n_ref, tgt_pos, n_tgt, window_info, n_windows, d, ploidy, (int)snv_packed_stride(d), out,
                                  (cudaStream_t)stream_);
}

int snv_pack_rows(int device, const void* x, int64_t rows, int64_t d, int dtype, int invert, uint32_t* out,
                  uint32_t* out_observed, void* stream_)
{
    if (!x || !out || rows < 0 || d <= 0) { set_error("snv_pack_rows: bad arguments"); return SNV_ERR_INVALID; }
    DeviceGuard g(device);
    if (!g.ok) { set_error("snv_pack_rows: cudaSetDevice failed"); return SNV_ERR_CUDA; }
    return pack_launch(x, rows, d, dtype, invert != 0, (int)snv_packed_stride(d), out, out_observed, (cudaStream_t)stream_);
}

}  // extern "C"
