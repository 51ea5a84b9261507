// tav_api.cu — the C ABI of libtavec (include/tavec.h): index lifecycle, append / adopt,
// the search dispatcher (row-scan path or tcgen05 path) and the shard-merge entry point.
//
// Reference surface this stands in for (src/typeagent/aitools/vectorbase.py):
//   add_embedding(s) :115-148 -> tav_append          clear :253-256        -> tav_clear
//   deserialize      :273-287 -> tav_append (bulk)   fuzzy_lookup_embedding :163-201 and
//   fuzzy_lookup_embedding_in_subset :203-230        -> tav_search
// There is no CPU path in this library: every entry point that computes needs a CUDA device.

#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <vector>

#include "tav_common.cuh"
#include "tav_internal.h"

namespace tav {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

static inline size_t dtype_size(int dt) { return dt == TAV_F32 ? 4 : 2; }
static inline bool dtype_ok(int dt) { return dt == TAV_F32 || dt == TAV_BF16 || dt == TAV_F16; }

#define TAV_CUDA(expr)                                                                     \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return _e == cudaErrorMemoryAllocation ? TAV_ERR_OOM : TAV_ERR_CUDA;           \
        }                                                                                  \
    } while (0)

// a grow-only device buffer
struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    cudaError_t ensure(size_t need) {
        if (need <= bytes) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
        size_t want = std::max(need, size_t(1) << 16);
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) bytes = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
};

// a grow-only pinned host buffer (staging for truly asynchronous H2D / D2H of small payloads)
struct PinBuf {
    void* p = nullptr;
    size_t bytes = 0;
    cudaError_t ensure(size_t need) {
        if (need <= bytes) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        bytes = 0;
        size_t want = std::max(need, size_t(1) << 16);
        cudaError_t e = cudaMallocHost(&p, want);
        if (e == cudaSuccess) bytes = want;
        return e;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        bytes = 0;
    }
};

constexpr int kMaxTimedChunks = 64;
constexpr size_t kPinnedStageLimit = size_t(8) << 20;  // payloads above this go straight from user memory
constexpr size_t kZeroCopyOutLimit = size_t(16) << 10; // results up to this size are written to host memory by the kernels

}  // namespace tav

using namespace tav;

struct tav_index {
    int device = 0;
    int dim = 0;
    int dtype = TAV_F32;
    int flags = 0;
    int64_t size = 0;
    int64_t capacity = 0;
    void* rows = nullptr;  // [capacity, dim] storage dtype
    bool adopted = false;

    // search workspace
    DevBuf queries;     // float32 [n_queries, dim]
    DevBuf subset;      // int64 [subset_len]
    DevBuf cand_keys;   // [qb, cand_stride] uint64
    DevBuf cand_count;  // [qb] uint32 + [qb] uint64 bounds
    DevBuf out_pack;    // device result staging for host outputs: [items | scores | counts | retry]
    PinBuf pin_in;      // pinned staging: queries (+ subset) on the way in
    PinBuf pin_out;     // pinned staging: packed results on the way out
    cudaEvent_t ev_pin_in = nullptr;  // completion of the last H2D that read pin_in
    bool pin_in_busy = false;
    DevBuf staging;     // append: source rows before conversion
    DevBuf mma_ws;      // tensor-core path workspace
    DevBuf retry;       // int32 [n_queries] flags of the last tensor-core search + [1] running total at the end
    int retry_capacity = 0;      // queries the flag array is sized for
    // float32 indexes: the rows as two fp16 planes for the tensor-core path (built lazily,
    // extended on append); split_flag[0] = a corpus value left the fp16 range (sticky),
    // split_flag[1] = a query value did (per search)
    DevBuf split_hi, split_lo, split_flag;
    int64_t split_rows = 0;      // rows [0, split_rows) of the planes are current
    int64_t split_cap = 0;
    bool last_split = false;     // the last tensor-core search used the planes
    int pending_queries = 0;     // > 0: a TAV_DEFER_RETRY search awaits tav_finish_search

    // timing of the last search
    cudaEvent_t ev_total[2] = {nullptr, nullptr};
    cudaEvent_t ev_chunk[kMaxTimedChunks][2];
    int ev_kind[kMaxTimedChunks];  // 0 dominant kernel, 1 sample pass, 2 auxiliary
    int timed_chunks = 0;
    int launches = 0;
    int path = 0;
    bool timing_valid = false;
    bool timing_on = false;  // record CUDA events around kernels (tav_set_timing)
};

// The pinned input staging may still be the source of an in-flight H2D copy when the previous
// search returned without synchronising (device outputs): wait for that copy before reuse.
static cudaError_t pin_in_acquire(tav_index* ix, size_t bytes) {
    if (ix->pin_in_busy) {
        cudaError_t e = cudaEventSynchronize(ix->ev_pin_in);
        if (e != cudaSuccess) return e;
        ix->pin_in_busy = false;
    }
    return ix->pin_in.ensure(bytes);
}

static int set_device(const tav_index* ix) {
    TAV_CUDA(cudaSetDevice(ix->device));
    return TAV_OK;
}

extern "C" {

int tav_abi_version(void) { return TAV_ABI_VERSION; }

const char* tav_last_error(void) { return g_error; }

int tav_device_count(int* out_count) {
    if (!out_count) return TAV_ERR_INVALID;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        n = 0;
    }
    *out_count = n;
    return TAV_OK;
}

int tav_create(int device, int dim, int store_dtype, int index_flags, int64_t reserve_rows,
               tav_index** out) {
    if (!out || dim < 0 || !dtype_ok(store_dtype) || reserve_rows < 0) {
        set_error("tav_create: invalid argument");
        return TAV_ERR_INVALID;
    }
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        set_error("tav_create: no CUDA device available (%s); libtavec has no CPU fallback",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return TAV_ERR_CUDA;
    }
    if (device < 0 || device >= n) {
        set_error("tav_create: device %d out of range (have %d)", device, n);
        return TAV_ERR_INVALID;
    }
    TAV_CUDA(cudaSetDevice(device));
    tav_index* ix = new (std::nothrow) tav_index();
    if (!ix) return TAV_ERR_OOM;
    ix->device = device;
    ix->dim = dim;
    ix->dtype = store_dtype;
    ix->flags = index_flags;
    for (auto& pr : ix->ev_chunk) pr[0] = pr[1] = nullptr;
    cudaError_t ce = cudaEventCreate(&ix->ev_total[0]);
    if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&ix->ev_pin_in, cudaEventDisableTiming);
    if (ce == cudaSuccess) ce = cudaEventCreate(&ix->ev_total[1]);
    for (int i = 0; ce == cudaSuccess && i < kMaxTimedChunks; ++i) {
        ce = cudaEventCreate(&ix->ev_chunk[i][0]);
        if (ce == cudaSuccess) ce = cudaEventCreate(&ix->ev_chunk[i][1]);
    }
    if (ce != cudaSuccess) {
        set_error("tav_create: stream/event creation failed: %s", cudaGetErrorString(ce));
        tav_destroy(ix);
        return TAV_ERR_CUDA;
    }
    *out = ix;
    if (reserve_rows > 0 && dim > 0) {
        int rc = tav_reserve(ix, reserve_rows);
        if (rc != TAV_OK) {
            tav_destroy(ix);
            *out = nullptr;
            return rc;
        }
    }
    return TAV_OK;
}

int tav_destroy(tav_index* ix) {
    if (!ix) return TAV_OK;
    cudaSetDevice(ix->device);
    cudaDeviceSynchronize();  // searches may still be in flight on the caller's streams
    if (ix->rows && !ix->adopted) cudaFree(ix->rows);
    for (DevBuf* b : {&ix->queries, &ix->subset, &ix->cand_keys, &ix->cand_count, &ix->out_pack,
                      &ix->staging, &ix->mma_ws, &ix->retry, &ix->split_hi, &ix->split_lo, &ix->split_flag})
        b->release();
    ix->pin_in.release();
    ix->pin_out.release();
    for (auto& ev : ix->ev_total)
        if (ev) cudaEventDestroy(ev);
    if (ix->ev_pin_in) cudaEventDestroy(ix->ev_pin_in);
    for (auto& pr : ix->ev_chunk)
        for (auto& ev : pr)
            if (ev) cudaEventDestroy(ev);
    delete ix;
    return TAV_OK;
}

int tav_clear(tav_index* ix) {
    if (!ix) return TAV_ERR_INVALID;
    if (ix->adopted) {
        ix->rows = nullptr;
        ix->adopted = false;
        ix->capacity = 0;
    }
    ix->size = 0;
    ix->split_rows = 0;
    return TAV_OK;
}

int tav_reserve(tav_index* ix, int64_t rows) {
    if (!ix || rows < 0) return TAV_ERR_INVALID;
    if (ix->adopted) {
        set_error("tav_reserve: index uses adopted device memory");
        return TAV_ERR_STATE;
    }
    if (rows <= ix->capacity) return TAV_OK;
    if (ix->dim <= 0) {
        set_error("tav_reserve: embedding size not known yet");
        return TAV_ERR_STATE;
    }
    if (int rc = set_device(ix)) return rc;
    const size_t row_bytes = static_cast<size_t>(ix->dim) * dtype_size(ix->dtype);
    void* fresh = nullptr;
    TAV_CUDA(cudaMalloc(&fresh, std::max<size_t>(static_cast<size_t>(rows) * row_bytes, 256)));
    if (ix->size > 0) {
        // in-order with whatever was enqueued on the index stream; then make it visible to all
        cudaError_t e = cudaMemcpy(fresh, ix->rows, static_cast<size_t>(ix->size) * row_bytes,
                                   cudaMemcpyDeviceToDevice);
        if (e != cudaSuccess) {
            cudaFree(fresh);
            set_error("tav_reserve: copy failed: %s", cudaGetErrorString(e));
            return TAV_ERR_CUDA;
        }
    }
    if (ix->rows) {
        cudaDeviceSynchronize();
        cudaFree(ix->rows);
    }
    ix->rows = fresh;
    ix->capacity = rows;
    return TAV_OK;
}

int tav_append(tav_index* ix, const void* rows, int64_t n, int dim, int src_dtype,
               int src_on_device, void* stream) {
    if (!ix || n < 0 || dim <= 0 || !dtype_ok(src_dtype) || (n > 0 && !rows)) {
        set_error("tav_append: invalid argument");
        return TAV_ERR_INVALID;
    }
    if (ix->adopted) {
        set_error("tav_append: index uses adopted device memory");
        return TAV_ERR_STATE;
    }
    if (ix->dim == 0) ix->dim = dim;  // first append fixes the width (vectorbase.py:119-121)
    if (dim != ix->dim) {
        set_error("Embedding size mismatch: expected %d, got %d", ix->dim, dim);
        return TAV_ERR_INVALID;
    }
    if (n == 0) return TAV_OK;
    if (int rc = set_device(ix)) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (ix->size + n > ix->capacity) {
        int64_t want = std::max<int64_t>(ix->size + n, ix->capacity * 2);
        want = std::max<int64_t>(want, 1024);
        TAV_CUDA(cudaStreamSynchronize(s));
        if (int rc = tav_reserve(ix, want)) return rc;
    }
    const size_t dst_row = static_cast<size_t>(ix->dim) * dtype_size(ix->dtype);
    const size_t src_row = static_cast<size_t>(ix->dim) * dtype_size(src_dtype);
    char* dst = static_cast<char*>(ix->rows) + static_cast<size_t>(ix->size) * dst_row;
    const bool plain = (src_dtype == ix->dtype) && !(ix->flags & TAV_NORMALIZE);
    if (plain) {
        TAV_CUDA(cudaMemcpyAsync(dst, rows, static_cast<size_t>(n) * src_row,
                                 src_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s));
    } else if (src_on_device) {
        TAV_CUDA(launch_convert(rows, src_dtype, dst, ix->dtype, n, ix->dim,
                                (ix->flags & TAV_NORMALIZE) ? 1 : 0, s));
    } else {
        // host source needing conversion: stage through a bounded device buffer, chunk by chunk
        const int64_t chunk_rows = std::max<int64_t>(1, (64ll << 20) / static_cast<int64_t>(src_row));
        TAV_CUDA(ix->staging.ensure(static_cast<size_t>(std::min(n, chunk_rows)) * src_row));
        for (int64_t done = 0; done < n; done += chunk_rows) {
            const int64_t m = std::min(chunk_rows, n - done);
            TAV_CUDA(cudaMemcpyAsync(ix->staging.p, static_cast<const char*>(rows) + done * src_row,
                                     static_cast<size_t>(m) * src_row, cudaMemcpyHostToDevice, s));
            TAV_CUDA(launch_convert(ix->staging.p, src_dtype, dst + done * dst_row, ix->dtype, m,
                                    ix->dim, (ix->flags & TAV_NORMALIZE) ? 1 : 0, s));
        }
    }
    if (!src_on_device) TAV_CUDA(cudaStreamSynchronize(s));  // the host buffer may be reused
    ix->size += n;
    return TAV_OK;
}

int tav_adopt_device(tav_index* ix, void* device_rows, int64_t n, int dim) {
    if (!ix || n < 0 || dim <= 0 || (n > 0 && !device_rows)) return TAV_ERR_INVALID;
    if (ix->flags & TAV_NORMALIZE) {
        set_error("tav_adopt_device: not available on a TAV_NORMALIZE index (rows are used as is)");
        return TAV_ERR_STATE;
    }
    if (ix->dim != 0 && ix->dim != dim) {
        set_error("Embedding size mismatch: expected %d, got %d", ix->dim, dim);
        return TAV_ERR_INVALID;
    }
    if (reinterpret_cast<uintptr_t>(device_rows) % 16 != 0) {
        set_error("tav_adopt_device: pointer must be 16-byte aligned");
        return TAV_ERR_INVALID;
    }
    if (ix->rows && !ix->adopted) {
        cudaSetDevice(ix->device);
        cudaDeviceSynchronize();
        cudaFree(ix->rows);
    }
    ix->dim = dim;
    ix->rows = device_rows;
    ix->split_rows = 0;
    ix->adopted = true;
    ix->size = n;
    ix->capacity = n;
    return TAV_OK;
}

int64_t tav_size(const tav_index* ix) { return ix ? ix->size : 0; }
int tav_dim(const tav_index* ix) { return ix ? ix->dim : 0; }
int tav_store_dtype(const tav_index* ix) { return ix ? ix->dtype : -1; }
int tav_device(const tav_index* ix) { return ix ? ix->device : -1; }

int tav_read_rows(tav_index* ix, int64_t first, int64_t n, float* out_host, void* stream) {
    if (!ix || n < 0 || (n > 0 && !out_host)) return TAV_ERR_INVALID;
    if (first < 0 || first + n > ix->size) {
        set_error("tav_read_rows: rows [%lld, %lld) out of range (size %lld)", (long long)first,
                  (long long)(first + n), (long long)ix->size);
        return TAV_ERR_RANGE;
    }
    if (n == 0) return TAV_OK;
    if (int rc = set_device(ix)) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t row = static_cast<size_t>(ix->dim) * dtype_size(ix->dtype);
    const char* src = static_cast<const char*>(ix->rows) + static_cast<size_t>(first) * row;
    if (ix->dtype == TAV_F32) {
        TAV_CUDA(cudaMemcpyAsync(out_host, src, static_cast<size_t>(n) * row, cudaMemcpyDeviceToHost, s));
    } else {
        const size_t bytes = static_cast<size_t>(n) * ix->dim * sizeof(float);
        TAV_CUDA(ix->staging.ensure(bytes));
        TAV_CUDA(launch_convert(src, ix->dtype, ix->staging.p, TAV_F32, n, ix->dim, 0, s));
        TAV_CUDA(cudaMemcpyAsync(out_host, ix->staging.p, bytes, cudaMemcpyDeviceToHost, s));
    }
    TAV_CUDA(cudaStreamSynchronize(s));
    return TAV_OK;
}

// Row-scan search of queries [q0, q0+nq) (device float32), any k: passes of <= kPassK hits.
static int scan_search(tav_index* ix, const float* d_queries, int nq_total, int k, float floor_score,
                       const int64_t* d_subset, int64_t n_scan, int64_t item_offset,
                       int64_t* d_items, float* d_scores, int32_t* d_counts,
                       const int32_t* only_flagged, cudaStream_t s) {
    (void)only_flagged;
    const int pass_k = std::min(k, kPassK);
    int qb = scan_max_queries(ix->dim, pass_k);
    if (qb < 1) {
        set_error("tav_search: embedding size %d too large for the row-scan kernel", ix->dim);
        return TAV_ERR_INVALID;
    }
    qb = std::min(qb, nq_total);
    // round qb down to a power of two (kernel instantiations 1/2/4/8)
    while (qb & (qb - 1)) qb &= qb - 1;
    const int grid = scan_grid(ix->device, ix->dtype, ix->dim, qb, pass_k, n_scan);
    const int cand_stride = grid * pass_k;
    TAV_CUDA(ix->cand_keys.ensure(static_cast<size_t>(qb) * cand_stride * sizeof(uint64_t)));
    if (ix->cand_count.bytes < 8 * (sizeof(uint32_t) + sizeof(uint64_t)) + 64) {
        // sized for the largest pass (8 queries) once, zeroed once; select_kernel re-zeroes the counters
        TAV_CUDA(ix->cand_count.ensure(8 * (sizeof(uint32_t) + sizeof(uint64_t)) + 64));
        TAV_CUDA(cudaMemsetAsync(ix->cand_count.p, 0, ix->cand_count.bytes, s));
    }
    uint64_t* d_bound = static_cast<uint64_t*>(ix->cand_count.p);
    uint32_t* d_count = reinterpret_cast<uint32_t*>(d_bound + 8);
    const int n_pass = (k + pass_k - 1) / pass_k;

    for (int q0 = 0; q0 < nq_total; q0 += qb) {
        const int nq = std::min(qb, nq_total - q0);
        for (int pass = 0; pass < n_pass; ++pass) {
            const int kk = std::min(pass_k, k - pass * pass_k);
            ScanArgs a{};
            a.corpus = ix->rows;
            a.dtype = ix->dtype;
            a.n_corpus = ix->size;
            a.dim = ix->dim;
            a.subset = d_subset;
            a.n_scan = n_scan;
            a.queries = d_queries + static_cast<size_t>(q0) * ix->dim;
            a.nq = nq;
            a.floor_score = floor_score;
            a.bound = pass > 0 ? d_bound : nullptr;
            a.k = kk;
            a.cand_keys = static_cast<uint64_t*>(ix->cand_keys.p);
            a.cand_stride = cand_stride;
            a.cand_count = d_count;
            a.grid = grid;
            const bool timed = ix->timing_on && ix->timed_chunks < kMaxTimedChunks;
            if (timed) TAV_CUDA(cudaEventRecord(ix->ev_chunk[ix->timed_chunks][0], s));
            TAV_CUDA(launch_scan(a, s));
            if (timed) {
                ix->ev_kind[ix->timed_chunks] = 0;
                TAV_CUDA(cudaEventRecord(ix->ev_chunk[ix->timed_chunks++][1], s));
            }
            SelectArgs sel{};
            sel.cand_keys = a.cand_keys;
            sel.cand_stride = cand_stride;
            sel.cand_count = d_count;
            sel.cand_count_reset = d_count;
            sel.nq = nq;
            sel.k = kk;
            sel.out_stride = k;
            sel.out_offset = pass * pass_k;
            sel.subset = d_subset;
            sel.item_offset = item_offset;
            sel.out_items = d_items + static_cast<size_t>(q0) * k;
            sel.out_scores = d_scores + static_cast<size_t>(q0) * k;
            sel.out_counts = d_counts + q0;
            sel.bound_out = n_pass > 1 ? d_bound : nullptr;
            sel.accumulate = pass > 0;
            TAV_CUDA(launch_select(sel, s));
            ix->launches += 2;
        }
    }
    return TAV_OK;
}

}  // extern "C"

// float32 index -> fp16 hi/lo planes covering rows [0, size); returns TAV_ERR_OOM when they do not fit
static int ensure_split_planes(tav_index* ix, cudaStream_t s) {
    const size_t plane_row = static_cast<size_t>(ix->dim) * 2;
    if (ix->split_flag.bytes == 0) {
        TAV_CUDA(ix->split_flag.ensure(2 * sizeof(int)));
        TAV_CUDA(cudaMemsetAsync(ix->split_flag.p, 0, 2 * sizeof(int), s));
    }
    if (ix->size > ix->split_cap) {
        const int64_t cap = std::max<int64_t>(ix->size, ix->capacity);
        ix->split_hi.release();
        ix->split_lo.release();
        ix->split_cap = 0;
        ix->split_rows = 0;
        cudaError_t e1 = ix->split_hi.ensure(static_cast<size_t>(cap) * plane_row);
        cudaError_t e2 = e1 == cudaSuccess ? ix->split_lo.ensure(static_cast<size_t>(cap) * plane_row) : e1;
        if (e2 != cudaSuccess) {
            cudaGetLastError();
            ix->split_hi.release();
            ix->split_lo.release();
            set_error("not enough device memory for the fp16 planes of the float32 index");
            return TAV_ERR_OOM;
        }
        ix->split_cap = cap;
        TAV_CUDA(cudaMemsetAsync(ix->split_flag.p, 0, sizeof(int), s));
    }
    if (ix->split_rows < ix->size) {
        const int64_t first = ix->split_rows, n = ix->size - first;
        TAV_CUDA(launch_split_rows(static_cast<const float*>(ix->rows) + first * ix->dim,
                                   static_cast<char*>(ix->split_hi.p) + static_cast<size_t>(first) * plane_row,
                                   static_cast<char*>(ix->split_lo.p) + static_cast<size_t>(first) * plane_row, n,
                                   ix->dim, static_cast<int*>(ix->split_flag.p), s));
        ix->split_rows = ix->size;
        ix->launches += 1;
    }
    return TAV_OK;
}

// Synchronising tail of a tensor-core search: read the retry flags, redo flagged queries exactly.
static int resolve_retries(tav_index* ix, const float* d_queries, int n_queries, int k, float min_score,
                           int64_t item_offset, int64_t* d_items, float* d_scores, int32_t* d_counts,
                           cudaStream_t s, int* redone) {
    std::vector<int32_t> host(static_cast<size_t>(n_queries) + 1);
    int32_t* flags = static_cast<int32_t*>(ix->retry.p);
    TAV_CUDA(cudaMemcpyAsync(host.data(), flags, static_cast<size_t>(n_queries) * sizeof(int32_t),
                             cudaMemcpyDeviceToHost, s));
    TAV_CUDA(cudaMemcpyAsync(&host[n_queries], flags + ix->retry_capacity, sizeof(int32_t),
                             cudaMemcpyDeviceToHost, s));
    int out_of_range[2] = {0, 0};  // split form: a corpus / query value beyond the fp16 range
    if (ix->last_split)
        TAV_CUDA(cudaMemcpyAsync(out_of_range, ix->split_flag.p, sizeof(out_of_range), cudaMemcpyDeviceToHost, s));
    TAV_CUDA(cudaStreamSynchronize(s));
    ix->pending_queries = 0;
    const bool redo_all = out_of_range[0] != 0 || out_of_range[1] != 0;
    const int total = host[n_queries];
    int n_flagged = 0;
    for (int q = 0; q < n_queries; ++q) {
        if (host[q]) ++n_flagged;
        if (!host[q] && !redo_all) continue;
        int rc = scan_search(ix, d_queries + static_cast<size_t>(q) * ix->dim, 1, k, min_score, nullptr,
                             ix->size, item_offset, d_items + static_cast<size_t>(q) * k,
                             d_scores + static_cast<size_t>(q) * k, d_counts + q, nullptr, s);
        if (rc != TAV_OK) return rc;
    }
    if (total != 0) TAV_CUDA(cudaMemsetAsync(flags + ix->retry_capacity, 0, sizeof(int32_t), s));
    if (redone) *redone = redo_all ? n_queries : n_flagged;
    if (total != n_flagged) {
        set_error("%d queries of earlier deferred searches needed the exact fallback but were never finished",
                  total - n_flagged);
        return TAV_ERR_STATE;
    }
    return TAV_OK;
}

extern "C" {

int tav_finish_search(tav_index* ix, const float* queries_device, int n_queries, int k, float min_score,
                      int64_t item_offset, int64_t* out_items, float* out_scores, int32_t* out_counts,
                      void* stream, int* redone) {
    if (!ix || n_queries < 0) return TAV_ERR_INVALID;
    if (redone) *redone = 0;
    if (ix->pending_queries == 0) return TAV_OK;  // nothing deferred (row-scan path, or already finished)
    if (n_queries != ix->pending_queries || !queries_device || !out_items || !out_scores || !out_counts || k < 1) {
        set_error("tav_finish_search: arguments must repeat the deferred tav_search call");
        return TAV_ERR_INVALID;
    }
    if (int rc = set_device(ix)) return rc;
    return resolve_retries(ix, queries_device, n_queries, k, min_score, item_offset, out_items, out_scores,
                           out_counts, static_cast<cudaStream_t>(stream), redone);
}

int tav_search(tav_index* ix, const float* queries, int n_queries, int k, float min_score,
               int flags, const int64_t* subset, int64_t subset_len, int64_t item_offset,
               int64_t* out_items, float* out_scores, int32_t* out_counts, void* stream) {
    if (!ix || n_queries < 0 || k < 1 || (n_queries > 0 && (!queries || !out_items || !out_scores || !out_counts))) {
        set_error("tav_search: invalid argument (k must be >= 1)");
        return TAV_ERR_INVALID;
    }
    if ((subset && subset_len < 0) || (!subset && subset_len != 0)) {
        set_error("tav_search: subset / subset_len mismatch");
        return TAV_ERR_INVALID;
    }
    if (n_queries == 0) return TAV_OK;
    if (int rc = set_device(ix)) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const bool q_dev = flags & TAV_QUERIES_ON_DEVICE, o_dev = flags & TAV_OUTPUTS_ON_DEVICE;
    const size_t nk = static_cast<size_t>(n_queries) * k;

    int64_t* d_items = out_items;
    float* d_scores = out_scores;
    int32_t* d_counts = out_counts;
    // host outputs: results are packed [items | scores | counts] in one device buffer so that a
    // single D2H copy (into pinned staging) brings them back
    const size_t off_scores = nk * sizeof(int64_t);
    const size_t off_counts = off_scores + ((nk * sizeof(float) + 7) & ~size_t(7));
    const size_t pack_bytes = off_counts + ((static_cast<size_t>(n_queries) * sizeof(int32_t) + 7) & ~size_t(7));
    // Small result sets are written by the kernels straight into the pinned host staging (zero
    // copy over PCIe: no D2H memcpy call on the single-lookup latency path).
    const bool zero_copy_out = !o_dev && pack_bytes <= kZeroCopyOutLimit;
    if (zero_copy_out) {
        TAV_CUDA(ix->pin_out.ensure(pack_bytes));
        char* base = static_cast<char*>(ix->pin_out.p);
        d_items = reinterpret_cast<int64_t*>(base);
        d_scores = reinterpret_cast<float*>(base + off_scores);
        d_counts = reinterpret_cast<int32_t*>(base + off_counts);
    } else if (!o_dev) {
        TAV_CUDA(ix->out_pack.ensure(pack_bytes));
        char* base = static_cast<char*>(ix->out_pack.p);
        d_items = reinterpret_cast<int64_t*>(base);
        d_scores = reinterpret_cast<float*>(base + off_scores);
        d_counts = reinterpret_cast<int32_t*>(base + off_counts);
    }

    const int64_t n_scan = subset ? subset_len : ix->size;
    ix->timed_chunks = 0;
    ix->launches = 0;
    ix->path = 0;
    ix->timing_valid = false;

    if (n_scan == 0 || ix->size == 0 || ix->dim == 0) {
        // empty corpus / empty subset: no hits (vectorbase.py:174-175, :214-215)
        if (o_dev) TAV_CUDA(cudaMemsetAsync(d_counts, 0, static_cast<size_t>(n_queries) * sizeof(int32_t), s));
        else memset(out_counts, 0, static_cast<size_t>(n_queries) * sizeof(int32_t));
        return TAV_OK;
    }
    if (n_scan > 0xFFFFFFFFll) {
        set_error("tav_search: more than 2^32 rows per index are not supported; shard the corpus");
        return TAV_ERR_INVALID;
    }

    // subset ordinals: validate on the host (numpy raises IndexError), then upload
    const int64_t* d_subset = nullptr;
    if (subset) {
        for (int64_t i = 0; i < subset_len; ++i) {
            if (subset[i] < -ix->size || subset[i] >= ix->size) {
                set_error("index %lld is out of bounds for axis 0 with size %lld",
                          (long long)subset[i], (long long)ix->size);
                return TAV_ERR_RANGE;
            }
        }
        const size_t sub_bytes = static_cast<size_t>(subset_len) * sizeof(int64_t);
        TAV_CUDA(ix->subset.ensure(sub_bytes));
        const void* src = subset;
        if (sub_bytes <= kPinnedStageLimit) {
            TAV_CUDA(pin_in_acquire(ix, ((sub_bytes + 15) & ~size_t(15)) +
                                            static_cast<size_t>(n_queries) * ix->dim * sizeof(float)));
            memcpy(ix->pin_in.p, subset, sub_bytes);
            src = ix->pin_in.p;
        }
        TAV_CUDA(cudaMemcpyAsync(ix->subset.p, src, sub_bytes, cudaMemcpyHostToDevice, s));
        d_subset = static_cast<const int64_t*>(ix->subset.p);
    }

    if (ix->timing_on) TAV_CUDA(cudaEventRecord(ix->ev_total[0], s));

    // queries -> device float32 (normalised in place when the index is TAV_NORMALIZE)
    const float* d_queries = queries;
    const size_t q_bytes = static_cast<size_t>(n_queries) * ix->dim * sizeof(float);
    if (!q_dev || (ix->flags & TAV_NORMALIZE)) {
        TAV_CUDA(ix->queries.ensure(q_bytes));
        if (ix->flags & TAV_NORMALIZE) {
            const void* src = queries;
            if (!q_dev) {
                TAV_CUDA(ix->staging.ensure(q_bytes));
                TAV_CUDA(cudaMemcpyAsync(ix->staging.p, queries, q_bytes, cudaMemcpyHostToDevice, s));
                src = ix->staging.p;
            }
            TAV_CUDA(launch_convert(src, TAV_F32, ix->queries.p, TAV_F32, n_queries, ix->dim, 1, s));
            ix->launches += 1;
        } else {
            const void* src = queries;
            if (q_bytes <= kPinnedStageLimit) {
                // via pinned staging: a pageable source would make the copy synchronous
                const size_t sub_bytes = subset ? static_cast<size_t>(subset_len) * sizeof(int64_t) : 0;
                const size_t sub_off = sub_bytes <= kPinnedStageLimit ? ((sub_bytes + 15) & ~size_t(15)) : 0;
                TAV_CUDA(pin_in_acquire(ix, sub_off + q_bytes));
                memcpy(static_cast<char*>(ix->pin_in.p) + sub_off, queries, q_bytes);
                src = static_cast<char*>(ix->pin_in.p) + sub_off;
            }
            TAV_CUDA(cudaMemcpyAsync(ix->queries.p, src, q_bytes, cudaMemcpyHostToDevice, s));
        }
        d_queries = static_cast<const float*>(ix->queries.p);
    }

    if (o_dev && (!q_dev || subset)) {  // no synchronisation at the end of this call
        TAV_CUDA(cudaEventRecord(ix->ev_pin_in, s));
        ix->pin_in_busy = true;
    }

    // path choice: tensor cores for batches on 16-bit storage, row scan otherwise
    bool use_mma = false, use_split = false;
    const bool mma_able = mma_supported(ix->dtype, ix->dim) || (ix->dtype == TAV_F32 && mma_split_supported(ix->dim));
    if (!(flags & TAV_FORCE_SCAN) && !subset && mma_able && k <= kPassK) {
        use_mma = (flags & TAV_FORCE_MMA) || (n_queries >= 16 && ix->size >= 4096);
        use_split = use_mma && ix->dtype == TAV_F32;
    }
    if ((flags & TAV_FORCE_MMA) && !use_mma) {
        set_error("tav_search: TAV_FORCE_MMA needs dim %% 8 == 0, no subset, k <= %d", kPassK);
        return TAV_ERR_INVALID;
    }
    if (use_split) {
        const int rc = ensure_split_planes(ix, s);
        if (rc == TAV_ERR_OOM && !(flags & TAV_FORCE_MMA)) {
            use_mma = use_split = false;  // a speed choice, not a correctness one: the exact row scan serves it
        } else if (rc != TAV_OK) {
            return rc;
        }
    }

    if (use_mma) {
        ix->path = use_split ? 3 : 2;
        ix->last_split = use_split;
        MmaArgs m{};
        m.device = ix->device;
        m.corpus = use_split ? ix->split_hi.p : ix->rows;
        m.corpus_lo = use_split ? ix->split_lo.p : nullptr;
        m.split = use_split ? 1 : 0;
        m.split_overflow = use_split ? static_cast<int*>(ix->split_flag.p) + 1 : nullptr;
        if (use_split) TAV_CUDA(cudaMemsetAsync(m.split_overflow, 0, sizeof(int), s));
        m.dtype = ix->dtype;
        m.n_corpus = ix->size;
        m.dim = ix->dim;
        m.queries = d_queries;
        m.nq = n_queries;
        m.floor_score = min_score;
        m.k = k;
        m.item_offset = item_offset;
        m.out_items = d_items;
        m.out_scores = d_scores;
        m.out_counts = d_counts;
        if (n_queries > ix->retry_capacity) {
            // [flags x capacity | total]; the running total lives right after the flags
            int32_t carried = 0;
            if (ix->retry.p)
                TAV_CUDA(cudaMemcpy(&carried, static_cast<int32_t*>(ix->retry.p) + ix->retry_capacity,
                                    sizeof(int32_t), cudaMemcpyDeviceToHost));
            const int cap = std::max(n_queries, 1024);
            TAV_CUDA(ix->retry.ensure((static_cast<size_t>(cap) + 1) * sizeof(int32_t)));
            TAV_CUDA(cudaMemcpy(static_cast<int32_t*>(ix->retry.p) + cap, &carried, sizeof(int32_t),
                                cudaMemcpyHostToDevice));
            ix->retry_capacity = cap;
        }
        m.retry_flags = static_cast<int32_t*>(ix->retry.p);
        m.retry_total = m.retry_flags + ix->retry_capacity;
        m.ev = ix->timing_on ? ix->ev_chunk : nullptr;
        m.ev_kind = ix->ev_kind;
        m.ev_max = kMaxTimedChunks;
        int ev_used = 0;
        m.ev_used = &ev_used;
        const size_t ws = mma_workspace_bytes(m);
        TAV_CUDA(ix->mma_ws.ensure(ws));
        int launches = 0;
        TAV_CUDA(launch_mma_search(m, ix->mma_ws.p, ws, s, &launches));
        ix->timed_chunks = ev_used;
        ix->launches += launches;
        // Queries the sampled admission threshold could not settle (fewer than k admitted rows
        // although rows were cut, or candidate overflow) are redone exactly by the row scan —
        // now, or in tav_finish_search when the caller defers the (synchronising) check.
        if ((flags & TAV_DEFER_RETRY) && o_dev && q_dev) {
            ix->pending_queries = n_queries;
        } else {
            int redone = 0;
            int rc = resolve_retries(ix, d_queries, n_queries, k, min_score, item_offset, d_items, d_scores,
                                     d_counts, s, &redone);
            if (rc != TAV_OK) return rc;
        }
    } else {
        ix->path = 1;
        int rc = scan_search(ix, d_queries, n_queries, k, min_score, d_subset, n_scan, item_offset,
                             d_items, d_scores, d_counts, nullptr, s);
        if (rc != TAV_OK) return rc;
    }
    if (ix->timing_on) TAV_CUDA(cudaEventRecord(ix->ev_total[1], s));
    ix->timing_valid = true;

    if (zero_copy_out) {
        TAV_CUDA(cudaStreamSynchronize(s));
        const char* h = static_cast<const char*>(ix->pin_out.p);
        memcpy(out_items, h, nk * sizeof(int64_t));
        memcpy(out_scores, h + off_scores, nk * sizeof(float));
        memcpy(out_counts, h + off_counts, static_cast<size_t>(n_queries) * sizeof(int32_t));
    } else if (!o_dev) {
        if (pack_bytes <= kPinnedStageLimit) {
            TAV_CUDA(ix->pin_out.ensure(pack_bytes));
            TAV_CUDA(cudaMemcpyAsync(ix->pin_out.p, ix->out_pack.p, pack_bytes, cudaMemcpyDeviceToHost, s));
            TAV_CUDA(cudaStreamSynchronize(s));
            const char* h = static_cast<const char*>(ix->pin_out.p);
            memcpy(out_items, h, nk * sizeof(int64_t));
            memcpy(out_scores, h + off_scores, nk * sizeof(float));
            memcpy(out_counts, h + off_counts, static_cast<size_t>(n_queries) * sizeof(int32_t));
        } else {
            TAV_CUDA(cudaMemcpyAsync(out_items, d_items, nk * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
            TAV_CUDA(cudaMemcpyAsync(out_scores, d_scores, nk * sizeof(float), cudaMemcpyDeviceToHost, s));
            TAV_CUDA(cudaMemcpyAsync(out_counts, d_counts, static_cast<size_t>(n_queries) * sizeof(int32_t),
                                     cudaMemcpyDeviceToHost, s));
            TAV_CUDA(cudaStreamSynchronize(s));
        }
    }
    return TAV_OK;
}

int tav_mma_scores(tav_index* ix, const float* queries, int n_queries, int flags, float* out_device,
                   void* stream) {
    if (!ix || n_queries < 1 || !queries || !out_device) return TAV_ERR_INVALID;
    const bool split = ix->dtype == TAV_F32;
    if (ix->size == 0 || !(split ? mma_split_supported(ix->dim) : mma_supported(ix->dtype, ix->dim))) {
        set_error("tav_mma_scores: needs a non-empty index with dim %% 8 == 0");
        return TAV_ERR_INVALID;
    }
    if (int rc = set_device(ix)) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const float* d_queries = queries;
    if (!(flags & TAV_QUERIES_ON_DEVICE)) {
        const size_t q_bytes = static_cast<size_t>(n_queries) * ix->dim * sizeof(float);
        TAV_CUDA(ix->queries.ensure(q_bytes));
        TAV_CUDA(cudaMemcpyAsync(ix->queries.p, queries, q_bytes, cudaMemcpyHostToDevice, s));
        d_queries = static_cast<const float*>(ix->queries.p);
    }
    if (split)
        if (int rc = ensure_split_planes(ix, s)) return rc;
    MmaArgs m{};
    m.device = ix->device;
    m.corpus = split ? ix->split_hi.p : ix->rows;
    m.corpus_lo = split ? ix->split_lo.p : nullptr;
    m.split = split ? 1 : 0;
    m.split_overflow = split ? static_cast<int*>(ix->split_flag.p) + 1 : nullptr;
    m.dtype = ix->dtype;
    m.n_corpus = ix->size;
    m.dim = ix->dim;
    m.queries = d_queries;
    m.nq = n_queries;
    m.k = 1;
    const size_t ws = mma_workspace_bytes(m);
    TAV_CUDA(ix->mma_ws.ensure(ws));
    TAV_CUDA(launch_mma_dump(m, ix->mma_ws.p, ws, out_device, s));
    TAV_CUDA(cudaStreamSynchronize(s));
    return TAV_OK;
}

int tav_merge_topk(int device, int n_lists, int n_queries, int k, const int64_t* items,
                   const float* scores, const int32_t* counts, int64_t items_stride,
                   int64_t scores_stride, int64_t counts_stride, int64_t* out_items,
                   float* out_scores, int32_t* out_counts, void* stream) {
    if (n_lists < 1 || n_queries < 0 || k < 1 || !items || !scores || !counts || !out_items ||
        !out_scores || !out_counts) {
        set_error("tav_merge_topk: invalid argument");
        return TAV_ERR_INVALID;
    }
    if (static_cast<int64_t>(n_lists) * k > 0x7FFFFFFFll || k > kPassK * 4) {
        set_error("tav_merge_topk: n_lists * k too large");
        return TAV_ERR_INVALID;
    }
    if (n_queries == 0) return TAV_OK;
    TAV_CUDA(cudaSetDevice(device));
    if (items_stride < 0 || scores_stride < 0 || counts_stride < 0) return TAV_ERR_INVALID;
    TAV_CUDA(launch_merge(n_lists, n_queries, k, items, scores, counts, items_stride, scores_stride,
                          counts_stride, out_items, out_scores, out_counts,
                          static_cast<cudaStream_t>(stream)));
    return TAV_OK;
}

int tav_set_timing(tav_index* ix, int enabled) {
    if (!ix) return TAV_ERR_INVALID;
    ix->timing_on = enabled != 0;
    ix->timing_valid = false;
    return TAV_OK;
}

int tav_timing_breakdown(tav_index* ix, float* ms, int* kinds, int capacity, int* n) {
    if (!ix || !n || capacity < 0) return TAV_ERR_INVALID;
    if (!ix->timing_valid || !ix->timing_on) {
        *n = 0;
        return TAV_OK;
    }
    if (int rc = set_device(ix)) return rc;
    TAV_CUDA(cudaEventSynchronize(ix->ev_total[1]));
    *n = ix->timed_chunks;
    for (int i = 0; i < ix->timed_chunks && i < capacity; ++i) {
        float v = 0.0f;
        TAV_CUDA(cudaEventElapsedTime(&v, ix->ev_chunk[i][0], ix->ev_chunk[i][1]));
        if (ms) ms[i] = v;
        if (kinds) kinds[i] = ix->ev_kind[i];
    }
    return TAV_OK;
}

int tav_last_timing(tav_index* ix, float* scan_ms, float* total_ms, int* launches, int* path) {
    if (!ix) return TAV_ERR_INVALID;
    if (!ix->timing_valid) {
        set_error("tav_last_timing: no timed search on this index yet");
        return TAV_ERR_STATE;
    }
    if (!ix->timing_on) {  // path / launch count only
        if (scan_ms) *scan_ms = -1.0f;
        if (total_ms) *total_ms = -1.0f;
        if (launches) *launches = ix->launches;
        if (path) *path = ix->path;
        return TAV_OK;
    }
    if (int rc = set_device(ix)) return rc;
    TAV_CUDA(cudaEventSynchronize(ix->ev_total[1]));
    float total = 0.0f, scan = 0.0f;
    TAV_CUDA(cudaEventElapsedTime(&total, ix->ev_total[0], ix->ev_total[1]));
    for (int i = 0; i < ix->timed_chunks; ++i) {
        if (ix->ev_kind[i] != 0) continue;
        float ms = 0.0f;
        TAV_CUDA(cudaEventElapsedTime(&ms, ix->ev_chunk[i][0], ix->ev_chunk[i][1]));
        scan += ms;
    }
    if (scan_ms) *scan_ms = scan;
    if (total_ms) *total_ms = total;
    if (launches) *launches = ix->launches;
    if (path) *path = ix->path;
    return TAV_OK;
}

}  // extern "C"
