// tav_internal.h — launch interfaces between the translation units of libtavec.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/tavec.h"

namespace tav {

struct MergeSync;  // tav_common.cuh

// ---- row-scan path (tav_scan.cu) -------------------------------------------------------
struct ScanArgs {
    const void* corpus;       // [n_corpus, dim] storage dtype, row-major dense
    int dtype;                // tav_dtype
    int64_t n_corpus;
    int dim;
    const int64_t* subset;    // device, or nullptr = all rows
    int64_t n_scan;           // rows scanned (subset_len or n_corpus)
    const float* queries;     // device float32 [nq, dim]
    int nq;                   // 1..8 queries scored per pass over the rows
    float floor_score;        // (float)min_score
    const uint64_t* bound;    // per query: admit only keys < bound (multi-pass), or nullptr
    int k;                    // hits kept per query this pass (<= kPassK)
    uint64_t* cand_keys;      // [nq, cand_stride] per-query global candidate buffers
    int cand_stride;
    uint32_t* cand_count;     // [nq], zero on entry (the select kernel re-zeroes it)
    int grid;                 // CTAs to launch (cand_stride >= grid * k)
    const uint32_t* row_mask; // optional: bit r set = corpus row r may be returned (predicate pushdown)
    int ties_low;             // 1: equal scores -> LOWER position first (reference predicate path)
    // single-launch form (launch_scan1): the last CTA merges all survivors and writes the hits
    int subset_in_params;     // the subset ordinals travel in the kernel parameters
    int fused;
    uint32_t* fused_ticket;   // device counter, zero on entry (self-resetting)
    int64_t item_offset;
    int64_t* out_items;       // [nq, k]   (device, or mapped pinned host memory)
    float* out_scores;
    int32_t* out_counts;
    uint32_t* done_flag;      // mapped pinned host word set to done_seq once the hits are written, or nullptr
    uint32_t done_seq;
    unsigned long long* trace;  // diagnostic (TAV_TRACE=1): %globaltimer stamps of the single-launch form's phases
};
constexpr int kFusedSelectMax = 8192;   // survivors the last CTA of the single-launch form can merge
constexpr int kFusedSelOut = 1024;      // ... of which it sorts at most this many after the histogram selection
constexpr int kParamQuerySmall = 1024;  // floats of query carried in a 4 KB kernel-parameter blob
constexpr int kParamQueryBig = 3072;    // ... in the 28 KB blob (with up to kParamSubsetMax ordinals)
constexpr int kParamSubsetMax = 4096;
constexpr int kParamSubsetSmall = 1024; // ordinals in the 8 KB blob (with a query of <= kParamQuerySmall floats)
bool scan1_fits(int dim, int k, int64_t n_scan, int64_t subset_len, bool has_subset);
int scan1_grid(int device, int dim, int k, int64_t n_scan);
// q_host: float32 [dim] on the host; sub_host: int64 [n_scan] validated ordinals or nullptr
cudaError_t launch_scan1(const ScanArgs& a, const float* q_host, const int64_t* sub_host, cudaStream_t s);
int scan_max_queries(int dim, int k);            // how many queries one pass can take (smem)
int scan_grid(int device, int dtype, int dim, int nq, int k, int64_t n_scan);
cudaError_t launch_scan(const ScanArgs& a, cudaStream_t s);

struct SelectArgs {
    const uint64_t* cand_keys;   // [nq, cand_stride]
    int cand_stride;
    const uint32_t* cand_count;  // [nq]
    uint32_t* cand_count_reset;  // same array: zeroed by the kernel once consumed
    int nq;
    int k;                       // hits this pass
    int out_stride;              // row stride of out_items/out_scores (the caller's total k)
    int out_offset;              // column where this pass starts
    const int64_t* subset;       // device copy of the caller's ordinals, or nullptr
    int64_t item_offset;
    int64_t* out_items;          // already offset to the first query of this chunk
    float* out_scores;
    int32_t* out_counts;
    uint64_t* bound_out;         // [nq] next-pass bound (last key, 0 if exhausted), or nullptr
    int accumulate;              // counts += n instead of counts = n
    int ties_low;                // keys carry ~position (see ScanArgs)
};
cudaError_t launch_select(const SelectArgs& a, cudaStream_t s);

// in place: hits sorted by score -> first hit of every group (row_to_group[item - item_offset]), the
// reference's chunk -> message fold (storage/memory/messageindex.py:185-207)
cudaError_t launch_fold_groups(int n_queries, int k, const int32_t* row_to_group, int64_t n_rows,
                               int64_t item_offset, int64_t* items, float* scores, int32_t* counts,
                               cudaStream_t s);

cudaError_t launch_merge(int n_lists, int n_queries, int k, const int64_t* items,
                         const float* scores, const int32_t* counts, int64_t items_stride,
                         int64_t scores_stride, int64_t counts_stride, int64_t* out_items,
                         float* out_scores, int32_t* out_counts, cudaStream_t s, const MergeSync* sync = nullptr);

// rows [n, dim] of src dtype -> dst dtype (RNE), optionally L2-normalised per row (fp32 math)
cudaError_t launch_convert(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t n,
                           int dim, int normalize, cudaStream_t s);

// ---- tensor-core path (tav_mma.cu) -----------------------------------------------------
struct MmaPlan;  // opaque: tensor maps + workspace for one (index, batch shape)
bool mma_supported(int dtype, int dim);   // bf16 / fp16 storage
bool mma_split_supported(int dim);        // float32 storage carried as two fp16 planes
// float32 rows -> fp16 planes hi, lo with x ~= hi + lo / 2048; *overflow |= 1 if some |x| > fp16 range
cudaError_t launch_split_rows(const float* src, void* hi, void* lo, int64_t n, int dim, int* overflow,
                              cudaStream_t s);
// returns cudaSuccess and fills outputs exactly like scan+select; see tav_mma.cu
struct MmaArgs {
    int device;
    const void* corpus;    // storage rows; for split float32 data: the hi plane (fp16)
    const void* corpus_lo; // split only: the lo plane (fp16)
    int split;             // 1: float32 index searched through its two fp16 planes
    int* split_overflow;   // split only: device flag set when a value left the fp16 range
    int dtype;             // storage dtype of the index (TAV_F32 when split)
    int64_t n_corpus;
    int dim;
    const float* queries;  // device float32 [nq, dim]
    int nq;
    float floor_score;
    int k;
    int64_t item_offset;
    int64_t* out_items;    // device [nq, k]
    float* out_scores;
    int32_t* out_counts;
    int32_t* retry_flags;  // device [nq]: set to 1 for queries the caller must redo with the row scan
    int32_t* retry_total;  // device [1]: incremented once per flagged query (never reset by the kernels)
    int32_t* retry_total_host;  // the same counter in mapped pinned memory (read by the host without a copy), or nullptr
    int* split_overflow_host;   // split only: mapped pinned twin of split_overflow, or nullptr
    const uint32_t* row_mask;  // optional device bitmask over corpus rows (bit set = row may be returned)
    int no_ts;             // 1: never use the Q-stationary (queries in tensor memory) form
    cudaEvent_t (*ev)[2];  // optional event pairs, one recorded around every kernel launched
    int* ev_kind;          //   kind per pair: 0 = dominant (MAIN) kernel, 1 = sample pass, 2 = auxiliary
    int ev_max;
    int* ev_used;
    int ev_main_only;      // 1: record events only around the dominant (MAIN) kernel
};
constexpr int kMmaMaxQueries = 32768;  // queries per launch_mma_search call (512 chunks of >= 128 ... callers slab)
size_t mma_workspace_bytes(const MmaArgs& a);
cudaError_t launch_mma_search(const MmaArgs& a, void* workspace, size_t workspace_bytes,
                              cudaStream_t s, int* launches);
// verification aid: every raw dot product of the tensor-core path, out[nq, n_corpus] on the device
cudaError_t launch_mma_dump(const MmaArgs& a, void* workspace, size_t workspace_bytes, float* out,
                            cudaStream_t s);

void set_error(const char* fmt, ...);

}  // namespace tav

// library-internal (not in include/tavec.h): device address of the "queries flagged for the exact redo"
// counters of the index's most recent search — `*count` int32 words, 2 words apart — or nullptr when that
// search ran on the row-scan path (nothing to flag).  The sharded search ships their sum with the
// published candidate list so that every rank learns, without a second exchange, whether some rank will
// correct its candidates at finish.
extern "C" const int32_t* tav_internal_retry_totals(tav_index* ix, int* count);

namespace tav {

}  // namespace tav
