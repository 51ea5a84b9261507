// tav_mma.cu — the tensor-core path of libtavec: batched query x corpus similarity as a dense
// bf16/fp16 contraction on tcgen05 (fp32 accumulators in TMEM) fed by TMA tiles from HBM, with
// the score threshold and top-k candidate selection fused into the TMEM epilogue.
//
// Reference semantics (aitools/vectorbase.py:163-190, per query): x = dot(row, q) in float32;
// score = clip((x+1)/2, 0, 1); keep score >= min_score; k best by score.  Products of bf16/fp16
// values are exact in float32, so on storage-rounded inputs only the summation order differs
// from the reference's sgemv.
//
// Shape of one CTA (persistent, one per SM, 192 threads):
//   warp 0     TMA producer: per (corpus tile, query block, 64-wide K slice) loads the query
//              slice [128 x 64] and the corpus slice [256 x 64] into a 4-stage smem ring
//              (128-byte swizzle), completing on an mbarrier.
//   warp 1     MMA issuer: one elected thread issues tcgen05.mma.kind::f16 M=128 N=256 K=16
//              (4 per stage), accumulating into one of two 256-column TMEM stages;
//              tcgen05.commit frees the smem slot / publishes the accumulator.
//   warps 2-5  epilogue: tcgen05.ld 32 lanes x 32 columns — a thread owns ONE query (its TMEM
//              lane) and walks that query's scores for the tile's 256 rows: one compare per
//              element against the query's admission threshold; admitted (dot, row) pairs go to
//              the query's global candidate buffer.  Overlaps the next tile's MMAs.
//
// Admission thresholds.  A first launch of the same kernel in SAMPLE mode scores a strided
// sample of corpus tiles and keeps, per query, the 16 largest dots in registers; a tiny kernel
// turns the 16th largest into a threshold (lowered to the bottom of its float32 score class and
// never below the caller's min_score) expected to admit ~2k..16k rows of the full corpus.  The
// MAIN launch then streams the whole corpus once; a finalize kernel maps the admitted dots to
// scores and selects the top k with the library's total order.  Exactness: every row not
// admitted scores strictly below every admitted row, so if at least k rows were admitted (or the
// threshold is the caller's min_score itself) the result is the exact top-k.  Queries for which
// neither holds (pathological score distributions) or whose buffer overflowed are flagged and
// redone by the exact row-scan path.
//
// Algorithmic bytes per search: N*D*2 (corpus, read once per <=256 queries) + queries + hits.

#include <float.h>
#include <math.h>

#include <algorithm>

#include "tav_common.cuh"
#include "tav_internal.h"
#include "tav_ptx.cuh"

namespace tav {

namespace {

constexpr int kBM = 128;   // queries per accumulator  (TMEM lanes)
constexpr int kBN = 256;   // corpus rows per tile     (TMEM columns per accumulator)
constexpr int kBK = 64;    // 16-bit elements per K slice = one 128-byte swizzle row
constexpr int kUmmaK = 16;
constexpr int kStages = 4;
constexpr int kABytes = kBM * kBK * 2;              // 16 KB
constexpr int kBBytes = kBN * kBK * 2;              // 32 KB
constexpr int kStageBytes = kABytes + kBBytes;      // 48 KB
constexpr int kMmaThreads = 192;
constexpr int kMaxMT = 2;                           // query blocks per launch (2 x 128 queries)
constexpr int kChunkQueries = kBM * kMaxMT;         // 256
constexpr int kSampleTop = 16;
constexpr int kTmemCols = 512;
constexpr size_t kSmemBytes = 1024 + static_cast<size_t>(kStages) * kStageBytes + 256;

enum Mode { kSample = 0, kMain = 1, kDump = 2 };

struct KernelArgs {
    int64_t n_rows;
    int n_work;            // tiles visited by this launch
    int64_t tile_mul;      // visited tile w -> corpus tile (w * tile_mul) / tile_div
    int64_t tile_div;
    int kb_count;          // ceil(dim / 64)
    int nq;                // valid queries in this chunk (<= 256)
    int mt;                // query blocks (1 or 2)
    const float* thr;      // MAIN: [256] admission threshold (raw dot) per query
    float* sample_top;     // SAMPLE: [gridDim.x, 256, 16]
    uint64_t* cand;        // MAIN: [256, capg]  (dot bits << 32 | row)
    uint32_t* cand_count;  // MAIN: [256]
    uint32_t capg;
    float* dump;           // DUMP: [nq, n_rows] raw dots
};

__device__ __forceinline__ void insert_top16(float (&top)[kSampleTop], float x) {
    // top[] sorted descending; x > top[15] on entry
#pragma unroll
    for (int i = 0; i < kSampleTop; ++i) {
        const float hi = fmaxf(top[i], x);
        x = fminf(top[i], x);
        top[i] = hi;
    }
}

template <int MODE>
__global__ void __launch_bounds__(kMmaThreads, 1)
mma_topk_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_c,
                const KernelArgs a, const uint32_t idesc) {
    extern __shared__ uint8_t smem_dyn[];
    // SWIZZLE_128B tiles need 1024-byte alignment
    uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + static_cast<size_t>(kStages) * kStageBytes);
    uint64_t* full = bars;                 // [kStages] TMA -> MMA
    uint64_t* empty = bars + kStages;      // [kStages] MMA -> TMA
    uint64_t* tfull = bars + 2 * kStages;  // [2] MMA -> epilogue
    uint64_t* tempty = tfull + 2;          // [2] epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            ptx::mbar_init(&full[s], 1);
            ptx::mbar_init(&empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(&tfull[s], 1);
            ptx::mbar_init(&tempty[s], 4);  // one arrive per epilogue warp
        }
        ptx::fence_mbar_init();
        ptx::prefetch_tensormap(&map_q);
        ptx::prefetch_tensormap(&map_c);
    }
    if (warp == 1) ptx::tmem_alloc(tmem_slot, kTmemCols);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= TMA producer =================
        if (ptx::elect_one()) {
            uint32_t stage = 0, phase = 0;
            for (int w = blockIdx.x; w < a.n_work; w += gridDim.x) {
                const int64_t tile = (static_cast<int64_t>(w) * a.tile_mul) / a.tile_div;
                const int32_t row0 = static_cast<int32_t>(tile * kBN);
                for (int m = 0; m < a.mt; ++m) {
                    for (int kb = 0; kb < a.kb_count; ++kb) {
                        ptx::mbar_wait(&empty[stage], phase ^ 1);
                        uint8_t* sa = tiles + static_cast<size_t>(stage) * kStageBytes;
                        ptx::mbar_expect_tx(&full[stage], kStageBytes);
                        ptx::tma_load_2d(sa, &map_q, &full[stage], kb * kBK, m * kBM, ptx::kEvictLast);
                        ptx::tma_load_2d(sa + kABytes, &map_c, &full[stage], kb * kBK, row0, ptx::kEvictFirst);
                        if (++stage == kStages) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (ptx::elect_one()) {
            uint32_t stage = 0, phase = 0, item = 0;
            for (int w = blockIdx.x; w < a.n_work; w += gridDim.x) {
                for (int m = 0; m < a.mt; ++m, ++item) {
                    const uint32_t as = item & 1, aphase = (item >> 1) & 1;
                    ptx::mbar_wait(&tempty[as], aphase ^ 1);  // epilogue drained this accumulator
                    ptx::tc_fence_after();
                    const uint32_t d_tmem = tmem_base + as * kBN;
                    for (int kb = 0; kb < a.kb_count; ++kb) {
                        ptx::mbar_wait(&full[stage], phase);
                        ptx::tc_fence_after();
                        const uint32_t sa = ptx::smem_u32(tiles + static_cast<size_t>(stage) * kStageBytes);
                        const uint64_t da = ptx::make_kmajor_sw128_desc(sa);
                        const uint64_t db = ptx::make_kmajor_sw128_desc(sa + kABytes);
#pragma unroll
                        for (int k = 0; k < kBK / kUmmaK; ++k) {
                            // advance 16 elements = 32 bytes along K inside the swizzle atom
                            const uint64_t koff = static_cast<uint64_t>((k * kUmmaK * 2) >> 4);
                            ptx::umma_f16(d_tmem, da + koff, db + koff, idesc, (kb | k) != 0 ? 1u : 0u);
                        }
                        ptx::umma_commit(&empty[stage]);  // smem slot reusable once these MMAs retire
                        if (++stage == kStages) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                    ptx::umma_commit(&tfull[as]);  // accumulator complete
                }
            }
        }
    } else {
        // ================= epilogue: one thread = one query (TMEM lane) =================
        const int quad = warp & 3;  // TMEM lane quadrant this warp may read
        const int lane_q = quad * 32 + lane;
        float tau[kMaxMT];
        float top[kMaxMT][kSampleTop];
#pragma unroll
        for (int m = 0; m < kMaxMT; ++m) {
            const int q = m * kBM + lane_q;
            tau[m] = INFINITY;
            if (MODE == kMain && m < a.mt && q < a.nq) tau[m] = a.thr[q];
#pragma unroll
            for (int i = 0; i < kSampleTop; ++i) top[m][i] = -INFINITY;
        }
        uint32_t item = 0;
        for (int w = blockIdx.x; w < a.n_work; w += gridDim.x) {
            const int64_t tile = (static_cast<int64_t>(w) * a.tile_mul) / a.tile_div;
            const int64_t row0 = tile * kBN;
            const int ncols = static_cast<int>(min(static_cast<int64_t>(kBN), a.n_rows - row0));
#pragma unroll
            for (int m = 0; m < kMaxMT; ++m) {
                if (m >= a.mt) break;
                const uint32_t as = item & 1, aphase = (item >> 1) & 1;
                ++item;
                const int q = m * kBM + lane_q;
                const bool q_valid = q < a.nq;
                ptx::mbar_wait(&tfull[as], aphase);
                ptx::tc_fence_after();
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * kBN;
                for (int c0 = 0; c0 < ncols; c0 += 32) {
                    uint32_t v[32];
                    ptx::tmem_ld_32x32(taddr + c0, v);
                    ptx::tmem_ld_wait();
                    const int nvalid = min(32, ncols - c0);
                    if (MODE == kMain) {
                        const float t = tau[m];
                        const uint32_t rbase = static_cast<uint32_t>(row0 + c0);
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const float x = __uint_as_float(v[i]);
                            if (x >= t && i < nvalid) {
                                const uint32_t slot = atomicAdd(&a.cand_count[q], 1u);
                                if (slot < a.capg)
                                    a.cand[static_cast<size_t>(q) * a.capg + slot] =
                                        (static_cast<uint64_t>(v[i]) << 32) | (rbase + i);
                            }
                        }
                    } else if (MODE == kSample) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const float x = __uint_as_float(v[i]);
                            if (x > top[m][kSampleTop - 1] && i < nvalid) insert_top16(top[m], x);
                        }
                    } else {
                        if (q_valid) {
#pragma unroll
                            for (int i = 0; i < 32; ++i)
                                if (i < nvalid)
                                    a.dump[static_cast<size_t>(q) * a.n_rows + row0 + c0 + i] = __uint_as_float(v[i]);
                        }
                    }
                }
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&tempty[as]);
            }
        }
        if (MODE == kSample) {
#pragma unroll
            for (int m = 0; m < kMaxMT; ++m) {
                if (m >= a.mt) break;
                const int q = m * kBM + lane_q;
                float* dst = a.sample_top + (static_cast<size_t>(blockIdx.x) * kChunkQueries + q) * kSampleTop;
#pragma unroll
                for (int i = 0; i < kSampleTop; ++i) dst[i] = top[m][i];
            }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        ptx::tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ---- float <-> order-preserving uint32 ------------------------------------------------------
__device__ __forceinline__ uint32_t float_to_ord(float f) {
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ord_to_float(uint32_t o) {
    return __uint_as_float((o & 0x80000000u) ? (o ^ 0x80000000u) : ~o);
}
// smallest dot x whose score clip((x+1)/2,0,1) is >= s: -inf when every x qualifies, +inf when none
__device__ float dot_floor_for_score(float s) {
    if (!(s > 0.0f)) return -INFINITY;
    if (s > 1.0f) return INFINITY;
    uint32_t lo = float_to_ord(-FLT_MAX), hi = float_to_ord(FLT_MAX);
    while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if (score_from_dot(ord_to_float(mid)) >= s) hi = mid;
        else lo = mid + 1;
    }
    return ord_to_float(lo);
}

__device__ __forceinline__ void store_rn(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
__device__ __forceinline__ void store_rn(__half* p, float v) { *p = __float2half_rn(v); }

// queries float32 [nq, dim] -> storage dtype [nq_pad, dim], rows >= nq zeroed
template <typename T>
__global__ void query_prep_kernel(const float* q, T* out, int nq, int nq_pad, int dim) {
    const int64_t total = static_cast<int64_t>(nq_pad) * dim;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int row = static_cast<int>(i / dim);
        const float v = row < nq ? q[i] : 0.0f;
        store_rn(out + i, v);
    }
}

// one CTA per query: 16th largest sampled dot -> admission threshold; resets the candidate counters
__global__ void __launch_bounds__(256)
threshold_kernel(const float* sample_top, int sample_ctas, int nq, float floor_score, int use_sample,
                 float* thr, float* floor_out, uint32_t* cand_count, int32_t* retry) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw);
    const int q = blockIdx.x, tid = threadIdx.x;
    const int n = sample_ctas * kSampleTop;
    int cap = 1;
    while (cap < n) cap <<= 1;
    if (cap < 2) cap = 2;
    float result = -INFINITY;
    if (use_sample) {
        for (int i = tid; i < cap; i += 256) {
            uint64_t key = 0;
            if (i < n) {
                const int cta = i / kSampleTop, j = i % kSampleTop;
                const float x = sample_top[(static_cast<size_t>(cta) * kChunkQueries + q) * kSampleTop + j];
                key = float_to_ord(x);
            }
            keys[i] = key;
        }
        bitonic_sort_desc<256>(keys, cap);
        if (n >= kSampleTop) result = ord_to_float(static_cast<uint32_t>(keys[kSampleTop - 1]));
    }
    if (tid == 0) {
        const float floor_x = dot_floor_for_score(floor_score);
        float t = floor_x;
        if (use_sample && result > -INFINITY && !(result != result)) {
            // bottom of the float32 score class of the sampled dot: rows below it score strictly less
            const float snapped = dot_floor_for_score(score_from_dot(result));
            t = fmaxf(snapped, floor_x);
        }
        thr[q] = t;
        floor_out[q] = floor_x;
        cand_count[q] = 0;
        retry[q] = 0;
    }
}

// one CTA per query: admitted (dot,row) pairs -> scores -> top-k, or flag the query for the row scan
__global__ void __launch_bounds__(kSelectThreads)
finalize_kernel(const uint64_t* cand, const uint32_t* cand_count, uint32_t capg, const float* thr,
                const float* floor_x, int k, int64_t item_offset, int64_t* out_items, float* out_scores,
                int32_t* out_counts, int32_t* retry) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw);
    __shared__ int s_cnt;
    __shared__ uint64_t s_admit;
    const int q = blockIdx.x, tid = threadIdx.x;
    const int cap = 1 << (32 - __clz(k + kSelectThreads - 1));
    const uint32_t total = cand_count[q];
    const bool overflow = total > capg;
    const bool starved = total < static_cast<uint32_t>(k) && thr[q] > floor_x[q];
    int64_t* items = out_items + static_cast<size_t>(q) * k;
    float* scores = out_scores + static_cast<size_t>(q) * k;
    if (overflow || starved) {  // CTA-uniform
        for (int j = tid; j < k; j += kSelectThreads) {
            items[j] = -1;
            scores[j] = 0.0f;
        }
        if (tid == 0) {
            out_counts[q] = 0;
            retry[q] = 1;
        }
        return;
    }
    if (tid == 0) {
        s_cnt = 0;
        s_admit = 0;
    }
    CandList l{keys, &s_cnt, &s_admit};
    const uint64_t* in = cand + static_cast<size_t>(q) * capg;
    int need = 0;
    for (uint32_t base = 0; base < total; base += kSelectThreads) {
        if (__syncthreads_or(need)) {
            need = 0;
            list_compact<kSelectThreads>(l, cap, k, 0);
        }
        const uint32_t i = base + tid;
        if (i < total) {
            const uint64_t e = in[i];
            const float x = __uint_as_float(static_cast<uint32_t>(e >> 32));
            const uint64_t key = make_key(score_from_dot(x), static_cast<uint32_t>(e));
            if (key >= s_admit) need |= list_push(l, key, cap - kSelectThreads);
        }
    }
    __syncthreads();
    list_compact<kSelectThreads>(l, cap, k, 0);
    const int n = s_cnt;
    for (int j = tid; j < k; j += kSelectThreads) {
        if (j < n) {
            items[j] = static_cast<int64_t>(key_pos(keys[j])) + item_offset;
            scores[j] = key_score(keys[j]);
        } else {
            items[j] = -1;
            scores[j] = 0.0f;
        }
    }
    if (tid == 0) out_counts[q] = n;
}

// ---- host side --------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
    return fn;
}

// row-major [rows, dim] 16-bit matrix; box = 64 elements x box_rows, 128-byte swizzle, zero OOB fill
bool encode_map(CUtensorMap* map, int dtype, const void* base, int64_t rows, int dim, int box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return false;
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(dim), static_cast<cuuint64_t>(rows)};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(dim) * 2};
    const cuuint32_t box[2] = {kBK, static_cast<cuuint32_t>(box_rows)};
    const cuuint32_t estr[2] = {1, 1};
    const CUtensorMapDataType dt =
        dtype == TAV_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    return fn(map, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

struct Plan {
    int sms;
    int n_tiles;
    int kb_count;
    int n_sample;      // tiles in the sample pass (0 = no sampling)
    int sample_ctas;
    int main_ctas;
    uint32_t capg;
    int nq_pad;        // all queries, padded to 128
    // workspace offsets
    size_t off_q, off_sample, off_thr, off_floor, off_count, off_cand, total;
};

Plan make_plan(int device, int64_t n_rows, int dim, int nq, int k) {
    Plan p{};
    p.sms = 148;
    cudaDeviceGetAttribute(&p.sms, cudaDevAttrMultiProcessorCount, device);
    p.n_tiles = static_cast<int>((n_rows + kBN - 1) / kBN);
    p.kb_count = (dim + kBK - 1) / kBK;
    const int64_t target = std::max<int64_t>(2048, 16ll * k);  // rows we aim to admit per query
    if (n_rows <= 16384) {
        p.n_sample = 0;
        p.capg = static_cast<uint32_t>(n_rows);
    } else {
        const int64_t sample_rows = (static_cast<int64_t>(kSampleTop) * n_rows + target - 1) / target;
        p.n_sample = static_cast<int>(std::min<int64_t>(p.n_tiles, std::max<int64_t>(1, (sample_rows + kBN - 1) / kBN)));
        p.capg = static_cast<uint32_t>(8 * target);
    }
    p.sample_ctas = std::max(1, std::min(p.n_sample, p.sms));
    p.main_ctas = std::min(p.n_tiles, p.sms);
    p.nq_pad = ((nq + kBM - 1) / kBM) * kBM;
    auto align = [](size_t v) { return (v + 255) & ~size_t(255); };
    size_t off = 0;
    p.off_q = off;
    off = align(off + static_cast<size_t>(p.nq_pad) * dim * 2);
    p.off_sample = off;
    off = align(off + static_cast<size_t>(p.sample_ctas) * kChunkQueries * kSampleTop * sizeof(float));
    p.off_thr = off;
    off = align(off + kChunkQueries * sizeof(float));
    p.off_floor = off;
    off = align(off + kChunkQueries * sizeof(float));
    p.off_count = off;
    off = align(off + kChunkQueries * sizeof(uint32_t));
    p.off_cand = off;
    off = align(off + static_cast<size_t>(kChunkQueries) * p.capg * sizeof(uint64_t));
    p.total = off;
    return p;
}

template <int MODE>
cudaError_t launch_kernel(const CUtensorMap& mq, const CUtensorMap& mc, const KernelArgs& ka, uint32_t idesc,
                          int grid, cudaStream_t s) {
    auto kern = mma_topk_kernel<MODE>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(kSmemBytes));
    if (e != cudaSuccess) return e;
    kern<<<grid, kMmaThreads, kSmemBytes, s>>>(mq, mc, ka, idesc);
    return cudaGetLastError();
}

cudaError_t prep_queries(const MmaArgs& a, void* dst, int nq_pad, cudaStream_t s) {
    const int64_t total = static_cast<int64_t>(nq_pad) * a.dim;
    const int grid = static_cast<int>(std::min<int64_t>((total + 255) / 256, 148 * 8));
    if (a.dtype == TAV_BF16)
        query_prep_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(a.queries, static_cast<__nv_bfloat16*>(dst), a.nq,
                                                              nq_pad, a.dim);
    else
        query_prep_kernel<__half><<<grid, 256, 0, s>>>(a.queries, static_cast<__half*>(dst), a.nq, nq_pad, a.dim);
    return cudaGetLastError();
}

}  // namespace

bool mma_supported(int dtype, int dim) {
    return (dtype == TAV_BF16 || dtype == TAV_F16) && dim >= 8 && dim % 8 == 0;
}

size_t mma_workspace_bytes(const MmaArgs& a) { return make_plan(a.device, a.n_corpus, a.dim, a.nq, a.k).total; }

cudaError_t launch_mma_search(const MmaArgs& a, void* workspace, size_t workspace_bytes, cudaStream_t s,
                              int* launches) {
    if (!mma_supported(a.dtype, a.dim) || a.k > kPassK || a.n_corpus >= (1ll << 31) ||
        reinterpret_cast<uintptr_t>(a.corpus) % 16 != 0)
        return cudaErrorInvalidValue;
    const Plan p = make_plan(a.device, a.n_corpus, a.dim, a.nq, a.k);
    if (workspace_bytes < p.total) return cudaErrorInvalidValue;
    char* ws = static_cast<char*>(workspace);
    void* d_q = ws + p.off_q;
    float* d_sample = reinterpret_cast<float*>(ws + p.off_sample);
    float* d_thr = reinterpret_cast<float*>(ws + p.off_thr);
    float* d_floor = reinterpret_cast<float*>(ws + p.off_floor);
    uint32_t* d_count = reinterpret_cast<uint32_t*>(ws + p.off_count);
    uint64_t* d_cand = reinterpret_cast<uint64_t*>(ws + p.off_cand);
    int n_launch = 0;

    cudaError_t e = prep_queries(a, d_q, p.nq_pad, s);
    if (e != cudaSuccess) return e;
    ++n_launch;

    CUtensorMap map_c;
    if (!encode_map(&map_c, a.dtype, a.corpus, a.n_corpus, a.dim, kBN)) return cudaErrorUnknown;
    const uint32_t idesc = ptx::make_idesc_f16(kBM, kBN, a.dtype == TAV_BF16 ? 1 : 0);

    int ev_used = 0;
    for (int q0 = 0; q0 < a.nq; q0 += kChunkQueries) {
        const int nq = std::min(kChunkQueries, a.nq - q0);
        const int mt = (nq + kBM - 1) / kBM;
        CUtensorMap map_q;
        const char* qbase = static_cast<const char*>(d_q) + static_cast<size_t>(q0) * a.dim * 2;
        if (!encode_map(&map_q, a.dtype, qbase, mt * kBM, a.dim, kBM)) return cudaErrorUnknown;

        KernelArgs ka{};
        ka.n_rows = a.n_corpus;
        ka.kb_count = p.kb_count;
        ka.nq = nq;
        ka.mt = mt;
        ka.thr = d_thr;
        ka.sample_top = d_sample;
        ka.cand = d_cand;
        ka.cand_count = d_count;
        ka.capg = p.capg;

        const bool timed = a.ev && ev_used < a.ev_max;
        if (timed) {
            e = cudaEventRecord(a.ev[ev_used][0], s);
            if (e != cudaSuccess) return e;
        }
        if (p.n_sample > 0) {
            ka.n_work = p.n_sample;
            ka.tile_mul = p.n_tiles;
            ka.tile_div = p.n_sample;
            e = launch_kernel<kSample>(map_q, map_c, ka, idesc, p.sample_ctas, s);
            if (e != cudaSuccess) return e;
            ++n_launch;
        }
        int cap = 2;
        while (cap < p.sample_ctas * kSampleTop) cap <<= 1;
        threshold_kernel<<<nq, 256, static_cast<size_t>(cap) * sizeof(uint64_t), s>>>(
            d_sample, p.sample_ctas, nq, a.floor_score, p.n_sample > 0 ? 1 : 0, d_thr, d_floor, d_count,
            a.retry_flags + q0);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        ++n_launch;

        ka.n_work = p.n_tiles;
        ka.tile_mul = 1;
        ka.tile_div = 1;
        e = launch_kernel<kMain>(map_q, map_c, ka, idesc, p.main_ctas, s);
        if (e != cudaSuccess) return e;
        ++n_launch;
        if (timed) {
            e = cudaEventRecord(a.ev[ev_used][1], s);
            if (e != cudaSuccess) return e;
            ++ev_used;
        }

        const size_t sel_smem = static_cast<size_t>(next_pow2(a.k + kSelectThreads)) * sizeof(uint64_t);
        e = cudaFuncSetAttribute(finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 static_cast<int>(sel_smem));
        if (e != cudaSuccess) return e;
        finalize_kernel<<<nq, kSelectThreads, sel_smem, s>>>(
            d_cand, d_count, p.capg, d_thr, d_floor, a.k, a.item_offset, a.out_items + static_cast<size_t>(q0) * a.k,
            a.out_scores + static_cast<size_t>(q0) * a.k, a.out_counts + q0, a.retry_flags + q0);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        ++n_launch;
    }
    if (launches) *launches = n_launch;
    if (a.ev_used) *a.ev_used = ev_used;
    return cudaSuccess;
}

// Debug / verification entry: all raw dot products of the tensor-core path, out[nq, n_rows] (device).
cudaError_t launch_mma_dump(const MmaArgs& a, void* workspace, size_t workspace_bytes, float* out, cudaStream_t s) {
    if (!mma_supported(a.dtype, a.dim) || a.n_corpus >= (1ll << 31)) return cudaErrorInvalidValue;
    const Plan p = make_plan(a.device, a.n_corpus, a.dim, a.nq, 1);
    if (workspace_bytes < p.total) return cudaErrorInvalidValue;
    char* ws = static_cast<char*>(workspace);
    void* d_q = ws + p.off_q;
    cudaError_t e = prep_queries(a, d_q, p.nq_pad, s);
    if (e != cudaSuccess) return e;
    CUtensorMap map_c;
    if (!encode_map(&map_c, a.dtype, a.corpus, a.n_corpus, a.dim, kBN)) return cudaErrorUnknown;
    const uint32_t idesc = ptx::make_idesc_f16(kBM, kBN, a.dtype == TAV_BF16 ? 1 : 0);
    for (int q0 = 0; q0 < a.nq; q0 += kChunkQueries) {
        const int nq = std::min(kChunkQueries, a.nq - q0);
        const int mt = (nq + kBM - 1) / kBM;
        CUtensorMap map_q;
        const char* qbase = static_cast<const char*>(d_q) + static_cast<size_t>(q0) * a.dim * 2;
        if (!encode_map(&map_q, a.dtype, qbase, mt * kBM, a.dim, kBM)) return cudaErrorUnknown;
        KernelArgs ka{};
        ka.n_rows = a.n_corpus;
        ka.kb_count = p.kb_count;
        ka.nq = nq;
        ka.mt = mt;
        ka.n_work = p.n_tiles;
        ka.tile_mul = 1;
        ka.tile_div = 1;
        ka.dump = out + static_cast<size_t>(q0) * a.n_corpus;
        e = launch_kernel<kDump>(map_q, map_c, ka, idesc, p.main_ctas, s);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace tav
