// WNN (weighted nearest neighbours, Hao et al.) building blocks -- reference muon/_core/preproc.py:264-640.
//
//  wnn_bandwidth : per cell, the n_bw (20) cells of lowest shared-neighbour overlap among those that share at
//                  least one kNN neighbour with it (largest Jaccard distance of the neighbour sets, ties towards
//                  the larger Euclidean distance), and the mean Euclidean distance to them = kernel bandwidth
//                  sigma_i.  The reference gets this from NN-descent with a custom metric
//                  (preproc.py:51-76, 408-470); here the candidate set is enumerated exactly through the
//                  transposed kNN graph (cells j with S_i n S_j != {}  <=>  j in U_{t in S_i} revnbrs(t)).
//  wnn_affinity_topk : per cell, union of the per-modality candidate lists, weighted affinity
//                  sum_m w_im exp(-||x^m_i - x^m_j|| / sigma^m_i), distance sqrt(0.5 (1 - affinity)) and the
//                  n_out smallest (preproc.py:569-604, incl. the numba top-k _sparse_csr_fast_knn_ :114-135).
//
// One warp per cell; a per-warp open-addressing hash table in shared memory holds the candidate set (keys =
// cell index, counts = |S_i n S_j|).  Arithmetic that decides orderings runs in fp64 like the reference (numpy
// float64); embeddings are fp32.
#include <float.h>
#include <math.h>

#include "common.cuh"

namespace mub {

constexpr int kWnnWarps = 4;
constexpr int kWnnH = 2048;            // hash slots per warp
constexpr int kWnnMaxCand = 1536;      // distinct candidates per cell before we report overflow
constexpr int kWnnMaxMod = 4;

struct WarpTable {
    int32_t* key;    // [H]  -1 = empty
    int32_t* cnt;    // [H]
    float* aux;      // [H]  per-candidate float (Euclidean distance / affinity distance)
    int32_t* aux2;   // [H]  per-candidate int (|S_j|)
    int32_t* list;   // [max_cand] occupied slots in insertion order
    int H;           // slots (power of two): kWnnH in shared memory, larger in the global-memory fallback
    int shift;       // 32 - log2(H)
    int max_cand;
};

__device__ __forceinline__ uint32_t wnn_hash(const WarpTable& t, int32_t j) {
    return ((uint32_t)j * 2654435761u) >> t.shift;
}

// insert j (count += 1); returns false on overflow.  Called by any subset of lanes.
__device__ __forceinline__ bool table_insert(const WarpTable& t, int32_t j, int* n_list) {
    uint32_t h = wnn_hash(t, j);
    for (int probe = 0; probe < t.H; ++probe) {
        const int32_t old = atomicCAS(&t.key[h], -1, j);
        if (old == -1) {  // new candidate
            const int pos = atomicAdd(n_list, 1);
            if (pos >= t.max_cand) return false;
            t.list[pos] = (int32_t)h;
            atomicAdd(&t.cnt[h], 1);
            return true;
        }
        if (old == j) {
            atomicAdd(&t.cnt[h], 1);
            return true;
        }
        h = (h + 1) & (t.H - 1);
    }
    return false;
}

__device__ __forceinline__ int table_find(const WarpTable& t, int32_t j) {
    uint32_t h = wnn_hash(t, j);
    for (int probe = 0; probe < t.H; ++probe) {
        const int32_t k = t.key[h];
        if (k == j) return (int)h;
        if (k == -1) return -1;
        h = (h + 1) & (t.H - 1);
    }
    return -1;
}

__device__ __forceinline__ float row_dist(const float* __restrict__ a, const float* __restrict__ b, int d) {
    double s = 0.0;
    for (int t = 0; t < d; ++t) {
        const double df = (double)a[t] - (double)b[t];
        s += df * df;
    }
    return (float)sqrt(s);
}

__host__ __device__ inline size_t table_bytes(int H, int max_cand) {
    return sizeof(int32_t) * (size_t)H * 3 + sizeof(float) * (size_t)H + sizeof(int32_t) * (size_t)max_cand + 16;
}
__device__ __forceinline__ WarpTable carve(unsigned char* base, int slot, int H, int max_cand) {
    unsigned char* p = base + (size_t)slot * table_bytes(H, max_cand);
    WarpTable t;
    t.key = reinterpret_cast<int32_t*>(p);
    t.cnt = t.key + H;
    t.aux2 = t.cnt + H;
    t.aux = reinterpret_cast<float*>(t.aux2 + H);
    t.list = reinterpret_cast<int32_t*>(t.aux + H);
    t.H = H;
    t.shift = 32 - (31 - __clz(H));
    t.max_cand = max_cand;
    return t;
}
constexpr size_t kWnnSmemPerWarp = sizeof(int32_t) * kWnnH * 3 + sizeof(float) * kWnnH + sizeof(int32_t) * kWnnMaxCand + 16;

// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kWnnWarps * 32)
wnn_bandwidth_kernel(const int64_t* __restrict__ g_indptr, const int32_t* __restrict__ g_indices,
                     const int64_t* __restrict__ t_indptr, const int32_t* __restrict__ t_indices,
                     const float* __restrict__ X, int64_t n, int d, int ld, int n_bw, double bbox,
                     double* __restrict__ sigma, int32_t* __restrict__ status,
                     const int64_t* __restrict__ cell_list, int64_t n_list_cells, unsigned char* gtable, int gH,
                     int gmax) {
    // Fast path: cell_list == NULL, tables in shared memory (kWnnH slots).  Fallback for hub cells: cell_list
    // holds the cells whose candidate set overflowed, tables live in global memory (gH slots per warp).
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int n_list_s[kWnnWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool big = (cell_list != nullptr);
    const WarpTable T = big ? carve(gtable, blockIdx.x * kWnnWarps + warp, gH, gmax)
                            : carve(smem_raw, warp, kWnnH, kWnnMaxCand);
    int* n_list = &n_list_s[warp];
    const double Nd = (double)n;
    for (int h = lane; h < T.H; h += 32) { T.key[h] = -1; T.cnt[h] = 0; }
    if (lane == 0) *n_list = 0;
    __syncwarp();

    const int64_t n_work = big ? n_list_cells : n;
    for (int64_t w = (int64_t)blockIdx.x * kWnnWarps + warp; w < n_work; w += (int64_t)gridDim.x * kWnnWarps) {
        const int64_t i = big ? cell_list[w] : w;
        const int64_t s0 = g_indptr[i], s1 = g_indptr[i + 1];
        const int si = (int)(s1 - s0);
        bool ok = true;
        // 1. enumerate cells sharing a neighbour with i through the transposed graph
        for (int64_t a = s0; a < s1; ++a) {
            const int32_t t = g_indices[a];
            const int64_t r0 = t_indptr[t], r1 = t_indptr[t + 1];
            for (int64_t b = r0 + lane; b < r1; b += 32) {
                const int32_t j = t_indices[b];
                if (j != (int32_t)i) ok &= table_insert(T, j, n_list);
            }
        }
        __syncwarp();
        const bool overflow = !__all_sync(0xffffffffu, ok) || *n_list > T.max_cand;
        if (overflow && lane == 0) atomicOr(status, big ? 2 : 1);   // bit0: needs the fallback, bit1: fallback too small
        const int nc = min(*n_list, T.max_cand);
        // 2. per candidate: |S_j| and the Euclidean distance to i
        const float* xi = X + (size_t)i * ld;
        for (int c = lane; c < nc; c += 32) {
            const int h = T.list[c];
            const int32_t j = T.key[h];
            T.aux2[h] = (int32_t)(g_indptr[j + 1] - g_indptr[j]);
            T.aux[h] = row_dist(xi, X + (size_t)j * ld, d);
        }
        __syncwarp();
        // 3. n_bw rounds of min-extraction on value = (N - jd N) + (bbox - e)/bbox, ties -> lower index
        double esum = 0.0;
        int taken = 0;
        for (; taken < n_bw && taken < nc; ++taken) {
            double bv = DBL_MAX;
            int bj = 0x7fffffff, bh = -1;
            for (int c = lane; c < nc; c += 32) {
                const int h = T.list[c];
                const int cn = T.cnt[h];
                if (cn <= 0) continue;               // already taken
                const double u = (double)(si + T.aux2[h] - cn);
                const double jd = (u - (double)cn) / u;
                const double v = (Nd - jd * Nd) + (bbox - (double)T.aux[h]) / bbox;
                const int32_t j = T.key[h];
                if (v < bv || (v == bv && j < bj)) { bv = v; bj = j; bh = h; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
                const int oh = __shfl_xor_sync(0xffffffffu, bh, o);
                if (ov < bv || (ov == bv && oj < bj)) { bv = ov; bj = oj; bh = oh; }
            }
            esum += (double)T.aux[bh];
            __syncwarp();
            if (lane == 0) T.cnt[bh] = -T.cnt[bh];   // mark as taken (sign flip keeps the key for lookups)
            __syncwarp();
        }
        // 4. fewer overlapping cells than n_bw: every other cell (i itself included) ties at N + 1 and a
        //    stable sort hands out the lowest indices first (what an exact search returns for this metric)
        for (int64_t j = 0; taken < n_bw && j < n; ++j) {
            if (table_find(T, (int32_t)j) >= 0) continue;
            esum += (double)row_dist(xi, X + (size_t)j * ld, d);
            ++taken;
        }
        if (lane == 0) sigma[i] = overflow ? -1.0 : esum / (double)n_bw;   // -1: recomputed by the fallback pass
        // 5. reset the touched slots
        __syncwarp();
        for (int c = lane; c < nc; c += 32) {
            const int h = T.list[c];
            T.key[h] = -1;
            T.cnt[h] = 0;
        }
        if (overflow) {                              // clear everything
            for (int h = lane; h < T.H; h += 32) { T.key[h] = -1; T.cnt[h] = 0; }
        }
        __syncwarp();
        if (lane == 0) *n_list = 0;
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------------
struct WnnMods {
    int n_mod;
    const float* rep[kWnnMaxMod];
    int dim[kWnnMaxMod];
    int ld[kWnnMaxMod];
    const int32_t* cand[kWnnMaxMod];    // [n x n_cand] candidate indices, -1 = none
    const double* sigma[kWnnMaxMod];    // [n]
};

__global__ void __launch_bounds__(kWnnWarps * 32)
wnn_affinity_topk_kernel(WnnMods M, const double* __restrict__ weight, int64_t n, int n_cand, int n_out,
                         int32_t* __restrict__ out_idx, double* __restrict__ out_dist, int32_t* __restrict__ status) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int n_list_s[kWnnWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const WarpTable T = carve(smem_raw, warp, kWnnH, kWnnMaxCand);
    int* n_list = &n_list_s[warp];
    for (int h = lane; h < kWnnH; h += 32) { T.key[h] = -1; T.cnt[h] = 0; }
    if (lane == 0) *n_list = 0;
    __syncwarp();

    for (int64_t i = (int64_t)blockIdx.x * kWnnWarps + warp; i < n; i += (int64_t)gridDim.x * kWnnWarps) {
        bool ok = true;
        for (int m = 0; m < M.n_mod; ++m) {
            const int32_t* row = M.cand[m] + (size_t)i * n_cand;
            for (int c = lane; c < n_cand; c += 32) {
                const int32_t j = row[c];
                if (j >= 0 && j != (int32_t)i) ok &= table_insert(T, j, n_list);
            }
        }
        __syncwarp();
        if (!__all_sync(0xffffffffu, ok) || *n_list > kWnnMaxCand) {
            if (lane == 0) atomicOr(status, 1);
        }
        const int nc = min(*n_list, kWnnMaxCand);
        // affinity of every candidate over all modalities (fp64 like the reference), distance in aux (float bits
        // are not enough for ordering ties only; the exact double is recomputed for the winners)
        for (int c = lane; c < nc; c += 32) {
            const int h = T.list[c];
            const int32_t j = T.key[h];
            double aff = 0.0;
            for (int m = 0; m < M.n_mod; ++m) {
                const float* a = M.rep[m] + (size_t)i * M.ld[m];
                const float* b = M.rep[m] + (size_t)j * M.ld[m];
                double s = 0.0;
                for (int t = 0; t < M.dim[m]; ++t) {
                    const double df = (double)a[t] - (double)b[t];
                    s += df * df;
                }
                aff += exp(-sqrt(s) / M.sigma[m][i]) * weight[(size_t)i * M.n_mod + m];
            }
            // keep the double in two int slots: cnt/aux2 are free here
            const double dist = sqrt(0.5 * (1.0 - aff));
            const long long bits = __double_as_longlong(dist);
            T.cnt[h] = (int32_t)(bits & 0xffffffffll);
            T.aux2[h] = (int32_t)(bits >> 32);
        }
        __syncwarp();
        for (int r = 0; r < n_out; ++r) {
            double bv = DBL_MAX;
            int bj = 0x7fffffff, bh = -1;
            for (int c = lane; c < nc; c += 32) {
                const int h = T.list[c];
                const int32_t j = T.key[h];
                if (j < 0) continue;                 // taken
                const long long bits = ((long long)T.aux2[h] << 32) | (unsigned int)T.cnt[h];
                const double v = __longlong_as_double(bits);
                if (v < bv || (v == bv && j < bj)) { bv = v; bj = j; bh = h; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
                const int oh = __shfl_xor_sync(0xffffffffu, bh, o);
                if (ov < bv || (ov == bv && oj < bj)) { bv = ov; bj = oj; bh = oh; }
            }
            if (lane == 0) {
                out_idx[(size_t)i * n_out + r] = (bh >= 0) ? bj : -1;
                out_dist[(size_t)i * n_out + r] = (bh >= 0) ? bv : INFINITY;
            }
            __syncwarp();
            if (bh >= 0 && lane == 0) T.key[bh] = -2 - bj;   // taken; still a non-empty slot for probing
            __syncwarp();
        }
        // reset
        for (int c = lane; c < nc; c += 32) {
            const int h = T.list[c];
            T.key[h] = -1;
            T.cnt[h] = 0;
        }
        if (*n_list > kWnnMaxCand) {
            for (int h = lane; h < kWnnH; h += 32) { T.key[h] = -1; T.cnt[h] = 0; }
        }
        __syncwarp();
        if (lane == 0) *n_list = 0;
        __syncwarp();
    }
}

static int wnn_grid(int64_t n) {
    int64_t want = (n + kWnnWarps - 1) / kWnnWarps, cap = (int64_t)sm_count() * 1;  // 1 CTA / SM (shared memory)
    int64_t g = want < cap ? want : cap;
    return (int)(g < 1 ? 1 : g);
}

}  // namespace mub

extern "C" {

size_t mub_wnn_bandwidth_workspace_bytes(int32_t table_slots, int32_t n_tables) {
    return mub::table_bytes(table_slots, table_slots / 2) * (size_t)n_tables;
}

int mub_wnn_bandwidth_f32(const int64_t* g_indptr, const int32_t* g_indices, const int64_t* t_indptr,
                          const int32_t* t_indices, const float* X, int64_t n, int32_t d, int32_t ld, int32_t n_bw,
                          double bbox_norm, double* sigma, int32_t* status, const int64_t* cell_list,
                          int64_t n_cells, void* workspace, int32_t table_slots, int32_t n_tables,
                          mub_stream_t stream) {
    MUB_REQUIRE(n >= 0 && d >= 1 && ld >= d && n_bw >= 1, "wnn_bandwidth: bad arguments");
    if (n == 0) return 0;
    MUB_REQUIRE(g_indptr && g_indices && t_indptr && t_indices && X && sigma && status, "wnn_bandwidth: null pointer");
    const size_t smem = mub::kWnnSmemPerWarp * mub::kWnnWarps;
    cudaError_t e = cudaFuncSetAttribute(mub::wnn_bandwidth_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        mub::set_error("wnn_bandwidth: %zu B of shared memory: %s", smem, cudaGetErrorString(e));
        return -2;
    }
    if (cell_list == nullptr) {
        mub::wnn_bandwidth_kernel<<<mub::wnn_grid(n), mub::kWnnWarps * 32, smem, (cudaStream_t)stream>>>(
            g_indptr, g_indices, t_indptr, t_indices, X, n, d, ld, n_bw, bbox_norm, sigma, status, nullptr, 0,
            nullptr, 0, 0);
        return mub::check_launch("wnn_bandwidth");
    }
    // fallback for hub cells: big tables in caller-provided global memory, n_tables warps
    if (n_cells <= 0) return 0;
    MUB_REQUIRE(workspace && table_slots >= 1024 && (table_slots & (table_slots - 1)) == 0 && n_tables >= mub::kWnnWarps,
                "wnn_bandwidth: fallback needs a workspace, a power-of-two table size and >= 4 tables");
    int grid = n_tables / mub::kWnnWarps;
    const int64_t want = (n_cells + mub::kWnnWarps - 1) / mub::kWnnWarps;
    if (grid > want) grid = (int)want;
    mub::wnn_bandwidth_kernel<<<grid, mub::kWnnWarps * 32, smem, (cudaStream_t)stream>>>(
        g_indptr, g_indices, t_indptr, t_indices, X, n, d, ld, n_bw, bbox_norm, sigma, status, cell_list, n_cells,
        (unsigned char*)workspace, table_slots, table_slots / 2);
    return mub::check_launch("wnn_bandwidth_fallback");
}

int mub_wnn_affinity_topk_f32(int32_t n_mod, const float* const* reps, const int32_t* dims, const int32_t* lds,
                              const int32_t* const* cands, const double* const* sigmas, const double* weight,
                              int64_t n, int32_t n_cand, int32_t n_out, int32_t* out_idx, double* out_dist,
                              int32_t* status, mub_stream_t stream) {
    MUB_REQUIRE(n_mod >= 1 && n_mod <= mub::kWnnMaxMod, "wnn_affinity_topk: 1 <= n_mod <= 4");
    MUB_REQUIRE(n >= 0 && n_cand >= 1 && n_out >= 1, "wnn_affinity_topk: bad sizes");
    if (n == 0) return 0;
    MUB_REQUIRE(reps && dims && lds && cands && sigmas && weight && out_idx && out_dist && status,
                "wnn_affinity_topk: null pointer");
    mub::WnnMods M;
    M.n_mod = n_mod;
    for (int m = 0; m < n_mod; ++m) {
        M.rep[m] = reps[m];
        M.dim[m] = dims[m];
        M.ld[m] = lds[m];
        M.cand[m] = cands[m];
        M.sigma[m] = sigmas[m];
    }
    const size_t smem = mub::kWnnSmemPerWarp * mub::kWnnWarps;
    cudaError_t e = cudaFuncSetAttribute(mub::wnn_affinity_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        mub::set_error("wnn_affinity_topk: %zu B of shared memory: %s", smem, cudaGetErrorString(e));
        return -2;
    }
    mub::wnn_affinity_topk_kernel<<<mub::wnn_grid(n), mub::kWnnWarps * 32, smem, (cudaStream_t)stream>>>(
        M, weight, n, n_cand, n_out, out_idx, out_dist, status);
    return mub::check_launch("wnn_affinity_topk");
}

}  // extern "C"
