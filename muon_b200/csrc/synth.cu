// Synthetic ATAC count generator (benchmark / test INPUT only; SURVEY App. E, adapted).
//
// Planted-topic Bernoulli model evaluated for every (cell, peak) pair with a counter-based
// hash, so any row range of any shard is reproducible and bit-identical to the numpy twin in
// muon_b200/_synth.py (only IEEE-exact float ops and integer hashing are used):
//   p_ij   = min(0.9, (0.5*beta_j + topic[t_i][j]) * row_scale_i)
//   keep   = hi32(h_ij) < uint32(p_ij * 2^32),  h_ij = mix64(seedmix + (row0+i)*n_cols + j)
//   count  = 1 + #{thresholds of a capped geometric(0.6) below lo32(h_ij)}  ("mostly 1 and 2")
// beta, topic, row_topic, row_scale are small tables produced on the host.
#include "common.cuh"

namespace mub {

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

__device__ __forceinline__ bool synth_entry(uint64_t key, float beta, float tw, float scale, float& value) {
    const float base = __fadd_rn(__fmul_rn(0.5f, beta), tw);
    const float p = fminf(__fmul_rn(base, scale), 0.9f);
    const uint32_t thr = __float2uint_rz(__fmul_rn(p, 4294967296.0f));
    const uint64_t h = mix64(key);
    const uint32_t hi = (uint32_t)(h >> 32), lo = (uint32_t)h;
    // capped geometric: P(1)=.6 P(2)=.24 P(3)=.096 P(4)=.0384 P(5)=rest
    value = 1.0f + (lo >= 2576980378u) + (lo >= 3607772529u) + (lo >= 4020089389u) + (lo >= 4185016133u);
    return hi < thr;
}

template <bool FILL>
__global__ void __launch_bounds__(256)
synth_kernel(int64_t row0, int64_t n_rows, int32_t n_cols, const float* __restrict__ beta,
             const float* __restrict__ topic, const int32_t* __restrict__ row_topic,
             const float* __restrict__ row_scale, uint64_t seedmix, int64_t* __restrict__ row_nnz,
             const int64_t* __restrict__ indptr, int32_t* __restrict__ indices, float* __restrict__ data) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int64_t n_warps = (int64_t)gridDim.x * 8;
    for (int64_t r = warp; r < n_rows; r += n_warps) {
        const float* trow = topic + (size_t)__ldg(row_topic + r) * n_cols;
        const float scale = __ldg(row_scale + r);
        const uint64_t rowkey = seedmix + (uint64_t)(row0 + r) * (uint64_t)n_cols;
        int64_t out = FILL ? __ldg(indptr + r) : 0;
        for (int32_t j0 = 0; j0 < n_cols; j0 += 32) {
            const int32_t j = j0 + lane;
            float val = 0.f;
            bool keep = false;
            if (j < n_cols) keep = synth_entry(rowkey + (uint64_t)j, __ldg(beta + j), __ldg(trow + j), scale, val);
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            if (FILL && keep) {
                const int64_t pos = out + __popc(m & ((1u << lane) - 1u));
                indices[pos] = j;
                data[pos] = val;
            }
            out += __popc(m);
        }
        if (!FILL && lane == 0) row_nnz[r] = out;
    }
}

static int synth_grid(int64_t n_rows) {
    int64_t want = (n_rows + 7) / 8, cap = (int64_t)sm_count() * 8;
    int64_t g = want < cap ? want : cap;
    return (int)(g < 1 ? 1 : g);
}

}  // namespace mub

extern "C" {

int mub_synth_count(int64_t row0, int64_t n_rows, int32_t n_cols, const float* beta, const float* topic,
                    const int32_t* row_topic, const float* row_scale, uint64_t seed, int64_t* row_nnz,
                    mub_stream_t stream) {
    MUB_REQUIRE(n_rows >= 0 && n_cols > 0, "synth_count: bad shape");
    if (n_rows == 0) return 0;
    mub::synth_kernel<false><<<mub::synth_grid(n_rows), 256, 0, (cudaStream_t)stream>>>(
        row0, n_rows, n_cols, beta, topic, row_topic, row_scale, mub::mix64(seed), row_nnz, nullptr, nullptr,
        nullptr);
    return mub::check_launch("synth_count");
}

int mub_synth_fill(int64_t row0, int64_t n_rows, int32_t n_cols, const float* beta, const float* topic,
                   const int32_t* row_topic, const float* row_scale, uint64_t seed, const int64_t* indptr,
                   int32_t* indices, float* data, mub_stream_t stream) {
    MUB_REQUIRE(n_rows >= 0 && n_cols > 0, "synth_fill: bad shape");
    if (n_rows == 0) return 0;
    mub::synth_kernel<true><<<mub::synth_grid(n_rows), 256, 0, (cudaStream_t)stream>>>(
        row0, n_rows, n_cols, beta, topic, row_topic, row_scale, mub::mix64(seed), nullptr, indptr, indices,
        data);
    return mub::check_launch("synth_fill");
}

}  // extern "C"
