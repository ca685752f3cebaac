// K4 -- tall-skinny Gram  G[l x l] = Y^T diag(w) Y,  Y[n x ld] row-major, ld in {32,64,128}.
//
// This is the only contraction over the (sharded) cell axis whose result is small: the l x l
// matrix that gets allreduced across cell shards.  It replaces the Gram hidden in
// scipy.linalg.svd(A @ eigvec) (scipy _svds.py:511-521, reached from muon/_atac/tools.py:53)
// and mofapy2's E[Z]^T E[Z] / E[W]^T diag(tau) E[W] (reached from muon/_core/tools.py:585).
//
// Each CTA owns a contiguous slab of rows, stages 32-row chunks in shared memory and keeps a
// (ld/16)^2 register tile per thread; slab partials are fp32, the cross-slab reduction runs in
// fp64 in a fixed order, so the result is deterministic and accurate to ~1e-6 relative.
// n*ld*4 bytes are read once: HBM-bound, ~0.04 ms per 256 MB operand at roofline.
#include "common.cuh"

namespace mub {

constexpr int kGramThreads = 256;
constexpr int kGramChunk = 32;

static int gram_grid(int64_t n) {
    int64_t want = (n + kGramChunk * 4 - 1) / (kGramChunk * 4);  // >= 128 rows per slab
    int64_t cap = (int64_t)sm_count() * 4;
    int64_t g = want < cap ? want : cap;
    return (int)(g < 1 ? 1 : g);
}

template <int LD>
__global__ void __launch_bounds__(kGramThreads)
gram_partial_kernel(const float* __restrict__ Y, const float* __restrict__ w, int64_t n, float* __restrict__ part) {
    constexpr int TR = LD / 16;
    __shared__ __align__(16) float tile[kGramChunk][LD];
    __shared__ float wt[kGramChunk];
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int64_t rows_per = (n + gridDim.x - 1) / gridDim.x;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per;
    const int64_t r1 = (r0 + rows_per < n) ? r0 + rows_per : n;
    float acc[TR][TR];
#pragma unroll
    for (int i = 0; i < TR; ++i)
#pragma unroll
        for (int j = 0; j < TR; ++j) acc[i][j] = 0.f;

    for (int64_t base = r0; base < r1; base += kGramChunk) {
        const int rows = (int)((r1 - base < kGramChunk) ? (r1 - base) : kGramChunk);
        // cooperative coalesced float4 load of rows x LD
        constexpr int V = LD / 4;
        for (int idx = tid; idx < kGramChunk * V; idx += kGramThreads) {
            const int r = idx / V, c4 = idx % V;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < rows) v = ld_stream4(reinterpret_cast<const float4*>(Y + (size_t)(base + r) * LD) + c4);
            *reinterpret_cast<float4*>(&tile[r][c4 * 4]) = v;
        }
        if (tid < kGramChunk) wt[tid] = (tid < rows) ? (w ? __ldg(w + base + tid) : 1.f) : 0.f;
        __syncthreads();
#pragma unroll 4
        for (int r = 0; r < kGramChunk; ++r) {
            float a[TR], b[TR];
            const float wr = wt[r];
#pragma unroll
            for (int i = 0; i < TR; ++i) a[i] = tile[r][ty * TR + i] * wr;
#pragma unroll
            for (int j = 0; j < TR; ++j) b[j] = tile[r][tx * TR + j];
#pragma unroll
            for (int i = 0; i < TR; ++i)
#pragma unroll
                for (int j = 0; j < TR; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    float* out = part + (size_t)blockIdx.x * LD * LD;
#pragma unroll
    for (int i = 0; i < TR; ++i)
#pragma unroll
        for (int j = 0; j < TR; ++j) out[(ty * TR + i) * LD + tx * TR + j] = acc[i][j];
}

__global__ void gram_reduce_kernel(const float* __restrict__ part, int n_part, int ld, int l, double* __restrict__ G) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= l * l) return;
    const int i = idx / l, j = idx % l;
    double s = 0.0;
    for (int p = 0; p < n_part; ++p) s += (double)part[(size_t)p * ld * ld + i * ld + j];
    G[idx] = s;
}

}  // namespace mub

extern "C" {

size_t mub_gram_workspace_bytes(int64_t n, int32_t ld) {
    if (n <= 0 || ld <= 0) return 0;
    return (size_t)mub::gram_grid(n) * ld * ld * sizeof(float);
}

int mub_gram_f32(const float* Y, const float* weights, int64_t n, int32_t ld, int32_t l, double* G,
                 void* workspace, mub_stream_t stream) {
    MUB_REQUIRE(ld == 32 || ld == 64 || ld == 128, "gram: ld must be 32, 64 or 128 (got %d)", ld);
    MUB_REQUIRE(l >= 1 && l <= ld, "gram: need 1 <= l <= ld");
    MUB_REQUIRE(n >= 0 && G, "gram: bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0) {
        cudaMemsetAsync(G, 0, sizeof(double) * l * l, s);
        return 0;
    }
    MUB_REQUIRE(Y && workspace, "gram: null pointer");
    MUB_REQUIRE(((uintptr_t)Y & 15) == 0, "gram: Y must be 16-byte aligned");
    const int grid = mub::gram_grid(n);
    float* part = (float*)workspace;
    switch (ld) {
        case 32: mub::gram_partial_kernel<32><<<grid, mub::kGramThreads, 0, s>>>(Y, weights, n, part); break;
        case 64: mub::gram_partial_kernel<64><<<grid, mub::kGramThreads, 0, s>>>(Y, weights, n, part); break;
        default: mub::gram_partial_kernel<128><<<grid, mub::kGramThreads, 0, s>>>(Y, weights, n, part); break;
    }
    int rc = mub::check_launch("gram_partial");
    if (rc) return rc;
    mub::gram_reduce_kernel<<<(l * l + 255) / 256, 256, 0, s>>>(part, grid, ld, l, G);
    return mub::check_launch("gram_reduce");
}

}  // extern "C"
