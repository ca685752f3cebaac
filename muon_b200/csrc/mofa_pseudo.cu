// MOFA non-gaussian likelihoods: Seeger pseudo-data and log-likelihood of a dense view.
//
// muon passes likelihoods=None to mofapy2, which guesses "poisson" for integer views and "bernoulli" for binary
// ones (muon/_core/tools.py:272-280).  mofapy2 then replaces the view, every iteration, by gaussian pseudo-data
// around zeta = E[Z] E[W]^T with a fixed precision kappa_d (Seeger & Bouchard 2012; mofapy2's Poisson_PseudoY /
// Bernoulli_PseudoY / Tau_Seeger nodes, restated in oracle/mofa_ref.py::mofa_ref_general):
//   poisson   : rate(z) = ln(1+e^z);  yhat = zeta - sigmoid(zeta) (1 - y/rate(zeta)) / kappa_d
//   bernoulli : yhat = zeta - (sigmoid(zeta) - y) / kappa_d
// The pseudo-data are dense N x D by construction (zeta is), so these views live as dense row-major fp32
// matrices; the contractions with them are plain GEMMs (cuBLAS through torch), this file holds the two
// elementwise passes around them: 12 B per element, HBM-bound.
#include <math.h>

#include "common.cuh"

namespace mub {

__device__ __forceinline__ float softplusf(float z) { return z > 20.f ? z : log1pf(__expf(z)); }
__device__ __forceinline__ float sigmoidf(float z) { return 1.f / (1.f + __expf(-z)); }

// kind: 1 = poisson, 2 = bernoulli.  zeta is overwritten with the pseudo-data.
__global__ void __launch_bounds__(256)
mofa_pseudo_kernel(float* __restrict__ zeta, const float* __restrict__ obs, const float* __restrict__ kappa,
                   int64_t n_elem, int32_t D, int32_t kind) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_elem; i += stride) {
        const float z = zeta[i], y = ld_stream(obs + i);
        const float k = __ldg(kappa + (int)(i % D));
        float out;
        if (kind == 1) {
            const float rate = fmaxf(softplusf(z), 1e-30f);
            out = z - sigmoidf(z) * (1.f - y / rate) / k;
        } else {
            out = z - (sigmoidf(z) - y) / k;
        }
        zeta[i] = out;
    }
}

// sum over elements of the log-likelihood at zeta (poisson: y ln rate - rate; bernoulli: y zeta - ln(1+e^zeta)),
// accumulated in double into *out
__global__ void __launch_bounds__(256)
mofa_loglik_kernel(const float* __restrict__ zeta, const float* __restrict__ obs, int64_t n_elem, int32_t kind,
                   double* out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_elem; i += stride) {
        const float z = ld_stream(zeta + i), y = ld_stream(obs + i);
        if (kind == 1) {
            const float rate = fmaxf(softplusf(z), 1e-30f);
            acc += (double)(y * logf(rate) - rate);
        } else {
            acc += (double)(y * z - softplusf(z));
        }
    }
    acc = warp_sum(acc);
    __shared__ double part[8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) part[w] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int j = 0; j < 8; ++j) s += part[j];
        atomicAdd(out, s);
    }
}

}  // namespace mub

extern "C" {

int mub_mofa_pseudo_f32(float* zeta, const float* obs, const float* kappa, int64_t n_rows, int32_t D, int32_t kind,
                        mub_stream_t stream) {
    MUB_REQUIRE(kind == 1 || kind == 2, "mofa_pseudo: kind must be 1 (poisson) or 2 (bernoulli)");
    MUB_REQUIRE(n_rows >= 0 && D >= 0, "mofa_pseudo: negative shape");
    const int64_t n = n_rows * (int64_t)D;
    if (n == 0) return 0;
    MUB_REQUIRE(zeta && obs && kappa, "mofa_pseudo: null pointer");
    int64_t want = (n + 255) / 256, cap = (int64_t)mub::sm_count() * 16;
    mub::mofa_pseudo_kernel<<<(int)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>(zeta, obs, kappa, n, D, kind);
    return mub::check_launch("mofa_pseudo");
}

int mub_mofa_loglik_f32(const float* zeta, const float* obs, int64_t n_rows, int32_t D, int32_t kind, double* out,
                        mub_stream_t stream) {
    MUB_REQUIRE(kind == 1 || kind == 2, "mofa_loglik: kind must be 1 (poisson) or 2 (bernoulli)");
    MUB_REQUIRE(n_rows >= 0 && D >= 0 && out, "mofa_loglik: bad argument");
    const int64_t n = n_rows * (int64_t)D;
    if (n == 0) return 0;
    MUB_REQUIRE(zeta && obs, "mofa_loglik: null pointer");
    int64_t want = (n + 255) / 256, cap = (int64_t)mub::sm_count() * 16;
    mub::mofa_loglik_kernel<<<(int)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>(zeta, obs, n, kind, out);
    return mub::check_launch("mofa_loglik");
}

}  // extern "C"
