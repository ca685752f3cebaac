// EXPERIMENTAL, NOT PART OF libmuon_b200.so.  The pipeline below ran once on a B200 at the very end of round 1
// (profiles/knn_tc_v1_vs_v2_100k.json: output bit-identical to v1, same time); the batched compaction added afterwards
// (v2_compact) has not run on hardware yet.
//
// Round-2 candidate for the tensor-core kNN candidate pass (knn_tc.cu: knn_tc_candidates_kernel, v1).  v1 is correct
// and bit-identical to the SIMT kernel but spends ~58 k cycles per 128 x 128 tile where MMA + epilogue need ~3 k:
// it stages every candidate tile with plain loads while all CTAs walk the same addresses in lockstep, and nothing
// overlaps.  v2 keeps v1's numerics (3xTF32 operands, error-band threshold, radix-select compaction, same re-rank)
// and changes the data movement:
//   * candidate tiles are 64 points (N = 64), pre-packed like v1, fetched by ONE 1-D TMA bulk copy each
//     (cp.async.bulk + mbarrier complete_tx, the pattern verified in spmm_panel.cu) into a 2-stage ring;
//   * two TMEM accumulator buffers: tcgen05.mma of tile t runs while the 128 threads filter tile t-1;
//   * every CTA starts its sweep over the candidate tiles at a different offset (no lockstep L2 hot spot);
//   * (added after the first hardware run showed v2 == v1 in time) the candidate-buffer compaction issues its
//     global loads in batches of 8 (v2_compact) -- NOT YET RUN ON HARDWARE.
// This file is a stand-alone harness: it compiles v1 (by including ../knn_tc.cu) and v2 into one binary, runs both on
// the same synthetic points and reports time and whether the final neighbour lists are identical.
//
//   make -C muon_b200/csrc probe2 && muon_b200/csrc/experimental/knn_tc_v2 [n=100000] [d=50] [k=201]
//
// All waits are bounded (a logic error shows up as status 4 / a mismatch, not as a hung GPU).
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../api.cu"
#include "../knn_tc.cu"

namespace mub {

constexpr int kV2N = 64;                 // candidates per MMA tile

__device__ __forceinline__ void v2_mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void v2_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc_smem_u32(bar)), "r"(bytes));
}
__device__ __forceinline__ bool v2_mbar_wait(uint64_t* bar, uint32_t parity) {      // bounded
    uint32_t done = 0;
    for (int spin = 0; spin < (1 << 22) && !done; ++spin) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done)
                     : "r"(tc_smem_u32(bar)), "r"(parity)
                     : "memory");
    }
    return done != 0;
}
__device__ __forceinline__ void v2_tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(tc_smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(tc_smem_u32(bar))
                 : "memory");
}

// pack with R rows per tile (v1's kernel is fixed at 128)
template <int R>
__global__ void v2_pack_kernel(const float* __restrict__ X, int64_t n, int d, int ld, int Kp, int split, float4* __restrict__ pk) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n_pad = (n + R - 1) / R * R;
    if (i >= n_pad) return;
    const int64_t tile = i / R;
    const int r = (int)(i % R);
    const int chunks = Kp / 4, chunks_total = split ? 3 * chunks : chunks;
    const bool live = i < n;
    for (int c = 0; c < chunks; ++c) {
        float v[4], big[4], small[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int dim = c * 4 + e;
            v[e] = (live && dim < d) ? X[(size_t)i * ld + dim] : 0.f;
            big[e] = __uint_as_float(__float_as_uint(v[e]) & 0xFFFFE000u);
            small[e] = v[e] - big[e];
        }
        float4* dst = pk + ((size_t)tile * chunks_total + c) * R + r;
        if (!split) {
            *dst = make_float4(v[0], v[1], v[2], v[3]);
        } else {
            const float4 b4 = make_float4(big[0], big[1], big[2], big[3]);
            const float4 s4 = make_float4(small[0], small[1], small[2], small[3]);
            dst[0] = b4;
            dst[(size_t)chunks * R] = (split == 1) ? b4 : s4;
            dst[(size_t)2 * chunks * R] = (split == 1) ? s4 : b4;
        }
    }
}

// tc_compact with the buffer loads batched: 8 independent L2 loads are issued before the first histogram update, so a
// pass costs ~cnt/8 memory round trips instead of ~cnt (the measured limiter of v1).
__device__ __forceinline__ void v2_compact(float* __restrict__ kb, int32_t* __restrict__ ib, int32_t* __restrict__ hist,
                                           int tid, int k, float slack, int& cnt, float& tau) {
    if (cnt < k) return;
    uint32_t H = 0;
    int below = 0;
#pragma unroll 1
    for (int s = 16; s >= 0; s -= 4) {
#pragma unroll
        for (int j = 0; j < 16; ++j) hist[j * kTcM + tid] = 0;
#pragma unroll 1
        for (int i0 = 0; i0 < cnt; i0 += 8) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = (i0 + u < cnt) ? __ldcg(kb + (size_t)(i0 + u) * kTcM + tid) : 0.f;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (i0 + u < cnt) {
                    const uint32_t p = tc_okey(v[u]) >> 12;
                    if ((p >> (s + 4)) == H) hist[((p >> s) & 15u) * kTcM + tid] += 1;
                }
            }
        }
        int acc = below, D = 15;
        bool found = false;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int hj = hist[j * kTcM + tid];
            if (!found) {
                if (acc + hj >= k) { D = j; found = true; }
                else acc += hj;
            }
        }
        below = acc;
        H = (H << 4) | (uint32_t)D;
    }
    float base = tc_from_okey((H << 12) | 0xFFFu);
    if (!(base <= FLT_MAX)) base = FLT_MAX;
    tau = base + slack;
    int w = 0;
#pragma unroll 1
    for (int i0 = 0; i0 < cnt; i0 += 8) {
        float v[8];
        int32_t jj[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const bool in = i0 + u < cnt;
            v[u] = in ? __ldcg(kb + (size_t)(i0 + u) * kTcM + tid) : INFINITY;
            jj[u] = in ? __ldcg(ib + (size_t)(i0 + u) * kTcM + tid) : -1;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (i0 + u < cnt && v[u] <= tau) {                  // w <= i0 + u: never overwrites an unread entry
                kb[(size_t)w * kTcM + tid] = v[u];
                ib[(size_t)w * kTcM + tid] = jj[u];
                ++w;
            }
        }
    }
    cnt = w;
}

// filter one 128 x 64 accumulator buffer (thread = TMEM lane = query)
__device__ __forceinline__ void v2_epilogue(uint32_t tmem_buf, int warp, const float* __restrict__ cn, int64_t c0, int64_t nc,
                                            float tau, float* __restrict__ kb, int32_t* __restrict__ ib, int tid, int& cnt,
                                            int& overflow) {
#pragma unroll
    for (int cc = 0; cc < kV2N; cc += 32) {
        uint32_t r[32];
        const uint32_t taddr = tmem_buf + ((uint32_t)(warp * 32) << 16) + (uint32_t)cc;
        __syncwarp();
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
            "%15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
              "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const float a = fmaf(-2.f, __uint_as_float(r[j]), cn[cc + j]);
            if (a <= tau && c0 + cc + j < nc) {
                if (cnt < kTcCap) {
                    kb[(size_t)cnt * kTcM + tid] = a;
                    ib[(size_t)cnt * kTcM + tid] = (int32_t)(c0 + cc + j);
                    ++cnt;
                } else {
                    overflow = 1;
                }
            }
        }
    }
}

__global__ void __launch_bounds__(kTcThreads)
knn_tc_candidates_v2_kernel(const float4* __restrict__ Xpk, const float* __restrict__ xnorm, int64_t nq,
                            const float4* __restrict__ Ypk64, const float* __restrict__ ynorm, int64_t nc, int Kp, int k,
                            float slack_rel, const unsigned int* __restrict__ ymax_bits, float* __restrict__ kbuf,
                            int32_t* __restrict__ ibuf, int32_t* __restrict__ cand, int32_t* __restrict__ cand_cnt,
                            int32_t* __restrict__ status) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t full_bar[2];       // TMA landed the candidate tile of stage s
    __shared__ __align__(8) uint64_t mma_bar[2];        // the MMAs into TMEM buffer s completed
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int chunks = Kp / 4;
    const uint32_t tile_bytes = (uint32_t)chunks * kV2N * 16;
    unsigned char* sQ = smem;                                        // [chunks][128][16 B]
    unsigned char* sC0 = smem + (size_t)chunks * kTcM * 16;          // 2 stages of [chunks][64][16 B]
    float* sCn = reinterpret_cast<float*>(sC0 + 2 * (size_t)tile_bytes);   // [2][64]
    int32_t* hist = reinterpret_cast<int32_t*>(sCn + 2 * kV2N);             // [16][128]

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(&tmem_base_s)), "r"(2 * kV2N));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        v2_mbar_init(&full_bar[0], 1);
        v2_mbar_init(&full_bar[1], 1);
        v2_mbar_init(&mma_bar[0], 1);
        v2_mbar_init(&mma_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem_base = tmem_base_s;

    float* kb = kbuf + (size_t)blockIdx.x * kTcCap * kTcM;
    int32_t* ib = ibuf + (size_t)blockIdx.x * kTcCap * kTcM;
    const float cmax = sqrtf(__uint_as_float(*ymax_bits));
    uint32_t idesc = 0;
    idesc |= 1u << 4;
    idesc |= 2u << 7;
    idesc |= 2u << 10;
    idesc |= (uint32_t)(kV2N >> 3) << 17;
    idesc |= (uint32_t)(kTcM >> 4) << 24;

    const int64_t n_qt = (nq + kTcM - 1) / kTcM;
    const int64_t T = (nc + kV2N - 1) / kV2N;                       // candidate tiles
    const int64_t off = T ? ((int64_t)blockIdx.x * T) / gridDim.x : 0;   // staggered start
    uint32_t it = 0;            // candidate tiles processed so far by this CTA: stage = it & 1, parity = (it >> 1) & 1
    bool dead = false;

    for (int64_t qt = blockIdx.x; qt < n_qt && !dead; qt += gridDim.x) {
        const int64_t row = qt * kTcM + tid;
        const bool live = row < nq;
        {
            const float4* src = Xpk + (size_t)qt * chunks * kTcM;
            float4* dst = reinterpret_cast<float4*>(sQ);
#pragma unroll 4
            for (int i = tid; i < chunks * kTcM; i += kTcThreads) dst[i] = __ldg(src + i);
            asm volatile("fence.proxy.async.shared::cta;");          // generic stores -> async proxy
        }
        const float qn = live ? xnorm[row] : 0.f;
        const float slack = slack_rel * sqrtf(qn) * cmax + 1e-5f * (qn + cmax * cmax);
        float tau = live ? INFINITY : -INFINITY;
        int cnt = 0, overflow = 0;
        if (T == 0) {
            if (live) cand_cnt[row] = 0;
            continue;
        }
        // prologue: fetch the first tile of this sweep
        if (tid == 0) {
            const int64_t ct = off % T;
            uint64_t* bar = &full_bar[it & 1];
            v2_mbar_expect_tx(bar, tile_bytes);
            v2_tma_bulk_g2s(sC0 + (size_t)(it & 1) * tile_bytes, Ypk64 + (size_t)ct * chunks * kV2N, tile_bytes, bar);
        }
        for (int64_t t = 0; t <= T; ++t) {
            // all threads finished filtering tile t-2 (TMEM buffer and norm slot about to be reused) and, on the first
            // iteration, finished writing sQ
            asm volatile("tcgen05.fence::before_thread_sync;");
            __syncthreads();
            asm volatile("tcgen05.fence::after_thread_sync;");
            if (t >= 1) {
                // wait for the MMAs of tile t-1 (also frees its shared-memory stage for the next TMA)
                const uint32_t pit = it - 1;
                if (__syncthreads_or(v2_mbar_wait(&mma_bar[pit & 1], (pit >> 1) & 1u) ? 0 : 1)) { dead = true; break; }
                asm volatile("tcgen05.fence::after_thread_sync;");
            }
            if (t < T) {
                const int64_t ct = (t + off) % T;
                if (t + 1 < T && tid == 0) {                          // prefetch tile t+1 into the other stage
                    const int64_t ct1 = (t + 1 + off) % T;
                    const uint32_t nit = it + 1;
                    uint64_t* bar = &full_bar[nit & 1];
                    v2_mbar_expect_tx(bar, tile_bytes);
                    v2_tma_bulk_g2s(sC0 + (size_t)(nit & 1) * tile_bytes, Ypk64 + (size_t)ct1 * chunks * kV2N, tile_bytes, bar);
                }
                if (tid < kV2N) {
                    const int64_t crow = ct * kV2N + tid;
                    sCn[(it & 1) * kV2N + tid] = crow < nc ? ynorm[crow] : INFINITY;
                }
                if (__syncthreads_or(v2_mbar_wait(&full_bar[it & 1], (it >> 1) & 1u) ? 0 : 1)) { dead = true; break; }
                if (tid == 0) {
                    asm volatile("tcgen05.fence::after_thread_sync;");
                    const uint32_t a_lbo = kTcM * 16, b_lbo = kV2N * 16, sbo = 128;
                    const uint32_t sCa = tc_smem_u32(sC0 + (size_t)(it & 1) * tile_bytes);
                    const uint32_t tbuf = tmem_base + (it & 1) * kV2N;
                    for (int ks = 0; ks < Kp / 8; ++ks) {
                        const uint64_t da = tc_umma_desc(tc_smem_u32(sQ) + (uint32_t)(2 * ks) * kTcM * 16, a_lbo, sbo);
                        const uint64_t db = tc_umma_desc(sCa + (uint32_t)(2 * ks) * kV2N * 16, b_lbo, sbo);
                        const uint32_t acc = ks > 0 ? 1u : 0u;
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                     "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tbuf),
                                     "l"(da), "l"(db), "r"(idesc), "r"(acc));
                    }
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                                     tc_smem_u32(&mma_bar[it & 1]))
                                 : "memory");
                }
            }
            if (t >= 1) {                                             // filter tile t-1 while the MMAs of tile t run
                const uint32_t pit = it - 1;
                if (__any_sync(0xffffffffu, cnt > kTcCap - kV2N)) {
                    v2_compact(kb, ib, hist, tid, k, slack, cnt, tau);
                    if (cnt > kTcCap - kV2N) {
                        overflow = 1;
                        cnt = kTcCap - kV2N;
                        atomicOr(status, 2);
                    }
                }
                const int64_t pct = (t - 1 + off) % T;
                v2_epilogue(tmem_base + (pit & 1) * kV2N, warp, sCn + (pit & 1) * kV2N, pct * kV2N, nc, tau, kb, ib, tid, cnt,
                            overflow);
            }
            if (t < T) ++it;
        }
        if (dead) break;
        v2_compact(kb, ib, hist, tid, k, slack, cnt, tau);
        if (live) {
            cand_cnt[row] = cnt;
            for (int i = 0; i < cnt; ++i) cand[(size_t)row * kTcCap + i] = ib[(size_t)i * kTcM + tid];
            if (overflow) atomicOr(status, 2);
        }
    }
    if (dead && tid == 0) atomicOr(status, 4);
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * kV2N));
    }
}

}  // namespace mub

// ------------------------------------------------------------------------------------------------------------------
__global__ void synth_points(float* X, int64_t n, int d, int n_clusters) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    auto h = [](uint64_t x) {
        x += 0x9E3779B97F4A7C15ull;
        x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
        x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
        return x ^ (x >> 31);
    };
    auto gauss = [&](uint64_t key) {                // sum of 4 uniforms, variance-normalised
        float s = 0.f;
        for (int t = 0; t < 4; ++t) s += (float)(h(key * 4 + t) >> 40) * (1.0f / 16777216.0f);
        return (s - 2.0f) * 1.7320508f;
    };
    const int c = (int)(h((uint64_t)i * 7919 + 13) % (uint64_t)n_clusters);
    float nrm = 0.f;
    for (int t = 0; t < d; ++t) {
        const float v = gauss((uint64_t)i * 131 + t) + 3.0f * gauss(0x100000000ull + (uint64_t)c * 131 + t);
        X[(size_t)i * d + t] = v;
        nrm += v * v;
    }
    nrm = rsqrtf(nrm);
    for (int t = 0; t < d; ++t) X[(size_t)i * d + t] *= nrm;
}

#define CK(x)                                                                       \
    do {                                                                            \
        cudaError_t e_ = (x);                                                       \
        if (e_ != cudaSuccess) {                                                    \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            return 2;                                                               \
        }                                                                           \
    } while (0)

int main(int argc, char** argv) {
    const int64_t n = argc > 1 ? atoll(argv[1]) : 100000;
    const int d = argc > 2 ? atoi(argv[2]) : 50;
    const int k = argc > 3 ? atoi(argv[3]) : 201;
    if (d > 64) {
        printf("harness covers the 3xTF32 variant only (d <= 64)\n");
        return 1;
    }
    float* X;
    CK(cudaMalloc(&X, sizeof(float) * n * d));
    synth_points<<<(unsigned)((n + 255) / 256), 256>>>(X, n, d, 30);
    CK(cudaDeviceSynchronize());

    // ---- v1 through the library entry point -----------------------------------------------------------------
    const size_t wsb = mub_knn_l2_tc_workspace_bytes(n, n, d);
    void* ws;
    int32_t *idx1, *idx2, *status;
    float *dist1, *dist2;
    CK(cudaMalloc(&ws, wsb));
    CK(cudaMalloc(&idx1, sizeof(int32_t) * n * k));
    CK(cudaMalloc(&idx2, sizeof(int32_t) * n * k));
    CK(cudaMalloc(&dist1, sizeof(float) * n * k));
    CK(cudaMalloc(&dist2, sizeof(float) * n * k));
    CK(cudaMalloc(&status, 4));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float ms1 = 0, ms2 = 0;
    int st1 = 0, st2 = 0;
    for (int rep = 0; rep < 2; ++rep) {
        CK(cudaMemset(status, 0, 4));
        cudaEventRecord(e0);
        if (mub_knn_l2_tc_f32(X, n, X, n, d, d, k, idx1, dist1, ws, wsb, status, nullptr)) {
            printf("v1 failed: %s\n", mub_last_error());
            return 2;
        }
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        cudaEventElapsedTime(&ms1, e0, e1);
    }
    CK(cudaMemcpy(&st1, status, 4, cudaMemcpyDeviceToHost));

    // ---- v2: same workspace layout for norms / buffers, own packed candidates --------------------------------
    const mub::TcLayout L = mub::tc_layout(n, n, d);
    unsigned char* w = (unsigned char*)ws;
    float* xnorm = (float*)(w + L.xnorm);
    float* ynorm = (float*)(w + L.ynorm);
    unsigned int* ymax = (unsigned int*)(w + L.ymax);
    int32_t* cnt = (int32_t*)(w + L.cnt);
    int32_t* cand = (int32_t*)(w + L.cand);
    float* kbuf = (float*)(w + L.kbuf);
    int32_t* ibuf = (int32_t*)(w + L.ibuf);
    float4* xpk = (float4*)(w + L.xpk);                 // still holds v1's packed queries (128-row tiles, [b|b|s])
    float4* ypk64;
    const int64_t n_pad64 = (n + 63) / 64 * 64;
    CK(cudaMalloc(&ypk64, sizeof(float) * n_pad64 * L.Ktot));
    const size_t smem = (size_t)(L.Ktot / 4) * (mub::kTcM + 2 * mub::kV2N) * 16 + 2 * mub::kV2N * sizeof(float) +
                        16 * mub::kTcM * sizeof(int32_t);
    CK(cudaFuncSetAttribute(mub::knn_tc_candidates_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int grid = mub::sm_count();
    const int64_t n_qt = (n + mub::kTcM - 1) / mub::kTcM;
    if (grid > n_qt) grid = (int)n_qt;                   // kbuf/ibuf are sized for 2 * SMs CTAs: enough
    for (int rep = 0; rep < 2; ++rep) {
        CK(cudaMemset(status, 0, 4));
        cudaEventRecord(e0);
        mub::v2_pack_kernel<64><<<(unsigned)(n_pad64 / 128 + 1), 128>>>(X, n, d, d, L.Kp, 2, ypk64);
        mub::knn_tc_candidates_v2_kernel<<<grid, mub::kTcThreads, smem>>>(xpk, xnorm, n, ypk64, ynorm, n, L.Ktot, k, 1.2e-4f, ymax,
                                                                         kbuf, ibuf, cand, cnt, status);
        mub::knn_tc_rerank_kernel<<<(unsigned)((n + mub::kTcRerankWarps - 1) / mub::kTcRerankWarps), mub::kTcRerankWarps * 32>>>(
            X, n, X, d, d, k, cand, cnt, idx2, dist2);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        cudaEventElapsedTime(&ms2, e0, e1);
    }
    CK(cudaMemcpy(&st2, status, 4, cudaMemcpyDeviceToHost));

    std::vector<int32_t> h1((size_t)n * k), h2((size_t)n * k);
    std::vector<float> g1((size_t)n * k), g2((size_t)n * k);
    CK(cudaMemcpy(h1.data(), idx1, sizeof(int32_t) * n * k, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(h2.data(), idx2, sizeof(int32_t) * n * k, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(g1.data(), dist1, sizeof(float) * n * k, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(g2.data(), dist2, sizeof(float) * n * k, cudaMemcpyDeviceToHost));
    size_t bad = 0;
    for (size_t i = 0; i < h1.size(); ++i) bad += (h1[i] != h2[i]) || (memcmp(&g1[i], &g2[i], 4) != 0);
    printf("{\"n\": %lld, \"d\": %d, \"k\": %d, \"v1_ms\": %.3f, \"v1_status\": %d, \"v2_ms\": %.3f, \"v2_status\": %d, "
           "\"mismatching_entries\": %zu, \"smem_v2\": %zu}\n",
           (long long)n, d, k, ms1, st1, ms2, st2, bad, smem);
    return 0;
}
