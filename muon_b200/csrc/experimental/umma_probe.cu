// Stand-alone probe (NOT part of libmuon_b200.so): one tcgen05.mma tile, D[128 x N] = A[128 x K] . B[N x K]^T, operands written
// to shared memory by plain stores in the canonical no-swizzle K-major layout, accumulator in TMEM, read back with
// tcgen05.ld.  Purpose: pin down the descriptor encodings (shared-memory matrix descriptor, instruction descriptor,
// TMEM addressing) on real sm_100a hardware before the round-2 kNN candidate kernel (the one GEMM-shaped op of the
// WNN row, reference muon/_core/preproc.py:520-528) is built on them.  Every wait is bounded: a wrong descriptor
// shows up as a numeric mismatch or a reported timeout, never as a hung GPU.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe umma_probe.cu && ./umma_probe
//
// Layout facts being tested (from the vendored CUTLASS headers, cute/arch/mma_sm100_desc.hpp and
// cute/atom/mma_traits_sm100.hpp, K-major "INTERLEAVE" = no swizzle):
//   * core matrix = 8 rows x 16 bytes, rows 16 B apart (128 B contiguous)
//   * SBO = byte distance between 8-row groups, LBO = byte distance between the two 16-byte K chunks of one MMA
//   * descriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout type [61,64) = 0
//   * instruction descriptor: c_format [4,6) (1 = f32), a/b_format [7,10)/[10,13) (1 = bf16, 2 = tf32),
//     a/b_major [15]/[16] (0 = K), N>>3 [17,23), M>>4 [24,29)
//   * accumulator: row i of D in TMEM lane i, column j in TMEM column base+j; warp w of the CTA may read lanes 32w..32w+31
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

constexpr int kM = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= 1ull << 46;                       // descriptor version 1 (Blackwell)
    return d;                              // base_offset 0, lbo_mode 0, layout type 0 (no swizzle)
}

// variant bit0: swap the roles of LBO and SBO (diagnostic)
template <int N, bool TF32>
__global__ void __launch_bounds__(128) umma_probe_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                         float* __restrict__ D, int K, int variant, int* status) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    constexpr int kElemBytes = TF32 ? 4 : 2;
    constexpr int kChunkElems = 16 / kElemBytes;          // elements per 16-byte K chunk
    constexpr int kUmmaK = 2 * kChunkElems;               // 32 bytes of K per instruction
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int chunks = K / kChunkElems;
    unsigned char* sA = smem;                              // [chunks][128 rows][16 B]
    unsigned char* sB = smem + (size_t)chunks * kM * 16;   // [chunks][N rows][16 B]

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(N < 32 ? 32 : N));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    // operands -> shared memory, canonical layout: offset(row, chunk) = chunk * rows * 16 + row * 16
    for (int c = 0; c < chunks; ++c) {
        {
            const float* src = A + (size_t)tid * K + c * kChunkElems;
            unsigned char* dst = sA + ((size_t)c * kM + tid) * 16;
            if (TF32) {
                *reinterpret_cast<float4*>(dst) = make_float4(src[0], src[1], src[2], src[3]);
            } else {
                __nv_bfloat16 v[8];
                for (int e = 0; e < 8; ++e) v[e] = __float2bfloat16(src[e]);
                *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<uint4*>(v);
            }
        }
        for (int r = tid; r < N; r += 128) {
            const float* src = B + (size_t)r * K + c * kChunkElems;
            unsigned char* dst = sB + ((size_t)c * N + r) * 16;
            if (TF32) {
                *reinterpret_cast<float4*>(dst) = make_float4(src[0], src[1], src[2], src[3]);
            } else {
                __nv_bfloat16 v[8];
                for (int e = 0; e < 8; ++e) v[e] = __float2bfloat16(src[e]);
                *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<uint4*>(v);
            }
        }
    }
    asm volatile("fence.proxy.async.shared::cta;");       // generic-proxy stores -> visible to the tensor core (async proxy)
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem_base = tmem_base_s;

    if (tid == 0) {
        uint32_t idesc = 0;
        idesc |= 1u << 4;                                  // D = f32
        idesc |= (TF32 ? 2u : 1u) << 7;                    // A format
        idesc |= (TF32 ? 2u : 1u) << 10;                   // B format
        idesc |= (uint32_t)(N >> 3) << 17;
        idesc |= (uint32_t)(kM >> 4) << 24;
        const uint32_t a_lbo = kM * 16, b_lbo = N * 16, sbo = 128;
        for (int ks = 0; ks < K / kUmmaK; ++ks) {
            const uint32_t a_addr = smem_u32(sA) + (uint32_t)(2 * ks) * kM * 16;
            const uint32_t b_addr = smem_u32(sB) + (uint32_t)(2 * ks) * N * 16;
            const uint64_t da = (variant & 1) ? make_desc(a_addr, sbo, a_lbo) : make_desc(a_addr, a_lbo, sbo);
            const uint64_t db = (variant & 1) ? make_desc(b_addr, sbo, b_lbo) : make_desc(b_addr, b_lbo, sbo);
            const uint32_t acc = ks > 0 ? 1u : 0u;
            if (TF32) {
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_base),
                             "l"(da), "l"(db), "r"(idesc), "r"(acc));
            } else {
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_base),
                             "l"(da), "l"(db), "r"(idesc), "r"(acc));
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    // bounded wait for the commit (phase 0)
    uint32_t done = 0;
    for (int spin = 0; spin < (1 << 23) && !done; ++spin) {      // test_wait never suspends: ~0.3 s worst case
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done)
                     : "r"(smem_u32(&bar)), "r"(0u)
                     : "memory");
    }
    done = __all_sync(0xffffffffu, done);                       // tcgen05.ld is .sync.aligned: keep each warp convergent
    if (!done && lane == 0) atomicExch(status, 1);
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (done) {
        for (int c0 = 0; c0 < N; c0 += 32) {
            uint32_t r[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                  "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                  "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                  "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const int row = warp * 32 + lane;
            for (int j = 0; j < 32 && c0 + j < N; ++j) D[(size_t)row * N + c0 + j] = __uint_as_float(r[j]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(N < 32 ? 32 : N));
    }
}

static float tf32_round(float x) {           // round-to-nearest-even to 10 mantissa bits (what cvt.rna.tf32 does)
    uint32_t u;
    memcpy(&u, &x, 4);
    u += 0xFFFu + ((u >> 13) & 1u);
    u &= 0xFFFFE000u;
    memcpy(&x, &u, 4);
    return x;
}
static float tf32_trunc(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    u &= 0xFFFFE000u;
    memcpy(&x, &u, 4);
    return x;
}
static float bf16_round(float x) { return __bfloat162float(__float2bfloat16(x)); }

template <int N, bool TF32>
static int run(int K, int variant, bool small_ints) {
    std::vector<float> A((size_t)kM * K), B((size_t)N * K), D((size_t)kM * N, -7.f);
    srand(1234 + K + N);
    for (auto& v : A) v = small_ints ? (float)(rand() % 5 - 2) : (float)rand() / RAND_MAX - 0.5f;
    for (auto& v : B) v = small_ints ? (float)(rand() % 5 - 2) : (float)rand() / RAND_MAX - 0.5f;
    float *dA, *dB, *dD;
    int* dS;
    cudaMalloc(&dA, A.size() * 4);
    cudaMalloc(&dB, B.size() * 4);
    cudaMalloc(&dD, D.size() * 4);
    cudaMalloc(&dS, 4);
    cudaMemset(dS, 0, 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dD, D.data(), D.size() * 4, cudaMemcpyHostToDevice);
    const int chunk_elems = TF32 ? 4 : 8;
    const size_t smem = (size_t)(K / chunk_elems) * (kM + N) * 16;
    auto kern = umma_probe_kernel<N, TF32>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<1, 128, smem>>>(dA, dB, dD, K, variant, dS);
    cudaError_t e = cudaDeviceSynchronize();
    int st = 0;
    if (e == cudaSuccess) {
        cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    }
    printf("%s N=%d K=%d variant=%d %s: ", TF32 ? "tf32" : "bf16", N, K, variant, small_ints ? "ints" : "rand");
    if (e != cudaSuccess) {
        printf("CUDA error %s\n", cudaGetErrorString(e));
        return 2;
    }
    if (st) {
        printf("TIMEOUT waiting for tcgen05.commit\n");
        return 2;
    }
    double err_exact = 0, err_rn = 0, err_tr = 0, ref_max = 0;
    for (int i = 0; i < kM; ++i)
        for (int j = 0; j < N; ++j) {
            double s = 0, s_rn = 0, s_tr = 0;
            for (int k = 0; k < K; ++k) {
                const float a = A[(size_t)i * K + k], b = B[(size_t)j * K + k];
                s += (double)a * b;
                if (TF32) {
                    s_rn += (double)tf32_round(a) * tf32_round(b);
                    s_tr += (double)tf32_trunc(a) * tf32_trunc(b);
                } else {
                    s_rn += (double)bf16_round(a) * bf16_round(b);
                    s_tr = s_rn;
                }
            }
            const double g = D[(size_t)i * N + j];
            err_exact = fmax(err_exact, fabs(g - s));
            err_rn = fmax(err_rn, fabs(g - s_rn));
            err_tr = fmax(err_tr, fabs(g - s_tr));
            ref_max = fmax(ref_max, fabs(s));
        }
    printf("max|D-ref| exact-inputs %.3e  rounded-inputs %.3e  truncated-inputs %.3e  (max|ref| %.3f)  %s\n", err_exact, err_rn,
           err_tr, ref_max, fmin(err_rn, err_tr) < 1e-3 * fmax(ref_max, 1.0) ? "OK" : "MISMATCH");
    const bool ok = fmin(err_rn, err_tr) < 1e-3 * fmax(ref_max, 1.0);
    if (!ok) {
        printf("   D[0][0..7]   =");
        for (int j = 0; j < 8; ++j) printf(" %9.4f", D[j]);
        printf("\n   D[1][0..7]   =");
        for (int j = 0; j < 8; ++j) printf(" %9.4f", D[N + j]);
        printf("\n   D[8][0..7]   =");
        for (int j = 0; j < 8; ++j) printf(" %9.4f", D[8 * N + j]);
        printf("\n   D[64][0..7]  =");
        for (int j = 0; j < 8; ++j) printf(" %9.4f", D[64 * N + j]);
        printf("\n");
    }
    cudaFree(dA);
    cudaFree(dB);
    cudaFree(dD);
    cudaFree(dS);
    return ok ? 0 : 1;
}

int main() {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, 0) != cudaSuccess) {
        printf("no CUDA device\n");
        return 3;
    }
    printf("device %s sm_%d%d\n", p.name, p.major, p.minor);
    int bad = 0;
    for (int variant = 0; variant < 2; ++variant) {
        bad += run<64, true>(8, variant, true) != 0;       // one instruction
        bad += run<64, true>(64, variant, true) != 0;      // 8 accumulating instructions
        bad += run<64, true>(64, variant, false) != 0;
        bad += run<128, true>(64, variant, false) != 0;
        bad += run<64, false>(64, variant, false) != 0;    // bf16
        bad += run<256, false>(128, variant, false) != 0;
    }
    printf("umma_probe: %d of 12 cases mismatched (variant 1 swaps LBO/SBO on purpose and is expected to fail)\n", bad);
    return 0;
}
