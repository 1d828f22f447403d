// Stand-alone probe of the tcgen05.mma kind::tf32 operand layouts used by k4_dense_kernel (no-swizzle, A MN-major,
// B K-major, M = 128, N = 64, K = 32 in four k-steps): A[m][k] = (m % 7 + 1) * (k + 1) is not needed -- simple patterns:
// A[m][k] = m + 1 if k == kk else 0 ; B[n][k] = n + 1 if k == kk else 0  =>  D[m][n] = (m + 1)(n + 1).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int M = 128, N = 64, K = 32;
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned long long umma_desc(unsigned saddr, unsigned lbo, unsigned sbo) {
    return (unsigned long long)((saddr >> 4) & 0x3FFFu) | ((unsigned long long)((lbo >> 4) & 0x3FFFu) << 16) |
           ((unsigned long long)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__device__ __forceinline__ int da_off(int m, int k) { return (k >> 3) * 4096 + (m >> 2) * 128 + (k & 7) * 16 + (m & 3) * 4; }
__device__ __forceinline__ int db_off(int n, int k) { return (n >> 3) * 1024 + (k >> 2) * 128 + (n & 7) * 16 + (k & 3) * 4; }

__global__ void __launch_bounds__(128) probe(float* out, int variant) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long mbar;
    __shared__ unsigned s_tmem;
    unsigned char* sA = smem;
    unsigned char* sB = smem + M * K * 4;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&s_tmem)), "n"(64) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < M * K; i += 128) {
        const int k = i / M, m = i % M;
        const int off = (variant == 4 || variant == 6) ? db_off(m, k) : da_off(m, k);     // 4: A stored K-major like B
        *reinterpret_cast<float*>(sA + off) = (float)((m % 16) + 1) * (float)(k % 4 + 1);
    }
    for (int i = tid; i < N * K; i += 128) {
        const int n = i / K, k = i % K;
        *reinterpret_cast<float*>(sB + db_off(n, k)) = (float)(n % 8 + 1) * (k < 8 ? 1.0f : 0.0f);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = s_tmem;
    constexpr unsigned IDESC = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (0u << 16) | ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
    long long t0 = clock64();
    if (variant == 7) {
        unsigned w[16];
        for (int j = 0; j < 16; ++j) w[j] = __float_as_uint((float)(tid * 100 + j));
        const unsigned ta = tmem + ((unsigned)(warp * 32) << 16);
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                     :: "r"(ta), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]),
                        "r"(w[8]), "r"(w[9]), "r"(w[10]), "r"(w[11]), "r"(w[12]), "r"(w[13]), "r"(w[14]), "r"(w[15]) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    if (tid == 0 && variant != 7) {
        for (int ks = 0; ks < K / 8; ++ks) {
            unsigned long long da, db;
            if (variant == 0) { da = umma_desc(smem_u32(sA) + ks * 4096, 4096, 128); db = umma_desc(smem_u32(sB) + ks * 256, 128, 1024); }
            else if (variant == 1) { da = umma_desc(smem_u32(sA) + ks * 4096, 128, 4096); db = umma_desc(smem_u32(sB) + ks * 256, 1024, 128); }
            else if (variant == 2) { da = umma_desc(smem_u32(sA) + ks * 4096, 4096, 128); db = umma_desc(smem_u32(sB) + ks * 256, 1024, 128); }
            else if (variant == 3) { da = umma_desc(smem_u32(sA) + ks * 4096, 128, 4096); db = umma_desc(smem_u32(sB) + ks * 256, 128, 1024); }
            else { da = umma_desc(smem_u32(sA) + ks * 256, 128, 1024); db = umma_desc(smem_u32(sB) + ks * 256, 128, 1024); }
            const unsigned en = ks ? 1u : 0u;
            const unsigned idesc = (variant == 4 || variant == 6) ? (IDESC & ~(1u << 15)) : IDESC;
            if (variant == 5 || variant == 6) {
                unsigned z = 0u;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
                             :: "r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(en), "r"(z) : "memory");
            } else
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                         :: "r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(en) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&mbar)) : "memory");
    }
    if (variant != 7) {
        const unsigned bar = smem_u32(&mbar);
        asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" :: "r"(bar), "r"(0u) : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) printf("  [dev] variant %d tmem base 0x%x waited %lld clk\n", variant, tmem, clock64() - t0);
    const unsigned taddr = tmem + ((unsigned)(warp * 32) << 16);
    for (int cc = 0; cc < N; cc += 16) {
        unsigned v[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                       "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                     : "r"(taddr + (unsigned)cc));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 16; ++j) out[tid * N + cc + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "n"(64) : "memory");
}

int main() {
    float* d;
    cudaMalloc(&d, M * N * 4);
    static float h[M * N];
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int variant = 0; variant < 8; ++variant) {
        cudaMemset(d, 0xFF, M * N * 4);
        probe<<<1, 128, (M * K + N * K) * 4 + 1024>>>(d, variant);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h, d, M * N * 4, cudaMemcpyDeviceToHost);
        // expected: D[m][n] = sum_k A[m][k] B[n][k] = (m%16+1)(n%8+1) * sum_{k<8} (k%4+1) = (m%16+1)(n%8+1)*20
        int bad = 0;
        for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) if (h[m * N + n] != (float)((m % 16 + 1) * (n % 8 + 1) * 20)) ++bad;
        printf("variant %d: %s  bad %d / %d   D[0][0..3] = %g %g %g %g  D[1][0] = %g D[5][3] = %g (expect %d; variant 7 = TMEM st/ld roundtrip: D[5][3] should be 503)\n", variant,
               cudaGetErrorString(e), bad, M * N, h[0], h[1], h[2], h[3], h[N], h[5 * N + 3], 6 * 4 * 20);
    }
    return 0;
}
