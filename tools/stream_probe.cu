// stream_probe.cu — how fast can one B200 stream an 80 MB event buffer into the SMs?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/stream_probe tools/stream_probe.cu && /tmp/stream_probe
//
// Variants: direct 128-bit loads vs. a TMA (1-D bulk copy) ring into shared memory, contiguous
// per-CTA spans vs. chunks interleaved across the grid, ring depth and CTAs/SM swept.  The result
// decides the staging scheme of K1.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

#define CK(x)                                                                      \
    do {                                                                           \
        cudaError_t e = (x);                                                       \
        if (e != cudaSuccess) {                                                    \
            printf("%s failed: %s\n", #x, cudaGetErrorString(e));                  \
            return 1;                                                              \
        }                                                                          \
    } while (0)

__device__ __forceinline__ unsigned smem_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint64_t evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ int4 ld_stream(const int4* p, uint64_t pol) {
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.s32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p), "l"(pol));
    return r;
}

// ---- direct loads ----------------------------------------------------------------------------
template <int U, bool INTERLEAVED>
__global__ void __launch_bounds__(256) k_ldg(const int4* __restrict__ buf, long long n, unsigned* sink) {
    const uint64_t pol = evict_first();
    unsigned acc = 0;
    const int chunk = 256 * U;
    if (INTERLEAVED) {
        for (long long base = (long long)blockIdx.x * chunk; base < n; base += (long long)gridDim.x * chunk) {
            int4 v[U];
#pragma unroll
            for (int k = 0; k < U; ++k) {
                long long i = base + k * 256 + threadIdx.x;
                v[k] = i < n ? ld_stream(buf + i, pol) : make_int4(0, 0, 0, 0);
            }
#pragma unroll
            for (int k = 0; k < U; ++k) acc ^= v[k].x ^ v[k].y ^ v[k].z ^ v[k].w;
        }
    } else {
        const long long per = ((n + gridDim.x - 1) / gridDim.x + 31) & ~31LL;
        const long long lo = per * blockIdx.x < n ? per * blockIdx.x : n;
        const long long hi = lo + per < n ? lo + per : n;
        for (long long base = lo; base < hi; base += chunk) {
            int4 v[U];
#pragma unroll
            for (int k = 0; k < U; ++k) {
                long long i = base + k * 256 + threadIdx.x;
                v[k] = i < hi ? ld_stream(buf + i, pol) : make_int4(0, 0, 0, 0);
            }
#pragma unroll
            for (int k = 0; k < U; ++k) acc ^= v[k].x ^ v[k].y ^ v[k].z ^ v[k].w;
        }
    }
    if (acc == 0x12345678u) *sink = acc;
}

// ---- TMA ring -----------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(
            smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_1d(void* dst, const void* src, unsigned bytes, uint64_t* bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
                 : "memory");
}

// chunk = CH int4 records; ring of S stages; 256 threads consume each stage with LDS.128
template <bool INTERLEAVED>
__global__ void __launch_bounds__(256) k_tma(const int4* __restrict__ buf, long long n, int S, int CH, unsigned* sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    unsigned char* ring = smem + 128;
    const uint64_t pol = evict_first();
    const int tid = threadIdx.x;
    long long lo, hi, stride;
    if (INTERLEAVED) {
        lo = (long long)blockIdx.x * CH;
        hi = n;
        stride = (long long)gridDim.x * CH;
    } else {
        const long long per = ((n + gridDim.x - 1) / gridDim.x + 31) & ~31LL;
        lo = per * blockIdx.x < n ? per * blockIdx.x : n;
        hi = lo + per < n ? lo + per : n;
        stride = CH;
    }
    const int n_chunks = lo < hi ? (int)((hi - lo + stride - 1) / stride) : 0;
    auto issue = [&](int c, int slot) {
        long long b = lo + c * stride;
        long long cnt = hi - b < CH ? hi - b : CH;
        mbar_expect_tx(full + slot, (unsigned)cnt * 16u);
        tma_1d(ring + (size_t)slot * CH * 16, buf + b, (unsigned)cnt * 16u, full + slot, pol);
    };
    if (tid == 0) {
        for (int s = 0; s < S; ++s) mbar_init(full + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int c = 0; c < S && c < n_chunks; ++c) issue(c, c);
    }
    __syncthreads();
    unsigned acc = 0;
    int slot = 0;
    unsigned phase = 0;
    for (int c = 0; c < n_chunks; ++c) {
        mbar_wait(full + slot, phase);
        const int4* st = reinterpret_cast<const int4*>(ring + (size_t)slot * CH * 16);
        long long b = lo + c * stride;
        int cnt = (int)(hi - b < CH ? hi - b : CH);
        for (int i = tid; i < cnt; i += 256) {
            int4 v = st[i];
            acc ^= v.x ^ v.y ^ v.z ^ v.w;
        }
        __syncthreads();
        if (tid == 0 && c + S < n_chunks) issue(c + S, slot);
        if (++slot == S) {
            slot = 0;
            phase ^= 1u;
        }
    }
    if (acc == 0x12345678u) *sink = acc;
}

int main() {
    const long long n = 5000000;  // records per buffer (80 MB)
    const int nbuf = 24;
    std::vector<int4*> bufs(nbuf);
    for (int i = 0; i < nbuf; ++i) {
        CK(cudaMalloc(&bufs[i], n * 16));
        CK(cudaMemset(bufs[i], i + 1, n * 16));
    }
    unsigned* sink;
    CK(cudaMalloc(&sink, 4));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    auto report = [&](const char* name, float ms) {
        double us = ms * 1e3 / nbuf;
        printf("%-52s %7.2f us/buffer  %7.1f GB/s\n", name, us, n * 16 / us / 1e3);
    };
#define RUN(name, launch)                                  \
    do {                                                   \
        for (int i = 0; i < nbuf; ++i) { launch; }         \
        CK(cudaDeviceSynchronize());                       \
        cudaEventRecord(e0);                               \
        for (int i = 0; i < nbuf; ++i) { launch; }         \
        cudaEventRecord(e1);                               \
        CK(cudaDeviceSynchronize());                       \
        float ms;                                          \
        cudaEventElapsedTime(&ms, e0, e1);                 \
        report(name, ms);                                  \
    } while (0)

    char name[128];
    for (int occ : {4, 8}) {
        snprintf(name, sizeof name, "ldg U=4 spans        grid=%dx%d", sms, occ);
        RUN(name, (k_ldg<4, false><<<sms * occ, 256>>>(bufs[i], n, sink)));
        snprintf(name, sizeof name, "ldg U=4 interleaved  grid=%dx%d", sms, occ);
        RUN(name, (k_ldg<4, true><<<sms * occ, 256>>>(bufs[i], n, sink)));
        snprintf(name, sizeof name, "ldg U=8 spans        grid=%dx%d", sms, occ);
        RUN(name, (k_ldg<8, false><<<sms * occ, 256>>>(bufs[i], n, sink)));
        snprintf(name, sizeof name, "ldg U=8 interleaved  grid=%dx%d", sms, occ);
        RUN(name, (k_ldg<8, true><<<sms * occ, 256>>>(bufs[i], n, sink)));
    }
    CK(cudaFuncSetAttribute(k_tma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(k_tma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    struct Cfg { int occ, S, CH; };
    for (Cfg c : {Cfg{4, 2, 1024}, Cfg{3, 3, 1024}, Cfg{2, 6, 1024}, Cfg{1, 12, 1024}, Cfg{2, 3, 2048}, Cfg{1, 6, 2048}, Cfg{4, 3, 512},
                  Cfg{2, 12, 512}, Cfg{4, 6, 512}, Cfg{1, 3, 4096}}) {
        size_t smem = 128 + (size_t)c.S * c.CH * 16;
        snprintf(name, sizeof name, "tma spans       occ=%d stages=%2d chunk=%3dKB (%3zuKB/SM)", c.occ, c.S, c.CH * 16 / 1024, smem * c.occ / 1024);
        RUN(name, (k_tma<false><<<sms * c.occ, 256, smem>>>(bufs[i], n, c.S, c.CH, sink)));
        snprintf(name, sizeof name, "tma interleaved occ=%d stages=%2d chunk=%3dKB (%3zuKB/SM)", c.occ, c.S, c.CH * 16 / 1024, smem * c.occ / 1024);
        RUN(name, (k_tma<true><<<sms * c.occ, 256, smem>>>(bufs[i], n, c.S, c.CH, sink)));
    }
    return 0;
}
