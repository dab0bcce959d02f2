// gather_probe.cu — what does the per-event random table look-up cost on its own?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/gather_probe.bin tools/gather_probe.cu && tools/gather_probe.bin
//
// 4.5 M random 4-byte reads from a 1.2 MB table (the packed rectify LUT of the default geometry), indices
// derived from a coalesced 16-byte event stream (80 MB, like K1) or from a hash (no stream), at several
// occupancies, with plain loads and with cp.async (LDGSTS) into shared memory; plus 1.33 M 64-bit
// atomicMax into a 6 MB region (the scatter).  Results: profiles/gather_probe_r01.txt.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ unsigned hash32(unsigned x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

// mode 0: indices from a hash (no stream)   1: indices from the event stream (x | y << 16 in word 0)
// mode 2: stream only (no gather)
template <int MODE, int U>
__global__ void __launch_bounds__(256) k_gather(const int4* __restrict__ ev, long long n, const int* __restrict__ lut, int cells, unsigned* sink) {
    unsigned acc = 0;
    for (long long base = (long long)blockIdx.x * 256 * U; base < n; base += (long long)gridDim.x * 256 * U) {
        int idx[U];
#pragma unroll
        for (int k = 0; k < U; ++k) {
            long long i = base + k * 256 + threadIdx.x;
            if (MODE == 0) {
                idx[k] = hash32((unsigned)i) % (unsigned)cells;
            } else {
                int4 r = i < n ? __ldcs(ev + i) : make_int4(0, 0, 0, 0);
                idx[k] = ((unsigned)r.x >> 16) * 640 + (r.x & 0xffff);
                acc ^= r.y ^ r.z ^ r.w;
            }
        }
        if (MODE != 2) {
#pragma unroll
            for (int k = 0; k < U; ++k) acc ^= __ldg(lut + idx[k]);
        } else {
#pragma unroll
            for (int k = 0; k < U; ++k) acc ^= idx[k];
        }
    }
    if (acc == 0x12345678u) *sink = acc;
}

__global__ void __launch_bounds__(256) k_scatter(unsigned long long* map, int cells, long long n, unsigned epoch) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        unsigned c = hash32((unsigned)i * 7u + 1u) % (unsigned)cells;
        unsigned long long key = ((unsigned long long)epoch << 48) | ((unsigned long long)i << 16) | 17ull;
        asm volatile("red.global.max.u64 [%0], %1;" ::"l"(map + c), "l"(key) : "memory");
    }
}

// the whole per-event path as ONE simple kernel: no shared memory, no TMA, no software pipeline -- U events per
// thread, latency hidden by occupancy alone (timing model only: ties / bounds violations are not handled)
template <int U>
__global__ void __launch_bounds__(256) k_simple(const int4* __restrict__ ev, long long n, const int* __restrict__ lut,
                                                const short* __restrict__ xmap_t, int col_stride, unsigned long long* map, int rect_w,
                                                unsigned t_lo, unsigned range, unsigned epoch) {
    const unsigned scale2 = 2u * 719u, d = 2u * range;
    const unsigned M = (unsigned)(((1ULL << (32 + (31 - __clz(d)))) + d - 1u) / d);
    const int sh = 31 - __clz(d);
    for (long long base = (long long)blockIdx.x * 256 * U; base < n; base += (long long)gridDim.x * 256 * U) {
        int4 r[U];
#pragma unroll
        for (int k = 0; k < U; ++k) {
            long long i = base + k * 256 + threadIdx.x;
            r[k] = i < n ? __ldcs(ev + i) : make_int4(0, 0, 0, 0);
        }
        int l[U];
#pragma unroll
        for (int k = 0; k < U; ++k) {
            const unsigned px = ((unsigned)r[k].x >> 16) * 640u + ((unsigned)r[k].x & 0xffffu);
            l[k] = (r[k].y == 1 && px < 307200u) ? __ldg(lut + px) : 0x7fff0000;
        }
#pragma unroll
        for (int k = 0; k < U; ++k) {
            const unsigned dt = (unsigned)r[k].z - t_lo;
            const unsigned q = __umulhi(dt * scale2 + range, M) >> sh;
            const int ycr = l[k] >> 16, xcr = (short)(l[k] & 0xffff);
            if ((unsigned)ycr < 1319u && q < 720u) {
                const int xp = __ldg(xmap_t + (long long)q * col_stride + ycr);
                const int disp = (short)(xp - xcr - 4242);
                if (disp >= 0) {
                    const unsigned idx = (unsigned)(base + k * 256 + threadIdx.x);
                    const unsigned long long key = ((unsigned long long)((epoch << 16) | (idx >> 16)) << 32) | ((idx << 16) | (unsigned)disp);
                    asm volatile("red.global.max.u64 [%0], %1;" ::"l"(map + (long long)ycr * rect_w + (xp - 4242)), "l"(key) : "memory");
                }
            }
        }
    }
}

template <typename F>
float time_us(F f, int reps = 20) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f();
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int i = 0; i < reps; ++i) f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms * 1000.f / reps;
}

int main() {
    const long long n = 5000000;
    const int cells = 640 * 480;
    const int frames = 8;
    std::vector<int4> h(n);
    unsigned s = 1;
    for (long long i = 0; i < n; ++i) {
        s = s * 1664525u + 1013904223u; unsigned x = (s >> 8) % 640;
        s = s * 1664525u + 1013904223u; unsigned y = (s >> 8) % 480;
        h[i] = make_int4((int)(x | (y << 16)), 1, (int)i, 0);
    }
    int4* d_ev; int* d_lut; unsigned* d_sink; unsigned long long* d_map;
    CK(cudaMalloc(&d_ev, n * 16 * frames));
    for (int f = 0; f < frames; ++f) CK(cudaMemcpy(d_ev + f * n, h.data(), n * 16, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_lut, cells * 4)); CK(cudaMemset(d_lut, 1, cells * 4));
    CK(cudaMalloc(&d_sink, 4));
    const int map_cells = 765000;
    CK(cudaMalloc(&d_map, (size_t)map_cells * 8)); CK(cudaMemset(d_map, 0, (size_t)map_cells * 8));
    int fi = 0;
    for (int occ : {2, 4, 8}) {
        int grid = 148 * occ;
        float t0 = time_us([&] { k_gather<0, 4><<<grid, 256>>>(d_ev, n * 9 / 10, d_lut, cells, d_sink); });
        float t1 = time_us([&] { k_gather<1, 4><<<grid, 256>>>(d_ev + (fi++ % frames) * n, n, d_lut, cells, d_sink); });
        float t2 = time_us([&] { k_gather<2, 4><<<grid, 256>>>(d_ev + (fi++ % frames) * n, n, d_lut, cells, d_sink); });
        float t3 = time_us([&] { k_gather<1, 8><<<grid, 256>>>(d_ev + (fi++ % frames) * n, n, d_lut, cells, d_sink); });
        printf("CTAs/SM=%d  gather only (4.5M, hash idx) %6.2f us | stream+gather U=4 %6.2f us | stream only %6.2f us | stream+gather U=8 %6.2f us\n", occ, t0, t1, t2, t3);
    }
    unsigned epoch = 1;
    float ts = time_us([&] { k_scatter<<<148 * 8, 256>>>(d_map, map_cells, 1330000, epoch++); });
    printf("scatter only: 1.33M red.max.u64 into %d cells: %6.2f us\n", map_cells, ts);
    // full simple kernel on realistic tables: needs the repo's golden tables dumped as raw files (optional)
    {
        FILE* f1 = fopen("gpurun_out/lut_xy.bin", "rb");
        FILE* f2 = fopen("gpurun_out/xmap_t.bin", "rb");
        if (f1 && f2) {
            std::vector<int> hl(cells);
            const int col_stride = 1320;
            std::vector<short> hx((size_t)720 * col_stride);
            size_t a = fread(hl.data(), 4, hl.size(), f1), b = fread(hx.data(), 2, hx.size(), f2);
            if (a == hl.size() && b == hx.size()) {
                short* d_x; unsigned long long* d_big;
                CK(cudaMemcpy(d_lut, hl.data(), cells * 4, cudaMemcpyHostToDevice));
                CK(cudaMalloc(&d_x, hx.size() * 2)); CK(cudaMemcpy(d_x, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice));
                CK(cudaMalloc(&d_big, (size_t)1760 * 1320 * 8)); CK(cudaMemset(d_big, 0, (size_t)1760 * 1320 * 8));
                // timestamps: word z = event index scaled to 0..16665
                for (long long i = 0; i < n; ++i) h[i].z = (int)(i * 16666 / n);
                for (int f = 0; f < frames; ++f) CK(cudaMemcpy(d_ev + f * n, h.data(), n * 16, cudaMemcpyHostToDevice));
                for (int occ : {3, 4, 6, 8}) {
                    float t4 = time_us([&] { k_simple<4><<<148 * occ, 256>>>(d_ev + (fi++ % frames) * n, n, d_lut, d_x, col_stride, d_big, 1760, 0u, 16665u, epoch++); });
                    float t2 = time_us([&] { k_simple<2><<<148 * occ, 256>>>(d_ev + (fi++ % frames) * n, n, d_lut, d_x, col_stride, d_big, 1760, 0u, 16665u, epoch++); });
                    float t8 = time_us([&] { k_simple<8><<<148 * occ, 256>>>(d_ev + (fi++ % frames) * n, n, d_lut, d_x, col_stride, d_big, 1760, 0u, 16665u, epoch++); });
                    printf("simple full path, CTAs/SM=%d: U=2 %6.2f us  U=4 %6.2f us  U=8 %6.2f us\n", occ, t2, t4, t8);
                }
            }
        }
        if (f1) fclose(f1);
        if (f2) fclose(f2);
    }
    CK(cudaDeviceSynchronize());
    return 0;
}
