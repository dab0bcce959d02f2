// Throughput of returning atomicAdd on ONE address (the batch kernel's ticket counters) from many SMs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/atomic_probe.bin tools/atomic_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
// every thread: `iters` atomics, `ilp` of them in flight; `stride_words` apart for thread groups (0 = all on one word)
__global__ void probe(unsigned* ctr, int iters, int ilp, int n_addr, int addr_stride_words, unsigned* sink) {
    unsigned* p = ctr + (blockIdx.x % n_addr) * addr_stride_words;
    unsigned acc = 0;
    for (int i = 0; i < iters; i += ilp) {
        unsigned v[4];
        for (int j = 0; j < ilp; ++j) v[j] = atomicAdd(p, 1u);
        for (int j = 0; j < ilp; ++j) acc += v[j];
    }
    if (acc == 0xdeadbeef) *sink = acc;
}
__global__ void probe_red(unsigned* ctr, int iters) {
    for (int i = 0; i < iters; ++i) atomicAdd(ctr, 1u);
}
int main() {
    unsigned* d;
    cudaMalloc(&d, 1 << 20);
    cudaMemset(d, 0, 1 << 20);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    auto run = [&](const char* name, int ctas, int thr, int iters, int ilp, int n_addr, int stride) {
        probe<<<ctas, thr>>>(d, 16, 1, 1, 0, d + 1000);
        cudaEventRecord(a);
        probe<<<ctas, thr>>>(d, iters, ilp, n_addr, stride, d + 1000);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        const double n = double(ctas) * thr * iters;
        printf("%-44s ctas %4d thr %3d ilp %d addrs %d stride %3d: %7.2f ns per atomic, %7.2f us round trip per thread\n", name, ctas, thr, ilp, n_addr, stride,
               ms * 1e6 / n, ms * 1e3 / (iters / ilp));
    };
    run("one word, 1 thread per CTA", 148, 1, 2000, 1, 1, 0);
    run("one word, 1 thread per CTA", 296, 1, 2000, 1, 1, 0);
    run("one word, 1 thread per CTA, 2 in flight", 296, 1, 2000, 2, 1, 0);
    run("one word, 4 warps per CTA (lane 0 each)", 296, 128, 500, 1, 1, 0);  // (all 128 threads issue: warp-aggregated by HW?)
    run("one word, 592 CTAs", 592, 1, 2000, 1, 1, 0);
    run("one word, 1184 CTAs", 1184, 1, 1000, 1, 1, 0);
    run("two words, same sector", 296, 1, 2000, 1, 2, 1);
    run("two words, same 128 B line", 296, 1, 2000, 1, 2, 16);
    run("two words, different lines", 296, 1, 2000, 1, 2, 64);
    run("eight words, different lines", 296, 1, 2000, 1, 8, 64);
    run("eight words, 1184 CTAs", 1184, 1, 1000, 1, 8, 64);
    cudaEventRecord(a);
    probe_red<<<296, 1>>>(d, 2000);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    printf("RED (no return), one word, 296 threads: %7.2f ns per atomic\n", ms * 1e6 / (296.0 * 2000));
    return 0;
}
