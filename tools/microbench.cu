// Integer-pipe microbenchmark for the Hamming roofline: POPC, LOP3 and the
// kernel's own CSA+POPC mix, per SM per clock.  nvcc -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(uint32_t* out, int iters, uint32_t seed)
{
    uint32_t a[8];
    for (int i = 0; i < 8; ++i) a[i] = seed * (threadIdx.x + 1) + i * 0x9e3779b9u;
    uint32_t acc = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (MODE == 0) {  // POPC only (dependent through add on a different pipe)
                acc += __popc(a[u] ^ acc);
            } else if (MODE == 1) {  // LOP3 only
                a[u] = (a[u] & a[(u + 1) & 7]) | (a[(u + 2) & 7] & (a[u] ^ a[(u + 1) & 7]));
            } else if (MODE == 2) {  // 4 LOP3 : 1 POPC  (the kernel's ratio 67:16)
                uint32_t x = a[u] ^ acc, y = a[(u + 1) & 7] ^ acc, z = a[(u + 2) & 7] ^ acc;
                uint32_t s = x ^ y ^ z, c = (x & y) | (z & (x ^ y));
                a[u] = s;
                acc += __popc(c);
            } else {  // independent popc, no xor: pure issue rate of POPC
                acc += __popc(a[u]);
                a[u] += 0x01010101u;
            }
        }
    }
    for (int i = 0; i < 8; ++i) acc ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
void run(const char* name, double ops_per_iter_thread)
{
    int dev_sms = 0, clk = 0;
    cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int blocks = dev_sms * 8, threads = 256, iters = 20000;
    uint32_t* out;
    cudaMalloc(&out, (size_t)blocks * threads * 4);
    k<MODE><<<blocks, threads>>>(out, 100, 1);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, iters, 3);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)blocks * threads * iters * ops_per_iter_thread;
    double per_s = ops / (ms * 1e-3);
    printf("%-28s %8.3f ms  %.3e ops/s  %.1f ops/clk/SM @%d MHz(max)\n", name, ms, per_s,
           per_s / dev_sms / (clk * 1e3), clk / 1000);
    cudaFree(out);
}

int main()
{
    run<0>("popc+xor+add (per popc)", 8);
    run<3>("popc+add independent", 8);
    run<1>("lop3-ish maj (per stmt)", 8);
    run<2>("csa mix (per popc)", 8);
    return 0;
}
