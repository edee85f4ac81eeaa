// scatter_gran.cu -- DRAM cost of partial-line writes on B200: each thread writes BYTES bytes of a
// distinct pseudo-random 128-byte line of a 4 GiB array.  Run under
//   ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum ./scatter_gran
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int BYTES>
__global__ void scatter(uint8_t* __restrict__ base, uint64_t nlines, uint64_t n)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t line = (i * 0x9E3779B97F4A7C15ull >> 20) % nlines;
    uint8_t* p = base + line * 128;
    const uint4 v = make_uint4((uint32_t)i, 1, 2, 3);
    if (BYTES == 4) *reinterpret_cast<uint32_t*>(p + 4 * (i & 31)) = (uint32_t)i;
    else
#pragma unroll
        for (int k = 0; k < BYTES / 16; ++k) reinterpret_cast<uint4*>(p + (BYTES == 128 ? 0 : BYTES == 96 ? (i & 1) * 32 : BYTES == 64 ? (i & 1) * 64 : (i & 3) * 32))[k] = v;
}

template <int BYTES> void run(uint8_t* d, uint64_t nlines, uint64_t n)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    scatter<BYTES><<<(unsigned)((n + 255) / 256), 256>>>(d, nlines, n);
    cudaEventRecord(e0);
    scatter<BYTES><<<(unsigned)((n + 255) / 256), 256>>>(d, nlines, n);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("write %3d B per line: %8.3f ms  %7.1f GB/s useful (%s)\n", BYTES, ms, n * (double)BYTES / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    const uint64_t bytes = 4ull << 30, nlines = bytes / 128, n = 16ull << 20;
    uint8_t* d;
    cudaMalloc(&d, bytes);
    cudaMemset(d, 1, bytes);
    run<4>(d, nlines, n);
    run<16>(d, nlines, n);
    run<32>(d, nlines, n);
    run<64>(d, nlines, n);
    run<96>(d, nlines, n);
    run<128>(d, nlines, n);
    return 0;
}
