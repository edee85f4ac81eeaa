// gather_gran.cu -- how many DRAM bytes does a 32-byte gather cost on B200?  Each thread reads one
// 32-byte sector of a distinct pseudo-random 128-byte line of a 4 GiB array.  Run under
//   ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum ./gather_gran
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33; return x; }

template <int MODE>
__global__ void gather(const uint8_t* __restrict__ base, uint64_t nlines, uint32_t* __restrict__ out, uint64_t n)
{
    __shared__ alignas(128) uint8_t sm[256 * 32];
    __shared__ alignas(8) uint64_t bar;
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t line = (i * 0x9E3779B97F4A7C15ull >> 20) % nlines;          // distinct-ish lines
    const uint8_t* p = base + line * 128 + 32 * (mix(i) & 3);
    uint32_t acc = 0;
    if (MODE == 0)
    {
        const uint4 a = *reinterpret_cast<const uint4*>(p), b = *reinterpret_cast<const uint4*>(p + 16);
        acc = a.x ^ a.y ^ a.z ^ a.w ^ b.x ^ b.y ^ b.z ^ b.w;
    }
    else if (MODE == 1)
    {
        uint4 a, b;
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "l"(p));
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p + 16));
        acc = a.x ^ a.y ^ a.z ^ a.w ^ b.x ^ b.y ^ b.z ^ b.w;
    }
    else if (MODE == 2)
    {
        const uint32_t s = (uint32_t)__cvta_generic_to_shared(sm + threadIdx.x * 32);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(p) : "memory");
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s + 16), "l"(p + 16) : "memory");
        asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
        const uint4 a = *reinterpret_cast<const uint4*>(sm + threadIdx.x * 32);
        acc = a.x ^ a.y ^ a.z ^ a.w;
    }
    else if (MODE == 3)
    {   // 1-D bulk copy (TMA engine), one per thread, all on one mbarrier
        const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
        if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(1)); }
        __syncthreads();
        if (threadIdx.x == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(32 * blockDim.x) : "memory");
        __syncthreads();
        const uint32_t s = (uint32_t)__cvta_generic_to_shared(sm + threadIdx.x * 32);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 32, [%2];" ::"r"(s), "l"(p), "r"(b) : "memory");
        uint32_t done = 0;
        while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(b) : "memory");
        const uint4 a = *reinterpret_cast<const uint4*>(sm + threadIdx.x * 32);
        acc = a.x ^ a.y ^ a.z ^ a.w;
    }
    else if (MODE == 4)
    {   // evict-first policy on the load
        uint64_t pol;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
        uint4 a, b;
        asm volatile("ld.global.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "l"(p), "l"(pol));
        asm volatile("ld.global.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p + 16), "l"(pol));
        acc = a.x ^ a.y ^ a.z ^ a.w ^ b.x ^ b.y ^ b.z ^ b.w;
    }
    else if (MODE == 5)
    {   // 16 bytes only
        const uint4 a = *reinterpret_cast<const uint4*>(p);
        acc = a.x ^ a.y ^ a.z ^ a.w;
    }
    else if (MODE == 6)
    {   // 64 bytes: sectors 0-1 or 2-3 of the line
        const uint8_t* q = base + line * 128 + 64 * (mix(i) & 1);
        const uint4 a = *reinterpret_cast<const uint4*>(q), b = *reinterpret_cast<const uint4*>(q + 16), c = *reinterpret_cast<const uint4*>(q + 32), d = *reinterpret_cast<const uint4*>(q + 48);
        acc = a.x ^ b.y ^ c.z ^ d.w;
    }
    if (acc == 0x12345678u) out[i & 1023] = acc;
}

template <int MODE> void run(const uint8_t* d, uint64_t nlines, uint32_t* out, uint64_t n, const char* name)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    gather<MODE><<<(unsigned)((n + 255) / 256), 256>>>(d, nlines, out, n);
    cudaEventRecord(e0);
    gather<MODE><<<(unsigned)((n + 255) / 256), 256>>>(d, nlines, out, n);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-28s %8.3f ms  %7.1f GB/s useful (%s)\n", name, ms, n * 32.0 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    const uint64_t bytes = 4ull << 30, nlines = bytes / 128, n = 16ull << 20;     // 16 M gathers of 32 B = 512 MB useful
    uint8_t* d; uint32_t* out;
    cudaMalloc(&d, bytes); cudaMalloc(&out, 4096);
    cudaMemset(d, 1, bytes);
    run<0>(d, nlines, out, n, "ld.v4 x2 (32 B)");
    run<1>(d, nlines, out, n, "ld.nc.no_allocate x2");
    run<2>(d, nlines, out, n, "cp.async.cg 16 x2");
    run<3>(d, nlines, out, n, "cp.async.bulk 32");
    run<4>(d, nlines, out, n, "ld evict_first x2");
    run<5>(d, nlines, out, n, "ld.v4 x1 (16 B)");
    run<6>(d, nlines, out, n, "ld.v4 x4 (64 B)");
    return 0;
}
