// Pipe-rate microbenchmarks used as roofline denominators for the non-tensor kernels
// (POPC / LOP3 / IADD3 integer pipes for the XOR-POPC scan, DMUL+DADD for the signature kernel).
// Prints one JSON object.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define ILP 8

__global__ void kPopc(uint32_t* out, uint32_t seed)
{
    uint32_t x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = seed + threadIdx.x * 977 + i * 131071;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) x[i] = __popc(x[i]) + 0x9e3779b9u * 0 + (x[i] << 7);   // popc + shift-add
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void kPopcOnly(uint32_t* out, uint32_t seed)
{
    uint32_t x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = seed + threadIdx.x * 977 + i * 131071;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) asm volatile("popc.b32 %0, %0;" : "+r"(x[i]));
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void kLop3(uint32_t* out, uint32_t seed)
{
    uint32_t x[ILP];
    uint32_t a = seed * 3 + threadIdx.x, b = seed * 7 + blockIdx.x;
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = seed + threadIdx.x * 977 + i * 131071;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(a), "r"(b));
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void kIadd(uint32_t* out, uint32_t seed)
{
    uint32_t x[ILP];
    uint32_t a = seed * 3 + threadIdx.x;
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = seed + threadIdx.x * 977 + i * 131071;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(a));
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// the scan's inner mix: 3 XOR + full adder (2 LOP3) + 2 POPC + 2 adds per 3 words
__global__ void kCsaMix(uint32_t* out, uint32_t seed)
{
    uint32_t x[6], ones = 0, twos = 0;
    uint32_t a = seed * 3 + threadIdx.x, b = seed * 7 + blockIdx.x, c = seed ^ 0x55aa;
#pragma unroll
    for (int i = 0; i < 6; i++) x[i] = seed + threadIdx.x * 977 + i * 131071;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int g = 0; g < 2; g++) {
            uint32_t p = x[3 * g] ^ a, q = x[3 * g + 1] ^ b, r = x[3 * g + 2] ^ c, s, m;
            asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(s) : "r"(p), "r"(q), "r"(r));
            asm volatile("lop3.b32 %0, %1, %2, %3, 0xe8;" : "=r"(m) : "r"(p), "r"(q), "r"(r));
            ones += __popc(s);
            twos += __popc(m);
            x[3 * g] += ones;
            x[3 * g + 1] ^= twos;
        }
        a += it;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ones + 2 * twos + x[0] + x[4];
}

__global__ void kDmulDadd(double* out, double seed)
{
    double x[ILP];
    const double m = 1e-9 * seed;
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = seed + threadIdx.x + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) x[i] = __dadd_rn(x[i], __dmul_rn(m, x[i]));   // DMUL + DADD, like the signature kernel
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void kDfma(double* out, double seed)
{
    double x[ILP];
    const double m = 1.0000001 + seed * 1e-9, u = 1e-9 * threadIdx.x;
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = seed + threadIdx.x + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) x[i] = __fma_rn(x[i], m, u);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F> float timeIt(F f)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    f();
    f();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(a);
        f();
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    return best;
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int blocks = p.multiProcessorCount * 8, threads = 256;
    void* buf;
    cudaMalloc(&buf, size_t(blocks) * threads * 8);
    const double lanes = double(blocks) * threads * ITERS * ILP;
    float t;
    printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, p.multiProcessorCount);
    t = timeIt([&] { kPopcOnly<<<blocks, threads>>>((uint32_t*)buf, 1); });
    printf(", \"popc_per_s\": %.4e", lanes / (t * 1e-3));
    t = timeIt([&] { kPopc<<<blocks, threads>>>((uint32_t*)buf, 1); });
    printf(", \"popc_plus_lea_per_s\": %.4e", lanes / (t * 1e-3));
    t = timeIt([&] { kLop3<<<blocks, threads>>>((uint32_t*)buf, 1); });
    printf(", \"lop3_per_s\": %.4e", lanes / (t * 1e-3));
    t = timeIt([&] { kIadd<<<blocks, threads>>>((uint32_t*)buf, 1); });
    printf(", \"iadd_per_s\": %.4e", lanes / (t * 1e-3));
    t = timeIt([&] { kCsaMix<<<blocks, threads>>>((uint32_t*)buf, 1); });
    printf(", \"csa_words_per_s\": %.4e", double(blocks) * threads * ITERS * 6 / (t * 1e-3));
    t = timeIt([&] { kDmulDadd<<<blocks, threads>>>((double*)buf, 1.0); });
    printf(", \"dmul_dadd_mac_per_s\": %.4e", lanes / (t * 1e-3));
    t = timeIt([&] { kDfma<<<blocks, threads>>>((double*)buf, 1.0); });
    printf(", \"dfma_per_s\": %.4e", lanes / (t * 1e-3));
    printf("}\n");
    return 0;
}
