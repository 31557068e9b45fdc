// Peak-rate microbenchmark for tcgen05.mma on this GPU: kind::i8 vs kind::f8f6f4 vs kind::f16, operands
// from (uninitialised) shared memory, M=128, N in {128,256}, cta_group::1, one CTA per SM issuing
// back-to-back MMAs into TMEM.  Gives the tensor-pipe roofline denominator for the int8 Hamming scan.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smemAddr(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint64_t makeDesc(uint32_t saddr)
{
    uint64_t d = 0;
    d |= uint64_t((saddr & 0x3FFFF) >> 4);
    d |= uint64_t(1) << 16;
    d |= uint64_t(1024 >> 4) << 32;
    d |= uint64_t(1) << 46;
    d |= uint64_t(2) << 61;
    return d;
}

template <int KIND>   // 0 = i8, 1 = f8f6f4 (e4m3), 2 = f16 (bf16)
__global__ void __launch_bounds__(128, 1) peak(uint32_t idesc, int iters, int nTile, int useTmemA, uint32_t* out)
{
    extern __shared__ uint8_t raw[];
    uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0x01010101u * (i & 1);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smemAddr(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smemAddr(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    if (threadIdx.x == 0) {
        const uint32_t aAddr = smemAddr(sm), bAddr = smemAddr(sm + 32 * 1024);
        for (int it = 0; it < iters; it++) {
            const uint32_t d = tmem + (it & 1) * nTile * 0;   // same accumulator: back-to-back dependent accumulate
            const uint64_t da = makeDesc(aAddr + (it & 3) * 32), db = makeDesc(bAddr + (it & 3) * 32);
            if (KIND == 0) {
                if (useTmemA)
                    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n}" ::"r"(d), "r"(tmem + 256 + (it & 3) * 8), "l"(db), "r"(idesc), "r"(it));
                else
                    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(it));
            } else if (KIND == 1) {
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(it));
            } else {
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(it));
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smemAddr(&bar)) : "memory");
        asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}" ::"r"(smemAddr(&bar)) : "memory");
        out[blockIdx.x] = 1;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

template <int KIND> double run(uint32_t idesc, int nTile, int kPer, int useTmemA, int sms, uint32_t* out)
{
    const int iters = 20000;
    const size_t smem = 100 * 1024;
    cudaFuncSetAttribute(peak<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    peak<KIND><<<sms, 128, smem>>>(idesc, 1000, nTile, useTmemA, out);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(a);
        peak<KIND><<<sms, 128, smem>>>(idesc, iters, nTile, useTmemA, out);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("\"error\": \"%s\", ", cudaGetErrorString(e)); return 0; }
    return 2.0 * 128 * nTile * kPer * double(iters) * sms / (best * 1e-3) / 1e12;
}

// Back-to-back launches for `seconds`; the rate over the second half (the GPU has reached its power-capped clocks by then).
template <int KIND> double sustained(uint32_t idesc, int nTile, int kPer, int useTmemA, int sms, uint32_t* out, double seconds)
{
    const int iters = 20000;
    const size_t smem = 100 * 1024;
    cudaFuncSetAttribute(peak<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    const double perLaunch = 2.0 * 128 * nTile * kPer * double(iters) * sms;
    // one launch is ~1 ms: warm for the first half, time the second half
    const int launches = int(seconds * 1000.0);
    for (int i = 0; i < launches / 2; i++) peak<KIND><<<sms, 128, smem>>>(idesc, iters, nTile, useTmemA, out);
    cudaEventRecord(a);
    for (int i = 0; i < launches / 2; i++) peak<KIND><<<sms, 128, smem>>>(idesc, iters, nTile, useTmemA, out);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    if (cudaGetLastError() != cudaSuccess) return 0;
    return perLaunch * (launches / 2) / (ms * 1e-3) / 1e12;
}

int main(int argc, char** argv)
{
    cudaDeviceProp p;
    const bool quick = argc > 1 && argv[1][0] == 'q';      // "quick": burst figures only
    cudaGetDeviceProperties(&p, 0);
    uint32_t* out;
    cudaMalloc(&out, 4096);
    const int sms = p.multiProcessorCount;
    auto idescI8 = [](int n) { return (2u << 4) | (1u << 7) | (1u << 10) | (uint32_t(n >> 3) << 17) | (8u << 24); };
    auto idescF8 = [](int n) { return (1u << 4) | (0u << 7) | (0u << 10) | (uint32_t(n >> 3) << 17) | (8u << 24); };   // e4m3 x e4m3 -> f32
    auto idescBf = [](int n) { return (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(n >> 3) << 17) | (8u << 24); };   // bf16 x bf16 -> f32
    printf("{\"gpu\": \"%s\", ", p.name);
    printf("\"i8_ss_n256_tops\": %.1f, ", run<0>(idescI8(256), 256, 32, 0, sms, out));
    printf("\"i8_ss_n128_tops\": %.1f, ", run<0>(idescI8(128), 128, 32, 0, sms, out));
    printf("\"i8_ts_n128_tops\": %.1f, ", run<0>(idescI8(128), 128, 32, 1, sms, out));
    printf("\"i8_ts_n256_tops\": %.1f, ", run<0>(idescI8(256), 256, 32, 1, sms, out));
    printf("\"f8_ss_n256_tflops\": %.1f, ", run<1>(idescF8(256), 256, 32, 0, sms, out));
    printf("\"bf16_ss_n256_tflops\": %.1f", run<2>(idescBf(256), 256, 16, 0, sms, out));
    if (!quick) {
        printf(", \"i8_ss_n256_tops_sustained\": %.1f", sustained<0>(idescI8(256), 256, 32, 0, sms, out, 3.0));
        printf(", \"i8_ts_n256_tops_sustained\": %.1f", sustained<0>(idescI8(256), 256, 32, 1, sms, out, 3.0));
        printf(", \"bf16_ss_n256_tflops_sustained\": %.1f", sustained<2>(idescBf(256), 256, 16, 0, sms, out, 3.0));
    }
    printf("}\n");
    return 0;
}
