// Scratch: which blocks of another kernel become resident next to a persistent Winograd-GEMM CTA?
// spin kernels of a given block size / register footprint record (SM id, start, end) in globaltimer ns;
// stamp kernels bracket the GEMM on its own stream.  Built and driven by scratch/corun/corun_probe.py.
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint64_t gtime() {
    uint64_t t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ uint32_t smid() {
    uint32_t s;
    asm volatile("mov.u32 %0, %smid;" : "=r"(s));
    return s;
}

extern "C" __global__ void stamp_kernel(uint64_t *out) { if (threadIdx.x == 0) *out = gtime(); }

// NR live floats per thread keep the register allocation at roughly NR + 20
template <int NR>
__global__ void spin_kernel(uint64_t *rec, float *sink, uint64_t ns, float seed) {
    float acc[NR];
#pragma unroll
    for (int i = 0; i < NR; ++i) acc[i] = seed + (float)i;
    const uint64_t t0 = gtime();
    uint64_t t = t0;
    while (t - t0 < ns) {
#pragma unroll
        for (int i = 0; i < NR; ++i) acc[i] = fmaf(acc[i], 1.0000001f, acc[(i + 1) % NR] * 1e-9f);
        t = gtime();
    }
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < NR; ++i) s += acc[i];
    if (s == 12345.678f) sink[threadIdx.x] = s;
    if (threadIdx.x == 0) {
        rec[3 * blockIdx.x] = smid();
        rec[3 * blockIdx.x + 1] = t0;
        rec[3 * blockIdx.x + 2] = t;
    }
}

extern "C" int launch_stamp(uint64_t *out, void *stream) {
    stamp_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(out);
    return (int)cudaGetLastError();
}

extern "C" int launch_spin(int nr, int blocks, int threads, uint64_t *rec, float *sink, uint64_t ns, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    // the GEMM runs with the maximum shared-memory carve-out: ask for the same split
    if (nr <= 16) { cudaFuncSetAttribute(spin_kernel<16>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared); spin_kernel<16><<<blocks, threads, 0, s>>>(rec, sink, ns, 1.0f); }
    else if (nr <= 64) { cudaFuncSetAttribute(spin_kernel<64>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared); spin_kernel<64><<<blocks, threads, 0, s>>>(rec, sink, ns, 1.0f); }
    else { cudaFuncSetAttribute(spin_kernel<128>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared); spin_kernel<128><<<blocks, threads, 0, s>>>(rec, sink, ns, 1.0f); }
    return (int)cudaGetLastError();
}

extern "C" int spin_regs(int nr) {
    cudaFuncAttributes a;
    if (nr <= 16) cudaFuncGetAttributes(&a, spin_kernel<16>);
    else if (nr <= 64) cudaFuncGetAttributes(&a, spin_kernel<64>);
    else cudaFuncGetAttributes(&a, spin_kernel<128>);
    return a.numRegs;
}
