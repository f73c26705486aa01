// Microbenchmark: shared-memory load throughput per SM for broadcast-style address patterns (sm_100a).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o lds_bcast lds_bcast.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

template<int W> struct Vec;
template<> struct Vec<1> { using T = float; };
template<> struct Vec<2> { using T = float2; };
template<> struct Vec<4> { using T = float4; };

__device__ __forceinline__ float sum(float v) { return v; }
__device__ __forceinline__ float sum(float2 v) { return v.x + v.y; }
__device__ __forceinline__ float sum(float4 v) { return v.x + v.y + v.z + v.w; }

__device__ __forceinline__ float ld(float const* p)
{
    float v; unsigned a = (unsigned) __cvta_generic_to_shared(p);
    asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ float ld(float2 const* p)
{
    float2 v; unsigned a = (unsigned) __cvta_generic_to_shared(p);
    asm volatile("ld.volatile.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a) : "memory");
    return v.x + v.y;
}
__device__ __forceinline__ float ld(float4 const* p)
{
    float4 v; unsigned a = (unsigned) __cvta_generic_to_shared(p);
    asm volatile("ld.volatile.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
    return v.x + v.y + v.z + v.w;
}
// pattern: 0 all-same, 1 two halves, 2 lane/4 (8 distinct consecutive), 3 all distinct consecutive,
//          4 quarter-uniform (4 distinct), 5 lane/2 (16 distinct), 6: 8 distinct with record stride 60 words
template<int W>
__global__ void k(int pattern, int iters, float* out, long long* cyc)
{
    extern __shared__ float4 sm4[];
    float* sm = reinterpret_cast<float*>(sm4);
    for(int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = float(i);
    __syncthreads();
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int idx;
    switch(pattern)
    {
    case 0: idx = 0; break;
    case 1: idx = (lane >> 4); break;
    case 2: idx = (lane >> 2); break;
    case 3: idx = lane; break;
    case 4: idx = (lane >> 3); break;
    case 5: idx = (lane >> 1); break;
    default: idx = (lane >> 2) * (60 / W); break;
    }
    using T = typename Vec<W>::T;
    T const* base = reinterpret_cast<T const*>(sm) + idx + warp * 4;
    float acc = 0.f;
    long long t0 = clock64();
#pragma unroll 1
    for(int i = 0; i < iters; ++i)
    {
#pragma unroll
        for(int u = 0; u < 16; ++u)
        {
            acc += ld(base + u * (16 / W) * 4);
        }
    }
    long long t1 = clock64();
    if(acc == 12345.678f) out[0] = acc;
    if(threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template<int W>
void run(int warps)
{
    float* out; long long* cyc;
    cudaMalloc(&out, 4); cudaMalloc(&cyc, 8 * 148);
    for(int p = 0; p <= 6; ++p)
    {
        int iters = 2000;
        cudaMemset(cyc, 0, 8 * 148);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        k<W><<<148, warps * 32, 65536>>>(p, iters, out, cyc);
        cudaEventRecord(e1);
        cudaError_t err = cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if(err != cudaSuccess) printf("ERR %s\n", cudaGetErrorString(err));
        long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double c = double(h[0]) / (double(iters) * 16 * warps);
        printf("LDS.%d warps=%d pattern=%d : %.2f cyc per warp-instruction per SM (raw %lld cyc, %.3f ms)\n", 32 * W, warps, p, c, h[0], ms);
    }
}

int main()
{
    cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(k<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    for(int w : {8, 16})
    {
        run<1>(w); run<2>(w); run<4>(w);
    }
    return 0;
}
