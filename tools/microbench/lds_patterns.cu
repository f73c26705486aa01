// Microbenchmark: cost of a warp-wide shared-memory load on sm_100a as a function of access width and of how many
// distinct addresses the 32 lanes ask for (the gather of the run kernel: lanes = particles of 1..4 cells / window
// parities; the record reads of phase 2: lanes = nodes).  Prints SM cycles per warp instruction at saturation
// (16 warps per SM, as the run kernel has).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o lds_patterns lds_patterns.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

template<int W>
__device__ __forceinline__ float ld(unsigned a);
template<>
__device__ __forceinline__ float ld<1>(unsigned a)
{
    float v;
    asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
    return v;
}
template<>
__device__ __forceinline__ float ld<2>(unsigned a)
{
    float x, y;
    asm volatile("ld.volatile.shared.v2.f32 {%0,%1}, [%2];" : "=f"(x), "=f"(y) : "r"(a) : "memory");
    return x;
}
template<>
__device__ __forceinline__ float ld<4>(unsigned a)
{
    float x, y, z, w;
    asm volatile("ld.volatile.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x), "=f"(y), "=f"(z), "=f"(w) : "r"(a) : "memory");
    return x;
}

// word offset of lane's address for pattern p (all multiples of 4 words so that every width is aligned)
__device__ int laneOffset(int p, int lane)
{
    switch(p)
    {
    case 0: return 0; // uniform
    case 1: return (lane >> 4) * 4; // 2 distinct, neighbouring quads, split at lane 16
    case 2: return (lane >> 4) * 20; // 2 distinct, neighbouring rows (20 words)
    case 3: return (lane >> 3) * 4; // 4 distinct neighbouring quads, quarter-warp uniform
    case 4: return ((lane * 5) >> 5) * 4 + ((lane & 1) ? 20 : 0); // 5 groups x 2 interleaved parities (10 distinct)
    case 5: return (lane >= 11 ? 4 : 0) + (lane >= 25 ? 4 : 0); // 3 cells with boundaries at lanes 11, 25
    case 6: return ((lane >= 11) + (lane >= 25)) * 4 + ((lane * 7 % 3 == 0) ? 20 : 0); // the same with a scattered second row
    case 7: return (lane >> 2) * 4; // 8 distinct quads
    case 8: return lane * 4; // 32 distinct consecutive quads (512 B for W=4)
    case 9: return (lane & 15) * 4; // two half warps ask for the same 16 quads
    case 10: return (lane % 12) * 4; // phase-2 like: 12 distinct
    case 11: return (lane & 3) * 4; // 4 distinct, interleaved
    default: return lane * 36; // record stride (36 words): all distinct, as the record writes / EmZ
    }
}

template<int W>
__global__ void __launch_bounds__(256, 2) k(int pattern, int iters, float* out, long long* cyc)
{
    extern __shared__ float4 sm4[];
    float* sm = reinterpret_cast<float*>(sm4);
    for(int i = threadIdx.x; i < 12288; i += blockDim.x)
        sm[i] = float(i);
    __syncthreads();
    int const lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned const base = (unsigned) __cvta_generic_to_shared(sm + laneOffset(pattern, lane) + warp * 1280);
    float acc = 0.f;
    __syncthreads();
    long long const t0 = clock64();
#pragma unroll 1
    for(int i = 0; i < iters; ++i)
    {
#pragma unroll
        for(int u = 0; u < 16; ++u)
            acc += ld<W>(base + u * 16);
    }
    __syncthreads();
    long long const t1 = clock64();
    if(acc == 12345.678f)
        out[0] = acc;
    if(threadIdx.x == 0)
        cyc[blockIdx.x] = t1 - t0;
}

template<int W>
void run()
{
    float* out;
    long long* cyc;
    cudaMalloc(&out, 4);
    cudaMalloc(&cyc, 8 * 296);
    static const char* names[] = {"uniform", "2 addr (lane 16), same row", "2 addr, two rows", "4 addr quarter-uniform", "10 addr, 2 rows interleaved",
                                  "3 cells (11/14/7 lanes)", "3 cells + scattered 2nd row", "8 addr (lane/4)", "32 consecutive quads", "16 quads twice",
                                  "12 addr (lane%12)", "4 addr interleaved", "32 addr stride 36 words"};
    cudaFuncSetAttribute(k<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    for(int p = 0; p <= 12; ++p)
    {
        int const iters = 2000;
        k<W><<<296, 256, 100 * 1024>>>(p, iters, out, cyc);
        cudaDeviceSynchronize();
        long long h[296];
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double s = 0;
        for(int i = 0; i < 296; ++i)
            s += double(h[i]);
        s /= 296;
        // 16 warps per SM, each iters*16 loads
        printf("LDS.%-3d %-32s %6.2f SM cycles per warp instruction\n", W * 32, names[p], s / (16.0 * iters * 16));
    }
    cudaFree(out);
    cudaFree(cyc);
}

int main()
{
    run<1>();
    run<2>();
    run<4>();
    cudaError_t e = cudaGetLastError();
    if(e != cudaSuccess)
        printf("error %s\n", cudaGetErrorString(e));
    return 0;
}
