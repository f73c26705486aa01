#include <cute/arch/copy_sm90_tma.hpp>
#include <cutlass/arch/barrier.h>
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
constexpr int PX = 12, TY = 11, TZ = 7;
__global__ void k(const __grid_constant__ CUtensorMap map, float* out, int ox, int oy, int oz, int words)
{
    extern __shared__ __align__(128) float tile[];
    __shared__ uint64_t bar;
    using Bar = cutlass::arch::ClusterTransactionBarrier;
    if(threadIdx.x == 0)
    {
        Bar::init(&bar, 1);
        cutlass::arch::fence_barrier_init();
    }
    __syncthreads();
    if(threadIdx.x == 0)
    {
        Bar::arrive_and_expect_tx(&bar, words * 4);
        cute::SM90_TMA_LOAD_3D::copy(&map, &bar, 0ull, tile, ox, oy, oz);
    }
    Bar::wait(&bar, 0);
    for(int i = threadIdx.x; i < words; i += blockDim.x)
        out[i] = tile[i];
}
int main(int argc, char** argv)
{
    int ox = argc > 1 ? atoi(argv[1]) : 7;
    int N[3] = {32, 32, 16};
    long long vol = (long long) N[0] * N[1] * N[2];
    std::vector<float> h(vol);
    for(size_t i = 0; i < h.size(); ++i) h[i] = float(i);
    float *dE, *dout;
    int const words = PX * TY * TZ;
    cudaMalloc(&dE, vol * 4); cudaMalloc(&dout, words * 4);
    cudaMemcpy(dE, h.data(), vol * 4, cudaMemcpyHostToDevice);
    using Encode = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, cuuint64_t const*, cuuint64_t const*, cuuint32_t const*, cuuint32_t const*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr; cudaDriverEntryPointQueryResult qr;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr);
    Encode encode = (Encode) fn;
    alignas(64) CUtensorMap map;
    cuuint64_t dims[3] = {cuuint64_t(N[0]), cuuint64_t(N[1]), cuuint64_t(N[2])};
    cuuint64_t strides[2] = {cuuint64_t(N[0]) * 4, cuuint64_t(N[0]) * N[1] * 4};
    cuuint32_t bx[3] = {PX, TY, TZ}, es[3] = {1, 1, 1};
    CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, dE, dims, strides, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode -> %d  qr %d\n", int(r), int(qr));
    for(int i = 0; i < 16; ++i) printf("%016llx ", ((unsigned long long*) &map)[i]);
    printf("\n");
    int dev; cudaGetDevice(&dev); cudaDeviceProp pr; cudaGetDeviceProperties(&pr, dev); printf("device %s cc %d.%d\n", pr.name, pr.major, pr.minor);
    int drv, rt; cudaDriverGetVersion(&drv); cudaRuntimeGetVersion(&rt); printf("driver %d runtime %d\n", drv, rt);
    k<<<1, 128, words * 4>>>(map, dout, ox, 7, 3, words);
    cudaError_t e = cudaDeviceSynchronize();
    printf("cutlass kernel: %s\n", cudaGetErrorString(e));
    std::vector<float> o(words);
    cudaMemcpy(o.data(), dout, words * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for(int z = 0; z < TZ; ++z) for(int y = 0; y < TY; ++y) for(int x = 0; x < PX; ++x)
        if(o[(z * TY + y) * PX + x] != float(((3 + z) * N[1] + 7 + y) * N[0] + ox + x)) ++bad;
    printf("mismatches: %d\n", bad);
    return 0;
}
