// stand-alone check of the TMA tile load helper (tma.cuh): nvcc -gencode arch=compute_100a,code=sm_100a -o tma_test tma_test.cu
#include "../../picongpu_b200/csrc/tma.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>
using namespace picstep;
constexpr int TY = 11, TZ = 7;
__device__ __forceinline__ void tmaLoad3(void* dst, CUtensorMap const* map, int x, int y, int z, uint64_t* bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smemAddr(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(z), "r"(smemAddr(bar)) : "memory");
}
__global__ void k(const __grid_constant__ TileMaps maps, CUtensorMap const* gm, float* out, int ox, int oy, int oz, int px, int rank, int words)
{
    extern __shared__ __align__(128) float tile[];
    __shared__ uint64_t bar;
    if(threadIdx.x == 0)
    {
        mbarInit(&bar, 1);
        int const ncomp = rank == 4 ? 3 : 1;
        mbarExpectTx(&bar, ncomp * px * TY * TZ * 4);
        CUtensorMap const* m = gm ? gm : &maps.E;
        if(rank == 4)
            tmaLoadTile(tile, m, ox, oy, oz, &bar);
        else
            tmaLoad3(tile, m, ox, oy, oz, &bar);
    }
    __syncthreads();
    mbarWait(&bar, 0);
    for(int i = threadIdx.x; i < words; i += blockDim.x)
        out[i] = tile[i];
}
int main(int argc, char** argv)
{
    int mode = argc > 1 ? atoi(argv[1]) : 0;
    int N[3] = {32, 32, 16};
    long long vol = (long long) N[0] * N[1] * N[2];
    std::vector<float> h(3 * vol);
    for(size_t i = 0; i < h.size(); ++i) h[i] = float(i);
    float *dE, *dout;
    int const px = mode == 4 ? 16 : 12;
    int const rank = mode == 3 ? 3 : 4;
    int const words = (rank == 4 ? 3 : 1) * px * TY * TZ;
    cudaMalloc(&dE, 3 * vol * 4); cudaMalloc(&dout, words * 4);
    cudaMemcpy(dE, h.data(), 3 * vol * 4, cudaMemcpyHostToDevice);
    using Encode = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, cuuint64_t const*, cuuint64_t const*, cuuint32_t const*, cuuint32_t const*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr; cudaDriverEntryPointQueryResult qr;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr);
    Encode encode = (Encode) fn;
    TileMaps maps;
    cuuint64_t dims[4] = {cuuint64_t(N[0]), cuuint64_t(N[1]), cuuint64_t(N[2]), 3};
    cuuint64_t strides[3] = {cuuint64_t(N[0]) * 4, cuuint64_t(N[0]) * N[1] * 4, cuuint64_t(vol) * 4};
    cuuint32_t bx[4] = {cuuint32_t(px), TY, TZ, 3}, es[4] = {1, 1, 1, 1};
    CUtensorMapL2promotion l2 = mode == 1 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    CUresult r = encode(&maps.E, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, dE, dims, strides, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("mode %d encode -> %d\n", mode, int(r));
    CUtensorMap* gm = nullptr;
    if(mode == 2)
    {
        cudaMalloc(&gm, sizeof(CUtensorMap));
        cudaMemcpy(gm, &maps.E, sizeof(CUtensorMap), cudaMemcpyHostToDevice);
    }
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, words * 4);
    k<<<1, 256, words * 4>>>(maps, gm, dout, 7, 7, 3, px, rank, words);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    std::vector<float> o(words);
    cudaMemcpy(o.data(), dout, words * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for(int c = 0; c < (rank == 4 ? 3 : 1); ++c) for(int z = 0; z < TZ; ++z) for(int y = 0; y < TY; ++y) for(int x = 0; x < px; ++x)
    {
        float exp = float(c * vol + ((3 + z) * N[1] + 7 + y) * N[0] + 7 + x);
        if(o[c * px * TY * TZ + (z * TY + y) * px + x] != exp) ++bad;
    }
    printf("mismatches: %d\n", bad);
    return bad != 0;
}
