// tma.cuh — TMA tile loads of the E/B supercell tiles (cp.async.bulk.tensor + mbarrier, sm_100a).
// One descriptor per field: a 4-D tensor (x, y, z, component) over the SoA planes of the field, box = the
// (TX rounded up to 4) x TY x TZ x 3 tile of one supercell incl. the interpolation margins.  The box lands densely in
// shared memory as [comp][z][y][x], which is the layout Tile<SHAPE> describes.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace picstep
{
    struct alignas(64) TileMaps
    {
        CUtensorMap E, B;
        int lead; // floats between the descriptor's base (the allocation) and element 0 of the field
    };

    __device__ __forceinline__ uint32_t smemAddr(void const* p)
    {
        return static_cast<uint32_t>(__cvta_generic_to_shared(p));
    }

    __device__ __forceinline__ void mbarInit(uint64_t* bar, uint32_t count)
    {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }

    __device__ __forceinline__ void mbarExpectTx(uint64_t* bar, uint32_t bytes)
    {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory");
    }

    __device__ __forceinline__ void mbarWait(uint64_t* bar, uint32_t parity)
    {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "WAIT_%=:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
            "@p bra DONE_%=;\n"
            "bra WAIT_%=;\n"
            "DONE_%=:\n"
            "}\n" ::"r"(smemAddr(bar)),
            "r"(parity)
            : "memory");
    }

    /** box of `map` at element coordinates (x, y, z, comp 0) -> dst (128-byte aligned shared memory); completes on bar.
     * x has to be a multiple of four floats (16-byte aligned box start), measured on B200: illegal instruction otherwise */
    __device__ __forceinline__ void tmaLoadTile(void* dst, CUtensorMap const* map, int x, int y, int z, uint64_t* bar)
    {
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(smemAddr(dst)),
            "l"(reinterpret_cast<uint64_t>(map)),
            "r"(x),
            "r"(y),
            "r"(z),
            "r"(0),
            "r"(smemAddr(bar))
            : "memory");
    }
} // namespace picstep
