// esirkepov.cuh — pieces of the Esirkepov current solver shared by the deposition kernels
// (fields/currentDeposition/relayPoint.hpp:48-63, Esirkepov/Esirkepov.hpp:147-242).
#pragma once
#include "common.cuh"
#include "shapes.cuh"

namespace picstep
{
    // relayPoint.hpp:48-63 (only the two assignment-cell indices are needed by Esirkepov)
    template<bool EVEN>
    __device__ __forceinline__ float relay(int& i1, int& i2, float x1, float x2)
    {
        if constexpr(EVEN)
        {
            i1 = __float2int_rd(x1);
            i2 = __float2int_rd(x2);
            return i1 == i2 ? x2 : float(max(i1, i2));
        }
        else
        {
            i1 = __float2int_rd(x1 + 0.5f);
            i2 = __float2int_rd(x2 + 0.5f);
            return i1 == i2 ? x2 : float(i1 + i2) / 2.0f;
        }
    }

    /** red.global.add.f32 on a pointer that is known to be global memory (atomicAdd on a generic pointer compiles to
     * an address-space test plus both the shared CAS loop and the global atomic) */
    __device__ __forceinline__ void redGlobal(float* p, float v)
    {
        asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
    }

    /** One rotated 1-D pass of Esirkepov (Esirkepov.hpp:147-242) for a single particle with global atomics
     * (red.global.add.f32); loop bounds and summation order are the reference's.  R0,R1,R2: original axes of the
     * rotated i,j,k.  `origin` points at the node of the particle's cell shifted by gridShift, `stride` are the
     * element strides of the three grid axes, p0/p1 the start/end point in that frame. */
    template<int SHAPE, int R0, int R1, int R2>
    __device__ __forceinline__ void esirkepov1DGlobal(
        float* __restrict__ origin,
        long long const stride[3],
        int const status[3],
        float const p0[3],
        float const p1[3],
        float currentSurfaceDensity)
    {
        using S = Shape<SHAPE>;
        if(p0[R2] == p1[R2])
            return;
        constexpr int begin = S::BEGIN, end = S::BEGIN + S::SUPP;
        float s0i[S::SUPP + 1], s1i[S::SUPP + 1], s0j[S::SUPP + 1], s1j[S::SUPP + 1], s0k[S::SUPP + 1], s1k[S::SUPP + 1];
        shapeOff<SHAPE>(p0[R0], !(status[R0] & 2), s0i);
        shapeOff<SHAPE>(p1[R0], !(status[R0] & 4), s1i);
        shapeOff<SHAPE>(p0[R1], !(status[R1] & 2), s0j);
        shapeOff<SHAPE>(p1[R1], !(status[R1] & 4), s1j);
        shapeOff<SHAPE>(p0[R2], !(status[R2] & 2), s0k);
        shapeOff<SHAPE>(p1[R2], !(status[R2] & 4), s1k);
        int const leaveI = status[R0] & 1, leaveJ = status[R1] & 1, leaveK = status[R2] & 1;
#pragma unroll
        for(int i = begin; i < end + 1; ++i)
            if(i < end + leaveI)
            {
                float const a0 = s0i[i - begin];
                float const da = s1i[i - begin] - a0;
#pragma unroll
                for(int j = begin; j < end + 1; ++j)
                    if(j < end + leaveJ)
                    {
                        float const b0 = s0j[j - begin];
                        float const db = s1j[j - begin] - b0;
                        float const tmp = -currentSurfaceDensity * (a0 * b0 + 0.5f * (da * b0 + a0 * db) + (1.0f / 3.0f) * db * da);
                        float acc = 0.0f;
#pragma unroll
                        for(int k = begin; k < end; ++k)
                            if(k < end + leaveK - 1)
                            {
                                float const W = (s1k[k - begin] - s0k[k - begin]) * tmp;
                                acc += W;
                                redGlobal(origin + i * stride[R0] + j * stride[R1] + k * stride[R2], acc);
                            }
                    }
            }
    }
    /** Slow path of the run kernel: one particle, all three components, arguments by value so that the caller keeps
     * nothing in local memory.  status = status[0] | status[1] << 3 | status[2] << 6. */
    template<int SHAPE>
    __device__ __noinline__ void esirkepovParticleGlobal(float* jx, float* jy, float* jz, long long strideY, long long strideZ, int statusPacked, float p0x, float p0y, float p0z, float p1x, float p1y, float p1z, float csdx, float csdy, float csdz)
    {
        long long const stride[3] = {1, strideY, strideZ};
        int const status[3] = {statusPacked & 7, (statusPacked >> 3) & 7, (statusPacked >> 6) & 7};
        float const p0[3] = {p0x, p0y, p0z}, p1[3] = {p1x, p1y, p1z};
        esirkepov1DGlobal<SHAPE, 1, 2, 0>(jx, stride, status, p0, p1, csdx);
        esirkepov1DGlobal<SHAPE, 2, 0, 1>(jy, stride, status, p0, p1, csdy);
        esirkepov1DGlobal<SHAPE, 0, 1, 2>(jz, stride, status, p0, p1, csdz);
    }

    /** emz::DepositCurrent::cptCurrent1D (EmZ/DepositCurrent.hpp:77-118) of one on-support segment with global atomics:
     * fixed loop bounds SUPP x SUPP x (SUPP-1), origin = node of the segment's assignment cell. */
    template<int SHAPE, int R0, int R1, int R2>
    __device__ __forceinline__ void emz1DGlobal(float* __restrict__ origin, long long const stride[3], float const p0[3], float const p1[3], float currentSurfaceDensity)
    {
        using S = Shape<SHAPE>;
        if(p0[R2] == p1[R2])
            return;
        constexpr int begin = S::BEGIN;
        float s0i[S::SUPP], s1i[S::SUPP], s0j[S::SUPP], s1j[S::SUPP], s0k[S::SUPP], s1k[S::SUPP];
        S::on(p0[R0], s0i);
        S::on(p1[R0], s1i);
        S::on(p0[R1], s0j);
        S::on(p1[R1], s1j);
        S::on(p0[R2], s0k);
        S::on(p1[R2], s1k);
#pragma unroll
        for(int i = 0; i < S::SUPP; ++i)
        {
            float const a0 = s0i[i];
            float const da = s1i[i] - a0;
#pragma unroll
            for(int j = 0; j < S::SUPP; ++j)
            {
                float const b0 = s0j[j];
                float const db = s1j[j] - b0;
                float const tmp = -currentSurfaceDensity * (a0 * b0 + 0.5f * (da * b0 + a0 * db) + (1.0f / 3.0f) * db * da);
                float acc = 0.0f;
#pragma unroll
                for(int k = 0; k < S::SUPP - 1; ++k)
                {
                    acc += (s1k[k] - s0k[k]) * tmp;
                    redGlobal(origin + (begin + i) * stride[R0] + (begin + j) * stride[R1] + (begin + k) * stride[R2], acc);
                }
            }
        }
    }

    /** Slow path of the run kernel for EmZ: one segment, all three components */
    template<int SHAPE>
    __device__ __noinline__ void emzSegmentGlobal(float* jx, float* jy, float* jz, long long strideY, long long strideZ, float p0x, float p0y, float p0z, float p1x, float p1y, float p1z, float csdx, float csdy, float csdz)
    {
        long long const stride[3] = {1, strideY, strideZ};
        float const p0[3] = {p0x, p0y, p0z}, p1[3] = {p1x, p1y, p1z};
        emz1DGlobal<SHAPE, 1, 2, 0>(jx, stride, p0, p1, csdx);
        emz1DGlobal<SHAPE, 2, 0, 1>(jy, stride, p0, p1, csdy);
        emz1DGlobal<SHAPE, 0, 1, 2>(jz, stride, p0, p1, csdz);
    }
} // namespace picstep
