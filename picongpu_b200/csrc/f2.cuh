// f2.cuh — two fp32 values in one 64-bit register pair.  sm_100a has packed FADD2 / FMUL2 / FFMA2: one issue slot
// for two IEEE fp32 operations, which is what an issue-bound kernel needs.  In the exact build (-fmad=false) every
// operation is done per half with separately rounded multiply and add, so both halves are bit-identical to the
// scalar code of the reference.
#pragma once
#include <cuda_runtime.h>

namespace picstep
{
    struct F2
    {
        float x, y;
        __device__ __forceinline__ F2()
        {
        }
        __device__ __forceinline__ F2(float a) : x(a), y(a)
        {
        }
        __device__ __forceinline__ F2(float a, float b) : x(a), y(b)
        {
        }
    };

#ifdef PICSTEP_EXACT
    __device__ __forceinline__ F2 operator+(F2 a, F2 b)
    {
        return F2(a.x + b.x, a.y + b.y);
    }
    __device__ __forceinline__ F2 operator-(F2 a, F2 b)
    {
        return F2(a.x - b.x, a.y - b.y);
    }
    __device__ __forceinline__ F2 operator*(F2 a, F2 b)
    {
        return F2(a.x * b.x, a.y * b.y);
    }
    /** a * b + c, multiply and add rounded separately (the exact build has no FMA contraction) */
    __device__ __forceinline__ F2 fma2(F2 a, F2 b, F2 c)
    {
        return F2(a.x * b.x + c.x, a.y * b.y + c.y);
    }
#else
    __device__ __forceinline__ F2 operator+(F2 a, F2 b)
    {
        F2 d;
        asm("{ .reg .b64 ra, rb, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; add.rn.f32x2 rd, ra, rb; mov.b64 {%0,%1}, rd; }"
            : "=f"(d.x), "=f"(d.y)
            : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
        return d;
    }
    __device__ __forceinline__ F2 operator-(F2 a, F2 b)
    {
        F2 d;
        asm("{ .reg .b64 ra, rb, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; sub.rn.f32x2 rd, ra, rb; mov.b64 {%0,%1}, rd; }"
            : "=f"(d.x), "=f"(d.y)
            : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
        return d;
    }
    __device__ __forceinline__ F2 operator*(F2 a, F2 b)
    {
        F2 d;
        asm("{ .reg .b64 ra, rb, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mul.rn.f32x2 rd, ra, rb; mov.b64 {%0,%1}, rd; }"
            : "=f"(d.x), "=f"(d.y)
            : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
        return d;
    }
    /** a * b + c as one FFMA2 */
    __device__ __forceinline__ F2 fma2(F2 a, F2 b, F2 c)
    {
        F2 d;
        asm("{ .reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mov.b64 rc, {%6,%7}; fma.rn.f32x2 rd, ra, rb, rc; "
            "mov.b64 {%0,%1}, rd; }"
            : "=f"(d.x), "=f"(d.y)
            : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
        return d;
    }
#endif
    __device__ __forceinline__ F2 operator-(F2 a)
    {
        return F2(-a.x, -a.y);
    }
    __device__ __forceinline__ F2 absT(F2 a)
    {
        return F2(fabsf(a.x), fabsf(a.y));
    }
    __device__ __forceinline__ float absT(float a)
    {
        return fabsf(a);
    }
} // namespace picstep
