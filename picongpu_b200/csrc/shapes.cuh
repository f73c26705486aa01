// shapes.cuh — B-spline assignment functions NGP..PCS evaluated on grid points.
// Arithmetic follows include/picongpu/particles/shapes/{NGP,CIC,TSC,PQS,PCS}.hpp term by term (the last weight is
// computed as 1 - sum(others), e.g. TSC.hpp:85) so that the exact build reproduces the reference bit for bit.
#pragma once
#include "common.cuh"
#include "f2.cuh"

namespace picstep
{
    template<int SHAPE>
    struct Shape;

    // ---- polynomial pieces (T = float or F2: two evaluations in packed registers) -------------------------------
    template<class T>
    __device__ __forceinline__ T tsc_inner(T a)
    {
        T const sq = a * a;
        return T(0.75f) - sq;
    }
    template<class T>
    __device__ __forceinline__ T tsc_outer(T a)
    {
        T const t = T(3.0f / 2.0f) - a;
        T const sq = t * t;
        return T(0.5f) * sq;
    }
    template<class T>
    __device__ __forceinline__ T pqs_inner(T a)
    {
        T const sq = a * a;
        T const cu = sq * a;
        return T(1.0f / 6.0f) * (T(4.0f) - T(6.0f) * sq + T(3.0f) * cu);
    }
    template<class T>
    __device__ __forceinline__ T pqs_outer(T a)
    {
        T const t = T(2.0f) - a;
        T const cu = t * t * t;
        return T(1.0f / 6.0f) * cu;
    }
    template<class T>
    __device__ __forceinline__ T pcs_inner(T a)
    {
        T const sq = a * a;
        return T(115.f / 192.f) + sq * (T(-5.f / 8.f) + T(1.0f / 4.0f) * sq);
    }
    template<class T>
    __device__ __forceinline__ T pcs_mid(T a)
    {
        return T(1.f / 96.f) * (T(55.f) + T(4.f) * a * (T(5.f) - T(2.f) * a * (T(15.f) + T(2.f) * a * (T(-5.f) + a))));
    }
    template<class T>
    __device__ __forceinline__ T pcs_outer(T a)
    {
        T const t = T(5.f) - T(2.f) * a;
        T const sq = t * t;
        T const q = sq * sq;
        return T(1.f / 384.f) * q;
    }

    // SUPP = support in cells, BEGIN = lowest grid offset.  on(x, v): values at BEGIN..BEGIN+SUPP-1 for a particle
    // on support (x in [-0.5,0.5) for odd, [0,1) for even support).
    template<>
    struct Shape<0>
    {
        static constexpr int SUPP = 1, BEGIN = 0;
        template<class T>
        __device__ __forceinline__ static void on(T, T* v)
        {
            v[0] = T(1.0f);
        }
    };
    template<>
    struct Shape<1>
    {
        static constexpr int SUPP = 2, BEGIN = 0;
        template<class T>
        __device__ __forceinline__ static void on(T x, T* v)
        {
            v[0] = T(1.0f) - x;
            v[1] = x;
        }
    };
    template<>
    struct Shape<2>
    {
        static constexpr int SUPP = 3, BEGIN = -1;
        template<class T>
        __device__ __forceinline__ static void on(T x, T* v)
        {
            v[0] = tsc_outer(absT(T(-1.f) - x));
            v[1] = tsc_inner(absT(x));
            v[2] = T(1.0f) - (v[0] + v[1]);
        }
    };
    template<>
    struct Shape<3>
    {
        static constexpr int SUPP = 4, BEGIN = -1;
        template<class T>
        __device__ __forceinline__ static void on(T x, T* v)
        {
            v[0] = pqs_outer(absT(T(-1.f) - x));
            v[1] = pqs_inner(x);
            v[3] = pqs_outer(T(2.f) - x);
            v[2] = T(1.0f) - (v[0] + v[1] + v[3]);
        }
    };
    template<>
    struct Shape<4>
    {
        static constexpr int SUPP = 5, BEGIN = -2;
        template<class T>
        __device__ __forceinline__ static void on(T x, T* v)
        {
            v[0] = pcs_outer(absT(T(-2.f) - x));
            v[1] = pcs_mid(absT(T(-1.f) - x));
            v[2] = pcs_inner(absT(x));
            v[4] = pcs_outer(T(2.f) - x);
            v[3] = T(1.0f) - (v[0] + v[1] + v[2] + v[4]);
        }
    };

    /** Off-support array (SUPP+1 values at BEGIN..BEGIN+SUPP): the particle may sit one assignment cell further,
     * in which case the on-support values move up one slot (ChargeAssignment::shapeArray, e.g. TSC.hpp:140-153). */
    template<int SHAPE>
    __device__ __forceinline__ void shapeOff(float xx, bool shifted, float* v)
    {
        constexpr int S = Shape<SHAPE>::SUPP;
        float t[S];
        Shape<SHAPE>::on(shifted ? xx - 1.0f : xx, t);
        v[S] = shifted ? t[S - 1] : 0.0f;
#pragma unroll
        for(int i = S - 1; i >= 1; --i)
            v[i] = shifted ? t[i - 1] : t[i];
        v[0] = shifted ? 0.0f : t[0];
    }

    /** General (off-support) assignment function value, ChargeAssignment::operator() — used for charge density. */
    template<int SHAPE>
    __device__ __forceinline__ float shapeEval(float x)
    {
        float const a = fabsf(x);
        if constexpr(SHAPE == 0)
            return float(-0.5f <= x && x < 0.5f);
        else if constexpr(SHAPE == 1)
            return a < 1.0f ? 1.0f - a : 0.0f;
        else if constexpr(SHAPE == 2)
        {
            float const r1 = tsc_inner(a), r2 = tsc_outer(a);
            return a < 0.5f ? r1 : (a < 1.5f ? r2 : 0.0f);
        }
        else if constexpr(SHAPE == 3)
        {
            float const r1 = pqs_inner(a), r2 = pqs_outer(a);
            return a < 1.0f ? r1 : (a < 2.0f ? r2 : 0.0f);
        }
        else
        {
            float const r1 = pcs_inner(a), r2 = pcs_mid(a), r3 = pcs_outer(a);
            float r = r3;
            if(a < 0.5f)
                r = r1;
            else if(a < 1.5f)
                r = r2;
            return a < 2.5f ? r : 0.0f;
        }
    }

    // Interpolation margins of a shape incl. the Yee stagger shift (FieldToParticleInterpolation.hpp:49-52)
    template<int SHAPE>
    struct GatherMargin
    {
        static constexpr int LO = Shape<SHAPE>::SUPP / 2;
        static constexpr int UP = (Shape<SHAPE>::SUPP + 1) / 2;
    };
    // Current deposition margins (Esirkepov.hpp:42-45)
    template<int SHAPE>
    struct CurrentMargin
    {
        static constexpr int LO = Shape<SHAPE>::SUPP / 2 + 1 - (Shape<SHAPE>::SUPP + 1) % 2;
        static constexpr int UP = (Shape<SHAPE>::SUPP + 1) / 2 + 1;
    };
} // namespace picstep
