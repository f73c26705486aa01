// pusher.cuh — device functions shared by the push kernels: field-to-particle interpolation on the Yee grid
// (FieldToParticleInterpolation.hpp:97-124), Boris / Vay pushers (particlePusherBoris.hpp:42-91,
// particlePusherVay.hpp:43-112), Gamma / Velocity (Gamma.hpp:30-38, Velocity.hpp:28-38).
#pragma once
#include "common.cuh"
#include "shapes.cuh"

namespace picstep
{
    // Yee stagger (include/picongpu/fields/YeeCell.hpp:70-130); component c of E sits at +0.5 along c,
    // component c of B at +0.5 along the two other axes.
    __device__ __forceinline__ float stagE(int comp, int d)
    {
        return comp == d ? 0.5f : 0.0f;
    }
    __device__ __forceinline__ float stagB(int comp, int d)
    {
        return comp == d ? 0.0f : 0.5f;
    }

    template<int SHAPE>
    struct Tile
    {
        static constexpr int LO = GatherMargin<SHAPE>::LO, UP = GatherMargin<SHAPE>::UP;
        static constexpr int TX = SCX + LO + UP, TY = SCY + LO + UP, TZ = SCZ + LO + UP;
        // rows are padded to a multiple of four floats: a TMA box row has to be a multiple of 16 bytes
        static constexpr int PX = (TX + 3) / 4 * 4;
        static constexpr int TV = PX * TY * TZ;
        // the B block [3][TZ][TY][PX] is followed by the E block at the next 128-byte boundary (TMA destination)
        static constexpr int HALF = (3 * TV + 31) / 32 * 32;
        static constexpr int WORDS = 2 * HALF;
        static constexpr uint32_t BYTES_PER_FIELD = 3u * TV * sizeof(float);
    };

    /** Interpolate one field component to the particle (FieldToParticleInterpolation.hpp:97-124 +
     * ShiftCoordinateSystem.hpp:54-79 + AssignedTrilinearInterpolation.hpp:54-86; x innermost). */
    template<int SHAPE, bool IS_B>
    __device__ __forceinline__ float gatherComp(float const* __restrict__ t, int comp, int lx, int ly, int lz, float px, float py, float pz)
    {
        using S = Shape<SHAPE>;
        using T = Tile<SHAPE>;
        constexpr bool even = (S::SUPP % 2) == 0;
        constexpr int begin = -S::SUPP / 2 + (S::SUPP + 1) % 2;
        float sx[S::SUPP], sy[S::SUPP], sz[S::SUPP];
        int base;
        {
            float const p[3] = {px, py, pz};
            int const l[3] = {lx, ly, lz};
            int sh[3];
            float q[3];
#pragma unroll
            for(int d = 0; d < 3; ++d)
            {
                float const fp = IS_B ? stagB(comp, d) : stagE(comp, d);
                float const v = p[d] - fp - 0.5f;
                if constexpr(even)
                    sh[d] = v >= -0.5f ? 0 : -1;
                else
                    sh[d] = v >= 0.0f ? 1 : 0;
                q[d] = v - float(sh[d]) + 0.5f;
                sh[d] += l[d] + T::LO + begin;
            }
            S::on(q[0], sx);
            S::on(q[1], sy);
            S::on(q[2], sz);
            base = (sh[2] * T::TY + sh[1]) * T::PX + sh[0];
        }
        float rz = 0.0f;
#pragma unroll
        for(int z = 0; z < S::SUPP; ++z)
        {
            float ry = 0.0f;
#pragma unroll
            for(int y = 0; y < S::SUPP; ++y)
            {
                float rx = 0.0f;
#pragma unroll
                for(int x = 0; x < S::SUPP; ++x)
                    rx += t[base + (z * T::TY + y) * T::PX + x] * sx[x];
                ry += rx * sy[y];
            }
            rz += ry * sz[z];
        }
        return rz;
    }

    /** All six components at once.  Production build: the components are interpolated in the pairs (B_c, E_c),
     * whose Yee staggers are complementary on every axis, with packed FFMA2 — the weights of the un-staggered and the
     * staggered frame of an axis are evaluated together as the register pairs {u,s} and {s,u}.  Same operations per
     * component as gatherComp (x innermost), half the issue slots.  Exact build: gatherComp per component. */
    template<int SHAPE>
    __device__ __forceinline__ void gatherEB(float const* __restrict__ tB, float const* __restrict__ tE, int lx, int ly, int lz, float px, float py, float pz, float Ef[3], float Bf[3])
    {
        using S = Shape<SHAPE>;
        using T = Tile<SHAPE>;
#ifdef PICSTEP_EXACT
#pragma unroll
        for(int k = 0; k < 3; ++k)
        {
            Bf[k] = gatherComp<SHAPE, true>(tB + k * T::TV, k, lx, ly, lz, px, py, pz);
            Ef[k] = gatherComp<SHAPE, false>(tE + k * T::TV, k, lx, ly, lz, px, py, pz);
        }
#else
        constexpr bool even = (S::SUPP % 2) == 0;
        constexpr int begin = -S::SUPP / 2 + (S::SUPP + 1) % 2;
        float const p[3] = {px, py, pz};
        int const l[3] = {lx, ly, lz};
        F2 wUS[3][S::SUPP], wSU[3][S::SUPP];
        int bU[3], bS[3];
#pragma unroll
        for(int d = 0; d < 3; ++d)
        {
            float const vu = p[d] - 0.0f - 0.5f, vs = p[d] - 0.5f - 0.5f;
            int shu, shs;
            if constexpr(even)
            {
                shu = vu >= -0.5f ? 0 : -1;
                shs = vs >= -0.5f ? 0 : -1;
            }
            else
            {
                shu = vu >= 0.0f ? 1 : 0;
                shs = vs >= 0.0f ? 1 : 0;
            }
            float const qu = vu - float(shu) + 0.5f, qs = vs - float(shs) + 0.5f;
            bU[d] = shu + l[d] + T::LO + begin;
            bS[d] = shs + l[d] + T::LO + begin;
            S::on(F2(qu, qs), wUS[d]);
            S::on(F2(qs, qu), wSU[d]);
        }
#pragma unroll
        for(int c = 0; c < 3; ++c)
        {
            // B_c is staggered on the axes != c, E_c on axis c: pair {B,E} uses {u,s} weights on axis c, {s,u} elsewhere
            int const bx = c == 0 ? bU[0] : bS[0], by = c == 1 ? bU[1] : bS[1], bz = c == 2 ? bU[2] : bS[2];
            int const ex = c == 0 ? bS[0] : bU[0], ey = c == 1 ? bS[1] : bU[1], ez = c == 2 ? bS[2] : bU[2];
            float const* __restrict__ pb = tB + c * T::TV + (bz * T::TY + by) * T::PX + bx;
            float const* __restrict__ pe = tE + c * T::TV + (ez * T::TY + ey) * T::PX + ex;
            F2 const* wx = c == 0 ? wUS[0] : wSU[0];
            F2 const* wy = c == 1 ? wUS[1] : wSU[1];
            F2 const* wz = c == 2 ? wUS[2] : wSU[2];
            F2 rz(0.0f);
#pragma unroll
            for(int z = 0; z < S::SUPP; ++z)
            {
                F2 ry(0.0f);
#pragma unroll
                for(int y = 0; y < S::SUPP; ++y)
                {
                    F2 rx(0.0f);
#pragma unroll
                    for(int x = 0; x < S::SUPP; ++x)
                    {
                        int const o = (z * T::TY + y) * T::PX + x;
                        rx = fma2(F2(pb[o], pe[o]), wx[x], rx);
                    }
                    ry = fma2(rx, wy[y], ry);
                }
                rz = fma2(ry, wz[z], rz);
            }
            Bf[c] = rz.x;
            Ef[c] = rz.y;
        }
#endif
    }

    __device__ __forceinline__ float norm2(float x, float y, float z)
    {
        float t = x * x;
        t += y * y;
        t += z * z;
        return t;
    }

    // Gamma.hpp:30-38
    __device__ __forceinline__ float gammaOf(float c, float ux, float uy, float uz, float mass)
    {
        float const c2 = c * c;
        float const r = ps_div(1.0f, mass * mass * c2);
        return ps_sqrt(1.0f + norm2(ux, uy, uz) * r);
    }

    // Velocity.hpp:28-38: v = p * rsqrt(m^2 + p^2/c^2)
    __device__ __forceinline__ void velocityOf(float rc2, float mass, float ux, float uy, float uz, float& vx, float& vy, float& vz)
    {
        float const t = ps_rsqrt(mass * mass + norm2(ux, uy, uz) * rc2);
        vx = t * ux;
        vy = t * uy;
        vz = t * uz;
    }

    // particlePusherBoris.hpp:42-91
    __device__ __forceinline__ void boris(DevParams const& P, float mass, float charge, float const E[3], float const B[3], float u[3])
    {
        float const QoM = ps_div(charge, mass);
        float const dt = P.dt;
        float m[3], t[3];
#pragma unroll
        for(int d = 0; d < 3; ++d)
            m[d] = u[d] + 0.5f * charge * E[d] * dt;
        float const gr = ps_div(1.0f, gammaOf(P.c, m[0], m[1], m[2], mass));
#pragma unroll
        for(int d = 0; d < 3; ++d)
            t[d] = 0.5f * QoM * B[d] * gr * dt;
        float const sf = ps_div(1.0f, 1.0f + norm2(t[0], t[1], t[2]));
        float s[3];
#pragma unroll
        for(int d = 0; d < 3; ++d)
            s[d] = 2.0f * t[d] * sf;
        float pr[3];
        pr[0] = m[0] + (m[1] * t[2] - m[2] * t[1]);
        pr[1] = m[1] + (m[2] * t[0] - m[0] * t[2]);
        pr[2] = m[2] + (m[0] * t[1] - m[1] * t[0]);
        float pl[3];
        pl[0] = m[0] + (pr[1] * s[2] - pr[2] * s[1]);
        pl[1] = m[1] + (pr[2] * s[0] - pr[0] * s[2]);
        pl[2] = m[2] + (pr[0] * s[1] - pr[1] * s[0]);
#pragma unroll
        for(int d = 0; d < 3; ++d)
            u[d] = pl[d] + 0.5f * charge * E[d] * dt;
    }

    // particlePusherVay.hpp:43-112 (sqrt section in fp64: sqrt_Vay = precision64Bit, param/pusher.param:62)
    __device__ __forceinline__ void vay(DevParams const& P, float rc2, float mass, float charge, float const E[3], float const B[3], float u[3])
    {
        float const factor = float(0.5 * double(charge) * double(P.dt));
        float v0[3];
        velocityOf(rc2, mass, u[0], u[1], u[2], v0[0], v0[1], v0[2]);
        float cr[3];
        cr[0] = v0[1] * B[2] - v0[2] * B[1];
        cr[1] = v0[2] * B[0] - v0[0] * B[2];
        cr[2] = v0[0] * B[1] - v0[1] * B[0];
        float mp[3];
#pragma unroll
        for(int d = 0; d < 3; ++d)
        {
            float const m0 = u[d] + factor * (E[d] + cr[d]);
            mp[d] = m0 + factor * E[d];
        }
        float const gp = gammaOf(P.c, mp[0], mp[1], mp[2], mass);
        double tau[3];
#pragma unroll
        for(int d = 0; d < 3; ++d)
            tau[d] = double(factor / mass * B[d]);
        double dpt = double(mp[0]) * tau[0];
        dpt += double(mp[1]) * tau[1];
        dpt += double(mp[2]) * tau[2];
        double const ustar = dpt / double(P.c * mass);
        double tau2 = tau[0] * tau[0];
        tau2 += tau[1] * tau[1];
        tau2 += tau[2] * tau[2];
        double const sigma = double(gp * gp) - tau2;
        double const gplus = sqrt(0.5 * (sigma + sqrt(sigma * sigma + 4.0 * (tau2 + ustar * ustar))));
        float t[3];
#pragma unroll
        for(int d = 0; d < 3; ++d)
            t[d] = float(tau[d] * (1.0 / gplus));
        float const s = ps_div(1.0f, 1.0f + norm2(t[0], t[1], t[2]));
        float dp = mp[0] * t[0];
        dp += mp[1] * t[1];
        dp += mp[2] * t[2];
        cr[0] = mp[1] * t[2] - mp[2] * t[1];
        cr[1] = mp[2] * t[0] - mp[0] * t[2];
        cr[2] = mp[0] * t[1] - mp[1] * t[0];
#pragma unroll
        for(int d = 0; d < 3; ++d)
            u[d] = s * (mp[d] + dp * t[d] + cr[d]);
    }

    // particlePusherHigueraCary.hpp:45-145 (auxiliary quantities in fp64: sqrt_HigueraCary = precision64Bit,
    // param/pusher.param:71; Gamma<> itself computes in fp32)
    __device__ __forceinline__ void higueraCary(DevParams const& P, float mass, float charge, float const E[3], float const B[3], float u[3])
    {
        float const dt = P.dt;
        float he[3], mm[3];
        double mmd[3], tau[3];
#pragma unroll
        for(int d = 0; d < 3; ++d)
        {
            he[d] = 0.5f * charge * E[d] * dt;
            mm[d] = u[d] + he[d];
            mmd[d] = double(mm[d]);
            tau[d] = double(ps_div(0.5f * B[d] * charge * dt, mass));
        }
        double const gm = double(gammaOf(P.c, mm[0], mm[1], mm[2], mass));
        double tau2 = tau[0] * tau[0];
        tau2 += tau[1] * tau[1];
        tau2 += tau[2] * tau[2];
        double const sigma = gm * gm - tau2;
        double dpt = mmd[0] * tau[0];
        dpt += mmd[1] * tau[1];
        dpt += mmd[2] * tau[2];
        double const ustar = dpt / double(mass * P.c);
        double const gplus = sqrt(0.5 * (sigma + sqrt(sigma * sigma + 4.0 * (tau2 + ustar * ustar))));
        double t[3];
#pragma unroll
        for(int d = 0; d < 3; ++d)
            t[d] = tau[d] / gplus;
        double t2 = t[0] * t[0];
        t2 += t[1] * t[1];
        t2 += t[2] * t[2];
        double const sf = 1.0 / (1.0 + t2);
        double dmt = mmd[0] * t[0];
        dmt += mmd[1] * t[1];
        dmt += mmd[2] * t[2];
        double const cr[3] = {mmd[1] * t[2] - mmd[2] * t[1], mmd[2] * t[0] - mmd[0] * t[2], mmd[0] * t[1] - mmd[1] * t[0]};
        double mp[3];
#pragma unroll
        for(int d = 0; d < 3; ++d)
            mp[d] = sf * (mmd[d] + dmt * t[d] + cr[d]);
        double const c2[3] = {mp[1] * t[2] - mp[2] * t[1], mp[2] * t[0] - mp[0] * t[2], mp[0] * t[1] - mp[1] * t[0]};
#pragma unroll
        for(int d = 0; d < 3; ++d)
        {
            float const diff = he[d] + float(c2[d]);
            u[d] = float(mp[d]) + diff;
        }
    }

    /** UsedParticlePusher dispatch: 0 Boris, 1 Vay, 2 HigueraCary */
    template<int PUSHER>
    __device__ __forceinline__ void pushMomentum(DevParams const& P, float rc2, float mass, float charge, float const E[3], float const B[3], float u[3])
    {
        if constexpr(PUSHER == 0)
            boris(P, mass, charge, E, B, u);
        else if constexpr(PUSHER == 1)
            vay(P, rc2, mass, charge, E, B, u);
        else
            higueraCary(P, mass, charge, E, B, u);
    }

} // namespace picstep
