// fields.cu — Maxwell solver (reference kernels K9/K10: KernelAddCurrentDensity, KernelUpdateField,
// include/picongpu/fields/MaxwellSolver/FDTD/FDTDBase.kernel:51-199, AddCurrentDensity.kernel:38-96), guard
// exchange kernels (K8: include/pmacc/fields/operations/{CopyGuardToExchange,AddExchangeToBorder}.hpp) and the
// field-side reductions used as parity observables.
//
// Fields are SoA planes (one float array per component, x fastest, guards included).  The curl stencils are
// pure streaming kernels: one thread per cell, x along the warp so every load is a coalesced 128-byte row
// segment; the +-1 (Yee) / +-2 (Lehe) neighbours in y and z come out of L1/L2.  Per cell and launch the
// algorithmic traffic is 12 B read + 24 B read/write.
#include "common.cuh"
#include "shapes.cuh"
#include "tma.cuh"

namespace picstep
{
    // MODE 0: forward difference (ForwardDerivative.hpp:58-63), 1: backward (BackwardDerivative.hpp:58-63),
    // 2: Lehe (Lehe/Derivative.hpp:113-146 along the Cherenkov-free direction, :225-236 otherwise)
    template<int MODE, int DIR>
    __device__ __forceinline__ float deriv(DevParams const& P, LeheCoeffs const& L, float const* __restrict__ f, long long i, long long const st[3])
    {
        float const h = P.cell[DIR];
        long long const s = st[DIR];
        if constexpr(MODE == 0)
            return (f[i + s] - f[i]) / h;
        else if constexpr(MODE == 1)
            return (f[i] - f[i - s]) / h;
        else
        {
            auto fwd = [&](long long j) { return (f[j + s] - f[j]) / h; };
            if(DIR == P.lehe_dir)
            {
                long long const s1 = st[(DIR + 1) % 3], s2 = st[(DIR + 2) % 3];
                float r = L.alpha[DIR] * fwd(i) + L.beta1[DIR] * fwd(i + s1);
                r = r + L.beta1[DIR] * fwd(i - s1);
                r = r + L.beta2[DIR] * fwd(i + s2);
                r = r + L.beta2[DIR] * fwd(i - s2);
                r = r + L.delta[DIR] * (f[i + 2 * s] - f[i - s]) / h;
                return r;
            }
            else
            {
                float const beta = 0.125f;
                float const alpha = 1.0f - 2.0f * beta;
                long long const sc = st[P.lehe_dir];
                float r = alpha * fwd(i) + beta * fwd(i + sc);
                r = r + beta * fwd(i - sc);
                return r;
            }
        }
    }

    // curl (differentiation/Curl.hpp:84-90) = (dFz/dy - dFy/dz, dFx/dz - dFz/dx, dFy/dx - dFx/dy)
    template<int MODE>
    __device__ __forceinline__ void curl(DevParams const& P, LeheCoeffs const& L, Field3 const& F, long long i, long long const st[3], float out[3])
    {
        float const dzdy = deriv<MODE, 1>(P, L, F.c[2], i, st);
        float const dydz = deriv<MODE, 2>(P, L, F.c[1], i, st);
        float const dxdz = deriv<MODE, 2>(P, L, F.c[0], i, st);
        float const dzdx = deriv<MODE, 0>(P, L, F.c[2], i, st);
        float const dydx = deriv<MODE, 0>(P, L, F.c[1], i, st);
        float const dxdy = deriv<MODE, 1>(P, L, F.c[0], i, st);
        out[0] = dzdy - dydz;
        out[1] = dxdz - dzdx;
        out[2] = dydx - dxdy;
    }

    // UpdateBHalfFunctor: B -= curlE * 0.5 * dt   (FDTDBase.kernel:115-121)
    template<int MODE>
    __global__ void __launch_bounds__(256) updateBHalfKernel(DevParams P, LeheCoeffs L, Field3 E, Field3 B)
    {
        int const x = blockIdx.x * blockDim.x + threadIdx.x;
        int const y = blockIdx.y * blockDim.y + threadIdx.y;
        int const z = blockIdx.z * blockDim.z + threadIdx.z;
        if(x >= P.n[0] || y >= P.n[1] || z >= P.n[2])
            return;
        long long const st[3] = {1, P.N[0], (long long) P.N[0] * P.N[1]};
        long long const i = fidx(P, x + P.g[0], y + P.g[1], z + P.g[2]);
        float cu[3];
        curl<MODE>(P, L, E, i, st, cu);
#pragma unroll
        for(int c = 0; c < 3; ++c)
            B.c[c][i] -= cu[c] * 0.5f * P.dt;
    }

    // UpdateEFunctor: E += curlB * c^2 * dt  (FDTDBase.kernel:74-81); optionally fused with the current term
    // E += coeff * J (currentInterpolation/None.hpp:60-64) when ADDJ (two separate roundings, same as two kernels)
    __global__ void __launch_bounds__(256) updateEKernel(DevParams P, LeheCoeffs L, Field3 E, Field3 B)
    {
        int const x = blockIdx.x * blockDim.x + threadIdx.x;
        int const y = blockIdx.y * blockDim.y + threadIdx.y;
        int const z = blockIdx.z * blockDim.z + threadIdx.z;
        if(x >= P.n[0] || y >= P.n[1] || z >= P.n[2])
            return;
        long long const st[3] = {1, P.N[0], (long long) P.N[0] * P.N[1]};
        long long const i = fidx(P, x + P.g[0], y + P.g[1], z + P.g[2]);
        float cu[3];
        curl<1>(P, L, B, i, st, cu);
        float const c2 = P.c * P.c;
#pragma unroll
        for(int c = 0; c < 3; ++c)
            E.c[c][i] += cu[c] * c2 * P.dt;
    }

    // ---- TMA-staged Yee update (the bandwidth-bound form the field solver runs in) -------------------------------
    // One CTA updates a brick of FD_TX x FD_TY x FD_TZ cells.  The source field (E for the B update, B for the E update)
    // arrives as ONE 4-D TMA box (x, y, z, component) incl. the one-cell halo the two-point differences need; the
    // destination is read-modified-written with 16-byte accesses straight from / to HBM.  Bricks start at x positions
    // that are multiples of four floats of the allocation (TMA box origins and the float4 accesses need 16-byte
    // alignment; the fields sit `lead` floats into their allocation for the particle tiles, common.cuh), cells outside
    // the active region are masked.  Per cell: 12 B read through TMA (x1.4 with halos, most of it L2 hits) + 24 B
    // read/write, +12 B when the current term is fused in.
    // KIND 0: B -= curl_forward(E) * 0.5 * dt   (UpdateBHalfFunctor, FDTDBase.kernel:115-121, ForwardDerivative.hpp:58-63)
    // KIND 1: E += curl_backward(B) * c^2 * dt  (UpdateEFunctor, FDTDBase.kernel:74-81, BackwardDerivative.hpp:58-63)
    //         ADDJ: followed by E += (-(1/eps0) * dt) * J (AddCurrentDensity.kernel:38-96 with currentInterpolation::None,
    //         None.hpp:60-64) -- two separately rounded additions, exactly as the reference's two kernels produce them.
    // The arithmetic per cell is the one of deriv<0/1> above, so the exact build stays bit-identical to the oracle.
    constexpr int FD_TX = 64, FD_TY = 8, FD_TZ = 4;
    constexpr int FD_BX = FD_TX + 8, FD_BY = FD_TY + 2, FD_BZ = FD_TZ + 1; // box: x from -4 to +67, y from -1, z from -1 (E update) or 0
    constexpr int FD_PLANE = FD_BX * FD_BY, FD_COMP = FD_PLANE * FD_BZ;
    constexpr uint32_t FD_BYTES = 3u * FD_COMP * sizeof(float);

    template<int KIND, bool ADDJ>
    __global__ void __launch_bounds__(256, 3) fdtdTmaKernel(DevParams P, Field3 D, Field3 J, const __grid_constant__ CUtensorMap srcMap, int lead)
    {
        extern __shared__ __align__(128) float box[];
        __shared__ uint64_t bar;
        // brick origin in allocation coordinates (x) / padded-grid coordinates (y, z)
        int const X0 = ((P.g[0] + lead) & ~3) + blockIdx.x * FD_TX;
        int const Y0 = P.g[1] + blockIdx.y * FD_TY, Z0 = P.g[2] + blockIdx.z * FD_TZ;
        constexpr int ZLO = KIND == 1 ? 1 : 0; // planes below the brick
        if(threadIdx.x == 0)
        {
            mbarInit(&bar, 1);
            mbarExpectTx(&bar, FD_BYTES);
            tmaLoadTile(box, &srcMap, X0 - 4, Y0 - 1, Z0 - ZLO, &bar);
        }
        __syncthreads();
        int const tx = threadIdx.x & 15, ty = (threadIdx.x >> 4) & 7, tz = threadIdx.x >> 7; // 16 x 8 x 2 threads, 4 cells in x each
        int const xa = X0 + 4 * tx; // allocation x of the thread's first cell
        int const xlo = P.g[0] + lead, xhi = xlo + P.n[0];
        bool const yok = Y0 + ty < P.g[1] + P.n[1];
        // two-point difference divided by the cell size: the exact build divides like the reference
        // (Forward/BackwardDerivative.hpp:58-63); the production build multiplies with the reciprocal -- 24 IEEE divisions
        // per four cells were a third of the kernel's instructions
#ifdef PICSTEP_EXACT
        float const hx = P.cell[0], hy = P.cell[1], hz = P.cell[2];
#    define FD_DIFF(a, b, h) (((a) - (b)) / (h))
#else
        float const hx = 1.0f / P.cell[0], hy = 1.0f / P.cell[1], hz = 1.0f / P.cell[2];
#    define FD_DIFF(a, b, h) (((a) - (b)) * (h))
#endif
        float const c2 = P.c * P.c;
        [[maybe_unused]] float const coeff = -(1.0f / P.eps0) * P.dt;
        bool const full = xa >= xlo && xa + 3 < xhi; // all four cells active: 16-byte accesses
        // The destination values are requested before the wait for the box, so that both travel together -- except in
        // the E update with the current term, where the 24 registers cost a resident CTA (measured at 256^3: B update
        // 175 -> 152 us with the early request, E + J update 197 -> 230 us).
        constexpr bool EARLY = !ADDJ;
        float4 dpre[FD_TZ / 2][3];
#pragma unroll
        for(int zz = 0; zz < FD_TZ; zz += 2)
        {
            int const z = zz + tz;
            if(EARLY && full && yok && Z0 + z < P.g[2] + P.n[2])
            {
                long long const gi = ((long long) (Z0 + z) * P.N[1] + (Y0 + ty)) * P.N[0] + (xa - lead);
#pragma unroll
                for(int c = 0; c < 3; ++c)
                    dpre[zz / 2][c] = *reinterpret_cast<float4 const*>(D.c[c] + gi);
            }
        }
        mbarWait(&bar, 0);
#pragma unroll
        for(int zz = 0; zz < FD_TZ; zz += 2)
        {
            int const z = zz + tz;
            if(!yok || Z0 + z >= P.g[2] + P.n[2] || xa + 3 < xlo || xa >= xhi)
                continue;
            // neighbour offsets inside the box: forward (+1) for the B update, backward (-1) for the E update
            constexpr int SGN = KIND == 0 ? 1 : -1;
            float const* b0 = box + ((z + ZLO) * FD_BY + (ty + 1)) * FD_BX + 4 + 4 * tx;
            float f[3][6], fy[3][4], fz[3][4]; // f: x-1 .. x+4 of the row; fy / fz: the row one step along y / z
#pragma unroll
            for(int c = 0; c < 3; ++c)
            {
                float const* r = b0 + c * FD_COMP;
                float4 const v = *reinterpret_cast<float4 const*>(r);
                f[c][1] = v.x;
                f[c][2] = v.y;
                f[c][3] = v.z;
                f[c][4] = v.w;
                f[c][0] = r[-1];
                f[c][5] = r[4];
                float4 const vy = *reinterpret_cast<float4 const*>(r + SGN * FD_BX);
                float4 const vz = *reinterpret_cast<float4 const*>(r + SGN * FD_PLANE);
                fy[c][0] = vy.x;
                fy[c][1] = vy.y;
                fy[c][2] = vy.z;
                fy[c][3] = vy.w;
                fz[c][0] = vz.x;
                fz[c][1] = vz.y;
                fz[c][2] = vz.z;
                fz[c][3] = vz.w;
            }
            long long const gi = ((long long) (Z0 + z) * P.N[1] + (Y0 + ty)) * P.N[0] + (xa - lead);
            float upd[3][4];
#pragma unroll
            for(int q = 0; q < 4; ++q)
            {
                // d<comp>d<axis>: two-point difference of component comp along axis, divided by the cell size
                float dzdy, dydz, dxdz, dzdx, dydx, dxdy;
                if constexpr(KIND == 0)
                {
                    dzdy = FD_DIFF(fy[2][q], f[2][q + 1], hy);
                    dydz = FD_DIFF(fz[1][q], f[1][q + 1], hz);
                    dxdz = FD_DIFF(fz[0][q], f[0][q + 1], hz);
                    dzdx = FD_DIFF(f[2][q + 2], f[2][q + 1], hx);
                    dydx = FD_DIFF(f[1][q + 2], f[1][q + 1], hx);
                    dxdy = FD_DIFF(fy[0][q], f[0][q + 1], hy);
                }
                else
                {
                    dzdy = FD_DIFF(f[2][q + 1], fy[2][q], hy);
                    dydz = FD_DIFF(f[1][q + 1], fz[1][q], hz);
                    dxdz = FD_DIFF(f[0][q + 1], fz[0][q], hz);
                    dzdx = FD_DIFF(f[2][q + 1], f[2][q], hx);
                    dydx = FD_DIFF(f[1][q + 1], f[1][q], hx);
                    dxdy = FD_DIFF(f[0][q + 1], fy[0][q], hy);
                }
                float const cu[3] = {dzdy - dydz, dxdz - dzdx, dydx - dxdy};
#pragma unroll
                for(int c = 0; c < 3; ++c)
                    upd[c][q] = KIND == 0 ? cu[c] * 0.5f * P.dt : cu[c] * c2 * P.dt;
            }
            if(full)
            {
#pragma unroll
                for(int c = 0; c < 3; ++c)
                {
                    float4* const dp = reinterpret_cast<float4*>(D.c[c] + gi);
                    float4 d;
                    if constexpr(EARLY)
                        d = dpre[zz / 2][c];
                    else
                        d = *dp;
                    if constexpr(KIND == 0)
                    {
                        d.x -= upd[c][0];
                        d.y -= upd[c][1];
                        d.z -= upd[c][2];
                        d.w -= upd[c][3];
                    }
                    else
                    {
                        d.x += upd[c][0];
                        d.y += upd[c][1];
                        d.z += upd[c][2];
                        d.w += upd[c][3];
                        if constexpr(ADDJ)
                        {
                            float4 const j = __ldcs(reinterpret_cast<float4 const*>(J.c[c] + gi));
                            d.x += coeff * j.x;
                            d.y += coeff * j.y;
                            d.z += coeff * j.z;
                            d.w += coeff * j.w;
                        }
                    }
                    *dp = d;
                }
            }
            else
            {
                // brick column that straddles the first / last active cell: cell by cell
#pragma unroll
                for(int q = 0; q < 4; ++q)
                    if(xa + q >= xlo && xa + q < xhi)
#pragma unroll
                        for(int c = 0; c < 3; ++c)
                        {
                            float d = D.c[c][gi + q];
                            if constexpr(KIND == 0)
                                d -= upd[c][q];
                            else
                            {
                                d += upd[c][q];
                                if constexpr(ADDJ)
                                    d += coeff * J.c[c][gi + q];
                            }
                            D.c[c][gi + q] = d;
                        }
            }
        }
    }

    // ---- PML absorber (convolutional PML, Yee curls) ----------------------------------------------------------------
    // Reference: fields/absorber/pml/Pml.kernel:60-160 (relative depth, graded sigma / kappa / alpha, coefficients b, c),
    // :420-476 UpdateEFunctor, :520-582 UpdateBHalfFunctor; hook FDTDBase.hpp:244-298.  One thread per cell of the whole
    // domain (the branch "not in the PML" is the plain Yee update), same float operations as the oracle's restatement.
    struct PmlCoeff
    {
        float kappa[3], b[3], c[3];
        bool inPml;
    };

    __device__ __forceinline__ float pmlRelativeDepth(float cellIdx, float nNeg, float nPos, int numLocalDomainCells, int numGuardCells)
    {
        float const zeroBasedIdx = cellIdx - float(numGuardCells);
        if(zeroBasedIdx < nNeg)
            return (nNeg - zeroBasedIdx) / nNeg;
        float const zeroBasedRightPMLStart = float(numLocalDomainCells - 2 * numGuardCells) - nPos;
        if(zeroBasedIdx > zeroBasedRightPMLStart)
            return (zeroBasedIdx - zeroBasedRightPMLStart) / nPos;
        return 0.0f;
    }

    __device__ __forceinline__ PmlCoeff pmlCoefficients(DevParams const& P, PmlDev const& M, float const idx[3])
    {
        PmlCoeff q;
#pragma unroll
        for(int d = 0; d < 3; ++d)
        {
            float sigma = 0.0f, alpha = 0.0f;
            q.kappa[d] = 1.0f;
            float const depth = pmlRelativeDepth(idx[d], float(M.thickness[d][0]), float(M.thickness[d][1]), P.N[d], P.g[d]);
            if(depth != 0.0f)
            {
                float const sk = powf(depth, M.sigmaKappaGradingOrder);
                sigma = M.sigmaMax[d] * sk;
                q.kappa[d] = 1.0f + (M.kappaMax[d] - 1.0f) * sk;
                float const ag = powf(1.0f - depth, M.alphaGradingOrder);
                alpha = M.alphaMax[d] * ag;
            }
            q.b[d] = expf(-(sigma / q.kappa[d] + alpha) * P.dt);
            q.c[d] = 0.0f;
            float const denominator = q.kappa[d] * (sigma + alpha * q.kappa[d]);
            if(denominator != 0.0f)
                q.c[d] = sigma * (q.b[d] - 1.0f) / denominator;
        }
        float prod = q.b[0] * q.b[1];
        prod = prod * q.b[2];
        q.inPml = prod != 1.0f;
        return q;
    }

    __global__ void __launch_bounds__(256) pmlUpdateEKernel(DevParams P, PmlDev M, Field3 E, Field3 B)
    {
        int const x = blockIdx.x * blockDim.x + threadIdx.x + P.g[0];
        int const y = blockIdx.y * blockDim.y + threadIdx.y + P.g[1];
        int const z = blockIdx.z * blockDim.z + threadIdx.z + P.g[2];
        if(x >= P.g[0] + P.n[0] || y >= P.g[1] + P.n[1] || z >= P.g[2] + P.n[2])
            return;
        long long const sy = P.N[0], sz = (long long) P.N[0] * P.N[1];
        long long const i = fidx(P, x, y, z);
        float const *bx = B.c[0], *by = B.c[1], *bz = B.c[2];
        float const dBzdy = (bz[i] - bz[i - sy]) / P.cell[1], dBydz = (by[i] - by[i - sz]) / P.cell[2];
        float const dBxdz = (bx[i] - bx[i - sz]) / P.cell[2], dBzdx = (bz[i] - bz[i - 1]) / P.cell[0];
        float const dBydx = (by[i] - by[i - 1]) / P.cell[0], dBxdy = (bx[i] - bx[i - sy]) / P.cell[1];
        float const c2 = P.c * P.c;
        float const idx[3] = {float(x), float(y), float(z)};
        PmlCoeff const q = pmlCoefficients(P, M, idx);
        if(q.inPml)
        {
            float const c2dt = c2 * P.dt;
            float* const pyx = M.psi + i;
            float* const pzx = pyx + P.vol;
            float* const pxy = pzx + P.vol;
            float* const pzy = pxy + P.vol;
            float* const pxz = pzy + P.vol;
            float* const pyz = pxz + P.vol;
            float const vyx = q.b[0] * *pyx + q.c[0] * dBzdx, vzx = q.b[0] * *pzx + q.c[0] * dBydx;
            float const vxy = q.b[1] * *pxy + q.c[1] * dBzdy, vzy = q.b[1] * *pzy + q.c[1] * dBxdy;
            float const vxz = q.b[2] * *pxz + q.c[2] * dBydz, vyz = q.b[2] * *pyz + q.c[2] * dBxdz;
            *pyx = vyx;
            *pzx = vzx;
            *pxy = vxy;
            *pzy = vzy;
            *pxz = vxz;
            *pyz = vyz;
            E.c[0][i] += c2dt * (dBzdy / q.kappa[1] - dBydz / q.kappa[2] + vxy - vxz);
            E.c[1][i] += c2dt * (dBxdz / q.kappa[2] - dBzdx / q.kappa[0] + vyz - vyx);
            E.c[2][i] += c2dt * (dBydx / q.kappa[0] - dBxdy / q.kappa[1] + vzx - vzy);
        }
        else
        {
            E.c[0][i] += (dBzdy - dBydz) * c2 * P.dt;
            E.c[1][i] += (dBxdz - dBzdx) * c2 * P.dt;
            E.c[2][i] += (dBydx - dBxdy) * c2 * P.dt;
        }
    }

    __global__ void __launch_bounds__(256) pmlUpdateBHalfKernel(DevParams P, PmlDev M, Field3 E, Field3 B, int updatePsi)
    {
        int const x = blockIdx.x * blockDim.x + threadIdx.x + P.g[0];
        int const y = blockIdx.y * blockDim.y + threadIdx.y + P.g[1];
        int const z = blockIdx.z * blockDim.z + threadIdx.z + P.g[2];
        if(x >= P.g[0] + P.n[0] || y >= P.g[1] + P.n[1] || z >= P.g[2] + P.n[2])
            return;
        long long const sy = P.N[0], sz = (long long) P.N[0] * P.N[1];
        long long const i = fidx(P, x, y, z);
        float const *ex = E.c[0], *ey = E.c[1], *ez = E.c[2];
        float const dEzdy = (ez[i + sy] - ez[i]) / P.cell[1], dEydz = (ey[i + sz] - ey[i]) / P.cell[2];
        float const dExdz = (ex[i + sz] - ex[i]) / P.cell[2], dEzdx = (ez[i + 1] - ez[i]) / P.cell[0];
        float const dEydx = (ey[i + 1] - ey[i]) / P.cell[0], dExdy = (ex[i + sy] - ex[i]) / P.cell[1];
        float const halfDt = 0.5f * P.dt;
        float const idx[3] = {0.5f + float(x), 0.5f + float(y), 0.5f + float(z)};
        PmlCoeff const q = pmlCoefficients(P, M, idx);
        if(q.inPml)
        {
            float* const pyx = M.psi + i;
            float* const pzx = pyx + P.vol;
            float* const pxy = pzx + P.vol;
            float* const pzy = pxy + P.vol;
            float* const pxz = pzy + P.vol;
            float* const pyz = pxz + P.vol;
            float vyx = *pyx, vzx = *pzx, vxy = *pxy, vzy = *pzy, vxz = *pxz, vyz = *pyz;
            if(updatePsi)
            {
                vyx = q.b[0] * vyx + q.c[0] * dEzdx;
                vzx = q.b[0] * vzx + q.c[0] * dEydx;
                vxy = q.b[1] * vxy + q.c[1] * dEzdy;
                vzy = q.b[1] * vzy + q.c[1] * dExdy;
                vxz = q.b[2] * vxz + q.c[2] * dEydz;
                vyz = q.b[2] * vyz + q.c[2] * dExdz;
                *pyx = vyx;
                *pzx = vzx;
                *pxy = vxy;
                *pzy = vzy;
                *pxz = vxz;
                *pyz = vyz;
            }
            B.c[0][i] += halfDt * (dEydz / q.kappa[2] - dEzdy / q.kappa[1] + vxz - vxy);
            B.c[1][i] += halfDt * (dEzdx / q.kappa[0] - dExdz / q.kappa[2] + vyx - vyz);
            B.c[2][i] += halfDt * (dExdy / q.kappa[1] - dEydx / q.kappa[0] + vzy - vzx);
        }
        else
        {
            B.c[0][i] -= (dEzdy - dEydz) * halfDt;
            B.c[1][i] -= (dExdz - dEzdx) * halfDt;
            B.c[2][i] -= (dEydx - dExdy) * halfDt;
        }
    }

    // ---- incident field (laser): PlaneWave profile through the YMin Huygens surface, Yee solver --------------------
    // Reference: fields/incidentField/Solver.hpp:190-395 (updateField: which plane, in-cell shifts, coefficients),
    // Solver.kernel:101-404 (UpdateFunctor; Yee: margin 1, single derivative coefficient 1), Functors.hpp
    // (BaseFunctorE::getCurrentTime, BaseSeparableFunctorE::operator(), ApproximateIncidentB), profiles/PlaneWave.hpp:93-130,
    // profiles/GaussianPulse.hpp:186-346.
    // One thread per cell of the updated plane.  Same float operations in the same order as the oracle's restatement;
    // sin / cos / exp are the device's (1-2 ulp from the host's).
    __device__ __forceinline__ float laserLongitudinal(LaserDev const& L, float time, float phaseShift)
    {
        float envelope = L.amplitude;
        float const mue = 0.5f * L.rampInit * L.pulseDuration;
        float const tau = L.pulseDuration * sqrtf(2.0f);
        float const endUpramp = mue;
        float const startDownramp = mue + L.nofocusConstant;
        float integrationCorrectionFactor = 0.0f;
        if(time > startDownramp)
        {
            float const exponent = (time - startDownramp) / tau;
            envelope *= expf(-0.5f * exponent * exponent);
            integrationCorrectionFactor = (time - startDownramp) / (L.omega * tau * tau);
        }
        else if(time < endUpramp)
        {
            float const exponent = (time - endUpramp) / tau;
            envelope *= expf(-0.5f * exponent * exponent);
            integrationCorrectionFactor = (time - endUpramp) / (L.omega * tau * tau);
        }
        float const timeOszi = time - endUpramp;
        float const phase = L.omega * timeOszi + L.phase + phaseShift;
        return (sinf(phase) + cosf(phase) * integrationCorrectionFactor) * envelope;
    }

    // WavepacketFunctorIncidentE::getLongitudinal (profiles/Wavepacket.hpp:122-151)
    __device__ __forceinline__ float wavepacketLongitudinal(LaserDev const& L, float time, float phaseShift)
    {
        float const endUpramp = -0.5f * L.nofocusConstant, startDownramp = 0.5f * L.nofocusConstant;
        float const mue = 0.5f * L.prm[0];
        float const runTime = time - mue;
        float const tau = L.pulseDuration * sqrtf(2.0f);
        float envelope = L.amplitude;
        float correctionFactor = 0.0f;
        if(runTime > startDownramp)
        {
            float const exponent = ((runTime - startDownramp) / L.pulseDuration / sqrtf(2.0f));
            envelope *= expf(-0.5f * exponent * exponent);
            correctionFactor = (runTime - startDownramp) / (tau * tau * L.omega);
        }
        else if(runTime < endUpramp)
        {
            float const exponent = ((runTime - endUpramp) / L.pulseDuration / sqrtf(2.0f));
            envelope *= expf(-0.5f * exponent * exponent);
            correctionFactor = (runTime - endUpramp) / (tau * tau * L.omega);
        }
        float const phase = L.omega * runTime + L.phase + phaseShift;
        return (sinf(phase) + correctionFactor * cosf(phase)) * envelope;
    }

    // PolynomFunctorIncidentE::getLongitudinal / polynomial (profiles/Polynom.hpp:112-136)
    __device__ __forceinline__ float polynomLongitudinal(LaserDev const& L, float time, float phaseShift)
    {
        float const riseTime = 0.5f * L.pulseDuration;
        float const tau = time / riseTime;
        float const phase = L.omega * (time - riseTime) + L.phase + phaseShift;
        float result = 0.0f;
        if(tau >= 0.0f && tau <= 1.0f)
            result = tau * tau * tau * (10.0f - 15.0f * tau + 6.0f * tau * tau);
        else if(tau > 1.0f && tau <= 2.0f)
            result = (2.0f - tau) * (2.0f - tau) * (2.0f - tau) * (4.0f - 9.0f * tau + 6.0f * tau * tau);
        float const amplitude = L.amplitude * result;
        return sinf(phase) * amplitude;
    }

    // ExpRampWithPrepulseLongitudinal::getEnvelope + ExpRampWithPrepulseFunctorIncidentE::getLongitudinal
    // (profiles/ExpRampWithPrepulse.hpp:157-290)
    __device__ __forceinline__ float expRampGauss(float t, float pulseDuration)
    {
        float const exponent = t / pulseDuration;
        return expf(-0.25f * exponent * exponent);
    }

    __device__ __forceinline__ float expRampExtrapolate(float t1, float a1, float t2, float a2, float t)
    {
        float const log1 = (t2 - t) * logf(a1);
        float const log2 = (t - t1) * logf(a2);
        return expf((log1 + log2) / (t2 - t1));
    }

    __device__ __forceinline__ float expRampLongitudinal(LaserDev const& L, float time, float phaseShift)
    {
        float const* q = L.prm;
        float const time_start_init = q[0], TIME_PREPULSE = q[1], TIME_PEAKPULSE = q[2], TIME_1 = q[3], TIME_2 = q[4], TIME_3 = q[5];
        float const PREPULSE_DURATION = q[6];
        float const endUpramp = TIME_PEAKPULSE - 0.5f * L.nofocusConstant, startDownramp = TIME_PEAKPULSE + 0.5f * L.nofocusConstant;
        float const runTime = time + time_start_init;
        float const phase = L.omega * runTime + L.phase + phaseShift;
        float const AMP_PREPULSE = sqrtf(q[7]), AMP_1 = sqrtf(q[8]), AMP_2 = sqrtf(q[9]), AMP_3 = sqrtf(q[10]);
        float env = 0.0f;
        bool const before_preupramp = runTime < time_start_init;
        bool const before_start = runTime < TIME_1;
        bool const before_peakpulse = runTime < endUpramp;
        bool const during_first_exp = (TIME_1 < runTime) && (runTime < TIME_2);
        bool const after_peakpulse = startDownramp <= runTime;
        if(before_preupramp)
            env = 0.0f;
        else if(before_start)
            env = AMP_1 * expRampGauss(runTime - TIME_1, L.pulseDuration);
        else if(before_peakpulse)
        {
            float const ramp_when_peakpulse = expRampExtrapolate(TIME_2, AMP_2, TIME_3, AMP_3, endUpramp);
            env += (1.0f - ramp_when_peakpulse) * expRampGauss(runTime - endUpramp, L.pulseDuration);
            env += AMP_PREPULSE * expRampGauss(runTime - TIME_PREPULSE, PREPULSE_DURATION);
            if(during_first_exp)
                env += expRampExtrapolate(TIME_1, AMP_1, TIME_2, AMP_2, runTime);
            else
                env += expRampExtrapolate(TIME_2, AMP_2, TIME_3, AMP_3, runTime);
        }
        else if(!after_peakpulse)
            env = 1.0f;
        else
            env = expRampGauss(runTime - startDownramp, L.pulseDuration);
        return cosf(phase) * L.amplitude * env;
    }

    __device__ __forceinline__ float separableLongitudinal(LaserDev const& L, float time, float phaseShift)
    {
        switch(L.profile)
        {
        case 2:
            return wavepacketLongitudinal(L, time, phaseShift);
        case 3:
            return polynomLongitudinal(L, time, phaseShift);
        case 4:
            return expRampLongitudinal(L, time, phaseShift);
        default:
            return laserLongitudinal(L, time, phaseShift);
        }
    }

    __device__ __forceinline__ float dot3(float const a[3], float const b[3])
    {
        float tmp = a[0] * b[0];
        tmp += a[1] * b[1];
        tmp += a[2] * b[2];
        return tmp;
    }

    // GaussianPulseFunctorIncidentE::simpleLaguerre (profiles/GaussianPulse.hpp:316-336)
    __device__ __forceinline__ float simpleLaguerre(unsigned n, float x)
    {
        if(n == 0)
            return 1.0f;
        unsigned currentN = 1;
        float laguerreNMinus1 = 1.0f;
        float laguerreN = 1.0f - x;
        while(currentN < n)
        {
            float const laguerreNPlus1 = ((2.0f * float(currentN) + 1.0f - x) * laguerreN - float(currentN) * laguerreNMinus1) / float(currentN + 1u);
            laguerreNMinus1 = laguerreN;
            laguerreN = laguerreNPlus1;
            currentN++;
        }
        return laguerreN;
    }

    // GaussianPulseFunctorIncidentE::getValue (profiles/GaussianPulse.hpp:208-308) + GaussianPulseEnvelope (:343-348)
    __device__ __forceinline__ float gaussianPulseValue(DevParams const& P, LaserDev const& L, float const posIn[3], float time, float phaseShift)
    {
        float pos[3] = {posIn[0], posIn[1], posIn[2]};
        time += L.timeShift;
        float const focusRelativeToOrigin[3] = {L.focus[0] - L.origin[0], L.focus[1] - L.origin[1], L.focus[2] - L.origin[2]};
        float const axis0[3] = {0.0f, 1.0f, 0.0f};
        float const distanceFocusRelativeToOrigin = dot3(focusRelativeToOrigin, axis0);
        float const focusPos = distanceFocusRelativeToOrigin - pos[0];
        float const w = L.w0 * sqrtf(1.0f + (focusPos / L.rayleighLength) * (focusPos / L.rayleighLength));
        float const phase = L.omega * (time - focusPos / P.c) + L.phase + phaseShift;
        if(L.tilted)
        {
            float const tiltTimeShift = phase / L.omega + focusPos / P.c;
            float const tiltPositionShift = P.c * tiltTimeShift / dot3(axis0, P.cell);
            pos[1] += L.tanTilt[0] * tiltPositionShift;
            pos[2] += L.tanTilt[1] * tiltPositionShift;
        }
        float const q[3] = {pos[0] * 0.0f, pos[1] * 1.0f, pos[2] * 1.0f};
        float transversalDistanceSquared = q[0] * q[0];
        transversalDistanceSquared += q[1] * q[1];
        transversalDistanceSquared += q[2] * q[2];
        float const R_inv = -focusPos / (L.rayleighLength * L.rayleighLength + focusPos * focusPos);
        float const xi = atanf(-focusPos / L.rayleighLength);
        float etrans = 0.0f;
        float const r2OverW2 = transversalDistanceSquared / w / w;
        float const r = 0.5f * transversalDistanceSquared * R_inv;
        float const twoPi = 6.28318530717958647692f;
        for(int m = 0; m < L.nModes; ++m)
            etrans += L.modes[m] * simpleLaguerre(unsigned(m), 2.0f * r2OverW2) * expf(-r2OverW2)
                * cosf(twoPi / L.waveLength * focusPos - twoPi / L.waveLength * r + (2.0f * float(m) + 1.0f) * xi + phase + L.modePhases[m]);
        float const shiftedTime = time - r / P.c;
        float const exponent = shiftedTime / (2.0f * L.pulseDuration);
        etrans *= expf(-exponent * exponent);
        float etrans_norm = 0.0f;
        for(int m = 0; m < L.nModes; ++m)
            etrans_norm += L.modes[m];
        float envelope = L.amplitude;
        envelope *= L.w0 / w;
        return envelope * etrans / etrans_norm;
    }

    // incident E at a fractional total cell index (Functors.hpp: BaseFunctorE::getCurrentTime / getInternalCoordinates)
    __device__ __forceinline__ void laserIncidentE(DevParams const& P, LaserDev const& L, float const idx[3], float out[3])
    {
        float const axis0[3] = {0.0f, 1.0f, 0.0f};
        float const shiftFromOrigin[3] = {idx[0] * P.cell[0] - L.origin[0], idx[1] * P.cell[1] - L.origin[1], idx[2] * P.cell[2] - L.origin[2]};
        float const distance = dot3(shiftFromOrigin, axis0);
        float const timeDelay = distance / L.phaseVelocity + L.timeDelay;
        float const time = L.currentTimeOrigin - timeDelay;
        out[0] = out[1] = out[2] = 0.0f;
        if(time < 0.0f)
            return;
        float a = 0.0f, b;
        if(L.profile != 1)
        {
            // BaseSeparableFunctorE::operator(); transversal: 1 (PlaneWave) or the Gaussian of
            // BaseSeparableTransversalGaussianFunctorE::getTransversal (Functors.hpp:525-532)
            float transversal = 1.0f;
            if(L.profile != 0)
            {
                float const r[3] = {0.0f / 1.0f, dot3(shiftFromOrigin, L.pol) / L.w0Axis[0], dot3(shiftFromOrigin, L.axis2) / L.w0Axis[1]};
                float r2 = r[0] * r[0];
                r2 += r[1] * r[1];
                r2 += r[2] * r[2];
                transversal = expf(-r2);
            }
            if(L.polarisation)
                a = separableLongitudinal(L, time, 1.57079632679489661923f) * transversal;
            b = separableLongitudinal(L, time, 0.0f) * transversal;
        }
        else
        {
            float const pos[3] = {dot3(shiftFromOrigin, axis0), dot3(shiftFromOrigin, L.pol), dot3(shiftFromOrigin, L.axis2)};
            if(L.polarisation)
                a = gaussianPulseValue(P, L, pos, time, 1.57079632679489661923f);
            b = gaussianPulseValue(P, L, pos, time, 0.0f);
        }
        if(L.polarisation == 0)
        {
#pragma unroll
            for(int d = 0; d < 3; ++d)
                out[d] = L.pol[d] * b;
        }
        else
        {
            float const rs2 = sqrtf(2.0f);
            float const p1[3] = {L.pol[0] / rs2, L.pol[1] / rs2, L.pol[2] / rs2};
            float const p2[3] = {1.0f * p1[2] - 0.0f * p1[1], 0.0f * p1[0] - 0.0f * p1[2], 0.0f * p1[1] - 1.0f * p1[0]};
#pragma unroll
            for(int d = 0; d < 3; ++d)
                out[d] = p1[d] * a + p2[d] * b;
        }
    }

    __global__ void __launch_bounds__(256) incidentKernel(DevParams P, Field3 F, LaserDev L)
    {
        int const x = L.lo[0] + blockIdx.x * blockDim.x + threadIdx.x, z = L.lo[1] + blockIdx.y;
        if(x >= L.hi[0] || z >= L.hi[1])
            return;
        bool const lastX = L.lastDomain[0] && x == L.hi[0] - 1, lastZ = L.lastDomain[1] && z == L.hi[1] - 1;
        // Solver.kernel:318-325 with incidentComponent1 = x, incidentComponent2 = z
        bool const apply1 = L.updatedIsE ? !lastZ : !lastX;
        bool const apply2 = L.updatedIsE ? !lastX : !lastZ;
        // in-cell shifts (Solver.hpp:360-372): -1 (E updated) / +1 (B updated) along y plus the Yee position of the
        // incident component: B_inc,x at (0, .5, .5), B_inc,z at (.5, .5, 0); E_inc,x at (.5, 0, 0), E_inc,z at (0, 0, .5)
        float const baseShift = L.updatedIsE ? -1.0f : 1.0f;
        float const base[3] = {float(x), L.planeTotal, float(z)};
        float i1[3], i2[3];
        if(L.updatedIsE)
        {
            i1[0] = base[0] + 0.0f, i1[1] = base[1] + (baseShift + 0.5f), i1[2] = base[2] + 0.5f;
            i2[0] = base[0] + 0.5f, i2[1] = base[1] + (baseShift + 0.5f), i2[2] = base[2] + 0.0f;
        }
        else
        {
            i1[0] = base[0] + 0.5f, i1[1] = base[1] + (baseShift + 0.0f), i1[2] = base[2] + 0.0f;
            i2[0] = base[0] + 0.0f, i2[1] = base[1] + (baseShift + 0.0f), i2[2] = base[2] + 0.5f;
        }
        float e1[3], e2[3];
        laserIncidentE(P, L, i1, e1);
        if(L.profile == 0)
        {
            // the plane wave depends on y only: both evaluations see the same point
            e2[0] = e1[0], e2[1] = e1[1], e2[2] = e1[2];
        }
        else
            laserIncidentE(P, L, i2, e2);
        float inc1, inc2; // incident components x and z
        if(L.updatedIsE)
        {
            // ApproximateIncidentB: cross((0,1,0), E) / c
            inc1 = (1.0f * e1[2] - 0.0f * e1[1]) / P.c;
            inc2 = (0.0f * e2[1] - 1.0f * e2[0]) / P.c;
        }
        else
        {
            inc1 = e1[0];
            inc2 = e2[2];
        }
        float rz = 0.0f, rx = 0.0f;
        if(apply1)
            rz += 1.0f * inc1;
        if(apply2)
            rx += 1.0f * inc2;
        rz *= L.baseCoefficient;
        rx *= -L.baseCoefficient;
        long long const i = fidx(P, x + P.g[0], L.plane, z + P.g[2]);
        F.c[2][i] += rz;
        F.c[0][i] += rx;
    }

    // KernelAddCurrentDensity + None: E += (-(1/eps0) * dt) * J   (FDTD.hpp:84-85)
    __global__ void __launch_bounds__(256) addCurrentKernel(DevParams P, Field3 E, Field3 J)
    {
        int const x = blockIdx.x * blockDim.x + threadIdx.x;
        int const y = blockIdx.y * blockDim.y + threadIdx.y;
        int const z = blockIdx.z * blockDim.z + threadIdx.z;
        if(x >= P.n[0] || y >= P.n[1] || z >= P.n[2])
            return;
        long long const i = fidx(P, x + P.g[0], y + P.g[1], z + P.g[2]);
        float const coeff = -(1.0f / P.eps0) * P.dt;
#pragma unroll
        for(int c = 0; c < 3; ++c)
            E.c[c][i] += coeff * J.c[c][i];
    }

    // KernelAddCurrentDensity + Binomial: E += coeff * (1-2-1 filter of J in x, y and z) with the reference's summation
    // order: corners (T=1), edges (D=2), faces (S=4), centre (M=8), times 1/64 (Binomial.hpp:62-110).  J guards hold
    // the neighbours' border values (second, "receive" exchange of FieldJ::asyncCommunication, FieldJ.x.cpp:118-141).
    __global__ void __launch_bounds__(256) addCurrentBinomialKernel(DevParams P, Field3 E, Field3 J)
    {
        int const x = blockIdx.x * blockDim.x + threadIdx.x;
        int const y = blockIdx.y * blockDim.y + threadIdx.y;
        int const z = blockIdx.z * blockDim.z + threadIdx.z;
        if(x >= P.n[0] || y >= P.n[1] || z >= P.n[2])
            return;
        long long const i = fidx(P, x + P.g[0], y + P.g[1], z + P.g[2]);
        long long const sy = P.N[0], sz = (long long) P.N[0] * P.N[1];
        float const coeff = -(1.0f / P.eps0) * P.dt;
#pragma unroll
        for(int c = 0; c < 3; ++c)
        {
            float const* __restrict__ j = J.c[c] + i;
            auto at = [&](int dx, int dy, int dz) { return j[dx + dy * sy + dz * sz]; };
            float t = at(-1, -1, -1) + at(+1, -1, -1);
            t += at(-1, +1, -1);
            t += at(+1, +1, -1);
            t += at(-1, -1, +1);
            t += at(+1, -1, +1);
            t += at(-1, +1, +1);
            t += at(+1, +1, +1);
            float d = at(-1, -1, 0) + at(+1, -1, 0);
            d += at(-1, +1, 0);
            d += at(+1, +1, 0);
            d += at(-1, 0, -1);
            d += at(+1, 0, -1);
            d += at(-1, 0, +1);
            d += at(+1, 0, +1);
            d += at(0, -1, -1);
            d += at(0, +1, -1);
            d += at(0, -1, +1);
            d += at(0, +1, +1);
            float f = at(-1, 0, 0) + at(+1, 0, 0);
            f += at(0, -1, 0);
            f += at(0, +1, 0);
            f += at(0, 0, -1);
            f += at(0, 0, +1);
            float avg = 1.0f * t + 2.0f * d;
            avg = avg + 4.0f * f;
            avg = avg + 8.0f * at(0, 0, 0);
            avg *= 1.0f / 64.0f;
            E.c[c][i] += coeff * avg;
        }
    }

    // exponential::KernelAbsorbBorder (Exponential.kernel:45-118) for all six faces in the order of Exponential.hpp:70-108
    // (x+, x-, y+, y-, z+, z-): every active cell closer than `cells` to an absorbing face is multiplied by
    // exp(-strength * factor), factor = cells-1 ... 1 from the face inwards.  The attenuation factors are tabulated on
    // the host with the same libm call the CPU reference makes: damp[(axis*2+side)*ABS_MAX + factor].
    __global__ void __launch_bounds__(256) absorbKernel(DevParams P, Field3 F, AbsorberDev A)
    {
        int const x = blockIdx.x * blockDim.x + threadIdx.x;
        int const y = blockIdx.y * blockDim.y + threadIdx.y;
        int const z = blockIdx.z * blockDim.z + threadIdx.z;
        if(x >= P.n[0] || y >= P.n[1] || z >= P.n[2])
            return;
        int const q[3] = {x, y, z};
        int fac[6];
        bool any = false;
#pragma unroll
        for(int a = 0; a < 3; ++a)
        {
            // positive side first (exchange types RIGHT, BOTTOM, BACK are odd), then negative
            int const fp = A.cells[a][1] > 0 ? q[a] - P.n[a] + A.cells[a][1] : 0;
            int const fn = A.cells[a][0] > 0 ? A.cells[a][0] - 1 - q[a] : 0;
            fac[2 * a] = fp > 0 ? fp : 0;
            fac[2 * a + 1] = fn > 0 ? fn : 0;
            any = any || fp > 0 || fn > 0;
        }
        if(!any)
            return;
        long long const i = fidx(P, x + P.g[0], y + P.g[1], z + P.g[2]);
#pragma unroll
        for(int c = 0; c < 3; ++c)
        {
            float v = F.c[c][i];
#pragma unroll
            for(int a = 0; a < 3; ++a)
            {
                if(fac[2 * a])
                    v = v * A.damp[(2 * a + 1) * ABS_MAX + fac[2 * a]];
                if(fac[2 * a + 1])
                    v = v * A.damp[(2 * a + 0) * ABS_MAX + fac[2 * a + 1]];
            }
            F.c[c][i] = v;
        }
    }

    // ---- guard exchange ------------------------------------------------------------------------------------------
    // One axis at a time (x, then y, then z for copies; the same order for the J reduction), every pass spanning the
    // full padded extent of the two other axes, so the 26 directions of the reference collapse into 3 passes.
    // slab geometry: `width` planes starting at plane `start` along `axis`.
    __device__ __forceinline__ long long slabIndex(DevParams const& P, int axis, int plane, int u, int v)
    {
        // (u,v) run over the two other axes in ascending axis order
        int c[3];
        c[axis] = plane;
        c[(axis == 0) ? 1 : 0] = u;
        c[(axis == 2) ? 1 : 2] = v;
        return fidx(P, c[0], c[1], c[2]);
    }

    __device__ __forceinline__ bool slabInRange(DevParams const& P, int axis, int u, int v)
    {
        int const au = (axis == 0) ? 1 : 0, av = (axis == 2) ? 1 : 2;
        return u >= P.tlo[au] && u < P.thi[au] && v >= P.tlo[av] && v < P.thi[av];
    }

    // local periodic wrap: dst planes <- (or +=) src planes, all three components
    template<bool ADD>
    __global__ void __launch_bounds__(256) haloLocalKernel(DevParams P, Field3 F, int ncomp, int axis, int srcStart, int dstStart, int width)
    {
        int const U = P.N[(axis == 0) ? 1 : 0], V = P.N[(axis == 2) ? 1 : 2];
        long long const per = (long long) U * V * width;
        long long const total = per * ncomp;
        for(long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long) gridDim.x * blockDim.x)
        {
            int const comp = int(t / per);
            long long r = t % per;
            int u, v, w;
            if(axis == 0)
            {
                w = int(r % width);
                r /= width;
                u = int(r % U);
                v = int(r / U);
            }
            else
            {
                u = int(r % U);
                r /= U;
                if(axis == 1)
                {
                    w = int(r % width);
                    v = int(r / width);
                }
                else
                {
                    v = int(r % V);
                    w = int(r / V);
                }
            }
            if(!slabInRange(P, axis, u, v))
                continue;
            long long const s = slabIndex(P, axis, srcStart + w, u, v), d = slabIndex(P, axis, dstStart + w, u, v);
            if(ADD)
                F.c[comp][d] += F.c[comp][s];
            else
                F.c[comp][d] = F.c[comp][s];
        }
    }

    // pack `width` planes into a contiguous buffer [comp][w][v][u] / unpack (copy or add)
    __global__ void __launch_bounds__(256) haloPackKernel(DevParams P, Field3 F, int ncomp, int axis, int start, int width, float* __restrict__ buf)
    {
        int const U = P.N[(axis == 0) ? 1 : 0], V = P.N[(axis == 2) ? 1 : 2];
        long long const per = (long long) U * V * width;
        long long const total = per * ncomp;
        for(long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long) gridDim.x * blockDim.x)
        {
            int const comp = int(t / per);
            long long r = t % per;
            int const u = int(r % U);
            r /= U;
            int const v = int(r % V);
            int const w = int(r / V);
            buf[t] = slabInRange(P, axis, u, v) ? F.c[comp][slabIndex(P, axis, start + w, u, v)] : 0.0f;
        }
    }

    template<bool ADD>
    __global__ void __launch_bounds__(256) haloUnpackKernel(DevParams P, Field3 F, int ncomp, int axis, int start, int width, float const* __restrict__ buf)
    {
        int const U = P.N[(axis == 0) ? 1 : 0], V = P.N[(axis == 2) ? 1 : 2];
        long long const per = (long long) U * V * width;
        long long const total = per * ncomp;
        for(long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long) gridDim.x * blockDim.x)
        {
            int const comp = int(t / per);
            long long r = t % per;
            int const u = int(r % U);
            r /= U;
            int const v = int(r % V);
            int const w = int(r / V);
            if(!slabInRange(P, axis, u, v))
                continue;
            long long const d = slabIndex(P, axis, start + w, u, v);
            if(ADD)
                F.c[comp][d] += buf[t];
            else
                F.c[comp][d] = buf[t];
        }
    }

    // ---- layout conversion (reference AoS float3 <-> SoA planes) -------------------------------------------------
    __global__ void __launch_bounds__(256) aosToSoaKernel(float const* __restrict__ aos, Field3 F, long long vol)
    {
        for(long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x; t < 3 * vol; t += (long long) gridDim.x * blockDim.x)
            F.c[t % 3][t / 3] = aos[t];
    }
    __global__ void __launch_bounds__(256) soaToAosKernel(Field3 F, float* __restrict__ aos, long long vol)
    {
        for(long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x; t < 3 * vol; t += (long long) gridDim.x * blockDim.x)
            aos[t] = F.c[t % 3][t / 3];
    }

    // ---- reductions ----------------------------------------------------------------------------------------------
    __device__ __forceinline__ double blockSum(double v)
    {
        __shared__ double ws[8];
        __syncthreads();
#pragma unroll
        for(int o = 16; o > 0; o >>= 1)
            v += __shfl_xor_sync(0xffffffffu, v, o);
        int const tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
        if((tid & 31) == 0)
            ws[tid >> 5] = v;
        __syncthreads();
        double s = 0;
        if(tid == 0)
            for(int i = 0; i < 8; ++i)
                s += ws[i];
        return s;
    }

    // EnergyFields.x.cpp:198-233: sum of squares over CORE+BORDER, out[0] += B^2, out[1] += E^2
    __global__ void __launch_bounds__(256) fieldEnergyKernel(DevParams P, Field3 E, Field3 B, double* __restrict__ out)
    {
        long long const ncell = (long long) P.n[0] * P.n[1] * P.n[2];
        double sB = 0, sE = 0;
        for(long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x; t < ncell; t += (long long) gridDim.x * blockDim.x)
        {
            int const x = int(t % P.n[0]), y = int((t / P.n[0]) % P.n[1]), z = int(t / ((long long) P.n[0] * P.n[1]));
            long long const i = fidx(P, x + P.g[0], y + P.g[1], z + P.g[2]);
#pragma unroll
            for(int c = 0; c < 3; ++c)
            {
                sB += double(B.c[c][i]) * double(B.c[c][i]);
                sE += double(E.c[c][i]) * double(E.c[c][i]);
            }
        }
        sB = blockSum(sB);
        sE = blockSum(sE);
        if(threadIdx.x == 0)
        {
            atomicAdd(&out[0], sB);
            atomicAdd(&out[1], sE);
        }
    }

    // EnergyParticles.x.cpp:100-131 + KinEnergy.hpp:38-68
    __global__ void __launch_bounds__(256) particleEnergyKernel(DevParams P, SpeciesDev S, uint32_t const* __restrict__ nPart, uint32_t const* __restrict__ inv, double* __restrict__ out)
    {
        uint32_t const n = *nPart;
        double ek = 0, et = 0;
        float const c2 = P.c * P.c;
        for(uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x)
        {
            // lazily re-sorted species: slot j of the frame runs lives at inv[j] (the sum does not care about the order,
            // but the buffer also holds the slots of particles that have left)
            uint32_t const i = inv ? inv[j] : j;
            float const ux = S.mom[0][i], uy = S.mom[1][i], uz = S.mom[2][i];
            float m2 = ux * ux;
            m2 += uy * uy;
            m2 += uz * uz;
            float const mass = S.mass_per_w * S.w[i];
            float const gamma = sqrtf(1.0f + m2 * (1.0f / (mass * mass * c2)));
            float kin;
            if(gamma < 1.005f) // GAMMA_THRESH, param/speciesConstants.param:39
                kin = m2 / (2.0f * mass);
            else
                kin = (gamma - 1.0f) * mass * c2;
            ek += double(kin);
            et += double(sqrtf(m2 + mass * mass * c2) * P.c);
        }
        ek = blockSum(ek);
        et = blockSum(et);
        if(threadIdx.x == 0)
        {
            atomicAdd(&out[0], ek);
            atomicAdd(&out[1], et);
        }
    }

    // ChargeDensity into a scalar grid (particleToGrid/ComputeGridValuePerFrame.hpp:60-134)
    template<int SHAPE>
    __global__ void __launch_bounds__(256) chargeDensityKernel(DevParams P, SpeciesDev S, uint32_t const* __restrict__ cellOff, float* __restrict__ rho)
    {
        using Sh = Shape<SHAPE>;
        constexpr int lo = Sh::SUPP / 2, up = (Sh::SUPP + 1) / 2;
        int const sc = blockIdx.x;
        int const scx = sc % P.nsc[0], scy = (sc / P.nsc[0]) % P.nsc[1], scz = sc / (P.nsc[0] * P.nsc[1]);
        uint32_t const p0 = cellOff[sc * SCVOL], p1 = cellOff[(sc + 1) * SCVOL];
        float const V = P.cell[0] * P.cell[1] * P.cell[2];
        for(uint32_t i = p0 + threadIdx.x; i < p1; i += blockDim.x)
        {
            int const lc = S.cell[i];
            int const cx = scx * SCX + lc % SCX + P.g[0], cy = scy * SCY + (lc / SCX) % SCY + P.g[1], cz = scz * SCZ + lc / (SCX * SCY) + P.g[2];
            float const px = S.pos[0][i], py = S.pos[1][i], pz = S.pos[2][i];
            float const attr = (S.charge_per_w * S.w[i]) / V;
            for(int oz = -lo; oz <= up; ++oz)
                for(int oy = -lo; oy <= up; ++oy)
                    for(int ox = -lo; ox <= up; ++ox)
                    {
                        float a = 1.0f;
                        a *= shapeEval<SHAPE>(float(ox) - px);
                        a *= shapeEval<SHAPE>(float(oy) - py);
                        a *= shapeEval<SHAPE>(float(oz) - pz);
                        atomicAdd(&rho[fidx(P, cx + ox, cy + oy, cz + oz)], a * attr);
                    }
        }
    }

    // ChargeConservation.tpp:122-136,205-259: max |div E * eps0 - rho| over CORE+BORDER (non-negative floats order
    // like their bit patterns, so atomicMax on the int view is exact)
    __global__ void __launch_bounds__(256) gaussResidualKernel(DevParams P, Field3 E, float const* __restrict__ rho, int* __restrict__ outMax)
    {
        long long const ncell = (long long) P.n[0] * P.n[1] * P.n[2];
        float const rw = 1.0f / P.cell[0], rh = 1.0f / P.cell[1], rd = 1.0f / P.cell[2];
        float mx = 0.0f;
        long long const sy = P.N[0], sz = (long long) P.N[0] * P.N[1];
        for(long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x; t < ncell; t += (long long) gridDim.x * blockDim.x)
        {
            int const x = int(t % P.n[0]), y = int((t / P.n[0]) % P.n[1]), z = int(t / ((long long) P.n[0] * P.n[1]));
            long long const i = fidx(P, x + P.g[0], y + P.g[1], z + P.g[2]);
            float const div = (E.c[0][i] - E.c[0][i - 1]) * rw + (E.c[1][i] - E.c[1][i - sy]) * rh + (E.c[2][i] - E.c[2][i - sz]) * rd;
            mx = fmaxf(mx, fabsf(div * P.eps0 - rho[i]));
        }
#pragma unroll
        for(int o = 16; o > 0; o >>= 1)
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if((threadIdx.x & 31) == 0)
            atomicMax(outMax, __float_as_int(mx));
    }

    // ---- launchers -----------------------------------------------------------------------------------------------
    static inline dim3 cellGrid(DevParams const& P, dim3 b)
    {
        return dim3((P.n[0] + b.x - 1) / b.x, (P.n[1] + b.y - 1) / b.y, (P.n[2] + b.z - 1) / b.z);
    }

    cudaError_t launchPmlUpdateE(DevParams const& P, PmlDev const& M, Field3 E, Field3 B, cudaStream_t st)
    {
        dim3 const b(32, 4, 2);
        pmlUpdateEKernel<<<dim3((P.n[0] + b.x - 1) / b.x, (P.n[1] + b.y - 1) / b.y, (P.n[2] + b.z - 1) / b.z), b, 0, st>>>(P, M, E, B);
        return cudaGetLastError();
    }

    cudaError_t launchPmlUpdateBHalf(DevParams const& P, PmlDev const& M, Field3 E, Field3 B, bool updatePsi, cudaStream_t st)
    {
        dim3 const b(32, 4, 2);
        pmlUpdateBHalfKernel<<<dim3((P.n[0] + b.x - 1) / b.x, (P.n[1] + b.y - 1) / b.y, (P.n[2] + b.z - 1) / b.z), b, 0, st>>>(P, M, E, B, updatePsi ? 1 : 0);
        return cudaGetLastError();
    }

    cudaError_t launchIncident(DevParams const& P, Field3 F, LaserDev const& L, cudaStream_t st)
    {
        dim3 const grid((L.hi[0] - L.lo[0] + 255) / 256, L.hi[1] - L.lo[1]);
        incidentKernel<<<grid, 256, 0, st>>>(P, F, L);
        return cudaGetLastError();
    }

    void fdtdBox(int box[3])
    {
        box[0] = FD_BX;
        box[1] = FD_BY;
        box[2] = FD_BZ;
    }

    /** TMA-staged Yee update: kind 0 B -= curl E dt/2 (src = map of E, dst = B), kind 1 E += curl B c^2 dt (src = map of B, dst = E), addJ: + coeff J */
    cudaError_t launchFdtdTma(int kind, bool addJ, DevParams const& P, Field3 dst, Field3 J, CUtensorMap const& srcMap, int lead, cudaStream_t st)
    {
        int const x0 = (P.g[0] + lead) & ~3;
        dim3 const grid((P.g[0] + lead + P.n[0] - x0 + FD_TX - 1) / FD_TX, (P.n[1] + FD_TY - 1) / FD_TY, (P.n[2] + FD_TZ - 1) / FD_TZ);
        size_t const smem = FD_BYTES;
        cudaError_t e = cudaSuccess;
        if(kind == 0)
        {
            e = cudaFuncSetAttribute(fdtdTmaKernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
            if(e == cudaSuccess)
                fdtdTmaKernel<0, false><<<grid, 256, smem, st>>>(P, dst, J, srcMap, lead);
        }
        else if(addJ)
        {
            e = cudaFuncSetAttribute(fdtdTmaKernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
            if(e == cudaSuccess)
                fdtdTmaKernel<1, true><<<grid, 256, smem, st>>>(P, dst, J, srcMap, lead);
        }
        else
        {
            e = cudaFuncSetAttribute(fdtdTmaKernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
            if(e == cudaSuccess)
                fdtdTmaKernel<1, false><<<grid, 256, smem, st>>>(P, dst, J, srcMap, lead);
        }
        return e != cudaSuccess ? e : cudaGetLastError();
    }

    cudaError_t launchUpdateBHalf(int solver, DevParams const& P, LeheCoeffs const& L, Field3 E, Field3 B, cudaStream_t st)
    {
        dim3 const b(32, 4, 2);
        if(solver == 1)
            updateBHalfKernel<2><<<cellGrid(P, b), b, 0, st>>>(P, L, E, B);
        else
            updateBHalfKernel<0><<<cellGrid(P, b), b, 0, st>>>(P, L, E, B);
        return cudaGetLastError();
    }

    cudaError_t launchUpdateE(DevParams const& P, LeheCoeffs const& L, Field3 E, Field3 B, cudaStream_t st)
    {
        dim3 const b(32, 4, 2);
        updateEKernel<<<cellGrid(P, b), b, 0, st>>>(P, L, E, B);
        return cudaGetLastError();
    }

    cudaError_t launchAddCurrent(DevParams const& P, Field3 E, Field3 J, bool binomial, cudaStream_t st)
    {
        dim3 const b(32, 4, 2);
        if(binomial)
            addCurrentBinomialKernel<<<cellGrid(P, b), b, 0, st>>>(P, E, J);
        else
            addCurrentKernel<<<cellGrid(P, b), b, 0, st>>>(P, E, J);
        return cudaGetLastError();
    }

    cudaError_t launchAbsorb(DevParams const& P, Field3 F, AbsorberDev const& A, cudaStream_t st)
    {
        dim3 const b(32, 4, 2);
        absorbKernel<<<cellGrid(P, b), b, 0, st>>>(P, F, A);
        return cudaGetLastError();
    }

    static inline int slabGrid(DevParams const& P, int axis, int width, int ncomp)
    {
        long long const per = (long long) P.N[(axis == 0) ? 1 : 0] * P.N[(axis == 2) ? 1 : 2] * width * ncomp;
        long long b = (per + 255) / 256;
        if(b > 148 * 8)
            b = 148 * 8;
        if(b < 1)
            b = 1;
        return int(b);
    }

    cudaError_t launchHaloLocal(bool add, DevParams const& P, Field3 F, int ncomp, int axis, int srcStart, int dstStart, int width, cudaStream_t st)
    {
        if(width <= 0)
            return cudaSuccess;
        if(add)
            haloLocalKernel<true><<<slabGrid(P, axis, width, ncomp), 256, 0, st>>>(P, F, ncomp, axis, srcStart, dstStart, width);
        else
            haloLocalKernel<false><<<slabGrid(P, axis, width, ncomp), 256, 0, st>>>(P, F, ncomp, axis, srcStart, dstStart, width);
        return cudaGetLastError();
    }

    cudaError_t launchHaloPack(DevParams const& P, Field3 F, int ncomp, int axis, int start, int width, float* buf, cudaStream_t st)
    {
        if(width <= 0)
            return cudaSuccess;
        haloPackKernel<<<slabGrid(P, axis, width, ncomp), 256, 0, st>>>(P, F, ncomp, axis, start, width, buf);
        return cudaGetLastError();
    }

    cudaError_t launchHaloUnpack(bool add, DevParams const& P, Field3 F, int ncomp, int axis, int start, int width, float const* buf, cudaStream_t st)
    {
        if(width <= 0)
            return cudaSuccess;
        if(add)
            haloUnpackKernel<true><<<slabGrid(P, axis, width, ncomp), 256, 0, st>>>(P, F, ncomp, axis, start, width, buf);
        else
            haloUnpackKernel<false><<<slabGrid(P, axis, width, ncomp), 256, 0, st>>>(P, F, ncomp, axis, start, width, buf);
        return cudaGetLastError();
    }

    cudaError_t launchAosToSoa(float const* aos, Field3 F, long long vol, cudaStream_t st)
    {
        aosToSoaKernel<<<148 * 8, 256, 0, st>>>(aos, F, vol);
        return cudaGetLastError();
    }
    cudaError_t launchSoaToAos(Field3 F, float* aos, long long vol, cudaStream_t st)
    {
        soaToAosKernel<<<148 * 8, 256, 0, st>>>(F, aos, vol);
        return cudaGetLastError();
    }

    cudaError_t launchFieldEnergy(DevParams const& P, Field3 E, Field3 B, double* out, cudaStream_t st)
    {
        fieldEnergyKernel<<<148 * 4, 256, 0, st>>>(P, E, B, out);
        return cudaGetLastError();
    }

    cudaError_t launchParticleEnergy(DevParams const& P, SpeciesDev S, uint32_t const* nPart, uint32_t const* inv, double* out, cudaStream_t st)
    {
        particleEnergyKernel<<<148 * 8, 256, 0, st>>>(P, S, nPart, inv, out);
        return cudaGetLastError();
    }

    cudaError_t launchChargeDensity(int shape, DevParams const& P, SpeciesDev S, uint32_t const* cellOff, float* rho, cudaStream_t st)
    {
        int const nscTot = P.nsc[0] * P.nsc[1] * P.nsc[2];
        switch(shape)
        {
        case 0:
            chargeDensityKernel<0><<<nscTot, 256, 0, st>>>(P, S, cellOff, rho);
            break;
        case 1:
            chargeDensityKernel<1><<<nscTot, 256, 0, st>>>(P, S, cellOff, rho);
            break;
        case 2:
            chargeDensityKernel<2><<<nscTot, 256, 0, st>>>(P, S, cellOff, rho);
            break;
        case 3:
            chargeDensityKernel<3><<<nscTot, 256, 0, st>>>(P, S, cellOff, rho);
            break;
        default:
            chargeDensityKernel<4><<<nscTot, 256, 0, st>>>(P, S, cellOff, rho);
            break;
        }
        return cudaGetLastError();
    }

    cudaError_t launchGaussResidual(DevParams const& P, Field3 E, float const* rho, int* outMax, cudaStream_t st)
    {
        gaussResidualKernel<<<148 * 4, 256, 0, st>>>(P, E, rho, outMax);
        return cudaGetLastError();
    }
} // namespace picstep
