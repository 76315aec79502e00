// Matrix-free cell operator, variant 1 ("general cell kernel"): one cell per n^(d-1) threads,
// register-tiled 1-D contractions with a shared-memory transpose between directions, scatter by
// FP64 atomics.  This is the simple, shape-agnostic kernel (any level, any cell count); the
// fused-epilogue fast path lives in op_v2.cuh and is validated against this one.
//
// Replaces the deal.II cell loop of the reference:
//   MassLaplaceOperatorMatrixFree::do_cell_integral_range   operator.h:379-421
//   ComplexMassLaplaceOperatorMatrixFree::vmult (fused)      operator.h:616-665
//   BatchedMassLaplaceOperatorMatrixFree::do_cell_integral   operator.h:841-880
//
// Math (Cartesian cell of side h, SURVEY 7 hard part 1): with the reference 1-D matrices Mh, Kh
//   A_cell = cm Mz My Mx + cl (Mz My Kx + Mz Ky Mx + Kz My Mx),  cm = mass h^d, cl = laplace h^(d-2)
// evaluated in 7 one-dimensional sweeps:
//   a = Mx u, b = Kx u;   p = My a, q = Ky a, r = My b;   out = Mz (cm p + cl (q + r)) + cl Kz p.
// COUPLED operators (dst_i = cl_i K u_i + M sum_j C_ij u_j) add one sweep per direction for the
// mixed mass input um_i = sum_j C_ij u_j.
#pragma once
#include "common.cuh"

namespace spirk
{
  // one copy per translation unit (the plane-streaming kernels are compiled in separate units, see build.py);
  // upload_fe_constants() fills all of them
  static __constant__ FeConst c_fe[SPIRK_MAX_DEGREE + 1];

  __device__ __forceinline__ bool on_bdry(const int i, const int n1) { return i == 0 || i == n1 - 1; }

  template <int K>
  struct CfgV1
  {
    static constexpr int n = K + 1, n2 = n * n, n3 = n2 * n;
    // cells per block, 3-D (n^2 threads per cell) and 2-D (n threads per cell)
    static constexpr int CPB3 = (K == 1) ? 32 : (K == 2) ? 14 : (K == 3) ? 8 : (K == 4) ? 5 : (K == 5) ? 4 : 3;
    static constexpr int CPB2 = 128 / n;
    static constexpr int T3   = CPB3 * n2;
    static constexpr int T2   = CPB2 * n;
  };

  // out[i] = sum_j A[i*n+j] in[j]   (A in constant memory, fully unrolled)
  template <int n>
  __device__ __forceinline__ void matvec(const double *__restrict__ A, const double (&in)[n], double (&out)[n])
  {
#pragma unroll
    for (int i = 0; i < n; ++i)
      {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < n; ++j)
          s = fma(A[i * n + j], in[j], s);
        out[i] = s;
      }
  }

  // dst_b = boundary ? src_b : 0   (replaces cell_loop's zero_dst + the constrained-DoF identity,
  // operator.h:301-309)
  static __global__ void k_init_dst(const Geo g, const int nb, double *__restrict__ dst, const double *__restrict__ src,
                             const long long stride)
  {
    const long long total = g.N * nb;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x)
      {
        const int       b  = e / g.N;
        const long long i  = e - b * g.N;
        const int       ix = i % g.n1, iy = (i / g.n1) % g.n1, iz = (g.dim == 3) ? i / ((long long)g.n1 * g.n1) : 1;
        const bool      bd = on_bdry(ix, g.n1) || on_bdry(iy, g.n1) || (g.dim == 3 && on_bdry(iz, g.n1));
        dst[b * stride + i] = bd ? src[b * stride + i] : 0.0;
      }
  }

  template <int K, bool COUPLED>
  __global__ void __launch_bounds__(CfgV1<K>::T3)
    k_apply3d_v1(const Geo g, const OpDev op, double *__restrict__ dst, const double *__restrict__ src, const long long stride)
  {
    using C             = CfgV1<K>;
    constexpr int n = C::n, n2 = C::n2, n3 = C::n3, CPB = C::CPB3;
    __shared__ double sA[CPB * n3], sB[CPB * n3], sC[CPB * n3], sD[CPB * n3];
    __shared__ double sE[COUPLED ? CPB * n3 : 1];
    const double *Mh = c_fe[K].Mh, *Kh = c_fe[K].Kh;

    const int       t = threadIdx.x, lc = t / n2, w = t % n2, w0 = w % n, w1 = w / n;
    const int       nc = g.nc, n1 = g.n1;
    const long long ncells = (long long)nc * nc * nc, nbatches = (ncells + CPB - 1) / CPB;
    double         *myA = sA + lc * n3, *myB = sB + lc * n3, *myC = sC + lc * n3, *myD = sD + lc * n3;
    double         *myE = sE + (COUPLED ? lc * n3 : 0);

    for (long long batch = blockIdx.x; batch < nbatches; batch += gridDim.x)
      {
        const long long c      = batch * CPB + lc;
        const bool      active = c < ncells;
        const int       cx = active ? (int)(c % nc) : 0, cy = active ? (int)((c / nc) % nc) : 0,
                  cz = active ? (int)(c / ((long long)nc * nc)) : 0;
        // phase-1 ownership: x-line (y = w0, z = w1)
        const int       gy1 = cy * K + w0, gz1 = cz * K + w1;
        const bool      line_bd = on_bdry(gy1, n1) || on_bdry(gz1, n1);
        const long long base1   = (long long)cx * K + (long long)n1 * (gy1 + (long long)n1 * gz1);

        for (int b = 0; b < op.nb; ++b)
          {
            // ---------------- phase 1: gather x-line, contract along x
            double u[n], um[n];
#pragma unroll
            for (int i = 0; i < n; ++i)
              {
                const bool bd = line_bd || (i == 0 && cx == 0) || (i == K && cx == nc - 1);
                u[i]          = (active && !bd) ? src[b * stride + base1 + i] : 0.0;
                um[i]         = 0.0;
              }
            if (COUPLED)
              {
                for (int j = 0; j < op.nb; ++j)
                  {
                    const double cij = op.cc[b * op.nb + j];
                    if (cij != 0.0)
                      {
#pragma unroll
                        for (int i = 0; i < n; ++i)
                          {
                            const bool bd = line_bd || (i == 0 && cx == 0) || (i == K && cx == nc - 1);
                            um[i] = fma(cij, (active && !bd) ? src[j * stride + base1 + i] : 0.0, um[i]);
                          }
                      }
                  }
              }
            {
              double a[n], bk[n];
              matvec<n>(Mh, u, a);
              matvec<n>(Kh, u, bk);
#pragma unroll
              for (int i = 0; i < n; ++i)
                {
                  myA[(w1 * n + w0) * n + i] = a[i];
                  myB[(w1 * n + w0) * n + i] = bk[i];
                }
              if (COUPLED)
                {
                  double am[n];
                  matvec<n>(Mh, um, am);
#pragma unroll
                  for (int i = 0; i < n; ++i)
                    myE[(w1 * n + w0) * n + i] = am[i];
                }
            }
            __syncthreads();
            // ---------------- phase 2: y-line (x = w0, z = w1)
            {
              double a[n], bk[n], p[n], q[n], r[n];
#pragma unroll
              for (int j = 0; j < n; ++j)
                {
                  a[j]  = myA[(w1 * n + j) * n + w0];
                  bk[j] = myB[(w1 * n + j) * n + w0];
                }
              matvec<n>(Mh, a, p);
              matvec<n>(Kh, a, q);
              matvec<n>(Mh, bk, r);
              const double cl = op.cl[b];
              if (COUPLED)
                {
                  double am[n], pm[n];
#pragma unroll
                  for (int j = 0; j < n; ++j)
                    am[j] = myE[(w1 * n + j) * n + w0];
                  matvec<n>(Mh, am, pm);
#pragma unroll
                  for (int j = 0; j < n; ++j)
                    myC[(w1 * n + j) * n + w0] = fma(cl, q[j] + r[j], pm[j]);
                }
              else
                {
                  const double cm = op.cm[b];
#pragma unroll
                  for (int j = 0; j < n; ++j)
                    myC[(w1 * n + j) * n + w0] = fma(cm, p[j], cl * (q[j] + r[j]));
                }
#pragma unroll
              for (int j = 0; j < n; ++j)
                myD[(w1 * n + j) * n + w0] = p[j];
            }
            __syncthreads();
            // ---------------- phase 3: z-line (x = w0, y = w1), scatter
            {
              double s[n], p[n], o1[n], o2[n];
#pragma unroll
              for (int z = 0; z < n; ++z)
                {
                  s[z] = myC[(z * n + w1) * n + w0];
                  p[z] = myD[(z * n + w1) * n + w0];
                }
              matvec<n>(Mh, s, o1);
              matvec<n>(Kh, p, o2);
              const double cl = op.cl[b];
              const int    gx = cx * K + w0, gy = cy * K + w1;
              if (active && !on_bdry(gx, n1) && !on_bdry(gy, n1))
                {
#pragma unroll
                  for (int z = 0; z < n; ++z)
                    {
                      const int gz = cz * K + z;
                      if (!on_bdry(gz, n1))
                        atomicAdd(dst + b * stride + gx + (long long)n1 * (gy + (long long)n1 * gz), fma(cl, o2[z], o1[z]));
                    }
                }
            }
          }
      }
  }

  template <int K, bool COUPLED>
  __global__ void __launch_bounds__(CfgV1<K>::T2)
    k_apply2d_v1(const Geo g, const OpDev op, double *__restrict__ dst, const double *__restrict__ src, const long long stride)
  {
    using C             = CfgV1<K>;
    constexpr int n = C::n, n2 = C::n2, CPB = C::CPB2;
    __shared__ double sA[CPB * n2], sB[CPB * n2];
    __shared__ double sE[COUPLED ? CPB * n2 : 1];
    const double *Mh = c_fe[K].Mh, *Kh = c_fe[K].Kh;

    const int       t = threadIdx.x, lc = t / n, w = t % n;
    const int       nc = g.nc, n1 = g.n1;
    const long long ncells = (long long)nc * nc, nbatches = (ncells + CPB - 1) / CPB;
    double         *myA = sA + lc * n2, *myB = sB + lc * n2, *myE = sE + (COUPLED ? lc * n2 : 0);

    for (long long batch = blockIdx.x; batch < nbatches; batch += gridDim.x)
      {
        const long long c      = batch * CPB + lc;
        const bool      active = (lc < CPB) && c < ncells;
        const int       cx = active ? (int)(c % nc) : 0, cy = active ? (int)(c / nc) : 0;
        const int       gy1     = cy * K + w;
        const bool      line_bd = on_bdry(gy1, n1);
        const long long base1   = (long long)cx * K + (long long)n1 * gy1;
        for (int b = 0; b < op.nb; ++b)
          {
            double u[n], um[n];
#pragma unroll
            for (int i = 0; i < n; ++i)
              {
                const bool bd = line_bd || (i == 0 && cx == 0) || (i == K && cx == nc - 1);
                u[i]          = (active && !bd) ? src[b * stride + base1 + i] : 0.0;
                um[i]         = 0.0;
              }
            if (COUPLED)
              for (int j = 0; j < op.nb; ++j)
                {
                  const double cij = op.cc[b * op.nb + j];
                  if (cij != 0.0)
                    {
#pragma unroll
                      for (int i = 0; i < n; ++i)
                        {
                          const bool bd = line_bd || (i == 0 && cx == 0) || (i == K && cx == nc - 1);
                          um[i] = fma(cij, (active && !bd) ? src[j * stride + base1 + i] : 0.0, um[i]);
                        }
                    }
                }
            {
              double a[n], bk[n];
              matvec<n>(Mh, u, a);
              matvec<n>(Kh, u, bk);
              if (lc < CPB)
                {
#pragma unroll
                  for (int i = 0; i < n; ++i)
                    {
                      myA[w * n + i] = a[i];
                      myB[w * n + i] = bk[i];
                    }
                  if (COUPLED)
                    {
                      double am[n];
                      matvec<n>(Mh, um, am);
#pragma unroll
                      for (int i = 0; i < n; ++i)
                        myE[w * n + i] = am[i];
                    }
                }
            }
            __syncthreads();
            if (lc < CPB)
              {
                double a[n], s[n], o1[n], o2[n];
                const double cl = op.cl[b];
#pragma unroll
                for (int j = 0; j < n; ++j)
                  {
                    a[j] = myA[j * n + w];
                    if (COUPLED)
                      s[j] = fma(cl, myB[j * n + w], myE[j * n + w]);
                    else
                      s[j] = fma(op.cm[b], a[j], cl * myB[j * n + w]);
                  }
                matvec<n>(Mh, s, o1);
                matvec<n>(Kh, a, o2);
                const int gx = cx * K + w;
                if (active && !on_bdry(gx, n1))
                  {
#pragma unroll
                    for (int j = 0; j < n; ++j)
                      {
                        const int gy = cy * K + j;
                        if (!on_bdry(gy, n1))
                          atomicAdd(dst + b * stride + gx + (long long)n1 * gy, fma(cl, o2[j], o1[j]));
                      }
                  }
              }
            __syncthreads();
          }
      }
  }

  // ---- unfused epilogues ------------------------------------------------------------------
  // dst = rhs - t
  static __global__ void k_residual_epilogue(const long long N, const int nb, double *__restrict__ dst, const double *__restrict__ rhs,
                                      const double *__restrict__ t, const long long stride, const long long tstride)
  {
    const long long total = N * nb;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x)
      {
        const int       b = e / N;
        const long long i = e - b * N;
        dst[b * stride + i] = rhs[b * stride + i] - t[b * tstride + i];
      }
  }

  struct ChebFactors
  {
    double f1[SPIRK_MAX_BLOCKS], f2[SPIRK_MAX_BLOCKS];
  };

  // x_new = (1 + f1) x - f1 x_old + f2 dinv (rhs - t)      (deal.II VectorUpdater)
  static __global__ void k_cheb_epilogue(const long long N, const int nb, double *x_new, const double *__restrict__ x,
                                  const double *x_old, const double *__restrict__ rhs, const double *__restrict__ dinv,
                                  const double *__restrict__ t, const long long stride, const long long tstride,
                                  const ChebFactors f, const long long dstride)
  {
    const long long total = N * nb;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x)
      {
        const int       b  = e / N;
        const long long i  = e - b * N, j = b * stride + i;
        const double    xo = x_old ? x_old[j] : 0.0;
        x_new[j]           = (1.0 + f.f1[b]) * x[j] - f.f1[b] * xo + f.f2[b] * dinv[b * dstride + i] * (rhs[j] - t[b * tstride + i]);
      }
  }

  // assembled 1-D diagonal entry of the reference matrix A (Mh or Kh) at global node i
  template <int K>
  __device__ __forceinline__ double diag1d(const double *A, const int i, const int n1)
  {
    constexpr int n = K + 1;
    if (i == 0)
      return A[0];
    if (i == n1 - 1)
      return A[K * n + K];
    const int l = i % K;
    return (l == 0) ? A[0] + A[K * n + K] : A[l * n + l];
  }

  // inverse diagonal (operator.h:361-373): abs(d) > 1e-10 ? 1/d : 1; Dirichlet entries 1
  template <int K>
  __global__ void k_inverse_diagonal(const Geo g, const double cm, const double cl, double *__restrict__ diag)
  {
    const double *Mh = c_fe[K].Mh, *Kh = c_fe[K].Kh;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < g.N; i += (long long)gridDim.x * blockDim.x)
      {
        const int  ix = i % g.n1, iy = (i / g.n1) % g.n1, iz = (g.dim == 3) ? (int)(i / ((long long)g.n1 * g.n1)) + g.zo0 : 1;
        const bool bd = on_bdry(ix, g.n1) || on_bdry(iy, g.n1) || (g.dim == 3 && on_bdry(iz, g.n1));
        double     d  = 0.0;
        if (!bd)
          {
            const double mx = diag1d<K>(Mh, ix, g.n1), kx = diag1d<K>(Kh, ix, g.n1);
            const double my = diag1d<K>(Mh, iy, g.n1), ky = diag1d<K>(Kh, iy, g.n1);
            if (g.dim == 3)
              {
                const double mz = diag1d<K>(Mh, iz, g.n1), kz = diag1d<K>(Kh, iz, g.n1);
                d = cm * mx * my * mz + cl * (kx * my * mz + mx * ky * mz + mx * my * kz);
              }
            else
              d = cm * mx * my + cl * (kx * my + mx * ky);
          }
        diag[i] = (fabs(d) > 1.0e-10) ? 1.0 / d : 1.0;
      }
  }
} // namespace spirk
