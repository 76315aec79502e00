// 1-D finite-element tables for FE_Q(k) on Gauss-Lobatto nodes with QGauss(k+1) quadrature
// (what the reference gets from deal.II: `FE_Q<dim> fe(k)`, `QGauss<dim> quadrature(k+1)`,
// main.cc:3028-3029).  Host-side, computed once per degree and uploaded to constant memory.
//
// On the refined hypercube every cell is the same axis-aligned cube of side h (MappingQ1,
// operator.h:262-263), so the cell matrices factor into the reference 1-D matrices
//   Mh = B^T W B   and   Kh = D^T W D     (B: values, D: derivatives at the Gauss points)
// scaled by powers of h.  The kernels use Mh / Kh directly ("Cartesian fast path").
#pragma once
#include <cmath>
#include <vector>

namespace spirk
{
  struct Fe1D
  {
    int                 k = 0, n = 0;
    std::vector<double> nodes;  // n GLL nodes on [0,1]
    std::vector<double> xq, wq; // n Gauss points / weights on [0,1]
    std::vector<double> B, D;   // n x n, row = quadrature point, col = basis function
    std::vector<double> Mh, Kh; // n x n reference mass / stiffness
    std::vector<double> P;      // (2k+1) x n two-level embedding
    std::vector<double> xe, we, Be; // QGauss(k+2) for the error norms, Be: (k+2) x n
  };

  namespace detail
  {
    // Legendre polynomial P_m and its first derivative by the three-term recurrence
    inline void legendre_pd(const int m, const long double x, long double &p, long double &dp)
    {
      long double a = 1.0L, b = x;
      if (m == 0)
        {
          p = 1.0L, dp = 0.0L;
          return;
        }
      for (int j = 1; j < m; ++j)
        {
          const long double c = ((2 * j + 1) * x * b - j * a) / (j + 1);
          a = b, b = c;
        }
      p  = b;
      dp = m * (a - x * b) / (1.0L - x * x);
    }

    inline std::vector<double> gauss_legendre(const int m, std::vector<double> &w)
    {
      std::vector<double> x(m);
      w.assign(m, 0.0);
      for (int i = 0; i < (m + 1) / 2; ++i)
        {
          long double z = std::cos(M_PIl * (i + 0.75L) / (m + 0.5L)), p, dp;
          for (int it = 0; it < 64; ++it)
            {
              legendre_pd(m, z, p, dp);
              const long double dz = p / dp;
              z -= dz;
              if (fabsl(dz) < 1e-19L)
                break;
            }
          legendre_pd(m, z, p, dp);
          const long double wt = 2.0L / ((1.0L - z * z) * dp * dp);
          x[m - 1 - i] = (double)(0.5L * (1.0L + z));
          x[i]         = (double)(0.5L * (1.0L - z));
          w[i] = w[m - 1 - i] = (double)(0.5L * wt);
        }
      return x;
    }

    inline std::vector<double> gauss_lobatto(const int k)
    {
      // interior nodes are the roots of P_k'; Newton with P_k'' from the Legendre ODE
      std::vector<double> x(k + 1);
      x[0] = 0.0, x[k] = 1.0;
      for (int i = 1; i <= k / 2; ++i)
        {
          long double z = std::cos(M_PIl * i / k), p, dp;
          for (int it = 0; it < 64; ++it)
            {
              legendre_pd(k, z, p, dp);
              const long double ddp = (2.0L * z * dp - (long double)k * (k + 1) * p) / (1.0L - z * z);
              const long double dz  = dp / ddp;
              z -= dz;
              if (fabsl(dz) < 1e-19L)
                break;
            }
          x[k - i] = (double)(0.5L * (1.0L + z));
          x[i]     = (double)(0.5L * (1.0L - z));
        }
      if (k % 2 == 0)
        x[k / 2] = 0.5;
      return x;
    }

    inline double lagr(const std::vector<double> &pts, const int i, const double x)
    {
      double v = 1.0;
      for (int j = 0; j < (int)pts.size(); ++j)
        if (j != i)
          v *= (x - pts[j]) / (pts[i] - pts[j]);
      return v;
    }

    inline double lagr_dx(const std::vector<double> &pts, const int i, const double x)
    {
      double sum = 0.0;
      for (int m = 0; m < (int)pts.size(); ++m)
        {
          if (m == i)
            continue;
          double term = 1.0 / (pts[i] - pts[m]);
          for (int j = 0; j < (int)pts.size(); ++j)
            if (j != i && j != m)
              term *= (x - pts[j]) / (pts[i] - pts[j]);
          sum += term;
        }
      return sum;
    }
  } // namespace detail

  inline Fe1D make_fe1d(const int k)
  {
    Fe1D f;
    f.k = k, f.n = k + 1;
    const int n = f.n;
    f.nodes = detail::gauss_lobatto(k);
    f.xq    = detail::gauss_legendre(n, f.wq);
    f.B.resize(n * n), f.D.resize(n * n), f.Mh.assign(n * n, 0.0), f.Kh.assign(n * n, 0.0);
    for (int q = 0; q < n; ++q)
      for (int i = 0; i < n; ++i)
        {
          f.B[q * n + i] = detail::lagr(f.nodes, i, f.xq[q]);
          f.D[q * n + i] = detail::lagr_dx(f.nodes, i, f.xq[q]);
        }
    for (int i = 0; i < n; ++i)
      for (int j = 0; j <= i; ++j)
        {
          double m = 0, s = 0;
          for (int q = 0; q < n; ++q)
            {
              m += f.wq[q] * f.B[q * n + i] * f.B[q * n + j];
              s += f.wq[q] * f.D[q * n + i] * f.D[q * n + j];
            }
          f.Mh[i * n + j] = f.Mh[j * n + i] = m;
          f.Kh[i * n + j] = f.Kh[j * n + i] = s;
        }
    const int m = 2 * k + 1;
    f.P.resize(m * n);
    for (int r = 0; r < m; ++r)
      {
        const double xf = (r <= k) ? 0.5 * f.nodes[r] : 0.5 + 0.5 * f.nodes[r - k];
        for (int i = 0; i < n; ++i)
          {
            const double v = detail::lagr(f.nodes, i, xf);
            f.P[r * n + i] = std::fabs(v) < 1e-15 ? 0.0 : v;
          }
      }
    f.xe = detail::gauss_legendre(k + 2, f.we);
    f.Be.resize((k + 2) * n);
    for (int q = 0; q < k + 2; ++q)
      for (int i = 0; i < n; ++i)
        f.Be[q * n + i] = detail::lagr(f.nodes, i, f.xe[q]);
    return f;
  }
} // namespace spirk
