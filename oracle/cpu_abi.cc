/*
 * ORACLE — TEST INFRASTRUCTURE ONLY ("port" oracle).  Never linked into the product.
 *
 * CPU implementation of include/spirk_b200.h that restates, in plain C++, what the reference
 * does through deal.II on the CPU: a *cell loop* with sum factorisation in deal.II's order of
 * operations (gather with Dirichlet DoFs read as zero -> interpolate to the Gauss points ->
 * collocation derivative -> multiply by JxW / J^-T J^-1 and the mass / Laplace scalings ->
 * integrate -> scatter-add skipping Dirichlet DoFs -> identity on Dirichlet DoFs).
 *   ref include/operator.h:298-310, 379-421 (scalar), 616-665 (complex), 841-880 (batched)
 *   ref include/preconditioner.h:266-282 (transfer), 353-373 (Chebyshev), 375-413 (coarse)
 *   deal.II conventions: SURVEY.md Appendix A (A2, A3, A4, A7, A9, A12)
 *
 * PARITY UNPINNED: the reference has no tests / golden vectors and deal.II cannot be built
 * here; this file is validated against the independent NumPy oracle (oracle/spirk_oracle.py,
 * assembled Kronecker form + sparse direct solves) in tests/test_oracle_cpu.py.
 *
 * Used only by: tests/ (as checker and as the CPU double that lets the C++ host logic be
 * exercised without a GPU) and bench.py's cpu_baseline / --impl reference legs.
 */
#include "../include/spirk_b200.h"

#include <omp.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <tuple>
#include <vector>

namespace
{
  thread_local std::string g_err;
  int fail(int code, const std::string &msg)
  {
    g_err = msg;
    return code;
  }

  // ---------------------------------------------------------------- 1-D FE tables
  double legendre(int n, double x, double *dp = nullptr)
  {
    double p0 = 1, p1 = x;
    if (n == 0)
      {
        if (dp)
          *dp = 0;
        return 1;
      }
    for (int j = 2; j <= n; ++j)
      {
        const double p2 = ((2 * j - 1) * x * p1 - (j - 1) * p0) / j;
        p0 = p1, p1 = p2;
      }
    if (dp)
      *dp = n * (x * p1 - p0) / (x * x - 1.0);
    return p1;
  }

  std::vector<double> gauss_points(int n, std::vector<double> &w)
  {
    std::vector<double> x(n);
    w.resize(n);
    for (int i = 0; i < n; ++i)
      {
        double z = -std::cos(M_PI * (i + 0.75) / (n + 0.5)), dp = 0;
        for (int it = 0; it < 100; ++it)
          {
            const double p = legendre(n, z, &dp), dz = p / dp;
            z -= dz;
            if (std::abs(dz) < 1e-16)
              break;
          }
        legendre(n, z, &dp);
        x[i] = 0.5 * (z + 1.0);
        w[i] = 1.0 / ((1.0 - z * z) * dp * dp);
      }
    for (int i = 0; i < n / 2; ++i)
      { // symmetrise
        const double a = 0.5 * (x[i] + 1.0 - x[n - 1 - i]);
        x[i] = a, x[n - 1 - i] = 1.0 - a;
        const double b = 0.5 * (w[i] + w[n - 1 - i]);
        w[i] = w[n - 1 - i] = b;
      }
    if (n % 2)
      x[n / 2] = 0.5;
    return x;
  }

  std::vector<double> gll_points(int k)
  { // k+1 Gauss-Lobatto points on [0,1]: end points + roots of P_k'
    std::vector<double> x(k + 1);
    x[0] = 0, x[k] = 1;
    for (int i = 1; i < k; ++i)
      {
        double z = -std::cos(M_PI * i / k);
        for (int it = 0; it < 100; ++it)
          {
            double       dp;
            const double p   = legendre(k, z, &dp);
            const double ddp = (2 * z * dp - k * (k + 1) * p) / (1 - z * z);
            const double dz  = dp / ddp;
            z -= dz;
            if (std::abs(dz) < 1e-16)
              break;
          }
        x[i] = 0.5 * (z + 1.0);
      }
    for (int i = 0; i <= k / 2; ++i)
      {
        const double a = 0.5 * (x[i] + 1.0 - x[k - i]);
        x[i] = a, x[k - i] = 1.0 - a;
      }
    return x;
  }

  double lagrange(const std::vector<double> &nodes, int i, double x)
  {
    double v = 1;
    for (size_t j = 0; j < nodes.size(); ++j)
      if ((int)j != i)
        v *= (x - nodes[j]) / (nodes[i] - nodes[j]);
    return v;
  }
  double lagrange_deriv(const std::vector<double> &nodes, int i, double x)
  {
    double s = 0;
    for (size_t m = 0; m < nodes.size(); ++m)
      if ((int)m != i)
        {
          double t = 1.0 / (nodes[i] - nodes[m]);
          for (size_t j = 0; j < nodes.size(); ++j)
            if ((int)j != i && j != m)
              t *= (x - nodes[j]) / (nodes[i] - nodes[j]);
          s += t;
        }
    return s;
  }

  struct FE
  {
    int                 k, n;
    std::vector<double> nodes, xq, wq;
    std::vector<double> B;    // n x n: B[q*n+i] = l_i(x_q)
    std::vector<double> Dcol; // n x n: collocation derivative at the Gauss points
    std::vector<double> P;    // (2k+1) x n prolongation
    std::vector<double> xe, we, Be; // error quadrature QGauss(k+2): Be[(k+2) x n]
  };

  const FE &get_fe(int k)
  {
    static std::map<int, std::unique_ptr<FE>> cache;
    static omp_lock_t                         lock;
    static bool                               init = (omp_init_lock(&lock), true);
    (void)init;
    omp_set_lock(&lock);
    auto &p = cache[k];
    if (!p)
      {
        p        = std::make_unique<FE>();
        FE &f    = *p;
        f.k      = k;
        f.n      = k + 1;
        const int n = f.n;
        f.nodes  = gll_points(k);
        f.xq     = gauss_points(n, f.wq);
        f.B.resize(n * n);
        f.Dcol.resize(n * n);
        for (int q = 0; q < n; ++q)
          for (int i = 0; i < n; ++i)
            {
              f.B[q * n + i]    = lagrange(f.nodes, i, f.xq[q]);
              f.Dcol[q * n + i] = lagrange_deriv(f.xq, i, f.xq[q]);
            }
        f.P.resize((2 * k + 1) * n);
        for (int j = 0; j < 2 * k + 1; ++j)
          {
            const double xf = (j <= k) ? 0.5 * f.nodes[j] : 0.5 + 0.5 * f.nodes[j - k];
            for (int i = 0; i < n; ++i)
              {
                double v = lagrange(f.nodes, i, xf);
                if (std::abs(v) < 1e-15)
                  v = 0;
                f.P[j * n + i] = v;
              }
          }
        f.xe = gauss_points(k + 2, f.we);
        f.Be.resize((k + 2) * n);
        for (int q = 0; q < k + 2; ++q)
          for (int i = 0; i < n; ++i)
            f.Be[q * n + i] = lagrange(f.nodes, i, f.xe[q]);
      }
    omp_unset_lock(&lock);
    return *p;
  }

  // apply the (m x n) matrix A (or its transpose, then A is n x m and we contract its rows)
  // along direction `dir` of a tensor whose extents along dir is n (in) -> m (out)
  inline void apply1d(const double *A, int m, int n, bool transpose, int dir, int dim, const int *ext_in,
                      const double *in, double *out)
  {
    int pre = 1, post = 1;
    for (int d = 0; d < dir; ++d)
      pre *= ext_in[d];
    for (int d = dir + 1; d < dim; ++d)
      post *= ext_in[d];
    for (int o = 0; o < post; ++o)
      for (int r = 0; r < m; ++r)
        for (int i = 0; i < pre; ++i)
          {
            double s = 0;
            for (int j = 0; j < n; ++j)
              s += (transpose ? A[j * m + r] : A[r * n + j]) * in[(o * n + j) * pre + i];
            out[(o * m + r) * pre + i] = s;
          }
  }

  struct Geo
  {
    int       dim, k, n, nc, n1;
    long long N;
    double    h;
    Geo(const spirk_level *l)
      : dim(l->dim)
      , k(l->degree)
      , n(l->degree + 1)
      , nc(l->n_cells_1d)
      , n1(l->degree * l->n_cells_1d + 1)
    {
      N = (long long)n1 * n1 * (dim == 3 ? n1 : 1);
      h = 1.0 / nc;
    }
    bool on_boundary(int ix, int iy, int iz) const
    {
      return ix == 0 || ix == n1 - 1 || iy == 0 || iy == n1 - 1 || (dim == 3 && (iz == 0 || iz == n1 - 1));
    }
  };

  int check_level(const spirk_level *l)
  {
    if (!l || (l->dim != 2 && l->dim != 3) || l->degree < 1 || l->degree > 6 || l->n_cells_1d < 1)
      return fail(SPIRK_ERR_INVALID, "bad level");
    if (((l->slab >> 8) & 0xff) > 1)
      return fail(SPIRK_ERR_UNSUPPORTED, "z-slab levels are not available on the CPU double");
    return SPIRK_OK;
  }

  // the cell integral for one cell and one block:
  //   out = B^T [ JxW * vm + sum_d Dcol_d^T ( lap * JxW/h^2 * Dcol_d v ) ],  v = B u, vm = B um
  struct CellKernel
  {
    const FE &fe;
    int       dim, n, nl;
    double    jxw_scale, grad_scale;
    std::vector<double> w3; // tensor-product quadrature weights
    std::vector<double> t0, t1, vq, vmq, acc, g;
    CellKernel(const Geo &geo)
      : fe(get_fe(geo.k))
      , dim(geo.dim)
      , n(geo.n)
    {
      nl = 1;
      for (int d = 0; d < dim; ++d)
        nl *= n;
      jxw_scale  = std::pow(geo.h, dim);
      grad_scale = 1.0 / (geo.h * geo.h);
      w3.resize(nl);
      for (int q = 0; q < nl; ++q)
        {
          int    r = q;
          double w = 1;
          for (int d = 0; d < dim; ++d)
            w *= fe.wq[r % n], r /= n;
          w3[q] = w;
        }
      for (auto *v : {&t0, &t1, &vq, &vmq, &acc, &g})
        v->resize(nl);
    }
    void interpolate(const double *u, double *out)
    {
      int ext[3] = {n, n, n};
      const double *in = u;
      double *bufs[2] = {t0.data(), t1.data()};
      for (int d = 0; d < dim; ++d)
        {
          double *o = (d == dim - 1) ? out : bufs[d % 2];
          apply1d(fe.B.data(), n, n, false, d, dim, ext, in, o);
          in = o;
        }
    }
    // u: Laplace input, um: mass input (already multiplied by mass scaling / coupling); either may be null
    void run(const double *u, const double *um, double lap, double *out)
    {
      int ext[3] = {n, n, n};
      std::fill(acc.begin(), acc.end(), 0.0);
      if (um)
        {
          interpolate(um, vmq.data());
          for (int q = 0; q < nl; ++q)
            acc[q] += vmq[q] * (w3[q] * jxw_scale);
        }
      if (u && lap != 0.0)
        {
          interpolate(u, vq.data());
          for (int d = 0; d < dim; ++d)
            {
              apply1d(fe.Dcol.data(), n, n, false, d, dim, ext, vq.data(), g.data());
              for (int q = 0; q < nl; ++q)
                g[q] *= lap * (w3[q] * jxw_scale) * grad_scale;
              apply1d(fe.Dcol.data(), n, n, true, d, dim, ext, g.data(), t0.data());
              for (int q = 0; q < nl; ++q)
                acc[q] += t0[q];
            }
        }
      const double *in = acc.data();
      double *bufs[2] = {t0.data(), t1.data()};
      for (int d = 0; d < dim; ++d)
        {
          double *o = (d == dim - 1) ? out : bufs[d % 2];
          apply1d(fe.B.data(), n, n, true, d, dim, ext, in, o);
          in = o;
        }
    }
  };

  // ---------------------------------------------------------------------------------------------------------
  // Vectorised cell loop: what deal.II's MatrixFree does on a CPU - batches of VL = 8 cells in
  // structure-of-arrays layout (one AVX-512 lane per cell, VectorizedArray<double>), sum factorisation with compile-time
  // extents: interpolate to the Gauss points (dim sweeps), collocation derivative / scaling by JxW and J^-T J^-1 /
  // transposed derivative (2 dim sweeps), integrate (dim sweeps); the mass term and the coupling of the blocks are mixed
  // at the quadrature points (operator.h:639-647).  Threads: rows of cells along x, coloured by the parity of the other
  // cell indices so that concurrently processed rows share no DoF.
  constexpr int VL = 8;
  template <int n, int dim>
  struct FastCell
  {
    static constexpr int nl = (dim == 3) ? n * n * n : n * n;
    // out[.., r, ..] = sum_j A[r][j] (or A[j][r]) in[.., j, ..] along direction d, for all VL lanes
    template <bool transpose, bool add, int d>
    static inline void sweep(const double *__restrict__ A, const double (*__restrict__ in)[VL], double (*__restrict__ out)[VL])
    {
      constexpr int pre = (d == 0) ? 1 : (d == 1 ? n : n * n), post = nl / (pre * n);
      for (int o = 0; o < post; ++o)
        for (int i = 0; i < pre; ++i)
          for (int r = 0; r < n; ++r)
            {
              double s[VL];
#pragma omp simd
              for (int l = 0; l < VL; ++l)
                s[l] = 0.0;
              for (int j = 0; j < n; ++j)
                {
                  const double a = transpose ? A[j * n + r] : A[r * n + j];
#pragma omp simd
                  for (int l = 0; l < VL; ++l)
                    s[l] += a * in[(o * n + j) * pre + i][l];
                }
#pragma omp simd
              for (int l = 0; l < VL; ++l)
                out[(o * n + r) * pre + i][l] = add ? out[(o * n + r) * pre + i][l] + s[l] : s[l];
            }
    }
  };

  template <int n, int dim>
  void cell_loop_fast(const Geo &geo, const spirk_opdesc *op, double *dst, const double *src, long long stride)
  {
    using FC          = FastCell<n, dim>;
    constexpr int nl  = FC::nl;
    const FE     &fe  = get_fe(geo.k);
    const int     k = geo.k, nc = geo.nc, n1 = geo.n1, nb = op->nb;
    const double  jxw = std::pow(geo.h, dim), gs = 1.0 / (geo.h * geo.h);
    double        w3[nl];
    for (int q = 0; q < nl; ++q)
      {
        int    r = q;
        double w = 1;
        for (int d = 0; d < dim; ++d)
          w *= fe.wq[r % n], r /= n;
        w3[q] = w * jxw;
      }
    const double *B = fe.B.data(), *D = fe.Dcol.data();
    const int     nrows = (dim == 3) ? nc * nc : nc; // rows of cells along x
    for (int colour = 0; colour < (dim == 3 ? 4 : 2); ++colour)
      {
#pragma omp parallel
        {
          // per thread: values at the nodes / quadrature points of every block, work arrays
          std::vector<double> store((size_t)(2 * nb + 4) * nl * VL);
          auto                arr = [&](int i) { return reinterpret_cast<double(*)[VL]>(store.data() + (size_t)i * nl * VL); };
          double(*t0)[VL] = arr(2 * nb), (*t1)[VL] = arr(2 * nb + 1), (*acc)[VL] = arr(2 * nb + 2), (*gq)[VL] = arr(2 * nb + 3);
#pragma omp for schedule(dynamic, 1)
          for (int row = 0; row < nrows; ++row)
            {
              const int cy = (dim == 3) ? row % nc : row, cz = (dim == 3) ? row / nc : 0;
              if ((cy & 1) + 2 * (cz & 1) != colour)
                continue;
              for (int cx0 = 0; cx0 < nc; cx0 += VL)
                {
                  const int nv = std::min(VL, nc - cx0);
                  // gather (Dirichlet DoFs read as zero) and interpolate every block to the Gauss points
                  for (int b = 0; b < nb; ++b)
                    {
                      double(*u)[VL] = arr(b), (*vq)[VL] = arr(nb + b);
                      for (int l = 0; l < nl; ++l)
                        {
                          const int lx = l % n, ly = (l / n) % n, lz = (dim == 3) ? l / (n * n) : 0;
                          const int iy = cy * k + ly, iz = (dim == 3) ? cz * k + lz : 0;
                          const bool bd_yz = iy == 0 || iy == n1 - 1 || (dim == 3 && (iz == 0 || iz == n1 - 1));
                          const double *p  = src + b * stride + (long long)n1 * (iy + (long long)n1 * iz);
                          for (int v = 0; v < VL; ++v)
                            {
                              const int ix = (cx0 + v) * k + lx;
                              u[l][v]      = (v < nv && !bd_yz && ix != 0 && ix != n1 - 1) ? p[ix] : 0.0;
                            }
                        }
                      FC::template sweep<false, false, 0>(B, u, t0);
                      if constexpr (dim == 3)
                        {
                          FC::template sweep<false, false, 1>(B, t0, t1);
                          FC::template sweep<false, false, 2>(B, t1, vq);
                        }
                      else
                        FC::template sweep<false, false, 1>(B, t0, vq);
                    }
                  for (int b = 0; b < nb; ++b)
                    {
                      double(*vq)[VL] = arr(nb + b);
                      // mass term at the quadrature points: mass_b v_b, or sum_j C_bj v_j for coupled operators
                      for (int q = 0; q < nl; ++q)
                        {
#pragma omp simd
                          for (int v = 0; v < VL; ++v)
                            acc[q][v] = 0.0;
                        }
                      for (int j = 0; j < nb; ++j)
                        {
                          const double c = (op->kind == SPIRK_OP_REAL) ? (j == b ? op->mass[b] : 0.0) : op->coupling[b * nb + j];
                          if (c == 0.0)
                            continue;
                          double(*vj)[VL] = arr(nb + j);
                          for (int q = 0; q < nl; ++q)
                            {
                              const double cw = c * w3[q];
#pragma omp simd
                              for (int v = 0; v < VL; ++v)
                                acc[q][v] += cw * vj[q][v];
                            }
                        }
                      const double lap = op->laplace[b];
                      if (lap != 0.0)
                        {
                          const auto scale = [&]() {
                            for (int q = 0; q < nl; ++q)
                              {
                                const double cw = lap * w3[q] * gs;
#pragma omp simd
                                for (int v = 0; v < VL; ++v)
                                  gq[q][v] *= cw;
                              }
                          };
                          FC::template sweep<false, false, 0>(D, vq, gq);
                          scale();
                          FC::template sweep<true, true, 0>(D, gq, acc);
                          FC::template sweep<false, false, 1>(D, vq, gq);
                          scale();
                          FC::template sweep<true, true, 1>(D, gq, acc);
                          if constexpr (dim == 3)
                            {
                              FC::template sweep<false, false, 2>(D, vq, gq);
                              scale();
                              FC::template sweep<true, true, 2>(D, gq, acc);
                            }
                        }
                      // integrate and scatter (constrained DoFs skipped); lanes one after the other: neighbouring cells of the
                      // batch share their x-face DoFs
                      double(*out)[VL] = arr(b); // (the nodal values of block b are no longer needed)
                      if constexpr (dim == 3)
                        {
                          FC::template sweep<true, false, 0>(B, acc, t0);
                          FC::template sweep<true, false, 1>(B, t0, t1);
                          FC::template sweep<true, false, 2>(B, t1, out);
                        }
                      else
                        {
                          FC::template sweep<true, false, 0>(B, acc, t0);
                          FC::template sweep<true, false, 1>(B, t0, out);
                        }
                      for (int l = 0; l < nl; ++l)
                        {
                          const int lx = l % n, ly = (l / n) % n, lz = (dim == 3) ? l / (n * n) : 0;
                          const int iy = cy * k + ly, iz = (dim == 3) ? cz * k + lz : 0;
                          if (iy == 0 || iy == n1 - 1 || (dim == 3 && (iz == 0 || iz == n1 - 1)))
                            continue;
                          double *p = dst + b * stride + (long long)n1 * (iy + (long long)n1 * iz);
                          for (int v = 0; v < nv; ++v)
                            {
                              const int ix = (cx0 + v) * k + lx;
                              if (ix != 0 && ix != n1 - 1)
                                p[ix] += out[l][v];
                            }
                        }
                    }
                }
            }
        }
      }
  }

  // the matrix-free cell loop (ref operator.h:298-310 + 379-421 / 616-665 / 841-880)
  void cell_loop(const spirk_level *lvl, const spirk_opdesc *op, double *dst, const double *src, long long stride)
  {
    const Geo geo(lvl);
    if (!std::getenv("SPIRK_CPU_SCALAR_CELLS"))
      {
        // zero dst, vectorised cell batches, identity on constrained DoFs (below)
        for (int b = 0; b < op->nb; ++b)
          {
            double *d = dst + b * stride;
#pragma omp parallel for
            for (long long i = 0; i < geo.N; ++i)
              d[i] = 0.0;
          }
#define SPIRK_FAST(NN)                                                        \
  case NN:                                                                    \
    if (geo.dim == 3)                                                         \
      cell_loop_fast<NN, 3>(geo, op, dst, src, stride);                       \
    else                                                                      \
      cell_loop_fast<NN, 2>(geo, op, dst, src, stride);                       \
    break;
        switch (geo.n)
          {
            SPIRK_FAST(2) SPIRK_FAST(3) SPIRK_FAST(4) SPIRK_FAST(5) SPIRK_FAST(6) SPIRK_FAST(7)
          }
#undef SPIRK_FAST
        for (int b = 0; b < op->nb; ++b)
          {
#pragma omp parallel for
            for (long long i = 0; i < geo.N; ++i)
              {
                const int ix = i % geo.n1, iy = (i / geo.n1) % geo.n1, iz = (geo.dim == 3) ? i / ((long long)geo.n1 * geo.n1) : 0;
                if (geo.on_boundary(ix, iy, iz))
                  dst[b * stride + i] = src[b * stride + i];
              }
          }
        return;
      }
    const int dim = geo.dim, n = geo.n, k = geo.k, nc = geo.nc, n1 = geo.n1, nb = op->nb;
    const long long N = geo.N;
    int nl = 1;
    for (int d = 0; d < dim; ++d)
      nl *= n;
    for (int b = 0; b < nb; ++b)
      {
        double *d = dst + b * stride;
#pragma omp parallel for
        for (long long i = 0; i < N; ++i)
          d[i] = 0.0;
      }
    const int nouter = nc; // outermost cell direction: z in 3-D, y in 2-D
    const int ncz = (dim == 3) ? nc : 1;
    (void)ncz;
    for (int phase = 0; phase < 2; ++phase)
      {
#pragma omp parallel
        {
          CellKernel          ck(geo);
          std::vector<double> u(nb * nl), um(nl), out(nl);
#pragma omp for schedule(static)
          for (int co = phase; co < nouter; co += 2)
            {
              const int ncy = (dim == 3) ? nc : 1;
              for (int cm = 0; cm < ncy; ++cm)
                for (int cx = 0; cx < nc; ++cx)
                  {
                    const int cy = (dim == 3) ? cm : co, cz = (dim == 3) ? co : 0;
                    // gather all blocks, Dirichlet read as zero
                    for (int b = 0; b < nb; ++b)
                      for (int l = 0; l < nl; ++l)
                        {
                          const int       ix = cx * k + l % n, iy = cy * k + (l / n) % n, iz = (dim == 3) ? cz * k + l / (n * n) : 0;
                          const long long gi = ix + (long long)n1 * (iy + (long long)n1 * iz);
                          u[b * nl + l]      = geo.on_boundary(ix, iy, iz) ? 0.0 : src[b * stride + gi];
                        }
                    for (int b = 0; b < nb; ++b)
                      {
                        bool have_mass = false;
                        if (op->kind == SPIRK_OP_REAL)
                          {
                            if (op->mass[b] != 0.0)
                              {
                                have_mass = true;
                                for (int l = 0; l < nl; ++l)
                                  um[l] = op->mass[b] * u[b * nl + l];
                              }
                          }
                        else
                          {
                            std::fill(um.begin(), um.end(), 0.0);
                            for (int j = 0; j < nb; ++j)
                              {
                                const double c = op->coupling[b * nb + j];
                                if (c != 0.0)
                                  {
                                    have_mass = true;
                                    for (int l = 0; l < nl; ++l)
                                      um[l] += c * u[j * nl + l];
                                  }
                              }
                          }
                        ck.run(&u[b * nl], have_mass ? um.data() : nullptr, op->laplace[b], out.data());
                        for (int l = 0; l < nl; ++l)
                          {
                            const int ix = cx * k + l % n, iy = cy * k + (l / n) % n, iz = (dim == 3) ? cz * k + l / (n * n) : 0;
                            if (!geo.on_boundary(ix, iy, iz))
                              dst[b * stride + ix + (long long)n1 * (iy + (long long)n1 * iz)] += out[l];
                          }
                      }
                  }
            }
        }
      }
    // identity on constrained DoFs
    for (int b = 0; b < nb; ++b)
      {
#pragma omp parallel for
        for (long long i = 0; i < N; ++i)
          {
            const int ix = i % n1, iy = (i / n1) % n1, iz = (dim == 3) ? i / ((long long)n1 * n1) : 0;
            if (geo.on_boundary(ix, iy, iz))
              dst[b * stride + i] = src[b * stride + i];
          }
      }
  }
} // namespace

struct spirk_ctx
{
  std::vector<double>                               scratch;
  long long                                         launches = 0;
  std::chrono::steady_clock::time_point             t0;
};
struct spirk_comm
{
  int rank, n_ranks;
};
// TEST-ONLY multi-rank hook: the CPU double has no transport of its own; a test process may register
// callbacks (e.g. torch.distributed / gloo from Python) that carry the collectives, so that the
// stage-parallel host logic can be exercised with world_size > 1 on CPU.
typedef void (*spirk_cpu_allreduce_fn)(double *buf, long long n);
typedef void (*spirk_cpu_allgather_fn)(double *recv, const double *send, long long n);
namespace
{
  int                    g_cpu_rank = 0, g_cpu_size = 1;
  spirk_cpu_allreduce_fn g_cpu_allreduce = nullptr;
  spirk_cpu_allgather_fn g_cpu_allgather = nullptr;
  thread_local spirk_comm *g_reduction_comm = nullptr;
  inline void reduce_scalar(double *r)
  {
    if (g_reduction_comm && g_reduction_comm->n_ranks > 1 && g_cpu_allreduce)
      g_cpu_allreduce(r, 1);
  }
} // namespace

extern "C" {

const char *spirk_backend(void) { return "cpu-oracle"; }
const char *spirk_last_error(void) { return g_err.c_str(); }

int spirk_ctx_create(spirk_ctx **ctx, int)
{
  *ctx = new spirk_ctx();
  return SPIRK_OK;
}
int spirk_ctx_destroy(spirk_ctx *ctx)
{
  delete ctx;
  return SPIRK_OK;
}
int       spirk_ctx_sync(spirk_ctx *) { return SPIRK_OK; }
long long spirk_ctx_launch_count(spirk_ctx *ctx) { return ctx->launches; }
int       spirk_ctx_timer_begin(spirk_ctx *ctx)
{
  ctx->t0 = std::chrono::steady_clock::now();
  return SPIRK_OK;
}
int spirk_ctx_timer_end(spirk_ctx *ctx, double *ms)
{
  *ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - ctx->t0).count();
  return SPIRK_OK;
}
int spirk_ctx_set_option(spirk_ctx *, const char *, int) { return SPIRK_OK; }

struct spirk_graph
{
  int dummy;
};
int spirk_graph_begin(spirk_ctx *) { return fail(SPIRK_ERR_UNSUPPORTED, "no graph capture on the CPU double"); }
int spirk_graph_end(spirk_ctx *, spirk_graph **) { return fail(SPIRK_ERR_UNSUPPORTED, "no graph capture"); }
int spirk_graph_launch(spirk_ctx *, spirk_graph *) { return fail(SPIRK_ERR_UNSUPPORTED, "no graph capture"); }
int spirk_graph_destroy(spirk_graph *) { return SPIRK_OK; }

int spirk_malloc(spirk_ctx *, double **ptr, size_t n)
{
  *ptr = (double *)std::calloc(n ? n : 1, sizeof(double));
  return *ptr ? SPIRK_OK : fail(SPIRK_ERR_NOMEM, "calloc");
}
int spirk_free(spirk_ctx *, double *ptr)
{
  std::free(ptr);
  return SPIRK_OK;
}
int spirk_copy_h2d(spirk_ctx *, double *dst, const double *src, size_t n)
{
  std::memcpy(dst, src, n * sizeof(double));
  return SPIRK_OK;
}
int spirk_copy_d2h(spirk_ctx *, double *dst, const double *src, size_t n)
{
  std::memcpy(dst, src, n * sizeof(double));
  return SPIRK_OK;
}
int spirk_malloc_host(spirk_ctx *c, double **p, size_t n) { return spirk_malloc(c, p, n); }
int spirk_free_host(spirk_ctx *c, double *p) { return spirk_free(c, p); }

long long spirk_level_n_dofs(const spirk_level *lvl) { return Geo(lvl).N; }

int spirk_op_apply(spirk_ctx *ctx, const spirk_level *lvl, const spirk_opdesc *op, double *dst, const double *src,
                   long long stride)
{
  if (int e = check_level(lvl))
    return e;
  if (!op || op->nb < 1 || op->nb > SPIRK_MAX_BLOCKS)
    return fail(SPIRK_ERR_INVALID, "bad opdesc");
  if (dst == src)
    return fail(SPIRK_ERR_INVALID, "dst aliases src");
  ctx->launches++;
  cell_loop(lvl, op, dst, src, stride);
  return SPIRK_OK;
}

int spirk_op_fuses_own_diagonal(spirk_ctx *, const spirk_level *, const spirk_opdesc *) { return 0; }

int spirk_op_apply_km(spirk_ctx *ctx, const spirk_level *lvl, int nb, double *dst, const double *v, const double *w, long long stride,
                      const double *laplace, const double *mass)
{
  if (int e = check_level(lvl))
    return e;
  const Geo geo(lvl);
  spirk_opdesc dk, dm;
  std::memset(&dk, 0, sizeof(dk)), std::memset(&dm, 0, sizeof(dm));
  dk.kind = dm.kind = SPIRK_OP_REAL, dk.nb = dm.nb = nb;
  for (int b = 0; b < nb; ++b)
    dk.laplace[b] = laplace[b], dm.mass[b] = mass[b];
  if (int e = spirk_op_apply(ctx, lvl, &dk, dst, v, stride))
    return e;
  std::vector<double> tmp((size_t)stride * nb);
  if (int e = spirk_op_apply(ctx, lvl, &dm, tmp.data(), w, stride))
    return e;
  for (int b = 0; b < nb; ++b)
    for (long long i = 0; i < geo.N; ++i)
      {
        const int ix = i % geo.n1, iy = (i / geo.n1) % geo.n1, iz = (geo.dim == 3) ? i / ((long long)geo.n1 * geo.n1) : 1;
        if (!geo.on_boundary(ix, iy, geo.dim == 3 ? iz : 1))
          dst[b * stride + i] += tmp[b * stride + i];
      }
  return SPIRK_OK;
}

int spirk_op_residual(spirk_ctx *ctx, const spirk_level *lvl, const spirk_opdesc *op, double *dst, const double *rhs,
                      const double *src, long long stride)
{
  const long long N = Geo(lvl).N;
  ctx->scratch.resize((size_t)op->nb * N);
  if (int e = spirk_op_apply(ctx, lvl, op, ctx->scratch.data(), src, N))
    return e;
  for (int b = 0; b < op->nb; ++b)
    {
      const double *t = ctx->scratch.data() + b * N;
#pragma omp parallel for
      for (long long i = 0; i < N; ++i)
        dst[b * stride + i] = rhs[b * stride + i] - t[i];
    }
  return SPIRK_OK;
}

static int cheb_step_impl(spirk_ctx *ctx, const spirk_level *lvl, const spirk_opdesc *op, double *x_new, const double *x,
                          const double *x_old, const double *rhs, const double *dinv, const double *diag_mass,
                          const double *diag_laplace, long long stride, const double *f1, const double *f2)
{
  const long long N = Geo(lvl).N;
  ctx->scratch.resize((size_t)op->nb * N);
  if (int e = spirk_op_apply(ctx, lvl, op, ctx->scratch.data(), x, N))
    return e;
  std::vector<double> own; // dinv == NULL: the operator's own inverse diagonal
  if (!dinv)
    {
      // (COUPLED: the diagonal of block b's own term, coupling[b][b] M + laplace[b] K)
      own.resize((size_t)op->nb * stride);
      for (int b = 0; b < op->nb; ++b)
        spirk_op_inverse_diagonal(ctx, lvl, own.data() + b * stride,
                                  diag_mass ? diag_mass[b] : (op->kind == SPIRK_OP_REAL ? op->mass[b] : op->coupling[b * op->nb + b]),
                                  diag_laplace ? diag_laplace[b] : op->laplace[b]);
      dinv = own.data();
    }
  for (int b = 0; b < op->nb; ++b)
    {
      const double *t = ctx->scratch.data() + b * N;
      const double  a1 = f1[b], a2 = f2[b];
#pragma omp parallel for
      for (long long i = 0; i < N; ++i)
        {
          const long long j  = b * stride + i;
          const double    xo = x_old ? x_old[j] : 0.0;
          // deal.II VectorUpdater: factor1_plus_1 * x - factor1 * x_old + factor2 * dinv * (rhs - Ax)
          x_new[j] = (1.0 + a1) * x[j] - a1 * xo + a2 * dinv[j] * (rhs[j] - t[i]);
        }
    }
  return SPIRK_OK;
}

// first two Chebyshev iterates from a zero start: x1 = f0 dinv rhs, x2 = x1 + f1 x1 + f2 dinv (rhs - A x1)
int spirk_op_cheb_step(spirk_ctx *ctx, const spirk_level *lvl, const spirk_opdesc *op, double *x_new, const double *x,
                       const double *x_old, const double *rhs, const double *dinv, long long stride, const double *f1,
                       const double *f2)
{
  return cheb_step_impl(ctx, lvl, op, x_new, x, x_old, rhs, dinv, nullptr, nullptr, stride, f1, f2);
}

int spirk_op_cheb_step_diag(spirk_ctx *ctx, const spirk_level *lvl, const spirk_opdesc *op, double *x_new, const double *x,
                            const double *x_old, const double *rhs, const double *diag_mass, const double *diag_laplace,
                            long long stride, const double *f1, const double *f2)
{
  if (!diag_mass || !diag_laplace)
    return fail(SPIRK_ERR_INVALID, "cheb_step_diag: the coefficients of the diagonal are required");
  return cheb_step_impl(ctx, lvl, op, x_new, x, x_old, rhs, nullptr, diag_mass, diag_laplace, stride, f1, f2);
}

static int cheb_first_impl(spirk_ctx *ctx, const spirk_level *lvl, const spirk_opdesc *op, double *x1, double *x2, const double *rhs,
                           const double *diag_mass, const double *diag_laplace, long long stride, const double *f0, const double *f1,
                           const double *f2)
{
  const long long N = Geo(lvl).N;
  for (int b = 0; b < op->nb; ++b)
    {
      if (int e = spirk_op_inverse_diagonal(ctx, lvl, x2 + b * stride,
                                            diag_mass ? diag_mass[b] : (op->kind == SPIRK_OP_REAL ? op->mass[b] : op->coupling[b * op->nb + b]),
                                            diag_laplace ? diag_laplace[b] : op->laplace[b]))
        return e;
      for (long long i = 0; i < N; ++i)
        x1[b * stride + i] = f0[b] * x2[b * stride + i] * rhs[b * stride + i];
    }
  return cheb_step_impl(ctx, lvl, op, x2, x1, nullptr, rhs, nullptr, diag_mass, diag_laplace, stride, f1, f2);
}

int spirk_op_cheb_first(spirk_ctx *ctx, const spirk_level *lvl, const spirk_opdesc *op, double *x1, double *x2, const double *rhs,
                        long long stride, const double *f0, const double *f1, const double *f2)
{
  return cheb_first_impl(ctx, lvl, op, x1, x2, rhs, nullptr, nullptr, stride, f0, f1, f2);
}

int spirk_op_cheb_first_diag(spirk_ctx *ctx, const spirk_level *lvl, const spirk_opdesc *op, double *x1, double *x2, const double *rhs,
                             const double *diag_mass, const double *diag_laplace, long long stride, const double *f0, const double *f1,
                             const double *f2)
{
  if (!diag_mass || !diag_laplace)
    return fail(SPIRK_ERR_INVALID, "cheb_first_diag: the coefficients of the diagonal are required");
  return cheb_first_impl(ctx, lvl, op, x1, x2, rhs, diag_mass, diag_laplace, stride, f0, f1, f2);
}

int spirk_op_inverse_diagonal(spirk_ctx *, const spirk_level *lvl, double *diag, double mass, double laplace)
{
  if (int e = check_level(lvl))
    return e;
  const Geo geo(lvl);
  const int dim = geo.dim, n = geo.n, k = geo.k, nc = geo.nc, n1 = geo.n1;
  int       nl  = 1;
  for (int d = 0; d < dim; ++d)
    nl *= n;
  // MatrixFreeTools::compute_diagonal: cell integral applied to unit vectors (A4)
  CellKernel          ck(geo);
  std::vector<double> e(nl), um(nl), out(nl), dloc(nl);
  for (int i = 0; i < nl; ++i)
    {
      std::fill(e.begin(), e.end(), 0.0);
      e[i] = 1.0;
      for (int l = 0; l < nl; ++l)
        um[l] = mass * e[l];
      ck.run(e.data(), mass != 0.0 ? um.data() : nullptr, laplace, out.data());
      dloc[i] = out[i];
    }
  for (long long i = 0; i < geo.N; ++i)
    diag[i] = 0.0;
  const int ncy = nc, ncz = (dim == 3) ? nc : 1;
  for (int cz = 0; cz < ncz; ++cz)
    for (int cy = 0; cy < ncy; ++cy)
      for (int cx = 0; cx < nc; ++cx)
        for (int l = 0; l < nl; ++l)
          {
            const int ix = cx * k + l % n, iy = cy * k + (l / n) % n, iz = (dim == 3) ? cz * k + l / (n * n) : 0;
            if (!geo.on_boundary(ix, iy, iz))
              diag[ix + (long long)n1 * (iy + (long long)n1 * iz)] += dloc[l];
          }
  for (long long i = 0; i < geo.N; ++i)
    diag[i] = (std::abs(diag[i]) > 1.0e-10) ? (1.0 / diag[i]) : 1.0;
  return SPIRK_OK;
}

int spirk_op_assemble_dense(spirk_ctx *ctx, const spirk_level *lvl, double mass, double laplace, double *host_matrix)
{
  const long long N = Geo(lvl).N;
  if (N > 4096)
    return fail(SPIRK_ERR_INVALID, "assemble_dense: level too large");
  spirk_opdesc op{};
  op.kind = SPIRK_OP_REAL, op.nb = 1, op.mass[0] = mass, op.laplace[0] = laplace;
  std::vector<double> e(N), col(N);
  for (long long j = 0; j < N; ++j)
    {
      std::fill(e.begin(), e.end(), 0.0);
      e[j] = 1.0;
      spirk_op_apply(ctx, lvl, &op, col.data(), e.data(), N);
      for (long long i = 0; i < N; ++i)
        host_matrix[i * N + j] = col[i];
    }
  return SPIRK_OK;
}

// ---------------------------------------------------------------- transfer (A9)
int spirk_mg_prolongate_add(spirk_ctx *, const spirk_level *lf, int nb, double *fine, long long fs, const double *coarse,
                            long long cs)
{
  if (int e = check_level(lf))
    return e;
  if (lf->n_cells_1d % 2)
    return fail(SPIRK_ERR_INVALID, "fine level must have an even number of cells");
  spirk_level lc = *lf;
  lc.n_cells_1d /= 2;
  const Geo gf(lf), gc(&lc);
  const FE &fe  = get_fe(gf.k);
  const int dim = gf.dim, n = gf.n, k = gf.k, m = 2 * k + 1, ncc = gc.nc;
  int       nlc = 1, nlf = 1;
  for (int d = 0; d < dim; ++d)
    nlc *= n, nlf *= m;
  const int nccz = (dim == 3) ? ncc : 1;
#pragma omp parallel
  {
    std::vector<double> u(nlc), t0(nlf), t1(nlf);
#pragma omp for collapse(2)
    for (int b = 0; b < nb; ++b)
      for (int cz = 0; cz < nccz; ++cz)
        for (int cy = 0; cy < ncc; ++cy)
          for (int cx = 0; cx < ncc; ++cx)
            {
              for (int l = 0; l < nlc; ++l)
                {
                  const int ix = cx * k + l % n, iy = cy * k + (l / n) % n, iz = (dim == 3) ? cz * k + l / (n * n) : 0;
                  u[l] = gc.on_boundary(ix, iy, iz) ? 0.0 : coarse[b * cs + ix + (long long)gc.n1 * (iy + (long long)gc.n1 * iz)];
                }
              int           ext[3] = {n, n, n};
              const double *in     = u.data();
              double       *bufs[2] = {t0.data(), t1.data()};
              double       *o      = nullptr;
              for (int d = 0; d < dim; ++d)
                {
                  o = bufs[d % 2];
                  apply1d(fe.P.data(), m, n, false, d, dim, ext, in, o);
                  ext[d] = m;
                  in     = o;
                }
              // each fine DoF written once: cell owns local fine indices [0, 2k), plus 2k at the far end
              const int mx = (cx == ncc - 1) ? m : m - 1, my = (cy == ncc - 1) ? m : m - 1,
                        mz = (dim == 3) ? ((cz == ncc - 1) ? m : m - 1) : 1;
              for (int jz = 0; jz < mz; ++jz)
                for (int jy = 0; jy < my; ++jy)
                  for (int jx = 0; jx < mx; ++jx)
                    {
                      const int fx = cx * 2 * k + jx, fy = cy * 2 * k + jy, fz = (dim == 3) ? cz * 2 * k + jz : 0;
                      fine[b * fs + fx + (long long)gf.n1 * (fy + (long long)gf.n1 * fz)] += o[jx + m * (jy + m * jz)];
                    }
            }
  }
  return SPIRK_OK;
}

int spirk_mg_restrict(spirk_ctx *, const spirk_level *lf, int nb, double *coarse, long long cs, const double *fine,
                      long long fs)
{
  if (int e = check_level(lf))
    return e;
  if (lf->n_cells_1d % 2)
    return fail(SPIRK_ERR_INVALID, "fine level must have an even number of cells");
  spirk_level lc = *lf;
  lc.n_cells_1d /= 2;
  const Geo gf(lf), gc(&lc);
  const FE &fe  = get_fe(gf.k);
  const int dim = gf.dim, n = gf.n, k = gf.k, m = 2 * k + 1, ncc = gc.nc;
  int       nlc = 1, nlf = 1;
  for (int d = 0; d < dim; ++d)
    nlc *= n, nlf *= m;
  const int nccz = (dim == 3) ? ncc : 1;
  for (int b = 0; b < nb; ++b)
    for (long long i = 0; i < gc.N; ++i)
      coarse[b * cs + i] = 0.0;
  std::vector<double> u(nlf), t0(nlf), t1(nlf);
  for (int b = 0; b < nb; ++b)
    for (int cz = 0; cz < nccz; ++cz)
      for (int cy = 0; cy < ncc; ++cy)
        for (int cx = 0; cx < ncc; ++cx)
          {
            // transpose of "write once": a shared fine DoF contributes through its owner cell only
            const int mx = (cx == ncc - 1) ? m : m - 1, my = (cy == ncc - 1) ? m : m - 1,
                      mz = (dim == 3) ? ((cz == ncc - 1) ? m : m - 1) : 1;
            std::fill(u.begin(), u.end(), 0.0);
            for (int jz = 0; jz < mz; ++jz)
              for (int jy = 0; jy < my; ++jy)
                for (int jx = 0; jx < mx; ++jx)
                  {
                    const int fx = cx * 2 * k + jx, fy = cy * 2 * k + jy, fz = (dim == 3) ? cz * 2 * k + jz : 0;
                    u[jx + m * (jy + m * jz)] = fine[b * fs + fx + (long long)gf.n1 * (fy + (long long)gf.n1 * fz)];
                  }
            int           ext[3] = {m, m, m};
            const double *in     = u.data();
            double       *bufs[2] = {t0.data(), t1.data()};
            double       *o      = nullptr;
            for (int d = 0; d < dim; ++d)
              {
                o = bufs[d % 2];
                apply1d(fe.P.data(), n, m, true, d, dim, ext, in, o);
                ext[d] = n;
                in     = o;
              }
            for (int l = 0; l < nlc; ++l)
              {
                const int ix = cx * k + l % n, iy = cy * k + (l / n) % n, iz = (dim == 3) ? cz * k + l / (n * n) : 0;
                if (!gc.on_boundary(ix, iy, iz))
                  coarse[b * cs + ix + (long long)gc.n1 * (iy + (long long)gc.n1 * iz)] += o[l];
              }
          }
  return SPIRK_OK;
}

int spirk_dense_matvec(spirk_ctx *, int n, int nb, double *y, const double *x, long long stride, const double *matrix,
                       long long matrix_stride)
{
  for (int b = 0; b < nb; ++b)
    for (int i = 0; i < n; ++i)
      {
        double s = 0;
        for (int j = 0; j < n; ++j)
          s += matrix[b * matrix_stride + (size_t)i * n + j] * x[b * stride + j];
        y[b * stride + i] = s;
      }
  return SPIRK_OK;
}

// ---------------------------------------------------------------- vector kernels
#define PFOR _Pragma("omp parallel for")
int spirk_vec_set(spirk_ctx *, double *x, long long n, double v)
{
  PFOR for (long long i = 0; i < n; ++i) x[i] = v;
  return SPIRK_OK;
}
int spirk_vec_copy(spirk_ctx *, double *d, const double *s, long long n)
{
  PFOR for (long long i = 0; i < n; ++i) d[i] = s[i];
  return SPIRK_OK;
}
int spirk_vec_scale(spirk_ctx *, double *x, long long n, double a)
{
  PFOR for (long long i = 0; i < n; ++i) x[i] *= a;
  return SPIRK_OK;
}
int spirk_vec_axpy(spirk_ctx *, double *y, double a, const double *x, long long n)
{
  PFOR for (long long i = 0; i < n; ++i) y[i] += a * x[i];
  return SPIRK_OK;
}
int spirk_vec_sadd(spirk_ctx *, double *y, double s, double a, const double *x, long long n)
{
  PFOR for (long long i = 0; i < n; ++i) y[i] = s * y[i] + a * x[i];
  return SPIRK_OK;
}
int spirk_vec_add2(spirk_ctx *, double *y, double a, const double *x, double b, const double *z, long long n)
{
  PFOR for (long long i = 0; i < n; ++i) y[i] += a * x[i] + b * z[i];
  return SPIRK_OK;
}
int spirk_vec_equ(spirk_ctx *, double *y, double a, const double *x, long long n)
{
  PFOR for (long long i = 0; i < n; ++i) y[i] = a * x[i];
  return SPIRK_OK;
}
int spirk_vec_scale_pointwise(spirk_ctx *, int nb, long long n, double *y, const double *d, const double *x,
                              long long stride, const double *f)
{
  for (int b = 0; b < nb; ++b)
    {
      const double a = f[b];
      PFOR for (long long i = 0; i < n; ++i) y[b * stride + i] = a * (d[b * stride + i] * x[b * stride + i]);
    }
  return SPIRK_OK;
}
int spirk_vec_dot(spirk_ctx *, const double *x, const double *y, long long n, double *r)
{
  double s = 0;
#pragma omp parallel for reduction(+ : s)
  for (long long i = 0; i < n; ++i)
    s += x[i] * y[i];
  *r = s;
  reduce_scalar(r);
  return SPIRK_OK;
}
int spirk_vec_add_and_dot(spirk_ctx *, double *v, double a, const double *V, const double *W, long long n, double *r)
{
  double s = 0;
#pragma omp parallel for reduction(+ : s)
  for (long long i = 0; i < n; ++i)
    {
      v[i] += a * V[i];
      s += v[i] * W[i];
    }
  *r = s;
  reduce_scalar(r);
  return SPIRK_OK;
}
int spirk_vec_sum(spirk_ctx *, const double *x, long long n, double *r)
{
  double s = 0;
#pragma omp parallel for reduction(+ : s)
  for (long long i = 0; i < n; ++i)
    s += x[i];
  *r = s;
  return SPIRK_OK;
}
// block vectors with a stride larger than the block length (z-slab vectors of the CUDA library): plain loops here
int spirk_vec_dot_strided(spirk_ctx *ctx, const double *x, const double *y, long long n, int nb, long long stride, double *r)
{
  double s = 0;
  for (int b = 0; b < nb; ++b)
    {
      double t = 0;
      if (int e = spirk_vec_dot(ctx, x + b * stride, y + b * stride, n, &t))
        return e;
      s += t;
    }
  *r = s;
  return SPIRK_OK;
}
int spirk_vec_sum_strided(spirk_ctx *ctx, const double *x, long long n, int nb, long long stride, double *r)
{
  double s = 0;
  for (int b = 0; b < nb; ++b)
    {
      double t = 0;
      if (int e = spirk_vec_sum(ctx, x + b * stride, n, &t))
        return e;
      s += t;
    }
  *r = s;
  return SPIRK_OK;
}
int spirk_vec_add_and_dot_strided(spirk_ctx *ctx, double *v, double a, const double *V, const double *W, long long n, int nb,
                                  long long stride, double *r)
{
  double s = 0;
  for (int b = 0; b < nb; ++b)
    {
      double t = 0;
      if (int e = spirk_vec_add_and_dot(ctx, v + b * stride, a, V + b * stride, (W == v ? v : W) + b * stride, n, &t))
        return e;
      s += t;
    }
  *r = s;
  return SPIRK_OK;
}
int spirk_gmres_mgs_strided(spirk_ctx *ctx, double *vv, const double *const *basis, int dim, long long n, int nb, long long stride,
                            double *h, double *norm)
{
  if (int e = spirk_vec_dot_strided(ctx, vv, basis[0], n, nb, stride, &h[0]))
    return e;
  for (int i = 1; i < dim; ++i)
    if (int e = spirk_vec_add_and_dot_strided(ctx, vv, -h[i - 1], basis[i - 1], basis[i], n, nb, stride, &h[i]))
      return e;
  double s = 0;
  if (int e = spirk_vec_add_and_dot_strided(ctx, vv, -h[dim - 1], basis[dim - 1], vv, n, nb, stride, &s))
    return e;
  *norm = std::sqrt(s);
  return SPIRK_OK;
}
int spirk_gmres_mgs(spirk_ctx *ctx, double *vv, const double *const *basis, int dim, long long n, double *h, double *norm)
{
  spirk_vec_dot(ctx, vv, basis[0], n, &h[0]);
  for (int i = 1; i < dim; ++i)
    spirk_vec_add_and_dot(ctx, vv, -h[i - 1], basis[i - 1], basis[i], n, &h[i]);
  double s;
  // vv.add_and_dot(-h(dim-1), q_{dim-1}, vv)
  spirk_vec_axpy(ctx, vv, -h[dim - 1], basis[dim - 1], n);
  spirk_vec_dot(ctx, vv, vv, n, &s);
  *norm = std::sqrt(s);
  return SPIRK_OK;
}
int spirk_mix(spirk_ctx *, int qo, int qi, double *dst, long long ds, const double *src, long long ss, long long n,
              const double *T, int add, double cutoff)
{
  if (qo > SPIRK_MAX_BLOCKS || qi > SPIRK_MAX_BLOCKS)
    return fail(SPIRK_ERR_INVALID, "mix: too many blocks");
#pragma omp parallel for
  for (long long e = 0; e < n; ++e)
    {
      double in[SPIRK_MAX_BLOCKS];
      for (int j = 0; j < qi; ++j)
        in[j] = src[j * ss + e];
      for (int i = 0; i < qo; ++i)
        {
          double t = 0;
          for (int j = 0; j < qi; ++j)
            if (std::abs(T[i * qi + j]) > cutoff)
              t += T[i * qi + j] * in[j];
          if (add)
            dst[i * ds + e] += t;
          else
            dst[i * ds + e] = t;
        }
    }
  return SPIRK_OK;
}

// ---------------------------------------------------------------- problem pieces
int spirk_problem_rhs_spatial(spirk_ctx *, const spirk_level *lvl, double *r)
{
  if (int e = check_level(lvl))
    return e;
  const Geo geo(lvl);
  const FE &fe = get_fe(geo.k);
  const int n = geo.n, k = geo.k, nc = geo.nc, n1 = geo.n1, dim = geo.dim;
  // separable: r = r1 (x) r1 (x) r1, r1_i = sum_cells sum_q phi_i(x_q) sin(2 pi x_q) w_q h
  std::vector<double> r1(n1, 0.0);
  for (int c = 0; c < nc; ++c)
    for (int q = 0; q < n; ++q)
      {
        const double s = std::sin(2.0 * M_PI * (c + fe.xq[q]) * geo.h) * fe.wq[q] * geo.h;
        for (int i = 0; i < n; ++i)
          r1[c * k + i] += fe.B[q * n + i] * s;
      }
  for (long long i = 0; i < geo.N; ++i)
    {
      const int ix = i % n1, iy = (i / n1) % n1, iz = (dim == 3) ? i / ((long long)n1 * n1) : 0;
      r[i] = geo.on_boundary(ix, iy, iz) ? 0.0 : r1[ix] * r1[iy] * (dim == 3 ? r1[iz] : 1.0);
    }
  return SPIRK_OK;
}

static double node_coord(const Geo &geo, const FE &fe, int i)
{
  const int c = std::min(i / geo.k, geo.nc - 1);
  return (c + fe.nodes[i - c * geo.k]) * geo.h;
}

int spirk_problem_interpolate_solution(spirk_ctx *, const spirk_level *lvl, double *u, double t)
{
  if (int e = check_level(lvl))
    return e;
  const Geo    geo(lvl);
  const FE    &fe = get_fe(geo.k);
  const int    n1 = geo.n1, dim = geo.dim;
  const double ft = (1.0 + std::sin(M_PI * t)) * std::exp(-0.5 * t);
  std::vector<double> s(n1);
  for (int i = 0; i < n1; ++i)
    s[i] = std::sin(2.0 * M_PI * node_coord(geo, fe, i));
  for (long long i = 0; i < geo.N; ++i)
    {
      const int ix = i % n1, iy = (i / n1) % n1, iz = (dim == 3) ? i / ((long long)n1 * n1) : 0;
      u[i]         = s[ix] * s[iy] * (dim == 3 ? s[iz] : 1.0) * ft;
    }
  return SPIRK_OK;
}

int spirk_problem_error_norms(spirk_ctx *, const spirk_level *lvl, const double *u, double t, double *l2, double *linf);
int spirk_problem_error_norms_partial(spirk_ctx *ctx, const spirk_level *lvl, const double *u, double t, double *l2sq, double *linf)
{
  double l2 = 0;
  if (int e = spirk_problem_error_norms(ctx, lvl, u, t, &l2, linf))
    return e;
  *l2sq = l2 * l2;
  return SPIRK_OK;
}
int spirk_problem_error_norms(spirk_ctx *, const spirk_level *lvl, const double *u, double t, double *l2, double *linf)
{
  if (int e = check_level(lvl))
    return e;
  const Geo    geo(lvl);
  const FE    &fe = get_fe(geo.k);
  const int    n = geo.n, k = geo.k, nc = geo.nc, n1 = geo.n1, dim = geo.dim, ne = k + 2;
  const double ft = (1.0 + std::sin(M_PI * t)) * std::exp(-0.5 * t);
  int          nl = 1, nq = 1;
  for (int d = 0; d < dim; ++d)
    nl *= n, nq *= ne;
  double    sum = 0, mx = 0;
  const int ncz = (dim == 3) ? nc : 1;
#pragma omp parallel
  {
    std::vector<double> ul(nl), t0(nq), t1(nq);
    double              lsum = 0, lmx = 0;
#pragma omp for collapse(2)
    for (int cz = 0; cz < ncz; ++cz)
      for (int cy = 0; cy < nc; ++cy)
        for (int cx = 0; cx < nc; ++cx)
          {
            for (int l = 0; l < nl; ++l)
              {
                const int ix = cx * k + l % n, iy = cy * k + (l / n) % n, iz = (dim == 3) ? cz * k + l / (n * n) : 0;
                ul[l]        = u[ix + (long long)n1 * (iy + (long long)n1 * iz)];
              }
            int           ext[3] = {n, n, n};
            const double *in     = ul.data();
            double       *bufs[2] = {t0.data(), t1.data()};
            double       *o      = nullptr;
            for (int d = 0; d < dim; ++d)
              {
                o = bufs[d % 2];
                apply1d(fe.Be.data(), ne, n, false, d, dim, ext, in, o);
                ext[d] = ne;
                in     = o;
              }
            for (int q = 0; q < nq; ++q)
              {
                const int    qx = q % ne, qy = (q / ne) % ne, qz = (dim == 3) ? q / (ne * ne) : 0;
                const double x = (cx + fe.xe[qx]) * geo.h, y = (cy + fe.xe[qy]) * geo.h, z = (cz + fe.xe[qz]) * geo.h;
                double       ex = std::sin(2 * M_PI * x) * std::sin(2 * M_PI * y) * ft, w = fe.we[qx] * fe.we[qy] * geo.h * geo.h;
                if (dim == 3)
                  ex *= std::sin(2 * M_PI * z), w *= fe.we[qz] * geo.h;
                const double d = o[q] - ex;
                lsum += w * d * d;
                lmx = std::max(lmx, std::abs(d));
              }
          }
#pragma omp critical
    {
      sum += lsum;
      mx = std::max(mx, lmx);
    }
  }
  *l2   = std::sqrt(sum);
  *linf = mx;
  return SPIRK_OK;
}

int spirk_constraints_set_zero(spirk_ctx *, const spirk_level *lvl, int nb, double *u, long long stride)
{
  const Geo geo(lvl);
  const int n1 = geo.n1, dim = geo.dim;
  for (int b = 0; b < nb; ++b)
    {
#pragma omp parallel for
      for (long long i = 0; i < geo.N; ++i)
        {
          const int ix = i % n1, iy = (i / n1) % n1, iz = (dim == 3) ? i / ((long long)n1 * n1) : 0;
          if (geo.on_boundary(ix, iy, iz))
            u[b * stride + i] = 0.0;
        }
    }
  return SPIRK_OK;
}

// ---------------------------------------------------------------- communication: single rank, or
// test-registered callbacks (see spirk_cpu_set_comm above)
int spirk_cpu_set_comm(int rank, int size, spirk_cpu_allreduce_fn allreduce, spirk_cpu_allgather_fn allgather)
{
  g_cpu_rank = rank, g_cpu_size = size, g_cpu_allreduce = allreduce, g_cpu_allgather = allgather;
  return SPIRK_OK;
}
int spirk_comm_unique_id(char *id)
{
  std::memset(id, 0, 128);
  return SPIRK_OK;
}
int spirk_comm_create(spirk_ctx *, const char *, int n_ranks, int rank, spirk_comm **comm)
{
  if (n_ranks != 1 && !(g_cpu_allreduce && g_cpu_allgather && n_ranks == g_cpu_size && rank == g_cpu_rank))
    return fail(SPIRK_ERR_UNSUPPORTED, "cpu oracle: multi-rank needs spirk_cpu_set_comm callbacks");
  *comm = new spirk_comm{rank, n_ranks};
  return SPIRK_OK;
}
int spirk_comm_destroy(spirk_comm *c)
{
  delete c;
  return SPIRK_OK;
}
int spirk_comm_rank(const spirk_comm *c, int *rank, int *n)
{
  *rank = c->rank, *n = c->n_ranks;
  return SPIRK_OK;
}
int spirk_comm_allreduce_sum(spirk_ctx *, spirk_comm *c, double *buf, long long n)
{
  if (c->n_ranks > 1)
    g_cpu_allreduce(buf, n);
  return SPIRK_OK;
}
// the CPU double has no spatial partition (z-slabs are a feature of the CUDA library): a split keeps the whole group
int spirk_comm_split(spirk_ctx *, spirk_comm *c, int, int, spirk_comm **out)
{
  if (c->n_ranks > 1)
    return fail(SPIRK_ERR_UNSUPPORTED, "comm_split: not available on the CPU double");
  *out = new spirk_comm{0, 1};
  return SPIRK_OK;
}
int spirk_comm_allreduce_max(spirk_ctx *, spirk_comm *c, double *, long long)
{
  return c->n_ranks > 1 ? fail(SPIRK_ERR_UNSUPPORTED, "allreduce_max: not available on the CPU double") : SPIRK_OK;
}
int spirk_halo_exchange(spirk_ctx *, spirk_comm *, const spirk_level *lvl, int, double *, long long, int, int)
{
  return lvl->slab ? fail(SPIRK_ERR_UNSUPPORTED, "z-slab levels are not available on the CPU double") : SPIRK_OK;
}
int spirk_comm_allgather(spirk_ctx *, spirk_comm *c, double *recv, const double *send, long long n)
{
  if (c->n_ranks > 1)
    g_cpu_allgather(recv, send, n);
  else if (recv != send)
    std::memcpy(recv, send, n * sizeof(double));
  return SPIRK_OK;
}
struct spirk_xbuf
{
  std::vector<double>       local; // [0, n): published blocks, [n, 2n): result region (all-to-all)
  long long                 n = 0;
  int                       rank = 0, n_ranks = 1;
  std::vector<spirk_xbuf *> group; // same-process virtual group (spirk_xbuf_create_virtual_group)
};
int spirk_comm_xbuf_create(spirk_ctx *, spirk_comm *c, long long n, spirk_xbuf **out)
{
  *out = new spirk_xbuf();
  (*out)->local.assign((size_t)2 * n, 0.0);
  (*out)->n = n, (*out)->rank = c->rank, (*out)->n_ranks = c->n_ranks;
  return SPIRK_OK;
}
int spirk_xbuf_create_virtual_group(spirk_ctx *, int n_ranks, long long n, spirk_xbuf **out)
{
  if (n_ranks < 1 || n_ranks > SPIRK_MAX_BLOCKS || n < 1)
    return fail(SPIRK_ERR_INVALID, "xbuf virtual group: n_ranks / n");
  std::vector<spirk_xbuf *> xs(n_ranks);
  for (int r = 0; r < n_ranks; ++r)
    {
      xs[r] = new spirk_xbuf();
      xs[r]->local.assign((size_t)2 * n, 0.0);
      xs[r]->n = n, xs[r]->rank = r, xs[r]->n_ranks = n_ranks;
    }
  for (int r = 0; r < n_ranks; ++r)
    xs[r]->group = xs, out[r] = xs[r];
  return SPIRK_OK;
}
int spirk_comm_xbuf_destroy(spirk_ctx *, spirk_xbuf *x)
{
  delete x;
  return SPIRK_OK;
}
double *spirk_comm_xbuf_local(spirk_xbuf *x) { return x->local.data(); }
// the blocks of all ranks, gathered: through the registered callback (real ranks) or from the virtual group
static int xbuf_gather(spirk_ctx *ctx, spirk_comm *c, spirk_xbuf *x, int m, long long n, std::vector<double> &all)
{
  all.assign((size_t)x->n_ranks * m * n, 0.0);
  if (!x->group.empty())
    {
      for (int r = 0; r < x->n_ranks; ++r)
        std::memcpy(&all[(size_t)r * m * n], x->group[r]->local.data(), (size_t)m * n * sizeof(double));
      return SPIRK_OK;
    }
  if (!c)
    return fail(SPIRK_ERR_INVALID, "mix_peer: a communicator is needed unless the buffer belongs to a virtual group");
  std::vector<double> mine(x->local.begin(), x->local.begin() + (size_t)m * n);
  return spirk_comm_allgather(ctx, c, all.data(), mine.data(), (long long)m * n);
}
// the CPU double has no peer memory: gather, then mix locally
int spirk_mix_peer(spirk_ctx *ctx, spirk_comm *c, spirk_xbuf *x, int qo, int m, double *dst, long long ds, long long n,
                   const double *T, int add, double cutoff)
{
  const int           qi = x->n_ranks * m;
  std::vector<double> all;
  if (int e = xbuf_gather(ctx, c, x, m, n, all))
    return e;
  return spirk_mix(ctx, qo, qi, dst, ds, all.data(), n, n, T, add, cutoff);
}
// all-to-all formulation, virtual groups only (real ranks use spirk_mix_peer_a2a): this rank's chunk of every block for all
// q outputs, written into the owners' result regions
int spirk_mix_peer_a2a_contract(spirk_ctx *ctx, spirk_xbuf *x, int m, long long n, const double *T, double cutoff)
{
  if (x->group.empty())
    return fail(SPIRK_ERR_UNSUPPORTED, "mix_peer_a2a_contract: the CPU double supports it for virtual groups only");
  const int       q  = x->n_ranks * m;
  const long long e0 = n * x->rank / x->n_ranks, e1 = n * (x->rank + 1) / x->n_ranks;
  for (int i = 0; i < q; ++i)
    for (long long e = e0; e < e1; ++e)
      {
        double t = 0.0;
        for (int j = 0; j < q; ++j)
          if (std::fabs(T[i * q + j]) > cutoff)
            t += T[i * q + j] * x->group[j / m]->local[(size_t)(j % m) * n + e];
        x->group[i / m]->local[(size_t)x->n + (size_t)(i % m) * n + e] = t;
      }
  (void)ctx;
  return SPIRK_OK;
}
int spirk_mix_peer_a2a_finish(spirk_ctx *, spirk_xbuf *x, int m, double *dst, long long ds, long long n, int add)
{
  for (int i = 0; i < m; ++i)
    for (long long e = 0; e < n; ++e)
      dst[i * ds + e] = (add ? dst[i * ds + e] : 0.0) + x->local[(size_t)x->n + (size_t)i * n + e];
  return SPIRK_OK;
}
int spirk_mix_peer_a2a(spirk_ctx *ctx, spirk_comm *c, spirk_xbuf *x, int m, double *dst, long long ds, long long n, const double *T,
                       int add, double cutoff)
{
  const int q = c->n_ranks * m;
  return spirk_mix_peer(ctx, c, x, m, m, dst, ds, n, T + (size_t)c->rank * m * q, add, cutoff);
}
int spirk_ctx_set_reduction_comm(spirk_ctx *, spirk_comm *c)
{
  g_reduction_comm = c;
  return SPIRK_OK;
}
}
