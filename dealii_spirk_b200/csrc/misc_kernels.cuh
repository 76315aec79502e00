// Multigrid transfer, vector kernels, reductions, stage mixing and problem-setup kernels.
#pragma once
#include "op_v1.cuh"

namespace spirk
{
#define SPIRK_GRID_STRIDE(i, n) \
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (n); i += (long long)gridDim.x * blockDim.x)

  // =========================================================================================
  // two-level transfer (deal.II MGTwoLevelTransfer set up in preconditioner.h:266-282)
  // One thread block per (coarse cell, vector block); sum-factorised with the (2k+1) x (k+1)
  // 1-D embedding P.  Prolongation writes every fine DoF exactly once (its owner coarse cell),
  // restriction is the exact transpose (atomics into a zeroed coarse vector).
  // =========================================================================================
  template <int K, int DIM>
  __global__ void __launch_bounds__(128)
    k_prolongate_add(const Geo gf, const int nb, double *__restrict__ fine, const long long fs,
                     const double *__restrict__ coarse, const long long cs)
  {
    constexpr int n = K + 1, m = 2 * K + 1;
    constexpr int NC = (DIM == 3) ? n * n * n : n * n;          // coarse local
    constexpr int N1 = (DIM == 3) ? n * n * m : n * m;          // x expanded
    constexpr int N2 = (DIM == 3) ? n * m * m : m * m;          // x,y expanded
    constexpr int N3 = (DIM == 3) ? m * m * m : 1;              // all expanded (3-D only)
    __shared__ double sc[NC], s1[N1], s2[N2];
    const double *P   = c_fe[K].P;
    const int     ncc = gf.nc / 2, n1c = K * ncc + 1, n1f = gf.n1;
    const long long ncells = (DIM == 3) ? (long long)ncc * ncc * ncc : (long long)ncc * ncc;
    for (long long wi = blockIdx.x; wi < ncells * nb; wi += gridDim.x)
      {
        const int       b = wi / ncells;
        const long long c = wi - b * ncells;
        const int       cx = c % ncc, cy = (c / ncc) % ncc, cz = (DIM == 3) ? c / ((long long)ncc * ncc) : 0;
        for (int l = threadIdx.x; l < NC; l += blockDim.x)
          {
            const int  ix = cx * K + l % n, iy = cy * K + (l / n) % n, iz = (DIM == 3) ? cz * K + l / (n * n) : 1;
            const bool bd = on_bdry(ix, n1c) || on_bdry(iy, n1c) || (DIM == 3 && on_bdry(iz, n1c));
            sc[l] = bd ? 0.0 : coarse[b * cs + ix + (long long)n1c * (iy + (DIM == 3 ? (long long)n1c * iz : 0))];
          }
        __syncthreads();
        // expand x: s1[(zy)*m + fx] = sum_i P[fx][i] sc[(zy)*n + i]
        for (int e = threadIdx.x; e < N1; e += blockDim.x)
          {
            const int fx = e % m, zy = e / m;
            double    s  = 0.0;
#pragma unroll
            for (int i = 0; i < n; ++i)
              s = fma(P[fx * n + i], sc[zy * n + i], s);
            s1[e] = s;
          }
        __syncthreads();
        // expand y: s2[(z*m + fy)*m + fx] = sum_j P[fy][j] s1[(z*n + j)*m + fx]
        for (int e = threadIdx.x; e < N2; e += blockDim.x)
          {
            const int fx = e % m, fy = (e / m) % m, z = e / (m * m);
            double    s  = 0.0;
#pragma unroll
            for (int j = 0; j < n; ++j)
              s = fma(P[fy * n + j], s1[(z * n + j) * m + fx], s);
            s2[e] = s;
          }
        __syncthreads();
        const int mx = (cx == ncc - 1) ? m : m - 1, my = (cy == ncc - 1) ? m : m - 1;
        if (DIM == 3)
          {
            const int mz = (cz == ncc - 1) ? m : m - 1;
            for (int e = threadIdx.x; e < N3; e += blockDim.x)
              {
                const int fx = e % m, fy = (e / m) % m, fz = e / (m * m);
                if (fx < mx && fy < my && fz < mz)
                  {
                    double s = 0.0;
#pragma unroll
                    for (int l = 0; l < n; ++l)
                      s = fma(P[fz * n + l], s2[(l * m + fy) * m + fx], s);
                    const long long gi = (cx * 2 * K + fx) + (long long)n1f * ((cy * 2 * K + fy) + (long long)n1f * (cz * 2 * K + fz));
                    fine[b * fs + gi] += s;
                  }
              }
          }
        else
          {
            for (int e = threadIdx.x; e < N2; e += blockDim.x)
              {
                const int fx = e % m, fy = e / m;
                if (fx < mx && fy < my)
                  fine[b * fs + (cx * 2 * K + fx) + (long long)n1f * (cy * 2 * K + fy)] += s2[e];
              }
          }
        __syncthreads();
      }
  }

  template <int K, int DIM>
  __global__ void __launch_bounds__(128)
    k_restrict(const Geo gf, const int nb, double *__restrict__ coarse, const long long cs,
               const double *__restrict__ fine, const long long fs)
  {
    constexpr int n = K + 1, m = 2 * K + 1;
    constexpr int NF = (DIM == 3) ? m * m * m : m * m;          // fine local
    constexpr int N2 = (DIM == 3) ? m * m * n : m * n;          // x contracted
    constexpr int N1 = (DIM == 3) ? m * n * n : n * n;          // x,y contracted
    constexpr int NC = (DIM == 3) ? n * n * n : n * n;
    __shared__ double sf[NF], s2[N2], s1[N1];
    const double *P   = c_fe[K].P;
    const int     ncc = gf.nc / 2, n1c = K * ncc + 1, n1f = gf.n1;
    const long long ncells = (DIM == 3) ? (long long)ncc * ncc * ncc : (long long)ncc * ncc;
    for (long long wi = blockIdx.x; wi < ncells * nb; wi += gridDim.x)
      {
        const int       b = wi / ncells;
        const long long c = wi - b * ncells;
        const int       cx = c % ncc, cy = (c / ncc) % ncc, cz = (DIM == 3) ? c / ((long long)ncc * ncc) : 0;
        const int       mx = (cx == ncc - 1) ? m : m - 1, my = (cy == ncc - 1) ? m : m - 1,
                  mz = (DIM == 3) ? ((cz == ncc - 1) ? m : m - 1) : 1;
        for (int e = threadIdx.x; e < NF; e += blockDim.x)
          {
            const int fx = e % m, fy = (e / m) % m, fz = (DIM == 3) ? e / (m * m) : 0;
            double    v  = 0.0;
            if (fx < mx && fy < my && fz < mz)
              v = fine[b * fs + (cx * 2 * K + fx) + (long long)n1f * ((cy * 2 * K + fy) + (DIM == 3 ? (long long)n1f * (cz * 2 * K + fz) : 0))];
            sf[e] = v;
          }
        __syncthreads();
        // contract x: s2[(zy)*n + i] = sum_fx P[fx][i] sf[(zy)*m + fx]
        for (int e = threadIdx.x; e < N2; e += blockDim.x)
          {
            const int i = e % n, zy = e / n;
            double    s = 0.0;
#pragma unroll
            for (int fx = 0; fx < m; ++fx)
              s = fma(P[fx * n + i], sf[zy * m + fx], s);
            s2[e] = s;
          }
        __syncthreads();
        // contract y: s1[(z*n + j)*n + i] = sum_fy P[fy][j] s2[(z*m + fy)*n + i]
        for (int e = threadIdx.x; e < N1; e += blockDim.x)
          {
            const int i = e % n, j = (e / n) % n, z = e / (n * n);
            double    s = 0.0;
#pragma unroll
            for (int fy = 0; fy < m; ++fy)
              s = fma(P[fy * n + j], s2[(z * m + fy) * n + i], s);
            s1[e] = s;
          }
        __syncthreads();
        for (int e = threadIdx.x; e < NC; e += blockDim.x)
          {
            const int i = e % n, j = (e / n) % n, l = (DIM == 3) ? e / (n * n) : 0;
            double    s;
            if (DIM == 3)
              {
                s = 0.0;
#pragma unroll
                for (int fz = 0; fz < m; ++fz)
                  s = fma(P[fz * n + l], s1[(fz * n + j) * n + i], s);
              }
            else
              s = s1[e];
            const int  ix = cx * K + i, iy = cy * K + j, iz = (DIM == 3) ? cz * K + l : 1;
            const bool bd = on_bdry(ix, n1c) || on_bdry(iy, n1c) || (DIM == 3 && on_bdry(iz, n1c));
            if (!bd)
              atomicAdd(coarse + b * cs + ix + (long long)n1c * (iy + (DIM == 3 ? (long long)n1c * iz : 0)), s);
          }
        __syncthreads();
      }
  }


  // =========================================================================================
  // two-level transfer, owner-computes form: three (two in 2-D) banded 1-D sweeps, one per direction
  // (R = P^T = R1 (x) R1 (x) R1, P = P1 (x) P1 (x) P1 with the GLOBAL 1-D embedding P1: fine node j in coarse cell e
  // takes the (2k+1) x (k+1) cell embedding, row j - 2k e).  One thread per OUTPUT entry: no zero-initialisation, no
  // atomics, bitwise reproducible; every sweep is coalesced along x.  Restriction contracts x first (N_f -> N_f / 2 ->
  // N_f / 4 -> N_f / 8), prolongation expands z first and adds into the fine vector in the last sweep.
  // `d` = direction of the sweep; ex / ey / ez = extents of the OUTPUT array; the input differs in extent `d` only.
  // =========================================================================================
  struct Sweep1D
  {
    int       d, ex, ey, ez; // direction, output extents
    int       n_in;          // input extent along d
    int       ncc;           // coarse cells along d (of the whole mesh)
    long long N_out, N_in;   // entries per block
    // z-slabs (d = 2 only): first coarse cell handled (grid.y counts the local ones) and the global index of local plane
    // 0 of the input / output array
    int ec0, in_z0, out_z0;
  };
  // Sweeps along y or z (d = 1, 2): one thread per (entry of the other two directions, coarse cell along d), lanes along
  // x: a thread reads the 2k+1 (+2k for the coarse vertex row) fine / k+1 coarse values of its cell once and produces the
  // cell's k coarse / 2k fine entries (index arithmetic once per cell, embedding entries as compile-time constants).
  // grid.x covers the other two directions (flattened), grid.y = coarse cells along d, grid.z = vector blocks.
  template <int K>
  __global__ void __launch_bounds__(256) k_restrict_1d(const Sweep1D w, double *__restrict__ out, const long long os,
                                                       const double *__restrict__ in, const long long is)
  {
    constexpr int  n = K + 1;
    const double  *P = c_fe[K].P;
    const unsigned n_other = (w.d == 1) ? w.ex * w.ez : w.ex * w.ey;
    const unsigned f       = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_other)
      return;
    const int ec = blockIdx.y + w.ec0, b = blockIdx.z;
    // input: extent n_in along d, the output's extents elsewhere
    long long off_in, off_out, sd_in, sd_out;
    if (w.d == 2)
      off_in = f, off_out = f, sd_in = sd_out = (long long)w.ex * w.ey;
    else
      {
        const unsigned ox = f % (unsigned)w.ex, oz = f / (unsigned)w.ex;
        off_in = ox + (long long)w.ex * w.n_in * oz, off_out = ox + (long long)w.ex * w.ey * oz, sd_in = sd_out = w.ex;
      }
    const double *p = in + b * is + off_in + (long long)(2 * K * ec - w.in_z0) * sd_in;
    double       *q = out + b * os + off_out + (long long)(K * ec - w.out_z0) * sd_out;
    double        v[2 * K + 1];
#pragma unroll
    for (int jl = 0; jl <= 2 * K; ++jl)
      v[jl] = p[jl * sd_in];
#pragma unroll
    for (int il = 1; il < K; ++il)
      {
        double s = 0.0;
#pragma unroll
        for (int jl = 0; jl <= 2 * K; ++jl)
          s = fma(P[jl * n + il], v[jl], s);
        q[il * sd_out] = s;
      }
    // the coarse vertex row at the bottom of the cell: the fine nodes of both adjacent coarse cells (the shared one once);
    // coarse Dirichlet entries are 0
    double s = 0.0;
    if (ec > 0)
      {
        s = v[0];
#pragma unroll
        for (int jl = 1; jl <= 2 * K; ++jl)
          s = fma(P[jl * n + 0], v[jl], s);
#pragma unroll
        for (int jl = 0; jl < 2 * K; ++jl)
          s = fma(P[jl * n + K], p[(jl - 2 * K) * sd_in], s);
      }
    q[0] = s;
    if (ec == w.ncc - 1)
      q[K * sd_out] = 0.0;
  }
  template <int K>
  __global__ void __launch_bounds__(256) k_prolongate_1d(const Sweep1D w, double *__restrict__ out, const long long os,
                                                         const double *__restrict__ in, const long long is)
  {
    constexpr int  n = K + 1;
    const double  *P = c_fe[K].P;
    const unsigned n_other = (w.d == 1) ? w.ex * w.ez : w.ex * w.ey;
    const unsigned f       = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_other)
      return;
    const int ec = blockIdx.y + w.ec0, b = blockIdx.z;
    long long off_in, off_out, sd_in, sd_out;
    if (w.d == 2)
      off_in = f, off_out = f, sd_in = sd_out = (long long)w.ex * w.ey;
    else
      {
        const unsigned ox = f % (unsigned)w.ex, oz = f / (unsigned)w.ex;
        off_in = ox + (long long)w.ex * w.n_in * oz, off_out = ox + (long long)w.ex * w.ey * oz, sd_in = sd_out = w.ex;
      }
    const double *p = in + b * is + off_in + (long long)(K * ec - w.in_z0) * sd_in;
    double       *q = out + b * os + off_out + (long long)(2 * K * ec - w.out_z0) * sd_out;
    double        c[n];
#pragma unroll
    for (int il = 0; il <= K; ++il)
      c[il] = p[il * sd_in];
    if (ec == 0) // coarse Dirichlet entries are read as 0
      c[0] = 0.0;
    if (ec == w.ncc - 1)
      c[K] = 0.0;
#pragma unroll
    for (int jl = 0; jl < 2 * K; ++jl)
      {
        double s = 0.0;
#pragma unroll
        for (int il = 0; il <= K; ++il)
          s = fma(P[jl * n + il], c[il], s);
        q[jl * sd_out] = s;
      }
    if (ec == w.ncc - 1)
      q[2 * K * sd_out] = c[K]; // the top fine plane coincides with the (Dirichlet) coarse one
  }
  // restriction along x: a block stages RPB fine rows in shared memory (coalesced), thread (lane, row) contracts the
  // cells lane, lane + 32, ... of its row into a staged coarse row, which is stored coalesced
  template <int K, int RPB>
  __global__ void __launch_bounds__(32 * RPB) k_restrict_x(const int n1f, const int ncc, const long long rows, const long long rows_per_block,
                                                           double *__restrict__ out, const long long os, const double *__restrict__ in,
                                                           const long long is)
  {
    constexpr int n = K + 1;
    extern __shared__ double srow[]; // RPB fine rows of n1f entries, RPB coarse rows of n1c entries
    const double *P   = c_fe[K].P;
    const int     n1c = K * ncc + 1;
    double       *fr  = srow + threadIdx.y * n1f, *cr = srow + RPB * n1f + threadIdx.y * n1c;
    for (long long r0 = (long long)blockIdx.x * RPB; r0 < rows; r0 += (long long)gridDim.x * RPB)
      {
        const long long R = r0 + threadIdx.y;
        const bool      live = R < rows;
        const long long b = live ? R / rows_per_block : 0, rl = R - b * rows_per_block;
        __syncwarp(); // (every warp owns one fine and one coarse row of the staging area: no block-wide barrier)
        if (live)
          for (int x = threadIdx.x; x < n1f; x += 32)
            fr[x] = in[b * is + rl * n1f + x];
        __syncwarp();
        if (live)
          {
            for (int ec = threadIdx.x; ec < ncc; ec += 32)
              {
                const double *p = fr + 2 * K * ec;
                double        v[2 * K + 1];
#pragma unroll
                for (int jl = 0; jl <= 2 * K; ++jl)
                  v[jl] = p[jl];
#pragma unroll
                for (int il = 1; il < K; ++il)
                  {
                    double s = 0.0;
#pragma unroll
                    for (int jl = 0; jl <= 2 * K; ++jl)
                      s = fma(P[jl * n + il], v[jl], s);
                    cr[K * ec + il] = s;
                  }
                double s = 0.0;
                if (ec > 0)
                  {
                    s = v[0];
#pragma unroll
                    for (int jl = 1; jl <= 2 * K; ++jl)
                      s = fma(P[jl * n + 0], v[jl], s);
#pragma unroll
                    for (int jl = 0; jl < 2 * K; ++jl)
                      s = fma(P[jl * n + K], p[jl - 2 * K], s);
                  }
                cr[K * ec] = s;
              }
            if (threadIdx.x == 0)
              cr[n1c - 1] = 0.0;
          }
        __syncwarp();
        if (live)
          for (int i = threadIdx.x; i < n1c; i += 32)
            out[b * os + rl * n1c + i] = cr[i];
      }
  }
  // prolongation along x, added into the fine vector: warp `row` stages its coarse row in shared memory
  template <int K, int RPB>
  __global__ void __launch_bounds__(32 * RPB) k_prolongate_x_add(const int n1f, const int ncc, const long long rows,
                                                                 const long long rows_per_block, double *__restrict__ out, const long long os,
                                                                 const double *__restrict__ in, const long long is)
  {
    constexpr int n = K + 1;
    extern __shared__ double srow[]; // RPB coarse rows of n1c entries, then P
    const int     n1c = K * ncc + 1;
    double       *P   = srow + (size_t)RPB * n1c;
    double       *cr  = srow + threadIdx.y * n1c;
    for (int t = threadIdx.y * 32 + threadIdx.x; t < (2 * K + 1) * n; t += 32 * RPB)
      P[t] = c_fe[K].P[t];
    __syncthreads();
    for (long long r0 = (long long)blockIdx.x * RPB; r0 < rows; r0 += (long long)gridDim.x * RPB)
      {
        const long long R = r0 + threadIdx.y;
        if (R >= rows)
          continue; // (no block-wide barrier below: the rows of a block are independent warps)
        const long long b = R / rows_per_block, rl = R - b * rows_per_block;
        __syncwarp();
        for (int i = threadIdx.x; i < n1c; i += 32)
          cr[i] = (i > 0 && i < n1c - 1) ? in[b * is + rl * n1c + i] : 0.0; // Dirichlet: 0
        __syncwarp();
        double *o = out + b * os + rl * n1f;
        for (int j = threadIdx.x; j < n1f; j += 32)
          {
            const int     ec = min(j / (2 * K), ncc - 1), jl = j - 2 * K * ec;
            const double *p  = cr + K * ec;
            double        s  = 0.0;
#pragma unroll
            for (int il = 0; il <= K; ++il)
              s = fma(P[jl * n + il], p[il], s);
            o[j] += s;
          }
      }
  }

  // y_b = A x_b, tiny dense (coarse-grid solve): one block per vector block
  __global__ void k_dense_matvec(const int n, double *__restrict__ y, const double *__restrict__ x, const long long stride,
                                 const double *__restrict__ A0, const long long matrix_stride)
  {
    extern __shared__ double sx[];
    const int                b = blockIdx.x;
    const double            *A = A0 + b * matrix_stride;
    for (int j = threadIdx.x; j < n; j += blockDim.x)
      sx[j] = x[b * stride + j];
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x)
      {
        double s = 0.0;
        for (int j = 0; j < n; ++j)
          s = fma(A[(size_t)i * n + j], sx[j], s);
        y[b * stride + i] = s;
      }
  }

  // =========================================================================================
  // vector kernels
  // =========================================================================================

  __global__ void k_set(double *x, const long long n, const double v) { SPIRK_GRID_STRIDE(i, n) x[i] = v; }
  __global__ void k_scale(double *x, const long long n, const double a) { SPIRK_GRID_STRIDE(i, n) x[i] *= a; }
  __global__ void k_axpy(double *y, const double a, const double *__restrict__ x, const long long n)
  {
    SPIRK_GRID_STRIDE(i, n) y[i] = fma(a, x[i], y[i]);
  }
  __global__ void k_sadd(double *y, const double s, const double a, const double *__restrict__ x, const long long n)
  {
    SPIRK_GRID_STRIDE(i, n) y[i] = s * y[i] + a * x[i];
  }
  __global__ void k_add2(double *y, const double a, const double *__restrict__ x, const double b,
                         const double *__restrict__ z, const long long n)
  {
    SPIRK_GRID_STRIDE(i, n) y[i] += a * x[i] + b * z[i];
  }
  __global__ void k_equ(double *y, const double a, const double *__restrict__ x, const long long n)
  {
    SPIRK_GRID_STRIDE(i, n) y[i] = a * x[i];
  }
  struct BlockFactors
  {
    double f[SPIRK_MAX_BLOCKS];
  };
  __global__ void k_scale_pointwise(const int nb, const long long n, double *y, const double *__restrict__ d,
                                    const double *__restrict__ x, const long long stride, const BlockFactors f)
  {
    SPIRK_GRID_STRIDE(e, n * nb)
    {
      const int       b = e / n;
      const long long j = b * stride + (e - b * n);
      y[j]              = f.f[b] * (d[j] * x[j]);
    }
  }

  // block-level sum of `v` over the block into partials[blockIdx.x + slot*gridDim.x]
  template <int T>
  __device__ __forceinline__ void block_reduce_store(double v, double *partials)
  {
    __shared__ double sw[T / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
      v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0)
      sw[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32)
      {
        v = (threadIdx.x < T / 32) ? sw[threadIdx.x] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
          v += __shfl_down_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0)
          *partials = v;
      }
    __syncthreads();
  }

  constexpr int RT = 256; // reduction block size
  __global__ void __launch_bounds__(RT) k_dot(const double *__restrict__ x, const double *__restrict__ y, const long long n,
                                              double *partials)
  {
    double s = 0.0;
    SPIRK_GRID_STRIDE(i, n) s = fma(x[i], y[i], s);
    block_reduce_store<RT>(s, partials + blockIdx.x);
  }
  __global__ void __launch_bounds__(RT) k_sum(const double *__restrict__ x, const long long n, double *partials)
  {
    double s = 0.0;
    SPIRK_GRID_STRIDE(i, n) s += x[i];
    block_reduce_store<RT>(s, partials + blockIdx.x);
  }
  // v += a V; partial of v . W     (W may alias v)
  __global__ void __launch_bounds__(RT) k_add_and_dot(double *v, const double a, const double *__restrict__ V, const double *W,
                                                      const long long n, double *partials)
  {
    double s = 0.0;
    SPIRK_GRID_STRIDE(i, n)
    {
      const double t = fma(a, V[i], v[i]);
      const double w = (W == v) ? t : W[i];
      v[i]           = t;
      s              = fma(t, w, s);
    }
    block_reduce_store<RT>(s, partials + blockIdx.x);
  }
  // the same with the coefficient read from device memory, negated: v -= (*coef) V (the Gram-Schmidt sweep keeps its
  // scalars on the device, so a sweep needs one host synchronisation instead of one per basis vector)
  __global__ void __launch_bounds__(RT) k_sub_and_dot_dev(double *v, const double *__restrict__ coef, const double *__restrict__ V,
                                                          const double *W, const long long n, double *partials)
  {
    const double a = -(*coef);
    double       s = 0.0;
    SPIRK_GRID_STRIDE(i, n)
    {
      const double t = fma(a, V[i], v[i]);
      const double w = (W == v) ? t : W[i];
      v[i]           = t;
      s              = fma(t, w, s);
    }
    block_reduce_store<RT>(s, partials + blockIdx.x);
  }
  // the reductions over a block vector whose blocks sit at a stride larger than their length (z-slab vectors carry ghost
  // planes between the blocks): blockIdx.y = block, partial sums in partials[blockIdx.y * gridDim.x + blockIdx.x]
  __global__ void __launch_bounds__(RT) k_dot_strided(const double *__restrict__ x, const double *__restrict__ y, const long long n,
                                                      const long long stride, double *partials)
  {
    const double *xb = x + blockIdx.y * stride, *yb = y + blockIdx.y * stride;
    double        s  = 0.0;
    SPIRK_GRID_STRIDE(i, n) s = fma(xb[i], yb[i], s);
    block_reduce_store<RT>(s, partials + blockIdx.y * gridDim.x + blockIdx.x);
  }
  __global__ void __launch_bounds__(RT) k_sum_strided(const double *__restrict__ x, const long long n, const long long stride, double *partials)
  {
    const double *xb = x + blockIdx.y * stride;
    double        s  = 0.0;
    SPIRK_GRID_STRIDE(i, n) s += xb[i];
    block_reduce_store<RT>(s, partials + blockIdx.y * gridDim.x + blockIdx.x);
  }
  // v += a V (a from the host, or -(*coef) from device memory); partial of v . W (W may alias v)
  __global__ void __launch_bounds__(RT) k_add_and_dot_strided(double *v, const double a_host, const double *__restrict__ coef,
                                                              const double *__restrict__ V, const double *W, const long long n,
                                                              const long long stride, double *partials)
  {
    const double    a = coef ? -(*coef) : a_host;
    const long long o = blockIdx.y * stride;
    double          s = 0.0;
    SPIRK_GRID_STRIDE(i, n)
    {
      const double t = fma(a, V[o + i], v[o + i]);
      const double w = (W == v) ? t : W[o + i];
      v[o + i]       = t;
      s              = fma(t, w, s);
    }
    block_reduce_store<RT>(s, partials + blockIdx.y * gridDim.x + blockIdx.x);
  }
  // result[slot] = sum of partials[0..nblocks)   (fixed order: deterministic)
  __global__ void __launch_bounds__(RT) k_finish(const double *__restrict__ partials, const int nblocks, double *result)
  {
    double s = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += RT)
      s += partials[i];
    block_reduce_store<RT>(s, result);
  }

  // stage mixing (main.cc:1100-1104, 1164-1168, 877-891, 1511-1529)
  struct MixMatrix
  {
    double T[SPIRK_MAX_BLOCKS * SPIRK_MAX_BLOCKS];
  };
  template <int QI>
  __global__ void k_mix(const int qo, double *dst, const long long ds, const double *__restrict__ src, const long long ss,
                        const long long n, const MixMatrix T, const int add)
  {
    SPIRK_GRID_STRIDE(e, n)
    {
      double in[QI];
#pragma unroll
      for (int j = 0; j < QI; ++j)
        in[j] = src[j * ss + e];
      for (int i = 0; i < qo; ++i)
        {
          double t = 0.0;
#pragma unroll
          for (int j = 0; j < QI; ++j)
            t = fma(T.T[i * QI + j], in[j], t); // entries below the cut-off were zeroed on the host
          if (add)
            dst[i * ds + e] += t;
          else
            dst[i * ds + e] = t;
        }
    }
  }

  // fused all-gather + stage mixing over NVLink peer memory: input block j lives in the exchange
  // buffer of rank j / m (peer-mapped pointer), block j % m; the loads of the remote blocks travel
  // over NVLink while the q x q contraction runs (no gathered copy is ever written)
  struct PeerPtrs
  {
    const double *p[SPIRK_MAX_BLOCKS];
  };
  template <int QI>
  __global__ void k_mix_peer(const int qo, const int m, double *dst, const long long ds, const PeerPtrs peers,
                             const long long n, const MixMatrix T, const int add)
  {
    SPIRK_GRID_STRIDE(e, n)
    {
      double in[QI];
#pragma unroll
      for (int j = 0; j < QI; ++j)
        in[j] = peers.p[j / m][(long long)(j % m) * n + e];
      for (int i = 0; i < qo; ++i)
        {
          double t = 0.0;
#pragma unroll
          for (int j = 0; j < QI; ++j)
            t = fma(T.T[i * QI + j], in[j], t);
          if (add)
            dst[i * ds + e] += t;
          else
            dst[i * ds + e] = t;
        }
    }
  }

  // fused all-to-all + stage mixing over NVLink peer memory: this rank contracts ITS chunk [e0, e1) of every stage
  // block (inputs read from the owners' exchange buffers) for ALL q outputs and writes each output chunk straight
  // into the owner's result region: 2 (R-1)/R n doubles cross NVLink per rank and mixing instead of (R-1) m n
  // for the gather formulation (k_mix_peer) - the reduce-scatter / all-gather split of the contraction
  struct PeerPtrsRW
  {
    double *p[SPIRK_MAX_BLOCKS];
  };
  template <int Q>
  __global__ void k_mix_a2a(const int m, const PeerPtrsRW peers, const long long n, const long long out_off, const long long e0,
                            const long long e1, const MixMatrix T)
  {
    for (long long e = e0 + blockIdx.x * (long long)blockDim.x + threadIdx.x; e < e1; e += (long long)gridDim.x * blockDim.x)
      {
        double in[Q];
#pragma unroll
        for (int j = 0; j < Q; ++j)
          in[j] = peers.p[j / m][(long long)(j % m) * n + e];
#pragma unroll
        for (int i = 0; i < Q; ++i)
          {
            double t = 0.0;
#pragma unroll
            for (int j = 0; j < Q; ++j)
              t = fma(T.T[i * Q + j], in[j], t);
            peers.p[i / m][out_off + (long long)(i % m) * n + e] = t;
          }
      }
  }
  // dst_i = [dst_i +] out_i (local result region -> destination blocks)
  __global__ void k_mix_finish(const int m, double *dst, const long long ds, const double *__restrict__ out, const long long n,
                               const int add)
  {
    SPIRK_GRID_STRIDE(e, n * m)
    {
      const int       i = e / n;
      const long long k = e - i * n;
      dst[i * ds + k]   = add ? dst[i * ds + k] + out[e] : out[e];
    }
  }

  // =========================================================================================
  // problem pieces
  // =========================================================================================
  // out[i] = boundary ? 0 : t1[ix] t1[iy] t1[iz] * scale   (separable load vector / nodal solution)
  __global__ void k_outer_product(const Geo g, const double *__restrict__ t1, const double scale, const int zero_bdry,
                                  double *__restrict__ out)
  {
    SPIRK_GRID_STRIDE(i, g.N)
    {
      const int  ix = i % g.n1, iy = (i / g.n1) % g.n1, iz = (g.dim == 3) ? (int)(i / ((long long)g.n1 * g.n1)) + g.zo0 : 1;
      const bool bd = on_bdry(ix, g.n1) || on_bdry(iy, g.n1) || (g.dim == 3 && on_bdry(iz, g.n1));
      double     v  = t1[ix] * t1[iy] * scale;
      if (g.dim == 3)
        v *= t1[iz];
      out[i] = (zero_bdry && bd) ? 0.0 : v;
    }
  }

  // y_b += x_b on interior DoFs (the Dirichlet rows of y keep their values)
  __global__ void k_add_interior(const Geo g, const int nb, double *y, const long long ys, const double *__restrict__ x, const long long xs)
  {
    SPIRK_GRID_STRIDE(e, g.N * nb)
    {
      const int       b  = e / g.N;
      const long long i  = e - b * g.N;
      const int       ix = i % g.n1, iy = (i / g.n1) % g.n1, iz = (g.dim == 3) ? (int)(i / ((long long)g.n1 * g.n1)) + g.zo0 : 1;
      if (!(on_bdry(ix, g.n1) || on_bdry(iy, g.n1) || (g.dim == 3 && on_bdry(iz, g.n1))))
        y[b * ys + i] += x[b * xs + i];
    }
  }

  __global__ void k_set_zero_bdry(const Geo g, const int nb, double *u, const long long stride)
  {
    SPIRK_GRID_STRIDE(e, g.N * nb)
    {
      const int       b  = e / g.N;
      const long long i  = e - b * g.N;
      const int       ix = i % g.n1, iy = (i / g.n1) % g.n1, iz = (g.dim == 3) ? (int)(i / ((long long)g.n1 * g.n1)) + g.zo0 : 1;
      if (on_bdry(ix, g.n1) || on_bdry(iy, g.n1) || (g.dim == 3 && on_bdry(iz, g.n1)))
        u[b * stride + i] = 0.0;
    }
  }

  // L2 / Linf error against the analytical solution with QGauss(k+2) (main.cc:3436-3469).
  // One block per cell; result[0] += sum w d^2 (atomicAdd), result[1] = max |d| (ordered-bit atomicMax).
  template <int K, int DIM>
  __global__ void __launch_bounds__(128) k_error_norms(const Geo g, const double *__restrict__ u, const double ft, double *result)
  {
    constexpr int n = K + 1, ne = K + 2;
    constexpr int NL = (DIM == 3) ? n * n * n : n * n;
    constexpr int N1 = (DIM == 3) ? n * n * ne : n * ne;
    constexpr int N2 = (DIM == 3) ? n * ne * ne : ne * ne;
    constexpr int NQ = (DIM == 3) ? ne * ne * ne : ne * ne;
    __shared__ double su[NL], s1[N1], s2[N2], red[128];
    const double *Be = c_fe[K].Be, *xe = c_fe[K].xe, *we = c_fe[K].we;
    const int     nc = g.nc, n1 = g.n1;
    // the cells of this z-slab (all cells of an unpartitioned level); u is indexed by local plane (one ghost plane above)
    const long long ncells = (DIM == 3) ? (long long)nc * nc * (g.L_hi - g.L_lo) : (long long)nc * nc;
    double          lsum = 0.0, lmax = 0.0;
    const double    twopi = 6.283185307179586476925286766559;
    for (long long c = blockIdx.x; c < ncells; c += gridDim.x)
      {
        const int cx = c % nc, cy = (c / nc) % nc, cz = (DIM == 3) ? (int)(c / ((long long)nc * nc)) + g.L_lo : 0;
        for (int l = threadIdx.x; l < NL; l += blockDim.x)
          {
            const int ix = cx * K + l % n, iy = cy * K + (l / n) % n, iz = (DIM == 3) ? cz * K + l / (n * n) - g.zo0 : 0;
            su[l]        = u[ix + (long long)n1 * (iy + (long long)n1 * iz)];
          }
        __syncthreads();
        for (int e = threadIdx.x; e < N1; e += blockDim.x)
          {
            const int qx = e % ne, zy = e / ne;
            double    s  = 0.0;
            for (int i = 0; i < n; ++i)
              s = fma(Be[qx * n + i], su[zy * n + i], s);
            s1[e] = s;
          }
        __syncthreads();
        for (int e = threadIdx.x; e < N2; e += blockDim.x)
          {
            const int qx = e % ne, qy = (e / ne) % ne, z = e / (ne * ne);
            double    s  = 0.0;
            for (int j = 0; j < n; ++j)
              s = fma(Be[qy * n + j], s1[(z * n + j) * ne + qx], s);
            s2[e] = s;
          }
        __syncthreads();
        for (int e = threadIdx.x; e < NQ; e += blockDim.x)
          {
            const int qx = e % ne, qy = (e / ne) % ne, qz = (DIM == 3) ? e / (ne * ne) : 0;
            double    v;
            if (DIM == 3)
              {
                v = 0.0;
                for (int l = 0; l < n; ++l)
                  v = fma(Be[qz * n + l], s2[(l * ne + qy) * ne + qx], v);
              }
            else
              v = s2[e];
            double ex = sin(twopi * (cx + xe[qx]) * g.h) * sin(twopi * (cy + xe[qy]) * g.h) * ft;
            double w  = we[qx] * we[qy] * g.h * g.h;
            if (DIM == 3)
              {
                ex *= sin(twopi * (cz + xe[qz]) * g.h);
                w *= we[qz] * g.h;
              }
            const double d = v - ex;
            lsum           = fma(w * d, d, lsum);
            lmax           = fmax(lmax, fabs(d));
          }
        __syncthreads();
      }
    red[threadIdx.x] = lsum;
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1)
      {
        if (threadIdx.x < o)
          red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
      }
    if (threadIdx.x == 0)
      atomicAdd(result, red[0]);
    __syncthreads();
    red[threadIdx.x] = lmax;
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1)
      {
        if (threadIdx.x < o)
          red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + o]);
        __syncthreads();
      }
    if (threadIdx.x == 0)
      atomicMax((unsigned long long *)(result + 1), (unsigned long long)__double_as_longlong(red[0]));
  }
} // namespace spirk
