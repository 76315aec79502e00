// Matrix-free cell operator, variant 2: fused-epilogue fast path for 3-D levels (no atomics, no
// zeroing pass, deterministic), written as a warp-specialised software pipeline.
//
// Work decomposition.  The cells are grouped into tile columns of TX x TY cells; one CTA (1024
// threads, one per SM) sweeps a column upwards in z, one cell layer (TX x TY cells,
// (k TX+1) x (k TY+1) x (k+1) nodes) at a time.  Three warp groups work on three consecutive layers
// at the same time, handing tiles to each other through double-buffered shared memory and named
// barriers (bar.arrive / bar.sync), so the FP64 pipe always has a phase to run while another group
// waits on shared or global memory:
//   group A : stages the layer's node planes of `src` with cp.async into a ring of 2k+1 planes
//             (Dirichlet nodes zero-filled) one layer ahead, then contracts along x, register-tiled:
//             one thread owns a cell segment of an x-line, a = Mx u, b = Kx u (the node shared by two
//             cells of the tile is finished by the right cell)
//   group B : the same along y:  p = My a, q = Ky a, r = My b;  s = cm p + cl (q + r)
//   group C : along z one thread owns the node column (X, Y):  o = Mz s + cl Kz p.  The top plane of
//             the layer stays in a REGISTER and is added to the bottom plane of the next layer, so
//             sums across z never touch memory; planes 0..k-1 are then complete inside the column
//             and the epilogue (plain / residual / Chebyshev update) is applied in place.
// Because the cells of a tile are congruent, summing the cell matrices over the tile before applying
// them is exact (sum of Kronecker products over a Cartesian set of cells), so groups A/B finish every
// line once instead of once per adjacent cell (about 56 instead of 68 FMA per unique DoF at k = 4).
// Node columns on the four side walls of a tile column also receive contributions from the
// neighbouring columns.  Their partial sums go to compact wall arrays (plain stores, one slot per
// contributing column) with the linear part of the epilogue already folded in by one designated
// contributor; a second small kernel adds the slots in a fixed order.  Long columns are cut into
// z-chunks for parallelism; a chunk recomputes the cell layer below it to obtain its incoming carry
// plane, so no z-walls exist.
//
// Replaces the deal.II cell loop + vector updates of the reference:
//   operator.h:298-310, 379-421 (vmult), 841-880 (batched); deal.II PreconditionChebyshev
//   vector_updates and the residual of Multigrid::level_v_step (A8).
#pragma once
#include "op_v1.cuh"

namespace spirk
{
  enum V2Mode
  {
    V2_APPLY    = 0, // dst = A src
    V2_RESIDUAL = 1, // dst = rhs - A src
    V2_CHEB     = 2  // dst = src + f1 (src - x_old) + f2 dinv (rhs - A src)
  };

  template <int K, int TX, int TY>
  struct CfgV2
  {
    static constexpr int n = K + 1, LX = K * TX + 1, LY = K * TY + 1, PL = LX * LY;
    static constexpr int itemsA = TX * LY * n, itemsB = LX * TY * n, itemsC = PL;
    static constexpr int IPT = 2; // work items per thread and layer (keeps the register tiles spill-free)
    static constexpr int TA = (((itemsA + IPT - 1) / IPT + 31) / 32) * 32, TB = (((itemsB + IPT - 1) / IPT + 31) / 32) * 32,
                         TC = (((itemsC + IPT - 1) / IPT + 31) / 32) * 32;
    static constexpr int TD      = 128; // epilogue / store warps
    static constexpr int threads = TA + TB + TC + TD;
    static constexpr int ring    = 2 * K + 1; // node planes of the current layer + the K new planes of the next one
    static constexpr size_t smem = sizeof(double) * (ring + 8 * n + 2 * K) * PL; // ring + 2 x (A,B) + 2 x (S,P) + 2 x OUT
    static_assert(threads <= 1024, "tile too large for one CTA");
  };

  struct V2Args
  {
    Geo          g;
    int          mode, nb;
    long long    stride;
    double      *dst;
    const double *src, *x_old, *rhs, *dinv;
    double       cm[SPIRK_MAX_BLOCKS], cl[SPIRK_MAX_BLOCKS], f1[SPIRK_MAX_BLOCKS], f2[SPIRK_MAX_BLOCKS];
    int          ntx, nty, nchunks, chunk_len;
    double      *WX, *WY, *WC; // wall slots
    long long    wx_block, wy_block, wc_block; // per-vector-block sizes of the wall arrays
  };

  __device__ __forceinline__ void bar_sync(const int id, const int count)
  {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
  }
  __device__ __forceinline__ void bar_arrive(const int id, const int count)
  {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
  }

  // epilogue for one DoF (index j into the vector incl. block offset); Ax = (A src)_j
  __device__ __forceinline__ void v2_epilogue(const V2Args &a, const int b, const long long j, const double Ax, const double x)
  {
    if (a.mode == V2_APPLY)
      a.dst[j] = Ax;
    else if (a.mode == V2_RESIDUAL)
      a.dst[j] = a.rhs[j] - Ax;
    else
      {
        const double xo = a.x_old ? a.x_old[j] : 0.0;
        a.dst[j]        = (1.0 + a.f1[b]) * x - a.f1[b] * xo + a.f2[b] * a.dinv[j] * (a.rhs[j] - Ax);
      }
  }

  template <int K, int TX, int TY, int MODE>
  __global__ void __launch_bounds__(CfgV2<K, TX, TY>::threads, 1) k_v2_main(const V2Args a)
  {
    using C             = CfgV2<K, TX, TY>;
    constexpr int n = C::n, LX = C::LX, LY = C::LY, PL = C::PL, RING = C::ring, IPT = C::IPT;
    constexpr int TA = C::TA, TB = C::TB, TC = C::TC, TD = C::TD;
    constexpr int NPF = (K * PL + TA - 1) / TA; // prefetch elements per group-A thread and layer
    // named barriers (0 is __syncthreads)
    constexpr int BAR_AB_FULL = 1, BAR_AB_EMPTY = 3, BAR_SP_FULL = 5, BAR_SP_EMPTY = 7, BAR_A = 9, BAR_OUT_FULL = 10,
                  BAR_OUT_EMPTY = 12;
    extern __shared__ double sm[];
    double *U = sm, *AB = sm + RING * PL, *SP = AB + 4 * n * PL, *OUT = SP + 4 * n * PL; // AB[buf][A|B][plane], SP[buf][S|P][plane], OUT[buf][plane]
    const double *Mh = c_fe[K].Mh, *Kh = c_fe[K].Kh;

    const int n1 = a.g.n1, nc = a.g.nc;
    int       bid = blockIdx.x;
    const int tx  = bid % a.ntx;
    bid /= a.ntx;
    const int ty = bid % a.nty;
    bid /= a.nty;
    const int chunk = bid % a.nchunks;
    const int b     = bid / a.nchunks;

    const int       gx0 = tx * K * TX, gy0 = ty * K * TY;
    const int       zb = chunk * a.chunk_len, ze = min(nc, zb + a.chunk_len);
    const int       z_first = (zb > 0 ? zb - 1 : zb), n_layers = ze - z_first;
    const long long boff = (long long)b * a.stride, plane = (long long)n1 * n1;
    const double   *src  = a.src + boff;

    if (threadIdx.x < TA)
      {
        // =========================== group A: staging + x-lines ===========================
        const int t = threadIdx.x;
        // prefetch descriptors of the K new planes of a layer: element e = t + i*TA
        const double *pf_base = src + plane * ((long long)K * (z_first + 1));
        int           pf_off[NPF], pf_goff[NPF];
#pragma unroll
        for (int i = 0; i < NPF; ++i)
          {
            const int e = t + i * TA;
            const int X = e % LX, Y = (e / LX) % LY, pl = 1 + e / PL;
            const int gx = gx0 + X, gy = gy0 + Y;
            pf_off[i]  = (e < K * PL) ? (Y * LX + X) | (pl << 16) | ((on_bdry(gx, n1) || on_bdry(gy, n1)) ? (1 << 30) : 0) : -1;
            pf_goff[i] = gx + n1 * gy + (int)plane * pl;
          }
        // prologue: all k+1 planes of the first layer
        for (int e = t; e < n * PL; e += TA)
          {
            const int  X = e % LX, Y = (e / LX) % LY, gz = K * z_first + e / PL;
            const int  gx = gx0 + X, gy = gy0 + Y;
            const bool bd = on_bdry(gx, n1) || on_bdry(gy, n1) || on_bdry(gz, n1);
            const unsigned int sp = (unsigned int)__cvta_generic_to_shared(U + (gz % RING) * PL + Y * LX + X);
            const int          sz = bd ? 0 : 8;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(sp), "l"(src + gx + (long long)n1 * gy + plane * gz), "r"(sz));
          }
        asm volatile("cp.async.commit_group;\n" ::);
        int slot0 = (K * z_first) % RING; // ring slot of plane 0 of the current layer
        for (int it = 0; it < n_layers; ++it)
          {
            asm volatile("cp.async.wait_all;\n" ::);
            bar_sync(BAR_A, TA); // this layer's planes are visible to group A; layer it-1 is fully consumed
            if (it + 1 < n_layers)
              {
#pragma unroll
                for (int i = 0; i < NPF; ++i)
                  if (pf_off[i] >= 0)
                    {
                      const int pl = (pf_off[i] >> 16) & 15;
                      const int gz = K * (z_first + it + 1) + pl;
                      int       sl = slot0 + K + pl;
                      sl -= (sl >= RING) ? RING : 0;
                      sl -= (sl >= RING) ? RING : 0;
                      const unsigned int sp = (unsigned int)__cvta_generic_to_shared(U + sl * PL + (pf_off[i] & 0xffff));
                      const int          sz = ((pf_off[i] >> 30) || gz == n1 - 1) ? 0 : 8;
                      asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(sp), "l"(pf_base + pf_goff[i]), "r"(sz));
                    }
                asm volatile("cp.async.commit_group;\n" ::);
                pf_base += K * plane;
              }
            const int buf = it & 1;
            if (it >= 2)
              bar_sync(BAR_AB_EMPTY + buf, TA + TB); // group B has released this buffer
#pragma unroll 1
            for (int w = 0; w < IPT; ++w)
              {
                const int e = t + w * TA;
                if (e < C::itemsA)
                  {
                    const int seg = e % TX, off = ((e / TX) % LY) * LX + K * seg, zl = e / (TX * LY);
                    int       sl  = slot0 + zl;
                    sl -= (sl >= RING) ? RING : 0;
                    const double *row = U + sl * PL + off;
                    double        u[n], av[n], bv[n];
#pragma unroll
                    for (int j = 0; j < n; ++j)
                      u[j] = row[j];
                    matvec<n>(Mh, u, av);
                    matvec<n>(Kh, u, bv);
                    if (seg > 0)
                      { // finish the node shared with the previous cell of the tile: its row K
                        double am = Mh[K * n + K] * u[0], ak = Kh[K * n + K] * u[0];
#pragma unroll
                        for (int j = 0; j < K; ++j)
                          {
                            const double up = row[j - K];
                            am              = fma(Mh[K * n + j], up, am);
                            ak              = fma(Kh[K * n + j], up, ak);
                          }
                        av[0] += am, bv[0] += ak;
                      }
                    double *oa = AB + (buf * 2 * n + zl) * PL + off, *ob = oa + n * PL;
#pragma unroll
                    for (int i = 0; i < K; ++i)
                      oa[i] = av[i], ob[i] = bv[i];
                    if (seg == TX - 1)
                      oa[K] = av[K], ob[K] = bv[K];
                  }
              }
            bar_arrive(BAR_AB_FULL + buf, TA + TB);
            slot0 += K;
            slot0 -= (slot0 >= RING) ? RING : 0;
          }
      }
    else if (threadIdx.x < TA + TB)
      {
        // =========================== group B: y-lines ===========================
        const int    t  = threadIdx.x - TA;
        const double cm = a.cm[b], cl = a.cl[b];
        for (int it = 0; it < n_layers; ++it)
          {
            const int buf = it & 1;
            bar_sync(BAR_AB_FULL + buf, TA + TB);
            if (it >= 2)
              bar_sync(BAR_SP_EMPTY + buf, TB + TC);
#pragma unroll 1
            for (int w = 0; w < IPT; ++w)
              {
                const int e = t + w * TB;
                if (e < C::itemsB)
                  {
                    const int     seg = (e / LX) % TY, off = (K * seg) * LX + e % LX, zl = e / (LX * TY);
                    const double *ca = AB + (buf * 2 * n + zl) * PL + off, *cb = ca + n * PL;
                    double        av[n], pv[n], q[n], sv[n];
#pragma unroll
                    for (int j = 0; j < n; ++j)
                      av[j] = ca[j * LX];
                    matvec<n>(Mh, av, pv);
                    matvec<n>(Kh, av, q);
                    if (seg > 0)
                      {
                        double pm = Mh[K * n + K] * av[0], qk = Kh[K * n + K] * av[0];
#pragma unroll
                        for (int j = 0; j < K; ++j)
                          {
                            const double ap = ca[(j - K) * LX];
                            pm              = fma(Mh[K * n + j], ap, pm);
                            qk              = fma(Kh[K * n + j], ap, qk);
                          }
                        pv[0] += pm, q[0] += qk;
                      }
                    // s = cm p + cl (q + My b): fold the b column into q with b pre-scaled by cl
#pragma unroll
                    for (int i = 0; i < n; ++i)
                      q[i] = fma(cm, pv[i], cl * q[i]);
#pragma unroll
                    for (int j = 0; j < n; ++j)
                      av[j] = cl * cb[j * LX];
#pragma unroll
                    for (int i = 0; i < n; ++i)
                      {
                        double acc = q[i];
#pragma unroll
                        for (int j = 0; j < n; ++j)
                          acc = fma(Mh[i * n + j], av[j], acc);
                        sv[i] = acc;
                      }
                    if (seg > 0)
                      {
                        double rm = Mh[K * n + K] * av[0];
#pragma unroll
                        for (int j = 0; j < K; ++j)
                          rm = fma(Mh[K * n + j], cl * cb[(j - K) * LX], rm);
                        sv[0] += rm;
                      }
                    double *os = SP + (buf * 2 * n + zl) * PL + off, *op = os + n * PL;
#pragma unroll
                    for (int i = 0; i < K; ++i)
                      os[i * LX] = sv[i], op[i * LX] = pv[i];
                    if (seg == TY - 1)
                      os[K * LX] = sv[K], op[K * LX] = pv[K];
                  }
              }
            if (it + 2 < n_layers)
              bar_arrive(BAR_AB_EMPTY + buf, TA + TB); // the A/B tile of this layer is consumed
            bar_arrive(BAR_SP_FULL + buf, TB + TC);
          }
      }
    else if (threadIdx.x < TA + TB + TC)
      {
        // =========================== group C: z-lines ===========================
        const int    t  = threadIdx.x - TA - TB;
        const double cl = a.cl[b];
        static_assert(IPT == 2, "group C keeps one carry register per item");
        double carry0 = 0.0, carry1 = 0.0;
        for (int it = 0; it < n_layers; ++it)
          {
            const int buf = it & 1;
            bar_sync(BAR_SP_FULL + buf, TB + TC);
            if (it >= 2)
              bar_sync(BAR_OUT_EMPTY + buf, TC + TD);
#pragma unroll 1
            for (int w = 0; w < IPT; ++w)
              {
                const int e = t + w * TC;
                if (e < C::itemsC)
                  {
                    const double *cs = SP + (buf * 2 * n) * PL + e, *cp = cs + n * PL;
                    double        s[n], p[n], val[n];
#pragma unroll
                    for (int z = 0; z < n; ++z)
                      s[z] = cs[z * PL], p[z] = cl * cp[z * PL];
#pragma unroll
                    for (int z = 0; z < n; ++z)
                      {
                        double acc = 0.0;
#pragma unroll
                        for (int j = 0; j < n; ++j)
                          acc = fma(Mh[z * n + j], s[j], fma(Kh[z * n + j], p[j], acc));
                        val[z] = acc;
                      }
                    val[0] += (w == 0) ? carry0 : carry1;
                    if (w == 0)
                      carry0 = val[K];
                    else
                      carry1 = val[K];
                    double *o = OUT + buf * K * PL + e;
#pragma unroll
                    for (int z = 0; z < K; ++z)
                      o[z * PL] = val[z];
                  }
              }
            if (it + 2 < n_layers)
              bar_arrive(BAR_SP_EMPTY + buf, TB + TC);
            bar_arrive(BAR_OUT_FULL + buf, TC + TD);
          }
      }
    else
      {
        // =========================== group D: epilogue / stores ===========================
        // streams the k finished node planes of a layer: interior nodes get the fused epilogue, wall
        // nodes their slot (E + F * partial, E only from the carrier = low side in x and y)
        const int    t  = threadIdx.x - TA - TB - TC;
        const double f1 = a.f1[b], f2 = a.f2[b];
        for (int it = 0; it < n_layers; ++it)
          {
            const int  buf = it & 1, zc = z_first + it;
            bar_sync(BAR_OUT_FULL + buf, TC + TD);
            if (zc >= zb)
              {
#pragma unroll 2
                for (int e = t; e < K * PL; e += TD)
                  {
                    const int       cX = e % LX, cY = (e / LX) % LY, z = e / PL;
                    const double    v  = OUT[buf * K * PL + e];
                    const int       gx = gx0 + cX, gy = gy0 + cY, gz = K * zc + z;
                    const long long j  = boff + gx + (long long)n1 * gy + plane * gz;
                    const bool      wallx = (cX == 0) || (cX == LX - 1), wally = (cY == 0) || (cY == LY - 1);
                    const bool      is_wall = wallx || wally;
                    double         *out     = a.dst + j;
                    bool            carrier = true;
                    if (is_wall)
                      {
                        const int wx = tx + (cX == 0 ? 0 : 1), dx = (cX == 0) ? 1 : 0;
                        const int wy = ty + (cY == 0 ? 0 : 1), dy = (cY == 0) ? 1 : 0;
                        if (wallx && wally)
                          out = a.WC + b * a.wc_block + (((long long)(dy * 2 + dx) * (a.ntx + 1) + wx) * (a.nty + 1) + wy) * n1 + gz,
                          carrier = (dx == 0 && dy == 0);
                        else if (wallx)
                          out = a.WX + b * a.wx_block + ((long long)dx * (a.ntx + 1) + wx) * plane + (long long)gz * n1 + gy,
                          carrier = (dx == 0);
                        else
                          out = a.WY + b * a.wy_block + ((long long)dy * (a.nty + 1) + wy) * plane + (long long)gz * n1 + gx,
                          carrier = (dy == 0);
                      }
                    double E = 0.0, F = 1.0;
                    if (MODE == V2_RESIDUAL)
                      {
                        F = -1.0;
                        E = carrier ? a.rhs[j] : 0.0;
                      }
                    else if (MODE == V2_CHEB)
                      {
                        F = -f2 * a.dinv[j];
                        if (carrier)
                          {
                            const double xo = a.x_old ? a.x_old[j] : 0.0;
                            E               = (1.0 + f1) * a.src[j] - f1 * xo - F * a.rhs[j];
                          }
                      }
                    double Ax = v;
                    if (!is_wall && gz == 0) // Dirichlet plane gz = 0: A is the identity there
                      Ax = a.src[j];
                    *out = fma(F, Ax, E);
                  }
              }
            if (it + 2 < n_layers)
              bar_arrive(BAR_OUT_EMPTY + buf, TC + TD);
          }
        // top plane of the domain (Dirichlet): interior node columns of the last chunk
        if (ze == nc)
          for (int e = t; e < PL; e += TD)
            {
              const int cX = e % LX, cY = e / LX;
              if (cX > 0 && cX < LX - 1 && cY > 0 && cY < LY - 1)
                {
                  const long long gi = (gx0 + cX) + (long long)n1 * (gy0 + cY) + plane * (n1 - 1);
                  const double    x  = src[gi];
                  v2_epilogue(a, b, boff + gi, x, x);
                }
            }
      }
  }

  // wall kernel: add the slots of every node on a tile wall in a fixed order (the linear part of the
  // epilogue is already folded into the carrier's slot); Dirichlet wall nodes get the identity.
  // grid: x = chunks of the (wall plane, in-plane coordinate) index, y = node plane gz, z = vector block
  template <int K, int TX, int TY>
  __global__ void __launch_bounds__(256) k_v2_walls(const V2Args a)
  {
    const int n1 = a.g.n1, gz = blockIdx.y, b = blockIdx.z;
    const int nx = (a.ntx + 1) * n1, ny = (a.nty + 1) * n1, ncn = (a.ntx + 1) * (a.nty + 1);
    const int i  = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nx + ny + ncn)
      return;
    const long long plane = (long long)n1 * n1;
    int             gx, gy;
    const double   *w0;
    long long       wst;
    int             nslots;
    if (i < nx)
      { // x-wall plane wx, in-plane coordinate gy
        const int wx = i / n1;
        gy           = i - wx * n1;
        if (gy % (K * TY) == 0)
          return; // corner line: handled below
        gx  = wx * K * TX;
        w0  = a.WX + b * a.wx_block + (long long)wx * plane + (long long)gz * n1 + gy;
        wst = (long long)(a.ntx + 1) * plane, nslots = 2;
      }
    else if (i < nx + ny)
      {
        const int r = i - nx, wy = r / n1;
        gx          = r - wy * n1;
        if (gx % (K * TX) == 0)
          return;
        gy  = wy * K * TY;
        w0  = a.WY + b * a.wy_block + (long long)wy * plane + (long long)gz * n1 + gx;
        wst = (long long)(a.nty + 1) * plane, nslots = 2;
      }
    else
      {
        const int r = i - nx - ny, wx = r / (a.nty + 1), wy = r - wx * (a.nty + 1);
        gx = wx * K * TX, gy = wy * K * TY;
        w0  = a.WC + b * a.wc_block + ((long long)wx * (a.nty + 1) + wy) * n1 + gz;
        wst = (long long)ncn * n1, nslots = 4;
      }
    const long long j  = (long long)b * a.stride + gx + (long long)n1 * gy + plane * gz;
    const bool      bd = on_bdry(gx, n1) || on_bdry(gy, n1) || on_bdry(gz, n1);
    if (bd)
      {
        const double x = a.src[j];
        v2_epilogue(a, b, j, x, x);
      }
    else
      a.dst[j] = (nslots == 2) ? (w0[0] + w0[wst]) : ((w0[0] + w0[wst]) + (w0[2 * wst] + w0[3 * wst]));
  }

  struct V2Scratch
  {
    double *buf = nullptr;
    size_t  cap = 0;
  };
  inline V2Scratch &v2_scratch(spirk_ctx *ctx)
  {
    static thread_local std::vector<std::pair<spirk_ctx *, V2Scratch>> tab;
    for (auto &p : tab)
      if (p.first == ctx)
        return p.second;
    tab.push_back({ctx, V2Scratch()});
    return tab.back().second;
  }

  template <int K, int TX, int TY>
  int v2_launch(spirk_ctx *ctx, V2Args &a)
  {
    using C = CfgV2<K, TX, TY>;
    static bool attr_set = false;
    if (!attr_set)
      {
        SPIRK_CUDA(cudaFuncSetAttribute(k_v2_main<K, TX, TY, V2_APPLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::smem));
        SPIRK_CUDA(cudaFuncSetAttribute(k_v2_main<K, TX, TY, V2_RESIDUAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::smem));
        SPIRK_CUDA(cudaFuncSetAttribute(k_v2_main<K, TX, TY, V2_CHEB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::smem));
        attr_set = true;
      }
    const int n1 = a.g.n1;
    a.ntx = a.g.nc / TX, a.nty = a.g.nc / TY;
    const long long columns = (long long)a.ntx * a.nty;
    // z-chunks: minimise the makespan (waves x layers per chunk incl. the recomputed layer below a chunk)
    {
      const long long slots = ctx->n_sms;
      long long       best  = -1;
      for (int nch = 1; nch <= std::max(1, a.g.nc / 4); ++nch)
        {
          const int       len   = (a.g.nc + nch - 1) / nch, real = (a.g.nc + len - 1) / len;
          const long long waves = (columns * real * a.nb + slots - 1) / slots;
          const long long cost  = waves * (len + (real > 1 ? 1 : 0) + 2); // + pipeline fill
          if (best < 0 || cost < best)
            best = cost, a.chunk_len = len, a.nchunks = real;
        }
    }
    a.wx_block = 2LL * (a.ntx + 1) * n1 * n1;
    a.wy_block = 2LL * (a.nty + 1) * n1 * n1;
    a.wc_block = 4LL * (a.ntx + 1) * (a.nty + 1) * n1;
    const size_t need = (size_t)(a.wx_block + a.wy_block + a.wc_block) * a.nb;
    V2Scratch   &sc   = v2_scratch(ctx);
    if (sc.cap < need)
      {
        if (sc.buf)
          {
            SPIRK_CUDA(cudaStreamSynchronize(ctx->stream));
            SPIRK_CUDA(cudaFree(sc.buf));
          }
        sc.buf = nullptr, sc.cap = 0;
        SPIRK_CUDA(cudaMalloc(&sc.buf, need * sizeof(double)));
        sc.cap = need;
      }
    a.WX = sc.buf;
    a.WY = a.WX + a.wx_block * a.nb;
    a.WC = a.WY + a.wy_block * a.nb;
    const long long grid = columns * a.nchunks * a.nb;
    if (a.mode == V2_APPLY)
      k_v2_main<K, TX, TY, V2_APPLY><<<(unsigned int)grid, C::threads, C::smem, ctx->stream>>>(a);
    else if (a.mode == V2_RESIDUAL)
      k_v2_main<K, TX, TY, V2_RESIDUAL><<<(unsigned int)grid, C::threads, C::smem, ctx->stream>>>(a);
    else
      k_v2_main<K, TX, TY, V2_CHEB><<<(unsigned int)grid, C::threads, C::smem, ctx->stream>>>(a);
    SPIRK_LAUNCH_CHECK(ctx);
    const int  per_plane = (a.ntx + 1) * n1 + (a.nty + 1) * n1 + (a.ntx + 1) * (a.nty + 1);
    const dim3 wgrid((per_plane + 255) / 256, n1, a.nb);
    k_v2_walls<K, TX, TY><<<wgrid, 256, 0, ctx->stream>>>(a);
    SPIRK_LAUNCH_CHECK(ctx);
    return SPIRK_OK;
  }

  // returns SPIRK_ERR_UNSUPPORTED when the level / operator shape is not covered; the caller then
  // uses the general variant-1 kernels.
  inline int v2_apply(spirk_ctx *ctx, const Geo &g, const spirk_opdesc *op, V2Mode mode, double *dst, const double *src,
                      const double *x_old, const double *rhs, const double *dinv, long long stride, const double *f1,
                      const double *f2)
  {
    if (g.dim != 3 || g.k != 4 || op->kind != SPIRK_OP_REAL || g.nc % 4 != 0 || g.nc < 8)
      return SPIRK_ERR_UNSUPPORTED;
    V2Args a;
    a.g = g, a.mode = mode, a.nb = op->nb, a.stride = stride;
    a.dst = dst, a.src = src, a.x_old = x_old, a.rhs = rhs, a.dinv = dinv;
    const double hd = g.h * g.h * g.h, hl = g.h;
    for (int b = 0; b < op->nb; ++b)
      {
        a.cm[b] = op->mass[b] * hd, a.cl[b] = op->laplace[b] * hl;
        a.f1[b] = f1 ? f1[b] : 0.0, a.f2[b] = f2 ? f2[b] : 0.0;
      }
    return v2_launch<4, 4, 4>(ctx, a);
  }
} // namespace spirk
