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
#include <cstdio>
#include <cstdlib>

#include "op_v1.cuh"

namespace spirk
{
  enum V2Mode
  {
    V2_APPLY    = 0, // dst = A src
    V2_RESIDUAL = 1, // dst = rhs - A src
    V2_CHEB     = 2, // dst = src + f1 (src - x_old) + f2 dinv (rhs - A src)
    V2_CHEB_OWN = 3, // the same with dinv = the operator's own inverse diagonal, computed on the fly
    // first two Chebyshev iterates in one pass (variant 3 only): x1 = f0 dinv src, dst = x1 + f1 x1 + f2 dinv (src - A x1)
    V2_CHEB_FIRST = 4
  };

  template <int K, int TX, int TY>
  struct CfgV2
  {
    static constexpr int n = K + 1, LX = K * TX + 1, LY = K * TY + 1;
    static constexpr int LXP = LX + 1;     // padded row pitch (even: 16-byte aligned rows for 128-bit LDS/STS)
    static constexpr int PLP = LXP * LY;   // padded plane
    static constexpr int NPG = (n + 1) / 2; // plane groups: a thread of group A/B owns <= 2 planes of its line
    static constexpr int pairsA = TX * LY, pairsB = LX * TY;
    static constexpr int TA = ((pairsA * NPG + 31) / 32) * 32, TB = ((pairsB * NPG + 31) / 32) * 32;
    static constexpr int TC = (((LX * LY + 2) / 3 + 31) / 32) * 32; // three node columns per thread
    static constexpr int TD = 320;                                 // epilogue / store warps
    static constexpr int threads = TA + TB + TC + TD;
    static constexpr int ringp   = 3 * K; // node-plane ring: 3 layers x K planes
    // operands staged with cp.async by the epilogue group: none (apply), rhs (residual), rhs + x_old (Chebyshev)
    __host__ __device__ static constexpr int nops_async(const int mode) { return mode >= 2 ? 2 : (mode == 1 ? 1 : 0); }
    static constexpr int NE_D = (K * (LX - 2) * (LY - 2) + TD - 1) / TD + (K * (2 * LX + 2 * (LY - 2)) + TD - 1) / TD;
    // ring + 2 x (A,B) + 2 x (S,P) + 2 x OUT + 3-deep ring of thread-private operand slots of the epilogue group
    static constexpr size_t smem(const int mode)
    {
      return sizeof(double) * ((size_t)(ringp + 8 * n + 2 * K) * PLP + 3 * (size_t)nops_async(mode) * NE_D * TD);
    }
    static_assert(threads <= 1024, "tile too large for one CTA");
    static_assert(K % 2 == 0, "128-bit shared accesses need an even degree");
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
        // dinv == NULL: only used for Dirichlet nodes when dinv == NULL (the inverse diagonal is 1 there)
        const double di = a.dinv ? a.dinv[j] : 1.0;
        a.dst[j]        = (1.0 + a.f1[b]) * x - a.f1[b] * xo + a.f2[b] * di * (a.rhs[j] - Ax);
      }
  }

  template <int K, int TX, int TY, int MODE>
  __global__ void __launch_bounds__(CfgV2<K, TX, TY>::threads, 1) k_v2_main(const V2Args a)
  {
    using C             = CfgV2<K, TX, TY>;
    constexpr int n = C::n, LX = C::LX, LY = C::LY, LXP = C::LXP, PLP = C::PLP, NPG = C::NPG;
    constexpr int TA = C::TA, TB = C::TB, TC = C::TC, TD = C::TD;
    // named barriers (0 is __syncthreads)
    constexpr int BAR_AB_FULL = 1, BAR_AB_EMPTY = 3, BAR_SP_FULL = 5, BAR_SP_EMPTY = 7, BAR_A = 9, BAR_OUT_FULL = 10,
                  BAR_OUT_EMPTY = 12;
    extern __shared__ __align__(16) double sm[];
    // U: ring of 3 layers x K node planes; AB[buf][A|B][plane]; SP[buf][S|P][plane]; OUT[buf][plane < K]
    double *U = sm, *AB = sm + C::ringp * PLP, *SP = AB + 4 * n * PLP, *OUT = SP + 4 * n * PLP;
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
        const int  t    = threadIdx.x;
        const int  pair = t % C::pairsA, pg = t / C::pairsA;
        const int  seg  = pair % TX, off = (pair / TX) * LXP + K * seg;
        const bool act  = pg < NPG;
        // prefetch descriptors of the K new planes of a layer: element e = t + i*TA of K x (LX x LY)
        constexpr int NPF = (K * LX * LY + TA - 1) / TA;
        int           pf_s[NPF], pf_g[NPF]; // smem offset (bit 30: x/y Dirichlet, bit 29: top plane of the layer); global offset
#pragma unroll
        for (int i = 0; i < NPF; ++i)
          {
            const int e = t + i * TA;
            const int X = e % LX, Y = (e / LX) % LY, pl = 1 + e / (LX * LY);
            const int gx = gx0 + X, gy = gy0 + Y;
            pf_g[i] = gx + n1 * gy + (int)plane * pl;
            pf_s[i] = (e < K * LX * LY) ? ((pl == K ? 0 : pl * PLP) + Y * LXP + X) | (pl == K ? (1 << 29) : 0) |
                                            ((on_bdry(gx, n1) || on_bdry(gy, n1)) ? (1 << 30) : 0)
                                        : -1;
          }
        // prologue: all k+1 planes of the first layer -> ring groups 0 (planes 0..K-1) and 1 (plane K)
        for (int e = t; e < n * LX * LY; e += TA)
          {
            const int  X = e % LX, Y = (e / LX) % LY, pl = e / (LX * LY), gz = K * z_first + pl;
            const int  gx = gx0 + X, gy = gy0 + Y;
            const bool bd = on_bdry(gx, n1) || on_bdry(gy, n1) || on_bdry(gz, n1);
            const unsigned int sp = (unsigned int)__cvta_generic_to_shared(U + pl * PLP + Y * LXP + X);
            const int          sz = bd ? 0 : 8;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(sp), "l"(src + gx + (long long)n1 * gy + plane * gz), "r"(sz));
          }
        asm volatile("cp.async.commit_group;\n" ::);
        const double *pf_base = src + plane * ((long long)K * (z_first + 1)); // plane 0 of the next layer
        int           cur = 0, nxt = K * PLP, nn = 2 * K * PLP;                  // ring group bases (elements)
        for (int it = 0; it < n_layers; ++it)
          {
            asm volatile("cp.async.wait_all;\n" ::);
            bar_sync(BAR_A, TA); // this layer's planes are visible to group A; layer it-1 is fully consumed
            if (it + 1 < n_layers)
              {
                const bool top = (z_first + it + 2 == nc); // plane K of the next layer is the Dirichlet top plane
#pragma unroll
                for (int i = 0; i < NPF; ++i)
                  if (pf_s[i] >= 0)
                    {
                      const bool last = (pf_s[i] >> 29) & 1;
                      const int  so   = (last ? nn : nxt) + (pf_s[i] & 0xffffff);
                      const unsigned int sp = (unsigned int)__cvta_generic_to_shared(U + so);
                      const int          sz = ((pf_s[i] >> 30) || (last && top)) ? 0 : 8;
                      asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(sp), "l"(pf_base + pf_g[i]), "r"(sz));
                    }
                asm volatile("cp.async.commit_group;\n" ::);
                pf_base += K * plane;
              }
            const int buf = it & 1;
            if (it >= 2)
              bar_sync(BAR_AB_EMPTY + buf, TA + TB); // group B has released this buffer
            if (act)
              {
#pragma unroll 1
                for (int w = 0; w < 2; ++w)
                  {
                    const int zl = 2 * pg + w;
                    if (zl < n)
                      {
                        const double *row = U + (zl < K ? cur + zl * PLP : nxt) + off;
                        double        u[n], av[n], bv[n];
#pragma unroll
                        for (int j = 0; j < K; j += 2)
                          {
                            const double2 v = *reinterpret_cast<const double2 *>(row + j);
                            u[j] = v.x, u[j + 1] = v.y;
                          }
                        u[K] = row[K];
                        matvec<n>(Mh, u, av);
                        matvec<n>(Kh, u, bv);
                        if (seg > 0)
                          { // finish the node shared with the previous cell of the tile: its row K
                            double am = Mh[K * n + K] * u[0], ak = Kh[K * n + K] * u[0];
#pragma unroll
                            for (int j = 0; j < K; j += 2)
                              {
                                const double2 v = *reinterpret_cast<const double2 *>(row + j - K);
                                am = fma(Mh[K * n + j], v.x, am), ak = fma(Kh[K * n + j], v.x, ak);
                                am = fma(Mh[K * n + j + 1], v.y, am), ak = fma(Kh[K * n + j + 1], v.y, ak);
                              }
                            av[0] += am, bv[0] += ak;
                          }
                        double *oa = AB + (buf * 2 * n + zl) * PLP + off, *ob = oa + n * PLP;
#pragma unroll
                        for (int i = 0; i < K; i += 2)
                          {
                            *reinterpret_cast<double2 *>(oa + i) = make_double2(av[i], av[i + 1]);
                            *reinterpret_cast<double2 *>(ob + i) = make_double2(bv[i], bv[i + 1]);
                          }
                        if (seg == TX - 1)
                          oa[K] = av[K], ob[K] = bv[K];
                      }
                  }
              }
            bar_arrive(BAR_AB_FULL + buf, TA + TB);
            const int tmp = cur;
            cur = nxt, nxt = nn, nn = tmp;
          }
      }
    else if (threadIdx.x < TA + TB)
      {
        // =========================== group B: y-lines ===========================
        const int    t    = threadIdx.x - TA;
        const int    pair = t % C::pairsB, pg = t / C::pairsB;
        const int    seg  = pair / LX, off = (K * seg) * LXP + pair % LX;
        const bool   act  = pg < NPG;
        const double cm = a.cm[b], cl = a.cl[b];
        for (int it = 0; it < n_layers; ++it)
          {
            const int buf = it & 1;
            bar_sync(BAR_AB_FULL + buf, TA + TB);
            if (it >= 2)
              bar_sync(BAR_SP_EMPTY + buf, TB + TC);
            if (act)
              {
#pragma unroll 1
                for (int w = 0; w < 2; ++w)
                  {
                    const int zl = 2 * pg + w;
                    if (zl < n)
                      {
                        const double *ca = AB + (buf * 2 * n + zl) * PLP + off, *cb = ca + n * PLP;
                        double        av[n], pv[n], q[n], sv[n];
#pragma unroll
                        for (int j = 0; j < n; ++j)
                          av[j] = ca[j * LXP];
                        matvec<n>(Mh, av, pv);
                        matvec<n>(Kh, av, q);
                        if (seg > 0)
                          {
                            double pm = Mh[K * n + K] * av[0], qk = Kh[K * n + K] * av[0];
#pragma unroll
                            for (int j = 0; j < K; ++j)
                              {
                                const double ap = ca[(j - K) * LXP];
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
                          av[j] = cl * cb[j * LXP];
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
                              rm = fma(Mh[K * n + j], cl * cb[(j - K) * LXP], rm);
                            sv[0] += rm;
                          }
                        // group C only needs s and cl * p
                        double *os = SP + (buf * 2 * n + zl) * PLP + off, *op = os + n * PLP;
#pragma unroll
                        for (int i = 0; i < K; ++i)
                          os[i * LXP] = sv[i], op[i * LXP] = cl * pv[i];
                        if (seg == TY - 1)
                          os[K * LXP] = sv[K], op[K * LXP] = cl * pv[K];
                      }
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
        // one thread owns up to three node columns; the top plane of a layer is carried in registers
        const int t  = threadIdx.x - TA - TB;
        const int e0 = t, e1 = t + TC, e2 = t + 2 * TC;
        const int o0 = (e0 / LX) * LXP + e0 % LX, o1 = (e1 / LX) * LXP + e1 % LX, o2 = (e2 / LX) * LXP + e2 % LX;
        const int ncol = (e2 < LX * LY) ? 3 : (e1 < LX * LY) ? 2 : (e0 < LX * LY) ? 1 : 0;
        double    carry0 = 0.0, carry1 = 0.0, carry2 = 0.0;
        for (int it = 0; it < n_layers; ++it)
          {
            const int buf = it & 1;
            bar_sync(BAR_SP_FULL + buf, TB + TC);
            if (it >= 2)
              bar_sync(BAR_OUT_EMPTY + buf, TC + TD);
#pragma unroll 1
            for (int w = 0; w < ncol; ++w)
              {
                const int     o  = (w == 0) ? o0 : (w == 1) ? o1 : o2;
                const double *cs = SP + (buf * 2 * n) * PLP + o, *cp = cs + n * PLP;
                double        s[n], p[n], val[n];
#pragma unroll
                for (int z = 0; z < n; ++z)
                  s[z] = cs[z * PLP], p[z] = cp[z * PLP];
#pragma unroll
                for (int z = 0; z < n; ++z)
                  {
                    double acc = 0.0;
#pragma unroll
                    for (int j = 0; j < n; ++j)
                      acc = fma(Mh[z * n + j], s[j], fma(Kh[z * n + j], p[j], acc));
                    val[z] = acc;
                  }
                val[0] += (w == 0) ? carry0 : (w == 1) ? carry1 : carry2;
                if (w == 0)
                  carry0 = val[K];
                else if (w == 1)
                  carry1 = val[K];
                else
                  carry2 = val[K];
                double *ov = OUT + buf * K * PLP + o;
#pragma unroll
                for (int z = 0; z < K; ++z)
                  ov[z * PLP] = val[z];
              }
            if (it + 2 < n_layers)
              bar_arrive(BAR_SP_EMPTY + buf, TB + TC);
            bar_arrive(BAR_OUT_FULL + buf, TC + TD);
          }
      }
    else
      {
        // =========================== group D: epilogue / stores ===========================
        // streams the k finished node planes of a layer.  Interior nodes: fused epilogue E + F * Ax with
        // plain stores.  Wall nodes: the slot value E + F * partial goes to the wall arrays (E only from
        // the carrier = the contributor on the low side in x and y).  Operand pipeline: rhs (and x_old)
        // are staged TWO layers ahead with cp.async into thread-private shared slots, x one layer ahead
        // in registers (an L2 hit: group A has just streamed it), D^-1 is either loaded one layer ahead
        // or - when the caller passes dinv == NULL - the operator's own inverse diagonal, which is the
        // same for every layer of the sweep and computed once per element.
        const int     t  = threadIdx.x - TA - TB - TC;
        const double  f1 = a.f1[b], f2 = a.f2[b];
        constexpr int IX = LX - 2, IY = LY - 2, NINT = K * IX * IY, NI = (NINT + TD - 1) / TD;
        constexpr int NW = 2 * LX + 2 * IY, NWALL = K * NW, NWI = (NWALL + TD - 1) / TD;
        constexpr int NE = NI + NWI, NAS = C::nops_async(MODE);
        double       *OPS = OUT + 2 * K * PLP + t; // [ring of 3][operand][element][thread]
        int           e_s[NE], e_g[NE]; // OUT-tile offset (-1: none; bit 30: plane z == 0), DoF offset in the layer
        int           w_o[NWI], w_k[NWI]; // wall elements: slot offset, kind | carrier << 4
        constexpr bool CHEB = (MODE == V2_CHEB || MODE == V2_CHEB_OWN), own_dinv = (MODE == V2_CHEB_OWN);
        double        dloc[own_dinv ? NE : 1]; // on-the-fly inverse diagonal (dinv == NULL)
        const double *Mh_ = Mh, *Kh_ = Kh;
        auto diag_of = [&](const int cX, const int cY, const int z) {
          // assembled 1-D diagonals at tile-local node (cX, cY) and plane z of a layer (not on the domain boundary)
          const int    lx = cX % K, ly = cY % K;
          const double mx = lx ? Mh_[lx * n + lx] : Mh_[0] + Mh_[K * n + K], kx = lx ? Kh_[lx * n + lx] : Kh_[0] + Kh_[K * n + K];
          const double my = ly ? Mh_[ly * n + ly] : Mh_[0] + Mh_[K * n + K], ky = ly ? Kh_[ly * n + ly] : Kh_[0] + Kh_[K * n + K];
          const double mz = z ? Mh_[z * n + z] : Mh_[0] + Mh_[K * n + K], kz = z ? Kh_[z * n + z] : Kh_[0] + Kh_[K * n + K];
          const double d  = a.cm[b] * mx * my * mz + a.cl[b] * (kx * my * mz + mx * ky * mz + mx * my * kz);
          return (fabs(d) > 1.0e-10) ? 1.0 / d : 1.0;
        };
#pragma unroll
        for (int i = 0; i < NI; ++i)
          {
            const int q = t + i * TD;
            const int z = q / (IX * IY), r = q % (IX * IY), cY = 1 + r / IX, cX = 1 + r % IX;
            e_s[i]  = (q < NINT) ? (z * PLP + cY * LXP + cX) | (z == 0 ? (1 << 30) : 0) : -1;
            e_g[i]  = (gx0 + cX) + n1 * (gy0 + cY) + (int)plane * z;
            dloc[i] = (own_dinv) ? diag_of(cX, cY, z % K) : 0.0;
          }
#pragma unroll
        for (int i = 0; i < NWI; ++i)
          {
            const int q = t + i * TD;
            const int z = q / NW, r = q % NW;
            int       cX, cY;
            if (r < LX)
              cX = r, cY = 0;
            else if (r < 2 * LX)
              cX = r - LX, cY = LY - 1;
            else
              cY = 1 + (r - 2 * LX) / 2, cX = ((r - 2 * LX) & 1) ? LX - 1 : 0;
            const int  gx = gx0 + cX, gy = gy0 + cY;
            const bool wallx = (cX == 0) || (cX == LX - 1), wally = (cY == 0) || (cY == LY - 1);
            const int  wx = tx + (cX == 0 ? 0 : 1), dx = (cX == 0) ? 1 : 0;
            const int  wy = ty + (cY == 0 ? 0 : 1), dy = (cY == 0) ? 1 : 0;
            e_s[NI + i]  = (q < NWALL) ? z * PLP + cY * LXP + cX : -1;
            e_g[NI + i]  = gx + n1 * gy + (int)plane * z;
            dloc[NI + i] = (own_dinv) ? diag_of(cX, cY, z % K) : 0.0;
            if (wallx && wally)
              w_o[i] = (((dy * 2 + dx) * (a.ntx + 1) + wx) * (a.nty + 1) + wy) * n1 + z, w_k[i] = 3 | ((dx == 0 && dy == 0) ? 16 : 0);
            else if (wallx)
              w_o[i] = (dx * (a.ntx + 1) + wx) * (int)plane + z * n1 + gy, w_k[i] = 1 | (dx == 0 ? 16 : 0);
            else
              w_o[i] = (dy * (a.nty + 1) + wy) * (int)plane + z * n1 + gx, w_k[i] = 2 | (dy == 0 ? 16 : 0);
          }
        const bool has_xo = a.x_old != nullptr;
        // cp.async staging of rhs (and x_old) of the layer with plane-0 DoF index lay_ into ring slot sb_
        auto stage = [&](const int sb_, const long long lay_, const bool valid) {
          if (NAS > 0 && valid)
            {
#pragma unroll
              for (int i = 0; i < NE; ++i)
                if (e_s[i] >= 0)
                  {
                    const long long    j  = lay_ + e_g[i];
                    const unsigned int sp = (unsigned int)__cvta_generic_to_shared(OPS + ((sb_ * NAS) * NE + i) * TD);
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sp), "l"(a.rhs + j));
                    if (NAS > 1 && has_xo)
                      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sp + NE * TD * 8), "l"(a.x_old + j));
                  }
            }
          asm volatile("cp.async.commit_group;\n" ::);
        };
        double xc[NE], xn[NE], dn[own_dinv ? 1 : NE]; // x of the current / next layer, explicit dinv of the next layer
        auto   load_x = [&](const long long lay_, double (&xv)[NE], double (&dv)[own_dinv ? 1 : NE]) {
          if (CHEB)
            {
#pragma unroll
              for (int i = 0; i < NE; ++i)
                if (e_s[i] >= 0)
                  {
                    xv[i] = a.src[lay_ + e_g[i]];
                    if (!own_dinv)
                      dv[own_dinv ? 0 : i] = a.dinv[lay_ + e_g[i]];
                  }
            }
        };
        // running layer bases
        long long lay = boff + plane * ((long long)K * z_first); // DoF index of plane 0 of the current layer
        double   *wxb = a.WX + b * a.wx_block + (long long)(K * z_first) * n1;
        double   *wyb = a.WY + b * a.wy_block + (long long)(K * z_first) * n1;
        double   *wcb = a.WC + b * a.wc_block + (K * z_first);
        stage(0, lay, true);
        stage(1, lay + K * plane, n_layers > 1);
        double dc[own_dinv ? 1 : NE];
        load_x(lay, xc, dc);
        int sb = 0;
        for (int it = 0; it < n_layers; ++it)
          {
            const int buf = it & 1, zc = z_first + it;
            int       sb2 = sb + 2;
            sb2 -= (sb2 >= 3) ? 3 : 0;
            stage(sb2, lay + 2 * K * plane, it + 2 < n_layers);
            if (it + 1 < n_layers)
              load_x(lay + K * plane, xn, dn);
            asm volatile("cp.async.wait_group 2;\n" ::);
            // E + F * Ax for this layer's elements (E only from the carrier on walls)
            double E[NE], F[NE];
            const double *ops = OPS + (sb * NAS) * NE * TD;
#pragma unroll
            for (int i = 0; i < NE; ++i)
              {
                E[i] = 0.0, F[i] = 1.0;
                if (e_s[i] >= 0 && MODE != V2_APPLY)
                  {
                    const bool   carrier = (i >= NI) ? (w_k[i >= NI ? i - NI : 0] & 16) : true;
                    const double rh      = ops[i * TD];
                    if (MODE == V2_RESIDUAL)
                      F[i] = -1.0, E[i] = carrier ? rh : 0.0;
                    else
                      {
                        double di = own_dinv ? dloc[own_dinv ? i : 0] : dc[own_dinv ? 0 : i];
                        if (zc == 0 && (e_s[i] >> 30) && i < NI)
                          di = 1.0; // Dirichlet plane gz = 0
                        const double xo = has_xo ? ops[(NE + i) * TD] : 0.0;
                        F[i]            = -f2 * di;
                        E[i]            = carrier ? (1.0 + f1) * xc[i] - f1 * xo - F[i] * rh : 0.0;
                      }
                  }
              }
            bar_sync(BAR_OUT_FULL + buf, TC + TD);
            if (zc >= zb)
              {
                const double *tile = OUT + buf * K * PLP;
#pragma unroll
                for (int i = 0; i < NE; ++i)
                  if (e_s[i] >= 0)
                    {
                      const long long j = lay + e_g[i];
                      double          v = tile[e_s[i] & 0xffffff];
                      if (i < NI)
                        {
                          if (zc == 0 && (e_s[i] >> 30)) // Dirichlet plane gz = 0: A is the identity there
                            v = CHEB ? xc[i] : a.src[j];
                          a.dst[j] = fma(F[i], v, E[i]);
                        }
                      else
                        {
                          const int kind = w_k[i >= NI ? i - NI : 0] & 3;
                          double   *slot = (kind == 1 ? wxb : kind == 2 ? wyb : wcb) + w_o[i >= NI ? i - NI : 0];
                          *slot          = fma(F[i], v, E[i]);
                        }
                    }
              }
            if (it + 2 < n_layers)
              bar_arrive(BAR_OUT_EMPTY + buf, TC + TD);
            lay += K * plane, wxb += K * n1, wyb += K * n1, wcb += K;
            sb = (sb == 2) ? 0 : sb + 1;
#pragma unroll
            for (int i = 0; i < NE; ++i)
              {
                xc[i] = xn[i];
                if (!own_dinv)
                  dc[own_dinv ? 0 : i] = dn[own_dinv ? 0 : i];
              }
          }
        // top plane of the domain (Dirichlet): interior node columns of the last chunk
        if (ze == nc)
          for (int r = t; r < IX * IY; r += TD)
            {
              const long long gi = (gx0 + 1 + r % IX) + (long long)n1 * (gy0 + 1 + r / IX) + plane * (n1 - 1);
              const double    x  = src[gi];
              v2_epilogue(a, b, boff + gi, x, x);
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
        SPIRK_CUDA(cudaFuncSetAttribute(k_v2_main<K, TX, TY, V2_APPLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::smem(V2_APPLY)));
        SPIRK_CUDA(cudaFuncSetAttribute(k_v2_main<K, TX, TY, V2_RESIDUAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::smem(V2_RESIDUAL)));
        SPIRK_CUDA(cudaFuncSetAttribute(k_v2_main<K, TX, TY, V2_CHEB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::smem(V2_CHEB)));
        SPIRK_CUDA(cudaFuncSetAttribute(k_v2_main<K, TX, TY, V2_CHEB_OWN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::smem(V2_CHEB_OWN)));
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
      k_v2_main<K, TX, TY, V2_APPLY><<<(unsigned int)grid, C::threads, C::smem(V2_APPLY), ctx->stream>>>(a);
    else if (a.mode == V2_RESIDUAL)
      k_v2_main<K, TX, TY, V2_RESIDUAL><<<(unsigned int)grid, C::threads, C::smem(V2_RESIDUAL), ctx->stream>>>(a);
    else if (a.dinv != nullptr)
      k_v2_main<K, TX, TY, V2_CHEB><<<(unsigned int)grid, C::threads, C::smem(V2_CHEB), ctx->stream>>>(a);
    else
      k_v2_main<K, TX, TY, V2_CHEB_OWN><<<(unsigned int)grid, C::threads, C::smem(V2_CHEB_OWN), ctx->stream>>>(a);
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
    if (g.dim != 3 || g.k != 4 || op->kind != SPIRK_OP_REAL || g.nc % 8 != 0 || g.nc < 8)
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
    if (ctx->opt_apply_variant == 3)
      return v2_launch<4, 4, 4>(ctx, a);
    return v2_launch<4, 8, 2>(ctx, a);
  }
} // namespace spirk
