// Matrix-free cell operator, variant 3: owner-computes plane streaming (3-D, uniform Cartesian mesh).
//
// The operator on the refined hypercube is a sum of Kronecker products of the GLOBAL banded 1-D
// matrices (assembled from the reference cell matrices Mh, Kh):
//   A = cm Mz My Mx + cl (Mz My Kx + Mz Ky Mx + Kz My Mx),   cm = mass h^3, cl = laplace h
// which factors into 7 one-dimensional banded sweeps
//   x :  a' = cl Mx u,             c = (cm Mx + cl Kx) u
//   y :  p' = My a',               w = My c + Ky a'
//   z :  out = Mz w + Kz p'
// A row of a banded 1-D matrix has 2k+1 entries at a vertex node (shared by two cells) and k+1 at a
// cell-interior node: 6 FMA per node and sweep at k = 4, about 44 FMA per DoF in total.
//
// Work decomposition.  A CTA owns a column of TX x TY cells (k TX x k TY node columns, "owner
// computes": nodes [k TX tx, k TX (tx+1)) x [k TY ty, ...)) and streams it upwards in z one NODE
// PLANE at a time through a ring of shared-memory slots:
//   stage   : the (k TX + k + 1) x (k TY + k + 1) staged nodes of `src` (owned + one cell of halo below, one node above)
//             arrive by TMA 2-D tile copies (cp.async.bulk.tensor, completion on an mbarrier), issued NBUF - 1 planes
//             ahead.  n1 is odd, so rows start at arbitrary 8-byte parity and no tensor map of row pitch n1 exists: the
//             maps view PAIRS of rows as super-rows of pitch 2 n1, the staged rows of even and of odd parity are one box
//             each, started at the 16-byte aligned element at or below the first wanted node (the consumers add the
//             parity, a compile-time function of the plane's position in its cell layer).  The rows of the epilogue
//             operands (rhs, x_old) of the same plane travel in the same slot.  No thread ever waits on a global load.
//   x-phase : one thread per (staged row, cell segment): 2k+1 loads, the k owned outputs of a', c.
//             Dirichlet / out-of-domain nodes are masked here (rows, planes and the two x-columns).  The warps without
//             an x task combine the pointwise operands of the Chebyshev epilogue in their ring slot meanwhile.
//   y+z     : one thread per (owned x, cell segment in y[, half of its nodes]), lanes along x: 2(2k+1) loads, the
//             outputs of p', w, and IMMEDIATELY their contribution to the z-sums of the current
//             cell layer, which live in registers ((k+1) planes x NPT nodes).  The linear part of the
//             fused epilogue (residual: rhs; Chebyshev: x, x_old, rhs) is folded into the z-sums
//             when the plane passes, so when the top plane of a layer has been added, planes 0..k-1
//             are final up to one scaling and are stored straight from registers, 256-byte
//             coalesced per warp.  The top plane's sums carry over to the next layer.
// No partial sums are accumulated in memory: no atomics, no zero-init pass, no second kernel; results are bitwise
// reproducible.  The price is the halo: (37/32)^2 of the plane is staged and 37/32 of the x-sweeps are computed for an
// 8 x 8-cell tile.
// Schedule: (block, layer range, column) work items drawn from an atomic counter by a fixed grid of 2 CTAs per SM, long
// ranges first and short ones for the tail; the two CTAs that meet at a range boundary exchange the partial z-sums of the
// shared vertex plane through a small buffer (nothing is recomputed; V3Args::dyn).  The even split and the z-lockstep
// schedule of round 1 (a piece that starts above layer 0 recomputes the layer below it) remain as options.
// Coupled pairs of blocks (NBC = 2: IRK q = 2 system matrix, complex level operators, K v + M w) stage the plane of both
// blocks and mix the mass sweeps in the x-phase.
//
// Replaces the deal.II cell loop + vector updates of the reference:
//   operator.h:298-310, 379-421 (vmult), 841-880 (batched); deal.II PreconditionChebyshev
//   vector_updates and the residual of Multigrid::level_v_step (A8).
#pragma once
#include <cuda.h>

#include <algorithm>
#include <cstdint>
#include <functional>
#include <vector>
#include <type_traits>

#include "op_v2.cuh"

namespace spirk
{
  __host__ __device__ constexpr int v3_nops(const int mode) { return mode == V2_RESIDUAL ? 1 : (mode == V2_CHEB_OWN ? 2 : 0); }

  // rows handled side by side by the epilogue pre-pass: the largest divisor of oy that is a multiple of k, even, and <= avail
  __host__ __device__ constexpr int v3_pre_rows(const int avail, const int oy, const int k)
  {
    int best = 0;
    for (int r = k; r <= oy && r <= avail; r += k)
      if (oy % r == 0 && r % 2 == 0)
        best = r;
    return best;
  }

  template <int K, int TX, int TY, int MODE, int NPT = K, int NBC = 1>
  struct CfgV3
  {
    static constexpr int n = K + 1, OX = K * TX, OY = K * TY, LXS = OX + K + 1, LYS = OY + K + 1;
    static constexpr int BW   = (LXS + 2) & ~1; // staged box: LXS nodes from a 16-byte aligned start, whole 16-byte chunks ...
    static constexpr int BH   = (LYS + 1) / 2;  // ... x every other row (even / odd rows are separate boxes, see v3_make_map)
    static constexpr int UB   = ((BW * BH + 15) / 16) * 16; // one box, padded to 128 bytes
    static constexpr int OW   = OX + 2;         // operand box: the owned nodes (+ alignment slack) of every other row
    static constexpr int OB   = OW * (OY / 2);
    static constexpr int PA   = OX | 1;         // odd pitch: lanes along y hit distinct banks
    static constexpr int NOPS = v3_nops(MODE);  // operand planes that travel with the staged plane
    // build-time overrides for kernel experiments (tools/build_variant.py): -DSPIRK_V3_NBUF / _NAC / _MINB
#ifdef SPIRK_V3_NBUF
    static constexpr int NBUF = SPIRK_V3_NBUF;
#else
    // ring depth (three / two planes in flight; 2 CTAs per SM must fit; coupled pairs stage the plane of both blocks, so
    // with operand planes in the slot only one plane is in flight)
    static constexpr int NBUF = (NBC > 1 && NOPS > 0) ? 2 : ((NOPS == 0 && NBC == 1) ? 4 : 3);
#endif
#ifdef SPIRK_V3_NAC
    static constexpr int NAC = SPIRK_V3_NAC;
#else
    static constexpr int NAC  = (NOPS == 2) ? 1 : 2; // a'/c tile double-buffered where shared memory allows
#endif
    // software pipeline: the x-phase of plane s + 1 runs in the same barrier interval as the y+z phase of plane s
    // (one __syncthreads per plane, two independent tasks per thread); needs the double-buffered tile
#ifndef SPIRK_V3_PIPE
#define SPIRK_V3_PIPE 0
#endif
    static constexpr bool PIPE = (NAC == 2) && (SPIRK_V3_PIPE != 0);
    // one code copy for the planes 1 .. K-1 of a cell layer (runtime plane position, z-coefficients from a shared-memory
    // table) instead of K-1 specialised copies: the plane loop of the fused modes is 67 KB of SASS, twice the 32 KB
    // instruction cache level (ncu: 1.65 "no instruction" stall cycles per issued instruction)
#ifndef SPIRK_V3_MIDLOOP
#define SPIRK_V3_MIDLOOP 0
#endif
    static constexpr bool MIDLOOP = (SPIRK_V3_MIDLOOP != 0) && (K >= 3);
    static constexpr int  NZT = MIDLOOP ? 2 * (K + 1) * (K + 1) : 0; // {M(z, zl), K'(z, zl)} by plane position zl
    // coupled operators (NBC > 1 blocks, apply only): the x-phase mixes the mass sweeps of ALL blocks, so a ring slot
    // holds the staged plane of every block
    static constexpr int OPB  = NBC * 2 * UB; // the operand planes follow the staged planes in the ring slot
    static constexpr int SLOT = OPB + NOPS * 2 * OB;
    static constexpr int NHALF = K / NPT;          // a cell segment in y is shared by NHALF threads of NPT nodes each
    static constexpr int NY   = OX * TY * NHALF; // y+z tasks (one thread each)
    static constexpr int NXT  = LYS * TX; // x-phase tasks
    static constexpr int NT   = ((NXT > NY ? NXT : NY) + 31) / 32 * 32; // one x task per thread; the threads beyond NY are helpers:
                                                                         // TMA issue, Dirichlet faces
    static constexpr int NH   = NT - NY;
#ifdef SPIRK_V3_MINB
    static constexpr int MINB = SPIRK_V3_MINB;
#else
    static constexpr int MINB = (NT > 256) ? 2 : (NT > 128 ? 3 : 6);
#endif
    static constexpr unsigned BYTES_U = NBC * 2 * BW * BH * 8, BYTES_O = 2 * OB * 8;
    static constexpr size_t   smem = 128 + sizeof(double) * (size_t)(NBUF * SLOT + NAC * 2 * LYS * PA + 3 * K * K * K + NZT + NBUF + 2);
    static_assert(NY % 32 == 0 && NT <= 1024 && K % NPT == 0 && NPT <= 4 && NBC <= 2, "tile shape");
    static constexpr int ISSUER = (NH > 0) ? NY : 0; // the thread that issues the TMA copies
    static_assert(UB % 16 == 0 && OB % 16 == 0 && OY % 2 == 0 && BW <= 256 && BH <= 256, "128-byte aligned TMA boxes");
  };

  // Mh / Kh are symmetric and persymmetric (made exact at upload): index the canonical copy of an entry so
  // that the compiler loads each distinct value once (FP64 instructions take no constant-bank operands)
  template <int K>
  __host__ __device__ constexpr int v3_canon(const int i, const int j)
  {
    int a = i < j ? i : j, b = i < j ? j : i;
    const int a2 = K - b, b2 = K - a;
    if (a2 < a || (a2 == a && b2 < b))
      a = a2, b = b2;
    return a * (K + 1) + b;
  }

  // position of the canonical copy of entry (i, j) in the compact list of distinct entries
  template <int K>
  __host__ __device__ constexpr int v3_cidx(const int i, const int j)
  {
    const int c   = v3_canon<K>(i, j);
    int       cnt = 0;
    for (int a = 0; a <= K; ++a)
      for (int b = a; b <= K; ++b)
        if (v3_canon<K>(a, b) == a * (K + 1) + b)
          {
            if (a * (K + 1) + b == c)
              return cnt;
            ++cnt;
          }
    return -1;
  }
  constexpr int V3_F0C = 8;  // coupled pairs: f0 of block b sits in cc[V3_F0C + b]
  constexpr int V3_NKP = 12; // distinct entries of K' per block (9 at k = 4, the only instantiated degree) + the vertex
                             // diagonal in the last slot

  static_assert(v3_cidx<4>(4, 4) < V3_NKP - 1 && v3_cidx<4>(2, 2) < V3_NKP - 1 && v3_cidx<4>(1, 3) < V3_NKP - 1, "compact coefficient list");

  struct V3Args
  {
    Geo           g;
    int           nb;
    long long     stride;
    double       *dst;
    const double *src, *x_old, *rhs, *dinv;
    double        cm[SPIRK_MAX_BLOCKS], cl[SPIRK_MAX_BLOCKS], f1[SPIRK_MAX_BLOCKS], f2[SPIRK_MAX_BLOCKS];
    // fused Chebyshev modes: D = diagonal of dm M + dl K per node class (h-scaled).  The operator's own diagonal by
    // default; the coefficients a stored inverse diagonal was computed with when the caller names them (the smoother
    // of a level operator whose coefficients changed after the multigrid set-up)
    double        dm[SPIRK_MAX_BLOCKS], dl[SPIRK_MAX_BLOCKS];
    // A = sc (Mz My K'x + Mz K'y Mx + K'z My Mx) with K' = Kh + cm / (3 cl) Mh, sc = cl  (cl = 0: K' = Mh / 3, sc = cm):
    // the mass term rides in the three stiffness terms, so no sweep scales its result
    double        sc[SPIRK_MAX_BLOCKS], kp[SPIRK_MAX_BLOCKS][V3_NKP];
    double        cc[16]; // coupled operators (<= 4 blocks): coupling * h^3, row-major; V2_CHEB_FIRST: f0 of the blocks
                          // (x1 = f0 dinv src; x1 is written to the `dinv` pointer, which that mode does not read) - in
                          // cc[b], for a coupled pair in cc[V3_F0C + b]
    int           coupled;
    int           km; // coupled pair whose second input is block b of ANOTHER vector (map tm_o0): dst_b = cl_b K u_b + cc[2b+1] M w_b
    int           ntx, nty;
    long long     W; // nb * columns * layers
    int           lock_nch, lock_len; // > 0: z-lockstep schedule (CTA = one column x one of lock_nch equal layer ranges)
    // dyn != 0: work-queue schedule.  Items = (block, layer range of lock_len layers, column), drawn from an atomic counter in
    // that order (concurrent CTAs work on neighbouring columns at similar heights); a range that starts above layer 0 does
    // NOT recompute the layer below: the two CTAs that meet at a range boundary exchange the partial z-sums of the shared
    // vertex plane through `carry` (whoever is ready first publishes, the other one adds and stores; a + b is commutative,
    // so the result is bitwise reproducible)
    int           dyn, n_items;
    // the ranges of a column: n_big ranges of lock_len layers, then lock_nch - n_big ranges of len_small layers; all long
    // items are drawn before the short ones, which fill the tail of the launch (items_big = nb * columns * n_big)
    int           n_big, len_small, items_big;
    int          *sched;  // [0] next item, [1] CTAs that are done (the last one resets both)
    int          *state;  // per range boundary: 0 idle, 1 claimed by the first arrival, 2 its partial sums are published
    double       *carry;  // per range boundary: OX * OY partial sums
    long long     rows_per_block; // stride / n1: the blocks continue the row sequence of block 0
    int           L_lo, L_hi;     // cell layers of this z-slab ([0, nc) for the whole mesh)
    int           zo0;            // global index of the first owned node plane = local plane 0 of every vector
    int           gh_lo;          // ghost planes below the owned range (the tensor maps start there)
    int           sh_src, sh_o0, sh_o1; // element shift of the 16-byte aligned map base below the vector
    int           pb[SPIRK_MAX_BLOCKS]; // parity of the global row of staged row 0 of an EVEN node plane of block b (set by
                                        // v3_launch_mode; K and the tile origins are even, n1 is odd: the parity alternates
                                        // from plane to plane and is a compile-time function of the plane's position ZL)
    alignas(64) CUtensorMap tm_src, tm_o0, tm_o1; // staged nodes; operand 0 (rhs | x_old); operand 1 (rhs)
  };

  static_assert(sizeof(V3Args) <= 4096, "kernel parameter space");

  __device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
  __device__ __forceinline__ void     mbar_init(const unsigned bar, const int count)
  {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
  }
  __device__ __forceinline__ void mbar_arrive_expect_tx(const unsigned bar, const unsigned bytes)
  {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
  }
  __device__ __forceinline__ bool mbar_try_wait(const unsigned bar, const unsigned parity)
  {
    unsigned ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity)
                 : "memory");
    return ok != 0;
  }
  // TMA 2-D tile copy global -> shared (box of the tensor map at element coordinates c0, c1), completion
  // counted in bytes on an mbarrier; out-of-bounds elements are zero-filled
  __device__ __forceinline__ void tma_g2s_2d(const unsigned dst, const CUtensorMap *map, const int c0, const int c1,
                                             const unsigned bar)
  {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(dst),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar)
                 : "memory");
  }

  // identity rows (Dirichlet nodes): A x = x
  template <int MODE>
  __device__ __forceinline__ void v3_identity(const V3Args &a, const double f1, const double f2, const long long j, const double f0 = 0.0)
  {
    const double x = a.src[j];
    if (MODE == V2_APPLY)
      a.dst[j] = x;
    else if (MODE == V2_RESIDUAL)
      a.dst[j] = a.rhs[j] - x;
    else if (MODE == V2_CHEB_FIRST)
      {
        const double x1 = f0 * x; // the inverse diagonal is 1 on Dirichlet nodes
        const_cast<double *>(a.dinv)[j] = x1;
        a.dst[j]        = fma(f2, x - x1, fma(f1, x1, x1));
      }
    else
      {
        const double xo = a.x_old ? a.x_old[j] : 0.0;
        const double di = (MODE == V2_CHEB) ? a.dinv[j] : 1.0;
        a.dst[j]        = fma(f2 * di, a.rhs[j] - x, fma(f1, x - xo, x));
      }
  }

  template <int K, int TX, int TY, int MODE, int NPT, int NBC>
  __global__ void __launch_bounds__(CfgV3<K, TX, TY, MODE, NPT, NBC>::NT, CfgV3<K, TX, TY, MODE, NPT, NBC>::MINB)
    k_v3(const __grid_constant__ V3Args a)
  {
    using C = CfgV3<K, TX, TY, MODE, NPT, NBC>;
    constexpr int n = C::n, OX = C::OX, OY = C::OY, LYS = C::LYS, BW = C::BW, UB = C::UB, OW = C::OW, OB = C::OB, PA = C::PA, NT = C::NT;
    constexpr int NOPS = C::NOPS, NBUF = C::NBUF, NAC = C::NAC, SLOT = C::SLOT, NY = C::NY, NH = C::NH, OPB = C::OPB;
    extern __shared__ __align__(16) double sm3_raw[];
    double   *sm3  = sm3_raw + (((128u - (smem_u32(sm3_raw) & 127u)) & 127u) >> 3); // TMA boxes: 128-byte aligned
    double   *RING = sm3, *AC = sm3 + NBUF * SLOT, *SDS = AC + NAC * 2 * LYS * PA, *SDI = SDS + K * K * K, *SD1 = SDI + K * K * K;
    double   *ZT   = SD1 + K * K * K; // (MIDLOOP) z-coefficients by runtime plane position
    uint64_t *BAR  = reinterpret_cast<uint64_t *>(ZT + C::NZT);
    int      *QS   = reinterpret_cast<int *>(BAR + NBUF); // [0] drawn item, [2] role at a range boundary
    const double *Mh = c_fe[K].Mh, *Kh = c_fe[K].Kh;
    const double  Mv = c_fe[K].Mv;
#define MC(i, j) Mh[v3_canon<K>(i, j)]
#define KC(i, j) kp[v3_cidx<K>(i, j)]

    const int       tid = threadIdx.x;
    const int       wrp = __shfl_sync(0xffffffffu, tid >> 5, 0); // warp index, provably warp-uniform
    const int       n1 = a.g.n1, nc = a.g.nc;
    const long long plane = (long long)n1 * n1;
    const int       ncols = a.ntx * a.nty;
    const bool      is_yz = tid < NY;                    // the other threads only take an x task, the TMA issue and the faces
    const int       xl = tid % OX, ys = (tid / OX) % TY; // y+z task: owned x, cell segment in y,
    const int       half = (tid / OX) / TY, i0 = half * NPT; // ... and which NPT of its K nodes (warp-uniform)

    if (tid == 0)
      {
#pragma unroll
        for (int i = 0; i < NBUF; ++i)
          mbar_init(smem_u32(BAR + i), 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
      }
    const unsigned ring_u32 = smem_u32(RING), bar_u32 = smem_u32(BAR);

    // work range in the (block, column, layer) space: an even split of the whole space, or - for vectors beyond
    // the L2 capacity - one column x one layer range per CTA with neighbouring CTAs on neighbouring columns at
    // the same height, so that the halo rows / columns shared by adjacent tiles meet in L2
    long long w = a.W * blockIdx.x / gridDim.x, w_end = a.W * (blockIdx.x + 1) / gridDim.x;
    if (a.lock_nch > 0 && !a.dyn)
      {
        const long long colb = blockIdx.x % ((long long)ncols), rest = blockIdx.x / ncols;
        const int       ch = (int)(rest % a.lock_nch), bb = (int)(rest / a.lock_nch);
        w     = ((long long)bb * ncols + colb) * nc + min(nc, ch * a.lock_len);
        w_end = ((long long)bb * ncols + colb) * nc + min(nc, (ch + 1) * a.lock_len);
      }
    int             b_tab = -1;
    unsigned        it0   = 0; // ring position of the piece's first plane (runs on across pieces)
    for (;;)
      {
        // ------------------------------------------------------------------ next piece: block b, column col, layers [L0, L1)
        int L0, L1, col, b;
        if (a.dyn)
          {
            // items are drawn from a queue (measured against dealing them round robin: 17 - 20 % faster at r = 6, 7; the
            // CTAs of one SM do not progress at the same rate)
            __syncthreads(); // the previous piece is finished (ring, a'/c tile, tables, QS)
            if (tid == 0)
              QS[0] = atomicAdd(a.sched, 1);
            __syncthreads();
            const int item = __shfl_sync(0xffffffffu, QS[0], 0); // (warp-uniform by construction: lets the compiler keep the piece's
                                                                 // indices and the per-block coefficients in uniform registers)
            if (item >= a.n_items)
              break;
            if (item < a.items_big)
              {
                col            = item % ncols;
                const int rest = item / ncols, ch = rest % a.n_big;
                b              = rest / a.n_big;
                L0 = min(a.L_hi, a.L_lo + ch * a.lock_len), L1 = min(a.L_hi, L0 + a.lock_len);
              }
            else
              {
                const int it = item - a.items_big, n_small = a.lock_nch - a.n_big;
                col            = it % ncols;
                const int rest = it / ncols, ch = rest % n_small;
                b              = rest / n_small;
                L0 = min(a.L_hi, a.L_lo + a.n_big * a.lock_len + ch * a.len_small), L1 = min(a.L_hi, L0 + a.len_small);
              }
          }
        else
          {
            if (w >= w_end)
              break;
            L0                 = (int)(w % nc);
            const long long t_ = w / nc;
            col = (int)(t_ % ncols), b = (int)(t_ / ncols);
            L1 = (int)min((long long)nc, L0 + (w_end - w));
            w += L1 - L0;
          }
        const int       tx = col % a.ntx, ty = col / a.ntx;
        const int       gx0 = tx * OX, gy0 = ty * OY;
        const double    cl = a.cl[b], f1 = a.f1[b], f2 = a.f2[b], sc = a.sc[b], sc_inv = 1.0 / sc;
        const double   *kp = a.kp[b];
        constexpr bool  CF = (MODE == V2_CHEB_FIRST);
        const double    f0 = CF ? a.cc[(NBC > 1 ? V3_F0C : 0) + b] : 0.0, kappa = CF ? (1.0 + (1.0 + f1) * f0 / f2) * sc_inv : 0.0;
        const double    Kv = kp[V3_NKP - 1];
        const long long boff = (long long)b * a.stride;
        const double   *src  = a.src + boff;
        // coupled pair: the staged plane of the block this item produces inside a ring slot (km: plane 0 = u_b)
        const int       bo   = (NBC > 1 && !a.km) ? b * (2 * UB) : 0;
        // range boundaries shared with another CTA of this launch (at the bottom of a z-slab the layer below is recomputed
        // from the ghost planes; the top plane of a slab belongs to the slab above)
        const bool      carry_in = a.dyn && L0 > a.L_lo, carry_out = a.dyn && L1 < a.L_hi;
        const int       zf   = (L0 > 0 && !carry_in) ? L0 - 1 : L0; // first layer that is processed (recomputed if < L0)
        const int       nsteps = 1 + K * (L1 - zf);        // node planes K zf .. K L1
        const bool      has_xo = (a.x_old != nullptr);
        const bool      edge_x = (tx == 0) || (tx == a.ntx - 1);

        if (!a.dyn)
          __syncthreads(); // the previous piece is finished with the ring, the a'/c tile and the tables
        if (C::MIDLOOP && b != b_tab)
          {
            for (int e = tid; e < n * n; e += NT)
              {
                const int z = e / n, zl = e % n;
                ZT[2 * (zl * n + z)] = Mh[v3_canon<K>(z, zl)], ZT[2 * (zl * n + z) + 1] = kp[v3_cidx<K>(z, zl)];
              }
            if (!(MODE == V2_CHEB_OWN || MODE == V2_CHEB_FIRST))
              b_tab = b; // (read after the first barrier of the plane loop at the earliest)
          }
        if ((MODE == V2_CHEB_OWN || MODE == V2_CHEB_FIRST) && b != b_tab)
          {
            // scaling by node class (position of the node inside its cell): -f2 / diag and its inverse
            for (int e = tid; e < K * K * K; e += NT)
              {
                const int    cx = e % K, cy = (e / K) % K, cz = e / (K * K);
                const double Kv0 = c_fe[K].Kv;
                const double mx = cx ? Mh[cx * n + cx] : Mv, kx = cx ? Kh[cx * n + cx] : Kv0;
                const double my = cy ? Mh[cy * n + cy] : Mv, ky = cy ? Kh[cy * n + cy] : Kv0;
                const double mz = cz ? Mh[cz * n + cz] : Mv, kz = cz ? Kh[cz * n + cz] : Kv0;
                const double d  = a.dm[b] * mx * my * mz + a.dl[b] * (kx * my * mz + mx * ky * mz + mx * my * kz);
                const double di = (fabs(d) > 1.0e-10) ? 1.0 / d : 1.0;
                SDS[e] = -f2 * di * sc, SDI[e] = 1.0 / (f2 * di * sc), SD1[e] = f0 * di;
              }
            b_tab = b;
            if (CF)
              __syncthreads(); // the x-phase of the first plane already reads the table
          }

        // ------------------------------------------------------------------ staging of one node plane (one thread)
        // Rows of all planes and blocks form ONE sequence of pitch n1 (row index R = y + n1 (z + n1 b)); the tensor
        // maps view pairs of rows as super-rows of pitch 2 n1 (a multiple of 16 bytes although n1 is odd), so the
        // even and the odd rows of a staged plane are one 2-D box each.
        const long long Rb0 = (long long)b * a.rows_per_block + (long long)n1 * (K * zf - a.zo0 + a.gh_lo) + (gy0 - K); // staged row 0 of step 0
        const bool      has_o0 = (MODE == V2_RESIDUAL) || (MODE == V2_CHEB_OWN && has_xo);
        auto            issue  = [&](const int sidx) {
          if (tid == C::ISSUER)
            {
              const int       P     = K * zf + sidx;
              const unsigned  slot  = (it0 + sidx) % NBUF;
              const bool      zpl   = (P <= 0) || (P >= n1 - 1);
              const bool      owned = (NOPS > 0) && !zpl && (P >= K * L0) && (P < K * L1);
              const unsigned  bar   = bar_u32 + 8 * slot, dst = ring_u32 + slot * (SLOT * 8);
              const long long Rb    = Rb0 + (long long)n1 * sidx, Ro = Rb + K;
              if (MODE == V2_CHEB_OWN) // the slot was written through the generic proxy (g over rhs)
                asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
              mbar_arrive_expect_tx(bar, (zpl ? 0u : C::BYTES_U) + (owned ? ((has_o0 ? 1u : 0u) + (NOPS == 2 ? 1u : 0u)) * C::BYTES_O : 0u));
              if (!zpl)
                {
#pragma unroll
                  for (int j = 0; j < NBC; ++j)
                    {
                      // coupled: the plane of every block (block j starts rows_per_block * (j - b) rows away)
                      const bool         second = (NBC > 1) && a.km && (j == 1); // the plane of w_b (same rows as u_b)
                      const long long    Rj  = (NBC == 1 || a.km) ? Rb : Rb + (long long)(j - b) * a.rows_per_block;
                      const CUtensorMap *tm  = second ? &a.tm_o0 : &a.tm_src;
                      const int          shj = second ? a.sh_o0 : a.sh_src;
                      // box t = the staged rows of parity t (global rows Rj + t, Rj + t + 2, ...: one half of the super-rows)
#pragma unroll
                      for (int t = 0; t < 2; ++t)
                        {
                          const long long G = Rj + t;
                          tma_g2s_2d(dst + j * (2 * UB * 8) + t * (UB * 8), tm, ((int)(G & 1) * n1 + gx0 - K + shj) & ~1, (int)(G >> 1), bar);
                        }
                    }
                }
              if (owned)
                {
#pragma unroll
                  for (int t = 0; t < 2; ++t)
                    {
                      const long long G = Ro + t; // operand rows of parity t
                      if (has_o0)
                        tma_g2s_2d(dst + (OPB + t * OB) * 8, &a.tm_o0, ((int)(G & 1) * n1 + gx0 + a.sh_o0) & ~1, (int)(G >> 1), bar);
                      if (NOPS == 2)
                        tma_g2s_2d(dst + (OPB + (2 + t) * OB) * 8, &a.tm_o1, ((int)(G & 1) * n1 + gx0 + a.sh_o1) & ~1, (int)(G >> 1), bar);
                    }
                }
            }
        };
        const int pb = a.pb[b];

        double acc[n][NPT]; // z-sums of the current layer: planes 0..K x the NPT owned nodes of this thread
#pragma unroll
        for (int z = 0; z < n; ++z)
#pragma unroll
          for (int i = 0; i < NPT; ++i)
            acc[z][i] = 0.0;

        for (int s = 0; s < NBUF && s < nsteps; ++s)
          issue(s);

        // running state of the plane loop (kept incrementally: no div / mod per step)
        unsigned slot = it0 % NBUF, phase = (it0 / NBUF) & 1; // ring slot of the plane and its mbarrier phase
        int      P    = K * zf;                              // node plane
        int      sbuf = 0;                                   // a'/c tile buffer
        int      s    = 0;

        // one node plane; ZL = position of the plane in its cell layer (0 only for the first plane of a piece),
        // Lc = the layer the plane belongs to as plane ZL (for ZL == K: the layer it completes)
        // shared-memory offset of staged row r / operand row ro of a plane whose staged row 0 has parity par_: box of the
        // row's parity, position inside the box, and the 8-byte parity of the row start (a box starts at the 16-byte
        // aligned element at or below the first wanted one; gx0, K even, n1 odd)
        // (the box is chosen by the parity of the STAGED row, so only the one-element alignment shift depends on the plane:
        // a box must start at a 16-byte aligned global address - an odd FP64 start coordinate faults, tools/tma_test.cu)
        auto urow = [&](const int r, const int par_, const int sh = -1) {
          return (r & 1) * UB + (r >> 1) * BW + (((sh < 0 ? a.sh_src : sh) ^ par_ ^ r) & 1);
        };
        auto orow = [&](const int ro, const int par_, const int sh) {
          // K is even: operand row 0 (= staged row K) has the parity of staged row 0
          return (ro & 1) * OB + (ro >> 1) * OW + ((sh ^ par_ ^ ro) & 1);
        };
        // x-phase of the plane xP staged in ring slot xslot: a = Mx u, c = K'x u into the tile buffer xbuf
        auto xphase = [&](const unsigned xslot, const unsigned xphase_bit, const int xpar, const int xP, const int xbuf,
                          const int xzl = 1) {
          const bool    xzpl = (xP <= 0) || (xP >= n1 - 1);
          const double *xub  = RING + xslot * SLOT;
          double       *XA = AC + (NAC == 2 ? xbuf * (2 * LYS * PA) : 0), *XC = XA + LYS * PA;
          while (!mbar_try_wait(bar_u32 + 8 * xslot, xphase_bit))
            ;
          if constexpr (MODE == V2_CHEB_FIRST)
            {
              // the first iterate x1 = f0 D^-1 b of the owned nodes of the plane is stored here, by the warps without an x task
              constexpr int PRE0 = ((C::NXT + 31) / 32 * 32 < NT) ? (C::NXT + 31) / 32 * 32 : 0, NPRE = NT - PRE0;
              const bool    xowned = (xzl > 0 || carry_in) && !xzpl && (xP >= K * L0) && (xP < K * L1);
              if (xowned && wrp >= PRE0 / 32)
                {
                  const double *sdz = SD1 + (xzl % K) * (K * K);
                  double       *d1p = const_cast<double *>(a.dinv) + boff + plane * (xP - a.zo0);
                  constexpr int TOT = OY * (OX / 2);
#pragma unroll
                  for (int k = 0; k < (TOT + NPRE - 1) / NPRE; ++k)
                    if (tid - PRE0 + k * NPRE < TOT)
                      {
                        const int     e = tid - PRE0 + k * NPRE;
                        const int     ro = e / (OX / 2), x0 = (e % (OX / 2)) * 2;
                        const double *pu = xub + bo + urow(K + ro, xpar) + K + x0;
                        const double *sd = sdz + (ro % K) * K + (x0 % K);
                        const int     gx = gx0 + x0, gy = gy0 + ro;
                        double       *dp = d1p + gx + (long long)n1 * gy;
                        if (gy != 0)
                          {
                            if (gx != 0)
                              dp[0] = sd[0] * pu[0];
                            dp[1] = sd[1] * pu[1];
                          }
                      }
                }
            }
          if constexpr (MODE == V2_CHEB_OWN)
            {
              // ---------------------------------------------------------- pointwise part of the Chebyshev epilogue,
              // g = (rhs + ((1 + f1) x - f1 x_old) / (f2 dinv)) / sc for the owned nodes of the plane, written over the rhs
              // operand in its ring slot, so that the y+z phase only subtracts g from its z-sums.  Done by warps that have
              // no x task (the x-phase occupies NXT of the NT threads): thread = (x mod OX/2, row mod RG), nodes (x, x + OX/2)
              // of the rows row + RG j - every node of a thread has the same node class (RG and OX/2 are multiples of K) and
              // the same row parity, all offsets but one base per operand are compile-time.
              constexpr int HX = OX / 2, PRE0 = ((C::NXT + 31) / 32 * 32 + 128 <= NT) ? (C::NXT + 31) / 32 * 32 : 0;
              constexpr int RG = v3_pre_rows((NT - PRE0) / HX, OY, K); // rows handled side by side
              static_assert(HX % K == 0 && RG > 0 && RG % K == 0 && RG % 2 == 0 && OY % RG == 0 && RG * HX <= NT - PRE0 && (RG * HX) % 32 == 0, "pre-pass thread layout");
              const bool    xowned = (xzl > 0 || carry_in) && !xzpl && (xP >= K * L0) && (xP < K * L1);
              if (xowned && wrp >= PRE0 / 32 && wrp < (PRE0 + RG * HX) / 32)
                {
                  const int     t_ = tid - PRE0, kx = t_ % HX, rs = t_ / HX; // rs < RG
                  const double *pu = xub + bo + urow(K + rs, xpar) + K + kx, *p0 = xub + OPB + orow(rs, xpar, a.sh_o0) + kx;
                  double       *p1 = RING + xslot * SLOT + OPB + 2 * OB + orow(rs, xpar, a.sh_o1) + kx;
                  const double  sd = SDI[(xzl % K) * (K * K) + (rs % K) * K + (kx % K)];
#pragma unroll
                  for (int j = 0; j < OY / RG; ++j)
#pragma unroll
                    for (int t = 0; t < 2; ++t)
                      {
                        // rows rs + RG j: same box, RG / 2 box rows further on
                        const int    ou = j * (RG / 2) * BW + t * HX, oo = j * (RG / 2) * OW + t * HX;
                        const double x = pu[ou], xo = has_xo ? p0[oo] : 0.0;
                        p1[oo]         = fma(fma(f1, x - xo, x), sd, sc_inv * p1[oo]);
                      }
                }
            }
          // -------------------------------------------------------------- x-phase: a = Mx u, c = K'x u
          for (int q = tid; q < C::NXT; q += NT)
            {
              const int seg = q / LYS, row = q - seg * LYS;
              const int gy  = gy0 - K + row;
              double   *oa = XA + row * PA + K * seg, *oc = XC + row * PA + K * seg;
              if (xzpl || gy <= 0 || gy >= n1 - 1)
                {
#pragma unroll
                  for (int i = 0; i < K; ++i)
                    oa[i] = 0.0, oc[i] = 0.0;
                }
              else
                {
                  if constexpr (NBC == 1)
                    {
                      // one block: row by row, every output stored as soon as it is complete (few live registers: the
                      // z-sums of this thread stay in registers across the x-phase)
                      const double *ur = xub + urow(row, xpar) + K * seg;
                      double        u[2 * K + 1];
#pragma unroll
                      for (int j = 0; j < 2 * K + 1; ++j)
                        u[j] = ur[j];
                      if (edge_x)
                        {
                          if (seg == 0 && tx == 0)
                            { // x < 0 (outside) and x = 0 (Dirichlet)
#pragma unroll
                              for (int j = 0; j <= K; ++j)
                                u[j] = 0.0;
                            }
                          if (seg == 1 && tx == 0)
                            u[0] = 0.0; // x = 0 seen from the second cell
                          if (seg == TX - 1 && tx == a.ntx - 1)
                            u[2 * K] = 0.0; // x = n1-1 (Dirichlet)
                        }
                      if (CF)
                        {
                          // first Chebyshev iterate x1 = f0 D^-1 b formed on the fly (node class = position in the cell)
                          const double *d1 = SD1 + ((xP % K) * K + (row % K)) * K;
#pragma unroll
                          for (int j = 0; j < 2 * K + 1; ++j)
                            u[j] *= d1[j % K];
                        }
                      {
                        // vertex row: two partial sums each (short dependency chains)
                        double a0 = Mv * u[K], k0 = Kv * u[K], m1 = MC(0, 1) * u[K + 1], k1 = KC(0, 1) * u[K + 1];
#pragma unroll
                        for (int j = 0; j < K; ++j)
                          a0 = fma(MC(K, j), u[j], a0), k0 = fma(KC(K, j), u[j], k0);
#pragma unroll
                        for (int j = 2; j <= K; ++j)
                          m1 = fma(MC(0, j), u[K + j], m1), k1 = fma(KC(0, j), u[K + j], k1);
                        oa[0] = a0 + m1, oc[0] = k0 + k1;
                      }
#pragma unroll
                      for (int i = 1; i < K; ++i)
                        {
                          double ai = MC(i, 0) * u[K], ki = KC(i, 0) * u[K];
#pragma unroll
                          for (int j = 1; j <= K; ++j)
                            ai = fma(MC(i, j), u[K + j], ai), ki = fma(KC(i, j), u[K + j], ki);
                          oa[i] = ai, oc[i] = ki;
                        }
                    }
                  else
                    {
                    double am[K], ak[K], cmix[K];
#pragma unroll
                    for (int i = 0; i < K; ++i)
                      cmix[i] = 0.0, am[i] = 0.0, ak[i] = 0.0;
#pragma unroll
                    for (int jb = 0; jb < NBC; ++jb)
                      {
                        // coupled: block jb's staged plane (its row parity differs by the parity of the block distance)
                        // (km: staged plane 0 = u_b, plane 1 = w_b, both at the rows of block b)
                        const int     pj = (NBC == 1 || a.km) ? xpar : (xpar ^ (((jb - b) & 1) & (int)(a.rows_per_block & 1)));
                        const double *ur = xub + jb * (2 * UB) + urow(row, pj, (a.km && jb == 1) ? a.sh_o0 : a.sh_src) + K * seg;
                        double        u[2 * K + 1];
#pragma unroll
                        for (int j = 0; j < 2 * K + 1; ++j)
                          u[j] = ur[j];
                        if (edge_x)
                          {
                            if (seg == 0 && tx == 0)
                              { // x < 0 (outside) and x = 0 (Dirichlet)
#pragma unroll
                                for (int j = 0; j <= K; ++j)
                                  u[j] = 0.0;
                              }
                            if (seg == 1 && tx == 0)
                              u[0] = 0.0; // x = 0 seen from the second cell
                            if (seg == TX - 1 && tx == a.ntx - 1)
                              u[2 * K] = 0.0; // x = n1-1 (Dirichlet)
                          }
                        if (CF)
                          {
                            // first Chebyshev iterate x1 = f0 D^-1 b formed on the fly (node class = position in the cell)
                            // (the blocks of a coupled pair share the diagonal, checked by v3_apply; block jb has its own f0)
                            const double *d1 = SD1 + ((((xP % K) + K) % K) * K + (row % K)) * K;
                            const double  rj = (jb == b) ? 1.0 : a.cc[V3_F0C + jb] / f0;
#pragma unroll
                            for (int j = 0; j < 2 * K + 1; ++j)
                              u[j] *= d1[j % K] * rj;
                          }
                        // mass sweep of this block; vertex row: two partial sums (short dependency chains)
                        double mj[K];
                        {
                          double m1 = MC(0, 1) * u[K + 1];
                          mj[0]     = Mv * u[K];
#pragma unroll
                          for (int j = 0; j < K; ++j)
                            mj[0] = fma(MC(K, j), u[j], mj[0]);
#pragma unroll
                          for (int j = 2; j <= K; ++j)
                            m1 = fma(MC(0, j), u[K + j], m1);
                          mj[0] += m1;
                        }
#pragma unroll
                        for (int i = 1; i < K; ++i)
                          {
                            mj[i] = MC(i, 0) * u[K];
#pragma unroll
                            for (int j = 1; j <= K; ++j)
                              mj[i] = fma(MC(i, j), u[K + j], mj[i]);
                          }
                        if (NBC > 1)
                          {
                            const double cbj = a.cc[(b * NBC + jb) % 16];
#pragma unroll
                            for (int i = 0; i < K; ++i)
                              cmix[i] = fma(cbj, mj[i], cmix[i]);
                          }
                        if (NBC == 1 || jb == (a.km ? 0 : b))
                          {
                            // stiffness sweep (K' for plain operators) of the block this CTA produces
#pragma unroll
                            for (int i = 0; i < K; ++i)
                              am[i] = mj[i];
                            double k1 = KC(0, 1) * u[K + 1];
                            ak[0]     = Kv * u[K];
#pragma unroll
                            for (int j = 0; j < K; ++j)
                              ak[0] = fma(KC(K, j), u[j], ak[0]);
#pragma unroll
                            for (int j = 2; j <= K; ++j)
                              k1 = fma(KC(0, j), u[K + j], k1);
                            ak[0] += k1;
#pragma unroll
                            for (int i = 1; i < K; ++i)
                              {
                                ak[i] = KC(i, 0) * u[K];
#pragma unroll
                                for (int j = 1; j <= K; ++j)
                                  ak[i] = fma(KC(i, j), u[K + j], ak[i]);
                              }
                          }
                      }
                    if (NBC > 1)
                      {
                        // dst_b = cl_b K u_b + M sum_j C_bj u_j: a = cl Mx u_b, c = cl Kx u_b + sum_j C_bj Mx u_j
#pragma unroll
                        for (int i = 0; i < K; ++i)
                          ak[i] = fma(cl, ak[i], cmix[i]), am[i] *= cl;
                      }
#pragma unroll
                    for (int i = 0; i < K; ++i)
                      oa[i] = am[i], oc[i] = ak[i];
                    }
                }
            }
        };

        // one node plane; ZL = position of the plane in its cell layer (0 only for the first plane of a piece),
        // Lc = the layer the plane belongs to as plane ZL (for ZL == K: the layer it completes)
        auto step = [&](auto zl_c, const int Lc, const int zl_rt = 0) {
          // ZLc < 0: one of the planes 1 .. K-1, position zl_rt at run time (MIDLOOP)
          constexpr int  ZLc = decltype(zl_c)::value;
          constexpr bool MID = (ZLc < 0);
          const int      ZL  = MID ? zl_rt : ZLc;
          const int      par = pb ^ (ZL & 1); // parity of the staged row 0 of this plane
          const bool    zpl = (P <= 0) || (P >= n1 - 1);
          const double *ub  = RING + slot * SLOT;
          const double *SA = AC + (NAC == 2 ? sbuf * (2 * LYS * PA) : 0), *SC = SA + LYS * PA;
          if (C::PIPE)
            {
              // the tile of this plane was filled in the previous interval; fill the other buffer with the next plane
              if (s + 1 < nsteps)
                {
                  const unsigned nslot = (slot + 1 == NBUF) ? 0 : slot + 1;
                  xphase(nslot, (nslot == 0) ? phase ^ 1 : phase, par ^ 1, P + 1, sbuf ^ 1);
                }
            }
          else
            {
              if (NAC == 1)
                {
                  // the a'/c tile of the previous plane is consumed: the warps that write it in the x-phase wait for all
                  // readers, the others only announce that they are done (and start on the epilogue pre-pass)
                  if (NT > (C::NXT + 31) / 32 * 32 && wrp >= (C::NXT + 31) / 32)
                    asm volatile("bar.arrive 2, %0;" ::"r"(NT) : "memory");
                  else
                    asm volatile("bar.sync 2, %0;" ::"r"(NT) : "memory");
                }
              xphase(slot, phase, par, P, sbuf, ZL);
              __syncthreads();
              // the slot of the previous plane is free now (its operands were read in the previous y+z phase)
              if (s >= 1 && s - 1 + NBUF < nsteps)
                issue(s - 1 + NBUF);
            }

          if (is_yz)
            {
              // -------------------------------------------------------------- linear part of the epilogue of this plane
              // the z-sums run on (A x) / sc - g with  g = rhs / sc (residual) | (rhs + ((1 + f1) x - f1 x_old) / (f2 dinv)) / sc
              // (Chebyshev); sc = the scalar factored out of the operator (see v3_apply)
              const bool owned = (NOPS > 0 || CF) && (ZL > 0 || carry_in) && !zpl && (P >= K * L0) && (P < K * L1);
              double     g[NPT];
#pragma unroll
              for (int i = 0; i < NPT; ++i)
                g[i] = 0.0;
              if (CF && owned)
                {
                  // b = src of this plane: g = kappa b (the first iterate x1 = f0 D^-1 b is stored during the x-phase)
#pragma unroll
                  for (int i = 0; i < NPT; ++i)
                    g[i] = kappa * ub[bo + urow(K + K * ys + i0 + i, par) + K + xl];
                }
              if (NOPS > 0 && owned)
                {
                  if (MODE == V2_RESIDUAL)
                    {
#pragma unroll
                      for (int i = 0; i < NPT; ++i)
                        g[i] = sc_inv * ub[OPB + orow(K * ys + i0 + i, par, a.sh_o0) + xl];
                    }
                  else
                    {
                      // formed during the x-phase (see xphase)
#pragma unroll
                      for (int i = 0; i < NPT; ++i)
                        g[i] = ub[OPB + 2 * OB + orow(K * ys + i0 + i, par, a.sh_o1) + xl];
                    }
                }
              // (subtracted BEFORE the y-sweep: g is dead while the sweep runs, and p / w go straight into the z-sums)
              if constexpr (MID)
                {
                  if (NOPS > 0 || CF)
                    {
#pragma unroll
                      for (int z = 1; z < K; ++z)
                        if (ZL == z)
                          {
#pragma unroll
                            for (int i = 0; i < NPT; ++i)
                              acc[z][i] -= g[i];
                          }
                    }
                }
              else if ((NOPS > 0 || CF) && (ZLc > 0 || carry_in))
                {
#pragma unroll
                  for (int i = 0; i < NPT; ++i)
                    acc[ZLc < 0 ? 0 : ZLc][i] -= g[i];
                }
              // -------------------------------------------------------------- y-sweep: p = My a, w = My c + K'y a
              double p[NPT], wv[NPT];
              {
                const double *ar = SA + (K * ys) * PA + xl, *cr = SC + (K * ys) * PA + xl;
#pragma unroll
                for (int h = 0; h < C::NHALF; ++h)
                  if (half == h)
                    {
                      // nodes h NPT .. h NPT + NPT - 1 of the segment; only the vertex node (0) reaches into the cell below
                      constexpr int JMIN_V = 0;
                      const int     jmin = (h == 0) ? JMIN_V : K;
                      double        av[2 * K + 1], cv[2 * K + 1];
#pragma unroll
                      for (int j = 0; j < 2 * K + 1; ++j)
                        if (j >= jmin)
                          av[j] = ar[j * PA], cv[j] = cr[j * PA];
#pragma unroll
                      for (int ii = 0; ii < NPT; ++ii)
                        {
                          const int i = h * NPT + ii;
                          if (i == 0 && NPT == K)
                            {
                              // vertex row: four partial sums for w, two for p (short dependency chains)
                              double p1 = MC(0, 1) * av[K + 1], wm1 = MC(0, 1) * cv[K + 1], wk0 = Kv * av[K], wk1 = KC(0, 1) * av[K + 1];
                              double p0 = Mv * av[K], w0 = Mv * cv[K];
#pragma unroll
                              for (int j = 0; j < K; ++j)
                                {
                                  p0  = fma(MC(K, j), av[j], p0);
                                  w0  = fma(MC(K, j), cv[j], w0);
                                  wk0 = fma(KC(K, j), av[j], wk0);
                                }
#pragma unroll
                              for (int j = 2; j <= K; ++j)
                                {
                                  p1  = fma(MC(0, j), av[K + j], p1);
                                  wm1 = fma(MC(0, j), cv[K + j], wm1);
                                  wk1 = fma(KC(0, j), av[K + j], wk1);
                                }
                              p[ii] = p0 + p1, wv[ii] = (w0 + wm1) + (wk0 + wk1);
                            }
                          else if (i == 0)
                            {
                              // vertex row, few registers (enough warps are resident to hide the longer chains)
                              double p0 = Mv * av[K], w0 = Mv * cv[K], wk0 = Kv * av[K];
#pragma unroll
                              for (int j = 0; j < K; ++j)
                                {
                                  p0  = fma(MC(K, j), av[j], p0);
                                  w0  = fma(MC(K, j), cv[j], w0);
                                  wk0 = fma(KC(K, j), av[j], wk0);
                                }
#pragma unroll
                              for (int j = 1; j <= K; ++j)
                                {
                                  p0  = fma(MC(0, j), av[K + j], p0);
                                  w0  = fma(MC(0, j), cv[K + j], w0);
                                  wk0 = fma(KC(0, j), av[K + j], wk0);
                                }
                              p[ii] = p0, wv[ii] = w0 + wk0;
                            }
                          else
                            {
                              // interior rows: M c and K' a separately
                              double wk = KC(i, 0) * av[K], pi = MC(i, 0) * av[K], wi = MC(i, 0) * cv[K];
#pragma unroll
                              for (int j = 1; j <= K; ++j)
                                {
                                  pi = fma(MC(i, j), av[K + j], pi);
                                  wi = fma(MC(i, j), cv[K + j], wi);
                                  wk = fma(KC(i, j), av[K + j], wk);
                                }
                              p[ii] = pi, wv[ii] = wi + wk;
                            }
                        }
                      // ------------------------------------------------------ z-accumulation: out = Mz w + K'z p
                      // (inside the branch of this half: p / w do not have to be merged across the two branches, which the
                      // register allocator did through the stack)
                      if constexpr (MID)
                        {
                          const double2 *zt = reinterpret_cast<const double2 *>(ZT) + ZL * n;
#pragma unroll
                          for (int z = 0; z < n; ++z)
                            {
                              const double2 mk = zt[z];
#pragma unroll
                              for (int i = 0; i < NPT; ++i)
                                acc[z][i] = fma(mk.x, wv[i], fma(mk.y, p[i], acc[z][i]));
                            }
                        }
                      else if constexpr (ZLc < K)
                        {
#pragma unroll
                          for (int z = 0; z < n; ++z)
#pragma unroll
                            for (int i = 0; i < NPT; ++i)
                              acc[z][i] = fma(MC(z, (ZLc < 0 ? 0 : ZLc)), wv[i], fma(KC(z, (ZLc < 0 ? 0 : ZLc)), p[i], acc[z][i]));
                        }
                    }
              }
              if constexpr (ZLc == K)
                {
#pragma unroll
                  for (int z = 0; z < n; ++z)
#pragma unroll
                    for (int i = 0; i < NPT; ++i)
                      acc[z][i] = fma(MC(z, K), wv[i], fma(KC(z, K), p[i], acc[z][i]));
                  if (Lc >= L0)
                    {
                      const int       gx = gx0 + xl, gy = gy0 + K * ys + i0; // first of the NPT nodes
                      const long long j0 = boff + gx + (long long)n1 * gy + plane * (K * Lc - a.zo0);
                      const bool      anyb = (gx == 0) || (gy == 0) || (Lc == 0);
                      const double   *sds  = SDS + i0 * K + (xl % K);
                      const double    sca  = (MODE == V2_APPLY) ? sc : -sc;
                      // final value of node i of plane z of the layer from its complete z-sum v
                      auto store_node = [&](const int z, const int i, const double v) {
                        const long long j = j0 + z * plane + i * n1;
                        if (anyb && ((gx == 0) || (i == 0 && gy == 0) || (z == 0 && Lc == 0)))
                          v3_identity<MODE>(a, f1, f2, j, f0);
                        else if (MODE == V2_CHEB)
                          {
                            // explicit inverse diagonal: operands straight from global memory
                            const double x = a.src[j], xo = has_xo ? a.x_old[j] : 0.0;
                            a.dst[j]       = fma(f2 * a.dinv[j], a.rhs[j] - sc * v, fma(f1, x - xo, x));
                          }
                        else
                          a.dst[j] = ((MODE == V2_CHEB_OWN || CF) ? sds[(z * K + i) * K] : sca) * v;
                      };
                      // The bottom plane of a range that starts above layer 0 (work-item schedule) also needs the z-sums of the
                      // layer below, the top plane of a range that ends below the top those of the layer above: the two CTAs
                      // meet at `state`; the first one publishes its partial sums, the second one adds them and stores.
                      auto combine = [&](const int Lb, const double(&part)[NPT]) {
                        const int off = Lb - a.L_lo, nbl = a.n_big * a.lock_len; // index of the range that starts at layer Lb
                        const int bnd = (b * ncols + col) * a.lock_nch + (off < nbl ? off / a.lock_len : a.n_big + (off - nbl) / a.len_small);
                        double   *cb  = a.carry + (size_t)bnd * (NY * NPT);
                        if (tid == 0)
                          QS[2] = atomicCAS(a.state + bnd, 0, 1);
                        asm volatile("bar.sync 1, %0;" ::"r"(NY) : "memory");
                        const int role = QS[2];
                        if (role == 0)
                          {
#pragma unroll
                            for (int i = 0; i < NPT; ++i)
                              cb[i * NY + tid] = part[i];
                            __threadfence();
                            asm volatile("bar.sync 1, %0;" ::"r"(NY) : "memory");
                            if (tid == 0)
                              atomicExch(a.state + bnd, 2);
                          }
                        else
                          {
                            if (tid == 0)
                              {
                                while (atomicAdd(a.state + bnd, 0) != 2)
                                  ;
                                __threadfence();
                              }
                            asm volatile("bar.sync 1, %0;" ::"r"(NY) : "memory");
                            // the plane is stored as plane 0 of layer Lb (this CTA sees it as z = K of layer Lb - 1 or as
                            // z = 0 of layer Lb: same node class, same index)
                            const long long jb = boff + gx + (long long)n1 * gy + plane * (K * Lb - a.zo0);
#pragma unroll
                            for (int i = 0; i < NPT; ++i)
                              {
                                const double    v = part[i] + __ldcg(cb + i * NY + tid);
                                const long long j = jb + i * n1;
                                if ((gx == 0) || (i == 0 && gy == 0))
                                  v3_identity<MODE>(a, f1, f2, j, f0);
                                else if (MODE == V2_CHEB)
                                  {
                                    const double x = a.src[j], xo = has_xo ? a.x_old[j] : 0.0;
                                    a.dst[j]       = fma(f2 * a.dinv[j], a.rhs[j] - sc * v, fma(f1, x - xo, x));
                                  }
                                else
                                  a.dst[j] = ((MODE == V2_CHEB_OWN || CF) ? sds[i * K] : sca) * v;
                              }
                            asm volatile("bar.sync 1, %0;" ::"r"(NY) : "memory");
                            if (tid == 0)
                              a.state[bnd] = 0; // ready for the next launch
                          }
                      };
                      const bool defer0 = carry_in && (Lc == L0);
                      if (!anyb && !defer0 && MODE != V2_CHEB)
                        {
                          double *dp = a.dst + j0;
#pragma unroll
                          for (int z = 0; z < K; ++z)
#pragma unroll
                            for (int i = 0; i < NPT; ++i)
                              dp[z * plane + i * n1] = ((MODE == V2_CHEB_OWN || CF) ? sds[(z * K + i) * K] : sca) * acc[z][i];
                        }
                      else
                        {
#pragma unroll
                          for (int z = 0; z < K; ++z)
                            if (!(z == 0 && defer0))
                              {
#pragma unroll
                                for (int i = 0; i < NPT; ++i)
                                  store_node(z, i, acc[z][i]);
                              }
                        }
                      if (defer0)
                        combine(L0, acc[0]);
                      if (carry_out && Lc == L1 - 1)
                        combine(L1, acc[K]);
                    }
                  // the top plane becomes the bottom plane of the next layer
#pragma unroll
                  for (int i = 0; i < NPT; ++i)
                    {
                      acc[0][i] = fma(MC(0, 0), wv[i], fma(KC(0, 0), p[i], acc[K][i]));
#pragma unroll
                      for (int z = 1; z < n; ++z)
                        acc[z][i] = fma(MC(z, 0), wv[i], KC(z, 0) * p[i]);
                    }
                }
            }
          if ((NH > 0 ? !is_yz : true) && ZLc == K && Lc >= L0)
            {
              // helpers (all threads without helper warps): Dirichlet faces x = n1-1 and y = n1-1 of the K planes of the
              // completed layer (owned by no tile)
              const int ht = (NH > 0) ? tid - NY : tid, hs = (NH > 0) ? NH : NT;
              if (tx == a.ntx - 1)
                {
                  const int oye = OY + (ty == a.nty - 1 ? 1 : 0);
                  for (int e = ht; e < K * oye; e += hs)
                    v3_identity<MODE>(a, f1, f2, boff + (n1 - 1) + (long long)n1 * (gy0 + e % oye) + plane * (K * Lc + e / oye - a.zo0), f0);
                }
              if (ty == a.nty - 1)
                for (int e = ht; e < K * OX; e += hs)
                  v3_identity<MODE>(a, f1, f2, boff + (gx0 + e % OX) + (long long)n1 * (n1 - 1) + plane * (K * Lc + e / OX - a.zo0), f0);
            }
          if (C::PIPE)
            {
              __syncthreads(); // the next tile is complete; this plane's ring slot and tile buffer are free
              if (s + NBUF < nsteps)
                issue(s + NBUF);
            }
#ifndef SPIRK_V3_NO_PIN
          // pin the z-sums here: without it the compiler sinks the accumulation below the next barrier and spills p / w
          // across it
          if (is_yz)
            {
#pragma unroll
              for (int z = 0; z < n; ++z)
#pragma unroll
                for (int i = 0; i < NPT; ++i)
                  asm volatile("" : "+d"(acc[z][i]));
            }
#endif
          // advance the running state
          ++s, ++P, sbuf ^= 1;
          if (++slot == NBUF)
            slot = 0, phase ^= 1;
        };

        if (C::PIPE)
          {
            xphase(slot, phase, pb, P, sbuf);
            __syncthreads();
          }
        step(std::integral_constant<int, 0>{}, zf);
        for (int Lc = zf; Lc < L1; ++Lc)
          {
            if constexpr (C::MIDLOOP)
              {
#pragma unroll 1
                for (int zl = 1; zl < K; ++zl)
                  step(std::integral_constant<int, -1>{}, Lc, zl);
                step(std::integral_constant<int, K>{}, Lc);
                continue;
              }
            step(std::integral_constant<int, 1>{}, Lc);
            if constexpr (K >= 2)
              step(std::integral_constant<int, (K >= 2 ? 2 : 1)>{}, Lc);
            if constexpr (K >= 3)
              step(std::integral_constant<int, (K >= 3 ? 3 : 1)>{}, Lc);
            if constexpr (K >= 4)
              step(std::integral_constant<int, (K >= 4 ? 4 : 1)>{}, Lc);
            if constexpr (K >= 5)
              step(std::integral_constant<int, (K >= 5 ? 5 : 1)>{}, Lc);
            if constexpr (K >= 6)
              step(std::integral_constant<int, (K >= 6 ? 6 : 1)>{}, Lc);
          }
        it0 += nsteps;
        // top plane of the domain (Dirichlet)
        if (L1 == nc)
          {
            const int oxe = OX + (tx == a.ntx - 1 ? 1 : 0), oye = OY + (ty == a.nty - 1 ? 1 : 0);
            for (int e = tid; e < oxe * oye; e += NT)
              v3_identity<MODE>(a, f1, f2, boff + (gx0 + e % oxe) + (long long)n1 * (gy0 + e / oxe) + plane * (n1 - 1 - a.zo0), f0);
          }
      }
    if (a.dyn && tid == 0 && atomicAdd(a.sched + 1, 1) == (int)gridDim.x - 1)
      {
        a.sched[0] = 0, a.sched[1] = 0; // the last CTA out: every other CTA has drawn its final (out of range) item
        __threadfence();
      }
#undef MC
#undef KC
  }

  // ---------------------------------------------------------------------------------------------------------
  // Tensor map over a block vector of `n_elems` doubles starting at `ptr` (8-byte aligned): 2-D view
  // [n_elems' / (2 n1)] x [2 n1] ("super-rows" = pairs of node rows) based at the 16-byte aligned address at or
  // below ptr; *shift = elements between that base and ptr (0 or 1, to be added to x coordinates).
  typedef CUresult (*PFN_tmap_encode_tiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                            const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                            CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  inline PFN_tmap_encode_tiled v3_encode_fn()
  {
    static PFN_tmap_encode_tiled fn = nullptr;
    if (!fn)
      {
        void                           *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
          fn = reinterpret_cast<PFN_tmap_encode_tiled>(p);
      }
    return fn;
  }
  inline int v3_make_map(CUtensorMap *map, int *shift, const double *ptr, const long long n_elems, const int n1, const int box_w,
                         const int box_h, const int l2promo)
  {
    // the solver applies the operator to the same few vectors over and over: keep the encoded maps
    struct Entry
    {
      const double *ptr;
      long long     n_elems;
      int           n1, box_w, box_h, l2promo, shift;
      CUtensorMap   map;
    };
    static thread_local std::vector<Entry> cache;
    static thread_local size_t             next = 0;
    for (const Entry &e : cache)
      if (e.ptr == ptr && e.n_elems == n_elems && e.n1 == n1 && e.box_w == box_w && e.box_h == box_h && e.l2promo == l2promo)
        {
          *map = e.map, *shift = e.shift;
          return SPIRK_OK;
        }
    PFN_tmap_encode_tiled enc = v3_encode_fn();
    if (!enc)
      return set_error(SPIRK_ERR_DEVICE, "cuTensorMapEncodeTiled is not available");
    const uintptr_t  p0 = reinterpret_cast<uintptr_t>(ptr);
    *shift              = (int)((p0 >> 3) & 1);
    void            *base = reinterpret_cast<void *>(p0 & ~(uintptr_t)15);
    const cuuint64_t dims[2]    = {(cuuint64_t)2 * n1, (cuuint64_t)((n_elems + *shift) / (2LL * n1))};
    const cuuint64_t strides[1] = {(cuuint64_t)2 * n1 * sizeof(double)};
    const cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)box_h}, estr[2] = {1, 1};
    alignas(64) CUtensorMap tm;
    const CUresult   r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)l2promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
      return set_error(SPIRK_ERR_DEVICE, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    *map = tm;
    Entry e{ptr, n_elems, n1, box_w, box_h, l2promo, *shift, tm};
    if (cache.size() < 64)
      cache.push_back(e);
    else
      cache[next++ % 64] = e;
    return SPIRK_OK;
  }

  // Ranges of a column for the work queue: n_big ranges of len_big layers followed by ranges of len_small layers.  The CTAs
  // draw all long items first; the short ones fill the tail (with equal ranges the last round of a launch leaves most CTA
  // slots idle: 1024 items on 296 slots = 3.46 rounds).  Chosen by simulating the queue (an item costs its layers + a fixed
  // start-up of about 1.25 layers: pipeline fill, the recomputed / exchanged boundary plane).
  inline void v3_plan_ranges(const long long cols, const int nloc, const long long slots, const int len_big, int &n_big, int &len_small)
  {
    struct Key
    {
      long long cols, slots;
      int       nloc, len_big, n_big, len_small;
    };
    static thread_local std::vector<Key> cache;
    for (const Key &k : cache)
      if (k.cols == cols && k.slots == slots && k.nloc == nloc && k.len_big == len_big)
        {
          n_big = k.n_big, len_small = k.len_small;
          return;
        }
    const double c0   = 1.25;
    double       best = 1e300;
    n_big = nloc / len_big, len_small = len_big;
    std::vector<double> heap;
    for (int nb_ = nloc / len_big; nb_ >= 0; --nb_)
      for (int ls = len_big; ls >= 1; ls /= 2)
        {
          const int rem = nloc - nb_ * len_big;
          if (rem <= 0 && ls != len_big)
            continue;
          const int ns = rem > 0 ? (rem + ls - 1) / ls : 0;
          // the queue: every item goes to the slot that becomes free first
          heap.assign((size_t)slots, 0.0);
          auto run = [&](const long long count, const double cost) {
            for (long long i = 0; i < count; ++i)
              {
                std::pop_heap(heap.begin(), heap.end(), std::greater<double>());
                heap.back() += cost;
                std::push_heap(heap.begin(), heap.end(), std::greater<double>());
              }
          };
          run(cols * nb_, len_big + c0);
          if (ns > 0)
            {
              run(cols * (ns - 1), ls + c0);
              run(cols, (rem - (ns - 1) * ls) + c0);
            }
          const double t = *std::max_element(heap.begin(), heap.end());
          if (t < best - 1e-9)
            best = t, n_big = nb_, len_small = ls;
        }
    if (cache.size() < 64)
      cache.push_back(Key{cols, slots, nloc, len_big, n_big, len_small});
  }

  template <int K, int TX, int TY, int MODE, int NPT, int NBC = 1>
  int v3_launch_mode(spirk_ctx *ctx, V3Args &a)
  {
    using C = CfgV3<K, TX, TY, MODE, NPT, NBC>;
    const size_t smem = C::smem + (size_t)std::max(0, ctx->opt_v3_smem_pad_kb) * 1024;
    static size_t attr_set = 0;
    if (attr_set != smem)
      {
        SPIRK_CUDA(cudaFuncSetAttribute(k_v3<K, TX, TY, MODE, NPT, NBC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = smem;
      }
    const int       l2p     = ctx->opt_v3_l2promo;
    // the maps span the ghost planes around the owned range of every block
    const long long below   = (long long)a.g.gh_lo * a.g.plane;
    const long long n_elems = (long long)(a.nb - 1) * a.stride + a.g.N + (long long)(a.g.gh_lo + a.g.gh_hi) * a.g.plane;
    a.L_lo = a.g.L_lo, a.L_hi = a.g.L_hi, a.zo0 = a.g.zo0, a.gh_lo = a.g.gh_lo;
    for (int b = 0; b < a.nb && b < SPIRK_MAX_BLOCKS; ++b) // row index of a plane: b rows_per_block + n1 (P - zo0 + gh_lo) + gy0 - K
      a.pb[b] = (int)(((long long)b * a.rows_per_block + a.gh_lo - a.zo0) & 1);
    if (int e = v3_make_map(&a.tm_src, &a.sh_src, a.src - below, n_elems, a.g.n1, C::BW, C::BH, l2p))
      return e;
    a.tm_o0 = a.tm_src, a.tm_o1 = a.tm_src, a.sh_o0 = a.sh_o1 = 0;
    if (NBC > 1 && a.km) // the second staged plane comes from block b of the vector passed as x_old
      if (int e = v3_make_map(&a.tm_o0, &a.sh_o0, a.x_old - below, n_elems, a.g.n1, C::BW, C::BH, l2p))
        return e;
    const double *o0 = (MODE == V2_RESIDUAL) ? a.rhs : (MODE == V2_CHEB_OWN ? a.x_old : nullptr);
    if (o0 != nullptr)
      if (int e = v3_make_map(&a.tm_o0, &a.sh_o0, o0 - below, n_elems, a.g.n1, C::OW, C::OY / 2, l2p))
        return e;
    if (MODE == V2_CHEB_OWN)
      if (int e = v3_make_map(&a.tm_o1, &a.sh_o1, a.rhs - below, n_elems, a.g.n1, C::OW, C::OY / 2, l2p))
        return e;
    // Schedules (option "v3_schedule"): 2 (default) work queue of (block, layer range, column) items with the partial sums
    // of the shared vertex planes exchanged between neighbouring ranges (no recomputation, any number of items per CTA);
    // 0 even static split of the (block, column, layer) space, 1 z-lockstep (one column x one of nch equal layer ranges per
    // CTA) - in both a piece that starts above layer 0 recomputes the layer below it.
    const long long slots = (long long)ctx->n_sms * C::MINB;
    long long       grid  = std::max(1LL, std::min(ctx->opt_v3_grid > 0 ? (long long)ctx->opt_v3_grid : slots, a.W / 4));
    a.lock_nch = 0, a.lock_len = 0, a.dyn = 0, a.n_items = 0, a.sched = nullptr, a.state = nullptr, a.carry = nullptr;
    a.n_big = 0, a.len_small = 1, a.items_big = 0;
    const long long cols  = (long long)a.nb * a.ntx * a.nty;
    const int       nloc  = a.g.L_hi - a.g.L_lo; // cell layers of this slab
    const int       force = (a.g.col_size > 1) ? 2 : ctx->opt_v3_schedule; // (z-slabs: the work queue only)
    if (force < 0 || force == 2)
      {
        // layers per item: as short as possible while the items do not outnumber the CTA slots, 8 at most (an item costs a
        // pipeline fill and up to two boundary exchanges; measured r = 3 .. 7, profiles/README.md)
        int len = 8;
        if (ctx->opt_v3_chunk > 0)
          len = ctx->opt_v3_chunk;
        else
          for (int c = 1; c < 8; c *= 2)
            if (cols * ((nloc + c - 1) / c) <= slots)
              {
                len = c;
                break;
              }
        len = std::max(1, std::min(len, nloc));
        a.lock_len = len, a.lock_nch = (nloc + len - 1) / len, a.n_big = a.lock_nch, a.len_small = len;
        if (ctx->opt_v3_chunk <= 0 && ctx->opt_v3_tail != 0 && cols * a.lock_nch > slots)
          {
            // more items than CTA slots: long ranges first, short ones for the tail of the launch
            int n_big, len_small;
            v3_plan_ranges(cols, nloc, slots, len, n_big, len_small);
            a.n_big = n_big, a.len_small = len_small;
            a.lock_nch = n_big + (nloc - n_big * len + len_small - 1) / len_small;
          }
        a.items_big = (int)(cols * a.n_big);
        a.dyn      = 1;
        a.n_items  = (int)(cols * a.lock_nch);
        if (int e = ensure_v3_queue(ctx, (size_t)a.n_items, (size_t)a.n_items * C::NY * NPT))
          return e;
        a.sched = ctx->d_v3_sched, a.state = ctx->d_v3_sched + 4, a.carry = ctx->d_v3_carry;
        grid    = std::max(1LL, std::min(ctx->opt_v3_grid > 0 ? (long long)ctx->opt_v3_grid : slots, (long long)a.n_items));
      }
    else if (ctx->opt_v3_grid <= 0)
      {
        const int nch = (int)std::max(1LL, std::min(slots / cols, (long long)a.g.nc / 8)); // ranges of >= 8 layers
        if (force == 1 || (force == 3 && 2 * cols * nch >= slots))
          {
            a.lock_len = (a.g.nc + nch - 1) / nch;
            a.lock_nch = (a.g.nc + a.lock_len - 1) / a.lock_len;
            grid       = cols * a.lock_nch;
          }
      }
    k_v3<K, TX, TY, MODE, NPT, NBC><<<(unsigned int)grid, C::NT, smem, ctx->stream>>>(a);
    SPIRK_LAUNCH_CHECK(ctx);
    return SPIRK_OK;
  }

#ifndef SPIRK_V3_INSTANTIATE
  // the kernels are instantiated in csrc/v3_mode*.cu (one translation unit per mode)
#define SPIRK_V3_EXTERN(MODE)                                                                \
  extern template int v3_launch_mode<4, 8, 8, MODE, 4, 1>(spirk_ctx *, V3Args &);           \
  extern template int v3_launch_mode<4, 8, 8, MODE, 2, 1>(spirk_ctx *, V3Args &);           \
  extern template int v3_launch_mode<4, 4, 4, MODE, 4, 1>(spirk_ctx *, V3Args &);
  SPIRK_V3_EXTERN(V2_APPLY)
  SPIRK_V3_EXTERN(V2_RESIDUAL)
  SPIRK_V3_EXTERN(V2_CHEB)
  SPIRK_V3_EXTERN(V2_CHEB_OWN)
  SPIRK_V3_EXTERN(V2_CHEB_FIRST)
#undef SPIRK_V3_EXTERN
#define SPIRK_V3_EXTERN(MODE)                                                                \
  extern template int v3_launch_mode<4, 8, 8, MODE, 4, 2>(spirk_ctx *, V3Args &);           \
  extern template int v3_launch_mode<4, 4, 4, MODE, 4, 2>(spirk_ctx *, V3Args &);
  SPIRK_V3_EXTERN(V2_APPLY)
  SPIRK_V3_EXTERN(V2_RESIDUAL)
  SPIRK_V3_EXTERN(V2_CHEB)
  SPIRK_V3_EXTERN(V2_CHEB_OWN)
  SPIRK_V3_EXTERN(V2_CHEB_FIRST)
#undef SPIRK_V3_EXTERN
  int v3_upload_constants_mode0(const FeConst *all);
  int v3_upload_constants_mode1(const FeConst *all);
  int v3_upload_constants_mode2(const FeConst *all);
  int v3_upload_constants_mode3(const FeConst *all);
  int v3_upload_constants_mode4(const FeConst *all);
  int v3_upload_constants_mode5(const FeConst *all);
  int v3_upload_constants_mode6(const FeConst *all);

  template <int K, int TX, int TY, int NPT>
  int v3_launch(spirk_ctx *ctx, V3Args &a, const V2Mode mode)
  {
    a.ntx = a.g.nc / TX, a.nty = a.g.nc / TY;
    a.W   = (long long)a.nb * a.ntx * a.nty * a.g.nc;
    if (a.coupled)
      {
        // coupled pair of blocks (IRK q = 2 system matrix, complex pair); the fused epilogues serve the smoother and the
        // residual of the complex level operators (operator.h:616-665 under preconditioner.h:353-373)
        if (mode == V2_APPLY)
          return v3_launch_mode<K, TX, TY, V2_APPLY, K, 2>(ctx, a);
        if (mode == V2_RESIDUAL)
          return v3_launch_mode<K, TX, TY, V2_RESIDUAL, K, 2>(ctx, a);
        if (mode == V2_CHEB_FIRST)
          return v3_launch_mode<K, TX, TY, V2_CHEB_FIRST, K, 2>(ctx, a);
        if (a.dinv != nullptr)
          return v3_launch_mode<K, TX, TY, V2_CHEB, K, 2>(ctx, a);
        return v3_launch_mode<K, TX, TY, V2_CHEB_OWN, K, 2>(ctx, a);
      }
    if (mode == V2_APPLY)
      return v3_launch_mode<K, TX, TY, V2_APPLY, NPT>(ctx, a);
    if (mode == V2_RESIDUAL)
      return v3_launch_mode<K, TX, TY, V2_RESIDUAL, NPT>(ctx, a);
    if (mode == V2_CHEB_FIRST)
      return v3_launch_mode<K, TX, TY, V2_CHEB_FIRST, NPT>(ctx, a);
    if (a.dinv != nullptr)
      return v3_launch_mode<K, TX, TY, V2_CHEB, NPT>(ctx, a);
    return v3_launch_mode<K, TX, TY, V2_CHEB_OWN, NPT>(ctx, a);
  }

  // returns SPIRK_ERR_UNSUPPORTED when the level / operator shape is not covered
  inline int v3_apply(spirk_ctx *ctx, const Geo &g, const spirk_opdesc *op, V2Mode mode, double *dst, const double *src,
                      const double *x_old, const double *rhs, const double *dinv, long long stride, const double *f1,
                      const double *f2, const double *f0 = nullptr, double *dst1 = nullptr, const double *diag_mass = nullptr,
                      const double *diag_laplace = nullptr)
  {
    const bool coupled = (op->kind == SPIRK_OP_COUPLED);
    if (g.dim != 3 || g.k != 4 || g.nc % 4 != 0 || g.nc < 8)
      return SPIRK_ERR_UNSUPPORTED;
    if (coupled && op->nb != 2)
      return SPIRK_ERR_UNSUPPORTED; // coupled blocks: pairs only (the planes of both blocks are staged)
    // fused first two Chebyshev iterates of a coupled pair: the first iterate of BOTH blocks is formed on the fly from one
    // node-class table, so the blocks must share their diagonal (the complex pair does: operator.h:560-575)
    {
      const double m0 = diag_mass ? diag_mass[0] : op->coupling[0], m1 = diag_mass ? diag_mass[1] : op->coupling[3];
      const double l0 = diag_laplace ? diag_laplace[0] : op->laplace[0], l1 = diag_laplace ? diag_laplace[1] : op->laplace[1];
      if (coupled && mode == V2_CHEB_FIRST && (m0 != m1 || l0 != l1))
        return SPIRK_ERR_UNSUPPORTED;
    }
    if (op->nb > 1 && stride % g.n1 != 0)
      return SPIRK_ERR_UNSUPPORTED; // the blocks must continue the row sequence of block 0 (one tensor map)
    V3Args a;
    a.g = g, a.nb = op->nb, a.stride = stride, a.rows_per_block = stride / g.n1, a.coupled = coupled ? 1 : 0, a.km = 0;
    a.dst = dst, a.src = src, a.x_old = x_old, a.rhs = rhs, a.dinv = (mode == V2_CHEB_FIRST) ? dst1 : dinv;
    if (mode == V2_CHEB_FIRST && (f0 == nullptr || dst1 == nullptr || dst1 == dst || dst1 == src))
      return SPIRK_ERR_UNSUPPORTED;
    const double hd = g.h * g.h * g.h, hl = g.h;
    constexpr int K = 4, n = K + 1;
    double        Ms[n * n], Ks[n * n];
    fe_host_sym(K, Ms, Ks);
    for (int b = 0; b < op->nb; ++b)
      {
        a.cm[b] = coupled ? 0.0 : op->mass[b] * hd, a.cl[b] = op->laplace[b] * hl;
        // node-class diagonal of the fused Chebyshev modes: block b's own term (coupled: coupling[b][b] M + laplace[b] K)
        // unless the caller names the coefficients of the diagonal
        a.dm[b] = (diag_mass ? diag_mass[b] : (coupled ? op->coupling[b * op->nb + b] : op->mass[b])) * hd;
        a.dl[b] = (diag_laplace ? diag_laplace[b] : op->laplace[b]) * hl;
        a.f1[b] = f1 ? f1[b] : 0.0, a.f2[b] = f2 ? f2[b] : 0.0;
        if (mode == V2_CHEB_FIRST)
          a.cc[(coupled ? V3_F0C : 0) + b] = f0[b];
        if (mode >= V2_CHEB && dinv == nullptr && a.f2[b] == 0.0)
          return SPIRK_ERR_UNSUPPORTED; // the folded Chebyshev epilogue divides by f2
        if (!coupled && a.cm[b] == 0.0 && a.cl[b] == 0.0)
          return SPIRK_ERR_UNSUPPORTED; // the zero operator has no scalar to factor out
        // coupled: dst_b = cl_b K u_b + M sum_j C_bj u_j with the plain matrices, nothing factored out
        const bool   lap   = coupled || (a.cl[b] != 0.0);
        const double gamma = (lap && !coupled) ? a.cm[b] / (3.0 * a.cl[b]) : 0.0;
        a.sc[b]            = coupled ? 1.0 : (lap ? a.cl[b] : a.cm[b]);
        for (int i = 0; i < n; ++i)
          for (int j = 0; j < n; ++j)
            a.kp[b][v3_cidx<K>(i, j)] = lap ? Ks[i * n + j] + gamma * Ms[i * n + j] : Ms[i * n + j] / 3.0;
        a.kp[b][V3_NKP - 1] = a.kp[b][v3_cidx<K>(K, K)] + a.kp[b][v3_cidx<K>(0, 0)];
        if (coupled)
          for (int j = 0; j < op->nb; ++j)
            a.cc[b * op->nb + j] = op->coupling[b * op->nb + j] * hd;
      }
    // 8 x 8-cell tiles (37/32 halo) on large levels; 4 x 4-cell tiles (21/16 halo, 4 x the columns) keep all SMs busy on
    // the coarser multigrid levels
    // option "v3_small_below"; measured (IRK q=2 step, r=6): 64 -> 48.2 ms, 32 -> 46.7 ms, 16 -> 47.0 ms
    if (g.nc % 8 != 0 || g.nc < ctx->opt_v3_small_below)
      return v3_launch<4, 4, 4, 4>(ctx, a, mode);
    // nodes per y+z thread on the 8 x 8 tile: 4 = 20 warps/SM at 96 registers, 2 = 32 warps/SM at 64 registers (some
    // spills); measured: 4 is better for the plain apply at r = 6, 2 for the fused epilogues
    // measured with the work-queue schedule (r = 6, nb = 2): apply 0.275 ms with 2 nodes per thread, 0.308 ms with 4
    const int npt = (ctx->opt_v3_npt == 2 || ctx->opt_v3_npt == 4) ? ctx->opt_v3_npt : 2;
    if (npt == 2)
      return v3_launch<4, 8, 8, 2>(ctx, a, mode);
    return v3_launch<4, 8, 8, 4>(ctx, a, mode);
  }

  // dst_b = laplace[b] K v_b + mass[b] M w_b for nb blocks of two block vectors with the same stride, in ONE pass over the
  // cells (24 B per DoF: v and w read, dst written) - the stage-parallel system matrix after the A_inv mixing of the
  // source, main.cc:1580-1592 (M commutes with the stage mixing).  Dirichlet rows: dst_b = v_b.
  inline int v3_apply_km(spirk_ctx *ctx, const Geo &g, const int nb, double *dst, const double *v, const double *w, const long long stride,
                         const double *laplace, const double *mass)
  {
    if (g.dim != 3 || g.k != 4 || g.nc % 4 != 0 || g.nc < 8 || nb < 1 || nb > 8 || stride % g.n1 != 0)
      return SPIRK_ERR_UNSUPPORTED;
    V3Args a;
    a.g = g, a.nb = nb, a.stride = stride, a.rows_per_block = stride / g.n1, a.coupled = 1, a.km = 1;
    a.dst = dst, a.src = v, a.x_old = w, a.rhs = nullptr, a.dinv = nullptr;
    const double hd = g.h * g.h * g.h, hl = g.h;
    constexpr int K = 4, n = K + 1;
    double        Ms[n * n], Ks[n * n];
    fe_host_sym(K, Ms, Ks);
    for (int b = 0; b < nb; ++b)
      {
        a.cm[b] = 0.0, a.cl[b] = laplace[b] * hl, a.f1[b] = a.f2[b] = 0.0, a.sc[b] = 1.0, a.dm[b] = a.dl[b] = 0.0;
        for (int i = 0; i < n; ++i)
          for (int j = 0; j < n; ++j)
            a.kp[b][v3_cidx<K>(i, j)] = Ks[i * n + j];
        a.kp[b][V3_NKP - 1] = a.kp[b][v3_cidx<K>(K, K)] + a.kp[b][v3_cidx<K>(0, 0)];
        a.cc[2 * b] = 0.0, a.cc[2 * b + 1] = mass[b] * hd;
      }
    if (g.nc % 8 != 0 || g.nc < ctx->opt_v3_small_below)
      return v3_launch<4, 4, 4, 4>(ctx, a, V2_APPLY);
    return v3_launch<4, 8, 8, 4>(ctx, a, V2_APPLY);
  }
#endif // SPIRK_V3_INSTANTIATE
} // namespace spirk
