// CPU check of dealii_spirk_b200/host/dealii_glue.h with a mock of deal.II's support-point map: a scrambled numbering of the
// FE_Q(k) nodes of the refined hypercube must be mapped back to the lexicographic index, for every degree.
#include <algorithm>
#include <array>
#include <cstdio>
#include <random>

#include "../dealii_spirk_b200/host/dealii_glue.h"

int main()
{
  for (int dim = 2; dim <= 3; ++dim)
    for (int k = 1; k <= 6; ++k)
      {
        const int  r     = (dim == 3) ? 2 : 3;
        const auto nodes = spirk_host::gauss_lobatto_nodes(k);
        const long long nc = 1LL << r, n1 = k * nc + 1, N = (dim == 3) ? n1 * n1 * n1 : n1 * n1;
        std::vector<long long> order(N);
        for (long long i = 0; i < N; ++i)
          order[i] = i;
        std::mt19937 gen(k + 10 * dim);
        std::shuffle(order.begin(), order.end(), gen);
        std::map<unsigned long long, std::array<double, 3>> sp; // "deal.II index" -> support point
        for (long long j = 0; j < N; ++j)
          {
            const long long lex = order[j];
            long long       a[3] = {lex % n1, (lex / n1) % n1, lex / (n1 * n1)};
            std::array<double, 3> p{};
            for (int d = 0; d < dim; ++d)
              {
                const long long c = std::min(a[d] / k, nc - 1), l = a[d] - c * k;
                p[d]              = ((double)c + nodes[l]) / (double)nc;
              }
            sp[j] = p;
          }
        const auto perm = (dim == 3) ? spirk_host::lexicographic_permutation<3>(sp, k, r) : spirk_host::lexicographic_permutation<2>(sp, k, r);
        for (long long j = 0; j < N; ++j)
          if (perm.at(j) != order[j])
            {
              std::printf("FAIL dim=%d k=%d j=%lld: %lld != %lld\n", dim, k, j, perm.at(j), order[j]);
              return 1;
            }
      }
  std::printf("glue ok\n");
  return 0;
}
