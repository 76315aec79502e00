// Glue between deal.II's DoF numbering and the lexicographic numbering of this library (INTEGRATION.md route A).
//
// The reference distributes FE_Q(k) DoFs with deal.II (main.cc:3374-3412); their numbers depend on the triangulation and
// the partition.  This library numbers the nodes of the refined hypercube lexicographically,
//   index = ix + n1 (iy + n1 iz),  n1 = k 2^r + 1,
// where (ix, iy, iz) counts the FE_Q(k) support points (Gauss-Lobatto nodes inside each cell of side 2^-r) along every
// axis.  A maintainer who keeps deal.II for the mesh and moves only the solve needs the permutation between the two; it
// follows from the support points alone (DoFTools::map_dofs_to_support_points).  The functions are templates over the
// point type (anything with operator[] / operator() returning the coordinate) so that this header has no deal.II include.
//
//   std::map<types::global_dof_index, Point<dim>> sp;                      // or the vector overload
//   DoFTools::map_dofs_to_support_points(MappingQ1<dim>(), dof_handler, sp);
//   const auto perm = spirk_host::lexicographic_permutation<dim>(sp, fe_degree, n_refinements);
//   for (auto i : locally_owned) lexicographic_values[perm[i]] = dealii_vector[i];     // deal.II -> library
//   ... solve on the GPU ...
//   for (auto i : locally_owned) dealii_vector[i] = lexicographic_values[perm[i]];     // library -> deal.II
#pragma once
#include <cmath>
#include <map>
#include <stdexcept>
#include <vector>

namespace spirk_host
{
  // Gauss-Lobatto nodes of FE_Q(k) on [0, 1] (k <= 6), the same closed forms deal.II's QGaussLobatto(k+1) yields
  inline std::vector<double> gauss_lobatto_nodes(const int k)
  {
    std::vector<double> x;
    switch (k)
      {
        case 1: x = {-1, 1}; break;
        case 2: x = {-1, 0, 1}; break;
        case 3: x = {-1, -std::sqrt(1.0 / 5.0), std::sqrt(1.0 / 5.0), 1}; break;
        case 4: x = {-1, -std::sqrt(3.0 / 7.0), 0, std::sqrt(3.0 / 7.0), 1}; break;
        case 5:
          {
            const double a = std::sqrt(1.0 / 3.0 - 2.0 * std::sqrt(7.0) / 21.0), b = std::sqrt(1.0 / 3.0 + 2.0 * std::sqrt(7.0) / 21.0);
            x = {-1, -b, -a, a, b, 1};
            break;
          }
        case 6:
          {
            const double a = std::sqrt(5.0 / 11.0 - 2.0 / 11.0 * std::sqrt(5.0 / 3.0)), b = std::sqrt(5.0 / 11.0 + 2.0 / 11.0 * std::sqrt(5.0 / 3.0));
            x = {-1, -b, -a, 0, a, b, 1};
            break;
          }
        default: throw std::invalid_argument("gauss_lobatto_nodes: degree 1..6");
      }
    for (double &v : x)
      v = 0.5 * (v + 1.0);
    return x;
  }

  // index along one axis of the FE_Q(k) support point with coordinate x in [0, 1] on the mesh with 2^r cells per direction
  inline long long axis_index(const double x, const int k, const int r, const std::vector<double> &nodes)
  {
    const long long nc = 1LL << r;
    const double    s  = x * (double)nc;                  // cell coordinate
    long long       c  = (long long)std::floor(s + 1e-9); // cell that contains x (an upper vertex belongs to the cell below)
    if (c >= nc)
      c = nc - 1;
    const double xi = s - (double)c; // reference coordinate in the cell
    int          best = 0;
    for (int i = 1; i <= k; ++i)
      if (std::fabs(nodes[i] - xi) < std::fabs(nodes[best] - xi))
        best = i;
    if (std::fabs(nodes[best] - xi) > 1e-6)
      throw std::runtime_error("axis_index: the point is not an FE_Q support point of this mesh");
    return c * k + best;
  }

  template <int dim, typename PointType>
  long long lexicographic_index(const PointType &p, const int k, const int r, const std::vector<double> &nodes)
  {
    const long long n1 = (long long)k * (1LL << r) + 1;
    long long       i  = 0;
    for (int d = dim - 1; d >= 0; --d)
      i = i * n1 + axis_index(p[d], k, r, nodes);
    return i;
  }

  // perm[deal.II index] = lexicographic index, from a map (or any range of pairs) index -> support point
  template <int dim, typename SupportPointMap>
  std::map<unsigned long long, long long> lexicographic_permutation(const SupportPointMap &support_points, const int k, const int r)
  {
    const std::vector<double>               nodes = gauss_lobatto_nodes(k);
    std::map<unsigned long long, long long> perm;
    for (const auto &ip : support_points)
      perm[(unsigned long long)ip.first] = lexicographic_index<dim>(ip.second, k, r, nodes);
    return perm;
  }
} // namespace spirk_host
