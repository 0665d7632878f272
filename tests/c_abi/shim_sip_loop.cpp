// A SIP assembly loop in the style of the reference's hand-written ones (examples/poisson.cc:745-905,
// include/poly_utils.h:2038-2132) written against include/polydeal_b200_shim.hpp: polytope_iterators(),
// ah.reinit(polytope), ah.reinit(polytope, f), ah.reinit_interface(...), get_dof_indices, and a host matrix
// filled entry by entry.  It is compared with what pd_assemble computes on the device for the same handler,
// and the LinearOperatorMG-style vmult hook is compared with the host matrix times x.
//
//   shim_sip_loop            on a GPU box: prints "SHIM LOOP OK <max relative difference>"
//   shim_sip_loop --host     without a GPU: exercises the host mirror through the shim and checks that the device
//                            entry points fail loudly (PD_ERR_NO_DEVICE); prints "SHIM HOST OK"
#include <polydeal_b200_shim.hpp>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>

using namespace polydeal_b200;

template <int dim>
static int
run(const bool host_only)
{
  // [0,1]^dim refined 3x / 2x (64 cells), irregular agglomerates: Morton cells in runs of varying length
  const unsigned int        n_refine = dim == 2 ? 3 : 2;
  AgglomerationHandler<dim> ah(0.0, 1.0, n_refine);
  const int                 n_cells = 1 << (n_refine * dim);
  const int                 runs[]  = {5, 3, 8, 1, 6, 4, 7, 2};
  for (int c = 0, k = 0; c < n_cells; ++k)
    {
      std::vector<int32_t> cells;
      for (int i = 0; i < runs[k % 8] && c < n_cells; ++i)
        cells.push_back(c++);
      ah.define_agglomerate(cells);
    }
  const unsigned int degree = 2;
  ah.initialize_fe_values(degree + 1, degree + 1);
  ah.distribute_agglomerated_dofs(PD_FE_DGQ, degree);
  const unsigned int dofs_per_cell = ah.n_dofs_per_cell();
  const unsigned int N             = ah.n_dofs();
  std::vector<int64_t> rowptr;
  std::vector<int32_t> cols;
  ah.create_agglomeration_sparsity_pattern(rowptr, cols);
  if (rowptr.size() != N + 1 || rowptr[N] != (int64_t)cols.size())
    return std::printf("bad sparsity pattern\n"), 1;
  {
    // MappingBox / BoundingBox round trip
    const auto       polytope = ah.polytope_iterators()[3];
    const auto       box      = polytope->get_bounding_box();
    const Point<dim> x        = box.unit_to_real(box.real_to_unit(box.hi));
    for (int d = 0; d < dim; ++d)
      if (std::fabs(x[d] - box.hi[d]) > 1e-15)
        return std::printf("bounding box round trip failed\n"), 1;
  }
  if (host_only)
    {
      try
        {
          ah.reinit(ah.polytope_iterators()[0]);
        }
      catch (const Error &e)
        {
          if (e.code != PD_ERR_NO_DEVICE)
            return std::printf("expected PD_ERR_NO_DEVICE, got %d (%s)\n", e.code, e.what()), 1;
          std::printf("SHIM HOST OK (%u polytopes, %u DoFs; device call refused: %s)\n", ah.n_agglomerates(), N, e.what());
          return 0;
        }
      return std::printf("a device call succeeded in --host mode\n"), 1;
    }

  const double penalty_constant = 10. * (degree + dim) * (degree + 1); // include/poly_utils.h:2018-2019
  std::map<std::pair<unsigned, unsigned>, double> system_matrix;
  auto distribute_local_to_global = [&](const std::vector<double> &M, const std::vector<unsigned int> &rows,
                                        const std::vector<unsigned int> &columns) {
    for (unsigned int i = 0; i < dofs_per_cell; ++i)
      for (unsigned int j = 0; j < dofs_per_cell; ++j)
        system_matrix[{rows[i], columns[j]}] += M[i * dofs_per_cell + j];
  };
  std::vector<unsigned int> local_dof_indices, local_dof_indices_neighbor;
  std::vector<double>       cell_matrix(dofs_per_cell * dofs_per_cell), M11(cell_matrix.size()), M12(cell_matrix.size()),
    M21(cell_matrix.size()), M22(cell_matrix.size());

  for (const auto &polytope : ah.polytope_iterators())
    {
      std::fill(cell_matrix.begin(), cell_matrix.end(), 0.);
      const auto &agglo_values = ah.reinit(polytope);
      for (unsigned int q_index : agglo_values.quadrature_point_indices())
        for (unsigned int i = 0; i < dofs_per_cell; ++i)
          for (unsigned int j = 0; j < dofs_per_cell; ++j)
            cell_matrix[i * dofs_per_cell + j] +=
              agglo_values.shape_grad(i, q_index) * agglo_values.shape_grad(j, q_index) * agglo_values.JxW(q_index);
      polytope->get_dof_indices(local_dof_indices);
      const double penalty = penalty_constant / std::fabs(polytope->diameter());
      for (unsigned int f = 0; f < polytope->n_faces(); ++f)
        {
          if (polytope->at_boundary(f))
            {
              const auto &fe_face = ah.reinit(polytope, f);
              const auto  normals = fe_face.get_normal_vectors();
              for (unsigned int q_index : fe_face.quadrature_point_indices())
                for (unsigned int i = 0; i < dofs_per_cell; ++i)
                  for (unsigned int j = 0; j < dofs_per_cell; ++j)
                    cell_matrix[i * dofs_per_cell + j] +=
                      (-fe_face.shape_value(i, q_index) * (fe_face.shape_grad(j, q_index) * normals[q_index]) -
                       (fe_face.shape_grad(i, q_index) * normals[q_index]) * fe_face.shape_value(j, q_index) +
                       penalty * fe_face.shape_value(i, q_index) * fe_face.shape_value(j, q_index)) *
                      fe_face.JxW(q_index);
            }
          else
            {
              const auto neigh_polytope = polytope->neighbor(f);
              if (!(polytope->id() < neigh_polytope->id())) // each interface once, from the smaller id
                continue;
              const unsigned int nofn     = polytope->neighbor_of_agglomerated_neighbor(f);
              const auto         fe_faces = ah.reinit_interface(polytope, neigh_polytope, f, nofn);
              const auto        &fe0 = fe_faces.first, &fe1 = fe_faces.second;
              const auto         normals = fe0.get_normal_vectors();
              for (unsigned int q = 0; q < fe0.n_quadrature_points; ++q) // two-sided alignment of the points
                for (int d = 0; d < dim; ++d)
                  if (std::fabs(fe0.quadrature_point(q)[d] - fe1.quadrature_point(q)[d]) > 1e-15)
                    return std::printf("interface points not aligned\n"), 1;
              std::fill(M11.begin(), M11.end(), 0.);
              std::fill(M12.begin(), M12.end(), 0.);
              std::fill(M21.begin(), M21.end(), 0.);
              std::fill(M22.begin(), M22.end(), 0.);
              for (unsigned int q = 0; q < fe0.n_quadrature_points; ++q)
                for (unsigned int i = 0; i < dofs_per_cell; ++i)
                  for (unsigned int j = 0; j < dofs_per_cell; ++j)
                    {
                      const double g0i = fe0.shape_grad(i, q) * normals[q], g0j = fe0.shape_grad(j, q) * normals[q];
                      const double g1i = fe1.shape_grad(i, q) * normals[q], g1j = fe1.shape_grad(j, q) * normals[q];
                      const double v0i = fe0.shape_value(i, q), v0j = fe0.shape_value(j, q);
                      const double v1i = fe1.shape_value(i, q), v1j = fe1.shape_value(j, q);
                      const unsigned ij = i * dofs_per_cell + j;
                      M11[ij] += (-0.5 * g0i * v0j - 0.5 * g0j * v0i + penalty * v0i * v0j) * fe0.JxW(q);
                      M12[ij] += (0.5 * g0i * v1j - 0.5 * g1j * v0i - penalty * v0i * v1j) * fe1.JxW(q);
                      M21[ij] += (-0.5 * g1i * v0j + 0.5 * g0j * v1i - penalty * v1i * v0j) * fe1.JxW(q);
                      M22[ij] += (0.5 * g1i * v1j + 0.5 * g1j * v1i + penalty * v1i * v1j) * fe1.JxW(q);
                    }
              neigh_polytope->get_dof_indices(local_dof_indices_neighbor);
              distribute_local_to_global(M11, local_dof_indices, local_dof_indices);
              distribute_local_to_global(M12, local_dof_indices, local_dof_indices_neighbor);
              distribute_local_to_global(M21, local_dof_indices_neighbor, local_dof_indices);
              distribute_local_to_global(M22, local_dof_indices_neighbor, local_dof_indices_neighbor);
            }
        }
      distribute_local_to_global(cell_matrix, local_dof_indices, local_dof_indices);
    }

  // the same matrix from the device kernels
  std::vector<double> values;
  assemble_dg_matrix(values, ah);
  double max_abs = 0., max_diff = 0.;
  for (unsigned int r = 0; r < N; ++r)
    for (int64_t k = rowptr[r]; k < rowptr[r + 1]; ++k)
      {
        const auto   it  = system_matrix.find({r, (unsigned)cols[k]});
        const double ref = it == system_matrix.end() ? 0. : it->second;
        max_abs          = std::fmax(max_abs, std::fabs(ref));
        max_diff         = std::fmax(max_diff, std::fabs(ref - values[k]));
      }
  for (const auto &e : system_matrix) // nothing outside the pattern
    {
      const int32_t *b = &cols[rowptr[e.first.first]], *en = &cols[rowptr[e.first.first + 1]];
      bool           found = false;
      for (const int32_t *p = b; p < en; ++p)
        found = found || (unsigned)*p == e.first.second;
      if (!found && std::fabs(e.second) > 0.)
        return std::printf("entry (%u,%u) outside the sparsity pattern\n", e.first.first, e.first.second), 1;
    }
  // the operator hook a solver would call
  const LinearOperatorMG op = linear_operator_mg(ah.device());
  std::vector<double>    x(N), y, y_ref(N, 0.);
  for (unsigned int i = 0; i < N; ++i)
    x[i] = std::sin(0.37 * i) + 0.01 * (i % 7);
  op.vmult(y, x);
  for (const auto &e : system_matrix)
    y_ref[e.first.first] += e.second * x[e.first.second];
  double y_abs = 0., y_diff = 0.;
  for (unsigned int i = 0; i < N; ++i)
    {
      y_abs  = std::fmax(y_abs, std::fabs(y_ref[i]));
      y_diff = std::fmax(y_diff, std::fabs(y_ref[i] - y[i]));
    }
  // agglomerated_quadrature: sum of the weights = volume of the domain
  double vol = 0.;
  for (const auto &polytope : ah.polytope_iterators())
    for (const double w : ah.agglomerated_quadrature(polytope).weights)
      vol += w;
  std::printf("dim %d: %u polytopes, %u DoFs, matrix rel. diff %.2e, vmult rel. diff %.2e, volume %.15f\n", dim,
              ah.n_agglomerates(), N, max_diff / max_abs, y_diff / y_abs, vol);
  if (max_diff > 1e-12 * max_abs || y_diff > 1e-12 * y_abs || std::fabs(vol - 1.) > 1e-13)
    return 1;
  return 0;
}

int
main(int argc, char **argv)
{
  const bool host_only = argc > 1 && std::strcmp(argv[1], "--host") == 0;
  try
    {
      if (run<2>(host_only) || run<3>(host_only))
        return 1;
    }
  catch (const Error &e)
    {
      std::printf("polydeal_b200::Error %d: %s\n", e.code, e.what());
      return 2;
    }
  std::printf(host_only ? "SHIM HOST OK\n" : "SHIM LOOP OK\n");
  return 0;
}
