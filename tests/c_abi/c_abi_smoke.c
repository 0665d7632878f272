/* Pure C99 caller of the C ABI (include/polydeal_b200.h): the host mirror of
 * AgglomerationHandler end to end -- grid, agglomerates, DoF numbering, flattening -- and the
 * behaviour of a compute entry point without a GPU.  Built and run by tests/test_host_mirror.py. */
#include <polydeal_b200.h>
#include <stdio.h>
#include <string.h>

#define CHECK(call)                                                        \
  do                                                                       \
    {                                                                      \
      const int rc_ = (call);                                              \
      if (rc_ != PD_OK)                                                    \
        {                                                                  \
          printf("FAILED %s -> %d: %s\n", #call, rc_, pd_last_error());    \
          return 1;                                                        \
        }                                                                  \
    }                                                                      \
  while (0)

int
main(void)
{
  /* test/polydeal/hp_structure_01.cc: 4x4 cells of [-1,1]^2 in 2x2 blocks, DGQ1 */
  const int32_t n[2] = {4, 4};
  const double  lo[2] = {-1., -1.}, hi[2] = {1., 1.};
  pdh_grid     *grid = NULL;
  pdh_handler  *ah   = NULL;
  CHECK(pdh_grid_create_structured(2, n, lo, hi, /*Morton (refine_global) order*/ 0, &grid));
  CHECK(pdh_handler_create(grid, &ah));
  {
    int b;
    for (b = 0; b < 4; ++b)
      {
        const int32_t cells[4] = {4 * b, 4 * b + 1, 4 * b + 2, 4 * b + 3}; /* one refined parent */
        if (pdh_define_agglomerate(ah, cells, 4) != b)
          {
            printf("FAILED define_agglomerate: %s\n", pd_last_error());
            return 1;
          }
      }
  }
  CHECK(pdh_initialize_fe_values(ah, 2, 2));
  CHECK(pdh_distribute_agglomerated_dofs(ah, PD_FE_DGQ, 1));
  if (pdh_n_polytopes(ah) != 4 || pdh_n_dofs(ah) != 16 || pdh_n_dofs_per_cell(ah) != 4)
    {
      printf("FAILED sizes\n");
      return 1;
    }
  {
    pdh_flatten_params prm;
    pd_mesh_desc       d;
    pd_handle         *h = NULL;
    int                rc;
    memset(&d, 0, sizeof d);
    prm.penalty_constant = -1.;
    prm.h_rule           = 0;
    prm.h_const          = 1.;
    prm.visit_rule       = 0;
    CHECK(pdh_flatten(ah, &prm, &d));
    if (d.dim != 2 || d.fe_degree != 1 || d.fe_kind != PD_FE_DGQ || d.n_polytopes != 4 || d.n_block_rows != 4 ||
        d.brow_ptr[4] != 12 /* every 2x2 block sees itself and two neighbours */)
      {
        printf("FAILED descriptor\n");
        return 1;
      }
    rc = pd_create(&d, &h);
    if (pd_device_count() == 0)
      {
        if (rc != PD_ERR_NO_DEVICE || h != NULL)
          {
            printf("FAILED: expected PD_ERR_NO_DEVICE without a GPU, got %d\n", rc);
            return 1;
          }
        printf("no device: %s\n", pd_last_error());
      }
    else
      {
        CHECK(rc);
        CHECK(pd_assemble(h, PD_ASSEMBLE_ALL, NULL));
        printf("assembled %lld values on the device\n", (long long)pd_nnz(h));
        pd_destroy(h);
      }
  }
  pdh_handler_destroy(ah);
  pdh_grid_destroy(grid);
  printf("C ABI OK\n");
  return 0;
}
