/* Drives libpq_b200.so exactly as a foreign host (Julia's ccall, src/backends/interactive.jl:32-75)
 * would, from plain C99: create -> save_tensor(s) -> contract -> permute -> save_output ->
 * load_tensor, and checks the numbers against a scalar triple loop.
 *
 *   exit 0  all checks passed          exit 2  no usable GPU (pq_create failed)
 *   exit 1  a check failed
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pq_b200.h"

#define CHECK(call)                                                                   \
  do {                                                                                \
    int rc_ = (call);                                                                 \
    if (rc_ != PQ_OK) {                                                               \
      fprintf(stderr, "%s -> %d: %s\n", #call, rc_, pq_last_error(h));                \
      return 1;                                                                       \
    }                                                                                 \
  } while (0)

int main(void) {
  pq_handle* h = NULL;
  int rc = pq_create(0, PQ_C128, &h);
  if (rc != PQ_OK) {
    fprintf(stderr, "pq_create -> %d (no usable GPU?)\n", rc);
    return 2;
  }
  printf("%s\n", pq_version());

  /* A[i,k] (8 x 6), B[k,j,l] (6 x 5 x 2), column-major complex doubles */
  enum { I = 8, K = 6, J = 5, L = 2 };
  double A[2 * I * K], B[2 * K * J * L], C[2 * I * J * L], P[2 * I * J * L];
  for (int x = 0; x < 2 * I * K; ++x) A[x] = sin(0.37 * x) + 0.01 * x;
  for (int x = 0; x < 2 * K * J * L; ++x) B[x] = cos(0.11 * x) - 0.02 * x;
  const int64_t da[2] = {I, K}, db[3] = {K, J, L};

  /* one tensor through pq_save_tensor, one through the batched call */
  CHECK(pq_save_tensor(h, "node_1", 2, da, A, PQ_HOST_C128));
  {
    const char* labels[1] = {"node_2"};
    const int ranks[1] = {3};
    const void* hosts[1] = {B};
    const int codes[1] = {PQ_HOST_C128};
    CHECK(pq_save_tensors(h, 1, labels, ranks, db, hosts, codes));
  }
  /* contract_tensors(b, :node_1, [-1, 1], :node_2, [1, -2, -3], :node_3): consumes both */
  const int32_t ai[2] = {-1, 1}, bi[3] = {1, -2, -3};
  CHECK(pq_contract(h, "node_1", ai, 2, "node_2", bi, 3, "node_3"));
  int rank = -1;
  int64_t dims[PQ_MAX_RANK];
  if (pq_tensor_info(h, "node_1", &rank, dims) != PQ_ERR_NOT_FOUND) {
    fprintf(stderr, "node_1 should have been consumed\n");
    return 1;
  }
  CHECK(pq_tensor_info(h, "node_3", &rank, dims));
  if (rank != 3 || dims[0] != I || dims[1] != J || dims[2] != L) {
    fprintf(stderr, "unexpected result shape\n");
    return 1;
  }
  CHECK(pq_save_output(h, "node_3", "result"));
  CHECK(pq_load_tensor(h, "result", C, PQ_HOST_C128));
  double worst = 0, norm = 0;
  for (int l = 0; l < L; ++l)
    for (int j = 0; j < J; ++j)
      for (int i = 0; i < I; ++i) {
        double re = 0, im = 0;
        for (int k = 0; k < K; ++k) {
          const double ar = A[2 * (i + I * k)], aim = A[2 * (i + I * k) + 1];
          const double br = B[2 * (k + K * (j + J * l))], bim = B[2 * (k + K * (j + J * l)) + 1];
          re += ar * br - aim * bim;
          im += ar * bim + aim * br;
        }
        const int o = 2 * (i + I * (j + J * l));
        worst = fmax(worst, fmax(fabs(C[o] - re), fabs(C[o + 1] - im)));
        norm = fmax(norm, fmax(fabs(re), fabs(im)));
        P[2 * (l + L * (j + J * i))] = re;   /* expected after permute [3, 2, 1] */
        P[2 * (l + L * (j + J * i)) + 1] = im;
      }
  if (!(worst <= 1e-12 * norm)) {
    fprintf(stderr, "contract: max abs error %.3e (scale %.3e)\n", worst, norm);
    return 1;
  }
  /* permute_tensor(b, :node_3, [3, 2, 1]); the alias :result keeps the old array */
  const int32_t axes[3] = {3, 2, 1};
  CHECK(pq_permute(h, "node_3", axes, 3));
  CHECK(pq_load_tensor(h, "node_3", C, PQ_HOST_C128));
  worst = 0;
  for (int x = 0; x < 2 * I * J * L; ++x) worst = fmax(worst, fabs(C[x] - P[x]));
  if (!(worst <= 1e-12 * norm)) {
    fprintf(stderr, "permute: max abs error %.3e\n", worst);
    return 1;
  }
  CHECK(pq_delete(h, "node_3"));
  CHECK(pq_delete(h, "missing_label_is_fine"));
  CHECK(pq_sync(h));
  pq_destroy(h);
  printf("abi_smoke ok\n");
  return 0;
}
