// bsg_la.cu -- Gram product and truncated SVD over the packed genotypes (filled in below).
#include "bsg_internal.cuh"

using namespace bsg;

extern "C" {

int bsg_tcrossprod(bsg_bed *h, const int *ind_row, int nr, const int *ind_col, int nc, const double *center,
                   const double *scale, double *K) {
  return fail(BSG_ERR_ARG, "bsg_tcrossprod: not implemented yet");
}

int bsg_randomsvd(bsg_bed *h, const int *ind_row, int nr, const int *ind_col, int nc, const double *center,
                  const double *scale, int k, double tol, int maxit, double *d, double *u, double *v,
                  double *center_out, double *scale_out, int *niter, int *nops) {
  return fail(BSG_ERR_ARG, "bsg_randomsvd: not implemented yet");
}

}  // extern "C"
