// bsg_la.cu -- the two dense consumers of the packed matvecs.
//
//   bsg_randomsvd   bed_randomSVD (R/autoSVD.R:205-219) -> bigstatsr::big_randomSVD -> RSpectra::svds
//                   [unvendored].  RSpectra runs an implicitly restarted Lanczos iteration on the smaller Gram
//                   operator (A A^T or A^T A), calling back into R twice per step.  Here the whole iteration
//                   stays on the device: thick-restart Lanczos (the explicit form of implicit restarting with
//                   exact shifts) with full re-orthogonalisation, ncv = max(2k+1, 20), the same stopping
//                   rule as ARPACK/Spectra (|Ritz residual| <= tol * max(eps^(2/3), |theta|)), operator =
//                   bsg_view_{c,}prodvec_dev.  Only ncv+1 doubles cross PCIe per step.
//   bsg_tcrossprod  bed_tcrossprodSelf (R/bed-tcrossprodSelf.R:21-52): K = sum_blocks X~_b X~_b^T.  Round-1
//                   version: decode a column block to fp64 on the device (read_bed_scaled semantics) and
//                   accumulate with cuBLAS DSYRK (a plain library SYRK; the int8 tensor-core Gram is the next
//                   step).
#include <cublas_v2.h>
#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "bsg_internal.cuh"

namespace bsg {

// ---- small deterministic vector kernels --------------------------------------------------------
// h[j] = <V[:, j], w>, one block per column, fixed-shape tree
__global__ void k_dots(const double *__restrict__ V, int64_t ld, int ncols, const double *__restrict__ w, int N,
                       double *__restrict__ h) {
  __shared__ double sh[32];
  const int j = blockIdx.x;
  if (j >= ncols) return;
  const double *v = V + (int64_t)j * ld;
  double acc = 0;
  for (int i = threadIdx.x; i < N; i += blockDim.x) acc += v[i] * w[i];
#pragma unroll
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int k = 0; k < (int)(blockDim.x >> 5); k++) t += sh[k];
    h[j] = t;
  }
}

// w[i] -= sum_j V[i, j] * h[j]
__global__ void k_axpys(const double *__restrict__ V, int64_t ld, int ncols, const double *__restrict__ h, int N,
                        double *__restrict__ w) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  double acc = 0;
  for (int j = 0; j < ncols; j++) acc += V[(int64_t)j * ld + i] * h[j];
  w[i] -= acc;
}

// out[:, c] = sum_j V[:, j] * S[j, c]   (S is ncv x kk column-major on device)
__global__ void k_combine(const double *__restrict__ V, int64_t ld, int ncv, const double *__restrict__ S, int lds,
                          int kk, int N, double *__restrict__ out, int64_t ldo) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int c = blockIdx.y;
  if (i >= N || c >= kk) return;
  double acc = 0;
  for (int j = 0; j < ncv; j++) acc += V[(int64_t)j * ld + i] * S[(int64_t)c * lds + j];
  out[(int64_t)c * ldo + i] = acc;
}

__global__ void k_scale_copy(const double *__restrict__ src, double alpha, int N, double *__restrict__ dst) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) dst[i] = src[i] * alpha;
}

__global__ void k_init_vec(int N, uint64_t seed, double *__restrict__ v) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  uint64_t x = seed + 0x9E3779B97F4A7C15ull * (uint64_t)(i + 1);
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  x ^= x >> 31;
  v[i] = ((double)(x >> 11) * (1.0 / 9007199254740992.0)) - 0.5;
}

// cyclic Jacobi eigen-decomposition of a small symmetric matrix (column-major n x n); eigenvalues in w,
// eigenvectors in the columns of Z; sorted descending.
static void jacobi_eigh(std::vector<double> A, int n, std::vector<double> &w, std::vector<double> &Z) {
  Z.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; i++) Z[(size_t)i * n + i] = 1.0;
  auto a = [&](int i, int j) -> double & { return A[(size_t)j * n + i]; };
  for (int sweep = 0; sweep < 60; sweep++) {
    double off = 0, diag = 0;
    for (int j = 0; j < n; j++)
      for (int i = 0; i < n; i++) (i == j ? diag : off) += a(i, j) * a(i, j);
    if (off <= 1e-30 * (diag + 1e-300)) break;
    for (int p = 0; p < n - 1; p++)
      for (int q = p + 1; q < n; q++) {
        double apq = a(p, q);
        if (fabs(apq) < 1e-300) continue;
        double theta = (a(q, q) - a(p, p)) / (2.0 * apq);
        double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < n; k++) {
          double akp = a(k, p), akq = a(k, q);
          a(k, p) = c * akp - s * akq;
          a(k, q) = s * akp + c * akq;
        }
        for (int k = 0; k < n; k++) {
          double apk = a(p, k), aqk = a(q, k);
          a(p, k) = c * apk - s * aqk;
          a(q, k) = s * apk + c * aqk;
        }
        for (int k = 0; k < n; k++) {
          double zkp = Z[(size_t)p * n + k], zkq = Z[(size_t)q * n + k];
          Z[(size_t)p * n + k] = c * zkp - s * zkq;
          Z[(size_t)q * n + k] = s * zkp + c * zkq;
        }
      }
  }
  std::vector<int> ord(n);
  for (int i = 0; i < n; i++) ord[i] = i;
  std::sort(ord.begin(), ord.end(), [&](int x, int y) { return a(x, x) > a(y, y); });
  w.resize(n);
  std::vector<double> Z2((size_t)n * n);
  for (int c = 0; c < n; c++) {
    w[c] = a(ord[c], ord[c]);
    memcpy(&Z2[(size_t)c * n], &Z[(size_t)ord[c] * n], n * sizeof(double));
  }
  Z.swap(Z2);
}

struct SvdWork {
  double *V = nullptr, *w = nullptr, *tmp = nullptr, *h = nullptr, *S = nullptr, *Y = nullptr;
  ~SvdWork() {
    void *p[] = {V, w, tmp, h, S, Y};
    for (void *q : p)
      if (q) cudaFree(q);
  }
};

}  // namespace bsg

using namespace bsg;

extern "C" {

// Extended form used by the multi-GPU host: `z_dev` (nr doubles, optional) is the buffer that holds the
// n-vector of partial products; after every local A (A^T x) the library synchronises its stream and calls
// reduce_cb(ctx), which must sum z_dev across ranks (e.g. NCCL all-reduce) and return once the sum is visible.
int bsg_randomsvd_ex(bsg_bed *h, const int *ind_row, int nr, const int *ind_col, int nc, const double *center,
                     const double *scale, int k, double tol, int maxit, double *d, double *u, double *v,
                     double *center_out, double *scale_out, int *niter, int *nops, double *z_dev,
                     bsg_reduce_cb reduce_cb, void *ctx, int ncol_total) {
  if (!h || !d) return fail(BSG_ERR_ARG, "null argument");
  BSG_TRY(bind_device(h));
  if (!ind_row) nr = h->n;
  if (!ind_col) nc = h->m;
  const int mtot = reduce_cb ? ncol_total : nc;  // columns of the whole (sharded) matrix
  if (k < 1 || k > std::min(nr, mtot)) return fail(BSG_ERR_ARG, "k must be in 1..min(n, m).");
  if (tol <= 0) tol = 1e-4;
  if (maxit <= 0) maxit = 1000;
  cudaStream_t s = h->stream;

  // ---- scaling: default bed_scaleBinom (R/binom-scaling.R:133-142), same fp64 formulas on the same integers
  std::vector<double> cen(nc), sca(nc);
  if (center && scale) {
    memcpy(cen.data(), center, (size_t)nc * sizeof(double));
    memcpy(sca.data(), scale, (size_t)nc * sizeof(double));
  } else {
    std::vector<double> sumX(nc), denoX(nc);
    std::vector<int> nona(nc);
    int n_bad = 0;
    BSG_TRY(bsg_colstats(h, ind_row, nr, ind_col, nc, sumX.data(), denoX.data(), nona.data(), &n_bad));
    for (int j = 0; j < nc; j++) {
      double af = sumX[j] / (2.0 * (double)nona[j]);
      cen[j] = 2.0 * af;
      sca[j] = sqrt(2.0 * af * (1.0 - af));
    }
  }
  if (center_out) memcpy(center_out, cen.data(), (size_t)nc * sizeof(double));
  if (scale_out) memcpy(scale_out, sca.data(), (size_t)nc * sizeof(double));

  bsg_view *view = nullptr;
  BSG_TRY(bsg_view_create(h, ind_row, nr, ind_col, nc, cen.data(), sca.data(), &view));
  struct ViewGuard {
    bsg_view *v;
    ~ViewGuard() { bsg_view_destroy(v); }
  } guard{view};

  // operator side: the smaller Gram matrix; a sharded matrix always iterates on the sample side
  const bool row_side = reduce_cb ? true : (nr <= nc);
  const int N = row_side ? nr : nc, Mo = row_side ? nc : nr;
  int ncv = std::max(2 * k + 1, 20);
  ncv = std::min(ncv, std::min(nr, mtot));
  if (ncv <= k) ncv = std::min(k + 1, std::min(nr, mtot));
  const bool full_space = ncv <= k;  // degenerate: k == min(n, m)

  SvdWork W;
  const int64_t ld = N;
  BSG_CUDA(cudaMalloc((void **)&W.V, (size_t)ld * (ncv + 1) * sizeof(double)));
  BSG_CUDA(cudaMalloc((void **)&W.tmp, (size_t)std::max(Mo, 1) * sizeof(double)));
  BSG_CUDA(cudaMalloc((void **)&W.h, (size_t)(ncv + 2) * sizeof(double)));
  BSG_CUDA(cudaMalloc((void **)&W.S, (size_t)ncv * ncv * sizeof(double)));
  BSG_CUDA(cudaMalloc((void **)&W.Y, (size_t)ld * ncv * sizeof(double)));
  double *wv = nullptr;  // work vector of length N; the caller's buffer when results are reduced across ranks
  if (reduce_cb && z_dev) {
    wv = z_dev;
  } else {
    BSG_CUDA(cudaMalloc((void **)&W.w, (size_t)N * sizeof(double)));
    wv = W.w;
  }

  int ops = 0;
  auto apply = [&](const double *x, double *out) -> int {  // out = H x
    if (row_side) {
      BSG_TRY(bsg_view_cprodvec_dev(view, x, W.tmp, s));
      BSG_TRY(bsg_view_prodvec_dev(view, W.tmp, out, s));
    } else {
      BSG_TRY(bsg_view_prodvec_dev(view, x, W.tmp, s));
      BSG_TRY(bsg_view_cprodvec_dev(view, W.tmp, out, s));
    }
    if (reduce_cb) {
      BSG_CUDA(cudaStreamSynchronize(s));
      reduce_cb(ctx);
    }
    ops++;
    return BSG_OK;
  };
  std::vector<double> hh(ncv + 2);
  const int TB = 256;
  auto gblocks = [&](int len) { return (len + TB - 1) / TB; };
  // orthogonalise wv against V[:, 0..cnt) (classical Gram-Schmidt, applied twice), returns coefficients and norm
  auto orth = [&](int cnt, std::vector<double> &coef, double &nrm) -> int {
    coef.assign(cnt, 0.0);
    for (int pass = 0; pass < 2 && cnt > 0; pass++) {
      k_dots<<<cnt, 512, 0, s>>>(W.V, ld, cnt, wv, N, W.h);
      k_axpys<<<gblocks(N), TB, 0, s>>>(W.V, ld, cnt, W.h, N, wv);
      count_launch(2);
      BSG_CUDA(cudaMemcpyAsync(hh.data(), W.h, (size_t)cnt * sizeof(double), cudaMemcpyDeviceToHost, s));
      BSG_CUDA(cudaStreamSynchronize(s));
      for (int j = 0; j < cnt; j++) coef[j] += hh[j];
    }
    k_dots<<<1, 512, 0, s>>>(wv, ld, 1, wv, N, W.h);
    count_launch();
    BSG_CUDA(cudaMemcpyAsync(hh.data(), W.h, sizeof(double), cudaMemcpyDeviceToHost, s));
    BSG_CUDA(cudaStreamSynchronize(s));
    nrm = sqrt(hh[0]);
    return BSG_OK;
  };

  // ---- start vector
  k_init_vec<<<gblocks(N), TB, 0, s>>>(N, 0x5EEDull, wv);
  count_launch();
  {
    std::vector<double> c0;
    double nrm = 0;
    BSG_TRY(orth(0, c0, nrm));
    k_scale_copy<<<gblocks(N), TB, 0, s>>>(wv, 1.0 / nrm, N, W.V);
    count_launch();
  }

  std::vector<double> T((size_t)ncv * ncv, 0.0), theta, Sm;
  int have = 0;      // number of basis vectors whose T column is complete
  int iters = 0, nconv = 0;
  double beta_last = 0;
  for (;;) {
    // ---- extend the Krylov basis to ncv vectors
    for (int j = have; j < ncv; j++) {
      BSG_TRY(apply(W.V + (int64_t)j * ld, wv));
      std::vector<double> coef;
      double nrm = 0;
      BSG_TRY(orth(j + 1, coef, nrm));
      for (int i = 0; i <= j; i++) {
        T[(size_t)j * ncv + i] = coef[i];
        T[(size_t)i * ncv + j] = coef[i];
      }
      beta_last = nrm;
      if (j + 1 < ncv) {
        T[(size_t)(j + 1) * ncv + j] = T[(size_t)j * ncv + j + 1] = nrm;
      }
      // next basis vector (also kept as the residual vector V[:, ncv] after the last step)
      double inv = nrm > 0 ? 1.0 / nrm : 0.0;
      k_scale_copy<<<gblocks(N), TB, 0, s>>>(wv, inv, N, W.V + (int64_t)(j + 1) * ld);
      count_launch();
    }
    have = ncv;
    // ---- Ritz pairs of the projected matrix
    jacobi_eigh(T, ncv, theta, Sm);
    const double eps23 = pow(2.220446049250313e-16, 2.0 / 3.0);
    nconv = 0;
    for (int i = 0; i < k; i++) {
      double res = fabs(beta_last * Sm[(size_t)i * ncv + (ncv - 1)]);
      if (res <= tol * std::max(eps23, fabs(theta[i]))) nconv++;
    }
    iters++;
    if (nconv >= k || iters >= maxit || full_space || ncv >= N) break;
    // ---- thick restart: keep nkeep Ritz vectors + the residual direction
    int nkeep = k + std::min(nconv, (ncv - k) / 2);
    if (nkeep == 1 && ncv > 3) nkeep = ncv / 2;
    nkeep = std::min(nkeep, ncv - 1);
    BSG_CUDA(cudaMemcpyAsync(W.S, Sm.data(), (size_t)ncv * ncv * sizeof(double), cudaMemcpyHostToDevice, s));
    dim3 grid(gblocks(N), nkeep);
    k_combine<<<grid, TB, 0, s>>>(W.V, ld, ncv, W.S, ncv, nkeep, N, W.Y, ld);
    count_launch();
    BSG_CUDA(cudaMemcpyAsync(W.V, W.Y, (size_t)ld * nkeep * sizeof(double), cudaMemcpyDeviceToDevice, s));
    BSG_CUDA(cudaMemcpyAsync(W.V + (int64_t)nkeep * ld, W.V + (int64_t)ncv * ld, (size_t)N * sizeof(double),
                             cudaMemcpyDeviceToDevice, s));
    BSG_CUDA(cudaStreamSynchronize(s));
    std::fill(T.begin(), T.end(), 0.0);
    for (int i = 0; i < nkeep; i++) {
      T[(size_t)i * ncv + i] = theta[i];
      double b = beta_last * Sm[(size_t)i * ncv + (ncv - 1)];
      T[(size_t)nkeep * ncv + i] = b;
      T[(size_t)i * ncv + nkeep] = b;
    }
    have = nkeep;
  }

  // ---- singular triplets
  BSG_CUDA(cudaMemcpyAsync(W.S, Sm.data(), (size_t)ncv * ncv * sizeof(double), cudaMemcpyHostToDevice, s));
  dim3 grid(gblocks(N), k);
  k_combine<<<grid, TB, 0, s>>>(W.V, ld, ncv, W.S, ncv, k, N, W.Y, ld);
  count_launch();
  std::vector<double> side((size_t)N * k), other((size_t)Mo * k);
  BSG_CUDA(cudaMemcpyAsync(side.data(), W.Y, (size_t)N * k * sizeof(double), cudaMemcpyDeviceToHost, s));
  double *d_other = nullptr;
  BSG_CUDA(cudaMalloc((void **)&d_other, (size_t)std::max(Mo, 1) * sizeof(double)));
  for (int c = 0; c < k; c++) {
    d[c] = sqrt(std::max(theta[c], 0.0));
    int rc = row_side ? bsg_view_cprodvec_dev(view, W.Y + (int64_t)c * ld, W.tmp, s)
                      : bsg_view_prodvec_dev(view, W.Y + (int64_t)c * ld, W.tmp, s);
    if (rc) {
      cudaFree(d_other);
      return rc;
    }
    k_scale_copy<<<gblocks(Mo), TB, 0, s>>>(W.tmp, d[c] > 0 ? 1.0 / d[c] : 0.0, Mo, d_other);
    count_launch();
    cudaMemcpyAsync(other.data() + (size_t)c * Mo, d_other, (size_t)Mo * sizeof(double), cudaMemcpyDeviceToHost, s);
    cudaStreamSynchronize(s);
  }
  cudaFree(d_other);
  BSG_CUDA(cudaStreamSynchronize(s));
  // deterministic sign: the entry of largest magnitude of each left vector (row side) is positive
  for (int c = 0; c < k; c++) {
    double *us = row_side ? side.data() + (size_t)c * N : other.data() + (size_t)c * Mo;
    int len = row_side ? N : Mo;
    double best = 0;
    for (int i = 0; i < len; i++)
      if (fabs(us[i]) > fabs(best)) best = us[i];
    if (best < 0) {
      for (int i = 0; i < N; i++) side[(size_t)c * N + i] = -side[(size_t)c * N + i];
      for (int i = 0; i < Mo; i++) other[(size_t)c * Mo + i] = -other[(size_t)c * Mo + i];
    }
  }
  const std::vector<double> &U = row_side ? side : other, &Vv = row_side ? other : side;
  if (u) memcpy(u, U.data(), (size_t)nr * k * sizeof(double));
  if (v) memcpy(v, Vv.data(), (size_t)nc * k * sizeof(double));
  if (niter) *niter = iters;
  if (nops) *nops = ops;
  return BSG_OK;
}

int bsg_randomsvd(bsg_bed *h, const int *ind_row, int nr, const int *ind_col, int nc, const double *center,
                  const double *scale, int k, double tol, int maxit, double *d, double *u, double *v,
                  double *center_out, double *scale_out, int *niter, int *nops) {
  return bsg_randomsvd_ex(h, ind_row, nr, ind_col, nc, center, scale, k, tol, maxit, d, u, v, center_out, scale_out,
                          niter, nops, nullptr, nullptr, nullptr, 0);
}

// ---------------------------------------------------------------------------------------------------
int bsg_tcrossprod(bsg_bed *h, const int *ind_row, int nr, const int *ind_col, int nc, const double *center,
                   const double *scale, double *K) {
  if (!h || !K) return fail(BSG_ERR_ARG, "null argument");
  if (!center || !scale) return fail(BSG_ERR_DIM, "Incompatibility between dimensions.");
  BSG_TRY(bind_device(h));
  if (!ind_row) nr = h->n;
  if (!ind_col) nc = h->m;
  cudaStream_t s = h->stream;
  const int *d_row = nullptr, *d_col = nullptr;
  BSG_TRY(upload_index(h, ind_row, nr, h->n, h->w_idx_row, &d_row));
  std::vector<int> iota;
  if (!ind_col) {  // column blocks are addressed through an explicit list
    iota.resize(nc);
    for (int j = 0; j < nc; j++) iota[j] = j + 1;
    ind_col = iota.data();
  }
  BSG_TRY(upload_index(h, ind_col, nc, h->m, h->w_idx_col, &d_col));
  size_t nn = (size_t)std::max(nc, 1);
  BSG_TRY(h->w_center.ensure(nn * sizeof(double)));
  BSG_TRY(h->w_scale.ensure(nn * sizeof(double)));
  BSG_CUDA(cudaMemcpyAsync(h->w_center.p, center, (size_t)nc * sizeof(double), cudaMemcpyHostToDevice, s));
  BSG_CUDA(cudaMemcpyAsync(h->w_scale.p, scale, (size_t)nc * sizeof(double), cudaMemcpyHostToDevice, s));
  // block of columns sized to ~1 GB of decoded doubles
  int blk = (int)std::max<int64_t>(64, std::min<int64_t>(nc > 0 ? nc : 1, ((int64_t)1 << 27) / std::max(nr, 1)));
  double *dK = nullptr, *dX = nullptr;
  BSG_CUDA(cudaMalloc((void **)&dK, (size_t)std::max(nr, 1) * std::max(nr, 1) * sizeof(double)));
  cudaError_t e = cudaMalloc((void **)&dX, (size_t)std::max(nr, 1) * blk * sizeof(double));
  if (e != cudaSuccess) {
    cudaFree(dK);
    return cuda_fail(e, "GRM block");
  }
  cublasHandle_t cb = nullptr;
  if (cublasCreate(&cb) != CUBLAS_STATUS_SUCCESS) {
    cudaFree(dK);
    cudaFree(dX);
    return fail(BSG_ERR_CUDA, "cublasCreate failed");
  }
  cublasSetStream(cb, s);
  cudaMemsetAsync(dK, 0, (size_t)nr * nr * sizeof(double), s);
  int rc = BSG_OK;
  const double one = 1.0;
  for (int j0 = 0; j0 < nc && !rc; j0 += blk) {
    int b = std::min(blk, nc - j0);
    rc = read_dense_scaled(h, d_row, nr, d_col + j0, b, h->w_center.as<double>() + j0, h->w_scale.as<double>() + j0,
                           dX, s);
    if (!rc && nr > 0 &&
        cublasDsyrk(cb, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, nr, b, &one, dX, nr, &one, dK, nr) != CUBLAS_STATUS_SUCCESS)
      rc = fail(BSG_ERR_CUDA, "cublasDsyrk failed");
  }
  if (!rc) {
    cudaError_t e2 = cudaMemcpyAsync(K, dK, (size_t)nr * nr * sizeof(double), cudaMemcpyDeviceToHost, s);
    if (e2 == cudaSuccess) e2 = cudaStreamSynchronize(s);
    if (e2 != cudaSuccess) rc = cuda_fail(e2, "GRM download");
  }
  cublasDestroy(cb);
  cudaFree(dK);
  cudaFree(dX);
  if (rc) return rc;
  // mirror the lower triangle
  for (int j = 0; j < nr; j++)
    for (int i = j + 1; i < nr; i++) K[(size_t)i * nr + j] = K[(size_t)j * nr + i];
  return BSG_OK;
}

}  // extern "C"
