// bsg_stats.cu -- column / row statistics and dense decodes behind the C ABI.
//   bed_colstats        src/bed-fun.cpp:9-46
//   bed_col_counts_cpp  src/bed-fun.cpp:51-69     bed_row_counts_cpp  src/bed-fun.cpp:72-98
//   read_bed            src/bed-mat-acc.cpp:8-26  read_bed_scaled     src/bed-mat-acc.cpp:30-49
//   snp_colstats        src/colstats.cpp:8-35
//   readbina2           src/read-plink.cpp:61-80  writebina           src/write-plink.cpp:13-52
// All sums here are sums of small integers, so they are exact and order independent: results are
// bit-identical to the reference's scalar loops.
#include <stdio.h>

#include <algorithm>
#include <vector>

#include "bsg_internal.cuh"

namespace bsg {

__global__ void k_gather_counts(const int32_t *__restrict__ cnt, const int *__restrict__ idx, int len,
                                int32_t *__restrict__ out) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= len) return;
  int src = idx ? idx[j] : j;
  reinterpret_cast<int4 *>(out)[j] = reinterpret_cast<const int4 *>(cnt)[src];
}

// sumX = c1 + 2 c2, xxSum = c1 + 4 c2, c = nr - c3, denoX = xxSum - sumX * sumX / c  (src/bed-fun.cpp:36-38)
__global__ void k_colstats_from_counts(const int32_t *__restrict__ cnt, int len, int nr, double *__restrict__ sumX,
                                       double *__restrict__ denoX, int *__restrict__ nona, int *__restrict__ n_bad) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  int bad = 0;
  if (j < len) {
    int c1 = cnt[4 * j + 1], c2 = cnt[4 * j + 2], c3 = cnt[4 * j + 3];
    double xSum = (double)c1 + 2.0 * (double)c2;
    double xxSum = (double)c1 + 4.0 * (double)c2;
    int c = nr - c3;
    sumX[j] = xSum;
    denoX[j] = xxSum - xSum * xSum / c;
    nona[j] = c;
    bad = (2 * (long long)c < nr);
  }
  unsigned b = __ballot_sync(0xffffffffu, bad);
  if ((threadIdx.x & 31) == 0 && b) atomicAdd(n_bad, __popc(b));
}

// FBM twin, no NA handling (src/colstats.cpp:24-31): x = code256[byte]; NA codes poison the sums like in R
__global__ void k_snp_colstats_from_counts(const int32_t *__restrict__ cnt, int len, int nr, double *__restrict__ sumX,
                                           double *__restrict__ denoX) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= len) return;
  int c1 = cnt[4 * j + 1], c2 = cnt[4 * j + 2], c3 = cnt[4 * j + 3];
  double xSum = (double)c1 + 2.0 * (double)c2;
  double xxSum = (double)c1 + 4.0 * (double)c2;
  if (c3 > 0) xSum = xxSum = nan("");
  sumX[j] = xSum;
  denoX[j] = xxSum - xSum * xSum / nr;
}

static bool host_identity(const int *ind, int len, int limit) {
  if (!ind) return true;
  if (len != limit) return false;
  for (int i = 0; i < len; i++)
    if (ind[i] != i + 1) return false;
  return true;
}

// counts for (ind_row, ind_col) into d_out4 [4 x nc]; uses the counts cached at staging when all rows are taken
int col_counts_dev(bsg_bed *h, const int *ind_row, int nr, const int *ind_col, int nc, int32_t **d_out) {
  cudaStream_t s = h->stream;
  const int *d_row = nullptr, *d_col = nullptr;
  BSG_TRY(upload_index(h, ind_col, nc, h->m, h->w_idx_col, &d_col));
  BSG_TRY(h->w_tmp0.ensure((size_t)(nc > 0 ? nc : 1) * 4 * sizeof(int32_t)));
  int32_t *out = h->w_tmp0.as<int32_t>();
  if (host_identity(ind_row, nr, h->n)) {
    if (nc > 0) {
      k_gather_counts<<<(nc + 255) / 256, 256, 0, s>>>(h->cntA, d_col, nc, out);
      count_launch();
    }
  } else {
    BSG_TRY(upload_index(h, ind_row, nr, h->n, h->w_idx_row, &d_row));
    BSG_TRY(counts_cols(h, d_row, nr, d_col, nc, out, s));
  }
  BSG_CUDA(cudaGetLastError());
  *d_out = out;
  return BSG_OK;
}

}  // namespace bsg

using namespace bsg;

#define FIX_DIMS()                         \
  if (!h) return fail(BSG_ERR_ARG, "null handle"); \
  if (h->fbm_generic && !generic_ok) BSG_PACKED_ONLY(h, "This entry point"); \
  BSG_TRY(bind_device(h));                 \
  if (!ind_row) nr = h->n;                 \
  if (!ind_col) nc = h->m;                 \
  if (nr < 0 || nc < 0) return fail(BSG_ERR_ARG, "negative length");

static const bool generic_ok = false;  // entry points that serve dosage FBMs shadow this with `true`

extern "C" {

int bsg_col_counts(bsg_bed *h, const int *ind_row, int nr, const int *ind_col, int nc, int *out) {
  FIX_DIMS();
  int32_t *d = nullptr;
  BSG_TRY(col_counts_dev(h, ind_row, nr, ind_col, nc, &d));
  BSG_CUDA(cudaMemcpyAsync(out, d, (size_t)nc * 4 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  BSG_CUDA(cudaStreamSynchronize(h->stream));
  return BSG_OK;
}

int bsg_row_counts(bsg_bed *h, const int *ind_row, int nr, const int *ind_col, int nc, int *out) {
  FIX_DIMS();
  cudaStream_t s = h->stream;
  const int *d_row = nullptr, *d_col = nullptr;
  BSG_TRY(upload_index(h, ind_row, nr, h->n, h->w_idx_row, &d_row));
  BSG_TRY(h->w_tmp0.ensure((size_t)(nr > 0 ? nr : 1) * 4 * sizeof(int32_t)));
  int32_t *d = h->w_tmp0.as<int32_t>();
  if (host_identity(ind_col, nc, h->m) && h->cntB) {
    if (nr > 0) {
      k_gather_counts<<<(nr + 255) / 256, 256, 0, s>>>(h->cntB, d_row, nr, d);
      count_launch();
    }
  } else if (nr > 0 && nc > 0) {
    // per-sample counts from the plane sums of the X-side kernel (any column multiset, single-copy handles included)
    BSG_TRY(row_counts_planes(h, ind_row, nr, ind_col, nc, d));
  } else {
    BSG_TRY(upload_index(h, ind_col, nc, h->m, h->w_idx_col, &d_col));
    BSG_TRY(counts_rows(h, d_row, nr, d_col, nc, d, s));
  }
  BSG_CUDA(cudaMemcpyAsync(out, d, (size_t)nr * 4 * sizeof(int), cudaMemcpyDeviceToHost, s));
  BSG_CUDA(cudaStreamSynchronize(s));
  return BSG_OK;
}

int bsg_colstats(bsg_bed *h, const int *ind_row, int nr, const int *ind_col, int nc, double *sumX, double *denoX,
                 int *nb_nona_col, int *n_bad) {
  FIX_DIMS();
  cudaStream_t s = h->stream;
  int32_t *d = nullptr;
  BSG_TRY(col_counts_dev(h, ind_row, nr, ind_col, nc, &d));
  size_t nn = (size_t)(nc > 0 ? nc : 1);
  BSG_TRY(h->w_tmp1.ensure(nn * sizeof(double)));
  BSG_TRY(h->w_tmp2.ensure(nn * sizeof(double)));
  BSG_TRY(h->w_tmp3.ensure(nn * sizeof(int) + 16));
  int *d_nona = h->w_tmp3.as<int>();
  int *d_bad = d_nona + nn;
  BSG_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int), s));
  if (nc > 0) {
    k_colstats_from_counts<<<(nc + 255) / 256, 256, 0, s>>>(d, nc, nr, h->w_tmp1.as<double>(), h->w_tmp2.as<double>(),
                                                          d_nona, d_bad);
    count_launch();
  }
  int bad = 0;
  BSG_CUDA(cudaMemcpyAsync(sumX, h->w_tmp1.p, (size_t)nc * sizeof(double), cudaMemcpyDeviceToHost, s));
  BSG_CUDA(cudaMemcpyAsync(denoX, h->w_tmp2.p, (size_t)nc * sizeof(double), cudaMemcpyDeviceToHost, s));
  BSG_CUDA(cudaMemcpyAsync(nb_nona_col, d_nona, (size_t)nc * sizeof(int), cudaMemcpyDeviceToHost, s));
  BSG_CUDA(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, s));
  BSG_CUDA(cudaStreamSynchronize(s));
  if (n_bad) *n_bad = bad;
  return BSG_OK;
}

int bsg_snp_colstats(bsg_bed *h, const int *ind_row, int nr, const int *ind_col, int nc, double *sumX, double *denoX) {
  const bool generic_ok = true;
  FIX_DIMS();
  cudaStream_t s = h->stream;
  size_t nn = (size_t)(nc > 0 ? nc : 1);
  BSG_TRY(h->w_tmp1.ensure(nn * sizeof(double)));
  BSG_TRY(h->w_tmp2.ensure(nn * sizeof(double)));
  int32_t *d = nullptr;
  if (h->fbm_generic) {  // dosage codes: fp64 sums of code256[byte] (bsg_generic.cu)
    const int *d_row = nullptr, *d_col = nullptr;
    BSG_TRY(upload_index(h, ind_row, nr, h->n, h->w_idx_row, &d_row));
    BSG_TRY(upload_index(h, ind_col, nc, h->m, h->w_idx_col, &d_col));
    BSG_TRY(generic_colstats(h, d_row, nr, d_col, nc, h->w_tmp1.as<double>(), h->w_tmp2.as<double>(), s));
  } else {
    BSG_TRY(col_counts_dev(h, ind_row, nr, ind_col, nc, &d));
  }
  if (nc > 0 && !h->fbm_generic) {
    k_snp_colstats_from_counts<<<(nc + 255) / 256, 256, 0, s>>>(d, nc, nr, h->w_tmp1.as<double>(),
                                                              h->w_tmp2.as<double>());
    count_launch();
  }
  BSG_CUDA(cudaMemcpyAsync(sumX, h->w_tmp1.p, (size_t)nc * sizeof(double), cudaMemcpyDeviceToHost, s));
  BSG_CUDA(cudaMemcpyAsync(denoX, h->w_tmp2.p, (size_t)nc * sizeof(double), cudaMemcpyDeviceToHost, s));
  BSG_CUDA(cudaStreamSynchronize(s));
  return BSG_OK;
}

int bsg_read_bed(bsg_bed *h, const int *ind_row, int nr, const int *ind_col, int nc, int na_val, int *out) {
  FIX_DIMS();
  cudaStream_t s = h->stream;
  const int *d_row = nullptr, *d_col = nullptr;
  BSG_TRY(upload_index(h, ind_row, nr, h->n, h->w_idx_row, &d_row));
  BSG_TRY(upload_index(h, ind_col, nc, h->m, h->w_idx_col, &d_col));
  size_t tot = (size_t)nr * nc;
  BSG_TRY(h->w_out.ensure((tot ? tot : 1) * sizeof(int)));
  BSG_TRY(read_dense(h, d_row, nr, d_col, nc, na_val, h->w_out.as<int>(), s));
  BSG_CUDA(cudaMemcpyAsync(out, h->w_out.p, tot * sizeof(int), cudaMemcpyDeviceToHost, s));
  BSG_CUDA(cudaStreamSynchronize(s));
  return BSG_OK;
}

// readbina2 (src/read-plink.cpp:61-80): the FBM.code256 bytes of X[ind_row, ind_col] (codes 0 / 1 / 2 / 3 = NA),
// nr x nc column-major, i.e. the contents of the .bk file snp_readBed2 fills.
int bsg_readbina2(bsg_bed *h, const int *ind_row, int nr, const int *ind_col, int nc, unsigned char *out) {
  FIX_DIMS();
  if (!out) return fail(BSG_ERR_ARG, "null argument");
  cudaStream_t s = h->stream;
  const int *d_row = nullptr, *d_col = nullptr;
  BSG_TRY(upload_index(h, ind_row, nr, h->n, h->w_idx_row, &d_row));
  BSG_TRY(upload_index(h, ind_col, nc, h->m, h->w_idx_col, &d_col));
  // column blocks of <= 1 GB through one device buffer
  const int blk = (int)std::max<int64_t>(1, std::min<int64_t>(nc > 0 ? nc : 1, ((int64_t)1 << 30) / std::max(nr, 1)));
  BSG_TRY(h->w_out.ensure((size_t)std::max(nr, 1) * blk));
  for (int j0 = 0; j0 < nc; j0 += blk) {
    const int b = std::min(blk, nc - j0);
    std::vector<int> iota;
    const int *dc = d_col ? d_col + j0 : nullptr;
    if (!d_col && j0 > 0) {  // identity columns beyond the first block: explicit list
      iota.resize(b);
      for (int j = 0; j < b; j++) iota[j] = j0 + j + 1;
      BSG_TRY(upload_index(h, iota.data(), b, h->m, h->w_idx_col, &dc));
    }
    BSG_TRY(read_bytes(h, d_row, nr, dc, b, h->w_out.as<uint8_t>(), s));
    BSG_CUDA(cudaMemcpyAsync(out + (size_t)j0 * nr, h->w_out.p, (size_t)nr * b, cudaMemcpyDeviceToHost, s));
    BSG_CUDA(cudaStreamSynchronize(s));
  }
  return BSG_OK;
}

// writebina (src/write-plink.cpp:13-52): write X[ind_row, ind_col] of a handle (bed- or FBM-staged) as a .bed file:
// magic 6C 1B 01, then ceil(nr / 4) bytes per selected column, byte for byte what the reference writes.
int bsg_writebina(bsg_bed *h, const char *path, const int *ind_row, int nr, const int *ind_col, int nc) {
  FIX_DIMS();
  if (!path) return fail(BSG_ERR_ARG, "null argument");
  cudaStream_t s = h->stream;
  const int *d_row = nullptr, *d_col = nullptr;
  BSG_TRY(upload_index(h, ind_row, nr, h->n, h->w_idx_row, &d_row));
  BSG_TRY(upload_index(h, ind_col, nc, h->m, h->w_idx_col, &d_col));
  const int nbytes = (nr + 3) / 4;
  FILE *f = fopen(path, "wb");
  if (!f) return fail(BSG_ERR_IO, "cannot open '%s' for writing.", path);
  const unsigned char magic[3] = {108, 27, 1};
  bool ok = fwrite(magic, 1, 3, f) == 3;
  const int blk = (int)std::max<int64_t>(1, std::min<int64_t>(nc > 0 ? nc : 1, ((int64_t)1 << 28) / std::max(nbytes, 1)));
  std::vector<uint8_t> host((size_t)std::max(nbytes, 1) * blk);
  int rc = h->w_out.ensure(host.size());
  for (int j0 = 0; j0 < nc && ok && !rc; j0 += blk) {
    const int b = std::min(blk, nc - j0);
    std::vector<int> iota;
    const int *dc = d_col ? d_col + j0 : nullptr;
    if (!d_col && j0 > 0) {
      iota.resize(b);
      for (int j = 0; j < b; j++) iota[j] = j0 + j + 1;
      rc = upload_index(h, iota.data(), b, h->m, h->w_idx_col, &dc);
      if (rc) break;
    }
    rc = pack_bed(h, d_row, nr, dc, b, h->w_out.as<uint8_t>(), s);
    if (rc) break;
    cudaError_t e = cudaMemcpyAsync(host.data(), h->w_out.p, (size_t)nbytes * b, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) {
      rc = cuda_fail(e, "writebina download");
      break;
    }
    ok = fwrite(host.data(), 1, (size_t)nbytes * b, f) == (size_t)nbytes * b;
  }
  if (fclose(f) != 0) ok = false;
  if (rc) return rc;
  if (!ok) return fail(BSG_ERR_IO, "short write to '%s'.", path);
  return BSG_OK;
}

int bsg_read_bed_scaled(bsg_bed *h, const int *ind_row, int nr, const int *ind_col, int nc, const double *center,
                        const double *scale, double *out) {
  FIX_DIMS();
  if (!center || !scale) return fail(BSG_ERR_DIM, "Incompatibility between dimensions.");
  cudaStream_t s = h->stream;
  const int *d_row = nullptr, *d_col = nullptr;
  BSG_TRY(upload_index(h, ind_row, nr, h->n, h->w_idx_row, &d_row));
  BSG_TRY(upload_index(h, ind_col, nc, h->m, h->w_idx_col, &d_col));
  size_t nn = (size_t)(nc > 0 ? nc : 1);
  BSG_TRY(h->w_center.ensure(nn * sizeof(double)));
  BSG_TRY(h->w_scale.ensure(nn * sizeof(double)));
  BSG_CUDA(cudaMemcpyAsync(h->w_center.p, center, (size_t)nc * sizeof(double), cudaMemcpyHostToDevice, s));
  BSG_CUDA(cudaMemcpyAsync(h->w_scale.p, scale, (size_t)nc * sizeof(double), cudaMemcpyHostToDevice, s));
  size_t tot = (size_t)nr * nc;
  BSG_TRY(h->w_out.ensure((tot ? tot : 1) * sizeof(double)));
  BSG_TRY(read_dense_scaled(h, d_row, nr, d_col, nc, h->w_center.as<double>(), h->w_scale.as<double>(),
                            h->w_out.as<double>(), s));
  BSG_CUDA(cudaMemcpyAsync(out, h->w_out.p, tot * sizeof(double), cudaMemcpyDeviceToHost, s));
  BSG_CUDA(cudaStreamSynchronize(s));
  return BSG_OK;
}

}  // extern "C"
