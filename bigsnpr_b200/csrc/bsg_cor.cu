// bsg_cor.cu -- windowed pairwise-complete correlations: corMat / ld_scores
// (src/corr.cpp:11-97,102-126 ; src/ld-scores.cpp:11-78,83-105).
//
// Every sum the reference accumulates per pair (nona, xSum, xxSum, ySum, yySum, xySum) is a sum of
// small integers, i.e. a handful of population counts over bit planes of the two packed columns:
//     valid_x = ~(lo&hi), x1 = lo&~hi, x2 = hi&~lo   (staged code: 1 -> 01, 2 -> 10, NA -> 11)
//     nona = |vx & vy|, xSum = |x1&vy| + 2|x2&vy|, xxSum = |x1&vy| + 4|x2&vy|  (same for y),
//     xySum = |x1&y1| + 2|x1&y2| + 2|x2&y1| + 4|x2&y2|.
// They are exact, so the fp64 epilogue below -- written in the reference's operation order
// (src/corr.cpp:77-80) -- returns bit-identical r.  Rows / columns subsets (any multiset) are first
// compacted into a temporary packed matrix so the pair kernel always runs on dense lines.
#include <math.h>
#include <stdlib.h>

#include <vector>

#include "bsg_internal.cuh"

namespace bsg {

// out line j = codes of (rows[i], cols[j]) for i < nr, packed 16 per word; pads are code 0.
__global__ void k_compact(const uint8_t *__restrict__ A, int64_t strideA, const int *__restrict__ rows, int nr,
                          const int *__restrict__ cols, int nc, uint8_t *__restrict__ out, int64_t stride_out) {
  int64_t words = stride_out / 4;
  int64_t total = (int64_t)nc * words;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t j = t / words, wq = t - j * words;
    const uint8_t *line = A + (int64_t)(cols ? cols[j] : (int)j) * strideA;
    uint32_t v = 0;
#pragma unroll 4
    for (int p = 0; p < 16; p++) {
      int64_t i = wq * 16 + p;
      if (i < nr) {
        int r = rows ? rows[i] : (int)i;
        v |= (uint32_t)((line[r >> 2] >> (2 * (r & 3))) & 3) << (2 * p);
      }
    }
    reinterpret_cast<uint32_t *>(out + j * stride_out)[wq] = v;
  }
}

struct PairSums {
  int nona, x1v, x2v, y1v, y2v, c11, c12, c21, c22;
};

__device__ __forceinline__ void pair_accum(uint32_t a, uint32_t b, PairSums &s) {
  const uint32_t M = 0x55555555u;
  uint32_t alo = a & M, ahi = (a >> 1) & M, blo = b & M, bhi = (b >> 1) & M;
  uint32_t av = M & ~(alo & ahi), bv = M & ~(blo & bhi);
  uint32_t a1 = alo & ~ahi, a2 = ahi & ~alo, b1 = blo & ~bhi, b2 = bhi & ~blo;
  s.nona += __popc(av & bv);
  s.x1v += __popc(a1 & bv);
  s.x2v += __popc(a2 & bv);
  s.y1v += __popc(b1 & av);
  s.y2v += __popc(b2 & av);
  s.c11 += __popc(a1 & b1);
  s.c12 += __popc(a1 & b2);
  s.c21 += __popc(a2 & b1);
  s.c22 += __popc(a2 & b2);
}

// one block per column j0; warps take neighbours j = j0-1-k (k < wlen[j0]) round robin; lanes stride
// over the words of the two lines.  band[boff[j0] + k] = r (or r^2 for LD), keep[...] = threshold test.
template <bool LD>
__global__ void __launch_bounds__(256) k_cor_pairs(const uint8_t *__restrict__ M, int64_t stride, int nrow, int ncol,
                                                   const int *__restrict__ wlen, const long long *__restrict__ boff,
                                                   const double *__restrict__ thr, double *__restrict__ band,
                                                   uint8_t *__restrict__ keep) {
  const int j0 = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  const int nw = wlen[j0];
  const int nvec = (int)((((int64_t)nrow + 3) / 4 + 15) / 16);
  const uint4 *la = reinterpret_cast<const uint4 *>(M + (int64_t)j0 * stride);
  for (int k = warp; k < nw; k += nwarp) {
    const int j = j0 - 1 - k;
    const uint4 *lb = reinterpret_cast<const uint4 *>(M + (int64_t)j * stride);
    PairSums s = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int v = lane; v < nvec; v += 32) {
      uint4 a = __ldg(la + v), b = __ldg(lb + v);
      pair_accum(a.x, b.x, s);
      pair_accum(a.y, b.y, s);
      pair_accum(a.z, b.z, s);
      pair_accum(a.w, b.w, s);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      s.nona += __shfl_xor_sync(0xffffffffu, s.nona, o);
      s.x1v += __shfl_xor_sync(0xffffffffu, s.x1v, o);
      s.x2v += __shfl_xor_sync(0xffffffffu, s.x2v, o);
      s.y1v += __shfl_xor_sync(0xffffffffu, s.y1v, o);
      s.y2v += __shfl_xor_sync(0xffffffffu, s.y2v, o);
      s.c11 += __shfl_xor_sync(0xffffffffu, s.c11, o);
      s.c12 += __shfl_xor_sync(0xffffffffu, s.c12, o);
      s.c21 += __shfl_xor_sync(0xffffffffu, s.c21, o);
      s.c22 += __shfl_xor_sync(0xffffffffu, s.c22, o);
    }
    if (lane == 0) {
      // pads (code 0) count as valid zeros on both sides: remove them from nona
      const int npad = nvec * 64 - nrow;
      const int nona = s.nona - npad;
      const double xSum = (double)s.x1v + 2.0 * (double)s.x2v;
      const double xxSum = (double)s.x1v + 4.0 * (double)s.x2v;
      const double ySum = (double)s.y1v + 2.0 * (double)s.y2v;
      const double yySum = (double)s.y1v + 4.0 * (double)s.y2v;
      const double xySum = (double)s.c11 + 2.0 * (double)s.c12 + 2.0 * (double)s.c21 + 4.0 * (double)s.c22;
      // src/corr.cpp:77-80 / src/ld-scores.cpp:63-66, same operation order
      const double num = xySum - xSum * ySum / nona;
      const double deno_x = xxSum - xSum * xSum / nona;
      const double deno_y = yySum - ySum * ySum / nona;
      const long long o = boff[j0] + k;
      if (LD) {
        band[o] = num * num / (deno_x * deno_y);
      } else {
        double r = num / sqrt(deno_x * deno_y);
        bool kp = isnan(r) || fabs(r) > thr[nona > 0 ? nona - 1 : 0];
        if (r > 1) r = 1; else if (r < -1) r = -1;
        band[o] = r;
        keep[o] = kp;
      }
    }
  }
}

// res[j] = 1 + sum_k band[j][k] + sum_{j0 > j, j in window(j0)} band[j0][j0-1-j]   (NaN skipped)
__global__ void k_ld_reduce(const double *__restrict__ band, const long long *__restrict__ boff,
                            const int *__restrict__ wlen, const int *__restrict__ reach, int ncol,
                            double *__restrict__ res) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= ncol) return;
  double acc = 1.0;
  for (int k = 0; k < wlen[j]; k++) {
    double v = band[boff[j] + k];
    if (!isnan(v)) acc += v;
  }
  for (int j0 = j + 1; j0 <= reach[j]; j0++) {
    int k = j0 - 1 - j;
    if (k < wlen[j0]) {
      double v = band[boff[j0] + k];
      if (!isnan(v)) acc += v;
    }
  }
  res[j] = acc;
}

struct Window {
  std::vector<int> wlen, reach;
  std::vector<long long> boff;
  long long total = 0;
};

// window of j0: j = j0-1 downto 0 while pos[j] >= pos[j0] - size   (src/corr.cpp:52-53), literal scan
static void build_window(const double *pos, int nc, double size, Window &w) {
  w.wlen.assign(nc, 0);
  w.reach.assign(nc, 0);
  w.boff.assign(nc + 1, 0);
  for (int j = 0; j < nc; j++) w.reach[j] = j;
  for (int j0 = 0; j0 < nc; j0++) {
    double pos_min = pos[j0] - size;
    int j = j0 - 1, c = 0;
    while (j >= 0 && pos[j] >= pos_min) {
      c++;
      j--;
    }
    w.wlen[j0] = c;
    if (c > 0 && w.reach[j0 - c] < j0) w.reach[j0 - c] = j0;
  }
  // reach[j] = largest j0 whose window contains j: windows are contiguous, take a running max from the left
  for (int j = 1; j < nc; j++)
    if (w.reach[j - 1] > w.reach[j] && w.reach[j - 1] > j) w.reach[j] = std::max(w.reach[j], w.reach[j - 1]);
  long long t = 0;
  for (int j = 0; j < nc; j++) {
    w.boff[j] = t;
    t += w.wlen[j];
  }
  w.boff[nc] = t;
  w.total = t;
}

template <class T>
static int to_dev(T **dst, const std::vector<T> &v, cudaStream_t s) {
  BSG_CUDA(cudaMalloc((void **)dst, (v.size() ? v.size() : 1) * sizeof(T)));
  if (v.size()) BSG_CUDA(cudaMemcpyAsync(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s));
  return BSG_OK;
}

struct CorScratch {
  uint8_t *M = nullptr;
  int *wlen = nullptr, *reach = nullptr;
  long long *boff = nullptr;
  double *thr = nullptr, *band = nullptr, *res = nullptr;
  uint8_t *keep = nullptr;
  ~CorScratch() {
    void *p[] = {M, wlen, reach, boff, thr, band, res, keep};
    for (void *q : p)
      if (q) cudaFree(q);
  }
};

static int cor_common(bsg_bed *h, const int *ind_row, int nr, const int *ind_col, int nc, double size,
                      const double *pos, const double *thr, bool ld, Window &w, CorScratch &sc) {
  cudaStream_t s = h->stream;
  const int *d_row = nullptr, *d_col = nullptr;
  BSG_TRY(upload_index(h, ind_row, nr, h->n, h->w_idx_row, &d_row));
  BSG_TRY(upload_index(h, ind_col, nc, h->m, h->w_idx_col, &d_col));
  build_window(pos, nc, size, w);
  int64_t stride = round_up(((int64_t)nr + 3) / 4, 16);
  if (stride < 16) stride = 16;
  BSG_CUDA(cudaMalloc((void **)&sc.M, (size_t)stride * (nc > 0 ? nc : 1)));
  if (nc > 0) {
    int64_t work = (int64_t)nc * (stride / 4);
    int grid = (int)std::min<int64_t>((work + 255) / 256, 148 * 32);
    k_compact<<<grid, 256, 0, s>>>(h->A, h->strideA, d_row, nr, d_col, nc, sc.M, stride);
    count_launch();
  }
  BSG_TRY(to_dev(&sc.wlen, w.wlen, s));
  BSG_TRY(to_dev(&sc.boff, w.boff, s));
  BSG_CUDA(cudaMalloc((void **)&sc.band, (size_t)(w.total ? w.total : 1) * sizeof(double)));
  if (!ld) {
    std::vector<double> t(thr, thr + nr);
    if (t.empty()) t.push_back(0.0);
    BSG_TRY(to_dev(&sc.thr, t, s));
    BSG_CUDA(cudaMalloc((void **)&sc.keep, (size_t)(w.total ? w.total : 1)));
  }
  if (nc > 0) {
    if (ld)
      k_cor_pairs<true><<<nc, 256, 0, s>>>(sc.M, stride, nr, nc, sc.wlen, sc.boff, nullptr, sc.band, nullptr);
    else
      k_cor_pairs<false><<<nc, 256, 0, s>>>(sc.M, stride, nr, nc, sc.wlen, sc.boff, sc.thr, sc.band, sc.keep);
    count_launch();
  }
  BSG_CUDA(cudaGetLastError());
  return BSG_OK;
}

}  // namespace bsg

using namespace bsg;

extern "C" {

int bsg_cor(bsg_bed *h, const int *ind_row, int nr, const int *ind_col, int nc, double size, const double *thr,
            const double *pos, int fill_diag, int64_t *p, int **pi, double **px) {
  if (!h || !thr || !pos || !p || !pi || !px) return fail(BSG_ERR_ARG, "null argument");
  BSG_TRY(bind_device(h));
  if (!ind_row) nr = h->n;
  if (!ind_col) nc = h->m;
  *pi = nullptr;
  *px = nullptr;
  Window w;
  CorScratch sc;
  BSG_TRY(cor_common(h, ind_row, nr, ind_col, nc, size, pos, thr, false, w, sc));
  std::vector<double> band((size_t)w.total);
  std::vector<uint8_t> keep((size_t)w.total);
  if (w.total) {
    BSG_CUDA(cudaMemcpyAsync(band.data(), sc.band, (size_t)w.total * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    BSG_CUDA(cudaMemcpyAsync(keep.data(), sc.keep, (size_t)w.total, cudaMemcpyDeviceToHost, h->stream));
  }
  BSG_CUDA(cudaStreamSynchronize(h->stream));
  // CSC assembly: ascending row index, diagonal last (rev() of src/corr.cpp:90-92)
  long long nnz = 0;
  for (int j0 = 0; j0 < nc; j0++) {
    p[j0] = nnz;
    for (int k = 0; k < w.wlen[j0]; k++) nnz += keep[(size_t)(w.boff[j0] + k)];
    nnz += fill_diag ? 1 : 0;
  }
  p[nc] = nnz;
  int *oi = (int *)malloc((size_t)(nnz ? nnz : 1) * sizeof(int));
  double *ox = (double *)malloc((size_t)(nnz ? nnz : 1) * sizeof(double));
  if (!oi || !ox) {
    free(oi);
    free(ox);
    return fail(BSG_ERR_ALLOC, "cannot allocate the correlation triplets");
  }
  for (int j0 = 0; j0 < nc; j0++) {
    long long o = p[j0];
    for (int k = w.wlen[j0] - 1; k >= 0; k--) {
      size_t b = (size_t)(w.boff[j0] + k);
      if (keep[b]) {
        oi[o] = j0 - 1 - k;
        ox[o] = band[b];
        o++;
      }
    }
    if (fill_diag) {
      oi[o] = j0;
      ox[o] = 1.0;
    }
  }
  *pi = oi;
  *px = ox;
  return BSG_OK;
}

int bsg_ld_scores(bsg_bed *h, const int *ind_row, int nr, const int *ind_col, int nc, double size, const double *pos,
                  double *out) {
  if (!h || !pos || !out) return fail(BSG_ERR_ARG, "null argument");
  BSG_TRY(bind_device(h));
  if (!ind_row) nr = h->n;
  if (!ind_col) nc = h->m;
  Window w;
  CorScratch sc;
  BSG_TRY(cor_common(h, ind_row, nr, ind_col, nc, size, pos, nullptr, true, w, sc));
  if (nc == 0) return BSG_OK;
  BSG_TRY(to_dev(&sc.reach, w.reach, h->stream));
  BSG_CUDA(cudaMalloc((void **)&sc.res, (size_t)nc * sizeof(double)));
  k_ld_reduce<<<(nc + 127) / 128, 128, 0, h->stream>>>(sc.band, sc.boff, sc.wlen, sc.reach, nc, sc.res);
  count_launch();
  BSG_CUDA(cudaMemcpyAsync(out, sc.res, (size_t)nc * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  BSG_CUDA(cudaStreamSynchronize(h->stream));
  return BSG_OK;
}

}  // extern "C"
