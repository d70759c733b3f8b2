/* gotorch_port.c -- CPU restatement of the reference's CPU engine (go/gotorch), float64.
 *
 * TEST / BASELINE INFRASTRUCTURE ONLY: bench.py's cpu_baseline and `--impl reference` legs time
 * it; nothing in kaldi_fp16_b200/ links or calls it.  The reference's CPU path is Go and there is
 * no Go toolchain in this image, so its loops are restated here in C, statement for statement,
 * with the same loop order, the same float64 arithmetic and the same threading:
 *   - MatMul: rows split over NumCPU goroutines, i-j-k loops      go/gotorch/ops.go:15-81
 *   - AffineLayer.Forward (MatMul + bias) / Backward (serial)     go/gotorch/layers.go:57-110
 *   - TDNNLayer.Forward / Backward (serial, clamped context)      go/gotorch/layers.go:444-524
 *   - ReLULayer                                                   go/gotorch/layers.go:134-153
 * Only MatMul is multi-threaded in the reference; TDNNLayer and every Backward are single
 * goroutine loops, and that is what is timed here.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* ---- MatMul (ops.go:49-81) */
typedef struct { const double *a, *b; double* c; int K, N, start, end; } mm_job;
static void* mm_worker(void* arg) {
  mm_job* j = (mm_job*)arg;
  for (int i = j->start; i < j->end; ++i)
    for (int n = 0; n < j->N; ++n) {
      double sum = 0.0;
      for (int k = 0; k < j->K; ++k) sum += j->a[(size_t)i * j->K + k] * j->b[(size_t)k * j->N + n];
      j->c[(size_t)i * j->N + n] = sum;
    }
  return NULL;
}
void gt_matmul(const double* a, const double* b, double* c, int M, int K, int N, int workers) {
  if ((long long)M * N * K <= 10000 || workers <= 1) {   /* matmulNaive */
    mm_job j = {a, b, c, K, N, 0, M};
    mm_worker(&j);
    return;
  }
  if (workers > M) workers = M;
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * workers);
  mm_job* jobs = (mm_job*)malloc(sizeof(mm_job) * workers);
  const int per = (M + workers - 1) / workers;
  int started = 0;
  for (int w = 0; w < workers; ++w) {
    int s = w * per, e = s + per;
    if (e > M) e = M;
    if (s >= e) break;
    jobs[w] = (mm_job){a, b, c, K, N, s, e};
    pthread_create(&th[w], NULL, mm_worker, &jobs[w]);
    ++started;
  }
  for (int w = 0; w < started; ++w) pthread_join(th[w], NULL);
  free(th);
  free(jobs);
}

/* ---- AffineLayer (layers.go:57-110) */
void gt_affine_forward(const double* x, const double* W, const double* bias, double* y, int batch, int in, int out,
                       int workers) {
  gt_matmul(x, W, y, batch, in, out, workers);
  for (int i = 0; i < batch; ++i)
    for (int j = 0; j < out; ++j) y[(size_t)i * out + j] += bias[j];
}
void gt_affine_backward(const double* x, const double* W, const double* gy, double* gW, double* gb, double* gx,
                        int batch, int in, int out) {
  for (int j = 0; j < out; ++j) gb[j] = 0;
  for (int i = 0; i < batch; ++i)
    for (int j = 0; j < out; ++j) gb[j] += gy[(size_t)i * out + j];
  memset(gW, 0, sizeof(double) * in * out);
  for (int b = 0; b < batch; ++b)
    for (int i = 0; i < in; ++i)
      for (int j = 0; j < out; ++j) gW[(size_t)i * out + j] += x[(size_t)b * in + i] * gy[(size_t)b * out + j];
  for (int b = 0; b < batch; ++b)
    for (int i = 0; i < in; ++i) {
      double sum = 0.0;
      for (int j = 0; j < out; ++j) sum += gy[(size_t)b * out + j] * W[(size_t)i * out + j];
      gx[(size_t)b * in + i] = sum;
    }
}

/* ---- TDNNLayer (layers.go:444-524): input [batch, T, in], weight [(nctx*in) x out], clamped context */
void gt_tdnn_forward(const double* x, const double* W, const double* bias, double* y, int batch, int T, int in,
                     int out, const int* ctx, int nctx) {
  for (int b = 0; b < batch; ++b)
    for (int t = 0; t < T; ++t)
      for (int o = 0; o < out; ++o) {
        double sum = bias[o];
        for (int ci = 0; ci < nctx; ++ci) {
          int tc = t + ctx[ci];
          if (tc < 0) tc = 0; else if (tc >= T) tc = T - 1;
          for (int i = 0; i < in; ++i)
            sum += x[((size_t)b * T + tc) * in + i] * W[((size_t)ci * in + i) * out + o];
        }
        y[((size_t)b * T + t) * out + o] = sum;
      }
}
void gt_tdnn_backward(const double* x, const double* W, const double* gy, double* gW, double* gb, double* gx,
                      int batch, int T, int in, int out, const int* ctx, int nctx) {
  memset(gx, 0, sizeof(double) * (size_t)batch * T * in);
  memset(gW, 0, sizeof(double) * (size_t)nctx * in * out);
  memset(gb, 0, sizeof(double) * out);
  for (int b = 0; b < batch; ++b)
    for (int t = 0; t < T; ++t)
      for (int o = 0; o < out; ++o) {
        const double g = gy[((size_t)b * T + t) * out + o];
        gb[o] += g;
        for (int ci = 0; ci < nctx; ++ci) {
          int tc = t + ctx[ci];
          if (tc < 0) tc = 0; else if (tc >= T) tc = T - 1;
          for (int i = 0; i < in; ++i) {
            const size_t in_idx = ((size_t)b * T + tc) * in + i;
            const size_t w_idx = ((size_t)ci * in + i) * out + o;
            gW[w_idx] += x[in_idx] * g;
            gx[in_idx] += W[w_idx] * g;
          }
        }
      }
}

/* ---- ReLULayer (layers.go:134-153) */
void gt_relu_forward(const double* x, double* y, size_t n) { for (size_t i = 0; i < n; ++i) y[i] = x[i] > 0 ? x[i] : 0; }
void gt_relu_backward(const double* x, const double* gy, double* gx, size_t n) { for (size_t i = 0; i < n; ++i) gx[i] = x[i] > 0 ? gy[i] : 0; }

/* deterministic pseudo-random fill in [-scale, scale) */
static void fill(double* p, size_t n, double scale, uint64_t* st) {
  for (size_t i = 0; i < n; ++i) {
    *st = *st * 6364136223846793005ULL + 1442695040888963407ULL;
    p[i] = (((double)((*st >> 11) & 0xFFFFFFFFFFFFFULL)) / 4503599627370496.0 * 2.0 - 1.0) * scale;
  }
}

/* TDNN-F stack (BASELINE configs[1]) in gotorch layers: per layer TDNNLayer(dim->bott, [-s,0]) ->
 * TDNNLayer(bott->dim, [0,s]) -> ReLU, forward then backward with dY = Y.  Returns wall seconds;
 * *checksum = sum of the final input-gradient (keeps the optimiser from dropping work). */
double gt_bench_tdnnf_stack(int layers, int dim, int bott, int stride, int batch, int T, int backward, double* checksum) {
  uint64_t st = 42;
  const size_t act = (size_t)batch * T * dim, bt = (size_t)batch * T * bott;
  const int ctx1[2] = {-stride, 0}, ctx2[2] = {0, stride};
  double** Wl = (double**)malloc(sizeof(double*) * layers);
  double** Wa = (double**)malloc(sizeof(double*) * layers);
  double** X = (double**)malloc(sizeof(double*) * (layers + 1));   /* layer inputs */
  double** B = (double**)malloc(sizeof(double*) * layers);
  double** Z = (double**)malloc(sizeof(double*) * layers);         /* pre-ReLU */
  double* bl = (double*)calloc(bott, sizeof(double));
  double* ba = (double*)calloc(dim, sizeof(double));
  for (int l = 0; l < layers; ++l) {
    Wl[l] = (double*)malloc(sizeof(double) * 2 * dim * bott);
    Wa[l] = (double*)malloc(sizeof(double) * 2 * bott * dim);
    fill(Wl[l], (size_t)2 * dim * bott, sqrt(2.0 / (2 * dim + bott)), &st);
    fill(Wa[l], (size_t)2 * bott * dim, sqrt(2.0 / (2 * bott + dim)), &st);
    B[l] = (double*)malloc(sizeof(double) * bt);
    Z[l] = (double*)malloc(sizeof(double) * act);
  }
  for (int l = 0; l <= layers; ++l) X[l] = (double*)malloc(sizeof(double) * act);
  fill(X[0], act, 1.0, &st);
  double* gW1 = (double*)malloc(sizeof(double) * 2 * dim * bott);
  double* gW2 = (double*)malloc(sizeof(double) * 2 * bott * dim);
  double* gb1 = (double*)malloc(sizeof(double) * bott);
  double* gb2 = (double*)malloc(sizeof(double) * dim);
  double* gA = (double*)malloc(sizeof(double) * act);
  double* gZ = (double*)malloc(sizeof(double) * act);
  double* gB = (double*)malloc(sizeof(double) * bt);

  const double t0 = now_s();
  for (int l = 0; l < layers; ++l) {
    gt_tdnn_forward(X[l], Wl[l], bl, B[l], batch, T, dim, bott, ctx1, 2);
    gt_tdnn_forward(B[l], Wa[l], ba, Z[l], batch, T, bott, dim, ctx2, 2);
    gt_relu_forward(Z[l], X[l + 1], act);
  }
  double cs = 0;
  if (backward) {
    memcpy(gA, X[layers], sizeof(double) * act);   /* dY = Y */
    for (int l = layers - 1; l >= 0; --l) {
      gt_relu_backward(Z[l], gA, gZ, act);
      gt_tdnn_backward(B[l], Wa[l], gZ, gW2, gb2, gB, batch, T, bott, dim, ctx2, 2);
      gt_tdnn_backward(X[l], Wl[l], gB, gW1, gb1, gA, batch, T, dim, bott, ctx1, 2);
    }
    for (size_t i = 0; i < act; ++i) cs += gA[i];
  } else {
    for (size_t i = 0; i < act; ++i) cs += X[layers][i];
  }
  const double dt = now_s() - t0;
  if (checksum) *checksum = cs;
  for (int l = 0; l < layers; ++l) { free(Wl[l]); free(Wa[l]); free(B[l]); free(Z[l]); }
  for (int l = 0; l <= layers; ++l) free(X[l]);
  free(Wl); free(Wa); free(X); free(B); free(Z); free(bl); free(ba);
  free(gW1); free(gW2); free(gb1); free(gb2); free(gA); free(gZ); free(gB);
  return dt;
}
