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
 *   - Conv1DLayer.Forward / Backward (serial 7-deep loops)         go/gotorch/cnn_tdnn.go:85-172, with the height axis
 *     and (time, height) tap list of the conv-relu-batchnorm layer (internal/nnet/forward.go:429-455) added: gotorch's
 *     own Conv1D convolves over time only, the CNN-TDNN front end needs time x height
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


/* ---- Conv layer: Conv1DLayer.Forward / Backward (cnn_tdnn.go:85-172) over [batch, T, Hin, Fin] -> [batch, T, Hout, Fout]
 * with (dt, dh) taps and height subsampling as forward.go:429-455 lays the patches out; weight [(ntaps*Fin) x Fout]
 * (tap-major rows), zero padding outside [0,T) x [0,Hin).  Same loop nest as the Go code: b, tOut, (hOut,) oc, k, ic. */
void gt_conv_forward(const double* x, const double* W, const double* bias, double* y, int batch, int T, int hin, int hout,
                     int sub, int fin, int fout, const int* dt, const int* dh, int ntaps) {
  for (int b = 0; b < batch; ++b)
    for (int t = 0; t < T; ++t)
      for (int ho = 0; ho < hout; ++ho)
        for (int oc = 0; oc < fout; ++oc) {
          double sum = 0.0;
          for (int k = 0; k < ntaps; ++k) {
            const int ti = t + dt[k], hi = ho * sub + dh[k];
            if (ti >= 0 && ti < T && hi >= 0 && hi < hin)
              for (int ic = 0; ic < fin; ++ic)
                sum += x[(((size_t)b * T + ti) * hin + hi) * fin + ic] * W[((size_t)k * fin + ic) * fout + oc];
          }
          sum += bias[oc];
          y[(((size_t)b * T + t) * hout + ho) * fout + oc] = sum;
        }
}
void gt_conv_backward(const double* x, const double* W, const double* gy, double* gW, double* gb, double* gx, int batch,
                      int T, int hin, int hout, int sub, int fin, int fout, const int* dt, const int* dh, int ntaps) {
  memset(gx, 0, sizeof(double) * (size_t)batch * T * hin * fin);
  memset(gW, 0, sizeof(double) * (size_t)ntaps * fin * fout);
  memset(gb, 0, sizeof(double) * fout);
  for (int b = 0; b < batch; ++b)
    for (int t = 0; t < T; ++t)
      for (int ho = 0; ho < hout; ++ho)
        for (int oc = 0; oc < fout; ++oc) {
          const double g = gy[(((size_t)b * T + t) * hout + ho) * fout + oc];
          gb[oc] += g;
          for (int k = 0; k < ntaps; ++k) {
            const int ti = t + dt[k], hi = ho * sub + dh[k];
            if (ti >= 0 && ti < T && hi >= 0 && hi < hin)
              for (int ic = 0; ic < fin; ++ic) {
                const size_t in_idx = (((size_t)b * T + ti) * hin + hi) * fin + ic;
                const size_t w_idx = ((size_t)k * fin + ic) * fout + oc;
                gW[w_idx] += x[in_idx] * g;
                gx[in_idx] += W[w_idx] * g;
              }
          }
        }
}

/* Full CNN-TDNN (BASELINE configs[2]) in gotorch layers, one sequence of T frames:
 *   idct affine 40x40 | ivector affine 100->200 broadcast -> [T x 40 x 6]
 *   6 x (conv 3x3 taps + ReLU): 6->64 (h 40), 64->64, 64->128 (h 40->20), 128->128, 128->256 (h 20->10), 256->256
 *   tdnnf7: TDNN(2560->256,{0}) TDNN(256->1536,{0}) ReLU;  11 x [TDNN(1536->160,{-3,0}) TDNN(160->1536,{0,3}) ReLU]
 *   prefinal-l affine 1536->256; chain: affine 256->1536, ReLU, affine 1536->256, affine 256->pdfs (xent branch: the
 *   same three GEMMs again, forward only).  Forward, then backward with dY = Y through the chain branch.
 * Batch-norm (identity at init) and the bypass add are elementwise passes of negligible cost next to the loops above
 * and are left out.  Affine layers use MatMul (row-parallel over `workers` threads, ops.go:49-81); TDNN / conv / every
 * Backward are single-goroutine loops in the reference and are timed as such.  Returns wall seconds. */
typedef struct { int kind; int in, out, nctx, ctx[2]; int hin, hout, sub, fin, fout; double *W, *b, *x, *y, *gW, *gb; } gt_layer;
enum { GT_AFFINE, GT_TDNN, GT_RELU, GT_CONV };
static const int kTapDt[9] = {-1, -1, -1, 0, 0, 0, 1, 1, 1}, kTapDh[9] = {-1, 0, 1, -1, 0, 1, -1, 0, 1};
static void gt_layer_alloc(gt_layer* l, int T, uint64_t* st) {
  size_t wn = 0, out_elems = (size_t)T * l->out;
  if (l->kind == GT_AFFINE) wn = (size_t)l->in * l->out;
  if (l->kind == GT_TDNN) wn = (size_t)l->nctx * l->in * l->out;
  if (l->kind == GT_CONV) { wn = (size_t)9 * l->fin * l->fout; out_elems = (size_t)T * l->hout * l->fout; }
  l->W = l->b = l->gW = l->gb = NULL;
  if (wn) {
    const int bo = l->kind == GT_CONV ? l->fout : l->out;
    l->W = (double*)malloc(sizeof(double) * wn); l->gW = (double*)malloc(sizeof(double) * wn);
    l->b = (double*)calloc(bo, sizeof(double)); l->gb = (double*)calloc(bo, sizeof(double));
    fill(l->W, wn, sqrt(6.0 / (double)(wn / bo + bo)), st);
  }
  l->y = (double*)malloc(sizeof(double) * out_elems);
}
static void gt_layer_forward(gt_layer* l, const double* x, int T, int workers) {
  l->x = (double*)x;
  switch (l->kind) {
    case GT_AFFINE: gt_affine_forward(x, l->W, l->b, l->y, T, l->in, l->out, workers); break;
    case GT_TDNN: gt_tdnn_forward(x, l->W, l->b, l->y, 1, T, l->in, l->out, l->ctx, l->nctx); break;
    case GT_RELU: gt_relu_forward(x, l->y, (size_t)T * l->out); break;
    case GT_CONV: gt_conv_forward(x, l->W, l->b, l->y, 1, T, l->hin, l->hout, l->sub, l->fin, l->fout, kTapDt, kTapDh, 9); break;
  }
}
static void gt_layer_backward(gt_layer* l, const double* gy, double* gx, int T) {
  switch (l->kind) {
    case GT_AFFINE: gt_affine_backward(l->x, l->W, gy, l->gW, l->gb, gx, T, l->in, l->out); break;
    case GT_TDNN: gt_tdnn_backward(l->x, l->W, gy, l->gW, l->gb, gx, 1, T, l->in, l->out, l->ctx, l->nctx); break;
    case GT_RELU: gt_relu_backward(l->x, gy, gx, (size_t)T * l->out); break;
    case GT_CONV: gt_conv_backward(l->x, l->W, gy, l->gW, l->gb, gx, 1, T, l->hin, l->hout, l->sub, l->fin, l->fout, kTapDt, kTapDh, 9); break;
  }
}
static gt_layer gt_mk(int kind, int in, int out) { gt_layer l; memset(&l, 0, sizeof(l)); l.kind = kind; l.in = in; l.out = out; l.nctx = 1; return l; }
static gt_layer gt_mk_tdnn(int in, int out, int c0, int c1, int nctx) { gt_layer l = gt_mk(GT_TDNN, in, out); l.nctx = nctx; l.ctx[0] = c0; l.ctx[1] = c1; return l; }
static gt_layer gt_mk_conv(int hin, int hout, int sub, int fin, int fout) {
  gt_layer l = gt_mk(GT_CONV, hin * fin, hout * fout); l.hin = hin; l.hout = hout; l.sub = sub; l.fin = fin; l.fout = fout; return l;
}
double gt_bench_cnn_tdnn(int T, int pdfs, int backward, int workers, double* checksum) {
  uint64_t st = 42;
  gt_layer L[96]; int n = 0;
  /* trunk (sequential): conv front end ... prefinal-l */
  L[n++] = gt_mk_conv(40, 40, 1, 6, 64);  L[n++] = gt_mk(GT_RELU, 2560, 2560);
  L[n++] = gt_mk_conv(40, 40, 1, 64, 64); L[n++] = gt_mk(GT_RELU, 2560, 2560);
  L[n++] = gt_mk_conv(40, 20, 2, 64, 128); L[n++] = gt_mk(GT_RELU, 2560, 2560);
  L[n++] = gt_mk_conv(20, 20, 1, 128, 128); L[n++] = gt_mk(GT_RELU, 2560, 2560);
  L[n++] = gt_mk_conv(20, 10, 2, 128, 256); L[n++] = gt_mk(GT_RELU, 2560, 2560);
  L[n++] = gt_mk_conv(10, 10, 1, 256, 256); L[n++] = gt_mk(GT_RELU, 2560, 2560);
  L[n++] = gt_mk_tdnn(2560, 256, 0, 0, 1); L[n++] = gt_mk_tdnn(256, 1536, 0, 0, 1); L[n++] = gt_mk(GT_RELU, 1536, 1536);
  for (int i = 0; i < 11; ++i) {
    L[n++] = gt_mk_tdnn(1536, 160, -3, 0, 2); L[n++] = gt_mk_tdnn(160, 1536, 0, 3, 2); L[n++] = gt_mk(GT_RELU, 1536, 1536);
  }
  L[n++] = gt_mk(GT_AFFINE, 1536, 256);                       /* prefinal-l */
  const int trunk = n;
  for (int br = 0; br < 2; ++br) {                            /* chain branch, then the xent branch (forward only) */
    L[n++] = gt_mk(GT_AFFINE, 256, 1536); L[n++] = gt_mk(GT_RELU, 1536, 1536); L[n++] = gt_mk(GT_AFFINE, 1536, 256);
    L[n++] = gt_mk(GT_AFFINE, 256, pdfs);
  }
  for (int i = 0; i < n; ++i) gt_layer_alloc(&L[i], T, &st);
  /* input side: idct affine on the 40 features, ivector affine (one row) broadcast into 5 of the 6 input filters */
  gt_layer idct = gt_mk(GT_AFFINE, 40, 40), ivl = gt_mk(GT_AFFINE, 100, 200);
  gt_layer_alloc(&idct, T, &st); gt_layer_alloc(&ivl, 1, &st);
  double* feats = (double*)malloc(sizeof(double) * (size_t)T * 40);
  double* ivec = (double*)malloc(sizeof(double) * 100);
  double* x0 = (double*)malloc(sizeof(double) * (size_t)T * 240);
  fill(feats, (size_t)T * 40, 10.0, &st); fill(ivec, 100, 1.0, &st);
  size_t maxw = 0;
  for (int i = 0; i < n; ++i) { size_t e = (size_t)T * (L[i].in > L[i].out ? L[i].in : L[i].out); if (e > maxw) maxw = e; }
  double* gA = (double*)malloc(sizeof(double) * maxw);
  double* gB = (double*)malloc(sizeof(double) * maxw);

  const double t0 = now_s();
  gt_layer_forward(&idct, feats, T, workers);
  gt_layer_forward(&ivl, ivec, 1, workers);
  for (int t = 0; t < T; ++t)
    for (int h = 0; h < 40; ++h) {
      x0[((size_t)t * 40 + h) * 6] = idct.y[(size_t)t * 40 + h];
      for (int f = 0; f < 5; ++f) x0[((size_t)t * 40 + h) * 6 + 1 + f] = ivl.y[f * 40 + h];
    }
  const double* cur = x0;
  for (int i = 0; i < trunk; ++i) { gt_layer_forward(&L[i], cur, T, workers); cur = L[i].y; }
  const double* pl = cur;
  for (int i = trunk; i < trunk + 4; ++i) { gt_layer_forward(&L[i], cur, T, workers); cur = L[i].y; }
  const double* out = cur;
  cur = pl;
  for (int i = trunk + 4; i < n; ++i) { gt_layer_forward(&L[i], cur, T, workers); cur = L[i].y; }
  double cs = 0;
  if (backward) {
    memcpy(gA, out, sizeof(double) * (size_t)T * pdfs);       /* dY = Y */
    double *gy = gA, *gx = gB;
    for (int i = trunk + 3; i >= 0; --i) { gt_layer_backward(&L[i], gy, gx, T); double* tmp = gy; gy = gx; gx = tmp; }
    /* gradient wrt the 5 ivector filters -> ivector-linear (the idct has no parameters) */
    double* giv = (double*)calloc(200, sizeof(double));
    for (int t = 0; t < T; ++t)
      for (int h = 0; h < 40; ++h)
        for (int f = 0; f < 5; ++f) giv[f * 40 + h] += gy[((size_t)t * 40 + h) * 6 + 1 + f];
    double* gi = (double*)malloc(sizeof(double) * 100);
    gt_layer_backward(&ivl, giv, gi, 1);
    for (int i = 0; i < 100; ++i) cs += gi[i];
    free(giv); free(gi);
  } else {
    for (size_t i = 0; i < (size_t)T * pdfs; ++i) cs += out[i];
  }
  const double dt = now_s() - t0;
  if (checksum) *checksum = cs;
  for (int i = 0; i < n; ++i) { free(L[i].W); free(L[i].b); free(L[i].gW); free(L[i].gb); free(L[i].y); }
  free(idct.W); free(idct.b); free(idct.gW); free(idct.gb); free(idct.y);
  free(ivl.W); free(ivl.b); free(ivl.gW); free(ivl.gb); free(ivl.y);
  free(feats); free(ivec); free(x0); free(gA); free(gB);
  return dt;
}
