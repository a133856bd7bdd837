/* ORACLE (test infrastructure, not product): plain-C restatement of the reference's UNet forward pass.
 *
 * Follows src/unet/model/unet.py of uibk-uncover/ws-unet:
 *   conv3x3_reflect  = nn.Conv2d(k=3, padding=1, padding_mode='reflect') + bias      unet.py:73,82-132
 *   relu             = F.relu                                                       unet.py:141-186
 *   maxpool2         = nn.MaxPool2d(2, 2)                                           unet.py:85,144
 *   upconv2          = nn.ConvTranspose2d(k=2, stride=2) + bias (no activation)     unet.py:74,113-131,177
 *   concat           = torch.cat([up, skip], dim=1)  (up first)                     unet.py:178
 *   head             = sigmoid(Conv2d 1x1)                                          unet.py:135,189
 * All tensors are NCHW fp32 like the reference; accumulation is in double so the oracle sits below the
 * reference's own fp32 rounding noise (1.4e-5 px). Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg may call this.
 */
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

/* minimal parallel-for on pthreads (libgomp is not in the image). WSO_THREADS overrides the thread count. */
typedef void (*wso_body)(int idx, void* ctx);
typedef struct { wso_body fn; void* ctx; int n; int next; pthread_mutex_t mu; } wso_job;
static void* wso_worker(void* p) {
  wso_job* j = (wso_job*)p;
  for (;;) {
    pthread_mutex_lock(&j->mu);
    const int i = j->next++;
    pthread_mutex_unlock(&j->mu);
    if (i >= j->n) return NULL;
    j->fn(i, j->ctx);
  }
}
int wso_num_threads(void) {
  const char* e = getenv("WSO_THREADS");
  int n = e ? atoi(e) : (int)sysconf(_SC_NPROCESSORS_ONLN);
  if (n < 1) n = 1;
  if (n > 256) n = 256;
  return n;
}
static void parallel_for(int n, wso_body fn, void* ctx) {
  int nt = wso_num_threads();
  if (nt > n) nt = n;
  wso_job j = {fn, ctx, n, 0, PTHREAD_MUTEX_INITIALIZER};
  if (nt <= 1) { wso_worker(&j); return; }
  pthread_t th[256];
  for (int t = 0; t < nt; ++t) pthread_create(&th[t], NULL, wso_worker, &j);
  for (int t = 0; t < nt; ++t) pthread_join(th[t], NULL);
}

static inline int reflect(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }

/* y[b][co][h][w] = bias[co] + sum_{ci,dy,dx} w[co][ci][dy][dx] * x[b][ci][reflect(h+dy-1)][reflect(w+dx-1)] */
typedef struct { const float *x, *w, *bias; float* y; int B, Cin, Cout, H, W, relu; } conv_ctx;
static void conv_body(int idx, void* vp) {
  const conv_ctx* c = (const conv_ctx*)vp;
  const float *x = c->x, *w = c->w, *bias = c->bias;
  float* y = c->y;
  const int Cin = c->Cin, Cout = c->Cout, H = c->H, W = c->W, relu = c->relu;
  const int b = idx / Cout, co = idx % Cout;
  {
    {
      double* acc = (double*)malloc(sizeof(double) * (size_t)H * W);
      for (int i = 0; i < H * W; ++i) acc[i] = bias ? bias[co] : 0.0;
      for (int ci = 0; ci < Cin; ++ci) {
        const float* xp = x + ((size_t)b * Cin + ci) * H * W;
        const float* wp = w + ((size_t)co * Cin + ci) * 9;
        for (int dy = 0; dy < 3; ++dy)
          for (int dx = 0; dx < 3; ++dx) {
            const double wv = wp[dy * 3 + dx];
            for (int h = 0; h < H; ++h) {
              const float* row = xp + (size_t)reflect(h + dy - 1, H) * W;
              double* arow = acc + (size_t)h * W;
              /* interior columns vectorise; the two border columns use the mirrored index */
              arow[0] += wv * row[reflect(dx - 1, W)];
              for (int c = 1; c < W - 1; ++c) arow[c] += wv * row[c + dx - 1];
              if (W > 1) arow[W - 1] += wv * row[reflect(W - 1 + dx - 1, W)];
            }
          }
      }
      float* yp = y + ((size_t)b * Cout + co) * H * W;
      for (int i = 0; i < H * W; ++i) {
        float v = (float)acc[i];
        yp[i] = (relu && v < 0.f) ? 0.f : v;
      }
      free(acc);
    }
  }
}
void wso_conv3x3_reflect(const float* x, const float* w, const float* bias, float* y, int B, int Cin, int Cout, int H,
                         int W, int relu) {
  conv_ctx c = {x, w, bias, y, B, Cin, Cout, H, W, relu};
  parallel_for(B * Cout, conv_body, &c);
}

void wso_maxpool2(const float* x, float* y, int B, int C, int H, int W) {
  const int Ho = H / 2, Wo = W / 2;
  for (int bc = 0; bc < B * C; ++bc) {
    const float* xp = x + (size_t)bc * H * W;
    float* yp = y + (size_t)bc * Ho * Wo;
    for (int h = 0; h < Ho; ++h)
      for (int c = 0; c < Wo; ++c) {
        float m = xp[(size_t)(2 * h) * W + 2 * c];
        const float v1 = xp[(size_t)(2 * h) * W + 2 * c + 1], v2 = xp[(size_t)(2 * h + 1) * W + 2 * c],
                    v3 = xp[(size_t)(2 * h + 1) * W + 2 * c + 1];
        if (v1 > m) m = v1;
        if (v2 > m) m = v2;
        if (v3 > m) m = v3;
        yp[(size_t)h * Wo + c] = m;
      }
  }
}

/* ConvTranspose2d k=2 s=2: y[b][co][2h+dy][2w+dx] = bias[co] + sum_ci x[b][ci][h][w] * w[ci][co][dy][dx] */
typedef struct { const float *x, *w, *bias; float* y; int B, Cin, Cout, H, W; } up_ctx;
static void up_body(int idx, void* vp) {
  const up_ctx* u = (const up_ctx*)vp;
  const float *x = u->x, *w = u->w, *bias = u->bias;
  float* y = u->y;
  const int Cin = u->Cin, Cout = u->Cout, H = u->H, W = u->W;
  const int b = idx / Cout, co = idx % Cout;
  {
    {
      for (int h = 0; h < H; ++h)
        for (int c = 0; c < W; ++c) {
          double a[4] = {bias[co], bias[co], bias[co], bias[co]};
          for (int ci = 0; ci < Cin; ++ci) {
            const double xv = x[(((size_t)b * Cin + ci) * H + h) * W + c];
            const float* wp = w + ((size_t)ci * Cout + co) * 4;
            a[0] += xv * wp[0];
            a[1] += xv * wp[1];
            a[2] += xv * wp[2];
            a[3] += xv * wp[3];
          }
          float* yp = y + ((size_t)b * Cout + co) * (4 * (size_t)H * W);
          yp[(size_t)(2 * h) * (2 * W) + 2 * c] = (float)a[0];
          yp[(size_t)(2 * h) * (2 * W) + 2 * c + 1] = (float)a[1];
          yp[(size_t)(2 * h + 1) * (2 * W) + 2 * c] = (float)a[2];
          yp[(size_t)(2 * h + 1) * (2 * W) + 2 * c + 1] = (float)a[3];
        }
    }
  }
}
void wso_upconv2(const float* x, const float* w, const float* bias, float* y, int B, int Cin, int Cout, int H, int W) {
  up_ctx u = {x, w, bias, y, B, Cin, Cout, H, W};
  parallel_for(B * Cout, up_body, &u);
}

void wso_concat(const float* a, const float* b, float* y, int B, int Ca, int Cb, int H, int W) {
  const size_t hw = (size_t)H * W;
  for (int n = 0; n < B; ++n) {
    memcpy(y + (size_t)n * (Ca + Cb) * hw, a + (size_t)n * Ca * hw, sizeof(float) * Ca * hw);
    memcpy(y + ((size_t)n * (Ca + Cb) + Ca) * hw, b + (size_t)n * Cb * hw, sizeof(float) * Cb * hw);
  }
}

/* sigmoid(conv1x1): y[b][0][h][w] */
void wso_head(const float* x, const float* w, const float* bias, float* y, int B, int C, int H, int W) {
  const size_t hw = (size_t)H * W;
  for (int b = 0; b < B; ++b)
    for (size_t i = 0; i < hw; ++i) {
      double z = bias[0];
      for (int c = 0; c < C; ++c) z += (double)w[c] * x[((size_t)b * C + c) * hw + i];
      y[(size_t)b * hw + i] = (float)(1.0 / (1.0 + exp(-z)));
    }
}

/* ---- Weighted-Stego arithmetic: src/ws/estimate.py:83-128 on one uint8 image, exact direct stencils in double.
 * kind: 0 KB, 1 AVG, 2 AVG9, 3 identity (estimate.py:31-52). xhat_in: external prediction (H-2)x(W-2) or NULL.
 * out[0] = beta_hat (clipped if clip), out[1] = l1, out[2] = unclipped beta_hat before bias correction. */
static double predict_kind(int kind, const double n[9]) {
  const double cross = n[1] + n[3] + n[5] + n[7], diag = n[0] + n[2] + n[6] + n[8];
  switch (kind) {
    case 0: return (2.0 * cross - diag) / 4.0;
    case 1: return (cross + diag) / 8.0;
    case 2: return (cross + diag + n[4]) / 9.0;
    default: return n[4];
  }
}

void wso_ws_attack(const unsigned char* img, int H, int W, int kind, const float* xhat_in, const float* xbias_in,
                   int weighted, int clip, int correct_bias, double* out) {
  double swr = 0, sw = 0, sl1 = 0, swb = 0;
  for (int y = 1; y < H - 1; ++y)
    for (int x = 1; x < W - 1; ++x) {
      double n[9], dn[9];
      for (int dy = 0; dy < 3; ++dy)
        for (int dx = 0; dx < 3; ++dx) {
          const unsigned char p = img[(size_t)(y + dy - 1) * W + (x + dx - 1)];
          n[dy * 3 + dx] = p;
          dn[dy * 3 + dx] = (double)(p ^ 1) - (double)p; /* x_bar - x, estimate.py:127 */
        }
      const double xv = n[4], xbar = (double)(img[(size_t)y * W + x] ^ 1);
      const size_t ci = (size_t)(y - 1) * (W - 2) + (x - 1);
      const double xhat = xhat_in ? (double)xhat_in[ci] : predict_kind(kind, n);
      double wgt = 1.0;
      if (weighted) {
        double s1 = 0, s2 = 0;
        for (int k = 0; k < 9; ++k)
          if (k != 4) { s1 += n[k]; s2 += n[k] * n[k]; }
        const double var = s2 / 8.0 - (s1 / 8.0) * (s1 / 8.0);
        wgt = weighted == 1 ? 1.0 / (5.0 + var) : 5.0 + var;
      }
      swr += wgt * (xv - xbar) * (xv - xhat);
      sw += wgt;
      sl1 += fabs(xv - xhat);
      if (correct_bias) {
        const double xb = xbias_in ? (double)xbias_in[ci] : predict_kind(kind, dn);
        swb += wgt * (xv - xbar) * xb;
      }
    }
  double beta = swr / sw;
  out[2] = beta;
  if (clip && beta < 0) beta = 0;
  if (correct_bias) beta -= beta * (swb / sw);
  out[0] = beta;
  out[1] = sl1 / ((double)(H - 2) * (W - 2));
}
