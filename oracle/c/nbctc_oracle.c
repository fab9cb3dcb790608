/*
 * nbctc_oracle.c -- plain-C float64 restatement of the reference's no-blank CTC losses.
 * TEST INFRASTRUCTURE ONLY (checker + reported CPU baseline); never linked into libnbctc.so.
 *
 * Parity: pinned.  tests/test_oracle_cport.py checks this file against oracle/restatement.py,
 * which is itself pinned to outputs of the reference (tests/golden/, tests/golden/make_golden.py).
 *
 * What it restates (file:line under /root/reference):
 *   log_softmax ................ NoBlankCTC.py:136          sigmoid ............ NoBlankBinaryCTC.py:146
 *   emission gather ............ NoBlankCTC.py:96-102       -BCELoss emission .. NoBlankBinaryCTC.py:109-112
 *   alpha step (logaddexp) ..... NoBlankCTC.py:71-87 (+ _logsumexp :16-19), initial state :92-93
 *   read-out alpha[T_b-1,L_b-1]  NoBlankCTC.py:58-68, :139  mean over batch .... NoBlankCTC.py:140
 * The gradient the reference gets from autograd (train.py:444) is evaluated in closed form:
 * beta recursion, gamma = exp(alpha+beta-ll), softmax - scatter(gamma).
 * One OpenMP thread per sequence (sequences are independent).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static double lae(double a, double b) {
  double m = a > b ? a : b;
  if (m == -INFINITY) return -INFINITY;
  return m + log1p(exp(-fabs(a - b)));
}

int nbctc_oracle_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* kind: 0 = NoBlankCTC (targets = int32 labels (B,Lmax)), 1 = NoBlankBinaryCTC (targets = float (B,Lmax,C)).
 * logits float32 (T,B,C); per_seq double (B); grad double (T,B,C) or NULL; w[b] multiplies sequence b's gradient. */
int nbctc_oracle(int kind, const float* logits, int64_t T, int64_t B, int64_t C, const void* targets, int64_t Lmax,
                 const int64_t* in_len, const int64_t* tgt_len, const double* w, double* per_seq, double* grad) {
  int err = 0;
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t b = 0; b < B; ++b) {
    const int64_t Tb = in_len[b], Lb = tgt_len[b];
    const int32_t* lab = kind == 0 ? (const int32_t*)targets + b * Lmax : NULL;
    const float* y = kind == 1 ? (const float*)targets + b * Lmax * C : NULL;
    int ok = Lb >= 1 && Lb <= Lmax && Tb >= Lb && Tb <= T;
    if (ok && kind == 0)
      for (int64_t s = 0; s < Lb; ++s) ok &= lab[s] >= 0 && lab[s] < C;
    if (grad)
      for (int64_t t = 0; t < T; ++t) memset(grad + (t * B + b) * C, 0, sizeof(double) * C);
    if (!ok) { per_seq[b] = INFINITY; continue; }
    double* E = (double*)malloc(sizeof(double) * Tb * Lb);     /* log emissions */
    double* A = (double*)malloc(sizeof(double) * Tb * Lb);     /* alpha */
    double* rowc = (double*)malloc(sizeof(double) * Tb);       /* lse (kind 0) */
    double* bcur = (double*)malloc(sizeof(double) * (Lb + 1));
    double* bnxt = (double*)malloc(sizeof(double) * (Lb + 1));
    double* ls = (double*)malloc(sizeof(double) * C);
    double* l1s = (double*)malloc(sizeof(double) * C);
    if (!E || !A || !rowc || !bcur || !bnxt || !ls || !l1s) { err = 1; per_seq[b] = NAN; goto done; }
    for (int64_t t = 0; t < Tb; ++t) {
      const float* x = logits + (t * B + b) * C;
      if (kind == 0) {
        double m = -INFINITY, s = 0.0;
        for (int64_t c = 0; c < C; ++c) m = x[c] > m ? x[c] : m;
        for (int64_t c = 0; c < C; ++c) s += exp((double)x[c] - m);
        rowc[t] = m + log(s);
        for (int64_t st = 0; st < Lb; ++st) E[t * Lb + st] = (double)x[lab[st]] - rowc[t];
      } else {
        for (int64_t c = 0; c < C; ++c) {
          double v = x[c];
          double a = -(fmax(-v, 0.0) + log1p(exp(-fabs(v))));   /* log sigmoid(v)     */
          double d = -(fmax(v, 0.0) + log1p(exp(-fabs(v))));    /* log (1-sigmoid(v)) */
          ls[c] = a < -100.0 ? -100.0 : a;                       /* nn.BCELoss clamp   */
          l1s[c] = d < -100.0 ? -100.0 : d;
        }
        for (int64_t st = 0; st < Lb; ++st) {
          const float* ys = y + st * C;
          double acc = 0.0;
          for (int64_t c = 0; c < C; ++c) acc += (double)ys[c] * ls[c] + (1.0 - (double)ys[c]) * l1s[c];
          E[t * Lb + st] = acc / (double)C;
        }
      }
    }
    for (int64_t s = 0; s < Lb; ++s) A[s] = s == 0 ? E[0] : -INFINITY;
    for (int64_t t = 1; t < Tb; ++t)
      for (int64_t s = 0; s < Lb; ++s) {
        double prev = A[(t - 1) * Lb + s];
        double adv = s > 0 ? A[(t - 1) * Lb + s - 1] : -INFINITY;
        A[t * Lb + s] = lae(prev, adv) + E[t * Lb + s];
      }
    {
      const double ll = A[(Tb - 1) * Lb + Lb - 1];
      per_seq[b] = -ll;
      if (grad && ll > -INFINITY) {
        const double wb = w ? w[b] : 1.0;
        for (int64_t s = 0; s <= Lb; ++s) bnxt[s] = -INFINITY;  /* holds beta_{t+1}+E_{t+1}; slot Lb = -inf */
        for (int64_t t = Tb - 1; t >= 0; --t) {
          const float* x = logits + (t * B + b) * C;
          double* g = grad + (t * B + b) * C;
          if (kind == 0)
            for (int64_t c = 0; c < C; ++c) g[c] = exp((double)x[c] - rowc[t]);
          else
            for (int64_t c = 0; c < C; ++c) g[c] = 1.0 / (1.0 + exp(-(double)x[c]));
          for (int64_t s = 0; s < Lb; ++s) {
            double beta = t == Tb - 1 ? (s == Lb - 1 ? 0.0 : -INFINITY) : lae(bnxt[s], bnxt[s + 1]);
            double gam = exp(A[t * Lb + s] + beta - ll);
            bcur[s] = beta + E[t * Lb + s];
            if (kind == 0) {
              g[lab[s]] -= gam;
            } else {
              const float* ys = y + s * C;
              for (int64_t c = 0; c < C; ++c) g[c] -= gam * (double)ys[c];
            }
          }
          bcur[Lb] = -INFINITY;
          { double* tmp = bcur; bcur = bnxt; bnxt = tmp; }
          const double sc = kind == 0 ? wb : wb / (double)C;
          for (int64_t c = 0; c < C; ++c) g[c] *= sc;
        }
      }
    }
  done:
    free(E); free(A); free(rowc); free(bcur); free(bnxt); free(ls); free(l1s);
  }
  return err;
}
