/*
 * pll_port.c — ORACLE (test infrastructure, NOT product code). See pll_port.h.
 *
 * Restates, in scalar C, the arithmetic of the reference's forked libpll for the NetRAX hot path.
 * For 4-state (DNA) data the operation ORDER of the kernels that actually run in the reference
 * (the AVX 4x4 bodies, reached from the AVX2 dispatcher: LIBPLL/core_partials_avx2.c:842-857)
 * is reproduced exactly — separate multiply and add, pairwise tree sums (p0+p1)+(p2+p3) — so that
 * CLVs and scaler counts are bit-identical to the reference (compiled with -ffp-contract=off).
 * For other state counts the generic serial order of LIBPLL/core_*.c is used (the reference's
 * AVX2+FMA lane order is not reproducible in scalar code; tolerance documented in tests).
 */
#include "pll_port.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static void *xcalloc(size_t n, size_t sz) {
  void *p = calloc(n ? n : 1, sz);
  if (!p) { fprintf(stderr, "pll_port: out of memory\n"); abort(); }
  return p;
}

port_partition *port_partition_create(unsigned states, unsigned rate_cats, unsigned sites,
                                      unsigned tips, unsigned edges) {
  port_partition *p = (port_partition *)xcalloc(1, sizeof(*p));
  unsigned i;
  p->states = states;
  p->states_padded = (states + 3) & ~3u; /* LIBPLL/pll.c:482-483 */
  p->rate_cats = rate_cats;
  p->sites = sites;
  p->tips = tips;
  p->edges = edges;
  p->freqs = (double *)xcalloc(p->states_padded, sizeof(double));
  p->subst_params = (double *)xcalloc(states * (states - 1) / 2, sizeof(double));
  p->eigenvecs = (double *)xcalloc((size_t)states * p->states_padded, sizeof(double));
  p->inv_eigenvecs = (double *)xcalloc((size_t)states * p->states_padded, sizeof(double));
  p->eigenvals = (double *)xcalloc(p->states_padded, sizeof(double));
  p->rates = (double *)xcalloc(rate_cats, sizeof(double));
  p->rate_weights = (double *)xcalloc(rate_cats, sizeof(double));
  for (i = 0; i < rate_cats; ++i) { /* LIBPLL/pll.c:779-784 */
    p->rates[i] = 1.0;
    p->rate_weights[i] = 1.0 / rate_cats;
  }
  p->nmodels = 1;
  p->cat_model = (unsigned *)xcalloc(rate_cats, sizeof(unsigned));
  p->m_freqs[0] = p->freqs; p->m_subst[0] = p->subst_params; p->m_eigenvecs[0] = p->eigenvecs;
  p->m_inv_eigenvecs[0] = p->inv_eigenvecs; p->m_eigenvals[0] = p->eigenvals;
  p->pattern_weights = (unsigned *)xcalloc(sites, sizeof(unsigned));
  for (i = 0; i < sites; ++i) p->pattern_weights[i] = 1;
  p->tipchars = (unsigned char **)xcalloc(tips, sizeof(unsigned char *));
  for (i = 0; i < tips; ++i) p->tipchars[i] = (unsigned char *)xcalloc(sites, 1);
  p->pmatrix = (double **)xcalloc(edges, sizeof(double *));
  for (i = 0; i < edges; ++i)
    p->pmatrix[i] = (double *)xcalloc((size_t)rate_cats * states * p->states_padded, sizeof(double));
  if (states == 4) { /* set_tipchars_4x4: the code IS the mask, LIBPLL/pll.c:875-900 */
    for (i = 0; i < 16; ++i) p->tipmap[i] = i;
    p->maxstates = 16;
  }
  return p;
}

void port_partition_destroy(port_partition *p) {
  if (p) free(p->invariant);
  unsigned i;
  if (!p) return;
  for (i = 0; i < p->tips; ++i) free(p->tipchars[i]);
  for (i = 0; i < p->edges; ++i) free(p->pmatrix[i]);
  for (i = 1; i < PORT_MAX_MODELS; ++i) {
    free(p->m_freqs[i]); free(p->m_subst[i]); free(p->m_eigenvecs[i]); free(p->m_inv_eigenvecs[i]); free(p->m_eigenvals[i]);
  }
  free(p->cat_model);
  free(p->tipchars); free(p->pmatrix); free(p->pattern_weights);
  free(p->freqs); free(p->subst_params); free(p->eigenvecs); free(p->inv_eigenvecs);
  free(p->eigenvals); free(p->rates); free(p->rate_weights);
  free(p);
}

/* ------------------------------------------------------------------------------------------
 * Discrete Gamma rates.  LIBPLL/gamma.c.  The reference uses Z. Yang's C conversions of
 * published algorithms: AS32 (Bhattacharjee 1970, incomplete gamma ratio), Algorithm 291
 * (Pike & Hill 1966, ln Gamma), AS70 (Odeh & Evans 1974, normal quantile), AS91 (Best & Roberts
 * 1975, chi-square quantile).  Their truncation constants (1e-8, .5e-6) are part of the result
 * at the 1e-7 level, so the same published constants are used here.
 * ---------------------------------------------------------------------------------------- */
int port_set_submodels(port_partition *p, unsigned n, const unsigned *cat_model, const double *freqs, const double *subst) {
  const unsigned states = p->states, sp = p->states_padded, nrates = states * (states - 1) / 2;
  unsigned m, i;
  if (n < 1 || n > PORT_MAX_MODELS) return 0;
  for (i = 0; i < p->rate_cats; ++i)
    if (cat_model[i] >= n) return 0;
  for (m = 0; m < n; ++m) {
    double sum = 0., *sf, *ss, *sev, *siv, *sva;
    if (!p->m_freqs[m]) {
      p->m_freqs[m] = (double *)xcalloc(sp, sizeof(double));
      p->m_subst[m] = (double *)xcalloc(nrates, sizeof(double));
      p->m_eigenvecs[m] = (double *)xcalloc((size_t)states * sp, sizeof(double));
      p->m_inv_eigenvecs[m] = (double *)xcalloc((size_t)states * sp, sizeof(double));
      p->m_eigenvals[m] = (double *)xcalloc(sp, sizeof(double));
    }
    /* pll_set_frequencies (LIBPLL/models.c:445-467): renormalise when |sum - 1| > PLL_MISC_EPSILON */
    for (i = 0; i < states; ++i) { p->m_freqs[m][i] = freqs[(size_t)m * states + i]; sum += p->m_freqs[m][i]; }
    if (fabs(sum - 1.0) > PORT_MISC_EPSILON)
      for (i = 0; i < states; ++i) p->m_freqs[m][i] /= sum;
    memcpy(p->m_subst[m], subst + (size_t)m * nrates, nrates * sizeof(double));
    /* pll_update_eigen(partition, m): the single-matrix routine run on matrix m's arrays */
    sf = p->freqs; ss = p->subst_params; sev = p->eigenvecs; siv = p->inv_eigenvecs; sva = p->eigenvals;
    p->freqs = p->m_freqs[m]; p->subst_params = p->m_subst[m]; p->eigenvecs = p->m_eigenvecs[m];
    p->inv_eigenvecs = p->m_inv_eigenvecs[m]; p->eigenvals = p->m_eigenvals[m];
    port_update_eigen(p);
    p->freqs = sf; p->subst_params = ss; p->eigenvecs = sev; p->inv_eigenvecs = siv; p->eigenvals = sva;
  }
  p->nmodels = n;
  for (i = 0; i < p->rate_cats; ++i) p->cat_model[i] = cat_model[i];
  return 1;
}

int port_update_invariant_sites(port_partition *p) { /* LIBPLL/models.c:651-750, PATTERN_TIP branch */
  unsigned i, j;
  uint32_t gap_state = 0, *inv;
  for (i = 0; i < p->states; ++i) gap_state = (gap_state << 1) | 1;
  if (!p->invariant) p->invariant = (int *)xcalloc(p->sites, sizeof(int));
  inv = (uint32_t *)xcalloc(p->sites, sizeof(uint32_t));
  for (j = 0; j < p->sites; ++j) inv[j] = gap_state;
  for (i = 0; i < p->tips; ++i)
    for (j = 0; j < p->sites; ++j) inv[j] &= p->tipmap[p->tipchars[i][j]];
  for (j = 0; j < p->sites; ++j)
    p->invariant[j] = (inv[j] == 0 || __builtin_popcount(inv[j]) > 1) ? -1 : __builtin_ctz(inv[j]);
  free(inv);
  return 1;
}

int port_set_prop_invar(port_partition *p, double prop_invar) { /* LIBPLL/models.c:495-543 */
  if (prop_invar < 0 || prop_invar >= 1) return 0; /* "Invalid proportion of invariant sites" */
  if (prop_invar > 0 && !p->invariant) port_update_invariant_sites(p);
  p->prop_invar = prop_invar;
  return 1;
}

static double ln_gamma_pike_hill(double alpha) { /* gamma.c:92-118 */
  double x = alpha, f = 0.0, z;
  if (x < 7.0) {
    f = 1.0;
    z = alpha - 1.0;
    for (;;) {
      z += 1.0;
      if (!(z < 7.0)) break;
      f *= z;
    }
    x = z;
    f = -log(f);
  }
  z = 1.0 / (x * x);
  return f + (x - 0.5) * log(x) - x + .918938533204673 +
         (((-.000595238095238 * z + .000793650793651) * z - .002777777777778) * z + .083333333333333) / x;
}

static double incomplete_gamma_as32(double x, double alpha, double ln_gamma_alpha) { /* gamma.c:27-90 */
  const double accurate = 1e-8, overflow = 1e30;
  double p = alpha, g = ln_gamma_alpha, factor, gin, rn, term, a, b, an, dif, pn[6];
  int i;
  if (x == 0) return 0;
  if (x < 0 || p <= 0) return -1;
  factor = exp(p * log(x) - x - g);
  if (!(x > 1 && x >= p)) { /* series expansion */
    gin = 1; term = 1; rn = p;
    do {
      rn += 1;
      term *= x / rn;
      gin += term;
    } while (term > accurate);
    return gin * factor / p;
  }
  /* continued fraction */
  a = 1 - p; b = a + x + 1; term = 0;
  pn[0] = 1; pn[1] = x; pn[2] = x + 1; pn[3] = x * b;
  gin = pn[2] / pn[3];
  for (;;) {
    a += 1; b += 2; term += 1;
    an = a * term;
    for (i = 0; i < 2; i++) pn[i + 4] = b * pn[i + 2] - an * pn[i];
    if (pn[5] != 0) {
      rn = pn[4] / pn[5];
      dif = fabs(gin - rn);
      if (dif <= accurate && dif <= accurate * rn) return 1 - factor * gin;
      gin = rn;
    }
    for (i = 0; i < 4; i++) pn[i] = pn[i + 2];
    if (fabs(pn[4]) >= overflow)
      for (i = 0; i < 4; i++) pn[i] /= overflow;
  }
}

static double point_normal_as70(double prob) { /* gamma.c:120-146 */
  const double a0 = -.322232431088, a1 = -1, a2 = -.342242088547, a3 = -.0204231210245,
               a4 = -.453642210148e-4, b0 = .0993484626060, b1 = .588581570495,
               b2 = .531103462366, b3 = .103537752850, b4 = .0038560700634;
  double p = prob, p1 = (p < 0.5 ? p : 1 - p), y, z;
  if (p1 < 1e-20) return -9999;
  y = sqrt(log(1 / (p1 * p1)));
  z = y + ((((y * a4 + a3) * y + a2) * y + a1) * y + a0) / ((((y * b4 + b3) * y + b2) * y + b1) * y + b0);
  return p < 0.5 ? -z : z;
}

static double point_chi2_as91(double prob, double v) { /* gamma.c:148-214 */
  const double e = .5e-6, aa = .6931471805;
  double p = prob, g, xx, c, ch, a, q, p1, p2, t, x, b, s1, s2, s3, s4, s5, s6;
  if (p < .000002 || p > .999998 || v <= 0) return -1;
  g = ln_gamma_pike_hill(v / 2);
  xx = v / 2; c = xx - 1;
  if (v < -1.24 * log(p)) {
    ch = pow(p * xx * exp(g + xx * aa), 1 / xx);
    if (ch - e < 0) return ch;
  } else if (v > .32) {
    x = point_normal_as70(p);
    p1 = 0.222222 / v;
    ch = v * pow(x * sqrt(p1) + 1 - p1, 3.0);
    if (ch > 2.2 * v + 6) ch = -2 * (log(1 - p) - c * log(.5 * ch) + g);
  } else {
    ch = 0.4; a = log(1 - p);
    do {
      q = ch; p1 = 1 + ch * (4.67 + ch); p2 = ch * (6.73 + ch * (6.66 + ch));
      t = -0.5 + (4.67 + 2 * ch) / p1 - (6.73 + ch * (13.32 + 3 * ch)) / p2;
      ch -= (1 - exp(a + g + .5 * ch + c * aa) * p2 / p1) / t;
    } while (fabs(q / ch - 1) - .01 > 0);
  }
  do {
    q = ch; p1 = .5 * ch;
    t = incomplete_gamma_as32(p1, xx, g);
    if (t < 0.0) return -1;
    p2 = p - t;
    t = p2 * exp(xx * aa + g + p1 - c * log(ch));
    b = t / ch; a = 0.5 * t - b * c;
    s1 = (210 + a * (140 + a * (105 + a * (84 + a * (70 + 60 * a))))) / 420;
    s2 = (420 + a * (735 + a * (966 + a * (1141 + 1278 * a)))) / 2520;
    s3 = (210 + a * (462 + a * (707 + 932 * a))) / 2520;
    s4 = (252 + a * (672 + 1182 * a) + c * (294 + a * (889 + 1740 * a))) / 5040;
    s5 = (84 + 264 * a + c * (175 + 606 * a)) / 2520;
    s6 = (120 + c * (346 + 127 * c)) / 5040;
    ch += t * (1 + 0.5 * t * s1 - b * c * (s1 - b * (s2 - b * (s3 - b * (s4 - b * (s5 - b * s6))))));
  } while (fabs(q / ch - 1) > e);
  return ch;
}

int port_compute_gamma_cats(double alpha, unsigned categories, double *out, int mode) {
  /* gamma.c:267-330; POINT_GAMMA(prob,alpha,beta) = PointChi2(prob, 2 alpha) / (2 beta) */
  unsigned i;
  double factor = alpha / alpha * categories, beta = alpha;
  if (alpha < 0.02 || categories < 1) return 0;
  if (categories == 1) { out[0] = 1.0; return 1; }
  if (mode == 1) { /* median */
    double middle = 1.0 / (2.0 * categories), t = 0.0;
    for (i = 0; i < categories; i++)
      out[i] = point_chi2_as91((double)(i * 2 + 1) * middle, 2.0 * alpha) / (2.0 * beta);
    for (i = 0; i < categories; i++) t += out[i];
    for (i = 0; i < categories; i++) out[i] *= factor / t;
    return 1;
  }
  if (mode == 0) { /* mean */
    double *gp = (double *)xcalloc(categories, sizeof(double));
    double lnga1 = ln_gamma_pike_hill(alpha + 1);
    for (i = 0; i < categories - 1; i++)
      gp[i] = point_chi2_as91((i + 1.0) / categories, 2.0 * alpha) / (2.0 * beta);
    for (i = 0; i < categories - 1; i++) gp[i] = incomplete_gamma_as32(gp[i] * beta, alpha + 1, lnga1);
    out[0] = gp[0] * factor;
    out[categories - 1] = (1 - gp[categories - 2]) * factor;
    for (i = 1; i < categories - 1; i++) out[i] = (gp[i] - gp[i - 1]) * factor;
    free(gp);
    return 1;
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * Eigen-decomposition of the symmetrised rate matrix.  LIBPLL/models.c:24-410.
 * mytred2 / mytqli are the textbook Householder reduction and implicit QL iteration
 * (EISPACK tred2/tql2 lineage) with 1-based loops; restated 0-based here.
 * ---------------------------------------------------------------------------------------- */
static void householder_tridiag(double **a, unsigned n, double *d, double *e) { /* models.c:99-178 */
  int i, j, k, l;
  for (i = (int)n - 1; i >= 1; i--) {
    double h = 0.0, scale = 0.0;
    l = i - 1;
    if (l > 0) {
      for (k = 0; k <= l; k++) scale += fabs(a[k][i]);
      if (scale == 0.0) {
        e[i] = a[l][i];
      } else {
        double f, g, hh;
        for (k = 0; k <= l; k++) {
          a[k][i] /= scale;
          h += a[k][i] * a[k][i];
        }
        f = a[l][i];
        g = (f > 0) ? -sqrt(h) : sqrt(h);
        e[i] = scale * g;
        h -= f * g;
        a[l][i] = f - g;
        f = 0.0;
        for (j = 0; j <= l; j++) {
          a[i][j] = a[j][i] / h;
          g = 0.0;
          for (k = 0; k <= j; k++) g += a[k][j] * a[k][i];
          for (k = j + 1; k <= l; k++) g += a[j][k] * a[k][i];
          e[j] = g / h;
          f += e[j] * a[j][i];
        }
        hh = f / (h + h);
        for (j = 0; j <= l; j++) {
          f = a[j][i];
          g = e[j] - hh * f;
          e[j] = g;
          for (k = 0; k <= j; k++) a[k][j] -= (f * e[k] + g * a[k][i]);
        }
      }
    } else {
      e[i] = a[l][i];
    }
    d[i] = h;
  }
  d[0] = 0.0;
  e[0] = 0.0;
  for (i = 0; i < (int)n; i++) {
    l = i - 1;
    if (d[i] != 0.0) {
      for (j = 0; j <= l; j++) {
        double g = 0.0;
        for (k = 0; k <= l; k++) g += a[k][i] * a[j][k];
        for (k = 0; k <= l; k++) a[j][k] -= g * a[i][k];
      }
    }
    d[i] = a[i][i];
    a[i][i] = 1.0;
    for (j = 0; j <= l; j++) a[i][j] = a[j][i] = 0.0;
  }
}

static void ql_implicit(double *d, double *e, unsigned n, double **z) { /* models.c:24-96 */
  int m, l, i, k, N = (int)n;
  double s, r, p, g, f, dd, c, b;
  for (i = 1; i < N; i++) e[i - 1] = e[i];
  e[N - 1] = 0.0;
  for (l = 0; l < N; l++) {
    do {
      for (m = l; m < N - 1; m++) {
        dd = fabs(d[m]) + fabs(d[m + 1]);
        if (fabs(e[m]) + dd == dd) break;
      }
      if (m != l) {
        g = (d[l + 1] - d[l]) / (2.0 * e[l]);
        r = sqrt((g * g) + 1.0);
        g = d[m] - d[l] + e[l] / (g + ((g < 0) ? -fabs(r) : fabs(r)));
        s = c = 1.0;
        p = 0.0;
        for (i = m - 1; i >= l; i--) {
          f = s * e[i];
          b = c * e[i];
          if (fabs(f) >= fabs(g)) {
            c = g / f;
            r = sqrt((c * c) + 1.0);
            e[i + 1] = f * r;
            c *= (s = 1.0 / r);
          } else {
            s = f / g;
            r = sqrt((s * s) + 1.0);
            e[i + 1] = g * r;
            s *= (c = 1.0 / r);
          }
          g = d[i + 1] - p;
          r = (d[i] - g) * s + 2.0 * c * b;
          p = s * r;
          d[i + 1] = g + p;
          g = c * r - b;
          for (k = 0; k < N; k++) {
            f = z[i + 1][k];
            z[i + 1][k] = s * z[i][k] + c * f;
            z[i][k] = c * z[i][k] - s * f;
          }
        }
        d[l] = d[l] - p;
        e[l] = g;
        e[m] = 0.0;
      }
    } while (m != l);
  }
}

int port_update_eigen(port_partition *p) {
  const unsigned states = p->states, sp = p->states_padded;
  const unsigned nparams = (states * states - states) / 2;
  unsigned i, j, k, inew, jnew, new_states = 0;
  double **a = (double **)xcalloc(states, sizeof(double *));
  double *pn = (double *)xcalloc(nparams, sizeof(double));
  double *d = (double *)xcalloc(states, sizeof(double));
  double *e = (double *)xcalloc(states, sizeof(double));
  double *nf = (double *)xcalloc(states, sizeof(double));
  double mean = 0;
  for (i = 0; i < states; ++i) a[i] = (double *)xcalloc(states, sizeof(double));

  /* create_ratematrix, models.c:182-262: sqrt(pi) Q sqrt(pi)^-1, normalised to mean rate 1 */
  memcpy(pn, p->subst_params, nparams * sizeof(double));
  if (pn[nparams - 1] > 0.0)
    for (i = 0; i < nparams; ++i) pn[i] /= pn[nparams - 1];
  k = 0;
  for (i = 0; i < states; ++i)
    for (j = i + 1; j < states; ++j) {
      double factor = (p->freqs[i] <= PORT_EIGEN_MINFREQ || p->freqs[j] <= PORT_EIGEN_MINFREQ) ? 0 : pn[k];
      k++;
      a[i][j] = a[j][i] = factor * sqrt(p->freqs[i] * p->freqs[j]);
      a[i][i] -= factor * p->freqs[j];
      a[j][j] -= factor * p->freqs[i];
    }
  for (i = 0; i < states; ++i) mean += p->freqs[i] * (-a[i][i]);
  for (i = 0; i < states; ++i)
    for (j = 0; j < states; ++j) a[i][j] /= mean;

  /* eliminate_zero_states, models.c:264-291 */
  for (i = 0; i < states; i++)
    if (p->freqs[i] > PORT_EIGEN_MINFREQ) nf[new_states++] = p->freqs[i];
  if (new_states < states) {
    for (i = 0, inew = 0; i < states; i++) {
      if (p->freqs[i] > PORT_EIGEN_MINFREQ) {
        for (j = 0, jnew = 0; j < states; j++)
          if (p->freqs[j] > PORT_EIGEN_MINFREQ) a[inew][jnew++] = a[i][j];
        inew++;
      }
    }
  }

  householder_tridiag(a, new_states, d, e);
  ql_implicit(d, e, new_states, a);

  for (i = 0, inew = 0; i < states; i++)
    p->eigenvals[i] = (p->freqs[i] > PORT_EIGEN_MINFREQ) ? d[inew++] : 0;
  for (i = 0; i < new_states; i++) nf[i] = sqrt(nf[i]);

  memset(p->eigenvecs, 0, (size_t)sp * states * sizeof(double));
  memset(p->inv_eigenvecs, 0, (size_t)sp * states * sizeof(double));
  if (new_states < states) { /* models.c:360-385 */
    for (i = 0; i < states; i++) p->eigenvecs[i * sp + i] = p->inv_eigenvecs[i * sp + i] = 1.;
    for (i = 0, inew = 0; i < states; i++) {
      if (p->freqs[i] > PORT_EIGEN_MINFREQ) {
        for (j = 0, jnew = 0; j < states; j++) {
          if (p->freqs[j] > PORT_EIGEN_MINFREQ) {
            p->eigenvecs[i * sp + j] = a[inew][jnew] * nf[jnew];
            p->inv_eigenvecs[i * sp + j] = a[jnew][inew] / nf[inew];
            jnew++;
          }
        }
        inew++;
      }
    }
  } else { /* models.c:387-398 */
    for (i = 0; i < states; i++)
      for (j = 0; j < states; j++) {
        p->eigenvecs[i * sp + j] = a[i][j] * nf[j];
        p->inv_eigenvecs[i * sp + j] = a[j][i] / nf[i];
      }
  }
  for (i = 0; i < states; ++i) free(a[i]);
  free(a); free(pn); free(d); free(e); free(nf);
  return 1;
}

/* pairwise tree sum used by every AVX 4x4 body (unpackhi/unpacklo add, permute2f128/blend add):
 * lanes end up as (p1+p0)+(p3+p2); IEEE addition is commutative, so (p0+p1)+(p2+p3). */
static inline double tree4(double p0, double p1, double p2, double p3) { return (p0 + p1) + (p2 + p3); }

int port_update_pmatrix(port_partition *p, unsigned edge, double t) {
  const unsigned states = p->states, sp = p->states_padded;
  unsigned n, j, k, m;
  double *expd = (double *)xcalloc(states, sizeof(double));
  double *temp = (double *)xcalloc((size_t)states * states, sizeof(double));
  if (t < 0) { free(expd); free(temp); return 0; }
  for (n = 0; n < p->rate_cats; ++n) {
    double *pmat = p->pmatrix[edge] + (size_t)n * states * sp;
    /* core_pmatrix.c:182-185: the category's matrix through params_indices */
    const double *eigenvals = p->m_eigenvals[p->cat_model[n]], *eigenvecs = p->m_eigenvecs[p->cat_model[n]],
                 *inv_eigenvecs = p->m_inv_eigenvecs[p->cat_model[n]];
    if (t > 0.) {
      /* core_pmatrix_avx.c:97-125: (eval*rate)*t, divided by (1-pinv) only if pinv > eps */
      for (j = 0; j < states; ++j) {
        double x = (eigenvals[j] * p->rates[n]) * t;
        if (p->prop_invar > PORT_MISC_EPSILON) x = x / (1.0 - p->prop_invar);
        expd[j] = expm1(x);
      }
      for (j = 0; j < states; ++j)
        for (k = 0; k < states; ++k) temp[j * states + k] = inv_eigenvecs[j * sp + k] * expd[k];
      if (states == 4) { /* core_pmatrix_avx.c:127-258: tree sum, identity added last */
        for (j = 0; j < 4; ++j)
          for (k = 0; k < 4; ++k) {
            double s = tree4(temp[j * 4 + 0] * eigenvecs[0 * sp + k], temp[j * 4 + 1] * eigenvecs[1 * sp + k],
                             temp[j * 4 + 2] * eigenvecs[2 * sp + k], temp[j * 4 + 3] * eigenvecs[3 * sp + k]);
            pmat[j * sp + k] = s + ((j == k) ? 1.0 : 0.0);
          }
      } else { /* core_pmatrix.c:205-217: identity first, then serial accumulation */
        for (j = 0; j < states; ++j)
          for (k = 0; k < states; ++k) {
            double s = (j == k) ? 1.0 : 0;
            for (m = 0; m < states; ++m) s += temp[j * states + m] * eigenvecs[m * sp + k];
            pmat[j * sp + k] = s;
          }
      }
    } else { /* core_pmatrix.c:220-226 */
      for (j = 0; j < states; ++j)
        for (k = 0; k < states; ++k) pmat[j * sp + k] = (j == k) ? 1 : 0;
    }
  }
  free(expd); free(temp);
  return 1;
}

/* Σ_{j∈mask} M[row][j] in the order the reference uses:
 * 4 states: masked load + tree sum (core_partials_avx.c:296-350, :1372-1400);
 * else: serial over set bits (core_partials.c:421-440). */
static inline double masked_rowsum(const double *row, unsigned states, uint32_t mask) {
  if (states == 4)
    return tree4((mask & 1) ? row[0] : 0.0, (mask & 2) ? row[1] : 0.0, (mask & 4) ? row[2] : 0.0,
                 (mask & 8) ? row[3] : 0.0);
  double s = 0;
  for (unsigned j = 0; j < states; ++j)
    if ((mask >> j) & 1) s += row[j];
  return s;
}

static inline double row_dot(const double *row, const double *clv, unsigned states) {
  if (states == 4) return tree4(row[0] * clv[0], row[1] * clv[1], row[2] * clv[2], row[3] * clv[3]);
  double s = 0;
  for (unsigned j = 0; j < states; ++j) s += row[j] * clv[j];
  return s;
}

void port_update_partials(const port_partition *p, double *parent_clv, unsigned *parent_scaler,
                          const port_operand *left, const port_operand *right) {
  const unsigned states = p->states, sp = p->states_padded, cats = p->rate_cats;
  const unsigned span = sp * cats;
  const int tiptip = (left->kind == 1 && right->kind == 1);
  unsigned n, k, i;
  for (n = 0; n < p->sites; ++n) {
    double *par = parent_clv + (size_t)n * span;
    int all_small = 1;
    for (k = 0; k < cats; ++k) {
      for (i = 0; i < states; ++i) {
        double x, y;
        /* fake operand: all-ones CLV times identity matrix == exactly 1.0 (RaxmlWrapper.cpp:156-187) */
        if (left->kind == 2) x = 1.0;
        else {
          const double *lrow = p->pmatrix[left->edge] + ((size_t)k * states + i) * sp;
          x = (left->kind == 1) ? masked_rowsum(lrow, states, p->tipmap[p->tipchars[left->tip][n]])
                                : row_dot(lrow, left->clv + (size_t)n * span + k * sp, states);
        }
        if (right->kind == 2) y = 1.0;
        else {
          const double *rrow = p->pmatrix[right->edge] + ((size_t)k * states + i) * sp;
          y = (right->kind == 1) ? masked_rowsum(rrow, states, p->tipmap[p->tipchars[right->tip][n]])
                                 : row_dot(rrow, right->clv + (size_t)n * span + k * sp, states);
        }
        par[k * sp + i] = x * y;
        all_small &= (par[k * sp + i] < PORT_SCALE_THRESHOLD);
      }
      for (i = states; i < sp; ++i) par[k * sp + i] = 0.0;
    }
    if (parent_scaler) {
      /* pll_fill_parent_scaler (core_partials.c) + per-site scaling (core_partials_avx.c:548-563);
       * tip-tip: scaler := 0 and NO scaling test (core_partials_avx.c:1003-1009) */
      unsigned s = 0;
      if (!tiptip) {
        if (left->kind == 0 && left->scaler) s += left->scaler[n];
        if (right->kind == 0 && right->scaler) s += right->scaler[n];
        if (all_small) {
          for (i = 0; i < span; ++i) par[i] *= PORT_SCALE_FACTOR;
          s += 1;
        }
      }
      parent_scaler[n] = s;
    }
  }
}

double port_root_loglikelihood(const port_partition *p, const double *clv, const unsigned *scaler,
                               double *persite_lnl) {
  const unsigned states = p->states, sp = p->states_padded, cats = p->rate_cats;
  double logl = 0;
  unsigned n, j, k;
  for (n = 0; n < p->sites; ++n) {
    double term = 0;
    for (j = 0; j < cats; ++j) {
      const double *c = clv + ((size_t)n * cats + j) * sp;
      const double *freqs = p->m_freqs[p->cat_model[j]]; /* frequencies[freqs_indices[j]] */
      double term_r;
      if (states == 4) /* core_likelihood_avx.c:232-245: mul, hadd, [0]+[2] */
        term_r = tree4(freqs[0] * c[0], freqs[1] * c[1], freqs[2] * c[2], freqs[3] * c[3]);
      else { /* core_likelihood.c:163-171 */
        term_r = 0;
        for (k = 0; k < states; ++k) term_r += c[k] * freqs[k];
      }
      if (p->prop_invar > 0) { /* core_likelihood.c:174-186 (the AVX / AVX2 kernels have the same lines) */
        const double inv_site_lk = (p->invariant[n] == -1) ? 0 : freqs[p->invariant[n]];
        term += p->rate_weights[j] * (term_r * (1 - p->prop_invar) + inv_site_lk * p->prop_invar);
      } else
        term += term_r * p->rate_weights[j];
    }
    term = log(term);
    if (scaler && scaler[n]) term += scaler[n] * log(PORT_SCALE_THRESHOLD);
    term *= p->pattern_weights[n];
    if (persite_lnl) persite_lnl[n] = term;
    logl += term;
  }
  return logl;
}

double port_edge_loglikelihood(const port_partition *p, const port_operand *parent,
                               const port_operand *child, unsigned edge, double *persite_lnl) {
  /* tip/inner dispatch as LIBPLL/likelihood.c:555-615: the tip, if any, plays "child" */
  const unsigned states = p->states, sp = p->states_padded, cats = p->rate_cats;
  const port_operand *in = parent, *ot = child;
  double logl = 0;
  unsigned n, i, j;
  if (parent->kind == 1) { in = child; ot = parent; }
  if (in->kind != 0) return -INFINITY; /* tip-tip: invalid */
  for (n = 0; n < p->sites; ++n) {
    double terma = 0, terminv = 0, site_lk;
    unsigned site_scalings = 0;
    if (in->scaler) site_scalings += in->scaler[n];
    if (ot->kind == 0 && ot->scaler) site_scalings += ot->scaler[n];
    for (i = 0; i < cats; ++i) {
      const double *clvp = in->clv + ((size_t)n * cats + i) * sp;
      const double *pm = p->pmatrix[edge] + (size_t)i * states * sp;
      const double *freqs = p->m_freqs[p->cat_model[i]];
      double terma_r = 0;
      for (j = 0; j < states; ++j) {
        double termb = (ot->kind == 1)
                           ? masked_rowsum(pm + j * sp, states, p->tipmap[p->tipchars[ot->tip][n]])
                           : row_dot(pm + j * sp, ot->clv + ((size_t)n * cats + i) * sp, states);
        terma_r += clvp[j] * freqs[j] * termb; /* core_likelihood.c:1433 */
      }
      if (p->prop_invar > 0) { /* core_likelihood_avx.c:471-486, core_likelihood_avx2.c:391-406 */
        terma += p->rate_weights[i] * terma_r * (1. - p->prop_invar);
        if (p->invariant[n] != -1) terminv += p->rate_weights[i] * freqs[p->invariant[n]] * p->prop_invar;
      } else
        terma += terma_r * p->rate_weights[i];
    }
    if (site_scalings) {
      if (terminv > 0.) { /* "undoing the scaling for non-variant likelihood term only" (core_likelihood_avx.c:493-501) */
        const unsigned capped = site_scalings < 4 ? site_scalings : 4; /* PLL_SCALE_RATE_MAXDIFF */
        double scale_factor = 1.0;
        unsigned s_;
        for (s_ = 0; s_ < capped; ++s_) scale_factor *= PORT_SCALE_THRESHOLD; /* scale_minlh[capped - 1] (:323-333) */
        site_lk = log(terma * scale_factor + terminv);
      } else {
        site_lk = log(terma);
        site_lk += site_scalings * log(PORT_SCALE_THRESHOLD);
      }
    } else
      site_lk = log(terma + terminv);
    site_lk *= p->pattern_weights[n];
    if (persite_lnl) persite_lnl[n] = site_lk;
    logl += site_lk;
  }
  return logl;
}

int port_update_sumtable(const port_partition *p, const port_operand *parent,
                         const port_operand *child, double *sumtable) {
  /* LIBPLL/derivatives.c:24-164: for tip-inner the TIP is always the "left/parent" operand of the
   * core kernel (lefterm uses freqs * inv_eigenvecs, righterm uses eigenvecs * inner clv). */
  const unsigned states = p->states, sp = p->states_padded, cats = p->rate_cats;
  const port_operand *l = parent, *r = child;
  unsigned n, i, j, k;
  if (parent->kind == 1 && child->kind == 1) return 0;
  if (child->kind == 1) { l = child; r = parent; }
  for (n = 0; n < p->sites; ++n) {
    for (i = 0; i < cats; ++i) {
      const double *cr = r->clv + ((size_t)n * cats + i) * sp;
      double *sum = sumtable + ((size_t)n * cats + i) * sp;
      /* core_derivatives.c:362-366: eigenvectors / frequencies of the category's matrix */
      const double *freqs = p->m_freqs[p->cat_model[i]], *eigenvecs = p->m_eigenvecs[p->cat_model[i]],
                   *inv_eigenvecs = p->m_inv_eigenvecs[p->cat_model[i]];
      for (j = 0; j < states; ++j) {
        double lefterm = 0, righterm = 0;
        if (l->kind == 1) { /* core_derivatives.c:616-627 */
          uint32_t ts = p->tipmap[p->tipchars[l->tip][n]];
          for (k = 0; k < states; ++k) {
            lefterm += (double)(ts & 1) * freqs[k] * inv_eigenvecs[k * sp + j];
            righterm += eigenvecs[j * sp + k] * cr[k];
            ts >>= 1;
          }
        } else { /* core_derivatives.c:447-456 */
          const double *cl = l->clv + ((size_t)n * cats + i) * sp;
          for (k = 0; k < states; ++k) {
            lefterm += cl[k] * freqs[k] * inv_eigenvecs[k * sp + j];
            righterm += eigenvecs[j * sp + k] * cr[k];
          }
        }
        sum[j] = lefterm * righterm;
      }
      for (j = states; j < sp; ++j) sum[j] = 0.0;
    }
  }
  return 1;
}

void port_compute_diagptable(const port_partition *p, double t, double *diagp) {
  /* core_derivatives.c:711-726 */
  unsigned i, j;
  for (i = 0; i < p->rate_cats; ++i) {
    double ki = p->rates[i] / (1.0 - p->prop_invar);
    const double *eigenvals = p->m_eigenvals[p->cat_model[i]]; /* core_derivatives.c:712 */
    for (j = 0; j < p->states; ++j) {
      diagp[0] = exp(eigenvals[j] * ki * t);
      diagp[1] = eigenvals[j] * ki * diagp[0];
      diagp[2] = eigenvals[j] * ki * eigenvals[j] * ki * diagp[0];
      diagp[3] = 0;
      diagp += 4;
    }
  }
}

int port_loglikelihood_derivatives(const port_partition *p, const double *sumtable,
                                   const double *diagptable, double *f, double *d_f, double *dd_f) {
  /* core_derivatives.c:643-694 (per-site lk0,lk1,lk2) and :840-867 (accumulation).
   * f follows the AVX2 kernel the reference runs (core_derivatives_avx2.c:1788-1802,1849-1874):
   * Σ w_n·log(lk0_n), no scaler term (Q1). */
  const unsigned states = p->states, sp = p->states_padded, cats = p->rate_cats;
  unsigned n, i, j;
  double F = 0, D1 = 0, D2 = 0;
  for (n = 0; n < p->sites; ++n) {
    const double *sum = sumtable + (size_t)n * cats * sp;
    const double *diagp = diagptable;
    double lk[3] = {0, 0, 0};
    for (i = 0; i < cats; ++i) {
      double c0 = 0, c1 = 0, c2 = 0;
      for (j = 0; j < states; ++j) {
        c0 += sum[j] * diagp[0];
        c1 += sum[j] * diagp[1];
        c2 += sum[j] * diagp[2];
        diagp += 4;
      }
      if (p->prop_invar > 0) { /* core_derivatives_avx2.c:1736-1749 (generic: core_derivatives.c:672-684) */
        c0 *= 1. - p->prop_invar; c1 *= 1. - p->prop_invar; c2 *= 1. - p->prop_invar;
        if (p->invariant && p->invariant[n] != -1) c0 += p->m_freqs[p->cat_model[i]][p->invariant[n]] * p->prop_invar;
      }
      lk[0] += c0 * p->rate_weights[i];
      lk[1] += c1 * p->rate_weights[i];
      lk[2] += c2 * p->rate_weights[i];
      sum += sp;
    }
    {
      double deriv1 = -lk[1] / lk[0];
      double deriv2 = deriv1 * deriv1 - lk[2] / lk[0];
      F += p->pattern_weights[n] * log(lk[0]);
      D1 += p->pattern_weights[n] * deriv1;
      D2 += p->pattern_weights[n] * deriv2;
    }
  }
  if (f) *f = F;
  *d_f = D1;
  *dd_f = D2;
  return 1;
}
