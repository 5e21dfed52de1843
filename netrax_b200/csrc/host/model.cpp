/*
 * model.cpp — host-side model helpers at the boundary (tiny, stay on the CPU):
 *   update_eigen       role of pll_update_eigen (LIBPLL/models.c:293-410): eigen-decomposition of the
 *                      symmetrised, mean-rate-normalised rate matrix.  The reference uses Householder + QL;
 *                      here a cyclic Jacobi solver (different algorithm, same decomposition up to sign /
 *                      order of eigenvectors, which cancels in P(t) = I + V^-1 diag(expm1) V).
 *   compute_gamma_cats role of pll_compute_gamma_cats (LIBPLL/gamma.c:267-330): discrete Gamma rates from the
 *                      published algorithms AS 91 (Best & Roberts 1975), AS 32 (Bhattacharjee 1970),
 *                      AS 70 (Odeh & Evans 1974) and Algorithm 291 (Pike & Hill 1966) with their published
 *                      truncation constants — these constants are visible in the rates at the 1e-7 level, so a
 *                      "more exact" quantile routine would break 1e-10 lnL parity with the reference.
 */
#include <cmath>

#include "netrax_likelihood_api.hpp"

namespace netrax {

/* pll_set_frequencies (LIBPLL/models.c:445-467): copy, and make sure the frequencies sum up to 1 — libpll's own empirical
 * amino-acid tables (pll_aa_freqs_*) sum to 1 + 1e-6, so this is not a corner case: without it every protein lnL is off by
 * ~1e-7 relative (found by pinning against test/out/protein-models.out). */
void set_frequencies(PartitionModel &m, const double *frequencies) {
  m.frequencies.assign(frequencies, frequencies + m.states);
  double sum = 0.;
  for (unsigned i = 0; i < m.states; ++i) sum += m.frequencies[i];
  if (std::fabs(sum - 1.0) > 1e-8 /* PLL_MISC_EPSILON */)
    for (unsigned i = 0; i < m.states; ++i) m.frequencies[i] /= sum;
  m.eigen_decomp_valid = false;
}

void update_eigen(PartitionModel &m) {
  const unsigned n = m.states, sp = m.states_padded;
  const unsigned nparams = n * (n - 1) / 2;
  if (m.frequencies.size() < n || m.subst_params.size() < nparams) throw std::runtime_error("update_eigen: model not set");
  for (unsigned i = 0; i < n; ++i)
    if (!(m.frequencies[i] > 1e-6)) throw std::runtime_error("update_eigen: state frequencies <= 1e-6 are not supported");
  std::vector<double> rate(m.subst_params.begin(), m.subst_params.begin() + nparams);
  if (rate[nparams - 1] > 0.0) { const double last = rate[nparams - 1]; for (double &r : rate) r /= last; }
  // A = sqrt(pi) Q sqrt(pi)^-1 (symmetric), Q normalised to one expected substitution per unit time
  std::vector<double> A((size_t)n * n, 0.0), V((size_t)n * n, 0.0);
  unsigned k = 0;
  for (unsigned i = 0; i < n; ++i)
    for (unsigned j = i + 1; j < n; ++j, ++k) {
      A[i * n + j] = A[j * n + i] = rate[k] * std::sqrt(m.frequencies[i] * m.frequencies[j]);
      A[i * n + i] -= rate[k] * m.frequencies[j];
      A[j * n + j] -= rate[k] * m.frequencies[i];
    }
  double mean = 0;
  for (unsigned i = 0; i < n; ++i) mean += m.frequencies[i] * (-A[i * n + i]);
  for (double &a : A) a /= mean;
  for (unsigned i = 0; i < n; ++i) V[i * n + i] = 1.0;
  // cyclic Jacobi: columns of V become the eigenvectors
  for (int sweep = 0; sweep < 100; ++sweep) {
    double off = 0;
    for (unsigned p = 0; p < n; ++p) for (unsigned q = p + 1; q < n; ++q) off += A[p * n + q] * A[p * n + q];
    if (off < 1e-300) break;
    for (unsigned p = 0; p < n; ++p)
      for (unsigned q = p + 1; q < n; ++q) {
        const double apq = A[p * n + q];
        if (std::fabs(apq) < 1e-300) continue;
        const double theta = (A[q * n + q] - A[p * n + p]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (unsigned r = 0; r < n; ++r) {  // A <- A J
          const double arp = A[r * n + p], arq = A[r * n + q];
          A[r * n + p] = c * arp - s * arq;
          A[r * n + q] = s * arp + c * arq;
        }
        for (unsigned r = 0; r < n; ++r) {  // A <- J^T A
          const double apr = A[p * n + r], aqr = A[q * n + r];
          A[p * n + r] = c * apr - s * aqr;
          A[q * n + r] = s * apr + c * aqr;
        }
        for (unsigned r = 0; r < n; ++r) {
          const double vrp = V[r * n + p], vrq = V[r * n + q];
          V[r * n + p] = c * vrp - s * vrq;
          V[r * n + q] = s * vrp + c * vrq;
        }
      }
  }
  m.eigenvals.assign(sp, 0.0);
  m.eigenvecs.assign((size_t)n * sp, 0.0);
  m.inv_eigenvecs.assign((size_t)n * sp, 0.0);
  for (unsigned e = 0; e < n; ++e) {
    m.eigenvals[e] = A[e * n + e];
    for (unsigned j = 0; j < n; ++j) {
      m.eigenvecs[e * sp + j] = V[j * n + e] * std::sqrt(m.frequencies[j]);      // row e: eigenvector e times sqrt(pi)
      m.inv_eigenvecs[j * sp + e] = V[j * n + e] / std::sqrt(m.frequencies[j]);  // column e
    }
  }
  m.eigen_decomp_valid = true;
}

void set_submodels(PartitionModel &m, unsigned n, const unsigned *ratecat_submodels, const double *freqs, const double *subst) {
  const unsigned S = m.states, NR = S * (S - 1) / 2;
  if (n < 1 || n > 16) throw std::runtime_error("set_submodels: 1..16 rate matrices");
  for (unsigned c = 0; n > 1 && c < m.rate_cats; ++c)
    if (ratecat_submodels[c] >= n) throw std::runtime_error("set_submodels: rate-matrix index of a category out of range");
  set_frequencies(m, freqs);
  m.subst_params.assign(subst, subst + NR);
  m.submodels.clear();
  m.ratecat_submodels.clear();
  for (unsigned i = 1; i < n; ++i) {
    PartitionModel tmp;   // pll_set_frequencies / pll_set_subst_params / pll_update_eigen with params_index i
    tmp.states = S; tmp.states_padded = m.states_padded;
    set_frequencies(tmp, freqs + (size_t)i * S);
    tmp.subst_params.assign(subst + (size_t)i * NR, subst + (size_t)(i + 1) * NR);
    update_eigen(tmp);
    m.submodels.push_back({tmp.frequencies, tmp.subst_params, tmp.eigenvecs, tmp.inv_eigenvecs, tmp.eigenvals});
  }
  if (n > 1) m.ratecat_submodels.assign(ratecat_submodels, ratecat_submodels + m.rate_cats);
  m.eigen_decomp_valid = false;
}

namespace {
double lnGamma(double a) {  // Algorithm 291
  double x = a, f = 0.0;
  if (x < 7.0) {
    f = 1.0;
    double z = a;
    while (z < 7.0) { f *= z; z += 1.0; }
    x = z;
    f = -std::log(f);
  }
  const double z = 1.0 / (x * x);
  return f + (x - 0.5) * std::log(x) - x + .918938533204673 +
         (((-.000595238095238 * z + .000793650793651) * z - .002777777777778) * z + .083333333333333) / x;
}

double incompleteGammaRatio(double x, double alpha, double lnga) {  // AS 32
  const double eps = 1e-8, big = 1e30;
  if (x == 0) return 0;
  if (x < 0 || alpha <= 0) return -1;
  const double factor = std::exp(alpha * std::log(x) - x - lnga);
  if (!(x > 1 && x >= alpha)) {
    double gin = 1, term = 1, rn = alpha;
    do { rn += 1; term *= x / rn; gin += term; } while (term > eps);
    return gin * factor / alpha;
  }
  double a = 1 - alpha, b = a + x + 1, term = 0, pn[6] = {1, x, x + 1, x * b, 0, 0};
  double gin = pn[2] / pn[3];
  for (;;) {
    a += 1; b += 2; term += 1;
    const double an = a * term;
    pn[4] = b * pn[2] - an * pn[0];
    pn[5] = b * pn[3] - an * pn[1];
    if (pn[5] != 0) {
      const double rn = pn[4] / pn[5], dif = std::fabs(gin - rn);
      if (dif <= eps && dif <= eps * rn) return 1 - factor * gin;
      gin = rn;
    }
    for (int i = 0; i < 4; ++i) pn[i] = pn[i + 2];
    if (std::fabs(pn[4]) >= big) for (int i = 0; i < 4; ++i) pn[i] /= big;
  }
}

double normalQuantile(double p) {  // AS 70
  const double a[5] = {-.322232431088, -1, -.342242088547, -.0204231210245, -.453642210148e-4};
  const double b[5] = {.0993484626060, .588581570495, .531103462366, .103537752850, .0038560700634};
  const double p1 = p < 0.5 ? p : 1 - p;
  if (p1 < 1e-20) return -9999;
  const double y = std::sqrt(std::log(1 / (p1 * p1)));
  const double z = y + ((((y * a[4] + a[3]) * y + a[2]) * y + a[1]) * y + a[0]) / ((((y * b[4] + b[3]) * y + b[2]) * y + b[1]) * y + b[0]);
  return p < 0.5 ? -z : z;
}

double chi2Quantile(double p, double v) {  // AS 91
  const double e = .5e-6, aa = .6931471805;
  if (p < .000002 || p > .999998 || v <= 0) return -1;
  const double g = lnGamma(v / 2), xx = v / 2, c = xx - 1;
  double ch;
  if (v < -1.24 * std::log(p)) {
    ch = std::pow(p * xx * std::exp(g + xx * aa), 1 / xx);
    if (ch - e < 0) return ch;
  } else if (v > .32) {
    const double x = normalQuantile(p), p1 = 0.222222 / v;
    ch = v * std::pow(x * std::sqrt(p1) + 1 - p1, 3.0);
    if (ch > 2.2 * v + 6) ch = -2 * (std::log(1 - p) - c * std::log(.5 * ch) + g);
  } else {
    ch = 0.4;
    const double a = std::log(1 - p);
    double q;
    do {
      q = ch;
      const double p1 = 1 + ch * (4.67 + ch), p2 = ch * (6.73 + ch * (6.66 + ch));
      const double t = -0.5 + (4.67 + 2 * ch) / p1 - (6.73 + ch * (13.32 + 3 * ch)) / p2;
      ch -= (1 - std::exp(a + g + .5 * ch + c * aa) * p2 / p1) / t;
    } while (std::fabs(q / ch - 1) - .01 > 0);
  }
  double q;
  do {
    q = ch;
    const double p1 = .5 * ch;
    double t = incompleteGammaRatio(p1, xx, g);
    if (t < 0) return -1;
    const double p2 = p - t;
    t = p2 * std::exp(xx * aa + g + p1 - c * std::log(ch));
    const double b = t / ch, a = 0.5 * t - b * c;
    const double s1 = (210 + a * (140 + a * (105 + a * (84 + a * (70 + 60 * a))))) / 420;
    const double s2 = (420 + a * (735 + a * (966 + a * (1141 + 1278 * a)))) / 2520;
    const double s3 = (210 + a * (462 + a * (707 + 932 * a))) / 2520;
    const double s4 = (252 + a * (672 + 1182 * a) + c * (294 + a * (889 + 1740 * a))) / 5040;
    const double s5 = (84 + 264 * a + c * (175 + 606 * a)) / 2520;
    const double s6 = (120 + c * (346 + 127 * c)) / 5040;
    ch += t * (1 + 0.5 * t * s1 - b * c * (s1 - b * (s2 - b * (s3 - b * (s4 - b * (s5 - b * s6))))));
  } while (std::fabs(q / ch - 1) > e);
  return ch;
}
}  // namespace

bool compute_gamma_cats(double alpha, unsigned cats, double *out, int mode) {
  if (alpha < 0.02 || cats < 1) return false;
  if (cats == 1) { out[0] = 1.0; return true; }
  const double factor = alpha / alpha * cats, beta = alpha;
  if (mode == 1) {  // PLL_GAMMA_RATES_MEDIAN
    double t = 0;
    for (unsigned i = 0; i < cats; ++i) { out[i] = chi2Quantile((double)(2 * i + 1) / (2.0 * cats), 2 * alpha) / (2 * beta); t += out[i]; }
    for (unsigned i = 0; i < cats; ++i) out[i] *= factor / t;
    return true;
  }
  if (mode != 0) return false;
  std::vector<double> g(cats);  // PLL_GAMMA_RATES_MEAN
  const double lnga1 = lnGamma(alpha + 1);
  for (unsigned i = 0; i + 1 < cats; ++i) g[i] = incompleteGammaRatio(chi2Quantile((i + 1.0) / cats, 2 * alpha) / (2 * beta) * beta, alpha + 1, lnga1);
  out[0] = g[0] * factor;
  out[cats - 1] = (1 - g[cats - 2]) * factor;
  for (unsigned i = 1; i + 1 < cats; ++i) out[i] = (g[i] - g[i - 1]) * factor;
  return true;
}

}  // namespace netrax
