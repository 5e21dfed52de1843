/*
 * backend_port.cpp — ORACLE (test infrastructure, NOT product code).
 * orc::Backend over the scalar restatement in pll_port.c (cpu_baseline.kind == "port").
 */
#include "netrax_port.hpp"
#include "pll_port.h"

#include <cmath>

#include <map>

namespace orc {
namespace {
struct PortBackend : Backend {
  std::vector<port_partition *> parts;
  ~PortBackend() override { for (auto *p : parts) port_partition_destroy(p); }
  const char *kind() const override { return "port"; }
  unsigned partitionCount() const override { return (unsigned)parts.size(); }
  unsigned sites(unsigned p) const override { return parts[p]->sites; }
  size_t clvEntries(unsigned p) const override { return (size_t)parts[p]->sites * parts[p]->rate_cats * parts[p]->states_padded; }
  unsigned statesPadded(unsigned p) const override { return parts[p]->states_padded; }
  unsigned rateCats(unsigned p) const override { return parts[p]->rate_cats; }
  unsigned states(unsigned p) const override { return parts[p]->states; }
  void setModel(unsigned p, const double *freqs, const double *subst, const double *rates, const double *weights) override {
    port_partition *pp = parts[p];
    double sum = 0.;   // pll_set_frequencies (LIBPLL/models.c:445-467): renormalise when |sum - 1| > PLL_MISC_EPSILON
    for (unsigned i = 0; i < pp->states; ++i) { pp->freqs[i] = freqs[i]; sum += freqs[i]; }
    if (std::fabs(sum - 1.0) > 1e-8) for (unsigned i = 0; i < pp->states; ++i) pp->freqs[i] /= sum;
    for (unsigned i = 0; i < pp->states * (pp->states - 1) / 2; ++i) pp->subst_params[i] = subst[i];
    for (unsigned i = 0; i < pp->rate_cats; ++i) { pp->rates[i] = rates[i]; pp->rate_weights[i] = weights[i]; }
    port_update_eigen(pp);
  }
  void getEigen(unsigned p, double *ev, double *iev, double *evals) const override {
    const port_partition *pp = parts[p];
    std::memcpy(ev, pp->eigenvecs, sizeof(double) * pp->states * pp->states_padded);
    std::memcpy(iev, pp->inv_eigenvecs, sizeof(double) * pp->states * pp->states_padded);
    std::memcpy(evals, pp->eigenvals, sizeof(double) * pp->states_padded);
  }
  void getRates(unsigned p, double *rates, double *weights, double *freqs) const override {
    const port_partition *pp = parts[p];
    std::memcpy(rates, pp->rates, sizeof(double) * pp->rate_cats);
    std::memcpy(weights, pp->rate_weights, sizeof(double) * pp->rate_cats);
    std::memcpy(freqs, pp->freqs, sizeof(double) * pp->states_padded);
  }
  void setSubmodels(unsigned p, unsigned n, const unsigned *cat_model, const double *freqs, const double *subst) override {
    if (!port_set_submodels(parts[p], n, cat_model, freqs, subst)) throw std::runtime_error("set_submodels: 1..16 rate matrices, every category's index below the matrix count");
  }
  void setPinv(unsigned p, double pinv) override {   // pll_update_invariant_sites_proportion (LIBPLL/models.c:495-543)
    if (!port_set_prop_invar(parts[p], pinv)) throw std::runtime_error("Invalid proportion of invariant sites");
  }
  void setCategoryRates(unsigned p, const double *rates) override { for (unsigned i = 0; i < parts[p]->rate_cats; ++i) parts[p]->rates[i] = rates[i]; }
  bool gammaRates(double alpha, unsigned cats, double *out, int mode) const override { return port_compute_gamma_cats(alpha, cats, out, mode) != 0; }
  void updatePmatrix(unsigned p, unsigned edge, double brlen) override { port_update_pmatrix(parts[p], edge, brlen); }
  const double *pmatrix(unsigned p, unsigned edge) const override { return parts[p]->pmatrix[edge]; }
  static port_operand conv(const Operand &o) {
    port_operand r;
    r.kind = o.kind; r.clv = o.clv; r.scaler = o.scaler; r.tip = o.tip; r.edge = o.edge;
    return r;
  }
  void updatePartials(unsigned p, double *pc, unsigned *ps, const Operand &l, const Operand &r) override {
    port_operand a = conv(l), b = conv(r);
    port_update_partials(parts[p], pc, ps, &a, &b);
  }
  double rootLogl(unsigned p, const double *clv, const unsigned *scaler, double *persite) override {
    return port_root_loglikelihood(parts[p], clv, scaler, persite);
  }
  double edgeLogl(unsigned p, const Operand &parent, const Operand &child, unsigned edge, double *persite) override {
    port_operand a = conv(parent), b = conv(child);
    return port_edge_loglikelihood(parts[p], &a, &b, edge, persite);
  }
  void sumtable(unsigned p, const Operand &parent, const Operand &child, double *out) override {
    port_operand a = conv(parent), b = conv(child);
    if (!port_update_sumtable(parts[p], &a, &b, out)) throw std::runtime_error("pll_update_sumtable() was called for the tip-tip case!");
  }
  void derivatives(unsigned p, const double *st, double brlen, bool want_f, double *f, double *d1, double *d2) override {
    std::vector<double> diag((size_t)parts[p]->rate_cats * parts[p]->states * 4);
    port_compute_diagptable(parts[p], brlen, diag.data());
    port_loglikelihood_derivatives(parts[p], st, diag.data(), want_f ? f : nullptr, d1, d2);
  }
};
}  // namespace

Backend *makePortBackend(unsigned tips, unsigned edges_plus_fake, const std::vector<PartitionDesc> &descs) {
  PortBackend *b = new PortBackend();
  for (const PartitionDesc &d : descs) {
    port_partition *pp = port_partition_create(d.states, d.rate_cats, d.sites, tips, edges_plus_fake);
    for (unsigned i = 0; i < d.sites; ++i) pp->pattern_weights[i] = d.pattern_weights.empty() ? 1 : d.pattern_weights[i];
    if (d.states == 4) {
      for (unsigned t = 0; t < tips; ++t)
        for (unsigned s = 0; s < d.sites; ++s) pp->tipchars[t][s] = (unsigned char)d.tip_masks[t][s];
    } else {  // code table in order of first appearance (role of pll charmap/tipmap, LIBPLL/pll.c:903-957)
      std::map<uint32_t, unsigned> code;
      for (unsigned t = 0; t < tips; ++t)
        for (unsigned s = 0; s < d.sites; ++s) {
          uint32_t m = d.tip_masks[t][s];
          auto it = code.find(m);
          if (it == code.end()) {
            if (code.size() >= 256) throw std::runtime_error("too many distinct tip states");
            unsigned c = (unsigned)code.size();
            it = code.emplace(m, c).first;
            pp->tipmap[c] = m;
          }
          pp->tipchars[t][s] = (unsigned char)it->second;
        }
      pp->maxstates = (unsigned)code.size();
    }
    b->parts.push_back(pp);
    b->setModel((unsigned)b->parts.size() - 1, d.freqs.data(), d.subst_params.data(), d.rates.data(), d.rate_weights.data());
  }
  return b;
}

}  // namespace orc
