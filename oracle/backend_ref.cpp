/*
 * backend_ref.cpp — ORACLE (test infrastructure, NOT product code).  Built only into oracle/_ref/.
 *
 * orc::Backend whose every arithmetic call is the REAL forked libpll of the reference, compiled
 * unmodified from /root/reference (oracle/Makefile target `ref`): pll_update_prob_matrices,
 * pll_update_partials_single, pll_compute_root_loglikelihood, pll_compute_edge_loglikelihood,
 * pll_update_sumtable, pll_compute_diagptable + pll_compute_loglikelihood_derivatives — the
 * explicit-pointer entry points NetRAX calls (LIBPLL/pll.h:731-842).  Partition attributes are
 * the ones NetRAX/raxml-ng use in every BASELINE config: AVX2 + PATTERN_TIP, per-site scalers,
 * no site repeats.  cpu_baseline.kind == "reference".
 */
#include "netrax_port.hpp"

extern "C" {
#include "pll.h"
}

#include <map>

namespace orc {
namespace {
struct RefBackend : Backend {
  std::vector<pll_partition_t *> parts;
  unsigned tips = 0, fake_clv = 0;
  std::vector<std::vector<unsigned>> pidx;   // per partition: params_indices (all zero unless setSubmodels installed a mixture)
  std::vector<unsigned> nmat;                // rate matrices in use per partition
  ~RefBackend() override { for (auto *p : parts) pll_partition_destroy(p); }
  const char *kind() const override { return "reference"; }
  unsigned partitionCount() const override { return (unsigned)parts.size(); }
  unsigned sites(unsigned p) const override { return parts[p]->sites; }
  size_t clvEntries(unsigned p) const override { return (size_t)parts[p]->sites * parts[p]->rate_cats * parts[p]->states_padded; }
  unsigned statesPadded(unsigned p) const override { return parts[p]->states_padded; }
  unsigned rateCats(unsigned p) const override { return parts[p]->rate_cats; }
  unsigned states(unsigned p) const override { return parts[p]->states; }
  void setModel(unsigned p, const double *freqs, const double *subst, const double *rates, const double *weights) override {
    pll_set_frequencies(parts[p], 0, freqs);
    pll_set_subst_params(parts[p], 0, subst);
    pll_set_category_rates(parts[p], rates);
    pll_set_category_weights(parts[p], weights);
    pll_update_eigen(parts[p], 0);
  }
  void getEigen(unsigned p, double *ev, double *iev, double *evals) const override {
    const pll_partition_t *pp = parts[p];
    std::memcpy(ev, pp->eigenvecs[0], sizeof(double) * pp->states * pp->states_padded);
    std::memcpy(iev, pp->inv_eigenvecs[0], sizeof(double) * pp->states * pp->states_padded);
    std::memcpy(evals, pp->eigenvals[0], sizeof(double) * pp->states_padded);
  }
  void getRates(unsigned p, double *rates, double *weights, double *freqs) const override {
    const pll_partition_t *pp = parts[p];
    std::memcpy(rates, pp->rates, sizeof(double) * pp->rate_cats);
    std::memcpy(weights, pp->rate_weights, sizeof(double) * pp->rate_cats);
    std::memcpy(freqs, pp->frequencies[0], sizeof(double) * pp->states_padded);
  }
  void setSubmodels(unsigned p, unsigned n, const unsigned *cat_model, const double *freqs, const double *subst) override {
    pll_partition_t *pp = parts[p];
    if (n < 1 || n > pp->rate_matrices) throw std::runtime_error("setSubmodels: number of rate matrices out of range");
    const unsigned nr = pp->states * (pp->states - 1) / 2;
    for (unsigned i = 0; i < n; ++i) {
      pll_set_frequencies(pp, i, freqs + (size_t)i * pp->states);
      pll_set_subst_params(pp, i, subst + (size_t)i * nr);
      pll_update_eigen(pp, i);
      if (i > 0 && !pll_update_invariant_sites_proportion(pp, i, pp->prop_invar[0])) throw std::runtime_error(pll_errmsg);
    }
    for (unsigned c = 0; c < pp->rate_cats; ++c) pidx[p][c] = n > 1 ? cat_model[c] : 0;
    nmat[p] = n;
  }
  void setPinv(unsigned p, double pinv) override {   // raxml-ng applies one pinv to every submodel (Model.cpp assign())
    for (unsigned i = 0; i < nmat[p]; ++i)
      if (!pll_update_invariant_sites_proportion(parts[p], i, pinv)) throw std::runtime_error(pll_errmsg);
  }
  void setCategoryRates(unsigned p, const double *rates) override { pll_set_category_rates(parts[p], rates); }
  bool gammaRates(double alpha, unsigned cats, double *out, int mode) const override { return pll_compute_gamma_cats(alpha, cats, out, mode) != 0; }
  void updatePmatrix(unsigned p, unsigned edge, double brlen) override {
    if (!pll_update_prob_matrices(parts[p], pidx[p].data(), &edge, &brlen, 1)) throw std::runtime_error(pll_errmsg);
  }
  const double *pmatrix(unsigned p, unsigned edge) const override { return parts[p]->pmatrix[edge]; }
  unsigned clvIndex(const Operand &o) const { return o.kind == 1 ? o.tip : (o.kind == 0 ? tips + 1 : fake_clv); }
  double *clvPtr(unsigned p, const Operand &o) const {
    return o.kind == 0 ? const_cast<double *>(o.clv) : (o.kind == 2 ? parts[p]->clv[fake_clv] : nullptr);
  }
  void updatePartials(unsigned p, double *pc, unsigned *ps, const Operand &l, const Operand &r) override {
    pll_operation_t op;  // LH/Operation.cpp:7-35
    op.parent_clv_index = tips + 1;
    op.parent_scaler_index = 0;
    op.child1_clv_index = clvIndex(l); op.child1_scaler_index = -1; op.child1_matrix_index = l.edge;
    op.child2_clv_index = clvIndex(r); op.child2_scaler_index = -1; op.child2_matrix_index = r.edge;
    pll_update_partials_single(parts[p], &op, 1, pc, clvPtr(p, l), clvPtr(p, r), ps,
                               l.kind == 0 ? const_cast<unsigned *>(l.scaler) : nullptr,
                               r.kind == 0 ? const_cast<unsigned *>(r.scaler) : nullptr);
  }
  double rootLogl(unsigned p, const double *clv, const unsigned *scaler, double *persite) override {
    return pll_compute_root_loglikelihood(parts[p], tips + 1, const_cast<double *>(clv), const_cast<unsigned *>(scaler), pidx[p].data(), persite);
  }
  double edgeLogl(unsigned p, const Operand &a, const Operand &b, unsigned edge, double *persite) override {
    return pll_compute_edge_loglikelihood(parts[p], clvIndex(a), clvPtr(p, a), a.kind == 0 ? const_cast<unsigned *>(a.scaler) : nullptr,
                                          clvIndex(b), clvPtr(p, b), b.kind == 0 ? const_cast<unsigned *>(b.scaler) : nullptr,
                                          edge, pidx[p].data(), persite);
  }
  void sumtable(unsigned p, const Operand &a, const Operand &b, double *out) override {
    if (!pll_update_sumtable(parts[p], clvIndex(a), clvPtr(p, a), clvIndex(b), clvPtr(p, b),
                             a.kind == 0 ? const_cast<unsigned *>(a.scaler) : nullptr,
                             b.kind == 0 ? const_cast<unsigned *>(b.scaler) : nullptr, pidx[p].data(), out))
      throw std::runtime_error(pll_errmsg);
  }
  void derivatives(unsigned p, const double *st, double brlen, bool want_f, double *f, double *d1, double *d2) override {
    // LH/LikelihoodDerivatives.cpp:75-88,98-106,146-170
    double **eigenvals = nullptr, *prop_invar = nullptr;
    pll_compute_eigenvals_and_prop_invar(parts[p], pidx[p].data(), &eigenvals, &prop_invar);
    double *diag = pll_compute_diagptable(parts[p]->states, parts[p]->rate_cats, brlen, prop_invar, parts[p]->rates, eigenvals);
    free(eigenvals);
    pll_compute_loglikelihood_derivatives(parts[p], 0, nullptr, 0, nullptr, brlen, pidx[p].data(), st,
                                          want_f ? f : nullptr, d1, d2, diag, prop_invar);
    pll_aligned_free(diag);
    free(prop_invar);
  }
};
}  // namespace

Backend *makeRefBackend(unsigned tips, unsigned edges_plus_fake, const std::vector<PartitionDesc> &descs) {
  RefBackend *b = new RefBackend();
  b->tips = tips;
  b->fake_clv = tips;  // clv_buffers: [0] = fake all-ones CLV, [1] = placeholder "inner" index
  for (const PartitionDesc &d : descs) {
    unsigned attrs = PLL_ATTRIB_ARCH_AVX2 | PLL_ATTRIB_PATTERN_TIP;  // RAXML/TreeInfo.cpp:646-676, SURVEY F4
    pll_partition_t *pp = pll_partition_create(tips, 2, d.states, d.sites, d.rate_cats /* room for one rate matrix per category */, edges_plus_fake, d.rate_cats, 1, attrs);
    if (!pp) throw std::runtime_error(std::string("pll_partition_create failed: ") + pll_errmsg);
    if (!d.pattern_weights.empty()) pll_set_pattern_weights(pp, d.pattern_weights.data());
    // tip states: synthetic char map code -> state mask (role of pll_map_nt / pll_map_aa)
    std::map<uint32_t, unsigned> code;
    std::vector<pll_state_t> map(256, 0);
    std::vector<std::string> seqs(tips, std::string(d.sites, '\0'));
    if (d.states == 4)   // a COMPLETE 16-entry map, as pll_map_nt is: with PATTERN_TIP + ARCH_AVX2 libpll sizes ttlookup from the largest
      for (uint32_t m = 1; m < 16; ++m) {   // mask in the map (LIBPLL/pll.c:352-396 tests ARCH_AVX only) while the 4x4 AVX lookup kernel it
        code.emplace(m, m);                 // dispatches to always writes all 16 x 16 entries (core_partials_avx.c:255-395): a map built from
        map[m] = m;                         // an alignment without T / gaps overflowed the heap (found with ASAN on 1-site partitions)
      }
    for (unsigned t = 0; t < tips; ++t)
      for (unsigned s = 0; s < d.sites; ++s) {
        uint32_t m = d.tip_masks[t][s];
        auto it = code.find(m);
        if (it == code.end()) {
          unsigned c = (unsigned)code.size() + 1;
          if (c > 127) throw std::runtime_error("too many distinct tip states");
          it = code.emplace(m, c).first;
          map[c] = m;
        }
        seqs[t][s] = (char)it->second;
      }
    for (unsigned t = 0; t < tips; ++t)
      if (!pll_set_tip_states(pp, t, map.data(), seqs[t].c_str())) throw std::runtime_error(std::string("pll_set_tip_states: ") + pll_errmsg);
    // fake CLV, SRC/RaxmlWrapper.cpp:156-187
    double *clv = pp->clv[b->fake_clv];
    for (unsigned n = 0; n < d.sites; ++n)
      for (unsigned i = 0; i < d.rate_cats; ++i) {
        for (unsigned j = 0; j < d.states; ++j) clv[j] = 1;
        clv += pp->states_padded;
      }
    b->parts.push_back(pp);
    b->pidx.emplace_back(64, 0u);
    b->nmat.push_back(1);
    b->setModel((unsigned)b->parts.size() - 1, d.freqs.data(), d.subst_params.data(), d.rates.data(), d.rate_weights.data());
  }
  return b;
}

}  // namespace orc
