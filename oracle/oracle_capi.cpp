/*
 * oracle_capi.cpp — ORACLE (test infrastructure, NOT product code).
 * Flat C API over netrax_port.cpp for ctypes (tests/, smoke(), bench.py cpu_baseline / --impl reference).
 * The function set deliberately has the same shape as the product's host C-ABI (include/netrax_b200.h,
 * prefix nrxh_) so that parity tests drive both sides with identical call sequences.
 */
#include "netrax_port.hpp"

#include <cstdio>
#include <string>

using namespace orc;

namespace {
thread_local std::string g_err;

struct Handle {
  std::string backend_kind;
  AnnotatedNetwork ann;
  std::vector<PartitionDesc> descs;
  std::vector<std::vector<double>> part_brlens;  // optional unlinked per-partition branch lengths
  bool net_set = false, inited = false;
  // branch-length optimisation state (role of the locals of optimize_branch, BranchLengthOptimization.cpp:345-420)
  std::vector<DisplayedTreeData> oldTrees;
  std::vector<std::vector<SumtableInfo>> sumtables;
};

template <class F> int guarded(F &&f) {
  try { f(); return 1; }
  catch (const std::exception &e) { g_err = e.what(); return 0; }
  catch (...) { g_err = "unknown error"; return 0; }
}
}  // namespace

extern "C" {

const char *orc_last_error() { return g_err.c_str(); }

void *orc_new(const char *backend) {
  Handle *h = new Handle();
  h->backend_kind = backend ? backend : "port";
  return h;
}

void orc_free(void *hv) { delete static_cast<Handle *>(hv); }

const char *orc_backend_kind(void *hv) {
  Handle *h = static_cast<Handle *>(hv);
  return h->ann.backend ? h->ann.backend->kind() : h->backend_kind.c_str();
}

int orc_has_ref_backend() {
#ifdef ORC_HAVE_REF
  return 1;
#else
  return 0;
#endif
}

int orc_set_network(void *hv, unsigned num_tips, unsigned num_nodes, unsigned root, unsigned num_edges,
                    const unsigned *src, const unsigned *tgt, const double *len, const double *prob,
                    unsigned num_ret, const unsigned *ret_node, const unsigned *ret_first, const unsigned *ret_second) {
  Handle *h = static_cast<Handle *>(hv);
  return guarded([&] {
    std::vector<Network::Edge> edges(num_edges);
    for (unsigned e = 0; e < num_edges; ++e) edges[e] = Network::Edge{src[e], tgt[e], len[e], prob ? prob[e] : 1.0};
    std::vector<unsigned> rn(ret_node, ret_node + num_ret), rf(ret_first, ret_first + num_ret), rs(ret_second, ret_second + num_ret);
    h->ann.network.build(num_tips, num_nodes, root, edges, rn, rf, rs);
    h->net_set = true;
  });
}

int orc_add_partition(void *hv, unsigned states, unsigned rate_cats, unsigned sites, const uint32_t *tip_masks,
                      const unsigned *weights, const double *freqs, const double *subst, const double *rates,
                      const double *rate_weights) {
  Handle *h = static_cast<Handle *>(hv);
  return guarded([&] {
    if (!h->net_set) throw std::runtime_error("set the network first");
    PartitionDesc d;
    d.states = states; d.rate_cats = rate_cats; d.sites = sites;
    d.freqs.assign(freqs, freqs + states);
    d.subst_params.assign(subst, subst + states * (states - 1) / 2);
    d.rates.assign(rates, rates + rate_cats);
    d.rate_weights.assign(rate_weights, rate_weights + rate_cats);
    if (weights) d.pattern_weights.assign(weights, weights + sites);
    else d.pattern_weights.assign(sites, 1);
    for (unsigned w : d.pattern_weights) h->ann.total_num_sites += w;
    { double sum = 0; for (unsigned w : d.pattern_weights) sum += w; h->ann.pattern_weight_sums.push_back(sum); }
    unsigned tips = h->ann.network.num_tips;
    d.tip_masks.resize(tips);
    for (unsigned t = 0; t < tips; ++t) d.tip_masks[t].assign(tip_masks + (size_t)t * sites, tip_masks + (size_t)(t + 1) * sites);
    h->descs.emplace_back(std::move(d));
  });
}

int orc_set_options(void *hv, int likelihood_variant, int brlen_linkage) {
  Handle *h = static_cast<Handle *>(hv);
  h->ann.options.likelihood_variant = likelihood_variant == 2 ? LikelihoodVariant::SARAH_PSEUDO : (likelihood_variant ? LikelihoodVariant::BEST_DISPLAYED_TREE : LikelihoodVariant::AVERAGE_DISPLAYED_TREES);
  h->ann.options.brlen_linkage = brlen_linkage;
  return 1;
}

int orc_set_partition_brlens(void *hv, unsigned p, const double *brlens) {
  Handle *h = static_cast<Handle *>(hv);
  return guarded([&] {
    size_t E = h->ann.network.edges.size();
    if (h->part_brlens.size() <= p) h->part_brlens.resize(p + 1);
    h->part_brlens[p].assign(brlens, brlens + E);
    h->part_brlens[p].push_back(0.0);  // fake branch
    if (h->inited) {
      h->ann.branch_lengths[p] = h->part_brlens[p];
      for (size_t e = 0; e < E; ++e) h->ann.pmatrix_valid[p][e] = 0;
      invalidateAllCLVs(h->ann);
    }
  });
}

int orc_set_reduce_callback(void *hv, void (*cb)(void *, double *, size_t, int), void *ctx) {
  Handle *h = static_cast<Handle *>(hv);
  h->ann.parallel_reduce_cb = cb;
  h->ann.parallel_context = ctx;
  return 1;
}

int orc_init(void *hv) {
  Handle *h = static_cast<Handle *>(hv);
  return guarded([&] {
    unsigned tips = h->ann.network.num_tips, ef = (unsigned)h->ann.network.edges.size() + 1;
    if (h->backend_kind == "port") h->ann.backend.reset(makePortBackend(tips, ef, h->descs));
#ifdef ORC_HAVE_REF
    else if (h->backend_kind == "ref") h->ann.backend.reset(makeRefBackend(tips, ef, h->descs));
#endif
    else throw std::runtime_error("unknown / unavailable oracle backend: " + h->backend_kind);
    if (!h->part_brlens.empty()) {
      std::vector<double> linked(h->ann.network.edges.size() + 1, 0.0);
      for (size_t e = 0; e < h->ann.network.edges.size(); ++e) linked[e] = h->ann.network.edges[e].length;
      h->ann.branch_lengths.assign(h->descs.size(), linked);
      for (size_t p = 0; p < h->part_brlens.size() && p < h->descs.size(); ++p)
        if (!h->part_brlens[p].empty()) h->ann.branch_lengths[p] = h->part_brlens[p];
    }
    init_annotated_network(h->ann);
    h->inited = true;
  });
}

int orc_compute_loglikelihood(void *hv, int incremental, int update_pmatrices, double *out) {
  Handle *h = static_cast<Handle *>(hv);
  return guarded([&] { *out = computeLoglikelihood(h->ann, incremental, update_pmatrices); });
}

int orc_naive_loglikelihood(void *hv, double *out, double *tree_logl, double *tree_logprob) {
  Handle *h = static_cast<Handle *>(hv);
  return guarded([&] {
    std::vector<double> tl, lp;
    *out = computeLoglikelihoodNaive(h->ann, &tl, &lp);
    if (tree_logl) std::copy(tl.begin(), tl.end(), tree_logl);
    if (tree_logprob) std::copy(lp.begin(), lp.end(), tree_logprob);
  });
}

/* per-site lnL of root displayed tree `tree`: out[p * stride + site] (same shape as the product's nrxh_persite_lnl) */
int orc_persite_lnl(void *hv, unsigned tree, double *out, unsigned stride) {
  Handle *h = static_cast<Handle *>(hv);
  return guarded([&] {
    for (unsigned p = 0; p < h->ann.partitionCount(); ++p) {
      if (h->ann.backend->sites(p) > stride) throw std::runtime_error("orc_persite_lnl: stride smaller than the partition");
      persiteLoglikelihood(h->ann, tree, p, out + (size_t)p * stride);
    }
  });
}

unsigned orc_num_partitions(void *hv) { return static_cast<Handle *>(hv)->ann.partitionCount(); }
unsigned orc_root(void *hv) { return static_cast<Handle *>(hv)->ann.network.root; }
unsigned orc_num_nodes(void *hv) { return (unsigned)static_cast<Handle *>(hv)->ann.network.num_nodes(); }

int orc_num_trees(void *hv, unsigned node) {
  Handle *h = static_cast<Handle *>(hv);
  return (int)h->ann.pernode_displayed_tree_data[node].num_active_displayed_trees;
}

int orc_tree_config(void *hv, unsigned node, unsigned tree, char *buf, unsigned buflen) {
  Handle *h = static_cast<Handle *>(hv);
  return guarded([&] {
    std::string s = configToString(h->ann.pernode_displayed_tree_data[node].displayed_trees.at(tree).treeLoglData.reticulationChoices, h->ann.network.num_reticulations());
    std::snprintf(buf, buflen, "%s", s.c_str());
  });
}

int orc_tree_info(void *hv, unsigned node, unsigned tree, double *logprob, double *partition_logl, int *flags) {
  Handle *h = static_cast<Handle *>(hv);
  return guarded([&] {
    const DisplayedTreeData &d = h->ann.pernode_displayed_tree_data[node].displayed_trees.at(tree);
    if (logprob) *logprob = d.treeLoglData.tree_logprob;
    if (partition_logl) std::copy(d.treeLoglData.tree_partition_logl.begin(), d.treeLoglData.tree_partition_logl.end(), partition_logl);
    if (flags) *flags = (d.clv_valid ? 1 : 0) | (d.treeLoglData.tree_logl_valid ? 2 : 0) | (d.treeLoglData.tree_logprob_valid ? 4 : 0);
  });
}

int orc_read_clv(void *hv, unsigned node, unsigned tree, unsigned p, double *out) {
  Handle *h = static_cast<Handle *>(hv);
  return guarded([&] {
    const DisplayedTreeData &d = h->ann.pernode_displayed_tree_data[node].displayed_trees.at(tree);
    if (d.isTip) throw std::runtime_error("tips have no CLV (PATTERN_TIP)");
    std::memcpy(out, d.clv_vector[p].p, d.clv_vector[p].n * sizeof(double));
  });
}

int orc_read_scaler(void *hv, unsigned node, unsigned tree, unsigned p, unsigned *out) {
  Handle *h = static_cast<Handle *>(hv);
  return guarded([&] {
    const DisplayedTreeData &d = h->ann.pernode_displayed_tree_data[node].displayed_trees.at(tree);
    if (d.isTip) throw std::runtime_error("tips have no scaler");
    std::memcpy(out, d.scale_buffer[p].p, d.scale_buffer[p].n * sizeof(unsigned));
  });
}

int orc_partition_loglh(void *hv, double *out) {
  Handle *h = static_cast<Handle *>(hv);
  std::copy(h->ann.partition_loglh.begin(), h->ann.partition_loglh.end(), out);
  return 1;
}

int orc_set_branch_length(void *hv, int partition, unsigned edge, double value) {
  Handle *h = static_cast<Handle *>(hv);
  return guarded([&] { setBranchLength(h->ann, partition, edge, value); invalidatePmatrixIndex(h->ann, edge); });
}

int orc_set_reticulation_prob(void *hv, unsigned r, double prob) {
  Handle *h = static_cast<Handle *>(hv);
  return guarded([&] { setReticulationProb(h->ann, r, prob); });
}

int orc_set_model(void *hv, unsigned p, const double *freqs, const double *subst, const double *rates, const double *weights) {
  Handle *h = static_cast<Handle *>(hv);
  return guarded([&] {
    h->ann.backend->setModel(p, freqs, subst, rates, weights);
    for (auto &v : h->ann.pmatrix_valid[p]) v = 0;
    invalidateAllCLVs(h->ann);
  });
}

int orc_get_eigen(void *hv, unsigned p, double *eigenvecs, double *inv_eigenvecs, double *eigenvals) {
  Handle *h = static_cast<Handle *>(hv);
  return guarded([&] { h->ann.backend->getEigen(p, eigenvecs, inv_eigenvecs, eigenvals); });
}

int orc_get_pmatrix(void *hv, unsigned p, unsigned edge, double *out) {
  Handle *h = static_cast<Handle *>(hv);
  return guarded([&] {
    const Backend &b = *h->ann.backend;
    std::memcpy(out, b.pmatrix(p, edge), sizeof(double) * b.rateCats(p) * b.states(p) * b.statesPadded(p));
  });
}

int orc_gamma_rates(double alpha, unsigned cats, int mode, double *out);  // below

/* ---- branch-length optimisation flow (BranchLengthOptimization.cpp:345-420) ---------------- */
int orc_brlen_prepare(void *hv, unsigned edge, double *old_logl) {
  Handle *h = static_cast<Handle *>(hv);
  return guarded([&] {
    AnnotatedNetwork &ann = h->ann;
    double l = computeLoglikelihood(ann, 1, 1);
    if (old_logl) *old_logl = l;
    h->oldTrees = extractOldTrees(ann, ann.network.root);
    ConfigSet restrictions = getRestrictionsActiveAliveBranch(ann, edge);
    updateCLVsVirtualRerootTrees(ann, ann.network.root, ann.network.edges[edge].source, ann.network.edges[edge].target, restrictions);
    ann.cached_logl_valid = false;
  });
}

int orc_brlen_logl(void *hv, unsigned edge, double *out) {
  Handle *h = static_cast<Handle *>(hv);
  return guarded([&] {
    h->ann.cached_logl_valid = false;
    *out = computeLoglikelihoodBrlenOpt(h->ann, h->oldTrees, edge, 1);
  });
}

int orc_brlen_sumtables(void *hv, unsigned edge, unsigned *count) {
  Handle *h = static_cast<Handle *>(hv);
  return guarded([&] {
    h->sumtables = computePartitionSumtables(h->ann, edge);
    if (count) *count = h->sumtables.empty() ? 0 : (unsigned)h->sumtables[0].size();
  });
}

int orc_brlen_read_sumtable(void *hv, unsigned p, unsigned idx, double *out, double *tree_prob, unsigned *left_tree, unsigned *right_tree) {
  Handle *h = static_cast<Handle *>(hv);
  return guarded([&] {
    const SumtableInfo &s = h->sumtables.at(p).at(idx);
    if (out) std::memcpy(out, s.sumtable.p, s.sumtable.n * sizeof(double));
    if (tree_prob) *tree_prob = s.tree_prob;
    if (left_tree) *left_tree = (unsigned)s.left_tree_idx;
    if (right_tree) *right_tree = (unsigned)s.right_tree_idx;
  });
}

/* network_derivative_func_multi (BranchLengthOptimization.cpp:171-200): set the proposal, then derivatives */
int orc_brlen_set_length(void *hv, int partition, unsigned edge, double value) {
  Handle *h = static_cast<Handle *>(hv);
  return guarded([&] {
    AnnotatedNetwork &ann = h->ann;
    if (ann.options.brlen_linkage == BRLEN_UNLINKED && partition >= 0) ann.branch_lengths[partition][edge] = value;
    else setBranchLength(ann, -1, edge, value);
    invalidPmatrixIndexOnly(ann, edge);
  });
}

int orc_brlen_derivatives(void *hv, unsigned edge, double *d1, double *d2, double *part_d1, double *part_d2, double *raw) {
  Handle *h = static_cast<Handle *>(hv);
  return guarded([&] {
    LoglDerivatives r = computeLoglikelihoodDerivatives(h->ann, h->sumtables, edge);
    if (d1) *d1 = r.logl_prime;
    if (d2) *d2 = r.logl_prime_prime;
    if (part_d1) std::copy(r.partition_logl_prime.begin(), r.partition_logl_prime.end(), part_d1);
    if (part_d2) std::copy(r.partition_logl_prime_prime.begin(), r.partition_logl_prime_prime.end(), part_d2);
    if (raw) {
      size_t n = h->sumtables.empty() ? 0 : h->sumtables[0].size();
      for (size_t p = 0; p < r.raw.size(); ++p)
        for (size_t k = 0; k < r.raw[p].size(); ++k) raw[p * 3 * n + k] = r.raw[p][k];
    }
  });
}

int orc_brlen_finish(void *hv, unsigned edge, double *final_logl) {
  Handle *h = static_cast<Handle *>(hv);
  return guarded([&] {
    h->sumtables.clear();
    h->oldTrees.clear();
    invalidatePmatrixIndex(h->ann, edge);
    double l = computeLoglikelihood(h->ann, 1, 1);
    if (final_logl) *final_logl = l;
  });
}

/* ---- the immediate callers (optimize_port.cpp) -------------------------------------------------------- */
int orc_optimize_branch(void *hv, unsigned edge, int method, unsigned max_iters, double *final_logl) {
  Handle *h = static_cast<Handle *>(hv);
  return guarded([&] { double l = optimize_branch(h->ann, edge, method, max_iters); if (final_logl) *final_logl = l; });
}
int orc_optimize_branches(void *hv, int max_iters, int max_iters_outside, int radius, int method, double *final_logl) {
  Handle *h = static_cast<Handle *>(hv);
  return guarded([&] { double l = optimize_branches(h->ann, max_iters, max_iters_outside, radius, method); if (final_logl) *final_logl = l; });
}
int orc_optimize_reticulation(void *hv, unsigned r, double *final_logl) {
  Handle *h = static_cast<Handle *>(hv);
  return guarded([&] { double l = optimize_reticulation(h->ann, r); if (final_logl) *final_logl = l; });
}
int orc_compute_pseudo_loglikelihood(void *hv, int incremental, int update_pmatrices, double *out) {
  Handle *h = static_cast<Handle *>(hv);
  return guarded([&] { *out = computePseudoLoglikelihood(h->ann, incremental, update_pmatrices); });
}
int orc_read_pseudo_clv(void *hv, unsigned node, unsigned p, double *out) {
  Handle *h = static_cast<Handle *>(hv);
  return guarded([&] {
    if (node >= h->ann.pseudo_clv.size() || p >= h->ann.pseudo_clv[node].size()) throw std::runtime_error("no pseudo CLV at that node");
    std::memcpy(out, h->ann.pseudo_clv[node][p].p, sizeof(double) * h->ann.backend->clvEntries(p));
  });
}
int orc_read_pseudo_scaler(void *hv, unsigned node, unsigned p, unsigned *out) {
  Handle *h = static_cast<Handle *>(hv);
  return guarded([&] {
    if (node >= h->ann.pseudo_scaler.size() || p >= h->ann.pseudo_scaler[node].size()) throw std::runtime_error("no pseudo CLV at that node");
    std::memcpy(out, h->ann.pseudo_scaler[node][p].p, sizeof(unsigned) * h->ann.backend->sites(p));
  });
}
int orc_score_network(void *hv, double *bic_score) {
  Handle *h = static_cast<Handle *>(hv);
  return guarded([&] { *bic_score = scoreNetwork(h->ann); });
}
int orc_set_scoring_sizes(void *hv, unsigned long long total_num_model_parameters, unsigned long long total_num_sites) {
  Handle *h = static_cast<Handle *>(hv);
  h->ann.total_num_model_parameters = (size_t)total_num_model_parameters;
  if (total_num_sites) h->ann.total_num_sites = (size_t)total_num_sites;
  return 1;
}
int orc_optimize_all_non_topology(void *hv, int type, double *bic_score) {
  Handle *h = static_cast<Handle *>(hv);
  return guarded([&] { optimizeAllNonTopology(h->ann, type); if (bic_score) *bic_score = scoreNetwork(h->ann); });
}
int orc_set_pinv(void *hv, unsigned p, double prop_invar) {
  Handle *h = static_cast<Handle *>(hv);
  return guarded([&] { setPinv(h->ann, p, prop_invar); });
}
int orc_set_params_to_optimize(void *hv, unsigned p, int mask) {
  Handle *h = static_cast<Handle *>(hv);
  return guarded([&] {
    if (p >= h->ann.partitionCount()) throw std::runtime_error("partition index out of range");
    if (h->ann.params_to_optimize.size() <= p) h->ann.params_to_optimize.resize(h->ann.partitionCount(), -1);
    h->ann.params_to_optimize[p] = mask;
  });
}
int orc_set_brlen_scaler(void *hv, unsigned p, double scaler) {
  Handle *h = static_cast<Handle *>(hv);
  return guarded([&] { setBrlenScaler(h->ann, p, scaler); });
}
int orc_set_submodels(void *hv, unsigned p, unsigned n, const unsigned *cat_model, const double *freqs, const double *subst) {
  Handle *h = static_cast<Handle *>(hv);
  return guarded([&] { setSubmodels(h->ann, p, n, cat_model, freqs, subst); });
}
int orc_set_alpha(void *hv, unsigned p, double alpha) {
  Handle *h = static_cast<Handle *>(hv);
  return guarded([&] { setAlpha(h->ann, p, alpha); });
}
int orc_get_alpha(void *hv, unsigned p, double *alpha) {
  Handle *h = static_cast<Handle *>(hv);
  return guarded([&] { *alpha = p < h->ann.alphas.size() ? h->ann.alphas[p] : 0.0; });
}
int orc_optimize_alpha(void *hv, double min_alpha, double max_alpha, double tolerance, double *final_logl) {
  Handle *h = static_cast<Handle *>(hv);
  return guarded([&] { double l = optimize_alpha(h->ann, min_alpha, max_alpha, tolerance); if (final_logl) *final_logl = l; });
}
int orc_optimize_pinv(void *hv, double min_pinv, double max_pinv, double tolerance, double *final_logl) {
  Handle *h = static_cast<Handle *>(hv);
  return guarded([&] { double l = optimize_pinv(h->ann, min_pinv, max_pinv, tolerance); if (final_logl) *final_logl = l; });
}
int orc_optimize_scalers(void *hv, double *bic_score) {
  Handle *h = static_cast<Handle *>(hv);
  return guarded([&] { double b = optimize_scalers(h->ann); if (bic_score) *bic_score = b; });
}
int orc_get_brlen_scalers(void *hv, double *out) {
  Handle *h = static_cast<Handle *>(hv);
  return guarded([&] { for (unsigned p = 0; p < h->ann.partitionCount(); ++p) out[p] = p < h->ann.brlen_scalers.size() ? h->ann.brlen_scalers[p] : 1.0; });
}
int orc_get_pinv(void *hv, unsigned p, double *out) {
  Handle *h = static_cast<Handle *>(hv);
  return guarded([&] { *out = p < h->ann.pinvs.size() ? h->ann.pinvs[p] : 0.0; });
}
int orc_optimize_reticulations(void *hv, int max_iters, double *final_logl) {
  Handle *h = static_cast<Handle *>(hv);
  return guarded([&] { double l = optimize_reticulations(h->ann, max_iters); if (final_logl) *final_logl = l; });
}
int orc_get_branch_lengths(void *hv, int partition, double *out) {
  Handle *h = static_cast<Handle *>(hv);
  return guarded([&] {
    const AnnotatedNetwork &ann = h->ann;
    const std::vector<double> &b = (partition < 0 || ann.options.brlen_linkage != BRLEN_UNLINKED) ? ann.linked_branch_lengths : ann.branch_lengths.at(partition);
    std::copy(b.begin(), b.begin() + ann.network.num_branches(), out);
  });
}
int orc_get_reticulation_probs(void *hv, double *out) {
  Handle *h = static_cast<Handle *>(hv);
  return guarded([&] { std::copy(h->ann.reticulation_probs.begin(), h->ann.reticulation_probs.begin() + h->ann.network.num_reticulations(), out); });
}
/* direct access to the two 1-D minimisers on a test function (pins opt_port.c against the reference's opt_algorithms.c) */
extern "C" {
#ifdef ORC_HAVE_REF
int pllmod_opt_minimize_newton_multi(unsigned int, double, double *, double, double, unsigned int, int *, void *, void(deriv_func)(void *, double *, double *, double *));
double pllmod_opt_minimize_brent(double, double, double, double, double *, double *, void *, double (*)(void *, double));
int pllmod_opt_minimize_brent_multi(unsigned int, int *, double *, double *, double *, double, double *, double *, double *, void *,
                                    double (*)(void *, double *, double *, int *), int);
#endif
int orcopt_brent_multi(unsigned int, int *, double *, double *, double *, double, double *, double *, double *, void *,
                       double (*)(void *, double *, double *, int *), int);
int orcopt_newton_multi(unsigned int, double, double *, double, double, unsigned int, int *, void *, void (*)(void *, double *, double *, double *));
double orcopt_brent(double, double, double, double, double *, double *, void *, double (*)(void *, double));
}
int orc_test_brent(int use_ref, double xmin, double xguess, double xmax, double xtol, double (*target)(void *, double), double *xopt) {
  double fx = 0, f2x = 0;
#ifdef ORC_HAVE_REF
  if (use_ref) { *xopt = pllmod_opt_minimize_brent(xmin, xguess, xmax, xtol, &fx, &f2x, nullptr, target); return 1; }
#endif
  if (use_ref) return 0;
  *xopt = orcopt_brent(xmin, xguess, xmax, xtol, &fx, &f2x, nullptr, target);
  return 1;
}
int orc_test_brent_multi(int use_ref, unsigned n, double xmin, double *x, double xmax, double xtol, double (*target)(void *, double *, double *, int *)) {
  std::vector<int> mask(n, 1);
#ifdef ORC_HAVE_REF
  if (use_ref) return pllmod_opt_minimize_brent_multi(n, mask.data(), &xmin, x, &xmax, xtol, x, nullptr, nullptr, nullptr, target, 1);
#endif
  if (use_ref) return 0;
  return orcopt_brent_multi(n, mask.data(), &xmin, x, &xmax, xtol, x, nullptr, nullptr, nullptr, target, 1);
}
int orc_test_newton(int use_ref, double xmin, double *x, double xmax, double tol, unsigned max_iters, void (*deriv)(void *, double *, double *, double *), int *status) {
#ifdef ORC_HAVE_REF
  if (use_ref) { *status = pllmod_opt_minimize_newton_multi(1, xmin, x, xmax, tol, max_iters, nullptr, nullptr, deriv); return 1; }
#endif
  if (use_ref) return 0;
  *status = orcopt_newton_multi(1, xmin, x, xmax, tol, max_iters, nullptr, nullptr, deriv);
  return 1;
}

unsigned long long orc_clv_update_count(void *hv) { return static_cast<Handle *>(hv)->ann.n_clv_updates; }
void orc_reset_counters(void *hv) { static_cast<Handle *>(hv)->ann.n_clv_updates = 0; }

}  // extern "C"

#include "pll_port.h"
extern "C" int orc_gamma_rates(double alpha, unsigned cats, int mode, double *out) {
  return port_compute_gamma_cats(alpha, cats, out, mode);
}
