/* capi.cpp — include/netrax_b200.h over netrax_likelihood_api.hpp (exceptions -> status + message). */
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../../../include/netrax_b200.h"
#include "host_internal.hpp"

using namespace netrax;

namespace {
thread_local std::string g_err;

struct Handle {
  AnnotatedNetwork ann;
  std::vector<PartitionInput> inputs;
  std::vector<std::vector<uint32_t>> masks, weights;
  std::vector<std::vector<double>> part_brlens;
  int device = 0;
  bool net_set = false, inited = false;
  std::vector<DisplayedTreeData> oldTrees;
  std::vector<std::vector<SumtableInfo>> sumtables;
  NetworkParams params{nullptr};
};

template <class F> int guarded(F &&f) {
  try { f(); return 1; }
  catch (const std::exception &e) { g_err = e.what(); return 0; }
  catch (...) { g_err = "unknown error"; return 0; }
}
Handle *H(void *h) { return static_cast<Handle *>(h); }
}  // namespace

extern "C" {

const char *nrxh_last_error(void) { return g_err.c_str(); }

void *nrxh_new(const char *options) {
  Handle *h = new Handle();
  if (options) {
    std::string o(options);
    size_t p = o.find("device=");
    if (p != std::string::npos) h->device = std::atoi(o.c_str() + p + 7);
    p = o.find("plan_cache=");
    if (p != std::string::npos) h->ann.use_plan_cache = std::atoi(o.c_str() + p + 11) != 0;
  }
  return h;
}

void nrxh_free(void *h) { delete H(h); }

int nrxh_set_network(void *hv, unsigned num_tips, unsigned num_nodes, unsigned root, unsigned num_edges, const unsigned *src,
                     const unsigned *tgt, const double *len, const double *prob, unsigned num_ret, const unsigned *ret_node,
                     const unsigned *ret_first, const unsigned *ret_second) {
  return guarded([&] {
    std::vector<Edge> edges(num_edges);
    for (unsigned e = 0; e < num_edges; ++e) { edges[e].source = src[e]; edges[e].target = tgt[e]; edges[e].length = len[e]; edges[e].prob = prob ? prob[e] : 1.0; }
    std::vector<size_t> rn(ret_node, ret_node + num_ret), rf(ret_first, ret_first + num_ret), rs(ret_second, ret_second + num_ret);
    H(hv)->ann.network = buildNetwork(num_tips, num_nodes, root, edges, rn, rf, rs);
    // nodes_by_index etc. point into the moved-from vectors' storage: rebuild the pointer tables
    Network &nw = H(hv)->ann.network;
    for (size_t i = 0; i < nw.nodes.size(); ++i) nw.nodes_by_index[i] = &nw.nodes[i];
    for (size_t i = 0; i < nw.edges.size(); ++i) nw.edges_by_index[i] = &nw.edges[i];
    for (size_t i = 0; i < nw.reticulations.size(); ++i) nw.reticulation_nodes[i] = &nw.nodes[nw.reticulations[i].node];
    nw.root = &nw.nodes[root];
    H(hv)->net_set = true;
  });
}

int nrxh_add_partition(void *hv, unsigned states, unsigned rate_cats, unsigned sites, const uint32_t *tip_masks, const unsigned *pw,
                       const double *freqs, const double *subst, const double *rates, const double *rate_weights) {
  return guarded([&] {
    Handle *h = H(hv);
    if (!h->net_set) throw std::runtime_error("set the network first");
    PartitionInput in;
    in.model.states = states; in.model.rate_cats = rate_cats; in.model.sites = sites;
    set_frequencies(in.model, freqs);
    in.model.subst_params.assign(subst, subst + states * (states - 1) / 2);
    in.model.rates.assign(rates, rates + rate_cats);
    in.model.rate_weights.assign(rate_weights, rate_weights + rate_cats);
    h->masks.emplace_back(tip_masks, tip_masks + (size_t)h->ann.network.num_tips() * sites);
    h->weights.emplace_back(pw ? std::vector<uint32_t>(pw, pw + sites) : std::vector<uint32_t>());
    h->inputs.push_back(in);
  });
}

int nrxh_set_options(void *hv, int variant, int linkage) {
  H(hv)->ann.options.likelihood_variant = variant == 2 ? LikelihoodVariant::SARAH_PSEUDO : (variant ? LikelihoodVariant::BEST_DISPLAYED_TREE : LikelihoodVariant::AVERAGE_DISPLAYED_TREES);
  H(hv)->ann.options.brlen_linkage = linkage;
  return 1;
}

int nrxh_set_partition_brlens(void *hv, unsigned p, const double *brlens) {
  return guarded([&] {
    Handle *h = H(hv);
    const size_t E = h->ann.network.num_branches();
    if (h->part_brlens.size() <= p) h->part_brlens.resize(p + 1);
    h->part_brlens[p].assign(brlens, brlens + E);
    h->part_brlens[p].push_back(0.0);
    if (h->inited) {
      h->ann.fake_treeinfo->branch_lengths[p] = h->part_brlens[p];
      for (size_t e = 0; e < E; ++e) h->ann.fake_treeinfo->pmatrix_valid[p][e] = 0;
      invalidateAllCLVs(h->ann);
    }
  });
}

int nrxh_set_reduce_callback(void *hv, nrxh_reduce_cb cb, void *ctx) {
  H(hv)->ann.fake_treeinfo->parallel_reduce_cb = cb;
  H(hv)->ann.fake_treeinfo->parallel_context = ctx;
  return 1;
}

int nrxh_comm_get_unique_id(uint8_t *id128) {
  if (nrx_comm_get_unique_id(id128)) return 1;
  g_err = nrx_last_error();
  return 0;
}

int nrxh_comm_init(void *hv, const uint8_t *id128, int rank, int nranks) {
  return guarded([&] {
    Handle *h = H(hv);
    if (!h->ann.engine) throw std::runtime_error("nrxh_comm_init: call nrxh_init first");
    netrax::detail::engineCheck(nrx_comm_init(h->ann.engine, id128, rank, nranks), "nrx_comm_init");
  });
}

int nrxh_init(void *hv) {
  return guarded([&] {
    Handle *h = H(hv);
    for (size_t p = 0; p < h->inputs.size(); ++p) {
      h->inputs[p].tip_masks = h->masks[p].data();
      h->inputs[p].pattern_weights = h->weights[p].empty() ? nullptr : h->weights[p].data();
    }
    if (!h->part_brlens.empty()) {
      std::vector<double> linked(h->ann.network.num_branches() + 1, 0.0);
      for (size_t e = 0; e < h->ann.network.num_branches(); ++e) linked[e] = h->ann.network.edges[e].length;
      h->ann.fake_treeinfo->branch_lengths.assign(h->inputs.size(), linked);
      for (size_t p = 0; p < h->part_brlens.size() && p < h->inputs.size(); ++p)
        if (!h->part_brlens[p].empty()) h->ann.fake_treeinfo->branch_lengths[p] = h->part_brlens[p];
    }
    init_annotated_network(h->ann, h->inputs, h->device);
    h->masks.clear(); h->masks.shrink_to_fit();
    h->inited = true;
  });
}

int nrxh_compute_loglikelihood(void *hv, int incremental, int update_pmatrices, double *out) {
  return guarded([&] { *out = computeLoglikelihood(H(hv)->ann, incremental, update_pmatrices); });
}

int nrxh_compute_loglikelihood_batch(void **handles, unsigned n, int incremental, int update_pmatrices, double *out) {
  return guarded([&] {
    std::vector<AnnotatedNetwork *> anns;
    for (unsigned i = 0; i < n; ++i) anns.push_back(&H(handles[i])->ann);
    const std::vector<double> r = computeLoglikelihoodBatch(anns, incremental, update_pmatrices);
    for (unsigned i = 0; i < n; ++i) out[i] = r[i];
  });
}

unsigned nrxh_num_partitions(void *hv) { return H(hv)->ann.fake_treeinfo->partition_count; }
unsigned nrxh_root(void *hv) { return (unsigned)H(hv)->ann.network.root->clv_index; }
unsigned nrxh_num_nodes(void *hv) { return (unsigned)H(hv)->ann.network.num_nodes(); }
static void checkNode(AnnotatedNetwork &ann, unsigned node) {
  if (node >= ann.pernode_displayed_tree_data.size()) throw std::runtime_error("node index " + std::to_string(node) + " out of range");
}
int nrxh_num_trees(void *hv, unsigned node) {
  if (node >= H(hv)->ann.pernode_displayed_tree_data.size()) { g_err = "node index " + std::to_string(node) + " out of range"; return -1; }
  return (int)H(hv)->ann.pernode_displayed_tree_data[node].num_active_displayed_trees;
}

int nrxh_tree_config(void *hv, unsigned node, unsigned tree, char *buf, unsigned buflen) {
  return guarded([&] {
    checkNode(H(hv)->ann, node);
    std::string s = toString(H(hv)->ann.pernode_displayed_tree_data[node].displayed_trees.at(tree).treeLoglData.reticulationChoices, H(hv)->ann.network.num_reticulations());
    std::snprintf(buf, buflen, "%s", s.c_str());
  });
}

int nrxh_tree_info(void *hv, unsigned node, unsigned tree, double *logprob, double *partition_logl, int *flags) {
  return guarded([&] {
    checkNode(H(hv)->ann, node);
    const DisplayedTreeData &d = H(hv)->ann.pernode_displayed_tree_data[node].displayed_trees.at(tree);
    if (logprob) *logprob = d.treeLoglData.tree_logprob;
    if (partition_logl) std::copy(d.treeLoglData.tree_partition_logl.begin(), d.treeLoglData.tree_partition_logl.end(), partition_logl);
    if (flags) *flags = (d.clv_valid ? 1 : 0) | (d.treeLoglData.tree_logl_valid ? 2 : 0) | (d.treeLoglData.tree_logprob_valid ? 4 : 0);
  });
}

int nrxh_set_lazy_rerooting(void *hv, int on) {
  return guarded([&] { H(hv)->ann.lazy_reroot = on != 0; });
}

int nrxh_set_score_only(void *hv, int on) {
  return guarded([&] { H(hv)->ann.score_only = on != 0; });
}

int nrxh_read_clv(void *hv, unsigned node, unsigned tree, unsigned p, double *out) {
  return guarded([&] {
    detail::flushPendingOps(H(hv)->ann);
    if (H(hv)->ann.root_clvs_stale) computeLoglikelihood(H(hv)->ann, 1, 1);   // a score-only evaluation did not store the root trees' CLVs
    const DisplayedTreeData &d = H(hv)->ann.pernode_displayed_tree_data[node].displayed_trees.at(tree);
    if (d.isTip) throw std::runtime_error("tips have no CLV (PATTERN_TIP)");
    detail::engineCheck(nrx_read_clv(H(hv)->ann.engine, p, d.slot, out), "nrx_read_clv");
  });
}

int nrxh_read_scaler(void *hv, unsigned node, unsigned tree, unsigned p, unsigned *out) {
  return guarded([&] {
    detail::flushPendingOps(H(hv)->ann);
    const DisplayedTreeData &d = H(hv)->ann.pernode_displayed_tree_data[node].displayed_trees.at(tree);
    if (d.isTip) throw std::runtime_error("tips have no scaler");
    detail::engineCheck(nrx_read_scaler(H(hv)->ann.engine, p, d.slot, out), "nrx_read_scaler");
  });
}

int nrxh_partition_loglh(void *hv, double *out) {
  const auto &v = H(hv)->ann.fake_treeinfo->partition_loglh;
  std::copy(v.begin(), v.end(), out);
  return 1;
}

static void setBranchLength(AnnotatedNetwork &ann, int partition, unsigned edge, double value) {
  FakeTreeinfo &ti = *ann.fake_treeinfo;
  if (edge >= ann.network.num_branches()) throw std::runtime_error("branch index " + std::to_string(edge) + " out of range (the network has " + std::to_string(ann.network.num_branches()) + " branches)");
  if (!(value >= 0.0)) throw std::runtime_error("negative or NaN branch length");
  if (partition >= 0 && ti.brlen_linkage != PLLMOD_COMMON_BRLEN_UNLINKED)   // writing one partition's copy would desynchronise a linked length
    throw std::runtime_error("per-partition branch lengths exist only under unlinked branch-length linkage");
  if (partition >= (int)ti.partition_count) throw std::runtime_error("partition index out of range");
  if (partition < 0) {
    ti.linked_branch_lengths[edge] = value;
    for (auto &b : ti.branch_lengths) b[edge] = value;
  } else {
    ti.branch_lengths.at(partition)[edge] = value;
  }
}

int nrxh_set_branch_length(void *hv, int partition, unsigned edge, double value) {
  return guarded([&] { setBranchLength(H(hv)->ann, partition, edge, value); invalidatePmatrixIndex(H(hv)->ann, edge); });
}

int nrxh_set_reticulation_prob(void *hv, unsigned r, double prob) {
  return guarded([&] {
    AnnotatedNetwork &ann = H(hv)->ann;
    if (r >= ann.network.num_reticulations()) throw std::runtime_error("reticulation index " + std::to_string(r) + " out of range");
    if (!(prob >= ann.options.brprob_min && prob <= ann.options.brprob_max))   // NetraxOptions::brprob_min / brprob_max: log(0) otherwise
      throw std::runtime_error("reticulation probability outside [brprob_min, brprob_max]");
    setReticulationProb(ann, r, prob);
  });
}
int nrxh_set_params_to_optimize(void *hv, unsigned p, int mask) {
  return guarded([&] { H(hv)->ann.fake_treeinfo->partitions.at(p).params_to_optimize = mask; });
}

static void pushModel(AnnotatedNetwork &ann, unsigned p) { pushPartitionModel(ann, p); }

int nrxh_set_model(void *hv, unsigned p, const double *freqs, const double *subst, const double *rates, const double *rw) {
  return guarded([&] {
    AnnotatedNetwork &ann = H(hv)->ann;
    PartitionModel &m = ann.fake_treeinfo->partitions.at(p);
    set_frequencies(m, freqs);
    m.subst_params.assign(subst, subst + m.states * (m.states - 1) / 2);
    m.rates.assign(rates, rates + m.rate_cats);
    m.rate_weights.assign(rw, rw + m.rate_cats);
    m.eigen_decomp_valid = false;  // pll_set_* clear eigen_decomp_valid; recomputed before the next P-matrix update
    pushModel(ann, p);
  });
}

int nrxh_optimize_pinv(void *hv, double min_pinv, double max_pinv, double tolerance, double *final_logl) {
  return guarded([&] { const double l = optimize_pinv(H(hv)->ann, min_pinv, max_pinv, tolerance); if (final_logl) *final_logl = l; });
}
int nrxh_optimize_scalers(void *hv, double *bic_score) {
  return guarded([&] { const double b = optimize_scalers(H(hv)->ann, true); if (bic_score) *bic_score = b; });
}
int nrxh_get_brlen_scalers(void *hv, double *out) {
  return guarded([&] {
    const FakeTreeinfo &ti = *H(hv)->ann.fake_treeinfo;
    for (unsigned p = 0; p < ti.partition_count; ++p) out[p] = p < ti.brlen_scalers.size() ? ti.brlen_scalers[p] : 1.0;
  });
}
int nrxh_get_pinv(void *hv, unsigned p, double *out) {
  return guarded([&] { *out = H(hv)->ann.fake_treeinfo->partitions.at(p).prop_invar; });
}
int nrxh_set_brlen_scaler(void *hv, unsigned p, double scaler) {
  return guarded([&] { set_brlen_scaler(H(hv)->ann, p, scaler); });
}

int nrxh_set_submodels(void *hv, unsigned p, unsigned n, const unsigned *ratecat_submodels, const double *freqs, const double *subst) {
  return guarded([&] {
    AnnotatedNetwork &ann = H(hv)->ann;
    set_submodels(ann.fake_treeinfo->partitions.at(p), n, ratecat_submodels, freqs, subst);
    pushModel(ann, p);
  });
}

int nrxh_get_eigen(void *hv, unsigned p, double *ev, double *iev, double *evals) {
  return guarded([&] {
    const PartitionModel &m = H(hv)->ann.fake_treeinfo->partitions.at(p);
    std::copy(m.eigenvecs.begin(), m.eigenvecs.end(), ev);
    std::copy(m.inv_eigenvecs.begin(), m.inv_eigenvecs.end(), iev);
    std::copy(m.eigenvals.begin(), m.eigenvals.end(), evals);
  });
}

int nrxh_set_eigen(void *hv, unsigned p, const double *ev, const double *iev, const double *evals) {
  return guarded([&] {
    AnnotatedNetwork &ann = H(hv)->ann;
    PartitionModel &m = ann.fake_treeinfo->partitions.at(p);
    m.eigenvecs.assign(ev, ev + (size_t)m.states * m.states_padded);
    m.inv_eigenvecs.assign(iev, iev + (size_t)m.states * m.states_padded);
    m.eigenvals.assign(evals, evals + m.states_padded);
    m.eigen_decomp_valid = true;
    pushModel(ann, p);
  });
}

int nrxh_get_pmatrix(void *hv, unsigned p, unsigned edge, double *out) {
  return guarded([&] { detail::engineCheck(nrx_get_pmatrix(H(hv)->ann.engine, p, edge, out), "nrx_get_pmatrix"); });
}

int nrxh_brlen_prepare(void *hv, unsigned edge, double *old_logl) {
  return guarded([&] {
    Handle *h = H(hv);
    AnnotatedNetwork &ann = h->ann;
    if (edge >= ann.network.num_branches()) throw std::runtime_error("nrxh_brlen_prepare: edge out of range");
    if (detail::lazyRerootPossible(ann, edge) && !ann.root_clvs_stale) {
      // lazy re-rooting: no evaluation from the root, only the root-directed trees this re-rooting reads are brought up to date;
      // the old lnL handed out is the one the previous step returned
      if (old_logl) *old_logl = ann.lazy_last_logl;
      h->oldTrees.clear();
      ReticulationConfigSet restrictions = getRestrictionsActiveAliveBranch(ann, edge);
      detail::validateRerootInputs(ann, edge);
      updateCLVsVirtualRerootTrees(ann, ann.network.root, &ann.network.nodes[ann.network.edges[edge].source],
                                   &ann.network.nodes[ann.network.edges[edge].target], restrictions);
      ann.cached_logl_valid = false;
      return;
    }
    const double l = computeLoglikelihood(ann, 1, 1);
    ann.lazy_last_logl = l;
    if (old_logl) *old_logl = l;
    h->oldTrees = extractOldTrees(ann, ann.network.root);
    ReticulationConfigSet restrictions = getRestrictionsActiveAliveBranch(ann, edge);
    updateCLVsVirtualRerootTrees(ann, ann.network.root, &ann.network.nodes[ann.network.edges.at(edge).source],
                                 &ann.network.nodes[ann.network.edges.at(edge).target], restrictions);
    ann.cached_logl_valid = false;
  });
}

int nrxh_brlen_logl(void *hv, unsigned edge, double *out) {
  return guarded([&] {
    H(hv)->ann.cached_logl_valid = false;
    try { *out = computeLoglikelihoodBrlenOpt(H(hv)->ann, H(hv)->oldTrees, edge, 1); }
    catch (const LazyRerootNeedsRoot &) {
      redoRerootFromRoot(H(hv)->ann, edge, H(hv)->oldTrees);
      *out = computeLoglikelihoodBrlenOpt(H(hv)->ann, H(hv)->oldTrees, edge, 1);
    }
  });
}

int nrxh_brlen_logl_sumtables(void *hv, unsigned edge, double *out, unsigned *count) {
  return guarded([&] {
    Handle *h = H(hv);
    h->ann.cached_logl_valid = false;
    try { *out = computeLoglikelihoodBrlenOptAndSumtables(h->ann, h->oldTrees, edge, h->sumtables, 1); }
    catch (const LazyRerootNeedsRoot &) {
      redoRerootFromRoot(h->ann, edge, h->oldTrees);
      *out = computeLoglikelihoodBrlenOptAndSumtables(h->ann, h->oldTrees, edge, h->sumtables, 1);
    }
    if (count) *count = h->sumtables.empty() ? 0 : (unsigned)h->sumtables[0].size();
  });
}

int nrxh_brlen_sumtables(void *hv, unsigned edge, unsigned *count) {
  return guarded([&] {
    H(hv)->sumtables = computePartitionSumtables(H(hv)->ann, edge);
    if (count) *count = H(hv)->sumtables.empty() ? 0 : (unsigned)H(hv)->sumtables[0].size();
  });
}

int nrxh_brlen_read_sumtable(void *hv, unsigned p, unsigned idx, double *out, double *tree_prob, unsigned *lt, unsigned *rt) {
  return guarded([&] {
    const SumtableInfo &s = H(hv)->sumtables.at(p).at(idx);
    if (out) detail::engineCheck(nrx_read_sumtable(H(hv)->ann.engine, p, s.index, out), "nrx_read_sumtable");
    if (tree_prob) *tree_prob = s.tree_prob;
    if (lt) *lt = (unsigned)s.left_tree_idx;
    if (rt) *rt = (unsigned)s.right_tree_idx;
  });
}

int nrxh_brlen_set_length(void *hv, int partition, unsigned edge, double value) {
  return guarded([&] {
    AnnotatedNetwork &ann = H(hv)->ann;
    if (ann.options.brlen_linkage == PLLMOD_COMMON_BRLEN_UNLINKED && partition >= 0) ann.fake_treeinfo->branch_lengths.at(partition)[edge] = value;
    else setBranchLength(ann, -1, edge, value);
    invalidPmatrixIndexOnly(ann, edge);
  });
}

int nrxh_brlen_derivatives(void *hv, unsigned edge, double *d1, double *d2, double *pd1, double *pd2, double *raw) {
  return guarded([&] {
    Handle *h = H(hv);
    const LoglDerivatives r = computeLoglikelihoodDerivatives(h->ann, h->sumtables, edge);
    if (d1) *d1 = r.logl_prime;
    if (d2) *d2 = r.logl_prime_prime;
    if (pd1) std::copy(r.partition_logl_prime.begin(), r.partition_logl_prime.end(), pd1);
    if (pd2) std::copy(r.partition_logl_prime_prime.begin(), r.partition_logl_prime_prime.end(), pd2);
    if (raw) {
      const size_t n = h->sumtables.empty() ? 0 : h->sumtables[0].size();
      for (size_t p = 0; p < r.raw.size(); ++p)
        for (size_t k = 0; k < r.raw[p].size(); ++k) raw[p * 3 * n + k] = r.raw[p][k];
    }
  });
}

int nrxh_brlen_finish(void *hv, unsigned edge, double *final_logl) {
  return guarded([&] {
    Handle *h = H(hv);
    h->sumtables.clear();
    if (detail::rerootSessionIsLazy(h->ann)) {   // the edge-rooted lnL at the final length; whatever became invalid stays so until it is read
      h->ann.cached_logl_valid = false;
      double l;
      try { l = computeLoglikelihoodBrlenOpt(h->ann, h->oldTrees, edge, 1); }
      catch (const LazyRerootNeedsRoot &) { redoRerootFromRoot(h->ann, edge, h->oldTrees); l = computeLoglikelihoodBrlenOpt(h->ann, h->oldTrees, edge, 1); }
      h->sumtables.clear();
      finishVirtualReroot(h->ann);
      h->ann.lazy_last_logl = l;
      if (final_logl) *final_logl = l;
      return;
    }
    h->oldTrees.clear();
    finishVirtualReroot(h->ann);   // restores the root-directed trees; invalidates above the edge only if its length changed
    const double l = computeLoglikelihood(h->ann, 1, 1);
    h->ann.lazy_last_logl = l;
    if (final_logl) *final_logl = l;
  });
}

int nrxh_brlen_sweep_order(void *hv, unsigned *edges_out) {
  return guarded([&] {
    const std::vector<size_t> order = detail::branchesInPreorder(H(hv)->ann);
    for (size_t i = 0; i < order.size(); ++i) edges_out[i] = (unsigned)order[i];
  });
}

int nrxh_reroot_stats(void *hv, unsigned long long *hits, unsigned long long *misses, unsigned *entries, unsigned *cached_slots) {
  return guarded([&] {
    const AnnotatedNetwork &ann = H(hv)->ann;
    size_t n = 0, s = 0;
    detail::rerootCacheSize(ann, &n, &s);
    if (hits) *hits = ann.reroot_hits;
    if (misses) *misses = ann.reroot_misses;
    if (entries) *entries = (unsigned)n;
    if (cached_slots) *cached_slots = (unsigned)s;
  });
}

int nrxh_lazy_reroot_stats(void *hv, unsigned long long *sessions, unsigned long long *fallbacks) {
  return guarded([&] {
    if (sessions) *sessions = H(hv)->ann.lazy_sessions;
    if (fallbacks) *fallbacks = H(hv)->ann.lazy_fallbacks;
  });
}

int nrxh_set_reroot_cache_slots(void *hv, long long max_slots) {
  return guarded([&] {
    AnnotatedNetwork &ann = H(hv)->ann;
    ann.reroot_cache_max_slots = max_slots < 0 ? SIZE_MAX : (size_t)max_slots;
    if (max_slots == 0) dropRerootCache(ann);
  });
}

/* ---- the immediate callers: src/optimization/{BranchLengthOptimization,ReticulationOptimization}.cpp ---- */
int nrxh_optimize_branch(void *hv, unsigned edge, int method, unsigned max_iters, double *final_logl) {
  return guarded([&] {
    const double l = optimize_branch(H(hv)->ann, edge, (BrlenOptMethod)method, max_iters);
    if (final_logl) *final_logl = l;
  });
}

int nrxh_optimize_branches(void *hv, int max_iters, int max_iters_outside, int radius, int method, double *final_logl) {
  return guarded([&] {
    AnnotatedNetwork &ann = H(hv)->ann;
    ann.options.brlenOptMethod = (BrlenOptMethod)method;
    const double l = optimize_branches(ann, max_iters, max_iters_outside, radius);
    if (final_logl) *final_logl = l;
  });
}

int nrxh_optimize_reticulation(void *hv, unsigned r, double *final_logl) {
  return guarded([&] {
    const double l = optimize_reticulation(H(hv)->ann, r);
    if (final_logl) *final_logl = l;
  });
}

int nrxh_compute_pseudo_loglikelihood(void *hv, int incremental, int update_pmatrices, double *out) {
  return guarded([&] { *out = computePseudoLoglikelihood(H(hv)->ann, incremental, update_pmatrices); });
}
int nrxh_read_pseudo_clv(void *hv, unsigned node, unsigned p, double *out) {
  return guarded([&] {
    AnnotatedNetwork &ann = H(hv)->ann;
    if (node >= ann.pseudo_slot.size() || ann.pseudo_slot[node] == UINT32_MAX) throw std::runtime_error("no pseudo CLV at that node");
    detail::engineCheck(nrx_read_clv(ann.engine, p, ann.pseudo_slot[node], out), "nrx_read_clv");
  });
}
int nrxh_read_pseudo_scaler(void *hv, unsigned node, unsigned p, unsigned *out) {
  return guarded([&] {
    AnnotatedNetwork &ann = H(hv)->ann;
    if (node >= ann.pseudo_slot.size() || ann.pseudo_slot[node] == UINT32_MAX) throw std::runtime_error("no pseudo CLV at that node");
    detail::engineCheck(nrx_read_scaler(ann.engine, p, ann.pseudo_slot[node], out), "nrx_read_scaler");
  });
}
int nrxh_score_network(void *hv, double *bic_score) {
  return guarded([&] { *bic_score = scoreNetwork(H(hv)->ann); });
}
int nrxh_set_scoring_sizes(void *hv, unsigned long long total_num_model_parameters, unsigned long long total_num_sites) {
  return guarded([&] {
    H(hv)->ann.total_num_model_parameters = (size_t)total_num_model_parameters;
    if (total_num_sites) H(hv)->ann.total_num_sites = (size_t)total_num_sites;
  });
}
int nrxh_optimize_all_non_topology(void *hv, int type, double *bic_score) {
  return guarded([&] {
    optimizeAllNonTopology(H(hv)->ann, (OptimizeAllNonTopologyType)type);
    if (bic_score) *bic_score = scoreNetwork(H(hv)->ann);
  });
}
double nrxh_likelihood_target_function(void *network_params, int incremental, int update_pmatrices, double **persite_lnl) {
  try { return network_logl_wrapper(network_params, incremental, update_pmatrices, persite_lnl); }
  catch (const std::exception &e) { g_err = e.what(); return -std::numeric_limits<double>::infinity(); }
}
void *nrxh_network_params(void *hv) {  // the likelihood_computation_params pointer for the handle's network
  H(hv)->params.ann_network = &H(hv)->ann;
  return &H(hv)->params;
}
int nrxh_set_pinv(void *hv, unsigned p, double prop_invar) {
  return guarded([&] { setPinv(H(hv)->ann, p, prop_invar); });
}
int nrxh_set_alpha(void *hv, unsigned p, double alpha) {
  return guarded([&] { setAlpha(H(hv)->ann, p, alpha); });
}
int nrxh_get_alpha(void *hv, unsigned p, double *alpha) {
  return guarded([&] { *alpha = H(hv)->ann.fake_treeinfo->partitions.at(p).alpha; });
}
int nrxh_optimize_alpha(void *hv, double min_alpha, double max_alpha, double tolerance, double *final_logl) {
  return guarded([&] {
    const double l = optimize_alpha(H(hv)->ann, min_alpha, max_alpha, tolerance);
    if (final_logl) *final_logl = l;
  });
}
int nrxh_optimize_reticulations(void *hv, int max_iters, double *final_logl) {
  return guarded([&] {
    const double l = optimize_reticulations(H(hv)->ann, max_iters);
    if (final_logl) *final_logl = l;
  });
}

int nrxh_get_branch_lengths(void *hv, int partition, double *out) {
  return guarded([&] {
    const FakeTreeinfo &ti = *H(hv)->ann.fake_treeinfo;
    const std::vector<double> &b = (partition < 0 || ti.brlen_linkage != PLLMOD_COMMON_BRLEN_UNLINKED) ? ti.linked_branch_lengths : ti.branch_lengths.at(partition);
    std::copy(b.begin(), b.begin() + H(hv)->ann.network.num_branches(), out);
  });
}

int nrxh_get_reticulation_probs(void *hv, double *out) {
  return guarded([&] {
    const AnnotatedNetwork &ann = H(hv)->ann;
    std::copy(ann.reticulation_probs.begin(), ann.reticulation_probs.begin() + ann.network.num_reticulations(), out);
  });
}

unsigned long long nrxh_clv_update_count(void *hv) { return H(hv)->ann.clv_site_updates; }
void nrxh_reset_counters(void *hv) { H(hv)->ann.clv_site_updates = 0; }
int nrxh_gamma_rates(double alpha, unsigned cats, int mode, double *out) {
  if (!compute_gamma_cats(alpha, cats, out, mode)) { g_err = "Invalid alpha value / GAMMA discretization mode"; return 0; }
  return 1;
}
/* device-free: the host's own eigendecomposition of (frequencies, exchangeabilities) — what nrx_set_model receives */
int nrxh_eigen_decompose(unsigned states, const double *freqs, const double *subst, double *ev, double *iev, double *evals) {
  return guarded([&] {
    if (states < 2 || states > 32) throw std::runtime_error("nrxh_eigen_decompose: states must be in 2..32");
    PartitionModel m;
    m.states = states;
    m.states_padded = (states + 3) & ~3u;
    set_frequencies(m, freqs);
    m.subst_params.assign(subst, subst + states * (states - 1) / 2);
    update_eigen(m);
    std::copy(m.eigenvecs.begin(), m.eigenvecs.end(), ev);
    std::copy(m.inv_eigenvecs.begin(), m.inv_eigenvecs.end(), iev);
    std::copy(m.eigenvals.begin(), m.eigenvals.end(), evals);
  });
}
/* device-free: the host's restatements of pll-modules' 1-D minimisers, the ones optimize_branch / optimize_reticulation run */
int nrxh_minimize_newton(double xmin, double *x, double xmax, double tolerance, unsigned max_iters,
                         void (*deriv)(void *, double *, double *, double *), void *ctx, int *converged) {
  return guarded([&] { const bool ok = netrax::detail::minimizeNewton(xmin, x, xmax, tolerance, max_iters, deriv, ctx); if (converged) *converged = ok ? 1 : 0; });
}
int nrxh_minimize_brent(double xmin, double xguess, double xmax, double xtol, double (*target)(void *, double), void *ctx, double *xopt) {
  return guarded([&] { *xopt = netrax::detail::minimizeBrent(xmin, xguess, xmax, xtol, target, ctx); });
}
int nrxh_minimize_brent_multi(unsigned n, double xmin, double *x, double xmax, double xtol,
                              double (*target)(void *, double *, double *, int *), void *ctx) {
  return guarded([&] { netrax::detail::minimizeBrentMulti(n, xmin, x, xmax, xtol, target, ctx); });
}
unsigned long long nrxh_launch_count(void *hv) { return nrx_launch_count(H(hv)->ann.engine); }
unsigned nrxh_num_slots(void *hv) { return H(hv)->ann.next_slot; }
int nrxh_profile_enable(void *hv, int on) { return nrx_profile_enable(H(hv)->ann.engine, on); }
int nrxh_profile_read(void *hv, double *ms, unsigned long long *l, unsigned long long *u, unsigned long long *b) {
  if (!nrx_profile_read(H(hv)->ann.engine, ms, l, u, b)) { g_err = nrx_last_error(); return 0; }
  return 1;
}
int nrxh_profile_read_kind(void *hv, int kind, double *ms, unsigned long long *l, unsigned long long *u, unsigned long long *b, unsigned long long *cb) {
  if (!nrx_profile_read_kind(H(hv)->ann.engine, kind, ms, l, u, b, cb)) { g_err = nrx_last_error(); return 0; }
  return 1;
}
int nrxh_persite_lnl(void *hv, unsigned tree, double *out, unsigned stride) {
  return guarded([&] {
    AnnotatedNetwork &ann = H(hv)->ann;
    detail::flushPendingOps(ann);
    NodeDisplayedTreeData &rd = ann.pernode_displayed_tree_data[ann.network.root->clv_index];
    DisplayedTreeData &t = rd.displayed_trees.at(tree);
    detail::setReticulationParents(ann.network, t.treeLoglData.reticulationChoices.configs[0]);
    size_t dtr = ann.network.root->clv_index;
    detail::collect_dead_nodes(ann.network, dtr, &dtr);
    NodeDisplayedTreeData &dd = ann.pernode_displayed_tree_data[dtr];
    uint32_t slot = UINT32_MAX;
    for (size_t k = 0; k < dd.num_active_displayed_trees; ++k)
      if (reticulationConfigsCompatible(t.treeLoglData.reticulationChoices, dd.displayed_trees[k].treeLoglData.reticulationChoices)) slot = dd.displayed_trees[k].slot;
    if (slot == UINT32_MAX) throw std::runtime_error("Found no suitable displayed tree");
    std::vector<double> tmp(ann.fake_treeinfo->partition_count);
    detail::engineCheck(nrx_tree_lnl(ann.engine, &slot, 1, tmp.data(), out, stride), "nrx_tree_lnl");
  });
}
void *nrxh_engine(void *hv) { return H(hv)->ann.engine; }
int nrxh_upload_alignment_u8(void *hv, unsigned p, const uint8_t *tipchars, const unsigned *pw) {
  return guarded([&] {   // asynchronous: the buffers are borrowed until the next evaluation has been collected
    AnnotatedNetwork &ann = H(hv)->ann;
    detail::engineCheck(nrx_set_tipchars_u8(ann.engine, p, tipchars), "nrx_set_tipchars_u8");
    if (pw) detail::engineCheck(nrx_set_pattern_weights_async(ann.engine, p, pw), "nrx_set_pattern_weights");
    invalidateAllCLVs(ann);
  });
}
int nrxh_stage_alignment_u8(void *hv, unsigned p, const uint8_t *tipchars, const unsigned *pw) {
  return guarded([&] { detail::engineCheck(nrx_stage_alignment_u8(H(hv)->ann.engine, p, tipchars, pw), "nrx_stage_alignment_u8"); });
}
int nrxh_commit_staged_alignment(void *hv) {
  return guarded([&] {
    AnnotatedNetwork &ann = H(hv)->ann;
    finishVirtualReroot(ann);
    detail::engineCheck(nrx_commit_staged_alignment(ann.engine), "nrx_commit_staged_alignment");
    invalidateAllCLVs(ann);
  });
}
int nrxh_upload_alignment_codes(void *hv, unsigned p, const uint8_t *codes, const uint32_t *tipmap, unsigned ncodes, const unsigned *pw) {
  return guarded([&] {   // any alphabet: code c = state set tipmap[c]
    AnnotatedNetwork &ann = H(hv)->ann;
    detail::engineCheck(nrx_set_tipcodes_u8(ann.engine, p, codes, tipmap, ncodes), "nrx_set_tipcodes_u8");
    if (pw) detail::engineCheck(nrx_set_pattern_weights_async(ann.engine, p, pw), "nrx_set_pattern_weights");
    invalidateAllCLVs(ann);
  });
}
int nrxh_timer_start(void *hv) { if (!nrx_timer_start(H(hv)->ann.engine)) { g_err = nrx_last_error(); return 0; } return 1; }
int nrxh_timer_stop(void *hv, double *ms) { if (!nrx_timer_stop(H(hv)->ann.engine, ms)) { g_err = nrx_last_error(); return 0; } return 1; }

}  // extern "C"
