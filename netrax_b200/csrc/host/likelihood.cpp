/*
 * likelihood.cpp — computeLoglikelihood and friends on the device engine.
 *
 * Control flow and semantics follow src/likelihood/ImprovedLoglikelihood.cpp, src/helper/InvalidationHelper.cpp,
 * src/helper/ReticulationConfigHelper.cpp and src/graph/AnnotatedNetwork.cpp:340-394 of the reference (cited per
 * function).  The arithmetic is NOT here: processNodeImproved emits nrx_op records (one per displayed tree the
 * reference would compute with one pll_update_partials_single per partition) and all trees of a node — and of
 * further independent nodes — go to the GPU in one nrx_update_clvs launch.
 */
#include <algorithm>
#include <cmath>
#include <cstring>
#include <exception>

#include "host_internal.hpp"

namespace netrax {

/* An evaluation plan = the launches of one full (non-incremental) traversal.  The enumeration of displayed
 * trees depends only on the topology, so as long as topology_changed() is not called the host re-issues the
 * recorded launches instead of redoing the O(4^k) config-set algebra per node. */
struct PlanCache {
  bool valid = false, recording = false;
  std::vector<std::vector<nrx_op>> batches;
  struct NodeSnap { std::vector<ReticulationConfigSet> configs; std::vector<uint32_t> slots; };
  std::vector<NodeSnap> nodes;
  uint64_t site_updates = 0;
  bool on_device = false;   // the batches live on the device as an engine plan (one CUDA-graph launch per replay)
  uint32_t engine_plan = 0;
  // fused K3: the root displayed trees' slots at recording time; their producing ops carry lnl_item = index + 1
  bool fused = false;
  std::vector<uint32_t> lnl_slots;
  bool run_pending = false;   // a replay whose device launches are issued together with the root lnLs (nrx_plan_evaluate_async)
  // which root displayed trees are evaluated and where their CLVs live (findFirstNodeWithTwoActiveChildren + findMatchingDisplayedTree
  // per tree: O(trees x (nodes + trees)) of host algebra, 192 trees on BASELINE config 5) — a function of the topology and, through
  // the min_interesting_tree_logprob filter, of the reticulation probabilities: kept per plan while `prob_epoch` stands, so that a
  // replay issues its launches at once instead of leaving the GPU idle behind that loop
  bool root_cached = false;
  uint64_t root_prob_epoch = 0;
  std::vector<uint32_t> root_slots;
  std::vector<size_t> root_which;
};

AnnotatedNetwork::~AnnotatedNetwork() {
  if (plan && plan->on_device && engine) nrx_plan_destroy(engine, plan->engine_plan);
  delete plan;
  detail::destroyRerootCache(*this);
  if (engine) nrx_engine_destroy(engine);
}

namespace detail {

void engineCheck(int ok, const char *what) {
  if (!ok) throw std::runtime_error(std::string(what) + ": " + nrx_last_error());
}

uint32_t allocSlot(AnnotatedNetwork &ann) {
  if (!ann.free_slots.empty()) { uint32_t s = ann.free_slots.back(); ann.free_slots.pop_back(); return s; }
  const uint32_t s = ann.next_slot++;
  if (s >= nrx_num_slots(ann.engine)) {
    const uint32_t want = std::max<uint32_t>(s + 1, nrx_num_slots(ann.engine) + std::max<uint32_t>(16, nrx_num_slots(ann.engine) / 2));
    if (!nrx_reserve_slots(ann.engine, want) && !nrx_reserve_slots(ann.engine, s + 1)) {
      // out of device memory: the memoised re-rooted trees are only a cache — give their slots back and take one of those
      const std::string err = nrx_last_error();
      ann.next_slot--;
      dropRerootCache(ann);
      if (ann.free_slots.empty()) throw std::runtime_error("nrx_reserve_slots: " + err);
      const uint32_t r = ann.free_slots.back();
      ann.free_slots.pop_back();
      return r;
    }
  }
  return s;
}

void releaseSlot(AnnotatedNetwork &ann, uint32_t slot) { ann.free_slots.push_back(slot); }

void flushPendingOps(AnnotatedNetwork &ann) {
  if (ann.pending_ops.empty()) return;
  engineCheck(nrx_update_clvs(ann.engine, ann.pending_ops.data(), (uint32_t)ann.pending_ops.size()), "nrx_update_clvs");
  uint64_t local_sites = 0;
  for (const PartitionModel &m : ann.fake_treeinfo->partitions) local_sites += m.sites;
  ann.clv_site_updates += local_sites * ann.pending_ops.size();
  if (ann.plan && ann.plan->recording) { ann.plan->batches.push_back(ann.pending_ops); ann.plan->site_updates += local_sites * ann.pending_ops.size(); }
  ann.pending_ops.clear();
  std::fill(ann.pending_parent.begin(), ann.pending_parent.end(), 0);
}

void reduceSum(AnnotatedNetwork &ann, double *data, size_t count) {
  // with an NCCL communicator attached the engine has already all-reduced its result on the device (C2-C4)
  if (ann.engine && nrx_comm_size(ann.engine) > 1) return;
  if (ann.fake_treeinfo->parallel_reduce_cb)
    ann.fake_treeinfo->parallel_reduce_cb(ann.fake_treeinfo->parallel_context, data, count, PLLMOD_COMMON_REDUCE_SUM);
}

/* SUM over site shards of values that live on the HOST (the Brent drivers' convergence flag, the scaler normalisation's
 * weight sums): through the engine's NCCL communicator when one is attached, else through parallel_reduce_cb */
void reduceHostSum(AnnotatedNetwork &ann, double *data, size_t count) {
  if (ann.engine && nrx_comm_size(ann.engine) > 1) engineCheck(nrx_comm_allreduce_sum(ann.engine, data, count), "nrx_comm_allreduce_sum");
  else if (ann.fake_treeinfo->parallel_reduce_cb)
    ann.fake_treeinfo->parallel_reduce_cb(ann.fake_treeinfo->parallel_context, data, count, PLLMOD_COMMON_REDUCE_SUM);
}

double logSumExp(const std::vector<double> &a) {
  double m = -std::numeric_limits<double>::infinity();
  for (double x : a) m = std::max(m, x);
  if (m == -std::numeric_limits<double>::infinity()) return m;
  double s = 0.0;
  for (double x : a) s += std::exp(x - m);
  return m + std::log(s);
}

/* ---- src/helper/ReticulationConfigHelper.cpp ----------------------------------------------------------- */
ReticulationConfigSet getRestrictionsToDismissNeighbor(AnnotatedNetwork &ann, size_t node, size_t neighbor) {  // :26-63
  const Network &nw = ann.network;
  ReticulationConfigSet res(ann.options.max_reticulations);
  ReticulationConfig r;
  auto fix = [&](size_t ret, bool second) { r.care |= 1u << ret; r.second = (r.second & ~(1u << ret)) | ((second ? 1u : 0u) << ret); };
  bool found = false;
  if (nw.nodes[node].type == NodeType::RETICULATION_NODE) {
    const ReticulationInfo &R = nw.reticulations[nw.nodes[node].reticulation_index];
    if (neighbor == R.first_parent) { fix(nw.nodes[node].reticulation_index, true); found = true; }
    else if (neighbor == R.second_parent) { fix(nw.nodes[node].reticulation_index, false); found = true; }
  }
  if (nw.nodes[neighbor].type == NodeType::RETICULATION_NODE) {
    const ReticulationInfo &R = nw.reticulations[nw.nodes[neighbor].reticulation_index];
    if (node == R.first_parent) { fix(nw.nodes[neighbor].reticulation_index, true); found = true; }
    else if (node == R.second_parent) { fix(nw.nodes[neighbor].reticulation_index, false); found = true; }
  }
  if (found) res.configs.push_back(r);
  return res;
}

ReticulationConfigSet getRestrictionsToTakeNeighbor(AnnotatedNetwork &ann, size_t node, size_t neighbor) {  // :65-96
  const Network &nw = ann.network;
  ReticulationConfigSet res(ann.options.max_reticulations);
  ReticulationConfig r;
  auto fix = [&](size_t ret, bool second) { r.care |= 1u << ret; r.second = (r.second & ~(1u << ret)) | ((second ? 1u : 0u) << ret); };
  if (nw.nodes[node].type == NodeType::RETICULATION_NODE) {
    const ReticulationInfo &R = nw.reticulations[nw.nodes[node].reticulation_index];
    if (neighbor == R.first_parent) fix(nw.nodes[node].reticulation_index, false);
    else if (neighbor == R.second_parent) fix(nw.nodes[node].reticulation_index, true);
  }
  if (nw.nodes[neighbor].type == NodeType::RETICULATION_NODE) {
    const ReticulationInfo &R = nw.reticulations[nw.nodes[neighbor].reticulation_index];
    if (node == R.first_parent) fix(nw.nodes[neighbor].reticulation_index, false);
    else if (node == R.second_parent) fix(nw.nodes[neighbor].reticulation_index, true);
  }
  res.configs.push_back(r);
  return res;
}

ReticulationConfigSet getTreeConfig(AnnotatedNetwork &ann, size_t tree_idx) {  // :196-211
  ReticulationConfigSet c(ann.options.max_reticulations);
  const size_t R = ann.network.num_reticulations();
  const uint32_t all = R >= 32 ? 0xffffffffu : ((1u << R) - 1);
  c.configs.push_back(ReticulationConfig{all, (uint32_t)tree_idx & all});
  return c;
}

bool isActiveBranch(AnnotatedNetwork &ann, const ReticulationConfigSet &rc, size_t pmatrix_index) {  // EdgeHelper.cpp:84-95
  const Edge &E = ann.network.edges[pmatrix_index];
  return reticulationConfigsCompatible(getRestrictionsToTakeNeighbor(ann, E.source, E.target), rc);
}

bool isActiveAliveBranch(AnnotatedNetwork &ann, const ReticulationConfigSet &rc, size_t pmatrix_index) {  // EdgeHelper.cpp:97-121
  setReticulationParents(ann.network, rc.configs[0]);
  const std::vector<char> dead = collect_dead_nodes(ann.network, ann.network.root->clv_index, nullptr);
  const Edge &E = ann.network.edges[pmatrix_index];
  return reticulationConfigsCompatible(getRestrictionsToTakeNeighbor(ann, E.source, E.target), rc) && !dead[E.source] && !dead[E.target];
}

bool clvValidCheck(AnnotatedNetwork &ann, size_t vroot, bool care_about_trees) {  // AnnotatedNetwork.cpp:340-357
  if (care_about_trees && ann.pernode_displayed_tree_data[vroot].num_active_displayed_trees == 0) return false;
  bool ok = true;
  for (unsigned p = 0; p < ann.fake_treeinfo->partition_count; ++p) ok &= (bool)ann.fake_treeinfo->clv_valid[p][vroot];
  return ok;
}

nrx_pair makePair(const DisplayedTreeData &a, const DisplayedTreeData &b) {
  nrx_pair q;
  q.a_kind = a.isTip ? NRX_TIP : NRX_CLV; q.a_idx = a.isTip ? a.tip : a.slot;
  q.b_kind = b.isTip ? NRX_TIP : NRX_CLV; q.b_idx = b.isTip ? b.tip : b.slot;
  return q;
}

}  // namespace detail

using namespace detail;

static void refreshLogprob(AnnotatedNetwork &ann, TreeLoglData &t) {
  if (!t.tree_logprob_valid) {
    t.tree_logprob = computeReticulationConfigLogProb(t.reticulationChoices, ann.first_parent_logprobs, ann.second_parent_logprobs);
    t.tree_logprob_valid = true;
  }
}

/* ---- src/helper/InvalidationHelper.cpp -------------------------------------------------------------------- */
void invalidateSingleClv(AnnotatedNetwork &ann, unsigned int clv_index) {  // :8-39
  for (unsigned p = 0; p < ann.fake_treeinfo->partition_count; ++p) ann.fake_treeinfo->clv_valid[p][clv_index] = 0;
  NodeDisplayedTreeData &nd = ann.pernode_displayed_tree_data[clv_index];
  for (size_t i = 0; i < nd.num_active_displayed_trees; ++i) {
    nd.displayed_trees[i].clv_valid = false;
    nd.displayed_trees[i].treeLoglData.tree_logl_valid = false;
  }
  nd.num_active_displayed_trees = 0;
  if (clv_index < ann.pseudo_clv_valid.size()) ann.pseudo_clv_valid[clv_index] = 0;  // :37
  ann.cached_logl_valid = false;
}

static void validateSingleClv(AnnotatedNetwork &ann, size_t clv_index) {  // :41-50
  for (unsigned p = 0; p < ann.fake_treeinfo->partition_count; ++p) ann.fake_treeinfo->clv_valid[p][clv_index] = 1;
}

static void invalidateHigher(AnnotatedNetwork &ann, size_t node, bool invalidate_myself) {  // :52-87
  Network &nw = ann.network;
  if (node == SIZE_MAX) return;
  if (node < nw.num_tips()) invalidate_myself = false;
  if (invalidate_myself) invalidateSingleClv(ann, (unsigned)node);
  if (node == nw.root->clv_index) return;
  if (nw.nodes[node].type == NodeType::RETICULATION_NODE) {
    invalidateHigher(ann, nw.reticulations[nw.nodes[node].reticulation_index].first_parent, true);
    invalidateHigher(ann, nw.reticulations[nw.nodes[node].reticulation_index].second_parent, true);
  } else {
    invalidateHigher(ann, activeParent(nw, node), true);
  }
  ann.cached_logl_valid = false;
}

void invalidateHigherCLVs(AnnotatedNetwork &ann, const Node *node, bool invalidate_myself) {
  invalidateHigher(ann, node ? node->clv_index : SIZE_MAX, invalidate_myself);
}

void invalidatePmatrixIndex(AnnotatedNetwork &ann, size_t pmatrix_index) {  // :157-177
  for (unsigned p = 0; p < ann.fake_treeinfo->partition_count; ++p) ann.fake_treeinfo->pmatrix_valid[p][pmatrix_index] = 0;
  invalidateHigher(ann, ann.network.edges[pmatrix_index].source, true);
}

void invalidPmatrixIndexOnly(AnnotatedNetwork &ann, size_t pmatrix_index) {  // :179-192
  for (unsigned p = 0; p < ann.fake_treeinfo->partition_count; ++p) ann.fake_treeinfo->pmatrix_valid[p][pmatrix_index] = 0;
  ann.cached_logl_valid = false;
}

bool allClvsValid(AnnotatedNetwork &ann, size_t clv_index) {  // :194-237
  for (unsigned p = 0; p < ann.fake_treeinfo->partition_count; ++p)
    if (!ann.fake_treeinfo->clv_valid[p][clv_index]) return false;
  NodeDisplayedTreeData &nd = ann.pernode_displayed_tree_data[clv_index];
  if (nd.num_active_displayed_trees == 0) return false;
  for (size_t i = 0; i < nd.num_active_displayed_trees; ++i) {
    DisplayedTreeData &dtd = nd.displayed_trees[i];
    refreshLogprob(ann, dtd.treeLoglData);
    if (!dtd.clv_valid && dtd.treeLoglData.tree_logprob < ann.options.min_interesting_tree_logprob) return false;
  }
  return true;
}

void invalidateAllCLVs(AnnotatedNetwork &ann) {  // :255-260
  ann.clv_epoch++;
  for (size_t i = ann.network.num_tips(); i < ann.network.num_nodes(); ++i) invalidateSingleClv(ann, (unsigned)i);
}

void invalidateTreeLogprobs(AnnotatedNetwork &ann) {  // :273-308
  uint32_t all = 0;
  for (size_t r = 0; r < ann.network.num_reticulations(); ++r) all |= 1u << r;
  for (NodeDisplayedTreeData &nd : ann.pernode_displayed_tree_data)
    for (size_t j = 0; j < nd.num_active_displayed_trees; ++j) {
      TreeLoglData &t = nd.displayed_trees[j].treeLoglData;
      bool has = false;
      for (const ReticulationConfig &c : t.reticulationChoices.configs) has |= (c.care & all) != 0;
      if (has) { t.tree_logprob_valid = false; refreshLogprob(ann, t); }
    }
}

void setReticulationProb(AnnotatedNetwork &ann, size_t r, double prob) {  // src/optimization/ReticulationOptimization.cpp:25-38
  if (ann.reticulation_probs[r] == prob) return;
  ann.reticulation_probs[r] = prob;
  ann.first_parent_logprobs[r] = std::log(prob);
  ann.second_parent_logprobs[r] = std::log(1.0 - prob);
  ann.cached_logl_valid = false;
  ann.prob_epoch++;  // which root trees are interesting (min_interesting_tree_logprob) may change
  ann.clv_epoch++;   // memoised re-rooted trees carry their tree_logprob
  invalidateTreeLogprobs(ann);
  if (ann.options.likelihood_variant == LikelihoodVariant::SARAH_PSEUDO)  // InvalidationHelper.cpp:296-301: the blend weights changed
    invalidateHigher(ann, ann.network.reticulations[r].node, false);
}

static void dropEnginePlan(AnnotatedNetwork &ann) {
  if (ann.plan && ann.plan->on_device) { nrx_plan_destroy(ann.engine, ann.plan->engine_plan); ann.plan->on_device = false; }
}

void topology_changed(AnnotatedNetwork &ann) {
  if (ann.plan) { ann.plan->valid = false; ann.plan->root_cached = false; dropEnginePlan(ann); }
  ann.clv_epoch++;
  ann.topology_epoch++;
  ann.node_version.clear();
  ann.travbuffer = reversed_topological_sort(ann.network);
}

/* ---- set-up --------------------------------------------------------------------------------------------- */
static void uploadModel(AnnotatedNetwork &ann, unsigned p) {
  PartitionModel &m = ann.fake_treeinfo->partitions[p];
  if (!m.eigen_decomp_valid) update_eigen(m);
  const size_t S = m.states, SP = m.states_padded, M = 1 + m.submodels.size();
  std::vector<double> freqs(M * SP, 0.0);
  std::copy(m.frequencies.begin(), m.frequencies.begin() + S, freqs.begin());
  if (M == 1) {
    engineCheck(nrx_set_model(ann.engine, p, freqs.data(), m.eigenvecs.data(), m.inv_eigenvecs.data(), m.eigenvals.data(),
                              m.rates.data(), m.rate_weights.data(), m.prop_invar), "nrx_set_model");
    return;
  }
  // one rate matrix per category (LG4M / LG4X): the matrices back to back, category -> matrix map = pllmod's param_indices
  std::vector<double> ev(m.eigenvecs), iev(m.inv_eigenvecs), evals(m.eigenvals);
  evals.resize(SP, 0.0);
  for (size_t i = 0; i + 1 < M; ++i) {
    const PartitionModel::SubModel &sm = m.submodels[i];
    std::copy(sm.frequencies.begin(), sm.frequencies.begin() + S, freqs.begin() + (i + 1) * SP);
    ev.insert(ev.end(), sm.eigenvecs.begin(), sm.eigenvecs.end());
    iev.insert(iev.end(), sm.inv_eigenvecs.begin(), sm.inv_eigenvecs.end());
    evals.insert(evals.end(), sm.eigenvals.begin(), sm.eigenvals.end());
    evals.resize((i + 2) * SP, 0.0);
  }
  std::vector<uint32_t> cm(m.ratecat_submodels.begin(), m.ratecat_submodels.end());
  engineCheck(nrx_set_model_mixture(ann.engine, p, (uint32_t)M, cm.data(), freqs.data(), ev.data(), iev.data(), evals.data(),
                                    m.rates.data(), m.rate_weights.data(), m.prop_invar), "nrx_set_model_mixture");
}

void set_brlen_scaler(AnnotatedNetwork &ann, unsigned p, double scaler) {
  FakeTreeinfo &ti = *ann.fake_treeinfo;
  if (ti.brlen_linkage != PLLMOD_COMMON_BRLEN_SCALED) throw std::runtime_error("Branch length scalers exist only in scaled branch length mode.");
  if (ti.brlen_scalers.size() < ti.partition_count) ti.brlen_scalers.resize(ti.partition_count, 1.0);
  ti.brlen_scalers.at(p) = scaler;
  for (auto &v : ti.pmatrix_valid.at(p)) v = 0;
  invalidateAllCLVs(ann);
}

void pushPartitionModel(AnnotatedNetwork &ann, unsigned p) {
  uploadModel(ann, p);
  for (auto &v : ann.fake_treeinfo->pmatrix_valid.at(p)) v = 0;
  invalidateAllCLVs(ann);
}

void init_annotated_network(AnnotatedNetwork &ann, const std::vector<PartitionInput> &parts, int device) {
  Network &nw = ann.network;
  FakeTreeinfo &ti = *ann.fake_treeinfo;
  const unsigned P = (unsigned)parts.size();
  if (nw.num_reticulations() > ann.options.max_reticulations) throw std::runtime_error("number of reticulations in the network is higher than maximum reticulation count");
  ti.partition_count = P;
  ti.brlen_linkage = ann.options.brlen_linkage;
  ti.partitions.clear();
  std::vector<nrx_partition_desc> descs(P);
  for (unsigned p = 0; p < P; ++p) {
    PartitionModel m = parts[p].model;
    m.states_padded = (m.states + 3) & ~3u;
    if (m.rate_weights.empty()) m.rate_weights.assign(m.rate_cats, 1.0 / m.rate_cats);
    ti.partitions.push_back(m);
    descs[p] = nrx_partition_desc{m.states, m.rate_cats, m.sites, (uint32_t)nw.num_tips(), (uint32_t)nw.num_branches() + 1};
  }
  ann.pattern_weight_sums.assign(P, 0.0);   // pll_partition_t::pattern_weight_sum of this shard
  for (unsigned p = 0; p < P; ++p)
    for (unsigned i = 0; i < parts[p].model.sites; ++i) ann.pattern_weight_sums[p] += parts[p].pattern_weights ? parts[p].pattern_weights[i] : 1;
  if (ann.total_num_sites == 0)
    for (unsigned p = 0; p < P; ++p) ann.total_num_sites += (size_t)ann.pattern_weight_sums[p];
  ann.engine = nrx_engine_create(descs.data(), P, device);
  if (!ann.engine) throw std::runtime_error(std::string("nrx_engine_create: ") + nrx_last_error());
  for (unsigned p = 0; p < P; ++p) {
    engineCheck(nrx_set_tips(ann.engine, p, parts[p].tip_masks), "nrx_set_tips");
    if (parts[p].pattern_weights) engineCheck(nrx_set_pattern_weights(ann.engine, p, parts[p].pattern_weights), "nrx_set_pattern_weights");
    uploadModel(ann, p);
  }
  ann.travbuffer = reversed_topological_sort(nw);
  ann.reticulation_probs.assign(ann.options.max_reticulations, 0.5);
  ann.first_parent_logprobs.assign(ann.options.max_reticulations, std::log(0.5));
  ann.second_parent_logprobs.assign(ann.options.max_reticulations, std::log(0.5));
  for (size_t i = 0; i < nw.num_reticulations(); ++i) {  // AnnotatedNetwork.cpp:89-101
    const double pr = nw.edges[nw.reticulations[i].first_edge].prob;
    ann.reticulation_probs[i] = pr;
    ann.first_parent_logprobs[i] = std::log(pr);
    ann.second_parent_logprobs[i] = std::log(1.0 - pr);
  }
  ti.clv_valid.assign(P, std::vector<char>(nw.num_nodes(), 0));
  ti.pmatrix_valid.assign(P, std::vector<char>(nw.num_branches() + 1, 0));
  ti.linked_branch_lengths.assign(nw.num_branches() + 1, 0.0);  // fake branch: length 0 (RaxmlWrapper.cpp:514-524)
  for (size_t e = 0; e < nw.num_branches(); ++e) ti.linked_branch_lengths[e] = nw.edges[e].length;
  if (ti.branch_lengths.size() != P) ti.branch_lengths.assign(P, ti.linked_branch_lengths);
  for (auto &b : ti.branch_lengths) b.resize(nw.num_branches() + 1, 0.0);
  ti.partition_loglh.assign(P, 0.0);
  pllmod_treeinfo_update_prob_matrices(ann, 1);
  for (unsigned p = 0; p < P; ++p) for (size_t j = 0; j < nw.num_tips(); ++j) ti.clv_valid[p][j] = 1;
  ann.pernode_displayed_tree_data.assign(nw.num_nodes(), NodeDisplayedTreeData());
  for (size_t i = 0; i < nw.num_tips(); ++i) {  // tips: one displayed tree, no CLV (PATTERN_TIP)
    DisplayedTreeData d;
    d.treeLoglData = TreeLoglData(P, ann.options.max_reticulations);
    d.treeLoglData.reticulationChoices.configs.push_back(ReticulationConfig{});
    d.isTip = true; d.tip = (uint32_t)i; d.clv_valid = true;
    ann.pernode_displayed_tree_data[i].displayed_trees.push_back(d);
    ann.pernode_displayed_tree_data[i].num_active_displayed_trees = 1;
  }
  ann.pending_parent.assign(nw.num_nodes(), 0);
  if (!ann.plan) ann.plan = new PlanCache();
  ann.cached_logl_valid = false;
}

int pllmod_treeinfo_update_prob_matrices(AnnotatedNetwork &ann, int update_all) {  // PLLMOD/tree/treeinfo.c:842-880
  FakeTreeinfo &ti = *ann.fake_treeinfo;
  for (unsigned p = 0; p < ti.partition_count; ++p) {
    std::vector<uint32_t> idx;
    std::vector<double> len;
    for (size_t m = 0; m < ti.pmatrix_valid[p].size(); ++m) {
      if (ti.pmatrix_valid[p][m] && !update_all) continue;
      idx.push_back((uint32_t)m);
      double p_brlen = ti.branch_lengths[p][m];
      if (ti.brlen_linkage == PLLMOD_COMMON_BRLEN_SCALED && p < ti.brlen_scalers.size()) p_brlen *= ti.brlen_scalers[p];  // :862-864
      len.push_back(p_brlen);
      ti.pmatrix_valid[p][m] = 1;
    }
    if (!idx.empty()) engineCheck(nrx_update_pmatrices(ann.engine, p, (uint32_t)idx.size(), idx.data(), len.data()), "nrx_update_pmatrices");
  }
  return 1;
}

/* ---- src/likelihood/ImprovedLoglikelihood.cpp ------------------------------------------------------------- */
static DisplayedTreeData *findDisplayedTree(AnnotatedNetwork &ann, size_t clv_index, const ReticulationConfigSet &rc) {  // :15-28
  NodeDisplayedTreeData &nd = ann.pernode_displayed_tree_data[clv_index];
  for (size_t i = 0; i < nd.num_active_displayed_trees; ++i)
    if (nd.displayed_trees[i].treeLoglData.reticulationChoices == rc) return &nd.displayed_trees[i];
  return nullptr;
}

static bool tree_already_present_and_fine(AnnotatedNetwork &ann, size_t clv_index, const ReticulationConfigSet &rc) {  // :30-57
  DisplayedTreeData *dtd = findDisplayedTree(ann, clv_index, rc);
  if (!dtd) return false;
  refreshLogprob(ann, dtd->treeLoglData);
  return dtd->clv_valid || dtd->treeLoglData.tree_logprob < ann.options.min_interesting_tree_logprob;
}

static DisplayedTreeData &add_displayed_tree(AnnotatedNetwork &ann, size_t clv_index, const ReticulationConfigSet &rc) {  // :59-113
  if (DisplayedTreeData *dtd = findDisplayedTree(ann, clv_index, rc)) return *dtd;
  NodeDisplayedTreeData &data = ann.pernode_displayed_tree_data[clv_index];
  data.num_active_displayed_trees++;
  if (data.num_active_displayed_trees > data.displayed_trees.size()) {
    DisplayedTreeData d;
    d.treeLoglData = TreeLoglData(ann.fake_treeinfo->partition_count, ann.options.max_reticulations);
    d.slot = allocSlot(ann);  // the slot stays with this entry for the lifetime of the network (reference: vector entry keeps its buffers)
    data.displayed_trees.push_back(d);
  }
  DisplayedTreeData &tree = data.displayed_trees[data.num_active_displayed_trees - 1];
  tree.clv_valid = false;
  tree.treeLoglData.reticulationChoices = rc;
  tree.treeLoglData.tree_logprob = computeReticulationConfigLogProb(rc, ann.first_parent_logprobs, ann.second_parent_logprobs);
  tree.treeLoglData.tree_logprob_valid = false;  // as in the reference: value set, flag untouched -> recomputed lazily
  tree.treeLoglData.tree_logl_valid = false;
  return tree;
}

static void operandOf(const DisplayedTreeData &t, uint32_t edge, uint32_t &kind, uint32_t &idx, uint32_t &e) {
  kind = t.isTip ? NRX_TIP : NRX_CLV;
  idx = t.isTip ? t.tip : t.slot;
  e = edge;
}

/* add_tree_single (:115-152) / add_tree_both (:154-192): instead of calling libpll, queue one nrx_op */
static void add_tree(AnnotatedNetwork &ann, size_t clv_index, size_t left_node, size_t left_tree, uint32_t left_edge,
                     long right_node, size_t right_tree, uint32_t right_edge, const ReticulationConfigSet &rc) {
  if (tree_already_present_and_fine(ann, clv_index, rc)) return;
  DisplayedTreeData &tree = add_displayed_tree(ann, clv_index, rc);
  nrx_op op{};
  op.parent_slot = tree.slot;
  operandOf(ann.pernode_displayed_tree_data[left_node].displayed_trees[left_tree], left_edge, op.left_kind, op.left_idx, op.left_edge);
  if (right_node >= 0) operandOf(ann.pernode_displayed_tree_data[right_node].displayed_trees[right_tree], right_edge, op.right_kind, op.right_idx, op.right_edge);
  else { op.right_kind = NRX_NONE; op.right_idx = 0; op.right_edge = (uint32_t)ann.network.num_branches(); }  // fake clv / fake pmatrix
  ann.pending_ops.push_back(op);
  tree.clv_valid = true;
}

static void processNodeImprovedSingleChild(AnnotatedNetwork &ann, Node *node, Node *child, const ReticulationConfigSet &extra) {  // :194-223
  const uint32_t edge = (uint32_t)edgeBetween(ann.network, child->clv_index, node->clv_index);
  ReticulationConfigSet restrictionsSet = getRestrictionsToTakeNeighbor(ann, node->clv_index, child->clv_index);
  if (!extra.configs.empty()) restrictionsSet = combineReticulationChoices(restrictionsSet, extra);
  NodeDisplayedTreeData &dc = ann.pernode_displayed_tree_data[child->clv_index];
  for (size_t i = 0; i < dc.num_active_displayed_trees; ++i) {
    const ReticulationConfigSet &crc = dc.displayed_trees[i].treeLoglData.reticulationChoices;
    if (reticulationConfigsCompatible(crc, restrictionsSet))
      add_tree(ann, node->clv_index, child->clv_index, i, edge, -1, 0, 0, combineReticulationChoices(crc, restrictionsSet));
  }
}

static ReticulationConfigSet getReticulationChoicesThisOnly(AnnotatedNetwork &ann, const ReticulationConfigSet &this_tree_config,
                                                            const ReticulationConfigSet &other_child_dead_settings, size_t parent,
                                                            size_t this_child, size_t other_child) {  // ReticulationConfigHelper.cpp:98-144
  ReticulationConfigSet res(ann.options.max_reticulations);
  ReticulationConfigSet restricted = combineReticulationChoices(this_tree_config, getRestrictionsToTakeNeighbor(ann, parent, this_child));
  if (restricted.configs.empty()) return res;
  const ReticulationConfigSet c1 = combineReticulationChoices(restricted, getRestrictionsToDismissNeighbor(ann, parent, other_child));
  res.configs.insert(res.configs.end(), c1.configs.begin(), c1.configs.end());
  restricted = combineReticulationChoices(restricted, getRestrictionsToTakeNeighbor(ann, parent, other_child));
  const ReticulationConfigSet c2 = combineReticulationChoices(restricted, other_child_dead_settings);
  res.configs.insert(res.configs.end(), c2.configs.begin(), c2.configs.end());
  simplifyReticulationChoices(res);
  return res;
}

static ReticulationConfigSet deadNodeSettings(AnnotatedNetwork &ann, const NodeDisplayedTreeData &dt, size_t parent, size_t child) {  // :146-194
  ReticulationConfigSet res(ann.options.max_reticulations);
  const ReticulationConfigSet notTaken = getRestrictionsToDismissNeighbor(ann, parent, child);
  res.configs.insert(res.configs.end(), notTaken.configs.begin(), notTaken.configs.end());
  const ReticulationConfigSet taken = getRestrictionsToTakeNeighbor(ann, parent, child);
  const size_t max_n_trees = (size_t)1 << ann.network.num_reticulations();
  for (size_t t = 0; t < max_n_trees; ++t) {
    const ReticulationConfigSet rc = getTreeConfig(ann, t);
    if (!reticulationConfigsCompatible(rc, taken)) continue;
    bool found = false;
    for (size_t i = 0; i < dt.num_active_displayed_trees && !found; ++i)
      found = reticulationConfigsCompatible(rc, dt.displayed_trees[i].treeLoglData.reticulationChoices);
    if (!found) res.configs.push_back(rc.configs[0]);
  }
  simplifyReticulationChoices(res);
  return res;
}

static void processNodeImprovedTwoChildren(AnnotatedNetwork &ann, Node *node, Node *left, Node *right, const ReticulationConfigSet &extra) {  // :225-345
  const size_t v = node->clv_index, l = left->clv_index, r = right->clv_index;
  const uint32_t le = (uint32_t)edgeBetween(ann.network, l, v), re = (uint32_t)edgeBetween(ann.network, r, v);
  ReticulationConfigSet both = combineReticulationChoices(getRestrictionsToTakeNeighbor(ann, v, l), getRestrictionsToTakeNeighbor(ann, v, r));
  if (!extra.configs.empty()) both = combineReticulationChoices(both, extra);
  NodeDisplayedTreeData &dl = ann.pernode_displayed_tree_data[l];
  NodeDisplayedTreeData &dr = ann.pernode_displayed_tree_data[r];

  for (size_t i = 0; i < dl.num_active_displayed_trees; ++i) {  // both children active
    const ReticulationConfigSet &lc = dl.displayed_trees[i].treeLoglData.reticulationChoices;
    if (!reticulationConfigsCompatible(lc, both)) continue;
    for (size_t j = 0; j < dr.num_active_displayed_trees; ++j) {
      const ReticulationConfigSet &rcfg = dr.displayed_trees[j].treeLoglData.reticulationChoices;
      if (!reticulationConfigsCompatible(rcfg, both)) continue;
      if (reticulationConfigsCompatible(lc, rcfg))
        add_tree(ann, v, l, i, le, (long)r, j, re, combineReticulationChoices(combineReticulationChoices(lc, rcfg), both));
    }
  }
  ReticulationConfigSet right_dead = deadNodeSettings(ann, dr, v, r);  // only left child
  if (!extra.configs.empty()) right_dead = combineReticulationChoices(right_dead, extra);
  for (size_t i = 0; i < dl.num_active_displayed_trees; ++i) {
    ReticulationConfigSet lo = getReticulationChoicesThisOnly(ann, dl.displayed_trees[i].treeLoglData.reticulationChoices, right_dead, v, l, r);
    if (!extra.configs.empty()) lo = combineReticulationChoices(lo, extra);
    if (!lo.configs.empty()) add_tree(ann, v, l, i, le, -1, 0, 0, lo);
  }
  ReticulationConfigSet left_dead = deadNodeSettings(ann, dl, v, l);  // only right child
  if (!extra.configs.empty()) left_dead = combineReticulationChoices(left_dead, extra);
  for (size_t i = 0; i < dr.num_active_displayed_trees; ++i) {
    ReticulationConfigSet ro = getReticulationChoicesThisOnly(ann, dr.displayed_trees[i].treeLoglData.reticulationChoices, left_dead, v, r, l);
    if (!extra.configs.empty()) ro = combineReticulationChoices(ro, extra);
    if (!ro.configs.empty()) add_tree(ann, v, r, i, re, -1, 0, 0, ro);
  }
}

void processNodeImproved(AnnotatedNetwork &ann, int incremental, Node *node, std::vector<Node *> &children,
                         const ReticulationConfigSet &extra, bool append) {  // :347-408
  const size_t v = node->clv_index;
  if (v < ann.network.num_tips()) return;
  if (incremental && allClvsValid(ann, v)) return;
  // a launch may only contain mutually independent ops: if a child's trees are still queued, or this node
  // already has queued ops (append mode), issue the queue first
  bool dep = ann.pending_parent[v];
  for (Node *c : children) dep |= (bool)ann.pending_parent[c->clv_index];
  if (dep) flushPendingOps(ann);
  if (!append) ann.pernode_displayed_tree_data[v].num_active_displayed_trees = 0;
  if (children.empty()) { validateSingleClv(ann, v); return; }
  if (children.size() == 1) processNodeImprovedSingleChild(ann, node, children[0], extra);
  else if (children.size() == 2) processNodeImprovedTwoChildren(ann, node, children[0], children[1], extra);
  else throw std::runtime_error("Node has too many children");
  ann.pending_parent[v] = 1;
  if (!rerootSessionOpen(ann) && v < ann.node_version.size()) ann.node_version[v]++;   // root-directed data recomputed: memoised re-rooted trees built on it miss from now on
  if (ann.pernode_displayed_tree_data[v].num_active_displayed_trees > ((size_t)1 << ann.network.num_reticulations()))
    throw std::runtime_error("Too many displayed trees stored at node " + std::to_string(v));
  validateSingleClv(ann, v);
}

/* computeDisplayedTreeLoglikelihood (:410-486) for ALL root trees: one K3 launch, one reduction.  Split in two so that
 * several networks can be in flight at once (computeLoglikelihoodBatch): Begin selects the trees and enqueues the device
 * work without blocking, End collects the [trees][partitions] sums. */
static void treeLoglikelihoodsBegin(AnnotatedNetwork &ann, Node *actRoot, bool replayed, std::vector<uint32_t> *slots_out) {
  NodeDisplayedTreeData &rd = ann.pernode_displayed_tree_data[actRoot->clv_index];
  std::vector<uint32_t> slots;
  std::vector<size_t> which;
  PlanCache *pcr = (replayed && ann.plan && ann.plan->valid && actRoot == ann.network.root) ? ann.plan : nullptr;
  const bool use_cached = pcr && pcr->root_cached && pcr->root_prob_epoch == ann.prob_epoch;
  if (use_cached) { slots = pcr->root_slots; which = pcr->root_which; }
  for (size_t i = 0; !use_cached && i < rd.num_active_displayed_trees; ++i) {
    DisplayedTreeData &t = rd.displayed_trees[i];
    refreshLogprob(ann, t.treeLoglData);
    if (t.treeLoglData.tree_logprob < ann.options.min_interesting_tree_logprob) continue;
    // findFirstNodeWithTwoActiveChildren (ReticulationConfigHelper.cpp:251-272): skip the dead path below the root
    setReticulationParents(ann.network, t.treeLoglData.reticulationChoices.configs[0]);
    size_t dtr = actRoot->clv_index;
    collect_dead_nodes(ann.network, actRoot->clv_index, &dtr);
    NodeDisplayedTreeData &dd = ann.pernode_displayed_tree_data[dtr];  // findMatchingDisplayedTree (:213-249)
    DisplayedTreeData *match = nullptr;
    size_t n_good = 0;
    for (size_t k = 0; k < dd.num_active_displayed_trees; ++k)
      if (reticulationConfigsCompatible(t.treeLoglData.reticulationChoices, dd.displayed_trees[k].treeLoglData.reticulationChoices)) { n_good++; match = &dd.displayed_trees[k]; }
    if (n_good > 1) throw std::runtime_error("Found multiple suitable trees");
    if (n_good == 0) throw std::runtime_error("Found no suitable displayed tree");
    slots.push_back(match->slot);
    which.push_back(i);
  }
  if (pcr && !use_cached) { pcr->root_slots = slots; pcr->root_which = which; pcr->root_prob_epoch = ann.prob_epoch; pcr->root_cached = true; }
  flushPendingOps(ann);
  if (slots_out) *slots_out = slots;
  if (!slots.empty()) {
    // after a plan replay whose ops carried lnl marks for exactly these trees, K2 has already written the per-site lnLs
    if (replayed && ann.plan->fused && ann.plan->on_device && slots == ann.plan->lnl_slots && nrx_supports_fused_lnl(ann.engine)) {
      if (ann.plan->run_pending) {   // CLVs + root lnLs of the replayed plan in one engine call (one launch when the plan tile-walks)
        ann.plan->run_pending = false;
        engineCheck(nrx_plan_evaluate_async(ann.engine, ann.plan->engine_plan, slots.data(), (uint32_t)slots.size()), "nrx_plan_evaluate");
      } else
        engineCheck(nrx_tree_lnl_fused_async(ann.engine, ann.plan->engine_plan, slots.data(), (uint32_t)slots.size()), "nrx_tree_lnl_fused");
    } else {
      if (ann.plan && ann.plan->run_pending) { ann.plan->run_pending = false; engineCheck(nrx_plan_run(ann.engine, ann.plan->engine_plan), "nrx_plan_run"); }
      engineCheck(nrx_tree_lnl_async(ann.engine, slots.data(), (uint32_t)slots.size()), "nrx_tree_lnl");
    }
  }
  if (ann.plan && ann.plan->run_pending) { ann.plan->run_pending = false; engineCheck(nrx_plan_run(ann.engine, ann.plan->engine_plan), "nrx_plan_run"); }
  ann.pending_root = actRoot->clv_index;
  ann.pending_trees = which;
  ann.pending_eval = true;
}

static void treeLoglikelihoodsEnd(AnnotatedNetwork &ann) {
  if (!ann.pending_eval) return;
  ann.pending_eval = false;
  NodeDisplayedTreeData &rd = ann.pernode_displayed_tree_data[ann.pending_root];
  const unsigned P = ann.fake_treeinfo->partition_count;
  const std::vector<size_t> &which = ann.pending_trees;
  std::vector<double> out(which.size() * P, 0.0);
  if (!which.empty()) engineCheck(nrx_result_wait(ann.engine, out.data(), (uint32_t)out.size()), "nrx_result_wait");
  reduceSum(ann, out.data(), out.size());  // C2: one reduction for all trees (reference: one per tree)
  for (size_t k = 0; k < which.size(); ++k) {
    TreeLoglData &t = rd.displayed_trees[which[k]].treeLoglData;
    for (unsigned p = 0; p < P; ++p) {
      const double tl = out[k * P + p];
      if (tl == -std::numeric_limits<double>::infinity()) throw std::runtime_error("negative infinity tree loglikelihood in partition " + std::to_string(p));
      t.tree_partition_logl[p] = tl;
    }
    t.tree_logl_valid = true;
  }
}

static void snapshotPlan(AnnotatedNetwork &ann) {
  PlanCache &pc = *ann.plan;
  pc.nodes.assign(ann.network.num_nodes(), PlanCache::NodeSnap());
  for (size_t v = ann.network.num_tips(); v < ann.network.num_nodes(); ++v) {
    NodeDisplayedTreeData &nd = ann.pernode_displayed_tree_data[v];
    for (size_t i = 0; i < nd.num_active_displayed_trees; ++i) {
      pc.nodes[v].configs.push_back(nd.displayed_trees[i].treeLoglData.reticulationChoices);
      pc.nodes[v].slots.push_back(nd.displayed_trees[i].slot);
    }
  }
  pc.valid = true;
}

/* hand the recorded batches to the engine: ops stay resident on the device, replay = one CUDA graph launch; the ops
 * that produce the root displayed trees' CLVs are marked so that K2 also emits their per-site lnL (fused K3) */
static void createEnginePlan(AnnotatedNetwork &ann, const std::vector<uint32_t> &root_slots) {
  PlanCache &pc = *ann.plan;
  dropEnginePlan(ann);
  std::vector<nrx_op> flat;
  std::vector<uint32_t> sizes;
  for (const std::vector<nrx_op> &b : pc.batches) { flat.insert(flat.end(), b.begin(), b.end()); sizes.push_back((uint32_t)b.size()); }
  pc.lnl_slots = root_slots;
  pc.fused = !root_slots.empty() && nrx_supports_fused_lnl(ann.engine);
  if (pc.fused) {
    std::vector<uint32_t> sorted = root_slots;
    std::sort(sorted.begin(), sorted.end());
    if (std::adjacent_find(sorted.begin(), sorted.end()) != sorted.end()) pc.fused = false;  // two root trees share a CLV: keep K3 separate
  }
  if (pc.fused) {
    size_t marked = 0;
    for (nrx_op &op : flat)
      for (size_t k = 0; k < root_slots.size(); ++k)
        if (op.parent_slot == root_slots[k]) { op.lnl_item = (uint32_t)k + 1; ++marked; }
    bool aa = false;
    for (const PartitionModel &m : ann.fake_treeinfo->partitions) aa = aa || m.states != 4;
    bool tiptip_marked = false;   // 20 states: tip x tip updates are a table kernel without the per-site epilogue
    for (const nrx_op &op : flat) tiptip_marked = tiptip_marked || (aa && op.lnl_item && op.left_kind == NRX_TIP && op.right_kind == NRX_TIP);
    if (marked != root_slots.size() || tiptip_marked) {  // a root tree's CLV is not produced by this traversal (cannot happen for a full one)
      for (nrx_op &op : flat) op.lnl_item = 0;
      pc.fused = false;
    }
  }
  engineCheck(nrx_plan_create(ann.engine, flat.data(), sizes.data(), (uint32_t)sizes.size(), &pc.engine_plan), "nrx_plan_create");
  pc.on_device = true;
}

static void replayPlan(AnnotatedNetwork &ann) {
  PlanCache &pc = *ann.plan;
  for (size_t v = ann.network.num_tips(); v < ann.network.num_nodes(); ++v) {
    NodeDisplayedTreeData &nd = ann.pernode_displayed_tree_data[v];
    const PlanCache::NodeSnap &sn = pc.nodes[v];
    for (size_t i = 0; i < sn.configs.size(); ++i) {
      DisplayedTreeData &t = nd.displayed_trees[i];  // entries (and their slots) persist, see add_displayed_tree
      t.treeLoglData.reticulationChoices = sn.configs[i];
      t.treeLoglData.tree_logprob_valid = false;
      t.treeLoglData.tree_logl_valid = false;
      t.clv_valid = true;
    }
    nd.num_active_displayed_trees = sn.configs.size();
    validateSingleClv(ann, v);
    if (v < ann.node_version.size()) ann.node_version[v]++;
  }
  uint64_t local_sites = 0;
  for (const PartitionModel &m : ann.fake_treeinfo->partitions) local_sites += m.sites;
  if (pc.on_device) {
    if (pc.fused && ann.score_only_now) ann.root_clvs_stale = true;   // this replay leaves the root displayed trees' CLVs unwritten
    if (pc.fused) pc.run_pending = true;   // treeLoglikelihoodsBegin, called next, issues it together with the root lnLs
    else engineCheck(nrx_plan_run(ann.engine, pc.engine_plan), "nrx_plan_run");
  }
  for (const std::vector<nrx_op> &b : pc.batches) {
    if (!pc.on_device) engineCheck(nrx_update_clvs(ann.engine, b.data(), (uint32_t)b.size()), "nrx_update_clvs");
    ann.clv_site_updates += local_sites * b.size();
  }
}

static void processPartitionsImproved(AnnotatedNetwork &ann, int incremental) {  // :488-519
  PlanCache &pc = *ann.plan;
  if (!incremental && ann.use_plan_cache && pc.valid) {
    replayPlan(ann);
  } else {
    const bool record = !incremental && ann.use_plan_cache;
    if (record) { pc.batches.clear(); pc.site_updates = 0; pc.recording = true; pc.valid = false; pc.root_cached = false; dropEnginePlan(ann); }
    for (Node *n : ann.travbuffer) {
      std::vector<Node *> children;
      for (size_t c : n->children) children.push_back(&ann.network.nodes[c]);
      processNodeImproved(ann, incremental, n, children, ReticulationConfigSet());
    }
    flushPendingOps(ann);
    if (record) {
      pc.recording = false;
      snapshotPlan(ann);
      std::vector<uint32_t> root_slots;
      treeLoglikelihoodsBegin(ann, ann.network.root, false, &root_slots);
      treeLoglikelihoodsEnd(ann);   // creating the engine plan synchronises anyway (first evaluation of a topology)
      createEnginePlan(ann, root_slots);
      return;
    }
    treeLoglikelihoodsBegin(ann, ann.network.root, false, nullptr);
    return;
  }
  treeLoglikelihoodsBegin(ann, ann.network.root, true, nullptr);
}

/* the mixing itself, over any sequence of trees (evaluateTrees used to deep-copy every root tree's TreeLoglData first: 192 x two
 * vectors per evaluation of BASELINE config 5, 20 us of host time behind which the GPU sat idle) */
template <class Get>
static double mixTreesPartition(AnnotatedNetwork &ann, size_t p, size_t n, Get get) {
  if (n == 0) throw std::runtime_error("evaluateTreesPartition: no trees");
  double pl;
  if (ann.options.likelihood_variant == LikelihoodVariant::AVERAGE_DISPLAYED_TREES) {
    // log(sum_t exp(tree_logprob_t + tree_partition_logl_t)): summed as the reference does, in log space (logSumExp's max-shift)
    double m = -std::numeric_limits<double>::infinity();
    for (size_t i = 0; i < n; ++i) {
      TreeLoglData &t = get(i);
      if (t.tree_logprob < ann.options.min_interesting_tree_logprob) continue;
      if (!t.tree_logl_valid) throw std::runtime_error("invalid tree logl");
      m = std::max(m, t.tree_logprob + t.tree_partition_logl[p]);
    }
    if (m == -std::numeric_limits<double>::infinity()) pl = m;
    else {
      double sum = 0.0;
      for (size_t i = 0; i < n; ++i) {
        TreeLoglData &t = get(i);
        if (t.tree_logprob < ann.options.min_interesting_tree_logprob) continue;
        sum += std::exp(t.tree_logprob + t.tree_partition_logl[p] - m);
      }
      pl = m + std::log(sum);
    }
  } else {
    pl = -std::numeric_limits<double>::infinity();
    for (size_t i = 0; i < n; ++i) {
      TreeLoglData &t = get(i);
      if (t.tree_logprob < ann.options.min_interesting_tree_logprob) continue;
      if (!t.tree_logl_valid) throw std::runtime_error("invalid tree logl");
      pl = std::max(pl, t.tree_logprob + t.tree_partition_logl[p]);
    }
  }
  ann.fake_treeinfo->partition_loglh[p] = pl;
  return pl;
}

double evaluateTreesPartition(AnnotatedNetwork &ann, size_t p, std::vector<TreeLoglData> &trees) {  // :521-604
  return mixTreesPartition(ann, p, trees.size(), [&](size_t i) -> TreeLoglData & { return trees[i]; });
}

static double evaluateTrees(AnnotatedNetwork &ann, Node *virtual_root) {  // :606-644
  NodeDisplayedTreeData &nd = ann.pernode_displayed_tree_data[virtual_root->clv_index];
  for (size_t i = 0; i < nd.num_active_displayed_trees; ++i) refreshLogprob(ann, nd.displayed_trees[i].treeLoglData);
  double network_logl = 0.0;
  for (unsigned p = 0; p < ann.fake_treeinfo->partition_count; ++p)
    network_logl += mixTreesPartition(ann, p, nd.num_active_displayed_trees, [&](size_t i) -> TreeLoglData & { return nd.displayed_trees[i].treeLoglData; });
  if (network_logl == -std::numeric_limits<double>::infinity()) throw std::runtime_error("Invalid network likelihood: negative infinity \n");
  ann.cached_logl = network_logl;
  ann.cached_logl_valid = true;
  return network_logl;
}

static bool reuseOldDisplayedTreesCheck(AnnotatedNetwork &ann, int incremental, size_t vroot) {  // AnnotatedNetwork.cpp:359-394
  if (!incremental) return false;
  if (!clvValidCheck(ann, vroot)) return false;
  NodeDisplayedTreeData &nd = ann.pernode_displayed_tree_data[vroot];
  for (size_t i = 0; i < nd.num_active_displayed_trees; ++i) {
    DisplayedTreeData &dtd = nd.displayed_trees[i];
    refreshLogprob(ann, dtd.treeLoglData);
    if ((!dtd.clv_valid || !dtd.treeLoglData.tree_logl_valid) && dtd.treeLoglData.tree_logprob >= ann.options.min_interesting_tree_logprob) return false;
  }
  return true;
}

/* computeLoglikelihoodImproved (:646-671) in two halves: Begin enqueues P-matrices, CLV updates, per-tree root lnLs,
 * the cross-rank reduction and the result copy on the network's own stream and returns; End waits for that stream and
 * mixes the trees.  computeLoglikelihoodImproved = Begin + End. */
static void computeLoglikelihoodImprovedBegin(AnnotatedNetwork &ann, int incremental, int update_pmatrices) {
  if (ann.pending_eval) throw std::runtime_error("computeLoglikelihoodBegin: an evaluation of this network is already in flight");
  finishVirtualReroot(ann);   // a re-rooting session the caller left open: the root-directed trees come back first
  if (ann.root_clvs_stale && (incremental || !ann.score_only)) {   // a score-only evaluation left the root trees' CLVs unwritten
    incremental = 0;
    ann.root_clvs_stale = false;
    ann.score_only_now = false;
  } else {
    ann.score_only_now = !incremental && ann.score_only;
  }
  if (!incremental) engineCheck(nrx_set_score_only(ann.engine, ann.score_only_now ? 1 : 0), "nrx_set_score_only");
  ann.begin_returns_cached = false;
  if (!incremental) invalidateAllCLVs(ann);
  const bool reuse = reuseOldDisplayedTreesCheck(ann, incremental, ann.network.root->clv_index);
  if (reuse) {
    if (ann.cached_logl_valid) ann.begin_returns_cached = true;
    return;
  }
  if (update_pmatrices) pllmod_treeinfo_update_prob_matrices(ann, !incremental);
  ann.begin_ran_traversal = true;
  processPartitionsImproved(ann, incremental);
}

static double computeLoglikelihoodImprovedEnd(AnnotatedNetwork &ann) {
  if (ann.begin_returns_cached) { ann.begin_returns_cached = false; return ann.cached_logl; }
  treeLoglikelihoodsEnd(ann);
  if (ann.begin_ran_traversal) {
    ann.begin_ran_traversal = false;
    if (!clvValidCheck(ann, ann.network.root->clv_index)) throw std::runtime_error("Invalid displayed trees after loglikelihood computation");
  }
  return evaluateTrees(ann, ann.network.root);
}

double computeLoglikelihoodImproved(AnnotatedNetwork &ann, int incremental, int update_pmatrices) {  // :646-671
  computeLoglikelihoodImprovedBegin(ann, incremental, update_pmatrices);
  return computeLoglikelihoodImprovedEnd(ann);
}

/* Batched scoring: the reference scores candidate networks one after the other
 * (src/search/Filtering.cpp:210-260: performMove -> optimise -> scoreNetwork -> undoMove); here every candidate is its
 * own AnnotatedNetwork = its own engine and CUDA stream, all evaluations are enqueued first and collected afterwards,
 * so the small kernels of different candidates overlap on the GPU and the host never idles on one stream at a time.
 * Result i is exactly computeLoglikelihood(*anns[i], incremental, update_pmatrices). */
std::vector<double> computeLoglikelihoodBatch(const std::vector<AnnotatedNetwork *> &anns, int incremental, int update_pmatrices) {
  for (AnnotatedNetwork *a : anns)
    if (a->options.likelihood_variant == LikelihoodVariant::SARAH_PSEUDO) throw std::runtime_error("computeLoglikelihoodBatch: the pseudo-likelihood variant is evaluated one network at a time");
  std::vector<double> out(anns.size(), 0.0);
  struct ModeGuard {  // throughput launch geometry while several networks are in flight
    const std::vector<AnnotatedNetwork *> &a;
    explicit ModeGuard(const std::vector<AnnotatedNetwork *> &a_) : a(a_) { if (a.size() > 1) for (AnnotatedNetwork *n : a) nrx_set_throughput_mode(n->engine, 1); }
    ~ModeGuard() { for (AnnotatedNetwork *n : a) nrx_set_throughput_mode(n->engine, 0); }
  } guard(anns);
  size_t begun = 0;
  try {
    for (; begun < anns.size(); ++begun) computeLoglikelihoodImprovedBegin(*anns[begun], incremental, update_pmatrices);
  } catch (...) {
    for (size_t i = 0; i < begun; ++i) { try { computeLoglikelihoodImprovedEnd(*anns[i]); } catch (...) {} }
    throw;
  }
  std::exception_ptr first;
  for (size_t i = 0; i < anns.size(); ++i) {
    try { out[i] = computeLoglikelihoodImprovedEnd(*anns[i]); }
    catch (...) { if (!first) first = std::current_exception(); anns[i]->pending_eval = false; }
  }
  if (first) std::rethrow_exception(first);
  return out;
}

/* ---- src/likelihood/PseudoLoglikelihood.cpp:57-226 ------------------------------------------------------------------
 * One CLV per node.  The reference runs up to three libpll updates per node and partition into scratch CLVs and merges
 * them on the host; here a node is ONE nrx_pseudo_op (the kernel does the three updates and the merge in registers), and
 * nodes whose children are final share a launch, like the displayed-tree ops of the exact evaluation. */
double computePseudoLoglikelihood(AnnotatedNetwork &ann, int incremental, int update_pmatrices) {
  Network &nw = ann.network;
  const unsigned P = ann.fake_treeinfo->partition_count;
  flushPendingOps(ann);
  finishVirtualReroot(ann);
  if (ann.pseudo_clv_valid.size() != nw.num_nodes()) {  // src/graph/AnnotatedNetwork.cpp:160-184
    ann.pseudo_clv_valid.assign(nw.num_nodes(), 0);
    for (size_t i = 0; i < nw.num_tips(); ++i) ann.pseudo_clv_valid[i] = 1;
    ann.pseudo_slot.assign(nw.num_nodes(), UINT32_MAX);
  }
  if (update_pmatrices) pllmod_treeinfo_update_prob_matrices(ann, !incremental);
  auto parentProb = [&](long child, size_t parent) {  // :95-118
    if (child < 0 || nw.nodes[child].type != NodeType::RETICULATION_NODE) return 1.0;
    const size_t r = nw.nodes[child].reticulation_index;
    return parent == nw.reticulations[r].first_parent ? ann.reticulation_probs[r] : 1.0 - ann.reticulation_probs[r];
  };
  auto operand = [&](long child, size_t parent, uint32_t &kind, uint32_t &idx, uint32_t &edge) {
    if (child < 0) { kind = NRX_NONE; idx = 0; edge = (uint32_t)nw.num_branches(); return; }   // fake clv behind the fake pmatrix
    edge = (uint32_t)edgeBetween(nw, (size_t)child, parent);
    if ((size_t)child < nw.num_tips()) { kind = NRX_TIP; idx = (uint32_t)child; }
    else { kind = NRX_CLV; idx = ann.pseudo_slot[child]; }
  };
  std::vector<nrx_pseudo_op> batch;
  std::vector<char> in_batch(nw.num_nodes(), 0);
  uint64_t local_sites = 0;
  for (const PartitionModel &m : ann.fake_treeinfo->partitions) local_sites += m.sites;
  auto flush = [&] {
    if (batch.empty()) return;
    engineCheck(nrx_update_pseudo_clvs(ann.engine, batch.data(), (uint32_t)batch.size()), "nrx_update_pseudo_clvs");
    ann.clv_site_updates += local_sites * batch.size();
    batch.clear();
    std::fill(in_batch.begin(), in_batch.end(), 0);
  };
  for (Node *node : ann.travbuffer) {
    const size_t v = node->clv_index;
    if (v < nw.num_tips()) continue;
    if (incremental && ann.pseudo_clv_valid[v]) continue;
    const std::vector<size_t> &children = node->children;   // getChildren: all children, whatever the active parents are
    if (children.empty() || children.size() > 2) throw std::runtime_error("computePseudoLoglikelihood: node with 0 or more than 2 children");
    const long left = (long)children[0], right = children.size() == 1 ? -1 : (long)children[1];
    if (in_batch[left] || (right >= 0 && in_batch[right])) flush();   // a launch holds mutually independent nodes only
    if (ann.pseudo_slot[v] == UINT32_MAX) ann.pseudo_slot[v] = allocSlot(ann);
    const double p_left = parentProb(left, v), p_right = parentProb(right, v);
    nrx_pseudo_op op{};
    op.parent_slot = ann.pseudo_slot[v];
    operand(left, v, op.left_kind, op.left_idx, op.left_edge);
    operand(right, v, op.right_kind, op.right_idx, op.right_edge);
    op.w[0] = p_left * p_right;
    op.w[1] = p_left * (1.0 - p_right);
    op.w[2] = (1.0 - p_left) * p_right;
    op.w[3] = (1.0 - p_left) * (1.0 - p_right);
    batch.push_back(op);
    in_batch[v] = 1;
    ann.pseudo_clv_valid[v] = 1;
  }
  flush();
  std::vector<double> partition_pseudo_logl(P, 0.0);
  const uint32_t root_slot = ann.pseudo_slot[nw.root->clv_index];
  engineCheck(nrx_tree_lnl(ann.engine, &root_slot, 1, partition_pseudo_logl.data(), nullptr, 0), "nrx_tree_lnl");
  reduceSum(ann, partition_pseudo_logl.data(), P);
  double pseudo_logl = 0.0;
  for (unsigned p = 0; p < P; ++p) { pseudo_logl += partition_pseudo_logl[p]; ann.fake_treeinfo->partition_loglh[p] = partition_pseudo_logl[p]; }
  return pseudo_logl;
}

double computeLoglikelihood(AnnotatedNetwork &ann, int incremental, int update_pmatrices) {  // LikelihoodComputation.cpp:18-33
  if (ann.options.likelihood_variant == LikelihoodVariant::SARAH_PSEUDO) return computePseudoLoglikelihood(ann, incremental, update_pmatrices);
  return computeLoglikelihoodImproved(ann, incremental, update_pmatrices);
}

}  // namespace netrax
