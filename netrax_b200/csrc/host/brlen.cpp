/*
 * brlen.cpp — edge-rooted likelihood, virtual re-rooting, sumtables and branch-length derivatives on the device
 * engine.  Semantics follow src/likelihood/VirtualRerooting.cpp and src/likelihood/LikelihoodDerivatives.cpp of
 * the reference (cited per function), including its quirks Q1, Q2 and Q6 — parity means
 * reproducing the reference as it behaves.  Device work is batched: all (source-tree, target-tree) pairs of an
 * edge go through ONE nrx_edge_lnl / nrx_sumtables / nrx_derivatives call and one reduction.
 */
#include <algorithm>
#include <cmath>
#include <queue>
#include <unordered_set>

#include "host_internal.hpp"

namespace netrax {
using namespace detail;

std::vector<DisplayedTreeData> extractOldTrees(AnnotatedNetwork &ann, Node *virtual_root) {  // BranchLengthOptimization.cpp:34-53
  if (!clvValidCheck(ann, virtual_root->clv_index))
    throw std::runtime_error("Cannot reuse old displayed trees before the extractOldTrees step. For some reason, they are invalidated at the root node " + std::to_string(virtual_root->clv_index));
  // The reference deep-copies every root CLV here although only treeLoglData is ever read from the copies
  // (VirtualRerooting.cpp:254-277,471-544); the metadata copy below is sufficient and moves no CLV bytes.
  std::vector<DisplayedTreeData> old;
  NodeDisplayedTreeData &nd = ann.pernode_displayed_tree_data[virtual_root->clv_index];
  for (size_t i = 0; i < nd.num_active_displayed_trees; ++i) {
    DisplayedTreeData d = nd.displayed_trees[i];
    d.slot = UINT32_MAX;
    old.push_back(d);
  }
  return old;
}

ReticulationConfigSet getRestrictionsActiveAliveBranch(AnnotatedNetwork &ann, size_t pmatrix_index) {  // ReticulationConfigHelper.cpp:319-331
  ReticulationConfigSet res;  // max_reticulations stays 0 as in the reference: simplify only removes duplicates
  for (size_t t = 0; t < ((size_t)1 << ann.network.num_reticulations()); ++t) {
    const ReticulationConfigSet tc = getTreeConfig(ann, t);
    if (isActiveAliveBranch(ann, tc, pmatrix_index)) res.configs.push_back(tc.configs[0]);
  }
  simplifyReticulationChoices(res);
  return res;
}

namespace {
struct PathToVirtualRoot {  // VirtualRerooting.cpp:12-19
  ReticulationConfigSet reticulationChoices;
  std::vector<size_t> path;
  std::vector<std::vector<size_t>> children;
};

std::vector<size_t> getParentPointers(AnnotatedNetwork &ann, size_t vroot) {  // ParentHelper.cpp:62-86
  std::vector<size_t> parent(ann.network.num_nodes(), SIZE_MAX);
  parent[vroot] = vroot;
  std::queue<size_t> q;
  q.push(vroot);
  while (!q.empty()) {
    const size_t a = q.front(); q.pop();
    for (size_t nb : activeNeighbors(ann.network, a))
      if (parent[nb] == SIZE_MAX) { q.push(nb); parent[nb] = a; }
  }
  parent[vroot] = SIZE_MAX;
  return parent;
}

std::vector<size_t> getCurrentChildren(AnnotatedNetwork &ann, size_t node, size_t parent, const ReticulationConfigSet &restrictions) {  // ChildrenHelper.cpp:122-152
  std::vector<size_t> res;
  for (size_t c : ann.network.nodes[node].neighbors) {
    if (c == parent) continue;
    if (reticulationConfigsCompatible(restrictions, getRestrictionsToTakeNeighbor(ann, node, c))) res.push_back(c);
  }
  if (res.size() > 2) throw std::runtime_error("getCurrentChildren: more than two children at node " + std::to_string(node));
  return res;
}

std::vector<PathToVirtualRoot> getPathsToVirtualRoot(AnnotatedNetwork &ann, size_t old_vr, size_t new_vr, size_t new_vr_back) {  // :51-129
  std::vector<PathToVirtualRoot> res;
  NodeDisplayedTreeData &old = ann.pernode_displayed_tree_data[old_vr];
  for (size_t i = 0; i < old.num_active_displayed_trees; ++i) {
    setReticulationParents(ann.network, old.displayed_trees[i].treeLoglData.reticulationChoices.configs[0]);
    const std::vector<size_t> parent = getParentPointers(ann, new_vr);
    PathToVirtualRoot ptvr;
    for (size_t a = old_vr; a != new_vr; a = parent[a]) {
      if (a == SIZE_MAX) throw std::runtime_error("new virtual root is not reachable from the old one");
      ptvr.path.push_back(a);
    }
    ptvr.path.push_back(new_vr);
    ReticulationConfigSet rs(ann.options.max_reticulations);
    rs.configs.push_back(ReticulationConfig{});
    for (size_t j = 0; j + 1 < ptvr.path.size(); ++j) rs = combineReticulationChoices(rs, getRestrictionsToTakeNeighbor(ann, ptvr.path[j], ptvr.path[j + 1]));
    ptvr.reticulationChoices = rs;
    for (size_t j = 0; j + 1 < ptvr.path.size(); ++j)
      ptvr.children.push_back(getCurrentChildren(ann, ptvr.path[j], ptvr.path[j] == new_vr_back ? new_vr : parent[ptvr.path[j]], rs));
    ptvr.children.push_back(getCurrentChildren(ann, new_vr, new_vr_back, rs));
    res.push_back(ptvr);
  }
  for (bool dup = true; dup;) {  // kick out duplicate paths, same removal order as the reference
    dup = false;
    for (size_t i = 0; i + 1 < res.size() && !dup; ++i)
      for (size_t j = i + 1; j < res.size(); ++j)
        if (res[i].path == res[j].path) { dup = true; std::swap(res[j], res.back()); res.pop_back(); break; }
  }
  return res;
}

struct NodeSaveInformation {  // :131-190
  std::vector<std::unordered_set<size_t>> pathNodesToRestore;
  std::unordered_set<size_t> nodesInDanger;
};

NodeSaveInformation computeNodeSaveInformation(const std::vector<PathToVirtualRoot> &paths) {
  NodeSaveInformation info;
  info.pathNodesToRestore.resize(paths.size());
  for (size_t p = 1; p < paths.size(); ++p) {
    std::unordered_set<size_t> &restore = info.pathNodesToRestore[p];
    for (size_t i = 0; i < paths[p].path.size(); ++i)
      for (size_t c : paths[p].children[i]) restore.insert(c);
    for (size_t n : paths[p].path) restore.erase(n);
    for (auto it = restore.begin(); it != restore.end();) {  // keep only nodes an EARLIER path overwrites
      bool overwritten = false;
      for (size_t q = 0; q < p && !overwritten; ++q)
        overwritten = std::find(paths[q].path.begin(), paths[q].path.end(), *it) != paths[q].path.end();
      it = overwritten ? std::next(it) : restore.erase(it);
    }
  }
  for (const auto &s : info.pathNodesToRestore) info.nodesInDanger.insert(s.begin(), s.end());
  return info;
}

/* deep copy of a node's displayed trees into temporary device slots / back (the reference copy-assigns
 * NodeDisplayedTreeData, i.e. memcpy's CLVs on the host: VirtualRerooting.cpp:211-220,234-238) */
struct SavedNode { std::vector<DisplayedTreeData> trees; size_t num_active = 0; };

struct SlotCopies {  // (dst, src) pairs of one save / restore step, issued as ONE device launch
  std::vector<uint32_t> dst, src;
  void add(uint32_t d, uint32_t s) { dst.push_back(d); src.push_back(s); }
  void run(AnnotatedNetwork &ann) {
    if (!dst.empty()) engineCheck(nrx_copy_slots(ann.engine, dst.data(), src.data(), (uint32_t)dst.size()), "nrx_copy_slots");
    dst.clear(); src.clear();
  }
};

SavedNode saveNode(AnnotatedNetwork &ann, size_t v, SlotCopies &copies) {
  SavedNode s;
  NodeDisplayedTreeData &nd = ann.pernode_displayed_tree_data[v];
  s.num_active = nd.num_active_displayed_trees;
  for (size_t i = 0; i < nd.num_active_displayed_trees; ++i) {
    DisplayedTreeData d = nd.displayed_trees[i];
    if (!d.isTip) {
      d.slot = allocSlot(ann);
      copies.add(d.slot, nd.displayed_trees[i].slot);
    }
    s.trees.push_back(d);
  }
  return s;
}

void restoreNode(AnnotatedNetwork &ann, size_t v, const SavedNode &s, SlotCopies &copies) {
  NodeDisplayedTreeData &nd = ann.pernode_displayed_tree_data[v];
  if (v < ann.network.num_tips()) return;  // tips are immutable
  for (size_t i = 0; i < s.trees.size(); ++i) {
    if (i >= nd.displayed_trees.size()) {
      DisplayedTreeData d;
      d.slot = allocSlot(ann);
      nd.displayed_trees.push_back(d);
    }
    const uint32_t own = nd.displayed_trees[i].slot;
    nd.displayed_trees[i] = s.trees[i];
    nd.displayed_trees[i].slot = own;  // entries keep their own slot; only the contents come back
    copies.add(own, s.trees[i].slot);
  }
  nd.num_active_displayed_trees = s.num_active;
}
}  // namespace

void updateCLVsVirtualRerootTrees(AnnotatedNetwork &ann, Node *old_virtual_root, Node *new_virtual_root,
                                  Node *new_virtual_root_back, ReticulationConfigSet &restrictions) {  // :192-252
  const size_t old_vr = old_virtual_root->clv_index, new_vr = new_virtual_root->clv_index, back = new_virtual_root_back->clv_index;
  if (ann.pernode_displayed_tree_data[old_vr].num_active_displayed_trees == 0) throw std::runtime_error("no displayed trees at the old virtual root");
  flushPendingOps(ann);
  const std::vector<PathToVirtualRoot> paths = getPathsToVirtualRoot(ann, old_vr, new_vr, back);
  const NodeSaveInformation info = computeNodeSaveInformation(paths);
  std::vector<SavedNode> buffered(ann.network.num_nodes());
  SlotCopies copies;
  for (size_t n : info.nodesInDanger) buffered[n] = saveNode(ann, n, copies);
  copies.run(ann);
  for (size_t p = 0; p < paths.size(); ++p) {
    if (!reticulationConfigsCompatible(paths[p].reticulationChoices, restrictions)) continue;
    if (!info.pathNodesToRestore[p].empty()) flushPendingOps(ann);
    for (size_t n : info.pathNodesToRestore[p]) restoreNode(ann, n, buffered[n], copies);
    copies.run(ann);
    for (size_t i = 0; i < paths[p].path.size(); ++i) {
      const bool appendMode = (p > 0) && (paths[p].path[i] == new_vr);
      std::vector<Node *> children;
      for (size_t c : paths[p].children[i]) children.push_back(&ann.network.nodes[c]);
      processNodeImproved(ann, 0, &ann.network.nodes[paths[p].path[i]], children, paths[p].reticulationChoices, appendMode);
    }
  }
  flushPendingOps(ann);
  for (size_t n : info.nodesInDanger)
    for (const DisplayedTreeData &d : buffered[n].trees)
      if (!d.isTip) releaseSlot(ann, d.slot);
  if (ann.pernode_displayed_tree_data[new_vr].num_active_displayed_trees == 0) throw std::runtime_error("no displayed trees at the new virtual root");
}

namespace {
const TreeLoglData &getMatchingTreeData(const std::vector<DisplayedTreeData> &trees, const ReticulationConfigSet &query) {  // ReticulationConfigHelper.cpp:290-302
  for (const DisplayedTreeData &t : trees)
    if (reticulationConfigsCompatible(query, t.treeLoglData.reticulationChoices)) return t.treeLoglData;
  throw std::runtime_error("No compatible old tree data found");
}

void updateTreeData(AnnotatedNetwork &ann, const std::vector<DisplayedTreeData> &oldTrees, TreeLoglData &td) {  // VirtualRerooting.cpp:254-277
  const TreeLoglData &old = getMatchingTreeData(oldTrees, td.reticulationChoices);
  td.tree_partition_logl = old.tree_partition_logl;
  td.tree_logprob = computeReticulationConfigLogProb(td.reticulationChoices, ann.first_parent_logprobs, ann.second_parent_logprobs);
  td.tree_logprob_valid = true;
  td.tree_logl_valid = old.tree_logl_valid;
}
}  // namespace

double computeLoglikelihoodBrlenOpt(AnnotatedNetwork &ann, const std::vector<DisplayedTreeData> &oldTrees, unsigned int pmatrix_index,
                                    int update_pmatrices, bool) {  // :348-585
  if (ann.cached_logl_valid) return ann.cached_logl;
  const size_t source = ann.network.edges[pmatrix_index].source, target = ann.network.edges[pmatrix_index].target;
  NodeDisplayedTreeData &sd = ann.pernode_displayed_tree_data[source];
  NodeDisplayedTreeData &td = ann.pernode_displayed_tree_data[target];
  const size_t ns = sd.num_active_displayed_trees, nt = td.num_active_displayed_trees;
  const unsigned P = ann.fake_treeinfo->partition_count;
  if (!clvValidCheck(ann, ann.network.root->clv_index, false))
    throw std::runtime_error("Cannot reuse old displayed trees. For some reason, they are invalidated at the root node " + std::to_string(ann.network.root->clv_index));
  if (update_pmatrices) pllmod_treeinfo_update_prob_matrices(ann, 0);
  std::vector<TreeLoglData> combined;
  std::vector<char> sseen(ns, 0), tseen(nt, 0);
  std::vector<nrx_pair> pairs;       // recomputeTreeData (:279-346), batched
  std::vector<size_t> pair_owner;
  for (size_t i = 0; i < ns; ++i)
    for (size_t j = 0; j < nt; ++j) {
      const ReticulationConfigSet &a = sd.displayed_trees[i].treeLoglData.reticulationChoices, &b = td.displayed_trees[j].treeLoglData.reticulationChoices;
      if (!reticulationConfigsCompatible(a, b)) continue;
      TreeLoglData c(P, ann.options.max_reticulations);
      c.reticulationChoices = combineReticulationChoices(a, b);
      if (!isActiveAliveBranch(ann, c.reticulationChoices, pmatrix_index)) continue;
      c.tree_logprob = computeReticulationConfigLogProb(c.reticulationChoices, ann.first_parent_logprobs, ann.second_parent_logprobs);
      c.tree_logprob_valid = true;
      if (c.tree_logprob >= ann.options.min_interesting_tree_logprob) {
        pairs.push_back(makePair(sd.displayed_trees[i], td.displayed_trees[j]));
        pair_owner.push_back(combined.size());
      }
      combined.push_back(c);
      sseen[i] = tseen[j] = 1;
    }
  flushPendingOps(ann);
  std::vector<double> out(pairs.size() * P, 0.0);
  if (!pairs.empty()) engineCheck(nrx_edge_lnl(ann.engine, pmatrix_index, pairs.data(), (uint32_t)pairs.size(), out.data()), "nrx_edge_lnl");
  reduceSum(ann, out.data(), out.size());  // C3
  for (size_t k = 0; k < pairs.size(); ++k) {
    TreeLoglData &c = combined[pair_owner[k]];
    for (unsigned p = 0; p < P; ++p) {
      if (out[k * P + p] == 0.0) throw std::runtime_error("bad partition logl");
      c.tree_partition_logl[p] = out[k * P + p];
    }
    c.tree_logl_valid = true;
  }
  for (size_t i = 0; i < ns; ++i)
    if (!sseen[i] && isActiveAliveBranch(ann, sd.displayed_trees[i].treeLoglData.reticulationChoices, pmatrix_index)) {
      updateTreeData(ann, oldTrees, sd.displayed_trees[i].treeLoglData);
      combined.push_back(sd.displayed_trees[i].treeLoglData);
    }
  for (size_t j = 0; j < nt; ++j)
    if (!tseen[j] && isActiveAliveBranch(ann, td.displayed_trees[j].treeLoglData.reticulationChoices, pmatrix_index)) {
      updateTreeData(ann, oldTrees, td.displayed_trees[j].treeLoglData);
      combined.push_back(td.displayed_trees[j].treeLoglData);
    }
  for (const DisplayedTreeData &o : oldTrees) {  // :471-502 trees fully present in the old trees only
    bool seen = false;
    for (const TreeLoglData &c : combined)
      if (reticulationConfigsCompatible(o.treeLoglData.reticulationChoices, c.reticulationChoices)) { seen = true; break; }
    if (!seen) {
      TreeLoglData c(P, ann.options.max_reticulations);
      c.reticulationChoices = o.treeLoglData.reticulationChoices;
      updateTreeData(ann, oldTrees, c);
      combined.push_back(c);
    }
  }
  for (size_t t = 0; t < ((size_t)1 << ann.network.num_reticulations()); ++t) {  // :513-544 trees partially present
    const ReticulationConfigSet tc = getTreeConfig(ann, t);
    bool seen = false;
    for (const TreeLoglData &c : combined) if (reticulationConfigsCompatible(tc, c.reticulationChoices)) { seen = true; break; }
    if (seen) continue;
    for (const DisplayedTreeData &o : oldTrees)
      if (reticulationConfigsCompatible(tc, o.treeLoglData.reticulationChoices)) {
        TreeLoglData c(P, ann.options.max_reticulations);
        c.reticulationChoices = tc;
        updateTreeData(ann, oldTrees, c);
        combined.push_back(c);
        break;
      }
  }
  double network_logl = 0;
  for (unsigned p = 0; p < P; ++p) network_logl += evaluateTreesPartition(ann, p, combined);
  ann.cached_logl = network_logl;
  ann.cached_logl_valid = true;
  return network_logl;
}

std::vector<std::vector<SumtableInfo>> computePartitionSumtables(AnnotatedNetwork &ann, unsigned int pmatrix_index) {  // LikelihoodDerivatives.cpp:291-344
  const unsigned P = ann.fake_treeinfo->partition_count;
  std::vector<std::vector<SumtableInfo>> res(P);
  const size_t source = ann.network.edges[pmatrix_index].source, target = ann.network.edges[pmatrix_index].target;
  NodeDisplayedTreeData &sd = ann.pernode_displayed_tree_data[source];
  NodeDisplayedTreeData &td = ann.pernode_displayed_tree_data[target];
  std::vector<nrx_pair> pairs;
  for (size_t i = 0; i < sd.num_active_displayed_trees; ++i)
    for (size_t j = 0; j < td.num_active_displayed_trees; ++j) {
      const ReticulationConfigSet &a = sd.displayed_trees[i].treeLoglData.reticulationChoices, &b = td.displayed_trees[j].treeLoglData.reticulationChoices;
      if (!reticulationConfigsCompatible(a, b)) continue;
      const ReticulationConfigSet restrictions = combineReticulationChoices(a, b);
      if (!isActiveBranch(ann, restrictions, pmatrix_index)) continue;
      if (computeReticulationConfigLogProb(restrictions, ann.first_parent_logprobs, ann.second_parent_logprobs) < ann.options.min_interesting_tree_logprob) continue;
      SumtableInfo si;
      si.tree_prob = computeReticulationConfigProb(restrictions, ann.first_parent_logprobs, ann.second_parent_logprobs);
      si.index = (uint32_t)pairs.size();
      si.left_tree_idx = i; si.right_tree_idx = j;
      pairs.push_back(makePair(sd.displayed_trees[i], td.displayed_trees[j]));
      for (unsigned p = 0; p < P; ++p) res[p].push_back(si);
    }
  flushPendingOps(ann);
  if (!pairs.empty()) engineCheck(nrx_sumtables(ann.engine, pairs.data(), (uint32_t)pairs.size()), "nrx_sumtables");
  return res;
}

LoglDerivatives computeLoglikelihoodDerivatives(AnnotatedNetwork &ann, const std::vector<std::vector<SumtableInfo>> &sumtables,
                                                unsigned int pmatrix_index) {  // :190-232 + computePartitionLhData :30-188
  FakeTreeinfo &ti = *ann.fake_treeinfo;
  const unsigned P = ti.partition_count;
  if (sumtables.size() != P) throw std::runtime_error("computeLoglikelihoodDerivatives: one sumtable list per partition expected");
  if (ann.options.brlen_linkage == PLLMOD_COMMON_BRLEN_SCALED)
    throw std::runtime_error("I believe this function currently does not work correctly with scaled branch lengths");
  LoglDerivatives out;
  out.logl_prime = out.logl_prime_prime = 0.0;
  out.partition_logl_prime.assign(P, 0.0);
  out.partition_logl_prime_prime.assign(P, 0.0);
  out.raw.assign(P, {});
  const size_t n = sumtables[0].size();
  std::vector<double> brlen(P);
  for (unsigned p = 0; p < P; ++p)
    brlen[p] = (ann.options.brlen_linkage == PLLMOD_COMMON_BRLEN_UNLINKED) ? ti.branch_lengths[p][pmatrix_index] : ti.linked_branch_lengths[pmatrix_index];
  std::vector<double> vals(n * P * 3, 0.0);
  if (n) engineCheck(nrx_derivatives(ann.engine, (uint32_t)n, brlen.data(), vals.data()), "nrx_derivatives");
  reduceSum(ann, vals.data(), vals.size());  // C4: one reduction for all displayed-tree pairs and partitions
  for (unsigned p = 0; p < P; ++p) {
    const bool single_tree_mode = (n == 1);
    double res_prime = 0.0, res_prime_prime = 0.0;
    // single-tree mode: the reference passes f = nullptr to libpll (LikelihoodDerivatives.cpp:104), so no f exists
    for (size_t i = 0; i < n; ++i) for (int k = 0; k < 3; ++k) out.raw[p].push_back((single_tree_mode && k == 0) ? 0.0 : vals[(i * P + p) * 3 + k]);
    if (single_tree_mode) {
      res_prime = vals[p * 3 + 1];
      res_prime_prime = vals[p * 3 + 2];
    } else if (n > 0 && ann.options.likelihood_variant == LikelihoodVariant::AVERAGE_DISPLAYED_TREES) {
      // lh_t = exp(f_t) p_t, lh_t' = lh_t f_t', lh_t'' = lh_t' f_t' + lh_t f_t''  (computeTreeDerivatives :13-23, on
      // libpll's NEGATED derivatives, Q6); quotient rule (:173-180).  Every term carries the common factor
      // exp(-M), M = max f_t, which cancels in both quotients — this keeps mpreal's range in plain doubles.
      double M = -std::numeric_limits<double>::infinity();
      for (size_t i = 0; i < n; ++i) M = std::max(M, vals[(i * P + p) * 3]);
      double S = 0, S1 = 0, S2 = 0;
      for (size_t i = 0; i < n; ++i) {
        const double f = vals[(i * P + p) * 3], d1 = vals[(i * P + p) * 3 + 1], d2 = vals[(i * P + p) * 3 + 2];
        const double lh = std::exp(f - M);
        const double lhp = lh * d1;
        const double lhpp = lhp * d1 + lh * d2;
        S += lh * sumtables[p][i].tree_prob;
        S1 += lhp * sumtables[p][i].tree_prob;
        S2 += lhpp * sumtables[p][i].tree_prob;
      }
      res_prime = S1 / S;
      res_prime_prime = (S2 * S - S1 * S1) / (S * S);
    } else if (n > 0) {  // BEST: the tree maximising tree_logl * tree_prob — a product, as the reference does (Q2)
      double best = -std::numeric_limits<double>::infinity();
      res_prime = res_prime_prime = best;
      for (size_t i = 0; i < n; ++i) {
        const double f = vals[(i * P + p) * 3];
        if (f * sumtables[p][i].tree_prob > best) { best = f * sumtables[p][i].tree_prob; res_prime = vals[(i * P + p) * 3 + 1]; res_prime_prime = vals[(i * P + p) * 3 + 2]; }
      }
    }
    out.partition_logl_prime[p] = res_prime;
    out.partition_logl_prime_prime[p] = res_prime_prime;
    out.logl_prime += res_prime;
    out.logl_prime_prime += res_prime_prime;
  }
  return out;
}

}  // namespace netrax
