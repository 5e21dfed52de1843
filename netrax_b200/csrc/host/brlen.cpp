/*
 * brlen.cpp — edge-rooted likelihood, virtual re-rooting, sumtables and branch-length derivatives on the device
 * engine.  Semantics follow src/likelihood/VirtualRerooting.cpp and src/likelihood/LikelihoodDerivatives.cpp of
 * the reference (cited per function), including its quirks Q1, Q2 and Q6 — parity means
 * reproducing the reference as it behaves.  Device work is batched: all (source-tree, target-tree) pairs of an
 * edge go through ONE nrx_edge_lnl / nrx_sumtables / nrx_derivatives call and one reduction.
 */
#include <algorithm>
#include <cmath>
#include <cstdlib>
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

static ReticulationConfigSet computeRestrictionsActiveAliveBranch(AnnotatedNetwork &ann, size_t pmatrix_index, bool *all_trees = nullptr) {  // ReticulationConfigHelper.cpp:319-331
  ReticulationConfigSet res;  // max_reticulations stays 0 as in the reference: simplify only removes duplicates
  for (size_t t = 0; t < ((size_t)1 << ann.network.num_reticulations()); ++t) {
    const ReticulationConfigSet tc = getTreeConfig(ann, t);
    if (isActiveAliveBranch(ann, tc, pmatrix_index)) res.configs.push_back(tc.configs[0]);
  }
  if (all_trees) *all_trees = res.configs.size() == ((size_t)1 << ann.network.num_reticulations());
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

}  // namespace

/* ---- shadow re-rooting + memoised re-rooted node data ---------------------------------------------------------------------------
 * The reference re-roots IN PLACE: processNodeImproved overwrites the displayed trees of every node on the root -> new-root paths,
 * nodes a later path still needs in their original form are deep-copied first and copied back (VirtualRerooting.cpp:131-190,
 * 211-220,234-238), and after the branch is optimised `invalidatePmatrixIndex` throws everything above the edge away so that the
 * next computeLoglikelihood recomputes it (BranchLengthOptimization.cpp:417-420).  On the device that is two HBM passes per path
 * node and edge that buy nothing, so here:
 *  (1) a re-rooting SESSION never overwrites a root-directed CLV: the node's NodeDisplayedTreeData is moved into a stash, the
 *      re-rooted trees go to fresh slots, "restoring" a node for a later path is a metadata copy, and finishVirtualReroot moves the
 *      stash back.  If the branch length is what it was, nothing is invalid afterwards;
 *  (2) what processNodeImproved produced for (node, ordered children, the children's data identity, the lengths of the edges to
 *      them, the path's restriction set) is MEMOISED: the same call in a later session — the neighbouring edge's path shares all
 *      but its last node — installs the cached trees instead of launching K2 again.  Identity of a child's data = the id of the
 *      cache entry installed there, or the node's version counter for root-directed data (bumped by every recomputation), so a
 *      changed branch length or CLV anywhere below simply misses; reticulation probabilities, models, tips and the topology
 *      bump `clv_epoch`, which empties the cache.  A sweep that visits the edges in pre-order therefore costs ~1 node update
 *      per edge instead of the whole path down and up again.  The enumeration (which trees, which configs, which op per tree) is
 *      the reference's, call for call: hits return exactly what the miss computed. */
struct RerootCache {
  struct Entry {
    uint64_t id = 0, last_use = 0, pinned_session = 0;
    size_t node = 0;
    std::vector<size_t> children;
    std::vector<uint64_t> child_stamp;
    std::vector<double> lengths;          // [child][partition]
    std::vector<ReticulationConfig> extra;
    size_t extra_max = 0;
    NodeDisplayedTreeData data;           // owns the slots of data.displayed_trees
  };
  std::vector<Entry> entries;
  /* the host algebra of one re-rooting (paths, their children and restriction sets, the restore sets) depends on the topology only:
   * kept per (old root, new root, node behind it) until topology_changed() */
  struct Plan { size_t old_vr, new_vr, back; std::vector<PathToVirtualRoot> paths; NodeSaveInformation info; };
  std::vector<Plan> plans;
  std::vector<std::pair<size_t, ReticulationConfigSet>> edge_restrictions;   // getRestrictionsActiveAliveBranch per branch
  std::vector<std::pair<size_t, bool>> edge_all_trees;                        // ... and whether the branch is active and alive in ALL displayed trees
  size_t base_slots = 0;                             // slots the network itself used when the first session opened (the memo's default budget)
  bool lazy_session = false;                         // the open session was prepared without evaluating the network from its root first
  uint64_t topology_epoch = 0;
  uint64_t next_id = 1, tick = 0, session = 0, epoch = 0;
  size_t cached_slots = 0;
  // the open session
  bool active = false;
  size_t edge = 0;
  std::vector<double> start_lengths;                 // [partition]
  std::vector<char> stashed;                         // [node]
  std::vector<NodeDisplayedTreeData> orig;           // [node] root-directed data while stashed
  std::vector<std::vector<char>> orig_valid;         // [node][partition]
  std::vector<size_t> alias_n;                       // [node] leading trees of the CURRENT data whose slots belong to the stash or to an entry
  std::vector<uint64_t> installed;                   // [node] id of the entry whose trees are installed (state SHOWS_ENTRY)
  enum State : char { UNTOUCHED = 0, OWN, SHOWS_ENTRY, SHOWS_ORIGINAL, MIXED };
  std::vector<char> state;                           // [node] what the node's current data is
  std::vector<size_t> touched;
};

namespace {
constexpr uint64_t STAMP_ENTRY = 1ull << 63;

RerootCache &rerootState(AnnotatedNetwork &ann) {
  if (!ann.reroot) ann.reroot = new RerootCache();
  RerootCache &rc = *ann.reroot;
  const size_t n = ann.network.num_nodes();
  if (rc.stashed.size() != n) {
    rc.stashed.assign(n, 0); rc.orig.assign(n, NodeDisplayedTreeData()); rc.orig_valid.assign(n, {}); rc.alias_n.assign(n, 0); rc.installed.assign(n, 0);
    rc.state.assign(n, RerootCache::UNTOUCHED);
  }
  if (ann.node_version.size() != n) ann.node_version.assign(n, 0);
  if (!rc.base_slots) rc.base_slots = ann.next_slot - ann.free_slots.size();
  return rc;
}

size_t rerootBudget(AnnotatedNetwork &ann) {
  if (ann.reroot_cache_max_slots != SIZE_MAX) return ann.reroot_cache_max_slots;
  if (const char *v = std::getenv("NRX_REROOT_CACHE_SLOTS")) return (size_t)std::max(0l, std::atol(v));
  // default: as many slots as the network itself uses, but no more than 24 GB of CLVs + scalers
  double slot_bytes = 0;
  for (const PartitionModel &m : ann.fake_treeinfo->partitions) slot_bytes += (double)m.sites * (m.rate_cats * m.states_padded * 8.0 + 4.0);
  const size_t by_mem = (size_t)(24e9 / std::max(1.0, slot_bytes));
  // (next_slot counts the memo's and the sessions' own slots too: the network's own ones are what was allocated before the first session)
  const size_t own = ann.reroot && ann.reroot->base_slots ? ann.reroot->base_slots : ann.next_slot;
  return std::min<size_t>(std::max<size_t>(own, 32), by_mem);
}

void releaseEntry(AnnotatedNetwork &ann, RerootCache &rc, RerootCache::Entry &e) {
  for (const DisplayedTreeData &d : e.data.displayed_trees)
    if (!d.isTip && d.slot != UINT32_MAX) { releaseSlot(ann, d.slot); rc.cached_slots--; }
  e.data = NodeDisplayedTreeData();
}

/* the node's current (re-rooted) data goes away: slots it owns itself are released, aliased ones stay with their owner */
void dropCurrent(AnnotatedNetwork &ann, RerootCache &rc, size_t v) {
  NodeDisplayedTreeData &nd = ann.pernode_displayed_tree_data[v];
  for (size_t i = rc.alias_n[v]; i < nd.displayed_trees.size(); ++i)
    if (!nd.displayed_trees[i].isTip && nd.displayed_trees[i].slot != UINT32_MAX) releaseSlot(ann, nd.displayed_trees[i].slot);
  nd = NodeDisplayedTreeData();
  rc.alias_n[v] = 0;
  rc.installed[v] = 0;
  rc.state[v] = RerootCache::OWN;
}

void stashNode(AnnotatedNetwork &ann, RerootCache &rc, size_t v) {
  if (rc.stashed[v]) return;
  rc.orig[v] = std::move(ann.pernode_displayed_tree_data[v]);
  ann.pernode_displayed_tree_data[v] = NodeDisplayedTreeData();
  rc.orig_valid[v].resize(ann.fake_treeinfo->partition_count);
  for (unsigned p = 0; p < ann.fake_treeinfo->partition_count; ++p) rc.orig_valid[v][p] = ann.fake_treeinfo->clv_valid[p][v];
  rc.stashed[v] = 1;
  rc.alias_n[v] = 0;
  rc.installed[v] = 0;
  rc.state[v] = RerootCache::OWN;
  rc.touched.push_back(v);
}

/* restoreNode of the reference (VirtualRerooting.cpp:234-238): the node shows its root-directed trees again — a metadata copy */
void showOriginal(AnnotatedNetwork &ann, RerootCache &rc, size_t v) {
  if (v < ann.network.num_tips() || !rc.stashed[v]) return;   // tips are immutable; an un-stashed node still holds its original
  dropCurrent(ann, rc, v);
  NodeDisplayedTreeData &nd = ann.pernode_displayed_tree_data[v];
  nd = rc.orig[v];
  nd.displayed_trees.resize(nd.num_active_displayed_trees);   // trees appended later must take slots of their own
  rc.alias_n[v] = nd.displayed_trees.size();
  rc.state[v] = RerootCache::SHOWS_ORIGINAL;
  for (unsigned p = 0; p < ann.fake_treeinfo->partition_count; ++p) ann.fake_treeinfo->clv_valid[p][v] = rc.orig_valid[v][p];
}

uint64_t childStamp(AnnotatedNetwork &ann, RerootCache &rc, size_t c) {
  if (c < ann.network.num_tips()) return 0;
  switch (rc.stashed[c] ? rc.state[c] : RerootCache::UNTOUCHED) {
    case RerootCache::UNTOUCHED: case RerootCache::SHOWS_ORIGINAL: return ann.node_version[c];
    case RerootCache::SHOWS_ENTRY: return STAMP_ENTRY | rc.installed[c];
    default: return STAMP_ENTRY | rc.next_id++;   // own or mixed data (append mode): never matches
  }
}

/* processNodeImproved(ann, 0, node, children, extra, false) of a re-rooting path, through the memo */
void processPathNode(AnnotatedNetwork &ann, RerootCache &rc, size_t v, std::vector<Node *> &children, const ReticulationConfigSet &extra) {
  if (v < ann.network.num_tips()) return;
  const unsigned P = ann.fake_treeinfo->partition_count;
  std::vector<size_t> ch;
  std::vector<uint64_t> stamps;
  std::vector<double> lengths;
  for (Node *c : children) {
    ch.push_back(c->clv_index);
    stamps.push_back(childStamp(ann, rc, c->clv_index));
    const size_t e = edgeBetween(ann.network, c->clv_index, v);
    for (unsigned p = 0; p < P; ++p) {
      double len = ann.fake_treeinfo->branch_lengths[p][e];
      if (ann.fake_treeinfo->brlen_linkage != PLLMOD_COMMON_BRLEN_UNLINKED) len = ann.fake_treeinfo->linked_branch_lengths[e];
      lengths.push_back(len);
    }
  }
  stashNode(ann, rc, v);
  const size_t budget = rerootBudget(ann);
  if (budget > 0)
    for (RerootCache::Entry &e : rc.entries)
      if (e.node == v && e.children == ch && e.child_stamp == stamps && e.lengths == lengths && e.extra_max == extra.max_reticulations && e.extra == extra.configs) {
        dropCurrent(ann, rc, v);
        ann.pernode_displayed_tree_data[v] = e.data;
        rc.alias_n[v] = e.data.displayed_trees.size();
        rc.installed[v] = e.id;
        rc.state[v] = RerootCache::SHOWS_ENTRY;
        e.last_use = ++rc.tick;
        e.pinned_session = rc.session;
        for (unsigned p = 0; p < P; ++p) ann.fake_treeinfo->clv_valid[p][v] = 1;
        ann.reroot_hits++;
        return;
      }
  ann.reroot_misses++;
  dropCurrent(ann, rc, v);
  processNodeImproved(ann, 0, &ann.network.nodes[v], children, extra, false);
  // the fresh trees become an entry (which owns their slots from now on); the node keeps showing them
  RerootCache::Entry e;
  e.id = rc.next_id++;
  e.last_use = ++rc.tick;
  e.pinned_session = rc.session;
  e.node = v; e.children = ch; e.child_stamp = stamps; e.lengths = lengths; e.extra = extra.configs; e.extra_max = extra.max_reticulations;
  e.data = ann.pernode_displayed_tree_data[v];
  for (const DisplayedTreeData &d : e.data.displayed_trees) if (!d.isTip && d.slot != UINT32_MAX) rc.cached_slots++;
  rc.alias_n[v] = e.data.displayed_trees.size();
  rc.installed[v] = e.id;
  rc.state[v] = RerootCache::SHOWS_ENTRY;
  rc.entries.push_back(std::move(e));
}

void evictRerootEntries(AnnotatedNetwork &ann, RerootCache &rc, size_t budget) {
  // superseded entries first (same node + children + restrictions as a younger entry: their stamps can never match again once the
  // data they were computed from has been recomputed), then least recently used
  while (rc.cached_slots > budget) {
    size_t victim = SIZE_MAX;
    for (size_t i = 0; i < rc.entries.size(); ++i) {
      if (rc.active && rc.entries[i].pinned_session == rc.session) continue;
      if (victim == SIZE_MAX || rc.entries[i].last_use < rc.entries[victim].last_use) victim = i;
    }
    if (victim == SIZE_MAX) break;
    releaseEntry(ann, rc, rc.entries[victim]);
    rc.entries.erase(rc.entries.begin() + victim);
  }
}
}  // namespace

ReticulationConfigSet getRestrictionsActiveAliveBranch(AnnotatedNetwork &ann, size_t pmatrix_index) {  // ReticulationConfigHelper.cpp:319-331, per topology
  RerootCache &rc = rerootState(ann);
  if (rc.topology_epoch != ann.topology_epoch) { rc.plans.clear(); rc.edge_restrictions.clear(); rc.edge_all_trees.clear(); rc.topology_epoch = ann.topology_epoch; }
  for (const auto &kv : rc.edge_restrictions) if (kv.first == pmatrix_index) return kv.second;
  bool all = false;
  rc.edge_restrictions.emplace_back(pmatrix_index, computeRestrictionsActiveAliveBranch(ann, pmatrix_index, &all));
  rc.edge_all_trees.emplace_back(pmatrix_index, all);
  return rc.edge_restrictions.back().second;
}

/* ---- lazy re-rooting (AnnotatedNetwork::lazy_reroot) --------------------------------------------------------------------------------
 * optimize_branch evaluates the network from its root before it re-roots and again after the branch is done
 * (src/optimization/BranchLengthOptimization.cpp:352,420): once a length has changed, that is the whole path from the branch up to
 * the root, per branch.  None of it is needed to optimise the NEXT branch when (a) the branch is active and alive in every displayed
 * tree — then computeLoglikelihoodBrlenOpt never falls back on the per-tree lnLs of the old root (VirtualRerooting.cpp:455-544 only
 * does for trees in which the branch is inactive or dead) — and (b) the re-rooting plan of the branch is known (it depends on the
 * topology only).  What a re-rooting does read are the root-directed trees of the nodes hanging off its paths; only those (and
 * whatever is invalid below them) are brought up to date here.  A pre-order sweep therefore recomputes a node's root-directed CLV
 * once, when the sweep leaves its subtree, instead of after every branch below it. */
namespace detail {
bool lazyRerootPossible(AnnotatedNetwork &ann, size_t pmatrix_index) {
  if (!ann.lazy_reroot || !ann.reroot) return false;
  RerootCache &rc = rerootState(ann);
  if (rc.topology_epoch != ann.topology_epoch) return false;
  bool all = false, known = false;
  for (const auto &kv : rc.edge_all_trees) if (kv.first == pmatrix_index) { all = kv.second; known = true; }
  if (!known || !all) return false;
  const size_t old_vr = ann.network.root->clv_index, new_vr = ann.network.edges[pmatrix_index].source, back = ann.network.edges[pmatrix_index].target;
  for (const RerootCache::Plan &pl : rc.plans) if (pl.old_vr == old_vr && pl.new_vr == new_vr && pl.back == back) return true;
  return false;
}

/* the root-directed displayed trees a re-rooting towards `pmatrix_index` reads — the children of its path nodes that are not on
 * the path themselves — and everything invalid below them, in post-order (incremental processNodeImproved: valid nodes return) */
void validateRerootInputs(AnnotatedNetwork &ann, size_t pmatrix_index) {
  RerootCache &rc = rerootState(ann);
  const size_t old_vr = ann.network.root->clv_index, new_vr = ann.network.edges[pmatrix_index].source, back = ann.network.edges[pmatrix_index].target;
  const RerootCache::Plan *plan = nullptr;
  for (const RerootCache::Plan &pl : rc.plans) if (pl.old_vr == old_vr && pl.new_vr == new_vr && pl.back == back) { plan = &pl; break; }
  if (!plan) throw std::runtime_error("validateRerootInputs: no re-rooting plan for this branch");
  finishVirtualReroot(ann);
  std::vector<char> need(ann.network.num_nodes(), 0);
  std::vector<size_t> stack;
  for (const PathToVirtualRoot &p : plan->paths)
    for (size_t i = 0; i < p.path.size(); ++i)
      for (size_t c : p.children[i])
        if (std::find(p.path.begin(), p.path.end(), c) == p.path.end() && !need[c]) { need[c] = 1; stack.push_back(c); }
  while (!stack.empty()) {
    const size_t v = stack.back(); stack.pop_back();
    for (size_t c : ann.network.nodes[v].children) if (!need[c]) { need[c] = 1; stack.push_back(c); }
  }
  pllmod_treeinfo_update_prob_matrices(ann, 0);
  for (Node *n : ann.travbuffer) {
    if (!need[n->clv_index]) continue;
    std::vector<Node *> children;
    for (size_t c : n->children) children.push_back(&ann.network.nodes[c]);
    processNodeImproved(ann, 1, n, children, ReticulationConfigSet());
  }
  flushPendingOps(ann);
  rc.lazy_session = true;   // picked up (and reset) by the updateCLVsVirtualRerootTrees call that follows
  ann.lazy_sessions++;
}
bool rerootSessionIsLazy(const AnnotatedNetwork &ann) { return ann.reroot && ann.reroot->active && ann.reroot->lazy_session; }
}  // namespace detail

void dropRerootCache(AnnotatedNetwork &ann) {
  if (!ann.reroot) return;
  RerootCache &rc = *ann.reroot;
  std::vector<RerootCache::Entry> keep;
  for (RerootCache::Entry &e : rc.entries) {
    if (rc.active && e.pinned_session == rc.session) { keep.push_back(std::move(e)); continue; }
    releaseEntry(ann, rc, e);
  }
  rc.entries.swap(keep);
}

namespace detail {
void destroyRerootCache(AnnotatedNetwork &ann) { delete ann.reroot; ann.reroot = nullptr; }
bool rerootSessionOpen(const AnnotatedNetwork &ann) { return ann.reroot && ann.reroot->active; }
void rerootCacheSize(const AnnotatedNetwork &ann, size_t *entries, size_t *slots) {
  *entries = ann.reroot ? ann.reroot->entries.size() : 0;
  *slots = ann.reroot ? ann.reroot->cached_slots : 0;
}
/* the branches in depth-first pre-order from the root: consecutive branches share all but the last node of their re-rooting
 * paths, so a sweep in this order finds the path's re-rooted trees memoised (the reference visits an unordered_set, i.e. in no
 * particular order: src/optimization/BranchLengthOptimization.cpp:423-476) */
std::vector<size_t> branchesInPreorder(const AnnotatedNetwork &ann) {
  const Network &nw = ann.network;
  std::vector<size_t> order, stack{nw.root->clv_index};
  std::vector<char> seen(nw.num_nodes(), 0), emitted(nw.num_branches(), 0);
  seen[nw.root->clv_index] = 1;
  // iterative DFS that emits (v -> c) right before descending into c
  std::vector<std::pair<size_t, size_t>> frames{{nw.root->clv_index, 0}};
  while (!frames.empty()) {
    const size_t v = frames.back().first, k = frames.back().second;
    if (k >= nw.nodes[v].children.size()) { frames.pop_back(); continue; }
    frames.back().second++;
    const size_t c = nw.nodes[v].children[k];
    const size_t e = edgeBetween(nw, c, v);
    if (e < emitted.size() && !emitted[e]) { emitted[e] = 1; order.push_back(e); }
    if (!seen[c]) { seen[c] = 1; frames.push_back({c, 0}); }
  }
  for (size_t e = 0; e < emitted.size(); ++e) if (!emitted[e]) order.push_back(e);
  return order;
}
}  // namespace detail

void finishVirtualReroot(AnnotatedNetwork &ann) {
  if (!ann.reroot) return;
  if (!ann.reroot->active) { ann.reroot->lazy_session = false; return; }
  RerootCache &rc = *ann.reroot;
  flushPendingOps(ann);
  for (size_t v : rc.touched) {
    dropCurrent(ann, rc, v);
    ann.pernode_displayed_tree_data[v] = std::move(rc.orig[v]);
    rc.orig[v] = NodeDisplayedTreeData();
    for (unsigned p = 0; p < ann.fake_treeinfo->partition_count; ++p) ann.fake_treeinfo->clv_valid[p][v] = rc.orig_valid[v][p];
    rc.stashed[v] = 0;
    rc.state[v] = RerootCache::UNTOUCHED;
  }
  rc.touched.clear();
  rc.active = false;
  rc.lazy_session = false;
  const size_t budget = rerootBudget(ann);
  if (budget == 0 || rc.epoch != ann.clv_epoch) dropRerootCache(ann);
  else evictRerootEntries(ann, rc, budget);
  FakeTreeinfo &ti = *ann.fake_treeinfo;
  bool changed = false;
  for (unsigned p = 0; p < ti.partition_count; ++p)
    changed |= ((ti.brlen_linkage == PLLMOD_COMMON_BRLEN_UNLINKED ? ti.branch_lengths[p][rc.edge] : ti.linked_branch_lengths[rc.edge]) != rc.start_lengths[p]);
  ann.cached_logl_valid = false;   // the cached value is the edge-rooted one; evaluateTrees re-mixes the root trees on the host
  if (changed) invalidatePmatrixIndex(ann, rc.edge);
  else pllmod_treeinfo_update_prob_matrices(ann, 0);   // lengths tried in between left the edge's P-matrix flagged invalid: refresh it now,
                                                       // because an evaluation that finds every CLV valid returns before its P-matrix update
}

void updateCLVsVirtualRerootTrees(AnnotatedNetwork &ann, Node *old_virtual_root, Node *new_virtual_root,
                                  Node *new_virtual_root_back, ReticulationConfigSet &restrictions) {  // :192-252
  const size_t old_vr = old_virtual_root->clv_index, new_vr = new_virtual_root->clv_index, back = new_virtual_root_back->clv_index;
  const bool lazy = ann.reroot && !ann.reroot->active && ann.reroot->lazy_session;   // validateRerootInputs ran for this branch: the root itself may be stale
  if (!lazy && ann.pernode_displayed_tree_data[old_vr].num_active_displayed_trees == 0) throw std::runtime_error("no displayed trees at the old virtual root");
  flushPendingOps(ann);
  finishVirtualReroot(ann);   // a session left open by the caller
  RerootCache &rc = rerootState(ann);
  rc.lazy_session = lazy;
  if (rc.epoch != ann.clv_epoch) { dropRerootCache(ann); rc.epoch = ann.clv_epoch; }
  if (rc.topology_epoch != ann.topology_epoch) { rc.plans.clear(); rc.edge_restrictions.clear(); rc.edge_all_trees.clear(); rc.topology_epoch = ann.topology_epoch; }
  const RerootCache::Plan *plan = nullptr;
  for (const RerootCache::Plan &pl : rc.plans) if (pl.old_vr == old_vr && pl.new_vr == new_vr && pl.back == back) { plan = &pl; break; }
  if (!plan) {
    if (lazy) throw std::runtime_error("lazy re-rooting without a plan");
    RerootCache::Plan pl{old_vr, new_vr, back, getPathsToVirtualRoot(ann, old_vr, new_vr, back), {}};
    pl.info = computeNodeSaveInformation(pl.paths);
    rc.plans.push_back(std::move(pl));
    plan = &rc.plans.back();
  }
  const std::vector<PathToVirtualRoot> &paths = plan->paths;
  const NodeSaveInformation &info = plan->info;
  rc.active = true;
  rc.session++;
  rc.edge = edgeBetween(ann.network, new_vr, back);
  rc.start_lengths.clear();
  for (unsigned p = 0; p < ann.fake_treeinfo->partition_count; ++p)
    rc.start_lengths.push_back(ann.fake_treeinfo->brlen_linkage == PLLMOD_COMMON_BRLEN_UNLINKED ? ann.fake_treeinfo->branch_lengths[p][rc.edge]
                                                                                                  : ann.fake_treeinfo->linked_branch_lengths[rc.edge]);
  try {
    for (size_t p = 0; p < paths.size(); ++p) {
      if (!reticulationConfigsCompatible(paths[p].reticulationChoices, restrictions)) continue;
      for (size_t n : info.pathNodesToRestore[p]) showOriginal(ann, rc, n);
      for (size_t i = 0; i < paths[p].path.size(); ++i) {
        const size_t v = paths[p].path[i];
        const bool appendMode = (p > 0) && (v == new_vr);
        std::vector<Node *> children;
        for (size_t c : paths[p].children[i]) children.push_back(&ann.network.nodes[c]);
        if (!appendMode) { processPathNode(ann, rc, v, children, paths[p].reticulationChoices); continue; }
        // further paths APPEND their trees to the new virtual root: the trees already there stay aliased, the new ones get own slots
        if (v >= ann.network.num_tips()) {
          if (!rc.stashed[v]) { stashNode(ann, rc, v); showOriginal(ann, rc, v); }   // no earlier path reached it: the reference appends to the root-directed trees
          rc.state[v] = RerootCache::MIXED;
        }
        processNodeImproved(ann, 0, &ann.network.nodes[v], children, paths[p].reticulationChoices, true);
      }
    }
    flushPendingOps(ann);
  } catch (...) {
    ann.pending_ops.clear();
    std::fill(ann.pending_parent.begin(), ann.pending_parent.end(), 0);
    finishVirtualReroot(ann);
    throw;
  }
  if (ann.pernode_displayed_tree_data[new_vr].num_active_displayed_trees == 0) throw std::runtime_error("no displayed trees at the new virtual root");
}

void redoRerootFromRoot(AnnotatedNetwork &ann, unsigned int pmatrix_index, std::vector<DisplayedTreeData> &oldTrees) {
  RerootCache &rc = rerootState(ann);
  finishVirtualReroot(ann);
  for (auto &kv : rc.edge_all_trees) if (kv.first == pmatrix_index) kv.second = false;   // this branch is not a lazy one after all
  ann.lazy_fallbacks++;
  ann.lazy_last_logl = computeLoglikelihood(ann, 1, 1);
  oldTrees = extractOldTrees(ann, ann.network.root);
  ReticulationConfigSet restrictions = getRestrictionsActiveAliveBranch(ann, pmatrix_index);
  updateCLVsVirtualRerootTrees(ann, ann.network.root, &ann.network.nodes[ann.network.edges[pmatrix_index].source],
                               &ann.network.nodes[ann.network.edges[pmatrix_index].target], restrictions);
  ann.cached_logl_valid = false;
}

namespace {
const TreeLoglData &getMatchingTreeData(const std::vector<DisplayedTreeData> &trees, const ReticulationConfigSet &query, bool lazy = false) {  // ReticulationConfigHelper.cpp:290-302
  if (lazy && trees.empty()) throw LazyRerootNeedsRoot();
  for (const DisplayedTreeData &t : trees)
    if (reticulationConfigsCompatible(query, t.treeLoglData.reticulationChoices)) return t.treeLoglData;
  throw std::runtime_error("No compatible old tree data found");
}

void updateTreeData(AnnotatedNetwork &ann, const std::vector<DisplayedTreeData> &oldTrees, TreeLoglData &td) {  // VirtualRerooting.cpp:254-277
  const TreeLoglData &old = getMatchingTreeData(oldTrees, td.reticulationChoices, rerootSessionIsLazy(ann));
  td.tree_partition_logl = old.tree_partition_logl;
  td.tree_logprob = computeReticulationConfigLogProb(td.reticulationChoices, ann.first_parent_logprobs, ann.second_parent_logprobs);
  td.tree_logprob_valid = true;
  td.tree_logl_valid = old.tree_logl_valid;
}
}  // namespace

static std::vector<std::vector<SumtableInfo>> buildSumtablePairs(AnnotatedNetwork &ann, unsigned int pmatrix_index, std::vector<nrx_pair> &pairs);

/* sumtables_out != nullptr: the sumtables of the branch are made in the same pass over the pairs' CLVs (nrx_edge_lnl_sumtables) */
static double brlenOptImpl(AnnotatedNetwork &ann, const std::vector<DisplayedTreeData> &oldTrees, unsigned int pmatrix_index,
                           int update_pmatrices, std::vector<std::vector<SumtableInfo>> *sumtables_out) {  // :348-585
  if (ann.cached_logl_valid && !sumtables_out) return ann.cached_logl;
  const size_t source = ann.network.edges[pmatrix_index].source, target = ann.network.edges[pmatrix_index].target;
  NodeDisplayedTreeData &sd = ann.pernode_displayed_tree_data[source];
  NodeDisplayedTreeData &td = ann.pernode_displayed_tree_data[target];
  const size_t ns = sd.num_active_displayed_trees, nt = td.num_active_displayed_trees;
  const unsigned P = ann.fake_treeinfo->partition_count;
  if (!rerootSessionIsLazy(ann) && !clvValidCheck(ann, ann.network.root->clv_index, false))
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
  bool fused_done = false;
  if (sumtables_out) {
    std::vector<nrx_pair> pairs5;
    *sumtables_out = buildSumtablePairs(ann, pmatrix_index, pairs5);
    // every edge-lnL pair (compatible, active AND alive, interesting) is also a sumtable pair (compatible, active, interesting)
    std::vector<int32_t> lnl_index(pairs5.size(), -1);
    size_t found = 0;
    for (size_t k = 0; k < pairs.size(); ++k)
      for (size_t q = 0; q < pairs5.size(); ++q)
        if (lnl_index[q] < 0 && pairs5[q].a_kind == pairs[k].a_kind && pairs5[q].a_idx == pairs[k].a_idx && pairs5[q].b_kind == pairs[k].b_kind && pairs5[q].b_idx == pairs[k].b_idx) {
          lnl_index[q] = (int32_t)k; ++found; break;
        }
    if (found == pairs.size() && !pairs5.empty()) {
      engineCheck(nrx_edge_lnl_sumtables(ann.engine, pmatrix_index, pairs5.data(), (uint32_t)pairs5.size(), lnl_index.data(), (uint32_t)pairs.size(), out.data()), "nrx_edge_lnl_sumtables");
      fused_done = true;
    } else if (!pairs5.empty()) {
      engineCheck(nrx_sumtables(ann.engine, pairs5.data(), (uint32_t)pairs5.size()), "nrx_sumtables");
    }
  }
  if (!fused_done && !pairs.empty()) engineCheck(nrx_edge_lnl(ann.engine, pmatrix_index, pairs.data(), (uint32_t)pairs.size(), out.data()), "nrx_edge_lnl");
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

double computeLoglikelihoodBrlenOpt(AnnotatedNetwork &ann, const std::vector<DisplayedTreeData> &oldTrees, unsigned int pmatrix_index,
                                    int update_pmatrices, bool) {
  return brlenOptImpl(ann, oldTrees, pmatrix_index, update_pmatrices, nullptr);
}

double computeLoglikelihoodBrlenOptAndSumtables(AnnotatedNetwork &ann, const std::vector<DisplayedTreeData> &oldTrees, unsigned int pmatrix_index,
                                                std::vector<std::vector<SumtableInfo>> &sumtables, int update_pmatrices) {
  return brlenOptImpl(ann, oldTrees, pmatrix_index, update_pmatrices, &sumtables);
}

/* the displayed-tree pairs computePartitionSumtables makes a sumtable for, in its order (LikelihoodDerivatives.cpp:291-344) */
static std::vector<std::vector<SumtableInfo>> buildSumtablePairs(AnnotatedNetwork &ann, unsigned int pmatrix_index, std::vector<nrx_pair> &pairs) {
  const unsigned P = ann.fake_treeinfo->partition_count;
  std::vector<std::vector<SumtableInfo>> res(P);
  const size_t source = ann.network.edges[pmatrix_index].source, target = ann.network.edges[pmatrix_index].target;
  NodeDisplayedTreeData &sd = ann.pernode_displayed_tree_data[source];
  NodeDisplayedTreeData &td = ann.pernode_displayed_tree_data[target];
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
  return res;
}

std::vector<std::vector<SumtableInfo>> computePartitionSumtables(AnnotatedNetwork &ann, unsigned int pmatrix_index) {  // LikelihoodDerivatives.cpp:291-344
  std::vector<nrx_pair> pairs;
  std::vector<std::vector<SumtableInfo>> res = buildSumtablePairs(ann, pmatrix_index, pairs);
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
