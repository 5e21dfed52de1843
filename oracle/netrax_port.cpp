/*
 * netrax_port.cpp — ORACLE (test infrastructure, NOT product code).  See netrax_port.hpp.
 *
 * Follows, function by function, the reference's network-likelihood layer.  LH = src/likelihood,
 * SRC = src (under /root/reference).  All arithmetic on CLVs goes through orc::Backend.
 *
 * mpfr::mpreal at the reference's default 53-bit precision is only an exponent-range
 * extension of double; it is emulated by XD below (double mantissa + separate binary exponent).
 */
#include "netrax_port.hpp"

#include <algorithm>
#include <cassert>
#include <climits>
#include <cmath>
#include <queue>
#include <sstream>
#include <unordered_set>

namespace orc {

/* ------------------------------------------------------------------------------------------
 * XD: value = m * 2^e, m normalised by frexp; stands in for mpfr::mpreal(53 bits).
 * ---------------------------------------------------------------------------------------- */
struct XD {
  double m = 0.0;
  long e = 0;
  XD() = default;
  XD(double v) { int ex = 0; m = std::frexp(v, &ex); e = ex; }
  static XD make(double m_, long e_) { XD r; int ex = 0; r.m = std::frexp(m_, &ex); r.e = (m_ == 0.0) ? 0 : e_ + ex; return r; }
  double toDouble() const { return (m == 0.0) ? 0.0 : std::ldexp(m, (int)std::max<long>(std::min<long>(e, 100000), -100000)); }
};
static XD operator*(const XD &a, const XD &b) { return XD::make(a.m * b.m, a.e + b.e); }
static XD operator/(const XD &a, const XD &b) { return XD::make(a.m / b.m, a.e - b.e); }
static XD operator+(const XD &a, const XD &b) {
  if (a.m == 0.0) return b;
  if (b.m == 0.0) return a;
  const XD &hi = (a.e >= b.e) ? a : b;
  const XD &lo = (a.e >= b.e) ? b : a;
  long d = hi.e - lo.e;
  if (d > 1100) return hi;
  return XD::make(hi.m + std::ldexp(lo.m, (int)-d), hi.e);
}
static XD operator-(const XD &a, const XD &b) { XD nb = b; nb.m = -nb.m; return a + nb; }
static XD xexp(double x) {  // exp with unbounded exponent
  static const double LN2_HI = 6.93147180369123816490e-01, LN2_LO = 1.90821492927058770002e-10, INV_LN2 = 1.44269504088896338700e+00;
  if (x == -std::numeric_limits<double>::infinity()) return XD(0.0);
  double k = std::nearbyint(x * INV_LN2);
  double r = (x - k * LN2_HI) - k * LN2_LO;
  return XD::make(std::exp(r), (long)k);
}
static double xlog(const XD &a) {  // log(m) + e*ln2
  return std::log(a.m) + (double)a.e * 6.931471805599453094e-01;
}

/* ------------------------------------------------------------------------------------------
 * ReticulationConfigSet algebra — SRC/graph/ReticulationConfigSet.cpp
 * ---------------------------------------------------------------------------------------- */
bool ConfigSet::operator==(const ConfigSet &o) const {  // ReticulationConfigSet.hpp:21-42
  if (max_reticulations != o.max_reticulations) return false;
  if (configs.size() != o.configs.size()) return false;
  for (size_t i = 0; i < configs.size(); ++i) {
    bool found = false;
    for (size_t j = 0; j < o.configs.size(); ++j)
      if (configs[i] == o.configs[j]) { found = true; break; }
    if (!found) return false;
  }
  return true;
}

static bool choicesCompatible(const Choices &l, const Choices &r) {  // .cpp:9-25
  for (size_t i = 0; i < l.size(); ++i) {
    if (l[i] != RS::DONT_CARE)
      if (r[i] != RS::DONT_CARE && r[i] != l[i]) return false;
    if (l[i] == RS::INVALID || r[i] == RS::INVALID) return false;
  }
  return true;
}

static Choices combineChoices(const Choices &l, const Choices &r) {  // .cpp:45-60
  Choices res = l;
  for (size_t i = 0; i < res.size(); ++i) {
    if (l[i] == RS::DONT_CARE) res[i] = r[i];
    else if (r[i] == RS::DONT_CARE) res[i] = l[i];
    else if (l[i] != r[i]) res[i] = RS::INVALID;
  }
  return res;
}

static bool validChoices(const Choices &c) {  // .cpp:62-69
  for (RS s : c) if (s == RS::INVALID) return false;
  return true;
}

static double choicesLogProb(const Choices &c, const std::vector<double> &first, const std::vector<double> &second) {  // .cpp:71-88
  double lp = 0;
  for (size_t i = 0; i < c.size(); ++i) {
    if (c[i] != RS::DONT_CARE) {
      if (c[i] == RS::TAKE_FIRST_PARENT) lp += first[i];
      else lp += second[i];
    }
  }
  return lp;
}

double computeReticulationConfigLogProb(const ConfigSet &c, const std::vector<double> &first, const std::vector<double> &second) {  // .cpp:98-113
  if (c.configs.size() == 1) return choicesLogProb(c.configs[0], first, second);
  XD prob(0.0);
  for (size_t i = 0; i < c.configs.size(); ++i) prob = prob + xexp(choicesLogProb(c.configs[i], first, second));
  return xlog(prob);
}

double computeReticulationConfigProb(const ConfigSet &c, const std::vector<double> &first, const std::vector<double> &second) {  // .cpp:115-130
  if (c.configs.size() == 1) return std::exp(choicesLogProb(c.configs[0], first, second));
  XD prob(0.0);
  for (size_t i = 0; i < c.configs.size(); ++i) prob = prob + xexp(choicesLogProb(c.configs[i], first, second));
  return prob.toDouble();
}

bool reticulationConfigsCompatible(const ConfigSet &l, const ConfigSet &r) {  // .cpp:132-142
  for (size_t i = 0; i < l.configs.size(); ++i)
    for (size_t j = 0; j < r.configs.size(); ++j)
      if (choicesCompatible(l.configs[i], r.configs[j])) return true;
  return false;
}

void simplifyReticulationChoices(ConfigSet &res) {  // .cpp:151-215
  bool shrinked = true;
  while (shrinked) {
    shrinked = false;
    for (size_t i = 0; i < res.configs.size() && !shrinked; ++i)
      for (size_t j = i + 1; j < res.configs.size(); ++j)
        if (res.configs[i] == res.configs[j]) {
          std::swap(res.configs[j], res.configs[res.configs.size() - 1]);
          res.configs.pop_back();
          shrinked = true;
          break;
        }
    if (shrinked) continue;
    for (size_t i = 0; i < res.configs.size() && !shrinked; ++i) {
      for (size_t j = 0; j < res.max_reticulations && !shrinked; ++j) {
        if (res.configs[i][j] != RS::DONT_CARE) {
          Choices query = res.configs[i];
          query[j] = (res.configs[i][j] == RS::TAKE_FIRST_PARENT) ? RS::TAKE_SECOND_PARENT : RS::TAKE_FIRST_PARENT;
          for (size_t k = 0; k < res.configs.size(); ++k) {
            if (k == i) continue;
            if (res.configs[k] == query) {
              res.configs[i][j] = RS::DONT_CARE;
              std::swap(res.configs[k], res.configs[res.configs.size() - 1]);
              res.configs.pop_back();
              shrinked = true;
              break;
            }
            if (res.configs[k] == res.configs[i]) {
              std::swap(res.configs[k], res.configs[res.configs.size() - 1]);
              res.configs.pop_back();
              shrinked = true;
              break;
            }
          }
        }
      }
    }
  }
}

ConfigSet combineReticulationChoices(const ConfigSet &l, const ConfigSet &r) {  // .cpp:217-231
  ConfigSet res(l.max_reticulations);
  for (size_t i = 0; i < l.configs.size(); ++i)
    for (size_t j = 0; j < r.configs.size(); ++j) {
      Choices c = combineChoices(l.configs[i], r.configs[j]);
      if (validChoices(c)) res.configs.emplace_back(c);
    }
  simplifyReticulationChoices(res);
  return res;
}

std::string configToString(const ConfigSet &c, size_t nret) {
  std::vector<std::string> rows;
  for (const Choices &ch : c.configs) {
    std::string s;
    for (size_t i = 0; i < nret; ++i)
      s += (ch[i] == RS::TAKE_FIRST_PARENT) ? '0' : (ch[i] == RS::TAKE_SECOND_PARENT) ? '1' : '-';
    rows.push_back(s);
  }
  std::sort(rows.begin(), rows.end());
  std::string out;
  for (size_t i = 0; i < rows.size(); ++i) { if (i) out += '|'; out += rows[i]; }
  return out;
}

/* ------------------------------------------------------------------------------------------
 * Network helpers — SRC/helper/*.cpp restated on the flat network
 * ---------------------------------------------------------------------------------------- */
std::vector<unsigned> Network::Node::neighbors() const {
  std::vector<unsigned> n;
  for (unsigned p : parents) if (std::find(n.begin(), n.end(), p) == n.end()) n.push_back(p);
  for (unsigned c : children) if (std::find(n.begin(), n.end(), c) == n.end()) n.push_back(c);
  return n;
}

void Network::build(unsigned ntips, unsigned nnodes, unsigned root_, const std::vector<Edge> &edges_,
                    const std::vector<unsigned> &ret_node, const std::vector<unsigned> &ret_first_edge,
                    const std::vector<unsigned> &ret_second_edge) {
  num_tips = ntips; root = root_; edges = edges_;
  nodes.assign(nnodes, Node());
  rets.clear();
  for (size_t r = 0; r < ret_node.size(); ++r) {
    Ret R;
    R.node = ret_node[r]; R.first_edge = ret_first_edge[r]; R.second_edge = ret_second_edge[r];
    R.first_parent = edges[R.first_edge].source; R.second_parent = edges[R.second_edge].source;
    if (edges[R.first_edge].target != R.node || edges[R.second_edge].target != R.node) throw std::runtime_error("reticulation edges do not end in the reticulation node");
    if (R.first_parent == R.second_parent) throw std::runtime_error("parallel reticulation arcs are not supported");
    R.child = UINT_MAX;
    nodes[R.node].is_ret = true; nodes[R.node].ret_index = (unsigned)r;
    nodes[R.node].parents = {R.first_parent, R.second_parent};
    rets.push_back(R);
  }
  for (size_t e = 0; e < edges.size(); ++e) {  // ascending pmatrix index == children order
    const Edge &E = edges[e];
    nodes[E.source].children.push_back(E.target);
    if (!nodes[E.target].is_ret) {
      if (!nodes[E.target].parents.empty()) throw std::runtime_error("non-reticulation node with two parents");
      nodes[E.target].parents.push_back(E.source);
    }
  }
  for (Ret &R : rets) {
    if (nodes[R.node].children.size() != 1) throw std::runtime_error("reticulation node must have exactly one child");
    R.child = nodes[R.node].children[0];
  }
  toggle.assign(rets.size(), 0);
}

unsigned Network::edgeBetween(unsigned a, unsigned b) const {
  for (size_t e = 0; e < edges.size(); ++e)
    if ((edges[e].source == a && edges[e].target == b) || (edges[e].source == b && edges[e].target == a)) return (unsigned)e;
  throw std::runtime_error("no edge between nodes");
}

unsigned Network::activeParent(unsigned n) const {  // ParentHelper.cpp:5-16
  const Node &N = nodes[n];
  if (N.is_ret) return toggle[N.ret_index] ? rets[N.ret_index].second_parent : rets[N.ret_index].first_parent;
  return N.parents.empty() ? UINT_MAX : N.parents[0];
}

std::vector<unsigned> Network::activeAliveChildren(const std::vector<bool> &dead, unsigned n) const {  // ChildrenHelper.cpp:58-80
  std::vector<unsigned> res;
  for (unsigned c : nodes[n].children) {
    if (dead[c]) continue;
    if (nodes[c].is_ret && activeParent(c) != n) continue;
    res.push_back(c);
  }
  return res;
}

std::vector<unsigned> Network::activeNeighbors(unsigned n) const {  // NeighborHelper.cpp:19-43
  std::vector<unsigned> res;
  for (unsigned nb : nodes[n].neighbors()) {
    if (nodes[nb].is_ret)
      if (n != rets[nodes[nb].ret_index].child && activeParent(nb) != n) continue;
    if (nodes[n].is_ret && nb != rets[nodes[n].ret_index].child)
      if (nb != activeParent(n)) continue;
    res.push_back(nb);
  }
  return res;
}

std::vector<bool> Network::collectDeadNodes(unsigned megablobRoot, unsigned *dtr) const {  // NetworkFunctions.cpp:228-279
  std::vector<bool> dead(nodes.size(), false);
  std::queue<unsigned> q;
  for (size_t i = 0; i < rets.size(); ++i) q.push(toggle[i] ? rets[i].first_parent : rets[i].second_parent);  // non-active parent
  while (!q.empty()) {
    unsigned u = q.front(); q.pop();
    if (activeAliveChildren(dead, u).empty()) {
      dead[u] = true;
      if (nodes[u].is_ret) { q.push(rets[nodes[u].ret_index].first_parent); q.push(rets[nodes[u].ret_index].second_parent); }
      else { unsigned p = activeParent(u); if (p != UINT_MAX) q.push(p); }
    }
  }
  unsigned dtroot = root;
  std::vector<unsigned> ch = activeAliveChildren(dead, dtroot);
  bool seenMegablobRoot = false;
  while (ch.size() == 1) {
    if (dtroot == megablobRoot) seenMegablobRoot = true;
    dead[dtroot] = true;
    dtroot = ch[0];
    ch = activeAliveChildren(dead, dtroot);
  }
  if (dtr) *dtr = seenMegablobRoot ? dtroot : megablobRoot;
  return dead;
}

std::vector<unsigned> Network::reversedTopologicalSort() const {  // NetworkFunctions.cpp:542-598 (Kahn on out-degrees)
  std::vector<unsigned> res, outdeg(nodes.size(), 0);
  std::queue<unsigned> q;
  for (size_t i = 0; i < nodes.size(); ++i) { outdeg[i] = (unsigned)nodes[i].children.size(); if (!outdeg[i]) q.push((unsigned)i); }
  while (!q.empty()) {
    unsigned a = q.front(); q.pop();
    res.push_back(a);
    for (unsigned p : nodes[a].parents) if (--outdeg[p] == 0) q.push(p);
  }
  if (res.size() != nodes.size()) throw std::runtime_error("Cycle in network detected");
  return res;
}

/* ---- SRC/helper/ReticulationConfigHelper.cpp ---------------------------------------------- */
static ConfigSet getRestrictionsToDismissNeighbor(AnnotatedNetwork &ann, unsigned node, unsigned neighbor) {  // :26-63
  const Network &nw = ann.network;
  ConfigSet res(ann.options.max_reticulations);
  Choices r(ann.options.max_reticulations, RS::DONT_CARE);
  bool found = false;
  if (nw.nodes[node].is_ret) {
    const Network::Ret &R = nw.rets[nw.nodes[node].ret_index];
    if (neighbor == R.first_parent) { r[nw.nodes[node].ret_index] = RS::TAKE_SECOND_PARENT; found = true; }
    else if (neighbor == R.second_parent) { r[nw.nodes[node].ret_index] = RS::TAKE_FIRST_PARENT; found = true; }
  }
  if (nw.nodes[neighbor].is_ret) {
    const Network::Ret &R = nw.rets[nw.nodes[neighbor].ret_index];
    if (node == R.first_parent) { r[nw.nodes[neighbor].ret_index] = RS::TAKE_SECOND_PARENT; found = true; }
    else if (node == R.second_parent) { r[nw.nodes[neighbor].ret_index] = RS::TAKE_FIRST_PARENT; found = true; }
  }
  if (found) res.configs.emplace_back(r);
  return res;
}

static ConfigSet getRestrictionsToTakeNeighbor(AnnotatedNetwork &ann, unsigned node, unsigned neighbor) {  // :65-96
  const Network &nw = ann.network;
  ConfigSet res(ann.options.max_reticulations);
  Choices r(ann.options.max_reticulations, RS::DONT_CARE);
  if (nw.nodes[node].is_ret) {
    const Network::Ret &R = nw.rets[nw.nodes[node].ret_index];
    if (neighbor == R.first_parent) r[nw.nodes[node].ret_index] = RS::TAKE_FIRST_PARENT;
    else if (neighbor == R.second_parent) r[nw.nodes[node].ret_index] = RS::TAKE_SECOND_PARENT;
  }
  if (nw.nodes[neighbor].is_ret) {
    const Network::Ret &R = nw.rets[nw.nodes[neighbor].ret_index];
    if (node == R.first_parent) r[nw.nodes[neighbor].ret_index] = RS::TAKE_FIRST_PARENT;
    else if (node == R.second_parent) r[nw.nodes[neighbor].ret_index] = RS::TAKE_SECOND_PARENT;
  }
  res.configs.emplace_back(r);
  return res;
}

static ConfigSet getTreeConfig(AnnotatedNetwork &ann, size_t tree_idx) {  // :196-211
  ConfigSet c(ann.options.max_reticulations);
  c.configs.emplace_back(Choices(ann.options.max_reticulations, RS::DONT_CARE));
  for (size_t i = 0; i < ann.network.num_reticulations(); ++i)
    c.configs[0][i] = (tree_idx & ((size_t)1 << i)) ? RS::TAKE_SECOND_PARENT : RS::TAKE_FIRST_PARENT;
  return c;
}

static ConfigSet getReticulationChoicesThisOnly(AnnotatedNetwork &ann, const ConfigSet &this_tree_config,
                                                const ConfigSet &other_child_dead_settings, unsigned parent,
                                                unsigned this_child, unsigned other_child) {  // :98-144
  ConfigSet res(ann.options.max_reticulations);
  ConfigSet restrictedConfig = combineReticulationChoices(this_tree_config, getRestrictionsToTakeNeighbor(ann, parent, this_child));
  if (restrictedConfig.configs.empty()) return res;
  ConfigSet combinedConfig = combineReticulationChoices(restrictedConfig, getRestrictionsToDismissNeighbor(ann, parent, other_child));
  for (const Choices &c : combinedConfig.configs) res.configs.emplace_back(c);
  restrictedConfig = combineReticulationChoices(restrictedConfig, getRestrictionsToTakeNeighbor(ann, parent, other_child));
  ConfigSet combinedConfig2 = combineReticulationChoices(restrictedConfig, other_child_dead_settings);
  for (const Choices &c : combinedConfig2.configs) res.configs.emplace_back(c);
  simplifyReticulationChoices(res);
  return res;
}

static ConfigSet deadNodeSettings(AnnotatedNetwork &ann, const NodeDisplayedTreeData &dt, unsigned parent, unsigned child) {  // :146-194
  ConfigSet res(ann.options.max_reticulations);
  ConfigSet notTaken = getRestrictionsToDismissNeighbor(ann, parent, child);
  for (const Choices &c : notTaken.configs) res.configs.emplace_back(c);
  ConfigSet taken = getRestrictionsToTakeNeighbor(ann, parent, child);
  size_t max_n_trees = (size_t)1 << ann.network.num_reticulations();
  for (size_t t = 0; t < max_n_trees; ++t) {
    ConfigSet rc = getTreeConfig(ann, t);
    if (!reticulationConfigsCompatible(rc, taken)) continue;
    bool found = false;
    for (size_t i = 0; i < dt.num_active_displayed_trees; ++i)
      if (reticulationConfigsCompatible(rc, dt.displayed_trees[i].treeLoglData.reticulationChoices)) { found = true; break; }
    if (!found) res.configs.emplace_back(rc.configs[0]);
  }
  simplifyReticulationChoices(res);
  return res;
}

static void setReticulationParents(Network &nw, const Choices &c) {  // ReticulationHelper.cpp:142-157
  for (size_t i = 0; i < nw.num_reticulations(); ++i) {
    if (c[i] == RS::TAKE_FIRST_PARENT) nw.toggle[i] = 0;
    else if (c[i] == RS::TAKE_SECOND_PARENT) nw.toggle[i] = 1;
  }
}

static unsigned findFirstNodeWithTwoActiveChildren(AnnotatedNetwork &ann, const ConfigSet &rc, unsigned oldRoot) {  // :251-272
  setReticulationParents(ann.network, rc.configs[0]);  // setReticulationState skips DONT_CARE: identical effect
  unsigned dtr = oldRoot;
  ann.network.collectDeadNodes(oldRoot, &dtr);
  return dtr;
}

static DisplayedTreeData &findMatchingDisplayedTree(AnnotatedNetwork &, const ConfigSet &rc, NodeDisplayedTreeData &data) {  // :213-249
  DisplayedTreeData *tree = nullptr;
  size_t n_good = 0;
  for (size_t i = 0; i < data.num_active_displayed_trees; ++i)
    if (reticulationConfigsCompatible(rc, data.displayed_trees[i].treeLoglData.reticulationChoices)) { n_good++; tree = &data.displayed_trees[i]; }
  if (n_good == 1) return *tree;
  if (n_good > 1) throw std::runtime_error("Found multiple suitable trees");
  throw std::runtime_error("Found no suitable displayed tree");
}

static const TreeLoglData &getMatchingTreeData(const std::vector<DisplayedTreeData> &trees, const ConfigSet &query) {  // :290-302
  for (size_t i = 0; i < trees.size(); ++i)
    if (reticulationConfigsCompatible(query, trees[i].treeLoglData.reticulationChoices)) return trees[i].treeLoglData;
  throw std::runtime_error("No compatible old tree data found");
}

static bool isActiveBranch(AnnotatedNetwork &ann, const ConfigSet &rc, unsigned pmatrix_index) {  // EdgeHelper.cpp:84-95
  const Network::Edge &E = ann.network.edges[pmatrix_index];
  return reticulationConfigsCompatible(getRestrictionsToTakeNeighbor(ann, E.source, E.target), rc);
}

static bool isActiveAliveBranch(AnnotatedNetwork &ann, const ConfigSet &rc, unsigned pmatrix_index) {  // EdgeHelper.cpp:97-121
  setReticulationParents(ann.network, rc.configs[0]);
  std::vector<bool> dead = ann.network.collectDeadNodes(ann.network.root, nullptr);
  const Network::Edge &E = ann.network.edges[pmatrix_index];
  return reticulationConfigsCompatible(getRestrictionsToTakeNeighbor(ann, E.source, E.target), rc) && !dead[E.source] && !dead[E.target];
}

ConfigSet getRestrictionsActiveAliveBranch(AnnotatedNetwork &ann, size_t pmatrix_index) {  // ReticulationConfigHelper.cpp:319-331
  ConfigSet res(ann.options.max_reticulations);  // NB: the reference leaves max_reticulations at 0 here, which
                                                 // makes simplifyReticulationChoices merge nothing but duplicates.
  res.max_reticulations = 0;
  for (size_t t = 0; t < ((size_t)1 << ann.network.num_reticulations()); ++t) {
    ConfigSet tc = getTreeConfig(ann, t);
    if (isActiveAliveBranch(ann, tc, (unsigned)pmatrix_index)) res.configs.emplace_back(tc.configs[0]);
  }
  simplifyReticulationChoices(res);
  return res;
}

/* ---- SRC/helper/InvalidationHelper.cpp ---------------------------------------------------- */
void invalidateSingleClv(AnnotatedNetwork &ann, unsigned clv_index) {  // :8-39
  for (unsigned p = 0; p < ann.partitionCount(); ++p) ann.clv_valid[p][clv_index] = 0;
  NodeDisplayedTreeData &nd = ann.pernode_displayed_tree_data[clv_index];
  for (size_t i = 0; i < nd.num_active_displayed_trees; ++i) {
    nd.displayed_trees[i].clv_valid = false;
    nd.displayed_trees[i].treeLoglData.tree_logl_valid = false;
  }
  nd.num_active_displayed_trees = 0;
  if (clv_index < ann.pseudo_clv_valid.size()) ann.pseudo_clv_valid[clv_index] = 0;  // InvalidationHelper.cpp:37
  ann.cached_logl_valid = false;
}

static void validateSingleClv(AnnotatedNetwork &ann, unsigned clv_index) {  // :41-50
  for (unsigned p = 0; p < ann.partitionCount(); ++p) ann.clv_valid[p][clv_index] = 1;
}

void invalidateHigherCLVs(AnnotatedNetwork &ann, unsigned node, bool invalidate_myself) {  // :52-87 (noVisited variant)
  Network &nw = ann.network;
  if (node == UINT_MAX) return;
  if (node < nw.num_tips) invalidate_myself = false;
  if (invalidate_myself) invalidateSingleClv(ann, node);
  if (node == nw.root) return;
  if (nw.nodes[node].is_ret) {
    invalidateHigherCLVs(ann, nw.rets[nw.nodes[node].ret_index].first_parent, true);
    invalidateHigherCLVs(ann, nw.rets[nw.nodes[node].ret_index].second_parent, true);
  } else {
    invalidateHigherCLVs(ann, nw.activeParent(node), true);
  }
  ann.cached_logl_valid = false;
}

void invalidatePmatrixIndex(AnnotatedNetwork &ann, size_t pmatrix_index) {  // :157-177
  for (unsigned p = 0; p < ann.partitionCount(); ++p) ann.pmatrix_valid[p][pmatrix_index] = 0;
  invalidateHigherCLVs(ann, ann.network.edges[pmatrix_index].source, true);
}

void invalidPmatrixIndexOnly(AnnotatedNetwork &ann, size_t pmatrix_index) {  // :179-192
  for (unsigned p = 0; p < ann.partitionCount(); ++p) ann.pmatrix_valid[p][pmatrix_index] = 0;
  ann.cached_logl_valid = false;
}

static bool allClvsValid(AnnotatedNetwork &ann, size_t clv_index) {  // :194-237 (no interesting-tree restriction)
  for (unsigned p = 0; p < ann.partitionCount(); ++p) if (!ann.clv_valid[p][clv_index]) return false;
  NodeDisplayedTreeData &nd = ann.pernode_displayed_tree_data[clv_index];
  if (nd.num_active_displayed_trees == 0) return false;
  for (size_t i = 0; i < nd.num_active_displayed_trees; ++i) {
    DisplayedTreeData &dtd = nd.displayed_trees[i];
    if (!dtd.treeLoglData.tree_logprob_valid) {
      dtd.treeLoglData.tree_logprob = computeReticulationConfigLogProb(dtd.treeLoglData.reticulationChoices, ann.first_parent_logprobs, ann.second_parent_logprobs);
      dtd.treeLoglData.tree_logprob_valid = true;
    }
    if (!dtd.clv_valid && dtd.treeLoglData.tree_logprob < ann.options.min_interesting_tree_logprob) return false;
  }
  return true;
}

void invalidateAllCLVs(AnnotatedNetwork &ann) {  // :255-260
  for (size_t i = ann.network.num_tips; i < ann.network.num_nodes(); ++i) invalidateSingleClv(ann, (unsigned)i);
}

void invalidateTreeLogprobs(AnnotatedNetwork &ann) {  // :273-308 (every reticulation)
  for (size_t r = 0; r < ann.network.num_reticulations(); ++r)
    for (size_t i = 0; i < ann.network.num_nodes(); ++i) {
      NodeDisplayedTreeData &nd = ann.pernode_displayed_tree_data[i];
      for (size_t j = 0; j < nd.num_active_displayed_trees; ++j) {
        DisplayedTreeData &dtd = nd.displayed_trees[j];
        bool has = false;
        for (const Choices &c : dtd.treeLoglData.reticulationChoices.configs) if (c[r] != RS::DONT_CARE) has = true;
        if (has) {
          dtd.treeLoglData.tree_logprob = computeReticulationConfigLogProb(dtd.treeLoglData.reticulationChoices, ann.first_parent_logprobs, ann.second_parent_logprobs);
          dtd.treeLoglData.tree_logprob_valid = true;
        }
      }
    }
}

void setReticulationProb(AnnotatedNetwork &ann, size_t r, double prob) {  // SRC/optimization/ReticulationOptimization.cpp:25-38
  if (ann.reticulation_probs[r] == prob) return;
  ann.reticulation_probs[r] = prob;
  ann.first_parent_logprobs[r] = std::log(prob);
  ann.second_parent_logprobs[r] = std::log(1.0 - prob);
  ann.cached_logl_valid = false;
  invalidateTreeLogprobs(ann);
  if (ann.options.likelihood_variant == LikelihoodVariant::SARAH_PSEUDO)  // InvalidationHelper.cpp:296-301
    invalidateHigherCLVs(ann, ann.network.rets[r].node, false);
}

void setBranchLength(AnnotatedNetwork &ann, int partition, size_t pmatrix_index, double value) {
  if (partition < 0) {
    ann.linked_branch_lengths[pmatrix_index] = value;
    for (unsigned p = 0; p < ann.partitionCount(); ++p) ann.branch_lengths[p][pmatrix_index] = value;
  } else {
    ann.branch_lengths[partition][pmatrix_index] = value;
  }
}

void updateProbMatrices(AnnotatedNetwork &ann, int update_all) {  // PLLMOD/tree/treeinfo.c:842-880
  for (unsigned p = 0; p < ann.partitionCount(); ++p)
    for (size_t m = 0; m < ann.network.edges.size() + 1; ++m) {
      if (ann.pmatrix_valid[p][m] && !update_all) continue;
      double p_brlen = ann.branch_lengths[p][m];
      if (ann.options.brlen_linkage == BRLEN_SCALED && p < ann.brlen_scalers.size()) p_brlen *= ann.brlen_scalers[p];  // :862-864
      ann.backend->updatePmatrix(p, (unsigned)m, p_brlen);
      ann.pmatrix_valid[p][m] = 1;
    }
}

/* ---- SRC/graph/AnnotatedNetwork.cpp:80-185, 340-394 ----------------------------------------- */
void init_annotated_network(AnnotatedNetwork &ann) {
  Network &nw = ann.network;
  const unsigned P = ann.partitionCount();
  ann.travbuffer = nw.reversedTopologicalSort();
  size_t R = nw.num_reticulations();
  ann.reticulation_probs.assign(ann.options.max_reticulations, 0.5);
  ann.first_parent_logprobs.assign(ann.options.max_reticulations, std::log(0.5));
  ann.second_parent_logprobs.assign(ann.options.max_reticulations, std::log(0.5));
  for (size_t i = 0; i < R; ++i) {
    double pr = nw.edges[nw.rets[i].first_edge].prob;
    ann.reticulation_probs[i] = pr;
    ann.first_parent_logprobs[i] = std::log(pr);
    ann.second_parent_logprobs[i] = std::log(1.0 - pr);
  }
  ann.clv_valid.assign(P, std::vector<char>(nw.num_nodes(), 0));
  ann.pmatrix_valid.assign(P, std::vector<char>(nw.edges.size() + 1, 0));
  // fake_init_collect_branch_lengths, SRC/RaxmlWrapper.cpp:514-537: fake branch length 0
  ann.linked_branch_lengths.assign(nw.edges.size() + 1, 0.0);
  for (size_t e = 0; e < nw.edges.size(); ++e) ann.linked_branch_lengths[e] = nw.edges[e].length;
  if (ann.branch_lengths.size() != P) ann.branch_lengths.assign(P, ann.linked_branch_lengths);
  ann.partition_loglh.assign(P, 0.0);
  updateProbMatrices(ann, 1);
  for (unsigned p = 0; p < P; ++p) for (unsigned j = 0; j < nw.num_tips; ++j) ann.clv_valid[p][j] = 1;
  ann.pernode_displayed_tree_data.assign(nw.num_nodes(), NodeDisplayedTreeData());
  for (unsigned i = 0; i < nw.num_tips; ++i) {
    DisplayedTreeData d;
    d.treeLoglData = TreeLoglData(P, ann.options.max_reticulations);
    d.treeLoglData.reticulationChoices.configs.emplace_back(Choices(ann.options.max_reticulations, RS::DONT_CARE));  // DisplayedTreeData tip ctor
    d.isTip = true; d.tip = i; d.clv_valid = true;
    d.clv_vector.resize(P); d.scale_buffer.resize(P);
    ann.pernode_displayed_tree_data[i].displayed_trees.emplace_back(std::move(d));
    ann.pernode_displayed_tree_data[i].num_active_displayed_trees = 1;
  }
  ann.cached_logl_valid = false;
}

static bool clvValidCheck(AnnotatedNetwork &ann, size_t vroot, bool care_about_trees = true) {  // :340-357
  if (care_about_trees && ann.pernode_displayed_tree_data[vroot].num_active_displayed_trees == 0) return false;
  bool ok = true;
  for (unsigned p = 0; p < ann.partitionCount(); ++p) ok &= (bool)ann.clv_valid[p][vroot];
  return ok;
}

static bool reuseOldDisplayedTreesCheck(AnnotatedNetwork &ann, int incremental, size_t vroot) {  // :359-394
  if (!incremental) return false;
  if (!clvValidCheck(ann, vroot)) return false;
  NodeDisplayedTreeData &nd = ann.pernode_displayed_tree_data[vroot];
  for (size_t i = 0; i < nd.num_active_displayed_trees; ++i) {
    DisplayedTreeData &dtd = nd.displayed_trees[i];
    if (!dtd.treeLoglData.tree_logprob_valid) {
      dtd.treeLoglData.tree_logprob = computeReticulationConfigLogProb(dtd.treeLoglData.reticulationChoices, ann.first_parent_logprobs, ann.second_parent_logprobs);
      dtd.treeLoglData.tree_logprob_valid = true;
    }
    if ((!dtd.clv_valid || !dtd.treeLoglData.tree_logl_valid) && dtd.treeLoglData.tree_logprob >= ann.options.min_interesting_tree_logprob) return false;
  }
  return true;
}

/* ------------------------------------------------------------------------------------------
 * LH/ImprovedLoglikelihood.cpp
 * ---------------------------------------------------------------------------------------- */
static DisplayedTreeData *findDisplayedTree(AnnotatedNetwork &ann, size_t clv_index, const ConfigSet &rc) {  // :15-28
  NodeDisplayedTreeData &nd = ann.pernode_displayed_tree_data[clv_index];
  for (size_t i = 0; i < nd.num_active_displayed_trees; ++i)
    if (nd.displayed_trees[i].treeLoglData.reticulationChoices == rc) return &nd.displayed_trees[i];
  return nullptr;
}

static bool tree_already_present_and_fine(AnnotatedNetwork &ann, size_t clv_index, const ConfigSet &rc) {  // :30-57
  DisplayedTreeData *dtd = findDisplayedTree(ann, clv_index, rc);
  if (!dtd) return false;
  if (!dtd->treeLoglData.tree_logprob_valid) {
    dtd->treeLoglData.tree_logprob = computeReticulationConfigLogProb(dtd->treeLoglData.reticulationChoices, ann.first_parent_logprobs, ann.second_parent_logprobs);
    dtd->treeLoglData.tree_logprob_valid = true;
  }
  return dtd->clv_valid || dtd->treeLoglData.tree_logprob < ann.options.min_interesting_tree_logprob;
}

static DisplayedTreeData &add_displayed_tree(AnnotatedNetwork &ann, size_t clv_index, const ConfigSet &rc) {  // :59-113
  DisplayedTreeData *dtd = findDisplayedTree(ann, clv_index, rc);
  if (dtd) return *dtd;
  NodeDisplayedTreeData &data = ann.pernode_displayed_tree_data[clv_index];
  data.num_active_displayed_trees++;
  const unsigned P = ann.partitionCount();
  if (data.num_active_displayed_trees > data.displayed_trees.size()) {
    DisplayedTreeData d;
    d.treeLoglData = TreeLoglData(P, ann.options.max_reticulations);
    d.clv_vector.resize(P); d.scale_buffer.resize(P);
    for (unsigned p = 0; p < P; ++p) {
      d.clv_vector[p].alloc(ann.backend->clvEntries(p));
      d.scale_buffer[p].alloc(ann.backend->sites(p));
    }
    data.displayed_trees.emplace_back(std::move(d));
  } else {
    DisplayedTreeData &d = data.displayed_trees[data.num_active_displayed_trees - 1];
    for (unsigned p = 0; p < P; ++p) { d.clv_vector[p].zero(); d.scale_buffer[p].zero(); }
  }
  DisplayedTreeData &tree = data.displayed_trees[data.num_active_displayed_trees - 1];
  tree.clv_valid = false;
  tree.treeLoglData.reticulationChoices = rc;
  tree.treeLoglData.tree_logprob = computeReticulationConfigLogProb(rc, ann.first_parent_logprobs, ann.second_parent_logprobs);
  return tree;
}

struct Op {  // pll_operation_t as built by LH/Operation.cpp:7-35
  unsigned parent;
  bool has1, has2;
  unsigned child1, child2, matrix1, matrix2;
};

static Op buildOperation(Network &nw, unsigned parent, int child1, int child2, unsigned fake_pmatrix) {
  Op op;
  op.parent = parent;
  op.has1 = child1 >= 0; op.has2 = child2 >= 0;
  op.child1 = op.has1 ? (unsigned)child1 : 0; op.matrix1 = op.has1 ? nw.edgeBetween((unsigned)child1, parent) : fake_pmatrix;
  op.child2 = op.has2 ? (unsigned)child2 : 0; op.matrix2 = op.has2 ? nw.edgeBetween((unsigned)child2, parent) : fake_pmatrix;
  return op;
}

static Operand makeOperand(const DisplayedTreeData &t, unsigned p, unsigned edge) {
  Operand o;
  if (t.isTip) { o.kind = 1; o.tip = t.tip; }
  else { o.kind = 0; o.clv = t.clv_vector[p].p; o.scaler = t.scale_buffer[p].p; }
  o.edge = edge;
  return o;
}

static void add_tree_single(AnnotatedNetwork &ann, size_t clv_index, const Op &op, unsigned child_node, size_t child_tree_idx, const ConfigSet &rc) {  // :115-152
  if (tree_already_present_and_fine(ann, clv_index, rc)) return;
  DisplayedTreeData &tree = add_displayed_tree(ann, clv_index, rc);
  // (re-fetch the child after a possible reallocation of the parent's vector: different nodes, safe)
  const DisplayedTreeData &childTree = ann.pernode_displayed_tree_data[child_node].displayed_trees[child_tree_idx];
  for (unsigned p = 0; p < ann.partitionCount(); ++p) {
    Operand l = makeOperand(childTree, p, op.matrix1);
    Operand r;  // fake clv, fake pmatrix, no scaler
    r.kind = 2; r.edge = ann.fakePmatrixIndex();
    ann.backend->updatePartials(p, tree.clv_vector[p].p, tree.scale_buffer[p].p, l, r);
    ann.n_clv_updates += ann.backend->sites(p);
  }
  tree.clv_valid = true;
}

static void add_tree_both(AnnotatedNetwork &ann, size_t clv_index, const Op &op, unsigned lnode, size_t li, unsigned rnode, size_t ri, const ConfigSet &rc) {  // :154-192
  if (tree_already_present_and_fine(ann, clv_index, rc)) return;
  DisplayedTreeData &tree = add_displayed_tree(ann, clv_index, rc);
  const DisplayedTreeData &lt = ann.pernode_displayed_tree_data[lnode].displayed_trees[li];
  const DisplayedTreeData &rt = ann.pernode_displayed_tree_data[rnode].displayed_trees[ri];
  for (unsigned p = 0; p < ann.partitionCount(); ++p) {
    Operand l = makeOperand(lt, p, op.matrix1), r = makeOperand(rt, p, op.matrix2);
    ann.backend->updatePartials(p, tree.clv_vector[p].p, tree.scale_buffer[p].p, l, r);
    ann.n_clv_updates += ann.backend->sites(p);
  }
  tree.clv_valid = true;
}

static void processNodeImprovedSingleChild(AnnotatedNetwork &ann, unsigned node, unsigned child, const ConfigSet &extra) {  // :194-223
  Op op = buildOperation(ann.network, node, (int)child, -1, ann.fakePmatrixIndex());
  ConfigSet restrictionsSet = getRestrictionsToTakeNeighbor(ann, node, child);
  if (!extra.configs.empty()) restrictionsSet = combineReticulationChoices(restrictionsSet, extra);
  NodeDisplayedTreeData &dc = ann.pernode_displayed_tree_data[child];
  for (size_t i = 0; i < dc.num_active_displayed_trees; ++i) {
    const ConfigSet &crc = dc.displayed_trees[i].treeLoglData.reticulationChoices;
    if (reticulationConfigsCompatible(crc, restrictionsSet)) {
      ConfigSet rc = combineReticulationChoices(crc, restrictionsSet);
      add_tree_single(ann, node, op, child, i, rc);
    }
  }
}

static void processNodeImprovedTwoChildren(AnnotatedNetwork &ann, unsigned node, unsigned left, unsigned right, const ConfigSet &extra) {  // :225-345
  Network &nw = ann.network;
  Op op_both = buildOperation(nw, node, (int)left, (int)right, ann.fakePmatrixIndex());
  ConfigSet both = getRestrictionsToTakeNeighbor(ann, node, left);
  both = combineReticulationChoices(both, getRestrictionsToTakeNeighbor(ann, node, right));
  if (!extra.configs.empty()) both = combineReticulationChoices(both, extra);
  NodeDisplayedTreeData &dl = ann.pernode_displayed_tree_data[left];
  NodeDisplayedTreeData &dr = ann.pernode_displayed_tree_data[right];

  for (size_t i = 0; i < dl.num_active_displayed_trees; ++i) {
    if (!reticulationConfigsCompatible(dl.displayed_trees[i].treeLoglData.reticulationChoices, both)) continue;
    for (size_t j = 0; j < dr.num_active_displayed_trees; ++j) {
      if (!reticulationConfigsCompatible(dr.displayed_trees[j].treeLoglData.reticulationChoices, both)) continue;
      if (reticulationConfigsCompatible(dl.displayed_trees[i].treeLoglData.reticulationChoices, dr.displayed_trees[j].treeLoglData.reticulationChoices)) {
        ConfigSet rc = combineReticulationChoices(dl.displayed_trees[i].treeLoglData.reticulationChoices, dr.displayed_trees[j].treeLoglData.reticulationChoices);
        rc = combineReticulationChoices(rc, both);
        add_tree_both(ann, node, op_both, left, i, right, j, rc);
      }
    }
  }

  Op op_left_only = buildOperation(nw, node, (int)left, -1, ann.fakePmatrixIndex());
  ConfigSet right_dead = deadNodeSettings(ann, dr, node, right);
  if (!extra.configs.empty()) right_dead = combineReticulationChoices(right_dead, extra);
  for (size_t i = 0; i < dl.num_active_displayed_trees; ++i) {
    ConfigSet lo = getReticulationChoicesThisOnly(ann, dl.displayed_trees[i].treeLoglData.reticulationChoices, right_dead, node, left, right);
    if (!extra.configs.empty()) lo = combineReticulationChoices(lo, extra);
    if (!lo.configs.empty()) add_tree_single(ann, node, op_left_only, left, i, lo);
  }

  Op op_right_only = buildOperation(nw, node, (int)right, -1, ann.fakePmatrixIndex());
  ConfigSet left_dead = deadNodeSettings(ann, dl, node, left);
  if (!extra.configs.empty()) left_dead = combineReticulationChoices(left_dead, extra);
  for (size_t i = 0; i < dr.num_active_displayed_trees; ++i) {
    ConfigSet ro = getReticulationChoicesThisOnly(ann, dr.displayed_trees[i].treeLoglData.reticulationChoices, left_dead, node, right, left);
    if (!extra.configs.empty()) ro = combineReticulationChoices(ro, extra);
    if (!ro.configs.empty()) add_tree_single(ann, node, op_right_only, right, i, ro);
  }
}

static void processNodeImproved(AnnotatedNetwork &ann, int incremental, unsigned node, const std::vector<unsigned> &children, const ConfigSet &extra, bool append = false) {  // :347-408
  if (node < ann.network.num_tips) return;
  if (incremental && allClvsValid(ann, node)) return;
  if (!append) ann.pernode_displayed_tree_data[node].num_active_displayed_trees = 0;
  if (children.empty()) { validateSingleClv(ann, node); return; }
  if (children.size() == 1) processNodeImprovedSingleChild(ann, node, children[0], extra);
  else processNodeImprovedTwoChildren(ann, node, children[0], children[1], extra);
  for (unsigned p = 0; p < ann.partitionCount(); ++p) ann.clv_valid[p][node] = 1;
  if (ann.pernode_displayed_tree_data[node].num_active_displayed_trees > ((size_t)1 << ann.network.num_reticulations()))
    throw std::runtime_error("Too many displayed trees stored at node");
  validateSingleClv(ann, node);
}

static void computeDisplayedTreeLoglikelihood(AnnotatedNetwork &ann, DisplayedTreeData &treeAtRoot, unsigned actRoot) {  // :410-486
  if (!treeAtRoot.treeLoglData.tree_logprob_valid) {
    treeAtRoot.treeLoglData.tree_logprob = computeReticulationConfigLogProb(treeAtRoot.treeLoglData.reticulationChoices, ann.first_parent_logprobs, ann.second_parent_logprobs);
    treeAtRoot.treeLoglData.tree_logprob_valid = true;
  }
  if (treeAtRoot.treeLoglData.tree_logprob < ann.options.min_interesting_tree_logprob) return;
  unsigned dtr = findFirstNodeWithTwoActiveChildren(ann, treeAtRoot.treeLoglData.reticulationChoices, actRoot);
  DisplayedTreeData &t = findMatchingDisplayedTree(ann, treeAtRoot.treeLoglData.reticulationChoices, ann.pernode_displayed_tree_data[dtr]);
  for (unsigned p = 0; p < ann.partitionCount(); ++p) {
    std::vector<double> persite(ann.backend->sites(p), 0.0);
    double tl = ann.backend->rootLogl(p, t.clv_vector[p].p, t.scale_buffer[p].p, persite.data());
    treeAtRoot.treeLoglData.tree_partition_logl[p] = tl;
  }
  if (ann.parallel_reduce_cb)  // LH/ImprovedLoglikelihood.cpp:472-483 (one reduce per displayed tree)
    ann.parallel_reduce_cb(ann.parallel_context, treeAtRoot.treeLoglData.tree_partition_logl.data(), ann.partitionCount(), 0);
  treeAtRoot.treeLoglData.tree_logl_valid = true;
}

/* Per-site lnL of root displayed tree `tree` in partition p: the `persite_lnl` array the reference hands to
 * pll_compute_root_loglikelihood and then discards (LH/ImprovedLoglikelihood.cpp:448-453; filled per pattern, pattern weight
 * applied, LIBPLL/core_likelihood.c:190-200).  Returns the partition sum. */
double persiteLoglikelihood(AnnotatedNetwork &ann, size_t tree, unsigned p, double *persite) {
  NodeDisplayedTreeData &rd = ann.pernode_displayed_tree_data[ann.network.root];
  if (tree >= rd.num_active_displayed_trees) throw std::runtime_error("persiteLoglikelihood: no such root displayed tree");
  DisplayedTreeData &treeAtRoot = rd.displayed_trees[tree];
  unsigned dtr = findFirstNodeWithTwoActiveChildren(ann, treeAtRoot.treeLoglData.reticulationChoices, ann.network.root);
  DisplayedTreeData &t = findMatchingDisplayedTree(ann, treeAtRoot.treeLoglData.reticulationChoices, ann.pernode_displayed_tree_data[dtr]);
  return ann.backend->rootLogl(p, t.clv_vector[p].p, t.scale_buffer[p].p, persite);
}

static void processPartitionsImproved(AnnotatedNetwork &ann, int incremental) {  // :488-519
  for (size_t i = 0; i < ann.travbuffer.size(); ++i) {
    unsigned n = ann.travbuffer[i];
    processNodeImproved(ann, incremental, n, ann.network.nodes[n].children, ConfigSet());
  }
  NodeDisplayedTreeData &rd = ann.pernode_displayed_tree_data[ann.network.root];
  for (size_t i = 0; i < rd.num_active_displayed_trees; ++i) computeDisplayedTreeLoglikelihood(ann, rd.displayed_trees[i], ann.network.root);
}

static double evaluateTreesPartition(AnnotatedNetwork &ann, size_t p, std::vector<TreeLoglData> &trees) {  // :521-604
  if (ann.options.likelihood_variant == LikelihoodVariant::AVERAGE_DISPLAYED_TREES) {
    XD partition_lh(0.0);
    for (TreeLoglData &t : trees) {
      if (t.tree_logprob < ann.options.min_interesting_tree_logprob) continue;
      if (!t.tree_logl_valid) throw std::runtime_error("invalid tree logl");
      partition_lh = partition_lh + xexp(t.tree_logprob) * xexp(t.tree_partition_logl[p]);
    }
    double pl = xlog(partition_lh);
    ann.partition_loglh[p] = pl;
    return pl;
  }
  double pl = -std::numeric_limits<double>::infinity();
  for (TreeLoglData &t : trees) {
    if (t.tree_logprob < ann.options.min_interesting_tree_logprob) continue;
    if (!t.tree_logl_valid) throw std::runtime_error("invalid tree logl");
    pl = std::max(pl, t.tree_logprob + t.tree_partition_logl[p]);
  }
  ann.partition_loglh[p] = pl;
  return pl;
}

static double evaluateTrees(AnnotatedNetwork &ann, unsigned virtual_root) {  // :606-644
  NodeDisplayedTreeData &nd = ann.pernode_displayed_tree_data[virtual_root];
  std::vector<TreeLoglData> tl;
  for (size_t i = 0; i < nd.num_active_displayed_trees; ++i) tl.emplace_back(nd.displayed_trees[i].treeLoglData);
  double network_logl = 0.0;
  for (unsigned p = 0; p < ann.partitionCount(); ++p) network_logl += evaluateTreesPartition(ann, p, tl);
  if (network_logl == -std::numeric_limits<double>::infinity()) throw std::runtime_error("Invalid network likelihood: negative infinity \n");
  ann.cached_logl = network_logl;
  ann.cached_logl_valid = true;
  return network_logl;
}

double computeLoglikelihood(AnnotatedNetwork &ann, int incremental, int update_pmatrices) {  // :646-671 + LikelihoodComputation.cpp:18-33
  if (ann.options.likelihood_variant == LikelihoodVariant::SARAH_PSEUDO) return computePseudoLoglikelihood(ann, incremental, update_pmatrices);
  if (!incremental) invalidateAllCLVs(ann);
  bool reuse = reuseOldDisplayedTreesCheck(ann, incremental, ann.network.root);
  if (reuse) {
    if (ann.cached_logl_valid) return ann.cached_logl;
  } else {
    if (update_pmatrices) updateProbMatrices(ann, !incremental);
    processPartitionsImproved(ann, incremental);
    if (!clvValidCheck(ann, ann.network.root)) throw std::runtime_error("Invalid displayed trees after loglikelihood computation");
  }
  return evaluateTrees(ann, ann.network.root);
}

/* ------------------------------------------------------------------------------------------
 * Naive cross-check (role of LH/NaiveLoglikelihood.cpp:35-115): every one of the 2^r displayed
 * trees is evaluated on its own by plain pruning over the active, alive part of the network; no
 * config sets, no CLV sharing.  Mixing as in the reference's naive path (:86-113).
 * ---------------------------------------------------------------------------------------- */
/* ---- LH/PseudoLoglikelihood.cpp ------------------------------------------------------------------------------------
 * One CLV per node: up to three libpll updates (both children / left only / right only) into scratch CLVs that SHARE the
 * node's scaler (the last executed update's scaler wins — kept as the reference has it), merged with the reticulation
 * probabilities of the children as weights; the fourth term is the all-ones fake CLV. */
static void merge_clvs(AnnotatedNetwork &ann, unsigned node, double w1, double w2, double w3, double w4) {  // :8-55
  for (unsigned p = 0; p < ann.partitionCount(); ++p) {
    double *clv = ann.pseudo_clv[node][p].p;
    const double *t1 = ann.tmp_clv_1[p].p, *t2 = ann.tmp_clv_2[p].p, *t3 = ann.tmp_clv_3[p].p;
    const size_t n = ann.backend->clvEntries(p);
    for (size_t i = 0; i < n; ++i) {
      double merged_entry = 0.0;
      if (w1 > 0.0) merged_entry += w1 * t1[i];
      if (w2 > 0.0) merged_entry += w2 * t2[i];
      if (w3 > 0.0) merged_entry += w3 * t3[i];
      if (w4 > 0.0) merged_entry += w4 * 1.0;   // partition->clv[fake_clv_index][i]
      clv[i] = merged_entry;
    }
  }
}

double computePseudoLoglikelihood(AnnotatedNetwork &ann, int incremental, int update_pmatrices) {  // :57-226
  Network &nw = ann.network;
  const unsigned P = ann.partitionCount();
  if (ann.pseudo_clv_valid.size() != nw.nodes.size()) {  // SRC/graph/AnnotatedNetwork.cpp:160-184
    ann.pseudo_clv_valid.assign(nw.nodes.size(), 0);
    for (unsigned i = 0; i < nw.num_tips; ++i) ann.pseudo_clv_valid[i] = 1;
    ann.pseudo_clv.clear(); ann.pseudo_scaler.clear();
    ann.pseudo_clv.resize(nw.nodes.size()); ann.pseudo_scaler.resize(nw.nodes.size());
    for (unsigned v = nw.num_tips; v < nw.nodes.size(); ++v) {
      ann.pseudo_clv[v].resize(P); ann.pseudo_scaler[v].resize(P);
      for (unsigned p = 0; p < P; ++p) { ann.pseudo_clv[v][p].alloc(ann.backend->clvEntries(p)); ann.pseudo_scaler[v][p].alloc(ann.backend->sites(p)); }
    }
    ann.tmp_clv_1.clear(); ann.tmp_clv_2.clear(); ann.tmp_clv_3.clear();
    ann.tmp_clv_1.resize(P); ann.tmp_clv_2.resize(P); ann.tmp_clv_3.resize(P);
    for (unsigned p = 0; p < P; ++p) { ann.tmp_clv_1[p].alloc(ann.backend->clvEntries(p)); ann.tmp_clv_2[p].alloc(ann.backend->clvEntries(p)); ann.tmp_clv_3[p].alloc(ann.backend->clvEntries(p)); }
  }
  if (update_pmatrices) updateProbMatrices(ann, !incremental);
  auto operandOf = [&](int child, unsigned p, unsigned parent) {
    Operand o;
    if (child < 0) { o.kind = 2; o.edge = ann.fakePmatrixIndex(); return o; }
    o.edge = nw.edgeBetween((unsigned)child, parent);
    if ((unsigned)child < nw.num_tips) { o.kind = 1; o.tip = (unsigned)child; }
    else { o.kind = 0; o.clv = ann.pseudo_clv[child][p].p; o.scaler = ann.pseudo_scaler[child][p].p; }
    return o;
  };
  auto parentProb = [&](int child, unsigned parent) {  // :95-118
    if (child < 0 || !nw.nodes[child].is_ret) return 1.0;
    const Network::Ret &R = nw.rets[nw.nodes[child].ret_index];
    return parent == R.first_parent ? ann.reticulation_probs[nw.nodes[child].ret_index] : 1.0 - ann.reticulation_probs[nw.nodes[child].ret_index];
  };
  for (unsigned node : ann.travbuffer) {
    if (node < nw.num_tips) continue;
    if (incremental && ann.pseudo_clv_valid[node]) continue;
    const std::vector<unsigned> &children = nw.nodes[node].children;
    if (children.empty() || children.size() > 2) throw std::runtime_error("computePseudoLoglikelihood: node with 0 or > 2 children");
    const int left = (int)children[0], right = children.size() == 1 ? -1 : (int)children[1];
    const double p_left = parentProb(left, node), p_right = parentProb(right, node);
    const double w1 = p_left * p_right, w2 = p_left * (1.0 - p_right), w3 = (1.0 - p_left) * p_right, w4 = (1.0 - p_left) * (1.0 - p_right);
    for (unsigned p = 0; p < P; ++p) {
      const Operand l = operandOf(left, p, node), r = operandOf(right, p, node), fake = operandOf(-1, p, node);
      unsigned *parent_scaler = ann.pseudo_scaler[node][p].p;
      if (w1 > 0.0) ann.backend->updatePartials(p, ann.tmp_clv_1[p].p, parent_scaler, l, r);      // case 1: take both
      if (w2 > 0.0) ann.backend->updatePartials(p, ann.tmp_clv_2[p].p, parent_scaler, l, fake);   // case 2: take left only
      if (w3 > 0.0) ann.backend->updatePartials(p, ann.tmp_clv_3[p].p, parent_scaler, fake, r);   // case 3: take right only
    }
    merge_clvs(ann, node, w1, w2, w3, w4);
    ann.pseudo_clv_valid[node] = 1;
  }
  std::vector<double> partition_pseudo_logl(P, 0.0);
  for (unsigned p = 0; p < P; ++p)
    partition_pseudo_logl[p] = ann.backend->rootLogl(p, ann.pseudo_clv[nw.root][p].p, ann.pseudo_scaler[nw.root][p].p, nullptr);
  if (ann.parallel_reduce_cb) ann.parallel_reduce_cb(ann.parallel_context, partition_pseudo_logl.data(), P, 0);
  double pseudo_logl = 0.0;
  for (unsigned p = 0; p < P; ++p) { pseudo_logl += partition_pseudo_logl[p]; ann.partition_loglh[p] = partition_pseudo_logl[p]; }
  return pseudo_logl;
}

namespace {
struct NaiveCtx {
  AnnotatedNetwork &ann;
  std::vector<bool> dead;
  std::vector<std::vector<ABuf<double>>> clv;     // [node][partition]
  std::vector<std::vector<ABuf<unsigned>>> scal;  // [node][partition]
};
void naiveUp(NaiveCtx &c, unsigned node) {
  Network &nw = c.ann.network;
  if (node < nw.num_tips) return;
  std::vector<unsigned> ch = nw.activeAliveChildren(c.dead, node);
  for (unsigned k : ch) naiveUp(c, k);
  const unsigned P = c.ann.partitionCount();
  c.clv[node].resize(P); c.scal[node].resize(P);
  for (unsigned p = 0; p < P; ++p) {
    c.clv[node][p].alloc(c.ann.backend->clvEntries(p));
    c.scal[node][p].alloc(c.ann.backend->sites(p));
    Operand o[2];
    for (int s = 0; s < 2; ++s) {
      if (s < (int)ch.size()) {
        unsigned k = ch[s];
        o[s].edge = nw.edgeBetween(k, node);
        if (k < nw.num_tips) { o[s].kind = 1; o[s].tip = k; }
        else { o[s].kind = 0; o[s].clv = c.clv[k][p].p; o[s].scaler = c.scal[k][p].p; }
      } else { o[s].kind = 2; o[s].edge = c.ann.fakePmatrixIndex(); }
    }
    c.ann.backend->updatePartials(p, c.clv[node][p].p, c.scal[node][p].p, o[0], o[1]);
  }
}
}  // namespace

double computeLoglikelihoodNaive(AnnotatedNetwork &ann, std::vector<double> *tree_logl, std::vector<double> *tree_logprob) {
  Network &nw = ann.network;
  const unsigned P = ann.partitionCount();
  updateProbMatrices(ann, 1);
  size_t n_trees = (size_t)1 << nw.num_reticulations();
  std::vector<std::vector<double>> tl(n_trees, std::vector<double>(P, 0.0));
  std::vector<double> lp(n_trees, 0.0);
  std::vector<unsigned char> saved = nw.toggle;
  for (size_t t = 0; t < n_trees; ++t) {
    for (size_t i = 0; i < nw.num_reticulations(); ++i) {
      nw.toggle[i] = (t >> i) & 1;
      lp[t] += nw.toggle[i] ? ann.second_parent_logprobs[i] : ann.first_parent_logprobs[i];
    }
    unsigned dtr = nw.root;
    NaiveCtx c{ann, nw.collectDeadNodes(nw.root, &dtr), {}, {}};
    c.clv.resize(nw.num_nodes()); c.scal.resize(nw.num_nodes());
    naiveUp(c, dtr);
    for (unsigned p = 0; p < P; ++p) tl[t][p] = ann.backend->rootLogl(p, c.clv[dtr][p].p, c.scal[dtr][p].p, nullptr);
  }
  nw.toggle = saved;
  double network_logl = 0;
  for (unsigned p = 0; p < P; ++p) {
    if (ann.options.likelihood_variant == LikelihoodVariant::AVERAGE_DISPLAYED_TREES) {
      XD s(0.0);
      for (size_t t = 0; t < n_trees; ++t) s = s + xexp(lp[t]) * xexp(tl[t][p]);
      network_logl += xlog(s);
    } else {
      double best = -std::numeric_limits<double>::infinity();
      for (size_t t = 0; t < n_trees; ++t) best = std::max(best, lp[t] + tl[t][p]);
      network_logl += best;
    }
  }
  if (tree_logl) { tree_logl->clear(); for (size_t t = 0; t < n_trees; ++t) for (unsigned p = 0; p < P; ++p) tree_logl->push_back(tl[t][p]); }
  if (tree_logprob) *tree_logprob = lp;
  return network_logl;
}

/* ------------------------------------------------------------------------------------------
 * LH/VirtualRerooting.cpp
 * ---------------------------------------------------------------------------------------- */
std::vector<DisplayedTreeData> extractOldTrees(AnnotatedNetwork &ann, unsigned vroot) {  // BranchLengthOptimization.cpp:34-53
  if (!clvValidCheck(ann, vroot)) throw std::runtime_error("Cannot reuse old displayed trees before the extractOldTrees step");
  std::vector<DisplayedTreeData> old;
  NodeDisplayedTreeData &nd = ann.pernode_displayed_tree_data[vroot];
  for (size_t i = 0; i < nd.num_active_displayed_trees; ++i) old.emplace_back(nd.displayed_trees[i]);
  return old;
}

namespace {
struct PathToVirtualRoot {  // :12-19
  ConfigSet reticulationChoices;
  std::vector<unsigned> path;
  std::vector<std::vector<unsigned>> children;
};

std::vector<unsigned> getParentPointers(AnnotatedNetwork &ann, unsigned vroot) {  // ParentHelper.cpp:62-86 (uses current toggles)
  const unsigned NONE = UINT_MAX;
  std::vector<unsigned> parent(ann.network.num_nodes(), NONE);
  parent[vroot] = vroot;
  std::queue<unsigned> q;
  q.push(vroot);
  while (!q.empty()) {
    unsigned a = q.front(); q.pop();
    for (unsigned nb : ann.network.activeNeighbors(a))
      if (parent[nb] == NONE) { q.push(nb); parent[nb] = a; }
  }
  parent[vroot] = NONE;
  return parent;
}

std::vector<unsigned> getChildrenIgnoreDirections(const Network &nw, unsigned node, unsigned myParent) {  // ChildrenHelper.cpp:23-35
  std::vector<unsigned> ch;
  for (unsigned nb : nw.nodes[node].neighbors()) if (nb != myParent) ch.push_back(nb);
  return ch;
}

std::vector<unsigned> getCurrentChildren(AnnotatedNetwork &ann, unsigned node, unsigned parent, const ConfigSet &restrictions) {  // ChildrenHelper.cpp:122-152
  std::vector<unsigned> res;
  for (unsigned c : getChildrenIgnoreDirections(ann.network, node, parent))
    if (reticulationConfigsCompatible(restrictions, getRestrictionsToTakeNeighbor(ann, node, c))) res.push_back(c);
  if (res.size() > 2) throw std::runtime_error("getCurrentChildren: more than two children");
  return res;
}

std::vector<PathToVirtualRoot> getPathsToVirtualRoot(AnnotatedNetwork &ann, unsigned old_vr, unsigned new_vr, unsigned new_vr_back) {  // :51-129
  std::vector<PathToVirtualRoot> res;
  NodeDisplayedTreeData &old = ann.pernode_displayed_tree_data[old_vr];
  for (size_t i = 0; i < old.num_active_displayed_trees; ++i) {
    PathToVirtualRoot ptvr;
    setReticulationParents(ann.network, old.displayed_trees[i].treeLoglData.reticulationChoices.configs[0]);
    std::vector<unsigned> parent = getParentPointers(ann, new_vr);
    std::vector<unsigned> path;  // getPathToVirtualRoot :21-33
    for (unsigned a = old_vr; a != new_vr; a = parent[a]) {
      if (a == UINT_MAX) throw std::runtime_error("virtual root not reachable");
      path.push_back(a);
    }
    path.push_back(new_vr);
    ConfigSet rs(ann.options.max_reticulations);
    rs.configs.emplace_back(Choices(ann.options.max_reticulations, RS::DONT_CARE));
    for (size_t j = 0; j + 1 < path.size(); ++j) rs = combineReticulationChoices(rs, getRestrictionsToTakeNeighbor(ann, path[j], path[j + 1]));
    ptvr.reticulationChoices = rs;
    ptvr.path = path;
    for (size_t j = 0; j + 1 < path.size(); ++j) {
      if (path[j] == new_vr_back) ptvr.children.emplace_back(getCurrentChildren(ann, path[j], new_vr, rs));
      else ptvr.children.emplace_back(getCurrentChildren(ann, path[j], parent[path[j]], rs));
    }
    ptvr.children.emplace_back(getCurrentChildren(ann, new_vr, new_vr_back, rs));
    res.emplace_back(ptvr);
  }
  bool dup = true;
  while (dup) {
    dup = false;
    for (size_t i = 0; i + 1 < res.size() && !dup; ++i)
      for (size_t j = i + 1; j < res.size(); ++j)
        if (res[i].path == res[j].path) { dup = true; std::swap(res[j], res[res.size() - 1]); res.pop_back(); break; }
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
    std::unordered_set<size_t> inPath;
    for (size_t i = 0; i < paths[p].path.size(); ++i) {
      inPath.emplace(paths[p].path[i]);
      for (unsigned c : paths[p].children[i]) info.pathNodesToRestore[p].emplace(c);
    }
    for (size_t n : inPath) info.pathNodesToRestore[p].erase(n);
    std::unordered_set<size_t> del;
    for (size_t m : info.pathNodesToRestore[p]) {
      bool save = false;
      for (size_t q = 0; q < p && !save; ++q)
        for (unsigned x : paths[q].path) if (x == m) { save = true; break; }
      if (!save) del.emplace(m);
    }
    for (size_t d : del) info.pathNodesToRestore[p].erase(d);
  }
  for (size_t p = 0; p < paths.size(); ++p) for (size_t n : info.pathNodesToRestore[p]) info.nodesInDanger.emplace(n);
  return info;
}
}  // namespace

void updateCLVsVirtualRerootTrees(AnnotatedNetwork &ann, unsigned old_vr, unsigned new_vr, unsigned new_vr_back, ConfigSet &restrictions) {  // :192-252
  std::vector<PathToVirtualRoot> paths = getPathsToVirtualRoot(ann, old_vr, new_vr, new_vr_back);
  NodeSaveInformation info = computeNodeSaveInformation(paths);
  std::vector<NodeDisplayedTreeData> buffered(ann.network.num_nodes());
  for (size_t n : info.nodesInDanger) buffered[n] = ann.pernode_displayed_tree_data[n];
  for (size_t p = 0; p < paths.size(); ++p) {
    if (!reticulationConfigsCompatible(paths[p].reticulationChoices, restrictions)) continue;
    for (size_t n : info.pathNodesToRestore[p]) ann.pernode_displayed_tree_data[n] = buffered[n];
    for (size_t i = 0; i < paths[p].path.size(); ++i) {
      bool appendMode = (p > 0) && (paths[p].path[i] == new_vr);
      processNodeImproved(ann, 0, paths[p].path[i], paths[p].children[i], paths[p].reticulationChoices, appendMode);
    }
  }
}

static void updateTreeData(AnnotatedNetwork &ann, const std::vector<DisplayedTreeData> &oldTrees, TreeLoglData &td) {  // :254-277
  const TreeLoglData &old = getMatchingTreeData(oldTrees, td.reticulationChoices);
  td.tree_partition_logl = old.tree_partition_logl;
  td.tree_logprob = computeReticulationConfigLogProb(td.reticulationChoices, ann.first_parent_logprobs, ann.second_parent_logprobs);
  td.tree_logprob_valid = true;
  td.tree_logl_valid = old.tree_logl_valid;
}

static void recomputeTreeData(AnnotatedNetwork &ann, size_t pmatrix_index, DisplayedTreeData &src, DisplayedTreeData &tgt, TreeLoglData &c) {  // :279-346
  c.tree_logprob = computeReticulationConfigLogProb(c.reticulationChoices, ann.first_parent_logprobs, ann.second_parent_logprobs);
  c.tree_logprob_valid = true;
  if (c.tree_logprob < ann.options.min_interesting_tree_logprob) return;
  for (unsigned p = 0; p < ann.partitionCount(); ++p) {
    Operand a = makeOperand(src, p, (unsigned)pmatrix_index), b = makeOperand(tgt, p, (unsigned)pmatrix_index);
    c.tree_partition_logl[p] = ann.backend->edgeLogl(p, a, b, (unsigned)pmatrix_index, nullptr);
  }
  if (ann.parallel_reduce_cb)  // LH/VirtualRerooting.cpp:326-344
    ann.parallel_reduce_cb(ann.parallel_context, c.tree_partition_logl.data(), ann.partitionCount(), 0);
  c.tree_logl_valid = true;
}

double computeLoglikelihoodBrlenOpt(AnnotatedNetwork &ann, const std::vector<DisplayedTreeData> &oldTrees, unsigned pmatrix_index, int update_pmatrices) {  // :348-585
  if (ann.cached_logl_valid) return ann.cached_logl;
  unsigned source = ann.network.edges[pmatrix_index].source, target = ann.network.edges[pmatrix_index].target;
  NodeDisplayedTreeData &sd = ann.pernode_displayed_tree_data[source];
  NodeDisplayedTreeData &td = ann.pernode_displayed_tree_data[target];
  size_t ns = sd.num_active_displayed_trees, nt = td.num_active_displayed_trees;
  std::vector<bool> sseen(ns, false), tseen(nt, false);
  if (!clvValidCheck(ann, ann.network.root, false)) throw std::runtime_error("Cannot reuse old displayed trees (root invalidated)");
  if (update_pmatrices) updateProbMatrices(ann, 0);
  std::vector<TreeLoglData> combined;
  const unsigned P = ann.partitionCount();
  for (size_t i = 0; i < ns; ++i)
    for (size_t j = 0; j < nt; ++j) {
      if (!reticulationConfigsCompatible(sd.displayed_trees[i].treeLoglData.reticulationChoices, td.displayed_trees[j].treeLoglData.reticulationChoices)) continue;
      TreeLoglData c(P, ann.options.max_reticulations);
      c.reticulationChoices = combineReticulationChoices(sd.displayed_trees[i].treeLoglData.reticulationChoices, td.displayed_trees[j].treeLoglData.reticulationChoices);
      if (isActiveAliveBranch(ann, c.reticulationChoices, pmatrix_index)) {
        recomputeTreeData(ann, pmatrix_index, sd.displayed_trees[i], td.displayed_trees[j], c);
        combined.emplace_back(c);
        sseen[i] = true; tseen[j] = true;
      }
    }
  for (size_t i = 0; i < ns; ++i)
    if (!sseen[i] && isActiveAliveBranch(ann, sd.displayed_trees[i].treeLoglData.reticulationChoices, pmatrix_index)) {
      updateTreeData(ann, oldTrees, sd.displayed_trees[i].treeLoglData);
      combined.emplace_back(sd.displayed_trees[i].treeLoglData);
    }
  for (size_t j = 0; j < nt; ++j)
    if (!tseen[j] && isActiveAliveBranch(ann, td.displayed_trees[j].treeLoglData.reticulationChoices, pmatrix_index)) {
      updateTreeData(ann, oldTrees, td.displayed_trees[j].treeLoglData);
      combined.emplace_back(td.displayed_trees[j].treeLoglData);
    }
  for (size_t i = 0; i < oldTrees.size(); ++i) {  // :471-502
    bool seen = false;
    for (size_t j = 0; j < combined.size(); ++j)
      if (reticulationConfigsCompatible(oldTrees[i].treeLoglData.reticulationChoices, combined[j].reticulationChoices)) { seen = true; break; }
    if (!seen) {
      TreeLoglData c(P, ann.options.max_reticulations);
      c.reticulationChoices = oldTrees[i].treeLoglData.reticulationChoices;
      updateTreeData(ann, oldTrees, c);
      combined.emplace_back(c);
    }
  }
  for (size_t t = 0; t < ((size_t)1 << ann.network.num_reticulations()); ++t) {  // :513-544
    ConfigSet tc = getTreeConfig(ann, t);
    bool seen = false;
    for (size_t i = 0; i < combined.size(); ++i) if (reticulationConfigsCompatible(tc, combined[i].reticulationChoices)) { seen = true; break; }
    if (!seen)
      for (size_t i = 0; i < oldTrees.size(); ++i)
        if (reticulationConfigsCompatible(tc, oldTrees[i].treeLoglData.reticulationChoices)) {
          TreeLoglData c(P, ann.options.max_reticulations);
          c.reticulationChoices = tc;
          updateTreeData(ann, oldTrees, c);
          combined.emplace_back(c);
          break;
        }
  }
  double network_logl = 0;
  for (unsigned p = 0; p < P; ++p) network_logl += evaluateTreesPartition(ann, p, combined);
  ann.cached_logl = network_logl;
  ann.cached_logl_valid = true;
  return network_logl;
}

/* ------------------------------------------------------------------------------------------
 * LH/LikelihoodDerivatives.cpp
 * ---------------------------------------------------------------------------------------- */
std::vector<std::vector<SumtableInfo>> computePartitionSumtables(AnnotatedNetwork &ann, unsigned pmatrix_index) {  // :291-344
  const unsigned P = ann.partitionCount();
  std::vector<std::vector<SumtableInfo>> res(P);
  unsigned source = ann.network.edges[pmatrix_index].source, target = ann.network.edges[pmatrix_index].target;
  NodeDisplayedTreeData &sd = ann.pernode_displayed_tree_data[source];
  NodeDisplayedTreeData &td = ann.pernode_displayed_tree_data[target];
  for (size_t i = 0; i < sd.num_active_displayed_trees; ++i)
    for (size_t j = 0; j < td.num_active_displayed_trees; ++j) {
      const ConfigSet &a = sd.displayed_trees[i].treeLoglData.reticulationChoices, &b = td.displayed_trees[j].treeLoglData.reticulationChoices;
      if (!reticulationConfigsCompatible(a, b)) continue;
      ConfigSet restrictions = combineReticulationChoices(a, b);
      if (!isActiveBranch(ann, restrictions, pmatrix_index)) continue;
      if (computeReticulationConfigLogProb(restrictions, ann.first_parent_logprobs, ann.second_parent_logprobs) < ann.options.min_interesting_tree_logprob) continue;
      for (unsigned p = 0; p < P; ++p) {  // computeSumtable :249-289
        SumtableInfo si;
        si.left_tree_idx = i; si.right_tree_idx = j;
        si.tree_prob = computeReticulationConfigProb(restrictions, ann.first_parent_logprobs, ann.second_parent_logprobs);
        si.sumtable.alloc(ann.backend->clvEntries(p));
        Operand l = makeOperand(sd.displayed_trees[i], p, pmatrix_index), r = makeOperand(td.displayed_trees[j], p, pmatrix_index);
        ann.backend->sumtable(p, l, r, si.sumtable.p);
        res[p].emplace_back(std::move(si));
      }
    }
  return res;
}

LoglDerivatives computeLoglikelihoodDerivatives(AnnotatedNetwork &ann, const std::vector<std::vector<SumtableInfo>> &sumtables, unsigned pmatrix_index) {  // :190-232 + computePartitionLhData :30-188
  const unsigned P = ann.partitionCount();
  LoglDerivatives out;
  out.logl_prime = 0.0; out.logl_prime_prime = 0.0;
  out.partition_logl_prime.assign(P, 0.0); out.partition_logl_prime_prime.assign(P, 0.0);
  out.raw.assign(P, {});
  if (ann.options.brlen_linkage == BRLEN_SCALED) throw std::runtime_error("I believe this function currently does not work correctly with scaled branch lengths");
  for (unsigned p = 0; p < P; ++p) {
    const std::vector<SumtableInfo> &st = sumtables[p];
    bool single_tree_mode = (st.size() == 1);
    double p_brlen = ann.branch_lengths[p][pmatrix_index];  // passed to libpll but unused there (diagptable is precomputed)
    (void)p_brlen;
    double branch_length = (ann.options.brlen_linkage == BRLEN_UNLINKED) ? ann.branch_lengths[p][pmatrix_index] : ann.linked_branch_lengths[pmatrix_index];
    XD lh_sum(0.0), lh_prime_sum(0.0), lh_prime_prime_sum(0.0);
    double best_score = -std::numeric_limits<double>::infinity(), best_prime = best_score, best_prime_prime = best_score;
    double res_prime = 0.0, res_prime_prime = 0.0;
    bool done = false;
    for (size_t i = 0; i < st.size(); ++i) {
      double f = 0.0, d1 = 0.0, d2 = 0.0;
      ann.backend->derivatives(p, st[i].sumtable.p, branch_length, !single_tree_mode, &f, &d1, &d2);
      if (ann.parallel_reduce_cb) {  // LH/LikelihoodDerivatives.cpp:108-143 (per-partition values of a linked run are reduced one at a time here)
        double v[3] = {f, d1, d2};
        ann.parallel_reduce_cb(ann.parallel_context, v, 3, 0);
        f = v[0]; d1 = v[1]; d2 = v[2];
      }
      out.raw[p].push_back(f); out.raw[p].push_back(d1); out.raw[p].push_back(d2);
      if (single_tree_mode) { res_prime = d1; res_prime_prime = d2; done = true; break; }
      if (ann.options.likelihood_variant == LikelihoodVariant::AVERAGE_DISPLAYED_TREES) {
        XD lh = xexp(f);                      // computeTreeDerivatives :13-23
        XD lhp = lh * XD(d1);
        XD lhpp = lhp * XD(d1) + lh * XD(d2);
        lh_sum = lh_sum + lh * XD(st[i].tree_prob);
        lh_prime_sum = lh_prime_sum + lhp * XD(st[i].tree_prob);
        lh_prime_prime_sum = lh_prime_prime_sum + lhpp * XD(st[i].tree_prob);
      } else if (f * st[i].tree_prob > best_score) {  // Q2
        best_score = f * st[i].tree_prob; best_prime = d1; best_prime_prime = d2;
      }
    }
    if (!done) {
      if (ann.options.likelihood_variant == LikelihoodVariant::AVERAGE_DISPLAYED_TREES) {
        res_prime = (lh_prime_sum / lh_sum).toDouble();
        res_prime_prime = ((lh_prime_prime_sum * lh_sum - lh_prime_sum * lh_prime_sum) / (lh_sum * lh_sum)).toDouble();
      } else { res_prime = best_prime; res_prime_prime = best_prime_prime; }
    }
    out.partition_logl_prime[p] = res_prime;
    out.partition_logl_prime_prime[p] = res_prime_prime;
    out.logl_prime += res_prime;
    out.logl_prime_prime += res_prime_prime;
  }
  return out;
}

}  // namespace orc
