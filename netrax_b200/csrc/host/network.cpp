/*
 * network.cpp — flat rooted network + the few graph helpers the likelihood layer needs.  The reference keeps a
 * pointer/link structure (src/graph/Network.hpp, src/helper/*.cpp); topology surgery is out of scope here, so
 * adjacency is stored as index lists.  Neighbour order = parents first, then children by ascending pmatrix index.
 */
#include <algorithm>
#include <queue>

#include "host_internal.hpp"

namespace netrax {

Network buildNetwork(size_t num_tips, size_t num_nodes, size_t root, const std::vector<Edge> &edges,
                     const std::vector<size_t> &ret_node, const std::vector<size_t> &ret_first_edge,
                     const std::vector<size_t> &ret_second_edge) {
  Network nw;
  nw.tipCount = num_tips;
  nw.nodes.resize(num_nodes);
  nw.edges = edges;
  for (size_t i = 0; i < num_nodes; ++i) nw.nodes[i].clv_index = i;
  for (size_t e = 0; e < edges.size(); ++e) {
    nw.edges[e].pmatrix_index = e;
    if (edges[e].source >= num_nodes || edges[e].target >= num_nodes) throw std::runtime_error("edge endpoint out of range");
  }
  for (size_t r = 0; r < ret_node.size(); ++r) {
    ReticulationInfo R{ret_node[r], 0, 0, SIZE_MAX, ret_first_edge[r], ret_second_edge[r]};
    if (R.first_edge >= edges.size() || R.second_edge >= edges.size() || R.node >= num_nodes) throw std::runtime_error("bad reticulation record");
    R.first_parent = edges[R.first_edge].source;
    R.second_parent = edges[R.second_edge].source;
    if (edges[R.first_edge].target != R.node || edges[R.second_edge].target != R.node) throw std::runtime_error("reticulation edges do not end in the reticulation node");
    if (R.first_parent == R.second_parent) throw std::runtime_error("parallel reticulation arcs are not supported");
    Node &n = nw.nodes[R.node];
    n.type = NodeType::RETICULATION_NODE;
    n.reticulation_index = r;
    n.parents = {R.first_parent, R.second_parent};
    nw.reticulations.push_back(R);
  }
  if (nw.reticulations.size() > 32) throw std::runtime_error("at most 32 reticulations are supported (max_reticulations)");
  for (const Edge &E : nw.edges) {
    nw.nodes[E.source].children.push_back(E.target);
    Node &t = nw.nodes[E.target];
    if (t.type != NodeType::RETICULATION_NODE) {
      if (!t.parents.empty()) throw std::runtime_error("non-reticulation node with two parents");
      t.parents.push_back(E.source);
    }
  }
  for (ReticulationInfo &R : nw.reticulations) {
    if (nw.nodes[R.node].children.size() != 1) throw std::runtime_error("Found a reticulation node that has no children");
    R.child = nw.nodes[R.node].children[0];
  }
  for (Node &n : nw.nodes) {
    if (n.children.size() > 2) throw std::runtime_error("The network is not bifurcating");
    for (size_t p : n.parents) if (std::find(n.neighbors.begin(), n.neighbors.end(), p) == n.neighbors.end()) n.neighbors.push_back(p);
    for (size_t c : n.children) if (std::find(n.neighbors.begin(), n.neighbors.end(), c) == n.neighbors.end()) n.neighbors.push_back(c);
  }
  nw.nodes_by_index.resize(num_nodes);
  for (size_t i = 0; i < num_nodes; ++i) nw.nodes_by_index[i] = &nw.nodes[i];
  nw.edges_by_index.resize(nw.edges.size());
  for (size_t e = 0; e < nw.edges.size(); ++e) nw.edges_by_index[e] = &nw.edges[e];
  for (const ReticulationInfo &R : nw.reticulations) nw.reticulation_nodes.push_back(&nw.nodes[R.node]);
  nw.active_parent_toggle.assign(nw.reticulations.size(), 0);
  nw.root = &nw.nodes[root];
  return nw;
}

std::vector<Node *> reversed_topological_sort(Network &nw) {  // Kahn on out-degrees, NetworkFunctions.cpp:542-598
  std::vector<Node *> res;
  std::vector<size_t> outdeg(nw.num_nodes());
  std::queue<size_t> q;
  for (size_t i = 0; i < nw.num_nodes(); ++i) { outdeg[i] = nw.nodes[i].children.size(); if (!outdeg[i]) q.push(i); }
  while (!q.empty()) {
    size_t a = q.front(); q.pop();
    res.push_back(&nw.nodes[a]);
    for (size_t p : nw.nodes[a].parents) if (--outdeg[p] == 0) q.push(p);
  }
  if (res.size() != nw.num_nodes()) throw std::runtime_error("Cycle in network detected");
  return res;
}

namespace detail {

size_t edgeBetween(const Network &nw, size_t a, size_t b) {  // getEdgeTo
  for (const Edge &E : nw.edges)
    if ((E.source == a && E.target == b) || (E.source == b && E.target == a)) return E.pmatrix_index;
  throw std::runtime_error("no edge between the two nodes");
}

size_t activeParent(const Network &nw, size_t n) {  // ParentHelper.cpp:5-16
  const Node &N = nw.nodes[n];
  if (N.type == NodeType::RETICULATION_NODE) {
    const ReticulationInfo &R = nw.reticulations[N.reticulation_index];
    return nw.active_parent_toggle[N.reticulation_index] ? R.second_parent : R.first_parent;
  }
  return N.parents.empty() ? SIZE_MAX : N.parents[0];
}

std::vector<size_t> activeAliveChildren(const Network &nw, const std::vector<char> &dead, size_t n) {  // ChildrenHelper.cpp:58-80
  std::vector<size_t> res;
  for (size_t c : nw.nodes[n].children) {
    if (dead[c]) continue;
    if (nw.nodes[c].type == NodeType::RETICULATION_NODE && activeParent(nw, c) != n) continue;
    res.push_back(c);
  }
  return res;
}

std::vector<size_t> activeNeighbors(const Network &nw, size_t n) {  // NeighborHelper.cpp:19-43
  std::vector<size_t> res;
  const Node &N = nw.nodes[n];
  for (size_t nb : N.neighbors) {
    const Node &B = nw.nodes[nb];
    if (B.type == NodeType::RETICULATION_NODE && n != nw.reticulations[B.reticulation_index].child && activeParent(nw, nb) != n) continue;
    if (N.type == NodeType::RETICULATION_NODE && nb != nw.reticulations[N.reticulation_index].child && nb != activeParent(nw, n)) continue;
    res.push_back(nb);
  }
  return res;
}

std::vector<char> collect_dead_nodes(const Network &nw, size_t megablobRoot, size_t *displayed_tree_root) {  // NetworkFunctions.cpp:228-279
  std::vector<char> dead(nw.num_nodes(), 0);
  std::queue<size_t> q;
  for (size_t i = 0; i < nw.num_reticulations(); ++i)
    q.push(nw.active_parent_toggle[i] ? nw.reticulations[i].first_parent : nw.reticulations[i].second_parent);
  while (!q.empty()) {
    size_t u = q.front(); q.pop();
    if (!activeAliveChildren(nw, dead, u).empty()) continue;
    dead[u] = 1;
    if (nw.nodes[u].type == NodeType::RETICULATION_NODE) {
      q.push(nw.reticulations[nw.nodes[u].reticulation_index].first_parent);
      q.push(nw.reticulations[nw.nodes[u].reticulation_index].second_parent);
    } else {
      size_t p = activeParent(nw, u);
      if (p != SIZE_MAX) q.push(p);
    }
  }
  size_t dtroot = nw.root->clv_index;
  std::vector<size_t> ch = activeAliveChildren(nw, dead, dtroot);
  bool seen = false;
  while (ch.size() == 1) {
    if (dtroot == megablobRoot) seen = true;
    dead[dtroot] = 1;
    dtroot = ch[0];
    ch = activeAliveChildren(nw, dead, dtroot);
  }
  if (displayed_tree_root) *displayed_tree_root = seen ? dtroot : megablobRoot;
  return dead;
}

void setReticulationParents(Network &nw, const ReticulationConfig &c) {  // ReticulationHelper.cpp:142-157 (DONT_CARE keeps the old toggle)
  for (size_t i = 0; i < nw.num_reticulations(); ++i)
    if ((c.care >> i) & 1) nw.active_parent_toggle[i] = (c.second >> i) & 1;
}

}  // namespace detail
}  // namespace netrax
