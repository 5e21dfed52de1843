/*
 * upper_seam_caller.cpp — a CALLER of the upper seam written the way NetRAX's own code and tests call it
 * (namespace netrax free functions on AnnotatedNetwork&; reference: test/src/LikelihoodTest.cpp:204-279,
 * test/src/BrlenOptTest.cpp:95-120,297-367, src/optimization/BranchLengthOptimization.cpp:345-420).  It includes ONLY
 * netrax_likelihood_api.hpp and links libnetrax_b200.so: what "the search, moves and optimisation layers call it
 * unchanged" means at the C++ level.  Input: a text file written by tests/test_cpp_upper_seam.py; output: "key value" lines.
 */
#include <cmath>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <stdexcept>
#include <vector>

#include "../../netrax_b200/csrc/host/netrax_likelihood_api.hpp"

using namespace netrax;

int main(int argc, char **argv) {
  if (argc < 2) { std::fprintf(stderr, "usage: upper_seam_caller <input.txt>\n"); return 2; }
  try {
    std::ifstream in(argv[1]);
    size_t num_tips, num_nodes, root, num_edges, num_ret, sites;
    in >> num_tips >> num_nodes >> root >> num_edges >> num_ret >> sites;
    std::vector<Edge> edges(num_edges);
    for (Edge &e : edges) in >> e.source >> e.target >> e.length >> e.prob;
    std::vector<size_t> ret_node(num_ret), ret_first(num_ret), ret_second(num_ret);
    for (size_t r = 0; r < num_ret; ++r) in >> ret_node[r] >> ret_first[r] >> ret_second[r];
    std::vector<uint32_t> masks(num_tips * sites), weights(sites);
    for (uint32_t &m : masks) in >> m;
    for (uint32_t &w : weights) in >> w;
    PartitionInput part;
    part.model.states = 4; part.model.rate_cats = 4; part.model.sites = (unsigned)sites;
    part.model.frequencies.resize(4); part.model.subst_params.resize(6); part.model.rates.resize(4);
    for (double &v : part.model.frequencies) in >> v;
    set_frequencies(part.model, std::vector<double>(part.model.frequencies).data());   // pll_set_frequencies
    for (double &v : part.model.subst_params) in >> v;
    for (double &v : part.model.rates) in >> v;
    if (!in) throw std::runtime_error("short input file");
    part.tip_masks = masks.data();
    part.pattern_weights = weights.data();

    AnnotatedNetwork ann_network;
    ann_network.network = buildNetwork(num_tips, num_nodes, root, edges, ret_node, ret_first, ret_second);
    ann_network.network.root = &ann_network.network.nodes[root];
    init_annotated_network(ann_network, {part}, 0);

    // LikelihoodTest.cpp:204-279: non-incremental == incremental
    const double full = computeLoglikelihood(ann_network, 0, 1);
    const double incremental = computeLoglikelihood(ann_network, 1, 1);
    std::printf("logl_full %.17g\nlogl_incremental %.17g\n", full, incremental);

    // BrlenOptTest.cpp:297-367: virtual re-rooting to every edge preserves the network lnL
    double worst = 0.0;
    for (size_t pmatrix_index = 0; pmatrix_index < ann_network.network.num_branches(); ++pmatrix_index) {
      std::vector<DisplayedTreeData> oldTrees = extractOldTrees(ann_network, ann_network.network.root);
      Node *new_virtual_root = &ann_network.network.nodes[ann_network.network.edges[pmatrix_index].source];
      Node *new_virtual_root_back = &ann_network.network.nodes[ann_network.network.edges[pmatrix_index].target];
      ReticulationConfigSet restrictions = getRestrictionsActiveAliveBranch(ann_network, pmatrix_index);
      updateCLVsVirtualRerootTrees(ann_network, ann_network.network.root, new_virtual_root, new_virtual_root_back, restrictions);
      ann_network.cached_logl_valid = false;
      const double brlenopt_logl = computeLoglikelihoodBrlenOpt(ann_network, oldTrees, (unsigned)pmatrix_index);
      worst = std::fmax(worst, std::fabs(brlenopt_logl - full));
      invalidatePmatrixIndex(ann_network, pmatrix_index);   // restore the network root (BranchLengthOptimization.cpp:413-418)
      computeLoglikelihood(ann_network);
    }
    std::printf("reroot_max_abs_diff %.3g\n", worst);

    // derivatives on edge 0 (LikelihoodDerivatives.hpp:82-87)
    {
      std::vector<DisplayedTreeData> oldTrees = extractOldTrees(ann_network, ann_network.network.root);
      ReticulationConfigSet restrictions = getRestrictionsActiveAliveBranch(ann_network, 0);
      updateCLVsVirtualRerootTrees(ann_network, ann_network.network.root, &ann_network.network.nodes[ann_network.network.edges[0].source],
                                   &ann_network.network.nodes[ann_network.network.edges[0].target], restrictions);
      ann_network.cached_logl_valid = false;
      computeLoglikelihoodBrlenOpt(ann_network, oldTrees, 0);
      std::vector<std::vector<SumtableInfo>> sumtables = computePartitionSumtables(ann_network, 0);
      LoglDerivatives d = computeLoglikelihoodDerivatives(ann_network, sumtables, 0);
      std::printf("edge0_logl_prime %.17g\nedge0_logl_prime_prime %.17g\n", d.logl_prime, d.logl_prime_prime);
      invalidatePmatrixIndex(ann_network, 0);
      computeLoglikelihood(ann_network);
    }

    // BrlenOptTest.cpp:95-120: ASSERT_GE(new_logl, old_logl); then the scores the search compares
    const double bic_before = scoreNetwork(ann_network);
    const double after_brlen = optimize_branches(ann_network, 32, 32, -1);
    const double after_probs = optimize_reticulations(ann_network, 10);
    std::printf("logl_after_brlen %.17g\nlogl_after_probs %.17g\nbic_before %.17g\nbic_after %.17g\n", after_brlen, after_probs, bic_before,
                scoreNetwork(ann_network));
    NetworkParams params(&ann_network);   // the slot pll-modules' optimisers re-enter through
    std::printf("likelihood_target_function %.17g\n", network_logl_wrapper(&params, 0, 1, nullptr));
    return 0;
  } catch (const std::exception &e) {
    std::fprintf(stderr, "ERROR: %s\n", e.what());
    return 1;
  }
}
