/*
 * netrax_likelihood_api.hpp — the UPPER SEAM of the drop-in: NetRAX's likelihood API with the
 * reference's own names, argument meaning and error behaviour (std::runtime_error), re-implemented on top of
 * the B200 device engine (include/nrx_engine.h).  Callers in src/optimization, src/search and pll-modules'
 * optimisers (through likelihood_target_function) use exactly these entry points:
 *
 *   computeLoglikelihood            src/likelihood/LikelihoodComputation.hpp:17-18
 *   computeLoglikelihoodBrlenOpt    src/likelihood/VirtualRerooting.hpp:8-12
 *   updateCLVsVirtualRerootTrees    src/likelihood/VirtualRerooting.hpp:13-16
 *   computePartitionSumtables       src/likelihood/LikelihoodDerivatives.hpp:86-87
 *   computeLoglikelihoodDerivatives src/likelihood/LikelihoodDerivatives.hpp:82-85
 *   evaluateTreesPartition          src/likelihood/ImprovedLoglikelihood.hpp:10-14
 *   invalidate* / allClvsValid      src/helper/Helper.hpp (InvalidationHelper.cpp)
 *
 * What differs from the reference, by design:
 *   - a displayed tree's CLV / scaler live in a device SLOT (all partitions), not in host vectors;
 *   - the (child-tree, child-tree) enumeration of processNodeImproved emits nrx_op records and ONE launch
 *     covers all displayed trees of a node (independent nodes are batched into the same launch);
 *   - per-tree root / edge lnLs of ALL trees are reduced in one kernel + one parallel_reduce_cb call instead
 *     of one call per tree;
 *   - reticulation configurations are (care, value) bit masks instead of enum vectors;
 *   - mpfr::mpreal (53 bit) is replaced by max-shifted double arithmetic.
 * Topology surgery (moves), I/O, search and the optimisers themselves are callers and out of scope.
 */
#pragma once

#include <cstddef>
#include <cstdint>
#include <limits>
#include <stdexcept>
#include <string>
#include <unordered_set>
#include <vector>

#include "../../../include/nrx_engine.h"

namespace netrax {

/* ---- src/graph/ReticulationConfigSet.hpp --------------------------------------------------------- */
enum class ReticulationState { DONT_CARE = 0, TAKE_FIRST_PARENT = 1, TAKE_SECOND_PARENT = 2, INVALID = 3 };

struct ReticulationConfig {  // one vector<ReticulationState> of the reference, as bit masks
  uint32_t care = 0;         // bit i set: reticulation i is fixed
  uint32_t second = 0;       // bit i set (only where care): TAKE_SECOND_PARENT, else TAKE_FIRST_PARENT
  bool operator==(const ReticulationConfig &o) const { return care == o.care && second == o.second; }
  ReticulationState operator[](size_t i) const {
    return !((care >> i) & 1) ? ReticulationState::DONT_CARE
                              : (((second >> i) & 1) ? ReticulationState::TAKE_SECOND_PARENT : ReticulationState::TAKE_FIRST_PARENT);
  }
};

struct ReticulationConfigSet {  // a big OR of configurations
  std::vector<ReticulationConfig> configs;
  size_t max_reticulations = 0;
  ReticulationConfigSet() = default;
  explicit ReticulationConfigSet(size_t m) : max_reticulations(m) {}
  bool empty() const { return configs.empty(); }
  bool operator==(const ReticulationConfigSet &o) const;
};

double computeReticulationConfigProb(const ReticulationConfigSet &c, const std::vector<double> &first, const std::vector<double> &second);
double computeReticulationConfigLogProb(const ReticulationConfigSet &c, const std::vector<double> &first, const std::vector<double> &second);
bool reticulationConfigsCompatible(const ReticulationConfigSet &l, const ReticulationConfigSet &r);
ReticulationConfigSet combineReticulationChoices(const ReticulationConfigSet &l, const ReticulationConfigSet &r);
void simplifyReticulationChoices(ReticulationConfigSet &res);
std::string toString(const ReticulationConfigSet &c, size_t n_reticulations);

/* ---- src/graph/{Node,Edge,Network}.hpp: the fields the likelihood layer reads ------------------------ */
enum class NodeType { BASIC_NODE = 0, RETICULATION_NODE = 1 };

struct Node {
  size_t clv_index = 0;
  NodeType type = NodeType::BASIC_NODE;
  size_t reticulation_index = 0;          // getReticulationData()->reticulation_index
  std::vector<size_t> parents;            // clv indices; reticulation: {first, second}
  std::vector<size_t> children;           // by ascending pmatrix index
  std::vector<size_t> neighbors;          // parents first, then children (the reference's link order)
  NodeType getType() const { return type; }
};

struct Edge {
  size_t pmatrix_index = 0;
  size_t source = 0, target = 0;  // getSource / getTarget
  double length = 0.0, prob = 1.0;
};

struct ReticulationInfo { size_t node, first_parent, second_parent, child, first_edge, second_edge; };

struct Network {
  std::vector<Node> nodes;
  std::vector<Edge> edges;
  std::vector<Node *> nodes_by_index;
  std::vector<Edge *> edges_by_index;
  std::vector<Node *> reticulation_nodes;
  std::vector<ReticulationInfo> reticulations;
  std::vector<unsigned char> active_parent_toggle;  // ReticulationData::active_parent_toggle (stateful)
  Node *root = nullptr;
  size_t tipCount = 0;
  size_t num_tips() const { return tipCount; }
  size_t num_nodes() const { return nodes.size(); }
  size_t num_branches() const { return edges.size(); }
  size_t num_reticulations() const { return reticulations.size(); }
};

Network buildNetwork(size_t num_tips, size_t num_nodes, size_t root, const std::vector<Edge> &edges,
                     const std::vector<size_t> &ret_node, const std::vector<size_t> &ret_first_edge,
                     const std::vector<size_t> &ret_second_edge);
std::vector<Node *> reversed_topological_sort(Network &network);  // src/helper/NetworkFunctions.cpp:542-598

/* ---- src/likelihood/LikelihoodVariant.hpp, src/NetraxOptions.hpp ---------------------------------------- */
enum class LikelihoodVariant { AVERAGE_DISPLAYED_TREES = 0, BEST_DISPLAYED_TREE = 1, SARAH_PSEUDO = 2 };
enum { PLLMOD_COMMON_BRLEN_LINKED = 0, PLLMOD_COMMON_BRLEN_SCALED = 1, PLLMOD_COMMON_BRLEN_UNLINKED = 2 };
enum { PLLMOD_COMMON_REDUCE_SUM = 0, PLLMOD_COMMON_REDUCE_MAX = 1, PLLMOD_COMMON_REDUCE_MIN = 2 };
enum class BrlenOptMethod { BRENT_NORMAL = 0, BRENT_REROOT = 1, NEWTON_RAPHSON = 2 };  // src/NetraxOptions.hpp:17-21

struct NetraxOptions {
  LikelihoodVariant likelihood_variant = LikelihoodVariant::AVERAGE_DISPLAYED_TREES;
  int brlen_linkage = PLLMOD_COMMON_BRLEN_LINKED;
  size_t max_reticulations = 32;
  double min_interesting_tree_logprob = -13.815510557964274;  // log(1e-6), NetraxOptions.hpp:108
  double brlen_min = 1e-6, brlen_max = 100.0;               // RAXML_BRLEN_MIN / MAX (RAXML/constants.hpp:14-15)
  double brprob_min = 1e-6, brprob_max = 1.0 - 1e-6;        // NetraxOptions.hpp:102-103
  double lh_epsilon = 0.1, tolerance = 0.1;                 // DEF_LH_EPSILON (NetraxOptions.hpp:104-105)
  BrlenOptMethod brlenOptMethod = BrlenOptMethod::NEWTON_RAPHSON;  // NetraxOptions.hpp:124-125
  bool save_memory = false;
};

/* ---- per-partition model: the fields of pll_partition_t the path reads -------------------------------- */
struct PartitionModel {
  unsigned states = 4, states_padded = 4, rate_cats = 4, sites = 0;
  std::vector<double> frequencies, subst_params, rates, rate_weights;
  double prop_invar = 0.0;  // pll_partition_t::prop_invar[0] (+I); the invariant-site indices are derived from the tips by the engine
  double alpha = 0.0;   // pllmod_treeinfo_t::alphas[p]; > 0: `rates` are the discrete-Gamma rates of this shape and optimize_alpha may change it
  int gamma_mode = 0;   // PLL_GAMMA_RATES_MEAN (raxml-ng default)
  /* pllmod_treeinfo_t::params_to_optimize[p] for the 1-D optimisers (bit 0: alpha, bit 1: pinv); -1 = not set: derived from the
   * current values (alpha > 0 / prop_invar > 0) as before.  Set from the model spec so that a +I partition whose proportion
   * is (or was optimised to) 0 stays a free parameter. */
  int params_to_optimize = -1;
  std::vector<double> eigenvecs, inv_eigenvecs, eigenvals;  // filled by update_eigen (own Jacobi solver) or set explicitly
  bool eigen_decomp_valid = false;
  /* Mixtures with one rate matrix per rate category (LG4M / LG4X): raxml-ng's Model::ratecat_submodels(), which NetRAX hands
   * to pll-modules as param_indices and from there to every libpll call (src/RaxmlWrapper.cpp:199-203,
   * LH/ImprovedLoglikelihood.cpp:452).  Rate matrix 0 is (frequencies, subst_params) above; matrices 1.. are `submodels`;
   * category c uses matrix ratecat_submodels[c] (empty = all 0, the single-matrix case of every BASELINE config). */
  struct SubModel {
    std::vector<double> frequencies, subst_params, eigenvecs, inv_eigenvecs, eigenvals;
  };
  std::vector<SubModel> submodels;
  std::vector<unsigned> ratecat_submodels;
};
/* installs rate matrices 1..n-1 (matrix 0 = freqs[0] / subst[0] too) and the category -> matrix map; n == 1 removes the mixture */
struct AnnotatedNetwork;
/* fake_treeinfo->brlen_scalers[p] = scaler (scaled linkage): the P-matrices of partition p use scaler x linked branch length
 * (pllmod_treeinfo_update_prob_matrices, PLLMOD/tree/treeinfo.c:862-864); what pll-modules' scaler optimiser
 * (pllmod_algo_opt_brlen_scalers_treeinfo, called by optimize_scalers, src/optimization/BranchLengthOptimization.cpp:581-599)
 * changes before it re-enters through likelihood_target_function */
void set_brlen_scaler(AnnotatedNetwork &ann, unsigned partition, double scaler);
void set_submodels(PartitionModel &m, unsigned n, const unsigned *ratecat_submodels, const double *freqs /*[n][states]*/,
                   const double *subst /*[n][states (states - 1) / 2]*/);
void set_frequencies(PartitionModel &m, const double *frequencies);                    // pll_set_frequencies (LIBPLL/models.c:445-467): renormalises when |sum - 1| > 1e-8
void update_eigen(PartitionModel &m);                                                 // role of pll_update_eigen
bool compute_gamma_cats(double alpha, unsigned cats, double *out_rates, int mode);    // role of pll_compute_gamma_cats

/* ---- the "fake treeinfo": the fields of pllmod_treeinfo_t (PLLMOD/tree/pll_tree.h:208-279) in use ------ */
struct FakeTreeinfo {
  unsigned partition_count = 0;
  std::vector<PartitionModel> partitions;
  std::vector<std::vector<double>> branch_lengths;   // [partition][edge + 1 fake]
  std::vector<double> linked_branch_lengths;         // [edge + 1 fake]
  std::vector<double> brlen_scalers;                 // [partition], PLLMOD_COMMON_BRLEN_SCALED only (PLLMOD/tree/pll_tree.h:236); empty = all 1
  int brlen_linkage = PLLMOD_COMMON_BRLEN_LINKED;
  std::vector<double> partition_loglh;
  std::vector<std::vector<char>> clv_valid;          // [partition][node]
  std::vector<std::vector<char>> pmatrix_valid;      // [partition][edge + 1 fake]
  /* pllmod's reduction hook (PLLMOD/tree/pll_tree.h:262-266): SUM over all site shards (ranks). */
  void (*parallel_reduce_cb)(void *context, double *data, size_t count, int op) = nullptr;
  void *parallel_context = nullptr;
};

/* ---- src/graph/{TreeLoglData,DisplayedTreeData,NodeDisplayedTreeData}.hpp ------------------------------ */
struct TreeLoglData {
  ReticulationConfigSet reticulationChoices;
  double tree_logprob = 0.0;
  bool tree_logprob_valid = false;
  std::vector<double> tree_partition_logl;
  bool tree_logl_valid = false;
  TreeLoglData() = default;
  TreeLoglData(size_t n_partitions, size_t max_reticulations) : reticulationChoices(max_reticulations), tree_partition_logl(n_partitions, 0.0) {}
};

struct DisplayedTreeData {
  TreeLoglData treeLoglData;
  uint32_t slot = UINT32_MAX;  // device CLV + scaler slot (replaces clv_vector / scale_buffer host pointers)
  bool clv_valid = false;
  bool isTip = false;
  uint32_t tip = 0;
};

struct NodeDisplayedTreeData {
  std::vector<DisplayedTreeData> displayed_trees;
  size_t num_active_displayed_trees = 0;
};

struct SumtableInfo {  // src/likelihood/LikelihoodDerivatives.hpp:9-74; the table itself is device sumtable slot `index`
  double tree_prob = 0.0;
  uint32_t index = 0;
  size_t left_tree_idx = 0, right_tree_idx = 0;
};

struct LoglDerivatives {
  double logl_prime = std::numeric_limits<double>::infinity();
  double logl_prime_prime = std::numeric_limits<double>::infinity();
  std::vector<double> partition_logl_prime, partition_logl_prime_prime;
  std::vector<std::vector<double>> raw;  // [partition][3 * sumtable]: (f, d1, d2) per displayed-tree pair
};

struct PlanCache;  // cached evaluation plan (ops per launch) for unchanged topology
struct RerootCache;  // virtual re-rooting session + memoised re-rooted node data (host/brlen.cpp)

struct AnnotatedNetwork {  // src/graph/AnnotatedNetwork.hpp:42-89
  NetraxOptions options;
  Network network;
  FakeTreeinfo fake_treeinfo_storage;
  FakeTreeinfo *fake_treeinfo = &fake_treeinfo_storage;
  std::vector<double> reticulation_probs, first_parent_logprobs, second_parent_logprobs;
  std::vector<NodeDisplayedTreeData> pernode_displayed_tree_data;
  std::vector<Node *> travbuffer;
  double cached_logl = 0.0;
  bool cached_logl_valid = false;
  size_t total_num_model_parameters = 0;  // parted_msa->total_free_model_params() (src/RaxmlWrapper.cpp:682-684); caller-set
  size_t total_num_sites = 0;             // parted_msa->total_sites(); default: sum of this process's pattern weights
  std::vector<double> pattern_weight_sums;   // [partition] pll_partition_t::pattern_weight_sum of this site shard
  double pattern_weight_sum(unsigned p) const { return pattern_weight_sums.at(p); }
  /* optimize_params (src/optimization/ModelOptimization.cpp:25-97): pll-modules' model optimisers in a NetRAX build; when
   * unset, optimizeModel runs the steps this repo implements itself: optimize_alpha, then optimize_pinv */
  double (*optimize_params_cb)(AnnotatedNetwork &) = nullptr;

  /* device side */
  nrx_engine *engine = nullptr;
  uint32_t next_slot = 0;
  std::vector<uint32_t> free_slots;
  std::vector<nrx_op> pending_ops;            // ops of the launch being assembled
  std::vector<char> pending_parent;           // [node] 1 if the node's trees are outputs of pending_ops
  PlanCache *plan = nullptr;
  bool use_plan_cache = true;
  /* score-only evaluation (candidate scoring reads nothing but the lnL): a full evaluation that replays a fused-K3 plan does not
   * store the CLVs of the root displayed trees.  They are then stale (`root_clvs_stale`): the next incremental evaluation, re-rooting
   * or CLV read-back first runs a full evaluation with the stores on. */
  bool score_only = false, root_clvs_stale = false, score_only_now = false;
  /* virtual re-rooting writes into shadow slots (the root-directed CLVs stay intact) and memoises the re-rooted node data */
  RerootCache *reroot = nullptr;
  std::vector<uint64_t> node_version;         // [node] bumped whenever the node's root-directed displayed trees are recomputed
  uint64_t prob_epoch = 1;                    // bumped by setReticulationProb: drops the plan's cached list of root trees to evaluate
  uint64_t topology_epoch = 1;                // bumped by topology_changed(): drops the per-branch re-rooting plans (paths, restriction sets)
  uint64_t clv_epoch = 0;                     // bumped by invalidateAllCLVs / setReticulationProb / topology_changed: drops the memoised data
  size_t reroot_cache_max_slots = SIZE_MAX;   // SIZE_MAX: default budget (env NRX_REROOT_CACHE_SLOTS); 0: nothing survives a re-rooting session
  uint64_t reroot_hits = 0, reroot_misses = 0;
  /* lazy re-rooting (opt-in): optimize_branch / the brlen_* flow neither evaluate the network from its root before a re-rooting nor
   * after the branch is done, when the branch is active and alive in every displayed tree; the lnL they return is then the edge-rooted
   * one (equal to the root's to rounding), and stale root-directed CLVs are recomputed when a later re-rooting reads them or by the
   * next computeLoglikelihood (host/brlen.cpp: lazyRerootPossible) */
  bool lazy_reroot = false;
  uint64_t lazy_sessions = 0, lazy_fallbacks = 0;   // re-rootings prepared lazily / of those, redone from the root because old per-tree lnLs were needed after all
  double lazy_last_logl = -std::numeric_limits<double>::infinity();
  uint64_t clv_site_updates = 0;              // Σ trees × local patterns actually launched
  std::vector<char> pseudo_clv_valid;         // [node]  (src/graph/AnnotatedNetwork.hpp:64; tips: valid)
  std::vector<uint32_t> pseudo_slot;          // [node]  device slot of the node's pseudo-likelihood CLV (UINT32_MAX: none yet)
  bool pending_eval = false;                  // root-tree lnLs enqueued on the engine stream, not yet collected (batched scoring)
  bool begin_returns_cached = false, begin_ran_traversal = false;
  size_t pending_root = 0;
  std::vector<size_t> pending_trees;
  ~AnnotatedNetwork();
  AnnotatedNetwork() = default;
  AnnotatedNetwork(const AnnotatedNetwork &) = delete;
  AnnotatedNetwork &operator=(const AnnotatedNetwork &) = delete;
};

/* set-up: role of init_annotated_network (src/graph/AnnotatedNetwork.cpp:80-185) + createNetworkPllTreeinfo */
struct PartitionInput {
  PartitionModel model;
  const uint32_t *tip_masks = nullptr;        // [tips][sites]
  const uint32_t *pattern_weights = nullptr;  // [sites] or null
};
void init_annotated_network(AnnotatedNetwork &ann, const std::vector<PartitionInput> &parts, int device);
void topology_changed(AnnotatedNetwork &ann);  // callers that edit the topology must call this (drops the plan cache)

/* ---- the likelihood API --------------------------------------------------------------------------------- */
double computeLoglikelihood(AnnotatedNetwork &ann_network, int incremental = 1, int update_pmatrices = 1);
double computeLoglikelihoodImproved(AnnotatedNetwork &ann_network, int incremental, int update_pmatrices);
double computePseudoLoglikelihood(AnnotatedNetwork &ann_network, int incremental = 1, int update_pmatrices = 1);  // src/likelihood/PseudoLoglikelihood.hpp
double scoreNetworkPseudo(AnnotatedNetwork &ann_network);  // src/likelihood/ComplexityScoring.cpp:69-79
/* batched scoring of candidate networks (one engine / stream each): result[i] == computeLoglikelihood(*anns[i], ...) */
std::vector<double> computeLoglikelihoodBatch(const std::vector<AnnotatedNetwork *> &anns, int incremental = 1, int update_pmatrices = 1);
void processNodeImproved(AnnotatedNetwork &ann_network, int incremental, Node *node, std::vector<Node *> &children,
                         const ReticulationConfigSet &extraRestrictions, bool append = false);
double evaluateTreesPartition(AnnotatedNetwork &ann_network, size_t partition_idx, std::vector<TreeLoglData> &treeLoglData);
int pllmod_treeinfo_update_prob_matrices(AnnotatedNetwork &ann_network, int update_all);  // PLLMOD/tree/treeinfo.c:842-880

std::vector<DisplayedTreeData> extractOldTrees(AnnotatedNetwork &ann_network, Node *virtual_root);
ReticulationConfigSet getRestrictionsActiveAliveBranch(AnnotatedNetwork &ann_network, size_t pmatrix_index);
void updateCLVsVirtualRerootTrees(AnnotatedNetwork &ann_network, Node *old_virtual_root, Node *new_virtual_root,
                                  Node *new_virtual_root_back, ReticulationConfigSet &restrictions);
/* ends the re-rooting session updateCLVsVirtualRerootTrees opened: the root-directed displayed trees come back (they were never
 * overwritten), and — only if the branch length of the session's edge differs from the one it started with — the edge's P-matrix
 * and the CLVs above it are invalidated (the reference's unconditional `invalidatePmatrixIndex` "restore the network root",
 * src/optimization/BranchLengthOptimization.cpp:417-419, recomputes what it had overwritten).  No-op without an open session. */
void finishVirtualReroot(AnnotatedNetwork &ann_network);
/* thrown by computeLoglikelihoodBrlenOpt inside a LAZY re-rooting session when a displayed tree turns out to need the per-tree lnL of
 * the old root (a source / target tree without a compatible partner): the caller redoes the re-rooting from an evaluated root */
struct LazyRerootNeedsRoot : std::runtime_error { LazyRerootNeedsRoot() : std::runtime_error("lazy re-rooting: the old root's per-tree lnLs are needed") {} };
void redoRerootFromRoot(AnnotatedNetwork &ann_network, unsigned int pmatrix_index, std::vector<DisplayedTreeData> &oldTrees);
void dropRerootCache(AnnotatedNetwork &ann_network);   // releases every memoised slot (keeps an open session's own data)
double computeLoglikelihoodBrlenOpt(AnnotatedNetwork &ann_network, const std::vector<DisplayedTreeData> &oldTrees,
                                    unsigned int pmatrix_index, int update_pmatrices = 1, bool print_extra_debug_info = false);
std::vector<std::vector<SumtableInfo>> computePartitionSumtables(AnnotatedNetwork &ann_network, unsigned int pmatrix_index);
/* computeLoglikelihoodBrlenOpt + computePartitionSumtables of the same branch in ONE pass over the displayed-tree pairs' CLVs (what
 * optimize_branch asks for back to back, src/optimization/BranchLengthOptimization.cpp:374-381): returns the edge-rooted lnL and
 * fills `sumtables`.  Same results as the two calls; partitions the fused kernel does not cover take the two engine calls. */
double computeLoglikelihoodBrlenOptAndSumtables(AnnotatedNetwork &ann_network, const std::vector<DisplayedTreeData> &oldTrees, unsigned int pmatrix_index,
                                                std::vector<std::vector<SumtableInfo>> &sumtables, int update_pmatrices = 1);
LoglDerivatives computeLoglikelihoodDerivatives(AnnotatedNetwork &ann_network,
                                                const std::vector<std::vector<SumtableInfo>> &sumtables,
                                                unsigned int pmatrix_index);

/* ---- the immediate callers: src/optimization/BranchLengthOptimization.hpp:12-32,
 *      ReticulationOptimization.hpp:16-18.  Same names and argument meaning; the 1-D minimisers they drive
 *      (pll-modules' Newton-Raphson and Brent, PLLMOD/optimize/opt_algorithms.c:133-261,1043-1254,1404-1429) are restated
 *      in host/optimize.cpp. ------------------------------------------------------------------------------- */
double optimize_branch(AnnotatedNetwork &ann_network, size_t pmatrix_index, BrlenOptMethod brlenOptMethod, unsigned int max_iters);
double optimize_branches(AnnotatedNetwork &ann_network, int max_iters, int max_iters_outside, int radius,
                         std::unordered_set<size_t> candidates, bool restricted_total_iters = false);
double optimize_branches(AnnotatedNetwork &ann_network, int max_iters, int max_iters_outside, int radius,
                         bool restricted_total_iters = false);
double optimize_reticulation(AnnotatedNetwork &ann_network, size_t reticulation_index);
double optimize_reticulations(AnnotatedNetwork &ann_network, int max_iters);
/* model-parameter loop: the model of partition p changed on the host -> upload it, invalidate its P-matrices and all CLVs */
void pushPartitionModel(AnnotatedNetwork &ann_network, unsigned partition);
void setAlpha(AnnotatedNetwork &ann_network, unsigned partition, double alpha);
void setPinv(AnnotatedNetwork &ann_network, unsigned partition, double prop_invar);   // pll_update_invariant_sites_proportion (LIBPLL/models.c:495-543)
/* the ALPHA step of optimize_params (src/optimization/ModelOptimization.cpp:56-65); defaults = PLLMOD_OPT_MIN/MAX_ALPHA, RAXML_PARAM_EPSILON */
double optimize_alpha(AnnotatedNetwork &ann_network, double min_alpha = 0.0201, double max_alpha = 100.0, double tolerance = 0.001);
/* the PINV step of optimize_params (src/optimization/ModelOptimization.cpp:67-76): pllmod_algo_opt_onedim_treeinfo(PLLMOD_OPT_PARAM_PINV)
 * over the partitions with +I (PLLMOD_OPT_MIN_PINV / MAX_PINV / RAXML_PARAM_EPSILON) */
double optimize_pinv(AnnotatedNetwork &ann_network, double min_pinv = 0.0, double max_pinv = 0.99, double tolerance = 0.001);
/* pllmod_algo_opt_brlen_scalers_treeinfo (PLLMOD/algorithm/pllmod_algorithm.c:869-960); returns the log-likelihood */
double optimize_brlen_scalers(AnnotatedNetwork &ann_network, double min_scaler, double max_scaler, double min_brlen, double max_brlen, double lh_epsilon);
double optimize_scalers(AnnotatedNetwork &ann_network, bool silent = true);   // src/optimization/BranchLengthOptimization.cpp:581-599; returns the BIC

/* ---- src/likelihood/ComplexityScoring.hpp:7-13 (what the search calls a network's score) ------------------------- */
size_t get_param_count(AnnotatedNetwork &ann_network);
size_t get_sample_size(AnnotatedNetwork &ann_network);
double aic(AnnotatedNetwork &ann_network, double logl);
double aicc(AnnotatedNetwork &ann_network, double logl);
double bic(AnnotatedNetwork &ann_network, double logl);
double scoreNetwork(AnnotatedNetwork &ann_network);

/* ---- src/optimization/Optimization.hpp:12-25: the non-topology optimisation rounds around the path ---------------- */
enum class OptimizeAllNonTopologyType { QUICK = 0, NORMAL = 1, SLOW = 2 };
void optimizeBranches(AnnotatedNetwork &ann_network, double brlen_smooth_factor = 1.0, bool silent = true, bool restricted_total_iters = false);
void optimizeModel(AnnotatedNetwork &ann_network, bool silent = true);
void optimizeReticulationProbs(AnnotatedNetwork &ann_network, bool silent = true);
void optimizeAllNonTopology(AnnotatedNetwork &ann_network, OptimizeAllNonTopologyType type, bool silent = true);

/* the slot pll-modules' optimisers re-enter through (src/RaxmlWrapper.cpp:21-26, src/RaxmlWrapper.hpp:20-24):
 * pllmod_treeinfo_t::likelihood_target_function = network_logl_wrapper, likelihood_computation_params = NetworkParams* */
struct NetworkParams {
  AnnotatedNetwork *ann_network;
  explicit NetworkParams(AnnotatedNetwork *a) : ann_network(a) {}
};
double network_logl_wrapper(void *network_params, int incremental, int update_pmatrices, double **persite_lnl);

/* ---- src/helper/InvalidationHelper.cpp -------------------------------------------------------------------- */
void invalidateSingleClv(AnnotatedNetwork &ann_network, unsigned int clv_index);
void invalidateHigherCLVs(AnnotatedNetwork &ann_network, const Node *node, bool invalidate_myself);
void invalidatePmatrixIndex(AnnotatedNetwork &ann_network, size_t pmatrix_index);
void invalidPmatrixIndexOnly(AnnotatedNetwork &ann_network, size_t pmatrix_index);
void invalidateAllCLVs(AnnotatedNetwork &ann_network);
void invalidateTreeLogprobs(AnnotatedNetwork &ann_network);
bool allClvsValid(AnnotatedNetwork &ann_network, size_t clv_index);
void setReticulationProb(AnnotatedNetwork &ann_network, size_t reticulation_idx, double prob);

}  // namespace netrax
