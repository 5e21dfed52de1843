/*
 * netrax_port.hpp — ORACLE (test infrastructure, NOT product code).
 *
 * C++ restatement of the NetRAX network-likelihood layer (SURVEY.md §8a rows a1–a18) on top of an
 * abstract kernel backend.  Two backends exist:
 *   backend_port.cpp — the scalar restatement in pll_port.c            (kind "port")
 *   backend_ref.cpp  — the reference's real forked libpll, _ref/ only  (kind "reference")
 * Parity status at THIS layer: the reference's own tests pin no golden lnL values, so
 * absolute network lnLs are pinned through (a) the real libpll underneath (backend_ref), (b) the
 * reference's own invariants (full == incremental, re-rooting preserves lnL, improved == naive
 * per-displayed-tree evaluation) — see tests/test_oracle_netrax.py.
 *
 * Reference paths: LH = /root/reference/src/likelihood, SRC = /root/reference/src.
 */
#pragma once

#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace orc {

/* ---- SRC/graph/ReticulationConfigSet.hpp:8-82 ------------------------------------------- */
enum class RS : unsigned char { DONT_CARE = 0, TAKE_FIRST_PARENT = 1, TAKE_SECOND_PARENT = 2, INVALID = 3 };
using Choices = std::vector<RS>;

struct ConfigSet {
  std::vector<Choices> configs;
  size_t max_reticulations = 0;
  ConfigSet() = default;
  explicit ConfigSet(size_t m) : max_reticulations(m) {}
  bool empty() const { return configs.empty(); }
  bool operator==(const ConfigSet &o) const;  // order-insensitive, as the reference
};

bool reticulationConfigsCompatible(const ConfigSet &l, const ConfigSet &r);
ConfigSet combineReticulationChoices(const ConfigSet &l, const ConfigSet &r);
void simplifyReticulationChoices(ConfigSet &res);
double computeReticulationConfigLogProb(const ConfigSet &c, const std::vector<double> &first,
                                        const std::vector<double> &second);
double computeReticulationConfigProb(const ConfigSet &c, const std::vector<double> &first,
                                     const std::vector<double> &second);
std::string configToString(const ConfigSet &c, size_t nret);

/* ---- flat rooted network (stands in for SRC/graph/Network.hpp; pointer surgery is out of scope) */
struct Network {
  struct Edge { unsigned source, target; double length, prob; };
  struct Node {
    bool is_ret = false;
    unsigned ret_index = 0;
    std::vector<unsigned> parents;   // 1 (basic) or 2 (reticulation: first, second); root: 0
    std::vector<unsigned> children;  // by ascending pmatrix index
    std::vector<unsigned> neighbors() const;  // parents first, then children (link order)
  };
  struct Ret { unsigned node, first_parent, second_parent, child, first_edge, second_edge; };
  unsigned num_tips = 0, root = 0;
  std::vector<Node> nodes;
  std::vector<Edge> edges;
  std::vector<Ret> rets;
  std::vector<unsigned char> toggle;  // ReticulationData::active_parent_toggle — STATEFUL like the reference

  size_t num_nodes() const { return nodes.size(); }
  size_t num_branches() const { return edges.size(); }
  size_t num_reticulations() const { return rets.size(); }
  unsigned edgeBetween(unsigned a, unsigned b) const;  // getEdgeTo
  unsigned activeParent(unsigned node) const;          // getActiveParent; UINT_MAX for root
  std::vector<unsigned> activeAliveChildren(const std::vector<bool> &dead, unsigned node) const;
  std::vector<unsigned> activeNeighbors(unsigned node) const;
  std::vector<bool> collectDeadNodes(unsigned megablobRoot, unsigned *displayed_tree_root) const;
  std::vector<unsigned> reversedTopologicalSort() const;
  void build(unsigned num_tips, unsigned num_nodes, unsigned root, const std::vector<Edge> &edges,
             const std::vector<unsigned> &ret_node, const std::vector<unsigned> &ret_first_edge,
             const std::vector<unsigned> &ret_second_edge);
};

/* ---- 32-byte aligned buffer with deep-copy semantics (DisplayedTreeData copies CLVs,
 *      SRC/graph/DisplayedTreeData.cpp:66-90) --------------------------------------------- */
template <class T> struct ABuf {
  T *p = nullptr;
  size_t n = 0;
  ABuf() = default;
  explicit ABuf(size_t n_) { alloc(n_); }
  ABuf(const ABuf &o) { alloc(o.n); if (n) std::memcpy(p, o.p, n * sizeof(T)); }
  ABuf(ABuf &&o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
  ABuf &operator=(const ABuf &o) { if (this != &o) { release(); alloc(o.n); if (n) std::memcpy(p, o.p, n * sizeof(T)); } return *this; }
  ABuf &operator=(ABuf &&o) noexcept { if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; } return *this; }
  ~ABuf() { release(); }
  void alloc(size_t n_) {
    n = n_;
    if (!n) { p = nullptr; return; }
    size_t bytes = ((n * sizeof(T) + 31) / 32) * 32;
    p = static_cast<T *>(std::aligned_alloc(32, bytes));
    if (!p) throw std::bad_alloc();
    std::memset(p, 0, bytes);
  }
  void release() { std::free(p); p = nullptr; n = 0; }
  void zero() { if (n) std::memset(p, 0, n * sizeof(T)); }
};

/* ---- kernel backend: the seven forked libpll entry points (SURVEY §8b lower seam) -------- */
struct Operand {
  int kind = 2;  // 0 inner CLV, 1 tip, 2 fake (all-ones CLV, identity P)
  const double *clv = nullptr;
  const unsigned *scaler = nullptr;
  unsigned tip = 0;
  unsigned edge = 0;
};

struct PartitionDesc {
  unsigned states = 4, rate_cats = 4, sites = 0;
  std::vector<double> freqs, subst_params, rates, rate_weights;
  std::vector<unsigned> pattern_weights;
  std::vector<std::vector<uint32_t>> tip_masks;  // [tip][site] state bit masks
};

struct Backend {
  virtual ~Backend() {}
  virtual const char *kind() const = 0;
  virtual unsigned partitionCount() const = 0;
  virtual unsigned sites(unsigned p) const = 0;
  virtual size_t clvEntries(unsigned p) const = 0;  // sites*rate_cats*states_padded
  virtual unsigned statesPadded(unsigned p) const = 0;
  virtual unsigned rateCats(unsigned p) const = 0;
  virtual unsigned states(unsigned p) const = 0;
  virtual void setModel(unsigned p, const double *freqs, const double *subst, const double *rates,
                        const double *weights) = 0;  // recomputes eigen
  virtual void getEigen(unsigned p, double *eigenvecs, double *inv_eigenvecs, double *eigenvals) const = 0;
  virtual void getRates(unsigned p, double *rates, double *weights, double *freqs) const = 0;
  /* n rate matrices, category c uses matrix cat_model[c] (libpll params_indices; LG4M / LG4X): pll_set_frequencies /
   * pll_set_subst_params / pll_update_eigen per matrix index */
  virtual void setSubmodels(unsigned p, unsigned n, const unsigned *cat_model, const double *freqs, const double *subst) = 0;
  virtual void setPinv(unsigned p, double prop_invar) = 0;                                 // pll_update_invariant_sites_proportion
  virtual void setCategoryRates(unsigned p, const double *rates) = 0;                       // pll_set_category_rates
  virtual bool gammaRates(double alpha, unsigned cats, double *out, int mode) const = 0;   // pll_compute_gamma_cats
  virtual void updatePmatrix(unsigned p, unsigned edge, double brlen) = 0;
  virtual const double *pmatrix(unsigned p, unsigned edge) const = 0;
  virtual void updatePartials(unsigned p, double *parent_clv, unsigned *parent_scaler,
                              const Operand &l, const Operand &r) = 0;
  virtual double rootLogl(unsigned p, const double *clv, const unsigned *scaler, double *persite) = 0;
  virtual double edgeLogl(unsigned p, const Operand &parent, const Operand &child, unsigned edge,
                          double *persite) = 0;
  virtual void sumtable(unsigned p, const Operand &parent, const Operand &child, double *out) = 0;
  virtual void derivatives(unsigned p, const double *sumtable, double brlen, bool want_f, double *f,
                           double *d1, double *d2) = 0;
};
Backend *makePortBackend(unsigned tips, unsigned edges_plus_fake, const std::vector<PartitionDesc> &parts);
#ifdef ORC_HAVE_REF
Backend *makeRefBackend(unsigned tips, unsigned edges_plus_fake, const std::vector<PartitionDesc> &parts);
#endif

/* ---- SRC/graph/TreeLoglData.hpp, DisplayedTreeData.hpp, NodeDisplayedTreeData.hpp -------- */
struct TreeLoglData {
  ConfigSet reticulationChoices;
  double tree_logprob = 0.0;
  bool tree_logprob_valid = false;
  std::vector<double> tree_partition_logl;
  bool tree_logl_valid = false;
  TreeLoglData() = default;
  TreeLoglData(size_t nparts, size_t max_ret) : reticulationChoices(max_ret), tree_partition_logl(nparts, 0.0) {}
};

struct DisplayedTreeData {
  TreeLoglData treeLoglData;
  std::vector<ABuf<double>> clv_vector;       // per partition (empty for tips: PATTERN_TIP)
  std::vector<ABuf<unsigned>> scale_buffer;   // per partition
  bool clv_valid = false;
  bool isTip = false;
  unsigned tip = 0;
};

struct NodeDisplayedTreeData {
  std::vector<DisplayedTreeData> displayed_trees;
  size_t num_active_displayed_trees = 0;
};

enum class LikelihoodVariant { AVERAGE_DISPLAYED_TREES = 0, BEST_DISPLAYED_TREE = 1, SARAH_PSEUDO = 2 };
enum { BRLEN_LINKED = 0, BRLEN_SCALED = 1, BRLEN_UNLINKED = 2 };  // PLLMOD_COMMON_BRLEN_*

struct Options {  // SRC/NetraxOptions.hpp
  LikelihoodVariant likelihood_variant = LikelihoodVariant::AVERAGE_DISPLAYED_TREES;
  int brlen_linkage = BRLEN_LINKED;
  size_t max_reticulations = 32;
  double min_interesting_tree_logprob = -13.815510557964274;  // log(1e-6), NetraxOptions.hpp:108
};

struct OptOptions {  // SRC/NetraxOptions.hpp:100-105 — what the optimisers read
  double brlen_min = 1e-6, brlen_max = 100.0, brprob_min = 1e-6, brprob_max = 1.0 - 1e-6, tolerance = 0.1;
};
enum { OPT_BRENT_NORMAL = 0, OPT_BRENT_REROOT = 1, OPT_NEWTON_RAPHSON = 2 };  // BrlenOptMethod, SRC/NetraxOptions.hpp:17-21

struct SumtableInfo {  // LH/LikelihoodDerivatives.hpp:9-74
  double tree_prob = 0.0;
  ABuf<double> sumtable;
  size_t left_tree_idx = 0, right_tree_idx = 0;
};

struct LoglDerivatives {  // LH/LikelihoodDerivatives.hpp:76-81
  double logl_prime = std::numeric_limits<double>::infinity();
  double logl_prime_prime = std::numeric_limits<double>::infinity();
  std::vector<double> partition_logl_prime, partition_logl_prime_prime;
  // raw per-sumtable values (f, d1, d2) per partition, kept for kernel-level parity tests
  std::vector<std::vector<double>> raw;  // [partition][3*sumtable]
};

struct AnnotatedNetwork {  // SRC/graph/AnnotatedNetwork.hpp:42-89 (fields the path reads)
  Options options;
  OptOptions opt;
  Network network;
  std::unique_ptr<Backend> backend;
  std::vector<double> reticulation_probs, first_parent_logprobs, second_parent_logprobs;
  std::vector<NodeDisplayedTreeData> pernode_displayed_tree_data;
  std::vector<unsigned> travbuffer;
  std::vector<std::vector<char>> clv_valid;       // [partition][node]  (fake_treeinfo->clv_valid)
  std::vector<std::vector<char>> pmatrix_valid;   // [partition][edge]
  std::vector<std::vector<double>> branch_lengths;  // [partition][edge+1] (fake_treeinfo->branch_lengths)
  std::vector<double> linked_branch_lengths;        // [edge+1]
  std::vector<double> brlen_scalers;                // [partition] pllmod_treeinfo_t::brlen_scalers (BRLEN_SCALED only; empty = all 1)
  std::vector<double> partition_loglh;
  // pseudo-likelihood state (SRC/graph/AnnotatedNetwork.hpp:64-68): one CLV + scaler per node and partition (the
  // reference keeps them in partition->clv[node] / scale_buffer[node]) and the three scratch CLVs
  std::vector<char> pseudo_clv_valid;                       // [node]
  std::vector<std::vector<ABuf<double>>> pseudo_clv;        // [node][partition]
  std::vector<std::vector<ABuf<unsigned>>> pseudo_scaler;   // [node][partition]
  std::vector<ABuf<double>> tmp_clv_1, tmp_clv_2, tmp_clv_3; // [partition]
  size_t total_num_model_parameters = 0, total_num_sites = 0;  // SRC/graph/AnnotatedNetwork.hpp:47-48
  std::vector<double> pinvs;   // partition->prop_invar[param_indices[p][0]] as last set (0 = no +I)
  std::vector<double> pattern_weight_sums;  // [partition] pll_partition_t::pattern_weight_sum
  std::vector<int> params_to_optimize;   // pllmod_treeinfo_t::params_to_optimize per partition (bit 0 alpha, bit 1 pinv); empty / -1 = derive from the values
  std::vector<double> alphas;  // fake_treeinfo->alphas (0 = no Gamma shape attached to the partition's rates)
  double cached_logl = 0;
  bool cached_logl_valid = false;
  // statistics for the benchmark metric (Σ trees(node) × patterns)
  uint64_t n_clv_updates = 0;
  // fake_treeinfo->parallel_reduce_cb / parallel_context (SRC/RaxmlWrapper.cpp:717-718): SUM over site shards; null = single process
  void (*parallel_reduce_cb)(void *context, double *data, size_t count, int op) = nullptr;
  void *parallel_context = nullptr;
  unsigned partitionCount() const { return backend->partitionCount(); }
  unsigned fakePmatrixIndex() const { return (unsigned)network.edges.size(); }
};

void init_annotated_network(AnnotatedNetwork &ann);  // SRC/graph/AnnotatedNetwork.cpp:80-185

/* upper seam, same names/arguments as the reference */
double computeLoglikelihood(AnnotatedNetwork &ann, int incremental = 1, int update_pmatrices = 1);
double computePseudoLoglikelihood(AnnotatedNetwork &ann, int incremental = 1, int update_pmatrices = 1);  // LH/PseudoLoglikelihood.cpp:57-226
double persiteLoglikelihood(AnnotatedNetwork &ann, size_t tree, unsigned p, double *persite);
double computeLoglikelihoodNaive(AnnotatedNetwork &ann, std::vector<double> *tree_logl, std::vector<double> *tree_logprob);
void invalidateSingleClv(AnnotatedNetwork &ann, unsigned clv_index);
void invalidateHigherCLVs(AnnotatedNetwork &ann, unsigned node, bool invalidate_myself);
void invalidatePmatrixIndex(AnnotatedNetwork &ann, size_t pmatrix_index);
void invalidPmatrixIndexOnly(AnnotatedNetwork &ann, size_t pmatrix_index);
void invalidateAllCLVs(AnnotatedNetwork &ann);
void invalidateTreeLogprobs(AnnotatedNetwork &ann);
void setReticulationProb(AnnotatedNetwork &ann, size_t ret, double prob);  // ReticulationOptimization.cpp:25-38
void setBranchLength(AnnotatedNetwork &ann, int partition /* -1: all/linked */, size_t pmatrix_index, double value);
void updateProbMatrices(AnnotatedNetwork &ann, int update_all);  // pllmod_treeinfo_update_prob_matrices

std::vector<DisplayedTreeData> extractOldTrees(AnnotatedNetwork &ann, unsigned virtual_root);
ConfigSet getRestrictionsActiveAliveBranch(AnnotatedNetwork &ann, size_t pmatrix_index);
void updateCLVsVirtualRerootTrees(AnnotatedNetwork &ann, unsigned old_virtual_root, unsigned new_virtual_root,
                                  unsigned new_virtual_root_back, ConfigSet &restrictions);
double computeLoglikelihoodBrlenOpt(AnnotatedNetwork &ann, const std::vector<DisplayedTreeData> &oldTrees,
                                    unsigned pmatrix_index, int update_pmatrices = 1);
std::vector<std::vector<SumtableInfo>> computePartitionSumtables(AnnotatedNetwork &ann, unsigned pmatrix_index);
LoglDerivatives computeLoglikelihoodDerivatives(AnnotatedNetwork &ann,
                                                const std::vector<std::vector<SumtableInfo>> &sumtables,
                                                unsigned pmatrix_index);

/* the immediate callers (optimize_port.cpp) */
double optimize_branch(AnnotatedNetwork &ann, size_t pmatrix_index, int method, unsigned max_iters);
double optimize_branches(AnnotatedNetwork &ann, int max_iters, int max_iters_outside, int radius, int method, bool restricted_total_iters = false);
double optimize_reticulation(AnnotatedNetwork &ann, size_t reticulation_index);
double optimize_reticulations(AnnotatedNetwork &ann, int max_iters);
double scoreNetwork(AnnotatedNetwork &ann);                               // LH/ComplexityScoring.cpp:57-67 (BIC)
void optimizeAllNonTopology(AnnotatedNetwork &ann, int type /* 0 QUICK, 1 NORMAL, 2 SLOW */);  // SRC/optimization/Optimization.cpp:118-214
void setPinv(AnnotatedNetwork &ann, unsigned partition, double prop_invar);
void setBrlenScaler(AnnotatedNetwork &ann, unsigned partition, double scaler);
void setSubmodels(AnnotatedNetwork &ann, unsigned partition, unsigned n, const unsigned *cat_model, const double *freqs, const double *subst);
void setAlpha(AnnotatedNetwork &ann, unsigned partition, double alpha);   // treeinfo_set_alpha (PLLMOD/algorithm/pllmod_algorithm.c:566-587)
double optimize_alpha(AnnotatedNetwork &ann, double min_alpha, double max_alpha, double tolerance);  // pllmod_algo_opt_onedim_treeinfo(ALPHA)
double optimize_pinv(AnnotatedNetwork &ann, double min_pinv, double max_pinv, double tolerance);     // pllmod_algo_opt_onedim_treeinfo(PINV)
double optimize_brlen_scalers(AnnotatedNetwork &ann, double min_scaler, double max_scaler, double min_brlen, double max_brlen, double lh_epsilon);
double optimize_scalers(AnnotatedNetwork &ann);  // SRC/optimization/BranchLengthOptimization.cpp:581-599; returns the BIC

}  // namespace orc
