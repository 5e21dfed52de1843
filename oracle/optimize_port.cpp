/*
 * optimize_port.cpp — ORACLE (test infrastructure, NOT product code).
 *
 * Restatement of the immediate callers of the likelihood path (SRC = /root/reference/src):
 *   optimize_branch / optimize_branches      SRC/optimization/BranchLengthOptimization.cpp:63-156,165-241,285-476,567-576
 *   optimize_reticulation(s)                 SRC/optimization/ReticulationOptimization.cpp:40-46,68-117
 *   optimize_pinv / optimize_scalers         the PINV step (SRC/optimization/ModelOptimization.cpp:67-76) and
 *                                            pllmod_algo_opt_brlen_scalers_treeinfo (pllmod_algorithm.c:869-960) behind
 *                                            SRC/optimization/BranchLengthOptimization.cpp:581-599
 *   optimize_alpha                           the ALPHA step of optimize_params (SRC/optimization/ModelOptimization.cpp:56-65) =
 *                                            pllmod_algo_opt_onedim_treeinfo (PLLMOD/algorithm/pllmod_algorithm.c:743-866) with
 *                                            target_func_onedim_treeinfo (algo_callback.c:295-363) and treeinfo_set_alpha (:566-587)
 * written the way the reference writes them: parameter structs + C callbacks handed to pll-modules' minimisers.
 * In the `_ref` build (ORC_HAVE_REF) the minimisers ARE the reference's pllmod_opt_minimize_newton_multi /
 * pllmod_opt_minimize_brent (opt_algorithms.c compiled where it lies); in the port build they are opt_port.c.
 */
#include <cmath>
#include <limits>
#include <stdexcept>
#include <vector>
#include <unordered_set>

#include "netrax_port.hpp"

extern "C" {
#ifdef ORC_HAVE_REF
int pllmod_opt_minimize_newton_multi(unsigned int xnum, double xmin, double *xguess, double xmax, double tolerance, unsigned int max_iters,
                                     int *converged, void *params, void(deriv_func)(void *, double *, double *, double *));
double pllmod_opt_minimize_brent(double xmin, double xguess, double xmax, double xtol, double *fx, double *f2x, void *params,
                                 double (*target_funk)(void *, double));
int pllmod_opt_minimize_brent_multi(unsigned int xnum, int *opt_mask, double *xmin, double *xguess, double *xmax, double xtol, double *xopt,
                                    double *fx, double *f2x, void *params, double (*target_funk)(void *, double *, double *, int *), int global_range);
#define MIN_NEWTON pllmod_opt_minimize_newton_multi
#define MIN_BRENT pllmod_opt_minimize_brent
#define MIN_BRENT_MULTI pllmod_opt_minimize_brent_multi
#else
int orcopt_newton_multi(unsigned int xnum, double xmin, double *xguess, double xmax, double tolerance, unsigned int max_iters, int *converged,
                        void *params, void (*deriv_func)(void *, double *, double *, double *));
double orcopt_brent(double xmin, double xguess, double xmax, double xtol, double *fx, double *f2x, void *params,
                    double (*target_funk)(void *, double));
int orcopt_brent_multi(unsigned int xnum, int *opt_mask, double *xmin, double *xguess, double *xmax, double xtol, double *xopt,
                       double *fx, double *f2x, void *params, double (*target_funk)(void *, double *, double *, int *), int global_range);
#define MIN_NEWTON orcopt_newton_multi
#define MIN_BRENT orcopt_brent
#define MIN_BRENT_MULTI orcopt_brent_multi
#endif
}

namespace orc {

namespace {
bool unlinked(const AnnotatedNetwork &ann) { return ann.options.brlen_linkage == BRLEN_UNLINKED; }
double currentBrlen(const AnnotatedNetwork &ann, size_t part, size_t edge) { return unlinked(ann) ? ann.branch_lengths[part][edge] : ann.linked_branch_lengths[edge]; }
void assignBrlen(AnnotatedNetwork &ann, size_t part, size_t edge, double v) {
  if (unlinked(ann)) ann.branch_lengths[part][edge] = v;
  else setBranchLength(ann, -1, edge, v);  // linked: all partitions read the one linked array
}

struct BrentBrlenParams {  // BranchLengthOptimization.cpp:55-61
  AnnotatedNetwork *ann;
  size_t edge, part;
  std::vector<DisplayedTreeData> *oldTrees;
  int method;
};

double brent_target_networks(void *p, double x) {  // :63-109
  BrentBrlenParams *q = static_cast<BrentBrlenParams *>(p);
  AnnotatedNetwork &ann = *q->ann;
  if (currentBrlen(ann, q->part, q->edge) == x)
    return q->method == OPT_BRENT_REROOT ? -1 * computeLoglikelihoodBrlenOpt(ann, *q->oldTrees, (unsigned)q->edge, 1) : -1 * computeLoglikelihood(ann);
  assignBrlen(ann, q->part, q->edge, x);
  if (q->method != OPT_BRENT_NORMAL) invalidPmatrixIndexOnly(ann, q->edge);
  else invalidatePmatrixIndex(ann, q->edge);
  return q->method == OPT_BRENT_REROOT ? -1 * computeLoglikelihoodBrlenOpt(ann, *q->oldTrees, (unsigned)q->edge, 1) : -1 * computeLoglikelihood(ann, 1, 1);
}

struct NewtonBrlenParams {  // :158-163
  AnnotatedNetwork *ann;
  size_t edge, part;
  std::vector<std::vector<SumtableInfo>> *sumtables;
  double new_brlen;
};

void network_derivative_func_multi(void *p, double *proposal, double *df, double *ddf) {  // :165-200
  (void)proposal;  // aliases q->new_brlen, as in the reference
  NewtonBrlenParams *q = static_cast<NewtonBrlenParams *>(p);
  AnnotatedNetwork &ann = *q->ann;
  assignBrlen(ann, q->part, q->edge, q->new_brlen);
  invalidPmatrixIndexOnly(ann, q->edge);
  LoglDerivatives d = computeLoglikelihoodDerivatives(ann, *q->sumtables, (unsigned)q->edge);
  if (unlinked(ann)) {
    // The reference writes all partition_count entries into the solver's 1-element arrays (:190-195); only element 0 is a
    // defined write and only element 0 is read back, so that is what is restated (quirk Q7).
    df[0] = d.partition_logl_prime[0];
    ddf[0] = d.partition_logl_prime_prime[0];
  } else {
    *df = d.logl_prime;
    *ddf = d.logl_prime_prime;
  }
}

double optimize_branch_partition(AnnotatedNetwork &ann, std::vector<DisplayedTreeData> &oldTrees, std::vector<std::vector<SumtableInfo>> &sumtables,
                                 size_t edge, size_t part, int method, unsigned max_iters) {  // :285-343
  ann.cached_logl_valid = false;
  double start_logl = method != OPT_BRENT_NORMAL ? computeLoglikelihoodBrlenOpt(ann, oldTrees, (unsigned)edge, 1) : computeLoglikelihood(ann);
  if (method == OPT_BRENT_NORMAL || method == OPT_BRENT_REROOT) {  // optimize_branch_brent :111-156
    BrentBrlenParams bp{&ann, edge, part, &oldTrees, method};
    double score = 0, f2x = 0;
    double nb = MIN_BRENT(ann.opt.brlen_min, currentBrlen(ann, part, edge), ann.opt.brlen_max, ann.opt.tolerance, &score, &f2x, &bp, &brent_target_networks);
    assignBrlen(ann, part, edge, nb);
    invalidatePmatrixIndex(ann, edge);
  } else {  // optimize_branch_newton_raphson :202-241
    double old_brlen = currentBrlen(ann, part, edge);
    double tolerance = ann.opt.brlen_min > 0 ? ann.opt.brlen_min / 10.0 : 1.0e-4;
    NewtonBrlenParams np{&ann, edge, part, &sumtables, old_brlen};
    MIN_NEWTON(1, ann.opt.brlen_min, &np.new_brlen, ann.opt.brlen_max, tolerance, max_iters, nullptr, &np, network_derivative_func_multi);
    double new_logl = computeLoglikelihoodBrlenOpt(ann, oldTrees, (unsigned)edge, 1);
    if (new_logl < start_logl) {
      assignBrlen(ann, part, edge, old_brlen);
      invalidPmatrixIndexOnly(ann, edge);
    }
  }
  return method != OPT_BRENT_NORMAL ? computeLoglikelihoodBrlenOpt(ann, oldTrees, (unsigned)edge, 1) : computeLoglikelihood(ann);
}

struct BrentBrprobParams { AnnotatedNetwork *ann; size_t ret; };  // ReticulationOptimization.cpp:19-22
double brent_target_networks_prob(void *p, double x) {  // :40-46
  BrentBrprobParams *q = static_cast<BrentBrprobParams *>(p);
  setReticulationProb(*q->ann, q->ret, x);
  return -1 * computeLoglikelihood(*q->ann, 1, 1);
}
}  // namespace

double optimize_branch(AnnotatedNetwork &ann, size_t edge, int method, unsigned max_iters) {  // :345-421
  double old_logl = computeLoglikelihood(ann);
  std::vector<DisplayedTreeData> oldTrees;
  std::vector<std::vector<SumtableInfo>> sumtables;
  if (method != OPT_BRENT_NORMAL) {
    oldTrees = extractOldTrees(ann, ann.network.root);
    ConfigSet restrictions = getRestrictionsActiveAliveBranch(ann, edge);
    updateCLVsVirtualRerootTrees(ann, ann.network.root, ann.network.edges[edge].source, ann.network.edges[edge].target, restrictions);
    ann.cached_logl_valid = false;
    double brlenopt_logl = computeLoglikelihoodBrlenOpt(ann, oldTrees, (unsigned)edge);
    if (std::fabs((double)(old_logl - brlenopt_logl >= 1E-3)))  // sic (:377): fabs of a comparison
      throw std::runtime_error("Something went wrong when rerooting CLVs during brlen optimization");
    if (method == OPT_NEWTON_RAPHSON) sumtables = computePartitionSumtables(ann, (unsigned)edge);
  }
  if (unlinked(ann)) {
    for (size_t p = 0; p < ann.partitionCount(); ++p) optimize_branch_partition(ann, oldTrees, sumtables, edge, p, method, max_iters);
  } else {
    optimize_branch_partition(ann, oldTrees, sumtables, edge, 0, method, max_iters);
  }
  if (method != OPT_BRENT_NORMAL) invalidatePmatrixIndex(ann, edge);
  return computeLoglikelihood(ann);
}

double optimize_branches(AnnotatedNetwork &ann, int max_iters, int max_iters_outside, int radius, int method, bool restricted_total_iters) {  // :423-476,567-576
  (void)radius;
  std::unordered_set<size_t> candidates;
  for (size_t i = 0; i < ann.network.num_branches(); ++i) candidates.emplace(i);
  double old_logl = computeLoglikelihood(ann, 1, 1);
  double start_logl = old_logl;
  std::vector<size_t> act_iters(ann.network.num_branches(), 0);
  size_t total_iters = 0;
  while (!candidates.empty()) {
    size_t edge = *candidates.begin();
    candidates.erase(candidates.begin());
    total_iters++;
    if (restricted_total_iters && total_iters >= (size_t)max_iters_outside) continue;
    if (act_iters[edge] >= (size_t)max_iters_outside) continue;
    act_iters[edge]++;
    old_logl = optimize_branch(ann, edge, method, (unsigned)max_iters);
  }
  if ((old_logl < start_logl) && (std::fabs(old_logl - start_logl) >= 1E-3)) throw std::runtime_error("Overall loglikelihood got worse");
  return old_logl;
}

double optimize_reticulation(AnnotatedNetwork &ann, size_t ret) {  // ReticulationOptimization.cpp:68-100
  computeLoglikelihood(ann, 1, 1);
  BrentBrprobParams bp{&ann, ret};
  double old_brprob = ann.reticulation_probs[ret];
  double score = 0, f2x = 0;
  setReticulationProb(ann, ret, 0.5);
  double nb = MIN_BRENT(ann.opt.brprob_min, old_brprob, ann.opt.brprob_max, ann.opt.tolerance, &score, &f2x, &bp, &brent_target_networks_prob);
  setReticulationProb(ann, ret, nb);
  return computeLoglikelihood(ann, 1, 1);
}

double optimize_reticulations(AnnotatedNetwork &ann, int max_iters) {  // :102-117
  double act_logl = computeLoglikelihood(ann, 1, 1);
  int act_iters = 0;
  while (act_iters < max_iters) {
    double loop_logl = act_logl;
    for (size_t i = 0; i < ann.network.num_reticulations(); ++i) loop_logl = optimize_reticulation(ann, i);
    act_iters++;
    if (loop_logl == act_logl) break;
    act_logl = loop_logl;
  }
  return act_logl;
}

void setPinv(AnnotatedNetwork &ann, unsigned p, double prop_invar) {
  if (ann.pinvs.size() < ann.partitionCount()) ann.pinvs.resize(ann.partitionCount(), 0.0);
  ann.pinvs[p] = prop_invar;
  ann.backend->setPinv(p, prop_invar);
  for (auto &v : ann.pmatrix_valid[p]) v = 0;
  invalidateAllCLVs(ann);
}

/* pllmod_treeinfo_t::brlen_scalers[p] (scaled branch-length linkage): P-matrices of partition p use scaler x linked length */
void setBrlenScaler(AnnotatedNetwork &ann, unsigned p, double scaler) {
  if (ann.options.brlen_linkage != BRLEN_SCALED) throw std::runtime_error("Branch length scalers exist only in scaled branch length mode.");
  if (ann.brlen_scalers.size() < ann.partitionCount()) ann.brlen_scalers.resize(ann.partitionCount(), 1.0);
  ann.brlen_scalers.at(p) = scaler;
  for (auto &v : ann.pmatrix_valid[p]) v = 0;
  invalidateAllCLVs(ann);
}

void setSubmodels(AnnotatedNetwork &ann, unsigned p, unsigned n, const unsigned *cat_model, const double *freqs, const double *subst) {
  ann.backend->setSubmodels(p, n, cat_model, freqs, subst);
  for (auto &v : ann.pmatrix_valid[p]) v = 0;
  invalidateAllCLVs(ann);
}

/* treeinfo_set_alpha: alpha -> discrete Gamma rates (mean mode, the raxml-ng default) -> partition rates; every P-matrix
 * and CLV of the partition is stale afterwards (the optimiser re-evaluates with incremental = 0 anyway) */
void setAlpha(AnnotatedNetwork &ann, unsigned p, double alpha) {
  if (ann.alphas.size() < ann.partitionCount()) ann.alphas.resize(ann.partitionCount(), 0.0);
  std::vector<double> rates(ann.backend->rateCats(p));
  if (!ann.backend->gammaRates(alpha, (unsigned)rates.size(), rates.data(), 0)) throw std::runtime_error("Invalid alpha value / GAMMA discretization mode");
  ann.alphas[p] = alpha;
  ann.backend->setCategoryRates(p, rates.data());
  for (auto &v : ann.pmatrix_valid[p]) v = 0;
  invalidateAllCLVs(ann);
}

namespace {
/* the three parameter kinds pllmod_algo_opt_onedim_treeinfo accepts (pllmod_algorithm.c:830-855) with their setters:
 * treeinfo_set_alpha (:566-587), treeinfo_set_pinv (:601-622), treeinfo_set_brlen_scaler (:636-647) */
enum OnedimParam { ONEDIM_ALPHA, ONEDIM_PINV, ONEDIM_BRLEN_SCALER };
struct OnedimOptParams { AnnotatedNetwork *ann; OnedimParam param; std::vector<unsigned> parts; };

void onedimSet(AnnotatedNetwork &ann, OnedimParam param, unsigned p, double x) {
  if (param == ONEDIM_ALPHA) setAlpha(ann, p, x);
  else if (param == ONEDIM_PINV) setPinv(ann, p, x);
  else ann.brlen_scalers.at(p) = x;   // no invalidation: the evaluation below runs with incremental = 0, which refreshes every P-matrix
}

double target_func_onedim(void *p, double *x, double *fx, int *converged) {  // algo_callback.c:295-363
  OnedimOptParams *q = static_cast<OnedimOptParams *>(p);
  AnnotatedNetwork &ann = *q->ann;
  double score = -std::numeric_limits<double>::infinity(), unconverged_flag = 0.;
  for (size_t j = 0; j < q->parts.size(); ++j) {
    if (converged && converged[j]) continue;
    unconverged_flag = 1.;
    if (x) onedimSet(ann, q->param, q->parts[j], x[j]);
  }
  if (x) score = -1 * computeLoglikelihood(ann, 0, 1);   // pllmod_treeinfo_compute_loglh(treeinfo, 0)
  if (fx) for (size_t j = 0; j < q->parts.size(); ++j) fx[j] = -1 * ann.partition_loglh[q->parts[j]];
  if (converged) {
    if (ann.parallel_reduce_cb) ann.parallel_reduce_cb(ann.parallel_context, &unconverged_flag, 1, 0);
    converged[q->parts.size()] = unconverged_flag > 0. ? 0 : 1;
  }
  return score;
}

double optimize_onedim(AnnotatedNetwork &ann, OnedimParam param, double min_value, double max_value, double tolerance) {  // pllmod_algorithm.c:743-818
  OnedimOptParams q{&ann, param, {}};
  const unsigned P = ann.partitionCount();
  if (ann.alphas.size() < P) ann.alphas.resize(P, 0.0);
  if (ann.pinvs.size() < P) ann.pinvs.resize(P, 0.0);
  if (param == ONEDIM_BRLEN_SCALER && ann.brlen_scalers.size() < P) ann.brlen_scalers.resize(P, 1.0);
  for (unsigned p = 0; p < P; ++p) {   // params_to_optimize[p] & param: raxml-ng sets ALPHA for +G, PINV for +I, the scaler bit under scaled linkage
    const int pto = p < ann.params_to_optimize.size() ? ann.params_to_optimize[p] : -1;   // (pllmod_algorithm.c:765-772)
    bool on = true;
    if (param == ONEDIM_ALPHA) on = pto >= 0 ? (pto & 1) != 0 && ann.alphas[p] > 0.0 : ann.alphas[p] > 0.0;
    else if (param == ONEDIM_PINV) on = pto >= 0 ? (pto & 2) != 0 : ann.pinvs[p] > 0.0;
    if (on) q.parts.push_back(p);
  }
  if (!q.parts.empty()) {
    std::vector<double> vals;
    for (unsigned p : q.parts) vals.push_back(param == ONEDIM_ALPHA ? ann.alphas[p] : param == ONEDIM_PINV ? ann.pinvs[p] : ann.brlen_scalers[p]);
    std::vector<int> mask(q.parts.size(), 1);
    MIN_BRENT_MULTI((unsigned)q.parts.size(), mask.data(), &min_value, vals.data(), &max_value, tolerance, vals.data(), nullptr, nullptr, &q,
                    &target_func_onedim, 1);
  }
  return computeLoglikelihood(ann, 0, 1);
}
}  // namespace

double optimize_alpha(AnnotatedNetwork &ann, double min_alpha, double max_alpha, double tolerance) {
  return optimize_onedim(ann, ONEDIM_ALPHA, min_alpha, max_alpha, tolerance);
}

double optimize_pinv(AnnotatedNetwork &ann, double min_pinv, double max_pinv, double tolerance) {  // SRC/optimization/ModelOptimization.cpp:67-76
  return optimize_onedim(ann, ONEDIM_PINV, min_pinv, max_pinv, tolerance);
}

/* pllmod_algo_opt_brlen_scalers_treeinfo (PLLMOD/algorithm/pllmod_algorithm.c:869-960) as it runs on NetRAX's fake treeinfo
 * (subnode_count == 0: the branches live in branch_lengths[0]) */
double optimize_brlen_scalers(AnnotatedNetwork &ann, double min_scaler, double max_scaler, double min_brlen, double max_brlen, double lh_epsilon) {
  if (ann.options.brlen_linkage != BRLEN_SCALED) throw std::runtime_error("Branch length scaler optimization works only in scaled branch length mode.");
  const unsigned P = ann.partitionCount();
  const size_t E = ann.network.edges.size();
  if (ann.brlen_scalers.size() < P) ann.brlen_scalers.resize(P, 1.0);
  const double old_loglh = computeLoglikelihood(ann, 0, 1);
  const std::vector<double> old_scalers = ann.brlen_scalers, old_brlen = ann.linked_branch_lengths;
  auto setLinked = [&](size_t e, double v) {   // branch_lengths[0][e] of the reference: the one array every partition reads; no invalidation
    ann.linked_branch_lengths[e] = v;
    for (auto &b : ann.branch_lengths) b[e] = v;
  };
  auto scaleBranchesAll = [&](double f) {   // pllmod_treeinfo_scale_branches_all (PLLMOD/tree/treeinfo.c:1132-1155)
    for (size_t e = 0; e < E; ++e) setLinked(e, ann.linked_branch_lengths[e] * f);
  };
  {  // fix_brlen_scalers (:649-706)
    double lowest = ann.brlen_scalers[0], highest = ann.brlen_scalers[0];
    for (double s : ann.brlen_scalers) { if (s < lowest) lowest = s; if (s > highest) highest = s; }
    if (lowest < min_scaler || highest > max_scaler) {
      const double global_scaler = lowest < min_scaler ? min_scaler / lowest : max_scaler / highest;
      for (double &s : ann.brlen_scalers) s *= global_scaler;
      scaleBranchesAll(1.0 / global_scaler);
    }
  }
  double loglh = optimize_onedim(ann, ONEDIM_BRLEN_SCALER, min_scaler, max_scaler, lh_epsilon);
  {  // pllmod_treeinfo_normalize_brlen_scalers (PLLMOD/tree/treeinfo.c:1186-1227)
    double sum_scalers = 0., sum_sites = 0.;
    for (unsigned p = 0; p < P; ++p) {
      const double pat_sites = ann.pattern_weight_sums.at(p);
      sum_sites += pat_sites;
      sum_scalers += ann.brlen_scalers[p] * pat_sites;
    }
    if (ann.parallel_reduce_cb) {
      ann.parallel_reduce_cb(ann.parallel_context, &sum_scalers, 1, 0);
      ann.parallel_reduce_cb(ann.parallel_context, &sum_sites, 1, 0);
    }
    const double mean_rate = sum_scalers / sum_sites;
    scaleBranchesAll(mean_rate);
    for (double &s : ann.brlen_scalers) s /= mean_rate;
  }
  bool brlen_fixed = false;   // fix_brlen_minmax (:708-740)
  for (size_t e = 0; e < E; ++e) {
    const double b = ann.linked_branch_lengths[e];
    if (b < min_brlen) { setLinked(e, min_brlen); brlen_fixed = true; }
    else if (b > max_brlen) { setLinked(e, max_brlen); brlen_fixed = true; }
  }
  if (brlen_fixed) {
    loglh = computeLoglikelihood(ann, 0, 1);
    if (loglh < old_loglh) {
      ann.brlen_scalers = old_scalers;
      for (size_t e = 0; e < E; ++e) setLinked(e, old_brlen[e]);
      loglh = computeLoglikelihood(ann, 0, 1);
    }
  }
  return loglh;
}

/* ---- LH/ComplexityScoring.cpp:7-67 and SRC/optimization/Optimization.cpp:17-214, with optimize_params reduced to its ALPHA
 * and PINV steps (the model steps restated; the product's default optimize_params hook does the same) ---------------------- */
double scoreNetwork(AnnotatedNetwork &ann) {
  const double logl = computeLoglikelihood(ann, 1, 1);
  size_t k = ann.total_num_model_parameters + ann.network.num_reticulations();
  if (ann.options.brlen_linkage == BRLEN_UNLINKED) k += ann.partitionCount() * ann.network.num_branches();
  else {
    k += ann.network.num_branches();
    if (ann.options.brlen_linkage == BRLEN_SCALED) k += ann.partitionCount() - 1;
  }
  const double n = (double)(ann.total_num_sites * ann.network.num_tips);
  const double bic = -2 * logl + (double)k * std::log(n);
  if (bic == std::numeric_limits<double>::infinity()) throw std::runtime_error("Invalid BIC score");
  return bic;
}

double optimize_scalers(AnnotatedNetwork &ann) {  // SRC/optimization/BranchLengthOptimization.cpp:581-599
  const double old_score = scoreNetwork(ann);
  if (ann.options.brlen_linkage == BRLEN_SCALED && ann.partitionCount() > 1) {
    optimize_brlen_scalers(ann, 0.01 /* RAXML_BRLEN_SCALER_MIN */, 100. /* RAXML_BRLEN_SCALER_MAX */, ann.opt.brlen_min, ann.opt.brlen_max, 0.001 /* RAXML_PARAM_EPSILON */);
    return scoreNetwork(ann);
  }
  return old_score;
}

namespace {
void optimizeBranches(AnnotatedNetwork &ann, double brlen_smooth_factor) {  // Optimization.cpp:17-38
  const double old_score = scoreNetwork(ann);
  const int max_iters = (int)(brlen_smooth_factor * 32);
  optimize_branches(ann, max_iters, max_iters, -1, OPT_NEWTON_RAPHSON, false);
  if (scoreNetwork(ann) - old_score > 1E-3) throw std::runtime_error("Complete brlenopt made BIC worse");
  optimize_scalers(ann);  // :37
}
void optimizeModel(AnnotatedNetwork &ann) {  // :72-84
  scoreNetwork(ann);
  optimize_alpha(ann, 0.0201, 100., 0.001);  // PLLMOD_OPT_MIN_ALPHA, PLLMOD_OPT_MAX_ALPHA, RAXML_PARAM_EPSILON
  optimize_pinv(ann, 0.0, 0.99, 0.001);      // PLLMOD_OPT_MIN_PINV, PLLMOD_OPT_MAX_PINV
  scoreNetwork(ann);
}
void optimizeReticulationProbs(AnnotatedNetwork &ann) {  // :91-108
  if (ann.network.num_reticulations() == 0) return;
  const double old_score = scoreNetwork(ann);
  optimize_reticulations(ann, 10);
  if (scoreNetwork(ann) - old_score > 1E-3) throw std::runtime_error("BIC got worse after optimizing reticulation probs");
}
}  // namespace

void optimizeAllNonTopology(AnnotatedNetwork &ann, int type) {  // :118-214
  int max_rounds_slow = 2, act_rounds_slow = 0;
  bool gotBetterSlow = true;
  while (gotBetterSlow) {
    gotBetterSlow = false;
    bool doBrlenOpt = true, doReticulationOpt = true, doModelOpt = true;
    double score_epsilon = 0.01;
    bool gotBetter = true;
    while (gotBetter) {
      gotBetter = false;
      double score_before = scoreNetwork(ann);
      if (doModelOpt) {
        double score_before_model = scoreNetwork(ann);
        optimizeModel(ann);
        double score_after_model = scoreNetwork(ann);
        if (score_before_model - score_after_model > score_epsilon) doModelOpt = false;
      }
      if (doReticulationOpt) {
        double score_before_probs = scoreNetwork(ann);
        optimizeReticulationProbs(ann);
        double score_after_probs = scoreNetwork(ann);
        if (score_before_probs - score_after_probs > score_epsilon) doReticulationOpt = false;
      }
      if (doBrlenOpt) {
        double score_before_branches = scoreNetwork(ann);
        optimizeBranches(ann, 1.0);
        double score_after_branches = scoreNetwork(ann);
        if (score_before_branches - score_after_branches > score_epsilon) doBrlenOpt = false;
      }
      double overall_improv = score_before - scoreNetwork(ann);
      if (overall_improv > score_epsilon && type != 0 && (doBrlenOpt || doReticulationOpt || doModelOpt)) gotBetter = true;
      if (overall_improv > score_epsilon && type == 2) gotBetterSlow = true;
    }
    act_rounds_slow++;
    if (act_rounds_slow >= max_rounds_slow) break;
  }
}

}  // namespace orc
