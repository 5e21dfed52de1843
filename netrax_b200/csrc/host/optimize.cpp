/*
 * optimize.cpp — the immediate callers of the likelihood path: branch-length
 * optimisation (Newton-Raphson on the sumtable derivatives, or Brent on the edge-rooted / full lnL) and
 * reticulation-probability optimisation (Brent on the re-mixed cached per-tree lnLs).
 *
 * Control flow follows src/optimization/BranchLengthOptimization.cpp and ReticulationOptimization.cpp of the
 * reference (cited per function) and keeps its observable quirks; the two 1-D minimisers are restatements of
 * pll-modules' algorithms (PLLMOD/optimize/opt_algorithms.c), written as small state machines:
 *   newtonMulti  <- pllmod_opt_minimize_newton_multi  (:133-261)
 *   brentSingle  <- pllmod_opt_minimize_brent -> brent_opt_alt / brent_opt_init / brent_opt_post_loop (:859-1254,1404-1429)
 *
 * One deviation and one reproduced quirk:
 *   D1  Brent: pllmod's single-variable wrapper never sets the "all converged" flag, so brent_opt_alt keeps calling
 *       the target with the last proposal until 101 loop iterations have passed.  Those calls do not change the
 *       optimiser state nor (the proposal being unchanged) the network state, so we stop at convergence.
 *   Q7  Newton-Raphson with UNLINKED branch lengths: the reference's derivative callback writes partition_count
 *       values into the solver's arrays of size 1 (BranchLengthOptimization.cpp:190-195; a heap overflow beyond three
 *       partitions) and the solver reads element 0, i.e. PARTITION 0's derivatives whatever partition is being
 *       optimised; the "reload the old length if the lnL got worse" guard (:318-331) then filters the result.  We
 *       reproduce the defined part of that behaviour (element 0) without the out-of-bounds writes.
 */
#include <cmath>
#include <functional>

#include "host_internal.hpp"

namespace netrax {
using namespace detail;

namespace {

/* libpll's PLL_MIN / PLL_MAX (LIBPLL/pll.h:67-68) with their operand order: they differ from std::min / std::max when an
 * operand is NaN (PLL_MIN(NaN, b) = b, std::min(NaN, b) = NaN), and the Newton step below does produce NaN — f = df = 0 on a
 * saturated branch gives dx = -0/0 — which the reference turns into a step of +dxmax instead of a NaN branch length. */
inline double pllMin(double a, double b) { return a < b ? a : b; }
inline double pllMax(double a, double b) { return a > b ? a : b; }

/* ---- pllmod_opt_minimize_newton_multi for xnum functions (opt_algorithms.c:133-261) ---------------------------- */
bool newtonMulti(unsigned xnum, double xmin, double *x, double xmax, double tolerance, unsigned max_iters,
                 const std::function<void(double *x, double *f, double *df)> &deriv) {
  const double dxmax = xmax / max_iters;
  std::vector<double> lo(xnum, xmin), hi(xnum, xmax), f(xnum, 0.0), df(xnum, 0.0);
  std::vector<char> done(xnum, 0);
  for (unsigned i = 0; i < xnum; ++i) x[i] = pllMax(pllMin(x[i], xmax), xmin);
  unsigned iter = 0;
  bool all = false;
  while (!all) {
    if (iter++ > max_iters) return false;  // "Exceeded maximum number of iterations"
    deriv(x, f.data(), df.data());
    all = true;
    for (unsigned i = 0; i < xnum; ++i) {
      if (done[i]) continue;
      if (!std::isfinite(f[i]) || !std::isfinite(df[i])) return false;  // "Wrong likelihood derivatives"
      double dx;
      if (df[i] > 0.0) {
        if (std::fabs(f[i]) < tolerance) { done[i] = 1; continue; }
        if (f[i] < 0.0) lo[i] = x[i]; else hi[i] = x[i];
        dx = -1 * f[i] / df[i];
      } else {
        dx = -1 * f[i] / std::fabs(df[i]);
      }
      dx = pllMax(pllMin(dx, dxmax), -dxmax);
      if (x[i] + dx < lo[i]) dx = lo[i] - x[i];
      if (x[i] + dx > hi[i]) dx = hi[i] - x[i];
      if (std::fabs(dx) < tolerance) { done[i] = 1; continue; }
      x[i] += dx;
      x[i] = pllMax(pllMin(x[i], xmax), xmin);
      all = all && done[i];
    }
  }
  return true;
}

/* ---- Brent's method, single variable, as pll-modules runs it (opt_algorithms.c:859-1254) ------------------------ */
struct BrentState {
  static constexpr double kGold = 0.3819660, kZeps = 1.0e-7;
  static constexpr int kItmax = 100;
  double tol = 0, a = 0, b = 0, d = 0, e = 0, u = 0, v = 0, w = 0, x = 0, fv = 0, fw = 0, fx = 0, startx = 0, fstartx = 0;
  static double sign(double mag, double s) { return s >= 0.0 ? std::fabs(mag) : -std::fabs(mag); }

  /* the part of an iteration before the target is evaluated: convergence test + next proposal u; false = converged */
  bool propose() {
    const double xm = 0.5 * (a + b);
    const double tol1 = tol * std::fabs(x) + kZeps, tol2 = 2.0 * tol1;
    if (std::fabs(x - xm) <= (tol2 - 0.5 * (b - a))) return false;
    if (std::fabs(e) > tol1) {
      double r = (x - w) * (fx - fv);
      double q = (x - v) * (fx - fw);
      double p = (x - v) * q - (x - w) * r;
      q = 2.0 * (q - r);
      if (q > 0.0) p = -p;
      q = std::fabs(q);
      const double etemp = e;
      e = d;
      if (std::fabs(p) >= std::fabs(0.5 * q * etemp) || p <= q * (a - x) || p >= q * (b - x)) {
        e = (x >= xm ? a - x : b - x);
        d = kGold * e;
      } else {
        d = p / q;
        const double t = x + d;
        if (t - a < tol2 || b - t < tol2) d = sign(tol1, xm - x);
      }
    } else {
      e = (x >= xm ? a - x : b - x);
      d = kGold * e;
    }
    u = (std::fabs(d) >= tol1 ? x + d : x + sign(tol1, d));
    return true;
  }

  bool init(double ax, double bx, double cx, double tol_, double fax, double fbx, double fcx) {
    tol = tol_;
    a = (ax < cx ? ax : cx);
    b = (ax > cx ? ax : cx);
    startx = x = bx;
    fstartx = fx = fbx;
    if (fax < fcx) { w = ax; fw = fax; v = cx; fv = fcx; }
    else { w = cx; fw = fcx; v = ax; fv = fax; }
    return propose();
  }

  bool absorb(double fu) {  // the part after the evaluation, then the next proposal
    if (fu <= fx) {
      if (u >= x) a = x; else b = x;
      v = w; w = x; x = u;
      fv = fw; fw = fx; fx = fu;
    } else {
      if (u < x) a = u; else b = u;
      if (fu <= fw || w == x) { v = w; w = u; fv = fw; fw = fu; }
      else if (fu <= fv || v == x || v == w) { v = u; fv = fu; }
    }
    return propose();
  }
};

double brentSingle(double xmin, double xguess, double xmax, double xtol, const std::function<double(double)> &target) {
  if (xguess < xmin) xguess = xmin;
  if (xguess > xmax) xguess = xmax;
  const double eps = xguess > 0 ? xguess * xtol * 50.0 : 2. * xtol;  // bracketing heuristic (:1138-1148)
  double ax = xguess - eps, cx = xguess + eps;
  if (ax < xmin) ax = xmin;
  if (cx > xmax) cx = xmax;
  double fa = target(ax);
  const double fb = target(xguess);
  double fc = target(cx);
  const double fmin = target(xmin), fmax = target(xmax);
  if (fa < fb || fc < fb) { fa = fmin; fc = fmax; ax = xmin; cx = xmax; }
  BrentState s;
  bool running = s.init(ax, xguess, cx, xtol, fa, fb, fc);
  for (int it = 0; running && it <= BrentState::kItmax; ++it) running = s.absorb(target(s.u));  // deviation D1: stop at convergence
  const double xopt = (s.fx > s.fstartx) ? s.startx : s.x;  // if the new score is worse, return the initial value
  target(xopt);
  return xopt;
}

/* pllmod_opt_minimize_brent_multi with global_range = 1 and every variable in the mask -> brent_opt_alt (opt_algorithms.c:
 * 1040-1254): one Brent search per variable, all of them advanced together, one target call per step for ALL variables.
 * target(x, converged, all_converged): converged == nullptr is the reference's target_funk(..., NULL) (bracketing and the final
 * call); otherwise the callee skips converged variables and reports through *all_converged what the reference reads from
 * converged[xnum].  x: in = the guesses, out = xopt (the start value where the search ended worse than it began). */
using BrentMultiTarget = std::function<std::vector<double>(const std::vector<double> &x, const std::vector<char> *converged, bool *all_converged)>;
void brentMulti(size_t n, double min_value, double max_value, double tolerance, std::vector<double> &x, const BrentMultiTarget &target) {
  std::vector<double> xguess = x, ax(n), cx(n), lmin(n, min_value), lmax(n, max_value);
  std::vector<char> converged(n, 0);
  bool all_converged = false;
  for (size_t j = 0; j < n; ++j) {
    if (xguess[j] < lmin[j]) xguess[j] = lmin[j];
    if (xguess[j] > lmax[j]) xguess[j] = lmax[j];
    const double eps = xguess[j] > 0 ? xguess[j] * tolerance * 50.0 : 2. * tolerance;  // bracketing heuristic (:1138-1148)
    ax[j] = xguess[j] - eps;
    if (ax[j] < lmin[j]) ax[j] = lmin[j];
    cx[j] = xguess[j] + eps;
    if (cx[j] > lmax[j]) cx[j] = lmax[j];
  }
  std::vector<double> fa = target(ax, nullptr, nullptr);
  const std::vector<double> fb = target(xguess, nullptr, nullptr);
  std::vector<double> fc = target(cx, nullptr, nullptr);
  const std::vector<double> fmin = target(lmin, nullptr, nullptr), fmax = target(lmax, nullptr, nullptr);
  std::vector<BrentState> st(n);
  for (size_t j = 0; j < n; ++j) {
    if (fa[j] < fb[j] || fc[j] < fb[j]) { fa[j] = fmin[j]; fc[j] = fmax[j]; ax[j] = lmin[j]; cx[j] = lmax[j]; }
    if (!st[j].init(ax[j], xguess[j], cx[j], tolerance, fa[j], fb[j], fc[j])) converged[j] = 1;
  }
  std::vector<double> u(n);
  for (int iter = 0; iter <= BrentState::kItmax; ++iter) {
    for (size_t j = 0; j < n; ++j) u[j] = st[j].u;
    const std::vector<double> fu = target(u, &converged, &all_converged);   // with every variable converged the callee sets nothing but still evaluates, as the reference does
    const bool iterate = !all_converged;
    for (size_t j = 0; j < n; ++j)
      if (!converged[j]) converged[j] = !st[j].absorb(fu[j]);
    if (!iterate) break;
  }
  for (size_t j = 0; j < n; ++j) x[j] = (st[j].fx > st[j].fstartx) ? st[j].startx : st[j].x;  // if the new score is worse, return the initial value
  target(x, nullptr, nullptr);
}

double &brlenRef(AnnotatedNetwork &ann, size_t partition_index, size_t pmatrix_index) {
  FakeTreeinfo &ti = *ann.fake_treeinfo;
  return ti.brlen_linkage == PLLMOD_COMMON_BRLEN_UNLINKED ? ti.branch_lengths[partition_index][pmatrix_index]
                                                          : ti.linked_branch_lengths[pmatrix_index];
}

/* In the reference the linked length is the only storage a linked/scaled analysis reads; our FakeTreeinfo keeps the
 * per-partition copies in sync (what pllmod_treeinfo_set_branch_length does, PLLMOD/tree/treeinfo.c:507-539). */
void storeBrlen(AnnotatedNetwork &ann, size_t partition_index, size_t pmatrix_index, double value) {
  FakeTreeinfo &ti = *ann.fake_treeinfo;
  if (ti.brlen_linkage == PLLMOD_COMMON_BRLEN_UNLINKED) ti.branch_lengths[partition_index][pmatrix_index] = value;
  else {
    ti.linked_branch_lengths[pmatrix_index] = value;
    for (auto &b : ti.branch_lengths) b[pmatrix_index] = value;
  }
}

/* optimize_branch_brent + brent_target_networks (BranchLengthOptimization.cpp:63-156) */
void optimizeBranchBrent(AnnotatedNetwork &ann, std::vector<DisplayedTreeData> &oldTrees, size_t pmatrix_index, size_t partition_index,
                         BrlenOptMethod method) {
  const double old_brlen = brlenRef(ann, partition_index, pmatrix_index);
  auto target = [&](double x) -> double {
    if (brlenRef(ann, partition_index, pmatrix_index) == x)
      return method == BrlenOptMethod::BRENT_REROOT ? -1 * computeLoglikelihoodBrlenOpt(ann, oldTrees, (unsigned)pmatrix_index, 1)
                                                    : -1 * computeLoglikelihood(ann);
    storeBrlen(ann, partition_index, pmatrix_index, x);
    if (method != BrlenOptMethod::BRENT_NORMAL) invalidPmatrixIndexOnly(ann, pmatrix_index);
    else invalidatePmatrixIndex(ann, pmatrix_index);
    return method == BrlenOptMethod::BRENT_REROOT ? -1 * computeLoglikelihoodBrlenOpt(ann, oldTrees, (unsigned)pmatrix_index, 1)
                                                  : -1 * computeLoglikelihood(ann, 1, 1);
  };
  const double new_brlen = brentSingle(ann.options.brlen_min, old_brlen, ann.options.brlen_max, ann.options.tolerance, target);
  storeBrlen(ann, partition_index, pmatrix_index, new_brlen);
  invalidatePmatrixIndex(ann, pmatrix_index);
}

/* optimize_branch_newton_raphson + network_derivative_func_multi (:165-241) */
void optimizeBranchNewtonRaphson(AnnotatedNetwork &ann, std::vector<std::vector<SumtableInfo>> &sumtables, size_t pmatrix_index,
                                 size_t partition_index, unsigned max_iters) {
  const double tolerance = ann.options.brlen_min > 0 ? ann.options.brlen_min / 10.0 : 1.0e-4 /* PLLMOD_OPT_TOL_BRANCH_LEN */;
  double new_brlen = brlenRef(ann, partition_index, pmatrix_index);
  const bool unlinked = ann.fake_treeinfo->brlen_linkage == PLLMOD_COMMON_BRLEN_UNLINKED;
  auto deriv = [&](double *x, double *df, double *ddf) {
    storeBrlen(ann, partition_index, pmatrix_index, x[0]);
    invalidPmatrixIndexOnly(ann, pmatrix_index);
    const LoglDerivatives d = computeLoglikelihoodDerivatives(ann, sumtables, (unsigned)pmatrix_index);
    if (unlinked) { df[0] = d.partition_logl_prime[0]; ddf[0] = d.partition_logl_prime_prime[0]; }  // quirk Q7
    else { df[0] = d.logl_prime; ddf[0] = d.logl_prime_prime; }
  };
  newtonMulti(1, ann.options.brlen_min, &new_brlen, ann.options.brlen_max, tolerance, max_iters, deriv);
  // the reference ignores the solver's status (libpll_reset_error) and keeps whatever length the last iterate stored
}

/* optimize_branch for one partition (:285-343) */
double optimizeBranchPartition(AnnotatedNetwork &ann, std::vector<DisplayedTreeData> &oldTrees, std::vector<std::vector<SumtableInfo>> &sumtables,
                               size_t pmatrix_index, size_t partition_index, BrlenOptMethod method, unsigned max_iters) {
  ann.cached_logl_valid = false;
  auto score = [&] { return method != BrlenOptMethod::BRENT_NORMAL ? computeLoglikelihoodBrlenOpt(ann, oldTrees, (unsigned)pmatrix_index, 1) : computeLoglikelihood(ann); };
  const double start_logl = score();
  if (method == BrlenOptMethod::BRENT_NORMAL || method == BrlenOptMethod::BRENT_REROOT) {
    optimizeBranchBrent(ann, oldTrees, pmatrix_index, partition_index, method);
  } else {
    const double old_brlen = brlenRef(ann, partition_index, pmatrix_index);
    optimizeBranchNewtonRaphson(ann, sumtables, pmatrix_index, partition_index, max_iters);
    const double new_logl = computeLoglikelihoodBrlenOpt(ann, oldTrees, (unsigned)pmatrix_index, 1);
    if (new_logl < start_logl) {  // NR did not converge: reload the old branch length
      storeBrlen(ann, partition_index, pmatrix_index, old_brlen);
      invalidPmatrixIndexOnly(ann, pmatrix_index);
    }
  }
  return score();
}

}  // namespace

/* the two single-variable minimisers above behind plain callbacks (device-free; nrxh_minimize_newton / nrxh_minimize_brent) */
namespace detail {
bool minimizeNewton(double xmin, double *x, double xmax, double tolerance, unsigned max_iters, void (*deriv)(void *, double *, double *, double *), void *ctx) {
  return newtonMulti(1, xmin, x, xmax, tolerance, max_iters, [&](double *xx, double *f, double *df) { deriv(ctx, xx, f, df); });
}
double minimizeBrent(double xmin, double xguess, double xmax, double xtol, double (*target)(void *, double), void *ctx) {
  return brentSingle(xmin, xguess, xmax, xtol, [&](double x) { return target(ctx, x); });
}
/* target: pll-modules' callback contract (algo_callback.c:295-363) — target(ctx, x, fx, converged) with converged == NULL for the
 * plain calls and converged[n] = "all converged" written by the callee otherwise */
void minimizeBrentMulti(unsigned n, double xmin, double *x, double xmax, double xtol, double (*target)(void *, double *, double *, int *), void *ctx) {
  std::vector<double> v(x, x + n);
  brentMulti(n, xmin, xmax, xtol, v, [&](const std::vector<double> &xs, const std::vector<char> *converged, bool *all_converged) {
    std::vector<double> in(xs), fx(n, 0.0);
    if (!converged) { target(ctx, in.data(), fx.data(), nullptr); return fx; }
    std::vector<int> flags(n + 1, 0);
    for (unsigned j = 0; j < n; ++j) flags[j] = (*converged)[j];
    target(ctx, in.data(), fx.data(), flags.data());
    *all_converged = flags[n] != 0;
    return fx;
  });
  std::copy(v.begin(), v.end(), x);
}
}  // namespace detail

/* optimize_branch (BranchLengthOptimization.cpp:345-421) */
double optimize_branch(AnnotatedNetwork &ann, size_t pmatrix_index, BrlenOptMethod method, unsigned int max_iters) {
  if (pmatrix_index >= ann.network.num_branches()) throw std::runtime_error("optimize_branch: pmatrix index out of range");
  // lazy re-rooting (opt-in, host/brlen.cpp): no evaluation from the root before and after a branch that is active and alive in every tree
  bool lazy = method != BrlenOptMethod::BRENT_NORMAL && !ann.root_clvs_stale && detail::lazyRerootPossible(ann, pmatrix_index);
  // (lazy: the reference's re-rooting sanity check below compares against the lnL from the root, which is not evaluated — skipped)
  const double old_logl = lazy ? -std::numeric_limits<double>::infinity() : computeLoglikelihood(ann);
  std::vector<DisplayedTreeData> oldTrees;
  std::vector<std::vector<SumtableInfo>> sumtables;
  if (method != BrlenOptMethod::BRENT_NORMAL) {  // step 1: the virtual re-rooting
    if (lazy) detail::validateRerootInputs(ann, pmatrix_index);
    else oldTrees = extractOldTrees(ann, ann.network.root);
    Node *new_virtual_root = &ann.network.nodes[ann.network.edges[pmatrix_index].source];
    Node *new_virtual_root_back = &ann.network.nodes[ann.network.edges[pmatrix_index].target];
    ReticulationConfigSet restrictions = getRestrictionsActiveAliveBranch(ann, pmatrix_index);
    updateCLVsVirtualRerootTrees(ann, ann.network.root, new_virtual_root, new_virtual_root_back, restrictions);
    ann.cached_logl_valid = false;
    // Newton-Raphson: the sumtables are made in the same pass over the pairs' CLVs as the edge-rooted lnL (the reference calls
    // computeLoglikelihoodBrlenOpt and computePartitionSumtables back to back, :374-381)
    const bool nr = method == BrlenOptMethod::NEWTON_RAPHSON;
    auto edge_logl = [&] { return nr ? computeLoglikelihoodBrlenOptAndSumtables(ann, oldTrees, (unsigned)pmatrix_index, sumtables)
                                     : computeLoglikelihoodBrlenOpt(ann, oldTrees, (unsigned)pmatrix_index); };
    double brlenopt_logl;
    try { brlenopt_logl = edge_logl(); }
    catch (const LazyRerootNeedsRoot &) {   // a displayed tree needs the old root's per-tree lnL after all: redo this branch from an evaluated root
      redoRerootFromRoot(ann, (unsigned)pmatrix_index, oldTrees);
      lazy = false;
      brlenopt_logl = edge_logl();
    }
    if (old_logl - brlenopt_logl >= 1E-3)  // the reference's `fabs(old_logl - brlenopt_logl >= 1E-3)` (:377): one-sided
      throw std::runtime_error("Something went wrong when rerooting CLVs during brlen optimization");
    // (sumtables: made above together with brlenopt_logl)
  }
  if (ann.fake_treeinfo->brlen_linkage == PLLMOD_COMMON_BRLEN_UNLINKED) {
    for (size_t p = 0; p < ann.fake_treeinfo->partition_count; ++p)
      optimizeBranchPartition(ann, oldTrees, sumtables, pmatrix_index, p, method, max_iters);
  } else {
    optimizeBranchPartition(ann, oldTrees, sumtables, pmatrix_index, 0, method, max_iters);
  }
  if (lazy) {
    ann.cached_logl_valid = false;
    const double l = computeLoglikelihoodBrlenOpt(ann, oldTrees, (unsigned)pmatrix_index, 1);
    finishVirtualReroot(ann);
    return ann.lazy_last_logl = l;
  }
  if (method != BrlenOptMethod::BRENT_NORMAL) finishVirtualReroot(ann);  // restore the network root: invalidatePmatrixIndex only if the length changed
  return ann.lazy_last_logl = computeLoglikelihood(ann);
}

/* optimize_branches_internal (:423-476); `radius` is unused there as well (the neighbour re-queueing is commented out).
 * The candidate set is a std::unordered_set<size_t> visited from begin(), exactly as in the reference, so the visiting
 * order is libstdc++'s — the same library a NetRAX build on this machine uses. */
double optimize_branches(AnnotatedNetwork &ann, int max_iters, int max_iters_outside, int radius, std::unordered_set<size_t> candidates,
                         bool restricted_total_iters) {
  (void)radius;
  for (size_t idx : candidates)
    if (idx >= ann.network.num_branches()) throw std::runtime_error("optimize_branches: candidate pmatrix index out of range");
  double old_logl = computeLoglikelihood(ann, 1, 1);
  const double start_logl = old_logl;
  std::vector<size_t> act_iters(ann.network.num_branches(), 0);
  const BrlenOptMethod method = ann.options.brlenOptMethod;
  size_t total_iters = 0;
  while (!candidates.empty()) {
    const size_t pmatrix_index = *candidates.begin();
    candidates.erase(candidates.begin());
    total_iters++;
    if (restricted_total_iters && total_iters >= (size_t)max_iters_outside) continue;
    if (act_iters[pmatrix_index] >= (size_t)max_iters_outside) continue;
    act_iters[pmatrix_index]++;
    old_logl = optimize_branch(ann, pmatrix_index, method, (unsigned)max_iters);
  }
  if (old_logl < start_logl && std::fabs(old_logl - start_logl) >= 1E-3) throw std::runtime_error("Overall loglikelihood got worse");
  return old_logl;
}

double optimize_branches(AnnotatedNetwork &ann, int max_iters, int max_iters_outside, int radius, bool restricted_total_iters) {  // :567-576
  std::unordered_set<size_t> candidates;
  for (size_t i = 0; i < ann.network.num_branches(); ++i) candidates.emplace(i);
  return optimize_branches(ann, max_iters, max_iters_outside, radius, candidates, restricted_total_iters);
}

/* optimize_reticulation (ReticulationOptimization.cpp:68-100): every Brent evaluation is setReticulationProb +
 * computeLoglikelihood(1, 1) — no CLV is invalid, so it re-mixes the cached per-tree lnLs on the host (row f2: 0 launches). */
double optimize_reticulation(AnnotatedNetwork &ann, size_t reticulation_index) {
  if (reticulation_index >= ann.network.num_reticulations()) throw std::runtime_error("optimize_reticulation: index out of range");
  computeLoglikelihood(ann, 1, 1);
  const double old_brprob = ann.reticulation_probs[reticulation_index];
  setReticulationProb(ann, reticulation_index, 0.5);  // "just for debug" in the reference; kept, it is part of the call sequence
  const double new_brprob = brentSingle(ann.options.brprob_min, old_brprob, ann.options.brprob_max, ann.options.tolerance, [&](double x) {
    setReticulationProb(ann, reticulation_index, x);
    return -1 * computeLoglikelihood(ann, 1, 1);
  });
  setReticulationProb(ann, reticulation_index, new_brprob);
  return computeLoglikelihood(ann, 1, 1);
}

double optimize_reticulations(AnnotatedNetwork &ann, int max_iters) {  // :102-117
  double act_logl = computeLoglikelihood(ann, 1, 1);
  for (int act_iters = 0; act_iters < max_iters;) {
    double loop_logl = act_logl;
    for (size_t i = 0; i < ann.network.num_reticulations(); ++i) loop_logl = optimize_reticulation(ann, i);
    act_iters++;
    if (loop_logl == act_logl) break;
    act_logl = loop_logl;
  }
  return act_logl;
}

/* ---- model-parameter loop: the ALPHA step of optimize_params (src/optimization/ModelOptimization.cpp:56-65)
 * = pllmod_algo_opt_onedim_treeinfo(PLLMOD_OPT_PARAM_ALPHA) (PLLMOD/algorithm/pllmod_algorithm.c:743-866): Brent for the
 * alphas of ALL partitions at once (brent_opt_alt, opt_algorithms.c:1043-1254, through BrentState above), each iterate =
 * new Gamma rates per unconverged partition (treeinfo_set_alpha, :566-587) + ONE full re-evaluation
 * (target_func_onedim_treeinfo, algo_callback.c:295-363) whose per-partition lnLs drive the per-partition Brent states.
 * On the device an iterate is: a < 1 KB model upload per partition, P-matrices of every edge (K1), the cached plan
 * replayed as one CUDA graph (K2 + fused K3) — no eigendecomposition, the rate matrix has not changed. */
void setAlpha(AnnotatedNetwork &ann, unsigned p, double alpha) {
  PartitionModel &m = ann.fake_treeinfo->partitions.at(p);
  std::vector<double> rates(m.rate_cats);
  if (!compute_gamma_cats(alpha, m.rate_cats, rates.data(), m.gamma_mode)) throw std::runtime_error("Invalid alpha value / GAMMA discretization mode");
  m.alpha = alpha;
  m.rates = rates;
  pushPartitionModel(ann, p);
}

void setPinv(AnnotatedNetwork &ann, unsigned p, double prop_invar) {
  if (!(prop_invar >= 0.0 && prop_invar < 1.0)) throw std::runtime_error("Invalid proportion of invariant sites (" + std::to_string(prop_invar) + ")");
  ann.fake_treeinfo->partitions.at(p).prop_invar = prop_invar;
  pushPartitionModel(ann, p);   // P-matrices depend on rates / (1 - pinv): everything of the partition is stale
}

/* pllmod_algo_opt_onedim_treeinfo (PLLMOD/algorithm/pllmod_algorithm.c:743-866): ONE Brent search per partition, all of them
 * advanced together, for the parameter kinds pll-modules supports there: ALPHA (treeinfo_set_alpha :566-587), PINV
 * (treeinfo_set_pinv :601-622) and BRANCH_LEN_SCALER (treeinfo_set_brlen_scaler :636-647). */
enum OnedimParam { ONEDIM_ALPHA, ONEDIM_PINV, ONEDIM_BRLEN_SCALER };

static double optimize_onedim(AnnotatedNetwork &ann, OnedimParam param, double min_value, double max_value, double tolerance) {
  FakeTreeinfo &ti = *ann.fake_treeinfo;
  std::vector<unsigned> parts;   // params_to_optimize[p] & param
  for (unsigned p = 0; p < ti.partition_count; ++p) {
    const PartitionModel &m = ti.partitions[p];
    bool on = true;   // params_to_optimize[p] & param (PLLMOD/algorithm/pllmod_algorithm.c:765-772)
    if (param == ONEDIM_ALPHA) on = m.params_to_optimize >= 0 ? (m.params_to_optimize & 1) != 0 && m.alpha > 0.0 : m.alpha > 0.0;
    else if (param == ONEDIM_PINV) on = m.params_to_optimize >= 0 ? (m.params_to_optimize & 2) != 0 : m.prop_invar > 0.0;
    if (on) parts.push_back(p);
  }
  if (param == ONEDIM_BRLEN_SCALER && ti.brlen_scalers.size() < ti.partition_count) ti.brlen_scalers.resize(ti.partition_count, 1.0);
  auto get = [&](unsigned p) { return param == ONEDIM_ALPHA ? ti.partitions[p].alpha : param == ONEDIM_PINV ? ti.partitions[p].prop_invar : ti.brlen_scalers[p]; };
  auto set = [&](unsigned p, double x) {
    if (param == ONEDIM_ALPHA) setAlpha(ann, p, x);
    else if (param == ONEDIM_PINV) setPinv(ann, p, x);
    else set_brlen_scaler(ann, p, x);
  };
  const size_t n = parts.size();
  if (n) {
    std::vector<double> x(n);
    for (size_t j = 0; j < n; ++j) x[j] = get(parts[j]);
    // target_func_onedim_treeinfo (algo_callback.c:295-363): set the unconverged partitions' parameters, one full evaluation,
    // per-partition scores
    brentMulti(n, min_value, max_value, tolerance, x, [&](const std::vector<double> &v, const std::vector<char> *converged, bool *all_converged) {
      double unconverged = 0.0;
      for (size_t j = 0; j < n; ++j) {
        if (converged && (*converged)[j]) continue;
        unconverged = 1.0;
        set(parts[j], v[j]);
      }
      computeLoglikelihood(ann, 0, 1);
      std::vector<double> fx(n);
      for (size_t j = 0; j < n; ++j) fx[j] = -1 * ti.partition_loglh[parts[j]];
      if (converged) {
        // every shard holds a slice of EVERY partition and the reduced per-partition lnLs, so the flags already agree under an
        // NCCL communicator; the callback is honoured because the reference calls it here
        if (ti.parallel_reduce_cb) ti.parallel_reduce_cb(ti.parallel_context, &unconverged, 1, PLLMOD_COMMON_REDUCE_SUM);
        *all_converged = !(unconverged > 0.0);
      }
      return fx;
    });
  }
  return computeLoglikelihood(ann, 0, 1);
}

double optimize_alpha(AnnotatedNetwork &ann, double min_alpha, double max_alpha, double tolerance) {
  return optimize_onedim(ann, ONEDIM_ALPHA, min_alpha, max_alpha, tolerance);
}

/* the PINV step of optimize_params (src/optimization/ModelOptimization.cpp:67-76): partitions with +I */
double optimize_pinv(AnnotatedNetwork &ann, double min_pinv, double max_pinv, double tolerance) {
  return optimize_onedim(ann, ONEDIM_PINV, min_pinv, max_pinv, tolerance);
}

/* pllmod_algo_opt_brlen_scalers_treeinfo (PLLMOD/algorithm/pllmod_algorithm.c:869-960) on the fake treeinfo (no subnodes: the
 * `else` branches that work on branch_lengths[0]).  Returns the log-likelihood. */
double optimize_brlen_scalers(AnnotatedNetwork &ann, double min_scaler, double max_scaler, double min_brlen, double max_brlen, double lh_epsilon) {
  FakeTreeinfo &ti = *ann.fake_treeinfo;
  if (ti.brlen_linkage != PLLMOD_COMMON_BRLEN_SCALED) throw std::runtime_error("Branch length scaler optimization works only in scaled branch length mode.");
  const unsigned P = ti.partition_count;
  const size_t E = ann.network.num_branches();
  if (ti.brlen_scalers.size() < P) ti.brlen_scalers.resize(P, 1.0);
  const double old_loglh = computeLoglikelihood(ann, 0, 1);
  const std::vector<double> old_scalers = ti.brlen_scalers, old_brlen = ti.linked_branch_lengths;
  auto scale_branches_all = [&](double f) {   // pllmod_treeinfo_scale_branches_all (PLLMOD/tree/treeinfo.c:1132-1155): no invalidation
    for (size_t e = 0; e < E; ++e) { ti.linked_branch_lengths[e] *= f; for (auto &b : ti.branch_lengths) b[e] = ti.linked_branch_lengths[e]; }
  };
  {  // fix_brlen_scalers (:649-706): force the scalers into [min, max], branches scaled by the inverse
    double lo = ti.brlen_scalers[0], hi = ti.brlen_scalers[0];
    for (unsigned p = 0; p < P; ++p) { lo = std::min(lo, ti.brlen_scalers[p]); hi = std::max(hi, ti.brlen_scalers[p]); }
    if (lo < min_scaler || hi > max_scaler) {
      const double g = lo < min_scaler ? min_scaler / lo : max_scaler / hi;
      for (unsigned p = 0; p < P; ++p) ti.brlen_scalers[p] *= g;
      scale_branches_all(1.0 / g);
    }
  }
  double loglh = optimize_onedim(ann, ONEDIM_BRLEN_SCALER, min_scaler, max_scaler, lh_epsilon);
  {  // pllmod_treeinfo_normalize_brlen_scalers (PLLMOD/tree/treeinfo.c:1186-1227): site-weighted mean scaler = 1
    double sum_scalers = 0., sum_sites = 0.;
    for (unsigned p = 0; p < P; ++p) {
      const double pat_sites = ann.pattern_weight_sum(p);   // this shard's pattern_weight_sum
      sum_sites += pat_sites;
      sum_scalers += ti.brlen_scalers[p] * pat_sites;
    }
    double sums[2] = {sum_scalers, sum_sites};   // the reference reduces them one after the other; one message here
    reduceHostSum(ann, sums, 2);
    sum_scalers = sums[0]; sum_sites = sums[1];
    const double mean_rate = sum_scalers / sum_sites;
    scale_branches_all(mean_rate);
    for (unsigned p = 0; p < P; ++p) ti.brlen_scalers[p] /= mean_rate;
  }
  bool brlen_fixed = false;   // fix_brlen_minmax (:708-740)
  for (size_t e = 0; e < E; ++e) {
    double &b = ti.linked_branch_lengths[e];
    if (b < min_brlen) { b = min_brlen; brlen_fixed = true; }
    else if (b > max_brlen) { b = max_brlen; brlen_fixed = true; }
    for (auto &pb : ti.branch_lengths) pb[e] = b;
  }
  if (brlen_fixed) {
    loglh = computeLoglikelihood(ann, 0, 1);
    if (loglh < old_loglh) {   // revert optimization and restore old values
      ti.brlen_scalers = old_scalers;
      ti.linked_branch_lengths = old_brlen;
      for (auto &pb : ti.branch_lengths) pb = old_brlen;
      loglh = computeLoglikelihood(ann, 0, 1);
    }
  }
  return loglh;
}

/* optimize_scalers (src/optimization/BranchLengthOptimization.cpp:581-599) */
double optimize_scalers(AnnotatedNetwork &ann, bool) {
  const double old_score = scoreNetwork(ann);
  if (ann.options.brlen_linkage == PLLMOD_COMMON_BRLEN_SCALED && ann.fake_treeinfo->partition_count > 1) {
    optimize_brlen_scalers(ann, 0.01 /* RAXML_BRLEN_SCALER_MIN */, 100. /* RAXML_BRLEN_SCALER_MAX */, ann.options.brlen_min, ann.options.brlen_max,
                           0.001 /* RAXML_PARAM_EPSILON */);
    return scoreNetwork(ann);
  }
  return old_score;
}

/* ---- src/likelihood/ComplexityScoring.cpp:7-67 --------------------------------------------------------------------- */
static double aic_(double logl, double k) { return -2 * logl + 2 * k; }
static double aicc_(double logl, double k, double n) { return aic_(logl, k) + (2 * k * k + 2 * k) / (n - k - 1); }
static double bic_(double logl, double k, double n) { return -2 * logl + k * std::log(n); }

size_t get_param_count(AnnotatedNetwork &ann) {  // :13-34
  size_t param_count = ann.total_num_model_parameters;
  param_count += ann.network.num_reticulations();  // reticulation probs as free parameters
  if (ann.fake_treeinfo->brlen_linkage == PLLMOD_COMMON_BRLEN_UNLINKED) param_count += ann.fake_treeinfo->partition_count * ann.network.num_branches();
  else {
    param_count += ann.network.num_branches();
    if (ann.fake_treeinfo->brlen_linkage == PLLMOD_COMMON_BRLEN_SCALED) param_count += ann.fake_treeinfo->partition_count - 1;
  }
  return param_count;
}
size_t get_sample_size(AnnotatedNetwork &ann) { return ann.total_num_sites * ann.network.num_tips(); }  // :36-38
double aic(AnnotatedNetwork &ann, double logl) { return aic_(logl, (double)get_param_count(ann)); }
double aicc(AnnotatedNetwork &ann, double logl) { return aicc_(logl, (double)get_param_count(ann), (double)get_sample_size(ann)); }
double bic(AnnotatedNetwork &ann, double logl) { return bic_(logl, (double)get_param_count(ann), (double)get_sample_size(ann)); }

double scoreNetwork(AnnotatedNetwork &ann) {  // :57-67
  const double logl = computeLoglikelihood(ann, 1, 1);
  const double bic_score = bic(ann, logl);
  if (bic_score == std::numeric_limits<double>::infinity()) throw std::runtime_error("Invalid BIC score");
  return bic_score;
}

double scoreNetworkPseudo(AnnotatedNetwork &ann) {  // :69-79
  const double logl = computePseudoLoglikelihood(ann, 1, 1);
  const double bic_score = bic(ann, logl);
  if (bic_score == std::numeric_limits<double>::infinity()) throw std::runtime_error("Invalid BIC score");
  return bic_score;
}

double network_logl_wrapper(void *network_params, int incremental, int update_pmatrices, double **) {  // RaxmlWrapper.cpp:21-26
  return computeLoglikelihood(*static_cast<NetworkParams *>(network_params)->ann_network, incremental, update_pmatrices);
}

/* ---- src/optimization/Optimization.cpp:17-216 ------------------------------------------------------------------------ */
void optimizeBranches(AnnotatedNetwork &ann, double brlen_smooth_factor, bool, bool restricted_total_iters) {  // :17-38
  const double old_score = scoreNetwork(ann);
  const int max_iters = (int)(brlen_smooth_factor * 32);  // RAXML_BRLEN_SMOOTHINGS
  optimize_branches(ann, max_iters, max_iters, -1 /* PLLMOD_OPT_BRLEN_OPTIMIZE_ALL */, restricted_total_iters);
  const double new_score = scoreNetwork(ann);
  if (new_score - old_score > 1E-3) throw std::runtime_error("Complete brlenopt made BIC worse");
  optimize_scalers(ann, true);  // :37 (a no-op unless the linkage is scaled and there are several partitions)
}

void optimizeModel(AnnotatedNetwork &ann, bool) {  // :72-84
  scoreNetwork(ann);
  if (ann.optimize_params_cb) ann.optimize_params_cb(ann);
  else { optimize_alpha(ann); optimize_pinv(ann); }   // the ALPHA and PINV steps of optimize_params, in its order
  scoreNetwork(ann);
}

void optimizeReticulationProbs(AnnotatedNetwork &ann, bool) {  // :91-108
  if (ann.network.num_reticulations() == 0) return;
  const double old_score = scoreNetwork(ann);
  optimize_reticulations(ann, 10);
  const double new_score = scoreNetwork(ann);
  if (new_score - old_score > 1E-3) throw std::runtime_error("BIC got worse after optimizing reticulation probs");
}

void optimizeAllNonTopology(AnnotatedNetwork &ann, OptimizeAllNonTopologyType type, bool silent) {  // :118-214
  const int max_rounds_slow = 2;
  int act_rounds_slow = 0;
  bool gotBetterSlow = true;
  while (gotBetterSlow) {
    gotBetterSlow = false;
    bool doBrlenOpt = true, doReticulationOpt = true, doModelOpt = true;
    const double score_epsilon = 0.01;
    bool gotBetter = true;
    while (gotBetter) {
      gotBetter = false;
      const double score_before = scoreNetwork(ann);
      if (doModelOpt) {
        const double before = scoreNetwork(ann);
        optimizeModel(ann, silent);
        if (before - scoreNetwork(ann) > score_epsilon) doModelOpt = false;   // as written in the reference (:161-164)
      }
      if (doReticulationOpt) {
        const double before = scoreNetwork(ann);
        optimizeReticulationProbs(ann, silent);
        if (before - scoreNetwork(ann) > score_epsilon) doReticulationOpt = false;
      }
      if (doBrlenOpt) {
        const double before = scoreNetwork(ann);
        optimizeBranches(ann, 1.0, silent);
        if (before - scoreNetwork(ann) > score_epsilon) doBrlenOpt = false;
      }
      const double overall_improv = score_before - scoreNetwork(ann);
      if (overall_improv > score_epsilon && type != OptimizeAllNonTopologyType::QUICK && (doBrlenOpt || doReticulationOpt || doModelOpt)) gotBetter = true;
      if (overall_improv > score_epsilon && type == OptimizeAllNonTopologyType::SLOW) gotBetterSlow = true;
    }
    if (++act_rounds_slow >= max_rounds_slow) break;
  }
}

}  // namespace netrax
