/* host_internal.hpp — helpers shared by the host translation units (not part of the public API). */
#pragma once
#include "netrax_likelihood_api.hpp"

namespace netrax {
namespace detail {
size_t edgeBetween(const Network &nw, size_t a, size_t b);
size_t activeParent(const Network &nw, size_t n);
std::vector<size_t> activeAliveChildren(const Network &nw, const std::vector<char> &dead, size_t n);
std::vector<size_t> activeNeighbors(const Network &nw, size_t n);
std::vector<char> collect_dead_nodes(const Network &nw, size_t megablobRoot, size_t *displayed_tree_root);
void setReticulationParents(Network &nw, const ReticulationConfig &c);

ReticulationConfigSet getRestrictionsToTakeNeighbor(AnnotatedNetwork &ann, size_t node, size_t neighbor);
ReticulationConfigSet getRestrictionsToDismissNeighbor(AnnotatedNetwork &ann, size_t node, size_t neighbor);
ReticulationConfigSet getTreeConfig(AnnotatedNetwork &ann, size_t tree_idx);
bool isActiveBranch(AnnotatedNetwork &ann, const ReticulationConfigSet &rc, size_t pmatrix_index);
bool isActiveAliveBranch(AnnotatedNetwork &ann, const ReticulationConfigSet &rc, size_t pmatrix_index);
bool clvValidCheck(AnnotatedNetwork &ann, size_t virtual_root, bool care_about_trees = true);
void flushPendingOps(AnnotatedNetwork &ann);
void engineCheck(int ok, const char *what);
uint32_t allocSlot(AnnotatedNetwork &ann);
void releaseSlot(AnnotatedNetwork &ann, uint32_t slot);
nrx_pair makePair(const DisplayedTreeData &a, const DisplayedTreeData &b);
void reduceSum(AnnotatedNetwork &ann, double *data, size_t count);
void reduceHostSum(AnnotatedNetwork &ann, double *data, size_t count);
bool minimizeNewton(double xmin, double *x, double xmax, double tolerance, unsigned max_iters, void (*deriv)(void *, double *, double *, double *), void *ctx);
double minimizeBrent(double xmin, double xguess, double xmax, double xtol, double (*target)(void *, double), void *ctx);
void minimizeBrentMulti(unsigned n, double xmin, double *x, double xmax, double xtol, double (*target)(void *, double *, double *, int *), void *ctx);
/* log(sum_t exp(a_t)) and friends without leaving double range (stands in for mpfr::mpreal) */
double logSumExp(const std::vector<double> &a);
void destroyRerootCache(AnnotatedNetwork &ann);          // host/brlen.cpp
void rerootCacheSize(const AnnotatedNetwork &ann, size_t *entries, size_t *slots);
std::vector<size_t> branchesInPreorder(const AnnotatedNetwork &ann);
bool rerootSessionOpen(const AnnotatedNetwork &ann);
bool rerootSessionIsLazy(const AnnotatedNetwork &ann);
bool lazyRerootPossible(AnnotatedNetwork &ann, size_t pmatrix_index);   // lazy_reroot on, branch active + alive in all trees, plan known
void validateRerootInputs(AnnotatedNetwork &ann, size_t pmatrix_index);  // brings the side children of the branch's re-rooting paths up to date    // between updateCLVsVirtualRerootTrees and finishVirtualReroot
}  // namespace detail
}  // namespace netrax
