/*
 * netrax_b200.h — host-level C-ABI of the B200 network-likelihood engine: NetRAX's likelihood API
 * (netrax_b200/csrc/host/netrax_likelihood_api.hpp, same names as the reference) flattened to plain C so that
 * any FFI (ctypes in this repo's tests / bench; a C++ NetRAX build links the C++ header directly) can drive it.
 * Every function returns 1 on success, 0 on failure (message in nrxh_last_error(), the text of the
 * std::runtime_error the reference would have thrown).
 *
 *   nrxh_new / nrxh_set_network  <- the Network inside an AnnotatedNetwork (src/graph/Network.hpp:20-84, numbering of
 *                                   convertNetworkToplevel, src/io/NetworkIO.cpp:59-330) + build_annotated_network
 *                                   (src/graph/AnnotatedNetwork.cpp:199-213)
 *   nrxh_add_partition           <- create_pll_partition's inputs (libs/raxml-ng/src/TreeInfo.cpp:629-707, called at src/RaxmlWrapper.cpp:195,252): states, rate categories,
 *                                   tip states, pattern weights, frequencies, exchangeabilities, category rates / weights
 *   nrxh_set_options             <- NetraxOptions::likelihood_variant / brlen_linkage (src/NetraxOptions.hpp:36,98)
 *   nrxh_set_partition_brlens    <- fake_treeinfo->branch_lengths[p] (createNetworkPllTreeinfoInternal, src/RaxmlWrapper.cpp:539-669)
 *   nrxh_init                    <- init_annotated_network (src/graph/AnnotatedNetwork.cpp:80-185) + createNetworkPllTreeinfo (:671)
 *   nrxh_read_clv / _scaler, nrxh_tree_info / _config, nrxh_num_trees
 *                                <- pernode_displayed_tree_data[node].displayed_trees[t]: clv_vector, scale_buffer,
 *                                   treeLoglData (src/graph/DisplayedTreeData.hpp:18-44, TreeLoglData.hpp) — what the tests compare
 *   nrxh_partition_loglh         <- fake_treeinfo->partition_loglh (src/likelihood/LikelihoodComputation.cpp: evaluateTrees)
 *   nrxh_persite_lnl             <- the persite_lnl output of pll_compute_root_loglikelihood (LIBPLL/likelihood.c:30-120)
 *   nrxh_compute_loglikelihood   <- netrax::computeLoglikelihood            src/likelihood/LikelihoodComputation.hpp:17
 *   nrxh_brlen_prepare           <- extractOldTrees + getRestrictionsActiveAliveBranch + updateCLVsVirtualRerootTrees
 *                                   (optimize_branch step 1, src/optimization/BranchLengthOptimization.cpp:355-373)
 *   nrxh_brlen_logl              <- netrax::computeLoglikelihoodBrlenOpt    src/likelihood/VirtualRerooting.hpp:8
 *   nrxh_brlen_sumtables         <- netrax::computePartitionSumtables       src/likelihood/LikelihoodDerivatives.hpp:86
 *   nrxh_brlen_set_length        <- network_derivative_func_multi's proposal step (BranchLengthOptimization.cpp:176-187)
 *   nrxh_brlen_derivatives       <- netrax::computeLoglikelihoodDerivatives src/likelihood/LikelihoodDerivatives.hpp:82
 *   nrxh_brlen_finish            <- invalidatePmatrixIndex + computeLoglikelihood (BranchLengthOptimization.cpp:413-419); the invalidation
 *                                   happens only when the branch length differs from the one nrxh_brlen_prepare found (nothing was overwritten)
 *   nrxh_set_branch_length / nrxh_set_reticulation_prob / nrxh_set_model
 *                                <- what optimize_branch / setReticulationProb / pll_set_* + invalidate do to the state
 *   nrxh_set_reduce_callback     <- fake_treeinfo->parallel_reduce_cb        src/RaxmlWrapper.cpp:717-718
 *   nrxh_comm_init               <- ParallelContext::init_mpi + mpi_allreduce libs/raxml-ng/src/ParallelContext.cpp:56-70,452-487
 */
#ifndef NETRAX_B200_H
#define NETRAX_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void (*nrxh_reduce_cb)(void *context, double *data, size_t count, int op /* 0 = SUM */);

const char *nrxh_last_error(void);
void *nrxh_new(const char *options /* "device=N;plan_cache=0|1" or NULL */);
void nrxh_free(void *h);
int nrxh_set_network(void *h, unsigned num_tips, unsigned num_nodes, unsigned root, unsigned num_edges, const unsigned *edge_source,
                     const unsigned *edge_target, const double *edge_length, const double *edge_prob, unsigned num_reticulations,
                     const unsigned *ret_node, const unsigned *ret_first_edge, const unsigned *ret_second_edge);
int nrxh_add_partition(void *h, unsigned states, unsigned rate_cats, unsigned sites, const uint32_t *tip_masks,
                       const unsigned *pattern_weights, const double *freqs, const double *subst_params, const double *rates,
                       const double *rate_weights);
int nrxh_set_options(void *h, int likelihood_variant /* 0 AVERAGE, 1 BEST, 2 SARAH_PSEUDO */, int brlen_linkage /* PLLMOD_COMMON_BRLEN_*: 0 linked, 1 scaled, 2 unlinked */);
int nrxh_set_partition_brlens(void *h, unsigned p, const double *brlens);
int nrxh_set_reduce_callback(void *h, nrxh_reduce_cb cb, void *context);
int nrxh_init(void *h);
/* site-shard communicator (after nrxh_init): NCCL all-reduce of the per-tree / per-pair partition sums inside the
 * engine, the role of ParallelContext::parallel_reduce_cb + MPI_Allreduce (libs/raxml-ng/src/ParallelContext.cpp:425-487).
 * With a communicator attached the reduce callback is not called. */
int nrxh_comm_get_unique_id(uint8_t *id128);
int nrxh_comm_init(void *h, const uint8_t *id128, int rank, int nranks);
int nrxh_compute_loglikelihood(void *h, int incremental, int update_pmatrices, double *out);
/* batched scoring of n networks (n handles, one engine / CUDA stream each): all evaluations are enqueued, then collected;
 * out[i] == nrxh_compute_loglikelihood(handles[i], ...) (reference: the sequential candidate loop of src/search/Filtering.cpp:210-260) */
int nrxh_compute_loglikelihood_batch(void **handles, unsigned n, int incremental, int update_pmatrices, double *out);
unsigned nrxh_num_partitions(void *h);
unsigned nrxh_root(void *h);
unsigned nrxh_num_nodes(void *h);
int nrxh_num_trees(void *h, unsigned node);
int nrxh_tree_config(void *h, unsigned node, unsigned tree, char *buf, unsigned buflen);
int nrxh_tree_info(void *h, unsigned node, unsigned tree, double *logprob, double *partition_logl, int *flags);
int nrxh_read_clv(void *h, unsigned node, unsigned tree, unsigned p, double *out);
int nrxh_read_scaler(void *h, unsigned node, unsigned tree, unsigned p, unsigned *out);
int nrxh_partition_loglh(void *h, double *out);
int nrxh_set_branch_length(void *h, int partition, unsigned edge, double value);
int nrxh_set_reticulation_prob(void *h, unsigned r, double prob);
/* pllmod_treeinfo_t::params_to_optimize[p] for the one-dimensional model optimisers: bit 0 alpha, bit 1 pinv (so that a +I partition
 * whose proportion is 0 stays a free parameter); -1 = derive from the current values */
int nrxh_set_params_to_optimize(void *h, unsigned p, int mask);
int nrxh_set_model(void *h, unsigned p, const double *freqs, const double *subst_params, const double *rates, const double *rate_weights);
/* Scaled branch-length linkage (brlen_linkage = 1): pllmod_treeinfo_t::brlen_scalers[p]; the P-matrices of partition p use
 * scaler x the linked branch length (PLLMOD/tree/treeinfo.c:862-864).  Fails unless the linkage is scaled. */
int nrxh_set_brlen_scaler(void *h, unsigned p, double scaler);
int nrxh_get_brlen_scalers(void *h, double *out /* [partitions] */);
/* optimize_scalers (src/optimization/BranchLengthOptimization.cpp:581-599) = pllmod_algo_opt_brlen_scalers_treeinfo with raxml-ng's
 * bounds: Brent over all partitions' scalers at once, scalers normalised to a site-weighted mean of 1; returns the BIC */
int nrxh_optimize_scalers(void *h, double *bic_score);
/* the PINV step of optimize_params (src/optimization/ModelOptimization.cpp:67-76) over the partitions with +I */
int nrxh_optimize_pinv(void *h, double min_pinv, double max_pinv, double tolerance, double *final_logl);
int nrxh_get_pinv(void *h, unsigned p, double *prop_invar);
/* Mixture with one rate matrix per rate category (LG4M / LG4X): what raxml-ng's Model::ratecat_submodels() is to NetRAX
 * (src/RaxmlWrapper.cpp:199-203 -> pllmod param_indices -> every libpll call).  n rate matrices, freqs [n][states],
 * subst [n][states (states - 1) / 2], category c uses matrix ratecat_submodels[c]; n == 1 goes back to a single matrix. */
int nrxh_set_submodels(void *h, unsigned p, unsigned n, const unsigned *ratecat_submodels, const double *freqs, const double *subst);
int nrxh_get_eigen(void *h, unsigned p, double *eigenvecs, double *inv_eigenvecs, double *eigenvals);
int nrxh_set_eigen(void *h, unsigned p, const double *eigenvecs, const double *inv_eigenvecs, const double *eigenvals);
int nrxh_get_pmatrix(void *h, unsigned p, unsigned edge, double *out);
int nrxh_brlen_prepare(void *h, unsigned edge, double *old_logl);
int nrxh_brlen_logl(void *h, unsigned edge, double *out);
int nrxh_brlen_sumtables(void *h, unsigned edge, unsigned *count);
/* nrxh_brlen_logl + nrxh_brlen_sumtables of the same branch in one pass over the displayed-tree pairs' CLVs
 * (netrax::computeLoglikelihoodBrlenOptAndSumtables; the reference makes the two calls back to back, BranchLengthOptimization.cpp:374-381) */
int nrxh_brlen_logl_sumtables(void *h, unsigned edge, double *out, unsigned *count);
int nrxh_brlen_read_sumtable(void *h, unsigned p, unsigned idx, double *out, double *tree_prob, unsigned *left_tree, unsigned *right_tree);
int nrxh_brlen_set_length(void *h, int partition, unsigned edge, double value);
int nrxh_brlen_derivatives(void *h, unsigned edge, double *d1, double *d2, double *part_d1, double *part_d2, double *raw);
int nrxh_brlen_finish(void *h, unsigned edge, double *final_logl);
/* Virtual re-rooting on the device never overwrites a root-directed CLV (the reference re-roots in place and recomputes,
 * src/likelihood/VirtualRerooting.cpp:192-252) and memoises the re-rooted trees of the path nodes between calls:
 *   nrxh_brlen_sweep_order       <- the candidate order of optimize_branches_internal (BranchLengthOptimization.cpp:423-476 visits an
 *                                   unordered_set): all branches in depth-first pre-order, edges_out[num_branches]
 *   nrxh_reroot_stats            <- memo hits / misses (processNodeImproved calls of re-rooting paths), live entries and the CLV slots they hold
 *   nrxh_set_reroot_cache_slots  <- slot budget of the memo (-1: default, 0: nothing survives a session) */
/* nrxh_set_score_only: full evaluations (incremental = 0) that replay the cached plan do not store the CLVs of the root displayed trees
 * — candidate scoring (src/search/Filtering.cpp:210-260: performMove, score, undoMove) reads nothing but the lnL.  The next incremental
 * evaluation, re-rooting or CLV read-back first re-evaluates with the stores on. */
int nrxh_set_score_only(void *h, int on);
/* nrxh_set_lazy_rerooting: optimize_branch / nrxh_brlen_prepare + _finish skip the evaluation from the network root before a re-rooting
 * and after the branch is done (src/optimization/BranchLengthOptimization.cpp:352,420) when the branch is active and alive in every
 * displayed tree and its re-rooting plan is known: only the root-directed CLVs the re-rooting reads are brought up to date, the lnL
 * returned is the edge-rooted one (computeLoglikelihoodBrlenOpt; equal to the root's to rounding), nrxh_brlen_prepare hands out the
 * lnL of the previous step.  Stale CLVs are recomputed when read, at the latest by the next nrxh_compute_loglikelihood. */
int nrxh_set_lazy_rerooting(void *h, int on);
int nrxh_lazy_reroot_stats(void *h, unsigned long long *sessions, unsigned long long *fallbacks);   /* lazily prepared re-rootings / of those, redone from an evaluated root */
int nrxh_brlen_sweep_order(void *h, unsigned *edges_out);
int nrxh_reroot_stats(void *h, unsigned long long *hits, unsigned long long *misses, unsigned *entries, unsigned *cached_slots);
int nrxh_set_reroot_cache_slots(void *h, long long max_slots);
/* The immediate callers of the path, same control flow as the reference:
 *   nrxh_optimize_branch(es)      <- netrax::optimize_branch / optimize_branches  src/optimization/BranchLengthOptimization.cpp:345-421,423-476,567-576
 *   nrxh_optimize_reticulation(s) <- netrax::optimize_reticulation(s)             src/optimization/ReticulationOptimization.cpp:68-117
 * method: 0 BRENT_NORMAL, 1 BRENT_REROOT, 2 NEWTON_RAPHSON (src/NetraxOptions.hpp:17-21; the reference's default is 2). */
int nrxh_optimize_branch(void *h, unsigned edge, int method, unsigned max_iters, double *final_logl);
int nrxh_optimize_branches(void *h, int max_iters, int max_iters_outside, int radius, int method, double *final_logl);
int nrxh_optimize_reticulation(void *h, unsigned r, double *final_logl);
int nrxh_optimize_reticulations(void *h, int max_iters, double *final_logl);
/* model-parameter loop: Gamma shape of partition p (treeinfo_set_alpha, PLLMOD/algorithm/pllmod_algorithm.c:566-587)
 * and the ALPHA step of optimize_params (src/optimization/ModelOptimization.cpp:56-65 = pllmod_algo_opt_onedim_treeinfo):
 * Brent over the alphas of all partitions that carry one, one full device re-evaluation per iterate. */
/* src/likelihood/PseudoLoglikelihood.hpp: the pseudo-likelihood (also what nrxh_compute_loglikelihood returns when the
 * likelihood variant is 2 = SARAH_PSEUDO, src/likelihood/LikelihoodComputation.cpp:23-27) + test hooks for its per-node CLVs */
int nrxh_compute_pseudo_loglikelihood(void *h, int incremental, int update_pmatrices, double *out);
int nrxh_read_pseudo_clv(void *h, unsigned node, unsigned p, double *out);
int nrxh_read_pseudo_scaler(void *h, unsigned node, unsigned p, unsigned *out);
/* src/likelihood/ComplexityScoring.hpp: BIC of the network (what the search compares) */
int nrxh_score_network(void *h, double *bic_score);
int nrxh_set_scoring_sizes(void *h, unsigned long long total_num_model_parameters, unsigned long long total_num_sites /* 0: keep */);
/* src/optimization/Optimization.hpp:23-25: model (alpha or the caller's optimize_params), reticulation probabilities and
 * branch lengths in rounds until the BIC stops improving; type 0 QUICK, 1 NORMAL, 2 SLOW */
int nrxh_optimize_all_non_topology(void *h, int type, double *bic_score);
/* The slot pll-modules' optimisers re-enter through: same signature as pllmod_treeinfo_t::likelihood_target_function
 * (PLLMOD/tree/pll_tree.h:267-271) / network_logl_wrapper (src/RaxmlWrapper.cpp:21-26); params from nrxh_network_params(h). */
double nrxh_likelihood_target_function(void *network_params, int incremental, int update_pmatrices, double **persite_lnl);
void *nrxh_network_params(void *h);
/* +I: proportion of invariant sites of partition p (pll_update_invariant_sites_proportion, LIBPLL/models.c:495-543) */
int nrxh_set_pinv(void *h, unsigned p, double prop_invar);
int nrxh_set_alpha(void *h, unsigned p, double alpha);
int nrxh_get_alpha(void *h, unsigned p, double *alpha);
int nrxh_optimize_alpha(void *h, double min_alpha, double max_alpha, double tolerance, double *final_logl);
int nrxh_get_branch_lengths(void *h, int partition /* -1: linked */, double *out /* [edges] */);
int nrxh_get_reticulation_probs(void *h, double *out /* [reticulations] */);
unsigned long long nrxh_clv_update_count(void *h);
void nrxh_reset_counters(void *h);
int nrxh_gamma_rates(double alpha, unsigned cats, int mode, double *out);
/* device-free: the host's eigendecomposition (role of pll_update_eigen, LIBPLL/models.c:293-410) of the reversible model
 * (freqs[states], subst[states (states - 1) / 2]): eigenvecs / inv_eigenvecs [states][states_padded], eigenvals [states_padded]
 * — exactly what the host hands to nrx_set_model */
int nrxh_eigen_decompose(unsigned states, const double *freqs, const double *subst, double *eigenvecs, double *inv_eigenvecs,
                         double *eigenvals);
/* device-free: the host library's restatements of pll-modules' single-variable minimisers — pllmod_opt_minimize_newton_multi
 * with xnum = 1 (PLLMOD/optimize/opt_algorithms.c:133-261; deriv(ctx, x, f', f'') as its deriv_func; *converged = PLL_SUCCESS /
 * PLL_FAILURE, x holds the last iterate either way) and pllmod_opt_minimize_brent (:1404-1429; stops at convergence instead of
 * re-evaluating the unchanged proposal up to iteration 101).  optimize_branch / optimize_reticulation run these. */
int nrxh_minimize_newton(double xmin, double *x, double xmax, double tolerance, unsigned max_iters,
                         void (*deriv)(void *ctx, double *x, double *f, double *df), void *ctx, int *converged);
int nrxh_minimize_brent(double xmin, double xguess, double xmax, double xtol, double (*target)(void *ctx, double x), void *ctx, double *xopt);
/* pllmod_opt_minimize_brent_multi(xnum = n, opt_mask all set, global_range = 1) (:1431-1459 -> brent_opt_alt :1040-1254), the
 * driver under optimize_alpha / optimize_pinv / optimize_scalers: target(ctx, x, fx, converged) with pll-modules' contract
 * (converged NULL on plain calls; otherwise skip converged[j] != 0 and write the "all converged" flag to converged[n]).
 * x: in = the guesses, out = the optima. */
int nrxh_minimize_brent_multi(unsigned n, double xmin, double *x, double xmax, double xtol,
                              double (*target)(void *ctx, double *x, double *fx, int *converged), void *ctx);
/* bench / profiling hooks */
unsigned long long nrxh_launch_count(void *h);
unsigned nrxh_num_slots(void *h);
int nrxh_profile_enable(void *h, int on);
int nrxh_profile_read(void *h, double *clv_ms, unsigned long long *launches, unsigned long long *site_updates, unsigned long long *bytes);
/* per kernel family (kind = NRX_PROF_* of nrx_engine.h): device ms, launches, units of work, algorithmic bytes */
int nrxh_profile_read_kind(void *h, int kind, double *ms, unsigned long long *launches, unsigned long long *units, unsigned long long *bytes,
                           unsigned long long *compulsory_bytes);
int nrxh_persite_lnl(void *h, unsigned tree, double *out /* [nparts][max_sites] */, unsigned stride);
void *nrxh_engine(void *h); /* the underlying nrx_engine* */
/* re-upload one partition's alignment slice from HOST buffers (tipchars: 1 byte per cell, DNA) + pattern weights */
int nrxh_upload_alignment_u8(void *h, unsigned p, const uint8_t *tipchars, const unsigned *pattern_weights);
/* the same for any alphabet (e.g. 20 states): 1-byte codes + the code -> state-set map; asynchronous like the call above */
/* double-buffered upload (4-state partitions): stage the NEXT alignment while the current one is being evaluated, commit before the
 * evaluation that should see it (nrx_stage_alignment_u8 / nrx_commit_staged_alignment; the role of re-reading the MSA slices between
 * analyses, src/main.cpp:715-736, without stalling the device) */
int nrxh_stage_alignment_u8(void *h, unsigned partition, const uint8_t *tipchars, const unsigned *pattern_weights);
int nrxh_commit_staged_alignment(void *h);
int nrxh_upload_alignment_codes(void *h, unsigned p, const uint8_t *codes, const uint32_t *tipmap, unsigned ncodes, const unsigned *pattern_weights);
int nrxh_timer_start(void *h);
int nrxh_timer_stop(void *h, double *elapsed_ms);

#ifdef __cplusplus
}
#endif
#endif
