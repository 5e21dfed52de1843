/*
 * nrx_engine.h — C-ABI of the B200 (sm_100a) device engine for NetRAX's network-likelihood hot path.
 *
 * This is the lower seam of the drop-in: the entry points below replace, for ALL displayed
 * trees of a network node at once, the seven explicit-pointer libpll calls NetRAX issues once per displayed
 * tree, per node, per partition (LIBPLL = libs/raxml-ng/libs/pll-modules/libs/libpll/src in the reference):
 *
 *   nrx_update_pmatrices  <- pll_update_prob_matrices            LIBPLL/pll.h:731-735, models.c:412-443
 *   nrx_update_clvs       <- pll_update_partials_single          LIBPLL/pll.h:798-806, partials.c:196-240
 *   nrx_tree_lnl          <- pll_compute_root_loglikelihood      LIBPLL/pll.h:752-757, likelihood.c:122-184
 *   nrx_edge_lnl          <- pll_compute_edge_loglikelihood      LIBPLL/pll.h:759-768, likelihood.c:555-615
 *   nrx_sumtables         <- pll_update_sumtable                 LIBPLL/pll.h:815-823, derivatives.c:246-326
 *   nrx_derivatives       <- pll_compute_diagptable + pll_compute_loglikelihood_derivatives
 *                                                                LIBPLL/pll.h:830-842, core_derivatives.c:696-959
 *   nrx_engine_create / nrx_set_tips / nrx_set_pattern_weights / nrx_set_model
 *                         <- pll_partition_create, pll_set_tip_states, pll_set_pattern_weights and the
 *                            model fields of pll_partition_t after pll_update_eigen (LIBPLL/pll.h:230-277,606-632)
 *
 * Conventions: plain C, no torch / C++ types; every call returns 1 on success and 0 on failure
 * (PLL_SUCCESS / PLL_FAILURE, LIBPLL/pll.h:75-76) with a thread-local message in nrx_last_error()
 * (mirrors __thread pll_errmsg, LIBPLL/pll.c:24-25).  A handle is single-threaded and stream-ordered;
 * host pointers are borrowed for the duration of the call only; the engine owns all device memory.
 * There is NO CPU fallback: every entry point fails if no CUDA device is usable.
 *
 * Data layout in HBM (per partition p):  CLV slot  = double[patterns][rate_cats][states_padded]
 * (states_padded = (states+3)&~3, the libpll AVX layout: DNA+G4 = 128 B per pattern), scaler slot =
 * uint32[patterns] (per-site scalers; PLL_ATTRIB_RATE_SCALERS is off in NetRAX, src/RaxmlWrapper.cpp:336),
 * tips = uint8 codes [tips][patterns] (PATTERN_TIP), P-matrices = double[edges][cats][states][states_padded].
 * A "slot" index addresses the same displayed-tree CLV in every partition, like the per-partition
 * clv_vector[] of the reference's DisplayedTreeData (src/graph/DisplayedTreeData.hpp:18-42).
 */
#ifndef NRX_ENGINE_H
#define NRX_ENGINE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nrx_engine nrx_engine;

typedef struct nrx_partition_desc {
  uint32_t states;    /* 4 (DNA) or 20 (protein); anything <= 32 runs on the generic kernels */
  uint32_t rate_cats; /* 1..16 */
  uint32_t patterns;  /* LOCAL pattern count: this rank's slice of the partition (may be 0) */
  uint32_t tips;
  uint32_t edges;     /* number of P-matrices INCLUDING the fake zero-length one (network edges + 1) */
} nrx_partition_desc;

enum { NRX_CLV = 0, NRX_TIP = 1, NRX_NONE = 2 };

/* One CLV update = one pll_operation_t (LIBPLL/pll.h:314-324) as built by src/likelihood/Operation.cpp:7-35,
 * with the output / input CLVs named by slot instead of by host pointer.  kind NRX_NONE is the reference's
 * "fake" all-ones CLV behind the fake identity P-matrix (ImprovedLoglikelihood.cpp:122,138-139). */
typedef struct nrx_op {
  uint32_t parent_slot;
  uint32_t left_kind, left_idx, left_edge;    /* idx: slot (NRX_CLV) or tip number (NRX_TIP) */
  uint32_t right_kind, right_idx, right_edge;
  uint32_t lnl_item; /* 0: none.  k+1: this CLV is root displayed tree k of the traversal — the kernel also emits its
                      * per-site likelihood term (first half of K3 fused into K2's epilogue; see nrx_plan_create / nrx_tree_lnl_fused) */
} nrx_op;

/* One pseudo-likelihood CLV update (src/likelihood/PseudoLoglikelihood.cpp:57-190): the node's CLV is the weighted blend
 * w[0] * update(left, right) + w[1] * update(left, fake) + w[2] * update(fake, right) + w[3] * 1 of up to three libpll
 * updates (each with its own per-site scaling test) — the reference runs them into scratch CLVs and merges on the host
 * (merge_clvs, :8-55).  The parent scaler is that of the LAST executed update, as in the reference. */
typedef struct nrx_pseudo_op {
  uint32_t parent_slot;
  uint32_t left_kind, left_idx, left_edge;
  uint32_t right_kind, right_idx, right_edge;
  uint32_t pad_;
  double w[4];
} nrx_pseudo_op;

/* An (a, b) operand pair on one edge: (source-tree, target-tree) of computeLoglikelihoodBrlenOpt /
 * computePartitionSumtables (src/likelihood/VirtualRerooting.cpp:403-439, LikelihoodDerivatives.cpp:312-341). */
typedef struct nrx_pair {
  uint32_t a_kind, a_idx, b_kind, b_idx;
} nrx_pair;

const char *nrx_last_error(void);
int nrx_device_count(void);

nrx_engine *nrx_engine_create(const nrx_partition_desc *parts, uint32_t nparts, int device);
void nrx_engine_destroy(nrx_engine *e);

/* tip_masks: [tips][patterns] state bit masks (bit k = state k; DNA 1..15).  Re-coded to uint8 on upload. */
int nrx_set_tips(nrx_engine *e, uint32_t p, const uint32_t *tip_masks);
/* Same, for callers that already hold libpll-style tipchars (1 byte per cell; DNA: the 4-bit mask itself), ASYNCHRONOUS: the
 * copy into the padded device rows and a device-side check — every code must denote a non-empty state set, as
 * pll_set_tip_states requires (LIBPLL/pll.c:875-957); the invariant-site table of +I is rebuilt too — are enqueued on the engine's
 * stream and the call returns.  `codes` is borrowed until the engine next synchronises (nrx_result_wait, nrx_sync, any
 * read-back); an illegal code makes THAT call fail with "Illegal state code in tip". */
int nrx_set_tipchars_u8(nrx_engine *e, uint32_t p, const uint8_t *codes);
/* the same for any alphabet: code c stands for the state set tipmap[c] (c < ncodes <= 256), the role of pll_map_aa & co. */
int nrx_set_tipcodes_u8(nrx_engine *e, uint32_t p, const uint8_t *codes, const uint32_t *tipmap, uint32_t ncodes);
int nrx_set_pattern_weights_async(nrx_engine *e, uint32_t p, const uint32_t *weights); /* borrowed like `codes` above */
/* Double-buffered form of the asynchronous upload (4-state partitions): stage copies the NEXT alignment (tips [tips][patterns] as state
 * masks and / or pattern weights; either may be NULL) into shadow buffers on a separate copy stream — it overlaps with every kernel
 * enqueued after this call, i.e. with the evaluation of the current alignment; commit makes the staged buffers the live ones
 * (the engine stream waits for the copies), validates the codes and rebuilds the invariant-site table.  Host buffers are borrowed until
 * the commit's evaluation has been collected. */
int nrx_stage_alignment_u8(nrx_engine *e, uint32_t partition, const uint8_t *codes, const uint32_t *weights);
int nrx_commit_staged_alignment(nrx_engine *e);
int nrx_set_pattern_weights(nrx_engine *e, uint32_t p, const uint32_t *weights);
/* eigenvecs / inv_eigenvecs: [states][states_padded]; eigenvals, freqs: [states_padded] (padding ignored);
 * prop_invar in [0, 1): proportion of invariant sites (+I; pll_update_invariant_sites_proportion, LIBPLL/models.c:495-543).
 * The invariant-pattern table (pll_update_invariant_sites, LIBPLL/models.c:651-760) is derived from the tips on upload. */
int nrx_set_model(nrx_engine *e, uint32_t p, const double *freqs, const double *eigenvecs,
                  const double *inv_eigenvecs, const double *eigenvals, const double *rates,
                  const double *rate_weights, double prop_invar);
/* The same for mixtures with one rate matrix per category (LG4M / LG4X; raxml-ng's ratecat_submodels forwarded by NetRAX as
 * libpll's params_indices, src/RaxmlWrapper.cpp:199-203, LIBPLL/core_pmatrix.c:182-185): category c uses matrix cat_model[c] <
 * nmodels <= 16; freqs / eigenvals are [nmodels][states_padded], eigenvecs / inv_eigenvecs [nmodels][states][states_padded].
 * nmodels == 1 (cat_model may be NULL) is nrx_set_model.  K1 and the generic K3-K6 kernels index the model by category;
 * K2 reads only P-matrices and keeps its fast paths. */
int nrx_set_model_mixture(nrx_engine *e, uint32_t p, uint32_t nmodels, const uint32_t *cat_model, const double *freqs,
                          const double *eigenvecs, const double *inv_eigenvecs, const double *eigenvals, const double *rates,
                          const double *rate_weights, double prop_invar);
/* K1: P(t) for n edges of partition p in one launch. */
int nrx_update_pmatrices(nrx_engine *e, uint32_t p, uint32_t n, const uint32_t *edge_idx, const double *brlen);
int nrx_get_pmatrix(nrx_engine *e, uint32_t p, uint32_t edge, double *out);
int nrx_set_pmatrix(nrx_engine *e, uint32_t p, uint32_t edge, const double *in); /* test hook */

/* CLV / scaler slot pool (all partitions): makes slots [0, nslots) exist. */
int nrx_reserve_slots(nrx_engine *e, uint32_t nslots);
uint32_t nrx_num_slots(nrx_engine *e);
int nrx_copy_slot(nrx_engine *e, uint32_t dst, uint32_t src); /* device-side deep copy (extractOldTrees etc.) */
/* n copies dst[i] <- src[i] (all partitions) in one launch per partition shape */
int nrx_copy_slots(nrx_engine *e, const uint32_t *dst, const uint32_t *src, uint32_t n);

/* K2: ONE launch per same-shape partition group updating `nops` CLVs (all displayed trees of a node —
 * or of several independent nodes — x patterns x rate categories).  The ops must be mutually independent. */
int nrx_update_clvs(nrx_engine *e, const nrx_op *ops, uint32_t nops);

/* K2p: `nops` mutually independent pseudo-likelihood updates in ONE launch per partition shape: reads each child once and
 * writes the merged CLV once (the reference: three kernel calls + a merge pass over three scratch CLVs). */
int nrx_update_pseudo_clvs(nrx_engine *e, const nrx_pseudo_op *ops, uint32_t nops);

/* Evaluation plan: the K2 batches of one whole traversal (the enumeration of displayed trees depends only on the
 * topology).  The ops stay resident on the device and nrx_plan_run replays all batches as ONE
 * CUDA graph launch (captured on first use) — small alignments are launch-latency-bound, not HBM-bound.
 * ops = the batches back to back, batch_sizes[nbatches]; batch b may depend on batches < b only. */
int nrx_plan_create(nrx_engine *e, const nrx_op *ops, const uint32_t *batch_sizes, uint32_t nbatches, uint32_t *plan_id);
int nrx_plan_run(nrx_engine *e, uint32_t plan_id);
/* 1 when plans may carry lnl_item marks (every partition runs on the pipelined 4-state kernel) */
int nrx_supports_fused_lnl(nrx_engine *e);
/* K3 for the n marked trees (slots[k] = the CLV slot of mark k+1) of the plan that has just run: log / scaler /
 * pattern weight and the sum over the per-site likelihood terms the K2 epilogue wrote (16 B per site read instead of
 * the whole CLV) -> out[n][nparts], bit-identical to nrx_tree_lnl on the same slots. */
int nrx_tree_lnl_fused(nrx_engine *e, uint32_t plan_id, const uint32_t *slots, uint32_t n, double *out);
/* nrx_plan_run + nrx_tree_lnl_fused_async in one call: the whole evaluation of a plan (all CLVs of the traversal + the n marked
 * trees' root lnLs; result collected with nrx_result_wait).  Small alignments are launch-latency-bound (one launch per
 * dependency level), so when the plan has a "tile-walk" form — every partition on the 4-state x 4-category kernels, the live
 * CLVs of a depth-first walk fit one block's shared memory — this is ONE launch: a block owns a tile of patterns and walks the
 * whole plan for it, children from shared memory, every CLV still written to its HBM slot (k_walk_dna4).  env NRX_WALK=0 / 1 / 2:
 * never / whenever possible / when possible and at most NRX_WALK_TILES tiles (default). */
int nrx_plan_evaluate_async(nrx_engine *e, uint32_t plan_id, const uint32_t *slots, uint32_t n);
int nrx_plan_destroy(nrx_engine *e, uint32_t plan_id);

/* K3: per-tree per-partition root lnL, out[n][nparts] (LOCAL sums: the caller all-reduces across ranks).
 * persite (optional): [n] pointers-free layout out_persite[(i * nparts + p) * max_patterns + site]. */
int nrx_tree_lnl(nrx_engine *e, const uint32_t *slots, uint32_t n, double *out, double *persite,
                 size_t persite_stride);
/* Asynchronous forms for batched scoring of several networks (one engine = one stream each): the launches, the
 * cross-rank all-reduce and the device->host copy are enqueued and the call returns; nrx_result_wait blocks on THIS
 * engine's stream and hands out the `count` = n * nparts doubles.  One result may be pending per engine. */
int nrx_tree_lnl_async(nrx_engine *e, const uint32_t *slots, uint32_t n);
int nrx_tree_lnl_fused_async(nrx_engine *e, uint32_t plan_id, const uint32_t *slots, uint32_t n);
int nrx_result_wait(nrx_engine *e, double *out, uint32_t count);
/* Launch geometry of K2: latency (default: a small launch spreads over all SMs, one tile per block) or throughput (several
 * engines share the GPU: fewer, longer-running blocks; measured +40 % evaluations/s with 32 networks in flight). */
int nrx_set_throughput_mode(nrx_engine *e, int on);
/* Score-only replays (candidate scoring, src/search/Filtering.cpp:210-260 reads nothing but the lnL): while on, the ops of a plan that
 * carry an lnl mark — the root displayed trees, whose per-site lnL K2 emits itself — do NOT store their CLV (25 % of the write
 * stream of BASELINE config 5); scalers and per-site terms are written as always.  The slots of those trees then hold stale CLVs:
 * the caller must re-evaluate without the flag before anything reads them (the host layer does, AnnotatedNetwork::score_only). */
int nrx_set_score_only(nrx_engine *e, int on);
/* K4: edge lnL for n operand pairs over P-matrix `edge`, out[n][nparts]. */
int nrx_edge_lnl(nrx_engine *e, uint32_t edge, const nrx_pair *pairs, uint32_t n, double *out);
/* K5: sumtables for n pairs into sumtable slots [0, n) (pool grows on demand). */
int nrx_sumtables(nrx_engine *e, const nrx_pair *pairs, uint32_t n);
/* K4 + K5 of one branch in ONE pass over the pairs' CLVs: sumtable slot i for pairs[i], and for the pairs with lnl_index[i] >= 0 the edge
 * lnL as nrx_edge_lnl gives it, out[lnl_index[i]][nparts] (n_lnl outputs; the reference calls pll_compute_edge_loglikelihood and
 * pll_update_sumtable on the same CLVs back to back: src/likelihood/VirtualRerooting.cpp:279-346, LikelihoodDerivatives.cpp:291-344). */
int nrx_edge_lnl_sumtables(nrx_engine *e, uint32_t edge, const nrx_pair *pairs, uint32_t n, const int32_t *lnl_index, uint32_t n_lnl, double *out);
/* K6: for sumtable slots [0, n): out[n][nparts][3] = (f, d(-lnL)/dt, d2(-lnL)/dt2) at brlen[p]
 * (f = sum w log lk0 without scaler term, the reference's AVX2 behaviour). */
int nrx_derivatives(nrx_engine *e, uint32_t n, const double *brlen_per_partition, double *out);

/* parity/debug read-back */
int nrx_read_clv(nrx_engine *e, uint32_t p, uint32_t slot, double *out);
int nrx_read_scaler(nrx_engine *e, uint32_t p, uint32_t slot, uint32_t *out);
int nrx_read_sumtable(nrx_engine *e, uint32_t p, uint32_t st_slot, double *out);
int nrx_sync(nrx_engine *e);

/* C2-C4: site-shard communicator.  Replaces the reference's MPI_Allreduce(MPI_IN_PLACE, MPI_DOUBLE, SUM) behind
 * parallel_reduce_cb (libs/raxml-ng/src/ParallelContext.cpp:425-487, installed at src/RaxmlWrapper.cpp:717-718):
 * once a communicator is attached, nrx_tree_lnl / nrx_edge_lnl / nrx_derivatives all-reduce their [n][nparts](x3)
 * result ON THE DEVICE with ONE ncclAllReduce (NVLink 5 / NVSwitch) on the engine's stream before the single
 * device->host copy, so every rank returns the global sums.  rank 0 obtains the 128-byte NCCL unique id and hands
 * it to the other ranks by any out-of-band channel (the launcher's store; MPI_Bcast in a NetRAX build).
 * libnccl.so.2 is resolved at run time (dlopen): the engine has no link-time NCCL dependency. */
int nrx_comm_get_unique_id(uint8_t *id128);
int nrx_comm_init(nrx_engine *e, const uint8_t *id128, int rank, int nranks);
int nrx_comm_size(nrx_engine *e); /* 1 when no communicator is attached */
/* SUM all-reduce of n host doubles across the ranks (staged through the device; for callers' own scalars) */
int nrx_comm_allreduce_sum(nrx_engine *e, double *host_inout, size_t n);

/* device-side access for callers that keep data on the GPU (NCCL all-reduce of the [n][nparts] results):
 * with a communicator attached (or NRX_ZEROCOPY=0) the last nrx_tree_lnl / nrx_edge_lnl / nrx_derivatives result also stays in this
 * device buffer; otherwise the reducing kernels write it straight into the engine's mapped host buffer and this one is not touched. */
void *nrx_result_device_ptr(nrx_engine *e);
void *nrx_stream(nrx_engine *e); /* cudaStream_t */
/* CUDA-event stopwatch on the engine's own stream (torch.cuda.Event cannot see this stream). */
int nrx_timer_start(nrx_engine *e);
int nrx_timer_stop(nrx_engine *e, double *elapsed_ms);
/* number of kernels launched by this engine so far (bench.py "gpu_launches") */
unsigned long long nrx_launch_count(nrx_engine *e);
/* device time (ms) spent in K2 launches since the last reset, measured with CUDA events on the engine
 * stream when profiling is enabled (bench roofline) */
int nrx_profile_enable(nrx_engine *e, int on);
int nrx_profile_read(nrx_engine *e, double *clv_ms, unsigned long long *clv_launches, unsigned long long *clv_site_updates,
                     unsigned long long *clv_bytes);
/* the same per kernel family: device ms (CUDA events around the family's launches), launches, units of work
 * (edges for K1, site-updates for K2, (item, pattern) pairs for K3-K6, outputs for the reduction),
 * ALGORITHMIC bytes (§8d table: every op / pair charged all of its operands) and COMPULSORY bytes (per launch, every
 * distinct operand CLV / tip row read once + every output written once: what has to cross the HBM pins when the ops of a
 * launch share operands through the L2 — the numerator of the roofline fraction) */
enum {
  NRX_PROF_K2 = 0,      /* CLV update (k_clv_*) */
  NRX_PROF_K1 = 1,      /* P-matrices (+ protein tip tables) */
  NRX_PROF_K3 = 2,      /* root lnL per tree (k_tree_lnl*) */
  NRX_PROF_K3F = 3,     /* second stage of the fused root lnL (k_term_lnl_sum) */
  NRX_PROF_K4 = 4,      /* edge lnL per pair */
  NRX_PROF_K5 = 5,      /* sumtables */
  NRX_PROF_K6 = 6,      /* derivatives per sumtable and Newton iterate */
  NRX_PROF_REDUCE = 7,  /* second-stage reduction of the per-block partial sums */
  NRX_PROF_COPY = 8,    /* slot copies of the virtual re-rooting save/restore */
  NRX_PROF_K45 = 9,     /* edge lnL + sumtables of the same pairs in one pass (k_edge_sum_dna4q) */
  NRX_PROF_KINDS = 10
};
int nrx_profile_read_kind(nrx_engine *e, int kind, double *ms, unsigned long long *launches, unsigned long long *units,
                          unsigned long long *bytes, unsigned long long *compulsory_bytes);

#ifdef __cplusplus
}
#endif
#endif
