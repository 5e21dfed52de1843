/*
 * nrx_engine.cu — C-ABI (include/nrx_engine.h) over the sm_100a kernels in nrx_kernels.cuh.
 * Device memory pool (CLV / scaler / sumtable slots), per-partition model + P-matrix storage, launch
 * geometry, deterministic two-stage reductions.  One handle = one GPU = one stream.
 */
#include "nrx_kernels.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <map>
#include <string>
#include <vector>

using namespace nrx;

namespace {
thread_local std::string g_err;

bool cuda_ok(cudaError_t e, const char *what) {
  if (e == cudaSuccess) return true;
  g_err = std::string(what) + ": " + cudaGetErrorString(e);
  return false;
}
#define CK(call) do { if (!cuda_ok((call), #call)) return 0; } while (0)

struct Part {
  nrx_partition_desc d{};
  uint32_t sp = 0, pat_pad = 0;  // pat_pad: patterns rounded up to a whole K2 tile (bulk copies always move full tiles)
  size_t clv_entries = 0, pmat_entries = 0;
  double *pmat = nullptr;
  double *pmat_pad = nullptr;   // DNA 4x4 partitions: P-matrices at k_walk_dna4's shared-memory pitch (written by K1 alongside pmat)
  double *tiplut = nullptr;  // 20-state partitions only
  bool invariant_stale = false;  // tips were replaced through nrx_set_tipchars_u8 while pinv == 0: recompute before +I is switched on
  double pinv = 0.0;        // proportion of invariant sites (+I)
  int *invariant = nullptr; // device [pat_pad]: pll_update_invariant_sites
  double *summat = nullptr, *sumlut = nullptr;  // 20-state partitions only: K5 operand matrices / tip table (PartView)
  std::vector<uint32_t> h_tipmap;               // code -> state mask, host copy
  std::vector<double> h_freqs, h_inv_eigenvecs, h_eigenvecs;
  uint8_t *tipchars = nullptr;
  uint32_t *tipmap = nullptr, *weights = nullptr;
  // double buffer of the alignment (nrx_stage_alignment_u8 / nrx_commit_staged_alignment): the next step's tips and weights are copied on
  // the engine's copy stream while the current step computes; commit swaps the pointers
  uint8_t *tipchars_next = nullptr;
  uint32_t *weights_next = nullptr;
  bool staged_tips = false, staged_weights = false;
  double *model = nullptr;  // freqs | eigenvecs | inv_eigenvecs | eigenvals | rates | rate_weights | diagp
  double *freqs = nullptr, *eigenvecs = nullptr, *inv_eigenvecs = nullptr, *eigenvals = nullptr, *rates = nullptr, *rate_weights = nullptr, *diagp = nullptr;
  std::vector<double> h_eigenvals, h_rates;   // h_eigenvals: [nmodels][states]
  uint32_t nmodels = 1, model_cap = 1;        // rate matrices in use / the model buffer has room for (PartView::nmodels)
  uint8_t cat_model[16] = {};                 // rate category -> rate matrix (libpll's params_indices)
  std::vector<void *> slot_mem;       // one allocation per slot: [clv | scaler]
  std::vector<double *> h_clv;
  std::vector<uint32_t *> h_scaler;
  std::vector<double *> h_sumtable;
  double **d_clv = nullptr;
  uint32_t **d_scaler = nullptr;
  double **d_sumtable = nullptr;
  uint32_t table_cap = 0, st_cap = 0;
  bool model_set = false, tips_set = false;
  std::vector<double> h_len;        // current branch length per edge (-1: never set) — the tile walk recomputes P-matrices from these
  std::vector<char> pend;           // edges whose P-matrix update is deferred (see nrx_update_pmatrices)
  uint32_t npend = 0;
  uint32_t tip_codes = 0;  // distinct tip codes in use
};

struct ShapeClass {
  uint32_t states, cats;
  std::vector<uint32_t> parts;     // partition indices
  PartView *d_views = nullptr;     // device array, same order
  uint32_t max_patterns = 0;
};
}  // namespace

/* ---- NCCL, resolved at run time ------------------------------------------------------------------------ */
namespace {
struct NcclApi {
  typedef struct { char internal[128]; } UniqueId;
  int (*GetUniqueId)(UniqueId *) = nullptr;
  int (*CommInitRank)(void **comm, int nranks, UniqueId id, int rank) = nullptr;
  int (*AllReduce)(const void *send, void *recv, size_t count, int dtype, int op, void *comm, cudaStream_t s) = nullptr;
  int (*CommDestroy)(void *comm) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  bool ok = false;
};
NcclApi &nccl() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (h) {
      api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
      api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
      api.AllReduce = (decltype(api.AllReduce))dlsym(h, "ncclAllReduce");
      api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
      api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
      api.ok = api.GetUniqueId && api.CommInitRank && api.AllReduce && api.CommDestroy && api.GetErrorString;
    }
  }
  return api;
}
constexpr int NCCL_FLOAT64 = 8, NCCL_SUM = 0;  // ncclDataType_t / ncclRedOp_t values (nccl.h)
}  // namespace

struct EnginePlan {
  bool alive = false;
  nrx_op *d_ops = nullptr;
  std::vector<size_t> offsets;
  std::vector<uint32_t> sizes;
  std::vector<char> tips;
  std::vector<uint32_t> ntt;   // leading tip-tip ops of each batch (20-state engines order them first)
  cudaGraphExec_t exec[4] = {nullptr, nullptr, nullptr, nullptr};   // [0] latency geometry, [1] throughput geometry (nrx_set_throughput_mode); +2: score-only (nrx_set_score_only)
  unsigned long long updates = 0, bytes = 0, cbytes = 0;
  uint32_t lnl_items = 0;    // number of ops carrying an lnl_item mark (fused K3)
  // node-centric K2 (k_clv_node_dna4): per batch, the groups of ops that share children; the batch's remaining ops stay in d_ops
  nrx_node_group *d_ngroups = nullptr;
  nrx_node_op *d_nops = nullptr;
  nrx_node_block *d_nblocks = nullptr;
  std::vector<uint32_t> node_block_off, node_block_cnt, node_ncmax;   // per batch: range in d_nblocks (count 0: no node launch), largest group (children)
  // tile-walk form of the plan (k_walk_dna4): the ops in a depth-first order with shared-memory buffer assignments
  nrx_walk_op *d_walk = nullptr;
  uint32_t walk_nops = 0, walk_nbuf = 0;
  size_t walk_smem = 0;
  unsigned long long walk_cbytes = 0;   // compulsory HBM bytes of one walk: every parent CLV + scaler written, every tip row read
};

struct nrx_engine {
  int device = 0;
  std::vector<EnginePlan> plans;
  double *d_fused = nullptr;   // per-site lnLs written by K2's fused epilogue: [items][nparts][max_patterns]
  size_t fused_cap = 0;
  bool use_graphs = true;  // env NRX_GRAPH=0: replay plans as individual launches
  void *comm = nullptr;  // ncclComm_t
  int comm_rank = 0, comm_size = 1;
  cudaStream_t stream = nullptr;
  std::vector<Part> parts;
  std::vector<ShapeClass> classes;
  uint32_t nslots = 0, nsumtables = 0;
  uint32_t max_patterns = 0;
  // staging
  void *h_stage = nullptr;      // pinned
  void *d_stage = nullptr;
  size_t stage_cap = 0, stage_off = 0;
  double *d_partial = nullptr;
  size_t partial_cap = 0;
  double *d_result = nullptr, *h_result = nullptr;
  size_t result_cap = 0;
  int *h_err = nullptr;            // mapped pinned flag raised by k_check_tipchars (illegal tip code in an asynchronous upload)
  uint32_t *d_tickets = nullptr;   // one self-resetting ticket counter per reduction output (item, partition): fused second stage
  bool fuse_reduce = true;         // env NRX_FUSE_REDUCE=0: separate k_reduce_partials launch (A/B)
  bool defer_pmat = false;         // P-matrix updates are deferred until a launch needs them: the tile walk computes them itself (one launch per evaluation)
  uint32_t node_maxc = NODE_MAXC, node_blocks = 0;   // env NRX_NODE_MAXC (children per group, <= 16), NRX_NODE_BLOCKS (block target per launch; 0 = 24 x SMs)
  cudaStream_t copy_stream = nullptr;   // host -> device copies of staged alignments (overlap with the kernels on `stream`)
  cudaEvent_t ev_main = nullptr, ev_copy = nullptr;
  double *d_result_map = nullptr;  // device address of h_result (pinned, mapped): reducing kernels write their result there (result_out)
  bool k6_ring = false;            // env NRX_K6_RING=1: K6 streams the sumtable through a cp.async.bulk ring (k_derivatives_dna4r) instead of registers — measured SLOWER (0.62 vs 0.69 of the HBM peak in the config-2 sweep, gpurun_out/r4c_*), kept for the A/B
  bool zero_copy = true;           // env NRX_ZEROCOPY=0: results go to d_result and are copied
  bool score_only = false;         // nrx_set_score_only: replays of a fused-K3 plan do not store the root displayed trees' CLVs (scalers and per-site terms only)
  uint32_t quad_total = 0;         // env NRX_QUAD_BLOCKS: blocks per launch of the quad kernels (0: 2 per SM)
  bool quad = true;                // env NRX_QUAD=0: thread-per-pattern k_tree_lnl_dna4 / k_edge_lnl_dna4 / k_derivatives_dna4 instead of the coalesced quad kernels (A/B)
  int node_mode = 0;               // env NRX_NODE=1: ops of a node that share children run on k_clv_node_dna4 (A/B; measured 10 % slower than the per-op kernel, profiles/r3a_node_centric_ab.md)
  int walk_mode = 2;               // env NRX_WALK: 0 never, 1 whenever the plan has a tile-walk form, 2 (default) when it has one and the launch is small enough
  uint32_t walk_max_tiles = 0;     // mode 2: use the walk up to this many tiles per partition (env NRX_WALK_TILES; default set in nrx_engine_create)
  double *d_persite = nullptr;
  size_t persite_cap = 0;
  unsigned long long launches = 0;
  int sm_count = 148;
  bool capturing_pdl = false, pdl_prev_is_k2 = false;  // plan capture in progress with programmatic dependent launches (env NRX_PDL=0 disables)
  bool use_pdl = true;
  bool throughput_mode = false;  // several engines share the GPU (batched scoring): fewer, longer-running blocks per launch
  uint32_t pending_result = 0;  // doubles of an enqueued, not yet collected result (nrx_*_async / nrx_result_wait)
  uint32_t k2_nt = 2;       // env NRX_K2_NT: 64-pattern sub-tiles per ring stage of k_clv_dna4_pipe2 (1 or 2; 2 measured 1-2 % faster)
  int k2_variant = 0;       // 0: k_clv_dna4_pipe2 (production); 1: k_clv_dna4_pipe (A/B baseline, env NRX_K2=1)
  bool aa_generic = false;  // env NRX_AA=generic: force the scalar kernel for 20-state partitions (A/B)
  bool aa_pipe = true;      // software-pipelined MMA loop for the CLV update (env NRX_AA_PIPE=0: straight loop, A/B; the sumtable always pipelines)
  bool aa_v1 = false;       // env NRX_AA=v1: K2 / K5 on the round-1 kernel k_aa20_dmma instead of the warp-specialised k_aa20_mma (A/B)
  uint32_t aa_blocks = 148 * 3 * 2;  // block-count target of k_aa20_dmma: two waves of 3 resident blocks per SM (A/B: profiles/r1e_all_configs.md)
  uint32_t aa2_blocks = 0;           // block-count target of k_aa20_mma; 0 = by launch size: 2 waves of 2 resident blocks per SM for small launches, 4 for big ones (env NRX_AA2_BLOCKS) (env NRX_AA2_BLOCKS)
  uint32_t k2_blocks = 2368; // block-count target of the pipelined kernel: 8 waves of 2 resident blocks per SM (measured best, profiles/)
  cudaEvent_t t0 = nullptr, t1 = nullptr;
  bool views_dirty = true;
  // profiling of K2
  bool prof = false;
  struct ProfKind {
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> events;
    double ms = 0;
    unsigned long long launches = 0, units = 0, bytes = 0, cbytes = 0;   // bytes: algorithmic (SURVEY §8d per op); cbytes: compulsory (each distinct operand once per launch)
  };
  ProfKind profk[NRX_PROF_KINDS];  // per kernel family (nrx_engine.h: NRX_PROF_K2 ...), CUDA events on the engine stream
};

namespace {

/* Pinned host / device staging ring: small host->device payloads (ops, slot lists, branch lengths) are
 * bump-allocated so consecutive launches never wait for each other; the stream is synchronised only when
 * the ring wraps. */
int stage_alloc(nrx_engine *e, size_t bytes, void **hptr, void **dptr) {
  bytes = (bytes + 255) & ~(size_t)255;
  if (bytes > e->stage_cap) {
    size_t cap = std::max<size_t>(bytes * 2, 8 << 20);
    CK(cudaStreamSynchronize(e->stream));
    if (e->h_stage) cudaFreeHost(e->h_stage);
    if (e->d_stage) cudaFree(e->d_stage);
    e->h_stage = e->d_stage = nullptr;
    e->stage_cap = 0;
    CK(cudaMallocHost(&e->h_stage, cap));
    CK(cudaMalloc(&e->d_stage, cap));
    e->stage_cap = cap;
    e->stage_off = 0;
  }
  if (e->stage_off + bytes > e->stage_cap) {
    CK(cudaStreamSynchronize(e->stream));
    e->stage_off = 0;
  }
  *hptr = (char *)e->h_stage + e->stage_off;
  *dptr = (char *)e->d_stage + e->stage_off;
  e->stage_off += bytes;
  return 1;
}

int ensure_result(nrx_engine *e, size_t n_doubles, size_t n_partial) {
  if (n_doubles > e->result_cap) {
    CK(cudaStreamSynchronize(e->stream));
    if (e->d_result) cudaFree(e->d_result);
    if (e->h_result) cudaFreeHost(e->h_result);
    size_t cap = std::max<size_t>(n_doubles, 4096);
    CK(cudaMalloc((void **)&e->d_result, cap * sizeof(double)));
    CK(cudaMallocHost((void **)&e->h_result, cap * sizeof(double)));
    e->d_result_map = nullptr;
    { void *dp = nullptr; if (cudaHostGetDevicePointer(&dp, e->h_result, 0) == cudaSuccess) e->d_result_map = (double *)dp; else cudaGetLastError(); }
    if (e->d_tickets) cudaFree(e->d_tickets);
    CK(cudaMalloc((void **)&e->d_tickets, cap * sizeof(uint32_t)));
    CK(cudaMemset(e->d_tickets, 0, cap * sizeof(uint32_t)));
    e->result_cap = cap;
  }
  if (n_partial > e->partial_cap) {
    if (e->d_partial) cudaFree(e->d_partial);
    size_t cap = std::max<size_t>(n_partial, 1 << 16);
    CK(cudaMalloc((void **)&e->d_partial, cap * sizeof(double)));
    e->partial_cap = cap;
  }
  return 1;
}

/* stage host bytes -> device on the engine stream; returns the device pointer */
template <class T> int upload(nrx_engine *e, const T *src, size_t n, T **dev) {
  void *h, *d;
  if (!stage_alloc(e, n * sizeof(T), &h, &d)) return 0;
  std::memcpy(h, src, n * sizeof(T));
  CK(cudaMemcpyAsync(d, h, n * sizeof(T), cudaMemcpyHostToDevice, e->stream));
  *dev = (T *)d;
  return 1;
}

static void prof_begin(nrx_engine *e, cudaEvent_t *ev0, cudaEvent_t *ev1) {
  *ev0 = *ev1 = nullptr;
  if (e->prof) { cudaEventCreate(ev0); cudaEventCreate(ev1); cudaEventRecord(*ev0, e->stream); }
}
static void prof_end(nrx_engine *e, cudaEvent_t ev0, cudaEvent_t ev1, unsigned long long launches, unsigned long long updates, unsigned long long bytes,
                     int kind = NRX_PROF_K2, unsigned long long cbytes = ~0ull /* default: nothing shared, compulsory == algorithmic */) {
  if (!e->prof) return;
  cudaEventRecord(ev1, e->stream);
  nrx_engine::ProfKind &k = e->profk[kind];
  k.events.emplace_back(ev0, ev1);
  k.launches += launches;
  k.units += updates;
  k.bytes += bytes;
  k.cbytes += (cbytes == ~0ull) ? bytes : cbytes;
}
/* algorithmic bytes of the pattern-streaming kernels K3-K6: per (item, pattern) `clvs` CLV-sized streams + `extra` bytes */
static unsigned long long stream_bytes(const nrx_engine *e, unsigned long long items, unsigned clvs, unsigned extra, unsigned long long *units) {
  unsigned long long b = 0;
  for (const Part &p : e->parts) {
    b += items * p.d.patterns * ((unsigned long long)clvs * p.d.rate_cats * p.sp * 8 + extra);
    *units += items * p.d.patterns;
  }
  return b;
}

/* partition shapes served by the pipelined 4-state kernel k_clv_dna4_pipe2<NT, CATS> */
static bool dna_pipe_cats(uint32_t states, uint32_t cats) { return states == 4 && (cats == 1 || cats == 2 || cats == 4 || cats == 8 || cats == 16); }

PartView make_view(const Part &p, uint32_t index) {
  PartView v{};
  v.states = p.d.states; v.sp = p.sp; v.cats = p.d.rate_cats; v.patterns = p.d.patterns; v.tips = p.d.tips; v.edges = p.d.edges;
  v.part_index = index; v.tip_pitch = p.pat_pad; v.tip_codes = p.tip_codes;
  v.pmat = p.pmat; v.pmat_pad = p.pmat_pad; v.tipchars = p.tipchars; v.tipmap = p.tipmap; v.weights = p.weights;
  v.freqs = p.freqs; v.eigenvecs = p.eigenvecs; v.inv_eigenvecs = p.inv_eigenvecs; v.eigenvals = p.eigenvals;
  v.rates = p.rates; v.rate_weights = p.rate_weights;
  v.clv = p.d_clv; v.scaler = p.d_scaler; v.sumtable = p.d_sumtable; v.diagp = p.diagp; v.tiplut = p.tiplut;
  v.summat = p.summat; v.sumlut = p.sumlut;
  v.pinv = p.pinv; v.invariant = p.invariant;
  v.nmodels = p.nmodels;
  std::memcpy(v.cat_model, p.cat_model, sizeof(v.cat_model));
  return v;
}

int refresh_views(nrx_engine *e) {
  if (!e->views_dirty) return 1;
  for (ShapeClass &c : e->classes) {
    std::vector<PartView> hv;
    for (uint32_t pi : c.parts) hv.push_back(make_view(e->parts[pi], pi));
    if (!c.d_views) CK(cudaMalloc((void **)&c.d_views, hv.size() * sizeof(PartView)));
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaMemcpy(c.d_views, hv.data(), hv.size() * sizeof(PartView), cudaMemcpyHostToDevice));
  }
  e->views_dirty = false;
  return 1;
}

int check_part(nrx_engine *e, uint32_t p) {
  if (!e) { g_err = "null engine"; return 0; }
  if (p >= e->parts.size()) { g_err = "partition index out of range"; return 0; }
  return 1;
}

uint32_t class_tip_codes(const nrx_engine *e, const ShapeClass &c);

/* K1b for a list of edges (nullptr = all edges) of a 20-state partition */
int refresh_tiplut(nrx_engine *e, uint32_t pi, const uint32_t *edges, uint32_t n) {
  Part &p = e->parts[pi];
  if (!p.tiplut || p.tip_codes > (uint32_t)AA_LUT_CODES) return 1;
  std::vector<uint32_t> all;
  if (!edges) { all.resize(p.d.edges); for (uint32_t i = 0; i < p.d.edges; ++i) all[i] = i; edges = all.data(); n = p.d.edges; }
  if (n == 0) return 1;
  uint32_t *d_idx;
  if (!upload(e, edges, n, &d_idx)) return 0;
  k_tip_lut20<<<n, 256, 0, e->stream>>>(make_view(p, pi), p.tiplut, d_idx);
  e->launches++;
  CK(cudaGetLastError());
  return 1;
}
/* pll_update_invariant_sites (LIBPLL/models.c:651-760): AND of all tips' state masks per pattern; exactly one state left ->
 * its index, otherwise -1.  masks[t * patterns + i] = state mask of tip t at pattern i. */
template <class T> int upload_invariant(nrx_engine *e, Part &p, const T *masks, const uint32_t *tipmap) {
  const size_t n = p.d.patterns;
  std::vector<int> inv(std::max<size_t>(1, n), -1);
  const uint32_t gap = (p.d.states >= 32) ? 0xffffffffu : ((1u << p.d.states) - 1);
  for (size_t i = 0; i < n; ++i) {
    uint32_t st = gap;
    for (uint32_t t = 0; t < p.d.tips; ++t) st &= tipmap ? tipmap[masks[(size_t)t * n + i]] : (uint32_t)masks[(size_t)t * n + i];
    inv[i] = (st == 0 || __builtin_popcount(st) > 1) ? -1 : __builtin_ctz(st);
  }
  CK(cudaStreamSynchronize(e->stream));
  if (n) CK(cudaMemcpy(p.invariant, inv.data(), n * sizeof(int), cudaMemcpyHostToDevice));
  return 1;
}

/* K5 operands of a 20-state partition: A_L[j][k] = pi_k Vinv[k][j], A_R[j][k] = V[j][k] and the tip table
 * sum_{k in code} pi_k Vinv[k][j] (serial k order as the reference's tip-inner sumtable loop, LIBPLL/core_derivatives.c:473-641) */
int refresh_summat(nrx_engine *e, uint32_t pi) {
  Part &p = e->parts[pi];
  if (!p.summat || !p.model_set) return 1;
  const uint32_t S = 20, SP = p.sp;
  std::vector<double> m(800, 0.0), lut((size_t)AA_LUT_CODES * 80, 0.0);
  for (uint32_t j = 0; j < S; ++j)
    for (uint32_t k = 0; k < S; ++k) {
      m[j * 20 + k] = p.h_freqs[k] * p.h_inv_eigenvecs[k * SP + j];
      m[400 + j * 20 + k] = p.h_eigenvecs[j * SP + k];
    }
  if (p.tips_set && p.tip_codes <= (uint32_t)AA_LUT_CODES)
    for (uint32_t code = 0; code < p.tip_codes; ++code)
      for (uint32_t j = 0; j < S; ++j) {
        double sum = 0.0;
        for (uint32_t k = 0; k < S; ++k) if ((p.h_tipmap[code] >> k) & 1u) sum += p.h_freqs[k] * p.h_inv_eigenvecs[k * SP + j];
        for (uint32_t c = 0; c < 4; ++c) lut[(size_t)code * 80 + c * 20 + j] = sum;
      }
  CK(cudaStreamSynchronize(e->stream));
  CK(cudaMemcpy(p.summat, m.data(), m.size() * sizeof(double), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(p.sumlut, lut.data(), lut.size() * sizeof(double), cudaMemcpyHostToDevice));
  return 1;
}
uint32_t tiles_for(uint64_t items, uint32_t per_block, uint32_t other_dims) {
  // enough blocks for >= ~8 waves over 148 SMs when the launch is big, one tile per block when it is small
  uint64_t full = (items + per_block - 1) / per_block;
  if (full == 0) full = 1;
  uint64_t want = full;
  const uint64_t target_blocks = 148ull * 8 * 8;
  if (full * other_dims > target_blocks) {
    want = std::max<uint64_t>(1, target_blocks / std::max<uint32_t>(1, other_dims));
    want = std::min<uint64_t>(want, full);
    // never give a block more than 16 tiles: keeps the tail short
    want = std::max<uint64_t>(want, (full + 15) / 16);
  }
  return (uint32_t)std::min<uint64_t>(want, 65535ull * 32);
}

}  // namespace

namespace {
uint32_t class_tip_codes(const nrx_engine *e, const ShapeClass &c) {
  uint32_t m = 0;
  for (uint32_t pi : c.parts) m = std::max(m, e->parts[pi].tip_codes);
  return m;
}
}  // namespace

extern "C" {

static int flush_pmatrices(nrx_engine *e);
static bool aa_dmma_class(const nrx_engine *e, const ShapeClass &c);

const char *nrx_last_error(void) { return g_err.c_str(); }

int nrx_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

nrx_engine *nrx_engine_create(const nrx_partition_desc *descs, uint32_t nparts, int device) {
  int ndev = 0;
  cudaError_t err = cudaGetDeviceCount(&ndev);
  if (err != cudaSuccess || ndev == 0) {
    g_err = std::string("nrx_engine_create: no usable CUDA device (") + cudaGetErrorString(err) + "); this engine has no CPU fallback";
    cudaGetLastError();
    return nullptr;
  }
  if (device < 0 || device >= ndev) { g_err = "nrx_engine_create: bad device index"; return nullptr; }
  if (!cuda_ok(cudaSetDevice(device), "cudaSetDevice")) return nullptr;
  nrx_engine *e = new nrx_engine();
  e->device = device;
  if (!cuda_ok(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking), "cudaStreamCreate")) { delete e; return nullptr; }
  if (!cuda_ok(cudaHostAlloc((void **)&e->h_err, sizeof(int), cudaHostAllocMapped), "cudaHostAlloc")) { delete e; return nullptr; }
  *e->h_err = 0;
  e->parts.resize(nparts);
  if (const char *v = std::getenv("NRX_K2")) e->k2_variant = std::atoi(v);
  if (const char *v = std::getenv("NRX_GRAPH")) e->use_graphs = std::atoi(v) != 0;
  if (const char *v = std::getenv("NRX_PDL")) e->use_pdl = std::atoi(v) != 0;
  if (const char *v = std::getenv("NRX_FUSE_REDUCE")) e->fuse_reduce = std::atoi(v) != 0;
  if (const char *v = std::getenv("NRX_ZEROCOPY")) e->zero_copy = std::atoi(v) != 0;
  if (const char *v = std::getenv("NRX_K6_RING")) e->k6_ring = std::atoi(v) != 0;
  if (const char *v = std::getenv("NRX_QUAD")) e->quad = std::atoi(v) != 0;
  if (const char *v = std::getenv("NRX_QUAD_BLOCKS")) e->quad_total = (uint32_t)std::max(0, std::atoi(v));
  if (const char *v = std::getenv("NRX_NODE")) e->node_mode = std::atoi(v);
  if (const char *v = std::getenv("NRX_NODE_MAXC")) e->node_maxc = (uint32_t)std::min(NODE_MAXC, std::max(2, std::atoi(v)));
  if (const char *v = std::getenv("NRX_NODE_BLOCKS")) e->node_blocks = (uint32_t)std::max(0, std::atoi(v));
  if (const char *v = std::getenv("NRX_WALK")) e->walk_mode = std::atoi(v);
  if (const char *v = std::getenv("NRX_WALK_TILES")) e->walk_max_tiles = (uint32_t)std::max(0, std::atoi(v));
  if (const char *v = std::getenv("NRX_K2_NT")) e->k2_nt = std::atoi(v) == 1 ? 1u : 2u;
  if (const char *v = std::getenv("NRX_AA")) { e->aa_generic = std::string(v) == "generic"; if (std::string(v) == "v2") e->aa_v1 = false; if (std::string(v) == "v1") e->aa_v1 = true; }
  if (const char *v = std::getenv("NRX_AA_PIPE")) e->aa_pipe = std::atoi(v) != 0;
  if (const char *v = std::getenv("NRX_AA2_BLOCKS")) e->aa2_blocks = (uint32_t)std::max(1, std::atoi(v));
  if (const char *v = std::getenv("NRX_AA_BLOCKS")) e->aa_blocks = (uint32_t)std::max(1, std::atoi(v));
  if (const char *v = std::getenv("NRX_K2_BLOCKS")) e->k2_blocks = (uint32_t)std::max(1, std::atoi(v));
  {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    e->sm_count = sms;
    if (!std::getenv("NRX_K2_BLOCKS")) e->k2_blocks = 16u * (uint32_t)sms;
    // measured (profiles/r2r_tile_walk.md): 10 k patterns (313 tiles) 73 vs 89 us per evaluation; at 100 k patterns the walk's serial
    // op chain per block loses 3x to the bandwidth-bound level-by-level kernels -> only while the launch is latency-bound
    if (!std::getenv("NRX_WALK_TILES")) e->walk_max_tiles = 4u * (uint32_t)sms;
    if (!cuda_ok(cudaFuncSetAttribute(k_walk_dna4, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024), "cudaFuncSetAttribute") ||
        !cuda_ok(cudaFuncSetAttribute(k_derivatives_dna4r, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(K6RingSmem)), "cudaFuncSetAttribute") ||
        !cuda_ok(cudaFuncSetAttribute(k_clv_node_dna4, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)node_smem_bytes(NODE_MAXC)), "cudaFuncSetAttribute") ||
        !cuda_ok(cudaFuncSetAttribute(k_clv_node_dna4, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared), "cudaFuncSetAttribute")) { delete e; return nullptr; }
    const int aa_smem = (int)(sizeof(AaSmem) + 2 * AA_LUT_CODES * 80 * sizeof(double));
    if (!cuda_ok(cudaFuncSetAttribute(k_aa20_dmma<AA_CLV>, cudaFuncAttributeMaxDynamicSharedMemorySize, aa_smem), "cudaFuncSetAttribute") ||
        !cuda_ok(cudaFuncSetAttribute(k_aa20_dmma<AA_SUM>, cudaFuncAttributeMaxDynamicSharedMemorySize, aa_smem), "cudaFuncSetAttribute") ||
        !cuda_ok(cudaFuncSetAttribute(k_aa20_dmma<AA_EDGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, aa_smem), "cudaFuncSetAttribute")) { delete e; return nullptr; }
    const int aa2_smem = (int)(sizeof(AaSmem2) + 2 * AA_LUT_CODES * 80 * sizeof(double));
    if (!cuda_ok(cudaFuncSetAttribute(k_aa20_mma<AA_CLV, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, aa2_smem), "cudaFuncSetAttribute") ||
        !cuda_ok(cudaFuncSetAttribute(k_aa20_mma<AA_CLV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, aa2_smem), "cudaFuncSetAttribute") ||
        !cuda_ok(cudaFuncSetAttribute(k_aa20_mma<AA_SUM, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, aa2_smem), "cudaFuncSetAttribute") ||
        !cuda_ok(cudaFuncSetAttribute(k_aa20_mma<AA_EDGE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, aa2_smem), "cudaFuncSetAttribute")) { delete e; return nullptr; }
    if (!cuda_ok(cudaFuncSetAttribute(k_clv_dna4_pipe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ClvPipeSmem)), "cudaFuncSetAttribute") ||
        !cuda_ok(cudaFuncSetAttribute(k_clv_dna4_pipe2<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PipeSmem<1, 4>)), "cudaFuncSetAttribute") ||
        !cuda_ok(cudaFuncSetAttribute(k_clv_dna4_pipe2<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PipeSmem<2, 4>)), "cudaFuncSetAttribute") ||
        !cuda_ok(cudaFuncSetAttribute(k_clv_dna4_pipe2<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PipeSmem<2, 1>)), "cudaFuncSetAttribute") ||
        !cuda_ok(cudaFuncSetAttribute(k_clv_dna4_pipe2<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PipeSmem<2, 2>)), "cudaFuncSetAttribute") ||
        !cuda_ok(cudaFuncSetAttribute(k_clv_dna4_pipe2<2, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PipeSmem<2, 8>)), "cudaFuncSetAttribute") ||
        !cuda_ok(cudaFuncSetAttribute(k_clv_dna4_pipe2<2, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PipeSmem<2, 16>)), "cudaFuncSetAttribute")) { delete e; return nullptr; }
  }
  for (uint32_t i = 0; i < nparts; ++i) {
    Part &p = e->parts[i];
    p.d = descs[i];
    if (p.d.states < 2 || p.d.states > 32 || p.d.rate_cats < 1 || p.d.rate_cats > 16) { g_err = "unsupported states / rate_cats"; nrx_engine_destroy(e); return nullptr; }
    p.sp = (p.d.states + 3) & ~3u;
    p.pat_pad = (p.d.patterns + 511) / 512 * 512;   // whole tiles of every kernel (k_clv_dna4_pipe2 with one rate category: 512 patterns per stage): bulk copies always move full tiles
    p.clv_entries = (size_t)p.d.patterns * p.d.rate_cats * p.sp;
    p.pmat_entries = (size_t)p.d.rate_cats * p.d.states * p.sp;
    e->max_patterns = std::max(e->max_patterns, p.d.patterns);
    const size_t S = p.d.states, SP = p.sp, C = p.d.rate_cats;
    const size_t model_doubles = SP + 2 * S * SP + SP + 2 * C + C * S * 4;
    bool ok = cuda_ok(cudaMalloc((void **)&p.pmat, std::max<size_t>(1, p.d.edges * p.pmat_entries) * sizeof(double)), "cudaMalloc pmat") &&
              cuda_ok(cudaMalloc((void **)&p.tipchars, std::max<size_t>(1, (size_t)p.d.tips * p.pat_pad)), "cudaMalloc tipchars") &&
              cuda_ok(cudaMalloc((void **)&p.tipmap, 256 * sizeof(uint32_t)), "cudaMalloc tipmap") &&
              cuda_ok(cudaMalloc((void **)&p.weights, std::max<size_t>(1, p.pat_pad) * sizeof(uint32_t)), "cudaMalloc weights") &&
              cuda_ok(cudaMalloc((void **)&p.model, model_doubles * sizeof(double)), "cudaMalloc model") &&
              cuda_ok(cudaMalloc((void **)&p.invariant, std::max<size_t>(1, p.pat_pad) * sizeof(int)), "cudaMalloc invariant");
    if (ok && p.d.states == 4 && p.d.rate_cats == 4)
      ok = cuda_ok(cudaMalloc((void **)&p.pmat_pad, std::max<size_t>(1, (size_t)p.d.edges * WALK_PE) * sizeof(double)), "cudaMalloc pmat_pad") &&
           cuda_ok(cudaMemset(p.pmat_pad, 0, std::max<size_t>(1, (size_t)p.d.edges * WALK_PE) * sizeof(double)), "cudaMemset pmat_pad");
    if (ok && p.d.states == 20 && p.d.rate_cats == 4)
      ok = cuda_ok(cudaMalloc((void **)&p.tiplut, (size_t)p.d.edges * AA_LUT_CODES * 80 * sizeof(double)), "cudaMalloc tiplut") &&
           cuda_ok(cudaMalloc((void **)&p.summat, 800 * sizeof(double)), "cudaMalloc summat") &&
           cuda_ok(cudaMalloc((void **)&p.sumlut, (size_t)AA_LUT_CODES * 80 * sizeof(double)), "cudaMalloc sumlut");
    if (!ok) { nrx_engine_destroy(e); return nullptr; }
    p.freqs = p.model; p.eigenvecs = p.freqs + SP; p.inv_eigenvecs = p.eigenvecs + S * SP; p.eigenvals = p.inv_eigenvecs + S * SP;
    p.rates = p.eigenvals + SP; p.rate_weights = p.rates + C; p.diagp = p.rate_weights + C;
    std::vector<uint32_t> ones(std::max<uint32_t>(1, p.pat_pad), 1);
    cudaMemcpy(p.weights, ones.data(), ones.size() * sizeof(uint32_t), cudaMemcpyHostToDevice);
    cudaMemset(p.pmat, 0, std::max<size_t>(1, p.d.edges * p.pmat_entries) * sizeof(double));
    cudaMemset(p.tipchars, 0, std::max<size_t>(1, (size_t)p.d.tips * p.pat_pad));
    // group by kernel shape
    bool found = false;
    for (ShapeClass &c : e->classes)
      if (c.states == p.d.states && c.cats == p.d.rate_cats) { c.parts.push_back(i); c.max_patterns = std::max(c.max_patterns, p.d.patterns); found = true; }
    if (!found) { ShapeClass c; c.states = p.d.states; c.cats = p.d.rate_cats; c.parts = {i}; c.max_patterns = p.d.patterns; e->classes.push_back(c); }
  }
  e->defer_pmat = e->walk_mode != 0 && e->classes.size() == 1 && e->classes[0].states == 4 && e->classes[0].cats == 4 && !std::getenv("NRX_NO_DEFER_PMAT");
  for (Part &p : e->parts) { p.h_len.assign(p.d.edges, -1.0); p.pend.assign(p.d.edges, 0); }
  return e;
}

void nrx_engine_destroy(nrx_engine *e) {
  if (!e) return;
  cudaSetDevice(e->device);
  if (e->stream) cudaStreamSynchronize(e->stream);
  if (e->copy_stream) { cudaStreamSynchronize(e->copy_stream); cudaStreamDestroy(e->copy_stream); cudaEventDestroy(e->ev_main); cudaEventDestroy(e->ev_copy); }
  if (e->comm && nccl().ok) nccl().CommDestroy(e->comm);
  for (EnginePlan &pl : e->plans) { for (cudaGraphExec_t &x : pl.exec) if (x) cudaGraphExecDestroy(x); cudaFree(pl.d_ops); cudaFree(pl.d_walk); cudaFree(pl.d_ngroups); cudaFree(pl.d_nops); cudaFree(pl.d_nblocks); }
  for (Part &p : e->parts) {
    cudaFree(p.invariant); cudaFree(p.pmat); cudaFree(p.pmat_pad); cudaFree(p.tiplut); cudaFree(p.summat); cudaFree(p.sumlut); cudaFree(p.tipchars); cudaFree(p.tipmap); cudaFree(p.weights); cudaFree(p.model);
    cudaFree(p.tipchars_next); cudaFree(p.weights_next);
    for (void *m : p.slot_mem) cudaFree(m);
    for (double *m : p.h_sumtable) cudaFree(m);
    cudaFree(p.d_clv); cudaFree(p.d_scaler); cudaFree(p.d_sumtable);
  }
  for (ShapeClass &c : e->classes) cudaFree(c.d_views);
  for (nrx_engine::ProfKind &k : e->profk) for (auto &ev : k.events) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
  if (e->h_stage) cudaFreeHost(e->h_stage);
  cudaFree(e->d_fused);
  cudaFree(e->d_stage); cudaFree(e->d_partial); cudaFree(e->d_result); cudaFree(e->d_persite); cudaFree(e->d_tickets);
  if (e->h_result) cudaFreeHost(e->h_result);
  if (e->h_err) cudaFreeHost(e->h_err);
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
}

int nrx_set_tips(nrx_engine *e, uint32_t pi, const uint32_t *tip_masks) {
  if (!check_part(e, pi)) return 0;
  CK(cudaSetDevice(e->device));
  Part &p = e->parts[pi];
  const size_t n = (size_t)p.d.tips * p.d.patterns;
  std::vector<uint8_t> codes(std::max<size_t>(1, n));
  std::vector<uint32_t> tipmap(256, 0);
  const uint32_t full = (p.d.states >= 32) ? 0xffffffffu : ((1u << p.d.states) - 1);
  if (p.d.states == 4) {  // the code IS the mask (set_tipchars_4x4, LIBPLL/pll.c:875-900)
    for (uint32_t i = 0; i < 16; ++i) tipmap[i] = i;
    for (size_t i = 0; i < n; ++i) {
      if (tip_masks[i] == 0 || tip_masks[i] > 15) { g_err = "Illegal state code in tip"; return 0; }
      codes[i] = (uint8_t)tip_masks[i];
    }
  } else {
    std::map<uint32_t, uint32_t> code;
    for (size_t i = 0; i < n; ++i) {
      const uint32_t m = tip_masks[i];
      if (m == 0 || (m & ~full)) { g_err = "Illegal state code in tip"; return 0; }
      auto it = code.find(m);
      if (it == code.end()) {
        if (code.size() >= 256) { g_err = "more than 256 distinct tip states"; return 0; }
        const uint32_t c = (uint32_t)code.size();
        it = code.emplace(m, c).first;
        tipmap[c] = m;
      }
      codes[i] = (uint8_t)it->second;
    }
  }
  CK(cudaStreamSynchronize(e->stream));
  if (n) CK(cudaMemcpy2D(p.tipchars, p.pat_pad, codes.data(), p.d.patterns, p.d.patterns, p.d.tips, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(p.tipmap, tipmap.data(), 256 * sizeof(uint32_t), cudaMemcpyHostToDevice));
  p.tip_codes = 0;
  for (uint32_t i = 0; i < 256; ++i) if (tipmap[i]) p.tip_codes = i + 1;
  p.tips_set = true;
  p.h_tipmap = tipmap;
  if (!upload_invariant(e, p, codes.data(), tipmap.data())) return 0;
  e->views_dirty = true;
  if (p.model_set && (!refresh_tiplut(e, pi, nullptr, 0) || !refresh_summat(e, pi))) return 0;  // the code -> state-set map may have changed
  return 1;
}

/* Asynchronous tip upload: codes [tips][patterns] (1 byte per cell) + the code -> state-set map.  Everything is enqueued on the
 * engine stream — the 2-D copy into the padded rows, the map, and k_check_tipchars, which validates every code and rebuilds the
 * invariant-site table on the device — and the call returns; `codes` is borrowed until the engine next synchronises
 * (nrx_result_wait / nrx_sync; pageable memory is staged by the driver before the call returns).  An illegal code surfaces as
 * a failure of that synchronising call. */
static int set_tipcodes_async(nrx_engine *e, uint32_t pi, const uint8_t *codes, const uint32_t *tipmap256, uint32_t ncodes) {
  Part &p = e->parts[pi];
  const size_t n = (size_t)p.d.tips * p.d.patterns;
  const uint32_t full = (p.d.states >= 32) ? 0xffffffffu : ((1u << p.d.states) - 1);
  std::vector<uint32_t> tipmap(tipmap256, tipmap256 + 256);
  const bool map_changed = !p.tips_set || tipmap != p.h_tipmap;
  if (n) CK(cudaMemcpy2DAsync(p.tipchars, p.pat_pad, codes, p.d.patterns, p.d.patterns, p.d.tips, cudaMemcpyHostToDevice, e->stream));
  if (map_changed) {
    uint32_t *d_map;
    if (!upload(e, tipmap.data(), 256, &d_map)) return 0;
    CK(cudaMemcpyAsync(p.tipmap, d_map, 256 * sizeof(uint32_t), cudaMemcpyDeviceToDevice, e->stream));
    p.h_tipmap = tipmap;
    if (p.tip_codes != ncodes) { p.tip_codes = ncodes; e->views_dirty = true; }
  }
  if (p.d.patterns) {
    const uint32_t blocks = (uint32_t)std::min<uint64_t>(((uint64_t)p.d.patterns + BLOCK - 1) / BLOCK, 8ull * e->sm_count);
    k_check_tipchars<<<blocks, BLOCK, 0, e->stream>>>(p.tipchars, p.d.tips, p.d.patterns, p.pat_pad, p.tipmap, full, p.invariant, e->h_err);
    e->launches++;
    CK(cudaGetLastError());
  }
  p.tips_set = true;
  p.invariant_stale = false;
  if (map_changed && p.model_set && (!refresh_tiplut(e, pi, nullptr, 0) || !refresh_summat(e, pi))) return 0;   // 20 states: the tables are per code
  return 1;
}

int nrx_set_tipchars_u8(nrx_engine *e, uint32_t pi, const uint8_t *codes) {
  if (!check_part(e, pi)) return 0;
  CK(cudaSetDevice(e->device));
  if (e->parts[pi].d.states != 4) { g_err = "nrx_set_tipchars_u8: only 4-state partitions store the mask as the code (use nrx_set_tipcodes_u8)"; return 0; }
  uint32_t tipmap[256] = {};
  for (uint32_t i = 1; i < 16; ++i) tipmap[i] = i;   // code == state mask; 0 and everything above 15 is illegal
  return set_tipcodes_async(e, pi, codes, tipmap, 16);
}

int nrx_set_tipcodes_u8(nrx_engine *e, uint32_t pi, const uint8_t *codes, const uint32_t *tipmap, uint32_t ncodes) {
  if (!check_part(e, pi)) return 0;
  CK(cudaSetDevice(e->device));
  Part &p = e->parts[pi];
  if (ncodes == 0 || ncodes > 256) { g_err = "nrx_set_tipcodes_u8: 1..256 codes"; return 0; }
  const uint32_t full = (p.d.states >= 32) ? 0xffffffffu : ((1u << p.d.states) - 1);
  uint32_t map[256] = {};
  for (uint32_t i = 0; i < ncodes; ++i) {
    if (tipmap[i] == 0 || (tipmap[i] & ~full)) { g_err = "Illegal state code in tip"; return 0; }
    map[i] = tipmap[i];
  }
  return set_tipcodes_async(e, pi, codes, map, ncodes);
}

int nrx_set_pattern_weights_async(nrx_engine *e, uint32_t pi, const uint32_t *w) {
  if (!check_part(e, pi)) return 0;
  CK(cudaSetDevice(e->device));
  Part &p = e->parts[pi];
  if (p.d.patterns) CK(cudaMemcpyAsync(p.weights, w, (size_t)p.d.patterns * sizeof(uint32_t), cudaMemcpyHostToDevice, e->stream));
  return 1;
}

/* Double-buffered alignment upload (4-state partitions, code == state mask).  stage: the copies go to the shadow buffers on the copy
 * stream, ordered after everything already enqueued on the engine stream (the shadow buffers were the live ones of the previous
 * step) — so they overlap with whatever the engine stream is given AFTER this call.  commit: the engine stream waits for the copies,
 * the buffers swap, the codes are validated and the invariant-site table rebuilt as in nrx_set_tipchars_u8. */
int nrx_stage_alignment_u8(nrx_engine *e, uint32_t pi, const uint8_t *codes, const uint32_t *weights) {
  if (!check_part(e, pi)) return 0;
  CK(cudaSetDevice(e->device));
  Part &p = e->parts[pi];
  if (p.d.states != 4) { g_err = "nrx_stage_alignment_u8: only 4-state partitions store the mask as the code"; return 0; }
  if (!e->copy_stream) {
    CK(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&e->ev_main, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&e->ev_copy, cudaEventDisableTiming));
  }
  if (codes && !p.tipchars_next) {
    CK(cudaMalloc((void **)&p.tipchars_next, std::max<size_t>(1, (size_t)p.d.tips * p.pat_pad)));
    CK(cudaMemset(p.tipchars_next, 0, std::max<size_t>(1, (size_t)p.d.tips * p.pat_pad)));
  }
  if (weights && !p.weights_next) {
    std::vector<uint32_t> ones(std::max<uint32_t>(1, p.pat_pad), 1);
    CK(cudaMalloc((void **)&p.weights_next, ones.size() * sizeof(uint32_t)));
    CK(cudaMemcpy(p.weights_next, ones.data(), ones.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
  }
  CK(cudaEventRecord(e->ev_main, e->stream));
  CK(cudaStreamWaitEvent(e->copy_stream, e->ev_main, 0));
  const size_t n = (size_t)p.d.tips * p.d.patterns;
  if (codes && n) CK(cudaMemcpy2DAsync(p.tipchars_next, p.pat_pad, codes, p.d.patterns, p.d.patterns, p.d.tips, cudaMemcpyHostToDevice, e->copy_stream));
  if (weights && p.d.patterns) CK(cudaMemcpyAsync(p.weights_next, weights, (size_t)p.d.patterns * sizeof(uint32_t), cudaMemcpyHostToDevice, e->copy_stream));
  p.staged_tips = p.staged_tips || codes != nullptr;
  p.staged_weights = p.staged_weights || weights != nullptr;
  return 1;
}

int nrx_commit_staged_alignment(nrx_engine *e) {
  if (!e) { g_err = "null engine"; return 0; }
  CK(cudaSetDevice(e->device));
  bool any = false;
  for (const Part &p : e->parts) any = any || p.staged_tips || p.staged_weights;
  if (!any) return 1;
  CK(cudaEventRecord(e->ev_copy, e->copy_stream));
  CK(cudaStreamWaitEvent(e->stream, e->ev_copy, 0));
  for (uint32_t pi = 0; pi < e->parts.size(); ++pi) {
    Part &p = e->parts[pi];
    if (p.staged_weights) { std::swap(p.weights, p.weights_next); p.staged_weights = false; e->views_dirty = true; }
    if (!p.staged_tips) continue;
    std::swap(p.tipchars, p.tipchars_next);
    p.staged_tips = false;
    e->views_dirty = true;
    uint32_t tipmap[256] = {};
    for (uint32_t i = 1; i < 16; ++i) tipmap[i] = i;
    std::vector<uint32_t> tm(tipmap, tipmap + 256);
    if (!p.tips_set || tm != p.h_tipmap) {
      uint32_t *d_map;
      if (!upload(e, tm.data(), 256, &d_map)) return 0;
      CK(cudaMemcpyAsync(p.tipmap, d_map, 256 * sizeof(uint32_t), cudaMemcpyDeviceToDevice, e->stream));
      p.h_tipmap = tm;
      if (p.tip_codes != 16) { p.tip_codes = 16; }
    }
    if (p.d.patterns) {
      const uint32_t blocks = (uint32_t)std::min<uint64_t>(((uint64_t)p.d.patterns + BLOCK - 1) / BLOCK, 8ull * e->sm_count);
      k_check_tipchars<<<blocks, BLOCK, 0, e->stream>>>(p.tipchars, p.d.tips, p.d.patterns, p.pat_pad, p.tipmap, 15u, p.invariant, e->h_err);
      e->launches++;
      CK(cudaGetLastError());
    }
    p.tips_set = true;
    p.invariant_stale = false;
  }
  return 1;
}

int nrx_set_pattern_weights(nrx_engine *e, uint32_t pi, const uint32_t *w) {
  if (!check_part(e, pi)) return 0;
  CK(cudaSetDevice(e->device));
  Part &p = e->parts[pi];
  CK(cudaStreamSynchronize(e->stream));
  if (p.d.patterns) CK(cudaMemcpy(p.weights, w, (size_t)p.d.patterns * sizeof(uint32_t), cudaMemcpyHostToDevice));
  return 1;
}

/* Model of partition pi with `nmodels` rate matrices: category c uses matrix cat_model[c] (libpll's params_indices; LG4M / LG4X).
 * freqs / eigenvals: [nmodels][states_padded]; eigenvecs / inv_eigenvecs: [nmodels][states][states_padded]. */
int nrx_set_model_mixture(nrx_engine *e, uint32_t pi, uint32_t nmodels, const uint32_t *cat_model, const double *freqs,
                          const double *eigenvecs, const double *inv_eigenvecs, const double *eigenvals, const double *rates,
                          const double *rate_weights, double prop_invar) {
  if (!check_part(e, pi)) return 0;
  if (!(prop_invar >= 0.0 && prop_invar < 1.0)) { g_err = "Invalid proportion of invariant sites"; return 0; }   // pll_update_invariant_sites_proportion
  CK(cudaSetDevice(e->device));
  if (!flush_pmatrices(e)) return 0;
  Part &p = e->parts[pi];
  const size_t S = p.d.states, SP = p.sp, C = p.d.rate_cats, M = nmodels;
  if (M < 1 || M > 16) { g_err = "nrx_set_model_mixture: 1..16 rate matrices"; return 0; }
  for (size_t c = 0; M > 1 && c < C; ++c)
    if (!cat_model || cat_model[c] >= M) { g_err = "nrx_set_model_mixture: rate-matrix index of a category out of range"; return 0; }
  CK(cudaStreamSynchronize(e->stream));
  if (M > p.model_cap) {   // grow the model buffer: freqs | eigenvecs | inv_eigenvecs | eigenvals (M blocks each) | rates | rate_weights | diagp
    double *nm = nullptr;
    CK(cudaMalloc((void **)&nm, (M * (2 * SP + 2 * S * SP) + 2 * C + C * S * 4) * sizeof(double)));
    cudaFree(p.model);
    p.model = nm;
    p.model_cap = (uint32_t)M;
    e->views_dirty = true;
  }
  const size_t Mc = p.model_cap;
  p.freqs = p.model; p.eigenvecs = p.freqs + Mc * SP; p.inv_eigenvecs = p.eigenvecs + Mc * S * SP; p.eigenvals = p.inv_eigenvecs + Mc * S * SP;
  p.rates = p.eigenvals + Mc * SP; p.rate_weights = p.rates + C; p.diagp = p.rate_weights + C;
  uint8_t cm[16] = {};
  for (size_t c = 0; M > 1 && c < C; ++c) cm[c] = (uint8_t)cat_model[c];
  if (p.pinv != prop_invar || p.nmodels != M || std::memcmp(cm, p.cat_model, sizeof(cm)) != 0) e->views_dirty = true;
  p.pinv = prop_invar;
  p.nmodels = (uint32_t)M;
  std::memcpy(p.cat_model, cm, sizeof(cm));
  std::vector<double> h((size_t)(p.rate_weights + C - p.model), 0.0);
  for (size_t m = 0; m < M; ++m) {
    std::memcpy(h.data() + m * SP, freqs + m * SP, S * sizeof(double));
    std::memcpy(h.data() + (p.eigenvecs - p.model) + m * S * SP, eigenvecs + m * S * SP, S * SP * sizeof(double));
    std::memcpy(h.data() + (p.inv_eigenvecs - p.model) + m * S * SP, inv_eigenvecs + m * S * SP, S * SP * sizeof(double));
    std::memcpy(h.data() + (p.eigenvals - p.model) + m * SP, eigenvals + m * SP, S * sizeof(double));
  }
  std::memcpy(h.data() + (p.rates - p.model), rates, C * sizeof(double));
  std::memcpy(h.data() + (p.rate_weights - p.model), rate_weights, C * sizeof(double));
  CK(cudaMemcpy(p.model, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice));
  p.h_eigenvals.assign(M * S, 0.0);
  for (size_t m = 0; m < M; ++m) std::memcpy(p.h_eigenvals.data() + m * S, eigenvals + m * SP, S * sizeof(double));
  p.h_rates.assign(rates, rates + C);
  p.h_freqs.assign(freqs, freqs + S);   // matrix 0: the single-matrix tables of the 20-state tensor-core K5 (unused by mixtures)
  p.h_eigenvecs.assign(eigenvecs, eigenvecs + S * SP);
  p.h_inv_eigenvecs.assign(inv_eigenvecs, inv_eigenvecs + S * SP);
  p.model_set = true;
  return refresh_summat(e, pi);
}

int nrx_set_model(nrx_engine *e, uint32_t pi, const double *freqs, const double *eigenvecs, const double *inv_eigenvecs,
                  const double *eigenvals, const double *rates, const double *rate_weights, double prop_invar) {
  return nrx_set_model_mixture(e, pi, 1, nullptr, freqs, eigenvecs, inv_eigenvecs, eigenvals, rates, rate_weights, prop_invar);
}

/* K1 launch for n edges of partition pi (inputs through the staging ring) */
static int launch_pmatrices(nrx_engine *e, uint32_t pi, uint32_t n, const uint32_t *edge_idx, const double *brlen) {
  Part &p = e->parts[pi];
  uint32_t *d_idx; double *d_len;
  if (!upload(e, edge_idx, n, &d_idx) || !upload(e, brlen, n, &d_len)) return 0;
  PartView v = make_view(p, pi);
  cudaEvent_t ev0, ev1;
  prof_begin(e, &ev0, &ev1);
  k_pmatrix<<<n, 128, p.d.rate_cats * p.d.states * sizeof(double), e->stream>>>(v, p.pmat, d_idx, d_len);
  e->launches++;
  CK(cudaGetLastError());
  if (!refresh_tiplut(e, pi, edge_idx, n)) return 0;  // K1b: tip tables of the updated edges for the DMMA kernel
  prof_end(e, ev0, ev1, 1, n, (unsigned long long)n * p.pmat_entries * 8, NRX_PROF_K1);
  return 1;
}
/* issue the deferred P-matrix updates (every entry point that reads P-matrices calls this first) */
static int flush_pmatrices(nrx_engine *e) {
  // several partitions of one shape class with pending edges (unlinked branch lengths): ONE launch for the class
  for (const ShapeClass &c : e->classes) {
    if (aa_dmma_class(e, c) || std::getenv("NRX_NO_K1_MULTI")) continue;   // 20-state classes refresh their tip tables per partition (K1b)
    uint32_t pending_parts = 0;
    for (uint32_t pi : c.parts) pending_parts += e->parts[pi].npend ? 1 : 0;
    if (pending_parts < 2) continue;
    std::vector<uint32_t> idx, off(1, 0);
    std::vector<double> len;
    uint32_t most = 0, smem_doubles = 0;
    for (uint32_t pi : c.parts) {
      Part &p = e->parts[pi];
      for (uint32_t m = 0; m < p.d.edges; ++m) if (p.pend[m]) { idx.push_back(m); len.push_back(p.h_len[m]); p.pend[m] = 0; }
      p.npend = 0;
      most = std::max<uint32_t>(most, (uint32_t)idx.size() - off.back());
      off.push_back((uint32_t)idx.size());
      smem_doubles = std::max<uint32_t>(smem_doubles, p.d.rate_cats * p.d.states);
    }
    if (!refresh_views(e)) return 0;
    uint32_t *d_idx, *d_off; double *d_len;
    if (!upload(e, idx.data(), idx.size(), &d_idx) || !upload(e, len.data(), len.size(), &d_len) || !upload(e, off.data(), off.size(), &d_off)) return 0;
    cudaEvent_t ev0, ev1;
    prof_begin(e, &ev0, &ev1);
    k_pmatrix_multi<<<dim3(most, (uint32_t)c.parts.size()), 128, smem_doubles * sizeof(double), e->stream>>>(c.d_views, d_idx, d_len, d_off);
    e->launches++;
    CK(cudaGetLastError());
    prof_end(e, ev0, ev1, 1, idx.size(), (unsigned long long)idx.size() * e->parts[c.parts[0]].pmat_entries * 8, NRX_PROF_K1);
  }
  for (uint32_t pi = 0; pi < e->parts.size(); ++pi) {
    Part &p = e->parts[pi];
    if (!p.npend) continue;
    std::vector<uint32_t> idx;
    std::vector<double> len;
    for (uint32_t m = 0; m < p.d.edges; ++m) if (p.pend[m]) { idx.push_back(m); len.push_back(p.h_len[m]); p.pend[m] = 0; }
    p.npend = 0;
    if (!launch_pmatrices(e, pi, (uint32_t)idx.size(), idx.data(), len.data())) return 0;
  }
  return 1;
}

int nrx_update_pmatrices(nrx_engine *e, uint32_t pi, uint32_t n, const uint32_t *edge_idx, const double *brlen) {
  if (!check_part(e, pi)) return 0;
  if (n == 0) return 1;
  CK(cudaSetDevice(e->device));
  Part &p = e->parts[pi];
  if (!p.model_set) { g_err = "nrx_update_pmatrices: model not set"; return 0; }
  for (uint32_t i = 0; i < n; ++i) {
    if (edge_idx[i] >= p.d.edges) { g_err = "nrx_update_pmatrices: edge index out of range"; return 0; }
    if (!(brlen[i] >= 0.0)) { g_err = "nrx_update_pmatrices: negative branch length"; return 0; }
  }
  for (uint32_t i = 0; i < n; ++i) p.h_len[edge_idx[i]] = brlen[i];
  if (e->defer_pmat && p.pinv == 0.0 && p.nmodels == 1) {
    // deferred: a full evaluation that tile-walks computes the P-matrices inside its one launch (nrx_plan_evaluate_async);
    // anything else that reads P-matrices flushes the pending edges through K1 first
    for (uint32_t i = 0; i < n; ++i) if (!p.pend[edge_idx[i]]) { p.pend[edge_idx[i]] = 1; p.npend++; }
    return 1;
  }
  if (p.npend) {   // keep the order of updates: an edge that is pending and updated again here must end with the newer length
    for (uint32_t i = 0; i < n; ++i) if (p.pend[edge_idx[i]]) { p.pend[edge_idx[i]] = 0; p.npend--; }
    if (!flush_pmatrices(e)) return 0;
  }
  return launch_pmatrices(e, pi, n, edge_idx, brlen);
}

int nrx_get_pmatrix(nrx_engine *e, uint32_t pi, uint32_t edge, double *out) {
  if (!check_part(e, pi)) return 0;
  CK(cudaSetDevice(e->device));
  if (!flush_pmatrices(e)) return 0;
  Part &p = e->parts[pi];
  if (edge >= p.d.edges) { g_err = "edge index out of range"; return 0; }
  CK(cudaStreamSynchronize(e->stream));
  CK(cudaMemcpy(out, p.pmat + (size_t)edge * p.pmat_entries, p.pmat_entries * sizeof(double), cudaMemcpyDeviceToHost));
  return 1;
}

int nrx_set_pmatrix(nrx_engine *e, uint32_t pi, uint32_t edge, const double *in) {
  if (!check_part(e, pi)) return 0;
  CK(cudaSetDevice(e->device));
  if (!flush_pmatrices(e)) return 0;
  Part &p = e->parts[pi];
  if (edge >= p.d.edges) { g_err = "edge index out of range"; return 0; }
  CK(cudaStreamSynchronize(e->stream));
  CK(cudaMemcpy(p.pmat + (size_t)edge * p.pmat_entries, in, p.pmat_entries * sizeof(double), cudaMemcpyHostToDevice));
  if (p.pmat_pad)   // keep the tile-walk copy in step (test hook: one 2-D copy, 4 category blocks of 16 doubles at pitch PCAT)
    CK(cudaMemcpy2D(p.pmat_pad + (size_t)edge * WALK_PE, PCAT * sizeof(double), in, 16 * sizeof(double), 16 * sizeof(double), 4, cudaMemcpyHostToDevice));
  return refresh_tiplut(e, pi, &edge, 1);
}

uint32_t nrx_num_slots(nrx_engine *e) { return e ? e->nslots : 0; }

int nrx_reserve_slots(nrx_engine *e, uint32_t nslots) {
  if (!e) { g_err = "null engine"; return 0; }
  if (nslots <= e->nslots) return 1;
  CK(cudaSetDevice(e->device));
  CK(cudaStreamSynchronize(e->stream));
  for (Part &p : e->parts) {
    const size_t clv_bytes = ((size_t)p.pat_pad * p.d.rate_cats * p.sp * sizeof(double) + 255) & ~(size_t)255;
    const size_t sc_bytes = ((size_t)p.pat_pad * sizeof(uint32_t) + 255) & ~(size_t)255;
    for (uint32_t s = (uint32_t)p.slot_mem.size(); s < nslots; ++s) {
      void *m = nullptr;
      cudaError_t err = cudaMalloc(&m, std::max<size_t>(256, clv_bytes + sc_bytes));
      if (err != cudaSuccess) {
        g_err = std::string("nrx_reserve_slots: out of device memory at slot ") + std::to_string(s) + " (" + cudaGetErrorString(err) + ")";
        cudaGetLastError();
        return 0;
      }
      cudaMemsetAsync(m, 0, std::max<size_t>(256, clv_bytes + sc_bytes), e->stream);
      p.slot_mem.push_back(m);
      p.h_clv.push_back((double *)m);
      p.h_scaler.push_back((uint32_t *)((char *)m + clv_bytes));
    }
    if (nslots > p.table_cap) {
      uint32_t cap = std::max<uint32_t>(nslots, p.table_cap * 2 + 64);
      cudaFree(p.d_clv); cudaFree(p.d_scaler);
      CK(cudaMalloc((void **)&p.d_clv, cap * sizeof(double *)));
      CK(cudaMalloc((void **)&p.d_scaler, cap * sizeof(uint32_t *)));
      p.table_cap = cap;
      e->views_dirty = true;
    }
    CK(cudaMemcpy(p.d_clv, p.h_clv.data(), p.h_clv.size() * sizeof(double *), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(p.d_scaler, p.h_scaler.data(), p.h_scaler.size() * sizeof(uint32_t *), cudaMemcpyHostToDevice));
  }
  e->nslots = nslots;
  return 1;
}

static int reserve_sumtables(nrx_engine *e, uint32_t n) {
  if (n <= e->nsumtables) return 1;
  CK(cudaStreamSynchronize(e->stream));
  for (Part &p : e->parts) {
    for (uint32_t s = (uint32_t)p.h_sumtable.size(); s < n; ++s) {
      double *m = nullptr;
      cudaError_t err = cudaMalloc((void **)&m, std::max<size_t>(256, p.clv_entries * sizeof(double)));
      if (err != cudaSuccess) { g_err = std::string("sumtable pool: ") + cudaGetErrorString(err); cudaGetLastError(); return 0; }
      p.h_sumtable.push_back(m);
    }
    if (n > p.st_cap) {
      uint32_t cap = std::max<uint32_t>(n, p.st_cap * 2 + 16);
      cudaFree(p.d_sumtable);
      CK(cudaMalloc((void **)&p.d_sumtable, cap * sizeof(double *)));
      p.st_cap = cap;
      e->views_dirty = true;
    }
    CK(cudaMemcpy(p.d_sumtable, p.h_sumtable.data(), p.h_sumtable.size() * sizeof(double *), cudaMemcpyHostToDevice));
  }
  e->nsumtables = n;
  return 1;
}

int nrx_copy_slot(nrx_engine *e, uint32_t dst, uint32_t src) {
  if (!e) { g_err = "null engine"; return 0; }
  if (dst >= e->nslots || src >= e->nslots) { g_err = "nrx_copy_slot: slot out of range"; return 0; }
  CK(cudaSetDevice(e->device));
  cudaEvent_t ev0, ev1;
  prof_begin(e, &ev0, &ev1);
  unsigned long long bytes = 0;
  for (Part &p : e->parts) {
    CK(cudaMemcpyAsync(p.h_clv[dst], p.h_clv[src], p.clv_entries * sizeof(double), cudaMemcpyDeviceToDevice, e->stream));
    CK(cudaMemcpyAsync(p.h_scaler[dst], p.h_scaler[src], (size_t)p.d.patterns * sizeof(uint32_t), cudaMemcpyDeviceToDevice, e->stream));
    bytes += 2 * (p.clv_entries * sizeof(double) + (size_t)p.d.patterns * sizeof(uint32_t));
  }
  prof_end(e, ev0, ev1, 2 * e->parts.size(), 1, bytes, NRX_PROF_COPY);
  return 1;
}

int nrx_copy_slots(nrx_engine *e, const uint32_t *dst, const uint32_t *src, uint32_t n) {
  if (!e) { g_err = "null engine"; return 0; }
  if (n == 0) return 1;
  std::vector<uint2> pairs(n);
  for (uint32_t i = 0; i < n; ++i) {
    if (dst[i] >= e->nslots || src[i] >= e->nslots) { g_err = "nrx_copy_slots: slot out of range"; return 0; }
    pairs[i] = make_uint2(dst[i], src[i]);
  }
  CK(cudaSetDevice(e->device));
  if (!refresh_views(e)) return 0;
  uint2 *d_pairs;
  if (!upload(e, pairs.data(), pairs.size(), &d_pairs)) return 0;
  cudaEvent_t ev0, ev1;
  prof_begin(e, &ev0, &ev1);
  unsigned long long bytes = 0;
  for (const Part &p : e->parts) bytes += 2ull * n * (p.clv_entries * sizeof(double) + (size_t)p.d.patterns * sizeof(uint32_t));
  unsigned long long launches = 0;
  for (const ShapeClass &c : e->classes) {
    if (c.max_patterns == 0) continue;
    const uint32_t z = (uint32_t)c.parts.size();
    const uint64_t units = (uint64_t)c.max_patterns * c.cats * ((c.states + 3) & ~3u) / 4;
    dim3 grid(tiles_for(units, BLOCK * 4, n * z), n, z);
    k_copy_slots<<<grid, BLOCK, 0, e->stream>>>(c.d_views, d_pairs);
    e->launches++; launches++;
    CK(cudaGetLastError());
  }
  prof_end(e, ev0, ev1, launches, n, bytes, NRX_PROF_COPY);
  return 1;
}

/* validation + byte accounting of ONE batch (= one launch per shape class) of CLV updates.
 *   *bytes  : ALGORITHMIC bytes, SURVEY §8d's per-op figures (every op charged both of its operands);
 *   *cbytes : COMPULSORY bytes of the launch — every distinct child CLV (+ scaler) / tip row read once, every parent CLV
 *             (+ scaler) written once.  The ops of a node share children (12 x 8 compatible pairs share 20 children), the
 *             re-reads hit the 126 MB L2, so this — not the algorithmic figure — is what must cross the HBM pins and is the
 *             numerator of the roofline fraction (the algorithmic numerator gave 1.38 "of peak"). */
static int check_ops(nrx_engine *e, const nrx_op *ops, uint32_t nops, unsigned long long *updates, unsigned long long *bytes,
                     unsigned long long *cbytes = nullptr) {
  std::vector<uint64_t> operands;   // (kind << 32 | idx) of every CLV / tip operand of the batch
  for (uint32_t i = 0; i < nops; ++i) {
    const nrx_op &o = ops[i];
    if (o.parent_slot >= e->nslots) { g_err = "nrx_update_clvs: parent slot out of range"; return 0; }
    const uint32_t kinds[2] = {o.left_kind, o.right_kind}, idx[2] = {o.left_idx, o.right_idx}, edges[2] = {o.left_edge, o.right_edge};
    for (int s = 0; s < 2; ++s) {
      if (kinds[s] == NRX_CLV && idx[s] >= e->nslots) { g_err = "nrx_update_clvs: child slot out of range"; return 0; }
      if (kinds[s] > NRX_NONE) { g_err = "nrx_update_clvs: bad operand kind"; return 0; }
      for (const Part &p : e->parts) {
        if (kinds[s] == NRX_TIP && idx[s] >= p.d.tips) { g_err = "nrx_update_clvs: tip index out of range"; return 0; }
        if (kinds[s] != NRX_NONE && edges[s] >= p.d.edges) { g_err = "nrx_update_clvs: edge index out of range"; return 0; }
      }
      if (kinds[s] != NRX_NONE) operands.push_back(((uint64_t)kinds[s] << 32) | idx[s]);
    }
    if (o.left_kind == NRX_NONE && o.right_kind == NRX_NONE) { g_err = "nrx_update_clvs: both operands absent"; return 0; }
    for (const Part &p : e->parts) {
      const unsigned long long Cb = (unsigned long long)p.d.rate_cats * p.sp * 8;
      unsigned long long b = Cb + 4;
      for (int s = 0; s < 2; ++s) b += (kinds[s] == NRX_CLV) ? Cb + 4 : (kinds[s] == NRX_TIP ? 1 : 0);
      if (o.left_kind == NRX_TIP && o.right_kind == NRX_TIP) b = Cb + 2 + 4;
      *bytes += b * p.d.patterns;
      *updates += p.d.patterns;
    }
  }
  if (cbytes) {
    std::sort(operands.begin(), operands.end());
    operands.erase(std::unique(operands.begin(), operands.end()), operands.end());
    unsigned long long clvs = 0, tips = 0;
    for (uint64_t o : operands) ((o >> 32) == NRX_CLV ? clvs : tips)++;
    for (const Part &p : e->parts) {
      const unsigned long long Cb = (unsigned long long)p.d.rate_cats * p.sp * 8;
      *cbytes += (clvs * (Cb + 4) + tips + (unsigned long long)nops * (Cb + 4)) * p.d.patterns;
    }
  }
  return 1;
}
/* compulsory bytes of a K4 / K5 launch over `n` operand pairs: every distinct CLV operand (+ its scaler when the kernel reads
 * it) / tip row once, plus `out_clvs` CLV-sized outputs per pair and `extra` bytes per (pair, pattern) */
static unsigned long long pair_cbytes(const nrx_engine *e, const nrx_pair *pairs, uint32_t n, bool scalers, unsigned out_clvs, unsigned extra) {
  std::vector<uint64_t> operands;
  for (uint32_t i = 0; i < n; ++i) {
    operands.push_back(((uint64_t)pairs[i].a_kind << 32) | pairs[i].a_idx);
    operands.push_back(((uint64_t)pairs[i].b_kind << 32) | pairs[i].b_idx);
  }
  std::sort(operands.begin(), operands.end());
  operands.erase(std::unique(operands.begin(), operands.end()), operands.end());
  unsigned long long clvs = 0, tips = 0, b = 0;
  for (uint64_t o : operands) ((o >> 32) == NRX_CLV ? clvs : tips)++;
  for (const Part &p : e->parts) {
    const unsigned long long Cb = (unsigned long long)p.d.rate_cats * p.sp * 8;
    b += (clvs * (Cb + (scalers ? 4 : 0)) + tips + (unsigned long long)n * (out_clvs * Cb + extra)) * p.d.patterns;
  }
  return b;
}

/* K2 launches (one per partition shape class) for `nops` device-resident ops; `with_tips`: some op has a tip operand */
static bool has_aa_dmma(const nrx_engine *e);
static bool aa_dmma_class(const nrx_engine *e, const ShapeClass &c);
static int launch_clv_batch(nrx_engine *e, const nrx_op *d_ops, uint32_t nops, bool with_tips, bool fused = false, uint32_t ntt = 0) {
  for (const ShapeClass &c : e->classes) {
    if (c.max_patterns == 0) continue;
    const uint32_t z = (uint32_t)c.parts.size();
    if (dna_pipe_cats(c.states, c.cats) && (c.cats == 4 || e->k2_variant == 0)) {
      if (e->k2_variant == 0 || e->k2_variant == 1) {
        // bulk-async pipeline: 2 resident blocks per SM; block b = (op b % nops, tile group b / nops)
        const uint32_t nt = (e->k2_variant == 0 && c.cats == 4) ? e->k2_nt : (e->k2_variant == 0 ? 2u : 1u);   // 256-item sub-tiles per ring stage
        const uint32_t tpx = nt * (BLOCK / c.cats);   // patterns per stage (4 categories: 64 per sub-tile)
        const uint32_t ntiles = (c.max_patterns + tpx - 1) / tpx;
        uint32_t groups = std::max<uint32_t>(1, (e->k2_blocks + nops * z - 1) / (nops * z));
        // >= 4 tiles per block amortise the pipeline fill — unless the whole launch fits the 2 x SMs resident blocks anyway:
        // then one tile per block (small alignments: 38 blocks walking 4 tiles each left 110 SMs idle)
        const uint32_t one_wave = e->throughput_mode ? 0u : (2u * (uint32_t)e->sm_count) / std::max<uint32_t>(1, nops * z);
        groups = std::min(groups, std::max<uint32_t>(std::max<uint32_t>(1, ntiles / 4), std::min(ntiles, one_wave)));
        dim3 grid(nops * groups, 1, z);
        double *fused_ptr = fused ? e->d_fused : nullptr;
        if (e->k2_variant == 0) {
          // inside a plan capture the launch is programmatically serialised behind the previous K2 launch (PDL)
          cudaLaunchConfig_t cfg{};
          cfg.gridDim = grid; cfg.blockDim = dim3(BLOCK); cfg.stream = e->stream;
          cudaLaunchAttribute attr[1];
          attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
          attr[0].val.programmaticStreamSerializationAllowed = 1;
          // only for launches that fit one wave of resident blocks (latency-bound); for big launches the early ring
          // fill of the stream-ordered form is worth more (config 2: 0.884 vs 0.915 ms per evaluation with PDL everywhere)
          const int pdl = (e->capturing_pdl && e->pdl_prev_is_k2 && (uint64_t)nops * groups * z <= 2ull * (uint64_t)e->sm_count) ? 1 : 0;
          cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
          const size_t stride = (size_t)e->max_patterns;
          const uint32_t np = (uint32_t)e->parts.size();
#define NRX_K2_LAUNCH(NT_, CATS_)                                                                                                    \
  do {                                                                                                                               \
    cfg.dynamicSmemBytes = sizeof(PipeSmem<NT_, CATS_>);                                                                             \
    CK(cudaLaunchKernelEx(&cfg, k_clv_dna4_pipe2<NT_, CATS_>, (const PartView *)c.d_views, d_ops, nops, groups, fused_ptr, stride, np, pdl | ((fused && e->score_only) ? 2 : 0))); \
  } while (0)
          if (c.cats == 4) { if (nt == 2) NRX_K2_LAUNCH(2, 4); else NRX_K2_LAUNCH(1, 4); }
          else if (c.cats == 1) NRX_K2_LAUNCH(2, 1);
          else if (c.cats == 2) NRX_K2_LAUNCH(2, 2);
          else if (c.cats == 8) NRX_K2_LAUNCH(2, 8);
          else NRX_K2_LAUNCH(2, 16);
#undef NRX_K2_LAUNCH
          e->pdl_prev_is_k2 = true;
        } else
          k_clv_dna4_pipe<<<grid, BLOCK, sizeof(ClvPipeSmem), e->stream>>>(c.d_views, d_ops, nops, groups, fused_ptr, (size_t)e->max_patterns, (uint32_t)e->parts.size());
      } else { g_err = "NRX_K2: unknown K2 variant (0 = k_clv_dna4_pipe2, 1 = k_clv_dna4_pipe)"; return 0; }
    } else if (c.states == 20 && c.cats == 4 && !e->aa_generic && class_tip_codes(e, c) <= (uint32_t)AA_LUT_CODES) {
      // protein.  The first `ntt` ops are tip-tip: a table product, written by a plain streaming kernel
      if (ntt) {
        const uint32_t want = std::max<uint32_t>(1, (148u * 8 + ntt * z - 1) / (ntt * z));
        uint32_t chunk = (c.max_patterns + want - 1) / want;
        chunk = std::max<uint32_t>(64, (chunk + 31) & ~31u);
        dim3 grid((c.max_patterns + chunk - 1) / chunk, ntt, z);
        k_clv_aa20_tiptip<<<grid, BLOCK, (size_t)2 * class_tip_codes(e, c) * 80 * sizeof(double), e->stream>>>(c.d_views, d_ops, chunk);
        if (ntt == nops) { e->launches++; CK(cudaGetLastError()); continue; }
        e->launches++;
      }
      // the rest on the FP64 tensor cores (DMMA): 3 resident blocks of 4+1 warps per SM, 8-pattern tiles
      const uint32_t rest = nops - ntt;
      const uint32_t ntiles = (c.max_patterns + AA_TP - 1) / AA_TP;
      // k_aa20_mma: >= ~48 tiles per block amortise the set-up (config 4 at 20 k patterns: 592 blocks 0.348 ms, 1184 blocks 0.374 ms;
      // at 200 k patterns 2.75 vs 2.67 ms the other way round)
      const uint32_t by_size = (uint32_t)std::min<uint64_t>(8ull * e->sm_count, std::max<uint64_t>(4ull * e->sm_count, (uint64_t)rest * z * ntiles / 48));
      const uint32_t target = e->aa_v1 ? e->aa_blocks : (e->aa2_blocks ? e->aa2_blocks : by_size);
      uint32_t groups = std::max<uint32_t>(1, (target + rest * z - 1) / (rest * z));
      groups = std::min(groups, std::max<uint32_t>(1, ntiles / 8));
      dim3 grid(rest * groups, 1, z);
      // tip-tip ops went to the table kernel above unless that path is disabled: then two tables are needed
      const int luts = !with_tips ? 0 : (has_aa_dmma(e) ? 1 : 2);
      const size_t lut_bytes = (size_t)luts * class_tip_codes(e, c) * 80 * sizeof(double);
      if (e->aa_v1) k_aa20_dmma<AA_CLV><<<grid, AA_THREADS, sizeof(AaSmem) + lut_bytes, e->stream>>>(c.d_views, d_ops + ntt, rest, groups, luts, nullptr, 0, 0.0);
      else {
        // inside a plan capture the launch is programmatically serialised behind the previous K2 launch (PDL): its blocks take the
        // SM slots the draining predecessor frees and run their prologue there; the loader warps wait for the predecessor's CLVs
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = grid; cfg.blockDim = dim3(AA2_THREADS); cfg.stream = e->stream; cfg.dynamicSmemBytes = sizeof(AaSmem2) + lut_bytes;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        const int pdl = (e->capturing_pdl && e->pdl_prev_is_k2) ? 1 : 0;
        cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
        const nrx_op *ops_rest = d_ops + ntt;
        double *ps = fused ? e->d_fused : nullptr;   // fused K3: the ops marked as root displayed trees also emit their per-site terms
        const uint32_t np = (uint32_t)e->parts.size();
        const size_t ps_stride = (size_t)e->max_patterns;
        if (e->aa_pipe) CK(cudaLaunchKernelEx(&cfg, k_aa20_mma<AA_CLV, true>, (const PartView *)c.d_views, ops_rest, rest, groups, luts, (double *)nullptr, np, 0.0, (double *)nullptr, (uint32_t *)nullptr, pdl, ps, ps_stride));
        else CK(cudaLaunchKernelEx(&cfg, k_aa20_mma<AA_CLV, false>, (const PartView *)c.d_views, ops_rest, rest, groups, luts, (double *)nullptr, np, 0.0, (double *)nullptr, (uint32_t *)nullptr, pdl, ps, ps_stride));
        e->pdl_prev_is_k2 = true;
      }
    } else {
      dim3 grid(tiles_for(c.max_patterns, BLOCK, nops * z), nops, z);
      k_clv_generic<<<grid, BLOCK, 0, e->stream>>>(c.d_views, d_ops, nullptr);
    }
    e->launches++;
    CK(cudaGetLastError());
  }
  return 1;
}

/* engines with a 20-state tensor-core class order each batch tip-tip first (ops of a batch are independent) */
static bool has_aa_dmma(const nrx_engine *e) {
  for (const ShapeClass &c : e->classes)
    if (c.states == 20 && c.cats == 4 && !e->aa_generic && class_tip_codes(e, c) <= (uint32_t)AA_LUT_CODES && !std::getenv("NRX_AA_NO_TT")) return true;
  return false;
}
static uint32_t order_tiptip_first(const nrx_engine *e, std::vector<nrx_op> &ops) {
  if (!has_aa_dmma(e)) return 0;
  auto mid = std::stable_partition(ops.begin(), ops.end(), [](const nrx_op &o) { return o.left_kind == NRX_TIP && o.right_kind == NRX_TIP; });
  return (uint32_t)(mid - ops.begin());
}

static bool any_tip(const nrx_op *ops, uint32_t nops) {
  for (uint32_t i = 0; i < nops; ++i) if (ops[i].left_kind == NRX_TIP || ops[i].right_kind == NRX_TIP) return true;
  return false;
}

int nrx_update_clvs(nrx_engine *e, const nrx_op *ops, uint32_t nops) {
  if (!e) { g_err = "null engine"; return 0; }
  if (nops == 0) return 1;
  CK(cudaSetDevice(e->device));
  if (!flush_pmatrices(e)) return 0;
  unsigned long long updates = 0, bytes = 0, cbytes = 0;
  if (!check_ops(e, ops, nops, &updates, &bytes, &cbytes)) return 0;
  if (!refresh_views(e)) return 0;
  std::vector<nrx_op> ordered(ops, ops + nops);
  const uint32_t ntt = order_tiptip_first(e, ordered);
  nrx_op *d_ops;
  if (!upload(e, ordered.data(), nops, &d_ops)) return 0;
  cudaEvent_t ev0, ev1;
  prof_begin(e, &ev0, &ev1);
  if (!launch_clv_batch(e, d_ops, nops, any_tip(ops, nops), false, ntt)) return 0;
  prof_end(e, ev0, ev1, e->classes.size(), updates, bytes, NRX_PROF_K2, cbytes);
  return 1;
}

int nrx_update_pseudo_clvs(nrx_engine *e, const nrx_pseudo_op *ops, uint32_t nops) {
  if (!e) { g_err = "null engine"; return 0; }
  if (nops == 0) return 1;
  CK(cudaSetDevice(e->device));
  if (!flush_pmatrices(e)) return 0;
  unsigned long long updates = 0, bytes = 0;
  for (uint32_t i = 0; i < nops; ++i) {   // same operand rules as nrx_update_clvs, except that both operands may be absent
    const nrx_pseudo_op &o = ops[i];
    nrx_op chk{};
    chk.parent_slot = o.parent_slot;
    chk.left_kind = o.left_kind; chk.left_idx = o.left_idx; chk.left_edge = o.left_edge;
    chk.right_kind = o.right_kind; chk.right_idx = o.right_idx; chk.right_edge = o.right_edge;
    if (chk.left_kind == NRX_NONE && chk.right_kind == NRX_NONE) { chk.left_kind = NRX_TIP; chk.left_idx = 0; chk.left_edge = 0; }
    if (!check_ops(e, &chk, 1, &updates, &bytes)) return 0;
    for (int k = 0; k < 4; ++k) if (!(o.w[k] >= 0.0 && o.w[k] <= 1.0)) { g_err = "nrx_update_pseudo_clvs: weight outside [0, 1]"; return 0; }
  }
  if (!refresh_views(e)) return 0;
  nrx_pseudo_op *d_ops;
  if (!upload(e, ops, nops, &d_ops)) return 0;
  cudaEvent_t ev0, ev1;
  prof_begin(e, &ev0, &ev1);
  for (const ShapeClass &c : e->classes) {
    if (c.max_patterns == 0) continue;
    const uint32_t z = (uint32_t)c.parts.size();
    if (c.states == 4 && c.cats == 4) {
      dim3 grid(tiles_for((uint64_t)c.max_patterns * 4, BLOCK, nops * z), nops, z);
      k_clv_pseudo_dna4<<<grid, BLOCK, 0, e->stream>>>(c.d_views, d_ops);
    } else {
      dim3 grid(tiles_for(c.max_patterns, BLOCK, nops * z), nops, z);
      k_clv_pseudo_generic<<<grid, BLOCK, 0, e->stream>>>(c.d_views, d_ops);
    }
    e->launches++;
    CK(cudaGetLastError());
  }
  prof_end(e, ev0, ev1, e->classes.size(), updates, bytes);
  return 1;
}

/* Node-centric grouping of ONE batch (k_clv_node_dna4).  Candidates are the inner x inner ops; ops with the same (left edge,
 * right edge) belong to one network node; their (left slot, right slot) pairs form a bipartite graph whose connected components
 * are (in practice complete) bipartite blocks — displayed trees that agree on the shared reticulations.  A component with
 * <= NODE_MAXC children becomes one group; a larger one is cut along its larger side (the smaller side whole when it has
 * <= NODE_MAXC / 2 children, else both sides in chunks of NODE_MAXC / 2).  A group is kept if it saves reads: 2 ops > children.
 * Returns the ops that stay on the per-op kernel in `rest`; appends groups / group ops. */
static void group_batch_for_node_kernel(const std::vector<nrx_op> &batch, std::vector<nrx_op> &rest, std::vector<nrx_node_group> &groups,
                                        std::vector<nrx_node_op> &gops, std::vector<uint32_t> &batch_groups, const size_t MAXC) {
  std::map<std::pair<uint32_t, uint32_t>, std::vector<uint32_t>> by_node;
  for (uint32_t i = 0; i < batch.size(); ++i) {
    const nrx_op &o = batch[i];
    if (o.left_kind == NRX_CLV && o.right_kind == NRX_CLV) by_node[{o.left_edge, o.right_edge}].push_back(i);
    else rest.push_back(o);
  }
  for (auto &kv : by_node) {
    const std::vector<uint32_t> &idx = kv.second;
    // connected components over (left slot, right slot)
    std::map<uint32_t, uint32_t> lcomp, rcomp;
    std::vector<uint32_t> comp(idx.size());
    uint32_t ncomp = 0;
    std::vector<uint32_t> parent;
    auto find = [&](uint32_t x) { while (parent[x] != x) { parent[x] = parent[parent[x]]; x = parent[x]; } return x; };
    for (size_t k = 0; k < idx.size(); ++k) {
      const nrx_op &o = batch[idx[k]];
      auto li = lcomp.find(o.left_idx), ri = rcomp.find(o.right_idx);
      uint32_t c;
      if (li == lcomp.end() && ri == rcomp.end()) { c = ncomp++; parent.push_back(c); }
      else if (li != lcomp.end() && ri != rcomp.end()) { const uint32_t a = find(li->second), b = find(ri->second); parent[b] = a; c = a; }
      else c = find(li != lcomp.end() ? li->second : ri->second);
      lcomp[o.left_idx] = c; rcomp[o.right_idx] = c; comp[k] = c;
    }
    std::map<uint32_t, std::vector<uint32_t>> comps;
    for (size_t k = 0; k < idx.size(); ++k) comps[find(comp[k])].push_back(idx[k]);
    for (auto &cv : comps) {
      std::vector<uint32_t> L, R;
      for (uint32_t i : cv.second) {
        if (std::find(L.begin(), L.end(), batch[i].left_idx) == L.end()) L.push_back(batch[i].left_idx);
        if (std::find(R.begin(), R.end(), batch[i].right_idx) == R.end()) R.push_back(batch[i].right_idx);
      }
      // chunk sizes per side
      size_t cl = L.size(), cr = R.size();
      if (L.size() + R.size() > MAXC) {
        if (L.size() <= MAXC / 2) cr = MAXC - L.size();
        else if (R.size() <= MAXC / 2) cl = MAXC - R.size();
        else cl = cr = MAXC / 2;
      }
      for (size_t l0 = 0; l0 < L.size(); l0 += cl)
        for (size_t r0 = 0; r0 < R.size(); r0 += cr) {
          const size_t l1 = std::min(L.size(), l0 + cl), r1 = std::min(R.size(), r0 + cr);
          std::vector<nrx_node_op> ops;
          std::vector<uint32_t> members;
          for (uint32_t i : cv.second) {
            const auto li = std::find(L.begin() + l0, L.begin() + l1, batch[i].left_idx), ri = std::find(R.begin() + r0, R.begin() + r1, batch[i].right_idx);
            if (li == L.begin() + l1 || ri == R.begin() + r1) continue;
            nrx_node_op no{};
            no.li = (uint16_t)(li - (L.begin() + l0)); no.rj = (uint16_t)(ri - (R.begin() + r0));
            no.parent_slot = batch[i].parent_slot; no.lnl_item = batch[i].lnl_item;
            ops.push_back(no);
            members.push_back(i);
          }
          if (ops.empty()) continue;
          // children actually referenced by this chunk's ops (a sparse component may leave some unused)
          const size_t children = (l1 - l0) + (r1 - r0);
          if (2 * ops.size() <= children + 1 || ops.size() > (size_t)NODE_MAXOPS) { for (uint32_t i : members) rest.push_back(batch[i]); continue; }
          nrx_node_group g{};
          g.nl = (uint32_t)(l1 - l0); g.nr = (uint32_t)(r1 - r0); g.nops = (uint32_t)ops.size();
          g.left_edge = kv.first.first; g.right_edge = kv.first.second; g.op_first = (uint32_t)gops.size();
          for (size_t a = l0; a < l1; ++a) g.child_slot[a - l0] = L[a];
          for (size_t a = r0; a < r1; ++a) g.child_slot[g.nl + (a - r0)] = R[a];
          batch_groups.push_back((uint32_t)groups.size());
          groups.push_back(g);
          gops.insert(gops.end(), ops.begin(), ops.end());
        }
    }
  }
}

/* Tile-walk form of a plan (k_walk_dna4): (1) order the ops depth-first from the CLVs nobody consumes (the root displayed trees),
 * children before parents, so that a CLV is consumed soon after it is produced; (2) run a linear-scan allocation of
 * shared-memory buffers over that order — a CLV needs a buffer from its op until its last consumer; (3) accept the form if
 * P-matrices + buffers + program + tip codes fit one block's shared memory.  Leaves pl.d_walk == nullptr when the plan has no
 * such form (other partition shapes, +I, a CLV produced twice or not at all, too many live CLVs): the level-by-level graph runs. */
static int build_walk_program(nrx_engine *e, EnginePlan &pl, const nrx_op *ops, size_t total) {
  if (e->classes.size() != 1 || e->classes[0].states != 4 || e->classes[0].cats != 4 || !nrx_supports_fused_lnl(e) || total == 0 || total > 60000) return 1;
  std::map<uint32_t, uint32_t> producer;   // slot -> op
  for (size_t i = 0; i < total; ++i) if (!producer.emplace(ops[i].parent_slot, (uint32_t)i).second) return 1;
  std::vector<uint32_t> uses(total, 0);
  for (size_t i = 0; i < total; ++i) {
    const uint32_t kinds[2] = {ops[i].left_kind, ops[i].right_kind}, idx[2] = {ops[i].left_idx, ops[i].right_idx};
    for (int c = 0; c < 2; ++c)
      if (kinds[c] == NRX_CLV) {
        auto it = producer.find(idx[c]);
        if (it == producer.end() || it->second >= i) return 1;   // child not produced by this traversal (or not before its consumer)
        uses[it->second]++;
      }
  }
  // depth-first emission (explicit stack), roots in plan order
  std::vector<uint32_t> order;
  std::vector<char> done(total, 0);
  for (size_t r = 0; r < total; ++r) {
    if (uses[r] != 0 || done[r]) continue;
    std::vector<std::pair<uint32_t, int>> stack{{(uint32_t)r, 0}};
    while (!stack.empty()) {
      auto &[op, stage] = stack.back();
      const uint32_t kinds[2] = {ops[op].left_kind, ops[op].right_kind}, idx[2] = {ops[op].left_idx, ops[op].right_idx};
      if (stage < 2) {
        const int c = stage++;
        if (kinds[c] == NRX_CLV) { const uint32_t q = producer[idx[c]]; if (!done[q]) { done[q] = 1; stack.push_back({q, 0}); } }
      } else {
        order.push_back(op);
        stack.pop_back();
      }
    }
    done[r] = 1;
  }
  if (order.size() != total) return 1;
  // linear-scan buffer allocation
  std::vector<uint16_t> buf(total, (uint16_t)WALK_NOBUF);
  std::vector<uint32_t> left(uses);
  std::vector<uint16_t> free_list;
  uint32_t nbuf = 0;
  std::vector<nrx_walk_op> prog(total);
  std::vector<char> tip_used;
  for (size_t k = 0; k < total; ++k) {
    const nrx_op &o = ops[order[k]];
    nrx_walk_op w{};
    w.parent_slot = o.parent_slot;
    w.kinds = (uint16_t)(o.left_kind | (o.right_kind << 2));
    w.left_idx = o.left_idx; w.right_idx = o.right_idx; w.left_edge = o.left_edge; w.right_edge = o.right_edge;
    w.lnl_item = o.lnl_item;
    w.lbuf = w.rbuf = (uint16_t)WALK_NOBUF;
    if (o.left_kind == NRX_CLV) w.lbuf = buf[producer[o.left_idx]];
    if (o.right_kind == NRX_CLV) w.rbuf = buf[producer[o.right_idx]];
    if (uses[order[k]] > 0) {   // the parent gets its buffer BEFORE the children's are released: no aliasing within an op
      if (free_list.empty()) { if (nbuf >= 4096) return 1; free_list.push_back((uint16_t)nbuf++); }
      std::sort(free_list.begin(), free_list.end(), std::greater<uint16_t>());
      buf[order[k]] = free_list.back(); free_list.pop_back();
    }
    w.pbuf = buf[order[k]];
    const uint32_t kinds[2] = {o.left_kind, o.right_kind}, idx[2] = {o.left_idx, o.right_idx};
    for (int c = 0; c < 2; ++c)
      if (kinds[c] == NRX_CLV) { const uint32_t q = producer[idx[c]]; if (--left[q] == 0) free_list.push_back(buf[q]); }
    prog[k] = w;
  }
  const Part &p0 = e->parts[e->classes[0].parts[0]];
  for (const Part &p : e->parts) if (p.d.edges != p0.d.edges || p.d.tips != p0.d.tips) return 1;
  const size_t smem = (size_t)p0.d.edges * WALK_PE * 8 + (size_t)nbuf * (WALK_TP * 128 + WALK_TP * 4) + (size_t)pl.lnl_items * (WALK_THREADS / 32) * 8 +
                      total * (sizeof(nrx_walk_op) + 16) + (size_t)p0.d.tips * WALK_TP;
  if (smem > 226 * 1024) return 1;
  CK(cudaMalloc((void **)&pl.d_walk, total * sizeof(nrx_walk_op)));
  CK(cudaMemcpy(pl.d_walk, prog.data(), total * sizeof(nrx_walk_op), cudaMemcpyHostToDevice));
  pl.walk_nops = (uint32_t)total; pl.walk_nbuf = nbuf; pl.walk_smem = (smem + 127) & ~(size_t)127;
  for (const Part &p : e->parts) pl.walk_cbytes += (unsigned long long)p.d.patterns * (total * 132ull + p.d.tips);
  return 1;
}

/* ---- evaluation plans: the K2 launches of a whole traversal, ops resident on the device, replayed as ONE CUDA graph ---- */
int nrx_plan_create(nrx_engine *e, const nrx_op *ops, const uint32_t *batch_sizes, uint32_t nbatches, uint32_t *plan_id) {
  if (!e) { g_err = "null engine"; return 0; }
  CK(cudaSetDevice(e->device));
  EnginePlan pl;
  size_t total = 0;
  std::vector<nrx_op> ordered;
  // node-centric K2 for the ops of a node that share children: single-class 4-state x 4-category engines
  const bool node_ok = e->node_mode != 0 && e->classes.size() == 1 && e->classes[0].states == 4 && e->classes[0].cats == 4 && e->k2_variant == 0;
  std::vector<nrx_node_group> ngroups;
  std::vector<nrx_node_op> ngops;
  std::vector<nrx_node_block> nblocks;
  std::vector<nrx_op> all_ops(ops, ops + [&] { size_t t = 0; for (uint32_t b = 0; b < nbatches; ++b) t += batch_sizes[b]; return t; }());   // original order (tile walk)
  size_t in_off = 0;
  for (uint32_t b = 0; b < nbatches; ++b) {
    if (batch_sizes[b] == 0) { g_err = "nrx_plan_create: empty batch"; return 0; }
    if (!check_ops(e, ops + in_off, batch_sizes[b], &pl.updates, &pl.bytes, &pl.cbytes)) return 0;
    std::vector<nrx_op> batch(ops + in_off, ops + in_off + batch_sizes[b]);
    in_off += batch_sizes[b];
    std::vector<uint32_t> bgroups;
    if (node_ok) {
      std::vector<nrx_op> rest;
      group_batch_for_node_kernel(batch, rest, ngroups, ngops, bgroups, e->node_maxc);
      batch.swap(rest);
    }
    // blocks of the node launch: ~8 waves of 3 resident blocks per SM in total, shared out by the groups' op counts, >= 4 tiles per block
    pl.node_block_off.push_back((uint32_t)nblocks.size());
    if (!bgroups.empty()) {
      const uint32_t ntiles = (e->max_patterns + NODE_TP - 1) / NODE_TP;
      uint64_t wsum = 0;
      for (uint32_t g : bgroups) wsum += ngroups[g].nops + ngroups[g].nl + ngroups[g].nr;
      const uint32_t target = e->node_blocks ? e->node_blocks : 24u * (uint32_t)e->sm_count;
      for (uint32_t g : bgroups) {
        const uint64_t w = ngroups[g].nops + ngroups[g].nl + ngroups[g].nr;
        uint32_t nb = (uint32_t)std::max<uint64_t>(1, (uint64_t)target * w / std::max<uint64_t>(1, wsum));
        nb = std::min(nb, std::max<uint32_t>(1, ntiles / 4));
        for (uint32_t k = 0; k < nb; ++k) nblocks.push_back(nrx_node_block{g, k, nb, 0});
      }
    }
    pl.node_block_cnt.push_back((uint32_t)nblocks.size() - pl.node_block_off.back());
    { uint32_t m = 0; for (uint32_t g : bgroups) m = std::max(m, ngroups[g].nl + ngroups[g].nr); pl.node_ncmax.push_back(m); }
    pl.offsets.push_back(total);
    pl.sizes.push_back((uint32_t)batch.size());
    pl.tips.push_back(any_tip(batch.data(), (uint32_t)batch.size()));
    pl.ntt.push_back(order_tiptip_first(e, batch));
    ordered.insert(ordered.end(), batch.begin(), batch.end());
    total += batch.size();
  }
  if (!ngroups.empty()) {
    // interleave the blocks of a batch's groups (group-fastest), so that the big and the small groups of a launch finish together
    for (size_t b = 0; b < pl.node_block_cnt.size(); ++b) {
      auto first = nblocks.begin() + pl.node_block_off[b], last = first + pl.node_block_cnt[b];
      std::stable_sort(first, last, [](const nrx_node_block &x, const nrx_node_block &y) { return x.tile0 < y.tile0; });
    }
    CK(cudaMalloc((void **)&pl.d_ngroups, ngroups.size() * sizeof(nrx_node_group)));
    CK(cudaMalloc((void **)&pl.d_nops, ngops.size() * sizeof(nrx_node_op)));
    CK(cudaMalloc((void **)&pl.d_nblocks, nblocks.size() * sizeof(nrx_node_block)));
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaMemcpy(pl.d_ngroups, ngroups.data(), ngroups.size() * sizeof(nrx_node_group), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(pl.d_nops, ngops.data(), ngops.size() * sizeof(nrx_node_op), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(pl.d_nblocks, nblocks.data(), nblocks.size() * sizeof(nrx_node_block), cudaMemcpyHostToDevice));
  }
  ops = ordered.data();
  if (total) {
    CK(cudaMalloc((void **)&pl.d_ops, total * sizeof(nrx_op)));
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaMemcpy(pl.d_ops, ops, total * sizeof(nrx_op), cudaMemcpyHostToDevice));
  }
  for (const nrx_op &o : all_ops) pl.lnl_items = std::max(pl.lnl_items, o.lnl_item);
  if (pl.lnl_items) {
    if (!nrx_supports_fused_lnl(e)) { cudaFree(pl.d_ops); g_err = "nrx_plan_create: lnl_item marks need every partition on the pipelined 4-state kernel"; return 0; }
    const size_t need = (size_t)pl.lnl_items * e->parts.size() * std::max<uint32_t>(1, e->max_patterns);
    if (need > e->fused_cap) {  // the buffer address is baked into captured graphs: re-capture them
      CK(cudaStreamSynchronize(e->stream));
      cudaFree(e->d_fused);
      e->d_fused = nullptr; e->fused_cap = 0;
      CK(cudaMalloc((void **)&e->d_fused, need * sizeof(double)));
      e->fused_cap = need;
      for (EnginePlan &o : e->plans) for (cudaGraphExec_t &x : o.exec) if (x) { cudaGraphExecDestroy(x); x = nullptr; }
    }
  }
  if (pl.lnl_items && e->walk_mode != 0 && !build_walk_program(e, pl, all_ops.data(), all_ops.size())) return 0;
  pl.alive = true;
  e->plans.push_back(pl);
  *plan_id = (uint32_t)e->plans.size() - 1;
  return 1;
}

/* one batch of a plan: the ops that share children on the node-centric kernel, the rest on the per-op kernels */
static int launch_plan_batch(nrx_engine *e, EnginePlan &pl, size_t b) {
  if (pl.sizes[b] && !launch_clv_batch(e, pl.d_ops + pl.offsets[b], pl.sizes[b], pl.tips[b], pl.lnl_items != 0, pl.ntt[b])) return 0;
  if (pl.node_block_cnt[b]) {
    const ShapeClass &c = e->classes[0];
    k_clv_node_dna4<<<dim3(pl.node_block_cnt[b], 1, (uint32_t)c.parts.size()), NODE_THREADS, node_smem_bytes(pl.node_ncmax[b]), e->stream>>>(
        c.d_views, pl.d_ngroups, pl.d_nops, pl.d_nblocks + pl.node_block_off[b], pl.lnl_items ? e->d_fused : nullptr, (size_t)e->max_patterns, (uint32_t)e->parts.size(), pl.node_ncmax[b]);
    e->launches++;
    e->pdl_prev_is_k2 = false;   // the next pipelined launch must not be programmatically serialised behind this one
    CK(cudaGetLastError());
  }
  return 1;
}

int nrx_plan_run(nrx_engine *e, uint32_t plan_id) {
  if (!e || plan_id >= e->plans.size() || !e->plans[plan_id].alive) { g_err = "nrx_plan_run: no such plan"; return 0; }
  CK(cudaSetDevice(e->device));
  if (!flush_pmatrices(e)) return 0;
  EnginePlan &pl = e->plans[plan_id];
  if (pl.sizes.empty()) return 1;
  if (!refresh_views(e)) return 0;
  cudaEvent_t ev0, ev1;
  prof_begin(e, &ev0, &ev1);
  unsigned long long per_run = 0;
  for (size_t b = 0; b < pl.sizes.size(); ++b) per_run += (pl.sizes[b] ? e->classes.size() : 0) + (pl.node_block_cnt[b] ? 1 : 0);
  cudaGraphExec_t &gexec = pl.exec[(e->throughput_mode ? 1 : 0) + (e->score_only ? 2 : 0)];
  if (e->use_graphs && !gexec) {  // capture the launches once per geometry; kernel arguments (views, resident ops) never change
    const unsigned long long l0 = e->launches;
    CK(cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal));
    // PDL only where every launch of the plan is the pipelined 4-state kernel (one shape class): the chain K2 -> K2 -> ...
    e->capturing_pdl = e->use_pdl && e->classes.size() == 1 &&
                       ((dna_pipe_cats(e->classes[0].states, e->classes[0].cats) && e->k2_variant == 0) || (aa_dmma_class(e, e->classes[0]) && !e->aa_v1));
    e->pdl_prev_is_k2 = false;
    int ok = 1;
    for (size_t b = 0; b < pl.sizes.size() && ok; ++b) ok = launch_plan_batch(e, pl, b);
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(e->stream, &graph);
    e->capturing_pdl = false;
    e->launches = l0;
    if (!ok || ce != cudaSuccess) { if (graph) cudaGraphDestroy(graph); if (ok) cuda_ok(ce, "cudaStreamEndCapture"); cudaGetLastError(); return 0; }
    const cudaError_t ie = cudaGraphInstantiate(&gexec, graph, 0);
    cudaGraphDestroy(graph);
    if (!cuda_ok(ie, "cudaGraphInstantiate")) return 0;
  }
  if (gexec) {
    CK(cudaGraphLaunch(gexec, e->stream));
    e->launches += per_run;
  } else {
    for (size_t b = 0; b < pl.sizes.size(); ++b)
      if (!launch_plan_batch(e, pl, b)) return 0;
  }
  prof_end(e, ev0, ev1, per_run, pl.updates, pl.bytes, NRX_PROF_K2, pl.cbytes);
  return 1;
}

int nrx_set_throughput_mode(nrx_engine *e, int on) {
  if (!e) { g_err = "null engine"; return 0; }
  e->throughput_mode = on != 0;
  return 1;
}

int nrx_set_score_only(nrx_engine *e, int on) {
  if (!e) { g_err = "null engine"; return 0; }
  e->score_only = on != 0;
  return 1;
}

int nrx_supports_fused_lnl(nrx_engine *e) {
  if (!e || (e->k2_variant != 0 && e->k2_variant != 1) || std::getenv("NRX_NO_FUSED_LNL")) return 0;
  for (const ShapeClass &c : e->classes)
    if (!(dna_pipe_cats(c.states, c.cats) && (c.cats == 4 || e->k2_variant == 0)) && !(aa_dmma_class(e, c) && !e->aa_v1 && std::getenv("NRX_AA_FUSED_LNL"))) return 0;   // 20 states: built, opt-in — the storer warp's extra 20 x 4
                                                                                           // multiply-adds per pattern cost K2 what the saved K3 pass gains (0.421 vs 0.422 ms at 20 k patterns, 2.813 vs 2.802 ms at 200 k; gpurun_out/r4e_*)
  for (const Part &p : e->parts) if (p.pinv > 0.0 || p.nmodels > 1) return 0;   // the K2 epilogue carries neither the invariant-site term nor per-category frequencies
  return 1;
}

int nrx_plan_destroy(nrx_engine *e, uint32_t plan_id) {
  if (!e || plan_id >= e->plans.size()) { g_err = "nrx_plan_destroy: no such plan"; return 0; }
  CK(cudaSetDevice(e->device));
  EnginePlan &pl = e->plans[plan_id];
  if (!pl.alive) return 1;
  CK(cudaStreamSynchronize(e->stream));
  for (cudaGraphExec_t &x : pl.exec) if (x) cudaGraphExecDestroy(x);
  cudaFree(pl.d_ops);
  cudaFree(pl.d_walk);
  cudaFree(pl.d_ngroups); cudaFree(pl.d_nops); cudaFree(pl.d_nblocks);
  pl = EnginePlan();
  return 1;
}

/* a partition of this class mixes several rate matrices over its categories (LG4M / LG4X): K3-K6 take the generic kernels,
 * which index the model by category; every fast kernel assumes one matrix */
static bool class_mixture(const nrx_engine *e, const ShapeClass &c) {
  for (uint32_t pi : c.parts) if (e->parts[pi].nmodels > 1) return true;
  return false;
}
static bool pow2_cats(const ShapeClass &c) { return c.cats <= 32 && (c.cats & (c.cats - 1)) == 0 && !std::getenv("NRX_NO_PC"); }

/* blocks along x for the reduction kernels (one value per call: the partial-sum layout uses gridDim.x as its stride) */
static uint32_t reduce_blocks(const nrx_engine *e, uint32_t items, bool quad_kernels = true) {
  // work items per (tree / pair): patterns for the thread-per-pattern kernels, patterns x categories where the class
  // runs thread-per-(pattern, category) or, for the 20-state tensor-core kernels, tile groups
  uint64_t work = 0;
  for (const ShapeClass &c : e->classes) {
    const bool dna4 = c.states == 4 && c.cats == 4;
    work = std::max<uint64_t>(work, (uint64_t)c.max_patterns * ((!dna4 && pow2_cats(c)) ? c.cats : 1));
  }
  const uint64_t full = std::max<uint64_t>(1, (work + BLOCK - 1) / BLOCK);
  // the quad kernels (DNA 4x4) prefetch their next pass: few long-lived blocks (quad_total, default 2 per SM = one resident wave; measured 296 / 592 / 1184 / 2368 blocks: K6 0.76 / 0.71 / 0.62 / 0.54 of the HBM peak); everything else
  // one pass per block where possible, ~32 blocks per SM in total
  bool all_quad = quad_kernels && e->quad && !e->classes.empty();
  for (const ShapeClass &c : e->classes) all_quad = all_quad && ((c.states == 4 && c.cats == 4) || (c.states == 20 && c.cats == 4 && pow2_cats(c))) && !class_mixture(e, c);
  const uint64_t total = all_quad ? (e->quad_total ? e->quad_total : 2ull * e->sm_count) : 148ull * 32;
  const uint64_t want = std::max<uint64_t>(1, total / std::max<uint32_t>(1, items));
  if (full <= want) return (uint32_t)full;
  const uint64_t passes = (full + want - 1) / want;       // every block makes the same number of passes (the last one maybe one less):
  return (uint32_t)((full + passes - 1) / passes);        // 391 chunks over 315 blocks would leave 239 blocks idle during the second pass

}

/* Where the last block of a reducing kernel writes the reduced values.  Without a communicator and with the fused second stage that is
 * the pinned host buffer itself (mapped into the device's address space): the result needs no device->host copy, one stream operation
 * and ~3 us less per synchronous call — the derivative sweep makes 440 of them on config 2.  Env NRX_ZEROCOPY=0: device buffer + copy. */
static double *result_out(nrx_engine *e, const uint32_t *tk) {
  return (tk && e->zero_copy && !e->comm && e->d_result_map) ? e->d_result_map : e->d_result;
}

/* second stage + cross-rank sum + device->host copy, all stream-ordered; the host blocks only in wait_result */
static int enqueue_reduction(nrx_engine *e, uint32_t total, uint32_t nblk, bool fused = false) {
  if (fused && result_out(e, e->d_tickets) != e->d_result) { e->pending_result = total; return 1; }   // the kernel wrote h_result directly
  if (!fused) {   // the second stage was not done by the last block of the reducing kernel itself
    cudaEvent_t ev0, ev1;
    prof_begin(e, &ev0, &ev1);
    k_reduce_partials<<<(total + 3) / 4, 128, 0, e->stream>>>(e->d_partial, e->d_result, nblk, total);
    e->launches++;
    CK(cudaGetLastError());
    prof_end(e, ev0, ev1, 1, total, (unsigned long long)total * nblk * 8, NRX_PROF_REDUCE);
  }
  if (e->comm) {  // C2-C4: one all-reduce over NVLink for all trees / pairs x partitions
    const int rc = nccl().AllReduce(e->d_result, e->d_result, total, NCCL_FLOAT64, NCCL_SUM, e->comm, e->stream);
    if (rc != 0) { g_err = std::string("ncclAllReduce: ") + nccl().GetErrorString(rc); return 0; }
  }
  CK(cudaMemcpyAsync(e->h_result, e->d_result, (size_t)total * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
  e->pending_result = total;
  return 1;
}
static int check_async_error(nrx_engine *e) {
  if (e->h_err && *e->h_err) { *e->h_err = 0; g_err = "Illegal state code in tip (asynchronous tip upload)"; return 0; }
  return 1;
}
static int wait_result(nrx_engine *e, uint32_t total, double *out) {
  CK(cudaStreamSynchronize(e->stream));
  if (!check_async_error(e)) { e->pending_result = 0; return 0; }
  if (out) std::memcpy(out, e->h_result, (size_t)total * sizeof(double));
  e->pending_result = 0;
  return 1;
}
static int finish_reduction(nrx_engine *e, uint32_t total, uint32_t nblk, double *out, bool fused = false) {
  return enqueue_reduction(e, total, nblk, fused) && wait_result(e, total, out);
}

static int tree_lnl_impl(nrx_engine *e, const uint32_t *slots, uint32_t n, double *out, double *persite, size_t persite_stride, bool async) {
  if (!e) { g_err = "null engine"; return 0; }
  if (n == 0) return 1;
  CK(cudaSetDevice(e->device));
  for (uint32_t i = 0; i < n; ++i) if (slots[i] >= e->nslots) { g_err = "nrx_tree_lnl: slot out of range"; return 0; }
  const uint32_t P = (uint32_t)e->parts.size();
  const uint32_t nblk = reduce_blocks(e, n * P);
  if (!ensure_result(e, (size_t)n * P, (size_t)n * P * nblk) || !refresh_views(e)) return 0;
  uint32_t *d_slots;
  if (!upload(e, slots, n, &d_slots)) return 0;
  double *d_ps = nullptr;
  if (persite) {
    const size_t need = (size_t)n * P * persite_stride;
    if (need > e->persite_cap) { cudaFree(e->d_persite); CK(cudaMalloc((void **)&e->d_persite, need * sizeof(double))); e->persite_cap = need; }
    CK(cudaMemsetAsync(e->d_persite, 0, need * sizeof(double), e->stream));
    d_ps = e->d_persite;
  }
  // every block of the launches below writes its partial sum (blocks without patterns write 0): no memset needed
  const double log_thresh = std::log(SCALE_THRESHOLD);
  uint32_t *tk = e->fuse_reduce ? e->d_tickets : nullptr;
  cudaEvent_t ev0, ev1;
  prof_begin(e, &ev0, &ev1);
  for (const ShapeClass &c : e->classes) {
    dim3 grid(nblk, n, (uint32_t)c.parts.size());
    const bool mix = class_mixture(e, c);
    if (!mix && c.states == 4 && c.cats == 4 && e->quad) k_tree_lnl_dna4q<<<grid, BLOCK, 0, e->stream>>>(c.d_views, d_slots, e->d_partial, P, log_thresh, d_ps, persite_stride, result_out(e, tk), tk);
    else if (!mix && c.states == 4 && c.cats == 4) k_tree_lnl_dna4<<<grid, BLOCK, 0, e->stream>>>(c.d_views, d_slots, e->d_partial, P, log_thresh, d_ps, persite_stride, result_out(e, tk), tk);
    else if (!mix && pow2_cats(c) && c.states == 20 && c.cats == 4 && e->quad) k_tree_lnl_aa20p<<<grid, BLOCK, 0, e->stream>>>(c.d_views, d_slots, e->d_partial, P, log_thresh, d_ps, persite_stride, result_out(e, tk), tk);
    else if (!mix && pow2_cats(c) && c.states == 20) k_tree_lnl_pc<20><<<grid, BLOCK, 0, e->stream>>>(c.d_views, d_slots, e->d_partial, P, log_thresh, d_ps, persite_stride, result_out(e, tk), tk);
    else if (!mix && pow2_cats(c)) k_tree_lnl_pc<0><<<grid, BLOCK, 0, e->stream>>>(c.d_views, d_slots, e->d_partial, P, log_thresh, d_ps, persite_stride, result_out(e, tk), tk);
    else k_tree_lnl<<<grid, BLOCK, 0, e->stream>>>(c.d_views, d_slots, e->d_partial, P, log_thresh, d_ps, persite_stride, result_out(e, tk), tk);
    e->launches++;
    CK(cudaGetLastError());
  }
  { unsigned long long u = 0; const unsigned long long b = stream_bytes(e, n, 1, persite ? 16 : 8, &u); prof_end(e, ev0, ev1, e->classes.size(), u, b, NRX_PROF_K3); }
  if (async) return enqueue_reduction(e, n * P, nblk, tk != nullptr);
  if (!finish_reduction(e, n * P, nblk, out, tk != nullptr)) return 0;
  if (persite) CK(cudaMemcpy(persite, e->d_persite, (size_t)n * P * persite_stride * sizeof(double), cudaMemcpyDeviceToHost));
  return 1;
}
int nrx_tree_lnl(nrx_engine *e, const uint32_t *slots, uint32_t n, double *out, double *persite, size_t persite_stride) {
  return tree_lnl_impl(e, slots, n, out, persite, persite_stride, false);
}
int nrx_tree_lnl_async(nrx_engine *e, const uint32_t *slots, uint32_t n) { return tree_lnl_impl(e, slots, n, nullptr, nullptr, 0, true); }
int nrx_result_wait(nrx_engine *e, double *out, uint32_t count) {
  if (!e) { g_err = "null engine"; return 0; }
  if (count != e->pending_result) { g_err = "nrx_result_wait: no pending result of that size"; return 0; }
  CK(cudaSetDevice(e->device));
  return wait_result(e, count, out);
}

static int tree_lnl_fused_impl(nrx_engine *e, uint32_t plan_id, const uint32_t *slots, uint32_t n, double *out, bool async) {
  if (!e || plan_id >= e->plans.size() || !e->plans[plan_id].alive) { g_err = "nrx_tree_lnl_fused: no such plan"; return 0; }
  if (n == 0) return 1;
  if (n > e->plans[plan_id].lnl_items || !e->d_fused) { g_err = "nrx_tree_lnl_fused: the plan carries fewer lnl marks"; return 0; }
  CK(cudaSetDevice(e->device));
  const uint32_t P = (uint32_t)e->parts.size();
  uint32_t nblk;
  {  // k_term_lnl_sum: a pass = 4 x BLOCK patterns, ~8 blocks per SM in total, every block the same number of passes
    const uint64_t chunks = std::max<uint64_t>(1, ((uint64_t)e->max_patterns + 4ull * BLOCK - 1) / (4ull * BLOCK));
    const uint64_t want = std::max<uint64_t>(1, (8ull * e->sm_count) / std::max<uint32_t>(1, n * P));
    const uint64_t passes = (chunks + want - 1) / want;
    nblk = (uint32_t)((chunks + passes - 1) / passes);
  }
  if (!ensure_result(e, (size_t)n * P, (size_t)n * P * nblk) || !refresh_views(e)) return 0;
  for (uint32_t i = 0; i < n; ++i) if (slots[i] >= e->nslots) { g_err = "nrx_tree_lnl_fused: slot out of range"; return 0; }
  uint32_t *d_slots;
  if (!upload(e, slots, n, &d_slots)) return 0;
  uint32_t *tk = e->fuse_reduce ? e->d_tickets : nullptr;
  cudaEvent_t ev0, ev1;
  prof_begin(e, &ev0, &ev1);
  for (const ShapeClass &c : e->classes) {
    dim3 grid(nblk, n, (uint32_t)c.parts.size());
    k_term_lnl_sum<<<grid, BLOCK, 0, e->stream>>>(c.d_views, d_slots, e->d_fused, (size_t)e->max_patterns, e->d_partial, P, std::log(SCALE_THRESHOLD), result_out(e, tk), tk);
    e->launches++;
    CK(cudaGetLastError());
  }
  { unsigned long long u = 0; const unsigned long long b = stream_bytes(e, n, 0, 16, &u); prof_end(e, ev0, ev1, e->classes.size(), u, b, NRX_PROF_K3F); }
  return async ? enqueue_reduction(e, n * P, nblk, tk != nullptr) : finish_reduction(e, n * P, nblk, out, tk != nullptr);
}
int nrx_tree_lnl_fused(nrx_engine *e, uint32_t plan_id, const uint32_t *slots, uint32_t n, double *out) {
  return tree_lnl_fused_impl(e, plan_id, slots, n, out, false);
}
int nrx_tree_lnl_fused_async(nrx_engine *e, uint32_t plan_id, const uint32_t *slots, uint32_t n) {
  return tree_lnl_fused_impl(e, plan_id, slots, n, nullptr, true);
}

/* Full evaluation of a plan whose root trees carry lnl marks: every CLV of the traversal + the n per-tree root lnLs, result
 * enqueued like nrx_tree_lnl_fused_async.  One k_walk_dna4 launch when the plan has a tile-walk form and the launch is small
 * enough to be latency-bound (or NRX_WALK=1), otherwise nrx_plan_run + nrx_tree_lnl_fused_async. */
int nrx_plan_evaluate_async(nrx_engine *e, uint32_t plan_id, const uint32_t *slots, uint32_t n) {
  if (!e || plan_id >= e->plans.size() || !e->plans[plan_id].alive) { g_err = "nrx_plan_evaluate: no such plan"; return 0; }
  EnginePlan &pl = e->plans[plan_id];
  const uint32_t ntiles = (e->max_patterns + WALK_TP - 1) / WALK_TP;
  const bool walk = pl.d_walk && n == pl.lnl_items && n > 0 && !e->throughput_mode && e->fuse_reduce &&
                    (e->walk_mode == 1 || (e->walk_mode == 2 && (uint64_t)ntiles * e->parts.size() <= e->walk_max_tiles));
  if (!walk) return nrx_plan_run(e, plan_id) && nrx_tree_lnl_fused_async(e, plan_id, slots, n);
  CK(cudaSetDevice(e->device));
  const uint32_t P = (uint32_t)e->parts.size();
  if (!ensure_result(e, (size_t)n * P, (size_t)n * P * ntiles) || !refresh_views(e)) return 0;
  for (uint32_t i = 0; i < n; ++i) if (slots[i] >= e->nslots) { g_err = "nrx_plan_evaluate: slot out of range"; return 0; }
  const ShapeClass &c = e->classes[0];
  // deferred P-matrix updates: the walk computes ALL P-matrices itself from the current branch lengths (K1 fused into the launch)
  int compute_p = 0;
  double *d_len = nullptr;
  {
    bool pending = false, complete = true;
    for (const Part &p : e->parts) { pending = pending || p.npend; for (double t : p.h_len) complete = complete && t >= 0.0; }
    if (pending && complete && (size_t)pl.walk_nbuf * WALK_TP * 16 >= (size_t)e->parts[0].d.edges * 16) {
      std::vector<double> lens;
      for (Part &p : e->parts) { lens.insert(lens.end(), p.h_len.begin(), p.h_len.end()); std::fill(p.pend.begin(), p.pend.end(), 0); p.npend = 0; }
      if (!upload(e, lens.data(), lens.size(), &d_len)) return 0;
      compute_p = 1;
    } else if (pending && !flush_pmatrices(e)) return 0;
  }
  cudaEvent_t ev0, ev1;
  prof_begin(e, &ev0, &ev1);
  k_walk_dna4<<<dim3(ntiles, 1, (uint32_t)c.parts.size()), WALK_BLOCK, pl.walk_smem, e->stream>>>(
      c.d_views, pl.d_walk, pl.walk_nops, pl.walk_nbuf, n, std::log(SCALE_THRESHOLD), e->d_partial, P, result_out(e, e->d_tickets), e->d_tickets, compute_p, d_len);
  e->launches++;
  CK(cudaGetLastError());
  prof_end(e, ev0, ev1, 1, pl.updates, pl.bytes, NRX_PROF_K2, pl.walk_cbytes);
  return enqueue_reduction(e, n * P, ntiles, true);
}

static int check_pairs(nrx_engine *e, const nrx_pair *pairs, uint32_t n, const char *who) {
  for (uint32_t i = 0; i < n; ++i) {
    const nrx_pair &q = pairs[i];
    if (q.a_kind == NRX_TIP && q.b_kind == NRX_TIP) { g_err = std::string(who) + " was called for the tip-tip case!"; return 0; }
    const uint32_t kinds[2] = {q.a_kind, q.b_kind}, idx[2] = {q.a_idx, q.b_idx};
    for (int s = 0; s < 2; ++s) {
      if (kinds[s] == NRX_CLV) { if (idx[s] >= e->nslots) { g_err = std::string(who) + ": slot out of range"; return 0; } }
      else if (kinds[s] == NRX_TIP) { for (const Part &p : e->parts) if (idx[s] >= p.d.tips) { g_err = std::string(who) + ": tip out of range"; return 0; } }
      else { g_err = std::string(who) + ": operand must be a CLV slot or a tip"; return 0; }
    }
  }
  return 1;
}

/* 20-state pairs as pseudo-ops of the DMMA kernel.  K4 (edge lnL): left = the parent CLV as it lies, right = the child
 * (CLV or tip, the tip always plays child: LIBPLL/likelihood.c:586-601) through P(edge).  K5 (sumtable): left = the
 * tip if there is one (LIBPLL/derivatives.c:70-98), parent_slot = sumtable index. */
static bool aa_dmma_class(const nrx_engine *e, const ShapeClass &c) {
  return c.states == 20 && c.cats == 4 && !e->aa_generic && class_tip_codes(e, c) <= (uint32_t)AA_LUT_CODES;
}
/* K4 / K5 on the tensor cores use category-independent frequency / eigen-matrix operands: single-matrix partitions only
 * (K2's AA_CLV mode reads only P-matrices and keeps running for mixtures) */
static bool aa_dmma_pairs_class(const nrx_engine *e, const ShapeClass &c) { return aa_dmma_class(e, c) && !class_mixture(e, c); }
static std::vector<nrx_op> pairs_to_ops(const nrx_pair *pairs, uint32_t n, uint32_t edge, bool tip_left, bool *any_tip_out) {
  std::vector<nrx_op> ops(n);
  *any_tip_out = false;
  for (uint32_t i = 0; i < n; ++i) {
    nrx_pair q = pairs[i];
    const bool swap = tip_left ? (q.b_kind == NRX_TIP) : (q.a_kind == NRX_TIP);
    if (swap) { std::swap(q.a_kind, q.b_kind); std::swap(q.a_idx, q.b_idx); }
    nrx_op o{};
    o.parent_slot = i;
    o.left_kind = q.a_kind; o.left_idx = q.a_idx; o.left_edge = edge;
    o.right_kind = q.b_kind; o.right_idx = q.b_idx; o.right_edge = edge;
    ops[i] = o;
    *any_tip_out |= (q.a_kind == NRX_TIP || q.b_kind == NRX_TIP);
  }
  return ops;
}

int nrx_edge_lnl(nrx_engine *e, uint32_t edge, const nrx_pair *pairs, uint32_t n, double *out) {
  if (!e) { g_err = "null engine"; return 0; }
  if (n == 0) return 1;
  CK(cudaSetDevice(e->device));
  if (!flush_pmatrices(e)) return 0;
  if (!check_pairs(e, pairs, n, "nrx_edge_lnl")) return 0;
  for (const Part &p : e->parts) if (edge >= p.d.edges) { g_err = "nrx_edge_lnl: edge out of range"; return 0; }
  const uint32_t P = (uint32_t)e->parts.size();
  uint32_t nblk = reduce_blocks(e, n * P);
  {  // engines whose partitions all run the tensor-core kernel: nblk = tile groups per pair; keep >= ~50 tiles per block so
     // that the per-block set-up (B fragments, tip table) is amortised (4736 blocks of 21 tiles ran at 7 % tensor pipe)
    bool only_aa = !e->classes.empty();
    for (const ShapeClass &c : e->classes) only_aa = only_aa && aa_dmma_pairs_class(e, c);
    if (only_aa && e->aa_v1) nblk = std::max<uint32_t>(1, std::min<uint32_t>(nblk, e->aa_blocks / std::max<uint32_t>(1, n * P)));
    else if (only_aa) {   // k_aa20_mma: 2 resident blocks per SM; 2 or 4 whole waves depending on the launch size (as for K2)
      const uint64_t ntiles = (e->max_patterns + AA_TP - 1) / AA_TP;
      const uint64_t blocks = e->aa2_blocks ? e->aa2_blocks : std::min<uint64_t>(8ull * e->sm_count, std::max<uint64_t>(4ull * e->sm_count, (uint64_t)n * P * ntiles / 48));
      nblk = (uint32_t)std::max<uint64_t>(1, blocks / std::max<uint32_t>(1, n * P));
    }
  }
  if (!ensure_result(e, (size_t)n * P, (size_t)n * P * nblk) || !refresh_views(e)) return 0;
  nrx_pair *d_pairs;
  if (!upload(e, pairs, n, &d_pairs)) return 0;
  // the tensor-core edge kernel lets surplus blocks exit before they write a partial sum: that path keeps the memset and
  // the separate second stage; every other kernel writes all partials and finishes the reduction in its last block
  bool any_aa = false;
  for (const ShapeClass &c : e->classes) any_aa = any_aa || aa_dmma_pairs_class(e, c);
  if (any_aa)   // the tensor-core kernels hand tiles to blocks: never more blocks per pair than tiles (every block writes its partial)
    for (const ShapeClass &c : e->classes) if (c.max_patterns) nblk = std::min<uint32_t>(nblk, (c.max_patterns + AA_TP - 1) / AA_TP);
  const bool aa_old = any_aa && e->aa_v1;   // k_aa20_dmma<AA_EDGE> lets surplus blocks exit early: memset + separate second stage
  uint32_t *tk = (e->fuse_reduce && !aa_old) ? e->d_tickets : nullptr;
  if (aa_old) CK(cudaMemsetAsync(e->d_partial, 0, (size_t)n * P * nblk * sizeof(double), e->stream));
  const double log_thresh = std::log(SCALE_THRESHOLD);
  cudaEvent_t ev0, ev1;
  prof_begin(e, &ev0, &ev1);
  for (const ShapeClass &c : e->classes) {
    dim3 grid(nblk, n, (uint32_t)c.parts.size());
    if (c.states == 4 && c.cats == 4 && !class_mixture(e, c) && e->quad) k_edge_lnl_dna4q<<<grid, BLOCK, 0, e->stream>>>(c.d_views, d_pairs, edge, e->d_partial, P, log_thresh, result_out(e, tk), tk);
    else if (c.states == 4 && c.cats == 4 && !class_mixture(e, c)) k_edge_lnl_dna4<<<grid, BLOCK, 0, e->stream>>>(c.d_views, d_pairs, edge, e->d_partial, P, log_thresh, result_out(e, tk), tk);
    else if (aa_dmma_pairs_class(e, c)) {  // FP64 tensor cores: block b = (pair b % n, tile group b / n), one partial per (pair, group)
      bool tips;
      const std::vector<nrx_op> ops = pairs_to_ops(pairs, n, edge, false, &tips);
      nrx_op *d_ops;
      if (!upload(e, ops.data(), ops.size(), &d_ops)) return 0;
      const size_t smem = sizeof(AaSmem) + (tips ? (size_t)class_tip_codes(e, c) * 80 * sizeof(double) : 0);   // pairs are never tip-tip: one table
      if (e->aa_v1) k_aa20_dmma<AA_EDGE><<<dim3(n * nblk, 1, (uint32_t)c.parts.size()), AA_THREADS, smem, e->stream>>>(c.d_views, d_ops, n, nblk, tips ? 1 : 0, e->d_partial, P, log_thresh);
      else k_aa20_mma<AA_EDGE, true><<<dim3(n * nblk, 1, (uint32_t)c.parts.size()), AA2_THREADS, sizeof(AaSmem2) + (tips ? (size_t)class_tip_codes(e, c) * 80 * sizeof(double) : 0), e->stream>>>(
            c.d_views, d_ops, n, nblk, tips ? 1 : 0, e->d_partial, P, log_thresh, result_out(e, tk), tk, 0);
    }
    else k_edge_lnl<<<grid, BLOCK, 0, e->stream>>>(c.d_views, d_pairs, edge, e->d_partial, P, log_thresh, result_out(e, tk), tk);
    e->launches++;
    CK(cudaGetLastError());
  }
  { unsigned long long u = 0; const unsigned long long b = stream_bytes(e, n, 2, 12, &u); prof_end(e, ev0, ev1, e->classes.size(), u, b, NRX_PROF_K4, pair_cbytes(e, pairs, n, true, 0, 4)); }
  return finish_reduction(e, n * P, nblk, out, tk != nullptr);
}

int nrx_sumtables(nrx_engine *e, const nrx_pair *pairs, uint32_t n) {
  if (!e) { g_err = "null engine"; return 0; }
  if (n == 0) return 1;
  CK(cudaSetDevice(e->device));
  if (!check_pairs(e, pairs, n, "pll_update_sumtable()")) return 0;
  if (!reserve_sumtables(e, n) || !refresh_views(e)) return 0;
  nrx_pair *d_pairs;
  if (!upload(e, pairs, n, &d_pairs)) return 0;
  cudaEvent_t ev0, ev1;
  prof_begin(e, &ev0, &ev1);
  for (const ShapeClass &c : e->classes) {
    if (c.max_patterns == 0) continue;
    const uint32_t z = (uint32_t)c.parts.size();
    if (c.states == 4 && c.cats == 4 && !class_mixture(e, c)) {
      dim3 grid(tiles_for((uint64_t)c.max_patterns * 4, BLOCK * RU, n * z), n, z);
      k_sumtable_dna4<<<grid, BLOCK, 0, e->stream>>>(c.d_views, d_pairs);
    } else if (aa_dmma_pairs_class(e, c)) {
      bool tips;
      const std::vector<nrx_op> ops = pairs_to_ops(pairs, n, 0, true, &tips);
      nrx_op *d_ops;
      if (!upload(e, ops.data(), ops.size(), &d_ops)) return 0;
      const uint32_t ntiles = (c.max_patterns + AA_TP - 1) / AA_TP;
      const uint32_t by_size = (uint32_t)std::min<uint64_t>(8ull * e->sm_count, std::max<uint64_t>(4ull * e->sm_count, (uint64_t)n * z * ntiles / 48));
      uint32_t groups = std::max<uint32_t>(1, ((e->aa_v1 ? e->aa_blocks : (e->aa2_blocks ? e->aa2_blocks : by_size)) + n * z - 1) / (n * z));
      groups = std::min(groups, std::max<uint32_t>(1, ntiles / 8));
      const size_t lut_bytes = tips ? (size_t)class_tip_codes(e, c) * 80 * sizeof(double) : 0;
      if (e->aa_v1) k_aa20_dmma<AA_SUM><<<dim3(n * groups, 1, z), AA_THREADS, sizeof(AaSmem) + lut_bytes, e->stream>>>(c.d_views, d_ops, n, groups, tips ? 1 : 0, nullptr, 0, 0.0);
      else k_aa20_mma<AA_SUM, true><<<dim3(n * groups, 1, z), AA2_THREADS, sizeof(AaSmem2) + lut_bytes, e->stream>>>(c.d_views, d_ops, n, groups, tips ? 1 : 0, nullptr, 0, 0.0, nullptr, nullptr, 0);
    } else {
      dim3 grid(tiles_for((uint64_t)c.max_patterns * c.cats, BLOCK, n * z), n, z);
      k_sumtable<<<grid, BLOCK, 0, e->stream>>>(c.d_views, d_pairs);
    }
    e->launches++;
    CK(cudaGetLastError());
  }
  { unsigned long long u = 0; const unsigned long long b = stream_bytes(e, n, 3, 0, &u); prof_end(e, ev0, ev1, e->classes.size(), u, b, NRX_PROF_K5, pair_cbytes(e, pairs, n, false, 1, 0)); }
  return 1;
}

/* K4 + K5 in one pass over the pairs' CLVs (k_edge_sum_dna4q); engines / pairs the fused kernel does not cover take the two calls */
int nrx_edge_lnl_sumtables(nrx_engine *e, uint32_t edge, const nrx_pair *pairs, uint32_t n, const int32_t *lnl_index, uint32_t n_lnl, double *out) {
  if (!e) { g_err = "null engine"; return 0; }
  if (n == 0) return 1;
  for (uint32_t i = 0; i < n; ++i) if (lnl_index[i] >= (int32_t)n_lnl) { g_err = "nrx_edge_lnl_sumtables: lnl index out of range"; return 0; }
  bool fusable = e->quad && !std::getenv("NRX_NO_EDGE_SUM");
  for (const ShapeClass &c : e->classes) fusable = fusable && c.states == 4 && c.cats == 4 && !class_mixture(e, c);
  for (uint32_t i = 0; i < n; ++i) fusable = fusable && pairs[i].a_kind == NRX_CLV;
  if (!fusable) {
    if (!nrx_sumtables(e, pairs, n)) return 0;
    if (n_lnl == 0) return 1;
    std::vector<nrx_pair> sub(n_lnl);
    for (uint32_t i = 0; i < n; ++i) if (lnl_index[i] >= 0) sub[lnl_index[i]] = pairs[i];
    return nrx_edge_lnl(e, edge, sub.data(), n_lnl, out);
  }
  CK(cudaSetDevice(e->device));
  if (!flush_pmatrices(e)) return 0;
  if (!check_pairs(e, pairs, n, "nrx_edge_lnl_sumtables")) return 0;
  for (const Part &p : e->parts) if (edge >= p.d.edges) { g_err = "nrx_edge_lnl_sumtables: edge out of range"; return 0; }
  const uint32_t P = (uint32_t)e->parts.size();
  const uint32_t nblk = reduce_blocks(e, n * P);
  if (!reserve_sumtables(e, n)) return 0;
  if (!ensure_result(e, (size_t)std::max<uint32_t>(1, n_lnl) * P, (size_t)std::max<uint32_t>(1, n_lnl) * P * nblk) || !refresh_views(e)) return 0;
  nrx_pair *d_pairs;
  int32_t *d_idx;
  if (!upload(e, pairs, n, &d_pairs) || !upload(e, lnl_index, n, &d_idx)) return 0;
  uint32_t *tk = e->fuse_reduce ? e->d_tickets : nullptr;
  const double log_thresh = std::log(SCALE_THRESHOLD);
  cudaEvent_t ev0, ev1;
  prof_begin(e, &ev0, &ev1);
  for (const ShapeClass &c : e->classes) {
    if (c.max_patterns == 0) continue;
    dim3 grid(nblk, n, (uint32_t)c.parts.size());
    k_edge_sum_dna4q<<<grid, BLOCK, 0, e->stream>>>(c.d_views, d_pairs, d_idx, edge, e->d_partial, P, log_thresh, result_out(e, tk), tk);
    e->launches++;
    CK(cudaGetLastError());
  }
  { unsigned long long u = 0; const unsigned long long b = stream_bytes(e, n, 3, 12, &u); prof_end(e, ev0, ev1, e->classes.size(), u, b, NRX_PROF_K45, pair_cbytes(e, pairs, n, true, 1, 4)); }
  if (n_lnl == 0) return 1;
  return finish_reduction(e, n_lnl * P, nblk, out, tk != nullptr);
}

int nrx_derivatives(nrx_engine *e, uint32_t n, const double *brlen, double *out) {
  if (!e) { g_err = "null engine"; return 0; }
  if (n == 0) return 1;
  CK(cudaSetDevice(e->device));
  if (n > e->nsumtables) { g_err = "nrx_derivatives: sumtables not computed"; return 0; }
  const uint32_t P = (uint32_t)e->parts.size();
  const uint32_t nblk = reduce_blocks(e, n * P);
  if (!ensure_result(e, (size_t)n * P * 3, (size_t)n * P * 3 * nblk) || !refresh_views(e)) return 0;
  // diagptable on the host with libm exp, exactly pll_compute_diagptable (LIBPLL/core_derivatives.c:711-726)
  for (uint32_t pi = 0; pi < P; ++pi) {
    Part &p = e->parts[pi];
    const uint32_t S = p.d.states, C = p.d.rate_cats;
    std::vector<double> diag((size_t)C * S * 4);
    double *dp = diag.data();
    for (uint32_t i = 0; i < C; ++i) {
      const double ki = p.h_rates[i] / (1.0 - p.pinv);   // LIBPLL/core_derivatives.c:716
      const double *ev = p.h_eigenvals.data() + (size_t)(p.nmodels > 1 ? p.cat_model[i] : 0) * S;   // eigenvals[params_indices[i]] (:712)
      for (uint32_t j = 0; j < S; ++j) {
        dp[0] = std::exp(ev[j] * ki * brlen[pi]);
        dp[1] = ev[j] * ki * dp[0];
        dp[2] = ev[j] * ki * ev[j] * ki * dp[0];
        dp[3] = 0;
        dp += 4;
      }
    }
    void *h_stage, *d_unused;   // pinned staging -> the partition's table in ONE copy (330 calls per derivative sweep)
    if (!stage_alloc(e, diag.size() * sizeof(double), &h_stage, &d_unused)) return 0;
    std::memcpy(h_stage, diag.data(), diag.size() * sizeof(double));
    CK(cudaMemcpyAsync(p.diagp, h_stage, diag.size() * sizeof(double), cudaMemcpyHostToDevice, e->stream));
  }
  uint32_t *tk = e->fuse_reduce ? e->d_tickets : nullptr;
  cudaEvent_t ev0, ev1;
  prof_begin(e, &ev0, &ev1);
  for (const ShapeClass &c : e->classes) {
    dim3 grid(nblk, n, (uint32_t)c.parts.size());
    const bool mix = class_mixture(e, c);   // (only the +I term of K6 reads frequencies; the generic kernel is the one that indexes them by category)
    if (!mix && c.states == 4 && c.cats == 4 && e->quad && e->k6_ring) k_derivatives_dna4r<<<grid, BLOCK, sizeof(K6RingSmem), e->stream>>>(c.d_views, e->d_partial, P, result_out(e, tk), tk);
    else if (!mix && c.states == 4 && c.cats == 4 && e->quad) k_derivatives_dna4q<<<grid, BLOCK, 0, e->stream>>>(c.d_views, e->d_partial, P, result_out(e, tk), tk);
    else if (!mix && c.states == 4 && c.cats == 4) k_derivatives_dna4<<<grid, BLOCK, 0, e->stream>>>(c.d_views, e->d_partial, P, result_out(e, tk), tk);
    else if (!mix && pow2_cats(c) && c.states == 20 && c.cats == 4 && e->quad) k_derivatives_aa20p<<<grid, BLOCK, ((size_t)c.cats * diag_stride(c.states) + c.cats) * sizeof(double), e->stream>>>(c.d_views, e->d_partial, P, result_out(e, tk), tk);
    else if (!mix && pow2_cats(c) && c.states == 20) k_derivatives_pc<20, 2><<<grid, BLOCK, ((size_t)c.cats * diag_stride(c.states) + c.cats) * sizeof(double), e->stream>>>(c.d_views, e->d_partial, P, result_out(e, tk), tk);
    else if (!mix && pow2_cats(c)) k_derivatives_pc<0, 2><<<grid, BLOCK, ((size_t)c.cats * diag_stride(c.states) + c.cats) * sizeof(double), e->stream>>>(c.d_views, e->d_partial, P, result_out(e, tk), tk);
    else k_derivatives<<<grid, BLOCK, 0, e->stream>>>(c.d_views, e->d_partial, P, result_out(e, tk), tk);
    e->launches++;
    CK(cudaGetLastError());
  }
  { unsigned long long u = 0; const unsigned long long b = stream_bytes(e, n, 1, 4, &u); prof_end(e, ev0, ev1, e->classes.size(), u, b, NRX_PROF_K6); }
  return finish_reduction(e, n * P * 3, nblk, out, tk != nullptr);
}

int nrx_read_clv(nrx_engine *e, uint32_t pi, uint32_t slot, double *out) {
  if (!check_part(e, pi)) return 0;
  if (slot >= e->nslots) { g_err = "slot out of range"; return 0; }
  CK(cudaSetDevice(e->device));
  CK(cudaStreamSynchronize(e->stream));
  Part &p = e->parts[pi];
  if (p.clv_entries) CK(cudaMemcpy(out, p.h_clv[slot], p.clv_entries * sizeof(double), cudaMemcpyDeviceToHost));
  return 1;
}

int nrx_read_scaler(nrx_engine *e, uint32_t pi, uint32_t slot, uint32_t *out) {
  if (!check_part(e, pi)) return 0;
  if (slot >= e->nslots) { g_err = "slot out of range"; return 0; }
  CK(cudaSetDevice(e->device));
  CK(cudaStreamSynchronize(e->stream));
  Part &p = e->parts[pi];
  if (p.d.patterns) CK(cudaMemcpy(out, p.h_scaler[slot], (size_t)p.d.patterns * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  return 1;
}

int nrx_read_sumtable(nrx_engine *e, uint32_t pi, uint32_t st, double *out) {
  if (!check_part(e, pi)) return 0;
  if (st >= e->nsumtables) { g_err = "sumtable slot out of range"; return 0; }
  CK(cudaSetDevice(e->device));
  CK(cudaStreamSynchronize(e->stream));
  Part &p = e->parts[pi];
  if (p.clv_entries) CK(cudaMemcpy(out, p.h_sumtable[st], p.clv_entries * sizeof(double), cudaMemcpyDeviceToHost));
  return 1;
}

int nrx_sync(nrx_engine *e) {
  if (!e) { g_err = "null engine"; return 0; }
  CK(cudaSetDevice(e->device));
  if (!flush_pmatrices(e)) return 0;
  CK(cudaStreamSynchronize(e->stream));
  return check_async_error(e);
}

int nrx_comm_get_unique_id(uint8_t *id128) {
  if (!nccl().ok) { g_err = "libnccl.so.2 could not be loaded"; return 0; }
  NcclApi::UniqueId id;
  const int rc = nccl().GetUniqueId(&id);
  if (rc != 0) { g_err = std::string("ncclGetUniqueId: ") + nccl().GetErrorString(rc); return 0; }
  std::memcpy(id128, id.internal, 128);
  return 1;
}

int nrx_comm_init(nrx_engine *e, const uint8_t *id128, int rank, int nranks) {
  if (!e) { g_err = "null engine"; return 0; }
  if (nranks < 1 || rank < 0 || rank >= nranks) { g_err = "nrx_comm_init: bad rank / nranks"; return 0; }
  if (e->comm) { g_err = "nrx_comm_init: communicator already attached"; return 0; }
  if (nranks == 1) { e->comm_rank = 0; e->comm_size = 1; return 1; }
  if (!nccl().ok) { g_err = "libnccl.so.2 could not be loaded"; return 0; }
  CK(cudaSetDevice(e->device));
  NcclApi::UniqueId id;
  std::memcpy(id.internal, id128, 128);
  const int rc = nccl().CommInitRank(&e->comm, nranks, id, rank);
  if (rc != 0) { e->comm = nullptr; g_err = std::string("ncclCommInitRank: ") + nccl().GetErrorString(rc); return 0; }
  e->comm_rank = rank; e->comm_size = nranks;
  return 1;
}

int nrx_comm_size(nrx_engine *e) { return e ? e->comm_size : 1; }

int nrx_comm_allreduce_sum(nrx_engine *e, double *host_inout, size_t n) {
  if (!e) { g_err = "null engine"; return 0; }
  if (!e->comm || n == 0) return 1;
  CK(cudaSetDevice(e->device));
  if (!ensure_result(e, n, 0)) return 0;
  std::memcpy(e->h_result, host_inout, n * sizeof(double));
  CK(cudaMemcpyAsync(e->d_result, e->h_result, n * sizeof(double), cudaMemcpyHostToDevice, e->stream));
  const int rc = nccl().AllReduce(e->d_result, e->d_result, n, NCCL_FLOAT64, NCCL_SUM, e->comm, e->stream);
  if (rc != 0) { g_err = std::string("ncclAllReduce: ") + nccl().GetErrorString(rc); return 0; }
  CK(cudaMemcpyAsync(e->h_result, e->d_result, n * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  std::memcpy(host_inout, e->h_result, n * sizeof(double));
  return 1;
}

void *nrx_result_device_ptr(nrx_engine *e) { return e ? e->d_result : nullptr; }
void *nrx_stream(nrx_engine *e) { return e ? (void *)e->stream : nullptr; }
unsigned long long nrx_launch_count(nrx_engine *e) { return e ? e->launches : 0; }

int nrx_timer_start(nrx_engine *e) {
  if (!e) { g_err = "null engine"; return 0; }
  CK(cudaSetDevice(e->device));
  if (!e->t0) { CK(cudaEventCreate(&e->t0)); CK(cudaEventCreate(&e->t1)); }
  CK(cudaEventRecord(e->t0, e->stream));
  return 1;
}

int nrx_timer_stop(nrx_engine *e, double *elapsed_ms) {
  if (!e || !e->t0) { g_err = "timer not started"; return 0; }
  CK(cudaSetDevice(e->device));
  CK(cudaEventRecord(e->t1, e->stream));
  CK(cudaEventSynchronize(e->t1));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e->t0, e->t1));
  if (elapsed_ms) *elapsed_ms = ms;
  return 1;
}

int nrx_profile_enable(nrx_engine *e, int on) {
  if (!e) { g_err = "null engine"; return 0; }
  e->prof = on != 0;
  for (nrx_engine::ProfKind &k : e->profk) {
    for (auto &ev : k.events) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
    k = nrx_engine::ProfKind();
  }
  return 1;
}

int nrx_profile_read_kind(nrx_engine *e, int kind, double *ms, unsigned long long *launches, unsigned long long *units, unsigned long long *bytes,
                          unsigned long long *compulsory_bytes) {
  if (!e) { g_err = "null engine"; return 0; }
  if (kind < 0 || kind >= NRX_PROF_KINDS) { g_err = "nrx_profile_read_kind: bad kind"; return 0; }
  CK(cudaSetDevice(e->device));
  CK(cudaStreamSynchronize(e->stream));
  nrx_engine::ProfKind &k = e->profk[kind];
  for (auto &ev : k.events) {
    float t = 0;
    cudaEventElapsedTime(&t, ev.first, ev.second);
    k.ms += t;
    cudaEventDestroy(ev.first); cudaEventDestroy(ev.second);
  }
  k.events.clear();
  if (ms) *ms = k.ms;
  if (launches) *launches = k.launches;
  if (units) *units = k.units;
  if (bytes) *bytes = k.bytes;
  if (compulsory_bytes) *compulsory_bytes = k.cbytes;
  return 1;
}

int nrx_profile_read(nrx_engine *e, double *clv_ms, unsigned long long *clv_launches, unsigned long long *clv_site_updates, unsigned long long *clv_bytes) {
  return nrx_profile_read_kind(e, NRX_PROF_K2, clv_ms, clv_launches, clv_site_updates, clv_bytes, nullptr);
}

}  // extern "C"
