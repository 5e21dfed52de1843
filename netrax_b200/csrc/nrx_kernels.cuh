/*
 * nrx_kernels.cuh — hand-written sm_100a kernels of the NetRAX network-likelihood hot path.
 *
 * All kernels are FP64, HBM-streaming ("stencil over CLVs"): one thread owns one
 * (pattern, rate category) pair = `states_padded` consecutive doubles, so a warp reads/writes 1 KB of
 * contiguous CLV per 256-bit LDG/STG (LDG.E.256 on sm_100a) — fully coalesced, no shared-memory staging
 * of CLV data needed because there is no reuse.  Per-edge P-matrix rows live in registers (DNA) and tip
 * lookup tables in shared memory.  Scaler decisions use warp ballots over the rate-category group.
 *
 * DNA (4 states) arithmetic mirrors the ORDER of the reference's AVX 4x4 kernels exactly —
 * separate multiply and add (__dmul_rn/__dadd_rn, never contracted into FMA) and pairwise sums
 * (p0+p1)+(p2+p3) — so CLVs and scaler counts are bit-identical to libpll given identical P-matrices
 * (LIBPLL/core_partials_avx.c:402-565,1310-1500,255-395,992-1030).
 */
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/nrx_engine.h"

namespace nrx {

constexpr double SCALE_FACTOR = 115792089237316195423570985008687907853269984665640564039457584007913129639936.0;  // 2^256
constexpr double SCALE_THRESHOLD = 1.0 / SCALE_FACTOR;
constexpr int BLOCK = 256;

/* Device view of one partition (array of these lives in HBM; blockIdx.z selects). */
struct PartView {
  uint32_t states, sp, cats, patterns, tips, edges;
  uint32_t part_index, tip_pitch;  // tip_pitch: row pitch of tipchars (patterns rounded up to a tile multiple)
  uint32_t tip_codes, pad_;        // number of distinct tip codes in use (entries of tipmap)
  const double *pmat;        // [edges][cats][states][sp]
  const uint8_t *tipchars;   // [tips][patterns]
  const uint32_t *tipmap;    // [256] code -> state mask
  const uint32_t *weights;   // [patterns]
  const double *freqs, *eigenvecs, *inv_eigenvecs, *eigenvals, *rates, *rate_weights;
  double *const *clv;        // [nslots] -> CLV
  uint32_t *const *scaler;   // [nslots] -> scaler
  double *const *sumtable;   // [nsumtables]
  const double *diagp;       // [cats][states][4] for the current derivative call
  const double *tiplut;      // 20-state partitions: [edges][AA_LUT_CODES][cats*20] sums of P rows over each tip code's states (K1b)
  const double *summat;      // 20-state partitions: K5 operand matrices [2][20][20]: A_L[j][k] = pi_k Vinv[k][j], A_R[j][k] = V[j][k]
  double pinv;               // proportion of invariant sites (+I, pll_partition_t::prop_invar); 0 = none
  const int *invariant;      // [patterns] state index of an invariant pattern (pll_update_invariant_sites, LIBPLL/models.c:651-760), else -1
  double *pmat_pad;          // 4-state x 4-category partitions: the same P-matrices at the bank-conflict-free pitch of k_walk_dna4's shared-memory table, [edges][4][PCAT]
  const double *sumlut;      // 20-state partitions: K5 tip table [AA_LUT_CODES][cats*20]: sum_{k in code} pi_k Vinv[k][j], replicated per category
  /* Mixtures with one rate matrix per category (LG4M / LG4X: raxml-ng's ratecat_submodels = libpll's params_indices[c],
   * src/RaxmlWrapper.cpp:199-203): freqs / eigenvecs / inv_eigenvecs / eigenvals hold `nmodels` blocks back to back (strides sp,
   * states*sp, states*sp, sp) and category c uses block cat_model[c].  nmodels == 1: the single-matrix layout every fast kernel
   * assumes; partitions with nmodels > 1 run K1 and the generic K3-K6 kernels (K2 only reads P-matrices, which are per category
   * already). */
  uint32_t nmodels;
  uint8_t cat_model[16];
};

__device__ __forceinline__ uint32_t cat_model_of(const PartView &pv, uint32_t c) { return pv.nmodels > 1 ? pv.cat_model[c] : 0u; }

__device__ __forceinline__ double tree4(double a, double b, double c, double d) {
  return __dadd_rn(__dadd_rn(a, b), __dadd_rn(c, d));
}

struct D4 { double x, y, z, w; };

/* +I (proportion of invariant sites), the three places libpll accounts for it (formulas identical in the generic, AVX and AVX2
 * kernels): root lnL per category  w (t (1 - pinv) + pi_inv pinv)            LIBPLL/core_likelihood.c:176-188
 *           edge lnL               terma += w t (1 - pinv); terminv += w pi_inv pinv, the site scaling undone on terma only
 *                                                                             LIBPLL/core_likelihood.c:523-560
 *           derivatives per category  (c0 (1 - pinv) + pi_inv pinv, c1 (1 - pinv), c2 (1 - pinv))   core_derivatives.c:676-686
 * invf = frequency of the pattern's invariant state, 0 if the pattern is variable. */
__device__ __forceinline__ double root_cat_term(double term_r, double w, double pinv, double invf) {
  if (pinv > 0.0) return __dmul_rn(w, __dadd_rn(__dmul_rn(term_r, __dsub_rn(1.0, pinv)), __dmul_rn(invf, pinv)));
  return __dmul_rn(term_r, w);
}
__device__ __forceinline__ void edge_cat_accum(double terma_r, double w, double pinv, double invf, bool has_inv, double &terma, double &terminv) {
  if (pinv > 0.0) {
    terma = __dadd_rn(terma, __dmul_rn(__dmul_rn(w, terma_r), __dsub_rn(1.0, pinv)));
    if (has_inv) terminv = __dadd_rn(terminv, __dmul_rn(__dmul_rn(w, invf), pinv));
  } else {
    terma = __dadd_rn(terma, __dmul_rn(terma_r, w));
  }
}
__device__ __forceinline__ double edge_site_lnl(double terma, double terminv, uint32_t s, double log_thresh) {
  if (s) {
    if (terminv > 0.0) {   // undo the scaling of the non-invariant term only, capped at PLL_SCALE_RATE_MAXDIFF = 4 steps
      double f = SCALE_THRESHOLD;
      const uint32_t capped = s < 4u ? s : 4u;
      for (uint32_t i = 1; i < capped; ++i) f = __dmul_rn(f, SCALE_THRESHOLD);
      return log(__dadd_rn(__dmul_rn(terma, f), terminv));
    }
    return __dadd_rn(log(terma), __dmul_rn((double)s, log_thresh));
  }
  return log(__dadd_rn(terma, terminv));
}
__device__ __forceinline__ void deriv_cat_pinv(double &c0, double &c1, double &c2, double pinv, double invf) {
  if (pinv > 0.0) {
    const double q = __dsub_rn(1.0, pinv);
    c0 = __dadd_rn(__dmul_rn(c0, q), __dmul_rn(invf, pinv));
    c1 = __dmul_rn(c1, q);
    c2 = __dmul_rn(c2, q);
  }
}

__device__ __forceinline__ D4 ldg256(const double *p) {
  D4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void stg256(double *p, const D4 &v) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w) : "memory");
}

/* ------------------------------------------------------------------------------------------------
 * K1  P(t) = I + V^-1 diag(expm1(lambda r_c t)) V      (LIBPLL/core_pmatrix.c:24-244, core_pmatrix_avx.c:42)
 * grid.x = edges to update, block = 128 threads looping over the cats*states*states outputs.
 * ---------------------------------------------------------------------------------------------- */
__device__ __forceinline__ void pmatrix_body(const PartView &pv, double *pmat_out, uint32_t edge, double t, double *expd /* shared [cats][states] */) {
  const uint32_t S = pv.states, SP = pv.sp, C = pv.cats;
  double *out = pmat_out + (size_t)edge * C * S * SP;
  for (uint32_t i = threadIdx.x; i < C * S; i += blockDim.x) {
    const uint32_t c = i / S, m = i % S;
    // (eval*rate)*t exactly as core_pmatrix_avx.c:104-108, divided by (1 - pinv) for +I partitions (:113-117, core_pmatrix.c:196-200)
    double a = __dmul_rn(__dmul_rn(pv.eigenvals[cat_model_of(pv, c) * SP + m], pv.rates[c]), t);
    if (pv.pinv > 1e-8 /* PLL_MISC_EPSILON */) a = __ddiv_rn(a, __dsub_rn(1.0, pv.pinv));
    expd[i] = expm1(a);
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < C * S * SP; i += blockDim.x) {
    const uint32_t c = i / (S * SP), j = (i / SP) % S, k = i % SP;
    double v = 0.0;
    if (k < S) {
      if (t > 0.0) {
        const double *ex = expd + c * S;
        const size_t mo = (size_t)cat_model_of(pv, c) * S * SP;   // this category's rate matrix (params_indices[c], core_pmatrix.c:160-166)
        const double *iev = pv.inv_eigenvecs + mo, *ev = pv.eigenvecs + mo;
        if (S == 4) {  // tree sum, identity added last (core_pmatrix_avx.c:127-258)
          double p0 = __dmul_rn(__dmul_rn(iev[j * SP + 0], ex[0]), ev[0 * SP + k]);
          double p1 = __dmul_rn(__dmul_rn(iev[j * SP + 1], ex[1]), ev[1 * SP + k]);
          double p2 = __dmul_rn(__dmul_rn(iev[j * SP + 2], ex[2]), ev[2 * SP + k]);
          double p3 = __dmul_rn(__dmul_rn(iev[j * SP + 3], ex[3]), ev[3 * SP + k]);
          v = __dadd_rn(tree4(p0, p1, p2, p3), (j == k) ? 1.0 : 0.0);
        } else {       // identity first, serial accumulation (core_pmatrix.c:205-217)
          v = (j == k) ? 1.0 : 0.0;
          for (uint32_t m = 0; m < S; ++m)
            v = __dadd_rn(v, __dmul_rn(__dmul_rn(iev[j * SP + m], ex[m]), ev[m * SP + k]));
        }
      } else {
        v = (j == k) ? 1.0 : 0.0;  // zero branch length: identity (core_pmatrix.c:220-226)
      }
    }
    out[i] = v;
    if (pv.pmat_pad && S == 4 && C == 4) pv.pmat_pad[(size_t)edge * 72 + c * 18 + j * 4 + k] = v;   // 72 = 4 * PCAT (k_walk_dna4)
  }
}

__global__ void k_pmatrix(PartView pv, double *pmat_out, const uint32_t *edge_idx, const double *brlen) {
  extern __shared__ double expd[];  // [cats][states]
  pmatrix_body(pv, pmat_out, edge_idx[blockIdx.x], brlen[blockIdx.x], expd);
}

/* the same for several partitions of one shape class in ONE launch (unlinked branch lengths: BASELINE config 3 updates the 10
 * partitions' P-matrices before every evaluation — ten launches with two small uploads each were 2.6 % of its step):
 * grid = (most edges of a partition, partitions); partition y updates edge_idx / brlen [offset[y], offset[y + 1]) */
__global__ void k_pmatrix_multi(const PartView *__restrict__ views, const uint32_t *__restrict__ edge_idx, const double *__restrict__ brlen,
                                const uint32_t *__restrict__ offset) {
  extern __shared__ double expd[];
  const uint32_t o0 = offset[blockIdx.y], n = offset[blockIdx.y + 1] - o0;
  if (blockIdx.x >= n) return;
  const PartView &pv = views[blockIdx.y];
  pmatrix_body(pv, const_cast<double *>(pv.pmat), edge_idx[o0 + blockIdx.x], brlen[o0 + blockIdx.x], expd);
}

/* Tip-code upload check, on the device (round 2: nrx_set_tipchars_u8 used to trust its input and left the invariant-site table
 * stale): every code must map to a non-empty state set (pll_set_tip_states rejects unknown characters, LIBPLL/pll.c:875-957) —
 * otherwise *err is raised (mapped host memory, read by the host at its next synchronisation) — and the invariant-site table
 * is rebuilt as pll_update_invariant_sites does (LIBPLL/models.c:651-760): AND of all tips' state sets per pattern; exactly one
 * state left -> its index, else -1.  thread = pattern: tip rows are read coalesced. */
__global__ void __launch_bounds__(BLOCK) k_check_tipchars(const uint8_t *__restrict__ tipchars, uint32_t tips, uint32_t patterns, uint32_t pitch,
                                                           const uint32_t *__restrict__ tipmap, uint32_t full_mask, int *__restrict__ invariant,
                                                           volatile int *err) {
  __shared__ uint32_t smap[256];
  for (int i = threadIdx.x; i < 256; i += BLOCK) smap[i] = tipmap[i];
  __syncthreads();
  for (uint32_t n = blockIdx.x * BLOCK + threadIdx.x; n < patterns; n += gridDim.x * BLOCK) {
    uint32_t st = full_mask;
    bool bad = false;
    for (uint32_t t = 0; t < tips; ++t) {
      const uint32_t m = smap[tipchars[(size_t)t * pitch + n]];
      bad |= (m == 0u);
      st &= m;
    }
    if (bad) *err = 1;
    invariant[n] = (st == 0u || __popc(st) > 1) ? -1 : (__ffs(st) - 1);
  }
}

/* ------------------------------------------------------------------------------------------------
 * K2  CLV update, DNA (4 states) x 4 rate categories fast path.
 * grid = (pattern tiles, ops, partitions of this shape); thread = one (pattern, category).
 * ---------------------------------------------------------------------------------------------- */
__device__ __forceinline__ void build_tip_lut4(double *lut /*[16][4][4]*/, const double *pm /*[4][4][4] of the edge*/, int tid) {
  for (int idx = tid; idx < 256; idx += BLOCK) {
    const int mask = idx >> 4, c = (idx >> 2) & 3, i = idx & 3;
    const double *row = pm + (c * 4 + i) * 4;
    // masked load + tree sum (core_partials_avx.c:1372-1400)
    lut[idx] = tree4((mask & 1) ? row[0] : 0.0, (mask & 2) ? row[1] : 0.0, (mask & 4) ? row[2] : 0.0, (mask & 8) ? row[3] : 0.0);
  }
}

/* P-matrix rows come from shared memory: [cat][row][col] with the category stride padded to 18 doubles so
 * the four categories of a warp hit disjoint banks (each LDS.128 is then a single conflict-free wavefront
 * with 8-lane broadcast).  Keeping the 2x16 doubles per thread in registers instead costs 64 registers and
 * halves occupancy — measured worse for an HBM-bound kernel. */
constexpr int PCAT = 18;  // padded doubles per category block (16 + 2)

__device__ __forceinline__ D4 matvec4(const double *__restrict__ P /* smem, this thread's category */, const D4 &v) {
  D4 r;
  const double2 *q = reinterpret_cast<const double2 *>(P);
  double2 a, b;
  a = q[0]; b = q[1]; r.x = tree4(__dmul_rn(a.x, v.x), __dmul_rn(a.y, v.y), __dmul_rn(b.x, v.z), __dmul_rn(b.y, v.w));
  a = q[2]; b = q[3]; r.y = tree4(__dmul_rn(a.x, v.x), __dmul_rn(a.y, v.y), __dmul_rn(b.x, v.z), __dmul_rn(b.y, v.w));
  a = q[4]; b = q[5]; r.z = tree4(__dmul_rn(a.x, v.x), __dmul_rn(a.y, v.y), __dmul_rn(b.x, v.z), __dmul_rn(b.y, v.w));
  a = q[6]; b = q[7]; r.w = tree4(__dmul_rn(a.x, v.x), __dmul_rn(a.y, v.y), __dmul_rn(b.x, v.z), __dmul_rn(b.y, v.w));
  return r;
}

/* ------------------------------------------------------------------------------------------------
 * K2, DNA 4x4, bulk-async pipelined version (round-1b production kernel; kept as the A/B baseline of k_clv_dna4_pipe2,
 * env NRX_K2=1.  Its predecessor, a register-streaming kernel — loads held in registers, 98 registers -> 16 warps/SM,
 * 0.68 of the copy peak, profiles/r1a_* — was removed).
 *
 * HBM-bound streaming needs many bytes in flight per SM; holding them in registers caps occupancy.  Here a
 * block owns one op and walks pattern tiles (TP patterns = 8 KB per CLV operand); one elected thread keeps
 * NSTAGE tiles of BOTH operands (+ their scalers / tip codes) in flight with cp.async.bulk (the TMA engine's
 * 1-D bulk copy: SASS UBLKCP) into a shared-memory ring, completion signalled on one mbarrier per stage.
 * All 256 threads consume: thread = (pattern, category), 32 B of each operand from the ring, P-matrix rows of
 * its category in REGISTERS (loaded once per block: the op is fixed), result written straight to HBM as one
 * 256-bit store (a warp writes 1 KB contiguous).  Blocks are numbered op-fastest so that the ops of a node
 * that share a child CLV touch the same tile at about the same time and the re-read hits the 126 MB L2.
 * Arithmetic order is identical to k_clv_dna4 (and to the reference's AVX kernel).
 * ---------------------------------------------------------------------------------------------- */
constexpr int TP = 64;        // patterns per tile (one (pattern, cat) item per thread)
constexpr int NSTAGE = 6;     // tiles in flight per block: 6 x 17 KB = 102 KB -> 2 blocks per SM

struct __align__(128) ClvStage {
  double l[TP * 16];
  double r[TP * 16];
  uint32_t scl[TP];
  uint32_t scr[TP];
  uint8_t tl[TP];
  uint8_t tr[TP];
};
struct __align__(128) ClvPipeSmem {
  ClvStage st[NSTAGE];
  double lutL[256];
  double lutR[256];
  unsigned long long full[NSTAGE];
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ D4 matvec4_reg(const double (&P)[16], const D4 &v) {
  D4 r;
  r.x = tree4(__dmul_rn(P[0], v.x), __dmul_rn(P[1], v.y), __dmul_rn(P[2], v.z), __dmul_rn(P[3], v.w));
  r.y = tree4(__dmul_rn(P[4], v.x), __dmul_rn(P[5], v.y), __dmul_rn(P[6], v.z), __dmul_rn(P[7], v.w));
  r.z = tree4(__dmul_rn(P[8], v.x), __dmul_rn(P[9], v.y), __dmul_rn(P[10], v.z), __dmul_rn(P[11], v.w));
  r.w = tree4(__dmul_rn(P[12], v.x), __dmul_rn(P[13], v.y), __dmul_rn(P[14], v.z), __dmul_rn(P[15], v.w));
  return r;
}

/* grid = (nops * groups, 1, partitions of this shape); block b: op = b % nops, tile group = b / nops */
/* Fused K3 (ops with lnl_item != 0, i.e. root displayed trees of a full traversal): the epilogue also computes the
 * tree's per-site likelihood term from the values it is about to store — same arithmetic and order as
 * k_tree_lnl_dna4 — and writes it to persite[(item * nparts_total + part) * persite_stride + site]; k_term_lnl_sum
 * then does the log / scaler / weight part densely from 16 B per site instead of K3 re-reading the 128 B CLV.
 * (Doing the log here as well was measured: it runs on 1 lane in 4 and lengthens every tile's critical path.) */
__global__ void __launch_bounds__(BLOCK, 2) k_clv_dna4_pipe(const PartView *__restrict__ parts, const nrx_op *__restrict__ ops,
                                                             uint32_t nops, uint32_t groups, double *__restrict__ persite,
                                                             size_t persite_stride, uint32_t nparts_total) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  ClvPipeSmem &sm = *reinterpret_cast<ClvPipeSmem *>(smem_raw);
  const PartView &pv = parts[blockIdx.z];
  const nrx_op op = ops[blockIdx.x % nops];
  const uint32_t grp = blockIdx.x / nops;
  const int tid = threadIdx.x, cat = tid & 3, lane = tid & 31;
  const uint32_t ntiles = (pv.patterns + TP - 1) / TP;
  if (grp >= ntiles) return;
  const uint32_t count = (ntiles - grp + groups - 1) / groups;  // my tiles: grp, grp + groups, ...
  const int lk = op.left_kind, rk = op.right_kind;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < NSTAGE; ++s) mbar_init(&sm.full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (lk == NRX_TIP) build_tip_lut4(sm.lutL, pv.pmat + (size_t)op.left_edge * 64, tid);
  if (rk == NRX_TIP) build_tip_lut4(sm.lutR, pv.pmat + (size_t)op.right_edge * 64, tid);
  double PL[16], PR[16];
  if (lk == NRX_CLV) {
    const double *src = pv.pmat + (size_t)op.left_edge * 64 + cat * 16;
#pragma unroll
    for (int i = 0; i < 16; ++i) PL[i] = src[i];
  }
  if (rk == NRX_CLV) {
    const double *src = pv.pmat + (size_t)op.right_edge * 64 + cat * 16;
#pragma unroll
    for (int i = 0; i < 16; ++i) PR[i] = src[i];
  }
  __syncthreads();

  const double *clvL = (lk == NRX_CLV) ? pv.clv[op.left_idx] : nullptr;
  const double *clvR = (rk == NRX_CLV) ? pv.clv[op.right_idx] : nullptr;
  const uint32_t *scL = (lk == NRX_CLV) ? pv.scaler[op.left_idx] : nullptr;
  const uint32_t *scR = (rk == NRX_CLV) ? pv.scaler[op.right_idx] : nullptr;
  const uint8_t *tipL = (lk == NRX_TIP) ? pv.tipchars + (size_t)op.left_idx * pv.tip_pitch : nullptr;
  const uint8_t *tipR = (rk == NRX_TIP) ? pv.tipchars + (size_t)op.right_idx * pv.tip_pitch : nullptr;
  double *par = pv.clv[op.parent_slot];
  uint32_t *psc = pv.scaler[op.parent_slot];
  const bool tiptip = (lk == NRX_TIP && rk == NRX_TIP);
  const unsigned quad = 0xFu << (lane & ~3);
  const bool emit = op.lnl_item != 0 && persite != nullptr;
  const uint32_t tx_bytes = ((lk == NRX_CLV) ? TP * 128u + TP * 4u : (lk == NRX_TIP ? (uint32_t)TP : 0u)) +
                            ((rk == NRX_CLV) ? TP * 128u + TP * 4u : (rk == NRX_TIP ? (uint32_t)TP : 0u));
  double f0 = 0, f1 = 0, f2 = 0, f3 = 0, wcat = 0;
  double *ps_out = nullptr;
  if (emit) {
    f0 = pv.freqs[0]; f1 = pv.freqs[1]; f2 = pv.freqs[2]; f3 = pv.freqs[3]; wcat = pv.rate_weights[cat];
    ps_out = persite + ((size_t)(op.lnl_item - 1) * nparts_total + pv.part_index) * persite_stride;
  }

  // every buffer is padded to a whole number of tiles by the engine, so full-tile copies never run off the end
  auto issue = [&](uint32_t k) {
    ClvStage &st = sm.st[k % NSTAGE];
    unsigned long long *bar = &sm.full[k % NSTAGE];
    const size_t p0 = (size_t)(grp + (size_t)k * groups) * TP;
    mbar_expect_tx(bar, tx_bytes);
    if (lk == NRX_CLV) { bulk_g2s(st.l, clvL + p0 * 16, TP * 128u, bar); bulk_g2s(st.scl, scL + p0, TP * 4u, bar); }
    else if (lk == NRX_TIP) bulk_g2s(st.tl, tipL + p0, TP, bar);
    if (rk == NRX_CLV) { bulk_g2s(st.r, clvR + p0 * 16, TP * 128u, bar); bulk_g2s(st.scr, scR + p0, TP * 4u, bar); }
    else if (rk == NRX_TIP) bulk_g2s(st.tr, tipR + p0, TP, bar);
  };
  if (tid == 0) {
    const uint32_t pre = count < (uint32_t)NSTAGE ? count : (uint32_t)NSTAGE;
    for (uint32_t k = 0; k < pre; ++k) issue(k);
  }

  const int pl = tid >> 2;  // pattern within the tile
  for (uint32_t k = 0; k < count; ++k) {
    const ClvStage &st = sm.st[k % NSTAGE];
    mbar_wait(&sm.full[k % NSTAGE], (k / NSTAGE) & 1u);
    const size_t site = (size_t)(grp + (size_t)k * groups) * TP + pl;
    const bool act = site < pv.patterns;
    D4 x, y, p;
    if (lk == NRX_CLV) x = matvec4_reg(PL, *reinterpret_cast<const D4 *>(st.l + tid * 4));
    else if (lk == NRX_TIP) x = *reinterpret_cast<const D4 *>(sm.lutL + ((st.tl[pl] & 15) * 4 + cat) * 4);
    if (rk == NRX_CLV) y = matvec4_reg(PR, *reinterpret_cast<const D4 *>(st.r + tid * 4));
    else if (rk == NRX_TIP) y = *reinterpret_cast<const D4 *>(sm.lutR + ((st.tr[pl] & 15) * 4 + cat) * 4);
    if (rk == NRX_NONE) p = x;
    else if (lk == NRX_NONE) p = y;
    else { p.x = __dmul_rn(x.x, y.x); p.y = __dmul_rn(x.y, y.y); p.z = __dmul_rn(x.z, y.z); p.w = __dmul_rn(x.w, y.w); }
    const bool small = act & (p.x < SCALE_THRESHOLD) & (p.y < SCALE_THRESHOLD) & (p.z < SCALE_THRESHOLD) & (p.w < SCALE_THRESHOLD);
    const unsigned b = __ballot_sync(0xffffffffu, small);
    const bool scale = !tiptip && ((b & quad) == quad);
    uint32_t s = 0;
    if (!tiptip) {
      if (lk == NRX_CLV) s += st.scl[pl];
      if (rk == NRX_CLV) s += st.scr[pl];
      s += scale ? 1u : 0u;
    }
    if (act) {
      if (scale) { p.x = __dmul_rn(p.x, SCALE_FACTOR); p.y = __dmul_rn(p.y, SCALE_FACTOR); p.z = __dmul_rn(p.z, SCALE_FACTOR); p.w = __dmul_rn(p.w, SCALE_FACTOR); }
      stg256(par + (site * 4 + cat) * 4, p);
      if (cat == 0) psc[site] = s;
    }
    if (emit) {  // warp-uniform (the op is fixed per block): per-site likelihood sum_c w_c sum_i pi_i clv[c][i], before the log
      double t = 0.0;
      if (act) t = __dmul_rn(tree4(__dmul_rn(f0, p.x), __dmul_rn(f1, p.y), __dmul_rn(f2, p.z), __dmul_rn(f3, p.w)), wcat);
      const double t1 = __shfl_down_sync(0xffffffffu, t, 1), t2 = __shfl_down_sync(0xffffffffu, t, 2), t3 = __shfl_down_sync(0xffffffffu, t, 3);
      if (act && cat == 0) ps_out[site] = __dadd_rn(__dadd_rn(__dadd_rn(t, t1), t2), t3);
    }
    // Block-wide barrier, then refill.  (Measured: releasing stages per warp through "empty" mbarriers instead, so that
    // warps run ahead independently, is 18 % SLOWER on config 5 — the lock-step keeps a block's 8 KB of stores and the
    // co-scheduled ops' reads of a shared child tile together in time.)
    __syncthreads();
    if (tid == 0 && k + NSTAGE < count) issue(k + NSTAGE);
  }
}

/* ------------------------------------------------------------------------------------------------
 * K2, DNA 4x4, bulk-async pipeline, SPECIALISED loop bodies (production; k_clv_dna4_pipe above is kept for A/B, env
 * NRX_K2=1).  Same pipeline, same arithmetic; what changes is the instruction stream: ncu (profiles/r1e_k2_*) showed
 * the kernel neither DRAM- nor FP64-bound but issuing ~135 instructions per (pattern, category) of which only 64 are
 * FP64 — block-uniform branches on the operand kinds inside the loop (and the register moves that merge them), and
 * 64-bit address arithmetic redone per tile.  Here the operand kinds of the block's op select one of eight loop
 * instantiations ONCE, outside the loop; output / scaler / per-site pointers and the ring stage advance incrementally.
 * ---------------------------------------------------------------------------------------------- */
struct PipeCtx {
  const double *clvL, *clvR;
  const uint32_t *scL, *scR;
  const uint8_t *tipL, *tipR;
  double *par;
  uint32_t *psc;
  double *ps_out;   // fused K3 output row of this op (or nullptr)
  uint32_t grp, groups, count, patterns;
  bool skip_clv;    // score-only evaluation: the CLV of a root displayed tree (EMIT) is not stored, only its per-site term and scaler
};

/* NT = 64-pattern sub-tiles per ring stage (1: 6 stages x 64 patterns, the geometry of k_clv_dna4_pipe; 2: 3 stages x 128
 * patterns — the same bytes in flight, half as many mbarrier waits / block barriers / refills per byte, two independent
 * items per thread between barriers). */
/* CATS = rate categories (1, 2, 4, 8 or 16; round 2: every power-of-two category count runs this kernel — a sub-tile is always 256
 * (pattern, category) items = 8 KB per CLV operand, i.e. 256 / CATS patterns; the reference's AVX kernels loop over rate_cats at
 * full SIMD speed too, LIBPLL/core_partials_avx.c:402-565). */
template <int NT, int CATS>
struct __align__(128) PipeStage {
  static constexpr int TPC = BLOCK / CATS;   // patterns per sub-tile
  double l[NT * BLOCK * 4];
  double r[NT * BLOCK * 4];
  uint32_t scl[NT * TPC];
  uint32_t scr[NT * TPC];
  uint8_t tl[NT * TPC < 16 ? 16 : NT * TPC];
  uint8_t tr[NT * TPC < 16 ? 16 : NT * TPC];
};
template <int NT, int CATS>
struct __align__(128) PipeSmem {
  static constexpr int NST = NSTAGE / NT;
  PipeStage<NT, CATS> st[NST];
  double lutL[16 * CATS * 4];
  double lutR[16 * CATS * 4];
  unsigned long long full[NST];
};

template <int CATS>
__device__ __forceinline__ void build_tip_lut(double *lut /*[16][CATS][4]*/, const double *pm /*[CATS][4][4] of the edge*/, int tid) {
  for (int idx = tid; idx < 16 * CATS * 4; idx += BLOCK) {
    const int mask = idx / (CATS * 4), c = (idx >> 2) % CATS, i = idx & 3;
    const double *row = pm + (c * 4 + i) * 4;
    lut[idx] = tree4((mask & 1) ? row[0] : 0.0, (mask & 2) ? row[1] : 0.0, (mask & 4) ? row[2] : 0.0, (mask & 8) ? row[3] : 0.0);
  }
}

template <int LK, int RK, int NT, int CATS>
__device__ __forceinline__ void pipe_issue(PipeSmem<NT, CATS> &sm, const PipeCtx &c, uint32_t k, uint32_t stage) {
  constexpr uint32_t TPX = NT * (BLOCK / CATS);   // patterns per stage
  constexpr uint32_t CB = CATS * 32u;             // CLV bytes per pattern
  constexpr uint32_t TIPB = TPX < 16u ? 16u : TPX;   // bulk copies move multiples of 16 bytes (rows are padded)
  PipeStage<NT, CATS> &st = sm.st[stage];
  unsigned long long *bar = &sm.full[stage];
  const size_t p0 = ((size_t)c.grp + (size_t)k * c.groups) * TPX;
  constexpr uint32_t tx = ((LK == NRX_CLV) ? TPX * CB + TPX * 4u : (LK == NRX_TIP ? TIPB : 0u)) +
                          ((RK == NRX_CLV) ? TPX * CB + TPX * 4u : (RK == NRX_TIP ? TIPB : 0u));
  mbar_expect_tx(bar, tx);
  if (LK == NRX_CLV) { bulk_g2s(st.l, c.clvL + p0 * (CATS * 4), TPX * CB, bar); bulk_g2s(st.scl, c.scL + p0, TPX * 4u, bar); }
  else if (LK == NRX_TIP) bulk_g2s(st.tl, c.tipL + p0, TIPB, bar);
  if (RK == NRX_CLV) { bulk_g2s(st.r, c.clvR + p0 * (CATS * 4), TPX * CB, bar); bulk_g2s(st.scr, c.scR + p0, TPX * 4u, bar); }
  else if (RK == NRX_TIP) bulk_g2s(st.tr, c.tipR + p0, TIPB, bar);
}

/* first ring fill, by one thread, BEFORE the block loads its P-matrices / builds its tip tables: the copies fly while the
 * prologue runs (small alignments are latency-bound: ~1 us per launch) */
template <int LK, int RK, int NT, int CATS>
__device__ __forceinline__ void pipe_prefetch(PipeSmem<NT, CATS> &sm, const PipeCtx &c) {
  constexpr uint32_t NST = PipeSmem<NT, CATS>::NST;
  const uint32_t pre = c.count < NST ? c.count : NST;
  for (uint32_t k = 0; k < pre; ++k) pipe_issue<LK, RK, NT, CATS>(sm, c, k, k);
}

template <int LK, int RK, bool EMIT, int NT, int CATS>
__device__ __forceinline__ void pipe_loop(PipeSmem<NT, CATS> &sm, const PipeCtx &c, const double (&PL)[16], const double (&PR)[16],
                                          double f0, double f1, double f2, double f3, double wcat) {
  constexpr uint32_t TPC = BLOCK / CATS;   // patterns per sub-tile
  constexpr uint32_t TPX = NT * TPC;
  constexpr uint32_t NST = PipeSmem<NT, CATS>::NST;
  const int tid = threadIdx.x, cat = tid & (CATS - 1), lane = tid & 31, pl = tid / CATS;
  constexpr bool tiptip = (LK == NRX_TIP && RK == NRX_TIP);
  // the lanes holding the categories of this thread's pattern (CATS <= 16: a group never straddles a warp)
  const unsigned quad = (CATS >= 32 ? 0xffffffffu : ((1u << CATS) - 1u)) << (lane & ~(CATS - 1));
  uint32_t site0 = c.grp * TPX + pl;                       // patterns < 2^32 (the first NST stages were issued by pipe_prefetch)
  const uint32_t site_step = c.groups * TPX;
  double *out0 = c.par + ((size_t)site0 * CATS + cat) * 4;
  const size_t out_step = (size_t)site_step * (CATS * 4);
  uint32_t stage = 0, phase = 0;
  for (uint32_t k = 0; k < c.count; ++k) {
    const PipeStage<NT, CATS> &st = sm.st[stage];
    mbar_wait(&sm.full[stage], phase);
#pragma unroll
    for (int u = 0; u < NT; ++u) {
      const uint32_t site = site0 + u * TPC;
      const int plu = pl + u * TPC, tu = tid + u * BLOCK;
      double *out = out0 + (size_t)u * BLOCK * 4;
      const bool act = site < c.patterns;
      D4 x, y, p;
      if (LK == NRX_CLV) x = matvec4_reg(PL, *reinterpret_cast<const D4 *>(st.l + tu * 4));
      else if (LK == NRX_TIP) x = *reinterpret_cast<const D4 *>(sm.lutL + ((st.tl[plu] & 15) * CATS + cat) * 4);
      if (RK == NRX_CLV) y = matvec4_reg(PR, *reinterpret_cast<const D4 *>(st.r + tu * 4));
      else if (RK == NRX_TIP) y = *reinterpret_cast<const D4 *>(sm.lutR + ((st.tr[plu] & 15) * CATS + cat) * 4);
      if (RK == NRX_NONE) p = x;
      else if (LK == NRX_NONE) p = y;
      else { p.x = __dmul_rn(x.x, y.x); p.y = __dmul_rn(x.y, y.y); p.z = __dmul_rn(x.z, y.z); p.w = __dmul_rn(x.w, y.w); }
      uint32_t s = 0;
      if (!tiptip) {
        const bool small = act & (p.x < SCALE_THRESHOLD) & (p.y < SCALE_THRESHOLD) & (p.z < SCALE_THRESHOLD) & (p.w < SCALE_THRESHOLD);
        const unsigned b = __ballot_sync(0xffffffffu, small);
        const bool scale = (b & quad) == quad;
        if (LK == NRX_CLV) s += st.scl[plu];
        if (RK == NRX_CLV) s += st.scr[plu];
        s += scale ? 1u : 0u;
        if (scale) { p.x = __dmul_rn(p.x, SCALE_FACTOR); p.y = __dmul_rn(p.y, SCALE_FACTOR); p.z = __dmul_rn(p.z, SCALE_FACTOR); p.w = __dmul_rn(p.w, SCALE_FACTOR); }
      }
      if (act) {
        if (!(EMIT && c.skip_clv)) stg256(out, p);
        if (cat == 0) c.psc[site] = s;
      }
      if (EMIT) {
        double t = 0.0;
        if (act) t = __dmul_rn(tree4(__dmul_rn(f0, p.x), __dmul_rn(f1, p.y), __dmul_rn(f2, p.z), __dmul_rn(f3, p.w)), wcat);
        double acc = t;   // categories in order: ((t0 + t1) + t2) + ...
#pragma unroll
        for (int i = 1; i < CATS; ++i) acc = __dadd_rn(acc, __shfl_down_sync(0xffffffffu, t, i));
        if (act && cat == 0) c.ps_out[site] = acc;
      }
    }
    __syncthreads();   // lock-step refill (see k_clv_dna4_pipe: per-warp release measured 18 % slower)
    if (tid == 0 && k + NST < c.count) pipe_issue<LK, RK, NT, CATS>(sm, c, k + NST, stage);
    site0 += site_step;
    out0 += out_step;
    if (++stage == NST) { stage = 0; phase ^= 1u; }
  }
}

/* pdl != 0: the launch carries the programmatic-stream-serialization attribute (plan graphs): the block may start while the
 * previous K2 launch still runs, does everything that does not depend on it (mbarrier init, P-matrices, tip tables),
 * lets ITS successor start (griddepcontrol.launch_dependents) and only then waits for the predecessor's CLVs
 * (griddepcontrol.wait) before the first ring fill.  Small alignments: ~2 us of prologue per launch off the critical path. */
template <int NT, int CATS = 4>
__global__ void __launch_bounds__(BLOCK, 2) k_clv_dna4_pipe2(const PartView *__restrict__ parts, const nrx_op *__restrict__ ops,
                                                              uint32_t nops, uint32_t groups, double *__restrict__ persite,
                                                              size_t persite_stride, uint32_t nparts_total, int flags /* bit 0: PDL launch, bit 1: score-only (root CLVs not stored) */) {
  const int pdl = flags & 1;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  PipeSmem<NT, CATS> &sm = *reinterpret_cast<PipeSmem<NT, CATS> *>(smem_raw);
  const PartView &pv = parts[blockIdx.z];
  const nrx_op op = ops[blockIdx.x % nops];
  const uint32_t grp = blockIdx.x / nops;
  const int tid = threadIdx.x, cat = tid & (CATS - 1);
  constexpr uint32_t TPX = NT * (BLOCK / CATS);
  const uint32_t ntiles = (pv.patterns + TPX - 1) / TPX;
  if (grp >= ntiles) return;
  const int lk = op.left_kind, rk = op.right_kind;
  PipeCtx c;
  c.clvL = (lk == NRX_CLV) ? pv.clv[op.left_idx] : nullptr;
  c.clvR = (rk == NRX_CLV) ? pv.clv[op.right_idx] : nullptr;
  c.scL = (lk == NRX_CLV) ? pv.scaler[op.left_idx] : nullptr;
  c.scR = (rk == NRX_CLV) ? pv.scaler[op.right_idx] : nullptr;
  c.tipL = (lk == NRX_TIP) ? pv.tipchars + (size_t)op.left_idx * pv.tip_pitch : nullptr;
  c.tipR = (rk == NRX_TIP) ? pv.tipchars + (size_t)op.right_idx * pv.tip_pitch : nullptr;
  c.par = pv.clv[op.parent_slot];
  c.psc = pv.scaler[op.parent_slot];
  c.grp = grp; c.groups = groups; c.patterns = pv.patterns;
  c.count = (ntiles - grp + groups - 1) / groups;
  c.ps_out = nullptr;
  c.skip_clv = false;
#define NRX_PIPE_PRE(L, R) case (L) * 3 + (R): pipe_prefetch<L, R, NT, CATS>(sm, c); break;
#define NRX_PIPE_PREFETCH()                                                                                     \
  switch (lk * 3 + rk) {                                                                                        \
    NRX_PIPE_PRE(NRX_CLV, NRX_CLV) NRX_PIPE_PRE(NRX_CLV, NRX_TIP) NRX_PIPE_PRE(NRX_CLV, NRX_NONE)               \
    NRX_PIPE_PRE(NRX_TIP, NRX_CLV) NRX_PIPE_PRE(NRX_TIP, NRX_TIP) NRX_PIPE_PRE(NRX_TIP, NRX_NONE)               \
    NRX_PIPE_PRE(NRX_NONE, NRX_CLV) NRX_PIPE_PRE(NRX_NONE, NRX_TIP)                                             \
    default: break;                                                                                             \
  }
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < PipeSmem<NT, CATS>::NST; ++s) mbar_init(&sm.full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (!pdl) { NRX_PIPE_PREFETCH() }   // stream-ordered launch: the children are final, fill the ring before the prologue
  }
  if (lk == NRX_TIP) build_tip_lut<CATS>(sm.lutL, pv.pmat + (size_t)op.left_edge * (CATS * 16), tid);
  if (rk == NRX_TIP) build_tip_lut<CATS>(sm.lutR, pv.pmat + (size_t)op.right_edge * (CATS * 16), tid);
  double PL[16], PR[16];
  if (lk == NRX_CLV) {
    const double *src = pv.pmat + (size_t)op.left_edge * (CATS * 16) + cat * 16;
#pragma unroll
    for (int i = 0; i < 16; ++i) PL[i] = src[i];
  }
  if (rk == NRX_CLV) {
    const double *src = pv.pmat + (size_t)op.right_edge * (CATS * 16) + cat * 16;
#pragma unroll
    for (int i = 0; i < 16; ++i) PR[i] = src[i];
  }
  if (pdl) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");   // every thread: the predecessor's CLVs / scalers are complete and visible
    if (tid == 0) { NRX_PIPE_PREFETCH() }
  }
#undef NRX_PIPE_PREFETCH
#undef NRX_PIPE_PRE
  __syncthreads();
  const bool emit = op.lnl_item != 0 && persite != nullptr;
  double f0 = 0, f1 = 0, f2 = 0, f3 = 0, wcat = 0;
  if (emit) {
    f0 = pv.freqs[0]; f1 = pv.freqs[1]; f2 = pv.freqs[2]; f3 = pv.freqs[3]; wcat = pv.rate_weights[cat];
    c.ps_out = persite + ((size_t)(op.lnl_item - 1) * nparts_total + pv.part_index) * persite_stride;
    c.skip_clv = (flags & 2) != 0;
  }
#define NRX_PIPE_CASE(L, R)                                                                  \
  case (L) * 3 + (R):                                                                        \
    if (emit) pipe_loop<L, R, true, NT, CATS>(sm, c, PL, PR, f0, f1, f2, f3, wcat);          \
    else pipe_loop<L, R, false, NT, CATS>(sm, c, PL, PR, f0, f1, f2, f3, wcat);              \
    break;
  switch (lk * 3 + rk) {
    NRX_PIPE_CASE(NRX_CLV, NRX_CLV)
    NRX_PIPE_CASE(NRX_CLV, NRX_TIP)
    NRX_PIPE_CASE(NRX_CLV, NRX_NONE)
    NRX_PIPE_CASE(NRX_TIP, NRX_CLV)
    NRX_PIPE_CASE(NRX_TIP, NRX_TIP)
    NRX_PIPE_CASE(NRX_TIP, NRX_NONE)
    NRX_PIPE_CASE(NRX_NONE, NRX_CLV)
    NRX_PIPE_CASE(NRX_NONE, NRX_TIP)
    default: break;
  }
#undef NRX_PIPE_CASE
}

/* ------------------------------------------------------------------------------------------------
 * K2, DNA 4x4, NODE-CENTRIC variant (round 2): for the ops of a network node that share children.
 *
 * k_clv_dna4_pipe2 gives every op (= displayed tree) its own block, so an op re-reads both of its children and recomputes
 * P_edge . child although the 96 displayed trees of a big node are the 12 x 8 compatible pairs of only 20 children; the re-reads are
 * served by the L2 most of the time, but at 1 M patterns 13 % more than the compulsory bytes cross the HBM pins (profiles/
 * r2b_k2_traffic_1M.md: 22.5 GB per step, 15.5 of them in the root node's launch).  The compatible (left tree, right tree) pairs of a
 * node form complete bipartite blocks (trees that agree on the shared reticulations); the host cuts the ops of a plan batch into such
 * GROUPS of <= NODE_MAXC distinct children (nrx_plan_create).  Here block = (group, tile of NODE_TP patterns):
 *   1. every distinct child tile (+ scalers) is bulk-copied into shared memory ONCE;
 *   2. x_i = P_left . left_i and y_j = P_right . right_j are computed ONCE per distinct child, in place (thread = (pattern, category)
 *      only ever touches its own 32 bytes of a tile, so no barrier is needed between the steps);
 *   3. every op (i, j) is a product x_i * y_j, the per-pattern scaling vote, and a streamed 256-bit store — ~30 instructions per item
 *      instead of ~95, HBM traffic = compulsory.
 * Arithmetic is the same instruction sequence per value as the pipelined kernel (matvec4_reg, separate multiply / add), so CLVs and
 * scalers stay bit-identical.  Three blocks per SM interleave one block's load phase with the others' store phases.
 * ---------------------------------------------------------------------------------------------- */
constexpr int NODE_TP = 32;          // patterns per tile: a child tile is 4 KB
constexpr int NODE_THREADS = NODE_TP * 4;
constexpr int NODE_MAXC = 16;        // distinct children per group (left + right)
constexpr int NODE_MAXOPS = 128;     // ops per group

struct nrx_node_op { uint16_t li, rj; uint32_t parent_slot; uint32_t lnl_item; uint32_t pad_; };   // 16 bytes; li / rj index the group's left / right children
struct nrx_node_group {             // 16-byte multiple: bulk-copied into shared memory together with its ops
  uint32_t nl, nr, nops, left_edge, right_edge, op_first, pad0_, pad1_;
  uint32_t child_slot[NODE_MAXC];   // nl left slots, then nr right slots
};
struct nrx_node_block { uint32_t group, tile0, stride, pad_; };   // block b works on tiles tile0, tile0 + stride, ... of its group

struct __align__(128) NodeSmemFixed {   // followed by double tile[ncmax][NODE_TP * 16] and uint32_t sc[ncmax][NODE_TP] (ncmax = the launch's largest group)
  nrx_node_group grp;
  nrx_node_op ops[NODE_MAXOPS];
  double *par[NODE_MAXOPS];
  uint32_t *psc[NODE_MAXOPS];
  const double *cclv[NODE_MAXC];
  const uint32_t *csc[NODE_MAXC];
  unsigned long long bar, bar_desc;
};
__host__ __device__ constexpr size_t node_smem_bytes(uint32_t ncmax) { return sizeof(NodeSmemFixed) + (size_t)ncmax * (NODE_TP * 128 + NODE_TP * 4); }

__global__ void __launch_bounds__(NODE_THREADS, 4) k_clv_node_dna4(const PartView *__restrict__ parts, const nrx_node_group *__restrict__ groups,
                                                                    const nrx_node_op *__restrict__ gops, const nrx_node_block *__restrict__ blocks,
                                                                    double *__restrict__ persite, size_t persite_stride, uint32_t nparts_total, uint32_t ncmax) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  NodeSmemFixed &sm = *reinterpret_cast<NodeSmemFixed *>(smem_raw);
  double *s_tile = reinterpret_cast<double *>(smem_raw + sizeof(NodeSmemFixed));          // [ncmax][NODE_TP * 16]
  uint32_t *s_sc = reinterpret_cast<uint32_t *>(s_tile + (size_t)ncmax * NODE_TP * 16);    // [ncmax][NODE_TP]
  const PartView &pv = parts[blockIdx.z];
  const nrx_node_block blk = blocks[blockIdx.x];
  const int tid = threadIdx.x, cat = tid & 3, pl = tid >> 2, lane = tid & 31;
  const uint32_t ntiles = (pv.patterns + NODE_TP - 1) / NODE_TP;
  if (blk.tile0 >= ntiles) return;
  if (tid == 0) {
    mbar_init(&sm.bar, 1);
    mbar_init(&sm.bar_desc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const nrx_node_group *g = groups + blk.group;
    const uint32_t nops = g->nops;   // (one global read; the copy below brings the rest)
    mbar_expect_tx(&sm.bar_desc, (uint32_t)sizeof(nrx_node_group) + nops * (uint32_t)sizeof(nrx_node_op));
    bulk_g2s(&sm.grp, g, (uint32_t)sizeof(nrx_node_group), &sm.bar_desc);
    bulk_g2s(sm.ops, gops + g->op_first, nops * (uint32_t)sizeof(nrx_node_op), &sm.bar_desc);
  }
  __syncthreads();
  mbar_wait(&sm.bar_desc, 0);
  const uint32_t nl = sm.grp.nl, nc = sm.grp.nl + sm.grp.nr, nops = sm.grp.nops;
  {   // input / output pointers of the group for this partition, fetched once (slot tables live in global memory)
    double *const *clv_tab = pv.clv;
    uint32_t *const *sc_tab = pv.scaler;
    for (uint32_t i = tid; i < nops; i += NODE_THREADS) { sm.par[i] = clv_tab[sm.ops[i].parent_slot]; sm.psc[i] = sc_tab[sm.ops[i].parent_slot]; }
    if ((uint32_t)tid < nc) { sm.cclv[tid] = clv_tab[sm.grp.child_slot[tid]]; sm.csc[tid] = sc_tab[sm.grp.child_slot[tid]]; }
  }
  double PL[16], PR[16];
  {
    const double *a = pv.pmat + (size_t)sm.grp.left_edge * 64 + cat * 16, *b = pv.pmat + (size_t)sm.grp.right_edge * 64 + cat * 16;
#pragma unroll
    for (int i = 0; i < 16; ++i) { PL[i] = a[i]; PR[i] = b[i]; }
  }
  const bool emit_any = persite != nullptr;
  double f0 = 0, f1 = 0, f2 = 0, f3 = 0, wcat = 0;
  if (emit_any) { f0 = pv.freqs[0]; f1 = pv.freqs[1]; f2 = pv.freqs[2]; f3 = pv.freqs[3]; wcat = pv.rate_weights[cat]; }
  const unsigned quad = 0xFu << (lane & ~3);
  __syncthreads();
  uint32_t phase = 0;
  for (uint32_t t = blk.tile0; t < ntiles; t += blk.stride) {
    const size_t p0 = (size_t)t * NODE_TP;
    if (tid == 0) {
      mbar_expect_tx(&sm.bar, nc * (uint32_t)(NODE_TP * 128 + NODE_TP * 4));
      for (uint32_t c = 0; c < nc; ++c) {
        bulk_g2s(s_tile + (size_t)c * NODE_TP * 16, sm.cclv[c] + p0 * 16, NODE_TP * 128u, &sm.bar);
        bulk_g2s(s_sc + c * NODE_TP, sm.csc[c] + p0, NODE_TP * 4u, &sm.bar);
      }
    }
    mbar_wait(&sm.bar, phase);
    phase ^= 1u;
    // step 2: P . child once per distinct child, in place (own 32 bytes)
    for (uint32_t c = 0; c < nc; ++c) {
      D4 *v = reinterpret_cast<D4 *>(s_tile + (size_t)c * NODE_TP * 16 + tid * 4);
      *v = (c < nl) ? matvec4_reg(PL, *v) : matvec4_reg(PR, *v);
    }
    // step 3: the ops
    const size_t site = p0 + pl;
    const bool act = site < pv.patterns;
    const size_t out_off = (site * 4 + cat) * 4;
    for (uint32_t i = 0; i < nops; ++i) {
      const nrx_node_op op = sm.ops[i];
      const D4 x = *reinterpret_cast<const D4 *>(s_tile + (size_t)op.li * NODE_TP * 16 + tid * 4), y = *reinterpret_cast<const D4 *>(s_tile + (size_t)(nl + op.rj) * NODE_TP * 16 + tid * 4);
      D4 p;
      p.x = __dmul_rn(x.x, y.x); p.y = __dmul_rn(x.y, y.y); p.z = __dmul_rn(x.z, y.z); p.w = __dmul_rn(x.w, y.w);
      const bool small = act & (p.x < SCALE_THRESHOLD) & (p.y < SCALE_THRESHOLD) & (p.z < SCALE_THRESHOLD) & (p.w < SCALE_THRESHOLD);
      const unsigned b = __ballot_sync(0xffffffffu, small);
      const bool scale = (b & quad) == quad;
      const uint32_t s = s_sc[op.li * NODE_TP + pl] + s_sc[(nl + op.rj) * NODE_TP + pl] + (scale ? 1u : 0u);
      if (scale) { p.x = __dmul_rn(p.x, SCALE_FACTOR); p.y = __dmul_rn(p.y, SCALE_FACTOR); p.z = __dmul_rn(p.z, SCALE_FACTOR); p.w = __dmul_rn(p.w, SCALE_FACTOR); }
      if (act) {
        stg256(sm.par[i] + out_off, p);
        if (cat == 0) sm.psc[i][site] = s;
      }
      if (emit_any && op.lnl_item) {   // fused K3, first half (as k_clv_dna4_pipe2)
        double tt = 0.0;
        if (act) tt = __dmul_rn(tree4(__dmul_rn(f0, p.x), __dmul_rn(f1, p.y), __dmul_rn(f2, p.z), __dmul_rn(f3, p.w)), wcat);
        const double t1 = __shfl_down_sync(0xffffffffu, tt, 1), t2 = __shfl_down_sync(0xffffffffu, tt, 2), t3 = __shfl_down_sync(0xffffffffu, tt, 3);
        if (act && cat == 0) persite[((size_t)(op.lnl_item - 1) * nparts_total + pv.part_index) * persite_stride + site] = __dadd_rn(__dadd_rn(__dadd_rn(tt, t1), t2), t3);
      }
    }
    __syncthreads();   // everyone is done with the tiles before the next load overwrites them
  }
}

/* ------------------------------------------------------------------------------------------------
 * Whole-evaluation "tile walk" (round 2), DNA 4x4: ONE launch computes every CLV of a traversal plan AND the per-tree root
 * lnLs.  The post-order dependencies of a likelihood traversal are per PATTERN: a block that owns a tile of WALK_TP patterns can
 * run the whole plan for it from the tips to the root without ever synchronising with another block.  So, instead of one
 * launch per dependency level (config 1: 12 dependent launches of ~3 us each = launch-latency-bound at 90-107 us per evaluation,
 * profiles/r2b_*), each block walks the op list in a depth-first order chosen by the host (nrx_plan_create) such that the CLVs
 * still needed later stay in a small set of shared-memory buffers:
 *   - children are read from SHARED MEMORY (the parent written a few ops earlier by the same block), never from HBM;
 *   - every CLV + scaler is still streamed out to its HBM slot (incremental updates, branch-length optimisation and the
 *     parity read-back use the slots exactly as before) — HBM traffic of an evaluation drops from (children read + parents
 *     written) to (parents written + tips read);
 *   - the P-matrices of all edges sit in shared memory (bulk-copied from the array K1 wrote), tip codes of the tile too;
 *   - ops marked as root displayed trees finish their per-site lnL in place (log, scaler term, pattern weight) and the block
 *     keeps one partial sum per tree; the last block (ticket) adds the partials of all tiles in a fixed order.
 * Arithmetic per (pattern, category) is the same instruction sequence as k_clv_dna4_pipe2 (P rows in registers per op, pairwise
 * sums, separate multiply / add), so CLVs and scalers stay bit-identical to libpll; per-tree lnLs are summed in a different
 * (fixed) order than k_tree_lnl_dna4's and agree with it to rounding.
 * grid = (tiles, 1, partitions of the class); block = WALK_TP x 4 threads: thread = (pattern, rate category).
 * ---------------------------------------------------------------------------------------------- */
constexpr int WALK_TP = 32;                 // patterns per tile
constexpr int WALK_THREADS = WALK_TP * 4;   // thread = (pattern, rate category): 4 CLV entries per op.  (One entry per thread —
                                            // 512 threads — was measured: 2.7x the instructions, the op decode is paid per warp, 57 vs 40 us.)
constexpr int WALK_HELPERS = 128;           // extra threads that share the prologue (P-matrices of all edges, tip codes, pointers) and then exit:
                                            // a third of the kernel was that prologue at five expm1 per thread (ncu, gpurun_out/r3i_walk.ncu-rep)
constexpr int WALK_BLOCK = WALK_THREADS + WALK_HELPERS;
__device__ __forceinline__ void walk_bar() { asm volatile("bar.sync 1, %0;" ::"n"(WALK_THREADS) : "memory"); }   // the 128 walking threads only
constexpr uint32_t WALK_NOBUF = 0xffffu;
constexpr int WALK_PE = 4 * PCAT;           // doubles per edge in the shared-memory P table (4 categories x (16 + 2 padding))

struct nrx_walk_op {   // 32 bytes; buffers index the block's shared-memory CLV buffers
  uint32_t parent_slot;
  uint16_t lbuf, rbuf, pbuf, kinds;   // kinds = left_kind | right_kind << 2; pbuf == WALK_NOBUF: nobody reads this CLV again
  uint32_t left_idx, right_idx;       // tip number for NRX_TIP
  uint32_t left_edge, right_edge;
  uint32_t lnl_item;
};

__device__ __forceinline__ D4 walk_tip_vec(const double *__restrict__ P /* this category's 4x4 block in shared memory, 16-byte aligned */, uint32_t m) {
  // masked row sums, same expression as build_tip_lut4 (core_partials_avx.c:1372-1400)
  const double2 *q = reinterpret_cast<const double2 *>(P);
  const bool m0 = m & 1, m1 = m & 2, m2 = m & 4, m3 = m & 8;
  D4 r;
  double2 a, b;
  a = q[0]; b = q[1]; r.x = tree4(m0 ? a.x : 0.0, m1 ? a.y : 0.0, m2 ? b.x : 0.0, m3 ? b.y : 0.0);
  a = q[2]; b = q[3]; r.y = tree4(m0 ? a.x : 0.0, m1 ? a.y : 0.0, m2 ? b.x : 0.0, m3 ? b.y : 0.0);
  a = q[4]; b = q[5]; r.z = tree4(m0 ? a.x : 0.0, m1 ? a.y : 0.0, m2 ? b.x : 0.0, m3 ? b.y : 0.0);
  a = q[6]; b = q[7]; r.w = tree4(m0 ? a.x : 0.0, m1 ? a.y : 0.0, m2 ? b.x : 0.0, m3 ? b.y : 0.0);
  return r;
}

struct WalkCtx {
  const double *sP; double *sClv; uint32_t *sSc; const uint8_t *sTip;
  int tid, pl, cat; bool act; unsigned quad;
};

/* one op, specialised on the operand kinds (no predication on them inside): the thread's category block of the parent CLV
 * (scaled if the pattern scales) and the pattern's scaler.  Expressions as in k_clv_dna4_pipe2 / build_tip_lut4. */
template <int LK, int RK>
__device__ __forceinline__ D4 walk_op(const WalkCtx &c, const nrx_walk_op &op, uint32_t &s_out) {
  D4 x, y, p;
  uint32_t s = 0;
  if (LK == NRX_CLV) { x = matvec4(c.sP + op.left_edge * WALK_PE + c.cat * PCAT, *reinterpret_cast<const D4 *>(c.sClv + ((size_t)op.lbuf * WALK_TP * 4 + c.tid) * 4)); s += c.sSc[op.lbuf * WALK_TP + c.pl]; }
  else if (LK == NRX_TIP) x = walk_tip_vec(c.sP + op.left_edge * WALK_PE + c.cat * PCAT, c.sTip[op.left_idx * WALK_TP + c.pl] & 15u);
  if (RK == NRX_CLV) { y = matvec4(c.sP + op.right_edge * WALK_PE + c.cat * PCAT, *reinterpret_cast<const D4 *>(c.sClv + ((size_t)op.rbuf * WALK_TP * 4 + c.tid) * 4)); s += c.sSc[op.rbuf * WALK_TP + c.pl]; }
  else if (RK == NRX_TIP) y = walk_tip_vec(c.sP + op.right_edge * WALK_PE + c.cat * PCAT, c.sTip[op.right_idx * WALK_TP + c.pl] & 15u);
  if (RK == NRX_NONE) p = x;
  else if (LK == NRX_NONE) p = y;
  else { p.x = __dmul_rn(x.x, y.x); p.y = __dmul_rn(x.y, y.y); p.z = __dmul_rn(x.z, y.z); p.w = __dmul_rn(x.w, y.w); }
  if (LK == NRX_TIP && RK == NRX_TIP) { s_out = 0; return p; }   // no scaling test in the tip-tip case (core_partials.c:371-470)
  const bool small = c.act & (p.x < SCALE_THRESHOLD) & (p.y < SCALE_THRESHOLD) & (p.z < SCALE_THRESHOLD) & (p.w < SCALE_THRESHOLD);
  const unsigned b = __ballot_sync(0xffffffffu, small);
  const bool scale = (b & c.quad) == c.quad;
  s_out = s + (scale ? 1u : 0u);
  if (scale) { p.x = __dmul_rn(p.x, SCALE_FACTOR); p.y = __dmul_rn(p.y, SCALE_FACTOR); p.z = __dmul_rn(p.z, SCALE_FACTOR); p.w = __dmul_rn(p.w, SCALE_FACTOR); }
  return p;
}

/* compute_p != 0: the branch lengths changed since K1 last ran — the block computes ALL P-matrices itself from `brlen`
 * ([partition][edges], same arithmetic as k_pmatrix, so bit-identical) instead of copying K1's table, and the blocks of tile 0
 * write them back to the partition's pmat / pmat_pad arrays for the kernels that run later (K4, incremental K2). */
__global__ void __launch_bounds__(WALK_BLOCK) k_walk_dna4(const PartView *__restrict__ parts, const nrx_walk_op *__restrict__ prog,
                                                            uint32_t nops, uint32_t nbuf, uint32_t nitems, double log_thresh,
                                                            double *__restrict__ partial /* [item][part][tiles] */, uint32_t nparts_total,
                                                            double *__restrict__ out /* [item][part] */, uint32_t *__restrict__ counter,
                                                            int compute_p, const double *__restrict__ brlen) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const PartView &pv = parts[blockIdx.z];
  const uint32_t ntiles = gridDim.x;
  const uint32_t tile = blockIdx.x;
  const int tid = threadIdx.x, pl = tid >> 2, cat = tid & 3, lane = tid & 31, warp = tid >> 5;
  const uint32_t edges = pv.edges, tips = pv.tips, patterns = pv.patterns;
  // shared-memory carve-up: P-matrices | CLV buffers | scaler buffers | per-tree block sums | program | output pointers | tip codes
  double *sP = reinterpret_cast<double *>(smem_raw);                                  // [edges][4][PCAT]: category blocks 18 doubles apart (disjoint banks)
  double *sClv = sP + (size_t)edges * WALK_PE;                                        // [nbuf][WALK_TP * 16]
  uint32_t *sSc = reinterpret_cast<uint32_t *>(sClv + (size_t)nbuf * WALK_TP * 16);  // [nbuf][WALK_TP]
  double *sAcc = reinterpret_cast<double *>(sSc + (size_t)nbuf * WALK_TP);            // [nitems][WALK_THREADS / 32]
  nrx_walk_op *sProg = reinterpret_cast<nrx_walk_op *>(sAcc + (size_t)nitems * (WALK_THREADS / 32));   // [nops]
  double **sPar = reinterpret_cast<double **>(sProg + nops);                          // [nops] this partition's parent CLV pointers
  uint32_t **sPsc = reinterpret_cast<uint32_t **>(sPar + nops);                       // [nops] ... and scaler pointers
  uint8_t *sTip = reinterpret_cast<uint8_t *>(sPsc + nops);                           // [tips][WALK_TP]
  __shared__ unsigned long long bar;
  __shared__ double sModel[16 + 16 + 4 + 4];   // inv_eigenvecs | eigenvecs | eigenvals | rates (compute_p)
  __shared__ int s_last;
  const size_t site0 = (size_t)tile * WALK_TP;
  const bool tile_live = site0 < patterns;   // partitions of a class may differ in length: surplus tiles only take their ticket

  if (tile_live) {
    if (tid == 0) {
      mbar_init(&bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      // the program and (unless this launch computes them) all P-matrices at the pitch K1 writes for this kernel: bulk copies
      const uint32_t pbytes = compute_p ? 0u : edges * (uint32_t)(WALK_PE * 8), gbytes = nops * (uint32_t)sizeof(nrx_walk_op);
      mbar_expect_tx(&bar, pbytes + gbytes);
      if (!compute_p) bulk_g2s(sP, pv.pmat_pad, pbytes, &bar);
      bulk_g2s(sProg, prog, gbytes, &bar);
    }
    // tip codes of this tile (rows are padded to whole 128-pattern tiles: always in bounds)
    if (compute_p) {   // model constants first: their latency hides under the other prologue loads
      if (tid < 16) { sModel[tid] = pv.inv_eigenvecs[tid]; sModel[16 + tid] = pv.eigenvecs[tid]; }
      if (tid < 4) { sModel[32 + tid] = pv.eigenvals[tid]; sModel[36 + tid] = pv.rates[tid]; }
    }
    for (uint32_t i = tid; i < tips * (WALK_TP / 4); i += WALK_BLOCK) {
      const uint32_t t = i / (WALK_TP / 4), w = i % (WALK_TP / 4);
      reinterpret_cast<uint32_t *>(sTip)[i] = *reinterpret_cast<const uint32_t *>(pv.tipchars + (size_t)t * pv.tip_pitch + site0 + 4 * w);
    }
    for (uint32_t i = tid; i < nitems * (WALK_THREADS / 32); i += WALK_BLOCK) sAcc[i] = 0.0;
    {   // the output pointers of every op, fetched up front (two dependent global loads in front of every store otherwise)
      double *const *clv_tab = pv.clv;
      uint32_t *const *sc_tab = pv.scaler;
      for (uint32_t i = tid; i < nops; i += WALK_BLOCK) { const uint32_t slot = prog[i].parent_slot; sPar[i] = clv_tab[slot]; sPsc[i] = sc_tab[slot]; }
    }
    if (compute_p) {
      /* K1 in the block (k_pmatrix's arithmetic, 4 states: (eval * rate) * t, expm1, pairwise sum of iev[j][m] ex[m] ev[m][k],
       * identity added last; t == 0 -> identity).  The expm1 values of an edge go through the padding doubles' neighbours:
       * stage 1 writes ex[c][m] into sP[edge][c][m] (entries 0..3), stage 2 reads them into registers before overwriting. */
      __syncthreads();
      const double *bl = brlen + (size_t)pv.part_index * edges;
      double *sEx = sClv;   // scratch: [edges][16] expm1 values (the CLV buffers are not in use yet); nbuf * 512 >= edges * 16 is checked by the host
      for (uint32_t i = tid; i < edges * 16u; i += WALK_BLOCK) {
        const uint32_t e = i >> 4, c = (i >> 2) & 3u, m = i & 3u;
        sEx[i] = expm1(__dmul_rn(__dmul_rn(sModel[32 + m], sModel[36 + c]), bl[e]));
      }
      __syncthreads();
      for (uint32_t i = tid; i < edges * 64u; i += WALK_BLOCK) {
        const uint32_t e = i >> 6, c = (i >> 4) & 3u, j = (i >> 2) & 3u, k = i & 3u;
        const double *ex = sEx + e * 16 + c * 4, *iev = sModel, *ev = sModel + 16;
        double v;
        if (bl[e] > 0.0) {
          const double p0 = __dmul_rn(__dmul_rn(iev[j * 4 + 0], ex[0]), ev[0 * 4 + k]);
          const double p1 = __dmul_rn(__dmul_rn(iev[j * 4 + 1], ex[1]), ev[1 * 4 + k]);
          const double p2 = __dmul_rn(__dmul_rn(iev[j * 4 + 2], ex[2]), ev[2 * 4 + k]);
          const double p3 = __dmul_rn(__dmul_rn(iev[j * 4 + 3], ex[3]), ev[3 * 4 + k]);
          v = __dadd_rn(tree4(p0, p1, p2, p3), (j == k) ? 1.0 : 0.0);
        } else {
          v = (j == k) ? 1.0 : 0.0;
        }
        sP[e * WALK_PE + c * PCAT + j * 4 + k] = v;
        if (tile == 0) { pv.pmat_pad[(size_t)e * WALK_PE + c * PCAT + j * 4 + k] = v; const_cast<double *>(pv.pmat)[i] = v; }
      }
    }
    __syncthreads();
    if (tid >= WALK_THREADS) return;   // the helpers are done; from here on only walk_bar() (128 threads) synchronises
    mbar_wait(&bar, 0);

    WalkCtx c;
    c.sP = sP; c.sClv = sClv; c.sSc = sSc; c.sTip = sTip;
    c.tid = tid; c.pl = pl; c.cat = cat;
    const size_t site = site0 + pl;
    c.act = site < patterns;
    c.quad = 0xFu << (lane & ~3);
    const double f0 = pv.freqs[0], f1 = pv.freqs[1], f2 = pv.freqs[2], f3 = pv.freqs[3], wcat = pv.rate_weights[cat];
    const double pw = c.act ? (double)pv.weights[site] : 0.0;
    const size_t out_off = (site * 4 + cat) * 4;
    for (uint32_t i = 0; i < nops; ++i) {
      const nrx_walk_op op = sProg[i];
      uint32_t s = 0;
      D4 p;
#define NRX_WALK_CASE(L, R) case (L) | ((R) << 2): p = walk_op<L, R>(c, op, s); break;
      switch (op.kinds) {
        NRX_WALK_CASE(NRX_CLV, NRX_CLV) NRX_WALK_CASE(NRX_CLV, NRX_TIP) NRX_WALK_CASE(NRX_TIP, NRX_CLV) NRX_WALK_CASE(NRX_TIP, NRX_TIP)
        NRX_WALK_CASE(NRX_CLV, NRX_NONE) NRX_WALK_CASE(NRX_NONE, NRX_CLV) NRX_WALK_CASE(NRX_TIP, NRX_NONE) NRX_WALK_CASE(NRX_NONE, NRX_TIP)
        default: p.x = p.y = p.z = p.w = 0.0; break;
      }
#undef NRX_WALK_CASE
      if (op.pbuf != WALK_NOBUF) {
        *reinterpret_cast<D4 *>(sClv + ((size_t)op.pbuf * WALK_TP * 4 + tid) * 4) = p;
        if (cat == 0) sSc[op.pbuf * WALK_TP + pl] = s;
      }
      if (c.act) {
        stg256(sPar[i] + out_off, p);
        if (cat == 0) sPsc[i][site] = s;
      }
      if (op.lnl_item) {   // root displayed tree: per-site lnL right here (same per-site arithmetic and order as k_tree_lnl_dna4)
        double t = 0.0;
        if (c.act) t = __dmul_rn(tree4(__dmul_rn(f0, p.x), __dmul_rn(f1, p.y), __dmul_rn(f2, p.z), __dmul_rn(f3, p.w)), wcat);
        const double t1 = __shfl_down_sync(0xffffffffu, t, 1), t2 = __shfl_down_sync(0xffffffffu, t, 2), t3 = __shfl_down_sync(0xffffffffu, t, 3);
        double lkv = 0.0;
        if (c.act && cat == 0) {
          lkv = log(__dadd_rn(__dadd_rn(__dadd_rn(t, t1), t2), t3));
          if (s) lkv = __dadd_rn(lkv, __dmul_rn((double)s, log_thresh));
          lkv = __dmul_rn(lkv, pw);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) lkv += __shfl_down_sync(0xffffffffu, lkv, off);
        if (lane == 0) sAcc[(op.lnl_item - 1) * (WALK_THREADS / 32) + warp] = lkv;
      }
      // A thread only ever reads the CLV entries it wrote itself (buffer index x tid) and the scaler its quad's category-0 lane
      // wrote: the op chain is independent per WARP (8 patterns), so a warp barrier orders everything — a block barrier here made
      // the four warps of a tile wait for each other 32 times per evaluation (ncu: 15 % of the samples on the barrier)
      __syncwarp();
    }
    walk_bar();
    // block sums of the marked trees, warps in order
    for (uint32_t it = tid; it < nitems; it += WALK_THREADS) {
      double sum = 0.0;
#pragma unroll
      for (int w = 0; w < WALK_THREADS / 32; ++w) sum += sAcc[it * (WALK_THREADS / 32) + w];
      partial[((size_t)it * nparts_total + pv.part_index) * ntiles + tile] = sum;
    }
  } else {
    if (tid >= WALK_THREADS) return;
    for (uint32_t it = tid; it < nitems; it += WALK_THREADS) partial[((size_t)it * nparts_total + pv.part_index) * ntiles + tile] = 0.0;
  }
  // second stage, fused: the block that draws the last ticket of this partition adds the tiles' partial sums in a fixed order
  walk_bar();
  if (tid == 0) {
    __threadfence();
    s_last = (atomicAdd(counter + pv.part_index, 1u) == ntiles - 1u) ? 1 : 0;
  }
  walk_bar();
  if (!s_last) return;
  __threadfence();
  for (uint32_t it = warp; it < nitems; it += WALK_THREADS / 32) {
    const double *pp = partial + ((size_t)it * nparts_total + pv.part_index) * ntiles;
    double sum = 0.0;
    for (uint32_t b = lane; b < ntiles; b += 32) sum += __ldcg(pp + b);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, off);
    if (lane == 0) out[(size_t)it * nparts_total + pv.part_index] = sum;
  }
  if (tid == 0) counter[pv.part_index] = 0u;
}

/* ------------------------------------------------------------------------------------------------
 * K2, protein: 20 states x 4 rate categories on the FP64 tensor cores (DMMA, mma.sync m8n8k4 f64).
 *
 * Here the per-node update IS a dense contraction: x[item][i] = sum_j P_c[i][j] * clv[item][j] with a 20x20
 * matrix per category (AI 3.3 flop/B, reference body LIBPLL/core_partials_avx2.c:630-818).  Per category:
 * D[8 items x 8 states] += A[8 items x 4] * B[4 x 8] with A = the CLV rows as they lie in memory (row-major),
 * B = P_c as it lies in memory ("col" layout) — 3 n-tiles (20 states padded to 24, padding rows of P are zero)
 * x 5 k-steps = 15 DMMAs per operand per 8 items.  A warp owns ONE category for the whole block, so its 2 x 15
 * B fragments (both child edges) stay in registers; the four category warps of a block share 8-pattern tiles.
 * A producer warp streams the tiles with per-row cp.async.bulk copies (640 B per pattern, destination pitch
 * 672 B so that the 8 rows of an A fragment fall into different banks) through an NSTAGE_AA-deep ring with
 * full/empty mbarriers; consumers pull their 10 A-fragment values + scalers into registers and release the
 * stage before the math.  Tip operands use a per-block table lut[code][cat][state] = sum_{j in mask} P[i][j]
 * (the reference's tip-inner precomputation, core_partials_avx2.c:344-420).  Scaling: all 80 entries of a
 * pattern < 2^-256 (core_partials.c:727-757) — AND over the quad, then over the 4 category warps through
 * shared flags and a 128-thread named barrier.
 * ---------------------------------------------------------------------------------------------- */
constexpr int AA_TP = 8;            // patterns per tile
constexpr int AA_PITCH = 84;        // doubles per staged row (80 + 4 padding)
constexpr int NSTAGE_AA = 4;
constexpr int AA_LUT_CODES = 32;    // tip table capacity (more distinct codes -> generic kernel)
constexpr int AA_THREADS = 160;     // 4 consumer warps (one per category) + 1 producer warp

struct __align__(128) AaStage {
  double l[AA_TP * AA_PITCH];
  double r[AA_TP * AA_PITCH];
  uint32_t scl[AA_TP];
  uint32_t scr[AA_TP];
  uint8_t tl[16];
  uint8_t tr[16];
};
struct __align__(128) AaSmem {
  AaStage st[NSTAGE_AA];
  unsigned long long full[NSTAGE_AA];
  unsigned long long empty[NSTAGE_AA];
  unsigned long long lutbar;
  uint32_t flags[8][4][AA_TP];
  double exch[2][4][AA_TP];   // AA_EDGE: per (category, pattern) weighted site-likelihood terms
  // followed by double lutL[AA_LUT_CODES*80], lutR[AA_LUT_CODES*80] when the launch has tip operands
};

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

/* K1b: tip tables of the edges whose P-matrix was just updated (grid.x = edges, one block each):
 * lut[edge][code][cat*20 + i] = sum_{j in states(code)} P_cat[i][j], serial order as core_partials.c:404-420. */
__global__ void k_tip_lut20(PartView pv, double *lut_out, const uint32_t *edge_idx) {
  const uint32_t edge = edge_idx[blockIdx.x];
  const double *pm = pv.pmat + (size_t)edge * 1600;
  double *lut = lut_out + (size_t)edge * AA_LUT_CODES * 80;
  for (uint32_t idx = threadIdx.x; idx < AA_LUT_CODES * 80; idx += blockDim.x) {
    const uint32_t code = idx / 80, ci = idx % 80;
    double sum = 0.0;
    if (code < pv.tip_codes) {
      const uint32_t mask = pv.tipmap[code];
      const double *row = pm + (size_t)ci * 20;
      for (int j = 0; j < 20; ++j) if ((mask >> j) & 1u) sum = __dadd_rn(sum, row[j]);
    }
    lut[idx] = sum;
  }
}

/* K2, protein, tip-tip ops: the parent CLV is a pure table product lutL[code_l] * lutR[code_r] (the reference's
 * tip-tip case, LIBPLL/core_partials.c:371-470: scaler := 0, no scaling test), i.e. a 640 B/pattern WRITE stream.
 * Both edge tables (K1b) sit in shared memory; thread = one double2 of the output, a warp writes 512 contiguous bytes. */
__global__ void __launch_bounds__(BLOCK) k_clv_aa20_tiptip(const PartView *__restrict__ parts, const nrx_op *__restrict__ ops, uint32_t chunk) {
  extern __shared__ __align__(16) double tt_lut[];
  const PartView &pv = parts[blockIdx.z];
  const nrx_op op = ops[blockIdx.y];
  const uint64_t p_lo = (uint64_t)blockIdx.x * chunk;
  if (p_lo >= pv.patterns) return;
  const uint64_t p_hi = (p_lo + chunk < pv.patterns) ? p_lo + chunk : pv.patterns;
  const uint32_t n_lut = pv.tip_codes * 80;
  double *lutL = tt_lut, *lutR = tt_lut + n_lut;
  const double *gl = pv.tiplut + (size_t)op.left_edge * AA_LUT_CODES * 80, *gr = pv.tiplut + (size_t)op.right_edge * AA_LUT_CODES * 80;
  for (uint32_t i = threadIdx.x; i < n_lut; i += BLOCK) { lutL[i] = gl[i]; lutR[i] = gr[i]; }
  __syncthreads();
  const uint8_t *tipL = pv.tipchars + (size_t)op.left_idx * pv.tip_pitch, *tipR = pv.tipchars + (size_t)op.right_idx * pv.tip_pitch;
  double2 *par = reinterpret_cast<double2 *>(pv.clv[op.parent_slot]);
  uint32_t *psc = pv.scaler[op.parent_slot];
  const double2 *l2 = reinterpret_cast<const double2 *>(lutL), *r2 = reinterpret_cast<const double2 *>(lutR);
  for (uint64_t f = p_lo * 40 + threadIdx.x; f < p_hi * 40; f += BLOCK) {
    const uint64_t pat = f / 40;
    const uint32_t r = (uint32_t)(f - pat * 40);
    const double2 a = l2[(uint32_t)tipL[pat] * 40 + r], b = r2[(uint32_t)tipR[pat] * 40 + r];
    double2 v;
    v.x = __dmul_rn(a.x, b.x); v.y = __dmul_rn(a.y, b.y);
    par[f] = v;
    if (r == 0) psc[pat] = 0u;
  }
}

/* The same pipeline serves all three 20-state contractions (MODE):
 *   AA_CLV  (K2): parent = (P_l . left) * (P_r . right), scaled, stored as a CLV slot.
 *   AA_SUM  (K5): sumtable = (A_L . left) * (A_R . right) with the category-independent eigen matrices of PartView::summat
 *                 (LIBPLL/core_derivatives.c:321-471 ii, :473-641 ti: the tip is the left operand), no scaling, stored
 *                 to sumtable slot op.parent_slot.
 *   AA_EDGE (K4): per pattern log( sum_c w_c sum_i pi_i left[c][i] (P_edge . right)[c][i] ) + scalers, times the pattern
 *                 weight, summed per block into partial[] (LIBPLL/core_likelihood.c:1191-1496 ii, :351-922 ti: the tip is
 *                 the child = right operand; the parent CLV = left operand is used as it lies, no matrix). */
enum { AA_CLV = 0, AA_SUM = 1, AA_EDGE = 2 };

template <int MODE>
__global__ void __launch_bounds__(AA_THREADS, 3) k_aa20_dmma(const PartView *__restrict__ parts, const nrx_op *__restrict__ ops,
                                                              uint32_t nops, uint32_t groups, int with_lut,
                                                              double *__restrict__ partial, uint32_t nparts_total, double log_thresh) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  AaSmem &sm = *reinterpret_cast<AaSmem *>(smem_raw);
  // with_lut: 0 = no tip operand in this launch, 1 = every op has at most ONE tip operand (both tables alias one
  // buffer of tip_codes x 640 B: 3 resident blocks per SM), 2 = tip-tip ops present (two tables)
  const PartView &pv = parts[blockIdx.z];
  double *lutL = reinterpret_cast<double *>(smem_raw + sizeof(AaSmem));
  double *lutR = (with_lut == 2) ? lutL + pv.tip_codes * 80 : lutL;
  const nrx_op op = ops[blockIdx.x % nops];
  const uint32_t grp = blockIdx.x / nops;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t ntiles = (pv.patterns + AA_TP - 1) / AA_TP;
  if (grp >= ntiles) return;
  const uint32_t count = (ntiles - grp + groups - 1) / groups;
  const int lk = op.left_kind, rk = op.right_kind;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < NSTAGE_AA; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 4); }
    mbar_init(&sm.lutbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (with_lut && tid == 0) {  // tip tables of the two child edges (precomputed by k_tip_lut20): one bulk copy each
    const uint32_t bytes = pv.tip_codes * 640u;
    const uint32_t tx = (lk == NRX_TIP ? bytes : 0u) + (rk == NRX_TIP ? bytes : 0u);
    if (tx) {
      mbar_expect_tx(&sm.lutbar, tx);
      if (lk == NRX_TIP) bulk_g2s(lutL, MODE == AA_SUM ? pv.sumlut : pv.tiplut + (size_t)op.left_edge * AA_LUT_CODES * 80, bytes, &sm.lutbar);
      if (rk == NRX_TIP) bulk_g2s(lutR, pv.tiplut + (size_t)op.right_edge * AA_LUT_CODES * 80, bytes, &sm.lutbar);
    }
  }
  const bool wait_lut = with_lut && (lk == NRX_TIP || rk == NRX_TIP);

  if (warp == 4) {
    /* ---------------- producer warp ---------------- */
    const double *clvL = (lk == NRX_CLV) ? pv.clv[op.left_idx] : nullptr;
    const double *clvR = (rk == NRX_CLV) ? pv.clv[op.right_idx] : nullptr;
    const uint32_t *scL = (lk == NRX_CLV) ? pv.scaler[op.left_idx] : nullptr;
    const uint32_t *scR = (rk == NRX_CLV) ? pv.scaler[op.right_idx] : nullptr;
    const uint8_t *tipL = (lk == NRX_TIP) ? pv.tipchars + (size_t)op.left_idx * pv.tip_pitch : nullptr;
    const uint8_t *tipR = (rk == NRX_TIP) ? pv.tipchars + (size_t)op.right_idx * pv.tip_pitch : nullptr;
    const uint32_t sc_bytes = (MODE == AA_SUM) ? 0u : AA_TP * 4u;   // sumtables ignore the scalers
    const uint32_t tx_bytes = ((lk == NRX_CLV) ? AA_TP * 640u + sc_bytes : (lk == NRX_TIP ? 16u : 0u)) +
                              ((rk == NRX_CLV) ? AA_TP * 640u + sc_bytes : (rk == NRX_TIP ? 16u : 0u));
    for (uint32_t k = 0; k < count; ++k) {
      const uint32_t s = k % NSTAGE_AA;
      if (k >= (uint32_t)NSTAGE_AA) mbar_wait(&sm.empty[s], ((k / NSTAGE_AA) - 1) & 1u);
      AaStage &st = sm.st[s];
      unsigned long long *bar = &sm.full[s];
      const size_t p0 = (size_t)(grp + (size_t)k * groups) * AA_TP;
      if (lane == 0) mbar_expect_tx(bar, tx_bytes);
      __syncwarp();
      if (lane < AA_TP) {           // lanes 0..7: left rows
        if (lk == NRX_CLV) bulk_g2s(st.l + lane * AA_PITCH, clvL + (p0 + lane) * 80, 640u, bar);
      } else if (lane < 2 * AA_TP) {  // lanes 8..15: right rows
        if (rk == NRX_CLV) bulk_g2s(st.r + (lane - AA_TP) * AA_PITCH, clvR + (p0 + lane - AA_TP) * 80, 640u, bar);
      } else if (lane == 16) {
        if (lk == NRX_CLV) { if (MODE != AA_SUM) bulk_g2s(st.scl, scL + p0, AA_TP * 4u, bar); }
        else if (lk == NRX_TIP) bulk_g2s(st.tl, tipL + (p0 & ~(size_t)15), 16u, bar);
      } else if (lane == 17) {
        if (rk == NRX_CLV) { if (MODE != AA_SUM) bulk_g2s(st.scr, scR + p0, AA_TP * 4u, bar); }
        else if (rk == NRX_TIP) bulk_g2s(st.tr, tipR + (p0 & ~(size_t)15), 16u, bar);
      }
    }
    return;
  }

  /* ---------------- consumer warps: warp = rate category ---------------- */
  const int cat = warp, item = lane >> 2, q = lane & 3;
  double BL[15], BR[15];   // B fragments: [ntile * 5 + kstep] = P_c[8*ntile + lane/4][4*kstep + lane%4]
  {
    const int i_base = lane >> 2, j_base = lane & 3;
#pragma unroll
    for (int n = 0; n < 3; ++n)
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const int i = 8 * n + i_base, j = 4 * k + j_base;
        if (MODE == AA_SUM) {
          BL[n * 5 + k] = (lk == NRX_CLV && i < 20) ? pv.summat[i * 20 + j] : 0.0;
          BR[n * 5 + k] = (rk == NRX_CLV && i < 20) ? pv.summat[400 + i * 20 + j] : 0.0;
        } else {
          BL[n * 5 + k] = (MODE == AA_CLV && lk == NRX_CLV && i < 20) ? pv.pmat[(size_t)op.left_edge * 1600 + (cat * 20 + i) * 20 + j] : 0.0;
          BR[n * 5 + k] = (rk == NRX_CLV && i < 20) ? pv.pmat[(size_t)op.right_edge * 1600 + (cat * 20 + i) * 20 + j] : 0.0;
        }
      }
  }
  double *par = (MODE == AA_SUM) ? pv.sumtable[op.parent_slot] : (MODE == AA_CLV ? pv.clv[op.parent_slot] : nullptr);
  uint32_t *psc = (MODE == AA_CLV) ? pv.scaler[op.parent_slot] : nullptr;
  const bool tiptip = (lk == NRX_TIP && rk == NRX_TIP);
  if (wait_lut) mbar_wait(&sm.lutbar, 0);
  double fr[6], wcat = 0.0, acc = 0.0;   // AA_EDGE: pi of this thread's six output states, rate weight of its category
  if (MODE == AA_EDGE) {
    wcat = pv.rate_weights[cat];
#pragma unroll
    for (int n = 0; n < 3; ++n)
#pragma unroll
      for (int h = 0; h < 2; ++h) { const int i = 8 * n + 2 * (lane & 3) + h; fr[2 * n + h] = (i < 20) ? pv.freqs[i] : 0.0; }
  }

  for (uint32_t k = 0; k < count; ++k) {
    const uint32_t s = k % NSTAGE_AA;
    const AaStage &st = sm.st[s];
    mbar_wait(&sm.full[s], (k / NSTAGE_AA) & 1u);
    const size_t p0 = (size_t)(grp + (size_t)k * groups) * AA_TP;
    const size_t site = p0 + item;
    const bool act = site < pv.patterns;
    // pull this warp's operands out of the stage, then release it
    double aL[5], aR[5];
    uint32_t codeL = 0, codeR = 0, sc = 0;
    double x[6], y[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) { x[i] = 0.0; y[i] = 0.0; }
    if (MODE == AA_EDGE) {   // the parent CLV enters as it lies: this thread's six output states, times pi
#pragma unroll
      for (int n = 0; n < 3; ++n) {
        const int i0 = 8 * n + 2 * q;
        if (i0 < 20) {
          const double2 v = *reinterpret_cast<const double2 *>(st.l + item * AA_PITCH + cat * 20 + i0);
          x[2 * n] = __dmul_rn(v.x, fr[2 * n]); x[2 * n + 1] = __dmul_rn(v.y, fr[2 * n + 1]);
        }
      }
      sc += st.scl[item];
    } else if (lk == NRX_CLV) {
#pragma unroll
      for (int kk = 0; kk < 5; ++kk) aL[kk] = st.l[item * AA_PITCH + cat * 20 + 4 * kk + q];
      if (MODE == AA_CLV) sc += st.scl[item];
    } else if (lk == NRX_TIP) codeL = st.tl[(p0 & 15) + item];
    if (rk == NRX_CLV) {
#pragma unroll
      for (int kk = 0; kk < 5; ++kk) aR[kk] = st.r[item * AA_PITCH + cat * 20 + 4 * kk + q];
      if (MODE != AA_SUM) sc += st.scr[item];
    } else if (rk == NRX_TIP) codeR = st.tr[(p0 & 15) + item];
    __syncwarp();
    if (lane == 0) mbar_arrive(&sm.empty[s]);

    // k-step outermost: the (up to) six accumulator chains advance together, so consecutive DMMAs are independent
#pragma unroll
    for (int kk = 0; kk < 5; ++kk) {
#pragma unroll
      for (int n = 0; n < 3; ++n) {
        if (MODE != AA_EDGE && lk == NRX_CLV) dmma884(x[2 * n], x[2 * n + 1], aL[kk], BL[n * 5 + kk]);
        if (rk == NRX_CLV) dmma884(y[2 * n], y[2 * n + 1], aR[kk], BR[n * 5 + kk]);
      }
    }
#pragma unroll
    for (int n = 0; n < 3; ++n) {
      const int i0 = 8 * n + 2 * q;   // this thread's two output states of n-tile n
      if (MODE != AA_EDGE && lk == NRX_TIP && i0 < 20) { const double2 v = *reinterpret_cast<const double2 *>(lutL + (codeL * 4 + cat) * 20 + i0); x[2 * n] = v.x; x[2 * n + 1] = v.y; }
      if (rk == NRX_TIP && i0 < 20) { const double2 v = *reinterpret_cast<const double2 *>(lutR + (codeR * 4 + cat) * 20 + i0); y[2 * n] = v.x; y[2 * n + 1] = v.y; }
    }
    double pz[6];
    bool small = true;
#pragma unroll
    for (int n = 0; n < 3; ++n) {
      const int i0 = 8 * n + 2 * q;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        double v;
        if (rk == NRX_NONE) v = x[2 * n + h];
        else if (lk == NRX_NONE) v = y[2 * n + h];
        else v = __dmul_rn(x[2 * n + h], y[2 * n + h]);
        pz[2 * n + h] = v;
        if (i0 < 20) small &= (v < SCALE_THRESHOLD);
      }
    }
    if (MODE == AA_EDGE) {
      // sum over this (pattern, category)'s 20 states: six per thread in state order, then the quad; weighted terms of
      // the four categories meet in shared memory (named barrier over the consumer warps), category order 0..3
      double t = 0.0;
#pragma unroll
      for (int i = 0; i < 6; ++i) t = __dadd_rn(t, pz[i]);   // padding states carry pi = 0
      t = __dadd_rn(t, __shfl_xor_sync(0xffffffffu, t, 1));
      t = __dadd_rn(t, __shfl_xor_sync(0xffffffffu, t, 2));
      if (q == 0) sm.exch[k & 1][cat][item] = (pv.pinv > 0.0) ? t : __dmul_rn(t, wcat);   // +I: the raw category term, weighted below
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (cat == 0 && q == 0 && act) {
        double lkv;
        if (pv.pinv > 0.0) {
          const int iv = pv.invariant[site];
          const double invf = iv < 0 ? 0.0 : pv.freqs[iv];
          double terma = 0.0, terminv = 0.0;
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) edge_cat_accum(sm.exch[k & 1][c4][item], pv.rate_weights[c4], pv.pinv, invf, iv >= 0, terma, terminv);
          lkv = edge_site_lnl(terma, terminv, sc, log_thresh);
        } else {
          const double term = __dadd_rn(__dadd_rn(__dadd_rn(sm.exch[k & 1][0][item], sm.exch[k & 1][1][item]), sm.exch[k & 1][2][item]), sm.exch[k & 1][3][item]);
          lkv = log(term);
          if (sc) lkv = __dadd_rn(lkv, __dmul_rn((double)sc, log_thresh));
        }
        acc += __dmul_rn(lkv, (double)pv.weights[site]);
      }
      continue;
    }
    if (MODE == AA_SUM) {
      if (act) {
        double *dst = par + site * 80 + cat * 20;
#pragma unroll
        for (int n = 0; n < 3; ++n) {
          const int i0 = 8 * n + 2 * q;
          if (i0 < 20) asm volatile("st.global.v2.f64 [%0], {%1,%2};" ::"l"(dst + i0), "d"(pz[2 * n]), "d"(pz[2 * n + 1]) : "memory");
        }
      }
      continue;
    }
    // all 20 states of (pattern, cat): AND over the quad; all 4 cats: flags in shared memory + a SPLIT named barrier.  A
    // pattern is scaled only if all four categories are small, so a warp none of whose eight items is small in ITS category
    // knows the answer (no scaling) without the others: it publishes its flags and only ARRIVES (non-blocking); a warp that
    // does need the other categories' flags SYNCs on the same barrier.  Underflow is rare, so the common path never waits
    // (the blocking 4-warp barrier per tile cost 10 % of the kernel).  Barrier ids / flag slots cycle over 8 tiles: warps of a
    // block cannot drift further apart than the NSTAGE_AA = 4 ring stages.
    unsigned b = __ballot_sync(0xffffffffu, small);
    const bool cat_small = ((b >> (lane & ~3)) & 0xFu) == 0xFu;
    bool scale = false;
    if (!tiptip) {
      const bool need = __any_sync(0xffffffffu, cat_small);
      if (q == 0) sm.flags[k & 7][cat][item] = cat_small ? 1u : 0u;
      const uint32_t bar_id = 1u + (k & 7u);
      if (!need) {
        __syncwarp();
        asm volatile("bar.arrive %0, 128;" ::"r"(bar_id) : "memory");
      } else {
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        scale = (sm.flags[k & 7][0][item] & sm.flags[k & 7][1][item] & sm.flags[k & 7][2][item] & sm.flags[k & 7][3][item]) != 0u;
      }
    }
    if (act) {
      double *dst = par + site * 80 + cat * 20;
#pragma unroll
      for (int n = 0; n < 3; ++n) {
        const int i0 = 8 * n + 2 * q;
        if (i0 < 20) {
          double v0 = pz[2 * n], v1 = pz[2 * n + 1];
          if (scale) { v0 = __dmul_rn(v0, SCALE_FACTOR); v1 = __dmul_rn(v1, SCALE_FACTOR); }
          asm volatile("st.global.v2.f64 [%0], {%1,%2};" ::"l"(dst + i0), "d"(v0), "d"(v1) : "memory");
        }
      }
      if (cat == 0 && q == 0) psc[site] = tiptip ? 0u : sc + (scale ? 1u : 0u);
    }
  }
  if (MODE == AA_EDGE && cat == 0) {   // the eight q == 0 lanes of the category-0 warp hold the block's sum
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, off);
    if (lane == 0) partial[((size_t)(blockIdx.x % nops) * nparts_total + pv.part_index) * groups + grp] = acc;
  }
}

/* ------------------------------------------------------------------------------------------------
 * K2 / K5, protein, round-2 kernel: k_aa20_mma<MODE> (MODE = AA_CLV or AA_SUM; AA_EDGE stays on k_aa20_dmma).
 *
 * What ncu said about k_aa20_dmma<AA_CLV> (profiles/r2a_aa20_dmma_stalls.md): the kernel is bound by the FP64 pipe — one
 * DMMA.8x8x4 holds an SM sub-partition's FP64 tensor path for 16 cycles (128 flop/clk/SM, the same peak as the DFMA pipe, and
 * DMUL competes for it) — yet that pipe was only 56-60 % busy: ~300 issued instructions per 30 DMMAs (address arithmetic,
 * predication on the operand kinds, the product / scaling-test / FSEL / partial-sector store epilogue, flags + named barrier)
 * executed in order by 3 warps per sub-partition.  Here the roles are split three ways so that the warps that own the
 * tensor pipe do almost nothing else:
 *   warps 4, 6  LOADERS  8-pattern tiles of the left / right child into an NIN-deep ring with 16-byte cp.async (see below);
 *   warps 0-3  MMA   (warp = rate category, both edges' B fragments in registers): 10 LDS of A fragments, 30 DMMAs, 6 DMUL,
 *                    the "< 2^-256" test (6 DSETP + ballot), then the UNSCALED products go to a shared-memory output tile
 *                    [8 patterns][80] (3 128-bit STS per lane) + one flag per (category, pattern);
 *   warp 5  STORER   waits for the four category warps of a tile, ANDs the flags (scaling needs all 80 entries of a pattern,
 *                    LIBPLL/core_partials.c:727-757; rows that do scale — rare — are multiplied in shared memory), writes
 *                    the scalers and issues ONE 5120-byte cp.async.bulk store per tile (8 consecutive patterns are contiguous in
 *                    the CLV): full-line, fully coalesced HBM writes by the TMA engine instead of 16-byte partial-sector stores.
 * All hand-offs are mbarriers (full_in / empty_in / full_out / empty_out); generic-proxy writes to the output tile are made
 * visible to the async proxy with fence.proxy.async before the storer is signalled.  Arithmetic (DMMA k-step order, products,
 * exact power-of-two scaling) is identical to k_aa20_dmma, so results are bit-identical to it.
 * ---------------------------------------------------------------------------------------------- */
constexpr int NIN_AA = 6;            // input ring stages (8 patterns x 2 operands x 672 B each)
constexpr int NOUT_AA = 3;           // output tiles in flight
constexpr int AA_OPITCH = 80;        // output rows dense (2-way conflict on 3 STS per lane and tile is noise): the tile leaves as ONE 5120-byte bulk store
constexpr int AA2_THREADS = 256;     // warpgroup 0: 4 MMA warps; warpgroup 1: left loader, storer, right loader, 1 idle warp (register donor)
constexpr int AA2_REGS_MMA = 200, AA2_REGS_AUX = 56;   // setmaxnreg budgets: 128 x 200 + 128 x 56 = 256 x 128 = the launch allocation

struct __align__(128) AaOut {
  double v[AA_TP * AA_OPITCH];
  uint32_t sc[AA_TP];              // sum of the children's scalers (written by the category-0 warp)
  uint32_t flag[4][AA_TP];         // 1: all 20 entries of (category, pattern) are below the scaling threshold
  double term[4][AA_TP];           // AA_EDGE: per (category, pattern) site-likelihood term sum_i pi_i parent_i (P child)_i
};
struct __align__(128) AaSmem2 {
  AaStage in[NIN_AA];
  AaOut out[NOUT_AA];
  unsigned long long full_in[NIN_AA], empty_in[NIN_AA], full_out[NOUT_AA], empty_out[NOUT_AA];
  unsigned long long lutbar;
  // followed by double lutL[tip_codes*80] (and lutR when with_lut == 2)
};

__device__ __forceinline__ void bulk_s2g(void *dst_gmem, const void *src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}

/* The MMA warps' loop, specialised on the operand kinds and software-pipelined: ncu on the straight version showed an MMA
 * warp only 38 % of its time in the DMMA region (24 % before it: barrier wait, A-fragment loads; 38 % after it: waiting for the
 * last DMMA, product, threshold test, stores, proxy fence), so two such warps kept a sub-partition's FP64 pipe ~50 % busy.
 * Here tile k+1's A-fragment loads and its 15-30 DMMAs are issued BEFORE tile k's epilogue, in one basic block (both barrier
 * waits first), with two accumulator sets: the epilogue's ~100 ALU / LSU instructions fill the issue slots under the 16-cycle
 * DMMAs instead of following them. */
struct AaFrag { double aL[5], aR[5]; uint32_t codeL, codeR, sc; };

template <int MODE, int LK, int RK, bool PIPE>
__device__ __forceinline__ void aa_mma_loop(AaSmem2 &sm, const PartView &pv, const nrx_op &op, const double *__restrict__ lutL,
                                            const double *__restrict__ lutR, uint32_t grp, uint32_t groups, uint32_t count, int cat, int lane) {
  const int item = lane >> 2, q = lane & 3;
  double BL[15], BR[15];   // B fragments: [ntile * 5 + kstep] = M[8*ntile + lane/4][4*kstep + lane%4]
#pragma unroll
  for (int n = 0; n < 3; ++n)
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const int i = 8 * n + item, j = 4 * k + q;
      if (MODE == AA_SUM) {
        BL[n * 5 + k] = (LK == NRX_CLV && i < 20) ? pv.summat[i * 20 + j] : 0.0;
        BR[n * 5 + k] = (RK == NRX_CLV && i < 20) ? pv.summat[400 + i * 20 + j] : 0.0;
      } else {
        BL[n * 5 + k] = (MODE == AA_CLV && LK == NRX_CLV && i < 20) ? pv.pmat[(size_t)op.left_edge * 1600 + (cat * 20 + i) * 20 + j] : 0.0;
        BR[n * 5 + k] = (RK == NRX_CLV && i < 20) ? pv.pmat[(size_t)op.right_edge * 1600 + (cat * 20 + i) * 20 + j] : 0.0;
      }
    }
  const uint32_t a_off = item * AA_PITCH + cat * 20 + q, o_off = item * AA_OPITCH + cat * 20 + 2 * q;
  double fr[6], wcat = 0.0;   // AA_EDGE: pi of this lane's six output states, rate weight of its category
  if (MODE == AA_EDGE) {
    wcat = pv.rate_weights[cat];
#pragma unroll
    for (int n = 0; n < 3; ++n)
#pragma unroll
      for (int h = 0; h < 2; ++h) { const int i = 8 * n + 2 * q + h; fr[2 * n + h] = (i < 20) ? pv.freqs[i] : 0.0; }
  }

  // pull tile k's operands out of its input stage (the stage's full barrier has been waited for) and release the stage
  double e_xl[2][6];   // AA_EDGE: pi-weighted parent entries of the two tiles in flight
  auto pull = [&](uint32_t k, AaFrag &f, int which) {
    const uint32_t s = k % NIN_AA;
    const AaStage &st = sm.in[s];
    const uint32_t p0lo = (uint32_t)(((size_t)(grp + (size_t)k * groups) * AA_TP) & 15u);
    f.sc = 0; f.codeL = 0; f.codeR = 0;
    if (MODE == AA_EDGE) {   // the parent CLV enters as it lies: this lane's six output states, times pi, parked in aL[0..2] / aR is the child
#pragma unroll
      for (int n = 0; n < 3; ++n) {
        const int i0 = 8 * n + 2 * q;
        double2 v = make_double2(0.0, 0.0);
        if (i0 < 20) v = *reinterpret_cast<const double2 *>(st.l + item * AA_PITCH + cat * 20 + i0);
        e_xl[which][2 * n] = __dmul_rn(v.x, fr[2 * n]); e_xl[which][2 * n + 1] = __dmul_rn(v.y, fr[2 * n + 1]);
      }
      f.sc += st.scl[item];
    } else if (LK == NRX_CLV) {
#pragma unroll
      for (int kk = 0; kk < 5; ++kk) f.aL[kk] = st.l[a_off + 4 * kk];
      if (MODE == AA_CLV) f.sc += st.scl[item];
    } else if (LK == NRX_TIP) f.codeL = st.tl[p0lo + item];
    if (RK == NRX_CLV) {
#pragma unroll
      for (int kk = 0; kk < 5; ++kk) f.aR[kk] = st.r[a_off + 4 * kk];
      if (MODE != AA_SUM) f.sc += st.scr[item];
    } else if (RK == NRX_TIP) f.codeR = st.tr[p0lo + item];
    __syncwarp();
    if (lane == 0) mbar_arrive(&sm.empty_in[s]);
  };
  auto mma = [&](const AaFrag &f, double (&x)[6], double (&y)[6]) {
#pragma unroll
    for (int i = 0; i < 6; ++i) { x[i] = 0.0; y[i] = 0.0; }
#pragma unroll
    for (int kk = 0; kk < 5; ++kk)
#pragma unroll
      for (int n = 0; n < 3; ++n) {
        if (MODE != AA_EDGE && LK == NRX_CLV) dmma884(x[2 * n], x[2 * n + 1], f.aL[kk], BL[n * 5 + kk]);
        if (RK == NRX_CLV) dmma884(y[2 * n], y[2 * n + 1], f.aR[kk], BR[n * 5 + kk]);
      }
  };
  // products, threshold test, output tile (its empty barrier has been waited for), flags, hand-over to the storer
  auto finish = [&](uint32_t k, const AaFrag &f, double (&x)[6], double (&y)[6], int which) {
    AaOut &ot = sm.out[k % NOUT_AA];
    if (MODE == AA_EDGE) {
      // sum over this (pattern, category)'s 20 states: six per lane in state order, then the quad (as k_aa20_dmma<AA_EDGE>)
      double t = 0.0;
#pragma unroll
      for (int n = 0; n < 3; ++n) {
        const int i0 = 8 * n + 2 * q;
        if (RK == NRX_TIP && i0 < 20) { const double2 v = *reinterpret_cast<const double2 *>(lutR + (f.codeR * 4 + cat) * 20 + i0); y[2 * n] = v.x; y[2 * n + 1] = v.y; }
        t = __dadd_rn(t, __dmul_rn(e_xl[which][2 * n], y[2 * n]));
        t = __dadd_rn(t, __dmul_rn(e_xl[which][2 * n + 1], y[2 * n + 1]));   // padding states carry pi = 0
      }
      t = __dadd_rn(t, __shfl_xor_sync(0xffffffffu, t, 1));
      t = __dadd_rn(t, __shfl_xor_sync(0xffffffffu, t, 2));
      if (q == 0) {
        ot.term[cat][item] = (pv.pinv > 0.0) ? t : __dmul_rn(t, wcat);   // +I: the raw category term, weighted by the storer
        if (cat == 0) ot.sc[item] = f.sc;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.full_out[k % NOUT_AA]);
      return;
    }
    bool small = true;
#pragma unroll
    for (int n = 0; n < 3; ++n) {
      const int i0 = 8 * n + 2 * q;
      if (LK == NRX_TIP && i0 < 20) { const double2 v = *reinterpret_cast<const double2 *>(lutL + (f.codeL * 4 + cat) * 20 + i0); x[2 * n] = v.x; x[2 * n + 1] = v.y; }
      if (RK == NRX_TIP && i0 < 20) { const double2 v = *reinterpret_cast<const double2 *>(lutR + (f.codeR * 4 + cat) * 20 + i0); y[2 * n] = v.x; y[2 * n + 1] = v.y; }
      double v0, v1;
      if (RK == NRX_NONE) { v0 = x[2 * n]; v1 = x[2 * n + 1]; }
      else if (LK == NRX_NONE) { v0 = y[2 * n]; v1 = y[2 * n + 1]; }
      else { v0 = __dmul_rn(x[2 * n], y[2 * n]); v1 = __dmul_rn(x[2 * n + 1], y[2 * n + 1]); }
      if (i0 < 20) {
        if (MODE == AA_CLV) small &= (v0 < SCALE_THRESHOLD) & (v1 < SCALE_THRESHOLD);
        *reinterpret_cast<double2 *>(ot.v + o_off + 8 * n) = make_double2(v0, v1);
      }
    }
    if (MODE == AA_CLV) {
      const unsigned b = __ballot_sync(0xffffffffu, small);
      if (q == 0) {
        ot.flag[cat][item] = (((b >> (lane & ~3)) & 0xFu) == 0xFu) ? 1u : 0u;
        if (cat == 0) ot.sc[item] = f.sc;
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the storer's bulk copy
    __syncwarp();
    if (lane == 0) mbar_arrive(&sm.full_out[k % NOUT_AA]);
  };
  auto wait_in = [&](uint32_t k) { mbar_wait(&sm.full_in[k % NIN_AA], (k / NIN_AA) & 1u); };
  auto wait_out = [&](uint32_t k) { mbar_wait(&sm.empty_out[k % NOUT_AA], ((k / NOUT_AA) & 1u) ^ 1u); };

  AaFrag f0, f1;
  double x0[6], y0[6], x1[6], y1[6];
  if (!PIPE) {   // straight order (measured: the CLV update is no faster pipelined — 3.27 vs 3.46 ms on config 4 at 200 k patterns —
                 // while the sumtable, which has no threshold test / flags in its epilogue, gains 17 %)
    for (uint32_t k = 0; k < count; ++k) {
      wait_in(k);
      pull(k, f0, 0);
      mma(f0, x0, y0);
      wait_out(k);
      finish(k, f0, x0, y0, 0);
    }
    return;
  }
  wait_in(0);
  pull(0, f0, 0);
  mma(f0, x0, y0);
  for (uint32_t k = 0; k < count; k += 2) {
    if (k + 1 < count) {   // tile k+1's DMMAs are issued first, tile k's epilogue fills the issue slots under them
      wait_in(k + 1);
      wait_out(k);
      pull(k + 1, f1, 1);
      mma(f1, x1, y1);
      finish(k, f0, x0, y0, 0);
    } else {
      wait_out(k);
      finish(k, f0, x0, y0, 0);
      break;
    }
    if (k + 2 < count) {
      wait_in(k + 2);
      wait_out(k + 1);
      pull(k + 2, f0, 0);
      mma(f0, x0, y0);
      finish(k + 1, f1, x1, y1, 1);
    } else {
      wait_out(k + 1);
      finish(k + 1, f1, x1, y1, 1);
    }
  }
}

template <int MODE, bool PIPE>
__global__ void __launch_bounds__(AA2_THREADS, 2) k_aa20_mma(const PartView *__restrict__ parts, const nrx_op *__restrict__ ops,
                                                              uint32_t nops, uint32_t groups, int with_lut,
                                                              double *__restrict__ partial, uint32_t nparts_total, double log_thresh,
                                                              double *__restrict__ red_out, uint32_t *__restrict__ counters, int pdl,
                                                              double *__restrict__ persite = nullptr, size_t persite_stride = 0) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  AaSmem2 &sm = *reinterpret_cast<AaSmem2 *>(smem_raw);
  // pdl: launched programmatically serialised behind the previous K2 launch of a plan graph — the successor may start its own
  // prologue (barriers, tip table, B fragments: all read data no K2 launch writes) while this grid drains; only the loader warps
  // read what the predecessor wrote and wait for it (griddepcontrol.wait) before their first copy
  if (pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const PartView &pv = parts[blockIdx.z];
  double *lutL = reinterpret_cast<double *>(smem_raw + sizeof(AaSmem2));
  double *lutR = (with_lut == 2) ? lutL + pv.tip_codes * 80 : lutL;
  const nrx_op op = ops[blockIdx.x % nops];
  const uint32_t grp = blockIdx.x / nops;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t ntiles = (pv.patterns + AA_TP - 1) / AA_TP;
  if (grp >= ntiles) return;
  const uint32_t count = (ntiles - grp + groups - 1) / groups;
  const int lk = op.left_kind, rk = op.right_kind;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < NIN_AA; ++s) { mbar_init(&sm.full_in[s], 64); mbar_init(&sm.empty_in[s], 4); }
#pragma unroll
    for (int s = 0; s < NOUT_AA; ++s) { mbar_init(&sm.full_out[s], 4); mbar_init(&sm.empty_out[s], 1); }
    mbar_init(&sm.lutbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (with_lut && tid == 0) {
    const uint32_t bytes = pv.tip_codes * 640u;
    const uint32_t tx = (lk == NRX_TIP ? bytes : 0u) + (rk == NRX_TIP ? bytes : 0u);
    if (tx) {
      mbar_expect_tx(&sm.lutbar, tx);
      if (lk == NRX_TIP) bulk_g2s(lutL, MODE == AA_SUM ? pv.sumlut : pv.tiplut + (size_t)op.left_edge * AA_LUT_CODES * 80, bytes, &sm.lutbar);
      if (rk == NRX_TIP) bulk_g2s(lutR, pv.tiplut + (size_t)op.right_edge * AA_LUT_CODES * 80, bytes, &sm.lutbar);
    }
  }
  const bool wait_lut = with_lut && (lk == NRX_TIP || rk == NRX_TIP);

  /* Register re-allocation between the roles (setmaxnreg, sm_90+): two resident blocks of 8 warps leave 128 registers per thread
   * at launch; the loader / storer warpgroup gives back down to 56 and the MMA warpgroup grows to 200 — room for both edges' B
   * fragments (60), TWO accumulator sets (48) and two tiles' A fragments (40) of the software-pipelined loop.  (Three blocks
   * of 4 MMA warps at 120 registers were measured too: no faster, the straight loop was the limit, not the warp count.) */
  if (warp >= 4) {   // (the role branches never re-join: ptxas budgets each side by its own setmaxnreg)
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(AA2_REGS_AUX));
    if (warp == 7) return;
  }

  if (warp == 4 || warp == 6) {
    /* ---------------- loaders: warp 4 streams the left operand, warp 6 the right one ----------------
     * ncu on the first cut (and on k_aa20_dmma, same producer): ONE warp issuing 18 per-row cp.async.bulk copies per 8-pattern
     * tile was the bottleneck — the uniform datapath issues them one lane at a time (ELECT / R2UR / UBLKCP / BRA.U.ANY loop, ~600
     * cycles per tile) while the MMA warps waited on full_in and the tensor pipe idled at 45 %.  The rows must land at a padded
     * pitch (672 B: the A-fragment reads of 8 rows then hit disjoint banks), which a single bulk copy cannot do; 16-byte cp.async
     * (LDGSTS) can: 10 fully coalesced 512-byte warp copies per operand and tile, completion counted on the stage's mbarrier
     * (cp.async.mbarrier.arrive.noinc, one arrival per lane of both loader warps = 64).  Keeping this loop free of register
     * spills mattered as much: at 40 registers it spilled its chunk offsets and K2 ran at 3.4 instead of 2.8 ms. */
    const bool left = warp == 4;
    const int kind = left ? lk : rk;
    const uint32_t idx = left ? op.left_idx : op.right_idx;
    const double *clv = (kind == NRX_CLV) ? pv.clv[idx] : nullptr;
    const uint32_t *sc = (kind == NRX_CLV && MODE != AA_SUM) ? pv.scaler[idx] : nullptr;
    const uint8_t *tip = (kind == NRX_TIP) ? pv.tipchars + (size_t)idx * pv.tip_pitch : nullptr;
    // chunk c = i * 32 + lane of a tile (320 chunks of 16 B): row c / 40, 16-byte column c % 40
    uint32_t soff[10], goff[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      const uint32_t c = (uint32_t)i * 32u + (uint32_t)lane;
      soff[i] = ((c / 40u) * AA_PITCH + (c % 40u) * 2u) * 8u;
      goff[i] = (c / 40u) * 80u + (c % 40u) * 2u;
    }
    uint32_t s = 0, ph = 0;
    if (pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
    for (uint32_t k = 0; k < count; ++k) {
      if (k >= (uint32_t)NIN_AA) mbar_wait(&sm.empty_in[s], ph ^ 1u);
      AaStage &st = sm.in[s];
      const size_t p0 = (size_t)(grp + (size_t)k * groups) * AA_TP;
      if (clv) {
        const double *g = clv + p0 * 80;
        const uint32_t d = smem_u32(left ? st.l : st.r);
#pragma unroll
        for (int i = 0; i < 10; ++i)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + soff[i]), "l"(g + goff[i]) : "memory");
        if (sc && lane < AA_TP)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32((left ? st.scl : st.scr) + lane)), "l"(sc + p0 + lane) : "memory");
      } else if (tip && lane == 0) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(left ? st.tl : st.tr)), "l"(tip + (p0 & ~(size_t)15)) : "memory");
      }
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&sm.full_in[s])) : "memory");
      if (++s == NIN_AA) { s = 0; ph ^= 1u; }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    return;
  }

  if (warp == 5) {
    /* ---------------- storer ---------------- */
    double *par = (MODE == AA_SUM) ? pv.sumtable[op.parent_slot] : (MODE == AA_CLV ? pv.clv[op.parent_slot] : nullptr);
    uint32_t *psc = (MODE == AA_CLV) ? pv.scaler[op.parent_slot] : nullptr;
    const bool tiptip = (lk == NRX_TIP && rk == NRX_TIP);
    double *ps_out = (MODE == AA_CLV && persite && op.lnl_item) ? persite + ((size_t)(op.lnl_item - 1) * nparts_total + pv.part_index) * persite_stride : nullptr;
    double edge_acc = 0.0, pend_t = 0.0, pend_w = 0.0;   // AA_EDGE: block sum; the pattern this lane finishes at the next flush
    uint32_t pend_s = 0;
    uint32_t o = 0, ph = 0, prev = 0;
    for (uint32_t k = 0; k < count; ++k) {
      AaOut &ot = sm.out[o];
      mbar_wait(&sm.full_out[o], ph);
      const size_t p0 = (size_t)(grp + (size_t)k * groups) * AA_TP;
      if (MODE == AA_CLV) {
        // lanes 0..7 = the tile's patterns: a pattern is scaled only if all four categories flagged it
        uint32_t f = 0, scv = 0;
        if (lane < AA_TP) {
          f = tiptip ? 0u : (ot.flag[0][lane] & ot.flag[1][lane] & ot.flag[2][lane] & ot.flag[3][lane]);
          scv = tiptip ? 0u : ot.sc[lane] + f;
          if (p0 + lane < pv.patterns) psc[p0 + lane] = scv;
        }
        const unsigned any = __ballot_sync(0xffffffffu, f != 0u);
        if (any) {   // rare: multiply the flagged rows by 2^256 in shared memory (exact), then hand them to the async proxy
          for (int r = 0; r < AA_TP; ++r)
            if ((any >> r) & 1u)
              for (int i = lane; i < 80; i += 32) ot.v[r * AA_OPITCH + i] = __dmul_rn(ot.v[r * AA_OPITCH + i], SCALE_FACTOR);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
        }
      }
      if (MODE == AA_CLV && ps_out) {
        // fused K3 (root displayed trees of a replayed plan, as the DNA kernel's epilogue does): the per-site likelihood term of the
        // tile as it leaves — lane = (pattern, category), pi-weighted sum over the 20 states in state order, category weight, categories
        // 0..3 in order (k_tree_lnl_aa20p's arithmetic); k_term_lnl_sum then takes log / scaler / weight from 16 B per site
        const int pp = lane >> 2, cc = lane & 3;
        const double *row = ot.v + pp * AA_OPITCH + cc * 20;
        double tr = 0.0;
#pragma unroll
        for (int i = 0; i < 20; ++i) tr = __dadd_rn(tr, __dmul_rn(row[i], __ldg(pv.freqs + i)));
        const double tw = __dmul_rn(tr, __ldg(pv.rate_weights + cc));
        double term = 0.0;
#pragma unroll
        for (int i = 0; i < 4; ++i) term = __dadd_rn(term, __shfl_sync(0xffffffffu, tw, (lane & ~3) + i));
        if (cc == 0 && p0 + pp < pv.patterns) ps_out[p0 + pp] = term;
      }
      if (MODE == AA_EDGE) {
        // K4: category sum (order 0..3) per pattern on lanes 0..7; the log / scaler term / pattern weight — ~150 dependent FP64
        // instructions — are batched over FOUR tiles so that all 32 lanes do them at once (one lane per pattern): done per tile on
        // 8 lanes, this warp was the slowest stage of the pipeline (K4 ran at 0.28 of the HBM peak, below k_aa20_dmma's 0.34)
        const uint32_t slot = k & 3u;
        double t = 0.0, w = 0.0;
        uint32_t scv = 0;
        if (lane < AA_TP && p0 + lane < pv.patterns) {
          const size_t site = p0 + lane;
          scv = ot.sc[lane];
          w = (double)pv.weights[site];
          if (pv.pinv > 0.0) {   // +I: finished here, per tile (terma / terminv do not batch as one number)
            const int iv = pv.invariant[site];
            const double invf = iv < 0 ? 0.0 : pv.freqs[iv];
            double terma = 0.0, terminv = 0.0;
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) edge_cat_accum(ot.term[c4][lane], pv.rate_weights[c4], pv.pinv, invf, iv >= 0, terma, terminv);
            edge_acc += __dmul_rn(edge_site_lnl(terma, terminv, scv, log_thresh), w);
            w = 0.0;
          } else {
            t = __dadd_rn(__dadd_rn(__dadd_rn(ot.term[0][lane], ot.term[1][lane]), ot.term[2][lane]), ot.term[3][lane]);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty_out[o]);   // nothing asynchronous reads the tile: free at once
        const double tt = __shfl_sync(0xffffffffu, t, lane & 7), ww = __shfl_sync(0xffffffffu, w, lane & 7);
        const uint32_t ss = __shfl_sync(0xffffffffu, scv, lane & 7);
        if ((uint32_t)(lane >> 3) == slot) { pend_t = tt; pend_w = ww; pend_s = ss; }
        if (slot == 3u || k + 1 == count) {
          if (pend_w != 0.0) {
            double lkv = log(pend_t);
            if (pend_s) lkv = __dadd_rn(lkv, __dmul_rn((double)pend_s, log_thresh));
            edge_acc += __dmul_rn(lkv, pend_w);
          }
          pend_w = 0.0;
        }
      } else if (lane == 0) {
        const uint32_t rows = (p0 + AA_TP <= pv.patterns) ? (uint32_t)AA_TP : (uint32_t)(pv.patterns - p0);   // consecutive patterns: one contiguous span
        bulk_s2g(par + p0 * 80, ot.v, rows * 640u);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        // all but the newest group have been READ out of shared memory: the previous tile's buffer can be rewritten
        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        if (k > 0) mbar_arrive(&sm.empty_out[prev]);
      }
      prev = o;
      if (++o == NOUT_AA) { o = 0; ph ^= 1u; }
    }
    if (MODE == AA_EDGE) {
      // block partial, then the fused second stage: the block drawing the last ticket of its (pair,
      // partition) sums all partials in the fixed lane-strided order (see finish_partials; one warp here, no block barrier)
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) edge_acc += __shfl_down_sync(0xffffffffu, edge_acc, off);
      const size_t oi = (size_t)(blockIdx.x % nops) * nparts_total + pv.part_index;
      int last = 0;
      if (lane == 0) {
        partial[oi * groups + grp] = edge_acc;
        if (counters) { __threadfence(); last = (atomicAdd(counters + oi, 1u) == groups - 1u) ? 1 : 0; }
      }
      last = __shfl_sync(0xffffffffu, last, 0);
      if (last) {
        __threadfence();
        double sum = 0.0;
        for (uint32_t b = lane; b < groups; b += 32) sum += __ldcg(partial + oi * groups + b);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, off);
        if (lane == 0) { red_out[oi] = sum; counters[oi] = 0u; }
      }
      return;
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // the stores must have landed before the block's shared memory goes away
    return;
  }

  /* ---------------- MMA warps: warp = rate category ---------------- */
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(AA2_REGS_MMA));
  if (wait_lut) mbar_wait(&sm.lutbar, 0);
  // the operand kinds are fixed per block: one specialised, software-pipelined loop per combination (no predication inside)
#define NRX_AA_CASE(L, R) case (L) * 3 + (R): aa_mma_loop<MODE, L, R, PIPE>(sm, pv, op, lutL, lutR, grp, groups, count, warp, lane); break;
  if (MODE == AA_EDGE) {   // pairs: the parent is always a CLV, the child a CLV or a tip
    switch (lk * 3 + rk) {
      NRX_AA_CASE(NRX_CLV, NRX_CLV) NRX_AA_CASE(NRX_CLV, NRX_TIP)
      default: break;
    }
    return;
  }
  switch (lk * 3 + rk) {
    NRX_AA_CASE(NRX_CLV, NRX_CLV) NRX_AA_CASE(NRX_CLV, NRX_TIP) NRX_AA_CASE(NRX_TIP, NRX_CLV)
    NRX_AA_CASE(NRX_CLV, NRX_NONE) NRX_AA_CASE(NRX_NONE, NRX_CLV) NRX_AA_CASE(NRX_TIP, NRX_NONE) NRX_AA_CASE(NRX_NONE, NRX_TIP)
    NRX_AA_CASE(NRX_TIP, NRX_TIP)
    default: break;
  }
#undef NRX_AA_CASE
}

/* ------------------------------------------------------------------------------------------------
 * K2 generic (any states <= 32, any cats): one thread per (pattern, category), P rows from L1/L2.
 * Used for protein data until the DMMA kernel takes over, and for unusual category counts.
 * ---------------------------------------------------------------------------------------------- */
__device__ __forceinline__ double masked_rowsum(const double *row, uint32_t S, uint32_t mask) {
  if (S == 4) return tree4((mask & 1) ? row[0] : 0.0, (mask & 2) ? row[1] : 0.0, (mask & 4) ? row[2] : 0.0, (mask & 8) ? row[3] : 0.0);
  double s = 0.0;
  for (uint32_t j = 0; j < S; ++j) if ((mask >> j) & 1) s = __dadd_rn(s, row[j]);
  return s;
}
__device__ __forceinline__ double row_dot(const double *row, const double *v, uint32_t S) {
  if (S == 4) return tree4(__dmul_rn(row[0], v[0]), __dmul_rn(row[1], v[1]), __dmul_rn(row[2], v[2]), __dmul_rn(row[3], v[3]));
  double s = 0.0;
  for (uint32_t j = 0; j < S; ++j) s = __dadd_rn(s, __dmul_rn(row[j], v[j]));
  return s;
}

__global__ void __launch_bounds__(BLOCK) k_clv_generic(const PartView *__restrict__ parts, const nrx_op *__restrict__ ops,
                                                        uint32_t *__restrict__ flags /* per-site "not all small" scratch */) {
  // Two-phase per pattern: (1) every (pattern,cat) thread writes its category block and atomically ANDs the
  // "all small" predicate into a per-pattern word; (2) handled by k_clv_generic_scale.  To stay simple and
  // correct for any category count, this kernel instead lets ONE thread own a whole pattern.
  const PartView &pv = parts[blockIdx.z];
  const nrx_op op = ops[blockIdx.y];
  const uint32_t S = pv.states, SP = pv.sp, C = pv.cats;
  const int lk = op.left_kind, rk = op.right_kind;
  const bool tiptip = (lk == NRX_TIP && rk == NRX_TIP);
  const double *pmL = pv.pmat + (size_t)op.left_edge * C * S * SP;
  const double *pmR = pv.pmat + (size_t)op.right_edge * C * S * SP;
  double *par = pv.clv[op.parent_slot];
  uint32_t *psc = pv.scaler[op.parent_slot];
  (void)flags;
  for (uint64_t n = (uint64_t)blockIdx.x * BLOCK + threadIdx.x; n < pv.patterns; n += (uint64_t)gridDim.x * BLOCK) {
    const uint32_t mL = (lk == NRX_TIP) ? pv.tipmap[pv.tipchars[(size_t)op.left_idx * pv.tip_pitch + n]] : 0;
    const uint32_t mR = (rk == NRX_TIP) ? pv.tipmap[pv.tipchars[(size_t)op.right_idx * pv.tip_pitch + n]] : 0;
    const double *cl = (lk == NRX_CLV) ? pv.clv[op.left_idx] + n * C * SP : nullptr;
    const double *cr = (rk == NRX_CLV) ? pv.clv[op.right_idx] + n * C * SP : nullptr;
    double *out = par + n * C * SP;
    bool all_small = true;
    for (uint32_t c = 0; c < C; ++c) {
      for (uint32_t i = 0; i < S; ++i) {
        const double *lrow = pmL + ((size_t)c * S + i) * SP, *rrow = pmR + ((size_t)c * S + i) * SP;
        double x = 1.0, y = 1.0;
        if (lk == NRX_CLV) x = row_dot(lrow, cl + c * SP, S); else if (lk == NRX_TIP) x = masked_rowsum(lrow, S, mL);
        if (rk == NRX_CLV) y = row_dot(rrow, cr + c * SP, S); else if (rk == NRX_TIP) y = masked_rowsum(rrow, S, mR);
        const double v = __dmul_rn(x, y);
        out[c * SP + i] = v;
        all_small &= (v < SCALE_THRESHOLD);
      }
      for (uint32_t i = S; i < SP; ++i) out[c * SP + i] = 0.0;
    }
    uint32_t s = 0;
    if (!tiptip) {
      if (lk == NRX_CLV) s += pv.scaler[op.left_idx][n];
      if (rk == NRX_CLV) s += pv.scaler[op.right_idx][n];
      if (all_small) {
        for (uint32_t i = 0; i < C * SP; ++i) out[i] = __dmul_rn(out[i], SCALE_FACTOR);
        s += 1;
      }
    }
    psc[n] = s;
  }
}

/* ------------------------------------------------------------------------------------------------
 * K2p  pseudo-likelihood CLV update (src/likelihood/PseudoLoglikelihood.cpp:57-190), fused: x = P_l . left and
 * y = P_r . right are computed ONCE; the three updates the reference runs (both / left only / right only, each with the
 * per-site scaling test of LIBPLL/core_partials*.c) and merge_clvs (:8-55, same term order, separate multiply and add)
 * happen in registers.  An absent operand is the fake all-ones CLV behind the identity matrix.  Scaler = that of the last
 * executed update (the reference's three calls share one scale buffer).
 * ---------------------------------------------------------------------------------------------- */
__global__ void __launch_bounds__(BLOCK) k_clv_pseudo_dna4(const PartView *__restrict__ parts, const nrx_pseudo_op *__restrict__ ops) {
  const PartView &pv = parts[blockIdx.z];
  const nrx_pseudo_op op = ops[blockIdx.y];
  const int tid = threadIdx.x, cat = tid & 3, lane = tid & 31;
  const uint64_t n_items = (uint64_t)pv.patterns * 4;
  __shared__ __align__(32) double lutL[256];
  __shared__ __align__(32) double lutR[256];
  __shared__ __align__(16) double sPL[4 * PCAT];
  __shared__ __align__(16) double sPR[4 * PCAT];
  const int lk = op.left_kind, rk = op.right_kind;
  if (lk == NRX_CLV) { if (tid < 64) sPL[(tid >> 4) * PCAT + (tid & 15)] = pv.pmat[(size_t)op.left_edge * 64 + tid]; }
  else if (lk == NRX_TIP) build_tip_lut4(lutL, pv.pmat + (size_t)op.left_edge * 64, tid);
  if (rk == NRX_CLV) { if (tid >= 64 && tid < 128) sPR[((tid - 64) >> 4) * PCAT + (tid & 15)] = pv.pmat[(size_t)op.right_edge * 64 + tid - 64]; }
  else if (rk == NRX_TIP) build_tip_lut4(lutR, pv.pmat + (size_t)op.right_edge * 64, tid);
  __syncthreads();
  const double *PL = sPL + cat * PCAT, *PR = sPR + cat * PCAT;
  const double *clvL = (lk == NRX_CLV) ? pv.clv[op.left_idx] : nullptr;
  const double *clvR = (rk == NRX_CLV) ? pv.clv[op.right_idx] : nullptr;
  const uint32_t *scL = (lk == NRX_CLV) ? pv.scaler[op.left_idx] : nullptr;
  const uint32_t *scR = (rk == NRX_CLV) ? pv.scaler[op.right_idx] : nullptr;
  const uint8_t *tipL = (lk == NRX_TIP) ? pv.tipchars + (size_t)op.left_idx * pv.tip_pitch : nullptr;
  const uint8_t *tipR = (rk == NRX_TIP) ? pv.tipchars + (size_t)op.right_idx * pv.tip_pitch : nullptr;
  double *par = pv.clv[op.parent_slot];
  uint32_t *psc = pv.scaler[op.parent_slot];
  const double w1 = op.w[0], w2 = op.w[1], w3 = op.w[2], w4 = op.w[3];
  const unsigned quad = 0xFu << (lane & ~3);
  const D4 ones{1.0, 1.0, 1.0, 1.0};
  const uint64_t span = ((n_items + BLOCK - 1) / BLOCK) * BLOCK;   // whole warps stay in the loop for the ballots
  for (uint64_t g = (uint64_t)blockIdx.x * BLOCK + tid; g < span; g += (uint64_t)gridDim.x * BLOCK) {
    const bool act = g < n_items;
    const uint64_t site = g >> 2;
    D4 x = ones, y = ones;
    uint32_t sl = 0, sr = 0;
    if (act) {
      if (lk == NRX_CLV) { x = matvec4(PL, ldg256(clvL + g * 4)); sl = scL[site]; }
      else if (lk == NRX_TIP) x = *reinterpret_cast<const D4 *>(lutL + ((tipL[site] & 15) * 4 + cat) * 4);
      if (rk == NRX_CLV) { y = matvec4(PR, ldg256(clvR + g * 4)); sr = scR[site]; }
      else if (rk == NRX_TIP) y = *reinterpret_cast<const D4 *>(lutR + ((tipR[site] & 15) * 4 + cat) * 4);
    }
    D4 m{0.0, 0.0, 0.0, 0.0};
    uint32_t s_out = 0;
    // one libpll update: value c (operands present: pl, pr), its scaling test unless both operands are tips, blend
    auto blend = [&](D4 c, bool tiptip, uint32_t s_in, double w) {
      const bool small = act & (c.x < SCALE_THRESHOLD) & (c.y < SCALE_THRESHOLD) & (c.z < SCALE_THRESHOLD) & (c.w < SCALE_THRESHOLD);
      const unsigned b = __ballot_sync(0xffffffffu, small);
      const bool scale = !tiptip && ((b & quad) == quad);
      if (scale) { c.x = __dmul_rn(c.x, SCALE_FACTOR); c.y = __dmul_rn(c.y, SCALE_FACTOR); c.z = __dmul_rn(c.z, SCALE_FACTOR); c.w = __dmul_rn(c.w, SCALE_FACTOR); }
      m.x = __dadd_rn(m.x, __dmul_rn(w, c.x)); m.y = __dadd_rn(m.y, __dmul_rn(w, c.y));
      m.z = __dadd_rn(m.z, __dmul_rn(w, c.z)); m.w = __dadd_rn(m.w, __dmul_rn(w, c.w));
      s_out = tiptip ? 0u : s_in + (scale ? 1u : 0u);
    };
    if (w1 > 0.0) {   // case 1: take both
      D4 c;
      if (rk == NRX_NONE) c = x; else if (lk == NRX_NONE) c = y;
      else { c.x = __dmul_rn(x.x, y.x); c.y = __dmul_rn(x.y, y.y); c.z = __dmul_rn(x.z, y.z); c.w = __dmul_rn(x.w, y.w); }
      blend(c, lk == NRX_TIP && rk == NRX_TIP, sl + sr, w1);
    }
    if (w2 > 0.0) blend(x, false, sl, w2);   // case 2: left child only (right = fake inner CLV)
    if (w3 > 0.0) blend(y, false, sr, w3);   // case 3: right child only
    if (w4 > 0.0) { m.x = __dadd_rn(m.x, w4); m.y = __dadd_rn(m.y, w4); m.z = __dadd_rn(m.z, w4); m.w = __dadd_rn(m.w, w4); }
    if (act) {
      stg256(par + g * 4, m);
      if (cat == 0 && (w1 > 0.0 || w2 > 0.0 || w3 > 0.0)) psc[site] = s_out;
    }
  }
}

/* any state / category count: one thread per pattern, two passes (scaling decisions of the three updates, then blend) */
__global__ void __launch_bounds__(BLOCK) k_clv_pseudo_generic(const PartView *__restrict__ parts, const nrx_pseudo_op *__restrict__ ops) {
  const PartView &pv = parts[blockIdx.z];
  const nrx_pseudo_op op = ops[blockIdx.y];
  const uint32_t S = pv.states, SP = pv.sp, C = pv.cats;
  const int lk = op.left_kind, rk = op.right_kind;
  const double *pmL = pv.pmat + (size_t)op.left_edge * C * S * SP;
  const double *pmR = pv.pmat + (size_t)op.right_edge * C * S * SP;
  double *par = pv.clv[op.parent_slot];
  uint32_t *psc = pv.scaler[op.parent_slot];
  const double w1 = op.w[0], w2 = op.w[1], w3 = op.w[2], w4 = op.w[3];
  const bool tiptip = (lk == NRX_TIP && rk == NRX_TIP);
  for (uint64_t n = (uint64_t)blockIdx.x * BLOCK + threadIdx.x; n < pv.patterns; n += (uint64_t)gridDim.x * BLOCK) {
    const uint32_t mL = (lk == NRX_TIP) ? pv.tipmap[pv.tipchars[(size_t)op.left_idx * pv.tip_pitch + n]] : 0;
    const uint32_t mR = (rk == NRX_TIP) ? pv.tipmap[pv.tipchars[(size_t)op.right_idx * pv.tip_pitch + n]] : 0;
    const double *cl = (lk == NRX_CLV) ? pv.clv[op.left_idx] + n * C * SP : nullptr;
    const double *cr = (rk == NRX_CLV) ? pv.clv[op.right_idx] + n * C * SP : nullptr;
    double *out = par + n * C * SP;
    auto xy = [&](uint32_t c, uint32_t i, double &x, double &y) {
      const double *lrow = pmL + ((size_t)c * S + i) * SP, *rrow = pmR + ((size_t)c * S + i) * SP;
      x = 1.0; y = 1.0;
      if (lk == NRX_CLV) x = row_dot(lrow, cl + c * SP, S); else if (lk == NRX_TIP) x = masked_rowsum(lrow, S, mL);
      if (rk == NRX_CLV) y = row_dot(rrow, cr + c * SP, S); else if (rk == NRX_TIP) y = masked_rowsum(rrow, S, mR);
    };
    bool small1 = true, small2 = true, small3 = true;
    for (uint32_t c = 0; c < C; ++c)
      for (uint32_t i = 0; i < S; ++i) {
        double x, y;
        xy(c, i, x, y);
        small1 &= (__dmul_rn(x, y) < SCALE_THRESHOLD);
        small2 &= (x < SCALE_THRESHOLD);
        small3 &= (y < SCALE_THRESHOLD);
      }
    const bool sc1 = !tiptip && small1, sc2 = small2, sc3 = small3;
    for (uint32_t c = 0; c < C; ++c) {
      for (uint32_t i = 0; i < S; ++i) {
        double x, y;
        xy(c, i, x, y);
        double m = 0.0;
        if (w1 > 0.0) { double v = __dmul_rn(x, y); if (sc1) v = __dmul_rn(v, SCALE_FACTOR); m = __dadd_rn(m, __dmul_rn(w1, v)); }
        if (w2 > 0.0) { double v = x; if (sc2) v = __dmul_rn(v, SCALE_FACTOR); m = __dadd_rn(m, __dmul_rn(w2, v)); }
        if (w3 > 0.0) { double v = y; if (sc3) v = __dmul_rn(v, SCALE_FACTOR); m = __dadd_rn(m, __dmul_rn(w3, v)); }
        if (w4 > 0.0) m = __dadd_rn(m, w4);
        out[c * SP + i] = m;
      }
      for (uint32_t i = S; i < SP; ++i) out[c * SP + i] = (w4 > 0.0) ? w4 : 0.0;   // padding of the scratch CLVs is 0, of the fake CLV 1
    }
    const uint32_t sl = (lk == NRX_CLV) ? pv.scaler[op.left_idx][n] : 0u, sr = (rk == NRX_CLV) ? pv.scaler[op.right_idx][n] : 0u;
    if (w3 > 0.0) psc[n] = sr + (sc3 ? 1u : 0u);
    else if (w2 > 0.0) psc[n] = sl + (sc2 ? 1u : 0u);
    else if (w1 > 0.0) psc[n] = tiptip ? 0u : sl + sr + (sc1 ? 1u : 0u);
  }
}

/* ------------------------------------------------------------------------------------------------
 * Block-level deterministic sum of up to 3 values; result valid in thread 0.
 * ---------------------------------------------------------------------------------------------- */
template <int N>
__device__ __forceinline__ void block_sum(double (&v)[N], double *smem /* [N][BLOCK/32] */) {
#pragma unroll
  for (int k = 0; k < N; ++k) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v[k] += __shfl_down_sync(0xffffffffu, v[k], off);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0)
    for (int k = 0; k < N; ++k) smem[k * (BLOCK / 32) + warp] = v[k];
  __syncthreads();
  if (threadIdx.x == 0)
    for (int k = 0; k < N; ++k) {
      double s = 0.0;
      for (int w = 0; w < BLOCK / 32; ++w) s += smem[k * (BLOCK / 32) + w];
      v[k] = s;
    }
}

/* second stage: out[item][part][k] = sum over blocks of partial[((item*nparts+part)*N + k)*nblk + b].  One WARP per
 * output: lane l accumulates b = l, l+32, ... in order, then a fixed shuffle tree — deterministic for a given nblk, and
 * 32 independent load chains instead of one thread walking nblk (up to 4736) values serially (was 20 us per call). */
__global__ void __launch_bounds__(128) k_reduce_partials(const double *__restrict__ partial, double *__restrict__ out, uint32_t nblk, uint32_t total) {
  const uint32_t i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= total) return;
  const double *p = partial + (size_t)i * nblk;
  double s = 0.0;
  for (uint32_t b = lane; b < nblk; b += 32) s += p[b];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
  if (lane == 0) out[i] = s;
}

/* The same second stage FUSED into the reducing kernels (round 2: the derivative sweep spent 550 of its 3089 launches in
 * k_reduce_partials): a block that has written its partial sums takes a ticket for its (item, partition) output; the block that
 * draws the last ticket sums all nblk partials — same lane-strided order + shuffle tree as k_reduce_partials, so the result is
 * bit-identical to the two-launch form and independent of WHICH block finishes last — writes out[] and re-arms the counter.
 * `partial` = this output's [N][nblk] block of partial sums, `out` = its N results.  counter == nullptr: two-launch form. */
template <int N>
__device__ __forceinline__ void finish_partials(const double *partial, double *out, uint32_t *counter, uint32_t nblk) {
  if (counter == nullptr) return;
  __shared__ int s_last;
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = (atomicAdd(counter, 1u) == nblk - 1u) ? 1 : 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < (uint32_t)N) {
    const double *p = partial + (size_t)warp * nblk;
    double s = 0.0;
    for (uint32_t b = lane; b < nblk; b += 32) s += __ldcg(p + b);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
    if (lane == 0) out[warp] = s;
  }
  if (threadIdx.x == 0) *counter = 0u;
}

/* ------------------------------------------------------------------------------------------------
 * K3  root lnL (LIBPLL/core_likelihood.c:25-209, 4x4: core_likelihood_avx.c:206-282).
 * K4  edge lnL (LIBPLL/core_likelihood.c:1191-1496 ii, :351-922 ti).
 * Generic over states/cats: one thread per pattern (these kernels read each CLV once; the log and the
 * reduction dominate).  grid = (tiles, items, partitions), partial sums per block, fixed-order stage 2.
 * ---------------------------------------------------------------------------------------------- */
__global__ void __launch_bounds__(BLOCK) k_tree_lnl(const PartView *__restrict__ parts, const uint32_t *__restrict__ slots,
                                                     double *__restrict__ partial, uint32_t nparts_total, double log_thresh,
                                                     double *__restrict__ persite, size_t persite_stride,
                                                     double *__restrict__ out, uint32_t *__restrict__ counters) {
  __shared__ double red[BLOCK / 32];
  const PartView &pv = parts[blockIdx.z];
  const uint32_t S = pv.states, SP = pv.sp, C = pv.cats;
  const uint32_t slot = slots[blockIdx.y];
  const double *clv = pv.clv[slot];
  const uint32_t *sc = pv.scaler[slot];
  double acc[1] = {0.0};
  for (uint64_t n = (uint64_t)blockIdx.x * BLOCK + threadIdx.x; n < pv.patterns; n += (uint64_t)gridDim.x * BLOCK) {
    const double *c = clv + n * C * SP;
    double term = 0.0;
    const double pinv = pv.pinv;
    const int iv = pinv > 0.0 ? pv.invariant[n] : -1;
    for (uint32_t j = 0; j < C; ++j) {
      const double *fr = pv.freqs + cat_model_of(pv, j) * SP;   // freqs_indices[j] (core_likelihood.c:150-160)
      double term_r;
      if (S == 4) {
        const D4 v = ldg256(c + j * 4);
        term_r = tree4(__dmul_rn(fr[0], v.x), __dmul_rn(fr[1], v.y), __dmul_rn(fr[2], v.z), __dmul_rn(fr[3], v.w));
      } else {
        term_r = 0.0;
        for (uint32_t k = 0; k < S; ++k) term_r = __dadd_rn(term_r, __dmul_rn(c[j * SP + k], fr[k]));
      }
      term = __dadd_rn(term, root_cat_term(term_r, pv.rate_weights[j], pinv, iv < 0 ? 0.0 : fr[iv]));
    }
    double lk = log(term);
    const uint32_t s = sc[n];
    if (s) lk = __dadd_rn(lk, __dmul_rn((double)s, log_thresh));
    lk = __dmul_rn(lk, (double)pv.weights[n]);
    if (persite) persite[((size_t)blockIdx.y * nparts_total + pv.part_index) * persite_stride + n] = lk;
    acc[0] += lk;
  }
  block_sum<1>(acc, red);
  const size_t oi = (size_t)blockIdx.y * nparts_total + pv.part_index;
  if (threadIdx.x == 0) partial[oi * gridDim.x + blockIdx.x] = acc[0];
  finish_partials<1>(partial + oi * gridDim.x, out + oi, counters ? counters + oi : nullptr, gridDim.x);
}

__global__ void __launch_bounds__(BLOCK) k_edge_lnl(const PartView *__restrict__ parts, const nrx_pair *__restrict__ pairs,
                                                     uint32_t edge, double *__restrict__ partial, uint32_t nparts_total,
                                                     double log_thresh, double *__restrict__ out, uint32_t *__restrict__ counters) {
  __shared__ double red[BLOCK / 32];
  const PartView &pv = parts[blockIdx.z];
  const uint32_t S = pv.states, SP = pv.sp, C = pv.cats;
  nrx_pair pr = pairs[blockIdx.y];
  // the tip, if any, plays "child" (LIBPLL/likelihood.c:586-601)
  if (pr.a_kind == NRX_TIP) { nrx_pair t = pr; pr.a_kind = t.b_kind; pr.a_idx = t.b_idx; pr.b_kind = t.a_kind; pr.b_idx = t.a_idx; }
  const double *clvp = pv.clv[pr.a_idx];
  const uint32_t *scp = pv.scaler[pr.a_idx];
  const double *clvc = (pr.b_kind == NRX_CLV) ? pv.clv[pr.b_idx] : nullptr;
  const uint32_t *scc = (pr.b_kind == NRX_CLV) ? pv.scaler[pr.b_idx] : nullptr;
  const uint8_t *tip = (pr.b_kind == NRX_TIP) ? pv.tipchars + (size_t)pr.b_idx * pv.tip_pitch : nullptr;
  const double *pm = pv.pmat + (size_t)edge * C * S * SP;
  double acc[1] = {0.0};
  for (uint64_t n = (uint64_t)blockIdx.x * BLOCK + threadIdx.x; n < pv.patterns; n += (uint64_t)gridDim.x * BLOCK) {
    const uint32_t mask = tip ? pv.tipmap[tip[n]] : 0;
    double terma = 0.0, terminv = 0.0;
    const double pinv = pv.pinv;
    const int iv = pinv > 0.0 ? pv.invariant[n] : -1;
    for (uint32_t i = 0; i < C; ++i) {
      const double *fr = pv.freqs + cat_model_of(pv, i) * SP;   // freqs_indices[i] (core_likelihood.c:1236-1240)
      const double *cp = clvp + (n * C + i) * SP;
      const double *cc = clvc ? clvc + (n * C + i) * SP : nullptr;
      double terma_r = 0.0;
      for (uint32_t j = 0; j < S; ++j) {
        const double *row = pm + ((size_t)i * S + j) * SP;
        const double termb = tip ? masked_rowsum(row, S, mask) : row_dot(row, cc, S);
        terma_r = __dadd_rn(terma_r, __dmul_rn(__dmul_rn(cp[j], fr[j]), termb));
      }
      edge_cat_accum(terma_r, pv.rate_weights[i], pinv, iv < 0 ? 0.0 : fr[iv], iv >= 0, terma, terminv);
    }
    const uint32_t s = scp[n] + (scc ? scc[n] : 0u);
    double lk;
    if (pinv > 0.0) lk = edge_site_lnl(terma, terminv, s, log_thresh);
    else { lk = log(terma); if (s) lk = __dadd_rn(lk, __dmul_rn((double)s, log_thresh)); }
    acc[0] += __dmul_rn(lk, (double)pv.weights[n]);
  }
  block_sum<1>(acc, red);
  const size_t oi = (size_t)blockIdx.y * nparts_total + pv.part_index;
  if (threadIdx.x == 0) partial[oi * gridDim.x + blockIdx.x] = acc[0];
  finish_partials<1>(partial + oi * gridDim.x, out + oi, counters ? counters + oi : nullptr, gridDim.x);
}

/* ------------------------------------------------------------------------------------------------
 * K5  sumtable (LIBPLL/core_derivatives.c:321-471 ii, :473-641 ti): thread per (pattern, category).
 * ---------------------------------------------------------------------------------------------- */
__global__ void __launch_bounds__(BLOCK) k_sumtable(const PartView *__restrict__ parts, const nrx_pair *__restrict__ pairs) {
  const PartView &pv = parts[blockIdx.z];
  const uint32_t S = pv.states, SP = pv.sp, C = pv.cats;
  nrx_pair pr = pairs[blockIdx.y];
  // tip-inner: the TIP is the "left" operand of the core kernel (LIBPLL/derivatives.c:70-98)
  if (pr.b_kind == NRX_TIP) { nrx_pair t = pr; pr.a_kind = t.b_kind; pr.a_idx = t.b_idx; pr.b_kind = t.a_kind; pr.b_idx = t.a_idx; }
  const double *clvl = (pr.a_kind == NRX_CLV) ? pv.clv[pr.a_idx] : nullptr;
  const uint8_t *tip = (pr.a_kind == NRX_TIP) ? pv.tipchars + (size_t)pr.a_idx * pv.tip_pitch : nullptr;
  const double *clvr = pv.clv[pr.b_idx];
  double *out = pv.sumtable[blockIdx.y];
  const uint64_t n_items = (uint64_t)pv.patterns * C;
  for (uint64_t g = (uint64_t)blockIdx.x * BLOCK + threadIdx.x; g < n_items; g += (uint64_t)gridDim.x * BLOCK) {
    const uint64_t n = g / C;
    const uint32_t cm = cat_model_of(pv, (uint32_t)(g % C));   // params_indices[i] (core_derivatives.c:362-366)
    const double *fr = pv.freqs + cm * SP, *iev = pv.inv_eigenvecs + (size_t)cm * S * SP, *ev = pv.eigenvecs + (size_t)cm * S * SP;
    const double *cr = clvr + g * SP;
    const double *cl = clvl ? clvl + g * SP : nullptr;
    const uint32_t mask = tip ? pv.tipmap[tip[n]] : 0;
    for (uint32_t j = 0; j < S; ++j) {
      double lefterm = 0.0, righterm = 0.0;
      for (uint32_t k = 0; k < S; ++k) {
        const double lv = tip ? (double)((mask >> k) & 1u) : cl[k];
        lefterm = __dadd_rn(lefterm, __dmul_rn(__dmul_rn(lv, fr[k]), iev[k * SP + j]));
        righterm = __dadd_rn(righterm, __dmul_rn(ev[j * SP + k], cr[k]));
      }
      out[g * SP + j] = __dmul_rn(lefterm, righterm);
    }
    for (uint32_t j = S; j < SP; ++j) out[g * SP + j] = 0.0;
  }
}

/* ------------------------------------------------------------------------------------------------
 * K6  first/second derivative + f per sumtable (LIBPLL/core_derivatives.c:643-694,840-867;
 * f as in core_derivatives_avx2.c:1788-1874: no scaler term).  thread per pattern.
 * ---------------------------------------------------------------------------------------------- */
__global__ void __launch_bounds__(BLOCK) k_derivatives(const PartView *__restrict__ parts, double *__restrict__ partial,
                                                        uint32_t nparts_total, double *__restrict__ out, uint32_t *__restrict__ counters) {
  __shared__ double red[3 * (BLOCK / 32)];
  const PartView &pv = parts[blockIdx.z];
  const uint32_t S = pv.states, SP = pv.sp, C = pv.cats;
  const double *st = pv.sumtable[blockIdx.y];
  double acc[3] = {0.0, 0.0, 0.0};
  for (uint64_t n = (uint64_t)blockIdx.x * BLOCK + threadIdx.x; n < pv.patterns; n += (uint64_t)gridDim.x * BLOCK) {
    const double *sum = st + n * C * SP;
    double lk0 = 0.0, lk1 = 0.0, lk2 = 0.0;
    const double pinv = pv.pinv;
    const int iv = pinv > 0.0 ? pv.invariant[n] : -1;
    for (uint32_t i = 0; i < C; ++i) {
      const double invf = iv < 0 ? 0.0 : pv.freqs[cat_model_of(pv, i) * SP + iv];   // freqs[params_indices[i]] (core_derivatives.c:676-686)
      double c0 = 0.0, c1 = 0.0, c2 = 0.0;
      const double *dg = pv.diagp + (size_t)i * S * 4;
      for (uint32_t j = 0; j < S; ++j) {
        const double v = sum[i * SP + j];
        c0 = __dadd_rn(c0, __dmul_rn(v, dg[j * 4 + 0]));
        c1 = __dadd_rn(c1, __dmul_rn(v, dg[j * 4 + 1]));
        c2 = __dadd_rn(c2, __dmul_rn(v, dg[j * 4 + 2]));
      }
      const double w = pv.rate_weights[i];
      deriv_cat_pinv(c0, c1, c2, pinv, invf);
      lk0 = __dadd_rn(lk0, __dmul_rn(c0, w));
      lk1 = __dadd_rn(lk1, __dmul_rn(c1, w));
      lk2 = __dadd_rn(lk2, __dmul_rn(c2, w));
    }
    const double pw = (double)pv.weights[n];
    const double d1 = -lk1 / lk0;
    const double d2 = d1 * d1 - lk2 / lk0;
    acc[0] += pw * log(lk0);
    acc[1] += pw * d1;
    acc[2] += pw * d2;
  }
  block_sum<3>(acc, red);
  const size_t oi = (size_t)blockIdx.y * nparts_total + pv.part_index;
  double *p = partial + oi * 3 * gridDim.x;
  if (threadIdx.x == 0) {
    p[0 * gridDim.x + blockIdx.x] = acc[0];
    p[1 * gridDim.x + blockIdx.x] = acc[1];
    p[2 * gridDim.x + blockIdx.x] = acc[2];
  }
  finish_partials<3>(p, out + oi * 3, counters ? counters + oi : nullptr, gridDim.x);
}

/* Slot copies of the virtual re-rooting save/restore (the reference copy-assigns NodeDisplayedTreeData, i.e. memcpy's
 * every CLV of a node on the host, LH/VirtualRerooting.cpp:211-220,234-238): all (dst, src) pairs x partitions in ONE
 * launch instead of two cudaMemcpyAsync per slot and partition.  grid = (chunks, pairs, partitions of this shape). */
__global__ void __launch_bounds__(BLOCK) k_copy_slots(const PartView *__restrict__ parts, const uint2 *__restrict__ dst_src) {
  const PartView &pv = parts[blockIdx.z];
  const uint2 ds = dst_src[blockIdx.y];
  const uint64_t n4 = (uint64_t)pv.patterns * pv.cats * pv.sp / 4;   // sp is a multiple of 4: 32-byte units
  const double *src = pv.clv[ds.y];
  double *dst = pv.clv[ds.x];
  for (uint64_t i = (uint64_t)blockIdx.x * BLOCK + threadIdx.x; i < n4; i += (uint64_t)gridDim.x * BLOCK) stg256(dst + i * 4, ldg256(src + i * 4));
  const uint32_t *ssrc = pv.scaler[ds.y];
  uint32_t *sdst = pv.scaler[ds.x];
  for (uint64_t i = (uint64_t)blockIdx.x * BLOCK + threadIdx.x; i < pv.patterns; i += (uint64_t)gridDim.x * BLOCK) sdst[i] = ssrc[i];
}

/* ------------------------------------------------------------------------------------------------
 * K3 / K6 for any state count when the category count is a power of two (the 20-state case): thread = one
 * (pattern, category) item = `sp` contiguous doubles read as 256-bit loads (a warp covers 32 consecutive items =
 * one contiguous span), the per-category results of a pattern are gathered IN CATEGORY ORDER from the adjacent
 * lanes (same summation order as the thread-per-pattern kernels above and as the reference), constants in shared
 * memory.  4x the threads and 1/4 of the load instructions of the thread-per-pattern version.  (A warp-per-pattern
 * variant of K6 with fully contiguous 640-byte warp loads and 20 active lanes was measured SLOWER: 0.22 vs 0.34 of the
 * HBM roofline at 200 k protein patterns — one load in flight per lane and a 27-shuffle chain per pattern.)
 * ---------------------------------------------------------------------------------------------- */
template <int SC /* compile-time state count (loops unroll, all loads issue up front); 0 = run time */>
__global__ void __launch_bounds__(BLOCK) k_tree_lnl_pc(const PartView *__restrict__ parts, const uint32_t *__restrict__ slots,
                                                        double *__restrict__ partial, uint32_t nparts_total, double log_thresh,
                                                        double *__restrict__ persite, size_t persite_stride,
                                                        double *__restrict__ out, uint32_t *__restrict__ counters) {
  __shared__ double red[BLOCK / 32];
  __shared__ double sfreq[32], swt[32];
  const PartView &pv = parts[blockIdx.z];
  const uint32_t S = SC ? (uint32_t)SC : pv.states, SP = SC ? (uint32_t)((SC + 3) & ~3) : pv.sp, C = pv.cats;
  const uint32_t slot = slots[blockIdx.y];
  const double *clv = pv.clv[slot];
  const uint32_t *sc = pv.scaler[slot];
  if (threadIdx.x < S) sfreq[threadIdx.x] = pv.freqs[threadIdx.x];
  if (threadIdx.x < C) swt[threadIdx.x] = pv.rate_weights[threadIdx.x];
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31, c = threadIdx.x & (C - 1);
  const uint64_t n_items = (uint64_t)pv.patterns * C;
  const uint64_t span = ((n_items + BLOCK - 1) / BLOCK) * BLOCK;  // whole warps stay in the loop for the shuffles
  const double pinv = pv.pinv;
  double acc[1] = {0.0};
  for (uint64_t g = (uint64_t)blockIdx.x * BLOCK + threadIdx.x; g < span; g += (uint64_t)gridDim.x * BLOCK) {
    double t = 0.0;
    if (g < n_items) {
      const double *v = clv + g * SP;
      double term_r = 0.0;
#pragma unroll
      for (uint32_t k = 0; k < SP; k += 4) {
        const D4 q = ldg256(v + k);
        term_r = __dadd_rn(term_r, __dmul_rn(q.x, sfreq[k]));
        if (k + 1 < S) term_r = __dadd_rn(term_r, __dmul_rn(q.y, sfreq[k + 1]));
        if (k + 2 < S) term_r = __dadd_rn(term_r, __dmul_rn(q.z, sfreq[k + 2]));
        if (k + 3 < S) term_r = __dadd_rn(term_r, __dmul_rn(q.w, sfreq[k + 3]));
      }
      double invf = 0.0;
      if (pinv > 0.0) { const int iv = pv.invariant[g / C]; invf = iv < 0 ? 0.0 : sfreq[iv]; }
      t = root_cat_term(term_r, swt[c], pinv, invf);
    }
    double term = 0.0;
    for (uint32_t i = 0; i < C; ++i) term = __dadd_rn(term, __shfl_sync(0xffffffffu, t, (lane & ~(C - 1)) + i));
    if (c == 0 && g < n_items) {
      const uint64_t n = g / C;
      double lk = log(term);
      const uint32_t s = sc[n];
      if (s) lk = __dadd_rn(lk, __dmul_rn((double)s, log_thresh));
      lk = __dmul_rn(lk, (double)pv.weights[n]);
      if (persite) persite[((size_t)blockIdx.y * nparts_total + pv.part_index) * persite_stride + n] = lk;
      acc[0] += lk;
    }
  }
  block_sum<1>(acc, red);
  const size_t oi = (size_t)blockIdx.y * nparts_total + pv.part_index;
  if (threadIdx.x == 0) partial[oi * gridDim.x + blockIdx.x] = acc[0];
  finish_partials<1>(partial + oi * gridDim.x, out + oi, counters ? counters + oi : nullptr, gridDim.x);
}

/* K6, thread = (pattern, category), NP patterns per thread.  ncu on the first version (profiles/r2a_k6_protein_200k.md): 54 % of the
 * warp stalls on the shared-memory scoreboard + 17 % MIO throttle, 36 M excessive shared wavefronts per launch — the diag table
 * rows of the four categories of a warp lay 640 B apart, i.e. in the SAME banks (4-way conflict on each of the 40 LDS per item),
 * and one item per thread kept only 5 x 32 B of HBM loads in flight.  Now: the per-category stride of the table is padded to
 * = 2 (mod 16) doubles so the four (broadcast) double2 reads of a warp fall into disjoint bank groups, and every diag entry read
 * from shared memory is applied to NP patterns (NP x 5 independent 256-bit loads in flight per thread).  Per-(pattern,
 * category) arithmetic and summation order are unchanged. */
__host__ __device__ constexpr uint32_t diag_stride(uint32_t S) { return ((S * 4 + 13) / 16) * 16 + 2; }

template <int SC, int NP>
__global__ void __launch_bounds__(BLOCK) k_derivatives_pc(const PartView *__restrict__ parts, double *__restrict__ partial,
                                                           uint32_t nparts_total, double *__restrict__ out, uint32_t *__restrict__ counters) {
  __shared__ double red[3 * (BLOCK / 32)];
  extern __shared__ __align__(16) double sdiag[];  // [cats][diag_stride(states)] (entries [state][4]) + [cats] rate weights
  const PartView &pv = parts[blockIdx.z];
  const uint32_t S = SC ? (uint32_t)SC : pv.states, SP = SC ? (uint32_t)((SC + 3) & ~3) : pv.sp, C = pv.cats;
  const uint32_t DS = diag_stride(S);
  double *swt = sdiag + (size_t)C * DS;
  for (uint32_t i = threadIdx.x; i < C * S * 4; i += BLOCK) sdiag[(i / (S * 4)) * DS + i % (S * 4)] = pv.diagp[i];
  if (threadIdx.x < C) swt[threadIdx.x] = pv.rate_weights[threadIdx.x];
  __syncthreads();
  const double *st = pv.sumtable[blockIdx.y];
  const uint32_t lane = threadIdx.x & 31, c = threadIdx.x & (C - 1);
  const double *dg = sdiag + (size_t)c * DS;
  const uint64_t n_items = (uint64_t)pv.patterns * C;
  const uint64_t span = ((n_items + (uint64_t)BLOCK * NP - 1) / ((uint64_t)BLOCK * NP)) * ((uint64_t)BLOCK * NP);   // whole warps stay in the loop for the shuffles
  const double pinv = pv.pinv;
  const double w = swt[c];
  double acc[3] = {0.0, 0.0, 0.0};
  for (uint64_t base = (uint64_t)blockIdx.x * BLOCK * NP; base < span; base += (uint64_t)gridDim.x * BLOCK * NP) {
    double c0[NP], c1[NP], c2[NP];
    const double *v[NP];
    bool on[NP];
#pragma unroll
    for (int u = 0; u < NP; ++u) {
      const uint64_t g = base + (uint64_t)u * BLOCK + threadIdx.x;   // BLOCK is a multiple of C: the category is the same for every u
      on[u] = g < n_items;
      v[u] = st + (on[u] ? g : 0) * SP;
      c0[u] = c1[u] = c2[u] = 0.0;
    }
#pragma unroll
    for (uint32_t k = 0; k < SP; k += 4) {
      D4 q[NP];
#pragma unroll
      for (int u = 0; u < NP; ++u) q[u] = ldg256(v[u] + k);
#pragma unroll
      for (uint32_t h = 0; h < 4; ++h)
        if (k + h < S) {
          const double2 d01 = *reinterpret_cast<const double2 *>(dg + (k + h) * 4);
          const double d2 = dg[(k + h) * 4 + 2];
#pragma unroll
          for (int u = 0; u < NP; ++u) {
            const double e = h == 0 ? q[u].x : (h == 1 ? q[u].y : (h == 2 ? q[u].z : q[u].w));
            c0[u] = __dadd_rn(c0[u], __dmul_rn(e, d01.x));
            c1[u] = __dadd_rn(c1[u], __dmul_rn(e, d01.y));
            c2[u] = __dadd_rn(c2[u], __dmul_rn(e, d2));
          }
        }
    }
#pragma unroll
    for (int u = 0; u < NP; ++u) {
      const uint64_t g = base + (uint64_t)u * BLOCK + threadIdx.x;
      double t0 = 0.0, t1 = 0.0, t2 = 0.0;
      if (on[u]) {
        if (pinv > 0.0) { const int iv = pv.invariant[g / C]; deriv_cat_pinv(c0[u], c1[u], c2[u], pinv, iv < 0 ? 0.0 : pv.freqs[iv]); }
        t0 = __dmul_rn(c0[u], w); t1 = __dmul_rn(c1[u], w); t2 = __dmul_rn(c2[u], w);
      }
      double lk0 = 0.0, lk1 = 0.0, lk2 = 0.0;
      for (uint32_t i = 0; i < C; ++i) {
        const int src = (int)((lane & ~(C - 1)) + i);
        lk0 = __dadd_rn(lk0, __shfl_sync(0xffffffffu, t0, src));
        lk1 = __dadd_rn(lk1, __shfl_sync(0xffffffffu, t1, src));
        lk2 = __dadd_rn(lk2, __shfl_sync(0xffffffffu, t2, src));
      }
      if (c == 0 && on[u]) {
        const double pw = (double)pv.weights[g / C];
        const double d1 = -lk1 / lk0;
        const double d2 = d1 * d1 - lk2 / lk0;
        acc[0] += pw * log(lk0);
        acc[1] += pw * d1;
        acc[2] += pw * d2;
      }
    }
  }
  block_sum<3>(acc, red);
  const size_t oi = (size_t)blockIdx.y * nparts_total + pv.part_index;
  double *p = partial + oi * 3 * gridDim.x;
  if (threadIdx.x == 0) {
    p[0 * gridDim.x + blockIdx.x] = acc[0];
    p[1 * gridDim.x + blockIdx.x] = acc[1];
    p[2 * gridDim.x + blockIdx.x] = acc[2];
  }
  finish_partials<3>(p, out + oi * 3, counters ? counters + oi : nullptr, gridDim.x);
}

/* K3 / K6 for 20 states x 4 categories, pipelined (round 2): same thread = (pattern, category) mapping, arithmetic and category
 * gather as k_tree_lnl_pc<20> / k_derivatives_pc<20, NP>, but few long-lived blocks (reduce_blocks' quad geometry) that load the
 * NEXT pass's five 256-bit chunks into registers before the current pass is consumed — the restructuring that took the DNA K6
 * from 0.49 to 0.76 of the HBM peak (profiles/r3d_k4_k5_k6_dna_sweep_ncu.md: one pass per block = load, log / division tail and
 * block reduction strictly one after the other). */
constexpr int AA_CH = 5;   // 256-bit chunks per (pattern, category) item: 20 states

__device__ __forceinline__ void aa_item_load(D4 (&v)[AA_CH], const double *__restrict__ src, uint64_t g, uint64_t n_items) {
  if (g < n_items) {
#pragma unroll
    for (int k = 0; k < AA_CH; ++k) v[k] = ldg256(src + g * 20 + k * 4);
  } else {
#pragma unroll
    for (int k = 0; k < AA_CH; ++k) { v[k].x = v[k].y = v[k].z = v[k].w = 0.0; }
  }
}

__global__ void __launch_bounds__(BLOCK, 2) k_derivatives_aa20p(const PartView *__restrict__ parts, double *__restrict__ partial,
                                                                 uint32_t nparts_total, double *__restrict__ out, uint32_t *__restrict__ counters) {
  __shared__ double red[3 * (BLOCK / 32)];
  extern __shared__ __align__(16) double sdiag[];  // [4][diag_stride(20)] (entries [state][4]) + [4] rate weights
  const PartView &pv = parts[blockIdx.z];
  constexpr uint32_t S = 20, C = 4;
  const uint32_t DS = diag_stride(S);
  double *swt = sdiag + (size_t)C * DS;
  for (uint32_t i = threadIdx.x; i < C * S * 4; i += BLOCK) sdiag[(i / (S * 4)) * DS + i % (S * 4)] = pv.diagp[i];
  if (threadIdx.x < C) swt[threadIdx.x] = pv.rate_weights[threadIdx.x];
  __syncthreads();
  const double *st = pv.sumtable[blockIdx.y];
  const uint32_t lane = threadIdx.x & 31, c = threadIdx.x & (C - 1);
  const double *dg = sdiag + (size_t)c * DS;
  const uint64_t n_items = (uint64_t)pv.patterns * C, stride = (uint64_t)gridDim.x * BLOCK;
  const double pinv = pv.pinv, w = swt[c];
  double acc[3] = {0.0, 0.0, 0.0};
  uint64_t base = (uint64_t)blockIdx.x * BLOCK;
  D4 nxt[AA_CH];
  aa_item_load(nxt, st, base + threadIdx.x, n_items);
  for (; base < n_items; base += stride) {
    D4 q[AA_CH];
#pragma unroll
    for (int k = 0; k < AA_CH; ++k) q[k] = nxt[k];
    const uint64_t g = base + threadIdx.x;
    if (base + stride < n_items) aa_item_load(nxt, st, g + stride, n_items);
    double c0 = 0.0, c1 = 0.0, c2 = 0.0;
#pragma unroll
    for (int k = 0; k < AA_CH; ++k) {
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        const double2 d01 = *reinterpret_cast<const double2 *>(dg + (k * 4 + h) * 4);
        const double d2 = dg[(k * 4 + h) * 4 + 2];
        const double e = h == 0 ? q[k].x : (h == 1 ? q[k].y : (h == 2 ? q[k].z : q[k].w));
        c0 = __dadd_rn(c0, __dmul_rn(e, d01.x));
        c1 = __dadd_rn(c1, __dmul_rn(e, d01.y));
        c2 = __dadd_rn(c2, __dmul_rn(e, d2));
      }
    }
    const bool on = g < n_items;
    double t0 = 0.0, t1 = 0.0, t2 = 0.0;
    if (on) {
      if (pinv > 0.0) { const int iv = pv.invariant[g / C]; deriv_cat_pinv(c0, c1, c2, pinv, iv < 0 ? 0.0 : pv.freqs[iv]); }
      t0 = __dmul_rn(c0, w); t1 = __dmul_rn(c1, w); t2 = __dmul_rn(c2, w);
    }
    double lk0 = 0.0, lk1 = 0.0, lk2 = 0.0;
#pragma unroll
    for (uint32_t i = 0; i < C; ++i) {
      const int src = (int)((lane & ~(C - 1)) + i);
      lk0 = __dadd_rn(lk0, __shfl_sync(0xffffffffu, t0, src));
      lk1 = __dadd_rn(lk1, __shfl_sync(0xffffffffu, t1, src));
      lk2 = __dadd_rn(lk2, __shfl_sync(0xffffffffu, t2, src));
    }
    if (c == 0 && on) {
      const double pw = (double)pv.weights[g / C];
      const double d1 = -lk1 / lk0;
      const double d2 = d1 * d1 - lk2 / lk0;
      acc[0] += pw * log(lk0);
      acc[1] += pw * d1;
      acc[2] += pw * d2;
    }
  }
  block_sum<3>(acc, red);
  const size_t oi = (size_t)blockIdx.y * nparts_total + pv.part_index;
  double *p = partial + oi * 3 * gridDim.x;
  if (threadIdx.x == 0) {
    p[0 * gridDim.x + blockIdx.x] = acc[0];
    p[1 * gridDim.x + blockIdx.x] = acc[1];
    p[2 * gridDim.x + blockIdx.x] = acc[2];
  }
  finish_partials<3>(p, out + oi * 3, counters ? counters + oi : nullptr, gridDim.x);
}

__global__ void __launch_bounds__(BLOCK, 2) k_tree_lnl_aa20p(const PartView *__restrict__ parts, const uint32_t *__restrict__ slots,
                                                              double *__restrict__ partial, uint32_t nparts_total, double log_thresh,
                                                              double *__restrict__ persite, size_t persite_stride,
                                                              double *__restrict__ out, uint32_t *__restrict__ counters) {
  __shared__ double red[BLOCK / 32];
  __shared__ double sfreq[32], swt[32];
  const PartView &pv = parts[blockIdx.z];
  constexpr uint32_t S = 20, C = 4;
  const uint32_t slot = slots[blockIdx.y];
  const double *clv = pv.clv[slot];
  const uint32_t *sc = pv.scaler[slot];
  if (threadIdx.x < S) sfreq[threadIdx.x] = pv.freqs[threadIdx.x];
  if (threadIdx.x < C) swt[threadIdx.x] = pv.rate_weights[threadIdx.x];
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31, c = threadIdx.x & (C - 1);
  const uint64_t n_items = (uint64_t)pv.patterns * C, stride = (uint64_t)gridDim.x * BLOCK;
  const double pinv = pv.pinv;
  double acc[1] = {0.0};
  uint64_t base = (uint64_t)blockIdx.x * BLOCK;
  D4 nxt[AA_CH];
  aa_item_load(nxt, clv, base + threadIdx.x, n_items);
  for (; base < n_items; base += stride) {
    D4 q[AA_CH];
#pragma unroll
    for (int k = 0; k < AA_CH; ++k) q[k] = nxt[k];
    const uint64_t g = base + threadIdx.x;
    if (base + stride < n_items) aa_item_load(nxt, clv, g + stride, n_items);
    double t = 0.0;
    if (g < n_items) {
      double term_r = 0.0;
#pragma unroll
      for (int k = 0; k < AA_CH; ++k) {
        term_r = __dadd_rn(term_r, __dmul_rn(q[k].x, sfreq[k * 4]));
        term_r = __dadd_rn(term_r, __dmul_rn(q[k].y, sfreq[k * 4 + 1]));
        term_r = __dadd_rn(term_r, __dmul_rn(q[k].z, sfreq[k * 4 + 2]));
        term_r = __dadd_rn(term_r, __dmul_rn(q[k].w, sfreq[k * 4 + 3]));
      }
      double invf = 0.0;
      if (pinv > 0.0) { const int iv = pv.invariant[g / C]; invf = iv < 0 ? 0.0 : sfreq[iv]; }
      t = root_cat_term(term_r, swt[c], pinv, invf);
    }
    double term = 0.0;
#pragma unroll
    for (uint32_t i = 0; i < C; ++i) term = __dadd_rn(term, __shfl_sync(0xffffffffu, t, (lane & ~(C - 1)) + i));
    if (c == 0 && g < n_items) {
      const uint64_t n = g / C;
      double lk = log(term);
      const uint32_t s = sc[n];
      if (s) lk = __dadd_rn(lk, __dmul_rn((double)s, log_thresh));
      lk = __dmul_rn(lk, (double)pv.weights[n]);
      if (persite) persite[((size_t)blockIdx.y * nparts_total + pv.part_index) * persite_stride + n] = lk;
      acc[0] += lk;
    }
  }
  block_sum<1>(acc, red);
  const size_t oi = (size_t)blockIdx.y * nparts_total + pv.part_index;
  if (threadIdx.x == 0) partial[oi * gridDim.x + blockIdx.x] = acc[0];
  finish_partials<1>(partial + oi * gridDim.x, out + oi, counters ? counters + oi : nullptr, gridDim.x);
}

/* ------------------------------------------------------------------------------------------------
 * K3-K6, DNA 4x4 specialisations.  Same arithmetic as the generic kernels above (same summation order), but
 * thread = (pattern, rate category) like K2: a warp reads 1 KB of contiguous CLV per 256-bit load, UNROLL
 * independent loads are issued before the first use, the four categories of a pattern are combined with quad
 * shuffles, and the small per-partition constants (frequencies, eigenvectors, P rows, diag table) sit in
 * shared memory.  grid = (tiles, items, partitions of this shape); per-block partial sums, fixed-order stage 2.
 * ---------------------------------------------------------------------------------------------- */
constexpr int RU = 4;  // k_sumtable_dna4: independent 256-bit loads in flight per operand per thread

/* K3 / K4 / K6 reduce over patterns and end in a log or a division per PATTERN, so here thread = pattern (all 32
 * lanes do the transcendental part) with the four category blocks of the pattern read as four independent 256-bit
 * loads (every 32-byte sector fetched is fully used). */
__global__ void __launch_bounds__(BLOCK) k_tree_lnl_dna4(const PartView *__restrict__ parts, const uint32_t *__restrict__ slots,
                                                          double *__restrict__ partial, uint32_t nparts_total, double log_thresh,
                                                          double *__restrict__ persite, size_t persite_stride,
                                                          double *__restrict__ out, uint32_t *__restrict__ counters) {
  __shared__ double red[BLOCK / 32];
  const PartView &pv = parts[blockIdx.z];
  const uint32_t slot = slots[blockIdx.y];
  const double *clv = pv.clv[slot];
  const uint32_t *sc = pv.scaler[slot];
  const double f0 = pv.freqs[0], f1 = pv.freqs[1], f2 = pv.freqs[2], f3 = pv.freqs[3];
  const double w0 = pv.rate_weights[0], w1 = pv.rate_weights[1], w2 = pv.rate_weights[2], w3 = pv.rate_weights[3];
  const double pinv = pv.pinv;
  double acc[1] = {0.0};
  for (uint64_t n = (uint64_t)blockIdx.x * BLOCK + threadIdx.x; n < pv.patterns; n += (uint64_t)gridDim.x * BLOCK) {
    const double *c = clv + n * 16;
    const D4 v0 = ldg256(c), v1 = ldg256(c + 4), v2 = ldg256(c + 8), v3 = ldg256(c + 12);
    const uint32_t s = sc[n];
    const double pw = (double)pv.weights[n];
    double invf = 0.0;
    if (pinv > 0.0) { const int iv = pv.invariant[n]; invf = iv < 0 ? 0.0 : pv.freqs[iv]; }
    double term = root_cat_term(tree4(__dmul_rn(f0, v0.x), __dmul_rn(f1, v0.y), __dmul_rn(f2, v0.z), __dmul_rn(f3, v0.w)), w0, pinv, invf);
    term = __dadd_rn(term, root_cat_term(tree4(__dmul_rn(f0, v1.x), __dmul_rn(f1, v1.y), __dmul_rn(f2, v1.z), __dmul_rn(f3, v1.w)), w1, pinv, invf));
    term = __dadd_rn(term, root_cat_term(tree4(__dmul_rn(f0, v2.x), __dmul_rn(f1, v2.y), __dmul_rn(f2, v2.z), __dmul_rn(f3, v2.w)), w2, pinv, invf));
    term = __dadd_rn(term, root_cat_term(tree4(__dmul_rn(f0, v3.x), __dmul_rn(f1, v3.y), __dmul_rn(f2, v3.z), __dmul_rn(f3, v3.w)), w3, pinv, invf));
    double lk = log(term);
    if (s) lk = __dadd_rn(lk, __dmul_rn((double)s, log_thresh));
    lk = __dmul_rn(lk, pw);
    if (persite) persite[((size_t)blockIdx.y * nparts_total + pv.part_index) * persite_stride + n] = lk;
    acc[0] += lk;
  }
  block_sum<1>(acc, red);
  const size_t oi = (size_t)blockIdx.y * nparts_total + pv.part_index;
  if (threadIdx.x == 0) partial[oi * gridDim.x + blockIdx.x] = acc[0];
  finish_partials<1>(partial + oi * gridDim.x, out + oi, counters ? counters + oi : nullptr, gridDim.x);
}

/* second half of the fused K3: log, scaler term and pattern weight on the per-site terms K2 wrote — same traversal and
 * accumulation order as k_tree_lnl_dna4, hence bit-identical partial sums */
constexpr int TU = 4;   // k_term_lnl_sum: patterns per thread and pass
__device__ __forceinline__ void term_load(double (&t)[TU], uint32_t (&s)[TU], uint32_t (&w)[TU], const double *__restrict__ ps,
                                          const uint32_t *__restrict__ sc, const uint32_t *__restrict__ wt, uint64_t base, uint64_t n_pat, int tid) {
#pragma unroll
  for (int u = 0; u < TU; ++u) {
    const uint64_t n = base + (uint64_t)u * BLOCK + tid;
    if (n < n_pat) { t[u] = ps[n]; s[u] = sc[n]; w[u] = wt[n]; }
    else { t[u] = 1.0; s[u] = 0u; w[u] = 0u; }   // log(1) * 0: contributes nothing
  }
}
/* four independent logs per thread and pass, the next pass's 16 bytes per pattern loaded before the current ones are consumed
 * (the one-pattern-per-thread form ran at 0.16-0.25 of the HBM peak: a log is ~40 dependent FP64 instructions) */
__global__ void __launch_bounds__(BLOCK) k_term_lnl_sum(const PartView *__restrict__ parts, const uint32_t *__restrict__ slots,
                                                         const double *__restrict__ terms, size_t persite_stride,
                                                         double *__restrict__ partial, uint32_t nparts_total, double log_thresh,
                                                         double *__restrict__ out, uint32_t *__restrict__ counters) {
  __shared__ double red[BLOCK / 32];
  const PartView &pv = parts[blockIdx.z];
  const double *ps = terms + ((size_t)blockIdx.y * nparts_total + pv.part_index) * persite_stride;
  const uint32_t *sc = pv.scaler[slots[blockIdx.y]];
  const uint32_t *wt = pv.weights;
  const uint64_t n_pat = pv.patterns, stride = (uint64_t)gridDim.x * BLOCK * TU;
  const int tid = threadIdx.x;
  double acc[1] = {0.0};
  uint64_t base = (uint64_t)blockIdx.x * BLOCK * TU;
  double tn[TU]; uint32_t sn[TU], wn[TU];
  term_load(tn, sn, wn, ps, sc, wt, base, n_pat, tid);
  for (; base < n_pat; base += stride) {
    double t[TU]; uint32_t s[TU], w[TU];
#pragma unroll
    for (int u = 0; u < TU; ++u) { t[u] = tn[u]; s[u] = sn[u]; w[u] = wn[u]; }
    if (base + stride < n_pat) term_load(tn, sn, wn, ps, sc, wt, base + stride, n_pat, tid);
#pragma unroll
    for (int u = 0; u < TU; ++u) {
      double lk = log(t[u]);
      if (s[u]) lk = __dadd_rn(lk, __dmul_rn((double)s[u], log_thresh));
      acc[0] += __dmul_rn(lk, (double)w[u]);
    }
  }
  block_sum<1>(acc, red);
  const size_t oi = (size_t)blockIdx.y * nparts_total + pv.part_index;
  if (threadIdx.x == 0) partial[oi * gridDim.x + blockIdx.x] = acc[0];
  finish_partials<1>(partial + oi * gridDim.x, out + oi, counters ? counters + oi : nullptr, gridDim.x);
}

__device__ __forceinline__ double edge_cat_term(const D4 &a, const D4 &y, double f0, double f1, double f2, double f3) {
  double r = __dmul_rn(__dmul_rn(a.x, f0), y.x);
  r = __dadd_rn(r, __dmul_rn(__dmul_rn(a.y, f1), y.y));
  r = __dadd_rn(r, __dmul_rn(__dmul_rn(a.z, f2), y.z));
  return __dadd_rn(r, __dmul_rn(__dmul_rn(a.w, f3), y.w));
}

__global__ void __launch_bounds__(BLOCK) k_edge_lnl_dna4(const PartView *__restrict__ parts, const nrx_pair *__restrict__ pairs,
                                                          uint32_t edge, double *__restrict__ partial, uint32_t nparts_total,
                                                          double log_thresh, double *__restrict__ out, uint32_t *__restrict__ counters) {
  __shared__ double red[BLOCK / 32];
  __shared__ __align__(32) double lut[256];
  __shared__ __align__(16) double sP[64];
  const PartView &pv = parts[blockIdx.z];
  nrx_pair pr = pairs[blockIdx.y];
  if (pr.a_kind == NRX_TIP) { nrx_pair t = pr; pr.a_kind = t.b_kind; pr.a_idx = t.b_idx; pr.b_kind = t.a_kind; pr.b_idx = t.a_idx; }
  const int tid = threadIdx.x;
  const bool tipc = pr.b_kind == NRX_TIP;
  if (tipc) build_tip_lut4(lut, pv.pmat + (size_t)edge * 64, tid);
  else if (tid < 64) sP[tid] = pv.pmat[(size_t)edge * 64 + tid];   // all lanes read the same row: broadcast, no conflicts
  __syncthreads();
  const double *clvp = pv.clv[pr.a_idx];
  const uint32_t *scp = pv.scaler[pr.a_idx];
  const double *clvc = tipc ? nullptr : pv.clv[pr.b_idx];
  const uint32_t *scc = tipc ? nullptr : pv.scaler[pr.b_idx];
  const uint8_t *tip = tipc ? pv.tipchars + (size_t)pr.b_idx * pv.tip_pitch : nullptr;
  const double f0 = pv.freqs[0], f1 = pv.freqs[1], f2 = pv.freqs[2], f3 = pv.freqs[3];
  const double w0 = pv.rate_weights[0], w1 = pv.rate_weights[1], w2 = pv.rate_weights[2], w3 = pv.rate_weights[3];
  const double pinv = pv.pinv;
  double acc[1] = {0.0};
  for (uint64_t n = (uint64_t)blockIdx.x * BLOCK + threadIdx.x; n < pv.patterns; n += (uint64_t)gridDim.x * BLOCK) {
    const double *cp = clvp + n * 16;
    const D4 a0 = ldg256(cp), a1 = ldg256(cp + 4), a2 = ldg256(cp + 8), a3 = ldg256(cp + 12);
    D4 y0, y1, y2, y3;
    uint32_t s = scp[n];
    if (tipc) {
      const double *l = lut + (tip[n] & 15) * 16;
      y0 = *reinterpret_cast<const D4 *>(l); y1 = *reinterpret_cast<const D4 *>(l + 4);
      y2 = *reinterpret_cast<const D4 *>(l + 8); y3 = *reinterpret_cast<const D4 *>(l + 12);
    } else {
      const double *cc = clvc + n * 16;
      const D4 b0 = ldg256(cc), b1 = ldg256(cc + 4), b2 = ldg256(cc + 8), b3 = ldg256(cc + 12);
      s += scc[n];
      y0 = matvec4(sP, b0); y1 = matvec4(sP + 16, b1); y2 = matvec4(sP + 32, b2); y3 = matvec4(sP + 48, b3);
    }
    const double pw = (double)pv.weights[n];
    double lk;
    if (pinv > 0.0) {
      const int iv = pv.invariant[n];
      const double invf = iv < 0 ? 0.0 : pv.freqs[iv];
      double terma = 0.0, terminv = 0.0;
      edge_cat_accum(edge_cat_term(a0, y0, f0, f1, f2, f3), w0, pinv, invf, iv >= 0, terma, terminv);
      edge_cat_accum(edge_cat_term(a1, y1, f0, f1, f2, f3), w1, pinv, invf, iv >= 0, terma, terminv);
      edge_cat_accum(edge_cat_term(a2, y2, f0, f1, f2, f3), w2, pinv, invf, iv >= 0, terma, terminv);
      edge_cat_accum(edge_cat_term(a3, y3, f0, f1, f2, f3), w3, pinv, invf, iv >= 0, terma, terminv);
      lk = edge_site_lnl(terma, terminv, s, log_thresh);
    } else {
      double term = __dmul_rn(edge_cat_term(a0, y0, f0, f1, f2, f3), w0);
      term = __dadd_rn(term, __dmul_rn(edge_cat_term(a1, y1, f0, f1, f2, f3), w1));
      term = __dadd_rn(term, __dmul_rn(edge_cat_term(a2, y2, f0, f1, f2, f3), w2));
      term = __dadd_rn(term, __dmul_rn(edge_cat_term(a3, y3, f0, f1, f2, f3), w3));
      lk = log(term);
      if (s) lk = __dadd_rn(lk, __dmul_rn((double)s, log_thresh));
    }
    acc[0] += __dmul_rn(lk, pw);
  }
  block_sum<1>(acc, red);
  const size_t oi = (size_t)blockIdx.y * nparts_total + pv.part_index;
  if (threadIdx.x == 0) partial[oi * gridDim.x + blockIdx.x] = acc[0];
  finish_partials<1>(partial + oi * gridDim.x, out + oi, counters ? counters + oi : nullptr, gridDim.x);
}

__global__ void __launch_bounds__(BLOCK) k_sumtable_dna4(const PartView *__restrict__ parts, const nrx_pair *__restrict__ pairs) {
  __shared__ double sV[16], sIV[16], sF[4];   // eigenvecs [j][k], inv_eigenvecs [k][j], freqs
  const PartView &pv = parts[blockIdx.z];
  nrx_pair pr = pairs[blockIdx.y];
  if (pr.b_kind == NRX_TIP) { nrx_pair t = pr; pr.a_kind = t.b_kind; pr.a_idx = t.b_idx; pr.b_kind = t.a_kind; pr.b_idx = t.a_idx; }
  const int tid = threadIdx.x;
  if (tid < 16) { sV[tid] = pv.eigenvecs[tid]; sIV[tid] = pv.inv_eigenvecs[tid]; }
  if (tid < 4) sF[tid] = pv.freqs[tid];
  __syncthreads();
  const bool tipl = pr.a_kind == NRX_TIP;
  const double *clvl = tipl ? nullptr : pv.clv[pr.a_idx];
  const uint8_t *tip = tipl ? pv.tipchars + (size_t)pr.a_idx * pv.tip_pitch : nullptr;
  const double *clvr = pv.clv[pr.b_idx];
  double *out = pv.sumtable[blockIdx.y];
  const uint64_t n_items = (uint64_t)pv.patterns * 4;
  for (uint64_t base = (uint64_t)blockIdx.x * BLOCK * RU; base < n_items; base += (uint64_t)gridDim.x * BLOCK * RU) {
    D4 a[RU], b[RU];
#pragma unroll
    for (int u = 0; u < RU; ++u) {
      const uint64_t g = base + (uint64_t)u * BLOCK + tid;
      if (g < n_items) {
        b[u] = ldg256(clvr + g * 4);
        if (!tipl) a[u] = ldg256(clvl + g * 4);
        else { const uint32_t m = tip[g >> 2] & 15; a[u].x = (double)(m & 1u); a[u].y = (double)((m >> 1) & 1u); a[u].z = (double)((m >> 2) & 1u); a[u].w = (double)((m >> 3) & 1u); }
      }
    }
#pragma unroll
    for (int u = 0; u < RU; ++u) {
      const uint64_t g = base + (uint64_t)u * BLOCK + tid;
      if (g >= n_items) continue;
      const double lf[4] = {__dmul_rn(a[u].x, sF[0]), __dmul_rn(a[u].y, sF[1]), __dmul_rn(a[u].z, sF[2]), __dmul_rn(a[u].w, sF[3])};
      const double rv[4] = {b[u].x, b[u].y, b[u].z, b[u].w};
      double o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        double le = 0.0, ri = 0.0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          le = __dadd_rn(le, __dmul_rn(lf[k], sIV[k * 4 + j]));
          ri = __dadd_rn(ri, __dmul_rn(sV[j * 4 + k], rv[k]));
        }
        o[j] = __dmul_rn(le, ri);
      }
      D4 ov; ov.x = o[0]; ov.y = o[1]; ov.z = o[2]; ov.w = o[3];
      stg256(out + g * 4, ov);
    }
  }
}

__device__ __forceinline__ void deriv_cat(const D4 &v, const double *dg, double w, double &lk0, double &lk1, double &lk2, bool first,
                                          double pinv = 0.0, double invf = 0.0) {
  const double sv[4] = {v.x, v.y, v.z, v.w};
  double c0 = 0.0, c1 = 0.0, c2 = 0.0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    c0 = __dadd_rn(c0, __dmul_rn(sv[j], dg[j * 4 + 0]));
    c1 = __dadd_rn(c1, __dmul_rn(sv[j], dg[j * 4 + 1]));
    c2 = __dadd_rn(c2, __dmul_rn(sv[j], dg[j * 4 + 2]));
  }
  deriv_cat_pinv(c0, c1, c2, pinv, invf);
  if (first) { lk0 = __dmul_rn(c0, w); lk1 = __dmul_rn(c1, w); lk2 = __dmul_rn(c2, w); }
  else { lk0 = __dadd_rn(lk0, __dmul_rn(c0, w)); lk1 = __dadd_rn(lk1, __dmul_rn(c1, w)); lk2 = __dadd_rn(lk2, __dmul_rn(c2, w)); }
}

__global__ void __launch_bounds__(BLOCK) k_derivatives_dna4(const PartView *__restrict__ parts, double *__restrict__ partial,
                                                             uint32_t nparts_total, double *__restrict__ out, uint32_t *__restrict__ counters) {
  __shared__ double red[3 * (BLOCK / 32)];
  __shared__ double sD[64];   // diag table [cat][state][4]: every lane reads the same entry (broadcast)
  const PartView &pv = parts[blockIdx.z];
  const int tid = threadIdx.x;
  if (tid < 64) sD[tid] = pv.diagp[tid];
  __syncthreads();
  const double *st = pv.sumtable[blockIdx.y];
  const double w0 = pv.rate_weights[0], w1 = pv.rate_weights[1], w2 = pv.rate_weights[2], w3 = pv.rate_weights[3];
  const double pinv = pv.pinv;
  double acc[3] = {0.0, 0.0, 0.0};
  for (uint64_t n = (uint64_t)blockIdx.x * BLOCK + threadIdx.x; n < pv.patterns; n += (uint64_t)gridDim.x * BLOCK) {
    const double *c = st + n * 16;
    const D4 v0 = ldg256(c), v1 = ldg256(c + 4), v2 = ldg256(c + 8), v3 = ldg256(c + 12);
    const double pw = (double)pv.weights[n];
    double lk0, lk1, lk2, invf = 0.0;
    if (pinv > 0.0) { const int iv = pv.invariant[n]; invf = iv < 0 ? 0.0 : pv.freqs[iv]; }
    deriv_cat(v0, sD, w0, lk0, lk1, lk2, true, pinv, invf);
    deriv_cat(v1, sD + 16, w1, lk0, lk1, lk2, false, pinv, invf);
    deriv_cat(v2, sD + 32, w2, lk0, lk1, lk2, false, pinv, invf);
    deriv_cat(v3, sD + 48, w3, lk0, lk1, lk2, false, pinv, invf);
    const double d1 = -lk1 / lk0;
    const double d2 = d1 * d1 - lk2 / lk0;
    acc[0] += pw * log(lk0);
    acc[1] += pw * d1;
    acc[2] += pw * d2;
  }
  block_sum<3>(acc, red);
  const size_t oi = (size_t)blockIdx.y * nparts_total + pv.part_index;
  double *p = partial + oi * 3 * gridDim.x;
  if (threadIdx.x == 0) {
    p[0 * gridDim.x + blockIdx.x] = acc[0];
    p[1 * gridDim.x + blockIdx.x] = acc[1];
    p[2 * gridDim.x + blockIdx.x] = acc[2];
  }
  finish_partials<3>(p, out + oi * 3, counters ? counters + oi : nullptr, gridDim.x);
}

/* ------------------------------------------------------------------------------------------------
 * K3 / K4 / K6, DNA 4x4, "quad" versions (round 2).  The thread-per-pattern kernels above read 32 bytes per lane at a 128-byte
 * stride: every sector is used, but a warp-wide load touches 32 different lines, and the L1 tag stage (one line per clock) then
 * tops out at about the HBM rate — measured 0.35-0.54 of the copy peak inside the derivative sweep
 * (profiles/r3b_sweep_shadow_reroot_memo_config2.md) against 0.78 for k_sumtable_dna4, which loads thread = (pattern, category).
 * Here every load is thread = (pattern, category) (a warp reads 1 KB contiguous = 8 lines), four independent 256-bit loads per
 * operand and thread are in flight, the per-category constants (P-matrix block, diag-table rows, weight) sit in registers because
 * a thread's category never changes, and the four categories of a pattern are brought together by a 3-shuffle transposing
 * reduction inside the quad that leaves lane q of the quad with the sum of the thread's q-th item — so all 32 lanes still take
 * one log / division each.  Category sum order is the fixed tree (c0 + c2) + (c1 + c3).
 * ---------------------------------------------------------------------------------------------- */
__device__ __forceinline__ double quad_gather_sum(double x0, double x1, double x2, double x3, int q) {
  const bool hi = (q & 2) != 0;
  double keepA = hi ? x2 : x0, keepB = hi ? x3 : x1;
  const double sendA = hi ? x0 : x2, sendB = hi ? x1 : x3;
  keepA = __dadd_rn(keepA, __shfl_xor_sync(0xffffffffu, sendA, 2));
  keepB = __dadd_rn(keepB, __shfl_xor_sync(0xffffffffu, sendB, 2));
  const bool odd = (q & 1) != 0;
  const double keep = odd ? keepB : keepA, send = odd ? keepA : keepB;
  return __dadd_rn(keep, __shfl_xor_sync(0xffffffffu, send, 1));
}

constexpr int QU = 4;   // items per thread and pass = lanes of a quad

/* one pass of a block = BLOCK * QU consecutive (pattern, category) items = 256 patterns; the next pass's loads are issued before
 * the current pass is consumed (register double buffer), blocks are few and long-lived (about two resident blocks per SM in total,
 * nrx_engine.cu:quad_blocks): ncu showed the one-pass-per-block form latency-bound — 36 % issue slots, 30 % FP64 pipe, 3.3 TB/s —
 * because a block's loads, its log / division tail and its block reduction ran strictly one after the other. */
__device__ __forceinline__ void quad_load(D4 (&v)[QU], const double *__restrict__ src, uint64_t base, uint64_t n_items, int tid) {
#pragma unroll
  for (int u = 0; u < QU; ++u) {
    const uint64_t g = base + (uint64_t)u * BLOCK + tid;
    if (g < n_items) v[u] = ldg256(src + g * 4);
    else { v[u].x = v[u].y = v[u].z = v[u].w = 0.0; }
  }
}

__global__ void __launch_bounds__(BLOCK, 2) k_derivatives_dna4q(const PartView *__restrict__ parts, double *__restrict__ partial,
                                                                 uint32_t nparts_total, double *__restrict__ out, uint32_t *__restrict__ counters) {
  __shared__ double red[3 * (BLOCK / 32)];
  const PartView &pv = parts[blockIdx.z];
  const int tid = threadIdx.x, q = tid & 3;
  double dg[12];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int k = 0; k < 3; ++k) dg[j * 3 + k] = pv.diagp[q * 16 + j * 4 + k];
  const double w = pv.rate_weights[q], pinv = pv.pinv;
  const double *st = pv.sumtable[blockIdx.y];
  const uint64_t n_items = (uint64_t)pv.patterns * 4, stride = (uint64_t)gridDim.x * BLOCK * QU;
  double acc[3] = {0.0, 0.0, 0.0};
  uint64_t base = (uint64_t)blockIdx.x * BLOCK * QU;
  D4 nxt[QU];
  quad_load(nxt, st, base, n_items, tid);
  for (; base < n_items; base += stride) {
    D4 v[QU];
#pragma unroll
    for (int u = 0; u < QU; ++u) v[u] = nxt[u];
    if (base + stride < n_items) quad_load(nxt, st, base + stride, n_items, tid);
    double l0[QU], l1[QU], l2[QU];
#pragma unroll
    for (int u = 0; u < QU; ++u) {
      const double sv[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
      double c0 = 0.0, c1 = 0.0, c2 = 0.0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        c0 = __dadd_rn(c0, __dmul_rn(sv[j], dg[j * 3 + 0]));
        c1 = __dadd_rn(c1, __dmul_rn(sv[j], dg[j * 3 + 1]));
        c2 = __dadd_rn(c2, __dmul_rn(sv[j], dg[j * 3 + 2]));
      }
      if (pinv > 0.0) {
        const uint64_t g = base + (uint64_t)u * BLOCK + tid;
        double invf = 0.0;
        if (g < n_items) { const int iv = pv.invariant[g >> 2]; invf = iv < 0 ? 0.0 : pv.freqs[iv]; }
        deriv_cat_pinv(c0, c1, c2, pinv, invf);
      }
      l0[u] = __dmul_rn(c0, w); l1[u] = __dmul_rn(c1, w); l2[u] = __dmul_rn(c2, w);
    }
    const double lk0 = quad_gather_sum(l0[0], l0[1], l0[2], l0[3], q);
    const double lk1 = quad_gather_sum(l1[0], l1[1], l1[2], l1[3], q);
    const double lk2 = quad_gather_sum(l2[0], l2[1], l2[2], l2[3], q);
    const uint64_t n = (base >> 2) + (uint64_t)q * (BLOCK / 4) + (tid >> 2);   // the pattern of this thread's q-th item
    if (n < pv.patterns) {
      const double pw = (double)pv.weights[n];
      const double d1 = -lk1 / lk0;
      const double d2 = d1 * d1 - lk2 / lk0;
      acc[0] += pw * log(lk0);
      acc[1] += pw * d1;
      acc[2] += pw * d2;
    }
  }
  block_sum<3>(acc, red);
  const size_t oi = (size_t)blockIdx.y * nparts_total + pv.part_index;
  double *p = partial + oi * 3 * gridDim.x;
  if (threadIdx.x == 0) {
    p[0 * gridDim.x + blockIdx.x] = acc[0];
    p[1 * gridDim.x + blockIdx.x] = acc[1];
    p[2 * gridDim.x + blockIdx.x] = acc[2];
  }
  finish_partials<3>(p, out + oi * 3, counters ? counters + oi : nullptr, gridDim.x);
}

__global__ void __launch_bounds__(BLOCK, 2) k_tree_lnl_dna4q(const PartView *__restrict__ parts, const uint32_t *__restrict__ slots,
                                                              double *__restrict__ partial, uint32_t nparts_total, double log_thresh,
                                                              double *__restrict__ persite, size_t persite_stride,
                                                              double *__restrict__ out, uint32_t *__restrict__ counters) {
  __shared__ double red[BLOCK / 32];
  const PartView &pv = parts[blockIdx.z];
  const int tid = threadIdx.x, q = tid & 3;
  const uint32_t slot = slots[blockIdx.y];
  const double *clv = pv.clv[slot];
  const uint32_t *sc = pv.scaler[slot];
  const double f0 = pv.freqs[0], f1 = pv.freqs[1], f2 = pv.freqs[2], f3 = pv.freqs[3];
  const double w = pv.rate_weights[q], pinv = pv.pinv;
  const uint64_t n_items = (uint64_t)pv.patterns * 4, stride = (uint64_t)gridDim.x * BLOCK * QU;
  double acc[1] = {0.0};
  uint64_t base = (uint64_t)blockIdx.x * BLOCK * QU;
  D4 nxt[QU];
  quad_load(nxt, clv, base, n_items, tid);
  for (; base < n_items; base += stride) {
    D4 v[QU];
#pragma unroll
    for (int u = 0; u < QU; ++u) v[u] = nxt[u];
    if (base + stride < n_items) quad_load(nxt, clv, base + stride, n_items, tid);
    double t[QU];
#pragma unroll
    for (int u = 0; u < QU; ++u) {
      double invf = 0.0;
      if (pinv > 0.0) {
        const uint64_t g = base + (uint64_t)u * BLOCK + tid;
        if (g < n_items) { const int iv = pv.invariant[g >> 2]; invf = iv < 0 ? 0.0 : pv.freqs[iv]; }
      }
      t[u] = root_cat_term(tree4(__dmul_rn(f0, v[u].x), __dmul_rn(f1, v[u].y), __dmul_rn(f2, v[u].z), __dmul_rn(f3, v[u].w)), w, pinv, invf);
    }
    const double term = quad_gather_sum(t[0], t[1], t[2], t[3], q);
    const uint64_t n = (base >> 2) + (uint64_t)q * (BLOCK / 4) + (tid >> 2);
    if (n < pv.patterns) {
      const uint32_t s = sc[n];
      double lk = log(term);
      if (s) lk = __dadd_rn(lk, __dmul_rn((double)s, log_thresh));
      lk = __dmul_rn(lk, (double)pv.weights[n]);
      if (persite) persite[((size_t)blockIdx.y * nparts_total + pv.part_index) * persite_stride + n] = lk;
      acc[0] += lk;
    }
  }
  block_sum<1>(acc, red);
  const size_t oi = (size_t)blockIdx.y * nparts_total + pv.part_index;
  if (threadIdx.x == 0) partial[oi * gridDim.x + blockIdx.x] = acc[0];
  finish_partials<1>(partial + oi * gridDim.x, out + oi, counters ? counters + oi : nullptr, gridDim.x);
}

/* K4: two items per thread and half-pass (a quad still collects four items before it transposes), P(edge) block of the thread's
 * category in shared memory at the padded pitch — the double buffer of both operands would not fit 128 registers otherwise */
constexpr int QH = 2;
__device__ __forceinline__ void quad_load_half(D4 (&a)[QH], D4 (&b)[QH], const double *__restrict__ clvp, const double *__restrict__ clvc,
                                               const uint8_t *__restrict__ tip, const double *__restrict__ lut, uint64_t base, uint64_t n_items, int tid, int q) {
#pragma unroll
  for (int u = 0; u < QH; ++u) {
    const uint64_t g = base + (uint64_t)u * BLOCK + tid;
    if (g < n_items) {
      a[u] = ldg256(clvp + g * 4);
      if (clvc) b[u] = ldg256(clvc + g * 4);
      else b[u] = *reinterpret_cast<const D4 *>(lut + (tip[g >> 2] & 15) * 16 + q * 4);
    } else { a[u].x = a[u].y = a[u].z = a[u].w = 0.0; b[u] = a[u]; }
  }
}

__global__ void __launch_bounds__(BLOCK, 2) k_edge_lnl_dna4q(const PartView *__restrict__ parts, const nrx_pair *__restrict__ pairs,
                                                              uint32_t edge, double *__restrict__ partial, uint32_t nparts_total,
                                                              double log_thresh, double *__restrict__ out, uint32_t *__restrict__ counters) {
  __shared__ double red[BLOCK / 32];
  __shared__ __align__(32) double lut[256];
  __shared__ __align__(16) double sP[4 * PCAT];
  const PartView &pv = parts[blockIdx.z];
  nrx_pair pr = pairs[blockIdx.y];
  if (pr.a_kind == NRX_TIP) { nrx_pair t = pr; pr.a_kind = t.b_kind; pr.a_idx = t.b_idx; pr.b_kind = t.a_kind; pr.b_idx = t.a_idx; }
  const int tid = threadIdx.x, q = tid & 3;
  const bool tipc = pr.b_kind == NRX_TIP;
  if (tipc) build_tip_lut4(lut, pv.pmat + (size_t)edge * 64, tid);
  else if (tid < 64) sP[(tid >> 4) * PCAT + (tid & 15)] = pv.pmat[(size_t)edge * 64 + tid];
  __syncthreads();
  const double *Pq = sP + q * PCAT;
  const double *clvp = pv.clv[pr.a_idx];
  const uint32_t *scp = pv.scaler[pr.a_idx];
  const double *clvc = tipc ? nullptr : pv.clv[pr.b_idx];
  const uint32_t *scc = tipc ? nullptr : pv.scaler[pr.b_idx];
  const uint8_t *tip = tipc ? pv.tipchars + (size_t)pr.b_idx * pv.tip_pitch : nullptr;
  const double f0 = pv.freqs[0], f1 = pv.freqs[1], f2 = pv.freqs[2], f3 = pv.freqs[3];
  const double w = pv.rate_weights[q], pinv = pv.pinv;
  const uint64_t n_items = (uint64_t)pv.patterns * 4, stride = (uint64_t)gridDim.x * BLOCK * QU;
  double acc[1] = {0.0};
  uint64_t base = (uint64_t)blockIdx.x * BLOCK * QU;
  D4 na[QH], nb[QH];
  quad_load_half(na, nb, clvp, clvc, tip, lut, base, n_items, tid, q);
  for (; base < n_items; base += stride) {
    double ta[QU], ti[QU];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      D4 a[QH], b[QH];
#pragma unroll
      for (int u = 0; u < QH; ++u) { a[u] = na[u]; b[u] = nb[u]; }
      // next half-pass: the second half of this pass, or the first half of the block's next pass
      const uint64_t nbase = h == 0 ? base + (uint64_t)QH * BLOCK : base + stride;
      if (h == 0 || nbase < n_items) quad_load_half(na, nb, clvp, clvc, tip, lut, nbase, n_items, tid, q);
#pragma unroll
      for (int u = 0; u < QH; ++u) {
        const D4 y = tipc ? b[u] : matvec4(Pq, b[u]);
        const double tr = edge_cat_term(a[u], y, f0, f1, f2, f3);
        const int k = h * QH + u;
        if (pinv > 0.0) {
          const uint64_t g = base + (uint64_t)k * BLOCK + tid;
          int iv = -1;
          if (g < n_items) iv = pv.invariant[g >> 2];
          const double invf = iv < 0 ? 0.0 : pv.freqs[iv];
          ta[k] = 0.0; ti[k] = 0.0;
          edge_cat_accum(tr, w, pinv, invf, iv >= 0, ta[k], ti[k]);
        } else { ta[k] = __dmul_rn(tr, w); ti[k] = 0.0; }
      }
    }
    const double terma = quad_gather_sum(ta[0], ta[1], ta[2], ta[3], q);
    double terminv = 0.0;
    if (pinv > 0.0) terminv = quad_gather_sum(ti[0], ti[1], ti[2], ti[3], q);
    const uint64_t n = (base >> 2) + (uint64_t)q * (BLOCK / 4) + (tid >> 2);
    if (n < pv.patterns) {
      uint32_t s = scp[n];
      if (!tipc) s += scc[n];
      double lk;
      if (pinv > 0.0) lk = edge_site_lnl(terma, terminv, s, log_thresh);
      else { lk = log(terma); if (s) lk = __dadd_rn(lk, __dmul_rn((double)s, log_thresh)); }
      acc[0] += __dmul_rn(lk, (double)pv.weights[n]);
    }
  }
  block_sum<1>(acc, red);
  const size_t oi = (size_t)blockIdx.y * nparts_total + pv.part_index;
  if (threadIdx.x == 0) partial[oi * gridDim.x + blockIdx.x] = acc[0];
  finish_partials<1>(partial + oi * gridDim.x, out + oi, counters ? counters + oi : nullptr, gridDim.x);
}

/* K6 with the sumtable streamed through a cp.async.bulk ring in shared memory (K2's recipe) instead of through registers: ncu on
 * k_derivatives_dna4q (profiles/r4b_sweep_kernels_after_ncu.md) showed 64 KB of loads in flight per SM — all 118 registers allow —
 * and a third of the samples waiting on memory at 4.75 TB/s.  Here one elected thread keeps K6R_STAGES tiles of 256 patterns
 * (32 KB each) in flight per block, 2 blocks per SM = 192 KB; the consumers read their items from shared memory (thread = (pattern,
 * category), conflict-free 256-bit reads) and finish exactly as the register version does (same arithmetic, same quad reduction).
 * MEASURED SLOWER and therefore opt-in (NRX_K6_RING=1): 16.5 vs 14.9 ms for the 330 K6 launches of the config-2 sweep (0.62 vs 0.69
 * of the HBM peak) — a 45-us launch of 18 blocks per pair does not amortise the ring's fill and its lock-step refill barrier. */
constexpr int K6R_TP = 256;       // patterns per tile
constexpr int K6R_STAGES = 3;
struct __align__(128) K6RingSmem {
  double st[K6R_STAGES][K6R_TP * 16];
  unsigned long long full[K6R_STAGES];
};

__global__ void __launch_bounds__(BLOCK, 2) k_derivatives_dna4r(const PartView *__restrict__ parts, double *__restrict__ partial,
                                                                 uint32_t nparts_total, double *__restrict__ out, uint32_t *__restrict__ counters) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  K6RingSmem &sm = *reinterpret_cast<K6RingSmem *>(smem_raw);
  __shared__ double red[3 * (BLOCK / 32)];
  const PartView &pv = parts[blockIdx.z];
  const int tid = threadIdx.x, q = tid & 3;
  double dg[12];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int k = 0; k < 3; ++k) dg[j * 3 + k] = pv.diagp[q * 16 + j * 4 + k];
  const double w = pv.rate_weights[q], pinv = pv.pinv;
  const double *st = pv.sumtable[blockIdx.y];
  const uint32_t patterns = pv.patterns;
  const uint32_t ntiles = (patterns + K6R_TP - 1) / K6R_TP, first = blockIdx.x, step = gridDim.x;
  const uint32_t count = first < ntiles ? (ntiles - first + step - 1) / step : 0u;
  auto issue = [&](uint32_t k, uint32_t stage) {
    const uint32_t p0 = (first + k * step) * K6R_TP;
    const uint32_t rows = patterns - p0 < (uint32_t)K6R_TP ? patterns - p0 : (uint32_t)K6R_TP;
    mbar_expect_tx(&sm.full[stage], rows * 128u);
    bulk_g2s(sm.st[stage], st + (size_t)p0 * 16, rows * 128u, &sm.full[stage]);
  };
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < K6R_STAGES; ++s) mbar_init(&sm.full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (uint32_t s = 0; s < (uint32_t)K6R_STAGES && s < count; ++s) issue(s, s);
  }
  __syncthreads();
  double acc[3] = {0.0, 0.0, 0.0};
  uint32_t stage = 0, phase = 0;
  for (uint32_t k = 0; k < count; ++k) {
    mbar_wait(&sm.full[stage], phase);
    const uint32_t p0 = (first + k * step) * K6R_TP;
    const double *tile = sm.st[stage];
    double l0[QU], l1[QU], l2[QU];
#pragma unroll
    for (int u = 0; u < QU; ++u) {
      const uint32_t it = (uint32_t)u * BLOCK + tid;   // item of the tile: pattern it / 4, category q
      const D4 v = *reinterpret_cast<const D4 *>(tile + (size_t)it * 4);
      const double sv[4] = {v.x, v.y, v.z, v.w};
      double c0 = 0.0, c1 = 0.0, c2 = 0.0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        c0 = __dadd_rn(c0, __dmul_rn(sv[j], dg[j * 3 + 0]));
        c1 = __dadd_rn(c1, __dmul_rn(sv[j], dg[j * 3 + 1]));
        c2 = __dadd_rn(c2, __dmul_rn(sv[j], dg[j * 3 + 2]));
      }
      if (pinv > 0.0) {
        const uint32_t n = p0 + (it >> 2);
        double invf = 0.0;
        if (n < patterns) { const int iv = pv.invariant[n]; invf = iv < 0 ? 0.0 : pv.freqs[iv]; }
        deriv_cat_pinv(c0, c1, c2, pinv, invf);
      }
      l0[u] = __dmul_rn(c0, w); l1[u] = __dmul_rn(c1, w); l2[u] = __dmul_rn(c2, w);
    }
    const double lk0 = quad_gather_sum(l0[0], l0[1], l0[2], l0[3], q);
    const double lk1 = quad_gather_sum(l1[0], l1[1], l1[2], l1[3], q);
    const double lk2 = quad_gather_sum(l2[0], l2[1], l2[2], l2[3], q);
    const uint32_t n = p0 + (uint32_t)q * (BLOCK / 4) + (tid >> 2);   // the pattern of this thread's q-th item (rows past the end of a ragged
    if (n < patterns) {                                               //  last tile hold stale data and are never owned)
      const double pw = (double)pv.weights[n];
      const double d1 = -lk1 / lk0;
      const double d2 = d1 * d1 - lk2 / lk0;
      acc[0] += pw * log(lk0);
      acc[1] += pw * d1;
      acc[2] += pw * d2;
    }
    __syncthreads();   // every thread has read the stage: refill it
    if (tid == 0 && k + K6R_STAGES < count) issue(k + K6R_STAGES, stage);
    if (++stage == (uint32_t)K6R_STAGES) { stage = 0; phase ^= 1u; }
  }
  block_sum<3>(acc, red);
  const size_t oi = (size_t)blockIdx.y * nparts_total + pv.part_index;
  double *p = partial + oi * 3 * gridDim.x;
  if (threadIdx.x == 0) {
    p[0 * gridDim.x + blockIdx.x] = acc[0];
    p[1 * gridDim.x + blockIdx.x] = acc[1];
    p[2 * gridDim.x + blockIdx.x] = acc[2];
  }
  finish_partials<3>(p, out + oi * 3, counters ? counters + oi : nullptr, gridDim.x);
}

/* K4 + K5 in one pass (round 2): the edge lnL of a displayed-tree pair and its sumtable read the same two CLVs.  optimize_branch
 * asks for both back to back (computeLoglikelihoodBrlenOpt, then computePartitionSumtables; BranchLengthOptimization.cpp:374-381), so
 * the host may issue them together (nrx_edge_lnl_sumtables): 2C read + C written per pair and pattern instead of (2C) + (2C + C).
 * Pair i writes sumtable slot i; lnl_index[i] >= 0: it is also edge-lnL output lnl_index[i] (the sumtable list is a superset of the
 * edge-lnL list: active vs. active-and-alive branch).  a must be a CLV (the source side of a branch always is), b a CLV or a tip.
 * Arithmetic per output as in k_edge_lnl_dna4q and k_sumtable_dna4. */
__device__ __forceinline__ void es_load_half(D4 (&a)[QH], D4 (&b)[QH], uint32_t (&code)[QH], const double *__restrict__ clva, const double *__restrict__ clvb,
                                             const uint8_t *__restrict__ tip, uint64_t base, uint64_t n_items, int tid) {
#pragma unroll
  for (int u = 0; u < QH; ++u) {
    const uint64_t g = base + (uint64_t)u * BLOCK + tid;
    code[u] = 0u;
    if (g < n_items) {
      a[u] = ldg256(clva + g * 4);
      if (clvb) b[u] = ldg256(clvb + g * 4);
      else { code[u] = tip[g >> 2] & 15u; b[u].x = b[u].y = b[u].z = b[u].w = 0.0; }
    } else { a[u].x = a[u].y = a[u].z = a[u].w = 0.0; b[u] = a[u]; }
  }
}

__global__ void __launch_bounds__(BLOCK, 2) k_edge_sum_dna4q(const PartView *__restrict__ parts, const nrx_pair *__restrict__ pairs,
                                                              const int32_t *__restrict__ lnl_index, uint32_t edge,
                                                              double *__restrict__ partial, uint32_t nparts_total, double log_thresh,
                                                              double *__restrict__ out, uint32_t *__restrict__ counters) {
  __shared__ double red[BLOCK / 32];
  __shared__ __align__(32) double lut[256];
  __shared__ __align__(16) double sP[4 * PCAT];
  __shared__ double sV[16], sIV[16], sF[4];
  const PartView &pv = parts[blockIdx.z];
  const nrx_pair pr = pairs[blockIdx.y];
  const int li = lnl_index[blockIdx.y];
  const int tid = threadIdx.x, q = tid & 3;
  const bool tipb = pr.b_kind == NRX_TIP;
  if (tipb) build_tip_lut4(lut, pv.pmat + (size_t)edge * 64, tid);
  else if (tid < 64) sP[(tid >> 4) * PCAT + (tid & 15)] = pv.pmat[(size_t)edge * 64 + tid];
  if (tid < 16) { sV[tid] = pv.eigenvecs[tid]; sIV[tid] = pv.inv_eigenvecs[tid]; }
  if (tid < 4) sF[tid] = pv.freqs[tid];
  __syncthreads();
  const double *Pq = sP + q * PCAT;
  const double *clva = pv.clv[pr.a_idx];
  const uint32_t *sca = pv.scaler[pr.a_idx];
  const double *clvb = tipb ? nullptr : pv.clv[pr.b_idx];
  const uint32_t *scb = tipb ? nullptr : pv.scaler[pr.b_idx];
  const uint8_t *tip = tipb ? pv.tipchars + (size_t)pr.b_idx * pv.tip_pitch : nullptr;
  double *st = pv.sumtable[blockIdx.y];
  const double f0 = sF[0], f1 = sF[1], f2 = sF[2], f3 = sF[3];
  const double w = pv.rate_weights[q], pinv = pv.pinv;
  const uint64_t n_items = (uint64_t)pv.patterns * 4, stride = (uint64_t)gridDim.x * BLOCK * QU;
  double acc[1] = {0.0};
  uint64_t base = (uint64_t)blockIdx.x * BLOCK * QU;
  D4 na[QH], nb[QH];
  uint32_t nc[QH];
  es_load_half(na, nb, nc, clva, clvb, tip, base, n_items, tid);
  for (; base < n_items; base += stride) {
    double ta[QU], ti[QU];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      D4 a[QH], b[QH];
      uint32_t code[QH];
#pragma unroll
      for (int u = 0; u < QH; ++u) { a[u] = na[u]; b[u] = nb[u]; code[u] = nc[u]; }
      const uint64_t nbase = h == 0 ? base + (uint64_t)QH * BLOCK : base + stride;
      if (h == 0 || nbase < n_items) es_load_half(na, nb, nc, clva, clvb, tip, nbase, n_items, tid);
#pragma unroll
      for (int u = 0; u < QH; ++u) {
        const int k = h * QH + u;
        const uint64_t g = base + (uint64_t)k * BLOCK + tid;
        // ---- K5: sumtable entry (k_sumtable_dna4: the tip, if any, is the left operand) ----
        D4 left = a[u], right = b[u];
        if (tipb) {
          right = a[u];
          left.x = (double)(code[u] & 1u); left.y = (double)((code[u] >> 1) & 1u); left.z = (double)((code[u] >> 2) & 1u); left.w = (double)((code[u] >> 3) & 1u);
        }
        const double lf[4] = {__dmul_rn(left.x, f0), __dmul_rn(left.y, f1), __dmul_rn(left.z, f2), __dmul_rn(left.w, f3)};
        const double rv[4] = {right.x, right.y, right.z, right.w};
        double o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          double le = 0.0, ri = 0.0;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            le = __dadd_rn(le, __dmul_rn(lf[kk], sIV[kk * 4 + j]));
            ri = __dadd_rn(ri, __dmul_rn(sV[j * 4 + kk], rv[kk]));
          }
          o[j] = __dmul_rn(le, ri);
        }
        if (g < n_items) { D4 ov; ov.x = o[0]; ov.y = o[1]; ov.z = o[2]; ov.w = o[3]; stg256(st + g * 4, ov); }
        // ---- K4: edge lnL term (k_edge_lnl_dna4q: a is the parent, b goes through P(edge)) ----
        ta[k] = 0.0; ti[k] = 0.0;
        if (li >= 0) {
          const D4 y = tipb ? *reinterpret_cast<const D4 *>(lut + code[u] * 16 + q * 4) : matvec4(Pq, b[u]);
          const double tr = edge_cat_term(a[u], y, f0, f1, f2, f3);
          if (pinv > 0.0) {
            int iv = -1;
            if (g < n_items) iv = pv.invariant[g >> 2];
            const double invf = iv < 0 ? 0.0 : pv.freqs[iv];
            edge_cat_accum(tr, w, pinv, invf, iv >= 0, ta[k], ti[k]);
          } else ta[k] = __dmul_rn(tr, w);
        }
      }
    }
    if (li >= 0) {
      const double terma = quad_gather_sum(ta[0], ta[1], ta[2], ta[3], q);
      double terminv = 0.0;
      if (pinv > 0.0) terminv = quad_gather_sum(ti[0], ti[1], ti[2], ti[3], q);
      const uint64_t n = (base >> 2) + (uint64_t)q * (BLOCK / 4) + (tid >> 2);
      if (n < pv.patterns) {
        uint32_t s = sca[n];
        if (!tipb) s += scb[n];
        double lk;
        if (pinv > 0.0) lk = edge_site_lnl(terma, terminv, s, log_thresh);
        else { lk = log(terma); if (s) lk = __dadd_rn(lk, __dmul_rn((double)s, log_thresh)); }
        acc[0] += __dmul_rn(lk, (double)pv.weights[n]);
      }
    }
  }
  if (li < 0) return;   // block-uniform: this pair has no edge-lnL output
  block_sum<1>(acc, red);
  const size_t oi = (size_t)li * nparts_total + pv.part_index;
  if (threadIdx.x == 0) partial[oi * gridDim.x + blockIdx.x] = acc[0];
  finish_partials<1>(partial + oi * gridDim.x, out + oi, counters ? counters + oi : nullptr, gridDim.x);
}

}  // namespace nrx
