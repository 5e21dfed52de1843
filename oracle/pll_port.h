/*
 * pll_port.h — ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the numerical kernels of the reference's forked libpll that sit on the
 * NetRAX network-likelihood hot path (SURVEY.md §2.3 K1–K6).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this; the product (netrax_b200/)
 * never links or imports anything under oracle/.
 *
 * Parity pinning: every function here is checked in tests/test_oracle_*.py against
 *   (i)  the golden numbers of libpll's own regression suite (test/out/derivatives.out, copied as
 *        numbers into tests/golden/libpll_derivatives_golden.json), and
 *   (ii) the real forked libpll compiled from /root/reference (oracle/_ref/libpll_ref.so).
 *
 * LIBPLL = /root/reference/libs/raxml-ng/libs/pll-modules/libs/libpll/src
 */
#ifndef PLL_PORT_H
#define PLL_PORT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* LIBPLL/pll.h:89-90 */
#define PORT_SCALE_FACTOR 115792089237316195423570985008687907853269984665640564039457584007913129639936.0 /* 2^256 */
#define PORT_SCALE_THRESHOLD (1.0 / PORT_SCALE_FACTOR)
#define PORT_EIGEN_MINFREQ 1e-6  /* LIBPLL/pll.h PLL_EIGEN_MINFREQ */
#define PORT_MISC_EPSILON 1e-8   /* LIBPLL/pll.h PLL_MISC_EPSILON */

/* One partition's model + data, the subset of pll_partition_t (LIBPLL/pll.h:230-277) the path reads.
 * Single rate matrix per partition (params_indices == 0 everywhere, as in every BASELINE config). */
#define PORT_MAX_MODELS 16

typedef struct port_partition {
  unsigned states;         /* 4 or 20 */
  unsigned states_padded;  /* (states+3)&~3, LIBPLL/pll.c:482-483 (AVX/AVX2) */
  unsigned rate_cats;
  unsigned sites;  /* patterns */
  unsigned tips;
  unsigned edges;  /* number of P-matrices incl. the fake one */
  double *freqs;          /* [states_padded] */
  double *subst_params;   /* [states*(states-1)/2] */
  double *eigenvecs;      /* [states*states_padded] */
  double *inv_eigenvecs;  /* [states*states_padded] */
  double *eigenvals;      /* [states_padded] */
  double *rates;          /* [rate_cats] */
  double *rate_weights;   /* [rate_cats] */
  double prop_invar;      /* +I: pll_partition_t::prop_invar[0] (one rate matrix) */
  int *invariant;         /* [sites] frequency index of an invariant pattern, -1 otherwise; NULL until +I is first used */
  /* several rate matrices, one per rate category (libpll rate_matrices + params_indices; LG4M / LG4X): matrix 0 is the set of
   * arrays above, matrices 1.. are allocated by port_set_submodels; cat_model[c] names the matrix of category c */
  unsigned nmodels;
  unsigned *cat_model;       /* [rate_cats] */
  double *m_freqs[PORT_MAX_MODELS], *m_subst[PORT_MAX_MODELS], *m_eigenvecs[PORT_MAX_MODELS], *m_inv_eigenvecs[PORT_MAX_MODELS],
      *m_eigenvals[PORT_MAX_MODELS];
  unsigned *pattern_weights; /* [sites] */
  unsigned char **tipchars;  /* [tips][sites]: code into tipmap (DNA: the 4-bit state mask itself) */
  uint32_t tipmap[256];      /* code -> state bit mask */
  unsigned maxstates;        /* number of codes in use (DNA: 16) */
  double **pmatrix;          /* [edges] -> [rate_cats][states][states_padded] */
} port_partition;

port_partition *port_partition_create(unsigned states, unsigned rate_cats, unsigned sites,
                                      unsigned tips, unsigned edges);
void port_partition_destroy(port_partition *p);

/* LIBPLL/models.c:651-750 pll_update_invariant_sites (PATTERN_TIP branch) and :495-543 pll_update_invariant_sites_proportion */
/* pll_set_frequencies / pll_set_subst_params / pll_update_eigen per matrix index + the params_indices every libpll call takes
 * (LIBPLL/models.c:445-493, core_pmatrix.c:182-185, core_likelihood.c:166, core_derivatives.c:362-366,712) */
int port_set_submodels(port_partition *p, unsigned n, const unsigned *cat_model, const double *freqs, const double *subst);
int port_update_invariant_sites(port_partition *p);
int port_set_prop_invar(port_partition *p, double prop_invar);
/* LIBPLL/gamma.c:267-330 pll_compute_gamma_cats, PLL_GAMMA_RATES_MEAN (mode 0) / MEDIAN (1) */
int port_compute_gamma_cats(double alpha, unsigned categories, double *out_rates, int mode);
/* LIBPLL/models.c:293-410 pll_update_eigen (Householder tridiagonalisation + QL, :24-180) */
int port_update_eigen(port_partition *p);
/* LIBPLL/models.c:412-443 + core_pmatrix.c:24-244 / core_pmatrix_avx.c:42 (4x4 op order) */
int port_update_pmatrix(port_partition *p, unsigned edge, double branch_length);

/* Child operand of a CLV update. kind: 0 inner CLV, 1 tip (tipchars), 2 absent ("fake" all-ones CLV
 * with identity P-matrix, LH/ImprovedLoglikelihood.cpp:122,138-139, src/RaxmlWrapper.cpp:156-187). */
typedef struct port_operand {
  int kind;
  const double *clv;         /* kind 0 */
  const unsigned *scaler;    /* kind 0, may be NULL */
  unsigned tip;              /* kind 1 */
  unsigned edge;             /* P-matrix index (ignored for kind 2) */
} port_operand;

/* LIBPLL/partials.c:196-240 pll_update_partials_single with PATTERN_TIP:
 * tip-tip (core_partials_avx.c:255,992), tip-inner (:1310), inner-inner (:402); per-site scaling. */
void port_update_partials(const port_partition *p, double *parent_clv, unsigned *parent_scaler,
                          const port_operand *left, const port_operand *right);

/* LIBPLL/likelihood.c:122-184 + core_likelihood.c:25-209 (root), no +I, no asc. bias. */
double port_root_loglikelihood(const port_partition *p, const double *clv, const unsigned *scaler,
                               double *persite_lnl);
/* LIBPLL/likelihood.c:555-615 + core_likelihood.c:1191-1496 (ii) / :351-922 (ti).
 * Exactly one of the two operands may be a tip (kind 1); tip-tip is invalid as in the reference. */
double port_edge_loglikelihood(const port_partition *p, const port_operand *parent,
                               const port_operand *child, unsigned edge, double *persite_lnl);
/* LIBPLL/derivatives.c:246-326 + core_derivatives.c:321-471 (ii) / :473-641 (ti).
 * sumtable has sites*rate_cats*states_padded doubles. */
int port_update_sumtable(const port_partition *p, const port_operand *parent,
                         const port_operand *child, double *sumtable);
/* LIBPLL/core_derivatives.c:696-728 pll_compute_diagptable: [rate_cats][states][4] */
void port_compute_diagptable(const port_partition *p, double branch_length, double *diagptable);
/* LIBPLL/derivatives.c:359-428 + core_derivatives.c:730-867 / core_derivatives_avx2.c:1538-1882.
 * Returns derivatives of MINUS lnL (Q6); *f (optional) is Σ w_n log(lk0_n) WITHOUT the scaler
 * term (quirk Q1: the AVX2 kernel is never given the scalers). */
int port_loglikelihood_derivatives(const port_partition *p, const double *sumtable,
                                   const double *diagptable, double *f, double *d_f, double *dd_f);

#ifdef __cplusplus
}
#endif
#endif
