/*
 * opt_port.c — ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the two 1-D minimisers of pll-modules that NetRAX's optimisers drive
 * (PLLMOD = /root/reference/libs/raxml-ng/libs/pll-modules/src):
 *   orcopt_newton_multi  <- pllmod_opt_minimize_newton_multi   PLLMOD/optimize/opt_algorithms.c:133-261
 *   orcopt_brent         <- pllmod_opt_minimize_brent          PLLMOD/optimize/opt_algorithms.c:1404-1429
 *   orcopt_brent_multi   <- pllmod_opt_minimize_brent_multi    PLLMOD/optimize/opt_algorithms.c:1448-1468 -> brent_opt_alt (any xnum)
 *                           -> brent_opt_alt (xnum = 1, global range) :1043-1254, brent_opt_init :859-939,
 *                              brent_opt_post_loop :941-1027
 * Same signatures as the originals so that the `_ref` build can bind the REAL functions instead
 * (oracle/Makefile compiles opt_algorithms.c where it lies; see optimize_port.cpp).  The restatement keeps the
 * original's evaluation sequence exactly — including brent_opt_alt's habit of calling the target with the last
 * proposal until 101 loop iterations have passed (the single-variable wrapper never raises the all-converged flag).
 * Pinned against the real functions by tests/test_oracle_optimize.py.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define BRENT_ITMAX 100
#define BRENT_GOLD 0.3819660
#define BRENT_ZEPS 1.0e-7

static double with_sign(double a, double b) { return b >= 0.0 ? fabs(a) : -fabs(a); }
/* PLL_MAX(PLL_MIN(v, hi), lo) with libpll's macros (LIBPLL/pll.h:67-68), operand order included: a NaN v (the Newton step
 * -0/0 when f = df = 0) comes out as hi, as in the reference */
static double clampd(double v, double lo, double hi) { const double m = v < hi ? v : hi; return m > lo ? m : lo; }

int orcopt_newton_multi(unsigned int xnum, double xmin, double *xguess, double xmax, double tolerance, unsigned int max_iters,
                        int *converged, void *params, void (*deriv_func)(void *, double *, double *, double *)) {
  unsigned int i, iter = 0;
  int all_converged = 0, error_flag = 0;
  const double dxmax = xmax / max_iters;
  double *xl = calloc(xnum, sizeof(double)), *xh = calloc(xnum, sizeof(double));
  double *f = calloc(xnum, sizeof(double)), *df = calloc(xnum, sizeof(double));
  int *own = NULL;
  double *x = xguess;
  if (!converged) converged = own = calloc(xnum, sizeof(int));
  for (i = 0; i < xnum; i++) { x[i] = clampd(x[i], xmin, xmax); xl[i] = xmin; xh[i] = xmax; }
  while (!all_converged && !error_flag) {
    if (iter++ > max_iters) { error_flag = 1; break; }
    deriv_func(params, x, f, df);
    all_converged = 1;
    for (i = 0; i < xnum; i++) {
      double dx;
      if (converged[i]) continue;
      if (!isfinite(f[i]) || !isfinite(df[i])) { error_flag = 1; break; }
      if (df[i] > 0.0) {
        if (fabs(f[i]) < tolerance) { converged[i] = 1; continue; }
        if (f[i] < 0.0) xl[i] = x[i]; else xh[i] = x[i];
        dx = -1 * f[i] / df[i];
      } else
        dx = -1 * f[i] / fabs(df[i]);
      dx = clampd(dx, -dxmax, dxmax);
      if (x[i] + dx < xl[i]) dx = xl[i] - x[i];
      if (x[i] + dx > xh[i]) dx = xh[i] - x[i];
      if (fabs(dx) < tolerance) { converged[i] = 1; continue; }
      x[i] += dx;
      x[i] = clampd(x[i], xmin, xmax);
      all_converged &= converged[i];
    }
  }
  free(xl); free(xh); free(f); free(df); free(own);
  return all_converged && !error_flag;
}

typedef struct {
  double startx, fstartx, tol, a, b, d, e, u, fu, v, w, x, fv, fw, fx;
} brent_state;

/* convergence test and next proposal (the "pre-loop" half shared by brent_opt_init and brent_opt_post_loop) */
static int brent_next(brent_state *s) {
  const double xm = 0.5 * (s->a + s->b);
  const double tol1 = s->tol * fabs(s->x) + BRENT_ZEPS, tol2 = 2.0 * tol1;
  if (fabs(s->x - xm) <= (tol2 - 0.5 * (s->b - s->a))) return 0;
  if (fabs(s->e) > tol1) {
    double r = (s->x - s->w) * (s->fx - s->fv), q = (s->x - s->v) * (s->fx - s->fw);
    double p = (s->x - s->v) * q - (s->x - s->w) * r, etemp;
    q = 2.0 * (q - r);
    if (q > 0.0) p = -p;
    q = fabs(q);
    etemp = s->e;
    s->e = s->d;
    if (fabs(p) >= fabs(0.5 * q * etemp) || p <= q * (s->a - s->x) || p >= q * (s->b - s->x))
      s->d = BRENT_GOLD * (s->e = (s->x >= xm ? s->a - s->x : s->b - s->x));
    else {
      s->d = p / q;
      s->u = s->x + s->d;
      if (s->u - s->a < tol2 || s->b - s->u < tol2) s->d = with_sign(tol1, xm - s->x);
    }
  } else
    s->d = BRENT_GOLD * (s->e = (s->x >= xm ? s->a - s->x : s->b - s->x));
  s->u = (fabs(s->d) >= tol1 ? s->x + s->d : s->x + with_sign(tol1, s->d));
  return 1;
}

double orcopt_brent(double xmin, double xguess, double xmax, double xtol, double *fx, double *f2x, void *params,
                    double (*target_funk)(void *, double)) {
  brent_state s;
  double ax, cx, fa, fb, fc, fxmin, fxmax, eps, xopt;
  int converged, iter_num = 0;
  (void)f2x;
  memset(&s, 0, sizeof s);
  if (xguess < xmin) xguess = xmin;
  if (xguess > xmax) xguess = xmax;
  eps = xguess > 0 ? xguess * xtol * 50.0 : 2. * xtol;
  ax = xguess - eps; if (ax < xmin) ax = xmin;
  cx = xguess + eps; if (cx > xmax) cx = xmax;
  fa = target_funk(params, ax);
  fb = target_funk(params, xguess);
  fc = target_funk(params, cx);
  fxmin = target_funk(params, xmin);
  fxmax = target_funk(params, xmax);
  if (fa < fb || fc < fb) { fa = fxmin; fc = fxmax; ax = xmin; cx = xmax; }
  s.tol = xtol;
  s.a = ax < cx ? ax : cx;
  s.b = ax > cx ? ax : cx;
  s.startx = s.x = xguess;
  s.fstartx = s.fx = fb;
  if (fa < fc) { s.w = ax; s.fw = fa; s.v = cx; s.fv = fc; } else { s.w = cx; s.fw = fc; s.v = ax; s.fv = fa; }
  converged = !brent_next(&s);
  do {  /* the original's loop runs until iter_num > ITMAX whether or not the variable has converged */
    const double fu = target_funk(params, s.u);
    if (!converged) {
      s.fu = fu;
      if (s.fu <= s.fx) {
        if (s.u >= s.x) s.a = s.x; else s.b = s.x;
        s.v = s.w; s.w = s.x; s.x = s.u;
        s.fv = s.fw; s.fw = s.fx; s.fx = s.fu;
      } else {
        if (s.u < s.x) s.a = s.u; else s.b = s.u;
        if (s.fu <= s.fw || s.w == s.x) { s.v = s.w; s.w = s.u; s.fv = s.fw; s.fw = s.fu; }
        else if (s.fu <= s.fv || s.v == s.x || s.v == s.w) { s.v = s.u; s.fv = s.fu; }
      }
      converged = !brent_next(&s);
    }
    iter_num++;
  } while (iter_num <= BRENT_ITMAX);
  xopt = (s.fx > s.fstartx) ? s.startx : s.x;
  fb = target_funk(params, xopt);
  if (fx) *fx = fb;
  return xopt;
}

/* Brent for xnum variables at once (one per partition), the target evaluating all of them in one call and reporting
 * per-variable scores: brent_opt_alt as pllmod_opt_minimize_brent_multi runs it (opt_algorithms.c:1043-1254).  The
 * target receives converged[] (xnum + 1 entries, the last one = "all converged", set by the target itself). */
int orcopt_brent_multi(unsigned int xnum, int *opt_mask, double *xmin, double *xguess, double *xmax, double xtol, double *xopt,
                       double *fx, double *f2x, void *params, double (*target_funk)(void *, double *, double *, int *), int global_range) {
  brent_state *st = calloc(xnum, sizeof(brent_state));
  double *ax = calloc(xnum, sizeof(double)), *cx = calloc(xnum, sizeof(double)), *fa = calloc(xnum, sizeof(double));
  double *fb = calloc(xnum, sizeof(double)), *fc = calloc(xnum, sizeof(double)), *fxmin = calloc(xnum, sizeof(double));
  double *fxmax = calloc(xnum, sizeof(double)), *lmin = calloc(xnum, sizeof(double)), *lmax = calloc(xnum, sizeof(double));
  double *u = calloc(xnum, sizeof(double)), *fu = calloc(xnum, sizeof(double));
  int *converged = calloc(xnum + 1, sizeof(int));
  unsigned int i;
  int iterate = 1, iter_num = 0;
  (void)f2x;
  for (i = 0; i < xnum; ++i) { lmin[i] = global_range ? *xmin : xmin[i]; lmax[i] = global_range ? *xmax : xmax[i]; }
  for (i = 0; i < xnum; ++i) {
    double eps;
    if (opt_mask && !opt_mask[i]) continue;
    if (xguess[i] < lmin[i]) xguess[i] = lmin[i];
    if (xguess[i] > lmax[i]) xguess[i] = lmax[i];
    eps = xguess[i] > 0 ? xguess[i] * xtol * 50.0 : 2. * xtol;
    ax[i] = xguess[i] - eps; if (ax[i] < lmin[i]) ax[i] = lmin[i];
    cx[i] = xguess[i] + eps; if (cx[i] > lmax[i]) cx[i] = lmax[i];
  }
  target_funk(params, ax, fa, NULL);
  target_funk(params, xguess, fb, NULL);
  target_funk(params, cx, fc, NULL);
  target_funk(params, lmin, fxmin, NULL);
  target_funk(params, lmax, fxmax, NULL);
  for (i = 0; i < xnum; ++i) {
    brent_state *s = &st[i];
    if (opt_mask && !opt_mask[i]) continue;
    if ((fa[i] < fb[i]) || (fc[i] < fb[i])) { fa[i] = fxmin[i]; fc[i] = fxmax[i]; ax[i] = lmin[i]; cx[i] = lmax[i]; }
    s->tol = xtol;
    s->a = ax[i] < cx[i] ? ax[i] : cx[i];
    s->b = ax[i] > cx[i] ? ax[i] : cx[i];
    s->startx = s->x = xguess[i];
    s->fstartx = s->fx = fb[i];
    if (fa[i] < fc[i]) { s->w = ax[i]; s->fw = fa[i]; s->v = cx[i]; s->fv = fc[i]; } else { s->w = cx[i]; s->fw = fc[i]; s->v = ax[i]; s->fv = fa[i]; }
    if (!brent_next(s)) converged[i] = 1;
  }
  while (iterate) {
    for (i = 0; i < xnum; ++i) u[i] = st[i].u;
    target_funk(params, u, fu, converged);
    iterate = !converged[xnum];
    for (i = 0; i < xnum; ++i) {
      brent_state *s = &st[i];
      if (opt_mask && !opt_mask[i]) continue;
      if (converged[i]) continue;
      s->fu = fu[i];
      if (s->fu <= s->fx) {
        if (s->u >= s->x) s->a = s->x; else s->b = s->x;
        s->v = s->w; s->w = s->x; s->x = s->u;
        s->fv = s->fw; s->fw = s->fx; s->fx = s->fu;
      } else {
        if (s->u < s->x) s->a = s->u; else s->b = s->u;
        if (s->fu <= s->fw || s->w == s->x) { s->v = s->w; s->w = s->u; s->fv = s->fw; s->fw = s->fu; }
        else if (s->fu <= s->fv || s->v == s->x || s->v == s->w) { s->v = s->u; s->fv = s->fu; }
      }
      converged[i] = !brent_next(s);
    }
    iter_num++;
    iterate &= (iter_num <= BRENT_ITMAX);
  }
  for (i = 0; i < xnum; ++i) xopt[i] = (st[i].fx > st[i].fstartx) ? st[i].startx : st[i].x;
  target_funk(params, xopt, fx, NULL);
  free(st); free(ax); free(cx); free(fa); free(fb); free(fc); free(fxmin); free(fxmax); free(lmin); free(lmax); free(u); free(fu); free(converged);
  return 1;
}
