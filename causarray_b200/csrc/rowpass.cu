// Host-side dispatch of the row-pass / line-search kernels (rowpass_impl.cuh) over
// KS = kpad / 4.
#include "rowpass_impl.cuh"

namespace ca {

static int pick_tc(int kpad) {
  // two staged chunks of tc x (kpad + 4) doubles; keep them under ~64 KB so that three CTAs
  // fit on an SM for kpad <= 28.
  int tc = CA_TC;
  while (tc > 16 && 2 * (size_t)tc * (kpad + 4 + 8) * 8 > 88 * 1024) tc >>= 1;
  return tc;
}

static int make_cfg(const RowProblem& rp, PassCfg& P) {
  CA_REQUIRE(rp.kpad % 8 == 4 && rp.kpad <= 68, "kpad must be congruent to 4 mod 8 and at most 68");
  CA_REQUIRE(rp.ldY % 8 == 0, "ldY must be a multiple of 8");
  P.Y = rp.Y;
  P.ldY = rp.ldY;
  P.ncols = rp.ncols;
  P.M = rp.M;
  P.nu = rp.nu;
  P.inv_nu = rp.inv_nu;
  P.log_nu = rp.log_nu;
  P.colparm = (!rp.nu_per_row && rp.family == CA_FAMILY_NB) ? reinterpret_cast<const double4*>(rp.colparm) : nullptr;
  CA_REQUIRE(rp.nu_per_row || rp.family != CA_FAMILY_NB || rp.colparm, "column parameters missing");
  P.nu_per_row = rp.nu_per_row;
  P.any_pois = rp.any_pois;
  P.family = rp.family;
  P.thres = rp.thres_disp;
  P.tc = pick_tc(rp.kpad);
  P.fm_tables = fastmath_tables();
  CA_REQUIRE(P.fm_tables, "fastmath tables unavailable");
  P.clip = nb_clip_constants();
  P.theta_out = rp.theta_out;
  P.theta_pos = rp.theta_pos;
  P.theta_ld = rp.theta_ld;
  return 0;
}

#define CA_DISPATCH_KS(ks, CALL)      \
  switch (ks) {                       \
    case 1: return CALL(1);           \
    case 3: return CALL(3);           \
    case 5: return CALL(5);           \
    case 7: return CALL(7);           \
    case 9: return CALL(9);           \
    case 11: return CALL(11);         \
    case 13: return CALL(13);         \
    case 15: return CALL(15);         \
    case 17: return CALL(17);         \
  }

int launch_rowpass(const RowProblem& rp, const double* X, const int* rowlist, const int* count, int nrows,
                   bool want_grad, int gc_lo, int gc_n, double* ell_out, double* G_out, int ldG,
                   cudaStream_t st) {
  if (nrows <= 0) return 0;
  PassCfg P;
  if (int rc = make_cfg(rp, P)) return rc;
  if (want_grad) {
    CA_REQUIRE(gc_n <= 64, "gradient block wider than 64 columns");
    CA_REQUIRE(ldG >= gc_n, "ldG too small");
  }
#define CALL_RP(K) launch_rowpass_ks<K>(P, X, rowlist, count, nrows, want_grad, gc_lo, gc_n, ell_out, G_out, ldG, st)
  CA_DISPATCH_KS(rp.kpad / 4, CALL_RP)
#undef CALL_RP
  set_error("unsupported kpad %d", rp.kpad);
  return -2;
}

static int dispatch_search(int ks, const PassCfg& P, const SearchArgs& A, cudaStream_t st) {
#define CALL_LS(K) launch_search_ks<K>(P, A, st)
  CA_DISPATCH_KS(ks, CALL_LS)
#undef CALL_LS
  set_error("unsupported kpad %d", 4 * ks);
  return -2;
}
static int dispatch_search_multi(int ks, const PassCfg& P, const SearchArgs& A, cudaStream_t st) {
#define CALL_LM(K) launch_search_multi_ks<K>(P, A, st)
  CA_DISPATCH_KS(ks, CALL_LM)
#undef CALL_LM
  set_error("unsupported kpad %d", 4 * ks);
  return -2;
}

int launch_search(const RowProblem& rp, const SearchLaunch& s, cudaStream_t st) {
  if (s.nrows <= 0) return 0;
  PassCfg P;
  if (int rc = make_cfg(rp, P)) return rc;
  SearchArgs A;
  A.X = s.X;
  A.g = s.g;
  A.f0 = s.f0;
  A.normg = s.normg;
  A.tys = s.tys;
  A.lam = s.lam;
  A.ldlam = s.ldlam;
  A.prox_lo = s.prox_lo;
  A.prox_hi = s.prox_hi;
  A.mv_lo = s.mv_lo;
  A.mv_hi = s.mv_hi;
  A.alpha_row = s.alpha_row;
  A.alpha = s.alpha;
  A.beta = s.beta;
  A.tol = s.tol;
  A.max_iters = s.max_iters;
  A.rowlist = s.rowlist;
  A.count = s.count;
  A.nrows = s.nrows;
  A.ell_fin = s.ell_fin;
  A.neval = s.neval;
  A.eval_total = s.eval_total;
  A.pend = s.pend;
  A.f01s = s.f01s;
  A.e01s = s.e01s;
  if (s.pend) {
    CA_REQUIRE(s.f01s && s.e01s && s.list2 && s.count2, "two-phase search needs all of its buffers");
    CA_CUDA(cudaMemsetAsync(s.pend, 0, (size_t)s.nrows, st));
  }
  int rc = dispatch_search(rp.kpad / 4, P, A, st);
  if (rc || !s.pend) return rc;
  // phase B over the rows phase A left pending
  if (s.neval ? launch_compact_by_key(s.pend, s.neval, s.nrows, s.list2, s.count2, st)
              : launch_compact(s.pend, s.nrows, s.list2, s.count2, st))
    return -1;
  A.rowlist = s.list2;
  A.count = s.count2;
  A.pend = nullptr;
  return dispatch_search_multi(rp.kpad / 4, P, A, st);
}

}  // namespace ca
