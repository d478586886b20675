// Device-resident state behind ca_gcate_t (shared with glm.cu / dr.cu so the
// raw counts uploaded once can be reused by the GLM and AIPW stages).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>
#include "devbuf.h"

using ca::DevBuf;

struct ca_gcate {
  int n, p, d, r, k, kpad, family;
  double thres;
  long ldp, ldn;
  cudaStream_t st;
  DevBuf<double> Yn, Ynt, nu, inv_nu, log_nu, colparm, tys_row, tys_col, tys_total, sparsity;
  DevBuf<double> A, B, Aprev, Bprev;
  DevBuf<double> P1, P2, lam, alpha_gene, l1part;
  int d1 = 0, r2 = 0, a = 0;
  bool has_P1 = false, has_P2 = false, has_lam = false, has_alpha_gene = false;
  DevBuf<uint8_t> gene_active, cell_active;
  DevBuf<int> list_c, list_g, list_clip, counts;  // counts: [0] cells, [1] genes, [2] clip, [3] n_act_c, [4] n_act_g
  DevBuf<double> ell_c, ell_g, ellfin_c, ellfin_g, Graw_c, Graw_g, g_c, g_g, f0_c, f0_g, ng_c, ng_g, W, scal;
  DevBuf<unsigned long long> evals;  // [0] cell evals, [1] gene evals, [2]/[3] cell / gene tile passes x 8
  DevBuf<uint8_t> ls_pend;           // two-phase line search scratch, max(n, p) entries each
  DevBuf<double> ls_f01, ls_e01;
  DevBuf<int> ls_list2;
  DevBuf<int> neval_g;               // candidates each gene needed in its last line search (list ordering key)
  int ldG = 64;
  // gene sharding over ranks (one process per GPU): this handle holds p of p_total genes
  void* comm = nullptr;
  void* peer = nullptr;              // ca::PeerArena: small reductions over NVLink peer memory (comm.h), or null
  int world = 1, rank = 0;
  long p_total = 0;                  // 0: not sharded (= p)
  DevBuf<double> ls_Xc, ls_t, ls_ell;
  DevBuf<int> ls_ne;
  int any_pois = 0;                  // some nu exceeds thres (the NB kernels then keep their Poisson switch)
  bool counts_loaded = false, factors_set = false;
  // stage-2 one-hot gene line search (search_onehot.cu): host copy of the observed columns of A,
  // treated cells in group order, their normalised counts and linear predictors
  std::vector<double> hostX;         // n x d (first d columns of A as uploaded by set_factors)
  bool s2_checked = false, s2_ok = false;
  int s2_ntr = 0, s2_a = 0;
  long s2_ld = 0;
  std::vector<int> s2_cells_host;
  DevBuf<int> s2_cells, s2_pos;
  DevBuf<signed char> s2_grp;
  DevBuf<double> s2_Ytr, s2_Th;
  // raw counts kept for the GLM / AIPW handles that share them
  DevBuf<double> Yraw, Yraw_t;
  bool raw_loaded = false;
};
