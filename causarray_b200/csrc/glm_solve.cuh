// Warp-level solution of one gene's (Jacobi-equilibrated) normal equations, shared by the
// launch-per-iteration solve kernel (glm.cu) and the persistent per-octet IRLS kernel
// (glm_irls_impl.cuh).  On entry L (kx x ld, ld = kx + 1, full matrix) holds X^T W X (+ ridge) and
// b holds X^T W z; on exit b holds the SCALED solution (multiply by sc[i]).  With l1 != nullptr the
// lasso step of the penalised IRLS is solved on the same system (see the comment inside).
// Returns true when the system is numerically singular.
#pragma once
#include "common.cuh"

namespace ca {

struct GlmSolveWork {
  double *L, *sc, *b, *invd, *Gc, *b0, *t2, *zf;
};

__host__ __device__ inline size_t glm_solve_work_doubles(int kx, bool lasso) {
  return (size_t)kx * (kx + 1) + 3 * kx + (lasso ? (size_t)kx * (kx + 1) + 3 * kx : 0);
}

// gc_elsewhere: the lasso refit's copy of the Gram matrix lives outside `base` (the persistent IRLS kernel
// keeps it in an L2-resident global scratch so that two CTAs fit on an SM); `base` then holds
// glm_solve_work_doubles(kx, true) - kx (kx + 1) doubles.
__device__ __forceinline__ GlmSolveWork glm_solve_carve(double* base, int kx, double* gc_elsewhere = nullptr) {
  const int ld = kx + 1;
  GlmSolveWork w;
  w.L = base;
  w.sc = w.L + kx * ld;
  w.b = w.sc + kx;
  w.invd = w.b + kx;   // reciprocals of the Cholesky diagonal
  w.Gc = gc_elsewhere ? gc_elsewhere : w.invd + kx;  // lasso refit: copy of the scaled Gram matrix (the factorisation is in place)
  w.b0 = gc_elsewhere ? w.invd + kx : w.Gc + kx * ld;
  w.t2 = w.b0 + kx;
  w.zf = w.t2 + kx;    // 1.0 where the coefficient is held at zero
  return w;
}

static __device__ __noinline__ bool glm_solve_system(const GlmSolveWork& W, int kx, const double* __restrict__ l1, int lane) {
  const int ld = kx + 1;
  double* L = W.L;
  double* sc = W.sc;
  double* b = W.b;
  double* invd = W.invd;
  double* Gc = W.Gc;
  double* b0 = W.b0;
  double* t2 = W.t2;
  double* zf = W.zf;
  __syncwarp();
  bool bad = false;
  for (int i = lane; i < kx; i += 32) {
    const double dgn = L[i * ld + i];
    sc[i] = dgn > 0.0 ? rsqrt(dgn) : 0.0;
    if (!(dgn > 0.0)) bad = true;
  }
  __syncwarp();
  // Jacobi scaling, one matrix row per lane
  for (int i = lane; i < kx; i += 32) {
    const double si = sc[i];
    for (int j = 0; j <= i; ++j) L[i * ld + j] *= si * sc[j];
    b[i] *= si;
    if (l1) {
      for (int j = 0; j <= i; ++j) Gc[i * ld + j] = L[i * ld + j];
      b0[i] = b[i];
    }
  }
  __syncwarp();
  if (l1) {  // mirror: the coordinate descent below reads full rows
    for (int i = lane; i < kx; i += 32)
      for (int j = 0; j < i; ++j) Gc[j * ld + i] = Gc[i * ld + j];
    __syncwarp();
  }
  // Right-looking Cholesky of the lower triangle, one matrix ROW per lane (rows lane and lane + 32):
  // the trailing update of row i is a run of FMAs over its columns with no index arithmetic (the
  // first version walked the trailing block as a flat index with a division and a modulo per
  // element and took 80 us per 26 x 26 system -- most of an IRLS tail iteration).
  // ld = kx + 1 is odd, so the row-per-lane accesses are bank-conflict free.
  auto factor = [&]() {
    bool sing = false;
    for (int j = 0; j < kx; ++j) {
      const double djj = L[j * ld + j];
      if (!(djj > 1e-13)) sing = true;
      const double inv = rsqrt(djj > 1e-300 ? djj : 1e-300);
      if (lane == 0) invd[j] = inv;
      for (int i = lane; i < kx; i += 32)
        if (i >= j) L[i * ld + j] *= inv;
      __syncwarp();
      for (int i = lane; i < kx; i += 32)
        if (i > j) {
          const double lij = L[i * ld + j];
          double* Li = L + i * ld;
          const double* Lj = L + j;
          int c = j + 1;
          for (; c + 3 <= i; c += 4) {  // four independent updates in flight
            const double a0 = Lj[c * ld], a1 = Lj[(c + 1) * ld], a2 = Lj[(c + 2) * ld], a3 = Lj[(c + 3) * ld];
            const double r0 = Li[c], r1 = Li[c + 1], r2 = Li[c + 2], r3 = Li[c + 3];
            Li[c] = fma(-lij, a0, r0);
            Li[c + 1] = fma(-lij, a1, r1);
            Li[c + 2] = fma(-lij, a2, r2);
            Li[c + 3] = fma(-lij, a3, r3);
          }
          for (; c <= i; ++c) Li[c] = fma(-lij, Lj[c * ld], Li[c]);
        }
      __syncwarp();
    }
    return sing;
  };
  if (factor()) bad = true;
  bad = __any_sync(0xffffffffu, bad);
  if (bad) return true;
  // column-oriented forward / backward substitution: lane i owns v[i]; the diagonal enters through
  // its stored reciprocal (after the scaled column the diagonal entry of L is djj * inv = 1 / inv)
  auto substitute = [&](double* v) {
    for (int j = 0; j < kx; ++j) {
      const double yj = v[j] * invd[j];
      __syncwarp();
      for (int i = lane; i < kx; i += 32) {
        if (i > j)
          v[i] = fma(-L[i * ld + j], yj, v[i]);
        else if (i == j)
          v[i] = yj;
      }
      __syncwarp();
    }
    for (int j = kx - 1; j >= 0; --j) {
      const double xj = v[j] * invd[j];
      __syncwarp();
      for (int i = lane; i < kx; i += 32) {
        if (i < j)
          v[i] = fma(-L[j * ld + i], xj, v[i]);
        else if (i == j)
          v[i] = xj;
      }
      __syncwarp();
    }
  };
  substitute(b);
  if (l1) {
    // Lasso step of the penalised IRLS (glmnet form): minimise the quadratic model
    //   1/2 b^T G b - r^T b + sum_k l1_k |b_k|
    // on the (scaled) kx x kx system.  Its fixed point over the IRLS iterations is the minimiser of
    // -loglike / n + sum_k alpha_k |beta_k|, the objective of statsmodels' fit_regularized
    // (elastic net with L1_wt = 1) that causarray/gcate_glm.py:209 falls back to.
    // Active-set rounds: with a sign pattern s (0 = coefficient held at zero) the optimality system is
    //   G_AA b_A = r_A - l1_A s_A,  b_Z = 0;
    // it is solved with the rows / columns of Z replaced by the identity, then checked: an active
    // coefficient whose sign came out wrong joins Z, a zero coefficient whose gradient exceeds its
    // weight rejoins A.  The unpenalised solution supplies the first pattern; one or two rounds
    // settle it.  Fallback (pattern still moving after 6 rounds): cyclic coordinate descent.
    double sg[2] = {0.0, 0.0};   // pattern of the (up to two) coordinates this lane owns
    for (int i = lane, q = 0; i < kx; i += 32, ++q) {
      const double lam = l1[i] * sc[i];
      sg[q] = b[i] > 0.0 ? 1.0 : (b[i] < 0.0 ? -1.0 : 0.0);
      if (lam == 0.0) sg[q] = 2.0;  // unpenalised coordinate: always active, no sign constraint
    }
    bool settled = false;
    for (int round = 0; round < 6 && !settled; ++round) {
      for (int i = lane, q = 0; i < kx; i += 32, ++q) {
        const double lam = l1[i] * sc[i];
        zf[i] = sg[q] == 0.0 ? 1.0 : 0.0;
        t2[i] = sg[q] == 0.0 ? 0.0 : b0[i] - (sg[q] == 2.0 ? 0.0 : lam * sg[q]);
      }
      __syncwarp();
      for (int i = lane; i < kx; i += 32)
        for (int j = 0; j <= i; ++j)
          L[i * ld + j] = (zf[i] != 0.0 || zf[j] != 0.0) ? (i == j ? 1.0 : 0.0) : Gc[i * ld + j];
      __syncwarp();
      factor();
      substitute(t2);
      bool chg = false;
      for (int i = lane, q = 0; i < kx; i += 32, ++q) {
        const double lam = l1[i] * sc[i];
        if (sg[q] == 2.0) continue;
        if (sg[q] != 0.0) {
          if (!((t2[i] > 0.0 && sg[q] > 0.0) || (t2[i] < 0.0 && sg[q] < 0.0))) {
            sg[q] = 0.0;
            chg = true;
          }
        } else {
          double gdot = 0.0;
          for (int j = 0; j < kx; ++j) gdot += Gc[i * ld + j] * t2[j];
          const double res = b0[i] - gdot;
          if (fabs(res) > lam * (1.0 + 1e-12)) {
            sg[q] = res > 0.0 ? 1.0 : -1.0;
            chg = true;
          }
        }
      }
      settled = !__any_sync(0xffffffffu, chg);
    }
    for (int i = lane; i < kx; i += 32) b[i] = t2[i];
    __syncwarp();
    if (!settled) {
      for (int sweep = 0; sweep < 2000; ++sweep) {
        double maxchg = 0.0, maxabs = 0.0;
        for (int k = 0; k < kx; ++k) {
          double part = 0.0;
          for (int j = lane; j < kx; j += 32) part += Gc[k * ld + j] * b[j];
          part = warp_sum(part);
          const double gkk = Gc[k * ld + k], old = b[k];
          const double r = b0[k] - (part - gkk * old);
          const double lam = l1[k] * sc[k];
          const double mag = fabs(r) - lam;
          const double nw = mag > 0.0 ? (r > 0.0 ? mag : -mag) / gkk : 0.0;
          __syncwarp();
          if (lane == 0) b[k] = nw;
          __syncwarp();
          maxchg = fmax(maxchg, fabs(nw - old));
          maxabs = fmax(maxabs, fabs(nw));
        }
        if (maxchg <= 1e-11 * (1.0 + maxabs)) break;
      }
    }
  }
  return false;
}

}  // namespace ca
