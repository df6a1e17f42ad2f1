// cpu_port.cpp -- C++ restatement of the reference's CPU path for ONE QuadraticOptimizer::optimize() call.
//
// TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/dpgo_oracle.py header): timed by bench.py's cpu_baseline leg
// and `--impl reference`, validated against the NumPy oracle in tests/.  It is labelled "CPU restatement of the
// reference path", never "reference": the reference binary cannot be built in this image (Eigen, SuiteSparse,
// ROPTLIB absent).  It mirrors the reference's algorithmic STRUCTURE and operation count:
//   * X*Q as Eigen evaluates dense * sparse(RowMajor): scalar CSR, int32 indices, one pass per product
//     (ref src/QuadraticProblem.cpp:59,65,72);
//   * the same (10 + j) products per RTR call: f + RieGradNorm before (:36-37), RieGradNorm in trustRegion (:65),
//     f + Grad in ROPTLIB's Run, one Hessian product per tCG iteration, f(x2) + Hessian(eta) + Grad(x2) after,
//     f + RieGradNorm at the end (:52-53);
//   * the preconditioner as an exact sparse factorisation of Q + 0.1 I with r right-hand sides per apply
//     (ref :37-41,75-87; CHOLMOD there, an RCM-ordered up-looking LDL^T here);
//   * QF retraction, tangent projection, ROPTLIB tCG recurrences.
// Single thread by default (the reference's default: ENABLE_OPENMP OFF, Eigen's product is serial);
// threads > 1 parallelises the products over output columns (OpenMP); the triangular solves stay sequential.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "sparse_ldl_oracle.h"

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

struct Problem {
  int n, d, r, dh, N, threads;
  std::vector<int> rowptr, colind;
  std::vector<double> val;
  std::vector<double> G;
  dpgo_oracle::SparseLDL ldl;
  long spmv_count = 0, solve_count = 0;
};

// Out = X * Q (+ G).  Q symmetric, so column c of Q is row c: Out(:,c) = sum_k X(:, col_k) * val_k over row c
// -- the same scalar work Eigen does, written gather-style so that threads never write the same column.
void xq(Problem &p, const double *X, double *out, bool addG) {
  const int r = p.r, N = p.N;
  p.spmv_count++;
#pragma omp parallel for num_threads(p.threads) schedule(static) if (p.threads > 1)
  for (int c = 0; c < N; ++c) {
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int q = p.rowptr[c]; q < p.rowptr[c + 1]; ++q) {
      const double v = p.val[q];
      const double *x = X + (size_t)p.colind[q] * r;
      for (int a = 0; a < r; ++a) acc[a] += x[a] * v;
    }
    double *o = out + (size_t)c * r;
    if (addG) for (int a = 0; a < r; ++a) o[a] = acc[a] + p.G[(size_t)c * r + a];
    else for (int a = 0; a < r; ++a) o[a] = acc[a];
  }
}

double dot(const Problem &p, const double *a, const double *b) {
  double s = 0;
  const size_t L = (size_t)p.r * p.N;
  for (size_t k = 0; k < L; ++k) s += a[k] * b[k];
  return s;
}

// tangent projection at X of Z (in place on out): per pose Z_Y - Y sym(Y^T Z_Y)
void project(const Problem &p, const double *X, const double *Z, double *out) {
  const int r = p.r, d = p.d, dh = p.dh;
  for (int i = 0; i < p.n; ++i) {
    const double *Y = X + (size_t)i * dh * r, *Zi = Z + (size_t)i * dh * r;
    double *O = out + (size_t)i * dh * r;
    double S[3][3];
    for (int a = 0; a < d; ++a)
      for (int b = 0; b < d; ++b) {
        double s = 0;
        for (int k = 0; k < r; ++k) s += Y[a * r + k] * Zi[b * r + k];
        S[a][b] = s;
      }
    for (int b = 0; b < d; ++b)
      for (int k = 0; k < r; ++k) {
        double s = Zi[b * r + k];
        for (int a = 0; a < d; ++a) s -= Y[a * r + k] * 0.5 * (S[a][b] + S[b][a]);
        O[b * r + k] = s;
      }
    for (int k = 0; k < r; ++k) O[d * r + k] = Zi[d * r + k];
  }
}

double f_of(Problem &p, const double *X, double *tmp) {
  xq(p, X, tmp, false);
  return 0.5 * dot(p, tmp, X) + dot(p, X, p.G.data());
}

void egrad(Problem &p, const double *X, double *EG) { xq(p, X, EG, true); }

double rgradnorm(Problem &p, const double *X, double *tmp, double *tmp2) {
  egrad(p, X, tmp);
  project(p, X, tmp, tmp2);
  return std::sqrt(dot(p, tmp2, tmp2));
}

// Riemannian Hessian-vector product (EucHessianEta + Stiefel::EucHvToHv + projection)
void rhess(Problem &p, const double *X, const double *EG, const double *V, double *out, double *tmp) {
  const int r = p.r, d = p.d, dh = p.dh;
  xq(p, V, tmp, false);
  for (int i = 0; i < p.n; ++i) {
    const double *Y = X + (size_t)i * dh * r, *E = EG + (size_t)i * dh * r, *Vi = V + (size_t)i * dh * r;
    double *T = tmp + (size_t)i * dh * r;
    double S[3][3];
    for (int a = 0; a < d; ++a)
      for (int b = 0; b < d; ++b) {
        double s = 0;
        for (int k = 0; k < r; ++k) s += Y[a * r + k] * E[b * r + k];
        S[a][b] = s;
      }
    for (int b = 0; b < d; ++b)
      for (int k = 0; k < r; ++k) {
        double s = 0;
        for (int a = 0; a < d; ++a) s += Vi[a * r + k] * 0.5 * (S[a][b] + S[b][a]);
        T[b * r + k] -= s;
      }
  }
  project(p, X, tmp, out);
}

// z = P_X( (Q + 0.1 I)^-1 v ): one solve per row of the r x N residual
void precondition(Problem &p, const double *X, const double *V, double *out, double *tmp) {
  const int r = p.r, N = p.N;
  p.solve_count++;
  // V is stored [column c][row a] = N right-hand-side entries interleaved by a: all r systems in one pass over L
  // (the sparse triangular solves are sequential along the elimination tree; only the products use threads)
  std::memcpy(tmp, V, sizeof(double) * (size_t)r * N);
  p.ldl.solve_multi(tmp, r);
  project(p, X, tmp, out);
}

// QF retraction (Householder-free: modified Gram-Schmidt twice, diag(R) > 0)
void retract(const Problem &p, const double *X, const double *eta, double *out) {
  const int r = p.r, d = p.d, dh = p.dh;
  for (int i = 0; i < p.n; ++i) {
    double W[3][8];
    const double *Xi = X + (size_t)i * dh * r, *Ei = eta + (size_t)i * dh * r;
    double *O = out + (size_t)i * dh * r;
    for (int c = 0; c < d; ++c)
      for (int k = 0; k < r; ++k) W[c][k] = Xi[c * r + k] + Ei[c * r + k];
    for (int pass = 0; pass < 2; ++pass)
      for (int c = 0; c < d; ++c) {
        for (int q = 0; q < c; ++q) {
          double s = 0;
          for (int k = 0; k < r; ++k) s += W[q][k] * W[c][k];
          for (int k = 0; k < r; ++k) W[c][k] -= s * W[q][k];
        }
        double nrm = 0;
        for (int k = 0; k < r; ++k) nrm += W[c][k] * W[c][k];
        nrm = std::sqrt(nrm);
        for (int k = 0; k < r; ++k) W[c][k] /= nrm;
      }
    for (int c = 0; c < d; ++c)
      for (int k = 0; k < r; ++k) O[c * r + k] = W[c][k];
    for (int k = 0; k < r; ++k) O[d * r + k] = Xi[d * r + k] + Ei[d * r + k];
  }
}

}  // namespace

extern "C" {

struct cpu_port_result {
  double f_init, gradnorm_init, f_opt, gradnorm_opt, relative_change;
  int tcg_iterations, tcg_status, outer_iterations, rejections, spmv, solves;
};

void *cpu_port_create(int n, int d, int r, const int *rowptr, const int *colind, const double *val, int threads) {
  Problem *p = new Problem();
  p->n = n; p->d = d; p->r = r; p->dh = d + 1; p->N = (d + 1) * n; p->threads = std::max(1, threads);
  p->rowptr.assign(rowptr, rowptr + p->N + 1);
  p->colind.assign(colind, colind + rowptr[p->N]);
  p->val.assign(val, val + rowptr[p->N]);
  p->G.assign((size_t)r * p->N, 0.0);
  std::vector<dpgo_oracle::Triplet> ent;
  std::vector<char> diag((size_t)p->N, 0);
  for (int i = 0; i < p->N; ++i)
    for (int q = rowptr[i]; q < rowptr[i + 1]; ++q) {
      const int c = colind[q];
      if (c < i) continue;                      // upper triangle only
      double v = val[q];
      if (c == i) { v += 0.1; diag[(size_t)i] = 1; }
      ent.push_back({i, c, v});
    }
  for (int i = 0; i < p->N; ++i)
    if (!diag[(size_t)i]) ent.push_back({i, i, 0.1});
  p->ldl.set_ordering(1, p->dh);                 // minimum degree on the pose graph (CHOLMOD uses AMD)
  p->ldl.factor(p->N, ent);                      // ref QuadraticProblem::setQ: factor Q + 0.1 I
  return p;
}

void cpu_port_destroy(void *h) { delete static_cast<Problem *>(h); }
void cpu_port_set_G(void *h, const double *G) {
  Problem *p = static_cast<Problem *>(h);
  if (G) std::memcpy(p->G.data(), G, sizeof(double) * p->G.size());
  else std::fill(p->G.begin(), p->G.end(), 0.0);
}
long cpu_port_nnzL(void *h) { return (long)static_cast<Problem *>(h)->ldl.nnzL(); }

// one optimize() call; algorithm 0 = RTR, 1 = RGD.  X col-major r x N.
int cpu_port_optimize(void *h, int algorithm, int tr_iterations, int max_inner, double tol, double radius0,
                      double rgd_step, const double *Xin, double *Xout, cpu_port_result *res) {
  Problem &p = *static_cast<Problem *>(h);
  const size_t L = (size_t)p.r * p.N;
  std::vector<double> X(Xin, Xin + L), EG(L), g(L), eta(L), rs(L), z(L), delta(L), Hd(L), X2(L), t1(L), t2(L), Heta(L);
  p.spmv_count = 0;
  p.solve_count = 0;
  std::memset(res, 0, sizeof(*res));
  res->tcg_status = -1;
  res->f_init = f_of(p, X.data(), t1.data());                       // ref :36
  res->gradnorm_init = rgradnorm(p, X.data(), t1.data(), t2.data());  // ref :37
  if (algorithm == 1) {
    egrad(p, X.data(), EG.data());
    project(p, X.data(), EG.data(), g.data());
    for (size_t k = 0; k < L; ++k) eta[k] = -rgd_step * g[k];
    retract(p, X.data(), eta.data(), X2.data());
    X.swap(X2);
    res->outer_iterations = 1;
  } else {
    const double gn0 = rgradnorm(p, X.data(), t1.data(), t2.data());  // ref :65
    if (gn0 >= tol) {
      const bool single = (tr_iterations == 1);
      double Delta = radius0;
      const double DeltaMax = single ? radius0 : 5.0 * radius0;
      int total_steps = 0, iter = 0;
      double f1 = f_of(p, X.data(), t1.data());                     // ROPTLIB Run: f(x1), Grad(x1)
      egrad(p, X.data(), EG.data());
      project(p, X.data(), EG.data(), g.data());
      double gn = std::sqrt(dot(p, g.data(), g.data()));
      while (true) {
        if (single && total_steps > 0) {                             // every Run() re-evaluates f and the gradient
          f1 = f_of(p, X.data(), t1.data());
          egrad(p, X.data(), EG.data());
          project(p, X.data(), EG.data(), g.data());
        }
        // ---- tCG
        std::fill(eta.begin(), eta.end(), 0.0);
        rs = g;
        precondition(p, X.data(), rs.data(), z.data(), t1.data());
        for (size_t k = 0; k < L; ++k) delta[k] = -z[k];
        double z_r = dot(p, z.data(), rs.data()), d_Pd = z_r, e_Pd = 0, e_Pe = 0;
        const double n0 = std::sqrt(dot(p, rs.data(), rs.data()));
        int status = 4;
        for (int j = 0; j < max_inner; ++j) {
          rhess(p, X.data(), EG.data(), delta.data(), Hd.data(), t1.data());
          res->tcg_iterations++;
          const double d_Hd = dot(p, delta.data(), Hd.data());
          const double alpha = z_r / d_Hd;
          const double e_new = e_Pe + 2 * alpha * e_Pd + alpha * alpha * d_Pd;
          if (d_Hd <= 0 || e_new >= Delta * Delta) {
            const double tau = (-e_Pd + std::sqrt(e_Pd * e_Pd + d_Pd * (Delta * Delta - e_Pe))) / d_Pd;
            for (size_t k = 0; k < L; ++k) eta[k] += tau * delta[k];
            status = (d_Hd <= 0) ? 0 : 1;
            break;
          }
          e_Pe = e_new;
          for (size_t k = 0; k < L; ++k) { eta[k] += alpha * delta[k]; rs[k] += alpha * Hd[k]; }
          const double nr = std::sqrt(dot(p, rs.data(), rs.data()));
          if (nr <= n0 * std::min(n0, 0.1)) { status = (0.1 < n0) ? 2 : 3; break; }
          precondition(p, X.data(), rs.data(), z.data(), t1.data());
          const double zr_new = dot(p, z.data(), rs.data());
          const double beta = zr_new / z_r;
          z_r = zr_new;
          for (size_t k = 0; k < L; ++k) delta[k] = -z[k] + beta * delta[k];
          e_Pd = beta * (e_Pd + alpha * d_Pd);
          d_Pd = z_r + beta * beta * d_Pd;
        }
        res->tcg_status = status;
        res->outer_iterations++;
        retract(p, X.data(), eta.data(), X2.data());
        const double f2 = f_of(p, X2.data(), t1.data());
        rhess(p, X.data(), EG.data(), eta.data(), Heta.data(), t1.data());
        const double denom = -dot(p, eta.data(), g.data()) - 0.5 * dot(p, eta.data(), Heta.data());
        const double rho = (f1 - f2) / denom;
        const bool accepted = rho > 0.1;
        if (single) {
          if (accepted) { X.swap(X2); egrad(p, X.data(), EG.data()); break; }   // Grad(x2) on acceptance
          res->rejections++;
          if (total_steps > 10) break;
          Delta /= 4;
          total_steps++;
        } else {
          if (rho < 0.25) Delta *= 0.25;
          else if (rho > 0.75 && (status == 0 || status == 1)) Delta = std::min(2 * Delta, DeltaMax);
          if (accepted) {
            X.swap(X2);
            f1 = f2;
            egrad(p, X.data(), EG.data());
            project(p, X.data(), EG.data(), g.data());
            gn = std::sqrt(dot(p, g.data(), g.data()));
          } else {
            res->rejections++;
          }
          if (gn < tol || ++iter >= tr_iterations) break;
        }
      }
    }
  }
  res->f_opt = f_of(p, X.data(), t1.data());                         // ref :52
  res->gradnorm_opt = rgradnorm(p, X.data(), t1.data(), t2.data());    // ref :53
  double ch = 0;
  for (size_t k = 0; k < L; ++k) ch += (X[k] - Xin[k]) * (X[k] - Xin[k]);
  res->relative_change = std::sqrt(ch / p.n);
  res->spmv = (int)p.spmv_count;
  res->solves = (int)p.solve_count;
  std::memcpy(Xout, X.data(), sizeof(double) * L);
  return 0;
}

}  // extern "C"
