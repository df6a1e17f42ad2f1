// sparse_ldl.h -- small sparse symmetric-positive-definite direct solver for HOST-side one-shot setup work
// (chordal initialisation's two least-squares problems; the reference uses SuiteSparse SPQR there,
// src/DPGO_utils.cpp:385,448, which is not available in this image).
// A = L D L^T with an up-looking factorisation on the elimination tree after a reverse Cuthill-McKee
// ordering (pose graphs are chains plus loop closures, so RCM keeps the profile small).
#ifndef DPGO_B200_SPARSE_LDL_H
#define DPGO_B200_SPARSE_LDL_H

#include <algorithm>
#include <cstddef>
#include <numeric>
#include <queue>
#include <stdexcept>
#include <vector>

namespace dpgo_host {

struct Triplet {
  int r, c;
  double v;
};

class SparseLDL {
 public:
  // entries: ONE triangle of the symmetric matrix (each off-diagonal pair given once, as (r,c) or (c,r));
  // duplicates of the same position are summed
  void factor(int n, const std::vector<Triplet> &entries) {
    n_ = n;
    // --- adjacency for the ordering
    std::vector<std::vector<int>> adj((size_t)n);
    for (const Triplet &t : entries)
      if (t.r != t.c) { adj[(size_t)t.r].push_back(t.c); adj[(size_t)t.c].push_back(t.r); }
    for (auto &a : adj) { std::sort(a.begin(), a.end()); a.erase(std::unique(a.begin(), a.end()), a.end()); }
    perm_ = rcm(adj);
    inv_.assign((size_t)n, 0);
    for (int k = 0; k < n; ++k) inv_[(size_t)perm_[(size_t)k]] = k;
    // --- permuted upper triangle in CSC (column k holds rows i <= k)
    std::vector<Triplet> up;
    up.reserve(entries.size());
    for (const Triplet &t : entries) {
      int i = inv_[(size_t)t.r], j = inv_[(size_t)t.c];
      if (i > j) std::swap(i, j);
      up.push_back({i, j, t.v});
    }
    std::sort(up.begin(), up.end(), [](const Triplet &a, const Triplet &b) { return a.c != b.c ? a.c < b.c : a.r < b.r; });
    Ap_.assign((size_t)n + 1, 0);
    Ai_.clear();
    Ax_.clear();
    int cur_col = -1;
    for (size_t q = 0; q < up.size(); ++q) {
      if (!Ai_.empty() && cur_col == up[q].c && Ai_.back() == up[q].r) Ax_.back() += up[q].v;
      else { Ai_.push_back(up[q].r); Ax_.push_back(up[q].v); Ap_[(size_t)up[q].c + 1]++; cur_col = up[q].c; }
    }
    for (int k = 0; k < n; ++k) Ap_[(size_t)k + 1] += Ap_[(size_t)k];
    symbolic();
    numeric();
  }

  // x <- A^-1 x for one right-hand side (length n)
  void solve(double *x) const {
    std::vector<double> y((size_t)n_);
    for (int k = 0; k < n_; ++k) y[(size_t)k] = x[perm_[(size_t)k]];
    for (int j = 0; j < n_; ++j) {
      const double yj = y[(size_t)j];
      for (int p = Lp_[(size_t)j]; p < Lp_[(size_t)j + 1]; ++p) y[(size_t)Li_[(size_t)p]] -= Lx_[(size_t)p] * yj;
    }
    for (int j = 0; j < n_; ++j) y[(size_t)j] /= D_[(size_t)j];
    for (int j = n_ - 1; j >= 0; --j) {
      double s = y[(size_t)j];
      for (int p = Lp_[(size_t)j]; p < Lp_[(size_t)j + 1]; ++p) s -= Lx_[(size_t)p] * y[(size_t)Li_[(size_t)p]];
      y[(size_t)j] = s;
    }
    for (int k = 0; k < n_; ++k) x[perm_[(size_t)k]] = y[(size_t)k];
  }
  size_t nnzL() const { return Li_.size(); }

 private:
  int n_ = 0;
  std::vector<int> perm_, inv_, Ap_, Ai_, Lp_, Li_, parent_, lnz_;
  std::vector<double> Ax_, Lx_, D_;

  static std::vector<int> rcm(const std::vector<std::vector<int>> &adj) {
    const int n = (int)adj.size();
    std::vector<int> order;
    order.reserve((size_t)n);
    std::vector<char> seen((size_t)n, 0);
    std::vector<int> by_degree((size_t)n);
    std::iota(by_degree.begin(), by_degree.end(), 0);
    std::stable_sort(by_degree.begin(), by_degree.end(), [&](int a, int b) { return adj[(size_t)a].size() < adj[(size_t)b].size(); });
    for (int start : by_degree) {
      if (seen[(size_t)start]) continue;
      // pseudo-peripheral start: two BFS sweeps from the lowest-degree unseen vertex
      int root = start;
      for (int sweep = 0; sweep < 2; ++sweep) {
        std::vector<int> dist((size_t)n, -1);
        std::queue<int> q;
        q.push(root);
        dist[(size_t)root] = 0;
        int far = root;
        while (!q.empty()) {
          int u = q.front();
          q.pop();
          for (int v : adj[(size_t)u])
            if (dist[(size_t)v] < 0 && !seen[(size_t)v]) {
              dist[(size_t)v] = dist[(size_t)u] + 1;
              q.push(v);
              if (dist[(size_t)v] > dist[(size_t)far] ||
                  (dist[(size_t)v] == dist[(size_t)far] && adj[(size_t)v].size() < adj[(size_t)far].size()))
                far = v;
            }
        }
        root = far;
      }
      std::queue<int> q;
      q.push(root);
      seen[(size_t)root] = 1;
      while (!q.empty()) {
        int u = q.front();
        q.pop();
        order.push_back(u);
        std::vector<int> nb;
        for (int v : adj[(size_t)u])
          if (!seen[(size_t)v]) { seen[(size_t)v] = 1; nb.push_back(v); }
        std::sort(nb.begin(), nb.end(), [&](int a, int b) { return adj[(size_t)a].size() < adj[(size_t)b].size(); });
        for (int v : nb) q.push(v);
      }
    }
    std::reverse(order.begin(), order.end());
    return order;
  }

  void symbolic() {
    const int n = n_;
    parent_.assign((size_t)n, -1);
    lnz_.assign((size_t)n, 0);
    std::vector<int> flag((size_t)n);
    for (int k = 0; k < n; ++k) {
      flag[(size_t)k] = k;
      for (int p = Ap_[(size_t)k]; p < Ap_[(size_t)k + 1]; ++p) {
        int i = Ai_[(size_t)p];
        if (i < k)
          for (; flag[(size_t)i] != k; i = parent_[(size_t)i]) {
            if (parent_[(size_t)i] == -1) parent_[(size_t)i] = k;
            lnz_[(size_t)i]++;
            flag[(size_t)i] = k;
          }
      }
    }
    Lp_.assign((size_t)n + 1, 0);
    for (int k = 0; k < n; ++k) Lp_[(size_t)k + 1] = Lp_[(size_t)k] + lnz_[(size_t)k];
    Li_.assign((size_t)Lp_[(size_t)n], 0);
    Lx_.assign((size_t)Lp_[(size_t)n], 0.0);
  }

  void numeric() {
    const int n = n_;
    D_.assign((size_t)n, 0.0);
    std::vector<double> y((size_t)n, 0.0);
    std::vector<int> flag((size_t)n), pattern((size_t)n), fill((size_t)n, 0);
    for (int k = 0; k < n; ++k) {
      int top = n;
      flag[(size_t)k] = k;
      for (int p = Ap_[(size_t)k]; p < Ap_[(size_t)k + 1]; ++p) {
        int i = Ai_[(size_t)p];
        if (i > k) continue;
        y[(size_t)i] += Ax_[(size_t)p];
        int len = 0;
        for (; flag[(size_t)i] != k; i = parent_[(size_t)i]) {
          pattern[(size_t)len++] = i;
          flag[(size_t)i] = k;
        }
        while (len > 0) pattern[(size_t)--top] = pattern[(size_t)--len];
      }
      D_[(size_t)k] = y[(size_t)k];
      y[(size_t)k] = 0.0;
      for (; top < n; ++top) {
        const int i = pattern[(size_t)top];
        const double yi = y[(size_t)i];
        y[(size_t)i] = 0.0;
        const int p2 = Lp_[(size_t)i] + fill[(size_t)i];
        for (int p = Lp_[(size_t)i]; p < p2; ++p) y[(size_t)Li_[(size_t)p]] -= Lx_[(size_t)p] * yi;
        const double lki = yi / D_[(size_t)i];
        D_[(size_t)k] -= lki * yi;
        Li_[(size_t)p2] = k;
        Lx_[(size_t)p2] = lki;
        fill[(size_t)i]++;
      }
      if (!(D_[(size_t)k] > 0.0)) throw std::runtime_error("SparseLDL: matrix is not positive definite");
    }
  }
};

}  // namespace dpgo_host
#endif
