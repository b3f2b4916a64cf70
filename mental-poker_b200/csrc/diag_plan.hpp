// Host-only plan for the prover's diagonal ciphertext products (kernel family K2 of SURVEY.md 2b;
// the E_k of the multi-exponentiation argument, SURVEY.md Appendix B.5, computed inside
// proof-essentials' `ShuffleArgument::prove`, reference call site
// src/discrete_log_cards/mod.rs:409-415).
//
//   E_k = sum_{i=1..m, j=0..m, m+j-i=k}  <C_i, A_j>,     k = 0 .. 2m-1
//
// with C_i the i-th chunk of n shuffled ciphertexts and A_j the j-th scalar row (A_0 = a0, the
// blinding row; A_j = b_j for j >= 1).  Written as polynomials in the exponent,
//   P(X) = sum_u P_u X^u  (P_u = C_{m-u}),   S(X) = sum_v S_v X^v  (S_v = A_{v+1}),  u, v in [0, m)
// the coefficient of X^w in P(X)*S(X) is E_{w+1} without its A_0 terms -- a product of two degree
// m-1 polynomials whose "coefficient multiplication" is the bilinear map <points, scalars>.  The
// schoolbook evaluation costs m^2 length-n inner products; Karatsuba's identity
//   (P_lo + X^h P_hi)(S_lo + X^h S_hi)
//        = P_lo S_lo + X^h [ (P_lo+P_hi)(S_lo+S_hi) - P_lo S_lo - P_hi S_hi ] + X^2h P_hi S_hi
// applied on every bit of the coefficient index needs 3^ceil(log2 m) of them (2187 instead of
// 16384 at m = 128), each over a *sum of point rows* and the matching *sum of scalar rows* (Bayer
// and Groth's own suggestion for the prover, section 6 of their paper).  Group elements are
// canonical, so the E_k bytes are unchanged.
//
// A leaf is a string of ternary digits, one per index bit b: 0 = "bit clear", 1 = "bit set",
// 2 = "either" (the (lo+hi) branch).  Its point / scalar rows are the sums over the index set
//   U = { u < m : (u & mask) == val },   mask = bits with digit != 2, val = bits with digit 1.
// Its product R_leaf contributes with sign +-1 to the output coefficients reached by choosing,
// per bit (h = 2^b):  digit 0 -> offset 0 (+) or h (-);  digit 1 -> offset 2h (+) or h (-);
// digit 2 -> offset h (+).
//
// No CUDA in this header: tests/host/host_shim.cpp compiles it with g++ and
// tests/test_host_diag_plan.py checks the identity over the integers.
#pragma once
#include <stdint.h>

#include <algorithm>
#include <vector>

namespace mp {

struct DiagPlan {
  int m = 0, levels = 0;
  // leaves with a non-empty index set, heaviest first (the evaluation kernels run one block row
  // per leaf: heavy rows are scheduled first)
  std::vector<uint32_t> leaf_mask, leaf_val, leaf_weight;
  std::vector<uint32_t> single;   // single[u] = leaf whose index set is exactly {u}
  // jobs: leaf l -> job l  (<leaf points, leaf scalars>);  job nleaf + (i-1) = <C_i, A_0>, i = 1..m
  // contributions to E_k in CSR form: entries job | (negative ? 1u<<31 : 0)
  std::vector<uint32_t> row_start;  // 2m + 1
  std::vector<uint32_t> entries;
  uint32_t nleaf() const { return (uint32_t)leaf_mask.size(); }
  uint32_t njobs() const { return nleaf() + (uint32_t)m; }
};

inline DiagPlan diag_plan_build(int m) {
  DiagPlan p;
  p.m = m;
  int L = 0;
  while ((1 << L) < m) L++;
  p.levels = L;
  const uint32_t M = 1u << L;
  uint32_t n3 = 1;
  for (int b = 0; b < L; b++) n3 *= 3;
  struct Leaf { uint32_t mask, val, weight, id; };
  std::vector<Leaf> leaves;
  for (uint32_t id = 0; id < n3; id++) {
    uint32_t mask = 0, val = 0, t = id;
    for (int b = 0; b < L; b++, t /= 3) {
      uint32_t d = t % 3;
      if (d != 2) mask |= 1u << b;
      if (d == 1) val |= 1u << b;
    }
    uint32_t weight = 0;
    const uint32_t free_bits = ~mask & (M - 1);
    uint32_t sub = 0;
    do {
      if ((val | sub) < (uint32_t)m) weight++;
      sub = (sub - free_bits) & free_bits;
    } while (sub != 0);
    if (weight) leaves.push_back(Leaf{mask, val, weight, id});
  }
  std::stable_sort(leaves.begin(), leaves.end(), [](const Leaf& a, const Leaf& b) { return a.weight > b.weight; });
  p.single.assign((size_t)m, 0);
  for (uint32_t l = 0; l < leaves.size(); l++) {
    p.leaf_mask.push_back(leaves[l].mask);
    p.leaf_val.push_back(leaves[l].val);
    p.leaf_weight.push_back(leaves[l].weight);
    if (leaves[l].mask == M - 1) p.single[leaves[l].val] = l;
  }
  // contributions, bucketed by k
  std::vector<std::vector<uint32_t>> rows((size_t)2 * m);
  for (uint32_t l = 0; l < leaves.size(); l++) {
    // expand the per-bit choices: (offset, sign) pairs
    std::vector<std::pair<uint32_t, uint32_t>> cur{{0u, 0u}}, nxt;
    for (int b = 0; b < L; b++) {
      const uint32_t h = 1u << b;
      const bool fixed = (leaves[l].mask >> b) & 1, set = (leaves[l].val >> b) & 1;
      nxt.clear();
      for (auto& c : cur) {
        if (!fixed) {
          nxt.push_back({c.first + h, c.second});
        } else {
          nxt.push_back({c.first + (set ? 2 * h : 0), c.second});
          nxt.push_back({c.first + h, c.second ^ 1u});
        }
      }
      cur.swap(nxt);
    }
    for (auto& c : cur) {
      const uint32_t k = c.first + 1;  // coefficient w of P*S is E_{w+1}
      // coefficients above 2m-2 of the padded product are identically zero: their contributions cancel
      if (k < (uint32_t)(2 * m)) rows[k].push_back(l | (c.second << 31));
    }
  }
  for (int i = 1; i <= m; i++) rows[(size_t)(m - i)].push_back((uint32_t)leaves.size() + (uint32_t)(i - 1));
  p.row_start.push_back(0);
  for (auto& r : rows) {
    p.entries.insert(p.entries.end(), r.begin(), r.end());
    p.row_start.push_back((uint32_t)p.entries.size());
  }
  return p;
}

// schoolbook (pre-shifted table, one bucket set per diagonal) vs Karatsuba (independent short
// jobs): cost in bucket additions per ciphertext component, used to pick the path
inline bool diag_use_karatsuba(int m, int n, int W_table, int W_leaf, int c_leaf) {
  int L = 0;
  while ((1 << L) < m) L++;
  double leaves = 1;
  for (int b = 0; b < L; b++) leaves *= 3;
  const double school = (double)m * (m + 1) * n * W_table;
  const double kara = (leaves + m) * (double)W_leaf * ((double)n + 3.0 * (double)(1u << (c_leaf - 1))) + (double)m * (1 << L) * n;
  return kara < school;
}

}  // namespace mp
