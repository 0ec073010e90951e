// Addition chains for the small-integer power acc^i of the Horner step (host side only).
//
// X_i = prod_j C_j^(i^j) is evaluated as acc <- acc^i * C_j (participant.rs:207-215, 423-434 compute the
// same element with t full exponentiations).  Raising to the integer i is a chain of modular products;
// a squaring costs the same as a multiplication in the fused Montgomery loop, so
// the chain LENGTH is what counts and the fixed 2-bit windows of round 1 (16.0 products for a 12-bit i)
// are replaced by a short addition chain per position:
//   * i <= TREE_LIMIT: the path to i in Knuth's power tree (TAOCP 4.6.3) -- 13.4 products on average
//     for i <= 4096, optimal for i < 77, and always a star chain (each element = previous + earlier);
//   * larger i: left-to-right sliding windows of 3 bits.
// A chain is lowered to a list of ops for the kernel (modp::hchain_body).  The accumulator lives in
// registers; chain elements that are needed again live in SLOTS shared-memory slots per lane group,
// assigned here by a linear scan over the elements' last uses.  One op:
//     if (save) slot[save-1] <- acc;  if (a) acc <- slot[a-1];  acc <- acc * B(b)
// with B(b) = slot[b] for b < SLOTS, the Montgomery one for b = SLOTS (padding) and C_j for b = SLOTS + 1
// (the last op of every Horner step).  Encoding: bits 0-3 b, 4-7 a, 8-11 save.
#pragma once
#include <stdint.h>
#include <algorithm>
#include <vector>

namespace modp_chain {

constexpr int SLOTS = 6;          // chain slots per lane group (the allocator needs at most 5 up to 2^31)
constexpr int OPS_MAX = 48;       // ops per Horner step, the product with C_j included
constexpr uint32_t B_ONE = SLOTS, B_CJ = SLOTS + 1;  // operand codes behind the chain slots
constexpr uint32_t TREE_LIMIT = 1u << 17;

struct Step {
  uint32_t a, b;  // element k (1-based step index) = element a + element b
};
struct Chain {
  std::vector<uint32_t> value;  // value[0] = 1
  std::vector<Step> steps;      // steps[k-1] builds value[k]
};

// parent[n] = the node n hangs below in the power tree (parent[1] = 0); covers 1..limit
struct PowerTree {
  std::vector<uint32_t> parent;
  uint32_t limit = 0;
  void build(uint32_t n) {
    if (n <= limit) return;
    n = std::min<uint32_t>(std::max<uint32_t>(n, 1024), TREE_LIMIT);
    parent.assign((size_t)n + 1, 0);
    std::vector<uint8_t> seen((size_t)n + 1, 0);
    std::vector<uint32_t> level{1}, next, path;
    seen[1] = 1;
    while (!level.empty()) {
      next.clear();
      for (uint32_t v : level) {
        path.clear();
        for (uint32_t m = v; m; m = parent[m]) path.push_back(m);
        for (size_t k = path.size(); k-- > 0;) {  // root first
          uint64_t s = (uint64_t)v + path[k];
          if (s <= n && !seen[s]) {
            seen[s] = 1;
            parent[s] = v;
            next.push_back((uint32_t)s);
          }
        }
      }
      level.swap(next);
    }
    limit = n;
  }
};

inline Chain tree_chain(uint32_t p, const PowerTree& T) {
  Chain c;
  for (uint32_t m = p; m; m = T.parent[m]) c.value.push_back(m);
  std::reverse(c.value.begin(), c.value.end());
  for (size_t k = 1; k < c.value.size(); ++k) {
    uint32_t d = c.value[k] - c.value[k - 1];
    uint32_t j = (uint32_t)(std::find(c.value.begin(), c.value.begin() + k, d) - c.value.begin());
    c.steps.push_back({(uint32_t)k - 1, j});
  }
  return c;
}

inline Chain window_chain(uint32_t p) {
  constexpr int W = 3;
  struct Win { uint32_t v; int len; };
  std::vector<Win> wins;
  int n = 32 - __builtin_clz(p);
  for (int k = n - 1; k >= 0;) {
    if (!((p >> k) & 1u)) {
      wins.push_back({0, 1});
      --k;
      continue;
    }
    int l = std::min(W, k + 1);
    while (!((p >> (k - l + 1)) & 1u)) --l;
    wins.push_back({(p >> (k - l + 1)) & ((1u << l) - 1u), l});
    k -= l;
  }
  uint32_t mx = 0;
  for (const Win& w : wins) mx = std::max(mx, w.v);
  Chain c;
  c.value.push_back(1);
  uint32_t idx[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // element index of the odd table entries
  auto add = [&](uint32_t a, uint32_t b) {
    c.value.push_back(c.value[a] + c.value[b]);
    c.steps.push_back({a, b});
    return (uint32_t)c.value.size() - 1;
  };
  if (mx > 1) {
    uint32_t two = add(0, 0), prev = 0;
    for (uint32_t odd = 3; odd <= mx; odd += 2) idx[odd] = prev = add(prev, two);
  }
  uint32_t cur = idx[wins[0].v];
  for (size_t w = 1; w < wins.size(); ++w) {
    for (int r = 0; r < wins[w].len; ++r) cur = add(cur, cur);
    if (wins[w].v) cur = add(cur, idx[wins[w].v]);
  }
  return c;
}

inline Chain chain_for(uint32_t p, const PowerTree& T) { return p <= T.limit ? tree_chain(p, T) : window_chain(p); }

// Lower a chain to kernel ops (without the final C_j op).  Returns false if it needs more than SLOTS slots.
// *sqr / *mul receive the number of doublings / general products.
inline bool lower(const Chain& c, std::vector<uint16_t>& ops, uint32_t* sqr, uint32_t* mul) {
  const size_t L = c.value.size();
  std::vector<int> last(L, -1), slot(L, -1);
  for (size_t k = 1; k < L; ++k) {
    const Step& s = c.steps[k - 1];
    last[s.b] = std::max(last[s.b], (int)k);
    if (s.a != k - 1) last[s.a] = std::max(last[s.a], (int)k);
  }
  uint32_t free_mask = (1u << SLOTS) - 1u;
  ops.clear();
  *sqr = *mul = 0;
  for (size_t k = 1; k < L; ++k) {
    const Step& s = c.steps[k - 1];
    uint32_t save = 0;
    const size_t e = k - 1;  // the element the accumulator holds on entry
    if (last[e] >= (int)k && slot[e] < 0) {
      if (!free_mask) return false;
      slot[e] = __builtin_ctz(free_mask);
      free_mask &= free_mask - 1;
      save = (uint32_t)slot[e] + 1;
    }
    if ((s.a != e && slot[s.a] < 0) || slot[s.b] < 0) return false;
    const uint32_t a = s.a == e ? 0u : (uint32_t)slot[s.a] + 1;
    ops.push_back((uint16_t)((save << 8) | (a << 4) | (uint32_t)slot[s.b]));
    if (s.a == s.b) ++*sqr; else ++*mul;
    for (size_t x = 0; x < L; ++x)
      if (last[x] == (int)k && slot[x] >= 0) {
        free_mask |= 1u << slot[x];
        slot[x] = -1;
        last[x] = -2;
      }
  }
  return true;
}

// Ops of one Horner step for position p: the chain, then acc * C_j.  Falls back to sliding windows if
// the tree chain does not fit the slots or OPS_MAX (never happens below 2^31; kept as a guard).
inline bool step_ops(uint32_t p, const PowerTree& T, std::vector<uint16_t>& ops, uint32_t* sqr, uint32_t* mul) {
  bool ok = lower(chain_for(p, T), ops, sqr, mul);
  if (!ok || ops.size() + 1 > (size_t)OPS_MAX) ok = lower(window_chain(p), ops, sqr, mul);
  if (!ok || ops.size() + 1 > (size_t)OPS_MAX) return false;
  ops.push_back((uint16_t)B_CJ);
  ++*mul;
  return true;
}

}  // namespace modp_chain
