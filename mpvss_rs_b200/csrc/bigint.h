// Small unsigned big-integer helper for the host-side scalar arithmetic of the PVSS
// phases (responses r = w - alpha*c mod order, Lagrange numerators/denominators,
// U masks).  The reference does this with num-bigint (modp.rs:180-192,
// participant.rs:255-264, 535-550, util.rs:33-64).  Little-endian u32 limbs,
// normalised (no leading zero limbs; zero is the empty vector).
#pragma once
#include <stdint.h>
#include <string.h>
#include <algorithm>
#include <vector>

namespace big {

using Int = std::vector<uint32_t>;

inline void trim(Int& a) {
  while (!a.empty() && a.back() == 0) a.pop_back();
}
inline Int from_u64(uint64_t v) {
  Int r;
  if (v) r.push_back((uint32_t)v);
  if (v >> 32) r.push_back((uint32_t)(v >> 32));
  return r;
}
inline Int from_le(const uint8_t* p, size_t nbytes) {
  Int r((nbytes + 3) / 4, 0);
  for (size_t i = 0; i < nbytes; ++i) r[i / 4] |= (uint32_t)p[i] << (8 * (i % 4));
  trim(r);
  return r;
}
inline Int from_be(const uint8_t* p, size_t nbytes) {
  Int r((nbytes + 3) / 4, 0);
  for (size_t i = 0; i < nbytes; ++i) r[i / 4] |= (uint32_t)p[nbytes - 1 - i] << (8 * (i % 4));
  trim(r);
  return r;
}
// fixed-width little-endian output (value must fit)
inline void to_le(const Int& a, uint8_t* out, size_t nbytes) {
  memset(out, 0, nbytes);
  for (size_t i = 0; i < a.size() * 4 && i < nbytes; ++i) out[i] = (uint8_t)(a[i / 4] >> (8 * (i % 4)));
}
inline void to_be(const Int& a, uint8_t* out, size_t nbytes) {
  memset(out, 0, nbytes);
  for (size_t i = 0; i < a.size() * 4 && i < nbytes; ++i) out[nbytes - 1 - i] = (uint8_t)(a[i / 4] >> (8 * (i % 4)));
}
inline size_t bit_length(const Int& a) {
  if (a.empty()) return 0;
  return 32 * (a.size() - 1) + (32 - __builtin_clz(a.back()));
}
inline bool is_zero(const Int& a) { return a.empty(); }
inline int cmp(const Int& a, const Int& b) {
  if (a.size() != b.size()) return a.size() < b.size() ? -1 : 1;
  for (size_t i = a.size(); i-- > 0;)
    if (a[i] != b[i]) return a[i] < b[i] ? -1 : 1;
  return 0;
}
inline Int add(const Int& a, const Int& b) {
  Int r(std::max(a.size(), b.size()) + 1, 0);
  uint64_t c = 0;
  for (size_t i = 0; i < r.size(); ++i) {
    c += (i < a.size() ? a[i] : 0u);
    c += (i < b.size() ? b[i] : 0u);
    r[i] = (uint32_t)c;
    c >>= 32;
  }
  trim(r);
  return r;
}
// a - b, requires a >= b
inline Int sub(const Int& a, const Int& b) {
  Int r(a.size(), 0);
  int64_t br = 0;
  for (size_t i = 0; i < a.size(); ++i) {
    int64_t d = (int64_t)a[i] - (i < b.size() ? b[i] : 0u) - br;
    br = d < 0;
    r[i] = (uint32_t)d;
  }
  trim(r);
  return r;
}
inline Int mul(const Int& a, const Int& b) {
  if (a.empty() || b.empty()) return Int();
  Int r(a.size() + b.size(), 0);
  for (size_t i = 0; i < a.size(); ++i) {
    uint64_t c = 0;
    for (size_t j = 0; j < b.size(); ++j) {
      c += (uint64_t)a[i] * b[j] + r[i + j];
      r[i + j] = (uint32_t)c;
      c >>= 32;
    }
    r[i + b.size()] = (uint32_t)c;
  }
  trim(r);
  return r;
}
inline Int shl1(const Int& a) { return add(a, a); }
inline Int shr1(const Int& a) {
  Int r(a);
  for (size_t i = 0; i < r.size(); ++i) r[i] = (r[i] >> 1) | (i + 1 < r.size() ? r[i + 1] << 31 : 0);
  trim(r);
  return r;
}
// Knuth algorithm D: a = q*b + r.  b must be non-zero.
inline void divmod(const Int& a, const Int& b, Int* q, Int* r) {
  if (cmp(a, b) < 0) {
    if (q) q->clear();
    if (r) *r = a;
    return;
  }
  if (b.size() == 1) {
    Int qq(a.size(), 0);
    uint64_t rem = 0;
    for (size_t i = a.size(); i-- > 0;) {
      uint64_t cur = rem << 32 | a[i];
      qq[i] = (uint32_t)(cur / b[0]);
      rem = cur % b[0];
    }
    trim(qq);
    if (q) *q = qq;
    if (r) *r = from_u64(rem);
    return;
  }
  int s = __builtin_clz(b.back());
  size_t n = b.size(), m = a.size() - n;
  Int v(n), u(a.size() + 1);
  for (size_t i = n; i-- > 0;) v[i] = (b[i] << s) | (s && i ? b[i - 1] >> (32 - s) : 0);
  u[a.size()] = s ? a.back() >> (32 - s) : 0;
  for (size_t i = a.size(); i-- > 0;) u[i] = (a[i] << s) | (s && i ? a[i - 1] >> (32 - s) : 0);
  Int qq(m + 1, 0);
  for (size_t j = m + 1; j-- > 0;) {
    uint64_t num = (uint64_t)u[j + n] << 32 | u[j + n - 1];
    uint64_t qhat = num / v[n - 1], rhat = num % v[n - 1];
    while (qhat >> 32 || qhat * v[n - 2] > (rhat << 32 | u[j + n - 2])) {
      --qhat;
      rhat += v[n - 1];
      if (rhat >> 32) break;
    }
    int64_t borrow = 0;
    uint64_t carry = 0;
    for (size_t i = 0; i < n; ++i) {
      uint64_t p = qhat * v[i] + carry;
      carry = p >> 32;
      int64_t t = (int64_t)u[i + j] - borrow - (uint32_t)p;
      u[i + j] = (uint32_t)t;
      borrow = t < 0;
    }
    int64_t t = (int64_t)u[j + n] - borrow - (int64_t)carry;
    u[j + n] = (uint32_t)t;
    if (t < 0) {
      --qhat;
      uint64_t c = 0;
      for (size_t i = 0; i < n; ++i) {
        c += (uint64_t)u[i + j] + v[i];
        u[i + j] = (uint32_t)c;
        c >>= 32;
      }
      u[j + n] += (uint32_t)c;
    }
    qq[j] = (uint32_t)qhat;
  }
  trim(qq);
  if (q) *q = qq;
  if (r) {
    Int rr(n);
    for (size_t i = 0; i < n; ++i) rr[i] = (u[i] >> s) | (s ? (uint32_t)((uint64_t)u[i + 1] << (32 - s)) : 0);
    trim(rr);
    *r = rr;
  }
}
inline Int mod(const Int& a, const Int& m) {
  Int r;
  divmod(a, m, nullptr, &r);
  return r;
}
inline Int mulmod(const Int& a, const Int& b, const Int& m) { return mod(mul(a, b), m); }
inline Int submod(const Int& a, const Int& b, const Int& m) {  // (a - b) mod m for a, b < m
  return cmp(a, b) >= 0 ? sub(a, b) : sub(add(a, m), b);
}
inline Int bxor(const Int& a, const Int& b) {
  Int r(std::max(a.size(), b.size()), 0);
  for (size_t i = 0; i < r.size(); ++i) r[i] = (i < a.size() ? a[i] : 0u) ^ (i < b.size() ? b[i] : 0u);
  trim(r);
  return r;
}
inline bool is_odd(const Int& a) { return !a.empty() && (a[0] & 1u); }

// a^-1 mod m by the extended Euclidean algorithm (util.rs:18-41); returns false when
// gcd(a, m) != 1.  Magnitudes only: the Bezout coefficient's sign alternates.
inline bool modinv(const Int& a, const Int& m, Int* out) {
  Int r0 = m, r1 = mod(a, m), t0, t1 = from_u64(1);
  bool neg = false;  // sign of t1 relative to (+): t1 positive at start
  while (!r1.empty()) {
    Int q, r2;
    divmod(r0, r1, &q, &r2);
    Int t2 = add(t0, mul(q, t1));
    r0 = r1; r1 = r2;
    t0 = t1; t1 = t2;
    neg = !neg;
  }
  if (!(r0.size() == 1 && r0[0] == 1)) return false;
  // t0 holds |coefficient of a|; its sign is +(-) when `neg` is true(false) after the final swap
  Int t = mod(t0, m);
  *out = neg ? t : (t.empty() ? t : sub(m, t));
  return true;
}

// a^-1 mod m for an ODD modulus by the binary extended Euclidean algorithm on fixed-width arrays: shifts and
// subtractions only, no division and no allocation inside the loop (a 2047-bit inverse in well under a
// millisecond, against several for the division-based modinv above).  Used for the one inversion at the root of a
// device-side batch inversion.  Returns false when gcd(a, m) != 1.
inline bool modinv_odd(const Int& a_in, const Int& m, Int* out) {
  if (m.empty() || !(m[0] & 1u)) return false;
  const size_t n = m.size() + 1;  // one spare limb for x + m
  auto is1 = [&](const std::vector<uint32_t>& x) {
    if (x[0] != 1) return false;
    for (size_t i = 1; i < n; ++i)
      if (x[i]) return false;
    return true;
  };
  auto is0 = [&](const std::vector<uint32_t>& x) {
    for (size_t i = 0; i < n; ++i)
      if (x[i]) return false;
    return true;
  };
  auto shr = [&](std::vector<uint32_t>& x) {
    for (size_t i = 0; i + 1 < n; ++i) x[i] = (x[i] >> 1) | (x[i + 1] << 31);
    x[n - 1] >>= 1;
  };
  auto addv = [&](std::vector<uint32_t>& x, const std::vector<uint32_t>& y) {
    uint64_t c = 0;
    for (size_t i = 0; i < n; ++i) {
      c += (uint64_t)x[i] + y[i];
      x[i] = (uint32_t)c;
      c >>= 32;
    }
  };
  auto subv = [&](std::vector<uint32_t>& x, const std::vector<uint32_t>& y) {  // x -= y, returns borrow
    uint64_t b = 0;
    for (size_t i = 0; i < n; ++i) {
      uint64_t d = (uint64_t)x[i] - y[i] - b;
      x[i] = (uint32_t)d;
      b = (d >> 32) & 1u;
    }
    return b != 0;
  };
  auto geq = [&](const std::vector<uint32_t>& x, const std::vector<uint32_t>& y) {
    for (size_t i = n; i-- > 0;)
      if (x[i] != y[i]) return x[i] > y[i];
    return true;
  };
  auto fixed = [&](const Int& v) {
    std::vector<uint32_t> r(n, 0);
    for (size_t i = 0; i < v.size() && i < n; ++i) r[i] = v[i];
    return r;
  };
  std::vector<uint32_t> M = fixed(m), u = fixed(mod(a_in, m)), v = M, x1(n, 0), x2(n, 0);
  if (is0(u)) return false;
  x1[0] = 1;
  auto halve = [&](std::vector<uint32_t>& x) {  // x / 2 mod m
    if (x[0] & 1u) addv(x, M);
    shr(x);
  };
  auto submod_v = [&](std::vector<uint32_t>& x, const std::vector<uint32_t>& y) {  // x = x - y mod m (x, y < m)
    if (subv(x, y)) addv(x, M);
  };
  while (!is1(u) && !is1(v)) {
    while (!(u[0] & 1u)) {
      shr(u);
      halve(x1);
    }
    while (!(v[0] & 1u)) {
      shr(v);
      halve(x2);
    }
    if (geq(u, v)) {
      subv(u, v);
      submod_v(x1, x2);
      if (is0(u)) return false;  // u == v > 1: common factor
    } else {
      subv(v, u);
      submod_v(x2, x1);
    }
  }
  Int r(is1(u) ? x1 : x2);
  trim(r);
  *out = r;
  return true;
}

}  // namespace big
