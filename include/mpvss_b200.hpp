// mpvss_b200.hpp -- C++17 host-side mirror of the reference's operator interface for the hot path, on top of
// the C ABI in mpvss_b200.h.  Header-only; link with -lmpvss_b200.
//
// The reference is a Rust crate and this image has no Rust toolchain, so the host side a maintainer would
// write in Rust is restated here in C++ with the reference's names, argument meaning and
// error behaviour:
//
//   reference (src/)                                          here
//   ModpGroup::new() / Secp256k1Group::new() / ...  -> Arc<G>   Group<G>::create() -> std::shared_ptr<Group<G>>
//   Participant::with_arc(group), initialize()                  Participant<G>(group), initialize()
//   distribute_secret(&secret, &publickeys, threshold)          distribute_secret(secret, publickeys, threshold)
//   extract_secret_share(&box, &sk, &w) -> Option<ShareBox>     extract_secret_share(box, sk, w) -> std::optional
//   verify_share(&sharebox, &box, &publickey) -> bool           verify_share(sharebox, box, publickey)
//   verify_distribution_shares(&box) -> bool                    verify_distribution_shares(box)
//   reconstruct(&[ShareBox], &box) -> Option<BigInt>            reconstruct(shareboxes, box) -> std::optional<Bytes>
//   DistributionSharesBox<G> (sharebox.rs:75-86)                DistributionSharesBox<G>
//   ShareBox<G> (sharebox.rs:22-27)                             ShareBox<G>
//   string_to_secret / string_from_secret (lib.rs:49-57)        same names
//
// Value representation: an Element or Scalar is its fixed-width boundary encoding (`Bytes`, see
// mpvss_b200.h); the maps of DistributionSharesBox are keyed by the public key's encoding, which is what
// `element_to_bytes(pubkey)` keys the reference's HashMaps with (participant.rs:197, 257, 409).  A secret is the
// big-endian byte string of the reference's BigInt (BigUint::to_bytes_be), at most one element long.
// Additions over the reference: the randomness-injection form `distribute_secret_with(coeffs, witnesses)` (the
// reference draws both from thread_rng: polynomial.rs:36, participant.rs:223) and the batch forms
// `extract_secret_shares` / `verify_shares`.  Every group operation runs on the GPU through the C ABI; there is
// no CPU path here either.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <optional>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

#include "mpvss_b200.h"

namespace mpvss {

using Bytes = std::vector<uint8_t>;

struct Error : std::runtime_error {
  int status;
  Error(int st, const std::string& what) : std::runtime_error(what), status(st) {}
};

namespace detail {
inline Bytes from_hex(const char* s) {
  auto nib = [](char c) { return c <= '9' ? c - '0' : (c | 32) - 'a' + 10; };
  Bytes out;
  for (size_t i = 0; s[i] && s[i + 1]; i += 2) out.push_back((uint8_t)(nib(s[i]) << 4 | nib(s[i + 1])));
  return out;
}
// big-endian comparison of equal-length strings: a < b
inline bool less_be(const Bytes& a, const Bytes& b) { return std::lexicographical_compare(a.begin(), a.end(), b.begin(), b.end()); }
inline bool is_zero(const Bytes& a) { return std::all_of(a.begin(), a.end(), [](uint8_t x) { return x == 0; }); }
inline Bytes reversed(Bytes a) { std::reverse(a.begin(), a.end()); return a; }
// a - 1 for a big-endian a > 0
inline Bytes minus_one_be(Bytes a) {
  for (size_t i = a.size(); i-- > 0;) if (a[i]--) break;
  return a;
}
// a >> 1 for a big-endian a
inline Bytes half_be(Bytes a) {
  uint8_t c = 0;
  for (auto& x : a) { uint8_t n = (uint8_t)(x & 1); x = (uint8_t)(x >> 1 | c << 7); c = n; }
  return a;
}
}  // namespace detail

// ---- group tags: the three `impl Group` of the reference -----------------------------------------------------
struct ModpGroup {  // src/groups/modp.rs:44-70 (RFC 3526 group 14)
  static constexpr int id = MPVSS_GROUP_MODP;
  static constexpr size_t EB = 256, SB = 256;
  static constexpr bool scalar_big_endian = false;
  static Bytes modulus_be() {
    return detail::from_hex(
        "ffffffffffffffffc90fdaa22168c234c4c6628b80dc1cd129024e088a67cc74020bbea63b139b22514a08798e3404dd"
        "ef9519b3cd3a431b302b0a6df25f14374fe1356d6d51c245e485b576625e7ec6f44c42e9a637ed6b0bff5cb6f406b7ed"
        "ee386bfb5a899fa5ae9f24117c4b1fe649286651ece45b3dc2007cb8a163bf0598da48361c55d39a69163fa8fd24cf5f"
        "83655d23dca3ad961c62f356208552bb9ed529077096966d670c354e4abc9804f1746c08ca18217c32905e462e36ce3b"
        "e39e772c180e86039b2783a2ec07a28fb5c55df06f4c52c9de2bcbf6955817183995497cea956ae515d2261898fa0510"
        "15728e5a8aacaa68ffffffffffffffff");
  }
  static Bytes order_be() { return detail::minus_one_be(modulus_be()); }  // q - 1 (modp.rs:101-103)
};
struct Secp256k1Group {  // src/groups/secp256k1.rs:44-55
  static constexpr int id = MPVSS_GROUP_SECP256K1;
  static constexpr size_t EB = 33, SB = 32;
  static constexpr bool scalar_big_endian = true;
  static Bytes order_be() { return detail::from_hex("fffffffffffffffffffffffffffffffebaaedce6af48a03bbfd25e8cd0364141"); }
};
struct Ristretto255Group {  // src/groups/ristretto255.rs:51-63
  static constexpr int id = MPVSS_GROUP_RISTRETTO255;
  static constexpr size_t EB = 32, SB = 32;
  static constexpr bool scalar_big_endian = false;
  static Bytes order_be() { return detail::from_hex("1000000000000000000000000000000014def9dea2f79cd65812631a5cf5d3ed"); }
};

// ---- Group<G>: one CUDA context, shared immutably like the reference's Arc<G> ----------------------------------
template <class G>
class Group {
 public:
  static std::shared_ptr<Group> create(int device = 0) { return std::shared_ptr<Group>(new Group(device)); }
  ~Group() { if (ctx_) mpvss_ctx_destroy(ctx_); }
  Group(const Group&) = delete;
  Group& operator=(const Group&) = delete;
  mpvss_ctx* ctx() const { return ctx_; }

  void check(int rc) const {
    if (rc != MPVSS_OK) throw Error(rc, std::string("mpvss_b200: ") + mpvss_last_error(ctx_));
  }
  static Bytes concat(const std::vector<Bytes>& v, size_t width) {
    Bytes out;
    out.reserve(v.size() * width);
    for (const auto& e : v) {
      if (e.size() != width) throw Error(MPVSS_ERR_ARG, "mpvss_b200: value has the wrong encoded width");
      out.insert(out.end(), e.begin(), e.end());
    }
    return out;
  }
  static std::vector<Bytes> split(const Bytes& b, size_t width) {
    std::vector<Bytes> out;
    for (size_t i = 0; i + width <= b.size(); i += width) out.emplace_back(b.begin() + (long)i, b.begin() + (long)(i + width));
    return out;
  }

  // -- batch forms of `trait Group` (src/group.rs:24-124) --
  // Group::exp for n (base, scalar) pairs
  std::vector<Bytes> batch_exp(const std::vector<Bytes>& bases, const std::vector<Bytes>& scalars) const {
    Bytes out(scalars.size() * G::EB);
    Bytes b = concat(bases, G::EB), s = concat(scalars, G::SB);
    check(mpvss_batch_exp(ctx_, b.data(), G::EB, s.data(), scalars.size(), out.data()));
    return split(out, G::EB);
  }
  Bytes exp(const Bytes& base, const Bytes& scalar) const { return batch_exp({base}, {scalar})[0]; }  // group.rs:58
  std::vector<Bytes> batch_fixed_base_exp(const std::vector<Bytes>& scalars, int generator = MPVSS_GEN_MAIN) const {
    Bytes out(scalars.size() * G::EB), s = concat(scalars, G::SB);
    check(mpvss_fixed_base_exp(ctx_, generator, s.data(), scalars.size(), out.data()));
    return split(out, G::EB);
  }
  std::vector<Bytes> batch_mul(const std::vector<Bytes>& a, const std::vector<Bytes>& b) const {
    Bytes out(a.size() * G::EB), x = concat(a, G::EB), y = concat(b, G::EB);
    check(mpvss_batch_mul(ctx_, x.data(), y.data(), a.size(), out.data()));
    return split(out, G::EB);
  }
  Bytes mul(const Bytes& a, const Bytes& b) const { return batch_mul({a}, {b})[0]; }  // group.rs:66
  // X_i = prod_j C_j^(i^j) (participant.rs:207-215)
  std::vector<Bytes> batch_poly_eval_in_exponent(const std::vector<Bytes>& commitments, const std::vector<int64_t>& positions) const {
    Bytes out(positions.size() * G::EB), c = concat(commitments, G::EB);
    check(mpvss_poly_eval_exp(ctx_, c.data(), commitments.size(), positions.data(), positions.size(), out.data()));
    return split(out, G::EB);
  }
  // Polynomial::get_value(i) % order (polynomial.rs:50-58, participant.rs:202)
  std::vector<Bytes> scalar_poly_eval(const std::vector<Bytes>& coeffs, const std::vector<int64_t>& positions) const {
    Bytes out(positions.size() * G::SB), c = concat(coeffs, G::SB);
    check(mpvss_scalar_poly_eval(ctx_, c.data(), coeffs.size(), positions.data(), positions.size(), out.data()));
    return split(out, G::SB);
  }
  Bytes multi_exp(const std::vector<Bytes>& bases, const std::vector<Bytes>& scalars) const {
    Bytes out(G::EB), b = concat(bases, G::EB), s = concat(scalars, G::SB);
    check(mpvss_multi_exp(ctx_, b.data(), s.data(), scalars.size(), out.data()));
    return out;
  }

  // Group::generate_private_key (modp.rs:162-174: below q and coprime to q - 1; secp256k1.rs:158-166 and
  // ristretto255.rs:227-235: non-zero below the order)
  Bytes generate_private_key() {
    const Bytes bound = G::id == MPVSS_GROUP_MODP ? ModpGroup::modulus_be() : G::order_be();
    const Bytes half = G::id == MPVSS_GROUP_MODP ? detail::half_be(G::order_be()) : Bytes();
    for (;;) {
      Bytes k = random_below_be(bound);
      if (detail::is_zero(k)) continue;
      if (G::id == MPVSS_GROUP_MODP && ((k.back() & 1) == 0 || k == half)) continue;  // gcd(k, 2g) must be 1
      return scalar_from_be(k);
    }
  }
  Bytes generate_public_key(const Bytes& private_key) const {  // modp.rs:176-178
    return batch_fixed_base_exp({private_key})[0];
  }
  // uniform scalar below the group's `order` (what Polynomial::init draws coefficients from)
  Bytes random_scalar() { return scalar_from_be(random_below_be(G::order_be())); }

  // big-endian integer -> boundary scalar encoding
  static Bytes scalar_from_be(const Bytes& be) {
    Bytes s(G::SB, 0);
    if (be.size() > G::SB) throw Error(MPVSS_ERR_ARG, "mpvss_b200: scalar too long");
    std::copy(be.begin(), be.end(), s.begin() + (long)(G::SB - be.size()));
    return G::scalar_big_endian ? s : detail::reversed(s);
  }

 private:
  explicit Group(int device) {
    int rc = mpvss_ctx_create(G::id, device, &ctx_);
    if (rc != MPVSS_OK) {
      std::string msg = ctx_ ? mpvss_last_error(ctx_) : "mpvss_ctx_create failed (no CUDA device?)";
      if (ctx_) mpvss_ctx_destroy(ctx_);
      ctx_ = nullptr;
      throw Error(rc, "mpvss_b200: " + msg);
    }
    std::random_device rd;
    std::seed_seq seq{rd(), rd(), rd(), rd(), rd(), rd(), rd(), rd()};
    rng_.seed(seq);
  }
  Bytes random_below_be(const Bytes& bound) {  // rejection sampling on the bound's bit length
    int top = 0;
    for (int b = 7; b >= 0; --b) if (bound[0] >> b & 1) { top = b + 1; break; }
    for (;;) {
      Bytes k(bound.size());
      for (auto& x : k) x = (uint8_t)rng_();
      k[0] &= (uint8_t)((1u << top) - 1u);
      if (detail::less_be(k, bound)) return k;
    }
  }
  mpvss_ctx* ctx_ = nullptr;
  std::mt19937_64 rng_;  // stands in for thread_rng; inject coefficients / witnesses where that matters
};

// ---- containers (src/sharebox.rs) -----------------------------------------------------------------------------
template <class G>
struct ShareBox {  // sharebox.rs:22-27
  Bytes publickey, share, challenge, response;
};

template <class G>
struct DistributionSharesBox {  // sharebox.rs:75-86
  std::vector<Bytes> commitments;
  std::map<Bytes, int64_t> positions;  // keyed by element_to_bytes(publickey)
  std::map<Bytes, Bytes> shares;
  std::vector<Bytes> publickeys;
  Bytes challenge;
  std::map<Bytes, Bytes> responses;
  Bytes U;  // element-size big-endian
};

// ---- Participant<G> (src/participant.rs:64-147 and the per-group entry points) ------------------------------
template <class G>
class Participant {
 public:
  explicit Participant(std::shared_ptr<Group<G>> group) : group_(std::move(group)) {}  // Participant::with_arc
  Bytes privatekey, publickey;

  void initialize() {  // participant.rs:139-146
    privatekey = group_->generate_private_key();
    publickey = group_->generate_public_key(privatekey);
  }
  void initialize_with(const Bytes& private_key) {
    privatekey = private_key;
    publickey = group_->generate_public_key(privatekey);
  }

  // participant.rs:160-286 / 1094-1274 / 1573-1717
  DistributionSharesBox<G> distribute_secret(const Bytes& secret, const std::vector<Bytes>& publickeys, uint32_t threshold) {
    std::vector<Bytes> coeffs(threshold), witnesses(publickeys.size());
    for (auto& c : coeffs) c = group_->random_scalar();             // Polynomial::init, polynomial.rs:34-47
    for (auto& w : witnesses) w = group_->generate_private_key();   // participant.rs:223
    return distribute_secret_with(secret, publickeys, threshold, coeffs, witnesses);
  }
  DistributionSharesBox<G> distribute_secret_with(const Bytes& secret, const std::vector<Bytes>& publickeys, uint32_t threshold,
                                                  const std::vector<Bytes>& coeffs, const std::vector<Bytes>& witnesses) {
    const size_t n = publickeys.size(), t = threshold;
    if (t > n) throw Error(MPVSS_ERR_ARG, "distribute_secret: threshold > number of participants");  // participant.rs:166 asserts
    if (coeffs.size() != t || witnesses.size() != n) throw Error(MPVSS_ERR_ARG, "distribute_secret: coefficient / witness count");
    Bytes comm(t * G::EB), shares(n * G::EB), chal(G::SB), resp(n * G::SB), u(G::EB);
    Bytes c = Group<G>::concat(coeffs, G::SB), w = Group<G>::concat(witnesses, G::SB), pk = Group<G>::concat(publickeys, G::EB);
    group_->check(mpvss_distribute(group_->ctx(), n, t, secret.data(), secret.size(), c.data(), w.data(), pk.data(), comm.data(),
                                   shares.data(), chal.data(), resp.data(), u.data(), nullptr));
    DistributionSharesBox<G> box;
    box.commitments = Group<G>::split(comm, G::EB);
    auto ys = Group<G>::split(shares, G::EB);
    auto rs = Group<G>::split(resp, G::SB);
    for (size_t i = 0; i < n; ++i) {  // participant.rs:196-248: position i + 1 in publickeys order
      box.positions[publickeys[i]] = (int64_t)i + 1;
      box.shares[publickeys[i]] = ys[i];
      box.responses[publickeys[i]] = rs[i];
    }
    box.publickeys = publickeys;
    box.challenge = chal;
    box.U = u;
    return box;
  }

  // participant.rs:399-455 / 1384-1442 / 1827-1885; mpvss.rs:90-144
  bool verify_distribution_shares(const DistributionSharesBox<G>& box, Bytes* digest_out = nullptr) const {
    const size_t n = box.publickeys.size(), t = box.commitments.size();
    std::vector<int64_t> pos;
    Bytes ys, rs;
    for (const auto& pk : box.publickeys) {  // a missing map entry makes the reference return false (participant.rs:415-420)
      auto p = box.positions.find(pk);
      auto y = box.shares.find(pk);
      auto r = box.responses.find(pk);
      if (p == box.positions.end() || y == box.shares.end() || r == box.responses.end()) return false;
      if (y->second.size() != G::EB || r->second.size() != G::SB) return false;
      pos.push_back(p->second);
      ys.insert(ys.end(), y->second.begin(), y->second.end());
      rs.insert(rs.end(), r->second.begin(), r->second.end());
    }
    if (n == 0 || t == 0 || box.challenge.size() != G::SB) return false;
    Bytes comm = Group<G>::concat(box.commitments, G::EB), pks = Group<G>::concat(box.publickeys, G::EB), dig(32);
    int ok = 0;
    group_->check(mpvss_verify_distribution(group_->ctx(), n, t, comm.data(), pos.data(), pks.data(), ys.data(), rs.data(),
                                            box.challenge.data(), &ok, nullptr, nullptr, nullptr, dig.data()));
    if (digest_out) *digest_out = dig;
    return ok == 1;
  }

  // batch form of extract_secret_share: one std::optional<ShareBox> per (private_key, w)
  std::vector<std::optional<ShareBox<G>>> extract_secret_shares(const DistributionSharesBox<G>& box, const std::vector<Bytes>& private_keys,
                                                                const std::vector<Bytes>& ws) const {
    const size_t n = private_keys.size();
    std::vector<std::optional<ShareBox<G>>> out(n);
    auto pks = group_->batch_fixed_base_exp(private_keys);  // participant.rs:306
    std::vector<size_t> live;
    Bytes ys, sk, w;
    for (size_t i = 0; i < n; ++i) {
      auto y = box.shares.find(pks[i]);                      // participant.rs:310: unknown key -> None
      if (y == box.shares.end()) continue;
      live.push_back(i);
      ys.insert(ys.end(), y->second.begin(), y->second.end());
      sk.insert(sk.end(), private_keys[i].begin(), private_keys[i].end());
      w.insert(w.end(), ws[i].begin(), ws[i].end());
    }
    const size_t m = live.size();
    if (!m) return out;
    Bytes pko(m * G::EB), so(m * G::EB), co(m * G::SB), ro(m * G::SB);
    std::vector<int> st(m);
    group_->check(mpvss_extract_shares(group_->ctx(), m, sk.data(), w.data(), ys.data(), pko.data(), so.data(), co.data(), ro.data(), st.data()));
    auto p2 = Group<G>::split(pko, G::EB), s2 = Group<G>::split(so, G::EB), c2 = Group<G>::split(co, G::SB), r2 = Group<G>::split(ro, G::SB);
    for (size_t j = 0; j < m; ++j)
      if (st[j] == MPVSS_OK) out[live[j]] = ShareBox<G>{p2[j], s2[j], c2[j], r2[j]};  // participant.rs:314: no inverse -> None
    return out;
  }
  // participant.rs:294-353 / 1282-1338 / 1725-1781
  std::optional<ShareBox<G>> extract_secret_share(const DistributionSharesBox<G>& box, const Bytes& private_key, const Bytes& w) const {
    return extract_secret_shares(box, {private_key}, {w})[0];
  }

  std::vector<bool> verify_shares(const std::vector<ShareBox<G>>& shareboxes, const DistributionSharesBox<G>& box,
                                  const std::vector<Bytes>& publickeys) const {
    const size_t n = shareboxes.size();
    std::vector<bool> res(n, false);
    std::vector<size_t> live;
    Bytes pk, s, ys, c, r;
    for (size_t i = 0; i < n; ++i) {
      auto y = box.shares.find(publickeys[i]);  // participant.rs:371-375
      if (y == box.shares.end()) continue;
      const auto& sb = shareboxes[i];
      if (sb.share.size() != G::EB || sb.challenge.size() != G::SB || sb.response.size() != G::SB) continue;
      live.push_back(i);
      pk.insert(pk.end(), publickeys[i].begin(), publickeys[i].end());
      s.insert(s.end(), sb.share.begin(), sb.share.end());
      ys.insert(ys.end(), y->second.begin(), y->second.end());
      c.insert(c.end(), sb.challenge.begin(), sb.challenge.end());
      r.insert(r.end(), sb.response.begin(), sb.response.end());
    }
    const size_t m = live.size();
    if (!m) return res;
    std::vector<int> ok(m);
    group_->check(mpvss_verify_shares(group_->ctx(), m, pk.data(), s.data(), ys.data(), c.data(), r.data(), ok.data()));
    for (size_t j = 0; j < m; ++j) res[live[j]] = ok[j] == 1;
    return res;
  }
  // participant.rs:361-386 / 1346-1371 / 1789-1814
  bool verify_share(const ShareBox<G>& sharebox, const DistributionSharesBox<G>& box, const Bytes& publickey) const {
    return verify_shares({sharebox}, box, {publickey})[0];
  }

  // participant.rs:462-519 / 1452-1513 / 1895-1950: the secret as a big-endian byte string without leading zeros
  std::optional<Bytes> reconstruct(const std::vector<ShareBox<G>>& share_boxes, const DistributionSharesBox<G>& box) const {
    if (share_boxes.size() < box.commitments.size()) return std::nullopt;  // participant.rs:469
    std::map<int64_t, Bytes> shares;                                        // BTreeMap, participant.rs:476-482
    for (const auto& sb : share_boxes) {
      auto p = box.positions.find(sb.publickey);
      if (p == box.positions.end()) return std::nullopt;
      shares[p->second] = sb.share;
    }
    std::vector<int64_t> pos;
    Bytes s;
    for (const auto& kv : shares) {
      pos.push_back(kv.first);
      s.insert(s.end(), kv.second.begin(), kv.second.end());
    }
    Bytes sec(G::EB);
    group_->check(mpvss_reconstruct(group_->ctx(), pos.size(), pos.data(), s.data(), box.U.data(), sec.data(), nullptr));
    size_t lead = 0;
    while (lead + 1 < sec.size() && sec[lead] == 0) ++lead;
    return Bytes(sec.begin() + (long)lead, sec.end());
  }

  const std::shared_ptr<Group<G>>& group() const { return group_; }

 private:
  std::shared_ptr<Group<G>> group_;
};

// lib.rs:49-57
inline Bytes string_to_secret(const std::string& message) { return Bytes(message.begin(), message.end()); }
inline std::string string_from_secret(const Bytes& secret) { return std::string(secret.begin(), secret.end()); }

}  // namespace mpvss
