// Test-only shim exposing the host-side helpers (bigint.h, sha2.h) of the product to ctypes.
#include "../../mpvss_rs_b200/csrc/bigint.h"
#include "../../mpvss_rs_b200/csrc/sha2.h"
#include "../../mpvss_rs_b200/csrc/modp_chain.h"

extern "C" {
// all numbers: little-endian byte strings of the given length; outputs `outlen` bytes
int hc_divmod(const uint8_t* a, size_t alen, const uint8_t* b, size_t blen, uint8_t* q, uint8_t* r, size_t outlen) {
  big::Int qq, rr;
  big::divmod(big::from_le(a, alen), big::from_le(b, blen), &qq, &rr);
  big::to_le(qq, q, outlen);
  big::to_le(rr, r, outlen);
  return 0;
}
int hc_mulmod(const uint8_t* a, size_t alen, const uint8_t* b, size_t blen, const uint8_t* m, size_t mlen, uint8_t* out,
              size_t outlen) {
  big::to_le(big::mulmod(big::from_le(a, alen), big::from_le(b, blen), big::from_le(m, mlen)), out, outlen);
  return 0;
}
int hc_modinv(const uint8_t* a, size_t alen, const uint8_t* m, size_t mlen, uint8_t* out, size_t outlen) {
  big::Int r;
  if (!big::modinv(big::from_le(a, alen), big::from_le(m, mlen), &r)) return 1;
  big::to_le(r, out, outlen);
  return 0;
}
int hc_modinv_odd(const uint8_t* a, size_t alen, const uint8_t* m, size_t mlen, uint8_t* out, size_t outlen) {
  big::Int r;
  if (!big::modinv_odd(big::from_le(a, alen), big::from_le(m, mlen), &r)) return 1;
  big::to_le(r, out, outlen);
  return 0;
}
void hc_sha256(const uint8_t* d, size_t n, size_t chunk, uint8_t* out) {
  sha2::Sha256 h;
  for (size_t i = 0; i < n; i += chunk) h.update(d + i, i + chunk <= n ? chunk : n - i);
  h.finalize(out);
}
void hc_sha512(const uint8_t* d, size_t n, size_t chunk, uint8_t* out) {
  sha2::Sha512 h;
  for (size_t i = 0; i < n; i += chunk) h.update(d + i, i + chunk <= n ? chunk : n - i);
  h.finalize(out);
}
// ops of one Horner step for position p (modp_chain.h); returns the op count or -1
int hc_chain_ops(uint32_t p, uint32_t tree_limit, uint16_t* ops_out, uint32_t* sqr, uint32_t* mul) {
  static modp_chain::PowerTree tree;
  tree.build(tree_limit);
  std::vector<uint16_t> ops;
  if (!modp_chain::step_ops(p, tree, ops, sqr, mul)) return -1;
  for (size_t i = 0; i < ops.size(); ++i) ops_out[i] = ops[i];
  return (int)ops.size();
}
}
