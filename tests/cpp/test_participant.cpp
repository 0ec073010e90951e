// The reference's own protocol tests, restated over include/mpvss_b200.hpp (C++ mirror of Participant<G> on the
// C ABI; every group operation runs in the CUDA library).  Same names and flow as
//   tests/mpvss_tests.rs:11            test_mpvss_distribute_verify_reconstruct
//   src/participant.rs:593, 703        test_end_to_end_modp, test_threshold_subset_modp_positions_1_and_3
//   src/participant.rs:752, 832        test_end_to_end_secp256k1, test_threshold_secp256k1
//   examples/mpvss_all_ristretto255.rs, examples/mpvss_sub_ristretto255.rs (as tests)
// plus the negative cases the reference's maps imply (unknown key -> None / false, too few shares -> None,
// tampered box -> false).  Needs a GPU; run by tests/test_cpp_mirror.py (-m gpu).  Prints one line per test.
#include <cstdio>
#include <functional>

#include "mpvss_b200.hpp"

using namespace mpvss;

static int failures = 0;
#define CHECK(cond, msg)                                                   \
  do {                                                                     \
    if (!(cond)) {                                                         \
      std::printf("    FAILED %s:%d: %s\n", __FILE__, __LINE__, msg);      \
      ++failures;                                                          \
      return;                                                              \
    }                                                                      \
  } while (0)

template <class G>
static std::vector<Participant<G>> make_participants(const std::shared_ptr<Group<G>>& group, size_t n) {
  std::vector<Participant<G>> ps;
  for (size_t i = 0; i < n; ++i) {
    ps.emplace_back(group);
    ps.back().initialize();
  }
  return ps;
}

// tests/mpvss_tests.rs:11-91
static void test_mpvss_distribute_verify_reconstruct() {
  auto group = Group<ModpGroup>::create();
  const std::string secret_message = "Hello MPVSS.";
  Participant<ModpGroup> dealer(group);
  dealer.initialize();
  auto p = make_participants(group, 3);
  auto box = dealer.distribute_secret(string_to_secret(secret_message), {p[0].publickey, p[1].publickey, p[2].publickey}, 3);
  CHECK(p[0].verify_distribution_shares(box), "p1 verifies the distribution");
  CHECK(p[1].verify_distribution_shares(box), "p2 verifies the distribution");
  CHECK(p[2].verify_distribution_shares(box), "p3 verifies the distribution");
  Bytes w = group->generate_private_key();
  auto s1 = p[0].extract_secret_share(box, p[0].privatekey, w);
  auto s2 = p[1].extract_secret_share(box, p[1].privatekey, w);
  auto s3 = p[2].extract_secret_share(box, p[2].privatekey, w);
  CHECK(s1 && s2 && s3, "every participant extracts its share");
  CHECK(p[0].verify_share(*s2, box, p[1].publickey), "p1 verifies s2");
  CHECK(p[1].verify_share(*s3, box, p[2].publickey), "p2 verifies s3");
  CHECK(p[2].verify_share(*s1, box, s1->publickey), "p3 verifies s1");
  std::vector<ShareBox<ModpGroup>> shares{*s1, *s2, *s3};
  for (auto& q : p) {
    auto r = q.reconstruct(shares, box);
    CHECK(r && string_from_secret(*r) == secret_message, "reconstructed message equals the original");
  }
}

// src/participant.rs:593-697 (modp), :752-826 (secp256k1); examples/mpvss_all_ristretto255.rs
template <class G>
static void test_end_to_end(const char* message) {
  auto group = Group<G>::create();
  Participant<G> dealer(group);
  dealer.initialize();
  auto p = make_participants(group, 3);
  std::vector<Bytes> publickeys{p[0].publickey, p[1].publickey, p[2].publickey};
  auto box = dealer.distribute_secret(string_to_secret(message), publickeys, 3);
  CHECK(dealer.verify_distribution_shares(box), "distribution is valid");
  CHECK(box.publickeys.size() == 3 && box.commitments.size() == 3 && box.shares.size() == 3, "box structure");
  CHECK(!detail::is_zero(box.U), "U is not zero");
  Bytes w = group->generate_private_key();
  std::vector<ShareBox<G>> s;
  for (auto& q : p) {
    auto sb = q.extract_secret_share(box, q.privatekey, w);
    CHECK(sb.has_value(), "share extracted");
    CHECK(sb->publickey == q.publickey, "sharebox carries the participant's key");
    CHECK(!detail::is_zero(sb->share), "share is not zero");
    s.push_back(*sb);
  }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      if (i != j) CHECK(p[i].verify_share(s[j], box, p[j].publickey), "every participant verifies the others' shares");
  auto r = dealer.reconstruct(s, box);
  CHECK(r && string_from_secret(*r) == message, "reconstructed message equals the original");
}

// src/participant.rs:703-746
static void test_threshold_subset_modp_positions_1_and_3() {
  auto group = Group<ModpGroup>::create();
  Participant<ModpGroup> dealer(group);
  dealer.initialize();
  auto p = make_participants(group, 3);
  const Bytes secret{0x01, 0xe2, 0x40};  // 123456
  auto box = dealer.distribute_secret(secret, {p[0].publickey, p[1].publickey, p[2].publickey}, 2);
  Bytes w = group->generate_private_key();
  auto s1 = p[0].extract_secret_share(box, p[0].privatekey, w);
  auto s3 = p[2].extract_secret_share(box, p[2].privatekey, w);
  CHECK(s1 && s3, "shares extracted");
  auto r = dealer.reconstruct({*s1, *s3}, box);
  CHECK(r && *r == secret, "threshold-2 reconstruction from positions 1 and 3 recovers the secret");
}

// src/participant.rs:832-903 (secp256k1); examples/mpvss_sub*.rs: 5 participants, threshold 3, shares 1, 3, 5
template <class G>
static void test_threshold(const char* message) {
  auto group = Group<G>::create();
  Participant<G> dealer(group);
  dealer.initialize();
  auto p = make_participants(group, 5);
  std::vector<Bytes> publickeys;
  for (auto& q : p) publickeys.push_back(q.publickey);
  auto box = dealer.distribute_secret(string_to_secret(message), publickeys, 3);
  CHECK(dealer.verify_distribution_shares(box), "distribution is valid");
  Bytes w = group->generate_private_key();
  auto s1 = p[0].extract_secret_share(box, p[0].privatekey, w);
  auto s3 = p[2].extract_secret_share(box, p[2].privatekey, w);
  auto s5 = p[4].extract_secret_share(box, p[4].privatekey, w);
  CHECK(s1 && s3 && s5, "three shares extracted");
  auto r = dealer.reconstruct({*s1, *s3, *s5}, box);
  CHECK(r && string_from_secret(*r) == message, "reconstructed message equals the original");
  // too few shares: participant.rs:469 returns None
  CHECK(!dealer.reconstruct({*s1, *s3}, box).has_value(), "two shares of a threshold-3 box give None");
}

// what the reference's maps and checks imply for bad inputs
template <class G>
static void test_rejections() {
  auto group = Group<G>::create();
  Participant<G> dealer(group);
  dealer.initialize();
  auto p = make_participants(group, 4);
  std::vector<Bytes> publickeys{p[0].publickey, p[1].publickey, p[2].publickey};
  auto box = dealer.distribute_secret(string_to_secret("reject me"), publickeys, 2);
  CHECK(dealer.verify_distribution_shares(box), "untampered box verifies");
  Bytes w = group->generate_private_key();
  // a key that is not in the box: extract -> None (participant.rs:310), verify_share -> false (:371-375)
  CHECK(!p[3].extract_secret_share(box, p[3].privatekey, w).has_value(), "outsider cannot extract");
  auto s1 = p[0].extract_secret_share(box, p[0].privatekey, w);
  CHECK(s1.has_value(), "insider extracts");
  CHECK(!dealer.verify_share(*s1, box, p[3].publickey), "share does not verify under an unknown key");
  CHECK(!dealer.verify_share(*s1, box, p[1].publickey), "share does not verify under somebody else's key");
  // tampering: swap two encrypted shares; flip a response bit; drop a map entry
  auto swapped = box;
  std::swap(swapped.shares[publickeys[0]], swapped.shares[publickeys[1]]);
  CHECK(!dealer.verify_distribution_shares(swapped), "swapped shares are rejected");
  auto flipped = box;
  flipped.responses[publickeys[2]][G::scalar_big_endian ? G::SB - 1 : 0] ^= 1;
  CHECK(!dealer.verify_distribution_shares(flipped), "a changed response is rejected");
  auto missing = box;
  missing.responses.erase(publickeys[1]);
  CHECK(!dealer.verify_distribution_shares(missing), "a missing response makes the box invalid (participant.rs:415-420)");
  auto bad_share = *s1;
  bad_share.response[G::scalar_big_endian ? G::SB - 1 : 0] ^= 1;
  CHECK(!dealer.verify_share(bad_share, box, p[0].publickey), "a changed share proof is rejected");
}

// injected randomness: the same coefficients and witnesses give the same box (what the parity tests build on)
template <class G>
static void test_injected_randomness_is_deterministic() {
  auto group = Group<G>::create();
  Participant<G> dealer(group);
  dealer.initialize();
  auto p = make_participants(group, 6);
  std::vector<Bytes> publickeys, coeffs, witnesses;
  for (auto& q : p) publickeys.push_back(q.publickey);
  for (int j = 0; j < 4; ++j) coeffs.push_back(group->random_scalar());
  for (int i = 0; i < 6; ++i) witnesses.push_back(group->generate_private_key());
  auto a = dealer.distribute_secret_with(string_to_secret("same"), publickeys, 4, coeffs, witnesses);
  auto b = dealer.distribute_secret_with(string_to_secret("same"), publickeys, 4, coeffs, witnesses);
  CHECK(a.commitments == b.commitments && a.shares == b.shares && a.challenge == b.challenge && a.responses == b.responses && a.U == b.U,
        "distribute_secret_with is a function of its inputs");
  // commitments are g^a_j and X_i = prod C_j^(i^j) equals g^P(i): the dealer's shortcut against the verifier's loop
  auto comm = group->batch_fixed_base_exp(coeffs, MPVSS_GEN_SUBGROUP);
  CHECK(comm == a.commitments, "commitments are subgroup_generator^coefficient (participant.rs:189-193)");
  std::vector<int64_t> pos{1, 2, 3, 4, 5, 6};
  auto x = group->batch_poly_eval_in_exponent(a.commitments, pos);
  auto gp = group->batch_fixed_base_exp(group->scalar_poly_eval(coeffs, pos), MPVSS_GEN_SUBGROUP);
  CHECK(x == gp, "prod_j C_j^(i^j) == g^(P(i) mod order) for every position");
  Bytes digest_a, digest_b;
  CHECK(dealer.verify_distribution_shares(a, &digest_a) && dealer.verify_distribution_shares(b, &digest_b) && digest_a == digest_b,
        "both boxes verify with the same transcript digest");
}

int main() {
  struct T { const char* name; std::function<void()> fn; };
  const T tests[] = {
      {"test_mpvss_distribute_verify_reconstruct", test_mpvss_distribute_verify_reconstruct},
      {"test_end_to_end_modp", [] { test_end_to_end<ModpGroup>("Hello MPVSS End-to-End Test!"); }},
      {"test_threshold_subset_modp_positions_1_and_3", test_threshold_subset_modp_positions_1_and_3},
      {"test_threshold_modp", [] { test_threshold<ModpGroup>("Threshold test modp!"); }},
      {"test_end_to_end_secp256k1", [] { test_end_to_end<Secp256k1Group>("Hello secp256k1 PVSS!"); }},
      {"test_threshold_secp256k1", [] { test_threshold<Secp256k1Group>("Threshold test secp256k1!"); }},
      {"test_end_to_end_ristretto255", [] { test_end_to_end<Ristretto255Group>("Hello Ristretto255 PVSS!"); }},
      {"test_threshold_ristretto255", [] { test_threshold<Ristretto255Group>("Threshold test ristretto255"); }},
      {"test_rejections_modp", test_rejections<ModpGroup>},
      {"test_rejections_secp256k1", test_rejections<Secp256k1Group>},
      {"test_rejections_ristretto255", test_rejections<Ristretto255Group>},
      {"test_injected_randomness_modp", test_injected_randomness_is_deterministic<ModpGroup>},
      {"test_injected_randomness_secp256k1", test_injected_randomness_is_deterministic<Secp256k1Group>},
      {"test_injected_randomness_ristretto255", test_injected_randomness_is_deterministic<Ristretto255Group>},
  };
  for (const auto& t : tests) {
    const int before = failures;
    try {
      t.fn();
    } catch (const std::exception& e) {
      std::printf("    EXCEPTION %s\n", e.what());
      ++failures;
    }
    std::printf("%s %s\n", failures == before ? "ok  " : "FAIL", t.name);
  }
  std::printf("%d failure(s)\n", failures);
  return failures ? 1 : 0;
}
