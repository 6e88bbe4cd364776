// Exercises include/typlonk_b200.hpp, the C++ host side above the C ABI, the way the reference's own tests exercise
// its Rust API (kzg/src/lib.rs:96-170, plonk/src/builder/test.rs:25-44, README.md:16-33).
//   mirror_test host           no device: tracer, permutation builder, Fr helpers, proof codec
//   mirror_test device         needs a GPU: prints `<name> <proof hex>` lines the pytest wrapper compares with the
//                              committed golden proofs, and self-checks verify / kzg / permutation behaviour
#include <cstdio>
#include <cstdlib>
#include <string>

#include "typlonk_b200.hpp"

using namespace typlonk;

static int failures = 0;
#define CHECK(name, cond)                          \
  do {                                             \
    bool ok__ = (cond);                            \
    std::printf("%s %s\n", ok__ ? "ok  " : "FAIL", name); \
    if (!ok__) failures++;                         \
  } while (0)

struct Pythagoras {  // README.md:16-27, plonk/src/builder/test.rs:12-23
  static constexpr size_t INPUTS = 3;
  template <class V>
  static void run(std::array<V, 3> in) {
    auto [a, b, c] = in;
    a = a.clone() * a;
    b = b.clone() * b;
    c = c.clone() * c;
    auto d = a + b;
    d.assert_eq(c);
  }
};
struct Additive {  // plonk/src/builder/test.rs:3-11
  static constexpr size_t INPUTS = 5;
  template <class V>
  static void run(std::array<V, 5> in) {
    auto [a, b, c, d, e] = in;
    auto x = (c + d) + e;
    a = a + b;
    a.assert_eq(x);
  }
};
template <int G>
struct MulChain {
  static constexpr size_t INPUTS = 2;
  template <class V>
  static void run(std::array<V, 2> in) {
    auto [x, y] = in;
    for (int i = 0; i < G; i++) x = x * y.clone();
  }
};

static void print_hex(const char* name, const std::vector<uint8_t>& raw) {
  std::printf("PROOF %s ", name);
  for (uint8_t b : raw) std::printf("%02x", b);
  std::printf("\n");
}

static int host_only() {
  // Fr helpers
  CHECK("Fr(-1) + canonical round trip", Fr::from_canonical(Fr(-1).to_canonical()) == Fr(-1));
  auto c = Fr(5).to_canonical();
  CHECK("Fr(5) canonical", c[0] == 5 && c[1] == 0 && c[31] == 0);
  std::array<uint8_t, 32> big;
  big.fill(0xff);
  bool threw = false;
  try {
    Fr::from_canonical(big);
  } catch (const Malformed&) {
    threw = true;
  }
  CHECK("non-canonical scalar rejected", threw);
  // permutation builder: the README circuit's constraints
  auto pb = PermutationBuilder::with_rows(4);
  pb.add_constrains({{{0, 0}, {1, 0}}, {{0, 1}, {1, 1}}, {{0, 2}, {1, 2}}, {{2, 0}, {0, 3}}, {{2, 1}, {1, 3}}, {{2, 3}, {2, 2}}});
  auto perm = pb.build(8);
  const uint64_t want[8][3] = {{8, 0, 3}, {9, 1, 11}, {10, 2, 19}, {16, 17, 18}, {4, 12, 20}, {5, 13, 21}, {6, 14, 22}, {7, 15, 23}};
  bool same = perm.perm.size() == 24;
  for (int j = 0; j < 8 && same; j++)
    for (int i = 0; i < 3; i++) same = same && perm.perm[j + i * 8] == want[j][i];
  CHECK("README permutation golden", same);
  CHECK("invalid tag is Err", !pb.add_constrain({0, 0}, {1, 4}));
  threw = false;
  try {
    pb.add_constrains({{{0, 0}, {9, 9}}});
  } catch (const Error& e) {
    threw = e.code == TP_ERR_INVALID_TAG;
  }
  CHECK("add_constrains unwraps", threw);
  // Fr::rand stream is deterministic and below r
  auto s1 = Fr::rand_stream(2, 9), s2 = Fr::rand_stream(2, 9);
  CHECK("rand_stream deterministic", s1 == s2 && s1[0] != s1[1]);
  // no device -> no context (there is no CPU fallback): only checked when asked to
  if (std::getenv("TYPLONK_EXPECT_NO_DEVICE")) {
    threw = false;
    try {
      Context ctx;
    } catch (const Error& e) {
      threw = e.code == TP_ERR_NO_DEVICE || e.code == TP_ERR_CUDA;
    }
    CHECK("no device -> Context throws", threw);
  }
  return failures;
}

template <class D, size_t I>
static void prove_and_print(const Context& ctx, const char* name, const Fr& tau, const std::array<Fr, I>& inputs,
                            const std::array<Fr, 9>& blinders, bool expect_verify) {
  auto circuit = build<D>(ctx, tau);
  auto proof = circuit.prove(inputs, {Fr(0)}, blinders);
  print_hex(name, proof.to_bytes());
  CHECK((std::string(name) + ": verify").c_str(), circuit.verify(proof) == expect_verify);
  auto back = Proof::from_bytes(proof.to_bytes());
  CHECK((std::string(name) + ": codec round trip").c_str(), back.fixed == proof.fixed && back.public_inputs == proof.public_inputs);
  Proof bad = proof;
  bad.fixed[192] ^= 1;  // a(zeta)
  CHECK((std::string(name) + ": tampered proof rejected").c_str(), !circuit.verify(bad));
}

static int device(const Context& ctx) {
  const Fr tau = Fr::rand_stream(1, 1)[0];
  std::array<Fr, 9> blinders;
  {
    auto b = Fr::rand_stream(2, 9);
    for (int i = 0; i < 9; i++) blinders[i] = b[i];
  }
  // kzg/src/lib.rs:96-108 `commit`: srs from secret 2, p = 1 + 2X + 3X^2, commit == p(2) G = 17 G; p(1) == 6
  {
    auto srs = Srs::from_secret(ctx, Fr(2), 10);
    KzgScheme scheme(srs);
    Poly p = {Fr(1), Fr(2), Fr(3)};
    auto com = scheme.commit(p);
    CHECK("commit(1 + 2X + 3X^2) == 17 G", com == scheme.commit({Fr(17)}));
    auto open = scheme.open(p, Fr(1));
    CHECK("p(1) == 6", open.y == Fr(6));
    CHECK("kzg verify accepts", scheme.verify(com, open, Fr(1)));
    CHECK("kzg verify rejects a wrong point", !scheme.verify(com, open, Fr(2)));
    CHECK("commit(0) is the point at infinity", scheme.commit({}).point.is_zero() && scheme.commit({Fr(0), Fr(0)}).point.is_zero());
    CHECK("identity == srs[0]", scheme.identity().point == srs.g1_ref(0, 1)[0]);
    bool threw = false;
    try {
      scheme.open({}, Fr(1));
    } catch (const Error& e) {
      threw = e.code == TP_ERR_EMPTY_POLY;
    }
    CHECK("open(empty) panics", threw);
    threw = false;
    try {
      scheme.commit(Poly(14, Fr(1)));
    } catch (const Error& e) {
      threw = e.code == TP_ERR_SRS_TOO_SHORT;
    }
    CHECK("commit longer than the SRS panics", threw);
    auto again = Srs::from_bytes(ctx, srs.to_bytes());
    CHECK("SRS wire round trip", again.g1_ref() == srs.g1_ref() && again.g2s_ref() == srs.g2s_ref());
  }
  // permutation crate: compile + prove; the grand product of a satisfied permutation returns to 1
  {
    auto pb = PermutationBuilder::with_rows(4);
    pb.add_constrains({{{0, 0}, {1, 1}}, {{2, 3}, {0, 2}}});
    auto compiled = pb.build(4).compile(ctx);
    std::array<std::vector<Fr>, 3> values;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 4; j++) values[i].push_back(Fr(10 * i + j + 1));
    values[1][1] = values[0][0];
    values[0][2] = values[2][3];
    auto z = compiled.prove(values, Fr(7), Fr(11));
    CHECK("grand product: n + 1 values, starts and ends at 1", z.size() == 5 && z[0] == Fr(1) && z[4] == Fr(1));
    values[1][1] = Fr(999);
    auto z2 = compiled.prove(values, Fr(7), Fr(11));
    CHECK("grand product: broken copy does not return to 1", z2[4] != Fr(1));
  }
  prove_and_print<Pythagoras, 3>(ctx, "readme_pythagoras_3_4_5", tau, {Fr(3), Fr(4), Fr(5)}, blinders, true);
  prove_and_print<Pythagoras, 3>(ctx, "readme_pythagoras_bad_3_4_6", tau, {Fr(3), Fr(4), Fr(6)}, blinders, false);  // test.rs:31-37
  prove_and_print<Additive, 5>(ctx, "additive_2_7_2_3_4", tau, {Fr(2), Fr(7), Fr(2), Fr(3), Fr(4)}, blinders, true);
  prove_and_print<MulChain<13>, 2>(ctx, "mulchain_13_gates", tau, {Fr(3), Fr(5)}, blinders, true);
  prove_and_print<MulChain<61>, 2>(ctx, "mulchain_61_gates", tau, {Fr(3), Fr(5)}, blinders, true);
  {
    auto circuit = build<Pythagoras>(ctx, tau);
    bool threw = false;
    try {
      circuit.prove({Fr(3), Fr(4), Fr(5)}, {Fr(1)}, blinders);
    } catch (const GateUnsatisfied&) {
      threw = true;
    }
    CHECK("non-zero public input panics like the reference", threw);
    auto random_setup = build<Pythagoras>(ctx);                  // Circuit::build() with a fresh tau, thread_rng blinders
    CHECK("README flow with random tau / blinders verifies", random_setup.verify(random_setup.prove({Fr(3), Fr(4), Fr(5)}, {Fr(0)})));
  }
  std::printf("launches %llu\n", (unsigned long long)ctx.launch_count());
  return failures;
}

int main(int argc, char** argv) {
  std::string mode = argc > 1 ? argv[1] : "host";
  int f;
  if (mode == "device") {
    f = device(Context());
  } else if (mode == "multi") {   // `multi 0,0,0`: the same program on a device group (three ranks on device 0)
    std::vector<int> devices;
    std::string list = argc > 2 ? argv[2] : "0,0";
    for (size_t pos = 0; pos <= list.size();) {
      size_t comma = list.find(',', pos);
      if (comma == std::string::npos) comma = list.size();
      devices.push_back(std::stoi(list.substr(pos, comma - pos)));
      pos = comma + 1;
    }
    Context ctx = Context::multi(devices);
    std::printf("ranks %d\n", ctx.ranks());
    f = device(ctx);
  } else {
    f = host_only();
  }
  std::printf("%d failures\n", f);
  return f ? 1 : 0;
}
