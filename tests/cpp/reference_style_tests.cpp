// The reference's own unit tests for the hot path, restated against the C++ host mirror
// (baby-plonk-rust_b200/host/baby_plonk.hpp -> libbpk.so).  Each test names the Rust test it follows.
// Built by __graft_entry__.build(); run on the GPU box by tests/test_gpu_cpp_host.py.
#include <cstdio>
#include <functional>
#include <string>
#include <vector>

#include "../../baby-plonk-rust_b200/host/baby_plonk.hpp"

using namespace baby_plonk;

static int failures = 0;
#define CHECK(cond)                                                         \
    do {                                                                    \
        if (!(cond)) {                                                      \
            std::printf("  CHECK FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            failures++;                                                     \
        }                                                                   \
    } while (0)

static bool panics(const std::function<void()>& f) {
    try {
        f();
    } catch (const Panic&) {
        return true;
    }
    return false;
}

// src/setup.rs:45-57  test_generate_srs
static void test_generate_srs() {
    Scalar tau = Scalar::from(2);
    Setup setup = Setup::generate_srs(8, tau);
    for (uint64_t i = 0; i < 8; i++) CHECK(setup.powers_of_x[i] == G1Projective::generator() * tau.pow(i));
}

// src/setup.rs:59-116  test_monomial_commit (the pairing identity is checked in G1: tau is known)
static void test_monomial_commit() {
    {
        Setup setup = Setup::generate_srs(2, Scalar::from(10));
        Polynomial poly({Scalar::from(2), Scalar::from(3)}, Basis::Monomial);
        G1Projective commitment = setup.commit(poly);
        G1Projective g1 = G1Projective::generator();
        CHECK(commitment == g1 * Scalar::from(2) + g1 * (Scalar::from(10) * Scalar::from(3)));
    }
    Scalar tau = Scalar::from(2);
    Setup setup = Setup::generate_srs(8, tau);
    Polynomial poly1({Scalar::zero(), Scalar::one()}, Basis::Monomial);
    CHECK(setup.commit(poly1) == G1Projective::generator() * tau);
    Polynomial p1({Scalar::one(), Scalar::from(2), Scalar::from(3)}, Basis::Monomial);
    Polynomial p3({Scalar::one().neg(), Scalar::one().neg(), Scalar::one().neg(), Scalar::from(3)}, Basis::Monomial);
    // e(commit(p1), [tau - 1]G2) == e(commit(p3), G2)   <=>   [tau - 1] commit(p1) == commit(p3)
    CHECK(setup.commit(p1) * (tau - Scalar::one()) == setup.commit(p3));
    // and p3 really is p1 * (x - 1) through impl Mul
    Polynomial p2({Scalar::one().neg(), Scalar::one()}, Basis::Monomial);
    CHECK(p1 * p2 == p3);
    CHECK(panics([&] { setup.commit(Polynomial({Scalar::one()}, Basis::Lagrange)); }));
}

// src/setup.rs:118-136 test_ntt + src/prover.rs:843-845: ntt / i_ntt are inverse to each other
static void test_ntt_round_trip() {
    std::vector<Scalar> v;
    for (uint64_t i = 0; i < 8; i++) v.push_back(Scalar::from(i * i + 3));
    std::vector<Scalar> e = ntt_381(v);
    CHECK(i_ntt_381(e) == v);
    // definition (utils.rs:63-81): e[1] = sum_j v[j] w^j
    Scalar w = root_of_unity(8), acc = Scalar::zero(), p = Scalar::one();
    for (auto& c : v) {
        acc = acc + c * p;
        p = p * w;
    }
    CHECK(e[1] == acc);
    Polynomial mono(v, Basis::Monomial);
    CHECK(mono.ntt().i_ntt() == mono);
    CHECK(panics([&] { mono.i_ntt(); }));
    CHECK(panics([&] { ntt_381(std::vector<Scalar>(6, Scalar::one())); }));
}

// src/polynomial.rs:437-451  (1 + x)^2 = 1 + 2x + x^2
static void test_polynomial_mul() {
    Polynomial a({Scalar::one(), Scalar::one()}, Basis::Monomial);
    Polynomial c = a * a;
    CHECK(c == Polynomial({Scalar::one(), Scalar::from(2), Scalar::one()}, Basis::Monomial));
    Polynomial l({Scalar::one(), Scalar::one()}, Basis::Lagrange);
    CHECK(panics([&] { l* l; }));
}

// src/utils.rs:238-242  omega_4^4 == 1
static void test_root_of_unity() {
    Scalar w = root_of_unity(4);
    CHECK(w.pow(4) == Scalar::one());
    CHECK(w.pow(2) != Scalar::one());
    CHECK(roots_of_unity(4).size() == 4);
    CHECK(find_next_power_of_two(1, 1) == 4);
}

// src/msm.rs: bucket_msm with the reference signature, zip truncation, and the panicking windows
static void test_bucket_msm() {
    G1Projective g = G1Projective::generator();
    std::vector<G1Projective> pts{g, g * Scalar::from(5), g * Scalar::from(7)};
    std::vector<Scalar> sc{Scalar::from(3), Scalar::from(4), Scalar::one().neg(), Scalar::from(99)};
    G1Projective expect = g * Scalar::from(3) + g * Scalar::from(20) + g * (Scalar::from(7) * Scalar::one().neg());
    CHECK(BucketMSM::bucket_msm(pts, sc, 256, 4) == expect);
    CHECK(BucketMSM::bucket_msm(pts, sc, 256, 8) == expect);
    CHECK(BucketMSM::bucket_msm({}, {}, 256, 4) == G1Projective::identity());
    CHECK(panics([&] { BucketMSM::bucket_msm(pts, sc, 3, 4); }));
    CHECK(panics([&] { BucketMSM::bucket_msm(pts, sc, 300, 4); }));
}

int main() {
    struct T { const char* name; void (*fn)(); } tests[] = {
        {"test_generate_srs", test_generate_srs},   {"test_monomial_commit", test_monomial_commit},
        {"test_ntt_round_trip", test_ntt_round_trip}, {"test_polynomial_mul", test_polynomial_mul},
        {"test_root_of_unity", test_root_of_unity}, {"test_bucket_msm", test_bucket_msm},
    };
    try {
        for (auto& t : tests) {
            int before = failures;
            t.fn();
            std::printf("%s ... %s\n", t.name, failures == before ? "ok" : "FAILED");
        }
    } catch (const std::exception& e) {
        std::printf("unexpected exception: %s\n", e.what());
        return 2;
    }
    std::printf(failures ? "FAILED (%d checks)\n" : "ALL PASSED\n", failures);
    return failures ? 1 : 0;
}
