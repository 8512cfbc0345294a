// CPU-only check of the C++ transcript and encodings (host/plonk_prover.hpp); no GPU call is made.
// Prints: merlin's conformance challenge, a challenge scalar of the PLONK schedule, the compressed generator.
#include <cstdio>

#include "../../baby-plonk-rust_b200/host/plonk_prover.hpp"

using namespace baby_plonk;

int main() {
    MerlinTranscript t("test protocol");
    t.append_message("some label", (const uint8_t*)"some data", 9);
    uint8_t c[32];
    t.challenge_bytes("challenge", c, 32);
    for (uint8_t b : c) std::printf("%02x", b);
    std::printf("\n");
    PlonkTranscript p;
    auto gen = g1_to_compressed(G1Projective::generator());
    p.append_point("a_1", gen);
    p.append_scalar("a_eval", Scalar::from(12345));
    auto ch = scalar_to_bytes(p.get_and_append_challenge("beta"));
    for (uint8_t b : ch) std::printf("%02x", b);
    std::printf("\n");
    for (uint8_t b : gen) std::printf("%02x", b);
    std::printf("\n");
    auto neg = g1_to_compressed(G1Projective::generator() * Scalar::from(5).neg());
    for (uint8_t b : neg) std::printf("%02x", b);
    std::printf("\n");
    return 0;
}
