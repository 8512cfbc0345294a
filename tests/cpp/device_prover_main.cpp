// Driver for the C++ device-resident prover (baby-plonk-rust_b200/host/plonk_prover.hpp): reads a circuit instance
// written by tests/test_gpu_cpp_host.py (little-endian u64 header + Montgomery limb columns), proves `reps` times and
// prints the 624 proof bytes as hex.  Layout of the input file:
//   u64 n, u64 n_pub, u64 srs_powers, u64 cache, u64 reps, Scalar tau,
//   5 selector columns, 3 sigma columns, 3 wire columns (n Scalars each), n_pub public inputs, 11 blinding scalars
#include <chrono>
#include <cstdio>
#include <fstream>
#include <iostream>

#include "../../baby-plonk-rust_b200/host/plonk_prover.hpp"

using namespace baby_plonk;

static std::vector<Scalar> read_col(std::ifstream& f, size_t n) {
    std::vector<Scalar> v(n);
    f.read((char*)v.data(), (std::streamsize)(n * sizeof(Scalar)));
    return v;
}

int main(int argc, char** argv) {
    if (argc < 2) {
        std::fprintf(stderr, "usage: %s instance.bin\n", argv[0]);
        return 2;
    }
    try {
        std::ifstream f(argv[1], std::ios::binary);
        uint64_t hdr[5];
        f.read((char*)hdr, sizeof hdr);
        const size_t n = hdr[0], n_pub = hdr[1], powers = hdr[2];
        const bool cache = hdr[3] != 0;
        const int reps = (int)hdr[4];
        Scalar tau;
        f.read((char*)tau.l, 32);
        std::vector<std::vector<Scalar>> sel, sig, wires;
        for (int i = 0; i < 5; i++) sel.push_back(read_col(f, n));
        for (int i = 0; i < 3; i++) sig.push_back(read_col(f, n));
        for (int i = 0; i < 3; i++) wires.push_back(read_col(f, n));
        std::vector<Scalar> pub = read_col(f, n_pub), blinding = read_col(f, 11);
        if (!f) throw Panic("short instance file");
        Setup setup = Setup::generate_srs(powers, tau);
        if (n >= 64) setup.precompute();
        DeviceProver prover(setup, n, sel, sig, cache);
        for (int r = 0; r < reps; r++) {
            auto t0 = std::chrono::steady_clock::now();
            Proof proof = prover.prove(wires, pub, blinding);
            double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            std::fprintf(stderr, "prove_ms %.3f\n", ms);
            auto bytes = proof.to_bytes();
            for (uint8_t b : bytes) std::printf("%02x", b);
            std::printf("\n");
        }
        return 0;
    } catch (const std::exception& e) {
        std::printf("PANIC %s\n", e.what());
        return 1;
    }
}
