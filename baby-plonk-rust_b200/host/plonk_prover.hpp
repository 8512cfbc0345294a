// plonk_prover.hpp -- C++ host side of the device-resident PLONK prover, layered on the
// C ABI of libbpk.so only (include/bpk.h; no CUDA runtime binding is needed: device memory goes through bpk_dev_*).
// It mirrors, in the compiled-language host layer, what the reference does in
//     Prover::prove, round_1 .. round_5      src/prover.rs:106-647
//     Transcript (merlin, label schedule)     src/transcript.rs:8-86, merlin 3.0.0 (Cargo dependency: STROBE-128)
//     G1Affine::to_compressed                 lib/bls12_381/src/g1.rs:221-242
//     Scalar::to_bytes / from_bytes           lib/bls12_381/src/scalar.rs:238-304
// and is the same algorithm as baby-plonk-rust_b200/prover.py (see its header for the round-by-round mapping onto
// the bpk_* entry points).  Polynomials stay in HBM; per round only commitments, evaluations and challenges cross
// PCIe.  Rust panics (assert!) become baby_plonk::Panic.
#pragma once
#include <array>

#include "baby_plonk.hpp"

namespace baby_plonk {

// ---- merlin transcript: STROBE-128 over keccak-f[1600] (published construction) ---------------------------
class Strobe128 {
    static constexpr int RATE = 166;
    uint8_t st[200];
    int pos = 0, begin = 0, flags = 0;

    static uint64_t rol(uint64_t v, int k) { return k ? (v << k) | (v >> (64 - k)) : v; }
    void permute() {
        static uint64_t rc[24];
        static int rho[25], pi[25];
        static bool init = false;
        if (!init) {
            int lfsr = 1;
            for (int r = 0; r < 24; r++) {
                uint64_t c = 0;
                for (int j = 0; j < 7; j++) {
                    if (lfsr & 1) c |= (uint64_t)1 << ((1 << j) - 1);
                    lfsr <<= 1;
                    if (lfsr & 0x100) lfsr ^= 0x171;
                }
                rc[r] = c;
            }
            for (int i = 0; i < 25; i++) rho[i] = 0;
            int x = 1, y = 0;
            for (int t = 0; t < 24; t++) {
                rho[x + 5 * y] = ((t + 1) * (t + 2) / 2) % 64;
                int nx = y, ny = (2 * x + 3 * y) % 5;
                x = nx;
                y = ny;
            }
            for (int a = 0; a < 5; a++)
                for (int b = 0; b < 5; b++) pi[a + 5 * b] = b + 5 * ((2 * a + 3 * b) % 5);
            init = true;
        }
        uint64_t s[25];
        for (int i = 0; i < 25; i++) {
            s[i] = 0;
            for (int k = 0; k < 8; k++) s[i] |= (uint64_t)st[8 * i + k] << (8 * k);
        }
        for (int r = 0; r < 24; r++) {
            uint64_t col[5], moved[25];
            for (int x = 0; x < 5; x++) col[x] = s[x] ^ s[x + 5] ^ s[x + 10] ^ s[x + 15] ^ s[x + 20];
            for (int x = 0; x < 5; x++) {
                uint64_t d = col[(x + 4) % 5] ^ rol(col[(x + 1) % 5], 1);
                for (int y = 0; y < 25; y += 5) s[x + y] ^= d;
            }
            for (int i = 0; i < 25; i++) moved[pi[i]] = rol(s[i], rho[i]);
            for (int y = 0; y < 25; y += 5)
                for (int x = 0; x < 5; x++) s[x + y] = moved[x + y] ^ (~moved[(x + 1) % 5 + y] & moved[(x + 2) % 5 + y]);
            s[0] ^= rc[r];
        }
        for (int i = 0; i < 25; i++)
            for (int k = 0; k < 8; k++) st[8 * i + k] = (uint8_t)(s[i] >> (8 * k));
    }
    void end_block() {
        st[pos] ^= (uint8_t)begin;
        st[pos + 1] ^= 0x04;
        st[RATE + 1] ^= 0x80;
        permute();
        pos = 0;
        begin = 0;
    }
    void absorb(const uint8_t* d, size_t n) {
        for (size_t k = 0; k < n; k++) {
            st[pos++] ^= d[k];
            if (pos == RATE) end_block();
        }
    }
    void squeeze(uint8_t* out, size_t n) {
        for (size_t k = 0; k < n; k++) {
            out[k] = st[pos];
            st[pos++] = 0;
            if (pos == RATE) end_block();
        }
    }
    void begin_op(int f, bool more) {
        if (more) {
            if (f != flags) throw Panic("continued STROBE operation with different flags");
            return;
        }
        uint8_t hdr[2] = {(uint8_t)begin, (uint8_t)f};
        begin = pos + 1;
        flags = f;
        absorb(hdr, 2);
        if ((f & (4 | 32)) && pos) end_block();
    }

public:
    explicit Strobe128(const char* protocol) {
        std::memset(st, 0, sizeof st);
        const uint8_t head[6] = {1, RATE + 2, 1, 0, 1, 96};
        std::memcpy(st, head, 6);
        std::memcpy(st + 6, "STROBEv1.0.2", 12);
        permute();
        meta_ad((const uint8_t*)protocol, std::strlen(protocol), false);
    }
    void meta_ad(const uint8_t* d, size_t n, bool more) { begin_op(16 | 2, more); absorb(d, n); }
    void ad(const uint8_t* d, size_t n, bool more) { begin_op(2, more); absorb(d, n); }
    void prf(uint8_t* out, size_t n) { begin_op(1 | 2 | 4, false); squeeze(out, n); }
};

class MerlinTranscript {  // merlin::Transcript: new / append_message / challenge_bytes
    Strobe128 strobe;
    void framed(const char* label, uint32_t len) {
        strobe.meta_ad((const uint8_t*)label, std::strlen(label), false);
        uint8_t le[4] = {(uint8_t)len, (uint8_t)(len >> 8), (uint8_t)(len >> 16), (uint8_t)(len >> 24)};
        strobe.meta_ad(le, 4, true);
    }

public:
    explicit MerlinTranscript(const char* label) : strobe("Merlin v1.0") {
        append_message("dom-sep", (const uint8_t*)label, std::strlen(label));
    }
    void append_message(const char* label, const uint8_t* msg, size_t n) {
        framed(label, (uint32_t)n);
        strobe.ad(msg, n, false);
    }
    void challenge_bytes(const char* label, uint8_t* out, size_t n) {
        framed(label, (uint32_t)n);
        strobe.prf(out, n);
    }
};

class PlonkTranscript : public MerlinTranscript {  // src/transcript.rs:65-86, Transcript::new(b"plonk") (src/prover.rs:112)
public:
    PlonkTranscript() : MerlinTranscript("plonk") {}
    void append_point(const char* label, const std::array<uint8_t, 48>& c) { append_message(label, c.data(), 48); }
    void append_scalar(const char* label, const Scalar& s);
    Scalar get_and_append_challenge(const char* label);
};

// ---- encodings ------------------------------------------------------------------------------------------
inline std::array<uint8_t, 32> scalar_to_bytes(const Scalar& s) {  // scalar.rs:292-304: canonical, little-endian
    bpk::fr_t c = bpk::from_mont(Scalar::to_fe(s));
    std::array<uint8_t, 32> out;
    for (int i = 0; i < 8; i++)
        for (int k = 0; k < 4; k++) out[4 * i + k] = (uint8_t)(c.l[i] >> (8 * k));
    return out;
}
inline bool scalar_from_bytes(const uint8_t b[32], Scalar* out) {  // scalar.rs:238-275: None unless canonical
    bpk::fr_t c;
    for (int i = 0; i < 8; i++) c.l[i] = (uint32_t)b[4 * i] | ((uint32_t)b[4 * i + 1] << 8) | ((uint32_t)b[4 * i + 2] << 16) |
                                         ((uint32_t)b[4 * i + 3] << 24);
    for (int i = 7; i >= 0; i--) {  // c < q ?
        uint32_t q = bpk::FrParams::mod(i);
        if (c.l[i] < q) break;
        if (c.l[i] > q) return false;
        if (i == 0) return false;  // equal to q
    }
    *out = Scalar::from_fe(bpk::to_mont(c));
    return true;
}
inline void PlonkTranscript::append_scalar(const char* label, const Scalar& s) {
    auto b = scalar_to_bytes(s);
    append_message(label, b.data(), 32);
}
inline Scalar PlonkTranscript::get_and_append_challenge(const char* label) {
    for (;;) {
        uint8_t raw[32];
        challenge_bytes(label, raw, 32);
        Scalar v;
        if (scalar_from_bytes(raw, &v) && v != Scalar::zero()) {
            append_message(label, raw, 32);
            return v;
        }
    }
}
inline std::array<uint8_t, 48> g1_to_compressed(const G1Projective& p) {  // g1.rs:221-242 on a normalised point
    std::array<uint8_t, 48> out{};
    if (G1Projective::fe(p.z).is_zero()) {
        out[0] = 0xC0;
        return out;
    }
    bpk::fp_t zi = bpk::inv(G1Projective::fe(p.z));
    bpk::fp_t xm = bpk::mul(G1Projective::fe(p.x), zi), ym = bpk::mul(G1Projective::fe(p.y), zi);
    bpk::fp_t x = bpk::from_mont(xm), y = bpk::from_mont(ym), ny = bpk::from_mont(bpk::neg(ym));
    for (int i = 0; i < 12; i++)
        for (int k = 0; k < 4; k++) out[47 - (4 * i + k)] = (uint8_t)(x.l[i] >> (8 * k));
    out[0] |= 0x80;
    bool largest = false;  // y > (p - 1) / 2  <=>  y > p - y
    for (int i = 11; i >= 0; i--) {
        if (y.l[i] != ny.l[i]) {
            largest = y.l[i] > ny.l[i];
            break;
        }
    }
    if (largest) out[0] |= 0x20;
    return out;
}

// ---- Proof (src/verifier.rs:23-40 field order) ----------------------------------------------------------
struct Proof {
    std::array<uint8_t, 48> a_1, b_1, c_1, z_1, t_lo_1, t_mid_1, t_hi_1, w_zeta_1, w_zeta_omega_1;
    Scalar a_bar, b_bar, c_bar, s1_bar, s2_bar, z_omega_bar;
    std::array<uint8_t, 624> to_bytes() const {
        std::array<uint8_t, 624> out;
        const std::array<uint8_t, 48>* pts[9] = {&a_1, &b_1, &c_1, &z_1, &t_lo_1, &t_mid_1, &t_hi_1, &w_zeta_1, &w_zeta_omega_1};
        for (int i = 0; i < 9; i++) std::memcpy(out.data() + 48 * i, pts[i]->data(), 48);
        const Scalar* sc[6] = {&a_bar, &b_bar, &c_bar, &s1_bar, &s2_bar, &z_omega_bar};
        for (int i = 0; i < 6; i++) {
            auto b = scalar_to_bytes(*sc[i]);
            std::memcpy(out.data() + 432 + 32 * i, b.data(), 32);
        }
        return out;
    }
};

// ---- device memory ----------------------------------------------------------------------------------------
class DevScalars {  // n Scalars in HBM
    char* p = nullptr;
    size_t n_ = 0;

public:
    DevScalars() = default;
    explicit DevScalars(size_t n, bool zero = false) : n_(n) {
        void* q = nullptr;
        check(bpk_dev_alloc(ctx(), n * 32, &q), "bpk_dev_alloc");
        p = (char*)q;
        if (zero) check(bpk_dev_zero(ctx(), p, n * 32), "bpk_dev_zero");
    }
    DevScalars(const DevScalars&) = delete;
    DevScalars& operator=(const DevScalars&) = delete;
    DevScalars(DevScalars&& o) noexcept : p(o.p), n_(o.n_) { o.p = nullptr; }
    DevScalars& operator=(DevScalars&& o) noexcept {
        if (this != &o) {
            if (p) bpk_dev_free(ctx(), p);
            p = o.p;
            n_ = o.n_;
            o.p = nullptr;
        }
        return *this;
    }
    ~DevScalars() {
        if (p) bpk_dev_free(ctx(), p);
    }
    void* at(size_t i = 0) const { return p + 32 * i; }
    size_t size() const { return n_; }
    void upload(size_t at_index, const Scalar* src, size_t count) {
        check(bpk_dev_upload(ctx(), at(at_index), src, count * 32), "bpk_dev_upload");
    }
    Scalar get(size_t i) const {
        Scalar s;
        check(bpk_dev_download(ctx(), s.l, at(i), 32), "bpk_dev_download");
        return s;
    }
};

// ---- Prover { group_order, setup, pk } (src/prover.rs:90-104) on one GPU ---------------------------------
class DeviceProver {
    uint64_t srs;
    size_t n, D, L, ratio;
    Scalar omega, shift, k1, k2;
    DevScalars pk_lagrange;  // 8 x n: ql qr qm qo qc s1 s2 s3 on H
    std::vector<Scalar> zh_inv;
    bool cache;
    DevScalars pre_coeffs, pre_evals;  // 8 x n coefficient forms, 10 x D coset evaluations (ql..qc s1..s3 L1 X)
    bool have_pre = false;

    static void vec(int op, void* a, const void* b, const Scalar* s, void* out, size_t cnt) {
        check(bpk_fr_vec_op(ctx(), op, a, b, s ? s->l : nullptr, out, cnt), "bpk_fr_vec_op");
    }
    static void axpy(void* acc, const Scalar& s, const void* p, size_t len) {  // acc[:len] += s * p[:len]
        if (s == Scalar::zero()) return;
        vec(4, acc, p, &s, acc, len);
    }
    static void add_const(void* elem, const Scalar& s) { vec(5, elem, nullptr, &s, elem, 1); }
    static void ntt(const void* in, void* out, size_t len, size_t batch, int flags, const Scalar* sh) {
        check(bpk_ntt_fr_dev(ctx(), in, out, len, batch, flags, sh ? sh->l : nullptr), "bpk_ntt_fr_dev");
    }
    static Scalar eval(const void* coeffs, size_t len, const Scalar& x) {
        Scalar out;
        check(bpk_fr_poly_eval(ctx(), coeffs, len, x.l, out.l), "bpk_fr_poly_eval");
        return out;
    }
    std::vector<std::array<uint8_t, 48>> commit_many(const std::vector<std::pair<const void*, size_t>>& items) const {
        const size_t k = items.size();
        std::vector<const void*> ptrs(k);
        std::vector<size_t> firsts(k, 0), lens(k);
        for (size_t i = 0; i < k; i++) {
            ptrs[i] = items[i].first;
            lens[i] = items[i].second;
        }
        DevScalars out((k * 18 * 8 + 31) / 32);
        check(bpk_msm_g1_dev_batch(ctx(), srs, k, ptrs.data(), firsts.data(), lens.data(), 1, out.at()),
              "bpk_msm_g1_dev_batch");
        std::vector<G1Projective> pts(k);
        check(bpk_dev_download(ctx(), pts.data(), out.at(), k * sizeof(G1Projective)), "bpk_dev_download");
        std::vector<std::array<uint8_t, 48>> res(k);
        for (size_t i = 0; i < k; i++) res[i] = g1_to_compressed(pts[i]);
        return res;
    }
    void preprocess() {  // per-circuit data of round 3 (the reference recomputes the i_ntt's on every prove)
        if (have_pre) return;
        pre_coeffs = DevScalars(8 * n);
        ntt(pk_lagrange.at(), pre_coeffs.at(), n, 8, 1, nullptr);
        pre_evals = DevScalars(10 * D, true);
        for (int r = 0; r < 8; r++)
            check(bpk_dev_copy(ctx(), pre_evals.at(r * D), pre_coeffs.at(r * n), n * 32), "bpk_dev_copy");
        std::vector<Scalar> l1(n, Scalar::from(n).invert());  // L1 = (1/n) sum X^i
        pre_evals.upload(8 * D, l1.data(), n);
        Scalar one = Scalar::one();
        pre_evals.upload(9 * D + 1, &one, 1);  // the polynomial X
        ntt(pre_evals.at(), pre_evals.at(), D, 10, 2, &shift);
        have_pre = cache;
    }

public:
    DeviceProver(const Setup& setup, size_t group_order, const std::vector<std::vector<Scalar>>& selectors,
                 const std::vector<std::vector<Scalar>>& sigmas, bool cache_preprocessed = false)
        : srs(setup.handle), n(group_order), cache(cache_preprocessed) {
        if (!is_power_of_two(n)) throw Panic("assertion failed: is_power_of_two(group_order)");
        if (setup.powers_of_x.size() < n + 6) throw Panic("SRS too short");
        if (selectors.size() != 5 || sigmas.size() != 3) throw Panic("5 selector and 3 sigma columns expected");
        D = 1;
        while (D < 3 * n + 6) D <<= 1;
        ratio = D / n;
        L = n + 8;
        omega = root_of_unity(n);
        shift = Scalar::from(7);  // multiplicative generator of Fr: outside every 2-power subgroup
        k1 = Scalar::from(2);
        k2 = Scalar::from(3);
        pk_lagrange = DevScalars(8 * n);
        for (int r = 0; r < 8; r++) {
            const std::vector<Scalar>& col = r < 5 ? selectors[r] : sigmas[r - 5];
            if (col.size() != n) throw Panic("column length != group order");
            pk_lagrange.upload(r * n, col.data(), n);
        }
        Scalar gn = shift.pow(n), wn = root_of_unity(D).pow(n), p = Scalar::one();
        for (size_t i = 0; i < ratio; i++) {
            zh_inv.push_back((gn * p - Scalar::one()).invert());
            p = p * wn;
        }
    }

    // wires = the (A, B, C) witness columns on H; blinding = b_1 .. b_11 (thread_rng in the reference, prover.rs:108-110)
    Proof prove(const std::vector<std::vector<Scalar>>& wires, const std::vector<Scalar>& public_inputs,
                const std::vector<Scalar>& b) {
        if (wires.size() != 3 || b.size() != 11) throw Panic("3 wire columns and 11 blinding scalars expected");
        PlonkTranscript tr;
        Proof pf;
        const Scalar blind_h[11] = {b[1], b[0], b[3], b[2], b[5], b[4], b[8], b[7], b[6], b[9], b[10]};
        DevScalars blind(11);
        blind.upload(0, blind_h, 11);
        DevScalars wv(5 * D, true);  // rows a b c z PI, zero-padded to the quotient domain
        auto row = [&](int r) { return (char*)wv.at((size_t)r * D); };

        // ---- round 1 (prover.rs:177-277)
        DevScalars W(3 * n), tmp(3 * n);
        for (int k = 0; k < 3; k++) {
            if (wires[k].size() != n) throw Panic("wire column length != group order");
            W.upload(k * n, wires[k].data(), n);
        }
        ntt(W.at(), tmp.at(), n, 3, 1, nullptr);
        for (int k = 0; k < 3; k++) {
            check(bpk_dev_copy(ctx(), row(k), tmp.at(k * n), n * 32), "bpk_dev_copy");
            vec(1, row(k), blind.at(2 * k), nullptr, row(k), 2);                                  // -(b_lo + b_hi X)
            check(bpk_dev_copy(ctx(), row(k) + 32 * n, blind.at(2 * k), 64), "bpk_dev_copy");  // +(..) X^n
        }
        {
            auto c = commit_many({{row(0), n + 2}, {row(1), n + 2}, {row(2), n + 2}});
            pf.a_1 = c[0];
            pf.b_1 = c[1];
            pf.c_1 = c[2];
        }
        tr.append_point("a_1", pf.a_1);
        tr.append_point("b_1", pf.b_1);
        tr.append_point("c_1", pf.c_1);
        const Scalar beta = tr.get_and_append_challenge("beta");
        const Scalar gamma = tr.get_and_append_challenge("gamma");

        // ---- round 2 (prover.rs:279-368)
        DevScalars Z(n + 1);
        check(bpk_plonk_grand_product(ctx(), W.at(0), W.at(n), W.at(2 * n), pk_lagrange.at(5 * n), pk_lagrange.at(6 * n),
                                      pk_lagrange.at(7 * n), n, beta.l, gamma.l, k1.l, k2.l, Z.at()),
              "bpk_plonk_grand_product");
        if (Z.get(n) != Scalar::one()) throw Panic("assertion `left == right` failed: z_values.pop() == Scalar::one()");
        char* z = row(3);
        ntt(Z.at(), z, n, 1, 1, nullptr);
        vec(1, z, blind.at(6), nullptr, z, 3);
        check(bpk_dev_copy(ctx(), z + 32 * n, blind.at(6), 96), "bpk_dev_copy");
        pf.z_1 = commit_many({{z, n + 3}})[0];
        tr.append_point("z_1", pf.z_1);
        const Scalar alpha = tr.get_and_append_challenge("z_1");  // sic: src/transcript.rs:24

        // ---- round 3 (prover.rs:370-500)
        preprocess();
        auto pkc = [&](int r) { return pre_coeffs.at((size_t)r * n); };  // ql qr qm qo qc s1 s2 s3
        {
            std::vector<Scalar> pi(n, Scalar::zero());
            for (size_t i = 0; i < public_inputs.size(); i++) pi[i] = public_inputs[i].neg();
            DevScalars pi_l(n);
            pi_l.upload(0, pi.data(), n);
            ntt(pi_l.at(), row(4), n, 1, 1, nullptr);
        }
        DevScalars keep(5 * L);  // coefficient forms for rounds 4-5
        for (int r = 0; r < 5; r++) check(bpk_dev_copy(ctx(), keep.at(r * L), row(r), L * 32), "bpk_dev_copy");
        ntt(wv.at(), wv.at(), D, 5, 2, &shift);
        DevScalars t(D);
        {
            std::vector<uint64_t> zh(4 * ratio);
            for (size_t i = 0; i < ratio; i++) std::memcpy(&zh[4 * i], zh_inv[i].l, 32);
            check(bpk_plonk_quotient_evals(ctx(), wv.at(), pre_evals.at(), D, n, beta.l, gamma.l, alpha.l, k1.l, k2.l,
                                           zh.data(), t.at()),
                  "bpk_plonk_quotient_evals");
        }
        ntt(t.at(), t.at(), D, 1, 3, &shift);
        auto kc = [&](int r) { return (char*)keep.at((size_t)r * L); };
        z = kc(3);
        // split_t_to_3pieces (prover.rs:454-500)
        DevScalars parts(3 * L, true);
        char *t_lo = (char*)parts.at(0), *t_mid = (char*)parts.at(L), *t_hi = (char*)parts.at(2 * L);
        check(bpk_dev_copy(ctx(), t_lo, t.at(0), n * 32), "bpk_dev_copy");
        check(bpk_dev_copy(ctx(), t_lo + 32 * n, blind.at(9), 32), "bpk_dev_copy");
        check(bpk_dev_copy(ctx(), t_mid, t.at(n), n * 32), "bpk_dev_copy");
        vec(1, t_mid, blind.at(9), nullptr, t_mid, 1);
        check(bpk_dev_copy(ctx(), t_mid + 32 * n, blind.at(10), 32), "bpk_dev_copy");
        check(bpk_dev_copy(ctx(), t_hi, t.at(2 * n), (n + 6) * 32), "bpk_dev_copy");
        vec(1, t_hi, blind.at(10), nullptr, t_hi, 1);
        {
            auto c = commit_many({{t_lo, n + 1}, {t_mid, n + 1}, {t_hi, n + 6}});
            pf.t_lo_1 = c[0];
            pf.t_mid_1 = c[1];
            pf.t_hi_1 = c[2];
        }
        tr.append_point("t_lo_1", pf.t_lo_1);
        tr.append_point("t_mid_1", pf.t_mid_1);
        tr.append_point("t_hi_1", pf.t_hi_1);
        const Scalar zeta = tr.get_and_append_challenge("zeta");

        // ---- round 4 (prover.rs:502-541)
        Scalar pi_zeta;
        {
            const void* ptrs[6] = {kc(0), kc(1), kc(2), pkc(5), pkc(6), kc(4)};
            const size_t lens[6] = {n + 2, n + 2, n + 2, n, n, n};
            Scalar ev[6];
            check(bpk_fr_poly_eval_many(ctx(), 6, ptrs, lens, zeta.l, ev[0].l), "bpk_fr_poly_eval_many");
            pf.a_bar = ev[0];
            pf.b_bar = ev[1];
            pf.c_bar = ev[2];
            pf.s1_bar = ev[3];
            pf.s2_bar = ev[4];
            pi_zeta = ev[5];
        }
        pf.z_omega_bar = eval(z, n + 3, zeta * omega);
        tr.append_scalar("a_eval", pf.a_bar);
        tr.append_scalar("b_eval", pf.b_bar);
        tr.append_scalar("c_eval", pf.c_bar);
        tr.append_scalar("s1_eval", pf.s1_bar);
        tr.append_scalar("s2_eval", pf.s2_bar);
        tr.append_scalar("z_shifted_eval", pf.z_omega_bar);
        const Scalar nu = tr.get_and_append_challenge("nu");

        // ---- round 5 (prover.rs:543-647)
        const Scalar one = Scalar::one();
        const Scalar zeta_n = zeta.pow(n), zh_zeta = zeta_n - one;
        const Scalar l1_zeta = zh_zeta * (Scalar::from(n) * (zeta - one)).invert();
        const Scalar f = (pf.a_bar + zeta * beta + gamma) * (pf.b_bar + zeta * beta * k1 + gamma) *
                         (pf.c_bar + zeta * beta * k2 + gamma);
        const Scalar g = (pf.a_bar + pf.s1_bar * beta + gamma) * (pf.b_bar + pf.s2_bar * beta + gamma) * pf.z_omega_bar;
        const Scalar a2 = alpha * alpha;
        DevScalars r(L, true);
        axpy(r.at(), pf.a_bar * pf.b_bar, pkc(2), n);
        axpy(r.at(), pf.a_bar, pkc(0), n);
        axpy(r.at(), pf.b_bar, pkc(1), n);
        axpy(r.at(), pf.c_bar, pkc(3), n);
        axpy(r.at(), one, pkc(4), n);
        axpy(r.at(), alpha * f + a2 * l1_zeta, z, n + 3);
        axpy(r.at(), (alpha * g * beta).neg(), pkc(7), n);
        axpy(r.at(), zh_zeta.neg(), t_lo, n + 1);
        axpy(r.at(), (zh_zeta * zeta_n).neg(), t_mid, n + 1);
        axpy(r.at(), (zh_zeta * zeta_n * zeta_n).neg(), t_hi, n + 6);
        add_const(r.at(), pi_zeta - alpha * g * (pf.c_bar + gamma) - a2 * l1_zeta);
        if (eval(r.at(), n + 6, zeta) != Scalar::zero())
            throw Panic("assertion `left == right` failed: r.coeffs_evaluate(zeta) == Scalar::zero()");
        Scalar nus[6];
        nus[0] = one;
        for (int k = 1; k < 6; k++) nus[k] = nus[k - 1] * nu;
        axpy(r.at(), nus[1], kc(0), n + 2);
        axpy(r.at(), nus[2], kc(1), n + 2);
        axpy(r.at(), nus[3], kc(2), n + 2);
        axpy(r.at(), nus[4], pkc(5), n);
        axpy(r.at(), nus[5], pkc(6), n);
        add_const(r.at(), (nus[1] * pf.a_bar + nus[2] * pf.b_bar + nus[3] * pf.c_bar + nus[4] * pf.s1_bar +
                           nus[5] * pf.s2_bar).neg());
        DevScalars w_zeta(L), w_zeta_omega(L);
        check(bpk_fr_poly_div_linear(ctx(), r.at(), n + 6, zeta.l, w_zeta.at()), "bpk_fr_poly_div_linear");
        add_const(z, pf.z_omega_bar.neg());
        check(bpk_fr_poly_div_linear(ctx(), z, n + 3, (zeta * omega).l, w_zeta_omega.at()), "bpk_fr_poly_div_linear");
        {
            auto c = commit_many({{w_zeta.at(), n + 5}, {w_zeta_omega.at(), n + 2}});
            pf.w_zeta_1 = c[0];
            pf.w_zeta_omega_1 = c[1];
        }
        tr.append_point("w_zeta_1", pf.w_zeta_1);
        tr.append_point("w_zeta_omega_1", pf.w_zeta_omega_1);
        (void)tr.get_and_append_challenge("mu");
        if (!cache) {
            pre_coeffs = DevScalars();
            pre_evals = DevScalars();
        }
        return pf;
    }
};

}  // namespace baby_plonk
