// Host-side exerciser for the csrc/ff.cuh + ec.cuh templates (carry flag emulated in C).
// Reads "<field> <op> <hexA> [hexB]" lines, prints the result limbs as one big-endian hex number.
// Driven by tests/test_host_templates.py, which compares against the Python oracle.
#include <cstdio>
#include <cstring>
#include <string>
#include <type_traits>
#include <iostream>
#include <sstream>
#include "../../baby-plonk-rust_b200/csrc/ff.cuh"
#include "../../baby-plonk-rust_b200/csrc/ec.cuh"
using namespace bpk;

template <class F> static F parse(const std::string& h) {
    F r = F::zero();
    int n = (int)h.size();
    for (int i = 0; i < n; i++) {
        char c = h[n - 1 - i];
        uint32_t v = (c >= '0' && c <= '9') ? c - '0' : (c >= 'a' && c <= 'f') ? c - 'a' + 10 : c - 'A' + 10;
        if (i / 8 < F::N) r.l[i / 8] |= v << (4 * (i % 8));
    }
    return r;
}
template <class F> static void show(const F& a) {
    for (int i = F::N - 1; i >= 0; i--) printf("%08x", a.l[i]);
}
template <class F> static int run_field(const std::string& op, std::istringstream& ss) {
    std::string ha, hb;
    ss >> ha;
    F a = parse<F>(ha), b = F::zero();
    if (op.substr(0, 3) == "mul" || op.substr(0, 3) == "add" || op.substr(0, 3) == "sub") { ss >> hb; b = parse<F>(hb); }
    F r;
    if (op == "mul") r = mul(a, b);
    else if (op == "mulcc") r = mul_cc(a, b);
    else if (op == "mullazy") r = reduce_once(mul_cc<typename F::params, false>(a, b));
    else if (op == "addlazy") r = reduce_once(add_lazy(a, b));
    else if (op == "sublazy") r = reduce_once(sub_lazy(a, b));
    else if (op == "sqr") r = sqr(a);
    else if (op == "add") r = add(a, b);
    else if (op == "sub") r = sub(a, b);
    else if (op == "neg") r = neg(a);
    else if (op == "inv") r = inv(a);
    else if (op == "from_mont") r = from_mont(a);
    else if (op == "to_mont") r = to_mont(a);
    else return 1;
    show(r); printf("\n");
    return 0;
}
static xyzz_t parse_xyzz(std::istringstream& ss) {
    std::string a, b, c, d; ss >> a >> b >> c >> d;
    xyzz_t p; p.X = parse<fp_t>(a); p.Y = parse<fp_t>(b); p.ZZ = parse<fp_t>(c); p.ZZZ = parse<fp_t>(d);
    return p;
}
static void show_xyzz(const xyzz_t& p) {
    show(p.X); printf(" "); show(p.Y); printf(" "); show(p.ZZ); printf(" "); show(p.ZZZ); printf("\n");
}
int main() {
    std::string line;
    while (std::getline(std::cin, line)) {
        std::istringstream ss(line);
        std::string field, op;
        ss >> field >> op;
        if (field == "fr") { if (run_field<fr_t>(op, ss)) return 1; }
        else if (field == "fp") { if (run_field<fp_t>(op, ss)) return 1; }
        else if (field == "g1") {
            // all coordinates Montgomery limbs as hex
            if (op == "madd") {           // xyzz + affine(x,y)
                xyzz_t p = parse_xyzz(ss); std::string hx, hy; ss >> hx >> hy;
                affine_t q; q.x = parse<fp_t>(hx); q.y = parse<fp_t>(hy);
                xyzz_madd(p, q); show_xyzz(p);
            } else if (op == "add") {
                xyzz_t p = parse_xyzz(ss); xyzz_t q = parse_xyzz(ss);
                xyzz_add(p, q); show_xyzz(p);
            } else if (op == "dbl") {
                xyzz_t p = parse_xyzz(ss); xyzz_dbl(p); show_xyzz(p);
            } else if (op == "to_affine") {
                xyzz_t p = parse_xyzz(ss); affine_t q = xyzz_to_affine(p);
                show(q.x); printf(" "); show(q.y); printf("\n");
            } else if (op == "aff_sum") {    // affine + affine -> XYZZ (generic case only)
                std::string a, b, c, d; ss >> a >> b >> c >> d;
                xyzz_t r = xyzz_from_affine_sum(parse<fp_t>(a), parse<fp_t>(b), parse<fp_t>(c), parse<fp_t>(d));
                show_xyzz(r);
            } else if (op == "aff_add") {    // batched-affine addition of one pair: prepare, invert, finish
                std::string a, b, c, d; ss >> a >> b >> c >> d;
                affine_t p, q; p.x = parse<fp_t>(a); p.y = parse<fp_t>(b); q.x = parse<fp_t>(c); q.y = parse<fp_t>(d);
                fp_t den;
                int kind = affine_add_prepare(p, q, den);
                affine_t r = affine_add_finish(kind, p, q, inv(den));
                // the lazy form, fed an inverse that is itself left in [0, 2p)
                fp_t di = inv(den), di2 = add(di, di);   // (2 di) / 2 ... simply use di + p when it fits below 2p
                fp_t dil = di;
                {   // di + p as a non-canonical representative of the same residue
                    uint64_t c = 0;
                    for (int i = 0; i < 12; i++) { uint64_t t = (uint64_t)di.l[i] + FpParams::mod(i) + c; dil.l[i] = (uint32_t)t; c = t >> 32; }
                }
                (void)di2;
                affine_t r2 = affine_add_finish_lazy(kind, p, q, dil);
                if (!(r.x == r2.x) || !(r.y == r2.y)) { printf("LAZY MISMATCH\n"); return 1; }
                printf("%d ", kind); show(r.x); printf(" "); show(r.y); printf("\n");
            } else if (op == "proj_to_affine") {
                std::string a, b, c; ss >> a >> b >> c;
                affine_t q = proj_to_affine(parse<fp_t>(a), parse<fp_t>(b), parse<fp_t>(c));
                show(q.x); printf(" "); show(q.y); printf("\n");
            } else return 1;
        } else return 1;
    }
    return 0;
}
