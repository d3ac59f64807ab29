// tests/cpp/host_layer.cpp — drives include/zkb200.hpp (the C++ mirror of SRS<G> / Radix2EvaluationDomain / the open rounds) the way
// a C++ caller would; inputs and outputs are flat little-endian u64 files exchanged with tests/test_gpu_cpp_layer.py.
//   in : n, then g (n x 8), h (8), coeffs (n + n/2) x 4, evals n x 4, blinders 2 x 4, u 4, u_inv 4
//   out: commit chunks 2 x 8, masked 2 x 8, commit_evaluations 8, fft of the first n coefficients n x 4 (and the ifft back, n x 4),
//        msm_bigint(coeffs[0..n]) as affine 8, round-0 L and R as affine 2 x 8, ip_l 4, ip_r 4, a0 4, b0 4, sg 8 after ALL rounds with u,
//        Srs::open of the first n coefficients with a stand-in sponge (U = g[7], every challenge = u): l_0, r_0, delta, sg 4 x 8, z1, z2 2 x 4
#include <cstdio>
#include <cstring>
#include <vector>

#include "zkb200.hpp"

using namespace zkb200;

static std::vector<uint64_t> slurp(const char* path) {
    std::vector<uint64_t> v;
    FILE* f = std::fopen(path, "rb");
    if (!f) return v;
    std::fseek(f, 0, SEEK_END);
    long sz = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    v.resize((size_t)sz / 8);
    if (std::fread(v.data(), 8, v.size(), f) != v.size()) v.clear();
    std::fclose(f);
    return v;
}
template <size_t K> static std::vector<std::array<uint64_t, K>> take(const std::vector<uint64_t>& src, size_t& pos, size_t count) {
    std::vector<std::array<uint64_t, K>> out(count);
    for (size_t i = 0; i < count; i++, pos += K) std::memcpy(out[i].data(), &src[pos], K * 8);
    return out;
}
template <class A> static void put(std::vector<uint64_t>& dst, const A& a) { dst.insert(dst.end(), a.begin(), a.end()); }

int main(int argc, char** argv) {
    if (argc != 3) return 2;
    try {
        const std::vector<uint64_t> in = slurp(argv[1]);
        if (in.empty()) return 3;
        size_t pos = 0;
        const size_t n = (size_t)in[pos++];
        auto g = take<8>(in, pos, n);
        const Affine h = take<8>(in, pos, 1)[0];
        auto coeffs = take<4>(in, pos, n + n / 2);
        auto evals = take<4>(in, pos, n);
        auto blinders = take<4>(in, pos, 2);
        const Fe u = take<4>(in, pos, 1)[0], u_inv = take<4>(in, pos, 1)[0];
        std::vector<uint64_t> out;

        Context ctx(0);
        Srs srs(ctx, ZK_PALLAS, g, h);
        PolyComm c = srs.commit_non_hiding(coeffs, 1);                  // n + n/2 coefficients: 2 chunks (ipa.rs:663-676)
        if (c.len() != 2) return 4;
        for (auto& ch : c.chunks) put(out, ch);
        BlindedCommitment m = srs.mask_custom(c, blinders);
        for (auto& ch : m.commitment.chunks) put(out, ch);
        bool threw = false;
        try { srs.mask_custom(c, {blinders[0]}); } catch (const BlindersDontMatch& e) { threw = e.blinders == 1 && e.commitment == 2; }
        if (!threw) return 5;
        srs.get_lagrange_basis_from_domain_size(n);
        put(out, srs.commit_evaluations_non_hiding(n, evals).chunks[0]);

        Radix2EvaluationDomain d(ctx, ZK_FQ, n);                        // Pallas scalars live in Fq
        std::vector<Fe> poly(coeffs.begin(), coeffs.begin() + (long)n);
        d.fft_in_place(poly);
        for (auto& e : poly) put(out, e);
        d.ifft_in_place(poly);
        for (auto& e : poly) put(out, e);

        Bases bases(ctx, ZK_PALLAS, g);
        std::vector<Fe> first(coeffs.begin(), coeffs.begin() + (long)n);
        put(out, into_affine(ZK_PALLAS, bases.msm(first)));            // Montgomery scalars, like ipa.rs:649

        IpaRounds rounds(ctx, bases, first, evals);
        bool first_round = true;
        while (rounds.len() > 1) {
            IpaRounds::LR lr = rounds.lr();
            if (first_round) {
                put(out, into_affine(ZK_PALLAS, lr.l));
                put(out, into_affine(ZK_PALLAS, lr.r));
                put(out, lr.ip_l);
                put(out, lr.ip_r);
                first_round = false;
            }
            rounds.fold(u, u_inv);
        }
        IpaRounds::Final fin = rounds.finish();
        put(out, fin.a0);
        put(out, fin.b0);
        put(out, fin.sg);

        // OpenProof::open through the C++ layer: the three callbacks are the caller's sponge (here a fixed derivation, the same the
        // Python side of the test uses through ctypes)
        OpenTranscript sponge;
        sponge.u_base = [&](const Fe&) { return g[7]; };
        sponge.round = [&](unsigned, const Affine&, const Affine&) { return u; };
        sponge.final_challenge = [&](const Affine&) { return u; };
        OpenPolynomial pl;
        pl.data = first.data(); pl.len = n; pl.domain_size = 0; pl.blinders = {blinders[0]};
        size_t k = 0;
        while ((size_t(1) << k) < n) k++;
        std::vector<Fe> draws(evals.begin(), evals.begin() + (long)(2 * k + 2));
        OpeningProof pr = srs.open({pl}, {u, u_inv}, u, u_inv, draws, sponge);
        if (pr.lr.size() != k) return 7;
        put(out, pr.lr[0].first);
        put(out, pr.lr[0].second);
        put(out, pr.delta);
        put(out, pr.sg);
        put(out, pr.z1);
        put(out, pr.z2);

        FILE* f = std::fopen(argv[2], "wb");
        if (!f) return 6;
        std::fwrite(out.data(), 8, out.size(), f);
        std::fclose(f);
        std::printf("ok %zu launches\n", (size_t)ctx.launch_count());
        return 0;
    } catch (const Error& e) {
        std::fprintf(stderr, "zkb200::Error %d: %s\n", e.code, e.what());
        return 1;
    }
}
