// tc5_program_check.cpp — CPU check of the tcgen05 step program (csrc/model_prep.hpp: prepare_tc5_program).
//
// Interprets the program exactly as k_prune_tc5 does — chain starts, one GEMM per step, MUL / PUSH_START / POP_MUL post-ops,
// leaf sources consumed in ring order, cherry sources in table order — but in FP64 on the host, with the cherry tables computed
// from their definition, and compares log z with a plain post-order Felsenstein recursion over the same P matrices on random
// codon columns.  A second pass runs the program in FP32 with split-TF32 products (what the kernel's arithmetic amounts to) to
// put a number on the rounding of the FP32-class path in decibans.  No GPU involved: this pins the host-side program logic.
//
// build: g++ -O2 -std=c++17 -o /tmp/tc5_program_check tools/tc5_program_check.cpp
// usage: tc5_program_check <model> [species-list] [n_windows] [gap_fraction]
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <random>
#include <set>
#include <memory>

#include "../phylocsfpp_b200/host/util.hpp"
#include "../phylocsfpp_b200/host/model.hpp"
#include "../phylocsfpp_b200/csrc/model_prep.hpp"

using namespace pcsf;

static void leaf_msg(const EcmHost &e, int leaf, int x, double *out) {
    for (int a = 0; a < 64; ++a) out[a] = x == 64 ? 1.0 : e.P[(size_t)leaf * 4096 + a * 64 + x];
}

// plain recursion: alpha of node i
static void alpha_ref(const ModelHost &m, const EcmHost &e, const uint8_t *ids, int i, double *out) {
    if (m.child1[i] < 0) { for (int a = 0; a < 64; ++a) out[a] = 0.0; return; }
    double acc[64];
    for (int a = 0; a < 64; ++a) acc[a] = 1.0;
    for (int c : {(int)m.child1[i], (int)m.child2[i]}) {
        double msg[64];
        if (m.child1[c] < 0) leaf_msg(e, c, ids[c], msg);
        else {
            double al[64];
            alpha_ref(m, e, ids, c, al);
            for (int a = 0; a < 64; ++a) {
                double s = 0.0;
                for (int b = 0; b < 64; ++b) s += e.P[(size_t)c * 4096 + a * 64 + b] * al[b];
                msg[a] = s;
            }
        }
        for (int a = 0; a < 64; ++a) acc[a] *= msg[a];
    }
    for (int a = 0; a < 64; ++a) out[a] = acc[a];
}

static float tf32_hi(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xffffe000u; float r; memcpy(&r, &u, 4); return r; }

struct Interp {
    const ModelHost &m;
    const EcmHost &e;
    const uint8_t *ids;
    size_t leaf_pos = 0, cherry_pos = 0, src_pos = 0;
    bool f32;
    Interp(const ModelHost &mm, const EcmHost &ee, const uint8_t *i, bool f) : m(mm), e(ee), ids(i), f32(f) {}
    void source(uint32_t s, double *out) {
        const Tc5Src &d = m.tc5_srcs[src_pos++];          // the kernel walks this list: it must describe the same source
        if ((s & T5_SRC_CHERRY) ? (d.l2 == 0xff || d.cherry != (s & 0x7fu) || d.l1 != (m.tc5_cherry_leaves[d.cherry] & 0xff) ||
                                   d.l2 != (m.tc5_cherry_leaves[d.cherry] >> 8))
                                : (d.l2 != 0xff || d.l1 != s))
            host::die("source list out of step with the program");
        if (s & T5_SRC_CHERRY) {
            const size_t k = cherry_pos++;
            if ((s & 0x7fu) != k) host::die("cherry consumed out of table order");
            const int c = m.tc5_cherries[k], l = m.tc5_cherry_leaves[k] & 0xff, r = m.tc5_cherry_leaves[k] >> 8;
            if (l != m.child1[c] || r != m.child2[c]) host::die("cherry leaves");
            double lm[64], rm[64];
            leaf_msg(e, l, ids[l], lm); leaf_msg(e, r, ids[r], rm);
            for (int a = 0; a < 64; ++a) {
                double t = 0.0;
                for (int b = 0; b < 64; ++b) t += e.cherry_P[k * 4096 + a * 64 + b] * lm[b] * rm[b];
                out[a] = f32 ? (double)(float)t : t;
            }
        } else {
            if (m.tc5_leaf_order[leaf_pos++] != (int)s) host::die("leaf consumed out of ring order");
            leaf_msg(e, (int)s, ids[s], out);
            if (f32) for (int a = 0; a < 64; ++a) out[a] = (double)(float)out[a];
        }
    }
    void gemm(int edge, const double *al, double *msg) {
        const double *P = e.P.data() + (size_t)edge * 4096;
        if (!f32) {
            for (int a = 0; a < 64; ++a) { double s = 0.0; for (int b = 0; b < 64; ++b) s += P[a * 64 + b] * al[b]; msg[a] = s; }
            return;
        }
        // split TF32: A = hi + lo (exact), B = hi(P) + lo(P) as written by to_tc5_tile; three products, FP32 accumulate
        float mx = 0.f;
        for (int b = 0; b < 64; ++b) mx = std::fmax(mx, (float)al[b]);
        int ex = 0;
        if (mx > 0.f) std::frexp(mx, &ex);
        const float sc = std::ldexp(1.0f, 1 - ex);
        for (int a = 0; a < 64; ++a) {
            float d0 = 0.f, d1 = 0.f;
            for (int b = 0; b < 64; ++b) {
                const float v = (float)al[b] * sc, ah = tf32_hi(v), alo = v - ah;
                const float ph = tf32_rna((float)P[a * 64 + b]), pl = tf32_rna((float)(P[a * 64 + b] - (double)ph));
                d0 += ah * ph; d0 += alo * ph; d1 += ah * pl;
            }
            msg[a] = (double)(d0 + d1) / (double)sc;
        }
    }
    // f32 mode: the running partial is renormalised by an exact power of two after every step (the exponent is carried as an
    // integer, as in the kernel), so FP32's range never matters and only its 24-bit rounding shows
    int renorm(double *R) {
        if (!f32) return 0;
        double mx = 0.0;
        for (int a = 0; a < 64; ++a) mx = std::fmax(mx, R[a]);
        int ex = 0;
        if (mx > 0.0) std::frexp(mx, &ex);
        for (int a = 0; a < 64; ++a) R[a] = (double)(float)std::ldexp(R[a], -ex);
        return ex;
    }
    double run() {
        double R[64], L[64], t[64];
        long E = 0;
        std::vector<std::vector<double>> stack;
        std::vector<long> estack;
        source(m.tc5_start & 0xffu, R);
        source((m.tc5_start >> 8) & 0xffu, L);
        for (int a = 0; a < 64; ++a) R[a] *= L[a];
        E += renorm(R);
        for (size_t s = 0; s < m.tc5_steps.size(); ++s) {
            const uint32_t w = m.tc5_steps[s], post = (w >> 16) & 3u;
            gemm(m.tc5_edges[s], R, t);
            if (post == T5_MUL) { source(w & 0xffu, L); for (int a = 0; a < 64; ++a) R[a] = t[a] * L[a]; }
            else if (post == T5_PUSH_START) {
                stack.emplace_back(t, t + 64);
                estack.push_back(E);
                E = 0;
                if ((int)stack.size() > m.tc5_max_stack) host::die("stack deeper than tc5_max_stack");
                source(w & 0xffu, R); source((w >> 8) & 0xffu, L);
                for (int a = 0; a < 64; ++a) R[a] *= L[a];
            } else if (post == T5_POP_MUL) {
                if (stack.empty()) host::die("pop from an empty stack");
                for (int a = 0; a < 64; ++a) R[a] = t[a] * stack.back()[a];
                E += estack.back();
                stack.pop_back(); estack.pop_back();
            } else host::die("step without a post-op");
            if (((w & T5_END) != 0) != (s + 1 == m.tc5_steps.size())) host::die("END flag");
            E += renorm(R);
        }
        if (!stack.empty() || leaf_pos != m.tc5_leaf_order.size() || cherry_pos != m.tc5_cherries.size() || src_pos != m.tc5_srcs.size()) host::die("program left-overs");
        double z = 0.0;
        for (int a = 0; a < 64; ++a) z += e.pi[a] * R[a];
        return std::log(z) + (double)E * 0.6931471805599453;
    }
};

int main(int argc, char **argv) {
    if (argc < 2) host::die("usage: tc5_program_check <model> [species] [n_windows] [gap_fraction]");
    const std::string species = argc > 2 ? argv[2] : "";
    const int nwin = argc > 3 ? atoi(argv[3]) : 200;
    const double gap = argc > 4 ? atof(argv[4]) : 0.3;
    host::Model hm;
    host::load_model(hm, argv[1], species, "");
    ModelHost m;
    const double *S[2] = {hm.c.S.data(), hm.nc.S.data()}, *f[2] = {hm.c.f.data(), hm.nc.f.data()};
    const std::string err = prepare_model(m, hm.tree.nl, hm.tree.child1.data(), hm.tree.child2.data(), hm.tree.bl.data(), hm.tree.bl64.data(), S, f);
    if (!err.empty()) host::die("%s", err.c_str());
    {   // the BLS program with tabulated subtrees (prepare_model has already compared it with the node-by-node program on 2000 masks)
        int n_tab = 0;
        for (const BlsInner &e : m.bls_short) n_tab += (e.flags & 4) != 0;
        printf("BLS program: %zu entries (%d tables, %zu doubles) for %zu inner nodes, stack depth %d (was %d)\n", m.bls_short.size(), n_tab,
               m.bls_tables.size(), m.bls_inner.size(), m.bls_short_depth, m.bls_depth);
    }
    std::mt19937 rng(12345);
    std::uniform_real_distribution<double> U(0, 1);
    double worst = 0.0, worst32 = 0.0;
    std::vector<uint8_t> ids(m.nl);
    for (int w = 0; w < nwin; ++w) {
        const int base = (int)(rng() % 64);
        for (int s = 0; s < m.nl; ++s) ids[s] = U(rng) < gap ? 64 : (U(rng) < 0.7 ? base : (int)(rng() % 64));
        double lz[2], lz5[2], lz32[2];
        for (int k = 0; k < 2; ++k) {
            double al[64];
            alpha_ref(m, m.ecm[k], ids.data(), m.n - 1, al);
            double z = 0.0;
            for (int a = 0; a < 64; ++a) z += m.ecm[k].pi[a] * al[a];
            lz[k] = std::log(z);
            lz5[k] = Interp(m, m.ecm[k], ids.data(), false).run();
            lz32[k] = Interp(m, m.ecm[k], ids.data(), true).run();
            worst = std::fmax(worst, std::fabs(lz5[k] - lz[k]));
        }
        const double db = 10.0 / std::log(10.0);
        worst32 = std::fmax(worst32, std::fabs(db * ((lz32[0] - lz32[1]) - (lz[0] - lz[1]))));
    }
    int npush = 0;
    for (uint32_t w : m.tc5_steps) npush += ((w >> 16) & 3u) == T5_PUSH_START;
    printf("%s%s%s: nl %d, gemm steps %zu (of %d inner edges), direct leaves %zu, cherries %zu, pushes %d, stack depth %d | "
           "FP64 program vs recursion: max |d log z| = %.3e | FP32-class emulation: max |d| = %.3e decibans\n",
           argv[1], species.empty() ? "" : " --species ", species.c_str(), m.nl, m.tc5_steps.size(), m.nl - 2, m.tc5_leaf_order.size(),
           m.tc5_cherries.size(), npush, m.tc5_max_stack, worst, worst32);
    return worst < 1e-9 ? 0 : 2;
}
