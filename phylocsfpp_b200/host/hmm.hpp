// PhyloCSF-HMM: smoothing of the raw per-codon scores into posterior log-odds tracks and candidate coding regions.
//
// Host-side restatement of
//   * estimate_hmm_params_for_genome + the exponential-mixture EM over inter-exon gaps with its 1-D Nelder-Mead
//     (reference src/estimate_hmm_parameter.hpp:44-340),
//   * get_coding_hmm (src/create_tracks.hpp:162-200),
//   * hmm::state_posterior_probabilities / get_best_path_by_viterbi (src/create_tracks.hpp:29-160),
//   * compute_log_odds, process_scores (src/create_tracks.hpp:226-314),
//   * wig_reader::get_next_scores (src/wig_file_reader.hpp:95-139): the HMM input is the TEXT of the raw tracks (the
//     "%.3f"-rounded, power-thresholded values), consecutive fixedStep runs that continue each other are joined.
// The operation order of every floating-point expression follows the reference so that the "%.3f" text is identical
// (tests/test_host_cli.py compares against files written by the reference itself).  Memory differs: one forward table,
// the backward vector is rolled, Viterbi back-pointers are packed four to a byte.
#pragma once

#include <algorithm>
#include <cfloat>
#include <cinttypes>
#include <cmath>
#include <list>
#include <map>
#include <random>

#include "util.hpp"

namespace host {

struct HmmParams {
    double coding_prior, coding_codons;
    double nc_weight[3], nc_codons[3];
};

struct Hmm {
    double init[4];
    double trans[4][4];
};

namespace hmm_detail {

// negative weighted log-likelihood of an exponential with mean 10^x (estimate_hmm_parameter.hpp:40-48)
inline double neg_loglik(const std::vector<uint32_t> &pts, const double *w, double x) {
    double f = 0.0;
    for (size_t p = 0; p < pts.size(); ++p) {
        const double tau = pow(10.0, x);
        const double ll = (-static_cast<double>(pts[p]) / tau - log(tau));
        f -= w[p] * ll;
    }
    return f;
}

// 1-D Nelder-Mead over two vertices, at most 30 steps (estimate_hmm_parameter.hpp:50-139)
inline double minimize_1d(const std::vector<uint32_t> &pts, const double *w, double guess, double xscale, double relxtol) {
    typedef std::pair<double, double> V;          // (x, f(x))
    const auto by_f = [](const V &a, const V &b) { return a.second < b.second; };
    std::vector<V> sx = {V(guess, neg_loglik(pts, w, guess)), V(guess + xscale, neg_loglik(pts, w, guess + xscale))};
    std::sort(sx.begin(), sx.end(), by_f);
    const double xtol = relxtol * xscale;
    const uint32_t max_steps = 30;
    bool grew_or_shrank = true;
    for (size_t it = 0; it <= max_steps; ++it) {
        if (!grew_or_shrank) {
            double lo = DBL_MAX, hi = DBL_MIN;    // (sic) DBL_MIN, as in the reference
            for (const V &v : sx) { lo = std::min(lo, v.first); hi = std::max(hi, v.first); }
            if (hi - lo < xtol) return sx[0].first;
        }
        const size_t n = sx.size() - 1;
        double centroid = 0.0;
        for (size_t i = 0; i < n; ++i) centroid += sx[i].first;
        centroid /= n;
        const double xr = centroid + (centroid - sx[n].first);
        const double fr = neg_loglik(pts, w, xr);
        grew_or_shrank = false;
        if (sx[0].second <= fr && fr < sx[n - 1].second) {
            sx[n] = V(xr, fr);
        } else if (fr < sx[0].second) {
            const double xe = centroid + 2 * (centroid - sx[n].first);
            const double fe = neg_loglik(pts, w, xe);
            if (fe < fr) { sx[n] = V(xe, fe); grew_or_shrank = true; }
            else sx[n] = V(xr, fr);
        } else {
            const double xc = centroid - .5 * (centroid - sx[n].first);
            const double fc = neg_loglik(pts, w, xc);
            if (fc < sx[n].second) {
                sx[n] = V(xc, fc);
            } else {
                grew_or_shrank = true;
                for (size_t i = 1; i < n + 1; ++i) {
                    const double nx = sx[0].first + 0.5 * (sx[i].first - sx[0].first);
                    sx[i] = V(nx, neg_loglik(pts, w, nx));
                }
            }
        }
        std::sort(sx.begin(), sx.end(), by_f);
    }
    die("nelder_mead did not converge in %u steps", max_steps);
}

}  // namespace hmm_detail

// estimate_hmm_parameter.hpp:243-340.  genome_length is a uint32_t in the reference.
inline HmmParams estimate_hmm_params(const std::string &exons_path, uint32_t genome_length) {
    FILE *fh = fopen(exons_path.c_str(), "r");
    if (!fh) die("could not open %s", exons_path.c_str());
    typedef std::pair<uint32_t, uint32_t> Range;          // (start, end); pair's ordering = the reference's comparator
    std::map<std::string, std::list<Range>> by_key;     // chrom:strand:phase
    char line[1024];
    while (fgets(line, sizeof line, fh)) {
        const char *delim = " \t";
        char *tok = strtok(line, delim);
        if (!tok) continue;
        std::string key(tok);
        for (int k = 0; k < 2; ++k) { tok = strtok(nullptr, delim); if (!tok) die("malformed line in %s", exons_path.c_str()); key += ":"; key += tok; }
        tok = strtok(nullptr, delim); if (!tok) die("malformed line in %s", exons_path.c_str());
        const uint32_t start = (uint32_t)atoi(tok);
        tok = strtok(nullptr, delim); if (!tok) die("malformed line in %s", exons_path.c_str());
        const uint32_t end = (uint32_t)atoi(tok);
        by_key[key].emplace_back(start, end);
    }
    fclose(fh);
    uint64_t num_exons = 0;
    size_t coding_nt = 0;
    std::vector<uint32_t> gaps;
    for (auto &kv : by_key) {
        std::list<Range> &ex = kv.second;
        ex.sort();
        // of two overlapping exons the shorter one is dropped (:286-310)
        size_t i = 0;
        auto tail = ex.begin(), head = std::next(ex.begin());
        while (i < ex.size() - 1) {
            if (head->first <= tail->second) {
                if (tail->second - tail->first >= head->second - head->first) head = ex.erase(head);
                else { tail = ex.erase(tail); head = std::next(tail); }
            } else { ++head; ++tail; ++i; }
        }
        // (sic) the iterator advances twice per pass: every other inter-exon gap is sampled (:311-322)
        for (auto it = ex.begin(); it != ex.end(); ++it) {
            const uint32_t end1 = it->second;
            ++it;
            if (it == ex.end()) break;
            if (it->first > end1 + 1) gaps.push_back(it->first - end1 - 1);
        }
        num_exons += ex.size();
        for (const Range &r : ex) coding_nt += r.second - r.first + 1;
    }
    // estimate_gap_mixture_model (:210-241): at most 20 000 gaps, libstdc++'s default engine seeded 0
    if (gaps.size() > 20000) {
        std::default_random_engine rng(0);
        std::shuffle(gaps.begin(), gaps.end(), rng);
        gaps.resize(20000);
    }
    const uint32_t guess_len[3] = {3000, 80000, 100};
    double prior[3] = {30, 10, 1}, x[3];
    const double psum = prior[0] + prior[1] + prior[2];
    for (int j = 0; j < 3; ++j) { prior[j] = prior[j] / psum; x[j] = log10(guess_len[j]); }
    // infer_mixture (:157-207): 20 EM rounds, the M-step of each mean is a 1-D Nelder-Mead
    const size_t np = gaps.size();
    std::vector<double> resp[3] = {std::vector<double>(np, 0.0), std::vector<double>(np, 0.0), std::vector<double>(np, 0.0)};
    for (int round = 0; round < 20; ++round) {
        for (size_t i = 0; i < np; ++i) {
            double lik[3];
            for (int j = 0; j < 3; ++j) {
                const double tau = pow(10.0, x[j]);
                const double ld = (-static_cast<double>(gaps[i]) / tau - log(tau));
                lik[j] = prior[j] * exp(ld);
            }
            const double total = lik[0] + lik[1] + lik[2];
            for (int j = 0; j < 3; ++j) resp[j][i] = (total != 0.0) ? lik[j] / total : 1.0 / 3;
        }
        for (int j = 0; j < 3; ++j) {
            double s = 0.0;
            for (size_t i = 0; i < np; ++i) s += resp[j][i];
            prior[j] = s / static_cast<double>(np);
        }
        for (int j = 0; j < 3; ++j) {
            if (x[j] == 0) continue;
            x[j] = hmm_detail::minimize_1d(gaps, resp[j].data(), x[j], 0.1, 0.001);
        }
    }
    HmmParams p;
    p.coding_prior = static_cast<double>(coding_nt) / static_cast<double>(genome_length) / 6.0;
    p.coding_codons = static_cast<double>(coding_nt) / static_cast<double>(num_exons) / 3.0;
    for (int j = 0; j < 3; ++j) { p.nc_weight[j] = prior[j]; p.nc_codons[j] = pow(10, x[j]) / 3; }
    return p;
}

// create_tracks.hpp:162-200: state 0 = coding, 1..3 = the three non-coding length classes
inline Hmm coding_hmm(const HmmParams &p) {
    Hmm h;
    double unnorm[3], c2nc[3], nc2c[3];
    for (int i = 0; i < 3; ++i) {
        unnorm[i] = p.nc_weight[i] * p.nc_codons[i];
        c2nc[i] = p.nc_weight[i] / p.coding_codons;
        nc2c[i] = 1.0 / p.nc_codons[i];
    }
    h.init[0] = p.coding_prior;
    const double usum = unnorm[0] + unnorm[1] + unnorm[2];
    for (int i = 0; i < 3; ++i) h.init[i + 1] = (1 - p.coding_prior) * unnorm[i] / usum;
    h.trans[0][0] = (1.0 - (c2nc[0] + c2nc[1] + c2nc[2]));
    for (int j = 0; j < 3; ++j) h.trans[0][j + 1] = c2nc[j];
    for (int i = 1; i < 4; ++i) {
        h.trans[i][0] = nc2c[i - 1];
        for (int j = 1; j < 4; ++j) h.trans[i][j] = (i == j) ? 1.0 - nc2c[i - 1] : 0.0;
    }
    return h;
}

inline double hmm_emit(int state, double score) { return state == 0 ? pow(10, (score / 10)) : 1; }

// compute_log_odds (create_tracks.hpp:226-235)
inline double hmm_log_odds(double prob) {
    const double MAX_LOG_ODDS = 15.0;
    if (prob < pow(10, -MAX_LOG_ODDS)) return -MAX_LOG_ODDS;
    if (prob > 1 - pow(10, -MAX_LOG_ODDS)) return MAX_LOG_ODDS;
    return log10(prob / (1 - prob));
}

// Posterior probability of the coding state at every codon of one contiguous run (scaled forward-backward,
// create_tracks.hpp:86-158).
inline void hmm_posterior_coding(const Hmm &h, const std::vector<double> &obs, std::vector<double> &post) {
    const size_t n = obs.size();
    post.resize(n);
    if (n == 0) return;
    std::vector<double> fwd(4 * n);
    for (int s = 0; s < 4; ++s) fwd[s] = h.init[s] * hmm_emit(s, obs[0]);
    for (size_t pos = 1; pos < n; ++pos) {
        const double *pf = &fwd[4 * (pos - 1)];
        double *cf = &fwd[4 * pos];
        const double e0 = hmm_emit(0, obs[pos]);
        double maxf = 0.0;
        for (int s = 0; s < 4; ++s) {
            double sum = 0.0;
            for (int q = 0; q < 4; ++q) sum += pf[q] * h.trans[q][s];
            cf[s] = sum * (s == 0 ? e0 : 1);
            maxf = std::max(maxf, cf[s]);
        }
        for (int s = 0; s < 4; ++s) cf[s] /= maxf;
    }
    double bw[4] = {1.0, 1.0, 1.0, 1.0}, nb[4];
    for (size_t pos = n; pos-- > 0;) {
        if (pos + 1 < n) {
            const double e0 = hmm_emit(0, obs[pos + 1]);
            double maxb = 0.0;
            for (int s = 0; s < 4; ++s) {
                double sum = 0.0;
                for (int q = 0; q < 4; ++q) sum += (h.trans[s][q] * (q == 0 ? e0 : 1) * bw[q]);
                nb[s] = sum;
                maxb = std::max(maxb, nb[s]);
            }
            for (int s = 0; s < 4; ++s) bw[s] = nb[s] / maxb;
        }
        const double *cf = &fwd[4 * pos];
        double total = 0.0;
        for (int s = 0; s < 4; ++s) total += (cf[s] * bw[s]);
        post[pos] = (cf[0] * bw[0]) / total;
    }
}

// Most probable state path (create_tracks.hpp:29-72; ties keep the lowest previous state, scores rescaled by the maximum).
inline void hmm_viterbi(const Hmm &h, const std::vector<double> &obs, std::vector<uint8_t> &path) {
    const size_t n = obs.size();
    path.resize(n);
    if (n == 0) return;
    std::vector<uint8_t> back(n);          // 4 x 2 bits per position
    double prev[4], cur[4];
    for (int s = 0; s < 4; ++s) prev[s] = h.init[s] * hmm_emit(s, obs[0]);
    for (size_t pos = 1; pos < n; ++pos) {
        double best_all = 0.0;
        uint8_t packed = 0;
        for (int s = 0; s < 4; ++s) {
            double best = 0.0;
            int arg = 0;
            for (int q = 0; q < 4; ++q) {
                const double v = prev[q] * h.trans[q][s];
                if (v > best) { arg = q; best = v; }
            }
            best *= hmm_emit(s, obs[pos]);
            packed |= (uint8_t)(arg << (2 * s));
            cur[s] = best;
            best_all = std::max(best_all, best);
        }
        back[pos] = packed;
        for (int s = 0; s < 4; ++s) prev[s] = cur[s] / best_all;
    }
    int st = -1;
    double mx = 0.0;
    for (int s = 0; s < 4; ++s) if (st < 0 || prev[s] > mx) { mx = prev[s]; st = s; }
    path[n - 1] = (uint8_t)st;
    for (size_t pos = n - 1; pos > 0; --pos) { st = (back[pos] >> (2 * st)) & 3; path[pos - 1] = (uint8_t)st; }
}

// One raw track (text) -> smoothed track text and/or region BED text (build_tracks.hpp:300-348, process_scores
// create_tracks.hpp:249-314).  The colour triple is always 0 in the reference (the value of computing_color_code is
// discarded) and so it is here.
inline void hmm_smooth_file(const Hmm &h, const std::string &raw_path, char strand, FILE *out_wig, FILE *out_bed) {
    FILE *fh = fopen(raw_path.c_str(), "r");
    if (!fh) die("Cannot open %s", raw_path.c_str());
    std::vector<double> scores, post;
    std::vector<uint8_t> path;
    std::string chr, text;
    uint64_t start = 0;
    const auto flush = [&] {
        if (scores.empty()) return;
        const uint32_t block = (uint32_t)start;          // positions pass through uint32_t (create_tracks.hpp:249-255)
        if (out_wig) {
            hmm_posterior_coding(h, scores, post);
            text.clear();
            char hdr[512];
            snprintf(hdr, sizeof hdr, "fixedStep chrom=%s start=%" PRId64 " step=3 span=3\n", chr.c_str(), (int64_t)start);
            text += hdr;
            for (double p : post) my_format(text, 3, (float)hmm_log_odds(p));
            fwrite(text.data(), 1, text.size(), out_wig);
        }
        if (out_bed) {
            hmm_viterbi(h, scores, path);
            uint32_t rs = 0;
            const auto emit = [&](uint32_t a, uint32_t b) {
                fprintf(out_bed, "%s\t%" PRIu32 "\t%" PRIu32 "\t%s:%" PRIu32 "-%" PRIu32 "\t0\t%c\t%" PRIu32 "\t%" PRIu32 "\t0,0,0\n", chr.c_str(), a, b,
                        chr.c_str(), a + 1, b, strand, a, b);
            };
            const size_t np = path.size();
            for (size_t i = 0; i + 1 < np; ++i) {
                const uint32_t i3 = (uint32_t)(3 * i);
                if (i == 0 && path[i] == 0) {
                    rs = block - 1;
                    if (path[i + 1] != 0) emit(rs, rs + 3);
                } else if (path[i + 1] == 0 && path[i] != 0) {
                    if (i != np - 2) rs = block + i3 + 2;
                    else emit(block + i3 + 5 - 3, block + i3 + 5);
                } else if (path[i + 1] != 0 && path[i] == 0) {
                    emit(rs, block + i3 + 2);
                } else if (i == np - 2 && path[i + 1] == 0 && path[i] == 0) {
                    emit(rs, block + i3 + 5);
                }
            }
        }
        scores.clear();
    };
    char buf[512];
    while (fgets(buf, sizeof buf, fh)) {
        if (buf[0] == 'f') {
            char c_chr[400];
            uint64_t s = 0;
            if (sscanf(buf + 16, "%399s %*6s%" SCNu64, c_chr, &s) != 2) die("malformed wig header in %s: %s", raw_path.c_str(), buf);
            if (!scores.empty() && (chr != c_chr || start + 3 * scores.size() != s)) flush();
            if (scores.empty()) { chr = c_chr; start = s; }
        } else if (buf[0] != '\n') {
            scores.push_back(strtod(buf, nullptr));
        }
    }
    flush();
    fclose(fh);
}

}  // namespace host
