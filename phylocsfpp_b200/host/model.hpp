// model.hpp — model loading: built-in parameter sets, external P.nh / P_coding.ECM / P_noncoding.ECM, --species,
// --mapping.  Mirrors load_model (reference src/models.hpp:1757-1856), empirical_codon_model::open
// (src/ecm.hpp:21-70), sequence_name_mapping (models.hpp:1468-1706) and update_sequence_name_mapping (:1709-1740).
// The built-in models ship as files under data/models/ in the reference's own external model format.
#pragma once

#include <unistd.h>

#include <fstream>
#include <map>
#include <sstream>
#include <unordered_map>

#include "../../include/phylocsf_b200.h"
#include "newick.hpp"

namespace host {

inline std::string data_dir() {
    if (const char *e = getenv("PHYLOCSF_B200_DATA")) return e;
    char buf[4096];
    const ssize_t n = readlink("/proc/self/exe", buf, sizeof buf - 1);
    if (n <= 0) return "data";
    buf[n] = 0;
    std::string p(buf);
    p = p.substr(0, p.find_last_of('/'));      // .../bin
    p = p.substr(0, p.find_last_of('/'));      // package root
    return p + "/data";
}

struct Ecm { std::vector<double> S, f; };   // S[64*64] symmetric, zero diagonal; f[64]

inline bool load_ecm(const std::string &path, Ecm &e) {
    std::ifstream fh(path);
    if (!fh) return false;
    e.S.assign(64 * 64, 0.0); e.f.assign(64, 0.0);
    std::string line;
    for (int line_id = 1; std::getline(fh, line); ++line_id) {
        if (line_id <= 63 || line_id == 65) {
            std::istringstream is(line);
            std::vector<double> vals;
            std::string tok;
            while (is >> tok) vals.push_back(std::stod(tok));
            if (line_id <= 63) {
                if ((int)vals.size() != line_id) die("%s: line %d has %zu entries, expected %d", path.c_str(), line_id, vals.size(), line_id);
                for (int j = 0; j < line_id; ++j) e.S[j * 64 + line_id] = e.S[line_id * 64 + j] = vals[j];
            } else {
                if (vals.size() != 64) die("%s: line 65 has %zu codon frequencies, expected 64", path.c_str(), vals.size());
                e.f = vals;
            }
        }
    }
    return true;
}

struct Model {
    std::string name;
    Ecm c, nc;
    std::unique_ptr<Node> root;
    FlatTree tree;
    std::unordered_map<std::string, uint16_t> seqid_to_phyloid;
    std::map<std::string, std::vector<std::string>> aliases;      // common name -> assembly names
    int nl() const { return tree.nl; }
};

inline std::vector<std::string> builtin_models() {
    std::vector<std::string> out;
    std::ifstream fh(data_dir() + "/models/INDEX");
    std::string ln;
    while (std::getline(fh, ln)) if (!ln.empty()) out.push_back(ln);
    return out;
}

inline void load_aliases(Model &m) {
    std::ifstream fh(data_dir() + "/species_aliases.tsv");
    std::string ln;
    while (std::getline(fh, ln)) {
        if (ln.empty()) continue;
        const size_t tab = ln.find('\t');
        const std::string common = ln.substr(0, tab);
        std::vector<std::string> &v = m.aliases[common];
        if (tab != std::string::npos)
            for (const std::string &a : split(ln.substr(tab + 1), ',')) if (!a.empty()) v.push_back(a);
    }
}

inline void update_sequence_name_mapping(Model &m, const std::string &path) {      // models.hpp:1709-1740
    std::ifstream fh(path);
    if (!fh) die("Could not open mapping file '%s'", path.c_str());
    std::string ln;
    while (std::getline(fh, ln)) {
        std::istringstream is(ln);
        std::string common, sci;
        if (!(is >> common >> sci)) continue;
        std::vector<std::string> &v = m.aliases[common];
        bool have = false;
        for (const std::string &x : v) have |= (x == sci);
        if (!have) v.push_back(sci);
    }
}

inline void load_model(Model &m, const std::string &name_or_path, const std::string &selected_species, const std::string &mapping_file) {
    load_aliases(m);
    if (!mapping_file.empty()) update_sequence_name_mapping(m, mapping_file);
    std::string prefix = name_or_path;
    const std::vector<std::string> builtin = builtin_models();
    for (const std::string &b : builtin) if (b == name_or_path) prefix = data_dir() + "/models/" + name_or_path;
    m.name = name_or_path;
    std::ifstream nh(prefix + ".nh");
    if (!load_ecm(prefix + "_coding.ECM", m.c) || !load_ecm(prefix + "_noncoding.ECM", m.nc) || !nh) {
        std::string all;
        for (const std::string &b : builtin) all += (all.empty() ? "" : ", ") + b;
        die("Could not open model files '%s{_coding.ECM,_noncoding.ECM,.nh}'. Pass the prefix to the model files without any file "
            "endings, or one of: %s", prefix.c_str(), all.c_str());
    }
    std::stringstream ss;
    ss << nh.rdbuf();
    m.root = newick_parse(ss.str());
    if (!selected_species.empty()) {                                                  // models.hpp:1791-1837
        std::vector<Node *> lv;
        newick_leaves(m.root.get(), lv);
        std::set<std::string> labels, selected;
        for (Node *l : lv) labels.insert(l->label);
        for (std::string s : split(selected_species, ',')) {
            s = lower(s);
            if (labels.count(s)) { selected.insert(s); continue; }
            bool found = false;
            for (const auto &kv : m.aliases)
                for (const std::string &alt : kv.second) if (alt == s) { found = true; selected.insert(kv.first); }
            if (!found) selected.insert(s);
        }
        std::string missing;
        for (const std::string &s : selected) if (!labels.count(s)) missing += (missing.empty() ? "" : ", ") + s;
        if (!missing.empty()) die("The following selected species are missing in the phylogenetic tree: %s", missing.c_str());
        newick_reduce(m.root.get(), selected);
    }
    m.tree = newick_flatten(m.root.get());
    for (int i = 0; i < m.tree.n; ++i) {
        const std::string &label = m.tree.labels[i];
        if (label.empty()) continue;
        m.seqid_to_phyloid.emplace(label, (uint16_t)i);
        auto it = m.aliases.find(label);
        if (it != m.aliases.end())
            for (const std::string &alt : it->second) m.seqid_to_phyloid.emplace(lower(alt), (uint16_t)i);
    }
}

inline pcsf_model *create_device_model(const Model &m, int device) {
    pcsf_model_desc d{(int32_t)m.tree.nl, m.tree.child1.data(), m.tree.child2.data(), m.tree.bl.data(), m.tree.bl64.data(),
                      m.c.S.data(), m.c.f.data(), m.nc.S.data(), m.nc.f.data()};
    pcsf_model *h = nullptr;
    if (pcsf_model_create(&d, device, &h) != PCSF_OK) die("pcsf_model_create (device %d): %s", device, pcsf_last_error());
    return h;
}

}  // namespace host
