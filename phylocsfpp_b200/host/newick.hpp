// newick.hpp — species tree: parsing, --species reduction, flattening.
//
// Semantics of the reference's src/newick.hpp: labels are lower-cased and branch lengths parsed as doubles
// (:31-91); strictly binary trees, root length 0 (:94-98); flatten (:218-229): leaves get ids 0..nl-1 in
// left-to-right DFS order, inner nodes nl..n-1 in post-order, root = n-1, flattened lengths stored as float
// (newick_elem, :20-29); reduce (:286-363): single-child chains are merged adding lengths in double.
#pragma once

#include <memory>
#include <set>
#include <string>
#include <vector>

#include "util.hpp"

namespace host {

struct Node {
    std::string label;
    double branch_length = 0.0;
    std::unique_ptr<Node> left, right;
    Node *parent = nullptr;
    int id = -1;
    bool leaf() const { return !left; }
};

struct NewickParser {
    std::string s;
    size_t i = 0;
    explicit NewickParser(const std::string &text) {
        for (char c : text) if (!isspace((unsigned char)c)) s.push_back(c);
    }
    char peek() const { return i < s.size() ? s[i] : 0; }
    double number() {
        size_t j = i;
        while (j < s.size() && (isdigit((unsigned char)s[j]) || s[j] == '.')) ++j;
        if (j == i) die("newick: number expected at offset %zu", i);
        const double v = std::stod(s.substr(i, j - i));
        i = j;
        return v;
    }
    std::unique_ptr<Node> subtree(Node *parent) {
        std::unique_ptr<Node> n(new Node);
        n->parent = parent;
        if (peek() == '(') {
            ++i;
            n->left = subtree(n.get());
            if (peek() != ',') die("newick: ',' expected at offset %zu", i);
            ++i;
            n->right = subtree(n.get());
            if (peek() != ')') die("newick: ')' expected at offset %zu (only binary trees are supported)", i);
            ++i;
            if (peek() == ':') { ++i; n->branch_length = number(); }
        } else {
            size_t j = i;
            while (j < s.size() && !strchr("(),:", s[j])) ++j;
            n->label = lower(s.substr(i, j - i));
            i = j;
            if (peek() != ':') die("newick: leaf without branch length at offset %zu", i);
            ++i;
            n->branch_length = number();
        }
        return n;
    }
};

inline std::unique_ptr<Node> newick_parse(const std::string &text) {
    NewickParser p(text);
    std::unique_ptr<Node> root = p.subtree(nullptr);
    if (root->branch_length != 0.0) die("newick: the root must not have a branch length");
    return root;
}

inline void newick_leaves(Node *n, std::vector<Node *> &out) {
    if (n->leaf()) { out.push_back(n); return; }
    newick_leaves(n->left.get(), out);
    newick_leaves(n->right.get(), out);
}

inline int newick_overlap(const Node *n, const std::set<std::string> &subset) {   // newick.hpp:264-270
    if (n->leaf()) return subset.count(n->label) ? 1 : 0;
    return newick_overlap(n->left.get(), subset) + newick_overlap(n->right.get(), subset);
}

inline void newick_reduce(Node *n, const std::set<std::string> &subset) {          // newick.hpp:286-363
    while (true) {
        if (n->leaf()) return;
        const int ol = newick_overlap(n->left.get(), subset), orr = newick_overlap(n->right.get(), subset);
        if (ol == 0 || orr == 0) {
            std::unique_ptr<Node> keep = std::move(ol == 0 ? n->right : n->left);
            n->left = std::move(keep->left);
            n->right = std::move(keep->right);
            if (n->left) { n->left->parent = n; n->right->parent = n; }
            else n->label = keep->label;
            if (n->parent) n->branch_length += keep->branch_length;
            continue;
        }
        newick_reduce(n->left.get(), subset);
        newick_reduce(n->right.get(), subset);
        return;
    }
}

struct FlatTree {
    int nl = 0, n = 0;
    std::vector<int16_t> child1, child2;
    std::vector<float> bl;
    std::vector<double> bl64;
    std::vector<std::string> labels;
};

inline FlatTree newick_flatten(Node *root) {
    std::vector<Node *> lv;
    newick_leaves(root, lv);
    FlatTree t;
    t.nl = (int)lv.size();
    t.n = 2 * t.nl - 1;
    for (int i = 0; i < t.nl; ++i) lv[i]->id = i;
    int counter = t.nl;
    struct Rec {
        static void annotate(Node *n, int &c) {
            if (n->leaf()) return;
            annotate(n->left.get(), c);
            annotate(n->right.get(), c);
            n->id = c++;
        }
        static void fill(Node *n, FlatTree &t) {
            t.bl[n->id] = (float)n->branch_length;
            t.bl64[n->id] = n->branch_length;
            t.labels[n->id] = n->label;
            if (!n->leaf()) {
                t.child1[n->id] = (int16_t)n->left->id;
                t.child2[n->id] = (int16_t)n->right->id;
                fill(n->left.get(), t);
                fill(n->right.get(), t);
            }
        }
    };
    Rec::annotate(root, counter);
    t.child1.assign(t.n, -1); t.child2.assign(t.n, -1);
    t.bl.assign(t.n, 0.f); t.bl64.assign(t.n, 0.0); t.labels.assign(t.n, "");
    Rec::fill(root, t);
    return t;
}

}  // namespace host
