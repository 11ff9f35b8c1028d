// Tree::reroot on the library (reference src/reroot.cpp:4-261): the topology transform and the re-inference of every
// block and nucleotide column with the new root's own state forced at the root.
//
// Restated from the reference:
//   * Tree::transform / transformHelper       src/panman.cpp:5831-5906: the tip's parent chain is turned upside down. The tip
//     leaves its parent; a NEW root "node_<k+1>" gets the children [tip, old parent]; every node on the chain loses the
//     child the chain came through and receives its own old parent as its LAST child; the old root stays if it still has
//     more than one child, otherwise it is deleted and its remaining child takes its place. A tip whose parent is the root
//     (or that is the root) changes nothing.
//   * block columns                            src/reroot.cpp:54-122: blockFitch*New with states 1 / 2 / 4 of EVERY leaf and
//     defaultState = the new root's state
//   * nucleotide columns                       src/reroot.cpp:134-224: every leaf takes part with its character ('-' and the
//     end marker 'x' are gaps); nucFitchBackwardPass(root, states, code, code) forces the root to the new root's character;
//     the assign pass starts from '-' (gap columns) or the consensus character (main columns)
//   * sort + greedy <= 6 run-merge             src/reroot.cpp:226-261 (the PanGraph form: block, pos + 1 / gapPos + 1, type)
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/panman_b200_host.h"
#include "host_tree.hpp"

namespace pmh {

// Returns "" or an error message. Node ids of `out` are a pre-order walk of the new tree (children in their new order);
// names and leaf rows travel with the nodes, so the rows of a code matrix keep their meaning.
std::string reroot_tree(const HostTree& in, int32_t tip, HostTree* out) {
    const int32_t n = in.n_nodes();
    if (tip < 0 || tip >= n) return "reroot: no such node";
    if (in.child_off[tip + 1] != in.child_off[tip]) return "reroot: node " + in.names[tip] + " is not a tip";  // src/reroot.cpp:10-13
    const int32_t par = in.parent[tip];
    if (par < 0 || par == in.root) {  // already the root / the root's child: the reference changes nothing (:5868-5877)
        *out = in;
        return "";
    }
    std::vector<std::vector<int32_t>> kids(size_t(n) + 1);
    for (int32_t v = 0; v < n; v++) kids[v].assign(in.child_idx.begin() + in.child_off[v], in.child_idx.begin() + in.child_off[v + 1]);
    auto drop = [&](int32_t from, int32_t child) {
        auto& k = kids[from];
        k.erase(std::find(k.begin(), k.end(), child));
    };
    // the chain par = p0, p1, ..., pk = old root
    std::vector<int32_t> chain;
    for (int32_t v = par; v >= 0; v = in.parent[v]) chain.push_back(v);
    drop(par, tip);
    for (size_t i = 0; i + 1 < chain.size(); i++) drop(chain[i + 1], chain[i]);
    const int32_t old_root = chain.back();
    bool root_dead = false;
    int32_t top = old_root;
    if (kids[old_root].size() <= 1) {  // (:5832-5842) a root left with a single child is replaced by that child
        if (kids[old_root].empty()) return "reroot: the old root has no other child";
        top = kids[old_root][0];
        kids[old_root].clear();
        root_dead = true;
    }
    for (size_t i = chain.size() - 1; i-- > 0;) kids[chain[i]].push_back(i + 2 == chain.size() ? top : chain[i + 1]);
    const int32_t new_root = n;
    kids[new_root] = {tip, par};
    size_t n_internal = 0;
    for (int32_t v = 0; v < n; v++) n_internal += in.child_off[v + 1] > in.child_off[v];
    // pre-order renumbering
    HostTree t;
    std::vector<int32_t> new_id(size_t(n) + 1, -1), stack{new_root}, order;
    while (!stack.empty()) {
        const int32_t v = stack.back();
        stack.pop_back();
        new_id[v] = int32_t(order.size());
        order.push_back(v);
        for (size_t k = kids[v].size(); k-- > 0;) stack.push_back(kids[v][k]);
    }
    if (int32_t(order.size()) != n + 1 - (root_dead ? 1 : 0)) return "reroot: internal error (nodes lost)";
    const int32_t m = int32_t(order.size());
    t.names.resize(m);
    t.parent.assign(m, -1);
    t.leaf_row.assign(m, -1);
    t.child_off.assign(m + 1, 0);
    for (int32_t i = 0; i < m; i++) {
        const int32_t v = order[i];
        t.names[i] = v == new_root ? "node_" + std::to_string(n_internal + 1) : in.names[v];  // newInternalNodeId, src/panman.hpp:793
        t.leaf_row[i] = v == new_root ? -1 : in.leaf_row[v];
        t.child_off[i + 1] = t.child_off[i] + int32_t(kids[v].size());
    }
    t.child_idx.resize(t.child_off[m]);
    for (int32_t i = 0; i < m; i++) {
        const int32_t v = order[i];
        for (size_t k = 0; k < kids[v].size(); k++) {
            const int32_t c = new_id[kids[v][k]];
            t.child_idx[t.child_off[i] + int32_t(k)] = c;
            t.parent[c] = i;
        }
    }
    t.root = 0;
    t.n_leaves = in.n_leaves;
    *out = std::move(t);
    return "";
}

int32_t find_node(const HostTree& t, const std::string& name) {
    for (int32_t v = 0; v < t.n_nodes(); v++)
        if (t.names[v] == name) return v;
    return -1;
}

}  // namespace pmh

namespace {
void set_err_(char* err, size_t n, const std::string& m) {
    if (err && n) {
        std::strncpy(err, m.c_str(), n - 1);
        err[n - 1] = 0;
    }
}
}  // namespace

extern "C" pmh_tree* pmh_tree_reroot(const pmh_tree* t, const char* leaf_name, char* err, size_t err_len) {
    if (!t || !leaf_name) {
        set_err_(err, err_len, "null argument");
        return nullptr;
    }
    try {
        const int32_t tip = pmh::find_node(t->t, leaf_name);
        if (tip < 0) {
            set_err_(err, err_len, std::string("Sequence with name ") + leaf_name + " not found!");  // src/reroot.cpp:5-8
            return nullptr;
        }
        pmh_tree* out = new pmh_tree();
        std::string e = pmh::reroot_tree(t->t, tip, &out->t);
        if (!e.empty()) {
            delete out;
            set_err_(err, err_len, e);
            return nullptr;
        }
        return out;
    } catch (const std::exception& ex) {
        set_err_(err, err_len, std::string("reroot: ") + ex.what());
        return nullptr;
    }
}
