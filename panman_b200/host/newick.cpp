#include <algorithm>

#include "host_tree.hpp"

namespace pmh {

bool HostTree::has_polytomy() const {  // reference src/panman.cpp:621-631
    for (int32_t v = 0; v < n_nodes(); v++)
        if (child_off[v + 1] - child_off[v] > 2) return true;
    return false;
}

// A piece that holds an odd number of apostrophes is glued to the following pieces until the count is even, so
// that commas inside quoted names do not split (reference src/panman.cpp:265-295).
void split_quote_aware(const std::string& s, char delim, std::vector<std::string>& words) {
    size_t start = 0, held = 0;
    bool holding = false;
    for (;;) {
        size_t end = s.find(delim, start);
        if (end == std::string::npos) break;
        size_t from = holding ? held : start;
        std::string sub = s.substr(from, end - from);
        bool odd = std::count(sub.begin(), sub.end(), '\'') % 2 == 1;
        if (!holding) {
            if (odd) { holding = true; held = start; }
            else words.push_back(sub);
        } else if (!odd) {
            holding = false;
            words.push_back(sub);
        }
        start = end + 1;
    }
    std::string last = s.substr(start);
    if (!last.empty()) words.push_back(last);
}

// Shape-only restatement (branch lengths do not reach the Fitch/Sankoff path):
//  * comma-split pieces; in each piece '(' opens an internal node, the characters before the first ':' or ')' that
//    are not parentheses form the leaf name, ')' closes a node (src/panman.cpp:332-385);
//  * internal nodes are named node_1, node_2, ... in order of their '(' (newInternalNodeId, src/panman.hpp:793-795);
//  * a node is appended to its parent's children when created, so children keep Newick order (src/panman.cpp:223-229);
//  * node ids here = creation order: the '(' of a piece first, then its leaf (src/panman.cpp:404-437).
std::string parse_newick(const std::string& newick, HostTree* out) {
    std::string s = newick;
    // stripString (src/panman.cpp:298-308): trailing blanks go (two characters per blank there), then leading blanks.
    while (!s.empty() && (s.back() == ' ' || s.back() == '\n' || s.back() == '\r')) s.pop_back();
    size_t lead = 0;
    while (lead < s.size() && s[lead] == ' ') lead++;
    s = s.substr(lead);

    std::vector<std::string> pieces;
    split_quote_aware(s, ',', pieces);
    HostTree t;
    std::vector<std::vector<int32_t>> kids;
    std::vector<int32_t> stack;
    size_t internal_counter = 0;
    for (const std::string& piece : pieces) {
        size_t n_open = 0, n_close = 0;
        bool stop = false, name_zone = false, has_apo = false;
        std::string leaf;
        for (char c : piece) {
            if (name_zone) {
                leaf += c;
                if (c == '\'') name_zone = false;
            } else if (c == '\'') {
                name_zone = has_apo = true;
                leaf += c;
            } else if (c == ':') {
                stop = true;
            } else if (c == '(') {
                n_open++;
            } else if (c == ')') {
                stop = true;
                n_close++;
            } else if (!stop) {
                leaf += c;
            }
        }
        if (has_apo && leaf.size() >= 2 && leaf.front() == '\'' && leaf.back() == '\'') leaf = leaf.substr(1, leaf.size() - 2);
        for (size_t j = 0; j < n_open; j++) {
            int32_t id = int32_t(t.names.size());
            t.names.push_back("node_" + std::to_string(++internal_counter));
            kids.emplace_back();
            t.parent.push_back(stack.empty() ? -1 : stack.back());
            if (!stack.empty()) kids[stack.back()].push_back(id);
            stack.push_back(id);
        }
        if (stack.empty()) return "incorrect Newick format: a leaf outside any parenthesis";
        int32_t id = int32_t(t.names.size());
        t.names.push_back(leaf);
        kids.emplace_back();
        t.parent.push_back(stack.back());
        kids[stack.back()].push_back(id);
        if (n_close > stack.size()) return "incorrect Newick format: unbalanced ')'";
        for (size_t j = 0; j < n_close; j++) stack.pop_back();
    }
    if (!stack.empty()) return "incorrect Newick format: unbalanced '('";  // src/panman.cpp:397-400
    if (t.names.empty()) return "empty tree";
    const int32_t n = t.n_nodes();
    t.child_off.assign(n + 1, 0);
    for (int32_t v = 0; v < n; v++) t.child_off[v + 1] = t.child_off[v] + int32_t(kids[v].size());
    t.child_idx.resize(t.child_off[n]);
    for (int32_t v = 0; v < n; v++) std::copy(kids[v].begin(), kids[v].end(), t.child_idx.begin() + t.child_off[v]);
    t.leaf_row.assign(n, -1);
    int32_t r = 0;
    for (int32_t v = 0; v < n; v++)
        if (kids[v].empty()) t.leaf_row[v] = r++;
    t.n_leaves = r;
    t.root = 0;
    *out = std::move(t);
    return "";
}

}  // namespace pmh
