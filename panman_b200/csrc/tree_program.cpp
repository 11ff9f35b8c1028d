#include "tree_program.h"

#include <algorithm>
#include <numeric>

namespace pmb {

std::string build_tree_program(int32_t n_nodes, int32_t root, const int32_t* child_off, const int32_t* child_idx,
                               const int32_t* leaf_row, int32_t chunk_nodes, int32_t inline_nodes, TreeProgram* out,
                               int32_t bwd_tail_chunks) {
    TreeProgram& P = *out;
    P = TreeProgram();
    if (n_nodes < 2) return "tree needs at least one internal node and one leaf";
    if (root < 0 || root >= n_nodes) return "root out of range";
    if (child_off[0] != 0) return "child_offsets[0] must be 0";
    if (chunk_nodes < 1) chunk_nodes = 1;
    if (n_nodes > int32_t(REF_IDX_MASK)) return "tree too large";
    const int32_t n_edges = child_off[n_nodes];
    if (n_edges != n_nodes - 1) return "child_index must hold exactly n_nodes - 1 entries (a tree)";

    // ---- validate: every non-root node is a child exactly once; leaves have rows; rows are a permutation ----
    std::vector<int32_t> parent(n_nodes, -1);
    std::vector<char> seen(n_nodes, 0);
    for (int32_t v = 0; v < n_nodes; v++) {
        if (child_off[v + 1] < child_off[v]) return "child_offsets not monotone";
        for (int32_t e = child_off[v]; e < child_off[v + 1]; e++) {
            int32_t c = child_idx[e];
            if (c < 0 || c >= n_nodes || c == root || seen[c]) return "child_index is not a tree";
            seen[c] = 1;
            parent[c] = v;
        }
    }
    int32_t n_rows = 0, n_internal = 0;
    for (int32_t v = 0; v < n_nodes; v++) {
        bool leaf = child_off[v] == child_off[v + 1];
        if (leaf) {
            if (leaf_row[v] < 0) return "leaf without a row";
            n_rows++;
        } else {
            if (leaf_row[v] != -1) return "internal node with a row";
            n_internal++;
        }
    }
    if (child_off[root] == child_off[root + 1]) return "root must be an internal node";
    {
        std::vector<char> row_seen(n_rows, 0);
        for (int32_t v = 0; v < n_nodes; v++)
            if (leaf_row[v] >= 0) {
                if (leaf_row[v] >= n_rows || row_seen[leaf_row[v]]) return "leaf_row must be a permutation of 0..n_leaves-1";
                row_seen[leaf_row[v]] = 1;
            }
    }

    // ---- pre-order from the root (also proves connectivity) ----
    std::vector<int32_t> pre;
    pre.reserve(n_nodes);
    {
        std::vector<int32_t> st{root};
        while (!st.empty()) {
            int32_t v = st.back();
            st.pop_back();
            pre.push_back(v);
            for (int32_t e = child_off[v + 1] - 1; e >= child_off[v]; e--) st.push_back(child_idx[e]);
        }
        if (int32_t(pre.size()) != n_nodes) return "tree is not connected";
    }

    // ---- internal subtree sizes, then children in processing order: internal heavy -> light ----
    std::vector<int32_t> isz(n_nodes, 0);
    for (int32_t i = n_nodes - 1; i >= 0; i--) {
        int32_t v = pre[i];
        if (child_off[v] == child_off[v + 1]) continue;
        isz[v] += 1;
        if (parent[v] >= 0) isz[parent[v]] += isz[v];
    }
    std::vector<int32_t> ichild_off(n_nodes + 1, 0), ichild;  // internal children only, processing order
    ichild.reserve(n_internal);
    for (int32_t v = 0; v < n_nodes; v++) {
        int32_t b = int32_t(ichild.size());
        for (int32_t e = child_off[v]; e < child_off[v + 1]; e++)
            if (isz[child_idx[e]] > 0) ichild.push_back(child_idx[e]);
        std::stable_sort(ichild.begin() + b, ichild.end(), [&](int32_t a, int32_t c) { return isz[a] > isz[c]; });
        ichild_off[v + 1] = int32_t(ichild.size());
    }

    // ---- cut into chunks: bottom subtrees (<= chunk_nodes) and heavy-path segments of the top tree ----
    if (inline_nodes < 0) inline_nodes = 0;
    if (inline_nodes >= chunk_nodes) inline_nodes = chunk_nodes - 1;
    std::vector<char> cut(n_nodes, 0), chain_cut(n_nodes, 0);
    std::vector<int32_t> run(n_nodes, 0);  // top nodes: ops of the node's chunk from the chunk's root down to this node
    for (int32_t i = 0; i < n_nodes; i++) {
        int32_t v = pre[i];
        if (isz[v] == 0) continue;
        const bool v_top = isz[v] > chunk_nodes;
        int32_t own = 1;  // this op plus the light subtrees evaluated inside the chunk
        if (v_top)
            for (int32_t e = ichild_off[v]; e < ichild_off[v + 1]; e++)
                if (isz[ichild[e]] <= inline_nodes) own += isz[ichild[e]];
        if (v == root) { cut[v] = 1; run[v] = own; continue; }
        int32_t p = parent[v];
        bool p_top = isz[p] > chunk_nodes;
        if (!p_top) continue;                         // inside a bottom subtree
        if (v_top) {
            cut[v] = (ichild[ichild_off[p]] != v);    // top node: stays iff heaviest child ...
            if (!cut[v] && run[p] >= chunk_nodes) cut[v] = chain_cut[v] = 1;  // ... and the path segment is not full yet
            run[v] = cut[v] ? own : run[p] + own;
        } else {
            cut[v] = isz[v] > inline_nodes;           // bottom subtree hanging off the top tree
        }
    }

    // ---- per-chunk op lists (post-order over uncut internal children), chunk levels ----
    struct TmpChunk {
        int32_t root_node;
        int32_t level;
        std::vector<int32_t> nodes;  // post-order
    };
    std::vector<TmpChunk> tmp;
    std::vector<int32_t> chunk_of(n_nodes, -1);
    // chunk roots in reverse pre-order => children chunks are created before their parents
    for (int32_t i = n_nodes - 1; i >= 0; i--) {
        int32_t r = pre[i];
        if (!cut[r]) continue;
        TmpChunk ch;
        ch.root_node = r;
        ch.level = 0;
        std::vector<std::pair<int32_t, int32_t>> st;  // node, next internal child cursor
        st.emplace_back(r, ichild_off[r]);
        while (!st.empty()) {
            auto& top = st.back();
            int32_t v = top.first;
            if (top.second < ichild_off[v + 1]) {
                int32_t c = ichild[top.second++];
                if (cut[c]) {
                    ch.level = std::max(ch.level, tmp[chunk_of[c]].level + 1);
                } else {
                    st.emplace_back(c, ichild_off[c]);
                }
            } else {
                ch.nodes.push_back(v);
                st.pop_back();
            }
        }
        for (int32_t v : ch.nodes) chunk_of[v] = int32_t(tmp.size());
        tmp.push_back(std::move(ch));
    }

    // ---- ticket orders ----
    const int32_t NC = int32_t(tmp.size());
    std::vector<int32_t> par_chunk(NC, -1), par_pos(NC, 0), n_kids(NC, 0);
    {
        std::vector<int32_t> pos_in_chunk(n_nodes, 0);
        for (auto& c : tmp)
            for (size_t k = 0; k < c.nodes.size(); k++) pos_in_chunk[c.nodes[k]] = int32_t(k);
        for (int32_t c = 0; c < NC; c++) {
            int32_t r = tmp[c].root_node;
            if (r == root) continue;
            par_chunk[c] = chunk_of[parent[r]];
            par_pos[c] = pos_in_chunk[parent[r]];
            n_kids[par_chunk[c]]++;
        }
    }
    // chunks were created children-first, so walking ids downwards visits parents before children
    std::vector<int64_t> t_end(NC, 0), est(NC, 0);  // ops left after the chunk ends (forward) / before it can start (backward)
    for (int32_t c = NC - 1; c >= 0; c--) {
        if (par_chunk[c] < 0) continue;
        int32_t q = par_chunk[c];
        int64_t after = int64_t(tmp[q].nodes.size()) - par_pos[c];  // parent's ops from the consuming op to its end
        if (chain_cut[tmp[c].root_node]) after = 1;  // chain segments do not wait for each other (speculative evaluation)
        t_end[c] = after + t_end[q];
        est[c] = est[q] + after;  // backward runs the parent chunk's ops in reverse: same count
    }
    std::vector<int32_t> fwd_order;
    fwd_order.reserve(NC);
    {
        // Kahn's algorithm with a max-heap on the remaining chain (own ops + t_end)
        auto key = [&](int32_t c) { return int64_t(tmp[c].nodes.size()) + t_end[c]; };
        auto cmp = [&](int32_t a, int32_t b) { return key(a) != key(b) ? key(a) < key(b) : a > b; };
        std::vector<int32_t> heap;
        std::vector<int32_t> pending = n_kids;
        for (int32_t c = 0; c < NC; c++)
            if (pending[c] == 0) heap.push_back(c);
        std::make_heap(heap.begin(), heap.end(), cmp);
        while (!heap.empty()) {
            std::pop_heap(heap.begin(), heap.end(), cmp);
            int32_t c = heap.back();
            heap.pop_back();
            fwd_order.push_back(c);
            int32_t q = par_chunk[c];
            if (q >= 0 && --pending[q] == 0) {
                heap.push_back(q);
                std::push_heap(heap.begin(), heap.end(), cmp);
            }
        }
        if (int32_t(fwd_order.size()) != NC) return "internal error: chunk order";
    }
    std::vector<int32_t> new_id(NC);
    for (int32_t k = 0; k < NC; k++) new_id[fwd_order[k]] = k;
    int32_t n_levels = 0;
    for (auto& c : tmp) n_levels = std::max(n_levels, c.level + 1);

    P.n_nodes = n_nodes;
    P.n_rows = n_rows;
    P.n_internal = n_internal;
    P.root = root;
    P.node_op.assign(n_nodes, -1);
    P.chunks.reserve(NC);
    int32_t op = 0;
    for (int32_t oi : fwd_order) {
        TmpChunk& c = tmp[oi];
        Chunk ck{op, op + int32_t(c.nodes.size()), 0, 0, -1, -1, 0, 0};
        for (int32_t v : c.nodes) P.node_op[v] = op++;
        P.chunks.push_back(ck);
    }
    if (op != n_internal) return "internal error: op count";
    P.bwd_order.resize(NC);
    std::iota(P.bwd_order.begin(), P.bwd_order.end(), 0);
    std::stable_sort(P.bwd_order.begin(), P.bwd_order.end(), [&](int32_t a, int32_t b) {
        int32_t oa = fwd_order[a], ob = fwd_order[b];
        if (est[oa] != est[ob]) return est[oa] < est[ob];
        return tmp[oa].nodes.size() > tmp[ob].nodes.size();
    });
    if (bwd_tail_chunks > 0) {
        // A chunk without child chunks can take any later ticket. The smallest of them go last, largest first: the
        // kernel then drains through short items instead of whatever happened to sit deepest in the tree.
        std::vector<int32_t> term;
        for (int32_t k = 0; k < NC; k++)
            if (n_kids[fwd_order[k]] == 0) term.push_back(k);
        std::stable_sort(term.begin(), term.end(),
                         [&](int32_t a, int32_t b) { return tmp[fwd_order[a]].nodes.size() < tmp[fwd_order[b]].nodes.size(); });
        if (int32_t(term.size()) > bwd_tail_chunks) term.resize(bwd_tail_chunks);
        std::vector<char> in_tail(NC, 0);
        for (int32_t k : term) in_tail[k] = 1;
        std::vector<int32_t> order;
        order.reserve(NC);
        for (int32_t k : P.bwd_order)
            if (!in_tail[k]) order.push_back(k);
        for (auto it = term.rbegin(); it != term.rend(); ++it) order.push_back(*it);
        P.bwd_order.swap(order);
    }
    P.level_order.resize(NC);
    std::iota(P.level_order.begin(), P.level_order.end(), 0);
    std::stable_sort(P.level_order.begin(), P.level_order.end(), [&](int32_t a, int32_t b) {
        int32_t oa = fwd_order[a], ob = fwd_order[b];
        if (tmp[oa].level != tmp[ob].level) return tmp[oa].level < tmp[ob].level;
        return tmp[oa].nodes.size() > tmp[ob].nodes.size();
    });
    P.level_chunk_begin.assign(n_levels + 1, 0);
    for (auto& c : tmp) P.level_chunk_begin[c.level + 1]++;
    for (int32_t l = 0; l < n_levels; l++) P.level_chunk_begin[l + 1] += P.level_chunk_begin[l];

    // ---- ops ----
    std::vector<int32_t> op_node(n_internal);
    for (int32_t v = 0; v < n_nodes; v++)
        if (P.node_op[v] >= 0) op_node[P.node_op[v]] = v;
    P.fwd_ops.resize(n_internal);
    P.bwd_ops.resize(n_internal);
    P.refs.reserve(n_nodes);
    P.bwd_leaves.reserve(n_rows);
    P.row_slot.assign(n_rows, -1);
    // which ops must park their assigned state for a child that does not follow them immediately (in reverse)
    std::vector<int32_t> fslot(n_internal, -1);
    std::vector<char> ext_child(n_internal, 0);  // some child of this op lives in another chunk
    int32_t n_fslots = 0;
    for (int32_t i = 0; i < n_internal; i++) {
        int32_t v = op_node[i];
        if (v == root) continue;
        int32_t pop = P.node_op[parent[v]];
        bool acc = (pop == i + 1) && chunk_of[v] == chunk_of[parent[v]];
        if (!acc && fslot[pop] < 0) fslot[pop] = n_fslots++;
        if (chunk_of[v] != chunk_of[parent[v]]) ext_child[pop] = 1;
    }
    P.n_fslots = n_fslots;
    int32_t max_arity = 0;
    for (int32_t i = 0; i < n_internal; i++) {
        int32_t v = op_node[i];
        FwdOp& f = P.fwd_ops[i];
        BwdOp& b = P.bwd_ops[i];
        f.ref_begin = int32_t(P.refs.size());
        f.flags = ((v == root) ? OPF_ROOT : 0) | (cut[v] ? OPF_SIGNAL : 0);
        b.node = v;
        b.flags = ((v == root) ? OPF_ROOT : 0) | (ext_child[i] ? OPF_SIGNAL_F : 0);
        b.leaf_begin = int32_t(P.bwd_leaves.size());
        b.leaf0_slot = b.leaf1_slot = 0;
        f.ref0 = f.ref1 = 0;
        f.row0 = f.row1 = 0;
        // leaves in Newick order, then internal children heavy -> light; the one computed by op i-1 of the
        // same chunk (if any) is taken from registers
        for (int32_t e = child_off[v]; e < child_off[v + 1]; e++) {
            int32_t c = child_idx[e];
            if (isz[c] == 0) {
                const int32_t slot = int32_t(P.bwd_leaves.size());  // consumption order of the forward program
                P.row_slot[leaf_row[c]] = slot;
                P.refs.push_back((REF_LEAF << 30) | uint32_t(slot));
                P.bwd_leaves.push_back(BwdLeaf{slot, c});
            }
        }
        for (int32_t e = ichild_off[v]; e < ichild_off[v + 1]; e++) {
            int32_t c = ichild[e];
            int32_t cop = P.node_op[c];
            bool acc = (cop == i - 1) && chunk_of[c] == chunk_of[v];
            bool ext = chunk_of[c] != chunk_of[v];
            if (chain_cut[c]) {
                P.refs.push_back((REF_CHAIN << 30) | uint32_t(cop));
                Chunk& ck = P.chunks[new_id[chunk_of[v]]];
                ck.chain_op = i;
                ck.chain_row = cop;
            } else {
                P.refs.push_back(acc ? (REF_ACC << 30) : ((REF_INT << 30) | (ext ? REF_EXT : 0u) | uint32_t(cop)));
            }
        }
        f.n_refs = int32_t(P.refs.size()) - f.ref_begin;
        if (f.n_refs > 0) f.ref0 = P.refs[f.ref_begin];
        if (f.n_refs > 1) f.ref1 = P.refs[f.ref_begin + 1];
        max_arity = std::max(max_arity, f.n_refs);
        if (f.n_refs == 2) {
            const uint32_t k0 = P.refs[f.ref_begin] >> 30, k1 = P.refs[f.ref_begin + 1] >> 30;
            int32_t type = FT_GENERIC;
            if (k0 == REF_LEAF && k1 == REF_LEAF) type = FT_LEAF_LEAF;
            else if (k0 == REF_LEAF && k1 == REF_ACC) type = FT_LEAF_ACC;
            else if (k0 == REF_LEAF && k1 == REF_INT) type = FT_LEAF_INT;
            else if (k0 == REF_INT && k1 == REF_ACC) type = FT_INT_ACC;
            f.flags |= type << OPF_TYPE_SHIFT;
        }
        f.max_arity_bits = f.n_refs <= 3 ? 2 : (f.n_refs <= 15 ? 4 : (f.n_refs <= 255 ? 8 : 20));
        b.n_leaves = int32_t(P.bwd_leaves.size()) - b.leaf_begin;
        if (b.n_leaves > 0) b.leaf0_slot = P.bwd_leaves[b.leaf_begin].row;
        if (b.n_leaves > 1) b.leaf1_slot = P.bwd_leaves[b.leaf_begin + 1].row;
        b.fslot_out = fslot[i];
        if (v == root) b.parent_ref = PARENT_ROOT;
        else {
            int32_t pop = P.node_op[parent[v]];
            bool acc = (pop == i + 1) && chunk_of[v] == chunk_of[parent[v]];
            b.parent_ref = acc ? PARENT_ACC : fslot[pop];
            if (chunk_of[v] != chunk_of[parent[v]]) b.flags |= OPF_PARENT_EXT;
            if (chain_cut[v]) {
                b.flags |= OPF_CHAIN_TOP;
                P.chunks[new_id[chunk_of[v]]].flags |= CHUNK_CHAIN_TOP;
            }
        }
    }
    P.max_arity = max_arity;
    for (const Chunk& ck : P.chunks) P.n_chain_segments += ck.chain_op >= 0;
    // the heavy path that starts at each chunk's root, as far as it stays inside the chunk
    for (const TmpChunk& c : tmp)
        for (int32_t v = c.root_node;;) {
            P.bwd_ops[P.node_op[v]].flags |= OPF_HEAVY;
            if (ichild_off[v] == ichild_off[v + 1]) break;
            const int32_t h = ichild[ichild_off[v]];
            if (cut[h]) break;
            v = h;
        }

    // ---- forward dependencies per chunk (external rows it reads), in consumption order; an external ref then
    //      carries its ordinal in that list instead of the row ----
    for (Chunk& ck : P.chunks) {
        ck.dep_begin = int32_t(P.deps.size());
        for (int32_t i = ck.op_begin; i < ck.op_end; i++) {
            FwdOp& f = P.fwd_ops[i];
            for (int32_t r = 0; r < f.n_refs; r++) {
                uint32_t& ref = P.refs[f.ref_begin + r];
                if ((ref >> 30) == REF_INT && (ref & REF_EXT)) {
                    const uint32_t k = uint32_t(P.deps.size()) - uint32_t(ck.dep_begin);
                    P.deps.push_back(int32_t(ref & REF_IDX_MASK));
                    ref = (REF_INT << 30) | REF_EXT | k;
                }
            }
            auto row_of = [&](uint32_t ref) {
                const uint32_t v = ref & REF_IDX_MASK;
                return int32_t(((ref >> 30) == REF_INT && (ref & REF_EXT)) ? uint32_t(P.deps[ck.dep_begin + v]) : v);
            };
            if (f.n_refs > 0) { f.ref0 = P.refs[f.ref_begin]; f.row0 = row_of(f.ref0); }
            if (f.n_refs > 1) { f.ref1 = P.refs[f.ref_begin + 1]; f.row1 = row_of(f.ref1); }
        }
        ck.dep_count = int32_t(P.deps.size()) - ck.dep_begin;
    }

    // ---- backward: parents inside the chunk are served from a per-warp shared-memory stack ----
    // Ops run in reverse; an op whose state is needed by a same-chunk child other than the very next op parks it in
    // stack entry d = number of such states currently parked (they are consumed strictly last-in-first-out because
    // the children's subtrees nest). Deeper than BWD_STACK_DEPTH (or needed by another chunk): the global slot.
    {
        std::vector<int32_t> pending(n_internal, 0);   // same-chunk non-ACC children still to come (reverse order)
        std::vector<int32_t> entry(n_internal, -1);
        for (const Chunk& ck : P.chunks) {
            for (int32_t i = ck.op_begin; i < ck.op_end; i++) {
                int32_t v = op_node[i];
                if (v == root) continue;
                int32_t pop = P.node_op[parent[v]];
                if (pop >= ck.op_begin && pop < ck.op_end && pop != i + 1) pending[pop]++;
            }
            int32_t depth = 0;
            for (int32_t i = ck.op_end - 1; i >= ck.op_begin; i--) {
                BwdOp& b = P.bwd_ops[i];
                int32_t v = op_node[i];
                if (v != root) {
                    int32_t pop = P.node_op[parent[v]];
                    bool same = pop >= ck.op_begin && pop < ck.op_end;
                    if (same && pop != i + 1 && entry[pop] >= 0) {
                        b.parent_ref = PARENT_STACK0 - entry[pop];
                        if (--pending[pop] == 0) depth--;   // last reader: the entry is free again
                    } else if (same && pop != i + 1) {
                        --pending[pop];
                    }
                }
                if (pending[i] > 0 && depth < BWD_STACK_DEPTH) {
                    entry[i] = depth++;
                    b.flags |= OPF_PUSH | (entry[i] << OPF_PUSH_SHIFT);
                    if (!ext_child[i]) b.fslot_out = -1;    // nobody reads the global slot
                }
            }
        }
    }
    return "";
}

}  // namespace pmb
