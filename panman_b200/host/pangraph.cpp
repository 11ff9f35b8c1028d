// Host adaptor for the PanGraph construction flow (panmanUtils -P pangraph.json -N tree.nwk): JSON -> per-block column
// batches in the pmb_run_nuc convention -> the block-level pass and one nucleotide pass per block on the device.
//
// Restated from the reference (nothing is copied; jsoncpp/TBB are not needed):
//   * JSON fields and how they are read            src/panman.cpp:6200-6258 (Pangraph::Pangraph): paths[].name / blocks[]
//     {id, strand}; blocks[]: sequence (upper-cased consensus), gaps {pos: length}, mutate [[{name, number}, [[pos(1-based),
//     char]]]], insert [[{..}, [[[pos, offset], string]]]], delete [[{..}, [[pos(1-based), length]]]]
//   * the aligned strings of a block                src/panman.cpp:1006-1045: consensus plus '-' filled gap slots (length+1
//     main positions, the last one '-'), then substitutions, insertions into the gap slots, deletions
//   * the column drivers                            src/panman.cpp:873-963 (blocks: 1 absent / 2 forward / 4 reverse) and
//     :1048-1232 (one column per main position j and per gap slot (j, k); sequences whose path lacks the block are
//     OMITTED from the state map; parent state = consensus character, '-' for gap slots)
// Which sequence the root is forced to where several qualify follows from the order in which the reference walks its maps:
//   * block level (src/panman.cpp:881-897, only with --reference): the LAST sequence whose name contains the reference string in
//     the walk of alignedSequences, a std::unordered_map -- reproduced with the same chain of containers (block_order.cpp);
//   * nucleotide level: the LAST such sequence in the walk of individualSequences (src/panman.cpp:1021-1022, 1131-1138), a
//     tbb::concurrent_unordered_map holding the sequences that own the block. Without --reference, gap columns and the
//     Sankoff branch have no override (guarded by reference.length()), while the Fitch main-column branch lacks that guard
//     (:1132): std::string::find("") matches everything and the root is forced to whichever owner is walked last.
//     TBB's map is a split-ordered list: it is walked in ascending order of the BIT-REVERSED hash of the key, whatever the
//     insertion order, and its default hasher for std::string is h = c ^ (h * 0x9E3779B97F4A7C15) over the characters
//     (tbb 2019_U9: tbb/internal/_concurrent_unordered_impl.h split_order_key_regular, _tbb_hash_compare_impl.h tbb_hasher).
//     TBB is not available here, so this rule is restated from its headers as we know them and is NOT pinned by executable
//     code; everything else of the flow is.
#include <cctype>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/panman_b200_host.h"
#include "host_tree.hpp"

namespace {

// ---------------------------------------------------------------- a small JSON reader (objects, arrays, strings, numbers)
struct JValue {
    enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
    bool b = false;
    double num = 0;
    std::string str;
    std::vector<JValue> arr;
    std::vector<std::pair<std::string, JValue>> obj;  // insertion order kept
    const JValue* get(const char* key) const {
        for (auto& kv : obj)
            if (kv.first == key) return &kv.second;
        return nullptr;
    }
};

struct JParser {
    const char* p;
    const char* end;
    std::string err;
    void ws() {
        while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) p++;
    }
    bool fail(const char* m) {
        if (err.empty()) err = std::string("JSON: ") + m;
        return false;
    }
    bool string(std::string& out) {
        if (p >= end || *p != '"') return fail("string expected");
        p++;
        out.clear();
        while (p < end && *p != '"') {
            if (*p == '\\') {
                if (++p >= end) return fail("bad escape");
                switch (*p) {
                case 'n': out += '\n'; break;
                case 't': out += '\t'; break;
                case 'r': out += '\r'; break;
                case 'b': out += '\b'; break;
                case 'f': out += '\f'; break;
                case 'u': {  // names and sequences are ASCII; keep the low byte
                    if (end - p < 5) return fail("bad \\u escape");
                    out += char(std::strtol(std::string(p + 1, p + 5).c_str(), nullptr, 16) & 0xFF);
                    p += 4;
                    break;
                }
                default: out += *p;
                }
                p++;
            } else {
                out += *p++;
            }
        }
        if (p >= end) return fail("unterminated string");
        p++;
        return true;
    }
    int depth = 0;
    bool value(JValue& v) {
        struct Depth {
            int& d;
            explicit Depth(int& x) : d(x) { d++; }
            ~Depth() { d--; }
        } guard(depth);
        if (depth > 64) return fail("nesting too deep");
        ws();
        if (p >= end) return fail("unexpected end");
        if (*p == '{') {
            v.kind = JValue::Obj;
            p++;
            ws();
            if (p < end && *p == '}') { p++; return true; }
            for (;;) {
                ws();
                std::string k;
                if (!string(k)) return false;
                ws();
                if (p >= end || *p != ':') return fail("':' expected");
                p++;
                v.obj.emplace_back(k, JValue());
                if (!value(v.obj.back().second)) return false;
                ws();
                if (p < end && *p == ',') { p++; continue; }
                if (p < end && *p == '}') { p++; return true; }
                return fail("',' or '}' expected");
            }
        }
        if (*p == '[') {
            v.kind = JValue::Arr;
            p++;
            ws();
            if (p < end && *p == ']') { p++; return true; }
            for (;;) {
                v.arr.emplace_back();
                if (!value(v.arr.back())) return false;
                ws();
                if (p < end && *p == ',') { p++; continue; }
                if (p < end && *p == ']') { p++; return true; }
                return fail("',' or ']' expected");
            }
        }
        if (*p == '"') {
            v.kind = JValue::Str;
            return string(v.str);
        }
        if (end - p >= 4 && !std::strncmp(p, "true", 4)) { v.kind = JValue::Bool; v.b = true; p += 4; return true; }
        if (end - p >= 5 && !std::strncmp(p, "false", 5)) { v.kind = JValue::Bool; v.b = false; p += 5; return true; }
        if (end - p >= 4 && !std::strncmp(p, "null", 4)) { v.kind = JValue::Null; p += 4; return true; }
        char* q = nullptr;
        v.num = std::strtod(p, &q);
        if (q == p) return fail("value expected");
        v.kind = JValue::Num;
        p = q;
        return true;
    }
};

// position of a key in the walk of a tbb::concurrent_unordered_map<std::string, ...> (see the header comment)
uint64_t tbb_walk_key(const std::string& s) {
    uint64_t h = 0;
    for (char ch : s) h = uint64_t(int64_t(ch)) ^ (h * 11400714819323198485ull);
    uint64_t r = 0;
    for (int b = 0; b < 64; b++) r |= ((h >> b) & 1ull) << (63 - b);
    return r | 1ull;
}

uint8_t code_of(unsigned char c) {  // getCodeFromNucleotide, src/panman.cpp:78-113: anything unlisted (incl. '-') -> 0
    static const struct T {
        uint8_t t[256];
        T() {
            std::memset(t, 0, sizeof t);
            const char* sym = "ACMGRSVTWYHKDBN";
            for (int i = 0; i < 15; i++) t[(unsigned char)sym[i]] = uint8_t(i + 1);
        }
    } tab;
    return tab.t[c];
}

struct BlockBatch {
    std::string id;
    int64_t n_cols = 0, stride = 0;
    std::vector<uint8_t> codes4, present, parent_code;
    std::vector<int8_t> root_override_fitch;  // see the header comment; empty = none anywhere
    bool any_override = false;
    std::vector<int32_t> col_pos, col_gap;    // (j, k); k = -1 for main columns
    // results of the last run: per-node lists, and their run-merged pieces (start column, mutInfo, nucs)
    std::vector<int64_t> off;
    std::vector<int32_t> pos;
    std::vector<uint8_t> tc;
    std::vector<int64_t> m_off;
    std::vector<int32_t> m_col;
    std::vector<uint8_t> m_info;
    std::vector<uint32_t> m_nucs;
};

}  // namespace

struct pmh_pangraph {
    pmh_tree tree;
    std::vector<BlockBatch> blocks;
    std::vector<uint8_t> block_states;  // n_leaves x n_blocks, one 3-state code per byte
    std::vector<int8_t> block_override; // n_blocks or empty
    BlockBatch block_level;             // results of the block-level pass
    bool has_reference = false;
    std::vector<int32_t> rotation_index;  // per leaf row: blocks a circular path was rotated by (Tree::rotationIndexes)
    std::vector<std::vector<pmh_nucmut>> nuc;      // Node::nucMutation after the run, in the reference's order
    std::vector<std::vector<pmh_blockmut>> blockmut;  // Node::blockMutation after the run, ascending block id
    mutable std::vector<pmh_wire_mutation> wire_muts;  // scratch of pmh_pangraph_wire (one node at a time)
    mutable std::vector<pmh_wire_nuc> wire_nucs;
};

namespace {

void set_err(char* err, size_t n, const std::string& m) {
    if (err && n) {
        std::strncpy(err, m.c_str(), n - 1);
        err[n - 1] = 0;
    }
}

std::string upper(std::string s) {
    for (char& c : s) c = char(std::toupper((unsigned char)c));
    return s;
}

}  // namespace

extern "C" {

pmh_pangraph* pmh_pangraph_load(const char* json, size_t json_len, const char* newick, const char* reference_c, char* err,
                                size_t err_len) {
    if (!json || !newick) { set_err(err, err_len, "null argument"); return nullptr; }
    try {
    const std::string reference = reference_c ? reference_c : "";
    std::unique_ptr<pmh_pangraph> g(new pmh_pangraph());
    g->has_reference = !reference.empty();
    {
        std::string nw(newick);
        size_t nl = nw.find('\n');
        if (nl != std::string::npos) nw.resize(nl);
        std::string e = pmh::parse_newick(nw, &g->tree.t);
        if (!e.empty()) { set_err(err, err_len, e); return nullptr; }
    }
    const pmh::HostTree& T = g->tree.t;
    JValue root;
    {
        JParser jp{json, json + json_len, "", 0};
        if (!jp.value(root)) { set_err(err, err_len, jp.err); return nullptr; }
    }
    const JValue* paths = root.get("paths");
    const JValue* blocks = root.get("blocks");
    if (!paths || !blocks || paths->kind != JValue::Arr || blocks->kind != JValue::Arr) {
        set_err(err, err_len, "PanGraph JSON needs \"paths\" and \"blocks\" arrays");
        return nullptr;
    }
    std::unordered_map<std::string, int32_t> row_of;
    for (int32_t v = 0; v < T.n_nodes(); v++)
        if (T.leaf_row[v] >= 0) row_of[T.names[v]] = T.leaf_row[v];
    // paths (src/panman.cpp:6203-6214), in JSON order; a path that is not a leaf of the tree still takes part in the block
    // ordering (as in the reference, which chains every path) but owns no row
    std::vector<pmh::PathIn> path_in;
    for (const JValue& path : paths->arr) {
        const JValue* name = path.get("name");
        const JValue* pb = path.get("blocks");
        const JValue* circ = path.get("circular");
        if (!name || name->kind != JValue::Str || !pb || pb->kind != JValue::Arr) { set_err(err, err_len, "path without name / blocks"); return nullptr; }
        pmh::PathIn pi;
        pi.name = name->str;
        pi.circular = circ && circ->kind == JValue::Bool && circ->b;
        for (const JValue& bl : pb->arr) {
            const JValue *id = bl.get("id"), *strand = bl.get("strand");
            if (!id || id->kind != JValue::Str) { set_err(err, err_len, "path block without id"); return nullptr; }
            pi.blocks.push_back(id->str);
            pi.strands.push_back(!strand || strand->kind != JValue::Bool || strand->b ? 1 : 0);
        }
        path_in.push_back(std::move(pi));
    }
    std::unordered_map<std::string, const JValue*> block_by_id;
    for (const JValue& blk : blocks->arr) {
        const JValue *id = blk.get("id"), *seq = blk.get("sequence");
        if (!id || id->kind != JValue::Str || !seq || seq->kind != JValue::Str) { set_err(err, err_len, "block without id / sequence"); return nullptr; }
        block_by_id[id->str] = &blk;
    }
    for (const pmh::PathIn& pi : path_in)
        for (const std::string& id : pi.blocks)
            if (!block_by_id.count(id)) { set_err(err, err_len, "path " + pi.name + " names the unknown block " + id); return nullptr; }
    // block columns: the consensus order of chain_align, duplicated blocks as columns of their own, circular paths rotated
    pmh::BlockOrder order;
    pmh::order_blocks(path_in, &order);
    g->rotation_index.assign(size_t(T.n_leaves), 0);
    for (auto& kv : order.rotation_index) {
        auto it = row_of.find(kv.first);
        if (it != row_of.end()) g->rotation_index[size_t(it->second)] = kv.second;
    }
    std::vector<std::string> name_of_row(size_t(T.n_leaves));
    for (int32_t v = 0; v < T.n_nodes(); v++)
        if (T.leaf_row[v] >= 0) name_of_row[size_t(T.leaf_row[v])] = T.names[v];
    auto matches_reference = [&](int32_t row) { return !reference.empty() && name_of_row[size_t(row)].find(reference) != std::string::npos; };
    const int32_t NB = int32_t(order.topo_ids.size()), L = T.n_leaves;
    g->blocks.resize(NB);
    g->block_states.assign(size_t(L) * NB, 0);
    if (!reference.empty()) g->block_override.assign(NB, -1);
    for (int32_t i = 0; i < NB; i++) {
        const JValue& blk = *block_by_id[order.topo_ids[i]];
        const JValue *seq = blk.get("sequence"), *gaps = blk.get("gaps");
        BlockBatch& B = g->blocks[i];
        B.id = order.topo_ids[i];
        const std::string cons = upper(seq->str);
        const int64_t len = int64_t(cons.size());
        std::map<int64_t, int64_t> gap_len;  // position -> slots, ascending
        if (gaps && gaps->kind == JValue::Obj)
            for (auto& kv : gaps->obj) {
                char* endp = nullptr;
                const long long pos = std::strtoll(kv.first.c_str(), &endp, 10);
                if (endp == kv.first.c_str() || *endp || pos < 0 || pos > len || kv.second.kind != JValue::Num || kv.second.num < 0) {
                    set_err(err, err_len, "block " + B.id + ": bad gap entry \"" + kv.first + "\"");
                    return nullptr;
                }
                gap_len[pos] = int64_t(kv.second.num);
            }
        // columns: main positions 0..len, then the gap slots in (position, slot) order
        std::map<std::pair<int64_t, int64_t>, int64_t> gap_col;
        for (int64_t j = 0; j <= len; j++) { B.col_pos.push_back(int32_t(j)); B.col_gap.push_back(-1); }
        for (auto& kv : gap_len)
            for (int64_t k = 0; k < kv.second; k++) {
                gap_col[{kv.first, k}] = int64_t(B.col_pos.size());
                B.col_pos.push_back(int32_t(kv.first));
                B.col_gap.push_back(int32_t(k));
            }
        B.n_cols = int64_t(B.col_pos.size());
        B.stride = ((B.n_cols + 1) / 2 + 15) / 16 * 16;
        B.codes4.assign(size_t(L) * size_t(B.stride), 0);
        B.present.assign(L, 0);
        B.parent_code.assign(B.n_cols, 0);
        for (int64_t j = 0; j < len; j++) B.parent_code[j] = code_of((unsigned char)cons[j]);
        // owners of this column: the paths aligned to it, with their strand and which occurrence of the block it is
        std::vector<std::vector<uint8_t>> rows(L);
        std::vector<int64_t> occurrence(size_t(L), 0);
        for (const pmh::PathIn& pi : path_in) {
            auto it = row_of.find(pi.name);
            if (it == row_of.end()) continue;
            auto al = order.aligned.find(pi.name);
            if (al == order.aligned.end() || al->second[size_t(i)] < 0) continue;
            const int32_t r = it->second;
            B.present[r] = 1;
            g->block_states[size_t(r) * NB + i] = order.strand[pi.name][size_t(i)] ? 1 : 2;
            occurrence[size_t(r)] = order.number[pi.name][size_t(i)];
            rows[r].assign(B.parent_code.begin(), B.parent_code.end());  // consensus, '-' in every gap slot
        }
        bool shape_ok = true;
        auto per_seq = [&](const char* field, auto&& fn) -> bool {
            const JValue* arr = blk.get(field);
            if (!arr || arr->kind != JValue::Arr) return true;
            for (const JValue& e : arr->arr) {
                if (e.kind != JValue::Arr || e.arr.size() != 2 || e.arr[1].kind != JValue::Arr) continue;
                const JValue *name = e.arr[0].get("name"), *number = e.arr[0].get("number");
                if (!name || name->kind != JValue::Str) continue;
                auto it = row_of.find(name->str);
                if (it == row_of.end() || rows[it->second].empty()) continue;
                // mutations are keyed by sequence AND occurrence (src/panman.cpp:1030-1044: [block][name][blockCounts])
                if (int64_t(number && number->kind == JValue::Num ? number->num : 1) != occurrence[size_t(it->second)]) continue;
                for (const JValue& m : e.arr[1].arr)
                    if (!fn(rows[it->second], m)) return false;
            }
            return true;
        };
        auto is_num = [](const JValue& v) { return v.kind == JValue::Num; };
        bool ok = per_seq("mutate", [&](std::vector<uint8_t>& row, const JValue& m) {
            if (m.kind != JValue::Arr || m.arr.size() < 2 || !is_num(m.arr[0]) || m.arr[1].kind != JValue::Str || m.arr[1].str.empty()) return shape_ok = false;
            const int64_t pos = int64_t(m.arr[0].num);
            if (pos < 1 || pos > len + 1) return false;
            row[pos - 1] = code_of((unsigned char)std::toupper((unsigned char)m.arr[1].str[0]));
            return true;
        });
        ok = ok && per_seq("insert", [&](std::vector<uint8_t>& row, const JValue& m) {
            if (m.kind != JValue::Arr || m.arr.size() < 2 || m.arr[0].kind != JValue::Arr || m.arr[0].arr.size() < 2 || !is_num(m.arr[0].arr[0]) ||
                !is_num(m.arr[0].arr[1]) || m.arr[1].kind != JValue::Str)
                return shape_ok = false;
            const int64_t pos = int64_t(m.arr[0].arr[0].num), off = int64_t(m.arr[0].arr[1].num);
            const std::string s = upper(m.arr[1].str);
            for (size_t t = 0; t < s.size(); t++) {
                auto it = gap_col.find({pos, off + int64_t(t)});
                if (it == gap_col.end()) return false;
                row[it->second] = code_of((unsigned char)s[t]);
            }
            return true;
        });
        ok = ok && per_seq("delete", [&](std::vector<uint8_t>& row, const JValue& m) {
            if (m.kind != JValue::Arr || m.arr.size() < 2 || !is_num(m.arr[0]) || !is_num(m.arr[1])) return shape_ok = false;
            const int64_t pos = int64_t(m.arr[0].num), ln = int64_t(m.arr[1].num);
            if (pos < 1 || ln < 0 || pos + ln - 1 > len + 1) return false;
            for (int64_t j = pos; j < pos + ln; j++) row[j - 1] = 0;
            return true;
        });
        if (!ok) {
            set_err(err, err_len, "block " + B.id + (shape_ok ? ": a mutation lies outside the block" : ": JSON: malformed mutation entry"));
            return nullptr;
        }
        // the owner walked last by the nucleotide-level driver, among all owners and among those matching --reference
        int32_t last_present = -1, ref_row = -1;
        uint64_t last_key = 0, ref_key = 0;
        for (int32_t r = 0; r < L; r++) {
            if (rows[r].empty()) continue;
            const uint64_t key = tbb_walk_key(name_of_row[size_t(r)]);
            if (last_present < 0 || key > last_key) { last_present = r; last_key = key; }
            if (matches_reference(r) && (ref_row < 0 || key > ref_key)) { ref_row = r; ref_key = key; }
            uint8_t* d = B.codes4.data() + size_t(r) * size_t(B.stride);
            for (int64_t c = 0; c < B.n_cols; c++) d[c >> 1] |= uint8_t(rows[r][c] << (4 * (c & 1)));
        }
        // root override (see the header comment)
        B.root_override_fitch.assign(B.n_cols, -1);
        if (!reference.empty()) {
            if (ref_row >= 0) {
                for (int64_t c = 0; c < B.n_cols; c++) B.root_override_fitch[c] = int8_t(rows[ref_row][c]);
                B.any_override = true;
            }
            for (const std::string& name : order.aligned_walk) {  // the last match of the block-level walk wins
                auto it = row_of.find(name);
                if (it != row_of.end() && matches_reference(it->second)) g->block_override[i] = int8_t(g->block_states[size_t(it->second) * NB + i]);
            }
        } else if (last_present >= 0) {
            for (int64_t c = 0; c <= len; c++) B.root_override_fitch[c] = int8_t(rows[last_present][c]);
            B.any_override = true;
        }
    }
    return g.release();
    } catch (const std::bad_alloc&) {
        set_err(err, err_len, "out of host memory");
        return nullptr;
    } catch (const std::exception& ex) {
        set_err(err, err_len, std::string("PanGraph JSON: ") + ex.what());
        return nullptr;
    }
}

void pmh_pangraph_free(pmh_pangraph* g) { delete g; }
const pmh_tree* pmh_pangraph_tree(const pmh_pangraph* g) { return &g->tree; }
int32_t pmh_pangraph_n_blocks(const pmh_pangraph* g) { return int32_t(g->blocks.size()); }
const char* pmh_pangraph_block_id(const pmh_pangraph* g, int32_t b) { return g->blocks[b].id.c_str(); }
const uint8_t* pmh_pangraph_block_states(const pmh_pangraph* g) { return g->block_states.data(); }
const int32_t* pmh_pangraph_rotation_index(const pmh_pangraph* g) { return g->rotation_index.data(); }
const int8_t* pmh_pangraph_block_override(const pmh_pangraph* g) { return g->block_override.empty() ? nullptr : g->block_override.data(); }
int64_t pmh_pangraph_n_cols(const pmh_pangraph* g, int32_t b) { return g->blocks[b].n_cols; }
const uint8_t* pmh_pangraph_codes4(const pmh_pangraph* g, int32_t b, int64_t* row_stride) {
    if (row_stride) *row_stride = g->blocks[b].stride;
    return g->blocks[b].codes4.data();
}
const uint8_t* pmh_pangraph_present(const pmh_pangraph* g, int32_t b) { return g->blocks[b].present.data(); }
const uint8_t* pmh_pangraph_parent_code(const pmh_pangraph* g, int32_t b) { return g->blocks[b].parent_code.data(); }
const int8_t* pmh_pangraph_root_override(const pmh_pangraph* g, int32_t b) { return g->blocks[b].root_override_fitch.data(); }
const int32_t* pmh_pangraph_col_pos(const pmh_pangraph* g, int32_t b) { return g->blocks[b].col_pos.data(); }
const int32_t* pmh_pangraph_col_gap(const pmh_pangraph* g, int32_t b) { return g->blocks[b].col_gap.data(); }

// One input of the shared runner: the column batch of a block as the passes see it.
struct BatchInput {
    const uint8_t* codes4;
    const uint8_t* present;        // NULL: every leaf takes part
    const int8_t* root_override;   // NULL: none
};

// Block-level pass, one nucleotide pass + device run-merge per block, then Node::blockMutation / Node::nucMutation.
static int run_batches(pmb_ctx* ctx, pmh_pangraph* g, int algo, const std::vector<BatchInput>& in, const uint8_t* block_states,
                       const int8_t* block_override, char* err, size_t err_len) {
    const pmh::HostTree& T = g->tree.t;
    int rc = pmb_set_tree(ctx, T.n_nodes(), T.root, T.child_off.data(), T.child_idx.data(), T.leaf_row.data());
    if (rc) { set_err(err, err_len, std::string("pmb_set_tree: ") + pmb_last_error(ctx)); return rc; }
    const int32_t NB = int32_t(g->blocks.size()), L = T.n_leaves;
    auto keep = [&](BlockBatch& B, const pmb_result& res) {
        B.off.assign(res.node_offsets, res.node_offsets + T.n_nodes() + 1);
        B.pos.assign(res.pos, res.pos + res.n_mut);
        B.tc.assign(res.type_code, res.type_code + res.n_mut);
    };
    pmb_result res;
    {   // block-level pass (src/panman.cpp:873-963, src/reroot.cpp:54-122): one 3-state column per block, parent state "absent"
        rc = pmb_run_block(ctx, algo, NB, L, block_states, block_override, &res);
        if (rc) { set_err(err, err_len, std::string("block pass: ") + pmb_last_error(ctx)); return rc; }
        keep(g->block_level, res);
    }
    // Node::blockMutation: BlockMut(blockId, (type, inversion)) of src/panman.hpp:467-484, applied at src/panman.cpp:971-980
    // (there in the iteration order of an unordered map; here ascending block id). Records arrive nuc-style: type 2 = block
    // insertion (code 2 = inserted inverted), type 1 = deletion, type 0 = inversion of a present block = (BD, true).
    g->blockmut.assign(T.n_nodes(), {});
    for (int32_t v = 0; v < T.n_nodes(); v++)
        for (int64_t k = g->block_level.off[v]; k < g->block_level.off[v + 1]; k++) {
            const uint8_t type = g->block_level.tc[k] >> 4, code = g->block_level.tc[k] & 15;
            pmh_blockmut m;
            m.primaryBlockId = g->block_level.pos[k];
            m.secondaryBlockId = -1;
            m.blockMutInfo = type == 2;
            m.inversion = type == 2 ? code == 2 : type == 0;
            g->blockmut[v].push_back(m);
        }
    for (int32_t i = 0; i < NB; i++) {
        BlockBatch& B = g->blocks[i];
        rc = pmb_run_nuc(ctx, algo, B.n_cols, L, in[i].codes4, B.stride, in[i].present, B.parent_code.data(), in[i].root_override,
                         nullptr, 0, 0, &res);
        if (rc) { set_err(err, err_len, "block " + B.id + ": " + pmb_last_error(ctx)); return rc; }
        keep(B, res);
        // greedy <= 6 run-merge on the device (src/panman.cpp:1236-1272): main positions merge on pos + 1, gap slots on equal
        // pos and gapPos + 1, so the first gap slot of every position never continues the column laid out before it
        std::vector<uint8_t> brk(size_t(B.n_cols), 0);
        for (int64_t c = 0; c < B.n_cols; c++) brk[c] = B.col_gap[c] == 0 ? 1 : 0;
        pmb_nucmut_result mr;
        rc = pmb_set_column_breaks(ctx, brk.data());
        if (!rc) rc = pmb_merge_runs(ctx, 0, 1, &mr);
        if (rc) { set_err(err, err_len, "block " + B.id + " run-merge: " + pmb_last_error(ctx)); return rc; }
        B.m_off.assign(mr.node_offsets, mr.node_offsets + T.n_nodes() + 1);
        B.m_col.assign(mr.nuc_position, mr.nuc_position + mr.n);
        B.m_info.assign(mr.mut_info, mr.mut_info + mr.n);
        B.m_nucs.assign(mr.nucs, mr.nucs + mr.n);
    }
    // Node::nucMutation: the non-gap pieces of every block first, then the gap pieces (the reference appends the merged
    // nonGapMutations map, then the merged gapMutations map; within a node both are sorted by block, position, gap slot)
    g->nuc.assign(T.n_nodes(), {});
    for (int pass = 0; pass < 2; pass++)
        for (int32_t i = 0; i < NB; i++) {
            const BlockBatch& B = g->blocks[i];
            for (int32_t v = 0; v < T.n_nodes(); v++)
                for (int64_t k = B.m_off[v]; k < B.m_off[v + 1]; k++) {
                    const int32_t c = B.m_col[k];
                    if ((B.col_gap[c] >= 0) != (pass == 1)) continue;
                    pmh_nucmut m;
                    m.nucPosition = B.col_pos[c];
                    m.nucGapPosition = B.col_gap[c];
                    m.primaryBlockId = i;
                    m.secondaryBlockId = -1;
                    m.mutInfo = B.m_info[k];
                    m.nucs = B.m_nucs[k];
                    g->nuc[v].push_back(m);
                }
        }
    return PMB_OK;
}

int pmh_pangraph_run(pmb_ctx* ctx, pmh_pangraph* g, int algo, char* err, size_t err_len) {
    if (!ctx || !g || (algo != PMB_ALGO_FITCH && algo != PMB_ALGO_SANKOFF)) { set_err(err, err_len, "bad argument"); return PMB_ERR_INVALID; }
    try {
        std::vector<BatchInput> in;
        for (BlockBatch& B : g->blocks) {
            // the Sankoff branch guards every override with reference.length(); the Fitch main-column branch does not
            const bool use_override = B.any_override && (algo == PMB_ALGO_FITCH || g->has_reference);
            in.push_back({B.codes4.data(), B.present.data(), use_override ? B.root_override_fitch.data() : nullptr});
        }
        return run_batches(ctx, g, algo, in, g->block_states.data(), g->block_override.empty() ? nullptr : g->block_override.data(), err,
                           err_len);
    } catch (const std::bad_alloc&) {
        set_err(err, err_len, "out of host memory");
        return PMB_ERR_OOM;
    } catch (const std::exception& ex) {
        set_err(err, err_len, std::string("pmh_pangraph_run: ") + ex.what());
        return PMB_ERR_INVALID;
    }
}

// Tree::reroot (reference src/reroot.cpp:4-261) on the loaded graph: the tree is transformed (pmh::reroot_tree), then EVERY
// block column and nucleotide column is inferred again by Fitch with the root forced to the new root's own state. Unlike
// the -P build, every leaf takes part in every column, with the characters the built PanMAT yields for it (the mutations of
// the last pmh_pangraph_run replayed root -> tip over the consensus), and the state "absent" at block level where it lacks
// the block. Afterwards pmh_pangraph_tree / _nucmut / _blockmut / _result describe the new tree.
int pmh_pangraph_reroot(pmb_ctx* ctx, pmh_pangraph* g, const char* leaf_name, char* err, size_t err_len) {
    if (!ctx || !g || !leaf_name) { set_err(err, err_len, "bad argument"); return PMB_ERR_INVALID; }
    try {
        const int32_t tip = pmh::find_node(g->tree.t, leaf_name);
        if (tip < 0) { set_err(err, err_len, std::string("Sequence with name ") + leaf_name + " not found!"); return PMB_ERR_INVALID; }
        pmh::HostTree nt;
        std::string e = pmh::reroot_tree(g->tree.t, tip, &nt);
        if (!e.empty()) { set_err(err, err_len, e); return PMB_ERR_INVALID; }
        const int32_t row = g->tree.t.leaf_row[tip];
        const pmh::HostTree& OT = g->tree.t;  // the tree the PanMAT was built on
        const int32_t NB = int32_t(g->blocks.size()), L = OT.n_leaves;
        for (const BlockBatch& B : g->blocks)
            if (int32_t(B.off.size()) != OT.n_nodes() + 1) { set_err(err, err_len, "reroot works on a built PanMAT: call pmh_pangraph_run first"); return PMB_ERR_NO_INPUT; }
        std::vector<std::vector<uint8_t>> codes(NB);
        std::vector<std::vector<int8_t>> over(NB);
        std::vector<int8_t> block_over(size_t(NB), 0);
        std::vector<BatchInput> in;
        for (int32_t i = 0; i < NB; i++) {
            const BlockBatch& B = g->blocks[i];
            // The sequences reroot works on are those the PanMAT yields (getSequenceFromReference, src/reroot.cpp:19-35): the
            // block consensus with the mutations on the path root -> tip applied in order. For a sequence that owns the block
            // that is its aligned string again; one that lacks it shows its nearest defined ancestor's characters.
            codes[i].assign(size_t(L) * size_t(B.stride), 0);
            std::vector<uint8_t> cur(B.parent_code.begin(), B.parent_code.end());
            struct Undo { int32_t col; uint8_t old; };
            std::vector<Undo> undo;
            struct Frame { int32_t node; int32_t next_child; size_t undo_mark; };
            std::vector<Frame> stack;
            auto enter = [&](int32_t v) {
                stack.push_back({v, OT.child_off[v], undo.size()});
                for (int64_t k = B.off[v]; k < B.off[v + 1]; k++) {
                    undo.push_back({B.pos[k], cur[B.pos[k]]});
                    cur[B.pos[k]] = B.tc[k] & 15;  // a deletion carries '-' = 0
                }
                if (OT.leaf_row[v] >= 0) {
                    uint8_t* d = codes[i].data() + size_t(OT.leaf_row[v]) * size_t(B.stride);
                    for (int64_t c = 0; c < B.n_cols; c++) d[c >> 1] |= uint8_t(cur[c] << (4 * (c & 1)));
                }
            };
            enter(OT.root);
            while (!stack.empty()) {
                Frame& f = stack.back();
                if (f.next_child < OT.child_off[f.node + 1]) {
                    enter(OT.child_idx[f.next_child++]);
                } else {
                    while (undo.size() > f.undo_mark) {
                        cur[undo.back().col] = undo.back().old;
                        undo.pop_back();
                    }
                    stack.pop_back();
                }
            }
            over[i].resize(size_t(B.n_cols));
            const uint8_t* nr = codes[i].data() + size_t(row) * size_t(B.stride);
            for (int64_t c = 0; c < B.n_cols; c++) over[i][c] = int8_t((nr[c >> 1] >> (4 * (c & 1))) & 15);
            block_over[i] = int8_t(g->block_states[size_t(row) * NB + i]);
            in.push_back({codes[i].data(), nullptr, over[i].data()});
        }
        g->tree.t = std::move(nt);
        return run_batches(ctx, g, PMB_ALGO_FITCH, in, g->block_states.data(), block_over.data(), err, err_len);
    } catch (const std::bad_alloc&) {
        set_err(err, err_len, "out of host memory");
        return PMB_ERR_OOM;
    } catch (const std::exception& ex) {
        set_err(err, err_len, std::string("pmh_pangraph_reroot: ") + ex.what());
        return PMB_ERR_INVALID;
    }
}

int64_t pmh_pangraph_wire(const pmh_pangraph* g, int32_t node, const pmh_wire_mutation** mutations, const pmh_wire_nuc** nucs) {
    if (!g || node < 0 || node >= int32_t(g->nuc.size())) return -1;
    try {
        pmh::build_wire(g->nuc[size_t(node)], g->blockmut[size_t(node)], &g->wire_muts, &g->wire_nucs);
    } catch (const std::exception&) {
        return -1;
    }
    if (mutations) *mutations = g->wire_muts.data();
    if (nucs) *nucs = g->wire_nucs.data();
    return int64_t(g->wire_muts.size());
}

int64_t pmh_pangraph_n_blockmut(const pmh_pangraph* g, int32_t node) { return node < int32_t(g->blockmut.size()) ? int64_t(g->blockmut[node].size()) : 0; }
const pmh_blockmut* pmh_pangraph_blockmut(const pmh_pangraph* g, int32_t node) { return g->blockmut[node].data(); }

int64_t pmh_pangraph_n_nucmut(const pmh_pangraph* g, int32_t node) { return int64_t(g->nuc[node].size()); }
const pmh_nucmut* pmh_pangraph_nucmut(const pmh_pangraph* g, int32_t node) { return g->nuc[node].data(); }

int64_t pmh_pangraph_result(const pmh_pangraph* g, int32_t block, const int64_t** node_offsets, const int32_t** pos,
                            const uint8_t** type_code) {
    const BlockBatch& B = block < 0 ? g->block_level : g->blocks[block];
    if (node_offsets) *node_offsets = B.off.data();
    if (pos) *pos = B.pos.data();
    if (type_code) *type_code = B.tc.data();
    return int64_t(B.pos.size());
}

}  // extern "C"
