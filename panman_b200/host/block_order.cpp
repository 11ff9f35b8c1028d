// Block order of a PanGraph build: which block column every (path, block occurrence) lands in.
//
// Restated from the reference (its own files are compiled verbatim only by the test infrastructure, never into the product):
//   * Pangraph::Pangraph, second half     src/panman.cpp:6259-6425: occurrence numbers of duplicated blocks; circular paths
//     are rotated against the first path; every further path is chained into the growing consensus order; ids are renumbered
//     in consensus order
//   * rotate_alignment / rotate_sample     src/rotation.cpp:14-110 (ALLOW_INVERSIONS is never defined: front rotation only)
//   * chaining / find_chain / build_consensus / chain_align   src/chaining.cpp:72-310
//   * getAlignedSequences / getAlignedStrandSequences / blockCounts   src/panman.cpp:6427-6465, 985-995
// Ties are part of the result (which of several equally good seeds, chains or rotations wins), and the reference decides them
// through library behaviour: the iteration order of std::unordered_map (which path comes first; which best-scoring seed ends
// the chain), std::sort's placement of equal keys (the shape of the range tree) and first-strictly-better scans. The same
// containers, hash functions and calls are therefore used here on purpose, so that with the same C++ library (libstdc++) the
// order is the same; tests/test_pangraph_host.py compares with the reference's compiled code on random path sets.
#include <algorithm>
#include <cstdint>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "host_tree.hpp"

namespace pmh {

namespace {

typedef std::pair<int, int> Seed;  // (index in the consensus order, index in the sample path) of an equal block id

// ---------------------------------------------------------------- circular paths (src/rotation.cpp)
// Semi-global alignment of `sample`, read circularly, against `consensus`: the second member of a cell is the sample index
// its alignment started from. Scores: match +5, mismatch -2, gap -1; the row is circular in j (column 0 continues the last).
Seed rotate_alignment(const std::vector<std::string>& consensus, const std::vector<std::string>& sample) {
    const size_t m = sample.size();
    std::vector<Seed> row(m, Seed(-1, -1)), next(m, Seed(-1, -1));
    Seed best(0, 0);
    for (size_t i = 0; i < consensus.size(); i++) {
        for (size_t j = 0; j < m; j++) {
            const size_t before = j == 0 ? m - 1 : j - 1;
            const int left = row[j].first - 1;
            const int up = j == 0 ? -1 : next[before].first - 1;
            const int diag = consensus[i] == sample[j] ? row[before].first + 5 : row[before].first - 2;
            const Seed via_up(up, j == 0 ? -1 : next[before].second);
            if (diag >= left) {
                next[j] = diag >= up ? Seed(diag, row[before].second == -1 ? int(j) : row[before].second) : via_up;
            } else {
                next[j] = left >= up ? Seed(left, row[j].second) : via_up;
            }
            if (next[j].first > best.first) best = next[j];
        }
        row = next;
    }
    return best;
}

// Rotates a path (block ids, strands, occurrence numbers) so that it starts where it aligns best with the first path.
void rotate_path(const std::vector<std::string>& base, std::vector<std::string>& ids, std::vector<int>& strands,
                 std::vector<size_t>& numbers, int* rotation_index) {
    const int rotate = rotate_alignment(base, ids).second;
    const size_t n = ids.size();
    *rotation_index = int((n - rotate) % n);  // size_t arithmetic, as the reference does it
    std::vector<std::string> r_ids;
    std::vector<int> r_strands;
    std::vector<size_t> r_numbers;
    for (size_t i = 0; i < n; i++) {
        const int index = int((i + rotate) % n);
        r_ids.push_back(ids[index]);
        r_strands.push_back(strands[index]);
        r_numbers.push_back(numbers[index]);
    }
    ids.swap(r_ids);
    strands.swap(r_strands);
    numbers.swap(r_numbers);
}

// ---------------------------------------------------------------- chaining (src/chaining.cpp)
struct SeedHash {  // the reference's hashPair: xor of the two std::hash values, or one of them when they are equal
    size_t operator()(const Seed& p) const {
        const size_t a = std::hash<int>{}(p.first), b = std::hash<int>{}(p.second);
        return a != b ? a ^ b : a;
    }
};
struct SeedLink {
    int score;
    Seed from;
};
typedef std::unordered_map<Seed, SeedLink, SeedHash> SeedMap;

struct RangeNode {
    Seed point;
    int left = -1, right = -1;
};

bool by_x(const Seed& a, const Seed& b) { return a.first < b.first; }
bool by_xy(const Seed& a, const Seed& b) { return a.first == b.first ? a.second < b.second : a.first < b.first; }

// A binary tree over the seeds keyed by x alone: the middle element of the (re-sorted) range is the node. The re-sort with an
// x-only comparison is kept: where several seeds share x, std::sort decides which of them becomes the node.
int build_range_tree(std::vector<Seed>& pts, int start, int end, std::vector<RangeNode>& nodes) {
    if (start > end) return -1;
    std::sort(pts.begin() + start, pts.begin() + end + 1, by_x);
    const int mid = (start + end) / 2;
    const int me = int(nodes.size());
    nodes.emplace_back();
    nodes[me].point = pts[mid];
    const int l = build_range_tree(pts, start, mid - 1, nodes);
    nodes[me].left = l;
    const int r = build_range_tree(pts, mid + 1, end, nodes);
    nodes[me].right = r;
    return me;
}

// pre-order: the node itself, then the left subtree if the window may reach it, then the right one
void query_range(const std::vector<RangeNode>& nodes, int at, const Seed& lo, const Seed& hi, std::vector<Seed>& out) {
    if (at < 0) return;
    const Seed& p = nodes[at].point;
    if (p.first >= lo.first && p.first <= hi.first && p.second >= lo.second && p.second <= hi.second) out.push_back(p);
    if (nodes[at].left >= 0 && lo.first <= p.first) query_range(nodes, nodes[at].left, lo, hi, out);
    if (nodes[at].right >= 0 && hi.first >= p.first) query_range(nodes, nodes[at].right, lo, hi, out);
}

// Best predecessor of a seed among the seeds in the window of K positions below and left of it: match bonus 50, cost = the
// distance skipped; candidates dominated by one already seen (the "barrier") are passed over; the first strictly better wins.
void link_seed(const std::vector<RangeNode>& nodes, int root, const Seed& point, SeedMap& map) {
    const int K = 4000, match = 50;
    if (point.first == 0 && point.second == 0) {
        map[point] = SeedLink{match, Seed(-1, -1)};
        return;
    }
    std::vector<Seed> found;
    query_range(nodes, root, Seed(point.first - K > 0 ? point.first - K : 0, point.second - K > 0 ? point.second - K : 0),
                Seed(point.first - 1, point.second - 1), found);
    int best = 10;
    Seed from(-1, -1);
    int x_b = -1, y_b = -1;
    for (size_t k = found.size(); k-- > 0;) {
        const Seed& p = found[k];
        if (p.first <= x_b && p.second <= y_b) continue;
        const int cost = -(point.first - p.first + point.second - p.second);
        if (cost + map[p].score + match > best) {
            best = cost + map[p].score + match;
            from = p;
        }
        if (x_b < p.first) x_b = p.first - 1;
        if (y_b < p.second) y_b = p.second - 1;
    }
    map[point] = SeedLink{best, from};
}

// the chain of seeds from the best-scoring one back to its origin (so: in descending order)
std::vector<Seed> best_chain(const std::vector<std::string>& consensus, const std::vector<std::string>& sample) {
    std::vector<Seed> chain, pts;
    for (size_t i = 0; i < consensus.size(); i++)
        for (size_t j = 0; j < sample.size(); j++)
            if (consensus[i] == sample[j]) pts.emplace_back(int(i), int(j));
    std::sort(pts.begin(), pts.end(), by_xy);
    std::vector<RangeNode> nodes;
    nodes.reserve(pts.size());
    const int root = build_range_tree(pts, 0, int(pts.size()) - 1, nodes);
    if (pts.empty()) return chain;
    SeedMap map;
    for (const Seed& p : pts) map[p] = SeedLink{-1, Seed(-1, -1)};
    for (const Seed& p : pts) link_seed(nodes, root, p, map);
    int max_score = -1;
    Seed at{};
    for (const auto& kv : map)  // the map's iteration order decides between equal scores
        if (kv.second.score > max_score) {
            max_score = kv.second.score;
            at = kv.first;
        }
    for (;;) {
        chain.push_back(at);
        at = map[at].from;
        if (at == Seed(-1, -1)) break;
    }
    return chain;
}

}  // namespace

void order_blocks(const std::vector<PathIn>& in, BlockOrder* out) {
    // the reference's containers: paths keyed by name, filled in JSON order, visited in the map's own order
    std::unordered_map<std::string, std::vector<std::string>> paths;
    std::unordered_map<std::string, std::vector<int>> strands;
    std::unordered_map<std::string, std::vector<size_t>> numbers;
    bool circular = false;
    for (const PathIn& p : in) {
        for (size_t k = 0; k < p.blocks.size(); k++) {
            paths[p.name].push_back(p.blocks[k]);
            strands[p.name].push_back(p.strands[k]);
        }
        circular = circular || p.circular;
    }
    auto number_occurrences = [&](const std::string& name, const std::vector<std::string>& ids) {
        std::unordered_map<std::string, size_t> seen;
        for (const std::string& b : ids) numbers[name].push_back(++seen[b]);
    };
    if (circular) {
        std::vector<std::string> base;
        int k = 0;
        for (auto& p : paths) {
            number_occurrences(p.first, p.second);
            if (k++ == 0) {
                out->rotation_index[p.first] = 0;
                base = p.second;
            } else {
                int rot = 0;
                rotate_path(base, p.second, strands[p.first], numbers[p.first], &rot);
                out->rotation_index[p.first] = rot;
            }
        }
    } else {
        for (auto& p : paths) {
            number_occurrences(p.first, p.second);
            out->rotation_index[p.first] = 0;
        }
    }
    // chaining: ids are handed out as blocks are met; the consensus order grows path by path
    size_t n_ids = 0;
    std::unordered_map<int, std::string> id_name;
    std::unordered_map<std::string, std::vector<int>> seq_ids;
    std::vector<std::string> consensus;
    std::vector<int> consensus_ids;
    int k = 0;
    for (const auto& p : paths) {
        out->visit.push_back(p.first);
        if (k++ == 0) {
            for (const std::string& b : p.second) {
                consensus.push_back(b);
                id_name[int(n_ids)] = b;
                seq_ids[p.first].push_back(int(n_ids));
                consensus_ids.push_back(int(n_ids));
                n_ids++;
            }
            continue;
        }
        const std::vector<std::string>& sample = p.second;
        std::vector<Seed> chain = best_chain(consensus, sample);
        std::vector<std::string> merged;
        std::vector<int> merged_ids;
        std::vector<int>& mine = seq_ids[p.first];
        int prev_c = -1, prev_s = -1;
        auto take_unmatched = [&](int c_end, int s_end) {  // consensus blocks first, then the sample's own (new ids)
            for (int j = prev_c + 1; j < c_end; j++) {
                merged.push_back(consensus[j]);
                merged_ids.push_back(consensus_ids[j]);
            }
            for (int j = prev_s + 1; j < s_end; j++) {
                merged.push_back(sample[j]);
                mine.push_back(int(n_ids));
                id_name[int(n_ids)] = sample[j];
                merged_ids.push_back(int(n_ids));
                n_ids++;
            }
        };
        for (size_t t = chain.size(); t-- > 0;) {
            const int c = chain[t].first, s = chain[t].second;
            take_unmatched(c, s);
            merged.push_back(consensus[c]);
            mine.push_back(consensus_ids[c]);
            merged_ids.push_back(consensus_ids[c]);
            prev_c = c;
            prev_s = s;
        }
        take_unmatched(int(consensus.size()), int(sample.size()));
        consensus.swap(merged);
        consensus_ids.swap(merged_ids);
    }
    // renumber in consensus order
    std::unordered_map<int, int> renum;
    for (size_t i = 0; i < consensus_ids.size(); i++) {
        renum[consensus_ids[i]] = int(i);
        out->topo_ids.push_back(id_name[consensus_ids[i]]);
    }
    // The block-level driver walks alignedSequences (src/panman.cpp:881), a std::unordered_map that getAlignedSequences fills
    // in the iteration order of intSequences, itself filled in the order the paths were visited: the same chain of
    // containers gives the same walk.
    {
        std::unordered_map<std::string, int> int_sequences, aligned_sequences;
        for (const std::string& name : out->visit) int_sequences[name];
        for (const auto& kv : int_sequences) aligned_sequences[kv.first];
        for (const auto& kv : aligned_sequences) out->aligned_walk.push_back(kv.first);
    }
    const size_t n = out->topo_ids.size();
    for (auto& p : paths) {
        std::vector<int>& ids = seq_ids[p.first];
        for (int& s : ids) s = renum[s];
        // two-pointer walk along the consensus order: an occurrence that would have to go backwards is dropped together with
        // everything behind it (src/panman.cpp:6427-6465)
        std::vector<int32_t>& al = out->aligned[p.first];
        std::vector<int32_t>& st = out->strand[p.first];
        std::vector<int32_t>& nu = out->number[p.first];
        al.assign(n, -1);
        st.assign(n, -1);
        nu.assign(n, 0);
        size_t p1 = 0, p2 = 0;
        while (p1 < n && p2 < ids.size()) {
            if (int(p1) == ids[p2]) {
                al[p1] = int32_t(p1);
                st[p1] = strands[p.first][p2];
                p2++;
            }
            p1++;
        }
        size_t ptr = 0;
        for (size_t i = 0; i < n; i++)
            if (al[i] != -1) nu[i] = int32_t(numbers[p.first][ptr++]);
    }
}

}  // namespace pmh
