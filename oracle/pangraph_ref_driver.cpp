// TEST INFRASTRUCTURE ONLY -- never linked into the product.
// The reference's own block-ordering code -- src/chaining.cpp (chain_align: seeds, 2-D range tree, chaining DP, consensus
// merge) and src/rotation.cpp (rotate_sample / rotate_alignment for circular paths) -- compiled VERBATIM into
// _ref/libpanman_pgorder.so (oracle/Makefile; TBB replaced by the serial stand-in in ref_shim/chain/), driven by a
// restatement of the part of Pangraph::Pangraph that calls it (src/panman.cpp:6259-6425: block numbers, rotation of circular
// paths against the first path, chaining of every further path into the growing consensus, re-numbering in consensus
// order) and of getAlignedSequences / getAlignedStrandSequences (:6427-6465).
// Like the reference, the paths live in a std::unordered_map<std::string, ...> and are visited in ITS iteration order:
// which path is "first" follows from libstdc++'s hashing, the same here as in a reference binary built with g++.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

// defined by the verbatim objects
void chain_align(std::vector<std::string>& consensus, std::vector<std::string>& sample, std::vector<int>& intSequenceConsensus,
                 std::vector<int>& intSequenceSample, size_t& numBlocks, std::vector<std::string>& consensus_new,
                 std::vector<int>& intSequenceConsensus_new, std::unordered_map<int, std::string>& intToString);
std::vector<std::string> rotate_sample(const std::vector<std::string>& consensus, std::vector<std::string>& sample,
                                       std::vector<int>& blockStrand, std::vector<size_t>& blockNumbers,
                                       std::unordered_map<std::string, int>& blockSizeMap, int& rotation_index, bool& invert);

namespace {
struct Order {
    std::vector<std::string> names;  // paths in the order given by the caller (JSON order)
    std::unordered_map<std::string, std::vector<std::string>> paths;
    std::unordered_map<std::string, std::vector<int>> strandPaths;
    std::unordered_map<std::string, std::vector<size_t>> blockNumbers;
    std::unordered_map<std::string, int> rotationIndexes;
    std::unordered_map<std::string, bool> sequenceInverted;
    std::unordered_map<std::string, std::vector<size_t>> intSequences;
    std::unordered_map<size_t, std::string> intIdToStringId;
    std::vector<size_t> topo;
    std::vector<std::string> visit;  // the iteration order of `paths`
    std::vector<std::string> aligned_walk;  // the iteration order of getAlignedSequences' result
};
}  // namespace

extern "C" {

// paths: n_paths names; blocks of path p are block_ids[path_off[p] .. path_off[p+1]) with strands[] (0/1); circular[p] != 0
// marks a circular path; block_len: consensus length per distinct block id (blockSizeMap), given as parallel arrays.
void* refpg_order(int n_paths, const char** names, const int64_t* path_off, const char** block_ids, const int32_t* strands,
                  const int32_t* circular, int n_blocks, const char** blk_ids, const int32_t* blk_len) {
    Order* o = new Order();
    bool any_circular = false;
    for (int p = 0; p < n_paths; p++) {
        o->names.push_back(names[p]);
        for (int64_t k = path_off[p]; k < path_off[p + 1]; k++) {
            o->paths[names[p]].push_back(block_ids[k]);
            o->strandPaths[names[p]].push_back(strands[k]);
        }
        if (circular[p]) any_circular = true;
    }
    std::unordered_map<std::string, int> blockSizeMap;
    for (int b = 0; b < n_blocks; b++) blockSizeMap[blk_ids[b]] = blk_len[b];
    // ---- src/panman.cpp:6259-6345
    if (any_circular) {
        std::vector<std::string> sample_base;
        int seq_count = 0;
        std::vector<std::string> sample_new;
        for (const auto& p : o->paths) {
            if (seq_count == 0) {
                std::unordered_map<std::string, size_t> baseBlockNumber;
                o->sequenceInverted[p.first] = false;
                o->rotationIndexes[p.first] = 0;
                for (const auto& block : p.second) {
                    o->blockNumbers[p.first].push_back(baseBlockNumber[block] + 1);
                    baseBlockNumber[block]++;
                    sample_base.push_back(block);
                }
            } else {
                std::unordered_map<std::string, size_t> baseBlockNumber;
                for (const auto& block : p.second) {
                    o->blockNumbers[p.first].push_back(baseBlockNumber[block] + 1);
                    baseBlockNumber[block]++;
                }
                std::vector<std::string> sample_dumy;
                sample_new.clear();
                for (const auto& block : p.second) sample_dumy.push_back(block);
                int rotation_index;
                bool invert = false;
                sample_new = rotate_sample(sample_base, sample_dumy, o->strandPaths[p.first], o->blockNumbers[p.first], blockSizeMap,
                                           rotation_index, invert);
                o->sequenceInverted[p.first] = invert;
                o->rotationIndexes[p.first] = rotation_index;
                o->paths[p.first] = sample_new;
            }
            seq_count++;
        }
    } else {
        for (auto p : o->paths) {
            std::unordered_map<std::string, size_t> baseBlockNumber;
            o->sequenceInverted[p.first] = false;
            o->rotationIndexes[p.first] = 0;
            for (const auto& block : p.second) {
                o->blockNumbers[p.first].push_back(baseBlockNumber[block] + 1);
                baseBlockNumber[block]++;
            }
        }
    }
    // ---- src/panman.cpp:6347-6425
    size_t numNodes = 0;
    std::unordered_map<int, std::string> intToString;
    int seqCount = 0;
    std::vector<std::string> consensus, sample, consensus_new;
    std::vector<int> intSequenceConsensus, intSequenceSample, intSequenceConsensusNew;
    for (const auto& p : o->paths) {
        o->visit.push_back(p.first);
        if (seqCount == 0) {
            for (const auto& block : p.second) {
                consensus.push_back(block);
                intToString[numNodes] = block;
                o->intSequences[p.first].push_back(numNodes);
                intSequenceConsensus.push_back(numNodes);
                numNodes++;
            }
        } else {
            intSequenceSample.clear();
            intSequenceConsensusNew.clear();
            sample.clear();
            consensus_new.clear();
            for (const auto& block : p.second) sample.push_back(block);
            chain_align(consensus, sample, intSequenceConsensus, intSequenceSample, numNodes, consensus_new, intSequenceConsensusNew,
                        intToString);
            for (auto& b : intSequenceSample) o->intSequences[p.first].push_back(b);
            consensus.clear();
            intSequenceConsensus.clear();
            for (auto& b : consensus_new) consensus.push_back(b);
            for (auto& b : intSequenceConsensusNew) intSequenceConsensus.push_back(b);
        }
        seqCount++;
    }
    int reorder = 0;
    std::unordered_map<int, int> order_map;
    for (auto& i : intSequenceConsensus) {
        order_map[i] = reorder;
        o->intIdToStringId[reorder] = intToString[i];
        o->topo.push_back(reorder);
        reorder++;
    }
    for (auto& m : o->intSequences)
        for (auto& s : m.second) s = order_map[s];
    {   // getAlignedSequences (src/panman.cpp:6447-6465) fills its result map in the iteration order of intSequences; the
        // block-level driver then walks that map (:881)
        std::unordered_map<std::string, std::vector<int>> alignedSequences;
        for (auto p : o->intSequences) alignedSequences[p.first].push_back(0);
        for (const auto& u : alignedSequences) o->aligned_walk.push_back(u.first);
    }
    return o;
}

void refpg_free(void* h) { delete static_cast<Order*>(h); }
int64_t refpg_n_topo(void* h) { return int64_t(static_cast<Order*>(h)->topo.size()); }
const char* refpg_topo_id(void* h, int64_t i) { return static_cast<Order*>(h)->intIdToStringId[size_t(i)].c_str(); }
const char* refpg_aligned_walk(void* h, int k) { return static_cast<Order*>(h)->aligned_walk[size_t(k)].c_str(); }
const char* refpg_visit(void* h, int k) { return static_cast<Order*>(h)->visit[size_t(k)].c_str(); }
int refpg_rotation_index(void* h, const char* name) { return static_cast<Order*>(h)->rotationIndexes[name]; }
int refpg_inverted(void* h, const char* name) { return static_cast<Order*>(h)->sequenceInverted[name] ? 1 : 0; }
int64_t refpg_path_len(void* h, const char* name) { return int64_t(static_cast<Order*>(h)->paths[name].size()); }
// per path after rotation: block id index into blk order is not needed -- ids come back as strings
const char* refpg_path_block(void* h, const char* name, int64_t k) { return static_cast<Order*>(h)->paths[name][size_t(k)].c_str(); }
int refpg_path_strand(void* h, const char* name, int64_t k) { return static_cast<Order*>(h)->strandPaths[name][size_t(k)]; }
int64_t refpg_path_number(void* h, const char* name, int64_t k) { return int64_t(static_cast<Order*>(h)->blockNumbers[name][size_t(k)]); }

// getAlignedSequences / getAlignedStrandSequences (src/panman.cpp:6427-6465) for one path: out[i] = topo id or -1, strand or -1;
// and the occurrence number of the matched path entry (blockCounts, src/panman.cpp:985-995), 0 where absent
void refpg_aligned(void* h, const char* name, int32_t* aligned, int32_t* strand, int32_t* number) {
    Order* o = static_cast<Order*>(h);
    const auto& seq = o->intSequences[name];
    const auto& st = o->strandPaths[name];
    const auto& nums = o->blockNumbers[name];
    size_t p1 = 0, p2 = 0, out = 0;
    const size_t n = o->topo.size();
    while (p1 < n && p2 < seq.size()) {
        if (o->topo[p1] == seq[p2]) {
            aligned[out] = int32_t(o->topo[p1]);
            strand[out] = st[p2];
            p2++;
        } else {
            aligned[out] = -1;
            strand[out] = -1;
        }
        out++;
        p1++;
    }
    while (out < n) {
        aligned[out] = -1;
        strand[out] = -1;
        out++;
    }
    int currentPtr = 0;
    for (size_t i = 0; i < n; i++) {
        number[i] = 0;
        if (aligned[i] != -1) number[i] = int32_t(nums[size_t(currentPtr++)]);
    }
}

}  // extern "C"
