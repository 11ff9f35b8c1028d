// Host-side tree with the reference's node numbering (see include/panman_b200_host.h).
#pragma once
#include <cstdint>
#include <string>
#include <unordered_map>
#include <vector>

namespace pmh {

struct HostTree {
    std::vector<std::string> names;
    std::vector<int32_t> parent, child_off, child_idx, leaf_row;
    int32_t root = 0;
    int32_t n_leaves = 0;
    int32_t n_nodes() const { return int32_t(names.size()); }
    bool has_polytomy() const;
};

// reference src/panman.cpp:265-295 (stringSplit, apostrophe-aware)
void split_quote_aware(const std::string& s, char delim, std::vector<std::string>& words);

// reference src/panman.cpp:310-450 (createTreeFromNewickString). Returns "" or an error message.
std::string parse_newick(const std::string& newick, HostTree* out);

// Block order of a PanGraph build (block_order.cpp; reference src/panman.cpp:6259-6465, src/chaining.cpp, src/rotation.cpp).
struct PathIn {
    std::string name;
    std::vector<std::string> blocks;  // block ids along the path
    std::vector<int> strands;         // 1 forward, 0 reverse
    bool circular = false;
};
struct BlockOrder {
    std::vector<std::string> topo_ids;  // block id of every block column, in consensus order (duplicated blocks repeat)
    // per path: for block column i the column itself or -1 (absent), the strand or -1, the occurrence number ("number" of
    // the PanGraph JSON) of the path entry that landed there or 0
    std::unordered_map<std::string, std::vector<int32_t>> aligned, strand, number;
    std::unordered_map<std::string, int> rotation_index;  // circular paths: blocks the path was rotated by
    std::vector<std::string> visit;                       // the order in which the paths were chained
    std::vector<std::string> aligned_walk;                // the order in which the block-level driver walks the sequences
};
void order_blocks(const std::vector<PathIn>& paths_in_json_order, BlockOrder* out);

// wire.cpp: one node's mutations as Tree::getNodesPreorder writes them (src/panman.cpp:2854-2929)
}  // namespace pmh
struct pmh_nucmut;
struct pmh_blockmut;
struct pmh_wire_mutation;
struct pmh_wire_nuc;
namespace pmh {
void build_wire(const std::vector<pmh_nucmut>& nuc, const std::vector<pmh_blockmut>& blk, std::vector<pmh_wire_mutation>* muts,
                std::vector<pmh_wire_nuc>* nucs);

// reference Tree::transform (src/panman.cpp:5831-5906): `tip` becomes the first child of a new root. Defined in reroot.cpp.
std::string reroot_tree(const HostTree& in, int32_t tip, HostTree* out);
int32_t find_node(const HostTree& t, const std::string& name);

}  // namespace pmh

// the opaque handle of include/panman_b200_host.h
struct pmh_tree {
    pmh::HostTree t;
};
