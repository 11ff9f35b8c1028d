// What the reference's writer puts on the wire for one node, without Cap'n Proto: Tree::getNodesPreorder
// (reference src/panman.cpp:2854-2929) groups a node's NucMut by (primaryBlockId, secondaryBlockId) in std::map order, keeps
// nucMutation order inside a group, stores every piece as {nucPosition, nucGapPosition / nucGapExist,
// mutInfo = ((nucs >> (24 - 4 * length)) << 8) + mutInfo} (:2866-2876), and marks a group with the node's BlockMut of that
// block if there is one: blockMutExist, blockMutInfo, blockInversion (:2881-2899; a group without one carries the writer's
// defaults blockMutInfo = true -- it assigns the integer 2 -- and blockInversion = true); blockId = primary << 32 (+ secondary),
// blockGapExist (:2901-2908). Nodes go out in pre-order (:2926-2928), which is the id order of pmh_tree.
#include <map>
#include <utility>
#include <vector>

#include "../../include/panman_b200_host.h"
#include "host_tree.hpp"

namespace pmh {

void build_wire(const std::vector<pmh_nucmut>& nuc, const std::vector<pmh_blockmut>& blk, std::vector<pmh_wire_mutation>* muts,
                std::vector<pmh_wire_nuc>* nucs) {
    struct Group {
        std::vector<pmh_wire_nuc> nucs;
        int info = 2;  // 2 = no block mutation at this node for the block
        bool inversion = false, has_inv = false;
    };
    std::map<std::pair<int32_t, int32_t>, Group> groups;
    for (const pmh_nucmut& m : nuc) {
        pmh_wire_nuc w;
        w.nucPosition = m.nucPosition;
        w.nucGapPosition = m.nucGapPosition != -1 ? m.nucGapPosition : 0;
        w.nucGapExist = m.nucGapPosition != -1;
        const int length = m.mutInfo >> 4;
        w.mutInfo = ((m.nucs >> (24 - 4 * length)) << 8) + m.mutInfo;
        Group& g = groups[{m.primaryBlockId, m.secondaryBlockId}];
        g.nucs.push_back(w);
        g.info = 2;
    }
    for (const pmh_blockmut& b : blk) {
        Group& g = groups[{b.primaryBlockId, b.secondaryBlockId}];
        g.info = b.blockMutInfo ? 1 : 0;
        g.inversion = b.inversion != 0;
        g.has_inv = true;
    }
    muts->clear();
    nucs->clear();
    for (auto& kv : groups) {
        pmh_wire_mutation w;
        w.blockMutExist = kv.second.info != 2;
        w.blockMutInfo = kv.second.info != 0;  // the writer hands the integer (0, 1 or 2) to a Bool setter
        w.blockInversion = kv.second.info != 2 ? (kv.second.has_inv && kv.second.inversion) : 1;
        const int32_t pb = kv.first.first, sb = kv.first.second;
        w.blockId = sb != -1 ? (int64_t(pb) << 32) + sb : (int64_t(pb) << 32);
        w.blockGapExist = sb != -1;
        w.nuc_begin = int64_t(nucs->size());
        nucs->insert(nucs->end(), kv.second.nucs.begin(), kv.second.nucs.end());
        w.nuc_end = int64_t(nucs->size());
        muts->push_back(w);
    }
}

}  // namespace pmh
