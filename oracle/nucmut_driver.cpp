// TEST INFRASTRUCTURE ONLY -- never linked into the product.
// Drives the reference's OWN struct NucMut (enums + struct, reference src/panman.hpp:27-313, extracted verbatim at build
// time into _ref/nucmut_extract.hpp) with restatements of the per-node post-processing loops that call its constructors:
//   MSA      : src/panman.cpp:1445-1466 (and :1625-1646)  -> NucMut(vector<tuple<int,int8_t,int8_t>>, start, end)
//   PanGraph : src/panman.cpp:1236-1253 (non-gap), :1255-1272 (gap) -> NucMut(vector<tuple<int x6>>, start, end)
//   wire     : src/panman.cpp:2866-2876 (getNodesPreorder): mutInfo on the wire = ((nucs >> (24 - 4 len)) << 8) + mutInfo,
//              and back through the reader constructor (src/panman.hpp:191-211)
// The loops are restated (they live inside the Tree constructor between TBB calls); the constructors, addNucCode, length()
// and the field layout are the reference's own compiled code. This pins oracle/fs_oracle.c's orc_merge_* and, through
// them, the device kernel merge_runs_kernel.
#include <algorithm>
#include <cstdint>
#include <tuple>
#include <vector>

#include "nucmut_shim.hpp"

namespace panmanUtils {
#include "nucmut_extract.hpp"
}

using panmanUtils::NucMut;

namespace {
void put(const std::vector<NucMut>& v, int32_t* pb, int32_t* sb, int32_t* pos, int32_t* gap, uint8_t* info, uint32_t* nucs) {
    for (size_t k = 0; k < v.size(); k++) {
        pb[k] = v[k].primaryBlockId;
        sb[k] = v[k].secondaryBlockId;
        pos[k] = v[k].nucPosition;
        gap[k] = v[k].nucGapPosition;
        info[k] = v[k].mutInfo;
        nucs[k] = v[k].nucs;
    }
}
}  // namespace

extern "C" {

// one node's MSA tuples (any order: sorted here like the reference does); outputs sized n
int64_t refnm_merge_msa(int64_t n, const int32_t* pos, const int8_t* type, const int8_t* code, int32_t* o_pb, int32_t* o_sb,
                        int32_t* o_pos, int32_t* o_gap, uint8_t* o_info, uint32_t* o_nucs) {
    if (n == 0) return 0;  // the reference's maps only hold nodes with at least one tuple
    std::vector<std::tuple<int, int8_t, int8_t>> u;
    for (int64_t i = 0; i < n; i++) u.emplace_back(pos[i], type[i], code[i]);
    std::vector<NucMut> nucMutation;
    std::sort(u.begin(), u.end());
    size_t currentStart = 0;
    for (size_t i = 1; i < u.size(); i++) {
        if (i - currentStart == 6 || std::get<0>(u[i]) != std::get<0>(u[i - 1]) + 1 || std::get<1>(u[i]) != std::get<1>(u[i - 1])) {
            nucMutation.emplace_back(u, currentStart, i);
            currentStart = i;
            continue;
        }
    }
    nucMutation.emplace_back(u, currentStart, u.size());
    put(nucMutation, o_pb, o_sb, o_pos, o_gap, o_info, o_nucs);
    return int64_t(nucMutation.size());
}

// one node's PanGraph 6-tuples (block, -1, pos, gapPos, type, code); gap = 0: the non-gap list, 1: the gap list
int64_t refnm_merge_pangraph(int gap, int64_t n, const int32_t* block, const int32_t* pos, const int32_t* gap_pos, const int32_t* type,
                             const int32_t* code, int32_t* o_pb, int32_t* o_sb, int32_t* o_pos, int32_t* o_gap, uint8_t* o_info,
                             uint32_t* o_nucs) {
    if (n == 0) return 0;
    std::vector<std::tuple<int, int, int, int, int, int>> u;
    for (int64_t i = 0; i < n; i++) u.emplace_back(block[i], -1, pos[i], gap_pos[i], type[i], code[i]);
    std::vector<NucMut> nucMutation;
    std::sort(u.begin(), u.end());
    size_t currentStart = 0;
    for (size_t i = 1; i < u.size(); i++) {
        bool split;
        if (!gap)
            split = i - currentStart == 6 || std::get<0>(u[i]) != std::get<0>(u[i - 1]) || std::get<2>(u[i]) != std::get<2>(u[i - 1]) + 1 ||
                    std::get<4>(u[i]) != std::get<4>(u[i - 1]);
        else
            split = i - currentStart == 6 || std::get<0>(u[i]) != std::get<0>(u[i - 1]) || std::get<2>(u[i]) != std::get<2>(u[i - 1]) ||
                    std::get<3>(u[i]) != std::get<3>(u[i - 1]) + 1 || std::get<4>(u[i]) != std::get<4>(u[i - 1]);
        if (split) {
            nucMutation.emplace_back(u, currentStart, i);
            currentStart = i;
            continue;
        }
    }
    nucMutation.emplace_back(u, currentStart, u.size());
    put(nucMutation, o_pb, o_sb, o_pos, o_gap, o_info, o_nucs);
    return int64_t(nucMutation.size());
}

// the value the writer stores (src/panman.cpp:2876) and the fields the reader constructor rebuilds from it (hpp:191-211)
uint32_t refnm_wire(uint8_t mut_info, uint32_t nucs, int32_t nuc_position, int32_t* back_pos, uint8_t* back_info, uint32_t* back_nucs,
                    int32_t* back_len, int32_t* back_type, int32_t* back_codes6) {
    NucMut mutation;
    mutation.nucPosition = nuc_position;
    mutation.nucGapPosition = -1;
    mutation.primaryBlockId = 0;
    mutation.secondaryBlockId = -1;
    mutation.mutInfo = mut_info;
    mutation.nucs = nucs;
    const uint32_t wire = (((mutation.nucs) >> (24 - mutation.length() * 4)) << 8) + mutation.mutInfo;
    panman::NucMut::Reader r;
    r.nucPosition = nuc_position;
    r.mutInfo = wire;
    r.nucGapExist = false;
    NucMut back(r, int64_t(0) << 32, false);
    *back_pos = back.nucPosition;
    *back_info = back.mutInfo;
    *back_nucs = back.nucs;
    *back_len = back.length();
    *back_type = int32_t(back.type());
    for (int i = 0; i < 6; i++) back_codes6[i] = back.getNucCode(i);
    return wire;
}

int refnm_sizeof(void) { return int(sizeof(NucMut)); }

}  // extern "C"
