// TEST INFRASTRUCTURE ONLY. Stand-ins for the two serialisation types struct NucMut's reader constructors mention
// (capnp `panman::NucMut::Reader`, protobuf `panmanOld::nucMut`; reference src/panman.hpp:191-233), so that the struct
// itself -- extracted VERBATIM from /root/reference/src/panman.hpp at build time into oracle/_ref/nucmut_extract.hpp by
// oracle/Makefile -- compiles without Cap'n Proto / protobuf. Nothing here restates reference logic.
#pragma once
#include <cstdint>
#include <tuple>
#include <vector>

namespace panman {
struct NucMut {
    struct Reader {
        int32_t nucPosition = 0, nucGapPosition = 0;
        bool nucGapExist = false;
        uint32_t mutInfo = 0;
        int32_t getNucPosition() const { return nucPosition; }
        int32_t getNucGapPosition() const { return nucGapPosition; }
        bool getNucGapExist() const { return nucGapExist; }
        uint32_t getMutInfo() const { return mutInfo; }
    };
};
}  // namespace panman

namespace panmanOld {
struct nucMut {
    int32_t nucposition_ = 0, nucgapposition_ = 0;
    bool nucgapexist_ = false;
    uint32_t mutinfo_ = 0;
    int32_t nucposition() const { return nucposition_; }
    int32_t nucgapposition() const { return nucgapposition_; }
    bool nucgapexist() const { return nucgapexist_; }
    uint32_t mutinfo() const { return mutinfo_; }
};
}  // namespace panmanOld
