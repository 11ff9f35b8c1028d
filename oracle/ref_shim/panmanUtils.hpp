// TEST INFRASTRUCTURE ONLY -- stand-in for the reference's src/panmanUtils.hpp.
//
// The reference's src/fitchSankoff.cpp is compiled VERBATIM, from where it lies
// under /root/reference, by oracle/Makefile (it is fed to g++ on stdin so that
// its `#include "panmanUtils.hpp"` resolves to this file instead of the real
// header, which needs TBB / Boost / jsoncpp / Cap'n Proto -- none of which exist
// in this image). This header declares only what that one translation unit
// touches:
//   * Node{identifier,parent,children}            (reference src/panman.hpp:555-598)
//   * Tree{root + the 19 Fitch/Sankoff methods}   (reference src/panman.hpp:846-902)
//   * NucMutationType / BlockMutationType values  (reference src/panman.hpp:46-72)
//   * SANKOFF_INF                                 (reference src/common.hpp:16)
//   * getNucleotideFromCode                       (reference src/panman.cpp:41-76,
//                                                  restated in oracle/ref_driver.cpp)
// Nothing here is copied from the reference: the declarations are the minimum
// needed for the member-function definitions in fitchSankoff.cpp to bind.
#pragma once

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <iostream>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

static const int SANKOFF_INF = 100000001;

namespace panmanUtils {

enum NucMutationType { NS = 0, ND = 1, NI = 2, NSNPS = 3, NSNPI = 4, NSNPD = 5, NNONE = 2000 };
enum BlockMutationType { BD = 0, BI = 1, BIn = 2, NONE = 1000 };

char getNucleotideFromCode(int code);

class Node {
  public:
    std::string identifier;
    Node* parent = nullptr;
    std::vector<Node*> children;
};

class Tree {
  public:
    Node* root = nullptr;

    using IntMap = std::unordered_map<std::string, int>;
    using VecMap = std::unordered_map<std::string, std::vector<int>>;
    using NucMutMap = std::unordered_map<std::string, std::pair<NucMutationType, char>>;
    using BlockMutMap = std::unordered_map<std::string, std::pair<BlockMutationType, bool>>;

    int nucFitchForwardPass(Node* node, IntMap& states, int refState = -1);
    int nucFitchForwardPassOpt(Node* node, IntMap& states);
    void nucFitchBackwardPass(Node* node, IntMap& states, int parentState, int defaultState = (1 << 28));
    void nucFitchBackwardPassOpt(Node* node, IntMap& states, int parentState, int defaultState = (1 << 28));
    void nucFitchAssignMutations(Node* node, IntMap& states, NucMutMap& mutations, int parentState);
    void nucFitchAssignMutationsOpt(Node* node, IntMap& states, NucMutMap& mutations, int parentState);

    std::vector<int> nucSankoffForwardPass(Node* node, VecMap& stateSets);
    std::vector<int> nucSankoffForwardPassOpt(Node* node, VecMap& stateSets);
    void nucSankoffBackwardPass(Node* node, VecMap& stateSets, IntMap& states, int parentPtr,
                                int defaultValue = (1 << 28));
    void nucSankoffBackwardPassOpt(Node* node, VecMap& stateSets, IntMap& states, int parentPtr,
                                   int defaultValue = (1 << 28));
    void nucSankoffAssignMutations(Node* node, IntMap& states, NucMutMap& mutations, int parentState);
    void nucSankoffAssignMutationsOpt(Node* node, IntMap& states, NucMutMap& mutations, int parentState);

    int blockFitchForwardPassNew(Node* node, IntMap& states);
    void blockFitchBackwardPassNew(Node* node, IntMap& states, int parentState, int defaultValue = (1 << 28));
    void blockFitchAssignMutationsNew(Node* node, IntMap& states, BlockMutMap& mutations, int parentState);

    std::vector<int> blockSankoffForwardPass(Node* node, VecMap& stateSets);
    void blockSankoffBackwardPass(Node* node, VecMap& stateSets, IntMap& states, int parentPtr,
                                  int defaultValue = (1 << 28));
    void blockSankoffAssignMutations(Node* node, IntMap& states, BlockMutMap& mutations, int parentState);
};

}  // namespace panmanUtils
