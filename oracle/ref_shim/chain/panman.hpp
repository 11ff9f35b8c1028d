// TEST INFRASTRUCTURE ONLY: what src/chaining.cpp needs from the reference's panman.hpp (which itself needs TBB, Boost,
// jsoncpp, Cap'n Proto and protobuf): the two Node members its tree-guided variant reads (src/panman.hpp:555-598).
#pragma once
#include <string>
#include <vector>
namespace panmanUtils {
struct Node {
    std::string identifier;
    std::vector<Node*> children;
};
}  // namespace panmanUtils
