// Host-side flattening of the tree into the schedule the kernels execute.
//
// The reference walks Node* pointers recursively, one column at a time (src/fitchSankoff.cpp:30-56,
// 96-171). Here the tree becomes a static "program":
//   * every internal node is one op; ops are grouped into chunks (connected pieces of the tree that
//     one warp evaluates sequentially for its 1024 columns);
//   * chunks are (a) whole bottom subtrees of at most chunk_nodes internal nodes and (b) heavy-path
//     segments of the remaining "top" tree (a top node stays in its parent's chunk iff it is the
//     parent's heaviest child), with tiny light subtrees inlined. This keeps the dependency chain
//     through the top of the tree inside as few warps as possible;
//   * a chunk may start before the chunks below it are finished: every cross-chunk read carries a
//     dependency (REF_EXT / OPF_PARENT_EXT) that the kernels resolve with per-(op, tile) done flags.
//     Chunks are listed in topological order (level-major), which is also the ticket order of the
//     persistent kernels: an item only ever waits for items with smaller tickets, so it cannot deadlock;
//   * within a chunk ops are in post-order with the lightest internal child last, so the most
//     recent result stays in registers (REF_ACC) and older siblings are re-read from the set matrix
//     while still L2-resident;
//   * the backward pass runs the same ops in reverse (parents before children);
//   * chain segments: a heavy path of the top tree longer than chunk_nodes ops (a caterpillar's spine, the backbone of
//     an unbalanced phylogeny) is cut into segments of about chunk_nodes ops. Taken literally each segment would wait
//     for the one below (forward) / above (backward) and a deep tree would run one segment at a time. The Fitch
//     kernels therefore evaluate a segment SPECULATIVELY: they track bounds on the unknown value entering the
//     segment (plane_math.h, FitchInterval / fitch_candidates_step) until, a few ops in, the value no longer depends
//     on it in any column -- in real alignments neighbouring leaves agree in almost every column -- continue exactly
//     from there, publish the segment's result, and only then wait for the neighbour to redo the few ops before
//     that point. Segments of one path thus run in parallel; results are exact, never approximate.
// Child order only affects scheduling: Fitch's AND/OR and Sankoff's sums are commutative, so results
// equal the reference's left-to-right recursion bit for bit.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace pmb {

enum : uint32_t { REF_INT = 0u, REF_LEAF = 1u, REF_ACC = 2u, REF_CHAIN = 3u };  // top 2 bits of a forward child ref
// REF_CHAIN: the heavy child of a node where a long heavy path was cut into segments; the index field holds the row
// (op) of that child, the top op of the segment below. Not listed in deps[]: the kernels either wait for it explicitly
// or evaluate the segment speculatively (see "chain segments" below).
constexpr uint32_t REF_EXT = 1u << 29;                          // REF_INT of another chunk: the index field holds the ORDINAL k of
                                                                // the dependency in the chunk's list; the row is deps[dep_begin + k]
constexpr uint32_t REF_IDX_MASK = (1u << 29) - 1u;
// BwdOp::parent_ref: >= 0 a global state slot (fslot); PARENT_ACC the previous op's registers; PARENT_ROOT the
// per-column parameters; <= PARENT_STACK0 entry (PARENT_STACK0 - parent_ref) of the warp's shared-memory state stack
enum : int32_t { PARENT_ACC = -1, PARENT_ROOT = -2, PARENT_STACK0 = -16 };
#ifndef PMB_BWD_STACK_DEPTH
#define PMB_BWD_STACK_DEPTH 3
#endif
constexpr int32_t BWD_STACK_DEPTH = PMB_BWD_STACK_DEPTH;   // entries of the per-warp state stack (640 B each)
enum : int32_t {
    OPF_ROOT = 1,
    OPF_SIGNAL = 2,      // forward: this op is a chunk root, publish its done flag after the store
    OPF_PARENT_EXT = 4,  // backward: the parent's state slot is written by another chunk, wait for it
    OPF_SIGNAL_F = 8,    // backward: another chunk reads this op's state slot, publish after the store
    OPF_TYPE_SHIFT = 8,  // forward: bits 8..11 hold the op's shape (FwdType) for the fast paths
    OPF_PUSH = 16,       // backward: park this op's assigned state in the warp's stack entry (flags >> OPF_PUSH_SHIFT) & 15
    OPF_PUSH_SHIFT = 12,
    OPF_CHAIN_TOP = 32,  // backward: top op of a chain segment, its parent is the bottom heavy op of the segment above
    OPF_HEAVY = 64       // backward: on the heavy path that starts at the chunk's root (inside the chunk)
};
// shape of a forward op; refs are stored in the order named (leaves first, then internal children heavy -> light)
enum FwdType : int32_t { FT_GENERIC = 0, FT_LEAF_LEAF = 1, FT_LEAF_ACC = 2, FT_LEAF_INT = 3, FT_INT_ACC = 4 };

struct FwdOp {  // 32 bytes; the op index is also the node's row ("slot") in the set matrix
    int32_t ref_begin;
    int32_t n_refs;
    int32_t flags;
    int32_t max_arity_bits;  // Sankoff counter width class for this op: 2, 4, 8 or 20
    uint32_t ref0, ref1;     // copies of the first two refs, so that binary ops never touch refs[]
    int32_t row0, row1;      // their rows (leaf slot / set-matrix op), resolved through deps[] for external refs
};

struct BwdOp {  // 32 bytes
    int32_t node;        // original node id (for emission)
    int32_t parent_ref;  // PARENT_ACC / PARENT_ROOT / fslot holding the parent's assigned state
    int32_t fslot_out;   // where to store this node's assigned state for later children, or -1
    int32_t leaf_begin;  // into bwd_leaves
    int32_t n_leaves;
    int32_t flags;
    int32_t leaf0_slot, leaf1_slot;  // slots of the first two leaf children (their node ids: bwd_leaves[leaf_begin + k].node)
};

struct BwdLeaf {
    int32_t row;  // leaf SLOT (see TreeProgram::row_slot)
    int32_t node;
};

enum : int32_t { CHUNK_CHAIN_TOP = 1 };
struct Chunk {  // 32 bytes
    int32_t op_begin, op_end;
    int32_t dep_begin, dep_count;  // forward: ops of other chunks whose rows this chunk reads (TreeProgram::deps)
    int32_t chain_op;              // op of this chunk holding the REF_CHAIN child (the deepest op of its heavy path), or -1
    int32_t chain_row;             // that child's row: the top op of the chain segment below
    int32_t flags;                 // CHUNK_CHAIN_TOP: the chunk is a chain segment with another segment above it
    int32_t pad1;
};

struct TreeProgram {
    int32_t n_nodes = 0, n_rows = 0, n_internal = 0, root = -1;
    int32_t n_fslots = 0;
    int32_t max_arity = 0;
    int32_t n_chain_segments = 0;             // chunks that receive a REF_CHAIN child
    std::vector<FwdOp> fwd_ops;
    std::vector<uint32_t> refs;
    std::vector<BwdOp> bwd_ops;
    std::vector<BwdLeaf> bwd_leaves;
    std::vector<Chunk> chunks;                // in forward ticket order (see below)
    std::vector<int32_t> deps;                // per chunk: external child ops, waited for once when the item starts
    // Ticket orders of the persistent kernels. Both are topological for their pass, so an item only waits for
    // smaller tickets. Forward = chunks[] order itself: critical-path list scheduling, among the chunks whose inputs
    // are all scheduled take the one with the longest remaining chain of ops up to the root. Backward = bwd_order:
    // ascending earliest start (ops executed on the way down from the root before the chunk's parent state exists).
    std::vector<int32_t> bwd_order;           // chunk indices
    std::vector<int32_t> level_order;         // chunk indices, level-major (for schedule = one launch per level)
    std::vector<int32_t> level_chunk_begin;   // n_levels + 1, into level_order
    std::vector<int32_t> node_op;             // node id -> op index, -1 for leaves
    std::vector<int32_t> row_slot;            // caller's leaf row -> slot in the packed leaf matrix: slots number the
                                              // leaves in the order the forward program reads them
    int n_levels() const { return int(level_chunk_begin.size()) - 1; }
};

// Returns "" on success, else a message. chunk_nodes = largest bottom subtree kept in one chunk (>= 1);
// inline_nodes = light subtrees up to this size are evaluated inside their parent's chunk (0 = never).
// bwd_tail_chunks = that many of the smallest chunks without child chunks get the LAST backward tickets, largest first,
// so that the persistent kernel drains through short items (0 = plain earliest-start order).
std::string build_tree_program(int32_t n_nodes, int32_t root, const int32_t* child_off, const int32_t* child_idx,
                               const int32_t* leaf_row, int32_t chunk_nodes, int32_t inline_nodes, TreeProgram* out,
                               int32_t bwd_tail_chunks = 0);

}  // namespace pmb
