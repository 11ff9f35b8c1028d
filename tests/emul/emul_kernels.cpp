// TEST INFRASTRUCTURE ONLY -- lane-by-lane CPU emulation of the CUDA kernels in
// panman_b200/csrc/pmb_kernels.cuh, sharing plane_math.h and tree_program.{h,cpp} with the product.
//
// Purpose: this build container has no GPU, and GPU minutes are rationed, so the bit-plane logic, the tree
// program (chunks, REF_ACC, fslots, levels) and the staging/gather index math are first validated here against
// the oracle. It is NOT a fallback: nothing under panman_b200/ links or loads this file, and the library
// refuses to compute without a device. The structure below deliberately mirrors the kernels one to one
// (same layouts, same op walk, a "warp" is a loop over 32 lanes).
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../panman_b200/csrc/plane_math.h"
#include "../../panman_b200/csrc/tree_program.h"

using namespace pmb;

namespace {

constexpr int TILE_COLS = 1024;
struct U4 { uint32_t x, y, z, w; };

struct Emu {
    TreeProgram P;
    int T = 0, flags = 0, algo = 0;  // flags: 1 block mode
    std::vector<U4> leaf_planes, sets, colparams, states;
    std::vector<uint32_t> fstore;  // [fslot][tile][160 words]
    std::vector<uint8_t> present;
    bool have_present = false;
    std::vector<unsigned> done, fdone;   // dependency flags, as in the kernels (value 1 = published this run)
    bool order_violation = false;        // an item read data of an item that had not run yet (would be a wait/deadlock)
    bool spec_mismatch = false;          // a speculatively resolved value differed from the exact one
    long long spec_items = 0, spec_resolved = 0;
    std::vector<unsigned long long> dir;
    std::vector<uint16_t> staging;
    unsigned long long pool = 0;
    unsigned error = 0;
};

void load16(const U4* base, int lane, uint32_t S[16]) {
    for (int j = 0; j < 4; j++) {
        U4 v = base[j * 32 + lane];
        S[4 * j] = v.x; S[4 * j + 1] = v.y; S[4 * j + 2] = v.z; S[4 * j + 3] = v.w;
    }
}
void store16(U4* base, int lane, const uint32_t S[16]) {
    for (int j = 0; j < 4; j++) base[j * 32 + lane] = U4{S[4 * j], S[4 * j + 1], S[4 * j + 2], S[4 * j + 3]};
}

// warp-level emit: lanes in order, bits in order => ascending column order
struct WarpMut { uint32_t mut[32], P[32][4], F[32][4]; };
void emit(Emu& E, int node, int tile, const WarpMut& w) {
    int total = 0;
    for (int l = 0; l < 32; l++) total += __builtin_popcount(w.mut[l]);
    if (!total) return;
    unsigned long long base = E.pool;
    E.pool += total;
    E.dir[(size_t)node * E.T + tile] = (base << 11) | (unsigned long long)total;
    E.staging.resize(E.pool);
    uint16_t* dst = E.staging.data() + base;
    for (int l = 0; l < 32; l++) {
        uint32_t t0, t1;
        mutation_type(w.P[l], w.F[l], t0, t1);
        uint32_t m = w.mut[l];
        while (m) {
            int b = __builtin_ctz(m);
            m &= m - 1;
            uint32_t code = ((w.F[l][0] >> b) & 1u) | (((w.F[l][1] >> b) & 1u) << 1) | (((w.F[l][2] >> b) & 1u) << 2) |
                            (((w.F[l][3] >> b) & 1u) << 3);
            uint32_t type = ((t0 >> b) & 1u) | (((t1 >> b) & 1u) << 1);
            *dst++ = uint16_t((l * 32 + b) | (code << 10) | (type << 14));
        }
    }
}

uint32_t present_mask(const Emu& E, int row) { return (!E.have_present || E.present[row]) ? 0xFFFFFFFFu : 0u; }

void store_state(Emu& E, int node, int tile, int lane, const uint32_t F[4], uint32_t vis) {
    if (E.states.empty()) return;
    U4* s = E.states.data() + ((size_t)node * E.T + tile) * 64;
    s[lane] = U4{F[0], F[1], F[2], F[3]};
    s[32 + lane] = U4{vis, 0, 0, 0};
}

template <int B>
void sankoff_fwd_op(Emu& E, const Chunk& ck, const FwdOp& f, int tile, int lane, uint32_t accG[16], uint32_t accH[16]) {
    const size_t T = E.T;
    SankoffFold<B> fold;
    fold.reset();
    for (int r = 0; r < f.n_refs; r++) {
        uint32_t ref = E.P.refs[f.ref_begin + r], kind = ref >> 30, idx = ref & REF_IDX_MASK;
        if (kind == REF_LEAF) {
            U4 c = E.leaf_planes[((size_t)tile * E.P.n_rows + idx) * 32 + lane];
            uint32_t cc[4] = {c.x, c.y, c.z, c.w}, pr = present_mask(E, idx);
            if ((E.flags & 1) && !pr) { cc[0] = cc[1] = cc[2] = cc[3] = 0; pr = 0xFFFFFFFFu; }
            fold.add_leaf(cc, pr);
        } else if (kind == REF_ACC) {
            fold.add_set(accG, sankoff_none(accG, accH));
        } else {
            if (kind == REF_CHAIN) {
                if (E.done[(size_t)tile * E.P.n_internal + idx] != 1) E.order_violation = true;
            } else {
                if (ref & REF_EXT) idx = uint32_t(E.P.deps[ck.dep_begin + idx]);
                if ((ref & REF_EXT) && E.done[(size_t)tile * E.P.n_internal + idx] != 1) E.order_violation = true;
            }
            const U4* base = E.sets.data() + ((size_t)tile * E.P.n_internal + idx) * 256;
            uint32_t G[16];
            load16(base, lane, G);
            uint32_t h0 = base[128 + lane].x;
            fold.add_set(G, h0 & ~G[0]);
        }
    }
    fold.finish(accG, accH);
}

void forward_item(Emu& E, int chunk, int tile) {
    const Chunk ck = E.P.chunks[chunk];
    const size_t T = E.T;
    auto row_of = [&](uint32_t ref) -> uint32_t {  // external refs name their ordinal in the chunk's dependency list
        uint32_t v = ref & REF_IDX_MASK;
        return (ref & REF_EXT) ? uint32_t(E.P.deps[ck.dep_begin + v]) : v;
    };
    static thread_local uint32_t acc[32][16], accH[32][16];
    memset(acc, 0, sizeof acc);
    memset(accH, 0, sizeof accH);
    // ---- chain segment, as in fitch_forward_kernel: bounds on the unknown input until every column is resolved
    const bool spec = E.algo == 0 && ck.chain_op >= 0 && !E.have_present;
    int first = ck.op_begin, resolved = -1;
    auto on_path = [&](const FwdOp& f, int op, int head) {
        for (int r = 0; r < f.n_refs; r++) {
            uint32_t ref = E.P.refs[f.ref_begin + r], kind = ref >> 30;
            if (kind == REF_CHAIN) return true;
            if (head < 0) continue;
            if (kind == REF_ACC && head == op - 1) return true;
            if (kind == REF_INT && !(ref & REF_EXT) && int(ref & REF_IDX_MASK) == head) return true;
        }
        return false;
    };
    auto fold_known = [&](const FwdOp& f, int op, int head, int lane, bool acc_regs, FitchFold& fold) {
        fold.reset();
        for (int r = 0; r < f.n_refs; r++) {
            uint32_t ref = E.P.refs[f.ref_begin + r], kind = ref >> 30, idx = ref & REF_IDX_MASK;
            if (kind == REF_CHAIN) continue;
            if (kind == REF_LEAF) {
                U4 c = E.leaf_planes[((size_t)tile * E.P.n_rows + idx) * 32 + lane];
                uint32_t cc[4] = {c.x, c.y, c.z, c.w};
                fold.add_leaf(cc, 0xFFFFFFFFu);
            } else if (kind == REF_ACC) {
                if (head >= 0 && head == op - 1) continue;
                if (acc_regs) fold.add_set(acc[lane]);
                else {
                    uint32_t S[16];
                    load16(E.sets.data() + ((size_t)tile * E.P.n_internal + (op - 1)) * 128, lane, S);
                    fold.add_set(S);
                }
            } else {
                if (!(ref & REF_EXT) && head >= 0 && int(idx) == head) continue;
                idx = row_of(ref);
                if ((ref & REF_EXT) && E.done[(size_t)tile * E.P.n_internal + idx] != 1) E.order_violation = true;
                uint32_t S[16];
                load16(E.sets.data() + ((size_t)tile * E.P.n_internal + idx) * 128, lane, S);
                fold.add_set(S);
            }
        }
    };
    auto root_ref = [&](int lane, uint32_t S[16]) {
        const U4* cp = E.colparams.data() + (size_t)tile * 128;
        U4 rc = cp[64 + lane];
        uint32_t rv = cp[96 + lane].y, r4[4] = {rc.x, rc.y, rc.z, rc.w}, d[16];
        decode16(r4, d);
        for (int k = 0; k < 16; k++) S[k] = (rv & d[k]) | (~rv & S[k]);
    };
    if (spec) {
        static thread_local FitchInterval iv[32];
        for (auto& x : iv) x.reset();
        int head = -1;
        first = ck.op_end;
        for (int op = ck.op_begin; op < ck.op_end; op++) {
            const FwdOp f = E.P.fwd_ops[op];
            const bool path = on_path(f, op, head);
            uint32_t open = 0;
            for (int lane = 0; lane < 32; lane++) {
                FitchFold fold;
                fold_known(f, op, head, lane, true, fold);
                if (!path) {
                    fold.finish(acc[lane]);
                    store16(E.sets.data() + ((size_t)tile * E.P.n_internal + op) * 128, lane, acc[lane]);
                    continue;
                }
                iv[lane].step(fold.A, fold.O);
                if ((f.flags & OPF_ROOT) && !(E.flags & 1)) { root_ref(lane, iv[lane].lo); root_ref(lane, iv[lane].hi); }
                open |= iv[lane].open();
            }
            if (!path) continue;
            head = op;
            if (!open) {
                for (int lane = 0; lane < 32; lane++) {
                    for (int k = 0; k < 16; k++) acc[lane][k] = iv[lane].lo[k];
                    store16(E.sets.data() + ((size_t)tile * E.P.n_internal + op) * 128, lane, acc[lane]);
                }
                if (f.flags & OPF_SIGNAL) E.done[(size_t)tile * E.P.n_internal + op] = 1;
                resolved = op;
                first = op + 1;
                break;
            }
        }
    }
    // ---- the same for Sankoff (sankoff_forward_kernel): bounds on the zero-excess set entering the segment
    bool spec_s = E.algo == 1 && ck.chain_op >= 0 && !E.have_present;
    auto known_g = [&](const FwdOp& f, int op, int head, int lane, bool acc_regs, uint32_t g[16]) {
        for (int r = 0; r < f.n_refs; r++) {
            uint32_t ref = E.P.refs[f.ref_begin + r], kind = ref >> 30, idx = ref & REF_IDX_MASK;
            if (kind == REF_CHAIN) continue;
            if (kind == REF_LEAF) {
                U4 c = E.leaf_planes[((size_t)tile * E.P.n_rows + idx) * 32 + lane];
                uint32_t cc[4] = {c.x, c.y, c.z, c.w};
                sankoff_leaf_g(cc, 0xFFFFFFFFu, g);
            } else if (kind == REF_ACC) {
                if (head >= 0 && head == op - 1) continue;
                if (acc_regs) for (int k = 0; k < 16; k++) g[k] = acc[lane][k];
                else load16(E.sets.data() + ((size_t)tile * E.P.n_internal + (op - 1)) * 256, lane, g);
            } else {
                if (!(ref & REF_EXT) && head >= 0 && int(idx) == head) continue;
                idx = row_of(ref);
                if ((ref & REF_EXT) && E.done[(size_t)tile * E.P.n_internal + idx] != 1) E.order_violation = true;
                load16(E.sets.data() + ((size_t)tile * E.P.n_internal + idx) * 256, lane, g);
            }
        }
    };
    if (spec_s) {
        static thread_local FitchInterval iv[32];
        for (auto& x : iv) x.reset();
        int head = -1;
        first = ck.op_end;
        for (int op = ck.op_begin; op < ck.op_end; op++) {
            const FwdOp f = E.P.fwd_ops[op];
            const bool path = on_path(f, op, head);
            if ((path && f.n_refs != 2) || (!path && f.max_arity_bits != 2)) { spec_s = false; break; }
            if (!path) {
                for (int lane = 0; lane < 32; lane++) {
                    SankoffFold<2> fold;
                    fold.reset();
                    for (int r = 0; r < f.n_refs; r++) {
                        uint32_t ref = E.P.refs[f.ref_begin + r], kind = ref >> 30, idx = ref & REF_IDX_MASK;
                        if (kind == REF_LEAF) {
                            U4 c = E.leaf_planes[((size_t)tile * E.P.n_rows + idx) * 32 + lane];
                            uint32_t cc[4] = {c.x, c.y, c.z, c.w};
                            fold.add_leaf(cc, 0xFFFFFFFFu);
                        } else if (kind == REF_ACC) {
                            fold.add_set(acc[lane], 0u);
                        } else {
                            idx = row_of(ref);
                            if ((ref & REF_EXT) && E.done[(size_t)tile * E.P.n_internal + idx] != 1) E.order_violation = true;
                            uint32_t G[16];
                            load16(E.sets.data() + ((size_t)tile * E.P.n_internal + idx) * 256, lane, G);
                            fold.add_set(G, 0u);
                        }
                    }
                    fold.finish(acc[lane], accH[lane]);
                    U4* base = E.sets.data() + ((size_t)tile * E.P.n_internal + op) * 256;
                    store16(base, lane, acc[lane]);
                    store16(base + 128, lane, accH[lane]);
                }
                continue;
            }
            uint32_t open = 0;
            for (int lane = 0; lane < 32; lane++) {
                uint32_t g[16];
                known_g(f, op, head, lane, true, g);
                for (int k = 0; k < 16; k++) g[k] = ~g[k];
                iv[lane].step(g, g);
                open |= iv[lane].open();
            }
            head = op;
            if (!open) {
                for (int lane = 0; lane < 32; lane++) {
                    for (int k = 0; k < 16; k++) { acc[lane][k] = ~iv[lane].lo[k]; accH[lane][k] = 0; }
                    U4* base = E.sets.data() + ((size_t)tile * E.P.n_internal + op) * 256;
                    store16(base, lane, acc[lane]);
                    store16(base + 128, lane, accH[lane]);
                }
                if (f.flags & OPF_SIGNAL) E.done[(size_t)tile * E.P.n_internal + op] = 1;
                resolved = op;
                first = op + 1;
                break;
            }
        }
        if (!spec_s) {
            first = ck.op_begin;
            memset(acc, 0, sizeof acc);
            memset(accH, 0, sizeof accH);
        }
    }
    for (int op = first; op < ck.op_end; op++) {
        const FwdOp f = E.P.fwd_ops[op];
        for (int lane = 0; lane < 32; lane++) {
            if (E.algo == 0) {
                const int type = E.have_present ? FT_GENERIC : ((f.flags >> OPF_TYPE_SHIFT) & 15);
                auto leafc = [&](int r, uint32_t cc[4]) {
                    uint32_t idx = E.P.refs[f.ref_begin + r] & REF_IDX_MASK;
                    U4 c = E.leaf_planes[((size_t)tile * E.P.n_rows + idx) * 32 + lane];
                    cc[0] = c.x; cc[1] = c.y; cc[2] = c.z; cc[3] = c.w;
                };
                auto intset = [&](int r, uint32_t X[16]) {
                    uint32_t ref = E.P.refs[f.ref_begin + r], idx = row_of(ref);
                    if ((ref & REF_EXT) && E.done[(size_t)tile * E.P.n_internal + idx] != 1) E.order_violation = true;
                    load16(E.sets.data() + ((size_t)tile * E.P.n_internal + idx) * 128, lane, X);
                };
                if (type == FT_LEAF_LEAF) {
                    uint32_t c0[4], c1[4];
                    leafc(0, c0); leafc(1, c1);
                    fitch_leaf_leaf(c0, c1, acc[lane]);
                } else if (type == FT_LEAF_ACC) {
                    uint32_t c0[4], X[16];
                    leafc(0, c0);
                    for (int k = 0; k < 16; k++) X[k] = acc[lane][k];
                    fitch_leaf_set(c0, X, acc[lane]);
                } else if (type == FT_LEAF_INT) {
                    uint32_t c0[4], X[16];
                    leafc(0, c0); intset(1, X);
                    fitch_leaf_set(c0, X, acc[lane]);
                } else if (type == FT_INT_ACC) {
                    uint32_t X[16], Y[16];
                    intset(0, X);
                    for (int k = 0; k < 16; k++) Y[k] = acc[lane][k];
                    fitch_set_set(X, Y, acc[lane]);
                } else {
                FitchFold fold;
                fold.reset();
                for (int r = 0; r < f.n_refs; r++) {
                    uint32_t ref = E.P.refs[f.ref_begin + r], kind = ref >> 30, idx = ref & REF_IDX_MASK;
                    if (kind == REF_LEAF) {
                        U4 c = E.leaf_planes[((size_t)tile * E.P.n_rows + idx) * 32 + lane];
                        uint32_t cc[4] = {c.x, c.y, c.z, c.w};
                        fold.add_leaf(cc, present_mask(E, idx));
                    } else if (kind == REF_ACC) {
                        fold.add_set(acc[lane]);
                    } else {
                        if (kind == REF_CHAIN) {
                            if (E.done[(size_t)tile * E.P.n_internal + idx] != 1) E.order_violation = true;
                        } else {
                            idx = row_of(ref);
                            if ((ref & REF_EXT) && E.done[(size_t)tile * E.P.n_internal + idx] != 1) E.order_violation = true;
                        }
                        uint32_t S[16];
                        load16(E.sets.data() + ((size_t)tile * E.P.n_internal + idx) * 128, lane, S);
                        fold.add_set(S);
                    }
                }
                fold.finish(acc[lane]);
                }
                if ((f.flags & OPF_ROOT) && !(E.flags & 1)) {
                    const U4* cp = E.colparams.data() + (size_t)tile * 128;
                    U4 rc = cp[64 + lane];
                    uint32_t rv = cp[96 + lane].y, r4[4] = {rc.x, rc.y, rc.z, rc.w}, d[16];
                    decode16(r4, d);
                    for (int k = 0; k < 16; k++) acc[lane][k] = (rv & d[k]) | (~rv & acc[lane][k]);
                }
                store16(E.sets.data() + ((size_t)tile * E.P.n_internal + op) * 128, lane, acc[lane]);
            } else {
                const int stype = (f.flags >> OPF_TYPE_SHIFT) & 15;
                if (stype != FT_GENERIC) {
                    uint32_t g1[16], g2[16], n1, n2;
                    auto leaf_child = [&](int r, uint32_t g[16], uint32_t& none) {
                        uint32_t idx = E.P.refs[f.ref_begin + r] & REF_IDX_MASK;
                        U4 c = E.leaf_planes[((size_t)tile * E.P.n_rows + idx) * 32 + lane];
                        uint32_t cc[4] = {c.x, c.y, c.z, c.w}, pr = present_mask(E, idx);
                        if ((E.flags & 1) && !pr) { cc[0] = cc[1] = cc[2] = cc[3] = 0; pr = 0xFFFFFFFFu; }
                        sankoff_leaf_g(cc, pr, g);
                        none = ~pr;
                    };
                    auto set_child = [&](int r, uint32_t g[16], uint32_t& none) {
                        uint32_t ref = E.P.refs[f.ref_begin + r], idx = row_of(ref);
                        if ((ref & REF_EXT) && E.done[(size_t)tile * E.P.n_internal + idx] != 1) E.order_violation = true;
                        const U4* base = E.sets.data() + ((size_t)tile * E.P.n_internal + idx) * 256;
                        load16(base, lane, g);
                        none = base[128 + lane].x & ~g[0];
                    };
                    auto acc_child = [&](uint32_t g[16], uint32_t& none) {
                        for (int k = 0; k < 16; k++) g[k] = acc[lane][k];
                        none = sankoff_none(acc[lane], accH[lane]);
                    };
                    if (stype == FT_LEAF_LEAF) { leaf_child(0, g1, n1); leaf_child(1, g2, n2); }
                    else if (stype == FT_LEAF_ACC) { leaf_child(0, g1, n1); acc_child(g2, n2); }
                    else if (stype == FT_LEAF_INT) { leaf_child(0, g1, n1); set_child(1, g2, n2); }
                    else { set_child(0, g1, n1); acc_child(g2, n2); }
                    sankoff_pair(g1, n1, g2, n2, acc[lane], accH[lane]);
                } else
                switch (f.max_arity_bits) {
                case 2: sankoff_fwd_op<2>(E, ck, f, tile, lane, acc[lane], accH[lane]); break;
                case 4: sankoff_fwd_op<4>(E, ck, f, tile, lane, acc[lane], accH[lane]); break;
                case 8: sankoff_fwd_op<8>(E, ck, f, tile, lane, acc[lane], accH[lane]); break;
                default: sankoff_fwd_op<20>(E, ck, f, tile, lane, acc[lane], accH[lane]); break;
                }
                U4* base = E.sets.data() + ((size_t)tile * E.P.n_internal + op) * 256;
                store16(base, lane, acc[lane]);
                store16(base + 128, lane, accH[lane]);
            }
        }
        if (f.flags & OPF_SIGNAL) E.done[(size_t)tile * E.P.n_internal + op] = 1;
    }
    if (spec) {  // redo the path ops before the resolved one with the real input
        if (E.done[(size_t)tile * E.P.n_internal + ck.chain_row] != 1) E.order_violation = true;
        static thread_local uint32_t S[32][16];
        for (int lane = 0; lane < 32; lane++) load16(E.sets.data() + ((size_t)tile * E.P.n_internal + ck.chain_row) * 128, lane, S[lane]);
        const int end = resolved >= 0 ? resolved : ck.op_end;
        int head = -1;
        for (int op = ck.chain_op; op < end; op++) {
            const FwdOp f = E.P.fwd_ops[op];
            if (!on_path(f, op, head)) continue;
            for (int lane = 0; lane < 32; lane++) {
                FitchFold fold;
                fold_known(f, op, head, lane, false, fold);
                fold.add_set(S[lane]);
                fold.finish(S[lane]);
                if ((f.flags & OPF_ROOT) && !(E.flags & 1)) root_ref(lane, S[lane]);
                store16(E.sets.data() + ((size_t)tile * E.P.n_internal + op) * 128, lane, S[lane]);
            }
            if (f.flags & OPF_SIGNAL) E.done[(size_t)tile * E.P.n_internal + op] = 1;
            head = op;
        }
        if (resolved >= 0) {  // the speculation must have produced what the exact evaluation gives
            const FwdOp f = E.P.fwd_ops[resolved];
            for (int lane = 0; lane < 32; lane++) {
                FitchFold fold;
                fold_known(f, resolved, head, lane, false, fold);
                fold.add_set(S[lane]);
                uint32_t X[16], Y[16];
                fold.finish(X);
                if ((f.flags & OPF_ROOT) && !(E.flags & 1)) root_ref(lane, X);
                load16(E.sets.data() + ((size_t)tile * E.P.n_internal + resolved) * 128, lane, Y);
                for (int k = 0; k < 16; k++)
                    if (X[k] != Y[k]) E.spec_mismatch = true;
            }
        }
    }
    if (spec_s) {  // redo the path ops up to and including the resolved one with the real input
        if (E.done[(size_t)tile * E.P.n_internal + ck.chain_row] != 1) E.order_violation = true;
        static thread_local uint32_t Gc[32][16];
        for (int lane = 0; lane < 32; lane++) load16(E.sets.data() + ((size_t)tile * E.P.n_internal + ck.chain_row) * 256, lane, Gc[lane]);
        const int end = resolved >= 0 ? resolved + 1 : ck.op_end;
        int head = -1;
        for (int op = ck.chain_op; op < end; op++) {
            const FwdOp f = E.P.fwd_ops[op];
            if (!on_path(f, op, head)) continue;
            for (int lane = 0; lane < 32; lane++) {
                uint32_t g[16], G[16], H[16];
                known_g(f, op, head, lane, false, g);
                sankoff_pair(g, 0u, Gc[lane], 0u, G, H);
                U4* base = E.sets.data() + ((size_t)tile * E.P.n_internal + op) * 256;
                if (op == resolved) {  // the speculated G must be what the exact evaluation gives
                    uint32_t Y[16];
                    load16(base, lane, Y);
                    for (int k = 0; k < 16; k++)
                        if (Y[k] != G[k]) E.spec_mismatch = true;
                }
                store16(base, lane, G);
                store16(base + 128, lane, H);
                for (int k = 0; k < 16; k++) Gc[lane][k] = G[k];
            }
            if (f.flags & OPF_SIGNAL) E.done[(size_t)tile * E.P.n_internal + op] = 1;
            head = op;
        }
    }
}

void backward_item(Emu& E, int chunk, int tile) {
    const Chunk ck = E.P.chunks[chunk];
    const size_t T = E.T;
    uint32_t accF[32][4] = {}, accVis[32] = {};
    static thread_local uint32_t stack[BWD_STACK_DEPTH][160];
    for (auto& e : stack) for (auto& w : e) w = 0xDEADBEEFu;
    const int J = E.algo == 0 ? 128 : 256;
    const int last = ck.op_end - 1;
    int resolved = -1;
    uint32_t specF[32][4] = {}, specVis[32] = {};
    if (!E.have_present && (ck.flags & CHUNK_CHAIN_TOP)) {
        E.spec_items++;
        static thread_local uint32_t Q[32][16];
        for (auto& q : Q) for (auto& w : q) w = 0xFFFFFFFFu;
        for (int op = last; op >= ck.op_begin; op--) {
            if (op != last && !(E.P.bwd_ops[op].flags & OPF_HEAVY)) continue;
            uint32_t open = 0;
            for (int lane = 0; lane < 32; lane++) {
                uint32_t S[16], H[16];
                load16(E.sets.data() + ((size_t)tile * E.P.n_internal + op) * J, lane, S);
                if (E.algo == 0) {
                    fitch_candidates_step(Q[lane], S);
                } else {
                    load16(E.sets.data() + ((size_t)tile * E.P.n_internal + op) * J + 128, lane, H);
                    sankoff_candidates_step(Q[lane], S, H);
                }
                open |= candidates_open(Q[lane]);
            }
            if (!open) {
                resolved = op;
                E.spec_resolved++;
                for (int lane = 0; lane < 32; lane++) {
                    encode16(Q[lane], specF[lane]);
                    specVis[lane] = E.colparams[(size_t)tile * 128 + 96 + lane].z;
                    for (int k = 0; k < 4; k++) specF[lane][k] &= specVis[lane];
                }
                break;
            }
        }
    }
    for (int range = (resolved >= 0 ? 0 : 1); range < 2; range++) {
    const int hi = range == 0 ? resolved : last;
    const int lo = (range == 1 && resolved >= 0) ? resolved : ck.op_begin;
    for (int op = hi; op >= lo; op--) {
        const BwdOp b = E.P.bwd_ops[op];
        const bool given = range == 0 && op == resolved, own_only = range == 1 && op == resolved;
        WarpMut wm;
        uint32_t Fw[32][4], visw[32];
        for (int lane = 0; lane < 32; lane++) {
            uint32_t G[16], H[16];
            const U4* base = E.sets.data() + ((size_t)tile * E.P.n_internal + op) * J;
            load16(base, lane, G);
            if (E.algo == 1) load16(base + 128, lane, H);
            uint32_t P[4] = {0, 0, 0, 0}, F[4], vis;
            if (given) {
                for (int k = 0; k < 4; k++) F[k] = specF[lane][k];
                vis = specVis[lane];
            } else if (b.parent_ref == PARENT_ROOT) {
                const U4* cp = E.colparams.data() + (size_t)tile * 128;
                U4 pc = cp[lane], ov = cp[32 + lane], fl = cp[96 + lane];
                P[0] = pc.x; P[1] = pc.y; P[2] = pc.z; P[3] = pc.w;
                uint32_t o4[4] = {ov.x, ov.y, ov.z, ov.w};
                const uint32_t ov_valid = fl.x & fl.z, colmask = fl.z;
                if (E.algo == 0) {
                    if (E.flags & 1) {
                        uint32_t v0;
                        fitch_assign(G, P, colmask, F, v0);
                        vis = (v0 | ov_valid) & colmask;
                        for (int k = 0; k < 4; k++) F[k] = ((ov_valid & o4[k]) | (~ov_valid & F[k])) & vis;
                    } else {
                        fitch_assign_root(G, o4, ov_valid, colmask, F, vis);
                    }
                } else {
                    uint32_t undefined;
                    sankoff_assign_root(G, H, o4, ov_valid, colmask, F, vis, undefined);
                    if (undefined && !(E.flags & 1)) E.error |= 1;
                }
            } else {
                uint32_t pvis;
                if (b.parent_ref == PARENT_ACC) {
                    for (int k = 0; k < 4; k++) P[k] = accF[lane][k];
                    pvis = accVis[lane];
                } else if (b.parent_ref <= PARENT_STACK0) {
                    const uint32_t* e = stack[PARENT_STACK0 - b.parent_ref];
                    for (int k = 0; k < 4; k++) P[k] = e[4 * lane + k];
                    pvis = e[128 + lane];
                } else {
                    if ((b.flags & OPF_PARENT_EXT) && E.fdone[(size_t)tile * std::max(1, E.P.n_fslots) + b.parent_ref] != 1) E.order_violation = true;
                    const uint32_t* fs = E.fstore.data() + ((size_t)tile * std::max(1, E.P.n_fslots) + b.parent_ref) * 160;
                    U4 a = reinterpret_cast<const U4*>(fs)[lane];
                    pvis = fs[128 + lane];
                    P[0] = a.x; P[1] = a.y; P[2] = a.z; P[3] = a.w;
                }
                if (E.algo == 0) fitch_assign(G, P, pvis, F, vis);
                else sankoff_assign(G, H, P, pvis, F, vis);
            }
            wm.mut[lane] = vis & differs4(F, P);
            for (int k = 0; k < 4; k++) { wm.P[lane][k] = P[k]; wm.F[lane][k] = F[k]; Fw[lane][k] = F[k]; }
            visw[lane] = vis;
            if (own_only) {
                if (F[0] != specF[lane][0] || F[1] != specF[lane][1] || F[2] != specF[lane][2] || F[3] != specF[lane][3] ||
                    vis != specVis[lane])
                    E.spec_mismatch = true;
                continue;
            }
            if (b.flags & OPF_PUSH) {
                uint32_t* e = stack[(b.flags >> OPF_PUSH_SHIFT) & 15];
                for (int k = 0; k < 4; k++) e[4 * lane + k] = F[k];
                e[128 + lane] = vis;
            }
            if (b.fslot_out >= 0) {
                uint32_t* fs = E.fstore.data() + ((size_t)tile * std::max(1, E.P.n_fslots) + b.fslot_out) * 160;
                reinterpret_cast<U4*>(fs)[lane] = U4{F[0], F[1], F[2], F[3]};
                fs[128 + lane] = vis;
            }
            store_state(E, b.node, tile, lane, F, vis);
        }
        if (!given) emit(E, b.node, tile, wm);
        if (own_only) continue;
        if (b.fslot_out >= 0 && (b.flags & OPF_SIGNAL_F)) E.fdone[(size_t)tile * std::max(1, E.P.n_fslots) + b.fslot_out] = 1;
        for (int l = 0; l < b.n_leaves; l++) {
            const BwdLeaf lf = E.P.bwd_leaves[b.leaf_begin + l];
            for (int lane = 0; lane < 32; lane++) {
                U4 c = E.leaf_planes[((size_t)tile * E.P.n_rows + lf.row) * 32 + lane];
                uint32_t cc[4] = {c.x, c.y, c.z, c.w}, pr = present_mask(E, lf.row);
                if (E.algo == 1 && (E.flags & 1) && !pr) { cc[0] = cc[1] = cc[2] = cc[3] = 0; pr = 0xFFFFFFFFu; }
                uint32_t lvis = visw[lane] & pr;
                wm.mut[lane] = lvis & differs4(cc, Fw[lane]);
                for (int k = 0; k < 4; k++) { wm.P[lane][k] = Fw[lane][k]; wm.F[lane][k] = cc[k]; }
                uint32_t m4[4] = {cc[0] & lvis, cc[1] & lvis, cc[2] & lvis, cc[3] & lvis};
                store_state(E, lf.node, tile, lane, m4, lvis);
            }
            emit(E, lf.node, tile, wm);
        }
        for (int lane = 0; lane < 32; lane++) {
            for (int k = 0; k < 4; k++) accF[lane][k] = Fw[lane][k];
            accVis[lane] = visw[lane];
        }
    }
    }
}

}  // namespace

extern "C" {

// Same inputs as pmb_run_nuc, except leaf_codes is one code per byte (n_rows x n_cols).
// Outputs: node_offsets (n_nodes+1), pos/type_code sized by the caller via a first call with pos == NULL
// (returns n_mut), states optional (n_nodes x n_cols). Returns n_mut >= 0, or a negative PMB_ERR_* code.
long long emul_run(int algo, int block_mode, int n_nodes, int root, const int32_t* child_off, const int32_t* child_idx,
                   const int32_t* leaf_row, int chunk_nodes, long long n_cols, const uint8_t* leaf_codes,
                   const uint8_t* leaf_present, const uint8_t* parent_code, const int8_t* root_override,
                   const int8_t* fwd_root_ref, long long col_base, long long* node_offsets, int32_t* pos,
                   uint8_t* type_code, uint8_t* states_out,
                   int32_t* prog_stats /* 8: chunks, levels, fslots, max_arity, chain segments, speculated backward items, of which resolved, 0 */,
                   int inline_nodes, int level_mode) {
    Emu E;
    std::string err = build_tree_program(n_nodes, root, child_off, child_idx, leaf_row, chunk_nodes, inline_nodes, &E.P);
    if (!err.empty()) return -1;
    const TreeProgram& P = E.P;
    if (prog_stats) {
        prog_stats[0] = int32_t(P.chunks.size());
        prog_stats[1] = P.n_levels();
        prog_stats[2] = P.n_fslots;
        prog_stats[3] = P.max_arity;
    }
    E.algo = algo;
    E.flags = block_mode ? 1 : 0;
    E.T = int((n_cols + TILE_COLS - 1) / TILE_COLS);
    const size_t T = E.T;
    // pack_leaves_kernel
    E.leaf_planes.assign((size_t)P.n_rows * T * 32, U4{0, 0, 0, 0});
    for (int r = 0; r < P.n_rows; r++)
        for (long long c = 0; c < n_cols; c++) {
            uint32_t code = leaf_codes[(size_t)r * n_cols + c] & 15u;
            U4& u = E.leaf_planes[((size_t)(c / TILE_COLS) * P.n_rows + P.row_slot[r]) * 32 + ((c % TILE_COLS) >> 5)];
            uint32_t bit = 1u << (c & 31);
            if (code & 1) u.x |= bit;
            if (code & 2) u.y |= bit;
            if (code & 4) u.z |= bit;
            if (code & 8) u.w |= bit;
        }
    E.have_present = leaf_present != nullptr;
    if (leaf_present) {  // indexed by leaf slot, like the kernels
        E.present.assign(P.n_rows, 0);
        for (int r = 0; r < P.n_rows; r++) E.present[P.row_slot[r]] = leaf_present[r];
    }
    // pack_colparams_kernel
    E.colparams.assign(T * 128, U4{0, 0, 0, 0});
    for (long long c = 0; c < n_cols; c++) {
        size_t tile = size_t(c / TILE_COLS);
        int lane = int((c % TILE_COLS) >> 5);
        uint32_t bit = 1u << (c & 31);
        U4* cp = E.colparams.data() + tile * 128;
        auto setcode = [&](U4& u, int code) {
            if (code & 1) u.x |= bit;
            if (code & 2) u.y |= bit;
            if (code & 4) u.z |= bit;
            if (code & 8) u.w |= bit;
        };
        setcode(cp[lane], parent_code[c] & 15);
        if (root_override && root_override[c] >= 0) { setcode(cp[32 + lane], root_override[c]); cp[96 + lane].x |= bit; }
        if (fwd_root_ref && fwd_root_ref[c] >= 0) { setcode(cp[64 + lane], fwd_root_ref[c]); cp[96 + lane].y |= bit; }
        cp[96 + lane].z |= bit;
    }
    E.sets.assign((size_t)P.n_internal * T * (algo == 0 ? 128 : 256), U4{0xDEADBEEFu, 0xDEADBEEFu, 0xDEADBEEFu, 0xDEADBEEFu});
    E.fstore.assign((size_t)std::max(1, P.n_fslots) * T * 160, 0xDEADBEEFu);
    if (states_out) E.states.assign((size_t)n_nodes * T * 64, U4{0xDEADBEEFu, 0xDEADBEEFu, 0xDEADBEEFu, 0xDEADBEEFu});
    E.dir.assign((size_t)n_nodes * T, 0ull);
    E.done.assign((size_t)P.n_internal * T, 0u);
    E.fdone.assign((size_t)std::max(1, P.n_fslots) * T, 0u);
    // persistent-kernel ticket order: chunks in schedule order, tiles fastest (backward: chunks reversed). An item
    // that finds a dependency flag unset here would have to wait for a LARGER ticket on the GPU: a scheduling bug.
    const int NC = int(P.chunks.size());
    for (int ch = 0; ch < NC; ch++)
        for (int t = 0; t < E.T; t++) forward_item(E, ch, t);
    for (int k = 0; k < NC; k++)
        for (int t = 0; t < E.T; t++) backward_item(E, P.bwd_order[k], t);
    if (level_mode) {  // the level-major order must be a valid schedule too (one launch per level)
        std::fill(E.done.begin(), E.done.end(), 0u);
        std::fill(E.fdone.begin(), E.fdone.end(), 0u);
        E.pool = 0;
        E.staging.clear();
        std::fill(E.dir.begin(), E.dir.end(), 0ull);
        const int L = P.n_levels();
        for (int l = 0; l < L; l++)
            for (int k = P.level_chunk_begin[l + 1] - 1; k >= P.level_chunk_begin[l]; k--)
                for (int t = E.T - 1; t >= 0; t--) forward_item(E, P.level_order[k], t);
        for (int l = L - 1; l >= 0; l--)
            for (int k = P.level_chunk_begin[l + 1] - 1; k >= P.level_chunk_begin[l]; k--)
                for (int t = E.T - 1; t >= 0; t--) backward_item(E, P.level_order[k], t);
    }
    if (prog_stats) {
        prog_stats[4] = 0;
        for (const Chunk& ck : P.chunks) prog_stats[4] += ck.chain_op >= 0;
        prog_stats[5] = int32_t(E.spec_items);
        prog_stats[6] = int32_t(E.spec_resolved);
        prog_stats[7] = 0;
    }
    if (E.order_violation) return -7;
    if (E.spec_mismatch) return -8;
    if (E.error & 1) return -4;
    // node_count + scan + gather
    long long run = 0;
    for (int v = 0; v < n_nodes; v++) {
        node_offsets[v] = run;
        for (size_t t = 0; t < T; t++) {
            unsigned long long d = E.dir[(size_t)v * T + t];
            int n = int(d & 0x7FFull);
            if (pos)
                for (int i = 0; i < n; i++) {
                    uint32_t r = E.staging[(d >> 11) + i];
                    pos[run + i] = int32_t(col_base + (long long)t * TILE_COLS + (r & 1023u));
                    type_code[run + i] = uint8_t(((r >> 14) << 4) | ((r >> 10) & 15u));
                }
            run += n;
        }
    }
    node_offsets[n_nodes] = run;
    if (states_out)
        for (int v = 0; v < n_nodes; v++)
            for (long long c = 0; c < n_cols; c++) {
                size_t tile = size_t(c / TILE_COLS);
                int lane = int((c % TILE_COLS) >> 5), bit = int(c & 31);
                const U4* s = E.states.data() + ((size_t)v * T + tile) * 64;
                U4 f = s[lane];
                uint32_t vis = s[32 + lane].x;
                uint32_t code = ((f.x >> bit) & 1u) | (((f.y >> bit) & 1u) << 1) | (((f.z >> bit) & 1u) << 2) | (((f.w >> bit) & 1u) << 3);
                states_out[(size_t)v * n_cols + c] = ((vis >> bit) & 1u) ? uint8_t(code) : uint8_t(0xFF);
            }
    return run;
}

}  // extern "C"

// schedule statistics of the tree program (no data involved): out[0] chunks, [1] levels, [2] fslots, [3] max arity,
// [4] largest chunk (ops), [5] ops in the root's chunk, [6] critical path in ops (longest dependency chain of
// chunks, counting every op of each chunk on it), [7] number of chunks with <= 2 ops
extern "C" int emul_prog_stats(int n_nodes, int root, const int32_t* child_off, const int32_t* child_idx, const int32_t* leaf_row,
                               int chunk_nodes, int inline_nodes, long long* out) {
    TreeProgram P;
    std::string err = build_tree_program(n_nodes, root, child_off, child_idx, leaf_row, chunk_nodes, inline_nodes, &P);
    if (!err.empty()) return -1;
    const int NC = int(P.chunks.size());
    std::vector<int> chunk_of_op(P.n_internal);
    for (int c = 0; c < NC; c++)
        for (int op = P.chunks[c].op_begin; op < P.chunks[c].op_end; op++) chunk_of_op[op] = c;
    std::vector<long long> cp(NC, 0);
    long long best = 0, largest = 0, tiny = 0;
    for (int c = 0; c < NC; c++) {  // schedule order is topological
        long long dep = 0;
        for (int op = P.chunks[c].op_begin; op < P.chunks[c].op_end; op++) {
            const FwdOp& f = P.fwd_ops[op];
            for (int r = 0; r < f.n_refs; r++) {
                uint32_t ref = P.refs[f.ref_begin + r];
                if ((ref >> 30) == REF_INT && (ref & REF_EXT))
                    dep = std::max(dep, cp[chunk_of_op[P.deps[P.chunks[c].dep_begin + (ref & REF_IDX_MASK)]]]);
            }
        }
        long long n = P.chunks[c].op_end - P.chunks[c].op_begin;
        cp[c] = dep + n;
        best = std::max(best, cp[c]);
        largest = std::max(largest, n);
        tiny += n <= 2;
    }
    out[0] = NC; out[1] = P.n_levels(); out[2] = P.n_fslots; out[3] = P.max_arity; out[4] = largest;
    out[5] = P.chunks[chunk_of_op[P.node_op[root]]].op_end - P.chunks[chunk_of_op[P.node_op[root]]].op_begin;
    out[6] = best; out[7] = tiny;
    return 0;
}

// Soundness of the speculation bounds (plane_math.h FitchInterval / fitch_candidates_step), column by column against
// plain integer arithmetic on random inputs: the true value always lies within the bounds, and wherever the bounds
// say "known" they are the true value. Returns 0, or the 1-based number of the first failing trial.
extern "C" int emul_speculation_selftest(unsigned long long seed, int trials) {
    auto rnd = [&]() {
        seed += 0x9E3779B97F4A7C15ull;
        unsigned long long z = seed;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    };
    auto rand_set = [&](int style) -> unsigned {  // non-empty 16-bit set
        unsigned s;
        if (style == 0) s = 1u << (rnd() % 4);                       // one of four one-hot states (alignment-like)
        else if (style == 1) s = 1u << (rnd() % 16);
        else if (style == 2) s = unsigned(rnd() & rnd() & 0xFFFFu);  // sparse
        else s = unsigned(rnd() & 0xFFFFu);
        return s ? s : 1u;
    };
    for (int t = 0; t < trials; t++) {
        const int steps = 1 + int(rnd() % 12);
        // ---- forward: 32 independent columns per trial
        unsigned S[32];
        FitchInterval iv;
        iv.reset();
        const int style = int(rnd() % 4);
        for (int j = 0; j < 32; j++) S[j] = rand_set(int(rnd() % 4));
        for (int s = 0; s < steps; s++) {
            unsigned A[32], O[32];
            uint32_t Ap[16] = {}, Op[16] = {};
            for (int j = 0; j < 32; j++) {
                const int nk = int(rnd() % 3);  // 0 known children: a unary node
                A[j] = 0xFFFFu;
                O[j] = 0;
                for (int c = 0; c < nk; c++) {
                    unsigned x = rand_set(style);
                    A[j] &= x;
                    O[j] |= x;
                }
                for (int k = 0; k < 16; k++) {
                    if ((A[j] >> k) & 1) Ap[k] |= 1u << j;
                    if ((O[j] >> k) & 1) Op[k] |= 1u << j;
                }
                S[j] = (S[j] & A[j]) ? (S[j] & A[j]) : (S[j] | O[j]);
            }
            iv.step(Ap, Op);
            const uint32_t open = iv.open();
            for (int j = 0; j < 32; j++) {
                unsigned lo = 0, hi = 0;
                for (int k = 0; k < 16; k++) {
                    lo |= ((iv.lo[k] >> j) & 1u) << k;
                    hi |= ((iv.hi[k] >> j) & 1u) << k;
                }
                if ((lo & ~S[j]) || (S[j] & ~hi)) return t + 1;
                if (!((open >> j) & 1u) && lo != S[j]) return t + 1;
            }
        }
        // ---- backward
        unsigned P[32];
        uint32_t Q[16];
        for (int k = 0; k < 16; k++) Q[k] = 0xFFFFFFFFu;
        for (int j = 0; j < 32; j++) P[j] = unsigned(rnd() % 16);
        for (int s = 0; s < steps; s++) {
            uint32_t Sp[16] = {};
            for (int j = 0; j < 32; j++) {
                unsigned x = rand_set(style);
                for (int k = 0; k < 16; k++)
                    if ((x >> k) & 1) Sp[k] |= 1u << j;
                if (!((x >> P[j]) & 1u)) P[j] = unsigned(__builtin_ctz(x));
            }
            fitch_candidates_step(Q, Sp);
            const uint32_t open = candidates_open(Q);
            uint32_t code[4];
            encode16(Q, code);
            for (int j = 0; j < 32; j++) {
                if (!((Q[P[j]] >> j) & 1u)) return t + 1;
                if (!((open >> j) & 1u)) {
                    unsigned c = ((code[0] >> j) & 1u) | (((code[1] >> j) & 1u) << 1) | (((code[2] >> j) & 1u) << 2) | (((code[3] >> j) & 1u) << 3);
                    if (c != P[j]) return t + 1;
                }
            }
        }
        // ---- backward, Sankoff: random excess vectors (at least one zero-excess state per column)
        {
            unsigned Ps[32];
            uint32_t Qs[16];
            for (int k = 0; k < 16; k++) Qs[k] = 0xFFFFFFFFu;
            for (int j = 0; j < 32; j++) Ps[j] = unsigned(rnd() % 16);
            for (int st = 0; st < steps; st++) {
                uint32_t Gp[16] = {}, Hp[16] = {};
                for (int j = 0; j < 32; j++) {
                    int e[16], z = -1;
                    const int zero_at = int(rnd() % 16);
                    for (int k = 0; k < 16; k++) {
                        e[k] = style == 0 ? 2 : int(rnd() % 3);
                        if (k == zero_at) e[k] = 0;
                        if (e[k] == 0 && z < 0) z = k;
                        if (e[k] > 0) Gp[k] |= 1u << j;
                        if (e[k] > 1) Hp[k] |= 1u << j;
                    }
                    const int sst = int(Ps[j]);
                    Ps[j] = unsigned(e[sst] == 0 ? sst : (e[sst] == 1 ? (sst < z ? sst : z) : z));
                }
                sankoff_candidates_step(Qs, Gp, Hp);
                const uint32_t open = candidates_open(Qs);
                uint32_t code[4];
                encode16(Qs, code);
                for (int j = 0; j < 32; j++) {
                    if (!((Qs[Ps[j]] >> j) & 1u)) return t + 1;
                    if (!((open >> j) & 1u)) {
                        unsigned c = ((code[0] >> j) & 1u) | (((code[1] >> j) & 1u) << 1) | (((code[2] >> j) & 1u) << 2) | (((code[3] >> j) & 1u) << 3);
                        if (c != Ps[j]) return t + 1;
                    }
                }
            }
        }
    }
    return 0;
}

// Structural invariants of the tree program (no data involved). Returns 0, or a positive code naming the first violated
// invariant (see the numbered comments). Used by hypothesis tests over random trees and parameters.
extern "C" int emul_prog_check(int n_nodes, int root, const int32_t* child_off, const int32_t* child_idx, const int32_t* leaf_row,
                               int chunk_nodes, int inline_nodes, int bwd_tail_chunks) {
    TreeProgram P;
    std::string err = build_tree_program(n_nodes, root, child_off, child_idx, leaf_row, chunk_nodes, inline_nodes, &P, bwd_tail_chunks);
    if (!err.empty()) return -1;
    const int NC = int(P.chunks.size()), NI = P.n_internal;
    std::vector<int> parent(n_nodes, -1);
    for (int v = 0; v < n_nodes; v++)
        for (int e = child_off[v]; e < child_off[v + 1]; e++) parent[child_idx[e]] = v;
    // 1. the chunks partition the ops, in order
    std::vector<int> chunk_of_op(NI, -1);
    int expect = 0;
    for (int c = 0; c < NC; c++) {
        if (P.chunks[c].op_begin != expect || P.chunks[c].op_end <= P.chunks[c].op_begin) return 1;
        for (int op = P.chunks[c].op_begin; op < P.chunks[c].op_end; op++) chunk_of_op[op] = c;
        expect = P.chunks[c].op_end;
    }
    if (expect != NI) return 1;
    // 2. node <-> op is a bijection on internal nodes; leaf rows <-> slots a permutation
    std::vector<int> op_node(NI, -1);
    for (int v = 0; v < n_nodes; v++) {
        const bool internal = child_off[v] != child_off[v + 1];
        if (internal != (P.node_op[v] >= 0)) return 2;
        if (internal) {
            if (op_node[P.node_op[v]] != -1) return 2;
            op_node[P.node_op[v]] = v;
        }
    }
    {
        std::vector<char> seen(P.n_rows, 0);
        for (int r = 0; r < P.n_rows; r++) {
            if (P.row_slot[r] < 0 || P.row_slot[r] >= P.n_rows || seen[P.row_slot[r]]) return 2;
            seen[P.row_slot[r]] = 1;
        }
    }
    std::vector<int> level_of(NC, -1), bwd_pos(NC, -1);
    for (int l = 0; l < P.n_levels(); l++)
        for (int k = P.level_chunk_begin[l]; k < P.level_chunk_begin[l + 1]; k++) level_of[P.level_order[k]] = l;
    for (int k = 0; k < NC; k++) {
        if (P.bwd_order[k] < 0 || P.bwd_order[k] >= NC || bwd_pos[P.bwd_order[k]] != -1) return 3;
        bwd_pos[P.bwd_order[k]] = k;
    }
    std::vector<char> has_child_chunk(NC, 0);
    for (int op = 0; op < NI; op++) {
        const FwdOp& f = P.fwd_ops[op];
        const BwdOp& b = P.bwd_ops[op];
        const int c = chunk_of_op[op], v = op_node[op];
        if (b.node != v) return 4;
        // 4. the refs name exactly the node's children
        if (f.n_refs != child_off[v + 1] - child_off[v]) return 4;
        int n_chain = 0;
        for (int r = 0; r < f.n_refs; r++) {
            const uint32_t ref = P.refs[f.ref_begin + r], kind = ref >> 30, idx = ref & REF_IDX_MASK;
            int child_op = -1;
            if (kind == REF_LEAF) continue;
            if (kind == REF_ACC) {
                if (op == P.chunks[c].op_begin) return 5;  // 5. the accumulator is the previous op of the same chunk
                child_op = op - 1;
            } else if (kind == REF_CHAIN) {
                n_chain++;
                child_op = int(idx);
                // 8. a chain child is the top op of an earlier chunk flagged as a chain segment with a segment above
                if (P.chunks[c].chain_op != op || P.chunks[c].chain_row != child_op) return 8;
                const int cc = chunk_of_op[child_op];
                if (cc >= c || child_op != P.chunks[cc].op_end - 1 || !(P.chunks[cc].flags & CHUNK_CHAIN_TOP)) return 8;
                if (!(P.bwd_ops[child_op].flags & OPF_CHAIN_TOP) || !(P.bwd_ops[child_op].flags & OPF_PARENT_EXT)) return 8;
                if (level_of[cc] >= level_of[c] || bwd_pos[cc] <= bwd_pos[c]) return 8;
                has_child_chunk[c] = 1;
            } else if (ref & REF_EXT) {
                if (int(idx) >= P.chunks[c].dep_count) return 6;
                child_op = P.deps[P.chunks[c].dep_begin + idx];
                const int cc = chunk_of_op[child_op];
                // 6. an external row comes from an earlier ticket, a lower level, and is published (chunk root)
                if (cc >= c || level_of[cc] >= level_of[c] || bwd_pos[cc] <= bwd_pos[c]) return 6;
                if (!(P.fwd_ops[child_op].flags & OPF_SIGNAL)) return 6;
                has_child_chunk[c] = 1;
            } else {
                child_op = int(idx);
                if (chunk_of_op[child_op] != c || child_op >= op) return 5;
            }
            if (parent[op_node[child_op]] != v) return 4;
        }
        if (n_chain > 1 || (n_chain == 1) != (P.chunks[c].chain_op == op)) return 8;
        if (((f.flags & OPF_ROOT) != 0) != (v == root)) return 4;
    }
    // 7. backward: replay every chunk and check that each op receives ITS parent's state
    for (int c = 0; c < NC; c++) {
        const Chunk& ck = P.chunks[c];
        int stack_owner[16];
        for (int& s : stack_owner) s = -1;
        for (int op = ck.op_end - 1; op >= ck.op_begin; op--) {
            const BwdOp& b = P.bwd_ops[op];
            const int v = b.node;
            if (v == root) {
                if (b.parent_ref != PARENT_ROOT) return 7;
            } else {
                const int pop = P.node_op[parent[v]];
                if (b.parent_ref == PARENT_ACC) {
                    if (pop != op + 1 || chunk_of_op[pop] != c) return 7;
                } else if (b.parent_ref <= PARENT_STACK0) {
                    const int e = PARENT_STACK0 - b.parent_ref;
                    if (e >= BWD_STACK_DEPTH || stack_owner[e] != pop) return 7;
                } else if (b.parent_ref >= 0) {
                    if (P.bwd_ops[pop].fslot_out != b.parent_ref) return 7;
                    const bool ext = chunk_of_op[pop] != c;
                    if (ext != ((b.flags & OPF_PARENT_EXT) != 0)) return 7;
                    if (ext && (!(P.bwd_ops[pop].flags & OPF_SIGNAL_F) || bwd_pos[chunk_of_op[pop]] >= bwd_pos[c])) return 7;
                    if (!ext && pop <= op) return 7;
                } else {
                    return 7;
                }
            }
            if (b.flags & OPF_PUSH) stack_owner[(b.flags >> OPF_PUSH_SHIFT) & 15] = op;
        }
        // 9. the heavy path flags: the chunk's last op is its root and carries OPF_HEAVY
        if (!(P.bwd_ops[ck.op_end - 1].flags & OPF_HEAVY)) return 9;
    }
    // (the sorted backward tail needs no check of its own: invariants 6-8 already require every child chunk to come
    //  later in the backward order than its parent, whatever moved where)
    (void)has_child_chunk;
    return 0;
}

// The record-parallel run-merge (rm_*_kernel in pmb_kernels.cuh), block by block as the kernels decompose it, for ONE node's
// position-sorted records: breaks -> latest break per block -> carry across blocks (exclusive max-scan) -> piece flags
// ((index - run start) % 6 == 0) -> piece counts per block -> prefix -> pieces filled by looking ahead along the flags.
// `block` = records per block (the kernels use 2048; the test also uses tiny blocks so that runs straddle many of them).
// Returns the number of pieces; outputs sized n by the caller.
extern "C" long long emul_merge_runs(long long n, const int32_t* pos, const uint8_t* tc, const uint8_t* col_break, long long col_base,
                                     int32_t* nuc_position, uint8_t* mut_info, uint32_t* nucs, long long block) {
    if (n <= 0) return 0;
    const long long nb = (n + block - 1) / block;
    auto is_break = [&](long long i) {
        if (i == 0) return true;  // the node's first record
        const int32_t p = pos[i];
        return p != pos[i - 1] + 1 || (tc[i] >> 4) != (tc[i - 1] >> 4) || (col_break && col_break[p - col_base]);
    };
    std::vector<long long> last(nb, -1), carry(nb, -1), base(nb, 0);
    std::vector<unsigned> counts(nb, 0);
    std::vector<uint8_t> flags(n, 0);
    for (long long b = 0; b < nb; b++)
        for (long long i = b * block; i < std::min(n, (b + 1) * block); i++)
            if (is_break(i)) last[b] = i;
    for (long long b = 1; b < nb; b++) carry[b] = std::max(carry[b - 1], last[b - 1]);
    for (long long b = 0; b < nb; b++) {
        long long rs = carry[b];
        for (long long i = b * block; i < std::min(n, (b + 1) * block); i++) {
            if (is_break(i)) rs = i;
            if ((i - rs) % 6 == 0) {
                flags[i] = 2;
                counts[b]++;
            }
        }
    }
    for (long long b = 1; b < nb; b++) base[b] = base[b - 1] + counts[b - 1];
    for (long long b = 0; b < nb; b++) {
        long long o = base[b];
        for (long long i = b * block; i < std::min(n, (b + 1) * block); i++) {
            if (!(flags[i] & 2)) continue;
            const uint32_t t = uint32_t(tc[i]) >> 4;
            uint32_t packed = (uint32_t(tc[i]) & 15u) << 20;
            int len = 1;
            for (; len < 6 && i + len < n && !(flags[i + len] & 2); len++) packed += (uint32_t(tc[i + len]) & 15u) << (4 * (5 - len));
            nuc_position[o] = pos[i];
            mut_info[o] = uint8_t((len << 4) + int(t));
            nucs[o] = packed;
            o++;
        }
    }
    return base[nb - 1] + counts[nb - 1];
}

// expand_runs_kernel (clade-run encoded leaves -> code planes), lane by lane, against pack_leaves_kernel's planes of the same
// matrix: returns the number of plane words that differ (0 = the two ingest paths build the same leaf matrix), -1 on a bad
// tree. events / item_off / n_seg / seg_rows: what pmb_runs_describe reports for the encoding of leaf_codes.
extern "C" long long emul_expand_runs(int n_nodes, int root, const int32_t* child_off, const int32_t* child_idx, const int32_t* leaf_row,
                                      int chunk_nodes, int inline_nodes, long long n_cols, const uint8_t* leaf_codes /* n_rows x n_cols */,
                                      const uint8_t* parent_code, const uint32_t* events, const long long* item_off, int n_seg,
                                      int seg_rows) {
    TreeProgram P;
    if (!build_tree_program(n_nodes, root, child_off, child_idx, leaf_row, chunk_nodes, inline_nodes, &P).empty()) return -1;
    const int T = int((n_cols + TILE_COLS - 1) / TILE_COLS), n_rows = P.n_rows;
    std::vector<U4> want((size_t)n_rows * T * 32, U4{0, 0, 0, 0}), got((size_t)n_rows * T * 32, U4{0xDEADBEEFu, 0xDEADBEEFu, 0xDEADBEEFu, 0xDEADBEEFu});
    std::vector<U4> colparams((size_t)T * 128, U4{0, 0, 0, 0});
    auto setcode = [](U4& u, uint32_t code, uint32_t bit) {
        if (code & 1) u.x |= bit;
        if (code & 2) u.y |= bit;
        if (code & 4) u.z |= bit;
        if (code & 8) u.w |= bit;
    };
    for (long long c = 0; c < n_cols; c++) {
        const uint32_t bit = 1u << (c & 31);
        const size_t tile = size_t(c / TILE_COLS), lane = size_t((c % TILE_COLS) >> 5);
        setcode(colparams[tile * 128 + lane], parent_code[c] & 15u, bit);
        for (int r = 0; r < n_rows; r++)
            setcode(want[(tile * n_rows + P.row_slot[r]) * 32 + lane], leaf_codes[(size_t)r * n_cols + c] & 15u, bit);
    }
    // the leaves in depth-first order (children in Newick order) -> their slots
    std::vector<int32_t> dfs_slot;
    {
        std::vector<int32_t> stack{root};
        while (!stack.empty()) {
            const int32_t v = stack.back();
            stack.pop_back();
            if (child_off[v] == child_off[v + 1]) {
                dfs_slot.push_back(P.row_slot[leaf_row[v]]);
                continue;
            }
            for (int32_t e = child_off[v + 1] - 1; e >= child_off[v]; e--) stack.push_back(child_idx[e]);
        }
        if (int(dfs_slot.size()) != n_rows) return -1;
    }
    const long long n_items = (long long)T * n_seg;
    for (long long item = 0; item < n_items; item++) {  // one warp each
        const long long tile = item / n_seg;
        const int seg = int(item % n_seg);
        const int r0 = seg * seg_rows, nr = std::min(seg_rows, n_rows - r0);
        U4 cur[32];
        for (int lane = 0; lane < 32; lane++) cur[lane] = colparams[(size_t)tile * 128 + lane];
        long long k = item_off[item];
        const long long k_end = item_off[item + 1];
        uint32_t ev[32];
        auto load = [&]() {
            for (int lane = 0; lane < 32; lane++) ev[lane] = (k + lane < k_end) ? events[k + lane] : 0xFFFFFFFFu;
        };
        load();
        int j = 0;
        uint32_t next_row = ev[0] >> 14;
        for (int rb = 0; rb < nr; rb += 32) {
            int my_slot[32];
            for (int lane = 0; lane < 32; lane++) my_slot[lane] = (rb + lane < nr) ? dfs_slot[size_t(r0 + rb + lane)] : 0;
            const int lim = std::min(32, nr - rb);
            for (int i = 0; i < lim; i++) {
                while (next_row == uint32_t(rb + i)) {
                    const uint32_t e = ev[j];
                    const int lane = int((e >> 9) & 31u);
                    const uint32_t bit = 1u << ((e >> 4) & 31u);
                    cur[lane].x ^= (e & 1u) ? bit : 0u;
                    cur[lane].y ^= (e & 2u) ? bit : 0u;
                    cur[lane].z ^= (e & 4u) ? bit : 0u;
                    cur[lane].w ^= (e & 8u) ? bit : 0u;
                    if (++j == 32) {
                        k += 32;
                        load();
                        j = 0;
                    }
                    next_row = ev[j] >> 14;
                }
                const int slot = my_slot[i];
                for (int lane = 0; lane < 32; lane++) got[((size_t)tile * n_rows + slot) * 32 + lane] = cur[lane];
            }
        }
    }
    long long bad = 0;
    for (size_t i = 0; i < want.size(); i++)
        bad += (want[i].x != got[i].x) + (want[i].y != got[i].y) + (want[i].z != got[i].z) + (want[i].w != got[i].w);
    return bad;
}
