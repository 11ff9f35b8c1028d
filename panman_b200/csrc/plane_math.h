// Bit-plane ("bit-sliced") arithmetic shared by the CUDA kernels and by the CPU emulation used in tests.
//
// One 32-bit word holds ONE bit of the state of 32 consecutive alignment columns, so a bitwise
// instruction advances 32 columns at once and a warp advances 1024. Layouts:
//   code planes   c[4]  : bit b of the 4-bit IUPAC code            (leaves, assigned states)
//   Fitch set     S[16] : S[k] bit j <=> state k is in the set of column j   (reference keeps this as an
//                         int bitmask per column, src/fitchSankoff.cpp:44-55)
//   Sankoff excess G[16], H[16] : e_k > 0, e_k > 1 where e_k = min(2, cost_k - min cost)
//                         (the literal int[16] vector is fitchSankoff.cpp:391-402)
// Everything here is pure integer logic: results are bit-exact by construction and are checked against
// the oracle in tests/.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PMB_HD __host__ __device__ __forceinline__
#else
#define PMB_HD inline
#endif

namespace pmb {

// 4 code planes -> 16 one-hot planes: d[k] bit j <=> code of column j == k
PMB_HD void decode16(const uint32_t c[4], uint32_t d[16]) {
    uint32_t lo[4], hi[4];
    lo[0] = ~c[0] & ~c[1];
    lo[1] = c[0] & ~c[1];
    lo[2] = ~c[0] & c[1];
    lo[3] = c[0] & c[1];
    hi[0] = ~c[2] & ~c[3];
    hi[1] = c[2] & ~c[3];
    hi[2] = ~c[2] & c[3];
    hi[3] = c[2] & c[3];
#pragma unroll
    for (int k = 0; k < 16; k++) d[k] = lo[k & 3] & hi[k >> 2];
}

// value of plane-array X at the per-column index given by code planes p: out bit j = X[p_j] bit j
PMB_HD uint32_t mux16(const uint32_t X[16], const uint32_t p[4]) {
    uint32_t t[8], u[4], w[2];
#pragma unroll
    for (int j = 0; j < 8; j++) t[j] = (X[2 * j + 1] & p[0]) | (X[2 * j] & ~p[0]);
#pragma unroll
    for (int j = 0; j < 4; j++) u[j] = (t[2 * j + 1] & p[1]) | (t[2 * j] & ~p[1]);
#pragma unroll
    for (int j = 0; j < 2; j++) w[j] = (u[2 * j + 1] & p[2]) | (u[2 * j] & ~p[2]);
    return (w[1] & p[3]) | (w[0] & ~p[3]);
}

// per column: index of the lowest k with X[k] set, as code planes; any = OR of all planes.
// (the reference's "while(!(state & cur)) cur <<= 1" loops, fitchSankoff.cpp:107-110, and the first strict
// minimum of the Sankoff argmin loops, :496-504, :518-526)
PMB_HD void lowest16(const uint32_t X[16], uint32_t code[4], uint32_t& any) {
    uint32_t any2[8], any4[4], any8[2];
    uint32_t b0_2[4], b0_3[2], b1_3[2];
#pragma unroll
    for (int j = 0; j < 8; j++) any2[j] = X[2 * j] | X[2 * j + 1];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        any4[j] = any2[2 * j] | any2[2 * j + 1];
        // bit0 inside a pair is ~X[even]; choose the lower pair when it has anything
        b0_2[j] = (any2[2 * j] & ~X[4 * j]) | (~any2[2 * j] & ~X[4 * j + 2]);
    }
#pragma unroll
    for (int j = 0; j < 2; j++) {
        any8[j] = any4[2 * j] | any4[2 * j + 1];
        b0_3[j] = (any4[2 * j] & b0_2[2 * j]) | (~any4[2 * j] & b0_2[2 * j + 1]);
        b1_3[j] = (any4[2 * j] & ~any2[4 * j]) | (~any4[2 * j] & ~any2[4 * j + 2]);
    }
    any = any8[0] | any8[1];
    code[0] = (any8[0] & b0_3[0]) | (~any8[0] & b0_3[1]);
    code[1] = (any8[0] & b1_3[0]) | (~any8[0] & b1_3[1]);
    code[2] = (any8[0] & ~any4[0]) | (~any8[0] & ~any4[2]);
    code[3] = ~any8[0];
#pragma unroll
    for (int b = 0; b < 4; b++) code[b] &= any;
}

PMB_HD uint32_t differs4(const uint32_t a[4], const uint32_t b[4]) {
    return (a[0] ^ b[0]) | (a[1] ^ b[1]) | (a[2] ^ b[2]) | (a[3] ^ b[3]);
}

// per column a < b for 4-bit numbers given as planes
PMB_HD uint32_t less4(const uint32_t a[4], const uint32_t b[4]) {
    uint32_t lt = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {  // from LSB to MSB: a higher bit overrides
        uint32_t x = a[k] ^ b[k];
        lt = (x & b[k]) | (~x & lt);
    }
    return lt;
}

// ---------------- Fitch ----------------

struct FitchFold {
    uint32_t A[16], O[16];
    PMB_HD void reset() {
#pragma unroll
        for (int k = 0; k < 16; k++) { A[k] = 0xFFFFFFFFu; O[k] = 0; }
    }
    PMB_HD void add_set(const uint32_t S[16]) {
#pragma unroll
        for (int k = 0; k < 16; k++) { A[k] &= S[k]; O[k] |= S[k]; }
    }
    // leaf given by code planes; present = all-ones / zero mask (an omitted leaf contributes the empty set,
    // which kills the intersection and is ignored by the union: fitchSankoff.cpp:33-36, 48-55)
    PMB_HD void add_leaf(const uint32_t c[4], uint32_t present) {
        uint32_t d[16];
        decode16(c, d);
#pragma unroll
        for (int k = 0; k < 16; k++) { A[k] &= d[k] & present; O[k] |= d[k] & present; }
    }
    // S = AND ? AND : OR  per column (fitchSankoff.cpp:52-55)
    PMB_HD void finish(uint32_t S[16]) const {
        uint32_t nz = 0;
#pragma unroll
        for (int k = 0; k < 16; k++) nz |= A[k];
#pragma unroll
        for (int k = 0; k < 16; k++) S[k] = A[k] | (O[k] & ~nz);
    }
};

// Fast paths for the three binary shapes that make up almost every op of a bifurcating tree. They are algebraic
// specialisations of FitchFold for PRESENT leaves (one-hot sets), a third of the instructions:
//   leaf,leaf : equal codes -> that code, else both                      S = d1 | (d2 & ~eq)
//   leaf,set  : the leaf's state in X -> only it, else X plus it         S = d | (X & ~X[code])
//   set,set   : intersection if anywhere non-empty, else union           S = (X&Y) | ((X|Y) & ~any(X&Y))
PMB_HD void fitch_leaf_leaf(const uint32_t c1[4], const uint32_t c2[4], uint32_t S[16]) {
    uint32_t d1[16], d2[16];
    decode16(c1, d1);
    decode16(c2, d2);
    const uint32_t ne = differs4(c1, c2);
#pragma unroll
    for (int k = 0; k < 16; k++) S[k] = d1[k] | (d2[k] & ne);
}
PMB_HD void fitch_leaf_set(const uint32_t c[4], const uint32_t X[16], uint32_t S[16]) {
    uint32_t d[16];
    decode16(c, d);
    const uint32_t hit = mux16(X, c);
#pragma unroll
    for (int k = 0; k < 16; k++) S[k] = d[k] | (X[k] & ~hit);
}
PMB_HD void fitch_set_set(const uint32_t X[16], const uint32_t Y[16], uint32_t S[16]) {
    uint32_t nz = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) nz |= X[k] & Y[k];
#pragma unroll
    for (int k = 0; k < 16; k++) S[k] = (X[k] & Y[k]) | ((X[k] | Y[k]) & ~nz);
}

// ---- speculative evaluation of chain segments (tree_program.h "chain segments") ----
// Forward: the set S entering a segment from the segment below is unknown; lo <= S <= hi (as sets, per column)
// bounds it. One op on the path combines S with its other, known, children: A = their intersection, O = their
// union (FitchFold before finish()). Per column, with every known child non-empty:
//   lo&A != 0 : certainly intersects                 -> [lo&A, hi&A]
//   hi&A == 0 : certainly disjoint (A may be empty)  -> [lo|O, hi|O]
//   else      : either                               -> [one & (lo|O), hi|O], one = hi&A if that is a single state
// (if it does intersect, the result is a non-empty subset of hi&A, hence exactly hi&A when that is one state). The
// value is known exactly where lo == hi: for one-hot leaves that happens as soon as two consecutive leaves agree.
struct FitchInterval {
    uint32_t lo[16], hi[16];
    PMB_HD void reset() {
#pragma unroll
        for (int k = 0; k < 16; k++) { lo[k] = 0; hi[k] = 0xFFFFFFFFu; }
    }
    PMB_HD void step(const uint32_t A[16], const uint32_t O[16]) {
        uint32_t sure = 0, maybe = 0, one = 0, two = 0;
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const uint32_t h = hi[k] & A[k];
            sure |= lo[k] & A[k];
            maybe |= h;
            two |= one & h;
            one |= h;
        }
        const uint32_t single = one & ~two, uni = ~maybe, either = maybe & ~sure;
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const uint32_t l = lo[k], h = hi[k];
            lo[k] = (sure & l & A[k]) | (uni & (l | O[k])) | (either & single & h & A[k] & (l | O[k]));
            hi[k] = (sure & h & A[k]) | (~sure & (h | O[k]));
        }
    }
    // columns where the value is not known exactly yet
    PMB_HD uint32_t open() const {
        uint32_t d = 0;
#pragma unroll
        for (int k = 0; k < 16; k++) d |= lo[k] ^ hi[k];
        return d;
    }
};

// Backward: the state a segment's top node receives from above is unknown; Q = the states it may be (one-hot planes).
// A node with set S != 0 keeps the parent's state if S has it, else takes lowest(S)  (fitch_assign below), so its own
// candidates are (Q & S) plus lowest(S) if some candidate is outside S. Known exactly where one candidate is left.
PMB_HD void fitch_candidates_step(uint32_t Q[16], const uint32_t S[16]) {
    uint32_t outside = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) outside |= Q[k] & ~S[k];
    uint32_t seen = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const uint32_t lowest = S[k] & ~seen;
        seen |= S[k];
        Q[k] = (Q[k] & S[k]) | (outside & lowest);
    }
}
PMB_HD uint32_t candidates_open(const uint32_t Q[16]) {  // columns with more (or fewer) than one candidate
    uint32_t one = 0, two = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        two |= one & Q[k];
        one |= Q[k];
    }
    return ~one | two;
}
PMB_HD void encode16(const uint32_t d[16], uint32_t c[4]) {  // one-hot planes -> code planes
    c[0] = c[1] = c[2] = c[3] = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        if (k & 1) c[0] |= d[k];
        if (k & 2) c[1] |= d[k];
        if (k & 4) c[2] |= d[k];
        if (k & 8) c[3] |= d[k];
    }
}

// Assigned state of a non-root node (or a block-mode root): parent state P (code planes) where visited pvis.
// F = P if P in S else lowest(S); vis = pvis & (S != 0)          (fitchSankoff.cpp:101-103, 115-123)
PMB_HD void fitch_assign(const uint32_t S[16], const uint32_t P[4], uint32_t pvis, uint32_t F[4], uint32_t& vis) {
    uint32_t low[4], any;
    lowest16(S, low, any);
    uint32_t hit = mux16(S, P);
    vis = pvis & any;
#pragma unroll
    for (int b = 0; b < 4; b++) F[b] = ((hit & P[b]) | (~hit & low[b])) & vis;
}

// Root, nuc mode: override if given, else lowest(S)                (fitchSankoff.cpp:98-99, 104-114)
PMB_HD void fitch_assign_root(const uint32_t S[16], const uint32_t ov[4], uint32_t ov_valid, uint32_t colmask,
                              uint32_t F[4], uint32_t& vis) {
    uint32_t low[4], any;
    lowest16(S, low, any);
    vis = (ov_valid | any) & colmask;
#pragma unroll
    for (int b = 0; b < 4; b++) F[b] = ((ov_valid & ov[b]) | (~ov_valid & low[b])) & vis;
}

// Mutation type planes from parent/child code planes: NI where the parent is '-', ND where the child is '-',
// NS otherwise (fitchSankoff.cpp:140-167, :683-698). Returned as 2 planes (type bit0, bit1): NS=0 ND=1 NI=2.
PMB_HD void mutation_type(const uint32_t P[4], const uint32_t F[4], uint32_t& t0, uint32_t& t1) {
    uint32_t pgap = ~(P[0] | P[1] | P[2] | P[3]);
    uint32_t fgap = ~(F[0] | F[1] | F[2] | F[3]);
    t1 = pgap;
    t0 = ~pgap & fgap;
}

// ---------------- Sankoff (2-bit excess form) ----------------

// NONE marker of a node whose every leaf below is omitted (the reference's all-INF vector,
// fitchSankoff.cpp:376-389): G[0] = 0 with H[0] = 1, a combination no real excess produces.
PMB_HD uint32_t sankoff_none(const uint32_t G[16], const uint32_t H[16]) { return H[0] & ~G[0]; }

template <int B>
struct SankoffFold {
    uint32_t cnt[16][B];  // bit-sliced counters r_k = #children with e_k > 0
    uint32_t all_none;
    PMB_HD void reset() {
#pragma unroll
        for (int k = 0; k < 16; k++)
#pragma unroll
            for (int b = 0; b < B; b++) cnt[k][b] = 0;
        all_none = 0xFFFFFFFFu;
    }
    PMB_HD void add_bit(int k, uint32_t g) {
        uint32_t carry = g;
#pragma unroll
        for (int b = 0; b < B; b++) {
            uint32_t t = cnt[k][b] & carry;
            cnt[k][b] ^= carry;
            carry = t;
        }
    }
    // internal child: its G planes and NONE plane (NONE children are skipped, fitchSankoff.cpp:398-400)
    PMB_HD void add_set(const uint32_t G[16], uint32_t none) {
#pragma unroll
        for (int k = 0; k < 16; k++) add_bit(k, G[k] & ~none);
        all_none &= none;
    }
    PMB_HD void add_leaf(const uint32_t c[4], uint32_t present) {
        uint32_t d[16];
        decode16(c, d);
#pragma unroll
        for (int k = 0; k < 16; k++) add_bit(k, ~d[k] & present);
        all_none &= ~present;
    }
    // e_k = min(2, r_k - min r)
    PMB_HD void finish(uint32_t G[16], uint32_t H[16]) const {
        uint32_t alive[16];
        uint32_t minbit[B];
#pragma unroll
        for (int k = 0; k < 16; k++) alive[k] = 0xFFFFFFFFu;
#pragma unroll
        for (int b = B - 1; b >= 0; b--) {
            uint32_t anyzero = 0;
#pragma unroll
            for (int k = 0; k < 16; k++) anyzero |= alive[k] & ~cnt[k][b];
            minbit[b] = ~anyzero;
#pragma unroll
            for (int k = 0; k < 16; k++) alive[k] &= ~(anyzero & cnt[k][b]);
        }
        // m + 1 (cannot overflow B bits on a state that matters: r_k <= children < 2^B)
        uint32_t mp1[B];
        uint32_t carry = 0xFFFFFFFFu;
#pragma unroll
        for (int b = 0; b < B; b++) {
            mp1[b] = minbit[b] ^ carry;
            carry &= minbit[b];
        }
#pragma unroll
        for (int k = 0; k < 16; k++) {
            uint32_t eq1 = ~carry;  // m+1 overflowed => nothing equals it
#pragma unroll
            for (int b = 0; b < B; b++) eq1 &= ~(cnt[k][b] ^ mp1[b]);
            G[k] = ~alive[k];
            H[k] = ~alive[k] & ~eq1;
        }
        // all children NONE: r = 0 everywhere gave G = H = 0; stamp the NONE marker
        H[0] |= all_none;
        G[0] &= ~all_none;
    }
};

// Two children (the bifurcating common case), closed form of SankoffFold<2>: with a_k, b_k = "child has excess > 0
// at state k" (0 for a NONE child), r_k = a_k + b_k in {0,1,2} and m = min_k r_k:
//   m = 0 (some state free in both):  e = r        -> G = a|b, H = a&b
//   m = 1:                            e = r - 1    -> G = a&b, H = 0
//   m = 2 (only if a&b everywhere):   e = 0        -> G = H = 0
PMB_HD void sankoff_pair(const uint32_t g1[16], uint32_t none1, const uint32_t g2[16], uint32_t none2, uint32_t G[16],
                         uint32_t H[16]) {
    uint32_t any0 = 0, all2 = 0xFFFFFFFFu;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const uint32_t a = g1[k] & ~none1, b = g2[k] & ~none2;
        any0 |= ~(a | b);
        all2 &= a & b;
    }
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const uint32_t a = g1[k] & ~none1, b = g2[k] & ~none2;
        G[k] = (any0 & (a | b)) | (~any0 & a & b & ~all2);
        H[k] = any0 & a & b;
    }
    const uint32_t both_none = none1 & none2;
    H[0] |= both_none;
    G[0] &= ~both_none;
}
// excess>0 planes of a leaf: every state but its own code (nothing, and NONE, when the leaf is omitted)
PMB_HD void sankoff_leaf_g(const uint32_t c[4], uint32_t present, uint32_t g[16]) {
    uint32_t d[16];
    decode16(c, d);
#pragma unroll
    for (int k = 0; k < 16; k++) g[k] = ~d[k] & present;
}

// Child pointer chosen by a parent in state P (fitchSankoff.cpp:518-529 in excess form):
// e[s]==0 -> s ; e[s]==1 -> min(s, z) ; e[s]==2 -> z, with z the lowest zero-excess state.
PMB_HD void sankoff_assign(const uint32_t G[16], const uint32_t H[16], const uint32_t P[4], uint32_t pvis,
                           uint32_t F[4], uint32_t& vis) {
    uint32_t zero[16], z[4], any;
#pragma unroll
    for (int k = 0; k < 16; k++) zero[k] = ~G[k];
    lowest16(zero, z, any);
    uint32_t gt0 = mux16(G, P), eq2 = mux16(H, P);
    uint32_t keep = ~gt0 | (~eq2 & less4(P, z));
    vis = pvis & ~sankoff_none(G, H);
#pragma unroll
    for (int b = 0; b < 4; b++) F[b] = ((keep & P[b]) | (~keep & z[b])) & vis;
}

// ---- speculative evaluation of chain segments, Sankoff (tree_program.h "chain segments") ----
// Forward: a parent's vector depends on a child only through the child's zero-excess set Z = {k : e_k = 0} (every
// child adds 0 to the parent's count r_k where k is in Z and 1 elsewhere). For a node with exactly two
// children, Z(parent) = Z1 & Z2 if that is non-empty, else Z1 | Z2 -- the Fitch rule -- so FitchInterval bounds the
// unknown Z entering a segment, with A = O = Z(known child). Where the bounds meet, G = ~Z is exact from there on;
// H of that node still depends on the unknown below it and is redone with the ops before it.
// Backward: the state a segment's top node receives is unknown; Q = the states it may be (one-hot planes). A node in
// parent state s takes s if e_s = 0, min(s, z) if e_s = 1, z if e_s = 2, z = lowest zero-excess state (sankoff_assign).
PMB_HD void sankoff_candidates_step(uint32_t Q[16], const uint32_t G[16], const uint32_t H[16]) {
    uint32_t to_z = 0, seen = 0;
    uint32_t keep[16], zhot[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const uint32_t zero = ~G[k];
        zhot[k] = zero & ~seen;     // k is the lowest zero-excess state
        seen |= zero;               // some zero-excess state <= k, i.e. NOT (k < z)
        const uint32_t one = G[k] & ~H[k];
        keep[k] = Q[k] & (zero | (one & ~seen));             // e = 0, or e = 1 and k < z: stays k
        to_z |= Q[k] & (H[k] | (one & seen));                // e = 2, or e = 1 and k > z: becomes z
    }
#pragma unroll
    for (int k = 0; k < 16; k++) Q[k] = keep[k] | (to_z & zhot[k]);
}

// Root: override if given, else the first minimum (fitchSankoff.cpp:492-507). undefined = columns where
// the reference would trip assert(minPtr != -1).
PMB_HD void sankoff_assign_root(const uint32_t G[16], const uint32_t H[16], const uint32_t ov[4], uint32_t ov_valid,
                                uint32_t colmask, uint32_t F[4], uint32_t& vis, uint32_t& undefined) {
    uint32_t zero[16], z[4], any;
#pragma unroll
    for (int k = 0; k < 16; k++) zero[k] = ~G[k];
    lowest16(zero, z, any);
    uint32_t none = sankoff_none(G, H);
    undefined = none & ~ov_valid & colmask;
    vis = (ov_valid | ~none) & colmask;
#pragma unroll
    for (int b = 0; b < 4; b++) F[b] = ((ov_valid & ov[b]) | (~ov_valid & z[b])) & vis;
}

}  // namespace pmb
