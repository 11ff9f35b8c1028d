/* TEST INFRASTRUCTURE ONLY -- CPU restatement ("port") of PanMAN's Fitch/Sankoff construction path.
 *
 * This file is the checker for libpanman_b200, never part of it: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * It restates, on flat arrays and without recursion (trees here reach depth 1e5), the
 * algorithms of the reference's src/fitchSankoff.cpp and the caller conventions of
 * src/panman.cpp. Every function cites the reference lines it follows.
 *
 * PARITY PINNING: the reference ships no golden vectors (its test/ holds three data files,
 * SURVEY.md section 4). This port is pinned instead against the reference's own
 * fitchSankoff.cpp compiled verbatim (oracle/_ref/libpanman_ref.so, see oracle/Makefile):
 * tests/test_oracle.py compares them column by column, and
 * tests/golden/ holds vectors generated from that verbatim build
 * (tests/golden/make_golden.py).
 *
 * Conventions shared with include/panman_b200.h:
 *   node ids 0..n_nodes-1, children of v = child_idx[child_off[v] .. child_off[v+1]) in Newick order;
 *   leaf_row[v] = row of leaf v in the code matrix, -1 for internal nodes;
 *   codes are the 4-bit IUPAC codes of reference src/panman.hpp:27-44 ('-' and anything unknown = 0);
 *   record type: NS=0 ND=1 NI=2 (reference src/panman.hpp:46-61).
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_INF 100000001 /* reference src/common.hpp:16 */
#define ORC_NO_DEFAULT (1 << 28) /* reference src/panman.hpp:851 default argument */

/* reference src/panman.cpp:78-113 */
int orc_code_from_nucleotide(int c) {
    switch (c) {
    case 'A': return 1;  case 'C': return 2;  case 'G': return 4;  case 'T': return 8;
    case 'R': return 5;  case 'Y': return 10; case 'S': return 6;  case 'W': return 9;
    case 'K': return 12; case 'M': return 3;  case 'B': return 14; case 'D': return 13;
    case 'H': return 11; case 'V': return 7;  case 'N': return 15;
    default: return 0;
    }
}

/* reference src/panman.cpp:41-76 */
int orc_nucleotide_from_code(int code) {
    static const char t[] = "-ACMGRSVTWYHKDBN";
    return (code >= 1 && code <= 15) ? t[code] : '-';
}

/* ---------- traversal orders (replace the reference's recursion) ---------- */

/* post[] receives the nodes in the order the reference's recursive forward pass finishes them:
 * children left to right, then the node (fitchSankoff.cpp:39-42). Returns the count. */
int orc_post_order(int n_nodes, int root, const int32_t* child_off, const int32_t* child_idx, int32_t* post) {
    int32_t* stack = (int32_t*)malloc(sizeof(int32_t) * (size_t)n_nodes);
    int32_t* next = (int32_t*)malloc(sizeof(int32_t) * (size_t)n_nodes);
    int sp = 0, k = 0;
    stack[sp] = root;
    next[sp] = child_off[root];
    sp++;
    while (sp > 0) {
        int v = stack[sp - 1];
        if (next[sp - 1] < child_off[v + 1]) {
            int c = child_idx[next[sp - 1]++];
            stack[sp] = c;
            next[sp] = child_off[c];
            sp++;
        } else {
            post[k++] = v;
            sp--;
        }
    }
    free(stack);
    free(next);
    return k;
}

static void parents_of(int n_nodes, const int32_t* child_off, const int32_t* child_idx, int32_t* parent) {
    for (int v = 0; v < n_nodes; v++) parent[v] = -1;
    for (int v = 0; v < n_nodes; v++)
        for (int e = child_off[v]; e < child_off[v + 1]; e++) parent[child_idx[e]] = v;
}

/* ---------- nuc / block Fitch, one column ---------- */

/* fwd: nucFitchForwardPass fitchSankoff.cpp:30-56 (block: blockFitchForwardPassNew :224-245).
 * S[v] for leaves is the input state (0 = leaf omitted from the map, :33-36). */
static void fitch_forward(int n, const int32_t* post, int root, const int32_t* child_off, const int32_t* child_idx,
                          int* S, int ref_state) {
    for (int i = 0; i < n; i++) {
        int v = post[i];
        int b = child_off[v], e = child_off[v + 1];
        if (b == e) continue; /* leaf: keeps its state */
        if (v == root && ref_state != -1) { /* :45-47 (children were evaluated first) */
            S[v] = ref_state;
            continue;
        }
        int orS = 0, andS = S[child_idx[b]]; /* :44 */
        for (int k = b; k < e; k++) {
            orS |= S[child_idx[k]];
            andS &= S[child_idx[k]];
        }
        S[v] = andS ? andS : orS; /* :52-55 */
    }
}

static int lowest_bit(int s) { /* the while-loops at :107-110, :118-121 */
    int cur = 1;
    while (!(s & cur)) cur <<= 1;
    return cur;
}

/* bwd: nucFitchBackwardPass :96-129 (block_mode: blockFitchBackwardPassNew :247-270, which has no
 * root-takes-lowest-bit case). F[v] = assigned one-hot state, 0 where the node is never assigned:
 * either its own set is 0 (:101-103) or an ancestor's was, so the recursion never reached it.
 * assign: nucFitchAssignMutations :131-171 -- visits exactly the nodes with F != 0. */
static void fitch_backward_assign(int n, const int32_t* post, int root, const int32_t* parent, const int* S,
                                  int parent_state, int default_state, int block_mode, int* F, int* mut_type,
                                  int* mut_code) {
    for (int i = n - 1; i >= 0; i--) { /* reverse post-order: parents before children */
        int v = post[i];
        int p;
        F[v] = 0;
        mut_type[v] = -1;
        mut_code[v] = 0;
        if (v == root) {
            p = parent_state;
            if (default_state != ORC_NO_DEFAULT) F[v] = default_state;        /* :98-99 */
            else if (S[v] == 0) continue;                                      /* :101-103 */
            else if (!block_mode) F[v] = lowest_bit(S[v]);                     /* :104-114 */
            else F[v] = (p & S[v]) ? p : lowest_bit(S[v]);                     /* :255-263 */
        } else {
            p = F[parent[v]];
            if (p == 0) continue;  /* parent never assigned => recursion never got here */
            if (S[v] == 0) continue;                                           /* :101-103 */
            F[v] = (p & S[v]) ? p : lowest_bit(S[v]);                          /* :115-123 */
        }
        if (F[v] == 0) continue;   /* default_state 0 cannot happen for one-hot input; :136-138 */
        if (p != F[v]) {                                                        /* :140 */
            int code = 0, cur = F[v];
            while (cur > 1) { cur >>= 1; code++; }
            if (p == 1) { mut_type[v] = 2; mut_code[v] = code; }               /* NI :141-151 */
            else if (F[v] == 1) { mut_type[v] = 1; mut_code[v] = 0; }          /* ND :152-154 */
            else { mut_type[v] = 0; mut_code[v] = code; }                      /* NS :155-166 */
        }
    }
}

/* One column, Fitch. leaf_state[v] used for leaves only (0 or negative = omitted). Arrays sized n_nodes.
 * out_fwd may be NULL. mut_code is the 4-bit code of the character the reference stores
 * (getCodeFromNucleotide(getNucleotideFromCode(code)) is the identity on 0..15). */
int orc_fitch_column(int n_nodes, int root, const int32_t* child_off, const int32_t* child_idx,
                     const int* leaf_state, int fwd_ref_state, int parent_state, int default_state, int block_mode,
                     int* out_fwd, int* out_final, int* out_mut_type, int* out_mut_code) {
    int32_t* post = (int32_t*)malloc(sizeof(int32_t) * (size_t)n_nodes);
    int32_t* parent = (int32_t*)malloc(sizeof(int32_t) * (size_t)n_nodes);
    int* S = (int*)malloc(sizeof(int) * (size_t)n_nodes);
    int n = orc_post_order(n_nodes, root, child_off, child_idx, post);
    parents_of(n_nodes, child_off, child_idx, parent);
    for (int v = 0; v < n_nodes; v++) S[v] = (child_off[v] == child_off[v + 1] && leaf_state[v] > 0) ? leaf_state[v] : 0;
    fitch_forward(n, post, root, child_off, child_idx, S, block_mode ? -1 : fwd_ref_state);
    if (out_fwd) memcpy(out_fwd, S, sizeof(int) * (size_t)n_nodes);
    fitch_backward_assign(n, post, root, parent, S, parent_state, default_state, block_mode, out_final, out_mut_type,
                          out_mut_code);
    free(post);
    free(parent);
    free(S);
    return 0;
}

/* ---------- nuc / block Sankoff, one column (literal min-plus form) ---------- */

/* fwd: nucSankoffForwardPass fitchSankoff.cpp:359-405; block (W=3): blockSankoffForwardPass :707-735,
 * whose missing leaf is {0,INF,INF} (:711-714) and which has no all-INF shortcut and no INF skip. */
static void sankoff_forward(int n, const int32_t* post, const int32_t* child_off, const int32_t* child_idx, int W,
                            int block_mode, int* cost /* n_nodes x W */) {
    for (int i = 0; i < n; i++) {
        int v = post[i];
        int b = child_off[v], e = child_off[v + 1];
        if (b == e) continue;
        int* cv = cost + (size_t)v * W;
        if (!block_mode) {
            int min_exists = 0; /* :376-389 */
            for (int k = b; k < e && !min_exists; k++)
                for (int s = 0; s < W; s++)
                    if (cost[(size_t)child_idx[k] * W + s] < ORC_INF) { min_exists = 1; break; }
            if (!min_exists) {
                for (int s = 0; s < W; s++) cv[s] = ORC_INF;
                continue;
            }
        }
        for (int s = 0; s < W; s++) { /* :391-402 / :723-732 */
            int acc = 0;
            for (int k = b; k < e; k++) {
                const int* cc = cost + (size_t)child_idx[k] * W;
                int mv = ORC_INF;
                for (int t = 0; t < W; t++) {
                    int cand = (s != t) + cc[t];
                    if (cand < mv) mv = cand;
                }
                if (block_mode || mv < ORC_INF) acc += mv;
            }
            cv[s] = acc;
        }
    }
}

/* bwd: nucSankoffBackwardPass :487-531, block: blockSankoffBackwardPass :737-786.
 * assign: nucSankoffAssignMutations :676-703, blockSankoffAssignMutations :788-818.
 * F[v] = assigned state index, -1 = none. Returns -2 where the reference would assert (:505). */
static int sankoff_backward_assign(int n, const int32_t* post, int root, const int32_t* parent, int W, int block_mode,
                                   const int* cost, int parent_state, int default_state, int* F, int* mut_type,
                                   int* mut_code) {
    for (int i = n - 1; i >= 0; i--) {
        int v = post[i];
        const int* cv = cost + (size_t)v * W;
        int p;
        F[v] = -1;
        mut_type[v] = -1;
        mut_code[v] = 0;
        if (v == root) {
            p = parent_state;
            if (default_state != ORC_NO_DEFAULT) F[v] = default_state; /* :492-493 */
            else {
                int mv = ORC_INF, mp = -1; /* :496-504 */
                for (int s = 0; s < W; s++)
                    if (cv[s] < mv) { mv = cv[s]; mp = s; }
                if (mp == -1) {
                    if (!block_mode) return -2; /* assert(minPtr != -1) :505 */
                    continue;                   /* block: states = -1, return :754-757 */
                }
                F[v] = mp;
            }
        } else {
            int pv = parent[v];
            p = F[pv];
            if (p == -1) continue; /* nuc: children of -1 get -1 (:513-516); block: never reached */
            if (block_mode) {      /* :761-771 */
                int exists = 0;
                for (int s = 0; s < W; s++) exists |= (cv[s] < ORC_INF);
                if (!exists) continue;
            }
            int mv = ORC_INF, mp = -1; /* parent picks the child's pointer :518-529 / :775-784 */
            for (int s = 0; s < W; s++) {
                int cand = (s != p) + cv[s];
                if (cand < mv) { mv = cand; mp = s; }
            }
            F[v] = mp;
            if (mp == -1) continue;
        }
        if (p != F[v]) { /* :683-698 */
            if (p == 0) { mut_type[v] = 2; mut_code[v] = F[v]; }
            else if (F[v] == 0) { mut_type[v] = 1; mut_code[v] = 0; }
            else { mut_type[v] = 0; mut_code[v] = F[v]; }
        }
    }
    return 0;
}

/* One column, Sankoff. leaf_code[v] for leaves: state index, <0 = omitted. W = 16 (nuc) or 3 (block). */
int orc_sankoff_column(int n_nodes, int root, const int32_t* child_off, const int32_t* child_idx, const int* leaf_code,
                       int W, int block_mode, int parent_state, int default_state, int* out_cost, int* out_final,
                       int* out_mut_type, int* out_mut_code) {
    int32_t* post = (int32_t*)malloc(sizeof(int32_t) * (size_t)n_nodes);
    int32_t* parent = (int32_t*)malloc(sizeof(int32_t) * (size_t)n_nodes);
    int* cost = (int*)malloc(sizeof(int) * (size_t)n_nodes * (size_t)W);
    int n = orc_post_order(n_nodes, root, child_off, child_idx, post);
    parents_of(n_nodes, child_off, child_idx, parent);
    for (int v = 0; v < n_nodes; v++) {
        int* cv = cost + (size_t)v * W;
        for (int s = 0; s < W; s++) cv[s] = ORC_INF;
        if (child_off[v] == child_off[v + 1]) {
            if (leaf_code[v] >= 0) cv[leaf_code[v]] = 0;
            else if (block_mode) cv[0] = 0; /* :711-714 */
        }
    }
    sankoff_forward(n, post, child_off, child_idx, W, block_mode, cost);
    if (out_cost) memcpy(out_cost, cost, sizeof(int) * (size_t)n_nodes * (size_t)W);
    int rc = sankoff_backward_assign(n, post, root, parent, W, block_mode, cost, parent_state, default_state, out_final,
                                     out_mut_type, out_mut_code);
    free(post);
    free(parent);
    free(cost);
    return rc;
}

/* ---------- batch over columns, same inputs as pmb_run_nuc (include/panman_b200.h) ---------- */

typedef struct {
    int32_t node;
    int32_t pos;
    uint8_t type, code;
} orc_rec;

typedef struct {
    /* inputs */
    int algo, n_nodes, root, block_mode;
    const int32_t *child_off, *child_idx, *leaf_row, *post, *parent;
    int n_post;
    int64_t n_cols, col_begin, col_end;
    const uint8_t* leaf_codes; /* n_rows x n_cols, one code per byte */
    const uint8_t* leaf_present;
    const uint8_t* parent_code;
    const int8_t *root_override, *fwd_root_ref;
    uint8_t* out_states; /* n_nodes x n_cols or NULL */
    /* outputs */
    orc_rec* recs;
    int64_t n_recs, cap;
    int rc;
} orc_job;

static void job_push(orc_job* j, int node, int64_t pos, int type, int code) {
    if (j->n_recs == j->cap) {
        j->cap = j->cap ? j->cap * 2 : 1024;
        j->recs = (orc_rec*)realloc(j->recs, sizeof(orc_rec) * (size_t)j->cap);
    }
    orc_rec r = {node, (int32_t)pos, (uint8_t)type, (uint8_t)code};
    j->recs[j->n_recs++] = r;
}

static void* job_run(void* arg) {
    orc_job* j = (orc_job*)arg;
    int n = j->n_nodes;
    int W = 16;
    int* S = (int*)malloc(sizeof(int) * (size_t)n);
    int* F = (int*)malloc(sizeof(int) * (size_t)n);
    int* mt = (int*)malloc(sizeof(int) * (size_t)n);
    int* mc = (int*)malloc(sizeof(int) * (size_t)n);
    int* cost = j->algo == 1 ? (int*)malloc(sizeof(int) * (size_t)n * W) : NULL;
    j->rc = 0;
    for (int64_t c = j->col_begin; c < j->col_end; c++) {
        int pc = j->parent_code[c];
        int ov = j->root_override ? j->root_override[c] : -1;
        if (j->algo == 0) {
            /* leaf state = 1 << code; '-' (code 0) => 1           reference panman.cpp:1409-1417 */
            for (int v = 0; v < n; v++) {
                int r = j->leaf_row[v];
                S[v] = (r >= 0 && (!j->leaf_present || j->leaf_present[r])) ? (1 << j->leaf_codes[(size_t)r * j->n_cols + c]) : 0;
            }
            int fr = (j->fwd_root_ref && j->fwd_root_ref[c] >= 0) ? (1 << j->fwd_root_ref[c]) : -1;
            fitch_forward(j->n_post, j->post, j->root, j->child_off, j->child_idx, S, j->block_mode ? -1 : fr);
            fitch_backward_assign(j->n_post, j->post, j->root, j->parent, S, 1 << pc, ov >= 0 ? (1 << ov) : ORC_NO_DEFAULT,
                                  j->block_mode, F, mt, mc);
            if (j->out_states)
                for (int v = 0; v < n; v++) {
                    int code = 0xFF;
                    if (F[v]) { code = 0; while ((F[v] >> code) > 1) code++; }
                    j->out_states[(size_t)v * j->n_cols + c] = (uint8_t)code;
                }
        } else {
            /* leaf vector: 0 at its code, INF elsewhere           reference panman.cpp:1574-1582 */
            for (int v = 0; v < n; v++) {
                int* cv = cost + (size_t)v * W;
                for (int s = 0; s < W; s++) cv[s] = ORC_INF;
                int r = j->leaf_row[v];
                if (r >= 0) {
                    if (!j->leaf_present || j->leaf_present[r]) cv[j->leaf_codes[(size_t)r * j->n_cols + c]] = 0;
                    else if (j->block_mode) cv[0] = 0;
                }
            }
            sankoff_forward(j->n_post, j->post, j->child_off, j->child_idx, W, j->block_mode, cost);
            int rc = sankoff_backward_assign(j->n_post, j->post, j->root, j->parent, W, j->block_mode, cost, pc,
                                             ov >= 0 ? ov : ORC_NO_DEFAULT, F, mt, mc);
            if (rc) { j->rc = rc; break; }
            if (j->out_states)
                for (int v = 0; v < n; v++) j->out_states[(size_t)v * j->n_cols + c] = (uint8_t)(F[v] < 0 ? 0xFF : F[v]);
        }
        for (int v = 0; v < n; v++)
            if (mt[v] >= 0) job_push(j, v, c, mt[v], mc[v]);
    }
    free(S);
    free(F);
    free(mt);
    free(mc);
    free(cost);
    return NULL;
}

typedef struct {
    int n_nodes;
    int64_t n_mut;
    int64_t* node_offsets; /* n_nodes + 1 */
    int32_t* pos;
    uint8_t* type_code; /* (type << 4) | code */
} orc_result;

/* Runs all columns on n_threads threads (contiguous column ranges). Result lists are per node in
 * ascending column order, which is what the reference's std::sort of (pos,type,code) tuples gives
 * (panman.cpp:1447) because a node has at most one record per column.
 * algo: 0 Fitch, 1 Sankoff. block_mode: 3-state block variants' root/leaf conventions. */
int orc_run(int algo, int block_mode, int n_nodes, int root, const int32_t* child_off, const int32_t* child_idx,
            const int32_t* leaf_row, int64_t n_cols, const uint8_t* leaf_codes, const uint8_t* leaf_present,
            const uint8_t* parent_code, const int8_t* root_override, const int8_t* fwd_root_ref, int n_threads,
            uint8_t* out_states, orc_result** out) {
    int32_t* post = (int32_t*)malloc(sizeof(int32_t) * (size_t)n_nodes);
    int32_t* parent = (int32_t*)malloc(sizeof(int32_t) * (size_t)n_nodes);
    int n_post = orc_post_order(n_nodes, root, child_off, child_idx, post);
    parents_of(n_nodes, child_off, child_idx, parent);
    if (n_threads < 1) n_threads = 1;
    if ((int64_t)n_threads > n_cols) n_threads = n_cols > 0 ? (int)n_cols : 1;
    orc_job* jobs = (orc_job*)calloc((size_t)n_threads, sizeof(orc_job));
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)n_threads);
    for (int t = 0; t < n_threads; t++) {
        orc_job* j = &jobs[t];
        j->algo = algo; j->n_nodes = n_nodes; j->root = root; j->block_mode = block_mode;
        j->child_off = child_off; j->child_idx = child_idx; j->leaf_row = leaf_row; j->post = post; j->parent = parent;
        j->n_post = n_post; j->n_cols = n_cols;
        j->col_begin = n_cols * t / n_threads; j->col_end = n_cols * (t + 1) / n_threads;
        j->leaf_codes = leaf_codes; j->leaf_present = leaf_present; j->parent_code = parent_code;
        j->root_override = root_override; j->fwd_root_ref = fwd_root_ref; j->out_states = out_states;
        if (n_threads == 1) job_run(j);
        else pthread_create(&th[t], NULL, job_run, j);
    }
    int rc = 0;
    int64_t total = 0;
    for (int t = 0; t < n_threads; t++) {
        if (n_threads > 1) pthread_join(th[t], NULL);
        if (jobs[t].rc) rc = jobs[t].rc;
        total += jobs[t].n_recs;
    }
    orc_result* r = (orc_result*)calloc(1, sizeof(orc_result));
    r->n_nodes = n_nodes;
    r->n_mut = total;
    r->node_offsets = (int64_t*)calloc((size_t)n_nodes + 1, sizeof(int64_t));
    r->pos = (int32_t*)malloc(sizeof(int32_t) * (size_t)(total ? total : 1));
    r->type_code = (uint8_t*)malloc((size_t)(total ? total : 1));
    for (int t = 0; t < n_threads; t++)
        for (int64_t k = 0; k < jobs[t].n_recs; k++) r->node_offsets[jobs[t].recs[k].node + 1]++;
    for (int v = 0; v < n_nodes; v++) r->node_offsets[v + 1] += r->node_offsets[v];
    int64_t* cur = (int64_t*)malloc(sizeof(int64_t) * (size_t)n_nodes);
    memcpy(cur, r->node_offsets, sizeof(int64_t) * (size_t)n_nodes);
    for (int t = 0; t < n_threads; t++) { /* thread order == column order: stable per node */
        for (int64_t k = 0; k < jobs[t].n_recs; k++) {
            orc_rec* q = &jobs[t].recs[k];
            int64_t at = cur[q->node]++;
            r->pos[at] = q->pos;
            r->type_code[at] = (uint8_t)((q->type << 4) | q->code);
        }
        free(jobs[t].recs);
    }
    free(cur);
    free(jobs);
    free(th);
    free(post);
    free(parent);
    *out = r;
    return rc;
}

int64_t orc_result_n_mut(const orc_result* r) { return r->n_mut; }
const int64_t* orc_result_offsets(const orc_result* r) { return r->node_offsets; }
const int32_t* orc_result_pos(const orc_result* r) { return r->pos; }
const uint8_t* orc_result_type_code(const orc_result* r) { return r->type_code; }
void orc_result_free(orc_result* r) {
    if (!r) return;
    free(r->node_offsets);
    free(r->pos);
    free(r->type_code);
    free(r);
}

/* ---------- run-merge into NucMut fields ---------- */

/* MSA form: reference panman.cpp:1445-1466 (and :1625-1646) + NucMut ctor panman.hpp:109-151.
 * Input: one node's records in ascending position (n of them). Output arrays sized >= n.
 * Returns the number of NucMut entries; entry i has nucPosition, mutInfo = (len<<4)+type,
 * nucs = sum code_k << (4*(5-k)); primaryBlockId 0, secondaryBlockId -1, nucGapPosition -1. */
int64_t orc_merge_msa(int64_t n, const int32_t* pos, const uint8_t* type_code, int32_t* nuc_position, uint8_t* mut_info,
                      uint32_t* nucs) {
    if (n == 0) return 0; /* the reference only ever merges nodes that own at least one tuple */
    int64_t out = 0, start = 0;
    for (int64_t i = 1; i <= n; i++) {
        int split = (i == n) || (i - start == 6) || (pos[i] != pos[i - 1] + 1) ||
                    ((type_code[i] >> 4) != (type_code[i - 1] >> 4)); /* :1451 */
        if (!split) continue;
        uint32_t packed = 0;
        for (int64_t k = start; k < i; k++) packed += (uint32_t)(type_code[k] & 0xF) << (4 * (5 - (k - start))); /* hpp:271-273 */
        nuc_position[out] = pos[start];
        mut_info[out] = (uint8_t)(((i - start) << 4) + (type_code[start] >> 4));
        nucs[out] = packed;
        out++;
        start = i;
    }
    return out;
}

/* PanGraph form: reference panman.cpp:1236-1272 + NucMut ctor panman.hpp:154-189.
 * Records are 6-tuples (block, -1, pos, gapPos, type, code) sorted lexicographically; gap = 0 merges the
 * non-gap list (adjacent = same block, pos+1, same type), gap = 1 the gap list (same block, same pos,
 * gapPos+1, same type). Outputs per entry: block, nucPosition, nucGapPosition, mutInfo, nucs. */
int64_t orc_merge_pangraph(int gap, int64_t n, const int32_t* block, const int32_t* pos, const int32_t* gap_pos,
                           const uint8_t* type_code, int32_t* o_block, int32_t* o_pos, int32_t* o_gap, uint8_t* mut_info,
                           uint32_t* nucs) {
    if (n == 0) return 0;
    int64_t out = 0, start = 0;
    for (int64_t i = 1; i <= n; i++) {
        int split;
        if (i == n) split = 1;
        else if (!gap)
            split = (i - start == 6) || block[i] != block[i - 1] || pos[i] != pos[i - 1] + 1 ||
                    (type_code[i] >> 4) != (type_code[i - 1] >> 4); /* :1242 */
        else
            split = (i - start == 6) || block[i] != block[i - 1] || pos[i] != pos[i - 1] ||
                    gap_pos[i] != gap_pos[i - 1] + 1 || (type_code[i] >> 4) != (type_code[i - 1] >> 4); /* :1261 */
        if (!split) continue;
        uint32_t packed = 0;
        for (int64_t k = start; k < i; k++) packed += (uint32_t)(type_code[k] & 0xF) << (4 * (5 - (k - start)));
        o_block[out] = block[start];
        o_pos[out] = pos[start];
        o_gap[out] = gap_pos[start];
        mut_info[out] = (uint8_t)(((i - start) << 4) + (type_code[start] >> 4));
        nucs[out] = packed;
        out++;
        start = i;
    }
    return out;
}
