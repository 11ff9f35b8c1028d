// TEST INFRASTRUCTURE ONLY -- C-ABI driver around the reference's own Fitch/Sankoff code.
//
// oracle/Makefile compiles the reference's src/fitchSankoff.cpp VERBATIM (from
// /root/reference, never copied into this repo) against oracle/ref_shim/panmanUtils.hpp
// and links it with this file into oracle/_ref/libpanman_ref.so. This file is OUR code:
// it builds a panmanUtils::Tree from flat arrays, and restates -- in its own words but
// keeping the reference's string-keyed unordered_maps, because that is where the
// reference spends its time -- the per-column caller loops:
//   * MSA / Fitch            reference src/panman.cpp:1381-1435
//   * MSA / Sankoff          reference src/panman.cpp:1568-1613   (--low-mem-mode)
//   * per-node sort          reference src/panman.cpp:1445-1448, 1625-1629
//   * generic one-column triples used by PanGraph/reroot callers
//                             reference src/panman.cpp:873-963, 1048-1232; src/reroot.cpp:54-224
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load the resulting library. The product (libpanman_b200) never does.
#include "panmanUtils.hpp"

#include <pthread.h>

#include <atomic>
#include <cstring>
#include <map>
#include <mutex>
#include <tuple>

// reference src/panman.cpp:41-76 (getNucleotideFromCode), restated as a table.
// Codes follow reference src/panman.hpp:27-44 (A1 C2 M3 G4 R5 S6 V7 T8 W9 Y10 H11 K12 D13 B14 N15).
char panmanUtils::getNucleotideFromCode(int code) {
    static const char table[17] = "-ACMGRSVTWYHKDBN";
    return (code >= 1 && code <= 15) ? table[code] : '-';
}

namespace {

// reference src/panman.cpp:78-113 (getCodeFromNucleotide): anything unlisted -> 0.
inline int codeOf(char c) {
    switch (c) {
    case 'A': return 1;  case 'C': return 2;  case 'G': return 4;  case 'T': return 8;
    case 'R': return 5;  case 'Y': return 10; case 'S': return 6;  case 'W': return 9;
    case 'K': return 12; case 'M': return 3;  case 'B': return 14; case 'D': return 13;
    case 'H': return 11; case 'V': return 7;  case 'N': return 15;
    default:  return 0;
    }
}

using Tuple3 = std::tuple<int, int8_t, int8_t>;  // (pos, type, code) -- reference panman.cpp:1363

struct RefTree {
    panmanUtils::Tree tree;
    std::vector<panmanUtils::Node> nodes;
    std::vector<std::vector<Tuple3>> muts;  // per node, after a batch run
    std::unordered_map<std::string, int> index;  // identifier -> node index
};

// ---- a tiny pthread parallel-for with big stacks (the reference recurses to tree depth) ----
struct PforCtx {
    std::atomic<int64_t>* next;
    int64_t n;
    int64_t grain;
    void (*fn)(int64_t, void*);
    void* arg;
};

void* pforWorker(void* p) {
    PforCtx* c = static_cast<PforCtx*>(p);
    for (;;) {
        int64_t b = c->next->fetch_add(c->grain);
        if (b >= c->n) break;
        int64_t e = std::min(c->n, b + c->grain);
        for (int64_t i = b; i < e; i++) c->fn(i, c->arg);
    }
    return nullptr;
}

void parallelFor(int64_t n, int nThreads, void (*fn)(int64_t, void*), void* arg) {
    if (nThreads < 1) nThreads = 1;
    std::atomic<int64_t> next(0);
    PforCtx ctx{&next, n, 1, fn, arg};
    if (n > int64_t(nThreads) * 64) ctx.grain = 8;
    std::vector<pthread_t> th(nThreads);
    pthread_attr_t attr;
    pthread_attr_init(&attr);
    pthread_attr_setstacksize(&attr, size_t(1) << 30);  // 1 GiB virtual; touched lazily
    for (int t = 0; t < nThreads; t++) pthread_create(&th[t], &attr, pforWorker, &ctx);
    for (int t = 0; t < nThreads; t++) pthread_join(th[t], nullptr);
    pthread_attr_destroy(&attr);
}

struct MsaJob {
    RefTree* rt;
    const std::map<std::string, std::string>* seqs;
    const std::string* consensus;
    std::string reference;
    int64_t nCols;
    uint8_t* outStates;  // n_nodes x nCols or null
    std::vector<std::mutex>* nodeMutex;
    std::unordered_map<std::string, int>* nodeIndex;
    bool sankoff;
    std::atomic<int> error{0};
};

// Walk exactly as nucFitchAssignMutations does (state 0 => stop) and record assigned codes.
void recordFitchStates(panmanUtils::Node* n, std::unordered_map<std::string, int>& states,
                       std::unordered_map<std::string, int>& nodeIndex, uint8_t* out, int64_t stride,
                       int64_t col) {
    std::vector<panmanUtils::Node*> stack{n};
    while (!stack.empty()) {
        panmanUtils::Node* v = stack.back();
        stack.pop_back();
        int s = states[v->identifier];
        if (s == 0) continue;
        int code = 0;
        while ((s >> code) > 1) code++;
        out[int64_t(nodeIndex[v->identifier]) * stride + col] = uint8_t(code);
        for (auto c : v->children) stack.push_back(c);
    }
}

void msaFitchColumn(int64_t i, void* p) {
    MsaJob& job = *static_cast<MsaJob*>(p);
    panmanUtils::Tree& T = job.rt->tree;
    // leaf states: 1 << code, '-' => 1                      (reference panman.cpp:1407-1417)
    std::unordered_map<std::string, int> states;
    std::unordered_map<std::string, std::pair<panmanUtils::NucMutationType, char>> mutations;
    for (const auto& u : *job.seqs) {
        char ch = u.second[i];
        states.insert({u.first, ch != '-' ? (1 << codeOf(ch)) : 1});
    }
    // refState only with --reference                         (reference panman.cpp:1419)
    int refState = -1;
    if (!job.reference.empty()) refState = 1 << codeOf(job.seqs->at(job.reference)[i]);
    int consState = 1 << codeOf((*job.consensus)[i]);
    T.nucFitchForwardPass(T.root, states, refState);          // :1420
    T.nucFitchBackwardPass(T.root, states, consState);        // :1424
    T.nucFitchAssignMutations(T.root, states, mutations, consState);  // :1425
    for (const auto& m : mutations) {                         // :1426-1430
        int idx = (*job.nodeIndex)[m.first];
        std::lock_guard<std::mutex> g((*job.nodeMutex)[idx]);
        job.rt->muts[idx].push_back(Tuple3(int(i), int8_t(m.second.first), int8_t(codeOf(m.second.second))));
    }
    if (job.outStates) recordFitchStates(T.root, states, *job.nodeIndex, job.outStates, job.nCols, i);
}

void msaSankoffColumn(int64_t i, void* p) {
    MsaJob& job = *static_cast<MsaJob*>(p);
    panmanUtils::Tree& T = job.rt->tree;
    std::unordered_map<std::string, std::vector<int>> stateSets;
    std::unordered_map<std::string, int> states;
    std::unordered_map<std::string, std::pair<panmanUtils::NucMutationType, char>> mutations;
    for (const auto& u : *job.seqs) {                         // reference panman.cpp:1574-1582
        std::vector<int> v(16, SANKOFF_INF);
        char ch = u.second[i];
        v[ch != '-' ? codeOf(ch) : 0] = 0;
        stateSets[u.first] = v;
    }
    int defaultState = -1;                                    // :1583-1596
    if (!job.reference.empty()) {
        auto it = job.seqs->find(job.reference);
        if (it == job.seqs->end()) { job.error = 1; return; }
        char ch = it->second[i];
        defaultState = ch != '-' ? codeOf(ch) : 0;
    }
    int consCode = codeOf((*job.consensus)[i]);
    T.nucSankoffForwardPass(T.root, stateSets);               // :1598
    if (defaultState != -1) T.nucSankoffBackwardPass(T.root, stateSets, states, consCode, defaultState);
    else                    T.nucSankoffBackwardPass(T.root, stateSets, states, consCode);   // :1600-1604
    T.nucSankoffAssignMutations(T.root, states, mutations, consCode);  // :1606
    for (const auto& m : mutations) {                         // :1607-1611
        int idx = (*job.nodeIndex)[m.first];
        std::lock_guard<std::mutex> g((*job.nodeMutex)[idx]);
        job.rt->muts[idx].push_back(Tuple3(int(i), int8_t(m.second.first), int8_t(codeOf(m.second.second))));
    }
    if (job.outStates) {
        for (const auto& s : states)
            if (s.second != -1)
                job.outStates[int64_t((*job.nodeIndex)[s.first]) * job.nCols + i] = uint8_t(s.second);
    }
}

struct BigStackCall {
    void (*fn)(void*);
    void* arg;
};
void* bigStackTramp(void* p) {
    BigStackCall* c = static_cast<BigStackCall*>(p);
    c->fn(c->arg);
    return nullptr;
}
void callOnBigStack(void (*fn)(void*), void* arg) {
    pthread_t th;
    pthread_attr_t attr;
    pthread_attr_init(&attr);
    pthread_attr_setstacksize(&attr, size_t(1) << 30);
    BigStackCall c{fn, arg};
    pthread_create(&th, &attr, bigStackTramp, &c);
    pthread_join(th, nullptr);
    pthread_attr_destroy(&attr);
}

struct ColumnCall {
    RefTree* rt;
    int mode;
    const int* leafVal;
    int fwdRef, parentState, defaultState;
    int *outFwd, *outFinal, *outMutType, *outMutArg;
    int rc;
};

void columnCallBody(void* p) {
    ColumnCall& c = *static_cast<ColumnCall*>(p);
    RefTree* rt = c.rt;
    panmanUtils::Tree& T = rt->tree;
    const int n = int(rt->nodes.size());
    const int NONE_DEFAULT = 1 << 28;
    for (int v = 0; v < n; v++) { c.outMutType[v] = -1; c.outMutArg[v] = 0; }
    if (c.mode == 0 || c.mode == 2) {
        std::unordered_map<std::string, int> states;
        for (int v = 0; v < n; v++)
            if (rt->nodes[v].children.empty() && c.leafVal[v] >= 0) states[rt->nodes[v].identifier] = c.leafVal[v];
        if (c.mode == 0) {
            std::unordered_map<std::string, std::pair<panmanUtils::NucMutationType, char>> muts;
            T.nucFitchForwardPass(T.root, states, c.fwdRef);
            if (c.outFwd) for (int v = 0; v < n; v++) c.outFwd[v] = states[rt->nodes[v].identifier];
            if (c.defaultState != NONE_DEFAULT) T.nucFitchBackwardPass(T.root, states, c.parentState, c.defaultState);
            else                                T.nucFitchBackwardPass(T.root, states, c.parentState);
            T.nucFitchAssignMutations(T.root, states, muts, c.parentState);
            for (auto& m : muts) {
                int idx = rt->index[m.first];
                c.outMutType[idx] = int(m.second.first);
                c.outMutArg[idx] = int(m.second.second);
            }
        } else {
            std::unordered_map<std::string, std::pair<panmanUtils::BlockMutationType, bool>> muts;
            T.blockFitchForwardPassNew(T.root, states);
            if (c.outFwd) for (int v = 0; v < n; v++) c.outFwd[v] = states[rt->nodes[v].identifier];
            if (c.defaultState != NONE_DEFAULT) T.blockFitchBackwardPassNew(T.root, states, c.parentState, c.defaultState);
            else                                T.blockFitchBackwardPassNew(T.root, states, c.parentState);
            T.blockFitchAssignMutationsNew(T.root, states, muts, c.parentState);
            for (auto& m : muts) {
                int idx = rt->index[m.first];
                c.outMutType[idx] = int(m.second.first);
                c.outMutArg[idx] = int(m.second.second);
            }
        }
        // assigned states: follow the assign pass' own traversal (state 0 => subtree not visited)
        for (int v = 0; v < n; v++) c.outFinal[v] = 0;
        std::vector<panmanUtils::Node*> stack{T.root};
        while (!stack.empty()) {
            panmanUtils::Node* v = stack.back();
            stack.pop_back();
            int s = states[v->identifier];
            if (s == 0) continue;
            c.outFinal[int(v - rt->nodes.data())] = s;
            for (auto ch : v->children) stack.push_back(ch);
        }
    } else {
        const int W = (c.mode == 1) ? 16 : 3;
        std::unordered_map<std::string, std::vector<int>> sets;
        std::unordered_map<std::string, int> states;
        for (int v = 0; v < n; v++)
            if (rt->nodes[v].children.empty() && c.leafVal[v] >= 0) {
                std::vector<int> vec(W, SANKOFF_INF);
                vec[c.leafVal[v]] = 0;
                sets[rt->nodes[v].identifier] = vec;
            }
        for (int v = 0; v < n; v++) c.outFinal[v] = -1;
        if (c.mode == 1) {
            std::unordered_map<std::string, std::pair<panmanUtils::NucMutationType, char>> muts;
            T.nucSankoffForwardPass(T.root, sets);
            if (c.outFwd) for (int v = 0; v < n; v++) {
                auto& vec = sets[rt->nodes[v].identifier];
                for (int k = 0; k < W; k++) c.outFwd[v * W + k] = vec[k];
            }
            if (c.defaultState == NONE_DEFAULT) {
                // the reference asserts a finite root minimum (fitchSankoff.cpp:505); report instead of aborting
                auto& rv = sets[T.root->identifier];
                bool any = false;
                for (int k = 0; k < W; k++) any = any || rv[k] < SANKOFF_INF;
                if (!any) { c.rc = -2; return; }
                T.nucSankoffBackwardPass(T.root, sets, states, c.parentState);
            } else {
                T.nucSankoffBackwardPass(T.root, sets, states, c.parentState, c.defaultState);
            }
            T.nucSankoffAssignMutations(T.root, states, muts, c.parentState);
            for (auto& m : muts) {
                int idx = rt->index[m.first];
                c.outMutType[idx] = int(m.second.first);
                c.outMutArg[idx] = int(m.second.second);
            }
        } else {
            std::unordered_map<std::string, std::pair<panmanUtils::BlockMutationType, bool>> muts;
            T.blockSankoffForwardPass(T.root, sets);
            if (c.outFwd) for (int v = 0; v < n; v++) {
                auto& vec = sets[rt->nodes[v].identifier];
                for (int k = 0; k < W; k++) c.outFwd[v * W + k] = vec[k];
            }
            if (c.defaultState != NONE_DEFAULT) T.blockSankoffBackwardPass(T.root, sets, states, c.parentState, c.defaultState);
            else                                T.blockSankoffBackwardPass(T.root, sets, states, c.parentState);
            T.blockSankoffAssignMutations(T.root, states, muts, c.parentState);
            for (auto& m : muts) {
                int idx = rt->index[m.first];
                c.outMutType[idx] = int(m.second.first);
                c.outMutArg[idx] = int(m.second.second);
            }
        }
        // assigned states: the assign passes stop at -1; mirror that traversal
        std::vector<panmanUtils::Node*> stack{T.root};
        while (!stack.empty()) {
            panmanUtils::Node* v = stack.back();
            stack.pop_back();
            auto it = states.find(v->identifier);
            if (it == states.end() || it->second == -1) continue;
            c.outFinal[int(v - rt->nodes.data())] = it->second;
            for (auto ch : v->children) stack.push_back(ch);
        }
    }
    c.rc = 0;
}

struct MsaCall {
    MsaJob* job;
    int nThreads;
};
void msaCallBody(void* p) {
    MsaCall& c = *static_cast<MsaCall*>(p);
    // n_threads == 1 reproduces the shipped serial loop (reference panman.cpp:1381; the
    // tbb::parallel_for on :1380 is commented out). For Sankoff the reference uses
    // tbb::parallel_for (:1568); pthreads over columns stand in for TBB, which is absent here.
    if (c.nThreads <= 1) {
        for (int64_t i = 0; i < c.job->nCols; i++) (c.job->sankoff ? msaSankoffColumn : msaFitchColumn)(i, c.job);
    } else {
        parallelFor(c.job->nCols, c.nThreads, c.job->sankoff ? msaSankoffColumn : msaFitchColumn, c.job);
    }
}

}  // namespace

extern "C" {

void* ref_tree_new(int n_nodes, const char* const* names, const int* parent, const int* child_off,
                   const int* child_idx) {
    RefTree* rt = new RefTree();
    rt->nodes.resize(n_nodes);
    rt->muts.resize(n_nodes);
    for (int v = 0; v < n_nodes; v++) {
        rt->nodes[v].identifier = names[v];
        rt->index[names[v]] = v;
        rt->nodes[v].parent = parent[v] < 0 ? nullptr : &rt->nodes[parent[v]];
        if (parent[v] < 0) rt->tree.root = &rt->nodes[v];
        for (int e = child_off[v]; e < child_off[v + 1]; e++) rt->nodes[v].children.push_back(&rt->nodes[child_idx[e]]);
    }
    return rt;
}

void ref_tree_free(void* h) { delete static_cast<RefTree*>(h); }

// One column through the reference's own call triple.
// mode: 0 nuc Fitch, 1 nuc Sankoff, 2 block Fitch, 3 block Sankoff.
// leaf_val[v] (leaves only): Fitch: state bitmask; Sankoff: state index; <0 => leaf omitted from the map.
// default_state: 1<<28 means "not given". Returns 0, or -2 where the reference would trip its assert.
int ref_column(void* h, int mode, const int* leaf_val, int fwd_ref_state, int parent_state, int default_state,
               int* out_fwd, int* out_final, int* out_mut_type, int* out_mut_arg) {
    ColumnCall c{static_cast<RefTree*>(h), mode, leaf_val, fwd_ref_state, parent_state, default_state,
                 out_fwd, out_final, out_mut_type, out_mut_arg, -1};
    callOnBigStack(columnCallBody, &c);
    return c.rc;
}

// Whole-MSA batch (all columns), results kept in the handle until the next batch call.
// algo: 0 Fitch (reference panman.cpp:1381-1435), 1 Sankoff (:1568-1613).
int ref_msa_run(void* h, int algo, int n_seqs, const char* const* seq_ids, const char* const* seqs,
                int64_t n_cols, const char* consensus, const char* reference_id, int n_threads,
                uint8_t* out_states) {
    RefTree* rt = static_cast<RefTree*>(h);
    std::map<std::string, std::string> seqMap;  // std::map, as the reference (panman.cpp:1280)
    for (int s = 0; s < n_seqs; s++) seqMap[seq_ids[s]] = std::string(seqs[s], size_t(n_cols));
    std::string cons(consensus, size_t(n_cols));
    std::unordered_map<std::string, int> nodeIndex;
    for (size_t v = 0; v < rt->nodes.size(); v++) nodeIndex[rt->nodes[v].identifier] = int(v);
    std::vector<std::mutex> nodeMutex(rt->nodes.size());
    for (auto& m : rt->muts) m.clear();
    if (out_states) std::memset(out_states, 0xFF, rt->nodes.size() * size_t(n_cols));
    MsaJob job;
    job.rt = rt;
    job.seqs = &seqMap;
    job.consensus = &cons;
    job.reference = reference_id ? reference_id : "";
    job.nCols = n_cols;
    job.outStates = out_states;
    job.nodeMutex = &nodeMutex;
    job.nodeIndex = &nodeIndex;
    job.sankoff = (algo == 1);
    MsaCall call{&job, n_threads};
    callOnBigStack(msaCallBody, &call);
    // per-node sort of the tuples                              (reference panman.cpp:1447, 1627)
    for (auto& m : rt->muts) std::sort(m.begin(), m.end());
    return job.error.load() ? -1 : 0;
}

void ref_result_counts(void* h, int64_t* counts) {
    RefTree* rt = static_cast<RefTree*>(h);
    for (size_t v = 0; v < rt->muts.size(); v++) counts[v] = int64_t(rt->muts[v].size());
}

void ref_result_copy(void* h, int32_t* pos, int8_t* type, int8_t* code) {
    RefTree* rt = static_cast<RefTree*>(h);
    int64_t k = 0;
    for (auto& m : rt->muts)
        for (auto& t : m) {
            pos[k] = std::get<0>(t);
            type[k] = std::get<1>(t);
            code[k] = std::get<2>(t);
            k++;
        }
}

}  // extern "C"
