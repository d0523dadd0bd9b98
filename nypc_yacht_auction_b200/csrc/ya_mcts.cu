// Yacht-Auction B200 engine -- batched MCTS (select / expand / backup) over a flat per-game node
// pool in HBM.  One warp per game; thousands of games advance one simulation per launch pair:
//
//   ya_mcts_select   walks root -> leaf with the reference's UCB rule, applies the (stochastic)
//                    transitions, allocates the leaf node, writes its feature row for the batched
//                    evaluator; paths that end in a terminal / dead-end node are backed up at once
//   [evaluator]      ONE batched forward for all leaves (PyTorch), or the uniform prior
//   ya_mcts_expand   masks + renormalises the policy into the leaf's legal-only prior row (numpy's
//                    pairwise float32 summation order), then backs the value up the recorded path
//
// Behavioural source of truth: /root/reference/MCTS.py (cited per function) with the quirks listed
// in SURVEY.md section 8a (Q1-Q9): state-keyed nodes (transpositions merge, tree persists across moves),
// in-search dice, lowest-index tie-break, three leaf fallbacks, un-negated 0 from dead-end revisits,
// and step-wise float32 / Python-double arithmetic decided per value exactly as numpy (NEP 50) does.
#include "ya_common.cuh"
#include "../../include/yacht_b200.h"
#include <math_constants.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace {

constexpr int kNodeWords = 16;        // 64-byte node record
constexpr int kCursorWords = 32;      // 128-byte per-game search cursor
constexpr int kMaxDepth = 16;
constexpr int kMinEdgeCap = 32;
constexpr int kWarpsPerBlock = 4;
#ifndef YA_SELECT_TEAM
#define YA_SELECT_TEAM 8
#endif
constexpr int kSelectTeam = YA_SELECT_TEAM;   // lanes per game in the descent kernels (8: four games per warp)
#ifndef YA_MCTS_MIN_BLOCKS
#define YA_MCTS_MIN_BLOCKS 8          // <= 64 registers: 32 resident warps per SM for the latency-bound tree walks
#endif

// node words
enum { N_KEY = 0, N_DESC = 8, N_VISITS = 9, N_PRIOR = 10, N_EDGES = 11, N_NEDGE = 12, N_KIND = 13, N_PCONST = 14 };
// N_KIND 0: N_PRIOR -> float32 prior row + group maxima.  N_KIND 1 (uniform evaluator): every legal move has the SAME prior
// (bits in N_PCONST), so N_PRIOR -> just a visited bitmask of ceil(L / 32) words: 0.4 KB instead of 12 KB per node.
// cursor words
enum { C_LEAF = 0, C_DEPTH = 8, C_KIND = 9, C_NODE = 10, C_PENDING = 11, C_PATH = 12, C_LEAF_DESC = 28, C_LEAF_ROW = 29 };   // C_PATH: kMaxDepth words
enum { KIND_DONE = 0, KIND_NEED_EVAL = 1, KIND_ERROR = 2, KIND_NEED_DRAW = 3 };
// meta words
enum { M_NODES = 0, M_TOP = 1, M_ROUND = 2 };
// error bits (OR-ed into err_flag)
enum { E_NODES_FULL = 1 << 8, E_ARENA_FULL = 1 << 9, E_DEPTH = 1 << 10, E_RULE = 1 << 11 };

struct View {
    uint32_t* nodes;
    uint16_t* ht;
    uint32_t* arena;
    uint32_t* meta;
    uint32_t* cur;
    int max_nodes, ht_size;
    uint32_t arena_words;
};

__device__ __forceinline__ View make_view(const ya_mcts_tree& t, int64_t g) {
    View v;
    v.nodes = t.nodes + g * (int64_t)t.max_nodes * kNodeWords;
    v.ht = t.ht + g * (int64_t)t.ht_size;
    v.arena = t.arena + g * t.arena_words;
    v.meta = t.meta + g * 4;
    v.cur = t.cursor + g * kCursorWords;
    v.max_nodes = t.max_nodes;
    v.ht_size = t.ht_size;
    v.arena_words = (uint32_t)t.arena_words;
    return v;
}

// A node's prior block in the arena: L float32 priors (sign bit = child visited), padded to a multiple of four
// words, then one word per 32 priors holding 1 + the bit pattern of the largest UNVISITED prior of that group
// (0 = none left).  UCB for an unvisited child, cpuct * P * sqrt(Ns + EPS), is monotone in P, so the argmax over
// up to 3,024 unvisited children is: best group from <= 95 words, then the 32 priors of that one group.
__device__ __forceinline__ int row_words(int L) { return ((L + 3) & ~3) + ((L + 31) >> 5); }
__device__ __forceinline__ int group_max_at(int L) { return (L + 3) & ~3; }

// A node's visited edges: ONE contiguous array in the arena, capacity C = 32, 64, 128, ... (it doubles when full; the
// old copy is abandoned in the bump arena): C x u16 legal index | C x u16 child | C x u32 Nsa (bit 31: Q is a Python
// float, else a numpy float32) | C x f64 Q = 16 bytes per edge.  Every lane can address any edge directly, so a node with
// 100 edges costs one round of independent loads, not a walk over a chunk list.
// child = 1 + the node the edge leads to, when getNextState is deterministic for it (no dice, no tie-break: everything
// except second bids that tie or close round 1, and the score move that ends a round -- which the search never makes,
// quirk Q3); 0 = not known / stochastic.  The child's record holds its canonical state (it IS the node key), so a descent
// through a known child skips the transition, the hash and the table probe: one dependent load instead of three.
__device__ __forceinline__ int edge_cap(int n) { return n <= kMinEdgeCap ? kMinEdgeCap : 1 << (32 - __clz(n - 1)); }
__device__ __forceinline__ int edge_words(int cap) { return 4 * cap; }
struct Edges {
    uint16_t* idx;
    uint16_t* child;
    uint32_t* nsa;
    double* q;
};
__device__ __forceinline__ Edges edges_at(uint32_t* base, int cap) {
    Edges e;
    e.idx = reinterpret_cast<uint16_t*>(base);
    e.child = reinterpret_cast<uint16_t*>(base + cap / 2);
    e.nsa = base + cap;
    e.q = reinterpret_cast<double*>(base + 2 * cap);
    return e;
}

// Row storage of a tree pool (one per pool, fixed by the caller):
//   ROWS_F32    float32 prior rows as above (any float32 policy: ya_mcts_expand; the bit-exact reference path)
//   ROWS_CONST  one prior value per node + a visited bitmask (ya_k_mcts_search_uniform)
//   ROWS_L16F / ROWS_L16B   the leaf's LEGAL policy-head logits as they left the tensor core, 16 bits each (IEEE half /
//               bfloat16), written straight into the row by the forward kernel's epilogue, + one exponent offset per node:
//               P[a] = 2^(l[a] * log2(e) + off), off = -max * log2(e) - log2(sum over legal of 2^((l - max) * log2(e))),
//               evaluated where a prior is needed (one FMA + one EX2, the same two instructions that would have produced
//               a stored float32 prior, so the values are bit-identical to storing them); half the bytes of a float32 row.
//               Layout from N_PRIOR: logit area | ceil(L / 32) group maxima (1 + bits of the largest unvisited P, 0 = none) |
//               ceil(L / 32) words of visited bits.  A leaf whose legal moves all underflowed (MCTS.py:97-101) becomes a
//               constant-prior node (N_KIND 1, P = 1 / L) on the same visited bits.
//               Logit area.  The forward kernel's epilogue owns 16 consecutive policy columns per 32-byte store, so the
//               area keeps every logit at its column's position modulo 16: a bid row is columns [0, 208) as they are; a
//               ten-dice row has one 272-slot block per OPEN category c, a copy of the 16-aligned column window that
//               contains the category's 252 columns [202 + 252 c, +252) -- its logits start pad(c) = (10 + 12 c) mod 16
//               slots into the block, the slots around them hold neighbouring columns and are never read; a five-dice
//               row (one legal move per open category) is compact.  Every store of the epilogue is a full, aligned sector.
enum { ROWS_F32 = 0, ROWS_CONST = 1, ROWS_L16F = 2, ROWS_L16B = 3 };
constexpr int kL16Block = 272;                                         // slots per open category of a ten-dice row
__device__ __forceinline__ int l16_logit_words(uint32_t desc) {
    if (desc >> 13) return __popc((desc >> 1) & 0xFFFu) * (kL16Block / 2);
    return (((ya_legal_count(desc) + 1) >> 1) + 3) & ~3;               // bids: 104 words = columns [0, 208)
}
__device__ __forceinline__ int l16_words(uint32_t desc) { return l16_logit_words(desc) + 2 * ((ya_legal_count(desc) + 31) >> 5); }
__device__ __forceinline__ int l16_pad(int cat) { return (10 + 12 * cat) & 15; }       // (202 + 252 cat) mod 16
// slot of the k-th legal move's logit inside the logit area (16-bit units)
__device__ __forceinline__ int l16_slot(uint32_t desc, int k) {
    if (!(desc >> 13)) return k;
    const int r = k / YA_N_SUBSET, sub = k - r * YA_N_SUBSET;
    const int cat = __fns((desc >> 1) & 0xFFFu, 0, r + 1);
    return kL16Block * r + l16_pad(cat) + sub;
}
template <int ROWS> __device__ __forceinline__ float l16_value(uint32_t h) {         // one 16-bit logit -> float32
    return ROWS == ROWS_L16F ? __half2float(__ushort_as_half((unsigned short)h)) : __uint_as_float(h << 16);
}
constexpr float kLog2e = 1.4426950408889634f;
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int ROWS> __device__ __forceinline__ float l16_prior(uint32_t h, float off) {
    return ex2_approx(fmaf(l16_value<ROWS>(h), kLog2e, off));
}

// writes priors [k0, k0 + 32) of a fresh row (one per lane; lanes with k >= L idle) and the group's maximum
__device__ __forceinline__ void store_prior_group(float* __restrict__ row, int L, int k0, float p, int lane) {
    const int k = k0 + lane;
    uint32_t e = 0;
    if (k < L) { row[k] = p; e = __float_as_uint(p) + 1u; }
    e = __reduce_max_sync(0xFFFFFFFFu, e);
    if (lane == 0) reinterpret_cast<uint32_t*>(row)[group_max_at(L) + (k0 >> 5)] = e;
}

// W lanes cooperate on one game: W = 32 is a warp per game (the expand kernels, whose work is streaming a prior row);
// W = 8 puts four games in a warp (the descent: mostly per-game scalar control flow -- rules, hashing, draws -- that
// every lane of a team executes redundantly, so narrower teams mean fewer instructions per game and four
// independent chains of dependent loads per warp).  All collectives are restricted to the team's lanes.
template <int W>
struct Team {
    int sub;            // this lane's index inside the team
    int shift;          // first lane of the team
    uint32_t mask;      // the team's lanes
    __device__ __forceinline__ static Team make() {
        const int lane = threadIdx.x & 31;
        Team t;
        t.sub = lane & (W - 1);
        t.shift = lane & ~(W - 1);
        t.mask = W == 32 ? 0xFFFFFFFFu : ((W == 32 ? 0u : ((1u << (W & 31)) - 1u)) << t.shift);
        return t;
    }
    __device__ __forceinline__ void sync() const { __syncwarp(mask); }
    __device__ __forceinline__ uint32_t ballot(bool p) const {
        const uint32_t b = __ballot_sync(mask, p) >> shift;
        return W == 32 ? b : (b & ((1u << (W & 31)) - 1u));
    }
    template <class T> __device__ __forceinline__ T bcast(T v, int src_sub) const { return __shfl_sync(mask, v, shift + src_sub); }
    template <class T> __device__ __forceinline__ T xor_(T v, int o) const { return __shfl_xor_sync(mask, v, o); }
    __device__ __forceinline__ uint32_t reduce_max(uint32_t v) const { return __reduce_max_sync(mask, v); }
};

// path word of one level of a descent: node index | legal index of the chosen move << 16 | deterministic transition << 31
__device__ __forceinline__ int path_node(uint32_t pe) { return (int)(pe & 0xFFFFu); }
__device__ __forceinline__ int path_move(uint32_t pe) { return (int)((pe >> 16) & 0xFFFu); }
__device__ __forceinline__ bool path_deterministic(uint32_t pe) { return (pe >> 31) != 0u; }

// A value travelling up the tree with the numeric type Python would give it.
struct Val {
    double d;        // value (exact float32 value when is_f32)
    bool is_f32;
};

__device__ __forceinline__ uint32_t state_hash(const YaState& s) {
    uint32_t h = 0x811C9DC5u;
#pragma unroll
    for (int i = 0; i < 8; ++i) h = (h ^ s.w[i]) * 0x01000193u;
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12;
    return h;
}

// open-addressing lookup (all lanes run it redundantly: uniform addresses broadcast)
__device__ __forceinline__ int ht_find(const View& v, const YaState& s, int* free_slot) {
    int mask = v.ht_size - 1;
    int slot = (int)(state_hash(s) & (uint32_t)mask);
    for (;;) {
        uint32_t e = v.ht[slot];
        if (e == 0) { *free_slot = slot; return -1; }
        const uint4* k = reinterpret_cast<const uint4*>(v.nodes + (int64_t)(e - 1) * kNodeWords);
        uint4 a = k[0], b = k[1];
        if (a.x == s.w[0] && a.y == s.w[1] && a.z == s.w[2] && a.w == s.w[3] &&
            b.x == s.w[4] && b.y == s.w[5] && b.z == s.w[6] && b.w == s.w[7]) return (int)(e - 1);
        slot = (slot + 1) & mask;
    }
}

__device__ __forceinline__ double es_as_double(float es) {       // getGameEnded returns Python floats
    if (es == 1.0f) return 1.0;
    if (es == -1.0f) return -1.0;
    return 1e-4;
}

// ---------------------------------------------------------------- UCB argmax (MCTS.py:117-133)
// ROWS: the row storage of this pool (constant-prior nodes: only ya_k_mcts_search_uniform creates them, and a tree pool is
// driven either by that kernel or by select / expand -- mcts.BatchedMCTS fixes the choice at construction).
// Returns (legal index << 16) | child of the arg-max edge (child = 0 for an unvisited move): ordering packed keys is
// ordering legal indices, so the lowest-index tie-break of MCTS.py:131 is untouched.
template <int ROWS, int W>
__device__ __forceinline__ int ucb_select(const View& v, const uint32_t* node, int L, float cpuct, const Team<W>& tm) {
    const int sub = tm.sub;
    const uint32_t visits = node[N_VISITS];
    const uint32_t* row = v.arena + node[N_PRIOR];
    const float sq_new = (float)sqrt((double)visits + 1e-8);       // math.sqrt(Ns + EPS), rounded when it meets float32
    const float sq_old = (float)sqrt((double)visits);
    float best = -CUDART_INF_F;
    int besti = 0x7FFFFFFF;
    const int n_edges = (int)node[N_NEDGE];
    auto team_argmax = [&] {                                        // (u desc, index asc) over the team
#pragma unroll
        for (int o = W / 2; o; o >>= 1) {
            float ou = tm.xor_(best, o);
            int oi = tm.xor_(besti, o);
            if (ou > best || (ou == best && oi < besti)) { best = ou; besti = oi; }
        }
    };
    if (ROWS != ROWS_F32 && (ROWS == ROWS_CONST || node[N_KIND] == 1)) {
        // constant prior p: every unvisited child has the same u, so the lowest unvisited index wins among them
        const float cp = __fmul_rn(cpuct, __uint_as_float(node[N_PCONST]));
        if (n_edges > 0) {
            const Edges ed = edges_at(v.arena + node[N_EDGES], edge_cap(n_edges));
            for (int e = sub; e < n_edges; e += W) {
                const int ai = ((int)ed.idx[e] << 16) | (int)ed.child[e];
                float x = __fdiv_rn(__fmul_rn(cp, sq_old), (float)(1u + (ed.nsa[e] & 0x7FFFFFFFu)));
                float u = __fadd_rn((float)ed.q[e], x);
                if (u > best || (u == best && ai < besti)) { best = u; besti = ai; }
            }
        }
        int first = 0x7FFFFFFF;
        for (int b = sub; b < ((L + 31) >> 5) && first == 0x7FFFFFFF; b += W) {
            const int nbits = min(32, L - 32 * b);
            const uint32_t open = ~row[b] & (nbits == 32 ? 0xFFFFFFFFu : ((1u << nbits) - 1u));
            if (open) first = 32 * b + __ffs(open) - 1;
        }
#pragma unroll
        for (int o = W / 2; o; o >>= 1) first = min(first, tm.xor_(first, o));
        if (first != 0x7FFFFFFF) {
            const float u = __fmul_rn(cp, sq_new);
            if (u > best || (u == best && (first << 16) < besti)) { best = u; besti = first << 16; }
        }
        team_argmax();
        return besti;
    }
    if (ROWS >= ROWS_L16F) {
        // 16-bit logit row: the same arithmetic on P = 2^(l * log2(e) + off), evaluated on the fly
        const uint32_t desc = node[N_DESC];
        const int lw = l16_logit_words(desc), nb = (L + 31) >> 5;
        const uint16_t* lg = reinterpret_cast<const uint16_t*>(row);
        const float off = __uint_as_float(node[N_PCONST]);
        // the group maxima (<= 95 words) and visited bits are needed after the visited-edge loop, whose loads depend on one
        // another (edge -> logit): start pulling their lines now (one 128-byte line per lane that has one)
        if (sub * 32 < 2 * nb) asm volatile("prefetch.global.L1 [%0];" ::"l"(row + lw + sub * 32));
        if (n_edges > 0) {
            const Edges ed = edges_at(v.arena + node[N_EDGES], edge_cap(n_edges));
            for (int e0 = 0; e0 < n_edges; e0 += 2 * W) {
                int ai[2];
                uint32_t nsa[2], lb[2];
                double q[2];
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    const int e = e0 + W * t + sub;
                    const bool ok = e < n_edges;
                    ai[t] = ok ? ((int)ed.idx[e] << 16) | (int)ed.child[e] : 0;
                    nsa[t] = ok ? ed.nsa[e] & 0x7FFFFFFFu : 0u;
                    q[t] = ok ? ed.q[e] : 0.0;
                }
#pragma unroll
                for (int t = 0; t < 2; ++t) lb[t] = lg[l16_slot(desc, ai[t] >> 16)];
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    if (e0 + W * t + sub < n_edges) {
                        float x = __fmul_rn(__fmul_rn(cpuct, l16_prior<ROWS>(lb[t], off)), sq_old);
                        x = __fdiv_rn(x, (float)(1u + nsa[t]));
                        float u = __fadd_rn((float)q[t], x);
                        if (u > best || (u == best && ai[t] < besti)) { best = u; besti = ai[t]; }
                    }
                }
            }
        }
        const uint32_t* gmax = row + lw;
        float gu = -CUDART_INF_F;
        int gb = 0x7FFFFFFF;
        for (int b = sub; b < nb; b += W) {
            uint32_t e = gmax[b];
            if (e) {
                float u = __fmul_rn(__fmul_rn(cpuct, __uint_as_float(e - 1u)), sq_new);
                if (u > gu) { gu = u; gb = b; }
            }
        }
#pragma unroll
        for (int o = W / 2; o; o >>= 1) {
            float ou = tm.xor_(gu, o);
            int ob = tm.xor_(gb, o);
            if (ou > gu || (ou == gu && ob < gb)) { gu = ou; gb = ob; }
        }
        if (gb != 0x7FFFFFFF) {
            const uint32_t seen = row[lw + nb + gb];
            constexpr int kPer = 32 / W;                                // consecutive logits of the group per lane
            const int i0 = (gb << 5) + sub * kPer;
            auto consider = [&](int i, uint32_t h) {
                if (i < L && !((seen >> (i & 31)) & 1u)) {
                    float u = __fmul_rn(__fmul_rn(cpuct, l16_prior<ROWS>(h, off)), sq_new);
                    if (u > best || (u == best && (i << 16) < besti)) { best = u; besti = i << 16; }
                }
            };
            if (kPer == 1) {
                if (i0 < L) consider(i0, lg[l16_slot(desc, i0)]);
            } else if (i0 < L) {
                // kPer (2, 4 or 8) consecutive legal moves starting at a multiple of kPer: 252 is a multiple of 4, so for
                // kPer <= 4 they lie in one category block at consecutive, even-aligned slots
                static_assert(kPer == 1 || kPer == 2 || kPer == 4, "a lane's moves must not straddle a category block");
                const int s0 = l16_slot(desc, i0);
#pragma unroll
                for (int t = 0; t < kPer / 2; ++t) {
                    const int i = i0 + 2 * t;
                    if (i < L) {                                        // the pair's word lies inside the logit area
                        const uint32_t w2 = row[(s0 >> 1) + t];
                        consider(i, w2 & 0xFFFFu);
                        consider(i + 1, w2 >> 16);
                    }
                }
            }
        }
        team_argmax();
        return besti;
    }
    // visited edges: u = Q + cpuct * P * sqrt(Ns) / (1 + Nsa); two edges per lane in flight
    if (n_edges > 0) {
        const Edges ed = edges_at(v.arena + node[N_EDGES], edge_cap(n_edges));
        for (int e0 = 0; e0 < n_edges; e0 += 2 * W) {
            int ai[2];
            uint32_t nsa[2], pb[2];
            double q[2];
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const int e = e0 + W * t + sub;
                const bool ok = e < n_edges;
                ai[t] = ok ? ((int)ed.idx[e] << 16) | (int)ed.child[e] : 0;
                nsa[t] = ok ? ed.nsa[e] & 0x7FFFFFFFu : 0u;
                q[t] = ok ? ed.q[e] : 0.0;
            }
#pragma unroll
            for (int t = 0; t < 2; ++t) pb[t] = row[ai[t] >> 16];
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                if (e0 + W * t + sub < n_edges) {
                    float p = __uint_as_float(pb[t] & 0x7FFFFFFFu);
                    float x = __fmul_rn(__fmul_rn(cpuct, p), sq_old);
                    x = __fdiv_rn(x, (float)(1u + nsa[t]));
                    float u = __fadd_rn((float)q[t], x);
                    if (u > best || (u == best && ai[t] < besti)) { best = u; besti = ai[t]; }
                }
            }
        }
    }
    // unvisited edges: u = cpuct * P * sqrt(Ns + EPS), float32.  u is monotone (non-strictly) in P, so the
    // lowest-index maximiser lives in the lowest group whose largest unvisited prior reaches the maximal u.
    {
        const uint32_t* gmax = row + group_max_at(L);
        const int ngroups = (L + 31) >> 5;
        float gu = -CUDART_INF_F;
        int gb = 0x7FFFFFFF;
        for (int b = sub; b < ngroups; b += W) {
            uint32_t e = gmax[b];
            if (e) {
                float u = __fmul_rn(__fmul_rn(cpuct, __uint_as_float(e - 1u)), sq_new);
                if (u > gu) { gu = u; gb = b; }                     // b ascends: ties keep the lower group
            }
        }
#pragma unroll
        for (int o = W / 2; o; o >>= 1) {
            float ou = tm.xor_(gu, o);
            int ob = tm.xor_(gb, o);
            if (ou > gu || (ou == gu && ob < gb)) { gu = ou; gb = ob; }
        }
        if (gb != 0x7FFFFFFF) {
#pragma unroll
            for (int t = 0; t < 32 / W; ++t) {                      // the 32 priors of that group
                int i = (gb << 5) + t * W + sub;
                if (i < L) {
                    uint32_t bits = row[i];
                    if (!(bits >> 31)) {
                        float u = __fmul_rn(__fmul_rn(cpuct, __uint_as_float(bits)), sq_new);
                        if (u > best || (u == best && (i << 16) < besti)) { best = u; besti = i << 16; }
                    }
                }
            }
        }
    }
    team_argmax();
    return besti;
}

// ---------------------------------------------------------------- backup (MCTS.py:152-164)
// Returns false if the arena overflowed.
template <int ROWS, int W>
__device__ __forceinline__ bool backup_edge(const View& v, uint32_t* node, int ai, const Val& val, uint32_t& arena_top,
                                            const Team<W>& tm, uint32_t child) {
    const int sub = tm.sub;
    int failed = 0, fresh = 0;
    // locate the edge
    const int n_edges = (int)node[N_NEDGE];
    const int cap = edge_cap(n_edges);
    uint32_t* base = v.arena + node[N_EDGES];
    int found = -1;
    if (n_edges > 0) {
        const Edges ed = edges_at(base, cap);
        for (int e0 = 0; e0 < n_edges && found < 0; e0 += 2 * W) {
            const int ea = e0 + sub, eb = e0 + W + sub;
            const int ia = ea < n_edges ? (int)ed.idx[ea] : -1, ib = eb < n_edges ? (int)ed.idx[eb] : -1;
            const uint32_t ma = tm.ballot(ia == ai), mb = tm.ballot(ib == ai);
            if (ma) found = e0 + __ffs(ma) - 1;
            else if (mb) found = e0 + W + __ffs(mb) - 1;
        }
    }
    // a first visit appends; a full array moves to one of twice the capacity first (all lanes copy)
    if (found < 0 && (n_edges == 0 || n_edges == cap)) {
        const int new_cap = n_edges == 0 ? kMinEdgeCap : 2 * cap;
        const uint32_t top = (arena_top + 1u) & ~1u;                 // 8-byte aligned: the Q values are doubles
        if (top + (uint32_t)edge_words(new_cap) > v.arena_words) {
            failed = 1;
        } else {
            uint32_t* nb = v.arena + top;
            if (n_edges > 0) {
                const Edges from = edges_at(base, cap), to = edges_at(nb, new_cap);
                for (int e = sub; e < n_edges; e += W) {
                    to.idx[e] = from.idx[e]; to.child[e] = from.child[e]; to.nsa[e] = from.nsa[e]; to.q[e] = from.q[e];
                }
            }
            tm.sync();
            if (sub == 0) node[N_EDGES] = top;
            arena_top = top + (uint32_t)edge_words(new_cap);
            base = nb;
        }
    }
    if (sub == 0) {
        if (found >= 0) {
            const Edges ed = edges_at(base, cap);
            uint32_t raw = ed.nsa[found];
            uint32_t nsa = raw & 0x7FFFFFFFu;
            bool q_f32 = (raw >> 31) == 0;
            double q = ed.q[found];
            if (!q_f32 && !val.is_f32) {
                // Python floats all the way: (Nsa * Q + v) / (Nsa + 1) in double
                q = __ddiv_rn(__dadd_rn(__dmul_rn((double)nsa, q), val.d), (double)(nsa + 1u));
            } else {
                // numpy float32 arithmetic; a Python-float operand is rounded to float32 first
                float prod = q_f32 ? __fmul_rn((float)nsa, (float)q) : (float)__dmul_rn((double)nsa, q);
                float sum = __fadd_rn(prod, (float)val.d);
                q = (double)__fdiv_rn(sum, (float)(nsa + 1u));
                q_f32 = true;
            }
            ed.q[found] = q;
            ed.nsa[found] = (nsa + 1u) | (q_f32 ? 0u : 0x80000000u);
            if (child && ed.child[found] == 0) ed.child[found] = (uint16_t)child;
        } else if (!failed) {
            const Edges ed = edges_at(base, edge_cap(n_edges + 1));
            ed.idx[n_edges] = (uint16_t)ai;
            ed.child[n_edges] = (uint16_t)child;
            ed.nsa[n_edges] = 1u | (val.is_f32 ? 0u : 0x80000000u);        // Qsa = v, Nsa = 1
            ed.q[n_edges] = val.d;
            node[N_NEDGE] = (uint32_t)(n_edges + 1);
            if (ROWS != ROWS_F32 && (ROWS == ROWS_CONST || node[N_KIND] == 1)) {
                v.arena[node[N_PRIOR] + (ai >> 5)] |= 1u << (ai & 31);      // constant-prior node: visited bitmask
            } else if (ROWS >= ROWS_L16F) {
                const int Ln = ya_legal_count(node[N_DESC]);
                v.arena[node[N_PRIOR] + l16_logit_words(node[N_DESC]) + ((Ln + 31) >> 5) + (ai >> 5)] |= 1u << (ai & 31);
                fresh = 1;
            } else {
                v.arena[node[N_PRIOR] + ai] |= 0x80000000u;                 // mark the prior entry as visited
                fresh = 1;
            }
        }
        node[N_VISITS] += 1;                                         // Ns[s] += 1
    }
    fresh = tm.bcast(fresh, 0);
    tm.sync();
    if (fresh) {                                                     // the child's group lost an unvisited prior
        const int L = ya_legal_count(node[N_DESC]);
        uint32_t* row = v.arena + node[N_PRIOR];
        uint32_t e = 0;
        if (ROWS >= ROWS_L16F) {
            const uint32_t desc = node[N_DESC];
            const int lw = l16_logit_words(desc), nb = (L + 31) >> 5;
            const uint32_t seen = row[lw + nb + (ai >> 5)];
            const float off = __uint_as_float(node[N_PCONST]);
            const uint16_t* lg = reinterpret_cast<const uint16_t*>(row);
            float m = -CUDART_INF_F;                                 // the group's largest unvisited logit; its maximum is P(that)
#pragma unroll
            for (int t = 0; t < 32 / W; ++t) {
                const int i = (ai & ~31) + t * W + sub;
                if (i < L && !((seen >> (i & 31)) & 1u)) m = fmaxf(m, l16_value<ROWS>(lg[l16_slot(desc, i)]));
            }
#pragma unroll
            for (int o = W / 2; o; o >>= 1) m = fmaxf(m, tm.xor_(m, o));
            e = m == -CUDART_INF_F ? 0u : __float_as_uint(ex2_approx(fmaf(m, kLog2e, off))) + 1u;
            if (sub == 0) row[lw + (ai >> 5)] = e;
        } else {
#pragma unroll
            for (int t = 0; t < 32 / W; ++t) {
                const int i = (ai & ~31) + t * W + sub;
                if (i < L) { uint32_t b = row[i]; if (!(b >> 31)) e = max(e, b + 1u); }
            }
            e = tm.reduce_max(e);
            if (sub == 0) row[group_max_at(L) + (ai >> 5)] = e;
        }
        tm.sync();
    }
    return !failed;
}

// last_node: the node the deepest edge of the path leads to (the new leaf, or a dead-end node), -1 if it ended in a
// terminal state (terminal states have no node).
template <int ROWS, int W>
__device__ __forceinline__ bool backup_path(const View& v, int depth, Val ret, uint32_t& arena_top, const Team<W>& tm,
                                            int last_node) {
    int below = last_node;                                           // node one level further down the path
    for (int d = depth - 1; d >= 0; --d) {
        uint32_t pe = v.cur[C_PATH + d];
        uint32_t* node = v.nodes + (int64_t)path_node(pe) * kNodeWords;
        const uint32_t child = (path_deterministic(pe) && below >= 0) ? (uint32_t)below + 1u : 0u;
        if (!backup_edge<ROWS>(v, node, path_move(pe), ret, arena_top, tm, child)) return false;
        below = path_node(pe);
        ret.d = -ret.d;                                              // return -v
    }
    return true;
}

// ---------------------------------------------------------------- select (MCTS.py:56-150)
struct Walk {
    uint32_t node_count, arena_top, pending;
    int depth, kind, err, leaf_node;
    uint32_t leaf_desc, leaf_row;        // mask descriptor and prior-row offset of the new leaf (for the expand kernel)
    Val ret;
    YaState leaf;
};

// Lazy pruning at round boundaries: inside rounds >= 2 the search never leaves the root's round
// (quirk Q3), so nothing stored for an earlier round >= 2 can be reached again.
template <int W>
__device__ __forceinline__ void prune_on_new_round(const View& v, const YaState& root, Walk& w, const Team<W>& tm) {
    uint32_t old_round = v.meta[M_ROUND], new_round = (uint32_t)ya_round(root);
    if (old_round == new_round) return;
    if (old_round >= 2 || new_round < old_round) {
        for (int i = tm.sub; i < v.ht_size / 2; i += W) reinterpret_cast<uint32_t*>(v.ht)[i] = 0u;
        w.node_count = 0;
        w.arena_top = 4;
    }
    tm.sync();                                                       // every lane has read the old round
    if (tm.sub == 0) v.meta[M_ROUND] = new_round;
    tm.sync();
}

// One descent from the canonical root.  On return: kind == KIND_NEED_EVAL (leaf allocated, w.leaf holds its
// state), KIND_DONE (terminal / dead end: w.ret is the value returned by the deepest call) or KIND_ERROR.
// Dice for a transition inside search: Philox (batched engine) or injected by the host (drop-in MCTS: the
// reference rolls in-search dice from the global numpy / random streams, yacht/YachtGame.py:154-159).
struct DrawSource {
    const uint8_t* inj;      // NULL = Philox; else {tie, rollA[5], rollB[5], valid bits (1 tie, 2 rolls)}
    int resume;              // continue the descent paused at a transition that needed draws
};

// One descent from the canonical root.  On return: kind == KIND_NEED_EVAL (leaf allocated, w.leaf holds its
// state), KIND_DONE (terminal / dead end: w.ret is the value returned by the deepest call), KIND_NEED_DRAW
// (injected mode only: the chosen transition needs dice the host has not supplied yet; w.leaf / w.depth /
// w.pending describe where to resume) or KIND_ERROR.
template <bool FEATURES, bool INJECT, int ROWS, int W>
__device__ __forceinline__ void descend(const View& v, YaState cur, Walk& w, uint64_t seed, uint32_t gid, uint32_t ep,
                                        uint32_t pl, uint32_t sim, float cpuct, float* __restrict__ feat_row,
                                        const Team<W>& tm, DrawSource src = DrawSource{nullptr, 0}) {
    const int lane = tm.sub;                                        // index inside the team
    w.depth = 0; w.kind = KIND_DONE; w.err = 0; w.leaf_node = -1; w.pending = 0;
    w.ret.d = 0.0; w.ret.is_f32 = false;
    bool pending = false;
    int a = 0;
    int known = -1;                                                  // node index of `cur` when it was reached through a cached child
    int have = (INJECT && src.inj && src.resume) ? src.inj[11] : 0;  // draws supplied for the paused transition
    if (INJECT && src.resume) {                                     // pick the paused transition up again
#pragma unroll
        for (int i = 0; i < 8; ++i) cur.w[i] = v.cur[C_LEAF + i];
        w.depth = (int)v.cur[C_DEPTH];
        a = (int)(v.cur[C_PENDING] & 0xFFFFu);
        pending = true;
    }
    for (;;) {
        if (!pending) {
            int idx = known, free_slot = 0;
            known = -1;
            if (idx < 0) {                                           // not reached through a cached child: look the state up
                float es = ya_game_ended(cur, 1);                   // Es[s], MCTS.py:79-83
                if (es != 0.0f) { w.ret.d = -es_as_double(es); w.ret.is_f32 = false; break; }
                idx = ht_find(v, cur, &free_slot);
            }
            if (idx < 0) {                                           // leaf: MCTS.py:84-115 (evaluation happens outside)
                uint32_t desc = ya_mask_desc(cur, 1);
                int L = ya_legal_count(desc);
                // 16-byte aligned prior rows; logit rows 32-byte aligned (the forward's epilogue stores whole sectors)
                uint32_t row_at = ROWS >= ROWS_L16F ? (w.arena_top + 7u) & ~7u : (w.arena_top + 3u) & ~3u;
                if ((int)w.node_count >= v.max_nodes) { w.err = E_NODES_FULL; w.kind = KIND_ERROR; break; }
                const uint32_t words = ROWS == ROWS_CONST ? (uint32_t)((L + 31) >> 5) : ROWS >= ROWS_L16F ? (uint32_t)l16_words(desc) : (uint32_t)row_words(L);
                if (row_at + words > v.arena_words) { w.err = E_ARENA_FULL; w.kind = KIND_ERROR; break; }
                idx = (int)w.node_count;
                if (lane == 0) {
                    uint32_t* nd = v.nodes + (int64_t)idx * kNodeWords;
#pragma unroll
                    for (int i = 0; i < 8; ++i) nd[N_KEY + i] = cur.w[i];
                    nd[N_DESC] = desc; nd[N_VISITS] = 0; nd[N_PRIOR] = row_at; nd[N_EDGES] = 0; nd[N_NEDGE] = 0;
                    nd[N_KIND] = ROWS == ROWS_CONST ? 1u : ROWS >= ROWS_L16F ? 2u : 0u; nd[N_PCONST] = 0;
                    v.ht[free_slot] = (uint16_t)(idx + 1);
                }
                w.node_count += 1;
                w.arena_top = row_at + words;
                if (FEATURES)
                    for (int f = lane; f < YA_N_FEATURE; f += W) feat_row[f] = ya_feature(cur, f);
                w.leaf_node = idx;
                w.leaf_desc = desc;
                w.leaf_row = row_at;
                w.kind = KIND_NEED_EVAL;
                tm.sync();
                break;
            }
            uint32_t* node = v.nodes + (int64_t)idx * kNodeWords;
            uint32_t desc = node[N_DESC];
            int L = ya_legal_count(desc);
            if (L == 0) { w.leaf_node = idx; w.ret.d = 0.0; w.ret.is_f32 = false; break; }   // MCTS.py:138-147: `return 0`, not negated
            if (w.depth >= kMaxDepth) { w.err = E_DEPTH; w.kind = KIND_ERROR; break; }
            const int key = ucb_select<ROWS>(v, node, L, cpuct, tm);
            const int ai = key >> 16, child = key & 0xFFFF;
            a = ya_nth_legal(desc, ai);
            // getNextState(s, 1, a) without dice or tie-break is a function of (s, a): its result can be cached on the edge
            const bool det = !INJECT && ya_draw_needs(cur, 1, a) == 0;
            if (lane == 0) v.cur[C_PATH + w.depth] = (uint32_t)idx | ((uint32_t)ai << 16) | (det ? 0x80000000u : 0u);
            if (det && child) {                                      // known child: its record holds the next canonical state
                const uint4* ck = reinterpret_cast<const uint4*>(v.nodes + (int64_t)(child - 1) * kNodeWords);
                const uint4 k0 = ck[0], k1 = ck[1];
                cur.w[0] = k0.x; cur.w[1] = k0.y; cur.w[2] = k0.z; cur.w[3] = k0.w;
                cur.w[4] = k1.x; cur.w[5] = k1.y; cur.w[6] = k1.z; cur.w[7] = k1.w;
                known = child - 1;
                ++w.depth;
                continue;
            }
        }
        pending = false;
        YaDraw d;
        d.roll_a = d.roll_b = d.tie = d.pick = 0;
        const int needs = ya_draw_needs(cur, 1, a);
        if (needs) {
            if (!INJECT) {
                d = ya_draw(seed, gid, ep, pl, YA_TAG_SEARCH, (uint32_t)w.depth, sim);
            } else {
                int missing = 0;
                if ((needs & YA_NEED_TIE) && !(have & 1)) missing |= YA_NEED_TIE;
                if ((needs & YA_NEED_ROLLS) && !(have & 2)) missing |= YA_NEED_ROLLS;
                if (missing) { w.kind = KIND_NEED_DRAW; w.pending = (uint32_t)a | ((uint32_t)missing << 16); break; }
                d.tie = src.inj[0];
                for (int i = 0; i < 5; ++i) {
                    d.roll_a |= (uint32_t)src.inj[1 + i] << (3 * i);
                    d.roll_b |= (uint32_t)src.inj[6 + i] << (3 * i);
                }
                have = 0;                                            // injected draws are single use
            }
        }
        int st;
        int np = ya_transition(cur, 1, a, d, &st);                   // getNextState(canonicalBoard, 1, a), MCTS.py:149
        if (st != YA_OK) { w.err = E_RULE | (1 << st); w.kind = KIND_ERROR; break; }
        if (np != 1) cur = ya_flip(cur);                             // getCanonicalForm(next_s, next_player), MCTS.py:150
        ++w.depth;
    }
    w.leaf = cur;
    tm.sync();
}

template <bool WRITE_LEAF_STATE, bool INJECT, int ROWS>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, YA_MCTS_MIN_BLOCKS)
ya_k_mcts_select(ya_mcts_tree tree, const uint4* __restrict__ states, int64_t stride, const int8_t* __restrict__ players,
                 const int32_t* __restrict__ ply, const uint32_t* __restrict__ episode, uint64_t seed, uint64_t game_base,
                 uint32_t sim, const uint32_t* __restrict__ sim_ptr, const uint64_t* __restrict__ game_base_ptr, float cpuct,
                 const uint8_t* __restrict__ active, float* __restrict__ features, uint8_t* __restrict__ need_eval, uint32_t* __restrict__ leaf_states,
                 int32_t* __restrict__ err_flag, const uint8_t* __restrict__ injected, int resume,
                 uint64_t* __restrict__ leaf_dst, uint32_t* __restrict__ leaf_desc) {
    const Team<kSelectTeam> tm = Team<kSelectTeam>::make();          // four games per warp
    const int lane = tm.sub;
    if (sim_ptr) sim = *sim_ptr;                                     // CUDA-graph replay: the counter lives in HBM,
    if (game_base_ptr) game_base = *game_base_ptr;                   // and so does the global id of this slice's first game
    const int64_t g = ((int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5)) * (32 / kSelectTeam) + (tm.shift / kSelectTeam);
    if (g >= tree.n) return;
    if (active && !active[g]) { if (lane == 0) { need_eval[g] = 0; if (leaf_dst) leaf_dst[g] = 0; } return; }
    View v = make_view(tree, g);
    YaState root = ya_load(states, stride, g);
    if (players[g] != 1) root = ya_flip(root);                       // getCanonicalForm of the root
    const uint32_t gid = (uint32_t)(game_base + g);
    const uint32_t ep = episode ? episode[g] : 0u, pl = ply ? (uint32_t)ply[g] : 0u;
    Walk w;
    w.node_count = v.meta[M_NODES];
    w.arena_top = v.meta[M_TOP];
    if (sim == 0 && !(INJECT && resume)) prune_on_new_round(v, root, w, tm);
    DrawSource src{INJECT ? injected + g * 12 : nullptr, INJECT ? resume : 0};
    descend<true, INJECT, ROWS>(v, root, w, seed, gid, ep, pl, sim, cpuct, features + g * YA_N_FEATURE, tm, src);
    if (w.kind == KIND_DONE) {
        if (!backup_path<ROWS>(v, w.depth, w.ret, w.arena_top, tm, w.leaf_node)) { w.err = E_ARENA_FULL; w.kind = KIND_ERROR; }
    }
    if (lane == 0) {
        v.meta[M_NODES] = w.node_count;
        v.meta[M_TOP] = w.arena_top;
        v.cur[C_DEPTH] = (uint32_t)w.depth;
        v.cur[C_KIND] = (uint32_t)w.kind;
        v.cur[C_NODE] = (uint32_t)w.leaf_node;
        if (w.kind == KIND_NEED_EVAL) { v.cur[C_LEAF_DESC] = w.leaf_desc; v.cur[C_LEAF_ROW] = w.leaf_row; }
        uint8_t code = w.kind == KIND_NEED_EVAL ? 1 : 0;
        if (INJECT && w.kind == KIND_NEED_DRAW) {                    // park the descent; tell the host what to draw
            v.cur[C_PENDING] = w.pending;
#pragma unroll
            for (int i = 0; i < 8; ++i) v.cur[C_LEAF + i] = w.leaf.w[i];
            code = (uint8_t)(0x10 | (((w.pending >> 16) & YA_NEED_TIE) ? 1 : 0) | (((w.pending >> 16) & YA_NEED_ROLLS) ? 2 : 0));
        }
        need_eval[g] = code;
        if (leaf_dst) {                                              // where the forward kernel's epilogue puts the legal logits
            leaf_dst[g] = w.kind == KIND_NEED_EVAL ? reinterpret_cast<uint64_t>(v.arena + w.leaf_row) : 0ull;
            leaf_desc[g] = w.kind == KIND_NEED_EVAL ? w.leaf_desc : 0u;
        }
        if (w.err && err_flag) atomicOr(err_flag, w.err);
        if (WRITE_LEAF_STATE && leaf_states && w.kind == KIND_NEED_EVAL) {
            uint4* o = reinterpret_cast<uint4*>(leaf_states);
            o[g] = make_uint4(w.leaf.w[0], w.leaf.w[1], w.leaf.w[2], w.leaf.w[3]);
            o[tree.n + g] = make_uint4(w.leaf.w[4], w.leaf.w[5], w.leaf.w[6], w.leaf.w[7]);
        }
    }
}

// ---------------------------------------------------------------- expand (MCTS.py:86-115) + backup
// validity of action i for a mask descriptor
__device__ __forceinline__ bool desc_valid(uint32_t desc, int i) {
    int r = ((i + 50) * 4162) >> 20;
    if (!((desc >> r) & 1u)) return false;
    if (r == 0 || (desc >> 13)) return true;
    return i == YA_N_BID + YA_N_SUBSET * (r - 1);
}

// does [lo, hi) contain a legal action?  (runs: 0 = bids [0,202), r = category r-1 [202+252(r-1), +252))
__device__ __forceinline__ bool range_has_legal(uint32_t desc, int lo, int hi) {
    int r0 = ((lo + 50) * 4162) >> 20, r1 = ((hi - 1 + 50) * 4162) >> 20;
    uint32_t runs = ((2u << r1) - 1u) & ~((1u << r0) - 1u);          // bits r0..r1
    uint32_t hit = desc & runs & 0x1FFFu;
    if (!hit) return false;
    if ((desc >> 13) || (hit & 1u)) return true;
    // five dice: only the first byte of each open category run is legal
    for (uint32_t m = hit >> 1; m; m &= m - 1) {
        int first = YA_N_BID + YA_N_SUBSET * (__ffs(m) - 1);
        if (first >= lo && first < hi) return true;
    }
    return false;
}

// sum of x[i] = valid(i) ? pi[i] : 0 over i < 3226 in numpy's float32 pairwise order
// (numpy/_core/src/umath/loops_utils.h.src: blocks of <= 128 with 8 accumulators, halves rounded
// down to a multiple of 8).  For n = 3226 that is a perfect binary tree over 32 blocks of 96 / 104 /
// 106 elements: 8 lanes run a block's accumulators, four blocks per pass.  `load(i)` returns pi[i]
// for a legal action i.
template <class Load>
__device__ __forceinline__ float masked_pairwise_sum(Load load, uint32_t desc, int lane) {
    const int sub = lane & 7, grp = lane >> 3;
    float my_leaf = 0.0f;
#pragma unroll 1
    for (int pass = 0; pass < 8; ++pass) {
        int leaf = pass * 4 + grp;
        int start = 0, n = YA_N_ACTION;
#pragma unroll
        for (int lvl = 4; lvl >= 0; --lvl) {
            int n2 = n / 2;
            n2 -= n2 % 8;
            if ((leaf >> lvl) & 1) { start += n2; n -= n2; } else { n = n2; }
        }
        int body = n - (n % 8);
        float r = 0.0f;
        // A block (<= 106 actions) spans at most two runs of the mask.  For bid / ten-dice rows a run is
        // all-legal or all-illegal: validity is one compare against the run boundary.
        const bool simple = (desc & 1u) || (desc >> 13);
        const int r0 = ((start + 50) * 4162) >> 20;
        const int bnd = 202 + 252 * r0;
        const bool v0 = (desc >> r0) & 1u, v1 = (desc >> (r0 + 1)) & 1u & (r0 < 12);
        auto valid = [&](int k) { return simple ? (k < bnd ? v0 : v1) : desc_valid(desc, k); };
        // a block whose actions are all illegal sums to exactly +0: skip it (bid rows touch 3 of 32 blocks)
        if (range_has_legal(desc, start, start + n)) {
            int i = start + sub;
            r = valid(i) ? load(i) : 0.0f;
            for (int t = 8; t < body; t += 8) {
                int k = start + t + sub;
                float x = valid(k) ? load(k) : 0.0f;
                r = __fadd_rn(r, x);
            }
        }
        // ((r0+r1)+(r2+r3)) + ((r4+r5)+(r6+r7))
        r = __fadd_rn(r, __shfl_xor_sync(0xFFFFFFFFu, r, 1));
        r = __fadd_rn(r, __shfl_xor_sync(0xFFFFFFFFu, r, 2));
        r = __fadd_rn(r, __shfl_xor_sync(0xFFFFFFFFu, r, 4));
        for (int k = start + body; k < start + n; ++k) {              // the n % 8 leftovers, in order
            float x = valid(k) ? load(k) : 0.0f;
            r = __fadd_rn(r, x);
        }
        // lane `leaf` keeps block `leaf`
        float from = __shfl_sync(0xFFFFFFFFu, r, ((lane & 3) << 3));
        if ((lane >> 2) == pass) my_leaf = from;
    }
    // perfect binary tree over the 32 block sums
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) my_leaf = __fadd_rn(my_leaf, __shfl_xor_sync(0xFFFFFFFFu, my_leaf, o));
    return my_leaf;
}

// Writes the leaf's legal-only prior row (MCTS.py:88-113).
// MODE 0: pi float32[3226] from the evaluator; MODE 1: uniform prior, nothing read.
template <int MODE>
__device__ __forceinline__ void write_prior_row(float* __restrict__ row, uint32_t desc, int L, const float* pi,
                                                float uniform_p, int lane) {
    float total;
    if (MODE == 0) {
        total = masked_pairwise_sum([pi](int i) { return pi[i]; }, desc, lane);
        if (total > 0.0f)                                            // Ps /= sum, MCTS.py:90-91
            for (int k0 = 0; k0 < L; k0 += 32) {
                const int k = k0 + lane;
                store_prior_group(row, L, k0, k < L ? __fdiv_rn(pi[ya_nth_legal(desc, k)], total) : 0.0f, lane);
            }
    } else {
        total = masked_pairwise_sum([uniform_p](int) { return uniform_p; }, desc, lane);
        if (total > 0.0f) {
            float p = __fdiv_rn(uniform_p, total);
            for (int k0 = 0; k0 < L; k0 += 32) store_prior_group(row, L, k0, p, lane);
        }
    }
    if (!(total > 0.0f)) {                                           // all legal moves masked: uniform over legal, :97-101
        float u = __fdiv_rn(1.0f, (float)L);
        for (int k0 = 0; k0 < L; k0 += 32) store_prior_group(row, L, k0, u, lane);
    }
}

template <int MODE>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
ya_k_mcts_expand(ya_mcts_tree tree, const float* __restrict__ pi_all, const float* __restrict__ value, float uniform_p,
                 float uniform_v, uint32_t* __restrict__ sim_counter, int32_t* __restrict__ err_flag) {
    const int lane = threadIdx.x & 31;
    const int64_t g = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (sim_counter && blockIdx.x == 0 && threadIdx.x == 0) *sim_counter += 1;   // next replay = next simulation
    if (g >= tree.n) return;
    View v = make_view(tree, g);
    if (v.cur[C_KIND] != KIND_NEED_EVAL) return;
    uint32_t* node = v.nodes + (int64_t)v.cur[C_NODE] * kNodeWords;
    const uint32_t desc = node[N_DESC];
    const int L = ya_legal_count(desc);
    if (L > 0)
        write_prior_row<MODE>(reinterpret_cast<float*>(v.arena + node[N_PRIOR]), desc, L,
                              MODE == 0 ? pi_all + g * YA_N_ACTION : nullptr, uniform_p, lane);
    __syncwarp();
    Val ret;
    ret.d = -(double)(MODE == 1 ? uniform_v : value[g]);             // return -v (numpy float32)
    ret.is_f32 = true;
    uint32_t arena_top = v.meta[M_TOP];
    bool ok = backup_path<ROWS_F32>(v, (int)v.cur[C_DEPTH], ret, arena_top, Team<32>::make(), (int)v.cur[C_NODE]);
    if (lane == 0) {
        v.meta[M_TOP] = arena_top;
        v.cur[C_KIND] = KIND_DONE;
        if (!ok && err_flag) atomicOr(err_flag, E_ARENA_FULL);
    }
}

// 16-bit policy-head logits -> the leaf's logit row + exponent offset (MCTS.py:86-101 after NNetWrapper.predict's softmax,
// yacht/NNet.py:193): softmax, mask and renormalisation collapse to
//     P[a] = exp(l[a] - max_all) / sum over legal a' of exp(l[a'] - max_all) = 2^(l[a] * log2(e) + off)
// with one offset per leaf, so neither float32 logits nor pi nor float32 priors ever exist in memory.
// FROM_DENSE = false: the forward kernel's epilogue has already written the LEGAL logits into the row (ya_nn_forward with
//   scatter targets); this kernel re-reads them (L2 hits: 2 * L bytes written a few microseconds ago), computes the offset
//   and the per-32 group maxima and backs the value up.
// FROM_DENSE = true: any evaluator that returns a dense [n][ld] 16-bit logit matrix; the legal entries are compacted into
//   the row here (asynchronous 4-byte copies through shared memory).
// max_all (over all 3,226 logits: it decides when every legal exp underflows and the reference falls back to the uniform
// row) comes from the forward kernel's epilogue (row_max) or, FROM_DENSE only, from one pass over the dense row.
constexpr int kLogitWarps = 4;
constexpr int kLogitCols = 3232;
template <bool F16, bool FROM_DENSE>
__global__ void __launch_bounds__(kLogitWarps * 32)
ya_k_mcts_expand_rows(ya_mcts_tree tree, const uint16_t* __restrict__ logits_all, int64_t ld,
                      const float* __restrict__ row_max, const float* __restrict__ value,
                      uint32_t* __restrict__ sim_counter, int32_t* __restrict__ err_flag) {
    constexpr int ROWS = F16 ? ROWS_L16F : ROWS_L16B;
    __shared__ __align__(16) uint32_t raw_all[kLogitWarps][kLogitCols / 2];
    __shared__ float group_logit[kLogitWarps][96];                     // largest logit of every group of 32 legal moves
    const int lane = threadIdx.x & 31;
    uint32_t* raw = raw_all[threadIdx.x >> 5];
    const int64_t g = (int64_t)blockIdx.x * kLogitWarps + (threadIdx.x >> 5);
    if (sim_counter && blockIdx.x == 0 && threadIdx.x == 0) *sim_counter += 1;
    if (g >= tree.n) return;
    View v = make_view(tree, g);
    if (v.cur[C_KIND] != KIND_NEED_EVAL) return;
    const uint32_t desc = v.cur[C_LEAF_DESC];                          // left by the descent next to the path: the leaf's
    const uint32_t leaf_row = v.cur[C_LEAF_ROW];                       // node record is not needed before the logits load
    const int L = ya_legal_count(desc);
    if (L > 0) {
        uint32_t* row = v.arena + leaf_row;
        const int lw = l16_logit_words(desc), nb = (L + 31) >> 5, nw_all = (L + 1) >> 1;
        const uint32_t raw_s = (uint32_t)__cvta_generic_to_shared(raw);
        const uint32_t* lg = FROM_DENSE ? reinterpret_cast<const uint32_t*>(logits_all + g * ld) : nullptr;   // two logits per word
        if (FROM_DENSE) {
            auto copy_word = [&](int dst_word, int src_word) {
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(raw_s + 4u * (uint32_t)dst_word), "l"(lg + src_word) : "memory");
            };
            if (desc & 1u) {                                           // bid row: actions 0..201
                for (int j = lane; j < YA_N_BID / 2; j += 32) copy_word(j, j);
            } else if (desc >> 13) {                                   // ten dice: 252 subsets per open category
                int k0 = 0;
                for (uint32_t open = (desc >> 1) & 0xFFFu; open; open &= open - 1, k0 += YA_N_SUBSET / 2) {
                    const int a0 = (YA_N_BID + (__ffs(open) - 1) * YA_N_SUBSET) / 2;
                    const uint32_t dst = raw_s + 4u * (uint32_t)(k0 + lane);
                    const uint32_t* src = lg + a0 + lane;              // 126 words: lanes 0..29 take a fourth one
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n\t"
                                 "cp.async.ca.shared.global [%0 + 128], [%1 + 128], 4;\n\t"
                                 "cp.async.ca.shared.global [%0 + 256], [%1 + 256], 4;" ::"r"(dst), "l"(src) : "memory");
                    if (lane < YA_N_SUBSET / 2 - 96)
                        asm volatile("cp.async.ca.shared.global [%0 + 384], [%1 + 384], 4;" ::"r"(dst), "l"(src) : "memory");
                }
            } else {                                                   // five dice: subset 0 of every open category
                if (lane < L) reinterpret_cast<uint16_t*>(raw)[lane] = reinterpret_cast<const uint16_t*>(lg)[ya_nth_legal(desc, lane)];
            }
        } else if (desc >> 13) {                                       // the row the forward kernel filled: 126 words per open
            int k0 = 0, blk = 0;                                       // category, pad / 2 words into its 136-word block
            for (uint32_t open = (desc >> 1) & 0xFFFu; open; open &= open - 1, k0 += YA_N_SUBSET / 2, blk += kL16Block / 2) {
                const uint32_t dst = raw_s + 4u * (uint32_t)(k0 + lane);
                const uint32_t* src = row + blk + (l16_pad(__ffs(open) - 1) >> 1) + lane;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n\t"
                             "cp.async.ca.shared.global [%0 + 128], [%1 + 128], 4;\n\t"
                             "cp.async.ca.shared.global [%0 + 256], [%1 + 256], 4;" ::"r"(dst), "l"(src) : "memory");
                if (lane < YA_N_SUBSET / 2 - 96)
                    asm volatile("cp.async.ca.shared.global [%0 + 384], [%1 + 384], 4;" ::"r"(dst), "l"(src) : "memory");
            }
        } else {                                                       // bid / five-dice rows are compact: 16-byte chunks
            for (int j = lane * 4; j < nw_all; j += 128)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(raw_s + 4u * (uint32_t)j), "l"(row + j) : "memory");
        }
        // While the logits are in flight: pull the lines the backup will walk (one lane per level of the path --
        // the node, the head of its edge index array, the logit group and group maximum of the chosen child).
        {
            const int depth = (int)v.cur[C_DEPTH];
            if (lane < depth) {
                const uint32_t pe = v.cur[C_PATH + lane];
                const uint32_t* nd = v.nodes + (int64_t)(pe & 0xFFFFu) * kNodeWords;
                const int ne = (int)nd[N_NEDGE], ai = path_move(pe);
                const uint32_t* prow = v.arena + nd[N_PRIOR];
                asm volatile("prefetch.global.L2 [%0];" ::"l"(prow + (l16_slot(nd[N_DESC], ai & ~31) >> 1)));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(prow + l16_logit_words(nd[N_DESC]) + (ai >> 5)));
                if (ne > 0) {
                    const uint32_t* eb = v.arena + nd[N_EDGES];
                    const int cap = edge_cap(ne);
                    for (int b = 0; b < ne * 2; b += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(eb) + b));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(eb + cap));                // Nsa / Q of a one-chunk node
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(eb + 2 * cap));
                }
            }
            __syncwarp();
        }
        float mx;
        if (row_max) {
            mx = row_max[g];
        } else if (FROM_DENSE) {
            mx = -CUDART_INF_F;
            for (int j = lane; j < YA_N_ACTION / 2; j += 32) {
                uint32_t w = lg[j];
                mx = fmaxf(mx, fmaxf(l16_value<ROWS>(w & 0xFFFFu), l16_value<ROWS>(w >> 16)));
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
        } else {
            mx = 0.0f;                                                 // not reachable: the entry point requires row_max
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncwarp();
        if (FROM_DENSE) {                                              // the compacted legal logits become the node's row
            if (desc >> 13) {
                int k0 = 0, blk = 0;
                for (uint32_t open = (desc >> 1) & 0xFFFu; open; open &= open - 1, k0 += YA_N_SUBSET / 2, blk += kL16Block / 2) {
                    uint32_t* dst = row + blk + (l16_pad(__ffs(open) - 1) >> 1);
                    for (int j = lane; j < YA_N_SUBSET / 2; j += 32) dst[j] = raw[k0 + j];
                }
            } else {
                for (int j = lane; j < nw_all; j += 32) row[j] = raw[j];
            }
        }
        // exp(l - max) = 2^(l * log2(e) - max * log2(e)): one FMA + one EX2 per logit; the normalisation is folded into
        // the exponent as well (P = 2^(.. - log2(total))).
        const float off_sum = -mx * kLog2e;
        const uint16_t* rc = reinterpret_cast<const uint16_t*>(raw);
        const bool pairs = !(L & 1), quads = !(L & 3);                 // ten-dice rows: L = 252 * open categories
        const int nw = L >> 1;
        // ONE pass over the legal logits: the softmax denominator and, per group of 32 legal moves, the largest logit
        // (the group maximum of the priors is P(largest logit): P is a monotone function of the logit)
        float* gl = group_logit[threadIdx.x >> 5];
        float total = 0.0f;
        if (quads) {
            float t1 = 0.0f, t2 = 0.0f, t3 = 0.0f;
            const uint2* raw2 = reinterpret_cast<const uint2*>(raw);
            const int nq = L >> 2;
            for (int j0 = 0; j0 < nq; j0 += 32) {                      // a lane owns four consecutive logits, eight lanes a group
                const int j = j0 + lane;
                float m = -CUDART_INF_F;
                if (j < nq) {
                    const uint2 w = raw2[j];
                    const float a0 = l16_value<ROWS>(w.x & 0xFFFFu), a1 = l16_value<ROWS>(w.x >> 16);
                    const float a2 = l16_value<ROWS>(w.y & 0xFFFFu), a3 = l16_value<ROWS>(w.y >> 16);
                    total += ex2_approx(fmaf(a0, kLog2e, off_sum));
                    t1 += ex2_approx(fmaf(a1, kLog2e, off_sum));
                    t2 += ex2_approx(fmaf(a2, kLog2e, off_sum));
                    t3 += ex2_approx(fmaf(a3, kLog2e, off_sum));
                    m = fmaxf(fmaxf(a0, a1), fmaxf(a2, a3));
                }
                m = fmaxf(m, __shfl_xor_sync(0xFFFFFFFFu, m, 1));
                m = fmaxf(m, __shfl_xor_sync(0xFFFFFFFFu, m, 2));
                m = fmaxf(m, __shfl_xor_sync(0xFFFFFFFFu, m, 4));
                if ((lane & 7) == 0 && j < nq) gl[j >> 3] = m;         // group of legal moves 4j .. 4j + 31
            }
            total = (total + t1) + (t2 + t3);
        } else if (pairs) {
            float t1 = 0.0f;
            for (int j0 = 0; j0 < nw; j0 += 32) {                      // two logits per lane, sixteen lanes a group
                const int j = j0 + lane;
                float m = -CUDART_INF_F;
                if (j < nw) {
                    const uint32_t w = raw[j];
                    const float a0 = l16_value<ROWS>(w & 0xFFFFu), a1 = l16_value<ROWS>(w >> 16);
                    total += ex2_approx(fmaf(a0, kLog2e, off_sum));
                    t1 += ex2_approx(fmaf(a1, kLog2e, off_sum));
                    m = fmaxf(a0, a1);
                }
#pragma unroll
                for (int o = 1; o < 16; o <<= 1) m = fmaxf(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
                if ((lane & 15) == 0 && j < nw) gl[j >> 4] = m;
            }
            total += t1;
        } else {
            for (int k0 = 0; k0 < L; k0 += 32) {
                const int k = k0 + lane;
                float m = -CUDART_INF_F;
                if (k < L) {
                    m = l16_value<ROWS>(rc[k]);
                    total += ex2_approx(fmaf(m, kLog2e, off_sum));
                }
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) m = fmaxf(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
                if (lane == 0) gl[k0 >> 5] = m;
            }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) total += __shfl_xor_sync(0xFFFFFFFFu, total, o);
        uint32_t* gm = row + lw;
        uint32_t* seen = row + lw + nb;
        uint32_t* node = v.nodes + (int64_t)v.cur[C_NODE] * kNodeWords;
        for (int b = lane; b < nb; b += 32) seen[b] = 0u;
        __syncwarp();
        if (total > 0.0f) {
            const float off_p = off_sum - __log2f(total);
            if (lane == 0) node[N_PCONST] = __float_as_uint(off_p);
            for (int b = lane; b < nb; b += 32) gm[b] = __float_as_uint(ex2_approx(fmaf(gl[b], kLog2e, off_p))) + 1u;
        } else if (lane == 0) {                                        // every legal move underflowed (MCTS.py:97-101): Ps = 1 / L,
            node[N_KIND] = 1u;                                         // a constant-prior node on the row's visited bits
            node[N_PRIOR] = leaf_row + (uint32_t)(lw + nb);
            node[N_PCONST] = __float_as_uint(__fdiv_rn(1.0f, (float)L));
        }
    }
    __syncwarp();
    Val ret;
    ret.d = -(double)value[g];
    ret.is_f32 = true;
    uint32_t arena_top = v.meta[M_TOP];
    bool ok = backup_path<ROWS>(v, (int)v.cur[C_DEPTH], ret, arena_top, Team<32>::make(), (int)v.cur[C_NODE]);
    if (lane == 0) {
        v.meta[M_TOP] = arena_top;
        v.cur[C_KIND] = KIND_DONE;
        if (!ok && err_flag) atomicOr(err_flag, E_ARENA_FULL);
    }
}

// ---------------------------------------------------------------- whole search, uniform prior
// BASELINE.json configs[2] (no network): nothing has to leave the SM between select and expand, so all
// numMCTSSims simulations of a move run inside ONE launch -- one warp walks, expands and backs up its
// game's tree num_sims times; node / row / edge data stay hot in L1/L2 across simulations.
__global__ void __launch_bounds__(kWarpsPerBlock * 32, YA_MCTS_MIN_BLOCKS)
ya_k_mcts_search_uniform(ya_mcts_tree tree, const uint4* __restrict__ states, int64_t stride, const int8_t* __restrict__ players,
                         const int32_t* __restrict__ ply, const uint32_t* __restrict__ episode, uint64_t seed,
                         uint64_t game_base, int num_sims, float cpuct, float uniform_p, float uniform_v,
                         const uint8_t* __restrict__ active, int32_t* __restrict__ err_flag) {
    const int lane = threadIdx.x & 31;
    const int64_t g = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (g >= tree.n) return;
    if (active && !active[g]) return;
    View v = make_view(tree, g);
    YaState root = ya_load(states, stride, g);
    if (players[g] != 1) root = ya_flip(root);
    const uint32_t gid = (uint32_t)(game_base + g);
    const uint32_t ep = episode ? episode[g] : 0u, pl = ply ? (uint32_t)ply[g] : 0u;
    Walk w;
    w.node_count = v.meta[M_NODES];
    w.arena_top = v.meta[M_TOP];
    const Team<32> tm = Team<32>::make();
    prune_on_new_round(v, root, w, tm);
    int err = 0;
    for (int sim = 0; sim < num_sims; ++sim) {
        descend<false, false, ROWS_CONST>(v, root, w, seed, gid, ep, pl, (uint32_t)sim, cpuct, nullptr, tm);
        if (w.kind == KIND_ERROR) { err = w.err; break; }
        Val ret = w.ret;
        if (w.kind == KIND_NEED_EVAL) {
            uint32_t* node = v.nodes + (int64_t)w.leaf_node * kNodeWords;
            const uint32_t desc = node[N_DESC];
            const int L = ya_legal_count(desc);
            if (L > 0) {
                // Ps = uniform_p * valids, renormalised (MCTS.py:88-101): one value for every legal move; only the
                // numpy-ordered sum decides its bits.  The node keeps that value and a visited bitmask.
                float total = masked_pairwise_sum([uniform_p](int) { return uniform_p; }, desc, lane);
                const float p = total > 0.0f ? __fdiv_rn(uniform_p, total) : __fdiv_rn(1.0f, (float)L);
                uint32_t* mask = v.arena + node[N_PRIOR];
                for (int b = lane; b < ((L + 31) >> 5); b += 32) mask[b] = 0u;
                if (lane == 0) node[N_PCONST] = __float_as_uint(p);
            }
            __syncwarp();
            ret.d = -(double)uniform_v;
            ret.is_f32 = true;
        }
        if (!backup_path<ROWS_CONST>(v, w.depth, ret, w.arena_top, tm, w.leaf_node)) { err = E_ARENA_FULL; break; }
    }
    if (lane == 0) {
        v.meta[M_NODES] = w.node_count;
        v.meta[M_TOP] = w.arena_top;
        v.cur[C_KIND] = KIND_DONE;
        if (err && err_flag) atomicOr(err_flag, err);
    }
}

// ---------------------------------------------------------------- root statistics (MCTS.py:40-42)
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
ya_k_mcts_root_counts(ya_mcts_tree tree, const uint4* __restrict__ states, int64_t stride, const int8_t* __restrict__ players,
                      int32_t* __restrict__ counts, int32_t* __restrict__ visits, double* __restrict__ qvals,
                      uint8_t* __restrict__ qkind) {
    const int lane = threadIdx.x & 31;
    const int64_t g = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (g >= tree.n) return;
    View v = make_view(tree, g);
    YaState cur = ya_load(states, stride, g);
    if (players[g] != 1) cur = ya_flip(cur);
    int32_t* row = counts + g * YA_N_ACTION;
    for (int i = lane; i < YA_N_ACTION; i += 32) {
        row[i] = 0;
        if (qvals) qvals[g * YA_N_ACTION + i] = 0.0;
        if (qkind) qkind[g * YA_N_ACTION + i] = 0;
    }
    __syncwarp();
    int free_slot;
    int idx = ht_find(v, cur, &free_slot);
    if (idx < 0) { if (lane == 0 && visits) visits[g] = -1; return; }
    const uint32_t* node = v.nodes + (int64_t)idx * kNodeWords;
    if (lane == 0 && visits) visits[g] = (int32_t)node[N_VISITS];
    uint32_t desc = node[N_DESC];
    const int n_edges = (int)node[N_NEDGE];
    if (n_edges == 0) return;
    const Edges ed = edges_at(v.arena + node[N_EDGES], edge_cap(n_edges));
    for (int e = lane; e < n_edges; e += 32) {
        int a = ya_nth_legal(desc, ed.idx[e]);
        uint32_t raw = ed.nsa[e];
        row[a] = (int32_t)(raw & 0x7FFFFFFFu);
        if (qvals) qvals[g * YA_N_ACTION + a] = ed.q[e];
        if (qkind) qkind[g * YA_N_ACTION + a] = (raw >> 31) ? 2 : 1;           // 1 = float32, 2 = Python float
    }
}

// ---------------------------------------------------------------- sparse root policy (training examples)
// The visited root edges as (action, Nsa) pairs in ascending action order, zero padded to k entries: the
// canonical sparse form of the pi vector Coach.executeEpisode records (Coach.py:60-63).  overflow[g] = number
// of visited edges when they do not fit k (the row then holds the k lowest actions).
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
ya_k_mcts_root_sparse(ya_mcts_tree tree, const uint4* __restrict__ states, int64_t stride, const int8_t* __restrict__ players,
                      int k, int16_t* __restrict__ actions, int32_t* __restrict__ counts, int32_t* __restrict__ overflow) {
    const int lane = threadIdx.x & 31;
    const int64_t g = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (g >= tree.n) return;
    View v = make_view(tree, g);
    YaState cur = ya_load(states, stride, g);
    if (players[g] != 1) cur = ya_flip(cur);
    int16_t* arow = actions + g * k;
    int32_t* crow = counts + g * k;
    for (int i = lane; i < k; i += 32) { arow[i] = 0; crow[i] = 0; }
    __syncwarp();
    int free_slot;
    int idx = ht_find(v, cur, &free_slot);
    if (idx < 0) { if (lane == 0 && overflow) overflow[g] = -1; return; }
    const uint32_t* node = v.nodes + (int64_t)idx * kNodeWords;
    const int n_edges = (int)node[N_NEDGE];
    if (lane == 0 && overflow) overflow[g] = n_edges > k ? n_edges : 0;
    // legal indices ascend with actions, so the rank of an edge = number of edges with a smaller legal index
    if (n_edges == 0) return;
    const Edges ed = edges_at(v.arena + node[N_EDGES], edge_cap(n_edges));
    for (int e = lane; e < n_edges; e += 32) {
        const int my = ed.idx[e];
        const uint32_t nsa = ed.nsa[e] & 0x7FFFFFFFu;
        int rank = 0;
        for (int j = 0; j < n_edges; ++j) rank += ed.idx[j] < my;
        if (rank < k) { arow[rank] = (int16_t)ya_nth_legal(node[N_DESC], my); crow[rank] = (int32_t)nsa; }
    }
}

// ---------------------------------------------------------------- action from visit counts
// temp = 1 (Coach.py:56-65): inverse CDF over the integer counts, r = (word * total) >> 32.
// temp = 0 (MCTS.py:44-49): uniformly among the arg-max actions, k = (word * ties) >> 32.
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
ya_k_mcts_pick(const int32_t* __restrict__ counts, const int32_t* __restrict__ ply, const uint32_t* __restrict__ episode,
               int64_t n, uint64_t seed, uint64_t game_base, int temp_threshold, int32_t* __restrict__ actions) {
    const int lane = threadIdx.x & 31;
    const int64_t g = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (g >= n) return;
    const int32_t* row = counts + g * YA_N_ACTION;
    uint32_t p = ply ? (uint32_t)ply[g] : 0u;
    YaDraw d = ya_draw(seed, (uint32_t)(game_base + g), episode ? episode[g] : 0u, p, YA_TAG_ACTION, 0, 0);
    const bool greedy = !((int)(p + 1) < temp_threshold);
    // each lane owns a contiguous slice of 101 actions
    const int per = (YA_N_ACTION + 31) / 32;
    const int lo = lane * per, hi = min(lo + per, YA_N_ACTION);
    long long sum = 0;
    int mx = 0;
    for (int i = lo; i < hi; ++i) { int c = row[i]; sum += c; mx = max(mx, c); }
#pragma unroll
    for (int o = 16; o; o >>= 1) mx = max(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
    long long weight = 0;                                            // what this lane contributes to the CDF
    if (greedy) { for (int i = lo; i < hi; ++i) weight += row[i] == mx; } else weight = sum;
    long long incl = weight;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        long long t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += t;
    }
    long long total = __shfl_sync(0xFFFFFFFFu, incl, 31);
    long long r = (long long)(((unsigned long long)d.pick * (unsigned long long)total) >> 32);
    long long before = incl - weight;
    int found = -1;
    if (total > 0 && r >= before && r < incl) {
        long long acc = before;
        for (int i = lo; i < hi; ++i) {
            acc += greedy ? (row[i] == mx) : row[i];
            if (acc > r) { found = i; break; }
        }
    }
    uint32_t m = __ballot_sync(0xFFFFFFFFu, found >= 0);
    int src = m ? __ffs(m) - 1 : 0;
    int a = __shfl_sync(0xFFFFFFFFu, found, src);
    if (lane == 0) actions[g] = m ? a : -1;
}

__global__ void ya_k_mcts_reset(ya_mcts_tree tree, const uint8_t* __restrict__ which) {
    const int64_t g = blockIdx.x;
    if (g >= tree.n || (which && !which[g])) return;
    View v = make_view(tree, g);
    for (int i = threadIdx.x; i < v.ht_size / 2; i += blockDim.x) reinterpret_cast<uint32_t*>(v.ht)[i] = 0u;
    if (threadIdx.x == 0) {
        v.meta[M_NODES] = 0; v.meta[M_TOP] = 4; v.meta[M_ROUND] = 0; v.meta[3] = 0;
        v.cur[C_KIND] = KIND_DONE; v.cur[C_DEPTH] = 0;
    }
}

inline int warp_blocks(int64_t n) { return (int)((n + kWarpsPerBlock - 1) / kWarpsPerBlock); }
inline int select_blocks(int64_t n) {                                 // 32 / kSelectTeam games per warp
    const int64_t per_block = kWarpsPerBlock * (32 / kSelectTeam);
    return (int)((n + per_block - 1) / per_block);
}

bool tree_ok(const ya_mcts_tree* t) {
    return t && t->n > 0 && t->max_nodes > 0 && t->max_nodes < 65535 && t->ht_size >= 2 * t->max_nodes &&
           (t->ht_size & (t->ht_size - 1)) == 0 && t->arena_words > 4 && (t->arena_words % 4) == 0 &&
           t->arena_words < (1ll << 32);
}

}  // namespace

extern "C" {

int ya_mcts_cursor_words(void) { return kCursorWords; }
int ya_mcts_node_words(void) { return kNodeWords; }

int ya_mcts_reset(const ya_mcts_tree* tree, const uint8_t* which, void* stream) {
    if (!tree_ok(tree)) return (int)cudaErrorInvalidValue;
    ya_k_mcts_reset<<<(unsigned)tree->n, 128, 0, (cudaStream_t)stream>>>(*tree, which);
    return (int)cudaGetLastError();
}

int ya_mcts_select(const ya_mcts_tree* tree, const uint32_t* states, int64_t stride, const int8_t* players,
                   const int32_t* ply, const uint32_t* episode, uint64_t seed, uint64_t game_base, uint32_t sim,
                   const uint32_t* sim_ptr, const uint64_t* game_base_ptr, float cpuct, const uint8_t* active,
                   float* features, uint8_t* need_eval, uint32_t* leaf_states, int rows, uint64_t* leaf_dst,
                   uint32_t* leaf_desc, int32_t* err_flag, void* stream) {
    if (!tree_ok(tree) || !(cpuct >= 0.0f)) return (int)cudaErrorInvalidValue;   // group maxima rely on u monotone in P
    if (rows != YA_ROWS_F32 && rows != YA_ROWS_FP16 && rows != YA_ROWS_BF16) return (int)cudaErrorInvalidValue;
    if (rows != YA_ROWS_F32 && ((tree->arena_words % 8) != 0 || (reinterpret_cast<uintptr_t>(tree->arena) & 31u)))
        return (int)cudaErrorMisalignedAddress;                                   // logit rows are 32-byte aligned
    if ((leaf_dst == nullptr) != (leaf_desc == nullptr)) return (int)cudaErrorInvalidValue;
    const dim3 grid(select_blocks(tree->n)), block(kWarpsPerBlock * 32);
    const cudaStream_t st = (cudaStream_t)stream;
    const uint4* sp = reinterpret_cast<const uint4*>(states);
#define YA_LAUNCH_SELECT(LEAF, ROWS)                                                                                      \
    ya_k_mcts_select<LEAF, false, ROWS><<<grid, block, 0, st>>>(*tree, sp, stride, players, ply, episode, seed, game_base, sim,   \
        sim_ptr, game_base_ptr, cpuct, active, features, need_eval, leaf_states, err_flag, nullptr, 0, leaf_dst, leaf_desc)
    if (rows == YA_ROWS_F32) {
        if (leaf_states) YA_LAUNCH_SELECT(true, ROWS_F32); else YA_LAUNCH_SELECT(false, ROWS_F32);
    } else if (rows == YA_ROWS_FP16) {
        if (leaf_states) YA_LAUNCH_SELECT(true, ROWS_L16F); else YA_LAUNCH_SELECT(false, ROWS_L16F);
    } else {
        if (leaf_states) YA_LAUNCH_SELECT(true, ROWS_L16B); else YA_LAUNCH_SELECT(false, ROWS_L16B);
    }
#undef YA_LAUNCH_SELECT
    return (int)cudaGetLastError();
}

int ya_mcts_select_injected(const ya_mcts_tree* tree, const uint32_t* states, int64_t stride, const int8_t* players,
                            uint32_t sim, float cpuct, const uint8_t* injected, int resume, float* features,
                            uint8_t* need_eval, uint32_t* leaf_states, int32_t* err_flag, void* stream) {
    if (!tree_ok(tree) || !injected || !leaf_states || !(cpuct >= 0.0f)) return (int)cudaErrorInvalidValue;
    ya_k_mcts_select<true, true, ROWS_F32><<<select_blocks(tree->n), kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
        *tree, reinterpret_cast<const uint4*>(states), stride, players, nullptr, nullptr, 0, 0, sim, nullptr, nullptr,
        cpuct, nullptr, features, need_eval, leaf_states, err_flag, injected, resume, nullptr, nullptr);
    return (int)cudaGetLastError();
}

int ya_mcts_expand(const ya_mcts_tree* tree, const float* pi, const float* value, int uniform, float uniform_p,
                   float uniform_v, uint32_t* sim_counter, int32_t* err_flag, void* stream) {
    if (!tree_ok(tree)) return (int)cudaErrorInvalidValue;
    if (uniform)
        ya_k_mcts_expand<1><<<warp_blocks(tree->n), kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
            *tree, nullptr, nullptr, uniform_p, uniform_v, sim_counter, err_flag);
    else
        ya_k_mcts_expand<0><<<warp_blocks(tree->n), kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
            *tree, pi, value, 0.0f, 0.0f, sim_counter, err_flag);
    return (int)cudaGetLastError();
}

int ya_mcts_expand_logits(const ya_mcts_tree* tree, const void* logits16, int fp16, int64_t ld, const float* row_max,
                          const float* value, uint32_t* sim_counter, int32_t* err_flag, void* stream) {
    if (!tree_ok(tree) || !value) return (int)cudaErrorInvalidValue;
    if (logits16 && (ld < kLogitCols || (ld % 8) != 0 || (reinterpret_cast<uintptr_t>(logits16) & 15u))) return (int)cudaErrorInvalidValue;
    if (!logits16 && !row_max) return (int)cudaErrorInvalidValue;       // rows filled by ya_nn_forward: its row_max comes with them
    const int blocks = (int)((tree->n + kLogitWarps - 1) / kLogitWarps);
    const cudaStream_t st = (cudaStream_t)stream;
    const uint16_t* lg = static_cast<const uint16_t*>(logits16);
    if (fp16) {
        if (lg) ya_k_mcts_expand_rows<true, true><<<blocks, kLogitWarps * 32, 0, st>>>(*tree, lg, ld, row_max, value, sim_counter, err_flag);
        else ya_k_mcts_expand_rows<true, false><<<blocks, kLogitWarps * 32, 0, st>>>(*tree, lg, ld, row_max, value, sim_counter, err_flag);
    } else {
        if (lg) ya_k_mcts_expand_rows<false, true><<<blocks, kLogitWarps * 32, 0, st>>>(*tree, lg, ld, row_max, value, sim_counter, err_flag);
        else ya_k_mcts_expand_rows<false, false><<<blocks, kLogitWarps * 32, 0, st>>>(*tree, lg, ld, row_max, value, sim_counter, err_flag);
    }
    return (int)cudaGetLastError();
}

int ya_mcts_search_uniform(const ya_mcts_tree* tree, const uint32_t* states, int64_t stride, const int8_t* players,
                           const int32_t* ply, const uint32_t* episode, uint64_t seed, uint64_t game_base, int num_sims,
                           float cpuct, float uniform_p, float uniform_v, const uint8_t* active, int32_t* err_flag,
                           void* stream) {
    if (!tree_ok(tree) || num_sims < 0 || !(cpuct >= 0.0f)) return (int)cudaErrorInvalidValue;
    ya_k_mcts_search_uniform<<<warp_blocks(tree->n), kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
        *tree, reinterpret_cast<const uint4*>(states), stride, players, ply, episode, seed, game_base, num_sims, cpuct,
        uniform_p, uniform_v, active, err_flag);
    return (int)cudaGetLastError();
}

int ya_mcts_root_counts(const ya_mcts_tree* tree, const uint32_t* states, int64_t stride, const int8_t* players,
                        int32_t* counts, int32_t* visits, double* qvals, uint8_t* qkind, void* stream) {
    if (!tree_ok(tree)) return (int)cudaErrorInvalidValue;
    ya_k_mcts_root_counts<<<warp_blocks(tree->n), kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
        *tree, reinterpret_cast<const uint4*>(states), stride, players, counts, visits, qvals, qkind);
    return (int)cudaGetLastError();
}

int ya_mcts_root_sparse(const ya_mcts_tree* tree, const uint32_t* states, int64_t stride, const int8_t* players,
                        int k, int16_t* actions, int32_t* counts, int32_t* overflow, void* stream) {
    if (!tree_ok(tree) || k <= 0) return (int)cudaErrorInvalidValue;
    ya_k_mcts_root_sparse<<<warp_blocks(tree->n), kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
        *tree, reinterpret_cast<const uint4*>(states), stride, players, k, actions, counts, overflow);
    return (int)cudaGetLastError();
}

int ya_mcts_pick_action(const int32_t* counts, const int32_t* ply, const uint32_t* episode, int64_t n, uint64_t seed,
                        uint64_t game_base, int temp_threshold, int32_t* actions, void* stream) {
    if (n <= 0) return 0;
    ya_k_mcts_pick<<<warp_blocks(n), kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
        counts, ply, episode, n, seed, game_base, temp_threshold, actions);
    return (int)cudaGetLastError();
}

}  // extern "C"
