// Yacht-Auction B200 engine -- packed state layout, Philox draw protocol and the rule
// primitives shared by every kernel (sm_100a only; no host fallback).
//
// Behavioural source of truth: /root/reference/yacht/YachtGame.py (cited per function).
//
// ---------------------------------------------------------------------------------------
// Packed state: 8 x u32 = 32 B per game, stored as two uint4 planes (structure of arrays):
//   plane0[g] = {w0,w1,w2,w3}   plane1[g] = {w4,w5,w6,w7}     plane1 = plane0 + stride
//
//   w0  [3:0]   round_no 1..13                       (YachtGame.py:135)
//       [4]     phase 0=BID 1=SCORE                  (:136)
//       [13:5]  p1_bid  : [5] present [6] target(0=A,1=B) [13:7] amount/500 (0..100)   (:141)
//       [22:14] p2_bid  : same layout                                                   (:142)
//   w1  [14:0]  rollA, die i at bits 3i (values 1..6, 0 = absent)   [29:15] rollB        (:138-139)
//   w2  p1.carry, die j at bits 3j, j = 0..9 in list order, 0 = empty slot              (:118)
//   w3  p2.carry
//   w4  p1: [11:0] used_mask  [24:12] bid_score/500 (13-bit two's complement)  [29:25] cat 8 (FULL_HOUSE) /1000
//   w5  p1: [17:0] cats 0..5 as face counts (3 bits each)  [22:18] cat 6 /1000  [27:23] cat 7 /1000
//           [28] cat 9 scored 15000  [29] cat 10 scored 30000  [30] cat 11 scored 50000  (:121)
//   w6,w7  p2, same as w4,w5
// The packing is injective on every state reachable by legal play (carry sizes 0/5/10); it
// is the MCTS node key in place of stringRepresentation (YachtGame.py:448-467).
// ---------------------------------------------------------------------------------------
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define YA_N_CAT 12
#define YA_N_BID_LEVEL 101
#define YA_N_BID 202
#define YA_N_SUBSET 252
#define YA_N_ACTION 3226
#define YA_N_FEATURE 59
#define YA_LAST_ROUND 13

// status codes written per game (0 = ok).  The Python host maps them onto the reference's
// exceptions (ValueError / RuntimeError / AssertionError, YachtGame.py:268-269,306-307,372,508).
#define YA_OK 0
#define YA_ERR_BID_RANGE 1
#define YA_ERR_SCORE_RANGE 2
#define YA_ERR_PHASE 3
#define YA_ERR_BID_ASSERT 4
#define YA_ERR_CARRY_OVERFLOW 5
#define YA_NEED_TIE 0x100
#define YA_NEED_ROLLS 0x200

// draw tags (oracle/philox.py restates the same protocol independently)
#define YA_TAG_INIT 0
#define YA_TAG_REAL 1
#define YA_TAG_ACTION 2
#define YA_TAG_SEARCH 3

struct YaState {
    uint32_t w[8];
};

struct YaDraw {          // one draw event
    uint32_t roll_a;     // 5 dice, 3 bits each
    uint32_t roll_b;
    uint32_t tie;        // 0 / 1
    uint32_t pick;       // raw 32-bit word for uniform index selection
};

__device__ const uint16_t ya_subset_mask[YA_N_SUBSET] = {
#include "ya_tables.inc"
};

// ------------------------------------------------------------------ load / store
__device__ __forceinline__ YaState ya_load(const uint4* __restrict__ base, int64_t stride, int64_t g) {
    uint4 a = base[g];
    uint4 b = base[stride + g];
    YaState s;
    s.w[0] = a.x; s.w[1] = a.y; s.w[2] = a.z; s.w[3] = a.w;
    s.w[4] = b.x; s.w[5] = b.y; s.w[6] = b.z; s.w[7] = b.w;
    return s;
}

__device__ __forceinline__ void ya_store(uint4* __restrict__ base, int64_t stride, int64_t g, const YaState& s) {
    base[g] = make_uint4(s.w[0], s.w[1], s.w[2], s.w[3]);
    base[stride + g] = make_uint4(s.w[4], s.w[5], s.w[6], s.w[7]);
}

// ------------------------------------------------------------------ field helpers
__device__ __forceinline__ int ya_round(const YaState& s) { return s.w[0] & 15; }
__device__ __forceinline__ int ya_phase(const YaState& s) { return (s.w[0] >> 4) & 1; }
__device__ __forceinline__ uint32_t ya_bid_slot(const YaState& s, int i) { return (s.w[0] >> (5 + 9 * i)) & 0x1FF; }
__device__ __forceinline__ int ya_dice_count(uint32_t c) {
    return __popc((c | (c >> 1) | (c >> 2)) & 0x09249249u);
}
__device__ __forceinline__ int ya_bank(uint32_t w4) { return ((int)(w4 << 7)) >> 19; }
__device__ __forceinline__ uint32_t ya_set_bank(uint32_t w4, int bank) {
    return (w4 & ~(0x1FFFu << 12)) | (((uint32_t)bank & 0x1FFFu) << 12);
}
__device__ __forceinline__ bool ya_bidding(const YaState& s) { return ya_phase(s) == 0 && ya_round(s) != YA_LAST_ROUND; }

// Swap p1<->p2 and their pending bids: getCanonicalForm(board, -1), YachtGame.py:430-442.
__device__ __forceinline__ YaState ya_flip(const YaState& s) {
    YaState o;
    uint32_t b0 = ya_bid_slot(s, 0), b1 = ya_bid_slot(s, 1);
    o.w[0] = (s.w[0] & 0x1Fu) | (b1 << 5) | (b0 << 14);
    o.w[1] = s.w[1];
    o.w[2] = s.w[3]; o.w[3] = s.w[2];
    o.w[4] = s.w[6]; o.w[5] = s.w[7];
    o.w[6] = s.w[4]; o.w[7] = s.w[5];
    return o;
}

// ------------------------------------------------------------------ Philox4x32-10 draw protocol
__device__ __forceinline__ void ya_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                          uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        c0 = h1 ^ c1 ^ k0;
        c1 = l1;
        c2 = h0 ^ c3 ^ k1;
        c3 = l0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ uint32_t ya_five_dice(uint32_t w) {
    uint32_t out = 0;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        out |= (1u + __umulhi(w, 6u)) << (3 * i);
        w *= 6u;
    }
    return out;
}

__device__ __forceinline__ YaDraw ya_draw(uint64_t seed, uint32_t game, uint32_t episode, uint32_t ply,
                                          uint32_t tag, uint32_t depth, uint32_t sim) {
    uint32_t o[4];
    ya_philox(game, episode, (ply & 0xFF) | ((tag & 0xFF) << 8) | ((depth & 0xFF) << 16), sim,
              (uint32_t)seed, (uint32_t)(seed >> 32), o);
    YaDraw d;
    d.roll_a = ya_five_dice(o[0]);
    d.roll_b = ya_five_dice(o[1]);
    d.tie = o[2] >> 31;
    d.pick = o[3];
    return d;
}

// ------------------------------------------------------------------ scoring (YachtGame.py:57-108)
// hist: six 4-bit face counters (face f at nibble f-1), pips: sum of the five dice.
// Returns the category score / 1000 via nibble bit tricks (no per-face loops).
__device__ __forceinline__ uint32_t ya_nibble_eq(uint32_t h, uint32_t v) {   // 0x1 per nibble == v (v in 0..7, nibbles <= 5)
    uint32_t x = h ^ (v * 0x111111u);
    // nibble zero test: (x | x>>1 | x>>2 | x>>3) & 1 per nibble is 0
    uint32_t nz = (x | (x >> 1) | (x >> 2) | (x >> 3)) & 0x111111u;
    return nz ^ 0x111111u;
}

__device__ __forceinline__ uint32_t ya_category_points_k(int cat, uint32_t hist, uint32_t pips) {
    if (cat < 6) return (uint32_t)(cat + 1) * ((hist >> (4 * cat)) & 0xF);
    if (cat == 6) return pips;
    if (cat == 7) return (((hist + 0x444444u) & 0x888888u) != 0) ? pips : 0;            // any count >= 4
    uint32_t e5 = ya_nibble_eq(hist, 5);
    if (cat == 8) {
        bool pair = (ya_nibble_eq(hist, 2) | e5) != 0;
        bool trip = (ya_nibble_eq(hist, 3) | e5) != 0;
        return (pair && trip) ? pips : 0;
    }
    if (cat == 11) return e5 ? 50u : 0u;
    // presence bits, face f -> bit f-1
    uint32_t nz = (hist | (hist >> 1) | (hist >> 2) | (hist >> 3)) & 0x111111u;
    uint32_t seen = (nz & 1u) | ((nz >> 3) & 2u) | ((nz >> 6) & 4u) | ((nz >> 9) & 8u) | ((nz >> 12) & 16u) | ((nz >> 15) & 32u);
    if (cat == 9) {
        bool ok = ((seen & 0x0Fu) == 0x0Fu) || ((seen & 0x1Eu) == 0x1Eu) || ((seen & 0x3Cu) == 0x3Cu);
        return ok ? 15u : 0u;
    }
    bool ok = ((seen & 0x1Fu) == 0x1Fu) || ((seen & 0x3Eu) == 0x3Eu);
    return ok ? 30u : 0u;
}

// Gather the dice at the positions of a 10-bit subset mask: histogram + pip sum.
__device__ __forceinline__ void ya_gather(uint32_t carry, uint32_t m, uint32_t& hist, uint32_t& pips) {
    hist = 0; pips = 0;
#pragma unroll
    for (int j = 0; j < 10; ++j) {
        uint32_t d = (carry >> (3 * j)) & 7u;
        if ((m >> j) & 1u) {
            hist += 1u << (4 * (d - 1));
            pips += d;
        }
    }
}

// write cat_scores[cat] = k*1000 into the packed score words
__device__ __forceinline__ void ya_put_score(uint32_t& w4, uint32_t& w5, int cat, uint32_t k, uint32_t hist) {
    if (cat < 6)        w5 |= ((hist >> (4 * cat)) & 0xFu) << (3 * cat);
    else if (cat == 6)  w5 |= k << 18;
    else if (cat == 7)  w5 |= k << 23;
    else if (cat == 8)  w4 |= k << 25;
    else                w5 |= (k ? 1u : 0u) << (28 + (cat - 9));
}

// cat_scores[cat] / 1000 read back from the packed words
__device__ __forceinline__ uint32_t ya_get_score_k(uint32_t w4, uint32_t w5, int cat) {
    if (cat < 6)  return (uint32_t)(cat + 1) * ((w5 >> (3 * cat)) & 7u);
    if (cat == 6) return (w5 >> 18) & 31u;
    if (cat == 7) return (w5 >> 23) & 31u;
    if (cat == 8) return (w4 >> 25) & 31u;
    if (cat == 9) return ((w5 >> 28) & 1u) * 15u;
    if (cat == 10) return ((w5 >> 29) & 1u) * 30u;
    return ((w5 >> 30) & 1u) * 50u;
}

// total_with_bonus() in units of 500 (YachtGame.py:125-130)
__device__ __forceinline__ int ya_total_500(uint32_t w4, uint32_t w5) {
    uint32_t upper = 0;
#pragma unroll
    for (int c = 0; c < 6; ++c) upper += (uint32_t)(c + 1) * ((w5 >> (3 * c)) & 7u);
    uint32_t rest = ((w5 >> 18) & 31u) + ((w5 >> 23) & 31u) + ((w4 >> 25) & 31u) +
                    ((w5 >> 28) & 1u) * 15u + ((w5 >> 29) & 1u) * 30u + ((w5 >> 30) & 1u) * 50u;
    uint32_t k = upper + rest + (upper >= 63u ? 35u : 0u);
    return 2 * (int)k + ya_bank(w4);
}

// getGameEnded(board, player): YachtGame.py:408-428.  player is +1 / -1.
__device__ __forceinline__ float ya_game_ended(const YaState& s, int player) {
    if ((s.w[4] & 0xFFFu) != 0xFFFu || (s.w[6] & 0xFFFu) != 0xFFFu) return 0.0f;
    int t0 = ya_total_500(s.w[4], s.w[5]);
    int t1 = ya_total_500(s.w[6], s.w[7]);
    if (t0 == t1) return 1e-4f;
    float lead = t0 > t1 ? 1.0f : -1.0f;
    return player == 1 ? lead : -lead;
}

// ------------------------------------------------------------------ legality (YachtGame.py:374-406)
// Mask descriptor for the player to move: bit 0 = all 202 bids legal; bits 1..12 = category
// c-1 open; bit 13 = all 252 subsets fit (10 dice) else only subset 0 (5 dice).
__device__ __forceinline__ uint32_t ya_mask_desc(const YaState& s, int player) {
    if (ya_bidding(s)) return 1u;
    if (ya_phase(s) != 1) return 0u;
    int me = player == 1 ? 0 : 1;
    int n = ya_dice_count(s.w[2 + me]);
    if (n < 5) return 0u;
    uint32_t open = (~s.w[4 + 2 * me]) & 0xFFFu;
    return (open << 1) | (n >= 10 ? (1u << 13) : 0u);
}

__device__ __forceinline__ int ya_legal_count(uint32_t desc) {
    if (desc & 1u) return YA_N_BID;
    return __popc((desc >> 1) & 0xFFFu) * ((desc >> 13) ? YA_N_SUBSET : 1);
}

// idx-th legal action in ascending action order
__device__ __forceinline__ int ya_nth_legal(uint32_t desc, int idx) {
    if (desc & 1u) return idx;
    uint32_t open = (desc >> 1) & 0xFFFu;
    int per = (desc >> 13) ? YA_N_SUBSET : 1;
    int k = idx / per, sub = idx - k * per;
    int cat = __fns(open, 0, k + 1);
    return YA_N_BID + cat * YA_N_SUBSET + sub;
}

// position of an action inside the legal list (inverse of ya_nth_legal); action must be legal
__device__ __forceinline__ int ya_legal_index(uint32_t desc, int action) {
    if (desc & 1u) return action;
    uint32_t open = (desc >> 1) & 0xFFFu;
    int q = action - YA_N_BID;
    int cat = q / YA_N_SUBSET, sub = q - cat * YA_N_SUBSET;
    int k = __popc(open & ((1u << cat) - 1u));
    return (desc >> 13) ? k * YA_N_SUBSET + sub : k;
}

__device__ __forceinline__ bool ya_is_legal(uint32_t desc, int action) {
    if (action < 0 || action >= YA_N_ACTION) return false;
    if (action < YA_N_BID) return desc & 1u;
    int q = action - YA_N_BID;
    int cat = q / YA_N_SUBSET, sub = q - cat * YA_N_SUBSET;
    if (!((desc >> (1 + cat)) & 1u)) return false;
    return (desc >> 13) ? true : sub == 0;
}

// ------------------------------------------------------------------ transition (YachtGame.py:260-372, 502-542)
// Which random draws would getNextState consume for (s, player, action)?  Returns a bitset of
// YA_NEED_TIE / YA_NEED_ROLLS (consumption order inside one call: tie, rollA, rollB).
__device__ __forceinline__ int ya_draw_needs(const YaState& s, int player, int action) {
    int me = player == 1 ? 0 : 1;
    int r = ya_round(s);
    if (ya_bidding(s)) {
        if (action < 0 || action >= YA_N_BID) return 0;
        uint32_t b0 = ya_bid_slot(s, 0), b1 = ya_bid_slot(s, 1);
        if (!((b0 | b1) & 1u)) return 0;                         // first bidder
        uint32_t mine = 1u | ((uint32_t)(action / YA_N_BID_LEVEL) << 1) | ((uint32_t)(action % YA_N_BID_LEVEL) << 2);
        uint32_t other = me == 0 ? b1 : b0;
        int needs = 0;
        if ((other & 1u) && mine == other) needs |= YA_NEED_TIE;  // same target, same amount
        if (r == 1) needs |= YA_NEED_ROLLS;
        return needs;
    }
    if (ya_phase(s) == 1) {
        if (action < YA_N_BID || action >= YA_N_ACTION) return 0;
        if (r == YA_LAST_ROUND || player != -1 || r + 1 == YA_LAST_ROUND) return 0;
        int q = action - YA_N_BID;
        int cat = q / YA_N_SUBSET;
        uint32_t m = ya_subset_mask[q - cat * YA_N_SUBSET];
        int n = ya_dice_count(s.w[2 + me]);
        if (((s.w[4 + 2 * me] >> cat) & 1u) || (31 - __clz(m)) >= n) return 0;   // silent no-op
        return YA_NEED_ROLLS;
    }
    return 0;
}

// Applies one ply in place.  `draw` must hold whatever ya_draw_needs() reported.
// Returns next_player (+1/-1); *status receives YA_OK or an error (state left unchanged on error).
__device__ __forceinline__ int ya_transition(YaState& s, int player, int action, const YaDraw& draw, int* status) {
    *status = YA_OK;
    const int me = player == 1 ? 0 : 1;
    const int r = ya_round(s);
    if (ya_bidding(s)) {
        if (action < 0 || action >= YA_N_BID) { *status = YA_ERR_BID_RANGE; return player; }
        uint32_t b0 = ya_bid_slot(s, 0), b1 = ya_bid_slot(s, 1);
        const bool first = !((b0 | b1) & 1u);
        uint32_t mine = 1u | ((uint32_t)(action / YA_N_BID_LEVEL) << 1) | ((uint32_t)(action % YA_N_BID_LEVEL) << 2);
        if (me == 0) b0 = mine; else b1 = mine;
        if (first) {                                               // :272-279
            s.w[0] = (s.w[0] & 0x1Fu) | (b0 << 5) | (b1 << 14);
            return -player;
        }
        if (!(b0 & b1 & 1u)) { *status = YA_ERR_BID_ASSERT; return player; }      // :508
        // ---- _resolve_bids_and_assign (:502-542)
        int t0 = (b0 >> 1) & 1, a0 = b0 >> 2, t1 = (b1 >> 1) & 1, a1 = b1 >> 2;
        int g0 = t0, g1 = t1;
        if (t0 == t1) {
            int win = a0 > a1 ? 0 : (a1 > a0 ? 1 : (int)draw.tie);
            if (win == 0) g1 = 1 - g0; else g0 = 1 - g1;
        }
        int n0 = ya_dice_count(s.w[2]), n1 = ya_dice_count(s.w[3]);
        if (n0 > 5 || n1 > 5) { *status = YA_ERR_CARRY_OVERFLOW; return player; }
        s.w[4] = ya_set_bank(s.w[4], ya_bank(s.w[4]) + (g0 == t0 ? -a0 : a0));
        s.w[6] = ya_set_bank(s.w[6], ya_bank(s.w[6]) + (g1 == t1 ? -a1 : a1));
        uint32_t pool_a = s.w[1] & 0x7FFFu, pool_b = (s.w[1] >> 15) & 0x7FFFu;
        s.w[2] |= (g0 == 0 ? pool_a : pool_b) << (3 * n0);
        s.w[3] |= (g1 == 0 ? pool_a : pool_b) << (3 * n1);
        if (r != 1) {                                              // :290-293
            s.w[0] = (s.w[0] & 0xFu) | (1u << 4) | (b0 << 5) | (b1 << 14);
        } else {                                                   // :295-301
            s.w[0] = 2u;
            s.w[1] = draw.roll_a | (draw.roll_b << 15);
        }
        return 1;
    }
    if (ya_phase(s) == 1) {
        if (action < YA_N_BID || action >= YA_N_ACTION) { *status = YA_ERR_SCORE_RANGE; return player; }
        int q = action - YA_N_BID;
        int cat = q / YA_N_SUBSET;
        uint32_t m = ya_subset_mask[q - cat * YA_N_SUBSET];
        uint32_t carry = s.w[2 + me];
        int n = ya_dice_count(carry);
        uint32_t& w4 = s.w[4 + 2 * me];
        uint32_t& w5 = s.w[5 + 2 * me];
        if (((w4 >> cat) & 1u) || (31 - __clz(m)) >= n) return -player;           // :312-324 silent no-op
        uint32_t hist, pips;
        ya_gather(carry, m, hist, pips);
        uint32_t k = ya_category_points_k(cat, hist, pips);
        ya_put_score(w4, w5, cat, k, hist);
        w4 |= 1u << cat;
        uint32_t kept = 0; int out = 0;
#pragma unroll
        for (int j = 0; j < 10; ++j) {
            uint32_t d = (carry >> (3 * j)) & 7u;
            if (!((m >> j) & 1u) && d) { kept |= d << (3 * out); ++out; }
        }
        s.w[2 + me] = kept;
        if (r == YA_LAST_ROUND) {                                  // :338-349
            bool done = (s.w[4] & 0xFFFu) == 0xFFFu && (s.w[6] & 0xFFFu) == 0xFFFu;
            return done ? 1 : -player;
        }
        if (player == -1) {                                        // :352-365
            if (r + 1 != YA_LAST_ROUND) {
                s.w[0] = (uint32_t)(r + 1);
                s.w[1] = draw.roll_a | (draw.roll_b << 15);
            } else {
                s.w[0] = (uint32_t)(r + 1) | (1u << 4);
            }
            return 1;
        }
        return -player;                                            // :366-369
    }
    *status = YA_ERR_PHASE;                                        // :372
    return player;
}

// ------------------------------------------------------------------ features (yacht/NNet.py:50-86)
// state_to_vec computes (d - 3.5) / 3.5 and round / 13 in Python doubles and numpy rounds them to float32.  Those
// are twenty distinct numbers: they are spelled out below as the float32 bit patterns of exactly that computation
// (float32((d - 3.5) / 3.5) for d = 1..6 is +-{5, 3, 1} / 7; float32(r / 13.0) for r = 0..13), so a feature costs a
// few selects instead of a double-precision division, and the lanes of a team writing different features of one
// state do not diverge into a dozen branches.
__device__ __forceinline__ float ya_die_feature(uint32_t d) {
    const int a = abs(2 * (int)d - 7);                                // 5, 3, 1 for d = 1|6, 2|5, 3|4
    const uint32_t mag = a == 5 ? 0x3F36DB6Eu : (a == 3 ? 0x3EDB6DB7u : 0x3E124925u);
    const float v = __uint_as_float(mag | (d < 4u ? 0x80000000u : 0u));
    return d ? v : -1.0f;                                             // empty slot
}

__device__ __forceinline__ float ya_round_feature(int r) {           // float32(r / 13.0)
    constexpr uint32_t k[16] = {0x00000000u, 0x3D9D89D9u, 0x3E1D89D9u, 0x3E6C4EC5u, 0x3E9D89D9u, 0x3EC4EC4Fu, 0x3EEC4EC5u, 0x3F09D89Eu,
                                0x3F1D89D9u, 0x3F313B14u, 0x3F44EC4Fu, 0x3F589D8Au, 0x3F6C4EC5u, 0x3F800000u, 0x3F800000u, 0x3F800000u};
    uint32_t bits = 0;
#pragma unroll
    for (int i = 0; i < 14; ++i) bits = r == i ? k[i] : bits;
    return __uint_as_float(bits);
}

__device__ __forceinline__ float ya_feature(const YaState& s, int f) {
    // dice features 3..32: ten dice of each carry, then the five dice of each open bundle (bidding phases only)
    const bool pool = f >= 23;
    const uint32_t dice_word = f < 13 ? s.w[2] : (pool ? s.w[1] : s.w[3]);
    const int dice_idx = f < 13 ? f - 3 : (pool ? f - 23 : f - 13);
    uint32_t d = (dice_word >> (3 * (dice_idx & 15))) & 7u;
    if (pool && !ya_bidding(s)) d = 0u;
    const float die = ya_die_feature(d);
    // used-category bits 33..56
    const uint32_t used_word = f < 45 ? s.w[4] : s.w[6];
    const float bit = (float)((used_word >> ((f < 45 ? f - 33 : f - 45) & 15)) & 1u);
    float out = f < 33 ? die : bit;
    if (f >= 57) out = (float)((double)(ya_bank(f == 57 ? s.w[4] : s.w[6]) * 500) * 1e-5);
    if (f < 3) out = f == 0 ? ya_round_feature(ya_round(s)) : (ya_phase(s) == f - 1 ? 1.0f : 0.0f);
    return out;
}
