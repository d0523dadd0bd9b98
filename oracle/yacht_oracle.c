/* TEST INFRASTRUCTURE ONLY -- bulk CPU oracle (plain C) of the Yacht-Auction rules.
 *
 * A second, independent restatement of /root/reference/yacht/YachtGame.py (lines cited per function)
 * on an unpacked struct, plus the Philox draw protocol and the random-legal policy, so that whole
 * batches of games (10^4 .. 10^6 plies) can be replayed on the host in milliseconds and compared with
 * the CUDA path.  It is pinned against the pure-Python oracle and the golden vectors produced by the
 * reference (tests/test_c_oracle.py).  Nothing in the product links or loads this file; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may.
 *
 * Build: make -C oracle   ->  oracle/_build/libyacht_oracle.so
 */
#include <stdint.h>
#include <string.h>
#include <stdio.h>

#define N_CAT 12
#define N_BID 202
#define N_SUBSET 252
#define N_ACTION 3226
#define LAST_ROUND 13

typedef struct {
    int ndice;
    int dice[16];           /* ordered carry (yacht/YachtGame.py:118) */
    int used;               /* 12-bit mask (:120) */
    int cats[N_CAT];        /* points (:121) */
    int bank;               /* bid_score (:123) */
} yo_side;

typedef struct {
    int rnd, phase;                 /* :135-136 */
    int pool_a[5], pool_b[5];       /* :138-139 */
    int bid_set[2], bid_target[2], bid_amount[2];   /* p1_bid / p2_bid (:141-142) */
    yo_side side[2];
} yo_state;

static int SUBSET_POS[N_SUBSET][5];
static int subsets_ready = 0;

static void init_subsets(void) {                    /* itertools.combinations(range(10), 5), :35 */
    if (subsets_ready) return;
    int k = 0;
    for (int a = 0; a < 10; ++a) for (int b = a + 1; b < 10; ++b) for (int c = b + 1; c < 10; ++c)
        for (int d = c + 1; d < 10; ++d) for (int e = d + 1; e < 10; ++e) {
            SUBSET_POS[k][0] = a; SUBSET_POS[k][1] = b; SUBSET_POS[k][2] = c; SUBSET_POS[k][3] = d; SUBSET_POS[k][4] = e;
            ++k;
        }
    subsets_ready = 1;
}

/* ------------------------------------------------------------------ Philox4x32-10 + draw protocol */
static void philox(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

typedef struct { int a[5], b[5], tie; uint32_t pick; } yo_draw;

static void five(uint32_t w, int out[5]) {
    for (int i = 0; i < 5; ++i) { uint64_t t = (uint64_t)w * 6u; out[i] = 1 + (int)(t >> 32); w = (uint32_t)t; }
}

static yo_draw make_draw(uint64_t seed, uint32_t game, uint32_t episode, uint32_t ply, uint32_t tag, uint32_t depth, uint32_t sim) {
    uint32_t c[4] = {game, episode, (ply & 0xFF) | ((tag & 0xFF) << 8) | ((depth & 0xFF) << 16), sim};
    philox(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    yo_draw d;
    five(c[0], d.a); five(c[1], d.b); d.tie = (int)(c[2] >> 31); d.pick = c[3];
    return d;
}

void yo_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
    philox(c, key[0], key[1]);
    memcpy(out, c, sizeof c);
}

/* ------------------------------------------------------------------ score_category, :57-108 */
int yo_category_points(int cat, const int five_dice[5]) {
    int hist[7] = {0}, pips = 0, mx = 0;
    for (int i = 0; i < 5; ++i) { hist[five_dice[i]]++; pips += five_dice[i]; }
    for (int f = 1; f <= 6; ++f) if (hist[f] > mx) mx = hist[f];
    if (cat < 6) return 1000 * (cat + 1) * hist[cat + 1];
    if (cat == 6) return 1000 * pips;
    if (cat == 7) return mx >= 4 ? 1000 * pips : 0;
    if (cat == 8) {
        int two = 0, three = 0;
        for (int f = 1; f <= 6; ++f) { if (hist[f] == 2 || hist[f] == 5) two = 1; if (hist[f] == 3 || hist[f] == 5) three = 1; }
        return (two && three) ? 1000 * pips : 0;
    }
    int run = 0, best = 0;
    for (int f = 1; f <= 6; ++f) { run = hist[f] ? run + 1 : 0; if (run > best) best = run; }
    if (cat == 9) return best >= 4 ? 15000 : 0;
    if (cat == 10) return best >= 5 ? 30000 : 0;
    return mx == 5 ? 50000 : 0;
}

/* ------------------------------------------------------------------ getInitBoard, :232-237 */
static void new_game(yo_state* s, const yo_draw* d) {
    memset(s, 0, sizeof *s);
    s->rnd = 1; s->phase = 0;
    memcpy(s->pool_a, d->a, sizeof s->pool_a);
    memcpy(s->pool_b, d->b, sizeof s->pool_b);
}

static int side_total(const yo_side* p) {           /* :125-130 */
    int upper = 0, all = 0;
    for (int c = 0; c < N_CAT; ++c) { all += p->cats[c]; if (c < 6) upper += p->cats[c]; }
    return all + (upper >= 63000 ? 35000 : 0) + p->bank;
}

/* getGameEnded(board, player), :408-428: 0, +1, -1, or 2 for the draw value 1e-4 */
static int outcome_code(const yo_state* s, int player) {
    if (s->side[0].used != 4095 || s->side[1].used != 4095) return 0;
    int t0 = side_total(&s->side[0]), t1 = side_total(&s->side[1]);
    if (t0 == t1) return 2;
    int lead = t0 > t1 ? 1 : -1;
    return player == 1 ? lead : -lead;
}

/* _resolve_bids_and_assign, :502-542 */
static void settle(yo_state* s, const yo_draw* d) {
    int t0 = s->bid_target[0], a0 = s->bid_amount[0], t1 = s->bid_target[1], a1 = s->bid_amount[1];
    int got[2] = {t0, t1};
    if (t0 == t1) {
        int win = a0 > a1 ? 0 : (a1 > a0 ? 1 : d->tie);
        got[1 - win] = 1 - got[win];
    }
    s->side[0].bank += got[0] == t0 ? -a0 : a0;
    s->side[1].bank += got[1] == t1 ? -a1 : a1;
    for (int p = 0; p < 2; ++p) {
        const int* pool = got[p] == 0 ? s->pool_a : s->pool_b;
        for (int i = 0; i < 5; ++i) s->side[p].dice[s->side[p].ndice++] = pool[i];
    }
}

/* getNextState, :260-372.  Returns next player, or 0 on the reference's exceptions (status set). */
static int next_state(yo_state* s, int player, int action, const yo_draw* d, int* status) {
    int me = player == 1 ? 0 : 1;
    *status = 0;
    if (s->phase == 0 && s->rnd != LAST_ROUND) {
        if (action < 0 || action >= N_BID) { *status = 1; return 0; }
        int first = !s->bid_set[0] && !s->bid_set[1];
        s->bid_set[me] = 1; s->bid_target[me] = action / 101; s->bid_amount[me] = (action % 101) * 500;
        if (first) return -player;
        if (!s->bid_set[0] || !s->bid_set[1]) { *status = 4; return 0; }
        if (s->side[0].ndice > 5 || s->side[1].ndice > 5) { *status = 5; return 0; }
        settle(s, d);
        if (s->rnd != 1) { s->phase = 1; return 1; }
        s->rnd += 1;
        s->bid_set[0] = s->bid_set[1] = 0;
        memcpy(s->pool_a, d->a, sizeof s->pool_a);
        memcpy(s->pool_b, d->b, sizeof s->pool_b);
        return 1;
    }
    if (s->phase == 1) {
        if (action < N_BID || action >= N_ACTION) { *status = 2; return 0; }
        int cat = (action - N_BID) / N_SUBSET, sub = (action - N_BID) % N_SUBSET;
        const int* pos = SUBSET_POS[sub];
        yo_side* p = &s->side[me];
        if (((p->used >> cat) & 1) || pos[4] >= p->ndice) return -player;         /* silent no-op, :312-324 */
        int chosen[5], keep[16], nk = 0;
        for (int i = 0; i < 5; ++i) chosen[i] = p->dice[pos[i]];
        for (int i = 0, j = 0; i < p->ndice; ++i) {
            if (j < 5 && pos[j] == i) { ++j; continue; }
            keep[nk++] = p->dice[i];
        }
        p->cats[cat] = yo_category_points(cat, chosen);
        p->used |= 1 << cat;
        p->ndice = nk;
        memcpy(p->dice, keep, sizeof(int) * nk);
        if (s->rnd == LAST_ROUND)
            return (s->side[0].used == 4095 && s->side[1].used == 4095) ? 1 : -player;
        if (player == -1) {
            s->rnd += 1;
            s->bid_set[0] = s->bid_set[1] = 0;
            if (s->rnd != LAST_ROUND) {
                memcpy(s->pool_a, d->a, sizeof s->pool_a);
                memcpy(s->pool_b, d->b, sizeof s->pool_b);
                s->phase = 0;
            } else {
                s->phase = 1;
            }
            return 1;
        }
        return -player;
    }
    *status = 3;
    return 0;
}

/* getValidMoves, :374-406 */
static int legal_mask(const yo_state* s, int player, uint8_t* v) {
    int count = 0;
    memset(v, 0, N_ACTION);
    if (s->phase == 0 && s->rnd != LAST_ROUND) { memset(v, 1, N_BID); return N_BID; }
    if (s->phase != 1) return 0;
    const yo_side* p = &s->side[player == 1 ? 0 : 1];
    if (p->ndice < 5) return 0;
    for (int cat = 0; cat < N_CAT; ++cat) {
        if ((p->used >> cat) & 1) continue;
        for (int sub = 0; sub < N_SUBSET; ++sub)
            if (SUBSET_POS[sub][4] < p->ndice) { v[N_BID + cat * N_SUBSET + sub] = 1; ++count; }
    }
    return count;
}

/* ------------------------------------------------------------------ packed layout, written
 * from the documented bit layout, independently of the CUDA headers */
static uint32_t pack_dice(const int* d, int n) { uint32_t w = 0; for (int i = 0; i < n; ++i) w |= (uint32_t)d[i] << (3 * i); return w; }

static void pack_state(const yo_state* s, uint32_t w[8]) {
    uint32_t b[2];
    for (int p = 0; p < 2; ++p)
        b[p] = s->bid_set[p] ? (1u | ((uint32_t)s->bid_target[p] << 1) | ((uint32_t)(s->bid_amount[p] / 500) << 2)) : 0u;
    w[0] = (uint32_t)s->rnd | ((uint32_t)s->phase << 4) | (b[0] << 5) | (b[1] << 14);
    w[1] = pack_dice(s->pool_a, 5) | (pack_dice(s->pool_b, 5) << 15);
    for (int p = 0; p < 2; ++p) {
        const yo_side* q = &s->side[p];
        w[2 + p] = pack_dice(q->dice, q->ndice);
        uint32_t w4 = (uint32_t)q->used | (((uint32_t)(q->bank / 500) & 0x1FFFu) << 12) | ((uint32_t)(q->cats[8] / 1000) << 25);
        uint32_t w5 = 0;
        for (int c = 0; c < 6; ++c) w5 |= (uint32_t)(q->cats[c] / (1000 * (c + 1))) << (3 * c);
        w5 |= (uint32_t)(q->cats[6] / 1000) << 18;
        w5 |= (uint32_t)(q->cats[7] / 1000) << 23;
        w5 |= (q->cats[9] ? 1u : 0u) << 28;
        w5 |= (q->cats[10] ? 1u : 0u) << 29;
        w5 |= (q->cats[11] ? 1u : 0u) << 30;
        w[4 + 2 * p] = w4;
        w[5 + 2 * p] = w5;
    }
}

/* stringRepresentation, :448-467 */
static int key_string(const yo_state* s, char* out, size_t cap) {
    char a[8] = "-", b[8] = "-", bids[2][16], dice[2][20], cats[2][128];
    if (s->pool_a[0]) { for (int i = 0; i < 5; ++i) a[i] = (char)('0' + s->pool_a[i]); a[5] = 0; }
    if (s->pool_b[0]) { for (int i = 0; i < 5; ++i) b[i] = (char)('0' + s->pool_b[i]); b[5] = 0; }
    for (int p = 0; p < 2; ++p) {
        if (s->bid_set[p]) snprintf(bids[p], sizeof bids[p], "%c%d", s->bid_target[p] ? 'B' : 'A', s->bid_amount[p]);
        else strcpy(bids[p], "-");
        for (int i = 0; i < s->side[p].ndice; ++i) dice[p][i] = (char)('0' + s->side[p].dice[i]);
        dice[p][s->side[p].ndice] = 0;
        int n = 0;
        for (int c = 0; c < N_CAT; ++c) n += snprintf(cats[p] + n, sizeof cats[p] - (size_t)n, c ? ",%d" : "%d", s->side[p].cats[c]);
    }
    return snprintf(out, cap, "r%d|ph%d|A%s|B%s|p1b%s|p2b%s|p1c%s|p2c%s|p1u%d|p2u%d|p1s%s|p2s%s|p1bid%d|p2bid%d",
                    s->rnd, s->phase, a, b, bids[0], bids[1], dice[0], dice[1], s->side[0].used, s->side[1].used,
                    cats[0], cats[1], s->side[0].bank, s->side[1].bank);
}

/* ------------------------------------------------------------------ batched driver
 * Plays `plies` plies of n games (global ids game_base + g) under the uniform random-legal policy with the
 * Philox draw protocol, exactly what ya_play_ply does.  Optional outputs per (ply, game):
 *   packed   uint32[plies][n][8]  packed state AFTER the ply (before a re-deal)
 *   actions  int32 [plies][n]
 *   legal    int32 [plies][n]     number of legal actions BEFORE the ply
 *   result   int8  [plies][n]     outcome code for player 1 after the ply (0 running, +1, -1, 2 draw)
 * With auto_reset a finished game is re-dealt (episode + 1).  Returns the number of plies played. */
long long yo_play_random(int64_t n, int plies, uint64_t seed, uint64_t game_base, int auto_reset,
                         uint32_t* packed, int32_t* actions, int32_t* legal, int8_t* result,
                         char* final_keys, int key_cap) {
    init_subsets();
    long long steps = 0;
#pragma omp parallel for schedule(static) reduction(+ : steps)
    for (int64_t g = 0; g < n; ++g) {
        uint32_t gid = (uint32_t)(game_base + (uint64_t)g);
        uint32_t episode = 0, ply = 0;
        yo_state s;
        yo_draw d0 = make_draw(seed, gid, 0, 0, 0, 0, 0);
        new_game(&s, &d0);
        int cur = 1, done = 0;
        uint8_t mask[N_ACTION];
        for (int t = 0; t < plies; ++t) {
            if (done) {
                if (packed) pack_state(&s, packed + ((int64_t)t * n + g) * 8);
                if (actions) actions[(int64_t)t * n + g] = -1;
                if (legal) legal[(int64_t)t * n + g] = 0;
                if (result) result[(int64_t)t * n + g] = (int8_t)outcome_code(&s, 1);
                continue;
            }
            int count = legal_mask(&s, cur, mask);
            int action = 0;
            if (count) {
                yo_draw da = make_draw(seed, gid, episode, ply, 2, 0, 0);
                int idx = (int)(((uint64_t)da.pick * (uint64_t)count) >> 32);
                for (int a = 0; a < N_ACTION; ++a) if (mask[a] && idx-- == 0) { action = a; break; }
            }
            yo_draw d = make_draw(seed, gid, episode, ply, 1, 0, 0);
            int status;
            int nxt = next_state(&s, cur, action, &d, &status);
            if (status) nxt = cur;
            cur = nxt;
            ++ply;
            ++steps;
            int res = outcome_code(&s, 1);
            if (packed) pack_state(&s, packed + ((int64_t)t * n + g) * 8);
            if (actions) actions[(int64_t)t * n + g] = action;
            if (legal) legal[(int64_t)t * n + g] = count;
            if (result) result[(int64_t)t * n + g] = (int8_t)res;
            if (res) {
                if (auto_reset) {
                    ++episode; ply = 0; cur = 1;
                    yo_draw dn = make_draw(seed, gid, episode, 0, 0, 0, 0);
                    new_game(&s, &dn);
                } else {
                    done = 1;
                }
            }
        }
        if (final_keys) key_string(&s, final_keys + g * (int64_t)key_cap, (size_t)key_cap);
    }
    return steps;
}

/* single-state helpers for cross-checks against the Python oracle */
int yo_subset_position(int sub, int i) { init_subsets(); return SUBSET_POS[sub][i]; }
