/* yacht_b200.h -- C ABI of the B200-native Yacht-Auction engine (libyacht_b200.so).
 *
 * This is the drop-in boundary for the reference's hot path.  The reference
 * (iyioon/NYPC-Yacht-Auction) is pure Python and has no FFI of its own; each entry point
 * below replaces the Python method cited next to it and is what a maintainer would bind
 * with ctypes from yacht/YachtGame.py / MCTS.py (see INTEGRATION.md for the stub).
 *
 * Conventions
 *   - Plain pointers and sizes only.  All `states`, `players`, ... pointers are DEVICE
 *     pointers unless the function name contains `_host_`.  The caller owns every buffer.
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).  Calls are
 *     asynchronous, never allocate and never synchronise (except the `_host_` variants).
 *   - Return value: 0 on success, otherwise a cudaError_t value.
 *   - Packed state: 8 x uint32 per game in two uint4 planes: words 0..3 of game g live at
 *     states[4*g .. 4*g+3], words 4..7 at states[4*(stride+g) ..].  `stride` is the number
 *     of games the buffer was allocated for (>= n).  Bit layout: csrc/ya_common.cuh and
 *     DESIGN.md "Packed state".
 *   - players are int8 +1 / -1 (Game.py:40-52).  Actions are int32 in [0, 3226)
 *     (YachtGame.py:27-42): 202 bids then 12 x 252 (category, 5-of-10 subset) scores.
 *   - Per-game status codes (int32): 0 ok; 1 ValueError "Invalid action in BID phase"
 *     (YachtGame.py:268-269); 2 ValueError "Invalid action in SCORE phase" (:306-307);
 *     3 RuntimeError "Invalid phase/state" (:372); 4 AssertionError in bid resolution (:508);
 *     5 carry would exceed 10 dice (outside legal play; engine limit); 0x100 / 0x200: the
 *     transition needs a tie-break / two dice rolls that were not injected (draw_mode 1).
 */
#ifndef YACHT_B200_H
#define YACHT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define YA_ABI_VERSION 2
#define YA_ACTION_SIZE 3226
#define YA_FEATURE_SIZE 59
#define YA_SCORE_TABLE_SIZE 3024

int ya_abi_version(void);
int ya_set_device(int device);

/* getInitBoard (YachtGame.py:232-237) for n games: round 1, bid phase, rollA then rollB drawn
 * from Philox(seed; game_base+g, episode[g], ply 0, tag INIT).  players/ply/episode may be NULL. */
int ya_init_states(uint32_t* states, int64_t stride, int8_t* players, int32_t* ply, const uint32_t* episode,
                   int64_t n, uint64_t seed, uint64_t game_base, void* stream);

/* getNextState (YachtGame.py:260-372) incl. _resolve_bids_and_assign (:502-542).
 * draw_mode 0: dice/tie-breaks from Philox(seed; game_base+g, episode[g], ply[g], tag, depth_sim[g]).
 * draw_mode 1: draws injected by the host through the reference's own hooks roll_five /
 *   tiebreak_uniform (:154-159): injected[g*12] = {tie, rollA[5], rollB[5], valid bits (1 tie, 2 rolls)}.
 * in/out may alias.  status[g] as documented above; on error the state is copied unchanged. */
int ya_next_state(const uint32_t* states_in, int64_t in_stride, const int8_t* players, const int32_t* actions,
                  uint32_t* states_out, int64_t out_stride, int8_t* next_players, int32_t* status, int64_t n,
                  int draw_mode, const uint8_t* injected, uint64_t seed, uint64_t game_base,
                  const uint32_t* episode, const int32_t* ply, uint32_t tag, const uint32_t* depth_sim, void* stream);

/* getValidMoves (YachtGame.py:374-406): masks is uint8[n][3226], 16-byte aligned. */
int ya_valid_moves(const uint32_t* states, int64_t stride, const int8_t* players, uint8_t* masks, int64_t n, void* stream);

/* getGameEnded (YachtGame.py:408-428): 0, +1, -1 or 1e-4 per game, from players[g]'s view. */
int ya_game_ended(const uint32_t* states, int64_t stride, const int8_t* players, float* out, int64_t n, void* stream);

/* getCanonicalForm (YachtGame.py:430-442): p1<->p2 and their pending bids swapped where players[g] == -1. */
int ya_canonical_form(const uint32_t* states_in, int64_t in_stride, const int8_t* players,
                      uint32_t* states_out, int64_t out_stride, int64_t n, void* stream);

/* state_to_vec (yacht/NNet.py:50-86) of canonical states: features is float32[n][59]. */
int ya_features(const uint32_t* states, int64_t stride, float* features, int64_t n, void* stream);

/* RandomYachtPlayer.play (yacht/YachtPlayers.py:174-183): uniform over the legal set (0 if empty),
 * index = (Philox word * count) >> 32 with tag ACTION. */
int ya_random_action(const uint32_t* states, int64_t stride, const int8_t* players, int32_t* actions, int64_t n,
                     uint64_t seed, uint64_t game_base, const uint32_t* episode, const int32_t* ply, void* stream);

/* Scoring-move enumeration (the 12 x 252 loop of yacht/YachtPlayers.py:134-169 and
 * score_category, YachtGame.py:57-108): scores is uint8[n][12][252] = points / 1000 of every
 * (category, 5-of-10 subset) for the player to move, 0 where the subset does not fit. */
int ya_enumerate_scores(const uint32_t* states, int64_t stride, const int8_t* players, uint8_t* scores, int64_t n,
                        void* stream);

/* GreedyYachtPlayer.play (yacht/YachtPlayers.py:186-214) for the player to move of every game: greedy
 * best-immediate-gain scoring (:134-169) and the value-gap bid heuristic (:39-129, including the bid
 * overflow quirk).  raw (may be NULL) receives the heuristic's own choice; when that is not a legal
 * action the reference plays a random legal move: with fallback != 0 that move is drawn on device
 * (Philox, tag ACTION), otherwise actions[g] = -1 and the host draws it. */
int ya_greedy_action(const uint32_t* states, int64_t stride, const int8_t* players, int32_t* actions, int32_t* raw,
                     int64_t n, int fallback, uint64_t seed, uint64_t game_base, const uint32_t* episode,
                     const int32_t* ply, void* stream);

/* One ply of Arena.playGame (Arena.py:49-71) with RandomYachtPlayer on both sides, for n games at
 * once: legal mask written to masks (uint8[n][3226], may be NULL), action sampled, transition
 * applied in place, outcome[g] = getGameEnded(board, 1) after the move (0 while running).  With
 * auto_reset, a finished game is re-dealt (episode[g]++, ply 0).  err_flag (may be NULL) receives
 * OR(1 << status) of any non-zero status. */
int ya_play_ply(uint32_t* states, int64_t stride, int8_t* players, int32_t* ply, uint32_t* episode,
                int32_t* actions, float* outcome, uint8_t* masks, int32_t* err_flag,
                int64_t n, uint64_t seed, uint64_t game_base, int auto_reset, void* stream);

/* ---- MCTS (MCTS.py:16-164) over a flat per-game node pool ------------------------------------
 * The caller (torch) allocates the pool once and passes it as a ya_mcts_tree:
 *   nodes   uint32[n][max_nodes][16]   64-byte node records (key = packed canonical state)
 *   ht      uint16[n][ht_size]         open-addressing table, ht_size = power of two >= 2*max_nodes
 *   arena   uint32[n][arena_words]     legal-only prior rows (float32, or 16-bit logits; + per-32 group maxima) + visited-edge
 *                                      arrays (index, Nsa, Qsa); arena_words a multiple of 8 and the base 32-byte aligned
 *   meta    uint32[n][4]               node count, arena top, round of the last root
 *   cursor  uint32[n][32]              path / leaf of the simulation in flight
 * One simulation of every game = ya_mcts_select -> evaluator -> ya_mcts_expand.  The tree persists
 * across moves (Coach.py:93 resets it per episode: call ya_mcts_reset) and is pruned at round
 * boundaries, which is exact because the reference's search never leaves a round >= 2.          */
typedef struct {
    uint32_t* nodes;
    uint16_t* ht;
    uint32_t* arena;
    uint32_t* meta;
    uint32_t* cursor;
    int64_t n;
    int32_t max_nodes;
    int32_t ht_size;
    int64_t arena_words;
} ya_mcts_tree;

/* Row storage of a pool (fixed for the life of its trees; ya_mcts_reset in between): float32 prior rows (any float32
 * policy through ya_mcts_expand: the bit-exact reference arithmetic), or the leaf's legal policy-head logits kept as 16-bit
 * values + one exponent offset per node (ya_mcts_expand_logits; half the bytes, priors evaluated on the fly with the same
 * FMA + EX2 that would have produced the stored float32 value). */
#define YA_ROWS_F32 0
#define YA_ROWS_FP16 2
#define YA_ROWS_BF16 3

int ya_mcts_cursor_words(void);
int ya_mcts_node_words(void);

/* MCTS.__init__ (MCTS.py:16-26): empty tables for the games flagged in `which` (NULL = all). */
int ya_mcts_reset(const ya_mcts_tree* tree, const uint8_t* which, void* stream);

/* The descent of MCTS.search (MCTS.py:56-150) for simulation `sim` of every active game, from the
 * canonical form of states[g].  Leaves that need the evaluator get need_eval[g] = 1, their
 * state_to_vec row in features[g][59] and (optionally) their packed state in leaf_states; all other
 * paths (terminal, dead end) are backed up inside this call.  err_flag bits: 0x100 node pool full,
 * 0x200 arena full, 0x400 path deeper than 16, 0x800 | (1 << status) rule error.  If sim_ptr (device)
 * is not NULL the simulation index is read from it, and if game_base_ptr (device) is not NULL the global id of game 0
 * is read from it instead of `game_base`, so one captured CUDA graph can be replayed for every simulation and for
 * every wave of games that shares the pool (a captured launch freezes its by-value arguments).
 * rows = YA_ROWS_*: how this pool stores priors.  leaf_dst / leaf_desc (device, uint64[n] / uint32[n], both or neither;
 * logit-row pools): for every leaf that needs the evaluator, the device address of its row's logit area and its legal-mask
 * descriptor (0 / 0 otherwise) -- the scatter targets ya_nn_forward's policy-head epilogue writes the legal logits to.
 * cpuct must be >= 0 (main.py:25 uses 1.5): the unvisited arg-max goes through per-group maxima of the priors,
 * which needs cpuct * P * sqrt(Ns + EPS) to be monotone in P; a negative value is rejected (invalid value). */
int ya_mcts_select(const ya_mcts_tree* tree, const uint32_t* states, int64_t stride, const int8_t* players,
                   const int32_t* ply, const uint32_t* episode, uint64_t seed, uint64_t game_base, uint32_t sim,
                   const uint32_t* sim_ptr, const uint64_t* game_base_ptr, float cpuct, const uint8_t* active,
                   float* features, uint8_t* need_eval, uint32_t* leaf_states, int rows, uint64_t* leaf_dst,
                   uint32_t* leaf_desc, int32_t* err_flag, void* stream);

/* The same descent with the dice of in-search transitions supplied by the HOST (the reference rolls them
 * from the global numpy / random streams through roll_five / tiebreak_uniform, yacht/YachtGame.py:154-159,
 * so a drop-in MCTS must consume those streams at exactly the same points).  injected[g*12] as in
 * ya_next_state.  When the chosen transition needs draws that are not marked valid the descent is parked
 * and need_eval[g] = 0x10 | (1: tie-break needed) | (2: two dice rolls needed); call again with resume = 1
 * and the draws filled in.  Injected draws are consumed by one transition.  need_eval[g] = 1 / 0 as above. */
int ya_mcts_select_injected(const ya_mcts_tree* tree, const uint32_t* states, int64_t stride, const int8_t* players,
                            uint32_t sim, float cpuct, const uint8_t* injected, int resume, float* features,
                            uint8_t* need_eval, uint32_t* leaf_states, int32_t* err_flag, void* stream);

/* Leaf expansion + backup (MCTS.py:86-115,152-164): pi is float32[n][3226] over ALL actions and
 * value float32[n], as NeuralNet.predict returns them (NeuralNet.py:27-37); with uniform != 0 every
 * leaf gets pi = uniform_p, v = uniform_v without reading memory (BASELINE.json configs[2]).
 * sim_counter (device, may be NULL) is incremented once per call. */
int ya_mcts_expand(const ya_mcts_tree* tree, const float* pi, const float* value, int uniform, float uniform_p,
                   float uniform_v, uint32_t* sim_counter, int32_t* err_flag, void* stream);

/* Leaf expansion + backup for logit-row pools (ya_mcts_select with rows = YA_ROWS_FP16 / YA_ROWS_BF16, fp16 saying which).
 * The softmax of NNetWrapper.predict (yacht/NNet.py:193), masking and renormalisation (MCTS.py:88-101) are fused:
 * P[a] = exp(l[a] - max) / sum over legal a' of exp(l[a'] - max) = 2^(l[a] * log2(e) + off), max taken over all 3,226
 * logits, off stored per node; neither float32 logits nor pi nor float32 priors are ever written to memory.
 *   logits16 == NULL: the rows were filled by ya_nn_forward's epilogue (scatter targets from ya_mcts_select); row_max
 *                     (written by the same forward) is required.
 *   logits16 != NULL: a dense 16-bit logit matrix [n][ld] from any evaluator (ld >= 3232 and a multiple of 8, e.g. the
 *                     head padded to 3232 columns; base 16-byte aligned); the legal entries are compacted into the rows
 *                     here.  row_max float32[n] = per-row maximum over the 3,226 columns, or NULL: the kernel scans the row. */
int ya_mcts_expand_logits(const ya_mcts_tree* tree, const void* logits16, int fp16, int64_t ld, const float* row_max,
                          const float* value, uint32_t* sim_counter, int32_t* err_flag, void* stream);

/* getActionProb's whole simulation loop (MCTS.py:37-38) in ONE launch for the uniform evaluator
 * (pi = uniform_p, v = uniform_v; BASELINE.json configs[2]): num_sims x (descent, expansion, backup) per
 * game without leaving the SM.  Same results as num_sims x (ya_mcts_select, ya_mcts_expand(uniform)).
 * With one prior value for every legal move of a node, the nodes this call creates keep that value and a visited
 * bitmask instead of a float32 prior row (0.4 KB instead of up to 12 KB): drive a tree either with this call or
 * with ya_mcts_select / ya_mcts_expand*, not both (ya_mcts_reset in between); the root_* readers accept both. */
int ya_mcts_search_uniform(const ya_mcts_tree* tree, const uint32_t* states, int64_t stride, const int8_t* players,
                           const int32_t* ply, const uint32_t* episode, uint64_t seed, uint64_t game_base, int num_sims,
                           float cpuct, float uniform_p, float uniform_v, const uint8_t* active, int32_t* err_flag,
                           void* stream);

/* counts[g][a] = Nsa[(root, a)] (MCTS.py:40-42), visits[g] = Ns[root] (-1 if the root is unknown);
 * optional qvals (float64) / qkind (1 = numpy float32, 2 = Python float) expose Qsa for tests. */
int ya_mcts_root_counts(const ya_mcts_tree* tree, const uint32_t* states, int64_t stride, const int8_t* players,
                        int32_t* counts, int32_t* visits, double* qvals, uint8_t* qkind, void* stream);

/* The same statistics in canonical sparse form for training examples (Coach.py:60-63): the visited root
 * edges as (action int16, Nsa int32) pairs in ascending action order, zero padded to k per game;
 * overflow[g] (may be NULL) = number of visited edges if they exceed k, -1 if the root is unknown, else 0. */
int ya_mcts_root_sparse(const ya_mcts_tree* tree, const uint32_t* states, int64_t stride, const int8_t* players,
                        int k, int16_t* actions, int32_t* counts, int32_t* overflow, void* stream);

/* Move selection from visit counts (Coach.py:56-65, MCTS.py:44-54): temp = (ply+1 < temp_threshold);
 * temp 1 samples proportionally to the counts, temp 0 picks uniformly among the arg-max actions;
 * randomness = Philox word of tag ACTION. */
int ya_mcts_pick_action(const int32_t* counts, const int32_t* ply, const uint32_t* episode, int64_t n, uint64_t seed,
                        uint64_t game_base, int temp_threshold, int32_t* actions, void* stream);

/* ---- leaf evaluator (YachtNNet.forward, yacht/pytorch/YachtNNet.py:62-70) ---- */
/* The whole YachtNNet.forward (yacht/pytorch/YachtNNet.py:62-70) for a wave of leaves as one persistent
 * tcgen05 kernel: features float32 [n][59] (state_to_vec rows; contiguous -- a 16-byte aligned matrix lets whole 128-row tiles
 * arrive by bulk copy, any other alignment works too) -> 16-bit logits [n][3232] (columns >= 3226 are padding: the dense
 * matrix holds -inf there) and values float32 [n] (tanh).  fp16 != 0: operands (activations, weight images) and logits are IEEE half,
 * the precision of the reference's CUDA predict (fp16 autocast, yacht/NNet.py:186-193); fp16 == 0: bfloat16.
 * Accumulation, bias, SiLU, LayerNorm and the residual sum are float32 in both modes; the residual stream is stored
 * between blocks as IEEE half (with half operands: exactly the next block's operand).  weight_blob / param_blob are built once on the host
 * (mcts.FusedYachtEvaluator): 128-byte-swizzled K-major images of W_in (K padded to 64), the 2*nblocks trunk
 * weights, the value head's first Linear and 26 policy-head tiles of 128 columns, every matrix split by output columns
 * into the halves the two CTAs of a pair hold (layout: mcts.FusedYachtEvaluator.pair_image; tests/test_forward_host.py); biases and LayerNorm
 * parameters as float32.  offsets (HOST pointer, int64[9]) = byte offsets {w_in, w_trunk, w_v, w_pi} and float
 * offsets {p_in, p_trunk, p_v, p_pi_ln, p_pi_bias}.  hidden width 256 only.  Rows are independent of the
 * batch they sit in (batch-invariant evaluator).  row_max float32[n] (may be NULL) receives each row's largest
 * logit over the 3,226 real columns, for ya_mcts_expand_logits.
 * scatter_dst / scatter_desc (device, uint64[n] / uint32[n], both or neither; written by ya_mcts_select): the policy-head
 * epilogue writes row g's LEGAL logits (legal-mask descriptor scatter_desc[g]), in legal order, to the device address
 * scatter_dst[g] -- the leaf's row in the tree pool -- and skips rows with address 0 (MCTS.py:86-101: the mask is applied
 * where the logits are produced).  logits16 may then be NULL: no dense logit matrix is written at all. */
int ya_nn_forward(const float* features, void* logits16, float* values, float* row_max, const void* weight_blob,
                  const float* param_blob, const int64_t* offsets, int nblocks, int64_t n, float eps, int fp16,
                  const uint64_t* scatter_dst, const uint32_t* scatter_desc, void* stream);
/* The same with the scheduling chosen by the caller.  tiles_per_cta = 1: one 128-leaf tile per CTA (a CTA pair = 256
 * leaves); 2: two tiles per CTA (512 leaves per pair): the tensor core works on one tile while the 16 warps run the other
 * tile's epilogue -- for waves of more than one tile per SM; 0: by size (ya_nn_forward).  Both schedules compute every
 * row with the same arithmetic in the same order: outputs are bit-identical. */
int ya_nn_forward_tiles(const float* features, void* logits16, float* values, float* row_max, const void* weight_blob,
                        const float* param_blob, const int64_t* offsets, int nblocks, int64_t n, float eps, int fp16,
                        const uint64_t* scatter_dst, const uint32_t* scatter_desc, int tiles_per_cta, void* stream);

/* ---- host-buffer variants (end-to-end path for callers that keep boards in host memory) ----
 * ya_host_create allocates the device mirror for n games once (no allocation per call);
 * ya_host_play_ply copies the packed states and side arrays host->device, runs ya_play_ply,
 * copies states / players / ply / episode / actions / outcome (and masks if requested and the
 * context was created with_masks) back and synchronises.  All pointers are HOST pointers; the
 * state buffer uses stride = n.
 *
 * Record variant: one 64-byte record per game, uint32[16] = {w0..w7 packed state, episode, ply,
 * player (+1/-1 as int32), action of the last ply (out), outcome of the last ply (out, float bits),
 * games finished / won by player 1 / won by player 2 (running tallies, in-out)}, so a slice of games is
 * ONE contiguous copy per direction; ya_host_play_ply_records pipelines 4 slices over 4 streams
 * (H2D, kernel, D2H overlap on the full-duplex link).  ya_host_play_plies_records plays `plies` plies
 * per call (e.g. 48 = Arena.playGames for every slot: one full game each, finished games re-dealt and
 * tallied) with a single round trip.  ya_play_ply_records is the device-side kernel
 * entry on records already in HBM.  With a context created with_masks the uint8[n][3226] mask is
 * materialised in HBM every ply (and copied to `masks` only if that pointer is not NULL). */
int ya_host_create(int64_t n, int with_masks, void** handle);
int ya_host_destroy(void* handle);
int ya_host_play_ply(void* handle, uint32_t* states, int8_t* players, int32_t* ply, uint32_t* episode,
                     int32_t* actions, float* outcome, uint8_t* masks, int32_t* err_flag,
                     uint64_t seed, uint64_t game_base, int auto_reset);
int ya_host_play_ply_records(void* handle, uint32_t* records, uint8_t* masks, int32_t* err_flag,
                             uint64_t seed, uint64_t game_base, int auto_reset);
int ya_host_play_plies_records(void* handle, uint32_t* records, int plies, uint8_t* masks, int32_t* err_flag,
                               uint64_t seed, uint64_t game_base, int auto_reset);
int ya_play_ply_records(uint32_t* records, uint8_t* masks, int32_t* err_flag, int64_t n, uint64_t seed,
                        uint64_t game_base, int auto_reset, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* YACHT_B200_H */
