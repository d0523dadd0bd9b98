// Yacht-Auction B200 engine -- environment kernels (init / transition / legal masks /
// random-legal policy / outcomes / canonical form / features / scoring-move enumeration and
// the fused "play one ply" kernel) and their C-ABI entry points (include/yacht_b200.h).
//
// Every entry point takes raw device pointers (torch owns the buffers), launches on the
// stream it is given and returns a cudaError_t as int.  No allocation, no host sync.
#include "ya_common.cuh"
#include "../../include/yacht_b200.h"
#include <atomic>
#include <mutex>

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ YaState ya_fresh_state(uint64_t seed, uint32_t gid, uint32_t episode) {
    // getInitBoard: YachtGame.py:232-237 (rollA then rollB)
    YaDraw d = ya_draw(seed, gid, episode, 0, YA_TAG_INIT, 0, 0);
    YaState s;
    s.w[0] = 1u;
    s.w[1] = d.roll_a | (d.roll_b << 15);
#pragma unroll
    for (int i = 2; i < 8; ++i) s.w[i] = 0u;
    return s;
}

// ------------------------------------------------------------------ init
__global__ void ya_k_init(uint4* __restrict__ states, int64_t stride, int8_t* __restrict__ players,
                          int32_t* __restrict__ ply, const uint32_t* __restrict__ episode,
                          int64_t n, uint64_t seed, uint64_t game_base) {
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    uint32_t ep = episode ? episode[g] : 0u;
    YaState s = ya_fresh_state(seed, (uint32_t)(game_base + g), ep);
    ya_store(states, stride, g, s);
    if (players) players[g] = 1;
    if (ply) ply[g] = 0;
}

// ------------------------------------------------------------------ getNextState
// draw_mode 0: Philox on device, counter (game, episode, ply, tag, depth, sim).
// draw_mode 1: draws injected by the host: inj[g*12 + 0] = tie, [1..5] rollA, [6..10] rollB,
//              [11] = bitset of what is valid (1 = tie, 2 = rolls).  If a needed draw is not
//              valid the state is left untouched and status = YA_NEED_* tells the host what to draw.
__global__ void ya_k_next_state(const uint4* __restrict__ in, int64_t in_stride,
                                const int8_t* __restrict__ players, const int32_t* __restrict__ actions,
                                uint4* __restrict__ out, int64_t out_stride, int8_t* __restrict__ next_players,
                                int32_t* __restrict__ status, int64_t n,
                                int draw_mode, const uint8_t* __restrict__ inj,
                                uint64_t seed, uint64_t game_base, const uint32_t* __restrict__ episode,
                                const int32_t* __restrict__ ply, uint32_t tag, const uint32_t* __restrict__ depth_sim) {
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    YaState s = ya_load(in, in_stride, g);
    int pl = players[g];
    int a = actions[g];
    int needs = ya_draw_needs(s, pl, a);
    YaDraw d;
    d.roll_a = d.roll_b = d.tie = d.pick = 0;
    int st = YA_OK;
    int np = pl;
    bool run = true;
    if (needs) {
        if (draw_mode == 0) {
            uint32_t ds = depth_sim ? depth_sim[g] : 0u;
            d = ya_draw(seed, (uint32_t)(game_base + g), episode ? episode[g] : 0u, ply ? (uint32_t)ply[g] : 0u,
                        tag, ds & 0xFF, ds >> 8);
        } else {
            const uint8_t* t = inj + g * 12;
            int have = t[11];
            int missing = 0;
            if ((needs & YA_NEED_TIE) && !(have & 1)) missing |= YA_NEED_TIE;
            if ((needs & YA_NEED_ROLLS) && !(have & 2)) missing |= YA_NEED_ROLLS;
            if (missing) { st = missing; run = false; }
            d.tie = t[0];
            for (int i = 0; i < 5; ++i) {
                d.roll_a |= (uint32_t)t[1 + i] << (3 * i);
                d.roll_b |= (uint32_t)t[6 + i] << (3 * i);
            }
        }
    }
    if (run) np = ya_transition(s, pl, a, d, &st);
    ya_store(out, out_stride, g, s);
    next_players[g] = (int8_t)np;
    status[g] = st;
}

// ------------------------------------------------------------------ legal-mask writer
// A game's mask row (uint8[3226], the reference's dtype: YachtGame.py:375) is 13 runs: 202
// bid bytes then 12 x 252 category bytes.  Per game we keep F (13 fill bits): run r is F[r]
// everywhere.  Bid rows and ten-dice rows are exactly that; five-dice rows (round 13 only) have
// F = 0 for the category runs and get their <= 12 single "subset 0" bytes patched afterwards.
//
// Rows are 3226 B apart (only 2-byte aligned), so the buffer is written as one flat byte stream
// with 16-byte vector stores.  8 rows = 25,808 B = 1613 vectors is the alignment period: which
// run(s) a vector covers and where the run boundary falls inside it depends only on the vector's
// index j in its 8-row superblock, never on game data.  That static part lives in a shared-memory
// table (built once per CTA); the per-vector work is: 1 table load, 1 descriptor load, 1 LUT load,
// a handful of logic ops and one 16-byte streaming store.
constexpr int kSuperGames = 8;
constexpr int kSuperVec = kSuperGames * YA_N_ACTION / 16;          // 1613

struct YaMaskSmem {
    uint4 tail_lut[16];              // [k-1]: 0xFF in bytes >= k (k = bytes of the first run in the vector)
    uint32_t fext[72];               // per local game: F[0..12] | F_next_game[0] << 13
    uint16_t tab[kSuperVec + 3];     // per vector of a superblock: local game | run << 3 | (k-1) << 7
};

__device__ __forceinline__ uint32_t ya_fill_bits(uint32_t desc) {
    uint32_t h = desc & 0x1FFFu;
    return (desc >> 13) ? h : (h & 1u);
}

__device__ __forceinline__ void ya_mask_smem_init(YaMaskSmem& sm, int tid, int nthr) {
    for (int j = tid; j < kSuperVec; j += nthr) {
        int b = j * 16;
        int lg = b / YA_N_ACTION, p = b - lg * YA_N_ACTION;
        int s = ((p + 50) * 4162) >> 20;            // run index: 0 = bids, 1..12 = categories
        int k = min(202 + 252 * s - p, 16);         // bytes of run s inside this vector
        sm.tab[j] = (uint16_t)(lg | (s << 3) | ((k - 1) << 7));
    }
    if (tid < 16) {
        uint32_t w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint32_t m = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) if (4 * q + i >= tid + 1) m |= 0xFFu << (8 * i);
            w[q] = m;
        }
        sm.tail_lut[tid] = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// fill[] (shared, ng entries valid) -> fext[]; call between two __syncthreads().
__device__ __forceinline__ void ya_mask_smem_link(YaMaskSmem& sm, const uint32_t* fill, int ng, int cap, int tid, int nthr) {
    for (int t = tid; t < cap; t += nthr) {
        uint32_t f = t < ng ? fill[t] : 0u;
        uint32_t nx = (t + 1 < ng) ? fill[t + 1] : 0u;
        sm.fext[t] = f | ((nx & 1u) << 13);
    }
}

// Writes rows of `ng` consecutive games starting at a 16-byte aligned address (ng <= 64).
__device__ __forceinline__ void ya_write_mask_chunk(uint8_t* __restrict__ out, int ng, const YaMaskSmem& sm,
                                                    int tid, int nthr) {
    const int total = ng * YA_N_ACTION;
    const int nvec = total >> 4;
    uint4* vout = reinterpret_cast<uint4*>(out);
    int j = tid, sb8 = 0;                           // nthr <= kSuperVec
    for (int v = tid; v < nvec; v += nthr) {
        uint32_t t = sm.tab[j];
        uint32_t two = sm.fext[sb8 + (t & 7u)] >> ((t >> 3) & 15u);
        uint32_t A = (two & 1u) * 0x01010101u;
        uint32_t X = ((two ^ (two >> 1)) & 1u) * 0x01010101u;
        uint4 nm = sm.tail_lut[t >> 7];
        __stcs(vout + v, make_uint4(A ^ (X & nm.x), A ^ (X & nm.y), A ^ (X & nm.z), A ^ (X & nm.w)));
        j += nthr;
        if (j >= kSuperVec) { j -= kSuperVec; sb8 += kSuperGames; }
    }
    // ragged tail (only when ng is not a multiple of 8): plain byte stores
    for (int i = (nvec << 4) + tid; i < total; i += nthr) {
        int gg = i / YA_N_ACTION, pp = i - gg * YA_N_ACTION;
        int s = ((pp + 50) * 4162) >> 20;
        out[i] = (uint8_t)((sm.fext[gg] >> s) & 1u);
    }
}

// Five-dice rows: set the "subset 0" byte of every open category (call after a __syncthreads()
// that follows ya_write_mask_chunk).  desc[] holds the ya_mask_desc words of the chunk.
__device__ __forceinline__ void ya_patch_five_dice_rows(uint8_t* __restrict__ out, int ng, const uint32_t* desc,
                                                        int tid, int nthr) {
    for (int t = tid; t < ng; t += nthr) {
        uint32_t d = desc[t];
        if ((d >> 13) || !(d & 0x1FFEu)) continue;
        uint8_t* row = out + (size_t)t * YA_N_ACTION + YA_N_BID;
        for (uint32_t open = (d >> 1) & 0xFFFu; open; open &= open - 1)
            row[(__ffs(open) - 1) * YA_N_SUBSET] = 1;
    }
}

constexpr int kMaskGames = 64;      // games per CTA in the mask kernels (multiple of 8)

__global__ void __launch_bounds__(kThreads)
ya_k_valid_moves(const uint4* __restrict__ states, int64_t stride, const int8_t* __restrict__ players,
                 uint8_t* __restrict__ masks, int64_t n) {
    __shared__ YaMaskSmem sm;
    __shared__ uint32_t desc[kMaskGames];
    __shared__ uint32_t fill[kMaskGames];
    int64_t g0 = (int64_t)blockIdx.x * kMaskGames;
    int ng = (int)min((int64_t)kMaskGames, n - g0);
    ya_mask_smem_init(sm, threadIdx.x, blockDim.x);
    if (threadIdx.x < ng) {
        YaState s = ya_load(states, stride, g0 + threadIdx.x);
        uint32_t d = ya_mask_desc(s, players[g0 + threadIdx.x]);
        desc[threadIdx.x] = d;
        fill[threadIdx.x] = ya_fill_bits(d);
    }
    __syncthreads();
    ya_mask_smem_link(sm, fill, ng, kMaskGames, threadIdx.x, blockDim.x);
    __syncthreads();
    uint8_t* out = masks + g0 * YA_N_ACTION;
    ya_write_mask_chunk(out, ng, sm, threadIdx.x, blockDim.x);
    __syncthreads();
    ya_patch_five_dice_rows(out, ng, desc, threadIdx.x, blockDim.x);
}

// ------------------------------------------------------------------ small per-game kernels
__global__ void ya_k_game_ended(const uint4* __restrict__ states, int64_t stride, const int8_t* __restrict__ players,
                                float* __restrict__ out, int64_t n) {
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    out[g] = ya_game_ended(ya_load(states, stride, g), players[g]);
}

__global__ void ya_k_canonical(const uint4* __restrict__ in, int64_t in_stride, const int8_t* __restrict__ players,
                               uint4* __restrict__ out, int64_t out_stride, int64_t n) {
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    YaState s = ya_load(in, in_stride, g);
    ya_store(out, out_stride, g, players[g] == 1 ? s : ya_flip(s));
}

__global__ void ya_k_features(const uint4* __restrict__ states, int64_t stride, float* __restrict__ feat, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * YA_N_FEATURE) return;
    int64_t g = i / YA_N_FEATURE;
    int f = (int)(i - g * YA_N_FEATURE);
    feat[i] = ya_feature(ya_load(states, stride, g), f);
}

__global__ void ya_k_random_action(const uint4* __restrict__ states, int64_t stride, const int8_t* __restrict__ players,
                                   int32_t* __restrict__ actions, int64_t n, uint64_t seed, uint64_t game_base,
                                   const uint32_t* __restrict__ episode, const int32_t* __restrict__ ply) {
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    YaState s = ya_load(states, stride, g);
    uint32_t desc = ya_mask_desc(s, players[g]);
    int count = ya_legal_count(desc);
    int a = 0;                                             // YachtPlayers.py:183: 0 when nothing is legal
    if (count) {
        YaDraw d = ya_draw(seed, (uint32_t)(game_base + g), episode ? episode[g] : 0u, ply ? (uint32_t)ply[g] : 0u,
                           YA_TAG_ACTION, 0, 0);
        a = ya_nth_legal(desc, (int)__umulhi(d.pick, (uint32_t)count));
    }
    actions[g] = a;
}

// ------------------------------------------------------------------ score table
// score_category depends only on the multiset of the five dice: 252 multisets.  A multiplicative
// perfect hash of the 24-bit face histogram ((hist * 0xB923B81B) >> 22, collision-free on the 252
// reachable histograms, found offline) indexes a 1024-slot table holding the 12 category scores / 1000
// as three packed words {cats 0-3, 4-7, 8-11}.  Built once per process and device by a tiny kernel.
constexpr uint32_t kHistHashMul = 0xB923B81Bu;
__device__ uint32_t g_score_table[3][1024];

__global__ void ya_k_build_score_table() {
    int t = blockIdx.x * blockDim.x + threadIdx.x;          // one ordered 5-tuple per thread (7776)
    if (t >= 7776) return;
    uint32_t hist = 0, pips = 0;
#pragma unroll
    for (int i = 0; i < 5; ++i) { uint32_t d = 1 + t % 6; t /= 6; hist += 1u << (4 * (d - 1)); pips += d; }
    uint32_t w[3] = {0, 0, 0};
#pragma unroll
    for (int c = 0; c < YA_N_CAT; ++c) w[c >> 2] |= ya_category_points_k(c, hist, pips) << (8 * (c & 3));
    uint32_t slot = (hist * kHistHashMul) >> 22;
    g_score_table[0][slot] = w[0]; g_score_table[1][slot] = w[1]; g_score_table[2][slot] = w[2];
}

// The table is per device (a __device__ array).  Thread-safe and stream-safe: the first caller on a device builds it
// under a mutex and waits for the build, so every later launch -- on any stream, from any thread -- finds it complete.
// ya_set_device builds it eagerly, which keeps the wait out of CUDA-graph captures.
int ya_ensure_score_table(cudaStream_t stream) {
    static std::atomic<bool> built[64];
    static std::mutex mu;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    if (dev < 0 || dev >= 64) return (int)cudaErrorInvalidDevice;
    if (built[dev].load(std::memory_order_acquire)) return 0;
    std::lock_guard<std::mutex> lock(mu);
    if (built[dev].load(std::memory_order_relaxed)) return 0;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &cap) != cudaSuccess) return (int)cudaGetLastError();
    if (cap != cudaStreamCaptureStatusNone) return (int)cudaErrorStreamCaptureUnsupported;   // call ya_set_device first
    ya_k_build_score_table<<<(7776 + 255) / 256, 256, 0, stream>>>();
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) return (int)e;
    built[dev].store(true, std::memory_order_release);
    return 0;
}

struct YaScoreSmem {
    uint32_t tab[3][1024];
    uint16_t smask[YA_N_SUBSET];
};

__device__ __forceinline__ void ya_score_smem_init(YaScoreSmem& sm, int tid, int nthr) {
    for (int i = tid; i < 3 * 1024; i += nthr) (&sm.tab[0][0])[i] = (&g_score_table[0][0])[i];
    for (int i = tid; i < YA_N_SUBSET; i += nthr) sm.smask[i] = ya_subset_mask[i];
}

// partial (histogram | pips << 24) sums over the dice of every 5-bit pattern: lane l serves pattern l of the
// low half (dice 0..4) and of the high half (dice 5..9); a subset's word is part[m & 31] + part[32 + (m >> 5)]
__device__ __forceinline__ void ya_build_parts(uint32_t dice10, uint32_t* part, int lane) {
    uint32_t lo = 0, hi = 0;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        uint32_t dl = (dice10 >> (3 * j)) & 7u, dh = (dice10 >> (3 * (j + 5))) & 7u;
        uint32_t wl = dl ? ((1u << (4 * (dl - 1))) | (dl << 24)) : 0u;
        uint32_t wh = dh ? ((1u << (4 * (dh - 1))) | (dh << 24)) : 0u;
        if ((lane >> j) & 1) { lo += wl; hi += wh; }
    }
    __syncwarp();
    part[lane] = lo;
    part[32 + lane] = hi;
    __syncwarp();
}

// ------------------------------------------------------------------ scoring-move enumeration
// out[g][cat][subset] = score_category(cat, dice at subset) / 1000 for the player to move, 0
// where the subset does not fit (YachtPlayers.py:134-169 enumerates exactly this 12 x 252 table).
// One warp per game.  A subset's face histogram (six 4-bit counters) and pip sum (bits 24..28) are
// additive over its dice, so they come from two 32-entry tables per game -- partial sums over dice
// 0..4 and 5..9, built by the 32 lanes in one step -- as T_lo[m & 31] + T_hi[m >> 5]: two shared-
// memory loads and an add per subset instead of a 10-step gather.  Each lane scores four consecutive
// subsets and writes one packed 32-bit word per category into the 3,024-byte tile, which leaves as
// 189 coalesced 16-byte stores.  Measured alternatives (65,536 games, this kernel: 47 us = 0.65 of the HBM copy peak):
// storing the words straight from registers (4-byte stores, 128 bytes per warp instruction, no tile): 60 us -- the
// unaligned partial sectors cost more than the tile's shared-memory round trip; one 16-byte table entry per subset and
// one 8-byte load for four dice masks (fewer, wider shared-memory loads): 49 us.
constexpr int kEnumWarps = 8;

// 4x4 byte transpose: in[k] = {b0,b1,b2,b3} of subset k  ->  out[b] = {in[0].b, in[1].b, in[2].b, in[3].b}
__device__ __forceinline__ void ya_transpose4(const uint32_t in[4], uint32_t out[4]) {
    uint32_t t0 = __byte_perm(in[0], in[1], 0x5140), t1 = __byte_perm(in[0], in[1], 0x7362);
    uint32_t t2 = __byte_perm(in[2], in[3], 0x5140), t3 = __byte_perm(in[2], in[3], 0x7362);
    out[0] = __byte_perm(t0, t2, 0x5410); out[1] = __byte_perm(t0, t2, 0x7632);
    out[2] = __byte_perm(t1, t3, 0x5410); out[3] = __byte_perm(t1, t3, 0x7632);
}

__global__ void __launch_bounds__(kEnumWarps * 32)
ya_k_enumerate_scores(const uint4* __restrict__ states, int64_t stride, const int8_t* __restrict__ players,
                      uint8_t* __restrict__ out, int64_t n) {
    __shared__ __align__(16) uint32_t tile[kEnumWarps][YA_N_CAT * YA_N_SUBSET / 4];
    __shared__ uint32_t part[kEnumWarps][64];
    __shared__ YaScoreSmem sm;
    ya_score_smem_init(sm, threadIdx.x, blockDim.x);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t nwarps = (int64_t)gridDim.x * kEnumWarps;
    for (int64_t g = (int64_t)blockIdx.x * kEnumWarps + warp; g < n; g += nwarps) {
        YaState s = ya_load(states, stride, g);
        const uint32_t carry = s.w[2 + (players[g] == 1 ? 0 : 1)];
        const int nd = ya_dice_count(carry);
        ya_build_parts(carry, part[warp], lane);
        uint32_t* t = tile[warp];
        for (int q = lane; q < YA_N_SUBSET / 4; q += 32) {          // subsets 4q .. 4q+3
            uint32_t w[3][4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t m = sm.smask[4 * q + k];
                const bool fits = (31 - __clz(m)) < nd;
                const uint32_t hist = (part[warp][m & 31u] + part[warp][32 + (m >> 5)]) & 0xFFFFFFu;
                const uint32_t slot = (hist * kHistHashMul) >> 22;
#pragma unroll
                for (int j = 0; j < 3; ++j) w[j][k] = fits ? sm.tab[j][slot] : 0u;
            }
#pragma unroll
            for (int j = 0; j < 3; ++j) {                            // categories 4j .. 4j+3
                uint32_t o[4];
                ya_transpose4(w[j], o);
#pragma unroll
                for (int b = 0; b < 4; ++b) t[(4 * j + b) * (YA_N_SUBSET / 4) + q] = o[b];
            }
        }
        __syncwarp();
        const uint4* src = reinterpret_cast<const uint4*>(t);
        uint4* dst = reinterpret_cast<uint4*>(out + g * (YA_N_CAT * YA_N_SUBSET));
        for (int i = lane; i < (YA_N_CAT * YA_N_SUBSET) / 16; i += 32) __stcs(dst + i, src[i]);
        __syncwarp();
    }
}

// ------------------------------------------------------------------ greedy heuristic player
// GreedyYachtPlayer (yacht/YachtPlayers.py:39-214): the scoring move maximising the immediate gain
// (score + 35000 when it crosses the upper-section bonus, first maximum wins) and the value-gap bid
// heuristic, both on top of the same table-driven subset enumeration.  One warp per game.
__device__ __forceinline__ void ya_greedy_best(uint32_t dice10, int nd, uint32_t used, uint32_t upper_before,
                                               uint32_t* part, const YaScoreSmem& sm, int lane, int& best, int& best_idx) {
    ya_build_parts(dice10, part, lane);
    best = -1; best_idx = 0x7FFFFFFF;
    for (int sub = lane; sub < YA_N_SUBSET; sub += 32) {
        const uint32_t m = sm.smask[sub];
        if ((31 - __clz(m)) >= nd) continue;
        const uint32_t hist = (part[m & 31u] + part[32 + (m >> 5)]) & 0xFFFFFFu;
        const uint32_t slot = (hist * kHistHashMul) >> 22;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const uint32_t w = sm.tab[j][slot];
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int c = 4 * j + b;
                if ((used >> c) & 1u) continue;
                const uint32_t sc = (w >> (8 * b)) & 0xFFu;
                int g = (int)sc;
                if (c < 6 && upper_before < 63u && upper_before + sc >= 63u) g += 35;   // YachtPlayers.py:158-162
                int idx = c * YA_N_SUBSET + sub;
                if (g > best || (g == best && idx < best_idx)) { best = g; best_idx = idx; }
            }
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        int ob = __shfl_xor_sync(0xFFFFFFFFu, best, o), oi = __shfl_xor_sync(0xFFFFFFFFu, best_idx, o);
        if (ob > best || (ob == best && oi < best_idx)) { best = ob; best_idx = oi; }
    }
}

__device__ __forceinline__ uint32_t ya_upper_k(uint32_t w5) {
    uint32_t u = 0;
#pragma unroll
    for (int c = 0; c < 6; ++c) u += (uint32_t)(c + 1) * ((w5 >> (3 * c)) & 7u);
    return u;
}

__global__ void __launch_bounds__(kEnumWarps * 32)
ya_k_greedy_action(const uint4* __restrict__ states, int64_t stride, const int8_t* __restrict__ players,
                   int32_t* __restrict__ actions, int32_t* __restrict__ raw, int64_t n, int fallback,
                   uint64_t seed, uint64_t game_base, const uint32_t* __restrict__ episode, const int32_t* __restrict__ ply) {
    __shared__ uint32_t part[kEnumWarps][64];
    __shared__ YaScoreSmem sm;
    ya_score_smem_init(sm, threadIdx.x, blockDim.x);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t nwarps = (int64_t)gridDim.x * kEnumWarps;
    for (int64_t g = (int64_t)blockIdx.x * kEnumWarps + warp; g < n; g += nwarps) {
        YaState s = ya_load(states, stride, g);
        const int pl = players[g];
        if (pl != 1) s = ya_flip(s);                              // the players see canonical boards (Arena.py:55-56)
        const uint32_t used = s.w[4] & 0xFFFu, upper = ya_upper_k(s.w[5]);
        const uint32_t desc = ya_mask_desc(s, 1);
        int a;
        if (ya_bidding(s)) {                                       // _choose_bid, YachtPlayers.py:98-129
            int val[2];
            const uint32_t carry = s.w[2];
            const int nc = ya_dice_count(carry);
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                const uint32_t bundle = (s.w[1] >> (15 * b)) & 0x7FFFu;
                if (ya_round(s) == 1) {                            // forward heuristic, :47-64
                    uint32_t dice = carry | (bundle << (3 * nc));
                    uint32_t hist, pips;
                    ya_gather(dice, (1u << (nc + 5)) - 1u, hist, pips);
                    int v = (int)pips;
                    if ((hist + 0x444444u) & 0x888888u) v += 6;
                    else if (ya_nibble_eq(hist, 3)) v += 3;
                    if (ya_category_points_k(9, hist, pips)) v += 5;
                    val[b] = v;
                } else {
                    int best, idx;
                    const int nd = nc + 5;
                    if (nd > 10) { val[b] = 0; continue; }
                    ya_greedy_best(carry | (bundle << (3 * nc)), nd, used, upper, part[warp], sm, lane, best, idx);
                    val[b] = best < 0 ? 0 : best;                   // :95 (nothing playable -> 0)
                }
            }
            const int target = val[0] >= val[1] ? 0 : 1;
            const int gap = abs(val[0] - val[1]);                  // thousands
            const int diff500 = ya_total_500(s.w[4], s.w[5]) - ya_total_500(s.w[6], s.w[7]);
            // bid_k = 0.5 * (gap / 1000.0) - 0.15 * (diff / 1000.0); bid = int(max(0, min(100000, round(1000 * bid_k))))
            const double bid_k = 0.5 * ((double)(gap * 1000) / 1000.0) - 0.15 * ((double)(diff500 * 500) / 1000.0);
            double bid = rint(1000.0 * bid_k);                     // Python round(): half to even
            bid = fmax(0.0, fmin(100000.0, bid));
            a = target * YA_N_BID_LEVEL + (int)bid / 500;          // can leave 0..201: quirk Q11
        } else {                                                   // _choose_scoring, :134-169
            const uint32_t carry = s.w[2];
            const int nd = ya_dice_count(carry);
            a = 0;
            if (nd >= 5) {
                int best, idx;
                ya_greedy_best(carry, nd, used, upper, part[warp], sm, lane, best, idx);
                if (best >= 0) a = YA_N_BID + idx;
            }
        }
        if (lane == 0) {
            if (raw) raw[g] = a;
            if (!ya_is_legal(desc, a)) {                           // :202-206 / :211-214: random legal move instead
                if (fallback) {
                    int count = ya_legal_count(desc);
                    a = 0;
                    if (count) {
                        YaDraw d = ya_draw(seed, (uint32_t)(game_base + g), episode ? episode[g] : 0u,
                                           ply ? (uint32_t)ply[g] : 0u, YA_TAG_ACTION, 0, 0);
                        a = ya_nth_legal(desc, (int)__umulhi(d.pick, (uint32_t)count));
                    }
                } else {
                    a = -1;                                        // the host falls back (np.random.choice, like the reference)
                }
            }
            actions[g] = a;
        }
    }
}

// ------------------------------------------------------------------ fused ply
// One launch = one ply of every game under the uniform random-legal policy
// (YachtPlayers.py:174-183 + Arena.py:49-71): legal mask materialised (optional), action
// sampled with Philox, transition applied, outcome reported, finished games re-dealt.
template <int GAMES>
__global__ void __launch_bounds__(kThreads, 4)
ya_k_play_ply(uint4* __restrict__ states, int64_t stride, int8_t* __restrict__ players, int32_t* __restrict__ ply,
              uint32_t* __restrict__ episode, int32_t* __restrict__ actions, float* __restrict__ outcome,
              uint8_t* __restrict__ masks, int32_t* __restrict__ err_flag,
              int64_t n, uint64_t seed, uint64_t game_base, int auto_reset) {
    static_assert(GAMES % kSuperGames == 0 && GAMES <= 64, "chunk must be whole superblocks");
    __shared__ YaMaskSmem sm;
    __shared__ uint32_t desc_s[GAMES];
    __shared__ uint32_t fill_s[GAMES];
    const int64_t g0 = (int64_t)blockIdx.x * GAMES;
    const int ng = (int)min((int64_t)GAMES, n - g0);
    if (masks) ya_mask_smem_init(sm, threadIdx.x, blockDim.x);
    for (int t = threadIdx.x; t < ng; t += blockDim.x) {
        const int64_t g = g0 + t;
        YaState s = ya_load(states, stride, g);
        int pl = players[g];
        uint32_t p = (uint32_t)ply[g], ep = episode[g];
        uint32_t gid = (uint32_t)(game_base + g);
        uint32_t desc = ya_mask_desc(s, pl);
        desc_s[t] = desc;
        fill_s[t] = ya_fill_bits(desc);
        int count = ya_legal_count(desc);
        int a = 0;
        if (count) {
            YaDraw da = ya_draw(seed, gid, ep, p, YA_TAG_ACTION, 0, 0);
            a = ya_nth_legal(desc, (int)__umulhi(da.pick, (uint32_t)count));
        }
        YaDraw d;
        d.roll_a = d.roll_b = d.tie = d.pick = 0;
        if (ya_draw_needs(s, pl, a)) d = ya_draw(seed, gid, ep, p, YA_TAG_REAL, 0, 0);
        int st;
        int np = ya_transition(s, pl, a, d, &st);
        if (st != YA_OK && err_flag) atomicOr(err_flag, 1 << st);
        float res = ya_game_ended(s, 1);
        p += 1;
        if (res != 0.0f && auto_reset) {
            ep += 1;
            p = 0;
            np = 1;
            s = ya_fresh_state(seed, gid, ep);
        }
        ya_store(states, stride, g, s);
        players[g] = (int8_t)np;
        ply[g] = (int32_t)p;
        episode[g] = ep;
        actions[g] = a;
        outcome[g] = res;
    }
    if (masks) {
        __syncthreads();
        ya_mask_smem_link(sm, fill_s, ng, GAMES, threadIdx.x, blockDim.x);
        __syncthreads();
        uint8_t* out = masks + g0 * YA_N_ACTION;
        ya_write_mask_chunk(out, ng, sm, threadIdx.x, blockDim.x);
        __syncthreads();
        ya_patch_five_dice_rows(out, ng, desc_s, threadIdx.x, blockDim.x);
    }
}

// Same ply, array-of-records I/O for the host-buffer path: one 64-byte record per game
// {w0..w7, episode, ply, player, action(out), outcome(out), finished, P1 wins, P2 wins} so a slice is ONE
// contiguous PCIe copy in each direction.  The mask (if requested) is still streamed to HBM.
template <int GAMES>
__global__ void __launch_bounds__(kThreads, 4)
ya_k_play_ply_records(uint4* __restrict__ rec, uint8_t* __restrict__ masks, int32_t* __restrict__ err_flag,
                      int64_t n, uint64_t seed, uint64_t game_base, int auto_reset) {
    __shared__ YaMaskSmem sm;
    __shared__ uint32_t desc_s[GAMES];
    __shared__ uint32_t fill_s[GAMES];
    const int64_t g0 = (int64_t)blockIdx.x * GAMES;
    const int ng = (int)min((int64_t)GAMES, n - g0);
    if (masks) ya_mask_smem_init(sm, threadIdx.x, blockDim.x);
    for (int t = threadIdx.x; t < ng; t += blockDim.x) {
        uint4* r = rec + (g0 + t) * 4;
        uint4 a = r[0], b = r[1], c = r[2], tally = r[3];
        YaState s;
        s.w[0] = a.x; s.w[1] = a.y; s.w[2] = a.z; s.w[3] = a.w;
        s.w[4] = b.x; s.w[5] = b.y; s.w[6] = b.z; s.w[7] = b.w;
        uint32_t ep = c.x, p = c.y;
        int pl = (int)c.z;
        uint32_t gid = (uint32_t)(game_base + g0 + t);
        uint32_t desc = ya_mask_desc(s, pl);
        desc_s[t] = desc;
        fill_s[t] = ya_fill_bits(desc);
        int count = ya_legal_count(desc);
        int act = 0;
        if (count) {
            YaDraw da = ya_draw(seed, gid, ep, p, YA_TAG_ACTION, 0, 0);
            act = ya_nth_legal(desc, (int)__umulhi(da.pick, (uint32_t)count));
        }
        YaDraw d;
        d.roll_a = d.roll_b = d.tie = d.pick = 0;
        if (ya_draw_needs(s, pl, act)) d = ya_draw(seed, gid, ep, p, YA_TAG_REAL, 0, 0);
        int st;
        int np = ya_transition(s, pl, act, d, &st);
        if (st != YA_OK && err_flag) atomicOr(err_flag, 1 << st);
        float res = ya_game_ended(s, 1);
        p += 1;
        if (res != 0.0f && auto_reset) { ep += 1; p = 0; np = 1; s = ya_fresh_state(seed, gid, ep); }
        r[0] = make_uint4(s.w[0], s.w[1], s.w[2], s.w[3]);
        r[1] = make_uint4(s.w[4], s.w[5], s.w[6], s.w[7]);
        r[2] = make_uint4(ep, p, (uint32_t)np, (uint32_t)act);
        if (res != 0.0f) { tally.y += 1; tally.z += res > 0.5f; tally.w += res < -0.5f; }   // finished / P1 wins / P2 wins
        r[3] = make_uint4(__float_as_uint(res), tally.y, tally.z, tally.w);
    }
    if (masks) {
        __syncthreads();
        ya_mask_smem_link(sm, fill_s, ng, GAMES, threadIdx.x, blockDim.x);
        __syncthreads();
        uint8_t* out = masks + g0 * YA_N_ACTION;
        ya_write_mask_chunk(out, ng, sm, threadIdx.x, blockDim.x);
        __syncthreads();
        ya_patch_five_dice_rows(out, ng, desc_s, threadIdx.x, blockDim.x);
    }
}

inline int blocks_for(int64_t n, int per) { return (int)((n + per - 1) / per); }

}  // namespace

// =================================================================== C ABI
extern "C" {

int ya_abi_version(void) { return YA_ABI_VERSION; }

int ya_set_device(int device) {
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return (int)e;
    return ya_ensure_score_table(cudaStreamPerThread);
}

int ya_init_states(uint32_t* states, int64_t stride, int8_t* players, int32_t* ply, const uint32_t* episode,
                   int64_t n, uint64_t seed, uint64_t game_base, void* stream) {
    if (n <= 0) return 0;
    ya_k_init<<<blocks_for(n, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<uint4*>(states), stride, players, ply, episode, n, seed, game_base);
    return (int)cudaGetLastError();
}

int ya_next_state(const uint32_t* states_in, int64_t in_stride, const int8_t* players, const int32_t* actions,
                  uint32_t* states_out, int64_t out_stride, int8_t* next_players, int32_t* status, int64_t n,
                  int draw_mode, const uint8_t* injected, uint64_t seed, uint64_t game_base,
                  const uint32_t* episode, const int32_t* ply, uint32_t tag, const uint32_t* depth_sim, void* stream) {
    if (n <= 0) return 0;
    ya_k_next_state<<<blocks_for(n, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const uint4*>(states_in), in_stride, players, actions,
        reinterpret_cast<uint4*>(states_out), out_stride, next_players, status, n,
        draw_mode, injected, seed, game_base, episode, ply, tag, depth_sim);
    return (int)cudaGetLastError();
}

int ya_valid_moves(const uint32_t* states, int64_t stride, const int8_t* players, uint8_t* masks, int64_t n, void* stream) {
    if (n <= 0) return 0;
    if ((reinterpret_cast<uintptr_t>(masks) & 15u) != 0) return (int)cudaErrorMisalignedAddress;
    ya_k_valid_moves<<<blocks_for(n, kMaskGames), kThreads, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const uint4*>(states), stride, players, masks, n);
    return (int)cudaGetLastError();
}

int ya_game_ended(const uint32_t* states, int64_t stride, const int8_t* players, float* out, int64_t n, void* stream) {
    if (n <= 0) return 0;
    ya_k_game_ended<<<blocks_for(n, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const uint4*>(states), stride, players, out, n);
    return (int)cudaGetLastError();
}

int ya_canonical_form(const uint32_t* states_in, int64_t in_stride, const int8_t* players,
                      uint32_t* states_out, int64_t out_stride, int64_t n, void* stream) {
    if (n <= 0) return 0;
    ya_k_canonical<<<blocks_for(n, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const uint4*>(states_in), in_stride, players, reinterpret_cast<uint4*>(states_out), out_stride, n);
    return (int)cudaGetLastError();
}

int ya_features(const uint32_t* states, int64_t stride, float* features, int64_t n, void* stream) {
    if (n <= 0) return 0;
    ya_k_features<<<blocks_for(n * YA_N_FEATURE, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const uint4*>(states), stride, features, n);
    return (int)cudaGetLastError();
}

int ya_random_action(const uint32_t* states, int64_t stride, const int8_t* players, int32_t* actions, int64_t n,
                     uint64_t seed, uint64_t game_base, const uint32_t* episode, const int32_t* ply, void* stream) {
    if (n <= 0) return 0;
    ya_k_random_action<<<blocks_for(n, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const uint4*>(states), stride, players, actions, n, seed, game_base, episode, ply);
    return (int)cudaGetLastError();
}

int ya_enumerate_scores(const uint32_t* states, int64_t stride, const int8_t* players, uint8_t* scores, int64_t n,
                        void* stream) {
    if (n <= 0) return 0;
    if ((reinterpret_cast<uintptr_t>(scores) & 15u) != 0) return (int)cudaErrorMisalignedAddress;
    int rc = ya_ensure_score_table((cudaStream_t)stream);
    if (rc) return rc;
    int blocks = (int)min((int64_t)148 * 5, (n + kEnumWarps - 1) / kEnumWarps);
    ya_k_enumerate_scores<<<blocks, kEnumWarps * 32, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const uint4*>(states), stride, players, scores, n);
    return (int)cudaGetLastError();
}

int ya_greedy_action(const uint32_t* states, int64_t stride, const int8_t* players, int32_t* actions, int32_t* raw,
                     int64_t n, int fallback, uint64_t seed, uint64_t game_base, const uint32_t* episode,
                     const int32_t* ply, void* stream) {
    if (n <= 0) return 0;
    int rc = ya_ensure_score_table((cudaStream_t)stream);
    if (rc) return rc;
    int blocks = (int)min((int64_t)148 * 8, (n + kEnumWarps - 1) / kEnumWarps);
    ya_k_greedy_action<<<blocks, kEnumWarps * 32, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const uint4*>(states), stride, players, actions, raw, n, fallback, seed, game_base, episode, ply);
    return (int)cudaGetLastError();
}

int ya_play_ply(uint32_t* states, int64_t stride, int8_t* players, int32_t* ply, uint32_t* episode,
                int32_t* actions, float* outcome, uint8_t* masks, int32_t* err_flag,
                int64_t n, uint64_t seed, uint64_t game_base, int auto_reset, void* stream) {
    if (n <= 0) return 0;
    if (masks && (reinterpret_cast<uintptr_t>(masks) & 15u) != 0) return (int)cudaErrorMisalignedAddress;
    constexpr int G = 64;
    ya_k_play_ply<G><<<blocks_for(n, G), kThreads, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<uint4*>(states), stride, players, ply, episode, actions, outcome, masks, err_flag,
        n, seed, game_base, auto_reset);
    return (int)cudaGetLastError();
}

int ya_play_ply_records(uint32_t* records, uint8_t* masks, int32_t* err_flag, int64_t n, uint64_t seed,
                        uint64_t game_base, int auto_reset, void* stream) {
    if (n <= 0) return 0;
    if (masks && (reinterpret_cast<uintptr_t>(masks) & 15u) != 0) return (int)cudaErrorMisalignedAddress;
    if ((reinterpret_cast<uintptr_t>(records) & 15u) != 0) return (int)cudaErrorMisalignedAddress;
    constexpr int G = 64;
    ya_k_play_ply_records<G><<<blocks_for(n, G), kThreads, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<uint4*>(records), masks, err_flag, n, seed, game_base, auto_reset);
    return (int)cudaGetLastError();
}

}  // extern "C"
