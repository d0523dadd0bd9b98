"""MCTS host side: the batched self-play searcher (thousands of lock-step games, one warp per game
on device) and the drop-in ``MCTS`` class with the reference's constructor / ``getActionProb`` /
``search`` surface (/root/reference/MCTS.py:16,28,56), both driving the same CUDA kernels
(csrc/ya_mcts.cu) through the C ABI.  No tree logic runs on the host.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib
from .engine import ACTION_SIZE, FEATURE_SIZE, BatchedYacht
from .layout import YachtBoard, boards_to_planes, pack_state, planes_to_boards

ERR_BITS = {0x100: "MCTS node pool full (raise max_nodes)", 0x200: "MCTS arena full (raise arena_mb_per_game)",
            0x400: "MCTS path deeper than 16", 0x800: "rule error inside search"}


def _next_pow2(x):
    p = 1
    while p < x:
        p *= 2
    return p


class TreePool:
    """Device storage of n per-game trees (include/yacht_b200.h: ya_mcts_tree)."""

    def __init__(self, n, num_sims, device, arena_mb_per_game=None, max_nodes=None, rows=0):
        self.n = int(n)
        self.rows = int(rows)                       # YA_ROWS_F32 (0) / YA_ROWS_FP16 (2) / YA_ROWS_BF16 (3)
        self.device = torch.device(device)
        lib = _lib.load()
        self.max_nodes = int(max_nodes or (6 * num_sims + 16))
        assert self.max_nodes < 65535
        self.ht_size = _next_pow2(2 * self.max_nodes)
        if arena_mb_per_game is None:
            # measured peak (random-init net, 100 sims, profiles/tools/arena_peak.py): 39 KB per simulation of the
            # longest-lived tree (rounds 1+2: six plies without pruning); 48 KB/sim leaves ~20 % head-room,
            # overflow is detected and raised.
            # 16-bit logit rows (ROWS_FP16 / ROWS_BF16) take 2.4 instead of 4.1 bytes per legal move: measured peak 22.4 KB
            # per simulation (2.24 MB per game at 100 sims), 28 KB/sim allocated.
            per_sim_kb = 48.0 if self.rows == 0 else 28.0
            arena_mb_per_game = max(num_sims * per_sim_kb / 1024.0, 0.25)
        words = int(arena_mb_per_game * 2 ** 20) // 4
        self.arena_words = words - (words % 8)          # every game's arena starts 32-byte aligned (logit rows)
        d = self.device
        self.nodes = torch.zeros((self.n, self.max_nodes, lib.ya_mcts_node_words()), dtype=torch.int32, device=d)
        self.ht = torch.zeros((self.n, self.ht_size), dtype=torch.int16, device=d)
        self.arena = torch.empty((self.n, self.arena_words), dtype=torch.int32, device=d)
        self.meta = torch.zeros((self.n, 4), dtype=torch.int32, device=d)
        self.cursor = torch.zeros((self.n, lib.ya_mcts_cursor_words()), dtype=torch.int32, device=d)
        self.struct = _lib.MctsTreeStruct(
            self.nodes.data_ptr(), self.ht.data_ptr(), self.arena.data_ptr(), self.meta.data_ptr(),
            self.cursor.data_ptr(), self.n, self.max_nodes, self.ht_size, self.arena_words)
        self.ref = ctypes.byref(self.struct)
        self.lib = lib
        self.reset()

    def bytes(self):
        return sum(t.numel() * t.element_size() for t in (self.nodes, self.ht, self.arena, self.meta, self.cursor))

    def reset(self, which=None):
        _lib.check(self.lib.ya_mcts_reset(self.ref, _lib.ptr(which), _lib.current_stream()), "ya_mcts_reset")

    def node_counts(self):
        return self.meta[:, 0]


class UniformEvaluator:
    """BASELINE.json configs[2]: P = 1/3226 (float32), v = 0; handled inside ya_mcts_expand."""
    uniform = True
    p = float(np.float32(1.0) / np.float32(ACTION_SIZE))
    v = 0.0


class FusedYachtEvaluator:
    """The yacht NNet forward (yacht/pytorch/YachtNNet.py:62-70) for a whole wave of leaves as ONE hand-written
    tcgen05 kernel on CTA pairs (csrc/ya_forward.cu): features in; tanh value, per-row max logit and the 16-bit policy
    logits out -- scattered straight into the leaves' rows of the tree pool (scatter=...), or as a dense matrix padded to
    3232 columns -- consumed by ya_mcts_expand_logits; rows do not depend on the batch they sit in.

    precision="fp16" (default): activations, weights and logits in IEEE half; accumulation, bias, SiLU, LayerNorm and the
    residual sum in float32 (the residual stream is stored as half between blocks) -- inside the error of the reference's
    CUDA predict (fp16 autocast, yacht/NNet.py:186-193).
    precision="bf16": the same kernel with bfloat16 operands (same tensor-core rate, 8-bit mantissa).
    tiles_per_cta: 1 = one 128-leaf tile per CTA, 2 = two (the tensor core works on one tile under the other tile's
    epilogue: for waves above one tile per SM), 0 = by wave size.  Scheduling only: the rows are bit-identical.
    Weights come from any module with YachtNNet's state dict (:25-52), hidden width 256 (main.py:40)."""
    uniform = False
    returns_logits = True
    supports_scatter = True
    PADDED = 3232

    def __init__(self, net, max_batch, precision="fp16", tiles_per_cta=0):
        if precision not in ("fp16", "bf16"):
            raise ValueError("precision must be 'fp16' or 'bf16', got %r" % (precision,))
        if tiles_per_cta not in (0, 1, 2):
            raise ValueError("tiles_per_cta must be 0 (by wave size), 1 or 2")
        self.tiles_per_cta = tiles_per_cta          # scheduling only: both kernels give bit-identical rows
        self.precision = precision
        self.fp16 = precision == "fp16"
        self.op_dtype = torch.float16 if self.fp16 else torch.bfloat16
        self.last_row_max = None
        sd = {k: v.detach() for k, v in net.state_dict().items()}
        dev = next(net.parameters()).device
        self.hidden = sd["inp.0.weight"].shape[0]
        self.nblocks = len({k.split(".")[1] for k in sd if k.startswith("blocks.")})
        a = sd["pi_head.2.weight"].shape[0]
        if self.hidden != 256 or sd["v_head.2.weight"].shape[0] != 128 or a > self.PADDED or sd["inp.0.weight"].shape[1] > 64:
            raise ValueError("FusedYachtEvaluator handles YachtNNet(hidden=256) with <= 64 inputs, a 128-wide value head and "
                             "<= 3232 actions (main.py:40-42)")
        self.lib = _lib.load()
        self.eps = 1e-5
        f32 = lambda t: t.to(device=dev, dtype=torch.float32).reshape(-1)
        images, params = [], []
        for i in range(self.nblocks):
            p = "blocks.%d." % i
            for fc, ln in (("fc1", "ln1"), ("fc2", "ln2")):
                images.append(self.pair_image(sd[p + fc + ".weight"].to(dev), 128))
                params.append(torch.stack([0.5 * sd[p + fc + ".bias"], sd[p + ln + ".weight"], sd[p + ln + ".bias"]]).float())
        # params[layer] = bias / 2 | gamma | beta: the kernel's SiLU works on (z + b) / 2
        trunk_w = torch.cat(images) if images else torch.zeros(0, dtype=torch.uint8, device=dev)
        trunk_p = torch.stack(params).to(dev).reshape(-1) if params else torch.zeros(0, device=dev)
        w_in = torch.zeros((256, 64), dtype=torch.float32, device=dev)
        w_in[:, :sd["inp.0.weight"].shape[1]].copy_(sd["inp.0.weight"])
        w_pi = torch.zeros((26 * 128, 256), dtype=torch.float32, device=dev)
        w_pi[:a].copy_(sd["pi_head.2.weight"])
        # the padding columns (>= the action count; their weights are zero) get a bias of -inf: their logits can never be a
        # row maximum, and nobody reads them (they are not legal moves)
        b_pi = torch.full((26 * 128,), float("-inf"), dtype=torch.float32, device=dev)
        b_pi[:a].copy_(sd["pi_head.2.bias"])
        wparts = [self.pair_image(w_in, 256), trunk_w, self.pair_image(sd["v_head.2.weight"].to(dev), 128),
                  self.pair_image(w_pi, 128, tile_major=True)]
        pparts = [f32(sd["inp.0.bias"]), f32(sd["inp.1.weight"]), f32(sd["inp.1.bias"]), trunk_p,
                  f32(sd["v_head.0.weight"]), f32(sd["v_head.0.bias"]), f32(sd["v_head.2.bias"]), f32(sd["v_head.4.weight"]),
                  torch.cat([f32(sd["v_head.4.bias"]), torch.zeros(3, device=dev)]),
                  f32(sd["pi_head.0.weight"]), f32(sd["pi_head.0.bias"]), b_pi]
        wsizes = [p.numel() for p in wparts]
        psizes = [p.numel() for p in pparts]
        self.fw_w = torch.cat(wparts).contiguous()
        self.fw_p = torch.cat(pparts).contiguous()
        w_off = [0, wsizes[0], wsizes[0] + wsizes[1], wsizes[0] + wsizes[1] + wsizes[2]]
        p_in, p_trunk = 0, sum(psizes[:3])
        p_v = p_trunk + psizes[3]
        p_pi_ln = p_v + sum(psizes[4:9])
        self.fw_off = (ctypes.c_int64 * 9)(*w_off, p_in, p_trunk, p_v, p_pi_ln, p_pi_ln + 512)
        assert sum(psizes[4:9]) == 772 and psizes[3] == 768 * 2 * self.nblocks
        self.device = dev
        self._alloc(int(max_batch), dev)

    def swizzled_image(self, w):
        """[rows][K] weight (K a multiple of 64) -> shared-memory image for tcgen05.mma in the operand format: K-blocks
        of [rows][64 x 16 bit], 16-byte chunk c of row r stored at chunk c ^ (r & 7) (128-byte swizzle)."""
        rows, k = w.shape
        w = w.to(self.op_dtype).contiguous().view(rows, k // 64, 8, 8)
        r = torch.arange(rows, device=w.device).view(rows, 1, 1)
        src_chunk = (torch.arange(8, device=w.device).view(1, 1, 8) ^ (r & 7)).expand(rows, k // 64, 8)
        img = torch.gather(w, 2, src_chunk.unsqueeze(-1).expand(rows, k // 64, 8, 8))
        return img.permute(1, 0, 2, 3).contiguous().view(torch.uint8).reshape(-1)

    def pair_image(self, w, n_tile, tile_major=False):
        """Weight [rows][K] for the CTA-pair kernel: the rows are consumed in blocks of n_tile output columns (one
        tcgen05.mma.cta_group::2 N extent), and of each block rank 0 of the pair holds the first n_tile / 2 rows, rank 1 the
        rest.  Layout: trunk / value / input = [rank][block][image of n_tile / 2 rows] (each CTA's share of a whole layer is
        contiguous); policy head (many blocks streamed one at a time) the same with rows = one tile, i.e. [tile][rank][image]."""
        rows, k = w.shape
        half = n_tile // 2
        blocks = w.reshape(rows // n_tile, 2, half, k)                  # [block][rank][half rows][K]
        if tile_major:                                                  # policy head
            order = blocks.reshape(-1, half, k)
        else:
            order = blocks.permute(1, 0, 2, 3).reshape(-1, half, k)    # rank-major
        return torch.cat([self.swizzled_image(b) for b in order])

    def _alloc(self, n, dev):
        self.max_batch = n
        self.logits = None                          # dense [n, 3232] logit matrix: only allocated if someone asks for it
        self.values = torch.empty(n, dtype=torch.float32, device=dev)
        self.row_max = torch.empty(n, dtype=torch.float32, device=dev)

    def with_private_buffers(self, max_batch):
        """Same weights, own activation buffers (one instance per concurrently running game group)."""
        import copy
        other = copy.copy(self)
        other._alloc(int(max_batch), self.device)
        return other

    @torch.no_grad()
    def __call__(self, features, need_eval=None, leaf_states=None, scatter=None):
        """scatter = (leaf_dst uint64[n], leaf_desc uint32[n]) from ya_mcts_select: the policy head's epilogue writes every
        leaf's LEGAL logits straight into its row of the tree pool and no dense logit matrix exists (returns None for it);
        otherwise the dense [n, 3232] 16-bit logits are returned."""
        n = features.shape[0]
        values = self.values[:n]
        self.last_row_max = self.row_max[:n]                              # consumed by ya_mcts_expand_logits
        logits = None
        if scatter is None:
            if self.logits is None:
                self.logits = torch.empty((self.max_batch, self.PADDED), dtype=self.op_dtype, device=self.device)
            logits = self.logits[:n]
        dst, desc = scatter if scatter is not None else (None, None)
        _lib.check(self.lib.ya_nn_forward_tiles(_lib.ptr(features), _lib.ptr(logits), _lib.ptr(values),
                                                _lib.ptr(self.last_row_max), _lib.ptr(self.fw_w),
                                                _lib.ptr(self.fw_p), self.fw_off, self.nblocks, n, self.eps,
                                                1 if self.fp16 else 0, _lib.ptr(dst), _lib.ptr(desc), self.tiles_per_cta,
                                                _lib.current_stream()),
                   "ya_nn_forward_tiles")
        return logits, values


class _Group:
    """A contiguous slice of the games with its own view of the tree pool, buffers and stream."""

    def __init__(self, mcts, g0, g1, evaluator, stream):
        pool, env = mcts.pool, mcts.env
        self.g0, self.g1, self.n = g0, g1, g1 - g0
        self.struct = _lib.MctsTreeStruct(
            pool.nodes[g0:g1].data_ptr(), pool.ht[g0:g1].data_ptr(), pool.arena[g0:g1].data_ptr(),
            pool.meta[g0:g1].data_ptr(), pool.cursor[g0:g1].data_ptr(), self.n, pool.max_nodes, pool.ht_size, pool.arena_words)
        self.ref = ctypes.byref(self.struct)
        self.states_ptr = ctypes.c_void_p(env.states.data_ptr() + 16 * g0)     # plane stride stays env.n
        self.players, self.ply, self.episode = env.players[g0:g1], env.ply[g0:g1], env.episode[g0:g1]
        self.features, self.need_eval = mcts.features[g0:g1], mcts.need_eval[g0:g1]
        self.leaf_dst = mcts.leaf_dst[g0:g1] if mcts.leaf_dst is not None else None
        self.leaf_desc = mcts.leaf_desc[g0:g1] if mcts.leaf_desc is not None else None
        self.sim_counter = torch.zeros(1, dtype=torch.int32, device=env.device)
        # global id of this slice's first game, in device memory: a captured launch freezes by-value arguments, so the
        # graph reads the base from here and stays valid when the pool moves on to the next wave of games
        self.base_dev = torch.full((1,), env.game_base + g0, dtype=torch.int64, device=env.device)
        self.evaluator = evaluator
        self.stream = stream


class BatchedMCTS:
    """numMCTSSims simulations for every game of a BatchedYacht per move, tree kept per episode
    (Coach.py:93) and pruned at round boundaries.

    With a network evaluator the games are split into `groups` independent slices whose
    select -> forward -> expand chains run on separate streams: the latency-bound tree kernels of one
    slice overlap the tensor-core forward of the other (games never interact, so this is exact)."""

    def __init__(self, env: BatchedYacht, num_sims, cpuct=1.5, evaluator=None, temp_threshold=15,
                 arena_mb_per_game=None, max_nodes=None, want_leaf_states=False, groups=None, forward_priority=True):
        self.env = env
        self.lib = env.lib
        self.num_sims = int(num_sims)
        self.cpuct = float(cpuct)
        self.temp_threshold = int(temp_threshold)
        self.evaluator = evaluator or UniformEvaluator()
        ev = self.evaluator
        # how the pool stores priors: evaluators that hand over raw 16-bit logits get logit rows (the format of their logits),
        # float32-policy evaluators and the uniform prior the float32 / constant rows of the bit-exact reference path
        self.rows = 0
        if getattr(ev, "returns_logits", False):
            dt = getattr(ev, "op_dtype", None) or getattr(getattr(ev, "logits", None), "dtype", torch.bfloat16)
            self.rows = 2 if dt == torch.float16 else 3
        self.scatter = bool(self.rows and getattr(ev, "supports_scatter", False) and not want_leaf_states)
        self.pool = TreePool(env.n, self.num_sims, env.device, arena_mb_per_game, max_nodes, rows=self.rows)
        d = env.device
        n = env.n
        self.leaf_dst = torch.zeros(n, dtype=torch.int64, device=d) if self.scatter else None
        self.leaf_desc = torch.zeros(n, dtype=torch.int32, device=d) if self.scatter else None
        self.features = torch.zeros((n, FEATURE_SIZE), dtype=torch.float32, device=d)
        self.need_eval = torch.zeros(n, dtype=torch.uint8, device=d)
        self.leaf_states = torch.zeros((2, n, 4), dtype=torch.int32, device=d) if want_leaf_states else None
        self.counts = torch.zeros((n, ACTION_SIZE), dtype=torch.int32, device=d)
        self.visits = torch.zeros(n, dtype=torch.int32, device=d)
        self.err_flag = torch.zeros(1, dtype=torch.int32, device=d)
        self.picked = torch.zeros(n, dtype=torch.int32, device=d)
        self.sims_run = 0
        self.graph = None
        # uniform evaluator: all simulations of a move in one kernel.  That kernel's nodes carry one prior value and a
        # visited bitmask instead of a prior row, so the choice holds for the life of the trees (set it before the
        # first search, or call pool.reset() when changing it).
        self.fuse_uniform = True
        uniform = getattr(self.evaluator, "uniform", False)
        if groups is None:
            groups = 1       # >1 overlaps slices on separate streams; measured gain on B200 is ~2 %, off by default
        if want_leaf_states or uniform:
            groups = 1
        self.groups = []
        per = (n + groups - 1) // groups
        for i in range(groups):
            g0, g1 = i * per, min(n, (i + 1) * per)
            if g0 >= g1:
                break
            ev = self.evaluator
            if i > 0 and hasattr(ev, "with_private_buffers"):
                ev = ev.with_private_buffers(g1 - g0)
            stream = torch.cuda.Stream(device=d) if groups > 1 else None
            self.groups.append(_Group(self, g0, g1, ev, stream))
            if groups > 1 and forward_priority:
                # a forward CTA needs a whole SM; at high priority its pending CTAs are placed before any further small
                # block of the other groups' tree kernels, so the SMs drain for it instead of starving it
                self.groups[-1].fwd_stream = torch.cuda.Stream(device=d, priority=-1)
        self.sim_counter = self.groups[0].sim_counter
        self._base_on_device = env.game_base

    def sync_game_base(self):
        """Publishes env.game_base to the device scalars the select kernel reads (call-free for the user: search()
        does it whenever env.game_base changed, e.g. coach.self_play_in_waves moving to the next wave)."""
        if self._base_on_device != self.env.game_base:
            for grp in self.groups:
                grp.base_dev.fill_(self.env.game_base + grp.g0)
            self._base_on_device = self.env.game_base

    def capture_graph(self):
        """Capture ONE simulation wave (select, evaluator forward, expand of every group) as a CUDA graph;
        the simulation index is a device counter the expand kernel bumps, so the same graph is replayed
        numMCTSSims times per move without any host-side launch work in between.  The global game ids come from
        device memory too (grp.base_dev), so the graph survives env.game_base changes (waves on one pool).
        Capture on EMPTY trees only: the warm-up simulations are rolled back by restoring the node tables, not the
        arena (edge statistics), which is only consistent when there was nothing in the arena before."""
        if int(self.pool.meta[:, 0].max().item()) != 0:
            raise _lib.YachtB200Error("capture_graph needs empty trees (call pool.reset() first, or capture before the first search)")
        self.sync_game_base()
        side = torch.cuda.Stream(device=self.env.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                 # warm-up outside capture (lazy inits of the evaluator)
            snap = [t.clone() for t in (self.pool.nodes, self.pool.ht, self.pool.meta, self.pool.cursor)]
            for _ in range(2):
                self.simulate(None)
            for t, s0 in zip((self.pool.nodes, self.pool.ht, self.pool.meta, self.pool.cursor), snap):
                t.copy_(s0)
            for grp in self.groups:
                grp.sim_counter.zero_()
            self.err_flag.zero_()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize(self.env.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.simulate(None)
        self.sims_run = 0

    # ------------------------------------------------------------------ one simulation wave
    def _simulate_group(self, grp, sim):
        env, s = self.env, _lib.current_stream()
        leaf = self.leaf_states if len(self.groups) == 1 else None
        _lib.check(self.lib.ya_mcts_select(
            grp.ref, grp.states_ptr, env.n, _lib.ptr(grp.players), _lib.ptr(grp.ply), _lib.ptr(grp.episode),
            env.seed, env.game_base + grp.g0, 0 if sim is None else sim, _lib.ptr(grp.sim_counter) if sim is None else None,
            _lib.ptr(grp.base_dev) if sim is None else None, self.cpuct, None, _lib.ptr(grp.features), _lib.ptr(grp.need_eval),
            _lib.ptr(leaf), self.rows, _lib.ptr(grp.leaf_dst), _lib.ptr(grp.leaf_desc), _lib.ptr(self.err_flag), s), "ya_mcts_select")
        counter = _lib.ptr(grp.sim_counter) if sim is None else None
        ev = grp.evaluator
        if getattr(ev, "uniform", False):
            _lib.check(self.lib.ya_mcts_expand(grp.ref, None, None, 1, ev.p, ev.v, counter, _lib.ptr(self.err_flag), s),
                       "ya_mcts_expand")
            return
        fwd = getattr(grp, "fwd_stream", None)                        # multi-group mode: the forward on its own high-priority stream
        if fwd is not None:
            fwd.wait_stream(torch.cuda.current_stream())
            ctx = torch.cuda.stream(fwd)
            ctx.__enter__()
        if self.scatter:                                              # legal logits go from the tensor core into the tree rows
            pi, v = ev(grp.features, grp.need_eval, leaf, scatter=(grp.leaf_dst, grp.leaf_desc))
        else:
            pi, v = ev(grp.features, grp.need_eval, leaf)
        if fwd is not None:
            ctx.__exit__(None, None, None)
            torch.cuda.current_stream().wait_stream(fwd)
        assert v.dtype == torch.float32 and v.is_contiguous() and v.shape == (grp.n,)
        if self.rows:
            assert self.scatter or (pi.dtype == (torch.float16 if self.rows == 2 else torch.bfloat16) and pi.is_contiguous()
                                    and pi.shape[0] == grp.n)
            row_max = getattr(ev, "last_row_max", None)               # per-row max logit, if the evaluator has it
            assert row_max is None or (row_max.dtype == torch.float32 and row_max.shape == (grp.n,))
            _lib.check(self.lib.ya_mcts_expand_logits(grp.ref, _lib.ptr(pi), 1 if self.rows == 2 else 0,
                                                      0 if pi is None else pi.shape[1], _lib.ptr(row_max), _lib.ptr(v),
                                                      counter, _lib.ptr(self.err_flag), s), "ya_mcts_expand_logits")
            return
        assert pi.dtype == torch.float32 and pi.is_contiguous() and pi.shape == (grp.n, ACTION_SIZE)
        _lib.check(self.lib.ya_mcts_expand(grp.ref, _lib.ptr(pi), _lib.ptr(v), 0, 0.0, 0.0, counter,
                                           _lib.ptr(self.err_flag), s), "ya_mcts_expand")

    def simulate(self, sim):
        if len(self.groups) == 1:
            self._simulate_group(self.groups[0], sim)
        else:
            cur = torch.cuda.current_stream()
            for grp in self.groups:
                grp.stream.wait_stream(cur)
                with torch.cuda.stream(grp.stream):
                    self._simulate_group(grp, sim)
            for grp in self.groups:
                cur.wait_stream(grp.stream)
        self.sims_run += self.env.n

    def search(self):
        """getActionProb's simulation loop (MCTS.py:37-38) for every game."""
        ev = self.evaluator
        self.sync_game_base()
        if getattr(ev, "uniform", False) and self.fuse_uniform:
            env = self.env                  # no network between select and expand: the whole loop is one launch
            _lib.check(self.lib.ya_mcts_search_uniform(
                self.pool.ref, _lib.ptr(env.states), env.n, _lib.ptr(env.players), _lib.ptr(env.ply), _lib.ptr(env.episode),
                env.seed, env.game_base, self.num_sims, self.cpuct, ev.p, ev.v, None, _lib.ptr(self.err_flag),
                _lib.current_stream()), "ya_mcts_search_uniform")
            self.sims_run += env.n * self.num_sims
            return
        if self.graph is not None:
            for grp in self.groups:
                grp.sim_counter.zero_()
            for _ in range(self.num_sims):
                self.graph.replay()
            self.sims_run += self.env.n * self.num_sims
            return
        for sim in range(self.num_sims):
            self.simulate(sim)

    def check_errors(self):
        e = int(self.err_flag.item())
        if e:
            msgs = [m for b, m in ERR_BITS.items() if e & b]
            raise _lib.YachtB200Error("MCTS kernel error %#x: %s" % (e, "; ".join(msgs)))

    def root_counts(self, with_q=False):
        env = self.env
        q = kind = None
        if with_q:
            q = torch.zeros((env.n, ACTION_SIZE), dtype=torch.float64, device=env.device)
            kind = torch.zeros((env.n, ACTION_SIZE), dtype=torch.uint8, device=env.device)
        _lib.check(self.lib.ya_mcts_root_counts(
            self.pool.ref, _lib.ptr(env.states), env.n, _lib.ptr(env.players), _lib.ptr(self.counts), _lib.ptr(self.visits),
            _lib.ptr(q), _lib.ptr(kind), _lib.current_stream()), "ya_mcts_root_counts")
        return (self.counts, self.visits, q, kind) if with_q else (self.counts, self.visits)

    def root_sparse(self, actions_out, counts_out, overflow=None):
        """Visited root edges as sorted (action, count) pairs, zero padded (canonical sparse pi)."""
        env = self.env
        _lib.check(self.lib.ya_mcts_root_sparse(
            self.pool.ref, _lib.ptr(env.states), env.n, _lib.ptr(env.players), actions_out.shape[-1], _lib.ptr(actions_out),
            _lib.ptr(counts_out), _lib.ptr(overflow), _lib.current_stream()), "ya_mcts_root_sparse")

    def pick_actions(self):
        env = self.env
        _lib.check(self.lib.ya_mcts_pick_action(
            _lib.ptr(self.counts), _lib.ptr(env.ply), _lib.ptr(env.episode), env.n, env.seed, env.game_base,
            self.temp_threshold, _lib.ptr(self.picked), _lib.current_stream()), "ya_mcts_pick_action")
        return self.picked

    def play_ply(self):
        """One move of Coach.executeEpisode (Coach.py:54-66) for every game: search, pick, step."""
        self.search()
        self.root_counts()
        actions = self.pick_actions()
        self.env.next_state(actions, check=False)
        return actions

    def new_episode(self):
        """Coach.py:93: a fresh tree per episode."""
        self.pool.reset()
        self.env.episode += 1
        self.env.reset()


# ======================================================================================= drop-in
class MCTS:
    """Drop-in for /root/reference/MCTS.py: ``MCTS(game, nnet, args)``, ``getActionProb(board, temp)``.

    The tree lives on the GPU (one game); ``nnet.predict(canonicalBoard)`` is called on the host for
    every leaf exactly like the reference does (MCTS.py:86), so any evaluator plugs in.  Dice rolled
    *inside* search (quirk Q1) come, by default, from the same hooks and global RNG streams as the
    reference's (args.search_dice = "hooks"): a seeded Coach.executeEpisode reproduces the reference bit
    for bit.  args.search_dice = "philox" uses the engine's counter stream keyed by (search_seed,
    tree_id, root index, sim, depth) instead (see DESIGN.md "Randomness")."""

    _next_tree_id = 0

    def __init__(self, game, nnet, args):
        self.game = game
        self.nnet = nnet
        self.args = args
        self.device = torch.device(getattr(game, "device", "cuda"))
        if not torch.cuda.is_available():
            raise _lib.YachtB200Error("no CUDA device: MCTS has no CPU fallback")
        self.lib = _lib.load()
        sims = int(args.numMCTSSims)
        self.pool = TreePool(1, max(sims, 1), self.device,
                             arena_mb_per_game=_arg(args, "arena_mb_per_game", None), max_nodes=_arg(args, "max_nodes", None))
        # "hooks" (default): dice rolled inside search come from game.roll_five / game.tiebreak_uniform, i.e. the
        # global numpy / random streams, consumed exactly like the reference; "philox": the engine's counter stream
        self.dice = str(_arg(args, "search_dice", "hooks"))
        assert self.dice in ("hooks", "philox")
        self.seed = int(_arg(args, "search_seed", 0))
        self.tree_id = int(_arg(args, "tree_id", MCTS._next_tree_id))
        MCTS._next_tree_id += 1
        d = self.device
        self.states = torch.zeros((2, 1, 4), dtype=torch.int32, device=d)
        self.players = torch.ones(1, dtype=torch.int8, device=d)
        self.ply = torch.zeros(1, dtype=torch.int32, device=d)
        self.features = torch.zeros((1, FEATURE_SIZE), dtype=torch.float32, device=d)
        self.need_eval = torch.zeros(1, dtype=torch.uint8, device=d)
        self.leaf_states = torch.zeros((2, 1, 4), dtype=torch.int32, device=d)
        self.err_flag = torch.zeros(1, dtype=torch.int32, device=d)
        self.counts = torch.zeros((1, ACTION_SIZE), dtype=torch.int32, device=d)
        self.visits = torch.zeros(1, dtype=torch.int32, device=d)
        self.pi_dev = torch.zeros((1, ACTION_SIZE), dtype=torch.float32, device=d)
        self.v_dev = torch.zeros(1, dtype=torch.float32, device=d)
        self.d_inj = torch.zeros(12, dtype=torch.uint8, device=d)
        self.root_index = -1          # number of getActionProb calls - 1 = the "ply" of the draw counter
        self.sim_index = 0
        self._root = None

    def _load_root(self, board):
        b = pack_state(board)
        if b != self._root:
            self.states.copy_(torch.from_numpy(boards_to_planes([b]).view(np.int32)))
            self._root = b

    def search(self, canonicalBoard):
        """One simulation from canonicalBoard (MCTS.py:56-164)."""
        self._load_root(canonicalBoard)
        if self.root_index < 0:
            self.root_index = 0
        s = _lib.current_stream()
        if self.dice == "philox":
            self.ply.fill_(self.root_index)
            _lib.check(self.lib.ya_mcts_select(
                self.pool.ref, _lib.ptr(self.states), 1, _lib.ptr(self.players), _lib.ptr(self.ply), None,
                self.seed, self.tree_id, self.sim_index, None, None, float(self.args.cpuct), None, _lib.ptr(self.features),
                _lib.ptr(self.need_eval), _lib.ptr(self.leaf_states), 0, None, None, _lib.ptr(self.err_flag), s), "ya_mcts_select")
            code = int(self.need_eval.item())
        else:
            # in-search dice through the reference's own hooks, drawn exactly when getNextState would draw them
            # (tie-break first, then rollA, rollB: yacht/YachtGame.py:287,298-299); the kernel parks the descent
            # at a transition that needs draws and resumes once they are injected
            from . import game as game_mod
            resume = 0
            while True:
                _lib.check(self.lib.ya_mcts_select_injected(
                    self.pool.ref, _lib.ptr(self.states), 1, _lib.ptr(self.players), self.sim_index, float(self.args.cpuct),
                    _lib.ptr(self.d_inj), resume, _lib.ptr(self.features), _lib.ptr(self.need_eval),
                    _lib.ptr(self.leaf_states), _lib.ptr(self.err_flag), s), "ya_mcts_select_injected")
                code = int(self.need_eval.item())
                if not code & 0x10:
                    break
                inj = np.zeros(12, dtype=np.uint8)
                if code & 1:
                    inj[0] = int(game_mod.tiebreak_uniform())
                    inj[11] |= 1
                if code & 2:
                    inj[1:6] = [int(x) for x in game_mod.roll_five()]
                    inj[6:11] = [int(x) for x in game_mod.roll_five()]
                    inj[11] |= 2
                self.d_inj.copy_(torch.from_numpy(inj))
                resume = 1
        self.sim_index += 1
        if code == 1:
            leaf = planes_to_boards(self.leaf_states.cpu().numpy().view(np.uint32))[0]
            pi, v = self.nnet.predict(leaf)                                   # MCTS.py:86
            pi = np.ascontiguousarray(np.asarray(pi, dtype=np.float32).reshape(ACTION_SIZE))
            self.pi_dev.copy_(torch.from_numpy(pi).unsqueeze(0))
            self.v_dev.fill_(float(np.float32(v)))
            _lib.check(self.lib.ya_mcts_expand(self.pool.ref, _lib.ptr(self.pi_dev), _lib.ptr(self.v_dev), 0, 0.0, 0.0, None,
                                               _lib.ptr(self.err_flag), s), "ya_mcts_expand")

    def getActionProb(self, canonicalBoard, temp=1):
        """MCTS.py:28-54."""
        self.root_index += 1
        self.sim_index = 0
        for _ in range(int(self.args.numMCTSSims)):
            self.search(canonicalBoard)
        e = int(self.err_flag.item())
        if e:
            raise _lib.YachtB200Error("MCTS kernel error %#x: %s" % (e, "; ".join(m for b, m in ERR_BITS.items() if e & b)))
        _lib.check(self.lib.ya_mcts_root_counts(
            self.pool.ref, _lib.ptr(self.states), 1, _lib.ptr(self.players), _lib.ptr(self.counts), _lib.ptr(self.visits),
            None, None, _lib.current_stream()), "ya_mcts_root_counts")
        counts = [int(x) for x in self.counts[0].cpu().numpy()]
        if temp == 0:
            bestAs = np.array(np.argwhere(counts == np.max(counts))).flatten()    # MCTS.py:45-49
            bestA = np.random.choice(bestAs)
            probs = [0] * len(counts)
            probs[bestA] = 1
            return probs
        counts = [x ** (1. / temp) for x in counts]                               # MCTS.py:51-54
        counts_sum = float(sum(counts))
        return [x / counts_sum for x in counts]

    def root_statistics(self):
        """(counts int32[3226], Ns, Q float64[3226], kind uint8[3226]) of the last searched root."""
        q = torch.zeros((1, ACTION_SIZE), dtype=torch.float64, device=self.device)
        kind = torch.zeros((1, ACTION_SIZE), dtype=torch.uint8, device=self.device)
        _lib.check(self.lib.ya_mcts_root_counts(
            self.pool.ref, _lib.ptr(self.states), 1, _lib.ptr(self.players), _lib.ptr(self.counts), _lib.ptr(self.visits),
            _lib.ptr(q), _lib.ptr(kind), _lib.current_stream()), "ya_mcts_root_counts")
        return self.counts[0].cpu().numpy(), int(self.visits.item()), q[0].cpu().numpy(), kind[0].cpu().numpy()

    def node_count(self):
        return int(self.pool.meta[0, 0].item())


def _arg(args, name, default):
    try:
        return args[name] if name in args else default
    except TypeError:
        return getattr(args, name, default)
