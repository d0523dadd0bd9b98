"""Parity of the CUDA environment kernels (through the C ABI) against the oracle and the golden
vectors produced by the reference.  Integer / byte work: bit-exact."""
import hashlib

import numpy as np
import pytest
import torch

from oracle import philox
from oracle import yacht_rules as yr

pytestmark = pytest.mark.gpu


def _engine(n, seed=0, game_base=0):
    from nypc_yacht_auction_b200.engine import BatchedYacht
    return BatchedYacht(n, seed=seed, game_base=game_base)


def _keys(env):
    from nypc_yacht_auction_b200.layout import string_key
    return [string_key(b) for b in env.boards()]


def _oracle_games(n, seed, game_base, episode=0):
    return [yr.new_game(philox.Draw(seed, game_base + g, episode, 0, philox.TAG_INIT)) for g in range(n)]


@pytest.mark.parametrize("n,seed,base", [(1, 0, 0), (13, 7, 3), (64, 0, 65530), (200, 2 ** 40 + 5, 10)])
def test_lockstep_random_games_match_oracle(n, seed, base):
    env = _engine(n, seed, base)
    boards = _oracle_games(n, seed, base)
    cur = [1] * n
    assert _keys(env) == [yr.key(b) for b in boards]
    for ply in range(48):
        masks = env.valid_moves().cpu().numpy()
        canon_dev = env.canonical()
        feats = env.features(canon_dev).cpu().numpy()
        acts = env.random_actions().cpu().numpy()
        ended = env.game_ended().cpu().numpy()
        from nypc_yacht_auction_b200.layout import planes_to_boards, string_key
        canon_keys = [string_key(b) for b in planes_to_boards(canon_dev.cpu().numpy().view(np.uint32))]
        exp_actions = []
        for g in range(n):
            canon = yr.canonical(boards[g], cur[g])
            assert canon_keys[g] == yr.key(canon)
            assert (masks[g] == yr.legal_mask(boards[g], cur[g])).all()
            assert feats[g].tobytes() == yr.features(canon).tobytes()
            assert ended[g] == np.float32(yr.outcome(boards[g], cur[g]))
            a = yr.random_legal_action(boards[g], cur[g], philox.Draw(seed, base + g, 0, ply, philox.TAG_ACTION))
            exp_actions.append(a)
            boards[g], cur[g] = yr.next_state(boards[g], cur[g], a, philox.Draw(seed, base + g, 0, ply, philox.TAG_REAL))
        assert acts.tolist() == exp_actions
        env.next_state(torch.from_numpy(acts))
        assert env.players.cpu().numpy().tolist() == cur
        assert _keys(env) == [yr.key(b) for b in boards]
    ended = env.game_ended().cpu().numpy()
    assert (ended != 0).all()
    assert ended.tolist() == [np.float32(yr.outcome(boards[g], cur[g])) for g in range(n)]


def test_fused_ply_matches_unfused_and_oracle():
    n, seed, base = 77, 11, 1000
    fused = _engine(n, seed, base)
    boards = _oracle_games(n, seed, base)
    cur = [1] * n
    ep = [0] * n
    ply = [0] * n
    masks = torch.empty((n, 3226), dtype=torch.uint8, device="cuda")
    for step in range(48 * 2 + 5):            # crosses two automatic re-deals
        exp_masks = np.stack([yr.legal_mask(boards[g], cur[g]) for g in range(n)])
        acts, outcome = fused.play_ply(masks=masks, auto_reset=True)
        assert (masks.cpu().numpy() == exp_masks).all(), step
        exp_a, exp_out = [], []
        for g in range(n):
            a = yr.random_legal_action(boards[g], cur[g], philox.Draw(seed, base + g, ep[g], ply[g], philox.TAG_ACTION))
            boards[g], cur[g] = yr.next_state(boards[g], cur[g], a, philox.Draw(seed, base + g, ep[g], ply[g], philox.TAG_REAL))
            ply[g] += 1
            exp_a.append(a)
            res = yr.outcome(boards[g], 1)
            exp_out.append(np.float32(res))
            if res != 0:
                ep[g] += 1
                ply[g] = 0
                cur[g] = 1
                boards[g] = yr.new_game(philox.Draw(seed, base + g, ep[g], 0, philox.TAG_INIT))
        assert acts.cpu().numpy().tolist() == exp_a
        assert outcome.cpu().numpy().tolist() == exp_out
        assert _keys(fused) == [yr.key(b) for b in boards]
        assert fused.players.cpu().numpy().tolist() == cur
    assert int(fused.err_flag.item()) == 0
    assert fused.episode.cpu().numpy().tolist() == ep


def test_golden_philox_traces(rules_golden):
    """The reference itself, driven by the injected Philox protocol (tests/golden/make_golden.py)."""
    meta, _ = rules_golden
    from nypc_yacht_auction_b200.layout import string_key
    for tr in meta["philox"]:
        env = _engine(1, tr["seed"], tr["game"])
        h = hashlib.sha256()
        actions = []
        for ply in range(48):
            canon = env.canonical()
            from nypc_yacht_auction_b200.layout import planes_to_boards
            ck = string_key(planes_to_boards(canon.cpu().numpy().view(np.uint32))[0])
            ones = torch.ones(1, dtype=torch.int8, device="cuda")
            mask = env.valid_moves(states=canon, players=ones).cpu().numpy()[0]
            h.update(ck.encode())
            h.update(mask.tobytes())
            a = env.random_actions()
            actions.append(int(a.item()))
            env.next_state(a)
        assert actions == tr["actions"]
        assert h.hexdigest() == tr["sha256"]
        b = env.boards()[0]
        assert string_key(b) == tr["final_key"]
        assert [b.p1.total_with_bonus(), b.p2.total_with_bonus()] == tr["totals"]
        ones = torch.ones(1, dtype=torch.int8, device="cuda")
        assert float(env.game_ended(players=ones).item()) == tr["ended_p1"]


def test_golden_seeded_traces_with_injected_draws(rules_golden):
    """MT19937-seeded reference games replayed on the GPU by injecting the recorded draws
    (draw_mode 1) and the recorded actions: every intermediate key and mask must match."""
    meta, arr = rules_golden
    from nypc_yacht_auction_b200.layout import string_key, planes_to_boards
    from nypc_yacht_auction_b200 import layout
    for tr in meta["seeded"]:
        env = _engine(1)
        a_roll, b_roll = tr["init_draws"][0][1], tr["init_draws"][1][1]
        start = layout.YachtBoard((1, sum(d << (3 * i) for i, d in enumerate(a_roll)) |
                                   (sum(d << (3 * i) for i, d in enumerate(b_roll)) << 15), 0, 0, 0, 0, 0, 0))
        env.load_boards([start], [1])
        feats = arr["features_seed%d" % tr["seed"]]
        for i, ply in enumerate(tr["plies"]):
            canon = env.canonical()
            assert string_key(planes_to_boards(canon.cpu().numpy().view(np.uint32))[0]) == ply["canon_key"]
            ones = torch.ones(1, dtype=torch.int8, device="cuda")
            mask = env.valid_moves(states=canon, players=ones).cpu().numpy()[0]
            exp = np.unpackbits(np.frombuffer(bytes.fromhex(ply["mask"]), dtype=np.uint8), bitorder="little")[:3226]
            assert (mask == exp).all()
            assert env.features(canon).cpu().numpy()[0].tobytes() == feats[i].tobytes()
            inj = np.zeros(12, dtype=np.uint8)
            rolls = [d[1] for d in ply["draws"] if d[0] == "roll"]
            ties = [d[1] for d in ply["draws"] if d[0] == "tie"]
            if ties:
                inj[0] = ties[0]
                inj[11] |= 1
            if rolls:
                inj[1:6] = rolls[0]
                inj[6:11] = rolls[1]
                inj[11] |= 2
            act = torch.tensor([ply["action"]], dtype=torch.int32)
            env.next_state(act, injected=torch.from_numpy(inj).cuda())
            assert _keys(env)[0] == ply["next_key"]
            assert int(env.players.item()) == ply["next_player"]
            assert float(env.game_ended().item()) == ply["ended"]
        assert _keys(env)[0] == tr["final_key"]


def test_injected_mode_reports_missing_draws():
    env = _engine(1, 3, 3)
    inj = torch.zeros(12, dtype=torch.uint8, device="cuda")
    env.next_state(torch.tensor([5], dtype=torch.int32), injected=inj)              # first bid: no draw
    before = _keys(env)[0]
    env.next_state(torch.tensor([5], dtype=torch.int32), injected=inj, check=False)  # same bid, round 1
    assert int(env.status.item()) == 0x300                                           # tie + rolls missing
    assert _keys(env)[0] == before


def test_error_statuses_match_reference_exceptions():
    env = _engine(4, 1, 1)
    with pytest.raises(ValueError):
        env.next_state(torch.tensor([202, 0, 0, 0], dtype=torch.int32))
    env = _engine(2, 1, 1)
    with pytest.raises(ValueError):
        env.next_state(torch.tensor([0, -1], dtype=torch.int32))
    # score phase: out-of-range raises, used category / unfit subset is a silent no-op
    from nypc_yacht_auction_b200 import layout

    class P:
        def __init__(self, carry, used):
            self.carry, self.used_mask, self.cat_scores, self.bid_score = carry, used, [0] * 12, 0

    class S:
        round_no, phase, rollA, rollB, p1_bid, p2_bid = 5, 1, [1, 2, 3, 4, 5], [6, 5, 4, 3, 2], None, None
        p1, p2 = P([1, 2, 3, 4, 5], 1), P([6, 6, 6, 6, 6, 1, 1, 1, 1, 1], 0)
    b = layout.pack_state(S())
    env = _engine(3)
    env.load_boards([b, b, b], [1, 1, 1])
    before = _keys(env)
    env.next_state(torch.tensor([202, 202 + 3 * 252 + 1, 202 + 2 * 252], dtype=torch.int32))
    after = _keys(env)
    assert after[0] == before[0] and after[1] == before[1] and after[2] != before[2]
    assert env.players.cpu().numpy().tolist() == [-1, -1, -1]
    env.load_boards([b, b, b], [1, 1, 1])
    with pytest.raises(ValueError):
        env.next_state(torch.tensor([5, 202, 202], dtype=torch.int32))
    # bid phase in round 13 cannot happen: RuntimeError (YachtGame.py:372)
    S.round_no, S.phase = 13, 0
    env.load_boards([layout.pack_state(S())] * 3, [1, 1, 1])
    with pytest.raises(RuntimeError):
        env.next_state(torch.tensor([202, 202, 202], dtype=torch.int32))


def test_enumerate_scores_matches_oracle(rules_golden):
    _, arr = rules_golden
    n = 40
    env = _engine(n, 5, 77)
    boards = _oracle_games(n, 5, 77)
    cur = [1] * n
    checked = 0
    for ply in range(48):
        table = env.enumerate_scores().cpu().numpy()
        for g in range(n):
            side = boards[g].sides[0 if cur[g] == 1 else 1]
            assert (table[g] == yr.score_table(side.dice)).all()
            checked += len(side.dice) == 10
        acts = env.random_actions()
        a = acts.cpu().numpy()
        for g in range(n):
            boards[g], cur[g] = yr.next_state(boards[g], cur[g], int(a[g]), philox.Draw(5, 77 + g, 0, ply, philox.TAG_REAL))
        env.next_state(acts)
    assert checked > 100


def test_score_kernel_exhaustive_table(rules_golden):
    """All 7776 five-dice tuples through the enumeration kernel (subset 0 of a 5-dice carry)
    against the reference's score_category table."""
    _, arr = rules_golden
    import itertools
    from nypc_yacht_auction_b200 import layout
    tuples = list(itertools.product(range(1, 7), repeat=5))
    boards = []
    for d in tuples:
        w2 = sum(v << (3 * i) for i, v in enumerate(d))
        boards.append(layout.YachtBoard((5 | (1 << 4), 0, w2, 0, 0, 0, 0, 0)))
    env = _engine(len(boards))
    env.load_boards(boards, [1] * len(boards))
    table = env.enumerate_scores().cpu().numpy()
    assert (table[:, :, 0] == arr["score_table_k"]).all()
    assert (table[:, :, 1:] == 0).all()
    assert hashlib.sha256((table[:, :, 0].astype(np.int32) * 1000).tobytes()).hexdigest() == \
        "ecb06dc09804d5a24e742b856b425a1782fe9aa7b1a087ba671b65d45ccc6e3a"


def test_full_size_properties():
    """65,536 games (BASELINE.json config 2): size-independent properties of the fused path."""
    n = 65536
    env = _engine(n, 123, 0)
    masks = torch.empty((n, 3226), dtype=torch.uint8, device="cuda")
    legal_counts = []
    for ply in range(48):
        acts, outcome = env.play_ply(masks=masks, auto_reset=False)
        cnt = masks.sum(dim=1, dtype=torch.int32)
        legal_counts.append(cnt)
        # the sampled action is legal
        assert bool(masks.gather(1, acts.long().unsqueeze(1)).all())
        if ply < 47:
            assert float(outcome.abs().sum()) == 0.0
    assert int(env.err_flag.item()) == 0
    # schedule of legal-move counts: 202 / (14-r)*252 / 1
    expect = [202, 202]
    for r in range(2, 13):
        expect += [202, 202, (14 - r) * 252, (14 - r) * 252]
    expect += [1, 1]
    for ply, cnt in enumerate(legal_counts):
        assert int(cnt.min()) == expect[ply] and int(cnt.max()) == expect[ply], ply
    assert bool((outcome != 0).all())
    boards = env.boards()[:50]
    for b, o in zip(boards, outcome[:50].cpu().numpy()):
        t1, t2 = b.p1.total_with_bonus(), b.p2.total_with_bonus()
        assert o == (np.float32(1e-4) if t1 == t2 else (1.0 if t1 > t2 else -1.0))
    # shard invariance: the same global games on a different base give the same results
    env2 = _engine(1000, 123, 40000)
    for ply in range(48):
        _, out2 = env2.play_ply(masks=None, auto_reset=False)
    assert torch.equal(out2, outcome[40000:41000])


def test_host_buffer_abi_matches_device_path():
    """ya_host_play_ply (host pointers, sliced over 4 copy/compute streams) gives exactly what the
    device-resident fused ply gives, including the masks, for a ragged batch size."""
    import ctypes
    from nypc_yacht_auction_b200 import _lib
    lib = _lib.load()
    n, seed, base = 1003, 9, 77
    dev_env = _engine(n, seed, base)
    dmask = torch.empty((n, 3226), dtype=torch.uint8, device="cuda")
    h = ctypes.c_void_p()
    _lib.check(lib.ya_host_create(n, 1, ctypes.byref(h)), "create")
    hs = dev_env.states.cpu().clone()
    hp = torch.ones(n, dtype=torch.int8)
    hply = torch.zeros(n, dtype=torch.int32)
    hep = torch.zeros(n, dtype=torch.int32)
    ha = torch.zeros(n, dtype=torch.int32)
    ho = torch.zeros(n, dtype=torch.float32)
    hm = torch.zeros((n, 3226), dtype=torch.uint8)
    herr = torch.zeros(1, dtype=torch.int32)
    for ply in range(50):
        acts, out = dev_env.play_ply(masks=dmask, auto_reset=True)
        _lib.check(lib.ya_host_play_ply(h, _lib.ptr(hs), _lib.ptr(hp), _lib.ptr(hply), _lib.ptr(hep), _lib.ptr(ha), _lib.ptr(ho),
                                        _lib.ptr(hm), _lib.ptr(herr), seed, base, 1), "host ply")
        assert torch.equal(hs, dev_env.states.cpu()) and torch.equal(hp, dev_env.players.cpu())
        assert torch.equal(ha, acts.cpu()) and torch.equal(ho, out.cpu()) and torch.equal(hm, dmask.cpu())
        assert torch.equal(hply, dev_env.ply.cpu()) and torch.equal(hep, dev_env.episode.cpu())
    assert int(herr.item()) == 0
    lib.ya_host_destroy(h)


def test_record_host_abi_matches_device_path():
    """ya_host_play_ply_records: 64-byte records, 4 pipelined slices -- same boards / actions / outcomes /
    masks as the device-resident fused ply (ragged batch so slices and superblocks are uneven)."""
    import ctypes
    from nypc_yacht_auction_b200 import _lib
    lib = _lib.load()
    n, seed, base = 1003, 21, 300
    dev_env = _engine(n, seed, base)
    dmask = torch.empty((n, 3226), dtype=torch.uint8, device="cuda")
    h = ctypes.c_void_p()
    _lib.check(lib.ya_host_create(n, 1, ctypes.byref(h)), "create")
    rec = torch.zeros((n, 16), dtype=torch.int32).pin_memory()
    st = dev_env.states.cpu()
    rec[:, 0:4] = st[0]
    rec[:, 4:8] = st[1]
    rec[:, 10] = 1
    hm = torch.zeros((n, 3226), dtype=torch.uint8).pin_memory()
    herr = torch.zeros(1, dtype=torch.int32)
    for ply in range(50):
        acts, out = dev_env.play_ply(masks=dmask, auto_reset=True)
        _lib.check(lib.ya_host_play_ply_records(h, _lib.ptr(rec), _lib.ptr(hm), _lib.ptr(herr), seed, base, 1), "records ply")
        st = dev_env.states.cpu()
        assert torch.equal(rec[:, 0:4], st[0]) and torch.equal(rec[:, 4:8], st[1])
        assert torch.equal(rec[:, 8], dev_env.episode.cpu()) and torch.equal(rec[:, 9], dev_env.ply.cpu())
        assert torch.equal(rec[:, 10], dev_env.players.cpu().int()) and torch.equal(rec[:, 11], acts.cpu())
        assert torch.equal(rec[:, 12].view(torch.float32), out.cpu())
        assert torch.equal(hm, dmask.cpu())
    assert int(herr.item()) == 0
    assert (rec[:, 13] == 1).all() and (rec[:, 14] + rec[:, 15] <= 1).all()      # one finished game per slot so far
    # a multi-ply call (46 more plies = the rest of episode 1) equals 46 single-ply calls on the device path
    _lib.check(lib.ya_host_play_plies_records(h, _lib.ptr(rec), 46, None, _lib.ptr(herr), seed, base, 1), "records plies")
    wins = torch.zeros(2, dtype=torch.int64)
    for ply in range(46):
        acts, out = dev_env.play_ply(masks=None, auto_reset=True)
    st = dev_env.states.cpu()
    assert torch.equal(rec[:, 0:4], st[0]) and torch.equal(rec[:, 4:8], st[1]) and torch.equal(rec[:, 8], dev_env.episode.cpu())
    assert (rec[:, 13] == 2).all() and int(herr.item()) == 0
    lib.ya_host_destroy(h)


def test_search_style_walks_match_oracle():
    """MCTS-style traversal (MCTS.py:149-150: getNextState(canonical, 1, a) then getCanonicalForm) from
    mid-game roots, including the quirk states real play never reaches (rounds that do not advance, five- and
    zero-dice positions, bogus round-12 terminals), with in-search Philox draws (tag SEARCH, depth, sim)."""
    from nypc_yacht_auction_b200.engine import BatchedYacht, TAG_SEARCH
    from nypc_yacht_auction_b200 import _lib
    from nypc_yacht_auction_b200.layout import string_key
    from conftest import to_oracle_board
    rng = np.random.default_rng(5)
    n, seed, base = 160, 77, 4000
    env = BatchedYacht(n, seed=seed, game_base=base)
    start_ply = rng.integers(0, 47, size=n)
    for t in range(47):                                   # bring game g to a random ply
        acts = env.random_actions()
        keep = torch.from_numpy((start_ply > t)).cuda()
        st_before, pl_before, ply_before = env.states.clone(), env.players.clone(), env.ply.clone()
        env.next_state(acts)
        env.states[:, ~keep] = st_before[:, ~keep]
        env.players[~keep] = pl_before[~keep]
        env.ply[~keep] = ply_before[~keep]
    canon = env.canonical()
    env.states.copy_(canon)
    env.players.fill_(1)
    boards = [to_oracle_board(b) for b in env.boards()]
    lib = env.lib
    ended_seen = dead_seen = 0
    for depth in range(7):
        masks = env.valid_moves().cpu().numpy()
        ended = env.game_ended().cpu().numpy()
        acts = np.zeros(n, dtype=np.int32)
        live = np.zeros(n, dtype=bool)
        for g in range(n):
            assert (masks[g] == yr.legal_mask(boards[g], 1)).all()
            assert ended[g] == np.float32(yr.outcome(boards[g], 1))
            legal = np.flatnonzero(masks[g])
            if ended[g] != 0:
                ended_seen += 1
            elif len(legal) == 0:
                dead_seen += 1
            else:
                acts[g] = int(legal[rng.integers(0, len(legal))])
                live[g] = True
        depth_sim = torch.full((n,), depth | (3 << 8), dtype=torch.int32, device="cuda")       # depth, sim = 3
        a_dev = torch.from_numpy(acts).cuda()
        out = torch.empty_like(env.states)
        nxt = torch.empty_like(env.players)
        _lib.check(lib.ya_next_state(_lib.ptr(env.states), n, _lib.ptr(env.players), _lib.ptr(a_dev), _lib.ptr(out), n,
                                     _lib.ptr(nxt), _lib.ptr(env.status), n, 0, None, seed, base, _lib.ptr(env.episode),
                                     _lib.ptr(env.ply), TAG_SEARCH, _lib.ptr(depth_sim), _lib.current_stream()), "next")
        live_t = torch.from_numpy(live).cuda()
        env.states[:, live_t] = out[:, live_t]
        env.players[live_t] = nxt[live_t]
        env.states.copy_(env.canonical())
        env.players.fill_(1)
        ply_host = env.ply.cpu().numpy()
        for g in range(n):
            if live[g]:
                d = philox.Draw(seed, base + g, 0, int(ply_host[g]), philox.TAG_SEARCH, depth, 3)
                nb, who = yr.next_state(boards[g], 1, int(acts[g]), d)
                boards[g] = yr.canonical(nb, who)
        assert _keys(env) == [yr.key(b) for b in boards], depth
    assert dead_seen > 0 and ended_seen > 0        # the walk did reach dead ends and (bogus) terminals
