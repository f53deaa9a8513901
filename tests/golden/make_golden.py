#!/usr/bin/env python
"""Generates the committed golden fixtures by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Outputs (all under tests/golden/, all small, all committed):

  notebook_games.json  the five rendered 2-player games recorded in
                       experiments/simple_tournament.ipynb (raw lines 809-2534), parsed
                       into deal / per-turn actions / per-turn boards, hands, scores.
  env_traces.npz       random-legal-play traces through the reference SechsNimmtEnv
                       for P = 2..10: deal, actions, rewards, done, hands, boards and the
                       full observation vectors after reset and after every step
                       (env.py:43-77, 174-212), plus include_summaries=False observations.
  kat.json             KAT-A / KAT-B of SURVEY.md §4 regenerated from the reference,
                       directed edge cases, illegal-move behaviour.
  mcs_exact.json       exact E[outcome | first card] by exhaustive enumeration through
                       the reference env for small roots (KAT-C and friends), and
                       reference-MCSAgent Monte-Carlo estimates for z-tests.
  policy_rollouts.npz  reference PolicyMCSAgent rollouts with a sharpened policy: per-first-card outcome
                       mean / variance / count, the root policy, the weights, the root.
  policy_vectors.npz   MultiHeadedMLP(48,(100,100),(1,)) weights (torch.manual_seed(0)),
                       input rows, SechsNimmtStateNormalization outputs, softmax probs,
                       and PUCTAgent._compute_pucts / _normalize_q vectors.

The reference is only ever *called*; none of its source is copied.
"""
import itertools
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_loader import REF_ROOT, load_reference  # noqa: E402

ref = load_reference()
Env = ref.env.SechsNimmtEnv


# ----------------------------------------------------------------------------------------
# 1. notebook games
# ----------------------------------------------------------------------------------------
CARD_RE = re.compile(r"(\d+)[ .:+#]?")


def _cards(text):
    return [int(m) - 1 for m in CARD_RE.findall(text)]


def parse_notebook_games():
    nb = json.load(open(os.path.join(REF_ROOT, "experiments", "simple_tournament.ipynb")))
    games = []
    for cell in nb["cells"]:
        if cell.get("cell_type") != "code":
            continue
        lines = []
        for out in cell.get("outputs", []):
            if out.get("name") == "stderr":
                lines += "".join(out["text"]).split("\n")
        if not any(l.startswith("Dealing cards") for l in lines):
            continue
        # split into render blocks and the "plays card" lines between them
        snapshots, plays, events = [], [], []
        cur_plays, cur_events = [], []
        i = 0
        names = {}
        while i < len(lines):
            l = lines[i]
            if l.startswith("Board:"):
                board = []
                i += 1
                while not lines[i].startswith("Players:"):
                    board.append(_cards(lines[i].replace("_", " ").replace("*", " ")))
                    i += 1
                i += 1
                hands, scores = [], []
                while lines[i].startswith("  ") and "Hornochsen" in lines[i]:
                    m = re.match(r"\s+(.*?)\s*\(player (\d+)\):\s+(-?\d+) Hornochsen, (.*)$", lines[i])
                    names[m.group(1).strip()] = int(m.group(2)) - 1
                    scores.append(int(m.group(3)))
                    rest = m.group(4)
                    hands.append([] if rest.startswith("no cards") else _cards(rest[len("cards"):]))
                    i += 1
                snapshots.append({"board": board, "hands": hands, "scores": scores})
                if len(snapshots) > 1:
                    plays.append(cur_plays)
                    events.append(cur_events)
                cur_plays, cur_events = [], []
                continue
            m = re.match(r"(.*?)\s*\(player (\d+)\) plays card (\d+)", l)
            if m:
                cur_plays.append([int(m.group(2)) - 1, int(m.group(3)) - 1])
            m = re.match(r"\s+\.\.\.chooses to replace row (\d+)", l)
            if m:
                cur_events.append(["replace", int(m.group(1)) - 1])
            m = re.match(r"\s+\.\.\.and gains (\d+) Hornochsen", l)
            if m:
                cur_events.append(["gains", int(m.group(1))])
            i += 1
        P = len(snapshots[0]["hands"])
        actions = []
        for turn in plays:
            a = [None] * P
            for p, c in turn:
                a[p] = c
            actions.append(a)
        games.append({"num_players": P, "snapshots": snapshots, "actions": actions, "events": events})
    return games


def replay_check_notebook(games):
    """The recorded games must replay exactly through today's reference env."""
    n_pen = 0
    for g in games:
        s0 = g["snapshots"][0]
        env = Env(g["num_players"], verbose=False)
        env.reset_to([list(r) for r in s0["board"]], [list(h) for h in s0["hands"]])
        for t, a in enumerate(g["actions"]):
            _, rew, done, _ = env.step(a)
            snap = g["snapshots"][t + 1]
            assert [list(map(int, r)) for r in env._board] == snap["board"], (t, env._board, snap["board"])
            assert [list(map(int, h)) for h in env._hands] == snap["hands"]
            assert list(map(int, env._scores)) == snap["scores"]
            n_pen += int(np.count_nonzero(rew))
        assert done
    return n_pen


# ----------------------------------------------------------------------------------------
# 2. random-play traces
# ----------------------------------------------------------------------------------------
def pad_hand(h):
    return list(map(int, h)) + [-1] * (10 - len(h))


def board_array(board):
    out = -np.ones((4, 6), dtype=np.int8)
    for r, cards in enumerate(board):
        for i, c in enumerate(cards):
            out[r, i] = c
    return out


def record_traces(P, n_games, seed):
    np.random.seed(seed)
    hamster = ref.random.DrunkHamster()
    env = Env(P, verbose=False)
    env_ns = Env(P, include_summaries=False, verbose=False)
    deal_hands = np.zeros((n_games, P, 10), np.int8)
    deal_rows = np.zeros((n_games, 4), np.int8)
    actions = np.zeros((n_games, 10, P), np.int8)
    rewards = np.zeros((n_games, 10, P), np.int8)
    done = np.zeros((n_games, 10), np.uint8)
    hands = -np.ones((n_games, 11, P, 10), np.int8)
    boards = -np.ones((n_games, 11, 4, 6), np.int8)
    scores = np.zeros((n_games, 11, P), np.int16)
    obs = np.zeros((n_games, 11, P, 47), np.int8)
    obs_ns = np.zeros((n_games, 11, P, 35), np.int8)
    for g in range(n_games):
        states, legal = env.reset()
        deal_hands[g] = np.array(env._hands)
        deal_rows[g] = [r[0] for r in env._board]
        env_ns.reset_to([list(r) for r in env._board], [list(h) for h in env._hands])
        st_ns, _ = env_ns._create_states()
        for t in range(11):
            assert all(s.dtype == np.int64 and s.shape == (47,) for s in states)
            obs[g, t] = np.array(states)
            obs_ns[g, t] = np.array(st_ns)
            hands[g, t] = np.array([pad_hand(h) for h in env._hands])
            boards[g, t] = board_array(env._board)
            scores[g, t] = env._scores
            for p in range(P):
                assert list(map(int, legal[p])) == list(map(int, env._hands[p]))
            if t == 10:
                break
            a = [int(hamster(states[p], legal_actions=legal[p])[0]) for p in range(P)]
            actions[g, t] = a
            (states, legal), rew, d, _ = env.step(a)
            (st_ns, _), rew2, d2, _ = env_ns.step(a)
            assert (rew == rew2).all() and d == d2
            assert rew.dtype == np.int32
            rewards[g, t] = rew
            done[g, t] = d
        assert d
    return dict(deal_hands=deal_hands, deal_rows=deal_rows, actions=actions, rewards=rewards, done=done,
                hands=hands, boards=boards, scores=scores, obs=obs, obs_ns=obs_ns)


# ----------------------------------------------------------------------------------------
# 3. KATs and edge cases
# ----------------------------------------------------------------------------------------
def index_policy_game(env, P):
    rewards = []
    for t in range(10):
        a = [int(env._hands[p][(t * (p + 1)) % len(env._hands[p])]) for p in range(P)]
        _, rew, done, _ = env.step(a)
        rewards.append(list(map(int, rew)))
    assert done
    return rewards


def make_kats():
    kat = {}
    # KAT-A
    np.random.seed(0)
    env = Env(4, verbose=False)
    env.reset()
    a = {"hands": [list(map(int, h)) for h in env._hands], "rows": [list(map(int, r)) for r in env._board]}
    a["rewards"] = index_policy_game(env, 4)
    a["totals"] = list(map(int, -env._scores))
    a["final_rows"] = [list(map(int, r)) for r in env._board]
    kat["A"] = a
    # KAT-B
    b = {}
    for P in (2, 4, 10):
        np.random.seed(1000 + P)
        env = Env(P, verbose=False)
        tot = np.zeros(P, np.int64)
        n_events = 0
        deals = []
        for g in range(2000):
            env.reset()
            deals.append([[list(map(int, h)) for h in env._hands], [int(r[0]) for r in env._board]])
            for rew in index_policy_game(env, P):
                tot += np.array(rew)
                n_events += sum(1 for x in rew if x != 0)
        b[str(P)] = {"score_sums": list(map(int, tot)), "events": n_events}
        # the deals are needed to replay KAT-B without the reference's RNG: store compactly
        np.save(os.path.join(HERE, f"katb_deals_p{P}.npy"),
                np.array([sum(d[0], []) + d[1] for d in deals], dtype=np.int8))
    kat["B"] = b

    # Directed edge cases, each run through the reference env
    edge = []

    def run_case(name, board, hands, acts):
        P = len(hands)
        env = Env(P, verbose=False)
        env.reset_to([list(r) for r in board], [list(h) for h in hands])
        steps = []
        for a in acts:
            _, rew, done, _ = env.step(list(a))
            steps.append({"actions": list(a), "rewards": list(map(int, rew)), "done": bool(done),
                          "board": [list(map(int, r)) for r in env._board],
                          "hands": [list(map(int, h)) for h in env._hands],
                          "scores": list(map(int, env._scores))})
        edge.append({"name": name, "board": board, "hands": hands, "steps": steps})

    # undercut picks the lowest-sum row, lowest index on ties (rows valued 10,1,1,1 -> row 1)
    run_case("undercut_tie_lowest_index", [[54, 65], [20], [30], [40]], [[3, 50], [100, 101]], [[3, 100]])
    # sixth card takes five, leaves the new card
    run_case("sixth_card", [[10, 11, 12, 13, 14], [30], [50], [70]], [[15, 90], [91, 92]], [[15, 91]])
    # undercut penalty excludes the played card (card 54 = value 7 played as undercut)
    run_case("undercut_excludes_played", [[60], [70], [80], [90]], [[54, 100], [101, 102]], [[54, 101]])
    # a later card in the same step sees the row left by an earlier one
    run_case("later_card_sees_new_row", [[60, 61, 62, 63, 64], [20], [30], [40]], [[65, 1], [66, 2]], [[65, 66]])
    # two undercuts in one step: the second sees the first one's replaced row
    run_case("double_undercut", [[50], [60], [70], [80]], [[1, 99], [2, 98]], [[1, 2]])
    # undercut takes a full 5-row (both conditions at once)
    run_case("undercut_full_row", [[90, 91, 92, 93, 94], [95, 96], [97, 98], [99, 100]], [[0, 5], [101, 102]], [[0, 101]])
    # all four rows full, 10 players: chain of sixth-card takes
    run_case("p10_chain", [[0, 1, 2, 3, 4], [20, 21, 22, 23, 24], [40, 41, 42, 43, 44], [60, 61, 62, 63, 64]],
             [[5 + i, 100 - i] for i in range(5)] + [[25 + i, 90 - i] for i in range(5)], [[5, 6, 7, 8, 9, 25, 26, 27, 28, 29]])
    # value-55 card (index 54) swept in a sixth-card take
    run_case("value_55", [[50, 51, 52, 53, 54], [10], [20], [30]], [[55, 99], [100, 101]], [[55, 100]])
    # equal row sums everywhere, undercut -> row 0
    run_case("undercut_all_equal", [[10], [20], [30], [40]], [[0, 50], [60, 61]], [[0, 60]])
    kat["edge"] = edge

    # illegal moves: raises before any mutation (env.py:68-69)
    env = Env(2, verbose=False)
    env.reset_to([[10], [20], [30], [40]], [[1, 2], [3, 4]])
    ill = {}
    try:
        env.step([1, 5])
        ill["raised"] = False
    except ref.env.InvalidMoveException as e:
        ill["raised"] = True
        ill["message"] = str(e)
    ill["board_after"] = [list(map(int, r)) for r in env._board]
    ill["hands_after"] = [list(map(int, h)) for h in env._hands]
    try:
        env.step([1])
        ill["short_asserts"] = False
    except AssertionError:
        ill["short_asserts"] = True
    try:
        Env(11)
        ill["p11_asserts"] = False
    except AssertionError:
        ill["p11_asserts"] = True
    e10 = Env(10, verbose=False)
    np.random.seed(5)
    e10.reset()
    ill["p10_uses_all_cards"] = sorted(sum([list(map(int, h)) for h in e10._hands], []) + [int(r[0]) for r in e10._board]) == list(range(104))
    kat["illegal"] = ill

    kat["card_values"] = [int(Env._card_value(c)) for c in range(104)]
    return kat


# ----------------------------------------------------------------------------------------
# 4. MCS exact enumeration + reference Monte-Carlo
# ----------------------------------------------------------------------------------------
def exact_mcs(board, own, available, P):
    """E[outcome | first card] for the reference's rollout distribution, by enumeration.

    Rollout law (agents/mcts.py:116-154): opponents hold a uniformly random ordered
    partition of a uniform sample of the available cards; every player (player 0 too)
    plays uniformly at random.  Enumerate opponent hands x all play orders.
    """
    n = len(own)
    totals = {a: [0, 0, 0] for a in own}  # sum, sumsq, count (weights equal per first card)
    opp_slots = (P - 1) * n
    for opp_cards in itertools.permutations(available, opp_slots) if opp_slots <= 2 else _opp_iter(available, P, n):
        opp_hands = [sorted(opp_cards[i * n:(i + 1) * n]) for i in range(P - 1)]
        for own_order in itertools.permutations(own):
            for opp_orders in itertools.product(*[itertools.permutations(h) for h in opp_hands]):
                env = Env(P, verbose=False)
                env.reset_to([list(r) for r in board], [list(own)] + [list(h) for h in opp_hands])
                out = 0
                for t in range(n):
                    a = [own_order[t]] + [o[t] for o in opp_orders]
                    _, rew, done, _ = env.step(a)
                    out += int(rew[0])
                tt = totals[own_order[0]]
                tt[0] += out
                tt[1] += out * out
                tt[2] += 1
    return {str(a): {"sum": t[0], "sumsq": t[1], "count": t[2], "mean": t[0] / t[2]} for a, t in totals.items()}


def _opp_iter(available, P, n):
    # unordered disjoint hands (order of opponents matters, order within a hand does not)
    def rec(rem, k):
        if k == 0:
            yield ()
            return
        for h in itertools.combinations(rem, n):
            rest = [c for c in rem if c not in h]
            for tail in rec(rest, k - 1):
                yield h + tail
    return rec(list(available), P - 1)


def reference_mcs_estimate(state, legal, n_rollouts, seed):
    """Runs the reference's own _draw_env/_play_out loop (agents/mcts.py:97-101) n times."""
    import torch
    np.random.seed(seed)
    agent = ref.mcts.MCSAgent(mc_max=n_rollouts, mc_per_card=n_rollouts)
    agent._initialize_game(torch.tensor(state, dtype=torch.float))
    agent._memorize_cards(torch.tensor(state, dtype=torch.float), list(legal))
    outcomes = {a: [] for a in legal}
    for _ in range(n_rollouts):
        env = agent._draw_env(list(legal), torch.tensor(state, dtype=torch.float))
        a, _, out = agent._play_out(env, outcomes)
        outcomes[int(a)].append(float(out))
    return {str(a): {"mean": float(np.mean(o)), "var": float(np.var(o)), "count": len(o)} for a, o in outcomes.items()}, \
        sorted(map(int, agent.available_cards))


def make_mcs(games):
    out = {}
    # KAT-C: notebook game 1 after turn 8 (0-based) => own hand of 2, P=2
    g = games[0]
    snap = g["snapshots"][8]
    board, own = snap["board"], snap["hands"][0]
    visible = set(own) | set(sum(board, []))
    # SURVEY KAT-C uses "all 85 cards not visible" as the available set
    avail = [c for c in range(104) if c not in visible]
    out["C"] = {"board": board, "own": own, "P": 2, "available": avail, "exact": exact_mcs(board, own, avail, 2)}
    # a 3-player, 2-card root with a small available set (keeps enumeration cheap)
    board2 = [[12, 33, 54], [60, 61, 62, 63, 70], [80], [5, 9]]
    own2 = [8, 71]
    avail2 = [2, 6, 10, 34, 55, 64, 72, 81, 90, 100, 103, 45]
    out["D"] = {"board": board2, "own": own2, "P": 3, "available": avail2, "exact": exact_mcs(board2, own2, avail2, 3)}
    # 2-player 3-card root
    board3 = [[20, 21, 22, 23, 24], [40, 43], [65], [88, 89, 98]]
    own3 = [25, 41, 99]
    avail3 = [0, 10, 26, 42, 44, 54, 66, 87, 90, 100, 101, 102]
    out["E"] = {"board": board3, "own": own3, "P": 2, "available": avail3, "exact": exact_mcs(board3, own3, avail3, 2)}

    # Reference Monte-Carlo on a 4-player opening position (KAT-A deal): z-test target
    np.random.seed(0)
    env = Env(4, verbose=False)
    states, legal = env.reset()
    est, avail = reference_mcs_estimate(np.array(states[0], dtype=np.float32), list(map(int, legal[0])), 6000, seed=123)
    out["MC4"] = {"state": list(map(int, states[0])), "legal": list(map(int, legal[0])), "available": avail,
                  "P": 4, "reference_mc": est, "rollouts": 6000}
    # ... and a mid-game one with stale card memory: play 4 turns with the agent watching
    np.random.seed(7)
    env = Env(3, verbose=False)
    states, legal = env.reset()
    import torch
    agent = ref.mcts.MCSAgent(mc_max=1, mc_per_card=1)
    for t in range(4):
        # the card-memory half of BaseMCAgent.forward (agents/mcts.py:47-49); the search itself is skipped
        st = torch.tensor(states[0], dtype=torch.float)
        if len(legal[0]) == agent.handsize:
            agent._initialize_game(st)
        agent._memorize_cards(st, list(map(int, legal[0])))
        a = [int(l[(3 * t + p) % len(l)]) for p, l in enumerate(legal)]
        (states, legal), _, _, _ = env.step(a)
    est, avail = reference_mcs_estimate(np.array(states[0], dtype=np.float32), list(map(int, legal[0])), 6000, seed=321)
    # reference_mcs_estimate re-initialises the memory; overwrite with the stale-memory set the agent really holds
    agent._memorize_cards(torch.tensor(states[0], dtype=torch.float), list(map(int, legal[0])))
    out["MC3"] = {"state": list(map(int, states[0])), "legal": list(map(int, legal[0])), "available": avail,
                  "agent_available_after_watching": sorted(map(int, agent.available_cards)),
                  "P": 3, "reference_mc": est, "rollouts": 6000}
    return out


# ----------------------------------------------------------------------------------------
# 5. Alpha0.5 policy vectors
# ----------------------------------------------------------------------------------------
def make_policy_vectors():
    import torch
    torch.manual_seed(0)
    agent = ref.mcts.PUCTAgent(mc_max=200)
    sd = {k: v.detach().numpy().copy() for k, v in agent.state_dict().items()}
    np.random.seed(11)
    env = Env(4, verbose=False)
    rows_in, rows_norm, logits, probs_all, seg = [], [], [], [], []
    for g in range(6):
        states, legal = env.reset()
        done = False
        while not done:
            for p in range(4):
                st = torch.tensor(states[p]).to(torch.float)
                la = list(map(int, legal[p]))
                batch = torch.cat([torch.cat((torch.tensor([float(a)]), st)).unsqueeze(0) for a in la], dim=0)
                norm = agent.preprocessor(batch)
                (lg,) = agent.actor(norm)
                pr = agent._compute_policy(la, st)
                rows_in.append(batch.numpy())
                rows_norm.append(norm.detach().numpy())
                logits.append(lg.detach().numpy().reshape(-1))
                probs_all.append(pr.detach().numpy())
                seg.append(len(la))
            a = [int(np.random.choice(l)) for l in legal]
            (states, legal), _, done, _ = env.step(a)
    out = {"w_" + k.replace(".", "_"): v for k, v in sd.items()}
    out["rows_in"] = np.concatenate(rows_in).astype(np.float32)
    out["rows_norm"] = np.concatenate(rows_norm).astype(np.float32)
    out["logits"] = np.concatenate(logits).astype(np.float32)
    out["probs"] = np.concatenate(probs_all).astype(np.float32)
    out["seg"] = np.array(seg, np.int32)

    # PUCT root rule vectors (agents/mcts.py:295-315)
    rng = np.random.RandomState(3)
    cases = []
    import warnings
    for case in range(40):
        n = int(rng.randint(2, 11))
        la = sorted(rng.choice(104, n, replace=False).tolist())
        n_out = int(rng.choice([0, 3, 9, 10, 11, 50, 200]))
        outcomes = {a: [] for a in la}
        for _ in range(n_out):
            outcomes[la[int(rng.randint(n))]].append(float(-rng.randint(0, 15)))
        if case % 8 == 7 and n_out >= 10:  # all-equal outcomes => 0/0 => NaN => choice 0
            outcomes = {a: [-3.0] * len(o) for a, o in outcomes.items()}
        pr = rng.dirichlet(np.ones(n)).astype(np.float32)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            pucts = agent._compute_pucts(la, outcomes, torch.tensor(pr))
            mx, mn, md = agent._normalize_q(outcomes)
        best, choice = -float("inf"), 0
        for i, pu in enumerate(pucts):
            if pu > best:
                best, choice = pu, i
        cases.append({"legal": la, "outcomes": {str(a): o for a, o in outcomes.items()}, "probs": pr.tolist(),
                      "pucts": [None if np.isnan(x) else float(x) for x in pucts],
                      "norm": [float(mx), float(mn), float(md)], "choice": choice})
    return out, cases


def make_policy_rollouts(n_rollouts=2500):
    """Reference PolicyMCSAgent rollouts (agents/mcts.py:91-154, 209-228) with a SHARPENED policy
    (head weights x 20, so that the policy is far from uniform and policy bugs would show) from a
    4-player mid-game root: per first card the mean / variance / count of the outcomes."""
    import torch
    torch.manual_seed(0)
    agent = ref.mcts.PolicyMCSAgent(mc_max=n_rollouts, mc_per_card=n_rollouts)
    with torch.no_grad():
        agent.actor.head_nets[0][0].weight *= 20.0
    np.random.seed(21)
    env = Env(4, verbose=False)
    states, legal = env.reset()
    for t in range(4):
        (states, legal), _, _, _ = env.step([int(np.random.choice(l)) for l in legal])
    state = torch.tensor(states[0], dtype=torch.float)
    legal0 = list(map(int, legal[0]))
    agent._initialize_game(state)
    agent._memorize_cards(state, legal0)
    root_probs = agent._compute_policy(legal0, state).detach().numpy()
    np.random.seed(22)
    torch.manual_seed(23)
    outcomes = {a: [] for a in legal0}
    for _ in range(n_rollouts):
        e = agent._draw_env(legal0, state)
        a, _, out = agent._play_out(e, outcomes)
        outcomes[int(a)].append(float(out))
    sd = {k: v.detach().numpy().copy() for k, v in agent.state_dict().items()}
    out = {"w_" + k.replace(".", "_"): v for k, v in sd.items()}
    out["state"] = np.array(states[0], np.int64)
    out["legal"] = np.array(legal0, np.int64)
    out["available"] = np.array(sorted(map(int, agent.available_cards)), np.int64)
    out["root_probs"] = root_probs.astype(np.float32)
    out["mean"] = np.array([np.mean(outcomes[a]) if outcomes[a] else np.nan for a in legal0])
    out["var"] = np.array([np.var(outcomes[a]) if outcomes[a] else np.nan for a in legal0])
    out["count"] = np.array([len(outcomes[a]) for a in legal0], np.int64)
    return out


def make_policy_train(n_episodes=3):
    """Reference PolicyMCSAgent.learn / _train (agents/mcts.py:230-261, agents/base.py:29-33): one agent, n_episodes
    four-player games; at every turn the agent's (state, legal cards) and a chosen card are recorded together with the
    log-probability the reference stores for it (Categorical(_compute_policy).log_prob, :212-215; constant 0 for the single
    last card, :52-53); after each episode the reference's own _train() runs (loss = -sum log_prob, Adam defaults).
    Saved: initial weights, the decisions, the loss of every episode, the weights after every episode."""
    import torch
    from torch.distributions import Categorical
    torch.manual_seed(5)
    agent = ref.mcts.PolicyMCSAgent(mc_max=10)
    agent.train()                                   # creates Adam with default kwargs (base.py:29-33)
    out = {"w0_" + k.replace(".", "_"): v.detach().numpy().copy() for k, v in agent.state_dict().items()}
    np.random.seed(31)
    env = Env(4, verbose=False)
    obs, legal_n, chosen, losses, ref_logp = [], [], [], [], []
    for ep in range(n_episodes):
        states, legal = env.reset()
        done = False
        while not done:
            st = torch.tensor(states[0]).to(torch.float)
            la = list(map(int, legal[0]))
            idx = int(np.random.randint(len(la)))
            if len(la) == 1:
                lp = torch.tensor(0.0)
            else:
                lp = Categorical(agent._compute_policy(la, st)).log_prob(torch.tensor(idx))
            obs.append(np.array(states[0], np.int64)); legal_n.append(len(la)); chosen.append(idx); ref_logp.append(float(lp.detach()))
            acts = [la[idx]] + [int(np.random.choice(l)) for l in legal[1:]]
            (states, legal), rewards, done, _ = env.step(acts)
            agent.history.store(log_prob=lp, reward=float(rewards[0]) * agent.r_factor)
        losses.append(agent._train())
        agent.history.clear()
        for k, v in agent.state_dict().items():
            out[f"w{ep + 1}_" + k.replace(".", "_")] = v.detach().numpy().copy()
    out["obs"] = np.array(obs, np.int8)
    out["n_legal"] = np.array(legal_n, np.int32)
    out["chosen"] = np.array(chosen, np.int32)
    out["log_prob"] = np.array(ref_logp, np.float64)
    out["loss"] = np.array(losses, np.float64)
    return out


def make_position_stats(n=400):
    """Tournament._compute_absolute_positions / _compute_relative_positions (tournament.py:240-256) and the winner rule
    (np.argmax, :141) of the unmodified reference on random score vectors with many ties, P = 2..10."""
    from ref_loader import load_tournament
    T = load_tournament().Tournament
    rng = np.random.RandomState(17)
    cases = []
    for _ in range(n):
        P = int(rng.randint(2, 11))
        scores = -rng.randint(0, 12, size=P).astype(np.int32)          # negative Hornochsen totals, ties are common
        cases.append({"scores": scores.tolist(), "absolute": [float(x) for x in T._compute_absolute_positions(scores)],
                      "relative": [float(x) for x in T._compute_relative_positions(scores)], "winner": int(np.argmax(scores))})
    return cases


def make_masked_policy(n_games=3):
    """MaskedReinforceAgent.forward (agents/policy.py:45-60): state -> SechsNimmtStateNormalization(action=False) ->
    MultiHeadedMLP(47, (100, 100), (104,)) -> logits of the legal cards -> softmax.  Weights, states, legal cards, the
    normalised states and the probabilities over the legal cards for every seat and turn of a few random games."""
    import torch
    from ref_loader import load_policy_agents
    pol = load_policy_agents()
    torch.manual_seed(9)
    agent = pol.MaskedReinforceAgent()
    out = {"w_" + k.replace(".", "_"): v.detach().numpy().copy() for k, v in agent.state_dict().items()}
    np.random.seed(41)
    env = Env(4, verbose=False)
    states_all, norm_all, probs_all, legal_all = [], [], [], []
    for g in range(n_games):
        states, legal = env.reset()
        done = False
        while not done:
            for p in range(4):
                st = torch.tensor(states[p]).to(torch.float)
                la = torch.tensor(list(map(int, legal[p])))
                norm = agent.preprocessor(st)
                (logits,) = agent.actor(norm)
                probs = agent.softmax(logits[la])
                row = np.zeros(10, np.float32)
                row[: len(la)] = probs.detach().numpy()
                states_all.append(np.array(states[p], np.int8)); norm_all.append(norm.detach().numpy()); probs_all.append(row)
                legal_all.append(len(la))
            (states, legal), _, done, _ = env.step([int(np.random.choice(l)) for l in legal])
    out["states"] = np.array(states_all, np.int8)
    out["norm"] = np.array(norm_all, np.float32)
    out["probs"] = np.array(probs_all, np.float32)
    out["n_legal"] = np.array(legal_all, np.int32)
    return out


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--only-masked-policy":
        mp = make_masked_policy()
        np.savez_compressed(os.path.join(HERE, "masked_policy.npz"), **mp)
        print("masked policy:", mp["states"].shape, mp["probs"][0])
        return
    if len(sys.argv) > 1 and sys.argv[1] == "--only-position-stats":
        json.dump(make_position_stats(), open(os.path.join(HERE, "position_stats.json"), "w"), separators=(",", ":"))
        print("position stats written")
        return
    if len(sys.argv) > 1 and sys.argv[1] == "--only-policy-train":
        pt = make_policy_train()
        np.savez_compressed(os.path.join(HERE, "policy_train.npz"), **pt)
        print("policy train: losses", pt["loss"], "decisions", pt["obs"].shape)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "--only-policy-rollouts":
        pr = make_policy_rollouts()
        np.savez_compressed(os.path.join(HERE, "policy_rollouts.npz"), **pr)
        print("policy rollouts:", dict(zip(pr["legal"].tolist(), zip(pr["mean"].round(3).tolist(), pr["count"].tolist()))), pr["root_probs"].round(3))
        return
    games = parse_notebook_games()
    assert len(games) == 5, len(games)
    n_pen = replay_check_notebook(games)
    print("notebook games:", len(games), "penalty events:", n_pen,
          "final scores:", [g["snapshots"][-1]["scores"] for g in games])
    json.dump(games, open(os.path.join(HERE, "notebook_games.json"), "w"), separators=(",", ":"))

    traces = {}
    for P in range(2, 11):
        tr = record_traces(P, 120, seed=4000 + P)
        for k, v in tr.items():
            traces[f"p{P}_{k}"] = v
    np.savez_compressed(os.path.join(HERE, "env_traces.npz"), **traces)
    print("env traces written")

    kat = make_kats()
    json.dump(kat, open(os.path.join(HERE, "kat.json"), "w"), separators=(",", ":"))
    print("KAT-A totals", kat["A"]["totals"], "KAT-B", kat["B"])

    mcs = make_mcs(games)
    json.dump(mcs, open(os.path.join(HERE, "mcs_exact.json"), "w"), separators=(",", ":"))
    print("KAT-C", {k: v["mean"] for k, v in mcs["C"]["exact"].items()})

    pol, puct_cases = make_policy_vectors()
    np.savez_compressed(os.path.join(HERE, "policy_vectors.npz"), **pol)
    json.dump(puct_cases, open(os.path.join(HERE, "puct_cases.json"), "w"), separators=(",", ":"))
    print("policy rows", pol["rows_in"].shape)
    pr = make_policy_rollouts()
    np.savez_compressed(os.path.join(HERE, "policy_rollouts.npz"), **pr)
    np.savez_compressed(os.path.join(HERE, "policy_train.npz"), **make_policy_train())
    json.dump(make_position_stats(), open(os.path.join(HERE, "position_stats.json"), "w"), separators=(",", ":"))


if __name__ == "__main__":
    main()
