"""6 nimmt! environments backed by the sm_100a kernels.

Two front ends over the same packed device state (include/nimmt_b200.h):

* :class:`BatchedSechsNimmtEnv` — B independent games stepped in lock-step on the GPU.  This is
  the throughput path (BASELINE.json configs 2 and 5).
* :class:`SechsNimmtEnv` — a drop-in for the reference class of the same name
  (``rl_6_nimmt/env.py:13``): identical constructor, ``reset / reset_to / step / render /
  _create_states``, identical return structure (lists of int64[47] arrays + legal-card lists,
  int32 rewards, Python bool done, ``{}``), identical exceptions.  It is a B = 1 view of the
  batched engine; every rule is evaluated on the device.

There is no CPU implementation of the rules in this package.
"""
import logging

import numpy as np
import torch

from . import _native as N

logger = logging.getLogger(__name__)

NUM_ROWS, NUM_CARDS, THRESHOLD, HAND_SIZE = 4, 104, 6, 10

_TORCH_DT = {torch.int8: N.DT_I8, torch.int16: N.DT_I16, torch.float32: N.DT_F32, torch.int64: N.DT_I64}


class InvalidMoveException(Exception):
    """Same name and meaning as rl_6_nimmt/env.py:9-10."""


class Discrete:
    """Duck-typed gym.spaces.Discrete (gym is not a dependency; only ``.n`` is read, agents/base.py:19)."""

    def __init__(self, n):
        self.n = n

    def __repr__(self):
        return f"Discrete({self.n})"


class Box:
    """Duck-typed gym.spaces.Box (only ``.shape`` is read, agents/base.py:18)."""

    def __init__(self, low, high, shape, dtype=np.float32):
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype

    def __repr__(self):
        return f"Box({self.low}, {self.high}, {self.shape}, {self.dtype})"


def unpack_packed_results(packed, num_players):
    """The bit records of nimmt_step_packed (uint8 [B, ceil((5P+2)/8)], little-endian: 5 bits of bull heads per player, then done,
    then illegal) -> (rewards int8 [B,P], done bool [B], illegal bool [B]).  Works on any device."""
    P = num_players
    rec = torch.zeros(packed.shape[0], dtype=torch.int64, device=packed.device)
    for b in range(packed.shape[1]):
        rec |= packed[:, b].to(torch.int64) << (8 * b)
    rewards = torch.stack([-((rec >> (5 * p)) & 31) for p in range(P)], dim=1).to(torch.int8)
    return rewards, ((rec >> (5 * P)) & 1).bool(), ((rec >> (5 * P + 1)) & 1).bool()


def _as_device(x, dtype, device):
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=dtype).contiguous()
    return torch.as_tensor(np.ascontiguousarray(x), dtype=dtype).to(device)


class BatchedSechsNimmtEnv:
    """B games of ``num_players`` players on one GPU.

    All methods enqueue work on the current CUDA stream and return device tensors; nothing
    synchronises unless stated.  ``seed`` keys the counter-based RNG used by :meth:`reset` and
    :meth:`random_actions`; ``game0`` is the global index of game 0, so a batch split over several
    GPUs deals exactly the games the unsplit batch would.
    """

    def __init__(self, num_games, num_players, include_summaries=True, device=None, seed=0, game0=0):
        assert num_players > 0                                      # env.py:19
        assert NUM_CARDS >= 10 * num_players + NUM_ROWS            # env.py:21
        assert num_games >= 0
        if not torch.cuda.is_available():
            raise N.NimmtNativeError("BatchedSechsNimmtEnv needs a CUDA device; there is no CPU fallback")
        self.lib = N.lib()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.num_games, self.num_players = int(num_games), int(num_players)
        self.include_summaries = bool(include_summaries)
        self.obs_len = self.lib.nimmt_obs_len(int(self.include_summaries))
        self.seed, self.game0 = int(seed), int(game0)
        self.episode = 0   # bumped by reset(): later deals differ from earlier ones
        self.turn = 0
        B, P = self.num_games, self.num_players
        with torch.cuda.device(self.device):
            self.state = torch.zeros(max(self.lib.nimmt_state_bytes(B, P), 16), dtype=torch.uint8, device=self.device)
            # rewards and the bit-packed done flags share one buffer so that step_host can return both in one D2H copy
            self._words = (B + 31) // 32
            self._out = torch.zeros(((B * P + 15) // 16 * 16 + 4 * self._words,), dtype=torch.uint8, device=self.device)
            self.rewards = self._out[: B * P].view(torch.int8).view(B, P)
            self._done_bits = self._out[(B * P + 15) // 16 * 16:].view(torch.int32)
            self.done = torch.zeros((B,), dtype=torch.uint8, device=self.device)
            self.illegal = torch.zeros((B,), dtype=torch.uint8, device=self.device)
            self._actions = torch.zeros((B, P), dtype=torch.uint8, device=self.device)
            # An all-zero buffer is not a game (rows of length 0): the kernels require a state written by deal / reset_to.
            # Deal a placeholder so that step() before reset() is merely a step of some game, never an out-of-range access.
            N.check(self.lib.nimmt_deal(N.ptr(self.state), B, P, 0, 0, self._stream()), "nimmt_deal")

    # -- helpers -----------------------------------------------------------------------------
    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _deal_seed(self):
        return (self.seed + 0x9E3779B97F4A7C15 * self.episode) & 0xFFFFFFFFFFFFFFFF

    # -- reference API, batched ----------------------------------------------------------------
    def reset(self, seed=None):
        """SechsNimmtEnv.reset (env.py:43-51) for every game; device RNG."""
        if seed is not None:
            self.seed, self.episode = int(seed), 0
        with torch.cuda.device(self.device):
            N.check(self.lib.nimmt_deal(N.ptr(self.state), self.num_games, self.num_players, self._deal_seed(),
                                        self.game0, self._stream()), "nimmt_deal")
        self.episode += 1
        self.turn = 0
        return self

    def reset_from_perm(self, perm):
        """_deal from caller-supplied shuffled decks, uint8 [B,104] (env.py:103-112)."""
        perm = _as_device(perm, torch.uint8, self.device)
        assert perm.shape == (self.num_games, NUM_CARDS)
        with torch.cuda.device(self.device):
            N.check(self.lib.nimmt_deal_from_perm(N.ptr(self.state), N.ptr(perm), self.num_games, self.num_players,
                                                  self._stream()), "nimmt_deal_from_perm")
        self.turn = 0
        return self

    def reset_to(self, board, hands, check=True):
        """SechsNimmtEnv.reset_to (env.py:53-62): board int8 [B,4,6], hands int8 [B,P,10], -1 padded."""
        board = _as_device(board, torch.int8, self.device)
        hands = _as_device(hands, torch.int8, self.device)
        assert board.shape == (self.num_games, NUM_ROWS, THRESHOLD), board.shape
        assert hands.shape == (self.num_games, self.num_players, HAND_SIZE), hands.shape
        invalid = torch.zeros((self.num_games,), dtype=torch.uint8, device=self.device) if check else None
        with torch.cuda.device(self.device):
            N.check(self.lib.nimmt_reset_to(N.ptr(self.state), N.ptr(board), N.ptr(hands), N.ptr(invalid), self.num_games,
                                            self.num_players, self._stream()), "nimmt_reset_to")
        if check and bool(invalid.any()):
            bad = int(torch.nonzero(invalid)[0])
            raise ValueError(f"reset_to: malformed board/hands for game {bad} (empty or over-long row, bad or duplicate card)")
        self.turn = 0
        return self

    def step(self, actions, check=False, rows=None):
        """SechsNimmtEnv.step (env.py:64-77) without the observation rebuild.

        ``rows`` (uint8 [B,P], optional): the free row choice of the real game, which the reference leaves as a TODO
        (env.py:156, README.md:11) — the row (0..3) each player takes IF their card undercuts every row, instead of the
        lowest-penalty rule.  None = the reference's rule.

        actions: uint8 [B,P] device tensor.  Returns (rewards int8 [B,P], done uint8 [B]) — views of
        buffers owned by the env, overwritten by the next step.  ``self.illegal`` flags games whose
        move was rejected (left untouched, env.py:68-69); with ``check=True`` the call synchronises
        and raises InvalidMoveException if any game was flagged.
        """
        assert actions.shape == (self.num_games, self.num_players) and actions.dtype == torch.uint8 and actions.is_cuda
        actions = actions.contiguous()
        with torch.cuda.device(self.device):
            if rows is None:
                N.check(self.lib.nimmt_step(N.ptr(self.state), N.ptr(actions), N.ptr(self.rewards), N.ptr(self.done),
                                            N.ptr(self.illegal), self.num_games, self.num_players, self._stream()), "nimmt_step")
            else:
                assert rows.shape == actions.shape and rows.dtype == torch.uint8 and rows.is_cuda
                rows = rows.contiguous()
                N.check(self.lib.nimmt_step_choice(N.ptr(self.state), N.ptr(actions), N.ptr(rows), N.ptr(self.rewards), N.ptr(self.done),
                                                   N.ptr(self.illegal), self.num_games, self.num_players, self._stream()), "nimmt_step_choice")
        self.turn += 1
        if check and bool(self.illegal.any()):
            bad = int(torch.nonzero(self.illegal)[0])
            raise InvalidMoveException(f"game {bad}: a played card is not in its owner's hand")
        return self.rewards, self.done

    def step_many(self, actions, rewards=None, done=None, illegal=None):
        """``T`` consecutive steps in one launch (nimmt_step_many): actions uint8 [T,B,P] device tensor -> (rewards int8 [T,B,P],
        done uint8 [T,B], illegal uint8 [T,B]), exactly what T calls of :meth:`step` on the slices produce.  For callers that
        hold the actions of several turns (replaying recorded games): the packed state is read and written once per launch,
        a 32-game tile stays in shared memory for all T turns."""
        T, B, P = actions.shape
        assert (B, P) == (self.num_games, self.num_players) and actions.dtype == torch.uint8 and actions.is_cuda and 1 <= T <= 10
        actions = actions.contiguous()
        dev = self.device
        rewards = torch.empty((T, B, P), dtype=torch.int8, device=dev) if rewards is None else rewards
        done = torch.empty((T, B), dtype=torch.uint8, device=dev) if done is None else done
        illegal = torch.empty((T, B), dtype=torch.uint8, device=dev) if illegal is None else illegal
        with torch.cuda.device(dev):
            N.check(self.lib.nimmt_step_many(N.ptr(self.state), N.ptr(actions), N.ptr(rewards), N.ptr(done), N.ptr(illegal), B, P, T,
                                             self._stream()), "nimmt_step_many")
        self.turn += T
        return rewards, done, illegal

    def step_random_many(self, turns, rewards=None, done=None, record_actions=False):
        """``turns`` consecutive :meth:`step_random` calls in one launch (nimmt_step_random_many): random-vs-random play with the
        state resident in shared memory.  Returns (rewards int8 [T,B,P], done uint8 [T,B], actions uint8 [T,B,P] or None)."""
        T, B, P, dev = int(turns), self.num_games, self.num_players, self.device
        rewards = torch.empty((T, B, P), dtype=torch.int8, device=dev) if rewards is None else rewards
        done = torch.empty((T, B), dtype=torch.uint8, device=dev) if done is None else done
        acts = torch.empty((T, B, P), dtype=torch.uint8, device=dev) if record_actions else None
        with torch.cuda.device(dev):
            N.check(self.lib.nimmt_step_random_many(N.ptr(self.state), N.ptr(acts), N.ptr(rewards), N.ptr(done), B, P, self._deal_seed(),
                                                    self.turn, self.game0, T, self._stream()), "nimmt_step_random_many")
        self.turn += T
        return rewards, done, acts

    # -- the compact transfer format (nimmt_step_packed): for hosts that feed actions and read results over PCIe every step --
    def packed_sizes(self):
        """(action bytes, result bytes) per game of the packed format: ceil(P/2) and ceil((5P+2)/8)."""
        P = self.num_players
        return (P + 1) // 2, (5 * P + 2 + 7) // 8

    @staticmethod
    def pack_slots(cards, dealt_hands):
        """cards [B,P] (any integer dtype) + the hands AS DEALT [B,P,10] (the first observation's hand blocks, ascending) ->
        uint8 [B, ceil(P/2)]: one 4-bit hand slot per player (15 where the card is not in the dealt hand: an illegal move)."""
        hit = dealt_hands.to(torch.int16) == cards.to(torch.int16).unsqueeze(2)
        slot = torch.where(hit.any(dim=2), hit.to(torch.uint8).argmax(dim=2), torch.full_like(cards, 15, dtype=torch.int64)).to(torch.uint8)
        if slot.shape[1] % 2:
            slot = torch.cat([slot, torch.zeros_like(slot[:, :1])], dim=1)
        return (slot[:, 0::2] | (slot[:, 1::2] << 4)).contiguous()

    def unpack_results(self, packed):
        """uint8 [B, ceil((5P+2)/8)] -> (rewards int8 [B,P], done bool [B], illegal bool [B])."""
        return unpack_packed_results(packed, self.num_players)

    def step_packed(self, slots, out=None):
        """:meth:`step` in the packed transfer format: slots uint8 [B, ceil(P/2)] device tensor (:meth:`pack_slots`) ->
        uint8 [B, ceil((5P+2)/8)] bit records (:meth:`unpack_results`).  B must be a multiple of 32."""
        ab, rb = self.packed_sizes()
        assert slots.shape == (self.num_games, ab) and slots.dtype == torch.uint8 and slots.is_cuda
        slots = slots.contiguous()
        if out is None:
            out = torch.empty((self.num_games, rb), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            N.check(self.lib.nimmt_step_packed(N.ptr(self.state), N.ptr(slots), N.ptr(out), self.num_games, self.num_players, self._stream()),
                    "nimmt_step_packed")
        self.turn += 1
        return out

    def step_host_packed(self, slots_host, results_host):
        """step_host in the packed format: H2D copy of ``slots_host`` (pinned uint8 [B, ceil(P/2)]), the step, ONE D2H copy of the
        bit records into ``results_host`` (pinned uint8 [B, ceil((5P+2)/8)]).  5 bytes per 4-player game cross PCIe instead of
        8.1.  Nothing synchronises."""
        if not hasattr(self, "_slots_dev"):
            ab, rb = self.packed_sizes()
            self._slots_dev = torch.empty((self.num_games, ab), dtype=torch.uint8, device=self.device)
            self._packed_dev = torch.empty((self.num_games, rb), dtype=torch.uint8, device=self.device)
        self._slots_dev.copy_(slots_host, non_blocking=True)
        self.step_packed(self._slots_dev, self._packed_dev)
        results_host.copy_(self._packed_dev, non_blocking=True)
        return results_host

    def host_out_buffer(self):
        """A pinned uint8 buffer for the packed form of step_host, plus views of its two parts:
        (buffer, rewards int8 [B,P], done_bits int32 [ceil(B/32)])."""
        B, P = self.num_games, self.num_players
        buf = torch.empty_like(self._out, device="cpu").pin_memory()
        off = (B * P + 15) // 16 * 16
        return buf, buf[: B * P].view(torch.int8).view(B, P), buf[off:].view(torch.int32)

    def step_host(self, actions_host, rewards_host, done_host=None):
        """step() for callers whose buffers live in (pinned) host memory — the end-to-end path.

        Enqueues, on the current stream: H2D copy of ``actions_host`` (uint8 [B,P]), the step kernel, and the
        D2H copies of the results.  Three result forms:
          step_host(a, rewards int8 [B,P], done uint8 [B])                one flag byte per game, two copies
          step_host(a, rewards int8 [B,P], done int32 [ceil(B/32)])       one bit per game (game b = bit b % 32 of
                                                                          word b // 32), two copies
          step_host(a, out)  with out from host_out_buffer()              rewards + done bits in ONE copy
        Nothing synchronises; the caller syncs the stream (or an event) before reading the host buffers.
        """
        self._actions.copy_(actions_host, non_blocking=True)
        self.step(self._actions)
        if done_host is None or done_host.dtype == torch.int32:
            with torch.cuda.device(self.device):
                N.check(self.lib.nimmt_pack_flags(N.ptr(self.done), N.ptr(self._done_bits), self.num_games, self._stream()),
                        "nimmt_pack_flags")
        if done_host is None:
            rewards_host.copy_(self._out, non_blocking=True)
            return rewards_host
        rewards_host.copy_(self.rewards, non_blocking=True)
        done_host.copy_(self._done_bits if done_host.dtype == torch.int32 else self.done, non_blocking=True)
        return rewards_host, done_host

    def random_actions(self, out=None, turn=None):
        """DrunkHamster for every seat (agents/random.py:8-10): uint8 [B,P]."""
        out = self._actions if out is None else out
        with torch.cuda.device(self.device):
            N.check(self.lib.nimmt_random_actions(N.ptr(self.state), N.ptr(out), self.num_games, self.num_players,
                                                  self._deal_seed(), self.turn if turn is None else int(turn), self.game0,
                                                  self._stream()), "nimmt_random_actions")
        return out

    def step_random(self, record_actions=False):
        """random_actions + step fused in one kernel (random-vs-random play)."""
        acts = self._actions if record_actions else None
        with torch.cuda.device(self.device):
            N.check(self.lib.nimmt_step_random(N.ptr(self.state), N.ptr(acts), N.ptr(self.rewards), N.ptr(self.done),
                                               self.num_games, self.num_players, self._deal_seed(), self.turn, self.game0,
                                               self._stream()), "nimmt_step_random")
        self.turn += 1
        return self.rewards, self.done

    def observe(self, dtype=torch.float32, out=None, n_legal=None):
        """SechsNimmtEnv._create_states (env.py:174-212): [B,P,L] tensor, L = 47 (35 without summaries).

        Legal actions of seat p in game b are the non-negative entries of obs[b,p,:10].
        """
        B, P, L = self.num_games, self.num_players, self.obs_len
        if out is None:
            out = torch.empty((B, P, L), dtype=dtype, device=self.device)
        assert out.shape == (B, P, L) and out.is_contiguous()
        with torch.cuda.device(self.device):
            N.check(self.lib.nimmt_observe(N.ptr(self.state), N.ptr(out), N.ptr(n_legal), B, P, int(self.include_summaries),
                                           _TORCH_DT[out.dtype], self._stream()), "nimmt_observe")
        return out

    def scores(self, out=None):
        """Cumulative Hornochsen per seat (env.py:32,167): uint8 [B,P]."""
        if out is None:
            out = torch.empty((self.num_games, self.num_players), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            N.check(self.lib.nimmt_scores(N.ptr(self.state), N.ptr(out), self.num_games, self.num_players, self._stream()),
                    "nimmt_scores")
        return out


class SechsNimmtEnv:
    """Drop-in for ``rl_6_nimmt.env.SechsNimmtEnv`` (env.py:13-256), rules evaluated on the GPU.

    ``reset()`` shuffles with ``np.random`` exactly as the reference does (env.py:103-104), so a
    program that seeds NumPy sees the same deals from both implementations.
    """

    metadata = {"render.modes": ["human"]}

    def __init__(self, num_players, num_rows=4, num_cards=104, threshold=6, include_summaries=True, player_names=None,
                 verbose=True, device=None, row_choice="lowest"):
        assert num_players > 0
        assert num_rows > 0
        assert num_cards >= 10 * num_players + num_rows
        if (num_rows, num_cards, threshold) != (NUM_ROWS, NUM_CARDS, THRESHOLD):
            raise NotImplementedError("the CUDA kernels are built for num_rows=4, num_cards=104, threshold=6 "
                                      "(the only values the reference's callers use)")
        self._num_players, self._num_rows, self._num_cards, self._threshold = num_players, num_rows, num_cards, threshold
        self._include_summaries = include_summaries
        self._player_names = player_names
        self.action_space = Discrete(num_cards)
        self.reward_range = (-float("inf"), 0)
        state_shape = (10 + 1 + int(include_summaries) * 3 * num_rows + num_rows * threshold,)
        self.observation_space = Box(low=-1.0, high=2.0, shape=state_shape, dtype=np.float32)
        self.spec = None
        self.verbose = verbose
        self._engine = None
        self._device = device
        # "lowest": the reference's rule (env.py:154-159).  "agent": the real game's free choice, the reference's TODO
        # (env.py:156) — step(action, rows=[...]) names the row each player takes if their card undercuts every row.
        assert row_choice in ("lowest", "agent")
        self._row_choice = row_choice
        self._cache = None  # (obs int64 [P,L] numpy, scores int32 [P]) of the current state
        self._record = None  # 512 bytes of mapped pinned host memory: what nimmt_step1 reports about the game

    # engine is created lazily: agents/base.py:11-12 builds a throw-away env just to read the spaces
    def _eng(self):
        if self._engine is None:
            self._engine = BatchedSechsNimmtEnv(1, self._num_players, self._include_summaries, device=self._device)
            board = -np.ones((1, NUM_ROWS, THRESHOLD), np.int8)
            board[0, :, 0] = np.arange(NUM_ROWS)  # placeholder position until reset()/reset_to()
            self._engine.reset_to(board, -np.ones((1, self._num_players, HAND_SIZE), np.int8), check=False)
        return self._engine

    def _launch1(self, cards=None, rows=None):
        """One kernel for the whole call (nimmt_step1): with ``cards`` the step itself, always the rewards / done / illegal /
        scores / every seat's observation of the resulting state, written straight into pinned host memory; one stream
        synchronisation, no copies.  Returns the record as a numpy view (valid until the next call)."""
        eng = self._eng()
        if self._record is None:
            self._record = torch.zeros(512, dtype=torch.uint8).pin_memory()
            self._record_np = self._record.numpy()
            self._cards = np.zeros(16, np.uint8)
            self._rows = np.zeros(16, np.uint8)
        ptr = rptr = None
        if cards is not None:
            self._cards[: self._num_players] = cards
            ptr = self._cards.ctypes.data
        if rows is not None:
            self._rows[: self._num_players] = rows
            rptr = self._rows.ctypes.data
        with torch.cuda.device(eng.device):
            stream = torch.cuda.current_stream(eng.device)
            N.check(eng.lib.nimmt_step1(N.ptr(eng.state), 0, ptr, rptr, self._num_players, int(self._include_summaries),
                                        self._record.data_ptr(), stream.cuda_stream), "nimmt_step1")
            stream.synchronize()
        rec = self._record_np
        P, L = self._num_players, eng.obs_len
        self._cache = (rec[32:32 + P * L].view(np.int8).reshape(P, L).astype(np.int64), rec[16:16 + P].astype(np.int32))
        return rec

    def _sync_cache(self):
        if self._cache is None:
            self._launch1()
        return self._cache

    # -- reference API -----------------------------------------------------------------------------
    def reset(self):
        cards = np.arange(0, self._num_cards, 1, dtype=np.int32)
        np.random.shuffle(cards)  # same RNG call as env.py:103-104
        if self.verbose:
            logger.debug("Dealing cards")
        self._eng().reset_from_perm(cards.astype(np.uint8)[None])
        self._cache = None
        return self._create_states()

    def reset_to(self, board, hands):
        assert len(board) == self._num_rows and len(hands) == self._num_players
        b = -np.ones((1, NUM_ROWS, THRESHOLD), np.int8)
        for r, cards in enumerate(board):
            b[0, r, : len(cards)] = cards
        h = -np.ones((1, self._num_players, HAND_SIZE), np.int8)
        for p, cards in enumerate(hands):
            h[0, p, : len(cards)] = sorted(int(c) for c in cards)
        self._eng().reset_to(b, h)
        self._cache = None
        return self._create_states()

    def step(self, action, rows=None):
        assert len(action) == self._num_players                     # env.py:67
        cards = [int(c) for c in action]
        P = self._num_players
        if self._row_choice == "agent":
            assert rows is not None and len(rows) == P, "row_choice='agent': step(action, rows) needs one row per player"
            rows = [int(r) if 0 <= int(r) < 256 else 255 for r in rows]
        else:
            assert rows is None, "rows are only accepted with row_choice='agent'"
        rec = self._launch1([c if 0 <= c < 256 else 255 for c in cards], rows)
        if rec[P + 1]:
            if rows is not None and any(not 0 <= r < NUM_ROWS for r in rows):
                raise InvalidMoveException(f"Row choices {rows} must be in 0..{NUM_ROWS - 1}")
            # the device rejected the step and left the game untouched (env.py:68-69, 117-118)
            hands = self._hands
            for player, card in enumerate(cards):
                if card not in hands[player]:
                    raise InvalidMoveException(f"Player {player + 1} tried to play card {card + 1}, but their hand is {hands[player]}")
            raise InvalidMoveException("illegal move")  # unreachable: device and host views agree
        if self.verbose and logger.isEnabledFor(logging.DEBUG):
            for card, player in sorted((c, p) for p, c in enumerate(cards)):
                logger.debug(f"{self._player_name(player)} plays card {card + 1}")
        rewards, done = rec[:P].view(np.int8).astype(np.int32), bool(rec[P])
        return self._create_states(), rewards, done, dict()

    def _create_states(self):
        obs, _ = self._sync_cache()
        player_states = [obs[p].copy() for p in range(self._num_players)]
        legal_actions = [[int(c) for c in obs[p, :HAND_SIZE] if c >= 0] for p in range(self._num_players)]
        return player_states, legal_actions

    # -- the attributes tests and probes poke at (env.py:30-32) -------------------------------------
    @property
    def _board(self):
        obs, _ = self._sync_cache()
        grid = obs[0, -NUM_ROWS * THRESHOLD:].reshape(NUM_ROWS, THRESHOLD)
        return [[int(c) for c in row if c >= 0] for row in grid]

    @property
    def _hands(self):
        obs, _ = self._sync_cache()
        return [[int(c) for c in obs[p, :HAND_SIZE] if c >= 0] for p in range(self._num_players)]

    @property
    def _scores(self):
        return self._sync_cache()[1].copy()

    def _is_done(self):
        return len(self._hands[0]) == 0

    @staticmethod
    def _card_value(card):
        assert 0 <= card < NUM_CARDS
        return N.lib().nimmt_card_value(int(card))

    # -- host-only presentation (env.py:79-97, 241-256) ----------------------------------------------
    def _format_card(self, card):
        marks = {1: " ", 2: ".", 3: ":", 5: "+", 7: "#"}
        return f"{card + 1:>3d}{marks[self._card_value(card)]}"

    def _player_name(self, player):
        if self._player_names is None:
            return f"Player {player + 1:d}"
        width = max(len(name) for name in self._player_names)
        return f"{self._player_names[player]:<{width}} (player {player + 1:d})"

    def render(self, mode="human"):
        rule = "-" * 120
        logger.info(rule)
        logger.info("Board:")
        for cards in self._board:
            shown = " ".join(self._format_card(c) for c in cards)
            logger.info("  " + shown + "   _ " * (self._threshold - len(cards) - 1) + "   * ")
        logger.info("Players:")
        scores, hands = self._scores, self._hands
        for player in range(self._num_players):
            held = "no cards " if not hands[player] else "cards " + " ".join(self._format_card(c) for c in hands[player])
            logger.info(f"  {self._player_name(player)}: {scores[player]:>3d} Hornochsen, " + held)
        if self._is_done():
            winner, loser = int(np.argmin(scores)), int(np.argmax(scores))
            logger.info(f"The game is over! {self._player_name(winner)} wins, {self._player_name(loser)} loses. Congratulations!")
        logger.info(rule)
