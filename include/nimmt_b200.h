/*
 * nimmt_b200.h — C ABI of the B200-native 6 nimmt! hot path.
 *
 * This is the whole drop-in boundary: plain pointers and sizes, no torch / C++ types.
 * The reference (coolo/rl-6-nimmt) is pure Python and has no FFI of its own; each entry
 * point below replaces a Python method of the reference, cited as file:line relative to
 * the reference repository.  INTEGRATION.md shows the ctypes binding a maintainer of the
 * reference would add.
 *
 * Conventions
 *   - All `state`, input and output pointers are DEVICE pointers owned by the caller
 *     (in this repo: torch.Tensor storage).  The library never allocates, frees or keeps
 *     a pointer beyond the stream work it enqueues.
 *   - `stream` is a cudaStream_t passed as void* (0 = default stream).  Calls only enqueue;
 *     they never synchronise.  Re-entrant; no global mutable state.
 *   - Return value: NIMMT_OK (0) or a negative NIMMT_E_* code.  Nothing throws.
 *   - Cards are 0-based (0..103) as in the reference (printed 1-based, env.py:244).
 *   - num_rows = 4, num_cards = 104, threshold = 6, hand size 10 are compiled in; they are
 *     the only values any reference caller uses (agents/mcts.py:21-23, play.py:17).
 *   - num_players P in [1, 10] (env.py:19-21: 10 P + 4 <= 104).
 *
 * Packed game state (HBM layout).  A hand only ever loses one card per step, so the ten cards a
 * player was dealt are stored once and a step writes one 32-bit word per player.  Games are stored in tiles of 32:
 *     per tile, one contiguous record of 384 P + 768 bytes:
 *       uint2  cards[P][32]         bytes 0..7 = the cards of hand slots 0..7, ascending (0x7F = none);
 *                                   written by deal / reset_to only
 *       uint32 meta[P][32]          bits 0..9 slot empty (played or never dealt), bits 10..17 the player's
 *                                   cumulative Hornochsen (<= 171), bits 18..24 / 25..31 cards of slots 8 / 9
 *                                   (127 = none)
 *       uint8  rows[32][24]         row record, slot-major: byte 4 j + r = j-th card (oldest first) of row r,
 *                                   j = 0..4; byte 20 + r = cards in the row (bits 0..2) | bull-head sum of the
 *                                   row (bits 3..7, <= 27)
 *   Total ceil(B / 32) * 32 * (12 P + 24) bytes.  A warp's accesses are contiguous in every plane and a tile is
 *   one contiguous run of HBM.
 *   PRECONDITION of every entry point that reads `state`: it was written by nimmt_deal, nimmt_deal_from_perm or
 *   nimmt_reset_to (an all-zero buffer is not a game: its rows have length 0 and stepping it is undefined).
 */
#ifndef NIMMT_B200_H
#define NIMMT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define NIMMT_API __attribute__((visibility("default")))
#else
#define NIMMT_API
#endif

#define NIMMT_OK 0
#define NIMMT_E_BADARG (-1)     /* NULL pointer, bad player count, negative size, bad dtype */
#define NIMMT_E_ALIGN (-2)      /* state pointer not 16-byte aligned */
#define NIMMT_E_CUDA (-3)       /* kernel launch failed; see nimmt_last_cuda_error() */
#define NIMMT_E_UNSUPPORTED (-4)

#define NIMMT_NUM_ROWS 4
#define NIMMT_NUM_CARDS 104
#define NIMMT_THRESHOLD 6
#define NIMMT_HAND 10
#define NIMMT_MAX_PLAYERS 10

/* observation element types for nimmt_observe */
#define NIMMT_DT_I8 0
#define NIMMT_DT_I16 1
#define NIMMT_DT_F32 2
#define NIMMT_DT_I64 3          /* the reference's dtype (np.hstack promotion, env.py:200) */

/* ABI version of this header; bumped on any signature/layout change. */
NIMMT_API int nimmt_abi_version(void);

/* Text of the last CUDA error seen by this thread's calls ("" if none). */
NIMMT_API const char *nimmt_last_cuda_error(void);

/* Bytes of packed state for B games of P players: (12 P + 24) bytes per game, whole tiles of 32 games.  Replaces the list
 * allocations of SechsNimmtEnv.__init__ (env.py:30-32). Returns 0 on bad arguments. */
NIMMT_API size_t nimmt_state_bytes(int64_t num_games, int num_players);

/* Length of one observation vector: 47, or 35 without summaries (env.py:37). */
NIMMT_API int nimmt_obs_len(int include_summaries);

/* Bull heads of one card — SechsNimmtEnv._card_value (env.py:224-239). Host function. */
NIMMT_API int nimmt_card_value(int card);

/* SechsNimmtEnv.reset / _deal (env.py:43-51, 99-112), batched: deals every game from a
 * counter-based RNG keyed by (seed, game0 + b), so results do not depend on launch geometry
 * or on how games are split over GPUs.  Hands are uniform 10-subsets, rows uniform single
 * cards, all distinct; scores zero. */
NIMMT_API int nimmt_deal(void *state, int64_t num_games, int num_players, uint64_t seed, uint64_t game0, void *stream);

/* SechsNimmtEnv._deal given the shuffled decks (env.py:103-112): perm is uint8 [B][104];
 * hand p = perm[10p .. 10p+9], row r = perm[103 - r].  Lets a host that shuffles with the
 * reference's own RNG reproduce the reference's deals bit for bit. */
NIMMT_API int nimmt_deal_from_perm(void *state, const uint8_t *perm, int64_t num_games, int num_players, void *stream);

/* SechsNimmtEnv.reset_to (env.py:53-62): board is int8 [B][4][6], hands int8 [B][P][10],
 * both -1 padded exactly like the observation layout (env.py:191-194, 209-210).  Copies (the
 * reference aliases the caller's lists).  Scores zero.  If `invalid` (uint8 [B]) is not NULL it
 * receives 1 for games whose input was malformed (empty row, > 5 cards in a row, card out of
 * range, duplicate card); such games are still written, with the malformed cards dropped. */
NIMMT_API int nimmt_reset_to(void *state, const int8_t *board, const int8_t *hands, uint8_t *invalid,
                   int64_t num_games, int num_players, void *stream);

/* SechsNimmtEnv.step without the observation rebuild (env.py:64-77, 114-172, 214-249):
 *   actions  uint8 [B][P]  card chosen by each player
 *   rewards  int8  [B][P]  minus the bull heads taken this step (<= 0)          (env.py:169)
 *   done     uint8 [B]     player 0's hand is empty                              (env.py:246-249)
 *   illegal  uint8 [B]     may be NULL; 1 if some card was not in its owner's hand — that game
 *                          is left untouched and its rewards are 0 (env.py:68-69 raises
 *                          InvalidMoveException before any mutation) */
NIMMT_API int nimmt_step(void *state, const uint8_t *actions, int8_t *rewards, uint8_t *done, uint8_t *illegal,
               int64_t num_games, int num_players, void *stream);

/* SechsNimmtEnv._create_states (env.py:174-212): obs [B][P][L] of `dtype`, L = nimmt_obs_len();
 * layout [hand10 | P | (len4 top4 sum4)? | board 4x6], -1 padded.  The legal actions of player p
 * are the non-negative entries of obs[b][p][0..9] (env.py:209).  n_legal (uint8 [B][P]) may be
 * NULL. */
NIMMT_API int nimmt_observe(const void *state, void *obs, uint8_t *n_legal, int64_t num_games, int num_players,
                  int include_summaries, int dtype, void *stream);

/* `turns` consecutive nimmt_step calls in ONE launch, for callers that hold the actions of several turns (replaying recorded
 * games, scripted opponents): actions uint8 [T][B][P], rewards int8 [T][B][P], done / illegal uint8 [T][B] (illegal may be
 * NULL) — exactly what T calls of nimmt_step on the slices would produce (an illegal turn leaves its game untouched and the
 * next turn's cards are checked against the unchanged hands).  The packed state is read and written once per launch instead of
 * once per step: a 32-game tile stays in shared memory for all T turns.  T <= 10; batches that are not a whole number of
 * 32-game tiles are stepped with one launch per turn (B must then be a multiple of 16 for T > 1). */
NIMMT_API int nimmt_step_many(void *state, const uint8_t *actions, int8_t *rewards, uint8_t *done, uint8_t *illegal,
                              int64_t num_games, int num_players, int turns, void *stream);

/* The same for random-vs-random play: `turns` consecutive nimmt_step_random calls (turn, turn + 1, ...) in one launch.
 * actions (may be NULL) uint8 [T][B][P] receives the cards drawn; rewards int8 [T][B][P]; done uint8 [T][B]. */
NIMMT_API int nimmt_step_random_many(void *state, uint8_t *actions, int8_t *rewards, uint8_t *done, int64_t num_games,
                                     int num_players, uint64_t seed, uint32_t turn, uint64_t game0, int turns, void *stream);

/* nimmt_step in a compact TRANSFER format, for hosts that feed actions and read results over PCIe every step (the step itself is
 * the same; about 5 bytes per 4-player game cross the bus instead of 8.1):
 *   slots    uint8 [B][ceil(P/2)]  one 4-bit HAND SLOT per player instead of a card byte: player p = nibble (p & 1) of byte p >> 1;
 *                                  slot s = the s-th card of the hand AS DEALT, ascending (what the first observation showed);
 *                                  a slot >= 10, or one whose card has been played, rejects the step like an illegal card
 *   results  uint8 [B][ceil((5P+2)/8)]  little-endian bit record per game: bits [5p, 5p+5) the bull heads player p took this
 *                                  step (reward = minus that, env.py:169), bit 5P done, bit 5P+1 illegal
 * nimmt_packed_bytes returns the two record sizes.  B must be a multiple of 32. */
NIMMT_API int nimmt_packed_bytes(int num_players, int *action_bytes, int *result_bytes);
NIMMT_API int nimmt_step_packed(void *state, const uint8_t *slots, uint8_t *results, int64_t num_games, int num_players,
                                void *stream);

/* nimmt_step with the FREE ROW CHOICE of the real game, which the reference marks as a TODO (env.py:154-159, ":156 TODO: In the
 * long term this should be up to the agents"; README.md:11): rows uint8 [B][P] names, for every player, the row (0..3) they
 * take IF their card is lower than every row's top card; it replaces the lowest-penalty rule of _pick_row_to_replace and is
 * ignored otherwise.  A value outside 0..3 rejects the game's step exactly like a card that is not in hand (illegal = 1, game
 * untouched).  With rows[b][p] = the lowest-penalty row this is nimmt_step bit for bit.  Everything else as nimmt_step. */
NIMMT_API int nimmt_step_choice(void *state, const uint8_t *actions, const uint8_t *rows, int8_t *rewards, uint8_t *done,
                                uint8_t *illegal, int64_t num_games, int num_players, void *stream);

/* SechsNimmtEnv.step of ONE game, complete, in one launch (env.py:64-77) — the B = 1 drop-in's path.  `cards_host` is a HOST
 * pointer to the P cards played (read during the call, passed to the kernel by value: no copy is enqueued); NULL = do not step,
 * only describe the current state (after reset / reset_to, env.py:51,62).  `record` is 512 bytes of device-ACCESSIBLE memory,
 * 16-byte aligned — typically mapped pinned host memory, so that one stream synchronisation is all the host needs:
 *     [0, P) rewards int8 | [P] done | [P+1] illegal (the game was left untouched, env.py:68-69) | [16, 16+P) cumulative
 *     Hornochsen uint8 | [32, 32 + P L) the observations of all seats, int8 [P][L], L = nimmt_obs_len(include_summaries)
 * `game` selects the game within `state` (0 for a one-game state).  `rows_host` (HOST pointer, P bytes, or NULL): the free
 * row choice of nimmt_step_choice. */
NIMMT_API int nimmt_step1(void *state, int64_t game, const uint8_t *cards_host, const uint8_t *rows_host, int num_players,
                          int include_summaries, void *record, void *stream);

/* Packs per-game flag bytes (the `done` or `illegal` output of nimmt_step: 0 / non-zero) into bits, game b ->
 * bit (b & 31) of bits[b >> 5]; bits has ceil(B / 32) words, unused high bits of the last word are 0.  The
 * reference returns `done` as one Python bool per env.step (env.py:75, 246-249); for a batch whose results go back
 * to host memory the bit form is 8x fewer PCIe bytes. */
NIMMT_API int nimmt_pack_flags(const uint8_t *flags, uint32_t *bits, int64_t num_games, void *stream);

/* Cumulative Hornochsen per player, SechsNimmtEnv._scores (env.py:32,167): uint8 [B][P]. */
NIMMT_API int nimmt_scores(const void *state, uint8_t *scores, int64_t num_games, int num_players, void *stream);

/* DrunkHamster.forward (agents/random.py:8-10) for every player of every game: a uniformly
 * random card of each hand, keyed by (seed, game0 + b, turn).  Empty hands yield 255. */
NIMMT_API int nimmt_random_actions(const void *state, uint8_t *actions, int64_t num_games, int num_players,
                         uint64_t seed, uint32_t turn, uint64_t game0, void *stream);

/* nimmt_random_actions followed by nimmt_step in one kernel (random-vs-random play,
 * play.py:36-45 with DrunkHamster agents).  `actions` may be NULL (not recorded). */
NIMMT_API int nimmt_step_random(void *state, uint8_t *actions, int8_t *rewards, uint8_t *done,
                      int64_t num_games, int num_players, uint64_t seed, uint32_t turn, uint64_t game0,
                      void *stream);

/* One root position of a Monte-Carlo search decision (BaseMCAgent state, agents/mcts.py:43-89). */
typedef struct nimmt_root {
    uint32_t own[4];        /* 104-bit mask of the deciding player's hand (legal_actions)       */
    uint32_t available[4];  /* 104-bit mask of BaseMCAgent.available_cards (agents/mcts.py:62-73) */
    uint8_t rows[4][6];     /* board rows as in the observation, 255-padded (mcts.py:75-85)     */
    uint8_t num_players;    /* int(state[10]) (mcts.py:87-89)                                    */
    uint8_t pad[7];
} nimmt_root;               /* 64 bytes */

/* BaseMCAgent._mcts with MCSAgent._choose_action_mc (agents/mcts.py:91-154, 187-188), batched
 * over D decisions: for every root and every legal first card, `rollouts_per_action` uniform
 * random playouts (stratified instead of the reference's uniform first-card choice; the law of
 * the outcome given the first card is identical — DESIGN.md §5).
 *   stats int64 [D][10][3] += (sum outcome, sum outcome^2, count) per first card, indexed by the
 *   card's rank within the own hand (ascending).  The caller zeroes stats.
 * All roots of one call share `num_players` (roots whose own num_players differs, or whose
 * available set is smaller than (P-1) * |own|, are skipped: their stats stay zero).
 * Rollout ids are striped over `world` ranks (id mod world == rank); summing the stats of all
 * ranks gives the same integers for any world size. */
NIMMT_API int nimmt_mcs_rollouts(const nimmt_root *roots, int num_roots, int num_players, int64_t rollouts_per_action,
                                 uint64_t seed, int rank, int world, int64_t *stats, void *stream);

/* ---- batched Monte-Carlo agents (the bookkeeping BaseMCAgent does on the host, for B games at once) ---- */

/* BaseMCAgent._initialize_game / _memorize_cards / _board_from_state (agents/mcts.py:62-89) for the agent at
 * `seat` of every game: available[b] (one 128-bit card mask per game, caller-owned, uint4 [B]) loses the
 * seat's hand and every card lying on the board now (with `initialize` != 0 it first becomes the full
 * deck, as at the start of a game, agents/mcts.py:47-48); then roots[b] is written for nimmt_mcs_rollouts /
 * nimmt_policy_rollouts.  Call once per decision, before the step, exactly as the agent's forward() runs. */
NIMMT_API int nimmt_mc_roots(const void *state, void *available, nimmt_root *roots, int64_t num_games, int num_players,
                             int seat, int initialize, void *stream);

/* BaseMCAgent._choose_action_from_outcomes (agents/mcts.py:156-165) and the single-card shortcut (:52-53):
 * actions[b][seat] = the legal card with the largest mean outcome in stats int64 [B][10][3] (strict '>' in
 * ascending card order; cards without rollouts never win).  Other seats' bytes of `actions` are untouched. */
NIMMT_API int nimmt_mc_choose(const void *state, const int64_t *stats, uint8_t *actions, int64_t num_games, int num_players,
                              int seat, void *stream);

/* Tournament._compute_elos (tournament.py:157-164) for B finished games, in order (game b+1 sees the ratings game b left):
 * multiplayer Elo as the reference's `multi_elo.calc_elo(players, k)` computes it — every pair of players is a two-player
 * match, K = k / (P - 1), S = 1 / 0.5 / 0 by place, E = 1 / (1 + 10^((R_opp - R_own) / 400)), all seats updated from the
 * ratings before the game.  (`multi_elo` is third-party and absent from the reference tree: parity unpinned.)
 *   scores  int32 [B][P]   session results (negative Hornochsen totals, play.py:69-74): higher = better place
 *   agents  int32 [B][P]   rating slot of every seat, or NULL (seat p = slot p); the seats of one game must be distinct slots
 *   ratings double [A]     in/out, on the device (Tournament starts every agent at elo_initial = 1600, k = elo_k = 32)
 *   history double [B][P]  may be NULL; the seat's rating after each game (what Tournament.elos accumulates) */
NIMMT_API int nimmt_elo_scan(const int32_t *scores, const int32_t *agents, double *ratings, int64_t num_games, int num_players,
                             double k, double *history, void *stream);

/* ---- Alpha0.5 leaf evaluation (the only dense GEMM on the path; tcgen05 tensor cores) ---- */

/* Bytes of the packed policy-weight blob consumed by the kernels below. */
NIMMT_API size_t nimmt_policy_weights_bytes(void);

/* HOST function.  Packs the fp32 parameters of PolicyMCSAgent.actor = MultiHeadedMLP(48, (100, 100), (1,))
 * (utils/nets.py:100-132; torch Linear layout: w1 [100][48], w2 [100][100], w3 [100]) into `blob_host`
 * (nimmt_policy_weights_bytes() bytes of host memory): SechsNimmtStateNormalization(action=True)
 * (utils/preprocessing.py:12-57) is folded into layer 1, matrices are rounded to bf16 and laid out as
 * tcgen05 K-major shared-memory operands, hidden width padded 100 -> 112 with zeros.  Copy the blob to
 * the device once per weight update. */
NIMMT_API int nimmt_policy_pack_weights(const float *w1, const float *b1, const float *w2, const float *b2,
                                        const float *w3, float b3, void *blob_host);

/* PolicyMCSAgent._compute_policy (agents/mcts.py:219-228) for D decisions at once.
 *   obs     int8  [D][47]  observation of the deciding player (nimmt_observe layout, NIMMT_DT_I8);
 *                          its legal cards are the non-negative entries of obs[d][0..9]
 *   weights device copy of the packed blob
 *   probs   float [D][10]  softmax over the legal cards, by hand slot; 0 for empty slots
 *   logits  float [D][10]  may be NULL; the pre-softmax head outputs
 * bf16 operands, fp32 accumulation: probabilities agree with the fp32 reference to ~2e-4. */
NIMMT_API int nimmt_policy_probs(const int8_t *obs, int64_t num_decisions, const void *weights, float *probs,
                                 float *logits, void *stream);

/* ---- state-only 47 -> 100 -> 100 -> 104 nets (the model-free agents' shape) on the same tensor-core tile ---- */

/* Bytes of the packed blob of a masked-policy net. */
NIMMT_API size_t nimmt_masked_weights_bytes(void);

/* HOST function.  Packs MultiHeadedMLP(47, (100, 100), (104,)) (MaskedReinforceAgent.actor, agents/policy.py:30-38; torch Linear
 * layout: w1 [100][47], w2 [100][100], w3 [104][100], b3 [104]) with SechsNimmtStateNormalization(action=False) folded into
 * layer 1, as nimmt_policy_pack_weights does for the 48 -> 100 -> 100 -> 1 net. */
NIMMT_API int nimmt_masked_pack_weights(const float *w1, const float *b1, const float *w2, const float *b2, const float *w3,
                                        const float *b3, void *blob_host);

/* MaskedReinforceAgent.forward up to the sampling (agents/policy.py:45-50) for D decisions: obs int8 [D][47] (4-byte aligned
 * base), probs float [D][10] = softmax of the net's card logits over the cards in hand, by hand slot (0 for empty slots);
 * logits float [D][10] may be NULL.  bf16 operands, fp32 accumulation. */
NIMMT_API int nimmt_masked_probs(const int8_t *obs, int64_t num_decisions, const void *weights, float *probs, float *logits,
                                 void *stream);

/* root-move rules of nimmt_policy_rollouts */
#define NIMMT_ROOT_PUCT 0        /* PUCTAgent._choose_action_mc (agents/mcts.py:276-293) */
#define NIMMT_ROOT_POLICY 1      /* PolicyMCSAgent._choose_action_mc (agents/mcts.py:209-217): sampled from the policy */
#define NIMMT_ROOT_STRATIFIED 2  /* rollout j starts with legal card j mod n (equal budgets; for evaluation and tests) */

/* BaseMCAgent._mcts (agents/mcts.py:91-103) for PolicyMCSAgent / PUCTAgent ("Alpha0.5"), D decisions at once:
 * n_mc sequential rollouts per root; in each, the opponents are dealt from the root's available cards
 * (agents/mcts.py:116-127), every move of every player is sampled from softmax(policy net) over that
 * player's legal cards (agents/mcts.py:139-147, 209-228) except player 0's first move, which follows
 * `root_rule`; the outcome (sum of player 0's rewards, agents/mcts.py:150) is filed under the first card.
 *   stats      int64 [D][10][3]  = (sum outcome, sum outcome^2, visits) per legal card (by rank in the own hand);
 *                                  OVERWRITTEN for playable roots, untouched for skipped ones (see nimmt_mcs_rollouts)
 *   root_probs float [D][10]      policy at the root (what PUCT uses as prior; log of it is the agent's log_prob)
 * The caller applies the final rule (_choose_action_from_outcomes, agents/mcts.py:156-165) to `stats`.
 * A search is inherently sequential (PUCT reads all earlier outcomes), so one decision is never split over
 * GPUs; shard the D roots instead.  With NIMMT_ROOT_PUCT n_mc <= 65535 (the root's outcome histogram, which the
 * median of _normalize_q is read from, counts in 16 bits); larger budgets return NIMMT_E_BADARG. */
NIMMT_API int nimmt_policy_rollouts(const nimmt_root *roots, int num_roots, int num_players, const void *weights, int n_mc,
                                    float c_puct, int root_rule, uint64_t seed, int64_t *stats, float *root_probs,
                                    void *stream);

/* PUCTAgent._choose_action_mc's root rule on its own (agents/mcts.py:276-315: _compute_pucts, _normalize_q and the strict-'>'
 * choice), for D decisions whose rollout outcomes so far are given — the same device code nimmt_policy_rollouts runs at the
 * root, callable (and checkable against the reference) by itself:
 *   offsets      int32 [D+1]   outcomes of decision d are entries offsets[d] .. offsets[d+1]-1 of the two lists below
 *   action_index int32 [...]   rank (0..9) of the first card the rollout started with, in the order the rollouts were played
 *   outcomes     int32 [...]   the rollout's outcome (sum of player 0's rewards, in [-171, 0])
 *   probs        float [D][10] policy prior by hand slot;  n_legal int32 [D] legal cards of the decision (1..10)
 *   pucts        double [D][10] receives the PUCT value of every legal card (NaN where the reference yields NaN), 0 beyond
 *   choice       int32 [D]     receives the selected hand slot
 * At most 65535 outcomes per decision. */
NIMMT_API int nimmt_puct_choose(const int32_t *offsets, const int32_t *action_index, const int32_t *outcomes, const float *probs,
                                const int32_t *n_legal, int num_decisions, float c_puct, double *pucts, int32_t *choice,
                                void *stream);

#ifdef __cplusplus
}
#endif
#endif /* NIMMT_B200_H */
