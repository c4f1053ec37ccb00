// cbx_layout.h -- HBM data layout of a batch and the parameter block of the kernels (host + device).
//
// Per-env dynamic state is a TILED structure of arrays of 32-bit words: word w of env e lives at
// state[((e / 32) * S + w) * 32 + e % 32], so the state of a tile of 32 consecutive envs is one contiguous block of
// S rows of 128 bytes (one row per state word) that moves with a single TMA bulk copy in each direction.  A CTA stages the S x 32 tile in shared memory (row-major, i.e. word-major / env-minor: thread
// e touching word w hits bank e -- conflict-free for the one-thread-per-env game logic), plays the step on it
// and streams it back.  What replaces the reference's Python objects:
//   discovered order + inverse map, agent_installed / ever_owned / not-running bitsets, 2-bit privilege levels,
//   reimaging countdowns (live env and the MARLon defender's stale copy), discovered-property bitsets,
//   2 bits per (node, vulnerability) replacing the last_attack timestamps, gathered-secret and cached-triple
//   bitsets, the ordered credential cache, and the wrappers' counters.
#ifndef CBX_LAYOUT_H_
#define CBX_LAYOUT_H_

#include <stdint.h>

#include "../../include/cbx.h"

#ifndef CBX_TILE
#define CBX_TILE 32          // envs per tile (a multiple of 32, at most CBX_THREADS: one game-logic thread per env)
#endif
#define CBX_THREADS 128      // threads per CTA
#define CBX_MAX_LEAK 64
#ifndef CBX_MIN_CTAS
#define CBX_MIN_CTAS 8       // resident CTAs per SM the step kernel is compiled for (caps registers at 64)
#endif
// kernel op bits
#define CBX_OP_RESET 1
#define CBX_OP_ATTACKER 2
#define CBX_OP_DEFENDER 4
#define CBX_OP_NOTIFY 8    // with CBX_OP_RESET: only raise the reset_request flags (EnvironmentEventSource.notify_reset)

struct cbx_layout {
  // dimensions
  int n, N, C, LEAK, P, L, R, nprops, nsecrets, ntriples, nservices;
  int LEAKS;   // leaked-credential slots actually staged: min(LEAK, longest LeakedCredentials list of the scenario); the rest of
               // the LEAK slots of an observation are always zero
  int Wn;      // words of a node bitset
  int PW;      // words of a property bitset (1 or 2)
  int AW;      // words of the attacked bits of one node: 2 bits per vulnerability
  int OW;      // words of the owned-by-discovery-index bitset (over N)
  // word offsets inside the per-env state
  int o_hdr;          // bits 0-7 n_discovered | 8-23 n_cached | 24 done | 25 att_reset_request | 26 def_reset_request |
                      //      27 has_breached_sla | 28 cyber_rewards non-empty | 29 rewards non-empty
  int o_stepcount;
  int o_att_ts, o_def_ts;
  int o_att_valid, o_att_invalid, o_def_valid, o_def_invalid;
  int o_last_cyber;   // f32: AttackerEnvWrapper.cyber_rewards[-1]
  int o_last_reward;  // f32: AttackerEnvWrapper.rewards[-1]
  int o_last_att;     // f32: DefenderEnvWrapper.__last_attacker_reward
  int o_att_return, o_def_return;  // f32 running episode returns (Monitor)
  int o_ep_sum;       // f32 sum of CyberBattleEnv.__episode_rewards
  int o_avail;        // bits 0-7 live not-running count at last tick | 8-15 shadow count at last tick | 16-23 previous shadow count
  int o_installed, o_everowned, o_notrunning;  // Wn words each
  int o_priv;         // 2 bits per node
  int o_tags;         // 4 bits per node (dynamic privilege_N tags)
  int o_cd_live;      // 8 bits per node: 0 = Running, k = Imaging with k ticks left
  int o_cd_shadow;
  int o_disc_order;   // 8 bits per discovery index: node
  int o_disc_idx;     // 8 bits per node: discovery index, 0xFF = undiscovered
  int o_props;        // n * PW
  int o_attacked;     // n * AW
  int o_gathered;     // ceil(nsecrets/32)
  int o_cached;       // ceil(ntriples/32): triple already in the cache
  int o_cache;        // 16 bits per cache slot: triple id
  int o_fw;           // live defender binding only (else -1): 2 words per firewall rule-list group -- port names with a rule,
                      // names whose first rule allows (cbx.h CBX_FX_*); part of the CyberBattleEnv's state (reset with it)
  int n_fw_groups;
  int o_cyber_begin;  // words [o_cyber_begin, S) (+ parts of hdr / avail) belong to the CyberBattleEnv and are re-initialised
                      // by CyberBattleEnv.reset(); the words before it belong to the MARLon wrappers and the stale copy
  int S;              // words per env
  // per-env staging words written by the game-logic thread for the encoder (word-major like the state tile):
  //   [0,8) scalars | 8 obs kind | 9 attacker done | 10 defender done | [11, 11+Wn) installed bits at defender done
  int g_leaked;       // LEAKS leaked-credential slots, one packed word each (leak_pack)
  int g_inst;         // Wn: agent_installed bits as the observation sees them (before the built-in defender moves)
  int g_priv;         // ceil(n/16): privilege levels as the observation sees them
  int G;              // staging words per env
  // byte sizes per env of the dense masks
  int sz_local, sz_remote, sz_connect;
};

struct cbx_fastdiv {  // x / d for x < 2^31: (m ? umulhi(x, m) : x) >> s
  uint32_t m, s;
};

struct cbx_enc_consts {  // divisors of the encoder, fixed per batch
  cbx_fastdiv d_leaked, d_cachem, d_props, d_priv, d_nprops, d_L, d_local, d_remote, d_connect, d_rowr, d_rowc, d_C, d_n,
      d_6n, d_svc;
  int desc_words;
  // warp-per-env fast path: every row of an env's remote / connect mask is either all zero or one and the same byte
  // string; a warp keeps that string in registers and stores it row by row
  int warp_env;                      // encoder variant: 0 generic, 1 warp-per-env, 2/3 statically specialised
  int debug_skip;                    // experiments only (env CBX_DEBUG_SKIP): bit0 small fields, 1 local, 2 remote, 3 connect, 4 defender
  int tmpl_unit_r, tmpl_unit_c;      // store granule in bytes: gcd(row length, 16)
};

struct cbx_smem_plan {  // shared-memory carve-up in 32-bit words
  int tables, state, stage, desc, acts, lut, bars, total_bytes;
};

// Pipelined step kernel (cbx_pipe.cuh): per CTA `wl` game-logic warps (one thread per env; they also lay out the small
// observation fields of their tile, env-major, and hand them to the TMA engine) run ahead of `we` encoder warps (which build
// the action-mask template rows of each env and hand those to the TMA engine) through `nslot` descriptor slots.
struct cbx_pipe_plan {
  int enabled;
  int wl, we, nslot;
  int gs;           // connect-mask rows per bulk copy: 1 when a row is a multiple of 16 bytes, 2 when it is 8 mod 16
  int logic_tma;    // 1 (default): the logic warps move their state tile / field images with TMA bulk copies too;
                    // 0 (CBX_PIPE_LOGIC_TMA=0): with plain 16-byte loads / stores (measured slower: HBM read latency
                    // under the write stream is ~12 us, the TMA queue hides more of it)
  // shared-memory carve-up in 32-bit words
  int tables, lut, bars, done_ring, zero, def_static;
  int lbufs, lbuf_words;    // per logic warp: state tile | staging | actions (aliased by the props image) | field images
  int dynamic;              // tiles after a logic warp's first come from a global ticket counter (cbx_params.tile_counter)
  int l_stage, l_acts;      // inside a logic buffer (the state tile is at 0)
  int i_scal, i_leak, i_cachem, i_props, i_priv, i_local;  // field images [32 envs][words per env], inside a logic buffer
  int i_fwin, i_fwout;      // live defender binding: the tile's firewall rows [32 envs][6 n bytes] (else -1: static CTA image)
  int def_svc;              // byte offset of the service rows inside the CTA's static defender image
  int slots, slot_words;    // per slot: descriptors [32][desc_words] | header (32 words)
  int s_hdr;
  int wbufs, wbuf_words;    // per encoder warp
  int b_remote, b_conn, b_inf;  // inside an encoder warp's buffer, in bytes (multiples of 16)
  int total_bytes;
};
// slot header words
enum { CBX_SH_ENC_MASK = 0, CBX_SH_TILE = 1 };

// Warp-per-tile kernel for large per-env state (cbx_wide.cuh): nothing big is staged; per warp a private area holds the
// staging words, the encoder descriptors and one image area -- the tile's actions while the game logic runs, then the packed
// field images (property bits + 2-bit privilege codes, credential-cache byte pairs, infected-node bytes), img_stride words
// per env.
#define CBX_WIDE_WARPS 14  // most warps per CTA (the kernel is compiled for 448 threads, one CTA per SM)
struct cbx_wide_plan {
  int enabled;
  int nwarps;                                      // warps per CTA: as many as fit (<= CBX_WIDE_WARPS)
  int dynamic;                                     // tiles after a warp's first come from the global ticket counter
  int warps, warp_words;                           // shared-memory carve-up in 32-bit words
  int w_stage, w_desc, w_img;                      // inside a warp's area
  int img_words, img_stride;                       // packed image words per env; the stride is odd (bank-conflict free columns)
  int nodes_per_chunk;                             // property / privilege images are built for this many nodes at a time
  int total_bytes;
};

struct cbx_params {
  cbx_layout lay;
  cbx_enc_consts enc;
  cbx_smem_plan plan;
  cbx_pipe_plan pipe;
  cbx_wide_plan wide;
  cbx_config cfg;
  int64_t n_envs;
  int64_t n_pad;          // n_envs rounded up to CBX_TILE
  int n_tiles;
  int table_words;        // scenario blob words (multiple of 4; the largest blob of a multi-scenario batch, others are padded)
  int slice_of_kind[3];
  int n_scenarios;        // > 1: cbx_batch_create_multi -- envs grouped by scenario in whole tiles
  int table_stride;       // words between the tables of consecutive scenarios (table_words + padded S + fwx_words)
  int fwx_words;          // firewall extension tables staged behind the initial state (live defender binding; else 0)
  const int32_t* tile_scn;  // [n_tiles] scenario of each tile (NULL for a single scenario)
  const uint32_t* tables; // per scenario: blob followed by the S-word initial state
  uint32_t* state;
  const int32_t* att_actions;
  const int32_t* def_actions;
  int act_i16;            // the two action arrays hold int16 elements (cbx_batch_step_i16 / _host_i16)
  uint8_t* host_results;  // optional mirror of the six result arrays in mapped host memory, cbx_batch_step_host's layout
  const double* scan_u;
  const double* detect_u;
  const uint8_t* reset_mask;  // reset kernel only
  float notify_last_reward;   // CBX_OP_NOTIFY
  cbx_views v;
  int* tile_counter;      // warp-per-tile kernel, dynamic tile order: [0] tickets handed out, [1] CTAs finished (both 0 between
                          // launches: the last CTA of a launch resets them; its launches are serialised)
  int* tickets;           // pipelined kernel, dynamic tile order: THIS launch's ticket counter, zero at launch (a slot of
                          // cbx_batch's ring; never reset in-kernel: with overlapped launches and a grid smaller than the
                          // machine any number of launches can be in flight at once)
  // Overlapped launches of the pipelined kernel (programmatic dependent launch): launch k+1's CTAs start on every SM the
  // moment launch k's CTA has left it and take a tile only after ALL of launch k's writes to that tile have completed --
  // tile_done[t] counts completions of tile t, (1 + encoder warps) per launch (the logic warp and every encoder warp add one
  // when their stores of the tile are complete); a launch with sequence number seq waits for parts * (seq - 1).
  uint32_t* tile_done;
  uint32_t seq;           // sequence number of this launch (1, 2, ...), counted per batch
  int overlap;            // 1: the protocol above is on (cbx_pipe_kernel only)
  int l2_hints;           // cbx_pipe_kernel: bit 0 = state tiles loaded / stored with L2 evict_last, bit 1 = the dense masks
                          // stored with evict_first
  unsigned long long* prof;  // optional: 16 cycle counters accumulated per phase by thread 0 of every CTA (NULL = off)
};

#endif  // CBX_LAYOUT_H_
