/* include/bpe_b200.h — C ABI of the B200-native BPE engine (libbpe_b200.so).
 *
 * Drop-in boundary for dbtreasure/zig-bpe's BasicTokenizer hot path. The reference has no
 * FFI of its own (it is one Zig struct, src/basic_tokenizer.zig:52-349); this header is the
 * thin layer its Zig host code binds with `extern fn` so that
 *   BasicTokenizer.train   (src/basic_tokenizer.zig:140-153, hot loop :172-306)
 *   BasicTokenizer.encode  (src/basic_tokenizer.zig:71-88)
 *   BasicTokenizer.decode  (src/basic_tokenizer.zig:90-138)
 * run on the GPU while the struct, allocator discipline, file I/O and merge (de)serialisation
 * (src/basic_tokenizer.zig:319-348) stay on the host side.
 *
 * Conventions: plain pointers and sizes only; every host output buffer is allocated by the
 * caller from the bounds documented per call; the library never returns memory it allocated;
 * all device memory is owned by the bpe_ctx. A bpe_ctx is not thread-safe; distinct contexts
 * are independent. There is no CPU fallback: every entry point fails with BPE_ERR_CUDA when
 * no usable sm_100-class device is present.
 */
#ifndef BPE_B200_H
#define BPE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bpe_ctx bpe_ctx;

/* = `Merge{ pair: CharPair{first,second}, new_token }` (src/basic_tokenizer.zig:12-15,40-43), 6 bytes */
typedef struct bpe_merge_t {
    uint16_t first;
    uint16_t second;
    uint16_t new_token;
} bpe_merge_t;

/* Fills the reference's TimeStats buckets (src/utils/time_statistics.zig:4-34) from CUDA-event
 * timings: pair counting -> just_count_pairs, argmax + tie-break -> sort_pairs,
 * merge + compaction -> replace_pair; generate_pairs is always 0 (pairs are formed in registers). */
typedef struct bpe_stats_t {
    double sort_pairs_ms;
    double replace_pair_ms;
    double generate_pairs_ms;
    double just_count_pairs_ms;
    uint64_t sort_pairs_calls;
    uint64_t replace_pair_calls;
    uint64_t generate_pairs_calls;
    uint64_t just_count_pairs_calls;
    double total_ms;          /* whole call, host clock */
    double device_ms;         /* resident-in-HBM -> merge list on host (CUDA events) */
    uint64_t scanned_slots;   /* sum over merge steps of resident token slots scanned (roofline numerator / 2 B) */
    uint64_t kernel_launches; /* kernels launched by this call */
    uint64_t tie_steps;       /* steps whose top count was shared by >= 2 pairs */
    uint64_t tie_slow_steps;  /* of those, resolved by the full table replay */
    uint64_t compactions;
    /* per-kernel-class device time, filled when option "profile" = 1 (CUDA events recorded on
     * the context's stream between launches, resolved after the run; no extra synchronisation):
     * [0] load+initial count  [1] argmax+ties  [2] tie occupancy kernels  [3] table replay
     * [4] halo  [5] merge  [6] apply deltas  [7] compaction  [8] table rebuild / zcnt rebuild
     * [9] host gap (status read-back until the next launch); encode with "encode_fuse": kernel_calls[9] / kernel_calls[8] =
     * fused level groups that went through / whose halos did not cover their levels (redone level by level)  [10] profile 3: slots scanned by the sampled
     * merge launches; encode: kernel_calls[10] = the encoder that produced the ids (0 passes, 1 segment kernel,
     * 2 tile kernel)  [11] encode: time of the tile / segment kernel; kernel_calls[11] = 1 when it produced the
     * result, 2 when it gave up (no common token at a seam, ...) and another encoder ran instead, 0 when not tried */
    double kernel_ms[12];
    uint64_t kernel_calls[12];
    uint64_t aeqb_steps;      /* train: steps whose merge has first == second (they take the run-chaining halo pass) */
} bpe_stats_t;

enum {
    BPE_OK = 0,
    BPE_ERR_INVALID_VOCAB = 1, /* TrainError.InvalidVocabSize (src/basic_tokenizer.zig:147-149) */
    BPE_ERR_OOM = 2,           /* TrainError.OutOfMemory: host or device allocation, or output cap too small */
    BPE_ERR_INVALID_TOKEN = 3, /* error.InvalidToken (src/basic_tokenizer.zig:101,125,135) */
    BPE_ERR_CUDA = 4,          /* CUDA / NCCL failure, or no device (no reference equivalent) */
    BPE_ERR_INVALID_ARG = 5,   /* null pointer, bad option name, ... */
    BPE_ERR_INTERNAL = 7       /* self-check failed (verify mode) */
};

/* ---- context ------------------------------------------------------------------------- */
/* One context = one GPU (CUDA device ordinal `device`) and one stream. */
int bpe_ctx_create(bpe_ctx** out, int device);
/* Multi-GPU: one process (or thread) per GPU; `rank`-th contiguous shard of the corpus goes to
 * rank `rank`. `nccl_unique_id` is the 128-byte ncclUniqueId obtained from
 * bpe_nccl_unique_id() on rank 0 and broadcast by the caller (torch.distributed, MPI, ...). */
int bpe_ctx_create_dist(bpe_ctx** out, int device, int rank, int world, const void* nccl_unique_id);
int bpe_nccl_unique_id(void* out128);
void bpe_ctx_destroy(bpe_ctx* ctx);
/* Message for the last non-OK status returned on this context ("" if none). ctx may be NULL
 * to get the message of the last failed bpe_ctx_create*. */
const char* bpe_last_error(const bpe_ctx* ctx);

/* Options (all default 0 unless noted):
 *   "verify_recount"      1: after every merge step recount all pairs from the token sequence
 *                            and compare with the incrementally maintained table (debug)
 *   "force_slow_tiebreak" 1: resolve every tie step with the full table replay
 *   "check_tiebreak"      1: on fast-path tie steps also run the replay and compare (debug)
 *   "compact_pct"         train: live/slots percentage below which the sequence is compacted between batches (default 93;
 *                         C3 on one B200: 1,257 ms per training with 85, 1,230 with 90, 1,216 with 93, 1,214 with 96)
 *   "encode_compact_pct"  the same for the level passes of encode (default 85)
 *   "table_log2"          log2 of the initial pair-table capacity (default: sized from n)
 *   "max_steps"           stop training after this many merges (0 = no limit)
 *   "encode_impl"         0 (default): the tile-resident kernel (4) for every list a trained tokenizer can write; when it
 *                         does not apply or gives up, the choice between the segment-resident kernel (3) and the level
 *                         passes (2) is made by cost: lists whose level schedule has more than "encode_seg_min_steps" steps
 *                         (default 450), and inputs of 4 GiB or more, go to the segment kernel. 4: tile-resident kernel — a
 *                         CTA keeps a window of the text ("encode_tile" bytes, default 7936, maximum 8192, + 2 x 64 bytes of
 *                         margin) in shared memory through all merges, one launch, the text is read once; neighbouring
 *                         windows are stitched at a token they share (BPE_ERR_INTERNAL if it does not apply).
 *                         1: one pass per merge, in list order. 2: level-
 *                         scheduled passes — merges that commute (no shared produced token, no token that is second of
 *                         one pair and first of another) share one pass. 3: segment-resident kernel (every thread encodes
 *                         a window of the text in shared memory, neighbouring windows are stitched at a token they share;
 *                         regular lists only; BPE_ERR_INTERNAL if it does not apply). All give the reference's result;
 *                         0 falls back from 3 to 2 by itself (irregular list, id 65535, windows without a common token).
 *   "encode_geom"         window geometry of the segment kernel: 4 (default) 64-byte cores + 2 x 32 bytes of margin, 64
 *                         threads per CTA; 0: same, 128 threads; 1: 128 + 2 x 32; 2 / 5: 96 + 2 x 32 with 128 / 64 threads;
 *                         3: 32 + 2 x 12 (tests)
 *   "encode_grid"         CTAs per SM of a level pass (default 24, several waves; they take the tiles round-robin; C3: 84.1 ms with 6, 80.1 ms with 24); 0: one CTA per
 *                         tile; negative: absolute CTA count (tests).
 *   "encode_filter"       candidate filter of a level pass. 2 (default): a 65,536-bit Bloom filter over the level's PAIRS, probed
 *                         with the pair a slot forms with its successor slot, so that practically only real occurrences reach
 *                         the hash look-up; 1: one role byte per id and an in-register successor filter (used when every id of
 *                         the list is below 16384); 0: 1-bit first-component map over all ids.
 *   "merge_impl"          0 (default): register-streaming merge kernel, one CTA per tile;
 *                         1: merge pass fed by a TMA ring (cp.async.bulk + mbarrier, persistent CTAs) —
 *                            measured slower on B200 for this access pattern, kept for comparison
 *                         2: candidate-scan path (single GPU, experimental, currently slower): one barrier-free
 *                            streaming scan queues the A's that can start an occurrence, a resolve + a write kernel
 *                            finish the step (no tiles / halos)
 *   "merge_filter"        1: the merge pass rules out, in registers, every A whose next slot holds a live token other than B;
 *                         a tile goes through the staged path only if a candidate is left. 0 (default): every tile that
 *                         holds an A is staged (measured 14 % faster on C3: early steps stage nearly every tile anyway and
 *                         the filter's instructions sit on the streaming path).
 *   "merge_direct"        K (default 3): a train step whose pair occurs fewer than K times per tile on average takes the
 *                         queue-less staged path (every thread looks its own A's up; no candidate masks, no
 *                         shared-memory queue); 0: always the queued path (round 2's first profile: 45 % of the merge
 *                         pass's instructions were candidate masks and queue traffic on steps with ~3 occurrences per tile)
 *   "encode_fuse"         K (2..8; default 0 = off): level passes of encode in groups of up to K consecutive levels per
 *                         residency of a tile (bpe_groupenc.cuh): one role look-up per slot per group, per-level queues,
 *                         30 live tokens of halo per side with a per-tile check that they covered the group (else the
 *                         group is redone level by level). Single GPU, u16 ids below 16,384. Written after the round's last
 *                         GPU lease: equal to the oracle under the CPU emulation (tests, tools/fuzz_emul_encode.py), not
 *                         yet run on a GPU, hence off.
 *   "merge_loop"          T (default 6): the train loop's merge pass runs ceil(tiles / T) CTAs, each takes T tiles (strided by
 *                         the grid size); the next tile's vectors are prefetched into registers and its halo by cp.async
 *                         while the staged tile is worked on, the control block is read once and the delta bins are
 *                         flushed once per CTA. 0: one CTA per tile (C3 on one B200: 1,417 ms per training against
 *                         1,253 ms with T = 6; T = 2 / 4 / 8 / 12: 1,349 / 1,268 / 1,256 / 1,264 ms).
 *   "merge_pairfilter"    1 (default): on the queue-less path a vector's A's are reduced, in registers, to those whose next slot
 *                         holds B or a hole (packed 16-bit minimum) before anything is looked up; 0: every A is looked up
 *   "merge_prestage"      1 (default): the train loop's merge pass copies every tile into shared memory before the barrier
 *                         that tells whether the tile holds an A at all, so a staged tile waits at one barrier, not two;
 *                         0: stage only tiles that hold an A
 *   "batch_steps"         merge steps enqueued between two read-backs of the control block (default 16)
 *   "stream_chunk_mb"     bpe_encode / host buffers: inputs of two chunks or more (default chunk 128 MiB) stream through the
 *                         GPU in chunks — copies overlap the tile-resident encoder, device memory stays at 8 bytes per chunk
 *                         byte ("stream_chunk_bytes" sets the chunk size in bytes, for tests)
 *   "count_limit_log2"    tests: lowers the limit (2^32 - 1) above which a pair count is refused
 *   "fuse_halo"           1 (default): the apply kernel of a step also gathers the tile halos of the next step (two launches
 *                         per merge step: merge, apply); steps whose merge has first == second take the stand-alone
 *                         run-chaining halo pass. 0: a halo launch before every merge pass.
 *   "xchg_impl"           multi-GPU per-step exchange. 0 (default): peer-memory mailboxes over NVLink (each rank
 *                         writes its deltas into every peer's mailbox and raises a flag; falls back to 1 when
 *                         peer access is unavailable); 1: NCCL all-reduce
 *   "time_phases"         1: fill the reference's TimeStats buckets (CUDA events recorded between the launches of the
 *                         loop, resolved after the run: two event records per launch, no extra synchronisation)
 *   "pdl"                 1 (default): the merge pass and the apply kernel of the train loop are launched as programmatic
 *                         dependents of each other (the next kernel's CTAs start and prefetch while the previous one drains)
 *   "cache_max_mb"        device memory the context keeps cached between calls (default 8192; the rest is freed when an
 *                         API call returns; 0 = keep nothing)
 *   "profile"             1: fill bpe_stats_t.kernel_ms / kernel_calls for every kernel class;
 *                         2: only the merge kernel (two event records per merge step)
 *                         3: the merge kernel of every 32nd step only; kernel_ms[10] then holds the token slots
 *                            those sampled launches scanned (lowest overhead; what bench.py uses)
 */
int bpe_ctx_set_option(bpe_ctx* ctx, const char* name, long value);

/* ---- train (src/basic_tokenizer.zig:140-153; loop :172-205) -------------------------------
 * text[n] bytes on the host. vocab_size as in the reference (u16; < 256 -> BPE_ERR_INVALID_VOCAB).
 * out_merges: capacity vocab_size-256 entries. out_counts (nullable): count of the winning pair
 * per merge, what the reference prints in verbose mode (:308-317). *out_n = merges learned
 * (fewer than vocab_size-256 when the sequence runs out of pairs, :188-191).
 * n == 0 or 1: zero merges (the reference underflows at :246 for n == 0; defined here as
 * "no pairs"). In a dist context every rank passes its own shard and receives the same merges. */
int bpe_train(bpe_ctx* ctx, const uint8_t* text, size_t n, uint16_t vocab_size,
              bpe_merge_t* out_merges, uint64_t* out_counts, size_t* out_n, bpe_stats_t* stats);
/* Same, text already resident in device memory (device pointer on ctx's device). */
int bpe_train_device(bpe_ctx* ctx, const uint8_t* d_text, size_t n, uint16_t vocab_size,
                     bpe_merge_t* out_merges, uint64_t* out_counts, size_t* out_n, bpe_stats_t* stats);

/* ---- encode (src/basic_tokenizer.zig:71-88) ----------------------------------------------
 * Applies merges[0..m) in list order, one left-to-right greedy pass each, exactly as the
 * reference does. out: capacity n ids. In a dist context every rank passes its contiguous shard
 * and receives the ids of its shard; the shards exchange their end tokens before every pass, so the
 * concatenation of the ranks' outputs equals the single-GPU encoding of the whole text (a token that
 * straddles two shards is emitted by the left one). */
int bpe_encode(bpe_ctx* ctx, const uint8_t* text, size_t n, const bpe_merge_t* merges, size_t m,
               uint16_t* out, size_t* out_n, bpe_stats_t* stats);
/* d_text and d_out (capacity n ids) are device pointers. */
int bpe_encode_device(bpe_ctx* ctx, const uint8_t* d_text, size_t n, const bpe_merge_t* merges, size_t m,
                      uint16_t* d_out, size_t* out_n, bpe_stats_t* stats);

/* ---- decode (src/basic_tokenizer.zig:90-138) ---------------------------------------------
 * ids < 256 are bytes; other ids expand through the first merge whose new_token matches
 * (findMerge :109-116); unknown id -> BPE_ERR_INVALID_TOKEN. bpe_decode_size returns the byte
 * length so the caller can allocate; bpe_decode writes at most cap bytes (BPE_ERR_OOM and the
 * needed size in *out_n if cap is too small). */
int bpe_decode_size(bpe_ctx* ctx, const uint16_t* toks, size_t n, const bpe_merge_t* merges, size_t m,
                    size_t* out_n);
int bpe_decode(bpe_ctx* ctx, const uint16_t* toks, size_t n, const bpe_merge_t* merges, size_t m,
               uint8_t* out, size_t cap, size_t* out_n, bpe_stats_t* stats);
/* d_toks and d_out are device pointers. */
int bpe_decode_device(bpe_ctx* ctx, const uint16_t* d_toks, size_t n, const bpe_merge_t* merges, size_t m,
                      uint8_t* d_out, size_t cap, size_t* out_n, bpe_stats_t* stats);

/* Library / build identification, e.g. "bpe_b200 0.1 sm_100a". */
const char* bpe_version(void);

#ifdef __cplusplus
}
#endif
#endif /* BPE_B200_H */
