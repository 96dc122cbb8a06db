/*
 * tsalign_b200.h -- C ABI of the B200-native template-switch aligner (libtsalign_b200.so).
 *
 * Drop-in boundary for the alignment hot path of sebschmi/template-switch-aligner.  The reference has no
 * FFI; its seam is the Rust generic function
 *     lib_tsalign::a_star_aligner::template_switch_distance_a_star_align   (lib_tsalign/src/a_star_aligner.rs:179-254)
 * called by the `tsalign align` CLI (tsalign/src/align/template_switch_distance_type_selectors.rs:389-450) and by
 * the run-time facade lib_tsalign::a_star_aligner::configurable_a_star_align::Aligner::align
 * (configurable_a_star_align.rs:214-236) used by the Python binding (python_bindings/src/lib.rs:97-142).
 * Each entry point below names the reference interface it replaces.  INTEGRATION.md shows the Rust `extern "C"`
 * block and the call-site change a maintainer would add.
 *
 * Ownership: the caller owns every input for the duration of a call; results are allocated by the library and
 * released with tsa_results_free.  No function throws or aborts across the ABI on bad input: integer status codes
 * plus a message buffer.  There is no CPU path: without a CUDA device every compute entry returns
 * TSA_ERR_NO_DEVICE.
 */
#ifndef TSALIGN_B200_H
#define TSALIGN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes --------------------------------------------------------------------------------------- */
enum {
    TSA_OK = 0,
    TSA_ERR_NO_DEVICE = 1,          /* no CUDA device / invalid device index */
    TSA_ERR_CONFIG_PARSE = 2,       /* lib_tsalign/src/error.rs: Error::Parser* */
    TSA_ERR_NOT_V_SHAPED_RQQR = 3,  /* error.rs: RQQROffsetCostsNotVShaped (config.rs:72-85) */
    TSA_ERR_NOT_V_SHAPED_RRQQ = 4,  /* error.rs: RRQQOffsetCostsNotVShaped */
    TSA_ERR_NOT_V_SHAPED_LENDIFF = 5, /* error.rs: LengthDifferenceCostsNotVShaped */
    TSA_ERR_ALPHABET = 6,           /* unknown alphabet id */
    TSA_ERR_INVALID_CHAR = 7,       /* a sequence character is not in the alphabet (align.rs:389-405; the facade panics) */
    TSA_ERR_INVALID_RANGE = 8,      /* offset > limit or limit > length */
    TSA_ERR_UNSUPPORTED = 9,        /* cost model outside the kernels' limits (see message) */
    TSA_ERR_ARGUMENT = 10,
    TSA_ERR_INTERNAL = 11           /* a consistency check of the library failed (never a wrong result) */
};

/* ---- alphabets: tsalign/src/align.rs:80-81,288-295 (-a/--alphabet) -------------------------------------- */
enum { TSA_ALPHABET_DNA = 0, TSA_ALPHABET_DNA_N = 1, TSA_ALPHABET_RNA = 2, TSA_ALPHABET_RNA_N = 3, TSA_ALPHABET_DNA_IUPAC = 4, TSA_ALPHABET_RNA_IUPAC = 5 };

/* ---- cost model: lib_tsalign::config::TemplateSwitchConfig (config.rs:24-49) ---------------------------- */
typedef struct tsa_config tsa_config;

/* TemplateSwitchConfig::read_plain / from_str (config/io.rs:21-31,265-275): parse a config.tsa text. */
tsa_config* tsa_config_parse(const char* text, size_t len, int alphabet, int* status, char* err, size_t errcap);
/* TemplateSwitchConfig::default() (config.rs:219-303): the cost model the Python binding uses when none is given. */
tsa_config* tsa_config_default(int alphabet);
/* Display for TemplateSwitchConfig (config/io.rs:223-263): config.tsa text; returns the needed size incl. NUL. */
size_t tsa_config_write(const tsa_config* cfg, char* out, size_t cap);
void tsa_config_free(tsa_config* cfg);
int tsa_config_alphabet(const tsa_config* cfg);

/* ---- options: the result-relevant part of the strategy selection (align.rs:104-223) --------------------- */
typedef struct {
    int32_t no_ts;                 /* --no-ts (MaxTemplateSwitchCount(0)) */
    int32_t device;                /* CUDA device index */
    uint64_t cost_limit;           /* --cost-limit, UINT64_MAX = none: optimal cost > limit -> ExceededCostLimit */
    uint64_t memory_limit;         /* --memory-limit in bytes, UINT64_MAX = none: bounds the resident HBM chunk */
    int32_t max_template_switches; /* 0 = default (64) */
    int32_t first_threshold;       /* 0 = default (12): first pruning threshold of the iterative deepening; tuning only, never changes results */
    int32_t no_traceback;          /* 1 = costs only (ops == NULL) */
    int32_t reserved;
    int32_t postprocess;           /* TSA_POST_* bits: what template_switch_distance_a_star_align does after the search (a_star_aligner.rs:238-253) */
    int32_t flags;                 /* TSA_FLAG_* bits */
    int32_t total_length_strategy; /* --ts-total-length-strategy (align.rs:112-118): 0 = maximise (the reference's default), 1 = none */
    int32_t descendant_strategy;   /* --ts-descendant-strategy (strategies/descendant.rs:22-104): 0 = any, 1 = allow-only-all-equal */
    int32_t force_label_correcting;/* --force-label-correcting (align.rs:119-122): accepted; a dense fill is exact either way */
    int32_t reserved2;
} tsa_options;

/* tsa_options.flags */
enum {
    TSA_FLAG_KEEP_FLANK_RUNS = 1   /* do not merge flank and non-flank variants of a primary operation into one run (the reference merges them,
                                      alignment_type.rs:101-121, which loses where a flank begins): lets a checker rescore flank alignments */
};

/* tsa_options.postprocess (host work per found alignment, O(length^2) like the reference's) */
enum {
    TSA_POST_EXTEND_BEYOND_RANGE = 1, /* extend_beyond_range_without_increasing_cost (alignment_result.rs:247-395); `tsalign align` default */
    TSA_POST_EQUAL_COST_RANGES = 2    /* compute_ts_equal_cost_ranges (alignment_result.rs:398-573); the reference always does this */
};

/* ---- one alignment problem: the arguments of Aligner::align (configurable_a_star_align.rs:214-236) ------ */
typedef struct {
    const char* reference; size_t reference_len;   /* ASCII, already upper-cased / skip characters removed */
    const char* query; size_t query_len;
    int64_t reference_offset, reference_limit;     /* AlignmentRange; limit < 0 means "end of sequence" */
    int64_t query_offset, query_limit;
} tsa_pair;

/* ---- result: generic_a_star::AStarResult (generic_a_star/src/lib.rs:164-187) + alignment ---------------- */
enum { TSA_FOUND_TARGET = 0, TSA_EXCEEDED_COST_LIMIT = 1, TSA_EXCEEDED_MEMORY_LIMIT = 2, TSA_NO_TARGET = 3 };

/* AlignmentType (template_switch_distance/alignment_type.rs:11-75) */
enum {
    TSA_OP_PRIMARY_INSERTION = 0, TSA_OP_PRIMARY_DELETION = 1, TSA_OP_PRIMARY_SUBSTITUTION = 2, TSA_OP_PRIMARY_MATCH = 3,
    TSA_OP_PRIMARY_FLANK_INSERTION = 4, TSA_OP_PRIMARY_FLANK_DELETION = 5, TSA_OP_PRIMARY_FLANK_SUBSTITUTION = 6, TSA_OP_PRIMARY_FLANK_MATCH = 7,
    TSA_OP_SECONDARY_INSERTION = 8, TSA_OP_SECONDARY_DELETION = 9, TSA_OP_SECONDARY_SUBSTITUTION = 10, TSA_OP_SECONDARY_MATCH = 11,
    TSA_OP_TS_ENTRANCE = 12, TSA_OP_TS_EXIT = 13
};

typedef struct {
    int64_t count;       /* run length as the reference emits it (a_star_aligner.rs:100-122) */
    int32_t type;        /* TSA_OP_* */
    int32_t primary;     /* entrance: 0 = Reference, 1 = Query */
    int32_t secondary;   /* entrance */
    int32_t direction;   /* entrance: 0 = Forward, 1 = Reverse */
    int64_t value;       /* entrance: first_offset; exit: anti_primary_gap */
    int8_t min_start, max_start, min_end, max_end; /* entrance: EqualCostRange (alignment_type.rs:77-99); (1,-1,1,-1) = invalid / not computed */
    int32_t reserved;
} tsa_op;

typedef struct {
    int32_t status;            /* TSA_OK or a per-pair TSA_ERR_* */
    int32_t result_type;       /* TSA_FOUND_TARGET ... */
    uint64_t cost;             /* FoundTarget: cost; ExceededCostLimit: the limit */
    int32_t template_switches; /* number of template switches on the returned path */
    int32_t reserved;
    tsa_op* ops;               /* run-length encoded alignment (NULL when no alignment was requested / found) */
    size_t n_ops;
    double duration_seconds;   /* share of the batch's wall time */
    char message[96];          /* human-readable reason when status != TSA_OK */
    int64_t reference_offset, reference_limit; /* the alignment range the ops span (wider than the input range after TSA_POST_EXTEND_BEYOND_RANGE) */
    int64_t query_offset, query_limit;
} tsa_result;

/* Batch form of template_switch_distance_a_star_align: aligns n independent pairs on one GPU. */
int tsa_align_batch(const tsa_config* cfg, const tsa_options* opt, const tsa_pair* pairs, size_t n, tsa_result* out, char* err, size_t errcap);
void tsa_results_free(tsa_result* results, size_t n);

/* ---- staged form (what bench.py times): inputs resident in HBM, run any number of times ------------------ */
typedef struct tsa_batch tsa_batch;
tsa_batch* tsa_batch_create(const tsa_config* cfg, const tsa_options* opt, const tsa_pair* pairs, size_t n, int* status, char* err, size_t errcap);
int tsa_batch_run(tsa_batch* batch);                              /* kernels only; synchronous */
int tsa_batch_fetch(tsa_batch* batch, tsa_result* out);           /* D2H + result assembly */
void tsa_batch_stats(const tsa_batch* batch, int64_t* launches, int64_t* jump_launches, int64_t* fill_launches, int32_t* layers, int64_t* h2d_bytes, int64_t* d2h_bytes);
/* device time (CUDA events on the engine's stream) of the two kernel families during the last tsa_batch_run */
void tsa_batch_timing(const tsa_batch* batch, double* jump_ms, double* fill_ms);
/* work done by the jump kernel during the last tsa_batch_run: chain pairs started / not pruned, rows filled, rows jumped */
void tsa_batch_work(const tsa_batch* batch, int64_t* chains_started, int64_t* chains_run, int64_t* rows_filled, int64_t* rows_jumped);
void tsa_batch_free(tsa_batch* batch);

/* ---- one very long pair without template switches, column-banded over several GPUs (BASELINE config 5) ----------------------
 * What `tsalign align --no-ts --memory-limit B` (tsalign/src/align.rs:57-223) is asked for on a pair whose search does not fit the
 * limit: the reference's A* gives up with AStarResult::ExceededMemoryLimit (generic_a_star/src/lib.rs:332-335,380-389).  Here the
 * limit (tsa_options.memory_limit, bytes per device) bounds what is resident -- checkpoint rows, boundary columns and the codes of
 * one tile instead of a code matrix -- and the alignment is still produced; TSA_EXCEEDED_MEMORY_LIMIT is returned only when not even
 * the coarsest checkpoint spacing fits.  The query columns are cut into bands, one per device; the last column of a band streams
 * into the next device's memory with plain 8-byte stores over NVLink (peer access), no collective. */
typedef struct {
    double forward_ms, trace_ms;       /* device time of the band's forward launch; host wall time of its share of the traceback */
    int64_t tiles, tile_cells;         /* tiles recomputed with codes by the traceback, and their cells */
    int64_t boundary_bytes_out;        /* bytes this band stored into the next device's memory (8 per row) */
    int64_t resident_bytes;            /* device memory the band held */
    int32_t interval, group;           /* rows between checkpoint rows / 256-column strips between kept boundary columns */
    int64_t speculated_tiles, speculated_used; /* tiles around the diagonal recomputed ahead of the walk in one launch, and how many the walk used */
    double speculate_ms;               /* device time of that launch */
} tsa_long_stats;
/* All bands driven by this process (one host thread; devices[k] = CUDA device of band k, neighbours need peer access).
 * interval / group: 0 = chosen by the library.  stats: n_devices entries or NULL.  opt->no_ts must be set. */
int tsa_align_long(const tsa_config* cfg, const tsa_options* opt, const tsa_pair* pair, const int32_t* devices, int32_t n_devices,
                   int32_t interval, int32_t group, tsa_result* out, tsa_long_stats* stats, char* err, size_t errcap);

/* The same, one band per process (torchrun: rank r drives device opt->device).  Protocol: every rank creates its band, exports the
 * handle of its incoming boundary buffer, connects to the handle of rank + 1 (ranks exchange the 64 bytes however they like),
 * a barrier, then every rank calls tsa_long_forward concurrently.  The traceback starts on the last rank with
 * {nn, mm, 0, 1, cost, 0} and is passed to tsa_long_owner(column) whenever a walk returns with status 0. */
typedef struct tsa_long tsa_long;
typedef struct { int32_t i, j, g, need; int64_t cost; int32_t status, reserved; } tsa_long_walk_state;   /* status: 0 handed to the left, 1 root reached, < 0 error */
tsa_long* tsa_long_create(const tsa_config* cfg, const tsa_options* opt, const tsa_pair* pair, int32_t rank, int32_t world,
                          int32_t interval, int32_t group, int* status, char* err, size_t errcap);
int tsa_long_ipc_export(const tsa_long* band, void* handle64);            /* cudaIpcGetMemHandle of the incoming boundary buffer */
int tsa_long_ipc_connect(tsa_long* band, const void* handle64);           /* maps the next rank's buffer (cudaIpcOpenMemHandle) */
int tsa_long_forward(tsa_long* band);
int tsa_long_cost(const tsa_long* band, uint64_t* cost, int32_t* result_type);   /* last rank only */
int tsa_long_owner(const tsa_long* band, int64_t column);                 /* rank whose band holds this column */
int tsa_long_walk(tsa_long* band, tsa_long_walk_state* state, uint8_t* ops, size_t cap, size_t* n_ops);   /* unit ops, walk order */
/* unit ops of all bands in walk order (last band first) -> run-length encoded result with the range of the pair */
int tsa_long_result(const tsa_long* band, uint64_t cost, const uint8_t* ops_walk_order, size_t n_ops, tsa_result* out);
void tsa_long_get_stats(const tsa_long* band, tsa_long_stats* stats);
void tsa_long_dims(const tsa_long* band, int64_t* rows, int64_t* columns);
void tsa_long_free(tsa_long* band);

/* Integer roofline probe: measured issue rate (lanes/s) of the DPX add-min instructions on all SMs. */
int tsa_measure_addmin_peak(int device, double* s16x2_lane_ops_per_s, double* s32_lane_ops_per_s);

/* Host-only entry of the post-processing (no device needed): ops in / out in place, `cap` = capacity of the ops array
 * (the run-length encoding can grow by a few entries), *n_ops updated; range in / out.  Returns TSA_OK, or TSA_ERR_ARGUMENT when
 * the capacity is too small or a character is not in the alphabet.  Used by the CLI's tests against the reference's result files. */
int tsa_postprocess(const tsa_config* cfg, const tsa_pair* pair, int32_t postprocess, tsa_op* ops, size_t* n_ops, size_t cap,
                    int64_t* reference_offset, int64_t* reference_limit, int64_t* query_offset, int64_t* query_limit, uint64_t* cost);

/* The four single-step moves compute_ts_equal_cost_ranges is built from (Alignment::move_template_switch_{start,end}_{backwards,
 * forwards}, alignment/template_switch_specifics.rs:30-589), host-only, exposed so that the reference's own unit vectors for them
 * (ibid. :1251-1410) can be replayed: which = 0 start backwards, 1 start forwards, 2 end backwards, 3 end forwards; *compact_index
 * points at the entrance and follows it.  Returns 1 if the move was possible, 0 if not, < 0 (-TSA_ERR_*) on bad arguments. */
int tsa_post_move(const tsa_config* cfg, const tsa_pair* pair, int which, tsa_op* ops, size_t* n_ops, size_t cap,
                  int64_t reference_offset, int64_t query_offset, size_t* compact_index, uint64_t* cost);

/* ---- misc ----------------------------------------------------------------------------------------------- */
int tsa_device_count(void);
const char* tsa_version(void);

#ifdef __cplusplus
}
#endif
#endif
