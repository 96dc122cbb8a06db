/*
 * tsa_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the reference's template-switch alignment path
 * (lib_tsalign::a_star_aligner + generic_a_star).  Only tests/, bench.py's
 * cpu_baseline / --impl reference legs and __graft_entry__.smoke() may load this
 * library, and only as the checker.  The product (template_switch_aligner_b200)
 * never links, imports or calls anything in oracle/.
 *
 * Parity status: the reference cannot be built here (no cargo/rustc).  The
 * restatement is pinned against the reference's own known answers
 * (tests/test_oracle_kat.py): lib_tsalign/src/tests.rs:38-194 (cost 10),
 * a_star_aligner/tests.rs:10-29 (1D2=2I, cost 9), the compute_cost vectors of
 * alignment_result/alignment/template_switch_specifics.rs:863-1410 and the
 * golden test_files .toml files (rescoring to the recorded cost).
 */
#ifndef TSA_ORACLE_H
#define TSA_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TSA_ORACLE_INF UINT64_MAX

/* Order of the five gap-affine tables (config.rs:35-39). */
enum { TSAO_TAB_PRIMARY = 0, TSAO_TAB_SEC_FWD = 1, TSAO_TAB_SEC_REV = 2, TSAO_TAB_LEFT_FLANK = 3, TSAO_TAB_RIGHT_FLANK = 4 };
/* Order of the six step functions (config.rs:43-48). */
enum { TSAO_FN_RQQR_OFFSET = 0, TSAO_FN_RRQQ_OFFSET = 1, TSAO_FN_LENGTH = 2, TSAO_FN_LENGTH_DIFFERENCE = 3, TSAO_FN_FWD_APG = 4, TSAO_FN_REV_APG = 5 };
/* Order of the base costs (config.rs:52-69). */
enum { TSAO_RRF = 0, TSAO_RQF = 1, TSAO_QRF = 2, TSAO_QQF = 3, TSAO_RRR = 4, TSAO_RQR = 5, TSAO_QRR = 6, TSAO_QQR = 7 };

/* Flattened TemplateSwitchConfig (config.rs:24-49).  Sequences are passed as
 * alphabet indices; `complement` maps index -> index. */
typedef struct {
    int32_t alphabet_size;
    const uint64_t* sub;        /* [5][A*A], row-major [c1*A + c2] (gap_affine.rs:148-157) */
    const uint64_t* open;       /* [5][A] */
    const uint64_t* ext;        /* [5][A] */
    uint64_t base[8];
    int32_t fn_len[6];
    const int64_t* fn_x[6];     /* breakpoints; fn_x[k][0] = INT64_MIN (isize) or 0 (Length) */
    const uint64_t* fn_c[6];
    int64_t left_flank_length;
    int64_t right_flank_length;
    const uint8_t* complement;  /* [A] */
} tsao_config;

/* Alignment op types (alignment_type.rs:11-75), same numbering as the product ABI. */
enum {
    TSAO_OP_PRIMARY_INSERTION = 0, TSAO_OP_PRIMARY_DELETION = 1, TSAO_OP_PRIMARY_SUBSTITUTION = 2, TSAO_OP_PRIMARY_MATCH = 3,
    TSAO_OP_PRIMARY_FLANK_INSERTION = 4, TSAO_OP_PRIMARY_FLANK_DELETION = 5, TSAO_OP_PRIMARY_FLANK_SUBSTITUTION = 6, TSAO_OP_PRIMARY_FLANK_MATCH = 7,
    TSAO_OP_SECONDARY_INSERTION = 8, TSAO_OP_SECONDARY_DELETION = 9, TSAO_OP_SECONDARY_SUBSTITUTION = 10, TSAO_OP_SECONDARY_MATCH = 11,
    TSAO_OP_TS_ENTRANCE = 12, TSAO_OP_TS_EXIT = 13
};

typedef struct {
    int64_t count;      /* RLE multiplicity as the reference emits it (a_star_aligner.rs:100-122) */
    int32_t type;       /* TSAO_OP_* */
    int32_t primary;    /* entrance: 0 = Reference, 1 = Query */
    int32_t secondary;  /* entrance */
    int32_t direction;  /* entrance: 0 = Forward, 1 = Reverse */
    int64_t value;      /* entrance: first_offset; exit: anti_primary_gap */
} tsao_op;

enum { TSAO_FOUND_TARGET = 0, TSAO_EXCEEDED_COST_LIMIT = 1, TSAO_EXCEEDED_MEMORY_LIMIT = 2, TSAO_NO_TARGET = 3 };

typedef struct {
    int32_t result_type;
    uint64_t cost;                 /* FoundTarget cost | cost_limit | max_cost (lib.rs:654-660) */
    uint64_t opened_nodes, closed_nodes, suboptimal_opened_nodes;
    uint64_t ts_total_length;      /* secondary maximisable score of the returned target */
    tsao_op* ops;
    int64_t n_ops;
} tsao_result;

typedef struct {
    int32_t no_ts;                 /* MaxTemplateSwitchCount(0), template_switch_count.rs:41-63 */
    int32_t total_length_maximise; /* MaxTemplateSwitchTotalLength => label-correcting */
    int32_t min_length_lookahead;  /* LookaheadTemplateSwitchMinLengthStrategy */
    int32_t force_label_correcting;
    uint64_t cost_limit;           /* UINT64_MAX = none */
    uint64_t memory_limit;         /* UINT64_MAX = none; bytes */
} tsao_options;

/* Restatement of template_switch_distance_a_star_align's search + backtrack
 * (a_star_aligner.rs:58-161) WITHOUT the post-processing steps. */
int tsao_astar_align(const tsao_config* cfg, const uint8_t* reference, int64_t n, const uint8_t* query, int64_t m,
                     int64_t ref_offset, int64_t ref_limit, int64_t qry_offset, int64_t qry_limit,
                     const tsao_options* opt, tsao_result* out);

/* Restatement of Alignment::compute_cost (template_switch_specifics.rs:591-835),
 * extended to flank ops (charged with the flank tables).  Returns UINT64_MAX on
 * overflow/inf, and sets *ok=0 when the walk leaves the sequences. */
uint64_t tsao_rescore(const tsao_config* cfg, const uint8_t* reference, int64_t n, const uint8_t* query, int64_t m,
                      int64_t ref_offset, int64_t qry_offset, const tsao_op* ops, int64_t n_ops,
                      int64_t* end_ref, int64_t* end_qry, int32_t* ok);

/* Same walk; as_searched != 0 charges forward entrances the way the search does (oc(0) + oc(o) - oc(+-1),
 * identifier.rs:290-319 + context.rs:392-462) instead of compute_cost's oc(o).  Identical for flat offset costs. */
uint64_t tsao_rescore_mode(const tsao_config* cfg, const uint8_t* reference, int64_t n, const uint8_t* query, int64_t m,
                           int64_t ref_offset, int64_t qry_offset, const tsao_op* ops, int64_t n_ops,
                           int64_t* end_ref, int64_t* end_qry, int32_t* ok, int32_t as_searched);

/* Scalar layered-DP statement of the same shortest-path problem;
 * the thing the CUDA kernels are diffed against at sizes A* cannot reach. */
int tsao_dp_align(const tsao_config* cfg, const uint8_t* reference, int64_t n, const uint8_t* query, int64_t m,
                  int64_t ref_offset, int64_t ref_limit, int64_t qry_offset, int64_t qry_limit,
                  const tsao_options* opt, tsao_result* out);

void tsao_result_free(tsao_result* r);

#ifdef __cplusplus
}
#endif
#endif
